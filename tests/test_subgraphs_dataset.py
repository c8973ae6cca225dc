"""CPU: the episode sampler (`Subgraphs`), the on-disk formats and the train.py CLI surface against the
UNMODIFIED reference classes (subgraph_data_processing.py:14-412, train.py:153-177) run on the same
dataset directory with the same seeds.  The reference side needs /root/reference (build container only)."""
import argparse
import os
import pickle
import random

import numpy as np
import pytest
import torch

from gmeta_b200 import data_io
from gmeta_b200.subgraph_data_processing import Subgraphs, collate
from oracle import ref_loader
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(ds):
    return argparse.Namespace(sample_nodes=ds.sample_nodes, link_pred_mode='True' if ds.link_pred else 'False',
                              task_setup=ds.task_setup)


def _write(tmp_path, kind, with_dgl):
    ds = H.tiny_dataset(kind)
    dgl = ref_loader.shim_dgl() if with_dgl else None
    root = data_io.write_synthetic_dataset(str(tmp_path / kind), ds, np.random.default_rng(5), dgl_module=dgl)
    return ds, root


def _seed(s):
    np.random.seed(s)
    random.seed(s)
    torch.manual_seed(s)


@pytest.mark.parametrize("kind", ['disjoint', 'shared', 'link', 'deep'])
def test_dataset_directory_round_trip(tmp_path, kind):
    ds, root = _write(tmp_path, kind, False)
    graphs = data_io.load_graphs(root)
    assert len(graphs) == len(ds.graphs)
    for a, b in zip(graphs, ds.graphs):
        assert a.n == b.n and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    feats = data_io.load_features(root)
    assert len(feats) == len(ds.feats) and all(np.array_equal(a, b) for a, b in zip(feats, ds.feats))
    info = data_io.load_labels(root)
    rows = data_io.read_item_csv(os.path.join(root, "train.csv"))
    assert rows and all(info[name] == int(lab) for name, lab in rows)
    if ds.link_pred:
        spt = data_io.read_item_csv(os.path.join(root, "train_spt.csv"))
        qry = data_io.read_item_csv(os.path.join(root, "train_qry.csv"))
        assert sorted(spt + qry) == sorted(rows)
    # one episode in the reference's 10-tuple layout
    _seed(3)
    db = Subgraphs(root, 'train', info, n_way=ds.n_way, k_shot=ds.k_spt, k_query=ds.k_qry, batchsz=4, args=_args(ds),
                   adjs=graphs, h=ds.h)
    assert len(db) == 4
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = db[0]
    S = len(xs.batch_num_nodes)
    assert ys.shape[0] == S == len(ns) == len(gs) and yq.shape[0] == len(xq.batch_num_nodes)
    assert ys.dtype == torch.int64 and cs.dtype == torch.int64
    assert tuple(cs.shape) == ((S, 2) if ds.link_pred else (S,))
    for k in range(S):                                   # centre index points at the item's node
        g, i = gs[k], int(db.support_x_batch[0][k // ds.k_spt][k % ds.k_spt].split('_')[1])
        c = int(cs[k][0]) if ds.link_pred else int(cs[k])
        assert ns[k][c] == i and len(ns[k]) == xs.batch_num_nodes[k]
    if ds.task_setup == 'Disjoint':
        assert sorted(set(ys.tolist())) == list(range(ds.n_way))      # relabelled 0..n_way-1 (:390-397)
    batch = collate([db[0], db[1]])
    assert len(batch) == 10 and all(len(b) == 2 for b in batch)


def test_unreadable_dgl_pickle_is_reported(tmp_path):
    root = tmp_path / "d"
    root.mkdir()
    with pytest.raises(FileNotFoundError):
        data_io.load_graphs(str(root))


def test_h_outside_1_2_3_raises_like_the_reference(tmp_path):
    ds, root = _write(tmp_path, 'disjoint', False)
    _seed(1)
    db = Subgraphs(root, 'train', data_io.load_labels(root), n_way=ds.n_way, k_shot=ds.k_spt, k_query=ds.k_qry,
                   batchsz=1, args=_args(ds), adjs=data_io.load_graphs(root), h=4)
    with pytest.raises(NameError):                       # subgraph_data_processing.py:300-311
        db[0]


@pytest.mark.parametrize("kind", ['disjoint', 'shared', 'link'])
def test_episodes_match_the_unmodified_reference(tmp_path, kind):
    if not ref_loader.available():
        pytest.skip("reference tree not mounted")
    _, _, sdp = ref_loader.load()
    ds, root = _write(tmp_path, kind, True)
    info = data_io.load_labels(root)
    with open(os.path.join(root, data_io.GRAPH_PKL), "rb") as f:
        dgl_graphs = pickle.load(f)
    kw = dict(n_way=ds.n_way, k_shot=ds.k_spt, k_query=ds.k_qry, batchsz=6, args=_args(ds), h=ds.h)
    ds.sample_nodes = 10 ** 6                            # no cap: sampled node sets are order dependent
    kw['args'] = _args(ds)
    _seed(11)
    ref = sdp.Subgraphs(root, 'train', info, adjs=dgl_graphs, **kw)
    _seed(11)
    ours = Subgraphs(root, 'train', info, adjs=data_io.load_graphs(root), **kw)   # graph_csr.npz path
    assert ours.support_x_batch == ref.support_x_batch and ours.query_x_batch == ref.query_x_batch
    assert [g.n for g in data_io.as_parent_graphs(dgl_graphs)] == [g.n for g in ours.G]
    for idx in range(3):
        _seed(100 + idx)
        r = ref[idx]
        _seed(100 + idx)
        o = ours[idx]
        for a, b in ((1, 1), (3, 3)):                    # labels (Disjoint: same random relabelling)
            assert torch.equal(r[a], o[b])
        assert r[8] == o[8] and r[9] == o[9]             # graph ids
        for gi, ci, ni in ((0, 4, 6), (2, 5, 7)):
            rg, og = r[gi], o[gi]
            assert list(rg.batch_num_nodes) == list(og.batch_num_nodes)
            roff = np.concatenate([[0], np.cumsum(rg.batch_num_nodes)])
            rs, rd = rg.edges()
            rs, rd = rs.numpy(), rd.numpy()
            odst = np.repeat(np.arange(og.n_nodes), np.diff(og.indptr))
            for k in range(len(rg.batch_num_nodes)):
                rn, on = np.asarray(r[ni][k]), np.asarray(o[ni][k])
                assert sorted(rn.tolist()) == on.tolist()            # same node set; ours ascending
                rc = np.atleast_1d(r[ci][k].numpy())
                oc = np.atleast_1d(o[ci][k].numpy())
                assert rn[rc].tolist() == on[oc].tolist()            # centres are the same parent nodes
                m = (rd >= roff[k]) & (rd < roff[k + 1])
                redges = sorted(zip(rn[rs[m] - roff[k]].tolist(), rn[rd[m] - roff[k]].tolist()))
                mo = (odst >= roff[k]) & (odst < roff[k + 1])
                oedges = sorted(zip(on[og.indices[mo] - roff[k]].tolist(), on[odst[mo] - roff[k]].tolist()))
                assert redges == oedges                              # same induced edge multiset


def _load_train():
    # by path: the reference's own train.py may be importable as `train` once ref_loader has run
    import importlib.util
    spec = importlib.util.spec_from_file_location("gmeta_b200_train", os.path.join(ROOT, "train.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cli_flags_and_defaults_match_the_reference():
    train = _load_train()
    got = vars(train.parse(["--data_dir", "x/", "--task_setup", "Disjoint"]))
    want = dict(epoch=10, n_way=3, k_spt=3, k_qry=24, task_num=8, meta_lr=1e-3, update_lr=1e-3, update_step=5,
                update_step_test=10, input_dim=1, hidden_dim=64, attention_size=32, data_dir="x/", no_finetune=True,
                task_setup='Disjoint', method='G-Meta', task_n=1, task_mode='False', val_result_report_steps=100,
                train_result_report_steps=30, num_workers=0, batchsz=1000, link_pred_mode='False', h=2,
                sample_nodes=1000)                                   # train.py:153-177
    for k, v in want.items():
        assert got[k] == v, k
    assert set(got) - set(want) == {"eval_batch", "device_extract", "aggregation"}
    a = train.parse(["--data_dir", "x/", "--task_setup", "Shared", "--link_pred", "True", "--hid", "128"])   # prefixes
    assert a.link_pred_mode == 'True' and a.hidden_dim == 128
    cfg = train.build_config([np.zeros((4, 5))], a, 2)                # train.py:67-75
    assert cfg == [('GraphConv', [5, 128]), ('GraphConv', [128, 128]), ('Linear', [128, 2]), ('LinkPred', [True])]
    if ref_loader.available():                                       # flag list of the reference source itself
        import re
        src = open(os.path.join(ref_loader.REFERENCE_DIR, "train.py")).read()
        flags = set(re.findall(r"add_argument\(['\"]--([a-z_]+)['\"]", src))
        assert flags == set(want)


@pytest.mark.parametrize("kind", ['disjoint', 'shared', 'link'])
def test_centre_requests_describe_the_same_episodes(tmp_path, kind):
    """Subgraphs.centre_requests (input of the device-side extraction) against __getitem__ of the same tasks."""
    ds, root = _write(tmp_path, kind, False)
    info = data_io.load_labels(root)
    db = Subgraphs(root, 'train', info, n_way=ds.n_way, k_shot=ds.k_spt, k_query=ds.k_qry, batchsz=4, args=_args(ds),
                   adjs=data_io.load_graphs(root), h=ds.h)
    _seed(21)
    eps = [db[i] for i in (2, 0)]
    _seed(21)
    req_s, req_q = db.centre_requests([2, 0])
    for req, (ix, iy, ic, inn, ig) in ((req_s, (0, 1, 4, 6, 8)), (req_q, (2, 3, 5, 7, 9))):
        assert req.sub_off.tolist() == np.concatenate([[0], np.cumsum([len(e[ig]) for e in eps])]).tolist()
        assert req.labels.tolist() == [int(v) for e in eps for v in e[iy]]
        assert req.graph_idx.tolist() == [g for e in eps for g in e[ig]]
        k = 0
        for e in eps:
            for s in range(len(e[ig])):
                c = np.atleast_1d(e[ic][s].numpy())
                assert req.centre_a[k] == e[inn][s][int(c[0])]
                if ds.link_pred:
                    assert req.centre_b[k] == e[inn][s][int(c[1])]
                k += 1


@pytest.mark.parametrize("kind", ['disjoint', 'shared', 'link'])
def test_sampler_against_committed_reference_fixture(tmp_path, kind):
    """Same check as test_episodes_match_the_unmodified_reference, against tests/golden/sampler_<kind>.json that
    oracle/make_golden_sampler.py recorded from the reference class -- runs where the reference is not mounted."""
    import json
    from oracle import make_golden_sampler as G
    with open(os.path.join(ROOT, "tests", "golden", "sampler_%s.json" % kind)) as f:
        gold = json.load(f)
    root = str(tmp_path / kind)
    ds = G.write_dataset(kind, root)
    info = data_io.load_labels(root)
    G.seed_all(G.SEED_TASKS)
    ours = Subgraphs(root, 'train', info, n_way=ds.n_way, k_shot=ds.k_spt, k_query=ds.k_qry, batchsz=G.BATCHSZ,
                     args=G.sampler_args(ds), adjs=data_io.load_graphs(root), h=ds.h)
    assert ours.support_x_batch == gold["support_x_batch"] and ours.query_x_batch == gold["query_x_batch"]
    for idx, ep in enumerate(gold["episodes"]):
        G.seed_all(G.SEED_EP + idx)
        o = ours[idx]
        assert o[1].tolist() == ep["y_spt"] and o[3].tolist() == ep["y_qry"]
        assert list(o[8]) == ep["g_spt"] and list(o[9]) == ep["g_qry"]
        for (gi, ci, ni), subs in zip(((0, 4, 6), (2, 5, 7)), ep["sets"]):
            og = o[gi]
            off = np.concatenate([[0], np.cumsum(og.batch_num_nodes)])
            odst = np.repeat(np.arange(og.n_nodes), np.diff(og.indptr))
            assert len(subs) == len(og.batch_num_nodes)
            for k, sub in enumerate(subs):
                on = np.asarray(o[ni][k])
                assert on.tolist() == sub["nodes"]
                assert on[np.atleast_1d(o[ci][k].numpy())].tolist() == sub["centres"]
                m = (odst >= off[k]) & (odst < off[k + 1])
                edges = sorted(zip(on[og.indices[m] - off[k]].tolist(), on[odst[m] - off[k]].tolist()))
                assert [list(e) for e in edges] == [list(e) for e in sub["edges"]]

"""CPU: host-side logic of the product (packing, tiling, extraction, parameter layout, label
validation) and the C-ABI surface (library loads and exports every symbol the header declares;
no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from gmeta_b200 import _lib, packing
from gmeta_b200.learner import Classifier, ModelSpec, tile_table
from gmeta_b200.packed import PackedSubgraphBatch, SubgraphCSR, csr_transpose
from gmeta_b200.subgraphs import ParentGraph, extract_subgraph, extract_subgraph_link_pred
from oracle import ref_loader
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gmeta_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gmeta_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 18
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(h, name), "libgmeta_b200.so does not export %s" % name
    assert sorted(_lib.exported_symbols()) == declared       # the ctypes table binds all of them
    assert _lib.lib().gmeta_version() >= 100
    assert _lib.lib().gmeta_error_string(-3).decode().startswith("shape")


def test_ctypes_structs_match_header_layout(tmp_path):
    """sizeof / offsetof of the ctypes mirrors against the header, measured with a gcc-compiled probe."""
    import subprocess
    probe = tmp_path / "probe.c"
    probe.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "gmeta_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gmeta_packed_set_t), sizeof(gmeta_model_t),'
        ' sizeof(gmeta_step_args_t), offsetof(gmeta_packed_set_t, centre_pos), offsetof(gmeta_step_args_t, pruned_forward),'
        ' offsetof(gmeta_step_args_t, meta_grad), offsetof(gmeta_step_args_t, workspace_bytes));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.PackedSet), ctypes.sizeof(_lib.Model), ctypes.sizeof(_lib.StepArgs),
            _lib.PackedSet.centre_pos.offset, _lib.StepArgs.pruned_forward.offset, _lib.StepArgs.meta_grad.offset,
            _lib.StepArgs.workspace_bytes.offset]
    assert got == want, (got, want)


def test_parameter_counts_match_reference_logs():
    """Trainable-parameter counts printed by the reference (test.ipynb:27,112,211,284,369)."""
    cases = [
        ([('GraphConv', [128, 256]), ('GraphConv', [256, 256]), ('Linear', [256, 3])], 99587),
        ([('GraphConv', [50, 128]), ('GraphConv', [128, 128]), ('Linear', [128, 2])], 23298),
        ([('GraphConv', [512, 128]), ('GraphConv', [128, 128]), ('Linear', [128, 3])], 82563),   # 3x128 head: test.ipynb:206,211
        ([('GraphConv', [5, 128]), ('GraphConv', [128, 128]), ('Linear', [128, 2]), ('LinkPred', [True])], 17794),
        ([('GraphConv', [1, 256]), ('GraphConv', [256, 256]), ('Linear', [256, 2]), ('LinkPred', [True])], 67330),
    ]
    for cfg, want in cases:
        net = Classifier(cfg)
        n = sum(int(np.prod(p.shape)) for p in net.parameters())
        assert n == want, (cfg, n)
        spec = ModelSpec(cfg)
        assert spec.n_params_padded >= n and all(o % 4 == 0 for o in spec.offsets)
        flat = spec.flatten(list(net.parameters()))
        for p, v in zip(net.parameters(), spec.unflatten(flat)):
            assert torch.equal(p.detach(), v)


def test_parameter_init_matches_reference_order_and_rng():
    if not ref_loader.available():
        pytest.skip("reference tree not mounted")
    learner, _, _ = ref_loader.load()
    cfg = [('GraphConv', [12, 16]), ('GraphConv', [16, 16]), ('Linear', [16, 4]), ('LinkPred', [True])]
    torch.manual_seed(222)
    ref = learner.Classifier(cfg)
    torch.manual_seed(222)
    ours = Classifier(cfg)
    for a, b in zip(ref.parameters(), ours.parameters()):
        assert a.shape == b.shape and torch.equal(a, b)


def test_tile_table_never_straddles_tasks():
    trp = np.array([0, 5, 5, 300, 428, 1000])
    row0, nrows, task = tile_table(trp)
    assert nrows.min() >= 1 and nrows.max() <= _lib.TILE_ROWS
    covered = np.concatenate([np.arange(r, r + n) for r, n in zip(row0, nrows)])
    assert np.array_equal(covered, np.arange(1000))
    for r, n, t in zip(row0, nrows, task):
        assert trp[t] <= r and r + n <= trp[t + 1]


def test_csr_transpose_and_batching():
    rng = np.random.default_rng(0)
    subs = []
    for n, e in [(5, 12), (1, 0), (9, 30)]:
        subs.append(SubgraphCSR.from_edges(rng.integers(0, n, e), rng.integers(0, n, e), n, centre=0))
    g = PackedSubgraphBatch.batch(subs)
    assert g.batch_num_nodes == [5, 1, 9] and g.n_nodes == 15 and g.n_edges == 42
    src, dst = g.edges()
    dense = np.zeros((15, 15), dtype=np.int64)
    np.add.at(dense, (dst, src), 1)
    tp, ti = g.t_indptr, g.t_indices
    dense_t = np.zeros((15, 15), dtype=np.int64)
    np.add.at(dense_t, (np.repeat(np.arange(15), np.diff(tp)), ti), 1)
    assert np.array_equal(dense.T, dense_t)
    assert (src[dst < 5] < 5).all() and (src[dst >= 6] >= 6).all()      # block diagonal
    tp2, ti2 = csr_transpose(g.indptr, g.indices, 15)
    assert np.array_equal(tp, tp2) and np.array_equal(ti, ti2)


@pytest.mark.parametrize("kind", H.TINY_KINDS)
def test_packing_layout(kind):
    ds = H.tiny_dataset(kind)
    mb = ds.sample_meta_batch(np.random.default_rng(4))
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    ps = packing.plan_set(xq, cq, 0)
    buf = np.zeros(ps.end, dtype=np.int32)
    rows = np.array([f.shape[0] for f in ds.feats])
    goff = np.concatenate([[0], np.cumsum(rows)])[:-1]
    packing.fill_set(buf, ps, xq, yq, cq, nq, gq, goff)
    o = ps.off
    indptr = buf[o["indptr"]:o["indptr"] + ps.N + 1]
    indices = buf[o["indices"]:o["indices"] + ps.E]
    assert indptr[0] == 0 and indptr[-1] == ps.E and (np.diff(indptr) >= 0).all()
    assert all(v % 4 == 0 for v in o.values())
    table = np.vstack(ds.feats)
    feat_row = buf[o["feat_row"]:o["feat_row"] + ps.N]
    centre = buf[o["centre_row"]:o["centre_row"] + ps.S * ps.cps]
    labels = buf[o["labels"]:o["labels"] + ps.S]
    trp = buf[o["task_row_ptr"]:o["task_row_ptr"] + ps.T + 1]
    tsp = buf[o["task_sub_ptr"]:o["task_sub_ptr"] + ps.T + 1]
    for t in range(ps.T):
        a, b = trp[t], trp[t + 1]
        assert b - a == xq[t].n_nodes
        # features reached through feat_row == the reference's host gather (meta.py:119-120)
        want = np.vstack([ds.feats[gq[t][j]][np.array(x)] for j, x in enumerate(nq[t])])
        assert np.array_equal(table[feat_row[a:b]], want)
        # edges stay inside the task, shifted by its row offset
        lo, hi = indptr[a], indptr[b]
        assert np.array_equal(indices[lo:hi], xq[t].indices + a)
        off = np.concatenate([[0], np.cumsum(xq[t].batch_num_nodes)])[:-1] + a
        c = cq[t].numpy()
        want_c = (c + off[:, None]).reshape(-1) if ps.cps == 2 else c + off
        assert np.array_equal(centre[tsp[t] * ps.cps:tsp[t + 1] * ps.cps], want_c)
        assert np.array_equal(labels[tsp[t]:tsp[t + 1]], yq[t].numpy())
    assert ps.cps == (2 if ds.link_pred else 1)


@pytest.mark.parametrize("kind", H.TINY_KINDS)
def test_active_rows_cover_every_row_with_a_gradient(kind):
    """act[L-1] = centre rows, act[l-1] = in-neighbours of act[l]: checked against autograd on the
    oracle -- every row of dL/dZ_l outside the list is exactly zero."""
    from oracle import gmeta_oracle as O
    ds = H.tiny_dataset(kind)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = ds.sample_task(np.random.default_rng(3))
    cfg = ds.config()
    L = sum(1 for n, _ in cfg if n == 'GraphConv')
    act = packing.active_rows(xq, cq, L)
    torch.manual_seed(0)
    params = O.init_params(cfg)
    feat = O.gather_features(ds.feats, gq, nq)
    g = H.to_ograph(xq)
    # pre-activations of every layer with a hook on their gradient
    hs, h, idx = [], feat, 0
    for l in range(L):
        norm = torch.pow(g.in_degrees().float().clamp(min=1), -0.5).unsqueeze(1)
        z = (g.aggregate_sum(h * norm) @ params[idx]) * norm + params[idx + 1]
        z.retain_grad()
        hs.append(z)
        h = torch.relu(z)
        idx += 2
    off = torch.cumsum(torch.LongTensor([0] + list(g.batch_num_nodes)), 0)[:-1]
    r = torch.cat((h[cq[:, 0] + off], h[cq[:, 1] + off]), 1) if ds.link_pred else h[cq + off]
    (r @ params[idx].T + params[idx + 1]).pow(2).sum().backward()
    for l in range(L):
        nz = np.nonzero(hs[l].grad.abs().sum(1).numpy())[0]
        assert set(nz.tolist()) <= set(act[l].tolist()), (kind, l)
        assert (np.diff(act[l]) > 0).all()


def test_label_validation_mirrors_reference_errors():
    y = torch.LongTensor
    assert packing.validate_labels([y([0, 0, 1, 1])], [y([0, 1, 1, 0])], 2) == 2
    with pytest.raises(RuntimeError):
        packing.validate_labels([y([0, 0, 1])], [y([0, 1])], 2)            # class 1 has < k_spt members
    with pytest.raises(RuntimeError):
        packing.validate_labels([y([0, 0, 1, 1])], [y([0, 1, 1])], 2)      # unbalanced query


def test_label_validation_in_the_library_agrees_with_the_numpy_check():
    """Meta.forward validates through gmeta_host_validate_labels (one call for all tasks, no interpreter lock held);
    same verdict and same class count as the numpy restatement on good, bad, large-valued and negative labels."""
    y = torch.LongTensor
    L = _lib.lib()
    cases = [
        ([y([0, 0, 1, 1])], [y([0, 1, 1, 0])], 2),
        ([y([0, 0, 1])], [y([0, 1])], 2),                                  # a support class below k_spt
        ([y([0, 0, 1, 1])], [y([0, 1, 1])], 2),                            # unbalanced query
        ([y([0, 0, 1, 1])], [y([0, 2, 0, 2])], 2),                         # different classes, same counts
        ([y([5, 5, 7, 7, 9, 9])], [y([9, 7, 5])], 2),
        ([y([-3, -3, 40000, 40000])], [y([40000, -3])], 2),                # outside the counting fast path
        ([y([2 ** 40, 2 ** 40, 1, 1])], [y([1, 2 ** 40, 1, 2 ** 40])], 2),
        ([y([0, 0, 1, 1]), y([3, 3, 4, 4, 5, 5])], [y([0, 1]), y([5, 4, 3])], 2),   # max classes over the tasks
        ([y([0, 0, 1, 1]), y([3, 3, 4, 4, 5])], [y([0, 1]), y([5, 4, 3])], 2),      # second task bad
        ([np.array([1, 1, 0, 0])], [[0, 1]], 2),                           # numpy array / plain list
    ]
    rng = np.random.default_rng(9)
    for _ in range(40):
        n_cls, k = int(rng.integers(1, 6)), int(rng.integers(1, 4))
        cls = rng.choice(50, n_cls, replace=False)
        ys = np.repeat(cls, k + rng.integers(0, 2, n_cls))                  # sometimes more than k_spt members
        yq = np.repeat(cls, int(rng.integers(1, 4)))
        if rng.random() < 0.3:
            yq = yq[:-1] if yq.shape[0] > 1 else yq                         # unbalance it
        if rng.random() < 0.2:
            yq = yq.copy()
            yq[yq == cls[0]] = 99                                           # other class, same counts
        cases.append(([y(rng.permutation(ys))], [y(rng.permutation(yq))], k))

    def verdict(fn):
        try:
            return fn()
        except RuntimeError as e:
            return str(e)
    for ys, yq, k in cases:
        a = verdict(lambda: packing.validate_labels(ys, yq, k))
        b = verdict(lambda: packing.validate_labels(ys, yq, k, L))
        assert a == b, (ys, yq, k, a, b)


def test_extraction_matches_reference_generate_subgraph():
    """Same node set / induced edges as Subgraphs.generate_subgraph when the neighbourhood is
    below the sampling cap (subgraph_data_processing.py:295-321), for h = 1, 2, 3."""
    if not ref_loader.available():
        pytest.skip("reference tree not mounted")
    _, _, sdp = ref_loader.load()
    dgl = ref_loader.shim_dgl()
    rng = np.random.default_rng(8)
    n, e = 300, 900
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    G = ParentGraph.from_edges(src, dst, n)
    Gd = dgl.DGLGraph(src, dst, n)
    for h in (1, 2, 3):
        ref = sdp.Subgraphs.__new__(sdp.Subgraphs)
        ref.h, ref.sample_nodes, ref.subgraphs = h, 10 ** 6, {}
        for i in rng.choice(n, 5, replace=False):
            sub, centre, parent = ref.generate_subgraph(Gd, int(i), "0_%d" % i)
            ours = extract_subgraph(G, int(i), h, 10 ** 6)
            assert sorted(parent) == ours.parent_nid.tolist()
            assert ours.parent_nid[ours.centre] == i and parent[centre] == i
            # induced edge multiset in parent ids
            rs, rd = sub.edges()
            ref_edges = sorted(zip(np.array(parent)[rs.numpy()].tolist(), np.array(parent)[rd.numpy()].tolist()))
            od = np.repeat(np.arange(ours.n), np.diff(ours.indptr))
            our_edges = sorted(zip(ours.parent_nid[ours.indices].tolist(), ours.parent_nid[od].tolist()))
            assert ref_edges == our_edges
    ref = sdp.Subgraphs.__new__(sdp.Subgraphs)
    ref.sample_nodes, ref.subgraphs = 10 ** 6, {}
    sub, centres, parent = ref.generate_subgraph_link_pred(Gd, 3, 77, "0_3_77")
    ours = extract_subgraph_link_pred(G, 3, 77, 10 ** 6)
    assert sorted(parent) == ours.parent_nid.tolist()
    assert [parent[c] for c in centres] == [3, 77] and ours.parent_nid[ours.centre].tolist() == [3, 77]


def test_extraction_sampling_cap():
    """Above the cap: exactly sample_nodes uniformly chosen nodes plus the centre (:312-314)."""
    rng = np.random.default_rng(1)
    n = 400
    src, dst = rng.integers(0, n, 6000), rng.integers(0, n, 6000)
    G = ParentGraph.from_edges(src, dst, n)
    full = extract_subgraph(G, 5, 2, 10 ** 6)
    assert full.n > 60
    s = extract_subgraph(G, 5, 2, 50, np.random.default_rng(0))
    assert s.n in (50, 51) and 5 in s.parent_nid and set(s.parent_nid) <= set(full.parent_nid)
    assert s.parent_nid[s.centre] == 5
    with pytest.raises(NameError):
        extract_subgraph(G, 5, 4, 50)              # the reference only defines h in {1,2,3}


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gmeta_b200")
    files = [os.path.join(pkg, fn) for fn in os.listdir(pkg) if fn.endswith(".py")] + [os.path.join(ROOT, "train.py")]
    files += [os.path.join(pkg, "csrc", fn) for fn in os.listdir(os.path.join(pkg, "csrc")) if fn.endswith((".cu", ".cuh"))]
    assert len(files) > 20
    for path in files:
        src = open(path).read()
        assert "oracle" not in src.replace("the oracle", ""), path
    # bench.py may use the oracle, but only for the CPU baseline / reference arm (cpu_arm)
    bench = open(os.path.join(ROOT, "bench.py")).read()
    assert bench.count("from oracle import") == 1
    assert bench.split("from oracle import")[0].rsplit("\ndef ", 1)[1].startswith("cpu_arm(")


@pytest.mark.parametrize("kind", H.TINY_KINDS)
def test_fast_packing_equals_reference_packing(kind):
    """Meta.upload_batch's path (threaded gmeta_host_pack_csr + active rows from the packed arrays + parent ids
    carried by the batch) against the plain per-task numpy packing, segment by segment."""
    ds = H.tiny_dataset(kind)
    mb = ds.sample_meta_batch(np.random.default_rng(4), 3)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    L = ds.h
    goff = np.concatenate([[0], np.cumsum([f.shape[0] for f in ds.feats])])[:-1]
    st = packing.Staging(torch.device("cpu"))
    for variant in ("carried ids", "different id lists", "id lists"):
        if variant == "different id lists":
            # the caller hands OTHER parent ids than the batch was built with (re-sampled / permuted lists):
            # they are what meta.py:119-120 would gather, so they must win over the ids the batch carries
            shift = lambda lists, gi: [[(np.asarray(a) + 1) % ds.feats[gi[t][j]].shape[0] for j, a in enumerate(lst)]  # noqa: E731
                                       for t, lst in enumerate(lists)]
            ns, nq = shift(ns, gs), shift(nq, gq)
            mb = (xs, ys, xq, yq, cs, cq, ns, nq, gs, gq)
        if variant == "id lists":
            for g in xs + xq:
                g.parent_ids = None
        ps_s, ps_q, end = packing.pack_meta_batch(st, mb, goff, L, _lib.lib(), n_threads=3)
        fast = st.host.numpy()
        for ps, (x, y, c, n, g) in ((ps_s, (xs, ys, cs, ns, gs)), (ps_q, (xq, yq, cq, nq, gq))):
            ref = packing.plan_set(x, c, 0, L)
            buf = np.zeros(ref.end, dtype=np.int32)
            packing.fill_set(buf, ref, x, y, c, n, g, goff)
            assert (ps.N, ps.E, ps.S, ps.T, ps.n_tiles, ps.cps, ps.max_rows_per_task) == \
                (ref.N, ref.E, ref.S, ref.T, ref.n_tiles, ref.cps, ref.max_rows_per_task)
            assert set(ps.sizes) == set(ref.sizes)
            for k, n_el in ref.sizes.items():
                assert ps.sizes[k] == n_el, k
                assert ps.off[k] % 4 == 0 and ps.off[k] + n_el <= end
                assert np.array_equal(fast[ps.off[k]:ps.off[k] + n_el], buf[ref.off[k]:ref.off[k] + n_el]), (variant, k)
            for l in range(L):
                assert ps.act[l]["n"] == ref.act[l]["n"] and ps.act[l]["n_tiles"] == ref.act[l]["n_tiles"]


def test_slim_host_pack_writes_the_same_host_segments_as_the_full_packer():
    """packing.pack_meta_batch_slim packs only what the host alone knows; those segments equal the full packer's and the
    device-derived ones are laid out behind them at their upper bounds (the device half is tested on the GPU)."""
    import torch
    from gmeta_b200 import packing
    from tests import helpers as H
    for kind in ('disjoint', 'link', 'shared'):
        ds = H.tiny_dataset(kind)
        mb = ds.sample_meta_batch(np.random.default_rng(3), 3)
        goff = np.concatenate([[0], np.cumsum([f.shape[0] for f in ds.feats])])[:-1]
        L = ds.h                                   # one GraphConv per hop
        full, slim = packing.Staging(torch.device("cpu")), packing.Staging(torch.device("cpu"))
        fs, fq, _ = packing.pack_meta_batch(full, mb, goff, L, _lib.lib())
        ss, sq, n_host, n_total = packing.pack_meta_batch_slim(slim, mb, goff, L, _lib.lib())
        a, b = full.host.numpy(), slim.host.numpy()
        assert n_host < n_total
        for f, s_ in ((fs, ss), (fq, sq)):
            assert (f.N, f.E, f.S, f.T, f.n_tiles, f.cps) == (s_.N, s_.E, s_.S, s_.T, s_.n_tiles, s_.cps)
            for k in packing.HOST_SEGS:
                n = f.sizes[k]
                assert s_.off[k] + n <= n_host
                assert np.array_equal(a[f.off[k]:f.off[k] + n], b[s_.off[k]:s_.off[k] + n]), k
            for k, cap in s_.cap.items():
                assert s_.off[k] >= n_host and cap >= f.sizes[k], k


@pytest.mark.parametrize("seed", range(6))
def test_host_packers_on_ragged_random_batches(seed):
    """Ragged inputs the tiny datasets do not produce: subgraphs of ONE node, subgraphs without edges, multi-edges,
    tasks of different subgraph counts, several parent graphs, two centres per subgraph.  The library packer
    (pack_meta_batch), the slim packer and the plain numpy packing (plan_set + fill_set) must agree segment by segment."""
    from gmeta_b200.packed import PackedSubgraphBatch, SubgraphCSR
    rng = np.random.default_rng(100 + seed)
    n_graphs = int(rng.integers(1, 4))
    graph_n = [int(rng.integers(40, 90)) for _ in range(n_graphs)]
    goff = np.concatenate([[0], np.cumsum(graph_n)])[:-1]
    link = bool(seed % 2)
    L = 1 + seed % 3
    T = int(rng.integers(1, 5))

    def subgraph(gi):
        n = int(rng.choice([1, 2, 3, 7, 20, 33]))
        e = 0 if rng.random() < 0.25 else int(rng.integers(0, 4 * n + 1))
        src, dst = rng.integers(0, n, e), rng.integers(0, n, e)          # multi-edges and self-loops may occur
        ids = rng.choice(graph_n[gi], n, replace=False)
        c = [int(rng.integers(0, n)), int(rng.integers(0, n))] if link else int(rng.integers(0, n))
        return SubgraphCSR.from_edges(src, dst, n, ids, c)

    def one_set(max_sub):
        xs, ys, cs, ns, gs = [], [], [], [], []
        for _ in range(T):
            S = int(rng.integers(1, max_sub + 1))
            gi = [int(rng.integers(0, n_graphs)) for _ in range(S)]
            subs = [subgraph(g) for g in gi]
            xs.append(PackedSubgraphBatch.batch(subs))
            ys.append(torch.LongTensor(rng.integers(0, 3, S)))
            cs.append(torch.LongTensor(np.array([s.centre for s in subs])))
            ns.append(xs[-1].parent_id_lists)
            gs.append(gi)
        return xs, ys, cs, ns, gs
    xs, ys, cs, ns, gs = one_set(4)
    xq, yq, cq, nq, gq = one_set(9)
    mb = (xs, ys, xq, yq, cs, cq, ns, nq, gs, gq)
    full, slim = packing.Staging(torch.device("cpu")), packing.Staging(torch.device("cpu"))
    ps_s, ps_q, end = packing.pack_meta_batch(full, mb, goff, L, _lib.lib(), n_threads=2)
    ss, sq, n_host, n_total = packing.pack_meta_batch_slim(slim, mb, goff, L, _lib.lib(), n_threads=2)
    fast, sl = full.host.numpy(), slim.host.numpy()
    for ps, s_, (x, y, c, n, g) in ((ps_s, ss, (xs, ys, cs, ns, gs)), (ps_q, sq, (xq, yq, cq, nq, gq))):
        ref = packing.plan_set(x, c, 0, L)
        buf = np.zeros(ref.end, dtype=np.int32)
        packing.fill_set(buf, ref, x, y, c, n, g, goff)
        assert (ps.N, ps.E, ps.S, ps.T, ps.n_tiles, ps.cps) == (ref.N, ref.E, ref.S, ref.T, ref.n_tiles, ref.cps)
        assert ps.cps == (2 if link else 1)
        for k, n_el in ref.sizes.items():
            assert ps.sizes[k] == n_el, k
            assert np.array_equal(fast[ps.off[k]:ps.off[k] + n_el], buf[ref.off[k]:ref.off[k] + n_el]), k
        for k in packing.HOST_SEGS:
            n_el = ref.sizes[k]
            assert np.array_equal(sl[s_.off[k]:s_.off[k] + n_el], buf[ref.off[k]:ref.off[k] + n_el]), k
        # the feature rows are the parent ids shifted by their graph's first row (meta.py:119-120)
        want = np.concatenate([np.concatenate([np.asarray(ids) + goff[gi] for ids, gi in zip(n[t], g[t])]) for t in range(ps.T)])
        assert np.array_equal(fast[ps.off["feat_row"]:ps.off["feat_row"] + ps.N], want)


def _tiny_meta(method='G-Meta'):
    import argparse
    from gmeta_b200.meta import Meta
    cfg = [('GraphConv', [12, 16]), ('GraphConv', [16, 16]), ('Linear', [16, 3])]
    args = argparse.Namespace(update_lr=0.05, meta_lr=1e-3, n_way=3, k_spt=2, k_qry=3, task_num=2, update_step=2,
                              update_step_test=2, method=method)
    return Meta(args, cfg), cfg


@pytest.mark.skipif(torch.cuda.is_available(), reason="the point is a box without a GPU")
def test_product_fails_loudly_without_a_gpu():
    """No CPU path and no silent fallback: on a box without a CUDA device Meta.forward / finetunning / Classifier.forward
    raise GMetaError instead of computing anything (the oracle is never reached: test_product_code_never_imports_the_oracle)."""
    ds = H.tiny_dataset('disjoint')
    mb = ds.sample_meta_batch(np.random.default_rng(1))
    m, _ = _tiny_meta()
    for call in (m.forward, m.finetunning, m.finetunning_batch):
        with pytest.raises(_lib.GMetaError, match="CUDA device"):
            call(*mb, ds.feats)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    with pytest.raises(_lib.GMetaError, match="CUDA device"):
        m.net(xs[0], cs[0], torch.zeros(xs[0].n_nodes, 12))


def test_unknown_method_fails_like_the_reference():
    """meta.py:236-246: only 'G-Meta' is defined; any other --method leaves `accs` unbound."""
    m, _ = _tiny_meta(method='MAML')
    with pytest.raises(UnboundLocalError):
        m.forward(*([None] * 11))
    with pytest.raises(UnboundLocalError):
        m.finetunning(*([None] * 11))

"""CPU, build container only: the oracle restatement (oracle/gmeta_oracle.py) against the UNMODIFIED
reference (G-Meta/learner.py + meta.py through oracle/ref_loader.py and the DGL stand-in) on FRESH
random inputs drawn in this test -- not on the committed fixtures.  Skips where /root/reference is
not mounted (the GPU box).  Single-threaded, so torch's index_add is reproducible and equality is
bit-for-bit.
"""
import numpy as np
import pytest
import torch

from oracle import gmeta_oracle as O
from oracle import ref_loader
from tests import helpers as H

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _to_dgl(dgl, p):
    s, d = p.edges()
    return dgl.DGLGraph(s, d, p.n_nodes, batch_num_nodes=p.batch_num_nodes)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_graphconv_random_multigraphs(seed):
    """GraphConv.forward (learner.py:25-56) + autograd, both branch orders, random multigraphs."""
    learner, _, _ = ref_loader.load()
    dgl = ref_loader.shim_dgl()
    import torch.nn.functional as F
    rng = np.random.default_rng(1000 + seed)
    for _ in range(4):
        n = int(rng.integers(5, 200))
        e = int(rng.integers(0, 6 * n))
        fi, fo = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        src, dst = rng.integers(0, n, size=e), rng.integers(0, n, size=e)
        x = torch.tensor(rng.standard_normal((n, fi), dtype=np.float32), requires_grad=True)
        w = torch.tensor(rng.standard_normal((fi, fo), dtype=np.float32) * 0.3, requires_grad=True)
        b = torch.tensor(rng.standard_normal(fo, dtype=np.float32) * 0.1, requires_grad=True)
        gy = torch.tensor(rng.standard_normal((n, fo), dtype=np.float32))
        y_ref = learner.GraphConv(fi, fo, activation=F.relu)(dgl.DGLGraph(src, dst, n), x, w, b)
        g_ref = torch.autograd.grad(y_ref, (x, w, b), gy)
        y = O.gcn_layer(O.OGraph(src, dst, n), x, w, b, fi, fo)
        g = torch.autograd.grad(y, (x, w, b), gy)
        assert torch.equal(y, y_ref), (n, e, fi, fo)
        for a, r in zip(g, g_ref):
            assert torch.equal(a, r), (n, e, fi, fo)


@pytest.mark.parametrize("seed", [0, 1])
def test_proto_losses_random(seed):
    """proto_loss_spt / proto_loss_qry (meta.py:28-79) incl. the gradient through the prototypes."""
    _, meta, _ = ref_loader.load()
    rng = np.random.default_rng(2000 + seed)
    for _ in range(5):
        ncls, ks, kq, d = int(rng.integers(2, 6)), int(rng.integers(1, 5)), int(rng.integers(1, 12)), int(rng.integers(2, 33))
        labels = rng.choice(60, ncls, replace=False)
        ys = torch.LongTensor(rng.permutation(np.repeat(labels, ks + int(rng.integers(0, 2)))))
        yq = torch.LongTensor(rng.permutation(np.repeat(labels, kq)))
        zs = torch.tensor(rng.standard_normal((ys.numel(), d), dtype=np.float32), requires_grad=True)
        zq = torch.tensor(rng.standard_normal((yq.numel(), d), dtype=np.float32), requires_grad=True)
        outs = []
        for mod in (meta, O):
            ls, acc_s, protos = mod.proto_loss_spt(zs, ys, ks)
            dzs = torch.autograd.grad(ls, zs, retain_graph=True)[0]
            lq, acc_q = mod.proto_loss_qry(zq, yq, protos)
            dzq, via = torch.autograd.grad(lq, (zq, zs))
            outs.append((ls.detach(), acc_s, protos.detach(), dzs, lq.detach(), acc_q, dzq, via))
        for a, r in zip(outs[1], outs[0]):
            assert torch.equal(a, r)


@pytest.mark.parametrize("kind,steps", [('disjoint', 2), ('disjoint', 5), ('wide', 3), ('shared', 3),
                                        ('deep', 2), ('link', 4)])
def test_meta_forward_and_finetunning_fresh_batches(kind, steps):
    """Meta.forward / Meta.finetunning (meta.py:101-234) on meta-batches sampled HERE (seeds differ from
    oracle/make_golden.py): accuracies, query losses, meta-gradient and the weights after Adam bit-equal,
    over two consecutive steps, at update_step in {2,3,4,5} (the K-step chain of meta.py:143-157)."""
    _, meta, _ = ref_loader.load()
    dgl = ref_loader.shim_dgl()
    ds = H.tiny_dataset(kind, seed=40 + steps)
    ds.update_step = steps
    ds.update_step_test = steps + 1
    rng = np.random.default_rng(300 + steps)
    torch.manual_seed(5)
    m = meta.Meta(ds.args(), ds.config())
    om = O.OracleMeta(ds.args(), ds.config(),
                      params=[p.detach().clone().requires_grad_(True) for p in m.net.parameters()])
    for it in range(2):
        mb = ds.sample_meta_batch(rng)
        xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
        dxs, dxq = [_to_dgl(dgl, x) for x in xs], [_to_dgl(dgl, x) for x in xq]
        oxs, oxq = [H.to_ograph(x) for x in xs], [H.to_ograph(x) for x in xq]
        fin_ref = m.finetunning(dxs, ys, dxq, yq, cs, cq, ns, nq, gs, gq, ds.feats)
        fin = om.finetunning(oxs, ys, oxq, yq, cs, cq, ns, nq, gs, gq, ds.feats)
        assert np.array_equal(np.asarray(fin_ref, dtype=np.float32), fin), (kind, it)
        acc_ref = m(dxs, ys, dxq, yq, cs, cq, ns, nq, gs, gq, ds.feats)
        acc = om.forward(oxs, ys, oxq, yq, cs, cq, ns, nq, gs, gq, ds.feats)
        assert np.array_equal(np.asarray(acc_ref, dtype=np.float32), acc), (kind, it)
        for p_ref, g, p in zip(m.net.parameters(), om.last_grads, om.vars):
            assert torch.equal(p_ref.grad, g), (kind, it)
            assert torch.equal(p_ref.detach(), p.detach()), (kind, it)


@pytest.mark.parametrize("link", [False, True])
def test_meta_on_degenerate_batches_and_nan_skip(link):
    """The inputs of tests/test_gpu_meta_ragged.py (one-node and edgeless subgraphs, a task without any edge, multi-edges,
    several parent graphs) through the unmodified reference and the oracle: finetunning, two forward steps bit-equal;
    then an all-NaN feature table: both skip the outer update (meta.py:163-164), parameters untouched."""
    import argparse
    from tests.test_gpu_meta_ragged import _ragged_meta_batch
    _, meta, _ = ref_loader.load()
    dgl = ref_loader.shim_dgl()
    mb, feats = _ragged_meta_batch(31 + int(link), link)
    cfg = [('GraphConv', [12, 16]), ('GraphConv', [16, 16]), ('Linear', [16, 3])] + ([('LinkPred', [True])] if link else [])
    args = argparse.Namespace(update_lr=0.05, meta_lr=1e-3, n_way=3, k_spt=2, k_qry=3, task_num=3, update_step=3,
                              update_step_test=4, method='G-Meta')
    torch.manual_seed(222)
    m = meta.Meta(args, cfg)
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in m.net.parameters():
            if p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=gen))
    om = O.OracleMeta(args, cfg, params=[p.detach().clone().requires_grad_(True) for p in m.net.parameters()])
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    dxs, dxq = [_to_dgl(dgl, x) for x in xs], [_to_dgl(dgl, x) for x in xq]
    oxs, oxq = [H.to_ograph(x) for x in xs], [H.to_ograph(x) for x in xq]
    one = lambda v: [v[1]]                                                                                # noqa: E731
    fin_ref = m.finetunning(one(dxs), one(ys), one(dxq), one(yq), one(cs), one(cq), one(ns), one(nq), one(gs), one(gq), feats)
    fin = om.finetunning(one(oxs), one(ys), one(oxq), one(yq), one(cs), one(cq), one(ns), one(nq), one(gs), one(gq), feats)
    assert np.array_equal(np.asarray(fin_ref, dtype=np.float32), fin)
    for it in range(2):
        acc_ref = m(dxs, ys, dxq, yq, cs, cq, ns, nq, gs, gq, feats)
        acc = om.forward(oxs, ys, oxq, yq, cs, cq, ns, nq, gs, gq, feats)
        assert np.array_equal(np.asarray(acc_ref, dtype=np.float32), acc), it
        for p_ref, g, p in zip(m.net.parameters(), om.last_grads, om.vars):
            assert torch.equal(p_ref.grad, g) and torch.equal(p_ref.detach(), p.detach()), it
    # NaN features: the mean query loss is NaN on both sides and neither applies the outer step
    bad = [np.full_like(f, np.nan) for f in feats]
    before = [p.detach().clone() for p in m.net.parameters()]
    m(dxs, ys, dxq, yq, cs, cq, ns, nq, gs, gq, bad)
    om.forward(oxs, ys, oxq, yq, cs, cq, ns, nq, gs, gq, bad)
    assert np.isnan(om.last_loss_q)
    for b, p_ref, p in zip(before, m.net.parameters(), om.vars):
        assert torch.equal(b, p_ref.detach()) and torch.equal(b, p.detach())

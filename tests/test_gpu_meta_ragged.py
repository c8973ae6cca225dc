"""GPU: Meta.forward / finetunning on DEGENERATE meta-batches the samplers rarely produce -- subgraphs of one node,
subgraphs without any edge, multi-edges and self-loops, centres without in-neighbours, tasks of different sizes, several
parent graphs -- against OracleMeta on the same inputs (the restatement of meta.py:101-234; accuracy vectors identical,
query loss within 1e-4, meta-gradient within the tolerance of tests/test_gpu_meta.py)."""
import argparse

import numpy as np
import pytest
import torch

from gmeta_b200.packed import PackedSubgraphBatch, SubgraphCSR
from oracle import gmeta_oracle as O
from tests import gpu_util as U
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _ragged_meta_batch(seed, link, n_way=3, k_spt=2, k_qry=3, tasks=3, f0=12):
    rng = np.random.default_rng(seed)
    graph_n = [50, 35, 64]
    feats = [rng.standard_normal((n, f0)).astype(np.float32) for n in graph_n]

    def subgraph(gi, kind):
        n = {0: 1, 1: int(rng.choice([2, 3, 5])), 2: int(rng.choice([9, 20, 33]))}[kind]
        e = 0 if (kind == 0 or rng.random() < 0.3) else int(rng.integers(1, 4 * n + 1))
        src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
        c = [int(rng.integers(0, n)), int(rng.integers(0, n))] if link else int(rng.integers(0, n))
        return SubgraphCSR.from_edges(src, dst, n, rng.choice(graph_n[gi], n, replace=False), c)

    def one_set(per_class, t):
        S = n_way * per_class
        labels = np.repeat(np.arange(n_way), per_class)
        rng.shuffle(labels)
        gi = [int(rng.integers(0, len(graph_n))) for _ in range(S)]
        kinds = rng.integers(0, 3, S)
        if t == 0:
            kinds[:] = 0                                   # a task made of one-node subgraphs only: no edge at all
        subs = [subgraph(g, int(k)) for g, k in zip(gi, kinds)]
        x = PackedSubgraphBatch.batch(subs)
        return x, torch.LongTensor(labels), torch.LongTensor(np.array([s.centre for s in subs])), x.parent_id_lists, gi
    out = [[] for _ in range(10)]
    for t in range(tasks):
        xs, ys, cs, ns, gs = one_set(k_spt, t)
        xq, yq, cq, nq, gq = one_set(k_qry, t)
        for k, v in zip((0, 1, 2, 3, 4, 5, 6, 7, 8, 9), (xs, ys, xq, yq, cs, cq, ns, nq, gs, gq)):
            out[k].append(v)
    return tuple(out), feats


def _nonzero_biases(params):
    # see tests/test_gpu_aggregation.py: rows without in-edges sit EXACTLY at the bias otherwise and their ReLU flips on noise
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for prm in params:
            if prm.dim() == 1:
                prm.copy_((0.05 * torch.randn(prm.shape, generator=gen)).to(prm.device))


@pytest.mark.parametrize("device_finish", [False, True])
@pytest.mark.parametrize("pruned", [True, False])
@pytest.mark.parametrize("link", [False, True])
def test_meta_on_degenerate_subgraphs_matches_oracle(link, pruned, device_finish):
    from gmeta_b200.meta import Meta
    mb, feats = _ragged_meta_batch(31 + int(link), link)
    cfg = [('GraphConv', [12, 16]), ('GraphConv', [16, 16]), ('Linear', [16, 3])] + ([('LinkPred', [True])] if link else [])
    args = argparse.Namespace(update_lr=0.05, meta_lr=1e-3, n_way=3, k_spt=2, k_qry=3, task_num=3, update_step=3,
                              update_step_test=4, method='G-Meta')
    args.pruned_forward = pruned
    args.device_finish = device_finish
    torch.manual_seed(222)
    m = Meta(args, cfg).to(U.dev())
    assert m.device_finish == device_finish
    m.return_meta_grad = True
    _nonzero_biases(m.net.parameters())
    params = [p.detach().cpu().clone().requires_grad_(True) for p in m.net.parameters()]
    om = O.OracleMeta(args, cfg, params=params)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    og = lambda xs_: [H.to_ograph(x) for x in xs_]                                                       # noqa: E731
    # finetunning first (it must leave the net untouched, meta.py:181), one task like train.py:125-129
    one = tuple([v[1]] for v in mb)
    want_f = om.finetunning(og(one[0]), one[1], og(one[2]), one[3], *one[4:], feats)
    got_f = m.finetunning(*one, feats)
    np.testing.assert_allclose(got_f, want_f, atol=1e-6)
    want = om.forward(og(xs), ys, og(xq), yq, cs, cq, ns, nq, gs, gq, feats)
    accs = m(*mb, feats)
    np.testing.assert_allclose(accs, want, atol=1e-6)
    assert abs(m.last["loss_q"] - om.last_loss_q) < 1e-4
    for k, (g, r) in enumerate(zip(m.last["meta_grad"], om.last_grads)):
        U.report("meta-grad[%d]" % k, g, r, 2e-5 + 1e-4 * float(r.abs().max()), 1e-3)
    # the second step carries the Adam state
    want2 = om.forward(og(xs), ys, og(xq), yq, cs, cq, ns, nq, gs, gq, feats)
    np.testing.assert_allclose(m(*mb, feats), want2, atol=1e-6)


@pytest.mark.parametrize("graphs", [True, False])
def test_nan_loss_skips_the_outer_update_like_the_reference(graphs):
    """meta.py:163-164: a NaN mean query loss leaves the parameters and the optimiser alone.  A meta-batch with one NaN
    feature row makes the loss NaN: parameters bit-identical afterwards, Adam's step count not advanced, and the next
    clean step equals the oracle's FIRST step (same accuracies, same gradient) -- with and without CUDA graphs."""
    from gmeta_b200.meta import Meta
    mb, feats = _ragged_meta_batch(77, False)
    cfg = [('GraphConv', [12, 16]), ('GraphConv', [16, 16]), ('Linear', [16, 3])]
    args = argparse.Namespace(update_lr=0.05, meta_lr=1e-3, n_way=3, k_spt=2, k_qry=3, task_num=3, update_step=3,
                              update_step_test=4, method='G-Meta', use_graphs=graphs, graph_host_batches=graphs)
    torch.manual_seed(222)
    m = Meta(args, cfg).to(U.dev())
    m.return_meta_grad = True
    _nonzero_biases(m.net.parameters())
    params = [p.detach().cpu().clone().requires_grad_(True) for p in m.net.parameters()]
    before = [p.detach().clone() for p in m.net.parameters()]
    bad = [f.copy() for f in feats]
    for f in bad:
        f[:] = np.nan                              # whichever rows the tasks touch
    accs_bad = m(*mb, bad)
    assert accs_bad.shape == (args.update_step + 1,)
    assert m.last["skipped"] and np.isnan(m.last["loss_q"])
    for a, b in zip(before, m.net.parameters()):
        assert torch.equal(a, b)
    assert m.meta_optim.step_count == 0
    om = O.OracleMeta(args, cfg, params=params)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    og = lambda xs_: [H.to_ograph(x) for x in xs_]                                                       # noqa: E731
    want = om.forward(og(xs), ys, og(xq), yq, cs, cq, ns, nq, gs, gq, feats)
    accs = m(*mb, feats)                           # new feature arrays: the device table is rebuilt from them
    np.testing.assert_allclose(accs, want, atol=1e-6)
    assert not m.last["skipped"] and m.meta_optim.step_count == 1
    for k, (g, r) in enumerate(zip(m.last["meta_grad"], om.last_grads)):
        U.report("meta-grad[%d]" % k, g, r, 2e-5 + 1e-4 * float(r.abs().max()), 1e-3)
    want2 = om.forward(og(xs), ys, og(xq), yq, cs, cq, ns, nq, gs, gq, feats)
    np.testing.assert_allclose(m(*mb, feats), want2, atol=1e-6)       # Adam state: one applied step on both sides

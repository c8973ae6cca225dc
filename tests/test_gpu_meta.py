"""GPU parity of the drop-in surface (Meta / Classifier / proto losses) against the golden
outputs of the unmodified reference and against the oracle on fresh inputs.

Stated tolerance (BASELINE.json north_star): logits within 1e-4 (fp32) of the reference CPU
path and identical argmax -> identical accuracy vectors.
"""
import copy
import os

import numpy as np
import pytest
import torch

from gmeta_b200 import _lib
from oracle import gmeta_oracle as O
from tests import gpu_util as U
from tests import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOGIT_TOL = 1e-4


def _meta_from_golden(kind, impl=_lib.IMPL_AUTO, dense_backward=False):
    from gmeta_b200.meta import Meta
    d = np.load(os.path.join(GOLD, "meta_%s.npz" % kind))
    ds = H.tiny_dataset(kind)
    args = ds.args()
    args.impl = impl
    args.dense_backward = dense_backward
    m = Meta(args, ds.config()).to(U.dev())
    with torch.no_grad():
        for k, p in enumerate(m.net.parameters()):
            p.copy_(torch.tensor(d['p0_%d' % k]))
    feats = [d['feat%d' % g] for g in range(int(d['n_graphs']))]
    return m, d, H.unpack_meta_batch(d), feats


@pytest.mark.parametrize("impl", [_lib.IMPL_AUTO, _lib.IMPL_SIMT])
@pytest.mark.parametrize("dense_backward", [False, True])
@pytest.mark.parametrize("kind", H.TINY_KINDS)
def test_meta_forward_matches_reference_golden(kind, dense_backward, impl):
    """Meta.forward == the reference's: accuracy vector identical, meta-gradient and step-0 support
    logits within tolerance, second step (Adam state carried) accuracies identical.  Both backward
    formulations (every row like autograd / structurally non-zero rows only) and both layer
    implementations (auto = tcgen05 where the shape allows / FFMA)."""
    m, d, mb, feats = _meta_from_golden(kind, impl, dense_backward)
    m.return_meta_grad = True
    m.keep_logits_spt0 = True
    accs = m(*mb, feats)
    assert accs.shape == d['accs'].shape
    n0 = mb[0][0].batch_num_nodes
    U.report("%s logits spt task0" % kind, m.last["logits_spt0"][:len(n0)], d['logits_spt0'], LOGIT_TOL)
    np.testing.assert_allclose(accs, d['accs'], atol=1e-6, err_msg="%s accs" % kind)
    for k, g in enumerate(m.last["meta_grad"]):
        ref = d['grad_%d' % k]
        U.report("%s meta-grad[%d]" % (kind, k), g, ref, 2e-5 + 1e-4 * float(np.abs(ref).max()), 1e-3)
    assert not m.last["skipped"]
    assert m.last["gpu_launches"] > 0
    accs2 = m(*mb, feats)
    np.testing.assert_allclose(accs2, d['accs2'], atol=1e-6, err_msg="%s accs step 2" % kind)


@pytest.mark.parametrize("kind", H.TINY_KINDS)
def test_finetunning_matches_reference_golden(kind):
    m, d, mb, feats = _meta_from_golden(kind)
    before = [p.detach().clone() for p in m.net.parameters()]
    fin = m.finetunning(*mb, feats)
    np.testing.assert_allclose(fin, d['finetune_accs'], atol=1e-6)
    for a, b in zip(before, m.net.parameters()):
        assert torch.equal(a, b)          # finetunning must not touch self.net (meta.py:181)


@pytest.mark.parametrize("kind", ['disjoint', 'link', 'deep'])
def test_classifier_autograd_path_matches_oracle(kind):
    """Classifier.forward + torch.autograd.grad (the call pattern of meta.py:122-126) + the public
    proto_loss_spt on the device vs the oracle on CPU."""
    from gmeta_b200.learner import Classifier
    from gmeta_b200.meta import proto_loss_qry, proto_loss_spt
    ds = H.tiny_dataset(kind)
    rng = np.random.default_rng(5)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = ds.sample_task(rng)
    torch.manual_seed(3)
    net = Classifier(ds.config()).to(U.dev())
    ref_vars = [p.detach().cpu().clone().requires_grad_(True) for p in net.parameters()]
    feat_s = O.gather_features(ds.feats, gs, ns)
    feat_q = O.gather_features(ds.feats, gq, nq)

    logits, _ = net(xs, cs, feat_s)
    loss, acc, protos = proto_loss_spt(logits, ys, ds.k_spt)
    grad = torch.autograd.grad(loss, net.parameters(), retain_graph=True)
    fast = [p - 0.05 * g for p, g in zip(net.parameters(), grad)]
    lq, aq = proto_loss_qry(net(xq, cq, feat_q, fast)[0], yq, protos)
    gq2 = torch.autograd.grad(lq, net.parameters())            # flows through fast AND the prototypes

    ol = O.classifier_forward(ds.config(), ref_vars, H.to_ograph(xs), cs, feat_s)
    oloss, oacc, oprotos = O.proto_loss_spt(ol, ys, ds.k_spt)
    ograd = torch.autograd.grad(oloss, ref_vars, retain_graph=True)
    ofast = [p - 0.05 * g for p, g in zip(ref_vars, ograd)]
    olq, oaq = O.proto_loss_qry(O.classifier_forward(ds.config(), ofast, H.to_ograph(xq), cq, feat_q), yq, oprotos)
    ogq2 = torch.autograd.grad(olq, ref_vars)

    U.report("logits", logits, ol.detach(), LOGIT_TOL)
    assert torch.equal(logits.argmax(1).cpu(), ol.argmax(1))
    U.report("loss_s", loss, oloss.detach(), 1e-5, 1e-5)
    assert abs(float(acc) - float(oacc)) < 1e-6 and abs(float(aq) - float(oaq)) < 1e-6
    U.report("loss_q", lq, olq.detach(), 1e-5, 1e-5)
    for k, (a, b) in enumerate(zip(grad, ograd)):
        U.report("inner grad[%d]" % k, a, b, 2e-5 + 1e-4 * float(b.abs().max()), 1e-3)
    for k, (a, b) in enumerate(zip(gq2, ogq2)):
        U.report("outer grad[%d]" % k, a, b, 2e-5 + 1e-4 * float(b.abs().max()), 1e-3)


def test_meta_forward_matches_oracle_midsize():
    """A C1-shaped batch (hidden 64, 128-d features, 4 tasks) vs the oracle on the same inputs:
    accuracies identical, loss within 1e-4, meta-gradient within tolerance."""
    from gmeta_b200.meta import Meta
    from gmeta_b200.synthetic import make_dataset
    ds = make_dataset('C1', scale=0.3)
    ds.update_lr = 0.05
    rng = np.random.default_rng(17)
    mb = ds.sample_meta_batch(rng, 4)
    torch.manual_seed(222)
    m = Meta(ds.args(), ds.config()).to(U.dev())
    m.return_meta_grad = True
    params = [p.detach().cpu().clone().requires_grad_(True) for p in m.net.parameters()]
    om = O.OracleMeta(ds.args(), ds.config(), params=params)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    want = om.forward([H.to_ograph(x) for x in xs], ys, [H.to_ograph(x) for x in xq], yq, cs, cq, ns, nq, gs, gq, ds.feats)
    accs = m(*mb, ds.feats)
    np.testing.assert_allclose(accs, want, atol=1e-6)
    assert abs(m.last["loss_q"] - om.last_loss_q) < 1e-4
    for k, (g, r) in enumerate(zip(m.last["meta_grad"], om.last_grads)):
        U.report("meta-grad[%d]" % k, g, r, 2e-5 + 1e-4 * float(r.abs().max()), 1e-3)


def test_pruned_forward_equals_full_forward():
    """Exact shortcut: computing only the rows the read-out depends on (layer l over its
    active rows, compact activations) gives the same accuracies, losses and meta-gradient as computing
    every row of every layer like the reference -- per-row arithmetic is identical, so the results
    agree to fp32 summation-order noise of the weight-gradient reductions."""
    from gmeta_b200.meta import Meta
    from gmeta_b200.synthetic import make_dataset
    # strict: both modes on the same layer kernel (FFMA).  AUTO: the full formulation takes the CTA-pair
    # path where the shape allows while the pruned mode does not; the two contractions round differently at the
    # 1e-6 level, which can flip the ReLU of a pre-activation that is ~0 and switch one row's contribution
    # (~1e-5) to a bias-gradient entry on or off -- hence the looser bound there.
    for impl, g_atol in ((_lib.IMPL_SIMT, 0.0), (_lib.IMPL_AUTO, 3e-5)):
        for name, scale, tasks in (('C1', 0.3, 4), ('C4', 1.0, 3)):
            ds = make_dataset(name, scale=scale)
            mb = ds.sample_meta_batch(np.random.default_rng(5), tasks)
            outs = []
            for pruned in (False, True):
                args = ds.args()
                args.pruned_forward = pruned
                args.impl = impl
                torch.manual_seed(222)
                m = Meta(args, ds.config()).to(U.dev())
                m.return_meta_grad = True
                # non-zero biases: with the reference's zero biases the last layer's pre-activation of a centre without
                # in-edges is the rounding noise of a mathematically-zero bias gradient (translation-invariant loss),
                # and its ReLU flips with the summation order (see tests/test_gpu_meta_configs.py)
                gen = torch.Generator().manual_seed(11)
                with torch.no_grad():
                    for prm in m.net.parameters():
                        if prm.dim() == 1:
                            prm.copy_((0.02 * torch.randn(prm.shape, generator=gen)).to(prm.device))
                accs = m(*mb, ds.feats)
                outs.append((accs, m.last["loss_q"], [g.clone() for g in m.last["meta_grad"]], m.last["gpu_launches"]))
            np.testing.assert_allclose(outs[0][0], outs[1][0], atol=1e-6, err_msg=name)
            assert abs(outs[0][1] - outs[1][1]) < (1e-6 if g_atol == 0.0 else 1e-5), name
            for k, (a, b) in enumerate(zip(outs[0][2], outs[1][2])):
                U.report("%s grad[%d] full vs pruned" % (name, k), b, a, g_atol + 1e-7 + 1e-5 * float(a.abs().max()), 1e-4)


def test_task_batching_equals_task_by_task():
    """Size-independent property: packing T tasks into one launch == running them one at a time
    (tasks only interact through the final sum, meta.py:155,161)."""
    from gmeta_b200.meta import Meta
    ds = H.tiny_dataset('disjoint')
    rng = np.random.default_rng(23)
    mb = ds.sample_meta_batch(rng, 4)
    torch.manual_seed(1)
    m = Meta(ds.args(), ds.config()).to(U.dev())
    m.return_meta_grad = True
    solo = copy.deepcopy(m)
    accs = m(*mb, ds.feats)
    g_all = [g.clone() for g in m.last["meta_grad"]]
    acc_sum, g_sum = 0, None
    for t in range(4):
        one = copy.deepcopy(solo)
        one.return_meta_grad = True
        a = one(*[[lst[t]] for lst in mb], ds.feats)
        acc_sum = acc_sum + a
        g = one.last["meta_grad"]
        g_sum = g if g_sum is None else [x + y for x, y in zip(g_sum, g)]
    np.testing.assert_allclose(accs, acc_sum / 4, atol=1e-6)
    for k, (a, b) in enumerate(zip(g_all, g_sum)):
        U.report("grad[%d] batched vs per-task" % k, a, b / 4, 1e-6, 1e-5)


def test_subgraph_order_and_node_permutation_equivariance():
    """Permuting the subgraphs of a batch permutes the logits; relabelling nodes inside a
    subgraph leaves its logit row unchanged (up to fp32 summation order)."""
    from gmeta_b200.learner import Classifier
    from gmeta_b200.packed import PackedSubgraphBatch, SubgraphCSR
    ds = H.tiny_dataset('disjoint')
    rng = np.random.default_rng(31)
    subs = [ds.subgraph(0, int(v)) for v in rng.choice(ds.graphs[0].n, 6, replace=False)]
    torch.manual_seed(5)
    net = Classifier(ds.config()).to(U.dev())

    def run(ss):
        g = PackedSubgraphBatch.batch(ss)
        feat = torch.tensor(np.vstack([ds.feats[0][s.parent_nid] for s in ss]))
        return net(g, torch.LongTensor([s.centre for s in ss]), feat)[0].detach().cpu()

    base = run(subs)
    perm = rng.permutation(6)
    U.report("subgraph permutation", run([subs[i] for i in perm]), base[perm], 1e-6)
    shuffled = []
    for s in subs:
        pi = rng.permutation(s.n)                      # new id of old node v is pi[v]
        dst = np.repeat(np.arange(s.n), np.diff(s.indptr))
        parent = np.empty(s.n, dtype=np.int64)
        parent[pi] = s.parent_nid
        shuffled.append(SubgraphCSR.from_edges(pi[s.indices], pi[dst], s.n, parent, int(pi[s.centre])))
    U.report("node relabelling", run(shuffled), base, 2e-6)


def test_two_streams_and_graph_replay_equal_the_serial_eager_step():
    """The schedule is an implementation detail: query forwards on the auxiliary stream and CUDA-graph replay of a
    device-resident batch give bit-identical step outputs, weights and Adam state to the one-stream eager step."""
    from gmeta_b200.meta import Meta
    from gmeta_b200.synthetic import make_dataset
    ds = make_dataset('C1', scale=0.3)
    ds.update_lr = 0.05
    mb = ds.sample_meta_batch(np.random.default_rng(3), 4)
    results = []
    for two, graphs in ((False, False), (True, False), (True, True)):
        args = ds.args()
        args.two_streams, args.use_graphs = two, graphs
        torch.manual_seed(222)
        m = Meta(args, ds.config()).to(U.dev())
        db = m.upload_batch(mb, ds.feats, own_buffer=True)
        outs = [m.step_device(db).clone() for _ in range(4)]       # eager, capture, replay, replay
        if graphs:
            assert any(v[0] is not None for v in m._graphs.values()), "the step must have been captured"
        torch.cuda.synchronize()
        results.append((torch.stack(outs).cpu(), [p.detach().cpu().clone() for p in m.net.parameters()],
                        m.meta_optim.step_count))
    for r in results[1:]:
        assert torch.equal(r[0], results[0][0])
        assert r[2] == results[0][2] == 4
        for a, b in zip(r[1], results[0][1]):
            assert torch.equal(a, b)


def test_prefetch_pipeline_equals_plain_forward():
    """Every way a host batch can reach the step gives exactly the accuracies, losses and weights of plain eager
    forward calls: the step as an updatable CUDA graph (prepared at the step, or one batch ahead while the previous step
    runs), Meta.prefetch with a one- or three-batch lookahead (packing and H2D on worker threads / the copy stream),
    the slim host packing with the device passes behind the copy; a prefetched batch that is not the one forward() then
    receives is dropped safely."""
    from gmeta_b200.meta import Meta
    ds = H.tiny_dataset('shared')
    rng = np.random.default_rng(12)
    mbs = [ds.sample_meta_batch(rng) for _ in range(7)]
    outs = []
    for mode in ("eager", "graph", "prefetch", "lookahead3", "lookahead3+device_finish", "mismatch"):
        torch.manual_seed(9)
        m = Meta(ds.args(), ds.config()).to(U.dev())
        m.graph_host_batches = mode != "eager"
        m.device_finish = mode.endswith("device_finish")
        accs = []
        for i, mb in enumerate(mbs):
            if mode == "prefetch" and i + 1 < len(mbs):
                m.prefetch(*mbs[i + 1], ds.feats)
            if mode.startswith("lookahead3"):
                for j in ((1, 2, 3) if i == 0 else (3,)):
                    if i + j < len(mbs):
                        m.prefetch(*mbs[i + j], ds.feats)
            if mode == "mismatch":
                m.prefetch(*mbs[(i + 2) % len(mbs)], ds.feats)         # not the batch that comes next
            accs.append((m(*mb, ds.feats), m.last["loss_q"]))
            assert m.last["gpu_launches"] > 0
        if mode.startswith("lookahead3"):
            assert m._step_graphs is not None                           # the steps ran as graphs prepared ahead
        outs.append((accs, [p.detach().cpu().clone() for p in m.net.parameters()]))
    for o in outs[1:]:
        for (a, la), (b, lb) in zip(o[0], outs[0][0]):
            assert np.array_equal(a, b) and la == lb
        for p, q in zip(o[1], outs[0][1]):
            assert torch.equal(p, q)


def test_step_graph_survives_topology_changes():
    """Host batches whose launch topology differs from step to step (task counts 3 / 1 / 2: a different number of tiles,
    split and unsplit weight gradients, empty launches dropped) go through the updatable step graph -- updated in place
    when the topology allows, rebuilt otherwise -- and give exactly the eager results."""
    import ctypes as C
    from gmeta_b200.meta import Meta
    ds = H.tiny_dataset('disjoint')
    rng = np.random.default_rng(21)
    mbs = [ds.sample_meta_batch(rng, t) for t in (3, 1, 2, 3, 3, 1, 2, 2)]
    outs = []
    for graphs in (False, True):
        torch.manual_seed(9)
        m = Meta(ds.args(), ds.config()).to(U.dev())
        m.graph_host_batches = graphs
        accs = []
        for i, mb in enumerate(mbs):
            m.global_task_num = len(mb[0])
            for j in ((1, 2) if i == 0 else (2,)):
                if graphs and i + j < len(mbs):
                    m.prefetch(*mbs[i + j], ds.feats)
            accs.append((m(*mb, ds.feats), m.last["loss_q"]))
        outs.append((accs, [p.detach().cpu().clone() for p in m.net.parameters()]))
        if graphs:
            ups = inst = 0
            for h in m._step_graphs:
                u, k = C.c_int32(), C.c_int32()
                _lib.check(_lib.lib().gmeta_step_graph_stats(h, C.byref(u), C.byref(k)))
                ups, inst = ups + u.value, inst + k.value
            assert ups + inst >= len(mbs) and inst >= 1
    for (a, la), (b, lb) in zip(outs[1][0], outs[0][0]):
        assert np.array_equal(a, b) and la == lb
    for p, q in zip(outs[1][1], outs[0][1]):
        assert torch.equal(p, q)


def test_parameters_alias_the_flat_buffer_and_survive_deepcopy():
    """The net's parameters are views of the flat theta buffer the kernels update in place; deepcopy (train.py:87,127)
    and in-place edits keep working, and a copy trains independently of the original."""
    from gmeta_b200.meta import Meta
    ds = H.tiny_dataset('disjoint')
    mb = ds.sample_meta_batch(np.random.default_rng(2))
    torch.manual_seed(4)
    m = Meta(ds.args(), ds.config()).to(U.dev())
    m(*mb, ds.feats)
    flat = m._theta_flat
    for p, off in zip(m.net.parameters(), m.spec.offsets):
        assert p.data_ptr() == flat.data_ptr() + 4 * off
    twin = copy.deepcopy(m)
    w_before = [p.detach().clone() for p in m.net.parameters()]
    a1 = twin(*mb, ds.feats)
    for p, q in zip(m.net.parameters(), w_before):
        assert torch.equal(p, q)                                   # training the copy leaves the original alone
    a2 = m(*mb, ds.feats)
    np.testing.assert_allclose(a1, a2, atol=0)
    for p, q in zip(m.net.parameters(), twin.net.parameters()):
        assert torch.equal(p, q)
    with torch.no_grad():
        for p in m.net.parameters():
            p.mul_(0.5)                                            # in-place edits reach the kernels
    a3 = m(*mb, ds.feats)
    assert a3.shape == a2.shape


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgmeta_b200.so")
    with pytest.raises(_lib.GMetaError):
        _lib.lib()


def test_update_step_one_is_rejected_like_the_reference():
    from gmeta_b200.meta import Meta
    ds = H.tiny_dataset('disjoint')
    ds.update_step = 1
    mb = ds.sample_meta_batch(np.random.default_rng(1), 2)
    m = Meta(ds.args(), ds.config()).to(U.dev())
    with pytest.raises(RuntimeError):
        m(*mb, ds.feats)

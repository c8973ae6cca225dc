"""ORACLE / TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/G-Meta/learner.py, meta.py via oracle/ref_loader.py + DGL stand-in)
on small seeded inputs.  Run in the build container (the reference tree is not on the
GPU box):      python -m oracle.make_golden
Single-threaded on purpose: torch's multi-threaded index_add makes the reference
itself non-reproducible in the last bits.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def to_dgl(dgl, p):
    s, d = p.edges()
    return dgl.DGLGraph(s, d, p.n_nodes, batch_num_nodes=p.batch_num_nodes)


def kat(learner, meta, dgl, F):
    """SURVEY Appendix D known-answer vectors, recomputed with the reference's functions."""
    g = dgl.DGLGraph([0, 1, 1, 2], [1, 0, 2, 1], 3)
    out = {}
    b = torch.tensor([0.5, -0.5])
    x = torch.tensor([[1., 0.], [0., 1.], [1., 1.]])
    w = torch.tensor([[1., 2.], [3., 4.]])
    out['d1'] = learner.GraphConv(2, 2, activation=F.relu)(g, x, w, b).numpy()
    x = torch.tensor([[1., 0., 2.], [0., 1., -1.], [1., 1., 0.]])
    w = torch.tensor([[1., 2.], [3., 4.], [-1., 0.5]])
    out['d1b'] = learner.GraphConv(3, 2, activation=F.relu)(g, x, w, b).numpy()
    z = torch.tensor([[0., 0.], [2., 0.], [0., 2.], [0., 4.]], requires_grad=True)
    y = torch.LongTensor([0, 0, 1, 1])
    loss, acc, protos = meta.proto_loss_spt(z, y, 2)
    out['d2_loss'], out['d2_acc'], out['d2_protos'] = loss.detach().numpy(), acc.numpy(), protos.detach().numpy()
    out['d2_dlogits'] = torch.autograd.grad(loss, z)[0].numpy()
    q = torch.tensor([[1., 1.], [0., 3.], [3., 0.], [1., 2.]])
    lq, aq = meta.proto_loss_qry(q, torch.LongTensor([0, 1, 0, 1]), protos.detach())
    out['d2q_loss'], out['d2q_acc'] = lq.numpy(), aq.numpy()
    np.savez(os.path.join(OUT, "kat.npz"), **out)


def layer_cases(learner, dgl, F):
    """GraphConv.forward + autograd on random multigraphs: both branch orders, zero-in-degree
    nodes, multi-edges, directed edges, non-multiple-of-4 widths."""
    rng = np.random.default_rng(11)
    out = {}
    for k, (n, e, fi, fo) in enumerate([(37, 90, 16, 32), (64, 200, 48, 24), (50, 60, 5, 16), (20, 300, 1, 8),
                                        (130, 700, 32, 32)]):
        src = rng.integers(0, n, size=e)
        dst = rng.integers(0, max(1, n - 3), size=e)   # last 3 nodes have in-degree 0
        src[:5], dst[:5] = src[5:10], dst[5:10]         # guaranteed multi-edges
        g = dgl.DGLGraph(src, dst, n)
        x = torch.tensor(rng.standard_normal((n, fi), dtype=np.float32), requires_grad=True)
        w = torch.tensor(rng.standard_normal((fi, fo), dtype=np.float32) * 0.3, requires_grad=True)
        b = torch.tensor(rng.standard_normal(fo, dtype=np.float32) * 0.1, requires_grad=True)
        y = learner.GraphConv(fi, fo, activation=F.relu)(g, x, w, b)
        gy = torch.tensor(rng.standard_normal((n, fo), dtype=np.float32))
        dx, dw, db = torch.autograd.grad(y, (x, w, b), gy)
        p = 'c%d_' % k
        out.update({p + 'src': src, p + 'dst': dst, p + 'n': np.array(n), p + 'x': x.detach().numpy(),
                    p + 'w': w.detach().numpy(), p + 'b': b.detach().numpy(), p + 'y': y.detach().numpy(),
                    p + 'gy': gy.numpy(), p + 'dx': dx.numpy(), p + 'dw': dw.numpy(), p + 'db': db.numpy()})
    out['n_cases'] = np.array(5)
    np.savez(os.path.join(OUT, "layer_cases.npz"), **out)


def loss_cases(meta):
    rng = np.random.default_rng(12)
    out = {}
    cases = [(3, 3, 24, 3), (2, 3, 10, 2), (5, 1, 4, 7), (4, 2, 6, 30)]   # (n_cls, k_spt, k_qry, logit dim)
    for k, (ncls, ks, kq, d) in enumerate(cases):
        labels = rng.choice(50, ncls, replace=False)
        ys = rng.permutation(np.repeat(labels, ks))
        yq = rng.permutation(np.repeat(labels, kq))
        zs = torch.tensor(rng.standard_normal((ncls * ks, d), dtype=np.float32), requires_grad=True)
        zq = torch.tensor(rng.standard_normal((ncls * kq, d), dtype=np.float32), requires_grad=True)
        ls, accs, protos = meta.proto_loss_spt(zs, torch.LongTensor(ys), ks)
        dzs = torch.autograd.grad(ls, zs, retain_graph=True)[0]
        lq, accq = meta.proto_loss_qry(zq, torch.LongTensor(yq), protos)
        dzq, dzs_via_protos = torch.autograd.grad(lq, (zq, zs))
        p = 'c%d_' % k
        out.update({p + 'ys': ys, p + 'yq': yq, p + 'zs': zs.detach().numpy(), p + 'zq': zq.detach().numpy(),
                    p + 'ks': np.array(ks), p + 'loss_s': ls.detach().numpy(), p + 'acc_s': accs.numpy(),
                    p + 'protos': protos.detach().numpy(), p + 'dzs': dzs.numpy(),
                    p + 'loss_q': lq.detach().numpy(), p + 'acc_q': accq.numpy(), p + 'dzq': dzq.numpy(),
                    p + 'dzs_via_protos': dzs_via_protos.numpy()})
    out['n_cases'] = np.array(len(cases))
    np.savez(os.path.join(OUT, "loss_cases.npz"), **out)


def meta_case(meta, dgl, kind):
    """Meta.forward / Meta.finetunning of the reference on one tiny meta-batch."""
    ds = H.tiny_dataset(kind)
    rng = np.random.default_rng(100)
    mb = ds.sample_meta_batch(rng)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    cfg, args = ds.config(), ds.args()
    torch.manual_seed(222)
    m = meta.Meta(args, cfg)
    params0 = [p.detach().clone().numpy() for p in m.net.parameters()]
    dxs, dxq = [to_dgl(dgl, x) for x in xs], [to_dgl(dgl, x) for x in xq]
    # logits of the plain Classifier.forward on task 0's support set (learner.py:134-194)
    feat_s0 = torch.Tensor(np.vstack([ds.feats[gs[0][j]][np.array(x)] for j, x in enumerate(ns[0])]))
    logits0 = m.net(dxs[0], cs[0], feat_s0)[0].detach().numpy()
    fin = m.finetunning(dxs, ys, dxq, yq, cs, cq, ns, nq, gs, gq, ds.feats)
    accs = m(dxs, ys, dxq, yq, cs, cq, ns, nq, gs, gq, ds.feats)
    grads = [p.grad.detach().clone().numpy() for p in m.net.parameters()]
    params1 = [p.detach().clone().numpy() for p in m.net.parameters()]
    accs2 = m(dxs, ys, dxq, yq, cs, cq, ns, nq, gs, gq, ds.feats)     # second step: Adam state carried
    out = H.pack_meta_batch(mb)
    out.update({'n_graphs': np.array(len(ds.feats)), 'n_params': np.array(len(params0)),
                'accs': np.asarray(accs, dtype=np.float64), 'accs2': np.asarray(accs2, dtype=np.float64),
                'finetune_accs': np.asarray(fin, dtype=np.float64), 'logits_spt0': logits0})
    for g, f in enumerate(ds.feats):
        out['feat%d' % g] = f
    for k in range(len(params0)):
        out['p0_%d' % k], out['grad_%d' % k], out['p1_%d' % k] = params0[k], grads[k], params1[k]
    np.savez_compressed(os.path.join(OUT, "meta_%s.npz" % kind), **out)


def main():
    torch.set_num_threads(1)
    learner, meta, _ = ref_loader.load()
    dgl = ref_loader.shim_dgl()
    import torch.nn.functional as F
    os.makedirs(OUT, exist_ok=True)
    kat(learner, meta, dgl, F)
    layer_cases(learner, dgl, F)
    loss_cases(meta)
    for kind in H.TINY_KINDS:
        meta_case(meta, dgl, kind)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

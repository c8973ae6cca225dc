"""Shared test helpers: meta-batch (de)serialisation for golden fixtures and adapters
from the product's PackedSubgraphBatch to the oracle's graph type."""
import numpy as np
import torch

from gmeta_b200.packed import PackedSubgraphBatch
from oracle import gmeta_oracle as O


def to_ograph(p):
    return O.OGraph.from_csr(p.indptr, p.indices, p.batch_num_nodes)


def tiny_dataset(kind, seed=7):
    """Small datasets covering each reference mode (node Disjoint / Shared, LinkPred,
    in>out 'matmul first' branch learner.py:34-40, h=1 and h=3)."""
    from gmeta_b200 import synthetic as S
    rng = np.random.default_rng(seed)
    if kind == 'disjoint':      # aggregate-first branch (16 <= 32), h=2
        g = S.er_graph(400, 1600, rng)
        return S._node_dataset('tiny_disjoint', [g], 16, 8, rng, task_setup='Disjoint', n_way=3, k_spt=2,
                               k_qry=4, h=2, hidden_dim=32, update_step=3, update_lr=0.05, meta_lr=1e-3,
                               task_num=3, sample_nodes=40, update_step_test=4)
    if kind == 'wide':          # matmul-first branch (48 > 24), skewed degrees, sampling cap active
        g = S.skewed_graph(500, 2500, rng)
        return S._node_dataset('tiny_wide', [g], 48, 6, rng, task_setup='Disjoint', n_way=2, k_spt=3,
                               k_qry=5, h=2, hidden_dim=24, update_step=2, update_lr=0.05, meta_lr=1e-3,
                               task_num=2, sample_nodes=30, update_step_test=3)
    if kind == 'shared':        # Shared labels, logits dim = total classes, odd feature width, h=1
        gs = [S.er_graph(150, 500, rng) for _ in range(3)]
        return S._node_dataset('tiny_shared', gs, 10, 2, rng, per_graph_classes=2, task_setup='Shared',
                               n_way=3, k_spt=3, k_qry=4, h=1, hidden_dim=20, update_step=4,
                               update_lr=0.05, meta_lr=5e-3, task_num=2, update_step_test=3)
    if kind == 'deep':          # h=3
        g = S.er_graph(300, 700, rng)
        return S._node_dataset('tiny_deep', [g], 12, 5, rng, task_setup='Disjoint', n_way=2, k_spt=2,
                               k_qry=3, h=3, hidden_dim=16, update_step=3, update_lr=0.05, meta_lr=1e-3,
                               task_num=2, sample_nodes=50, update_step_test=3)
    if kind == 'link':          # LinkPred: directed storage, pair readout, F0=5
        return S._link_dataset('tiny_link', 3, 120, 260, 5, rng, n_way=2, k_spt=4, k_qry=6, h=2,
                               hidden_dim=16, update_step=3, update_lr=0.05, meta_lr=5e-4, task_num=2,
                               sample_nodes=60, update_step_test=3)
    raise ValueError(kind)


TINY_KINDS = ['disjoint', 'wide', 'shared', 'deep', 'link']


def pack_meta_batch(mb):
    """10-tuple of lists -> flat dict of numpy arrays (npz-friendly)."""
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    d = {'task_num': np.array(len(xs))}
    for t in range(len(xs)):
        for tag, x, y, c, n, g in (('s', xs[t], ys[t], cs[t], ns[t], gs[t]), ('q', xq[t], yq[t], cq[t], nq[t], gq[t])):
            k = 't%d%s_' % (t, tag)
            d[k + 'indptr'] = x.indptr
            d[k + 'indices'] = x.indices
            d[k + 'bnn'] = np.array(x.batch_num_nodes, dtype=np.int64)
            d[k + 'y'] = y.numpy()
            d[k + 'c'] = c.numpy()
            d[k + 'nid'] = np.concatenate(n).astype(np.int64)
            d[k + 'gidx'] = np.array(g, dtype=np.int64)
    return d


def unpack_meta_batch(d):
    from gmeta_b200.packed import csr_transpose
    T = int(d['task_num'])
    out = [[] for _ in range(10)]
    for t in range(T):
        for tag, ix, iy, ic, inn, ig in (('s', 0, 1, 4, 6, 8), ('q', 2, 3, 5, 7, 9)):
            k = 't%d%s_' % (t, tag)
            indptr, indices, bnn = d[k + 'indptr'], d[k + 'indices'], d[k + 'bnn']
            n = indptr.shape[0] - 1
            tp, ti = csr_transpose(indptr, indices, n)
            out[ix].append(PackedSubgraphBatch(indptr.astype(np.int32), indices.astype(np.int32), tp, ti, bnn.tolist()))
            out[iy].append(torch.LongTensor(d[k + 'y']))
            out[ic].append(torch.LongTensor(d[k + 'c']))
            off = np.concatenate([[0], np.cumsum(bnn)])
            out[inn].append([d[k + 'nid'][off[j]:off[j + 1]] for j in range(len(bnn))])
            out[ig].append([int(v) for v in d[k + 'gidx']])
    return tuple(out)

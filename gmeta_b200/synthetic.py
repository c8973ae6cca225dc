"""Synthetic datasets + episode sampling shaped like BASELINE.json's configs.

There is no network for the real datasets, so bench.py and the tests run on
seeded synthetic graphs of the same shape.  Episodes come out in exactly the
10-tuple layout Subgraphs.__getitem__/collate hand to Meta.forward
(subgraph_data_processing.py:348-419), with a PackedSubgraphBatch where the
reference has a DGL batched graph.
"""
import numpy as np
import torch

from .packed import PackedSubgraphBatch
from .subgraphs import ParentGraph, extract_subgraph, extract_subgraph_link_pred


def _symmetrise(u, v, n):
    keep = u != v
    u, v = u[keep], v[keep]
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    key = np.unique(lo.astype(np.int64) * n + hi)
    lo, hi = key // n, key % n
    return np.concatenate([lo, hi]), np.concatenate([hi, lo])


def er_graph(n, m, rng):
    """Erdos-Renyi G(n, m): m undirected edges, symmetrised, deduped, no self loops."""
    u = rng.integers(0, n, size=m)
    v = rng.integers(0, n, size=m)
    src, dst = _symmetrise(u, v, n)
    return ParentGraph.from_edges(src, dst, n)


def skewed_graph(n, m, rng):
    """Heavy-tailed degrees: one endpoint uniform, the other ~ Pareto(1.5) weights."""
    w = rng.pareto(1.5, size=n) + 1.0
    p = w / w.sum()
    u = rng.integers(0, n, size=m)
    v = rng.choice(n, size=m, p=p)
    src, dst = _symmetrise(u, v, n)
    return ParentGraph.from_edges(src, dst, n)


def directed_link_graph(n, m, rng):
    """Edge i->j stored once with i<j, as link_process.py:45-47,83-85 stores them."""
    u = rng.integers(0, n, size=m)
    v = rng.integers(0, n, size=m)
    keep = u != v
    lo, hi = np.minimum(u[keep], v[keep]), np.maximum(u[keep], v[keep])
    key = np.unique(lo.astype(np.int64) * n + hi)
    return ParentGraph.from_edges(key // n, key % n, n), key // n, key % n


class SyntheticDataset(object):
    """graphs + features + labelled items, and the task hyper-parameters of one config."""

    def __init__(self, name, graphs, feats, task_setup, link_pred, n_way, k_spt, k_qry, h,
                 hidden_dim, update_step, update_lr, meta_lr, task_num, sample_nodes=1000,
                 update_step_test=10):
        self.name = name
        self.graphs = graphs
        self.feats = feats
        self.task_setup = task_setup
        self.link_pred = link_pred
        self.n_way, self.k_spt, self.k_qry, self.h = n_way, k_spt, k_qry, h
        self.hidden_dim, self.update_step, self.update_lr = hidden_dim, update_step, update_lr
        self.meta_lr, self.task_num, self.sample_nodes = meta_lr, task_num, sample_nodes
        self.update_step_test = update_step_test
        self.node_labels = None   # node classification: list (per graph) of int label arrays
        self.pairs = None         # link prediction: list (per graph) of (i, j, label) arrays
        self.labels_num = None
        self._memo = {}

    def config(self):
        """Model topology exactly as train.py:67-75 builds it."""
        cfg = [('GraphConv', [self.feats[0].shape[1], self.hidden_dim])]
        if self.h > 1:
            cfg = cfg + [('GraphConv', [self.hidden_dim, self.hidden_dim])] * (self.h - 1)
        cfg = cfg + [('Linear', [self.hidden_dim, self.labels_num])]
        if self.link_pred:
            cfg.append(('LinkPred', [True]))
        return cfg

    def args(self):
        import argparse
        return argparse.Namespace(
            update_lr=self.update_lr, meta_lr=self.meta_lr, n_way=self.n_way, k_spt=self.k_spt,
            k_qry=self.k_qry, task_num=self.task_num, update_step=self.update_step,
            update_step_test=self.update_step_test, method='G-Meta')

    # ---- subgraph memo (subgraph_data_processing.py:296-297,319) ----
    def subgraph(self, g, i, j=None, rng=np.random):
        key = (g, i, j)
        s = self._memo.get(key)
        if s is None:
            if j is None:
                s = extract_subgraph(self.graphs[g], i, self.h, self.sample_nodes, rng)
            else:
                s = extract_subgraph_link_pred(self.graphs[g], i, j, self.sample_nodes, rng)
            self._memo[key] = s
        return s

    # ---- one episode in Subgraphs.__getitem__ layout ----
    def sample_task(self, rng):
        if self.link_pred:
            g = int(rng.integers(len(self.graphs)))
            pi, pj, pl, is_spt = self.pairs[g]
            spt_items, qry_items = [], []
            for cls in rng.permutation(2):
                cand = np.nonzero((pl == cls) & is_spt)[0]
                spt_items += [(g, int(pi[k]), int(pj[k]), int(cls)) for k in rng.choice(cand, self.k_spt, replace=False)]
            for cls in rng.permutation(2):
                cand = np.nonzero((pl == cls) & ~is_spt)[0]
                qry_items += [(g, int(pi[k]), int(pj[k]), int(cls)) for k in rng.choice(cand, self.k_qry, replace=False)]
            relabel = None
        elif self.task_setup == 'Disjoint':
            lab = self.node_labels[0]
            classes = rng.choice(self.n_classes_total, self.n_way, replace=False)
            spt_items, qry_items = [], []
            for cls in classes:
                cand = self._by_class[int(cls)]
                pick = rng.choice(cand, self.k_spt + self.k_qry, replace=False)
                spt_items += [(0, int(v), None, int(cls)) for v in pick[:self.k_spt]]
                qry_items += [(0, int(v), None, int(cls)) for v in pick[self.k_spt:]]
            order = rng.permutation(np.unique([c for _, _, _, c in spt_items]))   # :390-397
            relabel = {int(l): k for k, l in enumerate(order)}
        else:  # Shared, node classification: one graph, all of its classes (:198-217)
            g = int(rng.integers(len(self.graphs)))
            lab = self.node_labels[g]
            spt_items, qry_items = [], []
            for cls in rng.permutation(np.unique(lab)):
                cand = np.nonzero(lab == cls)[0]
                pick = rng.choice(cand, self.k_spt + self.k_qry, replace=False)
                spt_items += [(g, int(v), None, int(cls)) for v in pick[:self.k_spt]]
                qry_items += [(g, int(v), None, int(cls)) for v in pick[self.k_spt:]]
            relabel = None

        def build(items):
            subs = [self.subgraph(g, i, j, rng) for g, i, j, _ in items]
            y = np.array([c if relabel is None else relabel[c] for _, _, _, c in items], dtype=np.int64)
            centre = np.array([s.centre for s in subs], dtype=np.int64)
            pb = PackedSubgraphBatch.batch(subs)
            return (pb, torch.LongTensor(y), torch.LongTensor(centre), pb.parent_id_lists, [g for g, _, _, _ in items])

        xs, ys, cs, ns, gs = build(spt_items)
        xq, yq, cq, nq, gq = build(qry_items)
        return xs, ys, xq, yq, cs, cq, ns, nq, gs, gq

    def sample_meta_batch(self, rng, task_num=None):
        """`collate` of task_num episodes (subgraph_data_processing.py:414-419)."""
        tasks = [self.sample_task(rng) for _ in range(task_num or self.task_num)]
        return tuple(map(list, zip(*tasks)))


def _node_dataset(name, graphs, f0, n_classes, rng, per_graph_classes=None, **kw):
    feats = [rng.standard_normal((g.n, f0), dtype=np.float32) for g in graphs]
    ds = SyntheticDataset(name, graphs, feats, link_pred=False, **kw)
    if per_graph_classes is None:
        ds.node_labels = [rng.integers(0, n_classes, size=g.n) for g in graphs]
        ds.n_classes_total = n_classes
        ds._by_class = {c: np.nonzero(ds.node_labels[0] == c)[0] for c in range(n_classes)}
        ds.labels_num = kw['n_way'] if kw['task_setup'] == 'Disjoint' else n_classes   # train.py:58-61
    else:
        ds.node_labels = [rng.integers(0, per_graph_classes, size=g.n) for g in graphs]
        ds.labels_num = per_graph_classes
    return ds


def _link_dataset(name, n_graphs, n, m, f0, rng, **kw):
    graphs, pairs = [], []
    for _ in range(n_graphs):
        g_pos, pi, pj = directed_link_graph(n, m, rng)
        e = pi.shape[0]
        ni = rng.integers(0, n, size=2 * e)
        nj = rng.integers(0, n, size=2 * e)
        ok = ni != nj
        ni, nj = ni[ok][:e], nj[ok][:e]
        # negative injection, following SEAL (link_process.py:83-85): negatives become edges too
        g = ParentGraph.from_edges(np.concatenate([pi, ni]), np.concatenate([pj, nj]), n)
        ai = np.concatenate([pi, ni])
        aj = np.concatenate([pj, nj])
        al = np.concatenate([np.ones(e, dtype=np.int64), np.zeros(ni.shape[0], dtype=np.int64)])
        is_spt = rng.random(ai.shape[0]) < 0.3                       # link_process.py:13
        graphs.append(g)
        pairs.append((ai, aj, al, is_spt))
    feats = [rng.standard_normal((g.n, f0), dtype=np.float32) for g in graphs]
    ds = SyntheticDataset(name, graphs, feats, task_setup='Shared', link_pred=True, **kw)
    ds.pairs = pairs
    ds.labels_num = 2
    return ds


def make_dataset(name, seed=222, scale=1.0):
    """C1..C5 of SURVEY 8d.  `scale` < 1 shrinks node/edge counts (tests only)."""
    rng = np.random.default_rng(seed)
    s = lambda x: max(16, int(x * scale))  # noqa: E731
    if name == 'C1':
        g = er_graph(s(10000), s(50000), rng)
        return _node_dataset('C1', [g], 128, 40, rng, task_setup='Disjoint', n_way=3, k_spt=3, k_qry=24,
                             h=2, hidden_dim=64, update_step=5, update_lr=1e-3, meta_lr=1e-3, task_num=4)
    if name == 'C2':
        g = skewed_graph(s(169343), s(1166243), rng)
        return _node_dataset('C2', [g], 128, 40, rng, task_setup='Disjoint', n_way=3, k_spt=3, k_qry=24,
                             h=2, hidden_dim=256, update_step=10, update_lr=1e-2, meta_lr=1e-3, task_num=32,
                             update_step_test=20)
    if name == 'C3':
        gs = [er_graph(s(2100), s(56000), rng) for _ in range(24)]
        return _node_dataset('C3', gs, 50, 2, rng, per_graph_classes=2, task_setup='Shared', n_way=3,
                             k_spt=3, k_qry=10, h=2, hidden_dim=128, update_step=10, update_lr=1e-2,
                             meta_lr=5e-3, task_num=4)
    if name == 'C4':
        return _link_dataset('C4', 41, s(1400), s(3000), 5, rng, n_way=2, k_spt=16, k_qry=32, h=2,
                             hidden_dim=128, update_step=10, update_lr=1e-2, meta_lr=5e-4, task_num=8)
    if name == 'C5':
        n_graphs = max(4, int(1840 * min(1.0, scale * 4))) if scale < 1 else 1840
        return _link_dataset('C5', n_graphs, s(790), s(4800), 1, rng, n_way=2, k_spt=16, k_qry=16, h=2,
                             hidden_dim=256, update_step=10, update_lr=5e-3, meta_lr=5e-4, task_num=64)
    raise ValueError(name)

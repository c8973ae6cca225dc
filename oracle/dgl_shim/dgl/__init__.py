"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the product path.

Minimal pure-torch CPU stand-in for the handful of `dgl==0.4.3post2` calls the
G-Meta reference makes (requirements.txt:2).  DGL itself is a third-party
dependency that is NOT vendored under /root/reference and is not installable
here, so its published semantics are restated:

  * `graph.update_all(fn.copy_src('h','m'), fn.sum('m','h'))`  (learner.py:38-39,44-45)
        out[v] = sum over in-edges (u -> v) of h[u]; multi-edges count with
        multiplicity; zero-in-degree rows are 0.
  * `graph.in_degrees()`                                      (learner.py:29)
  * `g.batch_num_nodes` (python list), `dgl.batch(list)`      (learner.py:161, subgraph_data_processing.py:399-406)
  * `G.in_edges(v)` -> (src, dst)                             (subgraph_data_processing.py:301-333)
  * `G.subgraph(nodes)` + `.parent_nid`                       (subgraph_data_processing.py:316-317,341-342)
  * `graph.local_var()`, `.ndata[...]`, `.to(device)`         (learner.py:27,37,154; meta.py:122)

With this package first on sys.path the *unmodified* reference files
G-Meta/learner.py, G-Meta/meta.py and G-Meta/subgraph_data_processing.py import
and run on CPU (see oracle/ref_loader.py).  Only tests/, oracle/make_golden.py and
bench.py's cpu_baseline / --impl reference leg may import this.
"""
import torch

from . import function  # noqa: F401  (dgl.function as fn)

__version__ = "0.4.3post2-shim"


class DGLGraph(object):
    """Directed multigraph held as COO edge lists (src[e] -> dst[e])."""

    def __init__(self, src=None, dst=None, num_nodes=0, batch_num_nodes=None):
        self._src = torch.as_tensor([] if src is None else src, dtype=torch.int64).reshape(-1)
        self._dst = torch.as_tensor([] if dst is None else dst, dtype=torch.int64).reshape(-1)
        self._n = int(num_nodes)
        self.ndata = {}
        self.batch_num_nodes = list(batch_num_nodes) if batch_num_nodes is not None else [self._n]
        self.parent_nid = None
        self._in_ptr = None   # lazily built CSR-by-destination for in_edges()
        self._in_src = None

    # ---- construction (offline builders, link_process.py:45-47,85) ----
    def add_nodes(self, n):
        self._n += int(n)
        self.batch_num_nodes = [self._n]
        self._in_ptr = None

    def add_edges(self, u, v):
        u = torch.as_tensor(u, dtype=torch.int64).reshape(-1)
        v = torch.as_tensor(v, dtype=torch.int64).reshape(-1)
        self._src = torch.cat([self._src, u])
        self._dst = torch.cat([self._dst, v])
        self._in_ptr = None

    def add_edge(self, u, v):
        self.add_edges([u], [v])

    # ---- queries ----
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def _build_in_csr(self):
        order = torch.argsort(self._dst, stable=True)
        self._in_src = self._src[order]
        counts = torch.bincount(self._dst, minlength=self._n)
        self._in_ptr = torch.zeros(self._n + 1, dtype=torch.int64)
        self._in_ptr[1:] = torch.cumsum(counts, 0)

    def in_edges(self, v):
        if self._in_ptr is None:
            self._build_in_csr()
        v = int(v)
        lo, hi = int(self._in_ptr[v]), int(self._in_ptr[v + 1])
        src = self._in_src[lo:hi]
        return src, torch.full_like(src, v)

    def subgraph(self, nodes):
        """Node-induced subgraph; node i of the result is nodes[i] of the parent."""
        nodes = torch.as_tensor(nodes, dtype=torch.int64).reshape(-1)
        local = torch.full((self._n,), -1, dtype=torch.int64)
        local[nodes] = torch.arange(nodes.numel(), dtype=torch.int64)
        keep = (local[self._src] >= 0) & (local[self._dst] >= 0)
        sub = DGLGraph(local[self._src[keep]], local[self._dst[keep]], nodes.numel())
        sub.parent_nid = nodes.clone()
        return sub

    # ---- message passing ----
    def local_var(self):
        g = DGLGraph.__new__(DGLGraph)
        g.__dict__.update(self.__dict__)
        g.ndata = dict(self.ndata)
        return g

    def to(self, device):
        return self

    def update_all(self, message_func, reduce_func):
        assert message_func.kind == "copy_src" and reduce_func.kind == "sum"
        h = self.ndata[message_func.src]
        out = torch.zeros((self._n,) + tuple(h.shape[1:]), dtype=h.dtype, device=h.device)
        out = out.index_add(0, self._dst.to(h.device), h[self._src.to(h.device)])
        self.ndata[reduce_func.out] = out


def batch(graph_list):
    """Disjoint union with node-id offsets, like dgl.batch."""
    srcs, dsts, nums, off = [], [], [], 0
    for g in graph_list:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        nums.append(g._n)
        off += g._n
    src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.int64)
    dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.int64)
    return DGLGraph(src, dst, off, batch_num_nodes=nums)

"""Host-side h-hop local-subgraph extraction (the producer next to the hot path).

Follows Subgraphs.generate_subgraph / generate_subgraph_link_pred
(subgraph_data_processing.py:295-346): closure over IN-edges to h hops (link
prediction: 2 hops from the first endpoint, 1 from the second -- a reference quirk --, union), uniform subsample to
`sample_nodes` re-adding the centre(s) when larger, node-induced subgraph,
memoised per item.  Integer work on numpy CSR; node order inside a subgraph is
ascending parent id (the reference's is python-`set` order, :303 -- logits are
permutation-equivariant, only the fp32 summation order inside a row can differ).
"""
import numpy as np

from .packed import SubgraphCSR


class ParentGraph(object):
    """A dataset graph held as int32 CSR by destination (in-neighbour lists)."""

    def __init__(self, indptr, indices, n):
        self.n = int(n)
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self._local = np.full(self.n, -1, dtype=np.int32)  # scratch for induced-subgraph renumbering

    @staticmethod
    def from_edges(src, dst, n):
        """Directed multigraph from COO src[e] -> dst[e] (edges kept as given)."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        order = np.argsort(dst, kind="stable")
        indptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(dst, minlength=n), out=indptr[1:])
        return ParentGraph(indptr, src[order].astype(np.int32), n)

    def number_of_nodes(self):
        return self.n

    def number_of_edges(self):
        return int(self.indices.shape[0])

    def in_neighbours(self, v):
        """`G.in_edges(v)[0]` (subgraph_data_processing.py:301)."""
        return self.indices[self.indptr[v]:self.indptr[v + 1]]

    def _rows_concat(self, rows):
        """Concatenated in-neighbour lists of `rows` and the per-row counts."""
        lo = self.indptr[rows]
        cnt = self.indptr[rows + 1] - lo
        tot = int(cnt.sum())
        if tot == 0:
            return np.zeros(0, dtype=np.int32), cnt
        starts = np.repeat(lo - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt)
        return self.indices[starts + np.arange(tot, dtype=np.int64)], cnt

    def khop_in_closure(self, seeds, h):
        """Union of the <=h-hop in-neighbourhoods of `seeds`, seeds included (ascending ids)."""
        seen = np.unique(np.asarray(seeds, dtype=np.int64))
        frontier = seen
        for _ in range(h):
            nb, _ = self._rows_concat(frontier)
            nb = np.unique(nb.astype(np.int64))
            frontier = np.setdiff1d(nb, seen, assume_unique=True)
            if frontier.size == 0:
                break
            seen = np.union1d(seen, frontier)
        return seen

    def induced(self, nodes):
        """Node-induced subgraph on `nodes` (`G.subgraph(nodes)`, :316): local id k <-> nodes[k];
        every parent edge with both endpoints selected is kept (with multiplicity)."""
        nodes = np.asarray(nodes, dtype=np.int64)
        k = nodes.shape[0]
        self._local[nodes] = np.arange(k, dtype=np.int32)
        nb, cnt = self._rows_concat(nodes)
        loc = self._local[nb]
        keep = loc >= 0
        dst = np.repeat(np.arange(k, dtype=np.int32), cnt)[keep]
        indptr = np.zeros(k + 1, dtype=np.int32)
        np.cumsum(np.bincount(dst, minlength=k), out=indptr[1:])
        self._local[nodes] = -1
        return indptr, loc[keep].astype(np.int32)


def extract_subgraph(G, i, h, sample_nodes, rng=np.random):
    """generate_subgraph (subgraph_data_processing.py:295-321) without the memo."""
    if h not in (1, 2, 3):
        raise NameError("h_hops_neighbor")  # the reference leaves it unbound for other h (:300-311)
    nodes = G.khop_in_closure([i], h)
    if nodes.shape[0] > sample_nodes:                                   # :312-314
        nodes = rng.choice(nodes, sample_nodes, replace=False)
        nodes = np.unique(np.append(nodes, [i]))
    indptr, indices = G.induced(nodes)
    centre = int(np.searchsorted(nodes, i))
    return SubgraphCSR(indptr, indices, nodes, centre)


def extract_subgraph_link_pred(G, i, j, sample_nodes, rng=np.random):
    """generate_subgraph_link_pred (:323-346).  Reproduces a reference quirk: the closure of the
    first endpoint is 2 hops (:327-329), but for the second endpoint the inner comprehension at
    :332 reads `G.in_edges(j)` for every first-hop neighbour, so only j's ONE-hop in-neighbourhood
    enters the union."""
    nodes = np.union1d(G.khop_in_closure([i], 2), G.khop_in_closure([j], 1))
    if nodes.shape[0] > sample_nodes:                                   # :337-339
        nodes = rng.choice(nodes, sample_nodes, replace=False)
        nodes = np.unique(np.append(nodes, [i, j]))
    indptr, indices = G.induced(nodes)
    centre = [int(np.searchsorted(nodes, i)), int(np.searchsorted(nodes, j))]
    return SubgraphCSR(indptr, indices, nodes, centre)

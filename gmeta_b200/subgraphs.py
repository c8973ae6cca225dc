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


class DeviceExtractor(object):
    """Device-side extraction of a whole batch of local subgraphs (gmeta_khop_select / gmeta_khop_build,
    csrc/khop.cu): the parent graphs live in HBM as ONE int32 CSR over their concatenated node ranges, a
    request is a centre (or a centre pair) in one of the graphs, and the result is the packed CSR of all
    requested subgraphs with batch offsets applied -- the layout the layer kernels consume -- without the
    subgraphs ever existing on the host.  Mirrors Subgraphs.generate_subgraph[_link_pred]
    (subgraph_data_processing.py:295-346); there is no CPU fallback."""

    def __init__(self, graphs, device=None):
        import torch
        from . import _lib
        self._lib = _lib
        self.L = _lib.lib()
        self.dev = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        ns = np.array([g.n for g in graphs], dtype=np.int64)
        self.node_off = np.concatenate([[0], np.cumsum(ns)])
        edge_off = np.concatenate([[0], np.cumsum([g.indices.shape[0] for g in graphs])])
        if self.node_off[-1] >= 2 ** 31 or edge_off[-1] >= 2 ** 31:
            raise _lib.GMetaError("parent graphs exceed the int32 index range")
        indptr = np.zeros(int(self.node_off[-1]) + 1, dtype=np.int32)
        indices = np.zeros(int(edge_off[-1]), dtype=np.int32)
        for k, g in enumerate(graphs):
            a, b = int(self.node_off[k]), int(self.node_off[k + 1])
            indptr[a + 1:b + 1] = g.indptr[1:] + edge_off[k]
            indices[int(edge_off[k]):int(edge_off[k + 1])] = g.indices.astype(np.int64) + a
        self.max_graph_nodes = int(ns.max())
        self.indptr = torch.from_numpy(indptr).to(self.dev)
        self.indices = torch.from_numpy(indices).to(self.dev)
        self._ws = None
        self._parent = None

    def _workspace(self, nb, slot):
        import torch
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(slot)
        if ws is None or ws.numel() < nb + 256:
            ws = self._ws[slot] = torch.empty(int(nb) + 256, dtype=torch.uint8, device=self.dev)
        return (ws.data_ptr() + 255) // 256 * 256

    def select(self, graph_idx, centre_a, centre_b=None, h=2, sample_nodes=1000, seed=222, slot=0):
        """First half of an extraction (gmeta_khop_select): closures, sampling and the exclusive node / edge
        sums of the requests, all left on the device.  `slot` picks the workspace, so that the selections of
        several sets can be in flight before their sizes are read with ONE copy (`totals`)."""
        import torch
        L = self.L
        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(self.dev, non_blocking=True)  # noqa: E731
        gi = np.asarray(graph_idx, dtype=np.int64)
        R = int(gi.shape[0])
        lo = self.node_off[gi]
        hi = self.node_off[gi + 1]
        hops_a, hops_b = (2, 1) if centre_b is not None else (int(h), 0)
        if centre_b is None and h not in (1, 2, 3):
            raise NameError("h_hops_neighbor")      # like the reference for other h (:300-311)
        # the request arrays of a set travel as ONE host->device copy
        req = np.stack([np.asarray(centre_a, dtype=np.int64) + lo,
                        (np.asarray(centre_b, dtype=np.int64) + lo) if centre_b is not None else lo, lo, hi]).astype(np.int32)
        d_req = i32(req)
        sel = {"R": R, "a": d_req[0], "b": d_req[1] if centre_b is not None else None, "lo": d_req[2], "hi": d_req[3],
               "req": d_req, "sample_nodes": sample_nodes, "slot": slot,
               "nb": L.gmeta_khop_workspace_bytes(R, sample_nodes, self.max_graph_nodes),
               "node_ptr": torch.empty(R + 1, dtype=torch.int32, device=self.dev),
               "edge_ptr": torch.empty(R + 1, dtype=torch.int32, device=self.dev),
               "closure_size": torch.empty(max(R, 1), dtype=torch.int32, device=self.dev)}
        sel["ws_ptr"] = self._workspace(sel["nb"], slot)
        st = torch.cuda.current_stream().cuda_stream
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        self._lib.check(L.gmeta_khop_select(ptr(self.indptr), ptr(self.indices), ptr(sel["a"]), ptr(sel["b"]), ptr(sel["lo"]),
                                            ptr(sel["hi"]), R, hops_a, hops_b, sample_nodes, self.max_graph_nodes, seed,
                                            ptr(sel["node_ptr"]), ptr(sel["edge_ptr"]), ptr(sel["closure_size"]),
                                            sel["ws_ptr"], sel["nb"], st), "khop_select")
        return sel

    def totals(self, sels):
        """Packed node / edge totals of the given selections: the one device->host copy of an extraction."""
        import torch
        t = torch.stack([x for s in sels for x in (s["node_ptr"][s["R"]], s["edge_ptr"][s["R"]])]).cpu()
        for i, s in enumerate(sels):
            s["N"], s["E"] = int(t[2 * i]), int(t[2 * i + 1])

    def build_into(self, sel, indptr, indices, feat_row, centre_row, parent=None):
        """Second half (gmeta_khop_build) straight into caller-provided int32 device tensors (views of the packed
        buffer): indptr [N+1], indices [>= E], feat_row [N], centre_row [R * cps]."""
        import torch
        N = sel["N"]
        if parent is None:
            if self._parent is None or self._parent.numel() < max(N, 1):
                self._parent = torch.empty(max(N, 1), dtype=torch.int32, device=self.dev)
            parent = self._parent
        st = torch.cuda.current_stream().cuda_stream
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        self._lib.check(self.L.gmeta_khop_build(ptr(self.indptr), ptr(self.indices), ptr(sel["a"]), ptr(sel["b"]), ptr(sel["lo"]),
                                                sel["R"], sel["sample_nodes"], self.max_graph_nodes, ptr(sel["node_ptr"]),
                                                ptr(sel["edge_ptr"]), ptr(indptr), ptr(indices), ptr(parent), ptr(feat_row),
                                                ptr(centre_row), sel["ws_ptr"], sel["nb"], st), "khop_build")

    def extract(self, graph_idx, centre_a, centre_b=None, h=2, sample_nodes=1000, seed=222):
        """graph_idx / centre_a / centre_b: int arrays of length R (node ids inside their graph).
        Node classification: `h` hops from centre_a.  Link prediction (centre_b given): 2 hops from a, 1 from b
        (the reference ignores h there, :327-333).  Returns a dict of device tensors (packed layout) and
        the per-request pointers."""
        import torch
        sel = self.select(graph_idx, centre_a, centre_b, h, sample_nodes, seed)
        self.totals([sel])                                            # the batch's only D2H: 8 bytes
        N, E, R = sel["N"], sel["E"], sel["R"]
        out = {"node_ptr": sel["node_ptr"], "edge_ptr": sel["edge_ptr"], "closure_size": sel["closure_size"], "N": N, "E": E,
               "indptr": torch.empty(N + 1, dtype=torch.int32, device=self.dev),
               "indices": torch.empty(max(E, 1), dtype=torch.int32, device=self.dev),
               "parent": torch.empty(max(N, 1), dtype=torch.int32, device=self.dev),
               "feat_row": torch.empty(max(N, 1), dtype=torch.int32, device=self.dev),
               "centre_row": torch.empty(max(R, 1) * (2 if sel["b"] is not None else 1), dtype=torch.int32, device=self.dev)}
        self.build_into(sel, out["indptr"], out["indices"], out["feat_row"], out["centre_row"], out["parent"])
        return out

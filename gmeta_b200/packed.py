"""Host-side containers for local subgraphs and their packed (batched) form.

`PackedSubgraphBatch` is what this build passes where the reference passes a DGL
batched graph (subgraph_data_processing.py:399-406 `dgl.batch(support_x)`); it
duck-types the two members the hot path reads from it -- `.batch_num_nodes`
(python list, learner.py:161-162) and `.to(device)` (meta.py:122) -- and carries
the adjacency as int32 CSR *by destination* (row v = in-neighbours of v, the
direction `update_all(copy_src, sum)` aggregates over, learner.py:38-45) plus the
by-source view the backward needs (dH_u gathers over out-edges).
"""
import numpy as np


def csr_transpose(indptr, indices, n):
    """CSR by destination -> CSR by source (stable: destinations ascending per row)."""
    indptr = np.asarray(indptr)
    indices = np.asarray(indices)
    dst = np.repeat(np.arange(n, dtype=np.int32), np.diff(indptr))
    order = np.argsort(indices, kind="stable")
    t_indices = dst[order].astype(np.int32)
    t_indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(indices, minlength=n), out=t_indptr[1:])
    return t_indptr, t_indices


class SubgraphCSR(object):
    """One node-induced local subgraph (subgraph_data_processing.py:316-321 result)."""

    __slots__ = ("n", "indptr", "indices", "t_indptr", "t_indices", "parent_nid", "centre")

    def __init__(self, indptr, indices, parent_nid, centre, t_indptr=None, t_indices=None):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.n = self.indptr.shape[0] - 1
        if t_indptr is None:
            t_indptr, t_indices = csr_transpose(self.indptr, self.indices, self.n)
        self.t_indptr = np.ascontiguousarray(t_indptr, dtype=np.int32)
        self.t_indices = np.ascontiguousarray(t_indices, dtype=np.int32)
        self.parent_nid = np.ascontiguousarray(parent_nid, dtype=np.int64)
        self.centre = centre  # int, or [i_local, j_local] in link-prediction mode

    @staticmethod
    def from_edges(src, dst, n, parent_nid=None, centre=0):
        """Build from a COO edge list src[e] -> dst[e] (multi-edges kept)."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        order = np.argsort(dst, kind="stable")
        indptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(dst, minlength=n), out=indptr[1:])
        if parent_nid is None:
            parent_nid = np.arange(n)
        return SubgraphCSR(indptr, src[order].astype(np.int32), parent_nid, centre)


class PackedSubgraphBatch(object):
    """Disjoint union of local subgraphs with node-id offsets applied (one task's set)."""

    def __init__(self, indptr, indices, t_indptr, t_indices, batch_num_nodes):
        self.indptr = indptr
        self.indices = indices
        self.t_indptr = t_indptr
        self.t_indices = t_indices
        self.batch_num_nodes = [int(x) for x in batch_num_nodes]
        self.n_nodes = int(indptr.shape[0] - 1)
        self.n_edges = int(indices.shape[0])
        self.parent_ids = None
        self.parent_id_lists = None

    @staticmethod
    def batch(subgraphs, id_lists=None):
        """`id_lists`: the per-subgraph parent-id sequences the caller will hand to Meta.forward as n_spt / n_qry
        next to this batch (default: the subgraphs' own `parent_nid` arrays).  The batch remembers those very
        objects: packing uses the pre-concatenated `parent_ids` only when it is later given the same objects."""
        ns = np.array([s.n for s in subgraphs], dtype=np.int64)
        es = np.array([s.indices.shape[0] for s in subgraphs], dtype=np.int64)
        n_off = np.concatenate([[0], np.cumsum(ns)])
        e_off = np.concatenate([[0], np.cumsum(es)])
        N, E = int(n_off[-1]), int(e_off[-1])
        indptr = np.empty(N + 1, dtype=np.int32)
        t_indptr = np.empty(N + 1, dtype=np.int32)
        indices = np.empty(E, dtype=np.int32)
        t_indices = np.empty(E, dtype=np.int32)
        indptr[0] = 0
        t_indptr[0] = 0
        for k, s in enumerate(subgraphs):
            a, b = int(n_off[k]), int(n_off[k + 1])
            ea, eb = int(e_off[k]), int(e_off[k + 1])
            np.add(s.indptr[1:], ea, out=indptr[a + 1:b + 1])
            np.add(s.t_indptr[1:], ea, out=t_indptr[a + 1:b + 1])
            np.add(s.indices, a, out=indices[ea:eb])
            np.add(s.t_indices, a, out=t_indices[ea:eb])
        pb = PackedSubgraphBatch(indptr, indices, t_indptr, t_indices, ns.tolist())
        # parent ids of all nodes in batch order (the n_spt / n_qry lists of an episode, concatenated once here)
        pb.parent_ids = np.concatenate([s.parent_nid for s in subgraphs]) if len(subgraphs) else np.zeros(0, np.int64)
        pb.parent_id_lists = id_lists if id_lists is not None else [s.parent_nid for s in subgraphs]
        return pb

    # duck-typing of the DGL batched graph members the hot path touches
    def to(self, device):
        return self

    def number_of_nodes(self):
        return self.n_nodes

    def number_of_edges(self):
        return self.n_edges

    def in_degrees(self):
        return np.diff(self.indptr)

    def edges(self):
        """COO (src, dst) in CSR order (for tests / adapters)."""
        dst = np.repeat(np.arange(self.n_nodes, dtype=np.int64), np.diff(self.indptr))
        return self.indices.astype(np.int64), dst

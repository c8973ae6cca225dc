"""A meta-batch assembled entirely in HBM: h-hop extraction on the device (csrc/khop.cu through
subgraphs.DeviceExtractor) straight into the packed-set layout the ProtoMAML driver consumes, so that a
training step needs only the centre node ids and the labels from the host -- the subgraphs themselves
never exist there (north_star; replaces Subgraphs.generate_subgraph + dgl.batch + the per-task feature
gather, subgraph_data_processing.py:295-346,399-406 and meta.py:119-122).

The extractor emits the packed CSR by destination, `feat_row` and `centre_row`.  What the driver needs
on top is structure-only integer work on those device arrays, done here with a handful of torch device ops
(sort / cumsum / unique / searchsorted: plumbing, not arithmetic):
  * the CSR by source (stable sort of the edge list by source: destinations ascending per row -- the same
    rule as packed.csr_transpose on the host),
  * the active-row lists per layer (centres, then the in-neighbours of the layer above) with their task
    pointers,
and three tiny host round trips (task row pointers and active-row counts: a few hundred bytes) for the tile
tables.  The result is written into ONE int32 device buffer with the segment names of packing._SEGS, so
Meta._enqueue runs unchanged.  For subgraphs below the `sample_nodes` cap the buffer is identical, segment
by segment, to what the host path (packing.pack_meta_batch of host-extracted subgraphs) uploads
(tests/test_gpu_device_batch.py); above the cap the extractor samples with its own counter-based hash, not
numpy's generator.
"""
import numpy as np
import torch

from . import packing
from .learner import tile_table


class CentreRequests(object):
    """The subgraph requests of one set (support or query) of a meta-batch: for task t the subgraphs
    [sub_off[t], sub_off[t+1]) of the flat lists; `centre_b` is given for link prediction only."""

    def __init__(self, graph_idx, centre_a, centre_b, sub_off, labels):
        self.graph_idx = np.asarray(graph_idx, dtype=np.int64)
        self.centre_a = np.asarray(centre_a, dtype=np.int64)
        self.centre_b = None if centre_b is None else np.asarray(centre_b, dtype=np.int64)
        self.sub_off = np.asarray(sub_off, dtype=np.int64)
        self.labels = np.asarray(labels, dtype=np.int64)

    @staticmethod
    def from_tasks(tasks):
        """tasks: list (per task) of (graph_idx [S_t], centre_a [S_t], centre_b [S_t] or None, labels [S_t])."""
        gi = np.concatenate([np.asarray(t[0], dtype=np.int64) for t in tasks])
        ca = np.concatenate([np.asarray(t[1], dtype=np.int64) for t in tasks])
        cb = None if tasks[0][2] is None else np.concatenate([np.asarray(t[2], dtype=np.int64) for t in tasks])
        y = np.concatenate([np.asarray(t[3], dtype=np.int64) for t in tasks])
        off = np.concatenate([[0], np.cumsum([len(t[0]) for t in tasks])])
        return CentreRequests(gi, ca, cb, off, y)

    @staticmethod
    def from_host_batch(graphs, centres, node_ids, graph_idx, labels):
        """The same requests a host-extracted set was built from (x_*, c_*, n_*, g_*, y_* of Meta.forward)."""
        tasks = []
        for g, c, n, gi, y in zip(graphs, centres, node_ids, graph_idx, labels):
            c = c.numpy() if hasattr(c, "numpy") else np.asarray(c)
            y = y.numpy() if hasattr(y, "numpy") else np.asarray(y)
            if c.ndim == 2:
                ca = [int(n[k][int(c[k, 0])]) for k in range(len(n))]
                cb = [int(n[k][int(c[k, 1])]) for k in range(len(n))]
            else:
                ca, cb = [int(n[k][int(c[k])]) for k in range(len(n))], None
            tasks.append((list(gi), ca, cb, y))
        return CentreRequests.from_tasks(tasks)


def _rows_concat_dev(indptr, indices, rows):
    """Concatenated in-neighbour lists of `rows` (device; one size round trip)."""
    lo = indptr[rows].long()
    cnt = indptr[rows + 1].long() - lo
    tot = int(cnt.sum())
    if tot == 0:
        return indices[:0].long()
    start = torch.repeat_interleave(lo - (torch.cumsum(cnt, 0) - cnt), cnt, output_size=tot)
    return indices[start + torch.arange(tot, device=indices.device)].long()


def pack_set_on_device(extractor, req, h, sample_nodes, n_layers, seed):
    """Extract one set and derive everything `gmeta_packed_set_t` needs.  Returns (PackedSetHost without
    offsets, dict name -> device tensor)."""
    dev = extractor.dev
    out = extractor.extract(req.graph_idx, req.centre_a, req.centre_b, h=h, sample_nodes=sample_nodes, seed=seed)
    N, E = out["N"], out["E"]
    ps = packing.PackedSetHost()
    ps.T, ps.S = int(req.sub_off.shape[0] - 1), int(req.graph_idx.shape[0])
    ps.N, ps.E = N, E
    ps.cps = 2 if req.centre_b is not None else 1
    ps.sub_off = req.sub_off
    ps.max_rows_per_task = int(np.diff(req.sub_off).max()) if ps.T else 0
    sub_off_dev = torch.as_tensor(req.sub_off, device=dev)
    node_off_dev = out["node_ptr"].long()[sub_off_dev]
    ps.node_off = node_off_dev.cpu().numpy().astype(np.int64)                 # round trip 1: T+1 ints
    ps.tiles = tile_table(ps.node_off)
    ps.n_tiles = int(ps.tiles[0].shape[0])
    indptr, indices = out["indptr"], out["indices"][:E]
    # CSR by source: stable sort of the edges by source keeps destinations ascending inside a row
    deg = (indptr[1:] - indptr[:-1]).long()
    dst = torch.repeat_interleave(torch.arange(N, device=dev, dtype=torch.int32), deg, output_size=E)
    order = torch.argsort(indices, stable=True)
    t_indices = dst[order]
    t_indptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    t_indptr[1:] = torch.cumsum(torch.bincount(indices.long(), minlength=N), 0).to(torch.int32)
    centre = out["centre_row"][:ps.S * ps.cps]
    seg = {"indptr": indptr, "indices": indices, "t_indptr": t_indptr, "t_indices": t_indices,
           "centre_row": centre, "feat_row": out["feat_row"][:N]}
    # small host-derived segments: shipped with ONE copy per meta-batch (build())
    hseg = {"tile_row0": ps.tiles[0], "tile_nrows": ps.tiles[1], "tile_task": ps.tiles[2], "task_row_ptr": ps.node_off,
            "task_sub_ptr": req.sub_off, "labels": req.labels}
    # active rows per layer: centres, then the in-neighbours of the layer above (sorted global row ids are grouped
    # by task because a task is a contiguous row range)
    ps.n_layers = n_layers
    ps.act = [{} for _ in range(n_layers)]
    rows = torch.unique(centre.long())
    per_layer = [None] * n_layers
    if n_layers:
        per_layer[n_layers - 1] = rows
        for l in range(n_layers - 1, 0, -1):
            rows = torch.unique(_rows_concat_dev(indptr, indices, rows))     # round trip 2 (per extra layer)
            per_layer[l - 1] = rows
        tptr = torch.stack([torch.searchsorted(r, node_off_dev) for r in per_layer]).cpu().numpy()   # round trip 3
        for l in range(n_layers):
            tiles = tile_table(tptr[l].astype(np.int64))
            ps.act[l] = {"n": int(per_layer[l].shape[0]), "n_tiles": int(tiles[0].shape[0])}
            seg["act_rows%d" % l] = per_layer[l].to(torch.int32)
            hseg["act_task_ptr%d" % l] = tptr[l]
            for k, arr in zip(("act_tile_row0", "act_tile_nrows", "act_tile_task"), tiles):
                hseg["%s%d" % (k, l)] = arr
        seg["centre_pos"] = torch.searchsorted(per_layer[n_layers - 1], centre.long()).to(torch.int32)
    else:
        seg["centre_pos"] = torch.zeros_like(centre)
    ps.sizes = {k: int(v.shape[0]) for k, v in list(hseg.items()) + list(seg.items())}
    return ps, seg, hseg


def build(extractor, req_spt, req_qry, h, sample_nodes, n_layers, seed=222):
    """Both sets of a meta-batch -> (ps_spt, ps_qry, one int32 device buffer holding every segment)."""
    sets = [pack_set_on_device(extractor, r, h, sample_nodes, n_layers, seed + 7919 * i)
            for i, r in enumerate((req_spt, req_qry))]
    # layout: [host-derived small segments of both sets | device-derived segments of both sets]
    off = 0
    for ps, seg, hseg in sets:
        for k in hseg:
            ps.off[k] = off
            off += packing._al(ps.sizes[k])
    n_host = off
    for ps, seg, hseg in sets:
        for k in seg:
            ps.off[k] = off
            off += packing._al(ps.sizes[k])
    host = np.zeros(max(n_host, 4), dtype=np.int32)
    for ps, seg, hseg in sets:
        for k, v in hseg.items():
            host[ps.off[k]:ps.off[k] + ps.sizes[k]] = v
    ints = torch.empty(max(off, 4), dtype=torch.int32, device=extractor.dev)
    ints[:host.shape[0]].copy_(torch.from_numpy(host))                      # the batch's one host->device copy of structure
    for ps, seg, hseg in sets:
        for k, v in seg.items():
            if ps.sizes[k]:
                ints[ps.off[k]:ps.off[k] + ps.sizes[k]] = v
        ps.end = off
    return sets[0][0], sets[1][0], ints

"""A meta-batch assembled entirely in HBM: h-hop extraction on the device (csrc/khop.cu through
subgraphs.DeviceExtractor) straight into the packed-set layout the ProtoMAML driver consumes, so that a
training step needs only the centre node ids and the labels from the host -- the subgraphs themselves
never exist there (north_star; replaces Subgraphs.generate_subgraph + dgl.batch + the per-task feature
gather, subgraph_data_processing.py:295-346,399-406 and meta.py:119-122).

The extractor (gmeta_khop_select / gmeta_khop_build) writes the packed CSR by destination, `feat_row` and
`centre_row` straight into their segments of ONE int32 device buffer; gmeta_packed_set_finish
(csrc/batch_assemble.cu) derives the rest there with a handful of CUDA passes:
  * the CSR by source (destinations ascending per row -- the rule of packed.csr_transpose on the host),
  * task row pointers and the row-tile table,
  * the active-row lists per layer (centres, then the in-neighbours of the layer above) with their task
    pointers, tile tables and the centre positions.
Segments whose length depends on the data (tiles, active rows) are laid out at their upper bounds, so nothing
has to be known on the host before the passes are enqueued; the host reads two small arrays per meta-batch: the
packed node / edge totals after the selection (to size the buffer) and the realised counts (tiles, active
rows) at the end.  The buffer has the segment names of packing._SEGS, so Meta._enqueue runs unchanged.  For
subgraphs below the `sample_nodes` cap every segment is identical to what the host path
(packing.pack_meta_batch of host-extracted subgraphs) uploads (tests/test_gpu_device_batch.py); above the cap
the extractor samples with its own counter-based hash, not numpy's generator.
"""
import numpy as np
import torch

import ctypes as C

from . import _lib, packing
from ._lib import TILE_ROWS


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


def _layout_set(ps, n_layers, off):
    """Segment offsets of one set at upper-bound sizes (ps.cap), 16-byte aligned; returns the next free offset."""
    cap_t = (ps.N + TILE_ROWS - 1) // TILE_ROWS + ps.T
    nc = ps.S * ps.cps
    ps.cap = {"indptr": ps.N + 1, "indices": ps.E, "t_indptr": ps.N + 1, "t_indices": ps.E, "tile_row0": cap_t,
              "tile_nrows": cap_t, "tile_task": cap_t, "task_row_ptr": ps.T + 1, "centre_row": nc, "feat_row": ps.N,
              "centre_pos": nc}
    for l in range(n_layers):
        ps.cap["act_rows%d" % l] = ps.N
        ps.cap["act_task_ptr%d" % l] = ps.T + 1
        for k in ("act_tile_row0", "act_tile_nrows", "act_tile_task"):
            ps.cap["%s%d" % (k, l)] = cap_t
    for k, n in ps.cap.items():
        ps.off[k] = off
        off += packing._al(n)
    return off


def build(extractor, req_spt, req_qry, h, sample_nodes, n_layers, seed=222, timings=None):
    """Both sets of a meta-batch -> (ps_spt, ps_qry, one int32 device buffer holding every segment).
    `timings` (debug): a dict that receives the wall-clock milliseconds of every phase, each closed by a device
    synchronisation (tools/devbatch_timeline.py)."""
    import time
    t_mark = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            torch.cuda.synchronize()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + 1e3 * (now - t_mark[0])
            t_mark[0] = now
    L = extractor.L
    dev = extractor.dev
    reqs = (req_spt, req_qry)
    sels = [extractor.select(r.graph_idx, r.centre_a, r.centre_b, h=h, sample_nodes=sample_nodes, seed=seed + 7919 * i,
                             slot=i) for i, r in enumerate(reqs)]
    lap("select")
    extractor.totals(sels)                                                  # device -> host: N, E of both sets
    lap("totals")
    sets = []
    # layout: [task_sub_ptr, labels of both sets (host-known) | counts of both sets | everything derived on the device]
    off = 0
    for r, sel in zip(reqs, sels):
        ps = packing.PackedSetHost()
        ps.T, ps.S = int(r.sub_off.shape[0] - 1), int(r.graph_idx.shape[0])
        ps.N, ps.E = sel["N"], sel["E"]
        ps.cps = 2 if r.centre_b is not None else 1
        ps.sub_off = r.sub_off
        ps.n_layers = n_layers
        for k, n in (("task_sub_ptr", ps.T + 1), ("labels", ps.S)):
            ps.off[k] = off
            off += packing._al(n)
        sets.append(ps)
    n_host = off
    for ps in sets:                                  # realised counts of both sets: adjacent, read back with one copy
        ps.off["counts"] = off
        off += packing._al(2 + 2 * n_layers)
    n_counts_end = off
    for ps in sets:
        off = _layout_set(ps, n_layers, off)
    host = np.zeros(max(n_host, 4), dtype=np.int32)
    for ps, r in zip(sets, reqs):
        host[ps.off["task_sub_ptr"]:ps.off["task_sub_ptr"] + ps.T + 1] = r.sub_off
        host[ps.off["labels"]:ps.off["labels"] + ps.S] = r.labels
    ints = torch.empty(max(off, 4), dtype=torch.int32, device=dev)
    ints[:host.shape[0]].copy_(torch.from_numpy(host), non_blocking=True)  # the batch's one host->device copy of structure
    base = ints.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    view = lambda ps, k: ints[ps.off[k]:ps.off[k] + max(ps.cap[k], 1)]      # noqa: E731
    lap("layout")
    for ps, sel in zip(sets, sels):
        extractor.build_into(sel, view(ps, "indptr"), view(ps, "indices"), view(ps, "feat_row"), view(ps, "centre_row"))
        lap("khop_build")
        nb = L.gmeta_packed_set_finish_workspace_bytes(ps.N, ps.E, n_layers)
        if getattr(extractor, "_finish_ws", None) is None or extractor._finish_ws.numel() < nb + 256:
            extractor._finish_ws = torch.empty(int(nb) + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (extractor._finish_ws.data_ptr() + 255) // 256 * 256
        seg = lambda k: base + 4 * ps.off[k]                                 # noqa: E731
        arr = lambda k: (C.c_void_p * max(n_layers, 1))(*[seg("%s%d" % (k, l)) for l in range(n_layers)])  # noqa: E731
        _lib.check(L.gmeta_packed_set_finish(seg("indptr"), seg("indices"), ps.N, ps.E, sel["node_ptr"].data_ptr(),
                                             seg("task_sub_ptr"), ps.T, seg("centre_row"), ps.S * ps.cps, n_layers,
                                             seg("t_indptr"), seg("t_indices"), seg("task_row_ptr"), seg("tile_row0"),
                                             seg("tile_nrows"), seg("tile_task"), arr("act_rows"), arr("act_task_ptr"),
                                             arr("act_tile_row0"), arr("act_tile_nrows"), arr("act_tile_task"),
                                             seg("centre_pos"), seg("counts"), ws_ptr, nb, st), "packed_set_finish")
        if n_layers == 0:
            view(ps, "centre_pos").zero_()
        lap("finish")
    # device -> host: the realised counts of both sets (adjacent segments: one copy)
    span = ints[n_host:n_counts_end].cpu().numpy()
    for ps in sets:
        cnt = span[ps.off["counts"] - n_host:][:2 + 2 * n_layers]
        ps.n_tiles = int(cnt[0])                     # cnt[1]: node rows of the largest task (not needed by the driver)
        ps.max_rows_per_task = int(np.diff(ps.sub_off).max()) if ps.T else 0     # readout rows (subgraphs) per task
        ps.act = [{"n": int(cnt[2 + l]), "n_tiles": int(cnt[2 + n_layers + l])} for l in range(n_layers)]
        ps.sizes = dict(ps.cap)
        ps.sizes.update({"task_sub_ptr": ps.T + 1, "labels": ps.S})
        for k in ("tile_row0", "tile_nrows", "tile_task"):
            ps.sizes[k] = ps.n_tiles
        for l in range(n_layers):
            ps.sizes["act_rows%d" % l] = ps.act[l]["n"]
            for k in ("act_tile_row0", "act_tile_nrows", "act_tile_task"):
                ps.sizes["%s%d" % (k, l)] = ps.act[l]["n_tiles"]
        ps.end = off
    lap("counts")
    return sets[0], sets[1], ints

"""Host-side packing of a meta-batch into the "packed set" layout of include/gmeta_b200.h.

Input is exactly what the reference's Meta.forward receives (meta.py:101, lists of length
task_num produced by subgraph_data_processing.py:348-419); output is ONE pinned int32 staging
buffer per meta-batch (so the host->device transfer is a single async copy) plus the counts the
C ABI needs.  All integer work, vectorised numpy; nothing here touches feature values.
"""
import numpy as np
import torch

from .learner import tile_table

_SEGS = ("indptr", "indices", "t_indptr", "t_indices", "tile_row0", "tile_nrows", "tile_task",
         "task_row_ptr", "task_sub_ptr", "centre_row", "feat_row", "labels", "centre_pos")


def _al(n):
    return (n + 3) // 4 * 4      # 16-byte aligned int32 segments


class PackedSetHost(object):
    """Segment offsets (in int32 elements) of one set inside the staging buffer + its counts."""

    def __init__(self):
        self.off = {}
        self.N = self.E = self.S = self.T = self.n_tiles = self.cps = 0
        self.max_rows_per_task = 0
        self.max_label = 0


def _flat_ids(node_ids):
    """n_spt[i]: list (per subgraph) of parent-id lists/arrays -> one int64 array."""
    if len(node_ids) == 0:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([np.asarray(x, dtype=np.int64) for x in node_ids])


def _rows_concat(indptr, indices, rows):
    """Concatenated in-neighbour lists of `rows` of a CSR-by-destination graph."""
    lo = indptr[rows].astype(np.int64)
    cnt = indptr[rows + 1].astype(np.int64) - lo
    tot = int(cnt.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64)
    starts = np.repeat(lo - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt)
    return indices[starts + np.arange(tot, dtype=np.int64)].astype(np.int64)


def active_rows(g, centre, n_layers):
    """Rows of one task's batched graph where dL/dZ_l can be non-zero, per GCN layer l: only the
    centre rows are read out (learner.py:166-170), so act[L-1] = centres and act[l-1] = the
    in-neighbours of act[l].  Local (task-relative) sorted row ids."""
    c = centre.numpy() if hasattr(centre, "numpy") else np.asarray(centre)
    sub_first = np.concatenate([[0], np.cumsum(g.batch_num_nodes)])[:-1]
    rows = np.unique((c.astype(np.int64) + (sub_first[:, None] if c.ndim == 2 else sub_first)).reshape(-1))
    out = [None] * n_layers
    out[n_layers - 1] = rows
    for l in range(n_layers - 1, 0, -1):
        rows = np.unique(_rows_concat(g.indptr, g.indices, rows))
        out[l - 1] = rows
    return out


def plan_set(graphs, centres, off0, n_layers=0):
    """Sizes and segment offsets for one set (spt or qry) of all tasks."""
    ps = PackedSetHost()
    ps.T = len(graphs)
    ns = np.array([g.n_nodes for g in graphs], dtype=np.int64)
    es = np.array([g.n_edges for g in graphs], dtype=np.int64)
    ss = np.array([len(g.batch_num_nodes) for g in graphs], dtype=np.int64)
    ps.node_off = np.concatenate([[0], np.cumsum(ns)])
    ps.edge_off = np.concatenate([[0], np.cumsum(es)])
    ps.sub_off = np.concatenate([[0], np.cumsum(ss)])
    ps.N, ps.E, ps.S = int(ps.node_off[-1]), int(ps.edge_off[-1]), int(ps.sub_off[-1])
    ps.max_rows_per_task = int(ss.max()) if ps.T else 0
    c0 = centres[0]
    ps.cps = 2 if (hasattr(c0, "dim") and c0.dim() == 2) or (isinstance(c0, np.ndarray) and c0.ndim == 2) else 1
    ps.tiles = tile_table(ps.node_off.astype(np.int64))
    ps.n_tiles = int(ps.tiles[0].shape[0])
    sizes = {"indptr": ps.N + 1, "indices": ps.E, "t_indptr": ps.N + 1, "t_indices": ps.E,
             "tile_row0": ps.n_tiles, "tile_nrows": ps.n_tiles, "tile_task": ps.n_tiles,
             "task_row_ptr": ps.T + 1, "task_sub_ptr": ps.T + 1, "centre_row": ps.S * ps.cps,
             "feat_row": ps.N, "labels": ps.S, "centre_pos": ps.S * ps.cps}
    # active rows of the backward pass (per layer), their task pointers and tile tables
    ps.n_layers = n_layers
    ps.act = []
    for l in range(n_layers):
        ps.act.append({})
    if n_layers:
        per_task = [active_rows(g, c, n_layers) for g, c in zip(graphs, centres)]
        for l in range(n_layers):
            cnt = np.array([pt[l].shape[0] for pt in per_task], dtype=np.int64)
            tptr = np.concatenate([[0], np.cumsum(cnt)])
            rows = (np.concatenate([pt[l] + ps.node_off[t] for t, pt in enumerate(per_task)])
                    if ps.T else np.zeros(0, dtype=np.int64))
            tiles = tile_table(tptr)
            ps.act[l] = {"rows": rows, "task_ptr": tptr, "tiles": tiles, "n": int(tptr[-1]),
                         "n_tiles": int(tiles[0].shape[0])}
            sizes["act_rows%d" % l] = ps.act[l]["n"]
            sizes["act_task_ptr%d" % l] = ps.T + 1
            for k in ("act_tile_row0", "act_tile_nrows", "act_tile_task"):
                sizes["%s%d" % (k, l)] = ps.act[l]["n_tiles"]
    off = off0
    for k in sizes:
        ps.off[k] = off
        off += _al(sizes[k])
    ps.sizes = sizes
    ps.end = off
    return ps


def fill_set(buf, ps, graphs, labels, centres, node_ids, graph_idx, graph_row_off):
    """Write one set's segments into the int32 numpy view `buf` of the staging buffer."""
    o, N, E = ps.off, ps.N, ps.E
    indptr = buf[o["indptr"]:o["indptr"] + N + 1]
    t_indptr = buf[o["t_indptr"]:o["t_indptr"] + N + 1]
    indices = buf[o["indices"]:o["indices"] + E]
    t_indices = buf[o["t_indices"]:o["t_indices"] + E]
    centre = buf[o["centre_row"]:o["centre_row"] + ps.S * ps.cps]
    feat_row = buf[o["feat_row"]:o["feat_row"] + N]
    lab = buf[o["labels"]:o["labels"] + ps.S]
    indptr[0] = 0
    t_indptr[0] = 0
    single_graph = graph_row_off.shape[0] == 1
    for t, g in enumerate(graphs):
        a, b = int(ps.node_off[t]), int(ps.node_off[t + 1])
        ea, eb = int(ps.edge_off[t]), int(ps.edge_off[t + 1])
        sa, sb = int(ps.sub_off[t]), int(ps.sub_off[t + 1])
        np.add(g.indptr[1:], ea, out=indptr[a + 1:b + 1])
        np.add(g.t_indptr[1:], ea, out=t_indptr[a + 1:b + 1])
        np.add(g.indices, a, out=indices[ea:eb])
        np.add(g.t_indices, a, out=t_indices[ea:eb])
        bnn = np.asarray(g.batch_num_nodes, dtype=np.int64)
        sub_first = np.concatenate([[0], np.cumsum(bnn)])[:-1] + a           # learner.py:161-163
        c = centres[t].numpy() if hasattr(centres[t], "numpy") else np.asarray(centres[t])
        if ps.cps == 2:
            centre[2 * sa:2 * sb] = (c.astype(np.int64) + sub_first[:, None]).reshape(-1)
        else:
            centre[sa:sb] = c.astype(np.int64) + sub_first
        ids = _flat_ids(node_ids[t])                                          # meta.py:119-120
        if single_graph:
            feat_row[a:b] = ids
        else:
            feat_row[a:b] = ids + np.repeat(graph_row_off[np.asarray(graph_idx[t], dtype=np.int64)], bnn)
        y = labels[t].numpy() if hasattr(labels[t], "numpy") else np.asarray(labels[t])
        lab[sa:sb] = y
    buf[o["tile_row0"]:o["tile_row0"] + ps.n_tiles] = ps.tiles[0]
    buf[o["tile_nrows"]:o["tile_nrows"] + ps.n_tiles] = ps.tiles[1]
    buf[o["tile_task"]:o["tile_task"] + ps.n_tiles] = ps.tiles[2]
    buf[o["task_row_ptr"]:o["task_row_ptr"] + ps.T + 1] = ps.node_off
    buf[o["task_sub_ptr"]:o["task_sub_ptr"] + ps.T + 1] = ps.sub_off
    if ps.n_layers:     # position of every centre row in the (globally sorted) active-row list of the last layer
        buf[o["centre_pos"]:o["centre_pos"] + ps.S * ps.cps] = np.searchsorted(ps.act[ps.n_layers - 1]["rows"], centre)
    for l in range(ps.n_layers):
        a = ps.act[l]
        buf[o["act_rows%d" % l]:o["act_rows%d" % l] + a["n"]] = a["rows"]
        buf[o["act_task_ptr%d" % l]:o["act_task_ptr%d" % l] + ps.T + 1] = a["task_ptr"]
        for k, arr in zip(("act_tile_row0", "act_tile_nrows", "act_tile_task"), a["tiles"]):
            buf[o["%s%d" % (k, l)]:o["%s%d" % (k, l)] + a["n_tiles"]] = arr


def validate_labels(y_spt, y_qry, k_spt, lib=None):
    """The equal-count requirements the reference enforces implicitly through torch.stack
    (meta.py:42,65-66): every support class has >= k_spt members, query classes are balanced
    and the same as the support classes.  With `lib` the check is one library call for all tasks."""
    if lib is not None and len(y_spt):
        import ctypes as C
        T = len(y_spt)
        keep = []
        vpa = lambda ys: (C.c_void_p * T)(*[_ptr_i64(y, keep) for y in ys])      # noqa: E731
        ns = np.array([len(y) for y in y_spt], dtype=np.int32)
        nq = np.array([len(y) for y in y_qry], dtype=np.int32)
        rc = lib.gmeta_host_validate_labels(T, vpa(y_spt), ns.ctypes.data, vpa(y_qry), nq.ctypes.data, k_spt)
        if rc == -1:
            raise RuntimeError("stack expects each tensor to be equal size: a support class has fewer "
                               "than k_spt=%d members (meta.py:42)" % k_spt)
        if rc == -2:
            raise RuntimeError("stack expects each tensor to be equal size: query classes are not "
                               "balanced (meta.py:65)")
        if rc == -3:
            raise RuntimeError("support and query sets must hold the same classes (meta.py:56-79)")
        return int(rc)
    max_classes = 1
    for ys, yq in zip(y_spt, y_qry):
        ys = ys.numpy() if hasattr(ys, "numpy") else np.asarray(ys)
        yq = yq.numpy() if hasattr(yq, "numpy") else np.asarray(yq)
        if ys.size and yq.size and ys.min() >= 0 and yq.min() >= 0 and max(ys.max(), yq.max()) < 4096:
            ns, nq = np.bincount(ys), np.bincount(yq)      # small non-negative labels: counting beats sorting
            cs, cq = np.flatnonzero(ns), np.flatnonzero(nq)
            ns, nq = ns[cs], nq[cq]
        else:
            cs, ns = np.unique(ys, return_counts=True)
            cq, nq = np.unique(yq, return_counts=True)
        if ns.min() < k_spt:
            raise RuntimeError("stack expects each tensor to be equal size: a support class has fewer "
                               "than k_spt=%d members (meta.py:42)" % k_spt)
        if nq.min() != nq.max():
            raise RuntimeError("stack expects each tensor to be equal size: query classes are not "
                               "balanced (meta.py:65)")
        if cs.shape[0] != cq.shape[0] or (cs != cq).any():
            raise RuntimeError("support and query sets must hold the same classes (meta.py:56-79)")
        max_classes = max(max_classes, int(cs.shape[0]))
    return max_classes


# ------------------------------------------------------------------------------------------------
# Fast path used by Meta.upload_batch: same layout and values as plan_set + fill_set, but the bulk
# CSR concatenation runs in the library on host threads (gmeta_host_pack_csr), the active-row lists
# are derived once from the PACKED arrays instead of per task, and the parent ids a
# PackedSubgraphBatch carries from batch time are used instead of re-concatenating the id lists.
# ------------------------------------------------------------------------------------------------
def _i32c(a):
    return a if (a.dtype == np.int32 and a.flags.c_contiguous) else np.ascontiguousarray(a, dtype=np.int32)


def _np(x):
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def _csr_ptrs(g):
    """Addresses of a batched graph's four int32 CSR arrays (indptr, indices, t_indptr, t_indices), cached on the graph
    together with the arrays they belong to (a replaced array invalidates the entry)."""
    c = getattr(g, "_gmeta_ptrs", None)
    arrs = (g.indptr, g.indices, g.t_indptr, g.t_indices)
    if c is not None and all(a is b for a, b in zip(c[0], arrs)):
        return c[1]
    fixed = tuple(_i32c(a) for a in arrs)
    p = tuple(a.ctypes.data for a in fixed)
    try:
        g._gmeta_ptrs = (arrs, p, fixed)
    except AttributeError:
        pass
    return p


def _ptr_i64(x, keep):
    """Address of `x` as a contiguous int64 array (torch tensor, numpy array or sequence); whatever had to be created for
    it is appended to `keep`.  The common case -- an int64 torch tensor or numpy array -- costs one attribute call: the
    packer threads run beside the thread that launches the steps, and every microsecond they spend in the interpreter is
    taken from it."""
    if isinstance(x, torch.Tensor):
        if x.dtype == torch.int64 and x.is_contiguous() and x.device.type == "cpu":
            return x.data_ptr()
        x = x.cpu().numpy()
    if not (isinstance(x, np.ndarray) and x.dtype == np.int64 and x.flags.c_contiguous):
        x = np.ascontiguousarray(x, dtype=np.int64)
        keep.append(x)
    return x.ctypes.data


def _plan_base(graphs, centres):
    ps = PackedSetHost()
    ps.T = len(graphs)
    ns = np.array([g.n_nodes for g in graphs], dtype=np.int64)
    es = np.array([g.n_edges for g in graphs], dtype=np.int64)
    ss = np.array([len(g.batch_num_nodes) for g in graphs], dtype=np.int64)
    ps.node_off = np.concatenate([[0], np.cumsum(ns)])
    ps.edge_off = np.concatenate([[0], np.cumsum(es)])
    ps.sub_off = np.concatenate([[0], np.cumsum(ss)])
    ps.N, ps.E, ps.S = int(ps.node_off[-1]), int(ps.edge_off[-1]), int(ps.sub_off[-1])
    ps.max_rows_per_task = int(ss.max()) if ps.T else 0
    c0 = centres[0]
    ps.cps = 2 if (hasattr(c0, "dim") and c0.dim() == 2) or (isinstance(c0, np.ndarray) and c0.ndim == 2) else 1
    ps.tiles = tile_table(ps.node_off.astype(np.int64))
    ps.n_tiles = int(ps.tiles[0].shape[0])
    ps.sizes = {"indptr": ps.N + 1, "indices": ps.E, "t_indptr": ps.N + 1, "t_indices": ps.E,
                "tile_row0": ps.n_tiles, "tile_nrows": ps.n_tiles, "tile_task": ps.n_tiles,
                "task_row_ptr": ps.T + 1, "task_sub_ptr": ps.T + 1, "centre_row": ps.S * ps.cps,
                "feat_row": ps.N, "labels": ps.S, "centre_pos": ps.S * ps.cps}
    return ps


def _act_capacity(ps, n_layers):
    """Upper bound (int32 elements) of the active-row segments of one set."""
    per_layer = _al(ps.N) + _al(ps.T + 1) + 3 * _al(ps.N // 128 + ps.T + 1)
    return n_layers * per_layer


def _fill_base(buf, ps, graphs, labels, centres, node_ids, graph_idx, graph_row_off, lib, n_threads, slim=False):
    """`slim`: only what the host alone knows (CSR by destination, centres, labels, feature rows, task pointers);
    the CSR by source and the tile tables are left to gmeta_packed_set_finish on the device."""
    import ctypes as C
    o, N, E, T = ps.off, ps.N, ps.E, ps.T
    ptrs = [(C.c_void_p * max(T, 1))(*col) for col in zip(*[_csr_ptrs(g) for g in graphs])] if T else [None] * 4
    base = buf.__array_interface__['data'][0]
    if slim:
        rc = lib.gmeta_host_pack_csr(T, ptrs[0], ptrs[1], None, None, ps.node_off.ctypes.data, ps.edge_off.ctypes.data,
                                     base + 4 * o["indptr"], base + 4 * o["indices"], None, None, n_threads)
    else:
        rc = lib.gmeta_host_pack_csr(T, ptrs[0], ptrs[1], ptrs[2], ptrs[3], ps.node_off.ctypes.data, ps.edge_off.ctypes.data,
                                     base + 4 * o["indptr"], base + 4 * o["indices"], base + 4 * o["t_indptr"],
                                     base + 4 * o["t_indices"], n_threads)
    if rc != 0:
        raise RuntimeError("gmeta_host_pack_csr failed (%d)" % rc)
    # centre rows, labels, task pointers and the row tiles: one library call (no interpreter lock held meanwhile)
    keep64 = []
    bnn = [g.batch_num_nodes if (isinstance(g.batch_num_nodes, np.ndarray) and g.batch_num_nodes.dtype == np.int64)
           else np.asarray(g.batch_num_nodes, dtype=np.int64) for g in graphs]
    n_sub_arr = np.array([b.shape[0] for b in bnn], dtype=np.int32)
    pa = lambda xs: (C.c_void_p * max(T, 1))(*[_ptr_i64(x, keep64) for x in xs])                        # noqa: E731
    vpa64 = pa
    cen, lab = centres, labels
    seg = lambda k: base + 4 * o[k]                                                                      # noqa: E731
    rc = lib.gmeta_host_pack_small(T, ps.cps, vpa64(bnn), n_sub_arr.ctypes.data, vpa64(cen), vpa64(lab),
                                   ps.node_off.ctypes.data, ps.sub_off.ctypes.data, seg("centre_row"), seg("labels"),
                                   seg("task_row_ptr"), seg("task_sub_ptr"), None if slim else seg("tile_row0"),
                                   None if slim else seg("tile_nrows"), None if slim else seg("tile_task"))
    if rc != 0:
        raise RuntimeError("gmeta_host_pack_small failed (%d): centres / labels / batch_num_nodes do not match the "
                           "batched graphs" % rc)
    single_graph = graph_row_off.shape[0] == 1
    all_ids = []
    for t, g in enumerate(graphs):
        ids = getattr(g, "parent_ids", None)
        # the ids a batch concatenated at batch time replace the id lists (meta.py:119-120 reads n_spt / n_qry) only
        # when the lists ARE the objects the batch was built from -- re-sampled, permuted or filtered id lists are
        # different objects and are honoured through the slow path
        src = getattr(g, "parent_id_lists", None)
        same = src is not None and (src is node_ids[t] or (
            len(src) == len(node_ids[t]) and all(a is b for a, b in zip(src, node_ids[t]))))
        if (not same or ids is None or ids.shape[0] != g.n_nodes or ids.dtype != np.int64
                or not ids.flags.c_contiguous):
            ids = np.ascontiguousarray(_flat_ids(node_ids[t]))                # meta.py:119-120
        all_ids.append(ids)
    vpa = lambda arrs: (C.c_void_p * max(T, 1))(*[a.__array_interface__['data'][0] for a in arrs])   # noqa: E731
    if single_graph:
        sub_ptr = sub_goff = n_sub = None
        rc = lib.gmeta_host_pack_feat_rows(T, vpa(all_ids), None, None, None, ps.node_off.ctypes.data,
                                           base + 4 * o["feat_row"], n_threads)
    else:
        sub_ptr = [np.concatenate([[0], np.cumsum(b)]).astype(np.int64) for b in bnn]
        sub_goff = [np.ascontiguousarray(graph_row_off[np.asarray(graph_idx[t], dtype=np.int64)], dtype=np.int64)
                    for t in range(T)]
        n_sub = np.array([len(b) for b in bnn], dtype=np.int32)
        rc = lib.gmeta_host_pack_feat_rows(T, vpa(all_ids), vpa(sub_ptr), vpa(sub_goff), n_sub.ctypes.data,
                                           ps.node_off.ctypes.data, base + 4 * o["feat_row"], n_threads)
    if rc != 0:
        raise RuntimeError("gmeta_host_pack_feat_rows failed (%d)" % rc)


def _plan_fill_act(buf, ps, off, n_layers, lib=None, staging=None):
    """Active rows per layer from the packed arrays (global sorted row ids are grouped by task because a task
    is a contiguous row range), their task pointers and tile tables, appended at `off`."""
    o = ps.off
    ps.n_layers = n_layers
    ps.act = [{} for _ in range(n_layers)]
    if n_layers == 0:
        return off
    indptr = buf[o["indptr"]:o["indptr"] + ps.N + 1]
    indices = buf[o["indices"]:o["indices"] + ps.E]
    centre = buf[o["centre_row"]:o["centre_row"] + ps.S * ps.cps]
    if lib is not None:
        # every layer's rows, task pointers, tile tables and the centre positions in ONE library call
        flags = staging.byte_map(ps.N) if staging is not None else np.zeros(ps.N, dtype=np.uint8)
        scratch = staging.i64_scratch(2 * (ps.N + 1)) if staging is not None else np.empty(2 * (ps.N + 1), dtype=np.int64)
        seg_off = np.zeros(n_layers * 5, dtype=np.int64)
        seg_n = np.zeros(n_layers * 5, dtype=np.int64)
        base = buf.__array_interface__['data'][0]
        end = lib.gmeta_host_active_rows(indptr.ctypes.data, indices.ctypes.data, ps.N, centre.ctypes.data, ps.S * ps.cps,
                                         ps.node_off.ctypes.data, ps.T, n_layers, flags.ctypes.data, scratch.ctypes.data,
                                         base, off, seg_off.ctypes.data, seg_n.ctypes.data, base + 4 * o["centre_pos"])
        if end < 0:
            raise RuntimeError("gmeta_host_active_rows failed (%d)" % end)
        for l in range(n_layers):
            for j, k in enumerate(("act_rows", "act_task_ptr", "act_tile_row0", "act_tile_nrows", "act_tile_task")):
                key = "%s%d" % (k, l)
                o[key] = int(seg_off[l * 5 + j])
                ps.sizes[key] = int(seg_n[l * 5 + j])
            ps.act[l] = {"n": int(seg_n[l * 5]), "n_tiles": int(seg_n[l * 5 + 2])}
        return int(end)
    rows = np.unique(centre.astype(np.int64))     # a few thousand centres: sorting is cheap
    per_layer = [None] * n_layers
    per_layer[n_layers - 1] = rows
    flags = None
    for l in range(n_layers - 1, 0, -1):
        # sorted distinct in-neighbours of ~10^3-10^5 rows through a byte map: cheaper than sorting them
        if flags is None:
            # byte map over the rows: kept (zeroed) in the staging object between steps when there is one
            flags = staging.byte_map(ps.N) if staging is not None else np.zeros(ps.N, dtype=np.uint8)
        if lib is not None:
            rows = np.ascontiguousarray(rows, dtype=np.int64)
            tot = int((indptr[rows + 1].astype(np.int64) - indptr[rows]).sum())
            scratch = np.empty(min(ps.N, tot) + 1, dtype=np.int64)
            n = lib.gmeta_host_active_in_neighbours(indptr.ctypes.data, indices.ctypes.data, rows.ctypes.data,
                                                    rows.shape[0], ps.N, flags.ctypes.data, scratch.ctypes.data)
            if n < 0:
                raise RuntimeError("gmeta_host_active_in_neighbours failed (%d)" % n)
            rows = scratch[:n].copy()
        else:
            flags[_rows_concat(indptr, indices, rows)] = 1
            rows = np.flatnonzero(flags)
            flags[rows] = 0
        per_layer[l - 1] = rows
    for l in range(n_layers):
        rows = per_layer[l]
        tptr = np.searchsorted(rows, ps.node_off).astype(np.int64)
        tiles = tile_table(tptr)
        a = {"rows": rows, "task_ptr": tptr, "tiles": tiles, "n": int(rows.shape[0]), "n_tiles": int(tiles[0].shape[0])}
        ps.act[l] = a
        for k, arr, n in (("act_rows", rows, a["n"]), ("act_task_ptr", tptr, ps.T + 1), ("act_tile_row0", tiles[0], a["n_tiles"]),
                          ("act_tile_nrows", tiles[1], a["n_tiles"]), ("act_tile_task", tiles[2], a["n_tiles"])):
            key = "%s%d" % (k, l)
            ps.sizes[key] = n
            o[key] = off
            buf[off:off + n] = arr
            off += _al(n)
    buf[o["centre_pos"]:o["centre_pos"] + ps.S * ps.cps] = np.searchsorted(per_layer[n_layers - 1], centre)
    return off


def pack_meta_batch(staging, batch, graph_row_off, n_layers, lib, n_threads=0):
    """Pack both sets of a collated meta-batch into the staging buffer; returns (ps_spt, ps_qry, n_int32)."""
    x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry = batch
    ps_s, ps_q = _plan_base(x_spt, c_spt), _plan_base(x_qry, c_qry)
    off = 0
    for ps in (ps_s, ps_q):
        for k, n in ps.sizes.items():
            ps.off[k] = off
            off += _al(n)
    buf = staging.reserve(off + _act_capacity(ps_s, n_layers) + _act_capacity(ps_q, n_layers))
    _fill_base(buf, ps_s, x_spt, y_spt, c_spt, n_spt, g_spt, graph_row_off, lib, n_threads)
    _fill_base(buf, ps_q, x_qry, y_qry, c_qry, n_qry, g_qry, graph_row_off, lib, n_threads)
    off = _plan_fill_act(buf, ps_s, off, n_layers, lib, staging)
    off = _plan_fill_act(buf, ps_q, off, n_layers, lib, staging)
    ps_s.end = ps_q.end = off
    return ps_s, ps_q, off


HOST_SEGS = ("indptr", "indices", "centre_row", "feat_row", "labels", "task_sub_ptr", "task_row_ptr")


def pack_meta_batch_slim(staging, batch, graph_row_off, n_layers, lib, n_threads=0):
    """The host half of a meta-batch for Meta.upload_batch: only the segments the host alone knows (HOST_SEGS) are
    packed -- they form the head of the buffer and are all that is copied to the device; every other segment (CSR by
    source, tile tables, active rows, centre positions) is laid out behind them at its upper-bound size and filled on
    the device by `finish_on_device`.  Same values, segment by segment, as pack_meta_batch
    (tests/test_gpu_meta.py::test_slim_pack_equals_full_pack).  Returns (ps_spt, ps_qry, n_host_int32, n_total_int32)."""
    x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry = batch
    sets = [_plan_base(x_spt, c_spt), _plan_base(x_qry, c_qry)]
    off = 0
    for ps in sets:
        ps.n_layers = n_layers
        for k in HOST_SEGS:
            ps.off[k] = off
            off += _al(ps.sizes[k])
    n_host = off
    for ps in sets:                       # realised counts of both sets: adjacent, read back with one copy
        ps.off["counts"] = off
        off += _al(2 + 2 * n_layers)
    for ps in sets:
        cap_t = (ps.N + 127) // 128 + ps.T
        ps.cap = {"t_indptr": ps.N + 1, "t_indices": ps.E, "tile_row0": cap_t, "tile_nrows": cap_t, "tile_task": cap_t,
                  "centre_pos": ps.S * ps.cps}
        for l in range(n_layers):
            ps.cap["act_rows%d" % l] = ps.N
            ps.cap["act_task_ptr%d" % l] = ps.T + 1
            for k in ("act_tile_row0", "act_tile_nrows", "act_tile_task"):
                ps.cap["%s%d" % (k, l)] = cap_t
        for k, n in ps.cap.items():
            ps.off[k] = off
            off += _al(n)
    buf = staging.reserve(off)
    _fill_base(buf, sets[0], x_spt, y_spt, c_spt, n_spt, g_spt, graph_row_off, lib, n_threads, slim=True)
    _fill_base(buf, sets[1], x_qry, y_qry, c_qry, n_qry, g_qry, graph_row_off, lib, n_threads, slim=True)
    sets[0].end = sets[1].end = off
    return sets[0], sets[1], n_host, off


def finish_on_device(lib, ints, sets, n_layers, workspace, stream):
    """Enqueue gmeta_packed_set_finish for both sets of a slim-packed batch living in the device tensor `ints`;
    `workspace(nbytes)` returns a 256-byte aligned device pointer valid until the passes have run."""
    import ctypes as C
    from . import _lib
    base = ints.data_ptr()
    for ps in sets:
        seg = lambda k: base + 4 * ps.off[k]                                 # noqa: E731
        arr = lambda k: (C.c_void_p * max(n_layers, 1))(*[seg("%s%d" % (k, l)) for l in range(n_layers)])  # noqa: E731
        nb = lib.gmeta_packed_set_finish_workspace_bytes(ps.N, ps.E, n_layers)
        _lib.check(lib.gmeta_packed_set_finish(seg("indptr"), seg("indices"), ps.N, ps.E, seg("task_row_ptr"), None, ps.T,
                                               seg("centre_row"), ps.S * ps.cps, n_layers, seg("t_indptr"), seg("t_indices"),
                                               seg("task_row_ptr"), seg("tile_row0"), seg("tile_nrows"), seg("tile_task"),
                                               arr("act_rows"), arr("act_task_ptr"), arr("act_tile_row0"),
                                               arr("act_tile_nrows"), arr("act_tile_task"), seg("centre_pos"), seg("counts"),
                                               workspace(nb), nb, stream), "packed_set_finish")


def apply_counts(sets, n_layers, counts, counts_off):
    """Realised tile / active-row counts (host copy of the counts segments starting at int32 offset `counts_off`)
    -> the size fields of the sets."""
    for ps in sets:
        cnt = counts[ps.off["counts"] - counts_off:][:2 + 2 * n_layers]
        assert int(cnt[0]) == ps.n_tiles, "tile table of the device pass disagrees with the host's"
        ps.act = [{"n": int(cnt[2 + l]), "n_tiles": int(cnt[2 + n_layers + l])} for l in range(n_layers)]
        for l in range(n_layers):
            ps.sizes["act_rows%d" % l] = ps.act[l]["n"]
            ps.sizes["act_task_ptr%d" % l] = ps.T + 1
            for k in ("act_tile_row0", "act_tile_nrows", "act_tile_task"):
                ps.sizes["%s%d" % (k, l)] = ps.act[l]["n_tiles"]


class Staging(object):
    """Grow-only pinned host buffer + device buffer for the packed integer arrays.  `copied` is the CUDA event of the
    last host->device copy out of the pinned buffer: the buffer must not be rewritten (the packer uses non-temporal
    stores straight into it) before that copy has finished -- `reserve` waits for it."""

    def __init__(self, device):
        self.device = device
        self.host = None
        self.dev = None
        self.copied = None

    def reserve(self, n):
        if self.copied is not None:
            self.copied.synchronize()          # the previous DMA out of this pinned buffer is done
        if self.host is None or self.host.numel() < n:
            cap = int(n * 1.25) + 1024
            self.host = torch.empty(cap, dtype=torch.int32, pin_memory=torch.cuda.is_available())
            self.dev = torch.empty(cap, dtype=torch.int32, device=self.device)
        return self.host.numpy()

    def finish_workspace(self, nbytes):
        """256-byte aligned device scratch of the slot for gmeta_packed_set_finish (grow-only)."""
        if getattr(self, "_fws", None) is None or self._fws.numel() < nbytes + 256:
            self._fws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return (self._fws.data_ptr() + 255) // 256 * 256

    def counts_buffer(self, n):
        """Pinned host landing zone for the realised counts of a batch (a few dozen ints)."""
        if getattr(self, "_counts", None) is None or self._counts.numel() < n:
            self._counts = torch.empty(max(n, 64), dtype=torch.int32, pin_memory=torch.cuda.is_available())
        return self._counts

    def i64_scratch(self, n):
        if getattr(self, "_i64", None) is None or self._i64.shape[0] < n:
            self._i64 = np.empty(int(n * 1.25) + 64, dtype=np.int64)
        return self._i64

    def byte_map(self, n):
        """Zeroed uint8 scratch of at least n entries; users hand it back zeroed."""
        if getattr(self, "_flags", None) is None or self._flags.shape[0] < n:
            self._flags = np.zeros(int(n * 1.25) + 64, dtype=np.uint8)
        return self._flags

    def upload(self, n, stream=None):
        """Async copy of the first n ints to the device buffer on `stream` (default: the current stream); returns the
        bytes moved.  The event recorded behind the copy guards the pinned buffer (reserve) and is what a consumer on
        another stream waits for."""
        st = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            self.dev[:n].copy_(self.host[:n], non_blocking=True)
            if self.copied is None:
                self.copied = torch.cuda.Event()
            self.copied.record(st)
        return n * 4

"""Helpers for the -m gpu parity tests: call the C ABI with torch device tensors."""
import numpy as np
import torch

from gmeta_b200 import _lib
from gmeta_b200.learner import tile_table
from gmeta_b200.packed import csr_transpose


def dev():
    return torch.device('cuda', torch.cuda.current_device())


def stream():
    return torch.cuda.current_stream().cuda_stream


# Device tensors made by i32()/f32() are kept alive until the end of the test: the C ABI takes raw
# pointers, so a temporary freed right after .data_ptr() could be recycled by torch's caching
# allocator for the next argument of the same call.
_KEEP = []


def i32(a):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(dev())
    _KEEP.append(t)
    return t


def f32(a):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev())
    _KEEP.append(t)
    return t


def p(t):
    return None if t is None else t.data_ptr()


class DevGraph(object):
    """CSR (by destination) + transpose + tile table on the device for T tasks."""

    def __init__(self, src, dst, n, task_row_ptr=None):
        src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
        order = np.argsort(dst, kind="stable")
        indptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(dst, minlength=n), out=indptr[1:])
        indices = src[order].astype(np.int32)
        t_indptr, t_indices = csr_transpose(indptr, indices, n)
        if task_row_ptr is None:
            task_row_ptr = np.array([0, n])
        self.task_row_ptr_h = np.asarray(task_row_ptr, dtype=np.int64)
        row0, nrows, task = tile_table(self.task_row_ptr_h)
        self.n, self.T, self.n_tiles = n, len(task_row_ptr) - 1, len(row0)
        self.indptr, self.indices, self.t_indptr, self.t_indices = i32(indptr), i32(indices), i32(t_indptr), i32(t_indices)
        self.tile_row0, self.tile_nrows, self.tile_task = i32(row0), i32(nrows), i32(task)
        self.task_row_ptr = i32(task_row_ptr)
        self.norm = torch.empty(n, dtype=torch.float32, device=dev())
        _lib.check(_lib.lib().gmeta_degree_norm(p(self.indptr), n, p(self.norm), stream()))


def layer_fwd(g, x, W, b, f_in, f_out, relu=1, transposed=False, trans_w=0, mask=None, row_map=None,
              w_stride=0, b_stride=0, ldw=None, impl=_lib.IMPL_SIMT, ld_out=None, dst_rows=None, tiles=None):
    """tiles = (row0, nrows, task) device tensors over the compact row list when dst_rows is given."""
    ld_out = ld_out or (f_out + 3) // 4 * 4
    n_out_rows = g.n if dst_rows is None else dst_rows.shape[0]
    out = torch.full((n_out_rows, ld_out), float('nan'), dtype=torch.float32, device=dev())
    t_row0, t_nrows, t_task = (g.tile_row0, g.tile_nrows, g.tile_task) if tiles is None else tiles
    ip, ix = (g.t_indptr, g.t_indices) if transposed else (g.indptr, g.indices)
    if ldw is None:
        ldw = f_in if trans_w else f_out
    if impl == _lib.IMPL_TCPAIR:
        return _layer_fwd_pair(g, x, W, b, f_in, f_out, relu, trans_w, mask, row_map, w_stride, b_stride, ldw, ld_out,
                               dst_rows, (t_row0, t_nrows, t_task), ip, ix, out)
    nb = _lib.lib().gmeta_gcn_layer_fwd_workspace_bytes(g.T, w_stride, f_in, f_out, impl)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev())
    rc = _lib.lib().gmeta_gcn_layer_fwd(p(x), x.shape[1], p(row_map), p(dst_rows), p(ip), p(ix), p(g.norm), p(t_row0),
                                        p(t_nrows), p(t_task), t_row0.shape[0], g.T, p(W), w_stride, ldw,
                                        trans_w, p(b), b_stride, f_in, f_out, relu, p(mask), p(out), ld_out, impl,
                                        p(ws), nb, stream())
    _lib.check(rc, "gcn_layer_fwd")
    return out


def _layer_fwd_pair(g, x, W, b, f_in, f_out, relu, trans_w, mask, row_map, w_stride, b_stride, ldw, ld_out, dst_rows,
                    tiles, ip, ix, out):
    """The CTA-pair tensor-core path through gmeta_gcn_layer_fwd_ex: per-row abs-max of the input, plan built
    inside the call, and the row abs-max it emits for the next layer checked against the output."""
    L = _lib.lib()
    t_row0, t_nrows, t_task = tiles
    n_rows, n_edges = out.shape[0], int(ix.shape[0])
    rmax_in = torch.empty(x.shape[0], dtype=torch.float32, device=dev())
    _lib.check(L.gmeta_row_absmax(p(x), x.shape[1], x.shape[0], f_in, p(rmax_in), stream()), "row_absmax")
    rmax_out = torch.full((n_rows,), float('nan'), dtype=torch.float32, device=dev())
    nb = L.gmeta_gcn_layer_fwd_ex_workspace_bytes(g.T, w_stride, t_row0.shape[0], n_rows, n_edges, f_in, f_out,
                                                  _lib.IMPL_TCPAIR)
    ws = torch.empty(max(nb, 16) + 256, dtype=torch.uint8, device=dev())
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    rc = L.gmeta_gcn_layer_fwd_ex(p(x), x.shape[1], p(row_map), p(dst_rows), p(ip), p(ix), p(g.norm), p(t_row0),
                                  p(t_nrows), p(t_task), t_row0.shape[0], g.T, p(W), w_stride, ldw, trans_w, p(b),
                                  b_stride, f_in, f_out, relu, p(mask), p(out), ld_out, _lib.IMPL_TCPAIR, ws_ptr, nb,
                                  n_rows, n_edges, p(rmax_in), p(rmax_out), None, stream())
    _lib.check(rc, "gcn_layer_fwd_ex")
    torch.cuda.synchronize()
    assert torch.equal(rmax_out, out[:, :f_out].abs().amax(1)), "row abs-max emitted for the next layer"
    return out


def layer_wgrad(g, x, dz, f_in, f_out, row_map=None, dst_rows=None, task_ptr=None):
    L = _lib.lib()
    T = g.T
    dW = torch.full((T, f_in, f_out), float('nan'), dtype=torch.float32, device=dev())
    db = torch.full((T, f_out), float('nan'), dtype=torch.float32, device=dev())
    nbytes = L.gmeta_gcn_layer_wgrad_workspace_bytes(T, f_in, f_out)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    rc = L.gmeta_gcn_layer_wgrad(p(x), x.shape[1], p(row_map), p(dst_rows), p(g.indptr), p(g.indices), p(g.norm),
                                 p(g.task_row_ptr if task_ptr is None else task_ptr), T, p(dz), dz.shape[1], f_in, f_out, p(dW), f_in * f_out, p(db),
                                 f_out, p(ws), nbytes, stream())
    _lib.check(rc, "gcn_layer_wgrad")
    return dW, db


def report(name, got, want, atol, rtol=0.0):
    got = got.detach().cpu().numpy() if hasattr(got, "detach") else np.asarray(got)
    want = want.detach().cpu().numpy() if hasattr(want, "detach") else np.asarray(want)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    lim = atol + rtol * np.abs(want)
    bad = ~(err <= lim)
    msg = "%s: max|err|=%.3e (ref max %.3e), %d/%d outside atol=%g rtol=%g" % (
        name, float(np.nanmax(err)) if err.size else 0.0, float(np.abs(want).max()) if want.size else 0.0,
        int(bad.sum()), err.size, atol, rtol)
    print(msg)
    assert not bad.any(), msg

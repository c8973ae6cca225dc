"""Drop-in for the reference's G-Meta/learner.py: `Classifier(config)` with the same
constructor, parameter order/initialisers, `forward(g, to_fetch, features, vars=None)`,
`parameters()` and `zero_grad(vars=None)` (learner.py:69-209) -- the arithmetic runs in the
hand-written sm_100a kernels behind the C ABI (include/gmeta_b200.h), including under
`torch.autograd.grad` (meta.py:125,149 style callers) through an autograd.Function.

`g` is a `PackedSubgraphBatch` (the stand-in for the DGL batched graph the reference gets from
`dgl.batch`, subgraph_data_processing.py:399-406).  There is no CPU path: tensors are moved to
the CUDA device (as learner.py:145 does) and a missing extension raises.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init

from . import _lib
from ._lib import TILE_ROWS

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')   # learner.py:10


def _round_up(x, m):
    return (x + m - 1) // m * m


class ModelSpec(object):
    """`config` list (train.py:67-75) -> layer dims and offsets into the flat parameter buffer."""

    def __init__(self, config):
        self.link_pred = config[-1][0] == 'LinkPred'                      # learner.py:78-79
        self.aggregation = _lib.AGG_GCN                                   # the reference's GraphConv; see Classifier
        self.conv = [tuple(p) for n, p in config if n == 'GraphConv']
        lin = [tuple(p) for n, p in config if n == 'Linear']
        if len(lin) != 1 or not 1 <= len(self.conv) <= _lib.MAX_LAYERS:
            raise ValueError("config must hold 1..%d 'GraphConv' entries and one 'Linear'" % _lib.MAX_LAYERS)
        if any(n not in ('GraphConv', 'Linear', 'LinkPred') for n, _ in config):
            raise ValueError("only 'GraphConv', 'Linear' and 'LinkPred' config entries are supported")
        self.hid = self.conv[-1][1]
        self.n_out = lin[0][1]
        self.lin_in = lin[0][0] * (2 if self.link_pred else 1)
        # creation order of learner.py:81-97: config order, (W, b) pairs
        self.shapes, self.offsets, off = [], [], 0
        for name, p in config:
            if name == 'GraphConv':
                shapes = [(p[0], p[1]), (p[1],)]
            elif name == 'Linear':
                shapes = [(p[1], self.lin_in), (p[1],)]
            else:
                continue
            for s in shapes:
                self.shapes.append(s)
                self.offsets.append(off)
                off = _round_up(off + int(np.prod(s)), 4)
        self.n_params_padded = off
        self.order = [n for n, _ in config if n in ('GraphConv', 'Linear')]

    def c_model(self):
        m = _lib.Model()
        m.n_layers = len(self.conv)
        k = 0
        li = 0
        for name in self.order:
            if name == 'GraphConv':
                m.f_in[li], m.f_out[li] = self.conv[li]
                m.w_off[li], m.b_off[li] = self.offsets[k], self.offsets[k + 1]
                li += 1
            else:
                m.wlin_off, m.blin_off = self.offsets[k], self.offsets[k + 1]
            k += 2
        m.n_out, m.link_pred, m.n_params_padded = self.n_out, int(self.link_pred), self.n_params_padded
        m.aggregation = int(getattr(self, "aggregation", _lib.AGG_GCN))
        return m

    def flatten(self, tensors, out=None):
        if out is None:
            out = torch.zeros(self.n_params_padded, dtype=torch.float32, device=tensors[0].device)
        for t, off, s in zip(tensors, self.offsets, self.shapes):
            out[off:off + t.numel()].copy_(t.detach().reshape(-1))
        return out

    def unflatten(self, flat):
        return [flat[off:off + int(np.prod(s))].view(s) for off, s in zip(self.offsets, self.shapes)]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def tile_table(task_row_ptr):
    """Row tiles of <= TILE_ROWS rows that never straddle a task."""
    n = np.diff(task_row_ptr).astype(np.int64)
    nt = (n + TILE_ROWS - 1) // TILE_ROWS
    task = np.repeat(np.arange(n.shape[0], dtype=np.int32), nt)
    first = np.concatenate([[0], np.cumsum(nt)])[:-1]
    k = np.arange(int(nt.sum()), dtype=np.int64) - np.repeat(first, nt)
    row0 = np.repeat(task_row_ptr[:-1].astype(np.int64), nt) + k * TILE_ROWS
    nrows = np.minimum(TILE_ROWS, np.repeat(task_row_ptr[1:].astype(np.int64), nt) - row0)
    return row0.astype(np.int32), nrows.astype(np.int32), task


class _DeviceGraph(object):
    """Device copy of one PackedSubgraphBatch as a single-task packed set (cached on the batch)."""

    def __init__(self, g, dev, aggregation=_lib.AGG_GCN):
        N = g.n_nodes
        self.aggregation = aggregation
        trp = np.array([0, N], dtype=np.int32)
        row0, nrows, task = tile_table(trp)
        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(dev)  # noqa: E731
        self.N, self.n_tiles = N, int(row0.shape[0])
        self.indptr, self.indices = i32(g.indptr), i32(g.indices)
        self.t_indptr, self.t_indices = i32(g.t_indptr), i32(g.t_indices)
        self.tile_row0, self.tile_nrows, self.tile_task = i32(row0), i32(nrows), i32(task)
        self.task_row_ptr = i32(trp)
        self.sub_off = torch.as_tensor(np.concatenate([[0], np.cumsum(g.batch_num_nodes)])[:-1]).to(dev)
        self.S = len(g.batch_num_nodes)
        self.task_sub_ptr = i32(np.array([0, self.S]))
        # norm: scale of a row as a source; norm_dst: as a destination (None = the same array: the reference's
        # symmetric GraphConv normalisation, learner.py:29-49)
        self.norm = torch.empty(N, dtype=torch.float32, device=dev)
        self.norm_dst = None
        if aggregation == _lib.AGG_GCN:
            _lib.check(_lib.lib().gmeta_degree_norm(_ptr(self.indptr), N, _ptr(self.norm), _stream()), "degree_norm")
        else:
            self.norm_dst = torch.empty(N, dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().gmeta_aggregation_norms(_ptr(self.indptr), N, aggregation, _ptr(self.norm),
                                                          _ptr(self.norm_dst), _stream()), "aggregation_norms")

    def scales(self, transposed):
        """(source scales, destination scales or None) of a layer call; a data gradient runs on the transposed graph
        with the two arrays swapped."""
        if transposed and self.norm_dst is not None:
            return self.norm_dst, self.norm
        return self.norm, self.norm_dst


def _plan(dg, transposed):
    """Structure-only layer plan of the CTA-pair tensor-core kernel (csrc/gcn_layer_pair.cu): per-row source
    records, tile pairs and the hub edge list.  One per orientation of the batched graph, built on first use and
    shared by every layer and every call on this graph."""
    key = "_plan_t" if transposed else "_plan"
    pl = getattr(dg, key, None)
    if pl is None:
        L = _lib.lib()
        ip, ix = (dg.t_indptr, dg.t_indices) if transposed else (dg.indptr, dg.indices)
        E = int(ix.shape[0])
        nb = L.gmeta_layer_plan_bytes(dg.n_tiles, 1, dg.N, E)
        buf = torch.empty(nb + 256, dtype=torch.uint8, device=dg.norm.device)
        ptr = (buf.data_ptr() + 255) // 256 * 256
        _lib.check(L.gmeta_layer_plan_build(_ptr(ip), _ptr(ix), _ptr(dg.scales(transposed)[0]), None, None, _ptr(dg.tile_row0),
                                            _ptr(dg.tile_nrows), _ptr(dg.tile_task), dg.n_tiles, 1, dg.N, E, ptr,
                                            _stream()), "layer_plan_build")
        pl = (buf, ptr)
        setattr(dg, key, pl)
    return pl[1]


def _row_absmax(x, f):
    r = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().gmeta_row_absmax(_ptr(x), x.shape[1], x.shape[0], f, _ptr(r), _stream()), "row_absmax")
    return r


def _layer(dg, inp, rmax_in, W, trans_w, ldw, b, fi, fo, relu, mask, out, impl, transposed=False, want_rmax=True):
    """One fused GCN layer over the whole batched graph through gmeta_gcn_layer_fwd_ex: with the input's per-row
    abs-max the library picks the CTA-pair tensor-core kernel where the shape allows (else 3xTF32 / FFMA) and
    returns the output's per-row abs-max for the next layer."""
    L = _lib.lib()
    ip, ix = (dg.t_indptr, dg.t_indices) if transposed else (dg.indptr, dg.indices)
    E = int(ix.shape[0])
    dev = inp.device
    use_ex = impl in (_lib.IMPL_AUTO, _lib.IMPL_TCPAIR) and rmax_in is not None
    rmax_out = torch.empty(dg.N, dtype=torch.float32, device=dev) if (use_ex and want_rmax) else None
    if use_ex:
        nb = L.gmeta_gcn_layer_fwd_ex_workspace_bytes(1, 0, dg.n_tiles, dg.N, E, fi, fo, impl)
    else:
        nb = L.gmeta_gcn_layer_fwd_workspace_bytes(1, 0, fi, fo, impl)
    scratch = torch.empty(max(nb, 16) + 256, dtype=torch.uint8, device=dev)
    sp = (scratch.data_ptr() + 255) // 256 * 256
    n_src, n_dst = dg.scales(transposed)
    _lib.check(L.gmeta_gcn_layer_fwd_nd(
        _ptr(inp), inp.shape[1], None, None, _ptr(ip), _ptr(ix), _ptr(n_src), _ptr(n_dst), _ptr(dg.tile_row0),
        _ptr(dg.tile_nrows), _ptr(dg.tile_task), dg.n_tiles, 1, _ptr(W), 0, ldw, trans_w, _ptr(b), 0, fi, fo, relu,
        _ptr(mask), _ptr(out), out.shape[1], impl, sp, nb, dg.N, E, _ptr(rmax_in) if use_ex else None,
        _ptr(rmax_out), _plan(dg, transposed) if use_ex else None, _stream()), "gcn_layer_fwd")
    return rmax_out


def _device_graph(g, dev, aggregation=_lib.AGG_GCN):
    dg = getattr(g, "_gmeta_dev", None)
    if dg is None or dg.norm.device != dev or dg.aggregation != aggregation:
        dg = _DeviceGraph(g, dev, aggregation)
        g._gmeta_dev = dg
    return dg


class _ClassifierFn(torch.autograd.Function):
    """logits = Linear(readout(GCN^h(features))) with per-call weights; backward = the kernels'
    own weight/data gradients (no autograd graph of small ops)."""

    @staticmethod
    def forward(ctx, spec, dg, centre_row, features, impl, *vars):
        L = _lib.lib()
        dev = features.device
        x = features.contiguous()
        acts, inp, ld_in = [], x, x.shape[1]
        n_conv = len(spec.conv)
        ws = [v.detach().contiguous() for v in vars]
        rmax = _row_absmax(x, spec.conv[0][0]) if impl in (_lib.IMPL_AUTO, _lib.IMPL_TCPAIR) else None
        for l, (fi, fo) in enumerate(spec.conv):
            ld_out = _round_up(fo, 4)
            out = torch.empty(dg.N, ld_out, dtype=torch.float32, device=dev)
            rmax = _layer(dg, inp, rmax, ws[2 * l], 0, fo, ws[2 * l + 1], fi, fo, 1, None, out, impl,
                          want_rmax=l + 1 < n_conv)
            acts.append(out)
            inp, ld_in = out, ld_out
        cps = 2 if spec.link_pred else 1
        logits = torch.empty(dg.S, spec.n_out, dtype=torch.float32, device=dev)
        _lib.check(L.gmeta_readout_linear_fwd(
            _ptr(acts[-1]), ld_in, spec.hid, _ptr(centre_row), cps, _ptr(dg.task_sub_ptr), 1, dg.S,
            _ptr(ws[2 * n_conv]), 0, _ptr(ws[2 * n_conv + 1]), 0, spec.n_out, _ptr(logits), _stream()),
            "readout_linear_fwd")
        ctx.spec, ctx.dg, ctx.centre_row, ctx.x, ctx.acts, ctx.ws, ctx.impl = spec, dg, centre_row, x, acts, ws, impl
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        L = _lib.lib()
        spec, dg, acts, ws, x = ctx.spec, ctx.dg, ctx.acts, ctx.ws, ctx.x
        dev = dlogits.device
        dlogits = dlogits.contiguous()
        n_conv = len(spec.conv)
        cps = 2 if spec.link_pred else 1
        grads = [torch.zeros_like(w) for w in ws]
        ld_top = acts[-1].shape[1]
        dz = torch.empty(dg.N, ld_top, dtype=torch.float32, device=dev)
        _lib.check(L.gmeta_readout_linear_bwd(
            _ptr(acts[-1]), ld_top, spec.hid, dg.N, None, _ptr(ctx.centre_row), cps, _ptr(dg.task_sub_ptr), 1, dg.S,
            _ptr(ws[2 * n_conv]), 0, spec.n_out, _ptr(dlogits), _ptr(grads[2 * n_conv]), 0,
            _ptr(grads[2 * n_conv + 1]), 0, _ptr(dz), _stream()), "readout_linear_bwd")
        for l in range(n_conv - 1, -1, -1):
            fi, fo = spec.conv[l]
            inp = x if l == 0 else acts[l - 1]
            nbytes = L.gmeta_gcn_layer_wgrad_workspace_bytes(1, fi, fo)
            wsb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.gmeta_gcn_layer_wgrad_nd(
                _ptr(inp), inp.shape[1], None, None, _ptr(dg.indptr), _ptr(dg.indices), _ptr(dg.norm), _ptr(dg.norm_dst),
                _ptr(dg.task_row_ptr), 1, _ptr(dz), dz.shape[1], fi, fo, _ptr(grads[2 * l]), 0,
                _ptr(grads[2 * l + 1]), 0, _ptr(wsb), nbytes, _stream()), "gcn_layer_wgrad")
            if l > 0:
                ld_lo = acts[l - 1].shape[1]
                dz_lo = torch.empty(dg.N, ld_lo, dtype=torch.float32, device=dev)
                rm = _row_absmax(dz, fo) if ctx.impl in (_lib.IMPL_AUTO, _lib.IMPL_TCPAIR) else None
                _layer(dg, dz, rm, ws[2 * l], 1, fo, None, fo, fi, 0, acts[l - 1], dz_lo, ctx.impl, transposed=True,
                       want_rmax=False)
                dz = dz_lo
        return (None, None, None, None, None) + tuple(grads)


class Classifier(nn.Module):
    def __init__(self, config, impl=_lib.IMPL_AUTO, aggregation="gcn"):
        """`aggregation`: "gcn" (the reference's symmetric-normalised GraphConv, the default and the only mode the
        reference has), "mean" (GraphSAGE-style mean over the in-neighbours) or "sum" -- the latter two have no
        reference counterpart (parity is against the oracle's restatement only)."""
        super(Classifier, self).__init__()
        self.vars = nn.ParameterList()
        self.config = config
        self.spec = ModelSpec(config)
        if aggregation not in _lib.AGGREGATIONS:
            raise ValueError("aggregation must be one of %s" % sorted(_lib.AGGREGATIONS))
        self.aggregation = aggregation
        self.spec.aggregation = _lib.AGGREGATIONS[aggregation]
        self.LinkPred_mode = self.spec.link_pred
        self.impl = impl
        for name, param in config:                                      # learner.py:81-97
            if name == 'Linear':
                w = nn.Parameter(torch.ones(param[1], param[0] * (2 if self.LinkPred_mode else 1)))
                init.kaiming_normal_(w)
                self.vars.append(w)
                self.vars.append(nn.Parameter(torch.zeros(param[1])))
            if name == 'GraphConv':
                w = nn.Parameter(torch.Tensor(param[0], param[1]))
                init.xavier_uniform_(w)
                self.vars.append(w)
                self.vars.append(nn.Parameter(torch.zeros(param[1])))

    def forward(self, g, to_fetch, features, vars=None):
        if vars is None:
            vars = self.vars
        if not torch.cuda.is_available():
            raise _lib.GMetaError("gmeta_b200 needs a CUDA device (there is no CPU path)")
        _lib.lib()
        dev = torch.device('cuda', torch.cuda.current_device())
        h = torch.as_tensor(features).float().to(dev)                    # learner.py:144-145
        dg = _device_graph(g, dev, self.spec.aggregation)
        to_fetch = torch.as_tensor(to_fetch).to(dev).long()
        if self.LinkPred_mode:                                           # learner.py:165-168
            centre = torch.stack((to_fetch[:, 0] + dg.sub_off, to_fetch[:, 1] + dg.sub_off), 1)
        else:                                                            # learner.py:170
            centre = to_fetch + dg.sub_off
        centre = centre.reshape(-1).to(torch.int32).contiguous()
        vars = [v if v.device == dev else v.to(dev) for v in vars]
        logits = _ClassifierFn.apply(self.spec, dg, centre, h, self.impl, *vars)
        return logits, logits                                            # learner.py:194

    def zero_grad(self, vars=None):                                      # learner.py:196-206
        with torch.no_grad():
            for p in (self.vars if vars is None else vars):
                if p.grad is not None:
                    p.grad.zero_()

    def parameters(self):                                                # learner.py:208-209
        return self.vars

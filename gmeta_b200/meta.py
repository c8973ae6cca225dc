"""Drop-in for the reference's G-Meta/meta.py: `Meta(args, config)` with `.forward(...)` /
`.finetunning(...)` of identical signature and return value (meta.py:236-244), plus the public
`euclidean_dist`, `proto_loss_spt`, `proto_loss_qry` (meta.py:14-79).

One call of `Meta.forward` packs the whole meta-batch (all tasks' support and query subgraph
batches) into one HBM-resident structure, and enqueues the complete first-order ProtoMAML
inner loop for every task at once through `gmeta_maml_step` (C++ driver over the sm_100a
kernels), followed -- across ranks -- by ONE all-reduce of [meta-grad | loss | accuracies] and
the fused Adam update.  Features are gathered on the device from a resident feature table by
parent id instead of on the CPU every step (meta.py:119-120).  No CPU fallback.
"""
import ctypes as C
from copy import deepcopy

import time

import numpy as np
import torch
from torch import nn

from . import _lib, packing
from .learner import Classifier, _ptr, _stream

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')   # meta.py:12


def _dev():
    if not torch.cuda.is_available():
        raise _lib.GMetaError("gmeta_b200 needs a CUDA device (there is no CPU path)")
    return torch.device('cuda', torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------
# public loss functions (meta.py:14-79) on the device
# ---------------------------------------------------------------------------------------------
def euclidean_dist(x, y):
    """Squared Euclidean distances between rows of x [N,D] and y [M,D] (meta.py:14-26)."""
    if x.size(1) != y.size(1):
        raise Exception                                                  # meta.py:20-21
    return torch.pow(x.unsqueeze(1) - y.unsqueeze(0), 2).sum(2)


def _labels_dev(y_t, dev):
    y = torch.as_tensor(y_t)
    return y.to(device=dev, dtype=torch.int32).contiguous()


class _ProtoSptFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, y_dev, n_support, n_classes):
        L = _lib.lib()
        dev, (S, D) = logits.device, logits.shape
        z = logits.detach().contiguous()
        ptr = torch.tensor([0, S], dtype=torch.int32, device=dev)
        cpos = torch.empty(S, dtype=torch.int32, device=dev)
        cocc = torch.empty(S, dtype=torch.int32, device=dev)
        ncls = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(L.gmeta_proto_label_prep(_ptr(y_dev), _ptr(ptr), 1, _ptr(cpos), _ptr(cocc), _ptr(ncls),
                                            _stream()), "proto_label_prep")
        protos = torch.zeros(n_classes, D, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        acc = torch.empty((), dtype=torch.float32, device=dev)
        dz = torch.empty_like(z)
        _lib.check(L.gmeta_proto_loss_spt(_ptr(z), D, _ptr(ptr), 1, _ptr(cpos), _ptr(cocc), _ptr(ncls),
                                          n_support, n_classes, S, 1.0, _ptr(protos), _ptr(loss),
                                          _ptr(acc), 1, _ptr(dz), _stream()), "proto_loss_spt")
        ctx.save_for_backward(dz, ptr, cpos, cocc)
        ctx.n_support, ctx.n_classes = n_support, n_classes
        ctx.mark_non_differentiable(acc)
        return loss, acc, protos

    @staticmethod
    def backward(ctx, g_loss, g_acc, g_protos):
        dz, ptr, cpos, cocc = ctx.saved_tensors
        grad = dz * g_loss
        if g_protos is not None:
            S, D = dz.shape
            extra = torch.empty_like(dz)
            gp = g_protos.contiguous()
            _lib.check(_lib.lib().gmeta_proto_grad_to_support(
                _ptr(gp), D, ctx.n_classes, _ptr(ptr), 1, _ptr(cpos), _ptr(cocc), ctx.n_support, S,
                _ptr(extra), _stream()), "proto_grad_to_support")
            grad = grad + extra
        return grad, None, None, None


class _ProtoQryFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, y_dev, prototypes):
        L = _lib.lib()
        dev, (S, D) = logits.device, logits.shape
        z = logits.detach().contiguous()
        P = prototypes.detach().contiguous()
        M = P.shape[0]
        ptr = torch.tensor([0, S], dtype=torch.int32, device=dev)
        cpos = torch.empty(S, dtype=torch.int32, device=dev)
        cocc = torch.empty(S, dtype=torch.int32, device=dev)
        ncls = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(L.gmeta_proto_label_prep(_ptr(y_dev), _ptr(ptr), 1, _ptr(cpos), _ptr(cocc), _ptr(ncls),
                                            _stream()), "proto_label_prep")
        nproto = torch.tensor([M], dtype=torch.int32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        acc = torch.empty((), dtype=torch.float32, device=dev)
        dz, dP = torch.empty_like(z), torch.empty_like(P)
        _lib.check(L.gmeta_proto_loss_qry(_ptr(z), D, _ptr(ptr), 1, _ptr(cpos), _ptr(nproto), _ptr(P), M, S,
                                          1.0, _ptr(loss), _ptr(acc), 1, _ptr(dz), _ptr(dP),
                                          _stream()), "proto_loss_qry")
        ctx.save_for_backward(dz, dP)
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g_loss, g_acc):
        dz, dP = ctx.saved_tensors
        return dz * g_loss, None, dP * g_loss


def proto_loss_spt(logits, y_t, n_support):
    """(loss, acc, prototypes) of meta.py:28-54, computed on the device (prototypes carry grad)."""
    dev = _dev()
    logits = logits.to(dev)
    y = torch.as_tensor(y_t)
    counts = torch.unique(y.cpu(), return_counts=True)[1]
    if int(counts.min()) < n_support:
        raise RuntimeError("stack expects each tensor to be equal size (meta.py:42)")
    return _ProtoSptFn.apply(logits, _labels_dev(y, dev), int(n_support), int(counts.numel()))


def proto_loss_qry(logits, y_t, prototypes):
    """(loss, acc) of meta.py:56-79 on the device."""
    dev = _dev()
    logits = logits.to(dev)
    y = torch.as_tensor(y_t)
    counts = torch.unique(y.cpu(), return_counts=True)[1]
    if int(counts.min()) != int(counts.max()):
        raise RuntimeError("stack expects each tensor to be equal size (meta.py:65)")
    return _ProtoQryFn.apply(logits, _labels_dev(y, dev), prototypes.to(dev))


# ---------------------------------------------------------------------------------------------
# device-resident state
# ---------------------------------------------------------------------------------------------
class FeatureTable(object):
    """All graphs' node features in one HBM table [sum_g N_g, ld] (ld = F0 rounded up to 4,
    zero padded), uploaded once and reused by every meta-step; row = graph_row_off[g] + node."""

    def __init__(self, feat, dev):
        feats = [np.asarray(f, dtype=np.float32) for f in feat]
        self.f0 = int(feats[0].shape[1])
        self.ld = (self.f0 + 3) // 4 * 4
        rows = np.array([f.shape[0] for f in feats], dtype=np.int64)
        self.graph_row_off = np.concatenate([[0], np.cumsum(rows)])[:-1]
        self.table = torch.zeros(int(rows.sum()), self.ld, dtype=torch.float32, device=dev)
        r = 0
        for f in feats:
            self.table[r:r + f.shape[0], :self.f0].copy_(torch.from_numpy(np.ascontiguousarray(f)))
            r += f.shape[0]
        self._keep = feat          # keeps id(feat) unique while cached
        # max |row| of the table, once: input scale bound of the CTA-pair tensor-core layer path
        self.rowmax = torch.empty(self.table.shape[0], dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().gmeta_row_absmax(self.table.data_ptr(), self.ld, self.table.shape[0], self.f0,
                                               self.rowmax.data_ptr(), _stream()), "row_absmax")


def _feat_fingerprint(feat):
    """Cheap identity + content fingerprint of the caller's feature arrays: length, and for a bounded sample of
    the arrays (first, last, up to eight evenly spaced) the data pointer, shape, dtype and a strided checksum of
    64 values.  The reference re-reads `feat` every step (meta.py:119-120); a resident table must notice when an
    array was swapped or rewritten in place.  `Meta.refresh_features` forces a re-upload for changes the sample
    cannot see."""
    n = len(feat)
    if n == 0:
        return (0,)
    fp = [n]
    for i in sorted(set([0, n - 1] + list(range(0, n, max(1, n // 8))))):
        a = np.asarray(feat[i])
        flat = a.reshape(-1)
        step = max(1, flat.shape[0] // 64)
        fp.append((i, a.__array_interface__['data'][0], a.shape, a.dtype.str,
                   float(np.asarray(flat[::step], dtype=np.float64).sum())))
    return tuple(fp)


def _keep_packers_off_the_step_cores():
    """Initialiser of the packer threads: when a rank has at least 8 host cores, the packers (and the C++ threads they
    start) stay off the first four cores of the rank's share, which are left to the thread that launches the steps and
    to the driver's own threads -- a launched step runs measurably slower on the DEVICE when every core is busy packing
    (DESIGN 4).  GMETA_B200_NO_PIN=1 disables it."""
    import os
    if os.environ.get("GMETA_B200_NO_PIN", "") not in ("", "0") or not hasattr(os, "sched_setaffinity"):
        return
    try:
        cores = sorted(os.sched_getaffinity(0))
        world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        rank = int(os.environ.get("LOCAL_RANK", "0")) % world
        per = len(cores) // world
        if per < 8:
            return
        mine = cores[rank * per:(rank + 1) * per]
        os.sched_setaffinity(0, mine[4:])
    except OSError:
        pass


class _DeviceBatch(object):
    """An uploaded meta-batch: segment plans of both sets + the device int32 buffer holding them."""
    __slots__ = ("T", "max_classes", "ft", "ps_s", "ps_q", "ints", "h2d_bytes", "resident", "ready", "pack_ms", "pending", "graph")


class FusedAdam(object):
    """`meta_optim` (meta.py:97): Adam(lr, betas=(0.9,0.999), eps=1e-8) over the flat parameter buffer with the
    reference's NaN-skip (meta.py:163-164).  Step count, gate and bias corrections live on the device
    (`gmeta_adam_step`): the count advances only when the update is applied -- like torch.optim.Adam, which the
    reference does not call on a skipped step -- and the launches are identical every step (CUDA-graph safe)."""

    def __init__(self, n_params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.betas, self.eps = lr, betas, eps
        self.n_params = n_params
        self.exp_avg = None
        self.exp_avg_sq = None
        self.state = None          # int32[8] on the device: [0] step count, [1] skipped flag of the last step

    def _state(self, dev):
        if self.exp_avg is None or self.exp_avg.device != dev:
            steps = 0 if self.state is None else int(self.state[0])
            self.exp_avg = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
            self.state = torch.zeros(8, dtype=torch.int32, device=dev)
            self.state[0] = steps

    @property
    def step_count(self):
        """Applied (non-skipped) steps so far; reads the device counter (synchronises)."""
        return 0 if self.state is None else int(self.state[0])

    def step(self, flat_param, flat_grad, loss_sum=None, loss_scale=1.0, acc_sums=None, step_out=None, grad_scale=1.0):
        """One Adam step on the flat buffer.  The update is skipped iff *loss_sum * loss_scale is NaN; with step_out
        (float[n_acc + 2]) the call also writes acc_sums * loss_scale, the gate value and the skipped flag."""
        self._state(flat_param.device)
        n_acc = 0 if acc_sums is None else int(acc_sums.numel())
        _lib.check(_lib.lib().gmeta_adam_step(
            _ptr(flat_param), _ptr(flat_grad), _ptr(self.exp_avg), _ptr(self.exp_avg_sq), self.n_params,
            self.lr, self.betas[0], self.betas[1], self.eps, _ptr(self.state), grad_scale,
            _ptr(loss_sum), loss_scale, _ptr(acc_sums), n_acc, _ptr(step_out), _stream()), "adam_step")

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "lr": self.lr, "betas": self.betas, "eps": self.eps}


class Meta(nn.Module):
    def __init__(self, args, config):
        super(Meta, self).__init__()
        self.update_lr = args.update_lr
        self.meta_lr = args.meta_lr
        self.n_way = args.n_way
        self.k_spt = args.k_spt
        self.k_qry = args.k_qry
        self.task_num = args.task_num
        self.update_step = args.update_step
        self.update_step_test = args.update_step_test
        self.impl = getattr(args, 'impl', _lib.IMPL_AUTO)
        # False (default): the backward skips rows whose gradient is structurally zero (only centre
        # rows are read out); True: back-propagate over every row like the reference's autograd.
        self.dense_backward = bool(getattr(args, 'dense_backward', False))
        # True (default): every forward computes only the rows the read-out depends on (exact: only centre
        # rows are read out, learner.py:166-170); False: every row of every layer, like the reference.
        self.pruned_forward = bool(getattr(args, 'pruned_forward', not self.dense_backward))

        # neighbourhood aggregation: "gcn" = the reference's GraphConv (default); "mean" / "sum" have no reference
        # counterpart (args.aggregation, train.py --aggregation)
        self.aggregation = str(getattr(args, 'aggregation', 'gcn'))
        self.net = Classifier(config, impl=self.impl, aggregation=self.aggregation)
        self.net = self.net.to(device)
        self.spec = self.net.spec
        self.meta_optim = FusedAdam(self.spec.n_params_padded, self.meta_lr)
        self.method = args.method

        self._staging = None
        self._feat_cache = None
        self._ws = None
        self._scratch = {}
        self._pool = None              # worker thread of `prefetch`
        self._copy_stream = None
        self._prefetched = None
        self._slot_i = 0
        self._theta_flat = None        # flat parameter buffer the net's parameters are views of
        self._aux_stream = None        # second stream for the query forwards (gmeta_step_args_t::aux_stream)
        self._graphs = {}              # captured meta-steps of device-resident batches
        self._step_graphs = None       # updatable graphs of host-batch steps (gmeta_step_graph_*): handles, round robin
        self._sg_i = 0
        self._cap_stream = None
        self._picked = None            # a prefetched batch picked up early, with its step graph prepared
        self._alloc_gen = 0            # bumped whenever a scratch / workspace buffer is re-allocated
        # replay device-resident meta-steps (step_device on a batch with its own buffer) from CUDA graphs
        self.use_graphs = bool(getattr(args, 'use_graphs', True))
        # host batches: True = the host packs only the CSR by destination and the device derives the rest behind the
        # copy (gmeta_packed_set_finish: 21 MB instead of 37 MB over the bus per C2 batch, 40% less packing);
        # False = everything packed on the host.  The device passes compete with the step they run beside (1 GPU, C2:
        # 5.2 vs 4.6 ms per step with the host packer), but with several ranks per box the HOST is what a step waits
        # for (8 ranks on 32 cores: 11 ms to pack a batch on one thread), so the default follows the cores a rank has.
        df = getattr(args, 'device_finish', None)
        if df is None:
            import os
            env = os.environ.get("GMETA_B200_DEVICE_FINISH")
            if env is not None:
                df = env not in ("", "0")
            else:
                df = (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))) < 8
        self.device_finish = bool(df)
        self.pack_workers = 3          # packer threads of `prefetch` (the callers' three-batch lookahead keeps them busy)
        # host batches: run the step as an updatable CUDA graph prepared one batch ahead (gmeta_step_graph_*)
        self.graph_host_batches = bool(getattr(args, 'graph_host_batches', True))
        self.prepare_ahead = True
        self.two_streams = bool(getattr(args, 'two_streams', True))
        self.last = {}                 # diagnostics of the most recent call (loss, launches, bytes)
        self.return_meta_grad = False  # tests: keep a copy of the reduced meta-gradient
        self.global_task_num = None    # set when ranks hold unequal task shares
        self.collective = True         # False: this instance steps on its own even inside an initialised process group

    # -- state that must not travel through copy.deepcopy(maml) (train.py:87,127) --
    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_feat_cache":
                new.__dict__[k] = v          # the resident table is read-only: copies share it (train.py:87,127)
            elif k in ("_staging", "_ws", "_scratch", "_extractor", "_theta_flat", "_aux_stream", "_graphs", "_pool",
                       "_copy_stream", "_prefetched", "_step_graphs", "_cap_stream", "_picked"):
                new.__dict__[k] = {} if k in ("_scratch", "_graphs") else None
            else:
                new.__dict__[k] = deepcopy(v, memo)
        return new

    # -- helpers --
    def _features(self, feat, dev):
        key = (id(feat), _feat_fingerprint(feat))
        if self._feat_cache is None or self._feat_cache[0] != key or self._feat_cache[1].table.device != dev:
            self._feat_cache = (key, FeatureTable(feat, dev))
        return self._feat_cache[1]

    def refresh_features(self, feat=None):
        """Drop the resident feature table (and re-upload `feat` now when given).  Call after modifying feature
        arrays in place in a way the sampled fingerprint cannot see; swapped arrays, resized arrays and rewritten
        arrays (normalisation, augmentation) are noticed by `_features` on its own."""
        self._feat_cache = None
        if feat is not None:
            self._features(feat, _dev())

    def _buf(self, name, shape, dtype, dev, zero=False):
        t = self._scratch.get(name)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype or t.device != dev:
            t = torch.empty(max(n, 1), dtype=dtype, device=dev)
            self._scratch[name] = t
            self._alloc_gen += 1           # captured graphs hold the old pointer
        v = t[:n].view(shape)
        if zero:
            v.zero_()
        return v

    def _c_set(self, ps, base_ptr, dev, tag):
        cs = _lib.PackedSet()
        cs.n_nodes, cs.n_edges, cs.n_tiles, cs.n_tasks = ps.N, ps.E, ps.n_tiles, ps.T
        cs.n_subgraphs, cs.centres_per_subgraph = ps.S, ps.cps
        for k in packing._SEGS:
            setattr(cs, k, base_ptr + 4 * ps.off[k])
        cs.norm = self._buf(tag + "norm", (ps.N,), torch.float32, dev).data_ptr()
        if self.spec.aggregation != _lib.AGG_GCN:      # separate destination-side scales (mean / sum aggregation)
            cs.norm_dst = self._buf(tag + "norm_dst", (ps.N,), torch.float32, dev).data_ptr()
        cs.class_pos = self._buf(tag + "cpos", (ps.S,), torch.int32, dev).data_ptr()
        cs.class_occ = self._buf(tag + "cocc", (ps.S,), torch.int32, dev).data_ptr()
        cs.n_classes = self._buf(tag + "ncls", (ps.T,), torch.int32, dev).data_ptr()
        for l in range(ps.n_layers):
            cs.n_act[l], cs.n_act_tiles[l] = ps.act[l]["n"], ps.act[l]["n_tiles"]
            for k in ("act_rows", "act_task_ptr", "act_tile_row0", "act_tile_nrows", "act_tile_task"):
                getattr(cs, k)[l] = base_ptr + 4 * ps.off["%s%d" % (k, l)]
            cs.row_pos[l] = self._buf("%srow_pos%d" % (tag, l), (ps.N,), torch.int32, dev).data_ptr()
        return cs

    # -- host batch -> device: packing + ONE pinned H2D copy, optionally one step ahead on a worker thread --
    def _pack_threads(self):
        """Host threads of one CSR packing call.  Synchronous uploads use the cores of the box shared between the ranks
        on it (8 at most).  Under `prefetch` the parallelism comes from the packer workers themselves -- one thread
        each: with every core busy packing, the step's own host side (launches, graph update, the driver's
        servicing of the running step) is starved and the DEVICE time of a step grows (measured on C2: 4.4 ms with
        5 threads per worker against 3.6 ms with one)."""
        import os
        if self._pool is not None:
            return 1
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        return max(1, min(8, (os.cpu_count() or 1) // local_world))

    def _slot(self, dev):
        """Next staging slot of a ring of six (pinned buffer + device buffer each): the step in flight, the batch picked
        up ahead of it, up to three pending prefetches and a spare."""
        if self._staging is None or self._staging[0].device != dev:
            self._staging = [packing.Staging(dev) for _ in range(6)]
            self._slot_i = 0
        self._slot_i = (self._slot_i + 1) % len(self._staging)
        slot = self._staging[self._slot_i]
        for e in [e for e in (self._prefetched or []) if e[3] is slot]:
            e[2].result()                      # an abandoned prefetch still owns this slot: retire it
            self._prefetched.remove(e)
        return slot

    def _pack_upload(self, batch, ft, slot, dev, own_buffer, copy_stream):
        """Pack `batch` into the slot's pinned buffer and start its host->device copy (worker or caller thread)."""
        torch.cuda.set_device(dev)
        x_spt, y_spt, x_qry, y_qry = batch[:4]
        db = _DeviceBatch()
        db.resident = bool(own_buffer)
        db.T = len(x_spt)
        db.max_classes = packing.validate_labels(y_spt, y_qry, self.k_spt, _lib.lib())
        db.ft = ft
        t0 = time.perf_counter()
        n_layers = len(self.spec.conv)
        lib = _lib.lib()
        if not self.device_finish:
            # everything packed on the host, ONE pinned copy
            db.ps_s, db.ps_q, n = packing.pack_meta_batch(slot, batch, ft.graph_row_off, n_layers, lib,
                                                          n_threads=self._pack_threads())
            db.pack_ms = 1e3 * (time.perf_counter() - t0)
            if own_buffer:
                db.ints = torch.empty(n, dtype=torch.int32, device=dev)
                db.ints.copy_(slot.host[:n], non_blocking=True)
                torch.cuda.current_stream().synchronize()     # the slot is reused by the next upload
                db.ready = None
            else:
                slot.upload(n, copy_stream)
                db.ints, db.ready = slot.dev, slot.copied
            db.pending = None
            db.h2d_bytes = n * 4
            return db
        # host: only what the host alone knows (CSR by destination, centres, labels, feature rows); the CSR by source,
        # tile tables and active rows are derived from it on the device, behind the copy, on the copy's stream
        db.ps_s, db.ps_q, n_host, n = packing.pack_meta_batch_slim(slot, batch, ft.graph_row_off, n_layers, lib,
                                                                   n_threads=self._pack_threads())
        db.pack_ms = 1e3 * (time.perf_counter() - t0)
        st = copy_stream if copy_stream is not None else torch.cuda.current_stream()
        sets = (db.ps_s, db.ps_q)
        c0, c1 = db.ps_s.off["counts"], db.ps_q.off["counts"] + 2 + 2 * n_layers
        counts = slot.counts_buffer(c1 - c0)
        with torch.cuda.stream(st):
            if own_buffer:
                db.ints = torch.empty(n, dtype=torch.int32, device=dev)
                db.ints[:n_host].copy_(slot.host[:n_host], non_blocking=True)
            else:
                slot.upload(n_host, st)
                db.ints = slot.dev
            packing.finish_on_device(lib, db.ints, sets, n_layers, slot.finish_workspace, st.cuda_stream)
            counts[:c1 - c0].copy_(db.ints[c0:c1], non_blocking=True)
            done = slot.copied if not own_buffer else torch.cuda.Event()
            done.record(st)                    # copy + device passes + counts: `ready` for the consumer's stream
        # the realised counts are read when the batch is picked up (upload_batch): a prefetching worker goes straight
        # on to the next batch instead of waiting for the device here
        db.pending = (done, sets, n_layers, counts, c0)
        db.ready = None if own_buffer else done
        db.h2d_bytes = n_host * 4
        return db

    @staticmethod
    def _finalize(db):
        """Wait for a batch's copy + device passes and take over the realised counts (tiles, active rows)."""
        pend = getattr(db, "pending", None)
        if pend is not None:
            done, sets, n_layers, counts, c0 = pend
            done.synchronize()
            packing.apply_counts(sets, n_layers, counts.numpy(), c0)
            db.pending = None
        return db

    def prefetch(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        """Start packing and uploading a meta-batch NOW, on a worker thread and a copy stream, while the current
        step runs; the next `forward(...)` / `finetunning_batch(...)` given the same lists picks it up.  Same
        arguments as `forward`.  (The reference's DataLoader hides its batch preparation behind the step in worker
        processes, train.py:96; this hides the part that is specific to this build.)"""
        from concurrent.futures import ThreadPoolExecutor
        dev = _dev()
        batch = (x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry)
        ft = self._features(feat, dev)
        if ft.f0 != self.spec.conv[0][0]:
            raise RuntimeError("feature width %d does not match the first GraphConv (%d)" % (ft.f0, self.spec.conv[0][0]))
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=self.pack_workers, initializer=_keep_packers_off_the_step_cores)
            self._copy_stream = torch.cuda.Stream(device=dev)
        if self._prefetched is None:
            self._prefetched = []
        slot = self._slot(dev)
        fut = self._pool.submit(self._pack_upload, batch, ft, slot, dev, False, self._copy_stream)
        # the caller's pattern is prefetch(batch i+3) followed by forward(batch i): four entries can be pending (the
        # lookahead keeps the packer threads and the copy of a batch off the step's critical path even when they
        # take longer than the step itself)
        while len(self._prefetched) >= 4:
            self._prefetched.pop(0)[2].result()            # an abandoned prefetch: let the worker finish with its slot
        self._prefetched.append((x_spt, feat, fut, slot))

    def upload_batch(self, batch, feat, own_buffer=False):
        """Pack one collated meta-batch (host, integer only) and copy it to the device with ONE
        async transfer from pinned memory.  With own_buffer=True the device copy gets its own
        allocation (so several batches can stay resident, e.g. for device-resident benchmarking);
        otherwise a slot of the staging ring is used.  A batch that `prefetch` already started is picked up."""
        dev = _dev()
        if self._prefetched is None:
            self._prefetched = []
        pend = self._prefetched
        if self._picked is not None and not own_buffer:
            x0, f0, db = self._picked
            self._picked = None
            if x0 is batch[0] and f0 is feat:
                self.pickup_wait_ms = (0.0, 0.0)
                self.host_pack_ms = db.pack_ms
                return db
        hit = next((e for e in pend if e[0] is batch[0] and e[1] is feat), None) if not own_buffer else None
        t0 = time.perf_counter()
        if hit is not None:
            pend.remove(hit)
            db = hit[2].result()
        else:
            ft = self._features(feat, dev)
            if ft.f0 != self.spec.conv[0][0]:
                raise RuntimeError("feature width %d does not match the first GraphConv (%d)"
                                   % (ft.f0, self.spec.conv[0][0]))
            db = self._pack_upload(batch, ft, self._slot(dev), dev, own_buffer, None)
        t1 = time.perf_counter()
        self._finalize(db)
        # where a pick-up waited: for the packer thread (ms), then for the copy + device passes of the batch (ms)
        self.pickup_wait_ms = (1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1))
        self.host_pack_ms = db.pack_ms
        return db

    def build_batch_on_device(self, graphs, req_spt, req_qry, feat, h, sample_nodes=1000, seed=222):
        """A meta-batch assembled entirely in HBM (device_batch.py): `graphs` are the dataset's ParentGraphs (their
        CSR is uploaded once and cached), `req_*` the CentreRequests of the support / query set.  The host ships
        only centre ids and labels; extraction, batching, the transposed CSR and the active-row lists are device
        work.  The result is interchangeable with `upload_batch`'s."""
        from . import device_batch
        from .subgraphs import DeviceExtractor
        dev = _dev()
        key = (id(graphs), len(graphs))
        if getattr(self, "_extractor", None) is None or self._extractor[0] != key:
            self._extractor = (key, DeviceExtractor(graphs, dev), graphs)
        db = _DeviceBatch()
        db.resident, db.ready, db.pack_ms = False, None, 0.0
        db.T = int(req_spt.sub_off.shape[0] - 1)
        T = db.T
        split = lambda r: [r.labels[r.sub_off[t]:r.sub_off[t + 1]] for t in range(T)]      # noqa: E731
        db.max_classes = packing.validate_labels(split(req_spt), split(req_qry), self.k_spt)
        db.ft = self._features(feat, dev)
        if db.ft.f0 != self.spec.conv[0][0]:
            raise RuntimeError("feature width %d does not match the first GraphConv (%d)"
                               % (db.ft.f0, self.spec.conv[0][0]))
        db.ps_s, db.ps_q, db.ints = device_batch.build(self._extractor[1], req_spt, req_qry, h, sample_nodes,
                                                       len(self.spec.conv), seed)
        db.h2d_bytes = int(8 * (req_spt.graph_idx.shape[0] + req_qry.graph_idx.shape[0]) *
                           (3 if req_spt.centre_b is not None else 2))
        return db

    def forward_device(self, graphs, req_spt, req_qry, feat, h, sample_nodes=1000, seed=222):
        """`forward` for a meta-batch given as centre requests: subgraph extraction included, on the device."""
        K = self.update_step
        db = self.build_batch_on_device(graphs, req_spt, req_qry, feat, h, sample_nodes, seed)
        host = self.step_device(db).cpu()
        self.last.update({"loss_q": float(host[K + 1]), "skipped": bool(host[K + 2] != 0),
                          "d2h_bytes": int(host.numel() * 4)})
        return host[:K + 1].numpy().astype(np.float32)

    def finetunning_batch_device(self, graphs, req_spt, req_qry, feat, h, sample_nodes=1000, seed=222):
        """`finetunning_batch` for episodes given as centre requests (extraction on the device)."""
        K = self.update_step_test
        if req_spt.sub_off.shape[0] <= 1:
            return np.zeros((0, K + 1), dtype=np.float32)
        db = self.build_batch_on_device(graphs, req_spt, req_qry, feat, h, sample_nodes, seed)
        theta = self._flat_theta(self.net.parameters(), _dev())
        acc_q, _, _ = self._enqueue(db, K, False, theta)
        host = acc_q.cpu()
        self.last["d2h_bytes"] = int(host.numel() * 4)
        return host.numpy().astype(np.float32)

    def _enqueue(self, db, steps, train, flat_theta, meta_grad=None, stats=None):
        """Enqueue the whole inner loop for an uploaded meta-batch.  Returns device tensors
        (acc_q [T,K+1], loss_q [T,K+1], meta_grad [P] or None).  `meta_grad` / `stats`: where the summed
        meta-gradient [P] and the step's scalars [K+2] (sum of last query losses, accuracy sums) go."""
        a, info, (acc_q, loss_q, meta_grad) = self._build_args(db, steps, train, flat_theta, meta_grad, stats)
        if getattr(db, "ready", None) is not None:
            torch.cuda.current_stream().wait_event(db.ready)          # the batch's H2D copy (maybe on the copy stream)
        L = _lib.lib()
        _lib.check(L.gmeta_maml_step(C.byref(a), _stream()), "maml_step")
        info["gpu_launches"] = L.gmeta_last_launch_count()
        self.last = info
        return acc_q, loss_q, meta_grad

    def _build_args(self, db, steps, train, flat_theta, meta_grad=None, stats=None):
        """The argument block of gmeta_maml_step for an uploaded meta-batch (allocates / grows the buffers it points
        to; launches nothing).  Returns (args, diagnostics, (acc_q, loss_q, meta_grad))."""
        L = _lib.lib()
        dev = _dev()
        T, ps_s, ps_q, ft = db.T, db.ps_s, db.ps_q, db.ft
        if train and steps < 2:
            raise RuntimeError("element 0 of tensors does not require grad and does not have a grad_fn "
                               "(update_step must be >= 2, as in the reference: meta.py:137-141,161)")
        base = db.ints.data_ptr()
        a = _lib.StepArgs()
        a.model = self.spec.c_model()
        a.spt = self._c_set(ps_s, base, dev, "s_")
        a.qry = self._c_set(ps_q, base, dev, "q_")
        a.feat_table, a.ld_feat = ft.table.data_ptr(), ft.ld
        a.feat_rowmax = ft.rowmax.data_ptr()
        a.theta = flat_theta.data_ptr()
        a.update_step, a.n_support, a.max_classes = steps, self.k_spt, db.max_classes
        a.spt_max_rows_per_task, a.qry_max_rows_per_task = ps_s.max_rows_per_task, ps_q.max_rows_per_task
        a.update_lr = self.update_lr
        a.grad_scale = 1.0 / (self._global_task_num(T) if train else 1)
        a.compute_meta_grad = 1 if train else 0
        a.dense_backward = 1 if self.dense_backward else 0
        a.pruned_forward = 1 if (self.pruned_forward and not self.dense_backward) else 0
        a.impl = self.impl
        P = self.spec.n_params_padded
        if train and meta_grad is None:
            meta_grad = self._buf("meta_grad", (P,), torch.float32, dev)
        loss_q = self._buf("loss_q", (T, steps + 1), torch.float32, dev)
        acc_q = self._buf("acc_q", (T, steps + 1), torch.float32, dev)
        loss_s = self._buf("loss_s", (T, steps), torch.float32, dev)
        a.meta_grad = _ptr(meta_grad)
        a.loss_q, a.acc_q, a.loss_s = loss_q.data_ptr(), acc_q.data_ptr(), loss_s.data_ptr()
        logits0 = None
        if getattr(self, "keep_logits_spt0", False):
            logits0 = self._buf("logits0", (ps_s.S, self.spec.n_out), torch.float32, dev)
        a.logits_spt0 = _ptr(logits0)
        a.step_stats = _ptr(stats)
        if self.two_streams and a.pruned_forward:      # full-formulation launches fill the chip on their own
            if self._aux_stream is None or self._aux_stream.device != dev:
                self._aux_stream = torch.cuda.Stream(device=dev)
            a.aux_stream = self._aux_stream.cuda_stream
        nbytes = L.gmeta_maml_step_workspace_bytes(C.byref(a))
        if nbytes < 0:
            raise _lib.GMetaError("invalid meta-step arguments")
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=dev)
            self._alloc_gen += 1
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        info = {"h2d_bytes": db.h2d_bytes, "gpu_launches": 0,
                "n_nodes": (ps_s.N, ps_q.N), "n_edges": (ps_s.E, ps_q.E), "workspace_bytes": int(nbytes),
                "logits_spt0": logits0, "loss_s": loss_s}
        return a, info, (acc_q, loss_q, meta_grad)

    def _run(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat,
             steps, train, flat_theta):
        db = self.upload_batch((x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry), feat)
        return self._enqueue(db, steps, train, flat_theta)

    def _global_task_num(self, local_tasks):
        """task_num of the whole meta-batch across ranks (meta.py:161).  Ranks hold equal shares
        (dist.shard_tasks), so no collective is needed unless `global_task_num` says otherwise."""
        if self.global_task_num is not None:
            return int(self.global_task_num)
        from . import dist
        return local_tasks * (dist.world_size() if self.collective else 1)

    def _flat_theta(self, params, dev):
        """The flat parameter buffer [W1 | b1 | ... | Wlin | blin] of the C ABI.  The net's parameters are kept as
        VIEWS of it, so neither the inner loop (reads theta) nor the fused Adam (updates it in place) needs a copy
        in or out; the aliasing is re-established whenever something replaced a parameter's storage (`.to()`,
        deepcopy, `p.data = ...`)."""
        params = list(params)
        flat = self._theta_flat
        ok = flat is not None and flat.device == dev and all(
            p.device == dev and p.data_ptr() == flat.data_ptr() + 4 * off and p.is_contiguous()
            for p, off in zip(params, self.spec.offsets))
        if not ok:
            flat = torch.zeros(self.spec.n_params_padded, dtype=torch.float32, device=dev)
            self.spec.flatten([p.to(dev) for p in params], flat)
            for p, v in zip(params, self.spec.unflatten(flat)):
                p.data = v
            self._theta_flat = flat
            self._alloc_gen += 1
        return flat

    def _prepare_step_graph(self, db):
        """Capture the training step of a host batch into the next updatable graph (no device work): called for the
        NEXT batch while the current step runs, so that a step costs the host one graph launch.  Returns
        (handle, diagnostics, validity key) or None when capture is not possible right now."""
        dev = _dev()
        L = _lib.lib()
        K, P = self.update_step, self.spec.n_params_padded
        theta = self._flat_theta(self.net.parameters(), dev)
        T_global = self._global_task_num(db.T)
        red = self._buf("reduce", (P + K + 2,), torch.float32, dev)
        a, info, _ = self._build_args(db, K, True, theta, meta_grad=red[:P], stats=red[P:])
        if self._step_graphs is None:
            self._step_graphs = []
            for _ in range(3):
                h = C.c_void_p()
                _lib.check(L.gmeta_step_graph_create(C.byref(h)), "step_graph_create")
                self._step_graphs.append(h)
            self._cap_stream = torch.cuda.Stream(device=dev)
        self._sg_i = (self._sg_i + 1) % len(self._step_graphs)
        h = self._step_graphs[self._sg_i]
        rc = L.gmeta_step_graph_prepare(h, C.byref(a), self._cap_stream.cuda_stream)
        if rc != 0:
            return None
        info["gpu_launches"] = L.gmeta_last_launch_count()
        db.graph = (h, info, (self._alloc_gen, theta.data_ptr(), T_global))
        return db.graph

    def _prepare_ahead(self):
        """While the step just launched runs: if the next prefetched batch has left the packer thread, pick it up and
        prepare its step graph now."""
        if not (self.use_graphs and self.graph_host_batches and self.prepare_ahead) or self._picked is not None \
                or not self._prefetched:
            return
        x0, feat, fut, slot = self._prefetched[0]
        t0 = time.perf_counter()
        try:
            # the device is busy with the step for a few milliseconds: waiting for the packer here costs nothing that
            # the next pick-up would not have to wait for anyway
            db = fut.result(timeout=0.003)
        except Exception:
            return
        self._prefetched.pop(0)
        t1 = time.perf_counter()
        db = self._finalize(db)
        t2 = time.perf_counter()
        self._prepare_step_graph(db)
        self._picked = (x0, feat, db)
        # diagnostics: waited for the packer thread, for the batch's copy, host time of capture + graph update (ms)
        self.ahead_ms = (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (time.perf_counter() - t2))

    # -- public API (meta.py:236-244) --
    def step_device(self, db):
        """One training meta-step on an uploaded batch, everything on the device: inner loop, the
        single all-reduce, fused Adam.  Returns a device tensor [K+3] = mean accuracies (K+1),
        mean query loss, skipped flag -- no host synchronisation.

        The inner loop of a batch that owns its device buffer (upload_batch(own_buffer=True): device-resident
        training loops) is captured into a CUDA graph on its second use and replayed afterwards: ~250 launches on
        two streams become one graph launch."""
        from . import dist
        dev = _dev()
        K = self.update_step
        P = self.spec.n_params_padded
        theta = self._flat_theta(self.net.parameters(), dev)
        T_global = self._global_task_num(db.T)
        # [meta-grad | sum_t loss_q^K | sum_t acc_q[0..K]] -- the one collective of a meta-step; the inner loop
        # writes straight into it
        red = self._buf("reduce", (P + K + 2,), torch.float32, dev)
        out = self._buf("step_out", (K + 3,), torch.float32, dev)
        key = None
        if self.use_graphs and getattr(db, "resident", False) and not getattr(self, "keep_logits_spt0", False):
            key = (db.ints.data_ptr(), db.ps_q.end, db.ft.table.data_ptr(), K, T_global, self._alloc_gen)
            if self._graphs and next(iter(self._graphs))[-1] != self._alloc_gen:
                self._graphs.clear()               # a buffer moved: every captured pointer set is stale
        entry = self._graphs.get(key) if key is not None else None
        sg = None
        if key is None and self.use_graphs and self.graph_host_batches and getattr(db, "ready", None) is not None \
                and not getattr(self, "keep_logits_spt0", False):        # host batches in a staging slot
            sg = getattr(db, "graph", None)
            if sg is None or sg[2] != (self._alloc_gen, theta.data_ptr(), T_global):
                sg = self._prepare_step_graph(db)
        if sg is not None:
            if getattr(db, "ready", None) is not None:
                torch.cuda.current_stream().wait_event(db.ready)
            _lib.check(_lib.lib().gmeta_step_graph_launch(sg[0], _stream()), "step_graph_launch")
            self.last = dict(sg[1])
        elif entry is not None and entry[0] is not None:
            entry[0].replay()
            self.last = dict(entry[1])
        else:
            if key is not None and entry is None and len(self._graphs) < 16:
                self._enqueue(db, K, True, theta, meta_grad=red[:P], stats=red[P:])     # first use: eager (allocates)
                if key[-1] == self._alloc_gen:
                    self._graphs[key] = (None, None)                                    # next use: capture
            elif key is not None and entry is not None:
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(db, K, True, theta, meta_grad=red[:P], stats=red[P:])
                self._graphs[key] = (g, dict(self.last))
                g.replay()
            else:
                self._enqueue(db, K, True, theta, meta_grad=red[:P], stats=red[P:])
        if self.collective:
            dist.allreduce_sum_(red)
        # meta.py:161-171: gate = sum loss / task_num, NaN -> skip; Adam on theta (the net's own storage)
        self.meta_optim.step(theta, red[:P], loss_sum=red[P:P + 1], loss_scale=1.0 / T_global, acc_sums=red[P + 1:],
                             step_out=out)
        self.last["gpu_launches"] = self.last.get("gpu_launches", 0) + 2          # Adam: prepare + apply
        if self.return_meta_grad:
            self.last["meta_grad"] = [g.clone() for g in self.spec.unflatten(red[:P])]
        return out

    def forward_ProtoMAML(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        K = self.update_step
        db = self.upload_batch((x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry), feat)
        out = self.step_device(db)
        self._prepare_ahead()                                             # the next batch's launches, while this step runs
        host = out.cpu()                                                  # the step's only D2H + sync
        self.last.update({"loss_q": float(host[K + 1]), "skipped": bool(host[K + 2] != 0),
                          "d2h_bytes": int(host.numel() * 4)})
        return host[:K + 1].numpy().astype(np.float32)

    def finetunning_ProtoMAML(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        dev = _dev()
        K = self.update_step_test
        theta = self._flat_theta(self.net.parameters(), dev)             # fine-tunes a copy (meta.py:181)
        one = lambda v: [v[0]]                                            # noqa: E731  (meta.py:182-191)
        acc_q, loss_q, _ = self._run(one(x_spt), one(y_spt), one(x_qry), one(y_qry), one(c_spt), one(c_qry),
                                     one(n_spt), one(n_qry), one(g_spt), one(g_qry), feat, K, False, theta)
        host = acc_q[0].cpu()
        self.last["d2h_bytes"] = int(host.numel() * 4)
        return host.numpy().astype(np.float32)                            # meta.py:232-234

    def finetunning_batch(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        """Fine-tune on E independent episodes at once: every episode starts from the current
        weights and gets its own fast-weight copy, exactly as E calls of `finetunning` would (meta.py:175-234
        deep-copies the net per call), but as ONE batched inner loop.  Returns np.float32 [E, update_step_test+1];
        `self.net` is not modified."""
        K = self.update_step_test
        if len(x_spt) == 0:
            return np.zeros((0, K + 1), dtype=np.float32)
        theta = self._flat_theta(self.net.parameters(), _dev())
        acc_q, _, _ = self._run(x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat,
                                K, False, theta)
        host = acc_q.cpu()
        self.last["d2h_bytes"] = int(host.numel() * 4)
        return host.numpy().astype(np.float32)

    def forward(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        if self.method == 'G-Meta':
            accs = self.forward_ProtoMAML(x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat)
        return accs

    def finetunning(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        if self.method == 'G-Meta':
            accs = self.finetunning_ProtoMAML(x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat)
        return accs

#!/usr/bin/env python
"""Micro-benchmark of the fused GCN layer kernel on a synthetic packed set shaped like the C2
query set (T tasks x ~38k rows, subgraph blocks of ~500 rows, ~1.8 in-edges per row).
Prints ms/launch, algorithmic GB/s and the fraction of the measured HBM peak for each impl."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmeta_b200 import _lib  # noqa: E402
from gmeta_b200.learner import tile_table  # noqa: E402


def synth(T, rows_per_task, block, deg, rng):
    sizes = np.full(T, rows_per_task)
    trp = np.concatenate([[0], np.cumsum(sizes)])
    N = int(trp[-1])
    e = int(N * deg)
    dst = np.sort(rng.integers(0, N, e))
    blk = dst // block
    src = np.minimum(blk * block + rng.integers(0, block, e), N - 1)
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=N), out=indptr[1:])
    return trp, indptr, src.astype(np.int32), N, e


def real_c2(T, rng):
    """Packed query set of T tasks of the C2 workload (skewed degrees, hubs)."""
    from gmeta_b200 import packing
    from gmeta_b200.synthetic import make_dataset
    ds = make_dataset('C2')
    mb = ds.sample_meta_batch(rng, T)
    xq, cq = mb[2], mb[5]
    ps = packing.plan_set(xq, cq, 0)
    buf = np.zeros(ps.end, dtype=np.int32)
    packing.fill_set(buf, ps, xq, mb[3], cq, mb[7], mb[9], np.array([0]))
    o = ps.off
    return (ps.node_off, buf[o["indptr"]:o["indptr"] + ps.N + 1].copy(), buf[o["indices"]:o["indices"] + ps.E].copy(),
            ps.N, ps.E)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--real", action="store_true", help="use a real C2 query set instead of the uniform synthetic")
    ap.add_argument("--tasks", type=int, default=32)
    ap.add_argument("--rows", type=int, default=38000)
    ap.add_argument("--fin", type=int, default=256)
    ap.add_argument("--fout", type=int, default=256)
    ap.add_argument("--deg", type=float, default=1.76)
    ap.add_argument("--impls", default="1,2")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ablate", default="", help="comma list of debug flag masks to time (tcgen05 impl)")
    ap.add_argument("--plan", type=int, default=1, help="1: reuse a prebuilt layer plan (impl 3); 0: rebuild per call")
    ap.add_argument("--profile-flags", type=int, default=0, help="ablation flags active during the --profile launch")
    ap.add_argument("--profile", action="store_true", help="per-role cycle counters of the tcgen05 kernel")
    a = ap.parse_args()
    L = _lib.lib()
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    if a.real:
        trp, indptr, indices, N, E = real_c2(a.tasks, rng)
    else:
        trp, indptr, indices, N, E = synth(a.tasks, a.rows, 500, a.deg, rng)
    row0, nrows, task = tile_table(trp)
    i32 = lambda x: torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32)).to(dev)  # noqa: E731
    d_indptr, d_indices, d_row0, d_nrows, d_task = i32(indptr), i32(indices), i32(row0), i32(nrows), i32(task)
    x = torch.randn(N, a.fin, device=dev)
    W = torch.randn(a.tasks, a.fin * a.fout + a.fout, device=dev) * 0.05
    out = torch.empty(N, a.fout, device=dev)
    norm = torch.empty(N, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.gmeta_degree_norm(d_indptr.data_ptr(), N, norm.data_ptr(), st))
    P = a.fin * a.fout + a.fout
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    alg = 4.0 * (N * a.fin + N * a.fout + E + N + 1 + a.tasks * (a.fin * a.fout + a.fout))
    res = {}
    outs = {}
    pb = L.gmeta_layer_plan_bytes(len(row0), a.tasks, N, E)
    plan = torch.empty(pb + 256, dtype=torch.uint8, device=dev)
    plan_ptr = (plan.data_ptr() + 255) // 256 * 256
    _lib.check(L.gmeta_layer_plan_build(d_indptr.data_ptr(), d_indices.data_ptr(), norm.data_ptr(), None, None,
                                        d_row0.data_ptr(), d_nrows.data_ptr(), d_task.data_ptr(), len(row0), a.tasks,
                                        N, E, plan_ptr, st))
    for impl in [int(v) for v in a.impls.split(",")]:
        nb = L.gmeta_gcn_layer_fwd_ex_workspace_bytes(a.tasks, P, len(row0), N, E, a.fin, a.fout, impl)
        ws = torch.empty(max(nb, 16) + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        rmax_in = torch.empty(N, device=dev)
        rmax_out = torch.empty(N, device=dev)
        _lib.check(L.gmeta_row_absmax(x.data_ptr(), a.fin, N, a.fin, rmax_in.data_ptr(), st))

        def launch():
            _lib.check(L.gmeta_gcn_layer_fwd_ex(x.data_ptr(), a.fin, None, None, d_indptr.data_ptr(), d_indices.data_ptr(),
                                                norm.data_ptr(), d_row0.data_ptr(), d_nrows.data_ptr(), d_task.data_ptr(),
                                                len(row0), a.tasks, W.data_ptr(), P, a.fout, 0,
                                                W.data_ptr() + 4 * a.fin * a.fout, P, a.fin, a.fout, 1, None,
                                                out.data_ptr(), a.fout, impl, ws_ptr, nb, N, E,
                                                rmax_in.data_ptr() if impl == 3 else None,
                                                rmax_out.data_ptr() if impl == 3 else None,
                                                plan_ptr if (impl == 3 and a.plan) else None, st))
        for _ in range(2):
            launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        if impl == 2 and a.profile:
            prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
            L.gmeta_debug_set_tc_profile(prof.data_ptr())
            launch()
            torch.cuda.synchronize()
            L.gmeta_debug_set_tc_profile(None)
            pr = prof.cpu().numpy().reshape(148, 16).astype(np.float64)
            names = ["prod0.prologue", "prod0.wait", "prod0.body", "prod7.prologue", "prod7.wait", "prod7.body",
                     "mma.wait_acc_empty", "mma.wait_a_full", "mma.wait_b_full", "mma.issue", "epi.wait_acc_full",
                     "epi.body", "epi.ldtm", "epi.store"]
            res["profile_kcycles_mean_per_cta"] = {n: round(float(pr[:, i].mean()) / 1e3, 1) for i, n in enumerate(names)}
        if impl == 3 and a.profile:
            prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
            L.gmeta_debug_set_pair_profile(prof.data_ptr())
            L.gmeta_debug_set_pair_flags(a.profile_flags)
            launch()
            L.gmeta_debug_set_pair_flags(0)
            torch.cuda.synchronize()
            L.gmeta_debug_set_pair_profile(None)
            pr = prof.cpu().numpy().reshape(148, 16).astype(np.float64)
            names = ["prod0.setup", "prod0.wait_empty", "prod0.body", "prod15.setup", "prod15.wait_empty", "prod15.body",
                     "mma.wait_w", "mma.wait_acc_empty", "mma.wait_a_full", "mma.issue", "epi.wait_acc_full", "epi.other",
                     "epi.stage", "epi.expand"]
            res["pair_profile_kcycles_mean_per_cta"] = {n: round(float(pr[:, i].mean()) / 1e3, 1) for i, n in enumerate(names)}
            res["pair_profile_leader_only"] = {n: round(float(pr[0::2, i].mean()) / 1e3, 1) for i, n in enumerate(names)}
            res["pair_profile_max_per_cta"] = {n: round(float(pr[:, i].max()) / 1e3, 1) for i, n in enumerate(names)}
            res["pair_profile_min_per_cta"] = {n: round(float(pr[:, i].min()) / 1e3, 1) for i, n in enumerate(names)}
            tot = pr[:, 10] + pr[:, 11] + pr[:, 12] + pr[:, 13]
            res["pair_profile_epi_total_kcycles"] = {"mean": round(float(tot.mean()) / 1e3, 1), "max": round(float(tot.max()) / 1e3, 1),
                                                     "min": round(float(tot.min()) / 1e3, 1)}
        if impl in (2, 3) and a.ablate:
            set_flags = L.gmeta_debug_set_tc_flags if impl == 2 else L.gmeta_debug_set_pair_flags
            for fl in [int(v) for v in a.ablate.split(",")]:
                set_flags(fl)
                launch()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.reps):
                    launch()
                e1.record()
                torch.cuda.synchronize()
                res["impl%d_ablate_%d_ms" % (impl, fl)] = round(e0.elapsed_time(e1) / a.reps, 4)
            set_flags(0)
            launch()
        if impl == 3:
            torch.cuda.synchronize()
            res["rowmax_max_abs_diff"] = float((rmax_out - out.abs().amax(1)).abs().max())
        outs[impl] = out.clone()
        res[impl] = {"ms": ms, "GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / peak}
    for k in outs:
        if k != 1 and 1 in outs:
            res["max_abs_diff_impl%d_vs_ffma" % k] = float((outs[1] - outs[k]).abs().max())
            res["max_abs_ref"] = float(outs[1].abs().max())
    print(json.dumps({"N": N, "E": E, "tiles": len(row0), "alg_GB": alg / 1e9, "res": res}))


if __name__ == "__main__":
    main()

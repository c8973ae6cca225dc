#!/usr/bin/env python
"""Runs a few device-resident meta-steps of a workload without CUDA graphs -- the command to put under
`ncu --metrics gpu__time_duration.sum` for a per-kernel launch list of the step (tools/ncu_summary.py launches)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmeta_b200.meta import Meta  # noqa: E402
from gmeta_b200.synthetic import make_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--full", action="store_true", help="the reference's formulation (pruned_forward=0)")
ap.add_argument("--one-stream", action="store_true")
a = ap.parse_args()
ds = make_dataset(a.workload)
mb = ds.sample_meta_batch(np.random.default_rng(1000), ds.task_num)
args = ds.args()
args.use_graphs = False
args.two_streams = not a.one_stream
args.pruned_forward = not a.full
torch.manual_seed(222)
m = Meta(args, ds.config()).to("cuda")
db = m.upload_batch(mb, ds.feats, own_buffer=True)
for _ in range(a.warmup):
    m.step_device(db)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = m.step_device(db)
e1.record()
torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / a.steps, "launches", m.last["gpu_launches"], "out", out.cpu().numpy()[-3:])

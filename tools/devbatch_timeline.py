"""Debug: where a device-assembled meta-batch spends its time (device_batch.build phases, each closed by a device
synchronisation, then the meta-step) on the C2 workload."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmeta_b200 import device_batch  # noqa: E402
from gmeta_b200.meta import Meta  # noqa: E402
from gmeta_b200.synthetic import make_dataset  # noqa: E402

ds = make_dataset(sys.argv[1] if len(sys.argv) > 1 else 'C2')
rng = np.random.default_rng(1000)
mb = ds.sample_meta_batch(rng, ds.task_num)
xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
req_s = device_batch.CentreRequests.from_host_batch(xs, cs, ns, gs, ys)
req_q = device_batch.CentreRequests.from_host_batch(xq, cq, nq, gq, yq)
torch.manual_seed(222)
m = Meta(ds.args(), ds.config()).to('cuda')
for _ in range(3):
    m.forward_device(ds.graphs, req_s, req_q, ds.feats, ds.h, ds.sample_nodes)
ex = m._extractor[1]
tm = {}
reps = 5
for _ in range(reps):
    device_batch.build(ex, req_s, req_q, ds.h, ds.sample_nodes, len(m.spec.conv), 222, timings=tm)
print("build phases, ms per meta-batch (synchronised):", {k: round(v / reps, 3) for k, v in tm.items()}, "sum %.3f" % (sum(tm.values()) / reps))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    db = m.build_batch_on_device(ds.graphs, req_s, req_q, ds.feats, ds.h, ds.sample_nodes)
torch.cuda.synchronize()
t1 = time.perf_counter()
for _ in range(reps):
    m.step_device(db)
torch.cuda.synchronize()
t2 = time.perf_counter()
for _ in range(reps):
    m.forward_device(ds.graphs, req_s, req_q, ds.feats, ds.h, ds.sample_nodes)
torch.cuda.synchronize()
t3 = time.perf_counter()
print("build_batch_on_device %.3f ms, step_device %.3f ms, forward_device %.3f ms" %
      (1e3 * (t1 - t0) / reps, 1e3 * (t2 - t1) / reps, 1e3 * (t3 - t2) / reps))

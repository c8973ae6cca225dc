"""Debug: does a concurrent host->device copy stream slow the (graph-replayed) meta-step down?  Device-resident C2 steps
alone, then with a background thread copying 36 MB pinned chunks on a side stream, then with a packer-like CPU load."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmeta_b200.meta import Meta  # noqa: E402
from gmeta_b200.synthetic import make_dataset  # noqa: E402

ds = make_dataset('C2')
rng = np.random.default_rng(1000)
batches = [ds.sample_meta_batch(rng, 32) for _ in range(3)]
torch.manual_seed(222)
m = Meta(ds.args(), ds.config()).to('cuda')
dbs = [m.upload_batch(b, ds.feats, own_buffer=True) for b in batches]
for i in range(60):
    m.step_device(dbs[i % 3])
torch.cuda.synchronize()


def timed(label, n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        m.step_device(dbs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    print("%-40s %.3f ms per step" % (label, e0.elapsed_time(e1) / n))


timed("alone")
stop = False
host = torch.empty(9 * 1024 * 1024, dtype=torch.int32, pin_memory=True)
dev = torch.empty_like(host, device='cuda')
side = torch.cuda.Stream()


def copier(pause):
    while not stop:
        with torch.cuda.stream(side):
            dev.copy_(host, non_blocking=True)
        side.synchronize()
        time.sleep(pause)


for pause in (0.0, 0.002):
    stop = False
    th = threading.Thread(target=copier, args=(pause,))
    th.start()
    timed("with 36 MB H2D copies (pause %.0f ms)" % (1e3 * pause))
    stop = True
    th.join()


def burner():
    a = np.random.rand(1 << 20)
    while not stop:
        a = a * 1.0000001 + 1e-9


stop = False
ths = [threading.Thread(target=burner) for _ in range(3)]
[t.start() for t in ths]
timed("with 3 numpy threads busy")
stop = True
[t.join() for t in ths]

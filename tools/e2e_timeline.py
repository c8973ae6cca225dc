"""Debug: host-side timeline of the e2e path (prefetch + forward) over a few steps."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmeta_b200.meta import Meta
from gmeta_b200.synthetic import make_dataset
if len(sys.argv) > 1:
    sys.setswitchinterval(float(sys.argv[1]))
ds = make_dataset('C2')
rng = np.random.default_rng(1000)
batches = [ds.sample_meta_batch(rng, 32) for _ in range(3)]
torch.manual_seed(222)
m = Meta(ds.args(), ds.config()).to('cuda')
for i in range(3):
    m(*batches[i % 3], ds.feats)
for mode in ("plain", "prefetch", "prefetch2"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    marks = []
    waits = []
    for i in range(12):
        a = time.perf_counter()
        if mode == "prefetch":
            m.prefetch(*batches[(i + 1) % 3], ds.feats)
        if mode == "prefetch2":
            if i == 0:
                m.prefetch(*batches[1], ds.feats)
            m.prefetch(*batches[(i + 2) % 3], ds.feats)
        b = time.perf_counter()
        db = m.upload_batch(batches[i % 3], ds.feats)
        c = time.perf_counter()
        out = m.step_device(db)
        d = time.perf_counter()
        host = out.cpu()
        e = time.perf_counter()
        marks.append((b - a, c - b, d - c, e - d, db.pack_ms))
        waits.append(m.pickup_wait_ms)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / 12
    mk = np.array(marks[2:]) * 1e3
    mk[:, 4] /= 1e3
    print(mode, "pickup waits (packer thread, device) ms: %.2f %.2f" % tuple(np.array(waits[2:]).mean(0)))
    print(mode, "ms/step %.2f" % (tot * 1e3), "prefetch_submit %.2f upload %.2f enqueue %.2f wait_result %.2f pack_ms %.2f" % tuple(mk.mean(0)))

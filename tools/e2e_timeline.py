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

batches = batches + [ds.sample_meta_batch(rng, 32) for _ in range(2)]


def forward_mode(label):
    # the public call (Meta.forward) with the two-batch lookahead: the next batch's step graph is prepared while the current
    # step runs
    import ctypes as C
    from gmeta_b200 import _lib
    NB = len(batches)
    for i in range(5):
        m.prefetch(*batches[(i + 3) % NB], ds.feats) if i else [m.prefetch(*batches[j], ds.feats) for j in (1, 2, 3)]
        m(*batches[i % NB], ds.feats)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    picked = 0
    fw, ah, gpu = [], [], []
    _orig_step = m.step_device
    _ev = []
    def _timed_step(db):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = _orig_step(db)
        e1.record()
        _ev.append((e0, e1))
        return out
    m.step_device = _timed_step
    for i in range(5, 5 + n):
        m.prefetch(*batches[(i + 3) % NB], ds.feats)
        picked += m._picked is not None
        ta = time.perf_counter()
        m(*batches[i % NB], ds.feats)
        fw.append(1e3 * (time.perf_counter() - ta))
        ah.append(getattr(m, "ahead_ms", (0, 0, 0)))
    torch.cuda.synchronize()
    print("device time of a step (events around step_device) ms %.2f" % np.mean([a.elapsed_time(b) for a, b in _ev[2:]]))
    print("forward() ms %.2f; prepare-ahead: wait packer %.2f, wait copy %.2f, capture+update %.2f" % ((np.mean(fw),) + tuple(np.array(ah).mean(0))))
    print(label, "forward + lookahead ms/step %.2f  picked-ahead %d/%d" % (1e3 * (time.perf_counter() - t0) / n, picked, n))
    if m._step_graphs:
        for h in m._step_graphs:
            u, k = C.c_int32(), C.c_int32()
            _lib.lib().gmeta_step_graph_stats(h, C.byref(u), C.byref(k))
            print("step graph: updates", u.value, "instantiations", k.value)


m.step_device_orig = m.step_device
forward_mode("graph prepared ahead")
m.step_device = m.step_device_orig
m.two_streams = False
forward_mode("graph prepared ahead, one stream")
m.two_streams = True
m.step_device = m.step_device_orig
m.prepare_ahead = False
forward_mode("graph prepared at the step")
m.step_device = m.step_device_orig
m.graph_host_batches = False
forward_mode("eager")

"""Stand-in for the HOST side of the other ranks of an 8-GPU job, for measuring one rank's end-to-end step on a 1-GPU box
under the host-memory traffic of a full node:

    for r in 1..7:  taskset -c <4 cores of rank r> python tools/host_contention.py --seconds 90 [--slim] &
    LOCAL_WORLD_SIZE=8 taskset -c 0-3 python bench.py --no-cpu-baseline --no-configs --no-device-extract

Each stand-in packs C2 meta-batches with the product's host packer (packing.pack_meta_batch, or the slim packer that
goes with Meta.device_finish) on `--threads` threads at the batch rate of a real rank (`--period-ms` per batch over all
threads), and reads the packed buffer once more per batch the way the H2D copy engine would.  No GPU is touched.
Prints the batch rate it sustained."""
import argparse
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--threads", type=int, default=3)
    ap.add_argument("--period-ms", type=float, default=3.3, help="one batch per this many ms over all threads")
    ap.add_argument("--slim", action="store_true")
    ap.add_argument("--batches", type=int, default=3)
    a = ap.parse_args()
    from gmeta_b200 import _lib, packing
    from gmeta_b200.synthetic import make_dataset
    ds = make_dataset("C2", scale=1.0)
    rng = np.random.default_rng(4000 + os.getpid() % 1000)
    batches = [ds.sample_meta_batch(rng, ds.task_num) for _ in range(a.batches)]
    goff = np.concatenate([[0], np.cumsum([f.shape[0] for f in ds.feats])])[:-1]
    n_layers = len([c for c in ds.config() if c[0] == "GraphConv"])
    lib = _lib.lib()
    done = [0] * a.threads
    t_end = time.perf_counter() + a.seconds

    def work(k):
        st = packing.Staging(torch.device("cpu"))
        sink = None
        period = a.period_ms * 1e-3 * a.threads
        nxt = time.perf_counter() + period * k / a.threads
        i = k
        while time.perf_counter() < t_end:
            if a.slim:
                n = packing.pack_meta_batch_slim(st, batches[i % a.batches], goff, n_layers, lib, 1)[2]
            else:
                n = packing.pack_meta_batch(st, batches[i % a.batches], goff, n_layers, lib, 1)[2]
            buf = st.host.numpy()[:n]
            if sink is None or sink.shape[0] < n:
                sink = np.empty(n, dtype=buf.dtype)
            np.copyto(sink[:n // 2], buf[:n // 2])     # about the traffic of one more read of the packed buffer
            done[k] += 1
            i += a.threads
            nxt += period
            d = nxt - time.perf_counter()
            if d > 0:
                time.sleep(d)
            else:
                nxt = time.perf_counter()

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(k,)) for k in range(a.threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    print("host_contention pid %d: %d batches in %.1f s = %.0f batches/s (%s)" % (
        os.getpid(), sum(done), dt, sum(done) / dt, "slim" if a.slim else "full"), flush=True)


if __name__ == "__main__":
    main()

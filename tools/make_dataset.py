#!/usr/bin/env python
"""Write one of the synthetic workloads C1..C5 (gmeta_b200/synthetic.py) as a dataset
directory in the reference's on-disk layout (features.npy, label.pkl, {train,val,test}.csv [+ *_spt /
*_qry for link prediction], graph_csr.npz), ready for `python train.py --data_dir <out>/ ...`.

    python tools/make_dataset.py C1 /tmp/c1 [--scale 0.2] [--items 200]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmeta_b200 import data_io  # noqa: E402
from gmeta_b200.synthetic import make_dataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("out")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--items", type=int, default=None, help="labelled items kept per class / graph")
    ap.add_argument("--seed", type=int, default=222)
    a = ap.parse_args()
    ds = make_dataset(a.workload, seed=a.seed, scale=a.scale)
    data_io.write_synthetic_dataset(a.out, ds, np.random.default_rng(a.seed), items_per_graph=a.items)
    flags = "--task_setup %s --n_way %d --k_spt %d --k_qry %d --hidden_dim %d --h %d --update_step %d --update_lr %g " \
            "--meta_lr %g --task_num %d" % (ds.task_setup, ds.n_way, ds.k_spt, ds.k_qry, ds.hidden_dim, ds.h,
                                            ds.update_step, ds.update_lr, ds.meta_lr, ds.task_num)
    if ds.link_pred:
        flags += " --link_pred_mode True"
    print("wrote %s: %d graph(s), %d nodes / %d edges in graph 0, feat=%d" % (
        a.out, len(ds.graphs), ds.graphs[0].n, ds.graphs[0].number_of_edges(), ds.feats[0].shape[1]))
    print("python train.py --data_dir %s/ %s" % (a.out.rstrip("/"), flags))


if __name__ == "__main__":
    main()

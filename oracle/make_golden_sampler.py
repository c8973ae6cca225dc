"""ORACLE / TEST INFRASTRUCTURE ONLY.

Golden fixtures for the episode sampler: runs the UNMODIFIED reference `Subgraphs`
(/root/reference/G-Meta/subgraph_data_processing.py:14-412, via oracle/ref_loader.py + DGL stand-in) on
seeded tiny dataset directories and records its pre-sampled task lists and the first episodes (labels,
centre parent ids, node sets, induced edges as parent-id pairs) in tests/golden/sampler_<kind>.json.
The dataset directory is regenerated from seeds by the test (tests/test_subgraphs_dataset.py), so the
fixture pins the product's sampler where the reference tree is not mounted (the GPU box).
Run in the build container:      python -m oracle.make_golden_sampler
"""
import argparse
import json
import os
import pickle
import random
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gmeta_b200 import data_io  # noqa: E402
from oracle import ref_loader  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
KINDS = ("disjoint", "shared", "link")
SEED_DS, SEED_TASKS, SEED_EP, BATCHSZ, N_EP = 5, 11, 100, 6, 2


def seed_all(s):
    np.random.seed(s)
    random.seed(s)
    torch.manual_seed(s)


def sampler_args(ds):
    return argparse.Namespace(sample_nodes=10 ** 6, link_pred_mode='True' if ds.link_pred else 'False',
                              task_setup=ds.task_setup)


def write_dataset(kind, root, dgl_module=None):
    ds = H.tiny_dataset(kind)
    data_io.write_synthetic_dataset(root, ds, np.random.default_rng(SEED_DS), dgl_module=dgl_module)
    return ds


def main():
    _, _, sdp = ref_loader.load()
    dgl = ref_loader.shim_dgl()
    for kind in KINDS:
        root = tempfile.mkdtemp()
        ds = write_dataset(kind, root, dgl)
        info = data_io.load_labels(root)
        with open(os.path.join(root, data_io.GRAPH_PKL), "rb") as f:
            graphs = pickle.load(f)
        seed_all(SEED_TASKS)
        ref = sdp.Subgraphs(root, 'train', info, n_way=ds.n_way, k_shot=ds.k_spt, k_query=ds.k_qry, batchsz=BATCHSZ,
                            args=sampler_args(ds), adjs=graphs, h=ds.h)
        out = {"support_x_batch": ref.support_x_batch, "query_x_batch": ref.query_x_batch, "episodes": []}
        for idx in range(N_EP):
            seed_all(SEED_EP + idx)
            r = ref[idx]
            ep = {"y_spt": r[1].tolist(), "y_qry": r[3].tolist(), "g_spt": list(r[8]), "g_qry": list(r[9]), "sets": []}
            for gi, ci, ni in ((0, 4, 6), (2, 5, 7)):
                g = r[gi]
                off = np.concatenate([[0], np.cumsum(g.batch_num_nodes)])
                src, dst = (t.numpy() for t in g.edges())
                subs = []
                for k in range(len(g.batch_num_nodes)):
                    nid = np.asarray(r[ni][k])
                    c = np.atleast_1d(r[ci][k].numpy())
                    m = (dst >= off[k]) & (dst < off[k + 1])
                    edges = sorted(zip(nid[src[m] - off[k]].tolist(), nid[dst[m] - off[k]].tolist()))
                    subs.append({"nodes": sorted(nid.tolist()), "centres": nid[c].tolist(), "edges": edges})
                ep["sets"].append(subs)
            out["episodes"].append(ep)
        with open(os.path.join(OUT, "sampler_%s.json" % kind), "w") as f:
            json.dump(out, f)
        print("wrote sampler_%s.json" % kind)


if __name__ == "__main__":
    main()

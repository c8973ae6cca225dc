#!/usr/bin/env python
"""One-off converter for the reference's graph_dgl.pkl (a pickle of DGL 0.4.3 graph objects,
train.py:43-44): run where `dgl` is importable; writes graph_csr.npz next to it, which
gmeta_b200.data_io.load_graphs reads without DGL.

    python tools/convert_dgl_graphs.py /path/to/dataset_dir
"""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmeta_b200 import data_io  # noqa: E402


def main():
    root = sys.argv[1]
    with open(os.path.join(root, data_io.GRAPH_PKL), "rb") as f:
        graphs = pickle.load(f)
    graphs = data_io.as_parent_graphs(graphs)
    data_io.save_graphs(root, graphs)
    print("wrote %s (%d graphs)" % (os.path.join(root, data_io.GRAPH_NPZ), len(graphs)))


if __name__ == "__main__":
    main()

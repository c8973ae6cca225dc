"""On-disk dataset formats of the reference (README.md:188-203, train.py:39-55) and a DGL-free
equivalent for the graphs.

A dataset directory holds

  features.npy     np.ndarray [N, F] (one graph) or an object array / list of per-graph [N_g, F]
                   arrays (train.py:41, :63-65)
  label.pkl        dict  item name -> int label; node items are "g_i", link items "g_i_j"
                   (subgraph_data_processing.py:349-358)
  {train,val,test}.csv      header + rows `index,name,label`  (loadCSV reads row[1], row[2];
                   subgraph_data_processing.py:118-147); link prediction additionally has
                   {mode}_spt.csv / {mode}_qry.csv (:36-40)
  graph_dgl.pkl    pickle of a list of DGL 0.4.3 graphs (train.py:43-44) -- only readable where a
                   `dgl` module is importable
  graph_csr.npz    (this build) the same graphs as int32 CSR by destination: n_graphs, and per graph
                   k: n_k, indptr_k, indices_k.  `tools/convert_dgl_graphs.py` writes it from
                   graph_dgl.pkl on a machine that has DGL; `load_graphs` prefers it.

Shared-label task mode (`--task_mode True`) reads label.pkl and the CSVs from `<root>/task<n>/`
(train.py:50-53).
"""
import csv
import os
import pickle

import numpy as np

from .subgraphs import ParentGraph

GRAPH_NPZ = "graph_csr.npz"
GRAPH_PKL = "graph_dgl.pkl"


def graph_from_dgl(g):
    """ParentGraph from anything with the DGL 0.4 surface `number_of_nodes()` + `edges()` (or
    `all_edges()`): edge u -> v is kept as stored, with multiplicity."""
    src, dst = g.all_edges() if hasattr(g, "all_edges") else g.edges()
    to_np = lambda t: t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)  # noqa: E731
    return ParentGraph.from_edges(to_np(src), to_np(dst), int(g.number_of_nodes()))


def as_parent_graphs(adjs):
    return [a if isinstance(a, ParentGraph) else graph_from_dgl(a) for a in adjs]


def save_graphs(root, graphs):
    d = {"n_graphs": np.array(len(graphs), dtype=np.int64)}
    for k, g in enumerate(graphs):
        d["n_%d" % k] = np.array(g.n, dtype=np.int64)
        d["indptr_%d" % k] = g.indptr.astype(np.int64)
        d["indices_%d" % k] = g.indices.astype(np.int32)
    np.savez_compressed(os.path.join(root, GRAPH_NPZ), **d)


def load_graphs(root):
    """List of ParentGraph for the dataset at `root` (graph_csr.npz, else graph_dgl.pkl)."""
    npz = os.path.join(root, GRAPH_NPZ)
    if os.path.isfile(npz):
        d = np.load(npz)
        return [ParentGraph(d["indptr_%d" % k], d["indices_%d" % k], int(d["n_%d" % k]))
                for k in range(int(d["n_graphs"]))]
    pkl = os.path.join(root, GRAPH_PKL)
    if not os.path.isfile(pkl):
        raise FileNotFoundError("neither %s nor %s under %s" % (GRAPH_NPZ, GRAPH_PKL, root))
    try:
        with open(pkl, "rb") as f:                                   # train.py:43-44
            graphs = pickle.load(f)
    except ModuleNotFoundError as e:
        raise RuntimeError(
            "%s is a pickle of DGL graph objects and needs the `dgl` package to be read (%s). "
            "Convert it once where DGL is installed:  python tools/convert_dgl_graphs.py %s" % (pkl, e, root))
    return as_parent_graphs(graphs)


def load_features(root):
    """features.npy as a list of float arrays, one per graph (train.py:41, :63-65)."""
    feat = np.load(os.path.join(root, "features.npy"), allow_pickle=True)
    if feat.dtype != object and feat.ndim == 2:
        return [feat]
    return [np.asarray(f) for f in feat]


def load_labels(root):
    with open(os.path.join(root, "label.pkl"), "rb") as f:
        return pickle.load(f)


def read_item_csv(path):
    """(name, label-string) per data row: column 1 and 2 of the reference CSVs; the header row is
    skipped (subgraph_data_processing.py:124-130)."""
    with open(path) as f:
        rd = csv.reader(f, delimiter=",")
        next(rd, None)
        return [(row[1], row[2]) for row in rd]


def write_item_csv(path, items):
    """items: iterable of (name, label).  Same three columns as the reference files."""
    with open(path, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["", "name", "label"])
        for k, (name, label) in enumerate(items):
            wr.writerow([k, name, label])


def write_synthetic_dataset(root, ds, rng, frac=(0.8, 0.1, 0.1), items_per_graph=None, dgl_module=None):
    """Write a SyntheticDataset (gmeta_b200.synthetic) in the reference's directory layout.

    Node tasks: Disjoint splits the CLASSES over train/val/test (node_process.py:60-80 does the same
    for arxiv), Shared splits the GRAPHS (tissue-PPI: node_process.py:82-100); link tasks split the
    graphs and write the *_spt / *_qry files (link_process.py:117-198).  With `dgl_module` given the
    graphs are also pickled as graph_dgl.pkl through that module's DGLGraph (tests use the oracle's
    shim to feed the unmodified reference loader)."""
    os.makedirs(root, exist_ok=True)
    feats = ds.feats
    if len(feats) == 1:
        np.save(os.path.join(root, "features.npy"), feats[0])
    else:
        arr = np.empty(len(feats), dtype=object)
        for k, f in enumerate(feats):
            arr[k] = f
        np.save(os.path.join(root, "features.npy"), arr, allow_pickle=True)
    save_graphs(root, ds.graphs)
    if dgl_module is not None:
        gl = []
        for g in ds.graphs:
            dst = np.repeat(np.arange(g.n, dtype=np.int64), np.diff(g.indptr))
            dg = dgl_module.DGLGraph()
            dg.add_nodes(g.n)
            dg.add_edges(g.indices.astype(np.int64), dst)
            gl.append(dg)
        with open(os.path.join(root, GRAPH_PKL), "wb") as f:
            pickle.dump(gl, f)
    label, split = {}, {"train": [], "val": [], "test": []}
    modes = ("train", "val", "test")

    def part(n):
        perm = rng.permutation(n)
        a, b = int(round(frac[0] * n)), int(round((frac[0] + frac[1]) * n))
        a = min(max(a, 1), n - 2)
        b = min(max(b, a + 1), n - 1)
        return {"train": perm[:a], "val": perm[a:b], "test": perm[b:]}

    if ds.link_pred:
        gpart = part(len(ds.graphs))
        spt = {m: [] for m in modes}
        qry = {m: [] for m in modes}
        for m in modes:
            for g in gpart[m]:
                pi, pj, pl, is_spt = ds.pairs[int(g)]
                idx = np.arange(pi.shape[0]) if items_per_graph is None else \
                    rng.choice(pi.shape[0], min(items_per_graph, pi.shape[0]), replace=False)
                for k in idx:
                    name = "%d_%d_%d" % (g, pi[k], pj[k])
                    if name in label:            # a random negative that repeats an earlier pair: first one wins
                        continue
                    label[name] = int(pl[k])
                    split[m].append((name, int(pl[k])))
                    (spt if is_spt[k] else qry)[m].append((name, int(pl[k])))
        for m in modes:
            write_item_csv(os.path.join(root, m + "_spt.csv"), spt[m])
            write_item_csv(os.path.join(root, m + "_qry.csv"), qry[m])
    elif ds.task_setup == "Disjoint":
        lab = ds.node_labels[0]
        cpart = part(ds.n_classes_total)
        for m in modes:
            for c in cpart[m]:
                nodes = np.nonzero(lab == c)[0]
                if items_per_graph is not None:
                    nodes = rng.choice(nodes, min(items_per_graph, nodes.shape[0]), replace=False)
                for v in nodes:
                    name = "0_%d" % v
                    label[name] = int(c)
                    split[m].append((name, int(c)))
    else:
        gpart = part(len(ds.graphs))
        for m in modes:
            for g in gpart[m]:
                lab = ds.node_labels[int(g)]
                nodes = np.arange(lab.shape[0]) if items_per_graph is None else \
                    rng.choice(lab.shape[0], min(items_per_graph, lab.shape[0]), replace=False)
                for v in nodes:
                    name = "%d_%d" % (g, v)
                    label[name] = int(lab[v])
                    split[m].append((name, int(lab[v])))
    for m in modes:
        write_item_csv(os.path.join(root, m + ".csv"), split[m])
    with open(os.path.join(root, "label.pkl"), "wb") as f:
        pickle.dump(label, f)
    return root

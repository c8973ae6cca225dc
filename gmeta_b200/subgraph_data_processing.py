"""Episode sampler over local subgraphs -- the `Subgraphs` dataset of the reference
(G-Meta/subgraph_data_processing.py:14-412) with the same constructor, the same CSV / label inputs,
the same pre-sampled task lists and the same 10-tuple per episode, so that `train.py` and any user
code written against the reference keep working.

What is different, deliberately:
  * graphs are `ParentGraph` CSR objects (DGL graphs are converted on the way in), a subgraph is a
    `SubgraphCSR` and a batch of them a `PackedSubgraphBatch` -- the containers the CUDA path
    consumes -- instead of DGL objects;
  * the h-hop extraction is vectorised integer work (gmeta_b200/subgraphs.py), not per-neighbour
    `.item()` loops; nodes of a subgraph are in ascending parent id (the reference's order is that of
    a python `set`, :303): logits are equivariant to that order;
  * `h` outside {1,2,3} raises NameError like the reference (:300-311 leave the name unbound).

The task lists draw from numpy's and python's global RNGs in the reference's order
(create_batch_* :150-292), so seeding both reproduces the reference's episodes item for item
(tests/test_subgraphs_dataset.py checks that against the unmodified reference class).
"""
import os
import random

import numpy as np
import torch
from torch.utils.data import Dataset

from .data_io import as_parent_graphs, read_item_csv
from .packed import PackedSubgraphBatch
from .subgraphs import extract_subgraph, extract_subgraph_link_pred


def _group(items):
    """The three dictionaries loadCSV returns (:118-147): label -> names, graph -> names,
    graph -> label -> names; insertion order is file order."""
    by_label, by_graph, by_graph_label = {}, {}, {}
    for name, label in items:
        g = int(name.split('_')[0])
        by_graph.setdefault(g, []).append(name)
        by_graph_label.setdefault(g, {}).setdefault(label, []).append(name)
        by_label.setdefault(label, []).append(name)
    return by_label, by_graph, by_graph_label


class Subgraphs(Dataset):
    def __init__(self, root, mode, subgraph2label, n_way, k_shot, k_query, batchsz, args, adjs, h):
        self.batchsz = batchsz                  # number of pre-sampled tasks, not of subgraphs
        self.n_way, self.k_shot, self.k_query = n_way, k_shot, k_query
        self.setsz = n_way * k_shot
        self.querysz = n_way * k_query
        self.h = h
        self.sample_nodes = args.sample_nodes
        print('shuffle DB :%s, b:%d, %d-way, %d-shot, %d-query, %d-hops' % (mode, batchsz, n_way, k_shot, k_query, h))
        self.subgraph2label = subgraph2label
        self.link_pred_mode = args.link_pred_mode == 'True'
        self.task_setup = args.task_setup
        self.G = as_parent_graphs(list(adjs))
        self.subgraphs = {}                     # memo: item name -> (SubgraphCSR, centre(s), parent ids)

        if self.link_pred_mode:
            _, graphs_spt, gl_spt = self.loadCSV(os.path.join(root, mode + '_spt.csv'))
            _, graphs_qry, gl_qry = self.loadCSV(os.path.join(root, mode + '_qry.csv'))
        dictLabels, dictGraphs, dictGraphsLabels = self.loadCSV(os.path.join(root, mode + '.csv'))

        if self.task_setup == 'Disjoint':
            self.data = [names for names in dictLabels.values()]          # one name list per class
            self.cls_num = len(self.data)
            self.create_batch_disjoint(self.batchsz)
        elif self.task_setup == 'Shared':
            if self.link_pred_mode:
                self.data_graph_spt, self.data_label_spt = self._per_graph(graphs_spt, gl_spt)
                self.graph_num_spt = len(self.data_graph_spt)
                self.cls_num_spt = len(self.data_label_spt[0])
                self.data_graph_qry, self.data_label_qry = self._per_graph(graphs_qry, gl_qry)
                self.graph_num_qry = len(self.data_graph_qry)
                self.cls_num_qry = len(self.data_label_qry[0])
                self.create_batch_LinkPred(self.batchsz)
            else:
                self.data_graph, self.data_label = self._per_graph(dictGraphs, dictGraphsLabels)
                self.graph_num = len(self.data_graph)
                self.cls_num = len(self.data_label[0])
                self.create_batch_shared(self.batchsz)

    @staticmethod
    def _per_graph(by_graph, by_graph_label):
        """[names of graph k] and [[names of label l of graph k]] in file order (:61-111)."""
        graphs = [names for names in by_graph.values()]
        labels = [[names for names in by_graph_label[g].values()] for g in by_graph.keys()]
        return graphs, labels

    def loadCSV(self, csvf):
        return _group(read_item_csv(csvf))

    # ---- pre-sampled task lists; RNG draw order as in the reference ----
    def _split(self, names, n_pick):
        """k_shot support + the rest query out of `n_pick` distinct items of one class (:167-176)."""
        idx = np.random.choice(len(names), n_pick, False)
        np.random.shuffle(idx)
        names = np.array(names)
        return names[idx[:self.k_shot]].tolist(), names[idx[self.k_shot:]].tolist()

    def create_batch_disjoint(self, batchsz):
        """n_way random classes per task, regardless of the graph they live in (:150-183)."""
        self.support_x_batch, self.query_x_batch = [], []
        for _ in range(batchsz):
            selected_cls = np.random.choice(self.cls_num, self.n_way, False)
            np.random.shuffle(selected_cls)
            support_x, query_x = [], []
            for cls in selected_cls:
                s, q = self._split(self.data[cls], self.k_shot + self.k_query)
                support_x.append(s)
                query_x.append(q)
            random.shuffle(support_x)           # python's RNG (:177-178); unseeded in train.py
            random.shuffle(query_x)
            self.support_x_batch.append(support_x)
            self.query_x_batch.append(query_x)

    def create_batch_shared(self, batchsz):
        """One random graph per task and ALL of its classes; n_way is not used (:185-244)."""
        k_shot, k_query = self.k_shot, self.k_query
        self.support_x_batch, self.query_x_batch = [], []
        for _ in range(batchsz):
            data = self.data_label[np.random.choice(self.graph_num, 1, False)[0]]
            selected_cls = np.arange(len(data))
            np.random.shuffle(selected_cls)
            support_x, query_x = [], []
            for cls in selected_cls:
                if len(data[cls]) >= k_shot + k_query:
                    s, q = self._split(data[cls], k_shot + k_query)
                    support_x.append(s)
                    query_x.append(q)
                elif len(data[cls]) >= k_shot:
                    # too few items of this class (:218-238, "not used in practice"): everything that is
                    # left after the support picks, topped up with random items of random classes
                    idx = np.arange(len(data[cls]))
                    np.random.shuffle(idx)
                    names = np.array(data[cls])
                    support_x.append(names[idx[:k_shot]].tolist())
                    q = names[idx[k_shot:]].tolist()
                    for _ in range(k_shot + k_query - len(data[cls]) + 1):
                        sub_cls = np.random.choice(selected_cls, 1)[0]
                        q = q + [np.array(data[sub_cls])[np.random.choice(len(data[sub_cls]), 1)[0]]]
                    query_x.append(q)
                else:
                    print('each class in a graph must have larger than k_shot entities in the current model')
            random.shuffle(support_x)
            random.shuffle(query_x)
            self.support_x_batch.append(support_x)
            self.query_x_batch.append(query_x)

    def create_batch_LinkPred(self, batchsz):
        """One random graph per task; support links from *_spt.csv, query links from *_qry.csv (:246-292)."""
        self.support_x_batch, self.query_x_batch = [], []
        for _ in range(batchsz):
            g = np.random.choice(self.graph_num_spt, 1, False)[0]
            data_spt, data_qry = self.data_label_spt[g], self.data_label_qry[g]
            cls_spt = np.arange(len(data_spt))
            np.random.shuffle(cls_spt)
            cls_qry = np.arange(len(data_qry))
            np.random.shuffle(cls_qry)
            support_x, query_x = [], []
            for cls in cls_spt:
                idx = np.random.choice(len(data_spt[cls]), self.k_shot, False)
                np.random.shuffle(idx)
                support_x.append(np.array(data_spt[cls])[idx].tolist())
            for cls in cls_qry:
                idx = np.random.choice(len(data_qry[cls]), self.k_query, False)
                np.random.shuffle(idx)
                query_x.append(np.array(data_qry[cls])[idx].tolist())
            random.shuffle(support_x)
            random.shuffle(query_x)
            self.support_x_batch.append(support_x)
            self.query_x_batch.append(query_x)

    # ---- subgraphs on the fly, memoised per item (:295-346) ----
    def generate_subgraph(self, G, i, item):
        hit = self.subgraphs.get(item)
        if hit is None:
            sub = extract_subgraph(G, i, self.h, self.sample_nodes)
            hit = (sub, sub.centre, list(sub.parent_nid))
            self.subgraphs[item] = hit
        return hit

    def generate_subgraph_link_pred(self, G, i, j, item):
        hit = self.subgraphs.get(item)
        if hit is None:
            sub = extract_subgraph_link_pred(G, i, j, self.sample_nodes)
            hit = (sub, list(sub.centre), list(sub.parent_nid))
            self.subgraphs[item] = hit
        return hit

    def _episode_half(self, task):
        items = [item for sublist in task for item in sublist]
        parts = [item.split('_') for item in items]
        if self.link_pred_mode:
            info = [self.generate_subgraph_link_pred(self.G[int(p[0])], int(p[1]), int(p[2]), item)
                    for p, item in zip(parts, items)]
        else:
            info = [self.generate_subgraph(self.G[int(p[0])], int(p[1]), item) for p, item in zip(parts, items)]
        graph_idx = [int(p[0]) for p in parts]
        x = [s for s, _, _ in info]
        y = np.array([self.subgraph2label[item] for item in items]).astype(np.int32)
        centre = np.array([c for _, c, _ in info]).astype(np.int32)
        node_idx = [n for _, _, n in info]
        return x, y, centre, node_idx, graph_idx

    def __getitem__(self, index):
        """One task: (spt graphs, spt labels, qry graphs, qry labels, spt centres, qry centres,
        spt parent ids, qry parent ids, spt graph ids, qry graph ids)  (:348-408)."""
        support_x, support_y, support_center, support_node_idx, support_graph_idx = \
            self._episode_half(self.support_x_batch[index])
        query_x, query_y, query_center, query_node_idx, query_graph_idx = \
            self._episode_half(self.query_x_batch[index])
        if self.task_setup == 'Disjoint':
            unique = np.unique(support_y)
            random.shuffle(unique)              # labels -> 0..n_way-1 in a random order (:390-397)
            support_y_relative = np.zeros(self.setsz)
            query_y_relative = np.zeros(self.querysz)
            for idx, l in enumerate(unique):
                support_y_relative[support_y == l] = idx
                query_y_relative[query_y == l] = idx
            support_y, query_y = support_y_relative, query_y_relative
        return (PackedSubgraphBatch.batch(support_x, support_node_idx), torch.LongTensor(support_y),
                PackedSubgraphBatch.batch(query_x, query_node_idx), torch.LongTensor(query_y),
                torch.LongTensor(support_center), torch.LongTensor(query_center),
                support_node_idx, query_node_idx, support_graph_idx, query_graph_idx)

    def __len__(self):
        return self.batchsz

    def centre_requests(self, indices):
        """The tasks `indices` as centre requests for device-side extraction (device_batch.CentreRequests for the
        support and the query set): graph id, centre node id(s) and label of every item, nothing else -- the
        subgraphs are then extracted, batched and consumed in HBM (Meta.forward_device).  Labels follow
        __getitem__ (Disjoint: relabelled 0..n_way-1 in a random order per task, :390-397)."""
        from .device_batch import CentreRequests

        def half(task, relabel):
            items = [item for sublist in task for item in sublist]
            parts = [item.split('_') for item in items]
            y = np.array([self.subgraph2label[item] for item in items]).astype(np.int64)
            if relabel is not None:
                y = np.array([relabel[int(v)] for v in y], dtype=np.int64)
            cb = [int(p[2]) for p in parts] if self.link_pred_mode else None
            return [int(p[0]) for p in parts], [int(p[1]) for p in parts], cb, y

        spt, qry = [], []
        for index in indices:
            relabel = None
            if self.task_setup == 'Disjoint':
                unique = np.unique([self.subgraph2label[item] for sub in self.support_x_batch[index] for item in sub])
                random.shuffle(unique)
                relabel = {int(l): k for k, l in enumerate(unique)}
            spt.append(half(self.support_x_batch[index], relabel))
            qry.append(half(self.query_x_batch[index], relabel))
        return CentreRequests.from_tasks(spt), CentreRequests.from_tasks(qry)


def collate(samples):
    """List of episodes -> ten lists (one entry per task), as train.py:26-29."""
    return tuple(map(list, zip(*samples)))

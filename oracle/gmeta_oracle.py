"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (torch fp32, autograd) of the G-Meta inner-loop hot path, written
from SURVEY.md section 3.2 / Appendix A and the reference files it cites.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module, and only as the checker / the CPU arm.

Pinning status: the reference holds no golden vectors or tests for this path
(SURVEY.md section 4, 8c).  This restatement is pinned instead against OUTPUTS OF THE
REFERENCE ITSELF RUN IN THE BUILD CONTAINER: tests/test_oracle_vs_reference.py
asserts bit-for-bit equality with the unmodified G-Meta/learner.py + G-Meta/meta.py
(imported through oracle/ref_loader.py with the DGL stand-in), and
oracle/make_golden.py commits those reference outputs as tests/golden/*.npz so the
pin travels to the GPU box.  The third-party arithmetic that is NOT under
/root/reference (dgl==0.4.3post2 SpMM/in_degrees/batch, requirements.txt:2) is
restated from its documented semantics -- that part of parity is unpinned.

Function -> reference map
  gcn_layer            learner.py:25-56   (GraphConv.forward)
  classifier_forward   learner.py:134-194 (Classifier.forward, GraphConv/Linear/LinkPred branches)
  init_params          learner.py:70-97   (creation order + initialisers)
  euclidean_dist       meta.py:14-26
  proto_loss_spt       meta.py:28-54
  proto_loss_qry       meta.py:56-79
  OracleMeta.forward   meta.py:101-173    (forward_ProtoMAML)
  OracleMeta.finetunning meta.py:175-234  (finetunning_ProtoMAML)
"""
import copy

import numpy as np
import torch
import torch.nn.functional as F


class OGraph(object):
    """Batched directed multigraph: COO edges src[e] -> dst[e] over n nodes."""

    def __init__(self, src, dst, n, batch_num_nodes=None):
        self.src = torch.as_tensor(np.asarray(src), dtype=torch.int64).reshape(-1)
        self.dst = torch.as_tensor(np.asarray(dst), dtype=torch.int64).reshape(-1)
        self.n = int(n)
        self.batch_num_nodes = list(batch_num_nodes) if batch_num_nodes is not None else [self.n]
        self._csr = None

    @staticmethod
    def from_csr(indptr, indices, batch_num_nodes=None):
        """indptr/indices = CSR by destination (row v lists the in-neighbours of v)."""
        indptr = np.asarray(indptr, dtype=np.int64)
        indices = np.asarray(indices, dtype=np.int64)
        n = indptr.shape[0] - 1
        dst = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
        return OGraph(indices, dst, n, batch_num_nodes)

    @staticmethod
    def batch(graphs):
        srcs, dsts, nums, off = [], [], [], 0
        for g in graphs:
            srcs.append(g.src + off)
            dsts.append(g.dst + off)
            nums.extend(g.batch_num_nodes)
            off += g.n
        return OGraph(torch.cat(srcs), torch.cat(dsts), off, nums)

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.n)

    def aggregate_sum(self, h, fast=False):
        """update_all(copy_src, sum): out[v] = sum_{(u->v)} h[u]  (learner.py:38-39,44-45)."""
        if fast:
            # same sum via a CSR SpMM (MKL): used for the timed CPU arm only, so
            # the baseline is not handicapped by a scalar index_add.
            if self._csr is None:
                order = torch.argsort(self.dst, stable=True)
                counts = torch.bincount(self.dst, minlength=self.n)
                crow = torch.zeros(self.n + 1, dtype=torch.int64)
                crow[1:] = torch.cumsum(counts, 0)
                self._csr = torch.sparse_csr_tensor(
                    crow, self.src[order], torch.ones(self.src.numel(), dtype=torch.float32),
                    size=(self.n, self.n))
            return torch.sparse.mm(self._csr, h)
        out = torch.zeros((self.n,) + tuple(h.shape[1:]), dtype=h.dtype)
        return out.index_add(0, self.dst, h[self.src])


def gcn_layer(g, feat, weight, bias, in_feats, out_feats, relu=True, fast=False, aggregation="gcn"):
    """ReLU(norm * A (norm * feat) W + b), norm = clamp(in_deg,1)^-1/2 (learner.py:25-56).

    `aggregation` other than "gcn" is NOT in the reference (its GraphConv has the symmetric normalisation only): "mean"
    = ReLU(D^-1 A feat W + b), "sum" = ReLU(A feat W + b), restated from their textbook definitions -- PARITY UNPINNED
    for these two modes (nothing under /root/reference computes them)."""
    if aggregation != "gcn":
        deg = g.in_degrees().float().clamp(min=1)
        rst = torch.matmul(g.aggregate_sum(feat, fast), weight)
        if aggregation == "mean":
            rst = rst * (1.0 / deg).reshape(-1, 1)
        elif aggregation != "sum":
            raise ValueError(aggregation)
        rst = rst + bias
        return F.relu(rst) if relu else rst
    norm = torch.pow(g.in_degrees().float().clamp(min=1), -0.5)       # :29
    norm = norm.reshape(norm.shape + (1,) * (feat.dim() - 1))          # :30-31
    feat = feat * norm                                                 # :32
    if in_feats > out_feats:                                           # :34-40 matmul first
        rst = g.aggregate_sum(torch.matmul(feat, weight), fast)
    else:                                                              # :41-47 aggregate first
        rst = torch.matmul(g.aggregate_sum(feat, fast), weight)
    rst = rst * norm                                                   # :49
    rst = rst + bias                                                   # :51
    if relu:
        rst = F.relu(rst)                                              # :53-54
    return rst


def is_link_pred(config):
    return config[-1][0] == 'LinkPred'                                 # learner.py:78-79


def init_params(config):
    """Parameter list in the reference's creation order with its initialisers
    (learner.py:81-97).  Consumes the global torch RNG exactly like Classifier.__init__."""
    lp = is_link_pred(config)
    out = []
    for name, param in config:
        if name == 'Linear':
            w = torch.ones(param[1], param[0] * (2 if lp else 1))     # :84-87
            torch.nn.init.kaiming_normal_(w)                          # :88
            out += [w, torch.zeros(param[1])]
        elif name == 'GraphConv':
            w = torch.Tensor(param[0], param[1])                      # :93
            torch.nn.init.xavier_uniform_(w)                          # :94
            out += [w, torch.zeros(param[1])]
    return [p.requires_grad_(True) for p in out]


def classifier_forward(config, vars, g, to_fetch, features, fast=False, aggregation="gcn"):
    """Classifier.forward (learner.py:134-194): h GraphConv layers over the whole batched
    graph, centre-row gather (pair concat in LinkPred mode), linear head."""
    lp = is_link_pred(config)
    n_conv = sum(1 for name, _ in config if name == 'GraphConv')
    h = features.float()                                               # :144
    idx = 0
    seen = 0
    for name, param in config:
        if name == 'GraphConv':
            h = gcn_layer(g, h, vars[idx], vars[idx + 1], param[0], param[1], True, fast, aggregation)
            idx += 2
            seen += 1
            if seen == n_conv:                                         # :159-170
                offset = torch.cumsum(torch.LongTensor([0] + list(g.batch_num_nodes)), dim=0)[:-1]
                if lp:
                    h = torch.cat((h[to_fetch[:, 0] + offset], h[to_fetch[:, 1] + offset]), 1)
                else:
                    h = h[to_fetch + offset]
        elif name == 'Linear':
            h = F.linear(h, vars[idx], vars[idx + 1])                  # :172-175
            idx += 2
    return h


def euclidean_dist(x, y):
    """Squared Euclidean distances [N,M] (meta.py:14-26)."""
    n, m, d = x.size(0), y.size(0), x.size(1)
    if d != y.size(1):
        raise Exception
    return torch.pow(x.unsqueeze(1).expand(n, m, d) - y.unsqueeze(0).expand(n, m, d), 2).sum(2)


def _proto_nll(dists, n_classes, n_query):
    log_p_y = F.log_softmax(-dists, dim=1).view(n_classes, n_query, -1)
    target = torch.arange(0, n_classes).view(n_classes, 1, 1).expand(n_classes, n_query, 1).long()
    loss = -log_p_y.gather(2, target).squeeze().view(-1).mean()
    _, y_hat = log_p_y.max(2)
    acc = y_hat.eq(target.squeeze()).float().mean()
    return loss, acc


def proto_loss_spt(logits, y_t, n_support):
    """Prototype loss on the support set (meta.py:28-54).  Returns (loss, acc, prototypes)."""
    classes = torch.unique(y_t)
    n_classes = len(classes)
    idxs = [y_t.eq(c).nonzero()[:n_support].squeeze(1) for c in classes]
    prototypes = torch.stack([logits[i].mean(0) for i in idxs])
    query_idxs = torch.stack([y_t.eq(c).nonzero()[:n_support] for c in classes]).view(-1)
    dists = euclidean_dist(logits[query_idxs], prototypes)
    loss, acc = _proto_nll(dists, n_classes, n_support)
    return loss, acc, prototypes


def proto_loss_qry(logits, y_t, prototypes):
    """Prototype loss of query logits against given prototypes (meta.py:56-79)."""
    classes = torch.unique(y_t)
    n_classes = len(classes)
    n_query = int(logits.shape[0] / n_classes)
    query_idxs = torch.stack([y_t.eq(c).nonzero() for c in classes]).view(-1)
    dists = euclidean_dist(logits[query_idxs], prototypes)
    return _proto_nll(dists, n_classes, n_query)


def gather_features(feat, graph_idx, node_ids):
    """np.vstack(feat[g][ids]) -> float tensor (meta.py:119-120)."""
    return torch.Tensor(np.vstack([feat[graph_idx[j]][np.array(x)] for j, x in enumerate(node_ids)]))


class OracleMeta(object):
    """First-order ProtoMAML step of Meta (meta.py:83-234) on CPU tensors.

    x_spt/x_qry are lists of OGraph (one batched graph per task); everything else
    has the reference's types.  `fast=True` swaps index_add for an MKL CSR SpMM
    (timed CPU arm only)."""

    def __init__(self, args, config, params=None, fast=False):
        self.update_lr = args.update_lr
        self.meta_lr = args.meta_lr
        self.k_spt = args.k_spt
        self.update_step = args.update_step
        self.update_step_test = args.update_step_test
        self.config = config
        self.fast = fast
        self.aggregation = str(getattr(args, "aggregation", "gcn"))      # not in the reference; see gcn_layer
        self.vars = params if params is not None else init_params(config)
        self.meta_optim = torch.optim.Adam(self.vars, lr=self.meta_lr)   # meta.py:97
        self.last_loss_q = None
        self.last_losses_q = None

    def _net(self, g, c, feat, vars):
        return classifier_forward(self.config, vars, g, c, feat, self.fast, getattr(self, "aggregation", "gcn"))

    def _inner(self, vars0, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, feat_spt, feat_qry, steps,
               losses_q, corrects):
        logits = self._net(x_spt, c_spt, feat_spt, vars0)                              # meta.py:122
        loss, _, protos = proto_loss_spt(logits, y_spt, self.k_spt)                    # :123
        grad = torch.autograd.grad(loss, vars0)                                        # :125
        fast_w = [p - self.update_lr * g for g, p in zip(grad, vars0)]                 # :126
        with torch.no_grad():                                                          # :129-134
            lq, aq = proto_loss_qry(self._net(x_qry, c_qry, feat_qry, vars0), y_qry, protos)
            losses_q[0] = losses_q[0] + lq
            corrects[0] = corrects[0] + aq
        with torch.no_grad():                                                          # :137-141
            lq, aq = proto_loss_qry(self._net(x_qry, c_qry, feat_qry, fast_w), y_qry, protos)
            losses_q[1] = losses_q[1] + lq
            corrects[1] = corrects[1] + aq
        for k in range(1, steps):                                                      # :143-157
            logits = self._net(x_spt, c_spt, feat_spt, fast_w)
            loss, _, protos = proto_loss_spt(logits, y_spt, self.k_spt)
            grad = torch.autograd.grad(loss, fast_w, retain_graph=True)
            fast_w = [p - self.update_lr * g for g, p in zip(grad, fast_w)]
            lq, aq = proto_loss_qry(self._net(x_qry, c_qry, feat_qry, fast_w), y_qry, protos)
            losses_q[k + 1] = losses_q[k + 1] + lq
            corrects[k + 1] = corrects[k + 1] + aq

    def forward(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        task_num = len(x_spt)
        K = self.update_step
        losses_q = [0 for _ in range(K + 1)]
        corrects = [0 for _ in range(K + 1)]
        for i in range(task_num):                                                      # :118
            feat_spt = gather_features(feat, g_spt[i], n_spt[i])                       # :119
            feat_qry = gather_features(feat, g_qry[i], n_qry[i])                       # :120
            self._inner(self.vars, x_spt[i], y_spt[i], x_qry[i], y_qry[i], c_spt[i], c_qry[i],
                        feat_spt, feat_qry, K, losses_q, corrects)
        loss_q = losses_q[-1] / task_num                                               # :161
        self.last_loss_q = float(loss_q.detach())
        self.last_losses_q = [float(l.detach()) for l in losses_q]
        if not torch.isnan(loss_q):                                                    # :163-169
            self.meta_optim.zero_grad()
            loss_q.backward()
            self.last_grads = [p.grad.detach().clone() for p in self.vars]
            self.meta_optim.step()
        return np.array([float(c) for c in corrects], dtype=np.float32) / task_num     # :171-173

    def finetunning(self, x_spt, y_spt, x_qry, y_qry, c_spt, c_qry, n_spt, n_qry, g_spt, g_qry, feat):
        K = self.update_step_test
        corrects = [0 for _ in range(K + 1)]
        losses_q = [0 for _ in range(K + 1)]
        vars0 = [p.detach().clone().requires_grad_(True) for p in self.vars]           # deepcopy(net) :181
        feat_spt = gather_features(feat, g_spt[0], n_spt[0])
        feat_qry = gather_features(feat, g_qry[0], n_qry[0])
        self._inner(vars0, x_spt[0], y_spt[0], x_qry[0], y_qry[0], c_spt[0], c_qry[0],
                    feat_spt, feat_qry, K, losses_q, corrects)
        self.last_losses_q = [float(l.detach()) for l in losses_q]
        return np.array([float(c) for c in corrects], dtype=np.float32)                # :232-234

    def clone(self):
        return copy.deepcopy(self)

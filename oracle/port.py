"""TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's pick-and-choose path.

This file is the ORACLE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it, and only as the checker or the timed CPU baseline. The
product (pc-gnn_b200/) never imports it.

It restates, step for step and with the same per-target algorithmic structure (Python loop over
targets, one sort per target per relation, Python set unions, dense [B,U] mask times feature
matrix), what /root/reference/src/layers.py, graphsage.py and utils.py compute, with the two
implementation-defined orders of the reference fixed the canonical way:
neighbour lists are id-ascending and sorts are stable, i.e. order = (|Δscore| fp32, id).

Pinned against the live reference by tests/golden/make_golden.py (fixtures under tests/golden/)
and, when /root/reference is present, by tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["choose_train", "choose_test", "intra_forward", "PortPCGNN", "PortGCN", "PortSAGE",
           "pick_step_port", "pick_step_replay", "xavier"]


# --------------------------------------------------------------------------- choose step
def _rank(center0: torch.Tensor, cand_scores0: torch.Tensor):
    """|center - cand| (fp32) and its stable ascending order (layers.py:657-658 / 685-687)."""
    diff = torch.abs(center0.repeat(cand_scores0.shape[0]) - cand_scores0)
    vals, order = torch.sort(diff, dim=0, descending=False, stable=True)
    return diff, vals, order


def choose_train(center_scores, center_labels, neigh_scores, neighs_list, minor_scores,
                 minor_list, sample_list, sample_rate):
    """Train-mode choose step, restating layers.py:633-697.

    center_scores [B,2]; neigh_scores: list of [d,2]; neighs_list: list of id lists (ascending);
    minor_scores [P,2] paired with minor_list; sample_list: kept count per target.
    Returns (list of sets, list of distance lists)."""
    picked, picked_diff = [], []
    for i in range(len(center_scores)):
        c0 = center_scores[i][0]
        ids = list(neighs_list[i])
        k = sample_list[i]
        diff, vals, order = _rank(c0, neigh_scores[i][:, 0])
        if len(ids) > k + 1:                                   # layers.py:662-666
            keep = [ids[j] for j in order.tolist()[:k]]
            keep_d = vals.tolist()[:k]
        else:                                                   # layers.py:667-672
            keep = ids
            keep_d = diff.tolist()
        if center_labels[i] == 1:                               # layers.py:675-691
            n_over = int(k * sample_rate)
            _, pvals, porder = _rank(c0, minor_scores[:, 0])
            keep = keep + [minor_list[j] for j in porder.tolist()[:n_over]]
            keep_d = keep_d + pvals.tolist()[:n_over]
        picked.append(set(keep))                                # layers.py:694
        picked_diff.append(keep_d)
    return picked, picked_diff


def choose_test(center_scores, neigh_scores, neighs_list, sample_list):
    """Eval-mode choose step, restating layers.py:700-738 (no oversampling)."""
    picked, picked_diff = [], []
    for i in range(len(center_scores)):
        ids = list(neighs_list[i])
        k = sample_list[i]
        diff, vals, order = _rank(center_scores[i][0], neigh_scores[i][:, 0])
        if len(ids) > k + 1:
            picked.append(set(ids[j] for j in order.tolist()[:k]))
            picked_diff.append(vals.tolist()[:k])
        else:
            picked.append(set(ids))
            picked_diff.append(diff.tolist())
    return picked, picked_diff


# --------------------------------------------------------------------------- aggregation
def _dense_mask_agg(feat_table: torch.Tensor, picked, norm: str):
    """Mean (or 1/sqrt) over each set via the dense-mask matmul of layers.py:593-624 /
    graphsage.py:80-95,212-231."""
    uniq = sorted(set.union(*picked))
    col = {n: j for j, n in enumerate(uniq)}
    mask = torch.zeros(len(picked), len(uniq))
    rows = [i for i, s in enumerate(picked) for _ in s]
    cols = [col[n] for s in picked for n in s]
    mask[rows, cols] = 1
    cnt = mask.sum(1, keepdim=True)
    mask = mask.div(cnt.sqrt() if norm == "rsqrt" else cnt)
    return mask.mm(feat_table[torch.tensor(uniq, dtype=torch.long)])


def intra_forward(feat_table, weight, nodes, labels, neighs_list, center_scores, neigh_scores,
                  pos_scores, train_pos, sample_list, rho, train_flag):
    """IntraAgg.forward restated (layers.py:562-630): choose -> mean over the union set ->
    concat self -> relu(cat @ W_r). Returns (to_feats [B,E], picked sets, diffs)."""
    if train_flag:
        picked, diffs = choose_train(center_scores, labels, neigh_scores, neighs_list, pos_scores,
                                     train_pos, sample_list, rho)
    else:
        picked, diffs = choose_test(center_scores, neigh_scores, neighs_list, sample_list)
    agg = _dense_mask_agg(feat_table, picked, "mean")
    self_feats = feat_table[torch.tensor(nodes, dtype=torch.long)]
    to_feats = F.relu(torch.cat((self_feats, agg), dim=1).mm(weight))
    return to_feats, picked, diffs


def xavier(rng, rows, cols):
    """Xavier-uniform init as numpy (the reference uses init.xavier_uniform_, layers.py:197,560)."""
    a = math.sqrt(6.0 / (rows + cols))
    return rng.uniform(-a, a, size=(rows, cols)).astype(np.float32)


class PortPCGNN:
    """PCALayer(InterAggR(IntraAgg x R)) restated (model.py:34-62, layers.py:207-291).

    graph: object with .n_rel and .row(r, v) -> ascending int array (pcgnn_b200.graph.RelGraph).
    params: dict with 'intra' (list of [2F,E]), 'inter' [(F+RE),E], 'clf_w' [2,F], 'clf_b' [2],
    'head' [2,E] as numpy float32. train_pos is used in the order given (pass it id-sorted for the
    canonical tie rule)."""

    def __init__(self, feat, graph, train_pos, params, rho=0.5, alpha=2.0, thresholds=None):
        self.feat = torch.from_numpy(np.ascontiguousarray(feat, dtype=np.float32))
        self.graph = graph
        self.R = graph.n_rel
        self.train_pos = [int(p) for p in train_pos]
        self.rho = rho
        self.alpha = alpha
        self.thresholds = thresholds or [0.5] * self.R
        t = lambda a: torch.from_numpy(np.array(a, dtype=np.float32)).requires_grad_(True)
        self.intra = [t(w) for w in params["intra"]]
        self.inter = t(params["inter"])
        self.clf_w = t(params["clf_w"])
        self.clf_b = t(params["clf_b"])
        self.head = t(params["head"])
        self.score_table = None     # optional injected [N,2] table (shared score bits)
        self.last = {}

    def parameters(self):
        return [self.head, self.inter, *self.intra, self.clf_w, self.clf_b]

    def named_grads(self):
        names = ["weight", "inter1.weight"] + [f"inter1.intra_agg{r + 1}.weight" for r in range(self.R)] \
            + ["inter1.label_clf.weight", "inter1.label_clf.bias"]
        return {n: p.grad.detach().numpy().copy() for n, p in zip(names, self.parameters())
                if p.grad is not None}

    def inter_forward(self, nodes, labels, train_flag=True, shared_table=True):
        nodes = [int(v) for v in nodes]
        g = self.graph
        neighs = [[g.row(r, v).tolist() for v in nodes] for r in range(self.R)]      # :216-219
        uniq = sorted(set(nodes).union(*[set(x) for rel in neighs for x in rel]))     # :226-227
        pos_of = {n: j for j, n in enumerate(uniq)}                                   # :240
        if shared_table:
            table = (self.score_table if self.score_table is not None
                     else F.linear(self.feat, self.clf_w, self.clf_b))
            self.last["score_table"] = table
            batch_scores = table[torch.tensor(uniq, dtype=torch.long)]
            pos_scores = table[torch.tensor(self.train_pos, dtype=torch.long)]
        else:                                                                         # :231-237
            batch_scores = F.linear(self.feat[torch.tensor(uniq, dtype=torch.long)], self.clf_w, self.clf_b)
            pos_scores = F.linear(self.feat[torch.tensor(self.train_pos, dtype=torch.long)],
                                  self.clf_w, self.clf_b)
        center = batch_scores[[pos_of[v] for v in nodes], :]                          # :243
        feats, sel, diffs = [], [], []
        for r in range(self.R):
            r_scores = [batch_scores[[pos_of[n] for n in ids], :].view(-1, 2) for ids in neighs[r]]  # :251
            k_list = [math.ceil(len(ids) * self.thresholds[r]) for ids in neighs[r]]                  # :260
            f_r, s_r, d_r = intra_forward(self.feat, self.intra[r], nodes, labels, neighs[r], center,
                                          r_scores, pos_scores, self.train_pos, k_list, self.rho,
                                          train_flag)
            feats.append(f_r)
            sel.append([sorted(s) for s in s_r])
            diffs.append(d_r)
        self_feats = self.feat[torch.tensor(nodes, dtype=torch.long)]                 # :273-277
        cat = torch.cat([self_feats] + feats, dim=1)                                  # :284
        combined = F.relu(cat.mm(self.inter).t())                                     # :289
        self.last.update(sel=sel, diffs=diffs, to_feats=feats)
        return combined, center

    def forward(self, nodes, labels, train_flag=True, **kw):
        emb, center = self.inter_forward(nodes, labels, train_flag, **kw)             # model.py:36
        return self.head.mm(emb).t(), center, emb                                     # model.py:38-39

    def loss(self, nodes, labels, train_flag=True, **kw):
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.long)
        logits, center, emb = self.forward(nodes, lab, train_flag, **kw)
        self.last.update(logits=logits, center=center, combined=emb)
        return F.cross_entropy(logits, lab) + self.alpha * F.cross_entropy(center, lab)  # model.py:54-61

    def step_loss_backward(self, nodes, labels):
        for p in self.parameters():
            p.grad = None
        loss = self.loss(nodes, labels, True)
        loss.backward()
        return float(loss.detach())


class _PortHomo:
    def __init__(self, feat, graph, enc_w, head_w):
        self.feat = torch.from_numpy(np.ascontiguousarray(feat, dtype=np.float32))
        self.graph = graph
        self.enc_w = torch.from_numpy(np.array(enc_w, dtype=np.float32)).requires_grad_(True)
        self.head = torch.from_numpy(np.array(head_w, dtype=np.float32)).requires_grad_(True)
        self.last = {}

    def parameters(self):
        return [self.head, self.enc_w]

    def forward(self, nodes):
        emb = self.encode([int(v) for v in nodes])
        self.last["combined"] = emb
        return self.head.mm(emb).t()

    def loss(self, nodes, labels):
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.long)
        return F.cross_entropy(self.forward(nodes), lab)

    def named_grads(self):
        return {"weight": self.head.grad.numpy().copy(), "enc.weight": self.enc_w.grad.numpy().copy()}


class PortGCN(_PortHomo):
    """GCN(GCNEncoder(GCNAggregator)) restated (graphsage.py:200-232, 259-275, 167-178):
    neighbours ∪ {self}, sum / sqrt(n), relu(W[E,F] @ agg^T)."""

    def encode(self, nodes):
        picked = [set(self.graph.row(0, v).tolist()) | {v} for v in nodes]
        agg = _dense_mask_agg(self.feat, picked, "rsqrt")
        return F.relu(self.enc_w.mm(agg.t()))


class PortSAGE(_PortHomo):
    """GraphSage(Encoder(MeanAggregator)) restated (graphsage.py:62-96, 127-150). With
    Encoder(gcn=True) (model_handler.py:98) there is no self concat; the aggregator itself was built
    with gcn=False (model_handler.py:97) so no self union either; num_sample is never passed
    (graphsage.py:133)."""

    def __init__(self, *a, concat_self=False, **kw):
        super().__init__(*a, **kw)
        self.concat_self = concat_self

    def encode(self, nodes):
        picked = [set(self.graph.row(0, v).tolist()) for v in nodes]
        agg = _dense_mask_agg(self.feat, picked, "mean")
        if self.concat_self:
            agg = torch.cat((self.feat[torch.tensor(nodes, dtype=torch.long)], agg), dim=1)
        return F.relu(self.enc_w.mm(agg.t()))


# --------------------------------------------------------------------------- pick step
def pick_step_port(idx_train, y_train, degree_of, size, rng=random):
    """Label-balanced sampler restated (utils.py:274-278): weight = deg / label-frequency where
    lf = sum(y) for positives and len(y) for negatives; `size` draws with replacement through
    the given `random`-module-like rng (consumes exactly `size` rng.random() doubles, like
    random.choices)."""
    y = np.asarray(y_train)
    deg = np.array([degree_of(int(v)) for v in idx_train])
    lf = (y.sum() - len(y)) * y + len(y)
    w = deg / lf
    return rng.choices(idx_train, weights=w, k=size)


def pick_step_replay(idx_train, weights, uniforms):
    """What random.choices does with its uniforms (CPython Lib/random.py `choices`):
    cum = accumulate(weights); total = cum[-1] + 0.0; index = bisect_right(cum, u * total, 0, n-1)."""
    import bisect

    cum = np.cumsum(np.asarray(weights, dtype=np.float64)).tolist()
    total = cum[-1] + 0.0
    hi = len(cum) - 1
    return [idx_train[bisect.bisect_right(cum, u * total, 0, hi)] for u in uniforms]

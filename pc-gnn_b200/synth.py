"""Seeded synthetic fraud graphs with the shapes BASELINE.json names.

The YelpChi / Amazon pickles the reference loads (/root/reference/src/utils.py:79-113)
are not available offline, so the bench and the parity tests run on generated
graphs of the same shape: per relation an undirected edge list
with Zipf endpoint popularity over a seeded permutation (so hubs exist), then
symmetrised + self loops exactly like ``sparse_to_adjlist`` (utils.py:233-239),
Bernoulli labels, a stratified train split and ``train_pos`` from it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .graph import RelGraph, csr_from_edges

__all__ = ["SynthSpec", "SynthData", "make_graph", "SPECS"]


@dataclass
class SynthSpec:
    name: str
    n_nodes: int
    feat_dim: int
    rel_edges: tuple          # undirected edge count per relation (before dedup)
    pos_rate: float
    zipf: float = 0.6
    n_unlabeled: int = 0      # leading nodes excluded from the split (amazon: model_handler.py:39-41)
    normalize: bool = False   # row-normalise features (amazon only: model_handler.py:59-60)
    train_ratio: float = 0.4


SPECS = {
    # C1: Amazon-shaped, 3 relations U-P-U / U-S-U / U-V-U
    "amazon": SynthSpec("amazon", 11944, 25, (175608, 3566479, 1036737), 0.095,
                        n_unlabeled=3305, normalize=True),
    # C2: YelpChi-shaped, R-U-R / R-T-R / R-S-R
    "yelp": SynthSpec("yelp", 45954, 32, (49315, 573616, 3402743), 0.145),
    # C3: YelpChi-shaped with the old 100-d features
    "yelp100": SynthSpec("yelp100", 45954, 100, (49315, 573616, 3402743), 0.145),
    # small shapes for tests
    "train_sig": SynthSpec("train_sig", 3000, 25, (5000, 60000, 20000), 0.12, zipf=0.7, normalize=True),
    "tiny": SynthSpec("tiny", 600, 12, (700, 5000, 2500), 0.2, zipf=0.8),
    "tiny_amz": SynthSpec("tiny_amz", 900, 25, (1500, 20000, 7000), 0.1, zipf=0.7,
                          n_unlabeled=100, normalize=True),
}


@dataclass
class SynthData:
    spec: SynthSpec
    graph: RelGraph            # R relations
    homo: RelGraph             # union graph (1 relation)
    feat: np.ndarray           # [N, F] float32
    labels: np.ndarray         # [N] int64
    idx_train: list
    y_train: np.ndarray
    idx_rest: list
    y_rest: np.ndarray
    train_pos: list
    extra: dict = field(default_factory=dict)


def _zipf_endpoints(rng, n, m, a):
    """m endpoints with P(rank k) ∝ (k+1)^-a over a random permutation of the nodes."""
    w = (np.arange(1, n + 1, dtype=np.float64)) ** (-a)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    perm = rng.permutation(n)
    return perm[np.searchsorted(cdf, rng.random(m), side="right").clip(0, n - 1)]


def row_normalize(x):
    """Dense restatement of ``normalize`` (utils.py:213-223): x / (rowsum + 0.01)."""
    rs = x.sum(1, keepdims=True).astype(np.float64) + 0.01
    return (x / rs).astype(np.float32)


def make_graph(spec: SynthSpec | str, seed: int = 72, *, dup_feature_frac: float = 0.0,
               edge_scale: float = 1.0, signal: float = 0.0, homophily: float = 0.0) -> SynthData:
    """Build one synthetic dataset. ``dup_feature_frac`` > 0 copies feature rows so
    that label scores collide and the tie rule (distance, id) is exercised (F6).
    ``signal`` / ``homophily`` > 0 make the labels learnable (for training-run comparisons): positives get their
    features shifted by ``signal`` along a fixed direction, and a fraction ``homophily`` of every relation's edges
    is rewired to an endpoint with the same label (fraud rings). With both 0 labels are independent noise."""
    if isinstance(spec, str):
        spec = SPECS[spec]
    rng = np.random.default_rng(seed)
    n = spec.n_nodes
    learnable = signal > 0 or homophily > 0
    if learnable:
        rng_l = np.random.default_rng(seed + 7919)
        labels_sig = (rng_l.random(n) < spec.pos_rate).astype(np.int64)
        labels_sig[:spec.n_unlabeled] = 0
        by_label = [np.nonzero(labels_sig == c)[0] for c in (0, 1)]
    ips, ixs = [], []
    for m in spec.rel_edges:
        m = max(1, int(m * edge_scale))
        u = _zipf_endpoints(rng, n, m, spec.zipf)
        v = _zipf_endpoints(rng, n, m, spec.zipf)
        if learnable and homophily > 0:
            flip = rng_l.random(m) < homophily
            for c in (0, 1):
                sel = flip & (labels_sig[u] == c)
                v[sel] = by_label[c][rng_l.integers(0, len(by_label[c]), int(sel.sum()))]
        ip, ix = csr_from_edges(n, u, v)
        ips.append(ip)
        ixs.append(ix)
    graph = RelGraph(n, ips, ixs)
    homo = graph.union()

    feat = rng.random((n, spec.feat_dim), dtype=np.float32)
    if dup_feature_frac > 0:
        k = int(n * dup_feature_frac)
        dst = rng.choice(n, k, replace=False)
        src = rng.choice(n, k, replace=True)
        feat[dst] = feat[src]
    if learnable and signal > 0:
        feat[labels_sig == 1] += np.float32(signal) * rng_l.random(spec.feat_dim, dtype=np.float32)
    if spec.normalize:
        feat = row_normalize(feat)

    labels = (rng.random(n) < spec.pos_rate).astype(np.int64)
    labels[:spec.n_unlabeled] = 0
    if learnable:
        labels = labels_sig
    # stratified split without sklearn: per class shuffle + cut (model_handler.py:38-48 uses
    # train_test_split(stratify=labels); only the class proportions matter for the shape).
    index = np.arange(spec.n_unlabeled, n)
    lab = labels[index]
    tr, rest = [], []
    for c in (0, 1):
        ids = index[lab == c]
        ids = ids[rng.permutation(len(ids))]
        cut = int(round(len(ids) * spec.train_ratio))
        tr.append(ids[:cut])
        rest.append(ids[cut:])
    idx_train = np.concatenate(tr)
    idx_train = idx_train[rng.permutation(len(idx_train))]
    idx_rest = np.concatenate(rest)
    idx_rest = idx_rest[rng.permutation(len(idx_rest))]
    y_train = labels[idx_train]
    # pos_neg_split (utils.py:256-271): positives in idx_train order
    train_pos = idx_train[y_train == 1].tolist()
    return SynthData(spec, graph, homo, feat, labels, idx_train.tolist(), y_train,
                     idx_rest.tolist(), labels[idx_rest], train_pos)

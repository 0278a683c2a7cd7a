"""Evaluation metrics computed where the probabilities live.

The reference's evaluation loop (/root/reference/src/utils.py:298-325) copies every batch's probabilities to the
host (`.data.cpu().numpy()`, :305) and hands Python lists to sklearn. Once a forward pass is a ~50 us graph replay
that per-batch device->host sync is the whole cost of a validation epoch, so the batches stay on the device and the
metrics are computed there (one sort + a few reductions), with ONE small device->host copy of the results.

AUC is the Mann-Whitney statistic with average ranks for tied scores, which is what sklearn.metrics.roc_auc_score
computes (trapezoidal ROC); F1 / recall / precision use sklearn's zero_division=0 convention.
"""
from __future__ import annotations

import torch

__all__ = ["binary_metrics"]


def binary_metrics(prob1: torch.Tensor, pred: torch.Tensor, labels: torch.Tensor) -> dict:
    """prob1: P(class 1) per node, pred: predicted class (0/1), labels: true class (0/1); any device.
    Returns python floats: auc, f1 (positive class), f1_macro, recall, precision, recall_macro, precision_macro,
    accuracy, gmean."""
    y = labels.reshape(-1).to(torch.int64)
    p = pred.reshape(-1).to(torch.int64)
    s = prob1.reshape(-1).to(torch.float64)
    n = y.numel()
    # ---- AUC: rank-sum with average ranks over ties
    s_sorted, order = torch.sort(s)
    _, inv, counts = torch.unique_consecutive(s_sorted, return_inverse=True, return_counts=True)
    cum = counts.cumsum(0).to(torch.float64)
    avg_rank = (2.0 * cum - counts.to(torch.float64) + 1.0) * 0.5          # mean of ranks start+1 .. cum (1-based)
    ranks = avg_rank[inv]
    y_sorted = y[order]
    n_pos = y.sum().to(torch.float64)
    n_neg = float(n) - n_pos
    auc = (ranks[y_sorted == 1].sum() - n_pos * (n_pos + 1.0) * 0.5) / (n_pos * n_neg)
    # ---- confusion counts
    tp = ((p == 1) & (y == 1)).sum().to(torch.float64)
    tn = ((p == 0) & (y == 0)).sum().to(torch.float64)
    fp = ((p == 1) & (y == 0)).sum().to(torch.float64)
    fn = ((p == 0) & (y == 1)).sum().to(torch.float64)

    def div(a, b):
        return torch.where(b > 0, a / b.clamp_min(1.0), torch.zeros_like(a))

    prec1, rec1 = div(tp, tp + fp), div(tp, tp + fn)
    prec0, rec0 = div(tn, tn + fn), div(tn, tn + fp)
    f1_1 = div(2 * prec1 * rec1, prec1 + rec1)
    f1_0 = div(2 * prec0 * rec0, prec0 + rec0)
    out = torch.stack([auc, f1_1, 0.5 * (f1_0 + f1_1), rec1, prec1, 0.5 * (rec0 + rec1), 0.5 * (prec0 + prec1),
                       (tp + tn) / float(max(n, 1)), torch.sqrt(rec1 * rec0)]).cpu().tolist()
    keys = ["auc", "f1", "f1_macro", "recall", "precision", "recall_macro", "precision_macro", "accuracy", "gmean"]
    return dict(zip(keys, out))

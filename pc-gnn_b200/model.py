"""PCALayer head — same surface as /root/reference/src/model.py:13-62 (one Pick-Choose-Aggregate
layer: class scores from the inter-relation embedding, GNN loss + lambda * label-similarity loss).

It is a caller of the hot path, not part of it; the reference's own ``src/model.py`` runs unchanged
on top of ``layers.InterAgg*`` (see ``shim.install``). This mirror exists because the reference
tree is not present on the GPU box.
"""
import torch
import torch.nn as nn
from torch.nn import init

__all__ = ["PCALayer"]


class PCALayer(nn.Module):
    def __init__(self, num_classes, inter1, lambda_1):
        super().__init__()
        self.inter1 = inter1
        self.xent = nn.CrossEntropyLoss()
        self.weight = nn.Parameter(torch.FloatTensor(num_classes, inter1.embed_dim))   # model.py:29
        init.xavier_uniform_(self.weight)
        self.lambda_1 = lambda_1
        self.epsilon = 0.1

    def forward(self, nodes, labels, train_flag=True):
        embeds1, label_scores = self.inter1(nodes, labels, train_flag)                 # model.py:36
        scores = self.weight.mm(embeds1)
        return scores.t(), label_scores

    def to_prob(self, nodes, labels, train_flag=True):
        gnn_logits, label_logits = self.forward(nodes, labels, train_flag)
        return torch.sigmoid(gnn_logits), torch.sigmoid(label_logits)                  # model.py:41-45

    def loss(self, nodes, labels, train_flag=True):
        inter = self.inter1
        if self.weight.is_cuda and self.weight.shape[0] == 2 and hasattr(inter, "engine") \
                and isinstance(self.xent, nn.CrossEntropyLoss):
            from .layers import HeadLossFn, _as_device_labels

            if train_flag and torch.is_grad_enabled() and hasattr(inter, "train_loss"):
                # dense part + head + both cross-entropies + every weight gradient in one pass (same math as below)
                loss = inter.train_loss(nodes, labels, self.weight, float(self.lambda_1))
                if loss is not None:
                    return loss
            # head + both cross-entropies fused into one kernel per direction (same math as below)
            embeds1, label_scores = inter(nodes, labels, train_flag)
            if embeds1.shape[1] > 0:
                lab = _as_device_labels(labels, embeds1.device)
                loss, _ = HeadLossFn.apply(inter.engine(), embeds1, self.weight, label_scores, lab,
                                           float(self.lambda_1))
                return loss
        gnn_scores, label_scores = self.forward(nodes, labels, train_flag)
        label_loss = self.xent(label_scores, labels.squeeze())                         # model.py:54
        gnn_loss = self.xent(gnn_scores, labels.squeeze())                             # model.py:59
        return gnn_loss + self.lambda_1 * label_loss                                   # model.py:61

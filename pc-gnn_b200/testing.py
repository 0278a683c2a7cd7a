"""Model builders shared by bench.py, __graft_entry__.smoke(), profiles/ and the tests: the product's
PCALayer(InterAggR(IntraAgg x R)) with given parameters (the construction model_handler.py:103-114 does)."""
import numpy as np

__all__ = ["build_cuda_pcgnn"]


def build_cuda_pcgnn(feat, graph, train_pos, params, rho=0.5, alpha=2.0, device="cuda", trainable_features=False):
    """PCALayer(InterAggR(IntraAgg x R)) of the product, parameters copied in."""
    import torch
    import torch.nn as nn

    from pcgnn_b200.layers import InterAgg1, InterAgg3, InterAgg5, IntraAgg
    from pcgnn_b200.model import PCALayer

    R = graph.n_rel
    F_ = feat.shape[1]
    E = params["inter"].shape[1]
    features = nn.Embedding(feat.shape[0], F_)
    features.weight = nn.Parameter(torch.from_numpy(np.ascontiguousarray(feat)).float(),
                                   requires_grad=trainable_features)
    intras = [IntraAgg(features, F_, E, train_pos, rho, cuda=True) for _ in range(R)]
    cls = {1: InterAgg1, 3: InterAgg3, 5: InterAgg5}[R]
    inter = cls(features, F_, E, train_pos, graph, intras, cuda=True)
    model = PCALayer(2, inter, alpha)
    with torch.no_grad():
        for ia, w in zip(intras, params["intra"]):
            ia.weight.copy_(torch.from_numpy(np.asarray(w)))
        inter.weight.copy_(torch.from_numpy(params["inter"]))
        inter.label_clf.weight.copy_(torch.from_numpy(params["clf_w"]))
        inter.label_clf.bias.copy_(torch.from_numpy(params["clf_b"]))
        model.weight.copy_(torch.from_numpy(params["head"]))
    return model.to(device)

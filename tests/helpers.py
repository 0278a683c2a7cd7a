"""Shared test helpers: golden fixtures, model builders for the oracle port and the CUDA product."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                if not os.path.basename(p).startswith("homo_"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def graph_of(g):
    from pcgnn_b200.graph import RelGraph

    n, R = int(g["n_nodes"]), int(g["n_rel"])
    ips, ixs = [], []
    for r in range(R):
        ip = g["indptr"][r * n:(r + 1) * n + 1]
        ips.append(ip - ip[0])
        ixs.append(g["indices"][ip[0]:ip[-1]])
    return RelGraph(n, ips, ixs)


def params_of(g):
    return dict(intra=[w for w in g["intra"]], inter=g["inter"], clf_w=g["clf_w"], clf_b=g["clf_b"], head=g["head"])


def split_sets(ptr, idx, R, B):
    return [[idx[ptr[r * B + i]:ptr[r * B + i + 1]].tolist() for i in range(B)] for r in range(R)]


def random_params(rng, F_, E, R):
    from oracle import port

    return dict(intra=[port.xavier(rng, 2 * F_, E) for _ in range(R)], inter=port.xavier(rng, F_ + R * E, E),
                clf_w=port.xavier(rng, 2, F_), clf_b=rng.uniform(-.1, .1, 2).astype(np.float32),
                head=port.xavier(rng, 2, E))


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


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

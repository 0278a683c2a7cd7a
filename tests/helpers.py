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


from pcgnn_b200.testing import build_cuda_pcgnn  # noqa: E402,F401  (shared with bench.py)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

"""Sampler and small data helpers — same names and meaning as /root/reference/src/utils.py.

  pick_step(idx_train, y_train, adj_list, size)      label-balanced sampler      (utils.py:274-278)
  pos_neg_split(nodes, labels)                       ids split by label          (utils.py:256-271)
  normalize(mx)                                      row-normalise features      (utils.py:213-223)
  sparse_to_adjlist_for_train(sp_matrix)             scipy matrix -> graph       (utils.py:244-254)
  set_seeds(seed)                                                                (utils.py:462-470)

``pick_step`` keeps the reference's random stream: it consumes exactly ``size`` doubles from
Python's global ``random`` (what ``random.choices`` does) and returns the same list the reference
returns for the same seed; the weighted search itself runs on the GPU (``pcg_pick_step``).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib
from .graph import RelGraph

__all__ = ["pick_step", "pick_step_device", "pos_neg_split", "normalize", "sparse_to_adjlist_for_train",
           "set_seeds", "pick_weights", "test", "load_data", "create_dir", "prob2pred", "conf_gmean"]


def _degrees(adj_list, idx_train):
    if isinstance(adj_list, RelGraph):
        ip = adj_list.indptr
        t = np.asarray(idx_train, dtype=np.int64)
        return (ip[t + 1] - ip[t]).astype(np.int64)
    return np.fromiter((len(adj_list[node]) for node in idx_train), dtype=np.int64, count=len(idx_train))


def pick_weights(idx_train, y_train, adj_list):
    """Sampling weight deg(v) / LF(label(v)) in float64, exactly as utils.py:275-277: the label
    frequency is sum(y) for positives and len(y) (not the negative count) for negatives."""
    y = np.asarray(y_train)
    lf_train = (y.sum() - len(y)) * y + len(y)
    return np.array(_degrees(adj_list, idx_train)) / lf_train


def pick_step_device(idx_train, y_train, adj_list, size, *, uniforms=None, seed=None, offset=0, device=None):
    """Device-resident pick step: int32 CUDA tensor of `size` sampled node ids.

    uniforms: the doubles to replay (bit-compatible with random.choices); if None a Philox4x32-10
    counter stream keyed by `seed` is drawn on the device (same distribution, different stream)."""
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    L = _lib.lib()
    w = pick_weights(idx_train, y_train, adj_list)
    cum = torch.from_numpy(np.cumsum(w.astype(np.float64))).to(device)    # sequential fp64 == itertools.accumulate
    ids = torch.from_numpy(np.asarray(idx_train, dtype=np.int32)).to(device)
    out = torch.empty(size, dtype=torch.int32, device=device)
    if uniforms is not None:
        u = torch.from_numpy(np.asarray(uniforms, dtype=np.float64)).to(device)
        rc = L.pcg_pick_step(cum.data_ptr(), cum.shape[0], u.data_ptr(), size, ids.data_ptr(), out.data_ptr(),
                             _lib.stream_ptr())
    else:
        if seed is None:
            seed = random.getrandbits(63)
        rc = L.pcg_pick_step_philox(cum.data_ptr(), cum.shape[0], int(seed), int(offset), size, ids.data_ptr(),
                                    out.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "pcg_pick_step")
    return out


def pick_step(idx_train, y_train, adj_list, size):
    """Drop-in for the reference's pick_step: same arguments, same returned list, same consumption of
    the global `random` stream (one random() per draw, as in CPython's random.choices)."""
    u = [random.random() for _ in range(size)]
    out = pick_step_device(idx_train, y_train, adj_list, size, uniforms=u)
    picked = out.cpu().tolist()
    if len(idx_train) and not isinstance(idx_train[0], int):
        lut = {int(v): v for v in idx_train}     # hand back the caller's own objects (e.g. numpy ints)
        picked = [lut[p] for p in picked]
    return picked


def pos_neg_split(nodes, labels):
    """(positive ids, negative ids) in `nodes` order (utils.py:256-271, without its O(n^2) remove)."""
    pos, neg = [], []
    for node, label in zip(nodes, labels):
        (pos if label == 1 else neg).append(node)
    return pos, neg


def normalize(mx):
    """Row-normalise: x / (rowsum + 0.01) (utils.py:213-223). Accepts scipy sparse or dense arrays."""
    import scipy.sparse as sp

    rowsum = np.array(mx.sum(1)) + 0.01
    r_inv = np.power(rowsum, -1).flatten()
    r_inv[np.isinf(r_inv)] = 0.
    return sp.diags(r_inv).dot(mx)


def sparse_to_adjlist_for_train(sp_matrix):
    """Self loops + symmetrisation like utils.py:244-254, but straight to a CSR ``RelGraph`` (one
    relation) instead of a dict of sets."""
    return RelGraph.from_scipy([sp_matrix])


def set_seeds(seed):

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)


def load_data(data_name, seed: int = 72):
    """Synthetic stand-in for the reference's dataset loader (utils.py:66-210; the YelpChi / Amazon files are
    not available offline): (homo, relation_list, feat_data, labels) with ``homo`` / ``relation_list`` as CSR
    ``RelGraph`` objects, which every consumer in this package accepts in place of dict-of-sets."""
    from .synth import SPECS, make_graph

    key = {"yelp": "yelp", "amazon": "amazon", "amazon_new": "amazon"}.get(data_name, data_name)
    if key not in SPECS:
        raise ValueError(f"no synthetic spec for dataset {data_name!r}")
    d = make_graph(key, seed=seed)
    return d.homo, d.graph, d.feat, d.labels


def test(test_nodes, labels, model, batch_size, result=None, epoch=None, epoch_best=None, flag=None,
         print_line=True):
    """Evaluation loop with the reference's signature and return value (utils.py:280-333): batched
    ``model.to_prob(nodes, labels, train_flag=False)`` then AUC / recall / macro-F1 / precision.

    The batches' probabilities stay on the device (the reference copies each batch to the host, utils.py:305) and
    the metrics are computed there (``metrics.binary_metrics``: one sort + reductions, one device->host copy)."""
    from .metrics import binary_metrics

    labels = np.asarray(labels)
    probs = []
    for start in range(0, len(test_nodes), batch_size):
        batch_nodes = test_nodes[start:start + batch_size]
        if len(batch_nodes) == 0:
            continue
        out = model.to_prob(batch_nodes, labels[start:start + batch_size], train_flag=False)
        probs.append((out[0] if isinstance(out, tuple) else out).detach())
    prob = torch.cat(probs, dim=0)
    m = binary_metrics(prob[:, 1], prob.argmax(dim=1), torch.as_tensor(labels, device=prob.device))
    f1, f1_macro, precision, recall, auc = m["f1"], m["f1_macro"], m["precision"], m["recall"], m["auc"]
    line = f"- F1: {f1:.4f}\t- Recall: {recall:.4f}\t- Precision: {precision:.4f}\t- AUC-ROC: {auc:.4f}\t- F1-macro: {f1_macro:.4f}\n"
    if result is not None:
        writer = getattr(result, "write_val_log" if flag == "val" else "write_test_log", None)
        if writer is not None:
            try:
                acc, pm, rm = m["accuracy"], m["precision_macro"], m["recall_macro"]
                if flag == "val":
                    writer(epoch, epoch_best, acc, f1, f1_macro, precision, pm, recall, rm, auc, line, print_line)
                else:
                    writer(epoch_best, acc, f1, f1_macro, precision, pm, recall, rm, auc, line, print_line)
            except TypeError:
                pass
    elif print_line:
        print(line, end="")
    return auc, recall, f1_macro, precision


# small helpers other reference modules import from src.utils (result_manager.py:8; utils.py:440-461)
def create_dir(dir_path):
    import os

    os.makedirs(dir_path, exist_ok=True)


def prob2pred(y_prob, thres=0.5):
    return (np.asarray(y_prob) >= thres).astype(np.int32)


def conf_gmean(conf):
    tn, fp, fn, tp = np.asarray(conf).ravel()
    return (tp * tn / ((tp + fn) * (tn + fp))) ** 0.5

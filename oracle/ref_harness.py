"""TEST INFRASTRUCTURE — live-reference harness (runs only where /root/reference exists).

Imports the UNMODIFIED reference (`/root/reference/src/{layers,model,graphsage,utils}.py`)
on CPU (`cuda=False`) and canonicalises the two things the reference leaves
implementation-defined, by patching *module globals of the imported modules*
(no file is edited, nothing is copied):

  1. neighbour order  — `set(adj_list[int(node)])` (layers.py:219) becomes a set whose
     iteration order is ascending id, so `r_list` (layers.py:246-248) is id-sorted;
  2. tie order        — `torch.sort(score_diff, dim=0, descending=False)` (layers.py:658,
     687, 722) is forwarded with `stable=True`, so equal distances keep list order.

Together: the kept neighbours are the first `num_sample` by (|Δscore| fp32, id), which is
the order the CUDA kernels implement. `train_pos` must be passed id-sorted by the caller so
pool position == id order (layers.py:232 vs :690). An optional shared score table makes both
sides see identical fp32 score bits (SURVEY.md §8c recipe step 4).

Only `tests/golden/make_golden.py` and the `not gpu` cross-check tests import this.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("PCGNN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "layers.py"))


class _SortedSet(set):
    """A set that iterates in ascending order (canonical neighbour order)."""

    def __iter__(self):
        return iter(sorted(set.__iter__(self)))


class _StableTorch:
    """Delegates to torch; only `sort` differs (stable=True)."""

    def __init__(self):
        self._t = torch

    def __getattr__(self, name):
        return getattr(self._t, name)

    def sort(self, x, dim=-1, descending=False, **kw):
        return self._t.sort(x, dim=dim, descending=descending, stable=True)


def load(canonical: bool = True):
    """Import the reference modules; returns a namespace (layers, model, graphsage, utils)."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[m]
    ns = types.SimpleNamespace()
    ns.layers = importlib.import_module("src.layers")
    ns.model = importlib.import_module("src.model")
    ns.graphsage = importlib.import_module("src.graphsage")
    ns.utils = importlib.import_module("src.utils")
    ns.canonical = canonical
    if canonical:
        ns.layers.set = _SortedSet
        ns.layers.torch = _StableTorch()
    # record what the choose step returns, per IntraAgg call
    ns.choose_log = []
    for fname in ("choose_step_neighs", "choose_step_test"):
        orig = getattr(ns.layers, fname)

        def wrap(*a, _orig=orig, **kw):
            neighs, diffs = _orig(*a, **kw)
            ns.choose_log.append(([sorted(int(x) for x in s) for s in neighs],
                                  [list(map(float, d)) for d in diffs]))
            return neighs, diffs

        setattr(ns.layers, fname, wrap)
    # drop the modules again so a later `import src...` (e.g. our shim tests) starts clean
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[m]
    return ns


class _TableFeatures(torch.nn.Module):
    """`features` callable (an nn.Embedding in the reference, model_handler.py:85-86) that
    tags every output with the ids it was asked for, so a label_clf hook can answer from a
    shared score table."""

    def __init__(self, feat: np.ndarray):
        super().__init__()
        self.emb = torch.nn.Embedding(feat.shape[0], feat.shape[1])
        self.emb.weight = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(feat)).float(),
                                             requires_grad=False)

    @property
    def weight(self):
        return self.emb.weight

    def forward(self, ids):
        out = self.emb(ids)
        out._pcg_ids = ids
        return out


def build_pcgnn(ns, feat, adj_lists, train_pos, embed_dim, rho, alpha=2.0, params=None,
                shared_scores=True):
    """PCALayer(InterAggR(IntraAgg x R)) from the reference classes, CPU.

    params (optional dict of numpy arrays): 'intra' list of [2F,E], 'inter' [(RE+F),E],
    'clf_w' [2,F], 'clf_b' [2], 'head' [2,E].
    With shared_scores the label_clf output is S[ids] where S = label_clf(ALL features) in one
    GEMM (same bits for every consumer; autograd still reaches label_clf).
    """
    L, M = ns.layers, ns.model
    R = len(adj_lists)
    F_ = feat.shape[1]
    features = _TableFeatures(feat)
    tp = sorted(int(x) for x in train_pos) if ns.canonical else list(train_pos)
    intras = [L.IntraAgg(features, F_, embed_dim, tp, rho, cuda=False) for _ in range(R)]
    cls = {1: L.InterAgg1, 3: L.InterAgg3, 5: L.InterAgg5}[R]
    inter = cls(features, F_, embed_dim, tp, adj_lists, intras, cuda=False)
    model = M.PCALayer(2, inter, alpha)
    if params is not None:
        with torch.no_grad():
            for ia, w in zip(intras, params["intra"]):
                ia.weight.copy_(torch.from_numpy(w))
            inter.weight.copy_(torch.from_numpy(params["inter"]))
            inter.label_clf.weight.copy_(torch.from_numpy(params["clf_w"]))
            inter.label_clf.bias.copy_(torch.from_numpy(params["clf_b"]))
            model.weight.copy_(torch.from_numpy(params["head"]))
    state = types.SimpleNamespace(table=None)
    if shared_scores:
        def hook(mod, inp, out):
            if state.table is None:
                state.table = torch.nn.functional.linear(features.weight, mod.weight, mod.bias)
            return state.table[inp[0]._pcg_ids]

        inter.label_clf.register_forward_hook(hook)
    model._pcg_state = state
    model._pcg_features = features
    return model


def run_pcgnn(ns, model, nodes, labels, train_flag=True, backward=True):
    """One forward (+ backward) of the reference model; returns numpy outputs.

    labels: numpy int64 [B]. Output keys: sel (list over relations of list over targets of
    sorted ids), diffs, combined [E,B], center [B,2], logits [B,2], loss, score_table [N,2],
    grads (dict name -> array)."""
    ns.choose_log.clear()
    model._pcg_state.table = None
    model.zero_grad(set_to_none=True)
    lab = torch.from_numpy(np.asarray(labels, dtype=np.int64))
    nodes = [int(v) for v in nodes]
    out = {}
    cap = {}
    h = model.inter1.register_forward_hook(lambda mod, inp, res: cap.update(comb=res[0], center=res[1]))
    try:
        if train_flag and backward:
            loss = model.loss(nodes, lab, train_flag)          # model.py:47-62, unmodified
            loss.backward()
            out["loss"] = float(loss.item())
            out["grads"] = {k: p.grad.detach().numpy().copy()
                            for k, p in model.named_parameters() if p.grad is not None}
            with torch.no_grad():
                logits = model.weight.mm(cap["comb"]).t()
        else:
            with torch.no_grad():
                logits, _ = model.forward(nodes, lab, train_flag)  # model.py:34-39
    finally:
        h.remove()
    comb, center = cap["comb"], cap["center"]
    out["combined"] = comb.detach().numpy().copy()
    out["center"] = center.detach().numpy().copy()
    out["logits"] = logits.detach().numpy().copy()
    out["sel"] = [c[0] for c in ns.choose_log]
    out["diffs"] = [c[1] for c in ns.choose_log]
    if model._pcg_state.table is not None:
        out["score_table"] = model._pcg_state.table.detach().numpy().copy()
    return out

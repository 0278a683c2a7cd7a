"""TEST INFRASTRUCTURE — ctypes binding of oracle/pcg_oracle.c (the C restatement)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpcg_oracle.so")
_lib = None


def build(force: bool = False):
    src = os.path.join(_HERE, "pcg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.pcgo_bound.restype = C.c_int64
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def choose(graph, score, targets, is_pos, *, thresh=None, rho=0.5, pool=None, train=True,
           entry_score=None, center_score=None, k_override=None, pool_score=None):
    """Selected id lists for every (relation, target) item, item w = r*B + i.
    Returns (sel_ptr int64 [R*B+1], sel_idx int32)."""
    L = lib()
    R, N = graph.n_rel, graph.n_nodes
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    B = len(targets)
    is_pos = np.ascontiguousarray(is_pos, dtype=np.uint8)
    thresh = np.ascontiguousarray(thresh if thresh is not None else [0.5] * R, dtype=np.float64)
    pool = np.ascontiguousarray(pool if pool is not None else [], dtype=np.int32)
    score = None if score is None else np.ascontiguousarray(score, dtype=np.float32)
    es = None if entry_score is None else np.ascontiguousarray(entry_score, dtype=np.float32)
    cs = None if center_score is None else np.ascontiguousarray(center_score, dtype=np.float32)
    ps = None if pool_score is None else np.ascontiguousarray(pool_score, dtype=np.float32)
    ko = None if k_override is None else np.ascontiguousarray(k_override, dtype=np.int32)
    if ko is None:
        cap = int(L.pcgo_bound(_p(graph.indptr, C.c_int64), C.c_int64(N), R, _p(targets, C.c_int32),
                               _p(is_pos, C.c_uint8), B, _p(thresh, C.c_double), C.c_double(rho),
                               len(pool), int(train)))
    else:
        deg = np.array([graph.indptr[r * N + t + 1] - graph.indptr[r * N + t] for r in range(R) for t in targets])
        cap = int(deg.sum() + len(pool) * B * R)
    sel_ptr = np.zeros(R * B + 1, dtype=np.int64)
    sel_idx = np.zeros(max(cap, 1), dtype=np.int32)
    rc = L.pcgo_choose(_p(graph.indptr, C.c_int64), _p(graph.indices, C.c_int32), _p(score, C.c_float),
                       _p(es, C.c_float), _p(cs, C.c_float), C.c_int64(N), R, _p(targets, C.c_int32),
                       _p(is_pos, C.c_uint8), B, _p(thresh, C.c_double), _p(ko, C.c_int32), C.c_double(rho),
                       _p(pool, C.c_int32), _p(ps, C.c_float), len(pool), int(train),
                       _p(sel_ptr, C.c_int64), _p(sel_idx, C.c_int32), C.c_int64(len(sel_idx)))
    if rc != 0:
        raise RuntimeError("pcgo_choose: capacity too small")
    return sel_ptr, sel_idx[:sel_ptr[-1]]


def aggregate(feat, sel_ptr, sel_idx, norm="mean"):
    """out[w] = sum over the id list / n (or / sqrt(n)); double accumulation, fp32 result."""
    L = lib()
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rows = len(sel_ptr) - 1
    out = np.zeros((rows, feat.shape[1]), dtype=np.float32)
    sel_idx = np.ascontiguousarray(sel_idx, dtype=np.int32)
    sel_ptr = np.ascontiguousarray(sel_ptr, dtype=np.int64)
    L.pcgo_aggregate(_p(feat, C.c_float), feat.shape[1], C.c_int64(feat.shape[1]), _p(sel_ptr, C.c_int64),
                     _p(sel_idx, C.c_int32), C.c_int64(rows), 1 if norm == "rsqrt" else 0, _p(out, C.c_float))
    return out


def select_all(graph, targets, add_self):
    """GCN/SAGE selection on relation 0 of `graph`: the whole row (∪ self)."""
    L = lib()
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    ip, ix = graph.relation(0)
    ip = np.ascontiguousarray(ip, dtype=np.int64)
    ix = np.ascontiguousarray(ix, dtype=np.int32)
    deg = ip[targets.astype(np.int64) + 1] - ip[targets]
    sel_ptr = np.zeros(len(targets) + 1, dtype=np.int64)
    sel_idx = np.zeros(int(deg.sum()) + len(targets) + 1, dtype=np.int32)
    L.pcgo_select_all(_p(ip, C.c_int64), _p(ix, C.c_int32), C.c_int64(graph.n_nodes), _p(targets, C.c_int32),
                      len(targets), int(add_self), _p(sel_ptr, C.c_int64), _p(sel_idx, C.c_int32))
    return sel_ptr, sel_idx[:sel_ptr[-1]]


def pick_replay(cum, u):
    L = lib()
    cum = np.ascontiguousarray(cum, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros(len(u), dtype=np.int64)
    L.pcgo_pick_replay(_p(cum, C.c_double), C.c_int64(len(cum)), _p(u, C.c_double), C.c_int64(len(u)),
                       _p(out, C.c_int64))
    return out

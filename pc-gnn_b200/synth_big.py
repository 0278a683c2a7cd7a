"""Procedural power-law fraud graph for config C5 (BASELINE.json: 10M nodes, 3 relations, ~1e9 CSR
entries, 10 % positives, 64-d features, CSR row-partitioned over the GPUs), generated ON the GPU.

A dict-of-sets (the reference's graph format, /root/reference/src/utils.py:226-254) or even a host edge
list is infeasible at this size, so the graph is defined by a counter-based hash: row ``v`` of relation
``r`` has ``d(r, v)`` draws (Pareto-distributed, so hubs of 1e5 entries exist) whose columns are
``col(r, v, j)`` (skewed towards popular nodes), plus the self loop that ``sparse_to_adjlist`` adds
(utils.py:233); duplicates inside a row are merged like a ``set`` merges them and rows are id-sorted.
Because every quantity is a pure function of (seed, relation, node, draw), ANY rank can generate ANY row
range and all ranks agree on labels, split and pool without communication. (Rows are not symmetrised:
that would need a global edge exchange; the kernels do not rely on symmetry.)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .graph import RelGraph

__all__ = ["BigSpec", "BigPartition", "make_partition", "hash_uniform"]

_M1 = 0xBF58476D1CE4E5B9 - (1 << 64)
_M2 = 0x94D049BB133111EB - (1 << 64)
_G = 0x9E3779B97F4A7C15 - (1 << 64)


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """logical shift right of an int64 tensor (torch's >> is arithmetic)."""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping multiplies)."""
    x = (x ^ _lsr(x, 30)) * _M1
    x = (x ^ _lsr(x, 27)) * _M2
    return x ^ _lsr(x, 31)


def hash_uniform(seed: int, a: torch.Tensor, b=0) -> torch.Tensor:
    """float64 in [0, 1) from (seed, a, b), a / b int64 tensors (or ints)."""
    h = _mix(_mix(a * _G + seed) + b * _M1)
    return _lsr(h, 11).to(torch.float64) * (1.0 / (1 << 53))


@dataclass
class BigSpec:
    nodes_per_rank: int = 1_250_000
    feat_dim: int = 64
    rel_mean_deg: tuple = (4.0, 24.0, 72.0)     # ~100 entries per node over the three relations -> 1e9 at 10M nodes
    pareto_shape: float = 1.6
    max_degree: int = 200_000
    col_skew: float = 2.0                       # column = N * u^skew through a multiplicative permutation
    pos_rate: float = 0.10
    train_ratio: float = 0.4
    seed: int = 72


@dataclass
class BigPartition:
    spec: BigSpec
    graph: RelGraph               # this rank's rows (device CSR, global column ids)
    n_global: int
    row_lo: int
    feat: torch.Tensor            # [n_global, F] fp32 on the device (replicated: 2.56 GB at 10M x 64)
    labels: torch.Tensor          # [n_global] int64 (device)
    train_pos: torch.Tensor       # global ids of the train positives, ascending (device int32): the pool
    own_train: torch.Tensor       # global ids of this rank's train nodes (device int64)
    own_weights: torch.Tensor     # pick_step weights deg / LF of own_train (utils.py:274-278), float64

    def sample_batches(self, n_batches: int, batch: int, seed: int):
        """Label-balanced batches of THIS rank's nodes (targets are routed to the owner of their rows by
        construction: every rank draws its share of the global batch from its own node range)."""
        g = torch.Generator(device=self.own_weights.device)
        g.manual_seed(seed)
        out = []
        for _ in range(n_batches):
            idx = torch.multinomial(self.own_weights, batch, replacement=True, generator=g)
            nodes = self.own_train[idx]
            out.append((nodes.to(torch.int32), self.labels[nodes]))
        return out


def _degrees(spec: BigSpec, r: int, rows: torch.Tensor) -> torch.Tensor:
    """Pareto draws per row (before the self loop and the in-row de-duplication)."""
    a = spec.pareto_shape
    xmin = spec.rel_mean_deg[r] * (a - 1.0) / a
    u = hash_uniform(spec.seed + 101 * (r + 1), rows)
    d = torch.floor(xmin * (1.0 - u).clamp_min(1e-12) ** (-1.0 / a)).to(torch.int64)
    return d.clamp_(0, spec.max_degree)


def _columns(spec: BigSpec, r: int, n_global: int, rows: torch.Tensor, j: torch.Tensor) -> torch.Tensor:
    u = hash_uniform(spec.seed + 977 * (r + 1), rows, j)
    rank_ = torch.floor(u ** spec.col_skew * n_global).to(torch.int64).clamp_(0, n_global - 1)
    # multiplicative permutation so that the popular nodes are spread over the id range (and the partitions)
    mult = 6_700_417
    while np.gcd(mult, n_global) != 1:
        mult += 2
    return (rank_ * mult + 12_345) % n_global


def make_partition(spec: BigSpec, rank: int, world: int, device) -> BigPartition:
    dev = torch.device(device)
    n_rows = spec.nodes_per_rank
    n_global = n_rows * world
    lo = rank * n_rows
    rows_g = torch.arange(lo, lo + n_rows, dtype=torch.int64, device=dev)
    ips, ixs = [], []
    total = 0
    deg_sum = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    for r in range(len(spec.rel_mean_deg)):
        d = _degrees(spec, r, rows_g)
        ptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
        torch.cumsum(d, 0, out=ptr[1:])
        t = int(ptr[-1].item())
        local = torch.repeat_interleave(torch.arange(n_rows, dtype=torch.int64, device=dev), d, output_size=t)
        j = torch.arange(t, dtype=torch.int64, device=dev) - ptr[local]
        col = _columns(spec, r, n_global, local + lo, j)
        del j
        key = torch.cat([local * n_global + col, torch.arange(n_rows, dtype=torch.int64, device=dev) * n_global + rows_g])
        del local, col
        key = torch.sort(key).values
        key = torch.unique_consecutive(key)
        row_of = key // n_global
        ix = (key - row_of * n_global).to(torch.int32)
        cnt = torch.bincount(row_of, minlength=n_rows)
        del key, row_of
        ip = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt, 0, out=ip[1:])
        deg_sum += cnt
        ips.append(ip[:-1] + total)
        ixs.append(ix)
        total += int(ix.shape[0])
    indptr = torch.cat(ips + [torch.tensor([total], dtype=torch.int64, device=dev)])
    indices = torch.cat(ixs)
    del ips, ixs
    graph = RelGraph.from_device_csr(n_rows, len(spec.rel_mean_deg), indptr, indices, row_lo=lo, n_global=n_global)

    all_ids = torch.arange(n_global, dtype=torch.int64, device=dev)
    labels = (hash_uniform(spec.seed + 7, all_ids) < spec.pos_rate).to(torch.int64)
    train = hash_uniform(spec.seed + 9, all_ids) < spec.train_ratio
    train_pos = torch.nonzero(train & (labels == 1)).flatten().to(torch.int32)
    n_train = int(train.sum().item())
    n_train_pos = int(train_pos.shape[0])
    own_train = torch.nonzero(train[lo:lo + n_rows]).flatten() + lo
    own_lab = labels[own_train]
    # pick_step: weight = deg_homo / LF(label), LF = #train positives for a positive, #train nodes otherwise
    # (utils.py:276); deg_homo ~ sum of the relation rows minus the shared self loops
    deg_homo = (deg_sum[own_train - lo] - (len(spec.rel_mean_deg) - 1)).to(torch.float64)
    lf = torch.where(own_lab == 1, float(n_train_pos), float(n_train))
    gen = torch.Generator(device=dev)
    gen.manual_seed(spec.seed)
    feat = torch.rand((n_global, spec.feat_dim), dtype=torch.float32, device=dev, generator=gen)
    return BigPartition(spec, graph, n_global, lo, feat, labels, train_pos, own_train, deg_homo / lf)

"""Target-node data parallelism: one process per GPU, graph/features/pool replicated, each balanced
mini-batch's targets sharded across ranks, ONE collective per step (sum of the small parameter
gradient). The reference has no distributed code at all; this is new design.

The path's forward is embarrassingly parallel over targets (each (target, relation) item reads only
shared read-only tables), so there is no data-path collective; the only exchange is the gradient.
Works with any torch.distributed backend (nccl on GPUs, gloo for the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_batch", "GradAllReduce", "global_mean_loss_scale", "PeerRegion", "PeerComm", "FusedAdam"]


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) of an n-item batch owned by `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(nodes, labels, rank: int, world: int):
    lo, hi = shard_range(len(nodes), rank, world)
    return nodes[lo:hi], labels[lo:hi]


def global_mean_loss_scale(local_n: int, global_n: int, world: int) -> float:
    """The reference's losses are means over the batch (nn.CrossEntropyLoss, model.py:26). With
    unequal shards the local mean must be re-weighted so that the SUM over ranks of
    scale * local_mean_loss equals the global-batch mean."""
    return float(local_n) / float(global_n) if global_n else 0.0


class GradAllReduce:
    """Flat-buffer gradient all-reduce: pack every trainable grad into one fp32 buffer, one
    all_reduce(SUM), unpack. C2: 26,818 floats (~0.1 MB), C3: 139,210 floats: latency-bound, so one
    bucket and no overlap."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        self.n = n
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros((n + 3) // 4 * 4, dtype=torch.float32, device=dev)   # padded: float4 kernels
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def attach(self):
        """Make every .grad a view into the flat buffer so no pack/unpack copies are needed."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
        return self

    def zero(self):
        self.flat.zero_()

    def __call__(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            for p, v in zip(self.params, self.views):
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)
                    p.grad = v
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat


class PeerRegion:
    """`nbytes` of zeroed device memory on every rank of `group`, each mapped by all the other ranks of the node
    (CUDA IPC; loads, stores and atomics on a peer's region travel over NVLink / NVSwitch).
    ``regions`` is a ctypes array of this process's mappings (own region at [rank]); None when world == 1."""

    def __init__(self, nbytes: int, group=None):
        import ctypes as C

        from . import _lib

        self.lib = _lib.lib()
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.regions = None
        self.own = None
        self._mapped = None
        if self.world == 1:
            return
        if self.world > 8:
            raise ValueError("the peer-memory exchange is written for one node (<= 8 GPUs)")
        own = C.c_void_p()
        _lib.check(self.lib.pcg_comm_alloc(C.byref(own), int(nbytes)), "pcg_comm_alloc")
        self.own = own.value
        buf = C.create_string_buffer(64)
        _lib.check(self.lib.pcg_comm_export(self.own, buf), "pcg_comm_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(buf.raw), group=group)
        ptrs, self._mapped = [], []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(self.own)
                continue
            q = C.c_void_p()
            _lib.check(self.lib.pcg_comm_import(handles[r], C.byref(q)), "pcg_comm_import")
            ptrs.append(q.value)
            self._mapped.append(q.value)
        self.regions = (C.c_void_p * self.world)(*ptrs)
        dist.barrier(group=group)      # nobody signals into a region that is not mapped everywhere yet

    def close(self):
        if self._mapped:
            for q in self._mapped:
                self.lib.pcg_comm_unmap(q)
            self._mapped = None
        if self.own:
            self.lib.pcg_comm_free(self.own)
            self.own = None


class PeerComm(PeerRegion):
    """Receive areas (one slot per rank, double buffered) + flags of the fused gradient exchange (csrc/pcg_comm.cu).
    world == 1: no region."""

    def __init__(self, n_params: int, group=None):
        from . import _lib

        super().__init__(int(_lib.lib().pcg_comm_region_bytes(n_params)), group)


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) on a flat replica of the parameters, fused with the
    data-parallel gradient mean into one kernel (``pcg_allreduce_adam``). The parameters' ``.data`` and
    ``.grad`` become views into flat buffers; ``step()`` is CUDA-graph capturable (the step count lives on
    the device) and leaves the gradients zeroed for the next backward."""

    def __init__(self, reducer: GradAllReduce, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, comm=None):
        from . import _lib

        self.lib = _lib.lib()
        self.reducer, self.comm = reducer, comm
        self.lr, self.betas, self.eps, self.wd = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        flat = reducer.flat
        self.param = torch.zeros_like(flat)
        o = 0
        with torch.no_grad():
            for p in reducer.params:
                n = p.numel()
                self.param[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.param[o:o + n].view_as(p)
                o += n
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.state = torch.zeros(2, dtype=torch.int32, device=flat.device)      # [step count, ticket]

    @property
    def steps(self) -> int:
        return int(self.state[0].item())

    def step(self, do_adam: bool = True):
        from . import _lib

        c = self.comm
        world = c.world if c is not None else 1
        rc = self.lib.pcg_allreduce_adam(self.reducer.flat.data_ptr(), self.param.data_ptr(), self.m.data_ptr(),
                                         self.v.data_ptr(), self.reducer.flat.numel(),
                                         c.regions if (c is not None and world > 1) else None,
                                         c.rank if c is not None else 0, world, self.state.data_ptr(),
                                         self.state[1:].data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps,
                                         self.wd, int(do_adam), _lib.stream_ptr())
        _lib.check(rc, "pcg_allreduce_adam")

"""CUDA-graph training step: the whole step (score table, pool sort, choose, aggregate, relation
transforms, combine, head, both losses, backward, Adam) captured once and replayed per batch.

The reference's trainer (/root/reference/src/model_handler.py:128-156) rebuilds Python lists and
launches ~70 small library kernels per batch; once the hot path is a handful of kernels the step is
launch-bound, so the launches are recorded into a graph and only the batch's ids/labels are copied
into static device buffers before each replay. Per-batch sizes that depend on the data (how many
neighbours survive the filter) stay on the device: the kernels write counts into a status block and
work inside a fixed-capacity slot buffer sized from the epoch's batches (``plan``).
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["GraphedTrainStep"]


class GraphedTrainStep:
    """model: pcgnn_b200.model.PCALayer (hot path: the InterAgg at ``model.inter1``) or graphsage.GCN / GraphSage
    (hot path: ``model.enc.aggregator``); optimizer must be capturable (e.g. Adam(capturable=True, fused=True))."""

    def __init__(self, model, optimizer, batch_size: int, cap_slots: int, reducer=None, world: int = 1,
                 warmup_batch=None, use_pdl: bool = True):
        """optimizer: a capturable torch optimizer (gradient mean over ranks by NCCL between two graphs), or a
        ``parallel.FusedAdam`` (gradient exchange over peer memory + Adam inside the one step graph)."""
        from .engine import PinnedStaging
        from .parallel import FusedAdam

        self.use_pdl = bool(use_pdl)
        self.fused = isinstance(optimizer, FusedAdam)
        self.model, self.opt, self.B = model, optimizer, int(batch_size)
        self.reducer, self.world = reducer, world
        # the module that sizes the per-step slot buffer: InterAgg for PC-GNN, the row aggregator for GCN / SAGE
        self._sizer = model.inter1 if hasattr(model, "inter1") else model.enc.aggregator
        dev = next(p for p in model.parameters() if p.requires_grad).device
        self.dev = dev
        self._sizer.cap_slots_hint = int(cap_slots)
        self.cap_slots = int(cap_slots)
        if hasattr(model, "inter1") and model.inter1.engine().score_group is not None:
            model.inter1.scores_external = True
        if hasattr(model, "inter1"):
            model.inter1.center_on_side_stream = True
        # ids and labels of a step share ONE static device buffer ([B int64 labels | B int32 ids]) so that a host batch
        # is one pinned staging copy + one H2D; the pinned buffers form a ring guarded by events: back-to-back run()
        # calls never rewrite a buffer whose copy is still queued behind the previous replay
        self._packed = torch.zeros(3 * self.B, dtype=torch.int32, device=dev)
        self.labels = self._packed[:2 * self.B].view(torch.int64)
        self.nodes = self._packed[2 * self.B:]
        self._pin = PinnedStaging(3 * self.B, torch.int32)
        self._pin_loss = torch.zeros(1, dtype=torch.float32, pin_memory=True)
        self._loss_event = torch.cuda.Event()
        self.loss = None
        if reducer is not None:
            # the captured graph writes the gradients through the parameters' .grad tensors: they must BE the views
            # of the reducer's flat buffer (otherwise the graph keeps accumulating into tensors nobody zeroes)
            for p_, v_ in zip(reducer.params, reducer.views):
                if p_.grad is None or p_.grad.data_ptr() != v_.data_ptr():
                    reducer.attach()
                    break
        # row-partitioned graph: the score slices are exchanged by an eager all-gather BETWEEN two graphs
        # (slice kernel | all-gather | everything else), so no collective is captured
        eng = model.inter1.engine() if hasattr(model, "inter1") else None
        self._xeng = eng if (eng is not None and eng.score_group is not None) else None
        self.g_pre = None
        if warmup_batch is not None:
            self.nodes.copy_(torch.as_tensor(np.asarray(warmup_batch[0], dtype=np.int32)))
            self.labels.copy_(torch.as_tensor(np.asarray(warmup_batch[1], dtype=np.int64)))
        self._capture()

    # -- one eager step on the static buffers (also what gets captured) -------------------------
    def _score_pre(self):
        inter = self.model.inter1
        self._xeng.set_features(inter.features.weight)
        self._xeng.score_local(inter.label_clf.weight, inter.label_clf.bias)

    def _fwd_bwd(self):
        if self._xeng is not None and not torch.cuda.is_current_stream_capturing():
            self._score_pre()                      # eager warm-up steps: slice, exchange, then the forward
            self._xeng.score_exchange()
        if self.fused:
            pass                                   # the fused step leaves the gradients zeroed
        elif self.reducer is not None:
            self.reducer.zero()
        else:
            self.opt.zero_grad(set_to_none=False)
        eng = self.model.inter1.engine() if (self.fused and hasattr(self.model, "inter1")) else None
        if eng is not None:
            # every parameter has exactly one producer kernel and the buffer is zero: let the kernels store the
            # gradients in place (no AccumulateGrad adds in the graph)
            eng.grad_sink = {p.data_ptr(): v for p, v in zip(self.reducer.params, self.reducer.views)}
        try:
            if eng is not None:
                eng.grads_in_sinks = False
            loss = self.model.loss(self.nodes, self.labels)
            if eng is None or not eng.grads_in_sinks:    # fused pass: the forward launch stored every gradient
                loss.backward()
        finally:
            if eng is not None:
                eng.grad_sink = None
        return loss.detach()

    def _capture(self):
        # programmatic dependent launch for the kernels recorded below (see include/pcgnn_b200.h: pcg_set_pdl).
        # Measured on C2 (profiles/README.md): the fused dense kernel and the exchange + Adam kernel gain from starting
        # under the tail of the kernel in front (mask 2 | 8: 91.2 -> 87.8 us per replay); the cluster-launched
        # weight-gradient kernel and the pool sort lose (mask 4: +10 us), so they keep full dependencies.
        from . import _lib

        import os

        mask = int(os.environ.get("PCG_PDL_MASK", "10")) if self.use_pdl else 0
        prev = _lib.lib().pcg_set_pdl(mask)
        try:
            self._capture_graphs()
        finally:
            _lib.lib().pcg_set_pdl(prev)

    def _capture_graphs(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):       # warm-up on a side stream (allocator, lazy inits, cuBLAS handles)
            for _ in range(3):
                if self.reducer is None:
                    for p in self.model.parameters():
                        if p.requires_grad and p.grad is None:
                            p.grad = torch.zeros_like(p)
                self._fwd_bwd()
                if self.fused:
                    self.reducer.zero()
                # no optimizer step during warm-up: parameters stay as the caller initialised them
            if self.fused:
                torch.cuda.current_stream(self.dev).wait_stream(s)
        if self.fused:
            torch.cuda.synchronize(self.dev)
            if self._xeng is not None:
                self.g_pre = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g_pre):
                    self._score_pre()
            self.g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fb):
                self.loss = self._fwd_bwd()
                self.opt.step()                    # gradient mean over the ranks + Adam, one kernel
            self.g_opt = None
            return
        with torch.cuda.stream(s):
            # Create the optimizer state OUTSIDE the graph: torch builds it lazily in the first step(), and a
            # captured lazy init would re-zero the moments on every replay. Step once, then undo it.
            params = [p for g in self.opt.param_groups for p in g["params"] if p.requires_grad]
            backup = [p.detach().clone() for p in params]
            had_state = len(self.opt.state) > 0
            saved = None
            if had_state:
                saved = {k: {n: (v.clone() if isinstance(v, torch.Tensor) else v) for n, v in st.items()}
                         for k, st in self.opt.state.items()}
            self.opt.step()
            with torch.no_grad():
                for p, b in zip(params, backup):
                    p.copy_(b)
                for k, st in self.opt.state.items():
                    for n, v in st.items():
                        if isinstance(v, torch.Tensor):
                            v.copy_(saved[k][n]) if had_state else v.zero_()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        if self._xeng is not None:
            self.g_pre = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_pre):
                self._score_pre()
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self.loss = self._fwd_bwd()
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt):
            self.opt.step()

    def _replay(self):
        if self.g_pre is not None:
            self.g_pre.replay()
            self._xeng.score_exchange()
        self.g_fb.replay()
        if self.fused:
            return self.loss
        if self.world > 1 and self.reducer is not None:
            self.reducer()                       # one NCCL all-reduce of the flat gradient
            self.reducer.flat.div_(self.world)
        self.g_opt.replay()
        return self.loss

    # -- public --------------------------------------------------------------------------------
    def run_device(self, nodes_dev: torch.Tensor, labels_dev: torch.Tensor):
        """Batch already in HBM (int32 ids, int64 labels). Returns the (device) loss tensor."""
        self.nodes.copy_(nodes_dev, non_blocking=True)
        self.labels.copy_(labels_dev, non_blocking=True)
        return self._replay()

    def run(self, nodes, labels):
        """Batch on the host (numpy arrays, or lists, of ids and labels): one pinned staging copy + one H2D + replay.
        Returns the (device) loss tensor."""
        B = self.B
        host = np.empty(3 * B, dtype=np.int32)
        host[:2 * B].view(np.int64)[:] = labels
        host[2 * B:] = nodes
        self._pin.upload(host, self._packed)
        return self._replay()

    def run_item(self, nodes, labels) -> float:
        """``run`` + the loss as a Python float (device->host through a pinned word and an event; cheaper than
        ``.item()``'s pageable copy)."""
        loss = self.run(nodes, labels)
        self._pin_loss.copy_(loss.reshape(1), non_blocking=True)
        self._loss_event.record(torch.cuda.current_stream(self.dev))
        self._loss_event.synchronize()
        return float(self._pin_loss[0])

    def overflowed(self) -> bool:
        """True if ANY replay since the last call needed more slots than the captured capacity (its results were
        incomplete: re-plan with a larger capacity). PC-GNN models: a workspace word every choose call ORs its flag
        into (``Engine.overflow_since_reset``); GCN / SAGE: the last replay's status block. Syncs."""
        if hasattr(self.model, "inter1"):
            return self.model.inter1.engine().overflow_since_reset()
        sel = self._sizer.last_selection
        return bool(sel is not None and sel.overflowed())

    @staticmethod
    def plan(engine, batches, thresholds, rho) -> int:
        """Slot capacity covering every batch of an epoch plan (host arithmetic on CSR offsets)."""
        return max(engine.slots_bound(np.asarray(n, dtype=np.int32), thresholds, rho, True) for n in batches)

"""Host-side shortcuts for the reference's OWN training loop (opt in with ``fastloop.enable()``).

The reference trains with (/root/reference/src/model_handler.py:124, :142-156)

    optimizer = torch.optim.Adam(filter(requires_grad, gnn_model.parameters()), lr=..., weight_decay=...)
    optimizer.zero_grad(); loss = gnn_model.loss(batch_nodes, labels); loss.backward(); optimizer.step()

Once ``model.loss`` is one graph replay (stepgraph.StepGraphCache) that loop is bound by what torch does on the host
around it (profiles/r02_replay_cost.txt): the autograd engine hand-off for a graph of ONE node (~140 us) and
``torch.optim.Adam.step`` for seven small tensors (~240 us). Both are pure bookkeeping here: the replay has already
computed every parameter gradient (dLoss = 1), and the package has a one-kernel Adam (``pcg_allreduce_adam``, the
optimizer of runtime.GraphedTrainStep). With ``enable()``:

  * the loss returned by ``model.loss`` is a ``torch.Tensor`` subclass whose ``backward()`` -- called the way the
    reference calls it: no arguments, once -- stores the replay's gradients into ``p.grad`` directly (it still carries
    the autograd node, so ``torch.autograd.backward``, ``(loss * 2).backward()`` etc. take the normal route);
  * a global optimizer pre-step hook recognises a plain ``torch.optim.Adam`` over exactly those parameters (no amsgrad /
    maximize / capturable / differentiable, no state yet), moves the parameters into one flat buffer (``p.data``
    becomes a view; ``state_dict`` / ``load_state_dict`` keep working) and runs the fused Adam kernel on it -- same
    formula as ``torch.optim.Adam`` with L2 weight decay, hyper-parameters read from ``param_groups`` on every step so
    schedulers work -- then clears the gradients so that torch's own ``step`` finds nothing left to do.

Anything else (another optimizer, gradient accumulation over several losses, hooks on the loss) falls through to
torch's own code paths unchanged. Moments live in this module's buffers, not in ``optimizer.state``.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["enable", "disable", "enabled", "StepLoss"]

_HOOK = None
_LAST = None          # the most recent fast backward: (params, flat gradient, views)


def enabled() -> bool:
    return _HOOK is not None


class StepLoss(torch.Tensor):
    """Loss of a replayed training step. ``backward()`` with no arguments hands out the gradients the replay stored."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    def backward(self, gradient=None, retain_graph=None, create_graph=False, inputs=None):
        job = self.__dict__.get("_pcg_job")
        if job is None or gradient is not None or create_graph or inputs is not None or not enabled():
            return super().backward(gradient, retain_graph, create_graph, inputs)
        self.__dict__["_pcg_job"] = None if not retain_graph else job
        flat_static, views, params = job
        global _LAST
        g = flat_static.clone()             # the static buffer is rewritten by the next replay
        for p, (off, n, shape) in zip(params, views):
            v = g[off:off + n].view(shape)
            if p.grad is None:
                p.grad = v
            else:
                p.grad.add_(v)
        _LAST = (params, g, views)


def wrap(loss: torch.Tensor, flat_static, views, params):
    out = loss.as_subclass(StepLoss)
    out.__dict__["_pcg_job"] = (flat_static, views, tuple(params))
    return out


class _FlatAdam:
    """Flat replica of the parameters + moments for ``pcg_allreduce_adam`` (world 1), in the layout of the replay's
    gradient buffer."""

    def __init__(self, params, views, n_flat, device):
        self.ids = tuple(id(p) for p in params)
        self.views = views
        self.param = torch.zeros(n_flat, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, (off, n, shape) in zip(params, views):
                self.param[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.param[off:off + n].view(shape)
        self.m = torch.zeros_like(self.param)
        self.v = torch.zeros_like(self.param)
        self.state = torch.zeros(2, dtype=torch.int32, device=device)    # [steps done, ticket]

    def step(self, g, lr, betas, eps, wd):
        rc = _lib.lib().pcg_allreduce_adam(g.data_ptr(), self.param.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                           g.numel(), None, 0, 1, self.state.data_ptr(), self.state[1:].data_ptr(),
                                           float(lr), float(betas[0]), float(betas[1]), float(eps), float(wd), 1,
                                           _lib.stream_ptr())
        _lib.check(rc, "pcg_allreduce_adam")


def _pre_step(opt, args, kwargs):
    global _LAST
    last, _LAST = _LAST, None
    if last is None or type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
        return None
    params, g, views = last
    grp = opt.param_groups[0]
    if grp.get("amsgrad") or grp.get("maximize") or grp.get("capturable") or grp.get("differentiable"):
        return None
    # torch passes the wrapper's positional arguments: (optimizer, [closure])
    if (len(args) > 1 and args[1] is not None) or kwargs.get("closure") is not None:
        return None
    mine = opt.__dict__.get("_pcg_flat")
    if mine is None and len(opt.state) > 0:
        return None                         # the optimizer already keeps moments of its own: leave it alone
    theirs = grp["params"]
    if len(theirs) != len(params) or {id(p) for p in theirs} != {id(p) for p in params}:
        return None
    for p, (off, n, shape) in zip(params, views):   # the gradients must still be the ones backward() stored
        if p.grad is None or p.grad.data_ptr() != g.data_ptr() + 4 * off:
            return None
    if mine is None or mine.ids != tuple(id(p) for p in params) or mine.param.numel() != g.numel():
        if mine is not None:
            return None
        mine = opt.__dict__["_pcg_flat"] = _FlatAdam(params, views, g.numel(), g.device)
    lr = grp["lr"]
    mine.step(g, float(lr) if not isinstance(lr, torch.Tensor) else float(lr.item()), grp["betas"], grp["eps"],
              grp["weight_decay"])
    for p in params:
        p.grad = None                       # torch's step() now finds nothing to update
    return None


def enable():
    """Turn the shortcuts on (process-wide; idempotent)."""
    global _HOOK
    if _HOOK is None:
        from torch.optim.optimizer import register_optimizer_step_pre_hook

        _HOOK = register_optimizer_step_pre_hook(_pre_step)


def disable():
    global _HOOK, _LAST
    if _HOOK is not None:
        _HOOK.remove()
        _HOOK = None
    _LAST = None

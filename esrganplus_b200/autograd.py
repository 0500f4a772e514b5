"""Training-mode (autograd) entry of the generator: one autograd.Function whose forward/backward are the
native esrp_rrdbnet_train_forward / esrp_rrdbnet_backward (SRRaGAN_model.py:120,140 drive them through
``self.netG(self.var_L)`` and ``l_g_total.backward()``).  Gradients land in ``.grad`` of the fp32 OIHW
Parameters, which is what ``torch.optim.Adam`` (SRRaGAN_model.py:82-89) consumes.

Data parallelism (one process per GPU): when ``torch.distributed`` is initialised and the module was
marked with ``esrganplus_b200.data_parallel(module)``, the flat gradient buffer the backward fills is
all-reduced (average) with ONE collective before the per-tensor views are handed to autograd — the only
exchange step of the path.
"""
from __future__ import annotations

from typing import Dict

import torch


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, eng, noise, seed, x, *params):
        y, token = eng.train_forward(x, noise, seed)
        ctx.eng = eng
        ctx.token = token
        ctx.module = module
        return y

    @staticmethod
    def backward(ctx, dy):
        needs = list(ctx.needs_input_grad[5:])
        grads, flat = ctx.eng.backward(dy, ctx.token, needs)
        allreduce_flat(ctx.module, flat)
        return (None, None, None, None, None) + tuple(grads)


class _GeneratorFnFlat(torch.autograd.Function):
    """The same forward / backward with ONE autograd input standing for all parameters (`anchor`, a dummy leaf): the
    backward writes the 771 gradients into the engine's persistent flat buffer and points each ``p.grad`` at its slice
    itself, instead of returning 771 tensors for autograd to accumulate (3-4 ms of host time per step at 4 crops per GPU,
    where the whole device step is ~15 ms).  Semantics: ``.grad`` is OVERWRITTEN by every backward (the reference's solver
    zeroes gradients before every backward, SRRaGAN_model.py:118,147); ``torch.autograd.grad`` w.r.t. parameters is not
    served by this path."""

    @staticmethod
    def forward(ctx, module, eng, noise, seed, x, anchor):
        y, token = eng.train_forward(x, noise, seed)
        ctx.eng, ctx.token, ctx.module = eng, token, module
        return y

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.eng
        flat = eng.backward_flat(dy, ctx.token)
        allreduce_flat(ctx.module, flat)
        views = eng.grad_views
        for p, v in zip(ctx.module._esrp_flat_plist, views):
            if p.grad is not v:
                p.grad = v
        return None, None, None, None, None, None


def enable_flat_grads(module, enable: bool = True):
    """Opt a generator in to `_GeneratorFnFlat` (see there).  Used by gan_step.GanTrainStep with the native solver; needs
    every parameter to require grad."""
    object.__setattr__(module, "_esrp_flat_grads", bool(enable))
    return module


def allreduce_flat(module, flat: torch.Tensor) -> None:
    """The one exchange step of data-parallel training: average the flat gradient buffer over the ranks of the
    module's process group (NCCL over NVLink on the GPU box; gloo in the CPU tests).  No-op unless the module was
    marked by `data_parallel` and torch.distributed is initialised."""
    group = getattr(module, "_dp_group", False)
    if group is False:
        return
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    pg = None if group is True else group
    ws = dist.get_world_size(pg)
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=pg)
        flat.mul_(1.0 / ws)


def generator_apply(module, x: torch.Tensor, params: Dict[str, torch.Tensor]) -> torch.Tensor:
    if x.requires_grad:
        raise NotImplementedError(
            "esrganplus_b200.RRDBNet: the gradient w.r.t. the LR input is not produced (the reference never asks "
            "for it, SRRaGAN_model.py:103-111); detach the input")
    eng = module._engine_for(x.device)
    eng.sync_weights(params, module.weights_epoch)
    plist = []
    for k in eng.keys:
        plist.append(params[k])
    noise = bool(module.training)
    if getattr(module, "_esrp_flat_grads", False) and all(p.requires_grad for p in plist):
        anchor = module.__dict__.get("_esrp_anchor")
        if anchor is None or anchor.device != x.device:
            anchor = torch.zeros((), device=x.device, requires_grad=True)
            object.__setattr__(module, "_esrp_anchor", anchor)
        object.__setattr__(module, "_esrp_flat_plist", plist)
        return _GeneratorFnFlat.apply(module, eng, noise, module.noise_seed() if noise else 0, x, anchor)
    return _GeneratorFn.apply(module, eng, noise, module.noise_seed() if noise else 0, x, *plist)


def data_parallel(module, process_group=True):
    """Mark a generator/discriminator for data-parallel training: its backward all-reduces (averages) the
    flat gradient buffer over `process_group` (True = the default group).  Parameters must start identical
    on every rank (broadcast them once, e.g. ``broadcast_parameters``)."""
    object.__setattr__(module, "_dp_group", process_group)
    return module


def broadcast_parameters(module, src: int = 0, process_group=None) -> None:
    import torch.distributed as dist
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)

"""Native solver arithmetic between the forward / backward passes (csrc/esrp_solver.cu).

* ``FlatAdam``: ``torch.optim.Adam`` arithmetic (SRRaGAN_model.py:82-89) as ONE kernel over flat fp32 storage.  It IS a
  ``torch.optim.Optimizer`` (``param_groups``, ``state_dict``), so the reference's ``lr_scheduler.MultiStepLR``
  (SRRaGAN_model.py:91-95) drives it unchanged.
* ``ragan_bce_terms`` / ``l1_loss``: the relativistic-average BCE terms (SRRaGAN_model.py:133-136,151-154) and the L1 pixel
  loss (:123) as single launches with analytic gradients, as ``torch.autograd.Function``s.

No CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Tuple

import torch

from . import _lib


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class FlatAdam(torch.optim.Optimizer):
    """Adam over the parameters of ONE module, kept in one flat fp32 buffer.

    At construction the parameters are re-homed: ``p.data`` becomes a view of a flat buffer whose slices are padded to
    16 bytes — the layout of the flat gradient buffer the native backward passes fill — so a step is one elementwise
    kernel over (parameters, gradients, exp_avg, exp_avg_sq) and, when the gradients already are views of one such
    buffer (they are after ``loss.backward()`` through the native modules), no gather either.  Parameter objects, names
    and ``state_dict()`` are unchanged.  ``module`` (optional): its derived bf16 weight cache is invalidated after
    every step (the kernel writes parameter memory without bumping tensor version counters).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, module: Optional[torch.nn.Module] = None):
        params = [p for p in params]
        if not params:
            raise ValueError("FlatAdam: no parameters")
        dev = params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in params):
            raise RuntimeError("FlatAdam: parameters must be fp32 tensors on one CUDA device")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.lib = _lib.load()
        self.device, self.module = dev, module
        self._plist: List[torch.nn.Parameter] = params
        self._offs, off = [], 0
        for p in params:
            self._offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self._total = off
        self._flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self._flat_m = torch.zeros_like(self._flat_p)
        self._flat_v = torch.zeros_like(self._flat_p)
        self._flat_g: Optional[torch.Tensor] = None
        self._t = 0
        self._step_tensor = torch.zeros((), dtype=torch.float32)
        self._rehome()
        for p, o in zip(params, self._offs):
            n = p.numel()
            self.state[p] = {"step": self._step_tensor, "exp_avg": self._flat_m[o:o + n].view(p.shape),
                             "exp_avg_sq": self._flat_v[o:o + n].view(p.shape)}

    def _rehome(self) -> None:
        base = self._flat_p.data_ptr()
        if all(p.data_ptr() == base + o * 4 for p, o in zip(self._plist, self._offs)):
            return   # (the common case, ~0.1 ms for 771 tensors)
        with torch.no_grad():
            for p, o in zip(self._plist, self._offs):
                n = p.numel()
                view = self._flat_p[o:o + n].view(p.shape)
                if p.data.data_ptr() != view.data_ptr():
                    view.copy_(p.data)
                    p.data = view

    def _flat_grads(self) -> torch.Tensor:
        """The gradients as one flat buffer in the padded layout: the buffer they are views of, or a gathered copy."""
        p0 = self._plist[0]
        if p0.grad is None:
            raise RuntimeError("FlatAdam.step: a parameter has no gradient")
        base = p0.grad.data_ptr() - self._offs[0] * 4
        ok = base % 16 == 0
        if ok:
            for p, o in zip(self._plist, self._offs):
                g = p.grad
                if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base + o * 4:
                    ok = False
                    break
        if ok:
            st = p0.grad.untyped_storage()
            lo = base - st.data_ptr()
            if lo >= 0 and lo + self._total * 4 <= st.nbytes():
                flat = torch.empty(0, dtype=torch.float32, device=self.device).set_(st, lo // 4, (self._total,))
                return flat
        if self._flat_g is None:
            self._flat_g = torch.zeros(self._total, dtype=torch.float32, device=self.device)
        views = [self._flat_g[o:o + p.numel()].view(p.shape) for p, o in zip(self._plist, self._offs)]
        grads = []
        for p in self._plist:
            if p.grad is None:
                raise RuntimeError("FlatAdam.step: a parameter has no gradient")
            grads.append(p.grad)
        torch._foreach_copy_(views, grads)
        return self._flat_g

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        grp = self.param_groups[0]
        self._rehome()   # (.to() / load_state_dict through .data may have replaced storages)
        g = self._flat_grads()
        self._t += 1
        self._step_tensor.fill_(float(self._t))
        b1, b2 = grp["betas"]
        _lib.check(self.lib.esrp_adam_flat(self._flat_p.data_ptr(), g.data_ptr(), self._flat_m.data_ptr(), self._flat_v.data_ptr(),
                                           self._total, float(grp["lr"]), float(b1), float(b2), float(grp["eps"]),
                                           float(grp["weight_decay"]), self._t, _stream(self.device)), "esrp_adam_flat")
        if self.module is not None:
            self.module.invalidate_weights()
        return loss

    def zero_grad(self, set_to_none: bool = True):
        # gradients are overwritten (not accumulated) by every native backward; dropping the references is enough
        for p in self._plist:
            p.grad = None


class _RaganTerms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_real, pred_fake, t_real, t_fake):
        lib = _lib.load()
        r, f = pred_real.contiguous().view(-1), pred_fake.contiguous().view(-1)
        n = r.numel()
        if f.numel() != n or r.dtype != torch.float32 or f.dtype != torch.float32 or not r.is_cuda:
            raise RuntimeError("ragan_bce_terms expects two fp32 CUDA logit tensors of the same size")
        out = torch.empty(4 + 4 * n, dtype=torch.float32, device=r.device)
        g = out[4:].view(4, n)
        _lib.check(lib.esrp_ragan_bce(r.data_ptr(), f.data_ptr(), n, float(t_real), float(t_fake), out.data_ptr(), g[0].data_ptr(),
                                      g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), _stream(r.device)), "esrp_ragan_bce")
        ctx.save_for_backward(g)
        ctx.shapes = (pred_real.shape, pred_fake.shape)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, gA, gB):
        (g,) = ctx.saved_tensors
        sr, sf = ctx.shapes
        d_real = d_fake = None
        if ctx.needs_input_grad[0]:
            d_real = (gA * g[0] + gB * g[2]).view(sr)
        if ctx.needs_input_grad[1]:
            d_fake = (gA * g[1] + gB * g[3]).view(sf)
        return d_real, d_fake, None, None


def ragan_bce_terms(pred_real: torch.Tensor, pred_fake: torch.Tensor, t_real: float, t_fake: float):
    """(BCEWithLogits(pred_real - mean(pred_fake), t_real), BCEWithLogits(pred_fake - mean(pred_real), t_fake)): the two
    relativistic terms of SRRaGAN_model.py:133-136 (t = 0, 1) and :151-154 (t = 1, 0), one launch, analytic gradients."""
    return _RaganTerms.apply(pred_real, pred_fake, t_real, t_fake)


class _L1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        if a.shape != b.shape or a.dtype != torch.float32 or b.dtype != torch.float32 or not a.is_cuda or a.numel() % 4:
            raise RuntimeError("l1_loss expects two fp32 CUDA tensors of one shape with numel % 4 == 0")
        a, b = a.contiguous(), b.contiguous()
        need = ctx.needs_input_grad[0]
        grad = torch.empty_like(a) if need else None
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        scratch = torch.empty(1, dtype=torch.float64, device=a.device)
        _lib.check(lib.esrp_l1_loss_grad(a.data_ptr(), b.data_ptr(), a.numel(), grad.data_ptr() if need else None, loss.data_ptr(),
                                         scratch.data_ptr(), _stream(a.device)), "esrp_l1_loss_grad")
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("l1_loss: gradient w.r.t. the target is not produced (SRRaGAN_model.py:123 never asks)")
        return (ctx.grad * g if ctx.grad is not None else None), None


def l1_loss(a: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.L1Loss()(a, target) (SRRaGAN_model.py:31-39,123): loss and d loss / d a in one pass."""
    return _L1.apply(a, target)

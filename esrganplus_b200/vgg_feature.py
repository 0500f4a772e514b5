"""Perceptual branch: `VGGFeatureExtractor` (codes/models/modules/architecture.py:279-307) forward and input gradient on
the tcgen05 conv kernels.

torchvision's vgg19().features[:35] is sixteen 3x3 / stride-1 / pad-1 convs (3 -> 64 -> 64 | 128 x2 | 256 x4 | 512 x4 |
512 x4, `|` = MaxPool2d(2,2)), each followed by ReLU except the last (feature_layer = 34 is conv5_4 BEFORE its ReLU,
networks.py:144-148).  On 128x128 HR crops that is 12.7 GFLOP per image, run three times per G step (real features, fake
features, gradient w.r.t. the fake image; SRRaGAN_model.py:127-130) with frozen weights (architecture.py:300-301).

The engine is the discriminator's (discriminator.py): NHWC bf16 activations, one `esrp_conv3x3_nhwc` launch per 32 / 64
output channels with the ReLU fused (act = 2), launch plans recorded on the first pass of a shape and replayed as CUDA
graphs, packed weight tiles rebuilt in place when a Parameter changes.  Added here: the input normalisation fused into the
NCHW -> NHWC converter, max-pool forward, and the gradient through [ReLU (, MaxPool)] as one element-wise pass per layer
that reads the layer's stored OUTPUT (csrc/esrp_vgg.cu).  Data gradients run on the same conv kernels over the
transposed / tap-flipped weight pack.  The features come back as fp32 NCHW like the reference's.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from . import conv as K
from .discriminator import DiscriminatorEngine, _Lease, _Layer, _Plan, _stream


class FeatureEngine(DiscriminatorEngine):
    """Launch plans of one VGGFeatureExtractor on one device."""

    def __init__(self, module: nn.Module, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        self.plans: Dict[tuple, List[_Plan]] = {}
        self.nside = 4
        self.side = (C.c_void_p * self.nside)()
        self.cap = (C.c_void_p * 1)()
        with torch.cuda.device(device):
            _lib.check(self.lib.esrp_streams_create(self.nside, self.side), "esrp_streams_create")
            _lib.check(self.lib.esrp_streams_create(1, self.cap), "esrp_streams_create")
        self._rec = self._keep = None
        self._ptr_sig = None
        self.epoch = 0
        self.layers: List[_Layer] = []
        self.pool_after: List[bool] = []     # MaxPool2d(2,2) after the layer's ReLU
        self.relu_after: List[bool] = []
        feats = list(module.features)
        for i, m in enumerate(feats):
            if isinstance(m, nn.Conv2d):
                if m.kernel_size != (3, 3) or m.stride != (1, 1) or m.padding != (1, 1) or m.groups != 1 or m.dilation != (1, 1):
                    raise NotImplementedError("VGGFeatureExtractor: 3x3 / stride-1 / pad-1 convs only")
                relu = i + 1 < len(feats) and isinstance(feats[i + 1], nn.ReLU)
                pool = relu and i + 2 < len(feats) and isinstance(feats[i + 2], nn.MaxPool2d)
                L = _Layer(m, None)
                L.act_code = 2 if relu else 0
                self.layers.append(L)
                self.relu_after.append(relu)
                self.pool_after.append(pool)
            elif isinstance(m, nn.MaxPool2d):
                if not (m.kernel_size in (2, (2, 2)) and m.stride in (2, (2, 2)) and m.padding in (0, (0, 0)) and not m.ceil_mode):
                    raise NotImplementedError("VGGFeatureExtractor: MaxPool2d(kernel_size=2, stride=2) only")
            elif not isinstance(m, nn.ReLU):
                raise NotImplementedError(f"VGGFeatureExtractor: unsupported layer {type(m).__name__} (use_bn=False only)")
        if not self.layers or self.relu_after[-1]:
            raise NotImplementedError("VGGFeatureExtractor: the feature layer must be a conv (features before the ReLU)")
        self.use_input_norm = bool(getattr(module, "use_input_norm", False))
        self._norm_sig = None
        self._scale = self._shift = None

    # the plan key of the base class looks at BatchNorm modes; there are none here
    def _acquire(self, module: nn.Module, shape) -> _Plan:
        sig = tuple(t.data_ptr() for t in list(module.parameters()) + list(module.buffers()))
        if sig != self._ptr_sig:
            self.plans.clear()
            self._ptr_sig = sig
        st = _stream(refresh=True)
        key = (tuple(shape), st)
        pool = self.plans.setdefault(key, [])
        for pl in pool:
            if not pl.busy:
                return pl
        if len(pool) >= 6:
            raise RuntimeError("VGGFeatureExtractor: more than 6 forward passes of one shape are waiting for their backward")
        pl = _Plan(key, st)
        pool.append(pl)
        return pl

    def _norm(self, module: nn.Module):
        """scale = 1 / std, shift = -mean / std (architecture.py:304-305), rebuilt IN PLACE when the buffers change."""
        if not self.use_input_norm:
            return None, None
        mean, std = module.mean, module.std
        sig = (mean.data_ptr(), mean._version, std.data_ptr(), std._version)
        if sig != self._norm_sig:
            with torch.no_grad():
                sc = (1.0 / std.detach().float().flatten()).contiguous()
                sh = (-mean.detach().float().flatten() / std.detach().float().flatten()).contiguous()
                if self._scale is None:
                    self._scale, self._shift = sc.to(self.device), sh.to(self.device)
                else:
                    self._scale.copy_(sc)
                    self._shift.copy_(sh)
            self._norm_sig = sig
        return self._scale, self._shift

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, module: nn.Module, x: torch.Tensor, lease: bool = False):
        if x.dim() != 4 or x.dtype != torch.float32 or x.device != self.device:
            raise RuntimeError("VGGFeatureExtractor forward expects an fp32 NCHW CUDA tensor")
        n, c, h, w = x.shape
        npool = sum(self.pool_after)
        if c != self.layers[0].cin or h % (1 << npool) or w % (1 << npool):
            raise RuntimeError(f"VGGFeatureExtractor: input must have {self.layers[0].cin} channels and a height / width divisible by "
                               f"{1 << npool} (got {tuple(x.shape)})")
        x = x.contiguous()
        self.epoch = module.__dict__.get("_esrp_epoch", 0)
        for L in self.layers:
            self._sync(L)
        self._norm(module)
        plan = self._acquire(module, x.shape)
        if plan.fwd is None:
            self._rec, self._keep = [], plan.keep
            try:
                self._record_forward(module, x, plan)
                plan.fwd = self._rec
            finally:
                self._rec = self._keep = None
        else:
            fn, args, what = plan.fwd[plan.x_idx]
            if plan.fwd_graph is None:
                if plan.x_in is None:
                    plan.x_in = torch.empty_like(x)
                plan.fwd[plan.x_idx] = (fn, (plan.x_in.data_ptr(),) + args[1:], what)
                plan.fwd_graph = self._capture(plan.fwd, plan.stream)
            if plan.fwd_graph:
                plan.x_in.copy_(x)
                _lib.check(self.lib.esrp_graph_launch(plan.fwd_graph, plan.stream), "esrp_graph_launch")
            else:
                plan.fwd[plan.x_idx] = (fn, (x.data_ptr(),) + args[1:], what)
                self._replay(plan.fwd)
        # plan.out is NHWC fp32; the reference returns NCHW
        return plan.out.permute(0, 3, 1, 2).contiguous(), (_Lease(plan) if lease else None)

    def _record_forward(self, module: nn.Module, x: torch.Tensor, plan: _Plan) -> None:
        st = _stream()
        n, c, h, w = x.shape
        L0 = self.layers[0]
        act = self._t(torch.empty((n, h, w, L0.cin_eff), dtype=torch.bfloat16, device=self.device))
        plan.x_idx = len(self._rec)
        sc, sh = self._scale, self._shift
        self._c(self.lib.esrp_nchw_f32_to_nhwc_bf16_affine, "esrp_nchw_f32_to_nhwc_bf16_affine", x.data_ptr(),
                sc.data_ptr() if sc is not None else None, sh.data_ptr() if sh is not None else None, act.data_ptr(), n, c, h, w,
                L0.cin_eff, st)
        saved = []
        for li, L in enumerate(self.layers):
            last = li == len(self.layers) - 1
            y, hv, wv = self._conv(L, act, fuse_act_bf16=not last)
            rec = dict(src=act, y=y, pooled=None)
            if self.pool_after[li]:
                pooled = self._t(torch.empty((n, hv // 2, wv // 2, L.cout), dtype=torch.bfloat16, device=self.device))
                self._c(self.lib.esrp_maxpool2x2_nhwc_bf16, "esrp_maxpool2x2_nhwc_bf16", y.data_ptr(), pooled.data_ptr(), n, hv, wv, L.cout, st)
                rec["pooled"] = pooled
                act = pooled
            else:
                act = y
            saved.append(rec)
        if act.dtype != torch.float32:
            raise RuntimeError("VGGFeatureExtractor: internal: the feature layer did not produce fp32 output")
        plan.saved = saved
        plan.out = act

    # -- backward: gradient w.r.t. the input image only (the weights are frozen, architecture.py:300-301) ----------
    def backward(self, module: nn.Module, plan: _Plan, dout: torch.Tensor):
        if _stream(refresh=True) != plan.stream:
            raise RuntimeError("VGGFeatureExtractor backward must run on the CUDA stream its forward ran on")
        if plan.dout_buf is None:
            plan.dout_buf = torch.empty(tuple(dout.shape), dtype=torch.float32, device=self.device)
        plan.dout_buf.copy_(dout)
        ent = plan.bwd.get("dx")
        if ent is None:
            self._rec, self._keep = [], plan.keep
            try:
                dx = self._record_backward(module, plan)
                ent = plan.bwd["dx"] = (self._rec, dx)
            finally:
                self._rec = self._keep = None
        else:
            gr = plan.bwd_graph.get("dx")
            if gr is None:
                gr = plan.bwd_graph["dx"] = self._capture(ent[0], plan.stream)
            if gr:
                _lib.check(self.lib.esrp_graph_launch(gr, plan.stream), "esrp_graph_launch")
            else:
                self._replay(ent[0])
        dx = ent[1]
        if self._scale is not None:   # d/dx of (x - mean) / std
            return dx * self._scale.view(1, -1, 1, 1)
        return dx.clone()

    def _record_backward(self, module: nn.Module, plan: _Plan) -> torch.Tensor:
        st = _stream()
        saved = plan.saved
        n, cf, hf, wf = plan.dout_buf.shape
        # NCHW fp32 gradient of the features -> NHWC bf16 (the layout of every dz below)
        d = self._t(torch.empty((n, hf, wf, cf), dtype=torch.bfloat16, device=self.device))
        self._c(self.lib.esrp_nchw_f32_to_nhwc_bf16_affine, "esrp_nchw_f32_to_nhwc_bf16_affine", plan.dout_buf.data_ptr(), None, None,
                d.data_ptr(), n, cf, hf, wf, cf, st)
        dx = None
        for li in range(len(self.layers) - 1, -1, -1):
            L, rec = self.layers[li], saved[li]
            y = rec["y"]
            if self.pool_after[li]:
                dz = self._t(torch.empty_like(y))
                self._c(self.lib.esrp_maxpool2x2_relu_bwd_nhwc_bf16, "esrp_maxpool2x2_relu_bwd_nhwc_bf16", y.data_ptr(), d.data_ptr(),
                        dz.data_ptr(), n, y.shape[1], y.shape[2], L.cout, st)
            elif self.relu_after[li]:
                dz = self._t(torch.empty_like(y))
                self._c(self.lib.esrp_relu_bwd_nhwc_bf16, "esrp_relu_bwd_nhwc_bf16", y.data_ptr(), d.data_ptr(), dz.data_ptr(), y.numel(), st)
            else:
                dz = d
            if li == 0:
                dx = self._t(torch.empty((n, L.cin, dz.shape[1], dz.shape[2]), dtype=torch.float32, device=self.device))
                self._dgrad(L, dz, dx)
                break
            d = self._dgrad(L, dz, None)
        return dx


class _FeatureFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, eng, x):
        out, lease = eng.forward(module, x, lease=True)
        ctx.module, ctx.eng, ctx.lease = module, eng, lease
        return out

    @staticmethod
    def backward(ctx, dout):
        lease = ctx.lease
        if lease is None or lease.plan is None:
            raise RuntimeError("VGGFeatureExtractor: backward called twice on the same forward (saved state was released)")
        dx = ctx.eng.backward(ctx.module, lease.plan, dout.contiguous())
        lease.release()
        ctx.lease = None
        return None, None, dx


def feature_apply(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    if any(p.requires_grad for p in module.features.parameters()):
        raise NotImplementedError("VGGFeatureExtractor: the feature weights are frozen (architecture.py:300-301); "
                                  "gradients w.r.t. them are not produced")
    engines = module.__dict__.setdefault("_engines", {})
    eng = engines.get(x.device)
    if eng is None:
        eng = FeatureEngine(module, x.device)
        engines[x.device] = eng
    if torch.is_grad_enabled() and x.requires_grad:
        return _FeatureFn.apply(module, eng, x)
    return eng.forward(module, x)[0]

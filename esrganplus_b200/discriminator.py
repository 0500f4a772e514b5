"""Discriminator_VGG_128 forward on the sm_100a kernels (reference: architecture.py:87-129).

Host orchestration only: every convolution is one or more `esrp_conv3x3_nhwc` launches, BatchNorm2d /
LeakyReLU / space-to-depth / Linear are the small kernels of csrc/esrp_dnet.cu.  PyTorch owns memory and
does the O(C) BatchNorm bookkeeping on [C]-sized vectors (mean/var -> scale/shift, running statistics).

* 3x3 stride-1 convs (features.0/5/11/17/23) run directly.
* 4x4 stride-2 pad-1 convs (features.2/8/14/20/26) are rewritten as the 2x2 conv over the space-to-depth
  tensor S[n, Y, X, (a,b,ci)] = in[n, 2Y+a-1, 2X+b-1, ci]: a 3x3 conv whose taps with ky = 0 or kx = 0 carry
  zero weights, evaluated on the (H/2+1) x (W/2+1) grid of S (the last row / column is discarded).
* Output channels are cut in slices of 32 and K in groups of chunks whose weights fit in shared memory;
  later K groups accumulate onto the fp32 result of the earlier ones through the residual input.

Forward only (inference and the no-grad D passes); the backward pass is not built yet.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from . import conv as K

SLICE = 32            # output channels per launch
MAX_CHUNKS = {_lib.LAYOUT_ROW: 3, _lib.LAYOUT_TILE: 8}   # K chunks per launch (weights must fit in smem)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _rearrange_k4s2(w4: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 4, 4] -> [Cout, 4*Cin, 3, 3] for the space-to-depth formulation (see module docstring)."""
    cout, cin = w4.shape[:2]
    w3 = torch.zeros((cout, 4 * cin, 3, 3), dtype=w4.dtype, device=w4.device)
    for A in (0, 1):
        for Bk in (0, 1):
            for a in (0, 1):
                for b in (0, 1):
                    blk = (a * 2 + b) * cin
                    w3[:, blk:blk + cin, A + 1, Bk + 1] = w4[:, :, 2 * A + a, 2 * Bk + b]
    return w3


class _Layer:
    def __init__(self, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d]):
        self.conv, self.bn = conv, bn
        self.k, self.stride = conv.kernel_size[0], conv.stride[0]
        self.cin, self.cout = conv.in_channels, conv.out_channels
        if (self.k, self.stride) not in ((3, 1), (4, 2)) or self.cout % SLICE:
            raise NotImplementedError("Discriminator_VGG_128: only k3s1 / k4s2 convs with Cout % 32 == 0")
        self.packed: Dict[Tuple[int, int, int], torch.Tensor] = {}   # (layout, slice, kgroup) -> packed weights
        self.bias_pad: Optional[torch.Tensor] = None
        self.w3: Optional[torch.Tensor] = None
        self.sig = None


class DiscriminatorEngine:
    def __init__(self, module: nn.Module, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        feats = list(module.features)
        self.layers: List[_Layer] = []
        i = 0
        while i < len(feats):
            m = feats[i]
            if isinstance(m, nn.Conv2d):
                bn = feats[i + 1] if i + 1 < len(feats) and isinstance(feats[i + 1], nn.BatchNorm2d) else None
                self.layers.append(_Layer(m, bn))
            i += 1
        self.fc0, self.fc1 = module.classifier[0], module.classifier[2]

    # -- weights ---------------------------------------------------------------------------------
    def _sync(self, L: _Layer) -> None:
        w, b = L.conv.weight, L.conv.bias
        sig = (w.data_ptr(), w._version, b.data_ptr() if b is not None else 0, b._version if b is not None else 0)
        if sig == L.sig:
            return
        if w.device != self.device or w.dtype != torch.float32:
            raise RuntimeError(f"Discriminator_VGG_128: parameters must be fp32 on {self.device}")
        with torch.no_grad():
            w3 = w.detach() if L.k == 3 else _rearrange_k4s2(w.detach())
            cin_eff = w3.shape[1]
            L.kc = 64 if cin_eff % 64 == 0 else 32
            pad = (-cin_eff) % L.kc
            if pad:
                w3 = torch.cat([w3, torch.zeros((w3.shape[0], pad, 3, 3), device=w3.device)], 1)
            L.w3 = w3.contiguous()
            L.cin_eff = L.w3.shape[1]
            L.bias_pad = torch.zeros(L.cout, device=self.device) if b is None else b.detach().clone()
        L.packed.clear()
        L.sig = sig

    def _packed(self, L: _Layer, layout: int, s: int, g: int, chunks: List[int]) -> torch.Tensor:
        key = (layout, s, g)
        t = L.packed.get(key)
        if t is None:
            t = K.pack_conv3x3_weights(L.w3, L.kc, SLICE, chunks, row0=s * SLICE, rows=SLICE, layout=layout)
            L.packed[key] = t
        return t

    # -- one conv layer: NHWC bf16 [n,h,w,cin_pad] -> fp32 NHWC on the conv grid (or bf16 when fused) ------
    def _conv(self, L: _Layer, act: torch.Tensor, fuse_act_bf16: bool):
        n, h, w, c = act.shape
        if L.k == 4:
            src = torch.empty((n, h // 2 + 1, w // 2 + 1, 4 * c), dtype=torch.bfloat16, device=self.device)
            _lib.check(self.lib.esrp_s2d_pad_nhwc_bf16(act.data_ptr(), src.data_ptr(), n, h, w, c, _stream()), "s2d_pad")
            hv, wv = h // 2, w // 2
        else:
            src, hv, wv = act, h, w
        gh, gw = src.shape[1], src.shape[2]
        assert src.shape[3] == L.cin_eff, (src.shape, L.cin_eff)
        layout = _lib.LAYOUT_ROW if gw > 64 else _lib.LAYOUT_TILE
        nchunks = L.cin_eff // L.kc
        per = MAX_CHUNKS[layout]
        groups = [list(range(c0, min(c0 + per, nchunks))) for c0 in range(0, nchunks, per)]
        fused = fuse_act_bf16 and len(groups) == 1
        out_f = None if fused else torch.empty((n, gh, gw, L.cout), dtype=torch.float32, device=self.device)
        out_b = torch.empty((n, gh, gw, L.cout), dtype=torch.bfloat16, device=self.device) if fused else None
        for s in range(L.cout // SLICE):
            for g, chs in enumerate(groups):
                last = g == len(groups) - 1
                lc0 = [ch * L.kc for ch in chs]
                call = K.ConvCall(n=n, h=gh, w=gw, srcs=[src], kc=L.kc, chunks=[(0, c0) for c0 in lc0], bn=SLICE,
                                  cout=SLICE, w_packed=self._packed(L, layout, s, g, lc0), w_layout=layout,
                                  bias=L.bias_pad[s * SLICE:(s + 1) * SLICE] if last else None,
                                  act=1 if fused else 0)
                if fused:
                    call.out_bf16, call.ob_c0 = out_b, s * SLICE
                else:
                    call.out_f32, call.of_c0 = out_f, s * SLICE
                    if g > 0:
                        call.r1, call.r1_c0, call.s1 = out_f, s * SLICE, 1.0
                call.launch()
        return (out_b if fused else out_f), hv, wv

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, module: nn.Module, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.dtype != torch.float32 or x.device != self.device:
            raise RuntimeError("Discriminator_VGG_128 forward expects an fp32 NCHW CUDA tensor")
        n = x.shape[0]
        if x.shape[2] != 128 or x.shape[3] != 128:
            raise RuntimeError("Discriminator_VGG_128 expects 128x128 inputs (classifier is Linear(512*4*4, 100))")
        act = K.nchw_f32_to_nhwc_bf16(x, 32)
        flat = None
        for li, L in enumerate(self.layers):
            self._sync(L)
            is_last = li == len(self.layers) - 1
            y, hv, wv = self._conv(L, act, fuse_act_bf16=L.bn is None)
            if L.bn is None:
                act = y[:, :hv, :wv, :] if (y.shape[1] != hv or y.shape[2] != wv) else y
                act = act.contiguous()
                continue
            gh, gw, c = y.shape[1], y.shape[2], L.cout
            bn = L.bn
            count = n * hv * wv
            if bn.training:
                sums = torch.empty(2 * c, dtype=torch.float64, device=self.device)
                _lib.check(self.lib.esrp_bn_stats_nhwc_f32(y.data_ptr(), n, hv, wv, gh, gw, c, sums.data_ptr(), _stream()),
                           "bn_stats")
                with torch.no_grad():
                    mean = sums[:c] / count
                    var = (sums[c:] / count - mean * mean).clamp_min_(0)          # biased (block.py:32 -> nn.BatchNorm2d)
                    if bn.track_running_stats and bn.running_mean is not None:
                        m = bn.momentum if bn.momentum is not None else 0.1
                        bn.running_mean.mul_(1 - m).add_(mean.float(), alpha=m)
                        bn.running_var.mul_(1 - m).add_((var * (count / max(count - 1, 1))).float(), alpha=m)
                        bn.num_batches_tracked.add_(1)
            else:
                mean, var = bn.running_mean.double(), bn.running_var.double()
            with torch.no_grad():
                rstd = torch.rsqrt(var + bn.eps)
                scale = (bn.weight.detach().double() * rstd).float().contiguous()
                shift = (bn.bias.detach().double() - mean * bn.weight.detach().double() * rstd).float().contiguous()
            nxt = torch.empty((n, hv, wv, c), dtype=torch.bfloat16, device=self.device)
            if is_last:
                flat = torch.empty((n, c * hv * wv), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.esrp_bn_apply_nhwc(y.data_ptr(), n, hv, wv, gh, gw, c, scale.data_ptr(), shift.data_ptr(), 1,
                                                   nxt.data_ptr(), flat.data_ptr() if is_last else None, _stream()),
                       "bn_apply")
            act = nxt
        h0 = torch.empty((n, self.fc0.out_features), dtype=torch.float32, device=self.device)
        out = torch.empty((n, self.fc1.out_features), dtype=torch.float32, device=self.device)
        for fc, src, dst, a in ((self.fc0, flat, h0, 1), (self.fc1, h0, out, 0)):
            wt = fc.weight.detach().contiguous()
            _lib.check(self.lib.esrp_linear_f32(src.data_ptr(), wt.data_ptr(), fc.bias.data_ptr() if fc.bias is not None else None,
                                                dst.data_ptr(), n, fc.in_features, fc.out_features, a, _stream()), "linear")
        return out


def discriminator_apply(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters()))
    if needs_grad:
        raise NotImplementedError(
            "esrganplus_b200: the Discriminator_VGG_128 backward pass is not built yet; call it under "
            "torch.no_grad() (forward only)")
    engines = module.__dict__.setdefault("_engines", {})
    eng = engines.get(x.device)
    if eng is None:
        eng = DiscriminatorEngine(module, x.device)
        engines[x.device] = eng
    return eng.forward(module, x.contiguous())

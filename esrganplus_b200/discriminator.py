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

Backward (SRRaGAN_model.py:140 G phase: data gradient only, D frozen; :167 D phase: everything): per layer,
BatchNorm(train)+LeakyReLU backward as a reduction + an apply pass that writes dz (bf16) on the conv's output
grid, the weight gradient as conv3x3_wgrad units (32 input channels x 64 output channels x 9 taps; a 4x4 conv's
units are scattered through the space-to-depth index map), the data gradient as esrp_conv3x3_nhwc over
esrp_pack_dgrad_weights, and for the stride-2 layers the inverse space-to-depth rearrangement.  Parameter
gradients are views of ONE flat fp32 buffer (what the data-parallel all-reduce operates on).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import conv as K

_UNIT_DT = np.dtype([("x", "u8"), ("x_ctotal", "i4"), ("x_c0", "i4"), ("dy", "u8"), ("dy_ctotal", "i4"), ("dy_c0", "i4"),
                     ("acc", "u8"), ("bias_acc", "u8")], align=True)
_SCAT_DT = np.dtype([("acc", "u8"), ("dst", "u8"), ("dst_off", "i8"), ("dst_index", "i4"), ("kind", "i4"), ("col0", "i4"),
                     ("ncols", "i4"), ("nci", "i4"), ("co0", "i4"), ("ci0", "i4"), ("w_i", "i4"), ("scale", "f4")], align=True)
assert _UNIT_DT.itemsize == C.sizeof(_lib.WgradUnit) and _SCAT_DT.itemsize == C.sizeof(_lib.ScatterEntry)
ACC_BLOCK = 9 * 64 * 32

SLICE = 32            # output channels per launch
MAX_CHUNKS = {_lib.LAYOUT_ROW: 3, _lib.LAYOUT_TILE: _lib.ESRP_MAX_CHUNKS}   # K chunks per launch (the row kernel keeps its weights resident in shared memory, the tile kernel streams them)


_STREAM = [None]
import os as _os
_USE_GRAPHS = _os.environ.get("ESRP_D_GRAPH", "1") != "0"   # replay discriminator passes as CUDA graphs (0: call by call)


class _StreamArg(int):
    """A cudaStream_t value that recorded call lists can recognise (and re-target when a list is captured into a CUDA
    graph on the library's own capture stream: the legacy default stream torch normally runs on cannot be captured)."""


def _stream(refresh: bool = False) -> int:
    """cudaStream_t of torch's current stream.  torch.cuda.current_stream() costs ~14 us, a forward makes ~150 launches:
    the engine refreshes the cached handle once per forward / backward call."""
    if refresh:
        _STREAM[0] = _StreamArg(torch.cuda.current_stream().cuda_stream)
    return _STREAM[0]


def _named_params(module: nn.Module):
    """(names, parameters) of the module, cached: the module tree is fixed after construction, and walking 69 (D) or
    771 (G) parameters through named_parameters() on every call is milliseconds of host time per training step."""
    c = module.__dict__.get("_esrp_named_params")
    if c is None or len(c[1]) == 0 or c[1][0] is not next(iter(module.parameters())):
        names, params = [], []
        for k, v in module.named_parameters():
            names.append(k)
            params.append(v)
        c = (names, params)
        module.__dict__["_esrp_named_params"] = c
    return c


def _rearrange_k4s2(w4: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[Cout, Cin, 4, 4] -> [Cout, 4*Cin, 3, 3] for the space-to-depth formulation (see module docstring)."""
    cout, cin = w4.shape[:2]
    w3 = out if out is not None else torch.zeros((cout, 4 * cin, 3, 3), dtype=w4.dtype, device=w4.device)
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
        self.packed_t: Dict[Tuple[int, int], torch.Tensor] = {}      # (layout, input-channel slice) -> dgrad weights
        self.wg_tab = None                                           # cached wgrad unit / scatter tables
        self.repack: List = []                                       # recorded C-ABI calls (fn, args, keep-alive) that rebuild every packed tile in place
        self.bias_pad: Optional[torch.Tensor] = None
        self.w3: Optional[torch.Tensor] = None
        self.sig = None
        self.act_code = 1   # fused activation of the bf16 output: 1 LeakyReLU(0.2) (this network), 2 ReLU (vgg_feature.py)


class _Plan:
    """One recorded forward (and, on demand, backward) of the discriminator for a batch size: every buffer it touches is
    owned by the plan and every C-ABI call it makes is stored with its arguments, so that a later call of the same shape
    REPLAYS the list (~150 ctypes calls) instead of re-deriving shapes, allocating tensors and building descriptors.
    A plan is leased to one forward/backward pair at a time (the GAN step keeps up to three alive)."""

    def __init__(self, key, stream: int):
        self.key, self.stream = key, stream
        self.keep: list = []          # tensors / descriptors / numpy tables the recorded calls point into
        self.fwd: Optional[list] = None
        self.x_idx = -1               # index of the recorded call that reads the caller's input tensor
        self.out: Optional[torch.Tensor] = None
        self.saved: Optional[list] = None
        self.nbt: list = []           # BatchNorm num_batches_tracked buffers to bump per training forward
        self.bwd: Dict[tuple, tuple] = {}
        self.dout_buf: Optional[torch.Tensor] = None
        self.busy = False
        # CUDA-graph replay (esrp_graph_*): the recorded call list captured once, then one launch per pass
        self.x_in: Optional[torch.Tensor] = None      # plan-owned copy of the caller's input (its address is baked into the graph)
        self.fwd_graph = None                         # cudaGraphExec_t, or False when capture failed (plain replay then)
        self.bwd_graph: Dict[tuple, object] = {}

    def __del__(self):
        try:
            lib = _lib.load()
            for g in [self.fwd_graph] + list(self.bwd_graph.values()):
                if g:
                    lib.esrp_graph_destroy(g)
        except Exception:
            pass


class _Lease:
    def __init__(self, plan: _Plan):
        self.plan = plan
        plan.busy = True

    def release(self):
        if self.plan is not None:
            self.plan.busy = False
            self.plan = None

    def __del__(self):
        self.release()


class DiscriminatorEngine:
    def __init__(self, module: nn.Module, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        self.plans: Dict[tuple, List[_Plan]] = {}
        self.nside = 4
        self.side = (C.c_void_p * self.nside)()
        self.cap = (C.c_void_p * 1)()   # capture stream of the CUDA-graph replays
        with torch.cuda.device(device):
            _lib.check(self.lib.esrp_streams_create(self.nside, self.side), "esrp_streams_create")
            _lib.check(self.lib.esrp_streams_create(1, self.cap), "esrp_streams_create")
        self._rec: Optional[list] = None     # call list being recorded
        self._keep: Optional[list] = None    # keep-alive list of the plan being recorded
        self._ptr_sig = None
        self.epoch = 0
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

    # -- record / replay -------------------------------------------------------------------------
    def _c(self, fn, what: str, *args) -> None:
        """Call a C-ABI function now and, while a plan is being recorded, remember (fn, args) for replay."""
        _lib.check(fn(*args), what)
        if self._rec is not None:
            self._rec.append((fn, args, what))

    def _op(self, fn, *args) -> None:
        """A framework-level op (e.g. tensor.zero_) that must be repeated on replay."""
        fn(*args)
        if self._rec is not None:
            self._rec.append((fn, args, None))

    def _t(self, obj):
        """Keep `obj` alive as long as the plan being recorded (its address is baked into recorded calls)."""
        if self._keep is not None:
            self._keep.append(obj)
        return obj

    @staticmethod
    def _replay(calls) -> None:
        for fn, args, what in calls:
            rc = fn(*args)
            if what is not None and rc != 0:
                _lib.check(rc, what)

    def _capture(self, calls, stream: int):
        """The recorded call list as an executable CUDA graph (None if graphs are switched off, False if capture fails: the
        caller then keeps replaying call by call).  ~150 ctypes calls of 10-15 us each become one launch.  The list is
        replayed onto the engine's own capture stream (every `_StreamArg` argument re-targeted); the graph is then launched
        on the caller's stream."""
        if not _USE_GRAPHS:
            return None
        if any(what is None for _, _, what in calls):
            return False   # a framework-level op in the list cannot be re-targeted
        lib = self.lib
        cap = self.cap[0]
        if lib.esrp_graph_begin(cap) != 0:
            return False
        try:
            self._replay([(fn, tuple(cap if isinstance(a, _StreamArg) else a for a in args), what) for fn, args, what in calls])
        except Exception:
            lib.esrp_graph_abort(cap)
            return False
        exe = C.c_void_p()
        if lib.esrp_graph_end(cap, C.byref(exe)) != 0 or not exe.value:
            return False
        return exe

    def _acquire(self, module: nn.Module, n: int) -> _Plan:
        # recorded calls bake in the addresses of Parameters / buffers (Linear weights, BatchNorm vectors): a plan is
        # only valid while those stay where they are
        sig = tuple(t.data_ptr() for t in list(_named_params(module)[1]) + list(module.buffers()))
        if sig != self._ptr_sig:
            self.plans.clear()
            self._ptr_sig = sig
        st = _stream(refresh=True)
        key = (n, tuple(bool(L.bn.training) for L in self.layers if L.bn is not None), st)
        pool = self.plans.setdefault(key, [])
        for pl in pool:
            if not pl.busy:
                return pl
        if len(pool) >= 6:
            raise RuntimeError("Discriminator_VGG_128: more than 6 forward passes of one shape are waiting for their backward")
        pl = _Plan(key, st)
        pool.append(pl)
        return pl

    # -- weights ---------------------------------------------------------------------------------
    def _sync(self, L: _Layer) -> None:
        """Derived weight tensors follow the Parameter (pointer / in-place version).  They are rebuilt IN PLACE: recorded
        launch plans keep pointing at the same packed tiles, bias vector and 3x3 view of a 4x4 weight."""
        w, b = L.conv.weight, L.conv.bias
        sig = (w.data_ptr(), w._version, b.data_ptr() if b is not None else 0, b._version if b is not None else 0, getattr(self, "epoch", 0))
        if sig == L.sig:
            return
        if w.device != self.device or w.dtype != torch.float32:
            raise RuntimeError(f"Discriminator_VGG_128: parameters must be fp32 on {self.device}")
        with torch.no_grad():
            if L.w3 is None:
                w3 = w.detach().clone() if L.k == 3 else _rearrange_k4s2(w.detach())
                cin_eff = w3.shape[1]
                L.kc = 64 if cin_eff % 64 == 0 else 32
                pad = (-cin_eff) % L.kc
                if pad:
                    w3 = torch.cat([w3, torch.zeros((w3.shape[0], pad, 3, 3), device=w3.device)], 1)
                L.w3 = w3.contiguous()
                L.cin_eff = L.w3.shape[1]
                L.bias_pad = torch.zeros(L.cout, device=self.device)
            elif L.k == 3:
                L.w3[:, :L.cin].copy_(w.detach())
            else:
                _rearrange_k4s2(w.detach(), out=L.w3)
            if b is not None:
                L.bias_pad.copy_(b.detach())
        st = _stream()
        for fn, args, _keep in L.repack:   # raw C-ABI calls recorded when each packed tensor was first built
            rc = fn(*args, st)
            if rc != 0:
                _lib.check(rc, "weight repack")
        L.sig = sig

    def _packed(self, L: _Layer, layout: int, s: int, g: int, chunks: List[int], sl: int = SLICE) -> torch.Tensor:
        key = (layout, s, g, sl)
        t = L.packed.get(key)
        if t is None:
            K.RECORD = L.repack   # the call that builds the tile is the call that rebuilds it (same source / destination pointers)
            try:
                t = K.pack_conv3x3_weights(L.w3, L.kc, sl, chunks, row0=s * sl, rows=sl, layout=layout)
            finally:
                K.RECORD = None
            L.packed[key] = t
        return t

    def _fan_out(self, launches: int, pixels: int) -> bool:
        """Independent launches of one layer go to the side streams when each would leave most of the GPU idle."""
        return self.nside > 0 and launches >= 2 and pixels <= 128 * 100

    @staticmethod
    def _slice_width(channels: int, layout: int, chunks_per_launch: int) -> int:
        """Output channels per launch: 64 (UMMA N = 192) where the kernel family allows it — the tile kernel streams
        its weights, the row kernel keeps them resident and only fits one 64-wide chunk — else 32."""
        if channels % 64 == 0 and (layout == _lib.LAYOUT_TILE or chunks_per_launch == 1):
            return 64
        return 32

    # -- one conv layer: NHWC bf16 [n,h,w,cin_pad] -> fp32 NHWC on the conv grid (or bf16 when fused) ------
    def _conv(self, L: _Layer, act: torch.Tensor, fuse_act_bf16: bool, keep: Optional[list] = None):
        n, h, w, c = act.shape
        st = _stream()
        if L.k == 4:
            src = self._t(torch.empty((n, h // 2 + 1, w // 2 + 1, 4 * c), dtype=torch.bfloat16, device=self.device))
            self._c(self.lib.esrp_s2d_pad_nhwc_bf16, "s2d_pad", act.data_ptr(), src.data_ptr(), n, h, w, c, st)
            hv, wv = h // 2, w // 2
        else:
            src, hv, wv = act, h, w
        gh, gw = src.shape[1], src.shape[2]
        assert src.shape[3] == L.cin_eff, (src.shape, L.cin_eff)
        if keep is not None:
            keep.append(src)   # the conv's operand: what the weight gradient contracts against
        layout = _lib.LAYOUT_ROW if gw > 64 else _lib.LAYOUT_TILE
        nchunks = L.cin_eff // L.kc
        per = MAX_CHUNKS[layout]
        groups = [list(range(c0, min(c0 + per, nchunks))) for c0 in range(0, nchunks, per)]
        fused = fuse_act_bf16 and len(groups) == 1
        out_f = None if fused else self._t(torch.empty((n, gh, gw, L.cout), dtype=torch.float32, device=self.device))
        out_b = self._t(torch.empty((n, gh, gw, L.cout), dtype=torch.bfloat16, device=self.device)) if fused else None
        SLICE = self._slice_width(L.cout, layout, max(len(g_) for g_ in groups))
        fan = self._fan_out(L.cout // SLICE, n * gh * gw)
        # every weight tile exists BEFORE the fork: a tile built lazily inside the fan-out region is packed on the main
        # stream after the side streams have branched off, i.e. unordered with the launch that reads it
        tiles = [[self._packed(L, layout, s, g, [ch * L.kc for ch in chs], SLICE) for g, chs in enumerate(groups)]
                 for s in range(L.cout // SLICE)]
        if fan:
            self._c(self.lib.esrp_streams_fork, "esrp_streams_fork", st, self.side, self.nside)
        main_st = st
        for s in range(L.cout // SLICE):
            st = self.side[s % self.nside] if fan else main_st   # the K groups of a slice chain on one stream
            for g, chs in enumerate(groups):
                last = g == len(groups) - 1
                lc0 = [ch * L.kc for ch in chs]
                call = K.ConvCall(n=n, h=gh, w=gw, srcs=[src], kc=L.kc, chunks=[(0, c0) for c0 in lc0], bn=SLICE,
                                  cout=SLICE, w_packed=tiles[s][g], w_layout=layout,
                                  bias=L.bias_pad[s * SLICE:(s + 1) * SLICE] if last else None,
                                  act=L.act_code if fused else 0)
                if fused:
                    call.out_bf16, call.ob_c0 = out_b, s * SLICE
                else:
                    call.out_f32, call.of_c0 = out_f, s * SLICE
                    if g > 0:
                        call.r1, call.r1_c0, call.s1 = out_f, s * SLICE, 1.0
                d = self._t(call.desc())
                self._c(self.lib.esrp_conv3x3_nhwc, "esrp_conv3x3_nhwc", C.byref(d), st)
        if fan:
            self._c(self.lib.esrp_streams_join, "esrp_streams_join", main_st, self.side, self.nside)
        return (out_b if fused else out_f), hv, wv

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, module: nn.Module, x: torch.Tensor, lease: bool = False):
        """Returns (logits, lease or None).  With lease=True the plan (which holds everything the backward needs) stays
        reserved until the lease is released."""
        if x.dim() != 4 or x.dtype != torch.float32 or x.device != self.device:
            raise RuntimeError("Discriminator_VGG_128 forward expects an fp32 NCHW CUDA tensor")
        n = x.shape[0]
        if x.shape[2] != 128 or x.shape[3] != 128:
            raise RuntimeError("Discriminator_VGG_128 expects 128x128 inputs (classifier is Linear(512*4*4, 100))")
        x = x.contiguous()
        self.epoch = module.__dict__.get("_esrp_epoch", 0)   # bumped by apply / .to() / load_state_dict / invalidate_weights
        for L in self.layers:
            self._sync(L)
        plan = self._acquire(module, n)
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
            if plan.nbt:
                torch._foreach_add_(plan.nbt, 1)
        return plan.out.clone(), (_Lease(plan) if lease else None)

    def _record_forward(self, module: nn.Module, x: torch.Tensor, plan: _Plan) -> None:
        st = _stream()
        n = x.shape[0]
        saved: list = []
        act = self._t(torch.empty((n, 128, 128, 32), dtype=torch.bfloat16, device=self.device))
        plan.x_idx = len(self._rec)
        self._c(self.lib.esrp_nchw_f32_to_nhwc_bf16, "esrp_nchw_f32_to_nhwc_bf16", x.data_ptr(), act.data_ptr(), n, x.shape[1], 128, 128,
                32, st)
        flat = None
        for li, L in enumerate(self.layers):
            is_last = li == len(self.layers) - 1
            y, hv, wv = self._conv(L, act, fuse_act_bf16=L.bn is None, keep=saved)
            rec = dict(src=saved.pop(), in_shape=tuple(act.shape), y=None, hv=hv, wv=wv)
            saved.append(rec)
            if L.bn is None:
                assert y.shape[1] == hv and y.shape[2] == wv
                act = y
                rec["out_act"] = act
                continue
            gh, gw, c = y.shape[1], y.shape[2], L.cout
            bn = L.bn
            count = n * hv * wv
            coef = self._t(torch.empty((7, c), dtype=torch.float32, device=self.device))
            batch_stats = bool(bn.training or bn.running_mean is None)
            sums = None
            if batch_stats:
                sums = self._t(torch.empty(2 * c, dtype=torch.float64, device=self.device))
                self._c(self.lib.esrp_bn_stats_nhwc_f32, "bn_stats", y.data_ptr(), n, hv, wv, gh, gw, c, sums.data_ptr(), st)
            track = batch_stats and bn.track_running_stats and bn.running_mean is not None
            m = bn.momentum if bn.momentum is not None else 0.1
            self._c(self.lib.esrp_bn_finalize, "bn_finalize", sums.data_ptr() if sums is not None else None, float(count),
                    bn.weight.data_ptr() if bn.weight is not None else None,
                    bn.bias.data_ptr() if bn.bias is not None else None, float(bn.eps), float(m), int(batch_stats),
                    bn.running_mean.data_ptr() if (track or not batch_stats) else None,
                    bn.running_var.data_ptr() if (track or not batch_stats) else None, c, coef.data_ptr(), st)
            if track:
                bn.num_batches_tracked.add_(1)
                plan.nbt.append(bn.num_batches_tracked)
            rec.update(y=y, coef=coef, count=count, batch_stats=batch_stats)
            nxt = self._t(torch.empty((n, hv, wv, c), dtype=torch.bfloat16, device=self.device))
            if is_last:
                flat = self._t(torch.empty((n, c * hv * wv), dtype=torch.float32, device=self.device))
            self._c(self.lib.esrp_bn_apply_nhwc, "bn_apply", y.data_ptr(), n, hv, wv, gh, gw, c, coef[2].data_ptr(), coef[3].data_ptr(),
                    1, nxt.data_ptr(), flat.data_ptr() if is_last else None, st)
            act = nxt
        h0 = self._t(torch.empty((n, self.fc0.out_features), dtype=torch.float32, device=self.device))
        out = self._t(torch.empty((n, self.fc1.out_features), dtype=torch.float32, device=self.device))
        for fc, src, dst, a in ((self.fc0, flat, h0, 1), (self.fc1, h0, out, 0)):
            if not fc.weight.is_contiguous():
                raise RuntimeError("Discriminator_VGG_128: classifier weights must be contiguous")
            self._c(self.lib.esrp_linear_f32, "linear", src.data_ptr(), fc.weight.data_ptr(),
                    fc.bias.data_ptr() if fc.bias is not None else None, dst.data_ptr(), n, fc.in_features, fc.out_features, a, st)
        saved.append(dict(flat=flat, h0=h0, n=n))
        plan.saved = saved
        plan.out = out

    # -- backward --------------------------------------------------------------------------------
    def _dgrad_packed(self, L: _Layer, layout: int, s: int, kc: int, sl: int = SLICE) -> torch.Tensor:
        key = (layout, s, sl)
        t = L.packed_t.get(key)
        if t is None:
            groups = [(L.w3, 32 * g, 1.0) for g in range(L.cout // 32)]
            K.RECORD = L.repack
            try:
                t = K.pack_dgrad_weights(groups, s * sl, min(sl, L.cin_eff - s * sl), kc, sl, layout=layout)
            finally:
                K.RECORD = None
            L.packed_t[key] = t
        return t

    def _wgrad_tables(self, L: _Layer):
        """Static part of the layer's wgrad unit table and scatter table (pointers are patched per call)."""
        if L.wg_tab is None:
            ng, nblk = L.cin_eff // 32, L.cout // 64
            units = np.zeros(ng * nblk, dtype=_UNIT_DT)
            scat = np.zeros(ng * nblk + nblk, dtype=_SCAT_DT)
            real_cin = L.cin
            i = 0
            for g in range(ng):
                for b in range(nblk):
                    units[i]["x_ctotal"], units[i]["x_c0"] = L.cin_eff, 32 * g
                    units[i]["dy_ctotal"], units[i]["dy_c0"] = L.cout, 64 * b
                    e = scat[i]
                    e["dst_index"], e["col0"], e["ncols"], e["co0"], e["ci0"], e["w_i"], e["scale"] = -1, 0, 64, 64 * b, 32 * g, real_cin, 1.0
                    if L.k == 3:
                        e["kind"] = 0
                        e["nci"] = max(0, min(32, real_cin - 32 * g))
                    else:
                        e["kind"], e["nci"] = 2, 32
                    i += 1
            for b in range(nblk):
                e = scat[ng * nblk + b]
                e["dst_index"], e["kind"], e["ncols"], e["scale"], e["dst_off"] = -1, 1, 64, 1.0, 64 * b
            L.wg_tab = (units, scat, ng, nblk)
        return L.wg_tab

    def _wgrad(self, L: _Layer, src: torch.Tensor, dz: torch.Tensor, dw: torch.Tensor, db: torch.Tensor) -> None:
        units0, scat0, ng, nblk = self._wgrad_tables(L)
        n, gh, gw, _ = dz.shape
        nu = ng * nblk
        acc = self._t(torch.empty(nu * ACC_BLOCK + nblk * 64, dtype=torch.float32, device=self.device))
        self._c(self.lib.esrp_memset_zero, "esrp_memset_zero", acc.data_ptr(), acc.numel() * acc.element_size(), _stream())
        units, scat = units0.copy(), scat0.copy()
        accp = acc.data_ptr()
        bias_base = accp + 4 * nu * ACC_BLOCK
        idx = np.arange(nu, dtype=np.uint64)
        units["x"], units["dy"] = src.data_ptr(), dz.data_ptr()
        units["acc"] = np.uint64(accp) + idx * np.uint64(4 * ACC_BLOCK)
        blk = (idx % np.uint64(nblk))
        units["bias_acc"] = np.where(idx < np.uint64(nblk), np.uint64(bias_base) + blk * np.uint64(256), np.uint64(0))
        scat["acc"][:nu] = units["acc"]
        scat["dst"][:nu] = dw.data_ptr()
        scat["acc"][nu:] = np.uint64(bias_base) + np.arange(nblk, dtype=np.uint64) * np.uint64(256)
        scat["dst"][nu:] = db.data_ptr()
        st = _stream()
        # one call covers whole 64-channel input chunks (2 groups x nblk units each); a launch holds <= 64 (chunk, slab) jobs
        slabs = (nblk + 1) // 2
        per_call = max(1, min(_lib.WGRAD_MAX_UNITS // (2 * nblk), 64 // slabs)) * 2 * nblk
        for u0 in range(0, nu, per_call):
            cnt = min(per_call, nu - u0)
            part = self._t(np.ascontiguousarray(units[u0:u0 + cnt]))
            self._c(self.lib.esrp_conv3x3_wgrad, "esrp_conv3x3_wgrad", part.ctypes.data_as(C.POINTER(_lib.WgradUnit)), cnt, n, gh, gw,
                    0, st)
        if L.k == 3 and L.cin % 32:
            self._c(self.lib.esrp_memset_zero, "esrp_memset_zero", dw.data_ptr(), dw.numel() * dw.element_size(), _stream())  # (never the case for D_VGG_128 beyond layer 0, whose 3 real channels are all written)
        self._t(scat)
        self._c(self.lib.esrp_wgrad_scatter, "esrp_wgrad_scatter", scat.ctypes.data_as(C.POINTER(_lib.ScatterEntry)), len(scat), st)

    def _dgrad(self, L: _Layer, dz: torch.Tensor, out_nchw: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """dz [n,gh,gw,cout] bf16 -> gradient of the conv's (space-to-depth) input [n,gh,gw,cin_eff] bf16, or for the
        first layer straight into the NCHW fp32 gradient of the image."""
        n, gh, gw, _ = dz.shape
        layout = _lib.LAYOUT_ROW if gw > 64 else _lib.LAYOUT_TILE
        kc = 64 if L.cout % 64 == 0 else 32
        nchunks = L.cout // kc
        if nchunks > MAX_CHUNKS[layout]:
            raise NotImplementedError("Discriminator_VGG_128 backward: K does not fit one launch for this layer width")
        chunks = [(0, c * kc) for c in range(nchunks)]
        st = _stream()
        if out_nchw is not None:
            wp = L.packed_t.get((layout, -1, 16))
            if wp is None:
                groups = [(L.w3, 32 * g, 1.0) for g in range(L.cout // 32)]
                K.RECORD = L.repack
                try:
                    wp = K.pack_dgrad_weights(groups, 0, L.cin, kc, 16, layout=layout)
                finally:
                    K.RECORD = None
                L.packed_t[(layout, -1, 16)] = wp
            d = self._t(K.ConvCall(n=n, h=gh, w=gw, srcs=[dz], kc=kc, chunks=chunks, bn=16, cout=L.cin, w_packed=wp, w_layout=layout,
                                   out_nchw=out_nchw).desc())
            self._c(self.lib.esrp_conv3x3_nhwc, "esrp_conv3x3_nhwc", C.byref(d), st)
            return None
        out = self._t(torch.empty((n, gh, gw, L.cin_eff), dtype=torch.bfloat16, device=self.device))
        SLICE = self._slice_width(L.cin_eff, layout, nchunks)
        fan = self._fan_out(L.cin_eff // SLICE, n * gh * gw)
        tiles = [self._dgrad_packed(L, layout, s, kc, SLICE) for s in range(L.cin_eff // SLICE)]   # before the fork (see _conv)
        if fan:
            self._c(self.lib.esrp_streams_fork, "esrp_streams_fork", st, self.side, self.nside)
        for s in range(L.cin_eff // SLICE):
            d = self._t(K.ConvCall(n=n, h=gh, w=gw, srcs=[dz], kc=kc, chunks=chunks, bn=SLICE, cout=SLICE,
                                   w_packed=tiles[s], w_layout=layout, out_bf16=out,
                                   ob_c0=s * SLICE).desc())
            self._c(self.lib.esrp_conv3x3_nhwc, "esrp_conv3x3_nhwc", C.byref(d), self.side[s % self.nside] if fan else st)
        if fan:
            self._c(self.lib.esrp_streams_join, "esrp_streams_join", st, self.side, self.nside)
        return out

    def backward(self, module: nn.Module, plan: _Plan, dout: torch.Tensor, need_dx: bool, need_params: bool):
        """Returns (dx NCHW fp32 or None, [per-parameter gradient views or None], flat gradient buffer or None); the
        gradients are views of a fresh copy of the plan's flat buffer, in named_parameters() order."""
        names, plist = _named_params(module)
        n = plan.saved[-1]["n"]
        if _stream(refresh=True) != plan.stream:
            raise RuntimeError("Discriminator_VGG_128 backward must run on the CUDA stream its forward ran on (the plan's "
                               "buffers and recorded launches belong to that stream)")
        if plan.dout_buf is None:
            plan.dout_buf = torch.empty((n, self.fc1.out_features), dtype=torch.float32, device=self.device)
        plan.dout_buf.copy_(dout)
        key = (bool(need_dx), bool(need_params))
        ent = plan.bwd.get(key)
        if ent is None:
            self._rec, self._keep = [], plan.keep
            try:
                dx, flatg, sizes = self._record_backward(module, plan, need_dx, need_params)
                ent = plan.bwd[key] = (self._rec, dx, flatg, sizes)
            finally:
                self._rec = self._keep = None
        else:
            gr = plan.bwd_graph.get(key)
            if gr is None:
                gr = plan.bwd_graph[key] = self._capture(ent[0], plan.stream)
            if gr:
                _lib.check(self.lib.esrp_graph_launch(gr, plan.stream), "esrp_graph_launch")
            else:
                self._replay(ent[0])
        _, dx, flatg, sizes = ent
        grads, flat = [None] * len(names), None
        if need_params:
            flat = flatg.clone()
            grads = [part[:p.numel()].view(p.shape) if part.numel() != p.numel() else part.view(p.shape)
                     for part, p in zip(flat.split(sizes), plist)]
        return (dx.clone() if dx is not None else None), grads, flat

    def _record_backward(self, module: nn.Module, plan: _Plan, need_dx: bool, need_params: bool):
        st = _stream()
        saved = plan.saved
        head = saved[-1]
        n, flat, h0 = head["n"], head["flat"], head["h0"]
        names, plist = _named_params(module)
        params = dict(zip(names, plist))
        grads: Dict[str, torch.Tensor] = {}
        flatg, sizes = None, []
        if need_params:
            offs, off = {}, 0
            for k in names:
                offs[k] = off
                sizes.append((params[k].numel() + 3) // 4 * 4)
                off += sizes[-1]
            flatg = self._t(torch.empty(off, dtype=torch.float32, device=self.device))
            grads = {k: flatg[offs[k]:offs[k] + params[k].numel()].view(params[k].shape) for k in names}
        gp = lambda k: grads[k].data_ptr() if need_params else None
        dout = plan.dout_buf
        # classifier (architecture.py:122-123)
        dh0 = self._t(torch.empty_like(h0))
        self._c(self.lib.esrp_linear_bwd_f32, "linear_bwd", dout.data_ptr(), None, h0.data_ptr(), self.fc1.weight.data_ptr(),
                dh0.data_ptr(), gp("classifier.2.weight"), gp("classifier.2.bias"), n, self.fc1.in_features,
                self.fc1.out_features, st)
        dflat = self._t(torch.empty_like(flat))
        self._c(self.lib.esrp_linear_bwd_f32, "linear_bwd", dh0.data_ptr(), h0.data_ptr(), flat.data_ptr(), self.fc0.weight.data_ptr(),
                dflat.data_ptr(), gp("classifier.0.weight"), gp("classifier.0.bias"), n, self.fc0.in_features,
                self.fc0.out_features, st)
        feats = list(module.features)
        idx_of = {id(m): i for i, m in enumerate(feats)}
        dout_b, dout_nchw = None, dflat
        dx = None
        for li in range(len(self.layers) - 1, -1, -1):
            L, rec = self.layers[li], saved[li]
            ci = idx_of[id(L.conv)]
            hv, wv, c = rec["hv"], rec["wv"], L.cout
            if L.bn is not None:
                y = rec["y"]
                gh, gw = y.shape[1], y.shape[2]
                bi = idx_of[id(L.bn)]
                coef = rec["coef"]   # rows a, b are (re)written by bn_bwd_finalize on every backward
                sums = self._t(torch.empty(2 * c, dtype=torch.float64, device=self.device))
                db_, dn_ = (dout_b.data_ptr() if dout_b is not None else None), (dout_nchw.data_ptr() if dout_nchw is not None else None)
                self._c(self.lib.esrp_bn_bwd_reduce, "bn_bwd_reduce", y.data_ptr(), db_, dn_, n, hv, wv, gh, gw, c, coef.data_ptr(),
                        sums.data_ptr(), st)
                self._c(self.lib.esrp_bn_bwd_finalize, "bn_bwd_finalize", sums.data_ptr(), float(rec["count"]), int(rec["batch_stats"]), c,
                        coef.data_ptr(), gp(f"features.{bi}.weight"), gp(f"features.{bi}.bias"), st)
                dz = self._t(torch.empty((n, gh, gw, c), dtype=torch.bfloat16, device=self.device))
                self._c(self.lib.esrp_bn_bwd_apply, "bn_bwd_apply", y.data_ptr(), db_, dn_, n, hv, wv, gh, gw, c, coef.data_ptr(),
                        dz.data_ptr(), st)
            else:
                dz = dout_b  # already times lrelu' (s2d_pad_bwd with the forward activation as sign reference)
            if need_params:
                self._wgrad(L, rec["src"], dz, grads[f"features.{ci}.weight"], grads[f"features.{ci}.bias"])
            if li == 0:
                if need_dx:
                    dx = self._t(torch.empty((n, L.cin, dz.shape[1], dz.shape[2]), dtype=torch.float32, device=self.device))
                    self._dgrad(L, dz, dx)
                break
            dsrc = self._dgrad(L, dz, None)
            if L.k == 4:
                _, ih, iw, ic = rec["in_shape"]
                din = self._t(torch.empty((n, ih, iw, ic), dtype=torch.bfloat16, device=self.device))
                prev = self.layers[li - 1]
                ref = saved[li - 1]["out_act"] if prev.bn is None else None
                self._c(self.lib.esrp_s2d_pad_bwd_nhwc_bf16, "s2d_pad_bwd", dsrc.data_ptr(), din.data_ptr(),
                        ref.data_ptr() if ref is not None else None, n, ih, iw, ic, st)
            else:
                if self.layers[li - 1].bn is None:
                    raise NotImplementedError("Discriminator_VGG_128 backward: an un-normalised layer must feed a stride-2 layer")
                din = dsrc
            dout_b, dout_nchw = din, None
        return dx, flatg, sizes


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, eng, x, *params):
        out, lease = eng.forward(module, x, lease=True)
        ctx.module, ctx.eng, ctx.lease = module, eng, lease
        return out

    @staticmethod
    def backward(ctx, dout):
        need_dx = ctx.needs_input_grad[2]
        needs = ctx.needs_input_grad[3:]
        need_params = any(needs)
        lease = ctx.lease
        if lease is None or lease.plan is None:
            raise RuntimeError("Discriminator_VGG_128: backward called twice on the same forward (saved state was released)")
        dx, grads, flat = ctx.eng.backward(ctx.module, lease.plan, dout, need_dx, need_params)
        lease.release()
        ctx.lease = None
        if flat is not None:
            from .autograd import allreduce_flat
            allreduce_flat(ctx.module, flat)
        return (None, None, dx) + tuple(g if nd else None for g, nd in zip(grads, needs))


def discriminator_apply(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    plist = _named_params(module)[1]
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in plist))
    engines = module.__dict__.setdefault("_engines", {})
    eng = engines.get(x.device)
    if eng is None:
        eng = DiscriminatorEngine(module, x.device)
        engines[x.device] = eng
    if needs_grad:
        return _DiscriminatorFn.apply(module, eng, x, *plist)
    return eng.forward(module, x)[0]

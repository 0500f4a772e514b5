"""Host-side helpers over the C ABI for one fused conv launch (PyTorch is plumbing only: it owns
device memory and the stream; all arithmetic happens in libesrp.so)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import Conv3x3Desc


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


# When a list is installed here, the weight-pack wrappers below append the raw C-ABI call they make as
# (function, arguments without the trailing stream, keep-alive objects): a caller that repacks the same tensors after every
# optimizer step (the discriminator engine) replays those calls instead of re-deriving them in Python (~55 us each).
RECORD: Optional[list] = None


def _i32_array(vals: Sequence[int]):
    arr = (C.c_int32 * len(vals))(*vals)
    return arr


def pack_conv3x3_weights(w_oihw: torch.Tensor, kc: int, bn: int, chunk_lc0: Sequence[int],
                         w_aux: Optional[torch.Tensor] = None, aux_chunks: int = 0, row0: int = 0,
                         rows: Optional[int] = None, transpose: bool = False,
                         layout: int = _lib.LAYOUT_TILE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[O, I, 3, 3] fp32 (device) -> pre-swizzled bf16 UMMA B tiles [chunk][ky][kx*bn + r (, conv1x1)][kc]
    (opaque bytes).  `w_aux` is the bias-free 1x1 conv [O, I_aux(,1,1)] fused on the first `aux_chunks`
    chunks; `transpose` packs the data-gradient operator (see include/esrp.h)."""
    lib = _lib.load()
    assert w_oihw.is_cuda and w_oihw.dtype == torch.float32 and w_oihw.dim() == 4
    w = w_oihw.contiguous()
    w_o, w_i, kh, kw = w.shape
    assert kh == 3 and kw == 3
    if rows is None:
        rows = min(bn, (w_i if transpose else w_o) - row0)
    aux_ptr, aux_cin = None, 0
    if w_aux is not None:
        assert w_aux.is_cuda and w_aux.dtype == torch.float32
        w_aux = w_aux.reshape(w_aux.shape[0], w_aux.shape[1]).contiguous()
        aux_ptr, aux_cin = w_aux.data_ptr(), w_aux.shape[1]
    nbytes = lib.esrp_packed_conv3x3_bytes(len(chunk_lc0), kc, bn, int(aux_chunks > 0))
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    assert out.numel() == nbytes and out.dtype == torch.uint8   # `out=`: repack in place (cached descriptors keep their pointer)
    lc0 = _i32_array(chunk_lc0)
    args = (w.data_ptr(), w_o, w_i, int(transpose), layout, row0, rows, kc, bn, len(chunk_lc0), lc0, aux_ptr, aux_cin, aux_chunks,
            out.data_ptr())
    _lib.check(lib.esrp_pack_conv3x3_weights(*args, _stream_ptr()), "esrp_pack_conv3x3_weights")
    if RECORD is not None:
        RECORD.append((lib.esrp_pack_conv3x3_weights, args, (w, w_aux, out, lc0)))
    return out


@dataclass
class ConvCall:
    """Python-side description of one esrp_conv3x3_nhwc launch."""

    n: int
    h: int
    w: int
    srcs: List[torch.Tensor]                 # 1 or 2 NHWC bf16 tensors
    kc: int
    chunks: List[tuple]                      # [(src_index, c0), ...]
    bn: int
    cout: int
    w_packed: torch.Tensor
    w_layout: int = _lib.LAYOUT_TILE         # must match the layout w_packed was packed with
    bias: Optional[torch.Tensor] = None      # fp32 [bn]
    aux_chunks: int = 0                      # leading chunks feeding the fused conv1x1 rows of w_packed
    act: int = 0
    s0: float = 1.0
    r1: Optional[torch.Tensor] = None
    r1_c0: int = 0
    s1: float = 1.0
    r2: Optional[torch.Tensor] = None
    r2_c0: int = 0
    s2: float = 1.0
    noise: int = 0
    noise_ctotal: int = 0
    noise_c0: int = 0
    sigma: float = 0.1
    seed: int = 0
    offset: int = 0
    out_bf16: Optional[torch.Tensor] = None
    ob_c0: int = 0
    out_f32: Optional[torch.Tensor] = None
    of_c0: int = 0
    out_nchw: Optional[torch.Tensor] = None
    variant: int = 0
    trace: Optional[torch.Tensor] = None     # int64 [3*1024] device tensor (CTA 0 timeline)
    # training extensions (include/esrp.h): masks are int16 tensors [n,h,w,bits/16]
    mask_out: Optional[torch.Tensor] = None
    mask_out_c0: int = 0
    mask_in: Optional[torch.Tensor] = None
    mask_in_c0: int = 0
    r2_pre: int = 0
    pre_bf16: Optional[torch.Tensor] = None
    pb_c0: int = 0
    pre_f32: Optional[torch.Tensor] = None
    pf_c0: int = 0
    # co-scheduled output slices (row kernel): slice s uses w_packed / bias + s*slice_stride bytes, channels + s*bn
    slices: int = 0
    slice_stride: int = 0
    f32_planar: int = 0                      # fp32 r1 / r2 / out_f32 are [n,h,c/4,w,4] in memory (include/esrp.h)
    k_valid: int = 0                         # input channels that carry non-zero weights (0 = all)
    ob_lo_c0: int = -1                       # >= 0: also store bf16(v - bf16(v)) at this channel offset of out_bf16 (tile kernel)
    _keep: list = field(default_factory=list, repr=False)

    def desc(self) -> Conv3x3Desc:
        d = Conv3x3Desc()
        d.n, d.h, d.w = self.n, self.h, self.w
        for i, s in enumerate(self.srcs):
            assert s.is_cuda and s.dtype == torch.bfloat16 and s.is_contiguous()
            assert s.shape[:3] == (self.n, self.h, self.w), (s.shape, (self.n, self.h, self.w))
            d.src[i] = s.data_ptr()
            d.src_ctotal[i] = s.shape[3]
        d.kc = self.kc
        d.num_chunks = len(self.chunks)
        for i, (si, c0) in enumerate(self.chunks):
            d.chunk_src[i] = si
            d.chunk_c0[i] = c0
        d.aux_chunks = self.aux_chunks
        d.bn, d.cout = self.bn, self.cout
        d.w_packed = self.w_packed.data_ptr()
        d.w_layout = self.w_layout
        d.bias = self.bias.data_ptr() if self.bias is not None else None
        d.act, d.s0 = self.act, self.s0
        for name in ("r1", "r2"):
            t = getattr(self, name)
            if t is not None:
                assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.bfloat16, torch.float32)
                setattr(d, name, t.data_ptr())
                setattr(d, name + "_is_f32", int(t.dtype == torch.float32))
                setattr(d, name + "_ctotal", t.shape[3])
                setattr(d, name + "_c0", getattr(self, name + "_c0"))
        d.s1, d.s2 = self.s1, self.s2
        d.noise, d.sigma, d.seed, d.offset = self.noise, self.sigma, self.seed, self.offset
        d.noise_ctotal, d.noise_c0 = (self.noise_ctotal or self.cout), self.noise_c0
        if self.out_bf16 is not None:
            assert self.out_bf16.dtype == torch.bfloat16 and self.out_bf16.is_contiguous()
            d.out_bf16, d.ob_ctotal, d.ob_c0 = self.out_bf16.data_ptr(), self.out_bf16.shape[3], self.ob_c0
        if self.out_f32 is not None:
            assert self.out_f32.dtype == torch.float32 and self.out_f32.is_contiguous()
            d.out_f32, d.of_ctotal, d.of_c0 = self.out_f32.data_ptr(), self.out_f32.shape[3], self.of_c0
        if self.out_nchw is not None:
            assert self.out_nchw.dtype == torch.float32 and self.out_nchw.is_contiguous()
            d.out_nchw = self.out_nchw.data_ptr()
        d.variant = self.variant
        d.trace = self.trace.data_ptr() if self.trace is not None else None
        for name in ("mask_out", "mask_in"):
            t = getattr(self, name)
            if t is not None:
                assert t.is_cuda and t.dtype == torch.int16 and t.is_contiguous() and t.shape[:3] == (self.n, self.h, self.w)
                setattr(d, name, t.data_ptr())
                setattr(d, name + "_ctotal", t.shape[3] * 16)
                setattr(d, name + "_c0", getattr(self, name + "_c0"))
        d.r2_pre = self.r2_pre
        if self.pre_bf16 is not None:
            assert self.pre_bf16.dtype == torch.bfloat16 and self.pre_bf16.is_contiguous()
            d.pre_bf16, d.pb_ctotal, d.pb_c0 = self.pre_bf16.data_ptr(), self.pre_bf16.shape[3], self.pb_c0
        if self.pre_f32 is not None:
            assert self.pre_f32.dtype == torch.float32 and self.pre_f32.is_contiguous()
            d.pre_f32, d.pf_ctotal, d.pf_c0 = self.pre_f32.data_ptr(), self.pre_f32.shape[3], self.pf_c0
        d.slices, d.slice_stride = self.slices, self.slice_stride
        d.f32_planar = self.f32_planar
        d.k_valid = self.k_valid
        d.out_lo, d.ob_lo_c0 = (1, self.ob_lo_c0) if self.ob_lo_c0 >= 0 else (0, 0)
        return d

    def launch(self) -> None:
        lib = _lib.load()
        d = self.desc()
        _lib.check(lib.esrp_conv3x3_nhwc(C.byref(d), _stream_ptr()), "esrp_conv3x3_nhwc")


def nchw_f32_to_nhwc_bf16(x: torch.Tensor, c_pad: int) -> torch.Tensor:
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float32
    x = x.contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.esrp_nchw_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), n, c, h, w, c_pad, _stream_ptr()),
               "esrp_nchw_f32_to_nhwc_bf16")
    return out


def nhwc_bf16_to_nchw_f32(x: torch.Tensor, c: int) -> torch.Tensor:
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    n, h, w, ct = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _lib.check(lib.esrp_nhwc_bf16_to_nchw_f32(x.data_ptr(), out.data_ptr(), n, c, h, w, ct, _stream_ptr()),
               "esrp_nhwc_bf16_to_nchw_f32")
    return out


def upsample2x_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.esrp_upsample2x_nhwc_bf16(x.data_ptr(), out.data_ptr(), n, h, w, c, _stream_ptr()),
               "esrp_upsample2x_nhwc_bf16")
    return out


# ------------------------------------------------------------------------------------------------
# backward-pass helpers (esrp_pack_dgrad_weights / esrp_conv3x3_wgrad / ... in include/esrp.h)
# ------------------------------------------------------------------------------------------------
def pack_dgrad_weights(groups: Sequence[Optional[tuple]], row0: int, rows: int, kc: int, bn: int,
                       layout: int = _lib.LAYOUT_TILE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """groups: per 32-channel K group either None (zero weights) or (w_oihw fp32 cuda, co0, scale)."""
    lib = _lib.load()
    arr = (_lib.DgradGroup * len(groups))()
    keep = []
    dev = None
    for i, g in enumerate(groups):
        if g is None:
            continue
        w, co0, scale = g
        assert w.is_cuda and w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3)
        w = w.contiguous()
        keep.append(w)
        dev = w.device
        arr[i].w, arr[i].w_o, arr[i].w_i, arr[i].co0, arr[i].scale = w.data_ptr(), w.shape[0], w.shape[1], co0, scale
    chunks = (len(groups) * 32 + kc - 1) // kc
    if out is None:
        out = torch.empty(lib.esrp_packed_conv3x3_bytes(chunks, kc, bn, 0), dtype=torch.uint8, device=dev)
    args = (arr, len(groups), layout, row0, rows, kc, bn, out.data_ptr())
    _lib.check(lib.esrp_pack_dgrad_weights(*args, _stream_ptr()), "esrp_pack_dgrad_weights")
    if RECORD is not None:
        RECORD.append((lib.esrp_pack_dgrad_weights, args, (keep, out, arr)))
    return out


def conv3x3_wgrad(units: Sequence[tuple], n: int, h: int, w: int, splits: int = 0) -> None:
    """units: (x bf16 NHWC, x_c0, dy bf16 NHWC, dy_c0, acc fp32 [9,64,32], bias_acc fp32 [64] or None)."""
    lib = _lib.load()
    arr = (_lib.WgradUnit * len(units))()
    for i, (x, x_c0, dy, dy_c0, acc, bacc) in enumerate(units):
        assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
        assert acc.dtype == torch.float32 and acc.is_contiguous() and acc.numel() == 9 * 64 * 32
        arr[i].x, arr[i].x_ctotal, arr[i].x_c0 = x.data_ptr(), x.shape[3], x_c0
        arr[i].dy, arr[i].dy_ctotal, arr[i].dy_c0 = dy.data_ptr(), dy.shape[3], dy_c0
        arr[i].acc = acc.data_ptr()
        arr[i].bias_acc = bacc.data_ptr() if bacc is not None else None
    _lib.check(lib.esrp_conv3x3_wgrad(arr, len(units), n, h, w, splits, _stream_ptr()), "esrp_conv3x3_wgrad")


def wgrad_scatter(entries: Sequence[dict]) -> None:
    lib = _lib.load()
    arr = (_lib.ScatterEntry * len(entries))()
    for i, e in enumerate(entries):
        arr[i].acc = e["acc"].data_ptr() + 4 * e.get("acc_off", 0)
        arr[i].dst = e["dst"].data_ptr()
        arr[i].dst_off = e.get("dst_off", 0)
        arr[i].dst_index = -1
        for k in ("kind", "col0", "ncols", "nci", "co0", "ci0", "w_i"):
            setattr(arr[i], k, e.get(k, 0))
        arr[i].scale = e.get("scale", 1.0)
    _lib.check(lib.esrp_wgrad_scatter(arr, len(entries), _stream_ptr()), "esrp_wgrad_scatter")


def conv1x1_bwd(x: torch.Tensor, dx2: torch.Tensor, d_c0: int, u: torch.Tensor, g: torch.Tensor,
                extra: Optional[torch.Tensor] = None, du_acc: Optional[torch.Tensor] = None) -> None:
    lib = _lib.load()
    nf = g.shape[-1]
    npx = g.numel() // nf
    _lib.check(lib.esrp_conv1x1_bwd(nf, x.data_ptr(), x.shape[-1], dx2.data_ptr(), dx2.shape[-1], d_c0, u.data_ptr(),
                                    g.data_ptr(), extra.data_ptr() if extra is not None else None,
                                    du_acc.data_ptr() if du_acc is not None else None, npx, _stream_ptr()),
               "esrp_conv1x1_bwd")


def upsample2x_bwd(dup: torch.Tensor, mask: Optional[torch.Tensor] = None, mask_c0: int = 0, want_f32: bool = False):
    lib = _lib.load()
    n, h2, w2, c = dup.shape
    h, w = h2 // 2, w2 // 2
    out_b = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dup.device)
    out_f = torch.empty((n, h, w, c), dtype=torch.float32, device=dup.device) if want_f32 else None
    _lib.check(lib.esrp_upsample2x_bwd_nhwc_bf16(dup.data_ptr(), n, h, w, c, mask.data_ptr() if mask is not None else None,
                                                 mask.shape[3] * 16 if mask is not None else 0, mask_c0, out_b.data_ptr(),
                                                 out_f.data_ptr() if out_f is not None else None, _stream_ptr()),
               "esrp_upsample2x_bwd_nhwc_bf16")
    return out_b, out_f

"""Parity of the backward-pass kernels (through the C ABI) with torch autograd on the same operands.

Tolerances: operands are bf16, accumulation fp32.  Data gradients follow the forward-conv contract
(2e-3 * max|ref| for fp32 outputs, 1e-2 for bf16 outputs); weight gradients sum n*h*w bf16 products in
fp32 with atomics (order-dependent), compared at 2e-3 * max|ref| against torch fp32 conv autograd run
on the same bf16-rounded operands.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from esrganplus_b200 import _lib
from esrganplus_b200 import conv as K

pytestmark = pytest.mark.gpu

SHAPES = [(2, 20, 27), (1, 16, 16), (3, 5, 7), (1, 33, 130), (2, 32, 32), (1, 64, 128)]


def _bits_from(t_nhwc_bool: torch.Tensor) -> torch.Tensor:
    """[n,h,w,c] bool -> int16 words [n,h,w,c/16], bit i of word j = channel 16 j + i."""
    n, h, w, c = t_nhwc_bool.shape
    b = t_nhwc_bool.reshape(n, h, w, c // 16, 16).to(torch.int32)
    weights = (2 ** torch.arange(16, device=b.device, dtype=torch.int32))
    words = (b * weights).sum(-1)
    words = torch.where(words >= 32768, words - 65536, words)
    return words.to(torch.int16).contiguous()


@pytest.mark.parametrize("layout", [_lib.LAYOUT_TILE, _lib.LAYOUT_ROW])
@pytest.mark.parametrize("shape", SHAPES[:5])
def test_mask_out_and_masked_dgrad(cuda_dev, layout, shape):
    """Forward conv saves the LeakyReLU selector bits; the data gradient of [conv_a | conv_b] (two source
    convs concatenated along K, one scaled by 0.2) applies them, with a pre-mask bf16/fp32 copy, an fp32
    residual added before the copy (r2_pre) and a final scale."""
    n, h, w = shape
    g = torch.Generator(device=cuda_dev).manual_seed(7 + h * w)
    x = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(32, 64, 3, 3, device=cuda_dev, generator=g) / 24.0
    bias = torch.randn(32, device=cuda_dev, generator=g) * 0.1
    wp = K.pack_conv3x3_weights(wt, 64, 32, [0], layout=layout)
    out = torch.zeros(n, h, w, 32, device=cuda_dev, dtype=torch.bfloat16)
    mask = torch.zeros(n, h, w, 8, device=cuda_dev, dtype=torch.int16)  # 128 bits per pixel
    K.ConvCall(n=n, h=h, w=w, srcs=[x], kc=64, chunks=[(0, 0)], bn=32, cout=32, w_packed=wp, w_layout=layout, bias=bias,
               act=1, out_bf16=out, mask_out=mask, mask_out_c0=32).launch()
    pre = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, padding=1)
    want_bits = _bits_from((pre > 0).permute(0, 2, 3, 1))
    # values within rounding distance of zero may legitimately differ; compare where |pre| is not tiny
    sure = (pre.abs() > 1e-3).permute(0, 2, 3, 1)
    got = mask[..., 2:4]
    got_b = torch.stack([(got[..., j].to(torch.int32) >> i) & 1 for j in range(2) for i in range(16)], -1).bool()
    want_b = (pre > 0).permute(0, 2, 3, 1)
    assert (got_b == want_b)[sure].all()
    assert (mask[..., :2] == 0).all() and (mask[..., 4:] == 0).all()
    assert want_bits.shape == got.shape

    # dgrad over DC = [dA (64 ch, conv A: 64 -> 96, scale 0.2) | dB (32 ch, conv B: 32 -> 96) | 32 unused]
    wa = torch.randn(64, 96, 3, 3, device=cuda_dev, generator=g) / 24.0
    wb = torch.randn(32, 96, 3, 3, device=cuda_dev, generator=g) / 17.0
    dc = torch.randn(n, h, w, 128, device=cuda_dev, generator=g).to(torch.bfloat16)
    res = torch.randn(n, h, w, 32, device=cuda_dev, generator=g)
    row0 = 64  # input-channel slice [64, 96) of both convs
    wpd = K.pack_dgrad_weights([(wa, 0, 0.2), (wa, 32, 0.2), (wb, 0, 1.0), None], row0, 32, 64, 32, layout=layout)
    o_b = torch.zeros(n, h, w, 32, device=cuda_dev, dtype=torch.bfloat16)
    o_f = torch.zeros(n, h, w, 32, device=cuda_dev)
    p_b = torch.zeros(n, h, w, 32, device=cuda_dev, dtype=torch.bfloat16)
    p_f = torch.zeros(n, h, w, 32, device=cuda_dev)
    K.ConvCall(n=n, h=h, w=w, srcs=[dc], kc=64, chunks=[(0, 0), (0, 64)], bn=32, cout=32, w_packed=wpd, w_layout=layout,
               r2=res, r2_pre=1, s2=0.5, pre_bf16=p_b, pre_f32=p_f, mask_in=mask, mask_in_c0=32, out_bf16=o_b,
               out_f32=o_f).launch()
    dA = dc[..., :64].float().permute(0, 3, 1, 2)
    dB = dc[..., 64:96].float().permute(0, 3, 1, 2)
    full = (F.conv_transpose2d(dA, (0.2 * wa).to(torch.bfloat16).float(), padding=1)[:, 64:96]
            + F.conv_transpose2d(dB, wb.to(torch.bfloat16).float(), padding=1)[:, 64:96]) + res.permute(0, 3, 1, 2)
    slope = torch.where(got_b, 1.0, 0.2).permute(0, 3, 1, 2)
    fin = 0.5 * full * slope
    sc = max(1.0, full.abs().max().item())
    assert (p_f.permute(0, 3, 1, 2) - full).abs().max().item() <= 2e-3 * sc
    assert (p_b.float().permute(0, 3, 1, 2) - full).abs().max().item() <= 1e-2 * sc
    assert (o_f.permute(0, 3, 1, 2) - fin).abs().max().item() <= 2e-3 * sc
    assert (o_b.float().permute(0, 3, 1, 2) - fin).abs().max().item() <= 1e-2 * sc


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_wgrad_units_match_autograd(cuda_dev, shape, splits):
    """dW and db of a 96 -> 64 conv from three 32-channel units (one of them reading a second tensor at a
    channel offset), scattered into OIHW, against torch autograd."""
    n, h, w = shape
    g = torch.Generator(device=cuda_dev).manual_seed(11 + h + w + splits)
    xa = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    xb = torch.randn(n, h, w, 128, device=cuda_dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(n, h, w, 192, device=cuda_dev, generator=g).to(torch.bfloat16)
    acc = torch.zeros(3, 9, 64, 32, device=cuda_dev)
    bacc = torch.zeros(64, device=cuda_dev)
    units = [(xa, 0, dy, 64, acc[0], bacc), (xa, 32, dy, 64, acc[1], None), (xb, 96, dy, 64, acc[2], None)]
    K.conv3x3_wgrad(units, n, h, w, splits)
    dw = torch.full((64, 96, 3, 3), float("nan"), device=cuda_dev)
    db = torch.full((64,), float("nan"), device=cuda_dev)
    ents = [dict(acc=acc[i], dst=dw, kind=0, col0=0, ncols=64, nci=32, co0=0, ci0=32 * i, w_i=96, scale=1.0) for i in range(3)]
    ents.append(dict(acc=bacc, dst=db, kind=1, ncols=64, scale=1.0))
    K.wgrad_scatter(ents)
    xcat = torch.cat([xa.float(), xb[..., 96:128].float()], 3).permute(0, 3, 1, 2).contiguous()
    wt = torch.zeros(64, 96, 3, 3, device=cuda_dev, requires_grad=True)
    bt = torch.zeros(64, device=cuda_dev, requires_grad=True)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = F.conv2d(xcat, wt, bt, padding=1)
    y.backward(dy[..., 64:128].float().permute(0, 3, 1, 2).contiguous())
    sc = max(1.0, wt.grad.abs().max().item())
    assert torch.isfinite(dw).all() and torch.isfinite(db).all()
    assert (dw - wt.grad).abs().max().item() <= 2e-3 * sc, (dw - wt.grad).abs().max().item() / sc
    assert (db - bt.grad).abs().max().item() <= 2e-3 * max(1.0, bt.grad.abs().max().item())


def test_wgrad_dy_block_hanging_over_channel_count(cuda_dev):
    """A 64-column dY block that starts 32 channels before the end of the tensor reads zeros beyond it."""
    n, h, w = 1, 12, 20
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    x = torch.randn(n, h, w, 32, device=cuda_dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(n, h, w, 32, device=cuda_dev, generator=g).to(torch.bfloat16)
    acc = torch.zeros(9, 64, 32, device=cuda_dev)
    K.conv3x3_wgrad([(x, 0, dy, 0, acc, None)], n, h, w, 0)
    assert (acc[:, 32:] == 0).all()
    wt = torch.zeros(32, 32, 3, 3, device=cuda_dev, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    got = acc[:, :32].permute(1, 2, 0).reshape(32, 32, 3, 3)
    assert (got - wt.grad).abs().max().item() <= 2e-3 * max(1.0, wt.grad.abs().max().item())


@pytest.mark.parametrize("nf", [64, 32])
def test_conv1x1_bwd(cuda_dev, nf):
    n, h, w = 2, 9, 13
    g = torch.Generator(device=cuda_dev).manual_seed(nf)
    x = torch.randn(n, h, w, nf, device=cuda_dev, generator=g).to(torch.bfloat16)
    dx2 = torch.randn(n, h, w, 96, device=cuda_dev, generator=g).to(torch.bfloat16)
    u = torch.randn(32, nf, device=cuda_dev, generator=g) / nf ** 0.5
    gbuf = torch.randn(n, h, w, nf, device=cuda_dev, generator=g)
    extra = torch.randn(n, h, w, nf, device=cuda_dev, generator=g)
    g0 = gbuf.clone()
    du = torch.zeros(32, nf, device=cuda_dev)
    K.conv1x1_bwd(x, dx2, 32, u, gbuf, extra, du)
    d = dx2[..., 32:64].float()
    want_g = g0 + d @ u.to(torch.bfloat16).float() + extra      # the kernel multiplies by the bf16 U the forward uses
    want_du = d.reshape(-1, 32).t() @ x.float().reshape(-1, nf)
    assert (gbuf - want_g).abs().max().item() <= 1e-4 * max(1.0, want_g.abs().max().item())
    assert (du - want_du).abs().max().item() <= 1e-4 * max(1.0, want_du.abs().max().item())


def test_upsample2x_bwd(cuda_dev):
    n, h, w, c = 2, 7, 9, 64
    g = torch.Generator(device=cuda_dev).manual_seed(1)
    dup = torch.randn(n, 2 * h, 2 * w, c, device=cuda_dev, generator=g).to(torch.bfloat16)
    bits = torch.rand(n, h, w, c, device=cuda_dev, generator=g) > 0.5
    mask = _bits_from(bits)
    ob, of = K.upsample2x_bwd(dup, mask, 0, want_f32=True)
    s = dup.float().reshape(n, h, 2, w, 2, c).sum((2, 4))
    assert (of - s).abs().max().item() <= 1e-5
    want = s * torch.where(bits, 1.0, 0.2)
    assert (ob.float() - want).abs().max().item() <= 1e-2 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 16), (128, 128, 8)])
def test_k4s2_layer_gradients_match_conv2d_autograd(cuda_dev, cin, cout, hw):
    """A 4x4 / stride-2 / pad-1 conv (architecture.py:95-119) runs as a 3x3 conv over the space-to-depth tensor:
    weight gradient = conv3x3_wgrad units scattered through the (a, b, tap) -> 4x4 index map (scatter kind 2),
    data gradient = transposed 3x3 conv + inverse space-to-depth; both against torch conv2d(stride=2) autograd on the
    same bf16 operands."""
    import torch.nn as nn
    from esrganplus_b200.discriminator import DiscriminatorEngine, _Layer
    n = 2
    g = torch.Generator(device=cuda_dev).manual_seed(cin + hw)
    conv = nn.Conv2d(cin, cout, 4, 2, 1).to(cuda_dev)
    eng = DiscriminatorEngine.__new__(DiscriminatorEngine)
    eng.lib, eng.device = _lib.load(), cuda_dev
    eng._rec = eng._keep = None   # no plan is being recorded: calls run directly
    eng.nside, eng.side = 0, None  # ... on the current stream
    L = _Layer(conv, None)
    eng._sync(L)
    act = torch.randn(n, hw, hw, cin, device=cuda_dev, generator=g).to(torch.bfloat16)
    src = torch.empty(n, hw // 2 + 1, hw // 2 + 1, 4 * cin, device=cuda_dev, dtype=torch.bfloat16)
    assert eng.lib.esrp_s2d_pad_nhwc_bf16(act.data_ptr(), src.data_ptr(), n, hw, hw, cin, torch.cuda.current_stream().cuda_stream) == 0
    dz = torch.zeros(n, hw // 2 + 1, hw // 2 + 1, cout, device=cuda_dev, dtype=torch.bfloat16)
    dz[:, :hw // 2, :hw // 2] = torch.randn(n, hw // 2, hw // 2, cout, device=cuda_dev, generator=g).to(torch.bfloat16)
    dw = torch.full_like(conv.weight, float("nan"))
    db = torch.full_like(conv.bias, float("nan"))
    eng._wgrad(L, src, dz, dw, db)
    dsrc = eng._dgrad(L, dz, None)
    din = torch.empty(n, hw, hw, cin, device=cuda_dev, dtype=torch.bfloat16)
    assert eng.lib.esrp_s2d_pad_bwd_nhwc_bf16(dsrc.data_ptr(), din.data_ptr(), None, n, hw, hw, cin,
                                              torch.cuda.current_stream().cuda_stream) == 0
    torch.backends.cudnn.allow_tf32 = False
    xr = act.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = conv.weight.detach().to(torch.bfloat16).float().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True)
    y = F.conv2d(xr, wr, br, stride=2, padding=1)
    y.backward(dz[:, :hw // 2, :hw // 2].float().permute(0, 3, 1, 2).contiguous())
    assert torch.isfinite(dw).all() and torch.isfinite(db).all()
    assert (dw - wr.grad).abs().max().item() <= 2e-3 * max(1.0, wr.grad.abs().max().item())
    assert (db - br.grad).abs().max().item() <= 2e-3 * max(1.0, br.grad.abs().max().item())
    got = din.float().permute(0, 3, 1, 2)
    assert (got - xr.grad).abs().max().item() <= 1e-2 * max(1.0, xr.grad.abs().max().item())

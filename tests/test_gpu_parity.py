"""Parity of the sm_100a path (through the C ABI in libesrp.so) with the oracle and with the golden
fixtures generated from the reference.

Stated tolerances (DESIGN.md "Precision contract"): operands are rounded to bf16, products are
accumulated in fp32 in tensor memory, the residual trunk is carried in fp32.
  * one conv vs torch fp32 conv2d on the SAME bf16-rounded operands: fp32 outputs within
    2e-3 * max|ref| (accumulation order only), bf16 outputs within 1e-2 * max|ref| (one rounding).
  * whole networks vs the fp32 reference: max|d| <= 6e-2 * std(ref) and PSNR >= 48 dB on the
    clamped [0,1] image (utils/util.py:107-114); the 8-bit PNG quantisation floor is 58.9 dB.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import esrganplus_b200 as E
from esrganplus_b200 import _lib
from esrganplus_b200 import conv as K
from oracle import esrgan_oracle as O

pytestmark = pytest.mark.gpu

NET_REL_TOL = 6e-2
NET_PSNR_DB = 48.0


def _golden(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


def _net_close(y: torch.Tensor, ref: torch.Tensor, what=""):
    assert y.shape == ref.shape, (y.shape, ref.shape)
    assert torch.isfinite(y).all(), what
    err = (y - ref).abs().max().item()
    rel = err / ref.std().item()
    psnr = O.psnr_255(y, ref)
    assert rel <= NET_REL_TOL and psnr >= NET_PSNR_DB, f"{what}: max|d|={err:.3e} rel={rel:.3e} psnr={psnr:.1f}"
    return rel, psnr


def test_extension_is_loaded_and_device_is_blackwell(cuda_dev):
    lib = _lib.load()
    assert lib.esrp_sm_count() > 0
    major, _ = torch.cuda.get_device_capability(cuda_dev)
    assert major == 10, "kernels are built for sm_100a only"


# ---------------------------------------------------------------------------------------------------
# one conv launch
# ---------------------------------------------------------------------------------------------------
def _ref_conv(srcs, chunks, kc, w, bias, act):
    x = torch.cat([srcs[si][..., c0:c0 + kc] for si, c0 in chunks], dim=3).float().permute(0, 3, 1, 2)
    y = F.conv2d(x.contiguous(), w.to(torch.bfloat16).float(), bias, padding=1)
    return F.leaky_relu(y, 0.2) if act else y


# (n, h, w), variant: default tiling plus forced M-tile widths / accumulator-slot counts, so that the
# x-halo column blocks (w > cw), every lane mapping (cw = 16..128) and partial tiles are all exercised
TILINGS = [((2, 20, 27), 0), ((1, 16, 16), 0), ((3, 5, 7), 0), ((1, 33, 130), 0), ((1, 9, 300), 0),
           ((2, 20, 27), _lib.variant_cwlog2(4)), ((2, 20, 27), _lib.variant_cwlog2(7) | _lib.variant_mt(2)),
           ((1, 40, 64), _lib.variant_cwlog2(6) | _lib.variant_mt(1)), ((1, 40, 64), _lib.variant_cwlog2(5)),
           ((2, 37, 128), _lib.variant_mt(3)), ((1, 128, 128), 0)]


# row-streaming kernel: ragged widths/heights, multi-column images, single-row images, many segments per CTA
ROW_TILINGS = [((2, 20, 27), 0), ((1, 16, 16), 0), ((3, 5, 7), 0), ((1, 33, 130), 0), ((1, 9, 300), 0),
               ((2, 37, 128), _lib.variant_mt(4)), ((1, 128, 128), 0), ((2, 1, 140), 0), ((1, 2, 64), _lib.variant_mt(4)),
               ((6, 40, 128), 0), ((3, 70, 260), 0)]
ALL_TILINGS = [(_lib.LAYOUT_TILE, s, v) for s, v in TILINGS] + [(_lib.LAYOUT_ROW, s, v) for s, v in ROW_TILINGS]


@pytest.mark.parametrize("kc,bn", [(64, 32), (64, 64), (32, 32), (32, 64), (64, 16), (32, 16)])
@pytest.mark.parametrize("layout,shape,variant", ALL_TILINGS)
def test_conv3x3_plain(cuda_dev, kc, bn, layout, variant, shape):
    torch.backends.cudnn.allow_tf32 = False
    n, h, w = shape
    if bn == 64 and layout == _lib.LAYOUT_TILE and (variant & 15) > 2:
        pytest.skip("TMEM: needs more accumulator slots of 192 columns than fit")
    g = torch.Generator(device=cuda_dev).manual_seed(kc * 1000 + bn * 10 + variant + h)
    s0 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    s1 = torch.randn(n, h, w, 128, device=cuda_dev, generator=g).to(torch.bfloat16)
    chunks = [(0, 0), (1, 64)] if kc == 64 else [(0, 0), (0, 32), (1, 32)]
    cin = kc * len(chunks)
    cout = bn if bn >= 32 else 3
    wt = torch.randn(cout, cin, 3, 3, device=cuda_dev, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, device=cuda_dev, generator=g)
    wp = K.pack_conv3x3_weights(wt, kc, bn, [i * kc for i in range(len(chunks))], layout=layout)
    bias_p = torch.zeros(bn, device=cuda_dev)
    bias_p[:cout] = bias
    ref = _ref_conv([s0, s1], chunks, kc, wt, bias, act=1)
    scale = max(1.0, ref.abs().max().item())
    if cout < 16:  # Cout=3 tail conv (HR_conv1): NCHW fp32 output straight to the caller's tensor
        out = torch.full((n, cout, h, w), float("nan"), device=cuda_dev)
        K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout, w_packed=wp, w_layout=layout,
                   bias=bias_p, act=1, out_nchw=out, variant=variant).launch()
        assert (out - ref).abs().max().item() <= 2e-3 * scale
        return
    out = torch.full((n, h, w, 192), float("nan"), device=cuda_dev, dtype=torch.bfloat16)
    K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout, w_packed=wp, w_layout=layout,
               bias=bias_p, act=1, out_bf16=out, ob_c0=64, variant=variant).launch()
    got = out[..., 64:64 + cout].float().permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() <= 1e-2 * scale
    # concat-free write: only the addressed channel slice is touched
    assert torch.isnan(out[..., :64].float()).all() and torch.isnan(out[..., 64 + cout:].float()).all()


@pytest.mark.parametrize("layout", [_lib.LAYOUT_TILE, _lib.LAYOUT_ROW])
@pytest.mark.parametrize("kc,bn", [(64, 32), (32, 32), (64, 64), (32, 16)])
@pytest.mark.parametrize("shape,variant", [((2, 20, 27), 0), ((1, 24, 128), 0), ((1, 6, 200), 0)])
def test_conv3x3_fused_epilogue(cuda_dev, kc, bn, shape, variant, layout):
    """bias + LeakyReLU + s0 + conv1x1 aux + fp32 residual + bf16 RRDB residual, both output twins
    (block.py:262-268, 291)."""
    n, h, w = shape
    if bn == 64 and layout == _lib.LAYOUT_ROW:
        pytest.skip("row layout: bn <= 32")
    g = torch.Generator(device=cuda_dev).manual_seed(99 + kc + bn + w)
    s0 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    s1 = torch.randn(n, h, w, 128, device=cuda_dev, generator=g).to(torch.bfloat16)
    chunks = [(0, 0), (1, 64)] if kc == 64 else [(0, 0), (0, 32), (1, 32)]
    cin, cout = kc * len(chunks), bn
    wt = torch.randn(cout, cin, 3, 3, device=cuda_dev, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, device=cuda_dev, generator=g)
    wa = torch.randn(cout, kc, 1, 1, device=cuda_dev, generator=g) / kc ** 0.5
    wp = K.pack_conv3x3_weights(wt, kc, bn, [i * kc for i in range(len(chunks))], w_aux=wa, aux_chunks=1, layout=layout)
    r1 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g)
    r2 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    out_b = torch.zeros((n, h, w, 64), device=cuda_dev, dtype=torch.bfloat16)
    out_f = torch.zeros((n, h, w, 64), device=cuda_dev)
    K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout, w_packed=wp, w_layout=layout, bias=bias,
               act=1, s0=0.5, aux_chunks=1, r1=r1, s1=0.25, r2=r2, s2=0.2, out_bf16=out_b,
               out_f32=out_f, variant=variant).launch()
    ref = _ref_conv([s0, s1], chunks, kc, wt, bias, act=1)
    aux = F.conv2d(s0[..., :kc].float().permute(0, 3, 1, 2).contiguous(), wa.to(torch.bfloat16).float())
    v = 0.5 * ref + aux + 0.25 * r1[..., :cout].permute(0, 3, 1, 2)
    v = 0.2 * v + r2[..., :cout].float().permute(0, 3, 1, 2)
    scale = max(1.0, v.abs().max().item())
    assert (out_f[..., :cout].permute(0, 3, 1, 2) - v).abs().max().item() <= 2e-3 * scale
    assert (out_b[..., :cout].float().permute(0, 3, 1, 2) - v).abs().max().item() <= 1e-2 * scale


@pytest.mark.parametrize("layout", [_lib.LAYOUT_TILE, _lib.LAYOUT_ROW])
def test_conv3x3_output_channel_slices_and_transposed_pack(cuda_dev, layout):
    """A 64-output conv issued as two 32-channel launches (how conv5 runs), and the data-gradient
    operator packed from the same OIHW tensor (transpose=True) against conv_transpose2d."""
    n, h, w = 1, 18, 40
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    s0 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(64, 64, 3, 3, device=cuda_dev, generator=g) / 24.0
    out = torch.zeros((n, h, w, 64), device=cuda_dev)
    for half in (0, 1):
        wp = K.pack_conv3x3_weights(wt, 32, 32, [0, 32], row0=32 * half, rows=32, layout=layout)
        K.ConvCall(n=n, h=h, w=w, srcs=[s0], kc=32, chunks=[(0, 0), (0, 32)], bn=32, cout=32, w_packed=wp,
                   w_layout=layout, out_f32=out, of_c0=32 * half).launch()
    x = s0.float().permute(0, 3, 1, 2).contiguous()
    wq = wt.to(torch.bfloat16).float()
    ref = F.conv2d(x, wq, padding=1)
    assert (out.permute(0, 3, 1, 2) - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    # dgrad: dX = conv_transpose2d(dY, W) == conv(dY, W'[ci][co] flipped)
    out2 = torch.zeros((n, h, w, 64), device=cuda_dev)
    for half in (0, 1):
        wpt = K.pack_conv3x3_weights(wt, 64, 32, [0], transpose=True, row0=32 * half, rows=32, layout=layout)
        K.ConvCall(n=n, h=h, w=w, srcs=[s0], kc=64, chunks=[(0, 0)], bn=32, cout=32, w_packed=wpt, w_layout=layout,
                   out_f32=out2, of_c0=32 * half).launch()
    ref2 = F.conv_transpose2d(x, wq, padding=1)
    assert (out2.permute(0, 3, 1, 2) - ref2).abs().max().item() <= 2e-3 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("shape", [(2, 20, 200), (3, 37, 130), (1, 1, 128), (16, 128, 128)])
@pytest.mark.parametrize("train_ext", [False, True])
@pytest.mark.parametrize("issuers", [0, _lib.VARIANT_ROW_ALT])   # both MMA-issuer protocols of the row kernel
def test_conv3x3_coscheduled_slices(cuda_dev, shape, train_ext, issuers):
    """conv5 of a dense block (block.py:258,268; RRDB skip block.py:291) as ONE launch over two 32-channel output
    slices (esrp_conv3x3_t::slices): same arithmetic as two launches, so the results must be bit-identical to them,
    and within the single-conv tolerance of torch conv2d."""
    n, h, w = shape
    g = torch.Generator(device=cuda_dev).manual_seed(1234 + w)
    t_in = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    gro = torch.randn(n, h, w, 128, device=cuda_dev, generator=g).to(torch.bfloat16)
    chunks = [(0, 0), (1, 0), (1, 64)]
    wt = torch.randn(64, 192, 3, 3, device=cuda_dev, generator=g) / (192 * 9) ** 0.5
    bias = torch.randn(64, device=cuda_dev, generator=g)
    lib = _lib.load()
    nbytes = lib.esrp_packed_conv3x3_bytes(3, 64, 32, 0)
    w_al = (nbytes + 1023) // 1024 * 1024
    stride = w_al + 1024
    buf = torch.zeros(2 * stride, dtype=torch.uint8, device=cuda_dev)
    biases = []
    for sl in (0, 1):
        K.pack_conv3x3_weights(wt, 64, 32, [0, 64, 128], row0=32 * sl, rows=32, layout=_lib.LAYOUT_ROW,
                               out=buf[sl * stride: sl * stride + nbytes])
        b = buf[sl * stride + w_al: sl * stride + w_al + 128].view(torch.float32)
        b.copy_(bias[32 * sl: 32 * sl + 32])
        biases.append(b)
    r1 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g)
    r2 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g)

    def run(sliced):
        ob = torch.zeros((n, h, w, 64), device=cuda_dev, dtype=torch.bfloat16)
        of = torch.zeros((n, h, w, 64), device=cuda_dev)
        pre = torch.zeros((n, h, w, 64), device=cuda_dev) if train_ext else None
        common = dict(n=n, h=h, w=w, srcs=[t_in, gro], kc=64, chunks=chunks, bn=32, cout=32, w_layout=_lib.LAYOUT_ROW,
                      s0=0.2, r1=r1, s1=1.0, r2=r2, s2=0.2, out_bf16=ob, out_f32=of, noise=1, noise_ctotal=64, seed=77,
                      offset=5 << 36, pre_f32=pre, variant=issuers)
        if sliced:
            K.ConvCall(w_packed=buf[:nbytes], bias=biases[0], slices=2, slice_stride=stride, **common).launch()
        else:
            for sl in (0, 1):
                K.ConvCall(w_packed=buf[sl * stride: sl * stride + nbytes], bias=biases[sl], r1_c0=32 * sl, r2_c0=32 * sl,
                           ob_c0=32 * sl, of_c0=32 * sl, noise_c0=32 * sl, pf_c0=32 * sl, **common).launch()
        return ob, of, pre

    ob1, of1, pre1 = run(True)
    ob2, of2, pre2 = run(False)
    assert torch.equal(of1, of2) and torch.equal(ob1, ob2)
    if not train_ext:
        # fp32 operands in the engine-private [n,h,c/4,w,4] layout (esrp_conv3x3_t::f32_planar): same values
        def to_planar(t):
            return t.view(n, h, w, 16, 4).permute(0, 1, 3, 2, 4).contiguous().view(n, h, w, 64)

        def from_planar(t):
            return t.view(n, h, 16, w, 4).permute(0, 1, 3, 2, 4).contiguous().view(n, h, w, 64)

        ob3 = torch.zeros_like(ob1)
        of3 = torch.zeros_like(of1)
        K.ConvCall(n=n, h=h, w=w, srcs=[t_in, gro], kc=64, chunks=chunks, bn=32, cout=32, w_layout=_lib.LAYOUT_ROW, s0=0.2,
                   r1=to_planar(r1), s1=1.0, r2=to_planar(r2), s2=0.2, out_bf16=ob3, out_f32=of3, noise=1, noise_ctotal=64,
                   seed=77, offset=5 << 36, w_packed=buf[:nbytes], bias=biases[0], slices=2, slice_stride=stride,
                   f32_planar=1, variant=issuers).launch()
        assert torch.equal(ob3, ob1) and torch.equal(from_planar(of3), of1)
    if train_ext:
        assert torch.equal(pre1, pre2)
    # against torch: v = 0.2*conv + r1; noise; 0.2*v + r2  (draws regenerated on the host)
    ref = _ref_conv([t_in, gro], chunks, 64, wt, bias, act=0)
    v = (0.2 * ref + r1.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous()
    if train_ext:
        scale = max(1.0, v.abs().max().item())
        assert (pre1 - v).abs().max().item() <= 2e-3 * scale
    if n * h * w <= 20000:
        z = np.empty(n * h * w * 64, dtype=np.float32)
        _lib.check(lib.esrp_philox_normal_host(77, 5 << 36, z.size, z.ctypes.data), "philox")
        zt = torch.from_numpy(z).view(n, h, w, 64).to(cuda_dev)
        v = v + zt * 0.1 * v
        v = 0.2 * v + r2
        scale = max(1.0, v.abs().max().item())
        assert (of1 - v).abs().max().item() <= 2e-3 * scale
        assert (ob1.float() - v).abs().max().item() <= 1e-2 * scale


@pytest.mark.parametrize("shape", [(2, 20, 200), (4, 37, 130), (2, 1, 128), (6, 5, 128), (16, 128, 128)])
@pytest.mark.parametrize("case", ["conv1_aux", "conv2_kvalid", "conv4_kvalid", "conv5_slices"])
def test_conv3x3_cta_pairs(cuda_dev, shape, case):
    """ESRP_VARIANT_PAIR: the dense-block convs (block.py:260-268) on clusters of two CTAs that share every MMA
    (tcgen05 cta_group::2: images i and i + n/2 are the two halves of M = 256, each CTA holds half of the weight rows).
    Same products as the single-CTA launch, rows cut differently over the CTAs (so other rows are summed as main + shadow
    block): equal to fp32 rounding, and within the single-conv tolerance of torch conv2d.  (The co-scheduled-slices case is
    planned on single CTAs whatever the variant says - plan_row.inl - so it checks that the request is harmless.)"""
    n, h, w = shape
    g = torch.Generator(device=cuda_dev).manual_seed(4321 + w + h)
    t_in = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    gro = torch.randn(n, h, w, 128, device=cuda_dev, generator=g).to(torch.bfloat16)
    lay = _lib.LAYOUT_ROW
    r1 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g)
    r2 = torch.randn(n, h, w, 64, device=cuda_dev, generator=g)
    kw = dict(n=n, h=h, w=w, srcs=[t_in, gro], kc=64, bn=32, cout=32, w_layout=lay)
    if case == "conv1_aux":
        chunks, cin, cout = [(0, 0)], 64, 32
        wt = torch.randn(32, 64, 3, 3, device=cuda_dev, generator=g) / 24.0
        wa = torch.randn(32, 64, 1, 1, device=cuda_dev, generator=g) / 8.0
        bias = torch.randn(32, device=cuda_dev, generator=g)
        wp = K.pack_conv3x3_weights(wt, 64, 32, [0], w_aux=wa, aux_chunks=1, layout=lay)
        kw.update(chunks=chunks, w_packed=wp, bias=bias, act=1, aux_chunks=1)
        ref = _ref_conv([t_in, gro], chunks, 64, wt, bias, act=1) + F.conv2d(t_in.float().permute(0, 3, 1, 2).contiguous(), wa.to(torch.bfloat16).float())
    elif case in ("conv2_kvalid", "conv4_kvalid"):
        cin = 96 if case == "conv2_kvalid" else 160
        nch = (cin + 63) // 64
        chunks, cout = [(0, 0), (1, 0), (1, 64)][:nch], 32
        wt = torch.randn(32, cin, 3, 3, device=cuda_dev, generator=g) / (cin * 9) ** 0.5
        bias = torch.randn(32, device=cuda_dev, generator=g)
        wp = K.pack_conv3x3_weights(wt, 64, 32, [64 * i for i in range(nch)], layout=lay)
        kw.update(chunks=chunks, w_packed=wp, bias=bias, act=1, k_valid=cin)
        x = torch.cat([t_in, gro], dim=3)[..., :cin].float().permute(0, 3, 1, 2).contiguous()
        ref = F.leaky_relu(F.conv2d(x, wt.to(torch.bfloat16).float(), bias, padding=1), 0.2)
    else:
        chunks, cin, cout = [(0, 0), (1, 0), (1, 64)], 192, 64
        wt = torch.randn(64, 192, 3, 3, device=cuda_dev, generator=g) / (192 * 9) ** 0.5
        bias = torch.randn(64, device=cuda_dev, generator=g)
        lib = _lib.load()
        nbytes = lib.esrp_packed_conv3x3_bytes(3, 64, 32, 0)
        w_al = (nbytes + 1023) // 1024 * 1024
        stride = w_al + 1024
        buf = torch.zeros(2 * stride, dtype=torch.uint8, device=cuda_dev)
        for sl in (0, 1):
            K.pack_conv3x3_weights(wt, 64, 32, [0, 64, 128], row0=32 * sl, rows=32, layout=lay, out=buf[sl * stride: sl * stride + nbytes])
            buf[sl * stride + w_al: sl * stride + w_al + 128].view(torch.float32).copy_(bias[32 * sl: 32 * sl + 32])
        kw.update(chunks=chunks, w_packed=buf[:nbytes], bias=buf[w_al: w_al + 128].view(torch.float32), slices=2, slice_stride=stride,
                  s0=0.2, r1=r1, s1=1.0, r2=r2, s2=0.2)
        ref = _ref_conv([t_in, gro], chunks, 64, wt, bias, act=0)
        ref = 0.2 * (0.2 * ref + r1.permute(0, 3, 1, 2)) + r2.permute(0, 3, 1, 2)
    outs = []
    for variant in (_lib.VARIANT_ROW_ALT, _lib.VARIANT_ROW_ALT | _lib.VARIANT_PAIR):
        ob = torch.zeros((n, h, w, 64), device=cuda_dev, dtype=torch.bfloat16)
        of = torch.zeros((n, h, w, 64), device=cuda_dev)
        K.ConvCall(out_bf16=ob, out_f32=of, variant=variant, **kw).launch()
        outs.append((ob, of))
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    for ob, of in outs:
        assert (of[..., :cout].permute(0, 3, 1, 2) - ref).abs().max().item() <= 2e-3 * scale
        assert (ob[..., :cout].float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 1e-2 * scale
    d = (outs[0][1] - outs[1][1]).abs()
    assert bool((d <= 2e-5 * (outs[0][1].abs() + 1.0)).all()), d.max().item()   # (fp32 sums of 576 .. 1 728 products in another order)


@pytest.mark.parametrize("cin", [96, 160])
def test_conv3x3_k_valid_skips_zero_weight_tail(cuda_dev, cin):
    """conv2 / conv4 of a dense block (Cin 96 / 160, block.py:254,256) in 64-channel chunks: the tail of the last chunk
    has zero weights; with k_valid (and the row-alternating issuers, ESRP_VARIANT_ROW_ALT) the kernel does not issue those MMAs.  Garbage (even NaN) in the skipped channels of
    the source must not reach the output, and the result equals the padded launch bit for bit."""
    n, h, w = 2, 21, 140
    g = torch.Generator(device=cuda_dev).manual_seed(cin)
    nch = (cin + 63) // 64
    t_in = torch.randn(n, h, w, 64, device=cuda_dev, generator=g).to(torch.bfloat16)
    gro = torch.randn(n, h, w, 128, device=cuda_dev, generator=g).to(torch.bfloat16)
    chunks = [(0, 0), (1, 0), (1, 64)][:nch]
    wt = torch.randn(32, cin, 3, 3, device=cuda_dev, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(32, device=cuda_dev, generator=g)
    wp = K.pack_conv3x3_weights(wt, 64, 32, [64 * i for i in range(nch)], layout=_lib.LAYOUT_ROW)
    outs = []
    for kv in (0, cin):
        out = torch.zeros((n, h, w, 32), device=cuda_dev)
        K.ConvCall(n=n, h=h, w=w, srcs=[t_in, gro], kc=64, chunks=chunks, bn=32, cout=32, w_packed=wp, w_layout=_lib.LAYOUT_ROW,
                   bias=bias, act=1, out_f32=out, k_valid=kv, variant=_lib.VARIANT_ROW_ALT).launch()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    x = torch.cat([t_in, gro], dim=3)[..., :cin].float().permute(0, 3, 1, 2).contiguous()
    ref = F.leaky_relu(F.conv2d(x, wt.to(torch.bfloat16).float(), bias, padding=1), 0.2)
    assert (outs[1].permute(0, 3, 1, 2) - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    gro2 = gro.clone()
    gro2[..., cin - 64:] = float("nan")   # channels whose weights are zero: never multiplied when k_valid is given
    out = torch.zeros((n, h, w, 32), device=cuda_dev)
    K.ConvCall(n=n, h=h, w=w, srcs=[t_in, gro2], kc=64, chunks=chunks, bn=32, cout=32, w_packed=wp, w_layout=_lib.LAYOUT_ROW,
               bias=bias, act=1, out_f32=out, k_valid=cin, variant=_lib.VARIANT_ROW_ALT).launch()
    assert torch.equal(out, outs[1])


def test_conv3x3_rejects_bad_arguments(cuda_dev):
    s0 = torch.zeros(1, 8, 8, 64, device=cuda_dev, dtype=torch.bfloat16)
    wp = torch.zeros(9 * 32 * 64 * 2, device=cuda_dev, dtype=torch.uint8)  # 3 ky x 96 rows x 64 ch bf16
    out = torch.zeros(1, 8, 8, 32, device=cuda_dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="kc"):
        K.ConvCall(n=1, h=8, w=8, srcs=[s0], kc=48, chunks=[(0, 0)], bn=32, cout=32, w_packed=wp, out_bf16=out).launch()
    with pytest.raises(RuntimeError, match="channel range"):
        K.ConvCall(n=1, h=8, w=8, srcs=[s0], kc=64, chunks=[(0, 32)], bn=32, cout=32, w_packed=wp, out_bf16=out).launch()


def test_layout_converters_bit_exact(cuda_dev):
    x = torch.randn(2, 3, 19, 23, device=cuda_dev)
    y = K.nchw_f32_to_nhwc_bf16(x, 32)
    ref = torch.zeros(2, 19, 23, 32, device=cuda_dev, dtype=torch.bfloat16)
    ref[..., :3] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(y, ref)
    z = torch.randn(2, 9, 11, 64, device=cuda_dev).to(torch.bfloat16)
    assert torch.equal(K.upsample2x_nhwc_bf16(z), z.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    assert torch.equal(K.nhwc_bf16_to_nchw_f32(z, 64), z.float().permute(0, 3, 1, 2).contiguous())


# ---------------------------------------------------------------------------------------------------
# whole generator vs fixtures from the reference
# ---------------------------------------------------------------------------------------------------
def _make(cls, sd, nf, nb, dev):
    net = cls(3, 3, nf, nb, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA", upsample_mode="upconv")
    net.load_state_dict(sd, strict=True)  # base_model.py:60-63
    net.eval()
    for _, p in net.named_parameters():   # test_image/test.py:19-20
        p.requires_grad = False
    return net.to(dev)


@pytest.mark.parametrize("cls", [E.RRDBNet, E.RRDB_Net])
def test_rrdbnet_config1_matches_reference_fixture(cuda_dev, golden_dir, cls):
    g = _golden(golden_dir, "rrdbnet_c1_nb1_nf32.npz")
    sd = O.synth_state_dict_g(3, 3, 32, 1, seed=21)
    net = _make(cls, sd, 32, 1, cuda_dev)
    y = net(torch.from_numpy(g["x"]).to(cuda_dev)).cpu()
    _net_close(y, torch.from_numpy(g["y"]), "config 1")
    # test_image/test.py:31-40 plumbing: uint8 BGR -> /255 -> RGB CHW -> model -> clamp -> *255 round
    img = g["img_u8"] * 1.0 / 255
    t = torch.from_numpy(np.transpose(img[:, :, [2, 1, 0]], (2, 0, 1))).float().unsqueeze(0).to(cuda_dev)
    out = net(t).data.squeeze().float().cpu().clamp_(0, 1).numpy()
    out = (np.transpose(out[[2, 1, 0], :, :], (1, 2, 0)) * 255.0).round().astype("uint8")
    diff = np.abs(out.astype(int) - g["out_u8"].astype(int))
    assert diff.max() <= 2 and (diff > 0).mean() < 0.25, (diff.max(), (diff > 0).mean())


def test_uint8_plumbing_on_device_matches_host_plumbing(cuda_dev, golden_dir):
    """RRDBNet.forward_uint8 (esrp_rrdbnet_forward_u8): the /255, BGR<->RGB, clamp, x255, round of
    test_image/test.py:31-40 on the device.  (i) bit-identical to the same plumbing done on the host around the fp32
    module call; (ii) within the 8-bit tolerance of the reference fixture; (iii) a batch of ragged-size images."""
    g = _golden(golden_dir, "rrdbnet_c1_nb1_nf32.npz")
    sd = O.synth_state_dict_g(3, 3, 32, 1, seed=21)
    net = _make(E.RRDBNet, sd, 32, 1, cuda_dev)

    def host_plumbing(img_u8):
        img = img_u8 * 1.0 / 255
        t = torch.from_numpy(np.transpose(img[:, :, [2, 1, 0]], (2, 0, 1))).float().unsqueeze(0).to(cuda_dev)
        out = net(t).data.squeeze().float().cpu().clamp_(0, 1).numpy()
        return (np.transpose(out[[2, 1, 0], :, :], (1, 2, 0)) * 255.0).round().astype("uint8")

    dev_out = net.forward_uint8(torch.from_numpy(g["img_u8"]).unsqueeze(0).to(cuda_dev)).cpu().numpy()[0]
    assert dev_out.shape == (128, 128, 3) and dev_out.dtype == np.uint8
    assert np.array_equal(dev_out, host_plumbing(g["img_u8"]))
    diff = np.abs(dev_out.astype(int) - g["out_u8"].astype(int))
    assert diff.max() <= 2 and (diff > 0).mean() < 0.25, (diff.max(), (diff > 0).mean())
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, size=(3, 19, 37, 3), dtype=np.uint8)
    outs = net.forward_uint8(torch.from_numpy(imgs).to(cuda_dev)).cpu().numpy()
    assert outs.shape == (3, 76, 148, 3)
    for i in range(3):
        assert np.array_equal(outs[i], host_plumbing(imgs[i])), i
    # RGB-ordered input/output (bgr=False) is the channel-reversed problem
    rgb = net.forward_uint8(torch.from_numpy(imgs[:, :, :, ::-1].copy()).to(cuda_dev), bgr=False).cpu().numpy()
    assert np.array_equal(rgb[:, :, :, ::-1], outs)


def test_rrdbnet_nb23_matches_reference_fixture(cuda_dev, golden_dir):
    g = _golden(golden_dir, "rrdbnet_nb23_nf64.npz")
    sd = O.synth_state_dict_g(3, 3, 64, 23, seed=31)
    net = _make(E.RRDBNet, sd, 64, 23, cuda_dev)
    _net_close(net(torch.from_numpy(g["x24"]).to(cuda_dev)).cpu(), torch.from_numpy(g["y24"]), "nb23 24x24")
    # ragged: batch 2 of 19x37, no dimension a multiple of the 16x16 CTA tile
    _net_close(net(torch.from_numpy(g["x_ragged"]).to(cuda_dev)).cpu(), torch.from_numpy(g["y_ragged"]), "nb23 ragged")


def test_conv3x3_split_precision_output(cuda_dev):
    """esrp_conv3x3_t::out_lo: hi = bf16(v) and lo = bf16(v - hi) land in two channel ranges of one tensor; hi + lo
    reproduces the fp32 result to 2^-16 relative, and a conv fed [A_hi | A_lo] x [W_hi | W_hi | W_lo] (three chunk groups)
    reproduces the fp32 conv of the UNROUNDED operands to ~1e-5 (bf16 operands alone: ~4e-3)."""
    n, h, w, cin, cout = 2, 20, 27, 64, 32
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), wt.double(), bias.double(), padding=1).float()
    xh = x.to(torch.bfloat16)
    xl = (x - xh.float()).to(torch.bfloat16)
    a = torch.cat([xh, xl], 1).permute(0, 2, 3, 1).contiguous().to(cuda_dev)                      # [n,h,w, hi 64 | lo 64]
    wh = wt.to(torch.bfloat16).float()
    wl = (wt - wh).to(torch.bfloat16).float()
    wcat = torch.cat([wh, wh, wl], 1).to(cuda_dev)
    wp = K.pack_conv3x3_weights(wcat, 32, 32, list(range(0, 3 * cin, 32)), layout=_lib.LAYOUT_TILE)
    chunks = [(0, c) for c in (0, 32)] + [(0, c) for c in (64, 96)] + [(0, c) for c in (0, 32)]
    out = torch.zeros((n, h, w, 2 * cout), dtype=torch.bfloat16, device=cuda_dev)
    of = torch.zeros((n, h, w, cout), device=cuda_dev)
    K.ConvCall(n=n, h=h, w=w, srcs=[a], kc=32, chunks=chunks, bn=32, cout=cout, w_packed=wp, w_layout=_lib.LAYOUT_TILE,
               bias=bias.to(cuda_dev), out_bf16=out, ob_c0=0, ob_lo_c0=cout, out_f32=of).launch()
    y32 = of.permute(0, 3, 1, 2).cpu()
    hi, lo = out[..., :cout].float(), out[..., cout:].float()
    assert torch.equal(hi, of.to(torch.bfloat16).float())
    assert torch.equal(lo, (of - hi).to(torch.bfloat16).float())
    scale = ref.abs().max().item()
    err_split = (y32 - ref).abs().max().item() / scale
    err_bf16 = (F.conv2d(xh.float(), wh, bias, padding=1) - ref).abs().max().item() / scale
    print(f"split-precision conv: max err {err_split:.2e} of max|ref| (bf16 operands: {err_bf16:.2e})")
    assert err_split <= 3e-5 and err_bf16 >= 20 * err_split
    # out_lo is a tile-kernel feature: the row layout rejects it
    wr = K.pack_conv3x3_weights(wcat, 32, 32, list(range(0, 3 * cin, 32)), layout=_lib.LAYOUT_ROW)
    with pytest.raises(RuntimeError):
        K.ConvCall(n=n, h=h, w=w, srcs=[a], kc=32, chunks=chunks[:3], bn=32, cout=cout, w_packed=wr, w_layout=_lib.LAYOUT_ROW,
                   out_bf16=out, ob_lo_c0=cout).launch()


@pytest.mark.parametrize("cls", [E.RRDBNet, E.RRDB_Net])
def test_fp32_parity_mode_matches_reference_fixtures(cuda_dev, golden_dir, cls):
    """forward_fp32_parity (esrganplus_b200/precise.py, split precision on the same tcgen05 kernels) against outputs of the
    reference itself.  Stated tolerance: max|d| <= 2e-4 std(ref) and PSNR >= 90 dB on the clamped image — three orders of
    magnitude inside the fast path's 6e-2 / 48 dB, and below the 8-bit quantisation step by 50 dB."""
    g = _golden(golden_dir, "rrdbnet_c1_nb1_nf32.npz")
    net = _make(cls, O.synth_state_dict_g(3, 3, 32, 1, seed=21), 32, 1, cuda_dev)
    cases = [("config 1", net, g["x"], g["y"])]
    if cls is E.RRDBNet:
        g23 = _golden(golden_dir, "rrdbnet_nb23_nf64.npz")
        net23 = _make(cls, O.synth_state_dict_g(3, 3, 64, 23, seed=31), 64, 23, cuda_dev)
        cases += [("nb23 24x24", net23, g23["x24"], g23["y24"]), ("nb23 ragged 2x19x37", net23, g23["x_ragged"], g23["y_ragged"])]
    for what, m, x, yref in cases:
        ref = torch.from_numpy(yref)
        y = m.forward_fp32_parity(torch.from_numpy(x).to(cuda_dev)).cpu()
        rel = (y - ref).abs().max().item() / ref.std().item()
        psnr = O.psnr_255(y, ref)
        fast = m(torch.from_numpy(x).to(cuda_dev)).cpu()
        rel_fast = (fast - ref).abs().max().item() / ref.std().item()
        print(f"fp32-parity mode, {what}: max|d|/std {rel:.2e}, PSNR {psnr:.1f} dB (fast path: {rel_fast:.2e}, {O.psnr_255(fast, ref):.1f} dB)")
        assert y.shape == ref.shape and rel <= 2e-4 and psnr >= 90.0, (what, rel, psnr)
    # test_image/test.py:31-40 around the parity mode: the reference's 8-bit output image, to within one level in < 1 % of the
    # pixels (the fast path needs +-2 levels and up to 25 %: test_rrdbnet_config1_matches_reference_fixture)
    img = g["img_u8"] * 1.0 / 255
    t = torch.from_numpy(np.transpose(img[:, :, [2, 1, 0]], (2, 0, 1))).float().unsqueeze(0).to(cuda_dev)
    out = net.forward_fp32_parity(t).data.squeeze().float().cpu().clamp_(0, 1).numpy()
    out = (np.transpose(out[[2, 1, 0], :, :], (1, 2, 0)) * 255.0).round().astype("uint8")
    diff = np.abs(out.astype(int) - g["out_u8"].astype(int))
    print(f"fp32-parity mode, 8-bit image: max level difference {diff.max()}, {100.0 * (diff > 0).mean():.3f} % of the pixels differ")
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01, (diff.max(), (diff > 0).mean())
    if cls is E.RRDBNet:   # the same through forward_uint8(fp32_parity=True): plumbing on the device
        dev_out = net.forward_uint8(torch.from_numpy(g["img_u8"]).unsqueeze(0).to(cuda_dev), fp32_parity=True).cpu().numpy()[0]
        d2 = np.abs(dev_out.astype(int) - g["out_u8"].astype(int))
        assert dev_out.shape == g["out_u8"].shape and d2.max() <= 1 and (d2 > 0).mean() < 0.01, (d2.max(), (d2 > 0).mean())
    net.train()
    with pytest.raises(RuntimeError):
        net.forward_fp32_parity(torch.from_numpy(g["x"]).to(cuda_dev))


def test_weight_cache_follows_parameter_updates(cuda_dev):
    sd_a = O.synth_state_dict_g(3, 3, 32, 1, seed=1)
    sd_b = O.synth_state_dict_g(3, 3, 32, 1, seed=2)
    net = _make(E.RRDBNet, sd_a, 32, 1, cuda_dev)
    x = torch.rand(1, 3, 16, 16)
    ya = net(x.to(cuda_dev)).cpu()
    net.load_state_dict(sd_b, strict=True)        # in-place copy_ bumps ._version -> repack
    yb = net(x.to(cuda_dev)).cpu()
    _net_close(ya, O.rrdbnet_forward(x, sd_a, 1), "sd_a")
    _net_close(yb, O.rrdbnet_forward(x, sd_b, 1), "sd_b")
    with torch.no_grad():
        for p in net.parameters():                # optimizer-style in-place update
            p.mul_(0.5)
    yc = net(x.to(cuda_dev)).cpu()
    sd_c = {k: v * 0.5 for k, v in sd_b.items()}
    _net_close(yc, O.rrdbnet_forward(x, sd_c, 1), "sd_c")


def test_train_mode_noise_matches_oracle_with_same_draws(cuda_dev):
    """nESRGAN+ noise (block.py:117-121): the kernel draws N(0,1) from Philox(seed, rdb_index<<36 + e/4);
    esrp_philox_normal_host regenerates the identical draws so the oracle can be fed the same tensor."""
    lib = _lib.load()
    nb, nf, n, h, w = 2, 32, 2, 12, 20
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=77)
    net = _make(E.RRDBNet, sd, nf, nb, cuda_dev)
    net.train()
    x = torch.rand(n, 3, h, w)
    torch.manual_seed(1234)
    with torch.no_grad():
        y = net(x.to(cuda_dev)).cpu()
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + net._step) & 0xFFFFFFFFFFFFFFFF
        noises = []
        for i in range(nb):
            row = []
            for r in range(3):
                buf = np.empty(n * h * w * nf, dtype=np.float32)
                assert lib.esrp_philox_normal_host(seed, (i * 3 + r) << 36, buf.size, buf.ctypes.data) == 0
                row.append(torch.from_numpy(buf).reshape(n, h, w, nf).permute(0, 3, 1, 2).contiguous())
            noises.append(row)
        ref = O.rrdbnet_forward(x, sd, nb, training=True, noises=noises)
        ref_eval = O.rrdbnet_forward(x, sd, nb)
    _net_close(y, ref, "train-mode noise")
    assert (ref - ref_eval).abs().max() > 3 * (y - ref).abs().max(), "noise must matter more than the tolerance"
    # draws are N(0,1)
    z = torch.cat([t.flatten() for row in noises for t in row])
    assert abs(z.mean().item()) < 0.02 and abs(z.std().item() - 1.0) < 0.02
    # eval() switches it off; a new call draws new noise; the same seed+step reproduces
    with torch.no_grad():
        y2 = net(x.to(cuda_dev)).cpu()
    assert not torch.equal(y, y2)
    net.eval()
    with torch.no_grad():
        _net_close(net(x.to(cuda_dev)).cpu(), ref_eval, "eval after train")


# ---------------------------------------------------------------------------------------------------
# BASELINE.json config 2 size: 16 tiles of 128x128, nb=23 nf=64 — size-independent properties
# ---------------------------------------------------------------------------------------------------
def test_config2_full_size_properties(cuda_dev):
    sd = O.synth_state_dict_g(3, 3, 64, 23, seed=31)
    net = _make(E.RRDBNet, sd, 64, 23, cuda_dev)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(16, 3, 128, 128, generator=g)
    xd = x.to(cuda_dev)
    y = net(xd)
    assert y.shape == (16, 3, 512, 512) and torch.isfinite(y).all()
    # Run-to-run and batch-composition stability of the default (launch-per-conv) path: the MMA issue order is fixed
    # by the issuers' turn tokens, so a forward reproduces bit for bit.
    assert torch.equal(net(xd), y), "run-to-run"

    def _same(a, b, what):
        d = (a - b).abs().max().item() / y.std().item()
        assert d <= NET_REL_TOL / 10, f"{what}: {d:.3e}"
    # tiles are independent units: a tile's result does not depend on its batch mates
    perm = torch.arange(15, -1, -1)
    _same(net(xd[perm])[perm], y, "batch permutation")
    y1 = net(xd[3:4])
    _same(y1[0], y[3], "single tile vs batch")
    # one tile against the fp32 oracle at full tile size
    ref = O.rrdbnet_forward(x[3:4], sd, 23)
    _net_close(y1.cpu(), ref, "config 2 tile 3")


@pytest.mark.parametrize("shape", [(2, 40, 72), (3, 24, 24)])
def test_repeated_forwards_of_one_plan_are_consistent(cuda_dev, shape):
    """A cached launch plan is reused by every later forward of its shape: with other input / output tensors, on another
    stream, after a weight update (the plan holds pointers to the packed tiles, which are rebuilt in place), after a
    detour through another shape — and inside a CUDA graph captured by the caller (tools/bench_fwd_graph.py measured
    13.0 vs 13.15 ms for config 2: the launch stream is not what bounds the step, so the engine does not capture its
    own)."""
    sd = O.synth_state_dict_g(3, 3, 32, 2, seed=5)
    net = _make(E.RRDBNet, sd, 32, 2, cuda_dev)
    g = torch.Generator().manual_seed(1)
    xa, xb = (torch.rand(*[shape[0], 3, shape[1], shape[2]], generator=g).to(cuda_dev) for _ in range(2))
    ya = net(xa)
    assert torch.equal(net(xa), ya)
    yb = net(xb)
    assert not torch.equal(yb, ya)
    assert torch.equal(net(xa), ya)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        ys = net(xb)
    st.synchronize()
    assert torch.equal(ys, yb)
    _net_close(yb.cpu(), O.rrdbnet_forward(xb.cpu(), sd, 2), "repeated forward vs oracle")
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(0.5)
    sd_half = {k: v * 0.5 for k, v in sd.items()}
    _net_close(net(xb).cpu(), O.rrdbnet_forward(xb.cpu(), sd_half, 2), "forward after a weight update")
    net(xa[:1])                       # a shape change rebuilds the plan
    yh = net(xb)
    assert torch.equal(net(xb), yh)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        net(xb)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            yg = net(xb)
        gr.replay()
    side.synchronize()
    assert torch.equal(yg, yh)


# ---------------------------------------------------------------------------------------------------
# persistent conv chain (csrc/conv3x3_chain.cuh): the dense-block convs of the trunk as phases of ONE launch
# ---------------------------------------------------------------------------------------------------
# ragged widths, several images per CTA range, CTA ranges that cross image boundaries, images wider than one
# 128-pixel tile (dep_all), fewer rows than SMs (conv5 keeps its own launch: another grid), one-row images
CHAIN_SHAPES = [(3, 37, 100), (2, 70, 128), (1, 9, 300), (5, 3, 130), (4, 64, 65), (2, 1, 128), (7, 128, 128)]


def _rel_rms(a, b):
    s = b.std().item()
    return (a - b).abs().max().item() / s, (a - b).pow(2).mean().sqrt().item() / s


@pytest.mark.parametrize("shape", CHAIN_SHAPES)
def test_chain_matches_per_conv_launches_and_oracle(cuda_dev, shape):
    """The opt-in persistent chain computes what the launch-per-conv path computes.  Its three issuer threads reach the
    tensor pipe in no fixed order, so two runs (and the two paths) agree up to fp32 addition order, which the bf16
    rounding of every conv output amplifies to what either path differs from a storage-precision emulation by
    (tools/chain_diag.py: rms 1e-3 of std at nb = 1, maximum 1e-2): bound the rms 10x below the parity tolerance."""
    n, h, w = shape
    nb = 2
    sd = O.synth_state_dict_g(3, 3, 64, nb, seed=41)
    net = _make(E.RRDBNet, sd, 64, nb, cuda_dev)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(n, 3, h, w, generator=g)
    xd = x.to(cuda_dev)
    eng = net._engine_for(cuda_dev)
    y_plain = net(xd)
    assert eng.num_chained_convs == 0, "the chain is opt-in"
    assert torch.equal(net(xd), y_plain), "the launch-per-conv path reproduces bit for bit"
    eng.set_chain(True)
    try:
        y = net(xd)
        assert eng.num_chained_convs >= 4 * 3 * nb, (eng.num_chained_convs, eng.num_launches)
        for _ in range(3):
            d, rms = _rel_rms(net(xd), y)
            assert d <= NET_REL_TOL and rms <= NET_REL_TOL / 10, f"chain run-to-run: max {d:.3e} rms {rms:.3e}"
        d, rms = _rel_rms(y, y_plain)
        assert d <= NET_REL_TOL and rms <= NET_REL_TOL / 10, f"chain vs one launch per conv: max {d:.3e} rms {rms:.3e}"
        if n * h * w <= 3 * 37 * 130:
            _net_close(y.cpu(), O.rrdbnet_forward(x, sd, nb), f"chain {shape}")
    finally:
        eng.set_chain(False)


def test_chain_nb23_many_phases_back_to_back(cuda_dev):
    """345 phases per launch, 12 launches back to back (flags are reset per launch), 6 images of 96 rows so that CTA
    ranges cross image boundaries; one tile checked against the fp32 oracle."""
    sd = O.synth_state_dict_g(3, 3, 64, 23, seed=31)
    net = _make(E.RRDBNet, sd, 64, 23, cuda_dev)
    g = torch.Generator().manual_seed(6)
    x = torch.rand(6, 3, 96, 128, generator=g)
    xd = x.to(cuda_dev)
    eng = net._engine_for(cuda_dev)
    eng.set_chain(True)
    y = net(xd)
    assert eng.num_chained_convs == 345 and eng.num_launches <= 12, (eng.num_chained_convs, eng.num_launches)
    for _ in range(11):
        d, rms = _rel_rms(net(xd), y)
        assert d <= NET_REL_TOL and rms <= NET_REL_TOL / 10, f"run-to-run: max {d:.3e} rms {rms:.3e}"
    _net_close(net(xd[2:3]).cpu(), O.rrdbnet_forward(x[2:3], sd, 23), "chain nb23 tile 2")
    eng.set_chain(False)
    d, rms = _rel_rms(net(xd), y)
    assert d <= NET_REL_TOL and rms <= NET_REL_TOL / 10, f"chain vs plain at nb=23: max {d:.3e} rms {rms:.3e}"


def test_tiled_inference_matches_oracle_per_crop(cuda_dev):
    """BASELINE config 3 in miniature: an LR image cut into independent crops (esrganplus_b200/tiled.py);
    parity is per crop (zero padding at crop edges), exactly like running the reference on each crop."""
    from esrganplus_b200 import tiled
    sd = O.synth_state_dict_g(3, 3, 32, 2, seed=9)
    net = _make(E.RRDBNet, sd, 32, 2, cuda_dev)
    img = torch.rand(1, 3, 40, 72)
    out = tiled.infer_tiled(lambda b: net(b.to(cuda_dev)).cpu(), img, 24)
    assert out.shape == (1, 3, 160, 288)
    for (y, x, th, tw) in tiled.crop_grid(40, 72, 24):
        ref = O.rrdbnet_forward(img[:, :, y:y + th, x:x + tw], sd, 2)
        _net_close(out[:, :, 4 * y:4 * (y + th), 4 * x:4 * (x + tw)], ref, f"crop {y},{x}")


# ---------------------------------------------------------------------------------------------------
# Discriminator_VGG_128 forward (architecture.py:87-129) vs fixtures from the reference
# ---------------------------------------------------------------------------------------------------
def test_discriminator_forward_matches_reference_fixture(cuda_dev, golden_dir):
    """Eval (running statistics) and train-mode (batch statistics + running-stat update) forward.  bf16
    operands through ten conv layers: logits within 5e-2 * max(1, |ref|); BatchNorm running statistics are
    fp32 reductions of fp32 conv outputs: within 2e-2 relative."""
    g = _golden(golden_dir, "dvgg128.npz")
    sd = O.synth_state_dict_d(3, 64, seed=41)
    d = E.Discriminator_VGG_128(3, 64, norm_type="batch", act_type="leakyrelu", mode="CNA")
    d.load_state_dict(sd, strict=True)
    d = d.to(cuda_dev)
    x = torch.from_numpy(g["x"]).to(cuda_dev)

    def close(a, ref, what, tol):
        ref = torch.from_numpy(np.asarray(ref))
        err = (a.cpu() - ref).abs().max().item()
        assert err <= tol * max(1.0, ref.abs().max().item()), f"{what}: {err:.3e} vs max {ref.abs().max().item():.3e}"

    d.eval()
    with torch.no_grad():
        y = d(x)
    assert y.shape == (4, 1)
    close(y, g["y_eval"], "eval logits", 5e-2)
    d.train()
    with torch.no_grad():
        yt = d(x)
    close(yt, g["y_train"], "train logits", 5e-2)
    after = d.state_dict()
    for k, v in after.items():
        if "running" in k:
            close(v, g["after." + k], k, 2e-2)
        elif "num_batches" in k:
            assert int(v) == int(g["after." + k])
    yg = d(x.clone().requires_grad_(True))          # gradient-requiring forward: same numbers, autograd node attached
    assert yg.requires_grad and yg.grad_fn is not None


# ---------------------------------------------------------------------------------------------------
# ESRP_PAIR=1: the engine's dense-block convs on CTA pairs (experimental, opt-in; the switch is read once per process)
# ---------------------------------------------------------------------------------------------------
_PAIR_SCRIPT = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g
net = E.RRDBNet(3, 3, 64, 2); net.load_state_dict(random_state_dict_g(3, 3, 64, 2, seed=5)); net = net.cuda().eval()
x = torch.rand(4, 3, 40, 136, generator=torch.Generator().manual_seed(3)).cuda()
with torch.no_grad():
    y = net(x)
torch.save({"y": y.cpu(), "pairs": net._engines[x.device].num_pair_launches}, sys.argv[2])
"""


@pytest.mark.parametrize("single", ["0", "1"])   # ESRP_PAIR_SINGLE: alternating issuers / one issuer thread, one commit per row
def test_engine_with_cta_pairs_matches_default_engine(cuda_dev, tmp_path, single):
    """Whole generator (nb = 2, a batch of four 40 x 136 tiles: two column blocks, ragged width) with ESRP_PAIR=1 against the
    default engine, each in its own process: the dense-block convs run as cta_group::2 pairs over images (i, i + 2); the
    result agrees with the default path to the network tolerance of section 4.3 (same products, other summation order)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for pair in ("0", "1"):
        f = tmp_path / f"y{pair}.pt"
        env = dict(os.environ, ESRP_PAIR=pair, ESRP_PAIR_SINGLE=single)
        r = subprocess.run([sys.executable, "-c", _PAIR_SCRIPT, root, str(f)], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(f))
    assert outs[0]["pairs"] == 0 and outs[1]["pairs"] == 2 * 3 * 4, (outs[0]["pairs"], outs[1]["pairs"])   # conv1..conv4 of every dense block
    # (a CTA streams only a few rows here, no output row is summed over two blocks: the results are normally identical)
    _net_close(outs[1]["y"], outs[0]["y"], "ESRP_PAIR=1 vs default")

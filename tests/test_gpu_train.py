"""Generator training path (esrp_rrdbnet_train_forward / esrp_rrdbnet_backward through the drop-in
RRDBNet module and torch autograd) against the fp32 oracle differentiated by torch autograd on the CPU.

Two references, two tolerances (per parameter tensor: relative L2 error and cosine of the gradient):

* the fp32 oracle itself (REL_L2_FP32 / COS_FP32).  The kernels' operands are bf16 (forward activations carry
  ~1e-2 relative error after a few blocks), and the LeakyReLU derivative is a sign bit: an
  activation whose pre-activation lies within that error of zero takes slope 1 instead of 0.2 (or vice
  versa), so a fraction f of flipped bits alone moves a gradient by ~0.8*sqrt(f) in relative L2 — a few per
  cent per layer, which is what is measured.  This bound states parity with the reference arithmetic.
* the same oracle graph evaluated at the kernels' storage precision (`_emulated_forward`: weights and every
  stored activation rounded to bf16 with a straight-through gradient, fp32 trunk), differentiated by torch
  autograd.  Sign bits then agree except where fp32 summation order moves a value across a bf16 rounding
  boundary, and what is left is mostly the bf16 rounding of the gradients between launches: REL_L2 / COS
  (measured: <= 0.5 % at the last conv, 2-8 % further up — sums of sign-alternating bf16 gradients such as the
  HR-resolution bias gradients are the noisiest).  This bound
  is the one that catches a wrong operator in the backward graph (a missing term shows up as tens of %).
"""
import numpy as np
import pytest
import torch

import esrganplus_b200 as E
from esrganplus_b200 import _lib
from oracle import esrgan_oracle as O

pytestmark = pytest.mark.gpu

REL_L2 = 0.12
COS = 0.992
REL_L2_LINEAR = 1.5e-2
REL_L2_FP32 = 0.25
COS_FP32 = 0.97


def _r(t):
    """bf16 storage rounding with a straight-through gradient."""
    return t + (t.bfloat16().float() - t).detach()


def _emulated_forward(x, sd, nb, training=False, noises=None):
    """oracle.rrdbnet_forward (architecture.py:47-78, block.py:260-291) at the precision the kernels store
    things in: bf16 weights / conv operands, fp32 accumulation, fp32 residual trunk."""
    import torch.nn.functional as F
    W = {k: (_r(v) if k.endswith("weight") else v) for k, v in sd.items()}
    lr = lambda t: F.leaky_relu(t, 0.2)
    cv = lambda t, key: F.conv2d(t, W[key + ".weight"], W[key + ".bias"], padding=1)
    fea_f = cv(_r(x), "model.0")
    cur_f = fea_f
    for i in range(nb):
        rr_f = cur_f
        for r in (1, 2, 3):
            p = f"model.1.sub.{i}.RDB{r}."
            xb = _r(cur_f)
            x1 = _r(lr(cv(xb, p + "conv1.0")))
            x2 = _r(lr(cv(torch.cat((xb, x1), 1), p + "conv2.0")) + F.conv2d(xb, W[p + "conv1x1.weight"]))
            x3 = _r(lr(cv(torch.cat((xb, x1, x2), 1), p + "conv3.0")))
            x4 = _r(lr(cv(torch.cat((xb, x1, x2, x3), 1), p + "conv4.0")) + x2)
            t = 0.2 * cv(torch.cat((xb, x1, x2, x3, x4), 1), p + "conv5.0") + cur_f
            if training:
                t = t + noises[i][r - 1] * (0.1 * t)
            cur_f = 0.2 * t + rr_f if r == 3 else t
    t = _r(cv(_r(cur_f), f"model.1.sub.{nb}") + fea_f)
    for key in ("model.3", "model.6"):
        t = _r(lr(cv(F.interpolate(t, scale_factor=2, mode="nearest"), key)))
    t = _r(lr(cv(t, "model.8")))
    return cv(t, "model.10")


def _make(sd, nf, nb, dev, upscale=4):
    net = E.RRDBNet(3, 3, nf, nb, upscale=upscale)
    net.load_state_dict(sd, strict=True)
    return net.to(dev)


def _oracle_grads(x, sd, nb, dy, training=False, noises=None, emulate=False):
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    fwd = _emulated_forward if emulate else O.rrdbnet_forward
    y = fwd(x, sdg, nb, training=training, noises=noises)
    (y * dy).sum().backward()
    return y.detach(), {k: v.grad for k, v in sdg.items()}


def _check(net, x, sd, nb, dy, what, training=False, noises=None, y=None):
    ref_y, ref_g = _oracle_grads(x, sd, nb, dy, training, noises)
    if y is not None:
        err = (y.detach().cpu() - ref_y).abs().max().item() / ref_y.std().item()
        assert err < 6e-2, err
    _compare(net, ref_g, what + " vs fp32 oracle", REL_L2_FP32, COS_FP32)
    emu_y, emu_g = _oracle_grads(x, sd, nb, dy, training, noises, emulate=True)
    if y is not None:
        err = (y.detach().cpu() - emu_y).abs().max().item() / emu_y.std().item()
        assert err < 3e-2, err
    _compare(net, emu_g, what + " vs bf16-storage oracle", REL_L2, COS)
    return ref_g


def _compare(net, ref_grads, what, REL_L2=REL_L2, COS=COS):
    worst = (0.0, None)
    for k, p in net.named_parameters():
        assert p.grad is not None, f"{what}: {k} got no gradient"
        g = p.grad.detach().cpu().double()
        r = ref_grads[k].double()
        assert torch.isfinite(g).all(), f"{what}: {k} gradient not finite"
        den = r.norm().item()
        rel = (g - r).norm().item() / max(den, 1e-30)
        cos = (g * r).sum().item() / max(g.norm().item() * den, 1e-30)
        if rel > worst[0]:
            worst = (rel, k)
        assert rel <= REL_L2 and cos >= COS, f"{what}: {k} rel_l2={rel:.3e} cos={cos:.5f}"
    print(f"{what}: worst rel_l2 {worst[0]:.3e} at {worst[1]}")


@pytest.mark.parametrize("nf,nb,shape", [(64, 1, (2, 20, 24)), (32, 1, (1, 32, 32)), (64, 2, (1, 18, 70)), (32, 2, (2, 9, 13))])
def test_generator_backward_matches_oracle_autograd(cuda_dev, nf, nb, shape):
    n, h, w = shape
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=5 + nf + nb)
    net = _make(sd, nf, nb, cuda_dev).eval()   # eval: noise off, gradients still required
    g = torch.Generator().manual_seed(h * w)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, 4 * h, 4 * w, generator=g)
    y = net(x.to(cuda_dev))
    assert y.requires_grad
    (y * dy.to(cuda_dev)).sum().backward()
    _check(net, x, sd, nb, dy, f"nf{nf} nb{nb} {shape}", y=y)


def test_generator_gradients_match_reference_fixture(cuda_dev, golden_dir):
    """Gradients the REFERENCE's own RRDBNet(3,3,64,1) produced under torch autograd (fp32, make_golden_grads.py) against
    the native backward on the same weights / input / output gradient.  Stated bf16 bound: per tensor the norm within
    20 %, the stored head of the gradient cos >= 0.95, full small tensors rel-L2 <= REL_L2_FP32."""
    import os
    g = np.load(os.path.join(golden_dir, "rrdbnet_grad_nb1_nf64.npz"))
    sd = O.synth_state_dict_g(3, 3, 64, 1, seed=11)
    net = _make(sd, 64, 1, cuda_dev).eval()
    y = net(torch.from_numpy(g["x"]).to(cuda_dev))
    ref_y = torch.from_numpy(g["y"])
    assert (y.detach().cpu() - ref_y).abs().max().item() <= 6e-2 * ref_y.std().item()
    (y * torch.from_numpy(g["gy"]).to(cuda_dev)).sum().backward()
    worst = {"norm": 0.0, "cos": 1.0, "rel": 0.0}
    for k, p in net.named_parameters():
        gr = p.grad.detach().cpu().double()
        nrm = float(g["norm." + k][0])
        dn = abs(gr.norm().item() - nrm) / nrm
        head = torch.from_numpy(g["head." + k]).double()
        a = gr.flatten()[:64]
        cos = (a * head).sum().item() / max(a.norm().item() * head.norm().item(), 1e-30)
        worst["norm"] = max(worst["norm"], dn)
        worst["cos"] = min(worst["cos"], cos)
        assert dn <= 0.2, f"{k}: gradient norm {gr.norm().item():.4e} vs reference {nrm:.4e}"
        assert cos >= 0.95, f"{k}: head cos {cos:.4f}"
        if "full." + k in g.files:
            r = torch.from_numpy(g["full." + k]).double()
            rel = (gr - r).norm().item() / r.norm().item()
            worst["rel"] = max(worst["rel"], rel)
            assert rel <= REL_L2_FP32, f"{k}: rel_l2 {rel:.3e}"
    print("reference gradient fixture: worst", worst)


# Stated bf16 bounds at the configured depth (measured on B200, printed by the test): see the docstring.
NB23_REL_L2_FP32, NB23_COS_FP32 = 0.25, 0.97    # measured: worst 0.149 / 0.989 (default init), 0.118 / 0.995 (training init)
NB23_REL_L2_EMU, NB23_COS_EMU = 0.15, 0.99     # measured: worst 0.087 / 0.996 (default init), 0.018 / 0.9998 (training init)


@pytest.mark.parametrize("init", ["default", "train"])
def test_generator_backward_nb23_config4_shape(cuda_dev, init):
    """Gradient parity at the depth SRRaGAN_model.py:120,140 actually drives: the 23-block generator at config-4 crop
    shape (batch 2 of 32x32 LR), all 771 tensors, against the fp32 oracle and against the oracle at the kernels' storage
    precision.  `train` = the reference's training init (kaiming x 0.1, zero bias: networks.py:103-104), `default` =
    torch-default-like weights (the harder case: 69 blocks of O(1) residual branches).  The bounds are the measured bf16
    numbers with head room, per tensor; the median over tensors is printed and asserted 2x tighter
    (measured medians: 0.073 / 0.057 against fp32, 0.043 / 0.008 against the storage-precision oracle)."""
    from esrganplus_b200.synth import random_state_dict_g
    nb, nf, n, h, w = 23, 64, 2, 32, 32
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=77) if init == "default" else \
        random_state_dict_g(3, 3, nf, nb, seed=31, scale=0.1, zero_bias=True)
    net = _make(sd, nf, nb, cuda_dev).eval()
    g = torch.Generator().manual_seed(123)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, 4 * h, 4 * w, generator=g)
    y = net(x.to(cuda_dev))
    (y * dy.to(cuda_dev)).sum().backward()
    for emulate, rel_tol, cos_tol, tag in ((False, NB23_REL_L2_FP32, NB23_COS_FP32, "fp32 oracle"),
                                           (True, NB23_REL_L2_EMU, NB23_COS_EMU, "bf16-storage oracle")):
        ref_y, ref_g = _oracle_grads(x, sd, nb, dy, emulate=emulate)
        err = (y.detach().cpu() - ref_y).abs().max().item() / ref_y.std().item()
        rels, coss = [], []
        for k, p in net.named_parameters():
            gr, r = p.grad.detach().cpu().double(), ref_g[k].double()
            den = r.norm().item()
            if den == 0.0:
                continue
            rels.append(((gr - r).norm().item() / den, k))
            coss.append(((gr * r).sum().item() / max(gr.norm().item() * den, 1e-30), k))
        rels.sort()
        coss.sort()
        med_rel, med_cos = rels[len(rels) // 2][0], coss[len(coss) // 2][0]
        print(f"nb23 {init} vs {tag}: forward rel {err:.3e}; rel_l2 median {med_rel:.3e} worst {rels[-1][0]:.3e} ({rels[-1][1]}); "
              f"cos median {med_cos:.5f} worst {coss[0][0]:.5f} ({coss[0][1]})")
        assert rels[-1][0] <= rel_tol and coss[0][0] >= cos_tol, (tag, rels[-1], coss[0])
        assert med_rel <= rel_tol / 2, (tag, med_rel)


def test_weight_cache_follows_writes_through_dot_data(cuda_dev):
    """The reference's init_weights (networks.py:30-44) writes through `.data` inside `net.apply(fn)`: no tensor version
    counter moves.  The derived bf16 weight cache must still follow (architecture._NativeWeights epoch)."""
    nf, nb = 32, 1
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=3)
    net = _make(sd, nf, nb, cuda_dev).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    x = torch.rand(1, 3, 16, 16)
    with torch.no_grad():
        y0 = net(x.to(cuda_dev)).cpu()

        def scale_conv(m):
            if m.__class__.__name__.find("Conv") != -1:   # the reference's class-name dispatch
                m.weight.data *= 0.5
                if m.bias is not None:
                    m.bias.data.zero_()
        versions = [p._version for p in net.parameters()]
        net.apply(scale_conv)
        assert versions == [p._version for p in net.parameters()], "writes through .data move no version counter"
        y1 = net(x.to(cuda_dev)).cpu()
        sd1 = {k: (v * 0.5 if k.endswith("weight") else torch.zeros_like(v)) for k, v in sd.items()}
        ref1 = O.rrdbnet_forward(x, sd1, nb)
        assert (y1 - ref1).abs().max().item() <= 6e-2 * ref1.std().item(), "stale weights after net.apply(init_fn)"
        assert (y1 - y0).abs().max().item() > 0.1 * ref1.std().item()
        # any other write through .data: the documented explicit call
        for p in net.parameters():
            p.data.mul_(2.0)
        net.invalidate_weights()
        y2 = net(x.to(cuda_dev)).cpu()
        sd2 = {k: v * 2.0 for k, v in sd1.items()}
        ref2 = O.rrdbnet_forward(x, sd2, nb)
        assert (y2 - ref2).abs().max().item() <= 6e-2 * ref2.std().item()


def test_generator_backward_train_mode_noise(cuda_dev):
    """Train mode: the backward regenerates the Philox draws of the forward (gradient flows through the noise
    scale, block.py:119 is_relative_detach=False); the oracle is fed the same draws."""
    lib = _lib.load()
    nf, nb, n, h, w = 64, 2, 2, 16, 16
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=91)
    net = _make(sd, nf, nb, cuda_dev).train()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, 4 * h, 4 * w, generator=g)
    torch.manual_seed(99)
    y = net(x.to(cuda_dev))
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + net._step) & 0xFFFFFFFFFFFFFFFF
    (y * dy.to(cuda_dev)).sum().backward()
    noises = []
    for i in range(nb):
        row = []
        for r in range(3):
            buf = np.empty(n * h * w * nf, dtype=np.float32)
            assert lib.esrp_philox_normal_host(seed, (i * 3 + r) << 36, buf.size, buf.ctypes.data) == 0
            row.append(torch.from_numpy(buf).reshape(n, h, w, nf).permute(0, 3, 1, 2).contiguous())
        noises.append(row)
    ref_g = _check(net, x, sd, nb, dy, "train-mode noise", training=True, noises=noises, y=y)
    # the noise must matter: the fp32 gradients with and without it differ by > 10 % (measured: 20 %)
    _, g_eval = _oracle_grads(x, sd, nb, dy)
    k = "model.1.sub.0.RDB1.conv1.0.weight"
    assert (g_eval[k] - ref_g[k]).norm() > 0.1 * ref_g[k].norm()


def test_generator_backward_frozen_parameters_and_reuse(cuda_dev):
    """requires_grad=False parameters get no gradient; a second forward/backward pair on the same module after
    an optimizer-style update uses the new weights (forward and data-gradient caches are both derived)."""
    nf, nb, n, h, w = 32, 1, 1, 12, 12
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=8)
    net = _make(sd, nf, nb, cuda_dev).eval()
    frozen = "model.1.sub.0.RDB2.conv3.0.weight"
    dict(net.named_parameters())[frozen].requires_grad_(False)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, 4 * h, 4 * w, generator=g)
    (net(x.to(cuda_dev)) * dy.to(cuda_dev)).sum().backward()
    assert dict(net.named_parameters())[frozen].grad is None
    dict(net.named_parameters())[frozen].requires_grad_(True)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(0.9)
    net.zero_grad()
    (net(x.to(cuda_dev)) * dy.to(cuda_dev)).sum().backward()
    sd2 = {k: v * 0.9 for k, v in sd.items()}
    _check(net, x, sd2, nb, dy, "after update")


def test_generator_inference_path_unchanged_by_training_plan(cuda_dev):
    """no_grad forward after a training step (different crop size => other weight layout) still matches."""
    nf, nb = 64, 1
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=12)
    net = _make(sd, nf, nb, cuda_dev).eval()
    x_small = torch.rand(1, 3, 16, 16)
    x_wide = torch.rand(1, 3, 8, 80)
    net(x_small.to(cuda_dev)).sum().backward()
    with torch.no_grad():
        yw = net(x_wide.to(cuda_dev)).cpu()
    ref = O.rrdbnet_forward(x_wide, sd, nb)
    assert (yw - ref).abs().max().item() / ref.std().item() < 6e-2
    net.zero_grad()
    y = net(x_small.to(cuda_dev))
    y.sum().backward()
    _check(net, x_small, sd, nb, torch.ones(1, 3, 64, 64), "after layout switch", y=y)


@pytest.mark.parametrize("sign", [1.0, -1.0])
@pytest.mark.parametrize("nf,nb,shape", [(64, 2, (1, 18, 70)), (32, 1, (2, 16, 16))])
def test_generator_backward_operator_graph_without_sign_sensitivity(cuda_dev, nf, nb, shape, sign):
    """Every activated conv gets a large bias of one sign, so every LeakyReLU sits on one branch (slope 1 or
    0.2) on both sides and no sign bit can differ: what remains is the linear operator graph of the backward
    (data-gradient groups and scales, conv1x1, x4 = .. + x2, RRDB / shortcut skips, noise scale, upsample sums,
    weight-gradient units and scatter), held to REL_L2_LINEAR against the bf16-storage oracle."""
    lib = _lib.load()
    n, h, w = shape
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=17 + nf)
    for k in list(sd):
        act = (".conv" in k and ".conv5" not in k and "conv1x1" not in k) or k.split(".")[1] in ("3", "6", "8")
        if k.endswith("weight"):
            sd[k] = sd[k] * 0.5
        elif act:
            sd[k] = torch.full_like(sd[k], 6.0 * sign)
    net = _make(sd, nf, nb, cuda_dev).train()
    g = torch.Generator().manual_seed(11)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, 4 * h, 4 * w, generator=g)
    torch.manual_seed(5)
    y = net(x.to(cuda_dev))
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + net._step) & 0xFFFFFFFFFFFFFFFF
    (y * dy.to(cuda_dev)).sum().backward()
    noises = []
    for i in range(nb):
        row = []
        for r in range(3):
            buf = np.empty(n * h * w * nf, dtype=np.float32)
            assert lib.esrp_philox_normal_host(seed, (i * 3 + r) << 36, buf.size, buf.ctypes.data) == 0
            row.append(torch.from_numpy(buf).reshape(n, h, w, nf).permute(0, 3, 1, 2).contiguous())
        noises.append(row)
    emu_y, emu_g = _oracle_grads(x, sd, nb, dy, True, noises, emulate=True)
    _compare(net, emu_g, f"one-sided activations ({sign:+.0f})", REL_L2_LINEAR, 0.9998)


def test_graph_replay_draws_new_noise_and_matches_oracle(cuda_dev):
    """From the second forward/backward pair of a shape on, the engine replays CUDA graphs of its launch plans; the
    Philox key of the noise launches is read from device memory, so a replay must (i) draw NEW noise per call,
    (ii) reproduce under the same torch seed, (iii) still match the oracle fed with that call's draws."""
    lib = _lib.load()
    nf, nb, n, h, w = 32, 1, 2, 16, 16
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=23)
    net = _make(sd, nf, nb, cuda_dev).train()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, 4 * h, 4 * w, generator=g)
    torch.manual_seed(7)
    outs = []
    for it in range(3):
        net.zero_grad()
        y = net(x.to(cuda_dev))
        (y * dy.to(cuda_dev)).sum().backward()
        outs.append(y.detach().cpu())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + net._step) & 0xFFFFFFFFFFFFFFFF
    noises = []
    for i in range(nb):
        row = []
        for r in range(3):
            buf = np.empty(n * h * w * nf, dtype=np.float32)
            assert lib.esrp_philox_normal_host(seed, (i * 3 + r) << 36, buf.size, buf.ctypes.data) == 0
            row.append(torch.from_numpy(buf).reshape(n, h, w, nf).permute(0, 3, 1, 2).contiguous())
        noises.append(row)
    _check(net, x, sd, nb, dy, "third (replayed) call", training=True, noises=noises, y=outs[2])
    # same torch seed + same call count => same draws
    net2 = _make(sd, nf, nb, cuda_dev).train()
    torch.manual_seed(7)
    with torch.no_grad():
        pass
    y2 = None
    for it in range(3):
        net2.zero_grad()
        y2 = net2(x.to(cuda_dev))
        (y2 * dy.to(cuda_dev)).sum().backward()
    assert torch.equal(y2.detach().cpu(), outs[2])


@pytest.mark.parametrize("upscale", [1, 2])
def test_generator_backward_other_upscales(cuda_dev, upscale):
    """architecture.py:51-69: upscale 2 has one upconv block, upscale 1 none (the HR convs then run at LR resolution and
    the shortcut gradient comes straight from HR_conv0's data gradient)."""
    import torch.nn.functional as F
    nf, nb, n, h, w = 32, 1, 2, 12, 20
    n_up = {1: 0, 2: 1}[upscale]
    full = O.synth_state_dict_g(3, 3, nf, nb, seed=33)
    # reference key layout for n_up upconv blocks (sequential() flattening): model.{3,6} upconvs, then HR_conv0 / HR_conv1
    remap = {"model.0": "model.0", "model.1": "model.1"}
    src_hr0, src_hr1 = "model.8", "model.10"
    sd = {}
    for k, v in full.items():
        head = ".".join(k.split(".")[:2])
        if head in ("model.0", "model.1"):
            sd[k] = v
    for u in range(n_up):
        for suf in ("weight", "bias"):
            sd[f"model.{3 + 3 * u}.{suf}"] = full[f"model.{3 + 3 * u}.{suf}"]
    for suf in ("weight", "bias"):
        sd[f"model.{2 + 3 * n_up}.{suf}"] = full[f"{src_hr0}.{suf}"]
        sd[f"model.{4 + 3 * n_up}.{suf}"] = full[f"{src_hr1}.{suf}"]
    net = E.RRDBNet(3, 3, nf, nb, upscale=upscale)
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_dev).eval()
    g = torch.Generator().manual_seed(upscale)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, upscale * h, upscale * w, generator=g)
    y = net(x.to(cuda_dev))
    (y * dy.to(cuda_dev)).sum().backward()

    def ref_forward(xx, sdd):
        cv = lambda t, key: F.conv2d(t, _r(sdd[key + ".weight"]), sdd[key + ".bias"], padding=1)
        fea = cv(_r(xx), "model.0")
        t = fea
        p = "model.1.sub.0."
        rr = t
        for r in (1, 2, 3):
            q = f"{p}RDB{r}."
            xb = _r(t)
            x1 = _r(F.leaky_relu(cv(xb, q + "conv1.0"), 0.2))
            x2 = _r(F.leaky_relu(cv(torch.cat((xb, x1), 1), q + "conv2.0"), 0.2) + F.conv2d(xb, _r(sdd[q + "conv1x1.weight"])))
            x3 = _r(F.leaky_relu(cv(torch.cat((xb, x1, x2), 1), q + "conv3.0"), 0.2))
            x4 = _r(F.leaky_relu(cv(torch.cat((xb, x1, x2, x3), 1), q + "conv4.0"), 0.2) + x2)
            t = 0.2 * cv(torch.cat((xb, x1, x2, x3, x4), 1), q + "conv5.0") + t
        t = 0.2 * t + rr
        t = _r(cv(_r(t), "model.1.sub.1") + fea)
        for u in range(n_up):
            t = _r(F.leaky_relu(cv(F.interpolate(t, scale_factor=2, mode="nearest"), f"model.{3 + 3 * u}"), 0.2))
        t = _r(F.leaky_relu(cv(t, f"model.{2 + 3 * n_up}"), 0.2))
        return cv(t, f"model.{4 + 3 * n_up}")

    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ry = ref_forward(x, sdg)
    (ry * dy).sum().backward()
    assert (y.detach().cpu() - ry.detach()).abs().max().item() <= 3e-2 * ry.detach().std().item()
    _compare(net, {k: v.grad for k, v in sdg.items()}, f"upscale {upscale} vs bf16-storage oracle", REL_L2, COS)

"""Discriminator_VGG_128 training path (forward with saved state + backward through torch autograd) against the
fp32 oracle differentiated by torch autograd on the CPU — same two-reference scheme as tests/test_gpu_train.py:
loose bounds against the fp32 oracle (LeakyReLU sign bits differ where bf16 forward error crosses zero), tight
bounds against the oracle graph evaluated at the kernels' storage precision, and a one-sided-activation case
(every LeakyReLU on one branch) that isolates the linear operator graph of the backward.
"""
import pytest
import torch
import torch.nn.functional as F

import esrganplus_b200 as E
from oracle import esrgan_oracle as O

pytestmark = pytest.mark.gpu


def _r(t):
    return t + (t.bfloat16().float() - t).detach()


class _GradRound(torch.autograd.Function):
    """Identity whose backward rounds the gradient to bf16: the kernels hand dz (gradient of a conv output) and the
    data gradient of a layer to the next launch as bf16 tensors.  Under BatchNorm dz sums to ~0 per channel, so this
    rounding noise is what dominates sum-like gradients (conv biases, weights against large-mean activations)."""

    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _emulated_d(x, sd, training=True, eps=1e-5):
    """oracle.discriminator_vgg128_forward (architecture.py:87-129) with bf16 weights / conv operands and bf16
    gradient hand-offs between launches."""
    t = _r(x)
    for conv_idx, stride, bn_idx in O._D_LAYOUT:
        key = f"features.{conv_idx}"
        w = _r(sd[key + ".weight"])
        t = _GradRound.apply(F.conv2d(t, w, sd[key + ".bias"], stride=stride, padding=1))
        if bn_idx is not None:
            bkey = f"features.{bn_idx}"
            if training:
                mean, var = t.mean(dim=(0, 2, 3)), t.var(dim=(0, 2, 3), unbiased=False)
            else:
                mean, var = sd[bkey + ".running_mean"], sd[bkey + ".running_var"]
            t = (t - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
            t = t * sd[bkey + ".weight"][None, :, None, None] + sd[bkey + ".bias"][None, :, None, None]
        t = F.leaky_relu(t, 0.2)
        if conv_idx != 26:
            t = _GradRound.apply(_r(t))
    t = t.reshape(t.shape[0], -1)
    t = F.leaky_relu(F.linear(t, sd["classifier.0.weight"], sd["classifier.0.bias"]), 0.2)
    return F.linear(t, sd["classifier.2.weight"], sd["classifier.2.bias"])


def _ref(x, sd, r, training, emulate):
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    xg = x.clone().requires_grad_(True)
    y = _emulated_d(xg, sdg, training) if emulate else O.discriminator_vgg128_forward(xg, sdg, training)[0]
    (y * r).sum().backward()
    return y.detach(), xg.grad, {k: v.grad for k, v in sdg.items() if v.requires_grad}


def _rel(a, b):
    a, b = a.double().cpu(), b.double()
    return (a - b).norm().item() / max(b.norm().item(), 1e-30), (a * b).sum().item() / max(a.norm().item() * b.norm().item(), 1e-30)


def _run(cuda_dev, sd, x, r, training, frozen=False):
    d = E.Discriminator_VGG_128(3, 64, norm_type="batch", act_type="leakyrelu", mode="CNA")
    d.load_state_dict(sd, strict=True)
    d = d.to(cuda_dev)
    d.train(training)
    if frozen:
        for p in d.parameters():
            p.requires_grad_(False)
    xg = x.to(cuda_dev).requires_grad_(True)
    y = d(xg)
    (y * r.to(cuda_dev)).sum().backward()
    return d, y.detach().cpu(), xg.grad


def _check(d, y, dx, ref, what, rel_tol, cos_tol, params=True, dx_tol=None):
    ry, rdx, rg = ref
    assert (y - ry).abs().max().item() <= 5e-2 * max(1.0, ry.abs().max().item()), what
    rel, cos = _rel(dx, rdx)
    worst = (rel, "dx")
    assert rel <= (dx_tol or rel_tol) and cos >= cos_tol, f"{what}: dx rel_l2={rel:.3e} cos={cos:.5f}"
    if params:
        for k, p in d.named_parameters():
            assert p.grad is not None, k
            assert torch.isfinite(p.grad).all(), k
            if k.startswith("features.") and k.endswith(".bias") and k.split(".")[1] in ("2", "5", "8", "11", "14", "17", "20", "23", "26"):
                # a conv bias in front of a train-mode BatchNorm has zero gradient (the reference gets rounding noise)
                assert p.grad.abs().max().item() <= 5e-2 * max(1e-3, rg[k.replace(".bias", ".weight")].abs().max().item()), k
                continue
            rel, cos = _rel(p.grad, rg[k])
            if rel > worst[0]:
                worst = (rel, k)
            assert rel <= rel_tol and cos >= cos_tol, f"{what}: {k} rel_l2={rel:.3e} cos={cos:.5f}"
    print(f"{what}: worst rel_l2 {worst[0]:.3e} at {worst[1]}")


def test_discriminator_backward_matches_oracle_autograd(cuda_dev):
    sd = O.synth_state_dict_d(3, 64, seed=41)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(4, 3, 128, 128, generator=g)
    r = torch.randn(4, 1, generator=g)
    d, y, dx = _run(cuda_dev, sd, x, r, training=True)
    _check(d, y, dx, _ref(x, sd, r, True, False), "D train vs fp32 oracle", 0.35, 0.93)
    _check(d, y, dx, _ref(x, sd, r, True, True), "D train vs bf16-storage oracle", 0.2, 0.985)


def test_discriminator_input_gradient_matches_reference_fixture(cuda_dev, golden_dir):
    """dvgg128.npz: the gradient of sum(D(x)) w.r.t. the image that the REFERENCE's Discriminator_VGG_128 produced in
    train mode (fp32).  Ten bf16 layers with BatchNorm over a batch of 4 make this the most sign-sensitive quantity of
    the path; the stated bf16 bound is the one of the fp32-oracle test above."""
    import os
    import numpy as np
    g = np.load(os.path.join(golden_dir, "dvgg128.npz"))
    sd = O.synth_state_dict_d(3, 64, seed=41)
    x = torch.from_numpy(g["x"])
    d, y, dx = _run(cuda_dev, sd, x, torch.ones(4, 1), training=True)
    ref_y = torch.from_numpy(g["y_train"])
    assert (y - ref_y).abs().max().item() <= 5e-2 * max(1.0, ref_y.abs().max().item())
    rel, cos = _rel(dx, torch.from_numpy(g["gx_train"]))
    print(f"D input gradient vs reference fixture: rel_l2 {rel:.3e} cos {cos:.5f}")
    assert rel <= 0.3 and cos >= 0.95, (rel, cos)   # measured 0.198 / 0.980


def test_discriminator_frozen_input_gradient_vs_fp32_oracle(cuda_dev):
    """G phase (SRRaGAN_model.py:115-116,140) against the fp32 oracle (not only its bf16-storage emulation)."""
    sd = O.synth_state_dict_d(3, 64, seed=43)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 128, 128, generator=g)
    r = torch.randn(2, 1, generator=g)
    d, y, dx = _run(cuda_dev, sd, x, r, training=True, frozen=True)
    ry, rdx, _ = _ref(x, sd, r, True, False)
    rel, cos = _rel(dx, rdx)
    print(f"D frozen input gradient vs fp32 oracle: rel_l2 {rel:.3e} cos {cos:.5f}")
    assert rel <= 0.35 and cos >= 0.94, (rel, cos)   # measured 0.221 / 0.975


def test_discriminator_backward_frozen_gives_input_gradient_only(cuda_dev):
    """G phase (SRRaGAN_model.py:115-116,140): D parameters frozen, the gradient flows to the image."""
    sd = O.synth_state_dict_d(3, 64, seed=43)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 128, 128, generator=g)
    r = torch.randn(2, 1, generator=g)
    d, y, dx = _run(cuda_dev, sd, x, r, training=True, frozen=True)
    assert all(p.grad is None for p in d.parameters())
    _check(d, y, dx, _ref(x, sd, r, True, True), "D frozen vs bf16-storage oracle", 0.15, 0.99, params=False)


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_discriminator_backward_operator_graph_without_sign_sensitivity(cuda_dev, sign):
    """BatchNorm beta (and the first conv's bias) large and of one sign, gamma small: every LeakyReLU sits on one
    branch, so sign bits cannot differ.  The data-gradient chain (BatchNorm backward, inverse space-to-depth, all
    ten transposed convs) is then held to 2e-2 at the image; parameter gradients keep a loose bound because this
    construction gives activations a mean of ~8x their spread, which bf16 storage resolves to only ~16 levels (their
    tight operator-level check is tests/test_gpu_backward.py::test_k4s2_layer_gradients_match_conv2d_autograd)."""
    sd = O.synth_state_dict_d(3, 64, seed=47)
    for k in list(sd):
        idx = k.split(".")[1]
        if k.startswith("features.") and idx in ("3", "6", "9", "12", "15", "18", "21", "24", "27"):
            if k.endswith(".weight"):
                sd[k] = torch.full_like(sd[k], 0.5)
            elif k.endswith(".bias"):
                sd[k] = torch.full_like(sd[k], 4.0 * sign)
        if k == "features.0.bias":
            sd[k] = torch.full_like(sd[k], 6.0 * sign)
        if k == "classifier.0.bias":
            sd[k] = torch.full_like(sd[k], 40.0 * sign)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(4, 3, 128, 128, generator=g)
    r = torch.randn(4, 1, generator=g)
    d, y, dx = _run(cuda_dev, sd, x, r, training=True)
    _check(d, y, dx, _ref(x, sd, r, True, True), f"D one-sided ({sign:+.0f}) vs bf16-storage oracle", 0.3, 0.96, dx_tol=2e-2)


def test_discriminator_eval_mode_backward(cuda_dev):
    """BatchNorm in eval mode (running statistics): dz = gamma * rstd * dzb, no batch terms."""
    sd = O.synth_state_dict_d(3, 64, seed=45)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 128, 128, generator=g)
    r = torch.randn(2, 1, generator=g)
    d, y, dx = _run(cuda_dev, sd, x, r, training=False)
    ry, rdx, rg = _ref(x, sd, r, False, True)
    rel, cos = _rel(dx, rdx)
    assert rel <= 0.12 and cos >= 0.99, (rel, cos)
    for k in ("features.27.weight", "features.27.bias", "features.26.weight", "features.26.bias", "classifier.0.weight"):
        rel, cos = _rel(dict(d.named_parameters())[k].grad, rg[k])
        assert rel <= 0.12 and cos >= 0.99, (k, rel, cos)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [3, 6])
def test_first_pass_of_a_new_discriminator_reads_no_unwritten_memory(cuda_dev, n):
    """The first pass of a shape builds the packed weight tiles lazily while it records its launch plan; the deep layers
    fan their output slices out to side streams.  A tile packed after the fork is unordered with the launch that reads
    it: with a fresh process's zeroed pages that went unnoticed, with recycled memory (a second model in one process:
    bench.py's weak + strong training legs) the first logits were NaN.  Prime the allocator with NaN blocks, then the
    recording pass must equal the replayed ones (up to the summation order of the BatchNorm statistics), forward and input
    gradient."""
    from esrganplus_b200.synth import random_state_dict_d
    blocks = [torch.full((sz // 4,), float("nan"), device=cuda_dev)
              for sz in [1 << 28] * 8 + [1 << 24] * 16 + [1 << 20] * 64 + [1 << 16] * 128 + [1 << 12] * 256 + [512] * 1024]
    torch.cuda.synchronize()
    del blocks
    d = E.Discriminator_VGG_128(3, 64)
    d.load_state_dict(random_state_dict_d(3, 64, seed=77), strict=True)
    d = d.to(cuda_dev).train()
    for p in d.parameters():
        p.requires_grad = False
    x = torch.rand(n, 3, 128, 128, generator=torch.Generator().manual_seed(n)).to(cuda_dev)
    outs = []
    for _ in range(3):
        xi = x.clone().requires_grad_(True)
        y = d(xi)
        y.sum().backward()
        outs.append((y.detach().clone(), xi.grad.clone()))
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[0][1]).all()
    for y, dx in outs[1:]:
        assert torch.allclose(y, outs[0][0], rtol=1e-4, atol=1e-5)
        assert (dx - outs[0][1]).norm().item() <= 1e-3 * outs[0][1].norm().item()

"""One full GAN training step (generator + discriminator forward/backward on the native kernels, the solver logic of
SRRaGAN_model.py:113-186 restated in esrganplus_b200/gan_step.py) against the fixture produced by the REFERENCE
SOLVER itself (tests/golden/make_golden_train_step.py: SRRaGANModel.optimize_parameters on CPU, noise off).

Tolerances: losses within 2 % (+1e-4 absolute), D logit means within 2e-2, per-tensor gradient norms within
40 % (G) / 25 % (D) (bias tensors 50 %), stored gradients cos >= 0.8 (0.7 for bias gradients: sums of sign-alternating terms).  The gradient bounds are loose on purpose: the fixture is the
fp32 reference, the discriminator here has random synthetic weights and BatchNorm over a batch of 2, and its input
gradient is then chaotic in the forward precision — the fp32 oracle and the same oracle at bf16 storage precision
differ from EACH OTHER by rel-L2 0.4 / cos 0.91 on this very input, while the kernels match the bf16-storage oracle
to 0.11 / 0.993 (tests/diag/diag_gphase.py; the per-network tests hold the tight bounds).  What this test pins is the
solver logic: phases, freezing, the relativistic losses, both Adam steps, BatchNorm statistics.
"""
import os

import numpy as np
import pytest
import torch

import esrganplus_b200 as E
from esrganplus_b200.gan_step import GanTrainStep
from oracle import esrgan_oracle as O

pytestmark = pytest.mark.gpu

NORM_TOL = {"g": 0.4, "d": 0.25}
# Bias gradients are sums of sign-alternating terms over all pixels: the norm of G's last bias (model.10.bias, fed by the
# chaotic D input gradient described above) lands 35 % or 42 % below the fp32 fixture depending only on the order in which
# the kernel accumulates the taps of a row in fp32 (two valid summation orders, ESRP_ROW_ALT=0 / 2; cos 0.963 / 0.968).
BIAS_NORM_TOL = 0.5
COS_TOL = {"g": 0.8, "d": 0.8}


def test_gan_train_step_matches_reference_solver(cuda_dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "train_step_nb1_nf32.npz"))
    netG = E.RRDBNet(3, 3, 32, 1)
    netG.load_state_dict(O.synth_state_dict_g(3, 3, 32, 1, seed=61), strict=True)
    netD = E.Discriminator_VGG_128(3, 64, norm_type="batch", act_type="leakyrelu", mode="CNA")
    netD.load_state_dict(O.synth_state_dict_d(3, 64, seed=62), strict=True)
    netG, netD = netG.to(cuda_dev).train(), netD.to(cuda_dev).train()
    # the fixture was made with GaussianNoise switched off (see make_golden_train_step.py): eval() on G does exactly that
    netG.eval()
    step = GanTrainStep(netG, netD)
    log = step.step(torch.from_numpy(g["lr"]).to(cuda_dev), torch.from_numpy(g["hr"]).to(cuda_dev))
    fake = step.fake_H.detach().cpu()
    ref_fake = torch.from_numpy(g["fake_H"])
    assert (fake - ref_fake).abs().max().item() <= 6e-2 * ref_fake.std().item()
    for k in ("l_g_pix", "l_g_gan", "l_d_real", "l_d_fake"):
        ref = float(g["log." + k])
        assert abs(log[k].item() - ref) <= 2e-2 * abs(ref) + 1e-4, (k, log[k].item(), ref)
    for k in ("D_real", "D_fake"):
        assert abs(log[k].item() - float(g["log." + k])) <= 2e-2, (k, log[k].item(), float(g["log." + k]))
    bad = []
    for tag, net in (("g", netG), ("d", netD)):
        names = [str(s) for s in g[f"names_{tag}"]]
        ref_norm = dict(zip(names, g[f"gradnorm_{tag}"]))
        ref_pn = dict(zip(names, g[f"paramnorm_after_{tag}"]))
        for k, p in net.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
            is_dead_bias = tag == "d" and k.endswith(".bias") and k.split(".")[1] in ("2", "5", "8", "11", "14", "17", "20", "23", "26")
            if not is_dead_bias:   # conv bias in front of a train-mode BatchNorm: true gradient is zero
                gn = p.grad.norm().item()
                bad += [(k, gn, ref_norm[k])] if abs(gn - ref_norm[k]) > (BIAS_NORM_TOL if k.endswith(".bias") else NORM_TOL[tag]) * ref_norm[k] + 1e-12 else []
            # Adam moved every parameter by at most lr per element: norms after the step agree closely
            assert abs(p.detach().norm().item() - ref_pn[k]) <= 1e-3 * ref_pn[k] + 1e-3, k
            fk = f"grad_{tag}.{k}"
            if fk in g.files:
                r = torch.from_numpy(g[fk]).double()
                a = p.grad.cpu().double()
                if r.norm().item() == 0.0:
                    assert a.norm().item() <= 1e-9, k
                    continue
                cos = (a * r).sum().item() / max(a.norm().item() * r.norm().item(), 1e-30)
                print(f"{tag} {k}: cos {cos:.4f} |g| {a.norm().item():.3e} |ref| {r.norm().item():.3e}")
                bad += [(k, cos)] if cos < (COS_TOL[tag] - (0.1 if k.endswith('.bias') else 0.0)) else []
    assert not bad, bad
    sd = netD.state_dict()
    for k in g.files:
        if k.startswith("after_d."):
            ref = torch.from_numpy(g[k])
            assert (sd[k[len("after_d."):]].cpu() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item()), k

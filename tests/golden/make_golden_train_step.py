#!/usr/bin/env python
"""Golden fixture for ONE full GAN training step, produced by the REFERENCE SOLVER ITSELF:
codes/models/SRRaGAN_model.py SRRaGANModel.optimize_parameters (:113-186) on CPU, ESRGAN+ recipe without the
perceptual branch (BASELINE.json config 4 at a size the CPU finishes in seconds: G nb=1 nf=32, D_VGG_128, bs 2).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_train_step.py

Interventions, all in memory (reference files untouched):
  * GaussianNoise.__init__ replaced (block.py:115 hard-codes .to('cuda'));
  * GaussianNoise.forward replaced by the identity: the reference draws its noise from torch's global generator,
    the kernels from Philox(seed, block) — noise parity is covered with injected draws in tests/test_gpu_train.py,
    here the step is made deterministic;
  * weights: deterministic synthetic state_dicts (oracle.synth_state_dict_*) loaded with load_state_dict(strict=True)
    after the solver's own init, so the fixture does not depend on torch's RNG stream.
Stored: inputs, the solver's log_dict, per-tensor gradient norms of G (after l_g_total.backward()) and D (after
l_d_total.backward()), full gradients of a few small tensors, and parameter norms after both Adam steps.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402
from oracle import esrgan_oracle as O  # noqa: E402

FULL = ["model.0.weight", "model.0.bias", "model.1.sub.0.RDB1.conv1x1.weight", "model.1.sub.0.RDB2.conv3.0.weight",
        "model.1.sub.0.RDB3.conv5.0.bias", "model.1.sub.1.weight", "model.10.weight", "model.10.bias"]
FULL_D = ["features.0.weight", "features.0.bias", "features.3.weight", "features.3.bias", "features.27.weight",
          "features.27.bias", "classifier.2.weight", "classifier.2.bias", "classifier.0.bias"]


def main():
    arch, B = import_reference()
    B.GaussianNoise.forward = lambda self, x: x
    from models.SRRaGAN_model import SRRaGANModel
    from options.options import NoneDict, dict_to_nonedict

    opt = dict_to_nonedict({
        "name": "golden", "model": "srragan", "scale": 4, "gpu_ids": None, "is_train": True,
        "path": {"root": "/tmp", "experiments_root": "/tmp/esrp_golden", "models": "/tmp/esrp_golden/models",
                 "training_state": "/tmp/esrp_golden/state", "log": "/tmp/esrp_golden", "val_images": "/tmp/esrp_golden/val"},
        "network_G": {"which_model_G": "RRDB_net", "norm_type": None, "mode": "CNA", "nf": 32, "nb": 1, "in_nc": 3,
                      "out_nc": 3, "gc": 32, "group": 1, "scale": 4},
        "network_D": {"which_model_D": "discriminator_vgg_128", "norm_type": "batch", "act_type": "leakyrelu",
                      "mode": "CNA", "nf": 64, "in_nc": 3},
        "train": {"lr_G": 1e-4, "weight_decay_G": 0, "beta1_G": 0.9, "lr_D": 1e-4, "weight_decay_D": 0, "beta1_D": 0.9,
                  "lr_scheme": "MultiStepLR", "lr_steps": [50000, 100000], "lr_gamma": 0.5,
                  "pixel_criterion": "l1", "pixel_weight": 1e-2, "feature_criterion": "l1", "feature_weight": 0,
                  "gan_type": "vanilla", "gan_weight": 5e-3, "manual_seed": 0, "niter": 10, "val_freq": 1000},
    })
    torch.manual_seed(0)
    model = SRRaGANModel(opt)
    sd_g = O.synth_state_dict_g(3, 3, 32, 1, seed=61)
    sd_d = O.synth_state_dict_d(3, 64, seed=62)
    model.netG.load_state_dict(sd_g, strict=True)
    model.netD.load_state_dict(sd_d, strict=True)
    g = torch.Generator().manual_seed(7)
    lr = torch.rand(2, 3, 32, 32, generator=g)
    hr = torch.rand(2, 3, 128, 128, generator=g)
    model.feed_data({"LR": lr, "HR": hr})
    model.optimize_parameters(1)
    log = {k: float(v) for k, v in model.get_current_log().items()}
    out = {"lr": lr.numpy(), "hr": hr.numpy(), "fake_H": model.fake_H.detach().numpy()}
    for k, v in log.items():
        out["log." + k] = np.float32(v)
    for tag, net, full in (("g", model.netG, FULL), ("d", model.netD, FULL_D)):
        names, gn, pn = [], [], []
        for k, p in net.named_parameters():
            names.append(k)
            gn.append(p.grad.norm().item())
            pn.append(p.detach().norm().item())
            if k in full:
                out[f"grad_{tag}.{k}"] = p.grad.numpy().astype(np.float32)
        out[f"names_{tag}"] = np.array(names)
        out[f"gradnorm_{tag}"] = np.array(gn, dtype=np.float64)
        out[f"paramnorm_after_{tag}"] = np.array(pn, dtype=np.float64)
    for k, v in model.netD.state_dict().items():
        if "running" in k and k.split(".")[1] in ("3", "27"):
            out["after_d." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "train_step_nb1_nf32.npz"), **out)
    print("log_dict:", log)
    print("wrote", os.path.join(HERE, "train_step_nb1_nf32.npz"), os.path.getsize(os.path.join(HERE, "train_step_nb1_nf32.npz")), "bytes")


if __name__ == "__main__":
    main()

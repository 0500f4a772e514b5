#!/usr/bin/env python
"""Fixture made by the REFERENCE's own VGGFeatureExtractor (build container only: needs /root/reference and torchvision):
`vgg_feature_34.npz`.

    python tests/golden/make_golden_vgg.py

`arch.VGGFeatureExtractor(feature_layer=34, use_bn=False, use_input_norm=True)` of
codes/models/modules/architecture.py:279-307 is imported unmodified.  Its constructor asks torchvision for the pretrained
vgg19 (:289), which needs the network; for the duration of the constructor `torchvision.models.vgg19` is replaced by one
that returns the same architecture un-initialised — the class then cuts `features[:35]` itself (:298) — and the synthetic
weights of oracle.synth_state_dict_vgg(seed 3) are loaded with load_state_dict(strict=True).
Stored: two images [1,3,128,128] in [0,1] ("fake", "real"), their features, the L1 feature loss of
SRRaGAN_model.py:128-130 (real detached) and its gradient w.r.t. the fake batch from torch autograd.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, rng_tensor, sd_digest, O  # noqa: E402


def main():
    import torchvision
    arch, _ = import_reference()
    orig = torchvision.models.vgg19
    torchvision.models.vgg19 = lambda pretrained=False, **kw: orig(weights=None)
    try:
        net = arch.VGGFeatureExtractor(feature_layer=34, use_bn=False, use_input_norm=True, device=torch.device("cpu"))
    finally:
        torchvision.models.vgg19 = orig
    sd = O.synth_state_dict_vgg(34, seed=3)
    net.load_state_dict(sd, strict=True)
    net.eval()
    fake = rng_tensor(601, 1, 3, 128, 128).requires_grad_(True)
    real = rng_tensor(602, 1, 3, 128, 128)
    real_fea = net(real).detach()
    fake_fea = net(fake)
    loss = torch.nn.L1Loss()(fake_fea, real_fea)
    loss.backward()
    out = {"fake": fake.detach().numpy(), "real": real.numpy(), "fake_fea": fake_fea.detach().numpy(), "real_fea": real_fea.numpy(),
           "loss": np.array(loss.item(), dtype="float64"), "dfake": fake.grad.numpy(), "sd_digest": np.array(sd_digest(sd)),
           "keys": np.array(list(net.state_dict().keys())), "repr": np.array(str(net))}
    path = os.path.join(HERE, "vgg_feature_34.npz")
    np.savez_compressed(path, **out)
    print("vgg_feature_34.npz", os.path.getsize(path), "loss", loss.item(), "fea std", fake_fea.std().item())


if __name__ == "__main__":
    main()

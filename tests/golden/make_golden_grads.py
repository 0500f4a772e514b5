#!/usr/bin/env python
"""Gradient fixtures made by the REFERENCE's own RRDBNet under torch autograd (build container only: needs
/root/reference): `rrdbnet_grad_nb1_nf64.npz`.

    python tests/golden/make_golden_grads.py

RRDBNet(3, 3, 64, 1) of codes/models/modules/architecture.py:47-78, imported unmodified (same GaussianNoise.__init__
shim as make_golden.py), synthetic weights (oracle.synth_state_dict_g seed 11: the weights of rdb64.npz), eval() so the
noise is off, x [2,3,16,16], loss = sum(y * gy).  Stored: x, gy, y, and for each of the 45 parameter tensors the gradient's
(norm, sum, max|.|), its first 64 elements, and the full gradient where the tensor has at most 4096 elements.
tests/test_gpu_train.py::test_generator_gradients_match_reference_fixture holds the native backward against it.
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
    arch, _ = import_reference()
    sd = O.synth_state_dict_g(3, 3, 64, 1, seed=11)
    net = arch.RRDBNet(in_nc=3, out_nc=3, nf=64, nb=1, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA",
                       upsample_mode="upconv")
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = rng_tensor(501, 2, 3, 16, 16)
    gy = rng_tensor(502, 2, 3, 64, 64, lo=-1, hi=1)
    y = net(x)
    (y * gy).sum().backward()
    out = {"x": x.numpy(), "gy": gy.numpy(), "y": y.detach().numpy(), "sd_digest": np.array(sd_digest(sd))}
    names = []
    for k, p in net.named_parameters():
        g = p.grad.detach()
        names.append(k)
        out["norm." + k] = np.array([g.norm().item(), g.sum().item(), g.abs().max().item()], dtype="float64")
        out["head." + k] = g.flatten()[:64].numpy()
        if g.numel() <= 4096:
            out["full." + k] = g.numpy()
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "rrdbnet_grad_nb1_nf64.npz"), **out)
    print("rrdbnet_grad_nb1_nf64.npz", os.path.getsize(os.path.join(HERE, "rrdbnet_grad_nb1_nf64.npz")), len(names), "tensors")


if __name__ == "__main__":
    main()

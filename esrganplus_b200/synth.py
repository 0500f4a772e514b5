"""Random-init weights with the reference's key set and shapes, for benchmarks and profiling scripts:
there is no network for checkpoints, so bench.py times the named architecture on synthetic weights (numpy PCG64, stable
across torch versions).  U(-1, 1) / sqrt(fan_in) magnitudes, i.e. torch's default conv init; `scale` < 1 mimics the
kaiming x 0.1 the reference applies to the generator before training (networks.py:103-104)."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch


def _u(rng, lo, hi, shape):
    return torch.from_numpy(rng.uniform(lo, hi, shape).astype("float32"))


def random_state_dict_g(in_nc: int, out_nc: int, nf: int, nb: int, gc: int = 32, upscale: int = 4, seed: int = 0,
                        scale: float = 1.0, zero_bias: bool = False) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(key, cout, cin, k, bias=True):
        b = scale / (cin * k * k) ** 0.5
        sd[key + ".weight"] = _u(rng, -b, b, (cout, cin, k, k))
        if bias:
            sd[key + ".bias"] = torch.zeros(cout) if zero_bias else _u(rng, -b, b, (cout,))

    conv("model.0", nf, in_nc, 3)
    for i in range(nb):
        for r in (1, 2, 3):
            p = f"model.1.sub.{i}.RDB{r}."
            conv(p + "conv1x1", gc, nf, 1, bias=False)
            for k in range(1, 5):
                conv(p + f"conv{k}.0", gc, nf + (k - 1) * gc, 3)
            conv(p + "conv5.0", nf, nf + 4 * gc, 3)
    conv(f"model.1.sub.{nb}", nf, nf, 3)
    n_up = {1: 0, 2: 1, 4: 2}[upscale]
    for u in range(n_up):
        conv(f"model.{3 + 3 * u}", nf, nf, 3)
    conv(f"model.{2 + 3 * n_up}", nf, nf, 3)
    conv(f"model.{4 + 3 * n_up}", out_nc, nf, 3)
    return sd


def random_state_dict_d(in_nc: int = 3, base_nf: int = 64, seed: int = 0) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    nf = base_nf
    plan = [(in_nc, nf, 3, False), (nf, nf, 4, True), (nf, 2 * nf, 3, True), (2 * nf, 2 * nf, 4, True),
            (2 * nf, 4 * nf, 3, True), (4 * nf, 4 * nf, 4, True), (4 * nf, 8 * nf, 3, True), (8 * nf, 8 * nf, 4, True),
            (8 * nf, 8 * nf, 3, True), (8 * nf, 8 * nf, 4, True)]
    idx = 0
    for cin, cout, k, bn in plan:
        b = (6.0 / (cin * k * k)) ** 0.5 * 0.5
        sd[f"features.{idx}.weight"] = _u(rng, -b, b, (cout, cin, k, k))
        sd[f"features.{idx}.bias"] = _u(rng, -0.1, 0.1, (cout,))
        idx += 1
        if bn:
            sd[f"features.{idx}.weight"] = _u(rng, 0.5, 1.5, (cout,))
            sd[f"features.{idx}.bias"] = _u(rng, -0.2, 0.2, (cout,))
            sd[f"features.{idx}.running_mean"] = torch.zeros(cout)
            sd[f"features.{idx}.running_var"] = torch.ones(cout)
            sd[f"features.{idx}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
            idx += 1
        idx += 1  # the LeakyReLU slot
    b = 1.0 / (8 * nf * 16) ** 0.5
    sd["classifier.0.weight"] = _u(rng, -b, b, (100, 8 * nf * 16))
    sd["classifier.0.bias"] = _u(rng, -b, b, (100,))
    sd["classifier.2.weight"] = _u(rng, -0.1, 0.1, (1, 100))
    sd["classifier.2.bias"] = _u(rng, -0.1, 0.1, (1,))
    return sd


def random_state_dict_vgg(feature_layer: int = 34, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic VGGFeatureExtractor weights (architecture.py:279-307; the pretrained torchvision file needs the network):
    uniform with the kaiming bound so that the activations keep their scale through the sixteen layers."""
    from .architecture import _VGG19_CFG
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {"mean": torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1),
                                   "std": torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)}
    idx, cin = 0, 3
    for v in _VGG19_CFG:
        if v == "M":
            idx += 1
            continue
        if idx <= feature_layer:
            b = (6.0 / (cin * 9)) ** 0.5
            sd[f"features.{idx}.weight"] = _u(rng, -b, b, (v, cin, 3, 3))
            sd[f"features.{idx}.bias"] = _u(rng, -0.05, 0.05, (v,))
        idx += 2
        cin = v
    return sd

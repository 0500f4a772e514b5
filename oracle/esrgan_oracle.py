"""CPU oracle for the ESRGAN+/nESRGAN+ hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under esrganplus_b200/ does.

A plain functional fp32 restatement (torch CPU ops on a state_dict, no nn.Module tree) of the
algorithm in the reference repository ncarraz/ESRGANplus; each function cites the reference
file:line it follows.  The convolution arithmetic itself is a third-party dependency of the
reference (PyTorch `nn.Conv2d` -> ATen; the reference pins only "PyTorch >= 1.0.0", README.md:20),
so `torch.nn.functional.conv2d` in fp32 is the arithmetic definition here too.

Parity pinning: the reference ships no tests and no reproducible golden vectors (its five PNG
goldens need an external weight download).  This oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF, run in the build container by tests/golden/make_golden.py (which imports
/root/reference/codes with an in-memory GaussianNoise ctor shim) and committed under
tests/golden/*.npz; tests/test_oracle.py checks this file against those fixtures bit-for-bit
(max-abs tolerance 0 on CPU, see the test).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def lrelu(x: Tensor) -> Tensor:
    """block.py:19-20 — nn.LeakyReLU(neg_slope=0.2)."""
    return F.leaky_relu(x, 0.2)


def conv(x: Tensor, sd: SD, key: str, stride: int = 1, pad: Optional[int] = None) -> Tensor:
    """block.py:137-138 — nn.Conv2d with zero padding (k-1)//2 (block.py:55-58)."""
    w = sd[key + ".weight"]
    b = sd.get(key + ".bias")
    if pad is None:
        pad = (w.shape[-1] - 1) // 2
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def gaussian_noise(x: Tensor, training: bool, noise: Optional[Tensor], sigma: float = 0.1) -> Tensor:
    """block.py:117-122 — y = x + N(0,1) * (sigma * x) in training mode, identity in eval.

    `noise` is the N(0,1) draw (same shape as x); the reference samples it from torch's global
    generator (`normal_()`), the oracle takes it as an argument so both sides of a parity test can
    be fed the same tensor.  training=True with noise=None draws from the global generator exactly
    like the reference (`torch.empty_like(x).normal_()` consumes the stream identically to
    `self.noise.repeat(*x.size()).normal_()`).
    """
    if not training or sigma == 0:
        return x
    if noise is None:
        noise = torch.empty_like(x).normal_()
    return x + noise * (sigma * x)


def rdb_forward(x: Tensor, sd: SD, prefix: str, training: bool = False,
                noise: Optional[Tensor] = None) -> Tensor:
    """block.py:260-268 — ResidualDenseBlock_5C.forward (ESRGAN+ variant with conv1x1 and x4+x2)."""
    x1 = lrelu(conv(x, sd, prefix + "conv1.0"))
    x2 = lrelu(conv(torch.cat((x, x1), 1), sd, prefix + "conv2.0"))
    x2 = x2 + F.conv2d(x, sd[prefix + "conv1x1.weight"])          # block.py:263, bias-free 1x1
    x3 = lrelu(conv(torch.cat((x, x1, x2), 1), sd, prefix + "conv3.0"))
    x4 = lrelu(conv(torch.cat((x, x1, x2, x3), 1), sd, prefix + "conv4.0"))
    x4 = x4 + x2                                                   # block.py:266
    x5 = conv(torch.cat((x, x1, x2, x3, x4), 1), sd, prefix + "conv5.0")  # no act in CNA mode (:253-258)
    return gaussian_noise(x5 * 0.2 + x, training, noise)           # block.py:268


def rrdb_forward(x: Tensor, sd: SD, prefix: str, training: bool = False, noises=None) -> Tensor:
    """block.py:287-291 — RRDB.forward: RDB3(RDB2(RDB1(x))) * 0.2 + x."""
    out = x
    for r in (1, 2, 3):
        nz = None if noises is None else noises[r - 1]
        out = rdb_forward(out, sd, f"{prefix}RDB{r}.", training, nz)
    return out * 0.2 + x


def rrdbnet_forward(x: Tensor, sd: SD, nb: int, training: bool = False, noises=None) -> Tensor:
    """architecture.py:47-78 — RRDBNet.forward for upscale=4, upsample_mode='upconv', norm None.

    Key layout (sequential() flattening, block.py:95-108): model.0 fea_conv; model.1.sub.{i} RRDBs;
    model.1.sub.{nb} LR_conv; model.3 / model.6 upconv convs; model.8 HR_conv0; model.10 HR_conv1.
    `noises`: optional list of nb lists of 3 N(0,1) tensors (training mode only).
    """
    fea = conv(x, sd, "model.0")                                   # architecture.py:55
    t = fea
    for i in range(nb):
        nz = None if noises is None else noises[i]
        t = rrdb_forward(t, sd, f"model.1.sub.{i}.", training, nz)
    t = fea + conv(t, sd, f"model.1.sub.{nb}")                     # ShortcutBlock, block.py:84-86
    for key in ("model.3", "model.6"):                             # upconv_blcok, block.py:315-322
        t = F.interpolate(t, scale_factor=2, mode="nearest")
        t = lrelu(conv(t, sd, key))
    t = lrelu(conv(t, sd, "model.8"))                              # HR_conv0, architecture.py:70
    return conv(t, sd, "model.10")                                 # HR_conv1, architecture.py:71


def n_rrdb_blocks(sd: SD) -> int:
    idx = {int(k.split(".")[3]) for k in sd if k.startswith("model.1.sub.") and ".RDB1." in k}
    return len(idx)


# features.{conv_idx}: (stride, bn_idx or None) — architecture.py:93-119 after sequential() flattening
_D_LAYOUT = [(0, 1, None), (2, 2, 3), (5, 1, 6), (8, 2, 9), (11, 1, 12), (14, 2, 15), (17, 1, 18),
             (20, 2, 21), (23, 1, 24), (26, 2, 27)]


def discriminator_vgg128_forward(x: Tensor, sd: SD, training: bool = False,
                                 momentum: float = 0.1, eps: float = 1e-5):
    """architecture.py:87-129 — Discriminator_VGG_128.forward.

    Returns (logits [B,1], new_buffers): in training mode BatchNorm2d (block.py:32) uses biased batch
    statistics for normalisation and updates running_mean / running_var (unbiased) with
    momentum 0.1; new_buffers maps the updated buffer keys to their new values.
    """
    new_buffers = {}
    t = x
    for conv_idx, stride, bn_idx in _D_LAYOUT:
        key = f"features.{conv_idx}"
        k = sd[key + ".weight"].shape[-1]
        t = conv(t, sd, key, stride=stride, pad=(k - 1) // 2)      # k=4 -> pad 1 (block.py:55-58)
        if bn_idx is not None:
            bkey = f"features.{bn_idx}"
            g, b = sd[bkey + ".weight"], sd[bkey + ".bias"]
            if training:
                mean = t.mean(dim=(0, 2, 3))
                var = t.var(dim=(0, 2, 3), unbiased=False)
                n = t.numel() // t.shape[1]
                new_buffers[bkey + ".running_mean"] = (1 - momentum) * sd[bkey + ".running_mean"] + momentum * mean
                new_buffers[bkey + ".running_var"] = (1 - momentum) * sd[bkey + ".running_var"] + momentum * var * n / max(n - 1, 1)
                new_buffers[bkey + ".num_batches_tracked"] = sd[bkey + ".num_batches_tracked"] + 1
            else:
                mean, var = sd[bkey + ".running_mean"], sd[bkey + ".running_var"]
            t = (t - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
            t = t * g[None, :, None, None] + b[None, :, None, None]
        t = lrelu(t)
    t = t.reshape(t.shape[0], -1)                                   # architecture.py:127
    t = lrelu(F.linear(t, sd["classifier.0.weight"], sd["classifier.0.bias"]))
    t = F.linear(t, sd["classifier.2.weight"], sd["classifier.2.bias"])
    return t, new_buffers



# torchvision.models.vgg19 configuration 'E' (the reference takes `model.features` of it, architecture.py:289,298):
# out-channels per conv, 'M' = MaxPool2d(2, 2); every conv is 3x3 / stride 1 / pad 1 followed by ReLU(inplace)
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]


def vgg19_feature_layout(feature_layer: int = 34):
    """[(index in `features`, kind, cin, cout)] of torchvision's vgg19().features[:feature_layer + 1]
    (architecture.py:298; feature_layer = 34 is conv5_4 BEFORE its ReLU, networks.py:144-148)."""
    out, idx, cin = [], 0, 3
    for v in VGG19_CFG:
        if v == "M":
            out.append((idx, "pool", cin, cin))
            idx += 1
        else:
            out.append((idx, "conv", cin, v))
            out.append((idx + 1, "relu", v, v))
            idx += 2
            cin = v
    return [e for e in out if e[0] <= feature_layer]


def vgg_feature_forward(x: Tensor, sd: SD, feature_layer: int = 34, use_input_norm: bool = True) -> Tensor:
    """VGGFeatureExtractor.forward (architecture.py:303-307): (x - mean) / std, then features[:feature_layer + 1]."""
    if use_input_norm:
        x = (x - sd["mean"]) / sd["std"]                      # architecture.py:304-305
    for idx, kind, _cin, _cout in vgg19_feature_layout(feature_layer):
        if kind == "conv":
            x = F.conv2d(x, sd[f"features.{idx}.weight"], sd[f"features.{idx}.bias"], stride=1, padding=1)
        elif kind == "relu":
            x = F.relu(x)
        else:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
    return x


def synth_state_dict_vgg(feature_layer: int = 34, seed: int = 0) -> SD:
    """Deterministic synthetic VGGFeatureExtractor weights (the pretrained torchvision file cannot be fetched here):
    kaiming-uniform-like convs so that activations keep their scale through 16 layers, small biases of both signs."""
    import numpy as np
    rng = np.random.default_rng(seed + 4241)
    sd: SD = {"mean": torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1),      # architecture.py:292
              "std": torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)}        # architecture.py:294
    for idx, kind, cin, cout in vgg19_feature_layout(feature_layer):
        if kind != "conv":
            continue
        bound = (6.0 / (cin * 9)) ** 0.5
        sd[f"features.{idx}.weight"] = torch.from_numpy(rng.uniform(-bound, bound, (cout, cin, 3, 3)).astype("float32"))
        sd[f"features.{idx}.bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, (cout,)).astype("float32"))
    return sd


def psnr_255(a: Tensor, b: Tensor) -> float:
    """utils/util.py:107-114 — calculate_psnr on [0,255] images (here: clamp to [0,1], scale)."""
    a = a.clamp(0, 1) * 255.0
    b = b.clamp(0, 1) * 255.0
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    import math
    return 20 * math.log10(255.0 / math.sqrt(mse))


def synth_state_dict_g(in_nc: int, out_nc: int, nf: int, nb: int, gc: int = 32, seed: int = 0,
                       scale: float = 1.0) -> SD:
    """Deterministic synthetic generator weights (numpy PCG64; stable across torch versions), with the
    reference's key set/shapes and torch-default-like magnitude U(-1,1)/sqrt(fan_in)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd: SD = {}

    def add_conv(key, cout, cin, k, bias=True):
        bound = scale / (cin * k * k) ** 0.5
        sd[key + ".weight"] = torch.from_numpy(rng.uniform(-bound, bound, (cout, cin, k, k)).astype("float32"))
        if bias:
            sd[key + ".bias"] = torch.from_numpy(rng.uniform(-bound, bound, (cout,)).astype("float32"))

    add_conv("model.0", nf, in_nc, 3)
    for i in range(nb):
        for r in (1, 2, 3):
            p = f"model.1.sub.{i}.RDB{r}."
            add_conv(p + "conv1x1", gc, nf, 1, bias=False)
            for k in range(1, 5):
                add_conv(p + f"conv{k}.0", gc, nf + (k - 1) * gc, 3)
            add_conv(p + "conv5.0", nf, nf + 4 * gc, 3)
    add_conv(f"model.1.sub.{nb}", nf, nf, 3)
    for key in ("model.3", "model.6", "model.8"):
        add_conv(key, nf, nf, 3)
    add_conv("model.10", out_nc, nf, 3)
    return sd


def synth_state_dict_d(in_nc: int = 3, base_nf: int = 64, seed: int = 0) -> SD:
    """Deterministic synthetic Discriminator_VGG_128 weights incl. non-trivial BN affine/running stats."""
    import numpy as np
    rng = np.random.default_rng(seed + 7919)
    sd: SD = {}
    chans = [(in_nc, base_nf, 3), (base_nf, base_nf, 4), (base_nf, base_nf * 2, 3), (base_nf * 2, base_nf * 2, 4),
             (base_nf * 2, base_nf * 4, 3), (base_nf * 4, base_nf * 4, 4), (base_nf * 4, base_nf * 8, 3),
             (base_nf * 8, base_nf * 8, 4), (base_nf * 8, base_nf * 8, 3), (base_nf * 8, base_nf * 8, 4)]
    for (conv_idx, _stride, bn_idx), (cin, cout, k) in zip(_D_LAYOUT, chans):
        bound = (6.0 / (cin * k * k)) ** 0.5 * 0.5
        sd[f"features.{conv_idx}.weight"] = torch.from_numpy(rng.uniform(-bound, bound, (cout, cin, k, k)).astype("float32"))
        sd[f"features.{conv_idx}.bias"] = torch.from_numpy(rng.uniform(-0.1, 0.1, (cout,)).astype("float32"))
        if bn_idx is not None:
            sd[f"features.{bn_idx}.weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, (cout,)).astype("float32"))
            sd[f"features.{bn_idx}.bias"] = torch.from_numpy(rng.uniform(-0.2, 0.2, (cout,)).astype("float32"))
            sd[f"features.{bn_idx}.running_mean"] = torch.from_numpy(rng.uniform(-0.1, 0.1, (cout,)).astype("float32"))
            sd[f"features.{bn_idx}.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, (cout,)).astype("float32"))
            sd[f"features.{bn_idx}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    b = 1.0 / (base_nf * 8 * 16) ** 0.5
    sd["classifier.0.weight"] = torch.from_numpy(rng.uniform(-b, b, (100, base_nf * 8 * 16)).astype("float32"))
    sd["classifier.0.bias"] = torch.from_numpy(rng.uniform(-b, b, (100,)).astype("float32"))
    sd["classifier.2.weight"] = torch.from_numpy(rng.uniform(-0.1, 0.1, (1, 100)).astype("float32"))
    sd["classifier.2.bias"] = torch.from_numpy(rng.uniform(-0.1, 0.1, (1,)).astype("float32"))
    return sd

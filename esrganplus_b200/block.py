"""Parameter-holding building blocks of the generator / discriminator module trees.

Host-side mirror of the reference's ``codes/models/modules/block.py`` (and ``test_image/block.py``):
the class names, attribute names and nesting are part of the state_dict wire format
(``model.1.sub.{i}.RDB{r}.conv{k}.0.weight`` ...), of ``init_weights``' class-name dispatch
(networks.py:30-44 matches 'Conv' / 'Linear' / 'BatchNorm2d' substrings, so every leaf that owns
weights is a stock ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear``) and of ``str(net)`` logging
(base_model.py:46).  None of these modules computes anything on its own: arithmetic happens in
libesrp.so, driven by the top-level classes in ``architecture.py``.  Calling ``forward`` on an
inner block raises, it never falls back to torch math.
"""
from __future__ import annotations

from typing import List, Optional

import torch.nn as nn

NEG_SLOPE = 0.2  # block.py:12-20: LeakyReLU(neg_slope=0.2, inplace=True) everywhere on this path


def layer_group(in_nc: int, out_nc: int, kernel_size: int = 3, stride: int = 1, bias: bool = True,
                norm_type: Optional[str] = None, act_type: Optional[str] = None) -> List[nn.Module]:
    """Conv2d [+ BatchNorm2d] [+ LeakyReLU] as a flat module list ('CNA' order, zero padding
    (k-1)//2 — block.py:55-58,125-151).  Only what RRDBNet / Discriminator_VGG_128 reach is
    supported; anything else raises like the reference does for unknown layer names."""
    mods: List[nn.Module] = [nn.Conv2d(in_nc, out_nc, kernel_size, stride=stride,
                                       padding=(kernel_size - 1) // 2, bias=bias)]
    if norm_type:
        if norm_type.lower() != "batch":
            raise NotImplementedError("normalization layer [{:s}] is not found".format(norm_type))
        mods.append(nn.BatchNorm2d(out_nc, affine=True))
    if act_type:
        if act_type.lower() != "leakyrelu":
            raise NotImplementedError("activation layer [{:s}] is not found".format(act_type))
        mods.append(nn.LeakyReLU(NEG_SLOPE, True))
    return mods


class _NoStandaloneForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is a parameter holder; run the enclosing RRDBNet / "
            "Discriminator_VGG_128 (its forward drives the fused sm_100a kernels)")


class GaussianNoise(_NoStandaloneForward):
    """block.py:110-122.  Multiplicative Gaussian noise y = x * (1 + sigma * N(0,1)) in training
    mode.  Holds no tensor at all (the reference pins a 0-dim tensor to 'cuda', :115, which is why
    it cannot be built without a driver); the samples come from Philox inside the conv5 epilogue."""

    def __init__(self, sigma: float = 0.1, is_relative_detach: bool = False):
        super().__init__()
        self.sigma = sigma
        self.is_relative_detach = is_relative_detach

    def extra_repr(self) -> str:
        return ""


class ResidualDenseBlock_5C(_NoStandaloneForward):
    """block.py:232-268: conv1x1 (bias-free) + five 3x3 convs, growth gc."""

    def __init__(self, nc: int, kernel_size: int = 3, gc: int = 32, stride: int = 1, bias: bool = True,
                 pad_type: str = "zero", norm_type: Optional[str] = None, act_type: str = "leakyrelu",
                 mode: str = "CNA", gaussian_noise: bool = True):
        super().__init__()
        if kernel_size != 3 or stride != 1 or pad_type != "zero" or norm_type or mode != "CNA":
            raise NotImplementedError("ResidualDenseBlock_5C: only k3 s1 zero-pad CNA without norm is on the hot path")
        self.noise = GaussianNoise() if gaussian_noise else None
        self.conv1x1 = nn.Conv2d(nc, gc, kernel_size=1, stride=1, bias=False)
        for k in range(1, 5):
            setattr(self, f"conv{k}", nn.Sequential(*layer_group(nc + (k - 1) * gc, gc, 3, bias=bias, act_type=act_type)))
        self.conv5 = nn.Sequential(*layer_group(nc + 4 * gc, nc, 3, bias=bias, act_type=None))


class RRDB(_NoStandaloneForward):
    """block.py:271-291: three dense blocks, out * 0.2 + x."""

    def __init__(self, nc: int, kernel_size: int = 3, gc: int = 32, stride: int = 1, bias: bool = True,
                 pad_type: str = "zero", norm_type: Optional[str] = None, act_type: str = "leakyrelu",
                 mode: str = "CNA", rrdb_noise: bool = False):
        super().__init__()
        for r in (1, 2, 3):
            setattr(self, f"RDB{r}", ResidualDenseBlock_5C(nc, kernel_size, gc, stride, bias, pad_type,
                                                           norm_type, act_type, mode))
        if rrdb_noise:  # test_image/block.py:250 — parameter-free, identity in eval()
            self.noise = GaussianNoise()


class ShortcutBlock(_NoStandaloneForward):
    """block.py:78-92: x + sub(x); the attribute name `sub` is part of the key format."""

    def __init__(self, submodule: nn.Module):
        super().__init__()
        self.sub = submodule

    def __repr__(self) -> str:
        return "Identity + \n|" + repr(self.sub).replace("\n", "\n|")

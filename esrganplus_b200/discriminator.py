"""Discriminator_VGG_128 execution (architecture.py:87-129 of the reference).  Kernels land here."""
from __future__ import annotations


def discriminator_apply(module, x):
    raise NotImplementedError("esrganplus_b200: Discriminator_VGG_128 kernels are not built yet")

"""esrganplus_b200 — B200-native (sm_100a) hot path of ESRGAN+ / nESRGAN+.

Public surface = the reference's network classes (see architecture.py) over libesrp.so
(C ABI in include/esrp.h).  Importing the package does not load the shared library; the first
forward does, and raises if it is missing.
"""
from .architecture import Discriminator_VGG_128, RRDB_Net, RRDBNet, VGGFeatureExtractor  # noqa: F401

__all__ = ["RRDBNet", "RRDB_Net", "Discriminator_VGG_128", "VGGFeatureExtractor"]

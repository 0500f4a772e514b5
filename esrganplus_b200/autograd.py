"""Training-mode (autograd) entry of the generator.  Backward kernels land here."""
from __future__ import annotations


def generator_apply(module, x, params):
    raise NotImplementedError(
        "esrganplus_b200: the generator backward pass (dgrad/wgrad kernels) is not built yet; "
        "run under torch.no_grad() / with requires_grad=False parameters for inference")

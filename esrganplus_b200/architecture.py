"""Drop-in network classes: same constructor signatures, module tree and state_dict keys as the
reference's ``codes/models/modules/architecture.py`` (RRDBNet :47-78, Discriminator_VGG_128 :87-129)
and ``test_image/architecture.py`` (RRDB_Net :7-38), with ``forward`` executed by the hand-written
sm_100a kernels in libesrp.so instead of nn.Conv2d / torch.cat / elementwise ATen ops.

``networks.define_G`` / ``define_D`` (networks.py:96-99,117-119), ``train.py``, ``test.py`` and
``test_image/test.py`` bind to these names; see INTEGRATION.md for the two-line shim.

There is no torch-math or CPU fallback: inputs must be CUDA fp32 NCHW tensors, and a missing
libesrp.so raises at the first forward.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import block as B


def _flat(*groups) -> nn.Sequential:
    """Concatenate module lists into ONE nn.Sequential: the flat indices are what produce the
    reference's key numbers (model.0/1/3/6/8/10, features.0/2/3/5/6/...; block.py:95-108)."""
    mods: List[nn.Module] = []
    for g in groups:
        mods.extend(g if isinstance(g, (list, tuple)) else [g])
    return nn.Sequential(*mods)


_LIVE_MODULES = None      # weakref.WeakSet of constructed native modules
_OPT_MODULES = {}         # id(optimizer) -> (weakref to the optimizer, [weakrefs to the native modules it updates])


def _after_optimizer_step(optimizer, args, kwargs):
    """Global ``torch.optim`` post-step hook: whatever optimizer just updated parameters of a native module, the module's
    derived weight cache is stale.  Needed because not every optimizer moves the tensors' version counters —
    ``torch.optim.Adam(fused=True)`` updates through ``torch._fused_adam_`` and leaves ``_version`` untouched (torch 2.11),
    so round 1's version-only check silently kept running the first step's weights."""
    import weakref
    ent = _OPT_MODULES.get(id(optimizer))
    if ent is None or ent[0]() is not optimizer:
        ids = {id(p) for g in optimizer.param_groups for p in g["params"]}
        mods = [weakref.ref(m) for m in list(_LIVE_MODULES or ()) if any(id(p) in ids for p in m.parameters())]
        try:
            ent = (weakref.ref(optimizer), mods)
        except TypeError:
            ent = (lambda: None, mods)
        if len(_OPT_MODULES) > 64:
            _OPT_MODULES.clear()
        _OPT_MODULES[id(optimizer)] = ent
    for r in ent[1]:
        m = r()
        if m is not None:
            m.invalidate_weights()


def _register_native_module(module) -> None:
    global _LIVE_MODULES
    import weakref
    if _LIVE_MODULES is None:
        _LIVE_MODULES = weakref.WeakSet()
        from torch.optim.optimizer import register_optimizer_step_post_hook
        register_optimizer_step_post_hook(_after_optimizer_step)
    _LIVE_MODULES.add(module)
    _OPT_MODULES.clear()   # (an optimizer seen before may cover this module as well)


class _NativeWeights:
    """Mixin of the top-level classes: the native engines keep DERIVED bf16 weight tiles and repack them when a
    Parameter's storage pointer or in-place version counter changes.  Writes through ``.data`` bump no version
    counter (``m.weight.data *= scale`` / ``init.kaiming_normal_(m.weight.data)`` in the reference's ``init_weights``,
    networks.py:30-44, run via ``net.apply(fn)``), so every bulk entry point that may hide such writes bumps an
    epoch the engines compare as well: ``apply``, ``_apply`` (``.to()`` / ``.cuda()``), ``load_state_dict``, and every
    ``torch.optim`` optimizer step that covers one of the module's parameters (a global post-step hook:
    ``Adam(fused=True)`` bumps no version counter either).  After any other ``.data`` write (EMA, weight interpolation,
    clipping) call ``invalidate_weights()``."""

    def invalidate_weights(self) -> None:
        object.__setattr__(self, "_esrp_epoch", self.__dict__.get("_esrp_epoch", 0) + 1)

    @property
    def weights_epoch(self) -> int:
        return self.__dict__.get("_esrp_epoch", 0)

    def apply(self, fn):
        r = super().apply(fn)
        self.invalidate_weights()
        return r

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        self.invalidate_weights()
        return r

    def load_state_dict(self, *args, **kwargs):
        r = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return r


class _GeneratorBase(_NativeWeights, nn.Module):
    def _build(self, in_nc, out_nc, nf, nb, gc, upscale, norm_type, act_type, mode, upsample_mode, rrdb_noise):
        if upsample_mode not in ("upconv", "pixelshuffle"):
            raise NotImplementedError("upsample mode [{:s}] is not found".format(upsample_mode))
        if upsample_mode != "upconv" or upscale == 3 or norm_type or (act_type or "").lower() != "leakyrelu":
            # accepted by the reference's signature but never selected by the ESRGAN+/nESRGAN+ configs
            # (options/train/train_ESRGANplus.json, test_image/test.py:15-16); no kernels exist for them
            raise NotImplementedError(
                "esrganplus_b200 implements the ESRGAN+ hot path only: upsample_mode='upconv', "
                "upscale in {1,2,4}, norm_type=None, act_type='leakyrelu'")
        n_up = int(math.log(upscale, 2))
        if 2 ** n_up != upscale:
            raise NotImplementedError(f"upscale={upscale} unsupported")
        # architecture.py:56 passes the literal gc=32 to every RRDB whatever `gc` says; keep that.
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, nb=nb, gc=32, upscale=upscale)
        trunk = nn.Sequential(*[B.RRDB(nf, 3, 32, 1, True, "zero", None, act_type, "CNA", rrdb_noise=rrdb_noise)
                                for _ in range(nb)],
                              *B.layer_group(nf, nf, 3))                       # LR_conv, no act (:58)
        ups = []
        for _ in range(n_up):                                                   # block.py:315-322
            ups += [nn.Upsample(scale_factor=2, mode="nearest")] + B.layer_group(nf, nf, 3, act_type=act_type)
        self.model = _flat(B.layer_group(in_nc, nf, 3),                         # fea_conv (:55)
                           B.ShortcutBlock(trunk), ups,
                           B.layer_group(nf, nf, 3, act_type=act_type),         # HR_conv0 (:70)
                           B.layer_group(nf, out_nc, 3))                        # HR_conv1 (:71)
        # native engines, one per device; kept out of the module/state machinery on purpose
        object.__setattr__(self, "_engines", {})
        object.__setattr__(self, "_step", 0)
        _register_native_module(self)

    # -- engine plumbing -----------------------------------------------------------------------
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engines"] = {}
        st.pop("_precise_engines", None)
        return st

    def _engine_for(self, device: torch.device):
        from .engine import GeneratorEngine
        eng = self._engines.get(device)
        if eng is None:
            c = self.cfg
            eng = GeneratorEngine(c["in_nc"], c["out_nc"], c["nf"], c["nb"], c["gc"], c["upscale"], device)
            self._engines[device] = eng
        return eng

    def noise_seed(self) -> int:
        """Philox key for this forward: torch's global seed mixed with a per-module call counter, so
        `torch.manual_seed` reproduces a training run (the reference draws from the global stream), and with the
        data-parallel rank: replicas seeded identically must not draw the same noise for their different samples
        (per-rank Philox offset = f(rank, step))."""
        object.__setattr__(self, "_step", self._step + 1)
        rank = 0
        if getattr(self, "_dp_group", False) is not False:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                g = self._dp_group
                rank = dist.get_rank(None if g is True else g)
        return (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._step + rank * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF

    def forward_uint8(self, img: torch.Tensor, bgr: bool = True, fp32_parity: bool = False) -> torch.Tensor:
        """Not part of the reference surface: 8-bit images in, 8-bit images out, with the host
        plumbing of test_image/test.py:31-40 (``/255``, BGR<->RGB, ``clamp_(0,1)``, ``(x*255).round()``) done on the
        device — a 4x smaller device->host copy and no numpy pass.  img: uint8 [n,h,w,in_nc] CUDA tensor (cv2 order when
        `bgr`); always the eval-mode network, no gradient.  `fp32_parity=True` runs the split-precision forward
        (forward_fp32_parity) inside the same plumbing."""
        if not img.is_cuda:
            raise RuntimeError("esrganplus_b200.RRDBNet runs on CUDA (sm_100a) only")
        if fp32_parity:
            # the same plumbing around the split-precision forward (the result then equals the reference's 8-bit image);
            # torch ops: this is the accuracy mode, not the fast one
            if img.dim() != 4 or img.dtype != torch.uint8:
                raise RuntimeError("forward_uint8 expects a uint8 [n,h,w,c] tensor")
            x = img.flip(-1) if bgr else img
            x = (x.permute(0, 3, 1, 2).to(torch.float64) * 1.0 / 255).float().contiguous()   # test_image/test.py:31-33
            y = self.forward_fp32_parity(x).clamp_(0, 1)                                      # :37
            y = (y.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8)                       # :38-39
            return (y.flip(-1) if bgr else y).contiguous()
        from .discriminator import _named_params
        names, plist = _named_params(self)
        eng = self._engine_for(img.device)
        eng.sync_weights(dict(zip(names, plist)), self.weights_epoch)
        return eng.forward_u8(img, bgr)

    def forward_fp32_parity(self, x: torch.Tensor) -> torch.Tensor:
        """Not part of the reference surface: the eval-mode forward in SPLIT PRECISION (esrganplus_b200/precise.py) — every
        value carried as two bf16 tensors, every conv as A_hi W_hi + A_lo W_hi + A_hi W_lo on the same tcgen05 kernels,
        fp32 accumulation and fp32 residuals.  Matches the reference's fp32 arithmetic to ~1e-5 of the output's spread
         at several times the cost of ``forward``.  No noise, no gradients."""
        if not x.is_cuda:
            raise RuntimeError("esrganplus_b200.RRDBNet runs on CUDA (sm_100a) only")
        if self.training:
            raise RuntimeError("forward_fp32_parity is an eval-mode path (call net.eval() first)")
        from .precise import PreciseGenerator
        engines = self.__dict__.setdefault("_precise_engines", {})
        eng = engines.get(x.device)
        if eng is None:
            eng = engines[x.device] = PreciseGenerator(self, x.device)
        with torch.no_grad():
            return eng.forward(self, x.contiguous())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("esrganplus_b200.RRDBNet runs on CUDA (sm_100a) only; there is no CPU path "
                               "(use the reference modules for CPU inference)")
        from .discriminator import _named_params
        names, plist = _named_params(self)
        params = self.__dict__.get("_esrp_params_dict")
        if params is None or params[0] is not plist:
            params = (plist, dict(zip(names, plist)))
            object.__setattr__(self, "_esrp_params_dict", params)
        params = params[1]
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in plist))
        if needs_grad:
            from .autograd import generator_apply
            return generator_apply(self, x, params)
        eng = self._engine_for(x.device)
        eng.sync_weights(params, self.weights_epoch)
        return eng.forward(x, self.training, self.noise_seed() if self.training else 0)


class RRDBNet(_GeneratorBase):
    """architecture.py:47-78.  RRDBNet(in_nc, out_nc, nf, nb, gc=32, upscale=4, norm_type=None,
    act_type='leakyrelu', mode='CNA', upsample_mode='upconv')."""

    def __init__(self, in_nc, out_nc, nf, nb, gc=32, upscale=4, norm_type=None, act_type="leakyrelu",
                 mode="CNA", upsample_mode="upconv"):
        super().__init__()
        self._build(in_nc, out_nc, nf, nb, gc, upscale, norm_type, act_type, mode, upsample_mode, rrdb_noise=False)


class RRDB_Net(_GeneratorBase):
    """test_image/architecture.py:7-38 — the standalone inference copy (extra, unused `res_scale`;
    its RRDB carries a second, parameter-free GaussianNoise that is the identity in eval())."""

    def __init__(self, in_nc, out_nc, nf, nb, gc=32, upscale=4, norm_type=None, act_type="leakyrelu",
                 mode="CNA", res_scale=1, upsample_mode="upconv"):
        super().__init__()
        self._build(in_nc, out_nc, nf, nb, gc, upscale, norm_type, act_type, mode, upsample_mode, rrdb_noise=True)

    def forward(self, x):
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("RRDB_Net is the inference copy (test_image/test.py); train with RRDBNet")
        return super().forward(x)


class Discriminator_VGG_128(_NativeWeights, nn.Module):
    """architecture.py:87-129.  Ten conv layers (k3 s1 / k4 s2 alternating, BatchNorm2d from the second
    on, LeakyReLU 0.2) 128x128 -> 4x4, then Linear(8192,100) + LeakyReLU + Linear(100,1)."""

    def __init__(self, in_nc, base_nf, norm_type="batch", act_type="leakyrelu", mode="CNA"):
        super().__init__()
        if mode != "CNA":
            raise NotImplementedError("Discriminator_VGG_128: only mode='CNA' is on the hot path")
        nf = base_nf
        plan = [(in_nc, nf, 3, 1, None), (nf, nf, 4, 2, norm_type),
                (nf, 2 * nf, 3, 1, norm_type), (2 * nf, 2 * nf, 4, 2, norm_type),
                (2 * nf, 4 * nf, 3, 1, norm_type), (4 * nf, 4 * nf, 4, 2, norm_type),
                (4 * nf, 8 * nf, 3, 1, norm_type), (8 * nf, 8 * nf, 4, 2, norm_type),
                (8 * nf, 8 * nf, 3, 1, norm_type), (8 * nf, 8 * nf, 4, 2, norm_type)]
        self.features = _flat(*[B.layer_group(ci, co, k, s, norm_type=nt, act_type=act_type)
                                for ci, co, k, s, nt in plan])
        self.classifier = nn.Sequential(nn.Linear(512 * 4 * 4, 100), nn.LeakyReLU(0.2, True), nn.Linear(100, 1))
        object.__setattr__(self, "_engines", {})
        _register_native_module(self)

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engines"] = {}
        return st

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("esrganplus_b200.Discriminator_VGG_128 runs on CUDA (sm_100a) only")
        from .discriminator import discriminator_apply
        return discriminator_apply(self, x)


# torchvision.models.vgg19 configuration 'E' (architecture.py:289): out-channels per conv, 'M' = MaxPool2d(2, 2)
_VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]


class VGGFeatureExtractor(_NativeWeights, nn.Module):
    """architecture.py:279-307: `(x - mean) / std`, then torchvision's vgg19().features[:feature_layer + 1] with frozen
    weights (feature_layer = 34: conv5_4 before its ReLU, networks.py:144-155).  Same constructor arguments, module tree
    (`features` = Sequential of Conv2d / ReLU(inplace) / MaxPool2d at torchvision's indices), buffers (`mean`, `std`) and
    state_dict keys as the reference, so `networks.define_F` builds it unchanged and a torchvision vgg19 checkpoint loads
    through `load_vgg19_state_dict`.

    One difference is unavoidable offline: the reference constructor downloads the pretrained torchvision weights
    (`pretrained=True`, :289).  This class does that only when `pretrained=True` is passed AND torchvision can supply
    them; by default the convs are torchvision-initialised (kaiming normal) and the caller loads weights.
    `use_bn=True` (vgg19_bn, feature_layer 49) is not on the ESRGAN+ path (networks.py:58 passes use_bn=False)."""

    def __init__(self, feature_layer=34, use_bn=False, use_input_norm=True, device=torch.device("cpu"), pretrained=False):
        super().__init__()
        if use_bn:
            raise NotImplementedError("VGGFeatureExtractor: use_bn=True (vgg19_bn) is not on the ESRGAN+ path")
        self.use_input_norm = use_input_norm
        if self.use_input_norm:
            self.register_buffer("mean", torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1).to(device))
            self.register_buffer("std", torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1).to(device))
        layers, cin = [], 3
        for v in _VGG19_CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                conv = nn.Conv2d(cin, v, kernel_size=3, padding=1)
                nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")   # torchvision/models/vgg.py
                nn.init.constant_(conv.bias, 0)
                layers += [conv, nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers[:feature_layer + 1])
        if not isinstance(self.features[-1], nn.Conv2d):
            raise NotImplementedError("VGGFeatureExtractor: feature_layer must index a conv (features before the ReLU)")
        if pretrained:
            import torchvision
            self.load_vgg19_state_dict(torchvision.models.vgg19(weights="IMAGENET1K_V1").state_dict())
        for _k, v in self.features.named_parameters():   # "No need to BP to variable" (:299-301)
            v.requires_grad = False
        object.__setattr__(self, "_engines", {})
        _register_native_module(self)

    def load_vgg19_state_dict(self, sd):
        """Load a torchvision vgg19 state_dict (`features.N.*` + `classifier.*`): the keys of the kept layers are identical."""
        own = self.state_dict()
        picked = {k: v for k, v in sd.items() if k in own}
        missing = [k for k in own if k.startswith("features.") and k not in picked]
        if missing:
            raise KeyError(f"load_vgg19_state_dict: missing {missing[:4]}...")
        self.load_state_dict({**own, **picked}, strict=True)

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engines"] = {}
        return st

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("esrganplus_b200.VGGFeatureExtractor runs on CUDA (sm_100a) only")
        from .vgg_feature import feature_apply
        return feature_apply(self, x)

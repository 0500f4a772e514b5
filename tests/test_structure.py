"""Host-side contract of the drop-in classes (no GPU): state_dict keys/shapes/dtypes and str(net)
against dumps taken from the reference (tests/golden/structure.json), init_weights-style dispatch
(networks.py:30-70), requires_grad toggling, strict loading, and the C-ABI symbol table."""
import copy
import ctypes
import functools
import hashlib
import json
import os
import re

import pytest
import torch
import torch.nn as nn
from torch.nn import init

import esrganplus_b200 as E
from esrganplus_b200 import _lib
from oracle import esrgan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def struct(golden_dir):
    return json.load(open(os.path.join(golden_dir, "structure.json")))


def _dump(net):
    return [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]


def test_generator_keys_match_reference(struct):
    g = E.RRDBNet(3, 3, 64, 23, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA",
                  upsample_mode="upconv")
    assert _dump(g) == struct["G_nb23_nf64"]
    assert len(g.state_dict()) == 771
    assert sum(p.numel() for p in g.parameters()) == struct["G_num_params"] == 16839299
    assert hashlib.sha256(str(g).encode()).hexdigest() == struct["G_repr_sha256"]
    # test_image/test.py:15-16 call form (positional + res_scale)
    t = E.RRDB_Net(3, 3, 64, 23, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA",
                   res_scale=1, upsample_mode="upconv")
    assert _dump(t) == struct["G_nb23_nf64"]


def test_gc_argument_is_ignored_like_reference():
    # architecture.py:56 passes literal gc=32
    g = E.RRDBNet(3, 3, 64, 1, gc=16)
    assert g.state_dict()["model.1.sub.0.RDB1.conv1.0.weight"].shape == (32, 64, 3, 3)


def test_discriminator_keys_match_reference(struct):
    d = E.Discriminator_VGG_128(3, 64, norm_type="batch", act_type="leakyrelu", mode="CNA")
    assert _dump(d) == struct["D_vgg128"]
    assert sum(p.numel() for p in d.parameters()) == struct["D_num_params"]
    assert hashlib.sha256(str(d).encode()).hexdigest() == struct["D_repr_sha256"]


def test_unknown_upsample_mode_raises():
    with pytest.raises(NotImplementedError):
        E.RRDBNet(3, 3, 64, 1, upsample_mode="bogus")


def test_strict_load_and_roundtrip():
    sd = O.synth_state_dict_g(3, 3, 32, 2, seed=5)
    net = E.RRDBNet(3, 3, 32, 2)
    net.load_state_dict(sd, strict=True)
    out = net.state_dict()
    assert list(out.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(out[k], sd[k]) and out[k].dtype == torch.float32
    d = E.Discriminator_VGG_128(3, 64)
    d.load_state_dict(O.synth_state_dict_d(3, 64, seed=1), strict=True)
    net2 = copy.deepcopy(net)
    assert net2._engines == {} and torch.equal(net2.state_dict()["model.0.weight"], sd["model.0.weight"])


def _kaiming(m, scale=1):
    # same dispatch rule as networks.py:30-44 (class-name substring)
    name = m.__class__.__name__
    if name.find("Conv") != -1 or name.find("Linear") != -1:
        init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
        m.weight.data *= scale
        if m.bias is not None:
            m.bias.data.zero_()
    elif name.find("BatchNorm2d") != -1:
        init.constant_(m.weight.data, 1.0)
        init.constant_(m.bias.data, 0.0)


def test_init_weights_dispatch_and_rng_order():
    """net.apply visits modules in the same post-order as the reference tree, so the same seed gives
    the same tensors: compare against a plain torch model with the reference's parameter order."""
    torch.manual_seed(123)
    g = E.RRDBNet(3, 3, 32, 1)
    torch.manual_seed(7)
    g.apply(functools.partial(_kaiming, scale=0.1))
    # replay: kaiming_normal_ over conv weights in module post-order == state_dict order here,
    # except conv1x1 is registered before conv1..5 (block.py:244) which state_dict order also reflects
    torch.manual_seed(7)
    for k, v in g.state_dict().items():
        if k.endswith(".weight"):
            ref = torch.empty_like(v)
            init.kaiming_normal_(ref, a=0, mode="fan_in")
            assert torch.equal(v, ref * 0.1), k
        else:
            assert torch.count_nonzero(v) == 0, k
    d = E.Discriminator_VGG_128(3, 64)
    d.apply(functools.partial(_kaiming, scale=1))
    assert torch.equal(d.state_dict()["features.3.weight"], torch.ones(64))


def test_requires_grad_toggle_and_modes():
    g = E.RRDBNet(3, 3, 32, 1)
    for _, p in g.named_parameters():
        p.requires_grad = False
    assert not any(p.requires_grad for p in g.parameters())
    assert g.train().training and not g.eval().training
    dp = nn.DataParallel(g)  # networks.py:107 wraps; base_model.py:44 unwraps via .module
    assert dp.module is g


def test_cpu_forward_raises_instead_of_falling_back():
    g = E.RRDBNet(3, 3, 32, 1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        g(torch.zeros(1, 3, 8, 8))
    d = E.Discriminator_VGG_128(3, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        d(torch.zeros(1, 3, 128, 128))


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "esrganplus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in src, f


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "esrp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(esrp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert os.path.exists(_lib.LIB_PATH), "libesrp.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/esrp.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    l = _lib.load()
    assert l.esrp_version() >= 100
    assert l.esrp_packed_conv3x3_bytes(3, 64, 32, 0) == 3 * 9 * 32 * 64 * 2
    assert l.esrp_packed_conv3x3_bytes(3, 64, 32, 1) == 3 * 12 * 32 * 64 * 2


def test_conv_desc_struct_matches_header_size():
    # field-by-field mirror of esrp_conv3x3_t; a size mismatch means the ctypes mirror drifted
    l = _lib.load()
    assert l.esrp_sizeof_conv3x3() == ctypes.sizeof(_lib.Conv3x3Desc)


def test_conv_descriptor_validation_of_optional_fields():
    """Argument checks of esrp_conv3x3_nhwc happen before anything touches a device (error behaviour of the C ABI:
    non-zero return + esrp_last_error()).  Covers the optional fields: slices / slice_stride, f32_planar, k_valid."""
    l = _lib.load()

    def desc(**kw):
        d = _lib.Conv3x3Desc()
        d.n, d.h, d.w = 1, 8, 200
        d.src[0] = 0x1000
        d.src_ctotal[0] = 64
        d.kc, d.num_chunks, d.bn, d.cout = 64, 1, 32, 32
        d.chunk_src[0], d.chunk_c0[0] = 0, 0
        d.w_packed = 0x2000
        d.w_layout = _lib.LAYOUT_ROW
        d.s0 = d.s1 = d.s2 = 1.0
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    def err(d):
        rc = l.esrp_conv3x3_nhwc(ctypes.byref(d), None)
        assert rc != 0
        return l.esrp_last_error().decode()

    assert "ESRP_LAYOUT_ROW" in err(desc(slices=2, slice_stride=1024, w_layout=_lib.LAYOUT_TILE))
    assert "slice_stride" in err(desc(slices=2, slice_stride=0))
    assert "slice_stride" in err(desc(slices=2, slice_stride=1000))
    assert "cout == bn" in err(desc(slices=2, slice_stride=1024, cout=16))
    assert "cout == bn" in err(desc(slices=5, slice_stride=1024))
    assert "out_nchw" in err(desc(slices=2, slice_stride=1024, out_nchw=0x3000))
    assert "f32_planar" in err(desc(f32_planar=1, w_layout=_lib.LAYOUT_TILE))
    assert "f32_planar" in err(desc(f32_planar=1, mask_out=0x3000, mask_out_ctotal=32))
    assert "k_valid" in err(desc(k_valid=65))
    assert "k_valid" in err(desc(k_valid=-1))
    assert "k_valid" in err(desc(num_chunks=2, k_valid=64, src_ctotal=(ctypes.c_int32 * 2)(128, 0),
                                 chunk_c0=(ctypes.c_int32 * _lib.ESRP_MAX_CHUNKS)(0, 64)))


def test_optimizer_steps_invalidate_the_derived_weight_cache():
    """torch.optim.Adam(fused=True) updates parameters without moving their version counters (torch 2.11): the engines'
    (pointer, version) signature cannot see it, the modules' invalidation epoch must (architecture._after_optimizer_step)."""
    import torch
    import esrganplus_b200 as E
    g = E.RRDBNet(3, 3, 32, 1)
    d = E.Discriminator_VGG_128(3, 64)
    for p in list(g.parameters()) + list(d.parameters()):
        p.grad = torch.zeros_like(p)
    for kw in ({"fused": True}, {}):
        opt_g = torch.optim.Adam(g.parameters(), lr=1e-4, **kw)
        e_g, e_d = g.weights_epoch, d.weights_epoch
        opt_g.step()
        assert g.weights_epoch > e_g and d.weights_epoch == e_d, kw
        opt_d = torch.optim.SGD(d.parameters(), lr=0.1)
        opt_d.step()
        assert d.weights_epoch > e_d


def test_vgg_feature_extractor_structure_matches_reference_fixture(golden_dir):
    """Keys and str(net) stored by tests/golden/make_golden_vgg.py from the reference's own VGGFeatureExtractor
    (architecture.py:279-307); frozen feature weights (:299-301); a torchvision vgg19 state_dict loads by key."""
    import numpy as np
    d = np.load(os.path.join(golden_dir, "vgg_feature_34.npz"))
    m = E.VGGFeatureExtractor(feature_layer=34, use_bn=False, use_input_norm=True, device=torch.device("cpu"))
    assert list(m.state_dict().keys()) == [str(k) for k in d["keys"]]
    assert str(m) == str(d["repr"])
    assert not any(p.requires_grad for p in m.parameters())
    tv = {k: torch.full_like(v, 0.5) for k, v in m.state_dict().items() if k.startswith("features.")}
    tv["features.35.weight"] = torch.zeros(1)          # layers beyond the cut and the classifier are ignored
    tv["classifier.0.weight"] = torch.zeros(1)
    m.load_vgg19_state_dict(tv)
    assert float(m.features[34].weight.mean()) == 0.5 and float(m.mean.flatten()[0]) == pytest.approx(0.485)
    with pytest.raises(KeyError):
        m.load_vgg19_state_dict({"features.0.weight": tv["features.0.weight"]})
    with pytest.raises(NotImplementedError):
        E.VGGFeatureExtractor(use_bn=True)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 32, 32))                    # no CPU fallback

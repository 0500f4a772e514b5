"""Diagnostic: does any pass read device memory it has not written?  The caching allocator is primed with NaN-filled
blocks (a fresh process hands out zeroed pages, which hides such reads), then the discriminator / generator passes run
and every plan-owned buffer is checked for non-finite values after the first (recording) and second (replayed) pass."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_d, random_state_dict_g

dev = torch.device("cuda:0")
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def poison():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    blocks = []
    for sz in [1 << 30] * 6 + [1 << 26] * 16 + [1 << 22] * 64 + [1 << 18] * 128 + [1 << 12] * 512 + [512] * 2048:
        blocks.append(torch.full((sz // 4,), float("nan"), device=dev))
    torch.cuda.synchronize()
    del blocks


def report(tag, eng):
    bad = 0
    for key, pool in eng.plans.items():
        for pi, pl in enumerate(pool):
            for i, t in enumerate(pl.keep):
                if isinstance(t, torch.Tensor) and t.is_floating_point():
                    nb = (~torch.isfinite(t.float())).sum().item()
                    if nb:
                        bad += 1
                        print(f"  {tag}: plan {pi} keep[{i}] {tuple(t.shape)} {t.dtype}: {nb} non-finite of {t.numel()}")
    print(f"{tag}: {bad} plan buffers with non-finite values", flush=True)


poison()
netD = E.Discriminator_VGG_128(3, 64)
netD.load_state_dict(random_state_dict_d(3, 64, seed=32), strict=True)
netD = netD.to(dev).train()
g = torch.Generator().manual_seed(5)
x = torch.rand(bs, 3, 128, 128, generator=g).to(dev)
poison()
for mode in ("nograd", "frozen_dx", "train"):
    for p in netD.parameters():
        p.requires_grad = mode == "train"
    xi = x.clone().requires_grad_(mode == "frozen_dx")
    for it in range(3):
        with torch.set_grad_enabled(mode != "nograd"):
            y = netD(xi)
        print(mode, it, "logits finite", bool(torch.isfinite(y).all()), float(y.mean()), flush=True)
        report(f"{mode} fwd {it}", netD._engines[dev])
        if mode != "nograd":
            y.mean().backward()
            if mode == "frozen_dx":
                print(mode, it, "dx finite", bool(torch.isfinite(xi.grad).all()), float(xi.grad.abs().mean()))
                xi.grad = None
            else:
                print(mode, it, "grads finite", all(bool(torch.isfinite(p.grad).all()) for p in netD.parameters()))
            report(f"{mode} bwd {it}", netD._engines[dev])
    poison()

netG = E.RRDBNet(3, 3, 64, 23)
netG.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31, scale=0.1, zero_bias=True), strict=True)
netG = netG.to(dev).train()
lr = torch.rand(bs, 3, 32, 32, generator=g).to(dev)
poison()
for it in range(3):
    y = netG(lr)
    print("G", it, "finite", bool(torch.isfinite(y).all()), float(y.abs().mean()), flush=True)
    y.mean().backward()
    print("G", it, "grads finite", all(bool(torch.isfinite(p.grad).all()) for p in netG.parameters()), flush=True)

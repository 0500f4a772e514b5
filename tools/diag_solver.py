#!/usr/bin/env python
"""Bisecting aid: losses of three GAN steps under the torch solver and under the native solver with single features off."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200 import autograd as A
from esrganplus_b200.gan_step import GanTrainStep
from oracle import esrgan_oracle as O
dev = torch.device("cuda:0")
mode = sys.argv[1]
netG = E.RRDBNet(3, 3, 32, 1); netG.load_state_dict(O.synth_state_dict_g(3, 3, 32, 1, seed=61), strict=True)
netD = E.Discriminator_VGG_128(3, 64); netD.load_state_dict(O.synth_state_dict_d(3, 64, seed=62), strict=True)
netG, netD = netG.to(dev).eval(), netD.to(dev).train()
if mode == "native_noflat":
    A.enable_flat_grads = lambda m, enable=True: m
step = GanTrainStep(netG, netD, native_solver=(mode != "torch"))
g = torch.Generator().manual_seed(9)
lr_img, hr_img = torch.rand(2, 3, 32, 32, generator=g).to(dev), torch.rand(2, 3, 128, 128, generator=g).to(dev)
out = []
p0d = [p.detach().clone() for p in netD.parameters()]
p0g = [p.detach().clone() for p in netG.parameters()]
for it in range(2):
    log = step.step(lr_img, hr_img)
    rec = {k: round(v.item(), 6) for k, v in log.items()}
    if it == 0:
        dd = torch.cat([(p.detach() - q).flatten() for p, q in zip(netD.parameters(), p0d)])
        dg = torch.cat([(p.detach() - q).flatten() for p, q in zip(netG.parameters(), p0g)])
        gd = torch.cat([p.grad.flatten() for p in netD.parameters()])
        rec.update(dD_abs_mean=dd.abs().mean().item(), dD_max=dd.abs().max().item(), dG_abs_mean=dg.abs().mean().item(), dG_max=dg.abs().max().item(),
                   gradD_norm=gd.norm().item(), gradD_absmean=gd.abs().mean().item())
        names = [k for k, _ in netD.named_parameters()]
        big = sorted(((p.detach() - q).abs().max().item(), k) for (k, p), q in zip(netD.named_parameters(), p0d))[-4:]
        rec["biggest_moves"] = big
    out.append(rec)
print(json.dumps({"mode": mode, "graph": os.environ.get("ESRP_D_GRAPH", "1"), "steps": out}))

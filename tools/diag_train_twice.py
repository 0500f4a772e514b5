"""Diagnostic: GAN steps at several batch sizes, one after another in ONE process (bench.py's weak + strong legs do
exactly that), printing the losses of every step.  usage: diag_train_twice.py 32 16 [...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.gan_step import GanTrainStep
from esrganplus_b200.synth import random_state_dict_d, random_state_dict_g

dev = torch.device("cuda:0")
for bs in [int(a) for a in sys.argv[1:]] or [32, 16]:
    netG = E.RRDBNet(3, 3, 64, 23)
    netG.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31, scale=0.1, zero_bias=True), strict=True)
    netD = E.Discriminator_VGG_128(3, 64)
    netD.load_state_dict(random_state_dict_d(3, 64, seed=32), strict=True)
    netG, netD = netG.to(dev).train(), netD.to(dev).train()
    step = GanTrainStep(netG, netD, native_solver=os.environ.get("DIAG_TORCH_SOLVER") != "1")
    calls = []
    orig = netD.forward
    def fwd(x, _o=orig, _c=calls):
        y = _o(x)
        _c.append((bool(torch.isfinite(x).all()), bool(torch.isfinite(y).all()), round(float(y.detach().mean()), 4)))
        return y
    netD.forward = fwd
    g = torch.Generator().manual_seed(100)
    lr = torch.rand(bs, 3, 32, 32, generator=g).to(dev)
    hr = torch.rand(bs, 3, 128, 128, generator=g).to(dev)
    for it in range(11):
        log = step.step(lr, hr)
        torch.cuda.synchronize()
        print(bs, it, {k: round(float(v), 5) for k, v in log.items()}, flush=True)
        if it < 2:
            print("   D calls (x finite, y finite, mean):", calls, flush=True)
        calls.clear()

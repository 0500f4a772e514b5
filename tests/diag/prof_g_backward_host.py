import sys, time
sys.path.insert(0, '/root/repo')
import torch
import esrganplus_b200 as E
from esrganplus_b200 import engine as EN
from esrganplus_b200.synth import random_state_dict_g
dev = torch.device('cuda:0')
net = E.RRDBNet(3, 3, 64, 23); net.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=1, scale=0.1)); net = net.to(dev).train()
x = torch.rand(32, 3, 32, 32, device=dev); dy = torch.randn(32, 3, 128, 128, device=dev)
orig = EN.GeneratorEngine.backward
T = {'c': 0.0, 'tot': 0.0, 'n': 0}
lib_bwd = None
def timed_backward(self, dy, token, needs):
    t0 = time.perf_counter()
    r = orig(self, dy, token, needs)
    T['tot'] += time.perf_counter() - t0; T['n'] += 1
    return r
EN.GeneratorEngine.backward = timed_backward
import ctypes
eng = None
for it in range(13):
    if it == 3:
        torch.cuda.synchronize(); T.update(c=0.0, tot=0.0, n=0); tb = 0.0; tf = 0.0
    net.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); y = net(x); t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter(); y.backward(dy); t3 = time.perf_counter()
    torch.cuda.synchronize()
    if it >= 3: tf += t1 - t0; tb += t3 - t2
print("fwd host ms", tf / 10 * 1e3, "bwd host ms (autograd total)", tb / 10 * 1e3, "engine.backward ms", T['tot'] / T['n'] * 1e3)
# C call only
eng = net._engines[dev]
import numpy as np

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_d, random_state_dict_g
dev = torch.device("cuda:0")
d = E.Discriminator_VGG_128(3, 64); d.load_state_dict(random_state_dict_d(3, 64, seed=32)); d = d.to(dev).train()
x = torch.rand(32, 3, 128, 128, device=dev, requires_grad=True)
for _ in range(2):
    d(x).sum().backward()
torch.cuda.synchronize()
torch.cuda.profiler.start()
d(x).sum().backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.gan_step import GanTrainStep
from esrganplus_b200.synth import random_state_dict_d, random_state_dict_g
dev = torch.device("cuda:0")
netG = E.RRDBNet(3, 3, 64, 23)
netG.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31, scale=0.1, zero_bias=True))
netD = E.Discriminator_VGG_128(3, 64); netD.load_state_dict(random_state_dict_d(3, 64, seed=32))
netG, netD = netG.to(dev).train(), netD.to(dev).train()
step = GanTrainStep(netG, netD)
lr = torch.rand(32, 3, 32, 32, device=dev); hr = torch.rand(32, 3, 128, 128, device=dev)
for _ in range(3): step.step(lr, hr)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step.step(lr, hr)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])

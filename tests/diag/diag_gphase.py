import sys, os, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import esrganplus_b200 as E
from oracle import esrgan_oracle as O
import test_gpu_train_d as T
dev = torch.device('cuda:0')
g = np.load('/root/repo/tests/golden/train_step_nb1_nf32.npz')
sd_d = O.synth_state_dict_d(3, 64, seed=62)
netD = E.Discriminator_VGG_128(3, 64); netD.load_state_dict(sd_d); netD = netD.to(dev).train()
for p in netD.parameters(): p.requires_grad = False
fake = torch.from_numpy(g['fake_H']).to(dev).requires_grad_(True)   # the REFERENCE's fake_H as D input
hr = torch.from_numpy(g['hr']).to(dev)
def gan(pred, real): return F.binary_cross_entropy_with_logits(pred, torch.full_like(pred, 1.0 if real else 0.0))
def gphase(dfn, fake):
    pg = dfn(fake); pr = dfn(hr).detach()
    return 5e-3 * (gan(pr - pg.mean(), False) + gan(pg - pr.mean(), True)) / 2
l = gphase(netD, fake); l.backward(); ours = fake.grad.detach().cpu()
for emu in (False, True):
    f2 = torch.from_numpy(g['fake_H']).requires_grad_(True)
    hrc = torch.from_numpy(g['hr'])
    dfn = (lambda t: T._emulated_d(t, sd_d, True)) if emu else (lambda t: O.discriminator_vgg128_forward(t, sd_d, True)[0])
    pg = dfn(f2); pr = dfn(hrc).detach()
    l2 = 5e-3 * (gan(pr - pg.mean(), False) + gan(pg - pr.mean(), True)) / 2
    l2.backward()
    print('emulated' if emu else 'fp32', 'loss', l.item(), l2.item(), 'dx rel/cos', T._rel(ours, f2.grad),
          'sum ours', ours.sum((0, 2, 3)).tolist(), 'ref', f2.grad.sum((0, 2, 3)).tolist(), 'pg', pg.flatten().tolist())

import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import esrgan_oracle as O
import test_gpu_train_d as T
dev = torch.device('cuda:0')
def prof(sd, x, r, what):
    d, y, dx = T._run(dev, sd, x, r, training=True)
    for emu in (False, True):
        ry, rdx, rg = T._ref(x, sd, r, True, emu)
        print(what, 'emulated' if emu else 'fp32', 'logits', (y - ry).abs().max().item(), 'dx', T._rel(dx, rdx))
        for k, p in d.named_parameters():
            rel, cos = T._rel(p.grad, rg[k])
            print(f"   {k:28s} rel {rel:.3e} cos {cos:.5f} |ref| {rg[k].norm().item():.3e} |g| {p.grad.norm().item():.3e}")
g = torch.Generator().manual_seed(2)
x = torch.rand(4, 3, 128, 128, generator=g); r = torch.randn(4, 1, generator=g)
prof(O.synth_state_dict_d(3, 64, seed=41), x, r, 'random')
sd = O.synth_state_dict_d(3, 64, seed=47)
for k in list(sd):
    idx = k.split(".")[1]
    if k.startswith("features.") and idx in ("3", "6", "9", "12", "15", "18", "21", "24", "27"):
        if k.endswith(".weight"): sd[k] = torch.full_like(sd[k], 0.5)
        elif k.endswith(".bias"): sd[k] = torch.full_like(sd[k], 4.0)
    if k == "features.0.bias": sd[k] = torch.full_like(sd[k], 6.0)
    if k == "classifier.0.bias": sd[k] = torch.full_like(sd[k], 40.0)
prof(sd, x, r, 'one-sided')

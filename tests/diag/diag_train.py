import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import esrganplus_b200 as E
from oracle import esrgan_oracle as O
from test_gpu_train import _make, _oracle_grads
dev = torch.device('cuda:0')
for nf, nb, (n, h, w) in [(64, 1, (2, 20, 24)), (32, 1, (1, 32, 32)), (64, 2, (1, 18, 70))]:
    sd = O.synth_state_dict_g(3, 3, nf, nb, seed=5 + nf + nb)
    net = _make(sd, nf, nb, dev).eval()
    g = torch.Generator().manual_seed(h * w)
    x = torch.rand(n, 3, h, w, generator=g)
    dy = torch.randn(n, 3, 4 * h, 4 * w, generator=g)
    y = net(x.to(dev))
    (y * dy.to(dev)).sum().backward()
    ref_y, ref_g = _oracle_grads(x, sd, nb, dy)
    print(nf, nb, (n, h, w), 'fwd rel', (y.detach().cpu() - ref_y).abs().max().item() / ref_y.std().item())
    for k, p in net.named_parameters():
        gg = p.grad.cpu().double(); r = ref_g[k].double()
        rel = (gg - r).norm().item() / r.norm().item()
        cos = (gg * r).sum().item() / (gg.norm().item() * r.norm().item())
        print(f"  {k:40s} rel {rel:.3e} cos {cos:.5f} |ref| {r.norm().item():.3e}")

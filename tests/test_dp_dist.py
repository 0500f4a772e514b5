"""Host logic of data-parallel training (esrganplus_b200/autograd.py: data_parallel / allreduce_flat /
broadcast_parameters), run on CPU with the gloo backend and world_size 2.  The flat gradient buffer a backward pass
fills is averaged across ranks with ONE collective and the per-tensor gradients are views of it, so they see the
reduced values; nothing else is exchanged.  (The kernels that fill the buffer are covered by the -m gpu tests.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from esrganplus_b200 import autograd as A


class _Fake(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Parameter(torch.zeros(3, 2))
        self.b = nn.Parameter(torch.zeros(5))
        self.register_buffer("stat", torch.zeros(2))


def test_allreduce_is_a_noop_without_process_group_or_mark():
    m = _Fake()
    flat = torch.arange(11, dtype=torch.float32)
    A.allreduce_flat(m, flat)                      # not marked
    A.data_parallel(m)
    A.allreduce_flat(m, flat)                      # marked, but torch.distributed not initialised
    assert torch.equal(flat, torch.arange(11, dtype=torch.float32))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _Fake()
    with torch.no_grad():
        m.a.fill_(float(rank + 1))
        m.b.fill_(float(10 * (rank + 1)))
        m.stat.fill_(float(rank + 7))
    A.broadcast_parameters(m, src=0)
    same = bool((m.a == 1).all() and (m.b == 10).all() and (m.stat == 7).all())
    A.data_parallel(m)
    # what a native backward hands over: one flat buffer, per-tensor gradients are views of it
    flat = torch.full((12,), float(rank + 1))
    ga, gb = flat[0:6].view(3, 2), flat[8:12 + 0][:4]
    A.allreduce_flat(m, flat)
    ok = bool(torch.allclose(flat, torch.full((12,), 1.5)) and torch.allclose(ga, torch.full((3, 2), 1.5)) and
              torch.allclose(gb, torch.full((4,), 1.5)))
    # an unmarked module on the same ranks is left alone
    m2 = _Fake()
    flat2 = torch.full((4,), float(rank))
    A.allreduce_flat(m2, flat2)
    ok2 = bool((flat2 == rank).all())
    q.put((rank, same, ok, ok2))
    dist.destroy_process_group()


def test_world2_gloo_flat_gradient_average_and_broadcast():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True, True), (1, True, True, True)], res

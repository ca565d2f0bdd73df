"""Throughput of the batched (meta-)training kernels (BASELINE.json configs[3] shape: one support + one query
word of 136 symbols per realisation, L=4).  Usage: python tools/bench_train.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import meta_viterbinet_b200 as mvn
from meta_viterbinet_b200.train import pack_params

dev = torch.device('cuda', 0)
L, S, N = 4, 16, 136
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                          torch.nn.Linear(50, S)).to(dev)
theta0 = pack_params(list(net.parameters()))


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for R in (1, 148, 296, 1184, 4736):
    tr = mvn.BatchedVNetTrainer(theta0.repeat(R, 1), L)
    ys, yq = torch.randn(R, N, device=dev), torch.randn(R, N, device=dev)
    ls = torch.randint(0, S, (R, N), device=dev, dtype=torch.int32)
    lq = torch.randint(0, S, (R, N), device=dev, dtype=torch.int32)
    reps = 20 if R <= 296 else 5
    t_maml = timeit(lambda: tr.meta_step(ys, ls, yq, lq, second_order=True), reps)
    t_fo = timeit(lambda: tr.meta_step(ys, ls, yq, lq, second_order=False), reps)
    t_sgd = timeit(lambda: tr.train_step(ys, ls), reps)
    print(f'R={R:5d}: MAML {t_maml:8.3f} ms ({R / t_maml * 1e3:10.0f} steps/s)  FO-MAML {t_fo:8.3f} ms ({R / t_fo * 1e3:10.0f} steps/s)  '
          f'train {t_sgd:8.3f} ms ({R / t_sgd * 1e3:10.0f} steps/s)', flush=True)

# CPU reference shape: the same step through torch autograd on the host (one realisation), for scale
cpu = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                          torch.nn.Linear(50, S))
opt = torch.optim.Adam(cpu.parameters(), lr=1e-3)
y, lab = torch.randn(N, 1), torch.randint(0, S, (N,))
ce = torch.nn.CrossEntropyLoss()
for name, second in (('MAML', True), ('FO-MAML', False)):
    t0 = time.perf_counter()
    for _ in range(20):
        params = list(cpu.parameters())
        x = torch.nn.functional.linear(y, params[0], params[1]).sigmoid()
        x = torch.relu(torch.nn.functional.linear(x, params[2], params[3]))
        loss_s = ce(torch.nn.functional.linear(x, params[4], params[5]), lab)
        g = torch.autograd.grad(loss_s, params, create_graph=second)
        fast = [p - 0.1 * gi for p, gi in zip(params, g)]
        x = torch.nn.functional.linear(y, fast[0], fast[1]).sigmoid()
        x = torch.relu(torch.nn.functional.linear(x, fast[2], fast[3]))
        loss_q = ce(torch.nn.functional.linear(x, fast[4], fast[5]), lab)
        mg = torch.autograd.grad(loss_q, params)
        for p, gi in zip(params, mg):
            p.grad = gi
        opt.step()
    dt = (time.perf_counter() - t0) / 20
    print(f'torch-CPU autograd {name}: {dt * 1e3:.2f} ms per step ({1 / dt:.0f} steps/s, {torch.get_num_threads()} threads)')

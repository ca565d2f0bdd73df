"""Fused true-MLSE mode of the ViterbiNet kernel (survivor masks in shared memory, in-kernel traceback) next to the
reference decision rule, L = 3..6.  Usage: python tools/bench_mlse.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import meta_viterbinet_b200 as mvn

dev = torch.device('cuda', 0)
T = bench.T
bits, y = bench.synth_frames(torch, dev, 1 << 20, 10, 1)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for L in (3, 4, 5, 6):
    torch.manual_seed(L)
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                              torch.nn.Linear(50, 2 ** L))
    w = [q.detach().to(dev).contiguous() for q in net.parameters()]
    for decision in ('reference', 'mlse', 'mlse_terminated'):
        ms = timeit(lambda: mvn.ops.vnet_decode(y, w, decision=decision))
        print(f'VNET L={L} {decision:16s}: {ms:7.3f} ms  {y.numel() / ms / 1e6:6.2f} Gsym/s', flush=True)

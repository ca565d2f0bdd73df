"""Fused ViterbiNet kernel over the memory-length sweep (BASELINE.json configs[4]).  Usage: python tools/bench_lsweep.py [Ls]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import meta_viterbinet_b200 as mvn

dev = torch.device('cuda', 0)
T = 120


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


Ls = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [3, 4, 5, 6, 7, 8]
for L in Ls:
    S = 2 ** L
    torch.manual_seed(L)
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(), torch.nn.Linear(50, S))
    w = [p.detach().to(dev).contiguous() for p in net.parameters()]
    fr = (1 << 20) if L <= 6 else (1 << 18)
    y = torch.randn(fr, T, device=dev) * 1.5
    for v in (('auto',) if L == 8 else ('auto', 'fma')):
        ms = t(lambda: mvn.ops.vnet_decode(y, w, variant=v))
        flop = 2 * (100 + 5000 + 50 * S) + 2 * S
        print(f'VNET L={L} {v:5s}: {ms:8.3f} ms {fr * T / ms / 1e6:7.2f} Gsym/s {fr * T / ms / 1e6 * flop / 1e3:7.1f} TFLOP/s algorithmic', flush=True)
    a = mvn.ops.vnet_decode(y[:65536], w)
    b = mvn.ops.vnet_decode(y[:65536], w, variant='fma' if L < 8 else 'fma_smem')
    print(f'   frames differing auto vs fma: {int((a != b).any(dim=1).sum())} of 65536; watchdog {mvn.ops.tc_timeout_status()}')

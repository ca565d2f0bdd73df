"""Times the default fused kernel of an alternative build of the library (experiments with -D switches).
Usage: python tools/time_lib.py <path/to/lib.so> [memory_length]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from meta_viterbinet_b200 import _lib

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import meta_viterbinet_b200 as mvn

L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device('cuda', 0)
w = bench.make_weights(torch, dev)
if L != 4:
    g = torch.Generator(device='cpu').manual_seed(5)
    w = w[:4] + [torch.randn(2 ** L, 50, generator=g).mul(0.3).to(dev), torch.randn(2 ** L, generator=g).mul(0.1).to(dev)]
bits, y = bench.synth_frames(torch, dev, bench.FRAMES, 10, 1)
out = torch.empty_like(y)
for _ in range(3):
    mvn.ops.vnet_decode(y, w, out=out) if 'out' in mvn.ops.vnet_decode.__code__.co_varnames else mvn.ops.vnet_decode(y, w)
ts = []
for _ in range(7):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    mvn.ops.vnet_decode(y, w)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
print(f'{sys.argv[1]} L={L}: median {ts[3]:.3f} ms  {y.numel() / ts[3] / 1e6:.3f} Gsym/s  (min {ts[0]:.3f} max {ts[-1]:.3f})')

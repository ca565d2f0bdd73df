"""Runs the round-2 side kernels a few times (for ncu): VA reference rule / fused MLSE, the cost-tensor stage loop at
64 / 128 / 256 states, fused ViterbiNet at 128 states and in MLSE mode.  Usage: python tools/run_misc.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import meta_viterbinet_b200 as mvn
from meta_viterbinet_b200.channel_taps import state_priors_table

dev = torch.device('cuda', 0)
T = bench.T
bits, y = bench.synth_frames(torch, dev, 1 << 18, 10, 1)
table = torch.as_tensor(state_priors_table(np.exp(-0.2 * np.arange(4)).reshape(1, 4), 4)).to(dev)
w = [torch.as_tensor(a).to(dev) for a in bench.load_weights(np, 10)]
for _ in range(2):
    mvn.ops.va_decode(y, table)
    mvn.ops.va_decode(y, table, decision='mlse_terminated')
    mvn.ops.vnet_decode(y, w, decision='mlse_terminated')
for L in (6, 7, 8):
    cost = torch.randn(1 << 14, T, 2 ** L, device=dev)
    for _ in range(2):
        mvn.ops.acs_decode(cost)
torch.manual_seed(7)
net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(), torch.nn.Linear(50, 128))
w7 = [p.detach().to(dev).contiguous() for p in net.parameters()]
for _ in range(2):
    mvn.ops.vnet_decode(y[:1 << 16], w7)
torch.cuda.synchronize()
print('ok')

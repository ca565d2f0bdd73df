"""Runs a few batched MAML / FO-MAML / plain steps (for ncu).  Usage: python tools/run_train.py [R]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import meta_viterbinet_b200 as mvn
from meta_viterbinet_b200.train import pack_params

R = int(sys.argv[1]) if len(sys.argv) > 1 else 148
dev = torch.device('cuda', 0)
L, S, N = 4, 16, 136
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                          torch.nn.Linear(50, S)).to(dev)
tr = mvn.BatchedVNetTrainer(pack_params(list(net.parameters())).repeat(R, 1), L)
ys, yq = torch.randn(R, N, device=dev), torch.randn(R, N, device=dev)
ls = torch.randint(0, S, (R, N), device=dev, dtype=torch.int32)
lq = torch.randint(0, S, (R, N), device=dev, dtype=torch.int32)
for _ in range(3):
    tr.meta_step(ys, ls, yq, lq, second_order=True)
    tr.train_step(ys, ls)
torch.cuda.synchronize()
print('ok')

"""Runs the Reed-Solomon decode kernel a few times on clean and corrupted words (for ncu).  Usage: python tools/run_rs.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import meta_viterbinet_b200 as mvn

dev = torch.device('cuda', 0)
k_bytes, nsym, n_words = 60, 8, 1 << 19
g = torch.Generator(device='cpu').manual_seed(1)
msg = torch.randint(0, 2, (n_words, 8 * k_bytes), generator=g).float().to(dev)
cw = mvn.ops.rs_encode(msg, nsym)
bad = cw.clone()
bad[:, 8 * 5 + 2] = 1 - bad[:, 8 * 5 + 2]
bad[:, 8 * 40 + 7] = 1 - bad[:, 8 * 40 + 7]
for _ in range(3):
    a = mvn.ops.rs_decode(cw, nsym)
    b = mvn.ops.rs_decode(bad, nsym)
torch.cuda.synchronize()
print('ok', bool((a == msg).all()), bool((b == msg).all()))

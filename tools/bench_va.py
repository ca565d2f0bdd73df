"""Times the classical VA kernel (both decision rules, output formats).  Usage: python tools/bench_va.py"""
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


def t(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


bits, y = bench.synth_frames(torch, dev, bench.FRAMES, 10, 1)
for L in (4, 3, 5, 6, 7, 8):
    table = torch.as_tensor(state_priors_table(np.exp(-0.2 * np.arange(L)).reshape(1, L), L)).to(dev)
    fr = bench.FRAMES if L <= 6 else bench.FRAMES // 4
    for d in ('reference', 'mlse_terminated'):
        for fmt in (mvn.OUT_F32, mvn.OUT_BITS):
            ms = t(lambda: mvn.ops.va_decode(y[:fr], table, decision=d, out_format=fmt))
            print(f'VA L={L} {d:16s} {"f32 " if fmt == mvn.OUT_F32 else "bits"}: {ms:7.3f} ms {fr * T / ms / 1e6:8.1f} Gsym/s '
                  f'{8 * fr * T / ms / 1e6 / 6553:.3f} of measured HBM (8 B/sym)', flush=True)

"""Times the non-headline kernels on a B200: classical VA (configs[1]) and the ACS-only stage loop on a cost
tensor, plus the L sweep of the fused kernel (configs[4]).  Usage: python tools/bench_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import meta_viterbinet_b200 as mvn
from meta_viterbinet_b200.channel_taps import channel_taps, state_priors_table

dev = torch.device('cuda', 0)


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


T = bench.T
frames = bench.FRAMES
bits, y = bench.synth_frames(torch, dev, frames, 10, 1)
for L in (4,):
    table = torch.as_tensor(state_priors_table(channel_taps(L, 0.2, 'time_decay'), L)).to(dev)
    for n_h in (1, 256):
        tab = table.repeat(n_h, 1).contiguous()
        ms = timeit(lambda: mvn.ops.va_decode(y, tab))
        print(f'VA   L={L} n_h={n_h:4d} frames={frames} T={T}: {ms:8.3f} ms  {frames * T / ms / 1e6:8.2f} Gsym/s  '
              f'{8 * frames * T / ms / 1e6:8.1f} GB/s (8 B/sym)', flush=True)
    cnt = mvn.ops.new_counters()
    ms = timeit(lambda: mvn.ops.va_decode(y, table, target=bits, counters=cnt, want_decoded=False))
    print(f'VA   L={L} fused BER, no decoded write: {ms:8.3f} ms  {frames * T / ms / 1e6:8.2f} Gsym/s', flush=True)
    ms = timeit(lambda: mvn.ops.va_decode(y, table, out_format=mvn.OUT_BITS))
    print(f'VA   L={L} bit-packed output: {ms:8.3f} ms  {frames * T / ms / 1e6:8.2f} Gsym/s', flush=True)

for L in (3, 4, 5, 6, 8):
    S = 2 ** L
    fr = (1 << 18) if L <= 5 else (1 << 16)
    cost = torch.randn(fr, T, S, device=dev)
    ms = timeit(lambda: mvn.ops.acs_decode(cost))
    byt = fr * T * (S * 4 + 4)
    print(f'ACS  L={L} frames={fr}: {ms:8.3f} ms  {fr * T / ms / 1e6:8.2f} Gsym/s  {byt / ms / 1e6:8.1f} GB/s', flush=True)
    del cost

for L in range(3, 9):
    S = 2 ** L
    torch.manual_seed(L)
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                              torch.nn.Linear(50, S))
    w = [p.detach().to(dev).contiguous() for p in net.parameters()]
    fr = frames if L <= 5 else frames // 4
    ms = timeit(lambda: mvn.ops.vnet_decode(y[:fr], w), reps=3)
    flop = 2 * (100 + 5000 + 50 * S) + 2 * S
    print(f'VNET L={L} frames={fr}: {ms:8.3f} ms  {fr * T / ms / 1e6:8.3f} Gsym/s  {flop * fr * T / ms / 1e9:7.2f} TFLOP/s', flush=True)

# Reed-Solomon (f2): words/s and HBM bytes (fp32 bit rows in and out), clean words and words with nsym/2 byte errors
import numpy as np
for k_bytes, nsym, n_words in ((15, 2, 1 << 20), (60, 8, 1 << 19), (223, 32, 1 << 17)):
    n = k_bytes + nsym
    g = torch.Generator(device='cpu').manual_seed(k_bytes)
    msg = torch.randint(0, 2, (n_words, 8 * k_bytes), generator=g).float().to(dev)
    ms = timeit(lambda: mvn.ops.rs_encode(msg, nsym), reps=3)
    cw = mvn.ops.rs_encode(msg, nsym)
    print(f'RS   encode k={k_bytes} nsym={nsym} words={n_words}: {ms:8.3f} ms  {n_words / ms / 1e3:8.2f} Mwords/s  '
          f'{4 * 8 * (k_bytes + n) * n_words / ms / 1e6:8.1f} GB/s', flush=True)
    for label, n_bad in (('clean', 0), (f'{nsym // 2} byte errors', nsym // 2)):
        rx = cw.clone()
        if n_bad:
            pos = torch.stack([torch.randperm(n, generator=g)[:n_bad] for _ in range(4096)]).to(dev)   # pattern pool
            pos = pos[torch.arange(n_words, device=dev) % 4096]
            rows = torch.arange(n_words, device=dev).unsqueeze(1).expand_as(pos)
            rx[rows, 8 * pos + 3] = 1 - rx[rows, 8 * pos + 3]
        ms = timeit(lambda: mvn.ops.rs_decode(rx, nsym), reps=3)
        ok = bool((mvn.ops.rs_decode(rx, nsym) == msg).all())
        print(f'RS   decode k={k_bytes} nsym={nsym} {label:>15s}: {ms:8.3f} ms  {n_words / ms / 1e3:8.2f} Mwords/s  '
              f'{4 * 8 * (k_bytes + n) * n_words / ms / 1e6:8.1f} GB/s  recovered={ok}', flush=True)
    del msg, cw, rx

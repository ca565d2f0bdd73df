"""End-to-end (pinned host buffers in/out) throughput of mvn_ctx_vnet_decode_host versus chunk size.
Usage: python tools/tune_e2e.py"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from meta_viterbinet_b200 import _lib

dev = torch.device('cuda', 0)
lib = _lib.load()
w = bench.make_weights(torch, dev)
_, y = bench.synth_frames(torch, dev, bench.FRAMES, 10, 1)
T = y.shape[1]
y_host = y.cpu().pin_memory()
out_host = torch.empty_like(y_host).pin_memory()
w_host = [t.cpu().contiguous() for t in w]
wave = 148 * 128
for waves in (0, 1, 2, 3, 4, 6, 8, 16):
    ctx = ctypes.c_void_p()
    _lib.check(lib.mvn_ctx_create(ctypes.byref(ctx), 0, waves * wave, T, bench.MEMORY_LENGTH))
    _lib.check(lib.mvn_ctx_set_vnet_weights_host(ctx, *[ctypes.c_void_p(t.data_ptr()) for t in w_host]))
    step = lambda: _lib.check(lib.mvn_ctx_vnet_decode_host(ctx, ctypes.c_void_p(y_host.data_ptr()), bench.FRAMES, T, T, 0,
                                                           ctypes.c_void_p(out_host.data_ptr())))
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f'chunk = {waves if waves else "auto"} waves: {dt * 1e3:7.2f} ms per 2^20 frames  {y.numel() / dt / 1e9:6.2f} Gsym/s  '
          f'({2 * y.numel() * 4 / dt / 1e9:5.1f} GB/s over PCIe, both directions)', flush=True)
    lib.mvn_ctx_destroy(ctx)

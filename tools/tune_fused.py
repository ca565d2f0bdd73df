"""Times every fused-kernel variant (ops.set_fused_variant) on the bench workload and checks that all of
them produce identical bits.  Usage (GPU box): python tools/tune_fused.py [frames]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import meta_viterbinet_b200 as mvn
from meta_viterbinet_b200 import _lib

frames = int(sys.argv[1]) if len(sys.argv) > 1 else bench.FRAMES
dev = torch.device('cuda', 0)
w = bench.make_weights(torch, dev)
bits, y = bench.synth_frames(torch, dev, frames, 10, 1)
lib = _lib.load()
names = {4: 'fma const M2 384 u10', 1: 'fma smem M2 256 u10', 2: 'fma const M2 320 u10', 3: 'tcgen05 warp-specialised fp16x2 (default)'}
ref = None
for v in (4, 1, 3):
    mvn.ops.set_fused_variant({x: k for k, x in mvn.ops.FUSED_VARIANTS.items()}[v])
    out = mvn.ops.vnet_decode(y, w)
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    same = bool(torch.equal(out, ref))
    nbad = int((out != ref).any(dim=1).sum())
    sub = y[:4096].contiguous()
    _, pri = mvn.ops.vnet_decode(sub, w, return_priors=True)
    if v == 4:
        pri_ref = pri
    rel = float(((pri - pri_ref).abs() / pri_ref.abs().amax(dim=-1, keepdim=True)).max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        mvn.ops.vnet_decode(y, w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rate = frames * bench.T / ms / 1e6
    print(f'variant {v} [{names[v]:26s}] {ms:8.3f} ms  {rate:7.3f} Gsym/s  {rate * bench.FLOP_PER_SYMBOL / 1e3:6.2f} TFLOP/s  '
          f'bits_equal_to_fma={same} frames_differing={nbad} priors_rel_vs_fma={rel:.2e}', flush=True)
mvn.ops.set_fused_variant({x: k for k, x in mvn.ops.FUSED_VARIANTS.items()}[0])
lib.mvn_tc_timeout_status.restype = ctypes.c_int
print('tc timeout flag:', lib.mvn_tc_timeout_status())

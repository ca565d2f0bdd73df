"""Runs one fused-kernel variant a few times (for ncu).  Usage: python tools/run_variant.py <variant> [frames]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import meta_viterbinet_b200 as mvn
from meta_viterbinet_b200 import _lib

v = int(sys.argv[1])
frames = int(sys.argv[2]) if len(sys.argv) > 2 else bench.FRAMES
dev = torch.device('cuda', 0)
w = bench.make_weights(torch, dev)
bits, y = bench.synth_frames(torch, dev, frames, 10, 1)
lib = _lib.load()
mvn.ops.set_fused_variant({x: k for k, x in mvn.ops.FUSED_VARIANTS.items()}[v])
for _ in range(3):
    out = mvn.ops.vnet_decode(y, w)
torch.cuda.synchronize()
print('ok', float(out.mean()))

"""Pipeline timeline of the tcgen05 fused kernel (debug library built with -DMVN_TC_TRACE).
Usage: python tools/tc_trace.py [memory_length]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from meta_viterbinet_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libmvn_trace.so')   # nvcc <build.py flags> -DMVN_TC_TRACE
import meta_viterbinet_b200 as mvn

dev = torch.device('cuda', 0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4   # memory length (2^L states)
torch.manual_seed(0)
_net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                           torch.nn.Linear(50, 2 ** L))
w = [q.detach().to(dev).contiguous() for q in _net.parameters()]
bits, y = bench.synth_frames(torch, dev, 148 * 128 * 4, 10, 1)
for _ in range(2):
    mvn.ops.vnet_decode(y, w)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (64 * 32))()
lib.mvn_debug_tc_trace(buf)
t = np.array(buf[:]).reshape(64, 32)
names = ['P start', 'P computed', 'P slot acquired', 'P stored', 'M a_full seen', 'M issued', 'C wait d_full', 'C d_full seen',
         'C h2->tmem done', 'C barrier passed', 'C mma2 issued', 'C d2_full seen', 'C slot released', 'C acs done']
base = t[8, 0]
print('stage ' + ' '.join(f'{n[:12]:>13s}' for n in names))
for s in range(8, 24):
    print(f'{s:5d} ' + ' '.join(f'{int(t[s, e] - base):13d}' for e in range(14)))
d = lambda a, b: float(np.median(t[16:56, b] - t[16:56, a]))
print('\nmedian segment lengths (cycles), stages 16..55:')
for a, b, label in [(0, 1, 'producer compute'), (1, 2, 'producer waits for slot'), (2, 3, 'producer tcgen05.st (+ wait::st unless late publish)'),
                    (3, 4, 'a_full arrive -> MMA warp wakes'), (4, 15, 'MMA warp waits for slot_free (consumers)'),
                    (15, 5, 'MMA warp issues the layer-2 MMAs'), (5, 7, 'commit -> converter sees d_full'),
                    (7, 8, 'converter: ld D, relu, split, st h2'), (8, 9, 'converter barrier'), (9, 10, 'issue the layer-3 MMAs'),
                    (10, 11, 'layer-3 issued -> consumer sees d2_full'), (11, 12, 'consumer: ld priors, release slot'), (12, 13, 'ACS'),
                    (6, 7, 'converter idle waiting for d_full')]:
    print(f'  {label:40s} {d(a, b):8.0f}')
late = t[20, 30] == 0   # MVN_TC_LATE_PUBLISH: wait::st / fence / arrive happen inside the next stage
print('  producer warp 0 tail: computed -> try_wait done -> fence -> st issued' + (' -> next start (late publish)' if late else ' -> wait::st done -> fence -> arrived -> next start'))
print('    ' + ' '.join(f'{float(np.median(t[16:56, b] - t[16:56, a])):6.0f}' for a, b in ([(1, 28), (28, 2), (2, 29)] if late else [(1, 28), (28, 2), (2, 29), (29, 30), (30, 3), (3, 31)]))
      + f' {float(np.median(t[17:57, 0] - t[16:56, 29 if late else 31])):6.0f}')
print(f'  d_full seen (slot use k) -> d_full seen (use k+1), same slot: {float(np.median(t[18:56, 7] - t[16:54, 7])):8.0f}')
print(f'  stage period (consumer)                  {float(np.median(np.diff(t[16:56, 13]))):8.0f}')
mhz = (t[60, 0] - t[4, 0]) / max(1, (t[60, 14] - t[4, 14])) * 1000.0
print(f'  effective SM clock during the kernel (clock64 / globaltimer over stages 4..60): {mhz:8.0f} MHz')
print('  A operand stored, per producer warp relative to warp 0 (median over stages 16..55, cycles); warps w, w+4, w+8 share a quadrant:')
print('    ' + ' '.join(f'{float(np.median(t[16:56, 16 + w] - t[16:56, 16])):6.0f}' for w in range(12 if t[20, 31] == 0 else 16)))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, meta_viterbinet_b200 as mvn
dev = torch.device('cuda', 0); T = 120
def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
bits, y = bench.synth_frames(torch, dev, 1 << 20, 10, 1)
for name, w in (('trained', [torch.as_tensor(a).to(dev) for a in bench.load_weights(np, 10)]), ('random-init', bench.make_weights(torch, dev))):
    ms = t(lambda: mvn.ops.vnet_decode(y, w))
    print(f'tcgen05 L=4 {name}: {ms:.3f} ms {(1 << 20) * T / ms / 1e6:.2f} Gsym/s  timeout {mvn.ops.tc_timeout_status()}', flush=True)
yb = y.clone(); yb[::7, 5] = 40.0; yb[3, 9] = float('nan'); yb[11, 2] = -1e6   # out-of-range samples take the clamped path
a = mvn.ops.vnet_decode(yb[:8192], w, variant='tcgen05'); b = mvn.ops.vnet_decode(yb[:8192], w, variant='fma')
print('out-of-range rows: frames differing tc vs fma', int((a != b).any(dim=1).sum()))

"""A/B of the two thread layouts of the cost-tensor stage loop (mvn_acs_decode_ex) for 8..256 states.
Usage: python tools/bench_acs.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import meta_viterbinet_b200 as mvn
dev = torch.device('cuda', 0); T = 120
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for L in (3, 4, 5, 6, 7, 8):
    S = 2 ** L
    fr = (1 << 18) if L <= 5 else (1 << 16)
    cost = torch.randn(fr, T, S, device=dev)
    byt = fr * T * (S * 4 + 4)
    for layout in ('lane_per_frame', 'states_on_lanes'):
        ms = t(lambda: mvn.ops.acs_decode(cost, layout=layout))
        print(f'ACS  L={L} {layout:16s} frames={fr}: {ms:8.3f} ms  {fr * T / ms / 1e6:8.2f} Gsym/s  {byt / ms / 1e6:8.1f} GB/s  '
              f'{byt / ms / 1e6 / 6553:.3f} of measured HBM', flush=True)
    del cost

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import meta_viterbinet_b200 as mvn
dev = torch.device('cuda', 0)
T = 120
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for L in (6, 7):
    S = 2 ** L
    torch.manual_seed(L)
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(), torch.nn.Linear(50, S))
    w = [p.detach().to(dev).contiguous() for p in net.parameters()]
    fr = 1 << 18
    y = torch.randn(fr, T, device=dev) * 1.5
    for v in ('auto', 'fma'):
        ms = t(lambda: mvn.ops.vnet_decode(y, w, variant=v))
        print(f'L={L} {v}: {ms:.3f} ms {fr * T / ms / 1e6:.2f} Gsym/s', flush=True)
    a = mvn.ops.vnet_decode(y[:65536], w, variant='auto'); b = mvn.ops.vnet_decode(y[:65536], w, variant='fma')
    print('frames differing tc vs fma', int((a != b).any(dim=1).sum()), 'of 65536; timeout', mvn.ops.tc_timeout_status())

import sys, torch, numpy as np, ctypes
sys.path.insert(0, '/root/repo')
import bench
import meta_viterbinet_b200 as mvn
dev = torch.device('cuda', 0)
for L in (4, 3, 5, 6, 1):
    torch.manual_seed(L)
    S = 2 ** L
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(), torch.nn.Linear(50, S))
    w = [p.detach().to(dev).contiguous() for p in net.parameters()]
    frames = 1 << 19
    _, y = bench.synth_frames(torch, dev, frames, 9, L)
    ref = mvn.ops.vnet_decode(y, w)
    mvn.ops.set_fused_variant('fma')
    fma = mvn.ops.vnet_decode(y[:65536], w)
    mvn.ops.set_fused_variant('auto')
    diff_frames = int((fma != ref[:65536]).any(dim=1).sum())
    bad = 0
    for it in range(12):
        out = mvn.ops.vnet_decode(y, w)
        bad += int((out != ref).sum())
    f = mvn._lib.load().mvn_tc_timeout_status; f.restype = ctypes.c_int
    print(f'L={L}: 12 repeats of {frames} frames x {y.shape[1]}: differing symbols vs first run {bad}; frames differing from the FMA kernel {diff_frames} of 65536; timeout flag {f()}', flush=True)

"""PCIe ceiling of the box: pinned host <-> device copies, one direction and both at once (for the e2e roofline)."""
import time
import torch

n = 1 << 28   # 1 GiB of fp32 = 2^28 floats
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device='cuda')
d_out = torch.empty(n, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


run(True, True, 1)
gb = n * 4 / 1e9
print(f'H2D alone : {gb / run(True, False):6.1f} GB/s')
print(f'D2H alone : {gb / run(False, True):6.1f} GB/s')
t = run(True, True)
print(f'both      : {gb / t:6.1f} GB/s in each direction at once ({2 * gb / t:6.1f} GB/s total)')

#!/usr/bin/env python
"""bench.py — decoded symbols/s of the fused ViterbiNet detection path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
  ViterbiNet, memory_length 4 (16 states), 2^20 synthetic ISI-AWGN frames x 120 symbols PER GPU, the 7..12 dB SNR sweep
  of the reference's plotter (plotter_main.py:117-122) with one reference-trained checkpoint per SNR point
  (tests/golden/ckpt_vnet_L4.npz, produced by the reference's own VNETTrainer.train(), trainer.py:455-490).
A "step" = one pass of the product's sweep API, `sweep.run_sweep`, over the whole job: the N x 2^20 frames are split over
the six SNR points, the (point, row block) work items are sharded over the ranks (no data-path collective), every item is
one launch of the fused priors-MLP (layers 2-3 on tcgen05 tensor cores) + ACS + decision kernel with the decoded words
written as fp32 [B,T] (the reference's dtype) and BER/FER counted in-kernel, and the [6,4] error counters are
all-reduced over NCCL once per step.

One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same metric through the C-ABI
host-buffer entry point (pinned host y in, decoded words out, copies inside the timed region) with the other host-side
forms beside it.  `--impl reference` times the torch-CPU port of the reference's VNETDetector forward
(oracle/torch_port.py) on the host cores, rank 0 only.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MEMORY_LENGTH = 4
N_STATES = 16
T = 120
FRAMES = 1 << 20
SNR_SWEEP = [7, 8, 9, 10, 11, 12]          # plotter_main.py:117-122
GAMMA = 0.2
FLOP_PER_SYMBOL = 2 * (100 + 5000 + 50 * N_STATES) + 2 * N_STATES     # 11 832
HBM_BYTES_PER_SYMBOL = 8                                              # fp32 y in, fp32 decoded out
METRIC = 'decoded symbols/sec, ViterbiNet L=4 (16-state)'
UNIT = 'symbols/s'
CKPT = os.path.join(ROOT, 'tests', 'golden', 'ckpt_vnet_L4.npz')
MAML_FLOP_PER_STEP = 12 * 2 * 5900 * 136   # fwd = 2*5900*N, FO part 3 fwd(s) + 3 fwd(q), HVP + 2*(fwd+bwd)(s), N = 136
TRAIN_FLOP_PER_STEP = 3 * 2 * 5900 * 136   # plain step: fwd + bwd


def env_int(name, default):
    return int(os.environ.get(name, default))


def load_weights(np, snr):
    """[W1,b1,W2,b2,W3,b3] trained by the reference's VNETTrainer.train() at this SNR (numpy fp32)."""
    g = np.load(CKPT)
    return [np.ascontiguousarray(g[f'snr{snr}_w{i}'], dtype=np.float32) for i in range(6)]


def synth_frames(torch, device, frames, snr_db, seed):
    """bits -> pad L zeros -> BPSK -> ISI (time_decay taps, gamma 0.2) -> AWGN; Appendix A of SURVEY.md."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    L = MEMORY_LENGTH
    bits = torch.randint(0, 2, (frames, T), generator=g, device=device, dtype=torch.int8)
    s = torch.ones((frames, T + L), device=device)
    s[:, :T] = 1.0 - 2.0 * bits.float()
    h = torch.exp(-GAMMA * torch.arange(L, device=device, dtype=torch.float32))
    y = torch.zeros((frames, T), device=device)
    for i in range(L):
        y += h[L - 1 - i] * s[:, i:i + T]
    y += (10 ** (-snr_db / 20.0)) * torch.randn((frames, T), generator=g, device=device)
    return bits.float().contiguous(), y.contiguous()


def make_weights(torch, device):
    """Random-init weights of the reference architecture (tools/ use them for kernel tuning; the bench itself decodes with
    the reference-trained checkpoints)."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50),
                              torch.nn.ReLU(), torch.nn.Linear(50, N_STATES))
    return [p.detach().to(device).contiguous() for p in net.parameters()]


class ClockSampler:
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix='mvn_clocks_', suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=open(self.path, 'w'),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(',')]
                if len(p) < 8:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for n, v in zip(names, p[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_port_rate(frames_per_chunk, chunks, weights_np, y_np, reps=1):
    """symbols/s of the torch-CPU port of VNETDetector.forward(y,'val') on all host threads."""
    import torch
    from oracle import torch_port as tp
    torch.set_num_threads(os.cpu_count() or 1)
    net = tp.make_net(N_STATES)
    tp.load_weights(net, weights_np)
    y = torch.as_tensor(y_np)
    best = None
    with torch.no_grad():
        tp.vnet_forward_val(net, y[:min(1024, y.shape[0])], T)            # warm-up
        for _ in range(reps):
            t0 = time.perf_counter()
            for c in range(chunks):
                tp.vnet_forward_val(net, y[c * frames_per_chunk:(c + 1) * frames_per_chunk], T)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return frames_per_chunk * chunks * T / best, best


def run_reference(args, rank):
    """Reference arm: the reference's CPU implementation (torch port, all host threads), rank 0 only."""
    if rank != 0:
        return 0
    import numpy as np
    import torch
    w = load_weights(np, 10)
    chunk, chunks = 16384, 8                        # bounded sample per step: 131 072 frames x 120 (~1.8 s)
    _, y = synth_frames(torch, 'cpu', chunk * chunks, 10, 3450002)
    y_np = y.numpy()
    for _ in range(max(args.warmup, 1)):
        cpu_port_rate(chunk, 1, w, y_np)
    dt = 0.0                                        # only the forward passes are timed (not net construction / warm-up)
    for _ in range(args.steps):
        dt += cpu_port_rate(chunk, chunks, w, y_np)[1]
    value = args.steps * chunk * chunks * T / dt
    cores = torch.get_num_threads()
    sample = (f'{chunk * chunks} frames x {T} symbols per step (chunks of {chunk}) at 10 dB with the 10 dB checkpoint, '
              f'torch {torch.__version__} CPU')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus):
    return {'workload': f'fused ViterbiNet priors-MLP+ACS+decision, memory_length 4 (16 states), {FRAMES} frames x '
                        f'{T} symbols per GPU, synthetic ISI-AWGN (time_decay taps, gamma {GAMMA}), SNR sweep '
                        f'{SNR_SWEEP[0]}..{SNR_SWEEP[-1]} dB through sweep.run_sweep (work items sharded over the ranks), '
                        f'one reference-trained checkpoint per SNR point',
            'frames_per_gpu': FRAMES, 'block_length': T, 'n_states': N_STATES, 'out_dtype': 'f32 [B,T]',
            'l2': 'inputs (503 MB y + 503 MB targets per GPU and step) are larger than the 126 MB L2; no flush needed',
            'parallelism': f'{len(SNR_SWEEP)} SNR points x row blocks sharded over {n_gpus} GPU(s), [6,4] counters all-reduced'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--frames', type=int, default=FRAMES, help=argparse.SUPPRESS)
    ap.add_argument('--no-cpu-baseline', action='store_true', help=argparse.SUPPRESS)
    ap.add_argument('--no-extras', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank, world, local = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)
    if args.impl == 'reference':
        return run_reference(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    import meta_viterbinet_b200 as mvn
    from meta_viterbinet_b200 import _lib, sweep

    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm'
    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    lib = _lib.load()
    stream = _lib.stream()

    # ---- the job: world x 2^20 frames split over the six SNR points; this rank's work items, resident in HBM
    frames_per_point = args.frames * world // len(SNR_SWEEP)
    items = sweep.work_items(len(SNR_SWEEP), world)
    mine = [items[k] for k in sweep.partition(len(items), world, rank)]
    w_np = {snr: load_weights(np, snr) for snr in SNR_SWEEP}
    w_dev = {snr: [torch.as_tensor(a).to(device) for a in w_np[snr]] for snr in SNR_SWEEP}
    data = {}
    for i, b, blocks in mine:
        rows = sweep.partition(frames_per_point, blocks, b)
        bits, y = synth_frames(torch, device, len(rows), SNR_SWEEP[i], 3450002 + 100 * i + b)   # noise_seed of config.yaml:40
        data[(i, rows.start)] = (y, bits, torch.empty_like(y))
    my_frames = sum(v[0].shape[0] for v in data.values())

    # The items of a rank are independent launches: they go round-robin to three streams, so that the last (partial) wave
    # of one launch overlaps the first waves of the next instead of leaving SMs idle (6 launches of 9.2 waves each at N=1).
    side = [torch.cuda.Stream(device) for _ in range(3)]
    turn = [0]

    def evaluate_block(snr, first, n, counters_row):
        y, bits, dec = data[(SNR_SWEEP.index(snr), first)]
        s = side[turn[0] % len(side)]
        turn[0] += 1
        s.wait_stream(torch.cuda.current_stream())      # run_sweep zeroes the counters on the current stream
        _lib.check(lib.mvn_vnet_decode(_lib.ptr(y), n, T, MEMORY_LENGTH, T, *[_lib.ptr(w) for w in w_dev[snr]], 0,
                                       _lib.ptr(dec), None, _lib.ptr(bits), T, 0, _lib.ptr(counters_row),
                                       ctypes.c_void_p(s.cuda_stream)))

    def join():
        for s in side:
            torch.cuda.current_stream().wait_stream(s)

    def step():
        return sweep.run_sweep(SNR_SWEEP, frames_per_point, evaluate_block, device=device, rank=rank, world_size=world,
                               before_reduce=join)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # measured FP32 peak (register-only FMA micro-benchmark) before the clock sampler starts
    peak_ffma, _ = _lib.fp32_peak(0, 2048)
    peak_ffma2, _ = _lib.fp32_peak(1, 2048)
    peak_fma = max(peak_ffma, peak_ffma2)

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        total = step()                                  # launches + the only collective: [6,4] int64 per step
    e1.record()
    barrier()
    launches = _lib.launch_count()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # kernel-only time of this rank's launches (the same three streams, no collective in between), for the roofline
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.zeros((len(SNR_SWEEP), 4), dtype=torch.int64, device=device)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        for (i, first), (y, _, _) in data.items():
            evaluate_block(SNR_SWEEP[i], first, y.shape[0], scratch[i])
        join()
    k1.record()
    torch.cuda.synchronize()
    kern_ms_per_step = k0.elapsed_time(k1) / args.steps
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    fr = torch.tensor([my_frames], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr)
    ms_total = float(t.item())
    job_frames = int(fr.item())
    symbols_per_step = job_frames * T
    value = symbols_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e through the host-buffer C-ABI calls: pinned buffers from mvn_host_alloc, every item of this rank per step
    def host_alloc(nbytes):
        p = ctypes.c_void_p()
        _lib.check(lib.mvn_host_alloc(ctypes.byref(p), max(nbytes, 16), 0))
        return p
    host = {}
    for key, (y, bits, dec) in data.items():
        nb = y.numel() * 4
        hy, hb, ho = host_alloc(nb), host_alloc(nb), host_alloc(nb)
        yc, bc = y.cpu(), bits.cpu()
        ctypes.memmove(hy, ctypes.c_void_p(yc.data_ptr()), nb)
        ctypes.memmove(hb, ctypes.c_void_p(bc.data_ptr()), nb)
        host[key] = (hy, hb, ho, y.shape[0])
    ctx = ctypes.c_void_p()
    _lib.check(lib.mvn_ctx_create(ctypes.byref(ctx), local, 0, T, MEMORY_LENGTH))
    cnt_host = (ctypes.c_uint64 * 4)()
    taps = (ctypes.c_double * MEMORY_LENGTH)(*[float(np.exp(-GAMMA * i)) for i in range(MEMORY_LENGTH)])

    def set_w(snr):
        _lib.check(lib.mvn_ctx_set_vnet_weights_host(ctx, *[a.ctypes.data_as(ctypes.c_void_p) for a in w_np[snr]]))

    def e2e_words(fmt):
        def fn():                       # the items stream through ONE pipeline (weight ring), one drain per step
            for (i, first), (hy, hb, ho, n) in host.items():
                set_w(SNR_SWEEP[i])
                _lib.check(lib.mvn_ctx_vnet_decode_host_async(ctx, hy, n, T, T, fmt, ho))
            _lib.check(lib.mvn_ctx_synchronize(ctx))
        return fn

    def e2e_counters():
        for (i, first), (hy, hb, ho, n) in host.items():
            set_w(SNR_SWEEP[i])
            _lib.check(lib.mvn_ctx_vnet_eval_host(ctx, hy, hb, n, T, T, T, 0, 0, None, cnt_host))

    def e2e_device_source():
        for (i, first), (hy, hb, ho, n) in host.items():
            set_w(SNR_SWEEP[i])
            _lib.check(lib.mvn_ctx_vnet_sweep_point(ctx, n, T, T, taps, 1, float(SNR_SWEEP[i]), 17 + first, 0, cnt_host))

    def timed(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return symbols_per_step * args.steps / float(dt.item())
    e2e_value = timed(e2e_words(0))
    # the words that came back through the host pipeline == the device-resident decode, EVERY row of every item
    e2e_ok = True
    for key, (hy, hb, ho, n) in host.items():
        back = torch.frombuffer((ctypes.c_char * (n * T * 4)).from_address(ho.value), dtype=torch.float32).reshape(n, T)
        e2e_ok = e2e_ok and bool(torch.equal(back, data[key][2].cpu()))
    e2e_bits_value = timed(e2e_words(1))
    e2e_cnt_value = timed(e2e_counters)
    e2e_src_value = timed(e2e_device_source)
    lib.mvn_ctx_destroy(ctx)
    for hy, hb, ho, n in host.values():
        for p in (hy, hb, ho):
            lib.mvn_host_free(p)
    ok_t = torch.tensor([1 if e2e_ok else 0], device=device)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    e2e_ok = bool(ok_t.item())

    extras = None
    if not args.no_extras:
        import bench_extras
        extras = bench_extras.run(torch, mvn, _lib, device, rank, world, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- sanity outside the timed region: a 4 096-frame subsample of this rank's first item against the oracle
    from oracle import viterbinet_oracle as orc
    (i0, first0), (y0, bits0, dec0) = next(iter(data.items()))
    snr0 = SNR_SWEEP[i0]
    sub = slice(0, 4096)
    dec_k, pri_k = mvn.ops.vnet_decode(y0[sub], w_dev[snr0], return_priors=True)
    own, _ = orc.vnet_decode_from_priors(pri_k.cpu().numpy())
    exact = orc.vnet_priors(y0[sub].cpu().numpy(), w_np[snr0], dtype=np.float64)
    pri_err = float(np.max(np.abs(pri_k.cpu().numpy() - exact) / np.max(np.abs(exact), axis=-1, keepdims=True)))
    sanity = {'frames': 4096, 'snr_db': snr0,
              'bits_equal_stage_loop_on_own_priors': bool(np.array_equal(dec_k.cpu().numpy(), own)),
              'same_bits_in_timed_output': bool(torch.equal(dec_k, dec0[sub])),
              'priors_max_err_rel_rowmax_vs_fp64': pri_err,
              'frames_differing_from_fp32_oracle_forward': int((orc.vnet_decode(y0[sub].cpu().numpy(), w_np[snr0]) != own).any(axis=1).sum())}
    assert sanity['bits_equal_stage_loop_on_own_priors'] and sanity['same_bits_in_timed_output'] and pri_err < 1e-5, sanity
    ber_by_snr = {}
    for i, snr in enumerate(SNR_SWEEP):
        be, fe, nb, nf = [int(v) for v in total[i].tolist()]
        ber_by_snr[str(snr)] = {'bit_errors': be, 'frame_errors': fe, 'bits': nb, 'frames': nf, 'ber': be / max(nb, 1)}
        assert nf == frames_per_point and 1e-3 < be / nb < 0.1, (snr, be, nb, nf)      # trained weights: BER ~ 1e-2

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    n_launch = len(data)
    k_ms = kern_ms_per_step / n_launch                       # average duration of one launch
    sym_per_launch = my_frames * T / n_launch
    achieved_tflops = FLOP_PER_SYMBOL * sym_per_launch / (k_ms * 1e-3) / 1e12
    peak_tflops = 2 * peak_fma / 1e12
    traffic = None
    try:   # DRAM bytes per launch from the ncu --set full capture (profiles/traffic.json holds bytes per frame)
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get('vnet_decode_bytes_per_frame') * my_frames / n_launch
    except Exception:
        pass
    # The default kernel runs layers 2 and 3 (98 % of the flops) on the tensor cores: fp16 tcgen05 MMAs on a two-piece
    # scaled split, i.e. 3 MMA chains on a 128 x 64 x 112 padded tile (layer 2) and 3 on a 128 x 16 x 64 tile (layer 3)
    # per 128 symbols -> 3*2*(64*112 + 16*64) = 49 152 executed tensor flop per symbol for 11 600 algorithmic ones;
    # sigmoid, split, ReLU, ACS and the decision are on the CUDA cores.  `achieved` is ALGORITHMIC flop (11 832 / symbol)
    # over the CUDA-event time of the launches; the tensor peak is MEASURED_PEAKS.json's bf16 burst figure.
    bf16_peak = peaks.get('bf16_tflops', 1590.0)
    tensor_executed = 3 * 2 * (64 * 112 + 16 * 64) * sym_per_launch / (k_ms * 1e-3) / 1e12
    roofline = {'bound': 'tensor', 'achieved': achieved_tflops, 'peak': bf16_peak, 'unit': 'TFLOP/s',
                'frac': achieved_tflops / bf16_peak, 'traffic': traffic,
                'kernel': 'vnet_decode_tc_kernel<4> (tcgen05 fp16x2-split layers 2+3, CUDA-core sigmoid/ACS)',
                'kernel_ms': k_ms, 'kernel_ms_note': 'kernel-only time of a step / launches per step (the launches of a step overlap their tails on three streams)', 'launches_per_step': n_launch, 'symbols_per_launch': sym_per_launch,
                'flop_per_symbol': FLOP_PER_SYMBOL,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst), of measured' if peaks else 'fallback 1590, of fallback',
                'tensor_executed': {'tflops': tensor_executed, 'frac': tensor_executed / bf16_peak,
                                    'note': 'executed fp16 MMA flop incl. the 3 chains of the two-piece split and tile padding'},
                'fp32': {'achieved': achieved_tflops, 'peak': peak_tflops, 'frac': achieved_tflops / peak_tflops,
                         'peak_source': 'measured live: register-only FMA micro-benchmark mvn_fp32_peak '
                                        f'(FFMA {2 * peak_ffma / 1e12:.1f}, FFMA2 {2 * peak_ffma2 / 1e12:.1f} TFLOP/s)'},
                'hbm': {'achieved': HBM_BYTES_PER_SYMBOL * sym_per_launch / (k_ms * 1e-3) / 1e9, 'peak': hbm_peak,
                        'unit': 'GB/s', 'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback'}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        chunk, chunks = 16384, 40                          # ~10 s of CPU work on this box's 16 cores
        n_cpu = min(chunk * chunks, y0.shape[0])
        chunks = max(1, n_cpu // chunk)
        rate, secs = cpu_port_rate(chunk, chunks, w_np[snr0], y0[:chunk * chunks].cpu().numpy(), reps=1)
        cpu_baseline = {'value': rate, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                        'sample': f'first {chunk * chunks} frames x {T} symbols of the {snr0} dB item, chunks of {chunk}, '
                                  f'one pass ({secs:.1f} s), oracle/torch_port.py (op-for-op port of '
                                  'VNETDetector.forward val), torch ' + torch.__version__}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(world),
            'roofline': roofline, 'cpu_baseline': cpu_baseline,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': job_frames * T * 4,
                    'd2h_bytes_per_step': job_frames * T * 4, 'matches_device_path_all_rows': e2e_ok,
                    'api': 'mvn_ctx_set_vnet_weights_host + mvn_ctx_vnet_decode_host_async per work item, one mvn_ctx_synchronize per '
                           'step (pinned host y in, fp32 words out, chunks of two kernel waves on a ring of 6 streams).  Host-fabric bound: profiles/r02_copy_ceiling_8gpu.txt measures '
                           'the raw concurrent pinned-copy ceiling of this box type at N = 1/2/4/8 (46.8 / 53.8 / 53.3 / '
                           '83.0 GB/s per direction summed over the GPUs) and this pipeline at 95-97 % of it at every N',
                    'forms': {'fp32_words_out': {'value': e2e_value, 'h2d': job_frames * T * 4, 'd2h': job_frames * T * 4},
                              'bit_packed_words_out': {'value': e2e_bits_value, 'h2d': job_frames * T * 4,
                                                       'd2h': job_frames * ((T + 31) // 32) * 4},
                              'y_and_targets_in_counters_out': {'value': e2e_cnt_value, 'h2d': 2 * job_frames * T * 4,
                                                                'd2h': 32 * len(SNR_SWEEP)},
                              'device_source_counters_out': {'value': e2e_src_value, 'h2d': 32, 'd2h': 32 * len(SNR_SWEEP),
                                                             'note': 'mvn_ctx_vnet_sweep_point: words, channel and noise '
                                                                     'generated on the device (not an e2e number in the '
                                                                     'contract sense: no host inputs); the Monte-Carlo '
                                                                     'sweep shape of the product'}}},
            'gpu_launches': launches, 'clocks': clocks, 'ber_by_snr': ber_by_snr, 'sanity': sanity,
            'extras': extras}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

"""Supplementary measurements carried by bench.py's JSON line under "extras" (never inside its timed region):

  config1  reference-scale latency: ViterbiNet / VA forward of the drop-in detector classes at B = 1 (the eval_by_word
           shape, trainer.py:295) and B = 300 (val_frames = 12, plotter_main.py:96-111)
  config2  classical VA kernel, 2^20 frames x 120 (BASELINE.json configs[1]): reference rule and fused MLSE traceback
  config4  Meta-ViterbiNet on COST2100 taps (BASELINE.json configs[3]): batched MAML / FO-MAML / plain steps per second
           at R = 1, 148, 4096 with the FP32 roofline fraction (SURVEY.md §8d flop count), and end-to-end
           eval_by_word blocks/s of R runs advancing in lock step with the plotter's per-block schedule
  config5  memory-length sweep L = 3..8 of the fused kernel (BASELINE.json configs[4]); with N > 1 ranks the L values are
           sharded over the ranks and gathered

Everything is timed with CUDA events on the launching stream after a warm-up call.
"""
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
T = 120


def _timeit(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _net(torch, S, device, seed):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                              torch.nn.Linear(50, S))
    return [p.detach().to(device).contiguous() for p in net.parameters()]


def l_sweep_point(torch, mvn, _lib, device, L, fp32_peak_tflops, hbm_peak):
    import bench
    S = 2 ** L
    frames = bench.FRAMES if L <= 6 else bench.FRAMES // 4
    w = _net(torch, S, device, L)
    g = torch.Generator(device=device)
    g.manual_seed(L)
    y = torch.randn((frames, T), generator=g, device=device) * 1.5
    dec = torch.empty_like(y)
    lib = _lib.load()
    st = _lib.stream()

    def run():
        _lib.check(lib.mvn_vnet_decode(_lib.ptr(y), frames, T, L, T, *[_lib.ptr(a) for a in w], 0, _lib.ptr(dec), None,
                                       None, 0, 0, None, st))
    ms = _timeit(torch, run, 5)
    flop = 2 * (100 + 5000 + 50 * S) + 2 * S
    rate = frames * T / (ms * 1e-3)
    cost = torch.randn((frames // 8, T, S), generator=g, device=device)
    ms_acs = _timeit(torch, lambda: mvn.ops.acs_decode(cost), 5)
    acs_gbs = (frames // 8) * T * (4 * S + 4) / (ms_acs * 1e-3) / 1e9
    return {'memory_length': L, 'n_states': S, 'frames': frames, 'symbols_per_s': rate, 'ms': ms,
            'tflops_algorithmic': rate * flop / 1e12, 'frac_of_measured_fp32_peak': rate * flop / 1e12 / fp32_peak_tflops,
            'acs_decode_cost_tensor': {'frames': frames // 8, 'gb_per_s': acs_gbs, 'frac_of_measured_hbm': acs_gbs / hbm_peak}}


def run(torch, mvn, _lib, device, rank, world, dist):
    import json
    import bench
    from meta_viterbinet_b200.channel_taps import state_priors_table
    from meta_viterbinet_b200.train import pack_params
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    p0, _ = _lib.fp32_peak(0, 2048)
    p1, _ = _lib.fp32_peak(1, 2048)
    fp32_peak = 2 * max(p0, p1) / 1e12
    out = {}

    # ---- config5: L sweep, sharded over the ranks
    mine = [L for k, L in enumerate(range(3, 9)) if k % world == rank]
    sweep_rows = [l_sweep_point(torch, mvn, _lib, device, L, fp32_peak, hbm_peak) for L in mine]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, sweep_rows)
        sweep_rows = sorted([r for part in gathered for r in part], key=lambda r: r['memory_length'])
    if rank != 0:
        return None
    out['config5_memory_length_sweep'] = {'note': f'fused priors-MLP+ACS+decision kernel, random-init nets, T = {T}; one L per '
                                                  f'rank round-robin over {world} GPU(s); fractions are of the measured peaks',
                                          'fp32_peak_tflops_measured': fp32_peak, 'rows': sweep_rows}

    lib = _lib.load()
    st = _lib.stream()
    L, S = 4, 16
    # ---- config1: reference-scale latency through the drop-in classes
    g = np.load(bench.CKPT)
    w10 = [torch.as_tensor(g[f'snr10_w{i}']).to(device) for i in range(6)]
    det = mvn.VNETDetector(S, {'val': T, 'train': T})
    with torch.no_grad():
        for p, a in zip(det.parameters(), w10):
            p.copy_(a)
    h300 = np.exp(-0.2 * np.arange(L)).reshape(1, L) * (0.8 + 0.2 * np.cos(2 * np.pi * np.arange(300).reshape(-1, 1) / np.array([51, 39, 33, 21])))
    table300 = torch.as_tensor(state_priors_table(h300, L)).to(device)
    lat = {}
    for B in (1, 300):
        _, yb = bench.synth_frames(torch, device, B, 10, 5 + B)
        with torch.no_grad():
            ms = _timeit(torch, lambda: det(yb, 'val'), 50)
        lat[f'vnet_forward_B{B}_us'] = 1e3 * ms
        tb = table300[:B] if B > 1 else table300[:1]
        ms = _timeit(torch, lambda: mvn.ops.va_decode(yb, tb.contiguous()), 50)
        lat[f'va_forward_B{B}_us'] = 1e3 * ms
    lat['note'] = ('wall time per detector call incl. the Python wrapper, CUDA events, 50 calls; the reference takes 10.2 ms '
                   '(B=1) / 27.5 ms (B=300) for the ViterbiNet forward on 8 CPU cores (SURVEY.md §6)')
    out['config1_reference_scale_latency'] = lat

    # ---- config2: classical VA, 2^20 frames
    frames = bench.FRAMES
    bits, y = bench.synth_frames(torch, device, frames, 10, 1)
    table = torch.as_tensor(state_priors_table(np.exp(-0.2 * np.arange(L)).reshape(1, L), L)).to(device)
    dec = torch.empty_like(y)
    va = {}
    for name, decision in (('reference_rule', 0), ('mlse_traceback_fused', 2)):
        ms = _timeit(torch, lambda: _lib.check(lib.mvn_va_decode_ex(_lib.ptr(y), frames, T, L, T, _lib.ptr(table), 1, 0,
                                                                    _lib.ptr(dec), None, 0, 0, None, decision, st)), 5)
        va[name] = {'symbols_per_s': frames * T / (ms * 1e-3), 'ms': ms, 'hbm_gb_per_s': 8 * frames * T / (ms * 1e-3) / 1e9,
                    'frac_of_measured_hbm': 8 * frames * T / (ms * 1e-3) / 1e9 / hbm_peak,
                    'ber': float((dec[:, 1:] != bits[:, 1:]).float().mean())}
    va['note'] = f'{frames} frames x {T}, 16 states, static time_decay taps, 10 dB; 8 algorithmic HBM bytes per symbol'
    out['config2_va_kernel'] = va
    del dec, y, bits

    # ---- config4: batched meta-training throughput + end-to-end eval_by_word on COST2100 taps
    N = 136
    theta0 = pack_params(_net(torch, S, device, 0))
    steps = {}
    for R in (1, 148, 4096):
        tr = mvn.BatchedVNetTrainer(theta0.repeat(R, 1), L)
        ys, yq = torch.randn(R, N, device=device), torch.randn(R, N, device=device)
        ls = torch.randint(0, S, (R, N), device=device, dtype=torch.int32)
        lq = torch.randint(0, S, (R, N), device=device, dtype=torch.int32)
        reps = 20 if R <= 148 else 4
        t_maml = _timeit(torch, lambda: tr.meta_step(ys, ls, yq, lq, second_order=True), reps)
        t_fo = _timeit(torch, lambda: tr.meta_step(ys, ls, yq, lq, second_order=False), reps)
        t_sgd = _timeit(torch, lambda: tr.train_step(ys, ls), reps)
        steps[f'R{R}'] = {'maml_steps_per_s': R / t_maml * 1e3, 'fo_maml_steps_per_s': R / t_fo * 1e3,
                          'train_steps_per_s': R / t_sgd * 1e3, 'maml_ms_per_launch': t_maml,
                          'maml_tflops': R / t_maml * 1e3 * bench.MAML_FLOP_PER_STEP / 1e12,
                          'maml_frac_of_measured_fp32_peak': R / t_maml * 1e3 * bench.MAML_FLOP_PER_STEP / 1e12 / fp32_peak,
                          'train_frac_of_measured_fp32_peak': R / t_sgd * 1e3 * bench.TRAIN_FLOP_PER_STEP / 1e12 / fp32_peak}
        del tr
    taps = np.load(os.path.join(ROOT, 'tests', 'golden', 'cost2100_taps.npz'))['taps']
    R, n_blocks, nsym = 148, 30, 2
    info = mvn.ops.random_bits(R * n_blocks, 120, seed=3).reshape(R, n_blocks, 120)
    cw = mvn.ops.rs_encode(info.reshape(R * n_blocks, 120), nsym)
    h = np.tile(taps[:n_blocks], (R, 1))                                   # block c of every run sees COST2100 block c
    rx = mvn.ops.channel_transmit(cw, h, 10.0, seed=11).reshape(R, n_blocks, N)
    tr = mvn.BatchedVNetTrainer(theta0.repeat(R, 1), L, lr=1e-3, meta_lr=0.1)
    # supervised warm-up so that detection works and the SER gate opens (as the reference's load_weights would provide)
    lab_words = cw.reshape(R, n_blocks, N)
    for it in range(150):
        c = it % n_blocks
        tr.train_step(rx[:, c].contiguous(), lab_words[:, c].contiguous())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ser = mvn.online.eval_by_word(tr, info, rx, nsym, 0.02, subframes_in_frame=25, self_supervised=True, iterations=200,
                                  restart_from_saved=True, online_meta=True, meta_subframes=5, meta_train_iterations=20,
                                  meta_j_num=10, window_size=1, second_order=True, weights_init='last_frame')
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out['config4_meta_viterbinet_cost2100'] = {
        'batched_steps': steps,
        'flop_per_maml_step': bench.MAML_FLOP_PER_STEP, 'flop_per_train_step': bench.TRAIN_FLOP_PER_STEP,
        'flop_note': 'SURVEY.md §8d: fwd = 2 x 5900 x N flop at N = 136 symbols; plain step = fwd + bwd = 3 fwd; MAML step = '
                     '3 fwd(support) + 3 fwd(query) + Hessian-vector product 2 (fwd + bwd)(support) = 12 fwd',
        'eval_by_word': {'runs': R, 'blocks_per_run': n_blocks, 'seconds': dt, 'blocks_per_s': R * n_blocks / dt,
                         'mean_ser': float(ser.mean()),
                         'schedule': 'plotter_main.py:96-111 / config.yaml:45,55-57: self_supervised 200 iterations per gated '
                                     'block, online_meta every 5 blocks with 20 x <=10 MAML steps, ser_thresh 0.02, RS(2), '
                                     'COST2100 taps per block (resources/cost2100_channel), 10 dB',
                         'cpu_reference_note': 'the reference spends ~68 ms per MAML step and ~2.2 ms per online-training '
                                               'iteration on 8 CPU cores (SURVEY.md §6): >= 0.44 s per gated block for ONE run'}}
    return out

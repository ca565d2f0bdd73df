"""Concurrent host<->device copy ceiling of the box at N = 1, 2, 4, 8 GPUs (one process per GPU, all started together),
next to the product's host-buffer decode pipeline under the same concurrency.

    python tools/copy_ceiling_mp.py [--gpus 1,2,4,8] [--frames 1048576] [--reps 6] > profiles/r02_copy_ceiling.txt

Answers VERDICT r1 weak #1: is the e2e path (mvn_ctx_vnet_decode_host) limited by the host fabric or by the pipeline?
Rows: raw pinned copies (mvn_copy_ceiling: same chunk size and stream ring as the pipeline, no kernel) with buffers from
cudaHostAlloc (mvn_host_alloc), write-combined input, and torch.pin_memory(); then the decode pipeline itself in its
three output forms (fp32 words, bit-packed words, counters only) and the on-device Monte-Carlo source.
"""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T, L = 120, 4


def worker(rank, n, frames, reps, barrier, results):
    import numpy as np
    import torch
    import bench
    from meta_viterbinet_b200 import _lib
    torch.cuda.set_device(rank)
    lib = _lib.load()
    nbytes = frames * T * 4
    chunk_frames = 2 * 128 * _lib.device_info()['sm_count']
    chunk = chunk_frames * T * 4

    def alloc(wc):
        p = ctypes.c_void_p()
        _lib.check(lib.mvn_host_alloc(ctypes.byref(p), nbytes, wc))
        return p
    h_in, h_out, h_wc = alloc(0), alloc(0), alloc(1)
    ctypes.memset(h_in, 1, nbytes)
    ctypes.memset(h_wc, 1, nbytes)
    t_in = torch.empty(frames * T, dtype=torch.float32).pin_memory()
    t_out = torch.empty(frames * T, dtype=torch.float32).pin_memory()
    out = {}

    def raw(name, src, dst, h2d, d2h):
        sec = ctypes.c_double()
        barrier.wait()
        _lib.check(lib.mvn_copy_ceiling(rank, src, dst, nbytes, chunk, h2d, d2h, reps, ctypes.byref(sec)))
        out[name] = nbytes * reps / sec.value / 1e9          # GB/s per direction

    raw('raw h2d only (cudaHostAlloc)', h_in, None, 1, 0)
    raw('raw d2h only (cudaHostAlloc)', None, h_out, 0, 1)
    raw('raw both (cudaHostAlloc)', h_in, h_out, 1, 1)
    raw('raw both (write-combined in)', h_wc, h_out, 1, 1)
    raw('raw h2d only (write-combined)', h_wc, None, 1, 0)
    raw('raw both (torch pin_memory)', ctypes.c_void_p(t_in.data_ptr()), ctypes.c_void_p(t_out.data_ptr()), 1, 1)

    # the product pipeline under the same concurrency
    w = [p.cpu().contiguous() for p in bench.make_weights(torch, 'cpu')]
    y = torch.randn(frames, T)
    ctypes.memmove(h_in, ctypes.c_void_p(y.data_ptr()), nbytes)
    ctypes.memmove(h_wc, ctypes.c_void_p(y.data_ptr()), nbytes)
    tgt = alloc(0)
    ctypes.memset(tgt, 0, nbytes)
    ctx = ctypes.c_void_p()
    _lib.check(lib.mvn_ctx_create(ctypes.byref(ctx), rank, 0, T, L))
    _lib.check(lib.mvn_ctx_set_vnet_weights_host(ctx, *[ctypes.c_void_p(a.data_ptr()) for a in w]))
    cnt = (ctypes.c_uint64 * 4)()
    taps = (ctypes.c_double * L)(*[float(np.exp(-0.2 * i)) for i in range(L)])

    def pipe(name, fn, gbytes):
        fn()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        out[name] = frames * T / dt / 1e9                    # G symbols/s
        out[name + ' [GB/s h2d]'] = gbytes / dt
    pipe('pipeline fp32 words out', lambda: _lib.check(lib.mvn_ctx_vnet_decode_host(ctx, h_in, frames, T, T, 0, h_out)), nbytes / 1e9)
    pipe('pipeline fp32 words out, WC input', lambda: _lib.check(lib.mvn_ctx_vnet_decode_host(ctx, h_wc, frames, T, T, 0, h_out)), nbytes / 1e9)
    pipe('pipeline bit-packed out', lambda: _lib.check(lib.mvn_ctx_vnet_decode_host(ctx, h_in, frames, T, T, 1, h_out)), nbytes / 1e9)
    pipe('pipeline y+targets in, counters out', lambda: _lib.check(lib.mvn_ctx_vnet_eval_host(ctx, h_in, tgt, frames, T, T, T, 0, 0, None, cnt)), 2 * nbytes / 1e9)
    pipe('device source, counters out', lambda: _lib.check(lib.mvn_ctx_vnet_sweep_point(ctx, frames, T, T, taps, 1, 10.0, 7 + rank, 0, cnt)), 0.0)
    lib.mvn_ctx_destroy(ctx)
    results[rank] = out


def main():
    import multiprocessing as mp
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', default='1,2,4,8')
    ap.add_argument('--frames', type=int, default=1 << 20)
    ap.add_argument('--reps', type=int, default=6)
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    mpc = mp.get_context('spawn')
    table = {}
    ns = [int(x) for x in args.gpus.split(',') if int(x) <= have]
    for n in ns:
        barrier = mpc.Barrier(n)
        results = mpc.Manager().dict()
        procs = [mpc.Process(target=worker, args=(r, n, args.frames, args.reps, barrier, results)) for r in range(n)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        if any(p.exitcode for p in procs):
            print(f'N={n}: a worker failed: exit codes {[p.exitcode for p in procs]}')
            continue
        table[n] = dict(results)
    keys = list(next(iter(table.values()))[0].keys()) if table else []
    print(f'{args.frames} frames x {T} fp32 = {args.frames * T * 4 / 1e6:.0f} MB per direction per GPU per pass; {args.reps} passes; '
          f'raw rows: GB/s per direction, pipeline rows: G symbols/s.  "sum" = aggregate over the N concurrent GPUs, '
          f'"min" = slowest GPU.')
    print(f'{"":44s}' + ''.join(f'   N={n}: sum     min' for n in ns))
    for k in keys:
        line = f'{k:44s}'
        for n in ns:
            vals = [table[n][r][k] for r in range(n)] if n in table else [float("nan")]
            line += f'   {sum(vals):10.1f} {min(vals):7.1f}'
        print(line)


if __name__ == '__main__':
    main()

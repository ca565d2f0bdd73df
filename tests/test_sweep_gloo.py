"""N>1 host logic on CPU: world_size-2 gloo run of the sharded sweep.  The decode itself is injected
(the oracle stands in for the CUDA kernels, which need a GPU); what is checked is the sharding, the
single all-reduce of the counters and that every rank ends with the single-process totals."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meta_viterbinet_b200 import sweep
from oracle import viterbinet_oracle as orc

L, T, FRAMES = 3, 24, 40
SNRS = [6.0, 8.0, 10.0]


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _evaluate_block(snr, first, n, row):
    """Deterministic per-(snr, frame) data so that any sharding sees the same frames."""
    h = np.exp(-0.2 * np.arange(L)).reshape(1, L)
    for f in range(first, first + n):
        rng = np.random.RandomState(int(snr * 1000) + f)
        bits = rng.randint(0, 2, size=(1, T))
        y = orc.isi_awgn(bits, h, snr, L, rng).astype(np.float32)
        dec = orc.va_decode(y, h, L)
        be, fe, nb, nf, _ = orc.error_counts(dec, bits)
        row += torch.tensor([be, fe, nb, nf], dtype=torch.int64)


def _worker(rank, world, port, points, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    total = sweep.run_sweep(points, FRAMES, _evaluate_block)
    torch.save(total, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.parametrize('world,points', [(2, SNRS), (2, SNRS[:1]), (3, SNRS[:2])])
def test_sharded_sweep_matches_single_process(tmp_path, world, points):
    single = sweep.run_sweep(points, FRAMES, _evaluate_block, rank=0, world_size=1)
    assert single[:, 3].tolist() == [FRAMES] * len(points)
    mp.spawn(_worker, args=(world, _free_port(), points, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f'rank{r}.pt'))
        assert torch.equal(got, single)
    ber, fer = sweep.rates(single)
    assert ber.shape == (len(points),) and float(ber.max()) <= 1.0


def test_partition_and_work_items():
    for n in (0, 1, 5, 17):
        for w in (1, 2, 3, 8):
            parts = [sweep.partition(n, w, r) for r in range(w)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sweep.work_items(6, 8) == [(i, b, 4) for i in range(6) for b in range(4)]     # 24 items, 3 per rank
    assert sweep.work_items(6, 4) == [(i, b, 2) for i in range(6) for b in range(2)]     # 12 items, 3 per rank
    assert sweep.work_items(6, 3) == [(i, 0, 1) for i in range(6)]                       # whole points, 2 per rank
    for points in range(1, 13):                                   # every rank gets exactly the same number of items
        for w in (1, 2, 3, 4, 6, 8):
            items = sweep.work_items(points, w)
            assert len({len(sweep.partition(len(items), w, r)) for r in range(w)}) == 1, (points, w)


def test_before_reduce_hook_runs_once_after_the_last_block():
    """run_sweep(before_reduce=...) — where a caller that spreads its launches over several streams joins them"""
    calls = []

    def block(point, first, n, row):
        calls.append(('block', point, first, n))
        row += torch.tensor([0, 0, n * 10, n])

    total = sweep.run_sweep([7, 8, 9], 100, block, rank=0, world_size=1, before_reduce=lambda: calls.append(('join',)))
    assert calls[-1] == ('join',) and sum(1 for c in calls if c[0] == 'join') == 1 and len(calls) == 4
    assert total[:, 3].tolist() == [100, 100, 100]

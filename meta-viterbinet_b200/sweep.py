"""Multi-GPU evaluation sweeps: frames, SNR points and channel realisations are independent, so the
work is sharded over ranks with NO data-path collective; the only exchange is one all-reduce (sum)
of the [n_points, 4] int64 error counters (bit errors, frame errors, bits, frames) at the end
(SURVEY.md §8e).  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The reference evaluates one SNR point after the other in one process (trainer.py:243-252,
gamma_eval); this module is the B200 replacement of that loop for large Monte-Carlo runs.
"""
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def partition(n_items: int, world_size: int, rank: int) -> range:
    """Balanced contiguous partition of range(n_items): the first n_items % world_size ranks get one more."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def work_items(n_points: int, world_size: int) -> List[Tuple[int, int, int]]:
    """(point index, block index, blocks per point), ordered so that partition() hands every rank the same number of
    items.  Whole points per rank when they divide evenly over the ranks (weights differ per SNR point,
    trainer.py:511, so a rank then stages each point's weights once); otherwise every point's frames are split into
    world / gcd(points, world) row blocks: points * blocks is the least common multiple of points and world, i.e. the
    coarsest split in which all ranks get exactly the same share (6 points on 8 ranks: 4 blocks per point, 3 items
    per rank — not 2 blocks, which would leave half the ranks with one item and the others with two)."""
    import math
    blocks = 1 if n_points % world_size == 0 else world_size // math.gcd(n_points, world_size)
    return [(i, b, blocks) for i in range(n_points) for b in range(blocks)]


def all_reduce_counters(counters: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the int64 counters over ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def run_sweep(points: Sequence, frames_per_point: int,
              evaluate_block: Callable[[object, int, int, torch.Tensor], None],
              device=None, rank: int = None, world_size: int = None, group=None,
              before_reduce: Callable[[], None] = None) -> torch.Tensor:
    """Evaluate every point on `frames_per_point` frames, sharded over the ranks.

    evaluate_block(point, first_frame, n_frames, counters_row) must ADD the four counts of frames
    [first_frame, first_frame + n_frames) of that point into counters_row (an int64[4] view) — e.g. by
    calling ops.vnet_decode(..., target=..., counters=counters_row).
    before_reduce() runs after this rank's last evaluate_block and before the all-reduce: a caller that spreads its
    launches over several CUDA streams joins them there.
    Returns the all-reduced [len(points), 4] int64 tensor (identical on every rank).
    """
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    counters = torch.zeros((len(points), 4), dtype=torch.int64, device=device)
    items = work_items(len(points), world_size)
    for k in partition(len(items), world_size, rank):
        i, b, blocks = items[k]
        rows = partition(frames_per_point, blocks, b)
        if len(rows):
            evaluate_block(points[i], rows.start, len(rows), counters[i])
    if before_reduce is not None:
        before_reduce()
    return all_reduce_counters(counters, group)


def rates(counters: torch.Tensor):
    """[n,4] counters -> (ber [n], fer [n]) as float64 tensors (exact integer ratios)."""
    c = counters.to(torch.float64).cpu()
    return c[:, 0] / c[:, 2].clamp(min=1), c[:, 1] / c[:, 3].clamp(min=1)

"""Drop-in for python_code/utils/metrics.py."""
from typing import Tuple

import numpy as np
import torch

from .. import ops


def calculate_error_rates(prediction: torch.Tensor, target: torch.Tensor) -> Tuple[float, float, torch.Tensor]:
    """Returns the ber, fer and error indices (metrics.py:7-17).

    The kernel accumulates exact integer counts; the rates are then formed the way the reference's
    fp32 ``torch.mean`` does (count / N in fp32, 1 - x in double, clamped at 0), so they are
    identical to the reference's floats while N < 2^24 elements.
    """
    counters, rows = ops.error_counts(prediction, target)
    bit_errs, frame_errs, bits, frames = [int(v) for v in counters.cpu().tolist()]
    bits_acc = float(np.float32(bits - bit_errs) / np.float32(bits))
    frames_acc = float(np.float32(frames - frame_errs) / np.float32(frames))
    err_idx = torch.nonzero(rows, as_tuple=False).reshape(-1)
    return max([1 - bits_acc, 0.0]), max([1 - frames_acc, 0.0]), err_idx

"""Host-side channel taps for the full-CSI Viterbi detector.

Vectorised restatement of ``estimate_channel`` (reference: python_code/channel/channel_estimation.py:11-49):
the reference calls it once per block inside every VADetector.forward (va_detector.py:54-58), which
is where 99 % of its VA time goes.  Here all blocks are evaluated in one numpy
expression in float64, the COST2100 tap files are read once, and static tables are cached.
Input generation is not on the GPU hot path; only the resulting [n_h, S] table is.
"""
import os
from functools import lru_cache

import numpy as np

COST_LENGTH = 300  # channel_estimation.py:8
COST2100_DIR = os.environ.get('MVN_COST2100_DIR', '')


@lru_cache(maxsize=4)
def _cost2100_table(directory: str, memory_length: int) -> np.ndarray:
    import scipy.io
    if not directory or not os.path.isdir(directory):
        raise ValueError('cost2100 taps requested but COST2100_DIR is not set '
                         '(set meta_viterbinet_b200.channel_taps.COST2100_DIR or $MVN_COST2100_DIR)')
    total = np.empty([COST_LENGTH, memory_length])
    for i in range(memory_length):
        for stem in (f'combined_h_{i}', f'h_{i}'):  # the code loads combined_h_i, the repo ships h_i
            path = os.path.join(directory, stem + '.mat')
            if os.path.exists(path):
                total[:, i] = scipy.io.loadmat(path)['h_channel_response_mag'].reshape(-1)
                break
        else:
            raise ValueError(f'no tap file for tap {i} in {directory}')
    return total


def channel_taps(memory_length: int, gamma: float, channel_coefficients: str, noisy_est_var: float = 0,
                 fading: bool = False, indices=0, fading_taps_type: int = 1) -> np.ndarray:
    """Taps for every block index in ``indices`` -> float64 [n, memory_length]."""
    idx = np.atleast_1d(np.asarray(indices))
    n = idx.shape[0]
    if channel_coefficients == 'time_decay':
        h = np.tile(np.exp(-gamma * np.arange(memory_length)).reshape(1, memory_length), (n, 1))
    elif channel_coefficients == 'cost2100':
        h = _cost2100_table(COST2100_DIR, memory_length)[idx].copy()
    else:
        raise ValueError('No such channel_coefficients value!!!')
    if noisy_est_var > 0:
        # same stream as the per-block draws of the reference (global numpy RNG, legacy normal())
        h[:, 1:] += np.random.normal(0, noisy_est_var ** 0.5, [n, memory_length - 1])
    if fading and channel_coefficients == 'time_decay':
        col = idx.reshape(-1, 1).astype(np.float64)
        if fading_taps_type == 1:
            periods = np.array([51, 39, 33, 21])
            h *= (0.8 + 0.2 * np.cos(2 * np.pi * col / periods)).reshape(n, memory_length)
        elif fading_taps_type == 2:
            periods = 5 * np.array([51, 39, 33, 21])
            periods = np.maximum(periods - 1.5 * col, 10 * np.ones(4)) - 1e-5
            h *= (0.8 + 0.2 * np.cos(np.pi * col / periods)).reshape(n, memory_length)
        else:
            raise ValueError("No such fading tap type!!!")
    return h


def estimate_channel(memory_length: int, gamma: float, channel_coefficients: str, noisy_est_var: float = 0,
                     fading: bool = False, index: int = 0, fading_taps_type: int = 1) -> np.ndarray:
    """Same signature and result ([1, memory_length]) as the reference's estimate_channel."""
    return channel_taps(memory_length, gamma, channel_coefficients, noisy_est_var, fading, [index], fading_taps_type)


def state_priors_table(h: np.ndarray, memory_length: int) -> np.ndarray:
    """Noiseless channel output per (tap block, state) -> fp32 [n_h, S].

    Same numpy expression as va_detector.py:42-50 (MSB-first state bits, BPSK 1-2b, float64 dot,
    cast to fp32) so the table is bit-identical; returned transposed ([n_h, S]) for the kernel.
    """
    if memory_length > 8:
        raise ValueError('memory_length > 8 does not fit the uint8 state expansion of the reference')
    n_states = 2 ** memory_length
    states = np.arange(n_states).astype(np.uint8).reshape(-1, 1)
    bits = np.unpackbits(states, axis=1).astype(int)[:, -memory_length:]
    symbols = 1 - 2 * bits
    table = np.dot(symbols, np.asarray(h, dtype=np.float64).T)   # [S, n_h] float64
    return np.ascontiguousarray(table.T.astype(np.float32))

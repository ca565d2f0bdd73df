"""Drop-in for python_code/utils/trellis_utils.py (create_transition_table, acs_block,
calculate_states) backed by the CUDA kernels."""
import numpy as np
import torch

from .. import ops


def create_transition_table(n_states: int) -> np.ndarray:
    """[n_states, 2]: previous state of state i and input bit b (trellis_utils.py:7-13):
    row j = [(2j) mod S, (2j+1) mod S]."""
    j = np.arange(n_states)
    return np.stack([(2 * j) % n_states, (2 * j + 1) % n_states], axis=1)


def acs_block(in_prob: torch.Tensor, llrs: torch.Tensor, transition_table: torch.Tensor, n_states: int):
    """Viterbi ACS block (trellis_utils.py:16-30) -> (values, indices) like torch.min(dim=2).
    ``transition_table`` is accepted for signature compatibility; the trellis is the reference's."""
    return ops.acs_block(in_prob, llrs, n_states)


def calculate_states(memory_length: int, transmitted_words: torch.Tensor) -> torch.Tensor:
    """Ground-truth state per symbol, flattened [B*T] int64 (trellis_utils.py:33-46)."""
    return ops.calculate_states(memory_length, transmitted_words)

"""Drop-in detector classes: same constructors, ``forward`` signatures and error behaviour as the
reference's ``VADetector`` (python_code/detectors/VA/va_detector.py), ``VNETDetector``
(python_code/detectors/VNET/vnet_detector.py) and ``META_VNETDetector``
(python_code/detectors/META_VNET/meta_vnet_detector.py); the work is done by the sm_100a kernels
behind the C ABI.  There is no torch/CPU fallback for the 'val' path.
"""
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import require_cuda
from .autograd import priors_function
from .channel_taps import channel_taps, state_priors_table
from .utils.trellis_utils import create_transition_table

HIDDEN1_SIZE = 100
HIDDEN2_SIZE = 50
# Below this many frames the fused one-lane-per-frame kernel cannot fill the GPU; the priors are
# then computed symbol-parallel and fed to the ACS kernel (both CUDA, same arithmetic).
SMALL_BATCH_FRAMES = 2048


_packed_cache = {}


def _packed(weights):
    """[P] packed copy of the six weight tensors, re-packed only when one of them changed (tensor identity + version)."""
    key = tuple((w.data_ptr(), w._version) for w in weights)
    hit = _packed_cache.get('k')
    if hit is None or hit[0] != key:
        from .train import pack_params
        hit = (key, pack_params(weights))
        _packed_cache['k'] = hit
    return hit[1]


def _decode_vnet(y, weights, n_stages):
    if n_stages > y.shape[1]:
        raise IndexError(f'index {y.shape[1]} is out of bounds for dimension 1 with size {y.shape[1]}')
    if y.shape[0] >= SMALL_BATCH_FRAMES:
        return ops.vnet_decode(y, weights, n_stages)
    L = int(weights[5].numel()).bit_length() - 1
    if L <= 5:       # one launch: a CTA per word, priors symbol-parallel into shared memory, then the stage loop
        from .train import detect_small
        return detect_small(y, _packed(weights), L, n_stages)
    priors = ops.vnet_priors(y, weights)
    return ops.acs_decode(-priors, n_stages)


class VADetector(nn.Module):
    """Classic Viterbi with full CSI (va_detector.py:13-100)."""

    def __init__(self, n_states: int, memory_length: int, transmission_length: int, val_words: int,
                 channel_type: str, noisy_est_var: float, fading: bool, fading_taps_type: int,
                 channel_coefficients: str):
        super(VADetector, self).__init__()
        self.memory_length = memory_length
        self.transmission_length = transmission_length
        self.val_words = val_words
        self.n_states = n_states
        self.channel_type = channel_type
        self.noisy_est_var = noisy_est_var
        self.fading = fading
        self.fading_taps_type = fading_taps_type
        self.channel_coefficients = channel_coefficients
        self.transition_table_array = create_transition_table(n_states)
        self.transition_table = torch.Tensor(self.transition_table_array).to(require_cuda())
        self._table_cache = {}

    def compute_state_priors(self, h: np.ndarray) -> torch.Tensor:
        """[S, n_h] fp32 like the reference (va_detector.py:42-50)."""
        if self.channel_type != 'ISI_AWGN':
            raise Exception('No such channel defined!!!')
        return torch.from_numpy(state_priors_table(h, self.memory_length).T.copy()).to(require_cuda())

    def _taps(self, gamma, phase):
        return channel_taps(self.memory_length, gamma, noisy_est_var=self.noisy_est_var, fading=self.fading,
                            indices=np.arange(self.val_words), fading_taps_type=self.fading_taps_type,
                            channel_coefficients=self.channel_coefficients[phase])

    def _table(self, gamma, phase, count):
        """Device table [n_h, S]; cached when the taps are deterministic (noisy_est_var == 0)."""
        if self.channel_type != 'ISI_AWGN':
            raise Exception('No such channel defined!!!')
        key = (gamma, phase, self.channel_coefficients[phase], self.fading, self.fading_taps_type)
        full = self._table_cache.get(key) if self.noisy_est_var == 0 else None
        if full is None:
            full = torch.from_numpy(state_priors_table(self._taps(gamma, phase), self.memory_length)).to(require_cuda())
            if self.noisy_est_var == 0:
                self._table_cache[key] = full
        return full if count is None else full[count:count + 1].contiguous()

    def forward(self, y: torch.Tensor, phase: str, snr: float = None, gamma: float = None,
                count: int = None) -> torch.Tensor:
        table = self._table(gamma, phase, count)
        if phase == 'val':
            if y.shape[0] % table.shape[0] != 0:
                raise RuntimeError(f'The size of tensor a ({y.shape[0]}) must match the size of tensor b '
                                   f'({(y.shape[0] // table.shape[0]) * table.shape[0]}) at non-singleton dimension 0')
            return ops.va_decode(y, table, self.transmission_length)
        else:
            raise NotImplementedError("No implemented training for this decoder!!!")


class VNETDetector(nn.Module):
    """ViterbiNet (vnet_detector.py:11-63).  ``net`` keeps the reference's parameter names/shapes
    (net.0/2/4.{weight,bias}) so checkpoints, optimizers and copy_model work unchanged."""

    def __init__(self, n_states: int, transmission_lengths: Dict[str, int]):
        super(VNETDetector, self).__init__()
        self.transmission_lengths = transmission_lengths
        self.n_states = n_states
        self.transition_table_array = create_transition_table(n_states)
        self.transition_table = torch.Tensor(self.transition_table_array).to(require_cuda())
        self.initialize_dnn()

    def initialize_dnn(self):
        layers = [nn.Linear(1, HIDDEN1_SIZE), nn.Sigmoid(), nn.Linear(HIDDEN1_SIZE, HIDDEN2_SIZE), nn.ReLU(),
                  nn.Linear(HIDDEN2_SIZE, self.n_states)]
        self.net = nn.Sequential(*layers).to(require_cuda())

    def weights(self):
        return [self.net[0].weight, self.net[0].bias, self.net[2].weight, self.net[2].bias,
                self.net[4].weight, self.net[4].bias]

    def forward(self, y: torch.Tensor, phase: str, snr: float = None, gamma: float = None,
                count: int = None) -> torch.Tensor:
        if phase == 'val':
            with torch.no_grad():
                return _decode_vnet(y, [w.detach() for w in self.weights()], self.transmission_lengths['val'])
        else:
            return priors_function(y, *self.weights())


class META_VNETDetector(nn.Module):
    """Functional ViterbiNet: the weights are passed per call (meta_vnet_detector.py:11-47)."""

    def __init__(self, n_states: int, transmission_lengths: Dict[str, int]):
        super(META_VNETDetector, self).__init__()
        self.transmission_lengths = transmission_lengths
        self.n_states = n_states
        self.transition_table_array = create_transition_table(n_states)
        self.transition_table = torch.Tensor(self.transition_table_array).to(require_cuda())

    def forward(self, y: torch.Tensor, phase: str, var: list) -> torch.Tensor:
        if phase == 'val':
            with torch.no_grad():
                return _decode_vnet(y, [w.detach() for w in var], self.transmission_lengths['val'])
        else:
            return priors_function(y, *var)

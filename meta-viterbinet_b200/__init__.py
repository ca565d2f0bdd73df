"""meta-viterbinet_b200 — B200-native detection hot path of Meta-ViterbiNet.

Import as ``meta_viterbinet_b200`` (the repo-root shim maps the importable name to this
directory, whose name contains a '-').
"""
from . import _lib, ops
from . import train
from . import sweep
from . import online
from . import ecc
from ._lib import MVNError, OUT_BITS, OUT_F32
from .train import BatchedVNetTrainer
from .detectors import META_VNETDetector, VADetector, VNETDetector
from .utils.metrics import calculate_error_rates
from .utils.trellis_utils import acs_block, calculate_states, create_transition_table

__all__ = ['VADetector', 'VNETDetector', 'META_VNETDetector', 'acs_block', 'calculate_states',
           'create_transition_table', 'calculate_error_rates', 'BatchedVNetTrainer', 'train', 'sweep', 'online', 'ecc', 'ops', 'MVNError', 'OUT_F32', 'OUT_BITS']

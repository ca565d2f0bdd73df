"""Reed-Solomon codec with the reference's call signatures (python_code/ecc/rs_main.py:9-37), on the GPU.

``encode(binary_word, nsym)`` / ``decode(binary_rx, nsym)`` take and return numpy bit arrays like the reference
(one word, 1-D) and additionally accept a batch ``[B, bits]`` (numpy or torch); torch CUDA inputs stay on the device.
"""
import numpy as np
import torch

from . import ops


def _run(fn, bits, nsym):
    is_np = not torch.is_tensor(bits)
    x = torch.as_tensor(np.asarray(bits)) if is_np else bits
    single = x.dim() == 1
    out = fn(x.reshape(1, -1) if single else x, nsym)
    out = out[0] if single else out
    return out.cpu().numpy().astype(int) if is_np else out


def encode(binary_word, nsym: int):
    """rs_main.py:9-18: message bits -> codeword bits (message followed by 8*nsym parity bits)."""
    return _run(ops.rs_encode, binary_word, nsym)


def decode(binary_rx, nsym: int):
    """rs_main.py:21-37: received codeword bits -> message bits (unchanged message part when uncorrectable)."""
    return _run(ops.rs_decode, binary_rx, nsym)

"""ctypes binding of the C ABI declared in ``include/mvn_b200.h``.

There is no CPU fallback: if ``libmvn_b200.so`` is missing or a tensor is not on a CUDA device the
call raises.  torch is used only for device memory and the current stream.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p, POINTER

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# MVN_LIB: load another build of the same library (kernel tuning: tools/ compare variants built with -D flags)
LIB_PATH = os.environ.get('MVN_LIB') or os.path.join(HERE, 'libmvn_b200.so')

OUT_F32 = 0
OUT_BITS = 1


class MVNError(RuntimeError):
    """A C-ABI call returned non-zero; the message is mvn_last_error()."""


_lib = None

_PROTOS = {
    'mvn_last_error': (c_char_p, []),
    'mvn_version': (c_int, []),
    'mvn_device_info': (c_int, [POINTER(c_int)] * 4),
    'mvn_acs_block': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    'mvn_acs_decode': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'mvn_acs_decode_ex': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'mvn_va_decode': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                              c_int, c_int, c_void_p, c_void_p]),
    'mvn_vnet_priors': (c_int, [c_void_p, c_int64, c_int] + [c_void_p] * 6 + [c_void_p, c_void_p]),
    'mvn_vnet_decode': (c_int, [c_void_p, c_int64, c_int, c_int, c_int] + [c_void_p] * 6 +
                        [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'mvn_va_decode_ex': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                 c_int, c_int, c_void_p, c_int, c_void_p]),
    'mvn_vnet_decode_ex': (c_int, [c_void_p, c_int64, c_int, c_int, c_int] + [c_void_p] * 6 +
                           [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    'mvn_tc_timeout_status': (c_int, []),
    'mvn_reset_tc_timeout': (c_int, []),
    'mvn_calculate_states': (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    'mvn_error_counts': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'mvn_ctx_create': (c_int, [POINTER(c_void_p), c_int, c_int64, c_int, c_int]),
    'mvn_ctx_destroy': (None, [c_void_p]),
    'mvn_ctx_set_vnet_weights_host': (c_int, [c_void_p] * 7),
    'mvn_ctx_vnet_decode_host': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    'mvn_ctx_vnet_decode_host_async': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    'mvn_ctx_synchronize': (c_int, [c_void_p]),
    'mvn_ctx_va_decode_host': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    'mvn_ctx_set_variant': (c_int, [c_void_p, c_int]),
    'mvn_ctx_set_decision': (c_int, [c_void_p, c_int]),
    'mvn_ctx_vnet_eval_host': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p]),
    'mvn_ctx_vnet_sweep_point': (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_double, c_uint64, c_int,
                                         c_void_p]),
    'mvn_host_alloc': (c_int, [POINTER(c_void_p), c_size_t, c_int]),
    'mvn_host_free': (c_int, [c_void_p]),
    'mvn_copy_ceiling': (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_int, c_int, POINTER(c_double)]),
    'mvn_launch_count': (c_int64, [c_int]),
    # include/mvn_b200_next.h
    'mvn_channel_transmit': (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_double, c_void_p, c_uint64,
                                     c_void_p, c_void_p]),
    'mvn_random_bits': (c_int, [c_void_p, c_int64, c_int, c_uint64, c_void_p]),
    'mvn_traceback': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'mvn_va_cost': (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'mvn_rs_decode': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'mvn_rs_encode': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    'mvn_fp32_peak': (c_int, [c_int, c_int, POINTER(c_double), POINTER(c_double), c_void_p]),
}


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python meta-viterbinet_b200/build.py` '
                '(there is no CPU fallback for the detection path)')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_PROTOS)


def check(rc):
    if rc != 0:
        raise MVNError(load().mvn_last_error().decode() or f'error code {rc}')


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('meta-viterbinet_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def dev_f32(t, name='tensor'):
    """Contiguous fp32 CUDA view/copy of t (moves host tensors to the current device)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    dev = require_cuda()
    if t.device.type != 'cuda':
        t = t.to(dev)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count(reset=False):
    return int(load().mvn_launch_count(1 if reset else 0))


def device_info():
    v = [c_int() for _ in range(4)]
    check(load().mvn_device_info(*[ctypes.byref(x) for x in v]))
    return dict(sm_count=v[0].value, sm_clock_khz=v[1].value, cc=(v[2].value, v[3].value))


def fp32_peak(mode=1, iters=4096):
    """Measured register-only FMA rate (FMA lane-ops/s).  mode 0 scalar FFMA, 1 packed FFMA2."""
    rate, ms = c_double(), c_double()
    check(load().mvn_fp32_peak(mode, iters, ctypes.byref(rate), ctypes.byref(ms), stream()))
    return rate.value, ms.value

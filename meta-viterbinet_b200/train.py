"""Batched (meta-)training of the ViterbiNet priors network on the GPU.

``BatchedVNetTrainer`` runs R independent realisations side by side (one CTA each): every step of the
reference's online loop depends on the previous one through the weights and the Adam state, so
the only parallelism is across runs (SNR points, channel realisations, seeds).

Reference semantics
  * meta_step      Trainer.meta_train_loop                  trainers/trainer.py:425-453
  * train_step     Trainer.run_train_loop / online_training trainers/trainer.py:492-505,
                                                            trainers/META_VNET/metavnet_trainer.py:52-64
  * loss           CrossEntropyLoss over all symbols        trainers/META_VNET/metavnet_trainer.py:41-50
  * labels         calculate_states                         utils/trellis_utils.py:33-46
  * optimiser      torch.optim.Adam defaults                trainers/trainer.py:167-169
"""
import ctypes
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

from . import _lib, ops
from ._lib import check, dev_f32, load, ptr, stream

_PROTOS = {
    'mvn_param_count': (c_int, [c_int]),
    'mvn_meta_workspace_bytes': (c_int64, [c_int, c_int, c_int]),
    'mvn_meta_step_batched': (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                                        c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                                        c_void_p]),
    'mvn_train_step_batched': (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                                         c_void_p, c_void_p, c_void_p]),
    'mvn_priors_backward_workspace_bytes': (c_int64, [c_int, c_int64]),
    'mvn_vnet_priors_backward': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'mvn_vnet_priors_backward2': (c_int, [c_void_p, c_int64, c_int] + [c_void_p] * 7),
    'mvn_vnet_detect_batched': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'mvn_vnet_detect_small': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}
_lib._PROTOS.update(_PROTOS)
if _lib._lib is not None:          # library already loaded: bind the extra prototypes now
    for _n, (_r, _a) in _PROTOS.items():
        _f = getattr(_lib._lib, _n)
        _f.restype, _f.argtypes = _r, _a


def param_count(memory_length: int) -> int:
    return int(load().mvn_param_count(memory_length))


def pack_params(weights) -> torch.Tensor:
    """[W1,b1,W2,b2,W3,b3] -> flat fp32 vector in torch parameter order."""
    return torch.cat([dev_f32(w).reshape(-1) for w in weights])


def unpack_params(theta: torch.Tensor, n_states: int):
    shapes = [(100, 1), (100,), (50, 100), (50,), (n_states, 50), (n_states,)]
    out, o = [], 0
    for s in shapes:
        n = 1
        for d in s:
            n *= d
        out.append(theta[..., o:o + n].reshape(theta.shape[:-1] + s))
        o += n
    return out


def detect_small(y, theta, memory_length, n_stages=None):
    """Single-launch VNETDetector.forward(y, 'val') for small batches: y [B,T], theta [P] packed weights."""
    y = dev_f32(y)
    B, T = y.shape
    n = T if n_stages is None else int(n_stages)
    dec = torch.empty((B, T), dtype=torch.float32, device=y.device)
    check(load().mvn_vnet_detect_small(ptr(theta), B, int(memory_length), ptr(y), T, n, ptr(dec), None, stream()))
    return dec


def state_labels(memory_length: int, tx: torch.Tensor) -> torch.Tensor:
    """int32 labels [R, N] for transmitted words tx [R, N] (each row one realisation's word(s))."""
    R, N = tx.shape
    return ops.calculate_states(memory_length, tx).reshape(R, N).to(torch.int32)


class BatchedVNetTrainer:
    """R independent copies of (theta, Adam state) trained side by side on one GPU."""

    def __init__(self, theta: torch.Tensor, memory_length: int, lr: float = 1e-3, meta_lr: float = 0.1):
        self.L = int(memory_length)
        self.S = 2 ** self.L
        self.P = param_count(self.L)
        theta = dev_f32(theta)
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        if theta.shape[1] != self.P:
            raise ValueError(f'theta must be [R, {self.P}]')
        self.theta = theta.clone().contiguous()
        self.R = theta.shape[0]
        self.lr, self.meta_lr = float(lr), float(meta_lr)
        self.reset_optimizer()
        nbytes = int(load().mvn_meta_workspace_bytes(self.L, self.R, 0))
        if nbytes < 0:
            raise _lib.MVNError(f'training kernels support memory_length 1..5, got {self.L}')
        self.workspace = torch.empty(nbytes // 4, dtype=torch.float32, device=self.theta.device)

    def reset_optimizer(self):
        """deep_learning_setup(): fresh Adam state (trainer.py:163-169)."""
        self.adam_m = torch.zeros_like(self.theta)
        self.adam_v = torch.zeros_like(self.theta)
        self.adam_step = torch.zeros(self.R, dtype=torch.int32, device=self.theta.device)

    def weights(self, r: int = 0):
        return unpack_params(self.theta[r], self.S)

    def _prep(self, y, tx):
        y = dev_f32(y).reshape(self.R, -1)
        lab = state_labels(self.L, dev_f32(tx).reshape(self.R, -1)) if tx.dtype != torch.int32 else tx
        return y, lab.contiguous()

    def train_step(self, y, tx, update=True, return_grad=False):
        """One run_train_loop iteration per realisation.  y, tx: [R, N].  Returns loss [R] (+ grad [R,P])."""
        y, lab = self._prep(y, tx)
        loss = torch.empty(self.R, dtype=torch.float32, device=y.device)
        grad = torch.empty_like(self.theta) if return_grad else None
        check(load().mvn_train_step_batched(ptr(self.theta), ptr(self.adam_m) if update else None, ptr(self.adam_v),
                                            ptr(self.adam_step), self.R, self.L, ptr(y), ptr(lab), y.shape[1], self.lr,
                                            ptr(loss), ptr(grad), ptr(self.workspace), stream()))
        return (loss, grad) if return_grad else loss

    def detect(self, y, n_stages=None, return_priors=False):
        """VNETDetector.forward(y, 'val') for every realisation with ITS OWN current weights, one launch.
        y [R, T] -> decoded [R, T] fp32 0/1 (and the priors [R, T, S] the decisions were taken on)."""
        y = dev_f32(y).reshape(self.R, -1)
        T = y.shape[1]
        n = T if n_stages is None else int(n_stages)
        dec = torch.empty((self.R, T), dtype=torch.float32, device=y.device)
        pri = torch.empty((self.R, T, 1 << self.L), dtype=torch.float32, device=y.device) if return_priors else None
        check(load().mvn_vnet_detect_batched(ptr(self.theta), self.R, self.L, ptr(y), T, n, ptr(dec), ptr(pri), stream()))
        return (dec, pri) if return_priors else dec

    def meta_step(self, y_s, tx_s, y_q, tx_q, second_order=True, update=True, return_grad=False):
        """One meta_train_loop step per realisation.  Support [R, Ns], query [R, Nq].  Returns query loss [R]."""
        y_s, lab_s = self._prep(y_s, tx_s)
        y_q, lab_q = self._prep(y_q, tx_q)
        loss = torch.empty(self.R, dtype=torch.float32, device=y_s.device)
        grad = torch.empty_like(self.theta) if return_grad else None
        check(load().mvn_meta_step_batched(ptr(self.theta), ptr(self.adam_m) if update else None, ptr(self.adam_v),
                                           ptr(self.adam_step), self.R, self.L, ptr(y_s), ptr(lab_s), y_s.shape[1],
                                           ptr(y_q), ptr(lab_q), y_q.shape[1], self.meta_lr, self.lr,
                                           1 if second_order else 0, ptr(loss), ptr(grad), ptr(self.workspace),
                                           stream()))
        return (loss, grad) if return_grad else loss


def _bwd_args(y, weights, grad_priors):
    S = weights[5].numel()
    L = S.bit_length() - 1
    theta = pack_params(weights).contiguous()
    g = dev_f32(grad_priors).reshape(-1, S)
    yf = dev_f32(y).reshape(-1)
    nbytes = int(load().mvn_priors_backward_workspace_bytes(L, yf.numel()))
    if nbytes < 0:
        raise _lib.MVNError(f'priors backward supports memory_length 1..5, got {L}')
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=yf.device)
    return S, L, theta, g, yf, ws


def priors_backward(y, weights, grad_priors):
    """Backward of the 'train'-phase priors: gradients w.r.t. [W1,b1,W2,b2,W3,b3] (shaped like them)."""
    S, L, theta, g, yf, ws = _bwd_args(y, weights, grad_priors)
    gt = torch.empty_like(theta)
    check(load().mvn_vnet_priors_backward(ptr(yf), yf.numel(), L, ptr(theta), ptr(g), ptr(gt), ptr(ws), stream()))
    return tuple(a.reshape(w.shape) for a, w in zip(unpack_params(gt, S), weights))


def priors_backward2(y, weights, grad_priors, upstream):
    """Backward of priors_backward (MAML, trainer.py:437 create_graph=True).  upstream: six tensors shaped like
    the weights.  Returns (grad w.r.t. grad_priors, six grads w.r.t. the weights)."""
    S, L, theta, g, yf, ws = _bwd_args(y, weights, grad_priors)
    u = pack_params(upstream).contiguous()
    gt2 = torch.empty_like(theta)
    ggp = torch.empty_like(g)
    check(load().mvn_vnet_priors_backward2(ptr(yf), yf.numel(), L, ptr(theta), ptr(g), ptr(u), ptr(gt2), ptr(ggp), ptr(ws),
                                           stream()))
    return ggp.reshape(grad_priors.shape), tuple(a.reshape(w.shape) for a, w in zip(unpack_params(gt2, S), weights))

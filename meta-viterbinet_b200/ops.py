"""Tensor-level wrappers around the C ABI (one function per entry point of include/mvn_b200.h)."""
import torch

from . import _lib
from ._lib import OUT_BITS, OUT_F32, check, dev_f32, load, ptr, stream


FUSED_VARIANTS = {'auto': 0, 'fma_smem': 1, 'fma_const_320': 2, 'tcgen05': 3, 'fma': 4}     # MVN_VARIANT_*
DECISIONS = {'reference': 0, 'mlse': 1, 'mlse_terminated': 2}                               # MVN_DECIDE_*
_default_variant = 'auto'


def set_fused_variant(name: str = 'auto') -> str:
    """Default implementation of the fused ViterbiNet kernel for calls that do not pass `variant` (tuning / testing):
    'auto' (library default: layers 2-3 on the tcgen05 tensor cores with an fp16 two-piece split, every memory
    length), 'tcgen05' (the same, explicitly), 'fma' (FP32 FMA pipe, constant-bank weights), 'fma_smem' (FP32 FMA
    pipe, weights in shared memory), 'fma_const_320'.  This is a Python-side default only: the C library itself keeps
    no such switch, the variant travels with every call (mvn_vnet_decode_ex).  Returns the previous selection."""
    global _default_variant
    if name not in FUSED_VARIANTS:
        raise ValueError(f'unknown fused-kernel variant {name!r}')
    old, _default_variant = _default_variant, name
    return old


def tc_timeout_status() -> bool:
    """True if the tensor-core kernel's pipeline watchdog fired on the current device since the last reset (check it
    after synchronising when the output of an asynchronous vnet_decode matters)."""
    return bool(load().mvn_tc_timeout_status())


def reset_tc_timeout():
    check(load().mvn_reset_tc_timeout())


def _mem_len(n_states: int) -> int:
    L = int(n_states).bit_length() - 1
    if n_states < 2 or (1 << L) != n_states:
        raise ValueError(f'n_states must be a power of two >= 2, got {n_states}')
    return L


def _new_out(B, T, out_format, device):
    if out_format == OUT_F32:
        return torch.empty((B, T), dtype=torch.float32, device=device)
    return torch.empty((B, (T + 31) // 32), dtype=torch.int32, device=device)


def new_counters(device=None):
    """[bit errors, frame errors, bits, frames] as int64 (the kernels treat them as uint64)."""
    return torch.zeros(4, dtype=torch.int64, device=device or _lib.require_cuda())


def acs_block(in_prob, llrs, n_states):
    """One ACS stage (trellis_utils.py:16-30): returns (values [B,S] fp32, indices [B,S] int64)."""
    L = _mem_len(n_states)
    in_prob = dev_f32(in_prob)
    llrs = dev_f32(llrs)
    B = in_prob.shape[0]
    if in_prob.dim() != 2 or in_prob.shape[1] != n_states:
        raise ValueError('in_prob must be [batch, n_states]')
    if llrs.dim() != 2 or llrs.shape[0] != B or llrs.shape[1] not in (1, n_states):
        raise ValueError('llrs must be [batch, n_states] or [batch, 1]')
    out = torch.empty_like(in_prob)
    idx = torch.empty((B, n_states), dtype=torch.int64, device=in_prob.device)
    check(load().mvn_acs_block(ptr(in_prob), ptr(llrs), llrs.shape[1], B, L, ptr(out), ptr(idx), stream()))
    return out, idx


ACS_LAYOUTS = {'auto': 0, 'lane_per_frame': 1, 'states_on_lanes': 2}                          # MVN_LAYOUT_*


def acs_decode(cost, n_stages=None, out_format=OUT_F32, return_final_pm=False, return_survivors=False, layout='auto'):
    """Stage loop on cost [B,T,S] (the a3 loop).  Returns decoded (+ final_pm, + survivors).  layout: thread layout of
    the kernel (ACS_LAYOUTS; 'auto' = lane per frame up to 64 states, states on lanes at 128 / 256)."""
    cost = dev_f32(cost)
    B, T, S = cost.shape
    L = _mem_len(S)
    n = T if n_stages is None else int(n_stages)
    dec = _new_out(B, T, out_format, cost.device)
    pm = torch.empty((B, S), dtype=torch.float32, device=cost.device) if return_final_pm else None
    sw = max(1, S // 64)
    surv = torch.zeros((B, n, sw), dtype=torch.int32, device=cost.device) if return_survivors else None
    check(load().mvn_acs_decode_ex(ptr(cost), B, T, L, n, out_format, ptr(dec), ptr(pm), ptr(surv), ACS_LAYOUTS[layout], stream()))
    res = (dec,)
    if return_final_pm:
        res += (pm,)
    if return_survivors:
        res += (surv,)
    return res if len(res) > 1 else dec


def va_decode(y, state_priors, n_stages=None, out_format=OUT_F32, target=None, pilot_period=0, counters=None,
              want_decoded=True, decision='reference'):
    """Fused full-CSI Viterbi.  state_priors [n_h, S] fp32 (row per tap block).  decision: 'reference' (the reference's
    running-argmin rule), 'mlse' / 'mlse_terminated' (in-kernel survivor traceback)."""
    y = dev_f32(y)
    sp = dev_f32(state_priors)
    B, T = y.shape
    n_h, S = sp.shape
    L = _mem_len(S)
    n = T if n_stages is None else int(n_stages)
    dec = _new_out(B, T, out_format, y.device) if want_decoded else None
    tgt, tT = None, 0
    if target is not None:
        tgt = dev_f32(target)
        tT = tgt.shape[1]
        if counters is None:
            raise ValueError('target given without counters')
    check(load().mvn_va_decode_ex(ptr(y), B, T, L, n, ptr(sp), n_h, out_format, ptr(dec), ptr(tgt), tT, pilot_period,
                                  ptr(counters), DECISIONS[decision], stream()))
    return dec


def _weights(weights):
    ws = [dev_f32(w) for w in weights]
    if len(ws) != 6:
        raise ValueError('expected [W1,b1,W2,b2,W3,b3]')
    S = ws[5].numel()
    shapes = [(100, 1), (100,), (50, 100), (50,), (S, 50), (S,)]
    for w, s in zip(ws, shapes):
        if tuple(w.shape) != s:
            raise ValueError(f'weight shape {tuple(w.shape)} != {s}')
    return ws, S


def vnet_priors(y, weights):
    """Priors of the ViterbiNet MLP: y [B,T] -> [B,T,S] (no autograd)."""
    y = dev_f32(y)
    ws, S = _weights(weights)
    L = _mem_len(S)
    out = torch.empty(tuple(y.shape) + (S,), dtype=torch.float32, device=y.device)
    check(load().mvn_vnet_priors(ptr(y), y.numel(), L, *[ptr(w) for w in ws], ptr(out), stream()))
    return out


def vnet_decode(y, weights, n_stages=None, out_format=OUT_F32, return_priors=False, target=None, pilot_period=0,
                counters=None, want_decoded=True, variant=None, decision='reference'):
    """Fused priors MLP + stage loop + decision.  variant: one of FUSED_VARIANTS (None = set_fused_variant's default);
    decision as in va_decode."""
    y = dev_f32(y)
    ws, S = _weights(weights)
    L = _mem_len(S)
    B, T = y.shape
    n = T if n_stages is None else int(n_stages)
    dec = _new_out(B, T, out_format, y.device) if want_decoded else None
    pri = torch.zeros((B, T, S), dtype=torch.float32, device=y.device) if return_priors else None
    tgt, tT = None, 0
    if target is not None:
        tgt = dev_f32(target)
        tT = tgt.shape[1]
        if counters is None:
            raise ValueError('target given without counters')
    var = variant or _default_variant
    if decision != 'reference' and (L > 6 or var not in ('auto', 'tcgen05')):
        # the in-kernel traceback lives in the tensor-core kernel (memory_length <= 6); 128 / 256 states and the FP32-FMA
        # variants take three launches: priors -> stage loop with exported survivors -> traceback kernel
        pri = vnet_priors(y, ws)
        dec = mlse_decode(-pri, n, terminated=(decision == 'mlse_terminated'), out_format=out_format)
        if tgt is not None:
            words = unpack_bits(dec, T) if out_format == OUT_BITS else dec
            error_counts(words[:, :tT], tgt, pilot_period=pilot_period, counters=counters, want_rows=False)
        return (dec, pri) if return_priors else dec
    check(load().mvn_vnet_decode_ex(ptr(y), B, T, L, n, *[ptr(w) for w in ws], out_format, ptr(dec), ptr(pri), ptr(tgt),
                                    tT, pilot_period, ptr(counters), FUSED_VARIANTS[var], DECISIONS[decision], stream()))
    return (dec, pri) if return_priors else dec


def calculate_states(memory_length, transmitted_words):
    tx = dev_f32(transmitted_words)
    B, T = tx.shape
    out = torch.empty(B * T, dtype=torch.int64, device=tx.device)
    check(load().mvn_calculate_states(ptr(tx), B, T, int(memory_length), ptr(out), stream()))
    return out


def error_counts(prediction, target, pilot_period=0, counters=None, want_rows=True):
    """Exact integer counts.  Returns (counters int64[4], row_errors uint8 [B] or None)."""
    p = dev_f32(prediction)
    t = dev_f32(target)
    if p.dim() != 2 or t.dim() != 2 or p.shape[0] != t.shape[0]:
        raise ValueError('prediction/target must be [rows, cols] with equal rows')
    if p.shape[1] != t.shape[1]:
        raise RuntimeError(f'The size of tensor a ({p.shape[1]}) must match the size of tensor b ({t.shape[1]})')
    B, T = p.shape
    if counters is None:
        counters = new_counters(p.device)
    rows = torch.empty(B, dtype=torch.uint8, device=p.device) if want_rows else None
    check(load().mvn_error_counts(ptr(p), p.stride(0), ptr(t), t.stride(0), B, T, pilot_period, ptr(counters),
                                  ptr(rows), stream()))
    return counters, rows


def unpack_bits(words, T):
    """[B, ceil(T/32)] int32 words -> [B,T] fp32 0/1 (host-side convenience for OUT_BITS)."""
    shifts = torch.arange(32, device=words.device, dtype=torch.int32)
    bits = (words.unsqueeze(-1) >> shifts) & 1
    return bits.reshape(words.shape[0], -1)[:, :T].float()


# ------------------------------------------------------------------ SURVEY.md §8(f) "next" rows
def _bind_next():
    return load()


def channel_transmit(bits, taps, snr_db, noise=None, seed=0):
    """ISI-AWGN channel on the device (channel_dataset.py:71,87-95 + channel.py:12-35): bits [B,T] 0/1 ->
    y [B,T] fp32.  taps [n_h, L] float64 (n_h = 1 or B, or any divisor pattern b mod n_h); noise [B,T] float64
    standard-normal samples for exact parity with a reference run, else Philox on the device with `seed`."""
    lib = _bind_next()
    bits = dev_f32(bits)
    B, T = bits.shape
    taps = torch.as_tensor(taps, dtype=torch.float64).to(bits.device).contiguous()
    if taps.dim() == 1:
        taps = taps.unsqueeze(0)
    n_h, L = taps.shape
    nz = None
    if noise is not None:
        nz = torch.as_tensor(noise, dtype=torch.float64).to(bits.device).contiguous()
        if tuple(nz.shape) != (B, T):
            raise ValueError('noise must be [B,T]')
    y = torch.empty((B, T), dtype=torch.float32, device=bits.device)
    check(lib.mvn_channel_transmit(ptr(bits), B, T, L, ptr(taps), n_h, float(snr_db), ptr(nz), int(seed), ptr(y), stream()))
    return y


def rs_decode(detected, nsym, return_status=False):
    """Reed-Solomon decoding of a batch of detected words (ecc/rs_main.py:21-37 per row, trainer.py:234-236):
    detected [B, 8*n_bytes] 0/1 -> [B, 8*(n_bytes-nsym)] fp32 0/1.  status (int32 [B]): 0 clean, 1 corrected,
    2 too many errors (returned unchanged), 3 locator with missing roots (partial correction, as the reference)."""
    lib = _bind_next()
    x = dev_f32(detected)
    if x.dim() != 2 or x.shape[1] % 8:
        raise ValueError('detected must be [B, 8*n_bytes]')
    B, n_bytes = x.shape[0], x.shape[1] // 8
    k = n_bytes - int(nsym)
    out = torch.empty((B, 8 * max(k, 0)), dtype=torch.float32, device=x.device)
    st = torch.empty(B, dtype=torch.int32, device=x.device) if return_status else None
    check(lib.mvn_rs_decode(ptr(x), B, x.shape[1], n_bytes, int(nsym), ptr(out), out.shape[1], ptr(st), stream()))
    return (out, st) if return_status else out


def rs_encode(words, nsym):
    """Systematic Reed-Solomon encoding (ecc/rs_main.py:9-18 per row): words [B, 8*k_bytes] 0/1 ->
    [B, 8*(k_bytes+nsym)] fp32 0/1, message bits followed by the parity bits."""
    lib = _bind_next()
    x = dev_f32(words)
    if x.dim() != 2 or x.shape[1] % 8:
        raise ValueError('words must be [B, 8*k_bytes]')
    B, k = x.shape[0], x.shape[1] // 8
    out = torch.empty((B, 8 * (k + int(nsym))), dtype=torch.float32, device=x.device)
    check(lib.mvn_rs_encode(ptr(x), B, x.shape[1], k, int(nsym), ptr(out), out.shape[1], stream()))
    return out


def mlse_decode(cost, n_stages=None, terminated=False, out_format=OUT_F32):
    """True-MLSE decoding of cost [B,T,S] by survivor traceback (the mode SURVEY.md §8f ranks 3rd; the reference's own
    rule is acs_decode).  terminated=True starts the traceback from state 0 (zero-padded words), else from the best
    final state."""
    lib = _bind_next()
    cost = dev_f32(cost)
    B, T, S = cost.shape
    L = _mem_len(S)
    n = T if n_stages is None else int(n_stages)
    if n == 0 or B == 0:                       # no stage: nothing to trace back, every column stays 0
        return torch.zeros_like(_new_out(B, T, out_format, cost.device))
    _, pm, surv = acs_decode(cost, n, return_final_pm=True, return_survivors=True)
    dec = _new_out(B, T, out_format, cost.device)
    check(lib.mvn_traceback(ptr(surv), ptr(pm), B, T, n, L, 0 if terminated else -1, out_format, ptr(dec), stream()))
    return dec


def va_mlse_decode(y, state_priors, n_stages=None, terminated=True, out_format=OUT_F32):
    """Full-CSI Viterbi with traceback: branch metrics as in va_decode, decisions by true MLSE.  One fused launch:
    the survivor bits live in shared memory and the traceback runs in the kernel (mvn_va_decode_ex)."""
    return va_decode(y, state_priors, n_stages, out_format, decision='mlse_terminated' if terminated else 'mlse')


def random_bits(B, T, seed=0, device=None):
    """Bernoulli(1/2) words [B,T] fp32 on the device (Philox; the role of word_rand_gen.randint, channel_dataset.py:67)."""
    out = torch.empty((B, T), dtype=torch.float32, device=device or _lib.require_cuda())
    check(load().mvn_random_bits(ptr(out), B, T, int(seed), stream()))
    return out

"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's detection hot path.

This file is the *checker*: a plain CPU restatement (numpy, IEEE fp32 with the same
operation order as the reference's torch expressions) of every row of SURVEY.md §8(a).
It is pinned against the live reference (``/root/reference``, imported by
``tests/golden/make_golden.py`` in the build container) through the committed fixtures in
``tests/golden/*.npz`` — see ``tests/test_oracle_golden.py``.  Parity status: PINNED (the
reference ships no tests or golden vectors of its own, SURVEY.md §4, so the pins are
outputs of the reference itself run on seeded inputs; the generating script is committed).

All citations are relative to the reference checkout root.
Nothing in the shipped package imports this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
HIDDEN1 = 100  # python_code/detectors/VNET/vnet_detector.py:7
HIDDEN2 = 50   # python_code/detectors/VNET/vnet_detector.py:8
LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))  # python_code/detectors/VA/va_detector.py:68


# --------------------------------------------------------------------------------------
# a1 / a2 / a3 — trellis primitives and the stage loop
# --------------------------------------------------------------------------------------
def transition_table(n_states: int) -> np.ndarray:
    """a1. python_code/utils/trellis_utils.py:7-13: two aranges back to back, viewed [S,2],
    i.e. row j = [(2j) mod S, (2j+1) mod S]."""
    j = np.arange(n_states)
    return np.stack([(2 * j) % n_states, (2 * j + 1) % n_states], axis=1)


def acs_stage(pm: np.ndarray, cost: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """a2. python_code/utils/trellis_utils.py:16-30.  tmp = pm + cost (per SOURCE state, fp32);
    out[j] = min over the two predecessors table[j]; survivor = argmin in {0,1}, tie -> 0
    (torch.min(dim) returns the first minimum)."""
    S = pm.shape[1]
    tab = transition_table(S)
    tmp = (pm + cost).astype(F32)
    cand = tmp[:, tab]                       # [B,S,2]
    surv = (cand[:, :, 1] < cand[:, :, 0]).astype(np.int64)
    return np.minimum(cand[:, :, 0], cand[:, :, 1]), surv


def acs_decode(cost: np.ndarray, n_stages: Optional[int] = None,
               return_survivors: bool = False):
    """a3. The loop shared by the three detectors (va_detector.py:83-98, vnet_detector.py:46-61,
    meta_vnet_detector.py:25-45): pm=0; per stage FIRST decide bit = argmin_s(pm) mod 2
    (lowest index on ties), THEN run the ACS with this stage's cost.  Columns >= n_stages of
    the output stay 0 (decoded_word is zeros(y.shape), loop runs transmission_length times).
    Returns (decoded fp32 [B,T], final pm fp32 [B,S]) (+ survivors int8 [B,n_stages,S])."""
    cost = np.asarray(cost, dtype=F32)
    B, T, S = cost.shape
    n = T if n_stages is None else n_stages
    pm = np.zeros((B, S), dtype=F32)
    dec = np.zeros((B, T), dtype=F32)
    survs = np.zeros((B, n, S), dtype=np.int8) if return_survivors else None
    for t in range(n):
        dec[:, t] = (np.argmin(pm, axis=1) % 2).astype(F32)
        pm, sv = acs_stage(pm, cost[:, t])
        if return_survivors:
            survs[:, t] = sv
    if return_survivors:
        return dec, pm, survs
    return dec, pm


def mlse_decode(cost: np.ndarray, n_stages: Optional[int] = None, start_state: int = -1):
    """SURVEY.md §8(f)3 — true MLSE over the reference trellis by survivor traceback (the indices
    trellis_utils.py:30 returns).  State(t) = sum_i b[t+i] 2^i, so the predecessor of j is (2j+sigma) mod S and the
    survivor choice sigma is the transmitted bit b[t].  start_state < 0: best final state (lowest index on ties).
    Returns (decoded fp32 [B,T], states int [B,n+1] along the surviving path, final pm)."""
    cost = np.asarray(cost, dtype=F32)
    B, T, S = cost.shape
    n = T if n_stages is None else n_stages
    _, pm, surv = acs_decode(cost, n, return_survivors=True)
    dec = np.zeros((B, T), dtype=F32)
    states = np.zeros((B, n + 1), dtype=np.int64)
    j = np.argmin(pm, axis=1) if start_state < 0 else np.full(B, start_state)
    states[:, n] = j
    rows = np.arange(B)
    for t in range(n - 1, -1, -1):
        sigma = surv[rows, t, j].astype(np.int64)
        j = (2 * j + sigma) % S
        dec[:, t] = sigma
        states[:, t] = j
    return dec, states, pm


def path_cost(cost: np.ndarray, states: np.ndarray) -> np.ndarray:
    """fp32 metric of a state sequence accumulated stage by stage like the ACS recursion does."""
    B, n = states.shape[0], states.shape[1] - 1
    acc = np.zeros(B, dtype=F32)
    for t in range(n):
        acc = (acc + cost[np.arange(B), t, states[:, t]]).astype(F32)
    return acc


# --------------------------------------------------------------------------------------
# a4 / a5 — classical VA with full CSI
# --------------------------------------------------------------------------------------
def va_state_priors(h: np.ndarray, memory_length: int) -> np.ndarray:
    """a4. va_detector.py:42-50.  Noiseless channel output per state: the last L columns of the
    8-bit big-endian expansion of s (MSB first), BPSK 1-2b (modulator.py:12), float64 dot with
    h^T, cast to fp32.  h: [n_h, L] float64  ->  [S, n_h] fp32."""
    S = 2 ** memory_length
    s = np.arange(S).astype(np.uint8).reshape(-1, 1)
    bits = np.unpackbits(s, axis=1).astype(int)[:, -memory_length:]
    sym = 1 - 2 * bits
    return np.dot(sym, np.asarray(h, dtype=np.float64).T).astype(F32)


def va_cost(y: np.ndarray, state_priors: np.ndarray) -> np.ndarray:
    """a5. va_detector.py:62-68.  Row r uses tap block r mod n_h (table tiled B//n_h times);
    d = y - sp; cost = d*d/2 - fp32(ln sqrt(2 pi)), every op rounded to fp32 separately."""
    y = np.asarray(y, dtype=F32)
    B = y.shape[0]
    n_h = state_priors.shape[1]
    if B % n_h != 0:
        raise RuntimeError("batch must be a multiple of the number of tap blocks")
    sp = np.tile(state_priors.T.astype(F32), (B // n_h, 1))       # [B,S]
    d = (y[:, :, None] - sp[:, None, :]).astype(F32)
    sq = (d * d).astype(F32)
    return (sq / F32(2) - F32(LOG_SQRT_2PI)).astype(F32)


def estimate_channel(memory_length: int, gamma: float, channel_coefficients: str = 'time_decay',
                     fading: bool = False, index: int = 0, fading_taps_type: int = 1,
                     cost2100_taps: Optional[np.ndarray] = None) -> np.ndarray:
    """python_code/channel/channel_estimation.py:11-49 without the noisy-estimate branch
    (noisy_est_var draws from the *global* numpy RNG, :35-36, which no seed controls).
    cost2100_taps: [300, L] array standing in for the four .mat files (:27-30)."""
    if channel_coefficients == 'time_decay':
        h = np.reshape(np.exp(-gamma * np.arange(memory_length)), [1, memory_length])
    elif channel_coefficients == 'cost2100':
        h = np.reshape(np.asarray(cost2100_taps, dtype=np.float64)[index], [1, memory_length]).copy()
    else:
        raise ValueError('No such channel_coefficients value!!!')
    if fading and channel_coefficients == 'time_decay':
        if fading_taps_type == 1:
            periods = np.array([51, 39, 33, 21])
            h = h * (0.8 + 0.2 * np.cos(2 * np.pi * index / periods)).reshape(1, memory_length)
        elif fading_taps_type == 2:
            periods = 5 * np.array([51, 39, 33, 21])
            periods = np.maximum(periods - 1.5 * index, 10 * np.ones(4)) - 1e-5
            h = h * (0.8 + 0.2 * np.cos(np.pi * index / periods)).reshape(1, memory_length)
        else:
            raise ValueError("No such fading tap type!!!")
    return h


def va_decode(y: np.ndarray, h: np.ndarray, memory_length: int,
              n_stages: Optional[int] = None) -> np.ndarray:
    """VADetector.forward(y,'val') given the stacked taps h [n_h, L] (va_detector.py:73-98)."""
    dec, _ = acs_decode(va_cost(y, va_state_priors(h, memory_length)), n_stages)
    return dec


# --------------------------------------------------------------------------------------
# a6 / a7 — ViterbiNet priors network
# --------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1 / (1 + np.exp(-x))


def vnet_priors(y: np.ndarray, weights: Sequence[np.ndarray], dtype=F32) -> np.ndarray:
    """a6/a7. vnet_detector.py:27-33,49 and meta_vnet_detector.py:27-33:
    p = W3 relu(W2 sigmoid(W1 y + b1) + b2) + b3 per received sample.
    weights = [W1 (100,1), b1 (100), W2 (50,100), b2 (50), W3 (S,50), b3 (S)].
    dtype=float64 gives the 'exact' priors used to state tolerances."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(w, dtype=dtype) for w in weights]
    B, T = y.shape
    x = np.asarray(y, dtype=dtype).reshape(-1, 1)
    h1 = _sigmoid(x @ W1.T + b1).astype(dtype)
    h2 = np.maximum(h1 @ W2.T + b2, 0).astype(dtype)
    p = (h2 @ W3.T + b3).astype(dtype)
    return p.reshape(B, T, W3.shape[0])


def vnet_decode_from_priors(priors: np.ndarray, n_stages: Optional[int] = None):
    """vnet_detector.py:51-61: the a3 loop on cost = -priors."""
    return acs_decode(-np.asarray(priors, dtype=F32), n_stages)


def vnet_decode(y: np.ndarray, weights: Sequence[np.ndarray], n_stages: Optional[int] = None):
    return vnet_decode_from_priors(vnet_priors(y, weights), n_stages)[0]


# --------------------------------------------------------------------------------------
# a8 / a9 — labels and loss
# --------------------------------------------------------------------------------------
def calculate_states(memory_length: int, tx: np.ndarray) -> np.ndarray:
    """a8. trellis_utils.py:33-46: state[b,t] = sum_{i<L} tx[b,t+i] 2^i, tx zero past the end,
    flattened row-major to [B*T] int64."""
    tx = np.asarray(tx)
    B, T = tx.shape
    padded = np.concatenate([tx, np.zeros((B, memory_length), dtype=tx.dtype)], axis=1)
    st = np.zeros((B, T), dtype=np.int64)
    for i in range(memory_length):
        st += padded[:, i:i + T].astype(np.int64) << i
    return st.reshape(-1)


def cross_entropy(priors: np.ndarray, labels: np.ndarray, dtype=np.float64) -> float:
    """a9. torch CrossEntropyLoss(), mean reduction (trainer.py:180-181):
    mean_n(logsumexp(p_n) - p_n[label_n])."""
    p = np.asarray(priors, dtype=dtype).reshape(-1, priors.shape[-1])
    m = p.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(p - m).sum(axis=1))
    return float(np.mean(lse - p[np.arange(p.shape[0]), labels]))


# --------------------------------------------------------------------------------------
# a10 / a11 — (meta-)training steps of the priors net, hand-derived
# --------------------------------------------------------------------------------------
def _forward_cache(y, labels, w, dtype):
    W1, b1, W2, b2, W3, b3 = [np.asarray(a, dtype=dtype) for a in w]
    x = np.asarray(y, dtype=dtype).reshape(-1, 1)
    a1 = x @ W1.T + b1
    h1 = _sigmoid(a1)
    a2 = h1 @ W2.T + b2
    h2 = np.maximum(a2, 0)
    z = h2 @ W3.T + b3
    zm = z - z.max(axis=1, keepdims=True)
    e = np.exp(zm)
    p = e / e.sum(axis=1, keepdims=True)
    N = x.shape[0]
    loss = float(np.mean(-np.log(p[np.arange(N), labels])))
    return dict(x=x, a1=a1, h1=h1, a2=a2, h2=h2, z=z, p=p, N=N, loss=loss)


def loss_and_grads(y, labels, w, dtype=np.float64):
    """CE loss of the priors net on flattened samples y with state labels, and its gradient
    w.r.t. [W1,b1,W2,b2,W3,b3] (what loss.backward() gives in trainer.py:503)."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(a, dtype=dtype) for a in w]
    c = _forward_cache(y, labels, w, dtype)
    N = c['N']
    dz = c['p'].copy()
    dz[np.arange(N), labels] -= 1
    dz /= N
    gW3 = dz.T @ c['h2']
    gb3 = dz.sum(0)
    dh2 = dz @ W3
    da2 = dh2 * (c['a2'] > 0)
    gW2 = da2.T @ c['h1']
    gb2 = da2.sum(0)
    dh1 = da2 @ W2
    s1 = c['h1'] * (1 - c['h1'])
    da1 = dh1 * s1
    gW1 = da1.T @ c['x']
    gb1 = da1.sum(0)
    c.update(dz=dz, da2=da2, dh1=dh1, s1=s1, da1=da1)
    return c['loss'], [gW1, gb1, gW2, gb2, gW3, gb3], c


def hessian_vector_product(y, labels, w, v, dtype=np.float64):
    """H(w) v for the same loss by forward-over-reverse differentiation (used for the
    second-order term of MAML, trainer.py:437 create_graph=True ... :444)."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(a, dtype=dtype) for a in w]
    V1, c1, V2, c2, V3, c3 = [np.asarray(a, dtype=dtype) for a in v]
    _, _, c = loss_and_grads(y, labels, w, dtype)
    N = c['N']
    x, h1, h2, p = c['x'], c['h1'], c['h2'], c['p']
    mask = (c['a2'] > 0)
    s1 = c['s1']
    Ra1 = x @ V1.T + c1
    Rh1 = s1 * Ra1
    Ra2 = h1 @ V2.T + Rh1 @ W2.T + c2
    Rh2 = mask * Ra2
    Rz = h2 @ V3.T + Rh2 @ W3.T + c3
    Rp = p * (Rz - (p * Rz).sum(axis=1, keepdims=True))
    Rdz = Rp / N
    hW3 = Rdz.T @ h2 + c['dz'].T @ Rh2
    hb3 = Rdz.sum(0)
    Rdh2 = c['dz'] @ V3 + Rdz @ W3
    Rda2 = Rdh2 * mask
    hW2 = Rda2.T @ h1 + c['da2'].T @ Rh1
    hb2 = Rda2.sum(0)
    Rdh1 = c['da2'] @ V2 + Rda2 @ W2
    s2 = s1 * (1 - 2 * h1)                       # sigma''
    Rda1 = Rdh1 * s1 + c['dh1'] * s2 * Ra1
    hW1 = Rda1.T @ x
    hb1 = Rda1.sum(0)
    return [hW1, hb1, hW2, hb2, hW3, hb3]


def adam_update(w, g, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, dtype=np.float64):
    """torch.optim.Adam with defaults (trainer.py:167-169).  state = dict(step, m, v) or None."""
    if state is None:
        state = dict(step=0, m=[np.zeros_like(np.asarray(a, dtype=dtype)) for a in w],
                     v=[np.zeros_like(np.asarray(a, dtype=dtype)) for a in w])
    b1, b2 = betas
    step = state['step'] + 1
    new_w, new_m, new_v = [], [], []
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for p, gi, m, v in zip(w, g, state['m'], state['v']):
        p = np.asarray(p, dtype=dtype)
        gi = np.asarray(gi, dtype=dtype).reshape(p.shape)
        m = b1 * m + (1 - b1) * gi
        v = b2 * v + (1 - b2) * gi * gi
        denom = np.sqrt(v) / math.sqrt(bc2) + eps
        new_w.append(p - (lr / bc1) * m / denom)
        new_m.append(m)
        new_v.append(v)
    return new_w, dict(step=step, m=new_m, v=new_v)


def train_step(y, tx, w, state, memory_length, lr=1e-3, dtype=np.float64):
    """a11. One run_train_loop iteration (trainer.py:492-505) with the Meta-ViterbiNet loss over
    ALL symbols (metavnet_trainer.py:41-50).  y, tx: [W,T].  Returns (loss, new_w, new_state)."""
    labels = calculate_states(memory_length, tx)
    loss, g, _ = loss_and_grads(np.asarray(y).reshape(-1), labels, w, dtype)
    new_w, state = adam_update(w, g, state, lr=lr, dtype=dtype)
    return loss, new_w, state


def maml_step(y_s, tx_s, y_q, tx_q, w, state, memory_length, meta_lr=0.1, lr=1e-3,
              second_order=True, dtype=np.float64):
    """a10. meta_train_loop (trainer.py:425-453): inner SGD step of size meta_lr on the support
    word(s), query loss at the adapted weights, gradient w.r.t. the ORIGINAL weights
    (g_q - meta_lr * H_s g_q when MAML=True, g_q for first order), Adam step.
    Returns (loss_query, meta_grad list, new_w, new_state)."""
    ls = calculate_states(memory_length, tx_s)
    lq = calculate_states(memory_length, tx_q)
    w = [np.asarray(a, dtype=dtype) for a in w]
    _, g_s, _ = loss_and_grads(np.asarray(y_s).reshape(-1), ls, w, dtype)
    w_fast = [p - meta_lr * g.reshape(p.shape) for p, g in zip(w, g_s)]
    loss_q, g_q, _ = loss_and_grads(np.asarray(y_q).reshape(-1), lq, w_fast, dtype)
    g_q = [g.reshape(p.shape) for g, p in zip(g_q, w)]
    if second_order:
        hv = hessian_vector_product(np.asarray(y_s).reshape(-1), ls, w, g_q, dtype)
        meta_g = [g - meta_lr * h.reshape(g.shape) for g, h in zip(g_q, hv)]
    else:
        meta_g = g_q
    new_w, state = adam_update(w, meta_g, state, lr=lr, dtype=dtype)
    return loss_q, meta_g, new_w, state


# --------------------------------------------------------------------------------------
# a12 — BER / FER
# --------------------------------------------------------------------------------------
def error_counts(prediction: np.ndarray, target: np.ndarray) -> Tuple[int, int, int, int, np.ndarray]:
    """Exact integer form of metrics.py:7-17: (bit errors, frame errors, n_bits, n_frames,
    indices of frames with any error).  .long() truncates toward zero."""
    p = np.trunc(np.asarray(prediction)).astype(np.int64)
    t = np.trunc(np.asarray(target)).astype(np.int64)
    neq = (p != t)
    rowsum = np.abs(p - t).sum(axis=1)
    bad = np.nonzero(rowsum)[0]
    return int(neq.sum()), int(bad.size), int(p.size), int(p.shape[0]), bad


def calculate_error_rates(prediction: np.ndarray, target: np.ndarray) -> Tuple[float, float, np.ndarray]:
    """a12. metrics.py:7-17: accuracies are fp32 means (count/N in fp32), then 1-x in double,
    clamped at 0.  Exact while N < 2^24."""
    be, fe, nb, nf, bad = error_counts(prediction, target)
    bits_acc = float(F32(nb - be) / F32(nb))
    frames_acc = float(F32(nf - fe) / F32(nf))
    return max([1 - bits_acc, 0.0]), max([1 - frames_acc, 0.0]), bad


# --------------------------------------------------------------------------------------
# Appendix A — channel, for synthetic inputs (NOT on the hot path)
# --------------------------------------------------------------------------------------
def isi_awgn(bits: np.ndarray, h: np.ndarray, snr_db: float, memory_length: int,
             rng: np.random.RandomState) -> np.ndarray:
    """channel_dataset.py:71,87-95 + channel.py:23-33: pad L zero bits, BPSK, y[t] =
    sum_i h[L-1-i] s[t+i] + 10^(-snr/20) n[t].  bits [B,T] in {0,1}; h [1,L] or [B,L]."""
    B, T = bits.shape
    L = memory_length
    c = np.concatenate([bits, np.zeros((B, L))], axis=1)
    s = 1 - 2 * c
    h = np.broadcast_to(np.asarray(h, dtype=np.float64), (B, L))
    conv = np.zeros((B, T))
    for i in range(L):
        conv += h[:, L - 1 - i:L - i] * s[:, i:i + T]
    w = (10 ** (snr_db / 10)) ** (-0.5) * rng.normal(0, 1, (B, T))
    return conv + w

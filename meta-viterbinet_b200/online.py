"""Word-by-word online evaluation, batched over independent runs.

The reference's ``Trainer.eval_by_word`` (trainers/trainer.py:267-354) walks through a sequence of coded words:
detect -> Reed-Solomon decode -> SER -> re-encode -> (if the SER is below a threshold) self-supervised online training
on the word just seen (metavnet_trainer.py:52-64 / vnet_trainer.py:49-59).  Every step depends on the weights the
previous one left behind, so one run is sequential; R runs (SNR points, channel realisations, seeds) are
independent.  ``eval_by_word`` below advances R runs in lock step with everything on the device: one batched
detection launch (per-run weights), one RS decode / encode launch, and one launch per training iteration for all runs;
the SER gate is a mask, so there is no host round trip per block.

Covered: pilots (``data_mask``), the SER gate, the label rule (detected word when 0 < SER <= threshold, re-encoded
word otherwise), training restarted from the saved weights (MetaViterbiNet flavour) or continued (ViterbiNet flavour),
Adam state carried across blocks as the reference's single optimizer does, and the periodic online META-training on the
replay buffer (``online_meta``, trainer.py:331-343): per-run buffers of the gated words (growing, or sliding when an
initial buffer is given), random query indices per run, support = the ``window_size`` words before the query with the
reference's wrap-around, one batched (FO-)MAML step per drawn index for all runs that have one, weights saved
afterwards.  With ``online_meta`` the gate decisions come to the host once per block (the buffers are ragged).
"""
import torch

from . import ops


def _default_draw(run, high, count):
    """torch.unique(torch.randint(low=0, high=high, size=[count])) — the reference's draw (trainer.py:336-338),
    from torch's global CPU generator, so a single run replays the reference's random stream."""
    return torch.unique(torch.randint(low=0, high=int(high), size=[int(count)])).tolist()


class _ReplayBuffers:
    """Per-run buffers of (received word, label word) rows on the device; lengths live on the host."""

    def __init__(self, R, T, capacity, device, init=None):
        n0 = 0 if init is None else int(init[0].shape[1])
        self.rx = torch.zeros((R, capacity + n0, T), dtype=torch.float32, device=device)
        self.tx = torch.zeros_like(self.rx)
        self.start = [0] * R
        self.length = [n0] * R
        self.sliding = init is not None              # not buffer_empty: the oldest word leaves when one enters
        if init is not None:
            self.tx[:, :n0] = ops.dev_f32(init[0])
            self.rx[:, :n0] = ops.dev_f32(init[1])

    def append(self, runs, rx, tx):
        if not runs:
            return
        r = torch.tensor(runs, device=self.rx.device)
        pos = torch.tensor([self.start[i] + self.length[i] for i in runs], device=self.rx.device)
        self.rx[r, pos] = rx[r]
        self.tx[r, pos] = tx[r]
        for i in runs:
            if self.sliding:
                self.start[i] += 1
            else:
                self.length[i] += 1

    def gather(self, rel_index):
        """rel_index: [R, n] positions inside each run's buffer (python-style negatives already resolved)."""
        base = torch.tensor(self.start, device=self.rx.device).unsqueeze(1)
        idx = (base + rel_index).unsqueeze(2).expand(-1, -1, self.rx.shape[2])
        return self.rx.gather(1, idx), self.tx.gather(1, idx)


def _snapshot(trainer):
    return [t.clone() for t in (trainer.theta, trainer.adam_m, trainer.adam_v, trainer.adam_step)]


def _restore_where_not(trainer, keep, active):
    """runs outside `active` ([R] bool, device) get their state back"""
    a = active.unsqueeze(1)
    trainer.theta.copy_(torch.where(a, trainer.theta, keep[0]))
    trainer.adam_m.copy_(torch.where(a, trainer.adam_m, keep[1]))
    trainer.adam_v.copy_(torch.where(a, trainer.adam_v, keep[2]))
    trainer.adam_step.copy_(torch.where(active, trainer.adam_step, keep[3]))


def _meta_round(trainer, buffers, saved, cfg, on_meta_step):
    """trainer.py:331-343 for every run whose buffer holds more than two words."""
    R = trainer.R
    dev = trainer.theta.device
    eligible = [r for r in range(R) if buffers.length[r] > 2]
    if not eligible:
        return
    el = torch.zeros(R, dtype=torch.bool, device=dev)
    el[eligible] = True
    init = cfg['weights_init']
    if isinstance(init, str) and init == 'last_frame':      # meta_weights_init (trainer.py:356-366)
        source = saved
    elif torch.is_tensor(init):                              # 'meta_training': the stored meta-trained weights
        source = ops.dev_f32(init).reshape(R, -1)
    elif callable(init):                                     # 'random': fresh weights AND a fresh optimizer (trainer.py:357-359)
        source = trainer.theta.clone()
        source[eligible] = ops.dev_f32(init(eligible)).reshape(len(eligible), -1)
        z = el.unsqueeze(1)
        trainer.adam_m.copy_(torch.where(z, torch.zeros_like(trainer.adam_m), trainer.adam_m))
        trainer.adam_v.copy_(torch.where(z, torch.zeros_like(trainer.adam_v), trainer.adam_v))
        trainer.adam_step.copy_(torch.where(el, torch.zeros_like(trainer.adam_step), trainer.adam_step))
    else:
        raise ValueError("weights_init must be 'last_frame', a [R,P] tensor of meta-trained weights, or a callable "
                         "runs -> [len(runs), P] fresh weights ('random')")
    trainer.theta.copy_(torch.where(el.unsqueeze(1), source, trainer.theta))
    window = cfg['window_size']
    for _ in range(cfg['meta_train_iterations']):
        draws = {r: cfg['draw'](r, buffers.length[r] - 2, cfg['meta_j_num']) for r in eligible}
        for k in range(max(len(v) for v in draws.values())):
            active_runs = [r for r in eligible if k < len(draws[r])]
            s_idx = torch.zeros((R, window), dtype=torch.long)
            q_idx = torch.zeros((R, 1), dtype=torch.long)
            for r in active_runs:
                j, n = int(draws[r][k]), buffers.length[r]
                for d in range(window):                      # j_hat + arange(-window-1, -1) + 1, negative indices wrap
                    p = j - window + d
                    s_idx[r, d] = p if p >= 0 else n + p
                q_idx[r, 0] = j
            rx_s, tx_s = buffers.gather(s_idx.to(dev))
            rx_q, tx_q = buffers.gather(q_idx.to(dev))
            T = rx_s.shape[2]
            lab_s = ops.calculate_states(trainer.L, tx_s.reshape(R * window, T)).reshape(R, window * T).to(torch.int32)
            lab_q = ops.calculate_states(trainer.L, tx_q.reshape(R, T)).reshape(R, T).to(torch.int32)
            active = torch.zeros(R, dtype=torch.bool, device=dev)
            active[active_runs] = True
            keep = _snapshot(trainer)
            trainer.meta_step(rx_s.reshape(R, window * T), lab_s.contiguous(), rx_q.reshape(R, T), lab_q.contiguous(),
                              second_order=cfg['second_order'])
            _restore_where_not(trainer, keep, active)
            if on_meta_step is not None:
                on_meta_step(active)
    saved.copy_(torch.where(el.unsqueeze(1), trainer.theta, saved))   # copy_model(detector -> saved_detector)


def eval_by_word(trainer, info_bits, received, n_symbols, ser_thresh, data_mask=None, subframes_in_frame=None,
                 self_supervised=True, iterations=200, restart_from_saved=True, on_block=None,
                 online_meta=False, meta_subframes=None, meta_train_iterations=1, meta_j_num=1, window_size=1,
                 second_order=True, weights_init='last_frame', init_buffer=None, draw=None, on_meta_step=None):
    """trainer: BatchedVNetTrainer holding the R runs' weights (updated in place, like ``self.detector``).
    info_bits [R, N, 8k] transmitted information bits, received [R, N, T] channel outputs (T = 8 (k + n_symbols)).
    data_mask [N] bool: False marks pilot words (known at the receiver, SER not counted; trainer.py:99-102 — or give
    ``subframes_in_frame`` and every word with index % subframes_in_frame == 0 is a pilot).
    online_meta: every ``meta_subframes`` words (trainer.py:331) run ``meta_train_iterations`` rounds of (FO-)MAML steps on
    the replay buffer with ``meta_j_num`` random query indices each (``draw(run, high, count)`` -> list of indices;
    default: the reference's torch.unique(torch.randint(...))), support = ``window_size`` preceding words, starting
    from ``weights_init`` ('last_frame' = the saved weights, a [R,P] tensor = the stored meta-trained weights, or a callable
    runs -> fresh weights = the reference's 'random', which also restarts the optimizer);
    ``trainer.meta_lr`` is the inner step.  init_buffer = (label words [R,n0,T], received [R,n0,T]) starts with a
    filled, sliding buffer (``buffer_empty: False``).
    Returns ser_by_word [R, N] float64 (0 for pilots), as trainer.py:354 returns per run."""
    info = ops.dev_f32(info_bits)
    rx = ops.dev_f32(received)
    R, N, T = rx.shape
    if info.shape[:2] != (R, N) or info.shape[2] + 8 * int(n_symbols) != T:
        raise ValueError('info_bits must be [R, N, T - 8 * n_symbols]')
    if data_mask is None:
        per = int(subframes_in_frame) if subframes_in_frame else 0
        data_mask = [(c % per != 0) if per else True for c in range(N)]
    saved = trainer.theta.clone()                        # self.saved_detector (trainer.py:277)
    buffers = _ReplayBuffers(R, T, N, rx.device, init_buffer) if online_meta else None
    meta_cfg = dict(meta_train_iterations=int(meta_train_iterations), meta_j_num=int(meta_j_num), window_size=int(window_size),
                    second_order=bool(second_order), weights_init=weights_init, draw=draw or _default_draw)
    ser_by_word = torch.zeros((R, N), dtype=torch.float64, device=rx.device)
    n_info = info.shape[2]
    for c in range(N):
        y = rx[:, c].contiguous()
        detected = trainer.detect(y)                      # self.detector(received_word, 'val', ...) with each run's weights
        if bool(data_mask[c]):
            decoded = ops.rs_decode(detected, n_symbols)
            # calculate_error_rates (metrics.py:11-17): accuracy = fp32 mean of the equal bits, ser = max(1 - acc, 0) in double
            acc32 = (decoded == info[:, c]).float().sum(dim=1) / n_info
            ser = (1.0 - acc32.double()).clamp(min=0.0)
            encoded = ops.rs_encode(decoded, n_symbols)
            label = torch.where((ser > 0).unsqueeze(1), detected, encoded)    # trainer.py:322-324
            ser_by_word[:, c] = ser
        else:                                             # pilot: the receiver knows the word (trainer.py:310-316)
            label = ops.rs_encode(info[:, c].contiguous(), n_symbols)
            ser = torch.zeros(R, dtype=torch.float64, device=rx.device)
        gate = ser <= float(ser_thresh)
        if online_meta:
            buffers.append([r for r, ok in enumerate(gate.cpu().tolist()) if ok], rx[:, c], label)   # trainer.py:319-329
            if c % int(meta_subframes) == 0 and c >= int(meta_subframes):
                _meta_round(trainer, buffers, saved, meta_cfg, on_meta_step)
        if self_supervised:
            keep = _snapshot(trainer)
            if restart_from_saved:                        # metavnet_trainer.py:59
                trainer.theta.copy_(saved)
            for _ in range(int(iterations)):
                trainer.train_step(y, label)
            _restore_where_not(trainer, keep, gate)       # runs whose word did not pass the gate keep their state
        if on_block is not None:
            on_block(c, ser, gate)
    return ser_by_word

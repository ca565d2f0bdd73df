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
Adam state carried across blocks as the reference's single optimizer does.  Not covered: the periodic online
META-training on the replay buffer (``online_meta``), which draws random support/query pairs per run.
"""
import torch

from . import ops


def eval_by_word(trainer, info_bits, received, n_symbols, ser_thresh, data_mask=None, subframes_in_frame=None,
                 self_supervised=True, iterations=200, restart_from_saved=True, on_block=None):
    """trainer: BatchedVNetTrainer holding the R runs' weights (updated in place, like ``self.detector``).
    info_bits [R, N, 8k] transmitted information bits, received [R, N, T] channel outputs (T = 8 (k + n_symbols)).
    data_mask [N] bool: False marks pilot words (known at the receiver, SER not counted; trainer.py:99-102 — or give
    ``subframes_in_frame`` and every word with index % subframes_in_frame == 0 is a pilot).
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
    ser_by_word = torch.zeros((R, N), dtype=torch.float64, device=rx.device)
    n_info = info.shape[2]
    for c in range(N):
        y = rx[:, c].contiguous()
        detected = trainer.detect(y)                      # self.detector(received_word, 'val', ...) with each run's weights
        if bool(data_mask[c]):
            decoded = ops.rs_decode(detected, n_symbols)
            ser32 = (decoded != info[:, c]).float().sum(dim=1) / n_info        # calculate_error_rates: fp32 mean
            encoded = ops.rs_encode(decoded, n_symbols)
            label = torch.where((ser32 > 0).unsqueeze(1), detected, encoded)  # trainer.py:322-324
            ser = ser32.double()
            ser_by_word[:, c] = ser
        else:                                             # pilot: the receiver knows the word (trainer.py:310-316)
            label = ops.rs_encode(info[:, c].contiguous(), n_symbols)
            ser = torch.zeros(R, dtype=torch.float64, device=rx.device)
        gate = ser <= float(ser_thresh)
        if self_supervised:
            keep = [t.clone() for t in (trainer.theta, trainer.adam_m, trainer.adam_v, trainer.adam_step)]
            if restart_from_saved:                        # metavnet_trainer.py:59
                trainer.theta.copy_(saved)
            for _ in range(int(iterations)):
                trainer.train_step(y, label)
            g = gate.unsqueeze(1)                         # runs whose word did not pass the gate keep their state
            trainer.theta.copy_(torch.where(g, trainer.theta, keep[0]))
            trainer.adam_m.copy_(torch.where(g, trainer.adam_m, keep[1]))
            trainer.adam_v.copy_(torch.where(g, trainer.adam_v, keep[2]))
            trainer.adam_step.copy_(torch.where(gate, trainer.adam_step, keep[3]))
        if on_block is not None:
            on_block(c, ser, gate)
    return ser_by_word

#!/usr/bin/env python
"""Golden fixture for the word-by-word online evaluation, generated FROM THE LIVE REFERENCE:
METAVNETTrainer.eval_by_word (trainers/trainer.py:267-354) with self-supervised online training
(metavnet_trainer.py:52-64), online_meta off, on a short sequence of RS-coded words over a fading ISI channel.

    python tests/golden/make_golden_online.py            # online.npz: runs a, b, m, n (time_decay taps)
    python tests/golden/make_golden_online.py cost2100   # online_cost2100.npz: runs c, d on the COST2100 taps
                                                         # (BASELINE.json configs[3]; channel_estimation.py:26-30)

Recorded: the words the reference's dataset drew (information bits, channel outputs), the detector weights before the
run (after a short supervised warm-up so that detection mostly works), the SER per word the reference returns and the
detector weights after every block.  Nothing is copied from the reference: it is only called.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    mg._import_reference()
    import torch
    from python_code.trainers.META_VNET.metavnet_trainer import METAVNETTrainer
    from python_code.ecc.rs_main import encode
    out = {}
    cost = len(sys.argv) > 1 and sys.argv[1] == 'cost2100'
    runs = ((('c', 10.0, False, False), ('d', 10.0, False, True)) if cost else
            (('a', 9.0, True, False), ('b', 6.0, False, False), ('m', 9.0, True, True), ('n', 7.0, True, True)))
    with tempfile.TemporaryDirectory() as tmp:
        if cost:
            import python_code.channel.channel_estimation as ce
            ce.COST2100_DIR = mg._cost2100_dir(tmp)
        for tag, snr, fading, meta in runs:
            wd = os.path.join(tmp, f'w_online_{tag}')
            os.makedirs(wd, exist_ok=True)
            torch.manual_seed(21)
            tr = METAVNETTrainer(memory_length=4, use_ecc=True, n_symbols=2, val_frames=3, subframes_in_frame=4,
                                 train_frames=1, val_block_length=120, fading_in_channel=fading, fading_in_decoder=False,
                                 channel_coefficients='cost2100' if cost else 'time_decay', self_supervised=True, self_supervised_iterations=4,
                                 ser_thresh=0.02 if tag != 'n' else 0.05, online_meta=meta, buffer_empty=True, lr=1e-3,
                                 weights_dir=wd, eval_mode='by_word', meta_subframes=4, meta_train_iterations=2,
                                 meta_j_num=3, window_size=1, MAML=(tag != 'n'), meta_lr=0.1, weights_init='last_frame')
            tr.deep_learning_setup()
            # supervised warm-up on separately drawn words, so that the run starts from a detector that mostly works
            with torch.no_grad():
                b, y = tr.channel_dataset['train'].__getitem__(snr_list=[snr], gamma=0.2)
            for step in range(120):
                w = step % b.shape[0]
                tx = torch.Tensor(encode(b[w].int().numpy(), tr.n_symbols).reshape(1, -1))
                tr.run_train_loop(tr.detector(y[w].reshape(1, -1), 'train'), tx)
            for i, p in enumerate(tr.detector.parameters()):
                out[f'{tag}_w0_{i}'] = p.detach().numpy().copy()
            # record what the dataset hands to eval_by_word
            ds = tr.channel_dataset['val']
            cls = type(ds)
            orig_fn = cls.__dict__['__getitem__'] if not hasattr(cls, '_mvn_orig') else cls._mvn_orig
            cls._mvn_orig = orig_fn
            drawn = {}

            def recording_getitem(self, snr_list, gamma, _orig=orig_fn, _drawn=drawn, _ds=ds):
                bb, yy = _orig(self, snr_list=snr_list, gamma=gamma)
                if self is _ds:
                    _drawn['b'], _drawn['y'] = bb.numpy().copy(), yy.numpy().copy()
                return bb, yy
            cls.__getitem__ = recording_getitem
            after = []
            orig_online = tr.online_training

            def recording_online(tx, rx, _o=orig_online, _after=after, _tr=tr):
                _o(tx, rx)
                _after.append(np.concatenate([p.detach().numpy().reshape(-1) for p in _tr.detector.parameters()]))
            tr.online_training = recording_online
            saved_after = []
            if meta:          # weights right after every online meta-training round (what copy_model saves, trainer.py:343)
                orig_loop = tr.meta_train_loop
                jhats = []

                def recording_loop(rx, tx, sidx, qidx, _o=orig_loop, _j=jhats, _s=saved_after, _tr=tr):
                    r = _o(rx, tx, sidx, qidx)
                    _j.append(int(qidx[0]))
                    _s.append(np.concatenate([p.detach().numpy().reshape(-1) for p in _tr.detector.parameters()]))
                    return r
                tr.meta_train_loop = recording_loop
            torch.manual_seed(1000 + len(tag) + ord(tag))
            out[f'{tag}_seed'] = np.array([1000 + len(tag) + ord(tag)])
            ser = tr.eval_by_word(snr, 0.2)
            if meta:
                out[f'{tag}_jhat'] = np.array(jhats)
                out[f'{tag}_theta_meta'] = np.stack(saved_after).astype(np.float32) if saved_after else np.zeros((0, 1), np.float32)
            cls.__getitem__ = orig_fn
            out[f'{tag}_bits'] = drawn['b'].astype(np.uint8)
            out[f'{tag}_y'] = drawn['y'].astype(np.float32)
            out[f'{tag}_ser'] = np.asarray(ser, dtype=np.float64)
            out[f'{tag}_theta_after'] = np.stack(after).astype(np.float32)
            out[f'{tag}_data_indices'] = tr.data_indices.numpy()
            out[f'{tag}_cfg'] = np.array([tr.memory_length, tr.n_symbols, tr.self_supervised_iterations, tr.ser_thresh, tr.lr])
            out[f'{tag}_meta_cfg'] = np.array([int(meta), tr.meta_subframes, tr.meta_train_iterations, tr.meta_j_num,
                                               tr.window_size, int(tr.MAML), tr.meta_lr])
            print(tag, 'ser by word', ser, 'trained after', len(after), 'blocks')
    mg.save('online_cost2100' if cost else 'online', **out)


if __name__ == '__main__':
    main()

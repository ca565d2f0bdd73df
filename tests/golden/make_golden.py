#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE LIVE REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference (read-only) with the three workarounds of SURVEY.md §0.6
(matplotlib stub, pre-created weights_dir, COST2100 symlink dir), calls the reference's own
functions on seeded inputs and stores inputs + outputs as small ``.npz`` files.  The fixtures
pin ``oracle/viterbinet_oracle.py`` (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_parity.py, GPU).  Nothing here is copied from the reference: it is only called.
"""
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get('MVN_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f'reference checkout not found at {REF}')
    mpl = types.ModuleType('matplotlib')
    mpl.rcParams = {}
    mpl.pyplot = types.ModuleType('matplotlib.pyplot')
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', mpl.pyplot)
    sys.path.insert(0, REF)


def _cost2100_dir(tmp):
    d = os.path.join(tmp, 'cost2100')
    os.makedirs(d, exist_ok=True)
    for i in range(4):
        dst = os.path.join(d, f'combined_h_{i}.mat')
        if not os.path.exists(dst):
            os.symlink(os.path.join(REF, 'resources', 'cost2100_channel', f'h_{i}.mat'), dst)
    return d


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}.npz  {os.path.getsize(path) / 1024:.1f} KB')


def main():
    _import_reference()
    import torch
    from python_code.utils.trellis_utils import create_transition_table, acs_block, calculate_states
    from python_code.utils.metrics import calculate_error_rates
    from python_code.detectors.VNET.vnet_detector import VNETDetector
    from python_code.detectors.META_VNET.meta_vnet_detector import META_VNETDetector
    import python_code.channel.channel_estimation as ce
    from python_code.trainers.VA.va_trainer import VATrainer
    from python_code.trainers.VNET.vnet_trainer import VNETTrainer
    from python_code.trainers.META_VNET.metavnet_trainer import METAVNETTrainer

    tmp = tempfile.mkdtemp(prefix='mvn_golden_')
    ce.COST2100_DIR = _cost2100_dir(tmp)
    torch.set_num_threads(1)

    # ---------------- a1-a3: transition tables + stage loop on arbitrary costs -------------
    out = {}
    rng = np.random.RandomState(1234)
    for L in range(3, 9):
        S = 2 ** L
        table = create_transition_table(S)
        out[f'table_L{L}'] = table
        ttab = torch.Tensor(table)
        for kind in ('rand', 'tie'):
            B, T = 6, 20
            if kind == 'rand':
                cost = rng.randn(B, T, S).astype(np.float32) * 2
            else:  # small integers -> many exact ties
                cost = rng.randint(0, 3, size=(B, T, S)).astype(np.float32)
            c = torch.tensor(cost)
            pm = torch.zeros(B, S)
            dec = torch.zeros(B, T)
            survs = np.zeros((B, T, S), dtype=np.int8)
            for i in range(T):
                dec[:, i] = torch.argmin(pm, dim=1) % 2
                pm, idx = acs_block(pm, c[:, i], ttab, S)
                survs[:, i] = idx.numpy()
            out[f'cost_{kind}_L{L}'] = cost
            out[f'dec_{kind}_L{L}'] = dec.numpy()
            out[f'pm_{kind}_L{L}'] = pm.numpy()
            out[f'surv_{kind}_L{L}'] = survs
    save('acs', **out)

    # ---------------- a8 / a12: labels and error rates -----------------------------------
    out = {}
    for L in (3, 4, 6, 8):
        tx = rng.randint(0, 2, size=(5, 37)).astype(np.float32)
        out[f'tx_L{L}'] = tx
        out[f'states_L{L}'] = calculate_states(L, torch.tensor(tx)).numpy()
    for k, (W, T, p) in enumerate([(24, 120, 0.02), (7, 33, 0.3), (300, 120, 0.001), (4, 16, 0.0)]):
        tgt = rng.randint(0, 2, size=(W, T)).astype(np.float32)
        flips = (rng.rand(W, T) < p)
        pred = np.where(flips, 1 - tgt, tgt).astype(np.float32)
        ber, fer, idx = calculate_error_rates(torch.tensor(pred), torch.tensor(tgt))
        out[f'pred_{k}'] = pred
        out[f'tgt_{k}'] = tgt
        out[f'ber_fer_{k}'] = np.array([ber, fer], dtype=np.float64)
        out[f'idx_{k}'] = idx.numpy()
    save('labels_metrics', **out)

    # ---------------- a4/a5 + a3: VADetector on the reference's own data generator ---------
    def va_case(name, **kw):
        wd = os.path.join(tmp, 'w_' + name)
        os.makedirs(wd, exist_ok=True)
        args = dict(val_frames=2, subframes_in_frame=25, val_SNR_start=10, val_SNR_end=10, gamma=0.2,
                    noisy_est_var=0, weights_dir=wd)
        args.update(kw)
        tr = VATrainer(**args)
        snr = float(tr.snr_range['val'][0])
        with torch.no_grad():
            b, y = tr.channel_dataset['val'].__getitem__(snr_list=[snr], gamma=tr.gamma)
            dec = tr.detector(y, 'val', snr, tr.gamma)
            W = y.shape[0]
            h = np.concatenate([ce.estimate_channel(tr.memory_length, tr.gamma, noisy_est_var=0,
                                                    fading=tr.fading_in_decoder, index=i,
                                                    fading_taps_type=tr.fading_taps_type,
                                                    channel_coefficients=tr.channel_coefficients['val'])
                                for i in range(W)], axis=0)
            sp = tr.detector.compute_state_priors(h).numpy()
            pri = tr.detector.compute_likelihood_priors(y, snr, tr.gamma, 'val').numpy()
            # single-word calls with count (the eval_by_word shape, trainer.py:295)
            dec_cnt = np.stack([tr.detector(y[i:i + 1], 'val', snr, tr.gamma, i)[0].numpy() for i in (0, 7, W - 1)])
        return {f'{name}_b': b.numpy(), f'{name}_y': y.numpy(), f'{name}_dec': dec.numpy(), f'{name}_h': h,
                f'{name}_sp': sp, f'{name}_cost_t0': pri[:, :3].copy(), f'{name}_dec_count': dec_cnt,
                f'{name}_meta': np.array([tr.memory_length, tr.transmission_lengths['val'], snr, tr.gamma,
                                          float(tr.fading_in_decoder), tr.fading_taps_type], dtype=np.float64)}

    out = {}
    out.update(va_case('L4_fade1_ecc', memory_length=4, use_ecc=True, n_symbols=2, fading_in_channel=True,
                       fading_in_decoder=True, fading_taps_type=1, channel_coefficients='time_decay'))
    out.update(va_case('L4_fade2', memory_length=4, use_ecc=False, fading_in_channel=True,
                       fading_in_decoder=True, fading_taps_type=2, channel_coefficients='time_decay'))
    out.update(va_case('L4_cost2100_ecc', memory_length=4, use_ecc=True, n_symbols=2, fading_in_channel=False,
                       fading_in_decoder=False, fading_taps_type=1, channel_coefficients='cost2100'))
    for L in (3, 5, 6, 7, 8):
        out.update(va_case(f'L{L}_static', memory_length=L, use_ecc=False, fading_in_channel=False,
                           fading_in_decoder=False, fading_taps_type=1, channel_coefficients='time_decay',
                           val_block_length=60))
    save('va', **out)

    # ---------------- a6 + a3: ViterbiNet, default-init and reference-trained weights ---------
    def weights_of(det):
        return [p.detach().numpy().copy() for p in det.parameters()]

    out = {}
    wd = os.path.join(tmp, 'w_vnet')
    os.makedirs(wd, exist_ok=True)
    torch.manual_seed(0)
    tr = VNETTrainer(memory_length=4, use_ecc=True, n_symbols=2, val_frames=2, train_frames=4,
                     train_minibatch_num=6, fading_in_channel=True, fading_in_decoder=True, fading_taps_type=1,
                     channel_coefficients='time_decay', val_SNR_start=10, val_SNR_end=10, weights_dir=wd)
    snr, gamma = 10.0, 0.2
    with torch.no_grad():
        b, y = tr.channel_dataset['val'].__getitem__(snr_list=[snr], gamma=gamma)
        out['init_w'] = np.array(weights_of(tr.detector), dtype=object)
        out['y'] = y.numpy()
        out['b'] = b.numpy()
        out['init_priors'] = tr.detector(y, 'train').numpy()
        out['init_dec'] = tr.detector(y, 'val').numpy()
    tr.train()                       # the reference's own training loop (Adam, CE on sampled states)
    ck = torch.load(os.path.join(wd, f'snr_{tr.snr_range["train"][0]}_gamma_{gamma}.pt'))
    tr.detector.load_state_dict(ck['model_state_dict'])
    with torch.no_grad():
        out['trained_w'] = np.array(weights_of(tr.detector), dtype=object)
        out['trained_priors'] = tr.detector(y, 'train').numpy()
        out['trained_dec'] = tr.detector(y, 'val').numpy()
        # shorter loop than y.shape[1]: transmission_lengths['val'] governs (vnet_detector.py:53)
        det_short = VNETDetector(16, {'val': 100, 'train': 100})
        det_short.load_state_dict(ck['model_state_dict'])
        out['trained_dec_T100'] = det_short(y, 'val').numpy()
    for k in ('init_w', 'trained_w'):
        ws = out.pop(k)
        for i, w in enumerate(ws):
            out[f'{k}{i}'] = w
    # other trellis sizes, default init
    for L in (3, 5, 6, 7, 8):
        torch.manual_seed(L)
        det = VNETDetector(2 ** L, {'val': 40, 'train': 40})
        yy = torch.tensor(rng.randn(8, 40).astype(np.float32) * 1.5)
        with torch.no_grad():
            for i, w in enumerate(weights_of(det)):
                out[f'L{L}_w{i}'] = w
            out[f'L{L}_y'] = yy.numpy()
            out[f'L{L}_priors'] = det(yy, 'train').numpy()
            out[f'L{L}_dec'] = det(yy, 'val').numpy()
    save('vnet', **out)

    # ---------------- a10 / a11: meta_train_loop and run_train_loop on the reference ---------
    out = {}
    for maml in (True, False):
        wd = os.path.join(tmp, f'w_meta_{maml}')
        os.makedirs(wd, exist_ok=True)
        torch.manual_seed(7)
        tr = METAVNETTrainer(memory_length=4, use_ecc=True, n_symbols=2, val_frames=1, train_frames=1,
                             fading_in_channel=False, fading_in_decoder=False, channel_coefficients='cost2100',
                             MAML=maml, meta_lr=0.1, lr=1e-3, window_size=1, weights_dir=wd)
        tr.deep_learning_setup()
        tag = 'maml' if maml else 'fo'
        with torch.no_grad():
            b, y = tr.channel_dataset['val'].__getitem__(snr_list=[10.0], gamma=0.2)
        # transmitted words for the loss are the RS-ENCODED words (trainer.py:407-410)
        from python_code.ecc.rs_main import encode
        tx = torch.cat([torch.Tensor(encode(w.int().numpy(), tr.n_symbols).reshape(1, -1)) for w in b], dim=0)
        out[f'{tag}_y'] = y.numpy()[:6]
        out[f'{tag}_tx'] = tx.numpy()[:6]
        for i, w in enumerate(weights_of(tr.detector)):
            out[f'{tag}_w0_{i}'] = w
        losses = []
        for step, j_hat in enumerate([1, 2, 4]):
            sidx = torch.tensor([j_hat - 1]).long()
            qidx = torch.tensor([j_hat]).long()
            lq = tr.meta_train_loop(y, tx, sidx, qidx)
            losses.append(float(lq))
            for i, p in enumerate(tr.detector.parameters()):
                out[f'{tag}_w{step + 1}_{i}'] = p.detach().numpy().copy()
                if step == 0:
                    out[f'{tag}_g1_{i}'] = p.grad.detach().numpy().copy()
        out[f'{tag}_loss_q'] = np.array(losses)
        out[f'{tag}_jhat'] = np.array([1, 2, 4])
    # plain training steps (online_training inner loop, metavnet_trainer.py:52-64)
    wd = os.path.join(tmp, 'w_sgd')
    os.makedirs(wd, exist_ok=True)
    torch.manual_seed(11)
    tr = METAVNETTrainer(memory_length=4, use_ecc=True, n_symbols=2, val_frames=1, train_frames=1,
                         fading_in_channel=False, fading_in_decoder=False, channel_coefficients='time_decay',
                         weights_dir=wd)
    tr.deep_learning_setup()
    with torch.no_grad():
        b, y = tr.channel_dataset['val'].__getitem__(snr_list=[10.0], gamma=0.2)
    from python_code.ecc.rs_main import encode
    tx = torch.Tensor(encode(b[3].int().numpy(), 2).reshape(1, -1))
    rx = y[3].reshape(1, -1)
    out['sgd_y'] = rx.numpy()
    out['sgd_tx'] = tx.numpy()
    for i, w in enumerate(weights_of(tr.detector)):
        out[f'sgd_w0_{i}'] = w
    losses = []
    for step in range(5):
        soft = tr.detector(rx, 'train')
        losses.append(tr.run_train_loop(soft, tx))
    for i, p in enumerate(tr.detector.parameters()):
        out[f'sgd_w5_{i}'] = p.detach().numpy().copy()
    out['sgd_losses'] = np.array(losses)
    save('meta', **out)


if __name__ == '__main__':
    main()

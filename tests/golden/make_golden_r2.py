#!/usr/bin/env python
"""Round-2 golden fixtures, generated FROM THE LIVE REFERENCE (build container only):

    python tests/golden/make_golden_r2.py [channel] [vnet4096] [config1] [ckpt]

* ``channel.npz``  — one ``ChannelModelDataset`` draw per tap model (time_decay static, fading types 1 and 2,
  COST2100) together with the ``RandomState(noise_seed).normal`` stream the reference consumed (regenerated from
  the same seed and asserted to reproduce ``y``).  Pins ``oracle.isi_awgn`` and ``mvn_channel_transmit``
  (reference: channel/channel.py:12-35, channel/channel_dataset.py:55-95, channel/modulator.py:12).
* ``vnet4096.npz`` — 4 100 words decoded by the reference's full ``VNETDetector.forward(y, 'val')`` with the
  reference-trained weights of ``vnet.npz`` (protocol (ii) of SURVEY.md §8c on a real sample).
* ``config1.npz``  — BASELINE.json configs[0] at its real size: ``val_frames=12`` -> 300 blocks, VA and ViterbiNet,
  with and without ECC, incl. the reference's SER on the data rows.
* ``ckpt_snr*.npz`` — weights trained by the reference's own ``VNETTrainer.train()`` (trainer.py:455-490) at every SNR
  of the plotter's sweep (plotter_main.py:117-122), T = 120 uncoded: the checkpoints bench.py decodes with.

Nothing is copied from the reference: it is only called.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def channel_fixture(tmp, torch, ce, VATrainer):
    out = {}
    cases = {
        'static_ecc': dict(use_ecc=True, n_symbols=2, fading_in_channel=False, channel_coefficients='time_decay'),
        'fade1': dict(use_ecc=False, fading_in_channel=True, fading_taps_type=1, channel_coefficients='time_decay'),
        'fade2_ecc': dict(use_ecc=True, n_symbols=2, fading_in_channel=True, fading_taps_type=2,
                          channel_coefficients='time_decay'),
        'cost2100': dict(use_ecc=False, fading_in_channel=False, channel_coefficients='cost2100'),
        'L6_static': dict(use_ecc=False, fading_in_channel=False, channel_coefficients='time_decay', memory_length=6,
                          val_block_length=60),
    }
    from python_code.ecc.rs_main import encode
    for name, kw in cases.items():
        wd = os.path.join(tmp, 'w_ch_' + name)
        os.makedirs(wd, exist_ok=True)
        args = dict(memory_length=4, val_frames=2, subframes_in_frame=25, val_SNR_start=9, val_SNR_end=9, gamma=0.2,
                    noisy_est_var=0, fading_in_decoder=False, weights_dir=wd)
        args.update(kw)
        tr = VATrainer(**args)
        ds = tr.channel_dataset['val']
        snr, L, T = 9.0, tr.memory_length, tr.transmission_lengths['val']
        database = []
        ds.get_snr_data(snr, tr.gamma, database)            # float64, before the cast of __getitem__
        b, y64 = database[0]
        W = b.shape[0]
        # what the reference's own generators produced, regenerated from the seeds (trainer.py:90-91)
        wr = np.random.RandomState(tr.word_seed)
        nr = np.random.RandomState(tr.noise_seed)
        b2 = np.concatenate([wr.randint(0, 2, size=(1, tr.block_lengths['val'])) for _ in range(W)])
        assert np.array_equal(b, b2)
        noise = np.concatenate([nr.normal(0, 1, (1, T)) for _ in range(W)])
        c = np.stack([np.asarray(encode(w, tr.n_symbols)).reshape(-1) if tr.use_ecc else w for w in b.astype(int)])
        h = np.concatenate([ce.estimate_channel(L, tr.gamma, channel_coefficients=tr.channel_coefficients['val'],
                                                noisy_est_var=0, fading=tr.fading_in_channel, index=i,
                                                fading_taps_type=tr.fading_taps_type) for i in range(W)])
        # the reference's own arithmetic on the regenerated noise reproduces its output bit for bit
        s = 1 - 2 * np.concatenate([c, np.zeros((W, L))], axis=1)
        chk = np.stack([(np.dot(h[i:i + 1, ::-1], np.concatenate([s[i:i + 1, k:-L + k] for k in range(L)], axis=0)) +
                         (10 ** (snr / 10)) ** (-0.5) * noise[i:i + 1])[0] for i in range(W)])
        assert np.array_equal(chk, y64), name
        out[f'{name}_b'] = b.astype(np.uint8)
        out[f'{name}_c'] = c.astype(np.uint8)
        out[f'{name}_h'] = h
        out[f'{name}_noise'] = noise
        out[f'{name}_y64'] = y64
        out[f'{name}_y'] = torch.Tensor(y64).numpy()        # the cast channel_dataset.py:103 applies
        out[f'{name}_meta'] = np.array([L, T, snr, tr.gamma, tr.noise_seed, tr.word_seed], dtype=np.float64)
        print(name, 'words', W, 'T', T)
    mg.save('channel', **out)
    # the COST2100 tap magnitudes themselves (resources/cost2100_channel/h_{0..3}.mat, 300 blocks x 4 taps): input DATA of
    # BASELINE.json configs[3], needed on the GPU box where the reference checkout does not exist
    import scipy.io
    taps = np.stack([scipy.io.loadmat(os.path.join(mg.REF, 'resources', 'cost2100_channel', f'h_{i}.mat'))['h_channel_response_mag'].reshape(-1)
                     for i in range(4)], axis=1)
    mg.save('cost2100_taps', taps=taps)


def vnet4096_fixture(tmp, torch, VNETTrainer):
    g = np.load(os.path.join(HERE, 'vnet.npz'))
    wd = os.path.join(tmp, 'w_v4096')
    os.makedirs(wd, exist_ok=True)
    tr = VNETTrainer(memory_length=4, use_ecc=True, n_symbols=2, val_frames=164, fading_in_channel=True,
                     fading_in_decoder=True, fading_taps_type=1, channel_coefficients='time_decay', val_SNR_start=10,
                     val_SNR_end=10, weights_dir=wd)
    sd = tr.detector.state_dict()
    for i, k in enumerate(sd):
        sd[k] = torch.tensor(g[f'trained_w{i}'])
    tr.detector.load_state_dict(sd)
    torch.set_num_threads(1)
    with torch.no_grad():
        b, y = tr.channel_dataset['val'].__getitem__(snr_list=[10.0], gamma=0.2)
        dec = tr.detector(y, 'val').numpy()
    print('vnet4096: words', y.shape, 'BER vs info bits (first 120 cols)', float((dec[:, 1:120] != b.numpy()[:, 1:120]).mean()))
    mg.save('vnet4096', y=y.numpy(), dec_packed=np.packbits(dec.astype(np.uint8), axis=1),
            b_packed=np.packbits(b.numpy().astype(np.uint8), axis=1), T=np.array([y.shape[1], b.shape[1]]))


def config1_fixture(tmp, torch, ce, VATrainer, VNETTrainer):
    from python_code.utils.metrics import calculate_error_rates
    g = np.load(os.path.join(HERE, 'vnet.npz'))
    out = {}
    for ecc in (False, True):
        tag = 'ecc' if ecc else 'raw'
        kw = dict(memory_length=4, use_ecc=ecc, n_symbols=2, val_frames=12, subframes_in_frame=25, fading_in_channel=True,
                  fading_in_decoder=True, fading_taps_type=1, channel_coefficients='time_decay', val_SNR_start=10,
                  val_SNR_end=10, gamma=0.2, noisy_est_var=0)
        wd = os.path.join(tmp, f'w_c1_va_{tag}')
        os.makedirs(wd, exist_ok=True)
        va = VATrainer(weights_dir=wd, **kw)
        with torch.no_grad():
            b, y = va.channel_dataset['val'].__getitem__(snr_list=[10.0], gamma=0.2)
            dec_va = va.detector(y, 'val', 10.0, 0.2)
        W = y.shape[0]
        assert W == 300
        h = np.concatenate([ce.estimate_channel(4, 0.2, channel_coefficients='time_decay', noisy_est_var=0, fading=True,
                                                index=i, fading_taps_type=1) for i in range(W)])
        wd = os.path.join(tmp, f'w_c1_vn_{tag}')
        os.makedirs(wd, exist_ok=True)
        vn = VNETTrainer(weights_dir=wd, **kw)
        sd = vn.detector.state_dict()
        for i, k in enumerate(sd):
            sd[k] = torch.tensor(g[f'trained_w{i}'])
        vn.detector.load_state_dict(sd)
        with torch.no_grad():
            dec_vn = vn.detector(y, 'val')
            pri = vn.detector(y[:25], 'train')
        out[f'{tag}_b'] = b.numpy().astype(np.uint8)
        out[f'{tag}_y'] = y.numpy()
        out[f'{tag}_h'] = h
        out[f'{tag}_dec_va'] = np.packbits(dec_va.numpy().astype(np.uint8), axis=1)
        out[f'{tag}_dec_vnet'] = np.packbits(dec_vn.numpy().astype(np.uint8), axis=1)
        out[f'{tag}_priors25'] = pri.numpy()
        if not ecc:     # uncoded: the detected word is compared with the information bits directly (trainer.py:238-239)
            for nm, d in (('va', dec_va), ('vnet', dec_vn)):
                ser, fer, idx = calculate_error_rates(d[va.data_indices], b[va.data_indices])
                out[f'{tag}_rates_{nm}'] = np.array([ser, fer])
                out[f'{tag}_erridx_{nm}'] = idx.numpy()
                print('config1', tag, nm, 'ser', ser, 'fer', fer)
        else:           # coded: the reference's own RS decode + SER (trainer.py:234-239)
            from python_code.ecc.rs_main import decode
            for nm, d in (('va', dec_va), ('vnet', dec_vn)):
                dw = torch.Tensor(np.array([decode(w, 2) for w in d.numpy()]))
                ser, fer, idx = calculate_error_rates(dw[va.data_indices], b[va.data_indices])
                out[f'{tag}_rates_{nm}'] = np.array([ser, fer])
                out[f'{tag}_erridx_{nm}'] = idx.numpy()
                print('config1', tag, nm, 'coded ser', ser, 'fer', fer)
        out[f'{tag}_data_indices'] = va.data_indices.numpy()
    mg.save('config1', **out)


def checkpoints(tmp, torch, VNETTrainer):
    """The reference's own training run per SNR point of the sweep; static time_decay channel, uncoded, T = 120 —
    the channel bench.py synthesises."""
    out = {}
    for snr in (7, 8, 9, 10, 11, 12):
        wd = os.path.join(tmp, f'w_ck_{snr}')
        os.makedirs(wd, exist_ok=True)
        torch.manual_seed(100 + snr)
        tr = VNETTrainer(memory_length=4, use_ecc=False, val_frames=4, train_frames=12, train_minibatch_num=25,
                         fading_in_channel=False, fading_in_decoder=False, channel_coefficients='time_decay',
                         train_SNR_start=snr, train_SNR_end=snr, val_SNR_start=snr, val_SNR_end=snr, weights_dir=wd)
        tr.train()
        ck = torch.load(os.path.join(wd, f'snr_{snr}_gamma_0.2.pt'))
        tr.detector.load_state_dict(ck['model_state_dict'])
        ser = tr.single_eval_at_point(float(snr), 0.2)
        for i, p in enumerate(tr.detector.parameters()):
            out[f'snr{snr}_w{i}'] = p.detach().numpy().copy()
        out[f'snr{snr}_ser'] = np.array([ser])
        print('checkpoint snr', snr, 'reference SER on 100 fresh blocks', ser)
    mg.save('ckpt_vnet_L4', **out)


def main():
    which = set(sys.argv[1:]) or {'channel', 'vnet4096', 'config1', 'ckpt'}
    mg._import_reference()
    import torch
    import python_code.channel.channel_estimation as ce
    from python_code.trainers.VA.va_trainer import VATrainer
    from python_code.trainers.VNET.vnet_trainer import VNETTrainer
    tmp = tempfile.mkdtemp(prefix='mvn_golden_r2_')
    ce.COST2100_DIR = mg._cost2100_dir(tmp)
    if 'channel' in which:
        channel_fixture(tmp, torch, ce, VATrainer)
    if 'vnet4096' in which:
        vnet4096_fixture(tmp, torch, VNETTrainer)
    if 'config1' in which:
        config1_fixture(tmp, torch, ce, VATrainer, VNETTrainer)
    if 'ckpt' in which:
        checkpoints(tmp, torch, VNETTrainer)


if __name__ == '__main__':
    main()

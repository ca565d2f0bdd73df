"""Pins oracle/viterbinet_oracle.py against the fixtures produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import viterbinet_oracle as orc
from conftest import load_golden


@pytest.mark.parametrize('L', range(3, 9))
@pytest.mark.parametrize('kind', ['rand', 'tie'])
def test_acs_loop_bit_exact(L, kind):
    g = load_golden('acs')
    assert np.array_equal(orc.transition_table(2 ** L), g[f'table_L{L}'])
    dec, pm, surv = orc.acs_decode(g[f'cost_{kind}_L{L}'], return_survivors=True)
    assert np.array_equal(dec, g[f'dec_{kind}_L{L}'])
    assert np.array_equal(pm.view(np.uint32), g[f'pm_{kind}_L{L}'].view(np.uint32))
    assert np.array_equal(surv, g[f'surv_{kind}_L{L}'])
    # structural facts the kernels rely on
    S = 2 ** L
    assert np.array_equal(pm[:, :S // 2], pm[:, S // 2:])
    assert np.all(dec[:, 0] == 0)


@pytest.mark.parametrize('L', [3, 4, 6, 8])
def test_calculate_states(L):
    g = load_golden('labels_metrics')
    assert np.array_equal(orc.calculate_states(L, g[f'tx_L{L}']), g[f'states_L{L}'])


@pytest.mark.parametrize('k', range(4))
def test_error_rates_identical(k):
    g = load_golden('labels_metrics')
    ber, fer, idx = orc.calculate_error_rates(g[f'pred_{k}'], g[f'tgt_{k}'])
    assert ber == g[f'ber_fer_{k}'][0] and fer == g[f'ber_fer_{k}'][1]
    assert np.array_equal(idx, g[f'idx_{k}'])


VA_CASES = ['L4_fade1_ecc', 'L4_fade2', 'L4_cost2100_ecc'] + [f'L{L}_static' for L in (3, 5, 6, 7, 8)]


@pytest.mark.parametrize('name', VA_CASES)
def test_va_bit_exact(name):
    g = load_golden('va')
    L, T = int(g[f'{name}_meta'][0]), int(g[f'{name}_meta'][1])
    h, y = g[f'{name}_h'], g[f'{name}_y']
    sp = orc.va_state_priors(h, L)
    assert np.array_equal(sp.view(np.uint32), g[f'{name}_sp'].view(np.uint32))
    cost = orc.va_cost(y, sp)
    assert np.array_equal(cost[:, :3].view(np.uint32), g[f'{name}_cost_t0'].view(np.uint32))
    assert np.array_equal(orc.va_decode(y, h, L, T), g[f'{name}_dec'])
    W = y.shape[0]
    for row, i in zip(g[f'{name}_dec_count'], (0, 7, W - 1)):
        assert np.array_equal(orc.va_decode(y[i:i + 1], h[i:i + 1], L, T)[0], row)


def test_va_taps_restated():
    g = load_golden('va')
    cases = {'L4_fade1_ecc': dict(fading=True, fading_taps_type=1),
             'L4_fade2': dict(fading=True, fading_taps_type=2),
             'L3_static': dict(fading=False)}
    for name, kw in cases.items():
        h = g[f'{name}_h']
        L = int(g[f'{name}_meta'][0])
        mine = np.concatenate([orc.estimate_channel(L, 0.2, 'time_decay', index=i, **kw) for i in range(h.shape[0])])
        assert np.array_equal(mine, h)


def _w(g, prefix):
    return [g[f'{prefix}{i}'] for i in range(6)]


def _rel_to_rowmax(a, ref):
    return np.max(np.abs(a - ref) / np.max(np.abs(ref), axis=-1, keepdims=True))


@pytest.mark.parametrize('tag', ['init', 'trained'])
def test_vnet_priors_and_decode(tag):
    g = load_golden('vnet')
    w, y = _w(g, f'{tag}_w'), g['y']
    pri = orc.vnet_priors(y, w)
    ref = g[f'{tag}_priors']
    # tolerance: 1e-5 relative to the row's max |prior|
    assert _rel_to_rowmax(pri, ref) < 1e-5
    # (i) the stage loop fed the REFERENCE's priors is bit-exact
    dec, _ = orc.vnet_decode_from_priors(ref)
    assert np.array_equal(dec, g[f'{tag}_dec'])
    # (ii) end to end: mismatches only through near-ties; tiny sample -> expect none
    assert np.mean(orc.vnet_decode(y, w) != g[f'{tag}_dec']) < 1e-3


def test_vnet_loop_length_from_config():
    g = load_golden('vnet')
    dec, _ = orc.vnet_decode_from_priors(g['trained_priors'], n_stages=100)
    assert np.array_equal(dec, g['trained_dec_T100'])
    assert np.all(dec[:, 100:] == 0)


@pytest.mark.parametrize('L', [3, 5, 6, 7, 8])
def test_vnet_other_trellis_sizes(L):
    g = load_golden('vnet')
    w = [g[f'L{L}_w{i}'] for i in range(6)]
    pri = orc.vnet_priors(g[f'L{L}_y'], w)
    assert _rel_to_rowmax(pri, g[f'L{L}_priors']) < 1e-5
    dec, _ = orc.vnet_decode_from_priors(g[f'L{L}_priors'])
    assert np.array_equal(dec, g[f'L{L}_dec'])


@pytest.mark.parametrize('tag,second', [('maml', True), ('fo', False)])
def test_maml_step_matches_reference(tag, second):
    g = load_golden('meta')
    y, tx = g[f'{tag}_y'], g[f'{tag}_tx']
    w = [g[f'{tag}_w0_{i}'] for i in range(6)]
    state = None
    for step, j in enumerate(g[f'{tag}_jhat']):
        loss, mg, w, state = orc.maml_step(y[j - 1:j], tx[j - 1:j], y[j:j + 1], tx[j:j + 1], w, state, 4,
                                           meta_lr=0.1, lr=1e-3, second_order=second)
        assert abs(loss - g[f'{tag}_loss_q'][step]) < 1e-5 * abs(loss)
        if step == 0:
            for i in range(6):
                ref = g[f'{tag}_g1_{i}']
                assert np.max(np.abs(mg[i].reshape(ref.shape) - ref)) < 2e-5 * np.max(np.abs(ref)) + 1e-9
        for i in range(6):
            ref = g[f'{tag}_w{step + 1}_{i}']
            # Adam's first steps are sign-like (|dw| = lr): compare on the scale of the update
            assert np.max(np.abs(w[i].reshape(ref.shape) - ref)) < 2e-5


def test_train_steps_match_reference():
    g = load_golden('meta')
    w = [g[f'sgd_w0_{i}'] for i in range(6)]
    state = None
    for step in range(5):
        loss, w, state = orc.train_step(g['sgd_y'], g['sgd_tx'], w, state, 4, lr=1e-3)
        assert abs(loss - g['sgd_losses'][step]) < 1e-5 * abs(loss)
    for i in range(6):
        ref = g[f'sgd_w5_{i}']
        assert np.max(np.abs(w[i].reshape(ref.shape) - ref)) < 2e-5


def test_hvp_against_finite_differences():
    rng = np.random.RandomState(0)
    w = [rng.randn(100, 1) * .5, rng.randn(100) * .5, rng.randn(50, 100) * .1, rng.randn(50) * .1,
         rng.randn(8, 50) * .1, rng.randn(8) * .1]
    v = [rng.randn(*a.shape) for a in w]
    y = rng.randn(40)
    lab = rng.randint(0, 8, 40)
    hv = orc.hessian_vector_product(y, lab, w, v)
    eps = 1e-5
    gp = orc.loss_and_grads(y, lab, [a + eps * b for a, b in zip(w, v)])[1]
    gm = orc.loss_and_grads(y, lab, [a - eps * b for a, b in zip(w, v)])[1]
    for h, p, m in zip(hv, gp, gm):
        fd = (p - m) / (2 * eps)
        assert np.max(np.abs(h - fd)) < 1e-6 * max(1.0, np.max(np.abs(fd)))


def test_torch_port_matches_oracle():
    """The CPU-baseline port (what bench.py times) gives the oracle's / reference's bits."""
    import torch
    from oracle import torch_port as tp
    g = load_golden('vnet')
    w, y = _w(g, 'trained_w'), g['y']
    net = tp.make_net(16)
    tp.load_weights(net, w)
    with torch.no_grad():
        dec = tp.vnet_forward_val(net, torch.tensor(y), y.shape[1]).numpy()
    assert np.array_equal(dec, g['trained_dec'])
    gv = load_golden('va')
    name = 'L4_fade1_ecc'
    sp = orc.va_state_priors(gv[f'{name}_h'], 4)
    with torch.no_grad():
        dec = tp.va_forward_val(torch.tensor(gv[f'{name}_y']), torch.tensor(sp), int(gv[f'{name}_meta'][1])).numpy()
    assert np.array_equal(dec, gv[f'{name}_dec'])


@pytest.mark.parametrize('L', [2, 3, 4])
def test_mlse_traceback_is_the_minimum_cost_path(L):
    """The oracle's traceback path attains the final minimum metric exactly, and no other bit sequence is cheaper
    (brute force over all sequences for a short trellis)."""
    import itertools
    rng = np.random.RandomState(L)
    S, T, B = 2 ** L, 7, 3
    cost = rng.randn(B, T, S).astype(np.float32)
    cost[1] = rng.randint(0, 3, size=(T, S))            # ties
    dec, states, pm = orc.mlse_decode(cost)
    assert np.array_equal(orc.path_cost(cost, states), pm.min(axis=1))
    assert np.array_equal(dec, (states[:, :-1] & 1).astype(np.float32))
    for b in range(B):
        best = np.inf
        for bits in itertools.product((0, 1), repeat=T + L - 1):
            st = [sum(bits[t + i] << i for i in range(L)) for t in range(T)]
            best = min(best, float(orc.path_cost(cost[b:b + 1], np.array([st + [0]]))[0]))
        assert best == float(pm[b].min())
    # terminated variant: forced final state 0
    dec0, states0, _ = orc.mlse_decode(cost, start_state=0)
    assert np.all(states0[:, -1] == 0)
    assert np.array_equal(orc.path_cost(cost, states0), pm[:, 0])


CHANNEL_CASES = ['static_ecc', 'fade1', 'fade2_ecc', 'cost2100', 'L6_static']


@pytest.mark.parametrize('name', CHANNEL_CASES)
def test_channel_simulator_pinned_on_reference_draw(name):
    """f1: orc.isi_awgn fed the reference's own RandomState(noise_seed) stream reproduces a recorded
    ChannelModelDataset draw (channel.py:12-35, channel_dataset.py:55-95): the fp32 words the dataset hands to the
    detectors bit for bit, and the float64 pre-cast values too at memory_length 4 (at L=6 numpy's BLAS dot sums the six
    taps in another order than the restatement: 1 ulp of float64, invisible after the fp32 cast)."""
    g = load_golden('channel')
    L, T, snr, gamma, noise_seed, _ = g[f'{name}_meta']
    L = int(L)
    c, h = g[f'{name}_c'].astype(np.float64), g[f'{name}_h']
    assert c.shape[1] == int(T)
    rng = np.random.RandomState(int(noise_seed))
    y = np.concatenate([orc.isi_awgn(c[i:i + 1], h[i:i + 1], snr, L, rng) for i in range(c.shape[0])])  # one draw per word
    assert np.array_equal(y.astype(np.float32).view(np.uint32), g[f'{name}_y'].view(np.uint32))
    if L == 4:
        assert np.array_equal(y, g[f'{name}_y64'])
    else:
        assert np.max(np.abs(y - g[f'{name}_y64'])) <= 2 ** -50
    # the recorded noise IS that stream, and the taps are the restated estimate_channel
    assert np.array_equal(np.random.RandomState(int(noise_seed)).normal(0, 1, g[f'{name}_noise'].shape), g[f'{name}_noise'])
    if name != 'cost2100':
        kw = {'fade1': dict(fading=True, fading_taps_type=1), 'fade2_ecc': dict(fading=True, fading_taps_type=2)}.get(name, {})
        mine = np.concatenate([orc.estimate_channel(L, gamma, 'time_decay', index=i, **kw) for i in range(h.shape[0])])
        assert np.array_equal(mine, h)


def test_vnet_4100_frames_full_forward_protocol():
    """Protocol (ii) of SURVEY.md §8c at a real sample size: 4 100 words decoded by the reference's full
    VNETDetector.forward (reference-trained weights); every frame the oracle decodes differently must be a near-tie."""
    from parity_utils import explain_mismatches, unpack_rows, PRIOR_RTOL
    g, v = load_golden('vnet4096'), load_golden('vnet')
    w, y, T = _w(v, 'trained_w'), g['y'], int(g['T'][0])
    ref = unpack_rows(g['dec_packed'], T)
    dec = orc.vnet_decode(y, w)
    exact = lambda rows: orc.vnet_priors(y[rows], w, dtype=np.float64)
    tol = PRIOR_RTOL * 40.0 * np.ones(y.shape[0])       # priors of the trained net reach |p| ~ 40
    assert explain_mismatches(dec, ref, exact, tol) <= 2


@pytest.mark.parametrize('tag', ['raw', 'ecc'])
def test_config1_300_blocks(tag):
    """BASELINE.json configs[0] at its real size (val_frames=12 -> 300 blocks, fading taps, 10 dB): VA bit-exact,
    ViterbiNet by protocol (ii), and the reference's SER / FER on the data rows (RS-decoded when coded)."""
    from oracle import rs_oracle
    from parity_utils import explain_mismatches, unpack_rows, PRIOR_RTOL
    g, v = load_golden('config1'), load_golden('vnet')
    y, h, b = g[f'{tag}_y'], g[f'{tag}_h'], g[f'{tag}_b'].astype(np.float32)
    T = y.shape[1]
    assert y.shape[0] == 300
    dec_va = orc.va_decode(y, h, 4, T)
    assert np.array_equal(dec_va, unpack_rows(g[f'{tag}_dec_va'], T))
    w = _w(v, 'trained_w')
    ref_vn = unpack_rows(g[f'{tag}_dec_vnet'], T)
    pri = orc.vnet_priors(y, w)
    assert _rel_to_rowmax(pri[:25], g[f'{tag}_priors25']) < 1e-5
    dec_vn, _ = orc.vnet_decode_from_priors(pri)
    exact = lambda rows: orc.vnet_priors(y[rows], w, dtype=np.float64)
    assert explain_mismatches(dec_vn, ref_vn, exact, PRIOR_RTOL * 40.0 * np.ones(300)) <= 1
    rows = g[f'{tag}_data_indices']
    for nm, d in (('va', dec_va), ('vnet', ref_vn)):
        msg = np.stack([rs_oracle.decode(wd.astype(int), 2) for wd in d]).astype(np.float32) if tag == 'ecc' else d
        ber, fer, idx = orc.calculate_error_rates(msg[rows], b[rows])
        assert ber == g[f'{tag}_rates_{nm}'][0] and fer == g[f'{tag}_rates_{nm}'][1]
        assert np.array_equal(idx, g[f'{tag}_erridx_{nm}'])


def test_reference_trained_checkpoints_decode_sanely():
    """the checkpoints bench.py decodes with (reference VNETTrainer.train() per SNR): oracle BER on a fresh synthetic
    draw is in the range the reference reported for them"""
    g = load_golden('ckpt_vnet_L4')
    rng = np.random.RandomState(5)
    bits = rng.randint(0, 2, size=(400, 120))
    h = np.exp(-0.2 * np.arange(4)).reshape(1, 4)
    for snr in (7, 10, 12):
        w = [g[f'snr{snr}_w{i}'] for i in range(6)]
        y = orc.isi_awgn(bits, h, float(snr), 4, rng).astype(np.float32)
        ber = float((orc.vnet_decode(y, w) != bits)[:, 1:].mean())
        assert 0.3 * g[f'snr{snr}_ser'][0] < ber < 3 * g[f'snr{snr}_ser'][0], (snr, ber)

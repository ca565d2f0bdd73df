"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on seeded inputs, against
the golden fixtures produced by the live reference, and — at full size — through size-independent
properties.  Run on the B200 box: ``python -m pytest tests -m gpu``.

Bars (BASELINE.json north_star / SURVEY.md §8c):
  * decoded bits, path metrics, survivor indices, labels, error counts: bit-exact;
  * priors: |kernel - reference| <= 1e-5 * (max |prior| of that symbol's row), fp32;
  * ViterbiNet decode: (i) the stage loop fed the kernel's own exported priors is bit-exact,
    (ii) against the full reference forward mismatching bits are counted and each must sit at a
    near-tie (competing metrics closer than the prior tolerance accumulated over the stages).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import viterbinet_oracle as orc

pytestmark = pytest.mark.gpu

PRIOR_RTOL = 1e-5


@pytest.fixture(scope='module')
def mvn():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import meta_viterbinet_b200 as m
    return m


@pytest.fixture(params=['tcgen05', 'fma', 'fma_smem'])
def fused_impl(request, mvn):
    """Every fused-ViterbiNet parity test runs on the tensor-core variant and on the FP32-FMA variants."""
    old = mvn.ops.set_fused_variant(request.param)
    yield request.param
    mvn.ops.set_fused_variant('auto')


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def rel_to_rowmax(a, ref):
    return float(np.max(np.abs(a - ref) / np.max(np.abs(ref), axis=-1, keepdims=True)))


def unpack_survivors(words, H):
    """[B,T,W] int32 -> [B,T,H] {0,1}"""
    w = words.cpu().numpy().view(np.uint32)
    bits = (w[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1
    return bits.reshape(w.shape[0], w.shape[1], -1)[..., :H].astype(np.int8)


# ------------------------------------------------------------------------------- a1-a3
@pytest.mark.parametrize('L', range(3, 9))
@pytest.mark.parametrize('kind', ['rand', 'tie'])
def test_acs_decode_golden(mvn, L, kind):
    g = load_golden('acs')
    cost = g[f'cost_{kind}_L{L}']
    dec, pm, surv = mvn.ops.acs_decode(cu(cost), return_final_pm=True, return_survivors=True)
    assert np.array_equal(dec.cpu().numpy(), g[f'dec_{kind}_L{L}'])
    assert np.array_equal(pm.cpu().numpy().view(np.uint32), g[f'pm_{kind}_L{L}'].view(np.uint32))
    H = 2 ** (L - 1)
    assert np.array_equal(unpack_survivors(surv, H), g[f'surv_{kind}_L{L}'][:, :, :H])
    assert np.array_equal(g[f'surv_{kind}_L{L}'][:, :, :H], g[f'surv_{kind}_L{L}'][:, :, H:])
    words = mvn.ops.acs_decode(cu(cost), out_format=mvn.OUT_BITS)
    assert np.array_equal(mvn.ops.unpack_bits(words, cost.shape[1]).cpu().numpy(), g[f'dec_{kind}_L{L}'])


@pytest.mark.parametrize('L', range(1, 9))
@pytest.mark.parametrize('B,T', [(1, 1), (33, 37), (257, 64), (1000, 45)])
def test_acs_decode_ragged_vs_oracle(mvn, L, B, T):
    rng = np.random.RandomState(100 * L + B)
    S = 2 ** L
    cost = (rng.randn(B, T, S) * 3).astype(np.float32)
    cost[::3] = rng.randint(-2, 3, size=cost[::3].shape)        # exact ties in a third of the frames
    n_stages = max(1, T - 5)
    ref_dec, ref_pm = orc.acs_decode(cost, n_stages)
    dec, pm = mvn.ops.acs_decode(cu(cost), n_stages, return_final_pm=True)
    assert np.array_equal(dec.cpu().numpy(), ref_dec)
    assert np.array_equal(pm.cpu().numpy(), ref_pm)


def test_acs_decode_empty(mvn):
    out = mvn.ops.acs_decode(torch.zeros(0, 8, 16).cuda())
    assert out.shape == (0, 8)


@pytest.mark.parametrize('L', [1, 3, 4, 8])
def test_acs_block_matches(mvn, L):
    rng = np.random.RandomState(L)
    S = 2 ** L
    pm = rng.randn(50, S).astype(np.float32)
    c = rng.randint(-1, 2, size=(50, S)).astype(np.float32)
    val, idx = mvn.acs_block(cu(pm), cu(c), None, S)
    rv, ri = orc.acs_stage(pm, c)
    assert np.array_equal(val.cpu().numpy(), rv) and np.array_equal(idx.cpu().numpy(), ri)
    assert idx.dtype == torch.int64
    val1, _ = mvn.acs_block(cu(pm), cu(c[:, :1]), None, S)      # [B,1] broadcast llrs
    assert np.array_equal(val1.cpu().numpy(), orc.acs_stage(pm, np.broadcast_to(c[:, :1], pm.shape))[0])


@pytest.mark.parametrize('L', [2, 3, 4, 5, 6, 7, 8])
def test_acs_decode_states_on_lanes(mvn, L):
    """The states-on-lanes layout (S/2 lanes per frame for 4..64 states, one warp per frame with S/64 metrics per lane at
    128 / 256 states; butterflies by shuffles, decision by REDUX.MIN + ballot): bits and final metrics bit-exact vs the
    oracle — random and exact-tie costs, +-0, ragged loop lengths, both output formats — and identical to the
    lane-per-frame kernel.  'auto' picks it at 128 / 256 states."""
    rng = np.random.RandomState(300 + L)
    S, B, T = 2 ** L, 77, 45
    cost = (rng.randn(B, T, S) * 2).astype(np.float32)
    cost[::3] = rng.randint(0, 3, size=cost[::3].shape)        # exact ties
    cost[1] = 0.0
    cost[2] = -0.0
    for n in (T, T - 13, 32, 1, 0):
        ref, pm_ref = orc.acs_decode(cost, n)
        for layout in ('states_on_lanes', 'lane_per_frame', 'auto'):
            dec, pm = mvn.ops.acs_decode(cu(cost), n, return_final_pm=True, layout=layout)
            assert np.array_equal(dec.cpu().numpy(), ref), (L, n, layout)
            assert np.array_equal(pm.cpu().numpy(), pm_ref), (L, n, layout)
            words = mvn.ops.acs_decode(cu(cost), n, out_format=mvn.OUT_BITS, layout=layout)
            assert np.array_equal(mvn.ops.unpack_bits(words, T).cpu().numpy(), ref)


# ------------------------------------------------------------------------------- a4/a5 VA
VA_CASES = ['L4_fade1_ecc', 'L4_fade2', 'L4_cost2100_ecc'] + [f'L{L}_static' for L in (3, 5, 6, 7, 8)]


@pytest.mark.parametrize('name', VA_CASES)
def test_va_decode_golden(mvn, name):
    from meta_viterbinet_b200.channel_taps import state_priors_table
    g = load_golden('va')
    L, T = int(g[f'{name}_meta'][0]), int(g[f'{name}_meta'][1])
    y, h = g[f'{name}_y'], g[f'{name}_h']
    table = cu(state_priors_table(h, L))
    dec = mvn.ops.va_decode(cu(y), table, T)
    assert np.array_equal(dec.cpu().numpy(), g[f'{name}_dec'])
    W = y.shape[0]
    for row, i in zip(g[f'{name}_dec_count'], (0, 7, W - 1)):        # eval_by_word shape: B=1, one tap block
        d1 = mvn.ops.va_decode(cu(y[i:i + 1]), table[i:i + 1].contiguous(), T)
        assert np.array_equal(d1.cpu().numpy()[0], row)
    words = mvn.ops.va_decode(cu(y), table, T, out_format=mvn.OUT_BITS)
    assert np.array_equal(mvn.ops.unpack_bits(words, y.shape[1]).cpu().numpy(), g[f'{name}_dec'])


@pytest.mark.parametrize('L', range(1, 9))
def test_va_decode_random_vs_oracle(mvn, L):
    rng = np.random.RandomState(L)
    n_h, reps, T = 7, 19, 50          # B = 133 (ragged), T % 4 != 0 -> scalar staging path
    h = np.abs(rng.randn(n_h, L)) + 0.1
    bits = rng.randint(0, 2, size=(n_h * reps, T))
    y = orc.isi_awgn(bits, np.tile(h, (reps, 1)), 8.0, L, rng).astype(np.float32)
    ref = orc.va_decode(y, h, L, T - 3)
    from meta_viterbinet_b200.channel_taps import state_priors_table
    dec = mvn.ops.va_decode(cu(y), cu(state_priors_table(h, L)), T - 3)
    assert np.array_equal(dec.cpu().numpy(), ref)


def test_va_batch_not_multiple_of_blocks_raises(mvn):
    with pytest.raises(mvn.MVNError):
        mvn.ops.va_decode(torch.zeros(10, 8).cuda(), torch.zeros(3, 16).cuda())


def test_va_detector_class_matches_reference_run(mvn):
    """VADetector built like va_trainer.py:32-40 reproduces the reference's decoded words."""
    g = load_golden('va')
    for name, kw in [('L4_fade1_ecc', dict(fading=True, fading_taps_type=1)),
                     ('L4_fade2', dict(fading=True, fading_taps_type=2)),
                     ('L6_static', dict(fading=False, fading_taps_type=1))]:
        L, T = int(g[f'{name}_meta'][0]), int(g[f'{name}_meta'][1])
        y = g[f'{name}_y']
        det = mvn.VADetector(n_states=2 ** L, memory_length=L, transmission_length=T, val_words=y.shape[0],
                             channel_type='ISI_AWGN', noisy_est_var=0, channel_coefficients={'train': 'time_decay',
                                                                                              'val': 'time_decay'}, **kw)
        out = det(cu(y), 'val', 10.0, 0.2)
        assert out.dtype == torch.float32 and out.shape == y.shape
        assert np.array_equal(out.cpu().numpy(), g[f'{name}_dec'])
        one = det(cu(y[7:8]), 'val', 10.0, 0.2, 7)
        assert np.array_equal(one.cpu().numpy()[0], g[f'{name}_dec_count'][1])
        sp = det.compute_state_priors(g[f'{name}_h'])
        assert np.array_equal(sp.cpu().numpy().view(np.uint32), g[f'{name}_sp'].view(np.uint32))
        with pytest.raises(NotImplementedError):
            det(cu(y), 'train', 10.0, 0.2)
    bad = mvn.VADetector(16, 4, 8, 4, 'OTHER', 0, False, 1, {'train': 'time_decay', 'val': 'time_decay'})
    with pytest.raises(Exception, match='No such channel defined'):
        bad(torch.zeros(4, 8).cuda(), 'val', 10.0, 0.2)


@pytest.mark.parametrize('L', [7, 8])
def test_va_states_on_lanes_counters(mvn, L):
    """128 / 256 states: the one-warp-per-frame VA kernel — bits bit-exact vs the oracle with several tap blocks, ragged
    loop length, bit-packed output, and the fused BER / FER counters with pilots and a target narrower than y."""
    from meta_viterbinet_b200.channel_taps import state_priors_table
    rng = np.random.RandomState(40 + L)
    B, T, Tt = 90, 70, 61
    h = (np.exp(-0.3 * np.arange(L)) * (1 + 0.1 * rng.randn(6, L))).astype(np.float64)     # 6 tap blocks, B % 6 == 0
    bits = rng.randint(0, 2, size=(B, T))
    y = orc.isi_awgn(bits, h[np.arange(B) % 6], 5.0, L, rng).astype(np.float32)
    y[:4] = np.round(y[:4])
    table = cu(state_priors_table(h, L))
    for n in (T, T - 9, 33):
        ref = orc.va_decode(y, h, L, n)
        cnt = mvn.ops.new_counters()
        tgt = bits[:, :Tt].astype(np.float32)
        dec = mvn.ops.va_decode(cu(y), table, n, target=cu(tgt), pilot_period=4, counters=cnt).cpu().numpy()
        assert np.array_equal(dec, ref), (L, n)
        keep = np.arange(B) % 4 != 0
        err = ref[keep][:, :Tt] != tgt[keep]
        assert cnt.tolist() == [int(err.sum()), int(err.any(axis=1).sum()), int(keep.sum()) * Tt, int(keep.sum())]
        words = mvn.ops.va_decode(cu(y), table, n, out_format=mvn.OUT_BITS)
        assert np.array_equal(mvn.ops.unpack_bits(words, T).cpu().numpy(), ref)


# ------------------------------------------------------------------------------- a6/a7 priors
def _w(g, prefix):
    return [g[f'{prefix}{i}'] for i in range(6)]


@pytest.mark.parametrize('tag', ['init', 'trained'])
def test_vnet_priors_golden(mvn, tag):
    g = load_golden('vnet')
    w, y = _w(g, f'{tag}_w'), g['y']
    pri = mvn.ops.vnet_priors(cu(y), [cu(a) for a in w]).cpu().numpy()
    assert rel_to_rowmax(pri, g[f'{tag}_priors']) < PRIOR_RTOL
    exact = orc.vnet_priors(y, w, dtype=np.float64)
    assert rel_to_rowmax(pri, exact) < PRIOR_RTOL


@pytest.mark.parametrize('L', [3, 5, 6, 7, 8])
def test_vnet_priors_other_sizes(mvn, L):
    g = load_golden('vnet')
    w = [g[f'L{L}_w{i}'] for i in range(6)]
    pri = mvn.ops.vnet_priors(cu(g[f'L{L}_y']), [cu(a) for a in w]).cpu().numpy()
    assert rel_to_rowmax(pri, g[f'L{L}_priors']) < PRIOR_RTOL


@pytest.mark.parametrize('L', [1, 2])
def test_vnet_priors_tiny_trellis(mvn, L):
    rng = np.random.RandomState(L)
    S = 2 ** L
    w = [rng.randn(100, 1) * .5, rng.randn(100) * .5, rng.randn(50, 100) * .1, rng.randn(50) * .1,
         rng.randn(S, 50) * .2, rng.randn(S) * .1]
    w = [a.astype(np.float32) for a in w]
    y = rng.randn(5, 13).astype(np.float32)
    pri = mvn.ops.vnet_priors(cu(y), [cu(a) for a in w]).cpu().numpy()
    assert rel_to_rowmax(pri, orc.vnet_priors(y, w, dtype=np.float64)) < PRIOR_RTOL


# ------------------------------------------------------------------------------- a6+a3 fused
from parity_utils import explain_mismatches as _explain_mismatches, unpack_rows  # noqa: E402


@pytest.mark.parametrize('tag', ['init', 'trained'])
def test_vnet_fused_decode_golden(mvn, fused_impl, tag):
    g = load_golden('vnet')
    w, y = _w(g, f'{tag}_w'), g['y']
    dec, pri = mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], return_priors=True)
    dec, pri = dec.cpu().numpy(), pri.cpu().numpy()
    # priors exported by the fused kernel == the priors kernel (same arithmetic), within tolerance of torch
    pri2 = mvn.ops.vnet_priors(cu(y), [cu(a) for a in w]).cpu().numpy()
    if fused_impl != 'tcgen05':     # same FMA arithmetic -> identical bits; the tensor-core path differs by ~1e-7
        assert np.array_equal(pri.view(np.uint32), pri2.view(np.uint32))
    assert rel_to_rowmax(pri, pri2) < PRIOR_RTOL
    assert rel_to_rowmax(pri, g[f'{tag}_priors']) < PRIOR_RTOL
    # (i) reference loop on the kernel's own priors: bit-exact
    ref_own, _ = orc.vnet_decode_from_priors(pri)
    assert np.array_equal(dec, ref_own)
    # (ii) against the reference's full forward
    tol = PRIOR_RTOL * np.abs(g[f'{tag}_priors']).max(axis=(1, 2))
    n_bad = _explain_mismatches(dec, g[f'{tag}_dec'], g[f'{tag}_priors'], tol)
    assert n_bad <= 1
    # small-batch path of the detector (priors kernel + ACS kernel) gives the same bits
    dec_small = mvn.ops.acs_decode(-cu(pri))
    assert np.array_equal(dec_small.cpu().numpy(), dec)


def test_vnet_fused_loop_length(mvn, fused_impl):
    g = load_golden('vnet')
    w, y = _w(g, 'trained_w'), g['y']
    dec = mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], n_stages=100).cpu().numpy()
    assert np.all(dec[:, 100:] == 0)
    n_bad = (dec != g['trained_dec_T100']).any(axis=1).sum()
    assert n_bad <= 1
    with pytest.raises(mvn.MVNError):
        mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], n_stages=y.shape[1] + 1)


@pytest.mark.parametrize('L', range(1, 9))
@pytest.mark.parametrize('B,T', [(1, 7), (70, 33), (515, 40)])
def test_vnet_fused_all_trellis_sizes(mvn, fused_impl, L, B, T):
    rng = np.random.RandomState(10 * L + B)
    S = 2 ** L
    w = [rng.randn(100, 1) * .7, rng.randn(100) * .5, rng.randn(50, 100) * .15, rng.randn(50) * .1,
         rng.randn(S, 50) * .3, rng.randn(S) * .1]
    w = [a.astype(np.float32) for a in w]
    y = (rng.randn(B, T) * 1.5).astype(np.float32)
    dec, pri = mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], n_stages=T - 1, return_priors=True)
    dec, pri = dec.cpu().numpy(), pri.cpu().numpy()
    exact = orc.vnet_priors(y, w, dtype=np.float64)
    assert rel_to_rowmax(pri[:, :T - 1], exact[:, :T - 1]) < PRIOR_RTOL
    ref_own, _ = orc.vnet_decode_from_priors(pri, T - 1)
    assert np.array_equal(dec, ref_own)
    words = mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], n_stages=T - 1, out_format=mvn.OUT_BITS)
    assert np.array_equal(mvn.ops.unpack_bits(words, T).cpu().numpy(), dec)


@pytest.mark.parametrize('L', [4, 7])
def test_vnet_fused_saturating_inputs(mvn, fused_impl, L):
    """Large first-layer weights and outliers in y: sigmoid exponents far beyond fp32 range for some (frame, stage,
    unit) triples.  The tensor-core kernel's producers take their clamp-free path only while |y| is inside the
    launch's safe bound, so this batch mixes both paths (and warps where only one lane is outside); every prior must
    stay finite and inside the tolerance, decisions exact on the kernel's own priors (vnet_detector.py:49-61)."""
    rng = np.random.RandomState(77 + L)
    S, B, T = 2 ** L, 515, 40
    w = [rng.randn(100, 1) * 6.0, rng.randn(100) * 8.0, rng.randn(50, 100) * .15, rng.randn(50) * .1,
         rng.randn(S, 50) * .3, rng.randn(S) * .1]
    w = [a.astype(np.float32) for a in w]
    y = (rng.randn(B, T) * 1.5).astype(np.float32)
    y[::7, ::5] *= 40.0          # single-lane outliers
    y[64:96] = y[64:96] * 0.01   # one warp tile entirely inside the bound
    y[3, 3], y[100, 10] = 3.0e4, -3.0e4
    dec, pri = mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], return_priors=True)
    dec, pri = dec.cpu().numpy(), pri.cpu().numpy()
    assert np.isfinite(pri).all()
    exact = orc.vnet_priors(y, w, dtype=np.float64)
    assert rel_to_rowmax(pri, exact) < PRIOR_RTOL
    ref_own, _ = orc.vnet_decode_from_priors(pri)
    assert np.array_equal(dec, ref_own)


def test_vnet_detector_classes(mvn):
    g = load_golden('vnet')
    w, y = _w(g, 'trained_w'), g['y']
    det = mvn.VNETDetector(16, {'val': y.shape[1], 'train': y.shape[1]})
    assert sorted(det.state_dict().keys()) == ['net.0.bias', 'net.0.weight', 'net.2.bias', 'net.2.weight',
                                               'net.4.bias', 'net.4.weight']
    with torch.no_grad():
        for p, a in zip(det.parameters(), w):
            p.copy_(cu(a))
    out = det(cu(y), 'val')
    assert out.shape == y.shape and out.dtype == torch.float32 and out.is_cuda
    assert (out.cpu().numpy() != g['trained_dec']).any(axis=1).sum() <= 1
    # batch above the small-batch threshold goes through the fused kernel: same bits
    reps = 2048 // y.shape[0] + 1
    big = det(cu(np.tile(y, (reps, 1))), 'val').cpu().numpy()
    assert np.array_equal(big[:y.shape[0]], big[-y.shape[0]:])          # replicas decode identically
    # fused (tensor-core layer 2) vs priors-kernel + ACS-kernel: priors agree to ~1e-7, so only a near-tie can differ
    assert (big[:y.shape[0]] != out.cpu().numpy()).any(axis=1).sum() <= 1
    pri = det(cu(y), 'train')
    assert pri.shape == y.shape + (16,)
    assert rel_to_rowmax(pri.detach().cpu().numpy(), g['trained_priors']) < PRIOR_RTOL
    meta = mvn.META_VNETDetector(16, {'val': y.shape[1], 'train': y.shape[1]})
    out_m = meta(cu(y), 'val', [cu(a) for a in w])
    assert np.array_equal(out_m.cpu().numpy(), out.cpu().numpy())


# ------------------------------------------------------------------------------- a8 / a12
@pytest.mark.parametrize('L', [3, 4, 6, 8])
def test_calculate_states_golden(mvn, L):
    g = load_golden('labels_metrics')
    st = mvn.calculate_states(L, cu(g[f'tx_L{L}']))
    assert st.dtype == torch.int64
    assert np.array_equal(st.cpu().numpy(), g[f'states_L{L}'])


@pytest.mark.parametrize('k', range(4))
def test_error_rates_golden(mvn, k):
    g = load_golden('labels_metrics')
    ber, fer, idx = mvn.calculate_error_rates(cu(g[f'pred_{k}']), cu(g[f'tgt_{k}']))
    assert ber == g[f'ber_fer_{k}'][0] and fer == g[f'ber_fer_{k}'][1]
    assert np.array_equal(idx.cpu().numpy(), g[f'idx_{k}'])


@pytest.mark.parametrize('L', [1, 4, 5, 6, 7, 8])
def test_vnet_fused_edge_shapes_with_counters(mvn, fused_impl, L):
    """no stage at all, one symbol, exact tile multiples, tile + 1, stages ending mid-tile; fused counters exact;
    the tensor-core kernel's pipeline-timeout flag stays clear"""
    import ctypes
    rng = np.random.RandomState(900 + L)
    S = 2 ** L
    w = [(rng.randn(*s) * sc).astype(np.float32) for s, sc in
         [((100, 1), .7), ((100,), .5), ((50, 100), .15), ((50,), .1), ((S, 50), .3), ((S,), .1)]]
    wd = [cu(a) for a in w]
    for B, T, n in ((3, 5, 0), (129, 64, 64), (1, 1, 1), (300, 97, 50), (128, 32, 32), (257, 120, 120)):
        y = (rng.randn(B, T) * 1.5).astype(np.float32)
        tgt = rng.randint(0, 2, size=(B, T)).astype(np.float32)
        cnt = mvn.ops.new_counters()
        dec, pri = mvn.ops.vnet_decode(cu(y), wd, n_stages=n, return_priors=True, target=cu(tgt), counters=cnt)
        dec, pri = dec.cpu().numpy(), pri.cpu().numpy()
        ref = orc.vnet_decode_from_priors(pri, n)[0] if n > 0 else np.zeros_like(dec)
        assert np.array_equal(dec, ref), (L, B, T, n)
        assert cnt.cpu().tolist() == [int((dec != tgt).sum()), int((dec != tgt).any(axis=1).sum()), B * T, B]
    f = mvn._lib.load().mvn_tc_timeout_status
    f.restype = ctypes.c_int
    assert f() == 0


def test_error_counts_pilots_and_fused_counters(mvn, fused_impl):
    g = load_golden('vnet')
    w, y = _w(g, 'trained_w'), g['y']
    rng = np.random.RandomState(3)
    tgt = rng.randint(0, 2, size=(y.shape[0], 120)).astype(np.float32)      # target narrower than y (ECC shape)
    dec = mvn.ops.vnet_decode(cu(y), [cu(a) for a in w])
    be, fe, nb, nf, _ = orc.error_counts(dec.cpu().numpy()[:, :120], tgt)
    cnt, rows = mvn.ops.error_counts(dec[:, :120], cu(tgt))
    assert cnt.tolist() == [be, fe, nb, nf]
    # pilot rows (index % 25 == 0) excluded, trainer.py:100-102
    keep = np.array([i for i in range(y.shape[0]) if i % 25 != 0])
    be, fe, nb, nf, _ = orc.error_counts(dec.cpu().numpy()[keep, :120], tgt[keep])
    cnt2, _ = mvn.ops.error_counts(dec[:, :120], cu(tgt), pilot_period=25)
    assert cnt2.tolist() == [be, fe, nb, nf]
    # fused accumulation inside the decode kernels
    c3 = mvn.ops.new_counters()
    mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], target=cu(tgt), pilot_period=25, counters=c3, want_decoded=False)
    assert c3.tolist() == [be, fe, nb, nf]
    gv = load_golden('va')
    from meta_viterbinet_b200.channel_taps import state_priors_table
    name = 'L4_fade2'
    yv, hv, bv = gv[f'{name}_y'], gv[f'{name}_h'], gv[f'{name}_b']
    c4 = mvn.ops.new_counters()
    dv = mvn.ops.va_decode(cu(yv), cu(state_priors_table(hv, 4)), target=cu(bv), counters=c4)
    be, fe, nb, nf, _ = orc.error_counts(dv.cpu().numpy(), bv)
    assert c4.tolist() == [be, fe, nb, nf]


# ------------------------------------------------------------------------------- host-buffer pipeline
def test_host_pipeline_matches_device_path(mvn):
    import ctypes
    from meta_viterbinet_b200 import _lib
    g = load_golden('vnet')
    w, y = _w(g, 'trained_w'), g['y']
    yy = np.ascontiguousarray(np.tile(y, (7, 1)))           # 350 frames, chunk 128 -> 3 chunks, ragged tail
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    _lib.check(lib.mvn_ctx_create(ctypes.byref(ctx), torch.cuda.current_device(), 128, yy.shape[1], 4))
    try:
        ws = [np.ascontiguousarray(a, dtype=np.float32) for a in w]
        _lib.check(lib.mvn_ctx_set_vnet_weights_host(ctx, *[a.ctypes.data_as(ctypes.c_void_p) for a in ws]))
        out = np.empty_like(yy)
        _lib.check(lib.mvn_ctx_vnet_decode_host(ctx, yy.ctypes.data_as(ctypes.c_void_p), yy.shape[0], yy.shape[1],
                                                yy.shape[1], 0, out.ctypes.data_as(ctypes.c_void_p)))
        dev = mvn.ops.vnet_decode(cu(yy), [cu(a) for a in w]).cpu().numpy()
        assert np.array_equal(out, dev)
        gv = load_golden('va')
        from meta_viterbinet_b200.channel_taps import state_priors_table
        name = 'L4_fade1_ecc'
        yv = np.ascontiguousarray(gv[f'{name}_y'])
        tab = state_priors_table(gv[f'{name}_h'], 4)
        ctx2 = ctypes.c_void_p()
        _lib.check(lib.mvn_ctx_create(ctypes.byref(ctx2), torch.cuda.current_device(), 50, yv.shape[1], 4))
        outv = np.empty_like(yv)
        _lib.check(lib.mvn_ctx_va_decode_host(ctx2, yv.ctypes.data_as(ctypes.c_void_p), yv.shape[0], yv.shape[1],
                                              yv.shape[1], tab.ctypes.data_as(ctypes.c_void_p), tab.shape[0], 0,
                                              outv.ctypes.data_as(ctypes.c_void_p)))
        lib.mvn_ctx_destroy(ctx2)
        assert np.array_equal(outv, gv[f'{name}_dec'])
    finally:
        lib.mvn_ctx_destroy(ctx)


# ------------------------------------------------------------------------------- full size
def test_full_size_properties(mvn, fused_impl):
    """BASELINE.json sizes (1M frames x 120): frames are independent, so a batch built from 4096
    distinct frames repeated in shuffled order must decode every copy exactly like the oracle decodes
    the distinct frame — for the VA (bit-exact) and for the fused ViterbiNet (own-priors protocol)."""
    from meta_viterbinet_b200.channel_taps import state_priors_table
    rng = np.random.RandomState(42)
    L, T, U, B = 4, 120, 4096, 1 << 20
    h = np.exp(-0.2 * np.arange(L)).reshape(1, L)
    bits = rng.randint(0, 2, size=(U, T))
    y_u = orc.isi_awgn(bits, h, 10.0, L, rng).astype(np.float32)
    perm = torch.randint(0, U, (B,), generator=torch.Generator().manual_seed(1))
    y = cu(y_u)[perm.cuda()].contiguous()
    ref = torch.as_tensor(orc.va_decode(y_u, h, L, T)).cuda()
    cnt = mvn.ops.new_counters()
    tgt = cu(bits.astype(np.float32))[perm.cuda()].contiguous()
    dec = mvn.ops.va_decode(y, cu(state_priors_table(h, L)), target=tgt, counters=cnt)
    assert torch.equal(dec, ref[perm.cuda()])
    be_u = (ref.cpu().numpy() != bits).sum(axis=1)
    assert cnt.tolist()[0] == int(be_u[perm.numpy()].sum()) and cnt.tolist()[3] == B
    g = load_golden('vnet')
    w = [cu(a) for a in _w(g, 'trained_w')]
    dec_u, pri_u = mvn.ops.vnet_decode(cu(y_u), w, return_priors=True)
    own, _ = orc.vnet_decode_from_priors(pri_u.cpu().numpy())
    assert np.array_equal(dec_u.cpu().numpy(), own)
    dec_big = mvn.ops.vnet_decode(y, w)
    assert torch.equal(dec_big, dec_u[perm.cuda()])
    words = mvn.ops.vnet_decode(y, w, out_format=mvn.OUT_BITS)
    assert torch.equal(mvn.ops.unpack_bits(words[:5000], T), dec_big[:5000])


# ------------------------------------------------------------------------------- protocol (ii) at real sizes
def test_vnet_4100_frames_against_reference_forward(mvn, fused_impl):
    """4 100 words decoded by the reference's full VNETDetector.forward (tests/golden/vnet4096.npz, reference-trained
    weights): every frame a kernel variant decodes differently must be explained as a near-tie under fp64 priors."""
    g, v = load_golden('vnet4096'), load_golden('vnet')
    w, y, T = _w(v, 'trained_w'), g['y'], int(g['T'][0])
    ref = unpack_rows(g['dec_packed'], T)
    dec, pri = mvn.ops.vnet_decode(cu(y), [cu(a) for a in w], return_priors=True)
    dec, pri = dec.cpu().numpy(), pri.cpu().numpy()
    exact = orc.vnet_priors(y, w, dtype=np.float64)
    assert rel_to_rowmax(pri, exact) < PRIOR_RTOL
    own, _ = orc.vnet_decode_from_priors(pri)
    assert np.array_equal(dec, own)                                                   # protocol (i)
    tol = PRIOR_RTOL * np.abs(exact).max(axis=(1, 2))
    assert _explain_mismatches(dec, ref, exact, tol) <= 2                             # protocol (ii)


def test_tensor_core_vs_fma_disagreements_are_near_ties(mvn):
    """2^18 frames with an UNTRAINED net (small margins, the worst case: tools/stress_tc.py saw 244 differing frames per
    2^20): every frame on which the tcgen05 kernel and the FP32-FMA kernel disagree must be a near-tie under fp64
    priors, and both must follow their own priors exactly on those frames."""
    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Linear(1, 100), torch.nn.Sigmoid(), torch.nn.Linear(100, 50), torch.nn.ReLU(),
                              torch.nn.Linear(50, 16))
    w = [p.detach().cuda().contiguous() for p in net.parameters()]
    w_np = [a.cpu().numpy() for a in w]
    rng = np.random.RandomState(11)
    B, T, L = 1 << 18, 120, 4
    bits = torch.randint(0, 2, (B, T), generator=torch.Generator().manual_seed(3)).float()
    h = np.exp(-0.2 * np.arange(L)).reshape(1, L)
    y = mvn.ops.channel_transmit(bits.cuda(), h, 9.0, seed=17)
    out = {}
    for name in ('tcgen05', 'fma'):
        mvn.ops.set_fused_variant(name)
        try:
            out[name] = mvn.ops.vnet_decode(y, w)
        finally:
            mvn.ops.set_fused_variant('auto')
    bad = torch.nonzero((out['tcgen05'] != out['fma']).any(dim=1)).reshape(-1).cpu().numpy()
    assert len(bad) < B // 1000
    if len(bad):
        yb = y[bad].cpu().numpy()
        exact = orc.vnet_priors(yb, w_np, dtype=np.float64)
        tol = PRIOR_RTOL * np.abs(exact).max(axis=(1, 2))
        n = _explain_mismatches(out['tcgen05'][bad].cpu().numpy(), out['fma'][bad].cpu().numpy(), exact, tol)
        assert n == len(bad)
        for name in ('tcgen05', 'fma'):           # each variant is exact on its own priors for exactly these frames
            mvn.ops.set_fused_variant(name)
            try:
                d, p = mvn.ops.vnet_decode(cu(yb), w, return_priors=True)
            finally:
                mvn.ops.set_fused_variant('auto')
            assert np.array_equal(d.cpu().numpy(), orc.vnet_decode_from_priors(p.cpu().numpy())[0])
            assert np.array_equal(d.cpu().numpy(), out[name][bad].cpu().numpy())      # batch position does not matter


@pytest.mark.parametrize('tag', ['raw', 'ecc'])
def test_config1_300_blocks(mvn, tag):
    """BASELINE.json configs[0] at its real size through the drop-in classes: 300 blocks (val_frames=12), fading taps,
    10 dB; VA bit-exact, ViterbiNet by protocol (ii), the reference's SER / FER on the data rows (RS-decoded on the GPU
    when coded) reproduced as identical floats."""
    g, v = load_golden('config1'), load_golden('vnet')
    y, b = g[f'{tag}_y'], g[f'{tag}_b'].astype(np.float32)
    T = y.shape[1]
    va = mvn.VADetector(16, 4, T, 300, 'ISI_AWGN', 0, True, 1, {'train': 'time_decay', 'val': 'time_decay'})
    dec_va = va(cu(y), 'val', 10.0, 0.2)
    assert np.array_equal(dec_va.cpu().numpy(), unpack_rows(g[f'{tag}_dec_va'], T))
    det = mvn.VNETDetector(16, {'val': T, 'train': T})
    with torch.no_grad():
        for p, a in zip(det.parameters(), _w(v, 'trained_w')):
            p.copy_(cu(a))
    dec_vn = det(cu(y), 'val')
    ref_vn = unpack_rows(g[f'{tag}_dec_vnet'], T)
    exact = orc.vnet_priors(y, _w(v, 'trained_w'), dtype=np.float64)
    assert _explain_mismatches(dec_vn.cpu().numpy(), ref_vn, exact, PRIOR_RTOL * np.abs(exact).max(axis=(1, 2))) <= 1
    assert rel_to_rowmax(det(cu(y[:25]), 'train').detach().cpu().numpy(), g[f'{tag}_priors25']) < PRIOR_RTOL
    rows = torch.as_tensor(g[f'{tag}_data_indices']).cuda()
    for nm, d in (('va', dec_va), ('vnet', cu(ref_vn))):
        msg = mvn.ecc.decode(d, 2) if tag == 'ecc' else d
        ber, fer, idx = mvn.calculate_error_rates(msg[rows], cu(b)[rows])
        assert ber == g[f'{tag}_rates_{nm}'][0] and fer == g[f'{tag}_rates_{nm}'][1]
        assert np.array_equal(idx.cpu().numpy(), g[f'{tag}_erridx_{nm}'])


# ------------------------------------------------------------------------------- shared state / sweeps
def test_constant_slots_are_safe_across_streams(mvn, fused_impl):
    """The fused kernel keeps its weights in two constant-bank slots; interleaved calls with DIFFERENT
    weights on different streams must not see each other's weights."""
    rng = np.random.RandomState(5)
    y = cu((rng.randn(3000, 64) * 1.5).astype(np.float32))
    sets = []
    for k in range(3):
        w = [rng.randn(100, 1) * .7, rng.randn(100) * .5, rng.randn(50, 100) * .15, rng.randn(50) * .1,
             rng.randn(16, 50) * .3, rng.randn(16) * .1]
        sets.append([cu(a.astype(np.float32)) for a in w])
    ref = [mvn.ops.vnet_decode(y, w) for w in sets]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    for rep in range(4):
        for k, st in enumerate(streams):
            with torch.cuda.stream(st):
                outs.append((k, mvn.ops.vnet_decode(y, sets[k])))
    torch.cuda.synchronize()
    for k, o in outs:
        assert torch.equal(o, ref[k])


def test_sweep_on_gpu_counts(mvn):
    g = load_golden('vnet')
    w = [cu(a) for a in _w(g, 'trained_w')]
    rng = np.random.RandomState(9)
    data = {}
    for snr in (8.0, 10.0):
        bits = rng.randint(0, 2, size=(300, 120))
        yy = orc.isi_awgn(bits, np.exp(-0.2 * np.arange(4)).reshape(1, 4), snr, 4, rng).astype(np.float32)
        data[snr] = (cu(yy), cu(bits.astype(np.float32)))

    def block(snr, first, n, row):
        yy, bb = data[snr]
        mvn.ops.vnet_decode(yy[first:first + n].contiguous(), w, target=bb[first:first + n].contiguous(), counters=row,
                            want_decoded=False)

    total = mvn.sweep.run_sweep([8.0, 10.0], 300, block, device='cuda')
    for i, snr in enumerate((8.0, 10.0)):
        yy, bb = data[snr]
        dec = mvn.ops.vnet_decode(yy, w).cpu().numpy()
        be, fe, nb, nf, _ = orc.error_counts(dec, bb.cpu().numpy())
        assert total[i].tolist() == [be, fe, nb, nf]
    # splitting a point into row blocks (more ranks than points) gives the same totals
    parts = torch.zeros(2, 4, dtype=torch.int64, device='cuda')
    for r in range(4):
        c = mvn.sweep.run_sweep([8.0, 10.0], 300, block, device='cuda', rank=r, world_size=4)
        parts += c
    assert torch.equal(parts, total)


# ------------------------------------------------------------------------------- SURVEY.md §8(f) next rows
@pytest.mark.parametrize('name', ['static_ecc', 'fade1', 'fade2_ecc', 'cost2100', 'L6_static'])
def test_channel_transmit_matches_reference_draw(mvn, name):
    """f1 pinned: mvn_channel_transmit fed the recorded RandomState(noise_seed) stream reproduces the words a
    ChannelModelDataset draw of the live reference produced (tests/golden/channel.npz, make_golden_r2.py) bit for bit."""
    g = load_golden('channel')
    L, T, snr = int(g[f'{name}_meta'][0]), int(g[f'{name}_meta'][1]), float(g[f'{name}_meta'][2])
    y = mvn.ops.channel_transmit(cu(g[f'{name}_c'].astype(np.float32)), g[f'{name}_h'], snr, noise=g[f'{name}_noise'])
    assert np.array_equal(y.cpu().numpy().view(np.uint32), g[f'{name}_y'].view(np.uint32))


@pytest.mark.parametrize('L', [1, 2, 4, 6, 8])
def test_channel_transmit_matches_oracle(mvn, L):
    """other memory lengths / per-word random taps: against the oracle (itself pinned on the reference draw above)"""
    rng = np.random.RandomState(L)
    B, T = 37, 136
    bits = rng.randint(0, 2, size=(B, T))
    h = np.abs(rng.randn(B, L)) + 0.1                      # one tap vector per word (fading)
    noise = np.random.RandomState(100 + L).normal(0, 1, (B, T))
    ref = orc.isi_awgn(bits, h, 9.0, L, np.random.RandomState(100 + L)).astype(np.float32)
    y = mvn.ops.channel_transmit(cu(bits.astype(np.float32)), h, 9.0, noise=noise)
    assert np.array_equal(y.cpu().numpy(), ref)
    h1 = np.exp(-0.2 * np.arange(L)).reshape(1, L)         # static channel (one tap row)
    ref2 = orc.isi_awgn(bits, h1, 10.0, L, np.random.RandomState(100 + L)).astype(np.float32)
    y2 = mvn.ops.channel_transmit(cu(bits.astype(np.float32)), h1, 10.0, noise=noise)
    assert np.array_equal(y2.cpu().numpy(), ref2)


def test_channel_transmit_device_noise_statistics(mvn):
    B, T, L, snr = 4096, 120, 4, 6.0
    bits = torch.zeros(B, T).cuda()
    h = np.array([[1.0, 0.0, 0.0, 0.0]])
    y = mvn.ops.channel_transmit(bits, h, snr, seed=123)
    n = (y - 1.0).double()                                 # all-zero bits -> symbol +1 on the main tap
    sigma = 10 ** (-snr / 20)
    assert abs(float(n.mean())) < 5 * sigma / np.sqrt(B * T)
    assert abs(float(n.std()) / sigma - 1) < 0.01
    y2 = mvn.ops.channel_transmit(bits, h, snr, seed=123)
    assert torch.equal(y, y2)                              # reproducible for a seed
    assert not torch.equal(y, mvn.ops.channel_transmit(bits, h, snr, seed=124))


@pytest.mark.parametrize('L', [1, 3, 4, 5, 6, 8])
def test_mlse_traceback_vs_oracle(mvn, L):
    rng = np.random.RandomState(L)
    S, B, T = 2 ** L, 45, 50
    cost = (rng.randn(B, T, S) * 2).astype(np.float32)
    cost[::4] = rng.randint(0, 3, size=cost[::4].shape)
    for n_stages, start in ((T, -1), (T - 7, -1), (T, 0)):
        ref, _, _ = orc.mlse_decode(cost, n_stages, start)
        dec = mvn.ops.mlse_decode(cu(cost), n_stages, terminated=(start == 0))
        assert np.array_equal(dec.cpu().numpy(), ref)
    words = mvn.ops.mlse_decode(cu(cost), out_format=mvn.OUT_BITS)
    assert np.array_equal(mvn.ops.unpack_bits(words, T).cpu().numpy(), orc.mlse_decode(cost)[0])


@pytest.mark.parametrize('L', range(1, 9))
def test_va_fused_mlse_vs_oracle(mvn, L):
    """decision = MLSE inside mvn_va_decode_ex (survivor masks in shared memory, in-kernel traceback): same bits as the
    oracle's traceback over the reference's acs_block survivors (trellis_utils.py:30), both start rules, loops that end
    mid-tile, bit-packed output, fused counters."""
    from meta_viterbinet_b200.channel_taps import state_priors_table
    rng = np.random.RandomState(50 + L)
    B, T = 77, 70
    h = (np.exp(-0.3 * np.arange(L)) * (1 + 0.1 * rng.randn(7, L))).astype(np.float64)     # 7 tap blocks, B % 7 == 0
    bits = rng.randint(0, 2, size=(B, T))
    y = orc.isi_awgn(bits, h[np.arange(B) % 7], 6.0, L, rng).astype(np.float32)
    y[:5] = np.round(y[:5])                                                               # coarse values: exact ties
    table = state_priors_table(h, L)
    cost = orc.va_cost(y, orc.va_state_priors(h, L))
    for n_stages, term in ((T, False), (T, True), (T - 9, False), (33, True), (1, False), (0, True)):
        ref = orc.mlse_decode(cost, n_stages, 0 if term else -1)[0]
        cnt = mvn.ops.new_counters()
        tgt = cu(bits.astype(np.float32))
        dec = mvn.ops.va_decode(cu(y), cu(table), n_stages, decision='mlse_terminated' if term else 'mlse', target=tgt,
                                pilot_period=5, counters=cnt).cpu().numpy()
        assert np.array_equal(dec, ref), (L, n_stages, term)
        keep = np.arange(B) % 5 != 0
        assert cnt.tolist() == [int((ref[keep] != bits[keep]).sum()), int((ref[keep] != bits[keep]).any(axis=1).sum()),
                                int(keep.sum()) * T, int(keep.sum())]
    words = mvn.ops.va_decode(cu(y), cu(table), decision='mlse', out_format=mvn.OUT_BITS)
    assert np.array_equal(mvn.ops.unpack_bits(words, T).cpu().numpy(), orc.mlse_decode(cost)[0])
    # per stage, the survivors the traceback walks are the reference's acs_block indices (checked through mvn_acs_decode)
    _, _, surv = mvn.ops.acs_decode(cu(cost), return_final_pm=True, return_survivors=True)
    H = max(1, 2 ** L // 2)
    assert np.array_equal(unpack_survivors(surv, H), orc.acs_decode(cost, return_survivors=True)[2][:, :, :H])


@pytest.mark.parametrize('L', range(1, 9))
def test_vnet_fused_mlse_vs_oracle(mvn, L):
    """decision = MLSE inside the fused ViterbiNet kernel (tensor-core kernel, memory_length <= 6: survivor masks of the
    consumer warps in shared memory, in-kernel traceback; 128 / 256 states: three launches): the bits are exactly the
    oracle's traceback over the kernel's own exported priors, both start rules, ragged loops, packed output, counters."""
    rng = np.random.RandomState(70 + L)
    S = 2 ** L
    w = [(rng.randn(*s) * sc).astype(np.float32) for s, sc in
         [((100, 1), .7), ((100,), .5), ((50, 100), .15), ((50,), .1), ((S, 50), .3), ((S,), .1)]]
    wd = [cu(a) for a in w]
    for B, T, n_stages, term in ((130, 70, 70, False), (130, 70, 61, True), (257, 33, 33, True), (5, 40, 1, False), (64, 32, 0, True)):
        y = (rng.randn(B, T) * 1.5).astype(np.float32)
        tgt = rng.randint(0, 2, size=(B, T)).astype(np.float32)
        cnt = mvn.ops.new_counters()
        dec, pri = mvn.ops.vnet_decode(cu(y), wd, n_stages=n_stages, return_priors=True, target=cu(tgt), counters=cnt,
                                       decision='mlse_terminated' if term else 'mlse')
        dec, pri = dec.cpu().numpy(), pri.cpu().numpy()
        ref = orc.mlse_decode(-pri, n_stages, 0 if term else -1)[0]
        assert np.array_equal(dec, ref), (L, B, T, n_stages, term)
        assert cnt.tolist() == [int((ref != tgt).sum()), int((ref != tgt).any(axis=1).sum()), B * T, B]
    words = mvn.ops.vnet_decode(cu(y), wd, out_format=mvn.OUT_BITS, decision='mlse')
    dec = mvn.ops.vnet_decode(cu(y), wd, decision='mlse')
    assert torch.equal(mvn.ops.unpack_bits(words, T), dec)
    f = mvn._lib.load().mvn_tc_timeout_status
    assert f() == 0


def test_va_fused_mlse_full_size(mvn):
    """2^20 frames x 120: replicas of 4096 oracle-decoded frames in shuffled order decode identically (one launch)"""
    from meta_viterbinet_b200.channel_taps import state_priors_table
    rng = np.random.RandomState(8)
    L, T, U, B = 4, 120, 4096, 1 << 20
    h = np.exp(-0.2 * np.arange(L)).reshape(1, L)
    bits = rng.randint(0, 2, size=(U, T))
    y_u = orc.isi_awgn(bits, h, 8.0, L, rng).astype(np.float32)
    ref = torch.as_tensor(orc.mlse_decode(orc.va_cost(y_u, orc.va_state_priors(h, L)), start_state=0)[0]).cuda()
    perm = torch.randint(0, U, (B,), generator=torch.Generator().manual_seed(2)).cuda()
    n0 = mvn._lib.launch_count(reset=True)
    dec = mvn.ops.va_mlse_decode(cu(y_u)[perm].contiguous(), cu(state_priors_table(h, L)), terminated=True)
    assert mvn._lib.launch_count() == 1
    assert torch.equal(dec, ref[perm])


def test_va_mlse_beats_the_reference_rule(mvn):
    """Traceback is strictly better than the reference's running-argmin rule on the same channel."""
    from meta_viterbinet_b200.channel_taps import state_priors_table
    rng = np.random.RandomState(3)
    L, B, T = 4, 4000, 120
    h = np.exp(-0.2 * np.arange(L)).reshape(1, L)
    bits = rng.randint(0, 2, size=(B, T))
    y = mvn.ops.channel_transmit(cu(bits.astype(np.float32)), h, 8.0, seed=5)
    table = cu(state_priors_table(h, L))
    ref_rule = mvn.ops.va_decode(y, table).cpu().numpy()
    mlse = mvn.ops.va_mlse_decode(y, table, terminated=True).cpu().numpy()
    cost = orc.va_cost(y.cpu().numpy(), orc.va_state_priors(h, L))
    assert np.array_equal(mlse, orc.mlse_decode(cost, start_state=0)[0])
    ber_rule = (ref_rule != bits)[:, 1:].mean()
    ber_mlse = (mlse != bits).mean()
    assert ber_mlse < 0.5 * ber_rule

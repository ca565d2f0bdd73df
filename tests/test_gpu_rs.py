"""GPU parity of the Reed-Solomon kernels: bit-exact against the fixtures recorded from the
reference's ecc/rs_main.py and against the oracle on larger random batches, through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rs_oracle as rs

pytestmark = pytest.mark.gpu
G = load_golden('rs')
CONFIGS = [tuple(int(v) for v in c) for c in G['configs']]


@pytest.fixture(scope='module')
def mvn():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import meta_viterbinet_b200 as m
    return m


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('k,nsym', CONFIGS)
def test_encode_decode_match_reference_fixtures(mvn, k, nsym):
    tag = f'k{k}_n{nsym}'
    tx = mvn.ops.rs_encode(cu(G[f'{tag}_msg']), nsym).cpu().numpy()
    assert np.array_equal(tx, G[f'{tag}_tx'])
    dec, st = mvn.ops.rs_decode(cu(G[f'{tag}_rx']), nsym, return_status=True)
    assert np.array_equal(dec.cpu().numpy(), G[f'{tag}_dec'])
    ref_st = [rs.decode(r, nsym, return_status=True)[1] for r in G[f'{tag}_rx']]
    assert st.cpu().tolist() == ref_st


@pytest.mark.parametrize('k,nsym,B', [(15, 2, 3000), (15, 4, 3000), (60, 8, 1500), (223, 32, 300), (100, 17, 400)])
def test_random_batches_match_oracle(mvn, k, nsym, B):
    rng = np.random.RandomState(100 * k + nsym)
    n = k + nsym
    msg = rng.randint(0, 2, size=(B, 8 * k)).astype(np.float32)
    tx = mvn.ops.rs_encode(cu(msg), nsym)
    rx = tx.cpu().numpy().astype(np.uint8)
    n_err = rng.randint(0, nsym // 2 + 3, size=B)
    for w in range(B):
        for p in rng.choice(n, size=n_err[w], replace=False):
            rx[w, 8 * p:8 * p + 8] ^= np.unpackbits(np.array([rng.randint(1, 256)], dtype=np.uint8))
    dec, st = mvn.ops.rs_decode(cu(rx.astype(np.float32)), nsym, return_status=True)
    dec, st = dec.cpu().numpy(), st.cpu().numpy()
    # every word within capacity is recovered (property, all rows) ...
    ok = n_err <= nsym // 2
    assert np.array_equal(dec[ok], msg[ok])
    assert np.all(st[ok] == (n_err[ok] > 0))
    # ... and a sample of rows, weighted towards the uncorrectable ones, is bit-exact with the oracle
    rows = np.concatenate([np.nonzero(~ok)[0][:120], np.nonzero(ok)[0][:40]])
    for w in rows:
        o, s = rs.decode(rx[w], nsym, return_status=True)
        assert np.array_equal(dec[w], o) and st[w] == s, (w, n_err[w], s, st[w])
        assert np.array_equal(tx[w].cpu().numpy(), rs.encode(msg[w].astype(np.uint8), nsym))


def test_reference_signatures_and_edge_cases(mvn):
    from meta_viterbinet_b200 import ecc
    k, nsym = 15, 4
    tag = f'k{k}_n{nsym}'
    word = G[f'{tag}_msg'][3].astype(int)
    cw = ecc.encode(word, nsym)                        # numpy in -> numpy out, one word (rs_main.py:9)
    assert isinstance(cw, np.ndarray) and np.array_equal(cw, G[f'{tag}_tx'][3])
    assert np.array_equal(ecc.decode(G[f'{tag}_rx'][5].astype(int), nsym), G[f'{tag}_dec'][5])
    assert ecc.decode(cu(G[f'{tag}_rx'].astype(np.float32)), nsym).is_cuda
    # empty batch, too long a message (the reference raises ValueError('Message is too long ...')), bad nsym
    assert mvn.ops.rs_decode(torch.zeros(0, 8 * 19).cuda(), nsym).shape == (0, 8 * 15)
    with pytest.raises(mvn.MVNError, match='too long'):
        mvn.ops.rs_encode(torch.zeros(2, 8 * 250).cuda(), 8)
    with pytest.raises(mvn.MVNError):
        mvn.ops.rs_decode(torch.zeros(2, 8 * 19).cuda(), 33)
    with pytest.raises(mvn.MVNError):
        mvn.ops.rs_decode(torch.zeros(2, 8 * 4).cuda(), 4)    # no message bytes left


def test_unaligned_rows_through_the_c_abi(mvn):
    """rows that are not 16-byte aligned take the scalar load/store path"""
    from meta_viterbinet_b200 import _lib
    lib = _lib.load()
    k, nsym = 15, 4
    tag = f'k{k}_n{nsym}'
    rx = G[f'{tag}_rx'].astype(np.float32)
    B, nb = rx.shape
    ld_in, ld_out = nb + 3, 8 * k + 1
    buf = torch.zeros(B * ld_in + 1, device='cuda')
    buf[1:].view(B, ld_in)[:, :nb] = cu(rx)
    out = torch.full((B * ld_out + 1,), 7.0, device='cuda')
    rc = lib.mvn_rs_decode(buf.data_ptr() + 4, B, ld_in, k + nsym, nsym, out.data_ptr() + 4, ld_out, None,
                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    got = out[1:].view(B, ld_out).cpu().numpy()
    assert np.array_equal(got[:, :8 * k], G[f'{tag}_dec']) and np.all(got[:, 8 * k:] == 7.0)


def test_detector_to_ber_chain_with_ecc(mvn):
    """coded evaluation as in trainer.py:226-240: encode -> channel -> VA detect -> RS decode -> error rates,
    all on the device; the decoded words equal the oracle's decode of the detector output."""
    rng = np.random.RandomState(7)
    k, nsym, L, B = 15, 4, 4, 512
    msg = rng.randint(0, 2, size=(B, 8 * k)).astype(np.float32)
    cw = mvn.ops.rs_encode(cu(msg), nsym)
    taps = np.exp(-0.2 * np.arange(L)).reshape(1, L)
    y = mvn.ops.channel_transmit(cw, taps, snr_db=9.0, seed=11)
    det = mvn.VADetector(2 ** L, L, cw.shape[1], B, 'ISI_AWGN', 0, False, 1, {'train': 'time_decay', 'val': 'time_decay'})
    detected = det(y, 'val', snr=9.0, gamma=0.2)
    dec = mvn.ops.rs_decode(detected, nsym)
    d_np = detected.cpu().numpy().astype(np.uint8)
    for w in range(0, B, 16):
        assert np.array_equal(dec[w].cpu().numpy(), rs.decode(d_np[w], nsym))
    raw_ber = float((detected[:, :8 * k].cpu().numpy() != msg).mean())
    coded_ber = float((dec.cpu().numpy() != msg).mean())
    assert coded_ber <= raw_ber

"""GPU: the batched word-by-word online evaluation (online.eval_by_word) replays two runs recorded
from the reference's METAVNETTrainer.eval_by_word (tests/golden/make_golden_online.py): same SER per word, same
weights after every block that passed the SER gate.  Both runs advance in lock step as R = 2 realisations with
different weights, words and gate decisions — which is exactly what the masks have to get right.

Tolerances: SER per word |d| <= 1e-7 (a count / 120 in fp32); weights |d| <= 2e-5 absolute after up to 5 x 4 Adam steps
(fp32, different summation order; one Adam step moves a weight by ~1e-3)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mvn():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import meta_viterbinet_b200 as m
    return m


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def pack(g, tag):
    return np.concatenate([g[f'{tag}_w0_{i}'].reshape(-1) for i in range(6)]).astype(np.float32)


def run(mvn, g, tags):
    L, nsym, iters, thresh, lr = g[f'{tags[0]}_cfg']
    tr = mvn.BatchedVNetTrainer(cu(np.stack([pack(g, t) for t in tags])), int(L), lr=float(lr))
    info = cu(np.stack([g[f'{t}_bits'] for t in tags]).astype(np.float32))
    rx = cu(np.stack([g[f'{t}_y'] for t in tags]))
    after = [[] for _ in tags]

    def on_block(c, ser, gate):
        for r, ok in enumerate(gate.cpu().tolist()):
            if ok:
                after[r].append(tr.theta[r].cpu().numpy().copy())
    ser = mvn.online.eval_by_word(tr, info, rx, int(nsym), float(thresh), subframes_in_frame=4, self_supervised=True,
                                  iterations=int(iters), restart_from_saved=True, on_block=on_block)
    return ser.cpu().numpy(), after


@pytest.mark.parametrize('tags', [('a',), ('b',), ('a', 'b'), ('b', 'a', 'b')])
def test_online_evaluation_replays_reference_runs(mvn, tags):
    g = load_golden('online')
    ser, after = run(mvn, g, tags)
    for r, t in enumerate(tags):
        assert np.max(np.abs(ser[r] - g[f'{t}_ser'])) < 1e-7, (t, ser[r], g[f'{t}_ser'])
        ref = g[f'{t}_theta_after']
        assert len(after[r]) == len(ref)
        for got, want in zip(after[r], ref):
            assert np.max(np.abs(got - want)) < 2e-5
        pilots = [c for c in range(ser.shape[1]) if c not in set(g[f'{t}_data_indices'].tolist())]
        assert np.all(ser[r, pilots] == 0)


def test_online_evaluation_without_training_is_plain_detection(mvn):
    g = load_golden('online')
    tr = mvn.BatchedVNetTrainer(cu(pack(g, 'a')[None]), 4)
    theta0 = tr.theta.clone()
    info, rx = cu(g['a_bits'][None].astype(np.float32)), cu(g['a_y'][None])
    ser = mvn.online.eval_by_word(tr, info, rx, 2, 0.02, subframes_in_frame=4, self_supervised=False).cpu().numpy()[0]
    assert torch.equal(tr.theta, theta0)
    w = [cu(g[f'a_w0_{i}']) for i in range(6)]
    for c in range(rx.shape[1]):
        if c % 4:
            dec = mvn.ecc.decode(mvn.ops.vnet_decode(rx[:, c], w), 2)
            assert abs(ser[c] - float((dec != info[:, c]).float().mean())) < 1e-7
    with pytest.raises(ValueError):
        mvn.online.eval_by_word(tr, info[:, :, :-8], rx, 2, 0.02)

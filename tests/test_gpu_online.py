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


def run_meta(mvn, g, tags, draw=None):
    t0 = tags[0]
    L, nsym, iters, thresh, lr = g[f'{t0}_cfg']
    _, meta_sub, meta_iters, j_num, window, maml, meta_lr = g[f'{t0}_meta_cfg']
    tr = mvn.BatchedVNetTrainer(cu(np.stack([pack(g, t) for t in tags])), int(L), lr=float(lr), meta_lr=float(meta_lr))
    info = cu(np.stack([g[f'{t}_bits'] for t in tags]).astype(np.float32))
    rx = cu(np.stack([g[f'{t}_y'] for t in tags]))
    after, meta_after = [[] for _ in tags], [[] for _ in tags]

    def on_block(c, ser, gate):
        for r, ok in enumerate(gate.cpu().tolist()):
            if ok:
                after[r].append(tr.theta[r].cpu().numpy().copy())

    def on_meta_step(active):
        for r, ok in enumerate(active.cpu().tolist()):
            if ok:
                meta_after[r].append(tr.theta[r].cpu().numpy().copy())
    ser = mvn.online.eval_by_word(tr, info, rx, int(nsym), float(thresh), subframes_in_frame=4, self_supervised=True,
                                  iterations=int(iters), restart_from_saved=True, on_block=on_block, online_meta=True,
                                  meta_subframes=int(meta_sub), meta_train_iterations=int(meta_iters), meta_j_num=int(j_num),
                                  window_size=int(window), second_order=bool(maml), weights_init='last_frame', draw=draw,
                                  on_meta_step=on_meta_step)
    return ser.cpu().numpy(), after, meta_after


def check_meta(g, tags, ser, after, meta_after):
    for r, t in enumerate(tags):
        assert np.max(np.abs(ser[r] - g[f'{t}_ser'])) < 1e-7, (t, ser[r], g[f'{t}_ser'])
        for got_list, ref in ((meta_after[r], g[f'{t}_theta_meta']), (after[r], g[f'{t}_theta_after'])):
            assert len(got_list) == len(ref)
            for got, want in zip(got_list, ref):
                assert np.max(np.abs(got - want)) < 3e-5


@pytest.mark.parametrize('tag', ['m', 'n'])
def test_online_meta_training_replays_reference_run(mvn, tag):
    """online_meta on (MAML for 'm', FO-MAML for 'n'): with the reference's seed the default draw reproduces its
    torch.randint stream, so query indices, every meta step, every online-training block and the SER agree."""
    g = load_golden('online')
    torch.manual_seed(int(g[f'{tag}_seed'][0]))
    drawn = []

    def draw(run, high, count):
        out = mvn.online._default_draw(run, high, count)
        drawn.extend(out)
        return out
    ser, after, meta_after = run_meta(mvn, g, (tag,), draw)
    assert drawn == g[f'{tag}_jhat'].tolist()
    check_meta(g, (tag,), ser, after, meta_after)


def test_online_meta_training_in_lock_step(mvn):
    """two copies of the recorded run side by side, each fed its recorded query indices: ragged buffers, masks"""
    g = load_golden('online')
    seq = {0: g['m_jhat'].tolist(), 1: g['m_jhat'].tolist()}
    _, _, meta_iters, j_num, *_ = g['m_meta_cfg']

    # the recorded list is flat: rebuild the per-round draws (rounds of unique indices) by replaying the generator once
    torch.manual_seed(int(g['m_seed'][0]))
    rounds = []

    def recording_draw(run, high, count):
        out = torch.unique(torch.randint(low=0, high=int(high), size=[int(count)])).tolist()
        rounds.append(out)
        return out
    run_meta(mvn, g, ('m',), recording_draw)
    assert [j for r in rounds for j in r] == seq[0]
    per_run = {0: list(rounds), 1: list(rounds)}
    ser, after, meta_after = run_meta(mvn, g, ('m', 'm'), lambda run, high, count: per_run[run].pop(0))
    check_meta(g, ('m', 'm'), ser, after, meta_after)


def test_online_meta_with_initial_sliding_buffer_and_wider_window(mvn):
    """buffer_empty: False (a filled buffer that slides) and window_size = 2: no reference recording for this
    configuration, so this checks the mechanics — buffer length stays constant, support indices wrap, weights move."""
    g = load_golden('online')
    L, nsym, iters, thresh, lr = g['a_cfg']
    tr = mvn.BatchedVNetTrainer(cu(np.stack([pack(g, 'a'), pack(g, 'b')])), int(L), lr=float(lr), meta_lr=0.1)
    theta0 = tr.theta.clone()
    info = cu(np.stack([g['a_bits'], g['b_bits']]).astype(np.float32))
    rx = cu(np.stack([g['a_y'], g['b_y']]))
    init_tx = mvn.ops.rs_encode(info[:, :5].reshape(10, -1), int(nsym)).reshape(2, 5, -1)
    seen = []
    ser = mvn.online.eval_by_word(tr, info, rx, int(nsym), float(thresh), subframes_in_frame=4, iterations=1,
                                  online_meta=True, meta_subframes=4, meta_train_iterations=1, meta_j_num=4, window_size=2,
                                  second_order=False, init_buffer=(init_tx, rx[:, :5]),
                                  draw=lambda run, high, count: (seen.append((run, high)) or [0, high - 1]))
    assert ser.shape == (2, 12) and all(h == 3 for _, h in seen) and len(seen) == 4   # 2 runs x 2 meta rounds, length 5
    assert not torch.equal(tr.theta, theta0)
    with pytest.raises(ValueError):
        mvn.online.eval_by_word(tr, info, rx, int(nsym), float(thresh), online_meta=True, meta_subframes=4,
                                weights_init='random', init_buffer=(init_tx, rx[:, :5]))
    # 'random' re-initialisation is a callable: fresh weights and a fresh optimizer for the runs that meta-train
    fresh = theta0 * 0.5
    calls = []
    tr2 = mvn.BatchedVNetTrainer(theta0, int(L), lr=float(lr), meta_lr=0.1)
    mvn.online.eval_by_word(tr2, info[:, :5], rx[:, :5], int(nsym), float(thresh), subframes_in_frame=4, iterations=1,
                            self_supervised=False, online_meta=True, meta_subframes=4, meta_train_iterations=1, meta_j_num=1,
                            second_order=False, init_buffer=(init_tx, rx[:, :5]), draw=lambda run, high, count: [1],
                            weights_init=lambda runs: (calls.append(list(runs)) or fresh[runs]))
    assert calls == [[0, 1]] and tr2.adam_step.tolist() == [1, 1]      # one round at word 4: restart, then one step


def test_online_evaluation_replays_cost2100_runs(mvn):
    """BASELINE.json configs[3]: the same replay on the COST2100 taps (tests/golden/online_cost2100.npz, recorded from
    METAVNETTrainer.eval_by_word with channel_coefficients='cost2100', channel_estimation.py:26-30): run c is
    self-supervised only, run d adds online meta-training (MAML) with the reference's torch.randint draws."""
    g = load_golden('online_cost2100')
    ser, after = run(mvn, g, ('c',))
    assert np.max(np.abs(ser[0] - g['c_ser'])) < 1e-7
    assert len(after[0]) == len(g['c_theta_after'])
    for got, want in zip(after[0], g['c_theta_after']):
        assert np.max(np.abs(got - want)) < 2e-5
    torch.manual_seed(int(g['d_seed'][0]))
    drawn = []

    def draw(run_, high, count):
        out = mvn.online._default_draw(run_, high, count)
        drawn.extend(out)
        return out
    ser, after, meta_after = run_meta(mvn, g, ('d',), draw)
    assert drawn == g['d_jhat'].tolist()
    check_meta(g, ('d',), ser, after, meta_after)

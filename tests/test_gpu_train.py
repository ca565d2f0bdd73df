"""GPU parity of the batched (meta-)training kernels against (a) the fixtures recorded from the
reference's own meta_train_loop / run_train_loop and (b) the fp64 oracle on random problems.

Tolerances ("loss and updated theta within 1e-5 relative, fp32, different summation
order"): loss |d| <= 1e-5 * |loss|; gradients |d| <= 2e-5 * max|g| per tensor; updated parameters
|d| <= 2e-5 absolute (Adam's first steps move every weight by ~lr = 1e-3, so this is 2 % of one step)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import viterbinet_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mvn():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import meta_viterbinet_b200 as m
    return m


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def pack(ws):
    return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in ws])


def assert_close_params(theta, ws_ref, atol):
    ref = pack(ws_ref)
    assert np.max(np.abs(theta - ref)) < atol, np.max(np.abs(theta - ref))


@pytest.mark.parametrize('tag,second', [('maml', True), ('fo', False)])
def test_meta_step_matches_reference_run(mvn, tag, second):
    g = load_golden('meta')
    y, tx = g[f'{tag}_y'], g[f'{tag}_tx']
    w0 = [g[f'{tag}_w0_{i}'] for i in range(6)]
    tr = mvn.BatchedVNetTrainer(cu(pack(w0)), 4, lr=1e-3, meta_lr=0.1)
    for step, j in enumerate(g[f'{tag}_jhat']):
        loss, grad = tr.meta_step(cu(y[j - 1:j]), cu(tx[j - 1:j]), cu(y[j:j + 1]), cu(tx[j:j + 1]),
                                  second_order=second, return_grad=True)
        ref_loss = g[f'{tag}_loss_q'][step]
        assert abs(float(loss[0]) - ref_loss) < 1e-5 * abs(ref_loss)
        if step == 0:
            gr = grad[0].cpu().numpy()
            o = 0
            for i in range(6):
                ref = g[f'{tag}_g1_{i}'].reshape(-1)
                assert np.max(np.abs(gr[o:o + ref.size] - ref)) < 2e-5 * np.max(np.abs(ref)) + 1e-9
                o += ref.size
        assert_close_params(tr.theta[0].cpu().numpy(), [g[f'{tag}_w{step + 1}_{i}'] for i in range(6)], 2e-5)
    assert tr.adam_step.tolist() == [3]


def test_train_steps_match_reference_run(mvn):
    g = load_golden('meta')
    w0 = [g[f'sgd_w0_{i}'] for i in range(6)]
    tr = mvn.BatchedVNetTrainer(cu(pack(w0)), 4, lr=1e-3)
    for step in range(5):
        loss = tr.train_step(cu(g['sgd_y']), cu(g['sgd_tx']))
        assert abs(float(loss[0]) - g['sgd_losses'][step]) < 1e-5 * abs(g['sgd_losses'][step])
    assert_close_params(tr.theta[0].cpu().numpy(), [g[f'sgd_w5_{i}'] for i in range(6)], 2e-5)


@pytest.mark.parametrize('L', [2, 3, 4, 5])
def test_batched_realisations_vs_oracle(mvn, L):
    """R realisations with different weights/data; N > 256 symbols exercises the blocked symbol loop."""
    rng = np.random.RandomState(L)
    S, R, Ns, Nq = 2 ** L, 5, 300, 136
    thetas, data = [], []
    for r in range(R):
        w = [rng.randn(100, 1) * .7, rng.randn(100) * .5, rng.randn(50, 100) * .15, rng.randn(50) * .1,
             rng.randn(S, 50) * .3, rng.randn(S) * .1]
        thetas.append([a.astype(np.float32) for a in w])
        data.append((rng.randn(1, Ns).astype(np.float32) * 1.5, rng.randint(0, 2, (1, Ns)).astype(np.float32),
                     rng.randn(1, Nq).astype(np.float32) * 1.5, rng.randint(0, 2, (1, Nq)).astype(np.float32)))
    theta0 = np.stack([pack(w) for w in thetas])
    ys, txs, yq, txq = [np.concatenate([d[i] for d in data]) for i in range(4)]
    for second in (True, False):
        tr = mvn.BatchedVNetTrainer(cu(theta0), L, lr=1e-3, meta_lr=0.1)
        loss, grad = tr.meta_step(cu(ys), cu(txs), cu(yq), cu(txq), second_order=second, return_grad=True)
        for r in range(R):
            lq, mg, new_w, _ = orc.maml_step(*data[r], thetas[r], None, L, meta_lr=0.1, lr=1e-3, second_order=second)
            assert abs(float(loss[r]) - lq) < 1e-5 * abs(lq)
            ref = np.concatenate([a.reshape(-1) for a in mg])
            assert np.max(np.abs(grad[r].cpu().numpy() - ref)) < 2e-5 * np.max(np.abs(ref))
            assert_close_params(tr.theta[r].cpu().numpy(), new_w, 3e-5)
    tr = mvn.BatchedVNetTrainer(cu(theta0), L, lr=1e-3)
    state = [None] * R
    ws = [list(w) for w in thetas]
    for it in range(3):
        loss = tr.train_step(cu(ys), cu(txs))
        for r in range(R):
            l, ws[r], state[r] = orc.train_step(data[r][0], data[r][1], ws[r], state[r], L, lr=1e-3)
            assert abs(float(loss[r]) - l) < 1e-5 * abs(l)
    for r in range(R):
        assert_close_params(tr.theta[r].cpu().numpy(), ws[r], 3e-5)


def test_train_phase_autograd_matches_torch(mvn):
    """VNETDetector(y,'train') -> CE loss -> backward gives the gradients torch computes for the same net."""
    g = load_golden('meta')
    w0 = [g[f'sgd_w0_{i}'] for i in range(6)]
    y, tx = g['sgd_y'], g['sgd_tx']
    det = mvn.VNETDetector(16, {'val': y.shape[1], 'train': y.shape[1]})
    with torch.no_grad():
        for p, a in zip(det.parameters(), w0):
            p.copy_(cu(a))
    labels = mvn.calculate_states(4, cu(tx))
    soft = det(cu(y), 'train')
    loss = torch.nn.CrossEntropyLoss()(soft.reshape(-1, 16), labels)
    loss.backward()
    ref_loss, ref_g, _ = orc.loss_and_grads(y.reshape(-1), orc.calculate_states(4, tx), w0)
    assert abs(float(loss) - ref_loss) < 1e-5 * abs(ref_loss)
    for p, r in zip(det.parameters(), ref_g):
        assert np.max(np.abs(p.grad.cpu().numpy().reshape(-1) - r.reshape(-1))) < 2e-5 * np.max(np.abs(r)) + 1e-9
    # the reference's own optimiser on top of our gradients reproduces its first recorded step
    opt = torch.optim.Adam(det.parameters(), lr=1e-3)
    opt.step()
    assert abs(float(loss) - g['sgd_losses'][0]) < 1e-5 * abs(g['sgd_losses'][0])
    # functional detector (FO-MAML path: create_graph=False)
    meta = mvn.META_VNETDetector(16, {'val': y.shape[1], 'train': y.shape[1]})
    var = [cu(a).requires_grad_(True) for a in w0]
    out = meta(cu(y), 'train', var)
    gr = torch.autograd.grad(torch.nn.CrossEntropyLoss()(out.reshape(-1, 16), labels), var)
    for a, r in zip(gr, ref_g):
        assert np.max(np.abs(a.cpu().numpy().reshape(-1) - r.reshape(-1))) < 2e-5 * np.max(np.abs(r)) + 1e-9


@pytest.mark.parametrize('tag,second', [('maml', True), ('fo', False)])
def test_meta_train_loop_through_autograd(mvn, tag, second):
    """trainer.py:425-453 verbatim in structure (functional detector, autograd.grad with create_graph=MAML,
    fast weights, query loss, Adam) on OUR detectors reproduces the reference's recorded run: the double backward
    of the CUDA priors is what makes the second-order case work."""
    g = load_golden('meta')
    y, tx = cu(g[f'{tag}_y']), cu(g[f'{tag}_tx'])
    T = y.shape[1]
    det = mvn.VNETDetector(16, {'val': T, 'train': T})
    meta = mvn.META_VNETDetector(16, {'val': T, 'train': T})
    with torch.no_grad():
        for p, i in zip(det.parameters(), range(6)):
            p.copy_(cu(g[f'{tag}_w0_{i}']))
    opt = torch.optim.Adam(det.parameters(), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()

    def calc_loss(soft, words):
        return crit(soft.reshape(-1, 16), mvn.calculate_states(4, words))

    for step, j in enumerate(g[f'{tag}_jhat']):
        params = list(det.parameters())
        loss_s = calc_loss(meta(y[j - 1:j], 'train', params), tx[j - 1:j])
        local = torch.autograd.grad(loss_s, params, create_graph=second)
        fast = [p - 0.1 * gr for gr, p in zip(local, params)]
        loss_q = calc_loss(meta(y[j:j + 1], 'train', fast), tx[j:j + 1])
        meta_grad = torch.autograd.grad(loss_q, params)
        ref_loss = g[f'{tag}_loss_q'][step]
        assert abs(float(loss_q) - ref_loss) < 1e-5 * abs(ref_loss)
        if step == 0:
            for i, a in enumerate(meta_grad):
                ref = g[f'{tag}_g1_{i}']
                assert np.max(np.abs(a.cpu().numpy() - ref)) < 2e-5 * np.max(np.abs(ref)) + 1e-9
        for p, a in zip(params, meta_grad):
            p.grad = a
        opt.step()
        got = pack([p.detach().cpu().numpy() for p in det.parameters()])
        assert_close_params(got, [g[f'{tag}_w{step + 1}_{i}'] for i in range(6)], 2e-5)


@pytest.mark.parametrize('L', [1, 2, 3, 5])
def test_double_backward_matches_torch(mvn, L):
    """gradient of <u, d loss/d theta> w.r.t. theta (through our kernels) == the same through a torch fp64 net."""
    S = 1 << L
    rng = np.random.default_rng(40 + L)
    n = 300
    y = rng.normal(size=n).astype(np.float32) * 1.5
    lab = rng.integers(0, S, size=n)
    ws = [rng.normal(size=s).astype(np.float32) * sc for s, sc in
          [((100, 1), 1.0), ((100,), 0.5), ((50, 100), 0.15), ((50,), 0.1), ((S, 50), 0.2), ((S,), 0.1)]]
    us = [rng.normal(size=w.shape).astype(np.float32) for w in ws]

    def run(fwd, dt, dev):
        var = [torch.tensor(w, dtype=dt, device=dev, requires_grad=True) for w in ws]
        yy = torch.tensor(y, dtype=dt, device=dev)
        loss = torch.nn.functional.cross_entropy(fwd(yy, var), torch.tensor(lab, device=dev))
        gr = torch.autograd.grad(loss, var, create_graph=True)
        inner = sum((a * torch.tensor(u, dtype=dt, device=dev)).sum() for a, u in zip(gr, us))
        return [t.detach().cpu().numpy().astype(np.float64) for t in torch.autograd.grad(inner, var)]

    def torch_fwd(yy, v):
        h1 = torch.sigmoid(yy.reshape(-1, 1) @ v[0].t() + v[1])
        return torch.relu(h1 @ v[2].t() + v[3]) @ v[4].t() + v[5]

    meta = mvn.META_VNETDetector(S, {'val': n, 'train': n})
    ours = run(lambda yy, v: meta(yy.reshape(1, -1), 'train', v).reshape(-1, S), torch.float32, 'cuda')
    ref = run(torch_fwd, torch.float64, 'cpu')
    for a, r in zip(ours, ref):
        assert np.max(np.abs(a - r)) < 5e-5 * np.max(np.abs(r)) + 1e-8, (np.max(np.abs(a - r)), np.max(np.abs(r)))


@pytest.mark.parametrize('L,T,n_stages', [(4, 136, 136), (4, 300, 290), (2, 40, 33), (5, 70, 70), (1, 9, 8)])
def test_batched_detection_with_per_realisation_weights(mvn, L, T, n_stages):
    """tr.detect(): every realisation decodes its word with its own weights in one launch; decisions are the
    reference recursion on the priors the kernel computed (bit-exact), priors within 1e-5 of fp64."""
    S, R = 1 << L, 7
    rng = np.random.RandomState(L * 100 + T)
    ws = [[(rng.randn(*s) * sc).astype(np.float32) for s, sc in
           [((100, 1), 0.8), ((100,), 0.5), ((50, 100), 0.15), ((50,), 0.1), ((S, 50), 0.3), ((S,), 0.1)]] for _ in range(R)]
    y = (rng.randn(R, T) * 1.5).astype(np.float32)
    tr = mvn.BatchedVNetTrainer(cu(np.stack([pack(w) for w in ws])), L)
    dec, pri = tr.detect(cu(y), n_stages=n_stages, return_priors=True)
    dec, pri = dec.cpu().numpy(), pri.cpu().numpy()
    assert np.all(dec[:, n_stages:] == 0)
    for r in range(R):
        exact = orc.vnet_priors(y[r:r + 1], ws[r], dtype=np.float64)[0]
        err = np.abs(pri[r, :n_stages] - exact[:n_stages]) / np.abs(exact[:n_stages]).max(axis=-1, keepdims=True)
        assert err.max() < 1e-5
        ref, _ = orc.vnet_decode_from_priors(pri[r:r + 1], n_stages)
        assert np.array_equal(dec[r:r + 1], ref)
        # the batch kernel with the same weights agrees except on near-ties (different fp32 summation order)
        one = mvn.ops.vnet_decode(cu(y[r:r + 1]), [cu(a) for a in ws[r]], n_stages=n_stages).cpu().numpy()
        assert (one != dec[r:r + 1]).sum() <= 1
    assert np.array_equal(tr.detect(cu(y), n_stages=n_stages).cpu().numpy(), dec)
    with pytest.raises(mvn.MVNError):
        tr.detect(cu(y), n_stages=T + 1)


def test_training_rejects_unsupported_trellis(mvn):
    with pytest.raises(mvn.MVNError):
        mvn.BatchedVNetTrainer(torch.zeros(1, mvn.train.param_count(7)).cuda(), 7)


def test_nan_loss_skips_the_update_like_run_train_loop(mvn):
    """run_train_loop returns before backward / optimizer.step when the loss is NaN (trainer.py:495-498): a run whose word
    holds a NaN sample keeps its weights and Adam state (and reports NaN), the other runs of the batch step normally."""
    rng = np.random.RandomState(2)
    w = [rng.randn(100, 1) * .7, rng.randn(100) * .5, rng.randn(50, 100) * .15, rng.randn(50) * .1,
         rng.randn(16, 50) * .3, rng.randn(16) * .1]
    theta0 = np.stack([pack(w)] * 3)
    tr = mvn.BatchedVNetTrainer(cu(theta0), 4, lr=1e-3)
    y = (rng.randn(3, 136) * 1.5).astype(np.float32)
    tx = rng.randint(0, 2, (3, 136)).astype(np.float32)
    tr.train_step(cu(y), cu(tx))                      # one clean step for every run
    before = [t.clone() for t in (tr.theta, tr.adam_m, tr.adam_v, tr.adam_step)]
    y[1, 17] = np.nan
    loss = tr.train_step(cu(y), cu(tx)).cpu().numpy()
    assert np.isnan(loss[1]) and np.isfinite(loss[0]) and np.isfinite(loss[2])
    for now, was in zip((tr.theta, tr.adam_m, tr.adam_v, tr.adam_step), before):
        assert torch.equal(now[1], was[1])            # untouched
        assert not torch.equal(now[0], was[0]) and not torch.equal(now[2], was[2])
    assert tr.adam_step.tolist() == [2, 1, 2]
    assert torch.isfinite(tr.theta).all()

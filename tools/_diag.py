import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import meta_viterbinet_b200 as mvn
from oracle import viterbinet_oracle as orc
def pack(w): return np.concatenate([np.asarray(a, dtype=np.float32).reshape(-1) for a in w])
def cu(a): return torch.as_tensor(np.ascontiguousarray(a)).cuda()
for L, Ns, Nq in ((5, 136, 136), (5, 300, 136), (5, 64, 32), (5, 136, 300)):
    rng = np.random.RandomState(L)
    S, R = 2 ** L, 2
    thetas, data = [], []
    for r in range(R):
        w = [rng.randn(100, 1) * .7, rng.randn(100) * .5, rng.randn(50, 100) * .15, rng.randn(50) * .1, rng.randn(S, 50) * .3, rng.randn(S) * .1]
        thetas.append([a.astype(np.float32) for a in w])
        data.append((rng.randn(1, Ns).astype(np.float32) * 1.5, rng.randint(0, 2, (1, Ns)).astype(np.float32),
                     rng.randn(1, Nq).astype(np.float32) * 1.5, rng.randint(0, 2, (1, Nq)).astype(np.float32)))
    theta0 = np.stack([pack(w) for w in thetas])
    ys, txs, yq, txq = [np.concatenate([d[i] for d in data]) for i in range(4)]
    segs = [0, 100, 200, 5200, 5250, 5250 + 50 * S, 5250 + 51 * S]
    tr = mvn.BatchedVNetTrainer(cu(theta0), L, lr=1e-3, meta_lr=0.1)
    loss, grad = tr.train_step(cu(ys), cu(txs), return_grad=True)
    for r in range(R):
        lab = orc.calculate_states(L, data[r][1])
        l, g, _ = orc.loss_and_grads(data[r][0].reshape(-1), lab, thetas[r])
        ref = np.concatenate([a.reshape(-1) for a in g]); gg = grad[r].cpu().numpy()
        errs = [float(np.max(np.abs(gg[a:b] - ref[a:b])) / np.max(np.abs(ref))) for a, b in zip(segs[:-1], segs[1:])]
        print(f'PLAIN L={L} Ns={Ns} r={r} loss err {abs(float(loss[r]) - l) / abs(l):.2e} grad seg errs ' + ' '.join(f'{e:.1e}' for e in errs))
    for second in (False,):
        tr = mvn.BatchedVNetTrainer(cu(theta0), L, lr=1e-3, meta_lr=0.1)
        loss, grad = tr.meta_step(cu(ys), cu(txs), cu(yq), cu(txq), second_order=second, return_grad=True)
        for r in range(R):
            lq, mg, new_w, _ = orc.maml_step(*data[r], thetas[r], None, L, meta_lr=0.1, lr=1e-3, second_order=second)
            ref = np.concatenate([a.reshape(-1) for a in mg]); g = grad[r].cpu().numpy()
            errs = [float(np.max(np.abs(g[a:b] - ref[a:b])) / np.max(np.abs(ref))) for a, b in zip(segs[:-1], segs[1:])]
            print(f'META  L={L} Ns={Ns} Nq={Nq} r={r} loss err {abs(float(loss[r]) - lq) / abs(lq):.2e} grad seg errs ' + ' '.join(f'{e:.1e}' for e in errs))

#!/usr/bin/env python
"""Golden fixtures for the Reed-Solomon row, generated FROM THE LIVE REFERENCE
(``python_code/ecc/rs_main.py`` encode / decode are only called, nothing is copied):

    python tests/golden/make_golden_rs.py

For every (message bytes, parity bytes) configuration: random messages, the reference's codewords, received
words with 0 .. nsym/2 + 2 corrupted bytes (so the beyond-capacity behaviour is pinned too) and the
reference's decoded words.  Stored as bits, one row per word.
"""
import os
import sys

import numpy as np

REF = os.environ.get('MVN_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = [(15, 2), (15, 4), (60, 8), (31, 6), (223, 32), (1, 2), (247, 8)]   # (k bytes, nsym)


def main():
    sys.path.insert(0, REF)
    from python_code.ecc.rs_main import encode, decode
    rng = np.random.RandomState(2024)
    out = {}
    for k, nsym in CONFIGS:
        n = k + nsym
        n_words = 24 if n > 100 else 64
        msg = rng.randint(0, 2, size=(n_words, 8 * k))
        tx = np.stack([encode(m, nsym) for m in msg])
        rx = tx.copy()
        n_err = np.zeros(n_words, dtype=np.int64)
        for w in range(n_words):
            e = w % (nsym // 2 + 3)                     # 0 .. nsym/2 + 2 corrupted bytes
            n_err[w] = e
            for p in rng.choice(n, size=min(e, n), replace=False):
                if w % 2:                               # single bit flip inside the byte
                    rx[w, 8 * p + rng.randint(8)] ^= 1
                else:                                   # random non-zero byte error
                    flip = rng.randint(1, 256)
                    rx[w, 8 * p:8 * p + 8] ^= np.unpackbits(np.array([flip], dtype=np.uint8))
        # plus words that are pure noise (far beyond capacity: exercises the "too many errors" exit)
        rx = np.concatenate([rx, rng.randint(0, 2, size=(16 if n > 100 else 96, 8 * n))])
        dec = np.stack([decode(r, nsym) for r in rx])
        tag = f'k{k}_n{nsym}'
        out[f'{tag}_msg'] = msg.astype(np.uint8)
        out[f'{tag}_tx'] = tx.astype(np.uint8)
        out[f'{tag}_rx'] = rx.astype(np.uint8)
        out[f'{tag}_dec'] = dec.astype(np.uint8)
        out[f'{tag}_nerr'] = n_err
        ok = (dec[:n_words] == msg).all(axis=1)
        print(tag, 'words', n_words, 'recovered', int(ok.sum()), 'within capacity', int((n_err <= nsym // 2).sum()))
    out['configs'] = np.array(CONFIGS)
    path = os.path.join(HERE, 'rs.npz')
    np.savez_compressed(path, **out)
    print(f'rs.npz {os.path.getsize(path) / 1024:.1f} KB')


if __name__ == '__main__':
    main()

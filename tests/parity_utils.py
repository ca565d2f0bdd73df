"""Shared by the CPU oracle tests and the GPU parity tests: protocol (ii) of SURVEY.md §8c."""
import numpy as np

from oracle import viterbinet_oracle as orc

PRIOR_RTOL = 1e-5


def explain_mismatches(dec_k, dec_ref, priors_ref, tol_rows):
    """Every frame whose bits differ from the reference's full forward must differ FIRST at a stage where, under the
    reference-grade priors `priors_ref`, the best even-state and best odd-state path metrics entering that stage are
    closer than the accumulated prior tolerance (2 * stages * tol of that frame) — a near-tie, the one case in which
    ulp-level prior differences may flip a decision (torch itself is not reproducible there).
    priors_ref may be a callable rows -> priors for just those rows (large fixtures).  Returns the number of
    differing frames."""
    bad = np.nonzero((dec_k != dec_ref).any(axis=1))[0]
    for b in bad:
        t = int(np.nonzero(dec_k[b] != dec_ref[b])[0][0])
        pr = priors_ref(np.array([b])) if callable(priors_ref) else priors_ref[b:b + 1]
        cost = -np.asarray(pr, dtype=np.float32)
        _, pm = orc.acs_decode(cost[:, :t], t)          # metrics entering stage t
        H = pm.shape[1] // 2
        v = pm[0, :H]
        even, odd = v[0::2].min(), v[1::2].min()
        gap = abs(float(even) - float(odd))
        assert gap <= 2 * t * tol_rows[b], f'frame {b} stage {t}: gap {gap} not a near-tie'
    return len(bad)


def unpack_rows(packed, T):
    return np.unpackbits(packed, axis=1)[:, :T].astype(np.float32)

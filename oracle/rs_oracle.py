"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's Reed-Solomon codec.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(meta-viterbinet_b200/) never does.  Pinned against the live reference through tests/golden/rs.npz
(tests/golden/make_golden_rs.py calls python_code/ecc/rs_main.py encode / decode on seeded words, including
words with more errors than the code can repair).

What the reference does (python_code/ecc/):
  * GF(2^8) with primitive polynomial 0x11d, generator element 2 (polynomials_manipulation.py:85-110);
  * bits <-> bytes with numpy packbits / unpackbits, most significant bit first (:119-125);
  * systematic encoding: parity = message(x) x^nsym mod g(x), g(x) = prod_{i<nsym} (x - 2^i) (rs_encoder.py:7-37,
    polynomials_manipulation.py:8-13), codeword = message bytes then parity bytes;
  * decoding (rs_main.py:21-37): syndromes S_i = r(2^i), i < nsym (rs_decoder.py:37-48); Berlekamp-Massey with the
    list-length update test (:150-203); more than nsym/2 errors claimed -> the received message bytes are returned
    unchanged (:200-203, rs_main.py:31-32); otherwise roots of the reversed locator among 2^0 .. 2^(n-1) (:206-218),
    Forney magnitudes from the positions FOUND (:88-147) — when the locator has fewer roots than its degree the
    reference still applies what it found, and so does this restatement.

Polynomials here are little-endian integer arrays (index = degree), the reference uses big-endian lists; the
list LENGTHS that drive the Berlekamp-Massey branch are tracked explicitly so the two agree in every case.
"""
import numpy as np

PRIM = 0x11D


def _tables():
    exp = np.zeros(512, dtype=np.int64)
    log = np.zeros(256, dtype=np.int64)
    x = 1
    for i in range(255):
        exp[i] = x
        log[x] = i
        x <<= 1
        if x & 0x100:
            x ^= PRIM
    exp[255:] = exp[np.arange(255, 512) - 255]
    return exp, log


EXP, LOG = _tables()


def gmul(a, b):
    """element-wise product in GF(2^8) (scalars or arrays)"""
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    return np.where((a == 0) | (b == 0), 0, EXP[(LOG[a] + LOG[b]) % 255])


def ginv(a):
    return int(EXP[(255 - LOG[a]) % 255])


def alpha(k):
    return int(EXP[k % 255])


def poly_eval(p_le, x):
    """p(x) for a little-endian coefficient array"""
    acc = 0
    for c in p_le[::-1]:
        acc = int(gmul(acc, x)) ^ int(c)
    return acc


def poly_mul(p_le, q_le):
    out = np.zeros(len(p_le) + len(q_le) - 1, dtype=np.int64)
    for d, c in enumerate(q_le):
        out[d:d + len(p_le)] ^= gmul(p_le, c)
    return out


def generator_poly(nsym):
    g = np.array([1], dtype=np.int64)
    for i in range(nsym):
        g = poly_mul(g, np.array([alpha(i), 1]))      # (x + 2^i), minus = plus
    return g


def bits_to_bytes(bits):
    bits = np.asarray(bits).astype(np.uint8).reshape(-1, 8)
    return np.packbits(bits, axis=1).reshape(-1).astype(np.int64)


def bytes_to_bits(b):
    return np.unpackbits(np.asarray(b, dtype=np.uint8).reshape(-1, 1), axis=1).reshape(-1)


def encode_bytes(msg, nsym):
    """message bytes (highest-degree coefficient first) -> codeword bytes = message | parity"""
    msg = np.asarray(msg, dtype=np.int64)
    if len(msg) + nsym > 255:
        raise ValueError('Message is too long (%i when max is 255)' % (len(msg) + nsym))
    g = generator_poly(nsym)                 # little-endian, monic: g[nsym] = 1
    rem = np.zeros(nsym, dtype=np.int64)     # remainder register, rem[nsym-1] = highest degree
    for m in msg:                            # LFSR division of m(x) x^nsym by g(x)
        fb = int(m) ^ int(rem[nsym - 1])
        rem[1:] = rem[:-1].copy()
        rem[0] = 0
        if fb:
            rem ^= gmul(g[:nsym], fb)
    return np.concatenate([msg, rem[::-1]])


def syndromes(word, nsym):
    """S_i = r(2^i), r(x) = sum_j word[j] x^(n-1-j)"""
    return np.array([poly_eval(np.asarray(word, dtype=np.int64)[::-1], alpha(i)) for i in range(nsym)], dtype=np.int64)


def berlekamp_massey(S, nsym):
    """Error locator (little-endian, lambda[0] = 1) or None when it claims more than nsym/2 errors.
    cur / old mirror the reference's err_loc / old_loc lists: arrays hold the coefficients by degree, the
    separate lengths are the reference's list lengths (leading zero coefficients count)."""
    cur, cur_len = np.zeros(nsym + 2, dtype=np.int64), 1
    old, old_len = np.zeros(nsym + 2, dtype=np.int64), 1
    cur[0] = old[0] = 1
    for i in range(nsym):
        delta = int(S[i])
        for j in range(1, cur_len):
            if i - j >= 0:
                delta ^= int(gmul(cur[j], S[i - j]))
        old = np.concatenate([[0], old[:-1]])        # old(x) <- x old(x)
        old_len += 1
        if delta:
            if old_len > cur_len:
                new = gmul(old, delta)
                old = gmul(cur, ginv(delta))
                cur = new
                cur_len, old_len = old_len, cur_len
            cur = cur ^ gmul(old, delta)
            cur_len = max(cur_len, old_len)
    while cur_len and cur[cur_len - 1] == 0:          # the reference drops leading zeros of its big-endian list
        cur_len -= 1
    if (cur_len - 1) * 2 > nsym:
        return None
    return cur[:cur_len]


def find_positions(lam, n):
    """Byte positions whose locator X = 2^(n-1-pos) is a root of the reversed locator, in the reference's order."""
    rev = lam[::-1]                                   # x^L lambda(1/x), little-endian
    return [n - 1 - i for i in range(n) if poly_eval(rev, alpha(i)) == 0]


def correct(word, S, positions):
    word = np.asarray(word, dtype=np.int64).copy()
    n = len(word)
    npos = len(positions)
    X = [alpha(n - 1 - p) for p in positions]
    loc = np.array([1], dtype=np.int64)
    for x in X:
        loc = poly_mul(loc, np.array([1, x]))          # prod (1 + X_j x)
    shifted = np.concatenate([[0], S])                 # S'(x) = sum S_i x^(i+1)
    omega = poly_mul(shifted, loc)[:npos + 1]          # mod x^(npos+1)
    for i, xi in enumerate(X):
        xi_inv = ginv(xi)
        den = 1
        for j, xj in enumerate(X):
            if j != i:
                den = int(gmul(den, 1 ^ int(gmul(xi_inv, xj))))
        if den == 0:
            raise ValueError('Could not find error magnitude')
        num = int(gmul(xi, poly_eval(omega, xi_inv)))
        word[positions[i]] ^= int(gmul(num, ginv(den))) if num else 0
    return word


def decode_bytes(word, nsym):
    """-> (message bytes, status): 0 clean, 1 corrected (all roots found), 2 more than nsym/2 errors claimed (returned
    unchanged), 3 locator with fewer roots than its degree (the reference applies the partial correction)."""
    word = np.asarray(word, dtype=np.int64)
    S = syndromes(word, nsym)
    lam = berlekamp_massey(S, nsym)
    if lam is None:
        return word[:-nsym].copy(), 2
    pos = find_positions(lam, len(word))
    fixed = correct(word, S, pos)
    status = 0 if len(lam) == 1 else (1 if len(pos) == len(lam) - 1 else 3)
    return fixed[:-nsym], status


def encode(bits, nsym):
    """ecc/rs_main.py:9-18 on one word of bits"""
    return bytes_to_bits(encode_bytes(bits_to_bytes(bits), nsym))


def decode(bits, nsym, return_status=False):
    """ecc/rs_main.py:21-37 on one received word of bits"""
    msg, status = decode_bytes(bits_to_bytes(bits), nsym)
    out = bytes_to_bits(msg)
    return (out, status) if return_status else out

"""ORACLE (test infrastructure only).

Bayer-Groth (Eurocrypt 2012) correct-shuffle argument exactly as SURVEY.md Appendix B restates
it, behind the reference's glue:

  shuffle_and_remask  barnett-smart-card-protocol/src/discrete_log_cards/mod.rs:380-418
  verify_shuffle      .../mod.rs:420-443
  remask              .../remasking.rs:9-22  (-> masking.rs:10-20 -> ElGamal::encrypt)

The bodies of `shuffle::ShuffleArgument::{prove,verify}` live in the un-vendored, unpinned git
dependency `proof-essentials` (Cargo.toml:18) which is absent here, so the transcript byte
order and the prover-randomness draw order are THIS repository's definition (Appendix B.6;
restated in include/mpshuffle.h).  PARITY UNPINNED at byte level against upstream.  What the
reference's own test pins (tests.rs:175-227) is reproduced by tests/test_oracle_protocol.py:
prove->verify == Ok, and a wrong output deck fails with "Hadamard Product (5.1)".

Everything here is plain Python big-int code meant to be obviously correct, not fast.
"""
import contextlib

from . import stark
from . import transcript as _transcript
from .stark import N as Q  # scalar-field modulus
from .transcript import FiatShamirRng, SHUFFLE_RNG_SEED

POINT_BYTES = 64  # C-ABI affine point of the active curve (x || y)


class _Bls12_377Group:
    """The names this module uses from `stark`, over BLS12-377 G1 (the reference's second instantiation,
    examples/parameter_selection.rs:25-29): 48-byte coordinates, so the ark `ToBytes` point is
    x || y || infinity = 97 bytes and the C-ABI point 96 bytes; scalars stay 32 bytes."""

    def __init__(self):
        from . import bls12_377 as b
        self.P, self.N, self.INF, self.G = b.P, b.N, None, b.G
        self.add, self.mul, self.msm, self.neg, self.sub = b.add, b.mul, b.msm, b.neg, b.sub
        self.point_to_bytes64, self.point_from_bytes64 = b.point_to_bytes, b.point_from_bytes
        self.fe_to_bytes = b.scalar_to_bytes  # only scalars are written through this name here
        self._fq = b.fe_to_bytes

    def fe_from_bytes(self, buf):
        return int.from_bytes(buf, "little")

    def point_to_bytes65(self, pt):
        if pt is None:
            return self._fq(0) + self._fq(1) + b"\x01"
        return self._fq(pt[0]) + self._fq(pt[1]) + b"\x00"


@contextlib.contextmanager
def curve(name):
    """`with curve("bls12_377"): ...` runs this module's protocol over BLS12-377 G1 instead of the Stark
    curve (group, scalar field of the challenges, byte widths).  Test infrastructure: not re-entrant."""
    global stark, Q, POINT_BYTES, CT_ZERO
    assert name in ("stark", "bls12_377")
    saved = (stark, Q, POINT_BYTES, _transcript.SCALAR_MODULUS)
    if name == "bls12_377":
        stark = _Bls12_377Group()
        Q, POINT_BYTES = stark.N, 96
        _transcript.SCALAR_MODULUS = stark.N
    try:
        yield stark
    finally:
        stark, Q, POINT_BYTES, _transcript.SCALAR_MODULUS = saved

# verification status codes (shared with include/mpshuffle.h)
OK = 0
ERR_HADAMARD = 1          # "Hadamard Product (5.1)"   (tests.rs:223-225)
ERR_ZERO = 2              # "Zero Argument (5.2)"
ERR_SVP = 3               # "Single Value Product (5.3)"
ERR_MULTIEXP = 4          # "Multi Exponentiation (4)"
ERR_STRINGS = {
    ERR_HADAMARD: "Hadamard Product (5.1)",
    ERR_ZERO: "Zero Argument (5.2)",
    ERR_SVP: "Single Value Product (5.3)",
    ERR_MULTIEXP: "Multi Exponentiation (4)",
}


# ----------------------------------------------------------------------------- building blocks
class Params:
    """DLCards `Parameters` (mod.rs:37-61) + what `setup` (mod.rs:105-121) creates:
    ElGamal generator g, Pedersen key (g_1..g_n, h), extra generator ghat."""

    def __init__(self, m, n, enc_g, ck_g, ck_h, ghat):
        assert len(ck_g) == n
        self.m, self.n = m, n
        self.enc_g, self.ck_g, self.ck_h, self.ghat = enc_g, list(ck_g), ck_h, ghat

    def to_bytes(self, pk):
        out = stark.point_to_bytes65(self.enc_g) + stark.point_to_bytes65(pk)
        for g in self.ck_g:
            out += stark.point_to_bytes65(g)
        return out + stark.point_to_bytes65(self.ck_h) + stark.point_to_bytes65(self.ghat)


def commit(pp, values, r):
    """Pedersen vector commitment com(v; r) = r*h + sum v_j*g_j, |v| <= n  (SURVEY.md A6)."""
    assert len(values) <= pp.n
    acc = stark.mul(pp.ck_h, r)
    for v, g in zip(values, pp.ck_g):
        acc = stark.add(acc, stark.mul(g, v))
    return acc


def encrypt(pp, pk, msg, r):
    """ElGamal::encrypt = (r*g, msg + r*pk)  (masking.rs:17)."""
    return (stark.mul(pp.enc_g, r), stark.add(msg, stark.mul(pk, r)))


def ct_add(c, d):
    return (stark.add(c[0], d[0]), stark.add(c[1], d[1]))


def ct_mul(c, k):
    return (stark.mul(c[0], k), stark.mul(c[1], k))


CT_ZERO = (stark.INF, stark.INF)


def ct_msm(cts, scalars):
    acc = CT_ZERO
    for c, k in zip(cts, scalars):
        acc = ct_add(acc, ct_mul(c, k))
    return acc


def remask(pp, pk, card, alpha):
    """remasking.rs:9-22: card + Enc(identity; alpha)."""
    return ct_add(card, encrypt(pp, pk, stark.INF, alpha))


def permute_array(mapping, arr):
    """proof-essentials Permutation::permute_array: out[i] = in[mapping[i]]."""
    return [arr[j] for j in mapping]


def shuffle_and_remask_deck(pp, pk, deck, masking_factors, mapping):
    """mod.rs:388-395."""
    permuted = permute_array(mapping, deck)
    return [remask(pp, pk, c, rho) for c, rho in zip(permuted, masking_factors)]


def ct_bytes(c):
    return stark.point_to_bytes65(c[0]) + stark.point_to_bytes65(c[1])


def pts_bytes(pts):
    return b"".join(stark.point_to_bytes65(p) for p in pts)


def bilinear(u, v, y):
    """u * v = sum_j u_j v_j y^j  (j = 1..n)  -- the map of Appendix B.3."""
    acc, yp = 0, 1
    for a, b in zip(u, v):
        yp = yp * y % Q
        acc = (acc + a * b % Q * yp) % Q
    return acc


def chunks(v, m, n):
    assert len(v) == m * n
    return [v[k * n:(k + 1) * n] for k in range(m)]


def lincomb(coeffs, vecs):
    n = len(vecs[0])
    return [sum(c * v[j] for c, v in zip(coeffs, vecs)) % Q for j in range(n)]


def pt_lincomb(coeffs, pts):
    return stark.msm(pts, coeffs)


class Rand:
    """Flat prover-randomness buffer consumed in the order of Appendix B.6."""

    def __init__(self, scalars):
        self.s, self.i = list(scalars), 0

    def one(self):
        v = self.s[self.i]
        self.i += 1
        return v % Q

    def vec(self, k):
        return [self.one() for _ in range(k)]


def prover_randomness_len(m, n):
    return 11 * m + 5 * n


# ----------------------------------------------------------------------------- B.4 zero argument
def zero_prove(pp, fs, rand, cA, cB, y, A, r, Bv, s):
    m, n = len(A), pp.n
    a0, bm1 = rand.vec(n), rand.vec(n)
    r0, sm1 = rand.one(), rand.one()
    t = [rand.one() if k != m + 1 else 0 for k in range(2 * m + 1)]
    Aext = [a0] + A                    # a_0 .. a_m
    Bext = Bv + [bm1]                  # b_1 .. b_{m+1}
    rext, sext = [r0] + r, s + [sm1]
    c_A0, c_Bm1 = commit(pp, a0, r0), commit(pp, bm1, sm1)
    d = [0] * (2 * m + 1)
    for i in range(m + 1):
        for j in range(1, m + 2):
            k = i + m + 1 - j
            d[k] = (d[k] + bilinear(Aext[i], Bext[j - 1], y)) % Q
    assert d[m + 1] == 0, "zero-argument witness does not satisfy the relation"
    c_D = [commit(pp, [d[k]], t[k]) for k in range(2 * m + 1)]
    fs.absorb(b"zero_argument" + pts_bytes([c_A0, c_Bm1] + c_D))
    x = fs.challenge()
    xp = [pow(x, k, Q) for k in range(2 * m + 1)]
    a = lincomb(xp[:m + 1], Aext)
    rr = sum(xp[i] * rext[i] for i in range(m + 1)) % Q
    b = lincomb([xp[m + 1 - j] for j in range(1, m + 2)], Bext)
    ss = sum(xp[m + 1 - j] * sext[j - 1] for j in range(1, m + 2)) % Q
    tt = sum(xp[k] * t[k] for k in range(2 * m + 1)) % Q
    return dict(c_A0=c_A0, c_Bm1=c_Bm1, c_D=c_D, a=a, b=b, r=rr, s=ss, t=tt)


def zero_verify(pp, fs, cA, cB, y, pf):
    m = len(cA)
    fs.absorb(b"zero_argument" + pts_bytes([pf["c_A0"], pf["c_Bm1"]] + pf["c_D"]))
    x = fs.challenge()
    xp = [pow(x, k, Q) for k in range(2 * m + 1)]
    if pf["c_D"][m + 1] is not stark.INF:
        return ERR_ZERO
    if pt_lincomb(xp[:m + 1], [pf["c_A0"]] + cA) != commit(pp, pf["a"], pf["r"]):
        return ERR_ZERO
    if pt_lincomb([xp[m + 1 - j] for j in range(1, m + 2)], cB + [pf["c_Bm1"]]) != commit(pp, pf["b"], pf["s"]):
        return ERR_ZERO
    if pt_lincomb(xp, pf["c_D"]) != commit(pp, [bilinear(pf["a"], pf["b"], y)], pf["t"]):
        return ERR_ZERO
    return OK


# ----------------------------------------------------------------------------- B.3 Hadamard argument
def hadamard_prove(pp, fs, rand, cA, c_b, A, r, b, s):
    m, n = len(A), pp.n
    Bv = [A[0]]
    for i in range(1, m):
        Bv.append([u * v % Q for u, v in zip(Bv[-1], A[i])])
    assert Bv[-1] == b
    sv = [r[0]] + [rand.one() for _ in range(m - 2)] + [s] if m >= 2 else [r[0]]
    c_B = [cA[0]] + [commit(pp, Bv[i], sv[i]) for i in range(1, m - 1)] + [c_b]
    fs.absorb(b"hadamard_argument" + pts_bytes([c_b] + c_B))
    x, y = fs.challenge(), fs.challenge()
    xp = [pow(x, k, Q) for k in range(m)]
    # zero-argument instance: A' = (a_2..a_m, -1), B' = (x b_1, .., x^{m-1} b_{m-1}, d)
    minus1 = [Q - 1] * n
    c_m1 = commit(pp, minus1, 0)
    zA = A[1:] + [minus1]
    zr = r[1:] + [0]
    D = [[xp[i] * v % Q for v in Bv[i - 1]] for i in range(1, m)]
    dlast = lincomb(xp[1:m], Bv[1:m])
    zB = D + [dlast]
    zs = [xp[i] * sv[i - 1] % Q for i in range(1, m)] + [sum(xp[i] * sv[i] for i in range(1, m)) % Q]
    c_Dv = [stark.mul(c_B[i - 1], xp[i]) for i in range(1, m)]
    c_Dl = pt_lincomb(xp[1:m], c_B[1:m])
    zero = zero_prove(pp, fs, rand, cA[1:] + [c_m1], c_Dv + [c_Dl], y, zA, zr, zB, zs)
    return dict(c_B=c_B, zero=zero)


def hadamard_verify(pp, fs, cA, c_b, pf):
    m, n = len(cA), pp.n
    c_B = pf["c_B"]
    if len(c_B) != m or c_B[0] != cA[0] or c_B[m - 1] != c_b:
        return ERR_HADAMARD
    fs.absorb(b"hadamard_argument" + pts_bytes([c_b] + c_B))
    x, y = fs.challenge(), fs.challenge()
    xp = [pow(x, k, Q) for k in range(m)]
    c_m1 = commit(pp, [Q - 1] * n, 0)
    c_Dv = [stark.mul(c_B[i - 1], xp[i]) for i in range(1, m)]
    c_Dl = pt_lincomb(xp[1:m], c_B[1:m])
    return zero_verify(pp, fs, cA[1:] + [c_m1], c_Dv + [c_Dl], y, pf["zero"])


# ----------------------------------------------------------------------------- B.5 single-value product
def svp_prove(pp, fs, rand, c_a, b, a, r):
    n = len(a)
    bk = [a[0]]
    for i in range(1, n):
        bk.append(bk[-1] * a[i] % Q)
    assert bk[-1] == b % Q
    d = rand.vec(n)
    r_d = rand.one()
    delta = [d[0]] + [rand.one() for _ in range(n - 2)] + [0]
    s1, sx = rand.one(), rand.one()
    c_d = commit(pp, d, r_d)
    c_delta = commit(pp, [(-delta[i] * d[i + 1]) % Q for i in range(n - 1)], s1)
    c_Delta = commit(pp, [(delta[i + 1] - a[i + 1] * delta[i] - bk[i] * d[i + 1]) % Q for i in range(n - 1)], sx)
    fs.absorb(b"single_value_product_argument" + pts_bytes([c_d, c_delta, c_Delta]))
    x = fs.challenge()
    at = [(x * a[i] + d[i]) % Q for i in range(n)]
    bt = [(x * bk[i] + delta[i]) % Q for i in range(n)]
    return dict(c_d=c_d, c_delta=c_delta, c_Delta=c_Delta, a=at, b=bt,
                r=(x * r + r_d) % Q, s=(x * sx + s1) % Q)


def svp_verify(pp, fs, c_a, b, pf):
    n = pp.n
    fs.absorb(b"single_value_product_argument" + pts_bytes([pf["c_d"], pf["c_delta"], pf["c_Delta"]]))
    x = fs.challenge()
    at, bt = pf["a"], pf["b"]
    if len(at) != n or len(bt) != n:
        return ERR_SVP
    if stark.add(stark.mul(c_a, x), pf["c_d"]) != commit(pp, at, pf["r"]):
        return ERR_SVP
    e = [(x * bt[i + 1] - bt[i] * at[i + 1]) % Q for i in range(n - 1)]
    if stark.add(stark.mul(pf["c_Delta"], x), pf["c_delta"]) != commit(pp, e, pf["s"]):
        return ERR_SVP
    if bt[0] != at[0] or bt[n - 1] != x * b % Q:
        return ERR_SVP
    return OK


# ----------------------------------------------------------------------------- B.2 product argument
def product_prove(pp, fs, rand, cA, b, A, r):
    m, n = len(A), pp.n
    s = rand.one()
    col = A[0]
    for i in range(1, m):
        col = [u * v % Q for u, v in zip(col, A[i])]
    c_b = commit(pp, col, s)
    had = hadamard_prove(pp, fs, rand, cA, c_b, A, r, col, s)
    svp = svp_prove(pp, fs, rand, c_b, b, col, s)
    return dict(c_b=c_b, hadamard=had, svp=svp)


def product_verify(pp, fs, cA, b, pf):
    st = hadamard_verify(pp, fs, cA, pf["c_b"], pf["hadamard"])
    if st != OK:
        return st
    return svp_verify(pp, fs, pf["c_b"], b, pf["svp"])


# ----------------------------------------------------------------------------- B.5' multi-exponentiation
def multiexp_prove(pp, pk, fs, rand, Cch, C, cA, A, r, rho):
    m, n = len(A), pp.n
    a0 = rand.vec(n)
    r0 = rand.one()
    b, s, tau = [0] * (2 * m), [0] * (2 * m), [0] * (2 * m)
    for k in range(2 * m):
        if k == m:
            b[k], s[k], tau[k] = 0, 0, rho % Q
        else:
            b[k], s[k], tau[k] = rand.one(), rand.one(), rand.one()
    Aext = [a0] + A
    c_A0 = commit(pp, a0, r0)
    c_B = [commit(pp, [b[k]], s[k]) for k in range(2 * m)]
    E = []
    for k in range(2 * m):
        acc = encrypt(pp, pk, stark.mul(pp.ghat, b[k]), tau[k])
        for i in range(1, m + 1):
            j = k - m + i
            if 0 <= j <= m:
                acc = ct_add(acc, ct_msm(Cch[i - 1], Aext[j]))
        E.append(acc)
    assert E[m] == C, "multi-exp witness does not open the statement"
    fs.absorb(b"multi_exponentiation_argument" + pts_bytes([c_A0] + c_B) + b"".join(ct_bytes(e) for e in E))
    x = fs.challenge()
    xp = [pow(x, k, Q) for k in range(2 * m)]
    return dict(c_A0=c_A0, c_B=c_B, E=E,
                a=lincomb(xp[:m + 1], Aext),
                r=sum(xp[j] * ([r0] + r)[j] for j in range(m + 1)) % Q,
                b=sum(xp[k] * b[k] for k in range(2 * m)) % Q,
                s=sum(xp[k] * s[k] for k in range(2 * m)) % Q,
                tau=sum(xp[k] * tau[k] for k in range(2 * m)) % Q)


def multiexp_verify(pp, pk, fs, Cch, C, cA, pf):
    m, n = len(cA), pp.n
    c_B, E = pf["c_B"], pf["E"]
    fs.absorb(b"multi_exponentiation_argument" + pts_bytes([pf["c_A0"]] + c_B) + b"".join(ct_bytes(e) for e in E))
    x = fs.challenge()
    xp = [pow(x, k, Q) for k in range(2 * m)]
    if c_B[m] is not stark.INF or E[m] != C:
        return ERR_MULTIEXP
    if pt_lincomb(xp[:m + 1], [pf["c_A0"]] + cA) != commit(pp, pf["a"], pf["r"]):
        return ERR_MULTIEXP
    if pt_lincomb(xp, c_B) != commit(pp, [pf["b"]], pf["s"]):
        return ERR_MULTIEXP
    lhs = ct_msm(E, xp)
    flat_scalars, flat_cts = [], []
    for i in range(1, m + 1):
        flat_scalars += [xp[m - i] * aj % Q for aj in pf["a"]]
        flat_cts += Cch[i - 1]
    rhs = ct_add(encrypt(pp, pk, stark.mul(pp.ghat, pf["b"]), pf["tau"]), ct_msm(flat_cts, flat_scalars))
    if lhs != rhs:
        return ERR_MULTIEXP
    return OK


# ----------------------------------------------------------------------------- B.1 shuffle argument
def _absorb_statement(fs, pp, pk, deck, deck2, c_A):
    fs.absorb(b"shuffle_argument" + pp.to_bytes(pk) + b"".join(ct_bytes(c) for c in deck)
              + b"".join(ct_bytes(c) for c in deck2) + pts_bytes(c_A))


def _product_statement(pp, c_A, c_B, x, y, z, N_cards):
    """c_D[k] = y*c_A[k] + c_B[k] + com(-z..-z; 0),  b* = prod_{i=1..N}(y*i + x^i - z)."""
    c_mz = commit(pp, [(-z) % Q] * pp.n, 0)
    c_D = [stark.add(stark.add(stark.mul(ca, y), cb), c_mz) for ca, cb in zip(c_A, c_B)]
    bstar, xi = 1, 1
    for i in range(1, N_cards + 1):
        xi = xi * x % Q
        bstar = bstar * ((y * i + xi - z) % Q) % Q
    return c_D, bstar


def shuffle_prove(pp, pk, deck, deck2, mapping, rho, rand_scalars, fs=None):
    """ShuffleArgument::prove (call site mod.rs:409-415).  `mapping` is the 0-based
    Permutation.mapping, `rho` the masking factors, `rand_scalars` the flat randomness."""
    m, n = pp.m, pp.n
    Nc = m * n
    assert len(deck) == len(deck2) == len(mapping) == len(rho) == Nc
    fs = fs or FiatShamirRng(SHUFFLE_RNG_SEED)
    rand = Rand(rand_scalars)
    r, s = rand.vec(m), rand.vec(m)
    a = [mapping[i] + 1 for i in range(Nc)]
    c_A = [commit(pp, ch, rk) for ch, rk in zip(chunks(a, m, n), r)]
    _absorb_statement(fs, pp, pk, deck, deck2, c_A)
    x = fs.challenge()
    b = [pow(x, ai, Q) for ai in a]
    c_B = [commit(pp, ch, sk) for ch, sk in zip(chunks(b, m, n), s)]
    fs.absorb(b"shuffle_argument_b" + pts_bytes(c_B))
    y, z = fs.challenge(), fs.challenge()
    d = [(y * ai + bi - z) % Q for ai, bi in zip(a, b)]
    t = [(y * rk + sk) % Q for rk, sk in zip(r, s)]
    c_D, bstar = _product_statement(pp, c_A, c_B, x, y, z, Nc)
    product = product_prove(pp, fs, rand, c_D, bstar, chunks(d, m, n), t)
    rho_star = (-sum(ri * bi for ri, bi in zip(rho, b))) % Q
    Chat = ct_msm(deck, [pow(x, i, Q) for i in range(1, Nc + 1)])
    multiexp = multiexp_prove(pp, pk, fs, rand, chunks(deck2, m, n), Chat, c_B, chunks(b, m, n), s, rho_star)
    assert rand.i == prover_randomness_len(m, n)
    return dict(c_A=c_A, c_B=c_B, product=product, multiexp=multiexp)


def shuffle_verify(pp, pk, deck, deck2, proof, fs=None):
    """ShuffleArgument::verify (call site mod.rs:437-442).  Returns a status code."""
    m, n = pp.m, pp.n
    Nc = m * n
    fs = fs or FiatShamirRng(SHUFFLE_RNG_SEED)
    c_A, c_B = proof["c_A"], proof["c_B"]
    _absorb_statement(fs, pp, pk, deck, deck2, c_A)
    x = fs.challenge()
    fs.absorb(b"shuffle_argument_b" + pts_bytes(c_B))
    y, z = fs.challenge(), fs.challenge()
    c_D, bstar = _product_statement(pp, c_A, c_B, x, y, z, Nc)
    st = product_verify(pp, fs, c_D, bstar, proof["product"])
    if st != OK:
        return st
    Chat = ct_msm(deck, [pow(x, i, Q) for i in range(1, Nc + 1)])
    return multiexp_verify(pp, pk, fs, chunks(deck2, m, n), Chat, c_B, proof["multiexp"])


def shuffle_and_remask(pp, pk, deck, masking_factors, mapping, rand_scalars):
    """DLCards::shuffle_and_remask (mod.rs:380-418)."""
    deck2 = shuffle_and_remask_deck(pp, pk, deck, masking_factors, mapping)
    return deck2, shuffle_prove(pp, pk, deck, deck2, mapping, masking_factors, rand_scalars)


# ----------------------------------------------------------------------------- flat proof encoding
def _p64(p):
    return stark.point_to_bytes64(p)


def _f32(v):
    return stark.fe_to_bytes(v % Q)


def proof_len(m, n):
    return (11 * m + 8) * POINT_BYTES + (5 * n + 9) * 32


def proof_to_bytes(pf):
    """Flat C-ABI proof layout (include/mpshuffle.h): 64-byte affine points (all-zero =
    identity), 32-byte LE canonical scalars, in the order of Appendix B."""
    z, sv, me = pf["product"]["hadamard"]["zero"], pf["product"]["svp"], pf["multiexp"]
    out = b"".join(_p64(p) for p in pf["c_A"] + pf["c_B"])
    out += _p64(pf["product"]["c_b"])
    out += b"".join(_p64(p) for p in pf["product"]["hadamard"]["c_B"])
    out += b"".join(_p64(p) for p in [z["c_A0"], z["c_Bm1"]] + z["c_D"])
    out += b"".join(_f32(v) for v in z["a"] + z["b"] + [z["r"], z["s"], z["t"]])
    out += b"".join(_p64(p) for p in [sv["c_d"], sv["c_delta"], sv["c_Delta"]])
    out += b"".join(_f32(v) for v in sv["a"] + sv["b"] + [sv["r"], sv["s"]])
    out += b"".join(_p64(p) for p in [me["c_A0"]] + me["c_B"])
    out += b"".join(_p64(e[0]) + _p64(e[1]) for e in me["E"])
    out += b"".join(_f32(v) for v in me["a"] + [me["r"], me["b"], me["s"], me["tau"]])
    return out


class NonCanonicalScalar(ValueError):
    """a serialized scalar is not a canonical residue (what ark-serialize's deserialiser rejects)"""


def proof_from_bytes(buf, m, n):
    assert len(buf) == proof_len(m, n)
    pos = [0]

    def pt():
        p = stark.point_from_bytes64(buf[pos[0]:pos[0] + POINT_BYTES])
        pos[0] += POINT_BYTES
        return p

    def fr():
        v = stark.fe_from_bytes(buf[pos[0]:pos[0] + 32])
        pos[0] += 32
        # `Proof: CanonicalDeserialize` (reference src/lib.rs:45-71): ark-serialize rejects a field element whose
        # integer is >= the modulus -- s and s + order must not both be accepted (proof malleability)
        if v >= stark.N:
            raise NonCanonicalScalar("proof scalar >= group order")
        return v

    pf = dict(c_A=[pt() for _ in range(m)], c_B=[pt() for _ in range(m)])
    c_b = pt()
    hB = [pt() for _ in range(m)]
    z = dict(c_A0=pt(), c_Bm1=pt(), c_D=[pt() for _ in range(2 * m + 1)])
    z["a"] = [fr() for _ in range(n)]
    z["b"] = [fr() for _ in range(n)]
    z["r"], z["s"], z["t"] = fr(), fr(), fr()
    sv = dict(c_d=pt(), c_delta=pt(), c_Delta=pt())
    sv["a"] = [fr() for _ in range(n)]
    sv["b"] = [fr() for _ in range(n)]
    sv["r"], sv["s"] = fr(), fr()
    me = dict(c_A0=pt(), c_B=[pt() for _ in range(2 * m)])
    me["E"] = [(pt(), pt()) for _ in range(2 * m)]
    me["a"] = [fr() for _ in range(n)]
    me["r"], me["b"], me["s"], me["tau"] = fr(), fr(), fr(), fr()
    pf["product"] = dict(c_b=c_b, hadamard=dict(c_B=hB, zero=z), svp=sv)
    pf["multiexp"] = me
    return pf

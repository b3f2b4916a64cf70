"""ORACLE (test infrastructure only -- never imported by the product path).

Stark-curve field / group arithmetic on Python big ints.

Restates the arithmetic the reference obtains from its un-vendored dependencies
(`starknet-curve`, `ark-ec 0.3`, `ark-ff 0.3`; reference crate
barnett-smart-card-protocol/Cargo.toml:11-12,20).  Constants are SURVEY.md
Appendix A7.  Byte layouts follow ark-ff/ark-ec 0.3 `ToBytes`:
field element = 32 bytes little-endian canonical (non-Montgomery); affine point
= x || y || infinity-flag (65 bytes), identity = (0, 1, true).

PARITY UNPINNED w.r.t. the upstream Rust crates (they cannot be built here and
the reference ships no golden vectors); pinned against mathematics: p, n prime,
G on curve, n*G = O, group laws (tests/test_oracle_math.py).
"""

P = 0x0800000000000011000000000000000000000000000000000000000000000001
N = 0x0800000000000010FFFFFFFFFFFFFFFFB781126DCAE7B2321E66A241ADC64D2F
A = 1
B = 0x06F21413EFBE40DE150E596D72F7A8C5609AD26C15C915C1F4CDFCB99CEE9E89
GX = 0x01EF15C18599971B7BECED415A40F0C7DEACFD9B0D1819E03D723D8BC943CFCA
GY = 0x005668060AA49730B7BE4801DF46EC62DE53ECD11ABE43A32873000C36E8DC1F
G = (GX, GY)
R256 = 1 << 256
INF = None  # affine identity


# ----------------------------------------------------------------------------- fields
def fq_inv(a):
    return pow(a, -1, P)


def fr_inv(a):
    return pow(a, -1, N)


def fq_sqrt(a):
    """Tonelli-Shanks over F_p (p - 1 = 2^192 * odd).  Returns a root or None."""
    a %= P
    if a == 0:
        return 0
    if pow(a, (P - 1) // 2, P) != 1:
        return None
    q, s = P - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (P - 1) // 2, P) != P - 1:
        z += 1
    m, c, t, r = s, pow(z, q, P), pow(a, q, P), pow(a, (q + 1) // 2, P)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % P
            i += 1
        b = pow(c, 1 << (m - i - 1), P)
        m, c = i, b * b % P
        t, r = t * c % P, r * b % P
    return r


# ----------------------------------------------------------------------------- curve
def is_on_curve(pt):
    if pt is INF:
        return True
    x, y = pt
    return (y * y - (x * x * x + A * x + B)) % P == 0


def neg(pt):
    if pt is INF:
        return INF
    return (pt[0], (-pt[1]) % P)


def add(p1, p2):
    """Affine short-Weierstrass addition (complete: handles O, P+P, P+(-P))."""
    if p1 is INF:
        return p2
    if p2 is INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = (3 * x1 * x1 + A) * fq_inv(2 * y1) % P
    else:
        lam = (y2 - y1) * fq_inv(x2 - x1) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def sub(p1, p2):
    return add(p1, neg(p2))


# Jacobian internals (speed only; results are always normalised to affine)
def _jdbl(X, Y, Z):
    if Y == 0 or Z == 0:
        return (1, 1, 0)
    XX = X * X % P
    YY = Y * Y % P
    YYYY = YY * YY % P
    ZZ = Z * Z % P
    S = 4 * X * YY % P
    M = (3 * XX + A * ZZ * ZZ) % P
    X3 = (M * M - 2 * S) % P
    Y3 = (M * (S - X3) - 8 * YYYY) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jadd_affine(X1, Y1, Z1, x2, y2):
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    H = (U2 - X1) % P
    r = (S2 - Y1) % P
    if H == 0:
        if r == 0:
            return _jdbl(X1, Y1, Z1)
        return (1, 1, 0)
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (r * r - HHH - 2 * V) % P
    Y3 = (r * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def _jto_affine(X, Y, Z):
    if Z == 0:
        return INF
    zi = fq_inv(Z)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def mul(pt, k):
    """k * pt, MSB-first double-and-add (the algorithm of ark-ec 0.3 `AffineCurve::mul`,
    SURVEY.md A2); k is reduced mod the group order first."""
    k %= N
    if pt is INF or k == 0:
        return INF
    x, y = pt
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jdbl(*acc)
        if bit == "1":
            acc = _jadd_affine(*acc, x, y)
    return _jto_affine(*acc)


def msm(points, scalars):
    """sum_i scalars[i] * points[i]  (obviously-correct definition; no windows)."""
    assert len(points) == len(scalars)
    acc = INF
    for pt, k in zip(points, scalars):
        acc = add(acc, mul(pt, k))
    return acc


# ----------------------------------------------------------------------------- bytes
def fe_to_bytes(a):
    return int(a).to_bytes(32, "little")


def fe_from_bytes(b):
    return int.from_bytes(b, "little")


def point_to_bytes65(pt):
    """ark-ec 0.3 GroupAffine::write = x || y || infinity."""
    if pt is INF:
        return fe_to_bytes(0) + fe_to_bytes(1) + b"\x01"
    return fe_to_bytes(pt[0]) + fe_to_bytes(pt[1]) + b"\x00"


def point_from_bytes65(b):
    assert len(b) == 65
    if b[64]:
        return INF
    return (fe_from_bytes(b[:32]), fe_from_bytes(b[32:64]))


def point_to_bytes64(pt):
    """C-ABI input layout: x || y, identity encoded as the all-zero 64 bytes
    ((0,0) is not on the curve since b != 0)."""
    if pt is INF:
        return bytes(64)
    return fe_to_bytes(pt[0]) + fe_to_bytes(pt[1])


def point_from_bytes64(b):
    assert len(b) == 64
    if b == bytes(64):
        return INF
    return (fe_from_bytes(b[:32]), fe_from_bytes(b[32:]))

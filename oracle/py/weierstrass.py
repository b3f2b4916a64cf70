"""ORACLE (test infrastructure only -- never imported by the product path).

Short-Weierstrass group arithmetic y^2 = x^3 + a*x + b over a prime field, on Python big ints,
parametrised by the curve: the generic form of `stark.py`, used for the second curve the
reference instantiates its protocol over (`ark_bls12_377::G1Projective`, reference
barnett-smart-card-protocol/examples/parameter_selection.rs:25-26; the trait is generic over
`C: ProjectiveCurve`, src/discrete_log_cards/mod.rs:86).

Byte layouts follow ark-ff / ark-ec 0.3 `ToBytes`: a field element is
`fe_bytes` bytes little-endian canonical (non-Montgomery); the C-ABI point is x || y with the
all-zero string for the identity ((0,0) is on neither curve: b != 0).

PARITY UNPINNED w.r.t. the upstream Rust crates (see stark.py); pinned against mathematics in
tests/test_oracle_bls12_377.py.
"""


class Curve:
    def __init__(self, name, p, n, a, b, g, fe_bytes, scalar_bytes=32):
        self.name, self.P, self.N, self.A, self.B, self.G = name, p, n, a, b, g
        self.fe_bytes, self.scalar_bytes = fe_bytes, scalar_bytes
        self.INF = None

    # ------------------------------------------------------------------------- affine law
    def is_on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        return (y * y - (x * x * x + self.A * x + self.B)) % self.P == 0

    def neg(self, pt):
        return None if pt is None else (pt[0], (-pt[1]) % self.P)

    def add(self, p1, p2):
        """Affine addition, complete (O, P + P, P + (-P))."""
        if p1 is None:
            return p2
        if p2 is None:
            return p1
        P = self.P
        x1, y1 = p1
        x2, y2 = p2
        if x1 == x2:
            if (y1 + y2) % P == 0:
                return None
            lam = (3 * x1 * x1 + self.A) * pow(2 * y1, -1, P) % P
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
        x3 = (lam * lam - x1 - x2) % P
        return (x3, (lam * (x1 - x3) - y1) % P)

    def sub(self, p1, p2):
        return self.add(p1, self.neg(p2))

    # ------------------------------------------------------------------------- scalar mul
    def _jdbl(self, X, Y, Z):
        P = self.P
        if Y == 0 or Z == 0:
            return (1, 1, 0)
        YY = Y * Y % P
        S = 4 * X * YY % P
        ZZ = Z * Z % P
        M = (3 * X * X + self.A * ZZ * ZZ) % P
        X3 = (M * M - 2 * S) % P
        return (X3, (M * (S - X3) - 8 * YY * YY) % P, 2 * Y * Z % P)

    def _jmadd(self, X1, Y1, Z1, x2, y2):
        P = self.P
        if Z1 == 0:
            return (x2, y2, 1)
        Z1Z1 = Z1 * Z1 % P
        H = (x2 * Z1Z1 - X1) % P
        r = (y2 * Z1 * Z1Z1 - Y1) % P
        if H == 0:
            return self._jdbl(X1, Y1, Z1) if r == 0 else (1, 1, 0)
        HH = H * H % P
        HHH = H * HH % P
        V = X1 * HH % P
        X3 = (r * r - HHH - 2 * V) % P
        return (X3, (r * (V - X3) - Y1 * HHH) % P, Z1 * H % P)

    def mul(self, pt, k):
        """k * pt by MSB-first double-and-add (ark-ec 0.3 `AffineCurve::mul`);
        k is reduced modulo the subgroup order first (all test points lie in the subgroup)."""
        k %= self.N
        if pt is None or k == 0:
            return None
        acc = (1, 1, 0)
        for bit in bin(k)[2:]:
            acc = self._jdbl(*acc)
            if bit == "1":
                acc = self._jmadd(*acc, pt[0], pt[1])
        X, Y, Z = acc
        if Z == 0:
            return None
        zi = pow(Z, -1, self.P)
        return (X * zi * zi % self.P, Y * zi * zi * zi % self.P)

    def msm(self, points, scalars):
        """sum_i scalars[i] * points[i], straight from the definition."""
        assert len(points) == len(scalars)
        acc = None
        for pt, k in zip(points, scalars):
            acc = self.add(acc, self.mul(pt, k))
        return acc

    # ------------------------------------------------------------------------- bytes
    def fe_to_bytes(self, a):
        return int(a).to_bytes(self.fe_bytes, "little")

    def scalar_to_bytes(self, k):
        return int(k).to_bytes(self.scalar_bytes, "little")

    def point_to_bytes(self, pt):
        """C-ABI layout: x || y, identity = all-zero bytes."""
        if pt is None:
            return bytes(2 * self.fe_bytes)
        return self.fe_to_bytes(pt[0]) + self.fe_to_bytes(pt[1])

    def point_from_bytes(self, b):
        assert len(b) == 2 * self.fe_bytes
        if b == bytes(2 * self.fe_bytes):
            return None
        return (int.from_bytes(b[:self.fe_bytes], "little"), int.from_bytes(b[self.fe_bytes:], "little"))

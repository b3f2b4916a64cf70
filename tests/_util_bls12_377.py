"""Synthetic inputs over BLS12-377 G1 (SURVEY.md section 8(d) recipe, second curve): chain points with known
discrete logs, scalars uniform below the group order from the seeded ChaCha20 stream."""
from oracle.py import bls12_377 as bls
from oracle.py.transcript import ChaCha20Rng

pb = bls.point_to_bytes
b32 = bls.scalar_to_bytes


class Stream:
    """ChaCha20 keyed by a u64 seed; scalar() = uniform in [0, r) by 253-bit mask-and-reject."""

    def __init__(self, seed):
        self.rng = ChaCha20Rng(int(seed).to_bytes(8, "little") + bytes(24))

    def scalar(self):
        while True:
            limbs = [self.rng.next_u64() for _ in range(4)]
            v = (limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)) & ((1 << 253) - 1)
            if v < bls.N:
                return v

    def below(self, bound):
        lim = (1 << 64) - ((1 << 64) % bound)
        while True:
            v = self.rng.next_u64()
            if v < lim:
                return v % bound


def chain_points(n, seed):
    """P_i = (s0 + i*s1)*G: distinct subgroup points, sum k_i*P_i = (sum k_i*(s0+i*s1))*G at any size."""
    st = Stream(seed)
    s0, s1 = st.scalar(), st.scalar()
    cur, step = bls.mul(bls.G, s0), bls.mul(bls.G, s1)
    pts = []
    for _ in range(n):
        pts.append(cur)
        cur = bls.add(cur, step)
    return s0, s1, pts, st


def scalars(st, n, kind="uniform"):
    if kind == "uniform":
        return [st.scalar() for _ in range(n)]
    if kind == "zero":
        return [0] * n
    if kind == "max":
        return [bls.N - 1] * n
    if kind == "small":
        return [st.below(1 << 16) for _ in range(n)]
    if kind == "same":
        return [st.scalar()] * n
    raise ValueError(kind)


def instance(m, n, seed):
    """A seeded shuffle instance over BLS12-377 G1 (same recipe as _util.instance); call it inside
    `with bayer_groth.curve("bls12_377")`."""
    from oracle.py import bayer_groth as bg
    N = m * n
    s0, s1, pts, st = chain_points(n + 3 + 2 * N, seed)
    pp = bg.Params(m, n, bls.G, pts[:n], pts[n], pts[n + 1])
    pk = pts[n + 2]
    deck = [(pts[n + 3 + 2 * i], pts[n + 4 + 2 * i]) for i in range(N)]
    perm = list(range(N))
    for i in range(N - 1, 0, -1):  # Fisher-Yates from the same stream
        j = st.below(i + 1)
        perm[i], perm[j] = perm[j], perm[i]
    rho = [st.scalar() for _ in range(N)]
    rnd = [st.scalar() for _ in range(bg.prover_randomness_len(m, n))]
    return pp, pk, deck, perm, rho, rnd

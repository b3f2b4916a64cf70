"""Shared helpers for the parity tests (synthetic inputs of SURVEY.md section 8(d))."""
from oracle.py import stark
from oracle.py.transcript import SeededStream

b32 = stark.fe_to_bytes
pb = stark.point_to_bytes64


def chain_points(n, seed):
    """P_i = (s0 + i*s1)*G from two seeded scalars: distinct, on-curve, and with known
    discrete logs, so sum k_i*P_i = (sum k_i*(s0+i*s1))*G is checkable at any size."""
    st = SeededStream(seed)
    s0, s1 = st.scalar(), st.scalar()
    cur, step = stark.mul(stark.G, s0), stark.mul(stark.G, s1)
    pts = []
    for _ in range(n):
        pts.append(cur)
        cur = stark.add(cur, step)
    return s0, s1, pts, st


def scalars(st, n, kind="uniform"):
    if kind == "uniform":
        return [st.scalar() for _ in range(n)]
    if kind == "zero":
        return [0] * n
    if kind == "max":
        return [stark.N - 1] * n
    if kind == "small":
        return [st.below(1 << 16) for _ in range(n)]
    if kind == "same":
        return [st.scalar()] * n
    raise ValueError(kind)


def instance(m, n, seed):
    """A seeded shuffle instance: params, pk, deck, permutation, masking factors, prover
    randomness.  Points are chain points (cheap), scalars from the seeded stream."""
    from oracle.py import bayer_groth as bg
    N = m * n
    s0, s1, pts, st = chain_points(n + 3 + 2 * N, seed)
    pp = bg.Params(m, n, stark.G, pts[:n], pts[n], pts[n + 1])
    pk = pts[n + 2]
    deck = [(pts[n + 3 + 2 * i], pts[n + 4 + 2 * i]) for i in range(N)]
    perm = st.permutation(N)
    rho = [st.scalar() for _ in range(N)]
    rnd = [st.scalar() for _ in range(bg.prover_randomness_len(m, n))]
    return pp, pk, deck, perm, rho, rnd

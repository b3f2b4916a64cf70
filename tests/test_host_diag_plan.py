"""CPU test of the host-side plan for the prover's diagonal ciphertext products
(csrc/diag_plan.hpp, compiled with g++): the Karatsuba leaves and the signed contribution lists must
reproduce  E_k = sum_{m+j-i=k} <C_i, A_j>  for every k.  The bilinear map <points, scalars> is
modelled over the integers mod a prime (any bilinear map obeys the same identity)."""
import ctypes
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = (1 << 61) - 1


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "host_shim.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_shim.cpp")])
    return ctypes.CDLL(out)


def plan(shim, m):
    sizes = (ctypes.c_uint32 * 2)()
    shim.h_diag_plan(m, sizes, None, None, None, None, None)
    nleaf, nent = sizes[0], sizes[1]
    mask, val = (ctypes.c_uint32 * nleaf)(), (ctypes.c_uint32 * nleaf)()
    single, rows, ent = (ctypes.c_uint32 * m)(), (ctypes.c_uint32 * (2 * m + 1))(), (ctypes.c_uint32 * nent)()
    shim.h_diag_plan(m, sizes, mask, val, single, rows, ent)
    return list(mask), list(val), list(single), list(rows), list(ent)


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 7, 8, 13, 16, 32])
def test_plan_reproduces_the_diagonals(shim, m):
    rng = random.Random(m)
    n = 3
    C = [[rng.randrange(Q) for _ in range(n)] for _ in range(m)]        # C[i-1] = chunk i
    A = [[rng.randrange(Q) for _ in range(n)] for _ in range(m + 1)]    # A[0] = a0, A[j] = b_j
    dot = lambda p, s: sum(a * b for a, b in zip(p, s)) % Q
    want = [0] * (2 * m)
    for i in range(1, m + 1):
        for j in range(m + 1):
            want[m + j - i] = (want[m + j - i] + dot(C[i - 1], A[j])) % Q
    mask, val, single, rows, ent = plan(shim, m)
    nleaf = len(mask)
    assert rows[0] == 0 and rows[-1] == len(ent) and len(rows) == 2 * m + 1
    levels = max(m - 1, 0).bit_length()
    assert nleaf <= 3 ** levels
    # leaf rows: sums over U = {u < m : u & mask == val};  P_u = C_{m-u},  S_v = A_{v+1}
    res = []
    for l in range(nleaf):
        U = [u for u in range(m) if (u & mask[l]) == val[l]]
        assert U, "plan must drop empty leaves"
        P = [sum(C[m - u - 1][c] for u in U) % Q for c in range(n)]
        S = [sum(A[u + 1][c] for u in U) % Q for c in range(n)]
        res.append(dot(P, S))
    for u in range(m):
        assert mask[single[u]] == (1 << levels) - 1 and val[single[u]] == u
    for i in range(1, m + 1):                                            # the blinding-row jobs
        res.append(dot(C[i - 1], A[0]))
    for k in range(2 * m):
        acc = 0
        for e in ent[rows[k]:rows[k + 1]]:
            acc += -res[e & 0x7fffffff] if e >> 31 else res[e & 0x7fffffff]
        assert acc % Q == want[k], f"E_{k} differs at m={m}"


def test_leaf_count_at_the_headline_size(shim):
    mask, val, single, rows, ent = plan(shim, 128)
    # every leaf expands to 2^(fixed digits) signed contributions: 5^7 in all, plus the m blinding-row jobs
    assert len(mask) == 3 ** 7 and len(ent) == 5 ** 7 + 128 and rows[-1] == len(ent)

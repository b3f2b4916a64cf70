"""Second curve: the word-level algorithms of csrc/fq_bls12_377.cuh + ec.cuh
(12-limb Montgomery product over carry-chain primitives, word-serial reduction with q = 1 mod 2^32, lazy
bounds, XYZZ formulas with a = 0) compiled for the HOST with g++ and checked against the big-int oracle."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle.py import bls12_377 as bls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q, R = bls.Q, bls.N
M384 = 1 << 384


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim377") / "host_shim_bls12_377.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMP_CURVE_BLS12_377", "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_shim_bls12_377.cpp")])
    return ctypes.CDLL(out)


def w(x, n=12):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def rd(buf, n=12):
    return sum(int(buf[i]) << (32 * i) for i in range(n))


def pw(pt):
    b = bls.point_to_bytes(pt)
    return (ctypes.c_uint32 * 24)(*[int.from_bytes(b[4 * i:4 * i + 4], "little") for i in range(24)])


def test_constants():
    x = bls.X_PARAM
    assert R == x ** 4 - x ** 2 + 1 and Q == (x - 1) ** 2 * R // 3 + x
    assert Q % (1 << 46) == 1 and (-pow(Q, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF
    assert (1 << 32) // ((Q >> 352) + 1) == 152 and 152 * Q < M384 < 153 * Q


def test_fq_mul_lazy_bounds(shim):
    rnd = random.Random(3)
    Rinv = pow(M384, -1, Q)
    cases = [(0, 0), (1, 1), (Q - 1, Q - 1), (Q, 2 * Q), (5 * Q - 1, 6 * Q - 1), (15 * Q, 2 * Q - 1),
             (M384 - 1, 0), ((1 << 380) - 1, (1 << 380) - 1), (30 * Q - 1, Q - 1)]
    cases += [(rnd.randrange(5 * Q), rnd.randrange(6 * Q)) for _ in range(3000)]
    # words of all ones / sparse words exercise every carry path of the chains
    cases += [(int("ffffffff" * k + "00000000" * (11 - k) + "00ffffff", 16) % (8 * Q), (1 << (32 * j)) - 1)
              for k in range(11) for j in range(1, 12)]
    out = (ctypes.c_uint32 * 12)()
    for a, b in cases:
        shim.h_fq_mul(w(a), w(b), out)
        r = rd(out)
        assert r % Q == a * b * Rinv % Q, (hex(a), hex(b))
        if a * b <= 30 * Q * Q:
            assert r < 2 * Q
    for a, _ in cases[:2000]:
        if a >= 5 * Q:  # outside the contract of fq_sqr ([x]^2 with x*x <= 30)
            continue
        shim.h_fq_sqr(w(a), out)
        assert rd(out) % Q == a * a * Rinv % Q


def test_reductions_and_sub(shim):
    rnd = random.Random(4)
    out = (ctypes.c_uint32 * 12)()
    for v in [0, Q, 2 * Q, M384 - 1, 152 * Q, 152 * Q - 1, 31 * Q] + [rnd.randrange(M384) for _ in range(3000)]:
        shim.h_fq_reduce_weak(w(v), out)
        r = rd(out)
        assert r < 2 * Q and r % Q == v % Q
        shim.h_fq_reduce_full(w(v), out)
        assert rd(out) == v % Q
    for kb in (2, 4, 6):
        for _ in range(300):
            a, b = rnd.randrange(2 * Q), rnd.randrange(kb * Q)
            shim.h_fq_sub(w(a), w(b), kb, out)
            assert rd(out) == a + kb * Q - b
    assert shim.h_fq_is_zero_mod_p_2(w(0)) == 1 and shim.h_fq_is_zero_mod_p_2(w(Q)) == 1
    assert shim.h_fq_is_zero_mod_p_2(w(1)) == 0 and shim.h_fq_is_zero_mod_p_2(w(Q + 1)) == 0


def test_inverse(shim):
    rnd = random.Random(5)
    out = (ctypes.c_uint32 * 12)()
    for v in [1, 2, Q - 1] + [rnd.randrange(1, Q) for _ in range(10)]:
        shim.h_fq_inv_canonical(w(v), out)
        assert rd(out) * v % Q == 1


def test_point_ops(shim):
    rnd = random.Random(6)
    pts = [bls.mul(bls.G, rnd.randrange(1, R)) for _ in range(5)]
    out = (ctypes.c_uint32 * 24)()
    assert shim.h_on_curve(pw(bls.G)) == 1
    assert shim.h_on_curve(pw((bls.G[0], bls.G[1] + 1))) == 0
    for a in pts:
        assert shim.h_on_curve(pw(a)) == 1
        for b in pts + [a, bls.neg(a), None]:
            shim.h_point_add(pw(a), pw(b), out)
            assert bytes(out) == bls.point_to_bytes(bls.add(a, b))
        shim.h_point_add(pw(None), pw(a), out)
        assert bytes(out) == bls.point_to_bytes(a)
        shim.h_dbl_affine(pw(a), out)
        assert bytes(out) == bls.point_to_bytes(bls.add(a, a))


def test_scalar_mul_and_general_add(shim):
    rnd = random.Random(7)
    out = (ctypes.c_uint32 * 24)()
    p, q = bls.mul(bls.G, 7), bls.mul(bls.G, rnd.randrange(1, R))
    for k in [0, 1, 2, R - 1, R, R + 1, rnd.randrange(1 << 256), rnd.randrange(R)]:
        shim.h_scalar_mul(pw(q), w(k, 8), out)
        assert bytes(out) == bls.point_to_bytes(bls.mul(q, k))
    for k1, k2 in [(3, 5), (rnd.randrange(R), rnd.randrange(R)), (5, R - 5), (0, 9)]:
        shim.h_lincomb2(pw(p), w(k1, 8), pw(q), w(k2, 8), out)
        assert bytes(out) == bls.point_to_bytes(bls.add(bls.mul(p, k1), bls.mul(q, k2)))
    # equal and opposite operands through the general addition
    shim.h_lincomb2(pw(p), w(6, 8), pw(p), w(6, 8), out)
    assert bytes(out) == bls.point_to_bytes(bls.mul(p, 12))
    shim.h_lincomb2(pw(p), w(6, 8), pw(p), w(R - 6, 8), out)
    assert bytes(out) == bls.point_to_bytes(None)

"""Property tests (hypothesis) of the 12-limb field code of csrc/fq_bls12_377.cuh compiled for the host: the
product over the carry-chain primitives, the parity-split word-serial reduction and the lazy-bound contract, on
edge-heavy distributions (values near 0, q, k*q and 2^384, sparse / all-ones limbs), plus the algebraic laws the
group formulas rely on."""
import ctypes
import os
import subprocess

import pytest
from hypothesis import given, settings, strategies as st

from oracle.py import bls12_377 as bls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = bls.P
R = 1 << 384
RINV = pow(R, -1, Q)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim377p") / "host_shim_bls12_377.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMP_CURVE_BLS12_377", "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_shim_bls12_377.cpp")])
    return ctypes.CDLL(out)


def w(x):
    return (ctypes.c_uint32 * 12)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(12)])


def rd(buf):
    return sum(int(buf[i]) << (32 * i) for i in range(12))


def edgy(bound):
    near = st.sampled_from([0, 1, 2, Q - 1, Q, Q + 1, 2 * Q - 1, 2 * Q, 30 * Q - 1, 152 * Q, (1 << 377) - 1, 1 << 377, R - 1])
    sparse = st.lists(st.sampled_from([0, 1, 0xFFFFFFFF, 0x80000000, 0xFFFFFFFE]), min_size=12, max_size=12).map(
        lambda ls: sum(l << (32 * i) for i, l in enumerate(ls)))
    return st.one_of(near, sparse, st.integers(0, R - 1)).map(lambda v: v % bound)


@settings(max_examples=500, deadline=None)
@given(a=edgy(5 * Q), b=edgy(6 * Q))
def test_fq_mul_any_lazy_operands(shim, a, b):
    out = (ctypes.c_uint32 * 12)()
    shim.h_fq_mul(w(a), w(b), out)
    r = rd(out)
    assert r < 2 * Q and r % Q == a * b * RINV % Q


@settings(max_examples=300, deadline=None)
@given(a=edgy(30 * Q), b=edgy(Q))
def test_fq_mul_wide_first_operand(shim, a, b):
    # the contract is on the PRODUCT of the bounds: [30] x [1] is as admissible as [5] x [6]
    out = (ctypes.c_uint32 * 12)()
    shim.h_fq_mul(w(a), w(b), out)
    r = rd(out)
    assert r < 2 * Q and r % Q == a * b * RINV % Q


@settings(max_examples=500, deadline=None)
@given(v=edgy(R))
def test_fq_reductions(shim, v):
    out = (ctypes.c_uint32 * 12)()
    shim.h_fq_reduce_weak(w(v), out)
    assert rd(out) < 2 * Q and rd(out) % Q == v % Q
    shim.h_fq_reduce_full(w(v), out)
    assert rd(out) == v % Q


@settings(max_examples=200, deadline=None)
@given(a=edgy(2 * Q), b=edgy(2 * Q), c=edgy(2 * Q))
def test_ring_laws_in_montgomery_form(shim, a, b, c):
    ab, ba, bc, ab_c, a_bc = [(ctypes.c_uint32 * 12)() for _ in range(5)]
    shim.h_fq_mul(w(a), w(b), ab)
    shim.h_fq_mul(w(b), w(a), ba)
    assert rd(ab) % Q == rd(ba) % Q
    shim.h_fq_mul(w(b), w(c), bc)
    shim.h_fq_mul(ab, w(c), ab_c)
    shim.h_fq_mul(w(a), bc, a_bc)
    assert rd(ab_c) % Q == rd(a_bc) % Q
    # distributivity through the lazy subtraction: a*(b - c) = a*b - a*c
    d, ad, ac, diff = [(ctypes.c_uint32 * 12)() for _ in range(4)]
    shim.h_fq_sub(w(b), w(c), 2, d)            # b + 2q - c, a [4] value
    shim.h_fq_mul(w(a), d, ad)
    shim.h_fq_mul(w(a), w(c), ac)
    shim.h_fq_sub(ab, ac, 2, diff)
    assert rd(ad) % Q == rd(diff) % Q


@settings(max_examples=40, deadline=None)
@given(v=edgy(Q))
def test_inverse(shim, v):
    out = (ctypes.c_uint32 * 12)()
    shim.h_fq_inv_canonical(w(v), out)
    assert rd(out) == (pow(v, -1, Q) if v else 0)

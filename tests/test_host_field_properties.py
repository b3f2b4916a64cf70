"""Property tests (hypothesis) of the word-level field code compiled for the host: the same source
nvcc compiles for the device (csrc/fq.cuh, fr.cuh).  Edge-heavy distributions: values near 0, p,
2^252 and 2^256, sparse limbs."""
import ctypes
import os
import subprocess

import pytest
from hypothesis import given, settings, strategies as st

from oracle.py import stark

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, N = stark.P, stark.N
R = 1 << 256


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "host_shim.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_shim.cpp")])
    return ctypes.CDLL(out)


def w(x):
    return (ctypes.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def rd(buf):
    return sum(int(buf[i]) << (32 * i) for i in range(8))


def edgy(bound):
    """integers in [0, bound) biased towards the edges and towards sparse limb patterns"""
    near = st.sampled_from([0, 1, 2, P - 1, P, P + 1, 2 * P, N - 1, N, (1 << 252) - 1, 1 << 252, R - 1])
    sparse = st.lists(st.sampled_from([0, 1, 0xFFFFFFFF, 0x80000000]), min_size=8, max_size=8).map(
        lambda ls: sum(l << (32 * i) for i, l in enumerate(ls)))
    return st.one_of(near, sparse, st.integers(0, R - 1)).map(lambda v: v % bound)


@settings(max_examples=400, deadline=None)
@given(a=edgy(5 * P), b=edgy(6 * P))
def test_fq_mul_any_lazy_operands(shim, a, b):
    out = (ctypes.c_uint32 * 8)()
    shim.h_fq_mul(w(a), w(b), out)
    r = rd(out)
    assert r < 2 * P and r % P == a * b * pow(R, -1, P) % P


@settings(max_examples=400, deadline=None)
@given(v=edgy(R))
def test_fq_reductions(shim, v):
    out = (ctypes.c_uint32 * 8)()
    shim.h_fq_reduce_weak(w(v), out)
    assert rd(out) < (1 << 252) and rd(out) % P == v % P
    shim.h_fq_reduce_full(w(v), out)
    assert rd(out) == v % P


@settings(max_examples=300, deadline=None)
@given(a=edgy(R), b=edgy(N))
def test_fr_ring_ops(shim, a, b):
    out, s, d, ng = [(ctypes.c_uint32 * 8)() for _ in range(4)]
    shim.h_fr_mul_canonical(w(a), w(b), out)     # first operand may be unreduced (< 2^256)
    assert rd(out) == a * b % N
    shim.h_fr_addsub_canonical(w(a % N), w(b), s, d, ng)
    assert rd(s) == (a + b) % N and rd(d) == (a - b) % N and rd(ng) == (-a) % N


@settings(max_examples=40, deadline=None)
@given(k=st.integers(1, N - 1), l=st.integers(1, N - 1))
def test_group_law_random_pairs(shim, k, l):
    p, q = stark.mul(stark.G, k), stark.mul(stark.G, l)
    out = (ctypes.c_uint32 * 16)()

    def pw(pt):
        b = stark.point_to_bytes64(pt)
        return (ctypes.c_uint32 * 16)(*[int.from_bytes(b[4 * i:4 * i + 4], "little") for i in range(16)])

    shim.h_point_add(pw(p), pw(q), out)
    assert bytes(out) == stark.point_to_bytes64(stark.add(p, q))

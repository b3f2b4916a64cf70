"""Word-level algorithms of csrc/fq.cuh + ec.cuh (sparse-prime Montgomery reduction, lazy
bounds, XYZZ formulas) compiled for the HOST with g++ and checked against the oracle.  The
same source is what nvcc compiles for the device; this catches logic errors without a GPU."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle.py import stark

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, N = stark.P, stark.N


@pytest.fixture(scope="module", params=["default", "avx2_blake2s", "scalar_blake2s"])
def shim(request, tmp_path_factory):
    # the default build picks the compression function at run time on x86 hosts (AVX-512VL multi-block
    # form, else the AVX2 row formulation); "avx2_blake2s" rules out the first, "scalar_blake2s"
    # forces the portable code
    out = str(tmp_path_factory.mktemp("shim") / f"host_shim_{request.param}.so")
    flags = {"scalar_blake2s": ["-DMP_BLAKE2S_FORCE_SCALAR"], "avx2_blake2s": ["-DMP_BLAKE2S_NO_AVX512"]}.get(request.param, [])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", *flags, "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_shim.cpp")])
    return ctypes.CDLL(out)


def w(x):
    return (ctypes.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def rd(buf, n=8):
    return sum(int(buf[i]) << (32 * i) for i in range(n))


def pw(pt):
    b = stark.point_to_bytes64(pt)
    return (ctypes.c_uint32 * 16)(*[int.from_bytes(b[4 * i:4 * i + 4], "little") for i in range(16)])


def test_fq_mul_lazy_bounds(shim):
    rnd = random.Random(3)
    Rinv = pow(1 << 256, -1, P)
    cases = [(0, 0), (1, 1), (P - 1, P - 1), (P, 2 * P), (5 * P - 1, 6 * P - 1), (15 * P, 2 * P - 1)]
    cases += [(rnd.randrange(5 * P), rnd.randrange(6 * P)) for _ in range(2000)]
    out = (ctypes.c_uint32 * 8)()
    for a, b in cases:
        shim.h_fq_mul(w(a), w(b), out)
        r = rd(out)
        assert r < 2 * P and r % P == a * b * Rinv % P


def test_reductions(shim):
    rnd = random.Random(4)
    out = (ctypes.c_uint32 * 8)()
    for v in [0, P, 2 * P, (1 << 256) - 1, 31 * P] + [rnd.randrange(1 << 256) for _ in range(2000)]:
        shim.h_fq_reduce_weak(w(v), out)
        r = rd(out)
        assert r < (1 << 252) and r % P == v % P
        shim.h_fq_reduce_full(w(v), out)
        assert rd(out) == v % P


def test_inverse(shim):
    rnd = random.Random(5)
    out = (ctypes.c_uint32 * 8)()
    for v in [1, 2, P - 1] + [rnd.randrange(1, P) for _ in range(20)]:
        shim.h_fq_inv_canonical(w(v), out)
        assert rd(out) * v % P == 1


def test_point_ops(shim):
    rnd = random.Random(6)
    pts = [stark.mul(stark.G, rnd.randrange(1, N)) for _ in range(6)]
    out = (ctypes.c_uint32 * 16)()
    for a in pts:
        assert shim.h_on_curve(pw(a)) == 1
        for b in pts + [a, stark.neg(a), None]:
            shim.h_point_add(pw(a), pw(b), out)
            assert bytes(out) == stark.point_to_bytes64(stark.add(a, b))
    bad = (ctypes.c_uint32 * 16)(*([5] + [0] * 7 + [7] + [0] * 7))
    assert shim.h_on_curve(bad) == 0
    for k in [0, 1, 2, N - 1, N, rnd.randrange(1 << 256), rnd.randrange(N)]:
        shim.h_scalar_mul(pw(pts[0]), w(k), out)
        assert bytes(out) == stark.point_to_bytes64(stark.mul(pts[0], k))
    k1, k2 = rnd.randrange(N), rnd.randrange(N)
    for q in (pts[1], pts[0]):  # distinct points, and the same point (general add -> doubling branch when k1 == k2)
        for kk2 in (k2, k1, N - k1):
            shim.h_lincomb2(pw(pts[0]), w(k1), pw(q), w(kk2), out)
            assert bytes(out) == stark.point_to_bytes64(stark.add(stark.mul(pts[0], k1), stark.mul(q, kk2)))


def test_fr_arithmetic(shim):
    rnd = random.Random(8)
    out, s, d, ng = [(ctypes.c_uint32 * 8)() for _ in range(4)]
    cases = [(0, 0), (1, N - 1), (N - 1, N - 1), (N, 5), ((1 << 256) - 1, 3)]
    cases += [(rnd.randrange(N), rnd.randrange(N)) for _ in range(1000)]
    for a, b in cases:
        shim.h_fr_mul_canonical(w(a), w(b), out)
        assert rd(out) == a * b % N
        shim.h_fr_addsub_canonical(w(a), w(b), s, d, ng)
        assert rd(s) == (a + b) % N and rd(d) == (a - b) % N and rd(ng) == (-a) % N


def test_host_transcript_matches_oracle(shim):
    import hashlib
    from oracle.py.transcript import FiatShamirRng
    shim.h_blake2s.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p]
    shim.h_fs_challenges.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_char_p]
    shim.h_fs_points_challenge.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
    out = ctypes.create_string_buffer(32)
    rnd = random.Random(9)
    for ln in [0, 1, 63, 64, 65, 127, 128, 129, 191, 192, 193, 1000, 4096, 4097]:
        msg = bytes(rnd.randrange(256) for _ in range(ln))
        for split in [0, 1, 64, ln // 2, ln]:
            shim.h_blake2s(msg, ln, split, out)
            assert out.raw == hashlib.blake2s(msg).digest(), (ln, split)
    for data in [b"", b"abc", bytes(range(200))]:
        buf = ctypes.create_string_buffer(32 * 4)
        shim.h_fs_challenges(data, len(data), 4, buf)
        fs = FiatShamirRng()
        if data:
            fs.absorb(data)
        assert buf.raw == b"".join(stark.fe_to_bytes(fs.challenge()) for _ in range(4))
    pts = [stark.mul(stark.G, 5), None, stark.mul(stark.G, 7)]
    shim.h_fs_points_challenge(b"".join(stark.point_to_bytes64(p) for p in pts), 3, out)
    fs = FiatShamirRng()
    fs.absorb(b"label" + b"".join(stark.point_to_bytes65(p) for p in pts))
    assert out.raw == stark.fe_to_bytes(fs.challenge())


def test_multi_stream_blake2s_equals_single_streams(shim, tmp_path):
    """Blake2sLanes (csrc/transcript.hpp: up to 8 equal-length streams hashed in lockstep, AVX2 / AVX-512VL) against
    hashlib for every lane count, lengths around the block and buffering boundaries, odd piece sizes, and a hand-over to
    the single-stream hasher in the middle of a stream; the scalar fallback through MP_BLAKE2S_LANES_SCALAR in a
    subprocess-free way is the same code path as lanes == 1 on a CPU without AVX2, so it is driven here by forcing it."""
    import hashlib, random
    rnd = random.Random(9)
    shim.h_blake2s_lanes.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64,
                                     ctypes.c_uint64, ctypes.c_char_p]
    stride = 5000
    data = bytes(rnd.getrandbits(8) for _ in range(8 * stride))
    out = ctypes.create_string_buffer(8 * 32)
    for lanes in range(1, 9):
        for ln in (0, 1, 63, 64, 65, 127, 128, 129, 1000, 4096, 4999):
            for piece, cut in ((64, ln), (1, ln), (37, ln), (4160, ln), (5000, ln), (200, ln // 2), (64, 64), (4160, 0)):
                shim.h_blake2s_lanes(data, stride, ln, lanes, max(piece, 1), min(cut, ln), out)
                for l in range(lanes):
                    want = hashlib.blake2s(data[l * stride:l * stride + ln]).digest()
                    assert out.raw[32 * l:32 * l + 32] == want, (lanes, ln, piece, cut, l)


def test_transcript_lanes_hand_over(shim):
    """TranscriptLanes: the first part of an absorb hashed for several transcripts at once, then handed to ordinary
    transcripts that finish it -- the challenges equal those of transcripts that absorbed everything themselves."""
    import random
    rnd = random.Random(10)
    pts = lambda k: bytes(rnd.getrandbits(8) for _ in range(64 * k))
    for lanes, n_shared, n_lane, n_tail in ((1, 3, 5, 2), (3, 70, 129, 1), (8, 515, 300, 4), (8, 1, 64, 0), (5, 0, 1, 1)):
        shared, lane, tail = pts(n_shared), pts(n_lane * lanes), pts(n_tail)
        if n_lane > 2:                                 # an identity point in lane 0 (all-zero bytes -> (0, 1, infinity))
            lane = bytes(64) + lane[64:]
        got = ctypes.create_string_buffer(32 * lanes)
        shim.h_fs_lanes_challenges(shared, ctypes.c_uint64(n_shared), lane, ctypes.c_uint64(n_lane), lanes, tail,
                                   ctypes.c_uint64(n_tail), got)
        for l in range(lanes):
            want = ctypes.create_string_buffer(32)
            shim.h_fs_single_challenge(shared, ctypes.c_uint64(n_shared), lane[64 * n_lane * l:64 * n_lane * (l + 1)],
                                       ctypes.c_uint64(n_lane), tail, ctypes.c_uint64(n_tail), want)
            assert got.raw[32 * l:32 * l + 32] == want.raw, (lanes, l)

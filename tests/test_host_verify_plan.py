"""CPU test of the HOST half of the GPU verifier (csrc/shuffle_host.hpp, compiled with g++): the
transcript schedule, the rewriting of every verifier check into "sum scalar*point == O" jobs and
the verdict order -- evaluated here with the Python oracle's group arithmetic instead of the GPU.
For a valid proof every job must sum to the identity; for tampered proofs the verdict must be the
oracle's."""
import ctypes
import json
import os
import subprocess

import pytest

from oracle.py import stark as _stark, bayer_groth as bg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDS = {"stark": json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json"))),
         "bls12_377": json.load(open(os.path.join(ROOT, "tests", "golden", "bls12_377_shuffle_vectors.json")))}
h = bytes.fromhex
# the module-level names below are rebound per curve by the `shim` fixture: the same test bodies run the host plan of
# the Stark instantiation and of the BLS12-377 one (csrc/shuffle_host.hpp compiled with -DMP_CURVE_BLS12_377)
stark, GOLD, PB = _stark, GOLDS["stark"], 64


@pytest.fixture(scope="module", params=["stark", "bls12_377"])
def shim(request, tmp_path_factory):
    global stark, GOLD, PB
    name = request.param
    out = str(tmp_path_factory.mktemp("shim_" + name) / "host_shim.so")
    src, flags = ("host_shim.cpp", []) if name == "stark" else ("host_shim_bls12_377.cpp", ["-DMP_CURVE_BLS12_377"])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", *flags, "-o", out,
                           os.path.join(ROOT, "tests", "host", src)])
    with bg.curve(name) as grp:
        stark, GOLD, PB = grp, GOLDS[name], 64 if name == "stark" else 96
        lib = ctypes.CDLL(out)
        if name != "stark":
            assert lib.h_point_bytes() == 96
        yield lib
    stark, GOLD, PB = _stark, GOLDS["stark"], 64


def pts(buf):
    return [stark.point_from_bytes64(buf[PB * i:PB * i + PB]) for i in range(len(buf) // PB)]


def scs(buf):
    return [int.from_bytes(buf[32 * i:32 * i + 32], "little") for i in range(len(buf) // 32)]


def run_plan(shim, fx, deck2=None, proof=None):
    m, n = fx["m"], fx["n"]
    N = m * n
    ck_g = h(fx["ck_g"])
    gsum = None
    for p in pts(ck_g):
        gsum = stark.add(gsum, p)
    T1 = 8 * m + 5 * n + 19
    g1_pts, g1_scal = ctypes.create_string_buffer(T1 * PB), ctypes.create_string_buffer(T1 * 32)
    lens, flags = (ctypes.c_int * 8)(), (ctypes.c_int * 5)()
    sx, s2 = ctypes.create_string_buffer(N * 32), ctypes.create_string_buffer(N * 32)
    ss, small = ctypes.create_string_buffer((2 * m + 3) * 32), ctypes.create_string_buffer((2 * m + 3) * 2 * PB)
    deck, deck2 = h(fx["deck"]), deck2 or h(fx["deck2"])
    cnt = shim.h_verify_plan(m, n, h(fx["enc_g"]), ck_g, h(fx["ck_h"]), h(fx["ghat"]), stark.point_to_bytes64(gsum),
                             h(fx["pk"]), deck, deck2, proof or h(fx["proof"]), g1_pts, g1_scal, lens, sx, s2, ss, small, flags)
    assert cnt == T1 and sum(lens) == T1
    # evaluate the eight G1 jobs with the oracle
    P, K, ids, off = pts(g1_pts.raw), scs(g1_scal.raw), [], 0
    for ln in lens:
        ids.append(int(stark.msm(P[off:off + ln], K[off:off + ln]) is stark.INF))
        off += ln
    # and the two ciphertext equations, component-wise
    D, D2, SM = pts(deck), pts(deck2), pts(small.raw)
    kx, k2, ks = scs(sx.raw), scs(s2.raw), scs(ss.raw)
    ct_ok = True
    for comp in (0, 1):
        e0 = stark.add(stark.msm(D[comp::2], kx), stark.mul(SM[comp], ks[0]))
        e1 = stark.add(stark.msm(D2[comp::2], k2), stark.msm(SM[2 + comp::2], ks[1:]))
        ct_ok = ct_ok and e0 is stark.INF and e1 is stark.INF
    return shim.h_verdict((ctypes.c_int * 8)(*ids), int(ct_ok), flags), ids, ct_ok, list(flags)


@pytest.mark.parametrize("idx", [0, 1])
def test_valid_proof_all_jobs_are_identity(shim, idx):
    status, ids, ct_ok, flags = run_plan(shim, GOLD["shuffle"][idx])
    assert ids == [1] * 8 and ct_ok and flags == [1] * 5 and status == 0


def test_verdicts_match_oracle_on_bad_inputs(shim):
    fx = GOLD["shuffle"][1]
    m, n = fx["m"], fx["n"]
    pp = bg.Params(m, n, stark.point_from_bytes64(h(fx["enc_g"])), pts(h(fx["ck_g"])), stark.point_from_bytes64(h(fx["ck_h"])),
                   stark.point_from_bytes64(h(fx["ghat"])))
    pk = stark.point_from_bytes64(h(fx["pk"]))
    deck = [tuple(pts(h(fx["deck"]))[2 * i:2 * i + 2]) for i in range(m * n)]
    good2 = h(fx["deck2"])
    # wrong shuffled deck (rotate the cards): reference negative case -> Hadamard
    wrong = good2[2 * PB:] + good2[:2 * PB]
    proof = h(fx["proof"])
    cases = [(wrong, proof)]
    for off in (len(proof) - 32 * 4, len(proof) - 32 * 5):   # multi-exp r, multi-exp a_n (low bytes: the scalars stay canonical)
        p2 = bytearray(proof)
        p2[off] ^= 1
        cases.append((good2, bytes(p2)))
    for d2, pf in cases:
        d2_pts = pts(d2)
        want = bg.shuffle_verify(pp, pk, deck, [tuple(d2_pts[2 * i:2 * i + 2]) for i in range(m * n)],
                                 bg.proof_from_bytes(pf, m, n))
        got, *_ = run_plan(shim, fx, d2, pf)
        assert got == want != 0


def test_contiguous_ciphertext_jobs(shim):
    """`assemble_ct_jobs` (the layout mp377_shuffle_verify hands to the batched MSM): both 2-component jobs sum to the
    identity for a valid proof and not for a rotated output deck."""
    if PB != 96:
        pytest.skip("only the BLS12-377 shim exports h_ct_jobs")
    fx = GOLD["shuffle"][1]
    m, n = fx["m"], fx["n"]
    N = m * n
    nct = 2 * N + 2 * m + 3
    good2 = h(fx["deck2"])
    for d2, expect in ((good2, True), (good2[2 * PB:] + good2[:2 * PB], False)):
        cts, scal = ctypes.create_string_buffer(nct * 2 * PB), ctypes.create_string_buffer(nct * 32)
        jobs = (ctypes.c_uint32 * 6)()
        cnt = shim.h_ct_jobs(m, n, h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]), h(fx["pk"]), h(fx["deck"]), d2,
                             h(fx["proof"]), cts, scal, jobs)
        assert cnt == nct and list(jobs) == [0, 0, N + 1, N + 1, N + 1, N + 2 * m + 2]
        P, K = pts(cts.raw), scs(scal.raw)
        ok = True
        for so, po, ln in (tuple(jobs[:3]), tuple(jobs[3:])):
            for comp in (0, 1):
                acc = stark.msm(P[2 * po + comp:2 * (po + ln):2], K[so:so + ln])
                ok = ok and acc is stark.INF
        assert ok == expect

"""GPU parity at BASELINE.json's full sizes, through size-independent properties: known discrete
logs for the 2^20-term MSM, prove -> verify round trips and the reference's negative case on a
2^16-card deck, window-range split == single launch."""
import ctypes
import os
import random

import numpy as np
import pytest

from oracle import c_oracle
from oracle.py import stark
from _util import b32, pb

pytestmark = pytest.mark.gpu
N = stark.N
G64 = pb(stark.G)


def rand_scalars(rng, k):
    a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8)
    a[:, 31] &= 0x07
    return a.tobytes()


def ints(buf):
    return [int.from_bytes(buf[32 * i:32 * i + 32], "little") for i in range(len(buf) // 32)]


def test_msm_2p20_known_discrete_logs(ctx, pkg):
    import torch
    n = 1 << 20
    rng = np.random.default_rng(3)
    logs_b = rand_scalars(rng, n)               # P_i = l_i * G, generated on the GPU
    pts = ctx.dbg_scalar_mul(G64 * n, logs_b)
    ks_b = rand_scalars(rng, n)
    e = sum(k * l for k, l in zip(ints(ks_b), ints(logs_b))) % N
    want = pb(stark.mul(stark.G, e))
    assert ctx.msm_g1(pts, ks_b, 0) == want
    assert ctx.msm_g1(pts, ks_b, 13) == want    # a different window size gives the same group element
    # window-range split across 8 "ranks" == single launch
    dev = torch.device("cuda:0")
    d_pts = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
    d_sc = torch.frombuffer(bytearray(ks_b), dtype=torch.uint8).to(dev)
    d_out = torch.zeros(64, dtype=torch.uint8, device=dev)
    c, world = 16, 8
    W = pkg.lib.mp_msm_num_windows(c)
    points, scs = b"", b""
    for r, s in pkg.dist.fold_scalars(c, W, world):
        b, e2 = pkg.dist.window_range(W, r, world)
        ctx.msm_g1_windows_device(d_pts.data_ptr(), d_sc.data_ptr(), n, d_out.data_ptr(), c, b, e2 - b)
        ctx.sync()
        points += bytes(d_out.cpu().numpy().tobytes())
        scs += s
    assert ctx.msm_g1(points, scs, 0) == want


def make_big_instance(ctx, m, n, seed):
    """Synthetic instance with every point generated on the GPU (s*G for seeded scalars s)."""
    Nc = m * n
    rng = np.random.default_rng(seed)
    npts = (n + 3) + 2 * Nc
    pts = ctx.dbg_scalar_mul(G64 * npts, rand_scalars(rng, npts))
    P = lambda i: pts[64 * i:64 * (i + 1)]
    return dict(m=m, n=n, N=Nc, ck_g=pts[:64 * n], ck_h=P(n), ghat=P(n + 1), pk=P(n + 2), deck=pts[64 * (n + 3):],
                perm=[int(v) for v in rng.permutation(Nc)], rho=rand_scalars(rng, Nc), rand=rand_scalars(rng, 11 * m + 5 * n))


@pytest.fixture(scope="module")
def headline(ctx):
    """BASELINE.json's headline configuration: one 2^16-card deck at (m, n) = (128, 512), proved once on
    the GPU through the host-buffer C ABI.  Shared by the tests below."""
    inst = make_big_instance(ctx, 128, 512, 5)
    ctx.set_params(128, 512, G64, inst["ck_g"], inst["ck_h"], inst["ghat"])
    inst["deck2"], inst["proof"] = ctx.shuffle_and_remask(inst["pk"], inst["deck"], inst["perm"], inst["rho"], inst["rand"])
    return inst


def oracle_args(inst):
    return (inst["m"], inst["n"], G64, inst["ck_g"], inst["ck_h"], inst["ghat"], inst["pk"])


def test_shuffle_2p16_round_trip_and_negative(ctx, pkg, headline, monkeypatch):
    inst = headline
    m, n, Nc = inst["m"], inst["n"], inst["N"]
    pk, deck, perm, rho, rand, deck2, proof = (inst[k] for k in ("pk", "deck", "perm", "rho", "rand", "deck2", "proof"))
    ctx.set_params(m, n, G64, inst["ck_g"], inst["ck_h"], inst["ghat"])
    assert ctx.verify_shuffle(pk, deck, deck2, proof) == 0
    # remask semantics on a sample of cards: out[i] - deck[perm[i]] == rho_i * (g, pk)
    rho_i = ints(rho)
    for i in random.Random(1).sample(range(Nc), 4):
        src = deck[128 * perm[i]:128 * perm[i] + 128]
        for comp, base in ((0, G64), (1, pk)):
            want = ctx.dbg_point_add(src[64 * comp:64 * comp + 64], ctx.dbg_scalar_mul(base, b32(rho_i[i])))
            assert deck2[128 * i + 64 * comp:128 * i + 64 * comp + 64] == want
    # determinism: same inputs -> same bytes
    deck2b, proofb = ctx.shuffle_and_remask(pk, deck, perm, rho, rand)
    assert deck2b == deck2 and proofb == proof
    # the two forms of the diagonal products (Karatsuba leaves, the default here, and the
    # schoolbook row products over the pre-shifted deck table) give the same proof bytes
    monkeypatch.setenv("MP_DIAG_KARATSUBA", "0")
    deck2c, proofc = ctx.shuffle_and_remask(pk, deck, perm, rho, rand)
    monkeypatch.delenv("MP_DIAG_KARATSUBA")
    assert deck2c == deck2 and proofc == proof
    # reference negative case (tests.rs:213-226): an unrelated output deck fails in the Hadamard argument
    wrong = deck[128:] + deck[:128]
    assert ctx.verify_shuffle(pk, deck, wrong, proof) == 1
    # a single swapped pair of output cards is still a wrong statement
    sw = bytearray(deck2)
    sw[0:128], sw[128:256] = deck2[128:256], deck2[0:128]
    assert ctx.verify_shuffle(pk, deck, bytes(sw), proof) != 0
    # tampered response scalar -> multi-exponentiation argument
    bad = bytearray(proof)
    bad[-1 - 32 * 3] ^= 1
    assert ctx.verify_shuffle(pk, deck, deck2, bytes(bad)) == 4


def test_shuffle_2p16_gpu_proof_against_the_c_oracle(ctx, headline):
    """The headline configuration against the ORACLE, not against the CUDA path itself: the C restatement
    (all host threads, Pippenger for its big sums) accepts the GPU's (128, 512) proof -- which checks every
    one of the 1 416 proof points and 2 569 scalars, the 7-level Karatsuba plan included, through the
    oracle's own transcript and equations -- and both verifiers give the same verdict on the reference's
    negative case (tests.rs:213-226) and on tampered proofs."""
    inst = headline
    co = c_oracle.COracle(threads=os.cpu_count() or 1, msm_mode=1)
    a = oracle_args(inst)
    pk, deck, deck2, proof = inst["pk"], inst["deck"], inst["deck2"], inst["proof"]
    ctx.set_params(inst["m"], inst["n"], G64, inst["ck_g"], inst["ck_h"], inst["ghat"])
    assert deck2 == co.remask(G64, pk, deck, inst["perm"], inst["rho"])
    assert co.verify(*a, deck, deck2, proof) == 0
    wrong = deck[128:] + deck[:128]
    tampered = []
    m, n = inst["m"], inst["n"]
    f1 = 64 * (5 * m + 4)                       # proof layout: points | 2n+3 scalars | 3 points | 2n+2 scalars | 6m+1 points | n+4 scalars
    f2 = f1 + 32 * (2 * n + 3) + 64 * 3
    f3 = f2 + 32 * (2 * n + 2) + 64 * (6 * m + 1)
    assert f3 + 32 * (n + 4) == len(proof)
    for off in (f1 + 7, f1 + 32 * (n + 1) + 3, f2 + 32 * 5 + 1, f2 + 32 * (2 * n + 1), f3 + 32 * 2 + 9, len(proof) - 1 - 32 * 3):
        bad = bytearray(proof)
        bad[off] ^= 1
        tampered.append(bytes(bad))
    sw = bytearray(proof)                      # two proof points exchanged (both still on the curve)
    sw[0:64], sw[64:128] = proof[64:128], proof[0:64]
    tampered.append(bytes(sw))
    seen = set()
    for d2, pr in [(wrong, proof)] + [(deck2, t) for t in tampered]:
        want = co.verify(*a, deck, d2, pr)
        got = ctx.verify_shuffle(pk, deck, d2, pr)
        assert got == want != 0
        seen.add(got)
    assert 1 in seen and len(seen) >= 2


def test_shuffle_2p16_resident_entry_points(ctx, pkg, headline):
    """mp_shuffle_and_remask_resident / mp_shuffle_verify_resident / mp_shuffle_prove_resident -- the entry
    points bench.py's `value` times -- give the bytes of the host-buffer path (which the test above checks
    against the oracle) and the same verdicts."""
    import torch
    inst = headline
    lib = pkg.lib
    m, n, Nc = inst["m"], inst["n"], inst["N"]
    pk, deck, deck2, proof = inst["pk"], inst["deck"], inst["deck2"], inst["proof"]
    ctx.set_params(m, n, G64, inst["ck_g"], inst["ck_h"], inst["ghat"])
    dev = torch.device("cuda:0")
    d_deck = torch.frombuffer(bytearray(deck), dtype=torch.uint8).to(dev)
    d_deck2 = torch.frombuffer(bytearray(deck2), dtype=torch.uint8).to(dev)
    perm_arr = (ctypes.c_uint32 * Nc)(*inst["perm"])
    out_deck = ctypes.create_string_buffer(128 * Nc)
    out_proof = ctypes.create_string_buffer(lib.mp_proof_len(m, n))
    pkg.check(ctx.h, lib.mp_shuffle_and_remask_resident(ctx.h, pk, deck, perm_arr, inst["rho"], inst["rand"], out_deck, out_proof,
                                                        d_deck.data_ptr()))
    assert out_deck.raw == deck2 and out_proof.raw == proof
    out_proof2 = ctypes.create_string_buffer(lib.mp_proof_len(m, n))
    pkg.check(ctx.h, lib.mp_shuffle_prove_resident(ctx.h, pk, deck, deck2, perm_arr, inst["rho"], inst["rand"], out_proof2,
                                                   d_deck2.data_ptr()))
    assert out_proof2.raw == proof
    assert lib.mp_shuffle_verify_resident(ctx.h, pk, deck, deck2, proof, d_deck.data_ptr(), d_deck2.data_ptr()) == 0
    wrong = deck[128:] + deck[:128]
    d_wrong = torch.frombuffer(bytearray(wrong), dtype=torch.uint8).to(dev)
    assert lib.mp_shuffle_verify_resident(ctx.h, pk, deck, wrong, proof, d_deck.data_ptr(), d_wrong.data_ptr()) == 1
    bad = bytearray(proof)
    bad[-1 - 32 * 3] ^= 1
    assert lib.mp_shuffle_verify_resident(ctx.h, pk, deck, deck2, bytes(bad), d_deck.data_ptr(), d_deck2.data_ptr()) == 4


@pytest.mark.parametrize("m,n,seed", [(64, 128, 21), (128, 512, 5)])
def test_large_deck_proof_is_byte_exact_vs_c_oracle(ctx, m, n, seed, request):
    """Byte-for-byte proof parity on the large-deck code path (device scalar kernels, Karatsuba diagonal plan
    with 6 levels at m = 64 and 7 levels / 2 187 leaves at m = 128 -- the headline configuration itself),
    against the C restatement's prover on all host threads (about 10 s and 1.5-3 min)."""
    if (m, n) == (128, 512):
        inst = request.getfixturevalue("headline")
    else:
        inst = make_big_instance(ctx, m, n, seed)
        ctx.set_params(m, n, G64, inst["ck_g"], inst["ck_h"], inst["ghat"])
        inst["deck2"], inst["proof"] = ctx.shuffle_and_remask(inst["pk"], inst["deck"], inst["perm"], inst["rho"], inst["rand"])
    co = c_oracle.COracle(threads=os.cpu_count() or 1, msm_mode=1)
    want = co.prove(*oracle_args(inst), inst["deck"], inst["deck2"], inst["perm"], inst["rho"], inst["rand"])
    assert inst["proof"] == want


def test_usage_errors(ctx, pkg):
    lib = pkg.lib
    fresh = pkg.Context(0)
    buf = ctypes.create_string_buffer(64)
    assert lib.mp_shuffle_verify(fresh.h, buf, buf, buf, buf) == -4          # MP_ERR_NO_PARAMS
    assert lib.mp_ctx_set_params(fresh.h, 1, 5, buf, buf, buf, buf) == -1     # m < 2
    assert lib.mp_msm_g1(fresh.h, None, None, 3, 0, buf) == -1               # null inputs with n > 0
    assert lib.mp_msm_g1(fresh.h, buf, buf, 1, 99, buf) == -1                # window out of range
    assert b"window_bits" in lib.mp_last_error_string(fresh.h)
    fresh.close()
    # commitments longer than the key are refused
    rng = np.random.default_rng(9)
    pts = ctx.dbg_scalar_mul(G64 * 7, rand_scalars(rng, 7))
    ctx.set_params(2, 4, G64, pts[:256], pts[256:320], pts[320:384])
    with pytest.raises(pkg.MpError):
        ctx.commit_batch(rand_scalars(rng, 5), rand_scalars(rng, 1), 5)
    assert ctx.commit_batch(b"", b"", 3) == b""                              # empty batch
    assert ctx.remask(pts[384:448], b"", [], b"") == b""                     # empty deck


def test_shuffle_2p18_round_trip(ctx):
    """Beyond the BASELINE sizes: 2^18 cards, (m, n) = (256, 1024) -- 67 M diagonal terms, several
    GB of sort and bucket space -- still proves, verifies and rejects a tampered deck."""
    m, n = 256, 1024
    Nc = m * n
    rng = np.random.default_rng(6)
    npts = (n + 3) + 2 * Nc
    pts = ctx.dbg_scalar_mul(G64 * npts, rand_scalars(rng, npts))
    P = lambda i: pts[64 * i:64 * (i + 1)]
    ck_g, ck_h, ghat, pk, deck = pts[:64 * n], P(n), P(n + 1), P(n + 2), pts[64 * (n + 3):]
    perm = [int(v) for v in rng.permutation(Nc)]
    rho, rand = rand_scalars(rng, Nc), rand_scalars(rng, 11 * m + 5 * n)
    ctx.set_params(m, n, G64, ck_g, ck_h, ghat)
    deck2, proof = ctx.shuffle_and_remask(pk, deck, perm, rho, rand)
    assert ctx.verify_shuffle(pk, deck, deck2, proof) == 0
    sw = bytearray(deck2)
    sw[0:128], sw[128:256] = deck2[128:256], deck2[0:128]
    assert ctx.verify_shuffle(pk, deck, bytes(sw), proof) != 0

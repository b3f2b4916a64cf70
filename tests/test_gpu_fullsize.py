"""GPU parity at BASELINE.json's full sizes, through size-independent properties: known discrete
logs for the 2^20-term MSM, prove -> verify round trips and the reference's negative case on a
2^16-card deck, window-range split == single launch."""
import ctypes
import random

import numpy as np
import pytest

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


def test_shuffle_2p16_round_trip_and_negative(ctx, pkg, monkeypatch):
    m, n = 128, 512
    Nc = m * n
    rng = np.random.default_rng(5)
    npts = (n + 3) + 2 * Nc
    pts = ctx.dbg_scalar_mul(G64 * npts, rand_scalars(rng, npts))
    P = lambda i: pts[64 * i:64 * (i + 1)]
    ck_g, ck_h, ghat, pk, deck = pts[:64 * n], P(n), P(n + 1), P(n + 2), pts[64 * (n + 3):]
    perm = [int(v) for v in rng.permutation(Nc)]
    rho, rand = rand_scalars(rng, Nc), rand_scalars(rng, 11 * m + 5 * n)
    ctx.set_params(m, n, G64, ck_g, ck_h, ghat)
    deck2, proof = ctx.shuffle_and_remask(pk, deck, perm, rho, rand)
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

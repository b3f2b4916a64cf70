"""Multi-GPU entry points of the C ABI (NCCL communicator behind mp_comm_*): two contexts on
two GPUs of one box, one host thread each -- the single-process form of "one rank per GPU".  Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped on a single-GPU box.  Every collective result is compared with the single-GPU entry
point, which the other suites compare with the oracle."""
import ctypes
import threading

import numpy as np
import pytest

from oracle.py import stark
from _util import b32, pb

pytestmark = pytest.mark.gpu
G64 = pb(stark.G)


def rand_scalars(rng, k):
    a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8)
    a[:, 31] &= 0x07
    return a.tobytes()


def run_ranks(pkg, nranks, body):
    """body(rank, ctx) on one thread per rank, contexts joined by one communicator; returns the per-rank results"""
    uid = pkg.Context.comm_unique_id()
    out, err = [None] * nranks, [None] * nranks

    def work(r):
        try:
            ctx = pkg.Context(r)
            ctx.comm_init(nranks, r, uid)
            out[r] = body(r, ctx)
            ctx.comm_destroy()
            ctx.close()
        except Exception as e:  # noqa: BLE001
            err[r] = e
    ts = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.fixture(scope="module")
def nranks():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return min(n, 4)


def test_msm_window_split_over_nccl(pkg, ctx, nranks):
    import torch
    n = 1 << 16
    rng = np.random.default_rng(3)
    pts = ctx.dbg_scalar_mul(G64 * n, rand_scalars(rng, n))
    ks = rand_scalars(rng, n)
    want = ctx.msm_g1(pts, ks, 0)

    def body(r, c):
        dev = torch.device("cuda", r)
        d_pts = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
        d_ks = torch.frombuffer(bytearray(ks), dtype=torch.uint8).to(dev)
        d_out = torch.zeros(64, dtype=torch.uint8, device=dev)
        res = []
        for wbits in (0, 13, 16):
            c.msm_g1_multi_device(d_pts.data_ptr(), d_ks.data_ptr(), n, d_out.data_ptr(), wbits)
            c.sync()
            res.append(bytes(d_out.cpu().numpy().tobytes()))
        return res
    for res in run_ranks(pkg, nranks, body):
        assert res == [want] * 3


def test_one_large_proof_across_gpus(pkg, ctx, nranks):
    """mp_shuffle_and_remask_multi / mp_shuffle_verify_multi: the same bytes and verdicts as the single-GPU calls."""
    m, n = 32, 512
    Nc = m * n
    rng = np.random.default_rng(11)
    npts = (n + 3) + 2 * Nc
    pts = ctx.dbg_scalar_mul(G64 * npts, rand_scalars(rng, npts))
    P = lambda i: pts[64 * i:64 * (i + 1)]
    ck_g, ck_h, ghat, pk, deck = pts[:64 * n], P(n), P(n + 1), P(n + 2), pts[64 * (n + 3):]
    perm = [int(v) for v in rng.permutation(Nc)]
    rho, rand = rand_scalars(rng, Nc), rand_scalars(rng, 11 * m + 5 * n)
    ctx.set_params(m, n, G64, ck_g, ck_h, ghat)
    deck2, proof = ctx.shuffle_and_remask(pk, deck, perm, rho, rand)
    assert ctx.verify_shuffle(pk, deck, deck2, proof) == 0
    bad = bytearray(proof)
    bad[-32 * 4] ^= 1
    wrong = deck[128:] + deck[:128]

    def body(r, c):
        c.set_params(m, n, G64, ck_g, ck_h, ghat)
        d2, pf = c.shuffle_and_remask_multi(pk, deck, perm, rho, rand)
        return d2 == deck2, pf == proof, c.verify_shuffle_multi(pk, deck, deck2, proof), c.verify_shuffle_multi(pk, deck, deck2, bytes(bad)), \
            c.verify_shuffle_multi(pk, deck, wrong, proof)
    for res in run_ranks(pkg, nranks, body):
        assert res == (True, True, 0, 4, 1)


def test_batch_verdicts_are_all_gathered(pkg, ctx, nranks):
    from oracle import c_oracle
    from _util import instance
    m, n, per = 3, 4, 3
    co = c_oracle.COracle(msm_mode=1)
    pp0, pk0, *_ = instance(m, n, 90)
    enc_g, ck_g, ck_h, ghat, pk = pb(pp0.enc_g), b"".join(map(pb, pp0.ck_g)), pb(pp0.ck_h), pb(pp0.ghat), pb(pk0)
    shards = []
    for r in range(nranks):
        decks = decks2 = proofs = b""
        for s in range(per):
            _, _, deck, perm, rho, rnd = instance(m, n, 90 + r * per + s)
            deck_b = b"".join(pb(c[0]) + pb(c[1]) for c in deck)
            rho_b, rnd_b = b"".join(map(b32, rho)), b"".join(map(b32, rnd))
            d2 = co.remask(enc_g, pk, deck_b, perm, rho_b)
            pf = co.prove(m, n, enc_g, ck_g, ck_h, ghat, pk, deck_b, d2, perm, rho_b, rnd_b)
            if (r, s) == (1, 2):
                pf = pf[:-32 * 4] + bytes([pf[-32 * 4] ^ 1]) + pf[-32 * 4 + 1:]   # one bad proof on rank 1
            decks += deck_b; decks2 += d2; proofs += pf
        shards.append((decks, decks2, proofs))
    want = [0] * (per * nranks)
    want[1 * per + 2] = 4

    def body(r, c):
        c.set_params(m, n, enc_g, ck_g, ck_h, ghat)
        return c.verify_shuffle_batch_multi(pk, *shards[r], nranks, host_threads=2)
    for res in run_ranks(pkg, nranks, body):
        assert res == want


def test_multi_entry_points_degenerate_to_one_rank(pkg, ctx):
    """Without a communicator a context is rank 0 of 1: the collective entry points give the single-GPU results
    (this is what runs on a 1-GPU box; the real collectives are the tests above)."""
    import torch
    assert pkg.lib.mp_comm_size(ctx.h) == 1 and pkg.lib.mp_comm_rank(ctx.h) == 0
    n = 3000
    rng = np.random.default_rng(8)
    pts = ctx.dbg_scalar_mul(G64 * n, rand_scalars(rng, n))
    ks = rand_scalars(rng, n)
    want = ctx.msm_g1(pts, ks, 0)
    dev = torch.device("cuda:0")
    d_pts = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
    d_ks = torch.frombuffer(bytearray(ks), dtype=torch.uint8).to(dev)
    d_out = torch.zeros(64, dtype=torch.uint8, device=dev)
    for wbits in (0, 9, 16):
        ctx.msm_g1_multi_device(d_pts.data_ptr(), d_ks.data_ptr(), n, d_out.data_ptr(), wbits)
        ctx.sync()
        assert bytes(d_out.cpu().numpy().tobytes()) == want
    # protocol entry points: golden 52-card instance
    import json, os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))["shuffle"][2]
    h = bytes.fromhex
    ctx.set_params(fx["m"], fx["n"], h(fx["enc_g"]), h(fx["ck_g"]), h(fx["ck_h"]), h(fx["ghat"]))
    deck2, proof = ctx.shuffle_and_remask_multi(h(fx["pk"]), h(fx["deck"]), fx["perm"], h(fx["rho"]), h(fx["rand"]))
    assert deck2.hex() == fx["deck2"] and proof.hex() == fx["proof"]
    assert ctx.verify_shuffle_multi(h(fx["pk"]), h(fx["deck"]), deck2, proof) == 0
    assert ctx.verify_shuffle_batch_multi(h(fx["pk"]), h(fx["deck"]) * 2, deck2 * 2, proof * 2, 1) == [0, 0]


def test_comm_init_needs_a_valid_shape(pkg):
    c = pkg.Context(0)
    with pytest.raises(pkg.MpError):
        c.comm_init(2, 5, bytes(128))       # rank outside [0, nranks)
    c.comm_destroy()                        # no communicator: a no-op
    c.close()

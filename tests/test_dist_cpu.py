"""world_size-2 gloo test (CPU) of the multi-GPU host logic: window-range split of one MSM with an
all-gather of the partials and a local fold; proof-index sharding.  The per-rank "engine" here is
the Python oracle (signed-window partial sums restated in a few lines), so no GPU is needed."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _signed_digits(k, c, W):
    out, carry = [], 0
    for w in range(W):
        raw = ((k >> (w * c)) & ((1 << c) - 1)) + carry
        if raw > (1 << (c - 1)):
            out.append(raw - (1 << c))
            carry = 1
        else:
            out.append(raw)
            carry = 0
    assert carry == 0
    return out


def _worker(rank, world, port, c, q, curve="stark"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    if curve == "stark":
        from oracle.py import stark
        from _util import chain_points, pb, b32
        PB, scalar_bits = 64, 253
    else:  # second curve: 96-byte points, 253-bit group order (+1 bit for the signed recoding)
        from oracle.py import bls12_377 as stark
        from _util_bls12_377 import chain_points, pb, b32
        stark.point_from_bytes64 = stark.point_from_bytes
        PB, scalar_bits = 96, 254
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s0, s1, pts, st = chain_points(24, 5)
    ks = [st.scalar() for _ in range(24)]
    ks[0] = stark.N - 1
    W = (scalar_bits + c - 1) // c
    digs = [_signed_digits(k, c, W) for k in ks]

    def partial(w_begin, w_count):  # sum_w 2^(c (w - w_begin)) * sum_i d_{i,w} P_i
        acc = None
        for w in range(w_begin + w_count - 1, w_begin - 1, -1):
            acc = stark.mul(acc, 1 << c) if acc is not None else None
            ws = stark.msm(pts, [d[w] % stark.N for d in digs])
            acc = stark.add(acc, ws)
        return pb(acc)

    def fold(points, scalars):
        n = len(scalars) // 32
        P = [stark.point_from_bytes64(points[PB * i:PB * i + PB]) for i in range(n)]
        K = [int.from_bytes(scalars[32 * i:32 * i + 32], "little") for i in range(n)]
        return pb(stark.msm(P, K))

    got = pkg.dist.window_split_msm(partial, fold, c, W, point_bytes=PB)
    want = pb(stark.msm(pts, ks))
    q.put((rank, got == want, pkg.dist.shard_range(4096, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,c,curve", [(2, 16, "stark"), (3, 13, "stark"), (2, 16, "bls12_377")])
def test_window_split_msm_gloo(world, c, curve, pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, c, q, curve)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    # proof-index shards tile [0, 4096) without gaps
    edges = [r[2] for r in res]
    assert edges[0][0] == 0 and edges[-1][1] == 4096
    assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_shard_helpers(pkg):
    d = pkg.dist
    assert [d.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [d.window_range(16, r, 8) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]
    assert d.window_range(3, 5, 8) == (3, 3)  # more ranks than windows: empty share
    fs = d.fold_scalars(16, 16, 8)
    assert len(fs) == 8 and int.from_bytes(fs[3][1], "little") == 1 << (16 * 6)

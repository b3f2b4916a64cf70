"""Multi-GPU sharding of the hot path: one process per GPU, plumbing by
`torch.distributed` (NCCL on the GPU box, gloo in the CPU tests).

* batches of independent proofs: proof-index split, no data-path collective (`shard_range`);
* one large MSM: window-range split.  Every rank holds the inputs, computes the windows
  `window_range(W, rank, world)` end to end (`mp_msm_g1_windows_device`), the 64-byte partial
  results are all-gathered, and  MSM = sum_r 2^(c * w_begin_r) * P_r  is folded with one tiny
  MSM.  EC addition is not an ncclRedOp, so the "reduce of partial sums" is bytes + a local fold.

The reference has no distributed path at all (single-threaded library); this module has no
reference counterpart to mirror.
"""


def shard_range(total, rank, world):
    """Contiguous, balanced [begin, end) of `total` items for `rank` of `world`."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def window_range(num_windows, rank, world):
    """Windows [w_begin, w_end) owned by `rank`; ranks beyond the window count get nothing."""
    return shard_range(num_windows, rank, world)


def fold_scalars(window_bits, num_windows, world):
    """Canonical 32-byte scalars 2^(c * w_begin_r) for every rank r that owns windows."""
    out = []
    for r in range(world):
        b, e = window_range(num_windows, r, world)
        if e > b:
            out.append((r, (1 << (window_bits * b)).to_bytes(32, "little")))
    return out


def window_split_msm(partial_fn, fold_fn, window_bits, num_windows, group=None, point_bytes=64):
    """Generic driver.  `partial_fn(w_begin, w_count) -> point_bytes bytes` computes this rank's partial
    (canonical affine, all-zero = identity); `fold_fn(points_bytes, scalars_bytes)` evaluates a small MSM.
    Returns the full MSM result on every rank.  point_bytes: 64 (Stark curve) or 96 (BLS12-377 G1)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = window_range(num_windows, rank, world)
    mine = partial_fn(b, e - b) if e > b else bytes(point_bytes)
    assert len(mine) == point_bytes
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    pts, scs = b"", b""
    for r, s in fold_scalars(window_bits, num_windows, world):
        pts += bytes(gathered[r].cpu().numpy().tobytes())
        scs += s
    return fold_fn(pts, scs)

#!/usr/bin/env python
"""Benchmark of the shuffle-proof hot path (BASELINE.json metric: shuffle proofs/s, prove+verify,
and MSM EC-adds/s, next to the CPU reference path on the same box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One step = `--decks` (default 8) independent synthetic N-card decks, each put through one
`shuffle_and_remask` (permute + remask + ShuffleArgument::prove) and one `verify_shuffle` of its output, by
the two batch entry points of the C ABI: a few worker contexts overlap one deck's serial Blake2s statement
hash (host) with the other decks' kernels (device).  Default workload: 2^16 cards, (m, n) = (128, 512) -- the
configuration BASELINE.json's target is quoted on.  With --gpus N > 1 (under torchrun) every rank proves and
verifies its own decks (proof-index split, weak scaling, no data-path collective).

`value`  : proofs/s with the decks already resident in HBM (mp_shuffle_*_batch_resident).
`e2e`    : proofs/s through the host-buffer C ABI (mp_shuffle_and_remask_batch + mp_shuffle_verify_batch):
           decks, permutations and scalars cross PCIe inside the timed region, decks and proofs come back.
`latency`: the strictly sequential single-deck step (round 1's headline), for reference.
In all of them the Fiat-Shamir transcript (Blake2s over the serialized decks) runs on the host inside the timed
region, as the reference design prescribes.  `--impl reference` RUNS the same (m, n) once, in full, in the C
restatement of the reference's CPU path on all host threads.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GX = 0x01EF15C18599971B7BECED415A40F0C7DEACFD9B0D1819E03D723D8BC943CFCA
GY = 0x005668060AA49730B7BE4801DF46EC62DE53ECD11ABE43A32873000C36E8DC1F
G64 = GX.to_bytes(32, "little") + GY.to_bytes(32, "little")
# BLS12-377 G1 generator, x || y as 48-byte little-endian coordinates (include/mpshuffle_bls12_377.h)
GEN_BLS12_377 = (0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF.to_bytes(48, "little")
                 + 0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6.to_bytes(48, "little")).hex()
METRIC = "shuffle proofs/sec (prove+verify)"
UNIT = "proofs/s"


# --------------------------------------------------------------------------------------------
# work model: point-scalar terms per proof, used ONLY to extrapolate the
# bounded CPU sample to the full workload
# --------------------------------------------------------------------------------------------
def work_terms(m, n):
    N = m * n
    prove_naive = 2 * N + 2 * m * (m + 1) * n + 4 * m * 2       # remask, diagonal CT-MSMs, Enc(b_k ghat; tau_k)
    prove_pedersen = (3 * m + 6) * (n + 1) + (4 * m + 1) * 2     # commitments (Pippenger on the CPU)
    verify_naive = 4 * N + 4 * m + 4 + (m + 1) + (m + 1) + (2 * m + 1) + (m + 1) + 2 * m + 3 * m + 6
    verify_pedersen = 5 * (n + 1) + 2 * 2 + 2 * (n + 1)
    return dict(prove_naive=prove_naive, prove_pedersen=prove_pedersen, verify_naive=verify_naive,
                verify_pedersen=verify_pedersen)


def rand_scalars(rng, k):
    import numpy as np
    a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8)
    a[:, 31] &= 0x07  # < 2^251 < group order
    return a.tobytes()


def make_instance(ctx, m, n, seed):
    """Synthetic instance generated on the GPU: every point is s*G for a seeded scalar s."""
    import numpy as np
    rng = np.random.default_rng(seed)
    N = m * n
    npts = (n + 3) + 2 * N
    pts = ctx.dbg_scalar_mul(G64 * npts, rand_scalars(rng, npts))
    P = lambda i: pts[64 * i:64 * (i + 1)]
    inst = dict(m=m, n=n, N=N, enc_g=G64, ck_g=pts[:64 * n], ck_h=P(n), ghat=P(n + 1), pk=P(n + 2),
                deck=pts[64 * (n + 3):], perm=[int(v) for v in rng.permutation(N)],
                rho=rand_scalars(rng, N), rand=rand_scalars(rng, 11 * m + 5 * n))
    return inst


class ClockSampler:
    """SM clocks / throttle reasons of this rank's GPU sampled DURING the timed region.  In-process NVML
    (nvidia-ml-py) every 20 ms: a sample costs microseconds, where spawning `nvidia-smi` from every rank
    ten times a second competes with the statement hash for the host cores -- that alone cost the
    8-GPU run 10 % (eight serial Blake2s chains share 16 vCPUs).  `nvidia-smi` stays as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip()]
        return int(ids[index]) if ids and all(v.strip().isdigit() for v in ids) and index < len(ids) else index

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                    n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
            return [str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in bits]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        return [c.strip() for c in out.strip().split(",")]

    def run(self):
        while not self.stop:
            try:
                self.rows.append(self.sample())
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.25)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        reasons = [nm for i, nm in enumerate(self.NAMES) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm), source="nvml" if self.nvml is not None else "nvidia-smi")


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------------------------
# CPU legs (oracle; the only place the bench executes oracle/)
# --------------------------------------------------------------------------------------------
def cpu_instance(m, n, seed, threads=None):
    """Instance for the CPU legs, built with the C oracle itself (no GPU needed): every point is s*G for a
    seeded scalar s, computed by oc_scalar_mul_batch on all host threads."""
    import numpy as np
    from oracle import c_oracle
    co = c_oracle.COracle(threads=threads or os.cpu_count() or 1, msm_mode=1)
    rng = np.random.default_rng(seed)
    N = m * n
    npts = (n + 3) + 2 * N
    pts = co.scalar_mul_batch(G64, rand_scalars(rng, npts))
    P = lambda i: pts[64 * i:64 * (i + 1)]
    return dict(m=m, n=n, N=N, enc_g=G64, ck_g=pts[:64 * n], ck_h=P(n), ghat=P(n + 1), pk=P(n + 2),
                deck=pts[64 * (n + 3):], perm=[int(v) for v in rng.permutation(N)],
                rho=rand_scalars(rng, N), rand=rand_scalars(rng, 11 * m + 5 * n))


def cpu_run(inst, threads, msm_mode, prove=True):
    """One remask + prove (optional) + verify of `inst` in the C restatement.  msm_mode 0 = the faithful cost
    model of the reference's CPU path (per-term double-and-add for ciphertext sums as proof-essentials does,
    ark-ec 0.3 Pippenger for commitments); 1 = best effort (Pippenger for the big sums too).
    -> (prove seconds or None, verify seconds, proof bytes)"""
    from oracle import c_oracle
    co = c_oracle.COracle(threads=threads, msm_mode=msm_mode)
    a = (inst["m"], inst["n"], inst["enc_g"], inst["ck_g"], inst["ck_h"], inst["ghat"], inst["pk"])
    tp = None
    if prove or "proof" not in inst:
        t0 = time.perf_counter()
        inst["deck2"] = co.remask(inst["enc_g"], inst["pk"], inst["deck"], inst["perm"], inst["rho"])
        inst["proof"] = co.prove(*a, inst["deck"], inst["deck2"], inst["perm"], inst["rho"], inst["rand"])
        tp = time.perf_counter() - t0
    t1 = time.perf_counter()
    ok = co.verify(*a, inst["deck"], inst["deck2"], inst["proof"])
    tv = time.perf_counter() - t1
    assert ok == 0, "the C restatement rejected its own proof"
    return tp, tv


def sample_shape(m, n, cap):
    """Smaller deck of the same m:n aspect with at most `cap` cards (bounded CPU samples)."""
    sm, sn = m, n
    while sm * sn > cap and sm > 2 and sn > 2:
        if sm >= 4:
            sm //= 2
        if sm * sn > cap and sn >= 4:
            sn //= 2
    return max(sm, 2), max(sn, 2)


# --------------------------------------------------------------------------------------------
def reference_arm(args, m, n, config):
    """`--impl reference`: the reference's own CPU path (C restatement of it, oracle/c -- the Rust crate cannot be
    built in this image) on all host threads, on the SAME (m, n): one full shuffle_and_remask + verify_shuffle
    is actually run (no extrapolation) in the faithful mode, which is the line's value, and once more in the
    best-effort mode (Pippenger for the big sums), reported beside it."""
    threads = os.cpu_count() or 1
    t00 = time.perf_counter()
    inst = cpu_instance(m, n, 1, threads)
    t_inst = time.perf_counter() - t00
    budget = args.ref_budget_s
    # predict the faithful prove from a small deck so that a slow host falls back instead of running for an hour
    sm, sn = sample_shape(m, n, 1024)
    small = cpu_instance(sm, sn, 2, threads)
    tps, tvs = cpu_run(small, threads, 0)
    ws, wf = work_terms(sm, sn), work_terms(m, n)
    predicted = tps * wf["prove_naive"] / ws["prove_naive"] + tvs * wf["verify_naive"] / ws["verify_naive"]
    extrapolated = predicted > budget
    if not extrapolated:
        tp, tv = cpu_run(inst, threads, 0)
        sample = (f"C restatement of the reference CPU path (oracle/c, not the Rust binary), faithful mode, {threads} threads: ONE full "
                  f"shuffle_and_remask + verify_shuffle at (m,n)=({m},{n}) actually run: prove {tp:.1f} s + verify {tv:.1f} s")
    else:
        # verify at full size is cheap enough to measure; the prover is scaled from the small deck
        tpb, _ = cpu_run(inst, threads, 1)             # best-effort prover makes the proof the verifier needs
        _, tv = cpu_run(inst, threads, 0, prove=False)
        tp = tps * wf["prove_naive"] / ws["prove_naive"]
        sample = (f"C restatement (oracle/c), faithful mode, {threads} threads: verify_shuffle at (m,n)=({m},{n}) measured ({tv:.1f} s); "
                  f"the faithful prover was predicted at {predicted:.0f} s > --ref-budget-s {budget:.0f} and is SCALED from a "
                  f"({sm},{sn}) deck by term count ({tp:.0f} s)")
    best = None
    if not extrapolated and (time.perf_counter() - t00) + 0.6 * (tp + tv) < 2.5 * budget:
        tp1, tv1 = cpu_run(inst, threads, 1)
        best = dict(value=1.0 / (tp1 + tv1), unit=UNIT, prove_s=tp1, verify_s=tv1, cores=threads, kind="port",
                    sample="same deck, best-effort mode: signed-window Pippenger for the ciphertext sums as well "
                           "(BASELINE.md section 3: so that the GPU speed-up is not flattered by a naive baseline)")
    elif extrapolated:
        best = dict(value=1.0 / (tpb + tv), unit=UNIT, prove_s=tpb, verify_s=None, cores=threads, kind="port",
                    sample="best-effort prover measured at full size; verify time of the faithful mode used")
    val = 1.0 / (tp + tv)
    line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=1, warmup=0, ms_per_step=1000.0 * (tp + tv),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64", data="synthetic", config=config,
                impl="reference", steps_requested=args.steps, warmup_requested=args.warmup, extrapolated=extrapolated,
                cpu_baseline=dict(value=val, unit=UNIT, cores=threads, kind="port", sample=sample, prove_s=tp, verify_s=tv,
                                  best_effort=best),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, instance_s=t_inst, wall_s=time.perf_counter() - t00)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=128)
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--decks", type=int, default=8, help="independent decks per step (pipelined over worker contexts)")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="verify each step's own proofs after proving them (two phases) instead of verifying the previous step's "
                         "proofs while this step's decks are proved")
    ap.add_argument("--workers", type=int, default=8, help="worker contexts per batch call of the headline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=600.0,
                    help="--impl reference: run the full-size faithful prover only if it is predicted to fit this many seconds")
    ap.add_argument("--batch52", type=int, default=512, help="proofs in the 52-card batch measurement (0 = skip)")
    ap.add_argument("--latency-steps", type=int, default=3, help="strictly sequential single-deck steps (latency + roofline kernel timing)")
    ap.add_argument("--sigma-cards", type=int, default=65536,
                    help="cards in the batched mask / remask / reveal and the wire-format measurements (SURVEY 8(f) ranks 1-2; 0 = skip)")
    ap.add_argument("--msm-logn", type=int, default=20, help="size of the MSM microbench reported beside the metric")
    ap.add_argument("--bls12-377-logn", type=int, default=20,
                    help="size of the BLS12-377 G1 MSM measurement (second curve, SURVEY 8(f) rank 3; 0 = skip)")
    ap.add_argument("--detail-file", default=os.path.join(ROOT, "gpurun_out", "bench_detail.json"),
                    help="sidecar for the bulky side measurements (sigma, wire, bls12_377)")
    args = ap.parse_args()
    m, n = args.m, args.n
    N = m * n
    Q = max(1, args.decks)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"{N}-card deck shuffle prove+verify, (m,n)=({m},{n}), Stark curve"
    config = dict(workload=workload, m=m, n=n, cards=N, decks_per_step=Q, l2="flushed between steps (256 MiB write)",
                  step=(f"{Q} independent decks per GPU per step: mp_shuffle_and_remask_batch of this step's decks on one context WHILE "
                        f"mp_shuffle_verify_batch checks the previous step's {Q} proofs on a second context (every deck proved once and verified "
                        "once; worker contexts overlap the serial Blake2s statement hashes with other decks' kernels)") if args.overlap else
                       (f"{Q} independent decks per GPU per step: mp_shuffle_and_remask_batch, then mp_shuffle_verify_batch of the same decks"),
                  sharding="proof-index split: independent decks per GPU, no data-path collective" if world > 1 else "single GPU")

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, m, n, config)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # keep stdout to the single JSON line: fd 1 is pointed at stderr for the whole run (NCCL prints its
    # version banner to stdout whatever NCCL_DEBUG_FILE says) and the line itself goes to a duplicate of the
    # original descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    host_threads = max(1, (os.cpu_count() or 1) // world)   # every rank hashes and schedules on its share of the host
    ctx = pkg.Context(local_rank)
    lib = pkg.lib
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    if world > 1:
        # the library's own NCCL communicator (mp_comm_*): rank 0 makes the id, torch.distributed only carries its 128 bytes
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(pkg.Context.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, src=0)
        ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))

    # Q decks: same parameters and key, independent decks / permutations / randomness
    inst = make_instance(ctx, m, n, seed=1 + rank)
    ctx.set_params(m, n, inst["enc_g"], inst["ck_g"], inst["ck_h"], inst["ghat"])
    rng = np.random.default_rng(1000 + rank)
    decks = inst["deck"] + ctx.dbg_scalar_mul(G64 * (2 * N * (Q - 1)), rand_scalars(rng, 2 * N * (Q - 1))) if Q > 1 else inst["deck"]
    perms = np.concatenate([np.asarray(inst["perm"], dtype=np.uint32)] + [rng.permutation(N).astype(np.uint32) for _ in range(Q - 1)])
    rhos = inst["rho"] + rand_scalars(rng, N * (Q - 1))
    rands = inst["rand"] + rand_scalars(rng, (11 * m + 5 * n) * (Q - 1))
    perm_p = perms.ctypes.data_as(ctypes.c_void_p)
    plen = lib.mp_proof_len(m, n)
    out_decks = ctypes.create_string_buffer(128 * N * Q)
    proofs = ctypes.create_string_buffer(plen * Q)
    statuses = (ctypes.c_int32 * Q)()
    d_decks = torch.frombuffer(bytearray(decks), dtype=torch.uint8).to(dev)
    # the shuffled decks of these inputs, resident for the verifier of `value` (the prover leaves its output on the host)
    pkg.check(ctx.h, lib.mp_shuffle_and_remask_batch(ctx.h, inst["pk"], decks, perm_p, rhos, rands, Q, out_decks, proofs, host_threads))
    d_decks2 = torch.frombuffer(bytearray(out_decks.raw), dtype=torch.uint8).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    phase = dict(prove=0.0, verify=0.0, n=0)
    # Two contexts on this GPU, as a service that proves and verifies streams of decks would hold them: `ctx` proves,
    # `ctx_v` verifies.  A step proves this step's Q decks and verifies the Q proofs of the PREVIOUS step at the same
    # time (every deck is proved once and verified once; the first timed step verifies the last warm-up step's proofs).
    # The verifier's serial statement hashes then overlap the prover's kernels instead of leaving the device idle.
    ctx_v = pkg.Context(local_rank)
    ctx_v.set_params(m, n, inst["enc_g"], inst["ck_g"], inst["ck_h"], inst["ghat"])
    bufs = [(out_decks, proofs), (ctypes.create_string_buffer(128 * N * Q), ctypes.create_string_buffer(plen * Q))]
    bufs[1][0].raw, bufs[1][1].raw = out_decks.raw, proofs.raw   # "previous step" of the very first step: the set-up pass
    state = dict(k=0)
    # host_threads of the two large-deck batch calls = the CPU threads each may keep busy: this rank's share of the host,
    # halved when the two calls run concurrently.  The library runs 8 worker contexts per call whatever the budget (they
    # sleep while their kernels run) and, when the budget is below half of that, hashes the statements of the 8 decks together
    # (multi-stream Blake2s) instead of one serial 23.5 ms pass per worker -- which is what a shared host cannot afford
    prove_threads = verify_threads = max(2, host_threads // 2) if args.overlap else max(2, host_threads)
    os.environ.setdefault("MP_PROVE_WORKERS", str(args.workers))   # read by the library at its first large-deck batch call

    def prove_call(resident, out_d, out_p):
        if resident:
            return lib.mp_shuffle_and_remask_batch_resident(ctx.h, inst["pk"], decks, perm_p, rhos, rands, Q, out_d, out_p,
                                                            prove_threads, d_decks.data_ptr())
        return lib.mp_shuffle_and_remask_batch(ctx.h, inst["pk"], decks, perm_p, rhos, rands, Q, out_d, out_p, prove_threads)

    def verify_call(resident, in_d, in_p):
        if resident:
            return lib.mp_shuffle_verify_batch_resident(ctx_v.h, inst["pk"], decks, in_d, in_p, Q, statuses, verify_threads,
                                                        d_decks.data_ptr(), d_decks2.data_ptr())
        return lib.mp_shuffle_verify_batch(ctx_v.h, inst["pk"], decks, in_d, in_p, Q, statuses, verify_threads)

    def step(resident):
        k = state["k"]
        state["k"] = k + 1
        cur, prev = bufs[k % 2], bufs[(k + 1) % 2]
        t_a = time.perf_counter()
        if args.overlap:
            res = {}

            def vrun():
                res["rc"] = verify_call(resident, *prev)       # BarnettSmartProtocol::verify_shuffle, previous step's proofs
                res["t"] = time.perf_counter()
            th = threading.Thread(target=vrun)
            th.start()
            rc = prove_call(resident, *cur)                    # BarnettSmartProtocol::shuffle_and_remask, this step's decks
            t_b = time.perf_counter()
            th.join()
            pkg.check(ctx.h, rc)
            pkg.check(ctx_v.h, res["rc"])
            t_p, t_v = t_b - t_a, res["t"] - t_a
        else:
            pkg.check(ctx.h, prove_call(resident, *cur))
            t_b = time.perf_counter()
            pkg.check(ctx_v.h, verify_call(resident, *cur))
            t_p, t_v = t_b - t_a, time.perf_counter() - t_b
        if any(statuses):
            raise SystemExit(f"bench: verify_shuffle rejected a valid proof (statuses {list(statuses)})")
        if resident:
            phase["prove"] += t_p
            phase["verify"] += t_v
            phase["n"] += 1
        return ctx.launches + ctx_v.launches

    def timed_steps(resident, steps):
        total_ms, launches = 0.0, 0
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launches += step(resident)     # synchronous: returns when every worker stream has drained
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms, launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(True)
        step(False)
    phase.update(prove=0.0, verify=0.0, n=0)
    # ---- `value`: decks resident in HBM
    barrier()
    with ClockSampler(local_rank) as clocks:
        ms_res, launches = timed_steps(True, args.steps)
    barrier()
    # ---- `e2e`: host buffers through the public C ABI
    ms_e2e, _ = timed_steps(False, args.steps)
    barrier()
    if world > 1:
        t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_e2e = t.tolist()
    value = world * Q * args.steps / (ms_res / 1e3)
    e2e = world * Q * args.steps / (ms_e2e / 1e3)

    # ---- strictly sequential single-deck steps: latency, and the roofline kernel timed without other streams
    perm_arr = (ctypes.c_uint32 * N)(*inst["perm"])
    deck2_buf, proof_buf = ctypes.create_string_buffer(128 * N), ctypes.create_string_buffer(plen)
    d_deck, d_deck2 = d_decks[:128 * N], d_decks2[:128 * N]

    def seq_step():
        pkg.check(ctx.h, lib.mp_shuffle_and_remask_resident(ctx.h, inst["pk"], inst["deck"], perm_arr, inst["rho"], inst["rand"],
                                                            deck2_buf, proof_buf, d_deck.data_ptr()))
        t_mid = time.perf_counter()
        rc = lib.mp_shuffle_verify_resident(ctx.h, inst["pk"], inst["deck"], deck2_buf, proof_buf, d_deck.data_ptr(), d_deck2.data_ptr())
        if pkg.check(ctx.h, rc) != 0:
            raise SystemExit(f"bench: verify_shuffle rejected a valid proof (status {rc})")
        return t_mid
    seq_step()
    ctx.profile_enable(True)
    ctx.profile_collect()
    lat_p = lat_v = 0.0
    for _ in range(max(1, args.latency_steps)):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t1 = seq_step()
        t2 = time.perf_counter()
        lat_p += t1 - t0
        lat_v += t2 - t1
    acc_ms, acc_adds, acc_launches = ctx.profile_collect_dominant()  # the prover's diagonal-product launches (bulk stream)
    ctx.profile_enable(False)
    lat_p, lat_v = lat_p / max(1, args.latency_steps), lat_v / max(1, args.latency_steps)

    # one large proof across the GPUs of the job: every rank proves and verifies THE SAME deck
    # (rank 0's), the prover's leaf products split by rank over the library's NCCL communicator
    one_proof = None
    if world > 1:
        try:
            ctx0 = pkg.Context(local_rank) if rank != 0 else None  # rank 0's instance on every rank
            inst0 = make_instance(ctx0 if ctx0 is not None else ctx, m, n, seed=1)
            if ctx0 is not None:
                ctx0.close()
            ctx.set_params(m, n, inst0["enc_g"], inst0["ck_g"], inst0["ck_h"], inst0["ghat"])  # one statement, every rank
            perm0 = (ctypes.c_uint32 * N)(*inst0["perm"])
            mp_t, mv_t = [], []
            for it in range(4):
                dist.barrier()
                t0 = time.perf_counter()
                pkg.check(ctx.h, lib.mp_shuffle_and_remask_multi(ctx.h, inst0["pk"], inst0["deck"], perm0, inst0["rho"], inst0["rand"],
                                                                 deck2_buf, proof_buf))
                t1 = time.perf_counter()
                rc = pkg.check(ctx.h, lib.mp_shuffle_verify_multi(ctx.h, inst0["pk"], inst0["deck"], deck2_buf, proof_buf))
                t2 = time.perf_counter()
                if rc != 0:
                    raise SystemExit("bench: multi-GPU verify rejected a valid proof")
                if it >= 1:
                    mp_t.append(t1 - t0)
                    mv_t.append(t2 - t1)
            t = torch.tensor([min(mp_t), min(mv_t)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            one_proof = dict(n_gpus=world, prove_ms=t[0].item() * 1e3, verify_ms=t[1].item() * 1e3,
                             single_gpu_prove_ms=lat_p * 1e3, single_gpu_verify_ms=lat_v * 1e3,
                             proof_sha=__import__("hashlib").sha256(proof_buf.raw).hexdigest()[:16],
                             note="mp_shuffle_and_remask_multi + mp_shuffle_verify_multi, host buffers, wall clock, max over ranks; "
                                  "the serial Blake2s statement hash (~23 ms each) is not split")
        except Exception as e:
            one_proof = dict(error=repr(e))

    # integer-pipe denominators measured in this run: bare IMAD.WIDE and bare XYZZ mixed additions
    def microbench(which, iters):
        best = 0.0
        for _ in range(2):
            t_ms, ops = ctx.dbg_bench(which, iters)
            best = max(best, ops / (t_ms / 1e3))
        return best
    try:
        imad_wide_peak, madd_peak, imad_cc_peak = microbench(0, 2000), microbench(3, 300), microbench(5, 1000)
    except Exception:
        imad_wide_peak = madd_peak = imad_cc_peak = None

    # MSM microbench (BASELINE config 5), collective when world > 1: run it on every rank
    try:
        msm_res = msm_microbench(ctx, torch, dev, stream, args.msm_logn, pkg, world, rank)
    except Exception as e:  # never lose the headline line to the side measurement
        msm_res = dict(error=repr(e))
    try:
        b52 = batch52_bench(pkg, ctx, torch, stream, args.batch52, rank, host_threads) if args.batch52 > 0 else None
        if b52 is not None and world > 1:
            t = torch.tensor([b52["prove_s"] + b52["verify_s"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            b52["proofs_per_s_all_gpus"] = world * b52["batch"] / t.item()
            b52["n_gpus"] = world
    except Exception as e:
        b52 = dict(error=repr(e))
    detail = {}
    if rank == 0:
        try:
            detail["sigma"] = sigma_bench(pkg, ctx, args.sigma_cards, not args.no_cpu_baseline) if args.sigma_cards > 0 else None
        except Exception as e:
            detail["sigma"] = dict(error=repr(e))
        try:
            detail["wire"] = wire_bench(pkg, ctx, args.sigma_cards, not args.no_cpu_baseline) if args.sigma_cards > 0 else None
        except Exception as e:
            detail["wire"] = dict(error=repr(e))
        try:
            detail["bls12_377"] = (bls12_377_bench(pkg, torch, dev, args.bls12_377_logn, not args.no_cpu_baseline)
                                   if args.bls12_377_logn > 0 else None)
        except Exception as e:
            detail["bls12_377"] = dict(error=repr(e))
    if rank == 0:
        peak, peak_src = load_peaks()
        bytes_per_add = 68.0  # 64 B affine point gather + 4 B sorted index
        imad_per_add = 740.0  # 8M + 2S: 8 x 74 + 2 x 46 IMAD.WIDE plus the reductions
        adds_per_s = acc_adds / (acc_ms / 1e3) if acc_ms > 0 else None
        achieved = bytes_per_add * adds_per_s / 1e9 if adds_per_s else None
        traffic, traffic_src = None, None
        try:  # dram bytes per launch of the same kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if (m, n) == (tj["m"], tj["n"]):
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        ms_per_proof = ms_res / (args.steps * Q)
        roofline = dict(bound="hbm", kernel="k_accumulate<2> (bucket accumulation of the prover's diagonal ciphertext products -- Karatsuba leaf jobs, XYZZ mixed adds)",
                        achieved=achieved, peak=peak, unit="GB/s", frac=(achieved / peak if achieved else None), traffic=traffic,
                        traffic_source=traffic_src, algorithmic_bytes_per_launch=(bytes_per_add * acc_adds / acc_launches if acc_launches else None),
                        peak_source=peak_src, launches=acc_launches,
                        avg_launch_ms=(acc_ms / acc_launches if acc_launches else None),
                        timed="CUDA events around the launch on its own stream, in the sequential single-deck steps of this run "
                              "(one launch per proof; in the pipelined steps it overlaps other decks' kernels)",
                        share_of_step=((acc_ms / acc_launches) / ms_per_proof if acc_launches and ms_per_proof else None),
                        ec_adds_per_s=adds_per_s,
                        int_pipe=dict(bound="integer pipe (IMAD.WIDE issue) -- the binding roofline of this kernel, SURVEY.md 8(d)",
                                      imad_wide_per_add=imad_per_add,
                                      achieved_imad_wide_per_s=(adds_per_s * imad_per_add if adds_per_s else None),
                                      peak_measured=imad_wide_peak, peak_source="mp_dbg_bench(0): bare IMAD.WIDE loop, this run",
                                      frac=(adds_per_s * imad_per_add / imad_wide_peak if adds_per_s and imad_wide_peak else None),
                                      peak_carry_chained=imad_cc_peak,
                                      frac_of_carry_chained=(adds_per_s * imad_per_add / imad_cc_peak if adds_per_s and imad_cc_peak else None),
                                      peak_carry_chained_source="mp_dbg_bench(5): IMAD.WIDE in mad.lo.cc / madc.hi.cc chains (the only form a "
                                                                "multi-limb product can use), this run",
                                      madd_microbench_per_s=madd_peak,
                                      madd_microbench_frac=(adds_per_s / madd_peak if adds_per_s and madd_peak else None)),
                        note="integer-pipe bound kernel (10 field multiplications per 68 B): the HBM fraction is structurally low, "
                             "int_pipe is the roofline that binds")
        # per deck -- prove: deck, perm, rho, randomness; verify: deck, shuffled deck, proof
        h2d = Q * (128 * N + 4 * N + 32 * N + 32 * (11 * m + 5 * n) + 128 * N * 2 + plen)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_res / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="u32", data="synthetic", config=config,
                    e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=Q * (128 * N + plen),
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=launches, clocks=clocks.summary(), roofline=roofline,
                    phases=dict(prove_ms_per_step=1e3 * phase["prove"] / max(phase["n"], 1), verify_ms_per_step=1e3 * phase["verify"] / max(phase["n"], 1),
                                overlapped=bool(args.overlap), worker_contexts_per_call=int(os.environ.get("MP_PROVE_WORKERS", args.workers)), prove_host_threads=prove_threads, verify_host_threads=verify_threads,
                                note="host wall clock of the two batch calls inside the steps of `value`" +
                                     (" (they run concurrently: a step lasts as long as the longer one)" if args.overlap else "")),
                    latency=dict(prove_ms=lat_p * 1e3, verify_ms=lat_v * 1e3, proofs_per_s_sequential=1.0 / (lat_p + lat_v),
                                 note="one deck at a time through mp_shuffle_and_remask_resident + mp_shuffle_verify_resident (round 1's headline step)"))
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_leg(m, n)
            except Exception as e:
                line["cpu_baseline"] = dict(error=repr(e))
        try:
            os.makedirs(os.path.dirname(args.detail_file), exist_ok=True)
            json.dump(detail, open(args.detail_file, "w"), indent=1)
            line["detail_file"] = os.path.relpath(args.detail_file, ROOT)
            line["detail_keys"] = {k: (list(v.keys())[:8] if isinstance(v, dict) else None) for k, v in detail.items()}
        except Exception:
            line["detail"] = detail
        # BASELINE configs 4 and 5 last, compact, so that they survive in the tail of the driver's record
        line["secondary"] = dict(
            msm=compact(msm_res, ["terms", "window_bits", "n_gpus", "ms", "ec_adds", "ec_adds_per_s", "accumulate_adds_per_s", "scaling", "error"]),
            one_proof_multi_gpu=one_proof,
            batch52=compact(b52, ["batch", "m", "n", "n_gpus", "proofs_per_s", "proofs_per_s_all_gpus", "prove_per_s", "verify_per_s",
                                  "all_verified", "host_threads", "single_proof_latency", "error"]))
        print(json.dumps(line), file=json_out)
        json_out.flush()
    if world > 1:
        dist.barrier()
    ctx_v.close()
    ctx.close()  # tears the library's communicator down as well
    if world > 1:
        dist.destroy_process_group()


def compact(d, keys):
    return None if d is None else {k: d[k] for k in keys if k in d}


def cpu_baseline_leg(m, n):
    """cpu_baseline of the GPU arm's line (rank 0, N = 1): a bounded sample on ONE core -- the reference is
    single-threaded (no `parallel` feature on any ark crate, Cargo.toml:8-21) -- in the faithful mode: the full
    prove + verify of a 1024-card deck of the same aspect, scaled to (m, n) by the point-scalar term count.  The
    full-size, all-threads measurement is `bench.py --impl reference`."""
    sm, sn = sample_shape(m, n, 1024)
    small = cpu_instance(sm, sn, 1, os.cpu_count() or 1)
    tp, tv = cpu_run(small, 1, 0)
    ws, wf = work_terms(sm, sn), work_terms(m, n)
    tp_full = tp * wf["prove_naive"] / ws["prove_naive"]
    tv_full = tv * wf["verify_naive"] / ws["verify_naive"]
    desc = (f"C restatement of the reference CPU path (oracle/c, not the Rust binary), faithful mode, 1 thread: full prove+verify of a "
            f"{sm * sn}-card deck (m={sm}, n={sn}) took {tp:.2f}s + {tv:.2f}s; scaled to (m={m}, n={n}) by point-scalar term count "
            f"(prove x{wf['prove_naive'] / ws['prove_naive']:.0f}, verify x{wf['verify_naive'] / ws['verify_naive']:.0f}) -> "
            f"{tp_full:.0f}s + {tv_full:.0f}s per proof.  The unscaled full-size run on all host threads is `--impl reference`")
    return dict(value=1.0 / (tp_full + tv_full), unit=UNIT, cores=1, kind="port", sample=desc)


def batch52_bench(pkg, ctx, torch, stream, batch, rank, host_threads=0):
    """BASELINE config "batch of independent 52-card proofs": `batch` decks of (m, n) = (4, 13)
    -- the reference's own test shape (tests.rs:178-179) -- proved by mp_shuffle_and_remask_batch
    and verified by mp_shuffle_verify_batch (this rank's shard of the proof-index split)."""
    import numpy as np
    m, n = 4, 13
    N = m * n
    inst = make_instance(ctx, m, n, seed=100 + rank)
    ctx2 = pkg.Context(ctx.device)
    ctx2.set_params(m, n, inst["enc_g"], inst["ck_g"], inst["ck_h"], inst["ghat"])
    rng = np.random.default_rng(200 + rank)
    npts = 2 * N * batch
    decks = ctx.dbg_scalar_mul(G64 * npts, rand_scalars(rng, npts))
    perms = np.concatenate([rng.permutation(N) for _ in range(batch)]).astype(np.uint32)
    rhos = rand_scalars(rng, N * batch)
    rands = rand_scalars(rng, (11 * m + 5 * n) * batch)
    perm_arr = perms.ctypes.data_as(ctypes.c_void_p)
    out_decks = ctypes.create_string_buffer(128 * N * batch)
    proofs = ctypes.create_string_buffer(pkg.lib.mp_proof_len(m, n) * batch)
    statuses = (ctypes.c_int32 * batch)()
    lib = pkg.lib
    res = {}
    for it in range(2):  # first pass warms the worker contexts
        t0 = time.perf_counter()
        pkg.check(ctx2.h, lib.mp_shuffle_and_remask_batch(ctx2.h, inst["pk"], decks, perm_arr, rhos, rands, batch, out_decks, proofs, host_threads))
        t1 = time.perf_counter()
        launches = ctx2.launches
        pkg.check(ctx2.h, lib.mp_shuffle_verify_batch(ctx2.h, inst["pk"], decks, out_decks, proofs, batch, statuses, host_threads))
        t2 = time.perf_counter()
        launches += ctx2.launches
        res = dict(batch=batch, m=m, n=n, prove_s=t1 - t0, verify_s=t2 - t1, proofs_per_s=batch / (t2 - t0),
                   prove_per_s=batch / (t1 - t0), verify_per_s=batch / (t2 - t1), all_verified=all(s == 0 for s in statuses),
                   gpu_launches=launches, host_threads=host_threads or os.cpu_count(), timing="host wall clock around the two C-ABI calls (host buffers)")
    # single 52-card proof latency (BASELINE config: one 52-card shuffle prove + verify on one GPU)
    one_deck, one_perm = decks[:128 * N], perms[:N].ctypes.data_as(ctypes.c_void_p)
    best = None
    for it in range(5):
        t0 = time.perf_counter()
        pkg.check(ctx2.h, lib.mp_shuffle_and_remask(ctx2.h, inst["pk"], one_deck, one_perm, rhos[:32 * N], rands[:32 * (11 * m + 5 * n)],
                                                    out_decks, proofs))
        t1 = time.perf_counter()
        ok = lib.mp_shuffle_verify(ctx2.h, inst["pk"], one_deck, out_decks, proofs)
        t2 = time.perf_counter()
        if it >= 2 and ok == 0:
            cur = dict(prove_ms=(t1 - t0) * 1e3, verify_ms=(t2 - t1) * 1e3)
            best = cur if best is None or cur["prove_ms"] + cur["verify_ms"] < best["prove_ms"] + best["verify_ms"] else best
    res["single_proof_latency"] = best
    ctx2.close()
    return res


def sigma_bench(pkg, ctx, n, cpu_baseline):
    """the sigma protocols either side of the shuffle, batched over a whole
    deck -- mask + proof, verify_mask, remask + proof, verify_remask, reveal token + proof (one player),
    verify_reveal for `n` cards through the C ABI with host buffers (host wall clock), beside the C
    restatement of the reference's per-card CPU path on a bounded sample."""
    import numpy as np
    rng = np.random.default_rng(300)
    sk = rand_scalars(rng, 1)
    pk = ctx.dbg_scalar_mul(G64, sk)
    cards = ctx.dbg_scalar_mul(G64 * n, rand_scalars(rng, n))
    r, om = rand_scalars(rng, n), [rand_scalars(rng, n) for _ in range(3)]
    alpha = rand_scalars(rng, n)
    res = {"cards": n, "timing": "host wall clock around each C-ABI call (host buffers)", "host_threads": os.cpu_count()}
    launches = 0
    for it in range(2):  # first pass warms the scratch buffers and the pk table
        t = [time.perf_counter()]
        masked, p1 = ctx.mask_batch(pk, cards, r, om[0]); t.append(time.perf_counter()); launches = ctx.launches
        s1 = ctx.verify_mask_batch(pk, cards, masked, p1); t.append(time.perf_counter()); launches += ctx.launches
        out, p2 = ctx.remask_prove_batch(pk, masked, alpha, om[1]); t.append(time.perf_counter()); launches += ctx.launches
        s2 = ctx.verify_remask_batch(pk, masked, out, p2); t.append(time.perf_counter()); launches += ctx.launches
        tok, p3 = ctx.reveal_batch(sk, pk, out, om[2]); t.append(time.perf_counter()); launches += ctx.launches
        s3 = ctx.verify_reveal_batch(pk, tok, out, p3); t.append(time.perf_counter()); launches += ctx.launches
    names = ["mask", "verify_mask", "remask", "verify_remask", "reveal", "verify_reveal"]
    for k, name in enumerate(names):
        res[name + "_per_s"] = n / (t[k + 1] - t[k])
    res["all_verified"] = not (any(s1) or any(s2) or any(s3))
    res["gpu_launches"] = launches
    res["proofs_per_s"] = 3 * n / (t[6] - t[0])  # three proofs made and checked per card
    if cpu_baseline:
        from oracle import c_oracle
        co = c_oracle.COracle(threads=1)
        k = min(n, 256)
        t0 = time.perf_counter()
        m2, q1 = co.mask_batch(G64, pk, cards[:64 * k], r[:32 * k], om[0][:32 * k])
        ok = co.verify_mask_batch(G64, pk, cards[:64 * k], m2, q1)
        o2, q2 = co.remask_prove_batch(G64, pk, m2, alpha[:32 * k], om[1][:32 * k])
        ok += co.verify_remask_batch(G64, pk, m2, o2, q2)
        t2, q3 = co.reveal_batch(G64, sk, pk, o2, om[2][:32 * k])
        ok += co.verify_reveal_batch(G64, pk, t2, o2, q3)
        dt = time.perf_counter() - t0
        same = (m2, q1, o2, q2, t2, q3) == (masked[:128 * k], p1[:160 * k], out[:128 * k], p2[:160 * k], tok[:64 * k], p3[:160 * k])
        res["cpu_baseline"] = dict(value=3 * k / dt, unit="proofs/s", cores=1, kind="port",
                                   sample=f"C restatement (oracle/c), 1 thread, the same six calls on the first {k} cards: {dt:.2f} s",
                                   bytes_identical_to_gpu=bool(same and not any(ok)))
    return res


def wire_bench(pkg, ctx, n_cards, cpu_baseline):
    """serialise a deck to the ark-serialize wire format (host) and
    deserialise it (GPU: one square root in F_p per point), host buffers, wall clock; beside the C
    restatement of the CPU path (Tonelli-Shanks as in ark-ff) on a bounded sample."""
    import numpy as np
    rng = np.random.default_rng(400)
    deck = ctx.dbg_scalar_mul(G64 * (2 * n_cards), rand_scalars(rng, 2 * n_cards))
    res = {"cards": n_cards, "points": 2 * n_cards, "timing": "host wall clock around each C-ABI call (host buffers)"}
    for it in range(2):
        t0 = time.perf_counter()
        ser = ctx.deck_serialize(deck)
        t1 = time.perf_counter()
        back = ctx.deck_deserialize(ser)
        t2 = time.perf_counter()
    res.update(serialize_points_per_s=2 * n_cards / (t1 - t0), deserialize_points_per_s=2 * n_cards / (t2 - t1),
               deserialize_ms=(t2 - t1) * 1e3, round_trip_ok=back == deck, gpu_launches=ctx.launches, wire_bytes=len(ser))
    if cpu_baseline:
        from oracle import c_oracle
        co = c_oracle.COracle(threads=1)
        k = min(2 * n_cards, 512)
        t0 = time.perf_counter()
        out, st = co.points_decompress(ser[8:8 + 32 * k])
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = dict(value=k / dt, unit="points/s", cores=1, kind="port",
                                   sample=f"C restatement (oracle/c, Tonelli-Shanks), 1 thread, first {k} points: {dt:.2f} s",
                                   bytes_identical_to_gpu=bool(out == deck[:64 * k] and not any(st)))
    return res


def bls12_377_bench(pkg, torch, dev, logn, cpu_baseline):
    """SURVEY.md section 8(f) rank 3 (group layer of the reference's second instantiation,
    `DLCards<ark_bls12_377::G1Projective>`, examples/parameter_selection.rs:25-29): a 2^logn-term G1 MSM with
    device-resident inputs (CUDA events on the context's stream, L2 flushed between iterations), the
    ciphertext MSM of a 2^16-card verify_shuffle, one batch of (m, n) = (128, 512) Pedersen commitments with
    host buffers, the field microbenchmarks, and the C restatement (oracle/c/bls12_377.c, ark-style
    Pippenger) on a bounded sample."""
    import numpy as np
    ctx = pkg.bls12_377.Context(dev.index or 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    rng = np.random.default_rng(11)
    g96 = bytes.fromhex(GEN_BLS12_377)
    n = 1 << logn
    nb = 4096
    base = ctx.dbg_scalar_mul(g96 * nb, rand_scalars(rng, nb))
    bases = torch.frombuffer(bytearray(base), dtype=torch.uint8).to(dev).repeat(n // nb).contiguous()
    scal_h = rand_scalars(rng, n)
    scal = torch.frombuffer(bytearray(scal_h), dtype=torch.uint8).to(dev)
    out = torch.zeros(192, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"curve": "BLS12-377 G1 (48-byte coordinates, 12 x 32-bit limbs)"}

    def timed(fn, reps=5):
        ts = []
        for it in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        return min(ts)

    c = 16 if logn >= 19 else (13 if logn >= 15 else 10)
    ctx.profile_enable(True)
    ctx.profile_collect()
    ms = timed(lambda: ctx.msm_g1_device(bases.data_ptr(), scal.data_ptr(), n, out.data_ptr(), c))
    acc_ms, acc_adds, acc_n = ctx.profile_collect()
    ctx.profile_enable(False)
    adds = ctx.last_msm_ec_adds
    launches = ctx.launches
    peak, peak_src = load_peaks()
    acc_rate = acc_adds / (acc_ms / 1e3) if acc_ms else None
    res["msm"] = dict(terms=n, window_bits=c, ms=ms, ec_adds=adds, ec_adds_per_s=adds / (ms / 1e3), terms_per_s=n / (ms / 1e3),
                      accumulate_ms_avg=acc_ms / max(acc_n, 1), accumulate_adds_per_s=acc_rate, gpu_launches=launches,
                      roofline=dict(bound="hbm", kernel="k_accumulate<1> (12-limb build)", unit="GB/s", peak=peak, peak_source=peak_src,
                                    achieved=(100.0 * acc_rate / 1e9 if acc_rate else None),
                                    frac=(100.0 * acc_rate / 1e9 / peak if acc_rate else None),
                                    note="algorithmic bytes per bucket addition: 96 B affine point + 4 B sorted index; "
                                         "integer-pipe bound (~10 field multiplications of 276 IMAD.WIDE each per addition)"),
                      result_x_prefix=bytes(out[:96].cpu().numpy().tobytes()).hex()[:16])
    # the verifier's dominant job at 2^16 cards: an N-term ciphertext MSM (2 components share digits and sort)
    nct = min(1 << 16, n // 2)
    msc = timed(lambda: ctx.ct_msm_device(bases.data_ptr(), scal.data_ptr(), nct, out.data_ptr(), 0))
    res["ct_msm"] = dict(ciphertexts=nct, window_bits=ctx.last_msm_window, ms=msc, ec_adds=ctx.last_msm_ec_adds,
                         ec_adds_per_s=ctx.last_msm_ec_adds / (msc / 1e3))
    # Pedersen commitments, (m, n) = (128, 512): m rows of n values over the constant key, host buffers
    m_rows, n_len = 128, 512
    ctx.set_commit_key(base[:96 * (n_len + 1)])
    vals, blinds = rand_scalars(rng, m_rows * n_len), rand_scalars(rng, m_rows)
    com = None
    for _ in range(3):
        t0 = time.perf_counter()
        com = ctx.pedersen_commit_batch(vals, blinds, n_len)
        dt = time.perf_counter() - t0
    res["pedersen"] = dict(rows=m_rows, length=n_len, ms=dt * 1e3, commitments_per_s=m_rows / dt, terms_per_s=m_rows * (n_len + 1) / dt,
                           gpu_launches=ctx.launches, timing="host wall clock around the C-ABI call (host buffers)")
    # The reference's own benchmark harness END TO END (examples/parameter_selection.rs:31-43,78-96): one 300-card deck over
    # this curve, (m, n) from (2, 150) to (30, 10); it times `shuffle_and_remask` with `Instant` and prints the proof
    # size.  Here: mp377_shuffle_and_remask (permute + remask + prove, host buffers, wall clock, best of 3) and
    # mp377_shuffle_verify of its output.  CPU column: the GROUP WORK of the same prover in the C restatement on one
    # core (per-term double-and-add inner products as proof-essentials does, Pippenger commitments), scaled by the
    # counts -- the restatement has no protocol driver over this curve, so transcript and scalar algebra are not in it.
    shapes = []
    pts300 = ctx.dbg_scalar_mul(g96 * 640, rand_scalars(rng, 640))
    deck300 = pts300[:96 * 600]
    for m_r, n_r in [(2, 150), (6, 50), (10, 30), (12, 25), (30, 10)]:
        Nr = m_r * n_r
        enc_g, ck_g, ck_h, ghat, pk = g96, pts300[96 * 600:96 * (600 + n_r)], pts300[96 * 630:96 * 631], pts300[96 * 631:96 * 632], pts300[96 * 632:96 * 633]
        if n_r > 30:
            ck_g = ctx.dbg_scalar_mul(g96 * n_r, rand_scalars(rng, n_r))
        perm = [int(v) for v in rng.permutation(Nr)]
        rho, rnd = rand_scalars(rng, Nr), rand_scalars(rng, 11 * m_r + 5 * n_r)
        ctx.set_params(m_r, n_r, enc_g, ck_g, ck_h, ghat)
        best_p = best_v = None
        for _ in range(3):
            t0 = time.perf_counter()
            deck2, proof = ctx.shuffle_and_remask(pk, deck300, perm, rho, rnd)
            t1 = time.perf_counter()
            ok = ctx.verify_shuffle(m_r, n_r, enc_g, ck_g, ck_h, ghat, pk, deck300, deck2, proof)
            t2 = time.perf_counter()
            best_p = t1 - t0 if best_p is None else min(best_p, t1 - t0)
            best_v = t2 - t1 if best_v is None else min(best_v, t2 - t1)
        entry = dict(m=m_r, n=n_r, cards=300, prove_ms=best_p * 1e3, verify_ms=best_v * 1e3, verified=(ok == 0),
                     proof_bytes_flat=len(proof), gpu_launches=ctx.launches)
        try:  # size of THIS repository's proof container (points compressed, no Vec prefixes); upstream's
            # `proof.serialized_size()` (parameter_selection.rs:93-96) adds 8 bytes per Vec field and is not reproduced
            entry["proof_bytes_repo_container"] = int(pkg.lib.mp377_proof_serialized_len(m_r, n_r))
        except Exception:
            pass
        if cpu_baseline:
            from oracle import c_oracle
            co = c_oracle.COracleBls12_377(threads=1)
            rows = rand_scalars(rng, n_r)
            t0 = time.perf_counter()
            co.msm(deck300[:192 * n_r], rows, 2, 0)   # one inner product <row of n ciphertexts, row of n scalars>
            t1 = time.perf_counter()
            co.msm(ck_h + ck_g, rand_scalars(rng, n_r + 1), 1, 1)   # one Pedersen commitment
            t2 = time.perf_counter()
            njobs, ncom = m_r * (m_r + 1), 4 * m_r + 5
            entry["cpu_group_work_ms_1core"] = ((t1 - t0) * njobs + (t2 - t1) * ncom) * 1e3
            entry["cpu_sample"] = (f"C restatement, 1 core: one of the {njobs} ciphertext inner products + one of the {ncom} "
                                   "commitments timed, scaled by the counts (group work of the prover only)")
        shapes.append(entry)
    res["reference_benchmark_shape"] = dict(
        source="examples/parameter_selection.rs:31-43,78-96 (BLS12-377 G1, 300 cards): shuffle_and_remask end to end + verify_shuffle",
        timing="host wall clock around mp377_shuffle_and_remask / mp377_shuffle_verify (host buffers), best of 3", shapes=shapes)
    # the rest of the trait over this curve at the reference benchmark's deck size and at a 2^14-card deck: the six
    # batched sigma calls (mod.rs:182-354) and deserialisation of a serialised deck (square roots + G1 membership)
    try:
        sig = {}
        sk = rand_scalars(rng, 1)
        pk1 = ctx.dbg_scalar_mul(g96, sk)
        for cards_n in (300, 16384):
            cards = ctx.dbg_scalar_mul(g96 * cards_n, rand_scalars(rng, cards_n))
            sc = [rand_scalars(rng, cards_n) for _ in range(6)]
            for it in range(2):
                t = [time.perf_counter()]
                masked, p1 = ctx.mask_batch(pk1, cards, sc[0], sc[1]); t.append(time.perf_counter())
                s1 = ctx.verify_mask_batch(pk1, cards, masked, p1); t.append(time.perf_counter())
                out, p2 = ctx.remask_prove_batch(pk1, masked, sc[2], sc[3]); t.append(time.perf_counter())
                s2 = ctx.verify_remask_batch(pk1, masked, out, p2); t.append(time.perf_counter())
                tok, p3 = ctx.reveal_batch(sk, pk1, out, sc[4]); t.append(time.perf_counter())
                s3 = ctx.verify_reveal_batch(pk1, tok, out, p3); t.append(time.perf_counter())
            names = ["mask", "verify_mask", "remask", "verify_remask", "reveal", "verify_reveal"]
            e = {nm + "_ms": (t[k + 1] - t[k]) * 1e3 for k, nm in enumerate(names)}
            e["all_verified"] = not (any(s1) or any(s2) or any(s3))
            e["proofs_per_s"] = 3 * cards_n / (t[6] - t[0])
            ser = pkg.bls12_377.deck_serialize(out)
            t0 = time.perf_counter()
            back = ctx.deck_deserialize(ser)
            e["deck_deserialize_ms"] = (time.perf_counter() - t0) * 1e3
            e["deck_round_trip_ok"] = back == out
            sig[str(cards_n)] = e
        res["sigma_and_wire"] = dict(timing="host wall clock around each mp377_* call (host buffers), second pass; verifiers and "
                                            "deserialisation include the G1 membership test of every point", cards=sig)
    except Exception as ex:
        res["sigma_and_wire"] = dict(error=repr(ex))
    mb = {}
    for which, name, iters in [(0, "fq_mul", 1000), (1, "madd", 300)]:
        best = 0
        for rep in range(2):
            t_ms, ops = ctx.dbg_bench(which, iters)
            best = max(best, ops / t_ms / 1e6)
        mb[name + "_G_per_s"] = best
    res["microbench"] = mb
    if cpu_baseline:
        from oracle import c_oracle
        co = c_oracle.COracleBls12_377(threads=1)
        k = min(n, 1 << 13)
        pts_h = base[:96 * min(k, nb)] * (k // min(k, nb))
        t0 = time.perf_counter()
        want = co.msm(pts_h, scal_h[:32 * k], 1, 1)
        dt = time.perf_counter() - t0
        got = ctx.msm_g1(pts_h, scal_h[:32 * k], 0)
        t0 = time.perf_counter()
        wantc = co.msm(base[:96 * (n_len + 1)], blinds[:32] + vals[:32 * n_len], 1, 1)
        dtc = time.perf_counter() - t0
        res["cpu_baseline"] = dict(value=k / dt, unit="MSM terms/s", cores=1, kind="port",
                                   sample=f"C restatement (oracle/c/bls12_377.c: ark-ff-style 6 x u64 Montgomery, ark-ec 0.3 "
                                          f"VariableBaseMSM), 1 thread, {k}-term G1 MSM: {dt:.2f} s; one ({n_len}+1)-term commitment: {dtc * 1e3:.1f} ms",
                                   commitments_per_s=1.0 / dtc,
                                   bytes_identical_to_gpu=bool(got == want and com[:96] == wantc))
    ctx.close()
    return res


def msm_microbench(ctx, torch, dev, stream, logn, pkg=None, world=1, rank=0):
    """2^logn-term variable-base MSM (BASELINE config 5).  world > 1: window-range split through the library's own
    multi-GPU entry point (mp_msm_g1_multi_device): every rank computes its share of the windows on replicated
    inputs, the 128-byte XYZZ partials are all-gathered over NCCL on the context's stream -- no host
    synchronisation in between -- and folded by one small kernel (strong scaling of ONE MSM)."""
    import numpy as np
    import torch.distributed as dist
    n = 1 << logn
    rng = np.random.default_rng(7)
    base = ctx.dbg_scalar_mul(G64 * 4096, rand_scalars(rng, 4096))
    bases = torch.frombuffer(bytearray(base), dtype=torch.uint8).to(dev).repeat(n // 4096).contiguous()
    scal = torch.frombuffer(bytearray(rand_scalars(rng, n)), dtype=torch.uint8).to(dev)
    out = torch.zeros(64, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    c = 16 if logn >= 19 else (13 if logn >= 15 else 10)
    times = []
    ctx.profile_enable(True)
    ctx.profile_collect()
    for it in range(6):
        flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if world == 1:
            ctx.msm_g1_device(bases.data_ptr(), scal.data_ptr(), n, out.data_ptr(), c)
        else:
            ctx.msm_g1_multi_device(bases.data_ptr(), scal.data_ptr(), n, out.data_ptr(), c)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        if it >= 3:
            times.append(ms)
    acc_ms, acc_adds, acc_n = ctx.profile_collect()
    ctx.profile_enable(False)
    best = min(times)
    adds = ctx.last_msm_ec_adds
    if world > 1:
        t = torch.tensor([best], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = t.item()
        a = torch.tensor([float(adds)], dtype=torch.float64, device=dev)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        adds = int(a.item())
    result = bytes(out.cpu().numpy().tobytes()).hex()
    return dict(terms=n, window_bits=c, n_gpus=world, ms=best, ec_adds=adds, ec_adds_per_s=adds / (best / 1e3),
                accumulate_ms_avg=acc_ms / max(acc_n, 1), accumulate_adds_per_s=acc_adds / (acc_ms / 1e3) if acc_ms else None,
                result_x_prefix=result[:16],
                scaling="strong (window-range split, NCCL all-gather of XYZZ partials on the stream + one fold kernel, mp_msm_g1_multi_device)" if world > 1 else "single GPU",
                note="device-resident canonical inputs -> canonical affine result, includes Montgomery conversion + on-curve check")


if __name__ == "__main__":
    main()

"""Host-side logic of bench.py that needs no GPU: work model, CPU-sample shapes, the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_work_model_matches_survey_appendix_c():
    w = bench.work_terms(128, 512)
    # diagonals 16 908 288, K1 262 144 (verify)
    assert 2 * 128 * 129 * 512 == 16908288 and w["prove_naive"] >= 16908288 + 2 * 65536
    assert w["verify_naive"] >= 4 * 65536 and w["verify_naive"] < 4 * 65536 + 4000


def test_sample_shapes_keep_aspect_and_budget():
    assert bench.sample_shape(128, 512, 1024) == (16, 64)
    assert bench.sample_shape(128, 512, 4096) == (32, 128)
    assert bench.sample_shape(128, 512, 16384) == (64, 256)
    assert bench.sample_shape(128, 512, 1 << 16) == (128, 512)
    assert bench.sample_shape(4, 13, 1024) == (4, 13)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--m", "4", "--n", "13"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "proofs/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["metric"] == bench.METRIC and d["higher_is_better"] is True
    # the arm RUNS the configuration it names: one full prove + verify, reported as one step, not extrapolated ...
    assert d["steps"] == 1 and d["warmup"] == 0 and d["steps_requested"] == 1 and d["extrapolated"] is False
    assert abs(d["ms_per_step"] - 1e3 * (d["cpu_baseline"]["prove_s"] + d["cpu_baseline"]["verify_s"])) < 1e-6
    assert d["ms_per_step"] / 1e3 <= d["wall_s"]
    # ... with the best-effort CPU number (Pippenger for the big sums) beside the faithful one
    assert d["cpu_baseline"]["best_effort"]["value"] > 0


def test_reference_arm_falls_back_to_a_scaled_prover_when_over_budget():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "5",
                          "--m", "8", "--n", "32", "--ref-budget-s", "0.001"], capture_output=True, text=True, timeout=600, check=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert d["extrapolated"] is True and "SCALED" in d["cpu_baseline"]["sample"] and d["steps"] == 1 and d["steps_requested"] == 20


def test_bls12_377_generator_constant():
    # the second-curve section of the bench builds its inputs from this constant (x || y, 48-byte LE coordinates)
    from oracle.py import bls12_377 as bls
    assert bytes.fromhex(bench.GEN_BLS12_377) == bls.point_to_bytes(bls.G)


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)

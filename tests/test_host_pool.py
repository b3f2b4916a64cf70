"""CPU stress test of the host thread pool behind the batched prover / verifier (csrc/host_pool.hpp, compiled with g++):
every item of every phase runs exactly once, `run` returns only after all of them, with varying item counts (0 and 1
included), thread counts that grow and shrink between phases, and several pools in use at once (one per worker
context)."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_pool_stress(tmp_path):
    out = str(tmp_path / "host_pool_test.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", out,
                           os.path.join(ROOT, "tests", "host", "host_pool_test.cpp")])
    lib = ctypes.CDLL(out)
    assert lib.h_pool_stress(1, 400, 1) == 0       # serial path
    assert lib.h_pool_stress(1, 3000, 16) == 0
    assert lib.h_pool_stress(4, 1500, 12) == 0     # four pools at once

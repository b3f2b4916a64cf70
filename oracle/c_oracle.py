"""ORACLE (test infrastructure only): ctypes loader for oracle/c/liboracle.so, the plain-C CPU
restatement of the reference's shuffle hot path (see oracle/c/oracle.c).  Imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "liboracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])


def load():
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    cp, u64, i32, vp = ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p
    lib.oc_msm.argtypes = [cp, cp, u64, i32, i32, cp]
    lib.oc_scalar_mul_batch.argtypes = [cp, cp, u64, cp]
    lib.oc_remask.argtypes = [cp, cp, cp, vp, cp, u64, cp]
    lib.oc_shuffle_prove.argtypes = [i32, i32, cp, cp, cp, cp, cp, cp, cp, vp, cp, cp, cp]
    lib.oc_shuffle_verify.argtypes = [i32, i32, cp, cp, cp, cp, cp, cp, cp, cp]
    lib.oc_pedersen_commit.argtypes = [i32, cp, cp, cp, i32, cp, cp]
    lib.oc_proof_len.restype = ctypes.c_size_t
    lib.oc_proof_len.argtypes = [i32, i32]
    lib.oc_prover_randomness_len.restype = ctypes.c_size_t
    lib.oc_prover_randomness_len.argtypes = [i32, i32]
    lib.oc_fr_mul.argtypes = [cp, cp, cp]
    lib.oc_fq_mul.argtypes = [cp, cp, cp]
    lib.oc_blake2s.argtypes = [cp, u64, cp]
    lib.oc_fs_challenges.argtypes = [cp, u64, i32, cp]
    lib.oc_on_curve.argtypes = [cp]
    lib.oc_mask_batch.argtypes = [cp, cp, cp, cp, cp, u64, cp, cp]
    lib.oc_verify_mask_batch.argtypes = [cp, cp, cp, cp, cp, u64, vp]
    lib.oc_remask_prove_batch.argtypes = [cp, cp, cp, cp, cp, u64, cp, cp]
    lib.oc_verify_remask_batch.argtypes = [cp, cp, cp, cp, cp, u64, vp]
    lib.oc_reveal_batch.argtypes = [cp, cp, cp, cp, cp, u64, cp, cp]
    lib.oc_verify_reveal_batch.argtypes = [cp, cp, cp, cp, cp, u64, vp]
    lib.oc_key_ownership_prove_batch.argtypes = [cp, cp, cp, cp, vp, cp, u64, cp]
    lib.oc_key_ownership_verify_batch.argtypes = [cp, cp, cp, vp, cp, u64, vp]
    lib.oc_points_compress.argtypes = [cp, u64, cp]
    lib.oc_points_decompress.argtypes = [cp, u64, cp, vp]
    lib.oracle_set_threads.argtypes = [i32]
    lib.oracle_set_msm_mode.argtypes = [i32]
    return lib


class COracle:
    """Byte-level interface (same layouts as include/mpshuffle.h)."""

    def __init__(self, threads=1, msm_mode=0):
        self.lib = load()
        self.set(threads, msm_mode)

    def set(self, threads=None, msm_mode=None):
        if threads is not None:
            self.lib.oracle_set_threads(threads)
        if msm_mode is not None:
            self.lib.oracle_set_msm_mode(msm_mode)

    @property
    def max_threads(self):
        return self.lib.oracle_max_threads()

    def msm(self, points, scalars, ncomp=1, mode=1):
        n = len(scalars) // 32
        out = ctypes.create_string_buffer(64 * ncomp)
        self.lib.oc_msm(points, scalars, n, ncomp, mode, out)
        return out.raw

    def scalar_mul_batch(self, base, scalars):
        """-> scalars[i] * base for every 32-byte scalar (synthetic-instance generator, all threads)"""
        n = len(scalars) // 32
        out = ctypes.create_string_buffer(64 * n)
        self.lib.oc_scalar_mul_batch(base, scalars, n, out)
        return out.raw

    def remask(self, enc_g, pk, deck, perm, rho):
        n = len(perm)
        arr = (ctypes.c_uint32 * n)(*perm)
        out = ctypes.create_string_buffer(128 * n)
        self.lib.oc_remask(enc_g, pk, deck, arr, rho, n, out)
        return out.raw

    def prove(self, m, n, enc_g, ck_g, ck_h, ghat, pk, deck, deck2, perm, rho, rand):
        arr = (ctypes.c_uint32 * len(perm))(*perm)
        out = ctypes.create_string_buffer(self.lib.oc_proof_len(m, n))
        rc = self.lib.oc_shuffle_prove(m, n, enc_g, ck_g, ck_h, ghat, pk, deck, deck2, arr, rho, rand, out)
        assert rc == 0
        return out.raw

    def verify(self, m, n, enc_g, ck_g, ck_h, ghat, pk, deck, deck2, proof):
        assert len(proof) == self.lib.oc_proof_len(m, n)
        return self.lib.oc_shuffle_verify(m, n, enc_g, ck_g, ck_h, ghat, pk, deck, deck2, proof)

    def commit(self, n, ck_g, ck_h, values, r):
        out = ctypes.create_string_buffer(64)
        self.lib.oc_pedersen_commit(n, ck_g, ck_h, values, len(values) // 32, r, out)
        return out.raw

    # ---- sigma protocols either side of the shuffle (oracle/py/sigma.py; reference mod.rs:132-354)
    def mask_batch(self, g, pk, cards, rs, omegas):
        n = len(rs) // 32
        masked, proofs = ctypes.create_string_buffer(128 * n), ctypes.create_string_buffer(160 * n)
        self.lib.oc_mask_batch(g, pk, cards, rs, omegas, n, masked, proofs)
        return masked.raw, proofs.raw

    def _statuses(self, fn, n, *args):
        st = (ctypes.c_int32 * max(n, 1))()
        fn(*args, n, st)
        return list(st)[:n]

    def verify_mask_batch(self, g, pk, cards, masked, proofs):
        return self._statuses(self.lib.oc_verify_mask_batch, len(proofs) // 160, g, pk, cards, masked, proofs)

    def remask_prove_batch(self, g, pk, deck, alphas, omegas):
        n = len(alphas) // 32
        out, proofs = ctypes.create_string_buffer(128 * n), ctypes.create_string_buffer(160 * n)
        self.lib.oc_remask_prove_batch(g, pk, deck, alphas, omegas, n, out, proofs)
        return out.raw, proofs.raw

    def verify_remask_batch(self, g, pk, deck, remasked, proofs):
        return self._statuses(self.lib.oc_verify_remask_batch, len(proofs) // 160, g, pk, deck, remasked, proofs)

    def reveal_batch(self, g, sk, pk, masked, omegas):
        n = len(omegas) // 32
        tokens, proofs = ctypes.create_string_buffer(64 * n), ctypes.create_string_buffer(160 * n)
        self.lib.oc_reveal_batch(g, sk, pk, masked, omegas, n, tokens, proofs)
        return tokens.raw, proofs.raw

    def verify_reveal_batch(self, g, pk, tokens, masked, proofs):
        return self._statuses(self.lib.oc_verify_reveal_batch, len(proofs) // 160, g, pk, tokens, masked, proofs)

    @staticmethod
    def _infos(infos):
        off = [0]
        for b in infos:
            off.append(off[-1] + len(b))
        return b"".join(infos), (ctypes.c_uint64 * len(off))(*off)

    def key_ownership_prove_batch(self, g, pks, sks, infos, omegas):
        n = len(infos)
        blob, off = self._infos(infos)
        proofs = ctypes.create_string_buffer(96 * n)
        self.lib.oc_key_ownership_prove_batch(g, pks, sks, blob, off, omegas, n, proofs)
        return proofs.raw

    def key_ownership_verify_batch(self, g, pks, infos, proofs):
        n = len(infos)
        blob, off = self._infos(infos)
        st = (ctypes.c_int32 * max(n, 1))()
        self.lib.oc_key_ownership_verify_batch(g, pks, blob, off, proofs, n, st)
        return list(st)[:n]

    # ---- wire format (oracle/py/wire.py)
    def points_compress(self, points):
        n = len(points) // 64
        out = ctypes.create_string_buffer(32 * n)
        self.lib.oc_points_compress(points, n, out)
        return out.raw

    def points_decompress(self, data):
        """-> (points n*64 with zeros for rejected items, statuses: 0 ok, 1 malformed, 2 not on the curve)"""
        n = len(data) // 32
        out = ctypes.create_string_buffer(64 * n)
        st = (ctypes.c_int32 * max(n, 1))()
        self.lib.oc_points_decompress(data, n, out, st)
        return out.raw, list(st)[:n]


class COracleBls12_377:
    """oracle/c/bls12_377.c: the CPU group arithmetic of the reference's BLS12-377 instantiation
    (48-byte coordinates, 96-byte points; include/mpshuffle_bls12_377.h layouts)."""

    def __init__(self, threads=1):
        so = os.path.join(_HERE, "c", "liboracle_bls12_377.so")
        if not os.path.exists(so):
            build()
        self.lib = ctypes.CDLL(so)
        cp, u64, i32 = ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int
        self.lib.oc377_msm.argtypes = [cp, cp, u64, i32, i32, cp]
        self.lib.oc377_fq_mul.argtypes = [cp, cp, cp]
        self.lib.oc377_set_threads.argtypes = [i32]
        self.lib.oc377_set_threads(threads)

    def set_threads(self, threads):
        self.lib.oc377_set_threads(threads)

    def msm(self, points: bytes, scalars: bytes, ncomp=1, mode=1) -> bytes:
        """mode 0: per-term double-and-add; mode 1: ark-ec 0.3 VariableBaseMSM."""
        n = len(scalars) // 32
        assert len(points) == 96 * n * ncomp
        out = ctypes.create_string_buffer(96 * ncomp)
        self.lib.oc377_msm(points, scalars, n, ncomp, mode, out)
        return out.raw

    def fq_mul(self, a: bytes, b: bytes) -> bytes:
        out = ctypes.create_string_buffer(48)
        self.lib.oc377_fq_mul(a, b, out)
        return out.raw

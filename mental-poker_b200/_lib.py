"""ctypes binding of libmpshuffle.so -- the same C ABI a Rust/cgo/JNI host would bind."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "lib", "libmpshuffle.so")

if not os.path.exists(lib_path):
    raise ImportError(
        f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  There is no CPU fallback for the hot path.")

lib = ctypes.CDLL(lib_path)

_vp, _i32, _u64, _cp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64, ctypes.c_char_p

# name -> (restype, argtypes); must list every symbol include/mpshuffle.h declares
SIGNATURES = {
    "mp_ctx_create": (_i32, [ctypes.POINTER(_vp), _i32]),
    "mp_ctx_destroy": (None, [_vp]),
    "mp_ctx_stream": (_vp, [_vp]),
    "mp_ctx_sync": (_i32, [_vp]),
    "mp_last_error_string": (_cp, [_vp]),
    "mp_verify_status_string": (_cp, [_i32]),
    "mp_last_kernel_launches": (_i32, [_vp]),
    "mp_msm_g1": (_i32, [_vp, _cp, _cp, _u64, _i32, _cp]),
    "mp_ct_msm": (_i32, [_vp, _cp, _cp, _u64, _i32, _cp]),
    "mp_msm_jobs": (_i32, [_vp, _cp, _u64, _i32, _cp, _u64, _vp, _u64, _i32, _cp]),
    "mp_msm_g1_device": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "mp_ct_msm_device": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "mp_msm_num_windows": (_i32, [_i32]),
    "mp_msm_g1_windows_device": (_i32, [_vp, _vp, _vp, _u64, _i32, _i32, _i32, _vp]),
    "mp_last_msm_ec_adds": (_u64, [_vp]),
    "mp_last_msm_window": (_i32, [_vp]),
    "mp_comm_unique_id": (_i32, [_cp]),
    "mp_comm_init": (_i32, [_vp, _i32, _i32, _cp]),
    "mp_comm_destroy": (_i32, [_vp]),
    "mp_comm_size": (_i32, [_vp]),
    "mp_comm_rank": (_i32, [_vp]),
    "mp_msm_g1_multi_device": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "mp_shuffle_verify_batch_multi": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp_shuffle_and_remask_multi": (_i32, [_vp, _cp, _cp, _vp, _cp, _cp, _cp, _cp]),
    "mp_shuffle_verify_multi": (_i32, [_vp, _cp, _cp, _cp, _cp]),
    "mp_ctx_set_params": (_i32, [_vp, _i32, _i32, _cp, _cp, _cp, _cp]),
    "mp_params_m": (_i32, [_vp]),
    "mp_params_n": (_i32, [_vp]),
    "mp_proof_len": (_u64, [_i32, _i32]),
    "mp_prover_randomness_len": (_u64, [_i32, _i32]),
    "mp_remask_batch": (_i32, [_vp, _cp, _cp, _vp, _cp, _u64, _cp]),
    "mp_pedersen_commit_batch": (_i32, [_vp, _cp, _cp, _u64, _u64, _cp]),
    "mp_shuffle_prove": (_i32, [_vp, _cp, _cp, _cp, _vp, _cp, _cp, _cp]),
    "mp_shuffle_and_remask": (_i32, [_vp, _cp, _cp, _vp, _cp, _cp, _cp, _cp]),
    "mp_shuffle_verify": (_i32, [_vp, _cp, _cp, _cp, _cp]),
    "mp_shuffle_and_remask_batch": (_i32, [_vp, _cp, _cp, _vp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp_shuffle_verify_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp_shuffle_and_remask_batch_resident": (_i32, [_vp, _cp, _cp, _vp, _cp, _cp, _u64, _cp, _cp, _i32, _vp]),
    "mp_shuffle_verify_batch_resident": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32, _vp, _vp]),
    "mp_shuffle_verify_resident": (_i32, [_vp, _cp, _cp, _cp, _cp, _vp, _vp]),
    "mp_shuffle_and_remask_resident": (_i32, [_vp, _cp, _cp, _vp, _cp, _cp, _cp, _cp, _vp]),
    "mp_shuffle_prove_resident": (_i32, [_vp, _cp, _cp, _cp, _vp, _cp, _cp, _cp, _vp]),
    "mp_mask_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp_verify_mask_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp_remask_prove_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp_verify_remask_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp_reveal_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp_verify_reveal_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp_key_ownership_prove_batch": (_i32, [_vp, _cp, _cp, _cp, ctypes.POINTER(_u64), _cp, _u64, _cp, _i32]),
    "mp_key_ownership_verify_batch": (_i32, [_vp, _cp, _cp, ctypes.POINTER(_u64), _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp_points_compress": (_i32, [_cp, _u64, _cp]),
    "mp_points_decompress": (_i32, [_vp, _cp, _u64, _cp, ctypes.POINTER(_i32)]),
    "mp_deck_serialized_len": (_u64, [_u64]),
    "mp_deck_serialize": (_i32, [_cp, _u64, _cp]),
    "mp_deck_deserialize": (_i32, [_vp, _cp, _u64, _cp, ctypes.POINTER(_u64)]),
    "mp_proof_serialized_len": (_u64, [_i32, _i32]),
    "mp_proof_serialize": (_i32, [_i32, _i32, _cp, _cp]),
    "mp_proof_deserialize": (_i32, [_vp, _i32, _i32, _cp, _cp]),
    "mp_profile_enable": (_i32, [_vp, _i32]),
    "mp_profile_collect_dominant": (_i32, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "mp_profile_collect": (_i32, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "mp_dbg_fq_mul": (_i32, [_vp, _cp, _cp, _u64, _cp]),
    "mp_dbg_point_add": (_i32, [_vp, _cp, _cp, _u64, _cp]),
    "mp_dbg_scalar_mul": (_i32, [_vp, _cp, _cp, _u64, _cp]),
    "mp_dbg_transcript_ms": (ctypes.c_double, [_u64]),
    "mp_dbg_bench": (_i32, [_vp, _i32, _i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class MpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mpshuffle error {code}: {msg}")
        self.code = code


def check(ctx_handle, code):
    if code < 0:
        msg = lib.mp_last_error_string(ctx_handle)
        raise MpError(code, msg.decode() if msg else "")
    return code


class Context:
    """Owns one `mp_ctx` (one CUDA device, one stream)."""

    def __init__(self, device=0):
        h = _vp()
        rc = lib.mp_ctx_create(ctypes.byref(h), device)
        if rc != 0:
            raise MpError(rc, f"mp_ctx_create(device={device}) failed: no usable CUDA device "
                              "(the engine has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib.mp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- plumbing
    @property
    def stream(self):
        return lib.mp_ctx_stream(self.h)

    def sync(self):
        check(self.h, lib.mp_ctx_sync(self.h))

    @property
    def launches(self):
        return lib.mp_last_kernel_launches(self.h)

    # --- multi-GPU (NCCL communicator behind the C ABI)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = lib.mp_comm_unique_id(buf)
        if rc != 0:
            raise MpError(rc, "mp_comm_unique_id failed (NCCL not available)")
        return buf.raw

    def comm_init(self, nranks, rank, uid: bytes):
        check(self.h, lib.mp_comm_init(self.h, nranks, rank, uid))

    def comm_destroy(self):
        check(self.h, lib.mp_comm_destroy(self.h))

    def msm_g1_multi_device(self, d_bases, d_scalars, n, d_out, window_bits=0):
        check(self.h, lib.mp_msm_g1_multi_device(self.h, d_bases, d_scalars, n, window_bits, d_out))

    def shuffle_and_remask_multi(self, pk, deck, perm, rho, rand):
        N = self.m * self.n
        arr = (ctypes.c_uint32 * N)(*perm)
        deck2 = ctypes.create_string_buffer(128 * N)
        proof = ctypes.create_string_buffer(lib.mp_proof_len(self.m, self.n))
        check(self.h, lib.mp_shuffle_and_remask_multi(self.h, pk, deck, arr, rho, rand, deck2, proof))
        return deck2.raw, proof.raw

    def verify_shuffle_multi(self, pk, deck, deck2, proof) -> int:
        return check(self.h, lib.mp_shuffle_verify_multi(self.h, pk, deck, deck2, proof))

    def verify_shuffle_batch_multi(self, pk, decks, decks2, proofs, nranks, host_threads=0):
        plen = lib.mp_proof_len(self.m, self.n)
        B = len(proofs) // plen
        st = (_i32 * (B * nranks))()
        check(self.h, lib.mp_shuffle_verify_batch_multi(self.h, pk, decks, decks2, proofs, B, st, host_threads))
        return list(st)

    # --- MSM (host buffers)
    def msm_g1(self, bases: bytes, scalars: bytes, window_bits=0) -> bytes:
        n = len(scalars) // 32
        assert len(bases) == 64 * n and len(scalars) == 32 * n
        out = ctypes.create_string_buffer(64)
        check(self.h, lib.mp_msm_g1(self.h, bases, scalars, n, window_bits, out))
        return out.raw

    def ct_msm(self, deck: bytes, scalars: bytes, window_bits=0) -> bytes:
        n = len(scalars) // 32
        assert len(deck) == 128 * n
        out = ctypes.create_string_buffer(128)
        check(self.h, lib.mp_ct_msm(self.h, deck, scalars, n, window_bits, out))
        return out.raw

    def msm_jobs(self, points: bytes, scalars: bytes, jobs, ncomp=1, window_bits=0) -> bytes:
        """jobs: list of (scalar_off, point_off, len).  -> len(jobs) * ncomp points (MultiExponentiationArgument's
        batched inner products; SURVEY.md 8(b) mp_msm_batch_shared_bases)"""
        flat = (ctypes.c_uint32 * (3 * len(jobs)))(*[v for j in jobs for v in j])
        out = ctypes.create_string_buffer(64 * ncomp * max(len(jobs), 1))
        check(self.h, lib.mp_msm_jobs(self.h, points, len(points) // (64 * ncomp), ncomp, scalars, len(scalars) // 32,
                                      flat, len(jobs), window_bits, out))
        return out.raw[:64 * ncomp * len(jobs)]

    # --- MSM (device pointers, asynchronous on self.stream)
    def msm_g1_device(self, d_bases, d_scalars, n, d_out, window_bits=0):
        check(self.h, lib.mp_msm_g1_device(self.h, d_bases, d_scalars, n, window_bits, d_out))

    def ct_msm_device(self, d_deck, d_scalars, n, d_out, window_bits=0):
        check(self.h, lib.mp_ct_msm_device(self.h, d_deck, d_scalars, n, window_bits, d_out))

    def msm_g1_windows_device(self, d_bases, d_scalars, n, d_out, window_bits, w_begin, w_count):
        check(self.h, lib.mp_msm_g1_windows_device(self.h, d_bases, d_scalars, n, window_bits, w_begin, w_count, d_out))

    @property
    def last_msm_ec_adds(self):
        return lib.mp_last_msm_ec_adds(self.h)

    @property
    def last_msm_window(self):
        return lib.mp_last_msm_window(self.h)

    # --- shuffle protocol: mirrors BarnettSmartProtocol::{setup, shuffle_and_remask,
    #     verify_shuffle} (reference src/lib.rs:74-78,181-197) at the byte level
    def set_params(self, m, n, enc_g: bytes, ck_g: bytes, ck_h: bytes, ghat: bytes):
        assert len(ck_g) == 64 * n
        check(self.h, lib.mp_ctx_set_params(self.h, m, n, enc_g, ck_g, ck_h, ghat))
        self.m, self.n = m, n

    def remask(self, pk: bytes, deck: bytes, perm, rho: bytes) -> bytes:
        n = len(perm)
        arr = (ctypes.c_uint32 * n)(*perm)
        out = ctypes.create_string_buffer(128 * n)
        check(self.h, lib.mp_remask_batch(self.h, pk, deck, arr, rho, n, out))
        return out.raw

    def commit_batch(self, values: bytes, blinds: bytes, length: int) -> bytes:
        k = len(blinds) // 32
        assert len(values) == 32 * k * length
        out = ctypes.create_string_buffer(64 * k)
        check(self.h, lib.mp_pedersen_commit_batch(self.h, values, blinds, k, length, out))
        return out.raw

    def shuffle_prove(self, pk, deck, deck2, perm, rho, rand) -> bytes:
        arr = (ctypes.c_uint32 * len(perm))(*perm)
        assert len(rand) == 32 * lib.mp_prover_randomness_len(self.m, self.n)
        out = ctypes.create_string_buffer(lib.mp_proof_len(self.m, self.n))
        check(self.h, lib.mp_shuffle_prove(self.h, pk, deck, deck2, arr, rho, rand, out))
        return out.raw

    def shuffle_and_remask(self, pk, deck, perm, rho, rand):
        """-> (shuffled deck bytes, proof bytes)"""
        N = self.m * self.n
        assert len(perm) == N and len(deck) == 128 * N and len(rho) == 32 * N
        arr = (ctypes.c_uint32 * N)(*perm)
        deck2 = ctypes.create_string_buffer(128 * N)
        proof = ctypes.create_string_buffer(lib.mp_proof_len(self.m, self.n))
        check(self.h, lib.mp_shuffle_and_remask(self.h, pk, deck, arr, rho, rand, deck2, proof))
        return deck2.raw, proof.raw

    def verify_shuffle(self, pk, deck, deck2, proof) -> int:
        """0 = Ok(()), > 0 = CryptoError::ProofVerificationError(status string)."""
        N = self.m * self.n
        assert len(deck) == 128 * N and len(deck2) == 128 * N
        assert len(proof) == lib.mp_proof_len(self.m, self.n)
        return check(self.h, lib.mp_shuffle_verify(self.h, pk, deck, deck2, proof))

    def shuffle_and_remask_batch(self, pk, decks, perms, rhos, rands, host_threads=0):
        """perms: flat list of B*N indices.  -> (shuffled decks bytes, proofs bytes)"""
        N = self.m * self.n
        B = len(perms) // N
        assert len(decks) == 128 * N * B and len(rhos) == 32 * N * B
        assert len(rands) == 32 * B * lib.mp_prover_randomness_len(self.m, self.n)
        arr = (ctypes.c_uint32 * (N * B))(*perms)
        out = ctypes.create_string_buffer(128 * N * B)
        proofs = ctypes.create_string_buffer(lib.mp_proof_len(self.m, self.n) * B)
        check(self.h, lib.mp_shuffle_and_remask_batch(self.h, pk, decks, arr, rhos, rands, B, out, proofs, host_threads))
        return out.raw, proofs.raw

    def verify_shuffle_batch(self, pk, decks, decks2, proofs, host_threads=0):
        """-> list of per-proof statuses"""
        plen = lib.mp_proof_len(self.m, self.n)
        B = len(proofs) // plen
        assert len(decks) == len(decks2) == 128 * self.m * self.n * B and len(proofs) == plen * B
        st = (_i32 * B)()
        check(self.h, lib.mp_shuffle_verify_batch(self.h, pk, decks, decks2, proofs, B, st, host_threads))
        return list(st)

    # --- batched sigma protocols either side of the shuffle (reference mod.rs:132-354)
    def _statuses(self, fn, n, *args, host_threads=0):
        st = (_i32 * max(n, 1))()
        check(self.h, fn(self.h, *args, n, st, host_threads))
        return list(st)[:n]

    def mask_batch(self, shared_key, cards, r, omega, host_threads=0):
        """-> (masked cards n*128, Chaum-Pedersen proofs n*160)"""
        n = len(r) // 32
        assert len(cards) == 64 * n and len(omega) == 32 * n
        out, proofs = ctypes.create_string_buffer(128 * n), ctypes.create_string_buffer(160 * n)
        check(self.h, lib.mp_mask_batch(self.h, shared_key, cards, r, omega, n, out, proofs, host_threads))
        return out.raw, proofs.raw

    def verify_mask_batch(self, shared_key, cards, masked, proofs, host_threads=0):
        n = len(proofs) // 160
        assert len(cards) == 64 * n and len(masked) == 128 * n
        return self._statuses(lib.mp_verify_mask_batch, n, shared_key, cards, masked, proofs, host_threads=host_threads)

    def remask_prove_batch(self, shared_key, deck, alpha, omega, host_threads=0):
        n = len(alpha) // 32
        assert len(deck) == 128 * n and len(omega) == 32 * n
        out, proofs = ctypes.create_string_buffer(128 * n), ctypes.create_string_buffer(160 * n)
        check(self.h, lib.mp_remask_prove_batch(self.h, shared_key, deck, alpha, omega, n, out, proofs, host_threads))
        return out.raw, proofs.raw

    def verify_remask_batch(self, shared_key, deck, remasked, proofs, host_threads=0):
        n = len(proofs) // 160
        assert len(deck) == len(remasked) == 128 * n
        return self._statuses(lib.mp_verify_remask_batch, n, shared_key, deck, remasked, proofs, host_threads=host_threads)

    def reveal_batch(self, sk, pk, masked, omega, host_threads=0):
        """-> (reveal tokens n*64, Chaum-Pedersen proofs n*160) of one player for n masked cards"""
        n = len(omega) // 32
        assert len(masked) == 128 * n and len(sk) == 32 and len(pk) == 64
        tokens, proofs = ctypes.create_string_buffer(64 * n), ctypes.create_string_buffer(160 * n)
        check(self.h, lib.mp_reveal_batch(self.h, sk, pk, masked, omega, n, tokens, proofs, host_threads))
        return tokens.raw, proofs.raw

    def verify_reveal_batch(self, pk, tokens, masked, proofs, host_threads=0):
        n = len(proofs) // 160
        assert len(tokens) == 64 * n and len(masked) == 128 * n
        return self._statuses(lib.mp_verify_reveal_batch, n, pk, tokens, masked, proofs, host_threads=host_threads)

    @staticmethod
    def _infos(infos):
        off = [0]
        for b in infos:
            off.append(off[-1] + len(b))
        return b"".join(infos), (_u64 * len(off))(*off)

    def key_ownership_prove_batch(self, pks, sks, infos, omega, host_threads=0):
        n = len(infos)
        assert len(pks) == 64 * n and len(sks) == 32 * n and len(omega) == 32 * n
        blob, off = self._infos(infos)
        proofs = ctypes.create_string_buffer(96 * n)
        check(self.h, lib.mp_key_ownership_prove_batch(self.h, pks, sks, blob, off, omega, n, proofs, host_threads))
        return proofs.raw

    def key_ownership_verify_batch(self, pks, infos, proofs, host_threads=0):
        n = len(infos)
        assert len(pks) == 64 * n and len(proofs) == 96 * n
        blob, off = self._infos(infos)
        st = (_i32 * max(n, 1))()
        check(self.h, lib.mp_key_ownership_verify_batch(self.h, pks, blob, off, proofs, n, st, host_threads))
        return list(st)[:n]

    # --- wire format (ark-serialize 0.3 compressed encodings; decompression on the GPU)
    @staticmethod
    def points_compress(points: bytes) -> bytes:
        n = len(points) // 64
        out = ctypes.create_string_buffer(32 * n)
        assert lib.mp_points_compress(points, n, out) == 0
        return out.raw

    def points_decompress(self, data: bytes, want_statuses=False):
        """-> points (n*64); with want_statuses=True -> (points, statuses, return code) without raising"""
        n = len(data) // 32
        out = ctypes.create_string_buffer(64 * n)
        st = (_i32 * max(n, 1))()
        rc = lib.mp_points_decompress(self.h, data, n, out, st)
        if want_statuses:
            return out.raw, list(st)[:n], rc
        check(self.h, rc)
        return out.raw

    @staticmethod
    def deck_serialize(deck: bytes) -> bytes:
        n = len(deck) // 128
        out = ctypes.create_string_buffer(lib.mp_deck_serialized_len(n))
        assert lib.mp_deck_serialize(deck, n, out) == 0
        return out.raw

    def deck_deserialize(self, data: bytes) -> bytes:
        cap = max((len(data) - 8) // 64, 0)
        out = ctypes.create_string_buffer(128 * max(cap, 1))
        n = _u64(cap)
        check(self.h, lib.mp_deck_deserialize(self.h, data, len(data), out, ctypes.byref(n)))
        return out.raw[:128 * n.value]

    @staticmethod
    def proof_serialize(m, n, proof: bytes) -> bytes:
        out = ctypes.create_string_buffer(lib.mp_proof_serialized_len(m, n))
        assert lib.mp_proof_serialize(m, n, proof, out) == 0
        return out.raw

    def proof_deserialize(self, m, n, data: bytes) -> bytes:
        assert len(data) == lib.mp_proof_serialized_len(m, n)
        out = ctypes.create_string_buffer(lib.mp_proof_len(m, n))
        check(self.h, lib.mp_proof_deserialize(self.h, m, n, data, out))
        return out.raw

    @staticmethod
    def status_string(code):
        return lib.mp_verify_status_string(code).decode()

    # --- measurement
    def profile_enable(self, on=True):
        check(self.h, lib.mp_profile_enable(self.h, 1 if on else 0))

    def profile_collect_dominant(self):
        """-> (ms, bucket additions, launches) of the dominant accumulate launches since the last collect"""
        ms, adds, n = ctypes.c_double(), _u64(), _u64()
        check(self.h, lib.mp_profile_collect_dominant(self.h, ctypes.byref(ms), ctypes.byref(adds), ctypes.byref(n)))
        return ms.value, adds.value, n.value

    def profile_collect(self):
        """-> (accumulate-kernel ms, bucket additions, launches) since the last collect"""
        ms, adds, n = ctypes.c_double(), _u64(), _u64()
        check(self.h, lib.mp_profile_collect(self.h, ctypes.byref(ms), ctypes.byref(adds), ctypes.byref(n)))
        return ms.value, adds.value, n.value

    # --- debug hooks
    def dbg_fq_mul(self, a: bytes, b: bytes) -> bytes:
        n = len(a) // 32
        out = ctypes.create_string_buffer(32 * n)
        check(self.h, lib.mp_dbg_fq_mul(self.h, a, b, n, out))
        return out.raw

    def dbg_point_add(self, p: bytes, q: bytes) -> bytes:
        n = len(p) // 64
        out = ctypes.create_string_buffer(64 * n)
        check(self.h, lib.mp_dbg_point_add(self.h, p, q, n, out))
        return out.raw

    def dbg_scalar_mul(self, p: bytes, k: bytes) -> bytes:
        n = len(p) // 64
        out = ctypes.create_string_buffer(64 * n)
        check(self.h, lib.mp_dbg_scalar_mul(self.h, p, k, n, out))
        return out.raw

    def dbg_bench(self, which, iters):
        ms, ops = ctypes.c_float(), ctypes.c_double()
        check(self.h, lib.mp_dbg_bench(self.h, which, iters, ctypes.byref(ms), ctypes.byref(ops)))
        return ms.value, ops.value

"""ctypes binding of the BLS12-377 G1 entry points of libmpshuffle.so (include/mpshuffle_bls12_377.h):
the group layer of the reference's second instantiation, `DLCards<ark_bls12_377::G1Projective>`
(reference examples/parameter_selection.rs:25-29).  Same conventions as `_lib.Context`, with 48-byte
coordinates / 96-byte points.  No CPU fallback: the context needs a CUDA device."""
import ctypes

from ._lib import lib, MpError

_vp, _i32, _u64, _cp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64, ctypes.c_char_p
_pd, _pu64 = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_u64)

FQ_BYTES, POINT_BYTES, SCALAR_BYTES = 48, 96, 32
CP_PROOF_BYTES, SCHNORR_PROOF_BYTES = 2 * POINT_BYTES + 32, POINT_BYTES + 32   # a | b | r ; commit | opening

# name -> (restype, argtypes); must list every symbol include/mpshuffle_bls12_377.h declares
SIGNATURES = {
    "mp377_ctx_create": (_i32, [ctypes.POINTER(_vp), _i32]),
    "mp377_ctx_destroy": (None, [_vp]),
    "mp377_ctx_stream": (_vp, [_vp]),
    "mp377_ctx_sync": (_i32, [_vp]),
    "mp377_last_error_string": (_cp, [_vp]),
    "mp377_last_kernel_launches": (_i32, [_vp]),
    "mp377_last_msm_ec_adds": (_u64, [_vp]),
    "mp377_last_msm_window": (_i32, [_vp]),
    "mp377_msm_num_windows": (_i32, [_i32]),
    "mp377_msm_g1": (_i32, [_vp, _cp, _cp, _u64, _i32, _cp]),
    "mp377_ct_msm": (_i32, [_vp, _cp, _cp, _u64, _i32, _cp]),
    "mp377_msm_g1_device": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "mp377_ct_msm_device": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "mp377_points_compress": (_i32, [_cp, _u64, _cp]),
    "mp377_deck_serialized_len": (_u64, [_u64]),
    "mp377_deck_serialize": (_i32, [_cp, _u64, _cp]),
    "mp377_proof_serialized_len": (_u64, [_i32, _i32]),
    "mp377_proof_serialize": (_i32, [_i32, _i32, _cp, _cp]),
    "mp377_points_decompress": (_i32, [_vp, _cp, _u64, _cp, ctypes.POINTER(_i32)]),
    "mp377_deck_deserialize": (_i32, [_vp, _cp, _u64, _cp, _pu64]),
    "mp377_proof_deserialize": (_i32, [_vp, _i32, _i32, _cp, _cp]),
    "mp377_proof_len": (_u64, [_i32, _i32]),
    "mp377_ctx_set_params": (_i32, [_vp, _i32, _i32, _cp, _cp, _cp, _cp]),
    "mp377_prover_randomness_len": (_u64, [_i32, _i32]),
    "mp377_remask_batch": (_i32, [_vp, _cp, _cp, _vp, _cp, _u64, _cp]),
    "mp377_shuffle_and_remask": (_i32, [_vp, _cp, _cp, _vp, _cp, _cp, _cp, _cp]),
    "mp377_shuffle_and_remask_batch": (_i32, [_vp, _cp, _cp, _vp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp377_subgroup_check": (_i32, [_vp, _cp, _u64, ctypes.POINTER(_i32)]),
    "mp377_mask_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp377_verify_mask_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp377_remask_prove_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp377_verify_remask_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp377_reveal_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, _cp, _cp, _i32]),
    "mp377_verify_reveal_batch": (_i32, [_vp, _cp, _cp, _cp, _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp377_key_ownership_prove_batch": (_i32, [_vp, _cp, _cp, _cp, ctypes.POINTER(_u64), _cp, _u64, _cp, _i32]),
    "mp377_key_ownership_verify_batch": (_i32, [_vp, _cp, _cp, ctypes.POINTER(_u64), _cp, _u64, ctypes.POINTER(_i32), _i32]),
    "mp377_shuffle_verify": (_i32, [_vp, _i32, _i32, _cp, _cp, _cp, _cp, _cp, _cp, _cp, _cp]),
    "mp377_msm_jobs": (_i32, [_vp, _cp, _u64, _i32, _cp, _u64, ctypes.POINTER(ctypes.c_uint32), _u64, _i32, _cp]),
    "mp377_msm_g1_windows_device": (_i32, [_vp, _vp, _vp, _u64, _i32, _i32, _i32, _vp]),
    "mp377_set_commit_key": (_i32, [_vp, _cp, _u64]),
    "mp377_pedersen_commit_batch": (_i32, [_vp, _cp, _cp, _u64, _u64, _cp]),
    "mp377_profile_enable": (_i32, [_vp, _i32]),
    "mp377_profile_collect": (_i32, [_vp, _pd, _pu64, _pu64]),
    "mp377_dbg_fq_mul": (_i32, [_vp, _cp, _cp, _u64, _cp]),
    "mp377_dbg_point_add": (_i32, [_vp, _cp, _cp, _u64, _cp]),
    "mp377_dbg_scalar_mul": (_i32, [_vp, _cp, _cp, _u64, _cp]),
    "mp377_dbg_bench": (_i32, [_vp, _i32, _i32, ctypes.POINTER(ctypes.c_float), _pd]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def _check(h, code):
    if code < 0:
        msg = lib.mp377_last_error_string(h)
        raise MpError(code, msg.decode() if msg else "")
    return code


# --- wire format, serialising half: host byte handling, no context, no GPU
def points_compress(points: bytes) -> bytes:
    n = len(points) // POINT_BYTES
    out = ctypes.create_string_buffer(max(1, FQ_BYTES * n))
    assert lib.mp377_points_compress(points, n, out) == 0
    return out.raw[:FQ_BYTES * n]


def deck_serialize(deck: bytes) -> bytes:
    n = len(deck) // (2 * POINT_BYTES)
    out = ctypes.create_string_buffer(lib.mp377_deck_serialized_len(n))
    assert lib.mp377_deck_serialize(deck, n, out) == 0
    return out.raw


def proof_serialize(m: int, n: int, proof: bytes) -> bytes:
    assert len(proof) == lib.mp377_proof_len(m, n)
    out = ctypes.create_string_buffer(lib.mp377_proof_serialized_len(m, n))
    assert lib.mp377_proof_serialize(m, n, proof, out) == 0
    return out.raw


class Context:
    """Owns one `mp377_ctx` (one CUDA device, one stream, one MSM workspace)."""

    def __init__(self, device=0):
        h = _vp()
        rc = lib.mp377_ctx_create(ctypes.byref(h), device)
        if rc != 0:
            raise MpError(rc, f"mp377_ctx_create(device={device}) failed: no usable CUDA device "
                              "(the engine has no CPU fallback)")
        self.h, self.device = h, device

    def close(self):
        if self.h:
            lib.mp377_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib.mp377_ctx_stream(self.h)

    def sync(self):
        _check(self.h, lib.mp377_ctx_sync(self.h))

    @property
    def launches(self):
        return lib.mp377_last_kernel_launches(self.h)

    @property
    def last_msm_ec_adds(self):
        return lib.mp377_last_msm_ec_adds(self.h)

    @property
    def last_msm_window(self):
        return lib.mp377_last_msm_window(self.h)

    # --- MSM
    def msm_g1(self, bases: bytes, scalars: bytes, window_bits=0) -> bytes:
        n = len(scalars) // SCALAR_BYTES
        assert len(bases) == POINT_BYTES * n and len(scalars) == SCALAR_BYTES * n
        out = ctypes.create_string_buffer(POINT_BYTES)
        _check(self.h, lib.mp377_msm_g1(self.h, bases, scalars, n, window_bits, out))
        return out.raw

    def ct_msm(self, deck: bytes, scalars: bytes, window_bits=0) -> bytes:
        n = len(scalars) // SCALAR_BYTES
        assert len(deck) == 2 * POINT_BYTES * n
        out = ctypes.create_string_buffer(2 * POINT_BYTES)
        _check(self.h, lib.mp377_ct_msm(self.h, deck, scalars, n, window_bits, out))
        return out.raw

    def msm_g1_device(self, d_bases, d_scalars, n, d_out, window_bits=0):
        _check(self.h, lib.mp377_msm_g1_device(self.h, d_bases, d_scalars, n, window_bits, d_out))

    def ct_msm_device(self, d_deck, d_scalars, n, d_out, window_bits=0):
        _check(self.h, lib.mp377_ct_msm_device(self.h, d_deck, d_scalars, n, window_bits, d_out))

    def msm_jobs(self, points: bytes, scalars: bytes, jobs, ncomp=1, window_bits=0) -> bytes:
        """jobs: list of (scalar_off, point_off, len); returns len(jobs) * ncomp points."""
        n_points, n_scalars = len(points) // (POINT_BYTES * ncomp), len(scalars) // SCALAR_BYTES
        flat = (ctypes.c_uint32 * (3 * len(jobs)))(*[v for j in jobs for v in j])
        out = ctypes.create_string_buffer(max(1, POINT_BYTES * ncomp * len(jobs)))
        _check(self.h, lib.mp377_msm_jobs(self.h, points, n_points, ncomp, scalars, n_scalars, flat, len(jobs), window_bits, out))
        return out.raw[:POINT_BYTES * ncomp * len(jobs)]

    def msm_g1_windows_device(self, d_bases, d_scalars, n, d_out, window_bits, w_begin, w_count):
        _check(self.h, lib.mp377_msm_g1_windows_device(self.h, d_bases, d_scalars, n, window_bits, w_begin, w_count, d_out))

    @staticmethod
    def msm_num_windows(window_bits):
        return lib.mp377_msm_num_windows(window_bits)

    # --- protocol
    def verify_shuffle(self, m, n, enc_g, ck_g, ck_h, ghat, pk, deck, shuffled_deck, proof) -> int:
        """-> 0 or an MP_VERIFY_* code (lib.mp_verify_status_string gives the reference's message)."""
        N = m * n
        assert len(ck_g) == n * POINT_BYTES and len(deck) == len(shuffled_deck) == 2 * N * POINT_BYTES
        assert len(proof) == lib.mp377_proof_len(m, n)
        return _check(self.h, lib.mp377_shuffle_verify(self.h, m, n, enc_g, ck_g, ck_h, ghat, pk, deck, shuffled_deck, proof))

    # --- BarnettSmartProtocol::{setup, shuffle_and_remask} over this curve
    def set_params(self, m, n, enc_g, ck_g, ck_h, ghat):
        assert len(ck_g) == n * POINT_BYTES
        _check(self.h, lib.mp377_ctx_set_params(self.h, m, n, enc_g, ck_g, ck_h, ghat))
        self.m, self.n = m, n

    def remask(self, pk, deck, perm, rho) -> bytes:
        n = len(perm)
        arr = (ctypes.c_uint32 * n)(*perm)
        out = ctypes.create_string_buffer(2 * POINT_BYTES * n)
        _check(self.h, lib.mp377_remask_batch(self.h, pk, deck, arr, rho, n, out))
        return out.raw

    def shuffle_and_remask(self, pk, deck, perm, rho, rand):
        """-> (shuffled deck bytes, proof bytes)"""
        N = self.m * self.n
        assert len(perm) == N and len(deck) == 2 * POINT_BYTES * N and len(rho) == 32 * N
        assert len(rand) == 32 * lib.mp377_prover_randomness_len(self.m, self.n)
        arr = (ctypes.c_uint32 * N)(*perm)
        deck2 = ctypes.create_string_buffer(2 * POINT_BYTES * N)
        proof = ctypes.create_string_buffer(lib.mp377_proof_len(self.m, self.n))
        _check(self.h, lib.mp377_shuffle_and_remask(self.h, pk, deck, arr, rho, rand, deck2, proof))
        return deck2.raw, proof.raw

    def shuffle_and_remask_batch(self, pk, decks, perms, rhos, rands, host_threads=0):
        N = self.m * self.n
        B = len(perms) // N
        arr = (ctypes.c_uint32 * (N * B))(*perms)
        out = ctypes.create_string_buffer(2 * POINT_BYTES * N * B)
        proofs = ctypes.create_string_buffer(lib.mp377_proof_len(self.m, self.n) * B)
        _check(self.h, lib.mp377_shuffle_and_remask_batch(self.h, pk, decks, arr, rhos, rands, B, out, proofs, host_threads))
        return out.raw, proofs.raw

    def subgroup_check(self, points: bytes):
        """-> (return code, per-point statuses: 0 in G1, 1 not a canonical curve point, 2 outside the subgroup)"""
        n = len(points) // POINT_BYTES
        st = (_i32 * max(n, 1))()
        rc = lib.mp377_subgroup_check(self.h, points, n, st)
        return rc, list(st)[:n]

    # --- wire format, deserialising half (square roots + G1 membership on the GPU)
    def points_decompress(self, data: bytes, want_statuses=False):
        """n*48 compressed bytes -> n*96 bytes; raises MpError if any item is rejected unless want_statuses, in which
        case -> (points, statuses: 0 ok, 1 malformed, 2 not on the curve, 3 outside G1, return code), as `_lib.Context`"""
        n = len(data) // FQ_BYTES
        assert len(data) == FQ_BYTES * n
        out = ctypes.create_string_buffer(POINT_BYTES * max(n, 1))
        st = (_i32 * max(n, 1))()
        rc = lib.mp377_points_decompress(self.h, data, n, out, st)
        if want_statuses:
            return out.raw[:POINT_BYTES * n], list(st)[:n], rc
        _check(self.h, rc)
        return out.raw[:POINT_BYTES * n]

    def deck_deserialize(self, data: bytes) -> bytes:
        n = ctypes.c_uint64((len(data) - 8) // (2 * FQ_BYTES) if len(data) >= 8 else 0)
        out = ctypes.create_string_buffer(2 * POINT_BYTES * max(n.value, 1))
        _check(self.h, lib.mp377_deck_deserialize(self.h, data, len(data), out, ctypes.byref(n)))
        return out.raw[:2 * POINT_BYTES * n.value]

    def proof_deserialize(self, m: int, n: int, data: bytes) -> bytes:
        assert len(data) == lib.mp377_proof_serialized_len(m, n)
        out = ctypes.create_string_buffer(lib.mp377_proof_len(m, n))
        _check(self.h, lib.mp377_proof_deserialize(self.h, m, n, data, out))
        return out.raw

    # --- batched sigma protocols either side of the shuffle (reference mod.rs:132-354), 96-byte points:
    #     Chaum-Pedersen proof = a | b | r = 224 bytes, Schnorr proof = commit | opening = 128 bytes
    def _statuses(self, fn, n, *args, host_threads=0):
        st = (_i32 * max(n, 1))()
        _check(self.h, fn(self.h, *args, n, st, host_threads))
        return list(st)[:n]

    def mask_batch(self, shared_key, cards, r, omega, host_threads=0):
        """-> (masked cards n*192, Chaum-Pedersen proofs n*224)"""
        n = len(r) // 32
        assert len(cards) == POINT_BYTES * n and len(omega) == 32 * n
        out, proofs = ctypes.create_string_buffer(2 * POINT_BYTES * n), ctypes.create_string_buffer(CP_PROOF_BYTES * n)
        _check(self.h, lib.mp377_mask_batch(self.h, shared_key, cards, r, omega, n, out, proofs, host_threads))
        return out.raw, proofs.raw

    def verify_mask_batch(self, shared_key, cards, masked, proofs, host_threads=0):
        n = len(proofs) // CP_PROOF_BYTES
        assert len(cards) == POINT_BYTES * n and len(masked) == 2 * POINT_BYTES * n
        return self._statuses(lib.mp377_verify_mask_batch, n, shared_key, cards, masked, proofs, host_threads=host_threads)

    def remask_prove_batch(self, shared_key, deck, alpha, omega, host_threads=0):
        n = len(alpha) // 32
        assert len(deck) == 2 * POINT_BYTES * n and len(omega) == 32 * n
        out, proofs = ctypes.create_string_buffer(2 * POINT_BYTES * n), ctypes.create_string_buffer(CP_PROOF_BYTES * n)
        _check(self.h, lib.mp377_remask_prove_batch(self.h, shared_key, deck, alpha, omega, n, out, proofs, host_threads))
        return out.raw, proofs.raw

    def verify_remask_batch(self, shared_key, deck, remasked, proofs, host_threads=0):
        n = len(proofs) // CP_PROOF_BYTES
        assert len(deck) == len(remasked) == 2 * POINT_BYTES * n
        return self._statuses(lib.mp377_verify_remask_batch, n, shared_key, deck, remasked, proofs, host_threads=host_threads)

    def reveal_batch(self, sk, pk, masked, omega, host_threads=0):
        """-> (reveal tokens n*96, Chaum-Pedersen proofs n*224) of one player for n masked cards"""
        n = len(omega) // 32
        assert len(masked) == 2 * POINT_BYTES * n and len(sk) == 32 and len(pk) == POINT_BYTES
        tokens, proofs = ctypes.create_string_buffer(POINT_BYTES * n), ctypes.create_string_buffer(CP_PROOF_BYTES * n)
        _check(self.h, lib.mp377_reveal_batch(self.h, sk, pk, masked, omega, n, tokens, proofs, host_threads))
        return tokens.raw, proofs.raw

    def verify_reveal_batch(self, pk, tokens, masked, proofs, host_threads=0):
        n = len(proofs) // CP_PROOF_BYTES
        assert len(tokens) == POINT_BYTES * n and len(masked) == 2 * POINT_BYTES * n
        return self._statuses(lib.mp377_verify_reveal_batch, n, pk, tokens, masked, proofs, host_threads=host_threads)

    @staticmethod
    def _infos(infos):
        off = [0]
        for b in infos:
            off.append(off[-1] + len(b))
        return b"".join(infos), (_u64 * len(off))(*off)

    def key_ownership_prove_batch(self, pks, sks, infos, omega, host_threads=0):
        n = len(infos)
        assert len(pks) == POINT_BYTES * n and len(sks) == 32 * n and len(omega) == 32 * n
        blob, off = self._infos(infos)
        proofs = ctypes.create_string_buffer(SCHNORR_PROOF_BYTES * n)
        _check(self.h, lib.mp377_key_ownership_prove_batch(self.h, pks, sks, blob, off, omega, n, proofs, host_threads))
        return proofs.raw

    def key_ownership_verify_batch(self, pks, infos, proofs, host_threads=0):
        n = len(infos)
        assert len(pks) == POINT_BYTES * n and len(proofs) == SCHNORR_PROOF_BYTES * n
        blob, off = self._infos(infos)
        return self._statuses(lib.mp377_key_ownership_verify_batch, n, pks, blob, off, proofs, host_threads=host_threads)

    # --- Pedersen
    def set_commit_key(self, ck: bytes):
        assert len(ck) % POINT_BYTES == 0 and len(ck) >= 2 * POINT_BYTES
        _check(self.h, lib.mp377_set_commit_key(self.h, ck, len(ck) // POINT_BYTES - 1))

    def pedersen_commit_batch(self, values: bytes, blinds: bytes, length: int) -> bytes:
        k = len(blinds) // SCALAR_BYTES
        assert len(values) == k * length * SCALAR_BYTES
        out = ctypes.create_string_buffer(max(1, POINT_BYTES * k))
        _check(self.h, lib.mp377_pedersen_commit_batch(self.h, values, blinds, k, length, out))
        return out.raw[:POINT_BYTES * k]

    # --- measurement / debug hooks
    def profile_enable(self, on=True):
        _check(self.h, lib.mp377_profile_enable(self.h, 1 if on else 0))

    def profile_collect(self):
        ms, adds, n = ctypes.c_double(), _u64(), _u64()
        _check(self.h, lib.mp377_profile_collect(self.h, ctypes.byref(ms), ctypes.byref(adds), ctypes.byref(n)))
        return ms.value, adds.value, n.value

    def dbg_fq_mul(self, a: bytes, b: bytes) -> bytes:
        n = len(a) // FQ_BYTES
        out = ctypes.create_string_buffer(FQ_BYTES * n)
        _check(self.h, lib.mp377_dbg_fq_mul(self.h, a, b, n, out))
        return out.raw

    def dbg_point_add(self, p: bytes, q: bytes) -> bytes:
        n = len(p) // POINT_BYTES
        out = ctypes.create_string_buffer(POINT_BYTES * n)
        _check(self.h, lib.mp377_dbg_point_add(self.h, p, q, n, out))
        return out.raw

    def dbg_scalar_mul(self, p: bytes, k: bytes) -> bytes:
        n = len(p) // POINT_BYTES
        out = ctypes.create_string_buffer(POINT_BYTES * n)
        _check(self.h, lib.mp377_dbg_scalar_mul(self.h, p, k, n, out))
        return out.raw

    def dbg_bench(self, which, iters):
        ms, ops = ctypes.c_float(), ctypes.c_double()
        _check(self.h, lib.mp377_dbg_bench(self.h, which, iters, ctypes.byref(ms), ctypes.byref(ops)))
        return ms.value, ops.value

/* mpshuffle.h -- C ABI of the B200 shuffle-proof engine (libmpshuffle.so).
 *
 * Drop-in boundary for the Bayer-Groth shuffle hot path of geometryxyz/mental-poker:
 *   BarnettSmartProtocol::shuffle_and_remask   reference src/lib.rs:181-188,
 *                                              impl src/discrete_log_cards/mod.rs:380-418
 *   BarnettSmartProtocol::verify_shuffle       reference src/lib.rs:191-197,
 *                                              impl src/discrete_log_cards/mod.rs:420-443
 * and the group arithmetic underneath (ShuffleArgument::{prove,verify},
 * MultiExponentiationArgument, PedersenCommitment::commit, ElGamal remask), which the
 * reference reaches through the un-vendored `proof-essentials` / arkworks crates.
 * The Rust-side binding a maintainer would add is shown in INTEGRATION.md and
 * bindings/rust/.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, caller allocates every buffer, the library never retains a
 *     caller pointer past return;
 *   - field elements: 32 bytes little-endian CANONICAL (non-Montgomery) integers, i.e.
 *     the byte string ark-ff 0.3 `ToBytes` writes;
 *   - affine G1 point: 64 bytes x || y; the identity is the all-zero 64 bytes
 *     ((0,0) is not on the Stark curve); ciphertext = c1 || c2 = 128 bytes;
 *   - return value int32: 0 = ok, > 0 = a verification check failed (MP_VERIFY_*),
 *     < 0 = usage or CUDA error (MP_ERR_*); no exceptions cross the boundary;
 *   - a context is used by one host thread at a time; all work is issued on the
 *     context's CUDA stream.  *_device variants take DEVICE pointers and are asynchronous
 *     on that stream (call mp_ctx_sync before reading results on the host).
 *   - there is NO CPU fallback: if no CUDA device is usable mp_ctx_create fails.
 */
#ifndef MPSHUFFLE_H
#define MPSHUFFLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mp_ctx mp_ctx;

#define MP_OK 0
/* verification failures: the strings are those of the reference's CryptoError
 * (MP_VERIFY_HADAMARD must map to exactly "Hadamard Product (5.1)",
 * reference src/discrete_log_cards/tests.rs:223-225) */
#define MP_VERIFY_HADAMARD 1
#define MP_VERIFY_ZERO 2
#define MP_VERIFY_SVP 3
#define MP_VERIFY_MULTIEXP 4
/* usage / runtime errors */
#define MP_ERR_INVALID_ARG (-1)
#define MP_ERR_CUDA (-2)
#define MP_ERR_NOT_ON_CURVE (-3)
#define MP_ERR_NO_PARAMS (-4)

/* ---- context ------------------------------------------------------------------------ */
/* Creates a context bound to CUDA device `device` (owns a stream and device scratch). */
int32_t mp_ctx_create(mp_ctx** out, int32_t device);
void mp_ctx_destroy(mp_ctx* ctx);
/* cudaStream_t of the context (as void*), for timing with CUDA events on the launch stream. */
void* mp_ctx_stream(mp_ctx* ctx);
int32_t mp_ctx_sync(mp_ctx* ctx);
/* Human-readable description of the last error on this context (never NULL). */
const char* mp_last_error_string(mp_ctx* ctx);
/* Message string of a positive verification status (e.g. "Hadamard Product (5.1)"). */
const char* mp_verify_status_string(int32_t status);
/* Kernels launched by the most recent entry-point call on this context. */
int32_t mp_last_kernel_launches(mp_ctx* ctx);

/* ---- variable-base MSM (replaces ark-ec scalar-mul loops / VariableBaseMSM) ---------- */
/* out = sum_i scalars[i] * bases[i].  window_bits = 0 picks the window automatically. */
int32_t mp_msm_g1(mp_ctx* ctx, const uint8_t* bases /* n*64 */, const uint8_t* scalars /* n*32 */,
                  uint64_t n, int32_t window_bits, uint8_t* out /* 64 */);
/* Ciphertext MSM: out = sum_i scalars[i] * deck[i] component-wise (2 G1 MSMs sharing digits). */
int32_t mp_ct_msm(mp_ctx* ctx, const uint8_t* deck /* n*128 */, const uint8_t* scalars /* n*32 */,
                  uint64_t n, int32_t window_bits, uint8_t* out /* 128 */);
/* Same with device-resident inputs/outputs (canonical byte layout), asynchronous. */
int32_t mp_msm_g1_device(mp_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                         int32_t window_bits, void* d_out);
int32_t mp_ct_msm_device(mp_ctx* ctx, const void* d_deck, const void* d_scalars, uint64_t n,
                         int32_t window_bits, void* d_out);
/* EC additions scheduled by the last MSM on this context (bucket adds + reduction adds),
 * and the window width it used. */
uint64_t mp_last_msm_ec_adds(mp_ctx* ctx);
int32_t mp_last_msm_window(mp_ctx* ctx);

/* ---- debug / parity hooks (exercise single device primitives; not used by the protocol) */
int32_t mp_dbg_fq_mul(mp_ctx* ctx, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out);
int32_t mp_dbg_point_add(mp_ctx* ctx, const uint8_t* p, const uint8_t* q, uint64_t n, uint8_t* out);
int32_t mp_dbg_scalar_mul(mp_ctx* ctx, const uint8_t* p, const uint8_t* k, uint64_t n, uint8_t* out);
/* integer-pipe microbenchmarks: returns milliseconds for `iters` dependent-chain iterations
 * on a full-chip grid; *ops receives the number of counted operations executed. */
int32_t mp_dbg_bench(mp_ctx* ctx, int32_t which, int32_t iters, float* ms, double* ops);

#ifdef __cplusplus
}
#endif
#endif /* MPSHUFFLE_H */

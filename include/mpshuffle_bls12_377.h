/* mpshuffle_bls12_377.h -- C ABI of the BLS12-377 G1 instantiation inside libmpshuffle.so.
 *
 * The reference's card protocol is generic over the curve, `DLCards<C: ProjectiveCurve>`
 * (reference src/discrete_log_cards/mod.rs:86), and its benchmark harness instantiates it over
 * `ark_bls12_377::G1Projective` / `ark_bls12_377::Fr` (examples/parameter_selection.rs:25-26).
 * These entry points are the group layer of that instantiation: the
 * variable-base MSM, the 2-component ciphertext MSM and the fixed-base batched Pedersen commitment
 * that ShuffleArgument::{prove,verify} / MultiExponentiationArgument / PedersenCommitment::commit
 * reduce to (call sites mod.rs:397-415,427-442; commit key setup mod.rs:111), plus verify_shuffle built on
 * them with host-side scalars (mp377_shuffle_verify), and shuffle_and_remask (mp377_shuffle_and_remask[_batch]: the
 * Stark build's lockstep prover compiled against this field), the batched sigma protocols either side of the shuffle
 * (mp377_mask_batch ... mp377_key_ownership_verify_batch) and the wire format in both directions (mp377_points_compress /
 * _decompress, deck and proof (de)serialisers).  Only the device-scalar large-deck prover (decks above 8 192 cards) and
 * the resident-buffer entry points are Stark-curve only.
 *
 * Conventions are those of mpshuffle.h with the sizes of this curve:
 *   - base-field element: 48 bytes little-endian canonical (ark-ff 0.3 `Fp384` `ToBytes`);
 *   - scalar: 32 bytes little-endian canonical, < r (253 bits);
 *   - affine G1 point: 96 bytes x || y, identity = 96 zero bytes ((0,0) is not on y^2 = x^3 + 1);
 *     ciphertext = c1 || c2 = 192 bytes.  Points are checked to be canonical and on the curve by every entry
 *     point; the group-layer entry points (MSM, commitments) do NOT test subgroup membership (G1 has cofactor
 *     0x170b5d44300000000000000000000000) -- mp377_subgroup_check does, and mp377_shuffle_verify applies it to
 *     every untrusted point, as ark-ec's CanonicalDeserialize does on the reference's side;
 *   - status codes MP_OK / MP_ERR_* of mpshuffle.h; no CPU fallback.
 */
#ifndef MPSHUFFLE_BLS12_377_H
#define MPSHUFFLE_BLS12_377_H

#include "mpshuffle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mp377_ctx mp377_ctx;

#define MP377_FQ_BYTES 48
#define MP377_POINT_BYTES 96
#define MP377_SCALAR_BYTES 32

int32_t mp377_ctx_create(mp377_ctx** out, int32_t device);
void mp377_ctx_destroy(mp377_ctx* ctx);
void* mp377_ctx_stream(mp377_ctx* ctx);
int32_t mp377_ctx_sync(mp377_ctx* ctx);
const char* mp377_last_error_string(mp377_ctx* ctx);
int32_t mp377_last_kernel_launches(mp377_ctx* ctx);
/* EC additions the last MSM call scheduled / the window width it used / windows for a width */
uint64_t mp377_last_msm_ec_adds(mp377_ctx* ctx);
int32_t mp377_last_msm_window(mp377_ctx* ctx);
int32_t mp377_msm_num_windows(int32_t window_bits);

/* out = sum_i scalars[i] * bases[i]  (replaces ark-ec 0.3 VariableBaseMSM / scalar-mul loops) */
int32_t mp377_msm_g1(mp377_ctx* ctx, const uint8_t* bases /* n*96 */, const uint8_t* scalars /* n*32 */,
                     uint64_t n, int32_t window_bits, uint8_t* out /* 96 */);
/* ciphertext MSM: component-wise over (c1, c2) pairs, both components share digits and sort */
int32_t mp377_ct_msm(mp377_ctx* ctx, const uint8_t* deck /* n*192 */, const uint8_t* scalars /* n*32 */,
                     uint64_t n, int32_t window_bits, uint8_t* out /* 192 */);
/* device-resident inputs / outputs (same byte layout), asynchronous on the context's stream */
int32_t mp377_msm_g1_device(mp377_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                            int32_t window_bits, void* d_out);
int32_t mp377_ct_msm_device(mp377_ctx* ctx, const void* d_deck, const void* d_scalars, uint64_t n,
                            int32_t window_bits, void* d_out);

/* Batch of MSMs over one point array and one scalar array (MultiExponentiationArgument's diagonal products:
 * m(m+1) inner products <row of n ciphertexts, row of n scalars>, one launch sequence for all of them):
 *   out[j] = sum_{t < len_j} scalars[scalar_off_j + t] * points[point_off_j + t]       (per component)
 * jobs = njobs x (scalar_off, point_off, len) as uint32; ncomp = 1 (96-byte points) or 2 (192-byte ciphertexts);
 * out = njobs * ncomp * 96 bytes. */
int32_t mp377_msm_jobs(mp377_ctx* ctx, const uint8_t* points, uint64_t n_points, int32_t ncomp,
                       const uint8_t* scalars, uint64_t n_scalars, const uint32_t* jobs, uint64_t njobs,
                       int32_t window_bits, uint8_t* out);

/* BarnettSmartProtocol::{setup, shuffle_and_remask} over this curve (reference src/lib.rs:74-78,181-188; impl
 * mod.rs:105-121,380-418; the instantiation and the call the reference's benchmark harness times,
 * examples/parameter_selection.rs:25-29,78-96).  Same contract as mp_ctx_set_params / mp_remask_batch /
 * mp_shuffle_and_remask / mp_shuffle_and_remask_batch of mpshuffle.h with 96-byte points and 192-byte ciphertexts:
 * proof_out receives mp377_proof_len(m, n) bytes, randomness = mp377_prover_randomness_len(m, n) = 11m + 5n scalars
 * drawn by the caller in the order of SURVEY.md Appendix B.6; proofs are byte-identical to the oracle's.  Decks up
 * to 8 192 cards (the lockstep, host-scalar prover; the device-scalar large-deck prover is Stark-only). */
int32_t mp377_ctx_set_params(mp377_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g /* 96 */, const uint8_t* ck_g /* n*96 */,
                             const uint8_t* ck_h /* 96 */, const uint8_t* ghat /* 96 */);
uint64_t mp377_prover_randomness_len(int32_t m, int32_t n);
int32_t mp377_remask_batch(mp377_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm, const uint8_t* rho,
                           uint64_t n_cards, uint8_t* out_deck);
int32_t mp377_shuffle_and_remask(mp377_ctx* ctx, const uint8_t* pk /* 96 */, const uint8_t* deck /* m*n*192 */, const uint32_t* perm,
                                 const uint8_t* rho /* m*n*32 */, const uint8_t* randomness, uint8_t* out_deck, uint8_t* proof_out);
int32_t mp377_shuffle_and_remask_batch(mp377_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                                       const uint8_t* rhos, const uint8_t* randomness, uint64_t batch, uint8_t* out_decks,
                                       uint8_t* proofs, int32_t host_threads);

/* Batched sigma protocols either side of the shuffle over this curve: BarnettSmartProtocol::{mask, verify_mask, remask,
 * verify_remask, compute_reveal_token, verify_reveal, prove_key_ownership, verify_key_ownership} (reference
 * src/lib.rs:88-175, impl mod.rs:132-354), n independent items per call.  Same contract as mp_mask_batch ...
 * mp_key_ownership_verify_batch of mpshuffle.h with 96-byte points: Chaum-Pedersen proof = a (96) | b (96) | r (32) =
 * 224 bytes, Schnorr proof = commit (96) | opening (32) = 128 bytes; the generator is enc_g of mp377_ctx_set_params
 * (which must have been called).  statuses[i] = MP_OK, MP_VERIFY_CHAUM_PEDERSEN / MP_VERIFY_SCHNORR, or
 * MP_VERIFY_MALFORMED for an item one of whose points (card, ciphertext, token, proof commitment) is not a canonical
 * point of G1 -- every point a verifier is handed goes through the membership test first, as the reference's
 * deserialiser would do, and a refused item does not stop the others; a bad KEY fails the call with
 * MP_ERR_NOT_ON_CURVE / MP_ERR_NOT_IN_SUBGROUP.  The provers take their inputs as trusted (on-curve is still checked).  Bytes are identical to oracle/py/sigma.py
 * under curve("bls12_377"). */
int32_t mp377_mask_batch(mp377_ctx* ctx, const uint8_t* shared_key /* 96 */, const uint8_t* cards /* n*96 */,
                         const uint8_t* r /* n*32 */, const uint8_t* omega /* n*32 */, uint64_t n,
                         uint8_t* out_masked /* n*192 */, uint8_t* out_proofs /* n*224 */, int32_t host_threads);
int32_t mp377_verify_mask_batch(mp377_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* masked,
                                const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t mp377_remask_prove_batch(mp377_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck /* n*192 */,
                                 const uint8_t* alpha /* n*32 */, const uint8_t* omega /* n*32 */, uint64_t n,
                                 uint8_t* out_deck /* n*192 */, uint8_t* out_proofs /* n*224 */, int32_t host_threads);
int32_t mp377_verify_remask_batch(mp377_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* remasked,
                                  const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t mp377_reveal_batch(mp377_ctx* ctx, const uint8_t* sk /* 32 */, const uint8_t* pk /* 96 */,
                           const uint8_t* masked /* n*192 */, const uint8_t* omega /* n*32 */, uint64_t n,
                           uint8_t* out_tokens /* n*96 */, uint8_t* out_proofs /* n*224 */, int32_t host_threads);
int32_t mp377_verify_reveal_batch(mp377_ctx* ctx, const uint8_t* pk, const uint8_t* tokens, const uint8_t* masked,
                                  const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t mp377_key_ownership_prove_batch(mp377_ctx* ctx, const uint8_t* pks /* n*96 */, const uint8_t* sks /* n*32 */,
                                        const uint8_t* infos, const uint64_t* info_offsets /* n+1 */,
                                        const uint8_t* omega /* n*32 */, uint64_t n, uint8_t* out_proofs /* n*128 */,
                                        int32_t host_threads);
int32_t mp377_key_ownership_verify_batch(mp377_ctx* ctx, const uint8_t* pks, const uint8_t* infos,
                                         const uint64_t* info_offsets, const uint8_t* proofs, uint64_t n,
                                         int32_t* statuses, int32_t host_threads);

/* Subgroup membership of n canonical points (n * 96 bytes): MP_OK iff every point is a canonical point of the curve
 * and lies in the order-r subgroup G1; otherwise MP_ERR_NOT_ON_CURVE / MP_ERR_NOT_IN_SUBGROUP.  statuses (optional,
 * n entries): 0 = in G1, 1 = not a canonical curve point, 2 = on the curve but outside G1.  One 127-bit
 * double-and-add per point (endomorphism test phi(P) == -[u^2]P, as ark-bls12-377 does). */
int32_t mp377_subgroup_check(mp377_ctx* ctx, const uint8_t* points, uint64_t n, int32_t* statuses);

/* BarnettSmartProtocol::verify_shuffle over this curve (reference src/lib.rs:191-197, impl mod.rs:420-443;
 * Parameters = (m, n, enc generator, commit key g_1..g_n / h, extra generator), mod.rs:37-61).  Returns
 * MP_OK, an MP_VERIFY_* code (mp_verify_status_string gives the reference's message, e.g.
 * "Hadamard Product (5.1)"), or MP_ERR_*.  Proof layout: the flat layout of mpshuffle.h with 96-byte
 * points, mp377_proof_len(m, n) = (11m + 8) * 96 + (5n + 9) * 32 bytes.  Untrusted inputs are validated as the
 * reference's deserialiser would: every point canonical, on the curve and in G1 (MP_ERR_NOT_ON_CURVE /
 * MP_ERR_NOT_IN_SUBGROUP), every proof scalar below r (MP_ERR_NOT_CANONICAL).  Host-scalar path: the O(N) scalar
 * work and the transcript run on the calling thread, the group work on the GPU. */
uint64_t mp377_proof_len(int32_t m, int32_t n);
int32_t mp377_shuffle_verify(mp377_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g /* 96 */,
                             const uint8_t* ck_g /* n*96 */, const uint8_t* ck_h /* 96 */, const uint8_t* ghat /* 96 */,
                             const uint8_t* pk /* 96 */, const uint8_t* deck /* m*n*192 */,
                             const uint8_t* shuffled_deck /* m*n*192 */, const uint8_t* proof);

/* Wire format (ark-serialize 0.3 compressed encodings, as the wire-format section of mpshuffle.h): compressed point =
 * x (48 bytes LE) with flags in the top bits of the last byte (bit 7: y is the larger of (y, -y); bit 6: infinity);
 * Vec<MaskedCard> = u64 LE length | c1 | c2 per card; proof = the REPOSITORY-PRIVATE container of mpshuffle.h (flat
 * layout, every point compressed, no Vec length prefixes), (11m + 8) * 48 + (5n + 9) * 32 bytes -- close to, but not,
 * what the reference's benchmark prints with `serialized_size()` (examples/parameter_selection.rs:93-96), which adds
 * 8 bytes per Vec field of the upstream proof struct.  Serialising is host byte handling and needs no context.
 * DEserialising runs on the GPU: one square root in F_q per point (q - 1 = 2^46 * t; windowed Tonelli-Shanks), then
 * the G1 membership test, as ark-serialize's CanonicalDeserialize does for this curve.  mp377_points_decompress
 * fills statuses[i] (may be NULL) with 0 ok, 1 malformed encoding (x >= q, stray flag bits), 2 x not on the curve,
 * 3 on the curve but outside G1, writes rejected items as 96 zero bytes and returns MP_ERR_NOT_ON_CURVE /
 * MP_ERR_NOT_IN_SUBGROUP; mp377_proof_deserialize also rejects scalars >= r (MP_ERR_NOT_CANONICAL). */
int32_t mp377_points_compress(const uint8_t* points /* n*96 */, uint64_t n, uint8_t* out /* n*48 */);
uint64_t mp377_deck_serialized_len(uint64_t n_cards);
int32_t mp377_deck_serialize(const uint8_t* deck /* n_cards*192 */, uint64_t n_cards, uint8_t* out);
uint64_t mp377_proof_serialized_len(int32_t m, int32_t n);
int32_t mp377_proof_serialize(int32_t m, int32_t n, const uint8_t* proof, uint8_t* out);
int32_t mp377_points_decompress(mp377_ctx* ctx, const uint8_t* in /* n*48 */, uint64_t n, uint8_t* out /* n*96 */,
                                int32_t* statuses /* n or NULL */);
/* *n_cards: in = capacity of out_deck in cards, out = cards in the buffer */
int32_t mp377_deck_deserialize(mp377_ctx* ctx, const uint8_t* in, uint64_t in_len, uint8_t* out_deck, uint64_t* n_cards);
int32_t mp377_proof_deserialize(mp377_ctx* ctx, int32_t m, int32_t n, const uint8_t* in, uint8_t* out_proof);

/* Window-range split of ONE MSM across GPUs: rank r computes the windows
 * [w_begin, w_begin + w_count) of the mp377_msm_num_windows(window_bits) windows end to end and returns
 *   P_r = sum_w 2^(c (w - w_begin)) * (window sum w);   MSM = sum_r 2^(c * w_begin_r) * P_r
 * after an all-gather of the 96-byte partials (mental-poker_b200/dist.py).  window_bits must be explicit. */
int32_t mp377_msm_g1_windows_device(mp377_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                                    int32_t window_bits, int32_t w_begin, int32_t w_count, void* d_out);

/* Pedersen commitments over a constant key (PedersenCommitment::setup / commit):
 * ck = h || G_1 .. G_len ((len+1)*96 bytes); builds the fixed-base window table once. */
int32_t mp377_set_commit_key(mp377_ctx* ctx, const uint8_t* ck, uint64_t len);
/* out[j] = blinds[j]*h + sum_i values[j*len + i]*G_{i+1},  j < k,  len <= key length */
int32_t mp377_pedersen_commit_batch(mp377_ctx* ctx, const uint8_t* values /* k*len*32 */,
                                    const uint8_t* blinds /* k*32 */, uint64_t k, uint64_t len,
                                    uint8_t* out /* k*96 */);

/* CUDA-event timing of the bucket-accumulation kernel on the launch stream (bench.py) */
int32_t mp377_profile_enable(mp377_ctx* ctx, int32_t on);
int32_t mp377_profile_collect(mp377_ctx* ctx, double* accumulate_ms, uint64_t* bucket_adds, uint64_t* launches);

/* parity helpers: one device field / group operation per element (tests only) */
int32_t mp377_dbg_fq_mul(mp377_ctx* ctx, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out);
int32_t mp377_dbg_point_add(mp377_ctx* ctx, const uint8_t* p, const uint8_t* q, uint64_t n, uint8_t* out);
int32_t mp377_dbg_scalar_mul(mp377_ctx* ctx, const uint8_t* p, const uint8_t* k, uint64_t n, uint8_t* out);
/* microbenchmarks: which = 0 fq_mul, 1 XYZZ mixed addition; returns ms and the operation count */
int32_t mp377_dbg_bench(mp377_ctx* ctx, int32_t which, int32_t iters, float* ms, double* ops);

#ifdef __cplusplus
}
#endif
#endif

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
/* sigma protocols either side of the shuffle: "Chaum-Pedersen" (reference masking.rs:103-105,
 * remasking.rs:110-112, reveal.rs:80-82) and "Schnorr Identification" (tests.rs:72-77) */
#define MP_VERIFY_CHAUM_PEDERSEN 5
#define MP_VERIFY_SCHNORR 6
/* batch verifiers only: item i is malformed -- a point off the curve / not canonical, or a scalar >= the group
 * order -- the per-item form of MP_ERR_NOT_ON_CURVE / MP_ERR_NOT_CANONICAL, so that one bad item does not fail the
 * other items of the batch (the reference fails only the offending item, at deserialisation) */
#define MP_VERIFY_MALFORMED 7
/* usage / runtime errors */
#define MP_ERR_INVALID_ARG (-1)
#define MP_ERR_CUDA (-2)
#define MP_ERR_NOT_ON_CURVE (-3)
#define MP_ERR_NO_PARAMS (-4)
/* a scalar of an untrusted input (a proof) is not a canonical residue, i.e. >= the group order: ark-serialize's
 * CanonicalDeserialize rejects such bytes before the reference's verifier ever sees them (s and s + order would
 * otherwise both verify -- proof malleability) */
#define MP_ERR_NOT_CANONICAL (-5)
/* BLS12-377 only (the Stark curve has cofactor 1): a point is on the curve but outside the order-r subgroup G1 */
#define MP_ERR_NOT_IN_SUBGROUP (-6)
/* NCCL is not available (libnccl.so.2 could not be loaded) or a collective failed */
#define MP_ERR_NCCL (-7)

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
/* Batch of MSMs over one point array and one scalar array -- SURVEY.md section 8(b)'s
 * `mp_msm_batch_shared_bases`, the surface of `MultiExponentiationArgument` named in BASELINE.json: its diagonal
 * products are m(m+1) inner products <row of n ciphertexts, row of n scalars> (reference call site
 * mod.rs:409-415), evaluated here by ONE launch sequence:
 *   out[j] = sum_{t < len_j} scalars[scalar_off_j + t] * points[point_off_j + t]       (per component)
 * jobs = njobs x (scalar_off, point_off, len) as uint32; ncomp = 1 (64-byte points) or 2 (128-byte ciphertexts);
 * out = njobs * ncomp * 64 bytes.  Jobs may overlap and share points or scalars. */
int32_t mp_msm_jobs(mp_ctx* ctx, const uint8_t* points, uint64_t n_points, int32_t ncomp, const uint8_t* scalars,
                    uint64_t n_scalars, const uint32_t* jobs, uint64_t njobs, int32_t window_bits, uint8_t* out);
/* Same with device-resident inputs/outputs (canonical byte layout), asynchronous. */
int32_t mp_msm_g1_device(mp_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                         int32_t window_bits, void* d_out);
int32_t mp_ct_msm_device(mp_ctx* ctx, const void* d_deck, const void* d_scalars, uint64_t n,
                         int32_t window_bits, void* d_out);
/* Window-range split of one large MSM across GPUs (BASELINE north_star):
 * with W = mp_msm_num_windows(c) signed c-bit windows, this computes only windows
 * [w_begin, w_begin + w_count) and returns  P = sum_w 2^(c*(w - w_begin)) * (window sum w),  so
 *   full MSM = sum over ranks of 2^(c * w_begin_r) * P_r
 * -- the ranks exchange 64-byte partials (all-gather) and fold them with one tiny MSM whose
 * scalars are the powers 2^(c * w_begin_r).  Inputs are replicated on every rank. */
int32_t mp_msm_num_windows(int32_t window_bits);
int32_t mp_msm_g1_windows_device(mp_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                                 int32_t window_bits, int32_t w_begin, int32_t w_count, void* d_out);
/* EC additions scheduled by the last MSM on this context (bucket adds + reduction adds),
 * and the window width it used. */
uint64_t mp_last_msm_ec_adds(mp_ctx* ctx);
int32_t mp_last_msm_window(mp_ctx* ctx);

/* ---- multi-GPU --------------------------------------------------------------
 * One context per GPU -- one process per GPU, or one thread per GPU inside one process -- joined by an NCCL
 * communicator.  Rank 0 calls mp_comm_unique_id and hands the 128 bytes to the other ranks over the host's own
 * channel (MPI, a socket, torch.distributed, a Rust mpsc); every rank then calls mp_comm_init on its context
 * (collective).  NCCL is bound at run time: without libnccl.so.2 these calls return MP_ERR_NCCL and everything
 * else keeps working.  A context without a communicator behaves as rank 0 of 1. */
#define MP_COMM_ID_BYTES 128
int32_t mp_comm_unique_id(uint8_t* id_out /* MP_COMM_ID_BYTES */);
int32_t mp_comm_init(mp_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id /* MP_COMM_ID_BYTES */);
int32_t mp_comm_destroy(mp_ctx* ctx);
int32_t mp_comm_size(mp_ctx* ctx);
int32_t mp_comm_rank(mp_ctx* ctx);
/* ONE variable-base MSM split by window range (BASELINE config 5).  Collective: every rank passes the same
 * device-resident inputs; rank r runs windows shard(W, r, nranks) end to end, the 128-byte XYZZ partials are
 * all-gathered on the context's stream (no host synchronisation) and folded by one small kernel; every rank
 * receives the canonical result in d_out (64 bytes, device).  Asynchronous on the context's stream. */
int32_t mp_msm_g1_multi_device(mp_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                               int32_t window_bits, void* d_out);
/* Batch of independent proofs, proof-index split (BASELINE config 4): every rank verifies ITS `batch_per_rank`
 * proofs (decks / shuffled_decks / proofs = this rank's shard) with mp_shuffle_verify_batch -- no data-path
 * collective -- and the verdicts are all-gathered: statuses_all receives nranks * batch_per_rank entries, rank-major,
 * on every rank. */
int32_t mp_shuffle_verify_batch_multi(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint8_t* shuffled_decks,
                                      const uint8_t* proofs, uint64_t batch_per_rank, int32_t* statuses_all,
                                      int32_t host_threads);
/* ONE large proof across the GPUs of the communicator (config 3).  Collective: every rank makes the same call on
 * the same inputs and receives the same bytes / verdict as the single-GPU entry point would return.  The prover's
 * diagonal ciphertext products (Karatsuba leaf products: independent MSMs, ~70 % of its device time at 2^16 cards)
 * are split by leaf index and their results all-gathered; the verifier's two ciphertext equations run on two
 * ranks.  Decks below the large-deck threshold (8 192 cards) run unsplit on every rank. */
int32_t mp_shuffle_and_remask_multi(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                    const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck,
                                    uint8_t* proof_out);
int32_t mp_shuffle_verify_multi(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* shuffled_deck,
                                const uint8_t* proof);

/* ---- shuffle protocol (the reference's hot path) ------------------------------------------
 * Data layouts.  A deck is N = m*n ElGamal ciphertexts, 128 bytes each (c1 || c2).  Scalars
 * (masking factors, prover randomness) are 32-byte little-endian canonical integers < group
 * order.  A permutation is N uint32, out[i] = in[perm[i]] (proof-essentials
 * `Permutation::permute_array`, reference mod.rs:388).
 *
 * Flat proof layout (mp_proof_len(m, n) = (11m+8)*64 + (5n+9)*32 bytes), P = 64-byte point,
 * F = 32-byte scalar, in the order of SURVEY.md Appendix B:
 *   c_A[m]P  c_B[m]P                                   shuffle argument first messages
 *   c_b P                                              product argument
 *   c_B'[m]P                                           Hadamard argument
 *   c_A0 P  c_B(m+1) P  c_D[2m+1]P  a[n]F b[n]F r F s F t F      zero argument
 *   c_d P  c_delta P  c_Delta P  a~[n]F b~[n]F r~ F s~ F         single-value product argument
 *   c_A0 P  c_B[2m]P  E[2m] (2P each)  a[n]F r F b F s F tau F   multi-exponentiation argument
 *
 * Prover randomness: the library never owns an RNG (the trait hands the caller's `rng` to the
 * prover, reference lib.rs:181-188); the host draws mp_prover_randomness_len(m, n) = 11m + 5n
 * scalars up front and passes them flat, consumed in this order: shuffle r[m], s[m]; product
 * s; Hadamard s_2..s_{m-1}; zero a_0[n], b_{m+1}[n], r_0, s_{m+1}, t_k (k = 0..2m, k != m+1);
 * single-value product d[n], r_d, delta_2..delta_{n-1}, s_1, s_x; multi-exp a_0[n], r_0,
 * (b_k, s_k, tau_k) for k = 0..2m-1, k != m.
 *
 * Fiat-Shamir transcript (host side, ark-marlin FiatShamirRng<Blake2s> seeded with
 * "Shuffle Proof", reference mod.rs:84,408,436): points are absorbed in the 65-byte ark-ec
 * encoding x || y || infinity.  Absorb order: "shuffle_argument" g pk G_1..G_n H ghat deck
 * deck' c_A -> x;  "shuffle_argument_b" c_B -> y, z;  "hadamard_argument" c_b c_B' -> x, y;
 * "zero_argument" c_A0 c_B(m+1) c_D -> x;  "single_value_product_argument" c_d c_delta c_Delta
 * -> x;  "multi_exponentiation_argument" c_A0 c_B E -> x.
 */
/* Binds DLCards `Parameters` (reference mod.rs:37-61, created by `setup`, mod.rs:105-121) to the
 * context: ElGamal generator, Pedersen key G_1..G_n and H, extra generator ghat, and (m, n). */
int32_t mp_ctx_set_params(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g /* 64 */,
                          const uint8_t* ck_g /* n*64 */, const uint8_t* ck_h /* 64 */,
                          const uint8_t* ghat /* 64 */);
int32_t mp_params_m(mp_ctx* ctx);
int32_t mp_params_n(mp_ctx* ctx);
uint64_t mp_proof_len(int32_t m, int32_t n);
uint64_t mp_prover_randomness_len(int32_t m, int32_t n);
/* `MaskedCard::remask` over a permuted deck (reference mod.rs:388-395, remasking.rs:9-22):
 * out[i] = deck[perm[i]] + (rho_i * g, rho_i * pk). */
int32_t mp_remask_batch(mp_ctx* ctx, const uint8_t* pk /* 64 */, const uint8_t* deck /* N*128 */,
                        const uint32_t* perm /* N */, const uint8_t* rho /* N*32 */, uint64_t n_cards,
                        uint8_t* out_deck /* N*128 */);
/* `PedersenCommitment::commit` for k vectors of `len` <= n values: out[i] = blinds[i]*H + sum_j
 * values[i][j]*G_j. */
int32_t mp_pedersen_commit_batch(mp_ctx* ctx, const uint8_t* values /* k*len*32 */,
                                 const uint8_t* blinds /* k*32 */, uint64_t k, uint64_t len,
                                 uint8_t* out /* k*64 */);
/* `ShuffleArgument::prove` (reference call site mod.rs:409-415). */
int32_t mp_shuffle_prove(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* shuffled_deck,
                         const uint32_t* perm, const uint8_t* rho, const uint8_t* randomness,
                         uint8_t* proof_out);
/* `BarnettSmartProtocol::shuffle_and_remask` (reference lib.rs:181-188, mod.rs:380-418). */
int32_t mp_shuffle_and_remask(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                              const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck,
                              uint8_t* proof_out);
/* `BarnettSmartProtocol::verify_shuffle` (reference lib.rs:191-197, mod.rs:420-443): returns
 * MP_OK, a positive MP_VERIFY_* code naming the first failing sub-argument, or a negative error. */
int32_t mp_shuffle_verify(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* shuffled_deck,
                          const uint8_t* proof);

/* `shuffle_and_remask` for `batch` independent decks under the same parameters and public key
 * (per-proof buffers concatenated; randomness is batch * mp_prover_randomness_len scalars).  Every
 * proof is byte-identical to the single-call result.  `host_threads` (0 = all hardware threads) is the
 * number of CPU threads the call may keep BUSY.  Decks above 8 192 cards run on up to 8 worker contexts
 * whatever the budget (a worker sleeps while its kernels run; host_threads == 1 means one worker); when
 * the budget is below half the number of workers -- several GPUs' callers sharing one host -- the serial
 * Blake2s passes over the statements (17 MB per proof at 2^16 cards) are computed for up to eight proofs
 * at once by one thread (multi-stream Blake2s) instead of one pass per worker.  The batch verifier below
 * treats host_threads the same way.  Small decks: the thread count of the per-proof host phases. */
int32_t mp_shuffle_and_remask_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                                    const uint8_t* rhos, const uint8_t* randomness, uint64_t batch,
                                    uint8_t* out_decks, uint8_t* proofs, int32_t host_threads);
/* `verify_shuffle` for `batch` independent proofs under the same parameters and public key
 * (BASELINE config: batch of 52-card proofs).  decks / shuffled_decks / proofs are the per-proof
 * buffers concatenated.  statuses[i] receives 0 or the MP_VERIFY_* code of proof i.  The
 * transcripts run on `host_threads` CPU threads (0 = all hardware threads); every group equation
 * of the whole batch is evaluated by two batched MSM launch sequences.  A point off the curve
 * anywhere in the batch fails the call with MP_ERR_NOT_ON_CURVE. */
int32_t mp_shuffle_verify_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks,
                                const uint8_t* shuffled_decks, const uint8_t* proofs, uint64_t batch,
                                int32_t* statuses, int32_t host_threads);
/* Same as mp_shuffle_verify / mp_shuffle_prove for decks that are ALREADY resident in HBM
 * (d_* = device pointers to the same canonical bytes): no deck crosses PCIe.  The host copies
 * are still required -- the Fiat-Shamir transcript hashes them on the CPU. */
int32_t mp_shuffle_verify_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck,
                                   const uint8_t* shuffled_deck, const uint8_t* proof, const void* d_deck,
                                   const void* d_shuffled_deck);
int32_t mp_shuffle_and_remask_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                       const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck,
                                       uint8_t* proof_out, const void* d_deck);
int32_t mp_shuffle_prove_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck,
                                  const uint8_t* shuffled_deck, const uint32_t* perm, const uint8_t* rho,
                                  const uint8_t* randomness, uint8_t* proof_out, const void* d_shuffled_deck);

/* The two batch entry points with the decks ALREADY resident in HBM (d_decks / d_shuffled_decks = device
 * pointers to batch * N * 128 canonical bytes, same layout as the host buffers).  The device copies are used
 * for decks above the small-deck threshold (8 192 cards), where the batch runs the single-proof path on a few
 * worker contexts so that one deck's serial statement hash (host) overlaps the other decks' kernels (device);
 * smaller decks are staged from the host buffers (a few KB each).  Proofs and verdicts are identical to the
 * host-buffer calls.  bench.py's `value` is measured through these two. */
int32_t mp_shuffle_and_remask_batch_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                                             const uint8_t* rhos, const uint8_t* randomness, uint64_t batch,
                                             uint8_t* out_decks, uint8_t* proofs, int32_t host_threads,
                                             const void* d_decks);
int32_t mp_shuffle_verify_batch_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks,
                                         const uint8_t* shuffled_decks, const uint8_t* proofs, uint64_t batch,
                                         int32_t* statuses, int32_t host_threads, const void* d_decks,
                                         const void* d_shuffled_decks);

/* ---- batched sigma protocols either side of the shuffle -----------
 * n independent items per call; item i uses the i-th entry of every array.  The generator g is the
 * ElGamal generator of mp_ctx_set_params.  Proof bytes: Chaum-Pedersen = a (64) | b (64) | r (32) =
 * 160 bytes, Schnorr = commit (64) | opening (32) = 96 bytes.  `omega` is the prover's randomness (one
 * scalar per proof, drawn by the caller from its own RNG in item order -- the library never owns an
 * RNG).  statuses[i] receives MP_OK or MP_VERIFY_CHAUM_PEDERSEN / MP_VERIFY_SCHNORR.  The per-proof
 * Fiat-Shamir transcripts run on `host_threads` CPU threads (0 = all hardware threads).  Verifiers: an
 * item with a point off the curve or a non-canonical coordinate gets statuses[i] = MP_VERIFY_MALFORMED and
 * the other items are still checked (the reference's deserialiser would refuse that one message); a bad
 * KEY -- shared by the whole call -- fails the call with MP_ERR_NOT_ON_CURVE, as does any bad input of a
 * prover.
 *
 *   mp_mask_batch / mp_verify_mask_batch       BarnettSmartProtocol::mask / verify_mask
 *                                              reference src/lib.rs:115-133, impl mod.rs:182-240
 *   mp_remask_prove_batch / mp_verify_remask_batch   ::remask / verify_remask
 *                                              reference src/lib.rs:136-154, impl mod.rs:242-299
 *   mp_reveal_batch / mp_verify_reveal_batch   ::compute_reveal_token / verify_reveal (one player,
 *                                              n masked cards)  src/lib.rs:157-175, impl mod.rs:301-354
 *   mp_key_ownership_prove_batch / _verify_batch   ::prove_key_ownership / verify_key_ownership
 *                                              src/lib.rs:88-104, impl mod.rs:132-165; info_offsets has
 *                                              n + 1 entries into the concatenated public-info bytes */
int32_t mp_mask_batch(mp_ctx* ctx, const uint8_t* shared_key /* 64 */, const uint8_t* cards /* n*64 */,
                      const uint8_t* r /* n*32 */, const uint8_t* omega /* n*32 */, uint64_t n,
                      uint8_t* out_masked /* n*128 */, uint8_t* out_proofs /* n*160 */, int32_t host_threads);
int32_t mp_verify_mask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* masked,
                             const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t mp_remask_prove_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck /* n*128 */,
                              const uint8_t* alpha /* n*32 */, const uint8_t* omega /* n*32 */, uint64_t n,
                              uint8_t* out_deck /* n*128 */, uint8_t* out_proofs /* n*160 */, int32_t host_threads);
int32_t mp_verify_remask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* remasked,
                               const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t mp_reveal_batch(mp_ctx* ctx, const uint8_t* sk /* 32 */, const uint8_t* pk /* 64 */,
                        const uint8_t* masked /* n*128 */, const uint8_t* omega /* n*32 */, uint64_t n,
                        uint8_t* out_tokens /* n*64 */, uint8_t* out_proofs /* n*160 */, int32_t host_threads);
int32_t mp_verify_reveal_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* tokens, const uint8_t* masked,
                               const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t mp_key_ownership_prove_batch(mp_ctx* ctx, const uint8_t* pks /* n*64 */, const uint8_t* sks /* n*32 */,
                                     const uint8_t* infos, const uint64_t* info_offsets /* n+1 */,
                                     const uint8_t* omega /* n*32 */, uint64_t n, uint8_t* out_proofs /* n*96 */,
                                     int32_t host_threads);
int32_t mp_key_ownership_verify_batch(mp_ctx* ctx, const uint8_t* pks, const uint8_t* infos,
                                      const uint64_t* info_offsets, const uint8_t* proofs, uint64_t n,
                                      int32_t* statuses, int32_t host_threads);

/* ---- wire format (Appendix A3) -------------------------------------
 * ark-serialize 0.3 encodings of what crosses the network in a round -- every public type of the trait
 * is bound by CanonicalSerialize + CanonicalDeserialize (reference src/lib.rs:45-71) and proof sizes are
 * measured with `serialized_size` (examples/parameter_selection.rs:95):
 *   compressed point = x (32 bytes LE) with flags in the top bits of the last byte: bit 7 = y is the
 *   larger of (y, -y), bit 6 = infinity;  Vec<T> = u64 LE length | items;  ciphertext = c1 | c2.
 * Serialisation is byte handling and runs on the host; DEserialisation needs one square root in F_p per
 * point (Tonelli-Shanks with a 192-bit two-adic part) and runs on the GPU.  mp_points_decompress fills
 * statuses[i] (may be NULL) with 0 ok, 1 malformed encoding, 2 x not on the curve and returns
 * MP_ERR_NOT_ON_CURVE if any item is rejected (rejected items are written as 64 zero bytes). */
int32_t mp_points_compress(const uint8_t* points /* n*64 */, uint64_t n, uint8_t* out /* n*32 */);
int32_t mp_points_decompress(mp_ctx* ctx, const uint8_t* in /* n*32 */, uint64_t n, uint8_t* out /* n*64 */,
                             int32_t* statuses /* n or NULL */);
/* Vec<MaskedCard>: 8 + 64 * n_cards bytes */
uint64_t mp_deck_serialized_len(uint64_t n_cards);
int32_t mp_deck_serialize(const uint8_t* deck /* n_cards*128 */, uint64_t n_cards, uint8_t* out);
/* *n_cards: in = capacity of out_deck in cards, out = cards in the buffer */
int32_t mp_deck_deserialize(mp_ctx* ctx, const uint8_t* in, uint64_t in_len, uint8_t* out_deck, uint64_t* n_cards);
/* REPOSITORY-PRIVATE proof container: the flat proof of this header with every point compressed and every
 * scalar as it is -- (11m+8)*32 + (5n+9)*32 bytes, NO length prefixes.  It is NOT the byte stream
 * `ZKProofShuffle::serialize` produces upstream: that struct is a nest of Vec<> fields (one u64 length prefix
 * each, order and nesting defined in the absent proof-essentials crate), so upstream's `serialized_size()` is
 * larger by 8 bytes per Vec field and the bytes do not round-trip with a Rust peer.  The element encodings
 * (compressed points, canonical scalars, the Vec<MaskedCard> deck format above) are ark-serialize 0.3's; only the
 * proof's framing is this repository's own.  Deserialisation validates like ark-serialize: points on the curve,
 * scalars below the group order (MP_ERR_NOT_ON_CURVE / MP_ERR_NOT_CANONICAL). */
uint64_t mp_proof_serialized_len(int32_t m, int32_t n);
int32_t mp_proof_serialize(int32_t m, int32_t n, const uint8_t* proof, uint8_t* out);
int32_t mp_proof_deserialize(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* in, uint8_t* out_proof);

/* ---- measurement ---------------------------------------------------------------------------
 * Per-launch CUDA-event timing (on the context's stream) of the bucket-accumulation kernel, the
 * dominant kernel of every MSM: collect returns the summed duration, the exact number of bucket
 * additions (mixed XYZZ adds) those launches executed, and the launch count, then resets. */
int32_t mp_profile_enable(mp_ctx* ctx, int32_t on);
int32_t mp_profile_collect(mp_ctx* ctx, double* accumulate_ms, uint64_t* bucket_adds, uint64_t* launches);
/* Same, restricted to the dominant launches (those within a factor two of the largest by additions):
 * what a per-kernel roofline should be quoted on when a step mixes one huge MSM with many tiny ones. */
int32_t mp_profile_collect_dominant(mp_ctx* ctx, double* ms, uint64_t* bucket_adds, uint64_t* launches);

/* ---- debug / parity hooks (exercise single device primitives; not used by the protocol) */
int32_t mp_dbg_fq_mul(mp_ctx* ctx, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out);
int32_t mp_dbg_point_add(mp_ctx* ctx, const uint8_t* p, const uint8_t* q, uint64_t n, uint8_t* out);
int32_t mp_dbg_scalar_mul(mp_ctx* ctx, const uint8_t* p, const uint8_t* k, uint64_t n, uint8_t* out);
/* host-side probe: milliseconds the Fiat-Shamir transcript needs to absorb n_points points */
double mp_dbg_transcript_ms(uint64_t n_points);
/* integer-pipe microbenchmarks: returns milliseconds for `iters` dependent-chain iterations
 * on a full-chip grid; *ops receives the number of counted operations executed.  which: 0 IMAD.WIDE,
 * 1 IMAD.LO, 2 fq_mul, 3 xyzz_madd, 4 fq_sqr, 5 carry-chained IMAD.WIDE pairs (counted as
 * wide multiply-adds), 6 IMAD.WIDE + IADD 1:1 (counted as wide multiply-adds), 7 DFMA (fma.rz.f64),
 * 8 DFMA + IMAD.WIDE 1:1 (counted as DFMAs), 9 DFMA + 32-bit add 1:1 (counted as DFMAs). */
int32_t mp_dbg_bench(mp_ctx* ctx, int32_t which, int32_t iters, float* ms, double* ops);

#ifdef __cplusplus
}
#endif
#endif /* MPSHUFFLE_H */

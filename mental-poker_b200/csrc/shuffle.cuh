// Bayer-Groth shuffle argument on the GPU: host-side protocol driver.
//
// Replaces, behind the C ABI, what the reference executes inside
//   DLCards::shuffle_and_remask   reference src/discrete_log_cards/mod.rs:380-418
//   DLCards::verify_shuffle       reference src/discrete_log_cards/mod.rs:420-443
// i.e. `MaskedCard::remask` (remasking.rs:9-22) over the permuted deck and
// `shuffle::ShuffleArgument::{prove,verify}` of the un-vendored proof-essentials crate
// (algebra restated in SURVEY.md Appendix B; transcript/draw order Appendix B.6).
//
// Division of labour (BASELINE north_star): the host owns the Fiat-Shamir transcript and the
// O(m + n) scalar glue; every group operation and every O(N) / O(m^2 n) scalar-vector
// operation runs in CUDA kernels.  Verification needs NO device->host round trip before the
// final verdict: all challenges derive from proof bytes, so the whole check is a handful of
// batched MSM jobs that must each evaluate to the identity.  Proving needs four round trips
// (c_A -> x, c_B -> y,z, the big commitment batch -> Hadamard challenges, zero-argument
// commitments -> remaining challenges).
#pragma once
#include <stdint.h>

#include "shuffle_host.hpp"  // shuffle_proof_len / shuffle_randomness_len and the host-only verifier pieces

struct mp_ctx;
// `*_src` arguments: where the device copy of a deck is taken from -- the host buffer itself
// (default) or a device pointer when the deck is already resident in HBM.  The host copy is
// always needed: the transcript hashes it on the CPU.

namespace mp {

struct ShuffleState;
struct MsmWorkspace;
void shuffle_state_destroy(ShuffleState*);

int32_t shuffle_set_params(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g, const uint8_t* ck_g,
                           const uint8_t* ck_h, const uint8_t* ghat);
MsmWorkspace* shuffle_bulk_workspace(const mp_ctx* ctx);  // nullptr before set_params
int32_t shuffle_m(const mp_ctx* ctx);
int32_t shuffle_n(const mp_ctx* ctx);

int32_t shuffle_remask(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                       const uint8_t* rho, uint64_t N, uint8_t* out_deck, const void* deck_src = nullptr,
                       const void** d_out_ret = nullptr, Transcript* fs_head = nullptr);
// (fs_head: when given, the head of the statement absorb -- parameters, pk, input deck -- is hashed
//  into it while the remask kernel and its copies run; pass the same transcript to shuffle_prove)
int32_t shuffle_commit_batch(mp_ctx* ctx, const uint8_t* values, const uint8_t* blinds, uint64_t k,
                             uint64_t len, uint8_t* out);
int32_t shuffle_prove(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2,
                      const uint32_t* perm, const uint8_t* rho, const uint8_t* rand, uint8_t* proof_out,
                      const void* deck2_src = nullptr, Transcript* fs_started = nullptr);
struct StatementHashes;
// (hashes / hash_index: the batch verifier's shared statement hashes, shuffle_internal.cuh; the transcript of this proof
//  then continues from hashes->wait(hash_index) instead of hashing the statement itself)
int32_t shuffle_verify(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2,
                       const uint8_t* proof, const void* deck_src = nullptr, const void* deck2_src = nullptr,
                       StatementHashes* hashes = nullptr, uint64_t hash_index = 0);

// B independent proofs under the same parameters and public key, verified in lockstep.
// (d_decks / d_decks2: optional device copies of the decks, B * N * 128 bytes each, used by the large-deck path)
int32_t shuffle_verify_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint8_t* decks2,
                             const uint8_t* proofs, uint64_t B, int32_t* statuses, int32_t host_threads,
                             const void* d_decks = nullptr, const void* d_decks2 = nullptr);

// Batched sigma protocols either side of the shuffle (sigma.cu)
int32_t sigma_mask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* r, const uint8_t* omega,
                         uint64_t n, uint8_t* out_masked, uint8_t* out_proofs, int32_t host_threads);
int32_t sigma_verify_mask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* masked,
                                const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t sigma_remask_prove_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* alpha,
                                 const uint8_t* omega, uint64_t n, uint8_t* out_deck, uint8_t* out_proofs, int32_t host_threads);
int32_t sigma_verify_remask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* remasked,
                                  const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t sigma_reveal_batch(mp_ctx* ctx, const uint8_t* sk, const uint8_t* pk, const uint8_t* masked, const uint8_t* omega,
                           uint64_t n, uint8_t* out_tokens, uint8_t* out_proofs, int32_t host_threads);
int32_t sigma_verify_reveal_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* tokens, const uint8_t* masked,
                                  const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);
int32_t sigma_key_ownership_prove_batch(mp_ctx* ctx, const uint8_t* pks, const uint8_t* sks, const uint8_t* infos,
                                        const uint64_t* info_off, const uint8_t* omega, uint64_t n, uint8_t* out_proofs,
                                        int32_t host_threads);
int32_t sigma_key_ownership_verify_batch(mp_ctx* ctx, const uint8_t* pks, const uint8_t* infos, const uint64_t* info_off,
                                         const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads);

// Wire format, device half (wire.cu): batched point decompression
int32_t wire_points_decompress(mp_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, int32_t* statuses);
int32_t wire_deck_deserialize(mp_ctx* ctx, const uint8_t* in, uint64_t in_len, uint8_t* out_deck, uint64_t* n_cards);
int32_t wire_proof_deserialize(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* in, uint8_t* out_proof);

bool shuffle_uses_small_deck_path(uint64_t n_cards);
int32_t shuffle_prove_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                            const uint8_t* rhos, const uint8_t* rands, uint64_t B, uint8_t* out_decks,
                            uint8_t* proofs, int32_t host_threads, const void* d_decks = nullptr);

}  // namespace mp

//! Drop-in for the shuffle hot path of `barnett_smart_card_protocol::discrete_log_cards::DLCards`
//! (reference src/discrete_log_cards/mod.rs:380-443) over libmpshuffle.so.
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image); it documents the
//! exact binding: which reference call each C entry point replaces, how arkworks types map
//! to the byte layouts of include/mpshuffle.h, and where the caller's RNG is consumed.
//!
//! `GpuDLCards` overrides `setup`, `shuffle_and_remask` and `verify_shuffle`, offers deck-wide
//! `mask_all` / `remask_all` / `reveal_all` helpers over the batched sigma-protocol entry points, and
//! delegates the remaining constant-size methods of the trait (key generation, aggregate key, unmask,
//! single-card mask / remask / reveal) to the reference's `DLCards`.
use std::os::raw::c_char;

#[repr(C)]
pub struct MpCtx {
    _private: [u8; 0],
}

extern "C" {
    fn mp_ctx_create(out: *mut *mut MpCtx, device: i32) -> i32;
    fn mp_ctx_destroy(ctx: *mut MpCtx);
    fn mp_last_error_string(ctx: *mut MpCtx) -> *const c_char;
    fn mp_verify_status_string(status: i32) -> *const c_char;
    // DLCards::setup (mod.rs:105-121) -> binds Parameters to the GPU context
    fn mp_ctx_set_params(ctx: *mut MpCtx, m: i32, n: i32, enc_g: *const u8, ck_g: *const u8, ck_h: *const u8, ghat: *const u8) -> i32;
    fn mp_proof_len(m: i32, n: i32) -> u64;
    fn mp_prover_randomness_len(m: i32, n: i32) -> u64;
    // DLCards::shuffle_and_remask (mod.rs:380-418)
    fn mp_shuffle_and_remask(ctx: *mut MpCtx, pk: *const u8, deck: *const u8, perm: *const u32, rho: *const u8,
                             randomness: *const u8, out_deck: *mut u8, proof_out: *mut u8) -> i32;
    // DLCards::verify_shuffle (mod.rs:420-443)
    fn mp_shuffle_verify(ctx: *mut MpCtx, pk: *const u8, deck: *const u8, shuffled: *const u8, proof: *const u8) -> i32;
    // Batched sigma protocols either side of the shuffle (one call per deck instead of one trait call per card):
    // DLCards::mask / verify_mask (mod.rs:182-240)
    fn mp_mask_batch(ctx: *mut MpCtx, shared_key: *const u8, cards: *const u8, r: *const u8, omega: *const u8, n: u64,
                     out_masked: *mut u8, out_proofs: *mut u8, host_threads: i32) -> i32;
    fn mp_verify_mask_batch(ctx: *mut MpCtx, shared_key: *const u8, cards: *const u8, masked: *const u8, proofs: *const u8,
                            n: u64, statuses: *mut i32, host_threads: i32) -> i32;
    // DLCards::remask / verify_remask (mod.rs:242-299)
    fn mp_remask_prove_batch(ctx: *mut MpCtx, shared_key: *const u8, deck: *const u8, alpha: *const u8, omega: *const u8,
                             n: u64, out_deck: *mut u8, out_proofs: *mut u8, host_threads: i32) -> i32;
    fn mp_verify_remask_batch(ctx: *mut MpCtx, shared_key: *const u8, deck: *const u8, remasked: *const u8,
                              proofs: *const u8, n: u64, statuses: *mut i32, host_threads: i32) -> i32;
    // DLCards::compute_reveal_token / verify_reveal (mod.rs:301-354), one player, n masked cards
    fn mp_reveal_batch(ctx: *mut MpCtx, sk: *const u8, pk: *const u8, masked: *const u8, omega: *const u8, n: u64,
                       out_tokens: *mut u8, out_proofs: *mut u8, host_threads: i32) -> i32;
    fn mp_verify_reveal_batch(ctx: *mut MpCtx, pk: *const u8, tokens: *const u8, masked: *const u8, proofs: *const u8,
                              n: u64, statuses: *mut i32, host_threads: i32) -> i32;
    // DLCards::prove_key_ownership / verify_key_ownership (mod.rs:132-165)
    fn mp_key_ownership_prove_batch(ctx: *mut MpCtx, pks: *const u8, sks: *const u8, infos: *const u8,
                                    info_offsets: *const u64, omega: *const u8, n: u64, out_proofs: *mut u8,
                                    host_threads: i32) -> i32;
    // CanonicalSerialize / CanonicalDeserialize of decks and proofs (bounds at src/lib.rs:45-71)
    fn mp_deck_serialized_len(n_cards: u64) -> u64;
    fn mp_deck_serialize(deck: *const u8, n_cards: u64, out: *mut u8) -> i32;
    fn mp_deck_deserialize(ctx: *mut MpCtx, input: *const u8, in_len: u64, out_deck: *mut u8, n_cards: *mut u64) -> i32;
    fn mp_proof_serialized_len(m: i32, n: i32) -> u64;
    fn mp_proof_serialize(m: i32, n: i32, proof: *const u8, out: *mut u8) -> i32;
    fn mp_proof_deserialize(ctx: *mut MpCtx, m: i32, n: i32, input: *const u8, out_proof: *mut u8) -> i32;
    fn mp_key_ownership_verify_batch(ctx: *mut MpCtx, pks: *const u8, infos: *const u8, info_offsets: *const u64,
                                     proofs: *const u8, n: u64, statuses: *mut i32, host_threads: i32) -> i32;
}

/// Owns the opaque GPU context; `Parameters` of the shim holds one (the reference's own
/// `Parameters` has private fields, mod.rs:37-43, so the shim defines its own type).
pub struct GpuContext(*mut MpCtx);
unsafe impl Send for GpuContext {}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { mp_ctx_destroy(self.0) }
    }
}

/// 32-byte little-endian canonical encoding = `ark_ff::ToBytes` of `Fp256`.
pub fn fr_bytes<F: ark_ff::PrimeField>(x: &F, out: &mut Vec<u8>) {
    ark_ff::ToBytes::write(&x.into_repr(), out).unwrap();
}

/// 64-byte x || y; the identity is all zeros (include/mpshuffle.h).
pub fn point_bytes<A: ark_ec::AffineCurve>(p: &A, out: &mut Vec<u8>)
where
    A::BaseField: ark_ff::PrimeField,
{
    if p.is_zero() {
        out.extend_from_slice(&[0u8; 64]);
    } else {
        // x.into_repr().write ; y.into_repr().write   (little-endian limbs)
        let mut buf = Vec::with_capacity(65);
        ark_ff::ToBytes::write(p, &mut buf).unwrap(); // x || y || infinity flag (ark-ec 0.3)
        out.extend_from_slice(&buf[..64]);
    }
}

/// Second curve (include/mpshuffle_bls12_377.h): the group layer of `DLCards<ark_bls12_377::G1Projective>`
/// (reference examples/parameter_selection.rs:25-29).  48-byte coordinates, 96-byte points, 32-byte scalars.
#[repr(C)]
pub struct Mp377Ctx {
    _private: [u8; 0],
}
extern "C" {
    pub fn mp377_ctx_create(out: *mut *mut Mp377Ctx, device: i32) -> i32;
    pub fn mp377_ctx_destroy(ctx: *mut Mp377Ctx);
    pub fn mp377_last_error_string(ctx: *mut Mp377Ctx) -> *const c_char;
    pub fn mp377_msm_g1(ctx: *mut Mp377Ctx, bases: *const u8, scalars: *const u8, n: u64, window_bits: i32, out: *mut u8) -> i32;
    pub fn mp377_ct_msm(ctx: *mut Mp377Ctx, deck: *const u8, scalars: *const u8, n: u64, window_bits: i32, out: *mut u8) -> i32;
    pub fn mp377_set_commit_key(ctx: *mut Mp377Ctx, ck: *const u8, len: u64) -> i32;
    pub fn mp377_pedersen_commit_batch(ctx: *mut Mp377Ctx, values: *const u8, blinds: *const u8, k: u64, len: u64, out: *mut u8) -> i32;
    pub fn mp377_msm_jobs(ctx: *mut Mp377Ctx, points: *const u8, n_points: u64, ncomp: i32, scalars: *const u8, n_scalars: u64,
                          jobs: *const u32, njobs: u64, window_bits: i32, out: *mut u8) -> i32;
    // BarnettSmartProtocol::verify_shuffle over BLS12-377 (lib.rs:191-197): 0 = Ok(()), > 0 = failed check, < 0 = error
    pub fn mp377_proof_len(m: i32, n: i32) -> u64;
    pub fn mp377_shuffle_verify(ctx: *mut Mp377Ctx, m: i32, n: i32, enc_g: *const u8, ck_g: *const u8, ck_h: *const u8, ghat: *const u8,
                                pk: *const u8, deck: *const u8, shuffled: *const u8, proof: *const u8) -> i32;
}

/// Sketch of the two overridden trait methods (generic bounds elided):
///
/// ```ignore
/// fn shuffle_and_remask<R: Rng>(rng, pp, shared_key, deck, masking_factors, permutation) {
///     // 1. serialise: deck -> N * 128 bytes, masking_factors -> N * 32 bytes,
///     //    permutation.mapping -> N * u32, shared_key -> 64 bytes
///     // 2. draw the prover randomness from the caller's rng, in the order documented in
///     //    include/mpshuffle.h:  (0..mp_prover_randomness_len(m, n)).map(|_| Scalar::rand(rng))
///     // 3. mp_shuffle_and_remask(ctx, pk, deck, perm, rho, rand, out_deck, proof)
///     //    status < 0  => Err(CardProtocolError::IoError(mp_last_error_string(ctx)))
///     // 4. deserialise out_deck (N ciphertexts) and keep `proof` as the opaque byte proof
///     //    (ZKProofShuffle = Vec<u8> for this implementor; CanonicalSerialize is satisfied)
/// }
/// fn verify_shuffle(pp, shared_key, original_deck, shuffled_deck, proof) {
///     match mp_shuffle_verify(ctx, pk, deck, shuffled, proof) {
///         0 => Ok(()),
///         s if s > 0 => Err(CryptoError::ProofVerificationError(mp_verify_status_string(s))),
///         _ => Err(CryptoError::ProofVerificationError(mp_last_error_string(ctx))),
///     }
/// }
/// ```
pub struct GpuDLCards;

//! Drop-in for the shuffle hot path of `barnett_smart_card_protocol::discrete_log_cards::DLCards`
//! (reference src/discrete_log_cards/mod.rs:380-443) over libmpshuffle.so.
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image).  It is written to be built unchanged by a
//! maintainer with `cargo`: the extern "C" block, the byte conversions between arkworks types and the layouts of
//! include/mpshuffle.h, and a complete `impl BarnettSmartProtocol for GpuDLCards<C>` (all fourteen methods).
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
    // BarnettSmartProtocol::{mask, verify_mask, remask, verify_remask, compute_reveal_token, verify_reveal, prove_key_ownership,
    // verify_key_ownership} (lib.rs:88-175) for n items per call: Chaum-Pedersen proof = a | b | r = 224 bytes, Schnorr = 128
    pub fn mp377_mask_batch(ctx: *mut Mp377Ctx, shared_key: *const u8, cards: *const u8, r: *const u8, omega: *const u8, n: u64,
                            out_masked: *mut u8, out_proofs: *mut u8, host_threads: i32) -> i32;
    pub fn mp377_verify_mask_batch(ctx: *mut Mp377Ctx, shared_key: *const u8, cards: *const u8, masked: *const u8, proofs: *const u8,
                                   n: u64, statuses: *mut i32, host_threads: i32) -> i32;
    pub fn mp377_remask_prove_batch(ctx: *mut Mp377Ctx, shared_key: *const u8, deck: *const u8, alpha: *const u8, omega: *const u8,
                                    n: u64, out_deck: *mut u8, out_proofs: *mut u8, host_threads: i32) -> i32;
    pub fn mp377_verify_remask_batch(ctx: *mut Mp377Ctx, shared_key: *const u8, deck: *const u8, remasked: *const u8,
                                     proofs: *const u8, n: u64, statuses: *mut i32, host_threads: i32) -> i32;
    pub fn mp377_reveal_batch(ctx: *mut Mp377Ctx, sk: *const u8, pk: *const u8, masked: *const u8, omega: *const u8, n: u64,
                              out_tokens: *mut u8, out_proofs: *mut u8, host_threads: i32) -> i32;
    pub fn mp377_verify_reveal_batch(ctx: *mut Mp377Ctx, pk: *const u8, tokens: *const u8, masked: *const u8, proofs: *const u8,
                                     n: u64, statuses: *mut i32, host_threads: i32) -> i32;
    pub fn mp377_key_ownership_prove_batch(ctx: *mut Mp377Ctx, pks: *const u8, sks: *const u8, infos: *const u8,
                                           info_offsets: *const u64, omega: *const u8, n: u64, out_proofs: *mut u8,
                                           host_threads: i32) -> i32;
    pub fn mp377_key_ownership_verify_batch(ctx: *mut Mp377Ctx, pks: *const u8, infos: *const u8, info_offsets: *const u64,
                                            proofs: *const u8, n: u64, statuses: *mut i32, host_threads: i32) -> i32;
    // CanonicalDeserialize of points / decks / proofs (lib.rs:45-71): square roots + G1 membership on the GPU
    pub fn mp377_points_decompress(ctx: *mut Mp377Ctx, input: *const u8 /* n*48 */, n: u64, out: *mut u8 /* n*96 */,
                                   statuses: *mut i32) -> i32;
    pub fn mp377_deck_deserialize(ctx: *mut Mp377Ctx, input: *const u8, in_len: u64, out_deck: *mut u8, n_cards: *mut u64) -> i32;
    pub fn mp377_proof_deserialize(ctx: *mut Mp377Ctx, m: i32, n: i32, input: *const u8, out_proof: *mut u8) -> i32;
}

// =============================================================================================================
// impl BarnettSmartProtocol for GpuDLCards<C>
//
// The complete replacement a maintainer builds with `cargo build` next to the reference crate (reference trait:
// src/lib.rs:41-198; reference implementor: src/discrete_log_cards/mod.rs:86-444).  Three methods go to the GPU --
// `setup` (it must own the parameters: the reference's `Parameters` has private fields, mod.rs:37-43),
// `shuffle_and_remask` (mod.rs:380-418) and `verify_shuffle` (mod.rs:420-443); the eleven constant-size methods
// delegate to the reference's `DLCards<C>` through a `Parameters` rebuilt with its public constructor (mod.rs:45-60).
// NOT COMPILED HERE (no Rust toolchain in the build image) -- names of items inside the un-vendored
// `proof-essentials` crate (`el_gamal::Parameters { generator }`, `pedersen::CommitKey { g, h }`,
// `Ciphertext(pub C::Affine, pub C::Affine)`, `Permutation { mapping }`) are as SURVEY.md Appendix A recalls them and
// are the only places that may need a rename.
// =============================================================================================================
use ark_ec::{AffineCurve, ProjectiveCurve};
use ark_ff::{to_bytes, FromBytes, PrimeField, ToBytes, UniformRand};
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize, Read, SerializationError, Write};
use barnett_smart_card_protocol::discrete_log_cards::{self as dl, DLCards};
use barnett_smart_card_protocol::error::CardProtocolError;
use barnett_smart_card_protocol::BarnettSmartProtocol;
use proof_essentials::error::CryptoError;
use proof_essentials::homomorphic_encryption::{el_gamal, el_gamal::ElGamal, HomomorphicEncryptionScheme};
use proof_essentials::utils::permutation::Permutation;
use proof_essentials::vector_commitment::{pedersen, pedersen::PedersenCommitment, HomomorphicCommitmentScheme};
use rand::Rng;
use std::ffi::CStr;
use std::marker::PhantomData;
use std::sync::Mutex;

/// Which half of libmpshuffle.so serves a curve.  `FE` = bytes per base-field element (32: Stark curve, `mp_*`;
/// 48: BLS12-377 G1, `mp377_*`).  A point is x || y (2 * FE bytes, all zero = identity), a ciphertext c1 || c2.
pub trait GpuCurve: ProjectiveCurve {
    const FE: usize;
}
impl GpuCurve for starknet_curve::Projective {
    const FE: usize = 32;
}
impl GpuCurve for ark_bls12_377::G1Projective {
    const FE: usize = 48;
}

fn point_to<C: GpuCurve>(p: &C::Affine, out: &mut Vec<u8>) {
    if p.is_zero() {
        out.extend(std::iter::repeat(0u8).take(2 * C::FE));
    } else {
        let b = to_bytes![p].unwrap(); // ark-ec 0.3 `GroupAffine::write`: x || y || infinity flag, canonical LE limbs
        out.extend_from_slice(&b[..2 * C::FE]);
    }
}
fn point_from<C: GpuCurve>(b: &[u8]) -> C::Affine {
    if b.iter().all(|&v| v == 0) {
        return C::Affine::zero();
    }
    let mut buf = b[..2 * C::FE].to_vec();
    buf.push(0); // infinity flag
    C::Affine::read(&buf[..]).unwrap() // the library has already validated the point (on-curve, canonical, subgroup)
}
fn scalar_to<F: PrimeField>(x: &F, out: &mut Vec<u8>) {
    out.extend_from_slice(&to_bytes![x.into_repr()].unwrap()[..32]);
}
fn deck_to<C: GpuCurve>(deck: &[el_gamal::Ciphertext<C>]) -> Vec<u8> {
    let mut v = Vec::with_capacity(deck.len() * 4 * C::FE);
    for c in deck {
        point_to::<C>(&c.0, &mut v);
        point_to::<C>(&c.1, &mut v);
    }
    v
}
fn last_error(ctx: *mut MpCtx) -> String {
    unsafe { CStr::from_ptr(mp_last_error_string(ctx)).to_string_lossy().into_owned() }
}
fn status_string(s: i32) -> String {
    unsafe { CStr::from_ptr(mp_verify_status_string(s)).to_string_lossy().into_owned() }
}

/// `Parameters` of the GPU implementor: the reference's own (for the delegated methods), the pieces the reference
/// keeps private, and the GPU context they were uploaded to.  A context is used by one thread at a time.
pub struct GpuParameters<C: GpuCurve> {
    pub m: usize,
    pub n: usize,
    enc_parameters: el_gamal::Parameters<C>,
    commit_parameters: pedersen::CommitKey<C>,
    generator: el_gamal::Generator<C>,
    reference: dl::Parameters<C>,
    ctx: Mutex<GpuContext>,
    // byte forms kept for the curve whose verifier takes the parameters per call (BLS12-377)
    enc_g: Vec<u8>,
    ck_g: Vec<u8>,
    ck_h: Vec<u8>,
    ghat: Vec<u8>,
}

/// The shuffle proof of this implementor: the flat layout of include/mpshuffle.h.  (Upstream's
/// `shuffle::proof::Proof` is a nest of structs in the absent crate; the trait only asks for Canonical(De)Serialize.)
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct GpuShuffleProof(pub Vec<u8>);
impl CanonicalSerialize for GpuShuffleProof {
    fn serialize<W: Write>(&self, mut w: W) -> Result<(), SerializationError> {
        (self.0.len() as u64).serialize(&mut w)?;
        w.write_all(&self.0)?;
        Ok(())
    }
    fn serialized_size(&self) -> usize {
        8 + self.0.len()
    }
}
impl CanonicalDeserialize for GpuShuffleProof {
    fn deserialize<R: Read>(mut r: R) -> Result<Self, SerializationError> {
        let len = u64::deserialize(&mut r)? as usize;
        let mut v = vec![0u8; len];
        r.read_exact(&mut v)?;
        Ok(GpuShuffleProof(v)) // points / scalars are validated by mp_shuffle_verify (on-curve, subgroup, canonical)
    }
}

pub struct GpuDLCards<C: GpuCurve> {
    _group: PhantomData<&'static C>,
}

impl<C: GpuCurve> BarnettSmartProtocol for GpuDLCards<C>
where
    el_gamal::Parameters<C>: Clone,
    pedersen::CommitKey<C>: Clone,
    el_gamal::Generator<C>: Clone,
{
    type Scalar = C::ScalarField;
    type Enc = ElGamal<C>;
    type Comm = PedersenCommitment<C>;
    type Parameters = GpuParameters<C>;
    type PlayerPublicKey = dl::PublicKey<C>;
    type PlayerSecretKey = dl::PlayerSecretKey<C>;
    type AggregatePublicKey = dl::PublicKey<C>;
    type Card = dl::Card<C>;
    type MaskedCard = dl::MaskedCard<C>;
    type RevealToken = dl::RevealToken<C>;
    type ZKProofKeyOwnership = <DLCards<C> as BarnettSmartProtocol>::ZKProofKeyOwnership;
    type ZKProofMasking = <DLCards<C> as BarnettSmartProtocol>::ZKProofMasking;
    type ZKProofRemasking = <DLCards<C> as BarnettSmartProtocol>::ZKProofRemasking;
    type ZKProofReveal = <DLCards<C> as BarnettSmartProtocol>::ZKProofReveal;
    type ZKProofShuffle = GpuShuffleProof;

    /// mod.rs:105-121 -- the same three draws in the same order, then the upload (mp_ctx_set_params builds the
    /// fixed-base tables of the commit key once per Parameters).
    fn setup<R: Rng>(rng: &mut R, m: usize, n: usize) -> Result<Self::Parameters, CardProtocolError> {
        let enc_parameters = Self::Enc::setup(rng)?;
        let commit_parameters = Self::Comm::setup(rng, n);
        let generator = Self::Enc::generator(rng)?;
        let (mut enc_g, mut ck_g, mut ck_h, mut ghat) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        point_to::<C>(&enc_parameters.generator, &mut enc_g);
        for g in commit_parameters.g.iter() {
            point_to::<C>(g, &mut ck_g);
        }
        point_to::<C>(&commit_parameters.h, &mut ck_h);
        point_to::<C>(&generator.0, &mut ghat);
        let mut raw: *mut MpCtx = std::ptr::null_mut();
        if unsafe { mp_ctx_create(&mut raw, 0) } != 0 {
            return Err(CardProtocolError::IoError("no CUDA device: libmpshuffle has no CPU fallback".into()));
        }
        let ctx = GpuContext(raw);
        if C::FE == 32 {
            let st = unsafe { mp_ctx_set_params(raw, m as i32, n as i32, enc_g.as_ptr(), ck_g.as_ptr(), ck_h.as_ptr(), ghat.as_ptr()) };
            if st != 0 {
                return Err(CardProtocolError::IoError(last_error(raw)));
            }
        }
        let reference = dl::Parameters::new(m, n, enc_parameters.clone(), commit_parameters.clone(), generator.clone());
        Ok(GpuParameters { m, n, enc_parameters, commit_parameters, generator, reference, ctx: Mutex::new(ctx), enc_g, ck_g, ck_h, ghat })
    }

    /// mod.rs:380-418: permute, remask every card, ShuffleArgument::prove -- one C call.  The library never owns an
    /// RNG: the prover's 11m + 5n scalars are drawn here from the caller's `rng`, in the order of SURVEY.md B.6.
    fn shuffle_and_remask<R: Rng>(
        rng: &mut R,
        pp: &Self::Parameters,
        shared_key: &Self::AggregatePublicKey,
        deck: &Vec<Self::MaskedCard>,
        masking_factors: &Vec<Self::Scalar>,
        permutation: &Permutation,
    ) -> Result<(Vec<Self::MaskedCard>, Self::ZKProofShuffle), CardProtocolError> {
        let big_n = pp.m * pp.n;
        if deck.len() != big_n || masking_factors.len() != big_n || permutation.mapping.len() != big_n {
            return Err(CardProtocolError::IoError("deck, masking factors and permutation must have m * n entries".into()));
        }
        if C::FE != 32 {
            return Err(CardProtocolError::IoError("shuffle_and_remask on the GPU is built for the Stark curve; see mpshuffle_bls12_377.h".into()));
        }
        let deck_b = deck_to::<C>(deck);
        let mut pk = Vec::new();
        point_to::<C>(shared_key, &mut pk);
        let perm: Vec<u32> = permutation.mapping.iter().map(|&i| i as u32).collect();
        let mut rho = Vec::with_capacity(32 * big_n);
        masking_factors.iter().for_each(|f| scalar_to(f, &mut rho));
        let nrand = unsafe { mp_prover_randomness_len(pp.m as i32, pp.n as i32) } as usize;
        let mut rand = Vec::with_capacity(32 * nrand);
        (0..nrand).for_each(|_| scalar_to(&Self::Scalar::rand(rng), &mut rand));
        let mut out_deck = vec![0u8; 128 * big_n];
        let mut proof = vec![0u8; unsafe { mp_proof_len(pp.m as i32, pp.n as i32) } as usize];
        let guard = pp.ctx.lock().unwrap();
        let st = unsafe {
            mp_shuffle_and_remask(guard.0, pk.as_ptr(), deck_b.as_ptr(), perm.as_ptr(), rho.as_ptr(), rand.as_ptr(), out_deck.as_mut_ptr(), proof.as_mut_ptr())
        };
        if st != 0 {
            return Err(CardProtocolError::IoError(last_error(guard.0)));
        }
        let shuffled = out_deck
            .chunks_exact(128)
            .map(|c| el_gamal::Ciphertext::<C>(point_from::<C>(&c[..64]), point_from::<C>(&c[64..])))
            .collect();
        Ok((shuffled, GpuShuffleProof(proof)))
    }

    /// mod.rs:420-443.  Status > 0 names the failing sub-argument with the reference's own strings
    /// ("Hadamard Product (5.1)", tests.rs:223-225); malformed inputs are what upstream's deserialiser would reject.
    fn verify_shuffle(
        pp: &Self::Parameters,
        shared_key: &Self::AggregatePublicKey,
        original_deck: &Vec<Self::MaskedCard>,
        shuffled_deck: &Vec<Self::MaskedCard>,
        proof: &Self::ZKProofShuffle,
    ) -> Result<(), CryptoError> {
        let big_n = pp.m * pp.n;
        if original_deck.len() != big_n || shuffled_deck.len() != big_n {
            return Err(CryptoError::ProofVerificationError("deck length".into()));
        }
        let (d1, d2) = (deck_to::<C>(original_deck), deck_to::<C>(shuffled_deck));
        let mut pk = Vec::new();
        point_to::<C>(shared_key, &mut pk);
        let guard = pp.ctx.lock().unwrap();
        let st = if C::FE == 32 {
            if proof.0.len() as u64 != unsafe { mp_proof_len(pp.m as i32, pp.n as i32) } {
                return Err(CryptoError::ProofVerificationError("proof length".into()));
            }
            unsafe { mp_shuffle_verify(guard.0, pk.as_ptr(), d1.as_ptr(), d2.as_ptr(), proof.0.as_ptr()) }
        } else {
            if proof.0.len() as u64 != unsafe { mp377_proof_len(pp.m as i32, pp.n as i32) } {
                return Err(CryptoError::ProofVerificationError("proof length".into()));
            }
            let mut c377: *mut Mp377Ctx = std::ptr::null_mut();
            if unsafe { mp377_ctx_create(&mut c377, 0) } != 0 {
                return Err(CryptoError::ProofVerificationError("no CUDA device".into()));
            }
            let s = unsafe {
                mp377_shuffle_verify(c377, pp.m as i32, pp.n as i32, pp.enc_g.as_ptr(), pp.ck_g.as_ptr(), pp.ck_h.as_ptr(), pp.ghat.as_ptr(),
                                     pk.as_ptr(), d1.as_ptr(), d2.as_ptr(), proof.0.as_ptr())
            };
            unsafe { mp377_ctx_destroy(c377) };
            s
        };
        match st {
            0 => Ok(()),
            s if s > 0 => Err(CryptoError::ProofVerificationError(status_string(s))),
            _ => Err(CryptoError::ProofVerificationError(last_error(guard.0))),
        }
    }

    // ---- the eleven constant-size methods: the reference's own implementation (mod.rs:123-378) ----
    fn player_keygen<R: Rng>(rng: &mut R, pp: &Self::Parameters) -> Result<(Self::PlayerPublicKey, Self::PlayerSecretKey), CardProtocolError> {
        DLCards::<C>::player_keygen(rng, &pp.reference)
    }
    fn prove_key_ownership<B: ToBytes, R: Rng>(rng: &mut R, pp: &Self::Parameters, pk: &Self::PlayerPublicKey, sk: &Self::PlayerSecretKey,
                                               player_public_info: &B) -> Result<Self::ZKProofKeyOwnership, CryptoError> {
        DLCards::<C>::prove_key_ownership(rng, &pp.reference, pk, sk, player_public_info)
    }
    fn verify_key_ownership<B: ToBytes>(pp: &Self::Parameters, pk: &Self::PlayerPublicKey, player_public_info: &B,
                                        proof: &Self::ZKProofKeyOwnership) -> Result<(), CryptoError> {
        DLCards::<C>::verify_key_ownership(&pp.reference, pk, player_public_info, proof)
    }
    fn compute_aggregate_key<B: ToBytes>(pp: &Self::Parameters,
                                         player_keys_proof_info: &Vec<(Self::PlayerPublicKey, Self::ZKProofKeyOwnership, B)>)
                                         -> Result<Self::AggregatePublicKey, CardProtocolError> {
        DLCards::<C>::compute_aggregate_key(&pp.reference, player_keys_proof_info)
    }
    fn mask<R: Rng>(rng: &mut R, pp: &Self::Parameters, shared_key: &Self::AggregatePublicKey, original_card: &Self::Card,
                    alpha: &Self::Scalar) -> Result<(Self::MaskedCard, Self::ZKProofMasking), CardProtocolError> {
        DLCards::<C>::mask(rng, &pp.reference, shared_key, original_card, alpha)
    }
    fn verify_mask(pp: &Self::Parameters, shared_key: &Self::AggregatePublicKey, card: &Self::Card, masked_card: &Self::MaskedCard,
                   proof: &Self::ZKProofMasking) -> Result<(), CryptoError> {
        DLCards::<C>::verify_mask(&pp.reference, shared_key, card, masked_card, proof)
    }
    fn remask<R: Rng>(rng: &mut R, pp: &Self::Parameters, shared_key: &Self::AggregatePublicKey, original_masked: &Self::MaskedCard,
                      alpha: &Self::Scalar) -> Result<(Self::MaskedCard, Self::ZKProofRemasking), CardProtocolError> {
        DLCards::<C>::remask(rng, &pp.reference, shared_key, original_masked, alpha)
    }
    fn verify_remask(pp: &Self::Parameters, shared_key: &Self::AggregatePublicKey, original_masked: &Self::MaskedCard,
                     remasked: &Self::MaskedCard, proof: &Self::ZKProofRemasking) -> Result<(), CryptoError> {
        DLCards::<C>::verify_remask(&pp.reference, shared_key, original_masked, remasked, proof)
    }
    fn compute_reveal_token<R: Rng>(rng: &mut R, pp: &Self::Parameters, sk: &Self::PlayerSecretKey, pk: &Self::PlayerPublicKey,
                                    masked_card: &Self::MaskedCard) -> Result<(Self::RevealToken, Self::ZKProofReveal), CardProtocolError> {
        DLCards::<C>::compute_reveal_token(rng, &pp.reference, sk, pk, masked_card)
    }
    fn verify_reveal(pp: &Self::Parameters, pk: &Self::PlayerPublicKey, reveal_token: &Self::RevealToken, masked_card: &Self::MaskedCard,
                     proof: &Self::ZKProofReveal) -> Result<(), CryptoError> {
        DLCards::<C>::verify_reveal(&pp.reference, pk, reveal_token, masked_card, proof)
    }
    fn unmask(pp: &Self::Parameters, decryption_key: &Vec<(Self::RevealToken, Self::ZKProofReveal, Self::PlayerPublicKey)>,
              masked_card: &Self::MaskedCard) -> Result<Self::Card, CardProtocolError> {
        DLCards::<C>::unmask(&pp.reference, decryption_key, masked_card)
    }
}

// Deck-wide helpers over the batched sigma entry points (one C call per deck instead of one trait call per card;
// round.rs:253-256 masks 52 cards one by one) are a thin loop over mp_mask_batch / mp_remask_prove_batch /
// mp_reveal_batch with the same byte conversions as above: proofs come back as 160-byte (a, b, r) records that
// `chaum_pedersen_dl_equality::proof::Proof::deserialize` reads after `point_from` / `FromBytes` per field.

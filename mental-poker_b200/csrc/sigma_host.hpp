// Host-only half of the batched sigma protocols either side of the shuffle (SURVEY.md section 8(f),
// rank 1): the Fiat-Shamir challenges and responses of
//   Schnorr identification     prove_key_ownership / verify_key_ownership   reference mod.rs:132-165
//   Chaum-Pedersen DL equality mask / remask / reveal and their verifiers   reference mod.rs:182-354
// (seeds mod.rs:80-83).  The protocol bodies live in the un-vendored `proof-essentials` crate; they are
// restated as in oracle/py/sigma.py [UPSTREAM-RECALL], transcript byte order = this repository's
// definition (PARITY UNPINNED against upstream bytes).  No CUDA here: tests/host/host_shim.cpp
// compiles this header with g++ (tests/test_host_sigma.py); the product includes it from sigma.cu.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/mpshuffle.h"
#include "fr.cuh"
#include "transcript.hpp"

namespace mp {

inline constexpr const char* kSeedKeyOwnership = "Key Ownership Proof";  // mod.rs:80
inline constexpr const char* kSeedMasking = "Masking Proof";             // mod.rs:81
inline constexpr const char* kSeedRemasking = "Remasking Proof";         // mod.rs:82
inline constexpr const char* kSeedReveal = "Reveal Proof";               // mod.rs:83

// widths of the curve fq.cuh selects: C-ABI point = x || y = 64 bytes (Stark) / 96 bytes (BLS12-377 G1); the
// transcript absorbs the ark-ec `ToBytes` form x || y || infinity flag, one byte longer
inline constexpr size_t kSigmaPt = 8 * (size_t)kFqLimbs, kSigmaPtFs = kSigmaPt + 1;
inline constexpr size_t kCpProofLen = 2 * kSigmaPt + 32;   // a | b | r          (160 bytes on the Stark curve)
inline constexpr size_t kSchnorrProofLen = kSigmaPt + 32;  // commit | opening   (96 bytes on the Stark curve)

// c = challenge after absorbing  "chaum_pedersen" | g | h | s0 | s1 | a | b  into a transcript seeded
// with `seeded` (passed by value: from_seed is hashed once per batch, not once per proof)
// C-ABI point (all-zero = identity) -> the ark-ec encoding the transcript absorbs
inline void point65(const uint8_t* p, uint8_t* out) {
  uint64_t w[kSigmaPt / 8], any = 0;
  memcpy(w, p, kSigmaPt);
  for (size_t k = 0; k < kSigmaPt / 8; k++) any |= w[k];
  if (any == 0) {
    memset(out, 0, kSigmaPtFs);  // identity = (0, 1, infinity)
    out[kSigmaPt / 2] = 1;
    out[kSigmaPt] = 1;
  } else {
    memcpy(out, p, kSigmaPt);
    out[kSigmaPt] = 0;
  }
}

inline fr cp_challenge(Transcript seeded, const uint8_t* g, const uint8_t* h, const uint8_t* s0, const uint8_t* s1,
                       const uint8_t* a, const uint8_t* b) {
  // one 404-byte message (Stark curve), fed in one piece: the long-input Blake2s path takes six of its seven blocks
  uint8_t msg[14 + 6 * kSigmaPtFs];
  memcpy(msg, "chaum_pedersen", 14);
  const uint8_t* pts[6] = {g, h, s0, s1, a, b};
  for (int k = 0; k < 6; k++) point65(pts[k], msg + 14 + kSigmaPtFs * k);
  seeded.begin();
  seeded.feed(msg, sizeof msg);
  seeded.end();
  return seeded.challenge();
}

// seed = "Key Ownership Proof" || info;  c = challenge after absorbing "schnorr_identity" | g | pk | commit
inline fr schnorr_challenge(const uint8_t* info, size_t info_len, const uint8_t* g, const uint8_t* pk, const uint8_t* commit) {
  std::string seed(kSeedKeyOwnership);
  seed.append(reinterpret_cast<const char*>(info), info_len);
  Transcript fs(seed.data(), seed.size());
  fs.begin();
  fs.feed_label("schnorr_identity");
  fs.feed_points64(g, 1);
  fs.feed_points64(pk, 1);
  fs.feed_points64(commit, 1);
  fs.end();
  return fs.challenge();
}

// canonical 32-byte scalar below the group order?
inline bool fr_bytes_canonical(const uint8_t* b) {
  uint32_t w[8];
  memcpy(w, b, 32);
  for (int i = 7; i >= 0; i--) {
    const uint32_t mi = fr_modulus_limb(i);
    if (w[i] != mi) return w[i] < mi;
  }
  return false;
}
inline fr fr_from_bytes(const uint8_t* b) {
  uint32_t w[8];
  memcpy(w, b, 32);
  return fr_from_canonical(w);
}
inline void fr_to_bytes(const fr& a, uint8_t* b) {
  uint32_t w[8];
  fr_to_canonical(a, w);
  memcpy(b, w, 32);
}

}  // namespace mp

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

inline constexpr size_t kCpProofLen = 160;       // a (64) | b (64) | r (32)
inline constexpr size_t kSchnorrProofLen = 96;   // commit (64) | opening (32)

// c = challenge after absorbing  "chaum_pedersen" | g | h | s0 | s1 | a | b  into a transcript seeded
// with `seeded` (passed by value: from_seed is hashed once per batch, not once per proof)
// 64-byte C-ABI point (all-zero = identity) -> the 65-byte ark-ec encoding the transcript absorbs
inline void point65(const uint8_t* p64, uint8_t* out65) {
  uint64_t w[8];
  memcpy(w, p64, 64);
  if ((w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0) {
    memset(out65, 0, 65);  // identity = (0, 1, infinity)
    out65[32] = 1;
    out65[64] = 1;
  } else {
    memcpy(out65, p64, 64);
    out65[64] = 0;
  }
}

inline fr cp_challenge(Transcript seeded, const uint8_t* g, const uint8_t* h, const uint8_t* s0, const uint8_t* s1,
                       const uint8_t* a, const uint8_t* b) {
  // one 404-byte message, fed in one piece: the long-input Blake2s path takes six of its seven blocks
  uint8_t msg[14 + 6 * 65];
  memcpy(msg, "chaum_pedersen", 14);
  const uint8_t* pts[6] = {g, h, s0, s1, a, b};
  for (int k = 0; k < 6; k++) point65(pts[k], msg + 14 + 65 * k);
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

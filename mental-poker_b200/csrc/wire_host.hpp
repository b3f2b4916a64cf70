// Host-only half of the wire format: ark-serialize 0.3
// compressed encodings of the objects the reference bounds by CanonicalSerialize / CanonicalDeserialize
// (reference src/lib.rs:45-71; proof sizes measured with `serialized_size`,
// examples/parameter_selection.rs:95).  Restated from recall [UPSTREAM-RECALL], as in oracle/py/wire.py:
//   compressed SW affine = x (32 B LE) | flags in the top bits of the last byte
//                          (bit 7: y is the larger of (y, -y); bit 6: infinity)
//   Vec<T> = u64 LE length | items;   ciphertext = c1 | c2
// Compression is byte handling (one 256-bit comparison per point) and stays on the host; decompression
// needs a square root in F_p per point and runs on the GPU (wire.cu).  No CUDA in this header:
// tests/host/host_shim.cpp compiles it with g++.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace mp {

inline constexpr uint8_t kWireFlagLarger = 0x80, kWireFlagInfinity = 0x40;

// canonical little-endian y > (p - 1) / 2 ?   (p - 1) / 2 = 2^250 + 17 * 2^191
inline bool wire_y_is_larger(const uint8_t* y32) {
  static const uint32_t half[8] = {0, 0, 0, 0, 0, 0x80000000u, 0x00000008u, 0x04000000u};
  uint32_t w[8];
  memcpy(w, y32, 32);
  for (int i = 7; i >= 0; i--)
    if (w[i] != half[i]) return w[i] > half[i];
  return false;
}

// 64-byte x || y (all-zero = identity)  ->  32-byte compressed
inline void wire_compress_point(const uint8_t* p64, uint8_t* out32) {
  bool zero = true;
  for (int i = 0; i < 64; i++) zero &= p64[i] == 0;
  if (zero) {
    memset(out32, 0, 32);
    out32[31] = kWireFlagInfinity;
    return;
  }
  memcpy(out32, p64, 32);
  if (wire_y_is_larger(p64 + 32)) out32[31] |= kWireFlagLarger;
}

// The flat proof of include/mpshuffle.h as runs of points (true) / scalars (false)
struct WireRun { bool points; size_t count; };
inline std::vector<WireRun> wire_proof_runs(int m, int n) {
  return {{true, 5 * (size_t)m + 4}, {false, 2 * (size_t)n + 3}, {true, 3}, {false, 2 * (size_t)n + 2},
          {true, 6 * (size_t)m + 1}, {false, (size_t)n + 4}};
}
inline uint64_t wire_proof_len(int m, int n) { return (uint64_t)(11 * (size_t)m + 8) * 32 + (uint64_t)(5 * (size_t)n + 9) * 32; }
inline uint64_t wire_deck_len(uint64_t n_cards) { return 8 + 64 * n_cards; }

inline void wire_proof_serialize(int m, int n, const uint8_t* proof, uint8_t* out) {
  for (const WireRun& r : wire_proof_runs(m, n)) {
    if (r.points) {
      for (size_t i = 0; i < r.count; i++, proof += 64, out += 32) wire_compress_point(proof, out);
    } else {
      memcpy(out, proof, 32 * r.count);
      proof += 32 * r.count;
      out += 32 * r.count;
    }
  }
}

}  // namespace mp

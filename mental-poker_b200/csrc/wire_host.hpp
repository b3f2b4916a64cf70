// Host-only half of the wire format: ark-serialize 0.3
// compressed encodings of the objects the reference bounds by CanonicalSerialize / CanonicalDeserialize
// (reference src/lib.rs:45-71; proof sizes measured with `serialized_size`,
// examples/parameter_selection.rs:95).  Restated from recall [UPSTREAM-RECALL], as in oracle/py/wire.py:
//   compressed SW affine = x (32 B LE; 48 B on BLS12-377) | flags in the top bits of the last byte
//                          (bit 7: y is the larger of (y, -y); bit 6: infinity)
//   Vec<T> = u64 LE length | items;   ciphertext = c1 | c2
// Compression is byte handling (one 256-bit comparison per point) and stays on the host; decompression
// needs a square root in F_p per point and runs on the GPU (wire.cu).  No CUDA in this header:
// tests/host/host_shim.cpp compiles it with g++.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "fq.cuh"

namespace mp {

inline constexpr uint8_t kWireFlagLarger = 0x80, kWireFlagInfinity = 0x40;
// widths of the curve fq.cuh selects: coordinate = compressed point = 32 bytes (Stark) / 48 bytes (BLS12-377, whose
// 377-bit prime leaves the top 7 bits of the last byte free), uncompressed C-ABI point = twice that
inline constexpr size_t kWireFe = 4 * (size_t)kFqLimbs, kWirePt = 2 * kWireFe;

// canonical little-endian y > (p - 1) / 2 ?
inline bool wire_y_is_larger(const uint8_t* y) {
#ifdef MP_CURVE_BLS12_377
  static const uint32_t half[kFqLimbs] = {0x00000000u, 0x42846000u, 0x18000000u, 0x0b85aea2u, 0xdd04a400u, 0x8f79b117u,
                                          0x807a89c7u, 0x8d116cf9u, 0x3650a49du, 0x631d82e0u, 0x0be28875u, 0x00d71d23u};
#else
  static const uint32_t half[kFqLimbs] = {0, 0, 0, 0, 0, 0x80000000u, 0x00000008u, 0x04000000u};  // 2^250 + 17 * 2^191
#endif
  uint32_t w[kFqLimbs];
  memcpy(w, y, kWireFe);
  for (int i = kFqLimbs - 1; i >= 0; i--)
    if (w[i] != half[i]) return w[i] > half[i];
  return false;
}

// x || y (all-zero = identity)  ->  compressed x with flags
inline void wire_compress_point(const uint8_t* p, uint8_t* out) {
  bool zero = true;
  for (size_t i = 0; i < kWirePt; i++) zero &= p[i] == 0;
  if (zero) {
    memset(out, 0, kWireFe);
    out[kWireFe - 1] = kWireFlagInfinity;
    return;
  }
  memcpy(out, p, kWireFe);
  if (wire_y_is_larger(p + kWireFe)) out[kWireFe - 1] |= kWireFlagLarger;
}

// The flat proof of include/mpshuffle.h as runs of points (true) / scalars (false)
struct WireRun { bool points; size_t count; };
inline std::vector<WireRun> wire_proof_runs(int m, int n) {
  return {{true, 5 * (size_t)m + 4}, {false, 2 * (size_t)n + 3}, {true, 3}, {false, 2 * (size_t)n + 2},
          {true, 6 * (size_t)m + 1}, {false, (size_t)n + 4}};
}
inline uint64_t wire_proof_len(int m, int n) { return (uint64_t)(11 * (size_t)m + 8) * kWireFe + (uint64_t)(5 * (size_t)n + 9) * 32; }
inline uint64_t wire_deck_len(uint64_t n_cards) { return 8 + 2 * kWireFe * n_cards; }

inline void wire_proof_serialize(int m, int n, const uint8_t* proof, uint8_t* out) {
  for (const WireRun& r : wire_proof_runs(m, n)) {
    if (r.points) {
      for (size_t i = 0; i < r.count; i++, proof += kWirePt, out += kWireFe) wire_compress_point(proof, out);
    } else {
      memcpy(out, proof, 32 * r.count);
      proof += 32 * r.count;
      out += 32 * r.count;
    }
  }
}

}  // namespace mp

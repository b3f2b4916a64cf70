// Host-side Fiat-Shamir transcript of the shuffle argument.  By design (BASELINE north_star)
// the transcript and protocol state stay on the host; the GPU only sees the challenges.
//
// Mirrors ark-marlin 0.3 `FiatShamirRng<Blake2s>` as the reference instantiates it
// (reference src/discrete_log_cards/mod.rs:9,12 and the seed at mod.rs:84,408,436):
//   from_seed(bytes):  seed = Blake2s-256(bytes);            rng = ChaCha20(seed)
//   absorb(bytes):     seed = Blake2s-256(bytes || seed);    rng = ChaCha20(seed)   (stream restarts)
//   challenge:         ark-ff 0.3 `Fp256::rand`: 4 x next_u64 taken as the raw MONTGOMERY
//                      representation, top 4 bits cleared, rejection-sampled below the modulus
// (SURVEY.md A1, A4, A5).  Point encoding inside absorbed data is ark-ec 0.3 `ToBytes`:
// x || y || infinity-flag = 65 bytes, identity = (0, 1, true).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#if (defined(__x86_64__) || defined(__i386__)) && defined(__GNUC__) && !defined(__CUDA_ARCH__) && \
    !defined(MP_BLAKE2S_FORCE_SCALAR)
#include <immintrin.h>
#define MP_BLAKE2S_X86 1
#endif

#include "fr.cuh"

namespace mp {

class Blake2s {
 public:
  Blake2s() { reset(); }
  void reset() {
    static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    memcpy(h_, iv, 32);
    h_[0] ^= 0x01010020u;  // 32-byte digest, unkeyed, sequential mode
    t_ = 0;
    fill_ = 0;
  }
  void update(const void* data, size_t len) {
    const uint8_t* in = static_cast<const uint8_t*>(data);
    if (len == 0) return;
    if (fill_ > 0) {
      size_t take = 64 - fill_;
      if (take > len) take = len;
      memcpy(buf_ + fill_, in, take);
      fill_ += take;
      in += take;
      len -= take;
      if (len == 0) return;  // keep a possibly-final block buffered
      t_ += 64;
      compress(buf_, false);
      fill_ = 0;
    }
#ifdef MP_BLAKE2S_X86
    if (len > 64 && use_avx512()) {  // all full blocks but the last in one call: the state stays in registers
      const size_t nblocks = (len - 1) / 64;
      compress_blocks_avx512(in, nblocks);
      in += 64 * nblocks;
      len -= 64 * nblocks;
    }
#endif
    while (len > 64) {  // strictly greater: the last block must go through finish()
      t_ += 64;
      compress(in, false);
      in += 64;
      len -= 64;
    }
    memcpy(buf_, in, len);
    fill_ = len;
  }
  void finish(uint8_t out[32]) {
    t_ += fill_;
    memset(buf_ + fill_, 0, 64 - fill_);
    compress(buf_, true);
    memcpy(out, h_, 32);
  }

 private:
  friend class Blake2sLanes;  // the multi-stream form below reads and writes (h, t, buffered tail)
  static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  // The statement absorb of a 2^16-card proof hashes ~17 MB on ONE core and nothing on the GPU can
  // start before it ends, so the compression function matters: on x86 hosts with AVX2 a 4-lane
  // row formulation (one state row per XMM register, rotations by 16 / 8 as byte shuffles, VEX
  // three-operand forms) runs ~1.6x faster than the scalar code; chosen once at run time.
  void compress(const uint8_t* block, bool last) {
#ifdef MP_BLAKE2S_X86
    static const bool use_avx = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("ssse3");
    if (use_avx) {
      compress_avx(block, last);
      return;
    }
#endif
    compress_scalar(block, last);
  }
#ifdef MP_BLAKE2S_X86
  static bool use_avx512() {
#ifdef MP_BLAKE2S_NO_AVX512
    return false;
#else
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") &&
                           getenv("MP_BLAKE2S_NO_AVX512") == nullptr;
    return ok;
#endif
  }
  // AVX-512VL form for long inputs.  The compression function is one dependent chain
  // (a -> d -> c -> b twice per G layer, 20 layers per block), so throughput is set by the latency of
  // that chain and nothing else; this version takes everything else off it:
  //   * rotations are single VPRORD instructions (AVX2 needs shift + shift + or for 12 and 7);
  //   * the whole message block sits in one ZMM register and each round's 16 words come out of ONE
  //     VPERMD with a per-round index vector (no scalar gathers);
  //   * the message word is added into row a BEFORE row b is ready (the asm barrier stops the compiler
  //     from re-associating a + m + b into (b + m) + a);
  //   * diagonalisation rotates rows a, c, d and leaves b -- the value computed last -- in place; the
  //     per-round index vectors are rotated to match.
  // 6 dependent single-cycle instructions per G layer = 240 cycles per 64-byte block.
  __attribute__((target("avx512f,avx512vl"))) void compress_blocks_avx512(const uint8_t* in, size_t nblocks) {
    alignas(64) static const uint32_t perm[10][16] = {
#define MP_R(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  {s0, s2, s4, s6, s1, s3, s5, s7, s14, s8, s10, s12, s15, s9, s11, s13},
        MP_R(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15) MP_R(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
        MP_R(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4) MP_R(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
        MP_R(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13) MP_R(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
        MP_R(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11) MP_R(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
        MP_R(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5) MP_R(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
#undef MP_R
    };
    const __m128i iv_lo = _mm_setr_epi32((int)0x6A09E667u, (int)0xBB67AE85u, (int)0x3C6EF372u, (int)0xA54FF53Au);
    const __m128i iv_hi = _mm_setr_epi32((int)0x510E527Fu, (int)0x9B05688Cu, (int)0x1F83D9ABu, (int)0x5BE0CD19u);
    __m128i h_lo = _mm_loadu_si128((const __m128i*)&h_[0]), h_hi = _mm_loadu_si128((const __m128i*)&h_[4]);
    uint64_t t = t_;
#define MP_PIN(x) __asm__("" : "+x"(x))
#define MP_G(ra, rb)                                          \
  row1 = _mm_add_epi32(row1, row2);                           \
  row4 = _mm_ror_epi32(_mm_xor_si128(row4, row1), ra);        \
  row3 = _mm_add_epi32(row3, row4);                           \
  row2 = _mm_ror_epi32(_mm_xor_si128(row2, row3), rb);
    for (size_t blk = 0; blk < nblocks; blk++, in += 64) {
      t += 64;
      const __m512i M = _mm512_loadu_si512((const void*)in);
      __m128i row1 = h_lo, row2 = h_hi, row3 = iv_lo;
      __m128i row4 = _mm_xor_si128(iv_hi, _mm_setr_epi32((int)(uint32_t)t, (int)(uint32_t)(t >> 32), 0, 0));
      __m512i P = _mm512_permutexvar_epi32(_mm512_load_si512((const void*)perm[0]), M);
      row1 = _mm_add_epi32(row1, _mm512_castsi512_si128(P));
      MP_PIN(row1);
#pragma GCC unroll 10
      for (int r = 0; r < 10; r++) {
        MP_G(16, 12)
        row1 = _mm_add_epi32(row1, _mm512_extracti32x4_epi32(P, 1));
        MP_PIN(row1);
        MP_G(8, 7)
        row1 = _mm_shuffle_epi32(row1, _MM_SHUFFLE(2, 1, 0, 3));
        row3 = _mm_shuffle_epi32(row3, _MM_SHUFFLE(0, 3, 2, 1));
        row4 = _mm_shuffle_epi32(row4, _MM_SHUFFLE(1, 0, 3, 2));
        row1 = _mm_add_epi32(row1, _mm512_extracti32x4_epi32(P, 2));
        MP_PIN(row1);
        MP_G(16, 12)
        row1 = _mm_add_epi32(row1, _mm512_extracti32x4_epi32(P, 3));
        MP_PIN(row1);
        MP_G(8, 7)
        row1 = _mm_shuffle_epi32(row1, _MM_SHUFFLE(0, 3, 2, 1));
        row3 = _mm_shuffle_epi32(row3, _MM_SHUFFLE(2, 1, 0, 3));
        row4 = _mm_shuffle_epi32(row4, _MM_SHUFFLE(1, 0, 3, 2));
        if (r < 9) {
          P = _mm512_permutexvar_epi32(_mm512_load_si512((const void*)perm[r + 1]), M);
          row1 = _mm_add_epi32(row1, _mm512_castsi512_si128(P));
          MP_PIN(row1);
        }
      }
      h_lo = _mm_xor_si128(h_lo, _mm_xor_si128(row1, row3));
      h_hi = _mm_xor_si128(h_hi, _mm_xor_si128(row2, row4));
    }
#undef MP_G
#undef MP_PIN
    _mm_storeu_si128((__m128i*)&h_[0], h_lo);
    _mm_storeu_si128((__m128i*)&h_[4], h_hi);
    t_ = t;
  }
  __attribute__((target("avx2,ssse3"))) void compress_avx(const uint8_t* block, bool last) {
    const __m128i r16 = _mm_setr_epi8(2, 3, 0, 1, 6, 7, 4, 5, 10, 11, 8, 9, 14, 15, 12, 13);
    const __m128i r8 = _mm_setr_epi8(1, 2, 3, 0, 5, 6, 7, 4, 9, 10, 11, 8, 13, 14, 15, 12);
    uint32_t m[16];
    memcpy(m, block, 64);
    __m128i row1 = _mm_loadu_si128((const __m128i*)&h_[0]);
    __m128i row2 = _mm_loadu_si128((const __m128i*)&h_[4]);
    __m128i row3 = _mm_setr_epi32((int)0x6A09E667u, (int)0xBB67AE85u, (int)0x3C6EF372u, (int)0xA54FF53Au);
    __m128i row4 = _mm_xor_si128(_mm_setr_epi32((int)0x510E527Fu, (int)0x9B05688Cu, (int)0x1F83D9ABu, (int)0x5BE0CD19u),
                                 _mm_setr_epi32((int)(uint32_t)t_, (int)(uint32_t)(t_ >> 32), last ? -1 : 0, 0));
    const __m128i s1 = row1, s2 = row2;
#define MP_G1(buf)                                                       \
  row1 = _mm_add_epi32(_mm_add_epi32(row1, buf), row2);                  \
  row4 = _mm_shuffle_epi8(_mm_xor_si128(row4, row1), r16);               \
  row3 = _mm_add_epi32(row3, row4);                                      \
  row2 = _mm_xor_si128(row2, row3);                                      \
  row2 = _mm_or_si128(_mm_srli_epi32(row2, 12), _mm_slli_epi32(row2, 20));
#define MP_G2(buf)                                                       \
  row1 = _mm_add_epi32(_mm_add_epi32(row1, buf), row2);                  \
  row4 = _mm_shuffle_epi8(_mm_xor_si128(row4, row1), r8);                \
  row3 = _mm_add_epi32(row3, row4);                                      \
  row2 = _mm_xor_si128(row2, row3);                                      \
  row2 = _mm_or_si128(_mm_srli_epi32(row2, 7), _mm_slli_epi32(row2, 25));
#define MP_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15)    \
  MP_G1(_mm_setr_epi32((int)m[s0], (int)m[s2], (int)m[s4], (int)m[s6]))                   \
  MP_G2(_mm_setr_epi32((int)m[s1], (int)m[s3], (int)m[s5], (int)m[s7]))                   \
  row4 = _mm_shuffle_epi32(row4, _MM_SHUFFLE(2, 1, 0, 3));                                \
  row3 = _mm_shuffle_epi32(row3, _MM_SHUFFLE(1, 0, 3, 2));                                \
  row2 = _mm_shuffle_epi32(row2, _MM_SHUFFLE(0, 3, 2, 1));                                \
  MP_G1(_mm_setr_epi32((int)m[s8], (int)m[s10], (int)m[s12], (int)m[s14]))                \
  MP_G2(_mm_setr_epi32((int)m[s9], (int)m[s11], (int)m[s13], (int)m[s15]))                \
  row4 = _mm_shuffle_epi32(row4, _MM_SHUFFLE(0, 3, 2, 1));                                \
  row3 = _mm_shuffle_epi32(row3, _MM_SHUFFLE(1, 0, 3, 2));                                \
  row2 = _mm_shuffle_epi32(row2, _MM_SHUFFLE(2, 1, 0, 3));
    MP_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    MP_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    MP_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    MP_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    MP_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    MP_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    MP_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    MP_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    MP_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    MP_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
#undef MP_ROUND
#undef MP_G1
#undef MP_G2
    _mm_storeu_si128((__m128i*)&h_[0], _mm_xor_si128(s1, _mm_xor_si128(row1, row3)));
    _mm_storeu_si128((__m128i*)&h_[4], _mm_xor_si128(s2, _mm_xor_si128(row2, row4)));
  }
#endif
  void compress_scalar(const uint8_t* block, bool last) {
    static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    uint32_t m[16], v[16];
    memcpy(m, block, 64);
    for (int i = 0; i < 8; i++) { v[i] = h_[i]; v[i + 8] = iv[i]; }
    v[12] ^= (uint32_t)t_;
    v[13] ^= (uint32_t)(t_ >> 32);
    if (last) v[14] = ~v[14];
#define MP_G(a, b, c, d, x, y)                          \
  do {                                                  \
    v[a] = v[a] + v[b] + m[x]; v[d] = rotr(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d];        v[b] = rotr(v[b] ^ v[c], 12); \
    v[a] = v[a] + v[b] + m[y]; v[d] = rotr(v[d] ^ v[a], 8);  \
    v[c] = v[c] + v[d];        v[b] = rotr(v[b] ^ v[c], 7);  \
  } while (0)
#define MP_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  MP_G(0, 4, 8, 12, s0, s1); MP_G(1, 5, 9, 13, s2, s3); MP_G(2, 6, 10, 14, s4, s5);    \
  MP_G(3, 7, 11, 15, s6, s7); MP_G(0, 5, 10, 15, s8, s9); MP_G(1, 6, 11, 12, s10, s11); \
  MP_G(2, 7, 8, 13, s12, s13); MP_G(3, 4, 9, 14, s14, s15)
    MP_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    MP_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    MP_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    MP_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    MP_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    MP_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    MP_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    MP_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    MP_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    MP_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
#undef MP_ROUND
#undef MP_G
    for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
  }
  uint32_t h_[8];
  uint64_t t_;
  uint8_t buf_[64];
  size_t fill_;
};

// Up to 8 Blake2s streams of EQUAL length hashed in lockstep: lane l of every vector register holds stream l, so one
// pass of the compression function advances all of them (AVX2; rotations are single instructions with AVX-512VL).
// A batch of 2^16-card proofs spends 2 x 23.5 ms of one core per proof on its two statement hashes (17 MB each, one
// serial Blake2s stream at 793 MB/s), and with several GPUs' workers on one host that hashing is what the GPUs wait
// for; the streams of different proofs are independent, so eight of them cost about one and a half times one.
// Same bytes in, same digests out as eight Blake2s objects: update() mirrors Blake2s::update (the last block stays
// buffered for finish), and the state of a lane can be moved into a Blake2s at any point (`extract`).
class Blake2sLanes {
 public:
  static constexpr int kLanes = 8;
  explicit Blake2sLanes(int lanes) : lanes_(lanes) {
    for (int l = 0; l < kLanes; l++) st_[l].reset();
  }
  int lanes() const { return lanes_; }
  // the same bytes for every lane (labels, parameters)
  void update_all(const void* data, size_t len) {
    const uint8_t* p[kLanes];
    for (int l = 0; l < kLanes; l++) p[l] = static_cast<const uint8_t*>(data);
    update(p, len);
  }
  // data[l] = `len` bytes for lane l (l < lanes())
  void update(const uint8_t* const* data, size_t len) {
    if (len == 0) return;
    const uint8_t* in[kLanes];
    for (int l = 0; l < kLanes; l++) in[l] = data[l < lanes_ ? l : 0];
    size_t fill = st_[0].fill_;
    if (fill > 0) {
      size_t take = 64 - fill;
      if (take > len) take = len;
      for (int l = 0; l < lanes_; l++) {
        memcpy(st_[l].buf_ + fill, in[l], take);
        st_[l].fill_ = fill + take;
        in[l] += take;
      }
      len -= take;
      if (len == 0) return;  // keep a possibly-final block buffered
      const uint8_t* b[kLanes];
      for (int l = 0; l < kLanes; l++) b[l] = st_[l < lanes_ ? l : 0].buf_;
      compress_blocks(b, 1);
      for (int l = 0; l < lanes_; l++) st_[l].fill_ = 0;
    }
    if (len > 64) {  // all full blocks but the last
      const size_t nblocks = (len - 1) / 64;
      compress_blocks(in, nblocks);
      for (int l = 0; l < kLanes; l++) in[l] += 64 * nblocks;
      len -= 64 * nblocks;
    }
    for (int l = 0; l < lanes_; l++) {
      memcpy(st_[l].buf_, in[l], len);
      st_[l].fill_ = len;
    }
  }
  // lane l's stream so far, as a single-stream hasher that can go on absorbing
  void extract(int l, Blake2s* out) const { *out = st_[l]; }
  static bool vectorised() {
#ifdef MP_BLAKE2S_X86
    static const bool ok = __builtin_cpu_supports("avx2") && !getenv("MP_BLAKE2S_LANES_SCALAR");
    return ok;
#else
    return false;
#endif
  }

 private:
  // nblocks full, non-final blocks from in[l] for every lane
  void compress_blocks(const uint8_t* const* in, size_t nblocks) {
#ifdef MP_BLAKE2S_X86
    if (vectorised()) {
      if (Blake2s::use_avx512()) compress_blocks_avx512vl(in, nblocks);
      else compress_blocks_avx2(in, nblocks);
      return;
    }
#endif
    for (int l = 0; l < lanes_; l++) {
      const uint8_t* p = in[l];
      for (size_t b = 0; b < nblocks; b++, p += 64) {
        st_[l].t_ += 64;
        st_[l].compress(p, false);
      }
    }
  }
#ifdef MP_BLAKE2S_X86
#define MP_LANES_BODY(ROR)                                                                                         \
    const __m256i iv0 = _mm256_set1_epi32((int)0x6A09E667u), iv1 = _mm256_set1_epi32((int)0xBB67AE85u),              \
                  iv2 = _mm256_set1_epi32((int)0x3C6EF372u), iv3 = _mm256_set1_epi32((int)0xA54FF53Au),              \
                  iv4 = _mm256_set1_epi32((int)0x510E527Fu), iv5 = _mm256_set1_epi32((int)0x9B05688Cu),              \
                  iv6 = _mm256_set1_epi32((int)0x1F83D9ABu), iv7 = _mm256_set1_epi32((int)0x5BE0CD19u);              \
    __m256i h[8];                                                                                                  \
    {                                                                                                              \
      alignas(32) uint32_t tmp[8][8];                                                                              \
      for (int w = 0; w < 8; w++)                                                                                  \
        for (int l = 0; l < 8; l++) tmp[w][l] = st_[l].h_[w];                                                      \
      for (int w = 0; w < 8; w++) h[w] = _mm256_load_si256((const __m256i*)tmp[w]);                                \
    }                                                                                                              \
    uint64_t t = st_[0].t_;                                                                                        \
    for (size_t blk = 0; blk < nblocks; blk++) {                                                                   \
      t += 64;                                                                                                     \
      __m256i m[16];                                                                                               \
      for (int half = 0; half < 2; half++) {                                                                       \
        __m256i r[8];                                                                                              \
        for (int l = 0; l < 8; l++) r[l] = _mm256_loadu_si256((const __m256i*)(in[l] + 64 * blk + 32 * half));     \
        const __m256i t0 = _mm256_unpacklo_epi32(r[0], r[1]), t1 = _mm256_unpackhi_epi32(r[0], r[1]);              \
        const __m256i t2 = _mm256_unpacklo_epi32(r[2], r[3]), t3 = _mm256_unpackhi_epi32(r[2], r[3]);              \
        const __m256i t4 = _mm256_unpacklo_epi32(r[4], r[5]), t5 = _mm256_unpackhi_epi32(r[4], r[5]);              \
        const __m256i t6 = _mm256_unpacklo_epi32(r[6], r[7]), t7 = _mm256_unpackhi_epi32(r[6], r[7]);              \
        const __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);                      \
        const __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);                      \
        const __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);                      \
        const __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);                      \
        m[8 * half + 0] = _mm256_permute2x128_si256(u0, u4, 0x20);                                                 \
        m[8 * half + 1] = _mm256_permute2x128_si256(u1, u5, 0x20);                                                 \
        m[8 * half + 2] = _mm256_permute2x128_si256(u2, u6, 0x20);                                                 \
        m[8 * half + 3] = _mm256_permute2x128_si256(u3, u7, 0x20);                                                 \
        m[8 * half + 4] = _mm256_permute2x128_si256(u0, u4, 0x31);                                                 \
        m[8 * half + 5] = _mm256_permute2x128_si256(u1, u5, 0x31);                                                 \
        m[8 * half + 6] = _mm256_permute2x128_si256(u2, u6, 0x31);                                                 \
        m[8 * half + 7] = _mm256_permute2x128_si256(u3, u7, 0x31);                                                 \
      }                                                                                                            \
      __m256i v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];              \
      __m256i v8 = iv0, v9 = iv1, v10 = iv2, v11 = iv3;                                                            \
      __m256i v12 = _mm256_xor_si256(iv4, _mm256_set1_epi32((int)(uint32_t)t));                                    \
      __m256i v13 = _mm256_xor_si256(iv5, _mm256_set1_epi32((int)(uint32_t)(t >> 32)));                            \
      __m256i v14 = iv6, v15 = iv7;                                                                                \
      MP_LR(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);                                                 \
      MP_LR(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);                                                 \
      MP_LR(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);                                                 \
      MP_LR(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);                                                 \
      MP_LR(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);                                                 \
      MP_LR(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);                                                 \
      MP_LR(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);                                                 \
      MP_LR(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);                                                 \
      MP_LR(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);                                                 \
      MP_LR(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);                                                 \
      h[0] = _mm256_xor_si256(h[0], _mm256_xor_si256(v0, v8));                                                     \
      h[1] = _mm256_xor_si256(h[1], _mm256_xor_si256(v1, v9));                                                     \
      h[2] = _mm256_xor_si256(h[2], _mm256_xor_si256(v2, v10));                                                    \
      h[3] = _mm256_xor_si256(h[3], _mm256_xor_si256(v3, v11));                                                    \
      h[4] = _mm256_xor_si256(h[4], _mm256_xor_si256(v4, v12));                                                    \
      h[5] = _mm256_xor_si256(h[5], _mm256_xor_si256(v5, v13));                                                    \
      h[6] = _mm256_xor_si256(h[6], _mm256_xor_si256(v6, v14));                                                    \
      h[7] = _mm256_xor_si256(h[7], _mm256_xor_si256(v7, v15));                                                    \
    }                                                                                                              \
    {                                                                                                              \
      alignas(32) uint32_t tmp[8][8];                                                                              \
      for (int w = 0; w < 8; w++) _mm256_store_si256((__m256i*)tmp[w], h[w]);                                      \
      for (int l = 0; l < 8; l++) {                                                                                \
        for (int w = 0; w < 8; w++) st_[l].h_[w] = tmp[w][l];                                                      \
        st_[l].t_ = t;                                                                                             \
      }                                                                                                            \
    }
#define MP_LG(a, b, c, d, x, y)                                                         \
  a = _mm256_add_epi32(_mm256_add_epi32(a, b), m[x]); d = ROR(_mm256_xor_si256(d, a), 16); \
  c = _mm256_add_epi32(c, d);                         b = ROR(_mm256_xor_si256(b, c), 12); \
  a = _mm256_add_epi32(_mm256_add_epi32(a, b), m[y]); d = ROR(_mm256_xor_si256(d, a), 8);  \
  c = _mm256_add_epi32(c, d);                         b = ROR(_mm256_xor_si256(b, c), 7)
#define MP_LR(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15)                      \
  MP_LG(v0, v4, v8, v12, s0, s1); MP_LG(v1, v5, v9, v13, s2, s3); MP_LG(v2, v6, v10, v14, s4, s5);       \
  MP_LG(v3, v7, v11, v15, s6, s7); MP_LG(v0, v5, v10, v15, s8, s9); MP_LG(v1, v6, v11, v12, s10, s11);  \
  MP_LG(v2, v7, v8, v13, s12, s13); MP_LG(v3, v4, v9, v14, s14, s15)
  __attribute__((target("avx2"))) void compress_blocks_avx2(const uint8_t* const* in, size_t nblocks) {
#define ROR(x, n) _mm256_or_si256(_mm256_srli_epi32((x), (n)), _mm256_slli_epi32((x), 32 - (n)))
    MP_LANES_BODY(ROR)
#undef ROR
  }
  __attribute__((target("avx2,avx512f,avx512vl"))) void compress_blocks_avx512vl(const uint8_t* const* in, size_t nblocks) {
#define ROR(x, n) _mm256_ror_epi32((x), (n))
    MP_LANES_BODY(ROR)
#undef ROR
  }
#undef MP_LR
#undef MP_LG
#undef MP_LANES_BODY
#endif
  int lanes_;
  Blake2s st_[kLanes];  // per lane: h, t, buffered tail (unused lanes shadow lane 0 and are never read back)
};

// rand_chacha `ChaCha20Rng`: 20 rounds, 64-bit block counter in words 12..13, stream id 0
class ChaCha20Stream {
 public:
  void seed(const uint8_t key[32]) {
    memcpy(key_, key, 32);
    counter_ = 0;
    pos_ = 16;
  }
  uint32_t next_u32() {
    if (pos_ >= 16) refill();
    return buf_[pos_++];
  }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
  }

 private:
  static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  void refill() {
    uint32_t s[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u};
    memcpy(s + 4, key_, 32);
    s[12] = (uint32_t)counter_;
    s[13] = (uint32_t)(counter_ >> 32);
    s[14] = 0;
    s[15] = 0;
    uint32_t x[16];
    memcpy(x, s, 64);
    auto qr = [&](int a, int b, int c, int d) {
      x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
      x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
      x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
      x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int i = 0; i < 10; i++) {
      qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
      qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) buf_[i] = x[i] + s[i];
    counter_++;
    pos_ = 0;
  }
  uint32_t key_[8];
  uint64_t counter_;
  uint32_t buf_[16];
  int pos_;
};

class Transcript {
 public:
  // FiatShamirRng::<Blake2s>::from_seed(&to_bytes![SHUFFLE_RNG_SEED])   mod.rs:84,408,436
  Transcript() : Transcript("Shuffle Proof", 13) {}
  // from_seed for the other proof types (mod.rs:80-83: "Key Ownership Proof" || info, "Masking Proof", ...)
  Transcript(const void* seed, size_t len) {
    Blake2s h;
    h.update(seed, len);
    h.finish(seed_);
    rng_.seed(seed_);
  }
  // absorb = begin(); feed()...; end()
  void begin() { pending_.reset(); }
  void feed(const void* data, size_t len) { pending_.update(data, len); }
  void feed_label(const char* s) { pending_.update(s, strlen(s)); }
  // C-ABI points (x || y, kPointBytes = 64 on the Stark curve / 96 on BLS12-377; all-zero = identity) -> the
  // ark-ec `ToBytes` encoding x || y || infinity flag, serialized in blocks of 64 points so the hash sees
  // few, large updates.  (The name keeps its Stark-curve origin.)
  void feed_points64(const uint8_t* pts, size_t count) {
    uint8_t buf[kEncChunk * kEncPoint];
    while (count > 0) {
      const size_t take = count < kEncChunk ? count : kEncChunk;
      encode_points(pts, take, buf);
      pending_.update(buf, take * kEncPoint);
      pts += kEncIn * take;
      count -= take;
    }
  }
  // `take` C-ABI points -> their ark-ec `ToBytes` encodings, back to back
  static constexpr size_t kEncIn = 8 * kFqLimbs, kEncPoint = kEncIn + 1, kEncChunk = 64;
  static void encode_points(const uint8_t* pts, size_t take, uint8_t* b) {
    constexpr size_t PB = kEncIn, FB = PB / 2;
    for (size_t i = 0; i < take; i++, b += PB + 1) {
      const uint8_t* p = pts + PB * i;
      uint64_t w[PB / 8], any = 0;
      memcpy(w, p, PB);
      for (size_t k = 0; k < PB / 8; k++) any |= w[k];
      if (any == 0) {
        memset(b, 0, PB + 1);  // identity = (0, 1, infinity)
        b[FB] = 1;
        b[PB] = 1;
      } else {
        memcpy(b, p, PB);
        b[PB] = 0;
      }
    }
  }
  // continue an absorb whose first bytes were hashed elsewhere (TranscriptLanes below): replaces begin()
  void adopt_pending(const Blake2s& started) { pending_ = started; }
  void end() {
    pending_.update(seed_, 32);
    pending_.finish(seed_);
    rng_.seed(seed_);
  }
  fr challenge() {
    for (;;) {
      uint64_t l[4];
      for (int i = 0; i < 4; i++) l[i] = rng_.next_u64();
      l[3] &= 0xFFFFFFFFFFFFFFFFull >> kFrShaveBits;
      fr c;
      for (int i = 0; i < 4; i++) {
        c.v[2 * i] = (uint32_t)l[i];
        c.v[2 * i + 1] = (uint32_t)(l[i] >> 32);
      }
      // accept iff raw < n
      bool lt = false;
      for (int i = 7; i >= 0; i--) {
        uint32_t mi = fr_modulus_limb(i);
        if (c.v[i] != mi) { lt = c.v[i] < mi; break; }
      }
      if (lt) return c;  // raw value IS the Montgomery representation
    }
  }

 private:
  uint8_t seed_[32];
  ChaCha20Stream rng_;
  Blake2s pending_;
};

// The first part of ONE absorb of up to 8 transcripts whose inputs have the same shape (the statements of the proofs of
// a batch: same parameters, decks of the same size), hashed in lockstep by Blake2sLanes.  hand_over(l, fs) moves lane
// l's hasher into `fs`, which goes on with feed() / feed_points64() / end() as if it had absorbed the bytes itself.
class TranscriptLanes {
 public:
  explicit TranscriptLanes(int lanes) : h_(lanes) {}
  void feed_all(const void* data, size_t len) { h_.update_all(data, len); }
  void feed_label_all(const char* s) { h_.update_all(s, strlen(s)); }
  void feed_points64_all(const uint8_t* pts, size_t count) {
    const uint8_t* p[Blake2sLanes::kLanes];
    for (int l = 0; l < Blake2sLanes::kLanes; l++) p[l] = pts;
    feed_points64(p, count);
  }
  // pts[l] = `count` C-ABI points of lane l
  void feed_points64(const uint8_t* const* pts, size_t count) {
    const int L = h_.lanes();
    const uint8_t* src[Blake2sLanes::kLanes];
    for (int l = 0; l < L; l++) src[l] = pts[l];
    const bool shared = [&] { for (int l = 1; l < L; l++) if (pts[l] != pts[0]) return false; return true; }();
    const uint8_t* enc[Blake2sLanes::kLanes];
    while (count > 0) {
      const size_t take = count < Transcript::kEncChunk ? count : Transcript::kEncChunk;
      for (int l = 0; l < L; l++) {
        if (l == 0 || !shared) Transcript::encode_points(src[l], take, buf_[l]);
        enc[l] = shared ? buf_[0] : buf_[l];
        src[l] += Transcript::kEncIn * take;
      }
      h_.update(enc, take * Transcript::kEncPoint);
      count -= take;
    }
  }
  void extract(int lane, Blake2s* out) const { h_.extract(lane, out); }
  void hand_over(int lane, Transcript* fs) const {
    Blake2s b;
    h_.extract(lane, &b);
    fs->adopt_pending(b);
  }

 private:
  Blake2sLanes h_;
  uint8_t buf_[Blake2sLanes::kLanes][Transcript::kEncChunk * Transcript::kEncPoint];
};

}  // namespace mp

// Host Blake2s throughput probe (run on the GPU box's CPU): scalar (product code) vs SSSE3 rows.
#include "../../mental-poker_b200/csrc/transcript.hpp"
#include <chrono>
#include <vector>
#include <stdio.h>
#if defined(__SSSE3__)
#include <tmmintrin.h>
static void compress_sse(uint32_t h[8], uint64_t t, const uint8_t* block, bool last) {
  const __m128i r16 = _mm_setr_epi8(2, 3, 0, 1, 6, 7, 4, 5, 10, 11, 8, 9, 14, 15, 12, 13);
  const __m128i r8 = _mm_setr_epi8(1, 2, 3, 0, 5, 6, 7, 4, 9, 10, 11, 8, 13, 14, 15, 12);
  uint32_t m[16];
  memcpy(m, block, 64);
  __m128i row1 = _mm_loadu_si128((const __m128i*)&h[0]), row2 = _mm_loadu_si128((const __m128i*)&h[4]);
  __m128i row3 = _mm_setr_epi32((int)0x6A09E667u, (int)0xBB67AE85u, (int)0x3C6EF372u, (int)0xA54FF53Au);
  __m128i row4 = _mm_xor_si128(_mm_setr_epi32((int)0x510E527Fu, (int)0x9B05688Cu, (int)0x1F83D9ABu, (int)0x5BE0CD19u),
                               _mm_setr_epi32((int)(uint32_t)t, (int)(uint32_t)(t >> 32), last ? -1 : 0, 0));
  const __m128i s1 = row1, s2 = row2;
#define G1(buf) row1 = _mm_add_epi32(_mm_add_epi32(row1, buf), row2); row4 = _mm_shuffle_epi8(_mm_xor_si128(row4, row1), r16); \
  row3 = _mm_add_epi32(row3, row4); row2 = _mm_xor_si128(row2, row3); row2 = _mm_or_si128(_mm_srli_epi32(row2, 12), _mm_slli_epi32(row2, 20));
#define G2(buf) row1 = _mm_add_epi32(_mm_add_epi32(row1, buf), row2); row4 = _mm_shuffle_epi8(_mm_xor_si128(row4, row1), r8); \
  row3 = _mm_add_epi32(row3, row4); row2 = _mm_xor_si128(row2, row3); row2 = _mm_or_si128(_mm_srli_epi32(row2, 7), _mm_slli_epi32(row2, 25));
#define RND(s0,s1,s2,s3,s4,s5,s6,s7,s8,s9,s10,s11,s12,s13,s14,s15) \
  G1(_mm_setr_epi32((int)m[s0],(int)m[s2],(int)m[s4],(int)m[s6])) G2(_mm_setr_epi32((int)m[s1],(int)m[s3],(int)m[s5],(int)m[s7])) \
  row4 = _mm_shuffle_epi32(row4, _MM_SHUFFLE(2,1,0,3)); row3 = _mm_shuffle_epi32(row3, _MM_SHUFFLE(1,0,3,2)); row2 = _mm_shuffle_epi32(row2, _MM_SHUFFLE(0,3,2,1)); \
  G1(_mm_setr_epi32((int)m[s8],(int)m[s10],(int)m[s12],(int)m[s14])) G2(_mm_setr_epi32((int)m[s9],(int)m[s11],(int)m[s13],(int)m[s15])) \
  row4 = _mm_shuffle_epi32(row4, _MM_SHUFFLE(0,3,2,1)); row3 = _mm_shuffle_epi32(row3, _MM_SHUFFLE(1,0,3,2)); row2 = _mm_shuffle_epi32(row2, _MM_SHUFFLE(2,1,0,3));
  RND(0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15) RND(14,10,4,8,9,15,13,6,1,12,0,2,11,7,5,3) RND(11,8,12,0,5,2,15,13,10,14,3,6,7,1,9,4)
  RND(7,9,3,1,13,12,11,14,2,6,5,10,4,0,15,8) RND(9,0,5,7,2,4,10,15,14,1,11,12,6,8,3,13) RND(2,12,6,10,0,11,8,3,4,13,7,5,15,14,1,9)
  RND(12,5,1,15,14,13,4,10,0,7,6,3,9,2,8,11) RND(13,11,7,14,12,1,3,9,5,0,15,4,8,6,2,10) RND(6,15,14,9,11,3,0,8,12,2,13,7,1,4,10,5)
  RND(10,2,8,4,7,6,1,5,15,11,9,14,3,12,13,0)
  _mm_storeu_si128((__m128i*)&h[0], _mm_xor_si128(s1, _mm_xor_si128(row1, row3)));
  _mm_storeu_si128((__m128i*)&h[4], _mm_xor_si128(s2, _mm_xor_si128(row2, row4)));
}
#endif
int main() {
  size_t bytes = 17 << 20;
  std::vector<uint8_t> buf(bytes, 7);
  for (int rep = 0; rep < 3; rep++) {
    auto t0 = std::chrono::steady_clock::now();
    mp::Blake2s h; h.update(buf.data(), buf.size()); uint8_t out[32]; h.finish(out);
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    printf("scalar: %.2f ms (%.0f MB/s) %02x", ms, bytes / 1e3 / ms, out[0]);
#if defined(__SSSE3__)
    uint32_t hh[8] = {0x6A09E667u ^ 0x01010020u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    t0 = std::chrono::steady_clock::now();
    uint64_t t = 0;
    for (size_t off = 0; off + 64 <= bytes; off += 64) { t += 64; compress_sse(hh, t, buf.data() + off, off + 64 == bytes); }
    t1 = std::chrono::steady_clock::now();
    ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    printf("   sse-rows: %.2f ms (%.0f MB/s) %02x", ms, bytes / 1e3 / ms, (unsigned)(hh[0] & 0xff));
#endif
    printf("\n");
  }
}

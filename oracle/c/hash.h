/* ORACLE (test infrastructure only).
 *
 * Fiat-Shamir transcript of the reference: ark-marlin 0.3 `FiatShamirRng<Blake2s>`
 * (reference barnett-smart-card-protocol/src/discrete_log_cards/mod.rs:9,12,408,436) =
 * Blake2s-256 (RFC 7693) re-seeding a rand_chacha ChaCha20 stream (RFC 7539 core, 64-bit
 * block counter), challenges drawn with ark-ff 0.3 `Fp256::rand`.
 * PARITY UNPINNED vs upstream; pinned to RFC vectors and the Python oracle.
 */
#ifndef ORACLE_HASH_H
#define ORACLE_HASH_H
#include <stdint.h>
#include <string.h>

#include "field.h"

/* ---------------------------------------------------------------- Blake2s-256 (unkeyed) */
typedef struct { uint32_t h[8]; uint64_t t; uint8_t buf[64]; size_t buflen; } blake2s_t;
static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void b2s_compress(blake2s_t* S, const uint8_t* block, int last) {
  uint32_t m[16], v[16];
  memcpy(m, block, 64);
  for (int i = 0; i < 8; i++) { v[i] = S->h[i]; v[i + 8] = B2S_IV[i]; }
  v[12] ^= (uint32_t)S->t;
  v[13] ^= (uint32_t)(S->t >> 32);
  if (last) v[14] = ~v[14];
#define B2S_G(a, b, c, d, x, y)                                   \
  v[a] += v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16);             \
  v[c] += v[d];       v[b] = rotr32(v[b] ^ v[c], 12);             \
  v[a] += v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);              \
  v[c] += v[d];       v[b] = rotr32(v[b] ^ v[c], 7);
  for (int r = 0; r < 10; r++) {
    const uint8_t* s = B2S_SIGMA[r];
    B2S_G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2S_G(1, 5, 9, 13, m[s[2]], m[s[3]])
    B2S_G(2, 6, 10, 14, m[s[4]], m[s[5]]) B2S_G(3, 7, 11, 15, m[s[6]], m[s[7]])
    B2S_G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2S_G(1, 6, 11, 12, m[s[10]], m[s[11]])
    B2S_G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2S_G(3, 4, 9, 14, m[s[14]], m[s[15]])
  }
#undef B2S_G
  for (int i = 0; i < 8; i++) S->h[i] ^= v[i] ^ v[i + 8];
}
static inline void b2s_init(blake2s_t* S) {
  memcpy(S->h, B2S_IV, 32);
  S->h[0] ^= 0x01010020u; /* digest 32 bytes, no key, fanout = depth = 1 */
  S->t = 0;
  S->buflen = 0;
}
static inline void b2s_update(blake2s_t* S, const uint8_t* in, size_t len) {
  while (len > 0) {
    if (S->buflen == 64) { /* buffer full and more input follows: not the last block */
      S->t += 64;
      b2s_compress(S, S->buf, 0);
      S->buflen = 0;
    }
    size_t take = 64 - S->buflen;
    if (take > len) take = len;
    memcpy(S->buf + S->buflen, in, take);
    S->buflen += take;
    in += take;
    len -= take;
  }
}
static inline void b2s_final(blake2s_t* S, uint8_t out[32]) {
  S->t += S->buflen;
  memset(S->buf + S->buflen, 0, 64 - S->buflen);
  b2s_compress(S, S->buf, 1);
  memcpy(out, S->h, 32);
}

/* ---------------------------------------------------------------- ChaCha20 (rand_chacha) */
typedef struct { uint32_t key[8]; uint64_t counter; uint32_t buf[16]; int pos; } chacha_t;
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static inline void chacha_block(chacha_t* c) {
  uint32_t st[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u};
  memcpy(st + 4, c->key, 32);
  st[12] = (uint32_t)c->counter; st[13] = (uint32_t)(c->counter >> 32); st[14] = 0; st[15] = 0;
  uint32_t x[16];
  memcpy(x, st, 64);
#define CC_QR(a, b, cc, d)                              \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);         \
  x[cc] += x[d]; x[b] = rotl32(x[b] ^ x[cc], 12);       \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);          \
  x[cc] += x[d]; x[b] = rotl32(x[b] ^ x[cc], 7);
  for (int i = 0; i < 10; i++) {
    CC_QR(0, 4, 8, 12) CC_QR(1, 5, 9, 13) CC_QR(2, 6, 10, 14) CC_QR(3, 7, 11, 15)
    CC_QR(0, 5, 10, 15) CC_QR(1, 6, 11, 12) CC_QR(2, 7, 8, 13) CC_QR(3, 4, 9, 14)
  }
#undef CC_QR
  for (int i = 0; i < 16; i++) c->buf[i] = x[i] + st[i];
  c->counter++;
  c->pos = 0;
}
static inline void chacha_seed(chacha_t* c, const uint8_t seed[32]) {
  memcpy(c->key, seed, 32);
  c->counter = 0;
  c->pos = 16;
}
static inline uint32_t chacha_u32(chacha_t* c) {
  if (c->pos >= 16) chacha_block(c);
  return c->buf[c->pos++];
}
static inline uint64_t chacha_u64(chacha_t* c) {
  uint64_t lo = chacha_u32(c);
  uint64_t hi = chacha_u32(c);
  return lo | (hi << 32);
}

/* ---------------------------------------------------------------- FiatShamirRng<Blake2s> */
typedef struct { uint8_t seed[32]; chacha_t rng; blake2s_t pending; } fsrng_t;
static inline void fs_from_seed(fsrng_t* fs, const uint8_t* bytes, size_t len) {
  blake2s_t S;
  b2s_init(&S);
  b2s_update(&S, bytes, len);
  b2s_final(&S, fs->seed);
  chacha_seed(&fs->rng, fs->seed);
}
/* absorb(data) = seed <- Blake2s(data || seed); streamed: begin, feed..., end */
static inline void fs_absorb_begin(fsrng_t* fs) { b2s_init(&fs->pending); }
static inline void fs_absorb_feed(fsrng_t* fs, const uint8_t* d, size_t len) { b2s_update(&fs->pending, d, len); }
static inline void fs_absorb_end(fsrng_t* fs) {
  b2s_update(&fs->pending, fs->seed, 32);
  b2s_final(&fs->pending, fs->seed);
  chacha_seed(&fs->rng, fs->seed);
}
/* Fr::rand: 4 x next_u64 as the raw MONTGOMERY representation, top 4 bits cleared, accept
 * iff < modulus.  Result is the field element whose Montgomery form is raw. */
static inline void fs_challenge(fsrng_t* fs, fe* out, const field_t* F) {
  for (;;) {
    uint64_t l[4];
    for (int i = 0; i < 4; i++) l[i] = chacha_u64(&fs->rng);
    l[3] &= 0xFFFFFFFFFFFFFFFFull >> 4;
    if (!limbs_geq(l, F->m)) { memcpy(out->l, l, 32); return; }
  }
}
#endif

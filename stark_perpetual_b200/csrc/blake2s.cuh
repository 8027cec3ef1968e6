// BLAKE2s-256 (RFC 7693), unkeyed, sequential mode -- the hash of the Merkle commitments and of the
// Fiat-Shamir channel (no reference symbol -- the reference's
// starkware/python/merkle_tree.py:4-44 is an update-tree hint helper, not a commitment).
// Same code on device (leaf / node kernels) and host (channel, proof assembly).
#pragma once
#include <stdint.h>

#include "fp.cuh"

struct B2s {
  uint32_t h[8];
};

SPG_HD uint32_t b2s_rotr(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(x, x, n);
#else
  return (x >> n) | (x << (32 - n));
#endif
}
SPG_HD uint32_t b2s_bswap(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, 0, 0x0123);
#else
  return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}

SPG_HD void b2s_init(B2s& s) {
  s.h[0] = 0x6A09E667u ^ 0x01010020u; s.h[1] = 0xBB67AE85u; s.h[2] = 0x3C6EF372u; s.h[3] = 0xA54FF53Au;
  s.h[4] = 0x510E527Fu; s.h[5] = 0x9B05688Cu; s.h[6] = 0x1F83D9ABu; s.h[7] = 0x5BE0CD19u;
}

// The xor / rotate half of G (8 of its 14 operations) can only run on the ALU pipe, which is what bounds the leaf
// kernel (96 % active).  On the device the six additions are therefore issued as multiply-adds by a one the
// compiler cannot see through (a __constant__ word), which run on the FMA pipe next to the ALU work.
#if defined(__CUDACC__)
static __constant__ uint32_t spg_b2s_one = 1;
#endif
SPG_HD uint32_t b2s_add(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && !defined(SPG_B2S_PLAIN_ADD)
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(spg_b2s_one), "r"(a));
  return r;
#else
  return a + b;
#endif
}
#define B2S_G(a, b, c, d, x, y)                                      \
  do {                                                               \
    a = b2s_add(b2s_add(a, b), (x)); d = b2s_rotr(d ^ a, 16);        \
    c = b2s_add(c, d);               b = b2s_rotr(b ^ c, 12);        \
    a = b2s_add(b2s_add(a, b), (y)); d = b2s_rotr(d ^ a, 8);         \
    c = b2s_add(c, d);               b = b2s_rotr(b ^ c, 7);         \
  } while (0)

#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  do {                                                                                  \
    B2S_G(v0, v4, v8, v12, m[s0], m[s1]);                                               \
    B2S_G(v1, v5, v9, v13, m[s2], m[s3]);                                               \
    B2S_G(v2, v6, v10, v14, m[s4], m[s5]);                                              \
    B2S_G(v3, v7, v11, v15, m[s6], m[s7]);                                              \
    B2S_G(v0, v5, v10, v15, m[s8], m[s9]);                                              \
    B2S_G(v1, v6, v11, v12, m[s10], m[s11]);                                            \
    B2S_G(v2, v7, v8, v13, m[s12], m[s13]);                                             \
    B2S_G(v3, v4, v9, v14, m[s14], m[s15]);                                             \
  } while (0)

// one compression: m = 16 little-endian message words, t = bytes hashed so far INCLUDING this block
SPG_HD void b2s_compress(B2s& s, const uint32_t (&m)[16], uint64_t t, bool last) {
  uint32_t v0 = s.h[0], v1 = s.h[1], v2 = s.h[2], v3 = s.h[3], v4 = s.h[4], v5 = s.h[5], v6 = s.h[6], v7 = s.h[7];
  uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
  uint32_t v12 = 0x510E527Fu ^ (uint32_t)t, v13 = 0x9B05688Cu ^ (uint32_t)(t >> 32);
  uint32_t v14 = last ? ~0x1F83D9ABu : 0x1F83D9ABu, v15 = 0x5BE0CD19u;
  B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
  B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
  B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
  B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
  B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
  B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
  B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
  B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
  B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
  B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
  s.h[0] ^= v0 ^ v8; s.h[1] ^= v1 ^ v9; s.h[2] ^= v2 ^ v10; s.h[3] ^= v3 ^ v11;
  s.h[4] ^= v4 ^ v12; s.h[5] ^= v5 ^ v13; s.h[6] ^= v6 ^ v14; s.h[7] ^= v7 ^ v15;
}

// message words of one field element serialised as 32 bytes big-endian (a canonical value in [0, p))
SPG_HD void b2s_felt_words(const Fp& a, uint32_t* m8) {
#pragma unroll
  for (int k = 0; k < 8; k++) m8[k] = b2s_bswap(a.v[7 - k]);
}

// ---- host-only: hash of an arbitrary byte string
static inline void b2s_hash_bytes(const uint8_t* data, size_t len, uint8_t out[32]) {
  B2s s;
  b2s_init(s);
  uint32_t m[16];
  size_t off = 0;
  while (len - off > 64) {
    for (int i = 0; i < 16; i++)
      m[i] = (uint32_t)data[off + 4 * i] | ((uint32_t)data[off + 4 * i + 1] << 8) |
             ((uint32_t)data[off + 4 * i + 2] << 16) | ((uint32_t)data[off + 4 * i + 3] << 24);
    off += 64;
    b2s_compress(s, m, off, false);
  }
  uint8_t blk[64] = {0};
  for (size_t i = 0; i < len - off; i++) blk[i] = data[off + i];
  for (int i = 0; i < 16; i++)
    m[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) |
           ((uint32_t)blk[4 * i + 3] << 24);
  b2s_compress(s, m, len, true);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)s.h[i]; out[4 * i + 1] = (uint8_t)(s.h[i] >> 8);
    out[4 * i + 2] = (uint8_t)(s.h[i] >> 16); out[4 * i + 3] = (uint8_t)(s.h[i] >> 24);
  }
}

// SHA-256 (FIPS 180-4) and HMAC-SHA256 (RFC 2104) with 32-byte keys: the hash under the RFC 6979 nonce derivation
// of the reference's `sign` (sign.cuh).  Written once for the CUDA kernel (ecdsa.cu) and the host emulation
// (tests/host_emul); streaming byte interface, state kept as big-endian words.
#pragma once
#include "fp.cuh"

struct Sha256 {
  uint32_t h[8];
  uint32_t w[16];      // current block, big-endian words
  uint32_t fill;       // bytes in the current block
  uint32_t total;      // bytes hashed so far
};

#define SPG_SHA256_K \
  0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, \
  0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, \
  0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, \
  0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, \
  0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, \
  0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, \
  0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, \
  0xc67178f2u \

#if defined(__CUDACC__)
static __constant__ uint32_t spg_sha256_k_dev[64] = {SPG_SHA256_K};
#endif
static const uint32_t spg_sha256_k_host[64] = {SPG_SHA256_K};
SPG_HD uint32_t sha256_k(int t) {
#if defined(__CUDA_ARCH__)
  return spg_sha256_k_dev[t];
#else
  return spg_sha256_k_host[t];
#endif
}

SPG_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

SPG_HD void sha256_init(Sha256& c) {
  c.h[0] = 0x6a09e667u; c.h[1] = 0xbb67ae85u; c.h[2] = 0x3c6ef372u; c.h[3] = 0xa54ff53au;
  c.h[4] = 0x510e527fu; c.h[5] = 0x9b05688cu; c.h[6] = 0x1f83d9abu; c.h[7] = 0x5be0cd19u;
  for (int i = 0; i < 16; i++) c.w[i] = 0;
  c.fill = 0; c.total = 0;
}
SPG_HD void sha256_compress(Sha256& c) {
  uint32_t w[16];
  for (int i = 0; i < 16; i++) w[i] = c.w[i];
  uint32_t a = c.h[0], b = c.h[1], cc = c.h[2], d = c.h[3], e = c.h[4], f = c.h[5], g = c.h[6], h = c.h[7];
  for (int t = 0; t < 64; t++) {
    uint32_t wt;
    if (t < 16) wt = w[t];
    else {
      const uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
      const uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      wt = w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
    }
    const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    const uint32_t ch = (e & f) ^ (~e & g);
    const uint32_t t1 = h + S1 + ch + sha256_k(t) + wt;
    const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    const uint32_t mj = (a & b) ^ (a & cc) ^ (b & cc);
    const uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = cc; cc = b; b = a; a = t1 + t2;
  }
  c.h[0] += a; c.h[1] += b; c.h[2] += cc; c.h[3] += d; c.h[4] += e; c.h[5] += f; c.h[6] += g; c.h[7] += h;
  for (int i = 0; i < 16; i++) c.w[i] = 0;
  c.fill = 0;
}
SPG_HD void sha256_byte(Sha256& c, uint32_t byte) {
  c.w[c.fill >> 2] |= (byte & 0xffu) << (24 - 8 * (c.fill & 3));
  c.fill++; c.total++;
  if (c.fill == 64) sha256_compress(c);
}
// 32 bytes given as 8 big-endian words
SPG_HD void sha256_words(Sha256& c, const uint32_t w[8]) {
  for (int i = 0; i < 8; i++)
    for (int k = 3; k >= 0; k--) sha256_byte(c, w[i] >> (8 * k));
}
SPG_HD void sha256_final(Sha256& c, uint32_t out[8]) {
  const uint32_t bits = c.total * 8;
  sha256_byte(c, 0x80);
  if (c.fill > 56) sha256_compress(c);
  c.w[15] = bits;                      // messages here are far shorter than 2^32 bits
  sha256_compress(c);
  for (int i = 0; i < 8; i++) out[i] = c.h[i];
}

// HMAC-SHA256 with a 32-byte key (8 big-endian words): start the inner hash / finish with the outer hash
SPG_HD void hmac_begin(Sha256& c, const uint32_t key[8]) {
  sha256_init(c);
  for (int i = 0; i < 8; i++) c.w[i] = key[i] ^ 0x36363636u;
  for (int i = 8; i < 16; i++) c.w[i] = 0x36363636u;
  c.total = 64;
  c.fill = 64;
  sha256_compress(c);
}
SPG_HD void hmac_end(Sha256& c, const uint32_t key[8], uint32_t out[8]) {
  uint32_t inner[8];
  sha256_final(c, inner);
  sha256_init(c);
  for (int i = 0; i < 8; i++) c.w[i] = key[i] ^ 0x5c5c5c5cu;
  for (int i = 8; i < 16; i++) c.w[i] = 0x5c5c5c5cu;
  c.total = 64;
  c.fill = 64;
  sha256_compress(c);
  sha256_words(c, inner);
  sha256_final(c, out);
}

// value (8 little-endian u32 limbs) -> 8 big-endian words of its 32-byte big-endian encoding
SPG_HD void u256_to_be_words(const uint32_t v[8], uint32_t w[8]) {
  for (int i = 0; i < 8; i++) w[i] = v[7 - i];
}


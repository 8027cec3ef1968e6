// Deterministic STARK-curve ECDSA signing, per-signature code shared by the CUDA kernel (ecdsa.cu) and the host
// emulation (tests/host_emul/emul_ecdsa.cpp).
//
// Reference restated: src/starkware/crypto/signature/signature.py:137-173 (sign and its three rejection rules),
// :117-134 (generate_k_rfc6979: the nibble rule, `extra_entropy` from the seed) on top of python-ecdsa 0.17.0
// `rfc6979.generate_k` (requirements pin; the package is not under /root/reference -- its algorithm is RFC 6979
// section 3.2 with HMAC-SHA256, qlen = 252, rolen = 32), :113-114 (inv_mod_curve_size).
#pragma once
#include "ecdsa.cuh"
#include "sha256.cuh"

// RFC 6979 nonce for (msg_hash, priv_key, seed) exactly as signature.py:117-134 produces it.
//   * signature.py:119-121 multiplies the hash by 16 when its bit length is 249..251 so that python-ecdsa's
//     bits2int (which drops the low bits of a 256-bit string) gives the hash back; for every msg_hash < 2^251 the
//     octet string that enters the DRBG is therefore msg_hash itself as 32 big-endian bytes (it is < n, so
//     bits2octets subtracts nothing);
//   * extra_entropy = the seed's minimal big-endian bytes (none for seed None or 0);
//   * candidate = the first 252 bits of each 32-byte HMAC output (>> 4); accepted when 1 <= k < n.
// k_out: 8 little-endian limbs.
SPG_HD void rfc6979_nonce(const uint32_t (&msg)[8], const uint32_t (&priv)[8], unsigned long long seed,
                          uint32_t (&k_out)[8]) {
  uint32_t pw[8], mw[8], K[8], V[8];
  u256_to_be_words(priv, pw);
  u256_to_be_words(msg, mw);
  for (int i = 0; i < 8; i++) { K[i] = 0; V[i] = 0x01010101u; }
  int seed_bytes = 0;
  while (seed_bytes < 8 && (seed >> (8 * seed_bytes)) != 0) seed_bytes++;
  Sha256 c;
  for (uint32_t sep = 0; sep < 2; sep++) {
    hmac_begin(c, K);
    sha256_words(c, V);
    sha256_byte(c, sep);
    sha256_words(c, pw);
    sha256_words(c, mw);
    for (int b = seed_bytes - 1; b >= 0; b--) sha256_byte(c, (uint32_t)(seed >> (8 * b)));
    uint32_t nk[8];
    hmac_end(c, K, nk);
    for (int i = 0; i < 8; i++) K[i] = nk[i];
    hmac_begin(c, K);
    sha256_words(c, V);
    hmac_end(c, K, V);
  }
  for (;;) {
    hmac_begin(c, K);
    sha256_words(c, V);
    hmac_end(c, K, V);
    uint32_t cand[8];
    for (int i = 0; i < 8; i++) {          // (V as a 256-bit big-endian integer) >> 4, little-endian limbs
      const uint32_t lo = V[7 - i], hi = (i < 7) ? V[6 - i] : 0u;
      cand[i] = (lo >> 4) | (hi << 28);
    }
    uint32_t nz = 0;
    for (int i = 0; i < 8; i++) nz |= cand[i];
    if (nz != 0 && !fn_geq_n(cand)) {
      for (int i = 0; i < 8; i++) k_out[i] = cand[i];
      return;
    }
    hmac_begin(c, K);
    sha256_words(c, V);
    sha256_byte(c, 0);
    uint32_t nk[8];
    hmac_end(c, K, nk);
    for (int i = 0; i < 8; i++) K[i] = nk[i];
    hmac_begin(c, K);
    sha256_words(c, V);
    hmac_end(c, K, V);
  }
}

// a + b mod n for canonical a, b
SPG_HD Fn fn_add(const Fn& a, const Fn& b) {
  const uint32_t n[8] = SPG_N_LIMBS;
  Fn r;
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  if (c || fn_geq_n(r.v)) {
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
      const uint64_t d = (uint64_t)r.v[i] - n[i] - br;
      r.v[i] = (uint32_t)d; br = (d >> 32) & 1;
    }
  }
  return r;
}

// sign(msg_hash, priv_key, seed) (signature.py:137-173).  Seed None and seed 0 are the same signature (no extra
// entropy on the first attempt, 1 on the second), so the seed is a plain integer here.
// status: 0 ok (r, s written, canonical little-endian limbs); 1 "Message not signable" (msg >= 2^251, :141);
//         2 private key outside [1, n) (outside the reference's domain: is_valid_stark_private_key);
//         3 no signature within SPG_SIGN_MAX_TRIES attempts (an attempt is rejected with probability ~2^-54 by the
//           2^251 bounds, so this does not happen).
#define SPG_SIGN_MAX_TRIES 128
SPG_HD int ecdsa_sign_one(const uint32_t (&msg)[8], const uint32_t (&priv)[8], unsigned long long seed,
                          const EcdsaTables& T, uint32_t (&r_out)[8], uint32_t (&s_out)[8]) {
  if (!u256_lt_2_251(msg)) return 1;
  if (u256_is_zero(priv) || fn_geq_n(priv)) return 2;
  Fn d, z0;
  for (int i = 0; i < 8; i++) { d.v[i] = priv[i]; z0.v[i] = msg[i]; }      // msg < 2^251 < n
  const Fn dm = fn_mul(d, T.r2_n, T.ninv);                                  // d R mod n
  for (int attempt = 0; attempt < SPG_SIGN_MAX_TRIES; attempt++, seed++) {
    uint32_t k[8];
    rfc6979_nonce(msg, priv, seed, k);
    // r = x(k G) as an integer; 1 <= r < 2^251 (:158-160)
    const Fp x = fp_from_mont(gen_mult(k, T).x);
    uint32_t r[8]; for (int i = 0; i < 8; i++) r[i] = x.v[i];
    if (u256_is_zero(r) || !u256_lt_2_251(r)) continue;
    // z = msg + r d mod n != 0 (:162-164)
    Fn rn; for (int i = 0; i < 8; i++) rn.v[i] = r[i];
    const Fn z = fn_add(z0, fn_mul(rn, dm, T.ninv));
    if (u256_is_zero(z.v)) continue;
    // w = k / z mod n, 1 <= w < 2^251 (:166-170)
    Fn kn; for (int i = 0; i < 8; i++) kn.v[i] = k[i];
    const Fn w = fn_mul(fn_mul(kn, T.r2_n, T.ninv), fn_inv(z, T), T.ninv);
    if (u256_is_zero(w.v) || !u256_lt_2_251(w.v)) continue;
    const Fn s = fn_inv(w, T);                                              // inv_mod_curve_size (:113-114, :172)
    for (int i = 0; i < 8; i++) { r_out[i] = r[i]; s_out[i] = s.v[i]; }
    return 0;
  }
  return 3;
}

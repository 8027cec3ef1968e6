// STARK-curve arithmetic  y^2 = x^3 + x + beta  over the STARK prime (alpha = 1).
// Reference semantics restated: src/starkware/crypto/signature/math_utils.py:59-100 (affine
// ec_add / ec_double / ec_mult with their assertions), signature.py:176-190, 296-318.
//
// Device code works in Jacobian coordinates (X : Y : Z), x = X / Z^2, y = Y / Z^3, Montgomery residues
// with the lazy bounds of fp.cuh (all values < 2p).  The reference's assertions are about AFFINE x / y
// being equal or zero; in Jacobian form  x1 == x2  <=>  X1 * Z2^2 == X2 * Z1^2  and  y == 0 <=> Y == 0,
// so every exceptional case of the reference is detected exactly, without inversions.
#pragma once
#include "fp.cuh"

struct APoint { Fp x, y; };          // affine, Montgomery form
struct JPoint { Fp X, Y, Z; };       // Jacobian, Z != 0

// a^-1 = a^(p-2), p - 2 = 2^251 + 2^196 + 2^192 - 1: 251 squarings + 13 multiplications
SPG_HD Fp fp_inv_chain(const Fp& a) {
  Fp x2 = fp_mul(fp_sqr(a), a);                       // 2^2 - 1
  Fp t = x2;
  for (int i = 0; i < 2; i++) t = fp_sqr(t);
  Fp x4 = fp_mul(t, x2);
  t = x4;
  for (int i = 0; i < 4; i++) t = fp_sqr(t);
  Fp x8 = fp_mul(t, x4);
  t = x8;
  for (int i = 0; i < 8; i++) t = fp_sqr(t);
  Fp x16 = fp_mul(t, x8);
  t = x16;
  for (int i = 0; i < 16; i++) t = fp_sqr(t);
  Fp x32 = fp_mul(t, x16);
  t = x32;
  for (int i = 0; i < 32; i++) t = fp_sqr(t);
  Fp x64 = fp_mul(t, x32);
  t = x64;
  for (int i = 0; i < 64; i++) t = fp_sqr(t);
  Fp x128 = fp_mul(t, x64);
  t = x128;
  for (int i = 0; i < 64; i++) t = fp_sqr(t);
  Fp x192 = fp_mul(t, x64);                           // a^(2^192 - 1)
  Fp e192 = fp_mul(x192, a);                          // a^(2^192)
  t = e192;
  for (int i = 0; i < 4; i++) t = fp_sqr(t);          // a^(2^196)
  Fp e196 = t;
  for (int i = 0; i < 55; i++) t = fp_sqr(t);         // a^(2^251)
  return fp_mul(fp_mul(t, e196), x192);
}

// Jacobian + affine, NO special cases (caller has checked x1 != x2).  8M + 3S.
// zz = Z1^2 and zzz = Z1^3 are passed in (callers cache them) together with u2 = x2 * zz.
SPG_HD JPoint ec_madd_nocheck(const JPoint& p, const APoint& q, const Fp& zzz, const Fp& u2) {
  Fp s2 = fp_mul(q.y, zzz);
  Fp h = fp_sub(u2, p.X);
  Fp r = fp_sub(s2, p.Y);
  Fp hh = fp_sqr(h);
  Fp hhh = fp_mul(h, hh);
  Fp v = fp_mul(p.X, hh);
  JPoint o;
  o.X = fp_sub(fp_sub(fp_sqr(r), hhh), fp_add(v, v));
  o.Y = fp_sub(fp_mul(r, fp_sub(v, o.X)), fp_mul(p.Y, hhh));
  o.Z = fp_mul(p.Z, h);
  return o;
}

// Jacobian + Jacobian, NO special cases (caller has checked x1 != x2, given u1, u2, and both Z^2).
SPG_HD JPoint ec_jadd_nocheck(const JPoint& p, const JPoint& q, const Fp& z1z1, const Fp& z2z2, const Fp& u1,
                              const Fp& u2) {
  Fp s1 = fp_mul(p.Y, fp_mul(q.Z, z2z2));
  Fp s2 = fp_mul(q.Y, fp_mul(p.Z, z1z1));
  Fp h = fp_sub(u2, u1);
  Fp r = fp_sub(s2, s1);
  Fp hh = fp_sqr(h);
  Fp hhh = fp_mul(h, hh);
  Fp v = fp_mul(u1, hh);
  JPoint o;
  o.X = fp_sub(fp_sub(fp_sqr(r), hhh), fp_add(v, v));
  o.Y = fp_sub(fp_mul(r, fp_sub(v, o.X)), fp_mul(s1, hhh));
  o.Z = fp_mul(fp_mul(p.Z, q.Z), h);
  return o;
}

// Jacobian doubling for alpha = 1 (caller has checked Y != 0).  4M + 6S... M = 3 X^2 + Z^4.
SPG_HD JPoint ec_jdouble_nocheck(const JPoint& p) {
  Fp xx = fp_sqr(p.X);
  Fp yy = fp_sqr(p.Y);
  Fp yyyy = fp_sqr(yy);
  Fp zz = fp_sqr(p.Z);
  Fp s = fp_mul(p.X, yy);
  s = fp_add(s, s); s = fp_add(s, s);                              // 4 X Y^2
  Fp m = fp_add(fp_add(fp_add(xx, xx), xx), fp_sqr(zz));           // 3 X^2 + alpha Z^4
  JPoint o;
  o.X = fp_sub(fp_sqr(m), fp_add(s, s));
  Fp y8 = fp_add(yyyy, yyyy); y8 = fp_add(y8, y8); y8 = fp_add(y8, y8);
  o.Y = fp_sub(fp_mul(m, fp_sub(s, o.X)), y8);
  Fp yz = fp_mul(p.Y, p.Z);
  o.Z = fp_add(yz, yz);
  return o;
}

// ---- affine arithmetic with one inversion per operation (host table building, witness generation)
SPG_HD APoint ec_affine_add(const APoint& a, const APoint& b) {     // math_utils.py:59-68, x1 != x2 assumed
  Fp m = fp_mul(fp_sub(a.y, b.y), fp_inv_chain(fp_sub(a.x, b.x)));
  APoint o;
  o.x = fp_sub(fp_sub(fp_sqr(m), a.x), b.x);
  o.y = fp_sub(fp_mul(m, fp_sub(a.x, o.x)), a.y);
  return o;
}
SPG_HD APoint ec_affine_double(const APoint& a) {                   // math_utils.py:79-88, y != 0 assumed
  Fp xx = fp_sqr(a.x);
  Fp num = fp_add(fp_add(fp_add(xx, xx), xx), fp_one());
  Fp m = fp_mul(num, fp_inv_chain(fp_add(a.y, a.y)));
  APoint o;
  o.x = fp_sub(fp_sub(fp_sqr(m), a.x), a.x);
  o.y = fp_sub(fp_mul(m, fp_sub(a.x, o.x)), a.y);
  return o;
}

// canonical 4 x u64 value >= p ?
SPG_HD bool spg_canon_geq_p(const uint32_t* v) {
  // p = 0x08000000 00000011 00000000 ... 00000001
  if (v[7] != SPG_P7) return v[7] > SPG_P7;
  if (v[6] != SPG_P6) return v[6] > SPG_P6;
  if (v[5] | v[4] | v[3] | v[2] | v[1]) return true;
  return v[0] >= SPG_P0;
}

#define SPG_N_CONST_POINTS 506
#define SPG_HASH_BITS 252      // N_ELEMENT_BITS_HASH, signature.py:50
#define SPG_ECDSA_BITS 251     // N_ELEMENT_BITS_ECDSA, signature.py:47

// ---- Pedersen accumulator (shared by the hash kernel and the host emulation test)
struct PedersenAcc {
  JPoint p;
  Fp zz, zzz;
  SPG_HD void init(const APoint& s) {
    p.X = s.x; p.Y = s.y; p.Z = fp_one(); zz = fp_one(); zzz = fp_one();
  }
};

// Absorb one canonical element (8 x u32 limbs, already checked < p) with table slice `tab` (252 points).
// Returns false on "Unhashable input." (signature.py:313).
SPG_HD bool pedersen_absorb(PedersenAcc& a, const uint32_t (&x)[8], const APoint* tab) {
  bool ok = true;
#pragma unroll 1
  for (int w = 0; w < 8; w++) {
    uint32_t word = x[0];
    // select limb w without dynamic register indexing
#pragma unroll
    for (int k = 1; k < 8; k++) word = (w == k) ? x[k] : word;
    const int nb = (w == 7) ? (SPG_HASH_BITS - 224) : 32;
#pragma unroll 1
    for (int b = 0; b < nb; b++) {
      const APoint q = tab[w * 32 + b];
      const Fp u2 = fp_mul(q.x, a.zz);
      if (fp_is_zero(fp_sub(u2, a.p.X))) ok = false;
      if ((word >> b) & 1u) {
        a.p = ec_madd_nocheck(a.p, q, a.zzz, u2);
        a.zz = fp_sqr(a.p.Z);
        a.zzz = fp_mul(a.zz, a.p.Z);
      }
    }
  }
  return ok;
}


// ---- the same absorption as ONE stream of set bits over both elements (n_elems = 1: x only).
// pedersen_absorb branches on the bit inside every step; in a warp some lane almost always has the bit set, so every step
// pays the 13-multiplication addition although half the lanes idle through it.  Here a thread jumps from set bit to set
// bit: the additions of a warp's lanes line up (k-th addition with k-th addition), and what differs per lane -- the
// collision checks of the steps skipped over, one multiplication each (signature.py:313 runs on EVERY step) -- is a
// short loop.  Same result, same "Unhashable input." detection, ~504 x 14 -> ~275 x 19 multiplications per warp and hash.
SPG_HD uint32_t ped_word(const uint32_t (&x)[8], const uint32_t (&y)[8], int wi) {
  uint32_t w = x[0];
#pragma unroll
  for (int k = 1; k < 8; k++) w = (wi == k) ? x[k] : w;
#pragma unroll
  for (int k = 0; k < 8; k++) w = (wi == 8 + k) ? y[k] : w;
  return w;
}
SPG_HD int ped_ctz(uint32_t w) {
#ifdef __CUDA_ARCH__
  return __ffs((int)w) - 1;
#else
  return __builtin_ctz(w);
#endif
}
// bits = steps per element (252 for the hash, 251 for mimic_ec_mult_air over the generator table); bits of an element
// above that are ignored, as the step loops ignore them
SPG_HD bool pedersen_absorb_stream(PedersenAcc& a, const uint32_t (&x)[8], const uint32_t (&y)[8], int n_elems, const APoint* tab,
                                   int bits = SPG_HASH_BITS) {
  bool ok = true;
  const int total = bits * n_elems, last_word = 8 * n_elems - 1;
  const uint32_t top_mask = (1u << (bits - 224)) - 1u;
  int s = 0;                      // next step whose collision check is due
  int wi = 0, base = 0;           // current word of the stream and the step index of its bit 0
  uint32_t word = x[0];
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    while (word == 0 && wi < last_word) {
      wi++;
      word = ped_word(x, y, wi);
      if ((wi & 7) == 7) word &= top_mask;
      base = (wi >> 3) * bits + (wi & 7) * 32;
    }
    if (word == 0) break;
    const int sb = base + ped_ctz(word);
    word &= word - 1;
    Fp u2 = fp_zero();
#pragma unroll 1
    for (; s <= sb; s++) {        // checks of the skipped steps and of step sb itself, all against the current sum
      u2 = fp_mul(tab[s].x, a.zz);
      if (fp_is_zero(fp_sub(u2, a.p.X))) ok = false;
    }
    a.p = ec_madd_nocheck(a.p, tab[sb], a.zzz, u2);
    a.zz = fp_sqr(a.p.Z);
    a.zzz = fp_mul(a.zz, a.p.Z);
  }
#pragma unroll 1
  for (; s < total; s++)
    if (fp_is_zero(fp_sub(fp_mul(tab[s].x, a.zz), a.p.X))) ok = false;
  return ok;
}
// ---- the stream with DEFERRED checks.  In pedersen_absorb_stream the check loop between two additions runs as long as the
// longest gap among a warp's lanes (6.4 steps on average for random scalars, against a mean gap of 2), so two thirds of
// its lane-slots idle.  Here the partial sums of the last SPG_PED_RING additions are kept (X and Z^2, thread-private), the
// checks of skipped steps queue up behind a second cursor and every addition serves at most SPG_PED_BUDGET of them: the
// per-round work is the same for every lane, and a check still compares its step's table point with the partial sum that
// was current at that step (signature.py:313).  A lane whose queue would outlive the ring (a zero run longer than ~16
// steps) drains it on the spot.
// MEASURED (profiles/r2ab6_pedersen_deferred_ab.json): 2^20 random pairs 56.4 -> 54.8 ms, the order pipeline 46.9 -> 49.3 ms
// -- the kernel is bound by the latency of the dependent addition chain (8 warps per SM), not by the idle lanes of the
// check loop, so this form is NOT the default (PEDERSEN_STREAM=2 selects it); it stays as the tested record of the attempt.
#define SPG_PED_RING 8
#define SPG_PED_BUDGET 2
struct PedCheckCursor {          // the next step to check: word index, bit position in it, remaining bits of that word
  int wi, pos, round;            // round = number of set bits below the step = index of the partial sum it is checked against
  uint32_t bits;
};
SPG_HD bool pedersen_absorb_deferred(PedersenAcc& a, const uint32_t (&x)[8], const uint32_t (&y)[8], int n_elems, const APoint* tab,
                                     int bits = SPG_HASH_BITS) {
  bool ok = true;
  const int last_word = 8 * n_elems - 1, top_len = bits - 224;
  const uint32_t top_mask = (1u << top_len) - 1u;
  Fp rx[SPG_PED_RING], rz[SPG_PED_RING];
  int k = 0;                                     // additions done so far = index of the current partial sum
  rx[0] = a.p.X; rz[0] = a.zz;
  int wi = 0, base = 0;
  uint32_t word = x[0];
  PedCheckCursor c = {0, 0, 0, x[0]};
  // one queued step: a set-bit step was checked inside its addition (it only advances the round); an unset one is
  // checked now against the partial sum of its round.  Returns true if a multiplication was spent.
  auto serve = [&]() -> bool {
    const int step = (c.wi >> 3) * bits + (c.wi & 7) * 32 + c.pos;
    bool spent = false;
    if (c.bits & 1u) c.round++;
    else {
      const int slot = c.round & (SPG_PED_RING - 1);
      if (fp_is_zero(fp_sub(fp_mul(tab[step].x, rz[slot]), rx[slot]))) ok = false;
      spent = true;
    }
    c.bits >>= 1;
    if (++c.pos == (((c.wi & 7) == 7) ? top_len : 32)) {
      c.pos = 0;
      c.wi++;
      c.bits = c.wi <= last_word ? ped_word(x, y, c.wi) : 0u;
      if ((c.wi & 7) == 7) c.bits &= top_mask;
    }
    return spent;
  };
  auto cursor_step = [&]() { return (c.wi >> 3) * bits + (c.wi & 7) * 32 + c.pos; };
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    while (word == 0 && wi < last_word) {
      wi++;
      word = ped_word(x, y, wi);
      if ((wi & 7) == 7) word &= top_mask;
      base = (wi >> 3) * bits + (wi & 7) * 32;
    }
    if (word == 0) break;
    const int sb = base + ped_ctz(word);
    word &= word - 1;
    // slot (k + 1) mod RING is about to be overwritten: everything still queued against partial sum k + 1 - RING goes first
#pragma unroll 1
    while (c.round <= k + 1 - SPG_PED_RING) serve();
    const APoint q = tab[sb];
    const Fp u2 = fp_mul(q.x, a.zz);
    if (fp_is_zero(fp_sub(u2, a.p.X))) ok = false;          // the set-bit step's own check
    a.p = ec_madd_nocheck(a.p, q, a.zzz, u2);
    a.zz = fp_sqr(a.p.Z);
    a.zzz = fp_mul(a.zz, a.p.Z);
    k++;
    rx[k & (SPG_PED_RING - 1)] = a.p.X; rz[k & (SPG_PED_RING - 1)] = a.zz;
    // serve the queue: steps up to sb have their partial sums in the ring
    int budget = SPG_PED_BUDGET;
#pragma unroll 1
    while (budget > 0 && cursor_step() <= sb) if (serve()) budget--;
  }
  const int total = bits * n_elems;
#pragma unroll 1
  while (cursor_step() < total) serve();
  return ok;
}
#ifndef PEDERSEN_STREAM
#define PEDERSEN_STREAM 1         // 0: the step-by-step absorption, 1: set-bit stream, 2: stream with deferred checks
#endif
// both elements (n_elems = 2) or x alone (n_elems = 1) into the accumulator
SPG_HD bool pedersen_absorb_elems(PedersenAcc& a, const uint32_t (&x)[8], const uint32_t (&y)[8], int n_elems, const APoint* tab) {
#if PEDERSEN_STREAM == 2
  return pedersen_absorb_deferred(a, x, y, n_elems, tab);
#elif PEDERSEN_STREAM
  return pedersen_absorb_stream(a, x, y, n_elems, tab);
#else
  bool ok = pedersen_absorb(a, x, tab);
  if (n_elems > 1) ok = pedersen_absorb(a, y, tab + SPG_HASH_BITS) && ok;
  return ok;
#endif
}

// One pedersen_hash(x, y) (signature.py:296-318) of canonical operands already known to be < p; cp = the 506-point table.
// Returns false on "Unhashable input." (signature.py:313); *out_canon receives the canonical hash.
SPG_HD bool pedersen_hash2_one(const uint32_t (&x)[8], const uint32_t (&y)[8], const APoint* cp, Fp* out_canon) {
  PedersenAcc a;
  a.init(cp[0]);
  if (!pedersen_absorb_elems(a, x, y, 2, cp + 2)) { *out_canon = fp_zero(); return false; }
  const Fp zi = fp_inv_chain(a.p.Z);
  *out_canon = fp_from_mont(fp_mul(a.p.X, fp_sqr(zi)));
  return true;
}

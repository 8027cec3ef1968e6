// STARK-curve ECDSA verification, per-signature code shared by the CUDA kernel (ecdsa.cu) and the host
// emulation test (tests/host_emul/emul_ecdsa.cpp).
// Reference restated: src/starkware/crypto/signature/signature.py:217-260 (verify), :176-190
// (mimic_ec_mult_air), :84-96 (get_y_coordinate), math_utils.py:36-47 (is_quad_residue / sqrt_mod, smaller
// root), :59-88 (ec_add / ec_double and their assertions).
//
// Every assertion of the reference is decided exactly in Jacobian coordinates (x1 == x2 <=> X1 Z2^2 == X2 Z1^2,
// y == 0 <=> Y == 0), so the result -- True / False / "raises" -- is the reference's for every input.
#pragma once
#include "ec.cuh"

// ------------------------------------------------------------------ arithmetic modulo the curve order n
// n = 0x0800000000000010 ffffffffffffffff b781126dcae7b232 1e66a241adc64d2f  (signature.py:55, params json)
struct Fn { uint32_t v[8]; };
#define SPG_N_LIMBS {0xadc64d2fu, 0x1e66a241u, 0xcae7b232u, 0xb781126du, 0xffffffffu, 0xffffffffu, 0x00000010u, 0x08000000u}
#define SPG_N_INV32 0xe8bde631u    // -n^-1 mod 2^32 (-n^-1 mod 2^64 = 0xbb6b3c4ce8bde631)

SPG_HD bool fn_geq_n(const uint32_t* a) {
  const uint32_t n[8] = SPG_N_LIMBS;
  for (int i = 7; i >= 0; i--) { if (a[i] != n[i]) return a[i] > n[i]; }
  return true;
}
SPG_HD bool u256_is_zero(const uint32_t* a) {
  uint32_t d = 0;
  for (int i = 0; i < 8; i++) d |= a[i];
  return d == 0;
}
// value < 2^251 ?
SPG_HD bool u256_lt_2_251(const uint32_t* a) { return (a[7] >> 27) == 0; }

// Montgomery product modulo n (CIOS, 32-bit limbs, 64-bit accumulators), fully reduced output
SPG_HD Fn fn_mul(const Fn& a, const Fn& b, uint32_t ninv) {
  const uint32_t n[8] = SPG_N_LIMBS;
  uint32_t t[10];
  for (int i = 0; i < 10; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
      uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c;
      t[j] = (uint32_t)s; c = s >> 32;
    }
    uint64_t s = (uint64_t)t[8] + c;
    t[8] = (uint32_t)s; t[9] = (uint32_t)(s >> 32);
    const uint32_t m = t[0] * ninv;
    c = ((uint64_t)m * n[0] + t[0]) >> 32;
    for (int j = 1; j < 8; j++) {
      uint64_t s2 = (uint64_t)m * n[j] + t[j] + c;
      t[j - 1] = (uint32_t)s2; c = s2 >> 32;
    }
    s = (uint64_t)t[8] + c;
    t[7] = (uint32_t)s;
    t[8] = t[9] + (uint32_t)(s >> 32);
    t[9] = 0;
  }
  Fn r;
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  if (t[8] || fn_geq_n(r.v)) {
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t d = (uint64_t)r.v[i] - n[i] - br;
      r.v[i] = (uint32_t)d; br = (d >> 32) & 1;
    }
  }
  return r;
}

// tables / constants living in device (or host) memory
struct EcdsaTables {
  const APoint* gen_doubles;   // G * 2^t, t < 252
  APoint shift, minus_shift;   // Montgomery affine
  Fp beta;                     // Montgomery
  Fn r2_n;                     // 2^512 mod n
  uint32_t ninv;               // -n^-1 mod 2^32
  // square roots: c = 3^q generates the 2-Sylow subgroup (order 2^192), ci = c^-1
  const Fp* sq_L;              // [256]      (c^(2^184))^k
  const Fp* sq_D;              // [24][256]  ci^(k 2^(8 i))
  const Fp* sq_Dh;             // [24][256]  ci^(k 2^(8 i) / 2)   (row 0: even k only)
};

// s^-1 mod n for 1 <= s < n (canonical in, canonical out): s^(n-2)
SPG_HD Fn fn_inv(const Fn& s, const EcdsaTables& T) {
  const uint32_t n[8] = SPG_N_LIMBS;
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = n[i];
  e[0] -= 2;                                   // n is odd and n[0] >= 2: no borrow
  const Fn sm = fn_mul(s, T.r2_n, T.ninv);     // to Montgomery
  Fn one; for (int i = 0; i < 8; i++) one.v[i] = 0; one.v[0] = 1;
  Fn acc = fn_mul(one, T.r2_n, T.ninv);        // R mod n
  for (int i = 251; i >= 0; i--) {
    acc = fn_mul(acc, acc, T.ninv);
    if ((e[i >> 5] >> (i & 31)) & 1) acc = fn_mul(acc, sm, T.ninv);
  }
  return fn_mul(acc, one, T.ninv);             // from Montgomery
}

// ------------------------------------------------------------------ square root (smaller root), p - 1 = 2^192 q
// Returns 0: a has no square root; 1: *y = min(root, p - root) in Montgomery form.
SPG_HD int fp_sqrt_min(const Fp& a, const EcdsaTables& T, Fp* y) {
  if (fp_is_zero(a)) { *y = fp_zero(); return 1; }
  // q = 2^59 + 17;  (q+1)/2 = 2^58 + 9;  w = a^((q-1)/2) = a^(2^58 + 8);  x0 = a w,  t = a w^2 = a^q
  Fp a8 = fp_sqr(fp_sqr(fp_sqr(a)));
  Fp p58 = a8;
  for (int i = 0; i < 55; i++) p58 = fp_sqr(p58);           // a^(2^58)
  const Fp w = fp_mul(p58, a8);
  Fp x = fp_mul(a, w);
  Fp r = fp_mul(x, w);                                       // a^q, in the subgroup of order 2^192
  // discrete logarithm of r to base c, 8 bits at a time from the bottom
  for (int i = 0; i < 24; i++) {
    Fp u = r;
    for (int k = 0; k < 184 - 8 * i; k++) u = fp_sqr(u);
    const Fp uc = fp_reduce(u);
    int e = -1;
    for (int k = 0; k < 256; k++) {
      if (T.sq_L[k].v[0] == uc.v[0] && fp_eq_raw(T.sq_L[k], uc)) { e = k; break; }
    }
    if (e < 0) return 0;                                     // unreachable for field elements
    if (i == 0 && (e & 1)) return 0;                         // odd logarithm: quadratic non-residue
    if (e) {
      r = fp_mul(r, T.sq_D[i * 256 + e]);
      x = fp_mul(x, T.sq_Dh[i * 256 + e]);
    }
  }
  // x^2 == a now; pick the smaller of x and p - x as integers (math_utils.py:47)
  const Fp xc = fp_from_mont(x);
  const Fp neg = fp_reduce(fp_neg(x));
  const Fp nc = fp_from_mont(neg);
  bool x_smaller = true;
  for (int i = 7; i >= 0; i--) { if (xc.v[i] != nc.v[i]) { x_smaller = xc.v[i] < nc.v[i]; break; } }
  *y = x_smaller ? fp_reduce(x) : neg;
  return 1;
}

// ------------------------------------------------------------------ mimic_ec_mult_air (signature.py:176-190)
// fixed base G: partial += bit ? G 2^t : 0 with the x-collision check on every step.  m: canonical scalar.
SPG_HD bool mimic_mult_gen(const uint32_t (&m)[8], const APoint& start, const EcdsaTables& T, JPoint* out) {
  if (u256_is_zero(m)) return false;                       // assert 0 < m
  PedersenAcc a;
  a.init(start);
  // (the set-bit walk of ec.cuh was tried here too: inside this 250-register kernel it made the order pipeline 3.6 % slower,
  // 48.1 vs 46.4 ms for 65536 orders, so the generator walk keeps the step loop)
  bool ok = true;
  for (int t = 0; t < SPG_ECDSA_BITS; t++) {
    const APoint q = T.gen_doubles[t];
    const Fp u2 = fp_mul(q.x, a.zz);
    if (fp_is_zero(fp_sub(u2, a.p.X))) ok = false;         // assert partial_sum.x != point.x
    uint32_t word = m[0];
#pragma unroll
    for (int k = 1; k < 8; k++) word = ((t >> 5) == k) ? m[k] : word;
    if ((word >> (t & 31)) & 1u) {
      a.p = ec_madd_nocheck(a.p, q, a.zzz, u2);
      a.zz = fp_sqr(a.p.Z);
      a.zzz = fp_mul(a.zz, a.p.Z);
    }
  }
  *out = a.p;
  return ok;
}

// variable base: point doubled every step (assert y != 0, math_utils.py:83)
SPG_HD bool mimic_mult_var(const uint32_t (&m)[8], JPoint pt, const APoint& start, JPoint* out) {
  if (u256_is_zero(m)) return false;
  JPoint ps;
  ps.X = start.x; ps.Y = start.y; ps.Z = fp_one();
  bool ok = true;
  for (int t = 0; t < SPG_ECDSA_BITS; t++) {
    const Fp z1z1 = fp_sqr(ps.Z), z2z2 = fp_sqr(pt.Z);
    const Fp u1 = fp_mul(ps.X, z2z2), u2 = fp_mul(pt.X, z1z1);
    if (fp_eq(u1, u2)) ok = false;                          // assert partial_sum.x != point.x
    uint32_t word = m[0];
#pragma unroll
    for (int k = 1; k < 8; k++) word = ((t >> 5) == k) ? m[k] : word;
    if ((word >> (t & 31)) & 1u) ps = ec_jadd_nocheck(ps, pt, z1z1, z2z2, u1, u2);
    if (fp_is_zero(pt.Y)) ok = false;                       // ec_double: assert y != 0
    pt = ec_jdouble_nocheck(pt);
  }
  *out = ps;
  return ok;
}

// Jacobian + Jacobian with the reference's x1 != x2 assertion
SPG_HD bool jadd_checked(const JPoint& p, const JPoint& q, JPoint* out) {
  const Fp z1z1 = fp_sqr(p.Z), z2z2 = fp_sqr(q.Z);
  const Fp u1 = fp_mul(p.X, z2z2), u2 = fp_mul(q.X, z1z1);
  if (fp_eq(u1, u2)) return false;
  *out = ec_jadd_nocheck(p, q, z1z1, z2z2, u1, u2);
  return true;
}

// the body of verify() for a full point key (signature.py:251-260); inputs already range-checked
SPG_HD bool verify_point(const uint32_t (&msg)[8], const uint32_t (&r)[8], const uint32_t (&w)[8], const APoint& Q,
                         const EcdsaTables& T) {
  JPoint zG, rQ, B, wB, fin;
  if (!mimic_mult_gen(msg, T.minus_shift, T, &zG)) return false;
  JPoint q; q.X = Q.x; q.Y = Q.y; q.Z = fp_one();
  if (!mimic_mult_var(r, q, T.shift, &rQ)) return false;
  if (!jadd_checked(zG, rQ, &B)) return false;
  if (!mimic_mult_var(w, B, T.shift, &wB)) return false;
  JPoint ms; ms.X = T.minus_shift.x; ms.Y = T.minus_shift.y; ms.Z = fp_one();
  if (!jadd_checked(wB, ms, &fin)) return false;
  // r == x  <=>  r * Z^2 == X   (r < 2^251 < p, no reduction mod n: signature.py:259-260)
  Fp rr; for (int i = 0; i < 8; i++) rr.v[i] = r[i];
  const Fp rm = fp_to_mont(rr);
  return fp_eq(fp_mul(rm, fp_sqr(fin.Z)), fin.X);
}

// status: 1 valid, 0 invalid (reference returns False), 2 precondition violated (reference raises)
// pub_y == nullptr: x-only key (signature.py:229-238)
SPG_HD int ecdsa_verify_one(const uint32_t (&msg)[8], const uint32_t (&r)[8], const uint32_t (&s)[8],
                            const uint32_t (&px)[8], const uint32_t* py, const EcdsaTables& T) {
  // assert 1 <= s < n ; w = s^-1 mod n ; assert 1 <= r, w < 2^251 ; assert 0 <= msg < 2^251   (:219-227)
  if (u256_is_zero(s) || fn_geq_n(s)) return 2;
  Fn sn; for (int i = 0; i < 8; i++) sn.v[i] = s[i];
  const Fn wn = fn_inv(sn, T);
  uint32_t w[8]; for (int i = 0; i < 8; i++) w[i] = wn.v[i];
  if (u256_is_zero(r) || !u256_lt_2_251(r)) return 2;
  if (u256_is_zero(w) || !u256_lt_2_251(w)) return 2;
  if (!u256_lt_2_251(msg)) return 2;
  // key coordinates must be field elements (the reference's Python ints would silently reduce mod p; keys
  // outside [0, p) are outside its documented domain and are reported as a precondition violation)
  if (spg_canon_geq_p(px) || (py && spg_canon_geq_p(py))) return 2;
  Fp xc; for (int i = 0; i < 8; i++) xc.v[i] = px[i];
  APoint Q;
  if (py) {
    // assert is_point_on_curve (:241) -- with Python ints, so coordinates >= p simply fail the equation
    // unless they are congruent; the reference compares y^2 % p with (x^3 + x + beta) % p
    Fp yc; for (int i = 0; i < 8; i++) yc.v[i] = py[i];
    Q.x = fp_to_mont(xc); Q.y = fp_to_mont(yc);
    const Fp rhs = fp_add(fp_add(fp_mul(fp_sqr(Q.x), Q.x), Q.x), T.beta);
    if (!fp_eq(fp_sqr(Q.y), rhs)) return 2;
    return verify_point(msg, r, w, Q, T) ? 1 : 0;
  }
  Q.x = fp_to_mont(xc);
  const Fp rhs = fp_add(fp_add(fp_mul(fp_sqr(Q.x), Q.x), Q.x), T.beta);
  Fp y;
  if (!fp_sqrt_min(rhs, T, &y)) return 0;                  // InvalidPublicKeyError -> False (:232-235)
  Q.y = y;
  if (verify_point(msg, r, w, Q, T)) return 1;
  Q.y = fp_reduce(fp_neg(y));
  return verify_point(msg, r, w, Q, T) ? 1 : 0;
}

// k * G for 0 < k < n (private_to_stark_key, signature.py:104-110; ec_mult never meets an exceptional case
// for such k).  Returns the affine point (Montgomery).
SPG_HD APoint gen_mult(const uint32_t (&k)[8], const EcdsaTables& T) {
  PedersenAcc a;
  bool started = false;
  for (int t = 0; t < 252; t++) {
    uint32_t word = k[0];
#pragma unroll
    for (int q = 1; q < 8; q++) word = ((t >> 5) == q) ? k[q] : word;
    if (!((word >> (t & 31)) & 1u)) continue;
    const APoint q = T.gen_doubles[t];
    if (!started) { a.init(q); started = true; continue; }
    const Fp u2 = fp_mul(q.x, a.zz);
    a.p = ec_madd_nocheck(a.p, q, a.zzz, u2);
    a.zz = fp_sqr(a.p.Z);
    a.zzz = fp_mul(a.zz, a.p.Z);
  }
  const Fp zi = fp_inv_chain(a.p.Z), zi2 = fp_sqr(zi);
  APoint o;
  o.x = fp_mul(a.p.X, zi2);
  o.y = fp_mul(a.p.Y, fp_mul(zi2, zi));
  return o;
}

// ------------------------------------------------------------------ math_utils.py:59-100, one operation
SPG_HD bool canon_to_mont(const uint32_t (&v)[8], Fp* mont) {
  if (spg_canon_geq_p(v)) return false;
  Fp c; for (int i = 0; i < 8; i++) c.v[i] = v[i];
  *mont = fp_to_mont(c);
  return true;
}
SPG_HD int u256_top_bit(const uint32_t (&m)[8]) {
  for (int k = 7; k >= 0; k--) {
    if (!m[k]) continue;
    int b = 31;
    while (!((m[k] >> b) & 1u)) b--;
    return 32 * k + b;
  }
  return -1;
}
// op 0: R = ec_add((ax, ay), (b0, b1));  op 1: R = ec_double((ax, ay));  op 2: R = ec_mult(b0, (ax, ay)).
// Inputs canonical, R Montgomery affine.  Returns 0 ok; 1 the reference's assertion fails (math_utils.py:64 x1 == x2,
// :84 y == 0); 2 a coordinate >= p; 3 m == 0.  scratch: this thread's column of parked doubles, scratch[k * stride].
SPG_HD int ec_op_one(int op, const uint32_t (&ax)[8], const uint32_t (&ay)[8], const uint32_t (&b0)[8], const uint32_t (&b1)[8],
                     JPoint* scratch, size_t stride, APoint* R) {
  R->x = fp_zero(); R->y = fp_zero();
  APoint P;
  if (!canon_to_mont(ax, &P.x) || !canon_to_mont(ay, &P.y)) return 2;
  if (op == 0) {
    APoint Q;
    if (!canon_to_mont(b0, &Q.x) || !canon_to_mont(b1, &Q.y)) return 2;
    if (fp_eq(P.x, Q.x)) return 1;                        // assert (x1 - x2) % p != 0
    *R = ec_affine_add(P, Q);
    return 0;
  }
  if (op == 1) {
    if (fp_is_zero(P.y)) return 1;                        // assert y % p != 0
    *R = ec_affine_double(P);
    return 0;
  }
  // ec_mult(m, P): m == 1 -> P;  m even -> ec_mult(m / 2, ec_double(P));  m odd -> ec_add(ec_mult(m - 1, P), P).
  // Unrolled: D_k = 2^k P for k up to the top bit (every doubling asserts y != 0), then acc = D_top and, from the
  // highest set bit below the top one down to bit 0, acc = ec_add(acc, D_k) (asserting acc.x != D_k.x).
  const int top = u256_top_bit(b0);
  if (top < 0) return 3;
  JPoint D; D.X = P.x; D.Y = P.y; D.Z = fp_one();
  for (int k = 0; k < top; k++) {
    if ((b0[k >> 5] >> (k & 31)) & 1u) scratch[(size_t)k * stride] = D;
    if (fp_is_zero(D.Y)) return 1;
    D = ec_jdouble_nocheck(D);
  }
  for (int k = top - 1; k >= 0; k--) {
    if (!((b0[k >> 5] >> (k & 31)) & 1u)) continue;
    const JPoint Q = scratch[(size_t)k * stride];
    if (!jadd_checked(D, Q, &D)) return 1;
  }
  const Fp zi = fp_inv_chain(D.Z), zi2 = fp_sqr(zi);
  R->x = fp_mul(D.X, zi2); R->y = fp_mul(D.Y, fp_mul(zi2, zi));
  return 0;
}

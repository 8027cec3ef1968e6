// 252-bit prime-field arithmetic for p = 2^251 + 17*2^192 + 1 (the STARK prime;
// reference: src/starkware/crypto/signature/signature.py:41, nothing_up_my_sleeve_gen.py:35).
//
// Representation: 8 x u32 little-endian limbs (memory layout == 4 x u64 LE), Montgomery form with
// R = 2^256.  Because p ~ 2^251 there are ~5 spare bits, so arithmetic is LAZY:
//   * fp_mul never does a final conditional subtraction: for inputs a, b < 2^254 (~8p) the
//     output is < p + a*b/R < 3p;  for inputs < 4p the output is < 1.5p.
//   * fp_add / fp_sub keep values in [0, 2p) given inputs in [0, 2p).
//   * fp_reduce brings a value < 4p to the canonical range [0, p).
// The prime is sparse: p = 1 + p3 * 2^192 with p3 = 2^59 + 17, so -p^-1 mod 2^64 = -1 and the
// Montgomery reduction needs no general multiplications, only shifts and adds (fp_redc).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SPG_HD __host__ __device__ __forceinline__
#define SPG_D __device__ __forceinline__
#else
#define SPG_HD inline
#define SPG_D inline
#endif

struct alignas(16) Fp {
  uint32_t v[8];
};

// p, 2p in 32-bit limbs
#define SPG_P0 0x00000001u
#define SPG_P6 0x00000011u
#define SPG_P7 0x08000000u
#define SPG_2P0 0x00000002u
#define SPG_2P6 0x00000022u
#define SPG_2P7 0x10000000u

SPG_HD Fp fp_zero() {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}
// R mod p (Montgomery one) = 0x07fffffffffffdf0 ffffffffffffffff ffffffffffffffff ffffffffffffffe1
SPG_HD Fp fp_one() {
  Fp r;
  r.v[0] = 0xffffffe1u; r.v[1] = 0xffffffffu; r.v[2] = 0xffffffffu; r.v[3] = 0xffffffffu;
  r.v[4] = 0xffffffffu; r.v[5] = 0xffffffffu; r.v[6] = 0xfffffdf0u; r.v[7] = 0x07ffffffu;
  return r;
}
// R^2 mod p = 0x07ffd4ab5e008810 ffffffffff6f8000 00000001330fffff fffffd737e000401
SPG_HD Fp fp_r2() {
  Fp r;
  r.v[0] = 0x7e000401u; r.v[1] = 0xfffffd73u; r.v[2] = 0x330fffffu; r.v[3] = 0x00000001u;
  r.v[4] = 0xff6f8000u; r.v[5] = 0xffffffffu; r.v[6] = 0x5e008810u; r.v[7] = 0x07ffd4abu;
  return r;
}

SPG_HD bool fp_eq_raw(const Fp& a, const Fp& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
  return d == 0;
}
SPG_HD bool fp_is_zero_raw(const Fp& a) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i];
  return d == 0;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------ device path (PTX carry chains)

// r = a - k*p if that is >= 0 else a, with kp given by its three non-trivial limbs.
SPG_D Fp fpd_csub(const Fp& a, uint32_t k0, uint32_t k6, uint32_t k7) {
  Fp t;
  uint32_t borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, 0;\n\t"
      "subc.cc.u32 %2, %11, 0;\n\t"
      "subc.cc.u32 %3, %12, 0;\n\t"
      "subc.cc.u32 %4, %13, 0;\n\t"
      "subc.cc.u32 %5, %14, 0;\n\t"
      "subc.cc.u32 %6, %15, %18;\n\t"
      "subc.cc.u32 %7, %16, %19;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=&r"(t.v[0]), "=&r"(t.v[1]), "=&r"(t.v[2]), "=&r"(t.v[3]), "=&r"(t.v[4]), "=&r"(t.v[5]),
        "=&r"(t.v[6]), "=&r"(t.v[7]), "=&r"(borrow)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(k0), "r"(k6), "r"(k7));
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : t.v[i];
  return r;
}

// a + b, inputs < 2p  -> output < 2p
SPG_D Fp fpd_add(const Fp& a, const Fp& b) {
  Fp s;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=&r"(s.v[0]), "=&r"(s.v[1]), "=&r"(s.v[2]), "=&r"(s.v[3]), "=&r"(s.v[4]), "=&r"(s.v[5]),
        "=&r"(s.v[6]), "=&r"(s.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]));
  return fpd_csub(s, SPG_2P0, SPG_2P6, SPG_2P7);
}
// a + b with no reduction (caller guarantees the sum stays < 2^256 and within the lazy bounds)
SPG_D Fp fpd_add_raw(const Fp& a, const Fp& b) {
  Fp s;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=&r"(s.v[0]), "=&r"(s.v[1]), "=&r"(s.v[2]), "=&r"(s.v[3]), "=&r"(s.v[4]), "=&r"(s.v[5]),
        "=&r"(s.v[6]), "=&r"(s.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]));
  return s;
}

// a - b (mod p), inputs < 2p -> output < 2p : computes a - b, adds 2p back on borrow.
SPG_D Fp fpd_sub(const Fp& a, const Fp& b) {
  Fp d;
  uint32_t borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=&r"(d.v[0]), "=&r"(d.v[1]), "=&r"(d.v[2]), "=&r"(d.v[3]), "=&r"(d.v[4]), "=&r"(d.v[5]),
        "=&r"(d.v[6]), "=&r"(d.v[7]), "=&r"(borrow)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]));
  // borrow is 0 or 0xffffffff: add (2p & borrow)
  uint32_t k0 = borrow & SPG_2P0, k6 = borrow & SPG_2P6, k7 = borrow & SPG_2P7;
  Fp r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, 0;\n\t"
      "addc.cc.u32 %2, %10, 0;\n\t"
      "addc.cc.u32 %3, %11, 0;\n\t"
      "addc.cc.u32 %4, %12, 0;\n\t"
      "addc.cc.u32 %5, %13, 0;\n\t"
      "addc.cc.u32 %6, %14, %17;\n\t"
      "addc.u32 %7, %15, %18;"
      : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
        "=&r"(r.v[6]), "=&r"(r.v[7])
      : "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]),
        "r"(d.v[7]), "r"(k0), "r"(k6), "r"(k7));
  return r;
}

// value < 4p -> canonical [0, p)
SPG_D Fp fpd_reduce(const Fp& a) {
  Fp r = fpd_csub(a, SPG_2P0, SPG_2P6, SPG_2P7);
  return fpd_csub(r, SPG_P0, SPG_P6, SPG_P7);
}

// a - b + K*p with no conditional step: requires b <= K*p and a + K*p < 2^256 (K a small compile-time integer)
SPG_D Fp fpd_sub_lazy(const Fp& a, const Fp& b, uint32_t K) {
  Fp d, r;
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;"
      : "=&r"(d.v[0]), "=&r"(d.v[1]), "=&r"(d.v[2]), "=&r"(d.v[3]), "=&r"(d.v[4]), "=&r"(d.v[5]),
        "=&r"(d.v[6]), "=&r"(d.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]));
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, 0;\n\t"
      "addc.cc.u32 %2, %10, 0;\n\t"
      "addc.cc.u32 %3, %11, 0;\n\t"
      "addc.cc.u32 %4, %12, 0;\n\t"
      "addc.cc.u32 %5, %13, 0;\n\t"
      "addc.cc.u32 %6, %14, %17;\n\t"
      "addc.u32 %7, %15, %18;"
      : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
        "=&r"(r.v[6]), "=&r"(r.v[7])
      : "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]),
        "r"(d.v[7]), "r"(K * SPG_P0), "r"(K * SPG_P6), "r"(K * SPG_P7));
  return r;
}

// a - q*p - (optionally) nothing else, q < 32, returning the borrow (0 / 0xffffffff) in *borrow
SPG_D Fp fpd_sub_qp(const Fp& a, uint32_t q, uint32_t* borrow) {
  Fp r;
  uint32_t bo;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, 0;\n\t"
      "subc.cc.u32 %2, %11, 0;\n\t"
      "subc.cc.u32 %3, %12, 0;\n\t"
      "subc.cc.u32 %4, %13, 0;\n\t"
      "subc.cc.u32 %5, %14, 0;\n\t"
      "subc.cc.u32 %6, %15, %18;\n\t"
      "subc.cc.u32 %7, %16, %19;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
        "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(bo)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]), "r"(q), "r"(q * SPG_P6), "r"(q << 27));
  *borrow = bo;
  return r;
}
// any 256-bit value -> an equivalent one below 2^252 (< 2p), 12 integer instructions, no conditional step:
// subtracts (q - 1) * p for q = floor(a / 2^251) >= 1
SPG_D Fp fpd_partial(const Fp& a) {
  uint32_t q = a.v[7] >> 27, bo;
  q -= (q != 0u);
  return fpd_sub_qp(a, q, &bo);
}
// any 256-bit value -> canonical [0, p): subtract floor(a / 2^251) * p, add p back if that went negative
SPG_D Fp fpd_reduce_full(const Fp& a) {
  uint32_t bo;
  Fp d = fpd_sub_qp(a, a.v[7] >> 27, &bo);
  Fp r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, 0;\n\t"
      "addc.cc.u32 %2, %10, 0;\n\t"
      "addc.cc.u32 %3, %11, 0;\n\t"
      "addc.cc.u32 %4, %12, 0;\n\t"
      "addc.cc.u32 %5, %13, 0;\n\t"
      "addc.cc.u32 %6, %14, %17;\n\t"
      "addc.u32 %7, %15, %18;"
      : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
        "=&r"(r.v[6]), "=&r"(r.v[7])
      : "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]),
        "r"(d.v[7]), "r"(bo & SPG_P0), "r"(bo & SPG_P6), "r"(bo & SPG_P7));
  return r;
}

// ---- 8x8 limb schoolbook product, even/odd column accumulators so that every mad.lo/mad.hi pair
// lands on an aligned 64-bit accumulator (ptxas fuses each pair into one IMAD.WIDE.U32[.X]).
#define SPG_ROW_MUL(acc, k, A0, A1, A2, A3, B)                                         \
  asm("{\n\t"                                                                           \
      ".reg .u64 w0, w1, w2, w3;\n\t"                                                   \
      "mul.wide.u32 w0, %8, %12;\n\t"                                                   \
      "mul.wide.u32 w1, %9, %12;\n\t"                                                   \
      "mul.wide.u32 w2, %10, %12;\n\t"                                                  \
      "mul.wide.u32 w3, %11, %12;\n\t"                                                  \
      "mov.b64 {%0, %1}, w0;\n\t"                                                       \
      "mov.b64 {%2, %3}, w1;\n\t"                                                       \
      "mov.b64 {%4, %5}, w2;\n\t"                                                       \
      "mov.b64 {%6, %7}, w3;\n\t"                                                       \
      "}"                                                                               \
      : "=&r"(acc[k]), "=&r"(acc[k + 1]), "=&r"(acc[k + 2]), "=&r"(acc[k + 3]), "=&r"(acc[k + 4]), \
        "=&r"(acc[k + 5]), "=&r"(acc[k + 6]), "=&r"(acc[k + 7])                             \
      : "r"(A0), "r"(A1), "r"(A2), "r"(A3), "r"(B))

#define SPG_ROW_MAD(acc, k, A0, A1, A2, A3, B)                                         \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                              \
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                             \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                            \
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                            \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                            \
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                            \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                            \
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                            \
      "addc.u32 %8, %8, 0;"                                                             \
      : "+r"(acc[k]), "+r"(acc[k + 1]), "+r"(acc[k + 2]), "+r"(acc[k + 3]), "+r"(acc[k + 4]), \
        "+r"(acc[k + 5]), "+r"(acc[k + 6]), "+r"(acc[k + 7]), "+r"(acc[k + 8])           \
      : "r"(A0), "r"(A1), "r"(A2), "r"(A3), "r"(B))

// Opaque copies of the two non-trivial high limbs of p (p6 = 0x11, p7 = 2^27).  Read from constant memory so
// that ptxas keeps the reduction's multiplications by them on the IMAD (FMA) pipe instead of strength-reducing
// them to shift/add sequences on the already saturated ALU pipe.
static __constant__ uint32_t spg_pk[2] = {SPG_P6, SPG_P7};

// t[0..15] = a * b + p * 2^256   (the p * 2^256 seed costs nothing -- it is the initial value of three
// accumulators -- and makes the subtractive reduction below non-negative)
SPG_D void fpd_mul_wide(uint32_t (&t)[16], const Fp& a, const Fp& b) {
  uint32_t E[17], O[16];
#pragma unroll
  for (int i = 8; i < 17; i++) E[i] = 0;
#pragma unroll
  for (int i = 8; i < 16; i++) O[i] = 0;
  E[8] = SPG_P0; E[14] = SPG_P6; E[15] = SPG_P7;
  // E[k] has weight 2^(32k); O[k] has weight 2^(32(k+1)).
  SPG_ROW_MUL(E, 0, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
  SPG_ROW_MUL(O, 0, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    if (i & 1) {
      SPG_ROW_MAD(E, i + 1, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
      SPG_ROW_MAD(O, i - 1, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
    } else {
      SPG_ROW_MAD(E, i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
      SPG_ROW_MAD(O, i, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
    }
  }
  t[0] = E[0];
  asm("add.cc.u32 %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32 %14, %29, %44;"
      : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]),
        "=&r"(t[8]), "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11]), "=&r"(t[12]), "=&r"(t[13]), "=&r"(t[14]),
        "=&r"(t[15])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]),
        "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
        "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]),
        "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
}

// ---- dedicated squaring: 28 cross products + 8 diagonal ones = 36 wide multiplies instead of 64.
// Chains of 3, 2 and 1 products (the 4-product chain is SPG_ROW_MAD); each ends with its carry into the next limb,
// which at that point holds nothing but earlier carries (chains run in ascending order of their multiplier limb).
#define SPG_CHAIN3(acc, k, A0, A1, A2, B)                                               \
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"                                              \
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"                                             \
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"                                             \
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"                                             \
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"                                             \
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"                                             \
      "addc.u32 %6, %6, 0;"                                                             \
      : "+r"(acc[k]), "+r"(acc[k + 1]), "+r"(acc[k + 2]), "+r"(acc[k + 3]), "+r"(acc[k + 4]), \
        "+r"(acc[k + 5]), "+r"(acc[k + 6])                                               \
      : "r"(A0), "r"(A1), "r"(A2), "r"(B))
#define SPG_CHAIN2(acc, k, A0, A1, B)                                                   \
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"                                               \
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"                                              \
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"                                              \
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"                                              \
      "addc.u32 %4, %4, 0;"                                                             \
      : "+r"(acc[k]), "+r"(acc[k + 1]), "+r"(acc[k + 2]), "+r"(acc[k + 3]), "+r"(acc[k + 4]) \
      : "r"(A0), "r"(A1), "r"(B))
#define SPG_CHAIN1(acc, k, A0, B)                                                       \
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"                                               \
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"                                              \
      "addc.u32 %2, %2, 0;"                                                             \
      : "+r"(acc[k]), "+r"(acc[k + 1]), "+r"(acc[k + 2])                                 \
      : "r"(A0), "r"(B))

// t[0..15] = a * a + p * 2^256 for a < 2^254: exactly what fpd_mul_wide(t, a, a) produces.
//   X = sum_{i<j} a_i a_j 2^(32(i+j))  on the even/odd accumulators (E: i+j even, O: i+j odd),
//   t = 2 X + sum_i (a_i^2 + seed_i) 2^(64 i),  seed = the limbs of p * 2^256 (they fit under the squares: a_7 < 2^30).
SPG_D void fpd_sqr_wide(uint32_t (&t)[16], const Fp& a) {
  uint32_t E[16], O[15];
#pragma unroll
  for (int i = 0; i < 16; i++) E[i] = 0;
#pragma unroll
  for (int i = 0; i < 15; i++) O[i] = 0;
  // E[k] has weight 2^(32k); O[k] has weight 2^(32(k+1)).
  SPG_ROW_MAD(O, 0, a.v[1], a.v[3], a.v[5], a.v[7], a.v[0]);
  SPG_CHAIN3(E, 2, a.v[2], a.v[4], a.v[6], a.v[0]);
  SPG_CHAIN3(O, 2, a.v[2], a.v[4], a.v[6], a.v[1]);
  SPG_CHAIN3(E, 4, a.v[3], a.v[5], a.v[7], a.v[1]);
  SPG_CHAIN3(O, 4, a.v[3], a.v[5], a.v[7], a.v[2]);
  SPG_CHAIN2(E, 6, a.v[4], a.v[6], a.v[2]);
  SPG_CHAIN2(O, 6, a.v[4], a.v[6], a.v[3]);
  SPG_CHAIN2(E, 8, a.v[5], a.v[7], a.v[3]);
  SPG_CHAIN2(O, 8, a.v[5], a.v[7], a.v[4]);
  SPG_CHAIN1(E, 10, a.v[6], a.v[4]);
  SPG_CHAIN1(O, 10, a.v[6], a.v[5]);
  SPG_CHAIN1(E, 12, a.v[7], a.v[5]);
  SPG_CHAIN1(O, 12, a.v[7], a.v[6]);
  // X[1] = O[0];  X[k] = E[k] + O[k-1], k = 2..15
  uint32_t X[16];
  X[1] = O[0];
  asm("add.cc.u32 %0, %14, %28;\n\t"
      "addc.cc.u32 %1, %15, %29;\n\t"
      "addc.cc.u32 %2, %16, %30;\n\t"
      "addc.cc.u32 %3, %17, %31;\n\t"
      "addc.cc.u32 %4, %18, %32;\n\t"
      "addc.cc.u32 %5, %19, %33;\n\t"
      "addc.cc.u32 %6, %20, %34;\n\t"
      "addc.cc.u32 %7, %21, %35;\n\t"
      "addc.cc.u32 %8, %22, %36;\n\t"
      "addc.cc.u32 %9, %23, %37;\n\t"
      "addc.cc.u32 %10, %24, %38;\n\t"
      "addc.cc.u32 %11, %25, %39;\n\t"
      "addc.cc.u32 %12, %26, %40;\n\t"
      "addc.u32 %13, %27, %41;"
      : "=&r"(X[2]), "=&r"(X[3]), "=&r"(X[4]), "=&r"(X[5]), "=&r"(X[6]), "=&r"(X[7]), "=&r"(X[8]), "=&r"(X[9]),
        "=&r"(X[10]), "=&r"(X[11]), "=&r"(X[12]), "=&r"(X[13]), "=&r"(X[14]), "=&r"(X[15])
      : "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]),
        "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]),
        "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
  // 2 X
  uint32_t Y[16];
  Y[1] = X[1] << 1;
#pragma unroll
  for (int k = 2; k < 16; k++) Y[k] = __funnelshift_l(X[k - 1], X[k], 1);
  // diagonal squares with the seed limbs of p * 2^256 (limb 8: p0 = 1; limbs 14, 15: p6, p7) as addends
  uint32_t D[16];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint64_t seed = (i == 4) ? (uint64_t)SPG_P0 : (i == 7) ? (((uint64_t)SPG_P7 << 32) | SPG_P6) : 0ull;
    uint64_t w;
    asm("mad.wide.u32 %0, %1, %1, %2;" : "=l"(w) : "r"(a.v[i]), "l"(seed));
    D[2 * i] = (uint32_t)w; D[2 * i + 1] = (uint32_t)(w >> 32);
  }
  t[0] = D[0];
  asm("add.cc.u32 %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32 %14, %29, %44;"
      : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]),
        "=&r"(t[8]), "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11]), "=&r"(t[12]), "=&r"(t[13]), "=&r"(t[14]),
        "=&r"(t[15])
      : "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]), "r"(Y[8]),
        "r"(Y[9]), "r"(Y[10]), "r"(Y[11]), "r"(Y[12]), "r"(Y[13]), "r"(Y[14]), "r"(Y[15]),
        "r"(D[1]), "r"(D[2]), "r"(D[3]), "r"(D[4]), "r"(D[5]), "r"(D[6]), "r"(D[7]), "r"(D[8]),
        "r"(D[9]), "r"(D[10]), "r"(D[11]), "r"(D[12]), "r"(D[13]), "r"(D[14]), "r"(D[15]));
}

SPG_D uint32_t spg_lo32(uint64_t x) { return (uint32_t)x; }
SPG_D uint32_t spg_hi32(uint64_t x) { return (uint32_t)(x >> 32); }
SPG_D uint64_t spg_mulw(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
SPG_D uint64_t spg_madw(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c; }

// Montgomery reduction of T' = T + p * 2^256 (T < 2^508, as produced by fpd_mul_wide): returns
// T * 2^-256 mod p, lazily (0 < result <= p + T / 2^256).  With p = 1 + p3 * 2^192, p3 = 2^59 + 17 = p6 + p7 * 2^32,
// p^-1 = 1 - p3 * 2^192 (mod 2^256):
//   V  = p3 * T[0..63] mod 2^64
//   M  = T[0..255] * p^-1 mod 2^256 = T[0..191] | ((T[192..255] - V) mod 2^64) << 192 ,  c = borrow of that
//   U  = p3 * M                                        (320 bits; its low 64 bits equal V)
//   result = T'[256..511] - c - (U >> 64)              (M * p == T mod 2^256, so the low halves cancel exactly)
// U is an 8 x 2 limb schoolbook on the IMAD pipe; every 64-bit column sum is below 2^60, so no carry chains are
// needed until the single merge of the even and odd columns.
SPG_D Fp fpd_redc_imad(const uint32_t (&t)[16]) {
  const uint32_t p6 = spg_pk[0], p7 = spg_pk[1];
  // low 96 bits of U: limbs 0 and 1 are V, limb 2 carries into the rest
  const uint64_t E0 = spg_mulw(t[0], p6);                       // weight 2^0
  const uint64_t O0 = spg_madw(t[0], p7, spg_mulw(t[1], p6));   // weight 2^32
  const uint32_t v0 = spg_lo32(E0);
  uint32_t v1, q2, w6, w7, nb;
  asm("add.cc.u32 %0, %2, %3;\n\t"
      "addc.u32 %1, %4, 0;"
      : "=&r"(v1), "=&r"(q2)
      : "r"(spg_hi32(E0)), "r"(spg_lo32(O0)), "r"(spg_hi32(O0)));
  // (w6, w7) = T[192..255] - V, borrow c folded into q2 (it has the weight of U's limb 2 = result limb 0)
  asm("sub.cc.u32 %0, %3, %5;\n\t"
      "subc.cc.u32 %1, %4, %6;\n\t"
      "subc.u32 %2, 0, 0;"
      : "=&r"(w6), "=&r"(w7), "=&r"(nb)
      : "r"(t[6]), "r"(t[7]), "r"(v0), "r"(v1));
  q2 -= nb;   // nb is 0 or 0xffffffff (= -c)
  const uint64_t E1 = spg_madw(t[1], p7, spg_mulw(t[2], p6));   // weight 2^64
  const uint64_t O1 = spg_madw(t[2], p7, spg_mulw(t[3], p6));
  const uint64_t E2 = spg_madw(t[3], p7, spg_mulw(t[4], p6));
  const uint64_t O2 = spg_madw(t[4], p7, spg_mulw(t[5], p6));
  const uint64_t E3 = spg_madw(t[5], p7, spg_mulw(w6, p6));
  const uint64_t O3 = spg_madw(w6, p7, spg_mulw(w7, p6));
  const uint64_t E4 = spg_mulw(w7, p7);
  // Y = (U >> 64) + c
  uint32_t Y[8];
  asm("add.cc.u32 %0, %8, %9;\n\t"
      "addc.cc.u32 %1, %10, %11;\n\t"
      "addc.cc.u32 %2, %12, %13;\n\t"
      "addc.cc.u32 %3, %14, %15;\n\t"
      "addc.cc.u32 %4, %16, %17;\n\t"
      "addc.cc.u32 %5, %18, %19;\n\t"
      "addc.cc.u32 %6, %20, %21;\n\t"
      "addc.u32 %7, %22, 0;"
      : "=&r"(Y[0]), "=&r"(Y[1]), "=&r"(Y[2]), "=&r"(Y[3]), "=&r"(Y[4]), "=&r"(Y[5]), "=&r"(Y[6]), "=&r"(Y[7])
      : "r"(q2), "r"(spg_lo32(E1)), "r"(spg_hi32(E1)), "r"(spg_lo32(O1)), "r"(spg_hi32(O1)), "r"(spg_lo32(E2)),
        "r"(spg_hi32(E2)), "r"(spg_lo32(O2)), "r"(spg_hi32(O2)), "r"(spg_lo32(E3)), "r"(spg_hi32(E3)),
        "r"(spg_lo32(O3)), "r"(spg_hi32(O3)), "r"(spg_lo32(E4)), "r"(spg_hi32(E4)));
  Fp r;
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;"
      : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
        "=&r"(r.v[6]), "=&r"(r.v[7])
      : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]),
        "r"(Y[0]), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
  return r;
}

// The same reduction with U = 17 * M + (M << 59) built from funnel shifts and additions only: no IMAD.WIDE at all
// (the fmaheavy pipe, which executes IMAD.WIDE at one warp instruction per 4 cycles, is the binding resource of
// every multiplication-heavy kernel; the ALU pipe runs these at one per 2 cycles and has head-room).
SPG_D Fp fpd_redc_shift(const uint32_t (&t)[16]) {
  uint32_t v0, v1;
  {
    uint32_t a0 = t[0] << 4, a1 = __funnelshift_l(t[0], t[1], 4), b1 = t[0] << 27;
    asm("add.cc.u32 %0, %2, %4;\n\t"
        "addc.u32 %1, %3, %5;\n\t"
        "add.u32 %1, %1, %6;"
        : "=&r"(v0), "=&r"(v1)
        : "r"(a0), "r"(a1), "r"(t[0]), "r"(t[1]), "r"(b1));
  }
  uint32_t w[8], nb;
#pragma unroll
  for (int i = 0; i < 6; i++) w[i] = t[i];
  asm("sub.cc.u32 %0, %3, %5;\n\t"
      "subc.cc.u32 %1, %4, %6;\n\t"
      "subc.u32 %2, 0, 0;"
      : "=&r"(w[6]), "=&r"(w[7]), "=&r"(nb)
      : "r"(t[6]), "r"(t[7]), "r"(v0), "r"(v1));
  // A = 17 * M = (M << 4) + M   (9 limbs)
  uint32_t s4[9];
  s4[0] = w[0] << 4;
#pragma unroll
  for (int i = 1; i < 8; i++) s4[i] = __funnelshift_l(w[i - 1], w[i], 4);
  s4[8] = w[7] >> 28;
  uint32_t A[9];
  asm("add.cc.u32 %0, %9, %18;\n\t"
      "addc.cc.u32 %1, %10, %19;\n\t"
      "addc.cc.u32 %2, %11, %20;\n\t"
      "addc.cc.u32 %3, %12, %21;\n\t"
      "addc.cc.u32 %4, %13, %22;\n\t"
      "addc.cc.u32 %5, %14, %23;\n\t"
      "addc.cc.u32 %6, %15, %24;\n\t"
      "addc.cc.u32 %7, %16, %25;\n\t"
      "addc.u32 %8, %17, 0;"
      : "=&r"(A[0]), "=&r"(A[1]), "=&r"(A[2]), "=&r"(A[3]), "=&r"(A[4]), "=&r"(A[5]), "=&r"(A[6]),
        "=&r"(A[7]), "=&r"(A[8])
      : "r"(s4[0]), "r"(s4[1]), "r"(s4[2]), "r"(s4[3]), "r"(s4[4]), "r"(s4[5]), "r"(s4[6]), "r"(s4[7]),
        "r"(s4[8]), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]));
  // B = M << 59 : limbs B[1..9];  Y = ((A + B) >> 64) + c, the borrow c = -nb entering at limb 2 of the sum
  uint32_t B[10];
  B[1] = w[0] << 27;
#pragma unroll
  for (int i = 1; i < 8; i++) B[i + 1] = __funnelshift_l(w[i - 1], w[i], 27);
  B[9] = w[7] >> 5;
  uint32_t U1, Y[8];
  asm("add.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, %25;"
      : "=&r"(U1), "=&r"(Y[0]), "=&r"(Y[1]), "=&r"(Y[2]), "=&r"(Y[3]), "=&r"(Y[4]), "=&r"(Y[5]), "=&r"(Y[6]),
        "=&r"(Y[7])
      : "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(A[8]), "r"(B[1]),
        "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]), "r"(B[8]), "r"(B[9]));
  (void)U1;
  // result = T'_hi - Y - c : c enters the subtraction chain as its initial borrow (0 - nb borrows exactly when
  // nb != 0).  Only sub-family instructions touch the flag here: ptxas does not convert an add.cc carry into a
  // subc borrow correctly (measured: tools/mulbench.cu).
  Fp r;
  uint32_t sink;
  asm("sub.cc.u32 %8, 0, %25;\n\t"
      "subc.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.u32 %7, %16, %24;"
      : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
        "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(sink)
      : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]),
        "r"(Y[0]), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]), "r"(nb));
  (void)sink;
  return r;
}

// Hybrid: 17 * M on the IMAD pipe (8 IMAD.WIDE, no carries), M << 59 by funnel shifts.
SPG_D Fp fpd_redc_hybrid(const uint32_t (&t)[16]) {
  const uint32_t p6 = spg_pk[0];
  uint32_t v0, v1;
  {
    uint32_t a0 = t[0] << 4, a1 = __funnelshift_l(t[0], t[1], 4), b1 = t[0] << 27;
    asm("add.cc.u32 %0, %2, %4;\n\t"
        "addc.u32 %1, %3, %5;\n\t"
        "add.u32 %1, %1, %6;"
        : "=&r"(v0), "=&r"(v1)
        : "r"(a0), "r"(a1), "r"(t[0]), "r"(t[1]), "r"(b1));
  }
  uint32_t w[8], nb;
#pragma unroll
  for (int i = 0; i < 6; i++) w[i] = t[i];
  asm("sub.cc.u32 %0, %3, %5;\n\t"
      "subc.cc.u32 %1, %4, %6;\n\t"
      "subc.u32 %2, 0, 0;"
      : "=&r"(w[6]), "=&r"(w[7]), "=&r"(nb)
      : "r"(t[6]), "r"(t[7]), "r"(v0), "r"(v1));
  // 17 * M = sum_k Ek * 2^(64k) + sum_k Ok * 2^(64k+32), each 64-bit term below 2^37
  uint64_t E[4], O[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { E[k] = spg_mulw(w[2 * k], p6); O[k] = spg_mulw(w[2 * k + 1], p6); }
  // B = M << 59: limbs B[1..9]
  uint32_t B[10];
  B[1] = w[0] << 27;
#pragma unroll
  for (int i = 1; i < 8; i++) B[i + 1] = __funnelshift_l(w[i - 1], w[i], 27);
  B[9] = w[7] >> 5;
  // S = E + B  (limbs 1..9 ; limb 0 of U is not needed), then Y = ((S + O << 32) >> 64) + c
  uint32_t S[10];
  asm("add.cc.u32 %0, %9, %16;\n\t"
      "addc.cc.u32 %1, %10, %17;\n\t"
      "addc.cc.u32 %2, %11, %18;\n\t"
      "addc.cc.u32 %3, %12, %19;\n\t"
      "addc.cc.u32 %4, %13, %20;\n\t"
      "addc.cc.u32 %5, %14, %21;\n\t"
      "addc.cc.u32 %6, %15, %22;\n\t"
      "addc.cc.u32 %7, 0, %23;\n\t"
      "addc.u32 %8, 0, %24;"
      : "=&r"(S[1]), "=&r"(S[2]), "=&r"(S[3]), "=&r"(S[4]), "=&r"(S[5]), "=&r"(S[6]), "=&r"(S[7]), "=&r"(S[8]), "=&r"(S[9])
      : "r"(spg_hi32(E[0])), "r"(spg_lo32(E[1])), "r"(spg_hi32(E[1])), "r"(spg_lo32(E[2])), "r"(spg_hi32(E[2])),
        "r"(spg_lo32(E[3])), "r"(spg_hi32(E[3])),
        "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]), "r"(B[8]), "r"(B[9]));
  uint32_t U1, Y[8];
  asm("add.cc.u32 %0, %9, %18;\n\t"
      "addc.cc.u32 %1, %10, %19;\n\t"
      "addc.cc.u32 %2, %11, %20;\n\t"
      "addc.cc.u32 %3, %12, %21;\n\t"
      "addc.cc.u32 %4, %13, %22;\n\t"
      "addc.cc.u32 %5, %14, %23;\n\t"
      "addc.cc.u32 %6, %15, %24;\n\t"
      "addc.cc.u32 %7, %16, %25;\n\t"
      "addc.u32 %8, %17, 0;"
      : "=&r"(U1), "=&r"(Y[0]), "=&r"(Y[1]), "=&r"(Y[2]), "=&r"(Y[3]), "=&r"(Y[4]), "=&r"(Y[5]), "=&r"(Y[6]), "=&r"(Y[7])
      : "r"(S[1]), "r"(S[2]), "r"(S[3]), "r"(S[4]), "r"(S[5]), "r"(S[6]), "r"(S[7]), "r"(S[8]), "r"(S[9]),
        "r"(spg_lo32(O[0])), "r"(spg_hi32(O[0])), "r"(spg_lo32(O[1])), "r"(spg_hi32(O[1])), "r"(spg_lo32(O[2])),
        "r"(spg_hi32(O[2])), "r"(spg_lo32(O[3])), "r"(spg_hi32(O[3])));
  (void)U1;
  Fp r;
  uint32_t sink;
  asm("sub.cc.u32 %8, 0, %25;\n\t"
      "subc.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.u32 %7, %16, %24;"
      : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
        "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(sink)
      : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]),
        "r"(Y[0]), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]), "r"(nb));
  (void)sink;
  return r;
}

#ifndef SPG_REDC
#define SPG_REDC 1     // 0: IMAD-pipe reduction, 1: shift/add reduction, 2: hybrid (measured in tools/mulbench.cu)
#endif
SPG_D Fp fpd_redc(const uint32_t (&t)[16]) {
#if SPG_REDC == 0
  return fpd_redc_imad(t);
#elif SPG_REDC == 1
  return fpd_redc_shift(t);
#else
  return fpd_redc_hybrid(t);
#endif
}

SPG_D Fp fpd_mul(const Fp& a, const Fp& b) {
  uint32_t t[16];
  fpd_mul_wide(t, a, b);
  return fpd_redc(t);
}
SPG_D Fp fpd_sqr(const Fp& a) {
#if defined(SPG_SQR_AS_MUL)
  return fpd_mul(a, a);
#else
  uint32_t t[16];
  fpd_sqr_wide(t, a);
  return fpd_redc(t);
#endif
}

#endif  // __CUDACC__

// ------------------------------------------------------------------ host path (unsigned __int128)
typedef unsigned __int128 spg_u128;

static inline void fp_to_u64(const Fp& a, uint64_t* w) {
  for (int i = 0; i < 4; i++) w[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
static inline Fp fp_from_u64(const uint64_t* w) {
  Fp r;
  for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)w[i]; r.v[2 * i + 1] = (uint32_t)(w[i] >> 32); }
  return r;
}
static const uint64_t SPG_P64[4] = {1ull, 0ull, 0ull, 0x0800000000000011ull};

static inline bool u256_geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
  return true;
}
static inline void u256_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) {
    spg_u128 d = (spg_u128)a[i] - b[i] - br;
    r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
  }
}
static inline void u256_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) {
    spg_u128 s = (spg_u128)a[i] + b[i] + c;
    r[i] = (uint64_t)s; c = (uint64_t)(s >> 64);
  }
}
// host functions always return canonical values
inline Fp fph_reduce(const Fp& a) {
  uint64_t w[4]; fp_to_u64(a, w);
  while (u256_geq(w, SPG_P64)) u256_sub(w, w, SPG_P64);
  return fp_from_u64(w);
}
inline Fp fph_add(const Fp& a, const Fp& b) {
  uint64_t x[4], y[4]; fp_to_u64(a, x); fp_to_u64(b, y);
  u256_add(x, x, y);
  return fph_reduce(fp_from_u64(x));
}
inline Fp fph_sub(const Fp& a, const Fp& b) {
  uint64_t x[4], y[4]; fp_to_u64(fph_reduce(a), x); fp_to_u64(fph_reduce(b), y);
  if (!u256_geq(x, y)) u256_add(x, x, SPG_P64);
  u256_sub(x, x, y);
  return fp_from_u64(x);
}
inline Fp fph_mul(const Fp& a, const Fp& b) {
  uint64_t x[4], y[4], t[9] = {0};
  fp_to_u64(a, x); fp_to_u64(b, y);
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      spg_u128 s = (spg_u128)x[j] * y[i] + t[i + j] + c;
      t[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64);
    }
    t[i + 4] = c;
  }
  // word-serial Montgomery, m = -t[k] (since -p^-1 = -1 mod 2^64)
  for (int k = 0; k < 4; k++) {
    uint64_t m = (uint64_t)0 - t[k];
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      spg_u128 s = (spg_u128)m * SPG_P64[j] + t[k + j] + c;
      t[k + j] = (uint64_t)s; c = (uint64_t)(s >> 64);
    }
    for (int j = k + 4; c && j < 9; j++) {
      spg_u128 s = (spg_u128)t[j] + c;
      t[j] = (uint64_t)s; c = (uint64_t)(s >> 64);
    }
  }
  return fph_reduce(fp_from_u64(t + 4));
}
inline Fp fph_sqr(const Fp& a) { return fph_mul(a, a); }

// Lazy operations on the host.  Default build: canonical results (same residues as the device).  With
// -DSPG_EMUL_LAZY (tests/host_emul) they reproduce the DEVICE representatives bit for bit and abort when a
// lazy bound is violated (overflow of 2^256 / 2^512 or a negative intermediate), so the bound bookkeeping of
// the NTT butterflies is checked on the CPU.
#if defined(SPG_EMUL_LAZY)
#include <stdio.h>
#include <stdlib.h>
static inline void spg_lazy_fail(const char* what) { fprintf(stderr, "lazy bound violated: %s\n", what); abort(); }
static inline void u256_mul_small(uint64_t* r5, const uint64_t* a, uint64_t k) {
  spg_u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (spg_u128)a[i] * k; r5[i] = (uint64_t)c; c >>= 64; }
  r5[4] = (uint64_t)c;
}
inline Fp fph_add_raw(const Fp& a, const Fp& b) {
  uint64_t x[4], y[4]; fp_to_u64(a, x); fp_to_u64(b, y);
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) { spg_u128 s = (spg_u128)x[i] + y[i] + c; x[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
  if (c) spg_lazy_fail("add_raw overflow");
  return fp_from_u64(x);
}
inline Fp fph_sub_lazy(const Fp& a, const Fp& b, uint32_t K) {
  uint64_t x[4], y[4], kp[5]; fp_to_u64(a, x); fp_to_u64(b, y);
  u256_mul_small(kp, SPG_P64, K);
  if (kp[4]) spg_lazy_fail("K*p overflow");
  if (!u256_geq(kp, y)) spg_lazy_fail("sub_lazy: b > K*p");
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) { spg_u128 s = (spg_u128)x[i] + kp[i] + c; x[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
  if (c) spg_lazy_fail("sub_lazy: a + K*p overflow");
  u256_sub(x, x, y);
  return fp_from_u64(x);
}
inline Fp fph_partial(const Fp& a) {
  uint64_t x[4], qp[5]; fp_to_u64(a, x);
  uint64_t q = x[3] >> 59;
  if (q) q -= 1;
  u256_mul_small(qp, SPG_P64, q);
  if (!u256_geq(x, qp)) spg_lazy_fail("partial negative");
  u256_sub(x, x, qp);
  return fp_from_u64(x);
}
inline Fp fph_reduce_full(const Fp& a) {
  uint64_t x[4]; fp_to_u64(a, x);
  uint64_t q = x[3] >> 59, qp[5];
  u256_mul_small(qp, SPG_P64, q);
  if (u256_geq(x, qp)) u256_sub(x, x, qp);
  else { u256_add(x, x, SPG_P64); u256_sub(x, x, qp); }
  if (u256_geq(x, SPG_P64)) spg_lazy_fail("reduce_full not canonical");
  return fp_from_u64(x);
}
// exact device representative: (T + p*2^256 - M*p) / 2^256 with M = T * p^-1 mod 2^256
inline Fp fph_mul_lazy(const Fp& a, const Fp& b) {
  uint64_t x[4], y[4], t[9] = {0};
  fp_to_u64(a, x); fp_to_u64(b, y);
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      spg_u128 s = (spg_u128)x[j] * y[i] + t[i + j] + c;
      t[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64);
    }
    t[i + 4] = c;
  }
  // T' = T + p * 2^256 must fit 512 bits
  { uint64_t c = 0;
    for (int i = 0; i < 4; i++) { spg_u128 s = (spg_u128)t[4 + i] + SPG_P64[i] + c; t[4 + i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    if (c) spg_lazy_fail("mul: T + p*2^256 overflow"); }
  const uint64_t p3 = SPG_P64[3];
  uint64_t m[4] = {t[0], t[1], t[2], 0};
  const uint64_t v = p3 * t[0];
  m[3] = t[3] - v;
  const uint64_t c = t[3] < v ? 1 : 0;
  // U = p3 * M (320 bits); result = T'_hi - c - (U >> 64)
  uint64_t u[5];
  u256_mul_small(u, m, p3);
  uint64_t y4[4] = {u[1], u[2], u[3], u[4]};
  { uint64_t cc = c;
    for (int i = 0; i < 4 && cc; i++) { y4[i] += cc; cc = (y4[i] == 0); }
    if (cc) spg_lazy_fail("mul: Y overflow"); }
  if (!u256_geq(t + 4, y4)) spg_lazy_fail("mul: negative result");
  uint64_t r[4];
  u256_sub(r, t + 4, y4);
  return fp_from_u64(r);
}
#else
inline Fp fph_add_raw(const Fp& a, const Fp& b) { return fph_add(a, b); }
inline Fp fph_sub_lazy(const Fp& a, const Fp& b, uint32_t) { return fph_sub(a, b); }
inline Fp fph_partial(const Fp& a) { return fph_reduce(a); }
inline Fp fph_reduce_full(const Fp& a) { return fph_reduce(a); }
inline Fp fph_mul_lazy(const Fp& a, const Fp& b) { return fph_mul(a, b); }
#endif

// ------------------------------------------------------------------ dispatch wrappers
#if defined(__CUDA_ARCH__)
#define SPG_DISPATCH(dev, hst) return dev
#else
#define SPG_DISPATCH(dev, hst) return hst
#endif
SPG_HD Fp fp_mul(const Fp& a, const Fp& b) { SPG_DISPATCH(fpd_mul(a, b), fph_mul(a, b)); }
SPG_HD Fp fp_sqr(const Fp& a) { SPG_DISPATCH(fpd_sqr(a), fph_mul(a, a)); }
SPG_HD Fp fp_add(const Fp& a, const Fp& b) { SPG_DISPATCH(fpd_add(a, b), fph_add(a, b)); }
SPG_HD Fp fp_sub(const Fp& a, const Fp& b) { SPG_DISPATCH(fpd_sub(a, b), fph_sub(a, b)); }
SPG_HD Fp fp_reduce(const Fp& a) { SPG_DISPATCH(fpd_reduce(a), fph_reduce(a)); }
// sum with no reduction (device: caller keeps the lazy bounds; host: canonical)
SPG_HD Fp fp_add_raw(const Fp& a, const Fp& b) { SPG_DISPATCH(fpd_add_raw(a, b), fph_add_raw(a, b)); }
// the lazy family used by the NTT butterflies (bounds: see each fpd_* function)
SPG_HD Fp fp_mul_lazy(const Fp& a, const Fp& b) { SPG_DISPATCH(fpd_mul(a, b), fph_mul_lazy(a, b)); }
SPG_HD Fp fp_sqr_lazy(const Fp& a) { SPG_DISPATCH(fpd_sqr(a), fph_mul_lazy(a, a)); }
SPG_HD Fp fp_sub_lazy(const Fp& a, const Fp& b, uint32_t K) { SPG_DISPATCH(fpd_sub_lazy(a, b, K), fph_sub_lazy(a, b, K)); }
SPG_HD Fp fp_partial(const Fp& a) { SPG_DISPATCH(fpd_partial(a), fph_partial(a)); }
SPG_HD Fp fp_reduce_full(const Fp& a) { SPG_DISPATCH(fpd_reduce_full(a), fph_reduce_full(a)); }

// ------------------------------------------------------------------ shared helpers
SPG_HD Fp fp_to_mont(const Fp& a) { return fp_mul(a, fp_r2()); }
SPG_HD Fp fp_from_mont(const Fp& a) {
  Fp one = fp_zero();
  one.v[0] = 1;
  return fp_reduce(fp_mul(a, one));
}
SPG_HD Fp fp_neg(const Fp& a) { return fp_sub(fp_zero(), a); }
// canonical equality of two lazy values
SPG_HD bool fp_eq(const Fp& a, const Fp& b) { return fp_eq_raw(fp_reduce(a), fp_reduce(b)); }
SPG_HD bool fp_is_zero(const Fp& a) { return fp_is_zero_raw(fp_reduce(a)); }

// a^e for a 256-bit exponent given as 8 x u32 LE (Montgomery in / out)
SPG_HD Fp fp_pow(const Fp& a, const uint32_t* e, int nlimbs) {
  Fp r = fp_one();
  bool started = false;
  for (int i = nlimbs * 32 - 1; i >= 0; i--) {
    if (started) r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? fp_mul(r, a) : a;
      started = true;
    }
  }
  return r;
}
SPG_HD Fp fp_pow_u64(const Fp& a, uint64_t e) {
  uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
  return fp_pow(a, w, 2);
}
// a^-1 = a^(p-2); p-2 = 0x0800000000000010 ffffffffffffffff ffffffffffffffff ffffffffffffffff
SPG_HD Fp fp_inv(const Fp& a) {
  const uint32_t e[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu,
                         0xffffffffu, 0xffffffffu, 0x00000010u, 0x08000000u};
  return fp_pow(a, e, 8);
}

// CPU emulation of the ECDSA verification kernel's per-thread code (csrc/ecdsa.cuh), compiled with g++.
// stdin: lines "msg r s pub_x pub_y|-" (hex canonical, no 0x);  "K priv" lines compute the public key x.
//        "S msg priv seed" lines (seed decimal) sign: the kernel's sign.cuh code (RFC 6979 nonce, k G, mod-n finish).
// stdout: status (0 False / 1 True / 2 raises), the key, or "status r s".
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../stark_perpetual_b200/csrc/sign.cuh"
#include "../../stark_perpetual_b200/csrc/curve_params.inc"

static Fp from_canon(const uint64_t* w) { return fp_to_mont(fp_from_u64(w)); }
static void parse_hex(const char* s, uint32_t w[8]) {
  memset(w, 0, 32);
  size_t n = strlen(s);
  for (size_t i = 0; i < n && i < 64; i++) {
    char c = s[n - 1 - i];
    uint32_t d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
    w[i / 8] |= d << (4 * (i % 8));
  }
}
int main() {
  // tables exactly as ecdsa.cu / ec.cu build them
  APoint shift, gen;
  shift.x = from_canon(SPG_BASE_POINTS[0][0]); shift.y = from_canon(SPG_BASE_POINTS[0][1]);
  gen.x = from_canon(SPG_BASE_POINTS[1][0]); gen.y = from_canon(SPG_BASE_POINTS[1][1]);
  std::vector<APoint> gd(SPG_ECDSA_BITS + 1);
  { APoint g = gen; for (int t = 0; t <= SPG_ECDSA_BITS; t++) { gd[t] = g; g = ec_affine_double(g); } }
  uint64_t three[4] = {3, 0, 0, 0};
  const Fp g3 = from_canon(three);
  const Fp c = fp_pow_u64(g3, (1ull << 59) + 17), ci = fp_inv(c);
  std::vector<Fp> tab(256 + 2 * 24 * 256, fp_one());
  Fp cl = c;
  for (int k = 0; k < 184; k++) cl = fp_sqr(cl);
  for (int k = 1; k < 256; k++) tab[k] = fp_mul(tab[k - 1], cl);
  Fp* D = tab.data() + 256; Fp* Dh = D + 24 * 256;
  Fp base = ci, base_h = ci;
  for (int i = 0; i < 24; i++) {
    for (int k = 1; k < 256; k++) D[i * 256 + k] = fp_mul(D[i * 256 + k - 1], base);
    if (i == 0) { for (int k = 2; k < 256; k += 2) Dh[k] = fp_mul(Dh[k - 2], ci); }
    else { for (int k = 1; k < 256; k++) Dh[i * 256 + k] = fp_mul(Dh[i * 256 + k - 1], base_h); }
    base_h = base;
    for (int k = 0; k < 7; k++) base_h = fp_sqr(base_h);
    for (int k = 0; k < 8; k++) base = fp_sqr(base);
  }
  EcdsaTables T;
  T.gen_doubles = gd.data(); T.shift = shift; T.minus_shift.x = shift.x; T.minus_shift.y = fp_neg(shift.y);
  T.beta = from_canon(SPG_BETA);
  static const uint32_t r2[8] = {0xea1c688du, 0x6021b3f1u, 0x14ce60b9u, 0x509cf64du, 0xf78bbabbu, 0xbaf0ab4cu, 0x2333766eu, 0x07d9e57cu};
  for (int i = 0; i < 8; i++) T.r2_n.v[i] = r2[i];
  T.ninv = SPG_N_INV32;
  T.sq_L = tab.data(); T.sq_D = D; T.sq_Dh = Dh;
  char a[5][128];
  while (scanf("%100s", a[0]) == 1) {
    if (a[0][0] == 'K') {
      if (scanf("%100s", a[1]) != 1) break;
      uint32_t k[8]; parse_hex(a[1], k);
      uint64_t out[4]; fp_to_u64(fp_from_mont(gen_mult(k, T).x), out);
      printf("%016llx%016llx%016llx%016llx\n", (unsigned long long)out[3], (unsigned long long)out[2],
             (unsigned long long)out[1], (unsigned long long)out[0]);
      continue;
    }
    if (a[0][0] == 'E') {                       // E op ax ay b0 b1 -> "status x y"   (math_utils ec_add / ec_double / ec_mult)
      char e[5][128];
      for (int k = 0; k < 5; k++) if (scanf("%100s", e[k]) != 1) return 1;
      uint32_t ax[8], ay[8], b0[8], b1[8];
      parse_hex(e[1], ax); parse_hex(e[2], ay); parse_hex(e[3], b0); parse_hex(e[4], b1);
      std::vector<JPoint> scratch(256);
      APoint R;
      const int st = ec_op_one(atoi(e[0]), ax, ay, b0, b1, scratch.data(), 1, &R);
      uint64_t ox[4], oy[4];
      fp_to_u64(fp_from_mont(R.x), ox); fp_to_u64(fp_from_mont(R.y), oy);
      printf("%d %016llx%016llx%016llx%016llx %016llx%016llx%016llx%016llx\n", st, (unsigned long long)ox[3], (unsigned long long)ox[2],
             (unsigned long long)ox[1], (unsigned long long)ox[0], (unsigned long long)oy[3], (unsigned long long)oy[2],
             (unsigned long long)oy[1], (unsigned long long)oy[0]);
      continue;
    }
    if (a[0][0] == 'Q') {                       // Q a -> "status y"   (math_utils sqrt_mod / is_quad_residue)
      if (scanf("%100s", a[1]) != 1) break;
      uint32_t av[8]; parse_hex(a[1], av);
      Fp am, y = fp_zero();
      int st = 0;
      if (!canon_to_mont(av, &am)) st = 2;
      else if (!fp_sqrt_min(am, T, &y)) { st = 1; y = fp_zero(); }
      uint64_t oy[4]; fp_to_u64(fp_from_mont(y), oy);
      printf("%d %016llx%016llx%016llx%016llx\n", st, (unsigned long long)oy[3], (unsigned long long)oy[2], (unsigned long long)oy[1],
             (unsigned long long)oy[0]);
      continue;
    }
    if (a[0][0] == 'S') {
      for (int k = 1; k < 4; k++) if (scanf("%100s", a[k]) != 1) return 1;
      uint32_t m[8], d[8], r[8] = {0}, s[8] = {0};
      parse_hex(a[1], m); parse_hex(a[2], d);
      const int st = ecdsa_sign_one(m, d, strtoull(a[3], nullptr, 10), T, r, s);
      printf("%d ", st);
      for (int k = 7; k >= 0; k--) printf("%08x", r[k]);
      printf(" ");
      for (int k = 7; k >= 0; k--) printf("%08x", s[k]);
      printf("\n");
      continue;
    }
    for (int k = 1; k < 5; k++) if (scanf("%100s", a[k]) != 1) return 1;
    uint32_t m[8], r[8], s[8], x[8], y[8];
    parse_hex(a[0], m); parse_hex(a[1], r); parse_hex(a[2], s); parse_hex(a[3], x);
    const bool has_y = a[4][0] != '-';
    if (has_y) parse_hex(a[4], y);
    printf("%d\n", ecdsa_verify_one(m, r, s, x, has_y ? y : nullptr, T));
  }
  return 0;
}

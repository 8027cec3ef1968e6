// CPU run of the ECDSA-builtin AIR's per-point code (csrc/ecdsa_air_point.cuh), compiled with g++.
// stdin, per case (hex canonical values): gx gy shift_x shift_y beta, the 2 public column values, 52 alpha powers, 25 cells
// at x, 25 cells at x w, 6 inverse zerofier values.  stdout: the composition value (canonical hex).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../stark_perpetual_b200/csrc/ecdsa_air_point.cuh"

static bool read_felt(Fp* out) {
  char s[128];
  if (scanf("%100s", s) != 1) return false;
  uint64_t w[4] = {0, 0, 0, 0};
  size_t n = strlen(s);
  for (size_t i = 0; i < n && i < 64; i++) {
    char c = s[n - 1 - i];
    uint64_t d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
    w[i / 16] |= d << (4 * (i % 16));
  }
  *out = fp_to_mont(fp_from_u64(w));
  return true;
}
int main() {
  Fp gx;
  while (read_felt(&gx)) {
    Fp gy, cur[SPG_EAIR_COLS], nxt[SPG_EAIR_COLS], iz[SPG_EAIR_NGROUPS];
    EcdsaAirConsts K;
    if (!read_felt(&gy) || !read_felt(&K.shift_x) || !read_felt(&K.shift_y) || !read_felt(&K.beta)) return 1;
    K.minus_shift_y = fp_neg(K.shift_y);
    Fp fm, fk;
    if (!read_felt(&fm) || !read_felt(&fk)) return 1;
    for (auto& a : K.alpha) if (!read_felt(&a)) return 1;
    for (auto& v : cur) if (!read_felt(&v)) return 1;
    for (auto& v : nxt) if (!read_felt(&v)) return 1;
    for (auto& z : iz) if (!read_felt(&z)) return 1;
    const Fp v = fp_from_mont(ecdsa_air_point(cur, nxt, gx, gy, fm, fk, K, iz));
    uint64_t o[4];
    fp_to_u64(v, o);
    printf("%016llx%016llx%016llx%016llx\n", (unsigned long long)o[3], (unsigned long long)o[2], (unsigned long long)o[1],
           (unsigned long long)o[0]);
  }
  return 0;
}

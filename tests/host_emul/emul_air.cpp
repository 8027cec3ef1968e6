// CPU run of the AIR kernel's per-point code (csrc/air_point.cuh), compiled with g++ -DSPG_EMUL_LAZY: the host
// arithmetic reproduces the device's lazy representatives and aborts on any violated bound.
// stdin, per case (hex canonical values): px py shift_x shift_y, 65 alpha powers, 5 x (X Y S M I Xn Yn Mn x0 out),
// 8 inverse zerofier values.  stdout: the composition value (canonical hex).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../stark_perpetual_b200/csrc/air_point.cuh"

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
  Fp px;
  while (read_felt(&px)) {
    Fp py, sx, sy, alpha[SPG_AIR_LANES * SPG_AIR_NCONSTR], iz[8];
    if (!read_felt(&py) || !read_felt(&sx) || !read_felt(&sy)) return 1;
    for (auto& a : alpha) if (!read_felt(&a)) return 1;
    AirAcc A;
    air_acc_init(A);
    for (int l = 0; l < SPG_AIR_LANES; l++) {
      AirRow r;
      Fp x0, out;
      Fp* cells[10] = {&r.X, &r.Y, &r.S, &r.M, &r.I, &r.Xn, &r.Yn, &r.Mn, &x0, &out};
      for (Fp* c : cells) if (!read_felt(c)) return 1;
      air_lane_accumulate(A, r, alpha + SPG_AIR_NCONSTR * l, px, py, sx, sy, x0, out);
    }
    for (auto& z : iz) if (!read_felt(&z)) return 1;
    const Fp v = fp_from_mont(air_combine(A, iz[0], iz[1], iz[2], iz[3], iz[4], iz[5], iz[6], iz[7]));
    uint64_t o[4];
    fp_to_u64(v, o);
    printf("%016llx%016llx%016llx%016llx\n", (unsigned long long)o[3], (unsigned long long)o[2], (unsigned long long)o[1],
           (unsigned long long)o[0]);
  }
  return 0;
}

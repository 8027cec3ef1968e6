// CPU run of the ECDSA-builtin AIR's witness generator (csrc/ecdsa_air_witness.cuh): the three kernels' per-thread code,
// thread by thread, compiled with g++.  argv[1] = log_n.  stdin: N/256 lines "msg r w key_x key_y" (hex canonical).
// stdout: the status word, then the trace, one canonical hex value per line, column-major [25][N].
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../stark_perpetual_b200/csrc/ecdsa_air_witness.cuh"
#include "../../stark_perpetual_b200/csrc/curve_params.inc"

static Fp parse_canon(const char* s) {          // canonical representation (the kernels' input form), not Montgomery
  uint64_t w[4] = {0, 0, 0, 0};
  size_t n = strlen(s);
  for (size_t i = 0; i < n && i < 64; i++) {
    char c = s[n - 1 - i];
    uint64_t d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
    w[i / 16] |= d << (4 * (i % 16));
  }
  return fp_from_u64(w);
}
int main(int argc, char** argv) {
  const unsigned log_n = argc > 1 ? (unsigned)atoi(argv[1]) : 9;
  const size_t n = (size_t)1 << log_n, nb = n >> 8;
  APoint shift, gen;
  shift.x = fp_to_mont(fp_from_u64(SPG_BASE_POINTS[0][0])); shift.y = fp_to_mont(fp_from_u64(SPG_BASE_POINTS[0][1]));
  gen.x = fp_to_mont(fp_from_u64(SPG_BASE_POINTS[1][0])); gen.y = fp_to_mont(fp_from_u64(SPG_BASE_POINTS[1][1]));
  std::vector<APoint> gd(SPG_EAIR_BITS + 1);
  { APoint g = gen; for (int t = 0; t <= SPG_EAIR_BITS; t++) { gd[t] = g; g = ec_affine_double(g); } }
  const Fp beta = fp_to_mont(fp_from_u64(SPG_BETA));
  std::vector<Fp> in[5];
  for (size_t b = 0; b < nb; b++) {
    char s[5][128];
    if (scanf("%100s %100s %100s %100s %100s", s[0], s[1], s[2], s[3], s[4]) != 5) return 1;
    for (int k = 0; k < 5; k++) in[k].push_back(parse_canon(s[k]));
  }
  std::vector<Fp> trace((size_t)SPG_EAIR_COLS * n, fp_zero()), cross(4 * nb, fp_zero());
  uint32_t status = 0;
  for (size_t i = 0; i < 2 * nb; i++)
    eair_walk_ab_thread(i, log_n, in[0].data(), in[1].data(), in[3].data(), in[4].data(), trace.data(), &status, gd.data(), shift, beta);
  for (size_t b = 0; b < nb; b++)
    eair_walk_c_thread(b, log_n, in[0].data(), in[1].data(), in[2].data(), trace.data(), cross.data(), &status, shift);
  for (size_t i = 0; i < 3 * (n / SPG_EAIR_WIT_ROWS); i++)
    eair_finish_thread(i, log_n, trace.data(), cross.data(), &status, gd.data());
  printf("%u\n", status);
  for (const Fp& v : trace) {
    uint64_t o[4];
    fp_to_u64(v, o);
    printf("%016llx%016llx%016llx%016llx\n", (unsigned long long)o[3], (unsigned long long)o[2], (unsigned long long)o[1],
           (unsigned long long)o[0]);
  }
  return 0;
}

// Integer-pipe microbenchmarks, second set: every operation consumes its own previous result, so ptxas cannot
// hoist or strength-reduce anything.  Each pattern's SASS was checked with cuobjdump before trusting the number.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/microbench2 tools/microbench2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
#define NCH 8

// plain IMAD.WIDE with a 64-bit addend, no carry flags: a = lo(a) * q + a
__global__ void __launch_bounds__(256) k_wide_acc(unsigned long long* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[NCH];
  for (int k = 0; k < NCH; k++) a[k] = i * 77u + k;
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"((unsigned)a[k]), "r"(q));
  }
  unsigned long long s = 0;
  for (int k = 0; k < NCH; k++) s ^= a[k];
  if (s == 0x1234567ull) out[i] = s;
}
// IMAD.WIDE without addend: a = lo(a) * hi(a)
__global__ void __launch_bounds__(256) k_wide_mul(unsigned long long* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[NCH];
  for (int k = 0; k < NCH; k++) a[k] = (i * 77u + k) | ((unsigned long long)(q + k) << 32);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++)
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[k]) : "r"((unsigned)a[k]), "r"((unsigned)(a[k] >> 32)));
  }
  unsigned long long s = 0;
  for (int k = 0; k < NCH; k++) s ^= a[k];
  if (s == 0x1234567ull) out[i] = s;
}
// 32-bit IMAD: a = a * q + a
__global__ void __launch_bounds__(256) k_imad32(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[NCH];
  for (int k = 0; k < NCH; k++) a[k] = i * 77u + k;
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[k]) : "r"(q));
  }
  unsigned s = 0;
  for (int k = 0; k < NCH; k++) s ^= a[k];
  if (s == 0x1234567u) out[i] = s;
}
// IMAD.HI: a = hi(a * q) + a
__global__ void __launch_bounds__(256) k_imadhi(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[NCH];
  for (int k = 0; k < NCH; k++) a[k] = i * 77u + k + 0x80000000u;
  q ^= i; q |= 0xf0000000u;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[k]) : "r"(q));
  }
  unsigned s = 0;
  for (int k = 0; k < NCH; k++) s ^= a[k];
  if (s == 0x1234567u) out[i] = s;
}
// plain IADD3 (no carries): a = a + b + rotating partner
__global__ void __launch_bounds__(256) k_iadd3(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[NCH];
  for (int k = 0; k < NCH; k++) a[k] = i * 77u + k;
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a[k]) : "r"(q), "r"(a[(k + 1) % NCH]));
  }
  unsigned s = 0;
  for (int k = 0; k < NCH; k++) s ^= a[k];
  if (s == 0x1234567u) out[i] = s;
}
// LOP3 / SHF: a = (a ^ q) funnel-shifted with neighbour
__global__ void __launch_bounds__(256) k_shf(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[NCH];
  for (int k = 0; k < NCH; k++) a[k] = i * 77u + k;
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a[k]) : "r"(a[(k + 1) % NCH]));
  }
  unsigned s = 0;
  for (int k = 0; k < NCH; k++) s ^= a[k];
  if (s == 0x1234567u) out[i] = s;
}
// 8-limb carry chain where the second operand is the previous result of ANOTHER chain (4 chains)
__global__ void __launch_bounds__(256) k_addc_chain(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[4][8];
  for (int r = 0; r < 4; r++) for (int k = 0; k < 8; k++) a[r][k] = i * 77u + k + r * q;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int o = (r + 1) & 3;
      asm volatile(
          "add.cc.u32 %0, %0, %8;\n\t"
          "addc.cc.u32 %1, %1, %9;\n\t"
          "addc.cc.u32 %2, %2, %10;\n\t"
          "addc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\t"
          "addc.cc.u32 %5, %5, %13;\n\t"
          "addc.cc.u32 %6, %6, %14;\n\t"
          "addc.u32 %7, %7, %15;"
          : "+r"(a[r][0]), "+r"(a[r][1]), "+r"(a[r][2]), "+r"(a[r][3]), "+r"(a[r][4]), "+r"(a[r][5]), "+r"(a[r][6]),
            "+r"(a[r][7])
          : "r"(a[o][0]), "r"(a[o][1]), "r"(a[o][2]), "r"(a[o][3]), "r"(a[o][4]), "r"(a[o][5]), "r"(a[o][6]), "r"(a[o][7]));
    }
  }
  unsigned s = 0;
  for (int r = 0; r < 4; r++) for (int k = 0; k < 8; k++) s ^= a[r][k];
  if (s == 0x1234567u) out[i] = s;
}
// carry rows (IMAD.WIDE.X): multiplier limbs are the row's own previous outputs
__global__ void __launch_bounds__(256) k_madc_rows(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[2][9];
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) a[r][k] = i * 77u + k + r;
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int o = r ^ 1;
      asm volatile(
          "mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
          "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
          "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
          "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
          "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
          "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
          "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
          "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
          "addc.u32 %8, %8, 0;"
          : "+r"(a[r][0]), "+r"(a[r][1]), "+r"(a[r][2]), "+r"(a[r][3]), "+r"(a[r][4]), "+r"(a[r][5]), "+r"(a[r][6]),
            "+r"(a[r][7]), "+r"(a[r][8])
          : "r"(a[o][0]), "r"(a[o][2]), "r"(a[o][4]), "r"(a[o][6]), "r"(q));
    }
  }
  unsigned s = 0;
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) s ^= a[r][k];
  if (s == 0x1234567u) out[i] = s;
}
// mixed: NCH plain IMAD.WIDE accumulations + NCH independent IADD3 pairs per iteration
__global__ void __launch_bounds__(256) k_mix(unsigned long long* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[NCH];
  unsigned c[NCH];
  for (int k = 0; k < NCH; k++) { a[k] = i * 77u + k; c[k] = i * 3u + k; }
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"((unsigned)a[k]), "r"(q));
      asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(c[k]) : "r"(q), "r"(c[(k + 1) % NCH]));
    }
  }
  unsigned long long s = 0;
  for (int k = 0; k < NCH; k++) s ^= a[k] ^ c[k];
  if (s == 0x1234567ull) out[i] = s;
}
// mixed: plain IMAD.WIDE + 2 carry-chain adds per product (the "multiply on FMA, accumulate on ALU" shape)
__global__ void __launch_bounds__(256) k_mul_addcc(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[2][9];
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) a[r][k] = i * 77u + k + r;
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int o = r ^ 1;
      unsigned long long p0, p1, p2, p3;
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a[o][0]), "r"(q));
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(a[o][2]), "r"(q));
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p2) : "r"(a[o][4]), "r"(q));
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p3) : "r"(a[o][6]), "r"(q));
      asm volatile(
          "add.cc.u32 %0, %0, %9;\n\t"
          "addc.cc.u32 %1, %1, %10;\n\t"
          "addc.cc.u32 %2, %2, %11;\n\t"
          "addc.cc.u32 %3, %3, %12;\n\t"
          "addc.cc.u32 %4, %4, %13;\n\t"
          "addc.cc.u32 %5, %5, %14;\n\t"
          "addc.cc.u32 %6, %6, %15;\n\t"
          "addc.cc.u32 %7, %7, %16;\n\t"
          "addc.u32 %8, %8, 0;"
          : "+r"(a[r][0]), "+r"(a[r][1]), "+r"(a[r][2]), "+r"(a[r][3]), "+r"(a[r][4]), "+r"(a[r][5]), "+r"(a[r][6]),
            "+r"(a[r][7]), "+r"(a[r][8])
          : "r"((unsigned)p0), "r"((unsigned)(p0 >> 32)), "r"((unsigned)p1), "r"((unsigned)(p1 >> 32)), "r"((unsigned)p2),
            "r"((unsigned)(p2 >> 32)), "r"((unsigned)p3), "r"((unsigned)(p3 >> 32)));
    }
  }
  unsigned s = 0;
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) s ^= a[r][k];
  if (s == 0x1234567u) out[i] = s;
}

template <class F>
static float time_it(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256;
  const double clk = khz * 1e3;
  void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
  const double thr = (double)blocks * threads * ITERS;
  auto rate = [&](float ms, double per_iter) { return thr * per_iter / (ms * 1e-3) / clk / sms; };
  printf("{\"sm_clock_khz\": %d, \"sms\": %d, \"unit\": \"thread-instructions per clock per SM\",\n", khz, sms);
  float ms;
  ms = time_it([&] { k_wide_acc<<<blocks, threads>>>((unsigned long long*)buf, 777u); });
  printf(" \"imad_wide_plain_acc\": %.2f,\n", rate(ms, NCH));
  ms = time_it([&] { k_wide_mul<<<blocks, threads>>>((unsigned long long*)buf, 777u); });
  printf(" \"imad_wide_mul_only\": %.2f,\n", rate(ms, NCH));
  ms = time_it([&] { k_imad32<<<blocks, threads>>>((unsigned*)buf, 777u); });
  printf(" \"imad32\": %.2f,\n", rate(ms, NCH));
  ms = time_it([&] { k_imadhi<<<blocks, threads>>>((unsigned*)buf, 777u); });
  printf(" \"imad_hi\": %.2f,\n", rate(ms, NCH));
  ms = time_it([&] { k_iadd3<<<blocks, threads>>>((unsigned*)buf, 777u); });
  printf(" \"iadd3_plain\": %.2f,\n", rate(ms, NCH));
  ms = time_it([&] { k_shf<<<blocks, threads>>>((unsigned*)buf, 777u); });
  printf(" \"shf\": %.2f,\n", rate(ms, NCH));
  ms = time_it([&] { k_addc_chain<<<blocks, threads>>>((unsigned*)buf, 777u); });
  printf(" \"iadd3_x_chain\": %.2f,\n", rate(ms, 32));
  ms = time_it([&] { k_madc_rows<<<blocks, threads>>>((unsigned*)buf, 777u); });
  printf(" \"imad_wide_x_rows\": %.2f,\n", rate(ms, 18));
  ms = time_it([&] { k_mix<<<blocks, threads>>>((unsigned long long*)buf, 777u); });
  printf(" \"mix_wide_plus_iadd3_total\": %.2f,\n", rate(ms, 2 * NCH));
  ms = time_it([&] { k_mul_addcc<<<blocks, threads>>>((unsigned*)buf, 777u); });
  printf(" \"mulwide_then_addcc_total\": %.2f, \"mulwide_then_addcc_note\": \"4 IMAD.WIDE + 9 IADD3.X per row\"}\n", rate(ms, 26));
  return 0;
}

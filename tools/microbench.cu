// Integer-pipe microbenchmarks for sm_100a: issue rates of the instruction patterns the 252-bit field
// arithmetic is built from.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/microbench tools/microbench.cu
// Prints instructions per clock per SM for each pattern (SM clock read from the device attribute; run at
// the default application clocks).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096

// A: plain mad.wide.u32, 8 independent accumulators
__global__ void __launch_bounds__(256) k_madwide(unsigned long long* out, unsigned m, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[8];
  for (int k = 0; k < 8; k++) a[k] = i + k;
  m += i; q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(m), "r"(q));
  }
  unsigned long long s = 0;
  for (int k = 0; k < 8; k++) s ^= a[k];
  if (s == 0x1234567ull) out[i] = s;
}

// B: carry-chained rows: mad.lo.cc / madc.hi.cc x4 + addc  (the SPG_ROW_MAD pattern), two independent rows
__global__ void __launch_bounds__(256) k_madc_rows(unsigned* out, unsigned m, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[2][9];
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) a[r][k] = i + k + r;
  unsigned x0 = m + i, x1 = m ^ i, x2 = m * 3 + i, x3 = m - i;
  q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 2; r++)
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
          : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(q));
  }
  unsigned s = 0;
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) s ^= a[r][k];
  if (s == 0x1234567u) out[i] = s;
}

// C: 8-limb add.cc chains, 4 independent chains
__global__ void __launch_bounds__(256) k_addc(unsigned* out, unsigned m) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[4][8], b[8];
  for (int r = 0; r < 4; r++) for (int k = 0; k < 8; k++) a[r][k] = i + k + r;
  for (int k = 0; k < 8; k++) b[k] = m * (k + 1) + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
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
          : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  }
  unsigned s = 0;
  for (int r = 0; r < 4; r++) for (int k = 0; k < 8; k++) s ^= a[r][k];
  if (s == 0x1234567u) out[i] = s;
}

// D: 32-bit mad.lo, 8 independent
__global__ void __launch_bounds__(256) k_madlo(unsigned* out, unsigned m, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[8];
  for (int k = 0; k < 8; k++) a[k] = i + k;
  m += i; q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(m), "r"(q));
  }
  unsigned s = 0;
  for (int k = 0; k < 8; k++) s ^= a[k];
  if (s == 0x1234567u) out[i] = s;
}

// E: mix: per iteration 8 plain mad.wide + 8 add.cc-chain limbs (independent): do the two pipes overlap?
__global__ void __launch_bounds__(256) k_mix(unsigned long long* out, unsigned m, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[8];
  unsigned c[8], b[8];
  for (int k = 0; k < 8; k++) { a[k] = i + k; c[k] = i * k; b[k] = m * (k + 1) + i; }
  m += i; q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(m), "r"(q));
    asm volatile(
        "add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  }
  unsigned long long s = 0;
  for (int k = 0; k < 8; k++) s ^= a[k] ^ c[k];
  if (s == 0x1234567ull) out[i] = s;
}

// F: mad.wide with 64-bit accumulate done as mul.wide + add.cc/addc (3 instr per product)
__global__ void __launch_bounds__(256) k_mulwide_add(unsigned long long* out, unsigned m, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[8];
  for (int k = 0; k < 8; k++) a[k] = i + k;
  m += i; q ^= i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      unsigned long long p;
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(m + k), "r"(q));
      a[k] += p;
      m ^= (unsigned)a[k];
    }
  }
  unsigned long long s = 0;
  for (int k = 0; k < 8; k++) s ^= a[k];
  if (s == 0x1234567ull) out[i] = s;
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
  printf("{\"sm_clock_khz\": %d, \"sms\": %d,\n", khz, sms);
  float ms;
  ms = time_it([&] { k_madwide<<<blocks, threads>>>((unsigned long long*)buf, 12345u, 777u); });
  printf(" \"mad_wide_plain_per_clk_sm\": %.2f,\n", rate(ms, 8));
  ms = time_it([&] { k_madc_rows<<<blocks, threads>>>((unsigned*)buf, 12345u, 777u); });
  printf(" \"imad_wide_x_rows_per_clk_sm\": %.2f, \"rows_note\": \"8 IMAD.WIDE.X + 2 addc per iteration, counted as 8\",\n", rate(ms, 8));
  ms = time_it([&] { k_addc<<<blocks, threads>>>((unsigned*)buf, 12345u); });
  printf(" \"iadd3_x_chain_per_clk_sm\": %.2f,\n", rate(ms, 32));
  ms = time_it([&] { k_madlo<<<blocks, threads>>>((unsigned*)buf, 12345u, 777u); });
  printf(" \"mad_lo_per_clk_sm\": %.2f,\n", rate(ms, 8));
  ms = time_it([&] { k_mix<<<blocks, threads>>>((unsigned long long*)buf, 12345u, 777u); });
  printf(" \"mix_8wide_8add_instr_per_clk_sm\": %.2f,\n", rate(ms, 16));
  ms = time_it([&] { k_mulwide_add<<<blocks, threads>>>((unsigned long long*)buf, 12345u, 777u); });
  printf(" \"mulwide_plus_add64_products_per_clk_sm\": %.2f}\n", rate(ms, 8));
  return 0;
}

// Third microbenchmark set: what makes IMAD.WIDE run at 2 vs 4 cycles per warp instruction on sm_100a?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/microbench3 tools/microbench3.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048

#define ROW(acc, x0, x1, x2, x3, b)                                                                  \
  asm volatile(                                                                                      \
      "mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                           \
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                          \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                         \
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                         \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                         \
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                         \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                         \
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                         \
      "addc.u32 %8, %8, 0;"                                                                          \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), \
        "+r"(acc[7]), "+r"(acc[8])                                                                   \
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b))

// MODE 0: two rows, shared multiplier q  (each row is 4 IMAD.WIDE: ptxas fuses every mad.lo.cc / madc.hi.cc pair)
// MODE 1: multiplier = a limb of the other row (changes every iteration)
// MODE 2: four rows in flight (more ILP), multiplier from the other row
// MODE 3: as 1 but 128 registers forced via a big dummy array? (occupancy) -- done through launch bounds instead
template <int MODE>
__global__ void __launch_bounds__(256) k_rows(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[4][9];
  for (int r = 0; r < 4; r++) for (int k = 0; k < 9; k++) a[r][k] = (i * 2654435761u) ^ (k * 0x9e3779b9u + r * 0x85ebca6bu);
  q ^= i * 0xc2b2ae35u;
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {
      ROW(a[0], a[1][0], a[1][2], a[1][4], a[1][6], q);
      ROW(a[1], a[0][0], a[0][2], a[0][4], a[0][6], q);
    } else if (MODE == 1) {
      ROW(a[0], a[1][0], a[1][2], a[1][4], a[1][6], a[1][7]);
      ROW(a[1], a[0][0], a[0][2], a[0][4], a[0][6], a[0][7]);
    } else {
      ROW(a[0], a[1][0], a[1][2], a[1][4], a[1][6], a[1][7]);
      ROW(a[1], a[2][0], a[2][2], a[2][4], a[2][6], a[2][7]);
      ROW(a[2], a[3][0], a[3][2], a[3][4], a[3][6], a[3][7]);
      ROW(a[3], a[0][0], a[0][2], a[0][4], a[0][6], a[0][7]);
    }
  }
  unsigned s = 0;
  for (int r = 0; r < 4; r++) for (int k = 0; k < 9; k++) s ^= a[r][k];
  if (s == 0x1234567u) out[i] = s;
}

// plain IMAD.WIDE (mul.wide) whose operands come from the neighbouring accumulator: no self-dependency
__global__ void __launch_bounds__(256) k_wide_ring(unsigned long long* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[8];
  for (int k = 0; k < 8; k++) a[k] = ((unsigned long long)(i * 2654435761u + k) << 32) | (i * 0x9e3779b9u + k * q);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int o = (k + 3) & 7;
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[k]) : "r"((unsigned)a[o]), "r"((unsigned)(a[o] >> 32)));
    }
  }
  unsigned long long s = 0;
  for (int k = 0; k < 8; k++) s ^= a[k];
  if (s == 0x1234567ull) out[i] = s;
}
// mad.wide with 64-bit addend, operands from the neighbour, OR-ed with 1 to keep values alive
__global__ void __launch_bounds__(256) k_wide_acc_ring(unsigned long long* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a[8];
  for (int k = 0; k < 8; k++) a[k] = ((unsigned long long)(i * 2654435761u + k) << 32) | (i * 0x9e3779b9u + k * q);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int o = (k + 3) & 7;
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"((unsigned)a[o]), "r"((unsigned)(a[o] >> 32)));
    }
  }
  unsigned long long s = 0;
  for (int k = 0; k < 8; k++) s ^= a[k];
  if (s == 0x1234567ull) out[i] = s;
}
// 32-bit IMAD ring
__global__ void __launch_bounds__(256) k_imad_ring(unsigned* out, unsigned q) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned a[8];
  for (int k = 0; k < 8; k++) a[k] = i * 2654435761u + k * q;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int o = (k + 3) & 7, o2 = (k + 5) & 7;
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(a[o]), "r"(a[o2]));
    }
  }
  unsigned s = 0;
  for (int k = 0; k < 8; k++) s ^= a[k];
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
  const int sms = prop.multiProcessorCount, threads = 256;
  const double clk = khz * 1e3;
  void* buf; cudaMalloc(&buf, (size_t)sms * 8 * threads * 8);
  printf("{\"unit\": \"SMSP cycles per warp instruction; a carry row = 4 mad.lo.cc/madc.hi.cc pairs = 4 IMAD.WIDE[.X] + 1 IADD3.X, counted as 4\",\n");
  for (int bps = 8; bps >= 2; bps /= 2) {      // blocks per SM: 16, 8, 4 warps per SMSP
    const int blocks = sms * bps;
    const double thr = (double)blocks * threads * ITERS;
    auto cyc = [&](float ms, double per_iter) { return (ms * 1e-3) * clk * sms * 4 * 32 / (thr * per_iter); };
    float ms;
    printf(" \"warps_per_smsp_%d\": {", bps * 2);
    ms = time_it([&] { k_rows<0><<<blocks, threads>>>((unsigned*)buf, 777u); });
    printf("\"rows_shared_q\": %.2f, ", cyc(ms, 8));
    ms = time_it([&] { k_rows<1><<<blocks, threads>>>((unsigned*)buf, 777u); });
    printf("\"rows_var_mult\": %.2f, ", cyc(ms, 8));
    ms = time_it([&] { k_rows<2><<<blocks, threads>>>((unsigned*)buf, 777u); });
    printf("\"rows4_var_mult\": %.2f, ", cyc(ms, 16));
    ms = time_it([&] { k_wide_ring<<<blocks, threads>>>((unsigned long long*)buf, 777u); });
    printf("\"mul_wide_ring\": %.2f, ", cyc(ms, 8));
    ms = time_it([&] { k_wide_acc_ring<<<blocks, threads>>>((unsigned long long*)buf, 777u); });
    printf("\"mad_wide_ring\": %.2f, ", cyc(ms, 8));
    ms = time_it([&] { k_imad_ring<<<blocks, threads>>>((unsigned*)buf, 777u); });
    printf("\"imad32_ring\": %.2f},\n", cyc(ms, 8));
  }
  printf(" \"sm_clock_khz\": %d}\n", khz);
  return 0;
}

// A/B harness for 252-bit Montgomery multiplication variants on sm_100a: every variant runs the same
// dependent-chain loop with opaque operands and is checked against the host (unsigned __int128) product.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/mulbench tools/mulbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../stark_perpetual_b200/csrc/fp.cuh"

template <int V>
SPG_D Fp mul_variant(const Fp& a, const Fp& b) {
  uint32_t t[16];
  fpd_mul_wide(t, a, b);
  if (V == 0) return fpd_redc_imad(t);
  if (V == 1) return fpd_redc_shift(t);
  if (V == 2) return fpd_redc_hybrid(t);
  if (V == 4) return fpd_sqr(a);                 // dedicated squaring (36 wide multiplies), shift reduction
  if (V == 7 || V == 8) {                        // dedicated squaring with the IMAD / hybrid reduction
    uint32_t t3[16];
    fpd_sqr_wide(t3, a);
    return V == 7 ? fpd_redc_imad(t3) : fpd_redc_hybrid(t3);
  }
  if (V == 5) return fpd_mul(a, a);              // squaring through the general product
  if (V == 6) {   // squaring product only
    uint32_t t2[16];
    fpd_sqr_wide(t2, a);
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t2[i] ^ t2[8 + i];
    r.v[7] &= 0x03ffffffu;
    return r;
  }
  if (V == 3) {   // product only: fold the halves so nothing is dead
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i] ^ t[8 + i];
    r.v[7] &= 0x07ffffffu;
    return r;
  }
  return a;
}

template <int V>
__global__ void __launch_bounds__(256) k_bench(Fp* io, int iters) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = io[2 * i], y = io[2 * i + 1];
  for (int k = 0; k < iters; k++) x = mul_variant<V>(x, y);
  io[2 * i] = x;
}

template <int V>
static void run(const char* name, Fp* d_io, const std::vector<Fp>& h_in, int blocks, int threads, int iters, bool check) {
  const size_t n = (size_t)blocks * threads;
  cudaMemcpy(d_io, h_in.data(), 2 * n * sizeof(Fp), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_bench<V><<<blocks, threads>>>(d_io, 8);      // warm-up + correctness data
  cudaDeviceSynchronize();
  size_t bad = 0;
  if (check) {
    std::vector<Fp> out(2 * n);
    cudaMemcpy(out.data(), d_io, 2 * n * sizeof(Fp), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < n; i += 97) {
      Fp x = h_in[2 * i], y = h_in[2 * i + 1];
      for (int k = 0; k < 8; k++) x = (V == 4 || V == 5 || V == 7 || V == 8) ? fph_mul(x, x) : fph_mul(x, y);
      if (!fp_eq_raw(fph_reduce(out[2 * i]), x)) bad++;
    }
  }
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    k_bench<V><<<blocks, threads>>>(d_io, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double mps = (double)n * iters / (best * 1e-3);
  double cyc = 148.0 * 4 * 32 * khz * 1e3 / mps;
  printf(" \"%s\": {\"mul_per_s\": %.4g, \"cycles_per_warp_mul_per_smsp\": %.1f, \"mismatches\": %zu},\n", name, mps, cyc, bad);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2000;
  const size_t n = (size_t)blocks * threads;
  std::vector<Fp> h(2 * n);
  uint64_t s = 88172645463325252ull;
  for (auto& e : h) {
    uint64_t w[4];
    for (int k = 0; k < 4; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w[k] = s; }
    w[3] &= 0x07ffffffffffffffull;
    e = fp_from_u64(w);
  }
  Fp* d; cudaMalloc(&d, 2 * n * sizeof(Fp));
  printf("{\n");
  run<0>("imad_redc", d, h, blocks, threads, iters, true);
  run<1>("shift_redc", d, h, blocks, threads, iters, true);
  run<2>("hybrid_redc", d, h, blocks, threads, iters, true);
  run<3>("product_only", d, h, blocks, threads, iters, false);
  run<4>("sqr_dedicated", d, h, blocks, threads, iters, true);
  run<5>("sqr_as_mul", d, h, blocks, threads, iters, true);
  run<6>("sqr_product_only", d, h, blocks, threads, iters, false);
  run<7>("sqr_imad_redc", d, h, blocks, threads, iters, true);
  run<8>("sqr_hybrid_redc", d, h, blocks, threads, iters, true);
  printf(" \"sms\": %d}\n", prop.multiProcessorCount);
  return 0;
}

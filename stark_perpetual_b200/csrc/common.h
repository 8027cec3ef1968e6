// Internal context shared by the translation units of libspg.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/spg.h"
#include "fp.cuh"

struct spg_ctx {
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_ms = 0.0;
  uint64_t launches = 0;
  std::string err;
  // NTT tables (device, Montgomery form)
  Fp* tw_fwd = nullptr;   // omega_1024^e, 512
  Fp* tw_inv = nullptr;   // omega_1024^-e, 512
  Fp* uniA = nullptr;     // omega_{2^26}^(i << 13), 8192
  Fp* uniB = nullptr;     // omega_{2^26}^i, 8192
  // curve tables
  Fp* const_points = nullptr;   // 506 x (x, y) Montgomery
  // scratch cache
  std::vector<void*> owned;
};

#define SPG_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + \
                 ":" + std::to_string(__LINE__);                                        \
      return SPG_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

#define SPG_ARG(cond, msg)             \
  do {                                 \
    if (!(cond)) {                     \
      ctx->err = std::string("bad argument: ") + (msg); \
      return SPG_E_ARG;                \
    }                                  \
  } while (0)

#define SPG_LAUNCH_CHECK()                          \
  do {                                              \
    ctx->launches++;                                \
    SPG_CUDA(cudaGetLastError());                   \
  } while (0)

// RAII device buffer (freed on scope exit)
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  template <class T> T* as() { return (T*)p; }
};

// host helpers (ctx.cu)
Fp spg_host_root_of_unity(int log_n);   // omega_{2^log_n}, Montgomery
Fp spg_host_from_u64(const uint64_t* canon);   // canonical -> Montgomery
void spg_host_to_u64(const Fp& mont, uint64_t* canon);

// ntt.cu
int spg_ntt_device(spg_ctx* ctx, const Fp* in, Fp* out, unsigned log_n, size_t ncols, size_t in_stride,
                   size_t out_stride, int inverse, int dit, unsigned long long coset_exp,
                   const Fp* scale_lo, const Fp* scale_hi);
int spg_bitrev_device(spg_ctx* ctx, const Fp* in, Fp* out, unsigned log_n, size_t ncols);

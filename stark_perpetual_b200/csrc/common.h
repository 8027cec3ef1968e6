// Internal context shared by the translation units of libspg.
#pragma once
#include <string.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/spg.h"
#include "fp.cuh"

struct spg_ctx {
  // Every C-ABI entry point holds this lock for its whole duration (SPG_LOCK): a context owns one stream, one temporary
  // pool, one error string and one set of timing events, so concurrent calls on the SAME context are serialised here
  // (ctypes releases the GIL around calls; the reference's functions are pure and callers may be threaded).  Distinct
  // contexts never share state and run concurrently.  Recursive because pipeline entry points call stage entry points.
  std::recursive_mutex mu;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-trace upload pipelined against the LDE (spg_prove with host buffers): copy stream + per-chunk events
  void* pin = nullptr;              // pinned host staging buffer for the small device->host reads the host waits on (spg_d2h_sync)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[8] = {nullptr};
  cudaEvent_t copy_gate = nullptr;
  // multi-GPU (sharded.cu): NCCL communicator bound at run time, its stream and the pipeline's events
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_ev[16] = {nullptr};
  double last_ms = 0.0;
  uint64_t launches = 0;
  std::string err;
  // NTT tables (device, Montgomery form)
  Fp* tw_fwd = nullptr;   // omega_2048^e, 1024 (SPG_TW_LOG in ntt.cuh)
  Fp* tw_inv = nullptr;   // omega_2048^-e, 1024
  Fp* uniA = nullptr;     // omega_{2^26}^(i << 13), 8192
  Fp* uniB = nullptr;     // omega_{2^26}^i, 8192
  // curve tables
  Fp* const_points = nullptr;   // 506 x (x, y) Montgomery
  Fp* gen_doubles = nullptr;    // G * 2^t, t < 251, (x, y) Montgomery
  std::vector<Fp> h_const_points;   // host copy of const_points
  std::vector<Fp> h_gen_doubles;    // host copy of gen_doubles
  Fp* sqrt_tables = nullptr;        // ECDSA square-root tables: L[256], D[24][256], Dh[24][256]
  std::vector<Fp> h_sqrt_tables;
  // scratch cache
  std::vector<void*> owned;
  bool own_stream = true;
  bool ntt_attr_set = false;
  bool ntt_tma_in = true;           // contiguous passes fetch their tile with one cp.async.bulk (TMA); SPG_NTT_TMA=0: per-thread LDG
  bool ntt_tma_strided = true;      // strided whole-workspace passes gather their tile through a tensor map (TMA); SPG_NTT_TMA2D=0 off
  bool ntt_tma_store = true;        // the strided pass also stores through the tensor map; SPG_NTT_TMA2D_STORE=0 off
  bool deep_pointwise = false;      // SPG_DEEP_POINTWISE=1: combine the 25 extended columns at every point (A/B of fri.cu's coefficient form)
  bool ntt_generic_only = false;    // SPG_NTT_GENERIC=1: route every pass through the generic tile kernel (A/B measurements)
  // growable scratch slots (device), kept until spg_destroy
  void* scratch_p[8] = {nullptr};
  size_t scratch_sz[8] = {0};
  // recycled temporary device buffers (DevBuf): cudaMalloc / cudaFree synchronise the device and cost
  // milliseconds, so temporaries of the entry points are kept and reused (all work is ordered on ctx->stream)
  struct PoolBlock { void* p; size_t size; bool in_use; };
  std::vector<PoolBlock> pool;
  // direct diagonal-twiddle tables of the coset transforms, per (log_n, log_blowup): [2^log_blowup][S][R] (see lde.cu)
  struct DiagTables { int log_n, log_blowup; Fp* t; Fp* t0; };   // t: [2^log_blowup][N] (strided pass), t0: [2^log_blowup][R]
  std::vector<DiagTables> diag_tables;
  // LDE scale tables cached per (log_n, offset, mont): lo[R] , hi[B]
  // inv_diag: direct diagonal table [N] of the inverse transform's strided pass with hi[] folded in (two-pass sizes; else null)
  struct LdeTables { int log_n; uint64_t offset[4]; int mont; Fp* lo; Fp* hi; Fp* inv_diag; };
  std::vector<LdeTables> lde_tables;
  // AIR tables cached per (log_n, chain_log)
  int air_log_n = -1, air_chain_log = -1;
  Fp* air_izt = nullptr;      // [7][4][seg] inverse zerofiers on cosets 0,2,4,6
  Fp* air_plde = nullptr;     // [8][2][512] periodic point columns on the LDE cosets
  Fp* air_ilast = nullptr;    // [4][N] 1 / (x - w^(N-1)) on cosets 0,2,4,6
  // tables of the ECDSA-builtin AIR cached per log_n (air_ecdsa.cu)
  int eair_log_n = -1;
  Fp* eair_izt = nullptr;     // [6][4][256] block-periodic inverse zerofiers on cosets 0,2,4,6
  Fp* eair_plde = nullptr;    // [8][2][256] lane A's periodic point 2^t G on the LDE cosets
  // per-stage device milliseconds of the last pipeline call (spg_stage_ms)
  double stage_ms[16] = {0};
  cudaEvent_t stage_ev[16][2] = {{nullptr}};
  bool stage_used[16] = {false};
};

// stage timers: events on the context stream; spg_stage_collect() after a stream sync fills stage_ms
static inline void spg_stage_begin(spg_ctx* ctx, int s) {
  if (!ctx->stage_ev[s][0]) { cudaEventCreate(&ctx->stage_ev[s][0]); cudaEventCreate(&ctx->stage_ev[s][1]); }
  cudaEventRecord(ctx->stage_ev[s][0], ctx->stream);
  ctx->stage_used[s] = true;
}
static inline void spg_stage_end(spg_ctx* ctx, int s) { cudaEventRecord(ctx->stage_ev[s][1], ctx->stream); }
static inline void spg_stage_reset(spg_ctx* ctx) {
  for (int s = 0; s < 16; s++) { ctx->stage_used[s] = false; ctx->stage_ms[s] = 0.0; }
}
static inline void spg_stage_collect(spg_ctx* ctx) {
  for (int s = 0; s < 16; s++) {
    if (!ctx->stage_used[s]) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->stage_ev[s][0], ctx->stage_ev[s][1]) == cudaSuccess) ctx->stage_ms[s] = ms;
  }
}

// device scratch slot `slot` of at least `bytes` (grown by reallocation; contents not preserved)
static inline cudaError_t spg_scratch(spg_ctx* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_sz[slot] < bytes) {
    if (ctx->scratch_p[slot]) cudaFree(ctx->scratch_p[slot]);
    ctx->scratch_p[slot] = nullptr; ctx->scratch_sz[slot] = 0;
    cudaError_t e = cudaMalloc(&ctx->scratch_p[slot], bytes);
    if (e != cudaSuccess) return e;
    ctx->scratch_sz[slot] = bytes;
  }
  *out = ctx->scratch_p[slot];
  return cudaSuccess;
}

// Device -> host copy the caller waits for (Merkle roots, out-of-domain values, openings: every Fiat-Shamir round trip).
// A copy into pageable memory is staged by the driver and costs tens of microseconds more than one into pinned memory;
// a proof has ~15 of them on its critical path, so they go through one pinned buffer of the context.
#define SPG_PIN_BYTES ((size_t)1 << 20)
static inline cudaError_t spg_d2h_sync(spg_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (!ctx->pin && cudaHostAlloc(&ctx->pin, SPG_PIN_BYTES, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->pin = nullptr; }
  cudaError_t e;
  if (ctx->pin && bytes <= SPG_PIN_BYTES) {
    if ((e = cudaMemcpyAsync(ctx->pin, src, bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    memcpy(dst, ctx->pin, bytes);
    return cudaSuccess;
  }
  if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  return cudaStreamSynchronize(s);
}

#define SPG_LOCK(ctx)                                   \
  std::unique_lock<std::recursive_mutex> spg_lock_;     \
  if (ctx) spg_lock_ = std::unique_lock<std::recursive_mutex>((ctx)->mu)

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

// RAII temporary device buffer, recycled through the context's pool (returned on scope exit, freed by spg_destroy)
struct DevBuf {
  void* p = nullptr;
  spg_ctx* owner = nullptr;
  ~DevBuf() {
    if (!p) return;
    for (auto& b : owner->pool) if (b.p == p) { b.in_use = false; return; }
  }
  cudaError_t alloc(spg_ctx* ctx, size_t bytes) {
    if (bytes < 256) bytes = 256;
    owner = ctx;
    spg_ctx::PoolBlock* best = nullptr;
    for (auto& b : ctx->pool)
      if (!b.in_use && b.size >= bytes && b.size <= 2 * bytes + (1u << 20) && (!best || b.size < best->size)) best = &b;
    if (best) { best->in_use = true; p = best->p; return cudaSuccess; }
    // keep the pool bounded: drop idle blocks before growing it past 64 entries
    if (ctx->pool.size() >= 64) {
      for (size_t i = 0; i < ctx->pool.size();) {
        if (!ctx->pool[i].in_use) { cudaFree(ctx->pool[i].p); ctx->pool.erase(ctx->pool.begin() + i); } else i++;
      }
    }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {   // out of memory: release every idle block and retry once
      for (size_t i = 0; i < ctx->pool.size();) {
        if (!ctx->pool[i].in_use) { cudaFree(ctx->pool[i].p); ctx->pool.erase(ctx->pool.begin() + i); } else i++;
      }
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      if (e != cudaSuccess) { p = nullptr; return e; }
    }
    ctx->pool.push_back({p, bytes, true});
    return cudaSuccess;
  }
  template <class T> T* as() { return (T*)p; }
};

// host helpers (ctx.cu)
Fp spg_host_root_of_unity(int log_n);   // omega_{2^log_n}, Montgomery
Fp spg_host_from_u64(const uint64_t* canon);   // canonical -> Montgomery
void spg_host_to_u64(const Fp& mont, uint64_t* canon);

// ctx.cu: in-place Montgomery -> canonical
int spg_from_mont_device(spg_ctx* ctx, Fp* data, size_t n);
// ntt.cu
int spg_ntt_device(spg_ctx* ctx, const Fp* in, Fp* out, unsigned log_n, size_t ncols, size_t in_stride,
                   size_t out_stride, int inverse, int dit, unsigned long long coset_exp,
                   const Fp* scale_lo, const Fp* scale_hi, const Fp* diag_table = nullptr, const Fp* diag_table0 = nullptr);
// ntt.cu: log2 of the shared-memory workspace (= the largest pass) the NTT kernels use for a 2^log_n transform
int spg_ntt_tile_log_ws(unsigned log_n);
// ntt.cu: fill the direct diagonal table of one pass of a two-pass transform (see ntt.cu)
int spg_ntt_build_diag_table(spg_ctx* ctx, unsigned log_n, int inverse, int dit, int pass_index, unsigned long long coset_exp,
                             const Fp* row_factor, Fp* table);
int spg_bitrev_device(spg_ctx* ctx, const Fp* in, Fp* out, unsigned log_n, size_t ncols);
// lde.cu: device-resident LDE, trace [C][N] -> out [B][C][N]; coeffs (optional) receives the scaled
// coefficient columns g^k c_k (bit-reversed order)
int spg_lde_device(spg_ctx* ctx, const Fp* trace, unsigned log_n, size_t C, unsigned log_blowup,
                   const uint64_t* offset_canon, Fp* out, Fp* coeffs, int mont = 0);
// the two phases; mont != 0 additionally multiplies by R = 2^256 (canonical input -> Montgomery output)
// out_stride: elements between output columns (0 = N: contiguous)
int spg_lde_coeffs_device(spg_ctx* ctx, const Fp* trace, unsigned log_n, size_t C, const uint64_t* offset_canon,
                          Fp* coeffs, int mont = 0, size_t out_stride = 0);
// sharded.cu
void spg_comm_destroy(spg_ctx* ctx);
// out_C / col0: the given columns are columns [col0, col0 + C) of a table with out_C columns per coset (0 = C)
int spg_lde_cosets_device(spg_ctx* ctx, const Fp* coeffs, unsigned log_n, size_t C, unsigned log_blowup, size_t j0,
                          size_t nj, Fp* out, size_t out_C = 0, size_t col0 = 0);

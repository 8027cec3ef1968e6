// Context lifetime, host-side constant tables, field-op entry points and throughput probes.
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ntt.cuh"

// ------------------------------------------------------------------ host helpers
static void exp_root(int log_n, uint32_t e[8]) {
  // (p - 1) >> log_n = 2^(251 - k) + 2^(196 - k) + 2^(192 - k)
  memset(e, 0, 32);
  int bits[3] = {251 - log_n, 196 - log_n, 192 - log_n};
  for (int b : bits) e[b >> 5] |= 1u << (b & 31);
}

Fp spg_host_from_u64(const uint64_t* canon) { return fp_to_mont(fp_from_u64(canon)); }
void spg_host_to_u64(const Fp& mont, uint64_t* canon) {
  Fp c = fp_from_mont(mont);
  fp_to_u64(c, canon);
}

Fp spg_host_root_of_unity(int log_n) {
  uint64_t three[4] = {3, 0, 0, 0};
  Fp g = spg_host_from_u64(three);
  uint32_t e[8];
  exp_root(log_n, e);
  return fp_pow(g, e, 8);
}

static int upload(spg_ctx* ctx, const std::vector<Fp>& h, Fp** d) {
  SPG_CUDA(cudaMalloc((void**)d, h.size() * sizeof(Fp)));
  SPG_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(Fp), cudaMemcpyHostToDevice));
  return SPG_OK;
}

int spg_curve_tables_init(spg_ctx* ctx);   // ec.cu

static int build_tables(spg_ctx* ctx) {
  // intra-tile twiddles omega_{2^SPG_TW_LOG}^(+-e), e < 2^(SPG_TW_LOG - 1)
  Fp w = spg_host_root_of_unity(SPG_TW_LOG);
  Fp wi = fp_inv(w);
  const int ntw = 1 << (SPG_TW_LOG - 1);
  std::vector<Fp> f(ntw), b(ntw);
  f[0] = b[0] = fp_one();
  for (int i = 1; i < ntw; i++) { f[i] = fp_mul(f[i - 1], w); b[i] = fp_mul(b[i - 1], wi); }
  int rc;
  if ((rc = upload(ctx, f, &ctx->tw_fwd))) return rc;
  if ((rc = upload(ctx, b, &ctx->tw_inv))) return rc;
  // universal two-level table of omega_{2^26}
  Fp u = spg_host_root_of_unity(26);
  std::vector<Fp> A(8192), B(8192);
  B[0] = fp_one();
  for (int i = 1; i < 8192; i++) B[i] = fp_mul(B[i - 1], u);
  Fp u13 = fp_mul(B[8191], u);
  A[0] = fp_one();
  for (int i = 1; i < 8192; i++) A[i] = fp_mul(A[i - 1], u13);
  if ((rc = upload(ctx, A, &ctx->uniA))) return rc;
  if ((rc = upload(ctx, B, &ctx->uniB))) return rc;
  return spg_curve_tables_init(ctx);
}

// ------------------------------------------------------------------ C-ABI: lifetime
extern "C" int spg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int spg_create(int device_ordinal, spg_ctx** out) {
  if (!out) return SPG_E_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_ordinal < 0 || device_ordinal >= n)
    return SPG_E_NODEVICE;   // no CPU fallback by design
  spg_ctx* ctx = new spg_ctx();
  ctx->device = device_ordinal;
  auto fail = [&](int rc) { *out = ctx; return rc; };   // caller can still read spg_last_error
  if (cudaSetDevice(device_ordinal) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return fail(SPG_E_CUDA); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { ctx->err = "cudaGetDeviceProperties failed"; return fail(SPG_E_CUDA); }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "stream"; return fail(SPG_E_CUDA); }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  { const char* e = getenv("SPG_NTT_GENERIC"); ctx->ntt_generic_only = e && e[0] == '1'; }
  { const char* e = getenv("SPG_DEEP_POINTWISE"); ctx->deep_pointwise = e && e[0] == '1'; }
  { const char* e = getenv("SPG_NTT_TMA2D_STORE"); ctx->ntt_tma_store = !(e && e[0] == '0'); }   // default on
  { const char* e = getenv("SPG_NTT_TMA2D"); ctx->ntt_tma_strided = !(e && e[0] == '0'); }   // default on; SPG_NTT_TMA2D=0: per-thread loads
  { const char* e = getenv("SPG_NTT_TMA"); ctx->ntt_tma_in = !(e && e[0] == '0'); }     // default on; SPG_NTT_TMA=0: per-thread loads
  int rc = build_tables(ctx);
  if (rc) return fail(rc);
  *out = ctx;
  return SPG_OK;
}

extern "C" void spg_destroy(spg_ctx* ctx) {
  if (!ctx) return;
  if (ctx->device >= 0) cudaSetDevice(ctx->device);
  cudaFree(ctx->tw_fwd); cudaFree(ctx->tw_inv); cudaFree(ctx->uniA); cudaFree(ctx->uniB);
  cudaFree(ctx->const_points); cudaFree(ctx->gen_doubles); cudaFree(ctx->sqrt_tables);
  for (auto& t : ctx->lde_tables) { cudaFree(t.lo); cudaFree(t.hi); cudaFree(t.inv_diag); }
  for (auto& t : ctx->diag_tables) { cudaFree(t.t); cudaFree(t.t0); }
  cudaFree(ctx->air_izt); cudaFree(ctx->air_plde); cudaFree(ctx->air_ilast);
  cudaFree(ctx->eair_izt); cudaFree(ctx->eair_plde);
  for (void* p : ctx->owned) cudaFree(p);
  for (void* p : ctx->scratch_p) cudaFree(p);
  for (auto& b : ctx->pool) cudaFree(b.p);
  spg_comm_destroy(ctx);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& e : ctx->copy_ev) if (e) cudaEventDestroy(e);
  if (ctx->copy_gate) cudaEventDestroy(ctx->copy_gate);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (auto& e : ctx->stage_ev) { if (e[0]) cudaEventDestroy(e[0]); if (e[1]) cudaEventDestroy(e[1]); }
  if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* spg_last_error(spg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" double spg_last_kernel_ms(spg_ctx* ctx) {
  SPG_LOCK(ctx); return ctx ? ctx->last_ms : 0.0; }
extern "C" uint64_t spg_launch_count(spg_ctx* ctx) {
  SPG_LOCK(ctx); return ctx ? ctx->launches : 0; }
extern "C" int spg_set_stream(spg_ctx* ctx, void* cuda_stream) {
  SPG_LOCK(ctx);
  if (!ctx) return SPG_E_ARG;
  SPG_CUDA(cudaSetDevice(ctx->device));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return SPG_OK;
}
extern "C" double spg_stage_ms(spg_ctx* ctx, int stage) {
  SPG_LOCK(ctx);
  return (ctx && stage >= 0 && stage < 16) ? ctx->stage_ms[stage] : 0.0;
}
extern "C" int spg_synchronize(spg_ctx* ctx) {
  SPG_LOCK(ctx);
  SPG_CUDA(cudaSetDevice(ctx->device));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

// ------------------------------------------------------------------ field ops
__global__ void k_field_op(int op, const Fp* a, const Fp* b, Fp* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op >= 5) {   // debug probes: raw Montgomery product / wide product halves / reduce
    Fp r;
    if (op == 5) r = fp_mul(a[i], b[i]);
    else if (op == 8) r = fp_reduce(a[i]);
    else if (op == 9) r = fp_sub_lazy(a[i], b[i], 2);
    else if (op == 10) r = fp_partial(a[i]);
    else if (op == 11) r = fp_reduce_full(a[i]);
    else if (op == 12) r = fp_add_raw(a[i], b[i]);
    else if (op == 13) r = fp_sqr(a[i]);
    else if (op == 14 || op == 15) {
      uint32_t t[16];
      fpd_sqr_wide(t, a[i]);
      for (int k = 0; k < 8; k++) r.v[k] = t[(op == 14 ? 0 : 8) + k];
    } else {
      uint32_t t[16];
      fpd_mul_wide(t, a[i], b[i]);
      for (int k = 0; k < 8; k++) r.v[k] = t[(op == 6 ? 0 : 8) + k];
    }
    out[i] = r;
    return;
  }
  Fp x = fp_to_mont(a[i]);
  Fp r;
  if (op == 3) {
    r = fp_inv(x);
  } else if (op == 4) {
    Fp e = b[i];
    r = fp_pow(x, e.v, 8);
  } else {
    Fp y = fp_to_mont(b[i]);
    r = op == 0 ? fp_mul(x, y) : (op == 1 ? fp_add(x, y) : fp_sub(x, y));
  }
  out[i] = fp_from_mont(r);
}

extern "C" int spg_field_op(spg_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out,
                            size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && a && out && op >= 0 && op <= 15, "spg_field_op");
  SPG_ARG(op == 3 || op == 8 || op == 10 || op == 11 || op >= 13 || b, "spg_field_op: b required");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  const Fp *da = (const Fp*)a, *db = (const Fp*)b;
  Fp* dout = (Fp*)out;
  DevBuf ba, bb, bo;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(ba.alloc(ctx, n * 32)); SPG_CUDA(bb.alloc(ctx, n * 32)); SPG_CUDA(bo.alloc(ctx, n * 32));
    SPG_CUDA(cudaMemcpyAsync(ba.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (b) SPG_CUDA(cudaMemcpyAsync(bb.p, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    da = ba.as<Fp>(); db = bb.as<Fp>(); dout = bo.as<Fp>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_field_op<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(op, da, db, dout, n);
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS))
    SPG_CUDA(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

__global__ void k_from_mont(Fp* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = fp_from_mont(p[i]);
}
int spg_from_mont_device(spg_ctx* ctx, Fp* data, size_t n) {
  if (!n) return SPG_OK;
  k_from_mont<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(data, n);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ throughput probes
template <int CH>
__global__ void __launch_bounds__(256) k_bench_mul(Fp* out, int iters) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x[CH], y;
  y = out[i];                       // opaque operand (the buffer is zero-filled; ptxas cannot fold it)
  y.v[0] ^= i; y.v[3] += 0x9e3779b9u * i; y.v[7] &= 0x07ffffffu;
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = fp_r2(); x[c].v[1] ^= i * 2654435761u + c; }
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = fp_mul(x[c], y);
  }
  Fp acc = x[0];
#pragma unroll
  for (int c = 1; c < CH; c++) acc = fp_add_raw(acc, x[c]);
  if (acc.v[7] == 0x12345678u) out[i] = acc;   // practically never; defeats dead-code elimination
}

// IMAD.WIDE issue rate: two carry-chained rows per iteration (the SPG_ROW_MAD pattern of the multiplication: every
// mad.lo.cc / madc.hi.cc pair becomes one IMAD.WIDE.U32[.X]), each row fed by the other row's previous values so
// that ptxas can neither hoist the products nor turn them into additions.  8 IMAD.WIDE per iteration.
__global__ void __launch_bounds__(256) k_bench_imad_wide(unsigned long long* out, int iters) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[2][9];
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) a[r][k] = (i * 2654435761u) ^ (k * 0x9e3779b9u + r * 0x85ebca6bu);
  for (int it = 0; it < iters; it++) {
    SPG_ROW_MAD(a[0], 0, a[1][0], a[1][2], a[1][4], a[1][6], a[1][7]);
    SPG_ROW_MAD(a[1], 0, a[0][0], a[0][2], a[0][4], a[0][6], a[0][7]);
  }
  uint32_t s = 0;
  for (int r = 0; r < 2; r++) for (int k = 0; k < 9; k++) s ^= a[r][k];
  if (s == 0x12345678u) out[i] = s;
}

extern "C" int spg_bench_field_mul(spg_ctx* ctx, int iters, int chains, double* mul_per_s,
                                   double* imad_wide_per_s) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && iters > 0 && (chains == 1 || chains == 2 || chains == 4), "spg_bench_field_mul");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const int threads = 256, blocks = ctx->sm_count * 8;
  DevBuf o; SPG_CUDA(o.alloc(ctx, (size_t)threads * blocks * 32));
  SPG_CUDA(cudaMemsetAsync(o.p, 0, (size_t)threads * blocks * 32, ctx->stream));
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {   // first repetition warms up
    SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (chains == 1) k_bench_mul<1><<<blocks, threads, 0, ctx->stream>>>(o.as<Fp>(), iters);
    else if (chains == 2) k_bench_mul<2><<<blocks, threads, 0, ctx->stream>>>(o.as<Fp>(), iters);
    else k_bench_mul<4><<<blocks, threads, 0, ctx->stream>>>(o.as<Fp>(), iters);
    SPG_LAUNCH_CHECK();
    SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  }
  ctx->last_ms = ms;
  if (mul_per_s) *mul_per_s = (double)threads * blocks * iters * chains / (ms * 1e-3);
  if (imad_wide_per_s) {
    for (int rep = 0; rep < 2; rep++) {
      SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
      k_bench_imad_wide<<<blocks, threads, 0, ctx->stream>>>((unsigned long long*)o.p, iters);
      SPG_LAUNCH_CHECK();
      SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
      SPG_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    }
    *imad_wide_per_s = (double)threads * blocks * iters * 8.0 / (ms * 1e-3);
  }
  return SPG_OK;
}

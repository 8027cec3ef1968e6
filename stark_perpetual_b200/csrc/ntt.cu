// NTT passes as CUDA kernels + the pass planner + the spg_ntt entry point.
// (SURVEY.md section 8 row p1 / BASELINE.json configs[1]; conventions in DESIGN.md.)
#include <cuda.h>

#include "common.h"
#include "ntt.cuh"

// Tile shapes, chosen by measurement on B200: radix-4 steps (4 elements per thread, 64 registers).  The default
// workspace is 2^10 elements (32 KB, 256 threads, 4 CTAs = 32 warps per SM); transforms of 2^21 and 2^22 points use a
// 2^11 workspace (64 KB, 512 threads, 2 CTAs per SM) so that they stay at two passes.  The radix-8 / 128-register
// shape of the first version is 7-9 % slower at 2^20 and above.
#ifndef NTT_LOG_WS
#define NTT_LOG_WS 10
#endif
#ifndef NTT_LOG_EPT
#define NTT_LOG_EPT 2
#endif

// 2^18 (BASELINE configs[1]) takes a 2^9 workspace so that both of its passes are whole-workspace tiles (k_ntt_tile)
int spg_ntt_tile_log_ws(unsigned log_n) { return (log_n == 21 || log_n == 22) ? 11 : log_n == 18 ? 9 : NTT_LOG_WS; }

// Synchronisation between the phases of a pass.  When the pass uses the whole workspace for one column (log_g = 0,
// log_r = LOG_WS) and every thread owns exactly one butterfly group per step, a step with sh + w <= 5 + LOG_EPT only
// touches rows inside the 2^(5 + LOG_EPT)-row window [warp * 128, warp * 128 + 128): warp w of the CTA reads and writes
// that window and nothing else.  Consecutive such phases therefore need a warp barrier only -- three of the six
// CTA-wide barriers of a 10-bit pass (the load / store phase is mapped onto the same windows).
template <bool DIT, int LOG_WS>
__global__ void __launch_bounds__((1 << LOG_WS) >> NTT_LOG_EPT, LOG_WS == 10 ? 4 : 2) k_ntt_pass(NttPass P) {
  typedef NttTile<LOG_WS, NTT_LOG_EPT> Tile;
  extern __shared__ uint4 smem_raw[];
  FpHalf* ws = reinterpret_cast<FpHalf*>(smem_raw);
  const int tid = threadIdx.x;
  const unsigned cta = blockIdx.x, col = blockIdx.y;
  constexpr int LW = 5 + NTT_LOG_EPT;                       // log2 rows per warp window
  const bool windowed = (P.log_g == 0 && P.log_r == LOG_WS);
  const int ns = Tile::n_steps(P);
  auto local = [&](int k) {                                 // phase k: -1 = load, ns = store, else butterfly step k
    if (!windowed) return false;
    if (k < 0 || k >= ns) return true;
    int w, sh;
    Tile::step_geom<DIT>(P, k, &w, &sh);
    return w == NTT_LOG_EPT && sh + w <= LW;
  };
  auto sync_between = [&](int a, int b) {
    if (local(a) && local(b)) __syncwarp(); else __syncthreads();
  };
#pragma unroll
  for (int j = 0; j < Tile::EPT; j++) {
    const int idx = windowed ? (((tid >> 5) << LW) | (j << 5) | (tid & 31)) : (j * Tile::NT + tid);
    Tile::load_one<DIT>(P, ws, cta, col, idx);
  }
  sync_between(-1, 0);
  for (int k = 0; k < ns; k++) {
    int w, sh;
    Tile::step_geom<DIT>(P, k, &w, &sh);
    Tile::step_w<DIT>(P, ws, tid, w, sh);
    sync_between(k, k + 1);
  }
#pragma unroll
  for (int j = 0; j < Tile::EPT; j++) {
    const int idx = windowed ? (((tid >> 5) << LW) | (j << 5) | (tid & 31)) : (j * Tile::NT + tid);
    Tile::store_one<DIT>(P, ws, cta, col, idx);
  }
}

// The compile-time tile (ntt.cuh NttTileCT): one CTA = one column's tile of 2^LOG_R elements, twiddles staged in shared
// memory after the workspace.  Barriers as in k_ntt_pass: consecutive phases that stay inside a warp's 128-row window
// need a warp barrier only; the first barrier is CTA-wide because every warp reads the staged twiddles.
template <bool DIT, int LOG_R, int K>
__device__ __forceinline__ void ntt_tile_steps(FpHalf* ws, const FpHalf* tws, int tid) {
  typedef NttTileCT<LOG_R> T;
  if constexpr (K < T::NS) {
    T::template step_k<DIT, K>(ws, tws, tid);
    constexpr bool warp_only = T::template step_local<DIT>(K) && T::template step_local<DIT>(K + 1);
    if (warp_only) __syncwarp(); else __syncthreads();
    ntt_tile_steps<DIT, LOG_R, K + 1>(ws, tws, tid);
  }
}

// TMA staging of the tile (default; SPG_NTT_TMA=0 selects the per-thread LDG path).  The tile's 2^LOG_R x 32 bytes are
// fetched by the copy engine into the workspace area in LINEAR element order while the threads stage the twiddles; each
// thread then takes its four elements out of shared memory, applies the load-phase factor and, after a CTA barrier,
// writes them back in the planar swizzled layout the butterflies use.
//   TMA_MODE 1  contiguous pass (log_s = 0): ONE cp.async.bulk of 32 KB, completion on an mbarrier;
//   TMA_MODE 2  strided pass (log_s = LOG_R, B = 1): the tile is column c of an [R][S] matrix of 32-byte elements --
//               four cp.async.bulk.tensor.4d copies of a {32 B x 1 x 256 rows x 1} box through a tensor map over
//               (word, c, r, column), in place of 2048 per-lane 32-byte loads 32 KB apart (32 L1 wavefronts each).
// Measured against the per-thread path in DESIGN.md section 3.
__device__ __forceinline__ uint32_t spg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool DIT, int LOG_R, int TMA_MODE>
__device__ __forceinline__ void ntt_tile_body(const NttPass& P, const CUtensorMap* tmap, const CUtensorMap* tmap_out) {
  typedef NttTileCT<LOG_R> T;
  extern __shared__ __align__(128) uint4 smem_raw[];
  FpHalf* ws = reinterpret_cast<FpHalf*>(smem_raw);
  FpHalf* tws = ws + 2 * T::R;
  const int tid = threadIdx.x;
  const unsigned cta = blockIdx.x, col = blockIdx.y;
  if (TMA_MODE != 0) {
    __shared__ __align__(8) unsigned long long mbar;     // static: placed in front of the 128-byte aligned dynamic area
    const uint32_t mb = spg_smem_u32(&mbar);
    constexpr uint32_t BYTES = (uint32_t)T::R * 32u;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BYTES) : "memory");
      if (TMA_MODE == 1) {
        const Fp* src = P.in + col * P.in_col_stride + ((unsigned long long)cta << LOG_R);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(spg_smem_u32(ws)), "l"(src), "r"(BYTES), "r"(mb) : "memory");
      } else {
#pragma unroll
        for (int q = 0; q < T::R / 256; q++)
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(spg_smem_u32(ws) + (uint32_t)q * 8192u), "l"(tmap), "r"(mb), "r"(0), "r"((int)cta), "r"(q * 256), "r"((int)col)
                       : "memory");
      }
    }
    T::stage_twiddles(P, tws, tid);          // overlaps the copy
    {
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb) : "memory");
    }
    Fp x[4];
    const Fp* lin = reinterpret_cast<const Fp*>(ws);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = T::io_row(tid, j);
      x[j] = lin[r];
      if (DIT) x[j] = (TMA_MODE == 1) ? T::G::apply_factors(P, x[j], cta, r, 0) : T::G::apply_factors(P, x[j], 0, r, cta);
    }
    __syncthreads();                         // every element has left the linear image before the swizzled one overwrites it
#pragma unroll
    for (int j = 0; j < 4; j++) T::ws_put(ws, T::G::swz(T::io_row(tid, j)), x[j]);
  } else {
    T::stage_twiddles(P, tws, tid);
#pragma unroll
    for (int j = 0; j < 4; j++) T::template load<DIT>(P, ws, cta, col, T::io_row(tid, j));
  }
  __syncthreads();
  ntt_tile_steps<DIT, LOG_R, 0>(ws, tws, tid);
  if (TMA_MODE == 2 && tmap_out) {
    // strided store through the tensor map: final values go back to shared memory in linear order, then four tensor copies
    // write the tile's rows (in place of 2048 per-lane 32-byte stores 32 KB apart)
    Fp y[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = T::io_row(tid, j);
      y[j] = T::ws_at(ws, T::G::swz(r));
      if (!DIT) y[j] = T::G::apply_factors(P, y[j], 0, r, cta);
      if (P.final_pass) y[j] = fp_reduce_full(y[j]);
    }
    __syncthreads();                         // the swizzled image has been read by everyone
    Fp* lin = reinterpret_cast<Fp*>(ws);
#pragma unroll
    for (int j = 0; j < 4; j++) lin[T::io_row(tid, j)] = y[j];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < T::R / 256; q++)
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(tmap_out), "r"(spg_smem_u32(ws) + (uint32_t)q * 8192u), "r"(0), "r"((int)cta), "r"(q * 256), "r"((int)col)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; j++) T::template store<DIT>(P, ws, cta, col, T::io_row(tid, j));
}

#ifndef NTT_TILE_MIN_CTAS
#define NTT_TILE_MIN_CTAS 4       // CTAs per SM of the 2^10 tile: 4 = 64 registers (measured against 3 = 80: DESIGN.md section 3)
#endif
// resident CTAs per SM asked of ptxas: 32 warps per SM (64 registers per thread) for every tile size
#define NTT_TILE_CTAS(LOG_R) ((LOG_R) == 10 ? NTT_TILE_MIN_CTAS : (LOG_R) == 9 ? 8 : 2)
template <bool DIT, int LOG_R, bool TMA_IN>
__global__ void __launch_bounds__((1 << LOG_R) / 4, NTT_TILE_CTAS(LOG_R)) k_ntt_tile(NttPass P) {
  typedef NttTileCT<LOG_R> T;
  extern __shared__ uint4 smem_raw[];
  FpHalf* ws = reinterpret_cast<FpHalf*>(smem_raw);
  FpHalf* tws = ws + 2 * T::R;
  const int tid = threadIdx.x;
  const unsigned cta = blockIdx.x, col = blockIdx.y;
  if (TMA_IN) {
    __shared__ __align__(8) unsigned long long mbar;
    const uint32_t mb = spg_smem_u32(&mbar);
    constexpr uint32_t BYTES = (uint32_t)T::R * 32u;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const Fp* src = P.in + col * P.in_col_stride + ((unsigned long long)cta << LOG_R);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(spg_smem_u32(ws)), "l"(src), "r"(BYTES), "r"(mb) : "memory");
    }
    T::stage_twiddles(P, tws, tid);          // overlaps the copy
    {
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb) : "memory");
    }
    Fp x[4];
    const Fp* lin = reinterpret_cast<const Fp*>(ws);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = T::io_row(tid, j);
      x[j] = lin[r];
      if (DIT) x[j] = T::G::apply_factors(P, x[j], cta, r, 0);
    }
    __syncthreads();                         // every element has left the linear image before the swizzled one overwrites it
#pragma unroll
    for (int j = 0; j < 4; j++) T::ws_put(ws, T::G::swz(T::io_row(tid, j)), x[j]);
  } else {
    T::stage_twiddles(P, tws, tid);
#pragma unroll
    for (int j = 0; j < 4; j++) T::template load<DIT>(P, ws, cta, col, T::io_row(tid, j));
  }
  __syncthreads();
  ntt_tile_steps<DIT, LOG_R, 0>(ws, tws, tid);
#pragma unroll
  for (int j = 0; j < 4; j++) T::template store<DIT>(P, ws, cta, col, T::io_row(tid, j));
}

// the strided-pass variant: same tile, gathered through a tensor map (ntt_tile_body<., ., 2>)
template <bool DIT, int LOG_R>
__global__ void __launch_bounds__((1 << LOG_R) / 4, NTT_TILE_CTAS(LOG_R))
    k_ntt_tile_tmap(NttPass P, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, int tma_store) {
  ntt_tile_body<DIT, LOG_R, 2>(P, &tmap, tma_store ? &tmap_out : nullptr);
}

// tensor map over the input of a strided whole-workspace pass (B = 1): dims (innermost first) 8 words, S columns, R rows,
// ncols trace columns; box {8, 1, 256, 1}.  Returns false when the driver entry point is unavailable.
typedef CUresult (*spg_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_strided_tmap(const NttPass& P, size_t ncols, CUtensorMap* out, bool for_output = false) {
  static spg_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (spg_encode_tiled_fn)p;
  }
  if (!fn) return false;
  const cuuint64_t S = 1ull << P.log_s, R = 1ull << P.log_r;
  const cuuint64_t dims[4] = {8, S, R, (cuuint64_t)ncols};
  const cuuint64_t strides[3] = {32, S * 32, (cuuint64_t)(for_output ? P.out_col_stride : P.in_col_stride) * 32};   // bytes, dims 1..3
  const cuuint32_t box[4] = {8, 1, 256, 1}, estr[4] = {1, 1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, for_output ? (void*)P.out : (void*)P.in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void k_bitrev(const Fp* in, Fp* out, unsigned log_n, size_t ncols) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)1 << log_n;
  if (i >= n * ncols) return;
  size_t col = i >> log_n, r = i & (n - 1);
  out[col * n + spg_bitrev((unsigned)r, (int)log_n)] = in[i];
}

// diag_table[(c << log_r) | r] = omega_{2^26}^(+- bitrev_R(r) * (c * ec + e0)) for the pass P, optionally times
// row_factor[r] (the LDE folds the per-tile part of its g^k scaling into the inverse transform's table)
__global__ void k_build_diag_table(NttPass P, Fp* __restrict__ table, const Fp* __restrict__ row_factor) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ((size_t)1 << (P.log_r + P.log_s))) return;
  Fp v = NttTile<NTT_LOG_WS, NTT_LOG_EPT>::diag_entry(P, idx);
  if (row_factor) v = fp_mul(v, row_factor[idx & ((1ull << P.log_r) - 1)]);
  table[idx] = fp_reduce(v);
}

// Direct diagonal table of pass `pass_index` (execution order) of a two-pass transform: forward DIT with coset exponent
// coset_exp (pass 0: the contiguous pass, 2^log_r entries; pass 1: the strided pass, N entries), or inverse DIF
// (pass 0: the strided pass, N entries).  row_factor: optional device table indexed by the row r of that pass.
int spg_ntt_build_diag_table(spg_ctx* ctx, unsigned log_n, int inverse, int dit, int pass_index, unsigned long long coset_exp,
                             const Fp* row_factor, Fp* table) {
  NttPass passes[8];
  const int np = spg_ntt_make_passes(passes, spg_ntt_tile_log_ws(log_n), nullptr, nullptr, log_n, 0, 0, inverse, dit, coset_exp,
                                     nullptr, nullptr, ctx->tw_fwd, ctx->tw_inv, ctx->uniA, ctx->uniB);
  SPG_ARG(np == 2 && pass_index >= 0 && pass_index < 2, "diag table: two-pass transforms only");
  const NttPass& P = passes[pass_index];
  SPG_ARG(P.use_diag, "diag table: pass has no diagonal factor");
  const size_t total = (size_t)1 << (P.log_r + P.log_s);
  k_build_diag_table<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(P, table, row_factor);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// launch of the compile-time tile for one pass: tensor-map kernel for strided passes, bulk-copy kernel for contiguous ones,
// per-thread loads when TMA is switched off
template <int LOG_R>
static void launch_tile(spg_ctx* ctx, const NttPass& P, dim3 grid, int dit, bool tma2d, const CUtensorMap& tin, const CUtensorMap& tout) {
  const int threads = (1 << LOG_R) / 4, smem = (1 << LOG_R) * 48, st = ctx->ntt_tma_store ? 1 : 0;
  if (tma2d) {
    if (dit) k_ntt_tile_tmap<true, LOG_R><<<grid, threads, smem, ctx->stream>>>(P, tin, tout, st);
    else k_ntt_tile_tmap<false, LOG_R><<<grid, threads, smem, ctx->stream>>>(P, tin, tout, st);
  } else if (ctx->ntt_tma_in && P.log_s == 0) {
    if (dit) k_ntt_tile<true, LOG_R, true><<<grid, threads, smem, ctx->stream>>>(P);
    else k_ntt_tile<false, LOG_R, true><<<grid, threads, smem, ctx->stream>>>(P);
  } else {
    if (dit) k_ntt_tile<true, LOG_R, false><<<grid, threads, smem, ctx->stream>>>(P);
    else k_ntt_tile<false, LOG_R, false><<<grid, threads, smem, ctx->stream>>>(P);
  }
}
template <int LOG_R>
static cudaError_t tile_attrs() {
  const int smem = (1 << LOG_R) * 48;
  cudaError_t e = cudaSuccess;
  auto set = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); };
  set((const void*)k_ntt_tile<false, LOG_R, false>); set((const void*)k_ntt_tile<true, LOG_R, false>);
  set((const void*)k_ntt_tile<false, LOG_R, true>); set((const void*)k_ntt_tile<true, LOG_R, true>);
  set((const void*)k_ntt_tile_tmap<false, LOG_R>); set((const void*)k_ntt_tile_tmap<true, LOG_R>);
  return e;
}

int spg_ntt_device(spg_ctx* ctx, const Fp* in, Fp* out, unsigned log_n, size_t ncols, size_t in_stride,
                   size_t out_stride, int inverse, int dit, unsigned long long coset_exp,
                   const Fp* scale_lo, const Fp* scale_hi, const Fp* diag_table, const Fp* diag_table0) {
  SPG_ARG(log_n <= 26, "NTT size above 2^26 not supported by the universal twiddle table");
  SPG_ARG(ncols < 65536, "too many columns in one NTT batch");
  if (ncols == 0) return SPG_OK;
  const int log_ws = spg_ntt_tile_log_ws(log_n);
  bool& attr_set = ctx->ntt_attr_set;
  if (!attr_set) {
    SPG_CUDA(cudaFuncSetAttribute(k_ntt_pass<false, NTT_LOG_WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << NTT_LOG_WS) * 32));
    SPG_CUDA(cudaFuncSetAttribute(k_ntt_pass<true, NTT_LOG_WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << NTT_LOG_WS) * 32));
    SPG_CUDA(cudaFuncSetAttribute(k_ntt_pass<false, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 11) * 32));
    SPG_CUDA(cudaFuncSetAttribute(k_ntt_pass<true, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 11) * 32));
    SPG_CUDA(tile_attrs<9>()); SPG_CUDA(tile_attrs<10>()); SPG_CUDA(tile_attrs<11>());
    attr_set = true;
  }
  NttPass passes[8];
  if (!dit && coset_exp != 0) { ctx->err = "coset shift only supported for DIT"; return SPG_E_ARG; }
  const int np = spg_ntt_make_passes(passes, log_ws, in, out, log_n, in_stride, out_stride, inverse, dit,
                                     coset_exp, scale_lo, scale_hi, ctx->tw_fwd, ctx->tw_inv, ctx->uniA, ctx->uniB);
  // direct diagonal tables (two-pass transforms): forward DIT -- diag_table0 for the contiguous first pass, diag_table for
  // the strided second pass; inverse DIF -- diag_table for the strided first pass
  if (np == 2 && dit && !inverse) {
    if (diag_table) passes[1].diag_table = diag_table;
    if (diag_table0 && passes[0].use_diag) passes[0].diag_table = diag_table0;
  } else if (np == 2 && !dit && inverse && diag_table) {
    passes[0].diag_table = diag_table;
  }
  const int gen_ws = log_ws == 9 ? NTT_LOG_WS : log_ws;     // the generic kernel exists for the 2^10 and 2^11 workspaces
  const int smem = (1 << gen_ws) * (int)sizeof(Fp), threads = (1 << gen_ws) >> NTT_LOG_EPT;
  for (int pi = 0; pi < np; pi++) {
    const NttPass& P = passes[pi];
    const size_t ctas = ((size_t)1 << log_n) >> (P.log_r + P.log_g);
    dim3 grid((unsigned)ctas, (unsigned)ncols);
    if (!ctx->ntt_generic_only && P.log_r == log_ws && P.log_g == 0) {     // the compile-time tile: whole-workspace tiles
      CUtensorMap tmap, tmap_out;
      const bool tma2d = ctx->ntt_tma_strided && P.log_s == P.log_r && (int)log_n == 2 * P.log_r && make_strided_tmap(P, ncols, &tmap) &&
                         make_strided_tmap(P, ncols, &tmap_out, true);
      if (log_ws == 11) launch_tile<11>(ctx, P, grid, dit, tma2d, tmap, tmap_out);
      else if (log_ws == 9) launch_tile<9>(ctx, P, grid, dit, tma2d, tmap, tmap_out);
      else launch_tile<NTT_LOG_WS>(ctx, P, grid, dit, tma2d, tmap, tmap_out);
    } else if (log_ws == 11) {
      if (dit) k_ntt_pass<true, 11><<<grid, threads, smem, ctx->stream>>>(P);
      else k_ntt_pass<false, 11><<<grid, threads, smem, ctx->stream>>>(P);
    } else {
      if (dit) k_ntt_pass<true, NTT_LOG_WS><<<grid, threads, smem, ctx->stream>>>(P);
      else k_ntt_pass<false, NTT_LOG_WS><<<grid, threads, smem, ctx->stream>>>(P);
    }
    SPG_LAUNCH_CHECK();
  }
  return SPG_OK;
}

int spg_bitrev_device(spg_ctx* ctx, const Fp* in, Fp* out, unsigned log_n, size_t ncols) {
  size_t total = ((size_t)1 << log_n) * ncols;
  k_bitrev<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(in, out, log_n, ncols);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

__global__ void k_fill(Fp* p, Fp v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" int spg_ntt(spg_ctx* ctx, uint64_t* data, unsigned log_n, size_t batch, int inverse, int order,
                       int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && data, "spg_ntt: null");
  SPG_ARG(order >= 0 && order <= 2, "spg_ntt: order");
  SPG_ARG(log_n <= 26, "spg_ntt: log_n");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (batch == 0) return SPG_OK;
  const size_t n = (size_t)1 << log_n, total = n * batch;
  Fp* d = (Fp*)data;
  DevBuf buf, tmp, sc;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(buf.alloc(ctx, total * 32));
    SPG_CUDA(cudaMemcpyAsync(buf.p, data, total * 32, cudaMemcpyHostToDevice, ctx->stream));
    d = buf.as<Fp>();
  }
  // inverse: scale by 1/N through the scale_hi hook of the contiguous pass
  const Fp* scale_hi = nullptr;
  if (inverse) {
    int lr, lb;
    spg_ntt_last_pass_geometry(log_n, spg_ntt_tile_log_ws(log_n), &lr, &lb);
    uint64_t nn[4] = {(uint64_t)n, 0, 0, 0};
    Fp ninv = fp_inv(spg_host_from_u64(nn));
    SPG_CUDA(sc.alloc(ctx, ((size_t)1 << lb) * 32));
    k_fill<<<(unsigned)((((size_t)1 << lb) + 255) / 256), 256, 0, ctx->stream>>>(sc.as<Fp>(), ninv, (size_t)1 << lb);
    SPG_LAUNCH_CHECK();
    scale_hi = sc.as<Fp>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (size_t c0 = 0; c0 < batch; c0 += 32768) {
    size_t nc = batch - c0 < 32768 ? batch - c0 : 32768;
    int rc = spg_ntt_device(ctx, d + c0 * n, d + c0 * n, log_n, nc, n, n, inverse, order == SPG_NTT_REV_TO_NAT, 0,
                            nullptr, scale_hi);
    if (rc) return rc;
  }
  if (order == SPG_NTT_NAT_TO_NAT && log_n > 0) {
    SPG_CUDA(tmp.alloc(ctx, total * 32));
    int rc = spg_bitrev_device(ctx, d, tmp.as<Fp>(), log_n, batch);
    if (rc) return rc;
    SPG_CUDA(cudaMemcpyAsync(d, tmp.p, total * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS))
    SPG_CUDA(cudaMemcpyAsync(data, d, total * 32, cudaMemcpyDeviceToHost, ctx->stream));
  if ((flags & SPG_NO_SYNC) && (flags & SPG_DEVICE_PTRS) && !inverse && order != SPG_NTT_NAT_TO_NAT) return SPG_OK;
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

// Low-degree extension of trace columns (SURVEY.md section 8 row p2 / BASELINE.json configs[2]).
// No reference symbol exists for this stage; conventions are this repo's own (DESIGN.md "LDE"):
//   column values are evaluations on <w_N> in natural order; the extension is evaluated on the B
//   cosets  g * w_{BN}^j * <w_N>,  j = 0..B-1 (g = 3 by default = FIELD_GEN, signature.py:42), each
//   in natural order, stored coset-major:  out[j][c][i] = f_c(g * w_{BN}^j * w_N^i).
//
// Data flow per column:  inverse DIF (natural -> bit-reversed) with the factors g^k / N folded into
// the store phase of its last pass, then per coset one forward DIT (bit-reversed -> natural) whose
// load phase applies w_{BN}^(j k).  No permutation or scaling pass touches HBM on its own.
#include <string.h>

#include "common.h"
#include "ntt.cuh"

static bool two_pass(unsigned log_n) {
  const unsigned ws = (unsigned)spg_ntt_tile_log_ws(log_n);
  return log_n > ws && log_n <= 2 * ws;
}

// lo[r] (contiguous-pass row factor), hi[b] (per-tile factor) and, for two-pass sizes, inv_diag: the inverse transform's
// inter-pass diagonal twiddle as a direct N-entry table with hi[] folded in -- the tile a value lands in is the row of
// the strided pass it leaves, so the g^k / N scaling costs ONE multiplication per element (lo) instead of two, and the
// diagonal twiddle one instead of two (no two-level lookup).  *hi_out is null when inv_diag carries it.
static int ensure_scale_tables(spg_ctx* ctx, unsigned log_n, const uint64_t* offset, int mont, const Fp** lo_out,
                               const Fp** hi_out, const Fp** inv_diag_out) {
  for (auto& t : ctx->lde_tables)
    if (t.log_n == (int)log_n && t.mont == mont && memcmp(t.offset, offset, 32) == 0) {
      *lo_out = t.lo; *hi_out = t.inv_diag ? nullptr : t.hi; *inv_diag_out = t.inv_diag;
      return SPG_OK;
    }
  std::vector<Fp> lo, hi;
  spg_lde_scale_tables(log_n, spg_ntt_tile_log_ws(log_n), spg_host_from_u64(offset), lo, hi);
  if (mont) for (auto& v : lo) v = fp_mul(v, fp_r2());     // extra factor R: canonical in -> Montgomery out
  const size_t R = lo.size(), B = hi.size();
  if (ctx->lde_tables.size() >= 8) {   // tiny cache: drop the oldest entry
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->lde_tables[0].lo); cudaFree(ctx->lde_tables[0].hi); cudaFree(ctx->lde_tables[0].inv_diag);
    ctx->lde_tables.erase(ctx->lde_tables.begin());
  }
  spg_ctx::LdeTables t;
  t.log_n = (int)log_n; t.mont = mont; memcpy(t.offset, offset, 32); t.lo = t.hi = t.inv_diag = nullptr;
  SPG_CUDA(cudaMalloc((void**)&t.lo, R * sizeof(Fp)));
  SPG_CUDA(cudaMalloc((void**)&t.hi, B * sizeof(Fp)));
  SPG_CUDA(cudaMemcpy(t.lo, lo.data(), R * sizeof(Fp), cudaMemcpyHostToDevice));
  SPG_CUDA(cudaMemcpy(t.hi, hi.data(), B * sizeof(Fp), cudaMemcpyHostToDevice));
  if (two_pass(log_n) && cudaMalloc((void**)&t.inv_diag, ((size_t)1 << log_n) * sizeof(Fp)) == cudaSuccess) {
    int rc = spg_ntt_build_diag_table(ctx, log_n, /*inverse=*/1, /*dit=*/0, /*pass=*/0, 0, t.hi, t.inv_diag);
    if (rc) { cudaFree(t.lo); cudaFree(t.hi); cudaFree(t.inv_diag); return rc; }
  } else {
    cudaGetLastError();
    t.inv_diag = nullptr;
  }
  ctx->lde_tables.push_back(t);
  *lo_out = t.lo; *hi_out = t.inv_diag ? nullptr : t.hi; *inv_diag_out = t.inv_diag;
  return SPG_OK;
}

// phase A: columns -> scaled coefficient columns  g^k c_k  (bit-reversed order)
int spg_lde_coeffs_device(spg_ctx* ctx, const Fp* trace, unsigned log_n, size_t C, const uint64_t* offset_canon,
                          Fp* coeffs, int mont, size_t out_stride) {
  static const uint64_t three[4] = {3, 0, 0, 0};
  const uint64_t* off = offset_canon ? offset_canon : three;
  const Fp *lo, *hi, *inv_diag;
  int rc = ensure_scale_tables(ctx, log_n, off, mont, &lo, &hi, &inv_diag);
  if (rc) return rc;
  const size_t n = (size_t)1 << log_n;
  if (out_stride == 0) out_stride = n;
  for (size_t c0 = 0; c0 < C; c0 += 32768) {
    const size_t nc = C - c0 < 32768 ? C - c0 : 32768;
    rc = spg_ntt_device(ctx, trace + c0 * n, coeffs + c0 * out_stride, log_n, nc, n, out_stride, /*inverse=*/1, /*dit=*/0, 0,
                        lo, hi, inv_diag);
    if (rc) return rc;
  }
  return SPG_OK;
}

// Direct diagonal-twiddle tables for the coset transforms of a two-pass LDE, built once per (log_n, log_blowup) and kept in
// the context: per coset one N-entry table for the strided second pass (8 x 32 MB at 2^20; a coset's table is re-read by
// every column of the launch and stays L2-resident) and one R-entry table for the contiguous first pass (the within-tile
// part of the coset shift w_{BN}^(j k); its per-tile part is already inside the second table).  They replace the
// two-level lookup's multiplication in the load phase of both passes: one multiplication per element per pass.
static int ensure_diag_tables(spg_ctx* ctx, unsigned log_n, unsigned log_blowup, const Fp** out, const Fp** out0, size_t* r0) {
  *out = *out0 = nullptr; *r0 = 0;
  if (!two_pass(log_n) || log_blowup > 3) return SPG_OK;      // single-pass or three-pass sizes: two-level lookup
  int lr, lb;
  spg_ntt_last_pass_geometry(log_n, spg_ntt_tile_log_ws(log_n), &lr, &lb);
  const size_t n = (size_t)1 << log_n, nb = (size_t)1 << log_blowup, R = (size_t)1 << lr;
  *r0 = R;
  for (auto& t : ctx->diag_tables)
    if (t.log_n == (int)log_n && t.log_blowup == (int)log_blowup) { *out = t.t; *out0 = t.t0; return SPG_OK; }
  if (ctx->diag_tables.size() >= 2) {
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->diag_tables[0].t); cudaFree(ctx->diag_tables[0].t0);
    ctx->diag_tables.erase(ctx->diag_tables.begin());
  }
  spg_ctx::DiagTables t;
  t.log_n = (int)log_n; t.log_blowup = (int)log_blowup; t.t = t.t0 = nullptr;
  if (cudaMalloc((void**)&t.t, nb * n * sizeof(Fp)) != cudaSuccess) { cudaGetLastError(); return SPG_OK; }   // optional
  if (cudaMalloc((void**)&t.t0, nb * R * sizeof(Fp)) != cudaSuccess) { cudaGetLastError(); cudaFree(t.t); return SPG_OK; }
  for (size_t j = 0; j < nb; j++) {
    const unsigned long long coset_exp = (unsigned long long)j << (SPG_UNI_LOG - (log_n + log_blowup));
    int rc = spg_ntt_build_diag_table(ctx, log_n, 0, 1, 1, coset_exp, nullptr, t.t + j * n);
    // coset 0 has no shift: its contiguous pass carries no factor at all (use_diag = 0)
    if (!rc && j) rc = spg_ntt_build_diag_table(ctx, log_n, 0, 1, 0, coset_exp, nullptr, t.t0 + j * R);
    if (rc) { cudaFree(t.t); cudaFree(t.t0); return rc; }
  }
  ctx->diag_tables.push_back(t);
  *out = t.t; *out0 = t.t0;
  return SPG_OK;
}

// phase B: scaled coefficients -> evaluations on cosets [j0, j0 + nj) of the 2^log_blowup cosets;
// out[(j - j0)][c][i]
int spg_lde_cosets_device(spg_ctx* ctx, const Fp* coeffs, unsigned log_n, size_t C, unsigned log_blowup, size_t j0,
                          size_t nj, Fp* out, size_t out_C, size_t col0) {
  if (out_C == 0) out_C = C;     // the columns given are the whole table
  SPG_ARG(log_n + log_blowup <= SPG_UNI_LOG, "spg_lde: log_n + log_blowup above 26");
  SPG_ARG(j0 + nj <= ((size_t)1 << log_blowup), "spg_lde: coset range");
  const size_t n = (size_t)1 << log_n;
  const Fp *diag = nullptr, *diag0 = nullptr;
  size_t r0 = 0;
  int rc0 = ensure_diag_tables(ctx, log_n, log_blowup, &diag, &diag0, &r0);
  if (rc0) return rc0;
  for (size_t j = j0; j < j0 + nj; j++) {
    const unsigned long long coset_exp = (unsigned long long)j << (SPG_UNI_LOG - (log_n + log_blowup));
    for (size_t c0 = 0; c0 < C; c0 += 32768) {
      const size_t nc = C - c0 < 32768 ? C - c0 : 32768;
      int rc = spg_ntt_device(ctx, coeffs + c0 * n, out + ((j - j0) * out_C + col0 + c0) * n, log_n, nc, n, n, /*inverse=*/0,
                              /*dit=*/1, coset_exp, nullptr, nullptr, diag ? diag + j * n : nullptr,
                              (diag0 && j) ? diag0 + j * r0 : nullptr);
      if (rc) return rc;
    }
  }
  return SPG_OK;
}

int spg_lde_device(spg_ctx* ctx, const Fp* trace, unsigned log_n, size_t C, unsigned log_blowup,
                   const uint64_t* offset_canon, Fp* out, Fp* coeffs, int mont) {
  SPG_ARG(log_n + log_blowup <= SPG_UNI_LOG, "spg_lde: log_n + log_blowup above 26");
  const size_t n = (size_t)1 << log_n;
  if (!coeffs) {
    void* p;
    SPG_CUDA(spg_scratch(ctx, 0, C * n * sizeof(Fp), &p));
    coeffs = (Fp*)p;
  }
  int rc = spg_lde_coeffs_device(ctx, trace, log_n, C, offset_canon, coeffs, mont);
  if (rc) return rc;
  return spg_lde_cosets_device(ctx, coeffs, log_n, C, log_blowup, 0, (size_t)1 << log_blowup, out);
}

// common tail of the device-pointer entry points
static int finish(spg_ctx* ctx, int flags) {
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_NO_SYNC)) {
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
    spg_stage_collect(ctx);
  }
  return SPG_OK;
}

extern "C" int spg_lde_coeffs(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, size_t n_cols,
                              const uint64_t* coset_offset, uint64_t* coeffs, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && trace && coeffs, "spg_lde_coeffs: null");
  SPG_ARG(flags & SPG_DEVICE_PTRS, "spg_lde_coeffs: device pointers only");
  SPG_ARG(log_n <= SPG_UNI_LOG, "spg_lde_coeffs: log_n");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n_cols == 0) return SPG_OK;
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = spg_lde_coeffs_device(ctx, (const Fp*)trace, log_n, n_cols, coset_offset, (Fp*)coeffs, (flags & SPG_MONT_OUT) ? 1 : 0);
  if (rc) return rc;
  return finish(ctx, flags);
}

extern "C" int spg_lde_cosets(spg_ctx* ctx, const uint64_t* coeffs, unsigned log_n, size_t n_cols,
                              unsigned log_blowup, size_t coset_begin, size_t coset_count, uint64_t* out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && coeffs && out, "spg_lde_cosets: null");
  SPG_ARG(flags & SPG_DEVICE_PTRS, "spg_lde_cosets: device pointers only");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n_cols == 0 || coset_count == 0) return SPG_OK;
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = spg_lde_cosets_device(ctx, (const Fp*)coeffs, log_n, n_cols, log_blowup, coset_begin, coset_count, (Fp*)out);
  if (rc) return rc;
  return finish(ctx, flags);
}

extern "C" int spg_lde(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, size_t n_cols, unsigned log_blowup,
                       const uint64_t* coset_offset, uint64_t* out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && trace && out, "spg_lde: null");
  SPG_ARG(log_n + log_blowup <= SPG_UNI_LOG && log_blowup <= 6, "spg_lde: size");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n_cols == 0) return SPG_OK;
  const size_t n = (size_t)1 << log_n, in_bytes = n_cols * n * 32, out_bytes = in_bytes << log_blowup;
  const Fp* din = (const Fp*)trace;
  Fp* dout = (Fp*)out;
  DevBuf bi, bo;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(bi.alloc(ctx, in_bytes)); SPG_CUDA(bo.alloc(ctx, out_bytes));
    SPG_CUDA(cudaMemcpyAsync(bi.p, trace, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    din = bi.as<Fp>(); dout = bo.as<Fp>();
  }
  spg_stage_reset(ctx);
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  void* cf;
  SPG_CUDA(spg_scratch(ctx, 0, in_bytes, &cf));
  spg_stage_begin(ctx, 0);
  int rc = spg_lde_coeffs_device(ctx, din, log_n, n_cols, coset_offset, (Fp*)cf, 0);
  if (rc) return rc;
  spg_stage_end(ctx, 0);
  spg_stage_begin(ctx, 1);
  rc = spg_lde_cosets_device(ctx, (const Fp*)cf, log_n, n_cols, log_blowup, 0, (size_t)1 << log_blowup, dout);
  if (rc) return rc;
  spg_stage_end(ctx, 1);
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS))
    SPG_CUDA(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (!((flags & SPG_NO_SYNC) && (flags & SPG_DEVICE_PTRS))) {
    SPG_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
    spg_stage_collect(ctx);
  }
  return SPG_OK;
}

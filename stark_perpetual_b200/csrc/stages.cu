// Stage-level C-ABI of the prover (device pointers only): the building blocks the multi-GPU host driver
// (stark_perpetual_b200/prover.py) sequences around its collectives.  Each GPU owns `n_cosets` consecutive
// cosets of the 8, starting at `first_coset`; tables passed here hold exactly those cosets.
// Scalars cross the boundary as canonical felts (4 x u64, host memory).  DESIGN.md "Multi-GPU".
#include <string.h>

#include "blake2s.cuh"
#include "stark_kernels.cuh"

static Fp gen_mont() { uint64_t three[4] = {3, 0, 0, 0}; return spg_host_from_u64(three); }

extern "C" int spg_stage_merkle(spg_ctx* ctx, const uint64_t* table, size_t n_cols, size_t rows, int n_cosets,
                                uint8_t* tree_out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && table && tree_out, "spg_stage_merkle: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  return spg_merkle_build_device(ctx, (const Fp*)table, (int)n_cols, rows, (uint32_t*)tree_out, n_cosets);
}

extern "C" int spg_stage_air(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const uint64_t* t_lde, int first_coset,
                             int jj0, int n_even, const uint64_t* x0, const uint64_t* outs, const uint64_t* alpha,
                             uint64_t* cp_out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && t_lde && x0 && outs && alpha && cp_out, "spg_stage_air: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  AirPublic pub;
  for (int l = 0; l < SPG_AIR_LANES; l++) { pub.x0[l] = spg_host_from_u64(x0 + 4 * l); pub.outs[l] = spg_host_from_u64(outs + 4 * l); }
  Fp apows[SPG_AIR_LANES * SPG_AIR_NCONSTR];
  const Fp a = spg_host_from_u64(alpha);
  apows[0] = fp_one();
  for (int k = 1; k < SPG_AIR_LANES * SPG_AIR_NCONSTR; k++) apows[k] = fp_mul(apows[k - 1], a);
  return spg_air_eval_device(ctx, log_n, chain_log, (const Fp*)t_lde, pub, apows, (Fp*)cp_out, first_coset, jj0, n_even);
}

extern "C" int spg_stage_cp_split(spg_ctx* ctx, unsigned log_n, const uint64_t* cp, int jj0, int n_even, uint64_t* hev) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && cp && hev, "spg_stage_cp_split: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  return spg_cp_split_device(ctx, log_n, (const Fp*)cp, (Fp*)hev, jj0, n_even);
}

// cols: host array of n_items device pointers (scaled, bit-reversed coefficient columns); pts: [n_pts][4]
// canonical; out: [n_items][4] canonical.  Evaluates column k at pts[pt_idx[k]] / g.
extern "C" int spg_stage_poly_eval(spg_ctx* ctx, unsigned log_n, const uint64_t* const* cols, const int* pt_idx,
                                   int n_items, const uint64_t* pts, int n_pts, uint64_t* out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && cols && pt_idx && pts && out && n_items > 0 && n_items <= 256 && n_pts > 0 && n_pts <= 16, "spg_stage_poly_eval");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const Fp ginv = fp_inv(gen_mont());
  std::vector<Fp> p(n_pts), res(n_items);
  for (int k = 0; k < n_pts; k++) p[k] = fp_mul(spg_host_from_u64(pts + 4 * k), ginv);
  int rc = spg_poly_eval_device(ctx, log_n, (const Fp* const*)cols, pt_idx, n_items, p.data(), n_pts, res.data());
  if (rc) return rc;
  for (int k = 0; k < n_items; k++) spg_host_to_u64(res[k], out + 4 * k);
  return SPG_OK;
}

// DEEP quotient on the local cosets.  oods: [54][4] canonical; inv_scratch: device, 3 * n_cosets * N felts.
extern "C" int spg_stage_deep(spg_ctx* ctx, unsigned log_n, const uint64_t* t_lde, const uint64_t* h_lde, int first_coset,
                              int n_cosets, const uint64_t* z, const uint64_t* gamma, const uint64_t* oods,
                              uint64_t* inv_scratch, uint64_t* out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && t_lde && h_lde && z && gamma && oods && inv_scratch && out, "spg_stage_deep: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const Fp zz = spg_host_from_u64(z), gm = spg_host_from_u64(gamma);
  Fp o[SPG_N_OODS];
  for (int k = 0; k < SPG_N_OODS; k++) o[k] = spg_host_from_u64(oods + 4 * k);
  void* ds;
  SPG_CUDA(spg_scratch(ctx, 6, 64 * sizeof(Fp), &ds));
  return spg_deep_stage_device(ctx, log_n, (const Fp*)t_lde, (const Fp*)h_lde, first_coset, n_cosets, zz, gm, o, (Fp*)inv_scratch,
                               (Fp*)ds, (Fp*)out);
}

// fold layer `layer_index` (0 = the DEEP quotient, rows = 2^log_rows per coset) by 8 with challenge beta
extern "C" int spg_stage_fri_fold(spg_ctx* ctx, const uint64_t* in, unsigned log_rows, int first_coset, int n_cosets,
                                  const uint64_t* beta, int layer_index, uint64_t* out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && in && beta && out && layer_index >= 0 && layer_index < 16, "spg_stage_fri_fold");
  SPG_CUDA(cudaSetDevice(ctx->device));
  Fp g_l = gen_mont();
  for (int k = 0; k < 3 * layer_index; k++) g_l = fp_sqr(g_l);
  return spg_fri_fold8_device(ctx, (const Fp*)in, log_rows, fp_mul(spg_host_from_u64(beta), fp_inv(g_l)), (Fp*)out,
                              first_coset, n_cosets);
}

// open `count` leaves (LOCAL leaf indices, host array) of a local table: leaves_out [count][8 n_cols 32] bytes,
// paths_out [count][levels 32] bytes with levels = log2(n_cosets rows / 8)   (host buffers)
extern "C" int spg_stage_open(spg_ctx* ctx, const uint64_t* table, size_t n_cols, size_t rows, int n_cosets,
                              const uint8_t* tree, const uint32_t* idx, int count, uint8_t* leaves_out, uint8_t* paths_out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && table && tree && idx && leaves_out && paths_out && count >= 0, "spg_stage_open");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (count == 0) return SPG_OK;
  const size_t n_leaves = (rows >> 3) * n_cosets;
  int levels = 0;
  while (((size_t)1 << levels) < n_leaves) levels++;
  const size_t lw = 8 * n_cols * 8, pw = (size_t)levels * 8;
  DevBuf di, dl, dp;
  SPG_CUDA(di.alloc(ctx, count * 4)); SPG_CUDA(dl.alloc(ctx, count * lw * 4)); SPG_CUDA(dp.alloc(ctx, count * pw * 4 + 16));
  SPG_CUDA(cudaMemcpyAsync(di.p, idx, count * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = spg_merkle_open_device(ctx, (const Fp*)table, (int)n_cols, rows, (const uint32_t*)tree, di.as<uint32_t>(), count,
                                  dl.as<uint32_t>(), dp.as<uint32_t>(), n_cosets);
  if (rc) return rc;
  SPG_CUDA(cudaMemcpyAsync(leaves_out, dl.p, count * lw * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (levels) SPG_CUDA(cudaMemcpyAsync(paths_out, dp.p, count * pw * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

// Prover self-check at the out-of-domain point (host arithmetic, the same routine spg_prove uses): recompute the
// composition from the 25 + 25 trace values at z and z*w and compare with sum z^m H_m(z^4).  oods: [54][4]
// canonical.  SPG_E_PROOF when the trace does not satisfy the AIR.
extern "C" int spg_stage_check_oods(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const uint64_t* x0, const uint64_t* outs,
                                    const uint64_t* alpha, const uint64_t* z, const uint64_t* oods) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && x0 && outs && alpha && z && oods, "spg_stage_check_oods: null");
  const int C = SPG_AIR_COLS;
  AirPublic pub;
  for (int l = 0; l < SPG_AIR_LANES; l++) { pub.x0[l] = spg_host_from_u64(x0 + 4 * l); pub.outs[l] = spg_host_from_u64(outs + 4 * l); }
  Fp apows[SPG_AIR_LANES * SPG_AIR_NCONSTR], o[SPG_N_OODS];
  const Fp a = spg_host_from_u64(alpha), zz = spg_host_from_u64(z);
  apows[0] = fp_one();
  for (int k = 1; k < SPG_AIR_LANES * SPG_AIR_NCONSTR; k++) apows[k] = fp_mul(apows[k - 1], a);
  for (int k = 0; k < SPG_N_OODS; k++) o[k] = spg_host_from_u64(oods + 4 * k);
  const Fp lhs = spg_air_composition_at_host(log_n, chain_log, pub, apows, zz, o, o + C, ctx->h_const_points);
  Fp rhs = fp_zero(), zp = fp_one();
  for (int m = 0; m < 4; m++) { rhs = fp_add(rhs, fp_mul(zp, o[2 * C + m])); zp = fp_mul(zp, zz); }
  if (!fp_eq(lhs, rhs)) { ctx->err = "trace does not satisfy the AIR (composition mismatch at the out-of-domain point)"; return SPG_E_PROOF; }
  return SPG_OK;
}

// Last FRI layer: vals = [8][2^log_rows_last][4] raw device representation (Montgomery) gathered from all ranks,
// coset-major.  Writes the 2^log_rows_last proof coefficients (32 bytes each, proof serialisation) to coeffs_out;
// SPG_E_PROOF when the layer is not of low degree.
extern "C" int spg_stage_last_layer(spg_ctx* ctx, const uint64_t* vals, unsigned log_rows_last, int n_folds, uint8_t* coeffs_out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && vals && coeffs_out && log_rows_last <= 6 && n_folds >= 0 && n_folds < 16, "spg_stage_last_layer");
  const size_t n_last = (size_t)1 << log_rows_last;
  std::vector<Fp> v(8 * n_last), coeffs;
  memcpy(v.data(), vals, v.size() * sizeof(Fp));
  if (!spg_fri_last_layer_host(v, log_rows_last, n_folds, coeffs)) {
    ctx->err = "trace does not satisfy the AIR (FRI last layer is not of low degree)";
    return SPG_E_PROOF;
  }
  for (size_t k = 0; k < n_last; k++)
    for (int w = 0; w < 8; w++) {
      const uint32_t x = coeffs[k].v[7 - w];
      coeffs_out[32 * k + 4 * w] = (uint8_t)(x >> 24); coeffs_out[32 * k + 4 * w + 1] = (uint8_t)(x >> 16);
      coeffs_out[32 * k + 4 * w + 2] = (uint8_t)(x >> 8); coeffs_out[32 * k + 4 * w + 3] = (uint8_t)x;
    }
  return SPG_OK;
}

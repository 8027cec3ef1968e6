// Shared declarations of the STARK prover stages (air.cu, fri.cu, merkle.cu, prove.cu).
// Protocol: DESIGN.md "Protocol"; CPU restatement: oracle/stark.py.  No reference symbol exists for any of
// these stages.
#pragma once
#include "common.h"
#include "ec.cuh"
#include "ntt.cuh"
#include "air_point.cuh"   // SPG_AIR_LANES, SPG_AIR_NCONSTR

#define SPG_AIR_COLS 25
#define SPG_LOG_BLOWUP 3
#define SPG_BLOWUP 8
#define SPG_FRI_LAST_MAX 64
#define SPG_N_OODS 54

// omega_{2^26}^E through the two-level table (device)
__device__ __forceinline__ Fp spg_uni_pow(const Fp* __restrict__ uniA, const Fp* __restrict__ uniB, unsigned long long E) {
  E &= (1ull << SPG_UNI_LOG) - 1;
  const unsigned hi = (unsigned)(E >> SPG_UNI_HALF), lo = (unsigned)(E & ((1u << SPG_UNI_HALF) - 1));
  if (lo == 0) return uniA[hi];
  if (hi == 0) return uniB[lo];
  return fp_mul(uniA[hi], uniB[lo]);
}

// ---- merkle.cu
int spg_merkle_build_device(spg_ctx* ctx, const Fp* table, int ncols, size_t rows, uint32_t* tree, int n_cosets = 8);
int spg_merkle_open_device(spg_ctx* ctx, const Fp* table, int ncols, size_t rows, const uint32_t* tree,
                           const uint32_t* d_idx, int count, uint32_t* d_leaves, uint32_t* d_paths, int n_cosets = 8);

// ---- fri.cu
// out[a][jj][i] = 1 / (x - A[a]),  x = g * w_{8N}^(j0 + jstep*jj + 8 i),  a < n_a, jj < nj, i < N  (Montgomery)
int spg_inv_x_minus_device(spg_ctx* ctx, unsigned log_n, int j0, int jstep, int nj, const Fp* d_A, int n_a, Fp* out);
// DEEP quotient stage over n_cosets consecutive cosets of the LDE domain starting at first_coset (tables hold exactly those
// cosets): out[j][i].  z, gamma, oods[54] Montgomery (host); inv_scratch: device, 2 * n_cosets * N felts (3 * with t_coef);
// d_small: device, >= 64.  t_coef / a_coef (optional): combine the coefficient columns before extending (fri.cu).
int spg_deep_stage_device(spg_ctx* ctx, unsigned log_n, const Fp* t_lde, const Fp* h_lde, int first_coset, int n_cosets,
                          const Fp& z, const Fp& gamma, const Fp* oods, Fp* inv_scratch, Fp* d_small, Fp* out,
                          const Fp* t_coef = nullptr, Fp* a_coef = nullptr);
// one FRI fold by 8 of n_cosets consecutive cosets starting at first_coset: in [n_cosets][rows] -> out
// [n_cosets][rows/8]; x = g_l * w_{8 rows}^(j + 8 i)
int spg_fri_fold8_device(spg_ctx* ctx, const Fp* in, unsigned log_rows, const Fp& beta_over_g /*beta / g_l, Mont*/, Fp* out,
                         int first_coset = 0, int n_cosets = 8);
// evaluate n_items polynomials given as scaled, bit-reversed coefficient columns: item k uses the device
// column h_cols[k] and the point h_pts[h_pt_idx[k]] (Montgomery):  h_out[k] = sum_pos col[pos] * w^bitrev(pos)
// (synchronises the stream; h_* are host arrays)
int spg_poly_eval_device(spg_ctx* ctx, unsigned log_n, const Fp* const* h_cols, const int* h_pt_idx, int n_items,
                         const Fp* h_pts, int n_pts, Fp* h_out);

// last FRI layer on the host: see fri.cu
bool spg_fri_last_layer_host(const std::vector<Fp>& vals, unsigned log_rows_last, int n_folds, std::vector<Fp>& coeffs);

// ---- air.cu
struct AirPublic {
  Fp x0[SPG_AIR_LANES], outs[SPG_AIR_LANES];   // Montgomery
};
// composition polynomial on the even cosets j = 2 jj, jj in [jj0, jj0 + n_even): cp[jj - jj0][i].  t_lde holds
// consecutive cosets starting at first_coset.
int spg_air_eval_device(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const Fp* t_lde, const AirPublic& pub,
                        const Fp* h_alpha_pows /*[65], host, Mont*/, Fp* cp, int first_coset = 0, int jj0 = 0,
                        int n_even = 4);
// split cp (even cosets jj0 .. jj0 + n_even) into the 4 chunk columns evaluated on g^4 <w_N>, natural order:
// hev[m][jj + 4 i'] (only the positions of those cosets are written)
int spg_cp_split_device(spg_ctx* ctx, unsigned log_n, const Fp* cp, Fp* hev, int jj0 = 0, int n_even = 4);
// host-side evaluation of the composition at an out-of-domain point (prover self-check)
Fp spg_air_composition_at_host(unsigned log_n, unsigned chain_log, const AirPublic& pub, const Fp* alpha_pows,
                               const Fp& z, const Fp* tz, const Fp* tzw, const std::vector<Fp>& const_points);
// witness generation for the Pedersen hash-chain AIR (air.cu)
int spg_pedersen_trace_device(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const Fp* d_x0_canon /*[5]*/,
                              const Fp* d_ys_canon /*[5][N/512]*/, Fp* trace /*[25][N] canonical*/, uint8_t* d_status);

// ---- air_ecdsa.cu: the ECDSA-builtin AIR (same protocol slots as the Pedersen hash-chain AIR).  Public input: the message
// hashes and the keys' x of all N/256 signatures, as two polynomials over the block-start rows.
int spg_eair_public_device(spg_ctx* ctx, unsigned log_n, const Fp* d_pub /*[2][N/256] canonical*/, Fp* coef_small /*[2][N/256]*/,
                           Fp* pub_lde /*[n_even][2][N]*/, int jj0 = 0, int n_even = 4);
int spg_eair_public_at_host(spg_ctx* ctx, unsigned log_n, const Fp* d_coef_small, const Fp& z, Fp* out /*[2]*/);
int spg_eair_eval_device(spg_ctx* ctx, unsigned log_n, const Fp* t_lde, const Fp* pub_lde, const Fp* h_alpha_pows /*[52]*/, Fp* cp,
                         int first_coset = 0, int jj0 = 0, int n_even = 4);
Fp spg_eair_composition_at_host(spg_ctx* ctx, unsigned log_n, const Fp* pub_z /*[2]*/, const Fp* alpha_pows, const Fp& z,
                                const Fp* tz, const Fp* tzw);
int spg_eair_trace_device(spg_ctx* ctx, unsigned log_n, const Fp* msg, const Fp* rr, const Fp* ww, const Fp* kx, const Fp* ky,
                          Fp* trace, uint32_t* d_status);

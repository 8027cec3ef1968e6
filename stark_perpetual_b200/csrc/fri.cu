// DEEP quotient, batched inversion over the evaluation domain, FRI fold-by-8 and out-of-domain
// polynomial evaluation (protocol in DESIGN.md, CPU restatement in
// oracle/stark.py -- no reference symbol exists for these stages).
#include "stark_kernels.cuh"

// ------------------------------------------------------------------ 1 / (x - A) over cosets
// Montgomery batch inversion: each thread owns `ch` points strided by the block size (coalesced) and handles up to
// three shifts A[a] at once -- the domain point x is stepped once per point for all of them, the running prefix
// products live in the output buffers themselves, and ONE Fermat inversion (of the product of the three running
// products) serves all chains of the thread.  Per output: (1 + 2/3) multiplications forward, (2 + 2/3) backward,
// plus 271 / (3 ch) for the inversion.
#ifndef INV_THREADS
#define INV_THREADS 128     // 160 registers per thread: three CTAs of 128 threads per SM (12 warps) instead of one of 256
#endif
template <int NA>
__global__ void __launch_bounds__(INV_THREADS) k_inv_x_minus(unsigned log_n, int j0, int jstep, int nj, const Fp* __restrict__ A,
                                                     Fp* __restrict__ out, int ch, Fp g, const Fp* __restrict__ uniA,
                                                     const Fp* __restrict__ uniB) {
  const size_t n = (size_t)1 << log_n;
  const int jj = blockIdx.y, a0 = blockIdx.z * NA, j = j0 + jstep * jj;
  const size_t i0 = (size_t)blockIdx.x * INV_THREADS * ch + threadIdx.x;
  const int sh = SPG_UNI_LOG - (int)log_n - SPG_LOG_BLOWUP;
  Fp x = fp_mul(g, spg_uni_pow(uniA, uniB, ((unsigned long long)j + 8ull * i0) << sh));
  const unsigned long long step_e = ((unsigned long long)INV_THREADS << (SPG_UNI_LOG - log_n));
  const Fp step = spg_uni_pow(uniA, uniB, step_e), stepinv = spg_uni_pow(uniA, uniB, 0ull - step_e);
  Fp av[NA], acc[NA];
  Fp* o[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) {
    av[a] = A[a0 + a];
    acc[a] = fp_one();
    o[a] = out + (((size_t)(a0 + a) * nj + jj) << log_n);
  }
  for (int k = 0; k < ch; k++) {
    const size_t i = i0 + (size_t)INV_THREADS * k;
    if (i < n) {
#pragma unroll
      for (int a = 0; a < NA; a++) { o[a][i] = acc[a]; acc[a] = fp_mul(acc[a], fp_sub(x, av[a])); }
    }
    x = fp_mul(x, step);
  }
  // one inversion for the NA chains: inv[a] = (prod of all) ^-1 * prod of the others
  Fp inv[NA];
  {
    Fp all = acc[0];
#pragma unroll
    for (int a = 1; a < NA; a++) all = fp_mul(all, acc[a]);
    const Fp ia = fp_inv_chain(all);
    if (NA == 1) inv[0] = ia;
    else if (NA == 2) { inv[0] = fp_mul(ia, acc[1]); inv[1] = fp_mul(ia, acc[0]); }
    else {
      inv[0] = fp_mul(ia, fp_mul(acc[1], acc[2 % NA]));
      inv[1] = fp_mul(ia, fp_mul(acc[0], acc[2 % NA]));
      inv[2 % NA] = fp_mul(ia, fp_mul(acc[0], acc[1]));
    }
  }
  for (int k = ch - 1; k >= 0; k--) {
    x = fp_mul(x, stepinv);
    const size_t i = i0 + (size_t)INV_THREADS * k;
    if (i < n) {
#pragma unroll
      for (int a = 0; a < NA; a++) {
        const Fp pre = o[a][i];
        o[a][i] = fp_reduce(fp_mul(inv[a], pre));
        inv[a] = fp_mul(inv[a], fp_sub(x, av[a]));
      }
    }
  }
}

static Fp host_gen() { uint64_t three[4] = {3, 0, 0, 0}; return spg_host_from_u64(three); }

int spg_inv_x_minus_device(spg_ctx* ctx, unsigned log_n, int j0, int jstep, int nj, const Fp* d_A, int n_a, Fp* out) {
  const size_t n = (size_t)1 << log_n;
  // points per thread: as many as keep the grid at a whole number of waves of sm_count * 3 resident CTAs (the single
  // Fermat inversion of a thread is amortised over 3 * ch outputs), between 16 and 96
  const size_t per_cta_min = INV_THREADS;
  int ch = 1;
  if (n > per_cta_min * 16) {
    const int groups = (n_a % 3 == 0) ? n_a / 3 : (n_a % 2 == 0) ? n_a / 2 : n_a;
    const double slots = (double)ctx->sm_count * 3;
    for (int waves = 1; waves <= 64; waves++) {
      const double tiles_f = waves * slots / ((double)nj * groups);
      const size_t tiles_w = tiles_f < 1.0 ? 1 : (size_t)tiles_f;
      const size_t c = (n + tiles_w * INV_THREADS - 1) / (tiles_w * INV_THREADS);
      ch = (int)c;
      if (c <= 96) break;
    }
    if (ch < 16) ch = 16;
  }
  const unsigned tiles = (unsigned)((n + (size_t)INV_THREADS * ch - 1) / ((size_t)INV_THREADS * ch));
  if (n_a % 3 == 0) {
    dim3 grid(tiles, (unsigned)nj, (unsigned)(n_a / 3));
    k_inv_x_minus<3><<<grid, INV_THREADS, 0, ctx->stream>>>(log_n, j0, jstep, nj, d_A, out, ch, host_gen(), ctx->uniA, ctx->uniB);
  } else if (n_a % 2 == 0) {
    dim3 grid(tiles, (unsigned)nj, (unsigned)(n_a / 2));
    k_inv_x_minus<2><<<grid, INV_THREADS, 0, ctx->stream>>>(log_n, j0, jstep, nj, d_A, out, ch, host_gen(), ctx->uniA, ctx->uniB);
  } else {
    dim3 grid(tiles, (unsigned)nj, (unsigned)n_a);
    k_inv_x_minus<1><<<grid, INV_THREADS, 0, ctx->stream>>>(log_n, j0, jstep, nj, d_A, out, ch, host_gen(), ctx->uniA, ctx->uniB);
  }
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ DEEP quotient
// Q(x) = (A(x) - K0) / (x - z) + (gamma^25 A(x) - K1) / (x - z w) + (C(x) - K2) / (x - z^4),  A = sum_c gamma^c T_c,
// C = sum_m gamma^(50+m) H_m.  Only TWO inverse tables are needed: the points of a coset are x_i = x_0 w^i, so
// x_i - z w = w (x_{i-1} - z) and 1 / (x_i - z w) = w^-1 / (x_{i-1} - z) -- the previous row of the first table; the
// factor w^-1 is folded into the constants (G1 = gamma^25 / w, K1' = K1 / w).  inv2: [2][n_cosets][N] = 1/(x - z), 1/(x - z^4).
__global__ void __launch_bounds__(256) k_deep(unsigned log_n, const Fp* __restrict__ t_lde, const Fp* __restrict__ h_lde,
                                              const Fp* __restrict__ inv2, const Fp* __restrict__ gamma,
                                              const Fp* __restrict__ K, Fp* __restrict__ out, int n_cosets) {
  const size_t n = (size_t)1 << log_n;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_cosets * n) return;
  const size_t j = idx >> log_n, i = idx & (n - 1);
  // Lazy accumulation (fp.cuh): products are below 2p, twelve of them are summed before a partial reduction.
  Fp a = fp_zero(), c = fp_zero();
  const Fp* tp = t_lde + (j * SPG_AIR_COLS << log_n) + i;
#pragma unroll 5
  for (int col = 0; col < SPG_AIR_COLS; col++) {
    a = fp_add_raw(a, fp_mul_lazy(tp[(size_t)col << log_n], gamma[col]));
    if (col % 12 == 11) a = fp_partial(a);
  }
  a = fp_partial(a);
  Fp b = fp_mul(a, gamma[SPG_AIR_COLS]);             // gamma^25 / w
  const Fp* hp = h_lde + (j * 4 << log_n) + i;
#pragma unroll
  for (int m = 0; m < 4; m++) c = fp_add_raw(c, fp_mul_lazy(hp[(size_t)m << log_n], gamma[2 * SPG_AIR_COLS + m]));
  c = fp_partial(c);
  a = fp_sub(a, K[0]); b = fp_sub(b, K[1]); c = fp_sub(c, K[2]);
  const Fp i1 = inv2[idx], i2 = inv2[(j << log_n) + ((i + n - 1) & (n - 1))], i3 = inv2[idx + (size_t)n_cosets * n];
  Fp q = fp_add(fp_add(fp_mul(a, i1), fp_mul(b, i2)), fp_mul(c, i3));
  out[idx] = fp_reduce(q);
}

// A = sum_c gamma^c T_c is a polynomial of degree < N like its summands: when all 8 cosets are on this GPU it is cheaper to
// combine the 25 COEFFICIENT columns once (25 N products) and extend the one combined column (16 single-column transform
// launches) than to combine the 25 extended columns at every one of the 8 N points (200 N products, 6.7 GB of reads).
__global__ void __launch_bounds__(256) k_combine_cols(unsigned log_n, const Fp* __restrict__ cols, const Fp* __restrict__ gamma,
                                                      Fp* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((size_t)1 << log_n)) return;
  Fp a = fp_zero();
#pragma unroll 5
  for (int col = 0; col < SPG_AIR_COLS; col++) {
    a = fp_add_raw(a, fp_mul_lazy(cols[((size_t)col << log_n) + i], gamma[col]));
    if (col % 12 == 11) a = fp_partial(a);
  }
  out[i] = fp_reduce(fp_partial(a));
}
// the quotient with A already extended: a_lde[j][i]
__global__ void __launch_bounds__(256) k_deep_combined(unsigned log_n, const Fp* __restrict__ a_lde, const Fp* __restrict__ h_lde,
                                                       const Fp* __restrict__ inv2, const Fp* __restrict__ gamma,
                                                       const Fp* __restrict__ K, Fp* __restrict__ out, int n_cosets) {
  const size_t n = (size_t)1 << log_n;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_cosets * n) return;
  const size_t j = idx >> log_n, i = idx & (n - 1);
  Fp a = fp_reduce(a_lde[idx]), c = fp_zero();
  Fp b = fp_mul(a, gamma[SPG_AIR_COLS]);             // gamma^25 / w
  const Fp* hp = h_lde + (j * 4 << log_n) + i;
#pragma unroll
  for (int m = 0; m < 4; m++) c = fp_add_raw(c, fp_mul_lazy(hp[(size_t)m << log_n], gamma[2 * SPG_AIR_COLS + m]));
  c = fp_partial(c);
  a = fp_sub(a, K[0]); b = fp_sub(b, K[1]); c = fp_sub(c, K[2]);
  const Fp i1 = inv2[idx], i2 = inv2[(j << log_n) + ((i + n - 1) & (n - 1))], i3 = inv2[idx + (size_t)n_cosets * n];
  out[idx] = fp_reduce(fp_add(fp_add(fp_mul(a, i1), fp_mul(b, i2)), fp_mul(c, i3)));
}

// The whole DEEP stage on n_cosets consecutive cosets starting at first_coset: constants on the host, two batched
// inverse tables, the quotient kernel.  z, gamma, oods[54]: Montgomery; inv_scratch: 2 n_cosets N felts (3 n_cosets N when
// t_coef is given); d_small: >= 64.  t_coef (optional): the trace's scaled coefficient columns [25][N] as spg_lde_coeffs
// leaves them, a_coef: N felts of scratch -- with both and all 8 cosets here, A is combined before the extension.
int spg_deep_stage_device(spg_ctx* ctx, unsigned log_n, const Fp* t_lde, const Fp* h_lde, int first_coset, int n_cosets,
                          const Fp& z, const Fp& gamma, const Fp* oods, Fp* inv_scratch, Fp* d_small, Fp* out,
                          const Fp* t_coef, Fp* a_coef) {
  const int C = SPG_AIR_COLS;
  const Fp wn = spg_host_root_of_unity((int)log_n), wn_inv = fp_inv(wn), z4 = fp_sqr(fp_sqr(z));
  Fp gp[SPG_N_OODS + 5];
  gp[0] = fp_one();
  for (int k = 1; k < SPG_N_OODS; k++) gp[k] = fp_mul(gp[k - 1], gamma);
  Fp K[3] = {fp_zero(), fp_zero(), fp_zero()};
  for (int c = 0; c < C; c++) { K[0] = fp_add(K[0], fp_mul(gp[c], oods[c])); K[1] = fp_add(K[1], fp_mul(gp[C + c], oods[C + c])); }
  for (int m = 0; m < 4; m++) K[2] = fp_add(K[2], fp_mul(gp[2 * C + m], oods[2 * C + m]));
  gp[C] = fp_mul(gp[C], wn_inv);                      // the kernel uses gamma[25] only as the factor of the second quotient
  gp[SPG_N_OODS] = K[0]; gp[SPG_N_OODS + 1] = fp_mul(K[1], wn_inv); gp[SPG_N_OODS + 2] = K[2];
  gp[SPG_N_OODS + 3] = z; gp[SPG_N_OODS + 4] = z4;
  SPG_CUDA(cudaMemcpyAsync(d_small, gp, sizeof(gp), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));       // gp is a stack object
  int rc = spg_inv_x_minus_device(ctx, log_n, first_coset, 1, n_cosets, d_small + SPG_N_OODS + 3, 2, inv_scratch);
  if (rc) return rc;
  const size_t total = (size_t)n_cosets << log_n;
  if (t_coef && a_coef && n_cosets == SPG_BLOWUP && !ctx->deep_pointwise) {
    const size_t n = (size_t)1 << log_n;
    Fp* a_lde = inv_scratch + 2 * total;
    k_combine_cols<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(log_n, t_coef, d_small, a_coef);
    SPG_LAUNCH_CHECK();
    if ((rc = spg_lde_cosets_device(ctx, a_coef, log_n, 1, SPG_LOG_BLOWUP, (size_t)first_coset, (size_t)n_cosets, a_lde))) return rc;
    k_deep_combined<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(log_n, a_lde, h_lde, inv_scratch, d_small,
                                                                             d_small + SPG_N_OODS, out, n_cosets);
  } else {
    k_deep<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(log_n, t_lde, h_lde, inv_scratch, d_small, d_small + SPG_N_OODS, out,
                                                                     n_cosets);
  }
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ FRI fold by 8
struct FoldConsts {
  Fp zi1, zi2, zi3;   // w_8^-1, w_8^-2, w_8^-3
  Fp inv8;
  Fp bg;              // beta / g_l
};

__global__ void __launch_bounds__(256) k_fri_fold8(const Fp* __restrict__ in, unsigned log_rows, FoldConsts K,
                                                   Fp* __restrict__ out, const Fp* __restrict__ uniA,
                                                   const Fp* __restrict__ uniB, int first_coset, int n_cosets) {
  const unsigned log_g = log_rows - 3;
  const size_t g = (size_t)1 << log_g;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_cosets * g) return;
  const size_t jl = idx >> log_g, ip = idx & (g - 1), j = jl + first_coset;
  const Fp* p = in + (jl << log_rows) + ip;
  Fp v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = p[(size_t)k << log_g];
  // size-8 inverse DFT, s[m] = sum_k w_8^(-m k) v[k]
  Fp a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { a[k] = fp_add(v[k], v[k + 4]); b[k] = fp_sub(v[k], v[k + 4]); }
  b[1] = fp_mul(b[1], K.zi1); b[2] = fp_mul(b[2], K.zi2); b[3] = fp_mul(b[3], K.zi3);
  Fp s[8];
  {
    // even m = 2 m': size-4 inverse DFT of a;  odd m = 2 m' + 1: of b
    Fp t0 = fp_add(a[0], a[2]), t1 = fp_sub(a[0], a[2]), t2 = fp_add(a[1], a[3]);
    Fp t3 = fp_mul(fp_sub(a[1], a[3]), K.zi2);           // w_4^-1 = w_8^-2
    s[0] = fp_add(t0, t2); s[4] = fp_sub(t0, t2); s[2] = fp_add(t1, t3); s[6] = fp_sub(t1, t3);
    t0 = fp_add(b[0], b[2]); t1 = fp_sub(b[0], b[2]); t2 = fp_add(b[1], b[3]);
    t3 = fp_mul(fp_sub(b[1], b[3]), K.zi2);
    s[1] = fp_add(t0, t2); s[5] = fp_sub(t0, t2); s[3] = fp_add(t1, t3); s[7] = fp_sub(t1, t3);
  }
  // t = beta / x,  x = g_l * w_{8 rows}^(j + 8 ip)
  const int sh = SPG_UNI_LOG - (int)log_rows - SPG_LOG_BLOWUP;
  const unsigned long long e = ((unsigned long long)j + 8ull * ip) << sh;
  const Fp t = fp_mul(K.bg, spg_uni_pow(uniA, uniB, 0ull - e));
  Fp acc = s[7];
#pragma unroll
  for (int m = 6; m >= 0; m--) acc = fp_add(fp_mul(acc, t), s[m]);
  out[idx] = fp_reduce(fp_mul(acc, K.inv8));
}

int spg_fri_fold8_device(spg_ctx* ctx, const Fp* in, unsigned log_rows, const Fp& beta_over_g, Fp* out, int first_coset,
                         int n_cosets) {
  SPG_ARG(log_rows >= 3 && log_rows + SPG_LOG_BLOWUP <= SPG_UNI_LOG, "fri fold: size");
  FoldConsts K;
  const Fp z = spg_host_root_of_unity(3), zi = fp_inv(z);
  K.zi1 = zi; K.zi2 = fp_mul(zi, zi); K.zi3 = fp_mul(K.zi2, zi);
  uint64_t eight[4] = {8, 0, 0, 0};
  K.inv8 = fp_inv(spg_host_from_u64(eight));
  K.bg = beta_over_g;
  const size_t total = ((size_t)1 << (log_rows - 3)) * n_cosets;
  k_fri_fold8<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(in, log_rows, K, out, ctx->uniA, ctx->uniB,
                                                                         first_coset, n_cosets);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ out-of-domain evaluation
#define SPG_EVAL_T 10
__global__ void k_eval_tables(const Fp* __restrict__ pts, int n_pts, unsigned log_n, unsigned t, Fp* __restrict__ A,
                              Fp* __restrict__ B) {
  const size_t na = (size_t)1 << t, nb = (size_t)1 << (log_n - t);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (na + nb) * n_pts) return;
  const int p = (int)(idx / (na + nb));
  const size_t r = idx - (size_t)p * (na + nb);
  const Fp w = pts[p];
  if (r < na) {
    const uint64_t e = (uint64_t)spg_bitrev((unsigned)r, (int)t) << (log_n - t);
    A[(size_t)p * na + r] = e ? fp_pow_u64(w, e) : fp_one();
  } else {
    const size_t hi = r - na;
    const uint64_t e = spg_bitrev((unsigned)hi, (int)(log_n - t));
    B[(size_t)p * nb + hi] = e ? fp_pow_u64(w, e) : fp_one();
  }
}

__device__ __forceinline__ Fp block_sum(Fp v, Fp* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (tid < s) sh[tid] = fp_add(sh[tid], sh[tid + s]);
    __syncthreads();
  }
  return sh[0];
}

__global__ void __launch_bounds__(256) k_poly_eval_partial(const Fp* const* __restrict__ cols, const int* __restrict__ pt_idx,
                                                           unsigned log_n, unsigned t, const Fp* __restrict__ A,
                                                           const Fp* __restrict__ B, Fp* __restrict__ partial) {
  __shared__ Fp sh[256];
  const size_t hi = blockIdx.x, n_hi = gridDim.x;
  const int item = blockIdx.y, p = pt_idx[item];
  const size_t L = (size_t)1 << t;
  const Fp* col = cols[item] + (hi << t);
  const Fp* a = A + (size_t)p * L;
  Fp acc = fp_zero();
  for (size_t lo = threadIdx.x; lo < L; lo += 256) acc = fp_add(acc, fp_mul(col[lo], a[lo]));
  const Fp s = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[(size_t)item * n_hi + hi] = fp_mul(s, B[(size_t)p * n_hi + hi]);
}

__global__ void __launch_bounds__(256) k_poly_eval_final(const Fp* __restrict__ partial, size_t n_hi, Fp* __restrict__ out) {
  __shared__ Fp sh[256];
  const int item = blockIdx.x;
  Fp acc = fp_zero();
  for (size_t h = threadIdx.x; h < n_hi; h += 256) acc = fp_add(acc, partial[(size_t)item * n_hi + h]);
  const Fp s = block_sum(acc, sh);
  if (threadIdx.x == 0) out[item] = fp_reduce(s);
}

int spg_poly_eval_device(spg_ctx* ctx, unsigned log_n, const Fp* const* h_cols, const int* h_pt_idx, int n_items,
                         const Fp* h_pts, int n_pts, Fp* h_out) {
  const unsigned t = log_n < SPG_EVAL_T ? log_n : SPG_EVAL_T;
  const size_t na = (size_t)1 << t, nb = (size_t)1 << (log_n - t);
  DevBuf dpts, dA, dB, dcols, dpidx, dpart, dout;
  SPG_CUDA(dpts.alloc(ctx, n_pts * sizeof(Fp))); SPG_CUDA(dA.alloc(ctx, n_pts * na * sizeof(Fp))); SPG_CUDA(dB.alloc(ctx, n_pts * nb * sizeof(Fp)));
  SPG_CUDA(dcols.alloc(ctx, n_items * sizeof(Fp*))); SPG_CUDA(dpidx.alloc(ctx, n_items * sizeof(int)));
  SPG_CUDA(dpart.alloc(ctx, n_items * nb * sizeof(Fp))); SPG_CUDA(dout.alloc(ctx, n_items * sizeof(Fp)));
  SPG_CUDA(cudaMemcpyAsync(dpts.p, h_pts, n_pts * sizeof(Fp), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(dcols.p, h_cols, n_items * sizeof(Fp*), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(dpidx.p, h_pt_idx, n_items * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  const size_t nt = (na + nb) * n_pts;
  k_eval_tables<<<(unsigned)((nt + 127) / 128), 128, 0, ctx->stream>>>(dpts.as<Fp>(), n_pts, log_n, t, dA.as<Fp>(), dB.as<Fp>());
  SPG_LAUNCH_CHECK();
  dim3 grid((unsigned)nb, (unsigned)n_items);
  k_poly_eval_partial<<<grid, 256, 0, ctx->stream>>>((const Fp* const*)dcols.p, dpidx.as<int>(), log_n, t, dA.as<Fp>(),
                                                     dB.as<Fp>(), dpart.as<Fp>());
  SPG_LAUNCH_CHECK();
  k_poly_eval_final<<<n_items, 256, 0, ctx->stream>>>(dpart.as<Fp>(), nb, dout.as<Fp>());
  SPG_LAUNCH_CHECK();
  SPG_CUDA(spg_d2h_sync(ctx, h_out, dout.p, n_items * sizeof(Fp), ctx->stream));
  return SPG_OK;
}

// ------------------------------------------------------------------ last FRI layer (host)
// in-place radix-2 inverse NTT on the host (natural order in and out), n = 2^log_n
static void host_intt(std::vector<Fp>& a, int log_n) {
  const size_t n = a.size();
  for (size_t i = 0; i < n; i++) { size_t r = spg_bitrev((unsigned)i, log_n); if (r > i) std::swap(a[i], a[r]); }
  const Fp w = fp_inv(spg_host_root_of_unity(log_n));
  for (size_t h = 1; h < n; h *= 2) {
    const Fp wh = fp_pow_u64(w, n / (2 * h));
    for (size_t b = 0; b < n; b += 2 * h) {
      Fp t = fp_one();
      for (size_t k = 0; k < h; k++) {
        const Fp u = a[b + k], v = fp_mul(a[b + k + h], t);
        a[b + k] = fp_add(u, v); a[b + k + h] = fp_sub(u, v);
        t = fp_mul(t, wh);
      }
    }
  }
  uint64_t nn[4] = {(uint64_t)n, 0, 0, 0};
  const Fp ninv = fp_inv(spg_host_from_u64(nn));
  for (auto& x : a) x = fp_mul(x, ninv);
}

// vals: the last layer [8 cosets][n_last] (Montgomery) on the domain g_l * <w_{8 n_last}>, g_l = 3^(8^n_folds).
// Interpolates it; returns false unless the upper 7/8 of the coefficients vanish; coeffs receives the n_last
// low ones (what the proof carries).
bool spg_fri_last_layer_host(const std::vector<Fp>& vals, unsigned log_rows_last, int n_folds, std::vector<Fp>& coeffs) {
  const size_t n_last = (size_t)1 << log_rows_last;
  std::vector<Fp> flat(8 * n_last);
  for (size_t j = 0; j < 8; j++) for (size_t i = 0; i < n_last; i++) flat[j + 8 * i] = vals[j * n_last + i];
  host_intt(flat, (int)log_rows_last + 3);
  uint64_t three[4] = {3, 0, 0, 0};
  Fp g_l = spg_host_from_u64(three);
  for (int k = 0; k < 3 * n_folds; k++) g_l = fp_sqr(g_l);
  const Fp gli = fp_inv(g_l);
  Fp s = fp_one();
  for (size_t k = 0; k < flat.size(); k++) { flat[k] = fp_mul(flat[k], s); s = fp_mul(s, gli); }
  for (size_t k = n_last; k < flat.size(); k++)
    if (!fp_is_zero(flat[k])) return false;
  coeffs.assign(flat.begin(), flat.begin() + n_last);
  return true;
}

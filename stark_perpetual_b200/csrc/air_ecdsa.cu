// The ECDSA-builtin AIR (second AIR of the prover): witness generation, constraint (composition) evaluation over the
// LDE cosets, host evaluation at the out-of-domain point.  SURVEY.md section 8 row p3; DESIGN.md section 5b.
// The reference has no AIR; what it pins is the computation every 256-row block encodes -- one `verify` call:
// signature.py:243-260 (three mimic_ec_mult_air loops, :176-190, then r == x), math_utils.py:59-88 (ec_add, ec_double).
// CPU twin (trace, constraints, verifier): oracle/stark_ecdsa.py.
#include "stark_kernels.cuh"
#include "ecdsa_air_point.cuh"
#include "ecdsa_air_witness.cuh"
#include "curve_params.inc"

static Fp eh_small(uint64_t v) { uint64_t w[4] = {v, 0, 0, 0}; return spg_host_from_u64(w); }

// ------------------------------------------------------------------ cached tables
// izt[g][jj][i mod 256], g < 6: the block-periodic inverse zerofiers on cosets 0, 2, 4, 6 (host arithmetic, one batch
// inversion: 4096 values once per trace size)
static int ensure_eair_tables(spg_ctx* ctx, unsigned log_n) {
  if (ctx->eair_log_n == (int)log_n) return SPG_OK;
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->eair_izt); cudaFree(ctx->eair_plde);
  ctx->eair_izt = ctx->eair_plde = nullptr;
  ctx->eair_log_n = -1;
  const size_t n = (size_t)1 << log_n;
  const int B = SPG_EAIR_BLOCK;
  const Fp one = fp_one(), g = eh_small(3);
  const Fp g256 = fp_pow_u64(g, n >> 8), gn = fp_pow_u64(g, n);
  const Fp w256 = spg_host_root_of_unity(8), w2048 = spg_host_root_of_unity(11), w8 = spg_host_root_of_unity(3);
  Fp wt[5];                                          // w_256^251 .. w_256^255
  for (int k = 0; k < 5; k++) wt[k] = fp_pow_u64(w256, SPG_EAIR_BITS + k);
  std::vector<Fp> den(4 * B * 4), tail5(4 * B), elast(4 * B);
  for (int jj = 0; jj < 4; jj++) {
    Fp u = fp_mul(g256, fp_pow_u64(w2048, 2 * jj));  // x^(N/256) = g256 w_2048^(j + 8 i), and w_2048^8 = w_256
    for (int i = 0; i < B; i++) {
      Fp t4 = one;
      for (int k = 0; k < 4; k++) t4 = fp_mul(t4, fp_sub(u, wt[k]));
      const Fp el = fp_sub(u, wt[4]), t5 = fp_mul(t4, el);
      const size_t q = (size_t)jj * B + i;
      den[4 * q] = t4; den[4 * q + 1] = t5; den[4 * q + 2] = fp_sub(u, one); den[4 * q + 3] = el;
      tail5[q] = t5; elast[q] = el;
      u = fp_mul(u, w256);
    }
  }
  std::vector<Fp> pre(den.size());
  Fp run = one;
  for (size_t k = 0; k < den.size(); k++) { pre[k] = run; run = fp_mul(run, den[k]); }
  Fp inv = fp_inv(run);
  std::vector<Fp> idn(den.size());
  for (size_t k = den.size(); k-- > 0;) { idn[k] = fp_mul(inv, pre[k]); inv = fp_mul(inv, den[k]); }
  std::vector<Fp> izt(6 * 4 * B);
  for (int jj = 0; jj < 4; jj++) {
    const Fp iz_all = fp_inv(fp_sub(fp_mul(gn, fp_pow_u64(w8, 2 * jj)), one));      // 1 / (x^N - 1) on coset 2 jj
    for (int i = 0; i < B; i++) {
      const size_t q = (size_t)jj * B + i;
      izt[0 * 4 * B + q] = fp_reduce(fp_mul(tail5[q], iz_all));        // step: rows t <= 250
      izt[1 * 4 * B + q] = fp_reduce(idn[4 * q]);                      // hold: rows 251 .. 254
      izt[2 * 4 * B + q] = fp_reduce(idn[4 * q + 1]);                  // zero: rows 251 .. 255
      izt[3 * 4 * B + q] = fp_reduce(idn[4 * q + 2]);                  // first: t = 0
      izt[4 * 4 * B + q] = fp_reduce(idn[4 * q + 3]);                  // last: t = 255
      izt[5 * 4 * B + q] = fp_reduce(fp_mul(elast[q], iz_all));        // thold: every row but t = 255
    }
  }
  SPG_CUDA(cudaMalloc((void**)&ctx->eair_izt, izt.size() * sizeof(Fp)));
  SPG_CUDA(cudaMemcpy(ctx->eair_izt, izt.data(), izt.size() * sizeof(Fp), cudaMemcpyHostToDevice));
  // lane A's periodic point 2^t G (0 on the padding rows), extended to the cosets: [8][2][256]
  std::vector<Fp> pcols(2 * B, fp_zero());
  for (int t = 0; t < SPG_EAIR_BITS; t++) { pcols[t] = ctx->h_gen_doubles[2 * t]; pcols[B + t] = ctx->h_gen_doubles[2 * t + 1]; }
  DevBuf dp, dcoef;
  SPG_CUDA(dp.alloc(ctx, pcols.size() * sizeof(Fp))); SPG_CUDA(dcoef.alloc(ctx, pcols.size() * sizeof(Fp)));
  SPG_CUDA(cudaMemcpyAsync(dp.p, pcols.data(), pcols.size() * sizeof(Fp), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMalloc((void**)&ctx->eair_plde, 8 * 2 * B * sizeof(Fp)));
  uint64_t off[4];
  spg_host_to_u64(g256, off);
  int rc = spg_lde_device(ctx, dp.as<Fp>(), 8, 2, SPG_LOG_BLOWUP, off, ctx->eair_plde, dcoef.as<Fp>(), 0);
  if (rc) return rc;
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));      // pcols is a host object
  ctx->eair_log_n = (int)log_n;
  return SPG_OK;
}

// ------------------------------------------------------------------ composition evaluation
// loaders handed to ecdsa_air_point: cell k of the row, group g's inverse zerofier
struct EairCells {
  const Fp* __restrict__ p; unsigned log_n;
  __device__ __forceinline__ Fp operator[](int k) const { return p[(size_t)k << log_n]; }
};
struct EairZerofiers {
  const Fp* __restrict__ izt;
  __device__ __forceinline__ Fp operator[](int g) const { return izt[(size_t)g * 4 * SPG_EAIR_BLOCK]; }
};
__global__ void __launch_bounds__(128) k_air_eval_ecdsa(unsigned log_n, const Fp* __restrict__ t_lde,
                                                        const EcdsaAirConsts* __restrict__ K, const Fp* __restrict__ izt,
                                                        const Fp* __restrict__ plde, const Fp* __restrict__ pub_lde,
                                                        Fp* __restrict__ cp, int first_coset, int jj0, int n_even) {
  const size_t n = (size_t)1 << log_n;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_even * n) return;
  const size_t jj = (idx >> log_n) + jj0, i = idx & (n - 1), j = 2 * jj;
  const size_t in = (i + 1) & (n - 1);
  const int B = SPG_EAIR_BLOCK;
  const Fp* base = t_lde + ((j - first_coset) * SPG_EAIR_COLS << log_n);
  const EairCells c = {base + i, log_n}, nx = {base + in, log_n};
  const EairZerofiers iz = {izt + jj * B + (i & (B - 1))};
  const Fp gx = plde[(j * 2 + 0) * B + (i & (B - 1))], gy = plde[(j * 2 + 1) * B + (i & (B - 1))];
  const Fp* pl = pub_lde + (((jj - jj0) * 2) << log_n) + i;          // public columns on coset 2 jj: [jj - jj0][2][N]
  cp[idx] = ecdsa_air_point(c, nx, gx, gy, pl[0], pl[n], *K, iz);
}

static void eair_consts(spg_ctx* ctx, const Fp* alpha_pows, EcdsaAirConsts& K) {
  for (int k = 0; k < SPG_EAIR_NALPHA; k++) K.alpha[k] = alpha_pows[k];
  K.shift_x = ctx->h_const_points[0]; K.shift_y = ctx->h_const_points[1];
  K.minus_shift_y = fp_reduce(fp_neg(K.shift_y));
  K.beta = spg_host_from_u64(SPG_BETA);
}

// The public columns.  d_pub: [2][N/256] canonical (message hashes, keys' x), device.  coef_small [2][N/256] receives the
// interpolants' coefficients c_k g^k (Montgomery, bit-reversed: what spg_lde_coeffs writes); pub_lde [n_even][2][N] their
// values on the cosets 2 jj, jj in [jj0, jj0 + n_even).  A polynomial of degree < N/256 over the size-N domain is the same
// coefficient vector zero-padded, and in bit-reversed order coefficient k of the small transform sits at position 256 q of
// the big one (q its small position): one strided copy, then the ordinary coset transforms.
int spg_eair_public_device(spg_ctx* ctx, unsigned log_n, const Fp* d_pub, Fp* coef_small, Fp* pub_lde, int jj0, int n_even) {
  const size_t n = (size_t)1 << log_n, nb = n >> 8;
  int rc = spg_lde_coeffs_device(ctx, d_pub, log_n - 8, 2, nullptr, coef_small, /*mont=*/1);
  if (rc) return rc;
  DevBuf pad;
  SPG_CUDA(pad.alloc(ctx, 2 * n * sizeof(Fp)));
  SPG_CUDA(cudaMemsetAsync(pad.p, 0, 2 * n * sizeof(Fp), ctx->stream));
  for (int c = 0; c < 2; c++)
    SPG_CUDA(cudaMemcpy2DAsync(pad.as<Fp>() + c * n, SPG_EAIR_BLOCK * sizeof(Fp), coef_small + c * nb, sizeof(Fp), sizeof(Fp), nb,
                               cudaMemcpyDeviceToDevice, ctx->stream));
  for (int e = 0; e < n_even; e++)
    if ((rc = spg_lde_cosets_device(ctx, pad.as<Fp>(), log_n, 2, SPG_LOG_BLOWUP, 2 * (size_t)(jj0 + e), 1, pub_lde + ((size_t)e * 2 << log_n))))
      return rc;
  return SPG_OK;
}

// the two public polynomials at an out-of-domain point z (host Horner over the bit-reversed coefficients): out[0] = F_msg(z),
// out[1] = F_key(z)
int spg_eair_public_at_host(spg_ctx* ctx, unsigned log_n, const Fp* d_coef_small, const Fp& z, Fp* out) {
  const unsigned lb = log_n - 8;
  const size_t nb = (size_t)1 << lb;
  std::vector<Fp> cf(2 * nb);
  SPG_CUDA(cudaMemcpyAsync(cf.data(), d_coef_small, cf.size() * sizeof(Fp), cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  const Fp y = fp_mul(z, fp_inv(eh_small(3)));       // the coefficients carry g^k
  Fp pw = fp_one();
  out[0] = out[1] = fp_zero();
  for (size_t k = 0; k < nb; k++) {
    size_t q = 0;
    for (unsigned b = 0; b < lb; b++) q |= ((k >> b) & 1) << (lb - 1 - b);
    out[0] = fp_add(out[0], fp_mul(cf[q], pw));
    out[1] = fp_add(out[1], fp_mul(cf[nb + q], pw));
    pw = fp_mul(pw, y);
  }
  return SPG_OK;
}

int spg_eair_eval_device(spg_ctx* ctx, unsigned log_n, const Fp* t_lde, const Fp* pub_lde, const Fp* h_alpha_pows, Fp* cp,
                         int first_coset, int jj0, int n_even) {
  if (n_even <= 0) return SPG_OK;
  SPG_ARG(2 * jj0 >= first_coset && jj0 + n_even <= 4, "ecdsa air eval: coset range");
  SPG_ARG(log_n >= 9 && log_n + SPG_LOG_BLOWUP <= SPG_UNI_LOG, "ecdsa air eval: size");
  int rc = ensure_eair_tables(ctx, log_n);
  if (rc) return rc;
  EcdsaAirConsts K;
  eair_consts(ctx, h_alpha_pows, K);
  void* dk;
  SPG_CUDA(spg_scratch(ctx, 7, sizeof(EcdsaAirConsts) + 4096, &dk));
  SPG_CUDA(cudaMemcpyAsync(dk, &K, sizeof(K), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));   // K is a stack object
  const size_t total = (size_t)n_even << log_n;
  k_air_eval_ecdsa<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(log_n, t_lde, (const EcdsaAirConsts*)dk, ctx->eair_izt,
                                                                            ctx->eair_plde, pub_lde, cp, first_coset, jj0, n_even);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ host evaluation at one point (prover self-check)
Fp spg_eair_composition_at_host(spg_ctx* ctx, unsigned log_n, const Fp* pub_z /*F_msg(z), F_key(z)*/, const Fp* alpha_pows,
                                const Fp& z, const Fp* tz, const Fp* tzw) {
  const uint64_t n = 1ull << log_n;
  const int B = SPG_EAIR_BLOCK;
  const Fp one = fp_one();
  EcdsaAirConsts K;
  eair_consts(ctx, alpha_pows, K);
  const Fp u = fp_pow_u64(z, n >> 8), w256 = spg_host_root_of_unity(8);
  const Fp iz_all = fp_inv(fp_sub(fp_pow_u64(z, n), one));
  Fp t4 = one;
  for (int t = SPG_EAIR_BITS; t < B - 1; t++) t4 = fp_mul(t4, fp_sub(u, fp_pow_u64(w256, t)));
  const Fp el = fp_sub(u, fp_pow_u64(w256, B - 1)), t5 = fp_mul(t4, el);
  Fp iz[SPG_EAIR_NGROUPS] = {fp_mul(t5, iz_all), fp_inv(t4), fp_inv(t5), fp_inv(fp_sub(u, one)), fp_inv(el), fp_mul(el, iz_all)};
  // periodic columns at z: Lagrange over the 256-th roots, L_r(u) = (u^256 - 1) w^r / (256 (u - w^r)), one batch inversion
  Fp gx = fp_zero(), gy = fp_zero();
  {
    const Fp c = fp_mul(fp_sub(fp_pow_u64(u, B), one), fp_inv(eh_small(B)));
    Fp wr[SPG_EAIR_BLOCK], den[SPG_EAIR_BLOCK], pre[SPG_EAIR_BLOCK];
    wr[0] = one;
    for (int r = 1; r < B; r++) wr[r] = fp_mul(wr[r - 1], w256);
    Fp run = one;
    for (int r = 0; r < B; r++) { den[r] = fp_sub(u, wr[r]); pre[r] = run; run = fp_mul(run, den[r]); }
    Fp inv_run = fp_inv(run);
    for (int r = B - 1; r >= 0; r--) {
      const Fp inv_r = fp_mul(inv_run, pre[r]);
      inv_run = fp_mul(inv_run, den[r]);
      if (r < SPG_EAIR_BITS) {
        const Fp lr = fp_mul(fp_mul(c, wr[r]), inv_r);
        gx = fp_add(gx, fp_mul(lr, ctx->h_gen_doubles[2 * r]));
        gy = fp_add(gy, fp_mul(lr, ctx->h_gen_doubles[2 * r + 1]));
      }
    }
  }
  return ecdsa_air_point(tz, tzw, gx, gy, pub_z[0], pub_z[1], K, iz);
}

// ------------------------------------------------------------------ witness generation (per-thread code: ecdsa_air_witness.cuh)
__global__ void __launch_bounds__(64) k_eair_walk_ab(unsigned log_n, const Fp* __restrict__ msg, const Fp* __restrict__ rr,
                                                     const Fp* __restrict__ kx, const Fp* __restrict__ ky, Fp* __restrict__ trace,
                                                     uint32_t* __restrict__ status, const APoint* __restrict__ gd,
                                                     APoint shift, Fp beta) {
  eair_walk_ab_thread((size_t)blockIdx.x * blockDim.x + threadIdx.x, log_n, msg, rr, kx, ky, trace, status, gd, shift, beta);
}
__global__ void __launch_bounds__(64) k_eair_walk_c(unsigned log_n, const Fp* __restrict__ msg, const Fp* __restrict__ rr,
                                                    const Fp* __restrict__ ww, Fp* __restrict__ trace, Fp* __restrict__ cross,
                                                    uint32_t* __restrict__ status, APoint shift) {
  eair_walk_c_thread((size_t)blockIdx.x * blockDim.x + threadIdx.x, log_n, msg, rr, ww, trace, cross, status, shift);
}
__global__ void __launch_bounds__(128) k_eair_finish(unsigned log_n, Fp* __restrict__ trace, const Fp* __restrict__ cross,
                                                     uint32_t* __restrict__ status, const APoint* __restrict__ gd) {
  eair_finish_thread((size_t)blockIdx.x * blockDim.x + threadIdx.x, log_n, trace, cross, status, gd);
}

// msg / r / w / key x / key y: [N/256] canonical (device); trace: [25][N] canonical (device), fully written
int spg_eair_trace_device(spg_ctx* ctx, unsigned log_n, const Fp* msg, const Fp* rr, const Fp* ww, const Fp* kx, const Fp* ky,
                          Fp* trace, uint32_t* d_status) {
  SPG_ARG(log_n >= 9 && log_n <= 23, "ecdsa air trace: size");
  const size_t n = (size_t)1 << log_n, nb = n >> 8;
  SPG_CUDA(cudaMemsetAsync(trace, 0, (size_t)SPG_EAIR_COLS * n * sizeof(Fp), ctx->stream));
  DevBuf cross;
  SPG_CUDA(cross.alloc(ctx, 4 * nb * sizeof(Fp)));
  SPG_CUDA(cudaMemsetAsync(cross.p, 0, 4 * nb * sizeof(Fp), ctx->stream));
  APoint shift; shift.x = ctx->h_const_points[0]; shift.y = ctx->h_const_points[1];
  const APoint* gd = (const APoint*)ctx->gen_doubles;
  k_eair_walk_ab<<<(unsigned)((2 * nb + 63) / 64), 64, 0, ctx->stream>>>(log_n, msg, rr, kx, ky, trace, d_status, gd, shift,
                                                                        spg_host_from_u64(SPG_BETA));
  SPG_LAUNCH_CHECK();
  k_eair_walk_c<<<(unsigned)((nb + 63) / 64), 64, 0, ctx->stream>>>(log_n, msg, rr, ww, trace, cross.as<Fp>(), d_status, shift);
  SPG_LAUNCH_CHECK();
  const size_t threads = 3 * (n / SPG_EAIR_WIT_ROWS);
  k_eair_finish<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(log_n, trace, cross.as<Fp>(), d_status, gd);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ C-ABI
extern "C" int spg_ecdsa_air_trace(spg_ctx* ctx, unsigned log_n, const uint64_t* msg, const uint64_t* r, const uint64_t* w,
                                   const uint64_t* key_x, const uint64_t* key_y, uint64_t* trace_out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && msg && r && w && key_x && key_y && trace_out, "spg_ecdsa_air_trace: null");
  SPG_ARG(log_n >= 9 && log_n <= 23, "spg_ecdsa_air_trace: log_n must be in [9, 23]");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << log_n, nb = n >> 8, bytes = (size_t)SPG_EAIR_COLS * n * 32;
  const uint64_t* in[5] = {msg, r, w, key_x, key_y};
  const Fp* din[5];
  DevBuf bin[5], dt, ds;
  SPG_CUDA(ds.alloc(ctx, 4));
  SPG_CUDA(cudaMemsetAsync(ds.p, 0, 4, ctx->stream));
  Fp* dtr = (Fp*)trace_out;
  if (!(flags & SPG_DEVICE_PTRS)) {
    for (int k = 0; k < 5; k++) {
      SPG_CUDA(bin[k].alloc(ctx, nb * 32));
      SPG_CUDA(cudaMemcpyAsync(bin[k].p, in[k], nb * 32, cudaMemcpyHostToDevice, ctx->stream));
      din[k] = bin[k].as<Fp>();
    }
    SPG_CUDA(dt.alloc(ctx, bytes));
    dtr = dt.as<Fp>();
  } else {
    for (int k = 0; k < 5; k++) din[k] = (const Fp*)in[k];
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = spg_eair_trace_device(ctx, log_n, din[0], din[1], din[2], din[3], din[4], dtr, ds.as<uint32_t>());
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) SPG_CUDA(cudaMemcpyAsync(trace_out, dtr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  uint32_t st = 0;
  SPG_CUDA(cudaMemcpyAsync(&st, ds.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  if (st & 1) { ctx->err = "spg_ecdsa_air_trace: a scalar is outside [1, 2^251) or a key is not on the curve"; return SPG_E_ARG; }
  if (st & 2) { ctx->err = "spg_ecdsa_air_trace: an assertion of mimic_ec_mult_air / ec_add fires (x collision)"; return SPG_E_ARG; }
  if (st & 4) { ctx->err = "spg_ecdsa_air_trace: a signature does not verify"; return SPG_E_ARG; }
  return SPG_OK;
}

// composition polynomial of an ECDSA-AIR trace on the cosets j = 0, 2, 4, 6 (parity entry point for the AIR stage):
// trace [25][N] canonical -> cp [4][N] canonical; msgs, key_x [N/256] canonical = the public input; alpha canonical.
extern "C" int spg_air_eval_ecdsa(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, const uint64_t* msgs, const uint64_t* key_x,
                                  const uint64_t* alpha, uint64_t* cp_out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && trace && msgs && key_x && alpha && cp_out, "spg_air_eval_ecdsa: null");
  SPG_ARG(log_n >= 9 && log_n <= 23, "spg_air_eval_ecdsa: size");
  SPG_ARG(!(flags & SPG_DEVICE_PTRS), "spg_air_eval_ecdsa: host pointers only");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << log_n, nb = n >> 8;
  const int C = SPG_EAIR_COLS;
  DevBuf dt, dl, dc, dcp, dpub, dcs, dpl;
  SPG_CUDA(dt.alloc(ctx, C * n * 32)); SPG_CUDA(dl.alloc(ctx, 8 * C * n * 32)); SPG_CUDA(dc.alloc(ctx, C * n * 32));
  SPG_CUDA(dcp.alloc(ctx, 4 * n * 32)); SPG_CUDA(dpub.alloc(ctx, 2 * nb * 32)); SPG_CUDA(dcs.alloc(ctx, 2 * nb * 32));
  SPG_CUDA(dpl.alloc(ctx, 8 * n * 32));
  SPG_CUDA(cudaMemcpyAsync(dt.p, trace, C * n * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(dpub.p, msgs, nb * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(dpub.as<Fp>() + nb, key_x, nb * 32, cudaMemcpyHostToDevice, ctx->stream));
  Fp apows[SPG_EAIR_NALPHA];
  const Fp a = spg_host_from_u64(alpha);
  apows[0] = fp_one();
  for (int k = 1; k < SPG_EAIR_NALPHA; k++) apows[k] = fp_mul(apows[k - 1], a);
  int rc = spg_lde_device(ctx, dt.as<Fp>(), log_n, C, SPG_LOG_BLOWUP, nullptr, dl.as<Fp>(), dc.as<Fp>(), 1);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  if ((rc = spg_eair_public_device(ctx, log_n, dpub.as<Fp>(), dcs.as<Fp>(), dpl.as<Fp>(), 0, 4))) return rc;
  rc = spg_eair_eval_device(ctx, log_n, dl.as<Fp>(), dpl.as<Fp>(), apows, dcp.as<Fp>(), 0, 0, 4);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  rc = spg_from_mont_device(ctx, dcp.as<Fp>(), 4 * n);
  if (rc) return rc;
  SPG_CUDA(cudaMemcpyAsync(cp_out, dcp.p, 4 * n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

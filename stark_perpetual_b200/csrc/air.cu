// The Pedersen hash-chain AIR: witness generation, constraint (composition) evaluation over the LDE
// cosets, and the split of the composition polynomial into degree < N chunks.
// (BASELINE.json configs[2].)  The reference has no AIR; what it pins is the
// step function each trace row encodes: signature.py:300-318 (pedersen_hash_as_point), math_utils.py:59-68
// (ec_add).  Constraint list and zerofiers: DESIGN.md "AIR", oracle/stark.py class Air.
#include "stark_kernels.cuh"

// ------------------------------------------------------------------ cached tables
struct AirTableConsts {
  Fp g256, g512, gseg;          // g^(N/256), g^(N/512), g^(N/seg)
  Fp w256[4];                   // w_256^252 .. w_256^255
  Fp w256_canon;                // w_256^251: c6 (M = 0) starts at row 251 -- canonical 251-bit unpacking (DESIGN.md "AIR")
  Fp w512_255, w512_511, wseg_inv;
  Fp iz_all[4];                 // 1 / (x^N - 1) on cosets 0, 2, 4, 6
};

// izt[t][jj][i], t: 0 step, 1 act, 2 pad, 3 mid, 4 link, 5 inst0, 6 seg0;  i < seg
__global__ void __launch_bounds__(128) k_air_tables(unsigned log_seg, AirTableConsts K, Fp* __restrict__ izt,
                                                    const Fp* __restrict__ uniA, const Fp* __restrict__ uniB) {
  const size_t seg = (size_t)1 << log_seg;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * seg) return;
  const size_t jj = idx >> log_seg, i = idx & (seg - 1);
  const unsigned long long e = 2ull * jj + 8ull * i;
  // x^(N/256) = g256 * w_2048^e, x^(N/512) = g512 * w_4096^e, x^(N/seg) = gseg * w_{8 seg}^e
  const Fp u256 = fp_mul(K.g256, spg_uni_pow(uniA, uniB, e << (SPG_UNI_LOG - 11)));
  const Fp u512 = fp_mul(K.g512, spg_uni_pow(uniA, uniB, e << (SPG_UNI_LOG - 12)));
  const Fp useg = fp_mul(K.gseg, spg_uni_pow(uniA, uniB, e << (SPG_UNI_LOG - 3 - log_seg)));
  const Fp e_step = fp_sub(u256, K.w256[3]);
  Fp z_pad = fp_sub(u256, K.w256[0]);
#pragma unroll
  for (int k = 1; k < 4; k++) z_pad = fp_mul(z_pad, fp_sub(u256, K.w256[k]));
  // five inversions with one Fermat chain
  const Fp a = fp_mul(z_pad, fp_sub(u256, K.w256_canon)), b = fp_sub(u512, K.w512_255), c = fp_sub(u512, K.w512_511), d = fp_sub(u512, fp_one()),
           f = fp_sub(useg, fp_one());
  const Fp p2 = fp_mul(a, b), p3 = fp_mul(p2, c), p4 = fp_mul(p3, d), p5 = fp_mul(p4, f);
  Fp inv = fp_inv_chain(p5);
  const Fp fi = fp_mul(inv, p4); inv = fp_mul(inv, f);
  const Fp di = fp_mul(inv, p3); inv = fp_mul(inv, d);
  const Fp ci = fp_mul(inv, p2); inv = fp_mul(inv, c);
  const Fp bi = fp_mul(inv, a);
  const Fp ai = fp_mul(inv, b);
  const size_t stride = 4 * seg;
  izt[0 * stride + idx] = fp_reduce(fp_mul(e_step, K.iz_all[jj]));
  izt[1 * stride + idx] = fp_reduce(fp_mul(z_pad, K.iz_all[jj]));
  izt[2 * stride + idx] = fp_reduce(ai);
  izt[3 * stride + idx] = fp_reduce(bi);
  izt[4 * stride + idx] = fp_reduce(fp_mul(fp_sub(useg, K.wseg_inv), ci));
  izt[5 * stride + idx] = fp_reduce(di);
  izt[6 * stride + idx] = fp_reduce(fi);
}

static Fp h_from_small(uint64_t v) { uint64_t w[4] = {v, 0, 0, 0}; return spg_host_from_u64(w); }

static int ensure_air_tables(spg_ctx* ctx, unsigned log_n, unsigned chain_log) {
  if (ctx->air_log_n == (int)log_n && ctx->air_chain_log == (int)chain_log) return SPG_OK;
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->air_izt); cudaFree(ctx->air_plde); cudaFree(ctx->air_ilast);
  ctx->air_izt = ctx->air_plde = ctx->air_ilast = nullptr;
  ctx->air_log_n = ctx->air_chain_log = -1;
  const size_t n = (size_t)1 << log_n;
  const unsigned log_seg = 9 + chain_log;
  const size_t seg = (size_t)1 << log_seg;
  const Fp g = h_from_small(3);
  AirTableConsts K;
  K.g256 = fp_pow_u64(g, n >> 8); K.g512 = fp_pow_u64(g, n >> 9); K.gseg = fp_pow_u64(g, n >> log_seg);
  const Fp w256 = spg_host_root_of_unity(8), w512 = spg_host_root_of_unity(9), wseg = spg_host_root_of_unity((int)log_seg);
  for (int k = 0; k < 4; k++) K.w256[k] = fp_pow_u64(w256, 252 + k);
  K.w256_canon = fp_pow_u64(w256, SPG_AIR_CANON_BITS);
  K.w512_255 = fp_pow_u64(w512, 255); K.w512_511 = fp_pow_u64(w512, 511); K.wseg_inv = fp_inv(wseg);
  const Fp gn = fp_pow_u64(g, n), w8 = spg_host_root_of_unity(3);
  for (int jj = 0; jj < 4; jj++) K.iz_all[jj] = fp_inv(fp_sub(fp_mul(gn, fp_pow_u64(w8, 2 * jj)), fp_one()));
  SPG_CUDA(cudaMalloc((void**)&ctx->air_izt, 7 * 4 * seg * sizeof(Fp)));
  k_air_tables<<<(unsigned)((4 * seg + 127) / 128), 128, 0, ctx->stream>>>(log_seg, K, ctx->air_izt, ctx->uniA, ctx->uniB);
  SPG_LAUNCH_CHECK();
  // periodic point columns PX, PY over one 512-row instance (0 on the padding rows), extended to the cosets
  std::vector<Fp> pcols(2 * 512, fp_zero());
  for (int e = 0; e < 2; e++)
    for (int t = 0; t < SPG_HASH_BITS; t++) {
      pcols[256 * e + t] = ctx->h_const_points[2 * (2 + SPG_HASH_BITS * e + t)];
      pcols[512 + 256 * e + t] = ctx->h_const_points[2 * (2 + SPG_HASH_BITS * e + t) + 1];
    }
  DevBuf dp;
  SPG_CUDA(dp.alloc(ctx, pcols.size() * sizeof(Fp)));
  SPG_CUDA(cudaMemcpyAsync(dp.p, pcols.data(), pcols.size() * sizeof(Fp), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMalloc((void**)&ctx->air_plde, 8 * 2 * 512 * sizeof(Fp)));
  uint64_t off[4];
  spg_host_to_u64(K.g512, off);
  DevBuf dcoef;
  SPG_CUDA(dcoef.alloc(ctx, 2 * 512 * sizeof(Fp)));
  int rc = spg_lde_device(ctx, dp.as<Fp>(), 9, 2, SPG_LOG_BLOWUP, off, ctx->air_plde, dcoef.as<Fp>(), 0);
  if (rc) return rc;
  // 1 / (x - w_N^(N-1)) on cosets 0, 2, 4, 6
  const Fp last = fp_inv(spg_host_root_of_unity((int)log_n));   // w^(N-1) = w^-1
  DevBuf dl;
  SPG_CUDA(dl.alloc(ctx, sizeof(Fp)));
  SPG_CUDA(cudaMemcpyAsync(dl.p, &last, sizeof(Fp), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMalloc((void**)&ctx->air_ilast, 4 * n * sizeof(Fp)));
  rc = spg_inv_x_minus_device(ctx, log_n, 0, 2, 4, dl.as<Fp>(), 1, ctx->air_ilast);
  if (rc) return rc;
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->air_log_n = (int)log_n; ctx->air_chain_log = (int)chain_log;
  return SPG_OK;
}

// ------------------------------------------------------------------ composition evaluation
#ifndef AIR_MIN_CTAS
#define AIR_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(128, AIR_MIN_CTAS) k_air_eval(unsigned log_n, unsigned log_seg, const Fp* __restrict__ t_lde,
                                                  const AirEvalConsts* __restrict__ K, const Fp* __restrict__ izt,
                                                  const Fp* __restrict__ plde, const Fp* __restrict__ ilast,
                                                  Fp* __restrict__ cp, int first_coset, int jj0, int n_even) {
  const size_t n = (size_t)1 << log_n, seg = (size_t)1 << log_seg;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_even * n) return;
  const size_t jj = (idx >> log_n) + jj0, i = idx & (n - 1), j = 2 * jj;
  const size_t in = (i + 1) & (n - 1);
  const Fp px = plde[(j * 2 + 0) * 512 + (i & 511)], py = plde[(j * 2 + 1) * 512 + (i & 511)];
  AirAcc A;
  air_acc_init(A);
  const Fp* base = t_lde + ((j - first_coset) * SPG_AIR_COLS << log_n);
#pragma unroll 1
  for (int l = 0; l < SPG_AIR_LANES; l++) {
    const Fp* c0 = base + ((size_t)(5 * l) << log_n);
    AirRow r;
    r.X = c0[i]; r.Y = c0[n + i]; r.S = c0[2 * n + i]; r.M = c0[3 * n + i]; r.I = c0[4 * n + i];
    r.Xn = c0[in]; r.Yn = c0[n + in]; r.Mn = c0[3 * n + in];
    air_lane_accumulate(A, r, K->alpha + SPG_AIR_NCONSTR * l, px, py, K->shift_x, K->shift_y, K->x0[l], K->outs[l]);
  }
  const size_t stride = 4 * seg, zi = jj * seg + (i & (seg - 1));
  const Fp acc = air_combine(A, izt[0 * stride + zi], izt[1 * stride + zi], izt[2 * stride + zi], izt[3 * stride + zi],
                             izt[4 * stride + zi], izt[5 * stride + zi], izt[6 * stride + zi], ilast[(jj << log_n) + i]);
  cp[idx] = acc;
}

int spg_air_eval_device(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const Fp* t_lde, const AirPublic& pub,
                        const Fp* h_alpha_pows, Fp* cp, int first_coset, int jj0, int n_even) {
  if (n_even <= 0) return SPG_OK;
  SPG_ARG(2 * jj0 >= first_coset && jj0 + n_even <= 4, "air eval: coset range");
  SPG_ARG(log_n >= 9 && log_n + SPG_LOG_BLOWUP <= SPG_UNI_LOG && 9 + chain_log <= log_n, "air eval: size");
  int rc = ensure_air_tables(ctx, log_n, chain_log);
  if (rc) return rc;
  AirEvalConsts K;
  for (int k = 0; k < SPG_AIR_LANES * SPG_AIR_NCONSTR; k++) K.alpha[k] = h_alpha_pows[k];
  for (int l = 0; l < SPG_AIR_LANES; l++) { K.x0[l] = pub.x0[l]; K.outs[l] = pub.outs[l]; }
  K.shift_x = ctx->h_const_points[0]; K.shift_y = ctx->h_const_points[1];
  void* dk;
  SPG_CUDA(spg_scratch(ctx, 7, sizeof(AirEvalConsts), &dk));
  SPG_CUDA(cudaMemcpyAsync(dk, &K, sizeof(K), cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));   // K is a stack object
  const size_t total = (size_t)n_even << log_n;
  k_air_eval<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(log_n, 9 + chain_log, t_lde, (const AirEvalConsts*)dk,
                                                                      ctx->air_izt, ctx->air_plde, ctx->air_ilast, cp,
                                                                      first_coset, jj0, n_even);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ chunk split
// cp[jj][i] on x = g w_{8N}^(2 jj + 8 i).  With CP(x) = sum_m x^m H_m(x^4):
//   H_m(x^4) x^m = 1/4 sum_k w_4^(-m k) CP(x w_4^k),  and  x w_4^k  is row i + k N/4 of the same coset.
__global__ void __launch_bounds__(256) k_cp_split(unsigned log_n, const Fp* __restrict__ cp, Fp* __restrict__ hev, Fp ginv,
                                                  Fp inv4, Fp iota_inv, const Fp* __restrict__ uniA,
                                                  const Fp* __restrict__ uniB, int jj0, int n_even) {
  const size_t n = (size_t)1 << log_n, q = n >> 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_even * q) return;
  const size_t e_loc = idx / q, ip = idx - e_loc * q, jj = e_loc + jj0;
  const Fp* p = cp + (e_loc << log_n) + ip;
  const Fp v0 = p[0], v1 = p[q], v2 = p[2 * q], v3 = p[3 * q];
  const Fp t0 = fp_add(v0, v2), t1 = fp_sub(v0, v2), t2 = fp_add(v1, v3), t3 = fp_mul(fp_sub(v1, v3), iota_inv);
  const Fp s0 = fp_add(t0, t2), s2 = fp_sub(t0, t2), s1 = fp_add(t1, t3), s3 = fp_sub(t1, t3);
  const int sh = SPG_UNI_LOG - (int)log_n - SPG_LOG_BLOWUP;
  const unsigned long long e = (2ull * jj + 8ull * ip) << sh;
  const Fp xi = fp_mul(ginv, spg_uni_pow(uniA, uniB, 0ull - e));     // 1 / x
  const Fp xi2 = fp_sqr(xi), xi3 = fp_mul(xi2, xi);
  const size_t pos = jj + 4 * ip;
  hev[pos] = fp_reduce(fp_mul(s0, inv4));
  hev[n + pos] = fp_reduce(fp_mul(fp_mul(s1, inv4), xi));
  hev[2 * n + pos] = fp_reduce(fp_mul(fp_mul(s2, inv4), xi2));
  hev[3 * n + pos] = fp_reduce(fp_mul(fp_mul(s3, inv4), xi3));
}

int spg_cp_split_device(spg_ctx* ctx, unsigned log_n, const Fp* cp, Fp* hev, int jj0, int n_even) {
  if (n_even <= 0) return SPG_OK;
  const Fp ginv = fp_inv(h_from_small(3)), inv4 = fp_inv(h_from_small(4));
  const Fp iota_inv = fp_inv(spg_host_root_of_unity(2));
  const size_t total = ((size_t)1 << log_n) / 4 * n_even;
  k_cp_split<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(log_n, cp, hev, ginv, inv4, iota_inv, ctx->uniA,
                                                                        ctx->uniB, jj0, n_even);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ host evaluation at one point (prover self-check)
Fp spg_air_composition_at_host(unsigned log_n, unsigned chain_log, const AirPublic& pub, const Fp* al_all, const Fp& z,
                               const Fp* tz, const Fp* tzw, const std::vector<Fp>& cpts) {
  const uint64_t n = 1ull << log_n;
  const unsigned log_seg = 9 + chain_log;
  const Fp one = fp_one();
  const Fp u256 = fp_pow_u64(z, n >> 8), u512 = fp_pow_u64(z, n >> 9), useg = fp_pow_u64(z, n >> log_seg);
  const Fp w256 = spg_host_root_of_unity(8), w512 = spg_host_root_of_unity(9), wseg = spg_host_root_of_unity((int)log_seg);
  const Fp iz_all = fp_inv(fp_sub(fp_pow_u64(z, n), one));
  Fp z_pad = one;
  for (int k = 252; k < 256; k++) z_pad = fp_mul(z_pad, fp_sub(u256, fp_pow_u64(w256, k)));
  const Fp iz_step = fp_mul(fp_sub(u256, fp_pow_u64(w256, 255)), iz_all);
  const Fp iz_act = fp_mul(z_pad, iz_all);
  const Fp iz_pad = fp_inv(fp_mul(z_pad, fp_sub(u256, fp_pow_u64(w256, SPG_AIR_CANON_BITS))));   // M = 0 on rows >= 251
  const Fp iz_mid = fp_inv(fp_sub(u512, fp_pow_u64(w512, 255)));
  const Fp iz_link = fp_mul(fp_sub(useg, fp_inv(wseg)), fp_inv(fp_sub(u512, fp_pow_u64(w512, 511))));
  const Fp iz_inst = fp_inv(fp_sub(u512, one)), iz_seg = fp_inv(fp_sub(useg, one));
  const Fp iz_last = fp_inv(fp_sub(z, fp_inv(spg_host_root_of_unity((int)log_n))));
  // periodic columns at z: barycentric-free, direct Lagrange over the 512-th roots: P(u) = sum_r v_r L_r(u),
  // L_r(u) = (u^512 - 1) w^r / (512 (u - w^r))
  Fp px = fp_zero(), py = fp_zero();
  {
    const Fp u512p = fp_sub(fp_pow_u64(u512, 512), one);
    const Fp c = fp_mul(u512p, fp_inv(h_from_small(512)));
    // 1 / (u - w^r) for all r with ONE inversion (Montgomery's trick): this check runs once per proof on the host
    Fp wr[512], den[512], pre[512];
    wr[0] = one;
    for (int r = 1; r < 512; r++) wr[r] = fp_mul(wr[r - 1], w512);
    Fp run = one;
    for (int r = 0; r < 512; r++) { den[r] = fp_sub(u512, wr[r]); pre[r] = run; run = fp_mul(run, den[r]); }
    Fp inv_run = fp_inv(run);
    for (int r = 511; r >= 0; r--) {
      const Fp inv_r = fp_mul(inv_run, pre[r]);
      inv_run = fp_mul(inv_run, den[r]);
      const int e = r >> 8, t = r & 255;
      if (t < SPG_HASH_BITS) {
        const Fp lr = fp_mul(fp_mul(c, wr[r]), inv_r);
        px = fp_add(px, fp_mul(lr, cpts[2 * (2 + SPG_HASH_BITS * e + t)]));
        py = fp_add(py, fp_mul(lr, cpts[2 * (2 + SPG_HASH_BITS * e + t) + 1]));
      }
    }
  }
  Fp acc = fp_zero();
  for (int l = 0; l < SPG_AIR_LANES; l++) {
    const Fp X = tz[5 * l], Y = tz[5 * l + 1], S = tz[5 * l + 2], M = tz[5 * l + 3], I = tz[5 * l + 4];
    const Fp Xn = tzw[5 * l], Yn = tzw[5 * l + 1], Mn = tzw[5 * l + 3];
    const Fp* al = al_all + SPG_AIR_NCONSTR * l;
    const Fp bit = fp_sub(M, fp_add(Mn, Mn)), nb = fp_sub(one, bit), dx = fp_sub(X, px);
    const Fp c1 = fp_mul(bit, fp_sub(bit, one));
    const Fp c2 = fp_mul(bit, fp_sub(fp_mul(S, dx), fp_sub(Y, py)));
    const Fp c3 = fp_add(fp_mul(bit, fp_sub(fp_sub(fp_sub(fp_sqr(S), X), px), Xn)), fp_mul(nb, fp_sub(Xn, X)));
    const Fp c4 = fp_add(fp_mul(bit, fp_sub(fp_sub(fp_mul(S, fp_sub(X, Xn)), Y), Yn)), fp_mul(nb, fp_sub(Yn, Y)));
    Fp s = fp_add(fp_add(fp_mul(al[0], c1), fp_mul(al[1], c2)), fp_add(fp_mul(al[2], c3), fp_mul(al[3], c4)));
    acc = fp_add(acc, fp_mul(s, iz_step));
    acc = fp_add(acc, fp_mul(fp_mul(al[4], fp_sub(fp_mul(I, dx), one)), iz_act));
    acc = fp_add(acc, fp_mul(fp_mul(al[5], M), iz_pad));
    acc = fp_add(acc, fp_mul(fp_add(fp_mul(al[6], fp_sub(Xn, X)), fp_mul(al[7], fp_sub(Yn, Y))), iz_mid));
    acc = fp_add(acc, fp_mul(fp_mul(al[8], fp_sub(Mn, X)), iz_link));
    acc = fp_add(acc, fp_mul(fp_add(fp_mul(al[9], fp_sub(X, cpts[0])), fp_mul(al[10], fp_sub(Y, cpts[1]))), iz_inst));
    acc = fp_add(acc, fp_mul(fp_mul(al[11], fp_sub(M, pub.x0[l])), iz_seg));
    acc = fp_add(acc, fp_mul(fp_mul(al[12], fp_sub(X, pub.outs[l])), iz_last));
  }
  return fp_reduce(acc);
}

// ------------------------------------------------------------------ witness generation
// Two kernels.  (1) k_pedersen_walk: one thread per (lane, segment) runs the 2^chain_log chained hashes of the segment
// step by step as signature.py:300-318 does, but in Jacobian coordinates (no inversion on the sequential path) and parks
// (X, Y, Z) of the partial sum BEFORE each row's step in the X, Y, S columns, the remaining scalar in M.  (2)
// k_pedersen_finish: one thread per 16 consecutive rows of a lane turns them into the affine witness -- x = X / Z^2,
// y = Y / Z^3, I = 1 / (x - px) = Z^2 / (X - px Z^2), S = bit ? (y - py) I : 0 -- with ONE Fermat inversion per 32 values
// (Montgomery's trick) instead of one per row.  The segment's hashes are chained through the affine x of the last
// row, which the walk needs: that one inversion per hash stays on the sequential path (4 per thread at chain_log 2).
#define SPG_WIT_ROWS 16
__global__ void __launch_bounds__(64) k_pedersen_walk(unsigned log_n, unsigned chain_log, const Fp* __restrict__ x0,
                                                      const Fp* __restrict__ ys, Fp* __restrict__ trace,
                                                      uint32_t* __restrict__ status, const APoint* __restrict__ cp) {
  const size_t n = (size_t)1 << log_n, inst = n >> 9, nseg = inst >> chain_log;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= SPG_AIR_LANES * nseg) return;
  const size_t l = idx / nseg, sgi = idx - l * nseg;
  Fp* X = trace + ((5 * l) << log_n);
  Fp *Y = X + n, *Z = X + 2 * n, *M = X + 3 * n;
  uint32_t st = 0;
  Fp a = x0[l];                                   // canonical
  for (size_t q = sgi << chain_log; q < ((sgi + 1) << chain_log); q++) {
    JPoint ps; ps.X = cp[0].x; ps.Y = cp[0].y; ps.Z = fp_one();
    for (int e = 0; e < 2; e++) {
      Fp v = e ? ys[l * inst + q] : a;            // canonical scalar, shifted right one bit per row
      if (spg_canon_geq_p(v.v)) st |= 1;
      else if (v.v[7] >> 27) st |= 4;             // >= 2^251: outside the AIR's canonical 251-bit unpacking
      for (int t = 0; t < 256; t++) {
        const size_t r = (q << 9) + 256 * e + t;
        X[r] = ps.X; Y[r] = ps.Y; Z[r] = ps.Z; M[r] = v;
        if (t < SPG_HASH_BITS && (v.v[0] & 1u)) {
          const APoint pt = cp[2 + SPG_HASH_BITS * e + t];
          const Fp zz = fp_sqr(ps.Z);
          ps = ec_madd_nocheck(ps, pt, fp_mul(zz, ps.Z), fp_mul(pt.x, zz));
        }
#pragma unroll
        for (int k = 0; k < 7; k++) v.v[k] = (v.v[k] >> 1) | (v.v[k + 1] << 31);
        v.v[7] >>= 1;
      }
    }
    const Fp zi = fp_inv_chain(ps.Z);             // Z = 0 only after an x-collision, which k_pedersen_finish reports
    a = fp_from_mont(fp_mul(ps.X, fp_sqr(zi)));
  }
  if (st) atomicOr(status, st);
}

__global__ void __launch_bounds__(128) k_pedersen_finish(unsigned log_n, Fp* __restrict__ trace, uint32_t* __restrict__ status,
                                                         const APoint* __restrict__ cp) {
  const size_t n = (size_t)1 << log_n, blocks_per_lane = n / SPG_WIT_ROWS;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= SPG_AIR_LANES * blocks_per_lane) return;
  const size_t l = idx / blocks_per_lane, r0 = (idx - l * blocks_per_lane) * SPG_WIT_ROWS;
  Fp* X = trace + ((5 * l) << log_n);
  Fp *Y = X + n, *S = X + 2 * n, *M = X + 3 * n, *I = X + 4 * n;
  // values to invert: Z_k and D_k = X_k - px Z_k^2 (1 on the padding rows), prefix products kept in local memory
  Fp pre[2 * SPG_WIT_ROWS];
  Fp run = fp_one();
  uint32_t st = 0;
#pragma unroll 1
  for (int k = 0; k < SPG_WIT_ROWS; k++) {
    const size_t r = r0 + k;
    const int t = (int)(r & 255), e = (int)((r >> 8) & 1);
    const Fp z = S[r];
    Fp d = fp_one();
    if (t < SPG_HASH_BITS) {
      d = fp_sub(X[r], fp_mul(cp[2 + SPG_HASH_BITS * e + t].x, fp_sqr(z)));
      if (fp_is_zero(d)) { st |= 2; d = fp_one(); }           // "Unhashable input." (signature.py:313)
    }
    pre[2 * k] = run; run = fp_mul(run, z);
    pre[2 * k + 1] = run; run = fp_mul(run, d);
  }
  if (fp_is_zero(run)) { st |= 2; run = fp_one(); }            // a partial sum at infinity (follows a collision)
  Fp inv = fp_inv_chain(run);
#pragma unroll 1
  for (int k = SPG_WIT_ROWS - 1; k >= 0; k--) {
    const size_t r = r0 + k;
    const int t = (int)(r & 255), e = (int)((r >> 8) & 1);
    const Fp z = S[r], x_j = X[r], y_j = Y[r];
    const Fp zz = fp_sqr(z);
    Fp d = fp_one();
    APoint pt; pt.x = fp_zero(); pt.y = fp_zero();
    if (t < SPG_HASH_BITS) {
      pt = cp[2 + SPG_HASH_BITS * e + t];
      d = fp_sub(x_j, fp_mul(pt.x, zz));
      if (fp_is_zero(d)) d = fp_one();
    }
    const Fp di = fp_mul(inv, pre[2 * k + 1]); inv = fp_mul(inv, d);
    const Fp zi = fp_mul(inv, pre[2 * k]); inv = fp_mul(inv, z);
    const Fp zi2 = fp_sqr(zi);
    const Fp x = fp_mul(x_j, zi2), y = fp_mul(y_j, fp_mul(zi2, zi));
    Fp s_out = fp_zero(), i_out = fp_zero();
    if (t < SPG_HASH_BITS) {
      const Fp ii = fp_mul(zz, di);                            // 1 / (x - px)
      i_out = fp_from_mont(ii);
      if (M[r].v[0] & 1u) s_out = fp_from_mont(fp_mul(fp_sub(y, pt.y), ii));
    }
    X[r] = fp_from_mont(x); Y[r] = fp_from_mont(y); S[r] = s_out; I[r] = i_out;
  }
  if (st) atomicOr(status, st);
}

int spg_pedersen_trace_device(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const Fp* d_x0, const Fp* d_ys, Fp* trace,
                              uint8_t* d_status) {
  SPG_ARG(log_n >= 9 && 9 + chain_log <= log_n, "pedersen trace: size");
  const size_t n = (size_t)1 << log_n, nseg = n >> (9 + chain_log);
  const size_t total = SPG_AIR_LANES * nseg;
  k_pedersen_walk<<<(unsigned)((total + 63) / 64), 64, 0, ctx->stream>>>(log_n, chain_log, d_x0, d_ys, trace, (uint32_t*)d_status,
                                                                        (const APoint*)ctx->const_points);
  SPG_LAUNCH_CHECK();
  const size_t blocks = SPG_AIR_LANES * (n / SPG_WIT_ROWS);
  k_pedersen_finish<<<(unsigned)((blocks + 127) / 128), 128, 0, ctx->stream>>>(log_n, trace, (uint32_t*)d_status,
                                                                              (const APoint*)ctx->const_points);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

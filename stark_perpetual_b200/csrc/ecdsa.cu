// Batched STARK-curve ECDSA verification and key derivation (BASELINE.json
// configs[4]).  Reference: src/starkware/crypto/signature/signature.py:217-260 (verify), :104-110
// (private_to_stark_key), :117-173 (sign).  One thread per signature; per-signature code in ecdsa.cuh / sign.cuh.
#include "common.h"
#include "curve_params.inc"
#include "ecdsa.cuh"
#include "sign.cuh"

static int ensure_ecdsa_tables(spg_ctx* ctx) {
  if (ctx->sqrt_tables) return SPG_OK;
  // c = 3^q, q = 2^59 + 17, generates the subgroup of order 2^192
  uint64_t three[4] = {3, 0, 0, 0};
  const Fp g = spg_host_from_u64(three);
  const Fp c = fp_pow_u64(g, (1ull << 59) + 17), ci = fp_inv(c);
  std::vector<Fp> tab(256 + 2 * 24 * 256, fp_one());
  Fp cl = c;
  for (int k = 0; k < 184; k++) cl = fp_sqr(cl);
  for (int k = 1; k < 256; k++) tab[k] = fp_mul(tab[k - 1], cl);
  Fp* D = tab.data() + 256;
  Fp* Dh = D + 24 * 256;
  Fp base = ci, base_h = ci;   // base = ci^(2^(8i)); base_h = ci^(2^(8i-1)) for i >= 1
  for (int i = 0; i < 24; i++) {
    for (int k = 1; k < 256; k++) D[i * 256 + k] = fp_mul(D[i * 256 + k - 1], base);
    if (i == 0) {
      for (int k = 2; k < 256; k += 2) Dh[k] = fp_mul(Dh[k - 2], ci);       // ci^(k/2), even k only
    } else {
      for (int k = 1; k < 256; k++) Dh[i * 256 + k] = fp_mul(Dh[i * 256 + k - 1], base_h);
    }
    base_h = base;
    for (int k = 0; k < 7; k++) base_h = fp_sqr(base_h);                     // ci^(2^(8i+7)) = ci^(2^(8(i+1)-1))
    for (int k = 0; k < 8; k++) base = fp_sqr(base);
  }
  ctx->h_sqrt_tables = tab;
  SPG_CUDA(cudaMalloc((void**)&ctx->sqrt_tables, tab.size() * sizeof(Fp)));
  SPG_CUDA(cudaMemcpy(ctx->sqrt_tables, tab.data(), tab.size() * sizeof(Fp), cudaMemcpyHostToDevice));
  return SPG_OK;
}

static EcdsaTables make_tables(spg_ctx* ctx, bool device) {
  EcdsaTables T;
  const Fp* cp = ctx->h_const_points.data();
  T.shift.x = cp[0]; T.shift.y = cp[1];
  T.minus_shift.x = cp[0]; T.minus_shift.y = fp_neg(cp[1]);
  T.beta = spg_host_from_u64(SPG_BETA);
  // 2^512 mod n
  static const uint32_t r2[8] = {0xea1c688du, 0x6021b3f1u, 0x14ce60b9u, 0x509cf64du, 0xf78bbabbu, 0xbaf0ab4cu, 0x2333766eu, 0x07d9e57cu};
  for (int i = 0; i < 8; i++) T.r2_n.v[i] = r2[i];
  T.ninv = SPG_N_INV32;
  if (device) {
    T.gen_doubles = (const APoint*)ctx->gen_doubles;
    T.sq_L = ctx->sqrt_tables; T.sq_D = ctx->sqrt_tables + 256; T.sq_Dh = ctx->sqrt_tables + 256 + 24 * 256;
  } else {
    T.gen_doubles = (const APoint*)ctx->h_gen_doubles.data();
    T.sq_L = ctx->h_sqrt_tables.data(); T.sq_D = T.sq_L + 256; T.sq_Dh = T.sq_D + 24 * 256;
  }
  return T;
}

__device__ __forceinline__ void load8(const uint64_t* src, uint32_t (&x)[8]) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  const uint4 lo = s[0], hi = s[1];
  x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w; x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
}

#ifndef ECDSA_MIN_CTAS
#define ECDSA_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(64, ECDSA_MIN_CTAS) k_ecdsa_verify(const uint64_t* __restrict__ msg, const uint64_t* __restrict__ r,
                                                     const uint64_t* __restrict__ s, const uint64_t* __restrict__ px,
                                                     const uint64_t* __restrict__ py, uint8_t* __restrict__ status,
                                                     size_t n, EcdsaTables T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t m[8], rr[8], ss[8], x[8], y[8];
  load8(msg + 4 * i, m); load8(r + 4 * i, rr); load8(s + 4 * i, ss); load8(px + 4 * i, x);
  if (py) load8(py + 4 * i, y);
  status[i] = (uint8_t)ecdsa_verify_one(m, rr, ss, x, py ? y : nullptr, T);
}

// status: 0 ok, 1 key outside (0, n)  (signature.py:105 asserts 0 < priv_key < EC_ORDER)
__global__ void __launch_bounds__(128) k_private_to_public(const uint64_t* __restrict__ priv, uint64_t* __restrict__ pub_x,
                                                           uint64_t* __restrict__ pub_y, uint8_t* __restrict__ status,
                                                           size_t n, EcdsaTables T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
  load8(priv + 4 * i, k);
  Fp res = fp_zero(), resy = fp_zero();
  uint8_t st = 0;
  if (u256_is_zero(k) || fn_geq_n(k)) st = 1;
  else { const APoint q = gen_mult(k, T); res = fp_from_mont(q.x); resy = fp_from_mont(q.y); }
  uint4* o = reinterpret_cast<uint4*>(pub_x + 4 * i);
  o[0] = make_uint4(res.v[0], res.v[1], res.v[2], res.v[3]);
  o[1] = make_uint4(res.v[4], res.v[5], res.v[6], res.v[7]);
  if (pub_y) {
    o = reinterpret_cast<uint4*>(pub_y + 4 * i);
    o[0] = make_uint4(resy.v[0], resy.v[1], resy.v[2], resy.v[3]);
    o[1] = make_uint4(resy.v[4], resy.v[5], resy.v[6], resy.v[7]);
  }
  status[i] = st;
}

// device-pointer building block (orders.cu): enqueue the verification kernel on the context stream
int spg_ecdsa_verify_device(spg_ctx* ctx, const uint64_t* msg, const uint64_t* r, const uint64_t* s, const uint64_t* px,
                            const uint64_t* py, uint8_t* status, size_t n) {
  int rc = ensure_ecdsa_tables(ctx);
  if (rc) return rc;
  k_ecdsa_verify<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(msg, r, s, px, py, status, n, make_tables(ctx, true));
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

extern "C" int spg_ecdsa_verify_batch(spg_ctx* ctx, const uint64_t* msg, const uint64_t* r, const uint64_t* s,
                                      const uint64_t* pub_x, const uint64_t* pub_y_or_null, uint8_t* status, size_t n,
                                      int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && msg && r && s && pub_x && status, "spg_ecdsa_verify_batch: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  int rc = ensure_ecdsa_tables(ctx);
  if (rc) return rc;
  const uint64_t *dm = msg, *dr = r, *ds = s, *dx = pub_x, *dy = pub_y_or_null;
  uint8_t* dst = status;
  DevBuf b[5], bs;
  if (!(flags & SPG_DEVICE_PTRS)) {
    const uint64_t* src[5] = {msg, r, s, pub_x, pub_y_or_null};
    const uint64_t** dstp[5] = {&dm, &dr, &ds, &dx, &dy};
    for (int k = 0; k < 5; k++) {
      if (!src[k]) continue;
      SPG_CUDA(b[k].alloc(ctx, n * 32));
      SPG_CUDA(cudaMemcpyAsync(b[k].p, src[k], n * 32, cudaMemcpyHostToDevice, ctx->stream));
      *dstp[k] = b[k].as<uint64_t>();
    }
    SPG_CUDA(bs.alloc(ctx, n));
    dst = bs.as<uint8_t>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_ecdsa_verify<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(dm, dr, ds, dx, dy, dst, n, make_tables(ctx, true));
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
  if ((flags & SPG_NO_SYNC) && (flags & SPG_DEVICE_PTRS)) return SPG_OK;
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

extern "C" int spg_private_to_stark_key_batch(spg_ctx* ctx, const uint64_t* priv, uint64_t* pub_x, uint64_t* pub_y_or_null,
                                              uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && priv && pub_x && status, "spg_private_to_stark_key_batch: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  int rc = ensure_ecdsa_tables(ctx);
  if (rc) return rc;
  const uint64_t* dp = priv; uint64_t* dx = pub_x; uint64_t* dy = pub_y_or_null; uint8_t* dst = status;
  DevBuf bp, bx, by, bs;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(bp.alloc(ctx, n * 32)); SPG_CUDA(bx.alloc(ctx, n * 32)); SPG_CUDA(bs.alloc(ctx, n));
    if (pub_y_or_null) { SPG_CUDA(by.alloc(ctx, n * 32)); dy = by.as<uint64_t>(); }
    SPG_CUDA(cudaMemcpyAsync(bp.p, priv, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    dp = bp.as<uint64_t>(); dx = bx.as<uint64_t>(); dst = bs.as<uint8_t>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_private_to_public<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(dp, dx, dy, dst, n, make_tables(ctx, true));
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(cudaMemcpyAsync(pub_x, dx, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (pub_y_or_null) SPG_CUDA(cudaMemcpyAsync(pub_y_or_null, dy, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
    // the staging copy of the private keys goes back to the context's buffer pool: scrub it first so that key material
    // does not outlive the call in recycled device memory
    SPG_CUDA(cudaMemsetAsync(bp.p, 0, n * 32, ctx->stream));
  }
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

// ------------------------------------------------------------------ the remaining building blocks of signature.py
// get_y_coordinate (signature.py:84-96): y = the smaller square root of x^3 + x + beta.
// status: 0 ok, 1 InvalidPublicKeyError (not a quadratic residue), 2 x is not a field element.
__global__ void __launch_bounds__(128) k_get_y(const uint64_t* __restrict__ x, uint64_t* __restrict__ y,
                                               uint8_t* __restrict__ status, size_t n, EcdsaTables T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t xl[8];
  load8(x + 4 * i, xl);
  Fp res = fp_zero();
  uint8_t st = 0;
  if (spg_canon_geq_p(xl)) st = 2;
  else {
    Fp xc; for (int k = 0; k < 8; k++) xc.v[k] = xl[k];
    const Fp xm = fp_to_mont(xc);
    const Fp rhs = fp_add(fp_add(fp_mul(fp_sqr(xm), xm), xm), T.beta);
    Fp ym;
    if (fp_sqrt_min(rhs, T, &ym)) res = fp_from_mont(ym); else st = 1;
  }
  uint4* o = reinterpret_cast<uint4*>(y + 4 * i);
  o[0] = make_uint4(res.v[0], res.v[1], res.v[2], res.v[3]);
  o[1] = make_uint4(res.v[4], res.v[5], res.v[6], res.v[7]);
  status[i] = st;
}

// mimic_ec_mult_air (signature.py:176-190): m * point + shift_point with the AIR's steps.
// status: 0 ok, 1 the reference raises AssertionError (m out of (0, 2^251), an x-collision, a doubling of y = 0),
// 2 a coordinate is not a field element.
__global__ void __launch_bounds__(64) k_mimic_mult(const uint64_t* __restrict__ m, const uint64_t* __restrict__ pt,
                                                   const uint64_t* __restrict__ shift, uint64_t* __restrict__ out,
                                                   uint8_t* __restrict__ status, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t ml[8], c[4][8];
  load8(m + 4 * i, ml);
  load8(pt + 8 * i, c[0]); load8(pt + 8 * i + 4, c[1]); load8(shift + 8 * i, c[2]); load8(shift + 8 * i + 4, c[3]);
  uint8_t st = 0;
  for (int k = 0; k < 4; k++) if (spg_canon_geq_p(c[k])) st = 2;
  Fp ox = fp_zero(), oy = fp_zero();
  if (!st) {
    if (u256_is_zero(ml) || !u256_lt_2_251(ml)) st = 1;
    else {
      Fp f[4];
      for (int k = 0; k < 4; k++) { Fp t; for (int q = 0; q < 8; q++) t.v[q] = c[k][q]; f[k] = fp_to_mont(t); }
      JPoint p; p.X = f[0]; p.Y = f[1]; p.Z = fp_one();
      APoint s; s.x = f[2]; s.y = f[3];
      JPoint r;
      if (!mimic_mult_var(ml, p, s, &r)) st = 1;
      else {
        const Fp zi = fp_inv_chain(r.Z), zi2 = fp_sqr(zi);
        ox = fp_from_mont(fp_mul(r.X, zi2));
        oy = fp_from_mont(fp_mul(r.Y, fp_mul(zi2, zi)));
      }
    }
  }
  uint4* o = reinterpret_cast<uint4*>(out + 8 * i);
  o[0] = make_uint4(ox.v[0], ox.v[1], ox.v[2], ox.v[3]); o[1] = make_uint4(ox.v[4], ox.v[5], ox.v[6], ox.v[7]);
  o[2] = make_uint4(oy.v[0], oy.v[1], oy.v[2], oy.v[3]); o[3] = make_uint4(oy.v[4], oy.v[5], oy.v[6], oy.v[7]);
  status[i] = st;
}

// host-buffer helper shared by the two entry points below: n items of `in_words` / `out_words` u64 words
template <class Launch>
static int run_simple(spg_ctx* ctx, const uint64_t* const* ins, const size_t* in_words, int n_in, uint64_t* out, size_t out_words,
                      uint8_t* status, size_t n, int flags, Launch launch) {
  SPG_CUDA(cudaSetDevice(ctx->device));
  int rc = ensure_ecdsa_tables(ctx);
  if (rc) return rc;
  const uint64_t* din[4] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t* dout = out; uint8_t* dst = status;
  DevBuf bi[4], bo, bs;
  for (int k = 0; k < n_in; k++) din[k] = ins[k];
  if (!(flags & SPG_DEVICE_PTRS)) {
    for (int k = 0; k < n_in; k++) {
      SPG_CUDA(bi[k].alloc(ctx, n * in_words[k] * 8));
      SPG_CUDA(cudaMemcpyAsync(bi[k].p, ins[k], n * in_words[k] * 8, cudaMemcpyHostToDevice, ctx->stream));
      din[k] = bi[k].as<uint64_t>();
    }
    SPG_CUDA(bo.alloc(ctx, n * out_words * 8)); SPG_CUDA(bs.alloc(ctx, n));
    dout = bo.as<uint64_t>(); dst = bs.as<uint8_t>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  launch(din, dout, dst);
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(cudaMemcpyAsync(out, dout, n * out_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

extern "C" int spg_get_y_coordinate_batch(spg_ctx* ctx, const uint64_t* x, uint64_t* y, uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && x && y && status, "spg_get_y_coordinate_batch: null");
  if (n == 0) return SPG_OK;
  const uint64_t* ins[1] = {x};
  const size_t words[1] = {4};
  return run_simple(ctx, ins, words, 1, y, 4, status, n, flags, [&](const uint64_t* const* d, uint64_t* o, uint8_t* st) {
    k_get_y<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d[0], o, st, n, make_tables(ctx, true));
  });
}

extern "C" int spg_mimic_ec_mult_air_batch(spg_ctx* ctx, const uint64_t* m, const uint64_t* point_xy, const uint64_t* shift_xy,
                                           uint64_t* out_xy, uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && m && point_xy && shift_xy && out_xy && status, "spg_mimic_ec_mult_air_batch: null");
  if (n == 0) return SPG_OK;
  const uint64_t* ins[3] = {m, point_xy, shift_xy};
  const size_t words[3] = {4, 8, 8};
  return run_simple(ctx, ins, words, 3, out_xy, 8, status, n, flags, [&](const uint64_t* const* d, uint64_t* o, uint8_t* st) {
    k_mimic_mult<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(d[0], d[1], d[2], o, st, n);
  });
}

// ------------------------------------------------------------------ sign (signature.py:137-173)
// One thread derives the RFC 6979 nonce (HMAC-SHA256), multiplies the generator, and finishes the signature modulo
// n; the reference's retry loop (bumped seed) runs inside the thread.
__global__ void __launch_bounds__(64) k_ecdsa_sign(const uint64_t* __restrict__ msg, const uint64_t* __restrict__ priv,
                                                   const uint64_t* __restrict__ seed, uint64_t* __restrict__ r,
                                                   uint64_t* __restrict__ s, uint8_t* __restrict__ status, size_t n,
                                                   EcdsaTables T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t m[8], d[8], ro[8], so[8];
  load8(msg + 4 * i, m); load8(priv + 4 * i, d);
  for (int k = 0; k < 8; k++) ro[k] = so[k] = 0;
  status[i] = (uint8_t)ecdsa_sign_one(m, d, seed ? seed[i] : 0ull, T, ro, so);
  uint4* o = reinterpret_cast<uint4*>(r + 4 * i);
  o[0] = make_uint4(ro[0], ro[1], ro[2], ro[3]); o[1] = make_uint4(ro[4], ro[5], ro[6], ro[7]);
  o = reinterpret_cast<uint4*>(s + 4 * i);
  o[0] = make_uint4(so[0], so[1], so[2], so[3]); o[1] = make_uint4(so[4], so[5], so[6], so[7]);
}

extern "C" int spg_sign_batch(spg_ctx* ctx, const uint64_t* msg, const uint64_t* priv, const uint64_t* seed_or_null,
                              uint64_t* r, uint64_t* s, uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && msg && priv && r && s && status, "spg_sign_batch: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  int rc = ensure_ecdsa_tables(ctx);
  if (rc) return rc;
  const uint64_t *dm = msg, *dp = priv, *dseed = seed_or_null;
  uint64_t *dr = r, *ds = s; uint8_t* dst = status;
  DevBuf bm, bp, bseed, bo, bs;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(bm.alloc(ctx, n * 32)); SPG_CUDA(bp.alloc(ctx, n * 32)); SPG_CUDA(bo.alloc(ctx, 2 * n * 32)); SPG_CUDA(bs.alloc(ctx, n));
    SPG_CUDA(cudaMemcpyAsync(bm.p, msg, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(bp.p, priv, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (seed_or_null) {
      SPG_CUDA(bseed.alloc(ctx, n * 8));
      SPG_CUDA(cudaMemcpyAsync(bseed.p, seed_or_null, n * 8, cudaMemcpyHostToDevice, ctx->stream));
      dseed = bseed.as<uint64_t>();
    }
    dm = bm.as<uint64_t>(); dp = bp.as<uint64_t>(); dr = bo.as<uint64_t>(); ds = dr + 4 * n; dst = bs.as<uint8_t>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_ecdsa_sign<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(dm, dp, dseed, dr, ds, dst, n, make_tables(ctx, true));
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(cudaMemcpyAsync(r, dr, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(s, ds, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
    // the staging copies of the private keys and seeds go back to the context's buffer pool: scrub them first so that
    // key material does not outlive the call in recycled device memory
    SPG_CUDA(cudaMemsetAsync(bp.p, 0, n * 32, ctx->stream));
    if (bseed.p) SPG_CUDA(cudaMemsetAsync(bseed.p, 0, n * 8, ctx->stream));
  }
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

// ------------------------------------------------------------------ math_utils.py:59-100 as batched device ops
// ec_add / ec_double / ec_mult with the reference's assertions reported as statuses.  ec_mult follows the reference's
// recursion exactly -- all doublings 2^k P first (each asserting y != 0), then the additions from the highest set bit
// down to the lowest, each asserting acc.x != (2^k P).x -- so the exceptional inputs are the reference's, not those of
// some other addition chain.  The doubles a thread needs again are parked in a global scratch array [k][thread].
__device__ __forceinline__ void store_canon(uint64_t* dst, const Fp& mont) {
  const Fp c = fp_from_mont(mont);
  uint4* o = reinterpret_cast<uint4*>(dst);
  o[0] = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
  o[1] = make_uint4(c.v[4], c.v[5], c.v[6], c.v[7]);
}

__global__ void __launch_bounds__(64) k_ec_op(int op, const uint64_t* __restrict__ a_xy, const uint64_t* __restrict__ b,
                                              uint64_t* __restrict__ out_xy, uint8_t* __restrict__ status, size_t n,
                                              JPoint* __restrict__ scratch, size_t stride) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t ax[8], ay[8], b0[8] = {0}, b1[8] = {0};
  load8(a_xy + 8 * i, ax); load8(a_xy + 8 * i + 4, ay);
  if (op == 0) { load8(b + 8 * i, b0); load8(b + 8 * i + 4, b1); }
  else if (op == 2) load8(b + 4 * i, b0);
  APoint R;
  const int st = ec_op_one(op, ax, ay, b0, b1, scratch + i, stride, &R);
  store_canon(out_xy + 8 * i, R.x); store_canon(out_xy + 8 * i + 4, R.y);
  status[i] = (uint8_t)st;
}

extern "C" int spg_ec_op_batch(spg_ctx* ctx, int op, const uint64_t* a_xy, const uint64_t* b, uint64_t* out_xy,
                               uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && a_xy && out_xy && status && op >= 0 && op <= 2, "spg_ec_op_batch: arguments");
  SPG_ARG(op == 1 || b, "spg_ec_op_batch: second operand required");
  SPG_ARG(!(flags & SPG_DEVICE_PTRS), "spg_ec_op_batch: host pointers only");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  const size_t chunk = 8192, b_words = op == 0 ? 8 : 4;
  DevBuf da, db, dout, dst, dscr;
  SPG_CUDA(da.alloc(ctx, chunk * 64)); SPG_CUDA(db.alloc(ctx, chunk * 64)); SPG_CUDA(dout.alloc(ctx, chunk * 64));
  SPG_CUDA(dst.alloc(ctx, chunk));
  if (op == 2) SPG_CUDA(dscr.alloc(ctx, 255 * chunk * sizeof(JPoint)));
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (size_t c0 = 0; c0 < n; c0 += chunk) {
    const size_t nc = n - c0 < chunk ? n - c0 : chunk;
    SPG_CUDA(cudaMemcpyAsync(da.p, a_xy + 8 * c0, nc * 64, cudaMemcpyHostToDevice, ctx->stream));
    if (b) SPG_CUDA(cudaMemcpyAsync(db.p, b + b_words * c0, nc * b_words * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_ec_op<<<(unsigned)((nc + 63) / 64), 64, 0, ctx->stream>>>(op, da.as<uint64_t>(), db.as<uint64_t>(), dout.as<uint64_t>(),
                                                                dst.as<uint8_t>(), nc, dscr.as<JPoint>(), chunk);
    SPG_LAUNCH_CHECK();
    SPG_CUDA(cudaMemcpyAsync(out_xy + 8 * c0, dout.p, nc * 64, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status + c0, dst.p, nc, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

// math_utils.py:36-47 is_quad_residue / sqrt_mod over the STARK prime: y[i] = the smaller square root of a[i].
// status: 0 ok, 1 not a quadratic residue, 2 a >= p.
__global__ void __launch_bounds__(128) k_field_sqrt(const uint64_t* __restrict__ a, uint64_t* __restrict__ y,
                                                    uint8_t* __restrict__ status, size_t n, EcdsaTables T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t av[8];
  load8(a + 4 * i, av);
  Fp am, r = fp_zero();
  uint8_t st = 0;
  if (!canon_to_mont(av, &am)) st = 2;
  else if (!fp_sqrt_min(am, T, &r)) { st = 1; r = fp_zero(); }
  store_canon(y + 4 * i, r);
  status[i] = st;
}

extern "C" int spg_field_sqrt_batch(spg_ctx* ctx, const uint64_t* a, uint64_t* y, uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && a && y && status, "spg_field_sqrt_batch: null");
  if (n == 0) return SPG_OK;
  const uint64_t* ins[1] = {a};
  const size_t words[1] = {4};
  return run_simple(ctx, ins, words, 1, y, 4, status, n, flags, [&](const uint64_t* const* d, uint64_t* o, uint8_t* st) {
    k_field_sqrt<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d[0], o, st, n, make_tables(ctx, true));
  });
}

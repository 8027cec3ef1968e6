// Batched Pedersen hash on the STARK curve (BASELINE.json configs[0]).
// Reference: src/starkware/crypto/signature/signature.py:296-318 (pedersen_hash_as_point),
// fast_pedersen_hash.py:47-52 (32-byte big-endian ABI), nothing_up_my_sleeve_gen.py:85-90 (table layout).
//
// One thread per hash.  The partial sum is kept in Jacobian coordinates; at every one of the 252 steps
// per element the reference's assertion `point.x != pt.x` (signature.py:313, checked even when the
// bit is 0) is evaluated exactly as  pt.x * Z^2 != X;  the final x needs one inversion per hash.
#include <string.h>

#include "common.h"
#include "curve_params.inc"
#include "ec.cuh"

// ------------------------------------------------------------------ constant-point table (host)
int spg_curve_tables_init(spg_ctx* ctx) {
  std::vector<APoint> pts;
  pts.reserve(SPG_N_CONST_POINTS);
  auto base = [&](int i) {
    APoint a;
    a.x = spg_host_from_u64(SPG_BASE_POINTS[i][0]);
    a.y = spg_host_from_u64(SPG_BASE_POINTS[i][1]);
    return a;
  };
  pts.push_back(base(0));   // SHIFT_POINT
  pts.push_back(base(1));   // EC_GEN
  const int chain[4] = {248, 4, 248, 4};
  for (int k = 0; k < 4; k++) {
    APoint q = base(2 + k);
    for (int i = 0; i < chain[k]; i++) {
      pts.push_back(q);
      q = ec_affine_double(q);
    }
  }
  // canonicalise the lazy host results (host functions already return canonical values)
  ctx->h_const_points.assign((const Fp*)pts.data(), (const Fp*)pts.data() + 2 * pts.size());
  SPG_CUDA(cudaMalloc((void**)&ctx->const_points, pts.size() * sizeof(APoint)));
  SPG_CUDA(cudaMemcpy(ctx->const_points, pts.data(), pts.size() * sizeof(APoint), cudaMemcpyHostToDevice));
  // doubling chain of the generator for ECDSA / key derivation: G * 2^t, t < 252
  std::vector<APoint> gd(SPG_ECDSA_BITS + 1);
  APoint g = base(1);
  for (int t = 0; t <= SPG_ECDSA_BITS; t++) { gd[t] = g; g = ec_affine_double(g); }
  ctx->h_gen_doubles.assign((const Fp*)gd.data(), (const Fp*)gd.data() + 2 * gd.size());
  SPG_CUDA(cudaMalloc((void**)&ctx->gen_doubles, gd.size() * sizeof(APoint)));
  SPG_CUDA(cudaMemcpy(ctx->gen_doubles, gd.data(), gd.size() * sizeof(APoint), cudaMemcpyHostToDevice));
  return SPG_OK;
}

// ------------------------------------------------------------------ device
__device__ __forceinline__ void load_canon(const uint64_t* src, uint32_t (&x)[8]) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4 lo = s[0], hi = s[1];
  x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w; x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
}

// chain_len elements per hash unit: h = H(e0, e1); h = H(h, e2); ... (chain_len >= 1; a single element
// hashes alone like the reference's variadic pedersen_hash(x)).  elems: [n][chain_len] canonical felts.
#ifndef PEDERSEN_MIN_CTAS
#define PEDERSEN_MIN_CTAS 3       // 163 registers, no spills, 12 warps per SM instead of 8 (183 registers): the set-bit walk is bound by
                                  // the latency of its dependent addition chain -- 2^20 pairs 56.3 -> 52.5 ms; 4 (128 registers, 96 bytes of
                                  // spills) is as fast at 2^20 but 10 % slower for 1024 pairs (profiles/r2ab6_pedersen_occupancy_ab.json)
#endif
#ifndef PEDERSEN_THREADS
#define PEDERSEN_THREADS 128
#endif
__global__ void __launch_bounds__(PEDERSEN_THREADS, PEDERSEN_MIN_CTAS) k_pedersen_chain(const uint64_t* __restrict__ elems, int chain_len,
                                                        uint64_t* __restrict__ out, uint8_t* __restrict__ status,
                                                        size_t n, const APoint* __restrict__ cp,
                                                        uint64_t* __restrict__ out_y = nullptr,
                                                        const uint64_t* __restrict__ second = nullptr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // `second` != null: pairs given as two separate arrays (x in elems, y in second), chain_len == 2
  const uint64_t* e = second ? elems + i * 4 : elems + i * (size_t)chain_len * 4;
  uint32_t x[8];
  uint8_t st = 0;
  // range checks first (signature.py:307)
  for (int k = 0; k < chain_len; k++) {
    load_canon((second && k == 1) ? second + i * 4 : e + 4 * k, x);
    if (spg_canon_geq_p(x)) st = 1;
  }
  Fp res = fp_zero(), res_y = fp_zero();
  if (!st) {
    load_canon(e, x);
    int k = 1;
    do {
      PedersenAcc a;
      a.init(cp[0]);
      uint32_t y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (k < chain_len) load_canon((second && k == 1) ? second + i * 4 : e + 4 * k, y);
      if (!pedersen_absorb_elems(a, x, y, k < chain_len ? 2 : 1, cp + 2)) { st = 2; break; }
      const Fp zi = fp_inv_chain(a.p.Z), zi2 = fp_sqr(zi);
      res = fp_from_mont(fp_mul(a.p.X, zi2));
      if (out_y) res_y = fp_from_mont(fp_mul(a.p.Y, fp_mul(zi2, zi)));     // pedersen_hash_as_point (signature.py:300)
#pragma unroll
      for (int q = 0; q < 8; q++) x[q] = res.v[q];
      k++;
    } while (k < chain_len);
    if (st) { res = fp_zero(); res_y = fp_zero(); }
  }
  uint4* o = reinterpret_cast<uint4*>(out + 4 * i);
  o[0] = make_uint4(res.v[0], res.v[1], res.v[2], res.v[3]);
  o[1] = make_uint4(res.v[4], res.v[5], res.v[6], res.v[7]);
  if (out_y) {
    o = reinterpret_cast<uint4*>(out_y + 4 * i);
    o[0] = make_uint4(res_y.v[0], res_y.v[1], res_y.v[2], res_y.v[3]);
    o[1] = make_uint4(res_y.v[4], res_y.v[5], res_y.v[6], res_y.v[7]);
  }
  status[i] = st;
}

// 32-byte big-endian <-> 4 x u64 little-endian limbs (fast_pedersen_hash.py:51-52, utils.py:414-451)
__global__ void k_be32_to_limbs(const uint8_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int l = 0; l < 4; l++) {
    uint64_t v = 0;
    for (int b = 0; b < 8; b++) v = (v << 8) | in[32 * i + 8 * (3 - l) + b];
    out[4 * i + l] = v;
  }
}
__global__ void k_limbs_to_be32(const uint64_t* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int l = 0; l < 4; l++) {
    uint64_t v = in[4 * i + l];
    for (int b = 0; b < 8; b++) out[32 * i + 8 * (3 - l) + b] = (uint8_t)(v >> (56 - 8 * b));
  }
}

int spg_pedersen_chain_device(spg_ctx* ctx, const uint64_t* elems, int chain_len, uint64_t* out, uint8_t* status,
                              size_t n, uint64_t* out_y = nullptr, const uint64_t* second = nullptr) {
  if (n == 0) return SPG_OK;
  const int threads = PEDERSEN_THREADS;
  k_pedersen_chain<<<(unsigned)((n + threads - 1) / threads), threads, 0, ctx->stream>>>(
      elems, chain_len, out, status, n, (const APoint*)ctx->const_points, out_y, second);
  SPG_LAUNCH_CHECK();
  return SPG_OK;
}

// ------------------------------------------------------------------ C-ABI
extern "C" int spg_pedersen_chain_batch(spg_ctx* ctx, const uint64_t* elems, size_t chain_len, uint64_t* out,
                                        uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && elems && out && status, "spg_pedersen_chain_batch: null");
  SPG_ARG(chain_len >= 1 && chain_len <= 1024, "spg_pedersen_chain_batch: chain_len");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  const uint64_t* de = elems; uint64_t* dout = out; uint8_t* dst = status;
  DevBuf be, bo, bs;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(be.alloc(ctx, n * chain_len * 32)); SPG_CUDA(bo.alloc(ctx, n * 32)); SPG_CUDA(bs.alloc(ctx, n));
    SPG_CUDA(cudaMemcpyAsync(be.p, elems, n * chain_len * 32, cudaMemcpyHostToDevice, ctx->stream));
    de = be.as<uint64_t>(); dout = bo.as<uint64_t>(); dst = bs.as<uint8_t>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = spg_pedersen_chain_device(ctx, de, (int)chain_len, dout, dst, n);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if ((flags & SPG_NO_SYNC) && (flags & SPG_DEVICE_PTRS)) return SPG_OK;
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

extern "C" int spg_pedersen_hash2_batch(spg_ctx* ctx, const uint64_t* x, const uint64_t* y, uint64_t* out,
                                        uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && x && y && out && status, "spg_pedersen_hash2_batch: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  // the kernel reads x and y from their own arrays (contiguous uploads; no interleaving copy)
  DevBuf bx, by, bo, bs;
  const uint64_t *dx = x, *dy = y;
  uint64_t* dout = out; uint8_t* dst = status;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(bx.alloc(ctx, n * 32)); SPG_CUDA(by.alloc(ctx, n * 32)); SPG_CUDA(bo.alloc(ctx, n * 32)); SPG_CUDA(bs.alloc(ctx, n));
    SPG_CUDA(cudaMemcpyAsync(bx.p, x, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(by.p, y, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    dx = bx.as<uint64_t>(); dy = by.as<uint64_t>(); dout = bo.as<uint64_t>(); dst = bs.as<uint8_t>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = spg_pedersen_chain_device(ctx, dx, 2, dout, dst, n, nullptr, dy);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

extern "C" int spg_pedersen_hash2_batch_be32(spg_ctx* ctx, const uint8_t* x, const uint8_t* y, uint8_t* out,
                                             uint8_t* status, size_t n) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && x && y && out && status, "spg_pedersen_hash2_batch_be32: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  DevBuf bx, by, lx, ly, lo, bo, bs;
  SPG_CUDA(bx.alloc(ctx, n * 32)); SPG_CUDA(by.alloc(ctx, n * 32)); SPG_CUDA(lx.alloc(ctx, n * 32)); SPG_CUDA(ly.alloc(ctx, n * 32));
  SPG_CUDA(lo.alloc(ctx, n * 32)); SPG_CUDA(bo.alloc(ctx, n * 32)); SPG_CUDA(bs.alloc(ctx, n));
  SPG_CUDA(cudaMemcpyAsync(bx.p, x, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(by.p, y, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned blocks = (unsigned)((n + 127) / 128);
  k_be32_to_limbs<<<blocks, 128, 0, ctx->stream>>>(bx.as<uint8_t>(), lx.as<uint64_t>(), n); SPG_LAUNCH_CHECK();
  k_be32_to_limbs<<<blocks, 128, 0, ctx->stream>>>(by.as<uint8_t>(), ly.as<uint64_t>(), n); SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  int rc = spg_pedersen_hash2_batch(ctx, lx.as<uint64_t>(), ly.as<uint64_t>(), lo.as<uint64_t>(), bs.as<uint8_t>(), n,
                                    SPG_DEVICE_PTRS);
  if (rc) return rc;
  k_limbs_to_be32<<<blocks, 128, 0, ctx->stream>>>(lo.as<uint64_t>(), bo.as<uint8_t>(), n); SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaMemcpyAsync(out, bo.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(status, bs.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

// ------------------------------------------------------------------ Merkle tree with Pedersen nodes
// node = pedersen_hash(left, right) (signature.py:296-318), the node function of the StarkEx state trees that
// src/services/perpetual/cairo/state/state.cairo:155-173 updates through merkle_multi_update; the reference's
// starkware/python/merkle_tree.py:4-44 builds the update-tree hints for exactly these trees.  A level is the chain
// kernel with chain_len = 2 applied to the previous level in place of a pair list (children are adjacent).
extern "C" int spg_pedersen_merkle_tree(spg_ctx* ctx, const uint64_t* leaves, size_t n_leaves, uint64_t* root_out,
                                        uint64_t* nodes_out, uint8_t* status_out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && leaves && root_out && status_out, "spg_pedersen_merkle_tree: null");
  SPG_ARG(n_leaves >= 2 && (n_leaves & (n_leaves - 1)) == 0, "spg_pedersen_merkle_tree: n_leaves must be a power of two >= 2");
  SPG_CUDA(cudaSetDevice(ctx->device));
  // nodes: n/2 + n/4 + ... + 1 = n - 1 internal nodes, level by level from the bottom; status per node
  DevBuf bl, bn, bs;
  const uint64_t* dl = leaves;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(bl.alloc(ctx, n_leaves * 32));
    SPG_CUDA(cudaMemcpyAsync(bl.p, leaves, n_leaves * 32, cudaMemcpyHostToDevice, ctx->stream));
    dl = bl.as<uint64_t>();
  }
  SPG_CUDA(bn.alloc(ctx, (n_leaves - 1) * 32)); SPG_CUDA(bs.alloc(ctx, n_leaves - 1));
  uint64_t* dn = bn.as<uint64_t>();
  uint8_t* ds = bs.as<uint8_t>();
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  const uint64_t* prev = dl;
  size_t off = 0;
  for (size_t m = n_leaves / 2; m >= 1; m /= 2) {
    int rc = spg_pedersen_chain_device(ctx, prev, 2, dn + 4 * off, ds + off, m);
    if (rc) return rc;
    prev = dn + 4 * off;
    off += m;
    if (m == 1) break;
  }
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  std::vector<uint8_t> st(n_leaves - 1);
  SPG_CUDA(cudaMemcpyAsync(st.data(), ds, n_leaves - 1, cudaMemcpyDeviceToHost, ctx->stream));
  const cudaMemcpyKind back = (flags & SPG_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  SPG_CUDA(cudaMemcpyAsync(root_out, dn + 4 * (n_leaves - 2), 32, back, ctx->stream));
  if (nodes_out) SPG_CUDA(cudaMemcpyAsync(nodes_out, dn, (n_leaves - 1) * 32, back, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  // a failed node (input >= p can only happen at the leaves; "Unhashable input." anywhere) poisons its ancestors:
  // report the worst status seen
  uint8_t worst = 0;
  for (uint8_t v : st) if (v > worst) worst = v;
  *status_out = worst;
  return SPG_OK;
}

// pedersen_hash_as_point (signature.py:300-318): both coordinates of the hash point of 1 or 2 elements per item.
extern "C" int spg_pedersen_hash_point_batch(spg_ctx* ctx, const uint64_t* elems, size_t n_elems, uint64_t* out_x, uint64_t* out_y,
                                             uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && elems && out_x && out_y && status, "spg_pedersen_hash_point_batch: null");
  SPG_ARG(n_elems == 1 || n_elems == 2, "spg_pedersen_hash_point_batch: the constant-point table covers one or two elements");
  SPG_ARG(!(flags & SPG_DEVICE_PTRS), "spg_pedersen_hash_point_batch: host pointers only");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  DevBuf be, bx, by, bs;
  SPG_CUDA(be.alloc(ctx, n * n_elems * 32)); SPG_CUDA(bx.alloc(ctx, n * 32)); SPG_CUDA(by.alloc(ctx, n * 32)); SPG_CUDA(bs.alloc(ctx, n));
  SPG_CUDA(cudaMemcpyAsync(be.p, elems, n * n_elems * 32, cudaMemcpyHostToDevice, ctx->stream));
  int rc = spg_pedersen_chain_device(ctx, be.as<uint64_t>(), (int)n_elems, bx.as<uint64_t>(), bs.as<uint8_t>(), n, by.as<uint64_t>());
  if (rc) return rc;
  SPG_CUDA(cudaMemcpyAsync(out_x, bx.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(out_y, by.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(status, bs.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPG_OK;
}

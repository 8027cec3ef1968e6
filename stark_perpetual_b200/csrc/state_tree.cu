// State-tree work of a perpetual batch on the device: position hashing and the sparse
// Merkle multi-update of the positions / orders trees.
// Reference: src/services/perpetual/cairo/position/hash.cairo:22-74 (position_hash_assets, position_hash),
// src/services/perpetual/cairo/state/state.cairo:143-173 (hash_position_updates -> merkle_multi_update over trees of
// height <= 64), src/starkware/python/merkle_tree.py:4-44 (shape of the update tree), constants of
// src/services/perpetual/cairo/definitions/constants.cairo:10-38.  Node function: pedersen_hash (signature.py:296-318).
#include <string.h>

#include <algorithm>

#include "common.h"
#include "ec.cuh"

__device__ __forceinline__ void st_load8(const uint64_t* src, uint32_t (&x)[8]) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  const uint4 lo = s[0], hi = s[1];
  x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w; x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
}
__device__ __forceinline__ void st_store8(uint64_t* dst, const Fp& c) {
  uint4* o = reinterpret_cast<uint4*>(dst);
  o[0] = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
  o[1] = make_uint4(c.v[4], c.v[5], c.v[6], c.v[7]);
}

// ------------------------------------------------------------------ position hash (hash.cairo:22-74)
// h = 0; for every asset: h = H(h, (asset_id * 2^64 + (funding_index + 2^63)) * 2^64 + (balance + 2^63));
// h = H(h, public_key); h = H(h, (collateral_balance + 2^63) * 2^16 + n_assets).  One thread per position.
__global__ void __launch_bounds__(128) k_position_hash(const uint64_t* __restrict__ public_key, const int64_t* __restrict__ collateral,
                                                       const uint64_t* __restrict__ offsets, const uint64_t* __restrict__ asset_id,
                                                       const int64_t* __restrict__ balance, const int64_t* __restrict__ funding,
                                                       uint64_t* __restrict__ out, uint8_t* __restrict__ status, size_t n,
                                                       const APoint* __restrict__ cp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t a0 = offsets[i], a1 = offsets[i + 1];
  uint32_t pk[8], h[8], y[8];
  st_load8(public_key + 4 * i, pk);
  uint8_t st = 0;
  if (spg_canon_geq_p(pk) || a1 < a0 || a1 - a0 >= (1u << 16)) st = 1;          // N_ASSETS_UPPER_BOUND = 2^16
  for (uint64_t a = a0; a < a1 && !st; a++)
    if (asset_id[2 * a + 1] >> 56) st = 1;                                        // ASSET_ID_UPPER_BOUND = 2^120
  Fp res = fp_zero();
  if (!st) {
    const uint64_t sign = 1ull << 63;
#pragma unroll
    for (int k = 0; k < 8; k++) h[k] = 0;
    bool ok = true;
    for (uint64_t a = a0; a <= a1 + 1 && ok; a++) {
      if (a < a1) {
        const uint64_t w0 = (uint64_t)balance[a] ^ sign, w1 = (uint64_t)funding[a] ^ sign, w2 = asset_id[2 * a], w3 = asset_id[2 * a + 1];
        y[0] = (uint32_t)w0; y[1] = (uint32_t)(w0 >> 32); y[2] = (uint32_t)w1; y[3] = (uint32_t)(w1 >> 32);
        y[4] = (uint32_t)w2; y[5] = (uint32_t)(w2 >> 32); y[6] = (uint32_t)w3; y[7] = (uint32_t)(w3 >> 32);
      } else if (a == a1) {
#pragma unroll
        for (int k = 0; k < 8; k++) y[k] = pk[k];
      } else {
        const uint64_t c = (uint64_t)collateral[i] ^ sign;                        // collateral_balance - BALANCE_LOWER_BOUND
        const uint64_t w0 = (c << 16) | (a1 - a0), w1 = c >> 48;
        y[0] = (uint32_t)w0; y[1] = (uint32_t)(w0 >> 32); y[2] = (uint32_t)w1; y[3] = (uint32_t)(w1 >> 32);
        y[4] = y[5] = y[6] = y[7] = 0;
      }
      ok = pedersen_hash2_one(h, y, cp, &res);
#pragma unroll
      for (int k = 0; k < 8; k++) h[k] = res.v[k];
    }
    if (!ok) { st = 2; res = fp_zero(); }
  }
  st_store8(out + 4 * i, res);
  status[i] = st;
}

extern "C" int spg_position_hash_batch(spg_ctx* ctx, const spg_positions* pos, uint64_t* hash_out, uint8_t* status, size_t n,
                                       int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && pos && hash_out && status, "spg_position_hash_batch: null");
  SPG_ARG(pos->public_key && pos->collateral_balance && pos->asset_offsets, "spg_position_hash_batch: null field");
  SPG_ARG(!(flags & SPG_DEVICE_PTRS), "spg_position_hash_batch: host pointers only");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  const size_t total = (size_t)pos->asset_offsets[n];
  SPG_ARG(total == 0 || (pos->asset_id && pos->balance && pos->cached_funding_index), "spg_position_hash_batch: null asset arrays");
  DevBuf dpk, dcol, doff, did, dbal, dfi, dout, dst;
  SPG_CUDA(dpk.alloc(ctx, n * 32)); SPG_CUDA(dcol.alloc(ctx, n * 8)); SPG_CUDA(doff.alloc(ctx, (n + 1) * 8));
  SPG_CUDA(did.alloc(ctx, total * 16 + 16)); SPG_CUDA(dbal.alloc(ctx, total * 8 + 8)); SPG_CUDA(dfi.alloc(ctx, total * 8 + 8));
  SPG_CUDA(dout.alloc(ctx, n * 32)); SPG_CUDA(dst.alloc(ctx, n));
  SPG_CUDA(cudaMemcpyAsync(dpk.p, pos->public_key, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(dcol.p, pos->collateral_balance, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(doff.p, pos->asset_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (total) {
    SPG_CUDA(cudaMemcpyAsync(did.p, pos->asset_id, total * 16, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(dbal.p, pos->balance, total * 8, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(dfi.p, pos->cached_funding_index, total * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_position_hash<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(dpk.as<uint64_t>(), dcol.as<int64_t>(), doff.as<uint64_t>(),
                                                                         did.as<uint64_t>(), dbal.as<int64_t>(), dfi.as<int64_t>(),
                                                                         dout.as<uint64_t>(), dst.as<uint8_t>(), n,
                                                                         (const APoint*)ctx->const_points);
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(hash_out, dout.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(status, dst.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

// ------------------------------------------------------------------ sparse Merkle multi-update (state.cairo:143-173)
// The update tree (merkle_tree.py:4-29) is walked bottom-up, one kernel launch per level: the nodes of level l + 1 are
// the distinct (key >> (l + 1)); each has one or two children in the update tree, a missing child being a SIBLING the
// caller supplies (the hash of the untouched subtree, what the Cairo hints read from the preimage dictionary).  Every
// node is hashed twice: with the previous values (-> prev_root, which the caller compares with the state's root as
// merkle_multi_update does) and with the new ones (-> new_root).
struct UpdateLevel {
  std::vector<uint64_t> index;      // node indices of this level, ascending
  std::vector<int64_t> lsrc, rsrc;  // per node: child position in the level below (>= 0) or -(1 + sibling number)
};

// host: levels 1..height of the update tree and the canonical sibling order (level by level from the leaves' parents
// up, ascending node index within a level)
static void build_update_levels(unsigned height, const uint64_t* keys, size_t n, std::vector<UpdateLevel>& levels,
                                std::vector<uint8_t>* sib_level, std::vector<uint64_t>* sib_index) {
  std::vector<uint64_t> cur(keys, keys + n);
  levels.clear();
  size_t n_sib = 0;
  for (unsigned l = 0; l < height; l++) {
    UpdateLevel L;
    for (size_t k = 0; k < cur.size();) {
      const uint64_t parent = cur[k] >> 1;
      const bool pair = (k + 1 < cur.size()) && (cur[k + 1] >> 1) == parent;
      L.index.push_back(parent);
      if (pair) { L.lsrc.push_back((int64_t)k); L.rsrc.push_back((int64_t)k + 1); k += 2; continue; }
      const bool is_right = cur[k] & 1;
      if (sib_level) { sib_level->push_back((uint8_t)l); sib_index->push_back(cur[k] ^ 1); }
      if (is_right) { L.lsrc.push_back(-(int64_t)(1 + n_sib)); L.rsrc.push_back((int64_t)k); }
      else { L.lsrc.push_back((int64_t)k); L.rsrc.push_back(-(int64_t)(1 + n_sib)); }
      n_sib++;
      k++;
    }
    cur = L.index;
    levels.push_back(std::move(L));
  }
}

static int check_keys(spg_ctx* ctx, unsigned height, const uint64_t* keys, size_t n) {
  SPG_ARG(height >= 1 && height <= 64, "merkle multi-update: height must be in [1, 64]");
  SPG_ARG(n >= 1, "merkle multi-update: no updates");
  for (size_t k = 0; k < n; k++) {
    SPG_ARG(height == 64 || (keys[k] >> height) == 0, "merkle multi-update: key outside the tree");
    SPG_ARG(k == 0 || keys[k] > keys[k - 1], "merkle multi-update: keys must be strictly increasing (a squashed dict)");
  }
  return SPG_OK;
}

extern "C" int spg_merkle_update_siblings(spg_ctx* ctx, unsigned height, const uint64_t* keys, size_t n, uint8_t* level_out,
                                          uint64_t* index_out, size_t cap, size_t* count_out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && keys && count_out, "spg_merkle_update_siblings: null");
  int rc = check_keys(ctx, height, keys, n);
  if (rc) return rc;
  std::vector<UpdateLevel> levels;
  std::vector<uint8_t> sl;
  std::vector<uint64_t> si;
  build_update_levels(height, keys, n, levels, &sl, &si);
  *count_out = sl.size();
  if (level_out && index_out) {
    SPG_ARG(cap >= sl.size(), "spg_merkle_update_siblings: output too small (call with null outputs for the count)");
    memcpy(level_out, sl.data(), sl.size());
    memcpy(index_out, si.data(), si.size() * 8);
  }
  return SPG_OK;
}

// values: [2][m] felts (prev plane, new plane) of the level below; out likewise for this level
__global__ void __launch_bounds__(128) k_merkle_update_level(const uint64_t* __restrict__ below, size_t m_below,
                                                             const uint64_t* __restrict__ siblings, const int64_t* __restrict__ lsrc,
                                                             const int64_t* __restrict__ rsrc, uint64_t* __restrict__ out, size_t m,
                                                             uint32_t* __restrict__ status, const APoint* __restrict__ cp) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * m) return;
  const size_t which = t / m, j = t - which * m;                 // which: 0 previous values, 1 new values
  const int64_t ls = lsrc[j], rs = rsrc[j];
  uint32_t x[8], y[8];
  st_load8(ls >= 0 ? below + 4 * (which * m_below + (size_t)ls) : siblings + 4 * (size_t)(-ls - 1), x);
  st_load8(rs >= 0 ? below + 4 * (which * m_below + (size_t)rs) : siblings + 4 * (size_t)(-rs - 1), y);
  Fp res = fp_zero();
  if (spg_canon_geq_p(x) || spg_canon_geq_p(y)) atomicOr(status, 1u);
  else if (!pedersen_hash2_one(x, y, cp, &res)) atomicOr(status, 2u);
  st_store8(out + 4 * (which * m + j), res);
}

extern "C" int spg_merkle_multi_update(spg_ctx* ctx, unsigned height, const uint64_t* keys, const uint64_t* prev_leaves,
                                       const uint64_t* new_leaves, size_t n, const uint64_t* siblings, size_t n_siblings,
                                       uint64_t* prev_root_out, uint64_t* new_root_out, uint64_t* nodes_out, uint8_t* status_out,
                                       int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && keys && prev_leaves && new_leaves && prev_root_out && new_root_out && status_out, "spg_merkle_multi_update: null");
  SPG_ARG(!(flags & SPG_DEVICE_PTRS), "spg_merkle_multi_update: host pointers only");
  int rc = check_keys(ctx, height, keys, n);
  if (rc) return rc;
  SPG_CUDA(cudaSetDevice(ctx->device));
  std::vector<UpdateLevel> levels;
  std::vector<uint8_t> sl;
  std::vector<uint64_t> si;
  build_update_levels(height, keys, n, levels, &sl, &si);
  SPG_ARG(n_siblings == sl.size(), "spg_merkle_multi_update: wrong number of siblings (see spg_merkle_update_siblings)");
  SPG_ARG(n_siblings == 0 || siblings, "spg_merkle_multi_update: null siblings");
  size_t total_nodes = 0, max_m = n;
  for (auto& L : levels) { total_nodes += L.index.size(); max_m = std::max(max_m, L.index.size()); }
  // one upload of all levels' child tables
  std::vector<int64_t> src(2 * total_nodes);
  {
    size_t off = 0;
    for (auto& L : levels) {
      memcpy(src.data() + off, L.lsrc.data(), L.lsrc.size() * 8);
      memcpy(src.data() + total_nodes + off, L.rsrc.data(), L.rsrc.size() * 8);
      off += L.index.size();
    }
  }
  DevBuf dsrc, dsib, dval, dstat;
  SPG_CUDA(dsrc.alloc(ctx, src.size() * 8)); SPG_CUDA(dsib.alloc(ctx, n_siblings * 32 + 32));
  SPG_CUDA(dval.alloc(ctx, 2 * (n + total_nodes) * 32)); SPG_CUDA(dstat.alloc(ctx, 4));
  SPG_CUDA(cudaMemcpyAsync(dsrc.p, src.data(), src.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (n_siblings) SPG_CUDA(cudaMemcpyAsync(dsib.p, siblings, n_siblings * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemsetAsync(dstat.p, 0, 4, ctx->stream));
  // value planes of every level, back to back: level 0 = the leaves [2][n], then each level [2][m_l]
  uint64_t* v = dval.as<uint64_t>();
  SPG_CUDA(cudaMemcpyAsync(v, prev_leaves, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(v + 4 * n, new_leaves, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  size_t below_off = 0, m_below = n, node_off = 0, val_off = 2 * n;
  for (auto& L : levels) {
    const size_t m = L.index.size();
    k_merkle_update_level<<<(unsigned)((2 * m + 127) / 128), 128, 0, ctx->stream>>>(
        v + 4 * below_off, m_below, dsib.as<uint64_t>(), dsrc.as<int64_t>() + node_off, dsrc.as<int64_t>() + total_nodes + node_off,
        v + 4 * val_off, m, dstat.as<uint32_t>(), (const APoint*)ctx->const_points);
    SPG_LAUNCH_CHECK();
    below_off = val_off; m_below = m; node_off += m; val_off += 2 * m;
  }
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  // the last level is the root: [prev, new]
  SPG_CUDA(cudaMemcpyAsync(prev_root_out, v + 4 * below_off, 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(new_root_out, v + 4 * (below_off + 1), 32, cudaMemcpyDeviceToHost, ctx->stream));
  if (nodes_out) SPG_CUDA(cudaMemcpyAsync(nodes_out, v + 8 * n, 2 * total_nodes * 32, cudaMemcpyDeviceToHost, ctx->stream));
  uint32_t st = 0;
  SPG_CUDA(cudaMemcpyAsync(&st, dstat.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  *status_out = (st & 2) ? 2 : (st & 1) ? 1 : 0;
  return SPG_OK;
}

extern "C" int spg_merkle_update_node_count(spg_ctx* ctx, unsigned height, const uint64_t* keys, size_t n, size_t* level_counts_out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && keys && level_counts_out, "spg_merkle_update_node_count: null");
  int rc = check_keys(ctx, height, keys, n);
  if (rc) return rc;
  std::vector<UpdateLevel> levels;
  build_update_levels(height, keys, n, levels, nullptr, nullptr);
  for (unsigned l = 0; l < height; l++) level_counts_out[l] = levels[l].index.size();
  return SPG_OK;
}

// ------------------------------------------------------------------ right-folded hash chain (program hash, row f-2)
// compute_hash_chain of cairo-lang (starkware/cairo/common/hash_chain.py, un-vendored; call site in the reference:
// src/starkware/cairo/bootloaders/program_hash_test_utils.py:7-9 via compute_program_hash_chain):
//   h(data[0], h(data[1], h(..., h(data[n-2], data[n-1]))))        -- inherently sequential: one thread per chain.
__global__ void __launch_bounds__(32) k_hash_chain_rfold(const uint64_t* __restrict__ data, size_t len, uint64_t* __restrict__ out,
                                                         uint8_t* __restrict__ status, size_t n, const APoint* __restrict__ cp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t* d = data + 4 * i * len;
  uint32_t h[8], x[8];
  uint8_t st = 0;
  st_load8(d + 4 * (len - 1), h);
  if (spg_canon_geq_p(h)) st = 1;
  Fp res;
#pragma unroll
  for (int k = 0; k < 8; k++) res.v[k] = h[k];
  for (size_t k = len - 1; k-- > 0 && !st;) {
    st_load8(d + 4 * k, x);
    if (spg_canon_geq_p(x)) { st = 1; break; }
    if (!pedersen_hash2_one(x, h, cp, &res)) { st = 2; break; }
#pragma unroll
    for (int q = 0; q < 8; q++) h[q] = res.v[q];
  }
  if (st) res = fp_zero();
  st_store8(out + 4 * i, res);
  status[i] = st;
}

extern "C" int spg_hash_chain_rfold_batch(spg_ctx* ctx, const uint64_t* data, size_t len, uint64_t* out, uint8_t* status, size_t n,
                                          int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && data && out && status && len >= 1, "spg_hash_chain_rfold_batch: arguments");
  SPG_ARG(!(flags & SPG_DEVICE_PTRS), "spg_hash_chain_rfold_batch: host pointers only");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  DevBuf dd, dout, dst;
  SPG_CUDA(dd.alloc(ctx, n * len * 32)); SPG_CUDA(dout.alloc(ctx, n * 32)); SPG_CUDA(dst.alloc(ctx, n));
  SPG_CUDA(cudaMemcpyAsync(dd.p, data, n * len * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_hash_chain_rfold<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>(dd.as<uint64_t>(), len, dout.as<uint64_t>(), dst.as<uint8_t>(), n,
                                                                       (const APoint*)ctx->const_points);
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(out, dout.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(status, dst.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

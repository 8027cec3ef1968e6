// Merkle commitment over coset-major tables with BLAKE2s (DESIGN.md "Merkle").
//
// A table is [8 cosets][ncols][rows] field elements (Montgomery form, canonical representatives).  Leaf
// (j, i'), i' < rows/8, is the BLAKE2s of the 8 rows i' + k*rows/8 (k = 0..7) of coset j, all columns of
// a row in order -- exactly the 8 points one FRI fold-by-8 consumes, so a query opens one leaf per table.
// Elements are serialised as 32 bytes big-endian (the reference's byte order, utils.py:414-451).
// Leaf index = j * rows/8 + i'; node = BLAKE2s(left || right).
#include "blake2s.cuh"
#include "common.h"

__global__ void __launch_bounds__(128) k_merkle_leaves(const Fp* __restrict__ table, int ncols, size_t rows,
                                                       size_t n_leaves, uint32_t* __restrict__ out) {
  const size_t leaf = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n_leaves) return;
  const size_t g = rows >> 3;
  const size_t j = leaf / g, ip = leaf - j * g;
  const Fp* base = table + j * (size_t)ncols * rows + ip;
  B2s s;
  b2s_init(s);
  const int nf = 8 * ncols;
  uint32_t m[16];
  int k = 0, c = 0;
  for (int f = 0; f < nf; f += 2) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const Fp v = base[(size_t)c * rows + (size_t)k * g];
      b2s_felt_words(v, m + 8 * h);
      if (++c == ncols) { c = 0; k++; }
    }
    b2s_compress(s, m, (uint64_t)(f + 2) * 32, f + 2 == nf);
  }
  uint4* o = reinterpret_cast<uint4*>(out + 8 * leaf);
  o[0] = make_uint4(s.h[0], s.h[1], s.h[2], s.h[3]);
  o[1] = make_uint4(s.h[4], s.h[5], s.h[6], s.h[7]);
}

// one level: out[i] = H(in[2i] || in[2i+1]); digests are 8 x u32 (the BLAKE2s state words, little-endian bytes)
__global__ void __launch_bounds__(128) k_merkle_nodes(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const uint4* p = reinterpret_cast<const uint4*>(in + 16 * i);
  uint4 a = p[0], b = p[1], c = p[2], d = p[3];
  uint32_t m[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
  B2s s;
  b2s_init(s);
  b2s_compress(s, m, 64, true);
  uint4* o = reinterpret_cast<uint4*>(out + 8 * i);
  o[0] = make_uint4(s.h[0], s.h[1], s.h[2], s.h[3]);
  o[1] = make_uint4(s.h[4], s.h[5], s.h[6], s.h[7]);
}

// tree over L = n_cosets * rows / 8 leaves (the table holds n_cosets consecutive cosets; 8 = the whole table,
// fewer = the sub-tree one GPU owns): L leaf digests followed by L/2, L/4, ... 1 nodes: 2L - 1 digests of
// 32 bytes.  Level l (0 = leaves) starts at digest offset  2L - (2L >> l).
// the top of a tree in ONE launch: from a level of n_in <= 1024 digests up to the root, one CTA, levels separated by
// CTA barriers (the upper ten levels of every tree are launch-latency, not work)
__global__ void __launch_bounds__(512) k_merkle_top(uint32_t* __restrict__ tree, size_t off, int n_in) {
  __shared__ uint32_t lvl[2][1024 * 8 / 2];     // ping-pong: at most 512 digests are ever written
  const int tid = threadIdx.x;
  const uint32_t* src = tree + 8 * off;
  size_t out_off = off + (size_t)n_in;
  int cur = 0;
  for (int n = n_in; n > 1; n >>= 1) {
    const int n_out = n >> 1;
    if (tid < n_out) {
      uint32_t m[16];
#pragma unroll
      for (int k = 0; k < 16; k++) m[k] = src[16 * tid + k];
      B2s st;
      b2s_init(st);
      b2s_compress(st, m, 64, true);
#pragma unroll
      for (int k = 0; k < 8; k++) { lvl[cur][8 * tid + k] = st.h[k]; tree[8 * (out_off + tid) + k] = st.h[k]; }
    }
    __syncthreads();
    src = lvl[cur];
    cur ^= 1;
    out_off += n_out;
  }
}

int spg_merkle_build_device(spg_ctx* ctx, const Fp* table, int ncols, size_t rows, uint32_t* tree, int n_cosets) {
  SPG_ARG(rows >= 8 && (rows & (rows - 1)) == 0, "merkle: rows must be a power of two >= 8");
  SPG_ARG(n_cosets == 1 || n_cosets == 2 || n_cosets == 4 || n_cosets == 8, "merkle: n_cosets");
  const size_t n_leaves = (rows >> 3) * n_cosets;
  k_merkle_leaves<<<(unsigned)((n_leaves + 127) / 128), 128, 0, ctx->stream>>>(table, ncols, rows, n_leaves, tree);
  SPG_LAUNCH_CHECK();
  size_t off = 0, n = n_leaves;
  while (n > 1) {
    if (n <= 1024) {
      k_merkle_top<<<1, 512, 0, ctx->stream>>>(tree, off, (int)n);
      SPG_LAUNCH_CHECK();
      break;
    }
    const size_t n_out = n / 2;
    k_merkle_nodes<<<(unsigned)((n_out + 127) / 128), 128, 0, ctx->stream>>>(tree + 8 * off, tree + 8 * (off + n), n_out);
    SPG_LAUNCH_CHECK();
    off += n;
    n = n_out;
  }
  return SPG_OK;
}

// gather `count` authentication paths: for query q with leaf index idx[q], out[q][l] = sibling at level l
__global__ void k_merkle_paths(const uint32_t* __restrict__ tree, size_t n_leaves, int levels, const uint32_t* __restrict__ idx,
                               int count, uint32_t* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * levels) return;
  const int q = t / levels, l = t % levels;
  const size_t off = 2 * n_leaves - ((2 * n_leaves) >> l);
  const size_t node = ((size_t)idx[q] >> l) ^ 1;
  const uint4* src = reinterpret_cast<const uint4*>(tree + 8 * (off + node));
  uint4* dst = reinterpret_cast<uint4*>(out + 8 * (size_t)t);
  dst[0] = src[0]; dst[1] = src[1];
}

// gather the leaf CONTENTS (8 rows x ncols elements, serialised big-endian) of `count` leaves
__global__ void k_merkle_open_leaves(const Fp* __restrict__ table, int ncols, size_t rows, const uint32_t* __restrict__ idx,
                                     int count, uint32_t* __restrict__ out) {
  const int nf = 8 * ncols;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * nf) return;
  const int q = t / nf, f = t % nf;
  const int k = f / ncols, c = f % ncols;
  const size_t g = rows >> 3, leaf = idx[q];
  const size_t j = leaf / g, ip = leaf - j * g;
  const Fp v = table[(j * (size_t)ncols + c) * rows + ip + (size_t)k * g];
  uint32_t w[8];
  b2s_felt_words(v, w);
  uint4* dst = reinterpret_cast<uint4*>(out + 8 * (size_t)t);
  dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
  dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

int spg_merkle_open_device(spg_ctx* ctx, const Fp* table, int ncols, size_t rows, const uint32_t* tree,
                           const uint32_t* d_idx, int count, uint32_t* d_leaves, uint32_t* d_paths, int n_cosets) {
  const size_t n_leaves = (rows >> 3) * n_cosets;
  int levels = 0;
  while (((size_t)1 << levels) < n_leaves) levels++;
  const int nf = 8 * ncols;
  k_merkle_open_leaves<<<(count * nf + 127) / 128, 128, 0, ctx->stream>>>(table, ncols, rows, d_idx, count, d_leaves);
  SPG_LAUNCH_CHECK();
  if (levels > 0) {
    k_merkle_paths<<<(count * levels + 127) / 128, 128, 0, ctx->stream>>>(tree, n_leaves, levels, d_idx, count, d_paths);
    SPG_LAUNCH_CHECK();
  }
  return SPG_OK;
}

// ------------------------------------------------------------------ C-ABI
extern "C" int spg_merkle_commit(spg_ctx* ctx, const uint64_t* table, size_t n_cols, size_t rows, uint8_t* root32,
                                 uint8_t* tree_out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && table && root32, "spg_merkle_commit: null");
  SPG_ARG(n_cols >= 1 && n_cols <= 4096, "spg_merkle_commit: n_cols");
  SPG_ARG(rows >= 8 && (rows & (rows - 1)) == 0, "spg_merkle_commit: rows must be a power of two >= 8");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = 8 * n_cols * rows * 32, tree_bytes = (2 * rows - 1) * 32;
  const Fp* dt = (const Fp*)table;
  DevBuf bt, btree;
  uint32_t* dtree = (uint32_t*)tree_out;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(bt.alloc(ctx, bytes));
    SPG_CUDA(cudaMemcpyAsync(bt.p, table, bytes, cudaMemcpyHostToDevice, ctx->stream));
    dt = bt.as<Fp>();
  }
  if (!(flags & SPG_DEVICE_PTRS) || !tree_out) {
    SPG_CUDA(btree.alloc(ctx, tree_bytes));
    dtree = btree.as<uint32_t>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = spg_merkle_build_device(ctx, dt, (int)n_cols, rows, dtree, 8);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPG_CUDA(cudaMemcpyAsync(root32, dtree + 8 * (2 * rows - 2), 32, cudaMemcpyDeviceToHost, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS) && tree_out)
    SPG_CUDA(cudaMemcpyAsync(tree_out, dtree, tree_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

// Multi-GPU prover: one process per GPU, all stages sequenced here in C++ with NCCL called directly (SURVEY.md section 8
// row (e); DESIGN.md "Multi-GPU").  The reference has no prover; this is the sharded form of
// prove.cu and produces byte-identical proofs.
//
// Sharding:
//   * trace columns are dealt CYCLICALLY -- rank r interpolates columns {c : c mod world = r} -- so the columns of
//     round k (k world .. k world + world - 1) are contiguous and ONE in-place all-gather per round exchanges them;
//   * the rounds are pipelined in chunks of ~8 columns on two streams: while chunk c is being all-gathered (NVLink),
//     chunk c+1 is interpolated and chunk c-1 is evaluated on this rank's cosets, so only the first chunk's exchange
//     is exposed;
//   * everything after that is per coset: each rank owns 8 / world consecutive cosets for the LDE, the Merkle
//     sub-trees, the AIR (even cosets), DEEP and the first FRI fold.  Small exchanges: 32-byte sub-tree roots, the
//     four composition chunks (N felts per even coset, broadcast by its owner), the 54 out-of-domain values (dealt
//     round-robin), the first folded FRI layer (N/8 rows per coset, all-gathered once; later layers are folded
//     redundantly by every rank) and the query openings of the two sharded tables (one small all-reduce).
//   * the Fiat-Shamir channel runs on every rank's host thread from the same gathered bytes: no broadcast of challenges.
// NCCL is bound at run time (dlopen of the libnccl the process already carries -- torch's -- or SPG_NCCL_LIB): libspg has no
// link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "prove_common.h"

// ------------------------------------------------------------------ NCCL, bound at run time
struct spg_nccl_id { char internal[128]; };
typedef void* spg_nccl_comm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(spg_nccl_id*) = nullptr;
  int (*CommInitRank)(spg_nccl_comm*, int, spg_nccl_id, int) = nullptr;
  int (*CommDestroy)(spg_nccl_comm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, spg_nccl_comm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, spg_nccl_comm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, spg_nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
enum { SPG_NCCL_UINT8 = 1, SPG_NCCL_SUM = 0 };     // ncclUint8, ncclSum (stable across NCCL 2.x)

static const char* nccl_load() {
  if (g_nccl.handle) return nullptr;
  const char* names[4] = {getenv("SPG_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy the process already loaded (torch's)
  for (int k = 0; !h && k < 3; k++)
    if (names[k]) h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
  if (!h) return "libnccl.so.2 not found (import torch first, or set SPG_NCCL_LIB)";
#define SPG_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return "NCCL symbol missing: " name
  SPG_SYM(GetUniqueId, "ncclGetUniqueId"); SPG_SYM(CommInitRank, "ncclCommInitRank"); SPG_SYM(CommDestroy, "ncclCommDestroy");
  SPG_SYM(AllGather, "ncclAllGather"); SPG_SYM(Broadcast, "ncclBroadcast"); SPG_SYM(AllReduce, "ncclAllReduce");
  SPG_SYM(GroupStart, "ncclGroupStart"); SPG_SYM(GroupEnd, "ncclGroupEnd"); SPG_SYM(GetErrorString, "ncclGetErrorString");
#undef SPG_SYM
  g_nccl.handle = h;
  return nullptr;
}

#define SPG_NCCL(call)                                                                          \
  do {                                                                                          \
    int r_ = (call);                                                                            \
    if (r_ != 0) {                                                                              \
      ctx->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error"); \
      return SPG_E_CUDA;                                                                        \
    }                                                                                           \
  } while (0)

extern "C" int spg_comm_unique_id(spg_ctx* ctx, uint8_t* id_out) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && id_out, "spg_comm_unique_id: null");
  const char* e = nccl_load();
  if (e) { ctx->err = e; return SPG_E_ARG; }
  spg_nccl_id id;
  SPG_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return SPG_OK;
}

extern "C" int spg_comm_init(spg_ctx* ctx, int rank, int world, const uint8_t* id128) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && rank >= 0 && rank < world && (world == 1 || world == 2 || world == 4 || world == 8), "spg_comm_init: world must be 1, 2, 4 or 8");
  SPG_CUDA(cudaSetDevice(ctx->device));
  ctx->comm_rank = rank; ctx->comm_world = world;
  if (!ctx->comm_stream) {
    SPG_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->comm_ev) SPG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (world == 1) return SPG_OK;
  SPG_ARG(id128, "spg_comm_init: unique id required");
  const char* e = nccl_load();
  if (e) { ctx->err = e; return SPG_E_ARG; }
  spg_nccl_id id;
  memcpy(id.internal, id128, 128);
  spg_nccl_comm comm = nullptr;
  SPG_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  return SPG_OK;
}

void spg_comm_destroy(spg_ctx* ctx) {
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((spg_nccl_comm)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  ctx->comm_stream = nullptr;
  for (auto& e : ctx->comm_ev) { if (e) cudaEventDestroy(e); e = nullptr; }
}

// in-place all-gather: rank r's piece already sits at base + r * bytes
static int ag_inplace(spg_ctx* ctx, void* base, size_t bytes, cudaStream_t s) {
  if (ctx->comm_world == 1) return SPG_OK;
  SPG_NCCL(g_nccl.AllGather((const char*)base + (size_t)ctx->comm_rank * bytes, base, bytes, SPG_NCCL_UINT8,
                            (spg_nccl_comm)ctx->nccl_comm, s));
  return SPG_OK;
}
static int bcast(spg_ctx* ctx, void* buf, size_t bytes, int root, cudaStream_t s) {
  if (ctx->comm_world == 1) return SPG_OK;
  SPG_NCCL(g_nccl.Broadcast(buf, buf, bytes, SPG_NCCL_UINT8, root, (spg_nccl_comm)ctx->nccl_comm, s));
  return SPG_OK;
}

// ------------------------------------------------------------------ small kernels of the sharded path
// composition chunks: pack[jj][m][ip] <-> hev[m][jj + 4 ip]   (jj: even coset 2 jj, ip < N / 4)
__global__ void k_hev_pack(const Fp* __restrict__ hev, Fp* __restrict__ pack, unsigned log_n, int jj) {
  const size_t q = (size_t)1 << (log_n - 2), idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * q) return;
  const size_t m = idx / q, ip = idx - m * q;
  pack[(size_t)jj * 4 * q + idx] = hev[(m << log_n) + (size_t)jj + 4 * ip];
}
__global__ void k_hev_unpack(const Fp* __restrict__ pack, Fp* __restrict__ hev, unsigned log_n, int jj) {
  const size_t q = (size_t)1 << (log_n - 2), idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * q) return;
  const size_t m = idx / q, ip = idx - m * q;
  hev[(m << log_n) + (size_t)jj + 4 * ip] = pack[(size_t)jj * 4 * q + idx];
}
// zero the rows of a [count][row_words] table whose flag is 0 (query openings this rank does not own)
__global__ void k_mask_rows(uint32_t* __restrict__ buf, size_t row_words, const uint8_t* __restrict__ keep, int count) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)count * row_words) return;
  if (!keep[idx / row_words]) buf[idx] = 0u;
}

static void top_levels(const std::vector<uint8_t>& roots /*world x 32*/, std::vector<std::vector<uint8_t>>& levels) {
  levels.clear();
  levels.push_back(roots);
  while (levels.back().size() > 32) {
    const std::vector<uint8_t>& p = levels.back();
    std::vector<uint8_t> nx(p.size() / 2);
    for (size_t i = 0; i < nx.size() / 32; i++) b2s_hash_bytes(p.data() + 64 * i, 64, nx.data() + 32 * i);
    levels.push_back(nx);
  }
}

// d_cols: this rank's columns (cyclic deal), [my_cols][N] canonical, on the device; h_cols: the same in host memory
// (then uploaded chunk by chunk, overlapped with the pipeline).  Appends the proof to `proof` on every rank.
static int prove_sharded_device(spg_ctx* ctx, Fp* d_cols, const Fp* h_cols, unsigned log_n, const AirSpec& air,
                                const uint64_t* outs_canon, unsigned n_queries, std::vector<uint8_t>& proof) {
  const unsigned chain_log = air.chain_log;
  const uint64_t* x0_canon = air.x0_canon;
  const bool pedersen = air.kind == 1;
  const int n_alpha = pedersen ? SPG_MAX_ALPHA : SPG_EAIR_NALPHA;
  SPG_ARG(log_n >= 9 && log_n + SPG_LOG_BLOWUP <= SPG_UNI_LOG, "spg_prove_sharded: log_n must be in [9, 23]");
  SPG_ARG(!pedersen || 9 + chain_log <= log_n, "spg_prove_sharded: chain_log");
  SPG_ARG(n_queries >= 1 && n_queries <= 1024, "spg_prove_sharded: n_queries");
  SPG_ARG(ctx->comm_world >= 1 && ctx->comm_stream, "spg_prove_sharded: call spg_comm_init first");
  const int rank = ctx->comm_rank, world = ctx->comm_world, cs = SPG_BLOWUP / world, first = rank * cs;
  const size_t n = (size_t)1 << log_n;
  const int C = SPG_AIR_COLS;
  cudaStream_t S = ctx->stream, CS = ctx->comm_stream;
  std::vector<unsigned> log_rows = {log_n};
  while ((1u << log_rows.back()) > SPG_FRI_LAST_MAX) log_rows.push_back(log_rows.back() - 3);
  const int n_folds = (int)log_rows.size() - 1;
  const size_t lt = (size_t)cs * (n >> 3);                         // leaves of this rank's sub-trees
  int local_levels = 0;
  while (((size_t)1 << local_levels) < lt) local_levels++;
  // ---- arena
  size_t need = 0;
  auto add = [&](size_t bytes) { need += (bytes + 255) & ~(size_t)255; };
  add(C * n * 32); add((size_t)cs * C * n * 32); add(2 * lt * 32);
  add(4 * n * 32); add(4 * n * 32); add(4 * n * 32); add(4 * n * 32); add((size_t)cs * 4 * n * 32); add(2 * lt * 32);
  add(3 * (size_t)cs * n * 32); add((size_t)cs * n * 32);
  for (int l = 1; l <= n_folds; l++) { add(((size_t)8 << log_rows[l]) * 32); add(((size_t)2 << log_rows[l]) * 32); }
  add(4096 * 32); add(64 * 32 * 8); add(8 * 64 * 32);
  add((size_t)n_queries * 4 * (2 + n_folds) + 4096 + n_queries);
  add((size_t)n_queries * (8 * C + 8 * 4 + 8 * n_folds) * 32 + (size_t)n_queries * (2 + n_folds) * log_n * 32 + 4096);
  void* block;
  SPG_CUDA(spg_scratch(ctx, 3, need + 65536, &block));
  Arena ar{(char*)block, need + 65536, 0};
  Fp* t_coef = ar.get<Fp>(C * n); Fp* t_lde = ar.get<Fp>((size_t)cs * C * n); uint32_t* tree_t = ar.get<uint32_t>(16 * lt);
  Fp* cp = ar.get<Fp>(4 * n); Fp* hev = ar.get<Fp>(4 * n); Fp* hpack = ar.get<Fp>(4 * n); Fp* h_coef = ar.get<Fp>(4 * n);
  Fp* h_lde = ar.get<Fp>((size_t)cs * 4 * n); uint32_t* tree_h = ar.get<uint32_t>(16 * lt);
  Fp* inv3 = ar.get<Fp>(3 * (size_t)cs * n); Fp* layer0 = ar.get<Fp>((size_t)cs * n);
  std::vector<Fp*> layers(n_folds + 1); std::vector<uint32_t*> trees(n_folds + 1, nullptr);
  layers[0] = layer0;
  for (int l = 1; l <= n_folds; l++) { layers[l] = ar.get<Fp>((size_t)8 << log_rows[l]); trees[l] = ar.get<uint32_t>((size_t)16 << log_rows[l]); }
  Fp* d_small = ar.get<Fp>(4096);
  uint32_t* d_roots = ar.get<uint32_t>(64 * 8);
  Fp* d_oods = ar.get<Fp>(8 * 64);
  SPG_ARG(d_oods != nullptr, "arena sizing");
  spg_stage_reset(ctx);

  // ---- public input, channel (identical to prove.cu)
  AirPublic pub;
  std::vector<uint8_t> seed;
  proof.insert(proof.end(), {'S', 'P', 'G', 'P'});
  put_u32(proof, (uint32_t)air.kind); put_u32(proof, log_n); put_u32(proof, chain_log); put_u32(proof, n_queries); put_u32(proof, (uint32_t)n_folds);
  if (pedersen) {
    for (int l = 0; l < SPG_AIR_LANES; l++) { pub.x0[l] = spg_host_from_u64(x0_canon + 4 * l); pub.outs[l] = spg_host_from_u64(outs_canon + 4 * l); }
    put_u32(seed, log_n); put_u32(seed, chain_log); put_u32(seed, n_queries);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(seed, pub.x0[l]);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(seed, pub.outs[l]);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(proof, pub.x0[l]);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(proof, pub.outs[l]);
  } else {
    put_ecdsa_statement(air, log_n, n_queries, seed, proof);
  }
  Channel ch(seed);

  int rc;
  // ---- 1. interpolate (my columns) -> all-gather (rounds) -> evaluate on my cosets, pipelined in chunks
  // chunks of whole rounds.  Device-resident columns: ~8 columns per chunk (large launches).  Host columns: growing
  // chunks (1, 1, 2, 4, ... rounds) so that the first interpolation starts after ONE column per rank has crossed PCIe and
  // the copy stays just ahead of the kernels (a rank's column takes ~0.6 ms to arrive, ~0.13 ms to interpolate)
  const int n_rounds = (C + world - 1) / world, rpc = world >= 8 ? 1 : 8 / world;
  std::vector<int> cb = {0};
  if (h_cols) { for (int sz = 1, first = 1; cb.back() < n_rounds; sz = first ? 1 : std::min(2 * sz, rpc), first = 0) cb.push_back(std::min(n_rounds, cb.back() + sz)); }
  else { while (cb.back() < n_rounds) cb.push_back(std::min(n_rounds, cb.back() + rpc)); }
  const int n_chunks = (int)cb.size() - 1;
  SPG_ARG(n_chunks <= 8, "chunk count");
  auto my_rounds = [&](int k0, int k1) { int c = 0; for (int k = k0; k < k1; k++) if (k * world + rank < C) c++; return c; };
  spg_stage_begin(ctx, ST_LDE);
  // the comm stream must not start before earlier work on the compute stream is done with the buffers
  SPG_CUDA(cudaEventRecord(ctx->comm_ev[15], S));
  SPG_CUDA(cudaStreamWaitEvent(CS, ctx->comm_ev[15], 0));
  if (h_cols) {
    // host trace: this rank's columns go up chunk by chunk on the copy stream while the pipeline below already works on
    // the chunks that have arrived
    if (!ctx->copy_stream) {
      SPG_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      for (auto& e : ctx->copy_ev) SPG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      SPG_CUDA(cudaEventCreateWithFlags(&ctx->copy_gate, cudaEventDisableTiming));
    }
    SPG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->comm_ev[15], 0));
    for (int c = 0; c < n_chunks; c++) {
      const int k0 = cb[c], k1 = cb[c + 1], mine = my_rounds(k0, k1);
      if (mine)
        SPG_CUDA(cudaMemcpyAsync(d_cols + (size_t)k0 * n, h_cols + (size_t)k0 * n, (size_t)mine * n * sizeof(Fp),
                                 cudaMemcpyHostToDevice, ctx->copy_stream));
      SPG_CUDA(cudaEventRecord(ctx->copy_ev[c], ctx->copy_stream));
    }
  }
  for (int c = 0; c < n_chunks; c++) {
    const int k0 = cb[c], k1 = cb[c + 1], mine = my_rounds(k0, k1);
    if (mine) {
      if (h_cols) SPG_CUDA(cudaStreamWaitEvent(S, ctx->copy_ev[c], 0));
      if ((rc = spg_lde_coeffs_device(ctx, d_cols + (size_t)k0 * n, log_n, (size_t)mine, nullptr,
                                      t_coef + ((size_t)k0 * world + rank) * n, /*mont=*/1, (size_t)world * n))) return rc;
    }
    SPG_CUDA(cudaEventRecord(ctx->comm_ev[c], S));
    SPG_CUDA(cudaStreamWaitEvent(CS, ctx->comm_ev[c], 0));
    for (int k = k0; k < k1 && world > 1; k++) {
      const int cols_in_round = std::min(world, C - k * world);
      if (cols_in_round == world) {
        if ((rc = ag_inplace(ctx, t_coef + (size_t)k * world * n, n * sizeof(Fp), CS))) return rc;
      } else {
        SPG_NCCL(g_nccl.GroupStart());
        for (int r = 0; r < cols_in_round; r++)
          if ((rc = bcast(ctx, t_coef + ((size_t)k * world + r) * n, n * sizeof(Fp), r, CS))) return rc;
        SPG_NCCL(g_nccl.GroupEnd());
      }
    }
    SPG_CUDA(cudaEventRecord(ctx->comm_ev[8 + (c & 3)], CS));
    if (c >= 1) {   // evaluate the previous chunk while this one is in flight
      const int pk0 = cb[c - 1], pk1 = cb[c];
      const size_t col0 = (size_t)pk0 * world, ncols = std::min((size_t)C, (size_t)pk1 * world) - col0;
      SPG_CUDA(cudaStreamWaitEvent(S, ctx->comm_ev[8 + ((c - 1) & 3)], 0));
      if ((rc = spg_lde_cosets_device(ctx, t_coef + col0 * n, log_n, ncols, SPG_LOG_BLOWUP, first, cs, t_lde, C, col0))) return rc;
    }
  }
  {
    const int pk0 = cb[n_chunks - 1];
    const size_t col0 = (size_t)pk0 * world, ncols = (size_t)C - col0;
    SPG_CUDA(cudaStreamWaitEvent(S, ctx->comm_ev[8 + ((n_chunks - 1) & 3)], 0));
    if ((rc = spg_lde_cosets_device(ctx, t_coef + col0 * n, log_n, ncols, SPG_LOG_BLOWUP, first, cs, t_lde, C, col0))) return rc;
  }
  spg_stage_end(ctx, ST_LDE);

  // commitment of a sharded table: local sub-tree, 32-byte root exchange, top levels on the host
  auto commit = [&](const Fp* table, int ncols, uint32_t* tree, std::vector<std::vector<uint8_t>>& top, uint8_t root[32]) -> int {
    int r;
    if ((r = spg_merkle_build_device(ctx, table, ncols, n, tree, cs))) return r;
    SPG_CUDA(cudaMemcpyAsync(d_roots + 8 * rank, tree + 8 * (2 * lt - 2), 32, cudaMemcpyDeviceToDevice, S));
    if ((r = ag_inplace(ctx, d_roots, 32, S))) return r;
    std::vector<uint8_t> roots(32 * world);
    SPG_CUDA(spg_d2h_sync(ctx, roots.data(), d_roots, roots.size(), S));
    top_levels(roots, top);
    memcpy(root, top.back().data(), 32);
    return SPG_OK;
  };
  uint8_t root[32];
  std::vector<std::vector<uint8_t>> top_t, top_h;
  spg_stage_begin(ctx, ST_MERKLE_T);
  if ((rc = commit(t_lde, C, tree_t, top_t, root))) return rc;
  spg_stage_end(ctx, ST_MERKLE_T);
  ch.absorb(root, 32);
  put_bytes(proof, root, 32);

  // ---- 2. composition on the even cosets this rank owns; chunk split; exchange of the chunk values; chunk LDE; commitment
  const Fp alpha = ch.draw_felt();
  Fp apows[SPG_MAX_ALPHA];
  apows[0] = fp_one();
  for (int k = 1; k < n_alpha; k++) apows[k] = fp_mul(apows[k - 1], alpha);
  int jj0 = -1, n_even = 0;
  for (int jj = 0; jj < 4; jj++)
    if (2 * jj >= first && 2 * jj < first + cs) { if (jj0 < 0) jj0 = jj; n_even++; }
  spg_stage_begin(ctx, ST_AIR);
  DevBuf e_pub, e_coef, e_lde;       // kind 2: the public columns (values, coefficients on every rank; this rank's even cosets)
  if (!pedersen) {
    const size_t nb = n >> 8;
    SPG_CUDA(e_pub.alloc(ctx, 2 * nb * 32)); SPG_CUDA(e_coef.alloc(ctx, 2 * nb * 32)); SPG_CUDA(e_lde.alloc(ctx, (size_t)std::max(n_even, 1) * 2 * n * 32));
    SPG_CUDA(cudaMemcpyAsync(e_pub.p, air.msgs_canon, nb * 32, cudaMemcpyHostToDevice, S));
    SPG_CUDA(cudaMemcpyAsync(e_pub.as<Fp>() + nb, air.keys_canon, nb * 32, cudaMemcpyHostToDevice, S));
    if ((rc = spg_eair_public_device(ctx, log_n, e_pub.as<Fp>(), e_coef.as<Fp>(), e_lde.as<Fp>(), std::max(jj0, 0), n_even))) return rc;
  }
  if (n_even) {
    if (pedersen) rc = spg_air_eval_device(ctx, log_n, chain_log, t_lde, pub, apows, cp, first, jj0, n_even);
    else rc = spg_eair_eval_device(ctx, log_n, t_lde, e_lde.as<Fp>(), apows, cp, first, jj0, n_even);
    if (rc) return rc;
    if ((rc = spg_cp_split_device(ctx, log_n, cp, hev, jj0, n_even))) return rc;
  }
  if (world > 1) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    for (int e = 0; e < n_even; e++) { k_hev_pack<<<blocks, 256, 0, S>>>(hev, hpack, log_n, jj0 + e); SPG_LAUNCH_CHECK(); }
    SPG_NCCL(g_nccl.GroupStart());
    for (int jj = 0; jj < 4; jj++)
      if ((rc = bcast(ctx, hpack + (size_t)jj * n, n * sizeof(Fp), (2 * jj) / cs, S))) return rc;
    SPG_NCCL(g_nccl.GroupEnd());
    for (int jj = 0; jj < 4; jj++) {
      if (jj >= jj0 && jj < jj0 + n_even && n_even) continue;
      k_hev_unpack<<<blocks, 256, 0, S>>>(hpack, hev, log_n, jj); SPG_LAUNCH_CHECK();
    }
  }
  spg_stage_end(ctx, ST_AIR);
  spg_stage_begin(ctx, ST_HLDE);
  {
    uint64_t three[4] = {3, 0, 0, 0}, off[4];
    const Fp g = spg_host_from_u64(three);
    spg_host_to_u64(fp_inv(fp_mul(fp_mul(g, g), g)), off);
    if (world == 1) {
      if ((rc = spg_lde_coeffs_device(ctx, hev, log_n, 4, off, h_coef, /*mont=*/0))) return rc;
    } else {
      // the four chunk columns are interpolated by four (or two) different ranks and broadcast: 32 MB per column at 2^20
      // against a replicated 4-column inverse transform
      for (int m = rank; m < 4; m += world)
        if ((rc = spg_lde_coeffs_device(ctx, hev + (size_t)m * n, log_n, 1, off, h_coef + (size_t)m * n, /*mont=*/0))) return rc;
      SPG_NCCL(g_nccl.GroupStart());
      for (int m = 0; m < 4; m++)
        if ((rc = bcast(ctx, h_coef + (size_t)m * n, n * sizeof(Fp), m % world, S))) return rc;
      SPG_NCCL(g_nccl.GroupEnd());
    }
    if ((rc = spg_lde_cosets_device(ctx, h_coef, log_n, 4, SPG_LOG_BLOWUP, first, cs, h_lde))) return rc;
  }
  spg_stage_end(ctx, ST_HLDE);
  spg_stage_begin(ctx, ST_MERKLE_H);
  if ((rc = commit(h_lde, 4, tree_h, top_h, root))) return rc;
  spg_stage_end(ctx, ST_MERKLE_H);
  ch.absorb(root, 32);
  put_bytes(proof, root, 32);

  // ---- 3. out-of-domain values: the 54 evaluations are dealt round-robin (every rank holds every coefficient column)
  const Fp z = ch.draw_felt();
  const Fp wn = spg_host_root_of_unity((int)log_n);
  const Fp zw = fp_mul(z, wn), z2 = fp_sqr(z), z4 = fp_sqr(z2);
  Fp oods[SPG_N_OODS];
  {
    uint64_t three[4] = {3, 0, 0, 0};
    const Fp ginv = fp_inv(spg_host_from_u64(three));
    Fp pts[3] = {fp_mul(z, ginv), fp_mul(zw, ginv), fp_mul(z4, ginv)};
    const Fp* cols[SPG_N_OODS]; int pidx[SPG_N_OODS];
    for (int c = 0; c < C; c++) { cols[c] = t_coef + ((size_t)c << log_n); pidx[c] = 0; cols[C + c] = cols[c]; pidx[C + c] = 1; }
    for (int m = 0; m < 4; m++) { cols[2 * C + m] = h_coef + ((size_t)m << log_n); pidx[2 * C + m] = 2; }
    const Fp* mc[SPG_N_OODS]; int mp[SPG_N_OODS]; int n_mine = 0;
    for (int k = rank; k < SPG_N_OODS; k += world) { mc[n_mine] = cols[k]; mp[n_mine] = pidx[k]; n_mine++; }
    const int per_rank = (SPG_N_OODS + world - 1) / world;
    Fp mine[SPG_N_OODS];
    for (int k = 0; k < per_rank; k++) mine[k] = fp_zero();
    spg_stage_begin(ctx, ST_OODS);
    if (n_mine && (rc = spg_poly_eval_device(ctx, log_n, mc, mp, n_mine, pts, 3, mine))) return rc;
    if (world == 1) {
      for (int k = 0; k < SPG_N_OODS; k++) oods[k] = mine[k];
    } else {
      std::vector<Fp> all((size_t)per_rank * world);
      SPG_CUDA(cudaMemcpyAsync(d_oods + (size_t)rank * per_rank, mine, per_rank * sizeof(Fp), cudaMemcpyHostToDevice, S));
      if ((rc = ag_inplace(ctx, d_oods, per_rank * sizeof(Fp), S))) return rc;
      SPG_CUDA(spg_d2h_sync(ctx, all.data(), d_oods, all.size() * sizeof(Fp), S));
      for (int r = 0; r < world; r++)
        for (int i = 0, k = r; k < SPG_N_OODS; k += world, i++) oods[k] = all[(size_t)r * per_rank + i];
    }
    spg_stage_end(ctx, ST_OODS);
  }
  {
    Fp pub_z[2];
    if (!pedersen && (rc = spg_eair_public_at_host(ctx, log_n, e_coef.as<Fp>(), z, pub_z))) return rc;
    const Fp lhs = pedersen ? spg_air_composition_at_host(log_n, chain_log, pub, apows, z, oods, oods + C, ctx->h_const_points)
                            : spg_eair_composition_at_host(ctx, log_n, pub_z, apows, z, oods, oods + C);
    Fp rhs = fp_zero(), zp = fp_one();
    for (int m = 0; m < 4; m++) { rhs = fp_add(rhs, fp_mul(zp, oods[2 * C + m])); zp = fp_mul(zp, z); }
    if (!fp_eq(lhs, rhs)) { ctx->err = "trace does not satisfy the AIR (composition mismatch at the out-of-domain point)"; return SPG_E_PROOF; }
  }
  {
    std::vector<uint8_t> b;
    for (int k = 0; k < SPG_N_OODS; k++) put_fp(b, oods[k]);
    ch.absorb(b.data(), b.size());
    put_bytes(proof, b.data(), b.size());
  }
  // ---- 4. DEEP quotient on my cosets
  const Fp gamma = ch.draw_felt();
  spg_stage_begin(ctx, ST_DEEP);
  if ((rc = spg_deep_stage_device(ctx, log_n, t_lde, h_lde, first, cs, z, gamma, oods, inv3, d_small, layer0))) return rc;
  spg_stage_end(ctx, ST_DEEP);
  // ---- 5. FRI: the first fold is sharded, its output all-gathered once; the rest is folded by every rank
  std::vector<uint8_t> fri_roots;
  {
    uint64_t three[4] = {3, 0, 0, 0};
    Fp g_l = spg_host_from_u64(three);
    spg_stage_begin(ctx, ST_FRI);
    for (int l = 1; l <= n_folds; l++) {
      const Fp beta = ch.draw_felt();
      const size_t rows_l = (size_t)1 << log_rows[l];
      if (l == 1) {
        if ((rc = spg_fri_fold8_device(ctx, layers[0], log_rows[0], fp_mul(beta, fp_inv(g_l)), layers[1] + (size_t)first * rows_l, first, cs))) return rc;
        if ((rc = ag_inplace(ctx, layers[1], (size_t)cs * rows_l * sizeof(Fp), S))) return rc;
      } else {
        if ((rc = spg_fri_fold8_device(ctx, layers[l - 1], log_rows[l - 1], fp_mul(beta, fp_inv(g_l)), layers[l]))) return rc;
      }
      if ((rc = spg_merkle_build_device(ctx, layers[l], 1, rows_l, trees[l]))) return rc;
      SPG_CUDA(spg_d2h_sync(ctx, root, trees[l] + 8 * (2 * rows_l - 2), 32, S));
      ch.absorb(root, 32);
      put_bytes(fri_roots, root, 32);
      for (int k = 0; k < 3; k++) g_l = fp_sqr(g_l);
    }
    spg_stage_end(ctx, ST_FRI);
    put_bytes(proof, fri_roots.data(), fri_roots.size());
    const unsigned lr = log_rows[n_folds];
    const size_t n_last = (size_t)1 << lr;
    std::vector<Fp> vals(8 * n_last), coeffs;
    SPG_CUDA(spg_d2h_sync(ctx, vals.data(), layers[n_folds], vals.size() * sizeof(Fp), S));
    if (!spg_fri_last_layer_host(vals, lr, n_folds, coeffs)) {
      ctx->err = "trace does not satisfy the AIR (FRI last layer is not of low degree)";
      return SPG_E_PROOF;
    }
    std::vector<uint8_t> b;
    for (size_t k = 0; k < n_last; k++) put_fp(b, coeffs[k]);
    ch.absorb(b.data(), b.size());
    put_bytes(proof, b.data(), b.size());
  }
  // ---- 6. queries.  The two sharded tables (trace, chunks): every rank opens the leaves of its own cosets into a fixed
  // layout, zeroes the rows it does not own, and one all-reduce (sum) assembles them; the FRI layers are replicated.
  spg_stage_begin(ctx, ST_QUERY);
  {
    const int nt = 2 + n_folds;
    std::vector<uint32_t> idx((size_t)nt * n_queries);
    std::vector<uint8_t> keep(n_queries);
    std::vector<int> owner(n_queries);
    for (unsigned q = 0; q < n_queries; q++) {
      const uint64_t id = ch.draw_index(n);
      uint64_t j = id / (n / 8), ip = id % (n / 8);
      owner[q] = (int)(j / cs);
      keep[q] = owner[q] == rank;
      // local leaf index in this rank's sub-tree (0 for leaves of other ranks: opened as a dummy and masked out)
      idx[q] = idx[n_queries + q] = keep[q] ? (uint32_t)((j - first) * (n / 8) + ip) : 0u;
      for (int l = 1; l <= n_folds; l++) {
        const uint64_t g8 = ((uint64_t)1 << log_rows[l]) / 8;
        ip %= g8;
        idx[(size_t)(1 + l) * n_queries + q] = (uint32_t)(j * g8 + ip);
      }
    }
    uint32_t* d_idx = ar.get<uint32_t>(idx.size() + 64);
    uint8_t* d_keep = ar.get<uint8_t>(n_queries + 64);
    std::vector<size_t> leaf_words(nt), path_words(nt), leaf_off(nt), path_off(nt);
    size_t total_words = 0, shared_words = 0;
    for (int t = 0; t < nt; t++) {
      const int ncols = t == 0 ? C : (t == 1 ? 4 : 1);
      const unsigned lg = t < 2 ? (unsigned)local_levels : log_rows[t - 1];
      leaf_words[t] = (size_t)8 * ncols * 8; path_words[t] = (size_t)lg * 8;
      leaf_off[t] = total_words; total_words += leaf_words[t] * n_queries;
      path_off[t] = total_words; total_words += path_words[t] * n_queries;
      if (t == 1) shared_words = total_words;          // tables 0 and 1 are the sharded ones
    }
    uint32_t* d_open = ar.get<uint32_t>(total_words + 64);
    SPG_ARG(d_idx && d_keep && d_open, "arena sizing (queries)");
    SPG_CUDA(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, S));
    SPG_CUDA(cudaMemcpyAsync(d_keep, keep.data(), n_queries, cudaMemcpyHostToDevice, S));
    for (int t = 0; t < nt; t++) {
      const Fp* table = t == 0 ? t_lde : (t == 1 ? h_lde : layers[t - 1]);
      const uint32_t* tree = t == 0 ? tree_t : (t == 1 ? tree_h : trees[t - 1]);
      const int ncols = t == 0 ? C : (t == 1 ? 4 : 1);
      const size_t rows = (size_t)1 << (t < 2 ? log_n : log_rows[t - 1]);
      if ((rc = spg_merkle_open_device(ctx, table, ncols, rows, tree, d_idx + (size_t)t * n_queries, (int)n_queries,
                                       d_open + leaf_off[t], d_open + path_off[t], t < 2 ? cs : 8))) return rc;
      if (t < 2 && world > 1) {
        k_mask_rows<<<(unsigned)((leaf_words[t] * n_queries + 255) / 256), 256, 0, S>>>(d_open + leaf_off[t], leaf_words[t], d_keep, (int)n_queries);
        SPG_LAUNCH_CHECK();
        if (path_words[t]) {
          k_mask_rows<<<(unsigned)((path_words[t] * n_queries + 255) / 256), 256, 0, S>>>(d_open + path_off[t], path_words[t], d_keep, (int)n_queries);
          SPG_LAUNCH_CHECK();
        }
      }
    }
    if (world > 1)
      SPG_NCCL(g_nccl.AllReduce(d_open, d_open, shared_words * 4, SPG_NCCL_UINT8, SPG_NCCL_SUM, (spg_nccl_comm)ctx->nccl_comm, S));
    std::vector<uint32_t> open(total_words);
    SPG_CUDA(spg_d2h_sync(ctx, open.data(), d_open, total_words * 4, S));
    const uint8_t* ob = (const uint8_t*)open.data();
    for (unsigned q = 0; q < n_queries; q++)
      for (int t = 0; t < nt; t++) {
        put_bytes(proof, ob + 4 * (leaf_off[t] + leaf_words[t] * q), 4 * leaf_words[t]);
        put_bytes(proof, ob + 4 * (path_off[t] + path_words[t] * q), 4 * path_words[t]);
        if (t < 2) {       // the levels above the per-rank sub-trees
          const std::vector<std::vector<uint8_t>>& top = t == 0 ? top_t : top_h;
          int node = owner[q];
          for (size_t lv = 0; lv + 1 < top.size(); lv++) { put_bytes(proof, top[lv].data() + 32 * (node ^ 1), 32); node >>= 1; }
        }
      }
  }
  spg_stage_end(ctx, ST_QUERY);
  SPG_CUDA(cudaStreamSynchronize(S));
  spg_stage_collect(ctx);
  return SPG_OK;
}

static int prove_sharded_entry(spg_ctx* ctx, const uint64_t* cols_local, unsigned log_n, const AirSpec& air, const uint64_t* outs,
                               unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len, int flags) {
  SPG_ARG(log_n >= 9 && log_n <= 23, "spg_prove_sharded: log_n must be in [9, 23]");
  SPG_ARG(ctx->comm_world >= 1, "spg_prove_sharded: call spg_comm_init first");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << log_n;
  int my_cols = 0;
  for (int c = ctx->comm_rank; c < SPG_AIR_COLS; c += ctx->comm_world) my_cols++;
  SPG_ARG(my_cols == 0 || cols_local, "spg_prove_sharded: null columns");
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  Fp* d_cols = (Fp*)cols_local;
  const Fp* h_cols = nullptr;
  if (!(flags & SPG_DEVICE_PTRS)) {
    void* p;
    SPG_CUDA(spg_scratch(ctx, 2, (size_t)std::max(my_cols, 1) * n * 32, &p));
    d_cols = (Fp*)p;
    h_cols = (const Fp*)cols_local;
  }
  std::vector<uint8_t> proof;
  int rc = prove_sharded_device(ctx, d_cols, h_cols, log_n, air, outs, n_queries, proof);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  *proof_len = proof.size();
  if (proof_out) {
    SPG_ARG(proof_cap >= proof.size(), "spg_prove_sharded: proof buffer too small (call with proof_out = NULL for the size)");
    memcpy(proof_out, proof.data(), proof.size());
  }
  return SPG_OK;
}

extern "C" int spg_prove_sharded(spg_ctx* ctx, const uint64_t* cols_local, unsigned log_n, unsigned chain_log, const uint64_t* x0,
                                 const uint64_t* outs, unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len,
                                 int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && x0 && outs && proof_len, "spg_prove_sharded: null");
  AirSpec air;
  air.kind = 1; air.chain_log = chain_log; air.x0_canon = x0;
  return prove_sharded_entry(ctx, cols_local, log_n, air, outs, n_queries, proof_out, proof_cap, proof_len, flags);
}

// the sharded prover over the ECDSA-builtin AIR: cols_local as for spg_prove_sharded (this rank's columns of the trace
// spg_ecdsa_air_trace writes, dealt cyclically); msgs, key_x: the whole public input on every rank (host).  Byte-identical
// to spg_prove_ecdsa.
extern "C" int spg_prove_ecdsa_sharded(spg_ctx* ctx, const uint64_t* cols_local, unsigned log_n, const uint64_t* msgs,
                                       const uint64_t* key_x, unsigned n_queries, uint8_t* proof_out, size_t proof_cap,
                                       size_t* proof_len, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && msgs && key_x && proof_len, "spg_prove_ecdsa_sharded: null");
  SPG_ARG(log_n >= 9 && log_n <= 23, "spg_prove_ecdsa_sharded: log_n must be in [9, 23]");
  for (size_t b = 0; b < (((size_t)1 << log_n) >> 8); b++)
    for (const uint64_t* v : {msgs + 4 * b, key_x + 4 * b}) {
      uint32_t lim[8];
      for (int q = 0; q < 4; q++) { lim[2 * q] = (uint32_t)v[q]; lim[2 * q + 1] = (uint32_t)(v[q] >> 32); }
      SPG_ARG(!spg_canon_geq_p(lim), "spg_prove_ecdsa_sharded: public value >= p");
    }
  AirSpec air;
  air.kind = 2; air.msgs_canon = msgs; air.keys_canon = key_x;
  return prove_sharded_entry(ctx, cols_local, log_n, air, nullptr, n_queries, proof_out, proof_cap, proof_len, flags);
}

// STARK prover for the Pedersen hash-chain AIR: host orchestration, Fiat-Shamir channel, proof assembly.
// Stages (all on the device): LDE -> Merkle -> composition (AIR) -> chunk split + LDE -> Merkle ->
// out-of-domain evaluation -> DEEP quotient -> FRI (fold by 8 + Merkle per layer) -> query openings.
// Protocol: DESIGN.md "Protocol"; CPU restatement and verifier: oracle/stark.py.  The reference repository has
// no prover; its only description of the prover boundary is the cairo-run artefact
// list of src/starkware/cairo/lang/cairo_cmake_rules.cmake:72-110 -- here the "runner output" is the trace.
#include <stdlib.h>
#include <string.h>

#include "prove_common.h"

// d_trace: [25][N] canonical felts on the device.  Appends the proof to `proof`.
// h_trace (optional): the same trace in HOST memory; it is then uploaded into d_trace in column chunks on a
// second stream while the LDE of the chunks already on the device runs (every column's transforms depend on that
// column only), so the host->device copy of the end-to-end path hides behind the first stage.
static int prove_device(spg_ctx* ctx, const Fp* d_trace, unsigned log_n, const AirSpec& air, unsigned n_queries,
                        std::vector<uint8_t>& proof, const Fp* h_trace = nullptr) {
  const unsigned chain_log = air.chain_log;
  const uint64_t* x0_canon = air.x0_canon;
  const bool pedersen = air.kind == 1;
  const int n_alpha = pedersen ? SPG_MAX_ALPHA : SPG_EAIR_NALPHA;
  SPG_ARG(log_n >= 9 && log_n + SPG_LOG_BLOWUP <= SPG_UNI_LOG, "spg_prove: log_n must be in [9, 23]");
  SPG_ARG(!pedersen || 9 + chain_log <= log_n, "spg_prove: chain_log");
  SPG_ARG(n_queries >= 1 && n_queries <= 1024, "spg_prove: n_queries");
  const size_t n = (size_t)1 << log_n;
  const int C = SPG_AIR_COLS;
  // FRI layer sizes
  std::vector<unsigned> log_rows = {log_n};
  while ((1u << log_rows.back()) > SPG_FRI_LAST_MAX) log_rows.push_back(log_rows.back() - 3);
  const int n_folds = (int)log_rows.size() - 1;
  // ---- arena
  size_t need = 0;
  auto add = [&](size_t bytes) { need += (bytes + 255) & ~(size_t)255; };
  add(C * n * 32); add(8 * C * n * 32); add(2 * n * 32);                 // t_coef, t_lde, tree_t
  add(4 * n * 32); add(4 * n * 32); add(4 * n * 32); add(8 * 4 * n * 32); add(2 * n * 32);   // cp, hev, h_coef, h_lde, tree_h
  add(3 * 8 * n * 32); add(8 * n * 32);                                  // inv3, layer0
  for (int l = 1; l <= n_folds; l++) { add(((size_t)8 << log_rows[l]) * 32); add(((size_t)2 << log_rows[l]) * 32); }
  add(4096 * 32);                                                         // small constants
  add(n_queries * 4 * (2 + n_folds) + 4096);
  add((size_t)n_queries * (8 * C + 8 * 4 + 8 * n_folds) * 32 + (size_t)n_queries * (2 + n_folds) * log_n * 32 + 4096);
  void* block;
  SPG_CUDA(spg_scratch(ctx, 1, need + 65536, &block));
  Arena ar{(char*)block, need + 65536, 0};
  Fp* t_coef = ar.get<Fp>(C * n); Fp* t_lde = ar.get<Fp>(8 * C * n); uint32_t* tree_t = ar.get<uint32_t>(16 * n);
  Fp* cp = ar.get<Fp>(4 * n); Fp* hev = ar.get<Fp>(4 * n); Fp* h_coef = ar.get<Fp>(4 * n); Fp* h_lde = ar.get<Fp>(32 * n);
  uint32_t* tree_h = ar.get<uint32_t>(16 * n);
  Fp* inv3 = ar.get<Fp>(24 * n); Fp* layer0 = ar.get<Fp>(8 * n);
  std::vector<Fp*> layers(n_folds + 1); std::vector<uint32_t*> trees(n_folds + 1, nullptr);
  layers[0] = layer0;
  for (int l = 1; l <= n_folds; l++) { layers[l] = ar.get<Fp>((size_t)8 << log_rows[l]); trees[l] = ar.get<uint32_t>((size_t)16 << log_rows[l]); }
  Fp* d_small = ar.get<Fp>(4096);
  SPG_ARG(d_small != nullptr, "arena sizing");
  spg_stage_reset(ctx);

  // ---- public input, channel
  Fp h_last[SPG_AIR_LANES];
  // upload chunks (columns).  A column takes ~0.6 ms to arrive over PCIe and ~1.07 ms to extend, and a chunk's transforms
  // start when its last byte is there: growing chunks (1, 2, 3, 5, ...) keep the copy just ahead of the kernels, so only
  // the first column's upload is exposed (the even 1, 4, 5, 5, 5, 5 split left the GPU idle for 1.4 ms after column 0)
  int n_chunks = 6, chunk_begin[9] = {0, 1, 3, 6, 11, 18, 25, 25, 25};     // 1, 2, 3, 5, 7, 7 (measured: tools/gpu_r2_chunks.sh)
  static_assert(sizeof(ctx->copy_ev) / sizeof(ctx->copy_ev[0]) >= 8, "one event per chunk");
  if (const char* e = getenv("SPG_UPLOAD_CHUNKS")) {     // A/B knob: comma-separated chunk sizes summing to 25, at most 8
    int sizes[8], k = 0, sum = 0;
    for (const char* q = e; *q && k < 8;) { sizes[k] = atoi(q); sum += sizes[k++]; while (*q && *q != ',') q++; if (*q) q++; }
    if (sum == SPG_AIR_COLS && k >= 1) { n_chunks = k; for (int i = 0; i < k; i++) chunk_begin[i + 1] = chunk_begin[i] + sizes[i]; }
  }
  static_assert(SPG_AIR_COLS == 25, "chunk table");
  if (h_trace) {
    for (int l = 0; l < SPG_AIR_LANES; l++) h_last[l] = h_trace[((size_t)(5 * l) << log_n) + (n - 1)];   // (unused by kind 2)
    if (!ctx->copy_stream) {
      SPG_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      for (auto& e : ctx->copy_ev) SPG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      SPG_CUDA(cudaEventCreateWithFlags(&ctx->copy_gate, cudaEventDisableTiming));
    }
    // the upload buffer may still be read by the previous call's kernels
    SPG_CUDA(cudaEventRecord(ctx->copy_gate, ctx->stream));
    SPG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_gate, 0));
    for (int k = 0; k < n_chunks; k++) {
      const size_t off = (size_t)chunk_begin[k] << log_n, cols = (size_t)(chunk_begin[k + 1] - chunk_begin[k]);
      SPG_CUDA(cudaMemcpyAsync((Fp*)d_trace + off, h_trace + off, (cols << log_n) * sizeof(Fp),
                               cudaMemcpyHostToDevice, ctx->copy_stream));
      SPG_CUDA(cudaEventRecord(ctx->copy_ev[k], ctx->copy_stream));
    }
  } else if (pedersen) {
    for (int l = 0; l < SPG_AIR_LANES; l++)
      SPG_CUDA(spg_d2h_sync(ctx, &h_last[l], d_trace + ((size_t)(5 * l) << log_n) + (n - 1), sizeof(Fp), ctx->stream));
  }
  AirPublic pub;
  std::vector<uint8_t> seed;
  proof.insert(proof.end(), {'S', 'P', 'G', 'P'});
  put_u32(proof, (uint32_t)air.kind); put_u32(proof, log_n); put_u32(proof, chain_log); put_u32(proof, n_queries); put_u32(proof, (uint32_t)n_folds);
  if (pedersen) {
    for (int l = 0; l < SPG_AIR_LANES; l++) {
      pub.x0[l] = spg_host_from_u64(x0_canon + 4 * l);
      uint64_t w[4]; fp_to_u64(h_last[l], w);
      pub.outs[l] = spg_host_from_u64(w);
    }
    put_u32(seed, log_n); put_u32(seed, chain_log); put_u32(seed, n_queries);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(seed, pub.x0[l]);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(seed, pub.outs[l]);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(proof, pub.x0[l]);
    for (int l = 0; l < SPG_AIR_LANES; l++) put_fp(proof, pub.outs[l]);
  } else {                                   // oracle/stark_ecdsa.py EcdsaAir.seed / .header
    put_ecdsa_statement(air, log_n, n_queries, seed, proof);
  }
  Channel ch(seed);

  int rc;
  uint8_t root[32];
  // ---- 1. trace LDE (canonical in, Montgomery out) + commitment
  spg_stage_begin(ctx, ST_LDE);
  if (h_trace) {
    for (int k = 0; k < n_chunks; k++) {
      const size_t c0 = (size_t)chunk_begin[k], off = c0 << log_n, cols = (size_t)(chunk_begin[k + 1] - chunk_begin[k]);
      SPG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[k], 0));
      if ((rc = spg_lde_coeffs_device(ctx, d_trace + off, log_n, cols, nullptr, t_coef + off, /*mont=*/1))) return rc;
      if ((rc = spg_lde_cosets_device(ctx, t_coef + off, log_n, cols, SPG_LOG_BLOWUP, 0, SPG_BLOWUP, t_lde, C, c0))) return rc;
    }
  } else {
    if ((rc = spg_lde_device(ctx, d_trace, log_n, C, SPG_LOG_BLOWUP, nullptr, t_lde, t_coef, /*mont=*/1))) return rc;
  }
  spg_stage_end(ctx, ST_LDE);
  spg_stage_begin(ctx, ST_MERKLE_T);
  if ((rc = spg_merkle_build_device(ctx, t_lde, C, n, tree_t))) return rc;
  spg_stage_end(ctx, ST_MERKLE_T);
  SPG_CUDA(spg_d2h_sync(ctx, root, tree_t + 8 * (2 * n - 2), 32, ctx->stream));
  ch.absorb(root, 32);
  put_bytes(proof, root, 32);
  // ---- 2. composition polynomial on cosets 0, 2, 4, 6; chunk split; chunk LDE; commitment
  const Fp alpha = ch.draw_felt();
  Fp apows[SPG_MAX_ALPHA];
  apows[0] = fp_one();
  for (int k = 1; k < n_alpha; k++) apows[k] = fp_mul(apows[k - 1], alpha);
  spg_stage_begin(ctx, ST_AIR);
  DevBuf e_pub, e_coef, e_lde;       // kind 2: the public columns (values, coefficients, cosets 0 2 4 6)
  if (pedersen) rc = spg_air_eval_device(ctx, log_n, chain_log, t_lde, pub, apows, cp);
  else {
    const size_t nb = n >> 8;
    SPG_CUDA(e_pub.alloc(ctx, 2 * nb * 32)); SPG_CUDA(e_coef.alloc(ctx, 2 * nb * 32)); SPG_CUDA(e_lde.alloc(ctx, 8 * n * 32));
    SPG_CUDA(cudaMemcpyAsync(e_pub.p, air.msgs_canon, nb * 32, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(e_pub.as<Fp>() + nb, air.keys_canon, nb * 32, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = spg_eair_public_device(ctx, log_n, e_pub.as<Fp>(), e_coef.as<Fp>(), e_lde.as<Fp>()))) return rc;
    rc = spg_eair_eval_device(ctx, log_n, t_lde, e_lde.as<Fp>(), apows, cp);
  }
  if (rc) return rc;
  if ((rc = spg_cp_split_device(ctx, log_n, cp, hev))) return rc;
  spg_stage_end(ctx, ST_AIR);
  spg_stage_begin(ctx, ST_HLDE);
  {
    // chunk values live on g^4 <w_N>; evaluate on g w_{8N}^j <w_N>: offset g / g^4 = g^-3
    uint64_t three[4] = {3, 0, 0, 0}, off[4];
    const Fp g = spg_host_from_u64(three);
    spg_host_to_u64(fp_inv(fp_mul(fp_mul(g, g), g)), off);
    if ((rc = spg_lde_device(ctx, hev, log_n, 4, SPG_LOG_BLOWUP, off, h_lde, h_coef, /*mont=*/0))) return rc;
  }
  spg_stage_end(ctx, ST_HLDE);
  spg_stage_begin(ctx, ST_MERKLE_H);
  if ((rc = spg_merkle_build_device(ctx, h_lde, 4, n, tree_h))) return rc;
  spg_stage_end(ctx, ST_MERKLE_H);
  SPG_CUDA(spg_d2h_sync(ctx, root, tree_h + 8 * (2 * n - 2), 32, ctx->stream));
  ch.absorb(root, 32);
  put_bytes(proof, root, 32);
  // ---- 3. out-of-domain sampling
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
    spg_stage_begin(ctx, ST_OODS);
    if ((rc = spg_poly_eval_device(ctx, log_n, cols, pidx, SPG_N_OODS, pts, 3, oods))) return rc;
    spg_stage_end(ctx, ST_OODS);
  }
  {
    // self-check: the composition recomputed on the host from the trace values at z must equal sum z^m H_m(z^4)
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
  // ---- 4. DEEP quotient
  const Fp gamma = ch.draw_felt();
  spg_stage_begin(ctx, ST_DEEP);
  // cp (4N felts) is free after the chunk split: its first N felts hold the combined coefficient column
  if ((rc = spg_deep_stage_device(ctx, log_n, t_lde, h_lde, 0, 8, z, gamma, oods, inv3, d_small, layer0, t_coef, cp))) return rc;
  spg_stage_end(ctx, ST_DEEP);
  // ---- 5. FRI
  std::vector<uint8_t> fri_roots;
  {
    uint64_t three[4] = {3, 0, 0, 0};
    Fp g_l = spg_host_from_u64(three);
    spg_stage_begin(ctx, ST_FRI);
    for (int l = 1; l <= n_folds; l++) {
      const Fp beta = ch.draw_felt();
      if ((rc = spg_fri_fold8_device(ctx, layers[l - 1], log_rows[l - 1], fp_mul(beta, fp_inv(g_l)), layers[l]))) return rc;
      if ((rc = spg_merkle_build_device(ctx, layers[l], 1, (size_t)1 << log_rows[l], trees[l]))) return rc;
      SPG_CUDA(spg_d2h_sync(ctx, root, trees[l] + 8 * (((size_t)2 << log_rows[l]) - 2), 32, ctx->stream));
      ch.absorb(root, 32);
      put_bytes(fri_roots, root, 32);
      for (int k = 0; k < 3; k++) g_l = fp_sqr(g_l);
    }
    spg_stage_end(ctx, ST_FRI);
    put_bytes(proof, fri_roots.data(), fri_roots.size());
    // last layer -> coefficients (host: at most 512 values)
    const unsigned lr = log_rows[n_folds];
    const size_t n_last = (size_t)1 << lr;
    std::vector<Fp> vals(8 * n_last), coeffs;
    SPG_CUDA(spg_d2h_sync(ctx, vals.data(), layers[n_folds], vals.size() * sizeof(Fp), ctx->stream));
    if (!spg_fri_last_layer_host(vals, lr, n_folds, coeffs)) {
      ctx->err = "trace does not satisfy the AIR (FRI last layer is not of low degree)";
      return SPG_E_PROOF;
    }
    std::vector<uint8_t> b;
    for (size_t k = 0; k < n_last; k++) put_fp(b, coeffs[k]);
    ch.absorb(b.data(), b.size());
    put_bytes(proof, b.data(), b.size());
  }
  // ---- 6. queries
  spg_stage_begin(ctx, ST_QUERY);
  {
    const int nt = 2 + n_folds;    // tables opened per query
    std::vector<uint32_t> idx((size_t)nt * n_queries);
    for (unsigned q = 0; q < n_queries; q++) {
      const uint64_t id = ch.draw_index(n);
      idx[q] = idx[n_queries + q] = (uint32_t)id;
      uint64_t j = id / (n / 8), ip = id % (n / 8);
      for (int l = 1; l <= n_folds; l++) {
        const uint64_t g8 = ((uint64_t)1 << log_rows[l]) / 8;
        ip %= g8;
        idx[(size_t)(1 + l) * n_queries + q] = (uint32_t)(j * g8 + ip);
      }
    }
    uint32_t* d_idx = ar.get<uint32_t>(idx.size() + 64);
    // per table: leaf words and path words
    std::vector<size_t> leaf_words(nt), path_words(nt), leaf_off(nt), path_off(nt);
    size_t total_words = 0;
    for (int t = 0; t < nt; t++) {
      const int ncols = t == 0 ? C : (t == 1 ? 4 : 1);
      const unsigned lg = t < 2 ? log_n : log_rows[t - 1];
      leaf_words[t] = (size_t)8 * ncols * 8; path_words[t] = (size_t)lg * 8;
      leaf_off[t] = total_words; total_words += leaf_words[t] * n_queries;
      path_off[t] = total_words; total_words += path_words[t] * n_queries;
    }
    uint32_t* d_open = ar.get<uint32_t>(total_words + 64);
    SPG_ARG(d_idx && d_open, "arena sizing (queries)");
    SPG_CUDA(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    for (int t = 0; t < nt; t++) {
      const Fp* table = t == 0 ? t_lde : (t == 1 ? h_lde : layers[t - 1]);
      const uint32_t* tree = t == 0 ? tree_t : (t == 1 ? tree_h : trees[t - 1]);
      const int ncols = t == 0 ? C : (t == 1 ? 4 : 1);
      const size_t rows = (size_t)1 << (t < 2 ? log_n : log_rows[t - 1]);
      if ((rc = spg_merkle_open_device(ctx, table, ncols, rows, tree, d_idx + (size_t)t * n_queries, (int)n_queries,
                                       d_open + leaf_off[t], d_open + path_off[t]))) return rc;
    }
    std::vector<uint32_t> open(total_words);
    SPG_CUDA(spg_d2h_sync(ctx, open.data(), d_open, total_words * 4, ctx->stream));
    const uint8_t* ob = (const uint8_t*)open.data();
    for (unsigned q = 0; q < n_queries; q++)
      for (int t = 0; t < nt; t++) {
        put_bytes(proof, ob + 4 * (leaf_off[t] + leaf_words[t] * q), 4 * leaf_words[t]);
        put_bytes(proof, ob + 4 * (path_off[t] + path_words[t] * q), 4 * path_words[t]);
      }
  }
  spg_stage_end(ctx, ST_QUERY);
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  spg_stage_collect(ctx);
  return SPG_OK;
}

// ------------------------------------------------------------------ C-ABI
extern "C" int spg_prove(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, unsigned chain_log, const uint64_t* x0,
                         unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && trace && x0 && proof_len, "spg_prove: null");
  SPG_ARG(log_n >= 9 && log_n <= 23, "spg_prove: log_n must be in [9, 23]");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << log_n, bytes = (size_t)SPG_AIR_COLS * n * 32;
  const Fp* d_trace = (const Fp*)trace;
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  const Fp* h_trace = nullptr;
  if (!(flags & SPG_DEVICE_PTRS)) {
    void* p;
    SPG_CUDA(spg_scratch(ctx, 2, bytes, &p));
    d_trace = (const Fp*)p;
    h_trace = (const Fp*)trace;       // uploaded chunk by chunk inside prove_device, overlapped with the LDE
  }
  std::vector<uint8_t> proof;
  AirSpec air;
  air.kind = 1; air.chain_log = chain_log; air.x0_canon = x0;
  int rc = prove_device(ctx, d_trace, log_n, air, n_queries, proof, h_trace);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  *proof_len = proof.size();
  if (proof_out) {
    SPG_ARG(proof_cap >= proof.size(), "spg_prove: proof buffer too small (call with proof_out = NULL for the size)");
    memcpy(proof_out, proof.data(), proof.size());
  }
  return SPG_OK;
}

// The same protocol over the ECDSA-builtin AIR (air_ecdsa.cu): trace [25][N] canonical felts as spg_ecdsa_air_trace writes
// it; msgs, key_x = the public input, [N/256] canonical felts each (host).  Proof header VERSION = 2, followed by the
// public input; verifier: oracle/stark.py verify.
extern "C" int spg_prove_ecdsa(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, const uint64_t* msgs, const uint64_t* key_x,
                               unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && trace && msgs && key_x && proof_len, "spg_prove_ecdsa: null");
  SPG_ARG(log_n >= 9 && log_n <= 23, "spg_prove_ecdsa: log_n must be in [9, 23]");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << log_n, bytes = (size_t)SPG_AIR_COLS * n * 32;
  const Fp* d_trace = (const Fp*)trace;
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  const Fp* h_trace = nullptr;
  if (!(flags & SPG_DEVICE_PTRS)) {
    void* p;
    SPG_CUDA(spg_scratch(ctx, 2, bytes, &p));
    d_trace = (const Fp*)p;
    h_trace = (const Fp*)trace;
  }
  AirSpec air;
  air.kind = 2; air.msgs_canon = msgs; air.keys_canon = key_x;
  for (size_t b = 0; b < (n >> 8); b++)
    for (const uint64_t* v : {msgs + 4 * b, key_x + 4 * b}) {
      uint32_t lim[8];
      for (int q = 0; q < 4; q++) { lim[2 * q] = (uint32_t)v[q]; lim[2 * q + 1] = (uint32_t)(v[q] >> 32); }
      SPG_ARG(!spg_canon_geq_p(lim), "spg_prove_ecdsa: public value >= p");
    }
  std::vector<uint8_t> proof;
  int rc = prove_device(ctx, d_trace, log_n, air, n_queries, proof, h_trace);
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  *proof_len = proof.size();
  if (proof_out) {
    SPG_ARG(proof_cap >= proof.size(), "spg_prove_ecdsa: proof buffer too small (call with proof_out = NULL for the size)");
    memcpy(proof_out, proof.data(), proof.size());
  }
  return SPG_OK;
}

extern "C" int spg_pedersen_chain_trace(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const uint64_t* x0,
                                        const uint64_t* ys, uint64_t* trace_out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && x0 && ys && trace_out, "spg_pedersen_chain_trace: null");
  SPG_ARG(log_n >= 9 && log_n <= 23 && 9 + chain_log <= log_n, "spg_pedersen_chain_trace: size");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << log_n, inst = n >> 9, bytes = (size_t)SPG_AIR_COLS * n * 32;
  DevBuf dx, dy, dt, ds;
  SPG_CUDA(dx.alloc(ctx, SPG_AIR_LANES * 32)); SPG_CUDA(ds.alloc(ctx, 4));
  SPG_CUDA(cudaMemcpyAsync(dx.p, x0, SPG_AIR_LANES * 32, cudaMemcpyHostToDevice, ctx->stream));
  SPG_CUDA(cudaMemsetAsync(ds.p, 0, 4, ctx->stream));
  const Fp* dys = (const Fp*)ys;
  Fp* dtr = (Fp*)trace_out;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(dy.alloc(ctx, SPG_AIR_LANES * inst * 32)); SPG_CUDA(dt.alloc(ctx, bytes));
    SPG_CUDA(cudaMemcpyAsync(dy.p, ys, SPG_AIR_LANES * inst * 32, cudaMemcpyHostToDevice, ctx->stream));
    dys = dy.as<Fp>(); dtr = dt.as<Fp>();
  }
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = spg_pedersen_trace_device(ctx, log_n, chain_log, dx.as<Fp>(), dys, dtr, ds.as<uint8_t>());
  if (rc) return rc;
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) SPG_CUDA(cudaMemcpyAsync(trace_out, dtr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  uint32_t st = 0;
  SPG_CUDA(cudaMemcpyAsync(&st, ds.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  if (st & 1) { ctx->err = "spg_pedersen_chain_trace: an input is >= p"; return SPG_E_ARG; }
  if (st & 2) { ctx->err = "spg_pedersen_chain_trace: Unhashable input."; return SPG_E_ARG; }
  if (st & 4) { ctx->err = "spg_pedersen_chain_trace: a hash input is >= 2^251 (outside the AIR's canonical 251-bit range)"; return SPG_E_ARG; }
  return SPG_OK;
}

// composition polynomial of a trace on the cosets j = 0, 2, 4, 6 (parity / bench entry point for the AIR stage):
// trace [25][N] canonical -> cp [4][N] canonical;  alpha canonical.
extern "C" int spg_air_eval(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, unsigned chain_log, const uint64_t* x0,
                            const uint64_t* outs, const uint64_t* alpha, uint64_t* cp_out, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && trace && x0 && outs && alpha && cp_out, "spg_air_eval: null");
  SPG_ARG(log_n >= 9 && log_n <= 23 && 9 + chain_log <= log_n, "spg_air_eval: size");
  SPG_ARG(!(flags & SPG_DEVICE_PTRS), "spg_air_eval: host pointers only");
  SPG_CUDA(cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << log_n;
  DevBuf dt, dl, dc, dcp;
  SPG_CUDA(dt.alloc(ctx, SPG_AIR_COLS * n * 32)); SPG_CUDA(dl.alloc(ctx, 8 * SPG_AIR_COLS * n * 32)); SPG_CUDA(dc.alloc(ctx, SPG_AIR_COLS * n * 32));
  SPG_CUDA(dcp.alloc(ctx, 4 * n * 32));
  SPG_CUDA(cudaMemcpyAsync(dt.p, trace, SPG_AIR_COLS * n * 32, cudaMemcpyHostToDevice, ctx->stream));
  AirPublic pub;
  for (int l = 0; l < SPG_AIR_LANES; l++) { pub.x0[l] = spg_host_from_u64(x0 + 4 * l); pub.outs[l] = spg_host_from_u64(outs + 4 * l); }
  Fp apows[SPG_AIR_LANES * SPG_AIR_NCONSTR];
  const Fp a = spg_host_from_u64(alpha);
  apows[0] = fp_one();
  for (int k = 1; k < SPG_AIR_LANES * SPG_AIR_NCONSTR; k++) apows[k] = fp_mul(apows[k - 1], a);
  int rc = spg_lde_device(ctx, dt.as<Fp>(), log_n, SPG_AIR_COLS, SPG_LOG_BLOWUP, nullptr, dl.as<Fp>(), dc.as<Fp>(), 1);
  if (rc) return rc;
  spg_stage_reset(ctx);
  spg_stage_begin(ctx, ST_AIR);
  rc = spg_air_eval_device(ctx, log_n, chain_log, dl.as<Fp>(), pub, apows, dcp.as<Fp>());
  if (rc) return rc;
  spg_stage_end(ctx, ST_AIR);
  // Montgomery -> canonical for the caller
  rc = spg_from_mont_device(ctx, dcp.as<Fp>(), 4 * n);
  if (rc) return rc;
  SPG_CUDA(cudaMemcpyAsync(cp_out, dcp.p, 4 * n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  spg_stage_collect(ctx);
  ctx->last_ms = ctx->stage_ms[ST_AIR];
  return SPG_OK;
}

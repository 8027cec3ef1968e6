// CPU emulation of the CUDA NTT pass kernels: runs the very same per-thread tile functions
// (csrc/ntt.cuh) sequentially under g++ and checks them against a textbook radix-2 NTT.
// Usage: emul_ntt <log_n> <inverse> <dit> <coset_j or -1> ; prints OK / FAIL.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../stark_perpetual_b200/csrc/ntt.cuh"

static void exp_root(int log_n, uint32_t e[8]) {
  memset(e, 0, 32);
  int bits[3] = {251 - log_n, 196 - log_n, 192 - log_n};
  for (int b : bits) e[b >> 5] |= 1u << (b & 31);
}
static Fp from_u64(uint64_t x) { uint64_t w[4] = {x, 0, 0, 0}; return fp_to_mont(fp_from_u64(w)); }
static Fp root(int log_n) { uint32_t e[8]; exp_root(log_n, e); return fp_pow(from_u64(3), e, 8); }

#ifndef EMUL_LOG_WS
#define EMUL_LOG_WS 10
#endif
#ifndef EMUL_LOG_EPT
#define EMUL_LOG_EPT 2
#endif
typedef NttTile<EMUL_LOG_WS, EMUL_LOG_EPT> Tile;

static int g_ct_passes = 0;      // passes that ran on the compile-time tile (NttTileCT)
template <bool DIT>
static void run_pass(const NttPass& P, unsigned ncols) {
  const size_t ctas = ((size_t)1 << P.log_n) >> (P.log_r + P.log_g);
  std::vector<FpHalf> ws(2 * Tile::WS);
#if EMUL_LOG_EPT == 2
  // the compile-time tile kernel (k_ntt_tile) whenever the pass fits it, phase by phase as the barriers order it
  typedef NttTileCT<EMUL_LOG_WS> CT;
  if (CT::fits(P) && !getenv("EMUL_GENERIC")) {
    std::vector<FpHalf> tws(2 * CT::HW);
    g_ct_passes++;
    for (unsigned col = 0; col < ncols; col++)
      for (unsigned cta = 0; cta < ctas; cta++) {
        for (int tid = 0; tid < CT::NT; tid++) CT::stage_twiddles(P, tws.data(), tid);
        for (int tid = 0; tid < CT::NT; tid++)
          for (int j = 0; j < 4; j++) CT::load<DIT>(P, ws.data(), cta, col, CT::io_row(tid, j));
        for (int k = 0; k < CT::NS; k++)
          for (int tid = 0; tid < CT::NT; tid++) CT::step_rt<DIT>(k, ws.data(), tws.data(), tid);
        for (int tid = 0; tid < CT::NT; tid++)
          for (int j = 0; j < 4; j++) CT::store<DIT>(P, ws.data(), cta, col, CT::io_row(tid, j));
      }
    return;
  }
#endif
  for (unsigned col = 0; col < ncols; col++)
    for (unsigned cta = 0; cta < ctas; cta++) {
      for (int tid = 0; tid < Tile::NT; tid++)
        for (int j = 0; j < Tile::EPT; j++) Tile::load_one<DIT>(P, ws.data(), cta, col, j * Tile::NT + tid);
      int ns = Tile::n_steps(P);
      for (int k = 0; k < ns; k++) {
        int w, sh;
        Tile::step_geom<DIT>(P, k, &w, &sh);
        for (int tid = 0; tid < Tile::NT; tid++) Tile::step_w<DIT>(P, ws.data(), tid, w, sh);
      }
      for (int tid = 0; tid < Tile::NT; tid++)
        for (int j = 0; j < Tile::EPT; j++) Tile::store_one<DIT>(P, ws.data(), cta, col, j * Tile::NT + tid);
    }
}

int main(int argc, char** argv) {
  int log_n = argc > 1 ? atoi(argv[1]) : 12;
  int inverse = argc > 2 ? atoi(argv[2]) : 0;
  int dit = argc > 3 ? atoi(argv[3]) : 0;
  int coset_j = argc > 4 ? atoi(argv[4]) : -1;
  int pattern = argc > 5 ? atoi(argv[5]) : 0;   // 0 random, 1 every element p - 1, 2 alternating p - 1 / 0
  const size_t n = (size_t)1 << log_n;
  const unsigned ncols = 2;
  // tables
  Fp w1024 = root(SPG_TW_LOG), w1024i = fp_inv(w1024);
  const int ntw = 1 << (SPG_TW_LOG - 1);
  std::vector<Fp> twf(ntw), twi(ntw), A(8192), B(8192);
  twf[0] = twi[0] = fp_one();
  for (int i = 1; i < ntw; i++) { twf[i] = fp_mul(twf[i - 1], w1024); twi[i] = fp_mul(twi[i - 1], w1024i); }
  Fp u = root(26);
  B[0] = fp_one();
  for (int i = 1; i < 8192; i++) B[i] = fp_mul(B[i - 1], u);
  Fp u13 = fp_mul(B[8191], u);
  A[0] = fp_one();
  for (int i = 1; i < 8192; i++) A[i] = fp_mul(A[i - 1], u13);
  // data
  std::vector<Fp> x(n * ncols), y(n * ncols), ref(n * ncols);
  uint64_t s = 88172645463325252ull;
  for (auto& e : x) {
    uint64_t wv[4];
    for (int k = 0; k < 4; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; wv[k] = s; }
    wv[3] &= 0x07ffffffffffffffull;
    e = fp_from_u64(wv);
  }
  if (pattern) {
    uint64_t pm1[4] = {0, 0, 0, 0x0800000000000011ull};
    for (size_t i = 0; i < x.size(); i++) x[i] = (pattern == 1 || (i & 1)) ? fp_from_u64(pm1) : fp_zero();
  }
  // reference: O(n log n) textbook, natural in / natural out
  Fp w = root(log_n);
  if (inverse) w = fp_inv(w);
  unsigned long long coset_exp = 0;
  Fp shift = fp_one();
  if (coset_j >= 0) {   // evaluate on coset  omega_{8N}^j : coefficient k scaled by shift^k
    coset_exp = (unsigned long long)coset_j << (26 - (log_n + 3));
    shift = fp_pow_u64(root(log_n + 3), (uint64_t)coset_j);
  }
  for (unsigned c = 0; c < ncols; c++) {
    std::vector<Fp> a(n);
    // input in natural order (for dit the kernel input is the bit-reversed permutation of it)
    Fp sp = fp_one();
    for (size_t i = 0; i < n; i++) { a[i] = fp_mul(x[c * n + i], sp); sp = fp_mul(sp, shift); }
    // iterative DIF, then un-bitreverse
    for (size_t h = n / 2; h >= 1; h /= 2) {
      Fp wh = fp_pow_u64(w, n / (2 * h));
      for (size_t b = 0; b < n; b += 2 * h) {
        Fp t = fp_one();
        for (size_t j = 0; j < h; j++) {
          Fp p = a[b + j], q = a[b + j + h];
          a[b + j] = fp_add(p, q);
          a[b + j + h] = fp_mul(fp_sub(p, q), t);
          t = fp_mul(t, wh);
        }
      }
    }
    Fp ninv = inverse ? fp_inv(from_u64(n)) : fp_one();
    for (size_t i = 0; i < n; i++) ref[c * n + spg_bitrev((unsigned)i, log_n)] = fp_mul(a[i], ninv);   // natural order
  }
  // kernel input
  std::vector<Fp> in(n * ncols);
  for (unsigned c = 0; c < ncols; c++)
    for (size_t i = 0; i < n; i++) in[c * n + (dit ? spg_bitrev((unsigned)i, log_n) : i)] = x[c * n + i];
  // inverse scaling through scale_hi
  int lr, lb;
  spg_ntt_last_pass_geometry(log_n, EMUL_LOG_WS, &lr, &lb);
  std::vector<Fp> sc((size_t)1 << lb, inverse ? fp_inv(from_u64(n)) : fp_one());
  NttPass passes[8];
  int np = spg_ntt_make_passes(passes, EMUL_LOG_WS, in.data(), y.data(), log_n, n, n, inverse, dit, coset_exp, nullptr,
                               inverse ? sc.data() : nullptr, twf.data(), twi.data(), A.data(), B.data());
  for (int pi = 0; pi < np; pi++) {
    if (dit) run_pass<true>(passes[pi], ncols); else run_pass<false>(passes[pi], ncols);
  }
  size_t bad = 0;
  for (unsigned c = 0; c < ncols; c++)
    for (size_t i = 0; i < n; i++) {
      size_t pos = dit ? i : spg_bitrev((unsigned)i, log_n);
      if (!fp_eq(y[c * n + pos], ref[c * n + i])) bad++;
    }
  printf("ct_passes=%d ", g_ct_passes);
  printf("log_n=%d inverse=%d dit=%d coset=%d passes=%d : %s (%zu bad)\n", log_n, inverse, dit, coset_j, np,
         bad ? "FAIL" : "OK", bad);
  return bad ? 1 : 0;
}

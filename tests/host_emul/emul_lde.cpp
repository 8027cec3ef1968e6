// CPU emulation of the LDE pipeline (csrc/lde.cu): the same pass planner, scale tables and per-thread
// tile functions, run sequentially under g++, checked against direct Horner evaluation of the
// interpolant on every coset point.  Usage: emul_lde <log_n> <log_blowup> ; prints OK / FAIL.
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
  int log_n = argc > 1 ? atoi(argv[1]) : 6;
  int log_blowup = argc > 2 ? atoi(argv[2]) : 3;
  const size_t n = (size_t)1 << log_n, nb = (size_t)1 << log_blowup;
  const unsigned C = 2;
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
  std::vector<Fp> x(n * C), coef(n * C), out(n * C * nb);
  uint64_t s = 88172645463325252ull;
  for (auto& e : x) {
    uint64_t wv[4];
    for (int k = 0; k < 4; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; wv[k] = s; }
    wv[3] &= 0x07ffffffffffffffull;
    e = fp_from_u64(wv);   // treated as Montgomery residues; the transform is linear so any representation works
  }
  const Fp g = from_u64(3);
  std::vector<Fp> lo, hi;
  spg_lde_scale_tables(log_n, EMUL_LOG_WS, g, lo, hi);
  // direct diagonal tables exactly as lde.cu / ntt.cu build them (two-pass sizes; argv[3] = 0 disables them)
  const bool use_tables = (argc > 3 ? atoi(argv[3]) : 1) && log_n > EMUL_LOG_WS && log_n <= 2 * EMUL_LOG_WS;
  auto build_table = [&](const NttPass& P, const Fp* row_factor) {
    std::vector<Fp> t((size_t)1 << (P.log_r + P.log_s));
    for (size_t idx = 0; idx < t.size(); idx++) {
      Fp v = Tile::diag_entry(P, idx);
      if (row_factor) v = fp_mul(v, row_factor[idx & (((size_t)1 << P.log_r) - 1)]);
      t[idx] = fp_reduce(v);
    }
    return t;
  };
  NttPass passes[8];
  int np = spg_ntt_make_passes(passes, EMUL_LOG_WS, x.data(), coef.data(), log_n, n, n, 1, 0, 0, lo.data(),
                               use_tables ? nullptr : hi.data(), twf.data(), twi.data(), A.data(), B.data());
  std::vector<Fp> inv_diag;
  if (use_tables) { inv_diag = build_table(passes[0], hi.data()); passes[0].diag_table = inv_diag.data(); }
  for (int pi = 0; pi < np; pi++) run_pass<false>(passes[pi], C);
  for (size_t j = 0; j < nb; j++) {
    unsigned long long coset_exp = (unsigned long long)j << (SPG_UNI_LOG - (log_n + log_blowup));
    np = spg_ntt_make_passes(passes, EMUL_LOG_WS, coef.data(), out.data() + j * C * n, log_n, n, n, 0, 1, coset_exp, nullptr,
                             nullptr, twf.data(), twi.data(), A.data(), B.data());
    std::vector<Fp> d0, d1;
    if (use_tables) {
      d1 = build_table(passes[1], nullptr); passes[1].diag_table = d1.data();
      if (passes[0].use_diag) { d0 = build_table(passes[0], nullptr); passes[0].diag_table = d0.data(); }
    }
    for (int pi = 0; pi < np; pi++) run_pass<true>(passes[pi], C);
  }
  // reference: textbook inverse transform (O(n^2) for small n would be slow; use the iterative DIF) then Horner
  size_t bad = 0, checked = 0;
  Fp wn = root(log_n), wni = fp_inv(wn), wb = root(log_n + log_blowup), ninv = fp_inv(from_u64(n));
  for (unsigned c = 0; c < C; c++) {
    std::vector<Fp> a(x.begin() + c * n, x.begin() + (c + 1) * n), ck(n);
    for (size_t h = n / 2; h >= 1; h /= 2) {
      Fp wh = fp_pow_u64(wni, n / (2 * h));
      for (size_t b = 0; b < n; b += 2 * h) {
        Fp t = fp_one();
        for (size_t k = 0; k < h; k++) {
          Fp p = a[b + k], q = a[b + k + h];
          a[b + k] = fp_add(p, q);
          a[b + k + h] = fp_mul(fp_sub(p, q), t);
          t = fp_mul(t, wh);
        }
      }
    }
    for (size_t i = 0; i < n; i++) ck[spg_bitrev((unsigned)i, log_n)] = fp_mul(a[i], ninv);
    // sample points (all of them for small n)
    size_t step = n * nb > 4096 ? (n * nb) / 509 : 1;
    for (size_t q = 0; q < n * nb; q += step) {
      size_t j = q / n, i = q % n;
      Fp pt = fp_mul(fp_mul(g, fp_pow_u64(wb, j)), fp_pow_u64(wn, i));
      Fp acc = fp_zero();
      for (size_t k = n; k-- > 0;) acc = fp_add(fp_mul(acc, pt), ck[k]);
      if (!fp_eq(acc, out[(j * C + c) * n + i])) bad++;
      checked++;
    }
  }
  printf("ct_passes=%d ", g_ct_passes);
  printf("lde log_n=%d log_blowup=%d : %s (%zu bad of %zu)\n", log_n, log_blowup, bad ? "FAIL" : "OK", bad, checked);
  return bad ? 1 : 0;
}

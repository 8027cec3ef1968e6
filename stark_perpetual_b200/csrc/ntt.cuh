// Tiled number-theoretic transform over the STARK prime.
//
// No reference symbol exists for this stage: the field and generator come from
// signature.py:41-42; the transform conventions are this repo's own (DESIGN.md "NTT").
//
// A size-2^n transform is a sequence of PASSES.  A pass views the column as [B][R][S]
// (element = b*R*S + r*S + c) and runs, for every (b, c), a size-R transform over r entirely in
// shared memory; consecutive passes are glued by the classic "four-step" diagonal twiddle
// w_{RS}^(bitrev(r) * c).  Two butterfly networks are provided:
//   DIF (Gentleman-Sande)  natural order in  -> bit-reversed order out
//   DIT (Cooley-Tukey)     bit-reversed in   -> natural order out
// so LDE (inverse DIF, then per-coset forward DIT) never needs a permutation pass.
//
// One CTA owns a WORKSPACE of 2^LOG_WS field elements in shared memory: R rows x G columns when
// S > 1 (G*32-byte global segments), or G whole contiguous tiles when S == 1.  Each thread holds 8
// elements in registers per STEP and does up to 3 butterfly stages (radix-8) between exchanges.
//
// Everything is written as per-thread functions of (tid, step) so the same code runs under the
// CPU emulation harness in tests/host_emul (g++, no GPU).
#pragma once
#include "fp.cuh"

#define SPG_TW_LOG 11          // intra-tile twiddle table: omega_2048^e, e < 1024 (passes of up to 11 bits)
#define SPG_UNI_LOG 26         // universal two-level table: omega_{2^26}^e = uniA[e >> 13] * uniB[e & 8191]
#define SPG_UNI_HALF 13

struct NttPass {
  const Fp* in;                // column 0
  Fp* out;
  unsigned long long in_col_stride, out_col_stride;   // elements between columns
  int log_n;                   // whole transform size
  int log_r, log_s;            // this pass: rows R, row stride S   (B = N / (R S))
  int log_g;                   // S > 1: columns per CTA;  S == 1: tiles per CTA
  int inverse;                 // use omega^-1
  const Fp* tw;                // omega_2048^e  (forward) or omega_2048^-e (inverse), 1024 entries
  const Fp* uniA;              // universal table, high part (8192 entries)
  const Fp* uniB;              // universal table, low part (8192 entries)
  // diagonal twiddle  omega_{2^26}^( +- bitrev_R(r) * (c * ec + e0) )  applied AFTER a DIF pass /
  // BEFORE a DIT pass.  use_diag = 0 disables it (last DIF pass / first DIT pass without coset).
  int use_diag;
  unsigned long long ec, e0;
  // optional element-wise scaling tables (Montgomery form), DIF: applied at the end of the pass,
  // DIT: applied at the start:  x *= lo[r] * hi[b]   (either may be null)
  const Fp* scale_lo;          // R entries
  const Fp* scale_hi;          // B entries
  int final_pass;              // last pass of the transform: store canonical values (otherwise any lazy representative)
  // optional direct table of the diagonal factor: diag_table[(c << log_r) | r]   (ROW order, so that a warp's consecutive
  // rows read consecutive entries -- coalesced)
  // = omega_{2^26}^(+- bitrev_R(r) * (c * ec + e0)) [* an extra per-row factor folded in by the LDE, see lde.cu]; saves the
  // two-level lookup's multiplication (null: use uniA/uniB).  The contiguous pass of a coset transform has c = 0: its
  // table holds just 2^log_r entries.
  const Fp* diag_table;
};

// One half (16 bytes) of a field element in the shared-memory workspace.  The workspace is PLANAR -- low halves
// in ws[0 .. WS), high halves in ws[WS .. 2 WS) -- and the slot index is XOR-swizzled, so that a warp's 16-byte
// accesses hit 8 distinct bank groups per quarter-warp for every power-of-two stride the butterflies use.
struct alignas(16) FpHalf {
  uint32_t v[4];
};

SPG_HD unsigned spg_bitrev(unsigned x, int bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
  unsigned r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

// LOG_EPT: log2 of the elements a thread holds per step (3: radix-8 steps, 2: radix-4 steps -- half the live
// registers, twice the threads per workspace).
template <int LOG_WS, int LOG_EPT = 3>
struct NttTile {
  static constexpr int WS = 1 << LOG_WS;
  static constexpr int EPT = 1 << LOG_EPT;
  static constexpr int NT = WS / EPT;

  // workspace slot of (row r, column/tile g)
  static SPG_HD int slot(const NttPass& P, int r, int g) {
    return P.log_s ? ((r << P.log_g) | g) : ((g << P.log_r) | r);
  }
  // swizzled position of a slot: the low 3 bits are XORed with the fold of all higher 3-bit groups, plus slot bit 4
  // into bit 2.  Over GF(2) this makes the 8 slots of a quarter-warp distinct modulo 8 for every access pattern of
  // the kernel: consecutive slots, any power-of-two stride >= 8, stride 4 (radix-4 step at shift 0) and the mixed
  // pattern of the radix-4 step at shift 2 (four consecutive slots, then +16), which the plain fold maps 2-way.
  static SPG_HD int swz(int s) {
    int u = s >> 3;
    return s ^ ((u ^ (u >> 3) ^ (u >> 6) ^ (u >> 9)) & 7) ^ ((u & 2) << 1);
  }
  static SPG_HD Fp ws_load(const FpHalf* ws, int s) {
    const int i = swz(s);
    const FpHalf lo = ws[i], hi = ws[WS + i];
    Fp x;
#pragma unroll
    for (int k = 0; k < 4; k++) { x.v[k] = lo.v[k]; x.v[4 + k] = hi.v[k]; }
    return x;
  }
  static SPG_HD void ws_store(FpHalf* ws, int s, const Fp& x) {
    const int i = swz(s);
    FpHalf lo, hi;
#pragma unroll
    for (int k = 0; k < 4; k++) { lo.v[k] = x.v[k]; hi.v[k] = x.v[4 + k]; }
    ws[i] = lo; ws[WS + i] = hi;
  }

  // global element offset (within a column) of workspace-linear index idx for CTA `cta`
  static SPG_HD unsigned long long gaddr(const NttPass& P, unsigned cta, int idx, int* r_out, int* g_out,
                                         unsigned* b_out, unsigned* c_out) {
    if (P.log_s) {
      int g = idx & ((1 << P.log_g) - 1), r = idx >> P.log_g;
      unsigned ctas_per_block = 1u << (P.log_s - P.log_g);
      unsigned b = cta / ctas_per_block, c = ((cta % ctas_per_block) << P.log_g) + g;
      *r_out = r; *g_out = g; *b_out = b; *c_out = c;
      return ((unsigned long long)b << (P.log_r + P.log_s)) + ((unsigned long long)r << P.log_s) + c;
    } else {
      int r = idx & ((1 << P.log_r) - 1), g = idx >> P.log_r;
      unsigned b = (cta << P.log_g) + g;
      *r_out = r; *g_out = g; *b_out = b; *c_out = 0;
      return ((unsigned long long)b << P.log_r) + r;
    }
  }

  // omega_{2^26}^E through the two-level table
  static SPG_HD Fp uni_pow(const NttPass& P, unsigned long long E) {
    E &= (1ull << SPG_UNI_LOG) - 1;
    unsigned hi = (unsigned)(E >> SPG_UNI_HALF), lo = (unsigned)(E & ((1u << SPG_UNI_HALF) - 1));
    if (lo == 0) return P.uniA[hi];
    if (hi == 0) return P.uniB[lo];
    return fp_mul_lazy(P.uniA[hi], P.uniB[lo]);
  }

  // entry idx = (c << log_r) | r of the direct diagonal table of pass P
  static SPG_HD Fp diag_entry(const NttPass& P, unsigned long long idx) {
    const unsigned long long k = spg_bitrev((unsigned)(idx & ((1ull << P.log_r) - 1)), P.log_r), c = idx >> P.log_r;
    unsigned long long E = k * (c * P.ec + P.e0);
    if (P.inverse) E = (0ull - E);
    return uni_pow(P, E);
  }

  // factor applied to element (b, r, c): diagonal twiddle and optional scale tables.  Whenever the pass has a
  // factor at all the element IS multiplied (by one if need be): a multiplication is what brings a lazy value
  // back below 2p, which the butterflies of the next pass rely on.
  static SPG_HD Fp apply_factors(const NttPass& P, Fp x, unsigned b, int r, unsigned c) {
    if (P.use_diag) {
      unsigned long long k = spg_bitrev((unsigned)r, P.log_r);
      if (P.diag_table) {
        x = fp_mul_lazy(x, P.diag_table[((unsigned long long)c << P.log_r) | (unsigned)r]);
      } else {
        unsigned long long E = k * ((unsigned long long)c * P.ec + P.e0);
        if (P.inverse) E = (0ull - E);
        x = fp_mul_lazy(x, uni_pow(P, E));
      }
    }
    if (P.scale_lo) x = fp_mul_lazy(x, P.scale_lo[r]);
    if (P.scale_hi) x = fp_mul_lazy(x, P.scale_hi[b]);
    return x;
  }

  // One butterfly step of width W bits at bit position sh of r, for thread tid.  EDGE = (sh == 0): the step
  // whose twiddles are partly trivial (first step of a DIT tile, last step of a DIF tile).
  //
  // Lazy arithmetic (fp.cuh): sums are plain 256-bit additions, differences add a multiple K*p instead of
  // branching, and only multiplications (or fp_partial) bring values back below 2p.  bd[] tracks, at compile
  // time (everything is unrolled), an upper bound in units of p for each register:
  //   DIT: tile inputs < 2p; the edge step ends below 10p, every later stage adds 2p  (<= 24p at the store);
  //   DIF: every step starts below 2p, sums grow to <= 16p inside a step and are cut back with fp_partial at
  //        its end; the edge step leaves <= 16p to the store phase.
  // Everything stays below 32p < 2^256 and within the operand bound of fp_mul_lazy (a * b < 2^508).
  template <int W, bool DIT, bool EDGE>
  static SPG_HD void step(const NttPass& P, FpHalf* ws, int tid, int sh) {
    constexpr int GROUPS = EPT >> W;    // groups of 2^W elements per thread
    constexpr int GS = 1 << W;
    const int t = P.log_r;
    if (EDGE) sh = 0;
#pragma unroll
    for (int gi = 0; gi < GROUPS; gi++) {
      int u = gi * NT + tid;
      if (u >= (1 << (t + P.log_g - W))) continue;   // workspace larger than the whole problem
      int g, x;
      if (P.log_s) { g = u & ((1 << P.log_g) - 1); x = u >> P.log_g; }
      else { x = u & ((1 << (t - W)) - 1); g = u >> (t - W); }
      int lo = x & ((1 << sh) - 1), hi = x >> sh;
      int rbase = (hi << (sh + W)) | lo;
      Fp v[GS];
      int bd[GS];
#pragma unroll
      for (int f = 0; f < GS; f++) { v[f] = ws_load(ws, slot(P, rbase | (f << sh), g)); bd[f] = 2; }
      if (DIT) {
#pragma unroll
        for (int s = 0; s < W; s++) {
#pragma unroll
          for (int f = 0; f < GS; f++) {
            if (f & (1 << s)) continue;
            const int f2 = f | (1 << s);
            const bool trivial = EDGE && (f & ((1 << s) - 1)) == 0;
            Fp tq;
            int btq;
            if (trivial) {
              if (bd[f2] > 4) { tq = fp_partial(v[f2]); btq = 2; }
              else { tq = v[f2]; btq = bd[f2]; }
            } else {
              int e = (((f & ((1 << s) - 1)) << sh) | lo) << (t - 1 - (sh + s));
              tq = fp_mul_lazy(v[f2], P.tw[e << (SPG_TW_LOG - t)]);
              btq = 2;
            }
            const Fp a = v[f];
            v[f] = fp_add_raw(a, tq);
            v[f2] = fp_sub_lazy(a, tq, (uint32_t)btq);
            bd[f] = bd[f2] = bd[f] + btq;
          }
        }
      } else {
#pragma unroll
        for (int s = W - 1; s >= 0; s--) {
#pragma unroll
          for (int f = 0; f < GS; f++) {
            if (f & (1 << s)) continue;
            const int f2 = f | (1 << s);
            const bool trivial = EDGE && (f & ((1 << s) - 1)) == 0;
            const Fp a = v[f], bb = v[f2];
            const int bsum = bd[f] + bd[f2];
            v[f] = fp_add_raw(a, bb);
            Fp d = fp_sub_lazy(a, bb, (uint32_t)bd[f2]);
            if (trivial) { v[f2] = d; bd[f2] = bsum; }
            else {
              int e = (((f & ((1 << s) - 1)) << sh) | lo) << (t - 1 - (sh + s));
              v[f2] = fp_mul_lazy(d, P.tw[e << (SPG_TW_LOG - t)]);
              bd[f2] = 2;
            }
            bd[f] = bsum;
          }
        }
        if (!EDGE) {
#pragma unroll
          for (int f = 0; f < GS; f++)
            if (bd[f] > 2) v[f] = fp_partial(v[f]);
        }
      }
#pragma unroll
      for (int f = 0; f < GS; f++) ws_store(ws, slot(P, rbase | (f << sh), g), v[f]);
    }
  }

  // dispatch on runtime width / edge
  template <bool DIT>
  static SPG_HD void step_w(const NttPass& P, FpHalf* ws, int tid, int w, int sh) {
    if (sh == 0) {
      if (LOG_EPT >= 3 && w == 3) step<(LOG_EPT >= 3 ? 3 : 1), DIT, true>(P, ws, tid, 0);
      else if (w == 2) step<2, DIT, true>(P, ws, tid, 0);
      else step<1, DIT, true>(P, ws, tid, 0);
    } else {
      if (LOG_EPT >= 3 && w == 3) step<(LOG_EPT >= 3 ? 3 : 1), DIT, false>(P, ws, tid, sh);
      else if (w == 2) step<2, DIT, false>(P, ws, tid, sh);
      else step<1, DIT, false>(P, ws, tid, sh);
    }
  }

  // load phase for linear index idx of this CTA.  Values entering the butterflies are below 2p: canonical
  // input, a multiplied value, or (DIF, pass > 0) what the previous pass stored.
  template <bool DIT>
  static SPG_HD void load_one(const NttPass& P, FpHalf* ws, unsigned cta, unsigned col, int idx) {
    int r, g; unsigned b, c;
    if (idx >= (1 << (P.log_r + P.log_g))) return;
    unsigned long long off = gaddr(P, cta, idx, &r, &g, &b, &c);
    Fp x = P.in[col * P.in_col_stride + off];
    if (DIT) x = apply_factors(P, x, b, r, c);
    ws_store(ws, slot(P, r, g), x);
  }
  template <bool DIT>
  static SPG_HD void store_one(const NttPass& P, const FpHalf* ws, unsigned cta, unsigned col, int idx) {
    int r, g; unsigned b, c;
    if (idx >= (1 << (P.log_r + P.log_g))) return;
    unsigned long long off = gaddr(P, cta, idx, &r, &g, &b, &c);
    Fp x = ws_load(ws, slot(P, r, g));
    if (!DIT) x = apply_factors(P, x, b, r, c);
    // a DIF pass that is not the last one always has a diagonal factor, so x < 2p there; a DIT pass that is not
    // the last one is followed by a pass that multiplies every element at its load
    P.out[col * P.out_col_stride + off] = P.final_pass ? fp_reduce_full(x) : x;
  }
  // number of butterfly steps and the (width, shift) of step k.  DIF walks the bits of r from the
  // top, DIT from the bottom; the short step (log_r mod 3) comes last for DIF and first for DIT... both
  // choices keep strides monotone.
  static SPG_HD int n_steps(const NttPass& P) { return (P.log_r + LOG_EPT - 1) / LOG_EPT; }
  template <bool DIT>
  static SPG_HD void step_geom(const NttPass& P, int k, int* w, int* sh) {
    constexpr int L = LOG_EPT;
    int t = P.log_r, rem = t % L, ns = (t + L - 1) / L;
    if (DIT) {
      // ascending strides: first step has width rem (if any)
      if (rem) { *w = (k == 0) ? rem : L; *sh = (k == 0) ? 0 : rem + L * (k - 1); }
      else { *w = L; *sh = L * k; }
    } else {
      // descending strides: last step has width rem (if any)
      if (rem && k == ns - 1) { *w = rem; *sh = 0; }
      else { *w = L; *sh = t - L * (k + 1); }
    }
  }
};

// ------------------------------------------------------------------ compile-time tile (the dominant geometry)
// A pass whose tile is exactly the workspace and holds one column: log_r = LOG_R, log_g = 0 -- both passes of every 2^20
// (LOG_R = 10) and 2^22 (LOG_R = 11) transform, i.e. every launch of the headline LDE.  Same butterflies, lazy bounds and
// workspace layout as NttTile<LOG_R, 2>; what changes is the instruction stream around them:
//   * every shift, mask and step geometry is a compile-time constant (the generic tile decodes them per element);
//   * the swizzle is GF(2)-linear, so swz(rbase | f << SH) = swz(rbase) ^ swz_c(f << SH): one swizzle per thread and step,
//     the per-element parts fold into constants;
//   * the intra-tile twiddles omega_{2^LOG_R}^e (e < 2^(LOG_R-1)) are staged once per CTA in shared memory, swizzled like the
//     workspace.  From global memory every lane of a warp reads a different 128-byte line of the twiddle table -- 32 L1
//     wavefronts per load instruction, which made the twiddle reads 83 % of the kernel's L1 lookups and kept the L1 data
//     pipe 72 % busy (profiles/r2a_ntt_l1_summary.txt); from shared memory a warp's 16-byte reads are conflict-free.
SPG_HD constexpr int spg_swz_c(int s) {
  const int u = s >> 3;
  return s ^ ((u ^ (u >> 3) ^ (u >> 6) ^ (u >> 9)) & 7) ^ ((u & 2) << 1);
}

template <int LOG_R>
struct NttTileCT {
  typedef NttTile<LOG_R, 2> G;
  static constexpr int R = 1 << LOG_R, NT = R / 4, HW = R / 2, REM = LOG_R & 1, NS = (LOG_R + 1) / 2;
  static constexpr int LWIN = 7;          // log2 rows of a warp's window (32 lanes x 4 elements)

  // geometry of butterfly step k (same sequence as NttTile<LOG_R, 2>::step_geom)
  template <bool DIT> static SPG_HD constexpr int step_w(int k) { return REM ? (DIT ? (k == 0 ? 1 : 2) : (k == NS - 1 ? 1 : 2)) : 2; }
  template <bool DIT> static SPG_HD constexpr int step_sh(int k) {
    return DIT ? (REM ? (k == 0 ? 0 : 1 + 2 * (k - 1)) : 2 * k) : ((REM && k == NS - 1) ? 0 : LOG_R - 2 * (k + 1));
  }
  // a step that only touches the 2^LWIN-row window of the executing warp (see k_ntt_pass)
  template <bool DIT> static SPG_HD constexpr bool step_local(int k) {
    return k >= NS ? true : (step_w<DIT>(k) == 2 && step_sh<DIT>(k) + 2 <= LWIN);
  }

  static SPG_HD void stage_twiddles(const NttPass& P, FpHalf* tws, int tid) {
#pragma unroll
    for (int e = tid; e < HW; e += NT) {
      const Fp w = P.tw[e << (SPG_TW_LOG - LOG_R)];
      const int i = G::swz(e);
      FpHalf lo, hi;
#pragma unroll
      for (int k = 0; k < 4; k++) { lo.v[k] = w.v[k]; hi.v[k] = w.v[4 + k]; }
      tws[i] = lo; tws[HW + i] = hi;
    }
  }
  static SPG_HD Fp tw_at(const FpHalf* tws, int i) {      // i: swizzled position
    const FpHalf lo = tws[i], hi = tws[HW + i];
    Fp x;
#pragma unroll
    for (int k = 0; k < 4; k++) { x.v[k] = lo.v[k]; x.v[4 + k] = hi.v[k]; }
    return x;
  }
  static SPG_HD Fp ws_at(const FpHalf* ws, int i) {
    const FpHalf lo = ws[i], hi = ws[R + i];
    Fp x;
#pragma unroll
    for (int k = 0; k < 4; k++) { x.v[k] = lo.v[k]; x.v[4 + k] = hi.v[k]; }
    return x;
  }
  static SPG_HD void ws_put(FpHalf* ws, int i, const Fp& x) {
    FpHalf lo, hi;
#pragma unroll
    for (int k = 0; k < 4; k++) { lo.v[k] = x.v[k]; hi.v[k] = x.v[4 + k]; }
    ws[i] = lo; ws[R + i] = hi;
  }

  // global element offset of row r of this CTA's tile (log_g = 0)
  static SPG_HD unsigned long long goff(const NttPass& P, unsigned cta, int r, unsigned* b, unsigned* c) {
    if (P.log_s) {
      *b = cta >> P.log_s; *c = cta & ((1u << P.log_s) - 1u);
      return ((unsigned long long)*b << (LOG_R + P.log_s)) + ((unsigned long long)r << P.log_s) + *c;
    }
    *b = cta; *c = 0;
    return ((unsigned long long)cta << LOG_R) + (unsigned)r;
  }
  // row handled by (tid, j) in the load / store phases: warp w owns rows [128 w, 128 w + 128)
  static SPG_HD int io_row(int tid, int j) { return ((tid >> 5) << LWIN) | (j << 5) | (tid & 31); }

  template <bool DIT>
  static SPG_HD void load(const NttPass& P, FpHalf* ws, unsigned cta, unsigned col, int r) {
    unsigned b, c;
    const unsigned long long off = goff(P, cta, r, &b, &c);
    Fp x = P.in[col * P.in_col_stride + off];
    if (DIT) x = G::apply_factors(P, x, b, r, c);
    ws_put(ws, G::swz(r), x);
  }
  template <bool DIT>
  static SPG_HD void store(const NttPass& P, const FpHalf* ws, unsigned cta, unsigned col, int r) {
    unsigned b, c;
    const unsigned long long off = goff(P, cta, r, &b, &c);
    Fp x = ws_at(ws, G::swz(r));
    if (!DIT) x = G::apply_factors(P, x, b, r, c);
    P.out[col * P.out_col_stride + off] = P.final_pass ? fp_reduce_full(x) : x;
  }

  // one butterfly step of width W at bit position SH (both compile-time); arithmetic identical to NttTile::step
  template <int W, int SH, bool DIT>
  static SPG_HD void step(FpHalf* ws, const FpHalf* tws, int tid) {
    constexpr bool EDGE = (SH == 0);
    constexpr int GS = 1 << W, GROUPS = 4 >> W, t = LOG_R;
#pragma unroll
    for (int gi = 0; gi < GROUPS; gi++) {
      const int x = gi * NT + tid;
      const int lo = x & ((1 << SH) - 1), hi = x >> SH;
      const int sbase = G::swz((hi << (SH + W)) | lo);
      Fp v[GS];
      int bd[GS];
#pragma unroll
      for (int f = 0; f < GS; f++) { v[f] = ws_at(ws, sbase ^ spg_swz_c(f << SH)); bd[f] = 2; }
      if (DIT) {
#pragma unroll
        for (int s = 0; s < W; s++) {
          const int k = t - 1 - (SH + s);
          const int tbase = EDGE ? 0 : G::swz(lo << k);
#pragma unroll
          for (int f = 0; f < GS; f++) {
            if (f & (1 << s)) continue;
            const int f2 = f | (1 << s);
            const int cf = f & ((1 << s) - 1);
            const bool trivial = EDGE && cf == 0;
            Fp tq;
            int btq;
            if (trivial) {
              if (bd[f2] > 4) { tq = fp_partial(v[f2]); btq = 2; }
              else { tq = v[f2]; btq = bd[f2]; }
            } else {
              tq = fp_mul_lazy(v[f2], tw_at(tws, tbase ^ spg_swz_c(cf << (SH + k))));
              btq = 2;
            }
            const Fp a = v[f];
            v[f] = fp_add_raw(a, tq);
            v[f2] = fp_sub_lazy(a, tq, (uint32_t)btq);
            bd[f] = bd[f2] = bd[f] + btq;
          }
        }
      } else {
#pragma unroll
        for (int s = W - 1; s >= 0; s--) {
          const int k = t - 1 - (SH + s);
          const int tbase = EDGE ? 0 : G::swz(lo << k);
#pragma unroll
          for (int f = 0; f < GS; f++) {
            if (f & (1 << s)) continue;
            const int f2 = f | (1 << s);
            const int cf = f & ((1 << s) - 1);
            const bool trivial = EDGE && cf == 0;
            const Fp a = v[f], bb = v[f2];
            const int bsum = bd[f] + bd[f2];
            v[f] = fp_add_raw(a, bb);
            const Fp d = fp_sub_lazy(a, bb, (uint32_t)bd[f2]);
            if (trivial) { v[f2] = d; bd[f2] = bsum; }
            else {
              v[f2] = fp_mul_lazy(d, tw_at(tws, tbase ^ spg_swz_c(cf << (SH + k))));
              bd[f2] = 2;
            }
            bd[f] = bsum;
          }
        }
        if (!EDGE) {
#pragma unroll
          for (int f = 0; f < GS; f++)
            if (bd[f] > 2) v[f] = fp_partial(v[f]);
        }
      }
#pragma unroll
      for (int f = 0; f < GS; f++) ws_put(ws, sbase ^ spg_swz_c(f << SH), v[f]);
    }
  }

  // step K (compile-time) and, for the host emulation, step k (run-time dispatch onto the same instantiations)
  template <bool DIT, int K>
  static SPG_HD void step_k(FpHalf* ws, const FpHalf* tws, int tid) {
    step<step_w<DIT>(K), step_sh<DIT>(K), DIT>(ws, tws, tid);
  }
  template <bool DIT>
  static SPG_HD void step_rt(int k, FpHalf* ws, const FpHalf* tws, int tid) {
    switch (k) {
      case 0: step_k<DIT, 0>(ws, tws, tid); break;
      case 1: step_k<DIT, 1>(ws, tws, tid); break;
      case 2: step_k<DIT, 2>(ws, tws, tid); break;
      case 3: step_k<DIT, 3>(ws, tws, tid); break;
      case 4: step_k<DIT, 4>(ws, tws, tid); break;
      default: if (NS > 5) step_k<DIT, (NS > 5 ? 5 : 0)>(ws, tws, tid); break;
    }
  }
  // can pass P run on this tile?
  static SPG_HD bool fits(const NttPass& P) { return P.log_r == LOG_R && P.log_g == 0; }
};

// ------------------------------------------------------------------ pass planner (host)
// Split log_n into per-pass tile sizes (each <= max_bits = log2 of the workspace, as even as possible, larger first).
static inline int spg_ntt_plan_bits(unsigned log_n, int bits[8], int max_bits) {
  if (log_n == 0) { bits[0] = 0; return 1; }
  int np = ((int)log_n + max_bits - 1) / max_bits;
  int base = log_n / np, extra = log_n % np;
  for (int i = 0; i < np; i++) bits[i] = base + (i < extra ? 1 : 0);
  return np;
}

// geometry of the pass that touches contiguous tiles (last for DIF, first for DIT):
// it sees the column as [2^log_b_hi blocks][2^log_r_lo rows]
static inline void spg_ntt_last_pass_geometry(unsigned log_n, int max_bits, int* log_r_lo, int* log_b_hi) {
  int bits[8];
  int np = spg_ntt_plan_bits(log_n, bits, max_bits);
  *log_r_lo = bits[np - 1];
  *log_b_hi = (int)log_n - bits[np - 1];
}

// Build the pass list of one transform.  bits[0] covers the most significant index bits (largest
// stride); DIF runs passes 0..np-1, DIT runs them in the opposite order (contiguous pass first).
// coset_exp (DIT only): the input at coefficient index k is multiplied by omega_{2^26}^(coset_exp*k).
// scale_lo / scale_hi are attached to the contiguous pass.
static inline int spg_ntt_make_passes(NttPass* passes, int log_ws, const Fp* in, Fp* out, unsigned log_n,
                                      unsigned long long in_stride, unsigned long long out_stride,
                                      int inverse, int dit, unsigned long long coset_exp, const Fp* scale_lo,
                                      const Fp* scale_hi, const Fp* tw_fwd, const Fp* tw_inv, const Fp* uniA,
                                      const Fp* uniB) {
  int bits[8];
  const int np = spg_ntt_plan_bits(log_n, bits, log_ws);
  for (int pi = 0; pi < np; pi++) {
    const int i = dit ? np - 1 - pi : pi;
    int log_s = 0;
    for (int k = i + 1; k < np; k++) log_s += bits[k];
    NttPass& P = passes[pi];
    P.in = (pi == 0) ? in : out;
    P.out = out;
    P.in_col_stride = (pi == 0) ? in_stride : out_stride;
    P.out_col_stride = out_stride;
    P.log_n = (int)log_n;
    P.log_r = bits[i];
    P.log_s = log_s;
    P.inverse = inverse;
    P.tw = inverse ? tw_inv : tw_fwd;
    P.uniA = uniA;
    P.uniB = uniB;
    const int log_b = (int)log_n - P.log_r - log_s;
    int log_g = log_ws - P.log_r;
    if (log_s) { if (log_g > log_s) log_g = log_s; }
    else { if (log_g > log_b) log_g = log_b; }
    P.log_g = log_g;
    P.ec = log_s ? (1ull << (SPG_UNI_LOG - (P.log_r + log_s))) : 0ull;
    // coefficient index k = bitrev_n(pos); the r field of pos contributes bitrev_R(r) << log_b
    P.e0 = coset_exp << log_b;
    P.use_diag = (log_s != 0) || (dit && coset_exp != 0);
    P.scale_lo = nullptr; P.scale_hi = nullptr;
    if (log_s == 0) { P.scale_lo = scale_lo; P.scale_hi = scale_hi; }
    P.final_pass = (pi == np - 1);
    P.diag_table = nullptr;
  }
  return np;
}

#include <vector>
// LDE scale tables for the store phase of the inverse DIF: output position b*R + r holds coefficient
// k = bitrev_R(r) * B + bitrev_B(b), which gets  g^k / N = lo[r] * hi[b].
static inline void spg_lde_scale_tables(unsigned log_n, int max_bits, const Fp& g_mont, std::vector<Fp>& lo, std::vector<Fp>& hi) {
  int lr, lb;
  spg_ntt_last_pass_geometry(log_n, max_bits, &lr, &lb);
  const size_t R = (size_t)1 << lr, B = (size_t)1 << lb;
  lo.resize(R); hi.resize(B);
  uint64_t nn[4] = {(uint64_t)1 << log_n, 0, 0, 0};
  const Fp ninv = fp_inv(fp_to_mont(fp_from_u64(nn)));
  std::vector<Fp> pw(B), pr(R);
  pw[0] = fp_one();
  for (size_t i = 1; i < B; i++) pw[i] = fp_mul(pw[i - 1], g_mont);
  for (size_t b = 0; b < B; b++) hi[b] = pw[spg_bitrev((unsigned)b, lb)];
  const Fp gB = fp_mul(pw[B - 1], g_mont);
  pr[0] = ninv;
  for (size_t i = 1; i < R; i++) pr[i] = fp_mul(pr[i - 1], gB);
  for (size_t r = 0; r < R; r++) lo[r] = pr[spg_bitrev((unsigned)r, lr)];
}

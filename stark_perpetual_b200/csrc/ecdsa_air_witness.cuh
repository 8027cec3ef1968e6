// Witness generation of the ECDSA-builtin AIR, per-thread code (the kernels of air_ecdsa.cu are thin wrappers; the same
// functions run on the host in tests/host_emul/emul_ecdsa_air_witness.cpp against the oracle twin's trace).
// What the reference pins here is every row: signature.py:176-190 (mimic_ec_mult_air), :243-260 (verify).
#pragma once
#include "ec.cuh"
#include "ecdsa_air_point.cuh"

#ifdef __CUDA_ARCH__
#define SPG_STATUS_OR(p, v) atomicOr((p), (v))
#else
#define SPG_STATUS_OR(p, v) (*(p) |= (v))
#endif

// Three kernels (the shape of the Pedersen witness, air.cu).  (1) k_eair_walk_ab / (2) k_eair_walk_c: one thread per
// mimic_ec_mult_air(m, point, shift) (signature.py:176-190) walks its 251 steps in JACOBIAN coordinates -- no inversion on
// the sequential path -- and parks, per row, the partial sum (X, Y, Z) in the lane's (PX, PY, SA) cells and the doubled
// point (X, Y, Z) in (QX, QY, SD), Montgomery form; lane A's point is the table 2^t G and has no cells.  Kernel (2) also
// does the two hand-over additions of row 255 (their slope and inverse go to a side buffer) and the closing r == x check.
// (3) k_eair_finish: one thread per 8 consecutive rows of a lane turns them into the affine witness with ONE Fermat
// inversion for the 32 values it needs (Montgomery's trick): 1/Z1, 1/Z2, and -- without waiting for the affine values --
// 1/(PX - QX) = Z1^2 Z2^2 / (X1 Z2^2 - X2 Z1^2) and 1/(2 QY) = Z2^3 / (2 Y2).  A zero among those is exactly an
// assertion of the reference firing (x collision, y = 0): status bit 2.
struct EairLaneCols { Fp *M, *PX, *PY, *QX, *QY, *SA, *SD, *I; };
#define SPG_EAIR_WIT_ROWS 8

SPG_HD void eair_walk(Fp m /*canonical scalar*/, JPoint ps, JPoint q, const APoint* gd,
                                 const EairLaneCols& L, size_t base, JPoint* out) {
#pragma unroll 1
  for (int t = 0; t < SPG_EAIR_BLOCK; t++) {
    const size_t r = base + t;
    L.M[r] = m;
    L.PX[r] = ps.X; L.PY[r] = ps.Y; L.SA[r] = ps.Z;
    if (t <= SPG_EAIR_BITS && !gd) { L.QX[r] = q.X; L.QY[r] = q.Y; L.SD[r] = q.Z; }
    if (t < SPG_EAIR_BITS) {
      if (m.v[0] & 1u) {
        const Fp z1z1 = fp_sqr(ps.Z);
        if (gd) {
          const APoint g = gd[t];
          ps = ec_madd_nocheck(ps, g, fp_mul(z1z1, ps.Z), fp_mul(g.x, z1z1));
        } else {
          const Fp z2z2 = fp_sqr(q.Z);
          ps = ec_jadd_nocheck(ps, q, z1z1, z2z2, fp_mul(ps.X, z2z2), fp_mul(q.X, z1z1));
        }
      }
      if (!gd) q = ec_jdouble_nocheck(q);
    }
#pragma unroll
    for (int k = 0; k < 7; k++) m.v[k] = (m.v[k] >> 1) | (m.v[k + 1] << 31);
    m.v[7] >>= 1;
  }
  *out = ps;
}

SPG_HD bool eair_scalar_ok(const Fp& m) {     // 0 < m < 2^251 (signature.py:180, :219-227)
  uint32_t any = 0;
  for (int k = 0; k < 8; k++) any |= m.v[k];
  return any != 0 && (m.v[7] >> 27) == 0;
}

// lanes A (z G) and B (r Q) of every block: one thread per (lane, block)
SPG_HD void eair_walk_ab_thread(size_t idx, unsigned log_n, const Fp* msg, const Fp* rr, const Fp* kx, const Fp* ky, Fp* trace,
                                uint32_t* status, const APoint* gd, APoint shift, Fp beta) {
  const size_t n = (size_t)1 << log_n, nb = n >> 8;
  if (idx >= 2 * nb) return;
  const size_t lane = idx / nb, b = idx - lane * nb, base = b << 8;
  auto col = [&](int c) { return trace + ((size_t)c << log_n); };
  JPoint out, ps, q;
  ps.X = shift.x; ps.Y = shift.y; ps.Z = fp_one();
  if (lane == 0) {
    const Fp m = msg[b];
    if (!eair_scalar_ok(m)) { SPG_STATUS_OR(status, 1u); return; }
    ps.Y = fp_neg(ps.Y);                                             // MINUS_SHIFT_POINT (signature.py:252)
    EairLaneCols L = {col(EA_AM), col(EA_APX), col(EA_APY), nullptr, nullptr, col(EA_ASA), nullptr, col(EA_AI)};
    eair_walk(m, ps, ps, gd, L, base, &out);
  } else {
    const Fp m = rr[b];
    q.X = fp_to_mont(kx[b]); q.Y = fp_to_mont(ky[b]); q.Z = fp_one();
    const Fp rhs = fp_add(fp_add(fp_mul(fp_sqr(q.X), q.X), q.X), beta);
    if (!eair_scalar_ok(m) || spg_canon_geq_p(kx[b].v) || spg_canon_geq_p(ky[b].v) || !fp_eq(fp_sqr(q.Y), rhs)) {
      SPG_STATUS_OR(status, 1u);
      return;
    }
    EairLaneCols L = {col(EA_BM), col(EA_BPX), col(EA_BPY), col(EA_BQX), col(EA_BQY), col(EA_BSA), col(EA_BSD), col(EA_BI)};
    eair_walk(m, ps, q, nullptr, L, base, &out);
  }
}

// affine x, y of a parked Jacobian point pair with one inversion; false if a Z is zero (a collision upstream)
SPG_HD bool eair_affine2(const JPoint& a, const JPoint& b, APoint* oa, APoint* ob) {
  const Fp zz = fp_mul(a.Z, b.Z);
  if (fp_is_zero(zz)) return false;
  const Fp iv = fp_inv_chain(zz), ia = fp_mul(iv, b.Z), ib = fp_mul(iv, a.Z);
  const Fp ia2 = fp_sqr(ia), ib2 = fp_sqr(ib);
  oa->x = fp_mul(a.X, ia2); oa->y = fp_mul(a.Y, fp_mul(ia2, ia));
  ob->x = fp_mul(b.X, ib2); ob->y = fp_mul(b.Y, fp_mul(ib2, ib));
  return true;
}

// lane C of every block (the signature of the block before), the two hand-over additions (cross[b] = slope and inverse of
// zG + rQ for block b's last row, slope and inverse of wB - S for lane C's), carriers and non-zero witnesses
SPG_HD void eair_walk_c_thread(size_t b, unsigned log_n, const Fp* msg, const Fp* rr, const Fp* ww, Fp* trace, Fp* cross,
                               uint32_t* status, APoint shift) {
  const size_t n = (size_t)1 << log_n, nb = n >> 8;
  if (b >= nb) return;
  const size_t sb = (b + nb - 1) % nb, base = b << 8, last_s = (sb << 8) + SPG_EAIR_BLOCK - 1;
  auto col = [&](int c) { return trace + ((size_t)c << log_n); };
  uint32_t st = 0;
  const Fp w = ww[sb], r_s = rr[sb], r_b = rr[b], z_b = msg[b];
  for (int t = 0; t < SPG_EAIR_BLOCK; t++) { col(EA_T1)[base + t] = r_b; col(EA_T2)[base + t] = r_s; }
  if (!eair_scalar_ok(w) || !eair_scalar_ok(r_b) || !eair_scalar_ok(z_b) || !eair_scalar_ok(r_s)) { SPG_STATUS_OR(status, 1u); return; }
  {
    const Fp zr = fp_mul(fp_to_mont(z_b), fp_to_mont(r_b)), wm = fp_to_mont(w);
    const Fp iv = fp_inv_chain(fp_mul(zr, wm));
    col(EA_V1)[base] = fp_from_mont(fp_mul(iv, wm));
    col(EA_V2)[base] = fp_from_mont(fp_mul(iv, zr));
  }
  // ec_add(zG, rQ) (signature.py:254): both partial sums are parked on the last row of block sb
  JPoint jz, jr;
  jz.X = col(EA_APX)[last_s]; jz.Y = col(EA_APY)[last_s]; jz.Z = col(EA_ASA)[last_s];
  jr.X = col(EA_BPX)[last_s]; jr.Y = col(EA_BPY)[last_s]; jr.Z = col(EA_BSA)[last_s];
  APoint zg, rq;
  if (!eair_affine2(jz, jr, &zg, &rq)) { SPG_STATUS_OR(status, 2u); return; }
  Fp d = fp_sub(zg.x, rq.x);
  if (fp_is_zero(d)) { st |= 2; d = fp_one(); }
  Fp di = fp_inv_chain(d);
  Fp s = fp_mul(fp_sub(zg.y, rq.y), di);
  cross[4 * sb + 0] = fp_from_mont(s); cross[4 * sb + 1] = fp_from_mont(di);
  JPoint sum, ps, wbj;
  sum.X = fp_sub(fp_sub(fp_sqr(s), zg.x), rq.x);
  sum.Y = fp_sub(fp_mul(s, fp_sub(zg.x, sum.X)), zg.y);
  sum.Z = fp_one();
  ps.X = shift.x; ps.Y = shift.y; ps.Z = fp_one();
  EairLaneCols L = {col(EA_CM), col(EA_CPX), col(EA_CPY), col(EA_CQX), col(EA_CQY), col(EA_CSA), col(EA_CSD), col(EA_CI)};
  eair_walk(w, ps, sum, nullptr, L, base, &wbj);
  // ec_add(wB, MINUS_SHIFT_POINT).x == r (signature.py:257-260)
  if (fp_is_zero(wbj.Z)) { SPG_STATUS_OR(status, st | 2u); return; }
  const Fp zi = fp_inv_chain(wbj.Z), zi2 = fp_sqr(zi);
  APoint wb; wb.x = fp_mul(wbj.X, zi2); wb.y = fp_mul(wbj.Y, fp_mul(zi2, zi));
  d = fp_sub(wb.x, shift.x);
  if (fp_is_zero(d)) { st |= 2; d = fp_one(); }
  di = fp_inv_chain(d);
  s = fp_mul(fp_add(wb.y, shift.y), di);
  cross[4 * b + 2] = fp_from_mont(s); cross[4 * b + 3] = fp_from_mont(di);
  const Fp x = fp_sub(fp_sub(fp_sqr(s), wb.x), shift.x);
  if (!fp_eq(x, fp_to_mont(r_s))) st |= 4;
  if (st) SPG_STATUS_OR(status, st);
}

// parked Jacobian rows -> the affine witness; one thread per (lane, 8 rows)
SPG_HD void eair_finish_thread(size_t idx, unsigned log_n, Fp* trace, const Fp* cross, uint32_t* status, const APoint* gd) {
  const size_t n = (size_t)1 << log_n, groups = n / SPG_EAIR_WIT_ROWS;
  if (idx >= 3 * groups) return;
  const int lane = (int)(idx / groups);
  const size_t r0 = (idx - (size_t)lane * groups) * SPG_EAIR_WIT_ROWS;
  const int c0 = lane == 0 ? EA_AM : lane == 1 ? EA_BM : EA_CM;
  auto col = [&](int c) { return trace + ((size_t)c << log_n); };
  Fp *M = col(c0), *PX = col(c0 + 1), *PY = col(c0 + 2);
  Fp *QX = lane ? col(c0 + 3) : nullptr, *QY = lane ? col(c0 + 4) : nullptr;
  Fp *SA = col(lane ? c0 + 5 : EA_ASA), *SD = lane ? col(c0 + 6) : nullptr, *I = col(lane ? c0 + 7 : EA_AI);
  const Fp one = fp_one();
  // values to invert per row: Z1, Z2, D = X1 Z2^2 - X2 Z1^2, E = 2 Y2 (1 where the row has none)
  Fp pre[4 * SPG_EAIR_WIT_ROWS];
  Fp run = one;
  uint32_t st = 0;
#pragma unroll 1
  for (int k = 0; k < SPG_EAIR_WIT_ROWS; k++) {
    const size_t r = r0 + k;
    const int t = (int)(r & (SPG_EAIR_BLOCK - 1));
    Fp z1 = SA[r], z2 = one, dd = one, ee = one;
    if (lane == 0) {
      if (t < SPG_EAIR_BITS) dd = fp_sub(PX[r], fp_mul(gd[t].x, fp_sqr(z1)));
    } else if (t <= SPG_EAIR_BITS) {
      z2 = SD[r];
      if (t < SPG_EAIR_BITS) {
        dd = fp_sub(fp_mul(PX[r], fp_sqr(z2)), fp_mul(QX[r], fp_sqr(z1)));
        ee = fp_add(QY[r], QY[r]);
      }
    }
    if (fp_is_zero(z1)) { st |= 2; z1 = one; }
    if (fp_is_zero(z2)) { st |= 2; z2 = one; }
    if (fp_is_zero(dd)) { st |= 2; dd = one; }                     // assert partial_sum[0] != point[0] (signature.py:183)
    if (fp_is_zero(ee)) { st |= 2; ee = one; }                     // ec_double's y != 0 (math_utils.py:80)
    pre[4 * k] = run; run = fp_mul(run, z1);
    pre[4 * k + 1] = run; run = fp_mul(run, z2);
    pre[4 * k + 2] = run; run = fp_mul(run, dd);
    pre[4 * k + 3] = run; run = fp_mul(run, ee);
  }
  Fp inv = fp_inv_chain(run);
#pragma unroll 1
  for (int k = SPG_EAIR_WIT_ROWS - 1; k >= 0; k--) {
    const size_t r = r0 + k;
    const int t = (int)(r & (SPG_EAIR_BLOCK - 1));
    const size_t b = r >> 8;
    const Fp x1 = PX[r], y1 = PY[r];
    Fp z1 = SA[r], z2 = one, dd = one, ee = one, x2 = fp_zero(), y2 = fp_zero();
    const bool has_q = lane != 0 && t <= SPG_EAIR_BITS, step = t < SPG_EAIR_BITS;
    if (lane == 0) {
      if (step) { x2 = gd[t].x; y2 = gd[t].y; dd = fp_sub(x1, fp_mul(x2, fp_sqr(z1))); }
    } else if (has_q) {
      z2 = SD[r]; x2 = QX[r]; y2 = QY[r];
      if (step) { dd = fp_sub(fp_mul(x1, fp_sqr(z2)), fp_mul(x2, fp_sqr(z1))); ee = fp_add(y2, y2); }
    }
    if (fp_is_zero(z1)) z1 = one;
    if (fp_is_zero(z2)) z2 = one;
    if (fp_is_zero(dd)) dd = one;
    if (fp_is_zero(ee)) ee = one;
    const Fp ie = fp_mul(inv, pre[4 * k + 3]); inv = fp_mul(inv, ee);
    const Fp id = fp_mul(inv, pre[4 * k + 2]); inv = fp_mul(inv, dd);
    const Fp iz2 = fp_mul(inv, pre[4 * k + 1]); inv = fp_mul(inv, z2);
    const Fp iz1 = fp_mul(inv, pre[4 * k]); inv = fp_mul(inv, z1);
    const Fp iz1s = fp_sqr(iz1);
    const Fp px = fp_mul(x1, iz1s), py = fp_mul(y1, fp_mul(iz1s, iz1));
    Fp qx = x2, qy = y2;                                             // lane A: the table point is affine already
    if (has_q) { const Fp iz2s = fp_sqr(iz2); qx = fp_mul(x2, iz2s); qy = fp_mul(y2, fp_mul(iz2s, iz2)); }
    Fp sa = fp_zero(), sd = fp_zero(), ii = fp_zero();
    if (step) {
      const Fp z1s = fp_sqr(z1), z2s = fp_sqr(z2);
      ii = fp_mul(fp_mul(z1s, z2s), id);                              // 1 / (px - qx)
      if (M[r].v[0] & 1u) sa = fp_mul(fp_sub(py, qy), ii);
      if (lane) {
        const Fp qq = fp_sqr(qx);
        sd = fp_mul(fp_add(fp_add(fp_add(qq, qq), qq), one), fp_mul(fp_mul(z2s, z2), ie));   // (3 qx^2 + 1) / (2 qy)
      }
      sa = fp_from_mont(sa); sd = fp_from_mont(sd); ii = fp_from_mont(ii);
    } else if (t == SPG_EAIR_BLOCK - 1 && lane != 1) {               // the hand-over cells of row 255 (already canonical)
      sa = cross[4 * b + (lane ? 2 : 0)]; ii = cross[4 * b + (lane ? 3 : 1)];
    }
    PX[r] = fp_from_mont(px); PY[r] = fp_from_mont(py);
    if (lane) {
      if (has_q) { QX[r] = fp_from_mont(qx); QY[r] = fp_from_mont(qy); }
      SD[r] = sd;
    }
    SA[r] = sa; I[r] = ii;
  }
  if (st) SPG_STATUS_OR(status, st);
}


// Per-point evaluation of the Pedersen hash-chain AIR's constraints (DESIGN.md "AIR"; CPU restatement:
// oracle/stark.py Air.composition), shared by the CUDA kernel k_air_eval (air.cu) and the host emulation
// (tests/host_emul/emul_air.cpp, where -DSPG_EMUL_LAZY checks every lazy bound below on adversarial inputs).
// No reference symbol exists for this stage; the step function it constrains is the
// reference's pedersen_hash loop (signature.py:308-317).
#pragma once
#include "fp.cuh"

#define SPG_AIR_LANES 5
#define SPG_AIR_NCONSTR 13
// Scalars are unpacked into 251 bits: c6 (M = 0) holds from row 251 of every 256-row element block on, so the bits spell
// an integer below 2^251 < p -- the unique canonical representative the reference hashes (signature.py:307).  With the
// 252 steps the reference walks, x and x + p (x < 2^252 - p) would both satisfy the constraints.
#define SPG_AIR_CANON_BITS 251

struct AirEvalConsts {
  Fp alpha[SPG_AIR_LANES * SPG_AIR_NCONSTR];
  Fp x0[SPG_AIR_LANES], outs[SPG_AIR_LANES];
  Fp shift_x, shift_y;
};

// one lane's eight cells at a point: current row (X, Y, slope S, remaining bits M, inverse I) and next row
struct AirRow { Fp X, Y, S, M, I, Xn, Yn, Mn; };
// accumulators per zerofier group
struct AirAcc { Fp step, act, pad, mid, link, inst, seg, last; };

SPG_HD void air_acc_init(AirAcc& A) {
  A.step = fp_zero(); A.act = fp_zero(); A.pad = fp_zero(); A.mid = fp_zero();
  A.link = fp_zero(); A.inst = fp_zero(); A.seg = fp_zero(); A.last = fp_zero();
}

// Adds lane l's thirteen constraints, each times its power of alpha, to the accumulators.
// Lazy arithmetic (fp.cuh): trace values, the periodic point (px, py), alpha powers and public values are canonical
// (< p); products are below 2p; sums are plain additions and differences add K*p (K = bound of the subtrahend in units
// of p, in brackets below); each accumulator is brought back below 2^252 once per lane, so it never exceeds
// 2p + 4 * 2p = 10p.
SPG_HD void air_lane_accumulate(AirAcc& A, const AirRow& r, const Fp* al, const Fp& px, const Fp& py, const Fp& shift_x,
                                const Fp& shift_y, const Fp& x0, const Fp& out) {
  const Fp one = fp_one();
  const Fp bit = fp_sub_lazy(r.M, fp_add_raw(r.Mn, r.Mn), 2);                    // [< 3p]
  const Fp dx = fp_sub_lazy(r.X, px, 1);                                         // [< 2p]
  const Fp dXn = fp_sub_lazy(r.Xn, r.X, 1), dYn = fp_sub_lazy(r.Yn, r.Y, 1);     // Xn - X, Yn - Y  [< 2p]
  // c1 = bit (bit - 1)
  const Fp c1 = fp_mul_lazy(bit, fp_sub_lazy(bit, one, 1));
  // c2 = bit (S (X - px) - (Y - py))
  const Fp c2 = fp_mul_lazy(bit, fp_sub_lazy(fp_mul_lazy(r.S, dx), fp_sub_lazy(r.Y, py, 1), 2));
  // c3 = bit (S^2 - px - 2 Xn) + (Xn - X)      [== bit (S^2 - X - px - Xn) + (1 - bit)(Xn - X)]
  const Fp c3 = fp_add_raw(fp_mul_lazy(bit, fp_sub_lazy(fp_sub_lazy(fp_sqr_lazy(r.S), px, 1), fp_add_raw(r.Xn, r.Xn), 2)), dXn);
  // c4 = bit (S (X - Xn) - 2 Yn) + (Yn - Y)    [== bit (S (X - Xn) - Y - Yn) + (1 - bit)(Yn - Y)]
  const Fp c4 = fp_add_raw(fp_mul_lazy(bit, fp_sub_lazy(fp_mul_lazy(r.S, fp_sub_lazy(r.X, r.Xn, 1)), fp_add_raw(r.Yn, r.Yn), 2)), dYn);
  A.step = fp_partial(fp_add_raw(fp_add_raw(A.step, fp_add_raw(fp_mul_lazy(al[0], c1), fp_mul_lazy(al[1], c2))),
                                 fp_add_raw(fp_mul_lazy(al[2], c3), fp_mul_lazy(al[3], c4))));
  // c5 = I (X - px) - 1
  A.act = fp_partial(fp_add_raw(A.act, fp_mul_lazy(al[4], fp_sub_lazy(fp_mul_lazy(r.I, dx), one, 1))));
  // c6 = M
  A.pad = fp_partial(fp_add_raw(A.pad, fp_mul_lazy(al[5], r.M)));
  // c7 = Xn - X, c8 = Yn - Y
  A.mid = fp_partial(fp_add_raw(A.mid, fp_add_raw(fp_mul_lazy(al[6], dXn), fp_mul_lazy(al[7], dYn))));
  // c9 = Mn - X
  A.link = fp_partial(fp_add_raw(A.link, fp_mul_lazy(al[8], fp_sub_lazy(r.Mn, r.X, 1))));
  // c10 = X - shift.x, c11 = Y - shift.y
  A.inst = fp_partial(fp_add_raw(A.inst, fp_add_raw(fp_mul_lazy(al[9], fp_sub_lazy(r.X, shift_x, 1)),
                                                    fp_mul_lazy(al[10], fp_sub_lazy(r.Y, shift_y, 1)))));
  // c12 = M - x0
  A.seg = fp_partial(fp_add_raw(A.seg, fp_mul_lazy(al[11], fp_sub_lazy(r.M, x0, 1))));
  // c13 = X - out
  A.last = fp_partial(fp_add_raw(A.last, fp_mul_lazy(al[12], fp_sub_lazy(r.X, out, 1))));
}

// sum over the zerofier groups of accumulator * inverse zerofier: eight products (< 16p), canonical result
SPG_HD Fp air_combine(const AirAcc& A, const Fp& iz_step, const Fp& iz_act, const Fp& iz_pad, const Fp& iz_mid,
                      const Fp& iz_link, const Fp& iz_inst, const Fp& iz_seg, const Fp& iz_last) {
  Fp acc = fp_mul_lazy(A.step, iz_step);
  acc = fp_add_raw(acc, fp_mul_lazy(A.act, iz_act));
  acc = fp_add_raw(acc, fp_mul_lazy(A.pad, iz_pad));
  acc = fp_add_raw(acc, fp_mul_lazy(A.mid, iz_mid));
  acc = fp_add_raw(acc, fp_mul_lazy(A.link, iz_link));
  acc = fp_add_raw(acc, fp_mul_lazy(A.inst, iz_inst));
  acc = fp_add_raw(acc, fp_mul_lazy(A.seg, iz_seg));
  acc = fp_add_raw(acc, fp_mul_lazy(A.last, iz_last));
  return fp_reduce_full(acc);
}

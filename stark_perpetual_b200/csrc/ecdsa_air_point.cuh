// Per-point evaluation of the ECDSA-builtin AIR's constraints (DESIGN.md "Second AIR"; CPU restatement:
// oracle/stark_ecdsa.py EcdsaAir.composition_per), shared by the CUDA kernel k_air_eval_ecdsa, the prover's host
// self-check (air_ecdsa.cu) and the host emulation (tests/host_emul/emul_ecdsa_air.cpp).
// No reference symbol exists for the constraint system; the steps it constrains are the
// reference's: signature.py:176-190 (mimic_ec_mult_air), :243-260 (verify), math_utils.py:59-88 (ec_add, ec_double).
#pragma once
#include "fp.cuh"

#define SPG_EAIR_COLS 25
#define SPG_EAIR_NALPHA 52
#define SPG_EAIR_BLOCK 256
#define SPG_EAIR_BITS 251          // N_ELEMENT_BITS_ECDSA, signature.py:47
#define SPG_EAIR_NGROUPS 6         // step, hold, zero, first, last, thold

// column indices: lane A (z G, shift -S; the point 2^t G is periodic), lane B (r Q), lane C (w (zG + rQ)), carriers
enum {
  EA_AM = 0, EA_APX, EA_APY, EA_ASA, EA_AI,
  EA_BM, EA_BPX, EA_BPY, EA_BQX, EA_BQY, EA_BSA, EA_BSD, EA_BI,
  EA_CM, EA_CPX, EA_CPY, EA_CQX, EA_CQY, EA_CSA, EA_CSD, EA_CI,
  EA_T1, EA_T2, EA_V1, EA_V2
};

struct EcdsaAirConsts {            // Montgomery form
  Fp alpha[SPG_EAIR_NALPHA];
  Fp shift_x, shift_y, minus_shift_y, beta;
};

// the five step constraints every lane has: bit, addition slope, x, y, x-distinctness (signature.py:183-185)
SPG_HD Fp eair_lane(const Fp* a, const Fp& M, const Fp& Mn, const Fp& PX, const Fp& PY, const Fp& PXn, const Fp& PYn,
                    const Fp& QX, const Fp& QY, const Fp& SA, const Fp& I) {
  const Fp one = fp_one();
  const Fp bit = fp_sub(M, fp_add(Mn, Mn)), nb = fp_sub(one, bit), dx = fp_sub(PX, QX);
  const Fp c1 = fp_mul(bit, fp_sub(bit, one));
  const Fp c2 = fp_mul(bit, fp_sub(fp_mul(SA, dx), fp_sub(PY, QY)));
  const Fp c3 = fp_add(fp_mul(bit, fp_sub(fp_sub(fp_sub(fp_sqr(SA), PX), QX), PXn)), fp_mul(nb, fp_sub(PXn, PX)));
  const Fp c4 = fp_add(fp_mul(bit, fp_sub(fp_sub(fp_mul(SA, fp_sub(PX, PXn)), PY), PYn)), fp_mul(nb, fp_sub(PYn, PY)));
  const Fp c5 = fp_sub(fp_mul(I, dx), one);
  Fp s = fp_add(fp_mul(a[0], c1), fp_mul(a[1], c2));
  s = fp_add(s, fp_add(fp_mul(a[2], c3), fp_mul(a[3], c4)));
  return fp_add(s, fp_mul(a[4], c5));
}

// the doubling of the point column (math_utils.py:79-88 with alpha = 1)
SPG_HD Fp eair_double(const Fp* a, const Fp& QX, const Fp& QY, const Fp& QXn, const Fp& QYn, const Fp& SD) {
  const Fp qq = fp_sqr(QX);
  const Fp sdy = fp_mul(SD, QY);
  const Fp d1 = fp_sub(fp_sub(fp_add(sdy, sdy), fp_add(fp_add(qq, qq), qq)), fp_one());
  const Fp d2 = fp_sub(fp_sub(fp_sqr(SD), fp_add(QX, QX)), QXn);
  const Fp d3 = fp_sub(fp_sub(fp_mul(SD, fp_sub(QX, QXn)), QY), QYn);
  return fp_add(fp_add(fp_mul(a[0], d1), fp_mul(a[1], d2)), fp_mul(a[2], d3));
}

// c / n: the 25 cells at x and at x w_N (anything indexable: an array on the host, a loader over the LDE table on the
// device, so that a cell is fetched where it is used and only one lane's cells are live at a time); (gx, gy): lane A's
// periodic point at x; (fm, fk): the public columns at x (the polynomials through the message hashes / the keys' x over the
// block-start rows); iz[6]: inverse zerofiers of the groups.  Returns the composition value (Montgomery, canonical).
template <class Cells, class Zerofiers>
SPG_HD Fp ecdsa_air_point(const Cells& c, const Cells& n, const Fp& gx, const Fp& gy, const Fp& fm, const Fp& fk,
                          const EcdsaAirConsts& K, const Zerofiers& iz) {
  const Fp* a = K.alpha;
  const Fp one = fp_one();
  // lane A                                                                                   alpha 0 .. 9
  Fp step = eair_lane(a, c[EA_AM], n[EA_AM], c[EA_APX], c[EA_APY], n[EA_APX], n[EA_APY], gx, gy, c[EA_ASA], c[EA_AI]);
  Fp hold = fp_add(fp_mul(a[5], fp_sub(n[EA_APX], c[EA_APX])), fp_mul(a[6], fp_sub(n[EA_APY], c[EA_APY])));
  Fp zero = fp_mul(a[7], c[EA_AM]);
  Fp first = fp_add(fp_mul(a[8], fp_sub(c[EA_APX], K.shift_x)), fp_mul(a[9], fp_sub(c[EA_APY], K.minus_shift_y)));
  // lanes B (alpha 10 .. 23) and C (alpha 24 .. 36): the same eight columns and the same constraints at the same relative
  // alpha positions, so ONE copy of the code serves both (the kernel is fetch-bound when everything is unrolled: a field
  // multiplication is ~190 instructions and a point needs ~100 of them)
#pragma unroll 1
  for (int l = 0; l < 2; l++) {
    const int cb = l ? EA_CM : EA_BM;
    const Fp* al = a + (l ? 24 : 10);
    const Fp M = c[cb], PX = c[cb + 1], PY = c[cb + 2], QX = c[cb + 3], QY = c[cb + 4];
    const Fp PXn = n[cb + 1], PYn = n[cb + 2];
    step = fp_add(step, eair_lane(al, M, n[cb], PX, PY, PXn, PYn, QX, QY, c[cb + 5], c[cb + 7]));
    step = fp_add(step, eair_double(al + 5, QX, QY, n[cb + 3], n[cb + 4], c[cb + 6]));
    hold = fp_add(hold, fp_add(fp_mul(al[8], fp_sub(PXn, PX)), fp_mul(al[9], fp_sub(PYn, PY))));
    zero = fp_add(zero, fp_mul(al[10], M));
    first = fp_add(first, fp_add(fp_mul(al[11], fp_sub(PX, K.shift_x)), fp_mul(al[12], fp_sub(PY, K.shift_y))));
    if (l == 0) {                                     // the key is a curve point (lane B's row 0)
      const Fp curve = fp_sub(fp_sub(fp_sub(fp_sqr(QY), fp_mul(fp_sqr(QX), QX)), QX), K.beta);
      first = fp_add(first, fp_mul(al[13], curve));
    }
  }
  // row 255: ec_add(zG, rQ) becomes lane C's point on the next row; ec_add(wB, -S).x == r      alpha 37 .. 44
  const Fp dab = fp_sub(c[EA_APX], c[EA_BPX]);
  Fp last = fp_add(fp_mul(a[37], fp_sub(fp_mul(c[EA_AI], dab), one)),
                   fp_mul(a[38], fp_sub(fp_mul(c[EA_ASA], dab), fp_sub(c[EA_APY], c[EA_BPY]))));
  last = fp_add(last, fp_mul(a[39], fp_sub(n[EA_CQX], fp_sub(fp_sub(fp_sqr(c[EA_ASA]), c[EA_APX]), c[EA_BPX]))));
  last = fp_add(last, fp_mul(a[40], fp_sub(n[EA_CQY], fp_sub(fp_mul(c[EA_ASA], fp_sub(c[EA_APX], n[EA_CQX])), c[EA_APY]))));
  const Fp dcs = fp_sub(c[EA_CPX], K.shift_x);
  last = fp_add(last, fp_add(fp_mul(a[41], fp_sub(fp_mul(c[EA_CI], dcs), one)),
                             fp_mul(a[42], fp_sub(fp_mul(c[EA_CSA], dcs), fp_add(c[EA_CPY], K.shift_y)))));
  last = fp_add(last, fp_mul(a[43], fp_sub(fp_sub(fp_sub(fp_sqr(c[EA_CSA]), c[EA_CPX]), K.shift_x), c[EA_T2])));
  last = fp_add(last, fp_mul(a[44], fp_sub(n[EA_T2], c[EA_T1])));
  // row 0: r into its carrier; scalars non-zero (`assert 0 < m`, signature.py:180)              alpha 45 .. 47
  first = fp_add(first, fp_mul(a[45], fp_sub(c[EA_T1], c[EA_BM])));
  first = fp_add(first, fp_mul(a[46], fp_sub(fp_mul(fp_mul(c[EA_V1], c[EA_AM]), c[EA_BM]), one)));
  first = fp_add(first, fp_mul(a[47], fp_sub(fp_mul(c[EA_V2], c[EA_CM]), one)));
  // carriers hold inside a block                                                               alpha 48, 49
  const Fp thold = fp_add(fp_mul(a[48], fp_sub(n[EA_T1], c[EA_T1])), fp_mul(a[49], fp_sub(n[EA_T2], c[EA_T2])));
  // public input on the block-start rows: the message hash and the key's x                     alpha 50, 51
  first = fp_add(first, fp_add(fp_mul(a[50], fp_sub(c[EA_AM], fm)), fp_mul(a[51], fp_sub(c[EA_BQX], fk))));
  Fp acc = fp_mul(step, iz[0]);
  acc = fp_add(acc, fp_mul(hold, iz[1]));
  acc = fp_add(acc, fp_mul(zero, iz[2]));
  acc = fp_add(acc, fp_mul(first, iz[3]));
  acc = fp_add(acc, fp_mul(last, iz[4]));
  acc = fp_add(acc, fp_mul(thold, iz[5]));
  return fp_reduce(acc);
}

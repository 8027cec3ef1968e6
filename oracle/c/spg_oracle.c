/* Plain-C CPU restatement of the prover-stage arithmetic (field, NTT, LDE, ...).
 * TEST INFRASTRUCTURE ONLY: used by tests/ as a faster checker than the Python oracle and by
 * bench.py's cpu_baseline / --impl reference legs.  PARITY UNPINNED for these stages: the
 * reference repository contains no prover; the field p and generator 3 are the
 * reference's (signature.py:41-42).  Validated against oracle/ntt.py in tests/test_oracle_ntt.py.
 *
 * Build: oracle/Makefile  ->  oracle/_build/libspg_oracle.so   (gcc -O3 -fopenmp)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;   /* little-endian limbs */

static const fe P = {{1ull, 0ull, 0ull, 0x0800000000000011ull}};
/* R = 2^256 */
static const fe R1 = {{0xffffffffffffffe1ull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0x07fffffffffffdf0ull}};
static const fe R2 = {{0xfffffd737e000401ull, 0x00000001330fffffull, 0xffffffffff6f8000ull, 0x07ffd4ab5e008810ull}};

static inline int geq(const fe* a, const fe* b) {
  for (int i = 3; i >= 0; i--) if (a->l[i] != b->l[i]) return a->l[i] > b->l[i];
  return 1;
}
static inline void sub_raw(fe* r, const fe* a, const fe* b) {
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) { u128 d = (u128)a->l[i] - b->l[i] - br; r->l[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
}
static inline void add_raw(fe* r, const fe* a, const fe* b) {
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) { u128 s = (u128)a->l[i] + b->l[i] + c; r->l[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
}
static inline void fe_add(fe* r, const fe* a, const fe* b) { add_raw(r, a, b); if (geq(r, &P)) sub_raw(r, r, &P); }
static inline void fe_sub(fe* r, const fe* a, const fe* b) {
  if (geq(a, b)) sub_raw(r, a, b); else { fe t; add_raw(&t, a, &P); sub_raw(r, &t, b); }
}
/* Montgomery product a*b/R mod p; -p^-1 mod 2^64 = -1 so m = -t[k] */
static inline void fe_mul(fe* r, const fe* a, const fe* b) {
  uint64_t t[9] = {0};
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) { u128 s = (u128)a->l[j] * b->l[i] + t[i + j] + c; t[i + j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    t[i + 4] = c;
  }
  for (int k = 0; k < 4; k++) {
    uint64_t m = (uint64_t)0 - t[k];
    /* t += m * p << 64k ; p = 1 + p3 * 2^192 */
    u128 s = (u128)t[k] + m; t[k] = (uint64_t)s; uint64_t c = (uint64_t)(s >> 64);
    for (int j = 1; j < 3; j++) { s = (u128)t[k + j] + c; t[k + j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    s = (u128)m * P.l[3] + t[k + 3] + c; t[k + 3] = (uint64_t)s; c = (uint64_t)(s >> 64);
    for (int j = k + 4; c && j < 9; j++) { s = (u128)t[j] + c; t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
  }
  fe o = {{t[4], t[5], t[6], t[7]}};
  if (geq(&o, &P)) sub_raw(&o, &o, &P);
  *r = o;
}
static void fe_pow(fe* r, const fe* a, const uint64_t e[4]) {
  fe acc = R1;
  for (int i = 255; i >= 0; i--) {
    fe_mul(&acc, &acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a);
  }
  *r = acc;
}
static void fe_inv(fe* r, const fe* a) {
  uint64_t e[4] = {0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0x0800000000000010ull};
  fe_pow(r, a, e);
}
static void to_mont(fe* r, const fe* a) { fe_mul(r, a, &R2); }
static void from_mont(fe* r, const fe* a) { fe one = {{1, 0, 0, 0}}; fe_mul(r, a, &one); }
static void root_of_unity(fe* r, int log_n) {   /* Montgomery form */
  fe three = {{3, 0, 0, 0}}, g;
  to_mont(&g, &three);
  uint64_t e[4] = {0, 0, 0, 0};
  int bits[3] = {251 - log_n, 196 - log_n, 192 - log_n};
  for (int i = 0; i < 3; i++) e[bits[i] >> 6] |= 1ull << (bits[i] & 63);
  fe_pow(r, &g, e);
}
static unsigned bitrev(unsigned x, int bits) {
  unsigned r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

/* out[i] = a[i] * b[i] mod p on canonical values (count field multiplications for the baseline) */
void spgo_mul_batch(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    fe x, y, z;
    memcpy(&x, a + 4 * i, 32); memcpy(&y, b + 4 * i, 32);
    to_mont(&x, &x); fe_mul(&z, &x, &y);   /* (xR)(y)/R = xy */
    memcpy(out + 4 * i, &z, 32);
  }
}

/* In-place DIF NTT core on Montgomery-or-canonical data (linear), natural in -> bit-reversed out.
 * tw: n/2 twiddles w^j in Montgomery form. */
static void dif_core(fe* a, size_t n, const fe* tw) {
  for (size_t h = n / 2; h >= 1; h /= 2) {
    size_t stride = n / (2 * h);
    for (size_t b = 0; b < n; b += 2 * h)
      for (size_t j = 0; j < h; j++) {
        fe u = a[b + j], v = a[b + j + h], d;
        fe_add(&a[b + j], &u, &v);
        fe_sub(&d, &u, &v);
        if (j) fe_mul(&a[b + j + h], &d, &tw[j * stride]); else a[b + j + h] = d;
      }
  }
}
static fe* make_twiddles(int log_n, int inverse) {
  size_t n = (size_t)1 << log_n, half = n > 1 ? n / 2 : 1;
  fe* tw = (fe*)malloc(half * sizeof(fe));
  fe w; root_of_unity(&w, log_n);
  if (inverse) fe_inv(&w, &w);
  tw[0] = R1;
  for (size_t i = 1; i < half; i++) fe_mul(&tw[i], &tw[i - 1], &w);
  return tw;
}

/* batch of `batch` vectors of 2^log_n canonical felts, in place.
 * order: 0 natural->bitrev, 1 bitrev->natural, 2 natural->natural.  inverse scales by 1/n. */
void spgo_ntt(uint64_t* data, unsigned log_n, size_t batch, int inverse, int order) {
  size_t n = (size_t)1 << log_n;
  fe* tw = make_twiddles((int)log_n, inverse);
  fe ninv, nn = {{(uint64_t)n, 0, 0, 0}};
  to_mont(&nn, &nn); fe_inv(&ninv, &nn);
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < batch; c++) {
    fe* a = (fe*)(data + 4 * n * c);
    fe* tmp = NULL;
    if (order == 1) {   /* bring to natural order first */
      tmp = (fe*)malloc(n * sizeof(fe));
      for (size_t i = 0; i < n; i++) tmp[bitrev((unsigned)i, (int)log_n)] = a[i];
      memcpy(a, tmp, n * sizeof(fe));
    }
    dif_core(a, n, tw);
    if (inverse) for (size_t i = 0; i < n; i++) fe_mul(&a[i], &a[i], &ninv);
    if (order != 0) {
      if (!tmp) tmp = (fe*)malloc(n * sizeof(fe));
      for (size_t i = 0; i < n; i++) tmp[bitrev((unsigned)i, (int)log_n)] = a[i];
      memcpy(a, tmp, n * sizeof(fe));
    }
    free(tmp);
  }
  free(tw);
}

/* LDE: trace [C][N] canonical, natural order on <w_N>  ->  out [B][C][N],
 * out[j][c][i] = f_c(g * w_{BN}^j * w_N^i)  (same convention as oracle/ntt.py lde()).  g canonical. */
void spgo_lde(const uint64_t* trace, unsigned log_n, size_t C, unsigned log_blowup, const uint64_t* g_canon,
              uint64_t* out) {
  size_t n = (size_t)1 << log_n, B = (size_t)1 << log_blowup;
  fe* twi = make_twiddles((int)log_n, 1);
  fe* twf = make_twiddles((int)log_n, 0);
  fe ninv, nn = {{(uint64_t)n, 0, 0, 0}}, g, wb;
  to_mont(&nn, &nn); fe_inv(&ninv, &nn);
  memcpy(&g, g_canon, 32); to_mont(&g, &g);
  root_of_unity(&wb, (int)(log_n + log_blowup));
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < C; c++) {
    fe* coef = (fe*)malloc(n * sizeof(fe));
    fe* tmp = (fe*)malloc(n * sizeof(fe));
    fe* work = (fe*)malloc(n * sizeof(fe));
    memcpy(tmp, trace + 4 * n * c, n * sizeof(fe));
    dif_core(tmp, n, twi);
    for (size_t i = 0; i < n; i++) fe_mul(&coef[bitrev((unsigned)i, (int)log_n)], &tmp[i], &ninv);
    /* coef now canonical * (Montgomery ninv) = canonical c_k */
    fe s = g;   /* Montgomery */
    for (size_t j = 0; j < B; j++) {
      fe acc = R1;
      for (size_t k = 0; k < n; k++) { fe_mul(&work[k], &coef[k], &acc); fe_mul(&acc, &acc, &s); }
      dif_core(work, n, twf);
      fe* o = (fe*)(out + 4 * n * (j * C + c));
      for (size_t i = 0; i < n; i++) o[bitrev((unsigned)i, (int)log_n)] = work[i];
      fe_mul(&s, &s, &wb);
    }
    free(coef); free(tmp); free(work);
  }
  free(twi); free(twf);
}

/* use `n` OpenMP threads from now on (torchrun exports OMP_NUM_THREADS=1, which would starve the CPU baseline) */
void spgo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int spgo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Horner evaluation of `batch` polynomials (canonical coefficients, natural order, n each) at one canonical point each:
 * out[b] = sum_k coef[b][k] * pt[b]^k.  Used by the 2^20 parity tests for the out-of-domain values. */
void spgo_poly_eval(const uint64_t* coef, size_t n, const uint64_t* pts, size_t batch, uint64_t* out) {
#pragma omp parallel for schedule(dynamic)
  for (size_t b = 0; b < batch; b++) {
    fe x, acc = {{0, 0, 0, 0}};
    memcpy(&x, pts + 4 * b, 32); to_mont(&x, &x);
    const fe* c = (const fe*)(coef + 4 * n * b);
    for (size_t k = n; k-- > 0;) { fe_mul(&acc, &acc, &x); fe_add(&acc, &acc, &c[k]); }   /* canonical * Mont / R = canonical */
    memcpy(out + 4 * b, &acc, 32);
  }
}

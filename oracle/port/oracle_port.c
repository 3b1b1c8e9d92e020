/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's MSM / NTT / field algorithms in plain C.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product (icicle_b200/) never
 * does.  Pinned against the unmodified reference CPU backend (oracle/_ref) and the committed golden vectors by
 * tests/test_oracle.py.  Every function cites the reference code it restates (paths relative to /root/reference/icicle).
 *
 * Representation: N little-endian 32-bit limbs in canonical standard form [0,p) (include/icicle/math/storage.h:36-48).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXL 24

typedef struct {
  int n;                 /* limbs */
  int bits;              /* modulus bit count */
  uint32_t p[MAXL];      /* modulus */
  uint32_t m[MAXL];      /* floor(2^(2*bits) / p)            include/icicle/fields/params_gen.h:24-33 (get_m) */
  uint32_t b3[MAXL];     /* 3*b of y^2 = x^3 + b (curves only) include/icicle/fields/field.h:23-56 (mul_weierstrass_b<.., true>) */
} port_field;

/* ---- multi-limb helpers (include/icicle/math/host_math.h:105-170 add_sub_limbs) ---- */
static uint32_t add_n(const uint32_t* a, const uint32_t* b, uint32_t* r, int n)
{
  uint64_t c = 0;
  for (int i = 0; i < n; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)c;
}
static uint32_t sub_n(const uint32_t* a, const uint32_t* b, uint32_t* r, int n)
{
  uint64_t br = 0;
  for (int i = 0; i < n; i++) { uint64_t t = (uint64_t)a[i] - b[i] - br; r[i] = (uint32_t)t; br = (t >> 32) & 1; }
  return (uint32_t)br;
}
static int is_zero_n(const uint32_t* a, int n) { uint32_t t = 0; for (int i = 0; i < n; i++) t |= a[i]; return t == 0; }
static int eq_n(const uint32_t* a, const uint32_t* b, int n) { return memcmp(a, b, 4 * n) == 0; }

/* schoolbook product, 2n limbs out (host_math.h:209-238 multiply_raw) */
static void mul_raw(const uint32_t* a, const uint32_t* b, uint32_t* r, int n)
{
  memset(r, 0, 8 * n);
  for (int i = 0; i < n; i++) {
    uint64_t c = 0;
    for (int j = 0; j < n; j++) { c += (uint64_t)a[i] * b[j] + r[i + j]; r[i + j] = (uint32_t)c; c >>= 32; }
    r[i + n] = (uint32_t)c;
  }
}

/* r = a + b mod p (include/icicle/math/modular_arithmetic.h:354-361) */
void port_add(const port_field* f, const uint32_t* a, const uint32_t* b, uint32_t* r)
{
  uint32_t s[MAXL], t[MAXL];
  uint32_t c = add_n(a, b, s, f->n);
  uint32_t br = sub_n(s, f->p, t, f->n);
  memcpy(r, (c || !br) ? t : s, 4 * f->n);
}
/* r = a - b mod p (modular_arithmetic.h:363-369) */
void port_sub(const port_field* f, const uint32_t* a, const uint32_t* b, uint32_t* r)
{
  uint32_t d[MAXL];
  if (sub_n(a, b, d, f->n)) add_n(d, f->p, d, f->n);
  memcpy(r, d, 4 * f->n);
}
/* r = -a mod p, with -0 = 0 (modular_arithmetic.h:587-597) */
void port_neg(const port_field* f, const uint32_t* a, uint32_t* r)
{
  if (is_zero_n(a, f->n)) { memset(r, 0, 4 * f->n); return; }
  sub_n(f->p, a, r, f->n);
}

/* r = a*b mod p: multiply_raw + multi-precision Barrett (host_math.h:437-470, modular_arithmetic.h:401-406,517-521):
 *   k = 2*bits - 32n ; l = ((xs >> k) * m) >> 32n ; r = xs - l*p (low n limbs) ; then at most two subtractions of p. */
void port_mul(const port_field* f, const uint32_t* a, const uint32_t* b, uint32_t* r)
{
  const int n = f->n;
  uint32_t xs[2 * MAXL], hi[MAXL + 1], l[2 * MAXL], lp[2 * MAXL], t[MAXL];
  mul_raw(a, b, xs, n);
  const int k = 2 * f->bits - 32 * n; /* >= 0 for every supported field except when bits <= 16n: handled by k < 0 branch */
  if (k >= 0) {
    const int ws = k / 32, bs = k % 32;
    for (int i = 0; i < n; i++) {
      uint64_t lo = (ws + i < 2 * n) ? xs[ws + i] : 0, up = (ws + i + 1 < 2 * n) ? xs[ws + i + 1] : 0;
      hi[i] = bs ? (uint32_t)((lo >> bs) | (up << (32 - bs))) : (uint32_t)lo;
    }
  } else { /* single-limb fields with < 16 bits do not occur; keep the general left shift for completeness */
    const int sh = -k;
    memset(hi, 0, sizeof(hi));
    for (int i = 0; i < n; i++) {
      int src = i - sh / 32;
      uint64_t lo = (src >= 0) ? xs[src] : 0, dn = (src - 1 >= 0) ? xs[src - 1] : 0;
      hi[i] = (sh % 32) ? (uint32_t)((lo << (sh % 32)) | (dn >> (32 - sh % 32))) : (uint32_t)lo;
    }
  }
  mul_raw(hi, f->m, l, n);          /* l_hi = l[n..2n) */
  mul_raw(l + n, f->p, lp, n);      /* only the low n limbs matter */
  sub_n(xs, lp, t, n);              /* r = xs - l*p mod 2^(32n); the true remainder is < 3p < 2^(32n) */
  /* at most two reductions (params_gen.h:54-70 num_of_reductions); a third iteration is harmless */
  for (int it = 0; it < 3; it++) {
    uint32_t u[MAXL];
    if (sub_n(t, f->p, u, n)) break;
    memcpy(t, u, 4 * n);
  }
  memcpy(r, t, 4 * n);
}

/* ---- NTT: forward out[k] = sum (in[i] g^i) w^(ik); inverse out[i] = g^-i N^-1 sum in[k] w^(-ik) -------------------------
 * Restates NttCpu::run (backend/cpu/include/ntt_cpu.h:69-232): coset multiply before a forward transform (:73,316-364),
 * radix-2 decimation-in-time butterflies on bit-reversed data (ntt_task.h:1161-1238, inverse twiddle index :1220-1222,
 * N^-1 scaling :1231-1235), coset multiply after an inverse transform (ntt_cpu.h:226).
 * `w` = root of unity of order N, `w_inv`, `n_inv`, `g` (coset generator for this direction: g or g^-1) are passed in
 * standard form by the caller (Python side computes them with integers). data: N elements, in place, natural order. */
static uint32_t brev(uint32_t x, int logn) { uint32_t r = 0; for (int i = 0; i < logn; i++) r |= ((x >> i) & 1u) << (logn - 1 - i); return r; }

void port_ntt(const port_field* f, uint32_t* data, int logn, const uint32_t* w_dir, const uint32_t* n_inv_or_null, const uint32_t* g_or_null,
              int inverse)
{
  const int n = f->n;
  const uint32_t N = 1u << logn;
  uint32_t one[MAXL] = {1};
  if (!inverse && g_or_null) { /* forward coset: x[i] *= g^i */
    uint32_t gp[MAXL];
    memcpy(gp, one, 4 * n);
    for (uint32_t i = 0; i < N; i++) { port_mul(f, data + (size_t)i * n, gp, data + (size_t)i * n); port_mul(f, gp, g_or_null, gp); }
  }
  /* bit-reverse permutation (ntt_cpu.h:286-296) */
  for (uint32_t i = 0; i < N; i++) {
    uint32_t j = brev(i, logn);
    if (j > i) { uint32_t t[MAXL]; memcpy(t, data + (size_t)i * n, 4 * n); memcpy(data + (size_t)i * n, data + (size_t)j * n, 4 * n); memcpy(data + (size_t)j * n, t, 4 * n); }
  }
  /* twiddle table tw[k] = w_dir^k, k < N/2 (cpu_ntt_domain.h:102-110) */
  uint32_t* tw = (uint32_t*)malloc((size_t)(N / 2 + 1) * n * 4);
  memcpy(tw, one, 4 * n);
  for (uint32_t k = 1; k < N / 2; k++) port_mul(f, tw + (size_t)(k - 1) * n, w_dir, tw + (size_t)k * n);
  for (int s = 0; s < logn; s++) {
    const uint32_t half = 1u << s, step = N >> (s + 1);
    for (uint32_t blk = 0; blk < N; blk += 2 * half)
      for (uint32_t j = 0; j < half; j++) {
        uint32_t *u = data + (size_t)(blk + j) * n, *v = data + (size_t)(blk + j + half) * n, t[MAXL], a[MAXL];
        port_mul(f, v, tw + (size_t)(j * step) * n, t);
        port_add(f, u, t, a);
        port_sub(f, u, t, v);
        memcpy(u, a, 4 * n);
      }
  }
  free(tw);
  if (inverse) {
    uint32_t gp[MAXL];
    memcpy(gp, n_inv_or_null, 4 * n);
    for (uint32_t i = 0; i < N; i++) {
      port_mul(f, data + (size_t)i * n, gp, data + (size_t)i * n);
      if (g_or_null) port_mul(f, gp, g_or_null, gp);
    }
  }
}

/* ---- group law: complete a = 0 formulas of Renes-Costello-Batina as written in the reference ------------------------------
 * Projective {X,Y,Z}, zero = (0,1,0) (include/icicle/curves/projective.h:26). */
typedef struct { uint32_t x[MAXL], y[MAXL], z[MAXL]; } pt;

static void pt_zero(const port_field* f, pt* r) { memset(r, 0, sizeof(*r)); r->y[0] = 1; }

/* projective.h:101-143 (operator+) */
static void pt_add(const port_field* f, const pt* p1, const pt* p2, pt* out)
{
  uint32_t t00[MAXL], t01[MAXL], t02[MAXL], t03[MAXL], t04[MAXL], t05[MAXL], t06[MAXL], t07[MAXL], t08[MAXL], t09[MAXL], t10[MAXL], t11[MAXL],
    t12[MAXL], t13[MAXL], t14[MAXL], t15[MAXL], t16[MAXL], t17[MAXL], t18[MAXL], t19[MAXL], t20[MAXL], t21[MAXL], t22[MAXL], t23[MAXL], a[MAXL], b[MAXL];
  pt r;
  port_mul(f, p1->x, p2->x, t00); port_mul(f, p1->y, p2->y, t01); port_mul(f, p1->z, p2->z, t02);
  port_add(f, p1->x, p1->y, t03); port_add(f, p2->x, p2->y, t04); port_mul(f, t03, t04, t05);
  port_add(f, t00, t01, t06); port_sub(f, t05, t06, t07);
  port_add(f, p1->y, p1->z, t08); port_add(f, p2->y, p2->z, t09); port_mul(f, t08, t09, t10);
  port_add(f, t01, t02, t11); port_sub(f, t10, t11, t12);
  port_add(f, p1->x, p1->z, t13); port_add(f, p2->x, p2->z, t14); port_mul(f, t13, t14, t15);
  port_add(f, t00, t02, t16); port_sub(f, t15, t16, t17);
  port_add(f, t00, t00, t18); port_add(f, t18, t00, t19);
  port_mul(f, f->b3, t02, t20); port_add(f, t01, t20, t21); port_sub(f, t01, t20, t22); port_mul(f, f->b3, t17, t23);
  port_mul(f, t12, t23, a); port_mul(f, t07, t22, b); port_sub(f, b, a, r.x);
  port_mul(f, t23, t19, a); port_mul(f, t22, t21, b); port_add(f, b, a, r.y);
  port_mul(f, t19, t07, a); port_mul(f, t21, t12, b); port_add(f, b, a, r.z);
  *out = r;
}
/* projective.h:73-99 (dbl) */
static void pt_dbl(const port_field* f, const pt* p, pt* out)
{
  uint32_t t0[MAXL], t1[MAXL], t2[MAXL], x3[MAXL], y3[MAXL], z3[MAXL];
  pt r;
  port_mul(f, p->y, p->y, t0);
  port_add(f, t0, t0, z3); port_add(f, z3, z3, z3); port_add(f, z3, z3, z3);
  port_mul(f, p->y, p->z, t1);
  port_mul(f, p->z, p->z, t2); port_mul(f, f->b3, t2, t2);
  port_mul(f, t2, z3, x3); port_add(f, t0, t2, y3); port_mul(f, t1, z3, z3);
  port_add(f, t2, t2, t1); port_add(f, t1, t2, t2); port_sub(f, t0, t2, t0);
  port_mul(f, t0, y3, y3); port_add(f, x3, y3, y3);
  port_mul(f, p->x, p->y, t1); port_mul(f, t0, t1, x3); port_add(f, x3, x3, x3);
  memcpy(r.x, x3, sizeof(x3)); memcpy(r.y, y3, sizeof(y3)); memcpy(r.z, z3, sizeof(z3));
  *out = r;
}

/* bits [lsb, lsb+width) of a scalar (modular_arithmetic.h:292-301 get_scalar_bits) */
static uint32_t scalar_bits(const uint32_t* s, int limbs, int lsb, int width)
{
  if (width <= 0) return 0;
  int li = lsb / 32, off = lsb % 32;
  uint64_t two = (li < limbs) ? s[li] : 0;
  if (li + 1 < limbs) two |= (uint64_t)s[li + 1] << 32;
  return (uint32_t)((two >> off) & ((width >= 32) ? 0xffffffffu : ((1u << width) - 1)));
}

/* Pippenger MSM, single worker (backend/cpu/src/curve/cpu_msm.hpp):
 *   parameters                         calc_optimal_parameters :199-223
 *   signed-digit bucket population     worker_run_phase1 :259-314 (negate trick :276-277, digit/carry :289-309, zero bases :282)
 *   line/triangle bucket sums          worker_collapse_segment :330-352
 *   Horner over bucket modules         phase3_final_accumulator :409-416
 * scalars: n x s_limbs standard form, `scalar_bits_total` = scalar field NBITS, `r_mod` = scalar field modulus (for the negate
 * trick); bases: n x 2 x fq.n affine; bitsize 0 = full.  Writes one projective point. */
void port_msm(const port_field* fq, const uint32_t* scalars, int s_limbs, int s_nbits, const uint32_t* r_mod, const uint32_t* bases, int n,
              int c, int bitsize, uint32_t* out)
{
  const int L = fq->n;
  const int scalar_size = bitsize ? bitsize : s_nbits;
  const int chopped = (scalar_size != s_nbits);
  const int size_with_carry = chopped ? scalar_size + 1 : scalar_size;
  const int nbm = (size_with_carry - 1) / c + 1;
  const uint32_t bm_size = 1u << (c - 1);
  const size_t total = (size_t)nbm * bm_size;
  pt* buckets = (pt*)malloc(total * sizeof(pt));
  uint8_t* busy = (uint8_t*)calloc(total, 1);
  for (int i = 0; i < n; i++) {
    uint32_t s[MAXL];
    memcpy(s, scalars + (size_t)i * s_limbs, 4 * s_limbs);
    int negate = (!chopped) && scalar_bits(s, s_limbs, scalar_size - 1, 1);
    if (negate) { uint32_t t[MAXL]; if (!is_zero_n(s, s_limbs)) { sub_n(r_mod, s, t, s_limbs); memcpy(s, t, 4 * s_limbs); } }
    const uint32_t* bx = bases + (size_t)i * 2 * L;
    if (is_zero_n(bx, 2 * L)) continue;
    pt base, base_neg;
    memset(&base, 0, sizeof(base));
    memcpy(base.x, bx, 4 * L); memcpy(base.y, bx + L, 4 * L); base.z[0] = 1;
    base_neg = base;
    port_neg(fq, base.y, base_neg.y);
    uint32_t carry = 0;
    int offset = 0;
    for (int bm = 0; bm < nbm; bm++) {
      int width = c < scalar_size - offset ? c : scalar_size - offset;
      uint32_t coeff = scalar_bits(s, s_limbs, offset, width) + carry;
      if ((coeff & ((1u << c) - 1)) != 0) {
        carry = coeff > bm_size;
        size_t idx = (size_t)bm_size * bm + (carry ? ((0u - coeff) & (bm_size - 1)) : (coeff & (bm_size - 1)));
        const pt* add = ((negate ^ (carry > 0)) ? &base_neg : &base);
        if (busy[idx]) pt_add(fq, &buckets[idx], add, &buckets[idx]);
        else { busy[idx] = 1; buckets[idx] = *add; }
      } else {
        carry = coeff >> c;
      }
      offset += c;
    }
  }
  /* per module: sum_k k*B_k with bucket index 0 holding the digit 2^(c-1) (cpu_msm.hpp:333-336) */
  pt result;
  pt_zero(fq, &result);
  for (int bm = nbm - 1; bm >= 0; bm--) {
    pt line, tri;
    pt_zero(fq, &line); pt_zero(fq, &tri);
    size_t base_i = (size_t)bm * bm_size;
    if (busy[base_i]) { line = buckets[base_i]; tri = line; }
    for (int64_t k = bm_size - 1; k > 0; k--) {
      if (busy[base_i + k]) pt_add(fq, &line, &buckets[base_i + k], &line);
      pt_add(fq, &tri, &line, &tri);
    }
    if (bm != nbm - 1) for (int j = 0; j < c; j++) pt_dbl(fq, &result, &result);
    pt_add(fq, &result, &tri, &result);
  }
  memcpy(out, result.x, 4 * L); memcpy(out + L, result.y, 4 * L); memcpy(out + 2 * L, result.z, 4 * L);
  free(buckets); free(busy);
}

/* element-wise helpers for the vec-ops parity tests (backend/cpu/src/field/cpu_vec_ops.cpp:100-140) */
void port_vec_op(const port_field* f, int op, const uint32_t* a, const uint32_t* b, uint64_t n, uint32_t* out)
{
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t *x = a + i * f->n, *y = b + i * f->n;
    uint32_t* r = out + i * f->n;
    if (op == 0) port_add(f, x, y, r);
    else if (op == 1) port_sub(f, x, y, r);
    else port_mul(f, x, y, r);
  }
}

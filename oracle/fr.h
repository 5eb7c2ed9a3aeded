/*
 * oracle/fr.h — TEST INFRASTRUCTURE (CPU oracle). Not linked into the product.
 *
 * Plain-C restatement of the BN254 scalar field element the reference calls FQ
 * (src/zkevm_specs/util/arithmetic.py:41-63; arithmetic itself lives in the absent
 * third-party py-ecc==6.0.0, py_ecc/fields/field_elements.py::FQ — restated from its
 * published semantics: values are ints reduced mod p, inv by extended Euclid with
 * inv(0)==0).  Pinned against Python big-int arithmetic in tests/test_oracle_fr.py.
 *
 * A cell is 4 little-endian uint64 limbs holding the canonical value (< p).
 */
#ifndef ORACLE_FR_H
#define ORACLE_FR_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr_t;

static const fr_t FR_P = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                           0x30644e72e131a029ull}};
static const fr_t FR_R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull,
                            0x0216d0b17f4e44a5ull}};
#define FR_N0 0xc2e1f593efffffffull

static inline fr_t fr_load(const uint64_t* p) { fr_t r; memcpy(r.l, p, 32); return r; }
static inline fr_t fr_u64(uint64_t v) { fr_t r = {{v, 0, 0, 0}}; return r; }
static inline fr_t fr_u128(uint64_t lo, uint64_t hi) { fr_t r = {{lo, hi, 0, 0}}; return r; }
static inline int fr_eq(fr_t a, fr_t b) {
  return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3];
}
static inline int fr_is_zero(fr_t a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
/* FQ == int  (py_ecc compares .n to the raw int) */
static inline int fr_eq_u64(fr_t a, uint64_t v) {
  return a.l[0] == v && (a.l[1] | a.l[2] | a.l[3]) == 0;
}
/* integer compare of .n : -1,0,1 */
static inline int fr_cmp(fr_t a, fr_t b) {
  for (int i = 3; i >= 0; i--) {
    if (a.l[i] < b.l[i]) return -1;
    if (a.l[i] > b.l[i]) return 1;
  }
  return 0;
}
/* .n < 2^bits  (bits <= 256) */
static inline int fr_fits_bits(fr_t a, int bits) {
  for (int i = 0; i < 4; i++) {
    int lo = 64 * i;
    if (bits <= lo) { if (a.l[i]) return 0; }
    else if (bits < lo + 64) { if (a.l[i] >> (bits - lo)) return 0; }
  }
  return 1;
}
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t* c) {
  u128 t = (u128)a + b + *c; *c = (uint64_t)(t >> 64); return (uint64_t)t;
}
static inline uint64_t sbb(uint64_t a, uint64_t b, uint64_t* br) {
  u128 t = (u128)a - b - *br; *br = (uint64_t)(t >> 64) & 1; return (uint64_t)t;
}
static inline fr_t fr_add(fr_t a, fr_t b) {
  fr_t s, d; uint64_t c = 0, br = 0;
  for (int i = 0; i < 4; i++) s.l[i] = adc(a.l[i], b.l[i], &c);
  for (int i = 0; i < 4; i++) d.l[i] = sbb(s.l[i], FR_P.l[i], &br);
  /* a,b < p < 2^254 so no carry out of 256 bits; subtract p if s >= p */
  return br ? s : d;
}
static inline fr_t fr_sub(fr_t a, fr_t b) {
  fr_t d; uint64_t br = 0, c = 0;
  for (int i = 0; i < 4; i++) d.l[i] = sbb(a.l[i], b.l[i], &br);
  if (br) for (int i = 0; i < 4; i++) d.l[i] = adc(d.l[i], FR_P.l[i], &c);
  return d;
}
static inline fr_t fr_neg(fr_t a) { fr_t z = {{0, 0, 0, 0}}; return fr_sub(z, a); }
/* Montgomery product a*b/2^256 mod p (CIOS) */
static inline fr_t fr_montmul(fr_t a, fr_t b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      u128 x = (u128)a.l[j] * b.l[i] + t[j] + c; t[j] = (uint64_t)x; c = (uint64_t)(x >> 64);
    }
    u128 y = (u128)t[4] + c; t[4] = (uint64_t)y; t[5] = (uint64_t)(y >> 64);
    uint64_t m = t[0] * FR_N0;
    u128 x = (u128)m * FR_P.l[0] + t[0]; c = (uint64_t)(x >> 64);
    for (int j = 1; j < 4; j++) {
      x = (u128)m * FR_P.l[j] + t[j] + c; t[j - 1] = (uint64_t)x; c = (uint64_t)(x >> 64);
    }
    y = (u128)t[4] + c; t[3] = (uint64_t)y; t[4] = t[5] + (uint64_t)(y >> 64);
  }
  fr_t r = {{t[0], t[1], t[2], t[3]}}, d; uint64_t br = 0;
  for (int i = 0; i < 4; i++) d.l[i] = sbb(r.l[i], FR_P.l[i], &br);
  return (t[4] || !br) ? d : r;
}
/* canonical a*b mod p */
static inline fr_t fr_mul(fr_t a, fr_t b) { return fr_montmul(fr_montmul(a, b), FR_R2); }
/* reduce an arbitrary 256-bit integer mod p (FQ(int) constructor) */
static inline fr_t fr_reduce256(fr_t a) {
  for (;;) { /* 2^256 / p < 6 */
    fr_t d; uint64_t br = 0;
    for (int i = 0; i < 4; i++) d.l[i] = sbb(a.l[i], FR_P.l[i], &br);
    if (br) return a;
    a = d;
  }
}
static inline fr_t fr_pow(fr_t a, const fr_t e) {
  fr_t r = fr_u64(1);
  for (int i = 255; i >= 0; i--) {
    r = fr_mul(r, r);
    if ((e.l[i / 64] >> (i % 64)) & 1) r = fr_mul(r, a);
  }
  return r;
}
/* prime_field_inv: inv(0) == 0 (py_ecc/utils.py); a^(p-2) gives the same values */
static inline fr_t fr_inv(fr_t a) {
  fr_t e = FR_P; e.l[0] -= 2;
  return fr_pow(a, e);
}
#endif

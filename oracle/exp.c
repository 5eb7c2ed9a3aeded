/*
 * oracle/exp.c — TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
 *
 * Restates verify_step / verify_exp_circuit, /root/reference/src/zkevm_specs/exp_circuit.py:14-97,
 * with mul_add_words of util/arithmetic.py:245-276 (field arithmetic, carries = (...)/2^128).
 * Row = 21 cells in the order of ExpCircuitRow (evm_circuit/table.py:519-535): q_usable, is_step,
 * identifier, is_last, base, exponent, exponentiation, a, b, c, d, q (each lo, hi), r.
 * A row stops at its first failing constraint.  Pinned by tests/golden/exp.npz.
 */
#include "common.h"

enum { X_USABLE, X_STEP, X_ID, X_LAST, X_BASE, X_EXPONENT = 6, X_EXPN = 8, X_A = 10, X_B = 12, X_C = 14, X_D = 16,
       X_Q = 18, X_R = 20, EXP_COLS };
typedef struct { fr_t lo, hi; } xword;
static const fr_t X_INV_2_128 = {{0x18ee753c76f9dc6full, 0x54ad7e14a329e70full, 0x2b16366f4f7684dfull,
                                  0x133100d71fdf3579ull}};
static int xw_domain(xword w) { return fr_fits_bits(w.lo, 128) && fr_fits_bits(w.hi, 128); }
static int xw_eq(xword a, xword b) { return fr_eq(a.lo, b.lo) && fr_eq(a.hi, b.hi); }
static int gate2(fr_t cond, xword a, xword b) {
  return fr_is_zero(fr_mul(cond, fr_sub(a.lo, b.lo))) && fr_is_zero(fr_mul(cond, fr_sub(a.hi, b.hi)));
}
static int gate1(fr_t cond, fr_t e) { return fr_is_zero(fr_mul(cond, e)); }
static int is01(fr_t v) { return fr_eq_u64(v, 0) || fr_eq_u64(v, 1); }
/* carries of mul_add_words(a, b, c, d) (a, b in the 128-bit-halves domain) */
static void mul_add(xword a, xword b, xword c, xword d, fr_t* clo, fr_t* chi) {
  fr_t a64[4] = {fr_u64(a.lo.l[0]), fr_u64(a.lo.l[1]), fr_u64(a.hi.l[0]), fr_u64(a.hi.l[1])};
  fr_t b64[4] = {fr_u64(b.lo.l[0]), fr_u64(b.lo.l[1]), fr_u64(b.hi.l[0]), fr_u64(b.hi.l[1])};
#define M(x, y) fr_mul(a64[x], b64[y])
  const fr_t t0 = M(0, 0), t1 = fr_add(M(0, 1), M(1, 0));
  const fr_t t2 = fr_add(fr_add(M(0, 2), M(1, 1)), M(2, 0));
  const fr_t t3 = fr_add(fr_add(fr_add(M(0, 3), M(1, 2)), M(2, 1)), M(3, 0));
#undef M
  const fr_t two64 = {{0, 1, 0, 0}};
  *clo = fr_mul(fr_sub(fr_add(fr_add(t0, fr_mul(t1, two64)), c.lo), d.lo), X_INV_2_128);
  *chi = fr_mul(fr_sub(fr_add(fr_add(fr_add(t2, fr_mul(t3, two64)), c.hi), *clo), d.hi), X_INV_2_128);
}

int orc_check_exp(const uint64_t* rows, uint64_t n_rows, uint64_t row_begin, uint64_t row_end, uint32_t* first_fail,
                  uint64_t* fail_count) {
  orc_result res_, *res = &res_; orc_result_init(res, first_fail, fail_count, XP_N_CONSTRAINTS);
#define XK(id, cond) do { if (!(cond)) { orc_fail(res, (id), i); goto next_row; } } while (0)
#define C(c) fr_load(ORC_CELL(rows, n_rows, c, i))
#define N(c) fr_load(ORC_CELL(rows, n_rows, c, j))
#define CW(c) ((xword){C(c), C((c) + 1)})
#define NW(c) ((xword){N(c), N((c) + 1)})
  for (uint64_t i = row_begin; i < row_end; i++) {
    const uint64_t j = (i + 1) % n_rows;
    const fr_t one = fr_u64(1), is_step = C(X_STEP), is_last = C(X_LAST), r = C(X_R);
    fr_t cond = fr_mul(is_step, fr_sub(one, is_last));
    XK(XP_BASE_SAME, gate2(cond, CW(X_BASE), NW(X_BASE)));
    XK(XP_A_EQ_NEXT_D, gate2(cond, CW(X_A), NW(X_D)));
    XK(XP_ID_SAME, gate1(cond, fr_sub(C(X_ID), N(X_ID))));
    XK(XP_LAST_BOOL, is01(fr_mul(is_step, is_last)));
    XK(XP_R_BOOL, is01(fr_mul(is_step, r)));
    {
      fr_t clo, chi;
      XK(XP_MUL_TO64, xw_domain(CW(X_A)) && xw_domain(CW(X_B)));
      mul_add(CW(X_A), CW(X_B), CW(X_C), CW(X_D), &clo, &chi);
      XK(XP_MUL_CARRY_LO, fr_fits_bits(clo, 72));
      XK(XP_MUL_CARRY_HI, fr_fits_bits(chi, 72));
    }
    XK(XP_EXP_EQ_D, gate2(is_step, CW(X_EXPN), CW(X_D)));
    XK(XP_C_ZERO, gate1(is_step, C(X_C)) && gate1(is_step, C(X_C + 1)));
    {
      fr_t clo, chi;
      XK(XP_PAR_R_WORD, fr_fits_bits(r, 128));
      XK(XP_PAR_TO64, xw_domain(CW(X_Q)));
      const xword two = {fr_u64(2), fr_u64(0)}, rw = {r, fr_u64(0)};
      mul_add(two, CW(X_Q), rw, CW(X_EXPONENT), &clo, &chi);
      XK(XP_PAR_CARRY_LO, fr_fits_bits(clo, 72));
      XK(XP_PAR_CARRY_HI, fr_fits_bits(chi, 72));
    }
    cond = fr_mul(fr_mul(is_step, fr_sub(one, is_last)), r);
    XK(XP_ODD_NEXT_LO, gate1(cond, fr_sub(N(X_EXPONENT), fr_sub(C(X_EXPONENT), one))));
    XK(XP_ODD_NEXT_HI, gate1(cond, fr_sub(N(X_EXPONENT + 1), C(X_EXPONENT + 1))));
    XK(XP_ODD_B_BASE, gate2(cond, CW(X_BASE), CW(X_B)));
    cond = fr_mul(fr_mul(is_step, fr_sub(one, is_last)), fr_sub(one, r));
    XK(XP_EVEN_NEXT_LO, gate1(cond, fr_sub(N(X_EXPONENT), C(X_Q))));
    XK(XP_EVEN_NEXT_HI, gate1(cond, fr_sub(N(X_EXPONENT + 1), C(X_Q + 1))));
    XK(XP_EVEN_A_EQ_B, gate2(cond, CW(X_A), CW(X_B)));
    XK(XP_LAST_EXP_LO2, gate1(is_last, fr_sub(C(X_EXPONENT), fr_u64(2))));
    XK(XP_LAST_EXP_HI0, gate1(is_last, C(X_EXPONENT + 1)));
    XK(XP_LAST_A_BASE, gate2(is_last, CW(X_BASE), CW(X_A)));
    XK(XP_LAST_B_BASE, gate2(is_last, CW(X_BASE), CW(X_B)));
  next_row:;
  }
  return 0;
}

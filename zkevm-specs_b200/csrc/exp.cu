// exp.cu — exponentiation-circuit row checker (one thread per row).
//
// Replaces the loop of verify_exp_circuit and its verify_step,
// src/zkevm_specs/exp_circuit.py:14-97.  Row = 21 cells in the order of ExpCircuitRow
// (evm_circuit/table.py:519-535): q_usable, is_step, identifier, is_last, base, exponent,
// exponentiation, a, b, c, d, q (lo, hi each), r; rotation {0,+1}.  Algorithmic bytes:
// 21 x 32 B = 672 B per row.  Two mul_add_words gates per row (words.cuh), evaluated for EVERY
// row like the reference (they are not under cs.condition); the other gates are cond*expr==0,
// decided without multiplication (Fr is an integral domain).  A row stops at its first failure.
#include "circuit.cuh"
#include "words.cuh"
#include "../../include/zk_constraints.h"
#include "../../include/zkcheck.h"

namespace zk {

enum { X_USABLE, X_STEP, X_ID, X_LAST, X_BASE, X_EXPONENT = 6, X_EXPN = 8, X_A = 10, X_B = 12, X_C = 14, X_D = 16,
       X_Q = 18, X_R = 20 };

#define XP_CHECK(id, cond)        \
  do {                            \
    if (!(cond)) {                \
      fail(res, (id), row);       \
      return;                     \
    }                             \
  } while (0)

template <int LAYOUT>
ZK_HD Word2 wword(const WitnessDev& w, u32 col, u64 row) { return Word2{wcell_l<LAYOUT>(w, col, row), wcell_l<LAYOUT>(w, col + 1, row)}; }
// is_step * v in {0, 1}  <=>  is_step == 0, or v == 0, or the field product equals 1
ZK_HD bool gated_bool(const Fr& gate, const Fr& v) {
  if (fr_is_zero(gate) || fr_is_zero(v)) return true;
  if (fr_eq_u64(gate, 1)) return fr_eq_u64(v, 1);
  return fr_eq_u64(fr_mul(gate, v), 1);
}

template <int LAYOUT>
ZK_HD void check_exp_row(const WitnessDev& w, const CheckRange& rg, const ResultDev& res, u64 i) {
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 j = rot_fwd(w, i, 1, wrap);
  const u64 row = rg.row_base + i;
  const Fr is_step = wcell_l<LAYOUT>(w, X_STEP, i), is_last = wcell_l<LAYOUT>(w, X_LAST, i), r = wcell_l<LAYOUT>(w, X_R, i);
  const Word2 base = wword<LAYOUT>(w, X_BASE, i), expo = wword<LAYOUT>(w, X_EXPONENT, i), a = wword<LAYOUT>(w, X_A, i), b = wword<LAYOUT>(w, X_B, i);
  const Word2 c = wword<LAYOUT>(w, X_C, i), d = wword<LAYOUT>(w, X_D, i), q = wword<LAYOUT>(w, X_Q, i);
  const Word2 n_expo = wword<LAYOUT>(w, X_EXPONENT, j);
  const bool step0 = fr_is_zero(is_step), last1 = fr_eq_u64(is_last, 1), last0 = fr_is_zero(is_last);
  // cond = is_step * (1 - is_last)
  const bool off1 = step0 || last1;
  XP_CHECK(XP_BASE_SAME, off1 || word_eq(base, wword<LAYOUT>(w, X_BASE, j)));
  XP_CHECK(XP_A_EQ_NEXT_D, off1 || word_eq(a, wword<LAYOUT>(w, X_D, j)));
  XP_CHECK(XP_ID_SAME, off1 || fr_eq(wcell_l<LAYOUT>(w, X_ID, i), wcell_l<LAYOUT>(w, X_ID, j)));
  XP_CHECK(XP_LAST_BOOL, gated_bool(is_step, is_last));
  XP_CHECK(XP_R_BOOL, gated_bool(is_step, r));
  {
    XP_CHECK(XP_MUL_TO64, word_in_domain(a) && word_in_domain(b));
    Fr clo, chi, ovf;
    mul_add_carries(a, b, c, d, &clo, &chi, &ovf);
    XP_CHECK(XP_MUL_CARRY_LO, fits_9_bytes(clo));
    XP_CHECK(XP_MUL_CARRY_HI, fits_9_bytes(chi));
  }
  XP_CHECK(XP_EXP_EQ_D, step0 || word_eq(wword<LAYOUT>(w, X_EXPN, i), d));
  XP_CHECK(XP_C_ZERO, step0 || (fr_is_zero(c.lo) && fr_is_zero(c.hi)));
  {
    XP_CHECK(XP_PAR_R_WORD, fr_fits128(r));
    XP_CHECK(XP_PAR_TO64, word_in_domain(q));
    Fr clo, chi, ovf;
    mul_add_carries(Word2{fr_u64(2), fr_u64(0)}, q, Word2{r, fr_u64(0)}, expo, &clo, &chi, &ovf);
    XP_CHECK(XP_PAR_CARRY_LO, fits_9_bytes(clo));
    XP_CHECK(XP_PAR_CARRY_HI, fits_9_bytes(chi));
  }
  {  // cond = is_step * (1 - is_last) * r
    const bool off = off1 || fr_is_zero(r);
    XP_CHECK(XP_ODD_NEXT_LO, off || fr_eq(n_expo.lo, fr_sub_u64(expo.lo, 1)));
    XP_CHECK(XP_ODD_NEXT_HI, off || fr_eq(n_expo.hi, expo.hi));
    XP_CHECK(XP_ODD_B_BASE, off || word_eq(base, b));
  }
  {  // cond = is_step * (1 - is_last) * (1 - r)
    const bool off = off1 || fr_eq_u64(r, 1);
    XP_CHECK(XP_EVEN_NEXT_LO, off || fr_eq(n_expo.lo, q.lo));
    XP_CHECK(XP_EVEN_NEXT_HI, off || fr_eq(n_expo.hi, q.hi));
    XP_CHECK(XP_EVEN_A_EQ_B, off || word_eq(a, b));
  }
  XP_CHECK(XP_LAST_EXP_LO2, last0 || fr_eq_u64(expo.lo, 2));
  XP_CHECK(XP_LAST_EXP_HI0, last0 || fr_is_zero(expo.hi));
  XP_CHECK(XP_LAST_A_BASE, last0 || word_eq(base, a));
  XP_CHECK(XP_LAST_B_BASE, last0 || word_eq(base, b));
}

#ifdef __CUDACC__
template <int LAYOUT>
__global__ void __launch_bounds__(128) k_check_exp(WitnessDev w, CheckRange rg, ResultDev res) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = rg.row_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rg.row_end; i += stride)
    check_exp_row<LAYOUT>(w, rg, res, i);
}
#endif

}  // namespace zk

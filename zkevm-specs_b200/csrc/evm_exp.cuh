// Gate program of EXP (group KG_ARITH), part of evm.cu (included there).
//   exp   evm_circuit/execution/exp.py:5-50 (exp_lookup table.py:797-814, ExpTableRow :539-548, mul_add_words
//         instruction.py:599-632, byte_size :492-494)
// ZK_TABLE_EXP holds the exp table (11 cells: is_step, identifier, is_last, base limbs 0..3, exponent lo / hi,
// exponentiation lo / hi); the hash index is keyed on the nine cells exp_lookup names.
#pragma once
namespace zk {

ZK_HD_NOINLINE int exp_lookup(const StepCtx& s, const Fr& identifier, u64 is_last, const Word2& base, const Word2& exponent, Word2* out) {
  Fr key[9] = {fr_u64(1), identifier, fr_u64(is_last), fr_u64(base.lo.l[0]), fr_u64(base.lo.l[1]), fr_u64(base.hi.l[0]),
               fr_u64(base.hi.l[1]), exponent.lo, exponent.hi};
  u32 r = 0;
  const int n = lookup<9>(s.t.exp, key, &r);
  if (n == 1) {
    out->lo = table_cell(s.t.exp.tab, 9, r);
    out->hi = table_cell(s.t.exp.tab, 10, r);
  }
  return n;
}

ZK_HD_NOINLINE void gadget_exp(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr sp = s.cur(S_SP), sp1 = fr_add_u64(sp, 1);
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 base = zero, exponent = zero, pushed = zero;
  if (!need1(s, true, stack_at(s, true, 0, 0, sp, &base), EV_EXP_RW0_UNSAT)) return;
  if (!need1(s, true, stack_at(s, true, 1, 0, sp1, &exponent), EV_EXP_RW1_UNSAT)) return;
  if (!need1(s, true, stack_at(s, true, 2, 1, sp1, &pushed), EV_EXP_RW2_UNSAT)) return;
  const bool hi0 = fr_is_zero(exponent.hi);
  if (hi0 && fr_is_zero(exponent.lo)) {
    EV_CHECK(EV_EXP_ZERO_LO, fr_eq_u64(pushed.lo, 1));
    EV_CHECK(EV_EXP_ZERO_HI, fr_is_zero(pushed.hi));
  } else if (hi0 && fr_eq_u64(exponent.lo, 1)) {
    EV_CHECK(EV_EXP_ONE_LO, fr_eq(pushed.lo, base.lo));
    EV_CHECK(EV_EXP_ONE_HI, fr_eq(pushed.hi, base.hi));
  } else {
    EV_CHECK(EV_EXP_BASE_TO64, word_in_domain(base));  // base.to_64s()
    const Fr identifier = fr_add_u64(s.cur(S_RWC), 3);
    const bool single = hi0 && fr_eq_u64(exponent.lo, 2);
    Word2 res = zero, int_res = zero;
    TX_LK(exp_lookup(s, identifier, single ? 1 : 0, base, exponent, &res), EV_EXP_FIRST_UNSAT);
    TX_LK(exp_lookup(s, identifier, 1, base, Word2{fr_u64(2), fr_u64(0)}, &int_res), EV_EXP_LAST_UNSAT);
    // mul_add_words(base, base, Word(0), int_res): the overflow it returns is not constrained here
    Fr carry_lo, carry_hi, overflow;
    mul_add_carries(base, base, zero, int_res, &carry_lo, &carry_hi, &overflow);
    EV_CHECK(EV_EXP_CARRY_LO, fits_9_bytes(carry_lo));
    EV_CHECK(EV_EXP_CARRY_HI, fits_9_bytes(carry_hi));
    EV_CHECK(EV_EXP_RESULT, word_eq(res, pushed));
  }
  EV_CHECK(EV_EXP_EXPONENT_BYTES, word_in_domain(exponent));  // byte_size: to_le_bytes
  const u64 v[4] = {exponent.lo.l[0], exponent.lo.l[1], exponent.hi.l[0], exponent.hi.l[1]};
  u64 size = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (v[k]) {
#ifdef __CUDA_ARCH__
      size = 8 * k + (u64)(8 - (__clzll((long long)v[k]) >> 3));
#else
      size = 8 * k + (u64)(8 - (__builtin_clzll(v[k]) >> 3));
#endif
    }
  }
  same_context_x_ni(s, opcode, fr_u64(3), fr_u64(1), fr_u64(1), false, fr_u64(0), fr_u64(50 * size));
}

}  // namespace zk

/* TEST INFRASTRUCTURE — CPU oracle, part of oracle/evm.c (included there).
 * EXP: evm_circuit/execution/exp.py:5-50 (exp_lookup table.py:797-814, ExpTableRow :539-548, mul_add_words
 * instruction.py:599-632, byte_size :492-494).  The exp table (11 cells: is_step, identifier, is_last, base limbs 0..3,
 * exponent lo / hi, exponentiation lo / hi) is set for the next call with orc_set_evm_exp_table.
 * Pinned by tests/golden/evm19.npz (1,126 verdicts of the reference's verify_step).
 */
static int exp_lookup(evm_env* e, fr_t identifier, uint64_t is_last, const fr_t limbs[4], word_t exponent, word_t* out) {
  fr_t key[9] = {fr_u64(1), identifier, fr_u64(is_last), limbs[0], limbs[1], limbs[2], limbs[3], exponent.lo, exponent.hi};
  uint32_t r; const int n = orc_lookup(&e->exp_ix, key, &r);
  if (n == 1) {
    out->lo = fr_load(ORC_CELL(e->exp_ix.cells, e->exp_ix.n_rows, 9, r));
    out->hi = fr_load(ORC_CELL(e->exp_ix.cells, e->exp_ix.n_rows, 10, r));
  }
  return n;
}
static void gadget_exp(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t base, exponent, pushed;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &base), EV_EXP_RW0_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &exponent), EV_EXP_RW1_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &pushed), EV_EXP_RW2_UNSAT, row)) return;
  const int hi0 = fr_is_zero(exponent.hi);
  if (hi0 && fr_is_zero(exponent.lo)) {
    CHECK(EV_EXP_ZERO_LO, fr_eq_u64(pushed.lo, 1));
    CHECK(EV_EXP_ZERO_HI, fr_is_zero(pushed.hi));
  } else if (hi0 && fr_eq_u64(exponent.lo, 1)) {
    CHECK(EV_EXP_ONE_LO, fr_eq(pushed.lo, base.lo));
    CHECK(EV_EXP_ONE_HI, fr_eq(pushed.hi, base.hi));
  } else {
    CHECK(EV_EXP_BASE_TO64, word_in_domain(base)); /* base.to_64s() */
    const fr_t limbs[4] = {fr_u64(base.lo.l[0]), fr_u64(base.lo.l[1]), fr_u64(base.hi.l[0]), fr_u64(base.hi.l[1])};
    const fr_t identifier = fr_add(rwc, fr_u64(3));
    const int single = hi0 && fr_eq_u64(exponent.lo, 2);
    word_t res, int_res; const word_t two = {fr_u64(2), fr_u64(0)};
    LK(exp_lookup(e, identifier, (uint64_t)single, limbs, exponent, &res), EV_EXP_FIRST_UNSAT);
    LK(exp_lookup(e, identifier, 1, limbs, two, &int_res), EV_EXP_LAST_UNSAT);
    { /* mul_add_words(base, base, Word(0), int_res): the overflow it returns is not constrained here */
      fr_t a64[4] = {limbs[0], limbs[1], limbs[2], limbs[3]};
#define M(x, y) fr_mul(a64[x], a64[y])
      const fr_t t0 = M(0, 0), t1 = fr_add(M(0, 1), M(1, 0));
      const fr_t t2 = fr_add(fr_add(M(0, 2), M(1, 1)), M(2, 0));
      const fr_t t3 = fr_add(fr_add(fr_add(M(0, 3), M(1, 2)), M(2, 1)), M(3, 0));
#undef M
      const fr_t two64 = {{0, 1, 0, 0}};
      const fr_t carry_lo = fr_mul(fr_sub(fr_add(t0, fr_mul(t1, two64)), int_res.lo), INV_2_128);
      const fr_t carry_hi = fr_mul(fr_sub(fr_add(fr_add(t2, fr_mul(t3, two64)), carry_lo), int_res.hi), INV_2_128);
      if (!fr_fits_bits(carry_lo, 72)) { orc_fail(e->res, EV_EXP_CARRY_LO, row); return; }
      if (!fr_fits_bits(carry_hi, 72)) { orc_fail(e->res, EV_EXP_CARRY_HI, row); return; }
    }
    CHECK(EV_EXP_RESULT, word_eq(res, pushed));
  }
  CHECK(EV_EXP_EXPONENT_BYTES, word_in_domain(exponent)); /* byte_size: to_le_bytes */
  const uint64_t v[4] = {exponent.lo.l[0], exponent.lo.l[1], exponent.hi.l[0], exponent.hi.l[1]};
  uint64_t size = 0;
  for (int k = 0; k < 32; k++) if ((v[k >> 3] >> (8 * (k & 7))) & 0xFF) size = (uint64_t)k + 1;
  same_context_x(e, i, row, opcode, fr_u64(3), one, one, 0, fr_u64(0), fr_u64(50 * size));
}

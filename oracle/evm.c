/*
 * oracle/evm.c — TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
 *
 * Restates verify_step and the hot gadgets of the reference's EVM circuit over the
 * column-major matrices of include/zkcheck.h:
 *   verify_steps / verify_step                  src/zkevm_specs/evm_circuit/main.py:14-63
 *   constrain_execution_state_transition        evm_circuit/instruction.py:189-204
 *   step_state_transition_in_same_context       evm_circuit/instruction.py:365-394, 206-264
 *   opcode_lookup / rw_lookup / stack_pop,push  evm_circuit/instruction.py:784-790, 792-824, 915-926
 *   add_sub                                     evm_circuit/execution/add_sub.py:5-24
 *   mul_div_mod (+ mul_add_words, compare_word) evm_circuit/execution/mul_div_mod.py:6-71,
 *                                               instruction.py:599-632, 447-463
 *   push                                        evm_circuit/execution/push.py:6-33
 *   pop                                         evm_circuit/execution/pop.py:4-14
 * Step = 13 cells in the order of StepState (evm_circuit/step.py:16-44), code_hash as lo,hi.
 * Pinned by tests/golden/evm.npz (verdicts of the reference's own verify_steps).
 */
#include "common.h"
#include "../include/zk_evm_spec.h"

enum { S_STATE, S_RWC, S_CALL_ID, S_IS_ROOT, S_IS_CREATE, S_HASH_LO, S_HASH_HI, S_PC, S_SP, S_GAS,
       S_MEM, S_REV, S_LOG, EVM_COLS };
/* bytecode table cells */
enum { B_HASH_LO, B_HASH_HI, B_TAG, B_INDEX, B_ISCODE, B_VALUE };
/* rw table cells */
enum { R_RWC, R_RW, R_TAG, R_ID, R_ADDR, R_FIELD, R_KEY_LO, R_KEY_HI, R_VAL_LO, R_VAL_HI, R_PREV_LO,
       R_PREV_HI, R_AUX_LO, R_AUX_HI };

static const int8_t ES_HALTS[ZK_ES_COUNT] = ZK_ES_HALTS_INIT;
static const int8_t ES_IMPL[ZK_ES_COUNT] = ZK_ES_IMPLEMENTED_INIT;
static const int16_t OPCODE_GAS[256] = ZK_OPCODE_GAS_INIT;

/* 2^-128 mod p */
static const fr_t INV_2_128 = {{0x18ee753c76f9dc6full, 0x54ad7e14a329e70full, 0x2b16366f4f7684dfull,
                                0x133100d71fdf3579ull}};

typedef struct {
  const uint64_t* steps; uint64_t n_steps;
  orc_index bytecode_ix, rw_ix, fixed_ix, copy_ix, keccak_ix;
  orc_index tx_ix, block_ix; /* tx table key (tx_id, tag, index); block table key (tag, block_number) */
  orc_index exp_ix;          /* exp table keyed on its first nine cells (exp_lookup, table.py:797-814) */
  const uint64_t* aux; uint64_t n_aux; /* step-aux side table: (step row, lo, hi), StepState.aux_data */
  orc_index rwc_ix;          /* rw table keyed on rw_counter alone (lookups with optional columns, evm_tx.h) */
  const uint8_t* rw_flags; /* bit0: value.is_word, bit1: value_prev.is_word */
  const uint8_t *tx_flags, *block_flags; /* bit0: value.is_word */
  const uint64_t* wd_tab; uint64_t n_wd; /* withdrawal table (id, validator_id, address, amount) */
  orc_result* res;
} evm_env;

typedef struct { fr_t lo, hi; } word_t;

static int bytecode_lookup(evm_env* e, fr_t hlo, fr_t hhi, uint64_t tag, fr_t index, uint64_t is_code,
                           fr_t* value) {
  fr_t key[5] = {hlo, hhi, fr_u64(tag), index, fr_u64(is_code)};
  uint32_t row; int n = orc_lookup(&e->bytecode_ix, key, &row);
  if (n == 1) *value = fr_load(ORC_CELL(e->bytecode_ix.cells, e->bytecode_ix.n_rows, B_VALUE, row));
  return n;
}
static int rw_lookup(evm_env* e, fr_t rwc, uint64_t rw, uint64_t tag, fr_t id, fr_t addr, word_t* value) {
  fr_t key[5] = {rwc, fr_u64(rw), fr_u64(tag), id, addr};
  uint32_t row; int n = orc_lookup(&e->rw_ix, key, &row);
  if (n == 1) {
    value->lo = fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, R_VAL_LO, row));
    value->hi = fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, R_VAL_HI, row));
  }
  return n;
}
/* A step stops at its FIRST failing constraint (the reference raises there), so at most one
 * id is recorded per step. */
#define CHECK(id, cond) do { if (!(cond)) { orc_fail(e->res, (id), row); return; } } while (0)
/* report a lookup outcome; returns 1 if exactly one row matched */
static int need1(evm_env* e, int n, int id_unsat, uint64_t row) {
  if (n == 0) orc_fail(e->res, id_unsat, row);
  if (n >= 2) orc_fail(e->res, id_unsat + 1, row);
  return n == 1;
}

/* ---- 256/512-bit integer helpers for the witness assignment of mul_div_mod.py:23-41 ---- */
typedef struct { uint64_t l[8]; } u512;
static void mul256(const uint64_t a[4], const uint64_t b[4], u512* out) {
  memset(out, 0, sizeof *out);
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      u128 x = (u128)a[i] * b[j] + out->l[i + j] + c; out->l[i + j] = (uint64_t)x; c = (uint64_t)(x >> 64);
    }
    out->l[i + 4] = c;
  }
}
static int cmp256(const uint64_t a[4], const uint64_t b[4]) {
  for (int i = 3; i >= 0; i--) { if (a[i] < b[i]) return -1; if (a[i] > b[i]) return 1; }
  return 0;
}
static void sub256(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
  uint64_t br = 0; for (int i = 0; i < 4; i++) out[i] = sbb(a[i], b[i], &br);
}
/* q = n / d (d != 0), schoolbook shift-subtract */
static void div256(const uint64_t n[4], const uint64_t d[4], uint64_t q[4]) {
  uint64_t r[4] = {0, 0, 0, 0}; memset(q, 0, 32);
  for (int bit = 255; bit >= 0; bit--) {
    uint64_t top = r[3] >> 63;
    for (int i = 3; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
    r[0] = (r[0] << 1) | ((n[bit / 64] >> (bit % 64)) & 1);
    if (top || cmp256(r, d) >= 0) { sub256(r, d, r); q[bit / 64] |= 1ull << (bit % 64); }
  }
}
static void word_to_u256(word_t w, uint64_t out[4]) { out[0] = w.lo.l[0]; out[1] = w.lo.l[1]; out[2] = w.hi.l[0]; out[3] = w.hi.l[1]; }
static word_t u256_to_word(const uint64_t v[4]) { word_t w = {fr_u128(v[0], v[1]), fr_u128(v[2], v[3])}; return w; }
static int word_in_domain(word_t w) { return fr_fits_bits(w.lo, 128) && fr_fits_bits(w.hi, 128); }

/* shared epilogue: instruction.py:365-394 */
/* general form: rw_counter delta is a field element, memory_word_size either stays or moves To a
 * value, and a dynamic gas cost is added to the opcode's constant cost */
static void same_context_x(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, fr_t d_rwc, fr_t d_pc, fr_t d_sp,
                           int mem_to, fr_t mem_value, fr_t dyn_gas);
static void same_context(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, uint64_t d_rwc, fr_t d_pc,
                         fr_t d_sp) {
  same_context_x(e, i, row, opcode, fr_u64(d_rwc), d_pc, d_sp, 0, fr_u64(0), fr_u64(0));
}
static void same_context_r(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, fr_t d_rwc, fr_t d_pc, fr_t d_sp,
                           int mem_to, fr_t mem_value, fr_t dyn_gas, uint64_t d_rev);
static void same_context_rl(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, fr_t d_rwc, fr_t d_pc, fr_t d_sp,
                            int mem_to, fr_t mem_value, fr_t dyn_gas, uint64_t d_rev, fr_t d_log);
static void same_context_x(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, fr_t d_rwc, fr_t d_pc, fr_t d_sp,
                           int mem_to, fr_t mem_value, fr_t dyn_gas) {
  same_context_r(e, i, row, opcode, d_rwc, d_pc, d_sp, mem_to, mem_value, dyn_gas, 0);
}
/* + reversible_write_counter = Transition.delta(d_rev) */
static void same_context_r(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, fr_t d_rwc, fr_t d_pc, fr_t d_sp,
                           int mem_to, fr_t mem_value, fr_t dyn_gas, uint64_t d_rev) {
  same_context_rl(e, i, row, opcode, d_rwc, d_pc, d_sp, mem_to, mem_value, dyn_gas, d_rev, fr_u64(0));
}
/* + log_id = Transition.delta(d_log) */
static void same_context_rl(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, fr_t d_rwc, fr_t d_pc, fr_t d_sp,
                            int mem_to, fr_t mem_value, fr_t dyn_gas, uint64_t d_rev, fr_t d_log) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
#define CUR(c) fr_load(ORC_CELL(S, n, c, i))
#define NXT(c) fr_load(ORC_CELL(S, n, c, j))
  fr_t key[4] = {fr_u64(ZK_FIXED_ResponsibleOpcode), CUR(S_STATE), opcode, fr_u64(0)};
  CHECK(EV_SC_RESP_OPCODE, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  int gas_cost = -1;
  if (fr_fits_bits(opcode, 8)) gas_cost = OPCODE_GAS[opcode.l[0]];
  if (gas_cost < 0) { orc_fail(e->res, EV_SC_OPCODE_VALUE, row); return; }
  fr_t gas_after = fr_sub(CUR(S_GAS), fr_add(fr_u64((uint64_t)gas_cost), dyn_gas));
  CHECK(EV_SC_GAS_RANGE, fr_fits_bits(gas_after, 64));
  CHECK(EV_SC_RWC, fr_eq(NXT(S_RWC), fr_add(CUR(S_RWC), d_rwc)));
  CHECK(EV_SC_PC, fr_eq(NXT(S_PC), fr_add(CUR(S_PC), d_pc)));
  CHECK(EV_SC_SP, fr_eq(NXT(S_SP), fr_add(CUR(S_SP), d_sp)));
  CHECK(EV_SC_GAS, fr_eq(NXT(S_GAS), gas_after));
  CHECK(EV_SC_MEM, fr_eq(NXT(S_MEM), mem_to ? mem_value : CUR(S_MEM)));
  CHECK(EV_SC_REV, fr_eq(NXT(S_REV), fr_add(CUR(S_REV), fr_u64(d_rev))));
  CHECK(EV_SC_LOG, fr_eq(NXT(S_LOG), fr_add(CUR(S_LOG), d_log)));
  CHECK(EV_SC_CALL_ID, fr_eq(NXT(S_CALL_ID), CUR(S_CALL_ID)));
  CHECK(EV_SC_IS_ROOT, fr_eq(NXT(S_IS_ROOT), CUR(S_IS_ROOT)));
  CHECK(EV_SC_IS_CREATE, fr_eq(NXT(S_IS_CREATE), CUR(S_IS_CREATE)));
  CHECK(EV_SC_CODE_HASH,
          fr_eq(NXT(S_HASH_LO), CUR(S_HASH_LO)) && fr_eq(NXT(S_HASH_HI), CUR(S_HASH_HI)));
}

/* add_words on two words, carry dropped (util/arithmetic.py:236-242) */
static word_t add_words2(word_t x, word_t y) {
  fr_t slo = fr_add(x.lo, y.lo);
  fr_t carry_lo = fr_u128(slo.l[2], slo.l[3]);
  fr_t shi = fr_add(fr_add(x.hi, y.hi), carry_lo);
  word_t r = {fr_u128(slo.l[0], slo.l[1]), fr_u128(shi.l[0], shi.l[1])};
  return r;
}
static int word_eq(word_t a, word_t b) { return fr_eq(a.lo, b.lo) && fr_eq(a.hi, b.hi); }

static void gadget_add(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const int is_sub = fr_eq_u64(opcode, 3);
  word_t a, b, c;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &a), EV_ADD_A_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(1)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(1)), &b),
             EV_ADD_B_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(1)), &c),
             EV_ADD_C_UNSAT, row)) return;
  CHECK(EV_ADD_SUM, word_eq(add_words2(is_sub ? c : a, b), is_sub ? a : c));
  same_context(e, i, row, opcode, 3, fr_u64(1), fr_u64(1));
}

/* Word((sel*lo, sel*hi)) with the constructor's < 2^128 assertion (arithmetic.py:110-114) */
static int word_select(word_t w, fr_t sel, word_t* out) {
  out->lo = fr_mul(sel, w.lo); out->hi = fr_mul(sel, w.hi);
  return word_in_domain(*out);
}

static void gadget_mul(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const fr_t one = fr_u64(1);
  /* mul_div_mod.py:14-16 */
  const fr_t is_mul = fr_mul(fr_mul(fr_sub(fr_u64(4), opcode), fr_sub(fr_u64(6), opcode)), fr_inv(fr_u64(8)));
  const fr_t is_div = fr_mul(fr_mul(fr_sub(opcode, fr_u64(2)), fr_sub(fr_u64(6), opcode)), fr_inv(fr_u64(4)));
  const fr_t is_mod = fr_mul(fr_mul(fr_sub(opcode, fr_u64(2)), fr_sub(opcode, fr_u64(4))), fr_inv(fr_u64(8)));
  word_t pop1, pop2, push;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &pop1), EV_MUL_POP1_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &pop2),
             EV_MUL_POP2_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &push),
             EV_MUL_PUSH_UNSAT, row)) return;
  const int in_domain = word_in_domain(pop1) && word_in_domain(pop2) && word_in_domain(push);
  /* witness assignment by branch, mul_div_mod.py:23-41 (Python int arithmetic) */
  word_t a, b, c, d;
  const word_t zero = {fr_u64(0), fr_u64(0)};
  if (fr_eq_u64(is_mul, 1)) { a = pop1; b = pop2; c = zero; d = push; }
  else if (!in_domain) { orc_fail(e->res, EV_MUL_WITNESS_DOMAIN, row); return; } /* needs > 512-bit ints */
  else if (fr_eq_u64(is_div, 1)) {
    d = pop1; b = pop2; a = push;
    uint64_t dv[4], bv[4], av[4], cv[4]; u512 prod;
    word_to_u256(d, dv); word_to_u256(b, bv); word_to_u256(a, av);
    mul256(bv, av, &prod);
    if (prod.l[4] | prod.l[5] | prod.l[6] | prod.l[7] || cmp256(prod.l, dv) > 0) {
      orc_fail(e->res, EV_MUL_WITNESS_NEG, row); return; /* Word(negative) */
    }
    sub256(dv, prod.l, cv); c = u256_to_word(cv);
  } else {
    d = pop1; b = pop2;
    uint64_t dv[4], bv[4], cv[4], qv[4], tv[4];
    word_to_u256(d, dv); word_to_u256(b, bv);
    if ((bv[0] | bv[1] | bv[2] | bv[3]) == 0) { c = d; a = zero; }
    else {
      c = push; word_to_u256(c, cv);
      if (cmp256(dv, cv) < 0) { orc_fail(e->res, EV_MUL_WITNESS_NEG, row); return; } /* floor div < 0 */
      sub256(dv, cv, tv); div256(tv, bv, qv); a = u256_to_word(qv);
    }
  }
  const fr_t divisor_is_zero = fr_u64(fr_is_zero(fr_add(b.lo, b.hi)));
  /* mul_add_words, instruction.py:599-632 */
  CHECK(EV_MUL_TO64, word_in_domain(a) && word_in_domain(b));
  fr_t a64[4] = {fr_u64(a.lo.l[0]), fr_u64(a.lo.l[1]), fr_u64(a.hi.l[0]), fr_u64(a.hi.l[1])};
  fr_t b64[4] = {fr_u64(b.lo.l[0]), fr_u64(b.lo.l[1]), fr_u64(b.hi.l[0]), fr_u64(b.hi.l[1])};
#define M(x, y) fr_mul(a64[x], b64[y])
  const fr_t t0 = M(0, 0);
  const fr_t t1 = fr_add(M(0, 1), M(1, 0));
  const fr_t t2 = fr_add(fr_add(M(0, 2), M(1, 1)), M(2, 0));
  const fr_t t3 = fr_add(fr_add(fr_add(M(0, 3), M(1, 2)), M(2, 1)), M(3, 0));
  const fr_t two64 = {{0, 1, 0, 0}};
  const fr_t carry_lo = fr_mul(fr_sub(fr_add(fr_add(t0, fr_mul(t1, two64)), c.lo), d.lo), INV_2_128);
  const fr_t carry_hi =
      fr_mul(fr_sub(fr_add(fr_add(fr_add(t2, fr_mul(t3, two64)), c.hi), carry_lo), d.hi), INV_2_128);
  fr_t overflow = carry_hi;
  overflow = fr_add(overflow, M(1, 3)); overflow = fr_add(overflow, M(2, 2));
  overflow = fr_add(overflow, M(3, 1)); overflow = fr_add(overflow, M(2, 3));
  overflow = fr_add(overflow, M(3, 2)); overflow = fr_add(overflow, M(3, 3));
#undef M
  if (!fr_fits_bits(carry_lo, 72)) { orc_fail(e->res, EV_MUL_CARRY_LO, row); return; }
  if (!fr_fits_bits(carry_hi, 72)) { orc_fail(e->res, EV_MUL_CARRY_HI, row); return; }
  /* the two constrain_equal of instruction.py:629-630 hold by construction of the carries */
  /* mul_div_mod.py:47-54 */
  if (!(fr_eq_u64(is_mul, 0) || fr_eq_u64(is_mul, 1))) { orc_fail(e->res, EV_MUL_SELECT, row); return; }
  /* pop1 == select_word(is_mul, a, d) and pop2 == b hold by the assignment above */
  word_t t_d, t_a, t_c, sum;
  const fr_t nz = fr_sub(one, divisor_is_zero);
  if (!word_select(d, is_mul, &t_d) || !word_select(a, fr_mul(is_div, nz), &t_a)) {
    orc_fail(e->res, EV_MUL_SELECT, row); return;
  }
  if (!word_select(c, fr_mul(is_mod, nz), &t_c)) { orc_fail(e->res, EV_MUL_SELECT, row); return; }
  sum.lo = fr_add(t_d.lo, t_a.lo); sum.hi = fr_add(t_d.hi, t_a.hi);
  if (!word_in_domain(sum)) { orc_fail(e->res, EV_MUL_SELECT, row); return; }
  sum.lo = fr_add(sum.lo, t_c.lo); sum.hi = fr_add(sum.hi, t_c.hi);
  if (!word_in_domain(sum)) { orc_fail(e->res, EV_MUL_SELECT, row); return; }
  CHECK(EV_MUL_PUSH_EQ, word_eq(push, sum));
  /* :57  is_mul * sum(c.to_le_bytes()) == 0 */
  uint64_t byte_sum = 0;
  for (int k = 0; k < 2; k++) for (int s = 0; s < 64; s += 8) {
    byte_sum += (c.lo.l[k] >> s) & 0xFF; byte_sum += (c.hi.l[k] >> s) & 0xFF;
  }
  CHECK(EV_MUL_C_ZERO, fr_is_zero(fr_mul(is_mul, fr_u64(byte_sum))));
  /* :60-61 compare_word(c, b) */
  const int hi_lt = fr_cmp(c.hi, b.hi) < 0, hi_eq = fr_eq(c.hi, b.hi), lo_lt = fr_cmp(c.lo, b.lo) < 0;
  const fr_t lt = fr_u64((uint64_t)(hi_lt + hi_eq * lo_lt));
  CHECK(EV_MUL_REM_LT,
          fr_is_zero(fr_mul(fr_mul(fr_sub(one, is_mul), nz), fr_sub(one, lt))));
  CHECK(EV_MUL_OVERFLOW, fr_is_zero(fr_mul(fr_sub(one, is_mul), overflow)));
  same_context(e, i, row, opcode, 3, one, one);
}

static void gadget_push(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), pc = CUR(S_PC);
  const fr_t hlo = CUR(S_HASH_LO), hhi = CUR(S_HASH_HI);
  const fr_t num_pushed = fr_sub(opcode, fr_u64(0x5f));
  fr_t code_length;
  if (!need1(e, bytecode_lookup(e, hlo, hhi, 1, fr_u64(0), 0, &code_length), EV_PUSH_LEN_UNSAT, row)) return;
  const fr_t left = fr_sub(fr_sub(code_length, pc), fr_u64(1));
  if (!fr_fits_bits(left, 64) || !fr_fits_bits(num_pushed, 64)) { orc_fail(e->res, EV_PUSH_CMP_RANGE, row); return; }
  const int oob = left.l[0] < num_pushed.l[0];
  const fr_t num_padding = oob ? fr_sub(num_pushed, left) : fr_u64(0);
  word_t value;
  if (!need1(e, rw_lookup(e, rwc, 1, ZK_TARGET_Stack, call_id, fr_sub(sp, fr_u64(1)), &value),
             EV_PUSH_RW_UNSAT, row)) return;
  if (!word_in_domain(value)) { orc_fail(e->res, EV_PUSH_VALUE_BYTES, row); return; }
  for (int idx = 0; idx < 32; idx++) {
    const uint64_t byte = idx < 16 ? (value.lo.l[idx / 8] >> (8 * (idx % 8))) & 0xFF
                                   : (value.hi.l[(idx - 16) / 8] >> (8 * (idx % 8))) & 0xFF;
    /* continuous_selectors: FQ(i < value.n) on the full integer (instruction.py:413-414) */
    const int is_pushed = fr_cmp(fr_u64((uint64_t)idx), num_pushed) < 0;
    const int is_padding = fr_cmp(fr_u64((uint64_t)idx), num_padding) < 0;
    const int base = EV_PUSH_B0_UNSAT + 4 * idx;
    if (is_pushed && !is_padding) {
      fr_t got;
      const fr_t index = fr_sub(fr_add(pc, num_pushed), fr_u64((uint64_t)idx));
      if (!need1(e, bytecode_lookup(e, hlo, hhi, 2, index, 0, &got), base, row)) return;
      if (!fr_eq_u64(got, byte)) { orc_fail(e->res, base + 2, row); return; }
    } else if (byte != 0) { orc_fail(e->res, base + 3, row); return; }
  }
  same_context(e, i, row, opcode, 1, fr_add(fr_u64(1), num_pushed), fr_neg(fr_u64(1)));
}

static void gadget_pop(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  word_t y;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), CUR(S_SP), &y), EV_POP_RW_UNSAT, row))
    return;
  same_context(e, i, row, opcode, 1, fr_u64(1), fr_u64(1));
}


/* ---- helpers of the copy gadgets ------------------------------------------------------------ */
/* word_to_fq(word, 5): 0 ok, 1 -> to_le_bytes OverflowError, 2 -> ConstraintUnsatFailure
 * (instruction.py:480-484) */
static int word_to_fq5(word_t w, fr_t* out) {
  if (!word_in_domain(w)) return 1;
  if ((w.lo.l[0] >> 40) || w.lo.l[1] || w.hi.l[0] || w.hi.l[1]) return 2;
  *out = fr_u64(w.lo.l[0]);
  return 0;
}
/* memory_gas_cost(size) for size < 2^32: size^2 // 512 + 3 size (instruction.py:1129-1136) */
static uint64_t memory_gas_cost(uint64_t size) { return size * size / 512 + 3 * size; }
/* memory_expansion_dynamic_length + memory_copier_gas_cost (instruction.py:1157-1192); returns the
 * failing id offset 0..3 (+1) or 0; ids are base_id + {MEMSIZE, MAX, WORDSIZE, GASCOST} */
static int copier_gas(evm_env* e, uint64_t i, uint64_t offset, uint64_t length, uint64_t per_word,
                      fr_t* next_mem, fr_t* gas) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const uint64_t cd_size = (offset + length + 31) / 32; /* offset, length < 2^40 */
  if (cd_size >> 32) return 1;
  const fr_t cur = CUR(S_MEM);
  if (!fr_fits_bits(cur, 32)) return 2;
  const uint64_t nxt = cur.l[0] < cd_size ? cd_size : cur.l[0];
  const uint64_t expansion = memory_gas_cost(nxt) - memory_gas_cost(cur.l[0]);
  const uint64_t words = (length + 31) / 32;
  if (words >> 32) return 3;
  const u128 g = (u128)words * per_word + expansion;
  if (g >> 64) return 4;
  *next_mem = fr_u64(nxt); *gas = fr_u64((uint64_t)g);
  return 0;
}
/* copy-table cells: is_first, src_id lo/hi, src_tag, dst_id lo/hi, dst_tag, src_addr, src_addr_end,
 * dst_addr, length, rlc_acc, rw_counter, rwc_inc; the ids are compared as Words (table.py:776-778) */
static int copy_lookup_w(evm_env* e, fr_t src_lo, fr_t src_hi, uint64_t src_tag, fr_t dst_id, uint64_t dst_tag, fr_t src_addr,
                         fr_t src_end, fr_t dst_addr, fr_t length, fr_t rwc, fr_t* rwc_inc, fr_t* rlc_acc) {
  fr_t key[11] = {src_lo, src_hi, fr_u64(src_tag), dst_id, fr_u64(0), fr_u64(dst_tag), src_addr, src_end,
                  dst_addr, length, rwc};
  uint32_t row; const int n = orc_lookup(&e->copy_ix, key, &row);
  if (n == 1) {
    *rlc_acc = fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 11, row));
    *rwc_inc = fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 13, row));
  }
  return n;
}
static int copy_lookup(evm_env* e, fr_t src_id, uint64_t src_tag, fr_t dst_id, uint64_t dst_tag, fr_t src_addr,
                       fr_t src_end, fr_t dst_addr, fr_t length, fr_t rwc, fr_t* rwc_inc, fr_t* rlc_acc) {
  /* ids are values here (hi = 0) */
  fr_t key[11] = {src_id, fr_u64(0), fr_u64(src_tag), dst_id, fr_u64(0), fr_u64(dst_tag), src_addr, src_end,
                  dst_addr, length, rwc};
  uint32_t row; const int n = orc_lookup(&e->copy_ix, key, &row);
  if (n == 1) {
    *rlc_acc = fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 11, row));
    *rwc_inc = fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 13, row));
  }
  return n;
}

static void gadget_sha3(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t off_w, size_w, val_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &off_w), EV_SHA_OFF_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &size_w), EV_SHA_SIZE_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &val_w), EV_SHA_VAL_UNSAT, row)) return;
  fr_t length, offset = fr_u64(0);
  int rc = word_to_fq5(size_w, &length);
  if (rc) { orc_fail(e->res, rc == 1 ? EV_SHA_LEN_BYTES : EV_SHA_LEN_RANGE, row); return; }
  if (!fr_is_zero(length)) {
    rc = word_to_fq5(off_w, &offset);
    if (rc) { orc_fail(e->res, rc == 1 ? EV_SHA_OFF_BYTES : EV_SHA_OFF_RANGE, row); return; }
  }
  fr_t rwc_inc = fr_u64(0), rlc_acc = fr_u64(0);
  if (!fr_is_zero(length)) {
    if (!need1(e, copy_lookup(e, call_id, ZK_COPY_Memory, call_id, ZK_COPY_RlcAcc, offset, fr_add(offset, length),
                              fr_u64(0), length, fr_add(rwc, fr_u64(3)), &rwc_inc, &rlc_acc), EV_SHA_COPY_UNSAT, row)) return;
  }
  { /* keccak_lookup(length, rlc_acc): key (state_tag=2, input_rlc, input_len) */
    fr_t key[3] = {fr_u64(2), rlc_acc, length};
    uint32_t hit; const int nk = orc_lookup(&e->keccak_ix, key, &hit);
    if (!need1(e, nk, EV_SHA_KECCAK_UNSAT, row)) return;
    CHECK(EV_SHA_HASH_EQ, fr_eq(fr_load(ORC_CELL(e->keccak_ix.cells, e->keccak_ix.n_rows, 3, hit)), val_w.lo) &&
                              fr_eq(fr_load(ORC_CELL(e->keccak_ix.cells, e->keccak_ix.n_rows, 4, hit)), val_w.hi));
  }
  fr_t next_mem, gas;
  rc = copier_gas(e, i, offset.l[0], length.l[0], ZK_GAS_COST_COPY_SHA3, &next_mem, &gas);
  if (rc) { orc_fail(e->res, EV_SHA_MEMSIZE_RANGE + rc - 1, row); return; }
  same_context_x(e, i, row, opcode, fr_add(fr_u64(3), rwc_inc), one, one, 1, next_mem, gas);
}

/* call_context_lookup(field_tag) at rw_counter + k: rw row (Read, CallContext, call_id, address = tag) */
static int call_context(evm_env* e, fr_t rwc, fr_t call_id, uint64_t field_tag, fr_t* value, int* is_word) {
  fr_t key[5] = {rwc, fr_u64(0), fr_u64(ZK_TARGET_CallContext), call_id, fr_u64(field_tag)};
  uint32_t r; const int n = orc_lookup(&e->rw_ix, key, &r);
  if (n == 1) {
    *value = fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, R_VAL_LO, r));
    *is_word = e->rw_flags ? (e->rw_flags[r] & 1) : 0;
  }
  return n;
}

static void gadget_calldatacopy(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t moff_w, doff_w, len_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &moff_w), EV_CDC_MOFF_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &doff_w), EV_CDC_DOFF_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(2)), &len_w), EV_CDC_LEN_UNSAT, row)) return;
  fr_t length, moff = fr_u64(0), doff;
  int rc = word_to_fq5(len_w, &length);
  if (rc) { orc_fail(e->res, rc == 1 ? EV_CDC_LEN_BYTES : EV_CDC_LEN_RANGE, row); return; }
  if (!fr_is_zero(length)) {
    rc = word_to_fq5(moff_w, &moff);
    if (rc) { orc_fail(e->res, rc == 1 ? EV_CDC_MOFF_BYTES : EV_CDC_MOFF_RANGE, row); return; }
  }
  rc = word_to_fq5(doff_w, &doff);
  if (rc) { orc_fail(e->res, rc == 1 ? EV_CDC_DOFF_BYTES : EV_CDC_DOFF_RANGE, row); return; }
  const fr_t is_root = CUR(S_IS_ROOT);
  const int root = !fr_is_zero(is_root); /* Python truthiness of the is_root attribute */
  fr_t src_id, cd_len, cd_off = fr_u64(0);
  int w;
  uint64_t k = 3;
  if (!need1(e, call_context(e, fr_add(rwc, fr_u64(k)), call_id, root ? ZK_CC_TxId : ZK_CC_CallerId, &src_id, &w), EV_CDC_CC1_UNSAT, row)) return;
  CHECK(EV_CDC_CC1_TYPE, !w); k++;
  if (!need1(e, call_context(e, fr_add(rwc, fr_u64(k)), call_id, ZK_CC_CallDataLength, &cd_len, &w), EV_CDC_CC2_UNSAT, row)) return;
  CHECK(EV_CDC_CC2_TYPE, !w); k++;
  if (!root) {
    if (!need1(e, call_context(e, fr_add(rwc, fr_u64(k)), call_id, ZK_CC_CallDataOffset, &cd_off, &w), EV_CDC_CC3_UNSAT, row)) return;
    CHECK(EV_CDC_CC3_TYPE, !w); k++;
  }
  fr_t next_mem, gas;
  rc = copier_gas(e, i, moff.l[0], length.l[0], ZK_GAS_COST_COPY, &next_mem, &gas);
  if (rc) { orc_fail(e->res, EV_CDC_MEMSIZE_RANGE + rc - 1, row); return; }
  CHECK(EV_CDC_SELECT_BOOL, fr_eq_u64(is_root, 0) || fr_eq_u64(is_root, 1));
  fr_t rwc_inc = fr_u64(0), unused;
  if (!fr_is_zero(length)) {
    if (!need1(e, copy_lookup(e, src_id, root ? ZK_COPY_TxCalldata : ZK_COPY_Memory, call_id, ZK_COPY_Memory,
                              fr_add(cd_off, doff), fr_add(cd_off, cd_len), moff, length, fr_add(rwc, fr_u64(k)),
                              &rwc_inc, &unused), EV_CDC_COPY_UNSAT, row)) return;
  }
  same_context_x(e, i, row, opcode, fr_add(fr_u64(k), rwc_inc), one, fr_u64(3), 1, next_mem, gas);
}


/* call_context_lookup_word at rw_counter, rw, call_id: value word + is_word flag */
static int call_context_w(evm_env* e, fr_t rwc, uint64_t rw, fr_t call_id, uint64_t field_tag, word_t* value, int* is_word) {
  fr_t key[5] = {rwc, fr_u64(rw), fr_u64(ZK_TARGET_CallContext), call_id, fr_u64(field_tag)};
  uint32_t r; const int n = orc_lookup(&e->rw_ix, key, &r);
  if (n == 1) {
    value->lo = fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, R_VAL_LO, r));
    value->hi = fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, R_VAL_HI, r));
    *is_word = e->rw_flags ? (e->rw_flags[r] & 1) : 0;
  }
  return n;
}
/* step_state_transition_to_restored_context (instruction.py:293-363) with caller_id = None;
 * rw_off = rw lookups the gadget already did, add_rev = curr state halts in success */
/* lookups at rw_counter + look_off + k (k = 0..11), next rw_counter = rw_counter + delta12 + 12 (the two differ in
 * return_revert.py's CREATE branch, whose rwc_delta forgets two lookups) */
static void restore_context_f(evm_env* e, uint64_t i, uint64_t row, fr_t look_off, fr_t delta12, fr_t ret_off, fr_t ret_len,
                              fr_t gas_left, int add_rev);
static void restore_context_x(evm_env* e, uint64_t i, uint64_t row, uint64_t rw_off, fr_t ret_off, fr_t ret_len,
                              fr_t gas_left, int add_rev, fr_t extra_delta) {
  restore_context_f(e, i, row, fr_u64(rw_off), fr_add(fr_u64(rw_off), extra_delta), ret_off, ret_len, gas_left, add_rev);
}
static void restore_context_f(evm_env* e, uint64_t i, uint64_t row, fr_t look_off, fr_t delta12, fr_t ret_off, fr_t ret_len,
                              fr_t gas_left, int add_rev) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC);
  const fr_t lrwc = fr_add(rwc, look_off);
  static const uint64_t READ_TAGS[8] = {ZK_CC_IsRoot, ZK_CC_IsCreate, ZK_CC_CodeHash, ZK_CC_ProgramCounter,
                                        ZK_CC_StackPointer, ZK_CC_GasLeft, ZK_CC_MemorySize, ZK_CC_ReversibleWriteCounter};
  static const uint64_t WRITE_TAGS[3] = {ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataOffset, ZK_CC_LastCalleeReturnDataLength};
  word_t v; int w;
  if (!need1(e, call_context_w(e, lrwc, 0, CUR(S_CALL_ID), ZK_CC_CallerId, &v, &w), EV_RST0_UNSAT, row)) return;
  CHECK(EV_RST0_CHECK, !w);
  const fr_t caller_id = v.lo;
  word_t vals[8]; int words[8];
  for (int k = 0; k < 8; k++) {
    if (!need1(e, call_context_w(e, fr_add(lrwc, fr_u64(1 + k)), 0, caller_id, READ_TAGS[k], &vals[k], &words[k]),
               EV_RST0_UNSAT + 3 * (1 + k), row)) return;
  }
  const fr_t expected[3] = {CUR(S_CALL_ID), ret_off, ret_len};
  for (int k = 0; k < 3; k++) {
    if (!need1(e, call_context_w(e, fr_add(lrwc, fr_u64(9 + k)), 1, caller_id, WRITE_TAGS[k], &v, &w),
               EV_RST0_UNSAT + 3 * (9 + k), row)) return;
    CHECK(EV_RST0_UNSAT + 3 * (9 + k) + 2, !w && fr_eq(v.lo, expected[k]));
  }
  CHECK(EV_RST_VALUE_TYPE, !words[0] && !words[1] && !words[3] && !words[4] && !words[5] && !words[6] && !words[7]);
  const fr_t rev = add_rev ? CUR(S_REV) : fr_u64(0);
  CHECK(EV_RST_RWC, fr_eq(NXT(S_RWC), fr_add(fr_add(rwc, delta12), fr_u64(12))));
  CHECK(EV_RST_CALL_ID, fr_eq(NXT(S_CALL_ID), caller_id));
  CHECK(EV_RST_IS_ROOT, fr_eq(NXT(S_IS_ROOT), vals[0].lo));
  CHECK(EV_RST_IS_CREATE, fr_eq(NXT(S_IS_CREATE), vals[1].lo));
  CHECK(EV_RST_CODE_HASH, fr_eq(NXT(S_HASH_LO), vals[2].lo) && fr_eq(NXT(S_HASH_HI), vals[2].hi));
  CHECK(EV_RST_PC, fr_eq(NXT(S_PC), vals[3].lo));
  CHECK(EV_RST_SP, fr_eq(NXT(S_SP), vals[4].lo));
  CHECK(EV_RST_GAS, fr_eq(NXT(S_GAS), fr_add(vals[5].lo, gas_left)));
  CHECK(EV_RST_MEM, fr_eq(NXT(S_MEM), vals[6].lo));
  CHECK(EV_RST_REV, fr_eq(NXT(S_REV), fr_add(vals[7].lo, rev)));
}
static void restore_context(evm_env* e, uint64_t i, uint64_t row, uint64_t rw_off, fr_t ret_off, fr_t ret_len,
                            fr_t gas_left, int add_rev) {
  restore_context_x(e, i, row, rw_off, ret_off, ret_len, gas_left, add_rev, fr_u64(0));
}

static void gadget_stop(evm_env* e, uint64_t i, uint64_t row) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  fr_t code_length;
  if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 1, fr_u64(0), 0, &code_length), EV_STOP_LEN_UNSAT, row)) return;
  const fr_t pc = CUR(S_PC);
  CHECK(EV_STOP_CMP_RANGE, fr_fits_bits(code_length, 64) && fr_fits_bits(pc, 64));
  if (code_length.l[0] > pc.l[0]) { /* lt + eq == 0 */
    fr_t opcode;
    if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 2, pc, 1, &opcode), EV_STOP_OP_UNSAT, row)) return;
    fr_t key[4] = {fr_u64(ZK_FIXED_ResponsibleOpcode), CUR(S_STATE), opcode, fr_u64(0)};
    CHECK(EV_STOP_RESP_OPCODE, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  }
  word_t v; int w;
  if (!need1(e, call_context_w(e, CUR(S_RWC), 0, CUR(S_CALL_ID), ZK_CC_IsSuccess, &v, &w), EV_STOP_CC_UNSAT, row)) return;
  CHECK(EV_STOP_CC_TYPE, !w);
  CHECK(EV_STOP_IS_SUCCESS, fr_eq_u64(v.lo, 1));
  const fr_t is_root = CUR(S_IS_ROOT);
  CHECK(EV_STOP_ROOT_ENDTX, fr_eq_u64(is_root, fr_eq_u64(NXT(S_STATE), ZK_ES_EndTx) ? 1 : 0));
  if (!fr_is_zero(is_root)) {
    CHECK(EV_STOP_RWC, fr_eq(NXT(S_RWC), fr_add(CUR(S_RWC), fr_u64(1))));
    CHECK(EV_STOP_CALL_ID, fr_eq(NXT(S_CALL_ID), CUR(S_CALL_ID)));
  } else {
    restore_context(e, i, row, 1, fr_u64(0), fr_u64(0), CUR(S_GAS), 1);
  }
}

/* memory (execution/memory.py:7-44): MLOAD / MSTORE / MSTORE8.  NB the byte values are NOT
 * constrained: `instruction.is_equal(memory_lookup(...), byte)` only computes a flag (memory.py:26,
 * 31-36), so each of the 1 / 32 memory lookups must merely exist, be unique and hold a value. */
static void gadget_memory(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t addr_w, val_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &addr_w), EV_MEM_ADDR_UNSAT, row)) return;
  CHECK(EV_MEM_ADDR_BYTES, word_in_domain(addr_w));
  CHECK(EV_MEM_ADDR_RANGE, (addr_w.hi.l[0] >> 32) == 0 && addr_w.hi.l[1] == 0); /* bytes 20..31 */
  fr_t address = addr_w.lo; address.l[2] = addr_w.hi.l[0]; /* lo + 2^128 * hi[0:4] < 2^160 */
  const int is_mload = fr_eq_u64(opcode, 0x51), is_mstore8 = fr_eq_u64(opcode, 0x53);
  const int is_store = !is_mload;
  if (is_mload) {
    if (!need1(e, rw_lookup(e, fr_add(rwc, one), 1, ZK_TARGET_Stack, call_id, sp, &val_w), EV_MEM_VAL_UNSAT, row)) return;
  } else {
    if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &val_w), EV_MEM_VAL_UNSAT, row)) return;
  }
  CHECK(EV_MEM_VAL_BYTES, word_in_domain(val_w));
  /* memory_expansion(offset = curr.memory_word_size, length = address + 1 + 31 * (1 - is_mstore8)):
   * memory_size = (length + offset + 31) // 32 must fit 4 bytes (instruction.py:1138-1155) */
  const fr_t cur_mem = CUR(S_MEM);
  const fr_t num = fr_add(fr_add(fr_add(address, fr_u64(is_mstore8 ? 1 : 32)), cur_mem), fr_u64(31));
  CHECK(EV_MEM_MEMSIZE_RANGE, fr_fits_bits(num, 37));
  const uint64_t mem_size = num.l[0] >> 5;
  CHECK(EV_MEM_MAX_RANGE, fr_fits_bits(cur_mem, 32));
  const uint64_t nxt = cur_mem.l[0] < mem_size ? mem_size : cur_mem.l[0];
  const fr_t gas = fr_u64(memory_gas_cost(nxt) - memory_gas_cost(cur_mem.l[0]));
  const int n_bytes = is_mstore8 ? 1 : 32;
  for (int k = 0; k < n_bytes; k++) {
    fr_t key[5] = {fr_add(rwc, fr_u64(2 + k)), fr_u64(is_store ? 1 : 0), fr_u64(ZK_TARGET_Memory), call_id,
                   fr_add(address, fr_u64(k))};
    uint32_t r; const int m = orc_lookup(&e->rw_ix, key, &r);
    if (!need1(e, m, EV_MEM_BYTE_UNSAT, row)) return;
    CHECK(EV_MEM_BYTE_TYPE, !(e->rw_flags && (e->rw_flags[r] & 1)));
  }
  same_context_x(e, i, row, opcode, fr_u64(is_mstore8 ? 3 : 34), one, fr_u64(is_store ? 2 : 0), 1, fr_u64(nxt), gas);
}

/* ---- simple same-context gadgets: msize.py, gas.py, iszero.py, comparator.py, jump.py, jumpi.py ---- */
static int word_is(word_t w, fr_t lo) { return fr_eq(w.lo, lo) && fr_is_zero(w.hi); }
static void gadget_msize(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t v = fr_mul(CUR(S_MEM), fr_u64(32)); /* memory_word_size * N_BYTES_WORD over the field */
  CHECK(EV_MSZ_WORD, fr_fits_bits(v, 128));
  word_t w;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_sub(CUR(S_SP), fr_u64(1)), &w), EV_MSZ_PUSH_UNSAT, row)) return;
  CHECK(EV_MSZ_EQ, word_is(w, v));
  same_context(e, i, row, opcode, 1, fr_u64(1), fr_neg(fr_u64(1)));
}
static void gadget_gas(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_GAS_OPCODE, fr_eq_u64(opcode, 0x5a));
  const fr_t v = fr_sub(CUR(S_GAS), fr_u64(2)); /* Opcode.GAS.constant_gas_cost() == 2 */
  CHECK(EV_GAS_WORD, fr_fits_bits(v, 128));
  word_t w;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_sub(CUR(S_SP), fr_u64(1)), &w), EV_GAS_PUSH_UNSAT, row)) return;
  CHECK(EV_GAS_EQ, word_is(w, v));
  same_context(e, i, row, opcode, 1, fr_u64(1), fr_neg(fr_u64(1)));
}
static void gadget_iszero(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  word_t v, w;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), CUR(S_SP), &v), EV_ISZ_POP_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(CUR(S_RWC), fr_u64(1)), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), CUR(S_SP), &w), EV_ISZ_PUSH_UNSAT, row)) return;
  CHECK(EV_ISZ_EQ, word_is(w, fr_u64(fr_is_zero(fr_add(v.lo, v.hi)) ? 1 : 0))); /* is_zero_word: field sum of the halves */
  same_context(e, i, row, opcode, 2, fr_u64(1), fr_u64(0));
}
static void gadget_cmp(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  const int is_eq = fr_eq_u64(opcode, 0x14), is_gt = fr_eq_u64(opcode, 0x11);
  word_t a, b, c;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &a), EV_CMP_A_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &b), EV_CMP_B_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &c), EV_CMP_C_UNSAT, row)) return;
  const word_t aa = is_gt ? b : a, bb = is_gt ? a : b;
  CHECK(EV_CMP_RANGE_LO, fr_fits_bits(aa.lo, 128) && fr_fits_bits(bb.lo, 128));
  CHECK(EV_CMP_RANGE_HI, fr_fits_bits(aa.hi, 128) && fr_fits_bits(bb.hi, 128));
  const int lt_lo = fr_cmp(aa.lo, bb.lo) < 0, eq_lo = fr_eq(aa.lo, bb.lo);
  const int lt_hi = fr_cmp(aa.hi, bb.hi) < 0, eq_hi = fr_eq(aa.hi, bb.hi);
  const int lt = lt_hi ? 1 : (eq_hi && lt_lo), eq = eq_lo && eq_hi;
  CHECK(EV_CMP_EQ, word_is(c, fr_u64(is_eq ? eq : lt)));
  same_context(e, i, row, opcode, 3, one, one);
}
static void gadget_jump(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_JMP_OPCODE, fr_eq_u64(opcode, 0x56));
  word_t dest;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), CUR(S_SP), &dest), EV_JMP_DEST_UNSAT, row)) return;
  CHECK(EV_JMP_DEST_HI, fr_is_zero(dest.hi));
  fr_t at;
  if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 2, dest.lo, 1, &at), EV_JMP_AT_UNSAT, row)) return;
  CHECK(EV_JMP_NOT_JUMPDEST, fr_eq_u64(at, 0x5b));
  /* program_counter = Transition.to(dest): next.pc == dest, i.e. delta dest - pc over the field */
  same_context(e, i, row, opcode, 1, fr_sub(dest.lo, CUR(S_PC)), fr_u64(1));
}
static void gadget_jumpi(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_JMPI_OPCODE, fr_eq_u64(opcode, 0x57));
  word_t dest, cond;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), CUR(S_SP), &dest), EV_JMPI_DEST_UNSAT, row)) return;
  CHECK(EV_JMPI_DEST_HI, fr_is_zero(dest.hi));
  if (!need1(e, rw_lookup(e, fr_add(CUR(S_RWC), fr_u64(1)), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_add(CUR(S_SP), fr_u64(1)), &cond), EV_JMPI_COND_UNSAT, row)) return;
  /* jumpi.py:20 `if instruction.is_zero_word(cond):` tests the truthiness of an FQ OBJECT (py_ecc's FQ
   * defines neither __bool__ nor __len__), which is always true: the reference takes the
   * fall-through branch (pc + 1) whatever cond is and never looks at the destination. */
  same_context(e, i, row, opcode, 2, fr_u64(1), fr_u64(2));
}

/* caller.py / callvalue.py / calldatasize.py / address.py / returndatasize.py: constrain the opcode, read
 * one call-context field (as a Word, or as a value wrapped by Word.from_lo), push it */
static void gadget_cc_push(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, uint64_t op, uint64_t field, int as_word) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_CCP_OPCODE, fr_eq_u64(opcode, op));
  word_t v, w; int is_word;
  if (!need1(e, call_context_w(e, CUR(S_RWC), 0, CUR(S_CALL_ID), field, &v, &is_word), EV_CCP_CC_UNSAT, row)) return;
  if (!as_word) {
    CHECK(EV_CCP_CC_TYPE, !is_word);
    CHECK(EV_CCP_WORD, fr_fits_bits(v.lo, 128));
    v.hi = fr_u64(0);
  }
  if (!need1(e, rw_lookup(e, fr_add(CUR(S_RWC), fr_u64(1)), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_sub(CUR(S_SP), fr_u64(1)), &w), EV_CCP_PUSH_UNSAT, row)) return;
  CHECK(EV_CCP_EQ, fr_eq(w.lo, v.lo) && fr_eq(w.hi, v.hi));
  same_context(e, i, row, opcode, 2, fr_u64(1), fr_neg(fr_u64(1)));
}
static void gadget_codesize(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_CSZ_OPCODE, fr_eq_u64(opcode, 0x38));
  fr_t len; word_t w;
  if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 1, fr_u64(0), 0, &len), EV_CSZ_LEN_UNSAT, row)) return;
  CHECK(EV_CSZ_WORD, fr_fits_bits(len, 128));
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_sub(CUR(S_SP), fr_u64(1)), &w), EV_CSZ_PUSH_UNSAT, row)) return;
  CHECK(EV_CSZ_EQ, word_is(w, len));
  same_context(e, i, row, opcode, 1, fr_u64(1), fr_neg(fr_u64(1)));
}

/* bitwise.py / not_.py / byte.py */
static unsigned word_byte(word_t w, int k) { /* k-th little-endian byte of a word in the 128-bit-halves domain */
  const fr_t c = k < 16 ? w.lo : w.hi; k &= 15;
  return (unsigned)((c.l[k >> 3] >> (8 * (k & 7))) & 0xFF);
}
static void gadget_bitwise(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t a, b, c;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &a), EV_BW_A_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &b), EV_BW_B_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &c), EV_BW_C_UNSAT, row)) return;
  CHECK(EV_BW_BYTES, word_in_domain(a) && word_in_domain(b) && word_in_domain(c));
  /* tag = BitwiseAnd + (opcode.n - AND) as a Python int; FixedTableTag(tag) must exist (1..16) */
  CHECK(EV_BW_TAG, fr_fits_bits(opcode, 8) && opcode.l[0] + 10 >= 0x16 + 1 && opcode.l[0] + 10 <= 0x16 + 16);
  const uint64_t tag = opcode.l[0] + ZK_FIXED_BitwiseAnd - 0x16;
  for (int k = 0; k < 32; k++) {
    fr_t key[4] = {fr_u64(tag), fr_u64(word_byte(a, k)), fr_u64(word_byte(b, k)), fr_u64(word_byte(c, k))};
    if (!need1(e, orc_lookup(&e->fixed_ix, key, 0), EV_BW_FIXED_UNSAT, row)) return;
  }
  same_context(e, i, row, opcode, 3, one, one);
}
static void gadget_not(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  word_t a, b;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), CUR(S_SP), &a), EV_NOT_A_UNSAT, row)) return;
  CHECK(EV_NOT_A_BYTES, word_in_domain(a));
  if (!need1(e, rw_lookup(e, fr_add(CUR(S_RWC), fr_u64(1)), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), CUR(S_SP), &b), EV_NOT_B_UNSAT, row)) return;
  CHECK(EV_NOT_B_BYTES, word_in_domain(b));
  for (int k = 0; k < 32; k++) {
    fr_t key[4] = {fr_u64(ZK_FIXED_BitwiseXor), fr_u64(word_byte(a, k)), fr_u64(word_byte(b, k)), fr_u64(255)};
    if (!need1(e, orc_lookup(&e->fixed_ix, key, 0), EV_NOT_FIXED_UNSAT, row)) return;
  }
  same_context(e, i, row, opcode, 2, fr_u64(1), fr_u64(0));
}
static void gadget_byte(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t a, b, c;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &a), EV_BYTE_A_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &b), EV_BYTE_B_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &c), EV_BYTE_C_UNSAT, row)) return;
  CHECK(EV_BYTE_BYTES, word_in_domain(a) && word_in_domain(b));
  unsigned msb = 0;
  for (int k = 1; k < 32; k++) msb += word_byte(a, k);
  const unsigned idx0 = word_byte(a, 0);
  const unsigned sel = (msb == 0 && idx0 < 32) ? word_byte(b, 31 - (int)idx0) : 0;
  CHECK(EV_BYTE_EQ, word_is(c, fr_u64(sel)));
  same_context(e, i, row, opcode, 3, one, one);
}

/* slt_sgt.py */
static void gadget_scmp(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  const int is_sgt = fr_eq_u64(opcode, 0x13);
  word_t a, b, c;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &a), EV_SCMP_A_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &b), EV_SCMP_B_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &c), EV_SCMP_C_UNSAT, row)) return;
  const word_t aa = is_sgt ? b : a, bb = is_sgt ? a : b;
  CHECK(EV_SCMP_BYTES, word_in_domain(aa) && word_in_domain(bb) && word_in_domain(c));
  CHECK(EV_SCMP_C_MSB, word_byte(c, 31) == 0);
  const int lt_lo = fr_cmp(aa.lo, bb.lo) < 0, lt_hi = fr_cmp(aa.hi, bb.hi) < 0, eq_hi = fr_eq(aa.hi, bb.hi);
  const int a_lt_b = lt_hi ? 1 : (eq_hi && lt_lo);
  const int a_neg = word_byte(aa, 31) >= 128, b_neg = word_byte(bb, 31) >= 128;
  const int expect = (a_neg && !b_neg) ? 1 : ((b_neg && !a_neg) ? 0 : a_lt_b);
  /* cc = the low 31 bytes of c as a field element; byte 31 is zero here, so cc == c */
  CHECK(EV_SCMP_EQ, fr_eq(c.lo, fr_u64(expect)) && fr_is_zero(c.hi));
  same_context(e, i, row, opcode, 3, one, one);
}
/* signextend.py: the byte-by-byte `is_equal` calls constrain nothing; what remains is the
 * sign_byte_lookup of the selected byte (signextend.py:44) */
static void gadget_signextend(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t index, value, result;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &index), EV_SEXT_IDX_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &value), EV_SEXT_VAL_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &result), EV_SEXT_RES_UNSAT, row)) return;
  CHECK(EV_SEXT_BYTES, word_in_domain(index) && word_in_domain(value) && word_in_domain(result));
  unsigned msb = 0;
  for (int k = 1; k < 32; k++) msb += word_byte(index, k);
  const unsigned idx0 = word_byte(index, 0);
  const unsigned sign_byte = idx0 < 31 ? (word_byte(value, (int)idx0) >> 7) * 0xFF : 0;
  const unsigned selected = (idx0 < 31 && msb == 0) ? word_byte(value, (int)idx0) : 0;
  fr_t key[4] = {fr_u64(ZK_FIXED_SignByte), fr_u64(selected), fr_u64(sign_byte), fr_u64(0)};
  if (!need1(e, orc_lookup(&e->fixed_ix, key, 0), EV_SEXT_SIGN_UNSAT, row)) return;
  same_context(e, i, row, opcode, 3, one, one);
}

/* block_ctx.py, origin.py, gasprice.py */
static void gadget_blockctx(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  uint64_t tag = 0;
  if (fr_fits_bits(opcode, 8)) switch (opcode.l[0]) {
    case 0x41: tag = 1; break; /* COINBASE -> Coinbase */
    case 0x42: tag = 4; break; /* TIMESTAMP */
    case 0x43: tag = 3; break; /* NUMBER */
    case 0x45: tag = 2; break; /* GASLIMIT */
    case 0x44: tag = 5; break; /* PREVRANDAO */
    case 0x48: tag = 6; break; /* BASEFEE */
    case 0x46: tag = 7; break; /* CHAINID */
    default: break;
  }
  CHECK(EV_BLK_OPCODE, tag != 0);
  fr_t key[2] = {fr_u64(tag), fr_u64(0)};
  uint32_t r; const int m = orc_lookup(&e->block_ix, key, &r);
  if (!need1(e, m, EV_BLK_CTX_UNSAT, row)) return;
  const fr_t lo = fr_load(ORC_CELL(e->block_ix.cells, e->block_ix.n_rows, 2, r)), hi = fr_load(ORC_CELL(e->block_ix.cells, e->block_ix.n_rows, 3, r));
  word_t w;
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_sub(CUR(S_SP), fr_u64(1)), &w), EV_BLK_PUSH_UNSAT, row)) return;
  CHECK(EV_BLK_EQ, fr_eq(w.lo, lo) && fr_eq(w.hi, hi));
  same_context(e, i, row, opcode, 1, fr_u64(1), fr_neg(fr_u64(1)));
}
static void gadget_txctx(evm_env* e, uint64_t i, uint64_t row, uint64_t op, uint64_t field) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  word_t v, w; int is_word;
  /* the call-context lookup comes BEFORE the opcode lookup here (origin.py:8-9) */
  if (!need1(e, call_context_w(e, CUR(S_RWC), 0, CUR(S_CALL_ID), ZK_CC_TxId, &v, &is_word), EV_TXC_TXID_UNSAT, row)) return;
  CHECK(EV_TXC_TXID_TYPE, !is_word);
  fr_t opcode;
  if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 2, CUR(S_PC), 1, &opcode), EV_OP_UNSAT, row)) return;
  CHECK(EV_TXC_OPCODE, fr_eq_u64(opcode, op));
  fr_t key[3] = {v.lo, fr_u64(field), fr_u64(0)};
  uint32_t r; const int m = orc_lookup(&e->tx_ix, key, &r);
  if (!need1(e, m, EV_TXC_TX_UNSAT, row)) return;
  const fr_t lo = fr_load(ORC_CELL(e->tx_ix.cells, e->tx_ix.n_rows, 3, r)), hi = fr_load(ORC_CELL(e->tx_ix.cells, e->tx_ix.n_rows, 4, r));
  if (!need1(e, rw_lookup(e, fr_add(CUR(S_RWC), fr_u64(1)), 1, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_sub(CUR(S_SP), fr_u64(1)), &w), EV_TXC_PUSH_UNSAT, row)) return;
  CHECK(EV_TXC_EQ, fr_eq(w.lo, lo) && fr_eq(w.hi, hi));
  same_context(e, i, row, opcode, 2, fr_u64(1), fr_neg(fr_u64(1)));
}

/* ---- shl_shr.py -------------------------------------------------------------------------------- */
/* unsigned big integers of 10 limbs (640 bits): Word.int_value() of arbitrary cells is < 2^382 and the
 * SHR remainder witness is dividend - quotient * 2^shift with shift < 256 */
typedef struct { uint64_t l[10]; } u640;
static u640 word_int(word_t w) { /* lo + hi * 2^128 as an integer */
  u640 r; memset(&r, 0, sizeof r);
  uint64_t c = 0;
  r.l[0] = w.lo.l[0]; r.l[1] = w.lo.l[1];
  r.l[2] = adc(w.lo.l[2], w.hi.l[0], &c); r.l[3] = adc(w.lo.l[3], w.hi.l[1], &c);
  r.l[4] = adc(w.hi.l[2], 0, &c); r.l[5] = adc(w.hi.l[3], 0, &c); r.l[6] = c;
  return r;
}
static u640 u640_shl(u640 a, unsigned s) { /* s < 256 */
  u640 r; memset(&r, 0, sizeof r);
  const unsigned ws = s >> 6, bs = s & 63;
  for (int k = 9; k >= 0; k--) {
    uint64_t v = 0;
    if (k >= (int)ws) {
      v = a.l[k - ws] << bs;
      if (bs && k - (int)ws - 1 >= 0) v |= a.l[k - ws - 1] >> (64 - bs);
    }
    r.l[k] = v;
  }
  return r;
}
static int u640_cmp(u640 a, u640 b) {
  for (int k = 9; k >= 0; k--) { if (a.l[k] < b.l[k]) return -1; if (a.l[k] > b.l[k]) return 1; }
  return 0;
}
static u640 u640_sub(u640 a, u640 b) {
  u640 r; uint64_t br = 0;
  for (int k = 0; k < 10; k++) {
    const u128 d = (u128)a.l[k] - b.l[k] - br;
    r.l[k] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
  }
  return r;
}
static void gadget_shl_shr(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  const word_t zero = {fr_u64(0), fr_u64(0)};
  word_t pop1, pop2, push;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &pop1), EV_SH_P1_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &pop2), EV_SH_P2_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, one), &push), EV_SH_PUSH_UNSAT, row)) return;
  /* gen_witness, shl_shr.py:103-129 */
  const fr_t is_shl = fr_sub(fr_u64(0x1c), opcode); /* Opcode.SHR - opcode over the field */
  CHECK(EV_SH_BYTES, word_in_domain(pop1));
  const unsigned shf0 = word_byte(pop1, 0);
  const int shf_lt256 = (pop1.lo.l[0] >> 8) == 0 && pop1.lo.l[1] == 0 && pop1.hi.l[0] == 0 && pop1.hi.l[1] == 0;
  word_t divisor = zero; /* Word(1 << shf0) if the shift is < 256 else Word(0) */
  if (shf_lt256) { if (shf0 < 128) divisor.lo.l[shf0 >> 6] = 1ull << (shf0 & 63); else divisor.hi.l[(shf0 - 128) >> 6] = 1ull << (shf0 & 63); }
  word_t dividend, quotient, remainder = zero;
  if (fr_eq_u64(is_shl, 1)) { dividend = push; quotient = pop2; }
  else {
    dividend = pop2; quotient = push;
    /* remainder = Word(dividend.int_value() - quotient.int_value() * divisor.int_value()) as Python ints */
    const u640 D = word_int(dividend);
    u640 QS; memset(&QS, 0, sizeof QS);
    if (shf_lt256) QS = u640_shl(word_int(quotient), shf0);
    if (u640_cmp(D, QS) < 0) { orc_fail(e->res, EV_SH_REM_NEG, row); return; } /* Word(negative).to_bytes */
    const u640 R = u640_sub(D, QS);
    CHECK(EV_SH_REM_WORD, (R.l[4] | R.l[5] | R.l[6] | R.l[7] | R.l[8] | R.l[9]) == 0); /* assert value < 256**32 */
    remainder.lo = fr_u128(R.l[0], R.l[1]); remainder.hi = fr_u128(R.l[2], R.l[3]);
  }
  /* check_witness, shl_shr.py:37-91 */
  const fr_t is_shr = fr_sub(one, is_shl);
  const int dz = fr_is_zero(fr_add(divisor.lo, divisor.hi));
  const fr_t nz = fr_u64(dz ? 0 : 1);
  word_t t1, t2, sum;
  /* :59-62 pop2 == quotient.select(is_shl) + dividend.select(is_shr) */
  CHECK(EV_SH_SELECT, word_select(quotient, is_shl, &t1) && word_select(dividend, is_shr, &t2));
  sum.lo = fr_add(t1.lo, t2.lo); sum.hi = fr_add(t1.hi, t2.hi);
  CHECK(EV_SH_SELECT, word_in_domain(sum));
  CHECK(EV_SH_POP2, word_eq(pop2, sum));
  /* :63-65 push == dividend.select(is_shl) + quotient.select(is_shr * (1 - divisor_is_zero)) */
  CHECK(EV_SH_SELECT, word_select(dividend, is_shl, &t1) && word_select(quotient, fr_mul(is_shr, nz), &t2));
  sum.lo = fr_add(t1.lo, t2.lo); sum.hi = fr_add(t1.hi, t2.hi);
  CHECK(EV_SH_SELECT, word_in_domain(sum));
  CHECK(EV_SH_PUSH_EQ, word_eq(push, sum));
  /* :66-76 hold by construction of shf0 / divisor (shift in the bytes domain) */
  /* :77-79 compare_word(remainder, divisor): both in the halves domain here */
  {
    const int hi_lt = fr_cmp(remainder.hi, divisor.hi) < 0, hi_eq = fr_eq(remainder.hi, divisor.hi);
    const int lo_lt = fr_cmp(remainder.lo, divisor.lo) < 0;
    CHECK(EV_SH_REM_LT, dz || (hi_lt + hi_eq * lo_lt) == 1);
  }
  CHECK(EV_SH_SHL_REM0, fr_is_zero(fr_mul(is_shl, fr_u64(fr_is_zero(fr_add(remainder.lo, remainder.hi)) ? 0 : 1))));
  /* :86 mul_add_words(quotient, divisor, remainder, dividend) */
  CHECK(EV_SH_TO64, word_in_domain(quotient));
  {
    const word_t a = quotient, b = divisor, c = remainder, d = dividend;
    fr_t a64[4] = {fr_u64(a.lo.l[0]), fr_u64(a.lo.l[1]), fr_u64(a.hi.l[0]), fr_u64(a.hi.l[1])};
    fr_t b64[4] = {fr_u64(b.lo.l[0]), fr_u64(b.lo.l[1]), fr_u64(b.hi.l[0]), fr_u64(b.hi.l[1])};
#define M(x, y) fr_mul(a64[x], b64[y])
    const fr_t t0 = M(0, 0), tt1 = fr_add(M(0, 1), M(1, 0));
    const fr_t tt2 = fr_add(fr_add(M(0, 2), M(1, 1)), M(2, 0));
    const fr_t tt3 = fr_add(fr_add(fr_add(M(0, 3), M(1, 2)), M(2, 1)), M(3, 0));
    const fr_t two64 = {{0, 1, 0, 0}};
    const fr_t carry_lo = fr_mul(fr_sub(fr_add(fr_add(t0, fr_mul(tt1, two64)), c.lo), d.lo), INV_2_128);
    const fr_t carry_hi = fr_mul(fr_sub(fr_add(fr_add(fr_add(tt2, fr_mul(tt3, two64)), c.hi), carry_lo), d.hi), INV_2_128);
    fr_t overflow = carry_hi;
    overflow = fr_add(overflow, M(1, 3)); overflow = fr_add(overflow, M(2, 2));
    overflow = fr_add(overflow, M(3, 1)); overflow = fr_add(overflow, M(2, 3));
    overflow = fr_add(overflow, M(3, 2)); overflow = fr_add(overflow, M(3, 3));
#undef M
    if (!fr_fits_bits(carry_lo, 72)) { orc_fail(e->res, EV_SH_CARRY_LO, row); return; }
    if (!fr_fits_bits(carry_hi, 72)) { orc_fail(e->res, EV_SH_CARRY_HI, row); return; }
    CHECK(EV_SH_OVERFLOW, fr_is_zero(fr_mul(is_shr, overflow)));
  }
  if (!dz) { /* :90-91 pow2_lookup(shf0, divisor_lo, divisor_hi) */
    fr_t key[4] = {fr_u64(ZK_FIXED_Pow2), fr_u64(shf0), divisor.lo, divisor.hi};
    if (!need1(e, orc_lookup(&e->fixed_ix, key, 0), EV_SH_POW2_UNSAT, row)) return;
  }
  same_context(e, i, row, opcode, 3, one, one);
}

static int state_is(fr_t s, uint64_t v) { return fr_eq_u64(s, v); }
#include "evm_tx.h"
#include "evm_err.h"
#include "evm_arith.h"
#include "evm_storage.h"
#include "evm_log.h"
#include "evm_exp.h"
#include "evm_return.h"
#include "evm_call.h"
#include "evm_create.h"

static void verify_step(evm_env* e, uint64_t i, uint64_t row, uint32_t flags) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t cs = CUR(S_STATE), ns = NXT(S_STATE);
  const int is_first = (flags & 2) && row == 0;
  const int is_last = (flags & 4) && i == n - 2;
  if (is_first) {
    CHECK(EV_FIRST_STATE, state_is(cs, ZK_ES_BeginTx) || state_is(cs, ZK_ES_EndBlock));
    CHECK(EV_FIRST_RWC, fr_eq_u64(CUR(S_RWC), 1));
  }
  if (is_last) CHECK(EV_LAST_STATE, state_is(cs, ZK_ES_EndBlock));
  else {
    const int c_halts = fr_fits_bits(cs, 16) && cs.l[0] < ZK_ES_COUNT && ES_HALTS[cs.l[0]];
    if (state_is(cs, ZK_ES_EndTx))
      CHECK(EV_TRANS_FROM_ENDTX, state_is(ns, ZK_ES_BeginTx) || state_is(ns, ZK_ES_EndBlock));
    else if (state_is(cs, ZK_ES_EndBlock))
      CHECK(EV_TRANS_FROM_ENDBLOCK, state_is(ns, ZK_ES_EndBlock));
    if (state_is(ns, ZK_ES_BeginTx)) CHECK(EV_TRANS_TO_BEGINTX, state_is(cs, ZK_ES_EndTx));
    else if (state_is(ns, ZK_ES_EndTx)) CHECK(EV_TRANS_TO_ENDTX, c_halts || state_is(cs, ZK_ES_BeginTx));
    else if (state_is(ns, ZK_ES_EndBlock))
      CHECK(EV_TRANS_TO_ENDBLOCK, state_is(cs, ZK_ES_EndTx) || state_is(cs, ZK_ES_EndBlock));
  }
  CHECK(EV_NOT_IMPLEMENTED, fr_fits_bits(cs, 16) && cs.l[0] < ZK_ES_COUNT && ES_IMPL[cs.l[0]]);
  const uint64_t st = cs.l[0];
  CHECK(EV_UNSUPPORTED_STATE, st == ZK_ES_ADD || st == ZK_ES_MUL || st == ZK_ES_PUSH || st == ZK_ES_POP ||
                                  st == ZK_ES_SHA3 || st == ZK_ES_CALLDATACOPY || st == ZK_ES_STOP ||
                                  st == ZK_ES_MEMORY || st == ZK_ES_MSIZE || st == ZK_ES_GAS || st == ZK_ES_ISZERO ||
                                  st == ZK_ES_CMP || st == ZK_ES_JUMP || st == ZK_ES_JUMPI || st == ZK_ES_CALLER ||
                                  st == ZK_ES_CALLVALUE || st == ZK_ES_CALLDATASIZE || st == ZK_ES_ADDRESS ||
                                  st == ZK_ES_RETURNDATASIZE || st == ZK_ES_CODESIZE || st == ZK_ES_BITWISE ||
                                  st == ZK_ES_NOT || st == ZK_ES_BYTE || st == ZK_ES_SCMP || st == ZK_ES_SIGNEXTEND ||
                                  st == ZK_ES_BlockCtx || st == ZK_ES_ORIGIN || st == ZK_ES_GASPRICE ||
                                  st == ZK_ES_SHL_SHR || st == ZK_ES_BeginTx || st == ZK_ES_EndTx || st == ZK_ES_EndBlock ||
                                  st == ZK_ES_ErrorStack || st == ZK_ES_ErrorInvalidOpcode || st == ZK_ES_ErrorOutOfGasConstant ||
                                  st == ZK_ES_ErrorInvalidJump || st == ZK_ES_SELFBALANCE || st == ZK_ES_ErrorOutOfGasSHA3 ||
                                  st == ZK_ES_ErrorOutOfGasStaticMemoryExpansion || st == ZK_ES_ErrorOutOfGasDynamicMemoryExpansion ||
                                  st == ZK_ES_ErrorOutOfGasLOG || st == ZK_ES_ErrorOutOfGasEXP || st == ZK_ES_ErrorReturnDataOutOfBound ||
                                  st == ZK_ES_BALANCE || st == ZK_ES_EXTCODEHASH || st == ZK_ES_EXTCODESIZE ||
                                  st == ZK_ES_ErrorOutOfGasAccountAccess || st == ZK_ES_CODECOPY || st == ZK_ES_RETURNDATACOPY ||
                                  st == ZK_ES_EXTCODECOPY || st == ZK_ES_ErrorOutOfGasMemoryCopy || st == ZK_ES_ADDMOD ||
                                  st == ZK_ES_MULMOD || st == ZK_ES_SDIV_SMOD || st == ZK_ES_SAR || st == ZK_ES_SLOAD || st == ZK_ES_SSTORE ||
                                  st == ZK_ES_CALLDATALOAD || st == ZK_ES_LOG || st == ZK_ES_ErrorWriteProtection || st == ZK_ES_BLOCKHASH ||
                                  st == ZK_ES_EXP || st == ZK_ES_ErrorMaxCodeSizeExceeded || st == ZK_ES_ErrorOutOfGasCodeStore ||
                                  st == ZK_ES_ErrorInvalidCreationCode || st == ZK_ES_RETURN ||
                                  st == ZK_ES_ErrorOutOfGasCall || st == ZK_ES_CALL_OP || st == ZK_ES_CREATE || st == ZK_ES_CREATE2 ||
                                  st == ZK_ES_ErrorOutOfGasSloadSstore || st == ZK_ES_ErrorOutOfGasCREATE ||
                                  st == ZK_ES_ErrorOutOfGasPrecompile || st == ZK_ES_ErrorGasUintOverflow);
  if (st == ZK_ES_BeginTx) { gadget_begin_tx(e, i, row, is_first); return; }
  if (st == ZK_ES_EndTx) { gadget_end_tx(e, i, row); return; }
  if (st == ZK_ES_EndBlock) { gadget_end_block(e, i, row, is_last); return; }
  if (st == ZK_ES_STOP) { gadget_stop(e, i, row); return; }
  if (st == ZK_ES_ErrorOutOfGasPrecompile) { gadget_error_oog_precompile(e, i, row); return; }
  if (st == ZK_ES_ORIGIN) { gadget_txctx(e, i, row, 0x32, ZK_TX_CallerAddress); return; }
  if (st == ZK_ES_GASPRICE) { gadget_txctx(e, i, row, 0x3a, ZK_TX_GasPrice); return; }
  fr_t opcode;
  if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 2, CUR(S_PC), 1, &opcode), EV_OP_UNSAT, row))
    return;
  if (st == ZK_ES_ADD) gadget_add(e, i, row, opcode);
  else if (st == ZK_ES_MUL) gadget_mul(e, i, row, opcode);
  else if (st == ZK_ES_PUSH) gadget_push(e, i, row, opcode);
  else if (st == ZK_ES_SHA3) gadget_sha3(e, i, row, opcode);
  else if (st == ZK_ES_CALLDATACOPY) gadget_calldatacopy(e, i, row, opcode);
  else if (st == ZK_ES_MEMORY) gadget_memory(e, i, row, opcode);
  else if (st == ZK_ES_MSIZE) gadget_msize(e, i, row, opcode);
  else if (st == ZK_ES_GAS) gadget_gas(e, i, row, opcode);
  else if (st == ZK_ES_ISZERO) gadget_iszero(e, i, row, opcode);
  else if (st == ZK_ES_CMP) gadget_cmp(e, i, row, opcode);
  else if (st == ZK_ES_JUMP) gadget_jump(e, i, row, opcode);
  else if (st == ZK_ES_JUMPI) gadget_jumpi(e, i, row, opcode);
  else if (st == ZK_ES_CALLER) gadget_cc_push(e, i, row, opcode, 0x33, ZK_CC_CallerAddress, 1);
  else if (st == ZK_ES_CALLVALUE) gadget_cc_push(e, i, row, opcode, 0x34, ZK_CC_Value, 1);
  else if (st == ZK_ES_CALLDATASIZE) gadget_cc_push(e, i, row, opcode, 0x36, ZK_CC_CallDataLength, 0);
  else if (st == ZK_ES_ADDRESS) gadget_cc_push(e, i, row, opcode, 0x30, ZK_CC_CalleeAddress, 1);
  else if (st == ZK_ES_RETURNDATASIZE) gadget_cc_push(e, i, row, opcode, 0x3d, ZK_CC_LastCalleeReturnDataLength, 0);
  else if (st == ZK_ES_CODESIZE) gadget_codesize(e, i, row, opcode);
  else if (st == ZK_ES_BITWISE) gadget_bitwise(e, i, row, opcode);
  else if (st == ZK_ES_NOT) gadget_not(e, i, row, opcode);
  else if (st == ZK_ES_BYTE) gadget_byte(e, i, row, opcode);
  else if (st == ZK_ES_SCMP) gadget_scmp(e, i, row, opcode);
  else if (st == ZK_ES_SIGNEXTEND) gadget_signextend(e, i, row, opcode);
  else if (st == ZK_ES_BlockCtx) gadget_blockctx(e, i, row, opcode);
  else if (st == ZK_ES_SHL_SHR) gadget_shl_shr(e, i, row, opcode);
  else if (st == ZK_ES_ErrorStack) gadget_error_stack(e, i, row, opcode);
  else if (st == ZK_ES_ErrorInvalidOpcode) gadget_error_invalid_opcode(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasConstant) gadget_error_oog_constant(e, i, row, opcode);
  else if (st == ZK_ES_ErrorInvalidJump) gadget_error_invalid_jump(e, i, row, opcode);
  else if (st == ZK_ES_SELFBALANCE) gadget_selfbalance(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasSHA3) gadget_error_oog_sha3(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasStaticMemoryExpansion) gadget_error_oog_static_memory(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasDynamicMemoryExpansion) gadget_error_oog_dynamic_memory(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasLOG) gadget_error_oog_log(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasEXP) gadget_error_oog_exp(e, i, row, opcode);
  else if (st == ZK_ES_ErrorReturnDataOutOfBound) gadget_error_return_data_oob(e, i, row, opcode);
  else if (st == ZK_ES_BALANCE) gadget_account_access(e, i, row, opcode, 0x31);
  else if (st == ZK_ES_EXTCODEHASH) gadget_account_access(e, i, row, opcode, 0x3f);
  else if (st == ZK_ES_EXTCODESIZE) gadget_account_access(e, i, row, opcode, 0x3b);
  else if (st == ZK_ES_ErrorOutOfGasAccountAccess) gadget_error_oog_account_access(e, i, row, opcode);
  else if (st == ZK_ES_CODECOPY) gadget_codecopy(e, i, row, opcode);
  else if (st == ZK_ES_RETURNDATACOPY) gadget_returndatacopy(e, i, row, opcode);
  else if (st == ZK_ES_EXTCODECOPY) gadget_extcodecopy(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasMemoryCopy) gadget_error_oog_memory_copy(e, i, row, opcode);
  else if (st == ZK_ES_ADDMOD) gadget_addmod_mulmod(e, i, row, opcode, 0);
  else if (st == ZK_ES_MULMOD) gadget_addmod_mulmod(e, i, row, opcode, 1);
  else if (st == ZK_ES_SDIV_SMOD) gadget_sdiv_smod(e, i, row, opcode);
  else if (st == ZK_ES_SAR) gadget_sar(e, i, row, opcode);
  else if (st == ZK_ES_SLOAD) gadget_sload(e, i, row, opcode);
  else if (st == ZK_ES_SSTORE) gadget_sstore(e, i, row, opcode);
  else if (st == ZK_ES_CALLDATALOAD) gadget_calldataload(e, i, row, opcode);
  else if (st == ZK_ES_LOG) gadget_log(e, i, row, opcode);
  else if (st == ZK_ES_ErrorWriteProtection) gadget_error_write_protection(e, i, row, opcode);
  else if (st == ZK_ES_BLOCKHASH) gadget_blockhash(e, i, row, opcode);
  else if (st == ZK_ES_EXP) gadget_exp(e, i, row, opcode);
  else if (st == ZK_ES_ErrorMaxCodeSizeExceeded || st == ZK_ES_ErrorOutOfGasCodeStore) gadget_error_code_store(e, i, row, opcode);
  else if (st == ZK_ES_ErrorInvalidCreationCode) gadget_error_invalid_creation_code(e, i, row, opcode);
  else if (st == ZK_ES_RETURN) gadget_return_revert(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasCall) gadget_error_oog_call(e, i, row, opcode);
  else if (st == ZK_ES_CALL_OP) gadget_callop(e, i, row, opcode);
  else if (st == ZK_ES_CREATE || st == ZK_ES_CREATE2) gadget_create(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasSloadSstore) gadget_error_oog_sload_sstore(e, i, row, opcode);
  else if (st == ZK_ES_ErrorOutOfGasCREATE) gadget_error_oog_create(e, i, row, opcode);
  else if (st == ZK_ES_ErrorGasUintOverflow) gadget_error_gas_uint_overflow(e, i, row, opcode);
  else gadget_pop(e, i, row, opcode);
}

int orc_check_evm_x(const uint64_t* steps, uint64_t n_steps, const uint64_t* bytecode_tab, uint64_t n_bytecode,
                    const uint64_t* rw_tab, uint64_t n_rw, const uint8_t* rw_flags, const uint64_t* fixed_tab,
                    uint64_t n_fixed, const uint64_t* copy_tab, uint64_t n_copy, const uint64_t* keccak_tab,
                    uint64_t n_keccak, uint64_t row_begin, uint64_t row_end, uint64_t row_base, uint32_t flags,
                    uint32_t* first_fail, uint64_t* fail_count);
int orc_check_evm(const uint64_t* steps, uint64_t n_steps, const uint64_t* bytecode_tab, uint64_t n_bytecode,
                  const uint64_t* rw_tab, uint64_t n_rw, const uint64_t* fixed_tab, uint64_t n_fixed,
                  uint64_t row_begin, uint64_t row_end, uint64_t row_base, uint32_t flags,
                  uint32_t* first_fail, uint64_t* fail_count) {
  return orc_check_evm_x(steps, n_steps, bytecode_tab, n_bytecode, rw_tab, n_rw, 0, fixed_tab, n_fixed, 0, 0, 0, 0,
                         row_begin, row_end, row_base, flags, first_fail, fail_count);
}
/* tx table (5 cells: tx_id, tag, index, value lo, hi) and block table (4 cells: tag, block number, value
 * lo, hi) of the NEXT orc_check_evm_x call on this thread (the ORIGIN / GASPRICE / BlockCtx gadgets) */
static __thread const uint64_t* g_tx_tab; static __thread uint64_t g_n_tx;
static __thread const uint64_t* g_block_tab; static __thread uint64_t g_n_block;
void orc_set_evm_context_tables(const uint64_t* tx_tab, uint64_t n_tx, const uint64_t* block_tab, uint64_t n_block) {
  g_tx_tab = tx_tab; g_n_tx = n_tx; g_block_tab = block_tab; g_n_block = n_block;
}
/* exp table (11 cells) of the NEXT call (EXP) */
static __thread const uint64_t* g_exp_tab; static __thread uint64_t g_n_exp;
void orc_set_evm_exp_table(const uint64_t* exp_tab, uint64_t n_exp) { g_exp_tab = exp_tab; g_n_exp = n_exp; }
/* StepState.aux_data of the NEXT call: rows of (step row, lo, hi), column-major like every table (CREATE / CREATE2) */
static __thread const uint64_t* g_aux_tab; static __thread uint64_t g_n_aux;
void orc_set_evm_step_aux(const uint64_t* aux_tab, uint64_t n_aux) { g_aux_tab = aux_tab; g_n_aux = n_aux; }
/* value type flags of the tx / block tables and the withdrawal table of the NEXT call (BeginTx / EndTx / EndBlock) */
static __thread const uint8_t *g_tx_flags, *g_block_flags; static __thread const uint64_t* g_wd_tab; static __thread uint64_t g_n_wd;
void orc_set_evm_block_tables(const uint8_t* tx_flags, const uint8_t* block_flags, const uint64_t* wd_tab, uint64_t n_wd) {
  g_tx_flags = tx_flags; g_block_flags = block_flags; g_wd_tab = wd_tab; g_n_wd = n_wd;
}
int orc_check_evm_x(const uint64_t* steps, uint64_t n_steps, const uint64_t* bytecode_tab, uint64_t n_bytecode,
                    const uint64_t* rw_tab, uint64_t n_rw, const uint8_t* rw_flags, const uint64_t* fixed_tab,
                    uint64_t n_fixed, const uint64_t* copy_tab, uint64_t n_copy, const uint64_t* keccak_tab,
                    uint64_t n_keccak, uint64_t row_begin, uint64_t row_end, uint64_t row_base, uint32_t flags,
                    uint32_t* first_fail, uint64_t* fail_count) {
  orc_result res; orc_result_init(&res, first_fail, fail_count, EV_N_CONSTRAINTS);
  if (row_end + 1 > n_steps && row_end > row_begin) return -1;
  evm_env env; env.steps = steps; env.n_steps = n_steps; env.res = &res;
  const uint32_t bk[5] = {0, 1, 2, 3, 4}, rk[5] = {0, 1, 2, 3, 4}, fk[4] = {0, 1, 2, 3};
  orc_index_build(&env.bytecode_ix, bytecode_tab, n_bytecode, 6, bk, 5);
  orc_index_build(&env.rw_ix, rw_tab, n_rw, 14, rk, 5);
  /* the fixed table (224,490 rows) is the same array call after call in the test-suite: keep its
   * sorted order, keyed on (pointer, rows, a checksum of sampled rows) */
  static __thread struct { const uint64_t* p; uint64_t n, sum; uint32_t* order; } fx_cache;
  uint64_t fx_sum = 0;
  for (uint64_t k = 0; k < 257 && n_fixed; k++) {
    const uint64_t r = (n_fixed - 1) * k / 256;
    for (uint32_t c = 0; c < 4; c++) fx_sum = fx_sum * 0x9E3779B97F4A7C15ull + ORC_CELL(fixed_tab, n_fixed, c, r)[0];
  }
  if (fx_cache.order && fx_cache.p == fixed_tab && fx_cache.n == n_fixed && fx_cache.sum == fx_sum) {
    env.fixed_ix.cells = fixed_tab; env.fixed_ix.n_rows = n_fixed; env.fixed_ix.n_cols = 4; env.fixed_ix.n_key = 4;
    for (uint32_t k = 0; k < 4; k++) env.fixed_ix.key_cols[k] = fk[k];
    env.fixed_ix.order = fx_cache.order;
  } else {
    orc_index_build(&env.fixed_ix, fixed_tab, n_fixed, 4, fk, 4);
    free(fx_cache.order);
    fx_cache.p = fixed_tab; fx_cache.n = n_fixed; fx_cache.sum = fx_sum; fx_cache.order = env.fixed_ix.order;
  }
  const uint32_t ck[11] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12}, kk[3] = {0, 1, 2};
  orc_index_build(&env.copy_ix, copy_tab, n_copy, 14, ck, 11);
  orc_index_build(&env.keccak_ix, keccak_tab, n_keccak, 5, kk, 3);
  const uint32_t tk[3] = {0, 1, 2}, blk[2] = {0, 1};
  orc_index_build(&env.tx_ix, g_tx_tab, g_n_tx, 5, tk, 3);
  orc_index_build(&env.block_ix, g_block_tab, g_n_block, 4, blk, 2);
  g_tx_tab = g_block_tab = 0; g_n_tx = g_n_block = 0;
  const uint32_t ek[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
  orc_index_build(&env.exp_ix, g_exp_tab, g_n_exp, 11, ek, 9);
  g_exp_tab = 0; g_n_exp = 0;
  env.aux = g_aux_tab; env.n_aux = g_n_aux; g_aux_tab = 0; g_n_aux = 0;
  const uint32_t ck0[1] = {0};
  orc_index_build(&env.rwc_ix, rw_tab, n_rw, 14, ck0, 1);
  env.tx_flags = g_tx_flags; env.block_flags = g_block_flags; env.wd_tab = g_wd_tab; env.n_wd = g_n_wd;
  g_tx_flags = g_block_flags = 0; g_wd_tab = 0; g_n_wd = 0;
  env.rw_flags = rw_flags;
  for (uint64_t i = row_begin; i < row_end; i++) verify_step(&env, i, row_base + i, flags);
  orc_index_free(&env.bytecode_ix); orc_index_free(&env.rw_ix); /* fixed_ix.order lives in fx_cache */
  orc_index_free(&env.copy_ix); orc_index_free(&env.keccak_ix); orc_index_free(&env.tx_ix); orc_index_free(&env.block_ix);
  orc_index_free(&env.rwc_ix); orc_index_free(&env.exp_ix);
  return 0;
}

/*
 * oracle/evm_err.h — TEST INFRASTRUCTURE (CPU oracle), included by evm.c.
 *
 * Restates constrain_error_state (/root/reference/src/zkevm_specs/evm_circuit/instruction.py:1426-1452) and the
 * gadgets execution/error_stack.py, error_invalid_opcode.py, error_oog_constant.py, error_invalid_jump.py and
 * selfbalance.py.  Pinned by tests/golden/evm12.npz.
 */

/* constrain_error_state(rw_counter_offset + curr.reversible_write_counter): `n_rw` = rw lookups the gadget did */
static void error_state_tail(evm_env* e, uint64_t i, uint64_t row, uint64_t n_rw) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC), rev = CUR(S_REV);
  word_t v; int w;
  if (!need1(e, call_context_w(e, fr_add(rwc, fr_u64(n_rw)), 0, CUR(S_CALL_ID), ZK_CC_IsSuccess, &v, &w), EV_ERR_CC_UNSAT, row)) return;
  CHECK(EV_ERR_CC_TYPE, !w);
  CHECK(EV_ERR_IS_SUCCESS, fr_is_zero(v.lo));
  const fr_t is_root = CUR(S_IS_ROOT);
  CHECK(EV_ERR_ROOT_ENDTX, fr_eq_u64(is_root, fr_eq_u64(NXT(S_STATE), ZK_ES_EndTx) ? 1 : 0));
  if (!fr_is_zero(is_root)) {
    CHECK(EV_ERR_RWC, fr_eq(NXT(S_RWC), fr_add(fr_add(rwc, fr_u64(n_rw + 1)), rev)));
    CHECK(EV_ERR_CALL_ID, fr_eq(NXT(S_CALL_ID), CUR(S_CALL_ID)));
  } else {
    restore_context_x(e, i, row, n_rw + 1, fr_u64(0), fr_u64(0), fr_u64(0), 0, rev);
  }
}

/* error_stack.py */
static void gadget_error_stack(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  fr_t key[4] = {fr_u64(ZK_FIXED_ResponsibleOpcode), CUR(S_STATE), opcode, CUR(S_SP)};
  CHECK(EV_ESTK_RESP_OPCODE, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  error_state_tail(e, i, row, 0);
}
/* error_invalid_opcode.py */
static void gadget_error_invalid_opcode(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  fr_t key[4] = {fr_u64(ZK_FIXED_ResponsibleOpcode), CUR(S_STATE), opcode, fr_u64(0)};
  CHECK(EV_EINV_RESP_OPCODE, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  error_state_tail(e, i, row, 0);
}
/* error_oog_constant.py */
static void gadget_error_oog_constant(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_EOGC_OPCODE_VALUE, fr_fits_bits(opcode, 8) && OPCODE_GAS[opcode.l[0]] >= 0);
  const fr_t gas = fr_u64((uint64_t)OPCODE_GAS[opcode.l[0]]);
  fr_t key[4] = {fr_u64(ZK_FIXED_OpcodeConstantGas), opcode, gas, fr_u64(0)};
  CHECK(EV_EOGC_GAS_UNSAT, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  const fr_t gas_left = CUR(S_GAS);
  CHECK(EV_EOGC_CMP_RANGE, fr_fits_bits(gas_left, 64) && fr_fits_bits(gas, 64));
  CHECK(EV_EOGC_NOT_ENOUGH, gas_left.l[0] < gas.l[0]);
  error_state_tail(e, i, row, 0);
}
/* error_invalid_jump.py: NB constrain_error_state sits INSIDE `if within_range == FQ(1)` (:24-33) — a destination at
 * or beyond the code length leaves the step's ending unconstrained */
static void gadget_error_invalid_jump(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  CHECK(EV_EJMP_OPCODE, fr_eq_u64(opcode, 0x56) || fr_eq_u64(opcode, 0x57));
  const int is_jumpi = fr_eq_u64(opcode, 0x57);
  fr_t code_length;
  if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 1, fr_u64(0), 0, &code_length), EV_EJMP_LEN_UNSAT, row)) return;
  word_t dest, cond;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &dest), EV_EJMP_DEST_UNSAT, row)) return;
  if (is_jumpi) {
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(1)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(1)), &cond), EV_EJMP_COND_UNSAT, row)) return;
    CHECK(EV_EJMP_COND_ZERO, !(fr_is_zero(cond.lo) && fr_is_zero(cond.hi)));
  }
  /* word_to_u64 = word_to_fq(word, 8) (instruction.py:480-484) */
  CHECK(EV_EJMP_DEST_DOMAIN, word_in_domain(dest));
  CHECK(EV_EJMP_DEST_U64, !(dest.lo.l[1] || dest.hi.l[0] || dest.hi.l[1]));
  const uint64_t d = dest.lo.l[0];
  CHECK(EV_EJMP_CMP_RANGE, fr_fits_bits(code_length, 64));
  if (d < code_length.l[0]) {
    /* bytecode_lookup_pair: key (hash, Byte, index), is_code not queried */
    fr_t key[4] = {CUR(S_HASH_LO), CUR(S_HASH_HI), fr_u64(2), fr_u64(d)};
    uint32_t r = 0; const int nn = orc_lookup_prefix(&e->bytecode_ix, key, 4, &r);
    if (!need1(e, nn, EV_EJMP_AT_UNSAT, row)) return;
    const fr_t value = fr_load(ORC_CELL(e->bytecode_ix.cells, e->bytecode_ix.n_rows, B_VALUE, r));
    const fr_t is_code = fr_load(ORC_CELL(e->bytecode_ix.cells, e->bytecode_ix.n_rows, B_ISCODE, r));
    /* is_code * FQ(value == JUMPDEST) == 0 */
    CHECK(EV_EJMP_IS_JUMPDEST, fr_is_zero(is_code) || !fr_eq_u64(value, 0x5b));
    error_state_tail(e, i, row, 1 + (uint64_t)is_jumpi);
  }
}
/* selfbalance.py */
static void gadget_selfbalance(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID);
  CHECK(EV_SBAL_OPCODE, fr_eq_u64(opcode, 0x47));
  uint32_t r;
  LK(cc_lookup(e, rwc, call_id, ZK_CC_CalleeAddress, &r), EV_SBAL_CC_UNSAT);
  const word_t callee = rw_value(e, r);
  /* word_to_address = word_to_fq(word, 20) */
  CHECK(EV_SBAL_ADDR_DOMAIN, word_in_domain(callee));
  CHECK(EV_SBAL_ADDR_RANGE, !((callee.hi.l[0] >> 32) || callee.hi.l[1]));
  fr_t address = {{callee.lo.l[0], callee.lo.l[1], callee.hi.l[0], 0}};
  LK(account_lookup(e, fr_add(rwc, fr_u64(1)), 0, address, ZK_ACC_Balance, &r), EV_SBAL_ACC_UNSAT);
  const word_t balance = rw_value(e, r);
  word_t w;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 1, ZK_TARGET_Stack, call_id, fr_sub(CUR(S_SP), fr_u64(1)), &w), EV_SBAL_PUSH_UNSAT, row)) return;
  CHECK(EV_SBAL_EQ, word_eq(w, balance));
  same_context(e, i, row, opcode, 3, fr_u64(1), fr_neg(fr_u64(1)));
}

/* ---- out-of-gas / out-of-bound error states ------------------------------------------------------------------ */
/* word_to_fq(word, n_bytes) for n_bytes <= 31: 0 ok, 1 -> OverflowError, 2 -> ConstraintUnsatFailure */
static int word_to_fq_nb(word_t w, int n_bytes, fr_t* out) {
  if (!word_in_domain(w)) return 1;
  const uint64_t v[4] = {w.lo.l[0], w.lo.l[1], w.hi.l[0], w.hi.l[1]};
  fr_t r = fr_u64(0);
  for (int k = 0; k < 32; k++) {
    const uint64_t b = (v[k >> 3] >> (8 * (k & 7))) & 0xFF;
    if (k >= n_bytes) { if (b) return 2; } else r.l[k >> 3] |= b << (8 * (k & 7));
  }
  *out = r;
  return 0;
}
#define W2FQ(word, nb, out, id_domain) do { const int rc_ = word_to_fq_nb((word), (nb), (out)); \
  if (rc_) { orc_fail(e->res, rc_ == 1 ? (id_domain) : (id_domain) + 1, row); return; } } while (0)
/* memory size in words of [offset, offset + length) and the gas of growing the memory to it (memory_expansion /
 * memory_expansion_dynamic_length, instruction.py:1138-1177); offset, length < 2^40 */
static int mem_expansion_gas(evm_env* e, uint64_t i, uint64_t words_needed, uint64_t* gas) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  if (words_needed >> 32) return 1;  /* range_check(memory_size, 4) */
  const fr_t cur = CUR(S_MEM);
  if (!fr_fits_bits(cur, 32)) return 2;  /* max(): compare asserts 4 bytes */
  const uint64_t nxt = cur.l[0] < words_needed ? words_needed : cur.l[0];
  *gas = memory_gas_cost(nxt) - memory_gas_cost(cur.l[0]);
  return 0;
}
#define MEMGAS(words, gas) do { const int rc_ = mem_expansion_gas(e, i, (words), (gas)); \
  if (rc_) { orc_fail(e->res, rc_ == 1 ? EV_EOOG_MEMSIZE_RANGE : EV_EOOG_MEM_MAX, row); return; } } while (0)
/* compare(gas_left, cost, 8) == lt, then constrain_error_state */
static void oog_finish(evm_env* e, uint64_t i, uint64_t row, u128 cost, uint64_t n_rw) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t gas_left = CUR(S_GAS);
  CHECK(EV_EOOG_CMP_RANGE, fr_fits_bits(gas_left, 64) && !(cost >> 64));
  CHECK(EV_EOOG_NOT_ENOUGH, gas_left.l[0] < (uint64_t)cost);
  error_state_tail(e, i, row, n_rw);
}
#define POP(k, out, id) do { if (!need1(e, rw_lookup(e, fr_add(CUR(S_RWC), fr_u64(k)), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), \
  fr_add(CUR(S_SP), fr_u64(k)), (out)), (id), row)) return; } while (0)

/* error_oog_sha3.py */
static void gadget_error_oog_sha3(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x20));
  word_t off_w, size_w;
  POP(0, &off_w, EV_EOOG_POP0_UNSAT);
  POP(1, &size_w, EV_EOOG_POP1_UNSAT);
  fr_t length, offset = fr_u64(0);  /* memory_offset_and_length: the length first (the second word) */
  W2FQ(size_w, 5, &length, EV_EOOG_W1_DOMAIN);
  if (!fr_is_zero(length)) W2FQ(off_w, 5, &offset, EV_EOOG_W0_DOMAIN);
  uint64_t expansion;
  MEMGAS((offset.l[0] + length.l[0] + 31) / 32, &expansion);
  const uint64_t words = (length.l[0] + 31) / 32;
  CHECK(EV_EOOG_WORDSIZE_RANGE, !(words >> 32));
  oog_finish(e, i, row, (u128)30 + (u128)words * 6 + expansion, 2);
}
/* error_oog_static_memory_expansion.py: `size = 1 if is_mstore8 else 32` tests the truthiness of an FQ object, which
 * is always true — the reference charges the expansion of ONE byte whatever the opcode (reproduced as written) */
static void gadget_error_oog_static_memory(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x51) || fr_eq_u64(opcode, 0x52) || fr_eq_u64(opcode, 0x53));
  word_t off_w;
  POP(0, &off_w, EV_EOOG_POP0_UNSAT);
  fr_t offset;
  W2FQ(off_w, 5, &offset, EV_EOOG_W0_DOMAIN);
  uint64_t expansion;
  MEMGAS((offset.l[0] + 1 + 31) / 32, &expansion);
  oog_finish(e, i, row, (u128)3 + expansion, 1);
}
/* error_oog_dynamic_memory_expansion.py */
static void gadget_error_oog_dynamic_memory(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0xf3) || fr_eq_u64(opcode, 0xfd));
  word_t off_w, size_w;
  POP(0, &off_w, EV_EOOG_POP0_UNSAT);
  POP(1, &size_w, EV_EOOG_POP1_UNSAT);
  fr_t length, offset = fr_u64(0);
  W2FQ(size_w, 5, &length, EV_EOOG_W1_DOMAIN);
  if (!fr_is_zero(length)) W2FQ(off_w, 5, &offset, EV_EOOG_W0_DOMAIN);
  uint64_t expansion;  /* memory_expansion: size 0 when length == 0 */
  MEMGAS(fr_is_zero(length) ? 0 : (offset.l[0] + length.l[0] + 31) / 32, &expansion);
  oog_finish(e, i, row, expansion, 2);
}
/* error_oog_log.py */
static void gadget_error_oog_log(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t topics = fr_sub(opcode, fr_u64(0xa0));
  {
    fr_t key[4] = {fr_u64(ZK_FIXED_Range5), topics, fr_u64(0), fr_u64(0)};
    CHECK(EV_EOOG_LOG_RANGE5, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  }
  word_t start_w, size_w;
  POP(0, &start_w, EV_EOOG_POP0_UNSAT);
  fr_t mstart, msize;
  W2FQ(start_w, 5, &mstart, EV_EOOG_W0_DOMAIN);
  POP(1, &size_w, EV_EOOG_POP1_UNSAT);
  W2FQ(size_w, 5, &msize, EV_EOOG_W1_DOMAIN);
  uint64_t expansion;
  MEMGAS((mstart.l[0] + msize.l[0] + 31) / 32, &expansion);
  /* topics is a fixed-table value here (0..4 with the reference's table), the cost stays far below 2^64 */
  oog_finish(e, i, row, (u128)375 + (u128)375 * topics.l[0] + (u128)8 * msize.l[0] + expansion, 2);
}
/* error_oog_exp.py */
static void gadget_error_oog_exp(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x0a));
  word_t expo;  /* stack_lookup(Read, 1): the first rw lookup of the step, at stack_pointer + 1 */
  if (!need1(e, rw_lookup(e, CUR(S_RWC), 0, ZK_TARGET_Stack, CUR(S_CALL_ID), fr_add(CUR(S_SP), fr_u64(1)), &expo), EV_EOOG_POP0_UNSAT, row)) return;
  CHECK(EV_EOOG_W0_DOMAIN, word_in_domain(expo));
  const uint64_t v[4] = {expo.lo.l[0], expo.lo.l[1], expo.hi.l[0], expo.hi.l[1]};
  uint64_t size = 0;
  for (int k = 0; k < 32; k++) if ((v[k >> 3] >> (8 * (k & 7))) & 0xFF) size = (uint64_t)k + 1;
  oog_finish(e, i, row, (u128)50 * size + 10, 1);
}
/* error_return_data_out_of_bound.py */
static void gadget_error_return_data_oob(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x3e));
  word_t off_w, len_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(1)), &off_w), EV_EOOG_POP0_UNSAT, row)) return;
  fr_t data_offset, length;
  W2FQ(off_w, 31, &data_offset, EV_EOOG_W0_DOMAIN);
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(1)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(2)), &len_w), EV_EOOG_POP1_UNSAT, row)) return;
  W2FQ(len_w, 31, &length, EV_EOOG_W1_DOMAIN);
  word_t v; int w;
  if (!need1(e, call_context_w(e, fr_add(rwc, fr_u64(2)), 0, call_id, ZK_CC_LastCalleeReturnDataLength, &v, &w), EV_EOOG_CC_UNSAT, row)) return;
  CHECK(EV_EOOG_CC_TYPE, !w);
  const fr_t rdl = v.lo, end = fr_add(data_offset, length);  /* < 2^249: no reduction */
  const int off_over = !fr_fits_bits(data_offset, 64), end_over = !fr_fits_bits(end, 64);
  CHECK(EV_EOOG_CMP_RANGE, fr_fits_bits(rdl, 248) && fr_fits_bits(end, 248));
  CHECK(EV_EOOG_NOT_ENOUGH, off_over || end_over || fr_cmp(rdl, end) < 0);
  error_state_tail(e, i, row, 3);
}

/* ---- BALANCE / EXTCODEHASH / EXTCODESIZE (balance.py, extcodehash.py, extcodesize.py) ------------------------------ */
static void gadget_account_access(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, uint64_t op) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  CHECK(EV_ACC_OPCODE, fr_eq_u64(opcode, op));
  word_t addr_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &addr_w), EV_ACC_POP_UNSAT, row)) return;
  fr_t address;
  W2FQ(addr_w, 20, &address, EV_ACC_ADDR_DOMAIN);
  uint32_t r;
  LK(cc_lookup(e, fr_add(rwc, fr_u64(1)), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_TXID_UNSAT);
  const fr_t tx_id = rw_cell(e, R_VAL_LO, r);
  LK(cc_lookup(e, fr_add(rwc, fr_u64(2)), call_id, ZK_CC_RwCounterEndOfReversion, &r), EV_ACC_REVEND_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_REVEND_UNSAT);
  const fr_t rev_end = rw_cell(e, R_VAL_LO, r);
  LK(cc_lookup(e, fr_add(rwc, fr_u64(3)), call_id, ZK_CC_IsPersistent, &r), EV_ACC_PERSIST_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_PERSIST_UNSAT);
  const fr_t is_persistent = rw_cell(e, R_VAL_LO, r);
  /* add_account_to_access_list(tx_id, address, reversion_info): state_write(TxAccessListAccount, tx_id, address, value = 1) */
  fr_t is_warm;
  {
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(4)), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id; key[R_ADDR] = address; key[R_VAL_LO] = one;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_VAL, &r), EV_ACC_AL_UNSAT);
    const uint32_t first = r;
    if (fr_is_zero(is_persistent)) { uint32_t r2; LK(reversion_lookup(e, fr_sub(rev_end, CUR(S_REV)), first, &r2), EV_ACC_AL_REV_UNSAT); }
    CHECK(EV_ACC_AL_PREV_TYPE, !rw_prev_is_word(e, first));
    is_warm = rw_cell(e, R_PREV_LO, first);
  }
  LK(account_lookup(e, fr_add(rwc, fr_u64(5)), 0, address, ZK_ACC_CodeHash, &r), EV_ACC_HASH_UNSAT);
  const word_t code_hash = rw_value(e, r);
  const int exists = !fr_is_zero(fr_add(code_hash.lo, code_hash.hi));  /* 1 - is_zero(lo + hi) */
  word_t expect = {fr_u64(0), fr_u64(0)};
  uint64_t n_rw = 6, d_rev = 0;
  if (op == 0x31) {  /* BALANCE */
    if (exists) {
      LK(account_lookup(e, fr_add(rwc, fr_u64(6)), 0, address, ZK_ACC_Balance, &r), EV_ACC_BAL_UNSAT);
      expect = rw_value(e, r);
      n_rw = 7;
    }
  } else if (op == 0x3f) {  /* EXTCODEHASH */
    expect = code_hash;
  } else {  /* EXTCODESIZE */
    if (exists) {
      fr_t len;
      if (!need1(e, bytecode_lookup(e, code_hash.lo, code_hash.hi, 1, fr_u64(0), 0, &len), EV_ACC_LEN_UNSAT, row)) return;
      CHECK(EV_ACC_SIZE_WORD, fr_fits_bits(len, 128));
      expect.lo = len;
    }
    d_rev = 1;
  }
  word_t w;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(n_rw)), 1, ZK_TARGET_Stack, call_id, sp, &w), EV_ACC_PUSH_UNSAT, row)) return;
  if (op == 0x3f) CHECK(EV_ACC_EQ, word_eq(expect, w)); else CHECK(EV_ACC_EQ, word_eq(expect, w));
  CHECK(EV_ACC_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  same_context_r(e, i, row, opcode, fr_u64(n_rw + 1), one, fr_u64(0), 0, fr_u64(0), fr_eq_u64(is_warm, 1) ? fr_u64(0) : fr_u64(2500), d_rev);
}
/* error_oog_account_access.py */
static void gadget_error_oog_account_access(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID);
  CHECK(EV_ACC_OPCODE, fr_eq_u64(opcode, 0x31) || fr_eq_u64(opcode, 0x3b) || fr_eq_u64(opcode, 0x3f));
  word_t addr_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, CUR(S_SP), &addr_w), EV_ACC_POP_UNSAT, row)) return;
  fr_t address;
  W2FQ(addr_w, 20, &address, EV_ACC_ADDR_DOMAIN);
  uint32_t r;
  LK(cc_lookup(e, fr_add(rwc, fr_u64(1)), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_TXID_UNSAT);
  const fr_t tx_id = rw_cell(e, R_VAL_LO, r);
  /* read_account_to_access_list: state_read(TxAccessListAccount, tx_id, address) */
  fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(2)), 0, ZK_TARGET_TxAccessListAccount);
  key[R_ID] = tx_id; key[R_ADDR] = address;
  LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR), &r), EV_ACC_AL_UNSAT);
  CHECK(EV_ACC_AL_PREV_TYPE, !rw_prev_is_word(e, r));
  const fr_t is_warm = rw_cell(e, R_PREV_LO, r);
  oog_finish(e, i, row, fr_eq_u64(is_warm, 1) ? 100 : 2600, 3);
}

/* ---- CODECOPY / RETURNDATACOPY / EXTCODECOPY (codecopy.py, returndatacopy.py, extcodecopy.py) ---------------------- */
#define CPY_POP(k, out) POP((k), (out), EV_CPY_POP0_UNSAT + 2 * (k))
/* memory_offset_and_length(moff_w, size_w) (instruction.py:1122-1127) */
#define CPY_OFFLEN(moff_w, size_w, moff, size) do { W2FQ((size_w), 5, (size), EV_CPY_SIZE_DOMAIN); *(moff) = fr_u64(0); \
  if (!fr_is_zero(*(size))) W2FQ((moff_w), 5, (moff), EV_CPY_MOFF_DOMAIN); } while (0)
#define CPY_GAS(moff, size, next_mem, gas) do { const int rc_ = copier_gas(e, i, (moff).l[0], (size).l[0], ZK_GAS_COST_COPY, (next_mem), (gas)); \
  if (rc_) { orc_fail(e->res, EV_CPY_MEMSIZE_RANGE + rc_ - 1, row); return; } } while (0)

static void gadget_codecopy(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), one = fr_u64(1);
  word_t moff_w, coff_w, size_w;
  CPY_POP(0, &moff_w); CPY_POP(1, &coff_w); CPY_POP(2, &size_w);
  fr_t moff, size, coff;
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  W2FQ(coff_w, 5, &coff, EV_CPY_OFF_DOMAIN);
  fr_t code_size;
  if (!need1(e, bytecode_lookup(e, CUR(S_HASH_LO), CUR(S_HASH_HI), 1, fr_u64(0), 0, &code_size), EV_CPY_LEN_UNSAT, row)) return;
  fr_t next_mem, gas;
  CPY_GAS(moff, size, &next_mem, &gas);
  fr_t rwc_inc = fr_u64(0), unused;
  if (!fr_is_zero(size)) {
    if (!need1(e, copy_lookup_w(e, CUR(S_HASH_LO), CUR(S_HASH_HI), ZK_COPY_Bytecode, call_id, ZK_COPY_Memory, coff, code_size, moff, size,
                                fr_add(rwc, fr_u64(3)), &rwc_inc, &unused), EV_CPY_COPY_UNSAT, row)) return;
  }
  same_context_x(e, i, row, opcode, fr_add(fr_u64(3), rwc_inc), one, fr_u64(3), 1, next_mem, gas);
}
static void gadget_returndatacopy(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), one = fr_u64(1);
  word_t moff_w, off_w, size_w;
  CPY_POP(0, &moff_w); CPY_POP(1, &off_w); CPY_POP(2, &size_w);
  word_t v; int w;
  static const uint64_t TAGS[3] = {ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataLength, ZK_CC_LastCalleeReturnDataOffset};
  fr_t cc[3];
  for (int k = 0; k < 3; k++) {
    if (!need1(e, call_context_w(e, fr_add(rwc, fr_u64(3 + k)), 0, call_id, TAGS[k], &v, &w), EV_CPY_CC0_UNSAT + 3 * k, row)) return;
    CHECK(EV_CPY_CC0_UNSAT + 3 * k + 2, !w);
    cc[k] = v.lo;
  }
  fr_t off8, size8;
  W2FQ(off_w, 8, &off8, EV_CPY_OFF_DOMAIN);
  W2FQ(size_w, 8, &size8, EV_CPY_SIZE8_DOMAIN);
  CHECK(EV_CPY_OOB_RANGE, fr_fits_bits(fr_sub(cc[1], fr_add(off8, size8)), 32));
  fr_t moff, size;
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  fr_t next_mem, gas;
  CPY_GAS(moff, size, &next_mem, &gas);
  fr_t rwc_inc = fr_u64(0), unused;
  if (!need1(e, copy_lookup(e, cc[0], ZK_COPY_Memory, call_id, ZK_COPY_Memory, cc[2], fr_add(cc[2], size), moff, size,
                            fr_add(rwc, fr_u64(6)), &rwc_inc, &unused), EV_CPY_COPY_UNSAT, row)) return;
  CHECK(EV_CPY_RWC_INC, fr_eq(rwc_inc, fr_add(size, size)));
  same_context_x(e, i, row, opcode, fr_add(fr_u64(6), rwc_inc), one, fr_u64(3), 1, next_mem, gas);
}
static void gadget_extcodecopy(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), one = fr_u64(1);
  word_t addr_w, moff_w, coff_w, size_w;
  CPY_POP(0, &addr_w);
  fr_t address;
  W2FQ(addr_w, 20, &address, EV_CPY_ADDR_DOMAIN);
  CPY_POP(1, &moff_w); CPY_POP(2, &coff_w); CPY_POP(3, &size_w);
  fr_t coff, moff, size;
  W2FQ(coff_w, 8, &coff, EV_CPY_OFF_DOMAIN);
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  uint32_t r;
  LK(cc_lookup(e, fr_add(rwc, fr_u64(4)), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_TXID_UNSAT);
  const fr_t tx_id = rw_cell(e, R_VAL_LO, r);
  LK(cc_lookup(e, fr_add(rwc, fr_u64(5)), call_id, ZK_CC_RwCounterEndOfReversion, &r), EV_ACC_REVEND_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_REVEND_UNSAT);
  const fr_t rev_end = rw_cell(e, R_VAL_LO, r);
  LK(cc_lookup(e, fr_add(rwc, fr_u64(6)), call_id, ZK_CC_IsPersistent, &r), EV_ACC_PERSIST_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_PERSIST_UNSAT);
  const fr_t is_persistent = rw_cell(e, R_VAL_LO, r);
  fr_t is_warm;
  {
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(7)), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id; key[R_ADDR] = address; key[R_VAL_LO] = one;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_VAL, &r), EV_ACC_AL_UNSAT);
    const uint32_t first = r;
    if (fr_is_zero(is_persistent)) { uint32_t r2; LK(reversion_lookup(e, fr_sub(rev_end, CUR(S_REV)), first, &r2), EV_ACC_AL_REV_UNSAT); }
    CHECK(EV_ACC_AL_PREV_TYPE, !rw_prev_is_word(e, first));
    is_warm = rw_cell(e, R_PREV_LO, first);
  }
  LK(account_lookup(e, fr_add(rwc, fr_u64(8)), 0, address, ZK_ACC_CodeHash, &r), EV_ACC_HASH_UNSAT);
  const word_t code_hash = rw_value(e, r);
  fr_t code_size = fr_u64(0);
  if (!fr_is_zero(fr_add(code_hash.lo, code_hash.hi))) {
    if (!need1(e, bytecode_lookup(e, code_hash.lo, code_hash.hi, 1, fr_u64(0), 0, &code_size), EV_ACC_LEN_UNSAT, row)) return;
  }
  fr_t next_mem, gas;
  CPY_GAS(moff, size, &next_mem, &gas);
  CHECK(EV_ACC_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  if (!fr_eq_u64(is_warm, 1)) gas = fr_add(gas, fr_u64(2500));
  fr_t rwc_inc = fr_u64(0), unused;
  if (!fr_is_zero(size)) {
    if (!need1(e, copy_lookup_w(e, code_hash.lo, code_hash.hi, ZK_COPY_Bytecode, call_id, ZK_COPY_Memory, coff, code_size, moff, size,
                                fr_add(rwc, fr_u64(9)), &rwc_inc, &unused), EV_CPY_COPY_UNSAT, row)) return;
  }
  same_context_x(e, i, row, opcode, fr_add(fr_u64(9), rwc_inc), one, fr_u64(4), 1, next_mem, gas);
}
/* error_oog_memory_copy.py: the external address goes through word_to_fq(.., N_BYTES_MEMORY_ADDRESS = 5) (:45), as written */
static void gadget_error_oog_memory_copy(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const int ext = fr_eq_u64(opcode, 0x3c);
  CHECK(EV_CPY_OPCODE, fr_eq_u64(opcode, 0x37) || fr_eq_u64(opcode, 0x39) || ext || fr_eq_u64(opcode, 0x3e));
  word_t addr_w = {fr_u64(0), fr_u64(0)}, moff_w, size_w;
  uint64_t k = 0;
  if (ext) { CPY_POP(0, &addr_w); k = 1; }
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(k)), &moff_w), EV_CPY_POP0_UNSAT + 2 * k, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k + 1)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(k + 2)), &size_w), EV_CPY_POP0_UNSAT + 2 * (k + 1), row)) return;
  uint64_t n_rw = k + 2, constant = 3;
  if (ext) {
    fr_t address;
    W2FQ(addr_w, 5, &address, EV_CPY_ADDR_DOMAIN);
    uint32_t r;
    LK(cc_lookup(e, fr_add(rwc, fr_u64(3)), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ACC_TXID_UNSAT);
    const fr_t tx_id = rw_cell(e, R_VAL_LO, r);
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(4)), 0, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id; key[R_ADDR] = address;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR), &r), EV_ACC_AL_UNSAT);
    CHECK(EV_ACC_AL_PREV_TYPE, !rw_prev_is_word(e, r));
    constant = fr_eq_u64(rw_cell(e, R_PREV_LO, r), 1) ? 100 : 2600;
    n_rw = 5;
  }
  fr_t moff, size;
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  fr_t next_mem, gas;
  CPY_GAS(moff, size, &next_mem, &gas);
  oog_finish(e, i, row, (u128)constant + gas.l[0], n_rw);
}

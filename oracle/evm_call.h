/* TEST INFRASTRUCTURE — CPU oracle, part of oracle/evm.c (included there).
 * CALL / CALLCODE / DELEGATECALL / STATICCALL: evm_circuit/execution/callop.py:12-341 with util/call_gadget.py:39-125
 * (transfer / add_balance / sub_balance instruction.py:987-1013,1111-1120, reversion_info :901-913, state_write :826-863,
 * step_state_transition_to_new_context :266-290).  The precompile branch (callop.py:158-277) reads StepState.aux_data, which
 * the 13-cell step layout of this build does not carry: it is reported as EV_CALL_PRECOMPILE (NotImplementedError).
 * Pinned by tests/golden/evm23.npz (2,241 verdicts of the reference's verify_step on 36 cases of its own test data).
 */
static int cc_rw_lookup(evm_env* e, fr_t rwc, uint64_t rw, fr_t call_id, uint64_t field, uint32_t* r) {
  fr_t key[14]; rw_key_init(key, rwc, rw, ZK_TARGET_CallContext);
  key[R_ID] = call_id; key[R_ADDR] = fr_u64(field);
  return rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR), r);
}
/* call_context_lookup(field, rw, call_id) at rw_counter + k: .value() of a non-Word row */
#define CALL_CCV(k, rw_, id_, field, out, base) do { uint32_t r_; LK(cc_rw_lookup(e, fr_add(rwc, fr_u64(k)), (rw_), (id_), (field), &r_), (base)); \
  NOT_WORD(rw_val_is_word(e, r_), (base)); *(out) = rw_cell(e, R_VAL_LO, r_); } while (0)
/* account_write_word(address, Balance, reversion_info) at rwc + k, its reversion row at rwc_rev when not persistent */
static int balance_write(evm_env* e, uint64_t row, fr_t rwc_k, fr_t address, fr_t is_persistent, fr_t rwc_rev, int id_base, uint32_t* r_out) {
  const int n = account_lookup(e, rwc_k, 1, address, ZK_ACC_Balance, r_out);
  if (n != 1) { orc_fail(e->res, n == 0 ? id_base : id_base + 1, row); return 0; }
  if (fr_is_zero(is_persistent)) {
    uint32_t r2; const int m = reversion_lookup(e, rwc_rev, *r_out, &r2);
    if (m != 1) { orc_fail(e->res, m == 0 ? id_base + 2 : id_base + 3, row); return 0; }
  }
  return 1;
}

static void gadget_callop(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  const int is_call = fr_eq_u64(opcode, 0xf1), is_callcode = fr_eq_u64(opcode, 0xf2), is_delegate = fr_eq_u64(opcode, 0xf4);
  const int is_staticcall = fr_eq_u64(opcode, 0xfa);
  {
    fr_t key[4] = {fr_u64(ZK_FIXED_ResponsibleOpcode), CUR(S_STATE), opcode, fr_u64(0)};
    CHECK(EV_CALL_RESP_OPCODE, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  }
  const fr_t callee_call_id = rwc;
  fr_t tx_id, rev_end, is_persistent, is_static, depth;
  uint32_t r;
  CALL_CCV(0, 0, call_id, ZK_CC_TxId, &tx_id, EV_CALL_TXID_UNSAT);
  CALL_CCV(1, 0, call_id, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_CALL_REVEND_UNSAT);
  CALL_CCV(2, 0, call_id, ZK_CC_IsPersistent, &is_persistent, EV_CALL_PERSIST_UNSAT);
  LK(cc_rw_lookup(e, fr_add(rwc, fr_u64(3)), 0, call_id, ZK_CC_CalleeAddress, &r), EV_CALL_SELF_UNSAT);
  const word_t ctx_caller_w = rw_value(e, r);
  fr_t ctx_caller;
  W2FQ(ctx_caller_w, 20, &ctx_caller, EV_CALL_SELF_DOMAIN);
  CALL_CCV(4, 0, call_id, ZK_CC_IsStatic, &is_static, EV_CALL_STATIC_UNSAT);
  CALL_CCV(5, 0, call_id, ZK_CC_Depth, &depth, EV_CALL_DEPTH_UNSAT);
  const word_t zero = {fr_u64(0), fr_u64(0)};
  word_t parent_caller_w = zero, parent_value = zero;
  uint64_t k = 6;
  if (is_delegate) {
    LK(cc_rw_lookup(e, fr_add(rwc, fr_u64(6)), 0, call_id, ZK_CC_CallerAddress, &r), EV_CALL_PCALLER_UNSAT);
    parent_caller_w = rw_value(e, r);
    LK(cc_rw_lookup(e, fr_add(rwc, fr_u64(7)), 0, call_id, ZK_CC_Value, &r), EV_CALL_PVALUE_UNSAT);
    parent_value = rw_value(e, r);
    k = 8;
  }
  /* CallGadget(instruction, FQ(1), ..), call_gadget.py:39-108 */
  CHECK(EV_CALL_OPCODE, is_call + is_callcode + is_delegate + is_staticcall == 1);
  const int has_value_pop = is_call || is_callcode;
  word_t w[8];
  uint64_t spo = 0;
  for (int f = 0; f < 7; f++) {
    if (f == 2 && !has_value_pop) { w[2] = zero; continue; }
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(spo)), &w[f]), EV_CALL_POP0_UNSAT + 2 * f, row)) return;
    k++; spo++;
  }
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(spo - 1)), &w[7]), EV_CALL_PUSH_UNSAT, row)) return;
  k++;
  CHECK(EV_CALL_RESULT_WORD, fr_is_zero(w[7].hi));
  const fr_t is_success = w[7].lo;
  CHECK(EV_CALL_RESULT_BOOL, fr_eq_u64(is_success, 0) || fr_eq_u64(is_success, 1));
  const int success = fr_eq_u64(is_success, 1);
  fr_t gas, callee;
  W2FQ(w[0], 8, &gas, EV_CALL_GAS_DOMAIN); /* after this, is_u64_gas == 1 */
  const int has_value = has_value_pop && !fr_is_zero(fr_add(w[2].lo, w[2].hi));
  W2FQ(w[1], 20, &callee, EV_CALL_CALLEE_DOMAIN);
  fr_t cd_off = fr_u64(0), cd_len, rd_off = fr_u64(0), rd_len;
  W2FQ(w[4], 5, &cd_len, EV_CALL_CDLEN_DOMAIN);
  if (!fr_is_zero(cd_len)) W2FQ(w[3], 5, &cd_off, EV_CALL_CDOFF_DOMAIN);
  W2FQ(w[6], 5, &rd_len, EV_CALL_RDLEN_DOMAIN);
  if (!fr_is_zero(rd_len)) W2FQ(w[5], 5, &rd_off, EV_CALL_RDOFF_DOMAIN);
  const uint64_t cd_words = (cd_off.l[0] + cd_len.l[0] + 31) / 32, rd_words = (rd_off.l[0] + rd_len.l[0] + 31) / 32;
  CHECK(EV_CALL_CD_MEMSIZE_RANGE, !(cd_words >> 32));
  const fr_t cur_mem = CUR(S_MEM);
  CHECK(EV_CALL_MEM_MAX, fr_fits_bits(cur_mem, 32));
  uint64_t next_mem = cur_mem.l[0] < cd_words ? cd_words : cur_mem.l[0];
  CHECK(EV_CALL_RD_MEMSIZE_RANGE, !(rd_words >> 32));
  next_mem = next_mem < rd_words ? rd_words : next_mem;
  const uint64_t expansion = memory_gas_cost(next_mem) - memory_gas_cost(cur_mem.l[0]);
  LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 0, callee, ZK_ACC_CodeHash, &r), EV_CALL_HASH_UNSAT);
  k++;
  const word_t callee_hash = rw_value(e, r);
  const word_t empty = {fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
  /* is_equal_word / is_zero_word: the FIELD SUM of the halves (differences) is zero, instruction.py:411-414, 489-490 */
  const int is_empty_hash = fr_is_zero(fr_add(fr_sub(callee_hash.lo, empty.lo), fr_sub(callee_hash.hi, empty.hi)));
  const int not_exists = fr_is_zero(fr_add(callee_hash.lo, callee_hash.hi));
  /* callop.py:41-55 */
  const fr_t callee_address = (is_callcode || is_delegate) ? ctx_caller : callee; /* < 2^160: address_to_word holds */
  const word_t callee_address_w = {fr_u128(callee_address.l[0], callee_address.l[1]), fr_u64(callee_address.l[2])};
  const word_t caller_address_w = is_delegate ? parent_caller_w : ctx_caller_w;
  CHECK(EV_CALL_CALLER_WORD, !is_delegate || word_in_domain(parent_caller_w)); /* select_word builds a Word: halves < 2^128 */
  fr_t caller_address;
  W2FQ(caller_address_w, 20, &caller_address, EV_CALL_CALLER_DOMAIN);
  /* add_account_to_access_list(tx_id, call.callee_address, reversion_info) */
  fr_t rev_count = CUR(S_REV);
  {
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(k)), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id; key[R_ADDR] = callee; key[R_VAL_LO] = one;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_VAL, &r), EV_CALL_AL_UNSAT);
    k++;
    const uint32_t first = r;
    if (fr_is_zero(is_persistent)) {
      uint32_t r2; LK(reversion_lookup(e, fr_sub(rev_end, rev_count), first, &r2), EV_CALL_AL_REV_UNSAT);
      rev_count = fr_add(rev_count, one);
    }
    CHECK(EV_CALL_AL_PREV_TYPE, !rw_prev_is_word(e, first));
    r = first;
  }
  const fr_t is_warm = rw_cell(e, R_PREV_LO, r);
  CHECK(EV_CALL_VALUE_STATIC, !has_value || fr_is_zero(is_static));
  /* callee's reversion info */
  fr_t callee_rev_end, callee_persistent;
  CALL_CCV(k, 0, callee_call_id, ZK_CC_RwCounterEndOfReversion, &callee_rev_end, EV_CALL_CREVEND_UNSAT); k++;
  CALL_CCV(k, 0, callee_call_id, ZK_CC_IsPersistent, &callee_persistent, EV_CALL_CPERSIST_UNSAT); k++;
  CHECK(EV_CALL_CPERSIST_EQ, fr_eq(callee_persistent, fr_mul(is_persistent, is_success)));
  if (success && fr_is_zero(is_persistent)) {
    CHECK(EV_CALL_CREVEND_EQ, fr_eq(callee_rev_end, fr_sub(rev_end, rev_count)));
    rev_count = fr_add(rev_count, one);
  }
  int insufficient = 0;
  if (has_value_pop) {
    LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 0, caller_address, ZK_ACC_Balance, &r), EV_CALL_BAL_UNSAT);
    k++;
    const word_t bal = rw_value(e, r);
    CHECK(EV_CALL_BAL_CMP_RANGE, word_in_domain(bal) && word_in_domain(w[2])); /* compare_word: 16-byte range asserts */
    insufficient = fr_cmp(bal.hi, w[2].hi) < 0 || (fr_eq(bal.hi, w[2].hi) && fr_cmp(bal.lo, w[2].lo) < 0);
  }
  CHECK(EV_CALL_DEPTH_RANGE, fr_fits_bits(depth, 16)); /* compare(depth, 1025, 2) */
  const int precheck_ok = depth.l[0] < 1025 && !insufficient;
  if (!precheck_ok) CHECK(EV_CALL_PRECHECK_SUCCESS, !success && fr_is_zero(is_success));
  if (is_call && precheck_ok) { /* transfer(caller, callee, value, callee_reversion_info) */
    fr_t crev = fr_u64(0);
    if (!balance_write(e, row, fr_add(rwc, fr_u64(k)), caller_address, callee_persistent, fr_sub(callee_rev_end, crev), EV_CALL_SEND_UNSAT, &r)) return;
    k++; crev = one;
    { word_t ws[2] = {rw_value(e, r), w[2]}; fr_t carry; const word_t sum = add_words_n(ws, 2, &carry);
      CHECK(EV_CALL_SEND_EQ, word_eq(rw_prev(e, r), sum)); CHECK(EV_CALL_SEND_CARRY, fr_is_zero(carry)); }
    if (!balance_write(e, row, fr_add(rwc, fr_u64(k)), callee_address, callee_persistent, fr_sub(callee_rev_end, crev), EV_CALL_RECV_UNSAT, &r)) return;
    k++;
    { word_t ws[2] = {rw_prev(e, r), w[2]}; fr_t carry; const word_t sum = add_words_n(ws, 2, &carry);
      CHECK(EV_CALL_RECV_EQ, word_eq(rw_value(e, r), sum)); CHECK(EV_CALL_RECV_CARRY, fr_is_zero(carry)); }
  }
  if (is_callcode && success) CHECK(EV_CALL_CALLCODE_BALANCE, !insufficient);
  /* gas: call.gas_cost(instruction, is_warm_access, is_call), EIP-150 */
  CHECK(EV_CALL_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  const uint64_t gas_cost = (fr_eq_u64(is_warm, 1) ? 100 : 2600) + (has_value ? 9000 + ((is_call && success && not_exists) ? 25000 : 0) : 0) + expansion;
  const fr_t gas_available = fr_sub(CUR(S_GAS), fr_u64(gas_cost));
  /* constant_divmod(gas_available, 64, 8): the quotient of the integer below p must fit 8 bytes */
  fr_t one_64th = {{(gas_available.l[0] >> 6) | (gas_available.l[1] << 58), (gas_available.l[1] >> 6) | (gas_available.l[2] << 58),
                    (gas_available.l[2] >> 6) | (gas_available.l[3] << 58), gas_available.l[3] >> 6}};
  CHECK(EV_CALL_GAS_64TH_RANGE, fr_fits_bits(one_64th, 64));
  const fr_t all_but = fr_sub(gas_available, one_64th);
  CHECK(EV_CALL_GAS_MIN_RANGE, fr_fits_bits(all_but, 64)); /* min(all_but_one_64th_gas, call.gas, 8) */
  fr_t callee_gas_left = all_but.l[0] < gas.l[0] ? all_but : gas;
  const int is_precompile = fr_fits_bits(callee, 64) && callee.l[0] >= 1 && callee.l[0] <= 9;
  const fr_t ns = NXT(S_STATE);
  const int next_is_precompile = fr_fits_bits(ns, 16) && ns.l[0] >= ZK_ES_ECRECOVER && ns.l[0] <= ZK_ES_ECRECOVER + 8;
  CHECK(EV_CALL_PRECOMPILE_STATE, is_precompile == next_is_precompile);
  const uint64_t sp_delta = 5 + (uint64_t)is_call + (uint64_t)is_callcode;
  const int no_callee_code = is_empty_hash + not_exists;
  if (!precheck_ok || (no_callee_code == 1 && !is_precompile)) {
    static const uint64_t TAGS[3] = {ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataOffset, ZK_CC_LastCalleeReturnDataLength};
    for (int t = 0; t < 3; t++) {
      fr_t v;
      CALL_CCV(k, 1, call_id, TAGS[t], &v, EV_CALL_LAST0_UNSAT + 4 * t); k++;
      CHECK(EV_CALL_LAST0_UNSAT + 4 * t + 3, fr_is_zero(v));
    }
    CHECK(EV_CALL_SAME_RWC, fr_eq(NXT(S_RWC), fr_add(rwc, fr_u64(k))));
    CHECK(EV_CALL_SAME_PC, fr_eq(NXT(S_PC), fr_add(CUR(S_PC), one)));
    CHECK(EV_CALL_SAME_SP, fr_eq(NXT(S_SP), fr_add(sp, fr_u64(sp_delta))));
    CHECK(EV_CALL_SAME_GAS, fr_eq(NXT(S_GAS), fr_sub(fr_add(CUR(S_GAS), fr_u64(has_value ? 2300 : 0)), fr_u64(gas_cost))));
    CHECK(EV_CALL_SAME_MEM, fr_eq_u64(NXT(S_MEM), next_mem));
    CHECK(EV_CALL_SAME_REV, fr_eq(NXT(S_REV), fr_add(CUR(S_REV), fr_u64(3))));
    CHECK(EV_CALL_SAME_CALL_ID, fr_eq(NXT(S_CALL_ID), call_id));
    CHECK(EV_CALL_SAME_IS_ROOT, fr_eq(NXT(S_IS_ROOT), CUR(S_IS_ROOT)));
    CHECK(EV_CALL_SAME_IS_CREATE, fr_eq(NXT(S_IS_CREATE), CUR(S_IS_CREATE)));
    CHECK(EV_CALL_SAME_CODE_HASH, fr_eq(NXT(S_HASH_LO), CUR(S_HASH_LO)) && fr_eq(NXT(S_HASH_HI), CUR(S_HASH_HI)));
    return;
  }
  if (is_precompile) { orc_fail(e->res, EV_CALL_PRECOMPILE, row); return; } /* needs StepState.aux_data */
  { /* save the caller's state: 5 call-context writes on the current call */
    const fr_t want[5] = {fr_add(CUR(S_PC), one), fr_add(sp, fr_u64(sp_delta)), fr_sub(fr_sub(CUR(S_GAS), fr_u64(gas_cost)), callee_gas_left),
                          fr_u64(next_mem), fr_add(CUR(S_REV), one)};
    static const uint64_t TAGS[5] = {ZK_CC_ProgramCounter, ZK_CC_StackPointer, ZK_CC_GasLeft, ZK_CC_MemorySize, ZK_CC_ReversibleWriteCounter};
    for (int t = 0; t < 5; t++) {
      fr_t v;
      CALL_CCV(k, 1, call_id, TAGS[t], &v, EV_CALL_SAVE0_UNSAT + 4 * t); k++;
      CHECK(EV_CALL_SAVE0_UNSAT + 4 * t + 3, fr_eq(v, want[t]));
    }
  }
  { /* the callee's context: 18 call-context reads compared as words (lo, hi) */
    const word_t value_w = is_delegate ? parent_value : w[2];
    CHECK(EV_CALL_VALUE_WORD, word_in_domain(value_w)); /* select_word builds a Word */
    const word_t want[18] = {{call_id, fr_u64(0)}, {tx_id, fr_u64(0)}, {fr_add(depth, one), fr_u64(0)}, caller_address_w, callee_address_w,
                             {cd_off, fr_u64(0)}, {cd_len, fr_u64(0)}, {rd_off, fr_u64(0)}, {rd_len, fr_u64(0)}, value_w, {is_success, fr_u64(0)},
                             {is_static, fr_u64(0)}, zero, zero, zero, zero, zero, callee_hash};
    static const uint64_t TAGS[18] = {ZK_CC_CallerId, ZK_CC_TxId, ZK_CC_Depth, ZK_CC_CallerAddress, ZK_CC_CalleeAddress, ZK_CC_CallDataOffset,
                                      ZK_CC_CallDataLength, ZK_CC_ReturnDataOffset, ZK_CC_ReturnDataLength, ZK_CC_Value, ZK_CC_IsSuccess,
                                      ZK_CC_IsStatic, ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataOffset, ZK_CC_LastCalleeReturnDataLength,
                                      ZK_CC_IsRoot, ZK_CC_IsCreate, ZK_CC_CodeHash};
    for (int t = 0; t < 18; t++) {
      LK(cc_rw_lookup(e, fr_add(rwc, fr_u64(k)), 0, callee_call_id, TAGS[t], &r), EV_CALL_CTX0_UNSAT + 3 * t); k++;
      CHECK(EV_CALL_CTX0_UNSAT + 3 * t + 2, word_eq(rw_value(e, r), want[t]));
    }
  }
  callee_gas_left = fr_add(callee_gas_left, fr_u64(has_value ? 2300 : 0));
  CHECK(EV_CALL_NC_RWC, fr_eq(NXT(S_RWC), fr_add(rwc, fr_u64(k))));
  CHECK(EV_CALL_NC_CALL_ID, fr_eq(NXT(S_CALL_ID), callee_call_id));
  CHECK(EV_CALL_NC_IS_ROOT, fr_is_zero(NXT(S_IS_ROOT)));
  CHECK(EV_CALL_NC_IS_CREATE, fr_is_zero(NXT(S_IS_CREATE)));
  CHECK(EV_CALL_NC_CODE_HASH, fr_eq(NXT(S_HASH_LO), callee_hash.lo) && fr_eq(NXT(S_HASH_HI), callee_hash.hi));
  CHECK(EV_CALL_NC_GAS, fr_eq(NXT(S_GAS), callee_gas_left));
  CHECK(EV_CALL_NC_REV, fr_eq_u64(NXT(S_REV), 2));
  CHECK(EV_CALL_NC_LOG, fr_eq(NXT(S_LOG), CUR(S_LOG)));
  CHECK(EV_CALL_NC_PC, fr_is_zero(NXT(S_PC)));
  CHECK(EV_CALL_NC_SP, fr_eq_u64(NXT(S_SP), 1024));
  CHECK(EV_CALL_NC_MEM, fr_is_zero(NXT(S_MEM)));
}

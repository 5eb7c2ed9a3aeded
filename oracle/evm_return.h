/* TEST INFRASTRUCTURE — CPU oracle, part of oracle/evm.c (included there).
 * RETURN / REVERT: evm_circuit/execution/return_revert.py:10-135 (copy_lookup table.py:760-787, account_write_word
 * instruction.py:973-985, step_state_transition_to_restored_context :292-363).  Reproduced as written, including:
 *   - `if instruction.curr.is_create and is_success` / `if not is_return` test the truthiness of FQ OBJECTS (always
 *     true): the deployment branch runs for REVERT too and the reversible write counter is never added to the delta;
 *   - the deployment branch's two lookups (CalleeAddress, the code-hash write) are not counted in rwc_delta, so the
 *     next step's rw_counter is two less than the lookups consumed;
 *   - the copy lookup towards the caller is unconditional (a zero-length return to a caller has no copy event to find).
 * Pinned by tests/golden/evm21.npz (1,301 verdicts of the reference's verify_step).
 */
static int copy_lookup_dw(evm_env* e, fr_t src_id, uint64_t src_tag, word_t dst_id, uint64_t dst_tag, fr_t src_addr, fr_t src_end,
                          fr_t dst_addr, fr_t length, fr_t rwc, fr_t* rwc_inc) {
  fr_t key[11] = {src_id, fr_u64(0), fr_u64(src_tag), dst_id.lo, dst_id.hi, fr_u64(dst_tag), src_addr, src_end, dst_addr, length, rwc};
  uint32_t r; const int n = orc_lookup(&e->copy_ix, key, &r);
  if (n == 1) *rwc_inc = fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 13, r));
  return n;
}
static void gadget_return_revert(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  const int is_return = fr_eq_u64(opcode, 0xf3);
  fr_t is_success;
  ST_CC(0, ZK_CC_IsSuccess, &is_success, EV_RET_SUCCESS_UNSAT);
  CHECK(EV_RET_SUCCESS_EQ, fr_eq_u64(is_success, (uint64_t)is_return));
  word_t off_w, len_w;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, sp, &off_w), EV_RET_POP0_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(2)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &len_w), EV_RET_POP1_UNSAT, row)) return;
  fr_t offset, length;
  W2FQ(off_w, 5, &offset, EV_RET_OFF_DOMAIN);
  W2FQ(len_w, 5, &length, EV_RET_LEN_DOMAIN);
  const fr_t ret_end = fr_add(offset, length);
  fr_t look = fr_u64(3), delta = fr_u64(3); /* rw_counter_offset / rwc_delta */
  fr_t gas_left = CUR(S_GAS);
  const int is_create = !fr_is_zero(CUR(S_IS_CREATE)), is_root = !fr_is_zero(CUR(S_IS_ROOT));
  uint32_t r;
  if (is_create) { /* A. the memory chunk becomes the deployed code */
    LK(cc_lookup(e, fr_add(rwc, fr_u64(3)), call_id, ZK_CC_CalleeAddress, &r), EV_RET_CALLEE_UNSAT);
    fr_t callee;
    W2FQ(rw_value(e, r), 20, &callee, EV_RET_CALLEE_DOMAIN);
    LK(account_lookup(e, fr_add(rwc, fr_u64(4)), 1, callee, ZK_ACC_CodeHash, &r), EV_RET_HASH_WRITE_UNSAT);
    const word_t code_hash = rw_value(e, r), code_hash_prev = rw_prev(e, r);
    const word_t empty = {fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
    CHECK(EV_RET_HASH_PREV, word_eq(code_hash_prev, empty));
    CHECK(EV_RET_HASH_CUR, fr_eq(code_hash.lo, CUR(S_HASH_LO)) && fr_eq(code_hash.hi, CUR(S_HASH_HI)));
    {
      fr_t key[4] = {fr_u64(ZK_FIXED_Range24_576), length, fr_u64(0), fr_u64(0)};
      CHECK(EV_RET_MAX_CODE_SIZE, orc_lookup(&e->fixed_ix, key, 0) >= 1);
    }
    gas_left = fr_sub(gas_left, fr_mul(length, fr_u64(200)));
    look = fr_u64(5);
    if (!fr_is_zero(length)) {
      fr_t inc;
      if (!need1(e, copy_lookup_dw(e, call_id, ZK_COPY_Memory, code_hash, ZK_COPY_Bytecode, offset, ret_end, fr_u64(0), length,
                                   fr_add(rwc, fr_u64(5)), &inc), EV_RET_COPY_CODE_UNSAT, row)) return;
      CHECK(EV_RET_COPY_CODE_INC, fr_eq(inc, length));
      look = fr_add(look, inc);
      delta = fr_add(delta, length);
      fr_t code_size;
      if (!need1(e, bytecode_lookup(e, code_hash.lo, code_hash.hi, 1, fr_u64(0), 0, &code_size), EV_RET_CODE_LEN_UNSAT, row)) return;
      CHECK(EV_RET_CODE_LEN_EQ, fr_eq(code_size, length));
    }
  }
  if (!is_root && !is_create) { /* D. the memory chunk is copied to the caller's memory */
    fr_t caller_off, caller_len;
    LK(cc_lookup(e, fr_add(rwc, look), call_id, ZK_CC_ReturnDataOffset, &r), EV_RET_RDO_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_RET_RDO_UNSAT);
    caller_off = rw_cell(e, R_VAL_LO, r);
    LK(cc_lookup(e, fr_add(rwc, fr_add(look, one)), call_id, ZK_CC_ReturnDataLength, &r), EV_RET_RDL_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_RET_RDL_UNSAT);
    caller_len = rw_cell(e, R_VAL_LO, r);
    CHECK(EV_RET_MIN_RANGE, fr_fits_bits(caller_len, 40)); /* min(return_length, caller_return_length, 5) */
    const fr_t copy_len = length.l[0] < caller_len.l[0] ? length : caller_len;
    fr_t inc, unused;
    if (!need1(e, copy_lookup(e, call_id, ZK_COPY_Memory, NXT(S_CALL_ID), ZK_COPY_Memory, offset, ret_end, caller_off, copy_len,
                              fr_add(rwc, fr_add(look, fr_u64(2))), &inc, &unused), EV_RET_COPY_UNSAT, row)) return;
    CHECK(EV_RET_COPY_INC, fr_eq(inc, fr_add(copy_len, copy_len)));
    look = fr_add(fr_add(look, fr_u64(2)), inc);
    delta = fr_add(fr_add(delta, fr_u64(2)), fr_add(copy_len, copy_len));
  }
  CHECK(EV_RET_ROOT_ENDTX, fr_eq_u64(CUR(S_IS_ROOT), fr_eq_u64(NXT(S_STATE), ZK_ES_EndTx) ? 1 : 0));
  /* memory_expansion_dynamic_length(return_offset, return_length) */
  const uint64_t words = (offset.l[0] + length.l[0] + 31) / 32; /* both below 2^40 */
  CHECK(EV_RET_MEMSIZE_RANGE, !(words >> 32));
  const fr_t cur_mem = CUR(S_MEM);
  CHECK(EV_RET_MEM_MAX, fr_fits_bits(cur_mem, 32));
  const uint64_t nxt_mem = cur_mem.l[0] < words ? words : cur_mem.l[0];
  const uint64_t expansion = memory_gas_cost(nxt_mem) - memory_gas_cost(cur_mem.l[0]);
  if (is_root) { /* B2 */
    fr_t is_persistent;
    LK(cc_lookup(e, fr_add(rwc, look), call_id, ZK_CC_IsPersistent, &r), EV_RET_PERSIST_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_RET_PERSIST_UNSAT);
    is_persistent = rw_cell(e, R_VAL_LO, r);
    CHECK(EV_RET_PERSIST_EQ, fr_eq_u64(is_persistent, (uint64_t)is_return));
    CHECK(EV_RET_RWC, fr_eq(NXT(S_RWC), fr_add(rwc, fr_add(delta, one))));
    CHECK(EV_RET_GAS, fr_eq(NXT(S_GAS), gas_left));
    CHECK(EV_RET_CALL_ID, fr_eq(NXT(S_CALL_ID), call_id));
  } else { /* C */
    restore_context_f(e, i, row, look, delta, offset, length, fr_sub(gas_left, fr_u64(expansion)), 1);
  }
}

/* ---- ErrorOutOfGasCall: error_oog_call.py:11-42 with util/call_gadget.py:39-125 (CallGadget, IS_SUCCESS_CALL = 0) ----
 * Pinned by tests/golden/evm22.npz (890 verdicts of the reference's verify_step). */
static void gadget_error_oog_call(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const uint64_t op = fr_fits_bits(opcode, 8) ? opcode.l[0] : 0x100;
  CHECK(EV_EOC_OPCODE, op == 0xf1 || op == 0xf2 || op == 0xf4 || op == 0xfa);
  const int has_value_pop = op == 0xf1 || op == 0xf2;
  fr_t tx_id;
  ST_CC(0, ZK_CC_TxId, &tx_id, EV_EOC_TXID_UNSAT);
  /* CallGadget: the stack words, call_gadget.py:53-66 */
  word_t w[8]; /* gas, callee, value, cd_offset, cd_length, rd_offset, rd_length, result */
  const word_t zero = {fr_u64(0), fr_u64(0)};
  uint64_t k = 1, spo = 0;
  for (int f = 0; f < 7; f++) {
    if (f == 2 && !has_value_pop) { w[2] = zero; continue; }
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(spo)), &w[f]), EV_EOC_POP0_UNSAT + 2 * f, row)) return;
    k++; spo++;
  }
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(spo - 1)), &w[7]), EV_EOC_PUSH_UNSAT, row)) return;
  k++;
  CHECK(EV_EOC_RESULT_WORD, fr_is_zero(w[7].hi));                       /* Word.from_lo(is_success) == result */
  CHECK(EV_EOC_RESULT_BOOL, fr_eq_u64(w[7].lo, 0) || fr_eq_u64(w[7].lo, 1));
  CHECK(EV_EOC_RESULT_ZERO, fr_is_zero(w[7].lo));                       /* IS_SUCCESS_CALL == 0 */
  fr_t gas, callee;
  W2FQ(w[0], 8, &gas, EV_EOC_GAS_DOMAIN);
  (void)gas;
  const int has_value = has_value_pop && !fr_is_zero(fr_add(w[2].lo, w[2].hi));
  W2FQ(w[1], 20, &callee, EV_EOC_CALLEE_DOMAIN);
  /* memory_offset_and_length x 2: the length first, the offset only when the length is not zero */
  fr_t cd_off = fr_u64(0), cd_len, rd_off = fr_u64(0), rd_len;
  W2FQ(w[4], 5, &cd_len, EV_EOC_CDLEN_DOMAIN);
  if (!fr_is_zero(cd_len)) W2FQ(w[3], 5, &cd_off, EV_EOC_CDOFF_DOMAIN);
  W2FQ(w[6], 5, &rd_len, EV_EOC_RDLEN_DOMAIN);
  if (!fr_is_zero(rd_len)) W2FQ(w[5], 5, &rd_off, EV_EOC_RDOFF_DOMAIN);
  /* memory_expansion_dynamic_length(cd_offset, cd_length, rd_offset, rd_length), instruction.py:1157-1181 */
  const uint64_t cd_words = (cd_off.l[0] + cd_len.l[0] + 31) / 32, rd_words = (rd_off.l[0] + rd_len.l[0] + 31) / 32;
  CHECK(EV_EOC_CD_MEMSIZE_RANGE, !(cd_words >> 32));
  const fr_t cur_mem = CUR(S_MEM);
  CHECK(EV_EOC_MEM_MAX, fr_fits_bits(cur_mem, 32));
  uint64_t nxt = cur_mem.l[0] < cd_words ? cd_words : cur_mem.l[0];
  CHECK(EV_EOC_RD_MEMSIZE_RANGE, !(rd_words >> 32));
  nxt = nxt < rd_words ? rd_words : nxt;
  const uint64_t expansion = memory_gas_cost(nxt) - memory_gas_cost(cur_mem.l[0]);
  uint32_t r;
  LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 0, callee, ZK_ACC_CodeHash, &r), EV_EOC_HASH_UNSAT);
  k++;
  { /* read_account_to_access_list: state_read(TxAccessListAccount, tx_id, callee) */
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(k)), 0, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id; key[R_ADDR] = callee;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR), &r), EV_EOC_AL_UNSAT);
    k++;
  }
  CHECK(EV_EOC_AL_PREV_TYPE, !rw_prev_is_word(e, r));
  const fr_t is_warm = rw_cell(e, R_PREV_LO, r);
  CHECK(EV_EOC_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  /* gas_cost(): is_success == 0 here, so the new-account term vanishes */
  const uint64_t cost = (fr_eq_u64(is_warm, 1) ? 100 : 2600) + (has_value ? 9000 : 0) + expansion;
  const fr_t gas_left = CUR(S_GAS);
  CHECK(EV_EOC_CMP_RANGE, fr_fits_bits(gas_left, 64));
  CHECK(EV_EOC_NOT_ENOUGH, gas_left.l[0] < cost);
  error_state_tail(e, i, row, k);
}

// Gate program of RETURN / REVERT (group KG_TX), part of evm.cu (included there).
//   return_revert   evm_circuit/execution/return_revert.py:10-135 (copy_lookup table.py:760-787, account_write_word
//                   instruction.py:973-985, step_state_transition_to_restored_context :292-363)
// Reproduced as written, including: `if instruction.curr.is_create and is_success` / `if not is_return` test the
// truthiness of FQ OBJECTS (always true) — the deployment branch runs for REVERT too and the reversible write counter is
// never added to the delta; the deployment branch's two lookups are not counted in rwc_delta (the next step starts two
// rw counters early); the copy lookup towards the caller is unconditional.
#pragma once
namespace zk {

// copy_lookup whose destination id is a Word (the code hash of the deployed contract)
ZK_HD_NOINLINE int copy_lookup_dw(const StepCtx& s, const Fr& src_id, u64 src_tag, const Word2& dst_id, u64 dst_tag, const Fr& src_addr,
                                  const Fr& src_end, const Fr& dst_addr, const Fr& length, const Fr& rwc, Fr* rwc_inc) {
  Fr key[11] = {src_id, fr_u64(0), fr_u64(src_tag), dst_id.lo, dst_id.hi, fr_u64(dst_tag), src_addr, src_end, dst_addr, length, rwc};
  u32 r = 0;
  const int n = lookup_sync<11>(s.t.copy, key, &r, s.mask, true);
  if (n == 1) *rwc_inc = table_cell(s.t.copy.tab, 13, r);
  return n;
}
// call_context_lookup(field) at rw_counter + off (a field element): the value cell of a non-Word row
#define RET_CC(off, field, out, base)                                                       \
  do {                                                                                      \
    u32 r_ = 0;                                                                             \
    TX_LK(cc_lookup_m(s, fr_add(rwc, (off)), call_id, (field), &r_), (base));               \
    TX_NOT_WORD(rw_flag(s, r_, 0), (base));                                                 \
    *(out) = rw_cell(s, R_VAL_LO, r_);                                                      \
  } while (0)

ZK_HD_NOINLINE void gadget_return_revert(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  const bool is_return = fr_eq_u64(opcode, 0xf3);
  Fr is_success;
  ST_CC(0, ZK_CC_IsSuccess, &is_success, EV_RET_SUCCESS_UNSAT);
  EV_CHECK(EV_RET_SUCCESS_EQ, fr_eq_u64(is_success, is_return ? 1 : 0));
  Word2 off_w{fr_u64(0), fr_u64(0)}, len_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 1, 0, sp, &off_w), EV_RET_POP0_UNSAT)) return;
  if (!need1(s, true, stack_at(s, true, 2, 0, fr_add_u64(sp, 1), &len_w), EV_RET_POP1_UNSAT)) return;
  Fr offset = fr_u64(0), length = fr_u64(0);
  EOOG_W2FQ(off_w, 5, &offset, EV_RET_OFF_DOMAIN);
  EOOG_W2FQ(len_w, 5, &length, EV_RET_LEN_DOMAIN);
  const Fr ret_end = fr_add(offset, length);
  Fr look = fr_u64(3), delta = fr_u64(3);  // rw_counter_offset / rwc_delta
  Fr gas_left = s.cur(S_GAS);
  const bool is_create = !fr_is_zero(s.cur(S_IS_CREATE)), is_root = !fr_is_zero(s.cur(S_IS_ROOT));
  u32 r = 0;
  if (is_create) {  // A. the memory chunk becomes the deployed code
    TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 3), call_id, ZK_CC_CalleeAddress, &r), EV_RET_CALLEE_UNSAT);
    Fr callee = fr_u64(0);
    EOOG_W2FQ(rw_word(s, R_VAL_LO, r), 20, &callee, EV_RET_CALLEE_DOMAIN);
    TX_LK(account_lookup_m(s, fr_add_u64(rwc, 4), 1, callee, ZK_ACC_CodeHash, &r), EV_RET_HASH_WRITE_UNSAT);
    const Word2 code_hash = rw_word(s, R_VAL_LO, r), code_hash_prev = rw_word(s, R_PREV_LO, r);
    const Word2 empty{fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
    EV_CHECK(EV_RET_HASH_PREV, word_eq(code_hash_prev, empty));
    EV_CHECK(EV_RET_HASH_CUR, fr_eq(code_hash.lo, s.cur(S_HASH_LO)) && fr_eq(code_hash.hi, s.cur(S_HASH_HI)));
    {
      Fr key[4] = {fr_u64(ZK_FIXED_Range24_576), length, fr_u64(0), fr_u64(0)};
      u32 rr = 0;
      EV_CHECK(EV_RET_MAX_CODE_SIZE, lookup<4>(s.t.fixed, key, &rr) >= 1);
    }
    gas_left = fr_sub(gas_left, fr_mul(length, fr_u64(200)));
    look = fr_u64(5);
    if (!fr_is_zero(length)) {
      Fr inc = fr_u64(0);
      if (!need1(s, true, copy_lookup_dw(s, call_id, ZK_COPY_Memory, code_hash, ZK_COPY_Bytecode, offset, ret_end, fr_u64(0), length,
                                         fr_add_u64(rwc, 5), &inc), EV_RET_COPY_CODE_UNSAT)) return;
      EV_CHECK(EV_RET_COPY_CODE_INC, fr_eq(inc, length));
      look = fr_add(look, inc);
      delta = fr_add(delta, length);
      Fr code_size = fr_u64(0);
      if (!need1(s, true, bytecode_lookup_ni(s, true, code_hash.lo, code_hash.hi, 1, fr_u64(0), 0, &code_size), EV_RET_CODE_LEN_UNSAT)) return;
      EV_CHECK(EV_RET_CODE_LEN_EQ, fr_eq(code_size, length));
    }
  }
  if (!is_root && !is_create) {  // D. the memory chunk is copied to the caller's memory
    Fr caller_off, caller_len;
    RET_CC(look, ZK_CC_ReturnDataOffset, &caller_off, EV_RET_RDO_UNSAT);
    RET_CC(fr_add_u64(look, 1), ZK_CC_ReturnDataLength, &caller_len, EV_RET_RDL_UNSAT);
    EV_CHECK(EV_RET_MIN_RANGE, fr_fits64(caller_len) && (caller_len.l[0] >> 40) == 0);  // min(.., .., 5)
    const Fr copy_len = length.l[0] < caller_len.l[0] ? length : caller_len;
    Fr inc = fr_u64(0), unused = fr_u64(0);
    if (!need1(s, true, copy_lookup(s, true, call_id, ZK_COPY_Memory, s.nxt(S_CALL_ID), ZK_COPY_Memory, offset, ret_end, caller_off, copy_len,
                                    fr_add(rwc, fr_add_u64(look, 2)), &inc, &unused), EV_RET_COPY_UNSAT)) return;
    const Fr twice = fr_add(copy_len, copy_len);
    EV_CHECK(EV_RET_COPY_INC, fr_eq(inc, twice));
    look = fr_add(fr_add_u64(look, 2), inc);
    delta = fr_add(fr_add_u64(delta, 2), twice);
  }
  EV_CHECK(EV_RET_ROOT_ENDTX, fr_eq_u64(s.cur(S_IS_ROOT), fr_eq_u64(s.nxt(S_STATE), ZK_ES_EndTx) ? 1 : 0));
  // memory_expansion_dynamic_length(return_offset, return_length); both below 2^40
  const u64 words = (offset.l[0] + length.l[0] + 31) / 32;
  EV_CHECK(EV_RET_MEMSIZE_RANGE, (words >> 32) == 0);
  u64 expansion = 0;
  EV_CHECK(EV_RET_MEM_MAX, mem_expansion_gas(s, words, &expansion) == 0);
  if (is_root) {  // B2
    Fr is_persistent;
    RET_CC(look, ZK_CC_IsPersistent, &is_persistent, EV_RET_PERSIST_UNSAT);
    EV_CHECK(EV_RET_PERSIST_EQ, fr_eq_u64(is_persistent, is_return ? 1 : 0));
    EV_CHECK(EV_RET_RWC, fr_eq(s.nxt(S_RWC), fr_add(rwc, fr_add_u64(delta, 1))));
    EV_CHECK(EV_RET_GAS, fr_eq(s.nxt(S_GAS), gas_left));
    EV_CHECK(EV_RET_CALL_ID, fr_eq(s.nxt(S_CALL_ID), call_id));
  } else {  // C
    restore_context_f(s, true, look, delta, offset, length, fr_sub(gas_left, fr_u64(expansion)), true);
  }
}

// ---- ErrorOutOfGasCall: error_oog_call.py:11-42 with util/call_gadget.py:39-125 (CallGadget, IS_SUCCESS_CALL = 0) ----
ZK_HD_NOINLINE void gadget_error_oog_call(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), sp = s.cur(S_SP);
  const u64 op = (fr_fits64(opcode) && opcode.l[0] < 256) ? opcode.l[0] : 0x100;
  EV_CHECK(EV_EOC_OPCODE, op == 0xf1 || op == 0xf2 || op == 0xf4 || op == 0xfa);
  const bool has_value_pop = op == 0xf1 || op == 0xf2;
  Fr tx_id;
  ST_CC(0, ZK_CC_TxId, &tx_id, EV_EOC_TXID_UNSAT);
  // CallGadget: gas, callee, value (CALL / CALLCODE only), cd_offset, cd_length, rd_offset, rd_length, then the pushed result
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 w[8] = {zero, zero, zero, zero, zero, zero, zero, zero};
  u64 k = 1, spo = 0;
#pragma unroll
  for (int f = 0; f < 7; f++) {
    if (f == 2 && !has_value_pop) continue;
    if (!need1(s, true, stack_at(s, true, k, 0, fr_add_u64(sp, spo), &w[f]), EV_EOC_POP0_UNSAT + 2 * f)) return;
    k++;
    spo++;
  }
  if (!need1(s, true, stack_at(s, true, k, 1, fr_add_u64(sp, spo - 1), &w[7]), EV_EOC_PUSH_UNSAT)) return;
  k++;
  EV_CHECK(EV_EOC_RESULT_WORD, fr_is_zero(w[7].hi));
  EV_CHECK(EV_EOC_RESULT_BOOL, fr_eq_u64(w[7].lo, 0) || fr_eq_u64(w[7].lo, 1));
  EV_CHECK(EV_EOC_RESULT_ZERO, fr_is_zero(w[7].lo));
  Fr gas = fr_u64(0), callee = fr_u64(0);
  EOOG_W2FQ(w[0], 8, &gas, EV_EOC_GAS_DOMAIN);
  const bool has_value = has_value_pop && !fr_is_zero(fr_add(w[2].lo, w[2].hi));
  EOOG_W2FQ(w[1], 20, &callee, EV_EOC_CALLEE_DOMAIN);
  // memory_offset_and_length x 2: the length first, the offset only when the length is not zero
  Fr cd_off = fr_u64(0), cd_len = fr_u64(0), rd_off = fr_u64(0), rd_len = fr_u64(0);
  EOOG_W2FQ(w[4], 5, &cd_len, EV_EOC_CDLEN_DOMAIN);
  if (!fr_is_zero(cd_len)) EOOG_W2FQ(w[3], 5, &cd_off, EV_EOC_CDOFF_DOMAIN);
  EOOG_W2FQ(w[6], 5, &rd_len, EV_EOC_RDLEN_DOMAIN);
  if (!fr_is_zero(rd_len)) EOOG_W2FQ(w[5], 5, &rd_off, EV_EOC_RDOFF_DOMAIN);
  // memory_expansion_dynamic_length(cd_offset, cd_length, rd_offset, rd_length)
  const u64 cd_words = (cd_off.l[0] + cd_len.l[0] + 31) / 32, rd_words = (rd_off.l[0] + rd_len.l[0] + 31) / 32;
  EV_CHECK(EV_EOC_CD_MEMSIZE_RANGE, (cd_words >> 32) == 0);
  const Fr cur_mem = s.cur(S_MEM);
  EV_CHECK(EV_EOC_MEM_MAX, fr_fits64(cur_mem) && (cur_mem.l[0] >> 32) == 0);
  EV_CHECK(EV_EOC_RD_MEMSIZE_RANGE, (rd_words >> 32) == 0);
  u64 expansion = 0;
  (void)mem_expansion_gas(s, cd_words > rd_words ? cd_words : rd_words, &expansion);
  u32 r = 0;
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 0, callee, ZK_ACC_CodeHash, &r), EV_EOC_HASH_UNSAT);
  k++;
  {  // read_account_to_access_list: state_read(TxAccessListAccount, tx_id, callee)
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, k), 0, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id;
    key[R_ADDR] = callee;
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR), &r), EV_EOC_AL_UNSAT);
    k++;
  }
  EV_CHECK(EV_EOC_AL_PREV_TYPE, !rw_flag(s, r, 1));
  const Fr is_warm = rw_cell(s, R_PREV_LO, r);
  const bool warm = fr_eq_u64(is_warm, 1);
  EV_CHECK(EV_EOC_WARM_BOOL, warm || fr_eq_u64(is_warm, 0));
  // gas_cost(): is_success == 0 here, so the new-account term vanishes
  const u64 cost = (warm ? 100 : 2600) + (has_value ? 9000 : 0) + expansion;
  const Fr gas_left = s.cur(S_GAS);
  EV_CHECK(EV_EOC_CMP_RANGE, fr_fits64(gas_left));
  EV_CHECK(EV_EOC_NOT_ENOUGH, gas_left.l[0] < cost);
  error_state_tail(s, k);
}

}  // namespace zk

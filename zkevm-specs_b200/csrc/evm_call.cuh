// Gate program of CALL / CALLCODE / DELEGATECALL / STATICCALL (group KG_TX), part of evm.cu (included there).
//   callop   evm_circuit/execution/callop.py:12-341 with util/call_gadget.py:39-125 (transfer / add_balance / sub_balance
//            instruction.py:987-1013, 1111-1120, reversion_info :901-913, state_write :826-863,
//            step_state_transition_to_new_context :266-290)
// Two endings: the callee has no code or the pre-check (depth, balance) fails -> the step stays in the caller with an empty
// return-data record; the callee has code -> the caller's state is saved in its call context, the callee's 18 context
// cells are read back and the next step starts the new call.  The precompile branch (callop.py:158-277) reads
// StepState.aux_data, which the 13-cell step layout of this build does not carry: EV_CALL_PRECOMPILE (NotImplementedError).
#pragma once
namespace zk {

ZK_HD_NOINLINE int cc_rw_lookup_m(const StepCtx& s, const Fr& rwc, u64 rw, const Fr& call_id, u64 field, u32* r) {
  Fr key[14];
  rw_key_init(key, rwc, rw, ZK_TARGET_CallContext);
  key[R_ID] = call_id;
  key[R_ADDR] = fr_u64(field);
  return rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR), r);
}
// call_context_lookup(field, rw, call_id) at rw_counter + k: .value() of a non-Word row
#define CALL_CCV(k, rw_, id_, field, out, base)                                             \
  do {                                                                                      \
    u32 r_ = 0;                                                                             \
    TX_LK(cc_rw_lookup_m(s, fr_add_u64(rwc, (k)), (rw_), (id_), (field), &r_), (base));     \
    TX_NOT_WORD(rw_flag(s, r_, 0), (base));                                                 \
    *(out) = rw_cell(s, R_VAL_LO, r_);                                                      \
  } while (0)
// account_write_word(address, Balance, reversion_info) at rwc_k, its reversion row at rwc_rev when not persistent
ZK_HD_NOINLINE bool balance_write(const StepCtx& s, const Fr& rwc_k, const Fr& address, const Fr& is_persistent, const Fr& rwc_rev, int id_base,
                                  u32* r_out) {
  const int n = account_lookup_m(s, rwc_k, 1, address, ZK_ACC_Balance, r_out);
  if (n != 1) {
    step_fail(s, n == 0 ? id_base : id_base + 1);
    return false;
  }
  if (fr_is_zero(is_persistent)) {
    u32 r2 = 0;
    const int m = reversion_lookup_m(s, rwc_rev, *r_out, &r2);
    if (m != 1) {
      step_fail(s, m == 0 ? id_base + 2 : id_base + 3);
      return false;
    }
  }
  return true;
}

ZK_HD_NOINLINE void gadget_callop(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  const bool is_call = fr_eq_u64(opcode, 0xf1), is_callcode = fr_eq_u64(opcode, 0xf2), is_delegate = fr_eq_u64(opcode, 0xf4);
  const bool is_staticcall = fr_eq_u64(opcode, 0xfa);
  EV_CHECK(EV_CALL_RESP_OPCODE, responsible_opcode(s, s.cur(S_STATE), opcode));
  const Fr callee_call_id = rwc;
  Fr tx_id, rev_end, is_persistent, is_static, depth;
  u32 r = 0;
  CALL_CCV(0, 0, call_id, ZK_CC_TxId, &tx_id, EV_CALL_TXID_UNSAT);
  CALL_CCV(1, 0, call_id, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_CALL_REVEND_UNSAT);
  CALL_CCV(2, 0, call_id, ZK_CC_IsPersistent, &is_persistent, EV_CALL_PERSIST_UNSAT);
  TX_LK(cc_rw_lookup_m(s, fr_add_u64(rwc, 3), 0, call_id, ZK_CC_CalleeAddress, &r), EV_CALL_SELF_UNSAT);
  const Word2 ctx_caller_w = rw_word(s, R_VAL_LO, r);
  Fr ctx_caller = fr_u64(0);
  EOOG_W2FQ(ctx_caller_w, 20, &ctx_caller, EV_CALL_SELF_DOMAIN);
  CALL_CCV(4, 0, call_id, ZK_CC_IsStatic, &is_static, EV_CALL_STATIC_UNSAT);
  CALL_CCV(5, 0, call_id, ZK_CC_Depth, &depth, EV_CALL_DEPTH_UNSAT);
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 parent_caller_w = zero, parent_value = zero;
  u64 k = 6;
  if (is_delegate) {
    TX_LK(cc_rw_lookup_m(s, fr_add_u64(rwc, 6), 0, call_id, ZK_CC_CallerAddress, &r), EV_CALL_PCALLER_UNSAT);
    parent_caller_w = rw_word(s, R_VAL_LO, r);
    TX_LK(cc_rw_lookup_m(s, fr_add_u64(rwc, 7), 0, call_id, ZK_CC_Value, &r), EV_CALL_PVALUE_UNSAT);
    parent_value = rw_word(s, R_VAL_LO, r);
    k = 8;
  }
  // CallGadget(instruction, FQ(1), ..)
  EV_CHECK(EV_CALL_OPCODE, (int)is_call + (int)is_callcode + (int)is_delegate + (int)is_staticcall == 1);
  const bool has_value_pop = is_call || is_callcode;
  Word2 w[8] = {zero, zero, zero, zero, zero, zero, zero, zero};
  u64 spo = 0;
#pragma unroll
  for (int f = 0; f < 7; f++) {
    if (f == 2 && !has_value_pop) continue;
    if (!need1(s, true, stack_at(s, true, k, 0, fr_add_u64(sp, spo), &w[f]), EV_CALL_POP0_UNSAT + 2 * f)) return;
    k++;
    spo++;
  }
  if (!need1(s, true, stack_at(s, true, k, 1, fr_add_u64(sp, spo - 1), &w[7]), EV_CALL_PUSH_UNSAT)) return;
  k++;
  EV_CHECK(EV_CALL_RESULT_WORD, fr_is_zero(w[7].hi));
  const Fr is_success = w[7].lo;
  const bool success = fr_eq_u64(is_success, 1);
  EV_CHECK(EV_CALL_RESULT_BOOL, success || fr_is_zero(is_success));
  Fr gas = fr_u64(0), callee = fr_u64(0);
  EOOG_W2FQ(w[0], 8, &gas, EV_CALL_GAS_DOMAIN);  // after this, is_u64_gas == 1
  const bool has_value = has_value_pop && !fr_is_zero(fr_add(w[2].lo, w[2].hi));
  EOOG_W2FQ(w[1], 20, &callee, EV_CALL_CALLEE_DOMAIN);
  Fr cd_off = fr_u64(0), cd_len = fr_u64(0), rd_off = fr_u64(0), rd_len = fr_u64(0);
  EOOG_W2FQ(w[4], 5, &cd_len, EV_CALL_CDLEN_DOMAIN);
  if (!fr_is_zero(cd_len)) EOOG_W2FQ(w[3], 5, &cd_off, EV_CALL_CDOFF_DOMAIN);
  EOOG_W2FQ(w[6], 5, &rd_len, EV_CALL_RDLEN_DOMAIN);
  if (!fr_is_zero(rd_len)) EOOG_W2FQ(w[5], 5, &rd_off, EV_CALL_RDOFF_DOMAIN);
  const u64 cd_words = (cd_off.l[0] + cd_len.l[0] + 31) / 32, rd_words = (rd_off.l[0] + rd_len.l[0] + 31) / 32;
  EV_CHECK(EV_CALL_CD_MEMSIZE_RANGE, (cd_words >> 32) == 0);
  const Fr cur_mem = s.cur(S_MEM);
  EV_CHECK(EV_CALL_MEM_MAX, fr_fits64(cur_mem) && (cur_mem.l[0] >> 32) == 0);
  EV_CHECK(EV_CALL_RD_MEMSIZE_RANGE, (rd_words >> 32) == 0);
  u64 next_mem = cur_mem.l[0] < cd_words ? cd_words : cur_mem.l[0];
  next_mem = next_mem < rd_words ? rd_words : next_mem;
  const u64 expansion = memory_gas_cost(next_mem) - memory_gas_cost(cur_mem.l[0]);
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 0, callee, ZK_ACC_CodeHash, &r), EV_CALL_HASH_UNSAT);
  k++;
  const Word2 callee_hash = rw_word(s, R_VAL_LO, r);
  const Word2 empty{fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
  // is_equal_word / is_zero_word: the FIELD SUM of the halves (differences) is zero
  const bool is_empty_hash = fr_is_zero(fr_add(fr_sub(callee_hash.lo, empty.lo), fr_sub(callee_hash.hi, empty.hi)));
  const bool not_exists = fr_is_zero(fr_add(callee_hash.lo, callee_hash.hi));
  const Fr callee_address = (is_callcode || is_delegate) ? ctx_caller : callee;  // < 2^160: address_to_word holds
  const Word2 callee_address_w{fr_u128(callee_address.l[0], callee_address.l[1]), fr_u64(callee_address.l[2])};
  const Word2 caller_address_w = is_delegate ? parent_caller_w : ctx_caller_w;
  EV_CHECK(EV_CALL_CALLER_WORD, !is_delegate || word_in_domain(parent_caller_w));
  Fr caller_address = fr_u64(0);
  EOOG_W2FQ(caller_address_w, 20, &caller_address, EV_CALL_CALLER_DOMAIN);
  // add_account_to_access_list(tx_id, call.callee_address, reversion_info)
  Fr rev_count = s.cur(S_REV);
  {
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, k), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id;
    key[R_ADDR] = callee;
    key[R_VAL_LO] = fr_u64(1);
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_VAL_LO) | ZK_RWM(R_VAL_HI), &r), EV_CALL_AL_UNSAT);
    k++;
    const u32 first = r;
    if (fr_is_zero(is_persistent)) {
      u32 r2 = 0;
      TX_LK(reversion_lookup_m(s, fr_sub(rev_end, rev_count), first, &r2), EV_CALL_AL_REV_UNSAT);
      rev_count = fr_add_u64(rev_count, 1);
    }
    EV_CHECK(EV_CALL_AL_PREV_TYPE, !rw_flag(s, first, 1));
    r = first;
  }
  const Fr is_warm = rw_cell(s, R_PREV_LO, r);
  EV_CHECK(EV_CALL_VALUE_STATIC, !has_value || fr_is_zero(is_static));
  Fr callee_rev_end, callee_persistent;
  CALL_CCV(k, 0, callee_call_id, ZK_CC_RwCounterEndOfReversion, &callee_rev_end, EV_CALL_CREVEND_UNSAT);
  k++;
  CALL_CCV(k, 0, callee_call_id, ZK_CC_IsPersistent, &callee_persistent, EV_CALL_CPERSIST_UNSAT);
  k++;
  EV_CHECK(EV_CALL_CPERSIST_EQ, fr_eq(callee_persistent, fr_mul_sel(is_persistent, is_success)));
  if (success && fr_is_zero(is_persistent)) {
    EV_CHECK(EV_CALL_CREVEND_EQ, fr_eq(callee_rev_end, fr_sub(rev_end, rev_count)));
    rev_count = fr_add_u64(rev_count, 1);
  }
  bool insufficient = false;
  if (has_value_pop) {
    TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 0, caller_address, ZK_ACC_Balance, &r), EV_CALL_BAL_UNSAT);
    k++;
    const Word2 bal = rw_word(s, R_VAL_LO, r);
    EV_CHECK(EV_CALL_BAL_CMP_RANGE, word_in_domain(bal) && word_in_domain(w[2]));
    insufficient = fr_lt(bal.hi, w[2].hi) || (fr_eq(bal.hi, w[2].hi) && fr_lt(bal.lo, w[2].lo));
  }
  EV_CHECK(EV_CALL_DEPTH_RANGE, fr_fits64(depth) && (depth.l[0] >> 16) == 0);
  const bool precheck_ok = depth.l[0] < 1025 && !insufficient;
  if (!precheck_ok) EV_CHECK(EV_CALL_PRECHECK_SUCCESS, fr_is_zero(is_success));
  if (is_call && precheck_ok) {  // transfer(caller, callee, value, callee_reversion_info)
    if (!balance_write(s, fr_add_u64(rwc, k), caller_address, callee_persistent, callee_rev_end, EV_CALL_SEND_UNSAT, &r)) return;
    k++;
    {
      const Word2 ws[2] = {rw_word(s, R_VAL_LO, r), w[2]};
      Fr carry;
      const Word2 sum = add_words_n(ws, 2, &carry);
      EV_CHECK(EV_CALL_SEND_EQ, word_eq(rw_word(s, R_PREV_LO, r), sum));
      EV_CHECK(EV_CALL_SEND_CARRY, fr_is_zero(carry));
    }
    if (!balance_write(s, fr_add_u64(rwc, k), callee_address, callee_persistent, fr_sub(callee_rev_end, fr_u64(1)), EV_CALL_RECV_UNSAT, &r)) return;
    k++;
    {
      const Word2 ws[2] = {rw_word(s, R_PREV_LO, r), w[2]};
      Fr carry;
      const Word2 sum = add_words_n(ws, 2, &carry);
      EV_CHECK(EV_CALL_RECV_EQ, word_eq(rw_word(s, R_VAL_LO, r), sum));
      EV_CHECK(EV_CALL_RECV_CARRY, fr_is_zero(carry));
    }
  }
  if (is_callcode && success) EV_CHECK(EV_CALL_CALLCODE_BALANCE, !insufficient);
  // gas: call.gas_cost(instruction, is_warm_access, is_call), EIP-150
  const bool warm = fr_eq_u64(is_warm, 1);
  EV_CHECK(EV_CALL_WARM_BOOL, warm || fr_is_zero(is_warm));
  const u64 gas_cost = (warm ? 100 : 2600) + (has_value ? 9000 + ((is_call && success && not_exists) ? 25000 : 0) : 0) + expansion;
  const Fr gas_available = fr_sub(s.cur(S_GAS), fr_u64(gas_cost));
  // constant_divmod(gas_available, 64, 8): the quotient of the integer below p must fit 8 bytes
  const Fr one_64th{{(gas_available.l[0] >> 6) | (gas_available.l[1] << 58), (gas_available.l[1] >> 6) | (gas_available.l[2] << 58),
                     (gas_available.l[2] >> 6) | (gas_available.l[3] << 58), gas_available.l[3] >> 6}};
  EV_CHECK(EV_CALL_GAS_64TH_RANGE, fr_fits64(one_64th));
  const Fr all_but = fr_sub(gas_available, one_64th);
  EV_CHECK(EV_CALL_GAS_MIN_RANGE, fr_fits64(all_but));  // min(all_but_one_64th_gas, call.gas, 8)
  Fr callee_gas_left = all_but.l[0] < gas.l[0] ? all_but : gas;
  const bool is_precompile = fr_fits64(callee) && callee.l[0] >= 1 && callee.l[0] <= 9;
  const Fr ns = s.nxt(S_STATE);
  const bool next_is_precompile = fr_fits64(ns) && ns.l[0] >= ZK_ES_ECRECOVER && ns.l[0] <= ZK_ES_ECRECOVER + 8;
  EV_CHECK(EV_CALL_PRECOMPILE_STATE, is_precompile == next_is_precompile);
  const u64 sp_delta = 5 + (is_call ? 1 : 0) + (is_callcode ? 1 : 0);
  const int no_callee_code = (is_empty_hash ? 1 : 0) + (not_exists ? 1 : 0);
  if (!precheck_ok || (no_callee_code == 1 && !is_precompile)) {
    const u64 TAGS[3] = {ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataOffset, ZK_CC_LastCalleeReturnDataLength};
#pragma unroll 1
    for (int t = 0; t < 3; t++) {
      Fr v;
      CALL_CCV(k, 1, call_id, TAGS[t], &v, EV_CALL_LAST0_UNSAT + 4 * t);
      k++;
      EV_CHECK(EV_CALL_LAST0_UNSAT + 4 * t + 3, fr_is_zero(v));
    }
    EV_CHECK(EV_CALL_SAME_RWC, fr_eq(s.nxt(S_RWC), fr_add_u64(rwc, k)));
    EV_CHECK(EV_CALL_SAME_PC, fr_eq(s.nxt(S_PC), fr_add_u64(s.cur(S_PC), 1)));
    EV_CHECK(EV_CALL_SAME_SP, fr_eq(s.nxt(S_SP), fr_add_u64(sp, sp_delta)));
    EV_CHECK(EV_CALL_SAME_GAS, fr_eq(s.nxt(S_GAS), fr_sub(fr_add_u64(s.cur(S_GAS), has_value ? 2300 : 0), fr_u64(gas_cost))));
    EV_CHECK(EV_CALL_SAME_MEM, fr_eq_u64(s.nxt(S_MEM), next_mem));
    EV_CHECK(EV_CALL_SAME_REV, fr_eq(s.nxt(S_REV), fr_add_u64(s.cur(S_REV), 3)));
    EV_CHECK(EV_CALL_SAME_CALL_ID, fr_eq(s.nxt(S_CALL_ID), call_id));
    EV_CHECK(EV_CALL_SAME_IS_ROOT, fr_eq(s.nxt(S_IS_ROOT), s.cur(S_IS_ROOT)));
    EV_CHECK(EV_CALL_SAME_IS_CREATE, fr_eq(s.nxt(S_IS_CREATE), s.cur(S_IS_CREATE)));
    EV_CHECK(EV_CALL_SAME_CODE_HASH, fr_eq(s.nxt(S_HASH_LO), s.cur(S_HASH_LO)) && fr_eq(s.nxt(S_HASH_HI), s.cur(S_HASH_HI)));
    return;
  }
  EV_CHECK(EV_CALL_PRECOMPILE, !is_precompile);  // needs StepState.aux_data
  {  // save the caller's state: 5 call-context writes on the current call
    const u64 TAGS[5] = {ZK_CC_ProgramCounter, ZK_CC_StackPointer, ZK_CC_GasLeft, ZK_CC_MemorySize, ZK_CC_ReversibleWriteCounter};
#pragma unroll 1
    for (int t = 0; t < 5; t++) {
      const Fr want = t == 0   ? fr_add_u64(s.cur(S_PC), 1)
                      : t == 1 ? fr_add_u64(sp, sp_delta)
                      : t == 2 ? fr_sub(fr_sub(s.cur(S_GAS), fr_u64(gas_cost)), callee_gas_left)
                      : t == 3 ? fr_u64(next_mem)
                               : fr_add_u64(s.cur(S_REV), 1);
      Fr v;
      CALL_CCV(k, 1, call_id, TAGS[t], &v, EV_CALL_SAVE0_UNSAT + 4 * t);
      k++;
      EV_CHECK(EV_CALL_SAVE0_UNSAT + 4 * t + 3, fr_eq(v, want));
    }
  }
  {  // the callee's context: 18 call-context reads compared as words (lo, hi)
    const Word2 value_w = is_delegate ? parent_value : w[2];
    EV_CHECK(EV_CALL_VALUE_WORD, word_in_domain(value_w));
    const u64 TAGS[18] = {ZK_CC_CallerId, ZK_CC_TxId, ZK_CC_Depth, ZK_CC_CallerAddress, ZK_CC_CalleeAddress, ZK_CC_CallDataOffset,
                          ZK_CC_CallDataLength, ZK_CC_ReturnDataOffset, ZK_CC_ReturnDataLength, ZK_CC_Value, ZK_CC_IsSuccess,
                          ZK_CC_IsStatic, ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataOffset, ZK_CC_LastCalleeReturnDataLength,
                          ZK_CC_IsRoot, ZK_CC_IsCreate, ZK_CC_CodeHash};
#pragma unroll 1
    for (int t = 0; t < 18; t++) {
      Word2 want = zero;
      switch (t) {
        case 0: want.lo = call_id; break;
        case 1: want.lo = tx_id; break;
        case 2: want.lo = fr_add_u64(depth, 1); break;
        case 3: want = caller_address_w; break;
        case 4: want = callee_address_w; break;
        case 5: want.lo = cd_off; break;
        case 6: want.lo = cd_len; break;
        case 7: want.lo = rd_off; break;
        case 8: want.lo = rd_len; break;
        case 9: want = value_w; break;
        case 10: want.lo = is_success; break;
        case 11: want.lo = is_static; break;
        case 17: want = callee_hash; break;
        default: break;
      }
      TX_LK(cc_rw_lookup_m(s, fr_add_u64(rwc, k), 0, callee_call_id, TAGS[t], &r), EV_CALL_CTX0_UNSAT + 3 * t);
      k++;
      EV_CHECK(EV_CALL_CTX0_UNSAT + 3 * t + 2, word_eq(rw_word(s, R_VAL_LO, r), want));
    }
  }
  callee_gas_left = fr_add_u64(callee_gas_left, has_value ? 2300 : 0);
  EV_CHECK(EV_CALL_NC_RWC, fr_eq(s.nxt(S_RWC), fr_add_u64(rwc, k)));
  EV_CHECK(EV_CALL_NC_CALL_ID, fr_eq(s.nxt(S_CALL_ID), callee_call_id));
  EV_CHECK(EV_CALL_NC_IS_ROOT, fr_is_zero(s.nxt(S_IS_ROOT)));
  EV_CHECK(EV_CALL_NC_IS_CREATE, fr_is_zero(s.nxt(S_IS_CREATE)));
  EV_CHECK(EV_CALL_NC_CODE_HASH, fr_eq(s.nxt(S_HASH_LO), callee_hash.lo) && fr_eq(s.nxt(S_HASH_HI), callee_hash.hi));
  EV_CHECK(EV_CALL_NC_GAS, fr_eq(s.nxt(S_GAS), callee_gas_left));
  EV_CHECK(EV_CALL_NC_REV, fr_eq_u64(s.nxt(S_REV), 2));
  EV_CHECK(EV_CALL_NC_LOG, fr_eq(s.nxt(S_LOG), s.cur(S_LOG)));
  EV_CHECK(EV_CALL_NC_PC, fr_is_zero(s.nxt(S_PC)));
  EV_CHECK(EV_CALL_NC_SP, fr_eq_u64(s.nxt(S_SP), 1024));
  EV_CHECK(EV_CALL_NC_MEM, fr_is_zero(s.nxt(S_MEM)));
}

}  // namespace zk

// Gate programs of LOG0..LOG4 / ErrorWriteProtection / BLOCKHASH (group KG_TX), part of evm.cu (included there).
//   log                      evm_circuit/execution/log.py:8-102 (tx_log_lookup_word instruction.py:708-720,
//                            copy_lookup with a TxLog destination table.py:760-787)
//   error_write_protection   evm_circuit/execution/error_write_protection.py:12-66
//   blockhash                evm_circuit/execution/blockhash.py:6-38
#pragma once
namespace zk {

enum { ZK_TXLOG_Address = 1, ZK_TXLOG_Topic = 2, ZK_TXLOG_Data = 3, ZK_BLOCK_Number = 3, ZK_BLOCK_HistoryHash = 8 };
// log_id << 48 as a field element
ZK_HD Fr log_id_shift(const Fr& log_id) { return fr_mul(log_id, Fr{{1ull << 48, 0, 0, 0}}); }
// rw_lookup(Write, TxLog, id = tx_id, address = index + (field << 32) + (log_id << 48), field_tag = 0, storage_key = Word(0))
ZK_HD_NOINLINE int tx_log_lookup_m(const StepCtx& s, const Fr& rwc_k, const Fr& tx_id, const Fr& log_shifted, u64 field, u64 index, u32* r) {
  Fr key[14];
  rw_key_init(key, rwc_k, 1, ZK_TARGET_TxLog);
  key[R_ID] = tx_id;
  key[R_ADDR] = fr_add_u64(log_shifted, index + (field << 32));
  return rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_FIELD) | ZK_RWM(R_KEY_LO) | ZK_RWM(R_KEY_HI), r);
}
// same context + log_id = Transition.delta(d_log), memory_word_size = Transition.to(mem_value)
ZK_HD_NOINLINE void same_context_l_ni(const StepCtx& s, const Fr& opcode, const Fr& d_rwc, const Fr& d_pc, const Fr& d_sp, const Fr& mem_value,
                                      const Fr& dyn_gas, const Fr& d_log) {
  same_context_x(s, opcode, d_rwc, d_pc, d_sp, true, mem_value, dyn_gas, 0, &d_log);
}

ZK_HD_NOINLINE void gadget_log(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  const Fr topics = fr_sub_u64(opcode, 0xa0);
  {
    Fr key[4] = {fr_u64(ZK_FIXED_Range5), topics, fr_u64(0), fr_u64(0)};
    u32 r = 0;
    EV_CHECK(EV_LOG_RANGE5, lookup<4>(s.t.fixed, key, &r) >= 1);
  }
  Word2 start_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  Fr mstart = fr_u64(0), msize = fr_u64(0);
  if (!need1(s, true, stack_at(s, true, 0, 0, sp, &start_w), EV_LOG_POP0_UNSAT)) return;
  EOOG_W2FQ(start_w, 8, &mstart, EV_LOG_START_DOMAIN);
  if (!need1(s, true, stack_at(s, true, 1, 0, fr_add_u64(sp, 1), &size_w), EV_LOG_POP1_UNSAT)) return;
  EOOG_W2FQ(size_w, 8, &msize, EV_LOG_SIZE_DOMAIN);
  Fr tx_id, is_static, is_persistent;
  ST_CC(2, ZK_CC_TxId, &tx_id, EV_LOG_TXID_UNSAT);
  ST_CC(3, ZK_CC_IsStatic, &is_static, EV_LOG_STATIC_UNSAT);
  EV_CHECK(EV_LOG_STATIC_NONZERO, fr_is_zero(is_static));
  u32 r = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 4), call_id, ZK_CC_CalleeAddress, &r), EV_LOG_CALLEE_UNSAT);
  const Word2 address = rw_word(s, R_VAL_LO, r);
  ST_CC(5, ZK_CC_IsPersistent, &is_persistent, EV_LOG_PERSIST_UNSAT);
  const bool persistent = !fr_is_zero(is_persistent);
  const Fr log_shifted = log_id_shift(fr_add_u64(s.cur(S_LOG), 1));
  u64 k = 6;
  if (persistent) {
    TX_LK(tx_log_lookup_m(s, fr_add_u64(rwc, k), tx_id, log_shifted, ZK_TXLOG_Address, 0, &r), EV_LOG_ADDR_UNSAT);
    EV_CHECK(EV_LOG_ADDR_EQ, word_eq(address, rw_word(s, R_VAL_LO, r)));
    k++;
  }
  const u64 n_topics = topics.l[0];  // 0..4: a row of the Range5 fixed table
#pragma unroll 1
  for (u64 t = 0; t < n_topics && t < 4; t++) {
    Word2 topic{fr_u64(0), fr_u64(0)};
    if (!need1(s, true, stack_at(s, true, k, 0, fr_add_u64(sp, 2 + t), &topic), EV_LOG_TOPIC_POP_UNSAT)) return;
    k++;
    if (persistent) {
      TX_LK(tx_log_lookup_m(s, fr_add_u64(rwc, k), tx_id, log_shifted, ZK_TXLOG_Topic, t, &r), EV_LOG_TOPIC_UNSAT);
      EV_CHECK(EV_LOG_TOPIC_EQ, word_eq(topic, rw_word(s, R_VAL_LO, r)));
      k++;
    }
  }
  Fr rwc_inc = fr_u64(0);
  if (!fr_is_zero(msize) && fr_eq_u64(is_persistent, 1)) {
    Fr unused = fr_u64(0);
    const Fr dst = fr_add_u64(log_shifted, (u64)ZK_TXLOG_Data << 32);
    if (!need1(s, true, copy_lookup(s, true, call_id, ZK_COPY_Memory, tx_id, ZK_COPY_TxLog, mstart, fr_add(mstart, msize), dst, msize,
                                    fr_add_u64(rwc, k), &rwc_inc, &unused), EV_LOG_COPY_UNSAT)) return;
  }
  // memory_expansion_dynamic_length(mstart, msize), instruction.py:1157-1181
  const unsigned __int128 words = ((unsigned __int128)mstart.l[0] + msize.l[0] + 31) / 32;
  EV_CHECK(EV_LOG_MEMSIZE_RANGE, (words >> 32) == 0);
  u64 expansion = 0;
  EV_CHECK(EV_LOG_MEM_MAX, mem_expansion_gas(s, (u64)words, &expansion) == 0);
  const u64 cur = s.cur(S_MEM).l[0], nxt = cur < (u64)words ? (u64)words : cur;
  const u64 gas = 375 + 375 * n_topics + 8 * msize.l[0] + expansion;
  same_context_l_ni(s, opcode, fr_add_u64(rwc_inc, k), fr_u64(1), fr_add_u64(topics, 2), fr_u64(nxt), fr_u64(gas), is_persistent);
}

ZK_HD_NOINLINE void gadget_error_write_protection(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const u64 op = (fr_fits64(opcode) && opcode.l[0] < 256) ? opcode.l[0] : 0x100;
  EV_CHECK(EV_EWP_OPCODE, op == 0x55 || op == 0xf0 || op == 0xf5 || op == 0xf1 || op == 0xff || (op >= 0xa0 && op <= 0xa4));
  Fr is_static;
  ST_CC(0, ZK_CC_IsStatic, &is_static, EV_EWP_STATIC_UNSAT);
  EV_CHECK(EV_EWP_NOT_STATIC, fr_eq_u64(is_static, 1));
  u64 n_rw = 1;
  if (op == 0xf1) {  // CALL: the transferred value (third stack word) must not be zero
    Word2 value{fr_u64(0), fr_u64(0)};
    if (!need1(s, true, stack_at(s, true, 1, 0, fr_add_u64(s.cur(S_SP), 2), &value), EV_EWP_VALUE_UNSAT)) return;
    EV_CHECK(EV_EWP_VALUE_ZERO, !(fr_is_zero(value.lo) && fr_is_zero(value.hi)));
    n_rw = 2;
  }
  error_state_tail(s, n_rw);
}

ZK_HD_NOINLINE void gadget_blockhash(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr sp = s.cur(S_SP);
  Word2 num_w{fr_u64(0), fr_u64(0)}, pushed{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, sp, &num_w), EV_BH_POP_UNSAT)) return;
  Fr number = fr_u64(0);
  EOOG_W2FQ(num_w, 8, &number, EV_BH_NUM_DOMAIN);
  u32 r = 0;
  TX_LK(block_lookup_m(s, ZK_BLOCK_Number, &r), EV_BH_CUR_UNSAT);
  TX_NOT_WORD(block_is_word(s, r), EV_BH_CUR_UNSAT);
  const Fr current = block_word(s, r).lo;
  if (!need1(s, true, stack_at(s, true, 1, 1, sp, &pushed), EV_BH_PUSH_UNSAT)) return;
  // compare(block_number, current, 8) and compare(current, 256 + block_number, 2): range asserts
  EV_CHECK(EV_BH_CMP1_RANGE, fr_fits64(current));
  const Fr limit = fr_add_u64(number, 256);
  EV_CHECK(EV_BH_CMP2_RANGE, (current.l[0] >> 16) == 0 && fr_fits64(limit) && (limit.l[0] >> 16) == 0);
  Word2 want{fr_u64(0), fr_u64(0)};
  if (number.l[0] < current.l[0] && current.l[0] < limit.l[0]) {
    Fr key[2] = {fr_u64(ZK_BLOCK_HistoryHash), number};
    TX_LK(lookup<2>(s.t.block, key, &r), EV_BH_HASH_UNSAT);
    want = block_word(s, r);
  }
  EV_CHECK(EV_BH_EQ, word_eq(pushed, want));
  same_context_ni(s, opcode, 2, fr_u64(1), fr_u64(0));
}

// ---- error_code_store.py:14-52 (ErrorMaxCodeSizeExceeded, ErrorOutOfGasCodeStore), error_invalid_creation_code.py:11-34 ----
ZK_HD_NOINLINE void gadget_error_code_store(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_ECS_OPCODE, fr_eq_u64(opcode, 0xf3));
  EV_CHECK(EV_ECS_IS_CREATE, fr_eq_u64(s.cur(S_IS_CREATE), 1));
  Word2 len_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, fr_add_u64(s.cur(S_SP), 1), &len_w), EV_ECS_LEN_UNSAT)) return;
  Fr length = fr_u64(0);
  EOOG_W2FQ(len_w, 5, &length, EV_ECS_LEN_DOMAIN);
  Fr is_static;
  ST_CC(1, ZK_CC_IsStatic, &is_static, EV_ECS_STATIC_UNSAT);
  EV_CHECK(EV_ECS_STATIC_NONZERO, fr_is_zero(is_static));
  EV_CHECK(EV_ECS_SIZE_RANGE, (length.l[0] >> 16) == 0);  // compare(MAX_CODE_SIZE, return_length, N_BYTES_STACK = 2)
  const Fr gas_left = s.cur(S_GAS);
  EV_CHECK(EV_ECS_GAS_RANGE, fr_fits64(gas_left));  // compare(gas_left, 200 * length, 8)
  EV_CHECK(EV_ECS_NEITHER, 24576 < length.l[0] || gas_left.l[0] < 200 * length.l[0]);
  error_state_tail(s, 2);
}
ZK_HD_NOINLINE void gadget_error_invalid_creation_code(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_ECS_OPCODE, fr_eq_u64(opcode, 0xf3));
  EV_CHECK(EV_ECS_IS_CREATE, fr_eq_u64(s.cur(S_IS_CREATE), 1));
  Word2 off_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, s.cur(S_SP), &off_w), EV_ECS_LEN_UNSAT)) return;
  Fr offset = fr_u64(0);
  EOOG_W2FQ(off_w, 5, &offset, EV_ECS_LEN_DOMAIN);
  Fr key[14];
  rw_key_init(key, fr_add_u64(s.cur(S_RWC), 1), 0, ZK_TARGET_Memory);
  key[R_ID] = s.cur(S_CALL_ID);
  key[R_ADDR] = offset;
  u32 r = 0;
  TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR), &r), EV_ECS_BYTE_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ECS_BYTE_UNSAT);
  EV_CHECK(EV_ECS_FIRST_BYTE, fr_eq_u64(rw_cell(s, R_VAL_LO, r), 0xEF));
  error_state_tail(s, 2);
}

}  // namespace zk

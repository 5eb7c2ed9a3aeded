// evm_err.cuh — gate programs of the error states and SELFBALANCE (included by evm.cu after evm_tx.cuh).
//
// constrain_error_state (evm_circuit/instruction.py:1426-1452) is the tail every error state shares: the call's
// IsSuccess is 0, a root call goes to EndTx, an internal call restores its caller's context — and the rw counter also
// advances by reversible_write_counter (the reverted writes consume counters without being looked up here).
//   execution/error_stack.py, error_invalid_opcode.py, error_oog_constant.py, error_invalid_jump.py, selfbalance.py
// These run in k_evm_group<KG_TX> (lane-private lookups, a step returns at its first failing constraint).
#pragma once

namespace zk {

ZK_HD_NOINLINE void error_state_tail(const StepCtx& s, u64 n_rw) {
  const Fr rwc = s.cur(S_RWC), rev = s.cur(S_REV), call_id = s.cur(S_CALL_ID);
  Word2 v{fr_u64(0), fr_u64(0)};
  bool w = false;
  if (!need1(s, true, call_context_w(s, true, fr_add_u64(rwc, n_rw), 0, call_id, ZK_CC_IsSuccess, &v, &w), EV_ERR_CC_UNSAT)) return;
  EV_CHECK(EV_ERR_CC_TYPE, !w);
  EV_CHECK(EV_ERR_IS_SUCCESS, fr_is_zero(v.lo));
  const Fr is_root = s.cur(S_IS_ROOT);
  EV_CHECK(EV_ERR_ROOT_ENDTX, fr_eq_u64(is_root, fr_eq_u64(s.nxt(S_STATE), ZK_ES_EndTx) ? 1 : 0));
  if (!fr_is_zero(is_root)) {
    EV_CHECK(EV_ERR_RWC, fr_eq(s.nxt(S_RWC), fr_add(fr_add_u64(rwc, n_rw + 1), rev)));
    EV_CHECK(EV_ERR_CALL_ID, fr_eq(s.nxt(S_CALL_ID), call_id));
  } else {
    restore_context_x(s, true, n_rw + 1, fr_u64(0), fr_u64(0), fr_u64(0), false, rev);
  }
}

ZK_HD_NOINLINE void gadget_error_stack(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  // responsible_opcode_lookup(opcode, stack_pointer): the aux cell names the pointer that under- / overflows
  Fr key[4] = {fr_u64(ZK_FIXED_ResponsibleOpcode), s.cur(S_STATE), opcode, s.cur(S_SP)};
  u32 r;
  EV_CHECK(EV_ESTK_RESP_OPCODE, lookup<4>(s.t.fixed, key, &r) >= 1);
  error_state_tail(s, 0);
}
ZK_HD_NOINLINE void gadget_error_invalid_opcode(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EINV_RESP_OPCODE, responsible_opcode(s, s.cur(S_STATE), opcode));
  error_state_tail(s, 0);
}
ZK_HD_NOINLINE void gadget_error_oog_constant(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EOGC_OPCODE_VALUE, fr_fits64(opcode) && opcode.l[0] < 256 && OPCODE_GAS(opcode.l[0]) >= 0);
  const u64 gas = (u64)OPCODE_GAS(opcode.l[0]);
  Fr key[4] = {fr_u64(ZK_FIXED_OpcodeConstantGas), opcode, fr_u64(gas), fr_u64(0)};
  u32 r;
  EV_CHECK(EV_EOGC_GAS_UNSAT, lookup<4>(s.t.fixed, key, &r) >= 1);
  const Fr gas_left = s.cur(S_GAS);
  EV_CHECK(EV_EOGC_CMP_RANGE, fr_fits64(gas_left));
  EV_CHECK(EV_EOGC_NOT_ENOUGH, gas_left.l[0] < gas);
  error_state_tail(s, 0);
}

// bytecode_lookup_pair (instruction.py:765-769): key (hash, Byte, index), is_code NOT queried -> (value, is_code).
// Positional table: the run of the hash gives the only candidate row; otherwise the index on the first four cells
// (built only when an ErrorInvalidJump step exists, api.cu).
ZK_HD_NOINLINE int bytecode_pair_m(const StepCtx& s, const Fr& hlo, const Fr& hhi, const Fr& index, Fr* value, Fr* is_code) {
  const IndexDev& ix = s.t.bytecode;
  if (ix.tab.n_rows == 0) return 0;
  u32 r = 0;
  if (pos_enabled(ix) && ix.pos_kind == ZK_POS_RUNS) {
    u32 head = 0, len = 0;
    const int nh = heads_probe(ix, hlo, hhi, &head, &len, s.mask, true);
    if (nh != 1 || !fr_fits64(index) || index.l[0] >= (u64)len) return 0;
    r = (u32)((u64)head + 1 + index.l[0]);
  } else {
    Fr key[4] = {hlo, hhi, fr_u64(2), index};
    const int n = lookup<4>(s.t.bytecode4, key, &r);
    if (n != 1) return n;
  }
  *value = table_cell(ix.tab, B_VALUE, r);
  *is_code = table_cell(ix.tab, B_ISCODE, r);
  return 1;
}
// NB constrain_error_state sits INSIDE `if within_range == FQ(1)` (error_invalid_jump.py:24-33): a destination at or
// beyond the code length leaves the step's ending unconstrained
ZK_HD_NOINLINE void gadget_error_invalid_jump(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EJMP_OPCODE, fr_eq_u64(opcode, 0x56) || fr_eq_u64(opcode, 0x57));
  const bool is_jumpi = fr_eq_u64(opcode, 0x57);
  const Fr hlo = s.cur(S_HASH_LO), hhi = s.cur(S_HASH_HI), sp = s.cur(S_SP);
  Fr code_length = fr_u64(0);
  if (!need1(s, true, bytecode_lookup_ni(s, true, hlo, hhi, 1, fr_u64(0), 0, &code_length), EV_EJMP_LEN_UNSAT)) return;
  Word2 dest{fr_u64(0), fr_u64(0)}, cond{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, sp, &dest), EV_EJMP_DEST_UNSAT)) return;
  if (is_jumpi) {
    if (!need1(s, true, stack_at(s, true, 1, 0, fr_add_u64(sp, 1), &cond), EV_EJMP_COND_UNSAT)) return;
    EV_CHECK(EV_EJMP_COND_ZERO, !(fr_is_zero(cond.lo) && fr_is_zero(cond.hi)));
  }
  Fr d = fr_u64(0);
  const int rc = word_to_fq_n(dest, 8, &d);  // word_to_u64
  EV_CHECK(EV_EJMP_DEST_DOMAIN, rc != 1);
  EV_CHECK(EV_EJMP_DEST_U64, rc != 2);
  EV_CHECK(EV_EJMP_CMP_RANGE, fr_fits64(code_length));
  if (d.l[0] < code_length.l[0]) {
    Fr value = fr_u64(0), is_code = fr_u64(0);
    if (!need1(s, true, bytecode_pair_m(s, hlo, hhi, d, &value, &is_code), EV_EJMP_AT_UNSAT)) return;
    EV_CHECK(EV_EJMP_IS_JUMPDEST, fr_is_zero(is_code) || !fr_eq_u64(value, 0x5b));  // is_code * (value == JUMPDEST) == 0
    error_state_tail(s, 1 + (is_jumpi ? 1 : 0));
  }
}

ZK_HD_NOINLINE void gadget_selfbalance(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_SBAL_OPCODE, fr_eq_u64(opcode, 0x47));
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, rwc, call_id, ZK_CC_CalleeAddress, &r), EV_SBAL_CC_UNSAT);
  const Word2 callee = rw_word(s, R_VAL_LO, r);
  Fr address = fr_u64(0);
  const int rc = word_to_fq_n(callee, 20, &address);  // word_to_address
  EV_CHECK(EV_SBAL_ADDR_DOMAIN, rc != 1);
  EV_CHECK(EV_SBAL_ADDR_RANGE, rc != 2);
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, 1), 0, address, ZK_ACC_Balance, &r), EV_SBAL_ACC_UNSAT);
  const Word2 balance = rw_word(s, R_VAL_LO, r);
  Word2 w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 2, 1, fr_sub_u64(s.cur(S_SP), 1), &w), EV_SBAL_PUSH_UNSAT)) return;
  EV_CHECK(EV_SBAL_EQ, word_eq(w, balance));
  same_context_ni(s, opcode, 3, fr_u64(1), fr_sub(fr_u64(0), fr_u64(1)));
}

// ---- out-of-gas / out-of-bound error states --------------------------------------------------------------------------
// execution/error_oog_sha3.py, error_oog_static_memory_expansion.py, error_oog_dynamic_memory_expansion.py,
// error_oog_log.py, error_oog_exp.py, error_return_data_out_of_bound.py.  They share the EV_EOOG_* ids (one per kind of
// constraint: a step reports one id and its execution state names the gadget).
#define EOOG_W2FQ(word, nb, out, id_domain)            \
  do {                                                 \
    const int rc_ = word_to_fq_n((word), (nb), (out)); \
    EV_CHECK((id_domain), rc_ != 1);                   \
    EV_CHECK((id_domain) + 1, rc_ != 2);               \
  } while (0)
// gas of growing the memory to `words_needed` words (memory_expansion / memory_expansion_dynamic_length,
// instruction.py:1138-1177): 0 ok, 1 memory size beyond 4 bytes, 2 max(): curr.memory_word_size beyond 4 bytes
ZK_HD int mem_expansion_gas(const StepCtx& s, u64 words_needed, u64* gas) {
  if (words_needed >> 32) return 1;
  const Fr cur = s.cur(S_MEM);
  if (!(fr_fits64(cur) && (cur.l[0] >> 32) == 0)) return 2;
  const u64 nxt = cur.l[0] < words_needed ? words_needed : cur.l[0];
  *gas = memory_gas_cost(nxt) - memory_gas_cost(cur.l[0]);
  return 0;
}
#define EOOG_MEMGAS(words, gas)                               \
  do {                                                        \
    const int rc_ = mem_expansion_gas(s, (words), (gas));     \
    EV_CHECK(EV_EOOG_MEMSIZE_RANGE, rc_ != 1);                \
    EV_CHECK(EV_EOOG_MEM_MAX, rc_ != 2);                      \
  } while (0)
// compare(gas_left, cost, 8) must say "less", then constrain_error_state; cost = hi * 2^64 + lo
ZK_HD_NOINLINE void oog_finish(const StepCtx& s, u64 cost_lo, u64 cost_hi, u64 n_rw) {
  const Fr gas_left = s.cur(S_GAS);
  EV_CHECK(EV_EOOG_CMP_RANGE, fr_fits64(gas_left) && cost_hi == 0);
  EV_CHECK(EV_EOOG_NOT_ENOUGH, gas_left.l[0] < cost_lo);
  error_state_tail(s, n_rw);
}
// stack_lookup(Read, sp_off) as the k-th rw lookup of the step
#define EOOG_STACK(k, sp_off, out, id)                                                                            \
  do {                                                                                                            \
    if (!need1(s, true, stack_at(s, true, (k), 0, fr_add_u64(s.cur(S_SP), (sp_off)), (out)), (id))) return;       \
  } while (0)

ZK_HD_NOINLINE void gadget_error_oog_sha3(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x20));
  Word2 off_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  EOOG_STACK(0, 0, &off_w, EV_EOOG_POP0_UNSAT);
  EOOG_STACK(1, 1, &size_w, EV_EOOG_POP1_UNSAT);
  Fr length = fr_u64(0), offset = fr_u64(0);  // memory_offset_and_length: the length (second word) first
  EOOG_W2FQ(size_w, 5, &length, EV_EOOG_W1_DOMAIN);
  if (!fr_is_zero(length)) EOOG_W2FQ(off_w, 5, &offset, EV_EOOG_W0_DOMAIN);
  u64 expansion = 0;
  EOOG_MEMGAS((offset.l[0] + length.l[0] + 31) / 32, &expansion);
  const u64 words = (length.l[0] + 31) / 32;
  EV_CHECK(EV_EOOG_WORDSIZE_RANGE, (words >> 32) == 0);
  oog_finish(s, 30 + words * 6 + expansion, 0, 2);  // words < 2^32, expansion < 2^56: no overflow
}
// `size = 1 if is_mstore8 else 32` tests the truthiness of an FQ object (always true): one byte whatever the opcode
ZK_HD_NOINLINE void gadget_error_oog_static_memory(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x51) || fr_eq_u64(opcode, 0x52) || fr_eq_u64(opcode, 0x53));
  Word2 off_w{fr_u64(0), fr_u64(0)};
  EOOG_STACK(0, 0, &off_w, EV_EOOG_POP0_UNSAT);
  Fr offset = fr_u64(0);
  EOOG_W2FQ(off_w, 5, &offset, EV_EOOG_W0_DOMAIN);
  u64 expansion = 0;
  EOOG_MEMGAS((offset.l[0] + 1 + 31) / 32, &expansion);
  oog_finish(s, 3 + expansion, 0, 1);
}
ZK_HD_NOINLINE void gadget_error_oog_dynamic_memory(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0xf3) || fr_eq_u64(opcode, 0xfd));
  Word2 off_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  EOOG_STACK(0, 0, &off_w, EV_EOOG_POP0_UNSAT);
  EOOG_STACK(1, 1, &size_w, EV_EOOG_POP1_UNSAT);
  Fr length = fr_u64(0), offset = fr_u64(0);
  EOOG_W2FQ(size_w, 5, &length, EV_EOOG_W1_DOMAIN);
  if (!fr_is_zero(length)) EOOG_W2FQ(off_w, 5, &offset, EV_EOOG_W0_DOMAIN);
  u64 expansion = 0;  // memory_expansion: size 0 when length == 0
  EOOG_MEMGAS(fr_is_zero(length) ? 0 : (offset.l[0] + length.l[0] + 31) / 32, &expansion);
  oog_finish(s, expansion, 0, 2);
}
ZK_HD_NOINLINE void gadget_error_oog_log(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr topics = fr_sub_u64(opcode, 0xa0);
  {
    Fr key[4] = {fr_u64(ZK_FIXED_Range5), topics, fr_u64(0), fr_u64(0)};
    u32 r;
    EV_CHECK(EV_EOOG_LOG_RANGE5, lookup<4>(s.t.fixed, key, &r) >= 1);
  }
  Word2 start_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  Fr mstart = fr_u64(0), msize = fr_u64(0);
  EOOG_STACK(0, 0, &start_w, EV_EOOG_POP0_UNSAT);
  EOOG_W2FQ(start_w, 5, &mstart, EV_EOOG_W0_DOMAIN);
  EOOG_STACK(1, 1, &size_w, EV_EOOG_POP1_UNSAT);
  EOOG_W2FQ(size_w, 5, &msize, EV_EOOG_W1_DOMAIN);
  u64 expansion = 0;
  EOOG_MEMGAS((mstart.l[0] + msize.l[0] + 31) / 32, &expansion);
  oog_finish(s, 375 + 375 * topics.l[0] + 8 * msize.l[0] + expansion, 0, 2);  // topics is a Range5 value
}
ZK_HD_NOINLINE void gadget_error_oog_exp(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x0a));
  Word2 expo{fr_u64(0), fr_u64(0)};
  EOOG_STACK(0, 1, &expo, EV_EOOG_POP0_UNSAT);  // stack_lookup(Read, 1): the step's first rw lookup
  EV_CHECK(EV_EOOG_W0_DOMAIN, word_in_domain(expo));  // byte_size: to_le_bytes
  const u64 v[4] = {expo.lo.l[0], expo.lo.l[1], expo.hi.l[0], expo.hi.l[1]};
  u64 size = 0;
  for (int k = 0; k < 32; k++)
    if ((v[k >> 3] >> (8 * (k & 7))) & 0xFF) size = (u64)k + 1;
  oog_finish(s, 50 * size + 10, 0, 1);
}
ZK_HD_NOINLINE void gadget_error_return_data_oob(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_EOOG_OPCODE, fr_eq_u64(opcode, 0x3e));
  Word2 off_w{fr_u64(0), fr_u64(0)}, len_w{fr_u64(0), fr_u64(0)};
  Fr data_offset = fr_u64(0), length = fr_u64(0);
  EOOG_STACK(0, 1, &off_w, EV_EOOG_POP0_UNSAT);
  EOOG_W2FQ(off_w, 31, &data_offset, EV_EOOG_W0_DOMAIN);
  EOOG_STACK(1, 2, &len_w, EV_EOOG_POP1_UNSAT);
  EOOG_W2FQ(len_w, 31, &length, EV_EOOG_W1_DOMAIN);
  Word2 v{fr_u64(0), fr_u64(0)};
  bool w = false;
  if (!need1(s, true, call_context_w(s, true, fr_add_u64(s.cur(S_RWC), 2), 0, s.cur(S_CALL_ID), ZK_CC_LastCalleeReturnDataLength, &v, &w),
             EV_EOOG_CC_UNSAT)) return;
  EV_CHECK(EV_EOOG_CC_TYPE, !w);
  const Fr rdl = v.lo, end = fr_add(data_offset, length);  // both < 2^248: the sum needs no reduction
  const bool off_over = !fr_fits64(data_offset), end_over = !fr_fits64(end);
  const bool fits31 = (rdl.l[3] >> 56) == 0 && (end.l[3] >> 56) == 0;  // compare(.., .., MAX_N_BYTES = 31)
  EV_CHECK(EV_EOOG_CMP_RANGE, fits31);
  EV_CHECK(EV_EOOG_NOT_ENOUGH, off_over || end_over || fr_lt(rdl, end));
  error_state_tail(s, 3);
}

// ---- BALANCE / EXTCODEHASH / EXTCODESIZE (balance.py, extcodehash.py, extcodesize.py) ------------------------------------
// pop an address, add it to the transaction's access list (state_write: the write row and, when the call is not
// persistent, its reversion row at rw_counter_end_of_reversion - reversible_write_counter), read the account, push.
// EXTCODESIZE alone advances reversible_write_counter (extcodesize.py:41), as written.
ZK_HD_NOINLINE void gadget_account_access(const StepCtx& s, u64 op) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_ACC_OPCODE, fr_eq_u64(opcode, op));
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  Word2 addr_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, sp, &addr_w), EV_ACC_POP_UNSAT)) return;
  Fr address = fr_u64(0);
  EOOG_W2FQ(addr_w, 20, &address, EV_ACC_ADDR_DOMAIN);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 1), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_TXID_UNSAT);
  const Fr tx_id = rw_cell(s, R_VAL_LO, r);
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 2), call_id, ZK_CC_RwCounterEndOfReversion, &r), EV_ACC_REVEND_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_REVEND_UNSAT);
  const Fr rev_end = rw_cell(s, R_VAL_LO, r);
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 3), call_id, ZK_CC_IsPersistent, &r), EV_ACC_PERSIST_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_PERSIST_UNSAT);
  const Fr is_persistent = rw_cell(s, R_VAL_LO, r);
  Fr is_warm;
  {
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, 4), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id;
    key[R_ADDR] = address;
    key[R_VAL_LO] = fr_u64(1);
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_VAL_LO) | ZK_RWM(R_VAL_HI), &r), EV_ACC_AL_UNSAT);
    const u32 first = r;
    if (fr_is_zero(is_persistent)) {
      u32 r2 = 0;
      TX_LK(reversion_lookup_m(s, fr_sub(rev_end, s.cur(S_REV)), first, &r2), EV_ACC_AL_REV_UNSAT);
    }
    EV_CHECK(EV_ACC_AL_PREV_TYPE, !rw_flag(s, first, 1));
    is_warm = rw_cell(s, R_PREV_LO, first);
  }
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, 5), 0, address, ZK_ACC_CodeHash, &r), EV_ACC_HASH_UNSAT);
  const Word2 code_hash = rw_word(s, R_VAL_LO, r);
  const bool exists = !fr_is_zero(fr_add(code_hash.lo, code_hash.hi));  // 1 - is_zero(lo + hi) over the field
  Word2 expect{fr_u64(0), fr_u64(0)};
  u64 n_rw = 6, d_rev = 0;
  if (op == 0x31) {
    if (exists) {
      TX_LK(account_lookup_m(s, fr_add_u64(rwc, 6), 0, address, ZK_ACC_Balance, &r), EV_ACC_BAL_UNSAT);
      expect = rw_word(s, R_VAL_LO, r);
      n_rw = 7;
    }
  } else if (op == 0x3f) {
    expect = code_hash;
  } else {
    if (exists) {
      Fr len = fr_u64(0);  // bytecode_length(code_hash): the Header row
      if (!need1(s, true, bytecode_lookup_ni(s, true, code_hash.lo, code_hash.hi, 1, fr_u64(0), 0, &len), EV_ACC_LEN_UNSAT)) return;
      EV_CHECK(EV_ACC_SIZE_WORD, fr_fits128(len));
      expect.lo = len;
    }
    d_rev = 1;
  }
  Word2 w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, n_rw, 1, sp, &w), EV_ACC_PUSH_UNSAT)) return;
  EV_CHECK(EV_ACC_EQ, word_eq(expect, w));
  EV_CHECK(EV_ACC_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  same_context_r_ni(s, opcode, fr_u64(n_rw + 1), fr_u64(1), fr_u64(0), fr_eq_u64(is_warm, 1) ? fr_u64(0) : fr_u64(2500), d_rev);
}
ZK_HD_NOINLINE void gadget_error_oog_account_access(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_ACC_OPCODE, fr_eq_u64(opcode, 0x31) || fr_eq_u64(opcode, 0x3b) || fr_eq_u64(opcode, 0x3f));
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID);
  Word2 addr_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, s.cur(S_SP), &addr_w), EV_ACC_POP_UNSAT)) return;
  Fr address = fr_u64(0);
  EOOG_W2FQ(addr_w, 20, &address, EV_ACC_ADDR_DOMAIN);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 1), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_TXID_UNSAT);
  const Fr tx_id = rw_cell(s, R_VAL_LO, r);
  Fr key[14];  // read_account_to_access_list: state_read(TxAccessListAccount, tx_id, address)
  rw_key_init(key, fr_add_u64(rwc, 2), 0, ZK_TARGET_TxAccessListAccount);
  key[R_ID] = tx_id;
  key[R_ADDR] = address;
  TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR), &r), EV_ACC_AL_UNSAT);
  EV_CHECK(EV_ACC_AL_PREV_TYPE, !rw_flag(s, r, 1));
  const Fr is_warm = rw_cell(s, R_PREV_LO, r);
  oog_finish(s, fr_eq_u64(is_warm, 1) ? 100 : 2600, 0, 3);
}

// ---- CODECOPY / RETURNDATACOPY / EXTCODECOPY (codecopy.py, returndatacopy.py, extcodecopy.py) ----------------------------
// copy_lookup with the source id compared as a Word (a code hash), table.py:776-778
ZK_HD_NOINLINE int copy_lookup_w(const StepCtx& s, const Fr& src_lo, const Fr& src_hi, u64 src_tag, const Fr& dst_id, u64 dst_tag,
                                 const Fr& src_addr, const Fr& src_end, const Fr& dst_addr, const Fr& length, const Fr& rwc,
                                 Fr* rwc_inc) {
  Fr key[11] = {src_lo, src_hi, fr_u64(src_tag), dst_id, fr_u64(0), fr_u64(dst_tag), src_addr, src_end, dst_addr, length, rwc};
  u32 r = 0;
  const int n = lookup_sync<11>(s.t.copy, key, &r, s.mask, true);
  if (n == 1) *rwc_inc = table_cell(s.t.copy.tab, 13, r);
  return n;
}
#define CPY_POP(k, out) EOOG_STACK((k), (k), (out), EV_CPY_POP0_UNSAT + 2 * (k))
// memory_offset_and_length(moff_w, size_w): the size first, the offset only when the size is not zero
#define CPY_OFFLEN(moff_w, size_w, moff, size)                          \
  do {                                                                  \
    EOOG_W2FQ((size_w), 5, (size), EV_CPY_SIZE_DOMAIN);                 \
    *(moff) = fr_u64(0);                                                \
    if (!fr_is_zero(*(size))) EOOG_W2FQ((moff_w), 5, (moff), EV_CPY_MOFF_DOMAIN); \
  } while (0)
#define CPY_GAS(moff, size, next_mem, gas)                                                        \
  do {                                                                                            \
    const int rc_ = copier_gas(s, (moff).l[0], (size).l[0], ZK_GAS_COST_COPY, (next_mem), (gas)); \
    if (rc_) {                                                                                    \
      step_fail(s, EV_CPY_MEMSIZE_RANGE + rc_ - 1);                                               \
      return;                                                                                     \
    }                                                                                             \
  } while (0)

ZK_HD_NOINLINE void gadget_codecopy(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), hlo = s.cur(S_HASH_LO), hhi = s.cur(S_HASH_HI);
  Word2 moff_w{fr_u64(0), fr_u64(0)}, coff_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  CPY_POP(0, &moff_w);
  CPY_POP(1, &coff_w);
  CPY_POP(2, &size_w);
  Fr moff, size, coff = fr_u64(0);
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  EOOG_W2FQ(coff_w, 5, &coff, EV_CPY_OFF_DOMAIN);
  Fr code_size = fr_u64(0);
  if (!need1(s, true, bytecode_lookup_ni(s, true, hlo, hhi, 1, fr_u64(0), 0, &code_size), EV_CPY_LEN_UNSAT)) return;
  Fr next_mem = fr_u64(0), gas = fr_u64(0);
  CPY_GAS(moff, size, &next_mem, &gas);
  Fr rwc_inc = fr_u64(0);
  if (!fr_is_zero(size)) {
    if (!need1(s, true, copy_lookup_w(s, hlo, hhi, ZK_COPY_Bytecode, call_id, ZK_COPY_Memory, coff, code_size, moff, size, fr_add_u64(rwc, 3), &rwc_inc),
               EV_CPY_COPY_UNSAT)) return;
  }
  same_context_x_ni(s, opcode, fr_add_u64(rwc_inc, 3), fr_u64(1), fr_u64(3), true, next_mem, gas);
}
ZK_HD_NOINLINE void gadget_returndatacopy(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID);
  Word2 moff_w{fr_u64(0), fr_u64(0)}, off_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  CPY_POP(0, &moff_w);
  CPY_POP(1, &off_w);
  CPY_POP(2, &size_w);
  const u64 TAGS[3] = {ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataLength, ZK_CC_LastCalleeReturnDataOffset};
  Fr cc[3];
  for (int k = 0; k < 3; k++) {
    Word2 v{fr_u64(0), fr_u64(0)};
    bool w = false;
    if (!need1(s, true, call_context_w(s, true, fr_add_u64(rwc, 3 + k), 0, call_id, TAGS[k], &v, &w), EV_CPY_CC0_UNSAT + 3 * k)) return;
    EV_CHECK(EV_CPY_CC0_UNSAT + 3 * k + 2, !w);
    cc[k] = v.lo;
  }
  Fr off8 = fr_u64(0), size8 = fr_u64(0);
  EOOG_W2FQ(off_w, 8, &off8, EV_CPY_OFF_DOMAIN);
  EOOG_W2FQ(size_w, 8, &size8, EV_CPY_SIZE8_DOMAIN);
  {  // range_check(return_data_length - (offset + size), N_BYTES_MEMORY_WORD_SIZE)
    const Fr d = fr_sub(cc[1], fr_add(off8, size8));
    EV_CHECK(EV_CPY_OOB_RANGE, fr_fits64(d) && (d.l[0] >> 32) == 0);
  }
  Fr moff, size;
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  Fr next_mem = fr_u64(0), gas = fr_u64(0);
  CPY_GAS(moff, size, &next_mem, &gas);
  Fr rwc_inc = fr_u64(0), unused = fr_u64(0);
  if (!need1(s, true, copy_lookup(s, true, cc[0], ZK_COPY_Memory, call_id, ZK_COPY_Memory, cc[2], fr_add(cc[2], size), moff, size,
                                  fr_add_u64(rwc, 6), &rwc_inc, &unused), EV_CPY_COPY_UNSAT)) return;
  EV_CHECK(EV_CPY_RWC_INC, fr_eq(rwc_inc, fr_add(size, size)));
  same_context_x_ni(s, opcode, fr_add_u64(rwc_inc, 6), fr_u64(1), fr_u64(3), true, next_mem, gas);
}
ZK_HD_NOINLINE void gadget_extcodecopy(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID);
  Word2 addr_w{fr_u64(0), fr_u64(0)}, moff_w{fr_u64(0), fr_u64(0)}, coff_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  CPY_POP(0, &addr_w);
  Fr address = fr_u64(0);
  EOOG_W2FQ(addr_w, 20, &address, EV_CPY_ADDR_DOMAIN);
  CPY_POP(1, &moff_w);
  CPY_POP(2, &coff_w);
  CPY_POP(3, &size_w);
  Fr coff = fr_u64(0), moff, size;
  EOOG_W2FQ(coff_w, 8, &coff, EV_CPY_OFF_DOMAIN);
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 4), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_TXID_UNSAT);
  const Fr tx_id = rw_cell(s, R_VAL_LO, r);
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 5), call_id, ZK_CC_RwCounterEndOfReversion, &r), EV_ACC_REVEND_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_REVEND_UNSAT);
  const Fr rev_end = rw_cell(s, R_VAL_LO, r);
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 6), call_id, ZK_CC_IsPersistent, &r), EV_ACC_PERSIST_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_PERSIST_UNSAT);
  const Fr is_persistent = rw_cell(s, R_VAL_LO, r);
  Fr is_warm;
  {
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, 7), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id;
    key[R_ADDR] = address;
    key[R_VAL_LO] = fr_u64(1);
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_VAL_LO) | ZK_RWM(R_VAL_HI), &r), EV_ACC_AL_UNSAT);
    const u32 first = r;
    if (fr_is_zero(is_persistent)) {
      u32 r2 = 0;
      TX_LK(reversion_lookup_m(s, fr_sub(rev_end, s.cur(S_REV)), first, &r2), EV_ACC_AL_REV_UNSAT);
    }
    EV_CHECK(EV_ACC_AL_PREV_TYPE, !rw_flag(s, first, 1));
    is_warm = rw_cell(s, R_PREV_LO, first);
  }
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, 8), 0, address, ZK_ACC_CodeHash, &r), EV_ACC_HASH_UNSAT);
  const Word2 code_hash = rw_word(s, R_VAL_LO, r);
  Fr code_size = fr_u64(0);
  if (!fr_is_zero(fr_add(code_hash.lo, code_hash.hi))) {
    if (!need1(s, true, bytecode_lookup_ni(s, true, code_hash.lo, code_hash.hi, 1, fr_u64(0), 0, &code_size), EV_ACC_LEN_UNSAT)) return;
  }
  Fr next_mem = fr_u64(0), gas = fr_u64(0);
  CPY_GAS(moff, size, &next_mem, &gas);
  EV_CHECK(EV_ACC_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  if (!fr_eq_u64(is_warm, 1)) gas = fr_add_u64(gas, 2500);
  Fr rwc_inc = fr_u64(0);
  if (!fr_is_zero(size)) {
    if (!need1(s, true, copy_lookup_w(s, code_hash.lo, code_hash.hi, ZK_COPY_Bytecode, call_id, ZK_COPY_Memory, coff, code_size, moff, size,
                                      fr_add_u64(rwc, 9), &rwc_inc), EV_CPY_COPY_UNSAT)) return;
  }
  same_context_x_ni(s, opcode, fr_add_u64(rwc_inc, 9), fr_u64(1), fr_u64(4), true, next_mem, gas);
}
// the external address goes through word_to_fq(.., N_BYTES_MEMORY_ADDRESS = 5) (error_oog_memory_copy.py:45), as written
ZK_HD_NOINLINE void gadget_error_oog_memory_copy(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const bool ext = fr_eq_u64(opcode, 0x3c);
  EV_CHECK(EV_CPY_OPCODE, fr_eq_u64(opcode, 0x37) || fr_eq_u64(opcode, 0x39) || ext || fr_eq_u64(opcode, 0x3e));
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID);
  Word2 addr_w{fr_u64(0), fr_u64(0)}, moff_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  const u64 k = ext ? 1 : 0;
  if (ext) CPY_POP(0, &addr_w);
  EOOG_STACK(k, k, &moff_w, EV_CPY_POP0_UNSAT + 2 * k);
  EOOG_STACK(k + 1, k + 2, &size_w, EV_CPY_POP0_UNSAT + 2 * (k + 1));
  u64 n_rw = k + 2, constant = 3;
  if (ext) {
    Fr address = fr_u64(0);
    EOOG_W2FQ(addr_w, 5, &address, EV_CPY_ADDR_DOMAIN);
    u32 r = 0;
    TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 3), call_id, ZK_CC_TxId, &r), EV_ACC_TXID_UNSAT);
    TX_NOT_WORD(rw_flag(s, r, 0), EV_ACC_TXID_UNSAT);
    const Fr tx_id = rw_cell(s, R_VAL_LO, r);
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, 4), 0, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id;
    key[R_ADDR] = address;
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR), &r), EV_ACC_AL_UNSAT);
    EV_CHECK(EV_ACC_AL_PREV_TYPE, !rw_flag(s, r, 1));
    constant = fr_eq_u64(rw_cell(s, R_PREV_LO, r), 1) ? 100 : 2600;
    n_rw = 5;
  }
  Fr moff, size;
  CPY_OFFLEN(moff_w, size_w, &moff, &size);
  Fr next_mem = fr_u64(0), gas = fr_u64(0);
  CPY_GAS(moff, size, &next_mem, &gas);
  const unsigned __int128 cost = (unsigned __int128)constant + gas.l[0];
  oog_finish(s, (u64)cost, (u64)(cost >> 64), n_rw);
}

}  // namespace zk

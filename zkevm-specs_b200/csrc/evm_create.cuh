// Gate program of CREATE / CREATE2 (group KG_TX), part of evm.cu (included there).
//   create   evm_circuit/execution/create.py:20-254 (generate_contract_address / generate_CREAET2_contract_address
//            instruction.py:1338-1352, transfer :1111-1120, add_account_to_access_list :1044-1057, memory_expansion
//            :1138-1155, step_state_transition_to_new_context :266-290)
// StepState.aux_data (the init code's hash, create.py:107) is not one of the 13 step cells: it is looked up in the step-aux
// side table ZK_TABLE_STEP_AUX (step row, lo, hi) by the step's row.  Three endings: pre-check failure (depth, balance,
// nonce) or address collision -> the step stays in the caller with an empty return-data record; creation without init
// code -> the same, after the transfer and the nonce write; creation with init code -> the memory chunk is copied into the
// bytecode next.code_hash, the caller's state is saved, the new call's ten context cells are read back and the next step
// starts the init code.  Reproduced as written: is_zero(is_static) (create.py:52) is computed and dropped; the caller's
// nonce write is not tied to nonce_prev + 1; the access-list write and both nonce writes carry no reversion row.
#pragma once
namespace zk {

// keccak(0xff ++ address (20, big endian) ++ salt (32, little endian) ++ code hash (32, little endian))[12:]; false when a
// word's integer lo + (hi << 128) does not fit 32 bytes (OverflowError in the reference)
ZK_HD_NOINLINE bool contract_address2(const Fr& address, const Word2& salt, const Word2& hash, Fr* out) {
  unsigned char buf[85];
  int n = 0;
  buf[n++] = 0xff;
  for (int k = 19; k >= 0; k--) buf[n++] = (unsigned char)(address.l[k >> 3] >> (8 * (k & 7)));
#pragma unroll 1
  for (int t = 0; t < 2; t++) {
    const Word2& w = t == 0 ? salt : hash;
    if ((w.hi.l[2] | w.hi.l[3]) != 0) return false;
    const Fr hi = fr_add(w.hi, fr_u128(w.lo.l[2], w.lo.l[3]));  // < 2^128 + 2^126: no wrap
    if ((hi.l[2] | hi.l[3]) != 0) return false;
    for (int k = 0; k < 16; k++) buf[n++] = (unsigned char)(w.lo.l[k >> 3] >> (8 * (k & 7)));
    for (int k = 0; k < 16; k++) buf[n++] = (unsigned char)(hi.l[k >> 3] >> (8 * (k & 7)));
  }
  u64 d[4];
  keccak256(buf, (u64)n, d);
  Fr r = fr_u64(0);
  for (int k = 0; k < 20; k++) {
    const int j = 31 - k;
    r.l[k >> 3] |= ((d[j >> 3] >> (8 * (j & 7))) & 0xFF) << (8 * (k & 7));
  }
  *out = r;
  return true;
}

ZK_HD_NOINLINE void gadget_create(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  const bool is_create = fr_eq_u64(opcode, 0xf0), is_create2 = fr_eq_u64(opcode, 0xf5);
  EV_CHECK(EV_CR_RESP_OPCODE, responsible_opcode(s, s.cur(S_STATE), opcode));
  const Fr callee_call_id = rwc;
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 w[5] = {zero, zero, zero, zero, zero};  // value, offset, size, salt, returned address
  u64 k = 0;
#pragma unroll
  for (int f = 0; f < 4; f++) {
    if (f == 3 && !is_create2) continue;
    if (!need1(s, true, stack_at(s, true, k, 0, fr_add_u64(sp, k), &w[f]), EV_CR_POP0_UNSAT + 2 * f)) return;
    k++;
  }
  if (!need1(s, true, stack_at(s, true, k, 1, fr_add_u64(sp, k - 1), &w[4]), EV_CR_PUSH_UNSAT)) return;
  k++;
  Fr offset = fr_u64(0), size = fr_u64(0);
  EOOG_W2FQ(w[1], 5, &offset, EV_CR_OFF_DOMAIN);
  EOOG_W2FQ(w[2], 5, &size, EV_CR_SIZE_DOMAIN);
  Fr depth, tx_id, is_success, is_static, rev_end, is_persistent;
  u32 r = 0;
  CALL_CCV(k, 0, call_id, ZK_CC_Depth, &depth, EV_CR_DEPTH_UNSAT);
  k++;
  CALL_CCV(k, 0, call_id, ZK_CC_TxId, &tx_id, EV_CR_TXID_UNSAT);
  k++;
  TX_LK(cc_rw_lookup_m(s, fr_add_u64(rwc, k), 0, call_id, ZK_CC_CallerAddress, &r), EV_CR_CALLER_UNSAT);
  k++;
  const Word2 caller_w = rw_word(s, R_VAL_LO, r);
  Fr caller = fr_u64(0);
  EOOG_W2FQ(caller_w, 20, &caller, EV_CR_CALLER_DOMAIN);
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 1, caller, ZK_ACC_Nonce, &r), EV_CR_NONCE_UNSAT);
  k++;
  TX_NOT_WORD(rw_flag(s, r, 0), EV_CR_NONCE_UNSAT);
  EV_CHECK(EV_CR_NONCE_PREV_TYPE, !rw_flag(s, r, 1));
  const Fr nonce = rw_cell(s, R_VAL_LO, r), nonce_prev = rw_cell(s, R_PREV_LO, r);
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 0, caller, ZK_ACC_Balance, &r), EV_CR_BAL_UNSAT);
  k++;
  TX_NOT_WORD(rw_flag(s, r, 0), EV_CR_BAL_UNSAT);
  const Fr balance = rw_cell(s, R_VAL_LO, r);
  CALL_CCV(k, 0, call_id, ZK_CC_IsSuccess, &is_success, EV_CR_SUCCESS_UNSAT);
  k++;
  CALL_CCV(k, 0, call_id, ZK_CC_IsStatic, &is_static, EV_CR_STATIC_UNSAT);  // create.py:52: read, never constrained
  k++;
  CALL_CCV(k, 0, call_id, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_CR_REVEND_UNSAT);
  k++;
  CALL_CCV(k, 0, call_id, ZK_CC_IsPersistent, &is_persistent, EV_CR_PERSIST_UNSAT);
  k++;
  const bool has_init_code = !fr_is_zero(size);
  // memory_expansion(offset, size); both below 2^40
  const u64 words = has_init_code ? (offset.l[0] + size.l[0] + 31) / 32 : 0;
  EV_CHECK(EV_CR_MEMSIZE_RANGE, (words >> 32) == 0);
  const Fr cur_mem = s.cur(S_MEM);
  EV_CHECK(EV_CR_MEM_MAX, fr_fits64(cur_mem) && (cur_mem.l[0] >> 32) == 0);
  const u64 next_mem = cur_mem.l[0] < words ? words : cur_mem.l[0];
  const u64 expansion = memory_gas_cost(next_mem) - memory_gas_cost(cur_mem.l[0]);
  const u64 word_len = (size.l[0] + 31) / 32;
  EV_CHECK(EV_CR_WORDLEN_RANGE, (word_len >> 32) == 0);
  const Fr gas_left = s.cur(S_GAS);
  const u64 gas_cost = 32000 + expansion + word_len * 2 + (is_create2 ? 6 * word_len : 0);
  const Fr gas_available = fr_sub(gas_left, fr_u64(gas_cost));
  const Fr one_64th{{(gas_available.l[0] >> 6) | (gas_available.l[1] << 58), (gas_available.l[1] >> 6) | (gas_available.l[2] << 58),
                     (gas_available.l[2] >> 6) | (gas_available.l[3] << 58), gas_available.l[3] >> 6}};
  EV_CHECK(EV_CR_GAS_64TH_RANGE, fr_fits64(one_64th));
  const Fr all_but = fr_sub(gas_available, one_64th);
  Fr callee_gas_left = all_but;
  if (fr_fits64(gas_left)) {  // is_u64_gas: min(all_but_one_64th_gas, gas_left, 8)
    EV_CHECK(EV_CR_GAS_MIN_RANGE, fr_fits64(all_but));
    callee_gas_left = all_but.l[0] < gas_left.l[0] ? all_but : gas_left;
  }
  EV_CHECK(EV_CR_DEPTH_RANGE, fr_fits64(depth) && (depth.l[0] >> 16) == 0);
  const Word2 bal_w{fr_u128(balance.l[0], balance.l[1]), fr_u128(balance.l[2], balance.l[3])};
  EV_CHECK(EV_CR_BAL_CMP_RANGE, word_in_domain(w[0]));
  const bool insufficient = fr_lt(bal_w.hi, w[0].hi) || (fr_eq(bal_w.hi, w[0].hi) && fr_lt(bal_w.lo, w[0].lo));
  EV_CHECK(EV_CR_NONCE_RANGE, fr_fits64(nonce_prev));
  const bool precheck_ok = depth.l[0] < 1025 && !insufficient && nonce_prev.l[0] < 0xFFFFFFFFFFFFFFFFull;
  const u64 sp_delta = 2 + (is_create2 ? 1 : 0);
  bool not_collision = false;
  if (precheck_ok) {
    const Word2 empty{fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
    Word2 code_hash = empty;
    if (has_init_code) {  // curr.aux_data
      const Fr key[1] = {fr_u64(s.row)};
      u32 ra = 0;
      EV_CHECK(EV_CR_AUX_MISSING, lookup<1>(s.t.aux, key, &ra) == 1);
      code_hash.lo = table_cell(s.t.aux.tab, 1, ra);
      code_hash.hi = table_cell(s.t.aux.tab, 2, ra);
    }
    Fr contract = fr_u64(0);
    if (is_create) contract = contract_address(caller, nonce);
    else EV_CHECK(EV_CR_ADDR2_DOMAIN, contract_address2(caller, w[3], code_hash, &contract));
    const Word2 contract_w{fr_u128(contract.l[0], contract.l[1]), fr_u64(contract.l[2])};
    {
      Fr key[14];
      rw_key_init(key, fr_add_u64(rwc, k), 1, ZK_TARGET_TxAccessListAccount);
      key[R_ID] = tx_id;
      key[R_ADDR] = contract;
      key[R_VAL_LO] = fr_u64(1);
      TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_VAL_LO) | ZK_RWM(R_VAL_HI), &r), EV_CR_AL_UNSAT);
      k++;
      EV_CHECK(EV_CR_AL_PREV_TYPE, !rw_flag(s, r, 1));
    }
    TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 0, contract, ZK_ACC_CodeHash, &r), EV_CR_CHASH_UNSAT);
    k++;
    const Word2 callee_hash = rw_word(s, R_VAL_LO, r);
    TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 0, contract, ZK_ACC_Nonce, &r), EV_CR_CNONCE_UNSAT);
    k++;
    TX_NOT_WORD(rw_flag(s, r, 0), EV_CR_CNONCE_UNSAT);
    const Fr callee_nonce = rw_cell(s, R_VAL_LO, r);
    // is_equal_word / is_zero_word: the FIELD SUM of the halves (differences) is zero
    const bool is_empty_hash = fr_is_zero(fr_add(fr_sub(callee_hash.lo, empty.lo), fr_sub(callee_hash.hi, empty.hi)));
    const bool is_zero_hash = fr_is_zero(fr_add(callee_hash.lo, callee_hash.hi));
    not_collision = fr_is_zero(callee_nonce) && (is_empty_hash || is_zero_hash);
    if (not_collision) {
      Fr ret = fr_u64(0);
      EOOG_W2FQ(w[4], 20, &ret, EV_CR_RETURN_DOMAIN);
      EV_CHECK(EV_CR_RETURN_EQ, fr_eq(ret, fr_mul_sel(contract, is_success)));
      Fr callee_rev_end, callee_persistent;
      CALL_CCV(k, 0, callee_call_id, ZK_CC_RwCounterEndOfReversion, &callee_rev_end, EV_CR_CREVEND_UNSAT);
      k++;
      CALL_CCV(k, 0, callee_call_id, ZK_CC_IsPersistent, &callee_persistent, EV_CR_CPERSIST_UNSAT);
      k++;
      EV_CHECK(EV_CR_CPERSIST_EQ, fr_eq(callee_persistent, fr_mul_sel(is_persistent, is_success)));
      // transfer(caller, contract, value, callee_reversion_info)
      if (!balance_write(s, fr_add_u64(rwc, k), caller, callee_persistent, callee_rev_end, EV_CR_SEND_UNSAT, &r)) return;
      k++;
      {
        const Word2 ws[2] = {rw_word(s, R_VAL_LO, r), w[0]};
        Fr carry;
        const Word2 sum = add_words_n(ws, 2, &carry);
        EV_CHECK(EV_CR_SEND_EQ, word_eq(rw_word(s, R_PREV_LO, r), sum));
        EV_CHECK(EV_CR_SEND_CARRY, fr_is_zero(carry));
      }
      if (!balance_write(s, fr_add_u64(rwc, k), contract, callee_persistent, fr_sub(callee_rev_end, fr_u64(1)), EV_CR_RECV_UNSAT, &r)) return;
      k++;
      {
        const Word2 ws[2] = {rw_word(s, R_PREV_LO, r), w[0]};
        Fr carry;
        const Word2 sum = add_words_n(ws, 2, &carry);
        EV_CHECK(EV_CR_RECV_EQ, word_eq(rw_word(s, R_VAL_LO, r), sum));
        EV_CHECK(EV_CR_RECV_CARRY, fr_is_zero(carry));
      }
      TX_LK(account_lookup_m(s, fr_add_u64(rwc, k), 1, contract, ZK_ACC_Nonce, &r), EV_CR_NEWNONCE_UNSAT);
      k++;
      TX_NOT_WORD(rw_flag(s, r, 0), EV_CR_NEWNONCE_UNSAT);
      EV_CHECK(EV_CR_NEWNONCE_PREV_TYPE, !rw_flag(s, r, 1));
      EV_CHECK(EV_CR_NEWNONCE_EQ, fr_eq_u64(rw_cell(s, R_VAL_LO, r), 1));
      if (has_init_code) {
        const Word2 next_hash{s.nxt(S_HASH_LO), s.nxt(S_HASH_HI)};
        Fr inc = fr_u64(0);
        if (!need1(s, true, copy_lookup_dw(s, call_id, ZK_COPY_Memory, next_hash, ZK_COPY_Bytecode, offset, fr_add(offset, size), fr_u64(0),
                                           size, fr_add_u64(rwc, k), &inc), EV_CR_COPY_UNSAT)) return;
        const Fr base = fr_add(rwc, inc);  // rw_counter + the copy's increment (a field element)
        Fr code_size = fr_u64(0);
        if (!need1(s, true, bytecode_lookup_ni(s, true, next_hash.lo, next_hash.hi, 1, fr_u64(0), 0, &code_size), EV_CR_CODE_LEN_UNSAT)) return;
        EV_CHECK(EV_CR_CODE_LEN_EQ, fr_eq(code_size, size));
        {  // save the caller's state: 5 call-context writes on the current call
          const u64 TAGS[5] = {ZK_CC_ProgramCounter, ZK_CC_StackPointer, ZK_CC_GasLeft, ZK_CC_MemorySize, ZK_CC_ReversibleWriteCounter};
#pragma unroll 1
          for (int t = 0; t < 5; t++) {
            const Fr want = t == 0   ? fr_add_u64(s.cur(S_PC), 1)
                            : t == 1 ? fr_add_u64(sp, sp_delta)
                            : t == 2 ? fr_sub(fr_sub(gas_left, fr_u64(gas_cost)), callee_gas_left)
                            : t == 3 ? fr_u64(next_mem)
                                     : fr_add_u64(s.cur(S_REV), 1);
            u32 r_ = 0;
            TX_LK(cc_rw_lookup_m(s, fr_add_u64(base, k), 1, call_id, TAGS[t], &r_), EV_CR_SAVE0_UNSAT + 4 * t);
            k++;
            TX_NOT_WORD(rw_flag(s, r_, 0), EV_CR_SAVE0_UNSAT + 4 * t);
            EV_CHECK(EV_CR_SAVE0_UNSAT + 4 * t + 3, fr_eq(rw_cell(s, R_VAL_LO, r_), want));
          }
        }
        {  // the new call's context: 10 call-context reads compared as words (lo, hi)
          const u64 TAGS[10] = {ZK_CC_CallerId, ZK_CC_TxId, ZK_CC_Depth, ZK_CC_CallerAddress, ZK_CC_CalleeAddress, ZK_CC_IsSuccess,
                                ZK_CC_IsStatic, ZK_CC_IsRoot, ZK_CC_IsCreate, ZK_CC_CodeHash};
#pragma unroll 1
          for (int t = 0; t < 10; t++) {
            Word2 want = zero;
            switch (t) {
              case 0: want.lo = call_id; break;
              case 1: want.lo = tx_id; break;
              case 2: want.lo = fr_add_u64(depth, 1); break;
              case 3: want = caller_w; break;
              case 4: want = contract_w; break;
              case 5: want.lo = is_success; break;
              case 8: want.lo = fr_u64(1); break;
              case 9: want = code_hash; break;
              default: break;
            }
            TX_LK(cc_rw_lookup_m(s, fr_add_u64(base, k), 0, callee_call_id, TAGS[t], &r), EV_CR_CTX0_UNSAT + 3 * t);
            k++;
            EV_CHECK(EV_CR_CTX0_UNSAT + 3 * t + 2, word_eq(rw_word(s, R_VAL_LO, r), want));
          }
        }
        EV_CHECK(EV_CR_NC_RWC, fr_eq(s.nxt(S_RWC), fr_add_u64(base, k)));
        EV_CHECK(EV_CR_NC_CALL_ID, fr_eq(s.nxt(S_CALL_ID), callee_call_id));
        EV_CHECK(EV_CR_NC_IS_ROOT, fr_is_zero(s.nxt(S_IS_ROOT)));
        EV_CHECK(EV_CR_NC_IS_CREATE, fr_eq_u64(s.nxt(S_IS_CREATE), 1));
        EV_CHECK(EV_CR_NC_GAS, fr_eq(s.nxt(S_GAS), callee_gas_left));
        EV_CHECK(EV_CR_NC_REV, fr_eq_u64(s.nxt(S_REV), 3));
        EV_CHECK(EV_CR_NC_LOG, fr_eq(s.nxt(S_LOG), s.cur(S_LOG)));
        EV_CHECK(EV_CR_NC_PC, fr_is_zero(s.nxt(S_PC)));
        EV_CHECK(EV_CR_NC_SP, fr_eq_u64(s.nxt(S_SP), 1024));
        EV_CHECK(EV_CR_NC_MEM, fr_is_zero(s.nxt(S_MEM)));
        return;
      }
    }
  }
  // pre-check failure, address collision, or nothing to run: the step stays in the caller's context
  if (!precheck_ok || !not_collision) EV_CHECK(EV_CR_FAIL_SUCCESS, fr_is_zero(is_success));
  {
    const u64 TAGS[3] = {ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataOffset, ZK_CC_LastCalleeReturnDataLength};
#pragma unroll 1
    for (int t = 0; t < 3; t++) {
      Fr v;
      CALL_CCV(k, 1, call_id, TAGS[t], &v, EV_CR_LAST0_UNSAT + 4 * t);
      k++;
      EV_CHECK(EV_CR_LAST0_UNSAT + 4 * t + 3, fr_is_zero(v));
    }
  }
  EV_CHECK(EV_CR_SAME_RWC, fr_eq(s.nxt(S_RWC), fr_add_u64(rwc, k)));
  EV_CHECK(EV_CR_SAME_PC, fr_eq(s.nxt(S_PC), fr_add_u64(s.cur(S_PC), 1)));
  EV_CHECK(EV_CR_SAME_SP, fr_eq(s.nxt(S_SP), fr_add_u64(sp, sp_delta)));
  EV_CHECK(EV_CR_SAME_REV, fr_eq(s.nxt(S_REV), fr_add_u64(s.cur(S_REV), not_collision ? 3 : 0)));
  EV_CHECK(EV_CR_SAME_GAS, fr_eq(s.nxt(S_GAS), fr_sub(gas_left, fr_u64(gas_cost))));
  EV_CHECK(EV_CR_SAME_MEM, fr_eq_u64(s.nxt(S_MEM), next_mem));
  EV_CHECK(EV_CR_SAME_CALL_ID, fr_eq(s.nxt(S_CALL_ID), call_id));
  EV_CHECK(EV_CR_SAME_IS_ROOT, fr_eq(s.nxt(S_IS_ROOT), s.cur(S_IS_ROOT)));
  EV_CHECK(EV_CR_SAME_IS_CREATE, fr_eq(s.nxt(S_IS_CREATE), s.cur(S_IS_CREATE)));
  EV_CHECK(EV_CR_SAME_CODE_HASH, fr_eq(s.nxt(S_HASH_LO), s.cur(S_HASH_LO)) && fr_eq(s.nxt(S_HASH_HI), s.cur(S_HASH_HI)));
}

// ErrorOutOfGasSloadSstore: error_oog_sload_sstore.py:16-60 (read_account_storage_to_access_list instruction.py:1088-1097,
// account_storage_read :1015-1026, constrain_error_state).  StepState.aux_data is the slot's committed value as an INT
// (Word(aux_data) splits it at bit 128), taken from ZK_TABLE_STEP_AUX by the step's row.
ZK_HD_NOINLINE void gadget_error_oog_sload_sstore(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  const bool is_sstore = fr_eq_u64(opcode, 0x55), is_sload = fr_eq_u64(opcode, 0x54);
  EV_CHECK(EV_ESS_OPCODE, is_sstore || is_sload);
  Word2 key_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, sp, &key_w), EV_ESS_KEY_UNSAT)) return;
  Fr tx_id, callee = fr_u64(0);
  ST_CC(1, ZK_CC_TxId, &tx_id, EV_ESS_TXID_UNSAT);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 2), call_id, ZK_CC_CalleeAddress, &r), EV_ESS_CALLEE_UNSAT);
  EOOG_W2FQ(rw_word(s, R_VAL_LO, r), 20, &callee, EV_ESS_CALLEE_DOMAIN);
  {
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, 3), 0, ZK_TARGET_TxAccessListAccountStorage);
    key[R_ID] = tx_id;
    key[R_ADDR] = callee;
    key[R_KEY_LO] = key_w.lo;
    key[R_KEY_HI] = key_w.hi;
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_KEY_LO) | ZK_RWM(R_KEY_HI), &r), EV_ESS_AL_UNSAT);
    TX_NOT_WORD(rw_flag(s, r, 0), EV_ESS_AL_UNSAT);
  }
  const Fr is_warm = rw_cell(s, R_VAL_LO, r);
  u64 gas_cost = 0, n_rw = 4;
  if (is_sload) gas_cost = fr_eq_u64(is_warm, 1) ? 100 : 2100;
  else {
    Word2 value{fr_u64(0), fr_u64(0)};
    if (!need1(s, true, stack_at(s, true, 4, 0, fr_add_u64(sp, 1), &value), EV_ESS_VAL_UNSAT)) return;
    {
      Fr key[14];
      rw_key_init(key, fr_add_u64(rwc, 5), 0, ZK_TARGET_AccountStorage);
      key[R_ID] = tx_id;
      key[R_ADDR] = callee;
      key[R_KEY_LO] = key_w.lo;
      key[R_KEY_HI] = key_w.hi;
      TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_KEY_LO) | ZK_RWM(R_KEY_HI), &r), EV_ESS_READ_UNSAT);
    }
    const Word2 value_prev = rw_word(s, R_VAL_LO, r);
    n_rw = 6;
    const Fr akey[1] = {fr_u64(s.row)};
    u32 ra = 0;
    EV_CHECK(EV_ESS_AUX_MISSING, lookup<1>(s.t.aux, akey, &ra) == 1);
    Word2 original{table_cell(s.t.aux.tab, 1, ra), table_cell(s.t.aux.tab, 2, ra)};
    // Word(lo + (hi << 128)): the integer must fit 32 bytes, and is split again at bit 128
    EV_CHECK(EV_ESS_AUX_RANGE, (original.hi.l[2] | original.hi.l[3]) == 0);
    original.hi = fr_add(original.hi, fr_u128(original.lo.l[2], original.lo.l[3]));  // < 2^128 + 2^126: no wrap
    original.lo = fr_u128(original.lo.l[0], original.lo.l[1]);
    EV_CHECK(EV_ESS_AUX_RANGE, (original.hi.l[2] | original.hi.l[3]) == 0);
    if (word_eq(value, value_prev)) gas_cost = 100;
    else if (word_eq(value_prev, original)) gas_cost = (fr_is_zero(original.lo) && fr_is_zero(original.hi)) ? 20000 : 2900;
    else gas_cost = 100;
    if (fr_is_zero(is_warm)) gas_cost += 2100;
  }
  const Fr gas_left = s.cur(S_GAS);
  EV_CHECK(EV_ESS_GAS_RANGE, fr_fits64(gas_left));
  const bool insufficient = gas_left.l[0] < gas_cost;
  if (is_sload) EV_CHECK(EV_ESS_SLOAD_NOT_OOG, insufficient);
  else EV_CHECK(EV_ESS_SSTORE_NOT_OOG, insufficient || gas_left.l[0] <= 2300);
  error_state_tail(s, n_rw);
}

// ErrorOutOfGasCREATE: error_oog_create.py:19-65.  In a root call the init code is priced as tx call data, one
// tx_calldata_lookup per byte (`for idx in range(size)`): the walk ends at the first index the tx table does not hold, so its
// trip count is bounded by the table, not by `size`.
ZK_HD_NOINLINE void gadget_error_oog_create(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  const bool is_create = fr_eq_u64(opcode, 0xf0), is_create2 = fr_eq_u64(opcode, 0xf5);
  EV_CHECK(EV_EOCR_OPCODE, is_create || is_create2);
  Word2 off_w{fr_u64(0), fr_u64(0)}, size_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, fr_add_u64(sp, 1), &off_w), EV_EOCR_OFF_UNSAT)) return;
  if (!need1(s, true, stack_at(s, true, 1, 0, fr_add_u64(sp, 2), &size_w), EV_EOCR_SIZE_UNSAT)) return;
  Fr offset = fr_u64(0), size = fr_u64(0), is_root;
  EOOG_W2FQ(size_w, 5, &size, EV_EOCR_SIZE_DOMAIN);
  if (!fr_is_zero(size)) EOOG_W2FQ(off_w, 5, &offset, EV_EOCR_OFF_DOMAIN);
  ST_CC(2, ZK_CC_IsRoot, &is_root, EV_EOCR_ROOT_UNSAT);
  u64 gas_cost = 0, n_rw = 3;
  if (fr_eq_u64(is_root, 1)) {
    Fr tx_id;
    ST_CC(3, ZK_CC_TxId, &tx_id, EV_EOCR_TXID_UNSAT);
    n_rw = 4;
    u64 nz = 0;
#pragma unroll 1
    for (u64 idx = 0; idx < size.l[0]; idx++) {
      u32 r = 0;
      Fr key[3] = {tx_id, fr_u64(ZK_TX_CallData), fr_u64(idx)};
      TX_LK(lookup<3>(s.t.tx, key, &r), EV_EOCR_BYTE_UNSAT);
      TX_NOT_WORD(tx_is_word(s, r), EV_EOCR_BYTE_UNSAT);
      nz += fr_is_zero(table_cell(s.t.tx.tab, 3, r)) ? 0 : 1;
    }
    gas_cost = 53000 + 16 * nz + 4 * (size.l[0] - nz);
  } else {
    u64 expansion = 0;
    const int rc_ = mem_expansion_gas(s, fr_is_zero(size) ? 0 : (offset.l[0] + size.l[0] + 31) / 32, &expansion);
    if (rc_) {
      step_fail(s, rc_ == 1 ? EV_EOCR_MEMSIZE_RANGE : EV_EOCR_MEM_MAX);
      return;
    }
    gas_cost = 32000 + expansion;
  }
  const u64 word_size = (size.l[0] + 31) / 32;
  EV_CHECK(EV_EOCR_WORDSIZE_RANGE, (word_size >> 32) == 0);
  gas_cost += 2 * word_size + (is_create2 ? 6 * word_size : 0);
  const bool exceeds = 49152 < size.l[0];
  const Fr gas_left = s.cur(S_GAS);
  EV_CHECK(EV_EOCR_GAS_RANGE, fr_fits64(gas_left));
  EV_CHECK(EV_EOCR_NOT_OOG, gas_left.l[0] < gas_cost || exceeds);
  error_state_tail(s, n_rw);
}

// ErrorOutOfGasPrecompile: execution/precompiles/error_oog_precompile.py:9-35 (no opcode lookup).  As written: only DATACOPY
// and BN254PAIRING can verify - for the other seven precompiles gas_cost stays a Python int and compare() raises
// AttributeError on it (ZK_ERR_VALUE here); the pairing count is the FIELD quotient len / 192, below 2^64 only when 192
// divides the integer.
ZK_HD_NOINLINE void gadget_error_oog_precompile(const StepCtx& s) {
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, rwc, call_id, ZK_CC_CalleeAddress, &r), EV_EOPC_CALLEE_UNSAT);
  Fr address = fr_u64(0), len;
  EOOG_W2FQ(rw_word(s, R_VAL_LO, r), 20, &address, EV_EOPC_CALLEE_DOMAIN);
  ST_CC(1, ZK_CC_CallDataLength, &len, EV_EOPC_CDLEN_UNSAT);
  EV_CHECK(EV_EOPC_NOT_PRECOMPILE, fr_fits64(address) && address.l[0] >= 1 && address.l[0] <= 9);
  const u64 a = address.l[0];
  u64 gas_cost = 0;
  bool cost_ok = true, cost_is_int = false;
  if (a == 8) {
    // 2^64 = 64 (mod 192): fold the four limbs with 64-bit arithmetic
    u64 rem = 0;
#pragma unroll
    for (int k = 3; k >= 0; k--) rem = ((rem * 64) % 192 + len.l[k] % 192) % 192;
    const u64 pairs = len.l[0] / 192;
    if (rem != 0 || !fr_fits64(len) || pairs > (0xFFFFFFFFFFFFFFFFull - 45000) / 34000) cost_ok = false;
    else gas_cost = 45000 + 34000 * pairs;
  } else if (a == 4) {
    // memory_copier_gas_cost(len, 0, 3): (len + 31) // 32 on the field sum, range-checked to 4 bytes
    const Fr t = fr_add_u64(len, 31);
    const Fr q{{(t.l[0] >> 5) | (t.l[1] << 59), (t.l[1] >> 5) | (t.l[2] << 59), (t.l[2] >> 5) | (t.l[3] << 59), t.l[3] >> 5}};
    EV_CHECK(EV_EOPC_WORDSIZE_RANGE, fr_fits64(q) && (q.l[0] >> 32) == 0);
    gas_cost = 15 + 3 * q.l[0];
  } else {
    cost_is_int = true;
  }
  const Fr gas_left = s.cur(S_GAS);
  EV_CHECK(EV_EOPC_GAS_LEFT_RANGE, fr_fits64(gas_left));
  EV_CHECK(EV_EOPC_GAS_INT, !cost_is_int);
  EV_CHECK(EV_EOPC_GAS_COST_RANGE, cost_ok);
  EV_CHECK(EV_EOPC_NOT_OOG, gas_left.l[0] < gas_cost);
  error_state_tail(s, 2);
}

// ErrorGasUintOverflow: error_gas_uint_overflow.py:19-171 with instruction.memory_size :1198-1305, calc_mem_size64 /
// _with_uint :1309-1327, safe_mul :1329-1331, to_word_size :1333-1336.  As written: `if is_dynamic_gas:` tests an FQ object
// (always true), so memory_size runs for every opcode and returns None (TypeError, ZK_ERR_VALUE) for one without a memory
// operand; an offset below 2^64 is read again with word_to_fq(.., 5) and raises beyond 5 bytes; `val.n < offset64.n` never
// holds over the field.  A memory size is offset (< 2^40) + length (< 2^64): 64 bits and a carry.
struct MemSize {
  u64 lo;
  bool carry, overflow;
};
// calc_mem_size64_with_uint: 0 ok, else the failing constraint id
ZK_HD_NOINLINE int cms_uint(const Word2& off_w, u64 len64, MemSize* m) {
  m->lo = 0;
  m->carry = m->overflow = false;
  if (len64 == 0) return 0;
  Fr off = fr_u64(0);
  const int rc = word_to_fq_n(off_w, 31, &off);
  if (rc) return rc == 1 ? EV_EGUO_OFF_DOMAIN : EV_EGUO_OFF_RANGE;
  if (!fr_fits64(off)) {
    m->overflow = true;
    return 0;
  }
  if ((off.l[0] >> 40) != 0) return EV_EGUO_OFF5_RANGE;  // word_to_fq(offset, 5)
  m->lo = off.l[0] + len64;
  m->carry = m->lo < len64;
  return 0;
}
ZK_HD_NOINLINE int cms(const Word2& off_w, const Word2& len_w, MemSize* m) {
  m->lo = 0;
  m->carry = m->overflow = false;
  Fr len = fr_u64(0);
  const int rc = word_to_fq_n(len_w, 31, &len);
  if (rc) return rc == 1 ? EV_EGUO_LEN_DOMAIN : EV_EGUO_LEN_RANGE;
  if (!fr_fits64(len)) {
    m->overflow = true;
    return 0;
  }
  return cms_uint(off_w, len.l[0], m);
}
ZK_HD_NOINLINE void gadget_error_gas_uint_overflow(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  const u64 op = (fr_fits64(opcode) && opcode.l[0] < 256) ? opcode.l[0] : 0x100;
  const bool is_create = op == 0xf0 || op == 0xf5;
  Fr cd_len, tx_id, is_root;
  ST_CC(0, ZK_CC_CallDataLength, &cd_len, EV_EGUO_CDLEN_UNSAT);
  ST_CC(1, ZK_CC_TxId, &tx_id, EV_EGUO_TXID_UNSAT);
  ST_CC(2, ZK_CC_IsRoot, &is_root, EV_EGUO_ROOT_UNSAT);
  bool calldata_of = false, initcode_of = false;
  if (fr_eq_u64(is_root, 1)) {
    // the walk ends at the first byte the tx table does not hold: its trip count is bounded by the table (< 2^32 rows),
    // so the intrinsic gas below stays far from 2^64 and plain 64-bit arithmetic is exact
    const u64 len = fr_fits64(cd_len) ? cd_len.l[0] : ~0ull;
    u64 nz = 0;
#pragma unroll 1
    for (u64 idx = 0; idx < len; idx++) {
      u32 r = 0;
      Fr key[3] = {tx_id, fr_u64(ZK_TX_CallData), fr_u64(idx)};
      TX_LK(lookup<3>(s.t.tx, key, &r), EV_EGUO_BYTE_UNSAT);
      TX_NOT_WORD(tx_is_word(s, r), EV_EGUO_BYTE_UNSAT);
      nz += fr_is_zero(table_cell(s.t.tx.tab, 3, r)) ? 0 : 1;
    }
    if (len > 0) {
      const u64 MAXU = ~0ull;
      u64 gas = is_create ? 53000 : 21000;
      const bool nz_of = (MAXU - gas) / 16 < nz;
      gas += nz * 16;
      bool z_of = false;
      if (!nz_of) {
        const u64 z = len - nz;
        z_of = (MAXU - gas) / 4 < z;
        gas += z * 4;
      }
      if (is_create) initcode_of = (MAXU - gas) / 2 < len / 32 + ((len % 32) ? 1 : 0);
      calldata_of = nz_of || z_of;
    }
  }
  // memory_size(opcode): the pops of each opcode, then calc_mem_size64 on (offset, length)
  int n_pop = 0, io = 0, il = 0, call = 0;
  bool mem32 = false;
  switch (op) {
    case 0x20: case 0xf3: case 0xfd: case 0xa0: case 0xa1: case 0xa2: case 0xa3: case 0xa4: n_pop = 2; io = 0; il = 1; break;
    case 0x37: case 0x3e: case 0x39: n_pop = 3; io = 1; il = 2; break;
    case 0x3c: n_pop = 4; io = 2; il = 3; break;
    case 0x51: n_pop = 1; io = 0; mem32 = true; break;
    case 0x52: case 0x53: n_pop = 2; io = 0; mem32 = true; break;
    case 0xf0: n_pop = 3; io = 1; il = 2; break;
    case 0xf5: n_pop = 4; io = 1; il = 2; break;
    case 0xf1: case 0xf2: n_pop = 7; call = 3; break;
    case 0xf4: case 0xfa: n_pop = 6; call = 2; break;
    default: step_fail(s, EV_EGUO_OPCODE); return;
  }
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 w[7] = {zero, zero, zero, zero, zero, zero, zero};
#pragma unroll 1
  for (int k = 0; k < n_pop; k++)
    if (!need1(s, true, stack_at(s, true, 3 + (u64)k, 0, fr_add_u64(sp, (u64)k), &w[k]), EV_EGUO_POP0_UNSAT + 2 * k)) return;
  MemSize m{0, false, false};
  int id_ = 0;
  if (call) {
    MemSize x, y;
    if ((id_ = cms(w[call + 2], w[call + 3], &x))) { step_fail(s, id_); return; }
    if (x.overflow) m.overflow = true;
    else {
      if ((id_ = cms(w[call], w[call + 1], &y))) { step_fail(s, id_); return; }
      if (y.overflow) m.overflow = true;
      else m = (x.carry != y.carry ? x.carry : x.lo > y.lo) ? x : y;
    }
  } else if (mem32) {
    if ((id_ = cms_uint(w[io], 32, &m))) { step_fail(s, id_); return; }
  } else {
    if ((id_ = cms(w[io], w[il], &m))) { step_fail(s, id_); return; }
  }
  const bool mul_of = !m.overflow && (m.carry || m.lo > ~0ull - 31);  // to_word_size(size) * 32 passes 2^64 - 1
  EV_CHECK(EV_EGUO_NOT_OVERFLOW, m.overflow || mul_of || calldata_of || initcode_of);
  error_state_tail(s, 3 + (u64)n_pop);
}

}  // namespace zk

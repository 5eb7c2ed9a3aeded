/* TEST INFRASTRUCTURE — CPU oracle, part of oracle/evm.c (included there).
 * CREATE / CREATE2: evm_circuit/execution/create.py:20-254 (generate_contract_address / generate_CREAET2_contract_address
 * instruction.py:1338-1352, transfer :1111-1120, add_account_to_access_list :1044-1057, memory_expansion :1138-1155,
 * step_state_transition_to_new_context :266-290).  StepState.aux_data (the init code's hash, create.py:107) is not one of
 * the 13 step cells: it comes from the step-aux side table (step row, lo, hi) set with orc_set_evm_step_aux.
 * Reproduced as written: `instruction.is_zero(is_static)` (create.py:52) is computed and dropped, so a CREATE inside a
 * STATICCALL is not rejected here; the caller's nonce write is not tied to nonce_prev + 1; the access-list write and both
 * nonce writes carry no reversion row.
 * Pinned by tests/golden/evm24.npz (verdicts of the reference's verify_step on 52 cases of its own test data).
 */
/* keccak(0xff ++ address (20, big endian) ++ salt (32, little endian) ++ code hash (32, little endian))[12:], instruction.py:1342-1352;
 * returns 0 when a word does not fit 32 bytes (OverflowError in the reference) */
static int contract_address2(fr_t address, word_t salt, word_t hash, fr_t* out) {
  if (!word_in_domain(salt) || !word_in_domain(hash)) {
    /* lo + (hi << 128) >= 2^256 needs hi >= 2^128, or hi == 2^128 - 1 with lo >= 2^128 carrying in */
    const word_t ws[2] = {salt, hash};
    for (int t = 0; t < 2; t++) {
      if (!fr_fits_bits(ws[t].hi, 128)) return 0;
      if (!fr_fits_bits(fr_add(ws[t].hi, fr_hi128(ws[t].lo)), 128)) return 0; /* < 2^128 + 2^126: no wrap */
    }
  }
  uint8_t buf[85]; int n = 0;
  buf[n++] = 0xff;
  for (int k = 19; k >= 0; k--) buf[n++] = (uint8_t)(address.l[k >> 3] >> (8 * (k & 7)));
  const word_t ws[2] = {salt, hash};
  for (int t = 0; t < 2; t++) {
    const fr_t hi = fr_add(ws[t].hi, fr_hi128(ws[t].lo));
    for (int k = 0; k < 16; k++) buf[n++] = (uint8_t)(ws[t].lo.l[k >> 3] >> (8 * (k & 7)));
    for (int k = 0; k < 16; k++) buf[n++] = (uint8_t)(hi.l[k >> 3] >> (8 * (k & 7)));
  }
  uint8_t h[32]; keccak256_small(buf, n, h);
  fr_t r = fr_u64(0);
  for (int k = 0; k < 20; k++) r.l[k >> 3] |= (uint64_t)h[31 - k] << (8 * (k & 7));
  *out = r;
  return 1;
}

static void gadget_create(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  const int is_create = fr_eq_u64(opcode, 0xf0), is_create2 = fr_eq_u64(opcode, 0xf5);
  {
    fr_t key[4] = {fr_u64(ZK_FIXED_ResponsibleOpcode), CUR(S_STATE), opcode, fr_u64(0)};
    CHECK(EV_CR_RESP_OPCODE, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  }
  const fr_t callee_call_id = rwc;
  const word_t zero = {fr_u64(0), fr_u64(0)};
  word_t w[5]; /* value, offset, size, salt, returned address */
  uint64_t k = 0;
  for (int f = 0; f < 4; f++) {
    if (f == 3 && !is_create2) { w[3] = zero; continue; }
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(k)), &w[f]), EV_CR_POP0_UNSAT + 2 * f, row)) return;
    k++;
  }
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 1, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(k - 1)), &w[4]), EV_CR_PUSH_UNSAT, row)) return;
  k++;
  fr_t offset, size;
  W2FQ(w[1], 5, &offset, EV_CR_OFF_DOMAIN);
  W2FQ(w[2], 5, &size, EV_CR_SIZE_DOMAIN);
  fr_t depth, tx_id, is_success, is_static, rev_end, is_persistent;
  uint32_t r;
  CALL_CCV(k, 0, call_id, ZK_CC_Depth, &depth, EV_CR_DEPTH_UNSAT); k++;
  CALL_CCV(k, 0, call_id, ZK_CC_TxId, &tx_id, EV_CR_TXID_UNSAT); k++;
  LK(cc_rw_lookup(e, fr_add(rwc, fr_u64(k)), 0, call_id, ZK_CC_CallerAddress, &r), EV_CR_CALLER_UNSAT); k++;
  const word_t caller_w = rw_value(e, r);
  fr_t caller;
  W2FQ(caller_w, 20, &caller, EV_CR_CALLER_DOMAIN);
  LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 1, caller, ZK_ACC_Nonce, &r), EV_CR_NONCE_UNSAT); k++;
  NOT_WORD(rw_val_is_word(e, r), EV_CR_NONCE_UNSAT);
  CHECK(EV_CR_NONCE_PREV_TYPE, !rw_prev_is_word(e, r));
  const fr_t nonce = rw_cell(e, R_VAL_LO, r), nonce_prev = rw_cell(e, R_PREV_LO, r);
  LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 0, caller, ZK_ACC_Balance, &r), EV_CR_BAL_UNSAT); k++;
  NOT_WORD(rw_val_is_word(e, r), EV_CR_BAL_UNSAT);
  const fr_t balance = rw_cell(e, R_VAL_LO, r);
  CALL_CCV(k, 0, call_id, ZK_CC_IsSuccess, &is_success, EV_CR_SUCCESS_UNSAT); k++;
  CALL_CCV(k, 0, call_id, ZK_CC_IsStatic, &is_static, EV_CR_STATIC_UNSAT); k++;
  (void)is_static; /* create.py:52 computes is_zero(is_static) and drops it */
  CALL_CCV(k, 0, call_id, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_CR_REVEND_UNSAT); k++;
  (void)rev_end;
  CALL_CCV(k, 0, call_id, ZK_CC_IsPersistent, &is_persistent, EV_CR_PERSIST_UNSAT); k++;
  const int has_init_code = !fr_is_zero(size);
  /* memory_expansion(offset, size) */
  const uint64_t words = has_init_code ? (offset.l[0] + size.l[0] + 31) / 32 : 0; /* both below 2^40 */
  CHECK(EV_CR_MEMSIZE_RANGE, !(words >> 32));
  const fr_t cur_mem = CUR(S_MEM);
  CHECK(EV_CR_MEM_MAX, fr_fits_bits(cur_mem, 32));
  const uint64_t next_mem = cur_mem.l[0] < words ? words : cur_mem.l[0];
  const uint64_t expansion = memory_gas_cost(next_mem) - memory_gas_cost(cur_mem.l[0]);
  const uint64_t word_len = (size.l[0] + 31) / 32;
  CHECK(EV_CR_WORDLEN_RANGE, !(word_len >> 32));
  const fr_t gas_left = CUR(S_GAS);
  const uint64_t gas_cost = 32000 + expansion + word_len * 2 + (is_create2 ? 6 * word_len : 0);
  const fr_t gas_available = fr_sub(gas_left, fr_u64(gas_cost));
  fr_t one_64th = {{(gas_available.l[0] >> 6) | (gas_available.l[1] << 58), (gas_available.l[1] >> 6) | (gas_available.l[2] << 58),
                    (gas_available.l[2] >> 6) | (gas_available.l[3] << 58), gas_available.l[3] >> 6}};
  CHECK(EV_CR_GAS_64TH_RANGE, fr_fits_bits(one_64th, 64));
  const fr_t all_but = fr_sub(gas_available, one_64th);
  fr_t callee_gas_left = all_but;
  if (fr_fits_bits(gas_left, 64)) { /* is_u64_gas: min(all_but_one_64th_gas, gas_left, 8) */
    CHECK(EV_CR_GAS_MIN_RANGE, fr_fits_bits(all_but, 64));
    callee_gas_left = all_but.l[0] < gas_left.l[0] ? all_but : gas_left;
  }
  CHECK(EV_CR_DEPTH_RANGE, fr_fits_bits(depth, 16));
  const word_t bal_w = {fr_u128(balance.l[0], balance.l[1]), fr_u128(balance.l[2], balance.l[3])};
  CHECK(EV_CR_BAL_CMP_RANGE, word_in_domain(w[0]));
  const int insufficient = fr_cmp(bal_w.hi, w[0].hi) < 0 || (fr_eq(bal_w.hi, w[0].hi) && fr_cmp(bal_w.lo, w[0].lo) < 0);
  CHECK(EV_CR_NONCE_RANGE, fr_fits_bits(nonce_prev, 64));
  const int precheck_ok = depth.l[0] < 1025 && !insufficient && nonce_prev.l[0] < 0xFFFFFFFFFFFFFFFFull;
  const uint64_t sp_delta = 2 + (uint64_t)is_create2;
  int not_collision = 0;
  fr_t look = fr_u64(0); /* rw_counter_offset once the copy lookup's increment (a field element) is in */
  if (precheck_ok) {
    word_t code_hash = {fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
    const word_t empty = code_hash;
    if (has_init_code) {
      int hits = 0;
      for (uint64_t a = 0; a < e->n_aux; a++) {
        const fr_t ar = fr_load(ORC_CELL(e->aux, e->n_aux, 0, a));
        if (fr_eq_u64(ar, row)) { hits++; code_hash.lo = fr_load(ORC_CELL(e->aux, e->n_aux, 1, a)); code_hash.hi = fr_load(ORC_CELL(e->aux, e->n_aux, 2, a)); }
      }
      CHECK(EV_CR_AUX_MISSING, hits == 1);
    }
    fr_t contract;
    if (is_create) contract = contract_address(caller, nonce);
    else CHECK(EV_CR_ADDR2_DOMAIN, contract_address2(caller, w[3], code_hash, &contract));
    const word_t contract_w = {fr_u128(contract.l[0], contract.l[1]), fr_u64(contract.l[2])};
    {
      fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(k)), 1, ZK_TARGET_TxAccessListAccount);
      key[R_ID] = tx_id; key[R_ADDR] = contract; key[R_VAL_LO] = one;
      LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_VAL, &r), EV_CR_AL_UNSAT); k++;
      CHECK(EV_CR_AL_PREV_TYPE, !rw_prev_is_word(e, r));
    }
    LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 0, contract, ZK_ACC_CodeHash, &r), EV_CR_CHASH_UNSAT); k++;
    const word_t callee_hash = rw_value(e, r);
    LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 0, contract, ZK_ACC_Nonce, &r), EV_CR_CNONCE_UNSAT); k++;
    NOT_WORD(rw_val_is_word(e, r), EV_CR_CNONCE_UNSAT);
    const fr_t callee_nonce = rw_cell(e, R_VAL_LO, r);
    const int is_empty_hash = fr_is_zero(fr_add(fr_sub(callee_hash.lo, empty.lo), fr_sub(callee_hash.hi, empty.hi)));
    const int is_zero_hash = fr_is_zero(fr_add(callee_hash.lo, callee_hash.hi));
    not_collision = fr_is_zero(callee_nonce) && (is_empty_hash || is_zero_hash);
    if (not_collision) {
      fr_t ret;
      W2FQ(w[4], 20, &ret, EV_CR_RETURN_DOMAIN);
      CHECK(EV_CR_RETURN_EQ, fr_eq(ret, fr_mul(is_success, contract)));
      fr_t callee_rev_end, callee_persistent;
      CALL_CCV(k, 0, callee_call_id, ZK_CC_RwCounterEndOfReversion, &callee_rev_end, EV_CR_CREVEND_UNSAT); k++;
      CALL_CCV(k, 0, callee_call_id, ZK_CC_IsPersistent, &callee_persistent, EV_CR_CPERSIST_UNSAT); k++;
      CHECK(EV_CR_CPERSIST_EQ, fr_eq(callee_persistent, fr_mul(is_persistent, is_success)));
      /* transfer(caller, contract, value, callee_reversion_info) */
      if (!balance_write(e, row, fr_add(rwc, fr_u64(k)), caller, callee_persistent, callee_rev_end, EV_CR_SEND_UNSAT, &r)) return;
      k++;
      { word_t ws[2] = {rw_value(e, r), w[0]}; fr_t carry; const word_t sum = add_words_n(ws, 2, &carry);
        CHECK(EV_CR_SEND_EQ, word_eq(rw_prev(e, r), sum)); CHECK(EV_CR_SEND_CARRY, fr_is_zero(carry)); }
      if (!balance_write(e, row, fr_add(rwc, fr_u64(k)), contract, callee_persistent, fr_sub(callee_rev_end, one), EV_CR_RECV_UNSAT, &r)) return;
      k++;
      { word_t ws[2] = {rw_prev(e, r), w[0]}; fr_t carry; const word_t sum = add_words_n(ws, 2, &carry);
        CHECK(EV_CR_RECV_EQ, word_eq(rw_value(e, r), sum)); CHECK(EV_CR_RECV_CARRY, fr_is_zero(carry)); }
      LK(account_lookup(e, fr_add(rwc, fr_u64(k)), 1, contract, ZK_ACC_Nonce, &r), EV_CR_NEWNONCE_UNSAT); k++;
      NOT_WORD(rw_val_is_word(e, r), EV_CR_NEWNONCE_UNSAT);
      CHECK(EV_CR_NEWNONCE_PREV_TYPE, !rw_prev_is_word(e, r));
      CHECK(EV_CR_NEWNONCE_EQ, fr_eq_u64(rw_cell(e, R_VAL_LO, r), 1));
      if (has_init_code) {
        const word_t next_hash = {NXT(S_HASH_LO), NXT(S_HASH_HI)};
        fr_t inc;
        if (!need1(e, copy_lookup_dw(e, call_id, ZK_COPY_Memory, next_hash, ZK_COPY_Bytecode, offset, fr_add(offset, size), fr_u64(0), size,
                                     fr_add(rwc, fr_u64(k)), &inc), EV_CR_COPY_UNSAT, row)) return;
        look = inc;
        fr_t code_size;
        if (!need1(e, bytecode_lookup(e, next_hash.lo, next_hash.hi, 1, fr_u64(0), 0, &code_size), EV_CR_CODE_LEN_UNSAT, row)) return;
        CHECK(EV_CR_CODE_LEN_EQ, fr_eq(code_size, size));
        {
          const fr_t want[5] = {fr_add(CUR(S_PC), one), fr_add(sp, fr_u64(sp_delta)), fr_sub(fr_sub(gas_left, fr_u64(gas_cost)), callee_gas_left),
                                fr_u64(next_mem), fr_add(CUR(S_REV), one)};
          static const uint64_t TAGS[5] = {ZK_CC_ProgramCounter, ZK_CC_StackPointer, ZK_CC_GasLeft, ZK_CC_MemorySize, ZK_CC_ReversibleWriteCounter};
          for (int t = 0; t < 5; t++) {
            uint32_t r_; LK(cc_rw_lookup(e, fr_add(fr_add(rwc, look), fr_u64(k)), 1, call_id, TAGS[t], &r_), EV_CR_SAVE0_UNSAT + 4 * t); k++;
            NOT_WORD(rw_val_is_word(e, r_), EV_CR_SAVE0_UNSAT + 4 * t);
            CHECK(EV_CR_SAVE0_UNSAT + 4 * t + 3, fr_eq(rw_cell(e, R_VAL_LO, r_), want[t]));
          }
        }
        {
          const word_t want[10] = {{call_id, fr_u64(0)}, {tx_id, fr_u64(0)}, {fr_add(depth, one), fr_u64(0)}, caller_w, contract_w,
                                   {is_success, fr_u64(0)}, zero, zero, {one, fr_u64(0)}, code_hash};
          static const uint64_t TAGS[10] = {ZK_CC_CallerId, ZK_CC_TxId, ZK_CC_Depth, ZK_CC_CallerAddress, ZK_CC_CalleeAddress, ZK_CC_IsSuccess,
                                            ZK_CC_IsStatic, ZK_CC_IsRoot, ZK_CC_IsCreate, ZK_CC_CodeHash};
          for (int t = 0; t < 10; t++) {
            LK(cc_rw_lookup(e, fr_add(fr_add(rwc, look), fr_u64(k)), 0, callee_call_id, TAGS[t], &r), EV_CR_CTX0_UNSAT + 3 * t); k++;
            CHECK(EV_CR_CTX0_UNSAT + 3 * t + 2, word_eq(rw_value(e, r), want[t]));
          }
        }
        CHECK(EV_CR_NC_RWC, fr_eq(NXT(S_RWC), fr_add(fr_add(rwc, look), fr_u64(k))));
        CHECK(EV_CR_NC_CALL_ID, fr_eq(NXT(S_CALL_ID), callee_call_id));
        CHECK(EV_CR_NC_IS_ROOT, fr_is_zero(NXT(S_IS_ROOT)));
        CHECK(EV_CR_NC_IS_CREATE, fr_eq_u64(NXT(S_IS_CREATE), 1));
        CHECK(EV_CR_NC_GAS, fr_eq(NXT(S_GAS), callee_gas_left));
        CHECK(EV_CR_NC_REV, fr_eq_u64(NXT(S_REV), 3));
        CHECK(EV_CR_NC_LOG, fr_eq(NXT(S_LOG), CUR(S_LOG)));
        CHECK(EV_CR_NC_PC, fr_is_zero(NXT(S_PC)));
        CHECK(EV_CR_NC_SP, fr_eq_u64(NXT(S_SP), 1024));
        CHECK(EV_CR_NC_MEM, fr_is_zero(NXT(S_MEM)));
        return;
      }
    }
  }
  /* pre-check failure, address collision, or nothing to run: stay in the caller's context */
  if (!precheck_ok || !not_collision) CHECK(EV_CR_FAIL_SUCCESS, fr_is_zero(is_success));
  {
    static const uint64_t TAGS[3] = {ZK_CC_LastCalleeId, ZK_CC_LastCalleeReturnDataOffset, ZK_CC_LastCalleeReturnDataLength};
    for (int t = 0; t < 3; t++) {
      fr_t v;
      CALL_CCV(k, 1, call_id, TAGS[t], &v, EV_CR_LAST0_UNSAT + 4 * t); k++;
      CHECK(EV_CR_LAST0_UNSAT + 4 * t + 3, fr_is_zero(v));
    }
  }
  CHECK(EV_CR_SAME_RWC, fr_eq(NXT(S_RWC), fr_add(rwc, fr_u64(k))));
  CHECK(EV_CR_SAME_PC, fr_eq(NXT(S_PC), fr_add(CUR(S_PC), one)));
  CHECK(EV_CR_SAME_SP, fr_eq(NXT(S_SP), fr_add(sp, fr_u64(sp_delta))));
  CHECK(EV_CR_SAME_REV, fr_eq(NXT(S_REV), fr_add(CUR(S_REV), fr_u64(not_collision ? 3 : 0))));
  CHECK(EV_CR_SAME_GAS, fr_eq(NXT(S_GAS), fr_sub(gas_left, fr_u64(gas_cost))));
  CHECK(EV_CR_SAME_MEM, fr_eq_u64(NXT(S_MEM), next_mem));
  CHECK(EV_CR_SAME_CALL_ID, fr_eq(NXT(S_CALL_ID), call_id));
  CHECK(EV_CR_SAME_IS_ROOT, fr_eq(NXT(S_IS_ROOT), CUR(S_IS_ROOT)));
  CHECK(EV_CR_SAME_IS_CREATE, fr_eq(NXT(S_IS_CREATE), CUR(S_IS_CREATE)));
  CHECK(EV_CR_SAME_CODE_HASH, fr_eq(NXT(S_HASH_LO), CUR(S_HASH_LO)) && fr_eq(NXT(S_HASH_HI), CUR(S_HASH_HI)));
}

/* ---- ErrorOutOfGasSloadSstore: error_oog_sload_sstore.py:16-60 (read_account_storage_to_access_list instruction.py:1088-1097,
 * account_storage_read :1015-1026, constrain_error_state).  StepState.aux_data is the slot's committed value as an INT
 * (Word(aux_data) splits it), taken from the step-aux side table.
 * Pinned by tests/golden/evm25.npz (verdicts of the reference's verify_step on all 34 cases of its own test). */
static void gadget_error_oog_sload_sstore(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const int is_sstore = fr_eq_u64(opcode, 0x55), is_sload = fr_eq_u64(opcode, 0x54);
  CHECK(EV_ESS_OPCODE, is_sstore + is_sload == 1);
  word_t key_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &key_w), EV_ESS_KEY_UNSAT, row)) return;
  fr_t tx_id, callee;
  ST_CC(1, ZK_CC_TxId, &tx_id, EV_ESS_TXID_UNSAT);
  uint32_t r;
  LK(cc_lookup(e, fr_add(rwc, fr_u64(2)), call_id, ZK_CC_CalleeAddress, &r), EV_ESS_CALLEE_UNSAT);
  W2FQ(rw_value(e, r), 20, &callee, EV_ESS_CALLEE_DOMAIN);
  {
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(3)), 0, ZK_TARGET_TxAccessListAccountStorage);
    key[R_ID] = tx_id; key[R_ADDR] = callee; key[R_KEY_LO] = key_w.lo; key[R_KEY_HI] = key_w.hi;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_KEY, &r), EV_ESS_AL_UNSAT);
    NOT_WORD(rw_val_is_word(e, r), EV_ESS_AL_UNSAT);
  }
  const fr_t is_warm = rw_cell(e, R_VAL_LO, r);
  uint64_t gas_cost, n_rw = 4;
  if (is_sload) gas_cost = fr_eq_u64(is_warm, 1) ? 100 : 2100;
  else {
    word_t value;
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(4)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(1)), &value), EV_ESS_VAL_UNSAT, row)) return;
    {
      fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(5)), 0, ZK_TARGET_AccountStorage);
      key[R_ID] = tx_id; key[R_ADDR] = callee; key[R_KEY_LO] = key_w.lo; key[R_KEY_HI] = key_w.hi;
      LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_KEY, &r), EV_ESS_READ_UNSAT);
    }
    const word_t value_prev = rw_value(e, r);
    n_rw = 6;
    word_t original = {fr_u64(0), fr_u64(0)};
    int hits = 0;
    for (uint64_t a = 0; a < e->n_aux; a++)
      if (fr_eq_u64(fr_load(ORC_CELL(e->aux, e->n_aux, 0, a)), row)) {
        hits++; original.lo = fr_load(ORC_CELL(e->aux, e->n_aux, 1, a)); original.hi = fr_load(ORC_CELL(e->aux, e->n_aux, 2, a));
      }
    CHECK(EV_ESS_AUX_MISSING, hits == 1);
    /* Word(lo + (hi << 128)): the integer must fit 32 bytes, and is split again at bit 128 */
    CHECK(EV_ESS_AUX_RANGE, fr_fits_bits(original.hi, 128));
    original.hi = fr_add(original.hi, fr_hi128(original.lo)); /* < 2^128 + 2^126: no wrap */
    original.lo = fr_lo128(original.lo);
    CHECK(EV_ESS_AUX_RANGE, fr_fits_bits(original.hi, 128));
    if (word_eq(value, value_prev)) gas_cost = 100;
    else if (word_eq(value_prev, original)) gas_cost = (fr_is_zero(original.lo) && fr_is_zero(original.hi)) ? 20000 : 2900;
    else gas_cost = 100;
    if (fr_is_zero(is_warm)) gas_cost += 2100;
  }
  const fr_t gas_left = CUR(S_GAS);
  CHECK(EV_ESS_GAS_RANGE, fr_fits_bits(gas_left, 64));
  const int insufficient = gas_left.l[0] < gas_cost;
  if (is_sload) CHECK(EV_ESS_SLOAD_NOT_OOG, insufficient);
  else CHECK(EV_ESS_SSTORE_NOT_OOG, insufficient || gas_left.l[0] <= 2300);
  error_state_tail(e, i, row, n_rw);
}

/* ---- ErrorOutOfGasCREATE: error_oog_create.py:19-65.  In a root call the init code is priced as tx call data, one
 * tx_calldata_lookup per byte (`for idx in range(size)`): the walk ends at the first index the tx table does not hold, so
 * it is bounded by the table, not by `size`.
 * Pinned by tests/golden/evm26.npz (verdicts of the reference's verify_step on all 8 cases of its own test). */
static void gadget_error_oog_create(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const int is_create = fr_eq_u64(opcode, 0xf0), is_create2 = fr_eq_u64(opcode, 0xf5);
  CHECK(EV_EOCR_OPCODE, is_create + is_create2 == 1);
  word_t off_w, size_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(1)), &off_w), EV_EOCR_OFF_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(1)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(2)), &size_w), EV_EOCR_SIZE_UNSAT, row)) return;
  fr_t offset = fr_u64(0), size, is_root;
  W2FQ(size_w, 5, &size, EV_EOCR_SIZE_DOMAIN);
  if (!fr_is_zero(size)) W2FQ(off_w, 5, &offset, EV_EOCR_OFF_DOMAIN);
  ST_CC(2, ZK_CC_IsRoot, &is_root, EV_EOCR_ROOT_UNSAT);
  uint64_t gas_cost, n_rw = 3;
  if (fr_eq_u64(is_root, 1)) {
    fr_t tx_id;
    ST_CC(3, ZK_CC_TxId, &tx_id, EV_EOCR_TXID_UNSAT);
    n_rw = 4;
    uint64_t nz = 0;
    for (uint64_t idx = 0; idx < size.l[0]; idx++) {
      uint32_t r;
      fr_t key[3] = {tx_id, fr_u64(ZK_TX_CallData), fr_u64(idx)};
      LK(orc_lookup(&e->tx_ix, key, &r), EV_EOCR_BYTE_UNSAT);
      NOT_WORD(tx_is_word(e, r), EV_EOCR_BYTE_UNSAT);
      nz += !fr_is_zero(tx_value(e, r).lo);
    }
    gas_cost = 53000 + 16 * nz + 4 * (size.l[0] - nz);
  } else {
    uint64_t expansion;
    const int rc_ = mem_expansion_gas(e, i, fr_is_zero(size) ? 0 : (offset.l[0] + size.l[0] + 31) / 32, &expansion);
    if (rc_) { orc_fail(e->res, rc_ == 1 ? EV_EOCR_MEMSIZE_RANGE : EV_EOCR_MEM_MAX, row); return; }
    gas_cost = 32000 + expansion;
  }
  const uint64_t word_size = (size.l[0] + 31) / 32;
  CHECK(EV_EOCR_WORDSIZE_RANGE, !(word_size >> 32));
  gas_cost += 2 * word_size + (is_create2 ? 6 * word_size : 0);
  const int exceeds = 49152 < size.l[0];
  const fr_t gas_left = CUR(S_GAS);
  CHECK(EV_EOCR_GAS_RANGE, fr_fits_bits(gas_left, 64));
  CHECK(EV_EOCR_NOT_OOG, gas_left.l[0] < gas_cost || exceeds);
  error_state_tail(e, i, row, n_rw);
}

/* ---- ErrorOutOfGasPrecompile: execution/precompiles/error_oog_precompile.py:9-35 (no opcode lookup; the reference has no
 * test of its own for it).  As written: only DATACOPY and BN254PAIRING can verify - for the other seven precompiles gas_cost
 * stays a Python int and compare() raises AttributeError on it; the pairing count is a FIELD quotient len / 192.
 * Pinned by tests/golden/evm27.npz (verdicts of the reference's verify_step). */
static void gadget_error_oog_precompile(evm_env* e, uint64_t i, uint64_t row) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID);
  uint32_t r;
  LK(cc_lookup(e, rwc, call_id, ZK_CC_CalleeAddress, &r), EV_EOPC_CALLEE_UNSAT);
  fr_t address, len;
  W2FQ(rw_value(e, r), 20, &address, EV_EOPC_CALLEE_DOMAIN);
  ST_CC(1, ZK_CC_CallDataLength, &len, EV_EOPC_CDLEN_UNSAT);
  CHECK(EV_EOPC_NOT_PRECOMPILE, fr_fits_bits(address, 64) && address.l[0] >= 1 && address.l[0] <= 9);
  const uint64_t a = address.l[0];
  uint64_t gas_cost = 0;
  int cost_ok = 1, cost_is_int = 0; /* cost_ok: gas_cost fits 8 bytes */
  if (a == 8) {
    /* 45000 + 34000 * (len / 192) over the field: below 2^64 only when 192 divides the integer len (and the sum is small) */
    u128 rem = 0;
    for (int k = 3; k >= 0; k--) rem = ((rem << 64) | len.l[k]) % 192;
    if (rem != 0 || !fr_fits_bits(len, 64)) cost_ok = 0;
    else {
      const u128 c = (u128)45000 + (u128)34000 * (len.l[0] / 192);
      if (c >> 64) cost_ok = 0; else gas_cost = (uint64_t)c;
    }
  } else if (a == 4) {
    /* memory_copier_gas_cost(len, 0, 3): (len + 31) // 32 on the field sum, range-checked to 4 bytes */
    const fr_t t = fr_add(len, fr_u64(31));
    const fr_t q = {{(t.l[0] >> 5) | (t.l[1] << 59), (t.l[1] >> 5) | (t.l[2] << 59), (t.l[2] >> 5) | (t.l[3] << 59), t.l[3] >> 5}};
    CHECK(EV_EOPC_WORDSIZE_RANGE, fr_fits_bits(q, 32));
    gas_cost = 15 + 3 * q.l[0];
  } else cost_is_int = 1;
  const fr_t gas_left = CUR(S_GAS);
  CHECK(EV_EOPC_GAS_LEFT_RANGE, fr_fits_bits(gas_left, 64));
  CHECK(EV_EOPC_GAS_INT, !cost_is_int);
  CHECK(EV_EOPC_GAS_COST_RANGE, cost_ok);
  CHECK(EV_EOPC_NOT_OOG, gas_left.l[0] < gas_cost);
  error_state_tail(e, i, row, 2);
}

/* ---- ErrorGasUintOverflow: error_gas_uint_overflow.py:19-171 with instruction.memory_size :1198-1305, calc_mem_size64 /
 * _with_uint :1309-1327, safe_mul :1329-1331, to_word_size :1333-1336.  As written: `if is_dynamic_gas:` tests an FQ object
 * (always true), so memory_size runs for every opcode and returns None (TypeError) for one without a memory operand; an
 * offset below 2^64 is read again with word_to_fq(.., 5) and raises beyond 5 bytes; `val.n < offset64.n` never holds over
 * the field.  Pinned by tests/golden/evm28.npz (the cases the reference's own test runs + the three memory opcodes). */
/* calc_mem_size64_with_uint: 0 ok (size / overflow set), else the failing constraint id */
static int cms_uint(word_t off_w, fr_t len64, u128* size, int* overflow) {
  *size = 0; *overflow = 0;
  if (fr_is_zero(len64)) return 0;
  fr_t off; int rc = word_to_fq_nb(off_w, 31, &off);
  if (rc) return rc == 1 ? EV_EGUO_OFF_DOMAIN : EV_EGUO_OFF_RANGE;
  if (!fr_fits_bits(off, 64)) { *overflow = 1; return 0; }
  if (word_to_fq_nb(off_w, 5, &off)) return EV_EGUO_OFF5_RANGE;
  *size = (u128)off.l[0] + len64.l[0];
  return 0;
}
static int cms(word_t off_w, word_t len_w, u128* size, int* overflow) {
  *size = 0; *overflow = 0;
  fr_t len; const int rc = word_to_fq_nb(len_w, 31, &len);
  if (rc) return rc == 1 ? EV_EGUO_LEN_DOMAIN : EV_EGUO_LEN_RANGE;
  if (!fr_fits_bits(len, 64)) { *overflow = 1; return 0; }
  return cms_uint(off_w, len, size, overflow);
}
static void gadget_error_gas_uint_overflow(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const uint64_t op = fr_fits_bits(opcode, 8) ? opcode.l[0] : 0x100;
  const int is_create = op == 0xf0 || op == 0xf5;
  fr_t cd_len, tx_id, is_root;
  ST_CC(0, ZK_CC_CallDataLength, &cd_len, EV_EGUO_CDLEN_UNSAT);
  ST_CC(1, ZK_CC_TxId, &tx_id, EV_EGUO_TXID_UNSAT);
  ST_CC(2, ZK_CC_IsRoot, &is_root, EV_EGUO_ROOT_UNSAT);
  int calldata_of = 0, initcode_of = 0;
  if (fr_eq_u64(is_root, 1)) {
    const uint64_t len = fr_fits_bits(cd_len, 64) ? cd_len.l[0] : ~0ull; /* a longer walk ends at the first missing byte */
    uint64_t nz = 0;
    for (uint64_t idx = 0; idx < len; idx++) {
      uint32_t r;
      fr_t key[3] = {tx_id, fr_u64(ZK_TX_CallData), fr_u64(idx)};
      LK(orc_lookup(&e->tx_ix, key, &r), EV_EGUO_BYTE_UNSAT);
      NOT_WORD(tx_is_word(e, r), EV_EGUO_BYTE_UNSAT);
      nz += !fr_is_zero(tx_value(e, r).lo);
    }
    if (len > 0) {
      const u128 MAXU = ~(uint64_t)0;
      u128 gas = is_create ? 53000 : 21000;
      const int nz_of = (uint64_t)((MAXU - gas) / 16) < nz;
      gas += (u128)nz * 16;
      int z_of = 0;
      if (!nz_of) {
        CHECK(EV_EGUO_CMP_RANGE, gas <= MAXU);
        const uint64_t z = len - nz;
        z_of = (uint64_t)((MAXU - gas) / 4) < z;
        gas += (u128)z * 4;
      }
      if (is_create) {
        CHECK(EV_EGUO_CMP_RANGE, gas <= MAXU);
        const uint64_t len_words = len / 32 + ((len % 32) ? 1 : 0);
        initcode_of = (uint64_t)((MAXU - gas) / 2) < len_words;
      }
      calldata_of = nz_of + z_of;
    }
  }
  /* memory_size(opcode): the pops of each opcode, then calc_mem_size64 on (offset, length) */
  int n_pop, io = -1, il = -1, mem32 = 0, call = 0;
  switch (op) {
    case 0x20: case 0xf3: case 0xfd: case 0xa0: case 0xa1: case 0xa2: case 0xa3: case 0xa4: n_pop = 2; io = 0; il = 1; break;
    case 0x37: case 0x3e: case 0x39: n_pop = 3; io = 1; il = 2; break;
    case 0x3c: n_pop = 4; io = 2; il = 3; break;
    case 0x51: n_pop = 1; io = 0; mem32 = 1; break;
    case 0x52: case 0x53: n_pop = 2; io = 0; mem32 = 1; break;
    case 0xf0: n_pop = 3; io = 1; il = 2; break;
    case 0xf5: n_pop = 4; io = 1; il = 2; break;
    case 0xf1: case 0xf2: n_pop = 7; call = 3; break;
    case 0xf4: case 0xfa: n_pop = 6; call = 2; break;
    default: orc_fail(e->res, EV_EGUO_OPCODE, row); return;
  }
  word_t w[7];
  for (int k = 0; k < n_pop; k++)
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(3 + k)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(k)), &w[k]), EV_EGUO_POP0_UNSAT + 2 * k, row)) return;
  u128 size = 0; int mem_of = 0, id_;
  if (call) {
    u128 x, y; int ofx, ofy;
    if ((id_ = cms(w[call + 2], w[call + 3], &x, &ofx))) { orc_fail(e->res, id_, row); return; }
    if (ofx) mem_of = 1;
    else {
      if ((id_ = cms(w[call], w[call + 1], &y, &ofy))) { orc_fail(e->res, id_, row); return; }
      if (ofy) mem_of = 1; else size = x > y ? x : y;
    }
  } else if (mem32) {
    if ((id_ = cms_uint(w[io], fr_u64(32), &size, &mem_of))) { orc_fail(e->res, id_, row); return; }
  } else {
    if ((id_ = cms(w[io], w[il], &size, &mem_of))) { orc_fail(e->res, id_, row); return; }
  }
  const int mul_of = size > (u128)(~(uint64_t)0 - 31); /* to_word_size(size) * 32 passes 2^64 - 1 */
  CHECK(EV_EGUO_NOT_OVERFLOW, mem_of + mul_of + calldata_of + initcode_of != 0);
  error_state_tail(e, i, row, 3 + (uint64_t)n_pop);
}

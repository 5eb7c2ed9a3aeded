/* TEST INFRASTRUCTURE — CPU oracle, part of oracle/evm.c (included there).
 * LOG0..LOG4 / ErrorWriteProtection / BLOCKHASH:
 *   log                      evm_circuit/execution/log.py:8-102 (tx_log_lookup_word instruction.py:708-720,
 *                            copy_lookup with a TxLog destination table.py:760-787)
 *   error_write_protection   evm_circuit/execution/error_write_protection.py:12-66
 *   blockhash                evm_circuit/execution/blockhash.py:6-38
 * Pinned by tests/golden/evm18.npz (1,229 verdicts of the reference's verify_step).
 */
enum { TXLOG_Address = 1, TXLOG_Topic = 2, TXLOG_Data = 3, BLOCK_Number = 3, BLOCK_HistoryHash = 8 };
static const fr_t TWO_48 = {{1ull << 48, 0, 0, 0}};
/* rw_lookup(Write, TxLog, id = tx_id, address = index + (field << 32) + (log_id << 48), field_tag = 0, storage_key = Word(0)) */
static int tx_log_lookup(evm_env* e, fr_t rwc_k, fr_t tx_id, fr_t log_id, uint64_t field, uint64_t index, uint32_t* r) {
  fr_t key[14]; rw_key_init(key, rwc_k, 1, ZK_TARGET_TxLog);
  key[R_ID] = tx_id;
  key[R_ADDR] = fr_add(fr_u64(index + (field << 32)), fr_mul(log_id, TWO_48));
  return rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM(R_FIELD) | RWM_KEY, r);
}

static void gadget_log(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  const fr_t topics = fr_sub(opcode, fr_u64(0xa0));
  {
    fr_t key[4] = {fr_u64(ZK_FIXED_Range5), topics, fr_u64(0), fr_u64(0)};
    CHECK(EV_LOG_RANGE5, orc_lookup(&e->fixed_ix, key, 0) >= 1);
  }
  word_t start_w, size_w;
  fr_t mstart, msize;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &start_w), EV_LOG_POP0_UNSAT, row)) return;
  W2FQ(start_w, 8, &mstart, EV_LOG_START_DOMAIN);
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &size_w), EV_LOG_POP1_UNSAT, row)) return;
  W2FQ(size_w, 8, &msize, EV_LOG_SIZE_DOMAIN);
  fr_t tx_id, is_static, is_persistent;
  ST_CC(2, ZK_CC_TxId, &tx_id, EV_LOG_TXID_UNSAT);
  ST_CC(3, ZK_CC_IsStatic, &is_static, EV_LOG_STATIC_UNSAT);
  CHECK(EV_LOG_STATIC_NONZERO, fr_is_zero(is_static));
  uint32_t r;
  LK(cc_lookup(e, fr_add(rwc, fr_u64(4)), call_id, ZK_CC_CalleeAddress, &r), EV_LOG_CALLEE_UNSAT);
  const word_t address = rw_value(e, r);
  ST_CC(5, ZK_CC_IsPersistent, &is_persistent, EV_LOG_PERSIST_UNSAT);
  const int persistent = !fr_is_zero(is_persistent);
  const fr_t log_id = fr_add(CUR(S_LOG), one);
  uint64_t k = 6;
  if (persistent) {
    LK(tx_log_lookup(e, fr_add(rwc, fr_u64(k)), tx_id, log_id, TXLOG_Address, 0, &r), EV_LOG_ADDR_UNSAT);
    CHECK(EV_LOG_ADDR_EQ, word_eq(address, rw_value(e, r)));
    k++;
  }
  const uint64_t n_topics = topics.l[0]; /* 0..4: a row of the Range5 fixed table */
  for (uint64_t t = 0; t < n_topics && t < 4; t++) {
    word_t topic;
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(2 + t)), &topic), EV_LOG_TOPIC_POP_UNSAT, row)) return;
    k++;
    if (persistent) {
      LK(tx_log_lookup(e, fr_add(rwc, fr_u64(k)), tx_id, log_id, TXLOG_Topic, t, &r), EV_LOG_TOPIC_UNSAT);
      CHECK(EV_LOG_TOPIC_EQ, word_eq(topic, rw_value(e, r)));
      k++;
    }
  }
  fr_t rwc_inc = fr_u64(0);
  if (!fr_is_zero(msize) && fr_eq_u64(is_persistent, 1)) {
    fr_t unused;
    const fr_t dst = fr_add(fr_u64((uint64_t)TXLOG_Data << 32), fr_mul(log_id, TWO_48));
    if (!need1(e, copy_lookup(e, call_id, ZK_COPY_Memory, tx_id, ZK_COPY_TxLog, mstart, fr_add(mstart, msize), dst, msize,
                              fr_add(rwc, fr_u64(k)), &rwc_inc, &unused), EV_LOG_COPY_UNSAT, row)) return;
  }
  /* memory_expansion_dynamic_length(mstart, msize), instruction.py:1157-1181 */
  const u128 words = ((u128)mstart.l[0] + msize.l[0] + 31) / 32;
  CHECK(EV_LOG_MEMSIZE_RANGE, !(words >> 32));
  const fr_t cur = CUR(S_MEM);
  CHECK(EV_LOG_MEM_MAX, fr_fits_bits(cur, 32));
  const uint64_t nxt = cur.l[0] < (uint64_t)words ? (uint64_t)words : cur.l[0];
  const uint64_t gas = 375 + 375 * n_topics + 8 * msize.l[0] + (memory_gas_cost(nxt) - memory_gas_cost(cur.l[0]));
  same_context_rl(e, i, row, opcode, fr_add(fr_u64(k), rwc_inc), one, fr_add(fr_u64(2), topics), 1, fr_u64(nxt), fr_u64(gas), 0, is_persistent);
}

static void gadget_error_write_protection(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  const uint64_t op = fr_fits_bits(opcode, 8) ? opcode.l[0] : 0x100;
  CHECK(EV_EWP_OPCODE, op == 0x55 || op == 0xf0 || op == 0xf5 || op == 0xf1 || op == 0xff || (op >= 0xa0 && op <= 0xa4));
  fr_t is_static;
  ST_CC(0, ZK_CC_IsStatic, &is_static, EV_EWP_STATIC_UNSAT);
  CHECK(EV_EWP_NOT_STATIC, fr_eq_u64(is_static, 1));
  uint64_t n_rw = 1;
  if (op == 0xf1) { /* CALL: the transferred value (third stack word) must not be zero */
    word_t value;
    if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(1)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(2)), &value), EV_EWP_VALUE_UNSAT, row)) return;
    CHECK(EV_EWP_VALUE_ZERO, !(fr_is_zero(value.lo) && fr_is_zero(value.hi)));
    n_rw = 2;
  }
  error_state_tail(e, i, row, n_rw);
}

static void gadget_blockhash(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t num_w, pushed;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &num_w), EV_BH_POP_UNSAT, row)) return;
  fr_t number;
  W2FQ(num_w, 8, &number, EV_BH_NUM_DOMAIN);
  uint32_t r;
  LK(block_lookup(e, BLOCK_Number, &r), EV_BH_CUR_UNSAT);
  NOT_WORD(block_is_word(e, r), EV_BH_CUR_UNSAT);
  const fr_t current = block_value(e, r).lo;
  if (!need1(e, rw_lookup(e, fr_add(rwc, one), 1, ZK_TARGET_Stack, call_id, sp, &pushed), EV_BH_PUSH_UNSAT, row)) return;
  /* compare(block_number, current, 8) and compare(current, 256 + block_number, 2): range asserts, blockhash.py:17-18 */
  CHECK(EV_BH_CMP1_RANGE, fr_fits_bits(current, 64));
  const fr_t limit = fr_add(fr_u64(256), number);
  CHECK(EV_BH_CMP2_RANGE, fr_fits_bits(current, 16) && fr_fits_bits(limit, 16));
  const int in_window = number.l[0] < current.l[0] && current.l[0] < limit.l[0];
  word_t want = {fr_u64(0), fr_u64(0)};
  if (in_window) {
    fr_t key[2] = {fr_u64(BLOCK_HistoryHash), number};
    LK(orc_lookup(&e->block_ix, key, &r), EV_BH_HASH_UNSAT);
    want = block_value(e, r);
  }
  CHECK(EV_BH_EQ, word_eq(pushed, want));
  same_context(e, i, row, opcode, 2, one, fr_u64(0));
}

/* ---- error_code_store.py:14-52 (ErrorMaxCodeSizeExceeded, ErrorOutOfGasCodeStore), error_invalid_creation_code.py:11-34 ----
 * Pinned by tests/golden/evm20.npz (700 verdicts of the reference's verify_step). */
static void gadget_error_code_store(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  CHECK(EV_ECS_OPCODE, fr_eq_u64(opcode, 0xf3));
  CHECK(EV_ECS_IS_CREATE, fr_eq_u64(CUR(S_IS_CREATE), 1));
  word_t len_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(1)), &len_w), EV_ECS_LEN_UNSAT, row)) return;
  fr_t length;
  W2FQ(len_w, 5, &length, EV_ECS_LEN_DOMAIN);
  fr_t is_static;
  ST_CC(1, ZK_CC_IsStatic, &is_static, EV_ECS_STATIC_UNSAT);
  CHECK(EV_ECS_STATIC_NONZERO, fr_is_zero(is_static));
  CHECK(EV_ECS_SIZE_RANGE, fr_fits_bits(length, 16)); /* compare(MAX_CODE_SIZE, return_length, N_BYTES_STACK = 2) */
  const int over = 24576 < length.l[0];
  const fr_t gas_left = CUR(S_GAS);
  CHECK(EV_ECS_GAS_RANGE, fr_fits_bits(gas_left, 64)); /* compare(gas_left, 200 * length, 8) */
  const int insufficient = gas_left.l[0] < 200 * length.l[0];
  CHECK(EV_ECS_NEITHER, over || insufficient);
  error_state_tail(e, i, row, 2);
}
static void gadget_error_invalid_creation_code(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP);
  CHECK(EV_ECS_OPCODE, fr_eq_u64(opcode, 0xf3));
  CHECK(EV_ECS_IS_CREATE, fr_eq_u64(CUR(S_IS_CREATE), 1));
  word_t off_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &off_w), EV_ECS_LEN_UNSAT, row)) return;
  fr_t offset;
  W2FQ(off_w, 5, &offset, EV_ECS_LEN_DOMAIN);
  fr_t key[5] = {fr_add(rwc, fr_u64(1)), fr_u64(0), fr_u64(ZK_TARGET_Memory), call_id, offset};
  uint32_t r;
  LK(orc_lookup(&e->rw_ix, key, &r), EV_ECS_BYTE_UNSAT);
  NOT_WORD(rw_val_is_word(e, r), EV_ECS_BYTE_UNSAT);
  CHECK(EV_ECS_FIRST_BYTE, fr_eq_u64(rw_cell(e, R_VAL_LO, r), 0xEF));
  error_state_tail(e, i, row, 2);
}

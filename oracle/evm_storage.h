/* TEST INFRASTRUCTURE — CPU oracle, part of oracle/evm.c (included there).
 * SLOAD / SSTORE / CALLDATALOAD:
 *   sload, sstore   evm_circuit/execution/storage.py:16-48, 51-145
 *                   (account_storage_read / _write instruction.py:1015-1042, add_account_storage_to_access_list :1071-1086,
 *                    tx_refund_write :940-950, state_write + its reversion row :826-863, reversion_info :901-913)
 *   calldataload    evm_circuit/execution/calldataload.py:8-55 (BufferReaderGadget util/memory_gadget.py:5-40)
 * Pinned by tests/golden/evm17.npz (1,654 verdicts of the reference's verify_step).
 */
#define ST_CC(k, field, out, base) do { uint32_t r_; LK(cc_lookup(e, fr_add(rwc, fr_u64(k)), call_id, (field), &r_), (base)); \
  NOT_WORD(rw_val_is_word(e, r_), (base)); *(out) = rw_cell(e, R_VAL_LO, r_); } while (0)

/* state_write(tag, id = tx_id [, address, storage_key] [, value = 1]) at rwc + k, with its reversion row when the
 * call is not persistent (the k_rev-th reversible write of the step) */
static int storage_state_write(evm_env* e, uint64_t row, fr_t rwc_k, uint64_t tag, fr_t tx_id, const fr_t* address, const word_t* key_w,
                               int value_one, fr_t is_persistent, fr_t rwc_rev, int id_base, uint32_t* r_out) {
  fr_t key[14]; rw_key_init(key, rwc_k, 1, tag);
  uint32_t mask = RWM_BASE | RWM(R_ID);
  key[R_ID] = tx_id;
  if (address) { key[R_ADDR] = *address; mask |= RWM(R_ADDR); }
  if (key_w) { key[R_KEY_LO] = key_w->lo; key[R_KEY_HI] = key_w->hi; mask |= RWM_KEY; }
  if (value_one) { key[R_VAL_LO] = fr_u64(1); mask |= RWM_VAL; }
  const int n = rw_lookup_m(e, key, mask, r_out);
  if (n != 1) { orc_fail(e->res, n == 0 ? id_base : id_base + 1, row); return 0; }
  if (fr_is_zero(is_persistent)) {
    uint32_t r2; const int m = reversion_lookup(e, rwc_rev, *r_out, &r2);
    if (m != 1) { orc_fail(e->res, m == 0 ? id_base + 2 : id_base + 3, row); return 0; }
  }
  return 1;
}

static void gadget_sload(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  CHECK(EV_ST_OPCODE, fr_eq_u64(opcode, 0x54));
  fr_t tx_id, rev_end, is_persistent, callee;
  ST_CC(0, ZK_CC_TxId, &tx_id, EV_ST_TXID_UNSAT);
  ST_CC(1, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_ST_REVEND_UNSAT);
  ST_CC(2, ZK_CC_IsPersistent, &is_persistent, EV_ST_PERSIST_UNSAT);
  uint32_t r;
  LK(cc_lookup(e, fr_add(rwc, fr_u64(3)), call_id, ZK_CC_CalleeAddress, &r), EV_ST_CALLEE_UNSAT);
  W2FQ(rw_value(e, r), 20, &callee, EV_ST_CALLEE_DOMAIN);
  word_t key_w, pushed;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(4)), 0, ZK_TARGET_Stack, call_id, sp, &key_w), EV_ST_KEY_UNSAT, row)) return;
  { /* account_storage_read: rw_lookup(Read, AccountStorage, tx_id, callee, storage_key = key) */
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(5)), 0, ZK_TARGET_AccountStorage);
    key[R_ID] = tx_id; key[R_ADDR] = callee; key[R_KEY_LO] = key_w.lo; key[R_KEY_HI] = key_w.hi;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_KEY, &r), EV_ST_READ_UNSAT);
  }
  const word_t value = rw_value(e, r);
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(6)), 1, ZK_TARGET_Stack, call_id, sp, &pushed), EV_ST_PUSH_UNSAT, row)) return;
  CHECK(EV_ST_READ_EQ, word_eq(value, pushed));
  if (!storage_state_write(e, row, fr_add(rwc, fr_u64(7)), ZK_TARGET_TxAccessListAccountStorage, tx_id, &callee, &key_w, 1, is_persistent,
                           fr_sub(rev_end, CUR(S_REV)), EV_ST_AL_UNSAT, &r)) return;
  CHECK(EV_ST_AL_PREV_TYPE, !rw_prev_is_word(e, r));
  const fr_t is_warm = rw_cell(e, R_PREV_LO, r);
  CHECK(EV_ST_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  same_context_r(e, i, row, opcode, fr_u64(8), one, fr_u64(0), 0, fr_u64(0), fr_u64(fr_eq_u64(is_warm, 1) ? 100 : 2100), 1);
}

static void gadget_sstore(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  CHECK(EV_ST_OPCODE, fr_eq_u64(opcode, 0x55));
  fr_t tx_id, is_static, rev_end, is_persistent, callee;
  ST_CC(0, ZK_CC_TxId, &tx_id, EV_ST_TXID_UNSAT);
  ST_CC(1, ZK_CC_IsStatic, &is_static, EV_ST_STATIC_UNSAT);
  CHECK(EV_ST_STATIC_NONZERO, fr_is_zero(is_static));
  ST_CC(2, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_ST_REVEND_UNSAT);
  ST_CC(3, ZK_CC_IsPersistent, &is_persistent, EV_ST_PERSIST_UNSAT);
  uint32_t r;
  LK(cc_lookup(e, fr_add(rwc, fr_u64(4)), call_id, ZK_CC_CalleeAddress, &r), EV_ST_CALLEE_UNSAT);
  W2FQ(rw_value(e, r), 20, &callee, EV_ST_CALLEE_DOMAIN);
  word_t key_w, val_w;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(5)), 0, ZK_TARGET_Stack, call_id, sp, &key_w), EV_ST_KEY_UNSAT, row)) return;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(6)), 0, ZK_TARGET_Stack, call_id, fr_add(sp, one), &val_w), EV_ST_VAL_UNSAT, row)) return;
  const fr_t rev0 = fr_sub(rev_end, CUR(S_REV)); /* rw_counter_of_reversion() counts down per reversible write */
  if (!storage_state_write(e, row, fr_add(rwc, fr_u64(7)), ZK_TARGET_AccountStorage, tx_id, &callee, &key_w, 0, is_persistent, rev0,
                           EV_ST_WRITE_UNSAT, &r)) return;
  const word_t value = rw_value(e, r), value_prev = rw_prev(e, r);
  const word_t original = {rw_cell(e, R_AUX_LO, r), rw_cell(e, R_AUX_HI, r)};
  CHECK(EV_ST_WRITE_EQ, word_eq(val_w, value));
  if (!storage_state_write(e, row, fr_add(rwc, fr_u64(8)), ZK_TARGET_TxAccessListAccountStorage, tx_id, &callee, &key_w, 1, is_persistent,
                           fr_sub(rev0, one), EV_ST_AL_UNSAT, &r)) return;
  CHECK(EV_ST_AL_PREV_TYPE, !rw_prev_is_word(e, r));
  const fr_t is_warm = rw_cell(e, R_PREV_LO, r);
  if (!storage_state_write(e, row, fr_add(rwc, fr_u64(9)), ZK_TARGET_TxRefund, tx_id, 0, 0, 0, is_persistent, fr_sub(rev0, fr_u64(2)),
                           EV_ST_REFUND_UNSAT, &r)) return;
  CHECK(EV_ST_REFUND_TYPE, !rw_val_is_word(e, r));
  CHECK(EV_ST_REFUND_PREV_TYPE, !rw_prev_is_word(e, r));
  const fr_t refund = rw_cell(e, R_VAL_LO, r), refund_prev = rw_cell(e, R_PREV_LO, r);
  /* storage.py:80-123: the EIP-3529 refund rule as nested selects over word (in)equalities */
  const int prev_zero = fr_is_zero(fr_add(value_prev.lo, value_prev.hi)), val_zero = fr_is_zero(fr_add(value.lo, value.hi));
  const int orig_zero = fr_is_zero(fr_add(original.lo, original.hi));
  const int orig_eq_val = word_eq(original, value), prev_eq_val = word_eq(value_prev, value), orig_eq_prev = word_eq(original, value_prev);
  const fr_t clears = fr_u64(4800);
  const fr_t nz_allne = prev_zero ? fr_sub(refund_prev, clears) : (val_zero ? fr_add(refund_prev, clears) : refund_prev);
  const fr_t nz_ne_ne = !orig_eq_val ? nz_allne : fr_add(nz_allne, fr_u64(2900 - 100));
  const fr_t ne_ne = !orig_zero ? nz_ne_ne : (orig_eq_val ? fr_add(refund_prev, fr_u64(20000 - 100)) : refund_prev);
  const fr_t refund_new = prev_eq_val ? refund_prev
                          : (orig_eq_prev ? ((!orig_zero && val_zero) ? fr_add(refund_prev, clears) : refund_prev) : ne_ne);
  CHECK(EV_ST_REFUND_EQ, fr_eq(refund, refund_new));
  const uint64_t warm_gas = (prev_eq_val || !orig_eq_prev) ? 100 : (orig_zero ? 20000 : 2900);
  CHECK(EV_ST_WARM_BOOL, fr_eq_u64(is_warm, 0) || fr_eq_u64(is_warm, 1));
  same_context_r(e, i, row, opcode, fr_u64(10), one, fr_u64(2), 0, fr_u64(0), fr_u64(warm_gas + (fr_eq_u64(is_warm, 1) ? 0 : 2100)), 3);
}

static void gadget_calldataload(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  CHECK(EV_CDL_OPCODE, fr_eq_u64(opcode, 0x35));
  word_t off_w;
  if (!need1(e, rw_lookup(e, rwc, 0, ZK_TARGET_Stack, call_id, sp, &off_w), EV_CDL_POP_UNSAT, row)) return;
  fr_t offset;
  W2FQ(off_w, 8, &offset, EV_CDL_OFF_DOMAIN);
  const int is_root = !fr_is_zero(CUR(S_IS_ROOT)); /* `if instruction.curr.is_root` on a field element: truthy unless zero */
  fr_t src_id, cd_len, cd_off = fr_u64(0);
  uint64_t k_rw = 3;
  ST_CC(1, is_root ? ZK_CC_TxId : ZK_CC_CallerId, &src_id, EV_CDL_CC0_UNSAT);
  ST_CC(2, ZK_CC_CallDataLength, &cd_len, EV_CDL_CC1_UNSAT);
  if (!is_root) { ST_CC(3, ZK_CC_CallDataOffset, &cd_off, EV_CDL_CC2_UNSAT); k_rw = 4; }
  const fr_t src_addr = fr_add(offset, cd_off), src_end = fr_add(cd_len, cd_off);
  /* BufferReaderGadget: min(addr_end, addr_start, 5) -> compare() asserts both fit 5 bytes (instruction.py:447-451) */
  CHECK(EV_CDL_END_RANGE, fr_fits_bits(src_end, 40));
  CHECK(EV_CDL_START_RANGE, fr_fits_bits(src_addr, 40));
  const uint64_t dist = src_end.l[0] > src_addr.l[0] ? src_end.l[0] - src_addr.l[0] : 0;
  const int n_read = dist < 32 ? (int)dist : 32;
  fr_t bytes_[32];
  for (int k = 0; k < n_read; k++) {
    uint32_t r;
    if (is_root) { /* tx_calldata_lookup(tx_id, src_addr + idx).value.value() */
      fr_t key[3] = {src_id, fr_u64(ZK_TX_CallData), fr_add(src_addr, fr_u64(k))};
      LK(orc_lookup(&e->tx_ix, key, &r), EV_CDL_BYTE_UNSAT);
      NOT_WORD(tx_is_word(e, r), EV_CDL_BYTE_UNSAT);
      bytes_[k] = tx_value(e, r).lo;
    } else {       /* memory_lookup(Read, src_addr + idx, caller id) */
      fr_t key[5] = {fr_add(rwc, fr_u64(k_rw)), fr_u64(0), fr_u64(ZK_TARGET_Memory), src_id, fr_add(src_addr, fr_u64(k))};
      LK(orc_lookup(&e->rw_ix, key, &r), EV_CDL_BYTE_UNSAT);
      NOT_WORD(rw_val_is_word(e, r), EV_CDL_BYTE_UNSAT);
      bytes_[k] = rw_cell(e, R_VAL_LO, r);
      k_rw++;
    }
  }
  word_t want = {fr_u64(0), fr_u64(0)};
  for (int k = 0; k < n_read; k++) { /* Word(bytes(calldata_word)): bytes() of a value > 255 -> ValueError */
    CHECK(EV_CDL_BYTES_VALUE, fr_fits_bits(bytes_[k], 8));
    fr_t* half = k < 16 ? &want.lo : &want.hi;
    half->l[(k & 15) >> 3] |= bytes_[k].l[0] << (8 * (k & 7));
  }
  word_t pushed;
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k_rw)), 1, ZK_TARGET_Stack, call_id, sp, &pushed), EV_CDL_PUSH_UNSAT, row)) return;
  CHECK(EV_CDL_EQ, word_eq(want, pushed));
  same_context(e, i, row, opcode, k_rw + 1, one, fr_u64(0));
}

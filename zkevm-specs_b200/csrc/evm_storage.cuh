// Gate programs of SLOAD / SSTORE / CALLDATALOAD (group KG_TX), part of evm.cu (included there).
//   sload, sstore   evm_circuit/execution/storage.py:16-48, 51-145
//                   (account_storage_read / _write instruction.py:1015-1042, add_account_storage_to_access_list :1071-1086,
//                    tx_refund_write :940-950, state_write + its reversion row :826-863, reversion_info :901-913)
//   calldataload    evm_circuit/execution/calldataload.py:8-55 (BufferReaderGadget util/memory_gadget.py:5-40)
// Every rw query of these gadgets names rw_counter, so each one is a positional read of a dense rw table (or a
// probe of the rw_counter index) followed by cell-by-cell confirmation of the named columns (rw_lookup_m).
#pragma once
namespace zk {

// call_context_lookup(field) at rwc + k: the value cell of a non-Word row
#define ST_CC(k, field, out, base)                                                       \
  do {                                                                                   \
    u32 r_ = 0;                                                                          \
    TX_LK(cc_lookup_m(s, fr_add_u64(s.cur(S_RWC), (k)), s.cur(S_CALL_ID), (field), &r_), (base)); \
    TX_NOT_WORD(rw_flag(s, r_, 0), (base));                                              \
    *(out) = rw_cell(s, R_VAL_LO, r_);                                                   \
  } while (0)

// state_write(tag, id = tx_id [, address, storage_key] [, value = 1]) at `rwc_k`, plus its reversion row at
// `rwc_rev` when the call is not persistent; ids id_base .. id_base + 3
ZK_HD_NOINLINE bool storage_state_write(const StepCtx& s, const Fr& rwc_k, u64 tag, const Fr& tx_id, const Fr* address, const Word2* key_w,
                                        bool value_one, const Fr& is_persistent, const Fr& rwc_rev, int id_base, u32* r_out) {
  Fr key[14];
  rw_key_init(key, rwc_k, 1, tag);
  u32 mask = ZK_RWM_BASE | ZK_RWM(R_ID);
  key[R_ID] = tx_id;
  if (address) {
    key[R_ADDR] = *address;
    mask |= ZK_RWM(R_ADDR);
  }
  if (key_w) {
    key[R_KEY_LO] = key_w->lo;
    key[R_KEY_HI] = key_w->hi;
    mask |= ZK_RWM(R_KEY_LO) | ZK_RWM(R_KEY_HI);
  }
  if (value_one) {
    key[R_VAL_LO] = fr_u64(1);
    mask |= ZK_RWM(R_VAL_LO) | ZK_RWM(R_VAL_HI);
  }
  const int n = rw_lookup_m(s, key, mask, r_out);
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

ZK_HD_NOINLINE void gadget_sload(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_ST_OPCODE, fr_eq_u64(opcode, 0x54));
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  Fr tx_id, rev_end, is_persistent, callee = fr_u64(0);
  ST_CC(0, ZK_CC_TxId, &tx_id, EV_ST_TXID_UNSAT);
  ST_CC(1, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_ST_REVEND_UNSAT);
  ST_CC(2, ZK_CC_IsPersistent, &is_persistent, EV_ST_PERSIST_UNSAT);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 3), call_id, ZK_CC_CalleeAddress, &r), EV_ST_CALLEE_UNSAT);
  EOOG_W2FQ(rw_word(s, R_VAL_LO, r), 20, &callee, EV_ST_CALLEE_DOMAIN);
  Word2 key_w{fr_u64(0), fr_u64(0)}, pushed{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 4, 0, sp, &key_w), EV_ST_KEY_UNSAT)) return;
  {  // account_storage_read: rw_lookup(Read, AccountStorage, tx_id, callee, storage_key = key)
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, 5), 0, ZK_TARGET_AccountStorage);
    key[R_ID] = tx_id;
    key[R_ADDR] = callee;
    key[R_KEY_LO] = key_w.lo;
    key[R_KEY_HI] = key_w.hi;
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_KEY_LO) | ZK_RWM(R_KEY_HI), &r), EV_ST_READ_UNSAT);
  }
  const Word2 value = rw_word(s, R_VAL_LO, r);
  if (!need1(s, true, stack_at(s, true, 6, 1, sp, &pushed), EV_ST_PUSH_UNSAT)) return;
  EV_CHECK(EV_ST_READ_EQ, fr_eq(value.lo, pushed.lo) && fr_eq(value.hi, pushed.hi));
  if (!storage_state_write(s, fr_add_u64(rwc, 7), ZK_TARGET_TxAccessListAccountStorage, tx_id, &callee, &key_w, true, is_persistent,
                           fr_sub(rev_end, s.cur(S_REV)), EV_ST_AL_UNSAT, &r)) return;
  EV_CHECK(EV_ST_AL_PREV_TYPE, !rw_flag(s, r, 1));
  const Fr is_warm = rw_cell(s, R_PREV_LO, r);
  const bool warm = fr_eq_u64(is_warm, 1);
  EV_CHECK(EV_ST_WARM_BOOL, warm || fr_eq_u64(is_warm, 0));
  same_context_r_ni(s, opcode, fr_u64(8), fr_u64(1), fr_u64(0), fr_u64(warm ? 100 : 2100), 1);
}

ZK_HD_NOINLINE void gadget_sstore(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_ST_OPCODE, fr_eq_u64(opcode, 0x55));
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), sp = s.cur(S_SP);
  Fr tx_id, is_static, rev_end, is_persistent, callee = fr_u64(0);
  ST_CC(0, ZK_CC_TxId, &tx_id, EV_ST_TXID_UNSAT);
  ST_CC(1, ZK_CC_IsStatic, &is_static, EV_ST_STATIC_UNSAT);
  EV_CHECK(EV_ST_STATIC_NONZERO, fr_is_zero(is_static));
  ST_CC(2, ZK_CC_RwCounterEndOfReversion, &rev_end, EV_ST_REVEND_UNSAT);
  ST_CC(3, ZK_CC_IsPersistent, &is_persistent, EV_ST_PERSIST_UNSAT);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 4), call_id, ZK_CC_CalleeAddress, &r), EV_ST_CALLEE_UNSAT);
  EOOG_W2FQ(rw_word(s, R_VAL_LO, r), 20, &callee, EV_ST_CALLEE_DOMAIN);
  Word2 key_w{fr_u64(0), fr_u64(0)}, val_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 5, 0, sp, &key_w), EV_ST_KEY_UNSAT)) return;
  if (!need1(s, true, stack_at(s, true, 6, 0, fr_add_u64(sp, 1), &val_w), EV_ST_VAL_UNSAT)) return;
  const Fr rev0 = fr_sub(rev_end, s.cur(S_REV));  // rw_counter_of_reversion() counts down per reversible write
  if (!storage_state_write(s, fr_add_u64(rwc, 7), ZK_TARGET_AccountStorage, tx_id, &callee, &key_w, false, is_persistent, rev0,
                           EV_ST_WRITE_UNSAT, &r)) return;
  const Word2 value = rw_word(s, R_VAL_LO, r), value_prev = rw_word(s, R_PREV_LO, r), original = rw_word(s, R_AUX_LO, r);
  EV_CHECK(EV_ST_WRITE_EQ, fr_eq(val_w.lo, value.lo) && fr_eq(val_w.hi, value.hi));
  if (!storage_state_write(s, fr_add_u64(rwc, 8), ZK_TARGET_TxAccessListAccountStorage, tx_id, &callee, &key_w, true, is_persistent,
                           fr_sub(rev0, fr_u64(1)), EV_ST_AL_UNSAT, &r)) return;
  EV_CHECK(EV_ST_AL_PREV_TYPE, !rw_flag(s, r, 1));
  const Fr is_warm = rw_cell(s, R_PREV_LO, r);
  if (!storage_state_write(s, fr_add_u64(rwc, 9), ZK_TARGET_TxRefund, tx_id, nullptr, nullptr, false, is_persistent, fr_sub(rev0, fr_u64(2)),
                           EV_ST_REFUND_UNSAT, &r)) return;
  EV_CHECK(EV_ST_REFUND_TYPE, !rw_flag(s, r, 0));
  EV_CHECK(EV_ST_REFUND_PREV_TYPE, !rw_flag(s, r, 1));
  const Fr refund = rw_cell(s, R_VAL_LO, r), refund_prev = rw_cell(s, R_PREV_LO, r);
  // storage.py:80-123: the EIP-3529 refund rule as nested selects over word (in)equalities
  const bool prev_zero = fr_is_zero(fr_add(value_prev.lo, value_prev.hi)), val_zero = fr_is_zero(fr_add(value.lo, value.hi));
  const bool orig_zero = fr_is_zero(fr_add(original.lo, original.hi));
  const bool orig_eq_val = fr_eq(original.lo, value.lo) && fr_eq(original.hi, value.hi);
  const bool prev_eq_val = fr_eq(value_prev.lo, value.lo) && fr_eq(value_prev.hi, value.hi);
  const bool orig_eq_prev = fr_eq(original.lo, value_prev.lo) && fr_eq(original.hi, value_prev.hi);
  const Fr clears = fr_u64(4800);
  const Fr nz_allne = prev_zero ? fr_sub(refund_prev, clears) : (val_zero ? fr_add(refund_prev, clears) : refund_prev);
  const Fr nz_ne_ne = !orig_eq_val ? nz_allne : fr_add_u64(nz_allne, 2900 - 100);
  const Fr ne_ne = !orig_zero ? nz_ne_ne : (orig_eq_val ? fr_add_u64(refund_prev, 20000 - 100) : refund_prev);
  const Fr refund_new = prev_eq_val ? refund_prev : (orig_eq_prev ? ((!orig_zero && val_zero) ? fr_add(refund_prev, clears) : refund_prev) : ne_ne);
  EV_CHECK(EV_ST_REFUND_EQ, fr_eq(refund, refund_new));
  const u64 warm_gas = (prev_eq_val || !orig_eq_prev) ? 100 : (orig_zero ? 20000 : 2900);
  const bool warm = fr_eq_u64(is_warm, 1);
  EV_CHECK(EV_ST_WARM_BOOL, warm || fr_eq_u64(is_warm, 0));
  same_context_r_ni(s, opcode, fr_u64(10), fr_u64(1), fr_u64(2), fr_u64(warm_gas + (warm ? 0 : 2100)), 3);
}

ZK_HD_NOINLINE void gadget_calldataload(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_CDL_OPCODE, fr_eq_u64(opcode, 0x35));
  const Fr rwc = s.cur(S_RWC), sp = s.cur(S_SP);
  Word2 off_w{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, 0, 0, sp, &off_w), EV_CDL_POP_UNSAT)) return;
  Fr offset = fr_u64(0);
  EOOG_W2FQ(off_w, 8, &offset, EV_CDL_OFF_DOMAIN);
  const bool is_root = !fr_is_zero(s.cur(S_IS_ROOT));  // `if instruction.curr.is_root`: truthy unless zero
  Fr src_id, cd_len, cd_off = fr_u64(0);
  u64 k_rw = 3;
  ST_CC(1, is_root ? ZK_CC_TxId : ZK_CC_CallerId, &src_id, EV_CDL_CC0_UNSAT);
  ST_CC(2, ZK_CC_CallDataLength, &cd_len, EV_CDL_CC1_UNSAT);
  if (!is_root) {
    ST_CC(3, ZK_CC_CallDataOffset, &cd_off, EV_CDL_CC2_UNSAT);
    k_rw = 4;
  }
  const Fr src_addr = fr_add(offset, cd_off), src_end = fr_add(cd_len, cd_off);
  // BufferReaderGadget: min(addr_end, addr_start, 5) -> compare() asserts both fit 5 bytes
  EV_CHECK(EV_CDL_END_RANGE, fr_fits64(src_end) && (src_end.l[0] >> 40) == 0);
  EV_CHECK(EV_CDL_START_RANGE, fr_fits64(src_addr) && (src_addr.l[0] >> 40) == 0);
  const u64 dist = src_end.l[0] > src_addr.l[0] ? src_end.l[0] - src_addr.l[0] : 0;
  const int n_read = dist < 32 ? (int)dist : 32;
  // the bytes are packed as they arrive; `wide` remembers the first one that is not a byte (bytes() raises only
  // after every lookup has been made)
  u64 w4[4] = {0, 0, 0, 0};
  bool wide = false;
#pragma unroll 1
  for (int k = 0; k < n_read; k++) {
    u32 r = 0;
    Fr b;
    if (is_root) {  // tx_calldata_lookup(tx_id, src_addr + idx).value.value()
      Fr key[3] = {src_id, fr_u64(ZK_TX_CallData), fr_add_u64(src_addr, (u64)k)};
      TX_LK(lookup<3>(s.t.tx, key, &r), EV_CDL_BYTE_UNSAT);
      TX_NOT_WORD(tx_is_word(s, r), EV_CDL_BYTE_UNSAT);
      b = table_cell(s.t.tx.tab, 3, r);
    } else {  // memory_lookup(Read, src_addr + idx, caller id)
      Fr key[14];
      rw_key_init(key, fr_add_u64(rwc, k_rw), 0, ZK_TARGET_Memory);
      key[R_ID] = src_id;
      key[R_ADDR] = fr_add_u64(src_addr, (u64)k);
      TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR), &r), EV_CDL_BYTE_UNSAT);
      TX_NOT_WORD(rw_flag(s, r, 0), EV_CDL_BYTE_UNSAT);
      b = rw_cell(s, R_VAL_LO, r);
      k_rw++;
    }
    wide = wide || !(fr_fits64(b) && b.l[0] < 256);
    const u64 sh = b.l[0] << (8 * (k & 7));
    w4[0] |= (k >> 3) == 0 ? sh : 0;
    w4[1] |= (k >> 3) == 1 ? sh : 0;
    w4[2] |= (k >> 3) == 2 ? sh : 0;
    w4[3] |= (k >> 3) == 3 ? sh : 0;
  }
  EV_CHECK(EV_CDL_BYTES_VALUE, !wide);
  Word2 pushed{fr_u64(0), fr_u64(0)};
  if (!need1(s, true, stack_at(s, true, k_rw, 1, sp, &pushed), EV_CDL_PUSH_UNSAT)) return;
  EV_CHECK(EV_CDL_EQ, fr_eq(pushed.lo, fr_u128(w4[0], w4[1])) && fr_eq(pushed.hi, fr_u128(w4[2], w4[3])));
  same_context_ni(s, opcode, k_rw + 1, fr_u64(1), fr_u64(0));
}

}  // namespace zk

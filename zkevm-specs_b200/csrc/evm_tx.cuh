// evm_tx.cuh — transaction-level gate programs of the EVM circuit (included by evm.cu).
//
//   BeginTx   src/zkevm_specs/evm_circuit/execution/begin_tx.py:23-267
//   EndTx     src/zkevm_specs/evm_circuit/execution/end_tx.py:7-86
//   EndBlock  src/zkevm_specs/evm_circuit/execution/end_block.py:67-183
// with the Instruction helpers behind them: rw_lookup with optional columns (instruction.py:792-824),
// state_write + its reversion write (:826-863), add_balance / sub_balance / transfer_with_gas_fee
// (:987-1109), mul_word_by_u64 (:587-597), sub_word (:576-585), add_words (util/arithmetic.py:236-242),
// word_to_fq (:480-484), generate_contract_address (:1338-1340: RLP + Keccak-256, csrc/keccak.cuh),
// step_state_transition_to_new_context (:266-290).
// These steps are a handful per transaction: they run in the KG_TX group kernel, one thread per step,
// lane-private lookups, early return at the first failing constraint (the reference raises there).
#pragma once

namespace zk {

enum { R_PREV_LO = 10, R_PREV_HI, R_AUX_LO, R_AUX_HI };
#define ZK_RWM(c) (1u << (c))
#define ZK_RWM_BASE (ZK_RWM(R_RWC) | ZK_RWM(R_RW) | ZK_RWM(R_TAG))

// Table-derived constants of end_block.py:68-105 (the reference computes them with list comprehensions
// over Python SETS: rows identical in every column count once), filled by k_evm_block_stats.
struct BlockStats {
  u32 max_txs;          // tx rows with tag CallerAddress
  u32 total_txs;        // ... whose value is not zero
  u32 invalid_txs;      // tx rows with tag TxInvalid and value 1
  u32 txinvalid_word;   // tx rows with tag TxInvalid whose value is Word-typed (value() asserts)
  u32 max_wds, total_wds;  // withdrawal rows / with amount != 0
  u32 max_rws;          // rw rows
  u32 pad;
};

// Every rw query names rw_counter: the only rows that can match sit at rw_counter - base of a dense
// table, or in the bucket run of an index keyed on rw_counter alone; the other named cells (bit c of
// `mask` = cell c of `key`) are confirmed one by one, distinct matches counted (0 / 1 / 2 = ambiguous).
ZK_HD_NOINLINE int rw_lookup_m(const StepCtx& s, const Fr* key, u32 mask, u32* row) {
  const IndexDev& ix = s.t.rw;
  const TableDev& t = ix.tab;
  if (t.n_rows == 0) return 0;
  if (pos_enabled(ix) && ix.pos_kind == ZK_POS_DENSE) {
    // dense head (every tag but Start) or the tail run of Start padding rows (lookup.cuh)
    const u64 split = ix.tail_key >= 0 ? ld_u32(ix.pos_ok + 1) : t.n_rows;
    const bool in_tail = ix.tail_key >= 0 && fr_eq_u64(key[ix.tail_col], ix.tail_val);
    const u64 first = in_tail ? split : 0, limit = in_tail ? t.n_rows - split : split;
    if (limit == 0) return 0;
    const u64 base = table_cell(t, R_RWC, first).l[0];
    if (!(fr_fits64(key[R_RWC]) && key[R_RWC].l[0] >= base && key[R_RWC].l[0] - base < limit)) return 0;
    const u64 cand = first + (key[R_RWC].l[0] - base);
    bool ok = true;
    for (int c = 1; c < 14; c++)
      if ((mask >> c) & 1) ok = ok && fr_eq(table_cell(t, c, cand), key[c]);
    *row = (u32)cand;
    return ok ? 1 : 0;
  }
  const IndexDev& rx = s.t.rw_rwc;  // hash index keyed on rw_counter alone
  const u64 mix = rlc_mix(key[R_RWC]);
  const u32 fp = (u32)(mix >> 32);
  u32 b = (u32)mix & rx.mask, first = 0;
  int found = 0;
  for (;;) {
    const u64 slot = ld_u64(&rx.slots[b]);
    if (slot == ZK_EMPTY_SLOT) break;
    if ((u32)(slot >> 32) == fp) {
      const u32 cand = (u32)slot;
      bool ok = fr_eq(table_cell(t, R_RWC, cand), key[R_RWC]);
      for (int c = 1; c < 14 && ok; c++)
        if ((mask >> c) & 1) ok = fr_eq(table_cell(t, c, cand), key[c]);
      if (ok) {
        if (!found) {
          found = 1;
          first = cand;
        } else if (!rows_identical(t, first, cand)) {
          found = 2;
          break;
        }
      }
    }
    b = (b + 1) & rx.mask;
  }
  *row = first;
  return found;
}
ZK_HD Fr rw_cell(const StepCtx& s, int c, u32 r) { return table_cell(s.t.rw.tab, c, r); }
ZK_HD Word2 rw_word(const StepCtx& s, int c_lo, u32 r) { return Word2{rw_cell(s, c_lo, r), rw_cell(s, c_lo + 1, r)}; }
ZK_HD bool rw_flag(const StepCtx& s, u32 r, int bit) { return s.t.rw.tab.flags && ((s.t.rw.tab.flags[r] >> bit) & 1); }
ZK_HD void rw_key_init(Fr key[14], const Fr& rwc, u64 rw, u64 tag) {
  for (int c = 0; c < 14; c++) key[c] = fr_u64(0);
  key[R_RWC] = rwc;
  key[R_RW] = fr_u64(rw);
  key[R_TAG] = fr_u64(tag);
}
ZK_HD_NOINLINE int cc_lookup_m(const StepCtx& s, const Fr& rwc, const Fr& call_id, u64 field, u32* r) {
  Fr key[14];
  rw_key_init(key, rwc, 0, ZK_TARGET_CallContext);
  key[R_ID] = call_id;
  key[R_ADDR] = fr_u64(field);
  return rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR), r);
}
ZK_HD_NOINLINE int receipt_lookup_m(const StepCtx& s, const Fr& rwc, u64 rw, const Fr& tx_id, u64 field, u32* r) {
  Fr key[14];
  rw_key_init(key, rwc, rw, ZK_TARGET_TxReceipt);
  key[R_ID] = tx_id;
  key[R_FIELD] = fr_u64(field);
  return rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_FIELD) | ZK_RWM(R_KEY_LO) | ZK_RWM(R_KEY_HI), r);
}
ZK_HD_NOINLINE int account_lookup_m(const StepCtx& s, const Fr& rwc, u64 rw, const Fr& address, u64 field, u32* r) {
  Fr key[14];
  rw_key_init(key, rwc, rw, ZK_TARGET_Account);
  key[R_ADDR] = address;
  key[R_FIELD] = fr_u64(field);
  return rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ADDR) | ZK_RWM(R_FIELD), r);
}
// the reversion write of state_write (instruction.py:848-861): every cell of the first row, value and
// value_prev swapped, at rw_counter_of_reversion
ZK_HD_NOINLINE int reversion_lookup_m(const StepCtx& s, const Fr& rwc_rev, u32 first, u32* r) {
  Fr key[14];
  for (int c = 0; c < 14; c++) key[c] = rw_cell(s, c, first);
  key[R_RWC] = rwc_rev;
  key[R_RW] = fr_u64(1);
  key[R_VAL_LO] = rw_cell(s, R_PREV_LO, first);
  key[R_VAL_HI] = rw_cell(s, R_PREV_HI, first);
  key[R_PREV_LO] = rw_cell(s, R_VAL_LO, first);
  key[R_PREV_HI] = rw_cell(s, R_VAL_HI, first);
  return rw_lookup_m(s, key, 0x3FFF, r);
}
ZK_HD_NOINLINE int tx_lookup_m(const StepCtx& s, const Fr& tx_id, u64 tag, u32* r) {
  Fr key[3] = {tx_id, fr_u64(tag), fr_u64(0)};
  return lookup<3>(s.t.tx, key, r);
}
ZK_HD Word2 tx_word(const StepCtx& s, u32 r) { return Word2{table_cell(s.t.tx.tab, 3, r), table_cell(s.t.tx.tab, 4, r)}; }
ZK_HD bool tx_is_word(const StepCtx& s, u32 r) { return s.t.tx.tab.flags && (s.t.tx.tab.flags[r] & 1); }
ZK_HD_NOINLINE int block_lookup_m(const StepCtx& s, u64 tag, u32* r) {
  Fr key[2] = {fr_u64(tag), fr_u64(0)};
  return lookup<2>(s.t.block, key, r);
}
ZK_HD Word2 block_word(const StepCtx& s, u32 r) { return Word2{table_cell(s.t.block.tab, 2, r), table_cell(s.t.block.tab, 3, r)}; }
ZK_HD bool block_is_word(const StepCtx& s, u32 r) { return s.t.block.tab.flags && (s.t.block.tab.flags[r] & 1); }

// lookup k of a gadget: ids base (unsat), base + 1 (ambiguous); .value() of its cell: base + 2
#define TX_LK(n_expr, base)                                  \
  do {                                                       \
    const int n_ = (n_expr);                                 \
    if (n_ != 1) {                                           \
      step_fail(s, n_ == 0 ? (base) : (base) + 1);           \
      return;                                                \
    }                                                        \
  } while (0)
#define TX_NOT_WORD(is_word, base) EV_CHECK((base) + 2, !(is_word))

// ---- word helpers ----
ZK_HD Fr fr_hi128(const Fr& a) { return fr_u128(a.l[2], a.l[3]); }  // a.n >> 128
ZK_HD Fr fr_lo128(const Fr& a) { return fr_u128(a.l[0], a.l[1]); }  // a.n mod 2^128
ZK_HD_NOINLINE bool mul_word_by_u64(const Word2& w, const Fr& m, Word2* out) {
  const Fr mm = fr_to_mont(m);
  const Fr t_lo = fr_montmul(mm, w.lo);
  const Fr t_hi = fr_add(fr_montmul(mm, w.hi), fr_hi128(t_lo));
  out->lo = fr_lo128(t_lo);
  out->hi = fr_lo128(t_hi);
  return fr_is_zero(fr_hi128(t_hi));
}
ZK_HD Word2 add_words_n(const Word2* ws, int n, Fr* carry) {
  Fr slo = fr_u64(0), shi = fr_u64(0);
  for (int k = 0; k < n; k++) {
    slo = fr_add(slo, ws[k].lo);
    shi = fr_add(shi, ws[k].hi);
  }
  shi = fr_add(shi, fr_hi128(slo));
  *carry = fr_hi128(shi);
  return Word2{fr_lo128(slo), fr_lo128(shi)};
}
// word_to_fq(word, n_bytes): 0 ok, 1 OverflowError (a half >= 2^128), 2 ConstraintUnsatFailure (bytes n.. not zero)
ZK_HD_NOINLINE int word_to_fq_n(const Word2& w, int n_bytes, Fr* out) {
  if (!word_in_domain(w)) return 1;
  const u64 v[4] = {w.lo.l[0], w.lo.l[1], w.hi.l[0], w.hi.l[1]};
  Fr r = fr_u64(0);
  for (int k = 0; k < 32; k++) {
    const u64 b = (v[k >> 3] >> (8 * (k & 7))) & 0xFF;
    if (k >= n_bytes) {
      if (b) return 2;
    } else {
      r.l[k >> 3] |= b << (8 * (k & 7));
    }
  }
  *out = r;
  return 0;
}
ZK_HD_NOINLINE bool sub_word(const Word2& a, const Word2& b, Word2* out) {
  const Fr two128 = Fr{{0, 0, 1, 0}};
  const bool borrow_lo = fr_lt(a.lo, b.lo);
  Fr dlo = fr_sub(a.lo, b.lo);
  if (borrow_lo) dlo = fr_add(dlo, two128);
  Fr bh = b.hi;  // b.hi.n + borrow_lo as an integer (both < p < 2^254: no wrap)
  if (borrow_lo) {
    u64 c = 1;
    for (int k = 0; k < 4; k++) bh.l[k] = adc64(bh.l[k], 0, c);
  }
  const bool borrow_hi = fr_lt(a.hi, bh);
  Fr dhi = fr_sub(a.hi, b.hi);
  if (borrow_lo) dhi = fr_sub_u64(dhi, 1);
  if (borrow_hi) dhi = fr_add(dhi, two128);
  out->lo = dlo;
  out->hi = dhi;
  return word_in_domain(*out);
}
// keccak(rlp([address (20 bytes big endian), nonce (int)]))[12:] as an integer; address < 2^160
ZK_HD_NOINLINE Fr contract_address(const Fr& address, const Fr& nonce) {
  unsigned char buf[64];
  int n = 1;
  buf[n++] = 0x94;
  for (int k = 19; k >= 0; k--) buf[n++] = (unsigned char)(address.l[k >> 3] >> (8 * (k & 7)));
  int nb = 32;
  while (nb > 0 && ((nonce.l[(nb - 1) >> 3] >> (8 * ((nb - 1) & 7))) & 0xFF) == 0) nb--;
  if (nb == 0) buf[n++] = 0x80;
  else if (nb == 1 && (nonce.l[0] & 0xFF) < 0x80) buf[n++] = (unsigned char)nonce.l[0];
  else {
    buf[n++] = (unsigned char)(0x80 + nb);
    for (int k = nb - 1; k >= 0; k--) buf[n++] = (unsigned char)(nonce.l[k >> 3] >> (8 * (k & 7)));
  }
  buf[0] = (unsigned char)(0xc0 + (n - 1));
  u64 d[4];
  keccak256(buf, (u64)n, d);
  Fr r = fr_u64(0);  // digest bytes 12..31, big endian
  for (int k = 0; k < 20; k++) {
    const int j = 31 - k;  // digest byte index of integer byte k
    r.l[k >> 3] |= ((d[j >> 3] >> (8 * (j & 7))) & 0xFF) << (8 * (k & 7));
  }
  return r;
}

// ================================= EndTx =================================
ZK_HD_NOINLINE void gadget_end_tx(const StepCtx& s) {
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), one = fr_u64(1);
  u32 r = 0;
  TX_LK(cc_lookup_m(s, rwc, call_id, ZK_CC_TxId, &r), EV_ETX_CC_TXID_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_CC_TXID_UNSAT);
  const Fr tx_id = rw_cell(s, R_VAL_LO, r);
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, 1), call_id, ZK_CC_IsPersistent, &r), EV_ETX_CC_PERSIST_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_CC_PERSIST_UNSAT);
  const Fr is_persistent = rw_cell(s, R_VAL_LO, r);
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_TxInvalid, &r), EV_ETX_TX_INVALID_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_ETX_TX_INVALID_UNSAT);
  const Fr is_invalid = tx_word(s, r).lo;
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_Gas, &r), EV_ETX_TX_GAS_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_ETX_TX_GAS_UNSAT);
  const Fr tx_gas = tx_word(s, r).lo;
  const Fr gas_used = fr_sub(tx_gas, s.cur(S_GAS));
  Fr max_refund = fr_u64(0);  // gas_used.n // 5, range-checked to 8 bytes
  {
    unsigned __int128 rem = 0;
    for (int k = 3; k >= 0; k--) {
      const unsigned __int128 cur = (rem << 64) | gas_used.l[k];
      max_refund.l[k] = (u64)(cur / 5);
      rem = cur % 5;
    }
    EV_CHECK(EV_ETX_MAXREFUND_RANGE, fr_fits64(max_refund));
  }
  {
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, 2), 0, ZK_TARGET_TxRefund);
    key[R_ID] = tx_id;
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID), &r), EV_ETX_REFUND_UNSAT);
    TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_REFUND_UNSAT);
  }
  const Fr refund = rw_cell(s, R_VAL_LO, r);
  EV_CHECK(EV_ETX_MIN_RANGE, fr_fits64(refund));
  const Fr eff = max_refund.l[0] < refund.l[0] ? max_refund : refund;
  const bool invalid1 = fr_eq_u64(is_invalid, 1);
  if (invalid1) EV_CHECK(EV_ETX_INVALID_REFUND0, fr_is_zero(eff));
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_GasPrice, &r), EV_ETX_TX_GASPRICE_UNSAT);
  const Word2 gas_price = tx_word(s, r);
  Word2 value;
  EV_CHECK(EV_ETX_MUL1_OVERFLOW, mul_word_by_u64(gas_price, fr_add(s.cur(S_GAS), eff), &value));
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_CallerAddress, &r), EV_ETX_TX_CALLER_UNSAT);
  Fr caller;
  {
    const int rc = word_to_fq_n(tx_word(s, r), 20, &caller);
    EV_CHECK(rc == 1 ? EV_ETX_CALLER_BYTES : EV_ETX_CALLER_RANGE, rc == 0);
  }
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, 3), 1, caller, ZK_ACC_Balance, &r), EV_ETX_BAL_CALLER_UNSAT);
  {
    const Word2 ws[2] = {rw_word(s, R_PREV_LO, r), value};
    Fr carry;
    const Word2 sum = add_words_n(ws, 2, &carry);
    EV_CHECK(EV_ETX_BAL1_EQ, word_eq(rw_word(s, R_VAL_LO, r), sum));
    EV_CHECK(EV_ETX_BAL1_CARRY, fr_is_zero(carry));
  }
  TX_LK(block_lookup_m(s, 6 /* BaseFee */, &r), EV_ETX_BLK_BASEFEE_UNSAT);
  Word2 tip, reward;
  EV_CHECK(EV_ETX_SUBWORD_RANGE, sub_word(gas_price, block_word(s, r), &tip));
  EV_CHECK(EV_ETX_MUL2_OVERFLOW, mul_word_by_u64(tip, gas_used, &reward));
  TX_LK(block_lookup_m(s, 1 /* Coinbase */, &r), EV_ETX_BLK_COINBASE_UNSAT);
  Fr coinbase;
  {
    const int rc = word_to_fq_n(block_word(s, r), 20, &coinbase);
    EV_CHECK(rc == 1 ? EV_ETX_COINBASE_BYTES : EV_ETX_COINBASE_RANGE, rc == 0);
  }
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, 4), 1, coinbase, ZK_ACC_Balance, &r), EV_ETX_BAL_COINBASE_UNSAT);
  {
    const Word2 ws[2] = {rw_word(s, R_PREV_LO, r), reward};
    Fr carry;
    const Word2 sum = add_words_n(ws, 2, &carry);
    EV_CHECK(EV_ETX_BAL2_EQ, word_eq(rw_word(s, R_VAL_LO, r), sum));
    EV_CHECK(EV_ETX_BAL2_CARRY, fr_is_zero(carry));
  }
  TX_LK(receipt_lookup_m(s, fr_add_u64(rwc, 5), 1, tx_id, ZK_RCPT_PostStateOrStatus, &r), EV_ETX_RCPT_STATUS_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_RCPT_STATUS_UNSAT);
  EV_CHECK(EV_ETX_STATUS, fr_eq(fr_mul(fr_sub(one, is_invalid), is_persistent), rw_cell(s, R_VAL_LO, r)));
  TX_LK(receipt_lookup_m(s, fr_add_u64(rwc, 6), 1, tx_id, ZK_RCPT_LogLength, &r), EV_ETX_RCPT_LOG_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_RCPT_LOG_UNSAT);
  const Fr log_id = rw_cell(s, R_VAL_LO, r);
  EV_CHECK(EV_ETX_LOGID, fr_eq(log_id, s.cur(S_LOG)));
  if (invalid1) EV_CHECK(EV_ETX_LOGID0, fr_is_zero(log_id));
  const bool first = fr_eq_u64(tx_id, 1);
  Fr cum = fr_u64(0);
  if (!first) {
    TX_LK(receipt_lookup_m(s, fr_add_u64(rwc, 7), 0, fr_sub(tx_id, one), ZK_RCPT_CumulativeGasUsed, &r), EV_ETX_RCPT_PREVCUM_UNSAT);
    TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_RCPT_PREVCUM_UNSAT);
    cum = rw_cell(s, R_VAL_LO, r);
  }
  TX_LK(receipt_lookup_m(s, fr_add_u64(rwc, first ? 7 : 8), 1, tx_id, ZK_RCPT_CumulativeGasUsed, &r), EV_ETX_RCPT_CUM_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_RCPT_CUM_UNSAT);
  EV_CHECK(EV_ETX_CUMGAS, fr_eq(fr_add(cum, gas_used), rw_cell(s, R_VAL_LO, r)));
  const Fr ns = s.nxt(S_STATE);
  if (fr_eq_u64(ns, ZK_ES_BeginTx)) {
    TX_LK(cc_lookup_m(s, fr_add_u64(rwc, first ? 8 : 9), s.nxt(S_RWC), ZK_CC_TxId, &r), EV_ETX_CC_NEXT_TXID_UNSAT);
    TX_NOT_WORD(rw_flag(s, r, 0), EV_ETX_CC_NEXT_TXID_UNSAT);
    EV_CHECK(EV_ETX_NEXT_TXID, fr_eq(rw_cell(s, R_VAL_LO, r), fr_add(tx_id, one)));
    EV_CHECK(EV_ETX_RWC_BEGINTX, fr_eq(s.nxt(S_RWC), fr_add_u64(rwc, first ? 9 : 10)));
  }
  if (fr_eq_u64(ns, ZK_ES_EndBlock)) {
    EV_CHECK(EV_ETX_RWC_ENDBLOCK, fr_eq(s.nxt(S_RWC), fr_add_u64(rwc, first ? 8 : 9)));
    EV_CHECK(EV_ETX_CALLID_ENDBLOCK, fr_eq(s.nxt(S_CALL_ID), call_id));
  }
}

// ================================= EndBlock =================================
ZK_HD Fr wd_cell(const StepCtx& s, int c, u32 r) { return table_cell(s.t.wd, c, r); }
// row r of a small table is the first of its kind (the reference's table is a set)
ZK_HD bool wd_first_of_kind(const StepCtx& s, u32 r) {
  for (u32 q = 0; q < r; q++)
    if (rows_identical(s.t.wd, q, r)) return false;
  return true;
}
ZK_HD_NOINLINE void gadget_end_block(const StepCtx& s, bool is_last) {
  const Fr rwc = s.cur(S_RWC), call_id = s.cur(S_CALL_ID), one = fr_u64(1);
  const BlockStats st = *s.t.stats;
  EV_CHECK(EV_EB_TXINVALID_TYPE, st.txinvalid_word == 0);
  const Fr total_valid = fr_sub(fr_u64(st.total_txs), fr_u64(st.invalid_txs));
  const bool is_empty = fr_eq_u64(rwc, 1);
  const Fr total_rws = is_empty ? fr_u64(0) : fr_add(rwc, one);
  if (!is_last) {
    EV_CHECK(EV_EB_RWC_SAME, fr_eq(s.nxt(S_RWC), rwc));
    EV_CHECK(EV_EB_CALLID_SAME, fr_eq(s.nxt(S_CALL_ID), call_id));
    return;
  }
  u32 r = 0;
  if (is_empty) {
    EV_CHECK(EV_EB_EMPTY_VALID_TXS, fr_is_zero(total_valid));
    EV_CHECK(EV_EB_EMPTY_WDS, st.total_wds == 0);
  } else {
    TX_LK(cc_lookup_m(s, rwc, call_id, ZK_CC_TxId, &r), EV_EB_CC_TXID_UNSAT);
    TX_NOT_WORD(rw_flag(s, r, 0), EV_EB_CC_TXID_UNSAT);
    EV_CHECK(EV_EB_TXID_EQ, fr_eq_u64(rw_cell(s, R_VAL_LO, r), st.total_txs));
    TX_LK(block_lookup_m(s, 2 /* GasLimit */, &r), EV_EB_BLK_GASLIMIT_UNSAT);
    TX_NOT_WORD(block_is_word(s, r), EV_EB_BLK_GASLIMIT_UNSAT);
    const Fr gas_limit = block_word(s, r).lo;
    TX_LK(receipt_lookup_m(s, fr_add(rwc, one), 0, fr_u64(st.total_txs), ZK_RCPT_CumulativeGasUsed, &r), EV_EB_RCPT_CUM_UNSAT);
    TX_NOT_WORD(rw_flag(s, r, 0), EV_EB_RCPT_CUM_UNSAT);
    const Fr cum = rw_cell(s, R_VAL_LO, r);
    EV_CHECK(EV_EB_GAS_CMP_RANGE, fr_fits64(gas_limit) && fr_fits64(cum));
    EV_CHECK(EV_EB_GAS_LIMIT, !(gas_limit.l[0] < cum.l[0]));
    // withdrawals in id order (sorted() is stable; equal ids keep table order): selection by (id, row)
    const u32 n_wd = (u32)s.t.wd.n_rows;
    u64 off = 2;
    bool have_prev = false;
    Fr prev_id = fr_u64(0);
    u32 prev_row = 0;
    for (u32 k = 0; k < n_wd; k++) {
      bool have = false;
      Fr best_id = fr_u64(0);
      u32 best = 0;
      for (u32 q = 0; q < n_wd; q++) {  // smallest (id, row) strictly after the previous pick
        const Fr id = wd_cell(s, 0, q);
        const bool after = !have_prev || fr_lt(prev_id, id) || (fr_eq(prev_id, id) && q > prev_row);
        if (!after) continue;
        if (!have || fr_lt(id, best_id) || (fr_eq(id, best_id) && q < best)) {
          have = true;
          best_id = id;
          best = q;
        }
      }
      if (!have) break;
      have_prev = true;
      prev_id = best_id;
      prev_row = best;
      const Fr amount = wd_cell(s, 3, best);
      if (fr_is_zero(amount) || !wd_first_of_kind(s, best)) continue;
      // Word(int(amount) * 10^9) must stay below 2^256
      u64 prod[5], c = 0;
      for (int q = 0; q < 4; q++) {
        const unsigned __int128 x = (unsigned __int128)amount.l[q] * 1000000000ull + c;
        prod[q] = (u64)x;
        c = (u64)(x >> 64);
      }
      prod[4] = c;
      EV_CHECK(EV_EB_WD_WORD, prod[4] == 0);
      const Word2 add{fr_u128(prod[0], prod[1]), fr_u128(prod[2], prod[3])};
      TX_LK(account_lookup_m(s, fr_add_u64(rwc, off), 1, wd_cell(s, 2, best), ZK_ACC_Balance, &r), EV_EB_WD_BAL_UNSAT);
      const Word2 ws[2] = {rw_word(s, R_PREV_LO, r), add};
      Fr carry;
      const Word2 sum = add_words_n(ws, 2, &carry);
      EV_CHECK(EV_EB_WD_BAL_EQ, word_eq(rw_word(s, R_VAL_LO, r), sum));
      EV_CHECK(EV_EB_WD_BAL_CARRY, fr_is_zero(carry));
      off++;
    }
  }
  if (st.total_txs != st.max_txs) {
    TX_LK(tx_lookup_m(s, fr_u64((u64)st.total_txs + 1), ZK_TX_CallerAddress, &r), EV_EB_TX_PAD_UNSAT);
    const Word2 v = tx_word(s, r);
    EV_CHECK(EV_EB_TX_PAD_ZERO, fr_is_zero(v.lo) && fr_is_zero(v.hi));
  }
  {
    Fr key[14];
    rw_key_init(key, one, 0, ZK_TARGET_Start);
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE, &r), EV_EB_START1_UNSAT);
    rw_key_init(key, fr_sub(fr_sub(fr_u64(st.max_rws), total_rws), fr_u64(st.total_wds)), 0, ZK_TARGET_Start);
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE, &r), EV_EB_START2_UNSAT);
  }
}

// ================================= BeginTx =================================
ZK_HD_NOINLINE void to_new_context(const StepCtx& s, const Fr& d_rwc, const Fr& call_id, bool is_create, const Word2& code_hash,
                                   const Fr& gas_left) {
  EV_CHECK(EV_BT_NC_RWC, fr_eq(s.nxt(S_RWC), fr_add(s.cur(S_RWC), d_rwc)));
  EV_CHECK(EV_BT_NC_CALL_ID, fr_eq(s.nxt(S_CALL_ID), call_id));
  EV_CHECK(EV_BT_NC_IS_ROOT, fr_eq_u64(s.nxt(S_IS_ROOT), 1));
  EV_CHECK(EV_BT_NC_IS_CREATE, fr_eq_u64(s.nxt(S_IS_CREATE), is_create ? 1 : 0));
  EV_CHECK(EV_BT_NC_CODE_HASH, fr_eq(s.nxt(S_HASH_LO), code_hash.lo) && fr_eq(s.nxt(S_HASH_HI), code_hash.hi));
  EV_CHECK(EV_BT_NC_GAS_LEFT, fr_eq(s.nxt(S_GAS), gas_left));
  EV_CHECK(EV_BT_NC_REV, fr_eq_u64(s.nxt(S_REV), 2));
  EV_CHECK(EV_BT_NC_LOG_ID, fr_is_zero(s.nxt(S_LOG)));
  EV_CHECK(EV_BT_NC_PC, fr_is_zero(s.nxt(S_PC)));
  EV_CHECK(EV_BT_NC_SP, fr_eq_u64(s.nxt(S_SP), 1024));
  EV_CHECK(EV_BT_NC_MEM, fr_is_zero(s.nxt(S_MEM)));
}
ZK_HD_NOINLINE void gadget_begin_tx(const StepCtx& s, bool is_first) {
  const Fr rwc = s.cur(S_RWC), call_id = rwc, one = fr_u64(1);
  const Word2 zero{fr_u64(0), fr_u64(0)};
  u32 r = 0;
  u64 off = 0;
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, off++), call_id, ZK_CC_TxId, &r), EV_BT_CC_TXID_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_BT_CC_TXID_UNSAT);
  const Fr tx_id = rw_cell(s, R_VAL_LO, r);
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, off++), call_id, 1 /* RwCounterEndOfReversion */, &r), EV_BT_CC_REVEND_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_BT_CC_REVEND_UNSAT);
  const Fr rev_end = rw_cell(s, R_VAL_LO, r);
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, off++), call_id, ZK_CC_IsPersistent, &r), EV_BT_CC_PERSIST_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_BT_CC_PERSIST_UNSAT);
  const Fr is_persistent = rw_cell(s, R_VAL_LO, r);
  u64 rev_count = 0;  // reversible_write_counter of ReversionInfo(call_id given) starts at 0
  TX_LK(cc_lookup_m(s, fr_add_u64(rwc, off++), call_id, 12 /* IsSuccess */, &r), EV_BT_CC_SUCCESS_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_BT_CC_SUCCESS_UNSAT);
  EV_CHECK(EV_BT_SUCCESS_EQ, fr_eq(rw_cell(s, R_VAL_LO, r), is_persistent));
  if (is_first) EV_CHECK(EV_BT_FIRST_TXID, fr_eq_u64(tx_id, 1));
  TX_LK(block_lookup_m(s, 1 /* Coinbase */, &r), EV_BT_BLK_COINBASE_UNSAT);
  Fr coinbase, caller, callee;
  {
    const int rc = word_to_fq_n(block_word(s, r), 20, &coinbase);
    EV_CHECK(rc == 1 ? EV_BT_COINBASE_BYTES : EV_BT_COINBASE_RANGE, rc == 0);
  }
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_CallerAddress, &r), EV_BT_TX_CALLER_UNSAT);
  const Word2 caller_word = tx_word(s, r);
  {
    const int rc = word_to_fq_n(caller_word, 20, &caller);
    EV_CHECK(rc == 1 ? EV_BT_CALLER_BYTES : EV_BT_CALLER_RANGE, rc == 0);
  }
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_CalleeAddress, &r), EV_BT_TX_CALLEE_UNSAT);
  const Word2 callee_word = tx_word(s, r);
  {
    const int rc = word_to_fq_n(callee_word, 20, &callee);
    EV_CHECK(rc == 1 ? EV_BT_CALLEE_BYTES : EV_BT_CALLEE_RANGE, rc == 0);
  }
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_IsCreate, &r), EV_BT_TX_ISCREATE_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_BT_TX_ISCREATE_UNSAT);
  const bool is_create = fr_eq_u64(tx_word(s, r).lo, 1);
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_Value, &r), EV_BT_TX_VALUE_UNSAT);
  const Word2 tx_val = tx_word(s, r);
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_CallDataLength, &r), EV_BT_TX_CDLEN_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_BT_TX_CDLEN_UNSAT);
  const Fr cd_len = tx_word(s, r).lo;
  EV_CHECK(EV_BT_CALLER_NONZERO, !fr_is_zero(caller));
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_TxInvalid, &r), EV_BT_TX_INVALID_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_BT_TX_INVALID_UNSAT);
  const Fr is_invalid = tx_word(s, r).lo;
  const bool invalid1 = fr_eq_u64(is_invalid, 1);
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_Nonce, &r), EV_BT_TX_NONCE_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_BT_TX_NONCE_UNSAT);
  const Fr tx_nonce = tx_word(s, r).lo;
  TX_LK(account_lookup_m(s, fr_add_u64(rwc, off++), 1, caller, ZK_ACC_Nonce, &r), EV_BT_ACC_NONCE_UNSAT);
  TX_NOT_WORD(rw_flag(s, r, 0), EV_BT_ACC_NONCE_UNSAT);
  EV_CHECK(EV_BT_ACC_NONCE_PREV_TYPE, !rw_flag(s, r, 1));
  const Fr nonce = rw_cell(s, R_VAL_LO, r), nonce_prev = rw_cell(s, R_PREV_LO, r);
  const bool nonce_valid = fr_eq(tx_nonce, nonce_prev);
  EV_CHECK(EV_BT_NONCE_EQ, fr_eq(nonce, fr_sub(fr_add(nonce_prev, one), is_invalid)));
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_Gas, &r), EV_BT_TX_GAS_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_BT_TX_GAS_UNSAT);
  const Fr tx_gas = tx_word(s, r).lo;
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_GasPrice, &r), EV_BT_TX_GASPRICE_UNSAT);
  Word2 gas_fee;
  EV_CHECK(EV_BT_GASFEE_OVERFLOW, mul_word_by_u64(tx_word(s, r), tx_gas, &gas_fee));
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_CallDataGasCost, &r), EV_BT_TX_CDGAS_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_BT_TX_CDGAS_UNSAT);
  const Fr cd_gas = tx_word(s, r).lo;
  Fr cost = fr_u64(21000);
  if (is_create) {  // constant_divmod(len + 31, 32, 8); 53000 + words * 2
    const Fr num = fr_add_u64(cd_len, 31);
    const Fr q{{(num.l[0] >> 5) | (num.l[1] << 59), (num.l[1] >> 5) | (num.l[2] << 59), (num.l[2] >> 5) | (num.l[3] << 59), num.l[3] >> 5}};
    EV_CHECK(EV_BT_INITCODE_RANGE, fr_fits64(q));
    cost = fr_add_u64(fr_add(q, q), 53000);  // q < 2^64: 2 q needs no reduction
  }
  TX_LK(tx_lookup_m(s, tx_id, ZK_TX_AccessListGasCost, &r), EV_BT_TX_ALGAS_UNSAT);
  TX_NOT_WORD(tx_is_word(s, r), EV_BT_TX_ALGAS_UNSAT);
  const Fr intrinsic = fr_add(fr_add(cd_gas, cost), tx_word(s, r).lo);
  EV_CHECK(EV_BT_GAS_CMP_RANGE, (tx_gas.l[3] >> 56) == 0 && (intrinsic.l[3] >> 56) == 0);
  const bool gas_not_enough = fr_lt(tx_gas, intrinsic);
  const Fr gas_left = gas_not_enough ? tx_gas : fr_sub(tx_gas, intrinsic);
  // (the reference derives it for every transaction; it can only matter, and is only used, for a creation)
  const Fr contract = is_create ? contract_address(caller, tx_nonce) : fr_u64(0);
  const Fr callee_address = is_create ? contract : callee;
  for (int k = 0; k < 3; k++) {  // access list: coinbase, caller, callee
    Fr key[14];
    rw_key_init(key, fr_add_u64(rwc, off++), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id;
    key[R_ADDR] = k == 0 ? coinbase : (k == 1 ? caller : callee_address);
    key[R_VAL_LO] = one;
    const int base = k == 0 ? EV_BT_AL_COINBASE_UNSAT : (k == 1 ? EV_BT_AL_CALLER_UNSAT : EV_BT_AL_CALLEE_UNSAT);
    TX_LK(rw_lookup_m(s, key, ZK_RWM_BASE | ZK_RWM(R_ID) | ZK_RWM(R_ADDR) | ZK_RWM(R_VAL_LO) | ZK_RWM(R_VAL_HI), &r), base);
    EV_CHECK(base + 2, !rw_flag(s, r, 1));
    EV_CHECK(base + 3, fr_is_zero(rw_cell(s, R_PREV_LO, r)));
  }
  // transfer_with_gas_fee(caller, callee_address, value, gas_fee, reversion_info)
  const Word2 t_value = invalid1 ? zero : tx_val, t_fee = invalid1 ? zero : gas_fee;
  const bool reverts = fr_is_zero(is_persistent);
  Word2 sender_prev;
  {
    TX_LK(account_lookup_m(s, fr_add_u64(rwc, off++), 1, caller, ZK_ACC_Balance, &r), EV_BT_BAL_SENDER_UNSAT);
    const u32 first = r;
    if (reverts) {
      u32 r2 = 0;
      TX_LK(reversion_lookup_m(s, fr_sub_u64(rev_end, rev_count), first, &r2), EV_BT_BAL_SENDER_REV_UNSAT);
      rev_count++;
    }
    const Word2 ws[3] = {rw_word(s, R_VAL_LO, first), t_value, t_fee};
    Fr carry;
    const Word2 sum = add_words_n(ws, 3, &carry);
    sender_prev = rw_word(s, R_PREV_LO, first);
    EV_CHECK(EV_BT_SENDER_EQ, word_eq(sender_prev, sum));
    EV_CHECK(EV_BT_SENDER_CARRY, fr_is_zero(carry));
  }
  {
    TX_LK(account_lookup_m(s, fr_add_u64(rwc, off++), 1, callee_address, ZK_ACC_Balance, &r), EV_BT_BAL_RECV_UNSAT);
    const u32 first = r;
    if (reverts) {
      u32 r2 = 0;
      TX_LK(reversion_lookup_m(s, fr_sub_u64(rev_end, rev_count), first, &r2), EV_BT_BAL_RECV_REV_UNSAT);
      rev_count++;
    }
    const Word2 ws[2] = {rw_word(s, R_PREV_LO, first), t_value};
    Fr carry;
    const Word2 sum = add_words_n(ws, 2, &carry);
    EV_CHECK(EV_BT_RECV_EQ, word_eq(rw_word(s, R_VAL_LO, first), sum));
    EV_CHECK(EV_BT_RECV_CARRY, fr_is_zero(carry));
  }
  Fr bal_prev31, val31, fee31;
  {
    int rc = word_to_fq_n(sender_prev, 31, &bal_prev31);
    EV_CHECK(rc == 1 ? EV_BT_BALPREV_BYTES : EV_BT_BALPREV_RANGE, rc == 0);
    rc = word_to_fq_n(tx_val, 31, &val31);
    EV_CHECK(rc == 1 ? EV_BT_VALUE_BYTES : EV_BT_VALUE_RANGE, rc == 0);
    rc = word_to_fq_n(gas_fee, 31, &fee31);
    EV_CHECK(rc == 1 ? EV_BT_FEE_BYTES : EV_BT_FEE_RANGE, rc == 0);
  }
  const Fr need = fr_add(val31, fee31);
  EV_CHECK(EV_BT_BAL_CMP_RANGE, (need.l[3] >> 56) == 0);
  const bool balance_not_enough = fr_lt(bal_prev31, need);
  const bool invalid_tx = !(!balance_not_enough && !gas_not_enough && nonce_valid);
  EV_CHECK(EV_BT_INVALID_FLAG, fr_eq_u64(is_invalid, invalid_tx ? 1 : 0));
  Word2 code_hash = zero;
  bool to_end_tx;
  if (is_create) {
    to_end_tx = invalid1 || fr_is_zero(cd_len);
    if (!to_end_tx) {
      const Fr at = fr_add_u64(rwc, off);
      Fr key[11] = {tx_id, fr_u64(0), fr_u64(ZK_COPY_TxCalldata), call_id, fr_u64(0), fr_u64(ZK_COPY_RlcAcc), fr_u64(0), cd_len,
                    fr_u64(0), cd_len, at};
      TX_LK(lookup<11>(s.t.copy, key, &r), EV_BT_COPY1_UNSAT);
      EV_CHECK(EV_BT_COPY1_RWC0, fr_is_zero(table_cell(s.t.copy.tab, 13, r)));
      const Fr rlc = table_cell(s.t.copy.tab, 11, r);
      Fr kk[3] = {fr_u64(2), rlc, cd_len};
      TX_LK(lookup<3>(s.t.keccak, kk, &r), EV_BT_KECCAK_UNSAT);
      code_hash.lo = table_cell(s.t.keccak.tab, 3, r);
      code_hash.hi = table_cell(s.t.keccak.tab, 4, r);
      Fr key2[11] = {tx_id, fr_u64(0), fr_u64(ZK_COPY_TxCalldata), code_hash.lo, code_hash.hi, fr_u64(ZK_COPY_Bytecode), fr_u64(0),
                     cd_len, fr_u64(0), cd_len, at};
      TX_LK(lookup<11>(s.t.copy, key2, &r), EV_BT_COPY2_UNSAT);
      EV_CHECK(EV_BT_COPY2_RWC0, fr_is_zero(table_cell(s.t.copy.tab, 13, r)));
    }
  } else {
    EV_CHECK(EV_BT_PRECOMPILE, !(fr_fits64(callee) && callee.l[0] >= 1 && callee.l[0] <= 9));
    TX_LK(account_lookup_m(s, fr_add_u64(rwc, off++), 0, callee, ZK_ACC_CodeHash, &r), EV_BT_ACC_CODEHASH_UNSAT);
    code_hash = rw_word(s, R_VAL_LO, r);
    // is_equal_word: the FIELD sum of the two half differences is zero (instruction.py:411-414, 489-490)
    const Word2 empty{fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
    const bool is_empty = fr_is_zero(fr_add(fr_sub(code_hash.lo, empty.lo), fr_sub(code_hash.hi, empty.hi)));
    to_end_tx = is_empty || invalid1;
  }
  if (to_end_tx) {
    EV_CHECK(EV_BT_PERSISTENT1, fr_eq_u64(is_persistent, 1));
    EV_CHECK(EV_BT_NEXT_ENDTX, fr_eq_u64(s.nxt(S_STATE), ZK_ES_EndTx));
    EV_CHECK(EV_BT_END_RWC, fr_eq(s.nxt(S_RWC), fr_add_u64(rwc, off)));
    EV_CHECK(EV_BT_END_CALLID, fr_eq(s.nxt(S_CALL_ID), call_id));
    return;
  }
  const Word2 addr_word = is_create ? Word2{fr_lo128(contract), fr_hi128(contract)} : callee_word;
  const u64 TAGS[13] = {4, 5, 6, 7, 8, 11, 14, 18, 19, 20, 15, 16, 17};
  for (int k = 0; k < 13; k++) {
    Word2 expect = zero;
    switch (k) {
      case 0: case 10: expect.lo = one; break;                       // Depth, IsRoot
      case 1: expect = caller_word; break;
      case 2: expect = addr_word; break;
      case 4: expect.lo = cd_len; break;
      case 5: expect = tx_val; break;
      case 11: expect.lo = fr_u64(is_create ? 1 : 0); break;
      case 12: expect = code_hash; break;
      default: break;                                                 // offsets, IsStatic, LastCallee*: 0
    }
    TX_LK(cc_lookup_m(s, fr_add_u64(rwc, off++), call_id, TAGS[k], &r), EV_BT_CTX0_UNSAT + 3 * k);
    EV_CHECK(EV_BT_CTX0_UNSAT + 3 * k + 2, word_eq(rw_word(s, R_VAL_LO, r), expect));
  }
  to_new_context(s, fr_u64(off), call_id, is_create, code_hash, gas_left);
}

}  // namespace zk

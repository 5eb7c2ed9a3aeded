/*
 * oracle/evm_tx.h — TEST INFRASTRUCTURE (CPU oracle), included by oracle/evm.c.
 *
 * Restates the transaction-level gadgets of the reference's EVM circuit:
 *   begin_tx     src/zkevm_specs/evm_circuit/execution/begin_tx.py:23-267
 *   end_tx       src/zkevm_specs/evm_circuit/execution/end_tx.py:7-86
 *   end_block    src/zkevm_specs/evm_circuit/execution/end_block.py:67-183
 * with the Instruction helpers they use: rw_lookup with optional columns (instruction.py:792-824),
 * state_write + reversion (:826-863), account / receipt / refund / access-list lookups (:723-755,
 * 937-1057), add_balance / sub_balance / transfer_with_gas_fee (:987-1109), mul_word_by_u64 (:587-597),
 * sub_word (:576-585), add_words (util/arithmetic.py:236-242), word_to_fq (:480-484), compare (:447-451),
 * constant_divmod (:440-445), generate_contract_address (:1338-1340; keccak + RLP restated here),
 * constrain_step_state_transition (:206-264), step_state_transition_to_new_context (:266-290).
 * Pinned by tests/golden/evm11.npz (verdicts of the reference's own verify_step).
 */
/* ---- rw lookups with optional columns: every query names rw_counter, so the rows are found through an
 * index on rw_counter alone and the remaining named cells are confirmed one by one ---- */
#define RWM(c) (1u << (c))
#define RWM_BASE (RWM(R_RWC) | RWM(R_RW) | RWM(R_TAG))
#define RWM_KEY (RWM(R_KEY_LO) | RWM(R_KEY_HI))
#define RWM_VAL (RWM(R_VAL_LO) | RWM(R_VAL_HI))
#define RWM_PREV (RWM(R_PREV_LO) | RWM(R_PREV_HI))
#define RWM_AUX (RWM(R_AUX_LO) | RWM(R_AUX_HI))
static fr_t rw_cell(evm_env* e, int c, uint32_t row) { return fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, c, row)); }
static int rw_lookup_m(evm_env* e, const fr_t key[14], uint32_t mask, uint32_t* row_out) {
  const orc_index* ix = &e->rwc_ix;
  uint64_t lo = 0, hi = ix->n_rows;
  while (lo < hi) {
    uint64_t mid = (lo + hi) / 2;
    if (fr_cmp(fr_load(ORC_CELL(ix->cells, ix->n_rows, R_RWC, ix->order[mid])), key[R_RWC]) < 0) lo = mid + 1; else hi = mid;
  }
  int found = 0; uint32_t first = 0;
  for (uint64_t j = lo; j < ix->n_rows; j++) {
    const uint32_t r = ix->order[j];
    if (!fr_eq(rw_cell(e, R_RWC, r), key[R_RWC])) break;
    int ok = 1;
    for (int c = 1; c < 14 && ok; c++) if ((mask >> c) & 1) ok = fr_eq(rw_cell(e, c, r), key[c]);
    if (!ok) continue;
    if (!found) { found = 1; first = r; }
    else if (!orc_rows_identical(ix, first, r)) { found = 2; break; }
  }
  if (found) *row_out = first;
  return found;
}
static word_t rw_value(evm_env* e, uint32_t r) { word_t w = {rw_cell(e, R_VAL_LO, r), rw_cell(e, R_VAL_HI, r)}; return w; }
static word_t rw_prev(evm_env* e, uint32_t r) { word_t w = {rw_cell(e, R_PREV_LO, r), rw_cell(e, R_PREV_HI, r)}; return w; }
static int rw_val_is_word(evm_env* e, uint32_t r) { return e->rw_flags && (e->rw_flags[r] & 1); }
static int rw_prev_is_word(evm_env* e, uint32_t r) { return e->rw_flags && (e->rw_flags[r] & 2); }

/* lookup k of a gadget: ids base (UNSAT), base + 1 (AMBIG); returns 1 iff exactly one row */
#define LK(n_expr, base) do { int n_ = (n_expr); if (n_ != 1) { orc_fail(e->res, n_ == 0 ? (base) : (base) + 1, row); return; } } while (0)
/* .value() of the WordOrValue just looked up: id base + 2 */
#define NOT_WORD(is_word, base) CHECK((base) + 2, !(is_word))

static void rw_key_init(fr_t key[14], fr_t rwc, uint64_t rw, uint64_t tag) {
  for (int c = 0; c < 14; c++) key[c] = fr_u64(0);
  key[R_RWC] = rwc; key[R_RW] = fr_u64(rw); key[R_TAG] = fr_u64(tag);
}
/* call_context_lookup_word (instruction.py:890-895) */
static int cc_lookup(evm_env* e, fr_t rwc, fr_t call_id, uint64_t field, uint32_t* r) {
  fr_t key[14]; rw_key_init(key, rwc, 0, ZK_TARGET_CallContext);
  key[R_ID] = call_id; key[R_ADDR] = fr_u64(field);
  return rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR), r);
}
/* tx_receipt_read / tx_receipt_write (instruction.py:723-755) */
static int receipt_lookup(evm_env* e, fr_t rwc, uint64_t rw, fr_t tx_id, uint64_t field, uint32_t* r) {
  fr_t key[14]; rw_key_init(key, rwc, rw, ZK_TARGET_TxReceipt);
  key[R_ID] = tx_id; key[R_FIELD] = fr_u64(field);
  return rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM(R_FIELD) | RWM_KEY, r);
}
/* account_write_word / account_read_word without reversion (instruction.py:957-985) */
static int account_lookup(evm_env* e, fr_t rwc, uint64_t rw, fr_t address, uint64_t field, uint32_t* r) {
  fr_t key[14]; rw_key_init(key, rwc, rw, ZK_TARGET_Account);
  key[R_ADDR] = address; key[R_FIELD] = fr_u64(field);
  return rw_lookup_m(e, key, RWM_BASE | RWM(R_ADDR) | RWM(R_FIELD), r);
}
/* the reversion write of state_write (instruction.py:848-861): every cell of the first row, value and
 * value_prev swapped, at rw_counter_of_reversion */
static int reversion_lookup(evm_env* e, fr_t rwc_rev, uint32_t first, uint32_t* r) {
  fr_t key[14];
  for (int c = 0; c < 14; c++) key[c] = rw_cell(e, c, first);
  key[R_RWC] = rwc_rev; key[R_RW] = fr_u64(1);
  key[R_VAL_LO] = rw_cell(e, R_PREV_LO, first); key[R_VAL_HI] = rw_cell(e, R_PREV_HI, first);
  key[R_PREV_LO] = rw_cell(e, R_VAL_LO, first); key[R_PREV_HI] = rw_cell(e, R_VAL_HI, first);
  return rw_lookup_m(e, key, 0x3FFF, r);
}
static int tx_lookup(evm_env* e, fr_t tx_id, uint64_t tag, uint32_t* r) {
  fr_t key[3] = {tx_id, fr_u64(tag), fr_u64(0)};
  return orc_lookup(&e->tx_ix, key, r);
}
static word_t tx_value(evm_env* e, uint32_t r) {
  word_t w = {fr_load(ORC_CELL(e->tx_ix.cells, e->tx_ix.n_rows, 3, r)), fr_load(ORC_CELL(e->tx_ix.cells, e->tx_ix.n_rows, 4, r))};
  return w;
}
static int tx_is_word(evm_env* e, uint32_t r) { return e->tx_flags && (e->tx_flags[r] & 1); }
static int block_lookup(evm_env* e, uint64_t tag, uint32_t* r) {
  fr_t key[2] = {fr_u64(tag), fr_u64(0)};
  return orc_lookup(&e->block_ix, key, r);
}
static word_t block_value(evm_env* e, uint32_t r) {
  word_t w = {fr_load(ORC_CELL(e->block_ix.cells, e->block_ix.n_rows, 2, r)), fr_load(ORC_CELL(e->block_ix.cells, e->block_ix.n_rows, 3, r))};
  return w;
}
static int block_is_word(evm_env* e, uint32_t r) { return e->block_flags && (e->block_flags[r] & 1); }

/* ---- word helpers ---- */
static fr_t fr_hi128(fr_t a) { return fr_u128(a.l[2], a.l[3]); }   /* a.n >> 128 */
static fr_t fr_lo128(fr_t a) { return fr_u128(a.l[0], a.l[1]); }   /* a.n mod 2^128 */
/* mul_word_by_u64 (instruction.py:587-597): returns 0 if quotient_hi != 0 */
static int mul_word_by_u64(word_t w, fr_t m, word_t* out) {
  const fr_t t_lo = fr_mul(w.lo, m);
  const fr_t t_hi = fr_add(fr_mul(w.hi, m), fr_hi128(t_lo));
  out->lo = fr_lo128(t_lo); out->hi = fr_lo128(t_hi);
  return fr_is_zero(fr_hi128(t_hi));
}
/* add_words (util/arithmetic.py:236-242) on n words: *carry = carry_hi */
static word_t add_words_n(const word_t* ws, int n, fr_t* carry) {
  fr_t slo = fr_u64(0), shi = fr_u64(0);
  for (int k = 0; k < n; k++) { slo = fr_add(slo, ws[k].lo); shi = fr_add(shi, ws[k].hi); }
  shi = fr_add(shi, fr_hi128(slo));
  word_t r = {fr_lo128(slo), fr_lo128(shi)};
  *carry = fr_hi128(shi);
  return r;
}
/* word_to_fq(word, n_bytes) (instruction.py:480-484): 0 ok, 1 OverflowError (a half >= 2^128), 2 raise */
static int word_to_fq_n(word_t w, int n_bytes, fr_t* out) {
  if (!word_in_domain(w)) return 1;
  uint8_t b[32];
  for (int k = 0; k < 16; k++) { b[k] = (uint8_t)(w.lo.l[k >> 3] >> (8 * (k & 7))); b[16 + k] = (uint8_t)(w.hi.l[k >> 3] >> (8 * (k & 7))); }
  for (int k = n_bytes; k < 32; k++) if (b[k]) return 2;
  fr_t v = fr_u64(0);
  for (int k = 0; k < n_bytes; k++) v.l[k >> 3] |= (uint64_t)b[k] << (8 * (k & 7));
  *out = v;
  return 0;
}
/* sub_word (instruction.py:576-585): 0 if the Word constructor's range assertion fails */
static int sub_word(word_t a, word_t b, word_t* out) {
  const int borrow_lo = fr_cmp(a.lo, b.lo) < 0;
  fr_t dlo = fr_sub(a.lo, b.lo);
  if (borrow_lo) dlo = fr_add(dlo, (fr_t){{0, 0, 1, 0}});  /* + 2^128 */
  /* borrow_hi = a.hi.n < b.hi.n + borrow_lo as integers (no wrap: both < p < 2^254) */
  fr_t bh = b.hi; if (borrow_lo) { uint64_t c = 1; for (int k = 0; k < 4 && c; k++) bh.l[k] = adc(bh.l[k], 0, &c); }
  const int borrow_hi = fr_cmp(a.hi, bh) < 0;
  fr_t dhi = fr_sub(a.hi, b.hi);
  if (borrow_lo) dhi = fr_sub(dhi, fr_u64(1));
  if (borrow_hi) dhi = fr_add(dhi, (fr_t){{0, 0, 1, 0}});
  out->lo = dlo; out->hi = dhi;
  return word_in_domain(*out);
}

/* ---- Keccak-256 + RLP for generate_contract_address (instruction.py:1338-1340; the published sponge) ---- */
static uint64_t rotl64(uint64_t v, int s) { return s ? (v << s) | (v >> (64 - s)) : v; }
static void keccak_f(uint64_t a[25]) {
  static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
    0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull,
    0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull,
    0x8000000000008089ull, 0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull,
    0x800000008000000aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) { uint64_t d = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1); for (int y = 0; y < 25; y += 5) a[x + y] ^= d; }
    for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(a[x + 5 * y], ROT[x][y]);
    for (int y = 0; y < 25; y += 5) for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
    a[0] ^= RC[round];
  }
}
static void keccak256_small(const uint8_t* msg, int len, uint8_t out[32]) { /* len < 136 */
  uint8_t blk[136]; memset(blk, 0, sizeof blk); memcpy(blk, msg, (size_t)len);
  blk[len] |= 0x01; blk[135] |= 0x80;
  uint64_t a[25]; memset(a, 0, sizeof a);
  for (int i = 0; i < 17; i++) { uint64_t v = 0; for (int k = 0; k < 8; k++) v |= (uint64_t)blk[8 * i + k] << (8 * k); a[i] = v; }
  keccak_f(a);
  for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(a[i] >> (8 * k));
}
/* keccak(rlp([address (20 bytes, big endian), nonce (int)]))[12:] as an integer; address < 2^160 */
static fr_t contract_address(fr_t address, fr_t nonce) {
  uint8_t buf[64]; int n = 1;
  buf[n++] = 0x94;
  for (int k = 19; k >= 0; k--) buf[n++] = (uint8_t)(address.l[k >> 3] >> (8 * (k & 7)));
  int nb = 32; while (nb > 0 && ((nonce.l[(nb - 1) >> 3] >> (8 * ((nb - 1) & 7))) & 0xFF) == 0) nb--;
  if (nb == 0) buf[n++] = 0x80;
  else if (nb == 1 && (nonce.l[0] & 0xFF) < 0x80) buf[n++] = (uint8_t)nonce.l[0];
  else { buf[n++] = (uint8_t)(0x80 + nb); for (int k = nb - 1; k >= 0; k--) buf[n++] = (uint8_t)(nonce.l[k >> 3] >> (8 * (k & 7))); }
  buf[0] = (uint8_t)(0xc0 + (n - 1));
  uint8_t h[32]; keccak256_small(buf, n, h);
  fr_t r = fr_u64(0);
  for (int k = 0; k < 20; k++) r.l[k >> 3] |= (uint64_t)h[31 - k] << (8 * (k & 7));
  return r;
}

/* ================================= EndTx ================================= */
static void gadget_end_tx(evm_env* e, uint64_t i, uint64_t row) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), one = fr_u64(1);
  uint32_t r;
  LK(cc_lookup(e, rwc, call_id, ZK_CC_TxId, &r), EV_ETX_CC_TXID_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ETX_CC_TXID_UNSAT);
  const fr_t tx_id = rw_cell(e, R_VAL_LO, r);
  LK(cc_lookup(e, fr_add(rwc, one), call_id, ZK_CC_IsPersistent, &r), EV_ETX_CC_PERSIST_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ETX_CC_PERSIST_UNSAT);
  const fr_t is_persistent = rw_cell(e, R_VAL_LO, r);
  LK(tx_lookup(e, tx_id, ZK_TX_TxInvalid, &r), EV_ETX_TX_INVALID_UNSAT); NOT_WORD(tx_is_word(e, r), EV_ETX_TX_INVALID_UNSAT);
  const fr_t is_invalid = tx_value(e, r).lo;
  LK(tx_lookup(e, tx_id, ZK_TX_Gas, &r), EV_ETX_TX_GAS_UNSAT); NOT_WORD(tx_is_word(e, r), EV_ETX_TX_GAS_UNSAT);
  const fr_t tx_gas = tx_value(e, r).lo;
  const fr_t gas_used = fr_sub(tx_gas, CUR(S_GAS));
  /* max_refund = gas_used.n // 5, range-checked to 8 bytes */
  fr_t max_refund;
  {
    u128 rem = 0; fr_t q = fr_u64(0);
    for (int k = 3; k >= 0; k--) { u128 cur = (rem << 64) | gas_used.l[k]; q.l[k] = (uint64_t)(cur / 5); rem = cur % 5; }
    if (!fr_fits_bits(q, 64)) { orc_fail(e->res, EV_ETX_MAXREFUND_RANGE, row); return; }
    max_refund = q;
  }
  {
    fr_t key[14]; rw_key_init(key, fr_add(rwc, fr_u64(2)), 0, ZK_TARGET_TxRefund); key[R_ID] = tx_id;
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID), &r), EV_ETX_REFUND_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_ETX_REFUND_UNSAT);
  }
  const fr_t refund = rw_cell(e, R_VAL_LO, r);
  CHECK(EV_ETX_MIN_RANGE, fr_fits_bits(refund, 64));
  const fr_t eff = fr_cmp(max_refund, refund) < 0 ? max_refund : refund;
  const int invalid1 = fr_eq_u64(is_invalid, 1);
  if (invalid1) CHECK(EV_ETX_INVALID_REFUND0, fr_is_zero(eff));
  LK(tx_lookup(e, tx_id, ZK_TX_GasPrice, &r), EV_ETX_TX_GASPRICE_UNSAT);
  const word_t gas_price = tx_value(e, r);
  word_t value;
  CHECK(EV_ETX_MUL1_OVERFLOW, mul_word_by_u64(gas_price, fr_add(CUR(S_GAS), eff), &value));
  LK(tx_lookup(e, tx_id, ZK_TX_CallerAddress, &r), EV_ETX_TX_CALLER_UNSAT);
  fr_t caller;
  { int rc = word_to_fq_n(tx_value(e, r), 20, &caller); if (rc) { orc_fail(e->res, rc == 1 ? EV_ETX_CALLER_BYTES : EV_ETX_CALLER_RANGE, row); return; } }
  LK(account_lookup(e, fr_add(rwc, fr_u64(3)), 1, caller, ZK_ACC_Balance, &r), EV_ETX_BAL_CALLER_UNSAT);
  {
    word_t ws[2] = {rw_prev(e, r), value}; fr_t carry; word_t sum = add_words_n(ws, 2, &carry);
    CHECK(EV_ETX_BAL1_EQ, word_eq(rw_value(e, r), sum));
    CHECK(EV_ETX_BAL1_CARRY, fr_is_zero(carry));
  }
  LK(block_lookup(e, 6 /* BaseFee */, &r), EV_ETX_BLK_BASEFEE_UNSAT);
  word_t tip, reward;
  CHECK(EV_ETX_SUBWORD_RANGE, sub_word(gas_price, block_value(e, r), &tip));
  CHECK(EV_ETX_MUL2_OVERFLOW, mul_word_by_u64(tip, gas_used, &reward));
  LK(block_lookup(e, 1 /* Coinbase */, &r), EV_ETX_BLK_COINBASE_UNSAT);
  fr_t coinbase;
  { int rc = word_to_fq_n(block_value(e, r), 20, &coinbase); if (rc) { orc_fail(e->res, rc == 1 ? EV_ETX_COINBASE_BYTES : EV_ETX_COINBASE_RANGE, row); return; } }
  LK(account_lookup(e, fr_add(rwc, fr_u64(4)), 1, coinbase, ZK_ACC_Balance, &r), EV_ETX_BAL_COINBASE_UNSAT);
  {
    word_t ws[2] = {rw_prev(e, r), reward}; fr_t carry; word_t sum = add_words_n(ws, 2, &carry);
    CHECK(EV_ETX_BAL2_EQ, word_eq(rw_value(e, r), sum));
    CHECK(EV_ETX_BAL2_CARRY, fr_is_zero(carry));
  }
  LK(receipt_lookup(e, fr_add(rwc, fr_u64(5)), 1, tx_id, ZK_RCPT_PostStateOrStatus, &r), EV_ETX_RCPT_STATUS_UNSAT);
  NOT_WORD(rw_val_is_word(e, r), EV_ETX_RCPT_STATUS_UNSAT);
  CHECK(EV_ETX_STATUS, fr_eq(fr_mul(fr_sub(one, is_invalid), is_persistent), rw_cell(e, R_VAL_LO, r)));
  LK(receipt_lookup(e, fr_add(rwc, fr_u64(6)), 1, tx_id, ZK_RCPT_LogLength, &r), EV_ETX_RCPT_LOG_UNSAT);
  NOT_WORD(rw_val_is_word(e, r), EV_ETX_RCPT_LOG_UNSAT);
  const fr_t log_id = rw_cell(e, R_VAL_LO, r);
  CHECK(EV_ETX_LOGID, fr_eq(log_id, CUR(S_LOG)));
  if (invalid1) CHECK(EV_ETX_LOGID0, fr_is_zero(log_id));
  const int first = fr_eq_u64(tx_id, 1);
  fr_t cum = fr_u64(0);
  if (!first) {
    LK(receipt_lookup(e, fr_add(rwc, fr_u64(7)), 0, fr_sub(tx_id, one), ZK_RCPT_CumulativeGasUsed, &r), EV_ETX_RCPT_PREVCUM_UNSAT);
    NOT_WORD(rw_val_is_word(e, r), EV_ETX_RCPT_PREVCUM_UNSAT);
    cum = rw_cell(e, R_VAL_LO, r);
  }
  LK(receipt_lookup(e, fr_add(rwc, fr_u64(first ? 7 : 8)), 1, tx_id, ZK_RCPT_CumulativeGasUsed, &r), EV_ETX_RCPT_CUM_UNSAT);
  NOT_WORD(rw_val_is_word(e, r), EV_ETX_RCPT_CUM_UNSAT);
  CHECK(EV_ETX_CUMGAS, fr_eq(fr_add(cum, gas_used), rw_cell(e, R_VAL_LO, r)));
  if (state_is(NXT(S_STATE), ZK_ES_BeginTx)) {
    LK(cc_lookup(e, fr_add(rwc, fr_u64(first ? 8 : 9)), NXT(S_RWC), ZK_CC_TxId, &r), EV_ETX_CC_NEXT_TXID_UNSAT);
    NOT_WORD(rw_val_is_word(e, r), EV_ETX_CC_NEXT_TXID_UNSAT);
    CHECK(EV_ETX_NEXT_TXID, fr_eq(rw_cell(e, R_VAL_LO, r), fr_add(tx_id, one)));
    CHECK(EV_ETX_RWC_BEGINTX, fr_eq(NXT(S_RWC), fr_add(rwc, fr_u64(first ? 9 : 10))));
  }
  if (state_is(NXT(S_STATE), ZK_ES_EndBlock)) {
    CHECK(EV_ETX_RWC_ENDBLOCK, fr_eq(NXT(S_RWC), fr_add(rwc, fr_u64(first ? 8 : 9))));
    CHECK(EV_ETX_CALLID_ENDBLOCK, fr_eq(NXT(S_CALL_ID), call_id));
  }
}

/* ================================= EndBlock ================================= */
static int wd_cmp_id(const void* pa, const void* pb);
static __thread evm_env* g_wd_env;
static fr_t wd_cell(evm_env* e, int c, uint32_t r) { return fr_load(ORC_CELL(e->wd_tab, e->n_wd, c, r)); }
static int wd_cmp_id(const void* pa, const void* pb) {
  const uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
  const int c = fr_cmp(wd_cell(g_wd_env, 0, a), wd_cell(g_wd_env, 0, b));
  return c ? c : (a < b ? -1 : a > b);
}
/* the reference's tables are Python sets: rows identical in every column count once.  `first_of_kind[r]` = 1
 * iff no identical row precedes r in the sorted order (len(table), the list comprehensions of end_block.py:68-105) */
static __thread const uint64_t* g_dd_cells; static __thread uint64_t g_dd_rows; static __thread uint32_t g_dd_cols;
static int dd_cmp(const void* pa, const void* pb) {
  const uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
  for (uint32_t c = 0; c < g_dd_cols; c++) {
    const int x = fr_cmp(fr_load(ORC_CELL(g_dd_cells, g_dd_rows, c, a)), fr_load(ORC_CELL(g_dd_cells, g_dd_rows, c, b)));
    if (x) return x;
  }
  return a < b ? -1 : a > b;
}
static uint8_t* first_of_kind(const uint64_t* cells, uint64_t n_rows, uint32_t n_cols, const uint8_t* flags) {
  uint8_t* out = (uint8_t*)calloc(n_rows ? n_rows : 1, 1);
  uint32_t* ord = (uint32_t*)malloc(sizeof(uint32_t) * (n_rows ? n_rows : 1));
  for (uint64_t k = 0; k < n_rows; k++) ord[k] = (uint32_t)k;
  g_dd_cells = cells; g_dd_rows = n_rows; g_dd_cols = n_cols;
  qsort(ord, n_rows, sizeof(uint32_t), dd_cmp);
  for (uint64_t k = 0; k < n_rows; k++) {
    int same = 0;
    if (k > 0) {
      same = 1;
      for (uint32_t c = 0; c < n_cols && same; c++)
        same = fr_eq(fr_load(ORC_CELL(cells, n_rows, c, ord[k])), fr_load(ORC_CELL(cells, n_rows, c, ord[k - 1])));
      if (same && flags) same = flags[ord[k]] == flags[ord[k - 1]];  /* WordOrValue.is_word is not part of row equality ... */
    }
    out[ord[k]] = !same;
  }
  free(ord);
  return out;
}
static void end_block_inner(evm_env* e, uint64_t i, uint64_t row, int is_last, const uint8_t* tx_first, const uint8_t* wd_first,
                            uint64_t max_rws);
static void gadget_end_block(evm_env* e, uint64_t i, uint64_t row, int is_last) {
  uint8_t* tx_first = first_of_kind(e->tx_ix.cells, e->tx_ix.n_rows, 5, 0);
  uint8_t* wd_first = first_of_kind(e->wd_tab, e->n_wd, 4, 0);
  uint8_t* rw_first = first_of_kind(e->rw_ix.cells, e->rw_ix.n_rows, 14, 0);
  uint64_t max_rws = 0;
  for (uint64_t k = 0; k < e->rw_ix.n_rows; k++) max_rws += rw_first[k];
  end_block_inner(e, i, row, is_last, tx_first, wd_first, max_rws);
  free(tx_first); free(wd_first); free(rw_first);
}
static void end_block_inner(evm_env* e, uint64_t i, uint64_t row, int is_last, const uint8_t* tx_first, const uint8_t* wd_first,
                            uint64_t max_rws) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), one = fr_u64(1);
  const orc_index* tx = &e->tx_ix;
  /* table-derived constants (end_block.py:68-105) */
  uint64_t max_txs = 0, total_txs = 0, invalid_txs = 0, total_wds = 0, max_wds = 0;
  for (uint64_t k = 0; k < e->n_wd; k++) max_wds += wd_first[k];
  for (uint64_t k = 0; k < e->n_wd; k++) if (wd_first[k] && !fr_is_zero(wd_cell(e, 3, (uint32_t)k))) total_wds++;
  for (uint64_t k = 0; k < tx->n_rows; k++)
    if (tx_first[k] && fr_eq_u64(fr_load(ORC_CELL(tx->cells, tx->n_rows, 1, k)), ZK_TX_CallerAddress)) {
      max_txs++;
      const word_t v = tx_value(e, (uint32_t)k);
      if (!(fr_is_zero(v.lo) && fr_is_zero(v.hi))) total_txs++;
    }
  for (uint64_t k = 0; k < tx->n_rows; k++)
    if (tx_first[k] && fr_eq_u64(fr_load(ORC_CELL(tx->cells, tx->n_rows, 1, k)), ZK_TX_TxInvalid)) {
      CHECK(EV_EB_TXINVALID_TYPE, !tx_is_word(e, (uint32_t)k));
      if (fr_eq_u64(tx_value(e, (uint32_t)k).lo, 1)) invalid_txs++;
    }
  const fr_t total_valid = fr_sub(fr_u64(total_txs), fr_u64(invalid_txs));
  const int is_empty = fr_eq_u64(rwc, 1);
  const fr_t total_rws = is_empty ? fr_u64(0) : fr_add(rwc, one);
  if (!is_last) {
    CHECK(EV_EB_RWC_SAME, fr_eq(NXT(S_RWC), rwc));
    CHECK(EV_EB_CALLID_SAME, fr_eq(NXT(S_CALL_ID), call_id));
    return;
  }
  uint32_t r;
  if (is_empty) {
    CHECK(EV_EB_EMPTY_VALID_TXS, fr_is_zero(total_valid));
    CHECK(EV_EB_EMPTY_WDS, total_wds == 0);
  } else {
    LK(cc_lookup(e, rwc, call_id, ZK_CC_TxId, &r), EV_EB_CC_TXID_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_EB_CC_TXID_UNSAT);
    CHECK(EV_EB_TXID_EQ, fr_eq_u64(rw_cell(e, R_VAL_LO, r), total_txs));
    LK(block_lookup(e, 2 /* GasLimit */, &r), EV_EB_BLK_GASLIMIT_UNSAT); NOT_WORD(block_is_word(e, r), EV_EB_BLK_GASLIMIT_UNSAT);
    const fr_t gas_limit = block_value(e, r).lo;
    LK(receipt_lookup(e, fr_add(rwc, one), 0, fr_u64(total_txs), ZK_RCPT_CumulativeGasUsed, &r), EV_EB_RCPT_CUM_UNSAT);
    NOT_WORD(rw_val_is_word(e, r), EV_EB_RCPT_CUM_UNSAT);
    const fr_t cum = rw_cell(e, R_VAL_LO, r);
    CHECK(EV_EB_GAS_CMP_RANGE, fr_fits_bits(gas_limit, 64) && fr_fits_bits(cum, 64));
    CHECK(EV_EB_GAS_LIMIT, !(gas_limit.l[0] < cum.l[0]));
    /* withdrawals in id order (sorted() is stable; equal ids keep table order here) */
    uint32_t* ord = (uint32_t*)malloc(sizeof(uint32_t) * (e->n_wd ? e->n_wd : 1));
    for (uint64_t k = 0; k < e->n_wd; k++) ord[k] = (uint32_t)k;
    g_wd_env = e; qsort(ord, e->n_wd, sizeof(uint32_t), wd_cmp_id);
    uint64_t off = 2;
    for (uint64_t k = 0; k < e->n_wd; k++) {
      const uint32_t w = ord[k];
      const fr_t amount = wd_cell(e, 3, w);
      if (!wd_first[w] || fr_is_zero(amount)) continue;
      /* Word(int(amount) * 10^9): must stay below 2^256 */
      uint64_t prod[5] = {0, 0, 0, 0, 0}; uint64_t c = 0;
      for (int q = 0; q < 4; q++) { u128 x = (u128)amount.l[q] * 1000000000ull + c; prod[q] = (uint64_t)x; c = (uint64_t)(x >> 64); }
      prod[4] = c;
      if (prod[4]) { free(ord); orc_fail(e->res, EV_EB_WD_WORD, row); return; }
      const word_t add = {fr_u128(prod[0], prod[1]), fr_u128(prod[2], prod[3])};
      int nn = account_lookup(e, fr_add(rwc, fr_u64(off)), 1, wd_cell(e, 2, w), ZK_ACC_Balance, &r);
      if (nn != 1) { free(ord); orc_fail(e->res, nn == 0 ? EV_EB_WD_BAL_UNSAT : EV_EB_WD_BAL_AMBIG, row); return; }
      word_t ws[2] = {rw_prev(e, r), add}; fr_t carry; word_t sum = add_words_n(ws, 2, &carry);
      if (!word_eq(rw_value(e, r), sum)) { free(ord); orc_fail(e->res, EV_EB_WD_BAL_EQ, row); return; }
      if (!fr_is_zero(carry)) { free(ord); orc_fail(e->res, EV_EB_WD_BAL_CARRY, row); return; }
      off++;
    }
    free(ord);
  }
  if (total_txs != max_txs) {
    LK(tx_lookup(e, fr_u64(total_txs + 1), ZK_TX_CallerAddress, &r), EV_EB_TX_PAD_UNSAT);
    const word_t v = tx_value(e, r);
    CHECK(EV_EB_TX_PAD_ZERO, fr_is_zero(v.lo) && fr_is_zero(v.hi));
  }
  {
    fr_t key[14]; rw_key_init(key, one, 0, ZK_TARGET_Start);
    LK(rw_lookup_m(e, key, RWM_BASE, &r), EV_EB_START1_UNSAT);
    rw_key_init(key, fr_sub(fr_sub(fr_u64(max_rws), total_rws), fr_u64(total_wds)), 0, ZK_TARGET_Start);
    LK(rw_lookup_m(e, key, RWM_BASE, &r), EV_EB_START2_UNSAT);
  }
  (void)max_wds;
}

/* ================================= BeginTx ================================= */
/* constrain_step_state_transition of step_state_transition_to_new_context (instruction.py:266-290) */
static void to_new_context(evm_env* e, uint64_t i, uint64_t row, fr_t d_rwc, fr_t call_id, int is_create, word_t code_hash,
                           fr_t gas_left) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  CHECK(EV_BT_NC_RWC, fr_eq(NXT(S_RWC), fr_add(CUR(S_RWC), d_rwc)));
  CHECK(EV_BT_NC_CALL_ID, fr_eq(NXT(S_CALL_ID), call_id));
  CHECK(EV_BT_NC_IS_ROOT, fr_eq_u64(NXT(S_IS_ROOT), 1));
  CHECK(EV_BT_NC_IS_CREATE, fr_eq_u64(NXT(S_IS_CREATE), (uint64_t)is_create));
  CHECK(EV_BT_NC_CODE_HASH, fr_eq(NXT(S_HASH_LO), code_hash.lo) && fr_eq(NXT(S_HASH_HI), code_hash.hi));
  CHECK(EV_BT_NC_GAS_LEFT, fr_eq(NXT(S_GAS), gas_left));
  CHECK(EV_BT_NC_REV, fr_eq_u64(NXT(S_REV), 2));
  CHECK(EV_BT_NC_LOG_ID, fr_is_zero(NXT(S_LOG)));
  CHECK(EV_BT_NC_PC, fr_is_zero(NXT(S_PC)));
  CHECK(EV_BT_NC_SP, fr_eq_u64(NXT(S_SP), 1024));
  CHECK(EV_BT_NC_MEM, fr_is_zero(NXT(S_MEM)));
}
static void gadget_begin_tx(evm_env* e, uint64_t i, uint64_t row, int is_first) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps; const uint64_t j = i + 1;
  const fr_t rwc = CUR(S_RWC), call_id = rwc, one = fr_u64(1);
  const word_t zero = {fr_u64(0), fr_u64(0)};
  uint32_t r; uint64_t off = 0;
#define NEXT_RWC() fr_add(rwc, fr_u64(off++))
  LK(cc_lookup(e, NEXT_RWC(), call_id, ZK_CC_TxId, &r), EV_BT_CC_TXID_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_BT_CC_TXID_UNSAT);
  const fr_t tx_id = rw_cell(e, R_VAL_LO, r);
  LK(cc_lookup(e, NEXT_RWC(), call_id, 1 /* RwCounterEndOfReversion */, &r), EV_BT_CC_REVEND_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_BT_CC_REVEND_UNSAT);
  const fr_t rev_end = rw_cell(e, R_VAL_LO, r);
  LK(cc_lookup(e, NEXT_RWC(), call_id, ZK_CC_IsPersistent, &r), EV_BT_CC_PERSIST_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_BT_CC_PERSIST_UNSAT);
  const fr_t is_persistent = rw_cell(e, R_VAL_LO, r);
  uint64_t rev_count = 0; /* reversible_write_counter of ReversionInfo(call_id given) starts at 0 */
  LK(cc_lookup(e, NEXT_RWC(), call_id, 12 /* IsSuccess */, &r), EV_BT_CC_SUCCESS_UNSAT); NOT_WORD(rw_val_is_word(e, r), EV_BT_CC_SUCCESS_UNSAT);
  CHECK(EV_BT_SUCCESS_EQ, fr_eq(rw_cell(e, R_VAL_LO, r), is_persistent));
  if (is_first) CHECK(EV_BT_FIRST_TXID, fr_eq_u64(tx_id, 1));
  LK(block_lookup(e, 1 /* Coinbase */, &r), EV_BT_BLK_COINBASE_UNSAT);
  fr_t coinbase, caller, callee;
  { int rc = word_to_fq_n(block_value(e, r), 20, &coinbase); if (rc) { orc_fail(e->res, rc == 1 ? EV_BT_COINBASE_BYTES : EV_BT_COINBASE_RANGE, row); return; } }
  LK(tx_lookup(e, tx_id, ZK_TX_CallerAddress, &r), EV_BT_TX_CALLER_UNSAT);
  const word_t caller_word = tx_value(e, r);
  { int rc = word_to_fq_n(caller_word, 20, &caller); if (rc) { orc_fail(e->res, rc == 1 ? EV_BT_CALLER_BYTES : EV_BT_CALLER_RANGE, row); return; } }
  LK(tx_lookup(e, tx_id, ZK_TX_CalleeAddress, &r), EV_BT_TX_CALLEE_UNSAT);
  const word_t callee_word = tx_value(e, r);
  { int rc = word_to_fq_n(callee_word, 20, &callee); if (rc) { orc_fail(e->res, rc == 1 ? EV_BT_CALLEE_BYTES : EV_BT_CALLEE_RANGE, row); return; } }
  LK(tx_lookup(e, tx_id, ZK_TX_IsCreate, &r), EV_BT_TX_ISCREATE_UNSAT); NOT_WORD(tx_is_word(e, r), EV_BT_TX_ISCREATE_UNSAT);
  const fr_t is_create_f = tx_value(e, r).lo; const int is_create = fr_eq_u64(is_create_f, 1);
  LK(tx_lookup(e, tx_id, ZK_TX_Value, &r), EV_BT_TX_VALUE_UNSAT);
  const word_t tx_val = tx_value(e, r);
  LK(tx_lookup(e, tx_id, ZK_TX_CallDataLength, &r), EV_BT_TX_CDLEN_UNSAT); NOT_WORD(tx_is_word(e, r), EV_BT_TX_CDLEN_UNSAT);
  const fr_t cd_len = tx_value(e, r).lo;
  CHECK(EV_BT_CALLER_NONZERO, !fr_is_zero(caller));
  LK(tx_lookup(e, tx_id, ZK_TX_TxInvalid, &r), EV_BT_TX_INVALID_UNSAT); NOT_WORD(tx_is_word(e, r), EV_BT_TX_INVALID_UNSAT);
  const fr_t is_invalid = tx_value(e, r).lo; const int invalid1 = fr_eq_u64(is_invalid, 1);
  LK(tx_lookup(e, tx_id, ZK_TX_Nonce, &r), EV_BT_TX_NONCE_UNSAT); NOT_WORD(tx_is_word(e, r), EV_BT_TX_NONCE_UNSAT);
  const fr_t tx_nonce = tx_value(e, r).lo;
  LK(account_lookup(e, NEXT_RWC(), 1, caller, ZK_ACC_Nonce, &r), EV_BT_ACC_NONCE_UNSAT);
  NOT_WORD(rw_val_is_word(e, r), EV_BT_ACC_NONCE_UNSAT);
  CHECK(EV_BT_ACC_NONCE_PREV_TYPE, !rw_prev_is_word(e, r));
  const fr_t nonce = rw_cell(e, R_VAL_LO, r), nonce_prev = rw_cell(e, R_PREV_LO, r);
  const int nonce_valid = fr_eq(tx_nonce, nonce_prev);
  CHECK(EV_BT_NONCE_EQ, fr_eq(nonce, fr_sub(fr_add(nonce_prev, one), is_invalid)));
  LK(tx_lookup(e, tx_id, ZK_TX_Gas, &r), EV_BT_TX_GAS_UNSAT); NOT_WORD(tx_is_word(e, r), EV_BT_TX_GAS_UNSAT);
  const fr_t tx_gas = tx_value(e, r).lo;
  LK(tx_lookup(e, tx_id, ZK_TX_GasPrice, &r), EV_BT_TX_GASPRICE_UNSAT);
  word_t gas_fee;
  CHECK(EV_BT_GASFEE_OVERFLOW, mul_word_by_u64(tx_value(e, r), tx_gas, &gas_fee));
  LK(tx_lookup(e, tx_id, ZK_TX_CallDataGasCost, &r), EV_BT_TX_CDGAS_UNSAT); NOT_WORD(tx_is_word(e, r), EV_BT_TX_CDGAS_UNSAT);
  const fr_t cd_gas = tx_value(e, r).lo;
  fr_t cost = fr_u64(21000);
  if (is_create) {
    /* constant_divmod(len + 31, 32, 8) */
    const fr_t num = fr_add(cd_len, fr_u64(31));
    fr_t q = {{(num.l[0] >> 5) | (num.l[1] << 59), (num.l[1] >> 5) | (num.l[2] << 59), (num.l[2] >> 5) | (num.l[3] << 59), num.l[3] >> 5}};
    if (!fr_fits_bits(q, 64)) { orc_fail(e->res, EV_BT_INITCODE_RANGE, row); return; }
    cost = fr_add(fr_u64(53000), fr_mul(q, fr_u64(2)));
  }
  LK(tx_lookup(e, tx_id, ZK_TX_AccessListGasCost, &r), EV_BT_TX_ALGAS_UNSAT); NOT_WORD(tx_is_word(e, r), EV_BT_TX_ALGAS_UNSAT);
  const fr_t intrinsic = fr_add(fr_add(cd_gas, cost), tx_value(e, r).lo);
  CHECK(EV_BT_GAS_CMP_RANGE, fr_fits_bits(tx_gas, 248) && fr_fits_bits(intrinsic, 248));
  const int gas_not_enough = fr_cmp(tx_gas, intrinsic) < 0;
  const fr_t gas_left = gas_not_enough ? tx_gas : fr_sub(tx_gas, intrinsic);
  const fr_t contract = contract_address(caller, tx_nonce);
  const fr_t callee_address = is_create ? contract : callee;
  /* access list: coinbase, caller, callee */
  for (int k = 0; k < 3; k++) {
    const fr_t a = k == 0 ? coinbase : (k == 1 ? caller : callee_address);
    fr_t key[14]; rw_key_init(key, NEXT_RWC(), 1, ZK_TARGET_TxAccessListAccount);
    key[R_ID] = tx_id; key[R_ADDR] = a; key[R_VAL_LO] = one;
    const int base = k == 0 ? EV_BT_AL_COINBASE_UNSAT : (k == 1 ? EV_BT_AL_CALLER_UNSAT : EV_BT_AL_CALLEE_UNSAT);
    LK(rw_lookup_m(e, key, RWM_BASE | RWM(R_ID) | RWM(R_ADDR) | RWM_VAL, &r), base);
    CHECK(base + 2, !rw_prev_is_word(e, r));
    CHECK(base + 3, fr_is_zero(rw_cell(e, R_PREV_LO, r)));
  }
  /* transfer_with_gas_fee(caller, callee_address, value, gas_fee, reversion_info) */
  const word_t t_value = invalid1 ? zero : tx_val, t_fee = invalid1 ? zero : gas_fee;
  const int reverts = fr_is_zero(is_persistent);
  word_t sender_prev;
  {
    LK(account_lookup(e, NEXT_RWC(), 1, caller, ZK_ACC_Balance, &r), EV_BT_BAL_SENDER_UNSAT);
    const uint32_t first = r;
    if (reverts) { uint32_t r2; LK(reversion_lookup(e, fr_sub(rev_end, fr_u64(rev_count)), first, &r2), EV_BT_BAL_SENDER_REV_UNSAT); rev_count++; }
    word_t ws[3] = {rw_value(e, first), t_value, t_fee}; fr_t carry; word_t sum = add_words_n(ws, 3, &carry);
    sender_prev = rw_prev(e, first);
    CHECK(EV_BT_SENDER_EQ, word_eq(sender_prev, sum));
    CHECK(EV_BT_SENDER_CARRY, fr_is_zero(carry));
  }
  {
    LK(account_lookup(e, NEXT_RWC(), 1, callee_address, ZK_ACC_Balance, &r), EV_BT_BAL_RECV_UNSAT);
    const uint32_t first = r;
    if (reverts) { uint32_t r2; LK(reversion_lookup(e, fr_sub(rev_end, fr_u64(rev_count)), first, &r2), EV_BT_BAL_RECV_REV_UNSAT); rev_count++; }
    word_t ws[2] = {rw_prev(e, first), t_value}; fr_t carry; word_t sum = add_words_n(ws, 2, &carry);
    CHECK(EV_BT_RECV_EQ, word_eq(rw_value(e, first), sum));
    CHECK(EV_BT_RECV_CARRY, fr_is_zero(carry));
  }
  fr_t bal_prev31, val31, fee31;
  { int rc = word_to_fq_n(sender_prev, 31, &bal_prev31); if (rc) { orc_fail(e->res, rc == 1 ? EV_BT_BALPREV_BYTES : EV_BT_BALPREV_RANGE, row); return; } }
  { int rc = word_to_fq_n(tx_val, 31, &val31); if (rc) { orc_fail(e->res, rc == 1 ? EV_BT_VALUE_BYTES : EV_BT_VALUE_RANGE, row); return; } }
  { int rc = word_to_fq_n(gas_fee, 31, &fee31); if (rc) { orc_fail(e->res, rc == 1 ? EV_BT_FEE_BYTES : EV_BT_FEE_RANGE, row); return; } }
  const fr_t need = fr_add(val31, fee31);
  CHECK(EV_BT_BAL_CMP_RANGE, fr_fits_bits(need, 248));
  const int balance_not_enough = fr_cmp(bal_prev31, need) < 0;
  const int invalid_tx = !(!balance_not_enough && !gas_not_enough && nonce_valid);
  CHECK(EV_BT_INVALID_FLAG, fr_eq_u64(is_invalid, (uint64_t)invalid_tx));
  word_t code_hash;
  int to_end_tx;
  if (is_create) {
    to_end_tx = invalid1 || fr_is_zero(cd_len);
    if (!to_end_tx) {
      const fr_t at = fr_add(rwc, fr_u64(off));
      fr_t key[11] = {tx_id, fr_u64(0), fr_u64(ZK_COPY_TxCalldata), call_id, fr_u64(0), fr_u64(ZK_COPY_RlcAcc), fr_u64(0), cd_len,
                      fr_u64(0), cd_len, at};
      LK(orc_lookup(&e->copy_ix, key, &r), EV_BT_COPY1_UNSAT);
      CHECK(EV_BT_COPY1_RWC0, fr_is_zero(fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 13, r))));
      const fr_t rlc = fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 11, r));
      fr_t kk[3] = {fr_u64(2), rlc, cd_len};
      LK(orc_lookup(&e->keccak_ix, kk, &r), EV_BT_KECCAK_UNSAT);
      code_hash.lo = fr_load(ORC_CELL(e->keccak_ix.cells, e->keccak_ix.n_rows, 3, r));
      code_hash.hi = fr_load(ORC_CELL(e->keccak_ix.cells, e->keccak_ix.n_rows, 4, r));
      fr_t key2[11] = {tx_id, fr_u64(0), fr_u64(ZK_COPY_TxCalldata), code_hash.lo, code_hash.hi, fr_u64(ZK_COPY_Bytecode), fr_u64(0),
                       cd_len, fr_u64(0), cd_len, at};
      LK(orc_lookup(&e->copy_ix, key2, &r), EV_BT_COPY2_UNSAT);
      CHECK(EV_BT_COPY2_RWC0, fr_is_zero(fr_load(ORC_CELL(e->copy_ix.cells, e->copy_ix.n_rows, 13, r))));
    }
  } else {
    if (fr_fits_bits(callee, 64) && callee.l[0] >= 1 && callee.l[0] <= 9) { orc_fail(e->res, EV_BT_PRECOMPILE, row); return; }
    LK(account_lookup(e, NEXT_RWC(), 0, callee, ZK_ACC_CodeHash, &r), EV_BT_ACC_CODEHASH_UNSAT);
    code_hash = rw_value(e, r);
    /* is_equal_word: the field sum of the two half differences is zero (instruction.py:411-414, 489-490) */
    const word_t empty = {fr_u128(0x7bfad8045d85a470ull, 0xe500b653ca82273bull), fr_u128(0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull)};
    const int is_empty = fr_is_zero(fr_add(fr_sub(code_hash.lo, empty.lo), fr_sub(code_hash.hi, empty.hi)));
    to_end_tx = is_empty || invalid1;
  }
  if (to_end_tx) {
    CHECK(EV_BT_PERSISTENT1, fr_eq_u64(is_persistent, 1));
    CHECK(EV_BT_NEXT_ENDTX, state_is(NXT(S_STATE), ZK_ES_EndTx));
    CHECK(EV_BT_END_RWC, fr_eq(NXT(S_RWC), fr_add(rwc, fr_u64(off))));
    CHECK(EV_BT_END_CALLID, fr_eq(NXT(S_CALL_ID), call_id));
    return;
  }
  const word_t addr_word = is_create ? (word_t){fr_lo128(contract), fr_hi128(contract)} : callee_word;
  const uint64_t TAGS[13] = {4, 5, 6, 7, 8, 11, 14, 18, 19, 20, 15, 16, 17};
  const word_t EXPECT[13] = {{one, fr_u64(0)}, caller_word, addr_word, zero, {cd_len, fr_u64(0)}, tx_val, zero, zero, zero, zero,
                             {one, fr_u64(0)}, {fr_u64((uint64_t)is_create), fr_u64(0)}, code_hash};
  for (int k = 0; k < 13; k++) {
    LK(cc_lookup(e, NEXT_RWC(), call_id, TAGS[k], &r), EV_BT_CTX0_UNSAT + 3 * k);
    CHECK(EV_BT_CTX0_UNSAT + 3 * k + 2, word_eq(rw_value(e, r), EXPECT[k]));
  }
  to_new_context(e, i, row, fr_u64(off), call_id, is_create, code_hash, gas_left);
#undef NEXT_RWC
}

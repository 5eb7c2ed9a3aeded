/*
 * oracle/state.c — TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
 *
 * Restates check_state_row and its per-tag helpers,
 * /root/reference/src/zkevm_specs/state_circuit.py:188-613, over the column-major matrix of
 * include/zkcheck.h.  Row = 57 cells in the order of state_circuit.Row (:63-96): rw_counter,
 * is_write, tag, id, address, field_tag, storage_key lo/hi, 10 address limbs, 32 key bytes,
 * value lo/hi, initial_value lo/hi, root lo/hi, lexicographic_ordering_selector.  Row flags:
 * bit0 value.is_word, bit1 initial_value.is_word.  Driver with wrap-around prev/next rows:
 * tests/test_state_circuit.py:17-38.  A row stops at its first failing constraint.
 * Pinned by tests/golden/state.npz.
 */
#include "common.h"
#include "../include/zk_evm_spec.h"

enum { T_RWC, T_IS_WRITE, T_TAG, T_ID, T_ADDR, T_FIELD_TAG, T_KEY_LO, T_KEY_HI, T_LIMB0, T_BYTE0 = 18,
       T_VAL_LO = 50, T_VAL_HI, T_INIT_LO, T_INIT_HI, T_ROOT_LO, T_ROOT_HI, T_SELECTOR, STATE_COLS };

typedef struct {
  const uint64_t* rows; uint64_t n; const uint8_t* flags;
  orc_index mpt_ix;
  orc_result* res;
} state_env;

#define SK(id, cond) do { if (!(cond)) { orc_fail(e->res, (id), row); return; } } while (0)
#define R(c) fr_load(ORC_CELL(e->rows, e->n, c, i))
#define P(c) fr_load(ORC_CELL(e->rows, e->n, c, ip))
#define N(c) fr_load(ORC_CELL(e->rows, e->n, c, in))

static int keys_eq(const state_env* e, uint64_t a, uint64_t b) {
  for (int c = T_TAG; c <= T_KEY_HI; c++)
    if (!fr_eq(fr_load(ORC_CELL(e->rows, e->n, c, a)), fr_load(ORC_CELL(e->rows, e->n, c, b)))) return 0;
  return 1;
}
static int word_eq2(const state_env* e, int c, uint64_t a, uint64_t b) {
  return fr_eq(fr_load(ORC_CELL(e->rows, e->n, c, a)), fr_load(ORC_CELL(e->rows, e->n, c, b))) &&
         fr_eq(fr_load(ORC_CELL(e->rows, e->n, c + 1, a)), fr_load(ORC_CELL(e->rows, e->n, c + 1, b)));
}
static int word_zero(const state_env* e, int c, uint64_t a) {
  return fr_is_zero(fr_load(ORC_CELL(e->rows, e->n, c, a))) && fr_is_zero(fr_load(ORC_CELL(e->rows, e->n, c + 1, a)));
}
/* value.value(): asserts the cell is not a Word, then returns lo */
static int is_word(const state_env* e, uint64_t a, int which) { return e->flags ? (e->flags[a] >> which) & 1 : 0; }

/* nominal ranges of the packed keys (state_circuit.py:21-39) */
static int keys_nominal(const state_env* e, uint64_t a) {
  fr_t tag = fr_load(ORC_CELL(e->rows, e->n, T_TAG, a)), id = fr_load(ORC_CELL(e->rows, e->n, T_ID, a));
  fr_t addr = fr_load(ORC_CELL(e->rows, e->n, T_ADDR, a)), ft = fr_load(ORC_CELL(e->rows, e->n, T_FIELD_TAG, a));
  if (!fr_fits_bits(tag, 4) || !fr_fits_bits(id, 28) || !fr_fits_bits(addr, 160) || !fr_fits_bits(ft, 16)) return 0;
  for (int b = 0; b < 32; b++) if (!fr_fits_bits(fr_load(ORC_CELL(e->rows, e->n, T_BYTE0 + b, a)), 8)) return 0;
  return 1;
}
/* pack = ((((tag*2^28+id)*2^160+address)*2^16+field_tag)*2^32 + key_int)*2^32 + rw_counter  (320-bit) */
typedef struct { uint64_t l[5]; } u320;
static void u320_shl_add(u320* v, int bits, const uint64_t* add, int add_limbs) {
  /* v = (v << bits) + add ; bits in {16,28,32,160} */
  u320 r = {{0, 0, 0, 0, 0}};
  int ws = bits / 64, bs = bits % 64;
  for (int k = 4; k >= 0; k--) {
    uint64_t x = 0;
    if (k - ws >= 0) x = v->l[k - ws] << bs;
    if (bs && k - ws - 1 >= 0) x |= v->l[k - ws - 1] >> (64 - bs);
    r.l[k] = x;
  }
  uint64_t c = 0;
  for (int k = 0; k < 5; k++) r.l[k] = adc(r.l[k], k < add_limbs ? add[k] : 0, &c);
  *v = r;
}
static u320 pack_keys(const state_env* e, uint64_t a) {
  u320 v = {{fr_load(ORC_CELL(e->rows, e->n, T_TAG, a)).l[0], 0, 0, 0, 0}};
  fr_t id = fr_load(ORC_CELL(e->rows, e->n, T_ID, a)), addr = fr_load(ORC_CELL(e->rows, e->n, T_ADDR, a));
  fr_t ft = fr_load(ORC_CELL(e->rows, e->n, T_FIELD_TAG, a)), rwc = fr_load(ORC_CELL(e->rows, e->n, T_RWC, a));
  u320_shl_add(&v, 28, id.l, 4);
  u320_shl_add(&v, 160, addr.l, 4);
  u320_shl_add(&v, 16, ft.l, 4);
  uint64_t key[4] = {0, 0, 0, 0};
  for (int b = 0; b < 32; b++) key[b / 8] |= fr_load(ORC_CELL(e->rows, e->n, T_BYTE0 + b, a)).l[0] << (8 * (b % 8));
  u320_shl_add(&v, 32, key, 4);
  u320_shl_add(&v, 32, rwc.l, 4);
  return v;
}
static int u320_lt(const u320* a, const u320* b) {
  for (int k = 4; k >= 0; k--) { if (a->l[k] < b->l[k]) return 1; if (a->l[k] > b->l[k]) return 0; }
  return 0;
}

static int mpt_lookup(state_env* e, uint64_t i, uint64_t ip, uint64_t proof_type) {
  fr_t key[12] = {R(T_ADDR), fr_u64(proof_type), R(T_KEY_LO), R(T_KEY_HI), R(T_ROOT_LO), R(T_ROOT_HI),
                  P(T_ROOT_LO), P(T_ROOT_HI), R(T_VAL_LO), R(T_VAL_HI), R(T_INIT_LO), R(T_INIT_HI)};
  return orc_lookup(&e->mpt_ix, key, 0);
}

static void check_row(state_env* e, uint64_t i, uint64_t row) {
  const uint64_t ip = (i + e->n - 1) % e->n, in = (i + 1) % e->n;
  const fr_t tag = R(T_TAG), id = R(T_ID), addr = R(T_ADDR), ft = R(T_FIELD_TAG), is_write = R(T_IS_WRITE);
  const fr_t rwc = R(T_RWC);
  SK(ST_TAG_RANGE, fr_fits_bits(tag, 8) && tag.l[0] >= 1 && tag.l[0] <= 12);
  SK(ST_ID_RANGE, fr_fits_bits(id, 28));
  SK(ST_FIELD_TAG_RANGE, fr_fits_bits(ft, 8) && ft.l[0] <= 24);
  fr_t acc = fr_u64(0);
  for (int k = 0; k < 10; k++) SK(ST_ADDR_LIMB_RANGE, fr_fits_bits(R(T_LIMB0 + k), 16));
  for (int k = 9; k >= 0; k--) acc = fr_add(fr_mul(acc, fr_u64(65536)), R(T_LIMB0 + k));
  SK(ST_ADDR_LIMBS, fr_eq(addr, acc));
  for (int k = 0; k < 32; k++) SK(ST_KEY_BYTE_RANGE, fr_fits_bits(R(T_BYTE0 + k), 8));
  uint64_t kb[4] = {0, 0, 0, 0};
  for (int b = 0; b < 32; b++) kb[b / 8] |= R(T_BYTE0 + b).l[0] << (8 * (b % 8));
  SK(ST_KEY_BYTES, fr_eq(R(T_KEY_LO), fr_u128(kb[0], kb[1])) && fr_eq(R(T_KEY_HI), fr_u128(kb[2], kb[3])));
  SK(ST_IS_WRITE_BOOL, fr_eq_u64(is_write, 0) || fr_eq_u64(is_write, 1));
  const uint64_t t = tag.l[0];
  for (int b = 0; b < 32; b++) SK(ST_PREV_KEY_BYTES, fr_fits_bits(P(T_BYTE0 + b), 8));
  if (t != ZK_ST_Start) {
    SK(ST_WITNESS_DOMAIN, keys_nominal(e, ip));
    u320 a = pack_keys(e, ip), b = pack_keys(e, i);
    SK(ST_LEX_ORDER, u320_lt(&a, &b));
  }
  const int same = keys_eq(e, i, ip);
  if (fr_eq_u64(is_write, 0) && same) SK(ST_READ_CONSISTENCY, word_eq2(e, T_VAL_LO, i, ip));
  if (same) SK(ST_INITIAL_CONSISTENCY, word_eq2(e, T_INIT_LO, i, ip));
  if (t != ZK_ST_Start) SK(ST_RWC_NONZERO, !fr_is_zero(rwc));
  const int key0 = word_zero(e, T_KEY_LO, i);
  const int root_same = word_eq2(e, T_ROOT_LO, i, ip);
  const int val_word = is_word(e, i, 0), init_word = is_word(e, i, 1);
  const fr_t val_lo = R(T_VAL_LO), val_hi = R(T_VAL_HI), init_lo = R(T_INIT_LO), init_hi = R(T_INIT_HI);
  const int first_read = !same && fr_eq_u64(is_write, 0);
  switch (t) {
    case ZK_ST_Start:
      SK(ST_START_FIELD_TAG0, fr_is_zero(ft)); SK(ST_START_ADDR0, fr_is_zero(addr)); SK(ST_START_ID0, fr_is_zero(id));
      SK(ST_START_KEY0, key0); SK(ST_START_VALUE_HI0, fr_is_zero(val_hi)); SK(ST_START_INIT_HI0, fr_is_zero(init_hi));
      SK(ST_START_RWC_INC, fr_is_zero(R(T_SELECTOR)) || fr_eq(rwc, fr_add(P(T_RWC), fr_u64(1))));
      SK(ST_START_VALUE0, !val_word && fr_is_zero(val_lo));
      SK(ST_START_INIT0, !init_word && fr_is_zero(init_lo));
      if (!fr_is_zero(R(T_SELECTOR))) SK(ST_START_ROOT_SAME, root_same);
      break;
    case ZK_ST_Memory:
      SK(ST_MEM_FIELD_TAG0, fr_is_zero(ft)); SK(ST_MEM_KEY0, key0);
      SK(ST_MEM_VALUE_HI0, fr_is_zero(val_hi)); SK(ST_MEM_INIT_HI0, fr_is_zero(init_hi));
      if (first_read) SK(ST_MEM_FIRST_READ0, !val_word && fr_is_zero(val_lo));
      SK(ST_MEM_ADDR_RANGE, fr_fits_bits(addr, 32));
      SK(ST_MEM_VALUE_BYTE, !val_word && fr_fits_bits(val_lo, 8));
      SK(ST_MEM_INIT0, !init_word && fr_is_zero(init_lo));
      SK(ST_MEM_ROOT_SAME, root_same);
      break;
    case ZK_ST_Stack:
      SK(ST_STK_FIELD_TAG0, fr_is_zero(ft)); SK(ST_STK_KEY0, key0);
      if (!same) SK(ST_STK_FIRST_WRITE, fr_eq_u64(is_write, 1));
      SK(ST_STK_PTR_RANGE, fr_fits_bits(addr, 16) && addr.l[0] <= 1023);
      if (fr_eq(tag, P(T_TAG)) && fr_eq(id, P(T_ID))) {
        fr_t d = fr_sub(addr, P(T_ADDR));
        SK(ST_STK_PTR_INC, fr_eq_u64(d, 0) || fr_eq_u64(d, 1));
      }
      SK(ST_STK_INIT0, fr_is_zero(init_lo) && fr_is_zero(init_hi));
      SK(ST_STK_ROOT_SAME, root_same);
      break;
    case ZK_ST_Storage: {
      SK(ST_STO_FIELD_TAG0, fr_is_zero(ft));
      const int non_exist = fr_is_zero(val_lo) && fr_is_zero(val_hi) && fr_is_zero(init_lo) && fr_is_zero(init_hi);
      if (!keys_eq(e, i, in)) {
        const int n = mpt_lookup(e, i, ip, non_exist ? ZK_MPT_NonExistingAccountProof : ZK_MPT_StorageMod);
        if (n != 1) { orc_fail(e->res, n == 0 ? ST_STO_MPT_UNSAT : ST_STO_MPT_AMBIG, row); return; }
      } else SK(ST_STO_ROOT_SAME, root_same);
      break;
    }
    case ZK_ST_CallContext:
      SK(ST_CC_ADDR0, fr_is_zero(addr)); SK(ST_CC_KEY0, key0);
      SK(ST_CC_FIELD_TAG_RANGE, fr_fits_bits(ft, 8) && ft.l[0] <= 24);
      if (first_read) SK(ST_CC_FIRST_READ0, !val_word && fr_is_zero(val_lo));
      SK(ST_CC_INIT0, fr_is_zero(init_lo) && fr_is_zero(init_hi));
      SK(ST_CC_ROOT_SAME, root_same);
      break;
    case ZK_ST_Account: {
      SK(ST_ACC_FIELD_TAG_VALUE, ft.l[0] >= 1 && ft.l[0] <= 4);
      SK(ST_ACC_ID0, fr_is_zero(id)); SK(ST_ACC_KEY0, key0);
      if (ft.l[0] == ZK_ACC_Nonce) { SK(ST_ACC_NONCE_VALUE_HI0, fr_is_zero(val_hi)); SK(ST_ACC_NONCE_INIT_HI0, fr_is_zero(init_hi)); }
      const int non_exist = fr_is_zero(val_lo) && fr_is_zero(val_hi) && fr_is_zero(init_lo) && fr_is_zero(init_hi) &&
                            ft.l[0] == ZK_ACC_CodeHash;
      if (!keys_eq(e, i, in)) {
        const int n = mpt_lookup(e, i, ip, non_exist ? ZK_MPT_NonExistingAccountProof : ft.l[0]);
        if (n != 1) { orc_fail(e->res, n == 0 ? ST_ACC_MPT_UNSAT : ST_ACC_MPT_AMBIG, row); return; }
      } else SK(ST_ACC_ROOT_SAME, root_same);
      break;
    }
    case ZK_ST_TxRefund:
      SK(ST_REF_ADDR0, fr_is_zero(addr)); SK(ST_REF_FIELD_TAG0, fr_is_zero(ft)); SK(ST_REF_KEY0, key0);
      SK(ST_REF_ROOT_SAME, root_same);
      SK(ST_REF_INIT0, fr_is_zero(init_lo) && fr_is_zero(init_hi));
      if (first_read) SK(ST_REF_FIRST_READ0, fr_is_zero(val_lo) && fr_is_zero(val_hi));
      break;
    case ZK_ST_TxAccessListAccount:
      SK(ST_ALA_FIELD_TAG0, fr_is_zero(ft)); SK(ST_ALA_KEY0, key0);
      SK(ST_ALA_VALUE_HI0, fr_is_zero(val_hi)); SK(ST_ALA_INIT_HI0, fr_is_zero(init_hi));
      SK(ST_ALA_ROOT_SAME, root_same);
      if (first_read) SK(ST_ALA_FIRST_READ0, !val_word && fr_is_zero(val_lo));
      break;
    case ZK_ST_TxAccessListAccountStorage:
      SK(ST_ALS_FIELD_TAG0, fr_is_zero(ft));
      SK(ST_ALS_VALUE_HI0, fr_is_zero(val_hi)); SK(ST_ALS_INIT_HI0, fr_is_zero(init_hi));
      SK(ST_ALS_ROOT_SAME, root_same);
      if (first_read) SK(ST_ALS_FIRST_READ0, !val_word && fr_is_zero(val_lo));
      break;
    case ZK_ST_TxLog:
      if (!fr_eq_u64(ft, ZK_LOG_Topic)) { SK(ST_LOG_VALUE_HI0, fr_is_zero(val_hi)); SK(ST_LOG_INIT_HI0, fr_is_zero(init_hi)); }
      SK(ST_LOG_IS_WRITE, fr_eq_u64(is_write, 1));
      SK(ST_LOG_ROOT_SAME, root_same);
      break;
    case ZK_ST_TxReceipt: {
      SK(ST_RCP_ADDR0, fr_is_zero(addr)); SK(ST_RCP_KEY0, key0);
      SK(ST_RCP_VALUE_HI0, fr_is_zero(val_hi)); SK(ST_RCP_INIT_HI0, fr_is_zero(init_hi));
      if (fr_eq_u64(ft, ZK_RCPT_PostStateOrStatus))
        SK(ST_RCP_STATUS_BOOL, !val_word && (fr_eq_u64(val_lo, 0) || fr_eq_u64(val_lo, 1)));
      const fr_t pid = P(T_ID);
      if (!fr_eq(id, pid) && fr_eq(tag, P(T_TAG))) {
        SK(ST_RCP_TXID_INC, fr_eq(id, fr_add(pid, fr_u64(1))));
        if (fr_eq_u64(ft, ZK_RCPT_CumulativeGasUsed))
          SK(ST_RCP_GAS_INC, !val_word && !is_word(e, ip, 0) && fr_cmp(val_lo, P(T_VAL_LO)) > 0);
      }
      if (!fr_eq(tag, P(T_TAG))) SK(ST_RCP_FIRST_TXID1, fr_eq_u64(id, 1));
      SK(ST_RCP_TXID_RANGE, fr_fits_bits(id, 16) && id.l[0] >= 1 && id.l[0] <= 2048);
      SK(ST_RCP_ROOT_SAME, root_same);
      break;
    }
    default:
      orc_fail(e->res, ST_TAG_UNREACHABLE, row);
  }
}

int orc_check_state(const uint64_t* rows, uint64_t n_rows, const uint8_t* flags, const uint64_t* mpt_tab,
                    uint64_t n_mpt, uint64_t row_begin, uint64_t row_end, uint32_t* first_fail,
                    uint64_t* fail_count) {
  orc_result res; orc_result_init(&res, first_fail, fail_count, ST_N_CONSTRAINTS);
  state_env env; env.rows = rows; env.n = n_rows; env.flags = flags; env.res = &res;
  const uint32_t k12[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
  orc_index_build(&env.mpt_ix, mpt_tab, n_mpt, 12, k12, 12);
  for (uint64_t i = row_begin; i < row_end; i++) check_row(&env, i, i);
  orc_index_free(&env.mpt_ix);
  return 0;
}

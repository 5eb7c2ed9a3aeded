/*
 * oracle/copy.c — TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
 *
 * Restates verify_row / verify_step / verify_copy_table,
 * /root/reference/src/zkevm_specs/copy_circuit.py:16-130, and the ConstraintSystem they use
 * (util/constraint_system.py:12-74: every gate is cond * expr == 0 over Fr).  Row = 20 cells
 * in the order of CopyCircuitRow (evm_circuit/table.py:472-491) with id as (lo, hi); the
 * WordOrValue.is_word bit of id and of table values travels as a per-row flag byte (bit 0).
 * A row stops at its first failing gate, like the reference's exception.
 * Pinned by tests/golden/copy.npz (verdicts of the reference's own functions).
 */
#include "common.h"
#include "../include/zk_evm_spec.h"

enum { K_QSTEP, K_FIRST, K_LAST, K_ID_LO, K_ID_HI, K_TAG, K_ADDR, K_SRC_END, K_BYTES_LEFT, K_VALUE,
       K_RLC_ACC, K_IS_CODE, K_IS_PAD, K_RWC, K_RWC_INC, K_IS_MEM, K_IS_BC, K_IS_TX, K_IS_LOG, K_IS_RLC,
       COPY_COLS };

typedef struct {
  const uint64_t* rows; uint64_t n; const uint8_t* row_flags;
  orc_index rw_ix, bc_ix, tx_ix;
  const uint8_t *rw_flags, *tx_flags;
  fr_t r;
  orc_result* res;
} copy_env;

#define CK(id, cond) do { if (!(cond)) { orc_fail(e->res, (id), row); return; } } while (0)
#define C0(c) fr_load(ORC_CELL(e->rows, e->n, c, i))
#define C1(c) fr_load(ORC_CELL(e->rows, e->n, c, j1))
#define C2(c) fr_load(ORC_CELL(e->rows, e->n, c, j2))

static int is_bool(fr_t v) { return fr_eq_u64(v, 0) || fr_eq_u64(v, 1); }
/* cond * expr == 0 over a field */
static int gate(fr_t cond, fr_t expr) { return fr_is_zero(fr_mul(cond, expr)); }

static void verify_row_and_step(copy_env* e, uint64_t i, uint64_t row) {
  const uint64_t j1 = (i + 1) % e->n, j2 = (i + 2) % e->n;
  const fr_t one = fr_u64(1);
  const fr_t q = C0(K_QSTEP), is_first = C0(K_FIRST), is_last = C0(K_LAST), tag = C0(K_TAG);
  CK(CP_BOOL_FIRST, is_bool(is_first));
  CK(CP_BOOL_LAST, is_bool(is_last));
  CK(CP_FIRST_NEEDS_STEP, gate(fr_sub(one, q), is_first));
  CK(CP_LAST_NOT_STEP, gate(q, is_last));
  CK(CP_IS_MEMORY, fr_eq(C0(K_IS_MEM), fr_u64(fr_eq_u64(tag, ZK_COPY_Memory))));
  CK(CP_IS_BYTECODE, fr_eq(C0(K_IS_BC), fr_u64(fr_eq_u64(tag, ZK_COPY_Bytecode))));
  CK(CP_IS_TX_CALLDATA, fr_eq(C0(K_IS_TX), fr_u64(fr_eq_u64(tag, ZK_COPY_TxCalldata))));
  CK(CP_IS_TX_LOG, fr_eq(C0(K_IS_LOG), fr_u64(fr_eq_u64(tag, ZK_COPY_TxLog))));
  CK(CP_IS_RLC_ACC, fr_eq(C0(K_IS_RLC), fr_u64(fr_eq_u64(tag, ZK_COPY_RlcAcc))));
  fr_t cond = fr_sub(one, fr_add(is_last, C1(K_LAST)));
  CK(CP_ID_SAME, gate(cond, fr_sub(C0(K_ID_LO), C2(K_ID_LO))) && gate(cond, fr_sub(C0(K_ID_HI), C2(K_ID_HI))));
  CK(CP_TAG_SAME, gate(cond, fr_sub(tag, C2(K_TAG))));
  CK(CP_ADDR_INC, gate(cond, fr_sub(fr_add(C0(K_ADDR), one), C2(K_ADDR))));
  CK(CP_SRC_END_SAME, gate(cond, fr_sub(C0(K_SRC_END), C2(K_SRC_END))));
  const fr_t rw_diff = fr_mul(fr_sub(one, C0(K_IS_PAD)), fr_add(C0(K_IS_MEM), C0(K_IS_LOG)));
  cond = fr_sub(one, is_last);
  CK(CP_RWC, gate(cond, fr_sub(fr_add(C0(K_RWC), rw_diff), C1(K_RWC))));
  CK(CP_RWC_INC_LEFT, gate(cond, fr_sub(fr_sub(C0(K_RWC_INC), rw_diff), C1(K_RWC_INC))));
  CK(CP_RLC_ACC_SAME, gate(cond, fr_sub(C0(K_RLC_ACC), C1(K_RLC_ACC))));
  CK(CP_RWC_INC_LAST, gate(is_last, fr_sub(C0(K_RWC_INC), rw_diff)));
  CK(CP_RLC_LAST, gate(fr_mul(is_last, C0(K_IS_RLC)), fr_sub(C0(K_RLC_ACC), C0(K_VALUE))));
  /* verify_step */
  CK(CP_BYTES_LEFT_LAST, gate(q, fr_mul(C1(K_LAST), fr_sub(one, C0(K_BYTES_LEFT)))));
  CK(CP_BYTES_LEFT_DEC,
     gate(q, fr_mul(fr_sub(one, C1(K_LAST)), fr_sub(fr_sub(C0(K_BYTES_LEFT), C2(K_BYTES_LEFT)), one))));
  CK(CP_PAD_VALUE0, gate(q, fr_mul(C0(K_IS_PAD), C0(K_VALUE))));
  if (fr_is_zero(C0(K_IS_LOG))) {
    CK(CP_LT_RANGE, fr_fits_bits(C0(K_ADDR), 40) && fr_fits_bits(C0(K_SRC_END), 40));
    const fr_t lt = fr_u64(fr_cmp(C0(K_ADDR), C0(K_SRC_END)) < 0);
    CK(CP_IS_PAD, gate(q, fr_sub(fr_sub(one, lt), C0(K_IS_PAD))));
  }
  CK(CP_NEXT_NOT_PAD, gate(q, C1(K_IS_PAD)));
  CK(CP_RW_VALUE_EQ, gate(fr_mul(q, fr_sub(one, C1(K_IS_RLC))), fr_sub(C0(K_VALUE), C1(K_VALUE))));
  CK(CP_FIRST_VALUE_EQ, gate(fr_mul(q, is_first), fr_sub(C0(K_VALUE), C1(K_VALUE))));
  CK(CP_RLC_STEP, gate(fr_mul(fr_mul(fr_sub(one, q), fr_sub(one, is_last)), C0(K_IS_RLC)),
                       fr_sub(C2(K_VALUE), fr_add(fr_mul(C0(K_VALUE), e->r), C1(K_VALUE)))));
  /* table lookups, copy_circuit.py:106-130 */
  const int id_is_word = e->row_flags ? (e->row_flags[i] & 1) : 0;
  const int not_pad = fr_eq_u64(C0(K_IS_PAD), 0);
  uint32_t hit;
  if (fr_eq_u64(C0(K_IS_MEM), 1) && not_pad) {
    CK(CP_MEM_ID_TYPE, !id_is_word);
    fr_t key[5] = {C0(K_RWC), fr_sub(one, q), fr_u64(ZK_TARGET_Memory), C0(K_ID_LO), C0(K_ADDR)};
    const int n = orc_lookup(&e->rw_ix, key, &hit);
    if (n != 1) { orc_fail(e->res, n == 0 ? CP_MEM_UNSAT : CP_MEM_AMBIG, row); return; }
    CK(CP_MEM_VALUE_TYPE, !(e->rw_flags && (e->rw_flags[hit] & 1)));
    CK(CP_MEM_VALUE, fr_eq(fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, 8, hit)), C0(K_VALUE)));
  }
  if (fr_eq_u64(C0(K_IS_BC), 1) && not_pad) {
    fr_t key[5] = {C0(K_ID_LO), C0(K_ID_HI), fr_u64(2), C0(K_ADDR), C0(K_IS_CODE)};
    const int n = orc_lookup(&e->bc_ix, key, &hit);
    if (n != 1) { orc_fail(e->res, n == 0 ? CP_BC_UNSAT : CP_BC_AMBIG, row); return; }
    CK(CP_BC_VALUE, fr_eq(fr_load(ORC_CELL(e->bc_ix.cells, e->bc_ix.n_rows, 5, hit)), C0(K_VALUE)));
  }
  if (fr_eq_u64(C0(K_IS_TX), 1) && not_pad) {
    CK(CP_TX_ID_TYPE, !id_is_word);
    fr_t key[3] = {C0(K_ID_LO), fr_u64(ZK_TX_CallData), C0(K_ADDR)};
    const int n = orc_lookup(&e->tx_ix, key, &hit);
    if (n != 1) { orc_fail(e->res, n == 0 ? CP_TX_UNSAT : CP_TX_AMBIG, row); return; }
    CK(CP_TX_VALUE_TYPE, !(e->tx_flags && (e->tx_flags[hit] & 1)));
    CK(CP_TX_VALUE, fr_eq(fr_load(ORC_CELL(e->tx_ix.cells, e->tx_ix.n_rows, 3, hit)), C0(K_VALUE)));
  }
  if (fr_eq_u64(C0(K_IS_LOG), 1)) {
    CK(CP_LOG_ID_TYPE, !id_is_word);
    fr_t key[5] = {C0(K_RWC), fr_u64(1), fr_u64(ZK_TARGET_TxLog), C0(K_ID_LO), C0(K_ADDR)};
    const int n = orc_lookup(&e->rw_ix, key, &hit);
    if (n != 1) { orc_fail(e->res, n == 0 ? CP_LOG_UNSAT : CP_LOG_AMBIG, row); return; }
    CK(CP_LOG_VALUE_TYPE, !(e->rw_flags && (e->rw_flags[hit] & 1)));
    CK(CP_LOG_VALUE, fr_eq(fr_load(ORC_CELL(e->rw_ix.cells, e->rw_ix.n_rows, 8, hit)), C0(K_VALUE)));
  }
}

int orc_check_copy(const uint64_t* rows, uint64_t n_rows, const uint8_t* row_flags, const uint64_t* rw_tab,
                   uint64_t n_rw, const uint8_t* rw_flags, const uint64_t* bytecode_tab, uint64_t n_bytecode,
                   const uint64_t* tx_tab, uint64_t n_tx, const uint8_t* tx_flags, const uint64_t r_[4],
                   uint64_t row_begin, uint64_t row_end, uint32_t* first_fail, uint64_t* fail_count) {
  orc_result res; orc_result_init(&res, first_fail, fail_count, CP_N_CONSTRAINTS);
  copy_env env; env.rows = rows; env.n = n_rows; env.row_flags = row_flags; env.rw_flags = rw_flags;
  env.tx_flags = tx_flags; env.r = fr_load(r_); env.res = &res;
  const uint32_t k5[5] = {0, 1, 2, 3, 4}, k3[3] = {0, 1, 2};
  orc_index_build(&env.rw_ix, rw_tab, n_rw, 14, k5, 5);
  orc_index_build(&env.bc_ix, bytecode_tab, n_bytecode, 6, k5, 5);
  orc_index_build(&env.tx_ix, tx_tab, n_tx, 5, k3, 3);
  for (uint64_t i = row_begin; i < row_end; i++) verify_row_and_step(&env, i, i);
  orc_index_free(&env.rw_ix); orc_index_free(&env.bc_ix); orc_index_free(&env.tx_ix);
  return 0;
}

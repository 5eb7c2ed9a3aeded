// copy.cu — copy-circuit row checker (one thread per row, warp-synchronous lookups).
//
// Replaces the loop of verify_copy_table (src/zkevm_specs/copy_circuit.py:92-130) with its
// verify_row (:23-59) and verify_step (:62-89).  Row = 20 cells in the order of CopyCircuitRow
// (evm_circuit/table.py:472-491), id as (lo, hi); rotations {0,+1,+2}.  Algorithmic bytes:
// 20 x 32 B = 640 B per row plus the table row it looks up.
// Every gate of the reference is `cond * expr == 0` over Fr (util/constraint_system.py:27-46);
// Fr is an integral domain, so a gate holds iff cond == 0 or expr == 0 and no multiplication is
// needed to decide it.  The one true Fr x Fr product is value * r in the RlcAcc gate (:89).
// A row stops at its first failing gate (the reference raises there).
#include "circuit.cuh"
#include "../../include/zk_constraints.h"
#include "../../include/zk_evm_spec.h"
#include "../../include/zkcheck.h"

namespace zk {

enum { K_QSTEP, K_FIRST, K_LAST, K_ID_LO, K_ID_HI, K_TAG, K_ADDR, K_SRC_END, K_BYTES_LEFT, K_VALUE,
       K_RLC_ACC, K_IS_CODE, K_IS_PAD, K_RWC, K_RWC_INC, K_IS_MEM, K_IS_BC, K_IS_TX, K_IS_LOG, K_IS_RLC };

struct CopyTables {
  IndexDev rw;        // key (rw_counter, rw, tag, id, address)
  IndexDev bytecode;  // key (hash_lo, hash_hi, tag, index, is_code)
  IndexDev tx;        // key (tx_id, field_tag, index)
};

ZK_HD bool fr_is_bool(const Fr& v) { return fr_fits64(v) && v.l[0] <= 1; }
// a * b over Fr with the 0/1 operands that selectors almost always are short-circuited
ZK_HD Fr fr_mul_sel(const Fr& a, const Fr& b) {
  if (fr_is_zero(a) || fr_is_zero(b)) return fr_u64(0);
  if (fr_eq_u64(a, 1)) return b;
  if (fr_eq_u64(b, 1)) return a;
  return fr_mul(a, b);
}
ZK_HD bool table_flag(const TableDev& t, u32 row, int bit) { return t.flags && ((t.flags[row] >> bit) & 1); }

#define CP_CHECK(id, cond)               \
  do {                                   \
    if (live && !(cond)) {               \
      if (record) fail(res, (id), row);  \
      live = false;                      \
    }                                    \
  } while (0)

// Warp-synchronous: every lane of `mask` calls it; lanes without a row pass live = false.
template <int LAYOUT>
ZK_HD void check_copy_row(const WitnessDev& w, const CheckRange& rg, const CopyTables& t, const Fr& r_mont,
                          const ResultDev& res, u64 i, bool live, unsigned mask) {
  const bool record = live;
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 j1 = rot_fwd(w, i, 1, wrap), j2 = rot_fwd(w, i, 2, wrap);
  const u64 row = rg.row_base + i;
  const Fr one = fr_u64(1);
  const Fr q = wcell_l<LAYOUT>(w, K_QSTEP, i), is_first = wcell_l<LAYOUT>(w, K_FIRST, i), is_last = wcell_l<LAYOUT>(w, K_LAST, i);
  const Fr tag = wcell_l<LAYOUT>(w, K_TAG, i), id_lo = wcell_l<LAYOUT>(w, K_ID_LO, i), id_hi = wcell_l<LAYOUT>(w, K_ID_HI, i);
  const Fr addr = wcell_l<LAYOUT>(w, K_ADDR, i), src_end = wcell_l<LAYOUT>(w, K_SRC_END, i), bytes_left = wcell_l<LAYOUT>(w, K_BYTES_LEFT, i);
  const Fr value = wcell_l<LAYOUT>(w, K_VALUE, i), rlc_acc = wcell_l<LAYOUT>(w, K_RLC_ACC, i), is_pad = wcell_l<LAYOUT>(w, K_IS_PAD, i);
  const Fr rwc = wcell_l<LAYOUT>(w, K_RWC, i), rwc_inc = wcell_l<LAYOUT>(w, K_RWC_INC, i);
  const Fr is_mem = wcell_l<LAYOUT>(w, K_IS_MEM, i), is_bc = wcell_l<LAYOUT>(w, K_IS_BC, i), is_tx = wcell_l<LAYOUT>(w, K_IS_TX, i);
  const Fr is_log = wcell_l<LAYOUT>(w, K_IS_LOG, i), is_rlc = wcell_l<LAYOUT>(w, K_IS_RLC, i);
  const Fr n_last = wcell_l<LAYOUT>(w, K_LAST, j1), n_value = wcell_l<LAYOUT>(w, K_VALUE, j1);
  const bool q0 = fr_is_zero(q);

  // ---- verify_row ----
  CP_CHECK(CP_BOOL_FIRST, fr_is_bool(is_first));
  CP_CHECK(CP_BOOL_LAST, fr_is_bool(is_last));
  CP_CHECK(CP_FIRST_NEEDS_STEP, fr_eq_u64(q, 1) || fr_is_zero(is_first));
  CP_CHECK(CP_LAST_NOT_STEP, q0 || fr_is_zero(is_last));
  CP_CHECK(CP_IS_MEMORY, fr_eq_u64(is_mem, fr_eq_u64(tag, ZK_COPY_Memory) ? 1 : 0));
  CP_CHECK(CP_IS_BYTECODE, fr_eq_u64(is_bc, fr_eq_u64(tag, ZK_COPY_Bytecode) ? 1 : 0));
  CP_CHECK(CP_IS_TX_CALLDATA, fr_eq_u64(is_tx, fr_eq_u64(tag, ZK_COPY_TxCalldata) ? 1 : 0));
  CP_CHECK(CP_IS_TX_LOG, fr_eq_u64(is_log, fr_eq_u64(tag, ZK_COPY_TxLog) ? 1 : 0));
  CP_CHECK(CP_IS_RLC_ACC, fr_eq_u64(is_rlc, fr_eq_u64(tag, ZK_COPY_RlcAcc) ? 1 : 0));
  {
    // cond = 1 - (is_last + next.is_last)
    const bool off = fr_eq_u64(fr_add(is_last, n_last), 1);
    CP_CHECK(CP_ID_SAME, off || (fr_eq(id_lo, wcell_l<LAYOUT>(w, K_ID_LO, j2)) && fr_eq(id_hi, wcell_l<LAYOUT>(w, K_ID_HI, j2))));
    CP_CHECK(CP_TAG_SAME, off || fr_eq(tag, wcell_l<LAYOUT>(w, K_TAG, j2)));
    CP_CHECK(CP_ADDR_INC, off || fr_eq(fr_add_u64(addr, 1), wcell_l<LAYOUT>(w, K_ADDR, j2)));
    CP_CHECK(CP_SRC_END_SAME, off || fr_eq(src_end, wcell_l<LAYOUT>(w, K_SRC_END, j2)));
  }
  const Fr rw_diff = fr_mul_sel(fr_sub(one, is_pad), fr_add(is_mem, is_log));
  {
    const bool off = fr_eq_u64(is_last, 1);  // cond = 1 - is_last
    CP_CHECK(CP_RWC, off || fr_eq(fr_add(rwc, rw_diff), wcell_l<LAYOUT>(w, K_RWC, j1)));
    CP_CHECK(CP_RWC_INC_LEFT, off || fr_eq(fr_sub(rwc_inc, rw_diff), wcell_l<LAYOUT>(w, K_RWC_INC, j1)));
    CP_CHECK(CP_RLC_ACC_SAME, off || fr_eq(rlc_acc, wcell_l<LAYOUT>(w, K_RLC_ACC, j1)));
  }
  CP_CHECK(CP_RWC_INC_LAST, fr_is_zero(is_last) || fr_eq(rwc_inc, rw_diff));
  CP_CHECK(CP_RLC_LAST, fr_is_zero(is_last) || fr_is_zero(is_rlc) || fr_eq(rlc_acc, value));
  // ---- verify_step ----
  CP_CHECK(CP_BYTES_LEFT_LAST, q0 || fr_is_zero(n_last) || fr_eq_u64(bytes_left, 1));
  CP_CHECK(CP_BYTES_LEFT_DEC,
           q0 || fr_eq_u64(n_last, 1) || fr_eq(bytes_left, fr_add_u64(wcell_l<LAYOUT>(w, K_BYTES_LEFT, j2), 1)));
  CP_CHECK(CP_PAD_VALUE0, q0 || fr_is_zero(is_pad) || fr_is_zero(value));
  if (fr_is_zero(is_log)) {
    // lt(addr, src_addr_end, 5) is evaluated (and range-asserts) whatever q_step is
    const u64 kMax = 1ull << 40;
    CP_CHECK(CP_LT_RANGE, fr_fits64(addr) && addr.l[0] < kMax && fr_fits64(src_end) && src_end.l[0] < kMax);
    const Fr want_pad = fr_u64(addr.l[0] < src_end.l[0] ? 0 : 1);  // 1 - lt
    CP_CHECK(CP_IS_PAD, q0 || fr_eq(is_pad, want_pad));
  }
  CP_CHECK(CP_NEXT_NOT_PAD, q0 || fr_is_zero(wcell_l<LAYOUT>(w, K_IS_PAD, j1)));
  CP_CHECK(CP_RW_VALUE_EQ, q0 || fr_eq_u64(wcell_l<LAYOUT>(w, K_IS_RLC, j1), 1) || fr_eq(value, n_value));
  CP_CHECK(CP_FIRST_VALUE_EQ, q0 || fr_is_zero(is_first) || fr_eq(value, n_value));
  if (live && !fr_eq_u64(q, 1) && !fr_eq_u64(is_last, 1) && !fr_is_zero(is_rlc)) {
    // next_write_value == write_value * r + next_read_value
    CP_CHECK(CP_RLC_STEP, fr_eq(wcell_l<LAYOUT>(w, K_VALUE, j2), fr_add(fr_montmul(value, r_mont), n_value)));
  }
  // ---- table lookups (copy_circuit.py:106-130), one warp-wide probe per table use ----
  const bool id_is_word = w.flags && (w.flags[i] & 1);
  const bool not_pad = fr_is_zero(is_pad);
  u32 hit = 0;
  {
    const bool need = live && fr_eq_u64(is_mem, 1) && not_pad;
    if (need) CP_CHECK(CP_MEM_ID_TYPE, !id_is_word);
    const bool go = need && live;
    Fr key[5] = {rwc, fr_sub(one, q), fr_u64(ZK_TARGET_Memory), id_lo, addr};
    const int n = lookup_sync<5>(t.rw, key, &hit, mask, go);
    if (go) {
      CP_CHECK(n == 0 ? CP_MEM_UNSAT : CP_MEM_AMBIG, n == 1);
      if (live) {
        CP_CHECK(CP_MEM_VALUE_TYPE, !table_flag(t.rw.tab, hit, 0));
        CP_CHECK(CP_MEM_VALUE, fr_eq(table_cell(t.rw.tab, 8, hit), value));
      }
    }
  }
  {
    const bool go = live && fr_eq_u64(is_bc, 1) && not_pad;
    Fr key[5] = {id_lo, id_hi, fr_u64(2), addr, wcell_l<LAYOUT>(w, K_IS_CODE, i)};
    const int n = lookup_sync<5>(t.bytecode, key, &hit, mask, go);
    if (go) {
      CP_CHECK(n == 0 ? CP_BC_UNSAT : CP_BC_AMBIG, n == 1);
      if (live) CP_CHECK(CP_BC_VALUE, fr_eq(table_cell(t.bytecode.tab, 5, hit), value));
    }
  }
  {
    const bool need = live && fr_eq_u64(is_tx, 1) && not_pad;
    if (need) CP_CHECK(CP_TX_ID_TYPE, !id_is_word);
    const bool go = need && live;
    Fr key[3] = {id_lo, fr_u64(ZK_TX_CallData), addr};
    const int n = lookup_sync<3>(t.tx, key, &hit, mask, go);
    if (go) {
      CP_CHECK(n == 0 ? CP_TX_UNSAT : CP_TX_AMBIG, n == 1);
      if (live) {
        CP_CHECK(CP_TX_VALUE_TYPE, !table_flag(t.tx.tab, hit, 0));
        CP_CHECK(CP_TX_VALUE, fr_eq(table_cell(t.tx.tab, 3, hit), value));
      }
    }
  }
  {
    const bool need = live && fr_eq_u64(is_log, 1);
    if (need) CP_CHECK(CP_LOG_ID_TYPE, !id_is_word);
    const bool go = need && live;
    Fr key[5] = {rwc, one, fr_u64(ZK_TARGET_TxLog), id_lo, addr};
    const int n = lookup_sync<5>(t.rw, key, &hit, mask, go);
    if (go) {
      CP_CHECK(n == 0 ? CP_LOG_UNSAT : CP_LOG_AMBIG, n == 1);
      if (live) {
        CP_CHECK(CP_LOG_VALUE_TYPE, !table_flag(t.rw.tab, hit, 0));
        CP_CHECK(CP_LOG_VALUE, fr_eq(table_cell(t.rw.tab, 8, hit), value));
      }
    }
  }
}

#ifdef __CUDACC__
template <int LAYOUT>
__global__ void __launch_bounds__(128, 4) k_check_copy(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ CopyTables t, const __grid_constant__ Fr r_mont,
             const __grid_constant__ ResultDev res) {
  const u64 n = rg.row_end - rg.row_begin;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (u64 first = 0; first < n; first += stride) {  // warp-uniform trip count
    const u64 k = first + tid;
    const bool live = k < n;
    check_copy_row<LAYOUT>(w, rg, t, r_mont, res, rg.row_begin + (live ? k : 0), live, 0xFFFFFFFFu);
  }
}
#endif

}  // namespace zk

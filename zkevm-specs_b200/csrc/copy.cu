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
// This is the GENERAL form: every cell a field element.  Warps whose rows hold only small integers in the integer-typed
// cells run check_copy_row_small below instead (same verdicts, a third of the registers).
template <int LAYOUT>
ZK_HD_NOINLINE void check_copy_row(const WitnessDev& w, const CheckRange& rg, const CopyTables& t, const Fr& r_mont,
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

// ---- the common case: flags are 0 / 1 and counters, addresses and tags are small integers ------------------------
// A cell that is an integer below 2^62: such cells compare, add and subtract as machine integers exactly like they do
// in the field (no wrap on either side).  check_copy_rows takes this path for a warp only if EVERY integer-typed cell
// its rows touch passes ld_small / ld_flag; otherwise the whole warp runs the general form above, so the verdict never
// depends on which path ran (tests/test_emu_parity.py runs the goldens' corruptions through the dispatcher).
template <int LAYOUT>
ZK_HD bool ld_small(const WitnessDev& w, u32 col, u64 row, u64* v) {
  const Fr c = wcell_l<LAYOUT>(w, col, row);
  *v = c.l[0];
  return (c.l[1] | c.l[2] | c.l[3]) == 0 && c.l[0] < (1ull << 62);
}
template <int LAYOUT>
ZK_HD bool ld_flag(const WitnessDev& w, u32 col, u64 row, bool* v) {
  const Fr c = wcell_l<LAYOUT>(w, col, row);
  *v = c.l[0] != 0;
  return (c.l[1] | c.l[2] | c.l[3]) == 0 && c.l[0] <= 1;
}
struct CopySmall {  // the integer-typed cells of rows i, i+1, i+2 that verify_row / verify_step read
  u64 tag, addr, src_end, bytes_left, rwc, rwc_inc;      // row i
  u64 n_rwc, n_rwc_inc;                                  // row i+1
  u64 nn_tag, nn_addr, nn_src_end, nn_bytes_left;        // row i+2
  bool q, is_first, is_last, is_pad, is_mem, is_bc, is_tx, is_log, is_rlc;  // row i
  bool n_last, n_pad, n_rlc;                             // row i+1
};
template <int LAYOUT>
ZK_HD bool load_copy_small(const WitnessDev& w, u64 i, u64 j1, u64 j2, CopySmall* c) {
  bool ok = ld_flag<LAYOUT>(w, K_QSTEP, i, &c->q);
  ok &= ld_flag<LAYOUT>(w, K_FIRST, i, &c->is_first);
  ok &= ld_flag<LAYOUT>(w, K_LAST, i, &c->is_last);
  ok &= ld_small<LAYOUT>(w, K_TAG, i, &c->tag);
  ok &= ld_small<LAYOUT>(w, K_ADDR, i, &c->addr);
  ok &= ld_small<LAYOUT>(w, K_SRC_END, i, &c->src_end);
  ok &= ld_small<LAYOUT>(w, K_BYTES_LEFT, i, &c->bytes_left);
  ok &= ld_flag<LAYOUT>(w, K_IS_PAD, i, &c->is_pad);
  ok &= ld_small<LAYOUT>(w, K_RWC, i, &c->rwc);
  ok &= ld_small<LAYOUT>(w, K_RWC_INC, i, &c->rwc_inc);
  ok &= ld_flag<LAYOUT>(w, K_IS_MEM, i, &c->is_mem);
  ok &= ld_flag<LAYOUT>(w, K_IS_BC, i, &c->is_bc);
  ok &= ld_flag<LAYOUT>(w, K_IS_TX, i, &c->is_tx);
  ok &= ld_flag<LAYOUT>(w, K_IS_LOG, i, &c->is_log);
  ok &= ld_flag<LAYOUT>(w, K_IS_RLC, i, &c->is_rlc);
  ok &= ld_flag<LAYOUT>(w, K_LAST, j1, &c->n_last);
  ok &= ld_flag<LAYOUT>(w, K_IS_PAD, j1, &c->n_pad);
  ok &= ld_flag<LAYOUT>(w, K_IS_RLC, j1, &c->n_rlc);
  ok &= ld_small<LAYOUT>(w, K_RWC, j1, &c->n_rwc);
  ok &= ld_small<LAYOUT>(w, K_RWC_INC, j1, &c->n_rwc_inc);
  ok &= ld_small<LAYOUT>(w, K_TAG, j2, &c->nn_tag);
  ok &= ld_small<LAYOUT>(w, K_ADDR, j2, &c->nn_addr);
  ok &= ld_small<LAYOUT>(w, K_SRC_END, j2, &c->nn_src_end);
  ok &= ld_small<LAYOUT>(w, K_BYTES_LEFT, j2, &c->nn_bytes_left);
  return ok;
}

// same gates, same order, same ids as check_copy_row, on the cells of CopySmall
template <int LAYOUT>
ZK_HD void check_copy_row_small(const WitnessDev& w, const CheckRange& rg, const CopyTables& t, const Fr& r_mont,
                                const ResultDev& res, u64 i, u64 j1, u64 j2, const CopySmall& c, bool live, unsigned mask) {
  const bool record = live;
  const u64 row = rg.row_base + i;
  const bool q0 = !c.q;
  // ---- verify_row ---- (CP_BOOL_FIRST / CP_BOOL_LAST hold: the flags are 0 / 1)
  CP_CHECK(CP_FIRST_NEEDS_STEP, c.q || !c.is_first);
  CP_CHECK(CP_LAST_NOT_STEP, q0 || !c.is_last);
  CP_CHECK(CP_IS_MEMORY, c.is_mem == (c.tag == ZK_COPY_Memory));
  CP_CHECK(CP_IS_BYTECODE, c.is_bc == (c.tag == ZK_COPY_Bytecode));
  CP_CHECK(CP_IS_TX_CALLDATA, c.is_tx == (c.tag == ZK_COPY_TxCalldata));
  CP_CHECK(CP_IS_TX_LOG, c.is_log == (c.tag == ZK_COPY_TxLog));
  CP_CHECK(CP_IS_RLC_ACC, c.is_rlc == (c.tag == ZK_COPY_RlcAcc));
  const Fr id_lo = wcell_l<LAYOUT>(w, K_ID_LO, i), id_hi = wcell_l<LAYOUT>(w, K_ID_HI, i);
  {
    const bool off = c.is_last != c.n_last;  // is_last + next.is_last == 1
    CP_CHECK(CP_ID_SAME, off || (fr_eq(id_lo, wcell_l<LAYOUT>(w, K_ID_LO, j2)) && fr_eq(id_hi, wcell_l<LAYOUT>(w, K_ID_HI, j2))));
    CP_CHECK(CP_TAG_SAME, off || c.tag == c.nn_tag);
    CP_CHECK(CP_ADDR_INC, off || c.addr + 1 == c.nn_addr);
    CP_CHECK(CP_SRC_END_SAME, off || c.src_end == c.nn_src_end);
  }
  const u64 rw_diff = c.is_pad ? 0 : (u64)c.is_mem + (u64)c.is_log;  // (1 - is_pad) * (is_memory + is_tx_log)
  const Fr value = wcell_l<LAYOUT>(w, K_VALUE, i), rlc_acc = wcell_l<LAYOUT>(w, K_RLC_ACC, i);
  {
    const bool off = c.is_last;
    CP_CHECK(CP_RWC, off || c.rwc + rw_diff == c.n_rwc);
    CP_CHECK(CP_RWC_INC_LEFT, off || (c.rwc_inc >= rw_diff && c.rwc_inc - rw_diff == c.n_rwc_inc));
    CP_CHECK(CP_RLC_ACC_SAME, off || fr_eq(rlc_acc, wcell_l<LAYOUT>(w, K_RLC_ACC, j1)));
  }
  CP_CHECK(CP_RWC_INC_LAST, !c.is_last || c.rwc_inc == rw_diff);
  CP_CHECK(CP_RLC_LAST, !c.is_last || !c.is_rlc || fr_eq(rlc_acc, value));
  // ---- verify_step ----
  CP_CHECK(CP_BYTES_LEFT_LAST, q0 || !c.n_last || c.bytes_left == 1);
  CP_CHECK(CP_BYTES_LEFT_DEC, q0 || c.n_last || c.bytes_left == c.nn_bytes_left + 1);
  CP_CHECK(CP_PAD_VALUE0, q0 || !c.is_pad || fr_is_zero(value));
  if (!c.is_log) {
    const u64 kMax = 1ull << 40;
    CP_CHECK(CP_LT_RANGE, c.addr < kMax && c.src_end < kMax);
    CP_CHECK(CP_IS_PAD, q0 || c.is_pad == !(c.addr < c.src_end));
  }
  CP_CHECK(CP_NEXT_NOT_PAD, q0 || !c.n_pad);
  const Fr n_value = wcell_l<LAYOUT>(w, K_VALUE, j1);
  CP_CHECK(CP_RW_VALUE_EQ, q0 || c.n_rlc || fr_eq(value, n_value));
  CP_CHECK(CP_FIRST_VALUE_EQ, q0 || !c.is_first || fr_eq(value, n_value));
  if (live && q0 && !c.is_last && c.is_rlc) {
    CP_CHECK(CP_RLC_STEP, fr_eq(wcell_l<LAYOUT>(w, K_VALUE, j2), fr_add(fr_montmul(value, r_mont), n_value)));
  }
  // ---- table lookups (copy_circuit.py:106-130) ----
  const bool id_is_word = w.flags && (w.flags[i] & 1);
  const Fr addr = fr_u64(c.addr), rwc = fr_u64(c.rwc);
  u32 hit = 0;
  {
    const bool need = live && c.is_mem && !c.is_pad;
    if (need) CP_CHECK(CP_MEM_ID_TYPE, !id_is_word);
    const bool go = need && live;
    Fr key[5] = {rwc, fr_u64(q0 ? 1 : 0), fr_u64(ZK_TARGET_Memory), id_lo, addr};
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
    const bool go = live && c.is_bc && !c.is_pad;
    Fr key[5] = {id_lo, id_hi, fr_u64(2), addr, go ? wcell_l<LAYOUT>(w, K_IS_CODE, i) : fr_u64(0)};
    const int n = lookup_sync<5>(t.bytecode, key, &hit, mask, go);
    if (go) {
      CP_CHECK(n == 0 ? CP_BC_UNSAT : CP_BC_AMBIG, n == 1);
      if (live) CP_CHECK(CP_BC_VALUE, fr_eq(table_cell(t.bytecode.tab, 5, hit), value));
    }
  }
  {
    const bool need = live && c.is_tx && !c.is_pad;
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
    const bool need = live && c.is_log;
    if (need) CP_CHECK(CP_LOG_ID_TYPE, !id_is_word);
    const bool go = need && live;
    Fr key[5] = {rwc, fr_u64(1), fr_u64(ZK_TARGET_TxLog), id_lo, addr};
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

// one row per lane; the warp picks ONE form for all its lanes (the lookups are warp-synchronous)
template <int LAYOUT>
ZK_HD void check_copy_rows(const WitnessDev& w, const CheckRange& rg, const CopyTables& t, const Fr& r_mont,
                           const ResultDev& res, u64 i, bool live, unsigned mask) {
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 j1 = rot_fwd(w, i, 1, wrap), j2 = rot_fwd(w, i, 2, wrap);
  CopySmall c;
  const bool ok = load_copy_small<LAYOUT>(w, i, j1, j2, &c) || !live;
#ifdef __CUDA_ARCH__
  const bool all_ok = __all_sync(mask, ok);
#else
  const bool all_ok = ok;
#endif
  if (all_ok) check_copy_row_small<LAYOUT>(w, rg, t, r_mont, res, i, j1, j2, c, live, mask);
  else check_copy_row<LAYOUT>(w, rg, t, r_mont, res, i, live, mask);
}

#ifdef __CUDACC__
#ifndef ZK_COPY_MINBLOCKS
#define ZK_COPY_MINBLOCKS 4
#endif
// Two kernels so that the common case does not pay the general form's registers.  k_check_copy_small: one row per
// lane; a warp whose rows all hold small integers where integers belong runs check_copy_row_small, any other warp
// appends its first row to `slow` and moves on.  k_check_copy_general then runs the general form over the listed warps
// (it finds the list empty on a well-formed witness and returns).
struct CopySlowList {
  u32* count;  // [1], cleared before the launch
  u32* rows;   // first row (relative to row_begin) of every deferred warp
};
template <int LAYOUT>
__global__ void __launch_bounds__(128, ZK_COPY_MINBLOCKS)
k_check_copy_small(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ CopyTables t,
                   const __grid_constant__ Fr r_mont, const __grid_constant__ ResultDev res, const __grid_constant__ CopySlowList slow) {
  const u64 n = rg.row_end - rg.row_begin;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  for (u64 first = 0; first < n; first += stride) {  // warp-uniform trip count
    const u64 k = first + tid;
    const bool live = k < n;
    const u64 i = rg.row_begin + (live ? k : 0);
    const u64 j1 = rot_fwd(w, i, 1, wrap), j2 = rot_fwd(w, i, 2, wrap);
    CopySmall c;
    const bool ok = load_copy_small<LAYOUT>(w, i, j1, j2, &c) || !live;
    if (__all_sync(0xFFFFFFFFu, ok)) {
      check_copy_row_small<LAYOUT>(w, rg, t, r_mont, res, i, j1, j2, c, live, 0xFFFFFFFFu);
    } else if ((threadIdx.x & 31) == 0) {
      slow.rows[atomicAdd(slow.count, 1u)] = (u32)k;
    }
  }
}
template <int LAYOUT>
__global__ void __launch_bounds__(128, 3)
k_check_copy_general(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ CopyTables t,
                     const __grid_constant__ Fr r_mont, const __grid_constant__ ResultDev res, const __grid_constant__ CopySlowList slow) {
  const u64 n = rg.row_end - rg.row_begin;
  const u32 count = ld_u32(slow.count);
  const u32 warps = (gridDim.x * blockDim.x) >> 5, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  for (u32 e = warp; e < count; e += warps) {
    const u64 k = (u64)slow.rows[e] + lane;
    const bool live = k < n;
    check_copy_row<LAYOUT>(w, rg, t, r_mont, res, rg.row_begin + (live ? k : 0), live, 0xFFFFFFFFu);
  }
}
#endif

}  // namespace zk

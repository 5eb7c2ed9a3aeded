// state.cu — state-circuit row checker (one thread per row, warp-synchronous MPT lookup).
//
// Replaces the per-row loop around check_state_row and its 11 per-tag helpers,
// src/zkevm_specs/state_circuit.py:188-613 (driver: tests/test_state_circuit.py:17-38).
// Row = 57 cells in the order of state_circuit.Row (:63-96); rotations {-1, 0, +1}; row flags
// bit0 = value.is_word, bit1 = initial_value.is_word (WordOrValue, arithmetic.py:171-189).
// Algorithmic bytes: 57 x 32 B = 1,824 B per row (+ 384 B per MPT-table row).  No Fr x Fr
// product at all: every multiplication of the reference is by a constant or a selector, the
// lexicographic ordering check is a 320-bit integer compare of the packed keys (:552-570).
// A row stops at its first failing constraint.
#include "circuit.cuh"
#include "../../include/zk_constraints.h"
#include "../../include/zk_evm_spec.h"
#include "../../include/zkcheck.h"

namespace zk {

enum { T_RWC, T_IS_WRITE, T_TAG, T_ID, T_ADDR, T_FIELD_TAG, T_KEY_LO, T_KEY_HI, T_LIMB0, T_BYTE0 = 18,
       T_VAL_LO = 50, T_VAL_HI, T_INIT_LO, T_INIT_HI, T_ROOT_LO, T_ROOT_HI, T_SELECTOR };

struct U320 {
  u64 l[5];
};
// acc += v << SHIFT  (v: 4 limbs; bits beyond 320 cannot occur for in-range keys)
template <int SHIFT>
ZK_HD void u320_add_shifted(U320& acc, const Fr& v) {
  constexpr int ws = SHIFT / 64, bs = SHIFT % 64;
  u64 c = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    u64 x = 0;
    if (k - ws >= 0 && k - ws < 4) x = v.l[k - ws] << bs;
    if (bs != 0 && k - ws - 1 >= 0 && k - ws - 1 < 4) x |= v.l[k - ws - 1] >> (64 - bs);
    acc.l[k] = adc64(acc.l[k], x, c);
  }
}
ZK_HD bool u320_lt(const U320& a, const U320& b) {
#pragma unroll
  for (int k = 4; k >= 0; k--) {
    if (a.l[k] != b.l[k]) return a.l[k] < b.l[k];
  }
  return false;
}
// 32 byte cells -> 256-bit integer; false if a byte is >= 256.  Eight cells (one 64-bit limb) at a time:
// eight independent 32-byte loads in flight, then folded, so the live registers stay bounded.
template <int LAYOUT>
ZK_HD bool gather_key_bytes(const WitnessDev& w, u64 row, Fr* out) {
  bool ok = true;
  u64 l0 = 0, l1 = 0, l2 = 0, l3 = 0;
#pragma unroll 1
  for (int q = 0; q < 4; q++) {  // not unrolled: the compiler would hoist all 32 loads (256 registers)
    Fr c[8];
#pragma unroll
    for (int b = 0; b < 8; b++) c[b] = wcell_l<LAYOUT>(w, T_BYTE0 + 8 * q + b, row);
    u64 limb = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      ok = ok && fr_fits64(c[b]) && c[b].l[0] < 256;
      limb |= (c[b].l[0] & 0xFF) << (8 * b);
    }
    l0 = q == 0 ? limb : l0;
    l1 = q == 1 ? limb : l1;
    l2 = q == 2 ? limb : l2;
    l3 = q == 3 ? limb : l3;
  }
  out->l[0] = l0;
  out->l[1] = l1;
  out->l[2] = l2;
  out->l[3] = l3;
  return ok;
}
// ((((tag*2^28+id)*2^160+address)*2^16+field_tag)*2^32 + key)*2^32 + rw_counter, as
// tag<<268 + id<<240 + address<<80 + field_tag<<64 + key<<32 + rw_counter  (:552-565: the storage
// key overlaps the higher fields, reproduced as written)
ZK_HD U320 pack_keys(const Fr& tag, const Fr& id, const Fr& addr, const Fr& ft, const Fr& key, const Fr& rwc) {
  U320 v{{rwc.l[0], rwc.l[1], rwc.l[2], rwc.l[3], 0}};
  u320_add_shifted<32>(v, key);
  u320_add_shifted<64>(v, ft);
  u320_add_shifted<80>(v, addr);
  u320_add_shifted<240>(v, id);
  u320_add_shifted<268>(v, tag);
  return v;
}
ZK_HD bool fits_bits(const Fr& a, int bits) {  // bits in (0, 256)
  const int w = bits >> 6, s = bits & 63;
  for (int k = 3; k > w; k--)
    if (a.l[k]) return false;
  return s == 0 ? (w < 4 ? a.l[w] == 0 : true) : (a.l[w] >> s) == 0;
}

#define ST_CHECK(id, cond)              \
  do {                                  \
    if (live && !(cond)) {              \
      if (record) fail(res, (id), row); \
      live = false;                     \
    }                                   \
  } while (0)

// Warp-synchronous: every lane of `mask` calls it; lanes without a row pass live = false.
// 10 address limb cells -> 160-bit integer (Sum limb_i * 2^(16 i), :507-509); false if a limb is >= 2^16
template <int LAYOUT>
ZK_HD bool gather_addr_limbs(const WitnessDev& w, u64 row, Fr* sum) {
  bool ok = true;
  Fr c[10];
#pragma unroll
  for (int k = 0; k < 10; k++) c[k] = wcell_l<LAYOUT>(w, T_LIMB0 + k, row);
  *sum = fr_u64(0);
#pragma unroll
  for (int k = 0; k < 10; k++) {
    ok = ok && fr_fits64(c[k]) && c[k].l[0] < 65536;
    const int bit = 16 * k;
    sum->l[bit >> 6] |= (c[k].l[0] & 0xFFFF) << (bit & 63);
  }
  return ok;
}
// `key_int` / `bytes_ok`: gather_key_bytes of row i; `p_key_int` / `p_bytes_ok`: of the previous row (the
// kernel hands the previous row's over through shared memory instead of gathering 32 cells twice)
template <int LAYOUT>
ZK_HD void check_state_row_core(const WitnessDev& w, const CheckRange& rg, const IndexDev& mpt, const ResultDev& res,
                                u64 i, bool live, unsigned mask, const Fr& key_int, bool bytes_ok, const Fr& p_key_int,
                                bool p_bytes_ok, const Fr& limb_sum, bool limbs_ok) {
  const bool record = live;
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 ip = rot_back(w, i, wrap), in = rot_fwd(w, i, 1, wrap);
  const u64 row = rg.row_base + i;
  const Fr rwc = wcell_l<LAYOUT>(w, T_RWC, i), is_write = wcell_l<LAYOUT>(w, T_IS_WRITE, i), tag = wcell_l<LAYOUT>(w, T_TAG, i);
  const Fr id = wcell_l<LAYOUT>(w, T_ID, i), addr = wcell_l<LAYOUT>(w, T_ADDR, i), ft = wcell_l<LAYOUT>(w, T_FIELD_TAG, i);
  const Fr key_lo = wcell_l<LAYOUT>(w, T_KEY_LO, i), key_hi = wcell_l<LAYOUT>(w, T_KEY_HI, i);
  const Fr p_tag = wcell_l<LAYOUT>(w, T_TAG, ip), p_id = wcell_l<LAYOUT>(w, T_ID, ip), p_addr = wcell_l<LAYOUT>(w, T_ADDR, ip);
  const Fr p_ft = wcell_l<LAYOUT>(w, T_FIELD_TAG, ip);

  ST_CHECK(ST_TAG_RANGE, fr_fits64(tag) && tag.l[0] >= 1 && tag.l[0] <= 12);
  ST_CHECK(ST_ID_RANGE, fits_bits(id, 28));
  ST_CHECK(ST_FIELD_TAG_RANGE, fr_fits64(ft) && ft.l[0] <= 24);
  ST_CHECK(ST_ADDR_LIMB_RANGE, limbs_ok);
  ST_CHECK(ST_ADDR_LIMBS, fr_eq(addr, limb_sum));
  ST_CHECK(ST_KEY_BYTE_RANGE, bytes_ok);
  ST_CHECK(ST_KEY_BYTES, fr_eq(key_lo, fr_u128(key_int.l[0], key_int.l[1])) &&
                             fr_eq(key_hi, fr_u128(key_int.l[2], key_int.l[3])));
  ST_CHECK(ST_IS_WRITE_BOOL, fr_fits64(is_write) && is_write.l[0] <= 1);
  ST_CHECK(ST_PREV_KEY_BYTES, p_bytes_ok);
  const u64 t = tag.l[0];
  const bool is_start = t == ZK_ST_Start;
  if (live && !is_start) {
    ST_CHECK(ST_WITNESS_DOMAIN, fits_bits(p_tag, 4) && fits_bits(p_id, 28) && fits_bits(p_addr, 160) && fits_bits(p_ft, 16));
    if (live) {
      const U320 a = pack_keys(p_tag, p_id, p_addr, p_ft, p_key_int, wcell_l<LAYOUT>(w, T_RWC, ip));
      const U320 b = pack_keys(tag, id, addr, ft, key_int, rwc);
      ST_CHECK(ST_LEX_ORDER, u320_lt(a, b));
    }
  }
  const bool same = fr_eq(tag, p_tag) && fr_eq(id, p_id) && fr_eq(addr, p_addr) && fr_eq(ft, p_ft) &&
                    fr_eq(key_lo, wcell_l<LAYOUT>(w, T_KEY_LO, ip)) && fr_eq(key_hi, wcell_l<LAYOUT>(w, T_KEY_HI, ip));
  const Fr val_lo = wcell_l<LAYOUT>(w, T_VAL_LO, i), val_hi = wcell_l<LAYOUT>(w, T_VAL_HI, i);
  const Fr init_lo = wcell_l<LAYOUT>(w, T_INIT_LO, i), init_hi = wcell_l<LAYOUT>(w, T_INIT_HI, i);
  const bool read = fr_is_zero(is_write);
  if (read && same)
    ST_CHECK(ST_READ_CONSISTENCY, fr_eq(val_lo, wcell_l<LAYOUT>(w, T_VAL_LO, ip)) && fr_eq(val_hi, wcell_l<LAYOUT>(w, T_VAL_HI, ip)));
  if (same)
    ST_CHECK(ST_INITIAL_CONSISTENCY, fr_eq(init_lo, wcell_l<LAYOUT>(w, T_INIT_LO, ip)) && fr_eq(init_hi, wcell_l<LAYOUT>(w, T_INIT_HI, ip)));
  if (!is_start) ST_CHECK(ST_RWC_NONZERO, !fr_is_zero(rwc));

  const bool key0 = fr_is_zero(key_lo) && fr_is_zero(key_hi);
  const bool root_same = fr_eq(wcell_l<LAYOUT>(w, T_ROOT_LO, i), wcell_l<LAYOUT>(w, T_ROOT_LO, ip)) &&
                         fr_eq(wcell_l<LAYOUT>(w, T_ROOT_HI, i), wcell_l<LAYOUT>(w, T_ROOT_HI, ip));
  const unsigned char fl = w.flags ? w.flags[i] : 0;
  const bool val_word = fl & 1, init_word = fl & 2;
  const bool first_read = !same && read;
  const bool val0 = fr_is_zero(val_lo) && fr_is_zero(val_hi), init0 = fr_is_zero(init_lo) && fr_is_zero(init_hi);
  const bool ft0 = fr_is_zero(ft), addr0 = fr_is_zero(addr), id0 = fr_is_zero(id);
  // keys of the next row differ? (last access of Storage / Account keys)
  const bool next_same = fr_eq(tag, wcell_l<LAYOUT>(w, T_TAG, in)) && fr_eq(id, wcell_l<LAYOUT>(w, T_ID, in)) &&
                         fr_eq(addr, wcell_l<LAYOUT>(w, T_ADDR, in)) && fr_eq(ft, wcell_l<LAYOUT>(w, T_FIELD_TAG, in)) &&
                         fr_eq(key_lo, wcell_l<LAYOUT>(w, T_KEY_LO, in)) && fr_eq(key_hi, wcell_l<LAYOUT>(w, T_KEY_HI, in));
  bool need_mpt = false;
  u64 proof_type = 0;
  int mpt_unsat_id = ST_STO_MPT_UNSAT;
  if (live) {
    switch (t) {
      case ZK_ST_Start: {
        const Fr sel = wcell_l<LAYOUT>(w, T_SELECTOR, i);
        ST_CHECK(ST_START_FIELD_TAG0, ft0);
        ST_CHECK(ST_START_ADDR0, addr0);
        ST_CHECK(ST_START_ID0, id0);
        ST_CHECK(ST_START_KEY0, key0);
        ST_CHECK(ST_START_VALUE_HI0, fr_is_zero(val_hi));
        ST_CHECK(ST_START_INIT_HI0, fr_is_zero(init_hi));
        ST_CHECK(ST_START_RWC_INC, fr_is_zero(sel) || fr_eq(rwc, fr_add_u64(wcell_l<LAYOUT>(w, T_RWC, ip), 1)));
        ST_CHECK(ST_START_VALUE0, !val_word && fr_is_zero(val_lo));
        ST_CHECK(ST_START_INIT0, !init_word && fr_is_zero(init_lo));
        if (!fr_is_zero(sel)) ST_CHECK(ST_START_ROOT_SAME, root_same);
        break;
      }
      case ZK_ST_Memory:
        ST_CHECK(ST_MEM_FIELD_TAG0, ft0);
        ST_CHECK(ST_MEM_KEY0, key0);
        ST_CHECK(ST_MEM_VALUE_HI0, fr_is_zero(val_hi));
        ST_CHECK(ST_MEM_INIT_HI0, fr_is_zero(init_hi));
        if (first_read) ST_CHECK(ST_MEM_FIRST_READ0, !val_word && fr_is_zero(val_lo));
        ST_CHECK(ST_MEM_ADDR_RANGE, fits_bits(addr, 32));
        ST_CHECK(ST_MEM_VALUE_BYTE, !val_word && fr_fits64(val_lo) && val_lo.l[0] < 256);
        ST_CHECK(ST_MEM_INIT0, !init_word && fr_is_zero(init_lo));
        ST_CHECK(ST_MEM_ROOT_SAME, root_same);
        break;
      case ZK_ST_Stack:
        ST_CHECK(ST_STK_FIELD_TAG0, ft0);
        ST_CHECK(ST_STK_KEY0, key0);
        if (!same) ST_CHECK(ST_STK_FIRST_WRITE, fr_eq_u64(is_write, 1));
        ST_CHECK(ST_STK_PTR_RANGE, fr_fits64(addr) && addr.l[0] <= 1023);
        if (fr_eq(tag, p_tag) && fr_eq(id, p_id)) {
          const Fr d = fr_sub(addr, p_addr);
          ST_CHECK(ST_STK_PTR_INC, fr_fits64(d) && d.l[0] <= 1);
        }
        ST_CHECK(ST_STK_INIT0, init0);
        ST_CHECK(ST_STK_ROOT_SAME, root_same);
        break;
      case ZK_ST_Storage:
        ST_CHECK(ST_STO_FIELD_TAG0, ft0);
        if (!next_same) {
          need_mpt = live;
          proof_type = (val0 && init0) ? ZK_MPT_NonExistingAccountProof : ZK_MPT_StorageMod;
          mpt_unsat_id = ST_STO_MPT_UNSAT;
        } else {
          ST_CHECK(ST_STO_ROOT_SAME, root_same);
        }
        break;
      case ZK_ST_CallContext:
        ST_CHECK(ST_CC_ADDR0, addr0);
        ST_CHECK(ST_CC_KEY0, key0);
        ST_CHECK(ST_CC_FIELD_TAG_RANGE, fr_fits64(ft) && ft.l[0] <= 24);
        if (first_read) ST_CHECK(ST_CC_FIRST_READ0, !val_word && fr_is_zero(val_lo));
        ST_CHECK(ST_CC_INIT0, init0);
        ST_CHECK(ST_CC_ROOT_SAME, root_same);
        break;
      case ZK_ST_Account:
        ST_CHECK(ST_ACC_FIELD_TAG_VALUE, ft.l[0] >= 1 && ft.l[0] <= 4);
        ST_CHECK(ST_ACC_ID0, id0);
        ST_CHECK(ST_ACC_KEY0, key0);
        if (ft.l[0] == ZK_ACC_Nonce) {
          ST_CHECK(ST_ACC_NONCE_VALUE_HI0, fr_is_zero(val_hi));
          ST_CHECK(ST_ACC_NONCE_INIT_HI0, fr_is_zero(init_hi));
        }
        if (!next_same) {
          need_mpt = live;
          proof_type = (val0 && init0 && ft.l[0] == ZK_ACC_CodeHash) ? ZK_MPT_NonExistingAccountProof : ft.l[0];
          mpt_unsat_id = ST_ACC_MPT_UNSAT;
        } else {
          ST_CHECK(ST_ACC_ROOT_SAME, root_same);
        }
        break;
      case ZK_ST_TxRefund:
        ST_CHECK(ST_REF_ADDR0, addr0);
        ST_CHECK(ST_REF_FIELD_TAG0, ft0);
        ST_CHECK(ST_REF_KEY0, key0);
        ST_CHECK(ST_REF_ROOT_SAME, root_same);
        ST_CHECK(ST_REF_INIT0, init0);
        if (first_read) ST_CHECK(ST_REF_FIRST_READ0, val0);
        break;
      case ZK_ST_TxAccessListAccount:
        ST_CHECK(ST_ALA_FIELD_TAG0, ft0);
        ST_CHECK(ST_ALA_KEY0, key0);
        ST_CHECK(ST_ALA_VALUE_HI0, fr_is_zero(val_hi));
        ST_CHECK(ST_ALA_INIT_HI0, fr_is_zero(init_hi));
        ST_CHECK(ST_ALA_ROOT_SAME, root_same);
        if (first_read) ST_CHECK(ST_ALA_FIRST_READ0, !val_word && fr_is_zero(val_lo));
        break;
      case ZK_ST_TxAccessListAccountStorage:
        ST_CHECK(ST_ALS_FIELD_TAG0, ft0);
        ST_CHECK(ST_ALS_VALUE_HI0, fr_is_zero(val_hi));
        ST_CHECK(ST_ALS_INIT_HI0, fr_is_zero(init_hi));
        ST_CHECK(ST_ALS_ROOT_SAME, root_same);
        if (first_read) ST_CHECK(ST_ALS_FIRST_READ0, !val_word && fr_is_zero(val_lo));
        break;
      case ZK_ST_TxLog:
        if (!fr_eq_u64(ft, ZK_LOG_Topic)) {
          ST_CHECK(ST_LOG_VALUE_HI0, fr_is_zero(val_hi));
          ST_CHECK(ST_LOG_INIT_HI0, fr_is_zero(init_hi));
        }
        ST_CHECK(ST_LOG_IS_WRITE, fr_eq_u64(is_write, 1));
        ST_CHECK(ST_LOG_ROOT_SAME, root_same);
        break;
      case ZK_ST_TxReceipt: {
        ST_CHECK(ST_RCP_ADDR0, addr0);
        ST_CHECK(ST_RCP_KEY0, key0);
        ST_CHECK(ST_RCP_VALUE_HI0, fr_is_zero(val_hi));
        ST_CHECK(ST_RCP_INIT_HI0, fr_is_zero(init_hi));
        if (fr_eq_u64(ft, ZK_RCPT_PostStateOrStatus))
          ST_CHECK(ST_RCP_STATUS_BOOL, !val_word && fr_fits64(val_lo) && val_lo.l[0] <= 1);
        const bool tag_same = fr_eq(tag, p_tag);
        if (!fr_eq(id, p_id) && tag_same) {
          ST_CHECK(ST_RCP_TXID_INC, fr_eq(id, fr_add_u64(p_id, 1)));
          if (fr_eq_u64(ft, ZK_RCPT_CumulativeGasUsed)) {
            const bool p_val_word = w.flags && (w.flags[ip] & 1);
            ST_CHECK(ST_RCP_GAS_INC, !val_word && !p_val_word && fr_lt(wcell_l<LAYOUT>(w, T_VAL_LO, ip), val_lo));
          }
        }
        if (!tag_same) ST_CHECK(ST_RCP_FIRST_TXID1, fr_eq_u64(id, 1));
        ST_CHECK(ST_RCP_TXID_RANGE, fr_fits64(id) && id.l[0] >= 1 && id.l[0] <= 2048);
        ST_CHECK(ST_RCP_ROOT_SAME, root_same);
        break;
      }
      default:
        ST_CHECK(ST_TAG_UNREACHABLE, false);
    }
  }
  // MPT lookup for the last access of a Storage / Account key (warp-wide probe)
  {
    const bool go = live && need_mpt;
    Fr key[12] = {addr, fr_u64(proof_type), key_lo, key_hi, wcell_l<LAYOUT>(w, T_ROOT_LO, i), wcell_l<LAYOUT>(w, T_ROOT_HI, i),
                  wcell_l<LAYOUT>(w, T_ROOT_LO, ip), wcell_l<LAYOUT>(w, T_ROOT_HI, ip), val_lo, val_hi, init_lo, init_hi};
    u32 hit;
    const int n = lookup_sync<12>(mpt, key, &hit, mask, go);
    if (go && n != 1 && record) fail(res, n == 0 ? mpt_unsat_id : mpt_unsat_id + 1, row);
  }
}

// whole row on one thread (tests/emu)
template <int LAYOUT>
ZK_HD void check_state_row_dev(const WitnessDev& w, const CheckRange& rg, const IndexDev& mpt, const ResultDev& res,
                               u64 i, bool live, unsigned mask) {
  Fr key_int, p_key_int, limb_sum;
  const bool bytes_ok = gather_key_bytes<LAYOUT>(w, i, &key_int);
  const bool p_bytes_ok = gather_key_bytes<LAYOUT>(w, rot_back(w, i, rg.flags & ZK_FLAG_WRAP), &p_key_int);
  const bool limbs_ok = gather_addr_limbs<LAYOUT>(w, i, &limb_sum);
  check_state_row_core<LAYOUT>(w, rg, mpt, res, i, live, mask, key_int, bytes_ok, p_key_int, p_bytes_ok, limb_sum, limbs_ok);
}

#ifdef __CUDACC__
// The state circuit runs as TWO kernels.  42 of a row's 57 cells are the 32 storage-key bytes and the 10
// address limbs: range-checked, folded into a 256-bit and a 160-bit integer, and otherwise unused.
// k_state_fold streams exactly those columns (few registers, every warp of the SM resident, eight
// 32-byte loads in flight per thread) and leaves 64 bytes per row: the key integer, the limb sum and the
// two range flags.  k_check_state then runs the gate program on the 15 remaining cells plus the folded
// words of its row and of the previous row (the ordering check packs both keys, state_circuit.py:552-570).
struct StateFold {  // one per resident row
  u64 key[4];
  u64 sum[3];
  u64 flags;  // bit 0: every key byte < 256; bit 1: every address limb < 2^16
};
template <int LAYOUT>
__global__ void __launch_bounds__(256) k_state_fold(WitnessDev w, StateFold* out) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x; row < w.n_rows; row += stride) {
    Fr key, sum;
    const bool bytes_ok = gather_key_bytes<LAYOUT>(w, row, &key);
    const bool limbs_ok = gather_addr_limbs<LAYOUT>(w, row, &sum);
    ulonglong4 a, b;
    a.x = key.l[0]; a.y = key.l[1]; a.z = key.l[2]; a.w = key.l[3];
    b.x = sum.l[0]; b.y = sum.l[1]; b.z = sum.l[2]; b.w = (bytes_ok ? 1ull : 0ull) | (limbs_ok ? 2ull : 0ull);
    ulonglong4* p = (ulonglong4*)(out + row);
    p[0] = a;
    p[1] = b;
  }
}
#ifndef ZK_STATE_MINBLOCKS
#define ZK_STATE_MINBLOCKS 3
#endif
template <int LAYOUT>
__global__ void __launch_bounds__(128, ZK_STATE_MINBLOCKS) k_check_state(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ IndexDev mpt, const __grid_constant__ ResultDev res,
              const StateFold* fold) {
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 n = rg.row_end - rg.row_begin;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (u64 first = 0; first < n; first += stride) {  // warp-uniform trip count (the MPT probe is warp-synchronous)
    const u64 k = first + tid;
    const bool live = k < n;
    const u64 i = rg.row_begin + (live ? k : 0);
    const u64* f = (const u64*)(fold + i);
    const u64* pf = (const u64*)(fold + rot_back(w, i, wrap));
    const Fr f0 = ld_cell(f), f1 = ld_cell(f + 4), p0 = ld_cell(pf), p1 = ld_cell(pf + 4);
    const Fr limb_sum{{f1.l[0], f1.l[1], f1.l[2], 0}};
    check_state_row_core<LAYOUT>(w, rg, mpt, res, i, live, 0xFFFFFFFFu, f0, f1.l[3] & 1, p0, p1.l[3] & 1, limb_sum,
                                 (f1.l[3] >> 1) & 1);
  }
}
#endif

}  // namespace zk

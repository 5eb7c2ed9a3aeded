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
#pragma unroll
  for (int q = 0; q < 4; q++) {
    Fr c[8];
#pragma unroll
    for (int b = 0; b < 8; b++) c[b] = wcell_l<LAYOUT>(w, T_BYTE0 + 8 * q + b, row);
    u64 limb = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      ok = ok && fr_fits64(c[b]) && c[b].l[0] < 256;
      limb |= (c[b].l[0] & 0xFF) << (8 * b);
    }
    out->l[q] = limb;
  }
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
// `key_int` / `bytes_ok`: gather_key_bytes of row i; `p_key_int` / `p_bytes_ok`: of the previous row (the
// kernel hands the previous row's over through shared memory instead of gathering 32 cells twice)
template <int LAYOUT>
ZK_HD void check_state_row_core(const WitnessDev& w, const CheckRange& rg, const IndexDev& mpt, const ResultDev& res,
                                u64 i, bool live, unsigned mask, const Fr& key_int, bool bytes_ok, const Fr& p_key_int,
                                bool p_bytes_ok) {
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
  {
    bool limbs_ok = true;
    Fr sum = fr_u64(0);
#pragma unroll
    for (int k = 0; k < 10; k++) {
      const Fr limb = wcell_l<LAYOUT>(w, T_LIMB0 + k, i);
      limbs_ok = limbs_ok && fr_fits64(limb) && limb.l[0] < 65536;
      const int bit = 16 * k;
      sum.l[bit >> 6] |= (limb.l[0] & 0xFFFF) << (bit & 63);
    }
    ST_CHECK(ST_ADDR_LIMB_RANGE, limbs_ok);
    ST_CHECK(ST_ADDR_LIMBS, fr_eq(addr, sum));
  }
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

// whole row on one thread (tests/emu, and rows the tiled kernel cannot serve from shared memory)
template <int LAYOUT>
ZK_HD void check_state_row_dev(const WitnessDev& w, const CheckRange& rg, const IndexDev& mpt, const ResultDev& res,
                               u64 i, bool live, unsigned mask) {
  Fr key_int, p_key_int;
  const bool bytes_ok = gather_key_bytes<LAYOUT>(w, i, &key_int);
  const bool p_bytes_ok = gather_key_bytes<LAYOUT>(w, rot_back(w, i, rg.flags & ZK_FLAG_WRAP), &p_key_int);
  check_state_row_core<LAYOUT>(w, rg, mpt, res, i, live, mask, key_int, bytes_ok, p_key_int, p_bytes_ok);
}

#ifdef __CUDACC__
// One thread per row, 128 consecutive rows per block iteration.  Every thread folds the 32 storage-key
// byte cells of ITS row once and leaves the 256-bit integer in shared memory for its successor (the
// ordering check packs the previous row's key too: state_circuit.py:552-570); the row before the tile is
// folded by warp 0, one byte cell per lane.
#define ZK_STATE_TILE 128
template <int LAYOUT>
__global__ void __launch_bounds__(ZK_STATE_TILE, 3) k_check_state(WitnessDev w, CheckRange rg, IndexDev mpt, ResultDev res) {
  __shared__ u64 s_key[ZK_STATE_TILE + 1][4];
  __shared__ unsigned char s_ok[ZK_STATE_TILE + 1];
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 n = rg.row_end - rg.row_begin;
  const u64 n_tiles = (n + ZK_STATE_TILE - 1) / ZK_STATE_TILE;
  const unsigned tid = threadIdx.x, lane = tid & 31;
  for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {  // block-uniform trip count
    const u64 first = rg.row_begin + tile * ZK_STATE_TILE;
    const u64 k = tile * ZK_STATE_TILE + tid;
    const bool live = k < n;
    const u64 i = live ? rg.row_begin + k : first;
    Fr key_int;
    const bool bytes_ok = gather_key_bytes<LAYOUT>(w, i, &key_int);
#pragma unroll
    for (int q = 0; q < 4; q++) s_key[tid + 1][q] = key_int.l[q];
    s_ok[tid + 1] = bytes_ok;
    if (tid < 32) {  // the row before the tile: lane b folds byte cell b
      const Fr c = wcell_l<LAYOUT>(w, T_BYTE0 + lane, rot_back(w, first, wrap));
      const bool okb = fr_fits64(c) && c.l[0] < 256;
      u64 v = (c.l[0] & 0xFF) << (8 * (lane & 7));
      v |= __shfl_xor_sync(0xFFFFFFFFu, v, 1);
      v |= __shfl_xor_sync(0xFFFFFFFFu, v, 2);
      v |= __shfl_xor_sync(0xFFFFFFFFu, v, 4);
      const unsigned all_ok = __all_sync(0xFFFFFFFFu, okb);
      if ((lane & 7) == 0) s_key[0][lane >> 3] = v;
      if (lane == 0) s_ok[0] = all_ok;
    }
    __syncthreads();
    const Fr p_key_int{{s_key[tid][0], s_key[tid][1], s_key[tid][2], s_key[tid][3]}};
    const bool p_bytes_ok = s_ok[tid];
    check_state_row_core<LAYOUT>(w, rg, mpt, res, i, live, 0xFFFFFFFFu, key_int, bytes_ok, p_key_int, p_bytes_ok);
    __syncthreads();
  }
}
#endif

}  // namespace zk

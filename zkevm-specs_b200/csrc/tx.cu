// tx.cu — the Fr parts of the tx circuit on the device.
//
// Replaces SignVerifyChip.verify (src/zkevm_specs/tx_circuit.py:205-243) and the copy constraints
// of verify_circuit (:253-289), one row per tx_index.  The ECDSA check itself (:147-158) is
// third-party curve math (eth_keys), its verdict enters as row flag bit 1.
// Row = 14 cells: address | pub_key_x (lo, hi) | pub_key_y (lo, hi) | Word(pub_key_hash) (lo, hi) |
// msg_hash (lo, hi) | Word(msg_hash_bytes) (lo, hi) | tx-table CallerAddress value | tx-table
// TxSignHash (lo, hi).  A 32-byte `bytes` field is held as the Word of its bytes (lo = bytes 0..15
// little-endian, hi = bytes 16..31): pub_key_x_bytes is little-endian (:128-129), so its cells are
// the coordinate as an integer, and byte i of a field is byte (i mod 16) of a cell.
// Keccak table (tx_circuit.py:36-61): (is_enabled, input_rlc, input_len, output lo, hi); lookup =
// set membership of the whole tuple.  64 Fr x Fr products per row (the public-key RLC, Horner with
// the challenge in Montgomery form).  A row stops at its first failing constraint.
#pragma once
#include "circuit.cuh"

namespace zk {

enum { X_ADDR, X_PKX_LO, X_PKX_HI, X_PKY_LO, X_PKY_HI, X_PKH_LO, X_PKH_HI, X_MSG_LO, X_MSG_HI, X_MSGB_LO, X_MSGB_HI,
       X_ROW_ADDR, X_ROW_HASH_LO, X_ROW_HASH_HI };

ZK_HD u64 cell_byte(const Fr& c, int k) { return (c.l[k >> 3] >> (8 * (k & 7))) & 0xFF; }  // k < 16

#define TX_CHECK(id, cond)      \
  do {                          \
    if (!(cond)) {              \
      fail(res, (id), row);     \
      return;                   \
    }                           \
  } while (0)

ZK_HD void check_tx_row(const WitnessDev& w, const CheckRange& rg, const IndexDev& keccak, const Fr& r_mont,
                        const ResultDev& res, u64 i) {
  const u64 row = rg.row_base + i;  // reported row (zkcheck.h: first_fail holds row_base + local row)
  const Fr address = wcell(w, X_ADDR, i);
  const bool np = !fr_is_zero(address);  // is_not_padding, :206
  const Fr pkx_lo = wcell(w, X_PKX_LO, i), pkx_hi = wcell(w, X_PKX_HI, i);
  const Fr pky_lo = wcell(w, X_PKY_LO, i), pky_hi = wcell(w, X_PKY_HI, i);
  const Fr pkh_lo = wcell(w, X_PKH_LO, i), pkh_hi = wcell(w, X_PKH_HI, i);
  const Fr msgb_lo = wcell(w, X_MSGB_LO, i), msgb_hi = wcell(w, X_MSGB_HI, i);
  const Fr msg_lo = wcell(w, X_MSG_LO, i), msg_hi = wcell(w, X_MSG_HI, i);
  TX_CHECK(TX_BYTE_DOMAIN, fr_fits128(pkx_lo) && fr_fits128(pkx_hi) && fr_fits128(pky_lo) && fr_fits128(pky_hi) &&
                               fr_fits128(pkh_lo) && fr_fits128(pkh_hi) && fr_fits128(msgb_lo) && fr_fits128(msgb_hi));
  // :215-226  RLC(reversed(x_be + y_be), r, 64): little-endian sequence = y_le then x_le, Horner
  // from the last element (linear_combine_bytes, util/arithmetic.py:9-24)
  Fr acc = fr_u64(0);
#pragma unroll 1
  for (int k = 63; k >= 0; k--) {
    const Fr& cell = k >= 48 ? pkx_hi : (k >= 32 ? pkx_lo : (k >= 16 ? pky_hi : pky_lo));
    acc = fr_add_u64(fr_montmul(acc, r_mont), cell_byte(cell, k & 15));
  }
  {
    const Fr z = fr_u64(0);
    Fr key[5] = {fr_u64(np ? 1 : 0), np ? acc : z, fr_u64(np ? 64 : 0), np ? pkh_lo : z, np ? pkh_hi : z};
    u32 hit;
    TX_CHECK(TX_KECCAK_LOOKUP, lookup<5>(keccak, key, &hit) >= 1);
  }
  {
    // :229-232  address == int.from_bytes(pub_key_hash[-20:], "big"): hash bytes 16..31 are the hi
    // cell, bytes 12..15 the top four bytes of the lo cell
    Fr a = fr_u64(0);
#pragma unroll
    for (int k = 0; k < 16; k++) a.l[k >> 3] |= cell_byte(pkh_hi, 15 - k) << (8 * (k & 7));
#pragma unroll
    for (int k = 0; k < 4; k++) a.l[2] |= cell_byte(pkh_lo, 15 - k) << (8 * k);
    TX_CHECK(TX_ADDRESS, fr_eq(a, address));
  }
  TX_CHECK(TX_MSG_HASH, np ? (fr_eq(msgb_lo, msg_lo) && fr_eq(msgb_hi, msg_hi)) : (fr_is_zero(msg_lo) && fr_is_zero(msg_hi)));
  const unsigned f = w.flags ? w.flags[i] : 0;
  TX_CHECK(TX_ECDSA, !(f & 2));
  TX_CHECK(TX_ROW_ADDR_TYPE, !(f & 1));
  TX_CHECK(TX_ROW_ADDR, fr_eq(wcell(w, X_ROW_ADDR, i), address));
  TX_CHECK(TX_ROW_HASH_LO, fr_eq(wcell(w, X_ROW_HASH_LO, i), msg_lo));
  TX_CHECK(TX_ROW_HASH_HI, fr_eq(wcell(w, X_ROW_HASH_HI, i), msg_hi));
}

// ---- sig circuit: sig_circuit.Row.verify (src/zkevm_specs/sig_circuit.py:64-104) ---------------
// Row = 21 cells: sig_v | recovered_addr | pub_key_x (lo, hi) | pub_key_y (lo, hi) | Word(pub_key_hash)
// (lo, hi) | msg_hash (lo, hi) | Word(msg_hash_bytes) (lo, hi) | is_valid | sig_r (lo, hi) | sig_s
// (lo, hi) | the ECDSA chip's r (lo, hi) and s (lo, hi).  Row flag bit 1 = ecdsa_chip.verify()
// (util/ec.py:109-117, eth_keys: third party) returned True.
enum { Y_V, Y_ADDR, Y_PKX_LO, Y_PKX_HI, Y_PKY_LO, Y_PKY_HI, Y_PKH_LO, Y_PKH_HI, Y_MSG_LO, Y_MSG_HI, Y_MSGB_LO,
       Y_MSGB_HI, Y_VALID, Y_R_LO, Y_R_HI, Y_S_LO, Y_S_HI, Y_CR_LO, Y_CR_HI, Y_CS_LO, Y_CS_HI };

// Word.int_value() == chip integer (sig_circuit.py:70-71): lo + hi * 2^128 as INTEGERS (the halves
// of a corrupted Word may exceed 2^128) against a 256-bit value held as two 128-bit cells
ZK_HD bool word_int_eq(const Fr& lo, const Fr& hi, const Fr& c_lo, const Fr& c_hi) {
  u64 c = 0;
  const u64 s2 = adc64(lo.l[2], hi.l[0], c), s3 = adc64(lo.l[3], hi.l[1], c), s4 = adc64(hi.l[2], 0, c),
            s5 = adc64(hi.l[3], 0, c);
  return lo.l[0] == c_lo.l[0] && lo.l[1] == c_lo.l[1] && s2 == c_hi.l[0] && s3 == c_hi.l[1] && s4 == 0 && s5 == 0 && c == 0;
}
ZK_HD void check_sig_row(const WitnessDev& w, const CheckRange& rg, const IndexDev& keccak, const Fr& r_mont,
                         const ResultDev& res, u64 i) {
  const u64 row = rg.row_base + i;
  const Fr pkx_lo = wcell(w, Y_PKX_LO, i), pkx_hi = wcell(w, Y_PKX_HI, i);
  const Fr pky_lo = wcell(w, Y_PKY_LO, i), pky_hi = wcell(w, Y_PKY_HI, i);
  const Fr pkh_lo = wcell(w, Y_PKH_LO, i), pkh_hi = wcell(w, Y_PKH_HI, i);
  const Fr msgb_lo = wcell(w, Y_MSGB_LO, i), msgb_hi = wcell(w, Y_MSGB_HI, i);
  TX_CHECK(SG_BYTE_DOMAIN, fr_fits128(pkx_lo) && fr_fits128(pkx_hi) && fr_fits128(pky_lo) && fr_fits128(pky_hi) &&
                               fr_fits128(pkh_lo) && fr_fits128(pkh_hi) && fr_fits128(msgb_lo) && fr_fits128(msgb_hi) &&
                               fr_fits128(wcell(w, Y_CR_LO, i)) && fr_fits128(wcell(w, Y_CR_HI, i)) &&
                               fr_fits128(wcell(w, Y_CS_LO, i)) && fr_fits128(wcell(w, Y_CS_HI, i)));
  TX_CHECK(SG_SIG_R_COPY, word_int_eq(wcell(w, Y_R_LO, i), wcell(w, Y_R_HI, i), wcell(w, Y_CR_LO, i), wcell(w, Y_CR_HI, i)));
  TX_CHECK(SG_SIG_S_COPY, word_int_eq(wcell(w, Y_S_LO, i), wcell(w, Y_S_HI, i), wcell(w, Y_CS_LO, i), wcell(w, Y_CS_HI, i)));
  const Fr v = wcell(w, Y_V, i);
  TX_CHECK(SG_V_BOOL, fr_is_zero(v) || fr_eq_u64(v, 1));
  Fr acc = fr_u64(0);
#pragma unroll 1
  for (int k = 63; k >= 0; k--) {
    const Fr& cell = k >= 48 ? pkx_hi : (k >= 32 ? pkx_lo : (k >= 16 ? pky_hi : pky_lo));
    acc = fr_add_u64(fr_montmul(acc, r_mont), cell_byte(cell, k & 15));
  }
  {
    Fr key[5] = {fr_u64(1), acc, fr_u64(64), pkh_lo, pkh_hi};
    u32 hit;
    TX_CHECK(SG_KECCAK_LOOKUP, lookup<5>(keccak, key, &hit) >= 1);
  }
  {
    Fr a = fr_u64(0);
#pragma unroll
    for (int k = 0; k < 16; k++) a.l[k >> 3] |= cell_byte(pkh_hi, 15 - k) << (8 * (k & 7));
#pragma unroll
    for (int k = 0; k < 4; k++) a.l[2] |= cell_byte(pkh_lo, 15 - k) << (8 * k);
    TX_CHECK(SG_ADDRESS, fr_eq(a, wcell(w, Y_ADDR, i)));
  }
  TX_CHECK(SG_MSG_HASH, fr_eq(msgb_lo, wcell(w, Y_MSG_LO, i)) && fr_eq(msgb_hi, wcell(w, Y_MSG_HI, i)));
  const unsigned f = w.flags ? w.flags[i] : 0;
  TX_CHECK(SG_ECDSA_VALID, fr_eq_u64(wcell(w, Y_VALID, i), (f >> 1) & 1));
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(128) k_check_sig(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ IndexDev keccak, const __grid_constant__ Fr r_mont, const __grid_constant__ ResultDev res) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = rg.row_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rg.row_end; i += stride)
    check_sig_row(w, rg, keccak, r_mont, res, i);
}
__global__ void __launch_bounds__(128) k_check_tx(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ IndexDev keccak, const __grid_constant__ Fr r_mont, const __grid_constant__ ResultDev res) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = rg.row_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rg.row_end; i += stride)
    check_tx_row(w, rg, keccak, r_mont, res, i);
}
#endif

}  // namespace zk

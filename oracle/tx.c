/*
 * oracle/tx.c — TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
 *
 * Restates the Fr parts of the tx circuit, /root/reference/src/zkevm_specs/tx_circuit.py:
 * SignVerifyChip.verify (:205-243) and the copy constraints of verify_circuit (:253-289).  The
 * ECDSA check (:147-158) is third-party curve math (eth_keys): its verdict is an input flag.
 * Row = 14 cells: address | pub_key_x (lo, hi) | pub_key_y (lo, hi) | Word(pub_key_hash) (lo, hi) |
 * msg_hash (lo, hi) | Word(msg_hash_bytes) (lo, hi) | tx-table CallerAddress value | tx-table
 * TxSignHash (lo, hi).  A 32-byte `bytes` field is held as the Word of its bytes (lo = bytes 0..15
 * little-endian, hi = bytes 16..31), so pub_key_x/y cells are the coordinates as integers
 * (pub_key_x_bytes is little-endian, :128-129) and byte i of a field is byte (i mod 16) of a cell.
 * Row flags: bit 0 = the CallerAddress cell is a Word, bit 1 = ecdsa_verify failed.
 * Keccak table (tx_circuit.py:36-61): rows (is_enabled, input_rlc, input_len, output lo, hi); the
 * lookup is set membership of the whole tuple.  A row stops at its first failing constraint.
 * Pinned by tests/golden/tx.npz.
 */
#include "common.h"
#include "lookup.h"

enum { T_ADDR, T_PKX_LO, T_PKX_HI, T_PKY_LO, T_PKY_HI, T_PKH_LO, T_PKH_HI, T_MSG_LO, T_MSG_HI, T_MSGB_LO, T_MSGB_HI,
       T_ROW_ADDR, T_ROW_HASH_LO, T_ROW_HASH_HI, TX_COLS };

static unsigned cell_byte(fr_t c, int k) { return (unsigned)((c.l[k >> 3] >> (8 * (k & 7))) & 0xFF); } /* k < 16 */

int orc_check_tx(const uint64_t* rows, uint64_t n_rows, const uint8_t* row_flags, const uint64_t* keccak,
                 uint64_t n_keccak, const uint64_t r_limbs[4], uint64_t row_begin, uint64_t row_end,
                 uint32_t* first_fail, uint64_t* fail_count) {
  orc_result res_, *res = &res_; orc_result_init(res, first_fail, fail_count, TX_N_CONSTRAINTS);
  const uint32_t kk[5] = {0, 1, 2, 3, 4};
  orc_index kix; orc_index_build(&kix, keccak, n_keccak, 5, kk, 5);
  const fr_t r = fr_load(r_limbs);
#define TK(id, cond) do { if (!(cond)) { orc_fail(res, (id), i); goto next_row; } } while (0)
#define C(c) fr_load(ORC_CELL(rows, n_rows, c, i))
  for (uint64_t i = row_begin; i < row_end; i++) {
    const fr_t address = C(T_ADDR);
    const int np = !fr_is_zero(address); /* is_not_padding = 1 - (address == 0), :206 */
    const fr_t halves[8] = {C(T_PKX_LO), C(T_PKX_HI), C(T_PKY_LO), C(T_PKY_HI), C(T_PKH_LO), C(T_PKH_HI),
                            C(T_MSGB_LO), C(T_MSGB_HI)};
    int dom = 1;
    for (int k = 0; k < 8; k++) dom = dom && fr_fits_bits(halves[k], 128);
    TK(TX_BYTE_DOMAIN, dom);
    /* :215-226  RLC(reversed(x_be + y_be), r, 64): little-endian sequence = y_le then x_le,
     * Horner from the last element (linear_combine_bytes, util/arithmetic.py:9-24) */
    fr_t acc = fr_u64(0);
    for (int k = 63; k >= 0; k--) {
      const fr_t cell = k >= 48 ? halves[1] : k >= 32 ? halves[0] : k >= 16 ? halves[3] : halves[2];
      acc = fr_add(fr_mul(acc, r), fr_u64(cell_byte(cell, k & 15)));
    }
    {
      const fr_t z = fr_u64(0);
      fr_t key[5] = {fr_u64(np), np ? acc : z, fr_u64(np ? 64 : 0), np ? halves[4] : z, np ? halves[5] : z};
      TK(TX_KECCAK_LOOKUP, orc_lookup(&kix, key, 0) >= 1);
    }
    /* :229-232  address == int.from_bytes(pub_key_hash[-20:], "big"): hash bytes 16..31 are the hi
     * cell, bytes 12..15 the top four bytes of the lo cell */
    {
      fr_t a = fr_u64(0);
      for (int k = 0; k < 16; k++) a.l[k >> 3] |= (uint64_t)cell_byte(halves[5], 15 - k) << (8 * (k & 7));
      for (int k = 0; k < 4; k++) a.l[2] |= (uint64_t)cell_byte(halves[4], 15 - k) << (8 * k);
      TK(TX_ADDRESS, fr_eq(a, address));
    }
    /* :236-239 */
    {
      const fr_t z = fr_u64(0);
      TK(TX_MSG_HASH, fr_eq(np ? halves[6] : z, C(T_MSG_LO)) && fr_eq(np ? halves[7] : z, C(T_MSG_HI)));
    }
    const unsigned f = row_flags ? row_flags[i] : 0;
    TK(TX_ECDSA, !(f & 2));
    TK(TX_ROW_ADDR_TYPE, !(f & 1));
    TK(TX_ROW_ADDR, fr_eq(C(T_ROW_ADDR), address));
    TK(TX_ROW_HASH_LO, fr_eq(C(T_ROW_HASH_LO), C(T_MSG_LO)));
    TK(TX_ROW_HASH_HI, fr_eq(C(T_ROW_HASH_HI), C(T_MSG_HI)));
  next_row:;
  }
  orc_index_free(&kix);
  return 0;
}

/* sig circuit, src/zkevm_specs/sig_circuit.py:64-104: Row.verify.  Row = 21 cells: sig_v |
 * recovered_addr | pub_key_x (lo, hi) | pub_key_y (lo, hi) | Word(pub_key_hash) (lo, hi) | msg_hash
 * (lo, hi) | Word(msg_hash_bytes) (lo, hi) | is_valid | sig_r (lo, hi) | sig_s (lo, hi) | chip r
 * (lo, hi) | chip s (lo, hi).  Row flag bit 1 = ecdsa_chip.verify() returned True.
 * Pinned by tests/golden/sig.npz. */
enum { G_V, G_ADDR, G_PKX_LO, G_PKX_HI, G_PKY_LO, G_PKY_HI, G_PKH_LO, G_PKH_HI, G_MSG_LO, G_MSG_HI, G_MSGB_LO,
       G_MSGB_HI, G_VALID, G_R_LO, G_R_HI, G_S_LO, G_S_HI, G_CR_LO, G_CR_HI, G_CS_LO, G_CS_HI, SIG_COLS };

/* Word.int_value() == chip integer (sig_circuit.py:70-71): lo + hi * 2^128 as INTEGERS (the halves of
 * a corrupted Word may exceed 2^128) against a 256-bit value held as two 128-bit cells */
static int word_int_eq(fr_t lo, fr_t hi, fr_t c_lo, fr_t c_hi) {
  uint64_t s[6], c = 0;
  s[0] = lo.l[0]; s[1] = lo.l[1];
  s[2] = adc(lo.l[2], hi.l[0], &c);
  s[3] = adc(lo.l[3], hi.l[1], &c);
  s[4] = adc(hi.l[2], 0, &c);
  s[5] = adc(hi.l[3], 0, &c);
  return s[0] == c_lo.l[0] && s[1] == c_lo.l[1] && s[2] == c_hi.l[0] && s[3] == c_hi.l[1] && s[4] == 0 && s[5] == 0 && c == 0;
}

int orc_check_sig(const uint64_t* rows, uint64_t n_rows, const uint8_t* row_flags, const uint64_t* keccak,
                  uint64_t n_keccak, const uint64_t r_limbs[4], uint64_t row_begin, uint64_t row_end,
                  uint32_t* first_fail, uint64_t* fail_count) {
  orc_result res_, *res = &res_; orc_result_init(res, first_fail, fail_count, SG_N_CONSTRAINTS);
  const uint32_t kk[5] = {0, 1, 2, 3, 4};
  orc_index kix; orc_index_build(&kix, keccak, n_keccak, 5, kk, 5);
  const fr_t r = fr_load(r_limbs);
  for (uint64_t i = row_begin; i < row_end; i++) {
    const fr_t halves[8] = {C(G_PKX_LO), C(G_PKX_HI), C(G_PKY_LO), C(G_PKY_HI), C(G_PKH_LO), C(G_PKH_HI),
                            C(G_MSGB_LO), C(G_MSGB_HI)};
    int dom = 1;
    for (int k = 0; k < 8; k++) dom = dom && fr_fits_bits(halves[k], 128);
    dom = dom && fr_fits_bits(C(G_CR_LO), 128) && fr_fits_bits(C(G_CR_HI), 128) && fr_fits_bits(C(G_CS_LO), 128) &&
          fr_fits_bits(C(G_CS_HI), 128); /* the chip's r, s are 32 `bytes` */
    TK(SG_BYTE_DOMAIN, dom);
    TK(SG_SIG_R_COPY, word_int_eq(C(G_R_LO), C(G_R_HI), C(G_CR_LO), C(G_CR_HI)));
    TK(SG_SIG_S_COPY, word_int_eq(C(G_S_LO), C(G_S_HI), C(G_CS_LO), C(G_CS_HI)));
    const fr_t v = C(G_V);
    TK(SG_V_BOOL, fr_eq_u64(v, 0) || fr_eq_u64(v, 1));
    fr_t acc = fr_u64(0);
    for (int k = 63; k >= 0; k--) {
      const fr_t cell = k >= 48 ? halves[1] : k >= 32 ? halves[0] : k >= 16 ? halves[3] : halves[2];
      acc = fr_add(fr_mul(acc, r), fr_u64(cell_byte(cell, k & 15)));
    }
    {
      fr_t key[5] = {fr_u64(1), acc, fr_u64(64), halves[4], halves[5]};
      TK(SG_KECCAK_LOOKUP, orc_lookup(&kix, key, 0) >= 1);
    }
    {
      fr_t a = fr_u64(0);
      for (int k = 0; k < 16; k++) a.l[k >> 3] |= (uint64_t)cell_byte(halves[5], 15 - k) << (8 * (k & 7));
      for (int k = 0; k < 4; k++) a.l[2] |= (uint64_t)cell_byte(halves[4], 15 - k) << (8 * k);
      TK(SG_ADDRESS, fr_eq(a, C(G_ADDR)));
    }
    TK(SG_MSG_HASH, fr_eq(halves[6], C(G_MSG_LO)) && fr_eq(halves[7], C(G_MSG_HI)));
    const unsigned f = row_flags ? row_flags[i] : 0;
    TK(SG_ECDSA_VALID, fr_eq_u64(C(G_VALID), (f >> 1) & 1));
  next_row:;
  }
  orc_index_free(&kix);
  return 0;
}

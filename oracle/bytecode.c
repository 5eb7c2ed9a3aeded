/*
 * oracle/bytecode.c — TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
 *
 * Restates check_bytecode_row and its four helpers,
 * /root/reference/src/zkevm_specs/bytecode_circuit.py:37-100, over the column-major
 * matrix of include/zkcheck.h (12 cells/row, order of Row, bytecode_circuit.py:15-27).
 * Driver loop with wrap-around next row: tests/test_bytecode_circuit.py:26-47.
 * Pinned by tests/golden/bytecode_*.npz (generated from the reference itself).
 */
#include "common.h"

enum { C_QFIRST, C_QLAST, C_HASH_LO, C_HASH_HI, C_TAG, C_INDEX, C_VALUE, C_ISCODE, C_PDL,
       C_RLC, C_LEN, C_PDS, BC_COLS };
#define TAG_HEADER 1 /* BytecodeFieldTag.Header, evm_circuit/table.py:170-176 */
#define TAG_BYTE 2

/* Word(EMPTY_HASH): keccak256("") split lo/hi, util/hash.py:13 */
static const fr_t EMPTY_LO = {{0x7bfad8045d85a470ull, 0xe500b653ca82273bull, 0, 0}};
static const fr_t EMPTY_HI = {{0x927e7db2dcc703c0ull, 0xc5d2460186f7233cull, 0, 0}};

int orc_check_bytecode(const uint64_t* cols, uint64_t n_rows, const uint64_t* push_tab,
                       uint64_t n_push, const uint64_t* keccak_tab, uint64_t n_keccak,
                       const uint64_t r_[4], uint64_t row_begin, uint64_t row_end,
                       uint32_t* first_fail, uint64_t* fail_count) {
  orc_result res; orc_result_init(&res, first_fail, fail_count, BC_N_CONSTRAINTS);
  const fr_t r = fr_load(r_);
  orc_index push_ix, kec_ix;
  const uint32_t pk[2] = {0, 1}, kk[5] = {0, 1, 2, 3, 4};
  orc_index_build(&push_ix, push_tab, n_push, 2, pk, 2);
  orc_index_build(&kec_ix, keccak_tab, n_keccak, 5, kk, 5);
#define CUR(c) fr_load(ORC_CELL(cols, n_rows, c, i))
#define NXT(c) fr_load(ORC_CELL(cols, n_rows, c, j))
  for (uint64_t i = row_begin; i < row_end; i++) {
    const uint64_t j = (i + 1) % n_rows;
    const fr_t tag = CUR(C_TAG), ntag = NXT(C_TAG);
    const int q_first = fr_eq_u64(CUR(C_QFIRST), 1);
    const int q_last0 = fr_eq_u64(CUR(C_QLAST), 0), q_last1 = fr_eq_u64(CUR(C_QLAST), 1);
    const int hash_same = fr_eq(NXT(C_HASH_LO), CUR(C_HASH_LO)) && fr_eq(NXT(C_HASH_HI), CUR(C_HASH_HI));
    const int hash_empty = fr_eq(CUR(C_HASH_LO), EMPTY_LO) && fr_eq(CUR(C_HASH_HI), EMPTY_HI);
    if (q_first) REQUIRE(&res, BC_FIRST_TAG, i, fr_eq_u64(tag, TAG_HEADER));
    if (q_last0) {
      if (fr_eq_u64(tag, TAG_HEADER)) {
        REQUIRE(&res, BC_HDR_VALUE_LEN, i, fr_eq(CUR(C_VALUE), CUR(C_LEN)));
        REQUIRE(&res, BC_HDR_INDEX0, i, fr_eq_u64(CUR(C_INDEX), 0));
        if (fr_eq_u64(ntag, TAG_BYTE)) { /* check_bytecode_row_header_to_byte :72-77 */
          REQUIRE(&res, BC_H2B_LEN, i, fr_eq(NXT(C_LEN), CUR(C_LEN)));
          REQUIRE(&res, BC_H2B_INDEX0, i, fr_eq_u64(NXT(C_INDEX), 0));
          REQUIRE(&res, BC_H2B_ISCODE, i, fr_eq_u64(NXT(C_ISCODE), 1));
          REQUIRE(&res, BC_H2B_HASH, i, hash_same);
          REQUIRE(&res, BC_H2B_RLC, i, fr_eq(NXT(C_RLC), NXT(C_VALUE)));
        }
        if (fr_eq_u64(ntag, TAG_HEADER)) { /* header_to_header :81-82 */
          REQUIRE(&res, BC_H2H_LEN0, i, fr_eq_u64(CUR(C_LEN), 0));
          REQUIRE(&res, BC_H2H_EMPTY_HASH, i, hash_empty);
        }
      }
      if (fr_eq_u64(tag, TAG_BYTE)) {
        fr_t key[2] = {CUR(C_VALUE), CUR(C_PDS)};
        REQUIRE(&res, BC_PUSH_TABLE, i, orc_lookup(&push_ix, key, 0) >= 1);
        REQUIRE(&res, BC_IS_CODE, i, fr_eq_u64(CUR(C_ISCODE), fr_is_zero(CUR(C_PDL)) ? 1 : 0));
        if (fr_eq_u64(ntag, TAG_BYTE)) { /* byte_to_byte :86-94 */
          REQUIRE(&res, BC_B2B_LEN, i, fr_eq(NXT(C_LEN), CUR(C_LEN)));
          REQUIRE(&res, BC_B2B_INDEX, i, fr_eq(NXT(C_INDEX), fr_add(CUR(C_INDEX), fr_u64(1))));
          REQUIRE(&res, BC_B2B_HASH, i, hash_same);
          REQUIRE(&res, BC_B2B_RLC, i,
                  fr_eq(NXT(C_RLC), fr_add(fr_mul(CUR(C_RLC), r), NXT(C_VALUE))));
          if (fr_eq_u64(CUR(C_ISCODE), 1))
            REQUIRE(&res, BC_B2B_PUSH_LEFT, i, fr_eq(NXT(C_PDL), CUR(C_PDS)));
          else
            REQUIRE(&res, BC_B2B_PUSH_LEFT, i, fr_eq(NXT(C_PDL), fr_sub(CUR(C_PDL), fr_u64(1))));
        }
        if (fr_eq_u64(ntag, TAG_HEADER)) { /* byte_to_header :98-100 */
          REQUIRE(&res, BC_B2H_INDEX, i, fr_eq(fr_add(CUR(C_INDEX), fr_u64(1)), CUR(C_LEN)));
          fr_t kkey[5] = {fr_u64(2), CUR(C_RLC), CUR(C_LEN), CUR(C_HASH_LO), CUR(C_HASH_HI)};
          REQUIRE(&res, BC_B2H_KECCAK, i, orc_lookup(&kec_ix, kkey, 0) >= 1);
        }
      }
    }
    if (q_last1) {
      REQUIRE(&res, BC_LAST_TAG, i, fr_eq_u64(tag, TAG_HEADER));
      REQUIRE(&res, BC_LAST_LEN0, i, fr_eq_u64(CUR(C_LEN), 0));
      REQUIRE(&res, BC_LAST_EMPTY_HASH, i, hash_empty);
    }
  }
#undef CUR
#undef NXT
  orc_index_free(&push_ix); orc_index_free(&kec_ix);
  return 0;
}

/* Fr known-answer hooks for tests/test_oracle_fr.py */
void orc_fr_mul(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
  fr_t r = fr_mul(fr_load(a), fr_load(b)); memcpy(out, r.l, 32);
}
void orc_fr_add(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
  fr_t r = fr_add(fr_load(a), fr_load(b)); memcpy(out, r.l, 32);
}
void orc_fr_sub(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
  fr_t r = fr_sub(fr_load(a), fr_load(b)); memcpy(out, r.l, 32);
}
void orc_fr_inv(const uint64_t a[4], uint64_t out[4]) {
  fr_t r = fr_inv(fr_load(a)); memcpy(out, r.l, 32);
}

// bytecode.cu — bytecode-circuit row checker (one thread per row).
//
// Replaces the per-row loop around check_bytecode_row and its four helpers,
// src/zkevm_specs/bytecode_circuit.py:37-100 (driver: tests/test_bytecode_circuit.py:26-47).
// Row = 12 cells in the order of bytecode_circuit.Row (:15-27); rotations {0,+1}.
// Algorithmic bytes: 12 cells x 32 B = 384 B per row (each cell counted once; the +1 rotation
// is served by L1/L2 because the neighbouring thread streams the same cells).
// Fr x Fr products: exactly one per Byte->Byte row (value_rlc * r).
#include "circuit.cuh"
#include "../../include/zk_constraints.h"
#include "../../include/zkcheck.h"

namespace zk {

enum { C_QFIRST, C_QLAST, C_HASH_LO, C_HASH_HI, C_TAG, C_INDEX, C_VALUE, C_ISCODE, C_PDL, C_RLC,
       C_LEN, C_PDS };
#define TAG_HEADER 1  // BytecodeFieldTag.Header (evm_circuit/table.py:170-176)
#define TAG_BYTE 2

// Word(EMPTY_HASH) = keccak256("") as lo/hi 128-bit halves (util/hash.py:13)
#define EMPTY_LO0 0x7bfad8045d85a470ull
#define EMPTY_LO1 0xe500b653ca82273bull
#define EMPTY_HI0 0x927e7db2dcc703c0ull
#define EMPTY_HI1 0xc5d2460186f7233cull

// (value, push_data_size) in push_table (bytecode_circuit.py:174-178).  The table has one row per byte
// value; when the verify pass found it regular (row v holds value v: ZK_POS_DENSE with base 0, 256 rows)
// its push_size column sits in shared memory (`s_push`, staged with one bulk copy) and membership is
// one indexed compare; otherwise the exact hash probe.
ZK_HD bool push_member(const IndexDev& push_ix, const u64* s_push, const Fr& value, const Fr& pds) {
  if (s_push) {
    if (!(fr_fits64(value) && value.l[0] < 256)) return false;
    const u64* c = s_push + 4 * value.l[0];
    return c[0] == pds.l[0] && c[1] == pds.l[1] && c[2] == pds.l[2] && c[3] == pds.l[3];
  }
  Fr pkey[2] = {value, pds};
  u32 hit;
  return lookup<2>(push_ix, pkey, &hit) >= 1;
}

template <int LAYOUT, class W>
ZK_HD void check_bytecode_row(const W& w, const CheckRange& rg, const IndexDev& push_ix, const u64* s_push,
                              const IndexDev& kec_ix, const Fr& r_mont, const ResultDev& res, u64 i) {
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 j = rot_fwd(w, i, 1, wrap);
  const u64 row = rg.row_base + i;

  const Fr q_first = wcell_l<LAYOUT>(w, C_QFIRST, i), q_last = wcell_l<LAYOUT>(w, C_QLAST, i);
  const Fr tag = wcell_l<LAYOUT>(w, C_TAG, i), ntag = wcell_l<LAYOUT>(w, C_TAG, j);
  const Fr hash_lo = wcell_l<LAYOUT>(w, C_HASH_LO, i), hash_hi = wcell_l<LAYOUT>(w, C_HASH_HI, i);
  const Fr index = wcell_l<LAYOUT>(w, C_INDEX, i), value = wcell_l<LAYOUT>(w, C_VALUE, i);
  const Fr is_code = wcell_l<LAYOUT>(w, C_ISCODE, i), pdl = wcell_l<LAYOUT>(w, C_PDL, i);
  const Fr rlc = wcell_l<LAYOUT>(w, C_RLC, i), len = wcell_l<LAYOUT>(w, C_LEN, i), pds = wcell_l<LAYOUT>(w, C_PDS, i);

  const bool is_hdr = fr_eq_u64(tag, TAG_HEADER), is_byte = fr_eq_u64(tag, TAG_BYTE);
  const bool n_hdr = fr_eq_u64(ntag, TAG_HEADER), n_byte = fr_eq_u64(ntag, TAG_BYTE);
  const bool hash_empty = fr_eq(hash_lo, fr_u128(EMPTY_LO0, EMPTY_LO1)) &&
                          fr_eq(hash_hi, fr_u128(EMPTY_HI0, EMPTY_HI1));

  if (fr_eq_u64(q_first, 1)) ZK_REQUIRE(res, BC_FIRST_TAG, row, is_hdr);

  if (fr_eq_u64(q_last, 0)) {
    if (is_hdr) {
      ZK_REQUIRE(res, BC_HDR_VALUE_LEN, row, fr_eq(value, len));
      ZK_REQUIRE(res, BC_HDR_INDEX0, row, fr_is_zero(index));
      if (n_hdr) {
        ZK_REQUIRE(res, BC_H2H_LEN0, row, fr_is_zero(len));
        ZK_REQUIRE(res, BC_H2H_EMPTY_HASH, row, hash_empty);
      }
    }
    if (is_byte) {
      u32 hit;
      ZK_REQUIRE(res, BC_PUSH_TABLE, row, push_member(push_ix, s_push, value, pds));
      ZK_REQUIRE(res, BC_IS_CODE, row, fr_eq_u64(is_code, fr_is_zero(pdl) ? 1 : 0));
      if (n_hdr) {
        ZK_REQUIRE(res, BC_B2H_INDEX, row, fr_eq(fr_add_u64(index, 1), len));
        Fr kkey[5] = {fr_u64(2), rlc, len, hash_lo, hash_hi};
        ZK_REQUIRE(res, BC_B2H_KECCAK, row, lookup<5>(kec_ix, kkey, &hit) >= 1);
      }
    }
    if ((is_hdr || is_byte) && n_byte) {
      const Fr nlen = wcell_l<LAYOUT>(w, C_LEN, j), nindex = wcell_l<LAYOUT>(w, C_INDEX, j);
      const Fr nvalue = wcell_l<LAYOUT>(w, C_VALUE, j), nrlc = wcell_l<LAYOUT>(w, C_RLC, j);
      const bool hash_same =
          fr_eq(wcell_l<LAYOUT>(w, C_HASH_LO, j), hash_lo) && fr_eq(wcell_l<LAYOUT>(w, C_HASH_HI, j), hash_hi);
      if (is_hdr) {
        ZK_REQUIRE(res, BC_H2B_LEN, row, fr_eq(nlen, len));
        ZK_REQUIRE(res, BC_H2B_INDEX0, row, fr_is_zero(nindex));
        ZK_REQUIRE(res, BC_H2B_ISCODE, row, fr_eq_u64(wcell_l<LAYOUT>(w, C_ISCODE, j), 1));
        ZK_REQUIRE(res, BC_H2B_HASH, row, hash_same);
        ZK_REQUIRE(res, BC_H2B_RLC, row, fr_eq(nrlc, nvalue));
      } else {
        ZK_REQUIRE(res, BC_B2B_LEN, row, fr_eq(nlen, len));
        ZK_REQUIRE(res, BC_B2B_INDEX, row, fr_eq(nindex, fr_add_u64(index, 1)));
        ZK_REQUIRE(res, BC_B2B_HASH, row, hash_same);
        // next.value_rlc == value_rlc * r + next.value : the one true Fr x Fr product
        ZK_REQUIRE(res, BC_B2B_RLC, row, fr_eq(nrlc, fr_add(fr_montmul(rlc, r_mont), nvalue)));
        const Fr npdl = wcell_l<LAYOUT>(w, C_PDL, j);
        const Fr want = fr_eq_u64(is_code, 1) ? pds : fr_sub_u64(pdl, 1);
        ZK_REQUIRE(res, BC_B2B_PUSH_LEFT, row, fr_eq(npdl, want));
      }
    }
  }
  if (fr_eq_u64(q_last, 1)) {
    ZK_REQUIRE(res, BC_LAST_TAG, row, is_hdr);
    ZK_REQUIRE(res, BC_LAST_LEN0, row, fr_is_zero(len));
    ZK_REQUIRE(res, BC_LAST_EMPTY_HASH, row, hash_empty);
  }
}

#ifdef __CUDACC__
template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_check_bytecode(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ IndexDev push_ix, const __grid_constant__ IndexDev kec_ix, const __grid_constant__ Fr r_mont,
                 const __grid_constant__ ResultDev res) {
  __shared__ alignas(32) u64 s_push[256 * 4];
  __shared__ alignas(8) u64 s_bar;
  // uniform: regular 256-row push table with a canonical push_size column -> stage it (8 KiB, UBLKCP)
  const bool staged = push_ix.tab.n_rows == 256 && push_ix.tab.width[1] == 32 && pos_enabled(push_ix) &&
                      push_ix.pos_kind == ZK_POS_DENSE && fr_is_zero(table_cell(push_ix.tab, 0, 0));
  if (staged) stage_to_smem(s_push, push_ix.tab.base + push_ix.tab.off[1], sizeof(s_push), &s_bar);
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = rg.row_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rg.row_end; i += stride)
    check_bytecode_row<LAYOUT>(w, rg, push_ix, staged ? s_push : nullptr, kec_ix, r_mont, res, i);
}

#endif

}  // namespace zk

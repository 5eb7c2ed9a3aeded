// circuit.cuh — shared kernel-side plumbing: resident witness matrices and row addressing.
#pragma once
#include "lookup.cuh"

namespace zk {

struct WitnessDev {
  const u64* cells;            // [n_cols][n_rows][4]
  u64 n_rows;                  // resident rows (incl. halos when sharded)
  const unsigned char* flags;  // optional per-row type flags
};

struct CheckRange {
  u64 row_begin, row_end;  // local rows to check
  u64 row_base;            // reported row = row_base + local
  u32 flags;               // ZK_FLAG_*
};

ZK_HD Fr wcell(const WitnessDev& w, u32 col, u64 row) {
  return ld_cell(w.cells + ((u64)col * w.n_rows + row) * 4);
}
// rotation by +k / -k: wraps modulo n_rows when the whole circuit is resident, otherwise the
// caller supplied halo rows (include/zkcheck.h)
ZK_HD u64 rot_fwd(const WitnessDev& w, u64 row, u32 k, bool wrap) {
  u64 j = row + k;
  if (j >= w.n_rows) j = wrap ? j % w.n_rows : w.n_rows - 1;
  return j;
}
ZK_HD u64 rot_back(const WitnessDev& w, u64 row, bool wrap) {
  return row ? row - 1 : (wrap ? w.n_rows - 1 : 0);
}

}  // namespace zk

// circuit.cuh — shared kernel-side plumbing: resident witness matrices and row addressing.
#pragma once
#include "lookup.cuh"

namespace zk {

#define ZK_MAX_COLS 64
struct WitnessDev {
  const unsigned char* base;   // column c: n_rows integers of width[c] bytes at base + off[c]
  u64 n_rows;                  // resident rows (incl. halos when sharded)
  const unsigned char* flags;  // optional per-row type flags
  u64 off[ZK_MAX_COLS];
  unsigned char width[ZK_MAX_COLS];  // 0 (constant column), 1, 2, 4, 8, 16 or 32 (fr.cuh:ld_col)
};

struct CheckRange {
  u64 row_begin, row_end;  // local rows to check
  u64 row_base;            // reported row = row_base + local
  u32 flags;               // ZK_FLAG_*
};

ZK_HD Fr wcell(const WitnessDev& w, u32 col, u64 row) {
  return ld_col(w.base + w.off[col], w.width[col], row);
}
// Storage layout a row-checker instance is compiled for.  L_CANON: every column holds canonical 32-byte
// cells (zk_upload_columns) — a cell load is one address add + one LDG.256; L_ANY: per-column widths
// (packed uploads) through the generic branch-free loader, ~20 more instructions per load.  The host
// picks the instance from the resident matrix's widths; results are identical.
enum { L_ANY = 0, L_CANON = 1 };
template <int LAYOUT>
ZK_HD Fr wcell_l(const WitnessDev& w, u32 col, u64 row) {
  if (LAYOUT == L_CANON) return ld_cell((const u64*)(w.base + w.off[col]) + row * 4);
  return ld_col(w.base + w.off[col], w.width[col], row);
}
template <int LAYOUT>
ZK_HD Fr tcell_l(const TableDev& t, u32 col, u64 row) {
  if (LAYOUT == L_CANON) return ld_cell((const u64*)(t.base + t.off[col]) + row * 4);
  return ld_col(t.base + t.off[col], t.width[col], row);
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

#ifdef __CUDACC__
// ---- shared-memory row tiles (canonical storage) ---------------------------------------------
// The row checkers are latency-bound when every thread loads its cells straight from HBM: occupancy
// is capped by the registers of the gate program, so too few loads are in flight (ncu: 17-24 % warps
// active, 32-41 % of DRAM throughput).  For canonical matrices the columns of a block of rows are
// instead streamed into shared memory by the bulk-copy engine (cp.async.bulk, one copy per column
// per tile, completion on an mbarrier) several tiles ahead of the compute warps, which then read
// their cells with LDS.  A TileDev is the view of one staged tile with the interface of WitnessDev.
struct TileDev {
  u32 smem;   // shared-window address of the stage buffer: column c at smem + c * cap * 32
  u32 cap;    // rows per column segment
  u64 row0;   // first staged row
  u64 n_rows; // resident rows of the whole matrix (for the rotations)
  const unsigned char* flags;
};
template <int LAYOUT>
__device__ __forceinline__ Fr wcell_l(const TileDev& t, u32 col, u64 row) {
  Fr r;
  const u32 a = t.smem + (col * t.cap + (u32)(row - t.row0)) * 32u;
  asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(r.l[0]), "=l"(r.l[1]) : "r"(a));
  asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(r.l[2]), "=l"(r.l[3]) : "r"(a + 16u));
  return r;
}
__device__ __forceinline__ u64 rot_fwd(const TileDev& t, u64 row, u32 k, bool) { return row + k; }  // tiles never serve edge rows
__device__ __forceinline__ u64 rot_back(const TileDev& t, u64 row, bool) { return row - 1; }

// NCOLS columns, R rows per tile plus PRE / POST halo rows, STAGES tiles in flight
template <int NCOLS, int R, int PRE, int POST, int STAGES>
struct TilePipe {
  static constexpr int CAP = R + PRE + POST;
  static constexpr u32 STAGE_BYTES = (u32)NCOLS * CAP * 32u;
  static constexpr u32 SMEM_BYTES = STAGE_BYTES * STAGES;
  // staged rows of the tile whose first checked row is r0: [lo, hi)
  __device__ static __forceinline__ void range(u64 r0, u64 n_rows, u64* lo, u64* hi) {
    *lo = r0 >= (u64)PRE ? r0 - PRE : 0;
    *hi = r0 + R + POST < n_rows ? r0 + R + POST : n_rows;
  }
  // called by the 32 lanes of ONE warp: lane l issues the copies of columns l, l + 32, ...
  __device__ static __forceinline__ void issue(const WitnessDev& w, u32 stage_smem, u32 bar, u64 r0) {
    u64 lo, hi;
    range(r0, w.n_rows, &lo, &hi);
    const u32 bytes = (u32)(hi - lo) * 32u;
    const unsigned lane = threadIdx.x & 31;
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * NCOLS) : "memory");
    __syncwarp();
    for (int c = lane; c < NCOLS; c += 32) {
      const unsigned char* src = w.base + w.off[c] + lo * 32;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       stage_smem + (u32)c * CAP * 32u),
                   "l"(src), "r"(bytes), "r"(bar)
                   : "memory");
    }
  }
};
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "ZK_TILE_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra ZK_TILE_WAIT;\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// Stage a small read-only table into shared memory with ONE bulk asynchronous copy
// (cp.async.bulk, the 1-D TMA path: SASS UBLKCP) completing on an mbarrier.  Called by every
// thread of the block; returns when the bytes are visible to all of them.  `bytes` must be a
// multiple of 16 and both addresses 16-byte aligned.
__device__ __forceinline__ void stage_to_smem(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
  const u32 bar_a = (u32)__cvta_generic_to_shared(bar);
  const u32 dst_a = (u32)__cvta_generic_to_shared(smem_dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
        "l"(gmem_src), "r"(bytes), "r"(bar_a)
        : "memory");
  }
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "ZK_STAGE_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@!p bra ZK_STAGE_WAIT;\n\t"
      "}" ::"r"(bar_a)
      : "memory");
}
#endif

}  // namespace zk

/*
 * oracle/lookup.h — TEST INFRASTRUCTURE (CPU oracle).
 *
 * Restates lookup()/TableRow.match (src/zkevm_specs/evm_circuit/table.py:864-884,
 * 389-401): exact match of the queried (non-None) columns over a SET of rows;
 * 0 matches => unsat, >1 => ambiguous.  The reference scans linearly; here the rows are
 * sorted once on the queried columns and probed by binary search (same answers, the
 * scan order never matters because the result depends only on the match count).
 * A table is a set: rows identical in EVERY column count once.
 */
#ifndef ORACLE_LOOKUP_H
#define ORACLE_LOOKUP_H
#include <stdint.h>
#include <stdlib.h>
#include "fr.h"

typedef struct {
  const uint64_t* cells; /* [n_cols][n_rows][4] */
  uint64_t n_rows;
  uint32_t n_cols;
  uint32_t n_key;
  uint32_t key_cols[16];
  uint32_t* order; /* row ids sorted by key */
} orc_index;

#define ORC_CELL(tab, nrows, c, i) ((tab) + (((uint64_t)(c) * (nrows) + (i)) * 4))

static __thread const orc_index* g_sort_ix;
static int orc_cmp_rows(const void* pa, const void* pb) {
  const orc_index* ix = g_sort_ix;
  uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
  for (uint32_t k = 0; k < ix->n_key; k++) {
    int c = fr_cmp(fr_load(ORC_CELL(ix->cells, ix->n_rows, ix->key_cols[k], a)),
                   fr_load(ORC_CELL(ix->cells, ix->n_rows, ix->key_cols[k], b)));
    if (c) return c;
  }
  return 0;
}
static inline void orc_index_build(orc_index* ix, const uint64_t* cells, uint64_t n_rows,
                                   uint32_t n_cols, const uint32_t* key_cols, uint32_t n_key) {
  ix->cells = cells; ix->n_rows = n_rows; ix->n_cols = n_cols; ix->n_key = n_key;
  for (uint32_t k = 0; k < n_key; k++) ix->key_cols[k] = key_cols[k];
  ix->order = (uint32_t*)malloc(sizeof(uint32_t) * (n_rows ? n_rows : 1));
  for (uint64_t i = 0; i < n_rows; i++) ix->order[i] = (uint32_t)i;
  g_sort_ix = ix;
  qsort(ix->order, n_rows, sizeof(uint32_t), orc_cmp_rows);
}
static inline void orc_index_free(orc_index* ix) { free(ix->order); ix->order = 0; }

static inline int orc_key_cmp(const orc_index* ix, uint32_t row, const fr_t* key) {
  for (uint32_t k = 0; k < ix->n_key; k++) {
    int c = fr_cmp(fr_load(ORC_CELL(ix->cells, ix->n_rows, ix->key_cols[k], row)), key[k]);
    if (c) return c;
  }
  return 0;
}
static inline int orc_rows_identical(const orc_index* ix, uint32_t a, uint32_t b) {
  for (uint32_t c = 0; c < ix->n_cols; c++)
    if (!fr_eq(fr_load(ORC_CELL(ix->cells, ix->n_rows, c, a)),
               fr_load(ORC_CELL(ix->cells, ix->n_rows, c, b)))) return 0;
  return 1;
}
/* returns number of DISTINCT matching rows capped at 2; *row = first match */
static inline int orc_lookup(const orc_index* ix, const fr_t* key, uint32_t* row) {
  uint64_t lo = 0, hi = ix->n_rows;
  while (lo < hi) {
    uint64_t mid = (lo + hi) / 2;
    if (orc_key_cmp(ix, ix->order[mid], key) < 0) lo = mid + 1; else hi = mid;
  }
  if (lo >= ix->n_rows || orc_key_cmp(ix, ix->order[lo], key) != 0) return 0;
  uint32_t first = ix->order[lo];
  if (row) *row = first;
  for (uint64_t j = lo + 1; j < ix->n_rows && orc_key_cmp(ix, ix->order[j], key) == 0; j++)
    if (!orc_rows_identical(ix, first, ix->order[j])) return 2;
  return 1;
}
/* the same on the first n_prefix key columns only (the remaining key columns are not queried): the index is sorted
 * on all its key columns, so the rows that match a prefix are contiguous */
static inline int orc_prefix_cmp(const orc_index* ix, uint32_t row, const fr_t* key, uint32_t n_prefix) {
  for (uint32_t k = 0; k < n_prefix; k++) {
    int c = fr_cmp(fr_load(ORC_CELL(ix->cells, ix->n_rows, ix->key_cols[k], row)), key[k]);
    if (c) return c;
  }
  return 0;
}
static inline int orc_lookup_prefix(const orc_index* ix, const fr_t* key, uint32_t n_prefix, uint32_t* row) {
  uint64_t lo = 0, hi = ix->n_rows;
  while (lo < hi) {
    uint64_t mid = (lo + hi) / 2;
    if (orc_prefix_cmp(ix, ix->order[mid], key, n_prefix) < 0) lo = mid + 1; else hi = mid;
  }
  if (lo >= ix->n_rows || orc_prefix_cmp(ix, ix->order[lo], key, n_prefix) != 0) return 0;
  uint32_t first = ix->order[lo];
  if (row) *row = first;
  for (uint64_t j = lo + 1; j < ix->n_rows && orc_prefix_cmp(ix, ix->order[j], key, n_prefix) == 0; j++)
    if (!orc_rows_identical(ix, first, ix->order[j])) return 2;
  return 1;
}
#endif

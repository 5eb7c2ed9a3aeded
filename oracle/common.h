/* oracle/common.h — TEST INFRASTRUCTURE (CPU oracle): result recording shared by the
 * per-circuit restatements. */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H
#include <stdint.h>
#include "fr.h"
#include "lookup.h"
#include "../include/zk_constraints.h"

typedef struct {
  uint32_t* first_fail;
  uint64_t* fail_count; /* may be NULL */
  uint32_t n;
} orc_result;

static inline void orc_result_init(orc_result* r, uint32_t* ff, uint64_t* fc, uint32_t n) {
  r->first_fail = ff; r->fail_count = fc; r->n = n;
  for (uint32_t i = 0; i < n; i++) { ff[i] = 0xFFFFFFFFu; if (fc) fc[i] = 0; }
}
static inline void orc_fail(orc_result* r, int id, uint64_t row) {
  if ((uint32_t)row < r->first_fail[id]) r->first_fail[id] = (uint32_t)row;
  if (r->fail_count) r->fail_count[id]++;
}
#define REQUIRE(res, id, row, cond) do { if (!(cond)) orc_fail((res), (id), (row)); } while (0)
#endif

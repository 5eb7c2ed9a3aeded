/* oracle/catalogue.c — TEST INFRASTRUCTURE: constraint counts and error classes as the
 * oracle sees them (from include/zk_constraints.h). */
#include "../include/zk_constraints.h"
#define CLS(id, cls, doc) cls,
static const int kBytecode[] = {ZK_BYTECODE_CONSTRAINTS(CLS)};
static const int kEvm[] = {ZK_EVM_CONSTRAINTS(CLS)};
static const int kCopy[] = {ZK_COPY_CONSTRAINTS(CLS)};
static const int kState[] = {ZK_STATE_CONSTRAINTS(CLS)};
static const int kExp[] = {ZK_EXP_CONSTRAINTS(CLS)};
static const int kTx[] = {ZK_TX_CONSTRAINTS(CLS)};
static const int kSig[] = {ZK_SIG_CONSTRAINTS(CLS)};
static const int kPi[] = {ZK_PI_CONSTRAINTS(CLS)};
int orc_n_constraints(int circuit) {
  switch (circuit) {
    case 0: return BC_N_CONSTRAINTS;
    case 1: return ST_N_CONSTRAINTS;
    case 2: return CP_N_CONSTRAINTS;
    case 3: return EV_N_CONSTRAINTS;
    case 4: return XP_N_CONSTRAINTS;
    case 5: return TX_N_CONSTRAINTS;
    case 6: return SG_N_CONSTRAINTS;
    case 7: return PI_N_CONSTRAINTS;
    default: return 0;
  }
}
int orc_constraint_class(int circuit, int idx) {
  if (idx < 0 || idx >= orc_n_constraints(circuit)) return -1;
  if (circuit == 4) return kExp[idx];
  if (circuit == 5) return kTx[idx];
  if (circuit == 6) return kSig[idx];
  if (circuit == 7) return kPi[idx];
  return circuit == 0 ? kBytecode[idx] : circuit == 1 ? kState[idx] : circuit == 2 ? kCopy[idx] : kEvm[idx];
}

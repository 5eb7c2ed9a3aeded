// pi.cu — public-inputs-circuit row checker (one thread per row).
//
// Replaces the loop around check_row, src/zkevm_specs/pi_circuit.py:150-321 (driver: verify_circuit :447-459,
// row_next = rows[(i + 1) % len(rows)]).  Row = 28 cells in the order of pi_circuit.Row (:105-133) with the
// tx-table and withdrawal-table rows a Row carries flattened behind it (include/zkcheck.h ZK_CIRCUIT_PI);
// rotation {0,+1}.  Algorithmic bytes: 28 x 32 B = 896 B per row + the keccak / gas-cost table rows.
// This is the one circuit of the reference already written as `selector * polynomial == 0` with inverse
// witnesses.  The four byte-accumulator gates every row evaluates are decided without the selector products
// (Fr is an integral domain: a product is zero iff a factor is), which leaves ONE Fr x Fr product per gate
// (next accumulator x challenge); the sections under `if selector != 0` (one row per tx field / calldata
// byte / withdrawal) evaluate the reference's polynomials as written.  A row stops at its first failure.
#include "circuit.cuh"
// fr_mul_sel comes from copy.cu (this file is part of the unity build of api.cu, after copy.cu)
#include "../../include/zk_constraints.h"
#include "../../include/zkcheck.h"

namespace zk {

enum { P_Q_BYTES_LAST, P_Q_TX_TABLE, P_Q_TX_CALLDATA, P_Q_TX_CALLDATA_START, P_Q_KECCAK, P_Q_VALUE_START, P_TXID_INV,
       P_VALUE_LO_INV, P_TXID_DIFF_INV, P_CD_GAS, P_IS_FINAL, P_Q_WD, P_BYTES, P_KRLC, P_VALUE_LC, P_DIGEST_LO,
       P_DIGEST_HI, P_Q_BYTE_EN, P_TX_ID, P_TX_TAG, P_TX_INDEX, P_TX_VAL_LO, P_TX_VAL_HI, P_WD_ID, P_WD_VALIDATOR,
       P_WD_ADDR_LO, P_WD_ADDR_HI, P_WD_AMOUNT, PI_COLS };
#define PI_TAG_CALLDATA_LENGTH 8  // TxContextFieldTag.CallDataLength (evm_circuit/table.py:147-167)
#define PI_GAS_NONZERO_BYTE 16    // GAS_COST_TX_CALL_DATA_PER_NON_ZERO_BYTE (util/param.py:64)
#define PI_GAS_ZERO_BYTE 4        // GAS_COST_TX_CALL_DATA_PER_ZERO_BYTE (util/param.py:66)

struct PiParams {
  Fr keccak_rand_mont;  // pi_circuit.keccak_rand (:836), Montgomery form
  Fr byte_base_mont;    // pi_circuit.byte_pow_base (:834), Montgomery form
  Fr circuit_len;       // Witness.circuit_len, canonical
};

#define PI_CHECK(id, cond)   \
  do {                       \
    if (!(cond)) {           \
      fail(res, (id), row);  \
      return;                \
    }                        \
  } while (0)

template <int LAYOUT>
ZK_HD void check_pi_row(const WitnessDev& w, const CheckRange& rg, const IndexDev& kec_ix, const IndexDev& gas_ix,
                        const PiParams& pp, const ResultDev& res, u64 i) {
  const bool wrap = rg.flags & ZK_FLAG_WRAP;
  const u64 j = rot_fwd(w, i, 1, wrap);
  const u64 row = rg.row_base + i;
#define C(c) wcell_l<LAYOUT>(w, (c), i)
#define N(c) wcell_l<LAYOUT>(w, (c), j)
#define M fr_mul
#define Z fr_is_zero
  const Fr one = fr_u64(1);
  const Fr en = C(P_Q_BYTE_EN), last = C(P_Q_BYTES_LAST), bytes = C(P_BYTES), krlc = C(P_KRLC), vlc = C(P_VALUE_LC);
  const Fr vstart = C(P_Q_VALUE_START), qk = C(P_Q_KECCAK);
  const Fr q_cd = C(P_Q_TX_CALLDATA), q_tx = C(P_Q_TX_TABLE), q_wd = C(P_Q_WD);
  const bool en0 = Z(en);
  // :162  en * last * (krlc - bytes)
  PI_CHECK(PI_RLC_LAST, en0 || Z(last) || fr_eq(krlc, bytes));
  // :165-170  en * (1 - last) * (krlc - (next.krlc * keccak_rand + bytes))
  PI_CHECK(PI_RLC_ACC, en0 || fr_eq_u64(last, 1) || fr_eq(krlc, fr_add(fr_montmul(N(P_KRLC), pp.keccak_rand_mont), bytes)));
  // :183-188  en * (1 - value_start) * (vlc - (next.vlc * byte_pow_base + bytes))
  PI_CHECK(PI_VALUE_ACC, en0 || fr_eq_u64(vstart, 1) || fr_eq(vlc, fr_add(fr_montmul(N(P_VALUE_LC), pp.byte_base_mont), bytes)));
  // :191-194  en * value_start * (vlc - bytes)
  PI_CHECK(PI_VALUE_START, en0 || Z(vstart) || fr_eq(vlc, bytes));
  {  // :197-203  the tuple (q, q*rlc, q*circuit_len, digest.select(q)) is a member of the keccak table
    Fr key[5];
    key[0] = qk;
    if (Z(qk)) {
      key[1] = key[2] = key[3] = key[4] = fr_u64(0);
    } else {
      const bool q1 = fr_eq_u64(qk, 1);
      key[1] = q1 ? krlc : M(qk, krlc);
      key[2] = q1 ? pp.circuit_len : M(qk, pp.circuit_len);
      key[3] = q1 ? C(P_DIGEST_LO) : M(qk, C(P_DIGEST_LO));
      key[4] = q1 ? C(P_DIGEST_HI) : M(qk, C(P_DIGEST_HI));
      PI_CHECK(PI_KECCAK_WORD, fr_fits128(key[3]) && fr_fits128(key[4]));
    }
    u32 hit;
    PI_CHECK(PI_KECCAK_LOOKUP, lookup<5>(kec_ix, key, &hit) >= 1);
  }
  if (!Z(q_cd)) {  // :207-294  one row per calldata byte
    const Fr tx_id = C(P_TX_ID), n_tx_id = N(P_TX_ID), tx_id_inv = C(P_TXID_INV), vlo = C(P_TX_VAL_LO);
    const Fr vlo_inv = C(P_VALUE_LO_INV), diff_inv = C(P_TXID_DIFF_INV), diff = fr_sub(n_tx_id, tx_id);
    const Fr nz = M(tx_id, tx_id_inv), neq = M(diff, diff_inv), b_nz = M(vlo, vlo_inv);
    PI_CHECK(PI_CD_TXID_INV, Z(tx_id) || fr_eq_u64(nz, 1));
    PI_CHECK(PI_CD_VALUE_INV, Z(vlo) || fr_eq_u64(b_nz, 1));
    PI_CHECK(PI_CD_DIFF_INV, Z(diff) || fr_eq_u64(neq, 1));
    const Fr n_nz = M(n_tx_id, N(P_TXID_INV)), nb_nz = M(N(P_TX_VAL_LO), N(P_VALUE_LO_INV));
    const Fr zr = fr_sub(one, nz), n_zr = fr_sub(one, n_nz), eq = fr_sub(one, neq);
    const Fr cg = C(P_CD_GAS), ncg = N(P_CD_GAS), fin = C(P_IS_FINAL), idx = C(P_TX_INDEX), n_idx = N(P_TX_INDEX);
    const bool zr0 = Z(zr);
    PI_CHECK(PI_CD_DEF_TXID, zr0 || Z(tx_id));
    PI_CHECK(PI_CD_DEF_NEXT_TXID, zr0 || Z(n_tx_id));
    PI_CHECK(PI_CD_DEF_FINAL, zr0 || Z(fin));
    PI_CHECK(PI_CD_DEF_GAS, zr0 || Z(cg));
    // gas_cost = 16 * is_byte_nonzero + 4 * (1 - is_byte_nonzero) = 4 + 12 * is_byte_nonzero
    // (is_byte_nonzero is 0 or 1 on every witness whose inverse cells are right: fr_mul_sel skips those products)
    const Fr gas = fr_add(fr_u64(PI_GAS_ZERO_BYTE), fr_mul_sel(fr_u64(PI_GAS_NONZERO_BYTE - PI_GAS_ZERO_BYTE), b_nz));
    const Fr n_gas = fr_add(fr_u64(PI_GAS_ZERO_BYTE), fr_mul_sel(fr_u64(PI_GAS_NONZERO_BYTE - PI_GAS_ZERO_BYTE), nb_nz));
    {  // :250-256 lookup(FixedU16Row): the table is the integers 0..65535
      const Fr v = fr_mul_sel(fr_mul_sel(neq, n_nz), fr_sub(diff, one));
      PI_CHECK(PI_CD_U16, fr_fits64(v) && v.l[0] < 65536);
    }
    const bool nz0 = Z(nz), eq0 = Z(eq), diff0 = Z(diff);
    PI_CHECK(PI_CD_IDX_SAME, nz0 || eq0 || fr_eq(n_idx, fr_add(idx, one)));
    PI_CHECK(PI_CD_IDX_NEXT, nz0 || diff0 || Z(n_idx));
    PI_CHECK(PI_CD_GAS_SAME, nz0 || eq0 || fr_eq(ncg, fr_add(cg, n_gas)));
    PI_CHECK(PI_CD_GAS_NEXT, nz0 || Z(n_nz) || diff0 || fr_eq(ncg, n_gas));
    PI_CHECK(PI_CD_GAS_LAST, nz0 || Z(n_zr) || Z(ncg));
    PI_CHECK(PI_CD_FINAL_SAME, nz0 || eq0 || Z(fin));
    PI_CHECK(PI_CD_FINAL_NEXT, nz0 || diff0 || fr_eq_u64(fin, 1));
    const bool qs0 = Z(C(P_Q_TX_CALLDATA_START));
    PI_CHECK(PI_CD_START_INDEX, qs0 || nz0 || Z(idx));
    PI_CHECK(PI_CD_START_GAS, qs0 || nz0 || fr_eq(cg, gas));
  }
  if (!Z(q_tx)) {  // :296-318  one row per tx-table field
    const Fr is_cdl = fr_sub_u64(C(P_TX_TAG), PI_TAG_CALLDATA_LENGTH), inv = C(P_TXID_INV), vlo = C(P_TX_VAL_LO);
    const Fr t = M(is_cdl, inv), len_nz = M(vlo, C(P_VALUE_LO_INV));
    PI_CHECK(PI_TX_CDL_INV, Z(is_cdl) || fr_eq_u64(t, 1));
    PI_CHECK(PI_TX_VALUE_INV, Z(vlo) || fr_eq_u64(len_nz, 1));
    const Fr cdl_row = fr_sub(one, t), len_z = fr_sub(one, len_nz), cost = N(P_TX_VAL_LO);
    PI_CHECK(PI_TX_ZERO_COST, Z(cdl_row) || Z(len_z) || Z(cost));
    const Fr qc = fr_mul_sel(cdl_row, len_nz);
    Fr key[3] = {fr_mul_sel(C(P_TX_ID), qc), qc, fr_mul_sel(cost, qc)};
    u32 hit;
    const int n = lookup<3>(gas_ix, key, &hit);
    PI_CHECK(PI_TX_GAS_LOOKUP, n >= 1);
    PI_CHECK(PI_TX_GAS_AMBIG, n <= 1);
  }
  if (!Z(q_wd)) {  // :320-323
    if (!Z(N(P_Q_WD))) PI_CHECK(PI_WD_NEXT_ID, fr_eq(N(P_WD_ID), fr_add(C(P_WD_ID), one)));
    PI_CHECK(PI_WD_AMOUNT, !Z(C(P_WD_AMOUNT)));
  }
#undef C
#undef N
#undef M
#undef Z
}

#ifdef __CUDACC__
template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_check_pi(const __grid_constant__ WitnessDev w, const __grid_constant__ CheckRange rg, const __grid_constant__ IndexDev kec_ix,
           const __grid_constant__ IndexDev gas_ix, const __grid_constant__ PiParams pp, const __grid_constant__ ResultDev res) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = rg.row_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < rg.row_end; i += stride)
    check_pi_row<LAYOUT>(w, rg, kec_ix, gas_ix, pp, res, i);
}
#endif

}  // namespace zk

/* TEST INFRASTRUCTURE — CPU oracle, part of oracle/evm.c (included there).
 * ADDMOD / MULMOD / SDIV_SMOD / SAR:
 *   addmod     evm_circuit/execution/addmod.py:21-68   (mul_add_words instruction.py:599-632, mul_add_words_512 :634-665)
 *   mulmod     evm_circuit/execution/mulmod.py:6-73
 *   sdiv_smod  evm_circuit/execution/sdiv_smod.py:6-117 (abs_word instruction.py:539-569)
 *   sar        evm_circuit/execution/sar.py:12-194
 * These gadgets DERIVE their witness (quotients, reduced operands, limb splits) from the stack words with
 * Python integer arithmetic, so for stack words in the halves domain (lo, hi < 2^128) every field identity they
 * state reduces to a statement about 256- / 512-bit integers:
 *   - carry = (low part of a*b + c - d) / 2^128 over the field passes range_check(.., 9) exactly when the
 *     128-bit slice of the integer a*b + c equals the slice of d (a non-multiple of 2^128 of magnitude < 2^200
 *     divided by 2^128 mod p never fits 72 bits; a multiple gives an integer < 2^67);
 *   - `overflow` (a sum of non-negative integers) is zero exactly when a*b + c < 2^256.
 * Words outside the halves domain (never produced by a state-circuit-checked rw table) are reported as
 * EV_AR_WITNESS_DOMAIN, at the step the reference fails on.
 * Pinned by tests/golden/evm16.npz (1,702 verdicts of the reference's verify_step).
 */
typedef struct { uint64_t l[4]; } u256;
static u256 w_int(word_t w) { u256 r = {{w.lo.l[0], w.lo.l[1], w.hi.l[0], w.hi.l[1]}}; return r; }
static int u256_is_zero(u256 a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static int u256_cmp(u256 a, u256 b) {
  for (int k = 3; k >= 0; k--) { if (a.l[k] < b.l[k]) return -1; if (a.l[k] > b.l[k]) return 1; }
  return 0;
}
static u256 u256_neg(u256 a) { /* 2^256 - a (0 for 0) */
  u256 r; uint64_t c = 1;
  for (int k = 0; k < 4; k++) { const u128 t = (u128)(~a.l[k]) + c; r.l[k] = (uint64_t)t; c = (uint64_t)(t >> 64); }
  return r;
}
static int u256_neg_sign(u256 a) { return (int)(a.l[3] >> 63); }
static u256 u256_abs(u256 a) { return u256_neg_sign(a) ? u256_neg(a) : a; }
/* out[0..8] = a * b + c */
static void u256_mul_add(u256 a, u256 b, u256 c, uint64_t out[9]) {
  for (int k = 0; k < 9; k++) out[k] = k < 4 ? c.l[k] : 0;
  for (int x = 0; x < 4; x++) {
    uint64_t carry = 0;
    for (int y = 0; y < 4; y++) {
      const u128 t = (u128)a.l[x] * b.l[y] + out[x + y] + carry;
      out[x + y] = (uint64_t)t; carry = (uint64_t)(t >> 64);
    }
    for (int k = x + 4; carry && k < 9; k++) { const u128 t = (u128)out[k] + carry; out[k] = (uint64_t)t; carry = (uint64_t)(t >> 64); }
  }
}
/* num (nn limbs) = quot * den + rem, den != 0: schoolbook binary long division */
static void u_divrem(const uint64_t* num, int nn, u256 den, uint64_t* quot, u256* rem) {
  uint64_t r[5] = {0, 0, 0, 0, 0};
  for (int k = 0; k < nn; k++) quot[k] = 0;
  for (int bit = nn * 64 - 1; bit >= 0; bit--) {
    for (int k = 4; k > 0; k--) r[k] = (r[k] << 1) | (r[k - 1] >> 63);
    r[0] = (r[0] << 1) | ((num[bit >> 6] >> (bit & 63)) & 1);
    int ge = r[4] != 0;
    if (!ge) { ge = 1; for (int k = 3; k >= 0; k--) { if (r[k] != den.l[k]) { ge = r[k] > den.l[k]; break; } } }
    if (ge) {
      uint64_t br = 0;
      for (int k = 0; k < 4; k++) { const u128 d = (u128)r[k] - den.l[k] - br; r[k] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
      r[4] -= br;
      quot[bit >> 6] |= 1ull << (bit & 63);
    }
  }
  for (int k = 0; k < 4; k++) rem->l[k] = r[k];
}
/* mul_add_words(a, b, c, d) on halves-domain words: 0 = holds with overflow 0, 1 = range_check(carry_lo),
 * 2 = range_check(carry_hi), 3 = holds with a non-zero overflow (instruction.py:599-632) */
static int mul_add_verdict(u256 a, u256 b, u256 c, u256 d) {
  uint64_t p[9]; u256_mul_add(a, b, c, p);
  if (p[0] != d.l[0] || p[1] != d.l[1]) return 1;
  if (p[2] != d.l[2] || p[3] != d.l[3]) return 2;
  return (p[4] | p[5] | p[6] | p[7] | p[8]) ? 3 : 0;
}
#define AR_STACK(k, rw_, sp_off, out) \
  if (!need1(e, rw_lookup(e, fr_add(rwc, fr_u64(k)), (rw_), ZK_TARGET_Stack, call_id, fr_add(sp, fr_u64(sp_off)), (out)), EV_AR_RW0_UNSAT + 2 * (k), row)) return

static void gadget_addmod_mulmod(evm_env* e, uint64_t i, uint64_t row, fr_t opcode, int is_mul) {
  const uint64_t* S = e->steps; const uint64_t n_ = e->n_steps; const uint64_t n = n_;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  CHECK(EV_AR_OPCODE, fr_eq_u64(opcode, is_mul ? 0x09 : 0x08)); /* addmod.py:23, mulmod.py:33 */
  word_t aw, bw, nw, rw_;
  AR_STACK(0, 0, 0, &aw); AR_STACK(1, 0, 1, &bw); AR_STACK(2, 0, 2, &nw); AR_STACK(3, 1, 2, &rw_);
  CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(nw));
  const u256 a = w_int(aw), b = w_int(bw), m = w_int(nw), r = w_int(rw_);
  if (u256_is_zero(m)) {
    /* n == 0: the witness is all zeros and the second operand is not looked at by ADDMOD (addmod.py:33-37,60);
     * mulmod.py:56 asserts a_reduced * b == k * n + r with a_reduced = k = 0, then mul_add_words_512 splits b
     * into 64-bit limbs (instruction.py:639) */
    if (is_mul) {
      CHECK(EV_AR_MULMOD_R, fr_is_zero(rw_.lo) && fr_is_zero(rw_.hi));
      CHECK(EV_AR_MULMOD_TO64, word_in_domain(bw));
    } else {
      CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(aw));
      CHECK(EV_AR_ADDMOD_ZERO, fr_is_zero(rw_.lo) && fr_is_zero(rw_.hi));
    }
  } else {
    CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(aw) && word_in_domain(bw) && word_in_domain(rw_));
    uint64_t q[8]; u256 a_red, r_true;
    u_divrem(a.l, 4, m, q, &a_red);
    if (is_mul) {
      uint64_t p[9]; const u256 zero = {{0, 0, 0, 0}};
      u256_mul_add(a_red, b, zero, p);
      u_divrem(p, 8, m, q, &r_true);
      CHECK(EV_AR_MULMOD_R, u256_cmp(r, r_true) == 0); /* mulmod.py:56; everything after it holds by construction */
    } else {
      uint64_t s[5]; uint64_t c = 0;
      for (int k = 0; k < 4; k++) { const u128 t = (u128)a_red.l[k] + b.l[k] + c; s[k] = (uint64_t)t; c = (uint64_t)(t >> 64); }
      s[4] = c;
      u_divrem(s, 5, m, q, &r_true);
      /* mul_add_words_512(d, n, r, overflow, a_reduced + b), addmod.py:47-50: carry_0 / carry_1 are integers exactly
       * when the low / high half of r is the true remainder's */
      if (r.l[0] != r_true.l[0] || r.l[1] != r_true.l[1]) { orc_fail(e->res, EV_AR_ADDMOD_CARRY0, row); return; }
      if (r.l[2] != r_true.l[2] || r.l[3] != r_true.l[3]) { orc_fail(e->res, EV_AR_ADDMOD_CARRY1, row); return; }
    }
  }
  same_context(e, i, row, opcode, 4, one, fr_u64(2));
}

static void gadget_sdiv_smod(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t p1w, p2w, pushw;
  AR_STACK(0, 0, 0, &p1w); AR_STACK(1, 0, 1, &p2w); AR_STACK(2, 1, 1, &pushw);
  const int is_sdiv = fr_eq_u64(opcode, 0x05); /* (SMOD - opcode) / 2 == 1 over the field, sdiv_smod.py:89 */
  CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(p1w) && word_in_domain(p2w));
  const u256 pop1 = w_int(p1w), pop2 = w_int(p2w);
  const int smod_by_zero = !is_sdiv && u256_is_zero(pop2); /* the pushed word is not looked at, sdiv_smod.py:106-113 */
  if (!smod_by_zero) CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(pushw));
  const u256 push = w_int(pushw), zero = {{0, 0, 0, 0}};
  const u256 pop1_abs = u256_abs(pop1), pop2_abs = u256_abs(pop2), push_abs = u256_abs(push);
  u256 quotient, remainder; const u256 divisor = pop2, dividend = pop1;
  if (is_sdiv) { /* remainder = Word(+-(|pop1| - |push| * |pop2|)), sdiv_smod.py:98-103 */
    uint64_t p[9]; u256_mul_add(push_abs, pop2_abs, zero, p);
    const u256 prod = {{p[0], p[1], p[2], p[3]}};
    const int neg = (p[4] | p[5] | p[6] | p[7] | p[8]) != 0 || u256_cmp(prod, pop1_abs) > 0;
    if (neg) { orc_fail(e->res, u256_neg_sign(pop1) ? EV_AR_SDIV_REM_WORD : EV_AR_SDIV_REM_NEG, row); return; }
    u256 x; uint64_t br = 0;
    for (int k = 0; k < 4; k++) { const u128 d = (u128)pop1_abs.l[k] - prod.l[k] - br; x.l[k] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    quotient = push; remainder = u256_neg_sign(pop1) ? u256_neg(x) : x;
  } else {
    if (u256_is_zero(pop2)) { quotient = zero; remainder = pop1; }
    else {
      uint64_t q[4]; u256 rem;
      u_divrem(pop1_abs.l, 4, pop2_abs, q, &rem);
      const u256 q0 = {{q[0], q[1], q[2], q[3]}};
      quotient = (u256_neg_sign(pop1) == u256_neg_sign(pop2)) ? q0 : u256_neg(q0);
      remainder = push;
    }
  }
  /* check_witness, sdiv_smod.py:33-79 (the abs_word constraints hold by construction in the halves domain) */
  const u256 q_abs = u256_abs(quotient), r_abs = u256_abs(remainder), dd_abs = u256_abs(dividend);
  const int v = mul_add_verdict(q_abs, pop2_abs, r_abs, dd_abs);
  if (v == 1) { orc_fail(e->res, EV_AR_SDIV_CARRY_LO, row); return; }
  if (v == 2) { orc_fail(e->res, EV_AR_SDIV_CARRY_HI, row); return; }
  CHECK(EV_AR_SDIV_OVERFLOW, v == 0);
  const int q_nz = !u256_is_zero(quotient), d_nz = !u256_is_zero(divisor), r_nz = !u256_is_zero(remainder);
  CHECK(EV_AR_SDIV_REM_LT, !d_nz || u256_cmp(r_abs, pop2_abs) < 0);
  CHECK(EV_AR_SDIV_SIGN_REM, !(q_nz && d_nz && r_nz) || u256_neg_sign(dividend) == u256_neg_sign(remainder));
  const int signed_overflow = u256_neg_sign(dd_abs);
  CHECK(EV_AR_SDIV_SIGN_QUOT, !(q_nz && d_nz && !signed_overflow) ||
                                  (u256_neg_sign(quotient) ^ u256_neg_sign(divisor)) == u256_neg_sign(dividend));
  same_context(e, i, row, opcode, 3, one, one);
}

static void gadget_sar(evm_env* e, uint64_t i, uint64_t row, fr_t opcode) {
  const uint64_t* S = e->steps; const uint64_t n = e->n_steps;
  const fr_t rwc = CUR(S_RWC), call_id = CUR(S_CALL_ID), sp = CUR(S_SP), one = fr_u64(1);
  word_t shw, aw, bw;
  AR_STACK(0, 0, 0, &shw); AR_STACK(1, 0, 1, &aw); AR_STACK(2, 1, 1, &bw);
  CHECK(EV_AR_SAR_BYTES, word_in_domain(shw) && word_in_domain(aw) && word_in_domain(bw)); /* to_le_bytes / to_64s */
  const u256 sh = w_int(shw), a = w_int(aw), b = w_int(bw);
  const unsigned shf0 = (unsigned)(sh.l[0] & 0xFF), div64 = shf0 >> 6, mod64 = shf0 & 63;
  const int lt256 = (sh.l[0] >> 8) == 0 && sh.l[1] == 0 && sh.l[2] == 0 && sh.l[3] == 0;
  const int is_neg = u256_neg_sign(a);
  const uint64_t fill = is_neg ? ~0ull : 0;
  u256 want = {{fill, fill, fill, fill}};
  if (lt256) { /* gen_witness, sar.py:180-183: an arithmetic shift right by shf0 */
    for (unsigned k = 0; k + div64 < 4; k++) {
      const uint64_t lo = a.l[k + div64], hi = (k + div64 + 1 < 4) ? a.l[k + div64 + 1] : fill;
      want.l[k] = mod64 ? (lo >> mod64) | (hi << (64 - mod64)) : lo;
    }
  }
  CHECK(EV_AR_SAR_RESULT, u256_cmp(b, want) == 0); /* sar.py:79-82 b64s[idx] == the limb of the pushed word */
  { /* sar.py:142-145 sign_byte_lookup, :151-152 pow2_lookup */
    fr_t k1[4] = {fr_u64(ZK_FIXED_SignByte), fr_u64(a.l[3] >> 56), fr_u64(is_neg ? 255 : 0), fr_u64(0)};
    if (!need1(e, orc_lookup(&e->fixed_ix, k1, 0), EV_AR_SAR_SIGN_UNSAT, row)) return;
    fr_t k2[4] = {fr_u64(ZK_FIXED_Pow2), fr_u64(mod64), fr_u64(1ull << mod64), fr_u64(0)};
    if (!need1(e, orc_lookup(&e->fixed_ix, k2, 0), EV_AR_SAR_POW_LO_UNSAT, row)) return;
    fr_t p_hi = fr_u64(0);
    if (mod64 == 0) p_hi.l[1] = 1; else p_hi.l[0] = 1ull << (64 - mod64);
    fr_t k3[4] = {fr_u64(ZK_FIXED_Pow2), fr_u64(64 - mod64), p_hi, fr_u64(0)};
    if (!need1(e, orc_lookup(&e->fixed_ix, k3, 0), EV_AR_SAR_POW_HI_UNSAT, row)) return;
  }
  same_context(e, i, row, opcode, 3, one, one);
}

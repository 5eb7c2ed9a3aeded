// Gate programs of ADDMOD / MULMOD / SDIV_SMOD / SAR (group KG_ARITH), part of evm.cu (included there).
//   addmod     evm_circuit/execution/addmod.py:21-68   (mul_add_words instruction.py:599-632, mul_add_words_512 :634-665)
//   mulmod     evm_circuit/execution/mulmod.py:6-73
//   sdiv_smod  evm_circuit/execution/sdiv_smod.py:6-117 (abs_word instruction.py:539-569)
//   sar        evm_circuit/execution/sar.py:12-194
// The reference derives these gadgets' witnesses (quotients, reduced operands, limb splits) from the stack words
// with Python integers and then states field identities about them.  For stack words in the halves domain
// (lo, hi < 2^128) each identity is a statement about 256- / 512-bit integers, which is what runs here, on four
// 64-bit limbs per word and without any field multiplication:
//   carry = (128-bit slice of a*b + c - slice of d) / 2^128 passes range_check(.., 9) iff the slices are equal
//   (a non-multiple of 2^128 below 2^200 divided by 2^128 mod p never fits 72 bits), and `overflow` is zero iff
//   a*b + c < 2^256.
// Words outside the halves domain (a state-circuit-checked rw table has none) are reported as
// EV_AR_WITNESS_DOMAIN at the step.
#pragma once
namespace zk {

struct U256 {
  u64 l[4];
};
ZK_HD U256 word_u256(const Word2& w) { return U256{{w.lo.l[0], w.lo.l[1], w.hi.l[0], w.hi.l[1]}}; }
ZK_HD bool u256_zero(const U256& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
ZK_HD bool u256_eq(const U256& a, const U256& b) {
  return ((a.l[0] ^ b.l[0]) | (a.l[1] ^ b.l[1]) | (a.l[2] ^ b.l[2]) | (a.l[3] ^ b.l[3])) == 0;
}
// a - b over 2^256; *borrow = (a < b)
ZK_HD U256 u256_sub(const U256& a, const U256& b, u64* borrow) {
  U256 r;
  u64 br = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const u64 d = a.l[k] - b.l[k], d2 = d - br;
    br = (u64)(a.l[k] < b.l[k]) | (u64)(d < br);
    r.l[k] = d2;
  }
  *borrow = br;
  return r;
}
ZK_HD bool u256_lt(const U256& a, const U256& b) {
  u64 br;
  (void)u256_sub(a, b, &br);
  return br != 0;
}
ZK_HD U256 u256_neg(const U256& a) {  // 2^256 - a (0 for 0)
  u64 br;
  return u256_sub(U256{{0, 0, 0, 0}}, a, &br);
}
ZK_HD int u256_sign(const U256& a) { return (int)(a.l[3] >> 63); }
ZK_HD U256 u256_abs(const U256& a) { return u256_sign(a) ? u256_neg(a) : a; }

ZK_HD void mul64(u64 a, u64 b, u64* lo, u64* hi) {
#ifdef __CUDA_ARCH__
  *lo = a * b;
  *hi = __umul64hi(a, b);
#else
  const unsigned __int128 t = (unsigned __int128)a * b;
  *lo = (u64)t;
  *hi = (u64)(t >> 64);
#endif
}
// p[0..8] = a * b + c (operand scanning, static limb indexes only)
ZK_HD void u256_mul_add(const U256& a, const U256& b, const U256& c, u64 (&p)[9]) {
#pragma unroll
  for (int k = 0; k < 9; k++) p[k] = k < 4 ? c.l[k] : 0;
#pragma unroll
  for (int x = 0; x < 4; x++) {
    u64 carry = 0;
#pragma unroll
    for (int y = 0; y < 4; y++) {
      u64 lo, hi;
      mul64(a.l[x], b.l[y], &lo, &hi);
      const u64 s1 = p[x + y] + lo;
      hi += (u64)(s1 < lo);
      const u64 s2 = s1 + carry;
      hi += (u64)(s2 < carry);
      p[x + y] = s2;
      carry = hi;
    }
#pragma unroll
    for (int k = x + 4; k < 9; k++) {
      const u64 s1 = p[k] + carry;
      carry = (u64)(s1 < carry);
      p[k] = s1;
    }
  }
}
// num (NN limbs) = quot * den + rem for den != 0: restoring binary division, one numerator limb per outer
// (unrolled) iteration so that every array index is a compile-time constant
template <int NN>
ZK_HD void u_divrem(const u64 (&num)[NN], const U256& den, u64 (&quot)[NN], U256* rem) {
  u64 r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
  for (int limb = NN - 1; limb >= 0; limb--) {
    const u64 word = num[limb];
    u64 qw = 0;
#pragma unroll 1
    for (int b = 63; b >= 0; b--) {
      r4 = (r4 << 1) | (r3 >> 63);
      r3 = (r3 << 1) | (r2 >> 63);
      r2 = (r2 << 1) | (r1 >> 63);
      r1 = (r1 << 1) | (r0 >> 63);
      r0 = (r0 << 1) | ((word >> b) & 1);
      u64 br;
      const U256 d = u256_sub(U256{{r0, r1, r2, r3}}, den, &br);
      if (r4 != 0 || br == 0) {
        r0 = d.l[0], r1 = d.l[1], r2 = d.l[2], r3 = d.l[3];
        r4 = 0;  // r < 2 * den, so after the subtraction the 5th limb is clear
        qw |= 1ull << b;
      }
    }
    quot[limb] = qw;
  }
  *rem = U256{{r0, r1, r2, r3}};
}
// mul_add_words(a, b, c, d) on halves-domain words (instruction.py:599-632): 0 = holds with overflow 0,
// 1 = range_check(carry_lo), 2 = range_check(carry_hi), 3 = holds with a non-zero overflow
ZK_HD int mul_add_verdict(const U256& a, const U256& b, const U256& c, const U256& d) {
  u64 p[9];
  u256_mul_add(a, b, c, p);
  if (p[0] != d.l[0] || p[1] != d.l[1]) return 1;
  if (p[2] != d.l[2] || p[3] != d.l[3]) return 2;
  return (p[4] | p[5] | p[6] | p[7] | p[8]) ? 3 : 0;
}
#define AR_STACK(k, rw, sp_off, out)                                                                                        \
  do {                                                                                                                      \
    if (!need1(s, true, stack_at(s, true, (k), (rw), fr_add_u64(s.cur(S_SP), (sp_off)), (out)), EV_AR_RW0_UNSAT + 2 * (k))) \
      return;                                                                                                               \
  } while (0)

ZK_HD_NOINLINE void gadget_addmod_mulmod(const StepCtx& s, bool is_mul) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  EV_CHECK(EV_AR_OPCODE, fr_eq_u64(opcode, is_mul ? 0x09 : 0x08));
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 aw = zero, bw = zero, nw = zero, rw = zero;
  AR_STACK(0, 0, 0, &aw);
  AR_STACK(1, 0, 1, &bw);
  AR_STACK(2, 0, 2, &nw);
  AR_STACK(3, 1, 2, &rw);
  EV_CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(nw));
  const U256 m = word_u256(nw);
  if (u256_zero(m)) {
    // n == 0: an all-zero witness; ADDMOD never looks at b (addmod.py:33-37,60), MULMOD asserts r == 0 and then
    // splits b into 64-bit limbs (mulmod.py:56,62)
    if (is_mul) {
      EV_CHECK(EV_AR_MULMOD_R, fr_is_zero(rw.lo) && fr_is_zero(rw.hi));
      EV_CHECK(EV_AR_MULMOD_TO64, word_in_domain(bw));
    } else {
      EV_CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(aw));
      EV_CHECK(EV_AR_ADDMOD_ZERO, fr_is_zero(rw.lo) && fr_is_zero(rw.hi));
    }
  } else {
    EV_CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(aw) && word_in_domain(bw) && word_in_domain(rw));
    const U256 a = word_u256(aw), b = word_u256(bw), r = word_u256(rw);
    U256 a_red, r_true;
    {
      const u64 num[4] = {a.l[0], a.l[1], a.l[2], a.l[3]};
      u64 q[4];
      u_divrem<4>(num, m, q, &a_red);
    }
    if (is_mul) {
      u64 p[9], q[8];
      u256_mul_add(a_red, b, U256{{0, 0, 0, 0}}, p);
      const u64 num[8] = {p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]};
      u_divrem<8>(num, m, q, &r_true);
      EV_CHECK(EV_AR_MULMOD_R, u256_eq(r, r_true));  // mulmod.py:56; the constraints after it hold by construction
    } else {
      u64 num[5], q[5], c = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const u64 s1 = a_red.l[k] + b.l[k], s2 = s1 + c;
        c = (u64)(s1 < b.l[k]) | (u64)(s2 < c);
        num[k] = s2;
      }
      num[4] = c;
      u_divrem<5>(num, m, q, &r_true);
      // mul_add_words_512(d, n, r, overflow, a_reduced + b): carry_0 / carry_1 are integers iff the low / high half
      // of the pushed word is the true remainder's
      EV_CHECK(EV_AR_ADDMOD_CARRY0, r.l[0] == r_true.l[0] && r.l[1] == r_true.l[1]);
      EV_CHECK(EV_AR_ADDMOD_CARRY1, r.l[2] == r_true.l[2] && r.l[3] == r_true.l[3]);
    }
  }
  same_context_ni(s, opcode, 4, fr_u64(1), fr_u64(2));
}

ZK_HD_NOINLINE void gadget_sdiv_smod(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 p1w = zero, p2w = zero, pushw = zero;
  AR_STACK(0, 0, 0, &p1w);
  AR_STACK(1, 0, 1, &p2w);
  AR_STACK(2, 1, 1, &pushw);
  const bool is_sdiv = fr_eq_u64(opcode, 0x05);  // (SMOD - opcode) / 2 == 1 over the field
  EV_CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(p1w) && word_in_domain(p2w));
  const U256 pop1 = word_u256(p1w), pop2 = word_u256(p2w), z256{{0, 0, 0, 0}};
  const bool smod_by_zero = !is_sdiv && u256_zero(pop2);  // the pushed word is not looked at (sdiv_smod.py:106-113)
  if (!smod_by_zero) EV_CHECK(EV_AR_WITNESS_DOMAIN, word_in_domain(pushw));
  const U256 push = word_u256(pushw);
  const U256 pop1_abs = u256_abs(pop1), pop2_abs = u256_abs(pop2);
  U256 quotient = z256, remainder = pop1;
  if (is_sdiv) {  // remainder = Word(+-(|pop1| - |push| * |pop2|))
    u64 p[9];
    u256_mul_add(u256_abs(push), pop2_abs, z256, p);
    const U256 prod{{p[0], p[1], p[2], p[3]}};
    u64 br;
    const U256 x = u256_sub(pop1_abs, prod, &br);
    if ((p[4] | p[5] | p[6] | p[7] | p[8]) != 0 || br) {
      step_fail(s, u256_sign(pop1) ? EV_AR_SDIV_REM_WORD : EV_AR_SDIV_REM_NEG);
      return;
    }
    quotient = push;
    remainder = u256_sign(pop1) ? u256_neg(x) : x;
  } else if (!u256_zero(pop2)) {
    const u64 num[4] = {pop1_abs.l[0], pop1_abs.l[1], pop1_abs.l[2], pop1_abs.l[3]};
    u64 q[4];
    U256 rem;
    u_divrem<4>(num, pop2_abs, q, &rem);
    const U256 q0{{q[0], q[1], q[2], q[3]}};
    quotient = (u256_sign(pop1) == u256_sign(pop2)) ? q0 : u256_neg(q0);
    remainder = push;
  }
  // check_witness, sdiv_smod.py:33-79 (the abs_word constraints hold by construction in the halves domain)
  const U256 r_abs = u256_abs(remainder), dd_abs = pop1_abs;
  const int v = mul_add_verdict(u256_abs(quotient), pop2_abs, r_abs, dd_abs);
  EV_CHECK(EV_AR_SDIV_CARRY_LO, v != 1);
  EV_CHECK(EV_AR_SDIV_CARRY_HI, v != 2);
  EV_CHECK(EV_AR_SDIV_OVERFLOW, v == 0);
  const bool q_nz = !u256_zero(quotient), d_nz = !u256_zero(pop2), r_nz = !u256_zero(remainder);
  EV_CHECK(EV_AR_SDIV_REM_LT, !d_nz || u256_lt(r_abs, pop2_abs));
  EV_CHECK(EV_AR_SDIV_SIGN_REM, !(q_nz && d_nz && r_nz) || u256_sign(pop1) == u256_sign(remainder));
  const bool signed_overflow = u256_sign(dd_abs) != 0;
  EV_CHECK(EV_AR_SDIV_SIGN_QUOT, !(q_nz && d_nz && !signed_overflow) || (u256_sign(quotient) ^ u256_sign(pop2)) == u256_sign(pop1));
  same_context_ni(s, opcode, 3, fr_u64(1), fr_u64(1));
}

ZK_HD_NOINLINE void gadget_sar(const StepCtx& s) {
  Fr opcode = fr_u64(0);
  if (!opcode_lookup_ni(s, true, &opcode)) return;
  const Word2 zero{fr_u64(0), fr_u64(0)};
  Word2 shw = zero, aw = zero, bw = zero;
  AR_STACK(0, 0, 0, &shw);
  AR_STACK(1, 0, 1, &aw);
  AR_STACK(2, 1, 1, &bw);
  EV_CHECK(EV_AR_SAR_BYTES, word_in_domain(shw) && word_in_domain(aw) && word_in_domain(bw));
  const U256 sh = word_u256(shw), a = word_u256(aw), b = word_u256(bw);
  const unsigned shf0 = (unsigned)(sh.l[0] & 0xFF), div64 = shf0 >> 6, mod64 = shf0 & 63;
  const bool lt256 = ((sh.l[0] >> 8) | sh.l[1] | sh.l[2] | sh.l[3]) == 0;
  const int is_neg = u256_sign(a);
  const u64 fill = is_neg ? ~0ull : 0;
  // gen_witness, sar.py:180-183: an arithmetic shift right by shf0 — limbs moved by div64 (a select chain, no
  // run-time array index), then a funnel shift by mod64
  u64 m[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    u64 v = fill;
#pragma unroll
    for (int j = 0; j < 4; j++) v = (lt256 && (unsigned)(k + (int)div64) == (unsigned)j) ? a.l[j] : v;
    m[k] = v;
  }
  U256 want;
#pragma unroll
  for (int k = 0; k < 4; k++) want.l[k] = mod64 ? (m[k] >> mod64) | (m[k + 1] << (64 - mod64)) : m[k];
  EV_CHECK(EV_AR_SAR_RESULT, u256_eq(b, want));
  {  // sar.py:142-145 sign_byte_lookup, :151-152 pow2_lookup
    u32 r = 0;
    Fr k1[4] = {fr_u64(ZK_FIXED_SignByte), fr_u64(a.l[3] >> 56), fr_u64(is_neg ? 255 : 0), fr_u64(0)};
    if (!need1(s, true, lookup_sync<4>(s.t.fixed, k1, &r, s.mask, true), EV_AR_SAR_SIGN_UNSAT)) return;
    Fr k2[4] = {fr_u64(ZK_FIXED_Pow2), fr_u64(mod64), fr_u64(1ull << mod64), fr_u64(0)};
    if (!need1(s, true, lookup_sync<4>(s.t.fixed, k2, &r, s.mask, true), EV_AR_SAR_POW_LO_UNSAT)) return;
    Fr p_hi = fr_u64(mod64 ? 1ull << (64 - mod64) : 0);
    p_hi.l[1] = mod64 ? 0 : 1;
    Fr k3[4] = {fr_u64(ZK_FIXED_Pow2), fr_u64(64 - mod64), p_hi, fr_u64(0)};
    if (!need1(s, true, lookup_sync<4>(s.t.fixed, k3, &r, s.mask, true), EV_AR_SAR_POW_HI_UNSAT)) return;
  }
  same_context_ni(s, opcode, 3, fr_u64(1), fr_u64(1));
}

}  // namespace zk

// Fp2 = Fp[u]/(u^2 + 1) for BLS12-381 (G2 coordinates, the twist, line coefficients).
#pragma once
#include "fp.cuh"

namespace b200 {

struct Fp2 {
    Fp c0, c1;
};

B200_HD Fp2 fp2_zero() { Fp2 r; r.c0 = fp_zero(); r.c1 = fp_zero(); return r; }
B200_HD Fp2 fp2_one() { Fp2 r; r.c0 = fp_one(); r.c1 = fp_zero(); return r; }
B200_HD bool fp2_is_zero(const Fp2& a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
B200_HD bool fp2_eq(const Fp2& a, const Fp2& b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
B200_HD void fp2_add(Fp2& r, const Fp2& a, const Fp2& b) { fp_add(r.c0, a.c0, b.c0); fp_add(r.c1, a.c1, b.c1); }
B200_HD void fp2_sub(Fp2& r, const Fp2& a, const Fp2& b) { fp_sub(r.c0, a.c0, b.c0); fp_sub(r.c1, a.c1, b.c1); }
B200_HD void fp2_neg(Fp2& r, const Fp2& a) { fp_neg(r.c0, a.c0); fp_neg(r.c1, a.c1); }
B200_HD void fp2_dbl(Fp2& r, const Fp2& a) { fp_dbl(r.c0, a.c0); fp_dbl(r.c1, a.c1); }
B200_HD void fp2_conj(Fp2& r, const Fp2& a) { r.c0 = a.c0; fp_neg(r.c1, a.c1); }

// Karatsuba: 3 Fp products
#if defined(B200_FP2_NOINLINE)
B200_HD_NOINLINE
#else
B200_HD
#endif
void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b) {
    Fp t0, t1, s0, s1, m;
    fp_mul(t0, a.c0, b.c0);
    fp_mul(t1, a.c1, b.c1);
    fp_add(s0, a.c0, a.c1);
    fp_add(s1, b.c0, b.c1);
    fp_mul(m, s0, s1);
    fp_sub(m, m, t0);
    fp_sub(m, m, t1);
    fp_sub(r.c0, t0, t1);
    r.c1 = m;
}
// complex squaring: 2 Fp products
#if defined(B200_FP2_NOINLINE)
B200_HD_NOINLINE
#else
B200_HD
#endif
void fp2_sqr(Fp2& r, const Fp2& a) {
    Fp s, d, m;
    fp_add(s, a.c0, a.c1);
    fp_sub(d, a.c0, a.c1);
    fp_mul(m, a.c0, a.c1);
    fp_mul(r.c0, s, d);
    fp_dbl(r.c1, m);
}
B200_HD void fp2_mul_fp(Fp2& r, const Fp2& a, const Fp& k) { fp_mul(r.c0, a.c0, k); fp_mul(r.c1, a.c1, k); }
// multiply by xi = 1 + u
B200_HD void fp2_mul_xi(Fp2& r, const Fp2& a) {
    Fp t0, t1;
    fp_sub(t0, a.c0, a.c1);
    fp_add(t1, a.c0, a.c1);
    r.c0 = t0; r.c1 = t1;
}
B200_BIG void fp2_inv(Fp2& r, const Fp2& a) {
    Fp n, t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
    fp_inv(n, n);
    fp_mul(r.c0, a.c0, n);
    fp_mul(t, a.c1, n);
    fp_neg(r.c1, t);
}
B200_BIG void fp2_pow(Fp2& r, const Fp2& a, const uint32_t* e) {
    Fp2 acc = fp2_one();
    bool started = false;
#pragma unroll 1
    for (int w = 11; w >= 0; w--) {
        const uint32_t word = e[w];
#pragma unroll 1
        for (int bit = 31; bit >= 0; bit--) {
            if (started) fp2_sqr(acc, acc);
            if ((word >> bit) & 1) {
                if (started) fp2_mul(acc, acc, a); else { acc = a; started = true; }
            }
        }
    }
    r = acc;
}
// halve in Fp (Montgomery form is linear): (a + (a odd ? p : 0)) >> 1
B200_HD void fp_half(Fp& r, const Fp& a) {
    const Fp p = fp_p();
    Fp t = a;
    uint32_t carry = 0;
    if (a.l[0] & 1) carry = fp_add_raw(t, a, p);
#pragma unroll
    for (int i = 0; i < 11; i++) t.l[i] = (t.l[i] >> 1) | (t.l[i + 1] << 31);
    t.l[11] = (t.l[11] >> 1) | (carry << 31);
    r = t;
}
// For d != 0 in Fp: one exponentiation t = d^((p-3)/4) gives both candidates: x = d*t (the square root if d is a
// residue) and t = 1/x.  Returns whether d is a quadratic residue.
B200_HD bool fp_sqrt_and_inv(Fp& x, Fp& xinv, const Fp& d) {
    Fp t, chk;
    fp_pow(t, d, B200_EXP_TABLE(exp_p_minus_3_div_4));
    fp_mul(x, d, t);
    fp_sqr(chk, x);
    xinv = t;
    return fp_eq(chk, d);
}
// Square root in Fp2 by the complex method: exactly two Fp exponentiations, no data-dependent retry (threads of a
// warp stay converged).  Either root may be returned: callers fix the sign (sgn0 / flag bit).
//
// fp2_sqrt_real: a = a0 real.  a0 = s^2, or a0 = -s^2 = (s u)^2.
B200_HD bool fp2_sqrt_real(Fp2& r, const Fp& a0) {
    Fp s, c;
    fp_pow(s, a0, B200_EXP_TABLE(exp_sqrt));   // s^2 = +-a0
    fp_sqr(c, s);
    if (fp_eq(c, a0)) { r.c0 = s; r.c1 = fp_zero(); } else { r.c0 = fp_zero(); r.c1 = s; }
    return true;
}
// fp2_sqrt_with_norm_root: w with w.c1 != 0 and s = sqrt(norm(w)).  With d = (w0 + s)/2 and d' = (w0 - s)/2,
// d d' = -w1^2/4, so exactly one of them is a residue, and ONE exponentiation t = d^((p-3)/4), x = d t serves both:
//   x^2 =  d: root = (x, w1 t / 2)            (1/x = t)
//   x^2 = -d: root = (-w1 t / 2, x)           (1/x = -t; (-w1 t/2)^2 = -w1^2/(4 d) = d')
B200_BIG bool fp2_sqrt_with_norm_root(Fp2& r, const Fp2& w, const Fp& s) {
    Fp d, t, x, chk, h;
    fp_add(d, w.c0, s);
    fp_half(d, d);
    fp_pow(t, d, B200_EXP_TABLE(exp_p_minus_3_div_4));
    fp_mul(x, d, t);
    fp_sqr(chk, x);
    fp_mul(h, w.c1, t);
    fp_half(h, h);
    Fp2 cand;
    if (fp_eq(chk, d)) { cand.c0 = x; cand.c1 = h; } else { fp_neg(cand.c0, h); cand.c1 = x; }
    Fp2 sq;
    fp2_sqr(sq, cand);
    r = cand;
    return fp2_eq(sq, w);
}
// returns false if `a` is not a square
B200_BIG bool fp2_sqrt(Fp2& r, const Fp2& a) {
    if (fp_is_zero(a.c1)) return fp2_sqrt_real(r, a.c0);
    Fp n, t, s, c;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);                           // norm; a is a square in Fp2 iff its norm is one in Fp
    fp_pow(s, n, B200_EXP_TABLE(exp_sqrt));
    fp_sqr(c, s);
    if (!fp_eq(c, n)) return false;
    return fp2_sqrt_with_norm_root(r, a, s);
}
// RFC 9380 sgn0 (m = 2)
B200_HD uint32_t fp2_sgn0(const Fp2& a) {
    uint32_t s0 = fp_parity(a.c0), z0 = fp_is_zero(a.c0) ? 1u : 0u, s1 = fp_parity(a.c1);
    return s0 | (z0 & s1);
}
// ZCash "lexicographically largest" for Fp2: compare c1 first, then c0
B200_HD bool fp2_is_lex_largest(const Fp2& a) {
    return fp_is_zero(a.c1) ? fp_is_lex_largest(a.c0) : fp_is_lex_largest(a.c1);
}

}  // namespace b200

// Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v), xi = 1 + u: the pairing target field (GT) of BLS12-381.
// In terms of w: coefficient c_h.a_j sits at w^(2j+h); w^6 = xi.
#pragma once
#include "fp2.cuh"

namespace b200 {

struct Fp6 {
    Fp2 c0, c1, c2;
};
struct Fp12 {
    Fp6 c0, c1;
};

B200_HD void fp6_add(Fp6& r, const Fp6& a, const Fp6& b) { fp2_add(r.c0, a.c0, b.c0); fp2_add(r.c1, a.c1, b.c1); fp2_add(r.c2, a.c2, b.c2); }
B200_HD void fp6_sub(Fp6& r, const Fp6& a, const Fp6& b) { fp2_sub(r.c0, a.c0, b.c0); fp2_sub(r.c1, a.c1, b.c1); fp2_sub(r.c2, a.c2, b.c2); }
B200_HD void fp6_neg(Fp6& r, const Fp6& a) { fp2_neg(r.c0, a.c0); fp2_neg(r.c1, a.c1); fp2_neg(r.c2, a.c2); }
B200_HD void fp6_dbl(Fp6& r, const Fp6& a) { fp2_dbl(r.c0, a.c0); fp2_dbl(r.c1, a.c1); fp2_dbl(r.c2, a.c2); }
// multiply by v: (c0, c1, c2) -> (xi*c2, c0, c1)
B200_HD void fp6_mul_v(Fp6& r, const Fp6& a) {
    Fp2 t;
    fp2_mul_xi(t, a.c2);
    r.c2 = a.c1; r.c1 = a.c0; r.c0 = t;
}
B200_BIG void fp6_mul(Fp6& r, const Fp6& a, const Fp6& b) {
    Fp2 v0, v1, v2, t0, t1, x0, x1, x2;
    fp2_mul(v0, a.c0, b.c0);
    fp2_mul(v1, a.c1, b.c1);
    fp2_mul(v2, a.c2, b.c2);
    fp2_add(t0, a.c1, a.c2); fp2_add(t1, b.c1, b.c2);
    fp2_mul(x0, t0, t1); fp2_sub(x0, x0, v1); fp2_sub(x0, x0, v2); fp2_mul_xi(x0, x0); fp2_add(x0, x0, v0);
    fp2_add(t0, a.c0, a.c1); fp2_add(t1, b.c0, b.c1);
    fp2_mul(x1, t0, t1); fp2_sub(x1, x1, v0); fp2_sub(x1, x1, v1); fp2_mul_xi(t0, v2); fp2_add(x1, x1, t0);
    fp2_add(t0, a.c0, a.c2); fp2_add(t1, b.c0, b.c2);
    fp2_mul(x2, t0, t1); fp2_sub(x2, x2, v0); fp2_sub(x2, x2, v2); fp2_add(x2, x2, v1);
    r.c0 = x0; r.c1 = x1; r.c2 = x2;
}
// a * (b0 + b1 v)
B200_BIG void fp6_mul_by_01(Fp6& r, const Fp6& a, const Fp2& b0, const Fp2& b1) {
    Fp2 v0, v1, t0, t1, x0, x1, x2;
    fp2_mul(v0, a.c0, b0);
    fp2_mul(v1, a.c1, b1);
    fp2_add(t0, a.c1, a.c2);
    fp2_mul(x0, t0, b1); fp2_sub(x0, x0, v1); fp2_mul_xi(x0, x0); fp2_add(x0, x0, v0);
    fp2_add(t0, a.c0, a.c1); fp2_add(t1, b0, b1);
    fp2_mul(x1, t0, t1); fp2_sub(x1, x1, v0); fp2_sub(x1, x1, v1);
    fp2_add(t0, a.c0, a.c2);
    fp2_mul(x2, t0, b0); fp2_sub(x2, x2, v0); fp2_add(x2, x2, v1);
    r.c0 = x0; r.c1 = x1; r.c2 = x2;
}
// a * (b1 v)
B200_BIG void fp6_mul_by_1(Fp6& r, const Fp6& a, const Fp2& b1) {
    Fp2 x0, x1, x2;
    fp2_mul(x0, a.c2, b1); fp2_mul_xi(x0, x0);
    fp2_mul(x1, a.c0, b1);
    fp2_mul(x2, a.c1, b1);
    r.c0 = x0; r.c1 = x1; r.c2 = x2;
}
B200_BIG void fp6_inv(Fp6& r, const Fp6& a) {
    Fp2 c0, c1, c2, t, d;
    fp2_sqr(c0, a.c0); fp2_mul(t, a.c1, a.c2); fp2_mul_xi(t, t); fp2_sub(c0, c0, t);
    fp2_sqr(c1, a.c2); fp2_mul_xi(c1, c1); fp2_mul(t, a.c0, a.c1); fp2_sub(c1, c1, t);
    fp2_sqr(c2, a.c1); fp2_mul(t, a.c0, a.c2); fp2_sub(c2, c2, t);
    fp2_mul(d, a.c2, c1); fp2_mul(t, a.c1, c2); fp2_add(d, d, t); fp2_mul_xi(d, d);
    fp2_mul(t, a.c0, c0); fp2_add(d, d, t);
    fp2_inv(d, d);
    fp2_mul(r.c0, c0, d); fp2_mul(r.c1, c1, d); fp2_mul(r.c2, c2, d);
}

B200_HD Fp12 fp12_one() {
    Fp12 r;
    r.c0.c0 = fp2_one(); r.c0.c1 = fp2_zero(); r.c0.c2 = fp2_zero();
    r.c1.c0 = fp2_zero(); r.c1.c1 = fp2_zero(); r.c1.c2 = fp2_zero();
    return r;
}
B200_HD bool fp12_is_one(const Fp12& a) {
    return fp2_eq(a.c0.c0, fp2_one()) && fp2_is_zero(a.c0.c1) && fp2_is_zero(a.c0.c2) && fp2_is_zero(a.c1.c0) &&
           fp2_is_zero(a.c1.c1) && fp2_is_zero(a.c1.c2);
}
B200_HD void fp12_conj(Fp12& r, const Fp12& a) { r.c0 = a.c0; fp6_neg(r.c1, a.c1); }
B200_BIG void fp12_mul(Fp12& r, const Fp12& a, const Fp12& b) {
    Fp6 v0, v1, t0, t1, x1;
    fp6_mul(v0, a.c0, b.c0);
    fp6_mul(v1, a.c1, b.c1);
    fp6_add(t0, a.c0, a.c1);
    fp6_add(t1, b.c0, b.c1);
    fp6_mul(x1, t0, t1);
    fp6_sub(x1, x1, v0);
    fp6_sub(x1, x1, v1);
    fp6_mul_v(t0, v1);
    fp6_add(r.c0, v0, t0);
    r.c1 = x1;
}
B200_BIG void fp12_sqr(Fp12& r, const Fp12& a) {
    Fp6 ab, t0, t1, x0;
    fp6_mul(ab, a.c0, a.c1);
    fp6_add(t0, a.c0, a.c1);
    fp6_mul_v(t1, a.c1);
    fp6_add(t1, t1, a.c0);
    fp6_mul(x0, t0, t1);
    fp6_sub(x0, x0, ab);
    fp6_mul_v(t0, ab);
    fp6_sub(r.c0, x0, t0);
    fp6_dbl(r.c1, ab);
}
// f * (A + B v + C v w): the sparse line value of the M-twist Miller loop (13 Fp2 products)
B200_BIG void fp12_mul_by_line(Fp12& r, const Fp12& f, const Fp2& A, const Fp2& B, const Fp2& C) {
    Fp6 v0, v1, t0, x1;
    Fp2 bc;
    fp6_mul_by_01(v0, f.c0, A, B);
    fp6_mul_by_1(v1, f.c1, C);
    fp6_add(t0, f.c0, f.c1);
    fp2_add(bc, B, C);
    fp6_mul_by_01(x1, t0, A, bc);
    fp6_sub(x1, x1, v0);
    fp6_sub(x1, x1, v1);
    fp6_mul_v(t0, v1);
    fp6_add(r.c0, v0, t0);
    r.c1 = x1;
}
B200_BIG void fp12_inv(Fp12& r, const Fp12& a) {
    Fp6 t0, t1;
    fp6_mul(t0, a.c0, a.c0);
    fp6_mul(t1, a.c1, a.c1);
    fp6_mul_v(t1, t1);
    fp6_sub(t0, t0, t1);
    fp6_inv(t0, t0);
    fp6_mul(r.c0, a.c0, t0);
    fp6_mul(t1, a.c1, t0);
    fp6_neg(r.c1, t1);
}

// Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part of the final
// exponentiation): three Fp4 squarings = 9 Fp2 squarings instead of 12 Fp2 products.
B200_HD void fp4_sqr(Fp2& c0, Fp2& c1, const Fp2& a, const Fp2& b) {
    Fp2 t0, t1, t2;
    fp2_sqr(t0, a);
    fp2_sqr(t1, b);
    fp2_mul_xi(t2, t1);
    fp2_add(c0, t2, t0);
    fp2_add(t2, a, b);
    fp2_sqr(t2, t2);
    fp2_sub(t2, t2, t0);
    fp2_sub(c1, t2, t1);
}
B200_BIG void fp12_cyclotomic_sqr(Fp12& r, const Fp12& f) {
    Fp2 z0 = f.c0.c0, z4 = f.c0.c1, z3 = f.c0.c2, z2 = f.c1.c0, z1 = f.c1.c1, z5 = f.c1.c2;
    Fp2 t0, t1, t2, t3;
    fp4_sqr(t0, t1, z0, z1);
    fp2_sub(z0, t0, z0); fp2_dbl(z0, z0); fp2_add(z0, z0, t0);
    fp2_add(z1, t1, z1); fp2_dbl(z1, z1); fp2_add(z1, z1, t1);
    fp4_sqr(t0, t1, z2, z3);
    fp4_sqr(t2, t3, z4, z5);
    fp2_sub(z4, t0, z4); fp2_dbl(z4, z4); fp2_add(z4, z4, t0);
    fp2_add(z5, t1, z5); fp2_dbl(z5, z5); fp2_add(z5, z5, t1);
    fp2_mul_xi(t0, t3);
    fp2_add(z2, t0, z2); fp2_dbl(z2, z2); fp2_add(z2, z2, t0);
    fp2_sub(z3, t2, z3); fp2_dbl(z3, z3); fp2_add(z3, z3, t2);
    r.c0.c0 = z0; r.c0.c1 = z4; r.c0.c2 = z3;
    r.c1.c0 = z2; r.c1.c1 = z1; r.c1.c2 = z5;
}

// frobenius^k, k in {1,2,3}: coefficient at w^i -> conj^k(coefficient) * xi^(i (p^k - 1)/6)
template <int K> B200_HD Fp2 frob_gamma(int i);
#define B200_DEF_GAMMA(K)                                                             \
    template <> B200_HD Fp2 frob_gamma<K>(int i) {                                    \
        const Fp2 g1 = B200_FP2_FROB##K##_1; const Fp2 g2 = B200_FP2_FROB##K##_2;     \
        const Fp2 g3 = B200_FP2_FROB##K##_3; const Fp2 g4 = B200_FP2_FROB##K##_4;     \
        const Fp2 g5 = B200_FP2_FROB##K##_5;                                          \
        return i == 1 ? g1 : i == 2 ? g2 : i == 3 ? g3 : i == 4 ? g4 : g5;            \
    }
B200_DEF_GAMMA(1)
B200_DEF_GAMMA(2)
B200_DEF_GAMMA(3)
#undef B200_DEF_GAMMA

template <int K> B200_HD void frob_coeff(Fp2& dst, const Fp2& src, int i) {
    Fp2 t;
    if (K & 1) fp2_conj(t, src); else t = src;
    if (i == 0) { dst = t; return; }
    const Fp2 g = frob_gamma<K>(i);
    fp2_mul(dst, t, g);
}
template <int K> B200_BIG void fp12_frobenius(Fp12& r, const Fp12& a) {
    // w-power index of each tower slot: c0.c0 -> 0, c1.c0 -> 1, c0.c1 -> 2, c1.c1 -> 3, c0.c2 -> 4, c1.c2 -> 5
    frob_coeff<K>(r.c0.c0, a.c0.c0, 0);
    frob_coeff<K>(r.c1.c0, a.c1.c0, 1);
    frob_coeff<K>(r.c0.c1, a.c0.c1, 2);
    frob_coeff<K>(r.c1.c1, a.c1.c1, 3);
    frob_coeff<K>(r.c0.c2, a.c0.c2, 4);
    frob_coeff<K>(r.c1.c2, a.c1.c2, 5);
}

}  // namespace b200

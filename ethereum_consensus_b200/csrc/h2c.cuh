// hash_to_G2 for the ciphersuite BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_ (the DST at
// /root/reference/ethereum-consensus/src/crypto/bls.rs:22): RFC 9380 hash_to_curve =
// expand_message_xmd(SHA-256) -> 2 x Fp2 -> simplified SWU on the 3-isogenous curve E'' -> 3-isogeny to E' ->
// add -> clear cofactor.  blst computes the same function for every `verify`/`sign` of the reference.
#pragma once
#include "groups.cuh"
#include "sha256_hd.cuh"

namespace b200 {

B200_HD void bls_dst_prime(uint8_t out[44]) {
    const char dst[44] = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";  // 43 chars + NUL
    for (int i = 0; i < 43; i++) out[i] = uint8_t(dst[i]);
    out[43] = 43;  // I2OSP(len(DST), 1)
}

// 64 big-endian bytes -> integer mod p, Montgomery form: hi(16 B) * 2^384 + lo(48 B)
B200_HD void fp_from_be64_mod_p(Fp& r, const uint8_t* b) {
    Fp lo, hi, t;
    fp_from_be_bytes_raw(lo, b + 16);
    for (int i = 0; i < 12; i++) hi.l[i] = 0;
    for (int i = 0; i < 4; i++) {
        const uint8_t* q = b + 12 - 4 * i;
        hi.l[i] = (uint32_t(q[0]) << 24) | (uint32_t(q[1]) << 16) | (uint32_t(q[2]) << 8) | q[3];
    }
    const Fp r2 = B200_FP_R2, r3 = B200_FP_R3;
    fp_mul_portable(lo, lo, r2);   // lo * R  (lo < 2^384 may exceed p: the portable product accepts it)
    fp_mul(t, hi, r3);    // hi * R * R
    fp_add(r, lo, t);
}

// hash_to_field(msg, count = 2) over Fp2: u0 = e0 + e1 u, u1 = e2 + e3 u
B200_BIG void hash_to_field_fp2(Fp2& u0, Fp2& u1, const uint8_t* msg, size_t len) {
    uint8_t dstp[44], b0[32], bi[32], tmp[32];
    bls_dst_prime(dstp);
    Sha256Ctx c;
    sha_init(c);
    const uint8_t zero64[64] = {0};
    sha_update(c, zero64, 64);
    sha_update(c, msg, len);
    const uint8_t lib[3] = {1, 0, 0};  // I2OSP(256, 2) || I2OSP(0, 1)
    sha_update(c, lib, 3);
    sha_update(c, dstp, 44);
    sha_final(c, b0);
    uint8_t uniform[256];
    for (int i = 1; i <= 8; i++) {
        sha_init(c);
        if (i == 1) sha_update(c, b0, 32);
        else {
            for (int k = 0; k < 32; k++) tmp[k] = b0[k] ^ bi[k];
            sha_update(c, tmp, 32);
        }
        const uint8_t idx = uint8_t(i);
        sha_update(c, &idx, 1);
        sha_update(c, dstp, 44);
        sha_final(c, bi);
        for (int k = 0; k < 32; k++) uniform[32 * (i - 1) + k] = bi[k];
    }
    fp_from_be64_mod_p(u0.c0, uniform);
    fp_from_be64_mod_p(u0.c1, uniform + 64);
    fp_from_be64_mod_p(u1.c0, uniform + 128);
    fp_from_be64_mod_p(u1.c1, uniform + 192);
}

// g(x) = x^3 + A'x + B' on E''
B200_HD void sswu_g(Fp2& r, const Fp2& x) {
    const Fp2 A = B200_FP2_SSWU_A, B = B200_FP2_SSWU_B;
    Fp2 t;
    fp2_sqr(t, x);
    fp2_add(t, t, A);
    fp2_mul(t, t, x);
    fp2_add(r, t, B);
}
// simplified SWU: t -> (x, y) on E''
B200_BIG void sswu_map(Fp2& x, Fp2& y, const Fp2& t) {
    const Fp2 Zc = B200_FP2_SSWU_Z;
    Fp2 t2, zt2, tv1, x1, gx, yy;
    fp2_sqr(t2, t);
    fp2_mul(zt2, Zc, t2);          // Z t^2
    fp2_sqr(tv1, zt2);             // Z^2 t^4
    fp2_add(tv1, tv1, zt2);
    if (fp2_is_zero(tv1)) {
        const Fp2 c = B200_FP2_SSWU_B_OVER_ZA;
        x1 = c;
    } else {
        const Fp2 c = B200_FP2_SSWU_NEG_B_OVER_A;
        Fp2 inv, one = fp2_one();
        fp2_inv(inv, tv1);
        fp2_add(inv, inv, one);
        fp2_mul(x1, c, inv);
    }
    sswu_g(gx, x1);
    // y = sqrt(g(x1)) if that is a square, else x = Z t^2 x1 and y = sqrt(g(x)) = t^3 sqrt(Z^3 g(x1)).  One shared
    // exponentiation s = norm(g(x1))^((p+1)/4) decides which: s^2 = norm (square) or s^2 = -norm, and then
    // s * sqrt(-norm(Z^3)) is the root of norm(Z^3 g(x1)).  A second one finishes the complex-method root: two
    // exponentiations per map and no divergent retry.
    Fp n, tt, s, c;
    fp_sqr(n, gx.c0);
    fp_sqr(tt, gx.c1);
    fp_add(n, n, tt);
    fp_pow(s, n, B200_EXP_TABLE(exp_sqrt));
    fp_sqr(c, s);
    const bool is_sq = fp_eq(c, n);
    Fp2 w = gx;
    if (!is_sq) {
        const Fp2 z3 = B200_FP2_SSWU_Z3;
        const Fp cs = B200_FP_SSWU_SQRT_NEG_NORM_Z3;
        fp2_mul(w, z3, gx);
        fp_mul(s, s, cs);
    }
    bool ok = !fp_is_zero(w.c1) && fp2_sqrt_with_norm_root(yy, w, s);
    if (!ok) fp2_sqrt(yy, w);      // real w (probability ~2^-381): the general routine
    if (is_sq) {
        x = x1;
    } else {
        Fp2 t3;
        fp2_mul(x, zt2, x1);
        fp2_mul(t3, t2, t);
        fp2_mul(yy, yy, t3);
    }
    if (fp2_sgn0(t) != fp2_sgn0(yy)) fp2_neg(yy, yy);
    y = yy;
}

#define B200_HORNER4(r, x, c0, c1, c2, c3) \
    { const Fp2 k0 = c0, k1 = c1, k2 = c2, k3 = c3; r = k3; fp2_mul(r, r, x); fp2_add(r, r, k2); fp2_mul(r, r, x); fp2_add(r, r, k1); fp2_mul(r, r, x); fp2_add(r, r, k0); }
#define B200_HORNER3(r, x, c0, c1, c2) \
    { const Fp2 k0 = c0, k1 = c1, k2 = c2; r = k2; fp2_mul(r, r, x); fp2_add(r, r, k1); fp2_mul(r, r, x); fp2_add(r, r, k0); }

// 3-isogeny E'' -> E', output in Jacobian coordinates (no inversion): Z = xd*yd
B200_BIG void iso3_map(G2Jac& out, const Fp2& x, const Fp2& y) {
    Fp2 xn, xd, yn, yd, t, yd2;
    B200_HORNER4(xn, x, B200_FP2_ISO_XNUM0, B200_FP2_ISO_XNUM1, B200_FP2_ISO_XNUM2, B200_FP2_ISO_XNUM3);
    B200_HORNER3(xd, x, B200_FP2_ISO_XDEN0, B200_FP2_ISO_XDEN1, B200_FP2_ISO_XDEN2);
    B200_HORNER4(yn, x, B200_FP2_ISO_YNUM0, B200_FP2_ISO_YNUM1, B200_FP2_ISO_YNUM2, B200_FP2_ISO_YNUM3);
    B200_HORNER4(yd, x, B200_FP2_ISO_YDEN0, B200_FP2_ISO_YDEN1, B200_FP2_ISO_YDEN2, B200_FP2_ISO_YDEN3);
    fp2_mul(out.z, xd, yd);
    fp2_sqr(yd2, yd);
    fp2_mul(t, xn, xd);
    fp2_mul(out.x, t, yd2);          // X = xn * xd * yd^2
    fp2_sqr(t, xd);
    fp2_mul(t, t, xd);               // xd^3
    fp2_mul(t, t, yd2);
    fp2_mul(t, t, yn);
    fp2_mul(out.y, t, y);            // Y = y * yn * xd^3 * yd^2
    if (fp2_is_zero(out.z)) jac_set_inf(out);
}

// first half, for j in {0, 1}: u_j = hash_to_field(msg)[j], Q_j = iso3(sswu(u_j)) in Jacobian coordinates
B200_BIG void hash_to_g2_map(G2Jac& out, const uint8_t* msg, size_t len, int j) {
    Fp2 u0, u1, x, y;
    hash_to_field_fp2(u0, u1, msg, len);
    sswu_map(x, y, j ? u1 : u0);
    iso3_map(out, x, y);
}
// second half: Q_0 + Q_1, clear the cofactor, normalise
B200_BIG void hash_to_g2_finish(G2Aff& out, const G2Jac& q0, const G2Jac& q1) {
    G2Jac s, c;
    jac_add(s, q0, q1);
    g2_clear_cofactor(c, s);
    jac_to_aff(out, c);
}

// full hash_to_curve -> affine G2 point
B200_BIG void hash_to_g2(G2Aff& out, const uint8_t* msg, size_t len) {
    Fp2 u0, u1, x, y;
    hash_to_field_fp2(u0, u1, msg, len);
    G2Jac q0, q1;
    sswu_map(x, y, u0);
    iso3_map(q0, x, y);
    sswu_map(x, y, u1);
    iso3_map(q1, x, y);
    jac_add(q0, q0, q1);
    g2_clear_cofactor(q1, q0);
    jac_to_aff(out, q1);
}

}  // namespace b200

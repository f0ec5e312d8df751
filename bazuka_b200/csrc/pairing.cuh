// bazuka_b200 — optimal-ate pairing on BLS12-381 for the Groth16 verifier (host arithmetic over Fp / Fp2 of ff.cuh).
//
// Replaces what `zk::groth16::groth16_verify` (/root/reference/src/zk/groth16/mod.rs:67-121) gets from bellman 0.14.0
// `prepare_verifying_key` / `verify_proof` and bls12_381 0.8.0 `multi_miller_loop` / `final_exponentiation`
// (un-vendored crates).  Written from the curve's definition, not from those crates:
//
//   Fp12 = Fp2[w] / (w^6 - xi), xi = 1 + u, as six Fp2 coefficients of w^k.
//   Miller loop over |x| = 0xd201000000010000 with the G2 point T kept in JACOBIAN coordinates on the twist
//   E'(Fp2): y^2 = x^3 + 4 xi, so no step needs an inversion.  A line is the sparse element
//        l0 + (c2 * xP) w^2 + (c3 * yP) w^3
//   scaled by an Fp2 factor (2YZ^3 for a tangent, Z*H for a chord) that the final exponentiation removes:
//        tangent at T = (X,Y,Z):  l0 = 3X^3 - 2Y^2,        c2 = -3X^2 Z^2,  c3 = 2YZ^3
//        chord T,Q (Q affine):    l0 = R x_Q - y_Q Z H,    c2 = -R,         c3 = Z H      (H = x_Q Z^2 - X, R = y_Q Z^3 - Y)
//   (l0, c2, c3) depend on the G2 point only: for the verifying key's gamma, delta (and beta) they are computed once
//   and cached in the prepared key; several pairs share the squarings of one loop.
//   Final exponentiation: easy part f^((p^6-1)(p^2+1)) by conjugation, one inversion and two Frobenius maps; hard part
//   through  3 (p^4-p^2+1)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3  — five powers by |x| — so what is computed is
//   e(P,Q)^3, a fixed power of the pairing: equalities between such values are what the verifier tests.
#pragma once
#include <stdint.h>
#include <vector>
#include "ec.cuh"

namespace bzk {
namespace pairing {

struct Fp12 {
    Fp2 c[6];  // sum c[k] w^k, w^6 = xi = 1 + u
};

BZK_HD Fp2 mul_xi(const Fp2 &a) { return Fp2{a.c0 - a.c1, a.c0 + a.c1}; }
BZK_HD Fp2 mul_fp(const Fp2 &a, const Fp &s) { return Fp2{a.c0 * s, a.c1 * s}; }
BZK_HD Fp2 conj2(const Fp2 &a) { return Fp2{a.c0, a.c1.neg()}; }

BZK_HD Fp12 f12_one() {
    Fp12 r;
    for (int k = 0; k < 6; k++) r.c[k] = Fp2::zero();
    r.c[0] = Fp2::one();
    return r;
}
BZK_HD Fp12 f12_fold(const Fp2 t[11]) {
    Fp12 r;
    for (int k = 0; k < 5; k++) r.c[k] = t[k] + mul_xi(t[k + 6]);
    r.c[5] = t[5];
    return r;
}
BZK_HD Fp12 f12_mul(const Fp12 &a, const Fp12 &b) {
    Fp2 t[11];
    for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) t[i + j] = t[i + j] + a.c[i] * b.c[j];
    return f12_fold(t);
}
BZK_HD Fp12 f12_sqr(const Fp12 &a) {
    Fp2 t[11];
    for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
    for (int i = 0; i < 6; i++) {
        t[2 * i] = t[2 * i] + a.c[i].sqr();
        for (int j = i + 1; j < 6; j++) t[i + j] = t[i + j] + (a.c[i] * a.c[j]).dbl();
    }
    return f12_fold(t);
}
// a * (l0 + l2 w^2 + l3 w^3)
BZK_HD Fp12 f12_mul_sparse(const Fp12 &a, const Fp2 &l0, const Fp2 &l2, const Fp2 &l3) {
    Fp2 t[11];
    for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
    for (int i = 0; i < 6; i++) {
        t[i] = t[i] + a.c[i] * l0;
        t[i + 2] = t[i + 2] + a.c[i] * l2;
        t[i + 3] = t[i + 3] + a.c[i] * l3;
    }
    return f12_fold(t);
}
BZK_HD Fp12 f12_conj(const Fp12 &a) {  // w -> -w: the p^6 Frobenius; the inverse on the cyclotomic subgroup
    Fp12 r = a;
    r.c[1] = a.c[1].neg(); r.c[3] = a.c[3].neg(); r.c[5] = a.c[5].neg();
    return r;
}
BZK_HD bool f12_eq(const Fp12 &a, const Fp12 &b) {
    for (int k = 0; k < 6; k++) if (a.c[k] != b.c[k]) return false;
    return true;
}

// ---- Frobenius: (c_k w^k)^p = conj(c_k) * xi^(k (p-1)/6) * w^k -------------------------------------------------
struct FrobConsts { Fp2 g[6]; };
static inline const FrobConsts &frob_consts() {
    static const FrobConsts K = [] {
        static const uint32_t e[12] = {0xfffff1c7u, 0x49aa7fffu, 0x72e35555u, 0x051caaaau, 0xd3c82906u, 0xe688231au,
                                       0x7deb831fu, 0xe613e1ebu, 0xb5e1f223u, 0x0c849bf3u, 0x5eeaa66fu, 0x045582fcu};  // (p-1)/6
        const Fp2 xi{Fp::one(), Fp::one()};
        Fp2 acc = Fp2::one();
        for (int i = 12 * 32 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * xi;
        }
        FrobConsts k;
        k.g[0] = Fp2::one();
        for (int j = 1; j < 6; j++) k.g[j] = k.g[j - 1] * acc;
        return k;
    }();
    return K;
}
static inline Fp12 f12_frob(const Fp12 &a) {
    const FrobConsts &K = frob_consts();
    Fp12 r;
    r.c[0] = conj2(a.c[0]);
    for (int k = 1; k < 6; k++) r.c[k] = conj2(a.c[k]) * K.g[k];
    return r;
}

// ---- inversion: a = A + w B over Fp6 = Fp2[v]/(v^3 - xi), v = w^2;  1/a = (A - wB) / (A^2 - v B^2) ---------------
struct Fp6 { Fp2 a, b, c; };  // a + b v + c v^2
static inline Fp6 f6_mul(const Fp6 &x, const Fp6 &y) {
    return Fp6{x.a * y.a + mul_xi(x.b * y.c + x.c * y.b), x.a * y.b + x.b * y.a + mul_xi(x.c * y.c), x.a * y.c + x.b * y.b + x.c * y.a};
}
static inline Fp6 f6_sub(const Fp6 &x, const Fp6 &y) { return Fp6{x.a - y.a, x.b - y.b, x.c - y.c}; }
static inline Fp6 f6_mul_v(const Fp6 &x) { return Fp6{mul_xi(x.c), x.a, x.b}; }
static inline Fp6 f6_inv(const Fp6 &x) {
    const Fp2 t0 = x.a.sqr() - mul_xi(x.b * x.c), t1 = mul_xi(x.c.sqr()) - x.a * x.b, t2 = x.b.sqr() - x.a * x.c;
    const Fp2 n = (x.a * t0 + mul_xi(x.c * t1 + x.b * t2)).inv();
    return Fp6{t0 * n, t1 * n, t2 * n};
}
static inline Fp12 f12_inv(const Fp12 &f) {
    const Fp6 A{f.c[0], f.c[2], f.c[4]}, B{f.c[1], f.c[3], f.c[5]};
    const Fp6 d = f6_inv(f6_sub(f6_mul(A, A), f6_mul_v(f6_mul(B, B))));
    const Fp6 ra = f6_mul(A, d), rb = f6_mul(B, d);
    Fp12 r;
    r.c[0] = ra.a; r.c[2] = ra.b; r.c[4] = ra.c;
    r.c[1] = rb.a.neg(); r.c[3] = rb.b.neg(); r.c[5] = rb.c.neg();
    return r;
}

// ---- final exponentiation -----------------------------------------------------------------------------------------
#define kAbsX 0xd201000000010000ULL   /* |x|, the BLS parameter (usable in host and device code) */
static inline Fp12 pow_abs_x(const Fp12 &g) {
    Fp12 acc = g;
    for (int i = 62; i >= 0; i--) {
        acc = f12_sqr(acc);
        if ((kAbsX >> i) & 1) acc = f12_mul(acc, g);
    }
    return acc;
}
// g in the cyclotomic subgroup: g^x with x = -|x|
static inline Fp12 pow_x(const Fp12 &g) { return f12_conj(pow_abs_x(g)); }

// f^(3 (p^12 - 1)/r)
static inline Fp12 final_exp(const Fp12 &f) {
    Fp12 g = f12_mul(f12_conj(f), f12_inv(f));       // f^(p^6 - 1)
    g = f12_mul(f12_frob(f12_frob(g)), g);            // ^(p^2 + 1): now cyclotomic, inverse = conjugate
    Fp12 a = f12_mul(pow_x(g), f12_conj(g));          // g^(x-1)
    a = f12_mul(pow_x(a), f12_conj(a));               // g^((x-1)^2)
    const Fp12 b = f12_mul(pow_x(a), f12_frob(a));    // ^(x+p)
    Fp12 c = f12_mul(pow_x(pow_x(b)), f12_frob(f12_frob(b)));
    c = f12_mul(c, f12_conj(b));                      // ^(x^2 + p^2 - 1)
    return f12_mul(c, f12_mul(f12_sqr(g), g));        // * g^3
}

// ---- lines of a G2 point: everything the Miller loop needs from Q, in loop order --------------------------------
struct LineCoeff { Fp2 l0, c2, c3; };
struct G2Lines {
    bool inf = true;
    std::vector<LineCoeff> steps;  // 63 tangents interleaved with the 5 chords (68 entries)
};
static inline void compute_lines(const Affine<Fp2> &Q, G2Lines &out) {
    out.steps.clear();
    out.inf = Q.is_inf();
    if (out.inf) return;
    out.steps.reserve(68);
    Fp2 X = Q.x, Y = Q.y, Z = Fp2::one();
    for (int i = 62; i >= 0; i--) {
        {   // tangent at T, then T <- 2T   (dbl-2009-l, a = 0)
            const Fp2 A = X.sqr(), B = Y.sqr(), C = B.sqr(), ZZ = Z.sqr();
            const Fp2 t = X + B;
            const Fp2 D = (t.sqr() - A - C).dbl();
            const Fp2 E = A.dbl() + A;
            const Fp2 Z3 = (Y * Z).dbl();
            LineCoeff L;
            L.l0 = E * X - B.dbl();
            L.c2 = (E * ZZ).neg();
            L.c3 = Z3 * ZZ;
            out.steps.push_back(L);
            const Fp2 X3 = E.sqr() - D.dbl();
            const Fp2 C8 = C.dbl().dbl().dbl();
            Y = E * (D - X3) - C8;
            X = X3;
            Z = Z3;
        }
        if ((kAbsX >> i) & 1) {  // chord through T and Q, then T <- T + Q   (mixed Jacobian addition)
            const Fp2 ZZ = Z.sqr();
            const Fp2 H = Q.x * ZZ - X, Rr = Q.y * ZZ * Z - Y;
            const Fp2 Z3 = Z * H;
            LineCoeff L;
            L.l0 = Rr * Q.x - Q.y * Z3;
            L.c2 = Rr.neg();
            L.c3 = Z3;
            out.steps.push_back(L);
            const Fp2 HH = H.sqr(), HHH = HH * H, V = X * HH;
            const Fp2 X3 = Rr.sqr() - HHH - V.dbl();
            Y = Rr * (V - X3) - Y * HHH;
            X = X3;
            Z = Z3;
        }
    }
}

// Miller loop of ONE pair with the G2 point walked on the fly and the G1 point in XYZZ form (x = X/ZZ, y = Y/ZZZ): every
// line is additionally scaled by ZZ*ZZZ (an Fp factor the final exponentiation removes), so no inversion is needed to
// feed a freshly scalar-multiplied point.  Same lines as compute_lines + multi_miller; no heap: usable in a kernel.
BZK_HD Fp12 miller_one_xyzz(const Xyzz<Fp> &P, const Affine<Fp2> &Q) {
    Fp12 f = f12_one();
    if (P.is_inf() || Q.is_inf()) return f;
    const Fp sx = P.X * P.ZZZ, sy = P.Y * P.ZZ, s0 = P.ZZ * P.ZZZ;   // (xP, yP, 1) * ZZ*ZZZ
    Fp2 X = Q.x, Y = Q.y, Z = Fp2::one();
    for (int i = 62; i >= 0; i--) {
        f = f12_sqr(f);
        {
            const Fp2 A = X.sqr(), B = Y.sqr(), C = B.sqr(), ZZ = Z.sqr();
            const Fp2 t = X + B;
            const Fp2 D = (t.sqr() - A - C).dbl();
            const Fp2 E = A.dbl() + A;
            const Fp2 Z3 = (Y * Z).dbl();
            f = f12_mul_sparse(f, mul_fp(E * X - B.dbl(), s0), mul_fp((E * ZZ).neg(), sx), mul_fp(Z3 * ZZ, sy));
            const Fp2 X3 = E.sqr() - D.dbl();
            Y = E * (D - X3) - C.dbl().dbl().dbl();
            X = X3;
            Z = Z3;
        }
        if ((kAbsX >> i) & 1) {
            const Fp2 ZZ = Z.sqr();
            const Fp2 H = Q.x * ZZ - X, Rr = Q.y * ZZ * Z - Y;
            const Fp2 Z3 = Z * H;
            f = f12_mul_sparse(f, mul_fp(Rr * Q.x - Q.y * Z3, s0), mul_fp(Rr.neg(), sx), mul_fp(Z3, sy));
            const Fp2 HH = H.sqr(), HHH = HH * H, V = X * HH;
            const Fp2 X3 = Rr.sqr() - HHH - V.dbl();
            Y = Rr * (V - X3) - Y * HHH;
            X = X3;
            Z = Z3;
        }
    }
    return f;
}

struct MillerPair {
    Affine<Fp> P;
    const G2Lines *lines;
};
// prod_k miller(P_k, Q_k); pairs with an identity on either side contribute 1
static inline Fp12 multi_miller(const MillerPair *pairs, size_t n) {
    Fp12 f = f12_one();
    size_t idx = 0;
    auto apply = [&](size_t at) {
        for (size_t k = 0; k < n; k++) {
            if (pairs[k].P.is_inf() || pairs[k].lines->inf) continue;
            const LineCoeff &L = pairs[k].lines->steps[at];
            f = f12_mul_sparse(f, L.l0, mul_fp(L.c2, pairs[k].P.x), mul_fp(L.c3, pairs[k].P.y));
        }
    };
    for (int i = 62; i >= 0; i--) {
        f = f12_sqr(f);
        apply(idx++);
        if ((kAbsX >> i) & 1) apply(idx++);
    }
    return f;
}

// sum_i [k_i] P_i with shared doublings (k canonical, little-endian 32-bit limbs)
static inline Xyzz<Fp> small_msm(const Affine<Fp> *pts, const Fr *k_canon, size_t n, int bits = 255) {
    Xyzz<Fp> acc = Xyzz<Fp>::inf();
    for (int i = bits - 1; i >= 0; i--) {
        acc = acc.dbl();
        for (size_t j = 0; j < n; j++)
            if ((k_canon[j].l[i >> 5] >> (i & 31)) & 1) acc.madd(pts[j]);
    }
    return acc;
}

}  // namespace pairing
}  // namespace bzk

// bazuka_b200 — Groth16 verifier (host arithmetic over libbzk's Fp / Fp2).
//
// Replaces `zk::groth16::groth16_verify` (/root/reference/src/zk/groth16/mod.rs:67-121), i.e. bellman
// 0.14.0 `prepare_verifying_key` + `verify_proof`:  e(A,B) = e(alpha,beta) * e(sum x_i ic_i, gamma) * e(C, delta),
// with the public inputs [commitment, height, prev_state, aux_data, next_state] in that order.
// Verification is a millisecond-scale scalar job the reference runs on the CPU under the node's lock
// (/root/reference/src/node/api/post_mpn_solution.rs:12); it stays on the host here too — there is nothing
// to parallelise over in ONE verification (a batched GPU verifier for chain sync is SURVEY §8f-4).
//
// Deliberately simple pairing: ate Miller loop over |x| with affine steps on the twist E'(Fp2)
// (slope in Fp2, one Fp inversion per step), lines embedded sparsely in Fp12 = Fp2[w]/(w^6 - (u+1)) as
// c0 + c2 w^2 + c3 w^3, and the final check  prod f_i ^ ((p^12-1)/r) == 1  done inversion-free as
//   conj(f)^E == f^E ,  E = (p^2+1) (p^4-p^2+1)/r ,   since f^(p^6) = conj(f).
// About 0.6 M Fp products per verification (tens of ms on one core) — correctness first.
#include "common.cuh"

namespace bzk {
namespace {

struct Fp12 {
    Fp2 c[6];  // sum c[k] w^k, w^6 = xi = 1 + u
};

static inline Fp2 mul_xi(const Fp2 &a) { return Fp2{a.c0 - a.c1, a.c0 + a.c1}; }

static Fp12 f12_one() {
    Fp12 r;
    for (int k = 0; k < 6; k++) r.c[k] = Fp2::zero();
    r.c[0] = Fp2::one();
    return r;
}
static Fp12 f12_mul(const Fp12 &a, const Fp12 &b) {
    Fp2 t[11];
    for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) t[i + j] = t[i + j] + a.c[i] * b.c[j];
    Fp12 r;
    for (int k = 0; k < 6; k++) r.c[k] = (k + 6 < 11) ? t[k] + mul_xi(t[k + 6]) : t[k];
    return r;
}
// a * (l0 + l2 w^2 + l3 w^3)
static Fp12 f12_mul_sparse(const Fp12 &a, const Fp2 &l0, const Fp2 &l2, const Fp2 &l3) {
    Fp2 t[11];
    for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
    for (int i = 0; i < 6; i++) {
        t[i] = t[i] + a.c[i] * l0;
        t[i + 2] = t[i + 2] + a.c[i] * l2;
        t[i + 3] = t[i + 3] + a.c[i] * l3;
    }
    Fp12 r;
    for (int k = 0; k < 6; k++) r.c[k] = (k + 6 < 11) ? t[k] + mul_xi(t[k + 6]) : t[k];
    return r;
}
static Fp12 f12_conj(const Fp12 &a) {  // w -> -w  (the p^6 Frobenius)
    Fp12 r = a;
    r.c[1] = a.c[1].neg(); r.c[3] = a.c[3].neg(); r.c[5] = a.c[5].neg();
    return r;
}
static bool f12_eq(const Fp12 &a, const Fp12 &b) {
    for (int k = 0; k < 6; k++) if (a.c[k] != b.c[k]) return false;
    return true;
}
static const uint32_t kFinalExp[64] = {
    0xc0705d6au, 0x8739e1cdu, 0xe0381a16u, 0x09a5256du, 0x61c791e2u, 0x9cf0f70au, 0x7903f76eu, 0x3a09c449u,
    0x3890f133u, 0x2d727156u, 0x6fec7760u, 0x224741b3u, 0x2a12bd40u, 0x338259c2u, 0x778e0de7u, 0x38ee1cd4u,
    0x188a20b0u, 0xc3b5ef4bu, 0xe2764d7bu, 0x1d615d49u, 0xd076117du, 0x816101ddu, 0x7ebe3afcu, 0xf007c01eu,
    0x935021c3u, 0x27d7bd90u, 0x57c0b15fu, 0xc3b5e2f5u, 0xc4f82384u, 0x5e886c94u, 0x11e63f56u, 0xee6a95dbu,
    0x4a9c4f6fu, 0x2b822f51u, 0xd21b73dau, 0x12d6a874u, 0xf499dffbu, 0x1304275eu, 0xbcb95d1fu, 0x967878feu,
    0x8b2f2922u, 0x4744497fu, 0xf0841855u, 0x85a2e707u, 0x6c802eecu, 0x9f0c5012u, 0xbd2fa489u, 0xfb46e197u,
    0x9bc5f61au, 0x548ce080u, 0x73beaa8cu, 0xcf56fb15u, 0x763bdf7cu, 0xad7375a3u, 0x179bdeccu, 0xe0ec9031u,
    0x3c48c1dau, 0x6579aea8u, 0x64cf5bb3u, 0xdbf85ae6u, 0x55ca7566u, 0x7b6f235cu, 0x14877503u, 0x000028b3u};
static Fp12 f12_pow_final(const Fp12 &a) {
    Fp12 acc = f12_one();
    bool started = false;
    for (int i = 64 * 32 - 1; i >= 0; i--) {
        if (started) acc = f12_mul(acc, acc);
        if ((kFinalExp[i >> 5] >> (i & 31)) & 1) {
            acc = started ? f12_mul(acc, a) : a;
            started = true;
        }
    }
    return acc;
}

// f *= miller(Q, P) for P in G1 (affine Fp), Q in G2 (affine on the twist); identity inputs contribute 1
static void miller_accumulate(Fp12 &f_total, const G1Affine &P, const G2Affine &Q) {
    if (P.is_inf() || Q.is_inf()) return;
    static const uint64_t X = 0xd201000000010000ULL;  // |x|
    Fp12 f = f12_one();
    Fp2 tx = Q.x, ty = Q.y;
    const Fp2 xp{P.x, Fp::zero()}, yp{P.y, Fp::zero()};
    auto line = [&](const Fp2 &lam) {
        // (lam*tx - ty) + (-lam*xP) w^2 + yP w^3   [the line scaled by w^3, which the final exponent kills]
        f = f12_mul_sparse(f, lam * tx - ty, (lam * xp).neg(), yp);
    };
    for (int i = 62; i >= 0; i--) {
        f = f12_mul(f, f);
        // tangent at T
        Fp2 x2 = tx.sqr();
        Fp2 lam = (x2.dbl() + x2) * ty.dbl().inv();
        line(lam);
        Fp2 nx = lam.sqr() - tx.dbl();
        ty = lam * (tx - nx) - ty;
        tx = nx;
        if ((X >> i) & 1) {
            Fp2 lam2 = (Q.y - ty) * (Q.x - tx).inv();
            line(lam2);
            Fp2 ax = lam2.sqr() - tx - Q.x;
            ty = lam2 * (tx - ax) - ty;
            tx = ax;
        }
    }
    f_total = f12_mul(f_total, f);
}

static G1Affine img_g1(const bzk_g1_affine *img) {
    if (img->infinity) return G1Affine::inf();
    G1Affine p;
    memcpy(p.x.l, img->x, 48);
    memcpy(p.y.l, img->y, 48);
    return p;
}
static G2Affine img_g2(const bzk_g2_affine *img) {
    if (img->infinity) return G2Affine::inf();
    G2Affine p;
    memcpy(p.x.c0.l, img->x, 48); memcpy(p.x.c1.l, img->x + 6, 48);
    memcpy(p.y.c0.l, img->y, 48); memcpy(p.y.c1.l, img->y + 6, 48);
    return p;
}
static bool on_curve_g1(const G1Affine &p) { return p.is_inf() || p.y.sqr() == p.x.sqr() * p.x + Fp::from_u32(4); }
static bool on_curve_g2(const G2Affine &p) {
    Fp four = Fp::from_u32(4);
    return p.is_inf() || p.y.sqr() == p.x.sqr() * p.x + Fp2{four, four};
}

}  // namespace
}  // namespace bzk

using namespace bzk;

extern "C" {

/* 1 = accepted, 0 = rejected (including malformed points), <0 = BZK_ERR_BAD_ARG */
int32_t bzk_groth16_verify(const bzk_g1_affine *alpha_g1, const bzk_g2_affine *beta_g2, const bzk_g2_affine *gamma_g2,
                           const bzk_g2_affine *delta_g2, const bzk_g1_affine *ic, size_t n_ic,
                           const bzk_fr *public_inputs, size_t n_inputs,
                           const bzk_g1_affine *proof_a, const bzk_g2_affine *proof_b, const bzk_g1_affine *proof_c) {
    if (!alpha_g1 || !beta_g2 || !gamma_g2 || !delta_g2 || !ic || !proof_a || !proof_b || !proof_c) return BZK_ERR_BAD_ARG;
    if (n_ic != n_inputs + 1 || (n_inputs && !public_inputs)) return BZK_ERR_BAD_ARG;
    const G1Affine A = img_g1(proof_a), C = img_g1(proof_c), al = img_g1(alpha_g1);
    const G2Affine B = img_g2(proof_b), be = img_g2(beta_g2), ga = img_g2(gamma_g2), de = img_g2(delta_g2);
    if (!on_curve_g1(A) || !on_curve_g1(C) || !on_curve_g2(B)) return 0;
    G1Xyzz acc = G1Xyzz::from_affine(img_g1(&ic[0]));
    for (size_t i = 0; i < n_inputs; i++) {
        Fr x;
        memcpy(x.l, &public_inputs[i], 32);
        const Fr k = x.from_mont();
        acc.add(scalar_mul(img_g1(&ic[i + 1]), k.l));
    }
    Fp12 f = f12_one();
    miller_accumulate(f, A, B);
    miller_accumulate(f, acc.to_affine().neg(), ga);
    miller_accumulate(f, C.neg(), de);
    miller_accumulate(f, al.neg(), be);
    return f12_eq(f12_pow_final(f12_conj(f)), f12_pow_final(f)) ? 1 : 0;
}

/* `check_proof` on the reference's byte images (/root/reference/src/zk/mod.rs:157-193): vk = bincode
 * `Groth16VerifyingKey` (878 + 97*len bytes, /root/reference/src/zk/groth16/mod.rs:22-31), proof = 387-byte
 * `Groth16Proof`; inputs = Montgomery scalars. */
int32_t bzk_groth16_verify_bytes(const uint8_t *vk, size_t vk_len, const bzk_fr *public_inputs, size_t n_inputs, const uint8_t *proof387) {
    if (!vk || !proof387 || vk_len < 878) return BZK_ERR_BAD_ARG;
    auto g1_at = [](const uint8_t *p) { bzk_g1_affine g; memset(&g, 0, sizeof g); memcpy(&g, p, 97); return g; };
    auto g2_at = [](const uint8_t *p) { bzk_g2_affine g; memset(&g, 0, sizeof g); memcpy(&g, p, 193); return g; };
    size_t off = 0;
    bzk_g1_affine alpha = g1_at(vk + off); off += 97;
    off += 97;  // beta_g1 (not used by the verifier)
    bzk_g2_affine beta = g2_at(vk + off); off += 193;
    bzk_g2_affine gamma = g2_at(vk + off); off += 193;
    off += 97;  // delta_g1
    bzk_g2_affine delta = g2_at(vk + off); off += 193;
    uint64_t n_ic;
    memcpy(&n_ic, vk + off, 8); off += 8;
    if (n_ic > 4096 || vk_len != off + 97 * n_ic) return BZK_ERR_BAD_ARG;
    std::vector<bzk_g1_affine> ic(n_ic);
    for (uint64_t i = 0; i < n_ic; i++) ic[i] = g1_at(vk + off + 97 * i);
    bzk_g1_affine a = g1_at(proof387), c = g1_at(proof387 + 290);
    bzk_g2_affine b = g2_at(proof387 + 97);
    return bzk_groth16_verify(&alpha, &beta, &gamma, &delta, ic.data(), (size_t)n_ic, public_inputs, n_inputs, &a, &b, &c);
}

}  // extern "C"

// Host-side shim over bazuka_b200/csrc/ff.cuh + ec.cuh for the CPU ("not gpu") test tier:
// runs the DEVICE multiplication algorithm (Fe::mul_evenodd, explicit-carry build) and the host
// group law on the CPU so their logic is checked against the oracle without a GPU.
#include <cstring>
#include "ec.cuh"
#include "witness_core.cuh"
#include <vector>
using namespace bzk;
template <class T, int W> static void bin(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n, int op) {
    for (size_t i = 0; i < n; i++) {
        T x, y, z;
        memcpy(x.l, a + W * i, 4 * W); memcpy(y.l, b + W * i, 4 * W);
        switch (op) {
            case 0: z = T::add_limbs32(x, y); break;   // the device text (explicit-carry build)
            case 1: z = T::sub_limbs32(x, y); break;
            case 2: z = T::mul_evenodd(x, y); break;
            case 4: z = x + y; break;                  // host fast paths (64-bit limbs)
            case 5: z = x - y; break;
            default: z = x * y; break;
        }
        memcpy(r + W * i, z.l, 4 * W);
    }
}
extern "C" {
// lazily reduced inner product (Fe::mul_wide / wide_accumulate / redc_wide): out = sum_k m[k]*s[k],
// m given UNSCALED in Montgomery form (the shim applies the 2^32 pre-scaling like the loader does)
void shim_fr_dot(const uint32_t *m, const uint32_t *s, size_t t, uint32_t *out) {
    uint32_t acc[17] = {0};
    Fr two32 = Fr::zero(); two32.l[1] = 1; two32 = two32.to_mont();
    for (size_t k = 0; k < t; k++) {
        Fr a, b; memcpy(a.l, m + 8 * k, 32); memcpy(b.l, s + 8 * k, 32);
        uint32_t w[16];
        Fr::mul_wide(w, a * two32, b);
        Fr::wide_accumulate(acc, w);
    }
    Fr r = Fr::redc_wide(acc); memcpy(out, r.l, 32);
}
void shim_fr(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n, int op) { bin<Fr, 8>(a, b, r, n, op); }
void shim_fp(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n, int op) { bin<Fp, 12>(a, b, r, n, op); }
void shim_fr_inv(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) { Fr x; memcpy(x.l, a + 8 * i, 32); Fr z = x.inv(); memcpy(r + 8 * i, z.l, 32); }
}
void shim_fr_inv_gcd(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) { Fr x; memcpy(x.l, a + 8 * i, 32); Fr z = x.inv_gcd(); memcpy(r + 8 * i, z.l, 32); }
}
void shim_fp_inv_gcd(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) { Fp x; memcpy(x.l, a + 12 * i, 48); Fp z = x.inv_gcd(); memcpy(r + 12 * i, z.l, 48); }
}
void shim_fp_inv(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) { Fp x; memcpy(x.l, a + 12 * i, 48); Fp z = x.inv(); memcpy(r + 12 * i, z.l, 48); }
}
// packed affine (x|y, 96 B): out = [k] p via XYZZ double-and-add, k canonical 8x u32
void shim_g1_mul(const uint32_t *p, const uint32_t *k, uint32_t *out) {
    G1Affine a; memcpy(&a, p, 96);
    G1Affine r = scalar_mul(a, k).to_affine();
    memcpy(out, &r, 96);
}
void shim_g2_mul(const uint32_t *p, const uint32_t *k, uint32_t *out) {
    G2Affine a; memcpy(&a, p, 192);
    G2Affine r = scalar_mul(a, k).to_affine();
    memcpy(out, &r, 192);
}
// acc (XYZZ from affine a) + b three ways: madd, add, and P+P / P-P exceptional cases
void shim_g1_add(const uint32_t *pa, const uint32_t *pb, uint32_t *out_madd, uint32_t *out_add) {
    G1Affine a, b; memcpy(&a, pa, 96); memcpy(&b, pb, 96);
    G1Xyzz x = G1Xyzz::from_affine(a); x = x.dbl(); x.madd(a.neg());  // 2a - a = a, non-trivial ZZ
    G1Xyzz y = x; y.madd(b);
    G1Affine r1 = y.to_affine(); memcpy(out_madd, &r1, 96);
    G1Xyzz z = G1Xyzz::from_affine(b); z = z.dbl(); z.madd(b.neg());
    G1Xyzz w = x; w.add(z);
    G1Affine r2 = w.to_affine(); memcpy(out_add, &r2, 96);
}

// the witness interpreter's per-slot loop (witness_core.cuh, the code the device kernel runs) on the host:
// one slot, values in a plain array; aux_out[n_ops] receives the block's variables (Montgomery)
void shim_witness_run(const int32_t *ops, uint32_t n_ops, const int32_t *lc_ptr, const int32_t *lc_slot, const int32_t *lc_coef,
                      const uint32_t *coefs, uint32_t n_raw, uint32_t n_ext, const uint32_t *jj_d, const uint32_t *raws,
                      const uint32_t *ext, uint32_t *aux_out) {
    struct Mem {
        std::vector<Fr> V;
        Fr *out_;
        Fr load(int32_t slot) const { return V[slot]; }
        void store(uint32_t slot, const Fr &v) { V[slot] = v; }
        void out(uint32_t j, const Fr &v) { out_[j] = v; }
        void prefetch(int32_t) const {}
    } mem;
    mem.V.assign(1 + n_ext + n_ops, Fr::zero());
    mem.out_ = (Fr *)aux_out;
    WitProgDev P{ops, lc_ptr, lc_slot, lc_coef, (const Fr *)coefs, n_ops, n_raw, n_ext};
    Fr d; memcpy(d.l, jj_d, 32);
    wit_run_slot(P, d, (const Fr *)raws, (const Fr *)ext, mem);
}

// the same program in the order the GPU kernel runs it (witness.cu k_witness_levels): the schedule of wit_build_schedule, level
// by level, the ops of a level BACKWARDS (any order inside a level must do); returns the level count, stats[0..3] =
// executed ops, widest level, levels holding an inversion (JJ / INVZ)
uint32_t shim_witness_run_levels(const int32_t *ops, uint32_t n_ops, const int32_t *lc_ptr, const int32_t *lc_slot, const int32_t *lc_coef,
                                 const uint32_t *coefs, uint32_t n_raw, uint32_t n_ext, const uint32_t *jj_d, const uint32_t *raws,
                                 const uint32_t *ext, uint32_t *aux_out, uint64_t *stats) {
    struct Mem {   // the device layout: {ONE, externals} in a side array, block variables in the z segment
        std::vector<Fr> pre;
        Fr *z;
        uint32_t block0;
        Fr load(int32_t slot) const { return (uint32_t)slot < block0 ? pre[slot] : z[slot - block0]; }
        void store(uint32_t slot, const Fr &v) { z[slot - block0] = v; }
        void out(uint32_t, const Fr &) {}
        void prefetch(int32_t) const {}
    } mem;
    mem.block0 = 1 + n_ext;
    mem.pre.assign(mem.block0, Fr::one());
    for (uint32_t k = 0; k < n_ext; k++) { Fr e; memcpy(e.l, ext + 8 * k, 32); mem.pre[1 + k] = e.to_mont(); }
    mem.z = (Fr *)aux_out;
    WitProgDev P{ops, lc_ptr, lc_slot, lc_coef, (const Fr *)coefs, n_ops, n_raw, n_ext};
    Fr d; memcpy(d.l, jj_d, 32);
    std::vector<int32_t> sops, level_ptr;
    const uint32_t n_levels = wit_build_schedule(ops, n_ops, lc_ptr, lc_slot, n_ext, sops, level_ptr);
    uint64_t widest = 0, heavy = 0;
    for (uint32_t L = 0; L < n_levels; L++) {
        bool inv = false;
        for (int32_t i = level_ptr[L + 1] - 1; i >= level_ptr[L]; i--) {
            const int32_t *o = sops.data() + (size_t)i * 8;
            inv |= (o[0] == W_JJ || o[0] == W_INVZ);
            wit_exec_op(P, d, (const Fr *)raws, (uint32_t)o[6], o[0], o[1], o[2], o[3], o[4], o[5], mem);
        }
        widest = std::max<uint64_t>(widest, (uint64_t)(level_ptr[L + 1] - level_ptr[L]));
        heavy += inv;
    }
    if (stats) { stats[0] = (uint64_t)level_ptr[n_levels]; stats[1] = widest; stats[2] = heavy; }
    return n_levels;
}
}

// ---- csrc/pairing.cuh on the host: e(P,Q)^3 as 6 Fp2 coefficients (12 x 12 u32, Montgomery), and its pieces
#include "pairing.cuh"
extern "C" {
static G1Affine shim_g1(const uint32_t *p) { G1Affine a; memcpy(a.x.l, p, 48); memcpy(a.y.l, p + 12, 48); return a; }
static G2Affine shim_g2(const uint32_t *p) {
    G2Affine a;
    memcpy(a.x.c0.l, p, 48); memcpy(a.x.c1.l, p + 12, 48); memcpy(a.y.c0.l, p + 24, 48); memcpy(a.y.c1.l, p + 36, 48);
    return a;
}
static void shim_f12_out(const pairing::Fp12 &f, uint32_t *out) {
    for (int k = 0; k < 6; k++) { memcpy(out + 24 * k, f.c[k].c0.l, 48); memcpy(out + 24 * k + 12, f.c[k].c1.l, 48); }
}
static pairing::Fp12 shim_f12_in(const uint32_t *in) {
    pairing::Fp12 f;
    for (int k = 0; k < 6; k++) { memcpy(f.c[k].c0.l, in + 24 * k, 48); memcpy(f.c[k].c1.l, in + 24 * k + 12, 48); }
    return f;
}
// g1: x|y (24 u32), g2: x.c0|x.c1|y.c0|y.c1 (48 u32), Montgomery; out: 144 u32
void shim_pairing_cubed(const uint32_t *g1, const uint32_t *g2, uint32_t *out) {
    pairing::G2Lines L;
    pairing::compute_lines(shim_g2(g2), L);
    pairing::MillerPair p{shim_g1(g1), &L};
    shim_f12_out(pairing::final_exp(pairing::multi_miller(&p, 1)), out);
}
void shim_miller(const uint32_t *g1, const uint32_t *g2, uint32_t *out) {
    pairing::G2Lines L;
    pairing::compute_lines(shim_g2(g2), L);
    pairing::MillerPair p{shim_g1(g1), &L};
    shim_f12_out(pairing::multi_miller(&p, 1), out);
}
void shim_final_exp(const uint32_t *in, uint32_t *out) { shim_f12_out(pairing::final_exp(shim_f12_in(in)), out); }
void shim_f12_inv_frob(const uint32_t *in, uint32_t *out_inv, uint32_t *out_frob) {
    shim_f12_out(pairing::f12_inv(shim_f12_in(in)), out_inv);
    shim_f12_out(pairing::f12_frob(shim_f12_in(in)), out_frob);
}
}

/* ORACLE — test infrastructure only.
 *
 * CPU restatement (plain C, 64-bit limbs, pthreads) of the arithmetic on Bazuka's MPN Groth16
 * proving path.  It is the checker for the CUDA library and the timed "bellman-equivalent
 * restatement" CPU baseline; it is never linked into, loaded by, or called from the product
 * (bazuka_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.
 *
 * PARITY STATUS: Poseidon is pinned by the reference's 16 known answers
 * (/root/reference/src/zk/poseidon/mod.rs:115-149) and the empty-MPN-root constant
 * (/root/reference/src/node/api/get_explorer_blocks.rs:29).  Fp/G1/G2 wire layout is pinned by
 * the three production verifying keys (/root/reference/src/config/blockchain.rs:32-37).  NTT
 * outputs, MSM results and proof bytes are **parity unpinned** by the reference (its tests use
 * OsRng and assert only verify_proof().is_ok()); they are pinned here against the big-integer
 * model in oracle/py and by algebraic identities.
 *
 * The arithmetic itself lives in crates the reference does not vendor — bellman 0.14.0,
 * bls12_381 0.8.0, ff 0.13 (/root/reference/Cargo.toml:19,27-28) — so each section below
 * restates the published algorithm and cites the reference call site it serves.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ fields */
#define F fr
#define NL 4
#define F_MODULUS { 0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL }
#include "mont_tmpl.h"
#undef F
#undef NL
#undef F_MODULUS

#define F fp
#define NL 6
#define F_MODULUS { 0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, \
                    0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL }
#include "mont_tmpl.h"
#undef F
#undef NL
#undef F_MODULUS

/* Fp2 = Fp[u]/(u^2+1), element = c0 | c1 (6+6 limbs) — bls12_381 0.8.0 `Fp2 {c0,c1}` image,
 * the order the reference's `((Fp,Fp),(Fp,Fp),bool)` G2 tuple exposes
 * (/root/reference/src/zk/groth16/mod.rs:25-27). */
static inline void fp2_add(u64 *r, const u64 *a, const u64 *b) { fp_add(r, a, b); fp_add(r + 6, a + 6, b + 6); }
static inline void fp2_sub(u64 *r, const u64 *a, const u64 *b) { fp_sub(r, a, b); fp_sub(r + 6, a + 6, b + 6); }
static inline void fp2_dbl(u64 *r, const u64 *a) { fp_dbl(r, a); fp_dbl(r + 6, a + 6); }
static inline void fp2_neg(u64 *r, const u64 *a) { fp_neg(r, a); fp_neg(r + 6, a + 6); }
static inline void fp2_copy(u64 *r, const u64 *a) { memcpy(r, a, 96); }
static inline int fp2_is_zero(const u64 *a) { return fp_is_zero(a) && fp_is_zero(a + 6); }
static inline int fp2_eq(const u64 *a, const u64 *b) { return fp_eq(a, b) && fp_eq(a + 6, b + 6); }
static inline void fp2_one(u64 *r) { fp_copy(r, fp_R1); fp_zero(r + 6); }
static inline void fp2_mul(u64 *r, const u64 *a, const u64 *b) {
    u64 t0[6], t1[6], t2[6], t3[6];
    fp_mul(t0, a, b); fp_mul(t1, a + 6, b + 6);
    fp_mul(t2, a, b + 6); fp_mul(t3, a + 6, b);
    fp_sub(r, t0, t1); fp_add(r + 6, t2, t3);
}
static inline void fp2_sqr(u64 *r, const u64 *a) { fp2_mul(r, a, a); }
static inline void fp2_inv(u64 *r, const u64 *a) {
    u64 n[6], t[6];
    fp_sqr(n, a); fp_sqr(t, a + 6); fp_add(n, n, t); fp_inv(n, n);
    fp_mul(r, a, n); fp_mul(t, a + 6, n); fp_neg(r + 6, t);
}
static inline void fp_one(u64 *r) { fp_copy(r, fp_R1); }

/* ------------------------------------------------------------------ groups */
static unsigned bellman_window_c(size_t n) {
    /* bellman 0.14.0 multiexp: c = 3 if n < 32 else ceil(ln n) */
    if (n < 32) return 3;
    return (unsigned)ceil(log((double)n));
}

#define E g1
#define FW 6
#define AFF_BYTES 104
#define FE_add fp_add
#define FE_sub fp_sub
#define FE_mul fp_mul
#define FE_sqr fp_sqr
#define FE_dbl fp_dbl
#define FE_neg fp_neg
#define FE_inv fp_inv
#define FE_is_zero fp_is_zero
#define FE_eq fp_eq
#define FE_copy fp_copy
#define FE_one fp_one
#include "ec_tmpl.h"
#undef E
#undef FW
#undef AFF_BYTES
#undef FE_add
#undef FE_sub
#undef FE_mul
#undef FE_sqr
#undef FE_dbl
#undef FE_neg
#undef FE_inv
#undef FE_is_zero
#undef FE_eq
#undef FE_copy
#undef FE_one

#define E g2
#define FW 12
#define AFF_BYTES 200
#define FE_add fp2_add
#define FE_sub fp2_sub
#define FE_mul fp2_mul
#define FE_sqr fp2_sqr
#define FE_dbl fp2_dbl
#define FE_neg fp2_neg
#define FE_inv fp2_inv
#define FE_is_zero fp2_is_zero
#define FE_eq fp2_eq
#define FE_copy fp2_copy
#define FE_one fp2_one
#include "ec_tmpl.h"
#undef E

/* ------------------------------------------------------------------ init */
static int g_init_done = 0;
static uint8_t G1_GEN_IMG[104], G2_GEN_IMG[200];

static void hex_to_limbs(u64 *out, int nl, const char *hex) {
    /* big-endian hex string (exactly nl*16 digits) -> little-endian limbs */
    for (int i = 0; i < nl; i++) {
        u64 v = 0;
        for (int k = 0; k < 16; k++) {
            char ch = hex[(nl - 1 - i) * 16 + k];
            v = (v << 4) | (u64)(ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10);
        }
        out[i] = v;
    }
}

void bzko_init(void) {
    if (g_init_done) return;
    fr_init(); fp_init();
    u64 t[6];
    memset(G1_GEN_IMG, 0, sizeof G1_GEN_IMG); memset(G2_GEN_IMG, 0, sizeof G2_GEN_IMG);
    static const char *g1h[2] = {
        "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
        "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1" };
    static const char *g2h[4] = {
        "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8",
        "13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e",
        "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801",
        "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be" };
    for (int i = 0; i < 2; i++) { hex_to_limbs(t, 6, g1h[i]); fp_to_mont(t, t); memcpy(G1_GEN_IMG + 48 * i, t, 48); }
    for (int i = 0; i < 4; i++) { hex_to_limbs(t, 6, g2h[i]); fp_to_mont(t, t); memcpy(G2_GEN_IMG + 48 * i, t, 48); }
    g_init_done = 1;
}

/* ------------------------------------------------------------------ exported: field ops */
/* all Fr/Fp arguments are Montgomery images unless a name says "canon" */
void bzko_fr_mul(const u64 *a, const u64 *b, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fr_mul(r + 4 * i, a + 4 * i, b + 4 * i); }
void bzko_fr_add(const u64 *a, const u64 *b, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fr_add(r + 4 * i, a + 4 * i, b + 4 * i); }
void bzko_fr_sub(const u64 *a, const u64 *b, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fr_sub(r + 4 * i, a + 4 * i, b + 4 * i); }
void bzko_fr_inv(const u64 *a, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fr_inv(r + 4 * i, a + 4 * i); }
void bzko_fr_to_mont(const u64 *a, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fr_to_mont(r + 4 * i, a + 4 * i); }
void bzko_fr_from_mont(const u64 *a, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fr_from_mont(r + 4 * i, a + 4 * i); }
void bzko_fp_mul(const u64 *a, const u64 *b, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fp_mul(r + 6 * i, a + 6 * i, b + 6 * i); }
void bzko_fp_add(const u64 *a, const u64 *b, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fp_add(r + 6 * i, a + 6 * i, b + 6 * i); }
void bzko_fp_sub(const u64 *a, const u64 *b, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fp_sub(r + 6 * i, a + 6 * i, b + 6 * i); }
void bzko_fp_inv(const u64 *a, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fp_inv(r + 6 * i, a + 6 * i); }
void bzko_fp_to_mont(const u64 *a, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fp_to_mont(r + 6 * i, a + 6 * i); }
void bzko_fp_from_mont(const u64 *a, u64 *r, size_t n) { bzko_init(); for (size_t i = 0; i < n; i++) fp_from_mont(r + 6 * i, a + 6 * i); }

/* SplitMix64 -> uniform Fr (Montgomery image), the generator of SURVEY.md §8(d) configs 2/3 */
static inline u64 splitmix_next(u64 *s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* canonical value = (256-bit LE integer from 4 draws) mod r, emitted in Montgomery form */
void bzko_fr_random(u64 seed, u64 *out, size_t n) {
    bzko_init();
    u64 s = seed;
    for (size_t i = 0; i < n; i++) {
        u64 v[4];
        for (int k = 0; k < 4; k++) v[k] = splitmix_next(&s);
        /* v < 2^256 < 5r : subtract r until canonical */
        while (fr_geq_p(v)) fr_sub_p(v, v);
        fr_to_mont(out + 4 * i, v);
    }
}

/* ------------------------------------------------------------------ exported: curve ops */
void bzko_g1_generator(uint8_t *img) { bzko_init(); memcpy(img, G1_GEN_IMG, 104); }
void bzko_g2_generator(uint8_t *img) { bzko_init(); memcpy(img, G2_GEN_IMG, 200); }

int bzko_g1_on_curve(const uint8_t *img) {
    bzko_init();
    if (img[96]) return 1;
    u64 x[6], y[6], l[6], r[6], b[6];
    memcpy(x, img, 48); memcpy(y, img + 48, 48);
    fp_sqr(l, y); fp_sqr(r, x); fp_mul(r, r, x); fp_set_u64(b, 4); fp_add(r, r, b);
    return fp_eq(l, r);
}
int bzko_g2_on_curve(const uint8_t *img) {
    bzko_init();
    if (img[192]) return 1;
    u64 x[12], y[12], l[12], r[12], b[12];
    memcpy(x, img, 96); memcpy(y, img + 96, 96);
    fp2_sqr(l, y); fp2_sqr(r, x); fp2_mul(r, r, x);
    fp_set_u64(b, 4); fp_set_u64(b + 6, 4); fp2_add(r, r, b);
    return fp2_eq(l, r);
}
/* out = a + b (affine images) */
void bzko_g1_add(const uint8_t *a, const uint8_t *b, uint8_t *out) {
    bzko_init(); g1_jac p, q; g1_from_affine(&p, a); g1_from_affine(&q, b); g1_add(&p, &p, &q); g1_to_affine(out, &p);
}
void bzko_g2_add(const uint8_t *a, const uint8_t *b, uint8_t *out) {
    bzko_init(); g2_jac p, q; g2_from_affine(&p, a); g2_from_affine(&q, b); g2_add(&p, &p, &q); g2_to_affine(out, &p);
}
/* out = [k] a, k = Montgomery Fr image */
void bzko_g1_mul(const uint8_t *a, const u64 *k_mont, uint8_t *out) {
    bzko_init(); u64 k[4]; fr_from_mont(k, k_mont);
    g1_jac p; g1_from_affine(&p, a); g1_mul_u256(&p, &p, k); g1_to_affine(out, &p);
}
void bzko_g2_mul(const uint8_t *a, const u64 *k_mont, uint8_t *out) {
    bzko_init(); u64 k[4]; fr_from_mont(k, k_mont);
    g2_jac p; g2_from_affine(&p, a); g2_mul_u256(&p, &p, k); g2_to_affine(out, &p);
}

/* n pseudo-random bases  P_i = [k_i] G,  k_i = SplitMix64(seed) Fr draws (same stream rule as
 * bzko_fr_random).  Threaded; fixed-base 8-bit window table. */
typedef struct { u64 seed; size_t lo, hi; uint8_t *out; const void *table; int g2; } gen_job;
static void *g1_gen_thread(void *arg_) {
    gen_job *job = arg_;
    const uint8_t (*table)[255][104] = job->table;
    size_t cnt = job->hi - job->lo;
    g1_jac *pts = malloc(sizeof(g1_jac) * cnt);
    u64 s = job->seed;
    for (size_t i = 0; i < job->lo * 4; i++) splitmix_next(&s);
    for (size_t i = 0; i < cnt; i++) {
        u64 v[4];
        for (int k = 0; k < 4; k++) v[k] = splitmix_next(&s);
        while (fr_geq_p(v)) fr_sub_p(v, v);
        g1_jac acc; g1_set_inf(&acc);
        for (int w = 0; w < 32; w++) {
            unsigned d = (v[w / 8] >> ((w % 8) * 8)) & 0xff;
            if (d) g1_madd(&acc, &acc, table[w][d - 1]);
        }
        pts[i] = acc;
    }
    g1_batch_to_affine(job->out + job->lo * 104, pts, cnt);
    free(pts);
    return NULL;
}
void bzko_g1_random_bases(u64 seed, uint8_t *out, size_t n, int threads) {
    bzko_init();
    /* table[w][d-1] = [d * 256^w] G */
    uint8_t (*table)[255][104] = malloc(32 * 255 * 104);
    g1_jac base; g1_from_affine(&base, G1_GEN_IMG);
    g1_jac *row = malloc(sizeof(g1_jac) * 255);
    for (int w = 0; w < 32; w++) {
        row[0] = base;
        for (int d = 1; d < 255; d++) g1_add(&row[d], &row[d - 1], &base);
        g1_batch_to_affine(&table[w][0][0], row, 255);
        g1_add(&base, &row[254], &base); /* 256 * base */
    }
    free(row);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    gen_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (gen_job){ seed, n * t / threads, n * (t + 1) / threads, out, table, 0 };
        pthread_create(&th[t], NULL, g1_gen_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(table);
}
static void *g2_gen_thread(void *arg_) {
    gen_job *job = arg_;
    const uint8_t (*table)[255][200] = job->table;
    size_t cnt = job->hi - job->lo;
    g2_jac *pts = malloc(sizeof(g2_jac) * cnt);
    u64 s = job->seed;
    for (size_t i = 0; i < job->lo * 4; i++) splitmix_next(&s);
    for (size_t i = 0; i < cnt; i++) {
        u64 v[4];
        for (int k = 0; k < 4; k++) v[k] = splitmix_next(&s);
        while (fr_geq_p(v)) fr_sub_p(v, v);
        g2_jac acc; g2_set_inf(&acc);
        for (int w = 0; w < 32; w++) {
            unsigned d = (v[w / 8] >> ((w % 8) * 8)) & 0xff;
            if (d) g2_madd(&acc, &acc, table[w][d - 1]);
        }
        pts[i] = acc;
    }
    g2_batch_to_affine(job->out + job->lo * 200, pts, cnt);
    free(pts);
    return NULL;
}
void bzko_g2_random_bases(u64 seed, uint8_t *out, size_t n, int threads) {
    bzko_init();
    uint8_t (*table)[255][200] = malloc(32 * 255 * 200);
    g2_jac base; g2_from_affine(&base, G2_GEN_IMG);
    g2_jac *row = malloc(sizeof(g2_jac) * 255);
    for (int w = 0; w < 32; w++) {
        row[0] = base;
        for (int d = 1; d < 255; d++) g2_add(&row[d], &row[d - 1], &base);
        g2_batch_to_affine(&table[w][0][0], row, 255);
        g2_add(&base, &row[254], &base);
    }
    free(row);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    gen_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (gen_job){ seed, n * t / threads, n * (t + 1) / threads, out, table, 1 };
        pthread_create(&th[t], NULL, g2_gen_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(table);
}

/* bellman multiexp: scalars are Montgomery Fr images (converted to canonical `Repr` first, as
 * the prover does with `.to_repr()` before calling multiexp) */
void bzko_msm_g1(const uint8_t *bases, const u64 *scalars_mont, size_t n, uint8_t *out, int threads) {
    bzko_init();
    u64 *e = malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) fr_from_mont(e + 4 * i, scalars_mont + 4 * i);
    g1_jac r; g1_multiexp(&r, bases, e, n, threads);
    g1_to_affine(out, &r);
    free(e);
}
void bzko_msm_g2(const uint8_t *bases, const u64 *scalars_mont, size_t n, uint8_t *out, int threads) {
    bzko_init();
    u64 *e = malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) fr_from_mont(e + 4 * i, scalars_mont + 4 * i);
    g2_jac r; g2_multiexp(&r, bases, e, n, threads);
    g2_to_affine(out, &r);
    free(e);
}
/* definition: sum of double-and-add products (arbiter for multiexp, small n only) */
void bzko_msm_g1_naive(const uint8_t *bases, const u64 *scalars_mont, size_t n, uint8_t *out) {
    bzko_init();
    g1_jac acc; g1_set_inf(&acc);
    for (size_t i = 0; i < n; i++) {
        u64 k[4]; fr_from_mont(k, scalars_mont + 4 * i);
        g1_jac p; g1_from_affine(&p, bases + 104 * i); g1_mul_u256(&p, &p, k); g1_add(&acc, &acc, &p);
    }
    g1_to_affine(out, &acc);
}
void bzko_msm_g2_naive(const uint8_t *bases, const u64 *scalars_mont, size_t n, uint8_t *out) {
    bzko_init();
    g2_jac acc; g2_set_inf(&acc);
    for (size_t i = 0; i < n; i++) {
        u64 k[4]; fr_from_mont(k, scalars_mont + 4 * i);
        g2_jac p; g2_from_affine(&p, bases + 200 * i); g2_mul_u256(&p, &p, k); g2_add(&acc, &acc, &p);
    }
    g2_to_affine(out, &acc);
}

/* ------------------------------------------------------------------ Poseidon
 * /root/reference/src/zk/poseidon/mod.rs:24-84.  `params` = the BZKPOSv1 table (canonical LE
 * constants); converted to Montgomery once per call batch. */
typedef struct { uint32_t t, rf, rp, nrc; u64 *rc; u64 *mds; } pos_params;
static pos_params g_pos[18];
static int g_pos_loaded = 0;

int bzko_poseidon_load(const uint8_t *blob, size_t len) {
    bzko_init();
    if (len < 12 || memcmp(blob, "BZKPOSv1", 8)) return -1;
    uint32_t n; memcpy(&n, blob + 8, 4);
    size_t off = 12;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t hdr[4]; memcpy(hdr, blob + off, 16); off += 16;
        uint32_t t = hdr[0];
        if (t < 2 || t > 17) return -2;
        pos_params *pp = &g_pos[t];
        pp->t = t; pp->rf = hdr[1]; pp->rp = hdr[2]; pp->nrc = hdr[3];
        free(pp->rc); free(pp->mds);
        pp->rc = malloc(32 * pp->nrc); pp->mds = malloc(32 * t * t);
        for (uint32_t k = 0; k < pp->nrc; k++) { u64 v[4]; memcpy(v, blob + off, 32); off += 32; fr_to_mont(pp->rc + 4 * k, v); }
        for (uint32_t k = 0; k < t * t; k++) { u64 v[4]; memcpy(v, blob + off, 32); off += 32; fr_to_mont(pp->mds + 4 * k, v); }
    }
    if (off != len) return -3;
    g_pos_loaded = 1;
    return 0;
}
static inline void fr_pow5(u64 *x) {
    u64 t[4]; fr_sqr(t, x); fr_sqr(t, t); fr_mul(x, x, t);
}
static void poseidon_permute(u64 *state, uint32_t t) {
    const pos_params *pp = &g_pos[t];
    u64 tmp[17 * 4];
    uint32_t off = 0;
    for (uint32_t rnd = 0; rnd < pp->rf + pp->rp; rnd++) {
        for (uint32_t i = 0; i < t; i++) fr_add(state + 4 * i, state + 4 * i, pp->rc + 4 * (off + i));
        off += t;
        if (rnd < pp->rf / 2 || rnd >= pp->rf / 2 + pp->rp) {
            for (uint32_t i = 0; i < t; i++) fr_pow5(state + 4 * i);
        } else {
            fr_pow5(state);
        }
        for (uint32_t j = 0; j < t; j++) {
            u64 acc[4] = {0, 0, 0, 0}, pr[4];
            for (uint32_t k = 0; k < t; k++) { fr_mul(pr, pp->mds + 4 * (j * t + k), state + 4 * k); fr_add(acc, acc, pr); }
            memcpy(tmp + 4 * j, acc, 32);
        }
        memcpy(state, tmp, 32 * t);
    }
}
typedef struct { const u64 *in; u64 *out; size_t lo, hi; uint32_t arity; } pos_job;
static void *pos_thread(void *arg_) {
    pos_job *job = arg_;
    uint32_t t = job->arity + 1;
    for (size_t h = job->lo; h < job->hi; h++) {
        u64 st[17 * 4];
        memset(st, 0, 32);
        memcpy(st + 4, job->in + 4 * job->arity * h, 32 * job->arity);
        poseidon_permute(st, t);
        memcpy(job->out + 4 * h, st + 4, 32);
    }
    return NULL;
}
/* n hashes of `arity` Montgomery inputs each (row-major [n][arity]) -> n Montgomery digests */
int bzko_poseidon(const u64 *in, size_t n, uint32_t arity, u64 *out, int threads) {
    if (!g_pos_loaded) return -1;
    if (arity < 1 || arity > 16) return -2;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pos_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (pos_job){ in, out, n * t / threads, n * (t + 1) / threads, arity };
        pthread_create(&th[t], NULL, pos_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* ------------------------------------------------------------------ NTT (bellman EvaluationDomain)
 * natural order in, natural order out; bit-reversal permutation then log n DIT passes
 * (bellman 0.14.0 domain.rs `serial_fft`), butterflies of one pass split across threads. */
static void fr_pow_u64(u64 *r, const u64 *a, u64 e) { u64 ee[1] = { e }; fr_pow(r, a, ee, 1); }
static void omega_for(u64 *w, unsigned log_n) {
    /* ROOT_OF_UNITY = 7^((r-1)/2^32); omega = ROOT_OF_UNITY^(2^(32-log_n)) */
    u64 seven[4], e[4];
    fr_set_u64(seven, 7);
    /* (r-1) >> 32 */
    u64 rm1[4]; memcpy(rm1, fr_P, 32); rm1[0] -= 1;
    for (int i = 0; i < 4; i++) e[i] = (rm1[i] >> 32) | (i < 3 ? rm1[i + 1] << 32 : 0);
    fr_pow(w, seven, e, 4);
    for (unsigned i = log_n; i < 32; i++) fr_sqr(w, w);
}
typedef struct { u64 *a; unsigned log_n; size_t m; const u64 *tw; size_t lo, hi; } ntt_job;
static void *ntt_pass_thread(void *arg_) {
    ntt_job *job = arg_;
    size_t m = job->m;
    /* butterfly index b in [lo,hi): group k = b / m, j = b % m */
    for (size_t b = job->lo; b < job->hi; b++) {
        size_t k = (b / m) * 2 * m, j = b % m;
        u64 *x = job->a + 4 * (k + j), *y = job->a + 4 * (k + j + m);
        u64 t[4];
        fr_mul(t, y, job->tw + 4 * j * ((((size_t)1 << job->log_n) / 2) / m));
        fr_sub(y, x, t);
        fr_add(x, x, t);
    }
    return NULL;
}
static void ntt_core(u64 *a, unsigned log_n, const u64 *omega, int threads) {
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; k++) {
        size_t rk = 0;
        for (unsigned b = 0; b < log_n; b++) rk |= ((k >> b) & 1) << (log_n - 1 - b);
        if (k < rk) { u64 t[4]; memcpy(t, a + 4 * k, 32); memcpy(a + 4 * k, a + 4 * rk, 32); memcpy(a + 4 * rk, t, 32); }
    }
    if (log_n == 0) return;
    /* twiddle table omega^i, i < n/2 */
    u64 *tw = malloc(32 * (n / 2 ? n / 2 : 1));
    memcpy(tw, fr_R1, 32);
    for (size_t i = 1; i < n / 2; i++) fr_mul(tw + 4 * i, tw + 4 * (i - 1), omega);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    for (size_t m = 1; m < n; m *= 2) {
        ntt_job jobs[256]; pthread_t th[256];
        int nt = (n / 2 < 4096) ? 1 : threads;
        for (int t = 0; t < nt; t++) {
            jobs[t] = (ntt_job){ a, log_n, m, tw, (n / 2) * t / nt, (n / 2) * (t + 1) / nt };
            if (nt > 1) pthread_create(&th[t], NULL, ntt_pass_thread, &jobs[t]);
        }
        if (nt == 1) ntt_pass_thread(&jobs[0]);
        else for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    }
    free(tw);
}
/* op: 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft   (EvaluationDomain method names) */
int bzko_ntt(u64 *a, unsigned log_n, int op, int threads) {
    bzko_init();
    if (log_n > 28) return -1;
    size_t n = (size_t)1 << log_n;
    u64 w[4], g[4], u[4];
    omega_for(w, log_n);
    fr_set_u64(g, 7);
    if (op == 2) { /* distribute_powers(7) */
        memcpy(u, fr_R1, 32);
        for (size_t i = 0; i < n; i++) { fr_mul(a + 4 * i, a + 4 * i, u); fr_mul(u, u, g); }
    }
    if (op == 1 || op == 3) fr_inv(w, w);
    ntt_core(a, log_n, w, threads);
    if (op == 1 || op == 3) {
        u64 minv[4]; fr_set_u64(minv, (u64)n); fr_inv(minv, minv);
        for (size_t i = 0; i < n; i++) fr_mul(a + 4 * i, a + 4 * i, minv);
    }
    if (op == 3) {
        u64 gi[4]; fr_inv(gi, g);
        memcpy(u, fr_R1, 32);
        for (size_t i = 0; i < n; i++) { fr_mul(a + 4 * i, a + 4 * i, u); fr_mul(u, u, gi); }
    }
    return 0;
}
/* a[i] *= (7^n - 1)^-1  — EvaluationDomain::divide_by_z_on_coset */
void bzko_divide_by_z_on_coset(u64 *a, unsigned log_n) {
    bzko_init();
    u64 g[4], z[4];
    fr_set_u64(g, 7); fr_pow_u64(z, g, (u64)1 << log_n); fr_sub(z, z, fr_R1); fr_inv(z, z);
    for (size_t i = 0; i < ((size_t)1 << log_n); i++) fr_mul(a + 4 * i, a + 4 * i, z);
}

/* ------------------------------------------------------------------ Groth16 (bellman 0.14.0 generator / prover)
 * Restates `groth16::generator::generate_parameters` and `groth16::prover::create_proof`
 * (un-vendored crate; reference call sites /root/reference/src/config/blockchain.rs:372-400 and
 * /root/reference/src/mpn/circuits/test.rs:135,175,215).  PARITY UNPINNED by the reference (no
 * proof bytes recorded anywhere); pinned against oracle/py/groth16.py and the pairing check.
 *
 * R1CS in CSR form, one matrix per side: rowptr[nrows+1] (u64), col[nnz] (u32 index into
 * z = inputs ++ aux, z[0] = ONE), val[nnz] (Montgomery Fr).  Both functions append bellman's
 * `Input(i) * 0 = 0` rows themselves. */
typedef struct { const u64 *rowptr; const uint32_t *col; const u64 *val; } csr_t;

static void csr_row_eval(u64 *out, const csr_t *m, size_t row, const u64 *z) {
    u64 acc[4] = {0, 0, 0, 0}, t[4];
    for (u64 k = m->rowptr[row]; k < m->rowptr[row + 1]; k++) {
        fr_mul(t, m->val + 4 * k, z + 4 * (size_t)m->col[k]);
        fr_add(acc, acc, t);
    }
    memcpy(out, acc, 32);
}
typedef struct { const csr_t *m; const u64 *z; u64 *out; size_t lo, hi; } eval_job;
static void *eval_thread(void *arg_) {
    eval_job *j = arg_;
    for (size_t r = j->lo; r < j->hi; r++) csr_row_eval(j->out + 4 * r, j->m, r, j->z);
    return NULL;
}
static void csr_eval(u64 *out, const csr_t *m, size_t nrows, const u64 *z, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    eval_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (eval_job){ m, z, out, nrows * t / threads, nrows * (t + 1) / threads };
        pthread_create(&th[t], NULL, eval_thread, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* out[i] = [k_i] base for Montgomery scalars k — fixed-base 8-bit windows, threaded */
typedef struct { const u64 *k; size_t lo, hi; uint8_t *out; const void *table; } fb_job;
static void *fb_g1_thread(void *arg_) {
    fb_job *job = arg_;
    const uint8_t (*table)[255][104] = job->table;
    size_t cnt = job->hi - job->lo;
    if (!cnt) return NULL;
    g1_jac *pts = malloc(sizeof(g1_jac) * cnt);
    for (size_t i = 0; i < cnt; i++) {
        u64 v[4]; fr_from_mont(v, job->k + 4 * (job->lo + i));
        g1_jac acc; g1_set_inf(&acc);
        for (int w = 0; w < 32; w++) { unsigned d = (v[w / 8] >> ((w % 8) * 8)) & 0xff; if (d) g1_madd(&acc, &acc, table[w][d - 1]); }
        pts[i] = acc;
    }
    g1_batch_to_affine(job->out + job->lo * 104, pts, cnt);
    free(pts);
    return NULL;
}
static void *fb_g2_thread(void *arg_) {
    fb_job *job = arg_;
    const uint8_t (*table)[255][200] = job->table;
    size_t cnt = job->hi - job->lo;
    if (!cnt) return NULL;
    g2_jac *pts = malloc(sizeof(g2_jac) * cnt);
    for (size_t i = 0; i < cnt; i++) {
        u64 v[4]; fr_from_mont(v, job->k + 4 * (job->lo + i));
        g2_jac acc; g2_set_inf(&acc);
        for (int w = 0; w < 32; w++) { unsigned d = (v[w / 8] >> ((w % 8) * 8)) & 0xff; if (d) g2_madd(&acc, &acc, table[w][d - 1]); }
        pts[i] = acc;
    }
    g2_batch_to_affine(job->out + job->lo * 200, pts, cnt);
    free(pts);
    return NULL;
}
void bzko_g1_fixed_base_mul(const uint8_t *base_img, const u64 *k_mont, size_t n, uint8_t *out, int threads) {
    bzko_init();
    uint8_t (*table)[255][104] = malloc(32 * 255 * 104);
    g1_jac base; g1_from_affine(&base, base_img);
    g1_jac *row = malloc(sizeof(g1_jac) * 255);
    for (int w = 0; w < 32; w++) {
        row[0] = base;
        for (int d = 1; d < 255; d++) g1_add(&row[d], &row[d - 1], &base);
        g1_batch_to_affine(&table[w][0][0], row, 255);
        g1_add(&base, &row[254], &base);
    }
    free(row);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    fb_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) { jobs[t] = (fb_job){ k_mont, n * t / threads, n * (t + 1) / threads, out, table }; pthread_create(&th[t], NULL, fb_g1_thread, &jobs[t]); }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(table);
}
void bzko_g2_fixed_base_mul(const uint8_t *base_img, const u64 *k_mont, size_t n, uint8_t *out, int threads) {
    bzko_init();
    uint8_t (*table)[255][200] = malloc(32 * 255 * 200);
    g2_jac base; g2_from_affine(&base, base_img);
    g2_jac *row = malloc(sizeof(g2_jac) * 255);
    for (int w = 0; w < 32; w++) {
        row[0] = base;
        for (int d = 1; d < 255; d++) g2_add(&row[d], &row[d - 1], &base);
        g2_batch_to_affine(&table[w][0][0], row, 255);
        g2_add(&base, &row[254], &base);
    }
    free(row);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    fb_job jobs[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) { jobs[t] = (fb_job){ k_mont, n * t / threads, n * (t + 1) / threads, out, table }; pthread_create(&th[t], NULL, fb_g2_thread, &jobs[t]); }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(table);
}

static unsigned log2_ceil(size_t n) { unsigned e = 0; size_t m = 1; while (m < n) { m <<= 1; e++; } return e; }

/* Setup scalars.  Outputs (Montgomery Fr): h_k[m-1], at[nv], bt[nv], ext_ic[num_inputs],
 * ext_l[num_aux];  returns log2(m).  The caller turns them into points with the fixed-base muls. */
int bzko_groth16_setup_scalars(u64 num_inputs, u64 num_aux, u64 ncons,
                               const u64 *a_rp, const uint32_t *a_col, const u64 *a_val,
                               const u64 *b_rp, const uint32_t *b_col, const u64 *b_val,
                               const u64 *c_rp, const uint32_t *c_col, const u64 *c_val,
                               const u64 *toxic /* tau, alpha, beta, gamma, delta */,
                               u64 *h_k, u64 *at, u64 *bt, u64 *ext_ic, u64 *ext_l, int threads) {
    bzko_init();
    const u64 *tau = toxic, *alpha = toxic + 4, *beta = toxic + 8, *gamma = toxic + 12, *delta = toxic + 16;
    size_t rows = ncons + num_inputs, nv = num_inputs + num_aux;
    unsigned log_m = log2_ceil(rows);
    size_t m = (size_t)1 << log_m;
    u64 *lag = malloc(32 * m);
    memcpy(lag, fr_R1, 32);
    for (size_t i = 1; i < m; i++) fr_mul(lag + 4 * i, lag + 4 * (i - 1), tau);
    u64 zt[4]; fr_mul(zt, lag + 4 * (m - 1), tau); fr_sub(zt, zt, fr_R1);       /* tau^m - 1 */
    u64 dinv[4], ginv[4], zd[4];
    fr_inv(dinv, delta); fr_inv(ginv, gamma); fr_mul(zd, zt, dinv);
    for (size_t i = 0; i + 1 < m; i++) fr_mul(h_k + 4 * i, lag + 4 * i, zd);    /* tau^i Z(tau)/delta */
    bzko_ntt(lag, log_m, 1, threads);                                             /* L_j(tau) */
    u64 *ct = calloc(nv, 32);
    memset(at, 0, 32 * nv); memset(bt, 0, 32 * nv);
    const u64 *rp[3] = { a_rp, b_rp, c_rp }; const uint32_t *cl[3] = { a_col, b_col, c_col }; const u64 *vl[3] = { a_val, b_val, c_val };
    u64 *dst[3] = { at, bt, ct };
    for (int s = 0; s < 3; s++)
        for (size_t j = 0; j < ncons; j++)
            for (u64 k = rp[s][j]; k < rp[s][j + 1]; k++) {
                u64 t[4]; fr_mul(t, vl[s] + 4 * k, lag + 4 * j);
                fr_add(dst[s] + 4 * (size_t)cl[s][k], dst[s] + 4 * (size_t)cl[s][k], t);
            }
    for (size_t i = 0; i < num_inputs; i++) fr_add(at + 4 * i, at + 4 * i, lag + 4 * (ncons + i));  /* Input(i)*0=0 rows */
    for (size_t v = 0; v < nv; v++) {
        u64 e[4], t[4];
        fr_mul(e, beta, at + 4 * v); fr_mul(t, alpha, bt + 4 * v); fr_add(e, e, t); fr_add(e, e, ct + 4 * v);
        if (v < num_inputs) fr_mul(ext_ic + 4 * v, e, ginv); else fr_mul(ext_l + 4 * (v - num_inputs), e, dinv);
    }
    free(ct); free(lag);
    return (int)log_m;
}

/* a/b/c evaluations (padded to m) and the h coefficient vector (m-1 scalars, Montgomery) */
int bzko_groth16_h(u64 num_inputs, u64 num_aux, u64 ncons,
                   const u64 *a_rp, const uint32_t *a_col, const u64 *a_val,
                   const u64 *b_rp, const uint32_t *b_col, const u64 *b_val,
                   const u64 *c_rp, const uint32_t *c_col, const u64 *c_val,
                   const u64 *z, u64 *h_out, int threads) {
    bzko_init();
    size_t rows = ncons + num_inputs;
    unsigned log_m = log2_ceil(rows);
    size_t m = (size_t)1 << log_m;
    u64 *ev[3];
    csr_t mats[3] = { { a_rp, a_col, a_val }, { b_rp, b_col, b_val }, { c_rp, c_col, c_val } };
    for (int s = 0; s < 3; s++) { ev[s] = calloc(m, 32); csr_eval(ev[s], &mats[s], ncons, z, threads); }
    for (size_t i = 0; i < num_inputs; i++) memcpy(ev[0] + 4 * (ncons + i), z + 4 * i, 32);
    for (int s = 0; s < 3; s++) { bzko_ntt(ev[s], log_m, 1, threads); bzko_ntt(ev[s], log_m, 2, threads); }
    for (size_t i = 0; i < m; i++) { fr_mul(ev[0] + 4 * i, ev[0] + 4 * i, ev[1] + 4 * i); fr_sub(ev[0] + 4 * i, ev[0] + 4 * i, ev[2] + 4 * i); }
    bzko_divide_by_z_on_coset(ev[0], log_m);
    bzko_ntt(ev[0], log_m, 3, threads);
    memcpy(h_out, ev[0], 32 * (m - 1));
    for (int s = 0; s < 3; s++) free(ev[s]);
    return (int)log_m;
}

/* final assembly from the five MSM answers (affine images) — tail of create_proof */
void bzko_groth16_assemble(const uint8_t *alpha_g1, const uint8_t *beta_g1, const uint8_t *beta_g2,
                           const uint8_t *delta_g1, const uint8_t *delta_g2,
                           const uint8_t *a_ans, const uint8_t *b1_ans, const uint8_t *b2_ans,
                           const uint8_t *h_ans, const uint8_t *l_ans, const u64 *r_mont, const u64 *s_mont,
                           uint8_t *out_a, uint8_t *out_b, uint8_t *out_c) {
    bzko_init();
    u64 r[4], s[4], rs[4], rsm[4];
    fr_from_mont(r, r_mont); fr_from_mont(s, s_mont); fr_mul(rsm, r_mont, s_mont); fr_from_mont(rs, rsm);
    g1_jac ga, gc, t, aa, b1;
    g2_jac gb, t2, b2;
    g1_from_affine(&t, delta_g1); g1_mul_u256(&ga, &t, r); g1_from_affine(&t, alpha_g1); g1_add(&ga, &ga, &t);
    g1_from_affine(&aa, a_ans); g1_add(&ga, &ga, &aa);
    g2_from_affine(&t2, delta_g2); g2_mul_u256(&gb, &t2, s); g2_from_affine(&t2, beta_g2); g2_add(&gb, &gb, &t2);
    g2_from_affine(&b2, b2_ans); g2_add(&gb, &gb, &b2);
    g1_from_affine(&t, delta_g1); g1_mul_u256(&gc, &t, rs);
    g1_from_affine(&t, alpha_g1); g1_mul_u256(&t, &t, s); g1_add(&gc, &gc, &t);
    g1_from_affine(&t, beta_g1); g1_mul_u256(&t, &t, r); g1_add(&gc, &gc, &t);
    g1_mul_u256(&t, &aa, s); g1_add(&gc, &gc, &t);
    g1_from_affine(&b1, b1_ans); g1_mul_u256(&t, &b1, r); g1_add(&gc, &gc, &t);
    g1_from_affine(&t, h_ans); g1_add(&gc, &gc, &t);
    g1_from_affine(&t, l_ans); g1_add(&gc, &gc, &t);
    g1_to_affine(out_a, &ga); g2_to_affine(out_b, &gb); g1_to_affine(out_c, &gc);
}

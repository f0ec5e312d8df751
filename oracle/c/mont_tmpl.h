/* ORACLE (test infrastructure only — never linked into the product library).
 *
 * Macro-templated Montgomery prime field on 64-bit limbs (unsigned __int128 products).
 * Instantiated twice by field.c:  Fr (4 limbs, R = 2^256)  and  Fp (6 limbs, R = 2^384).
 *
 * Restates what the reference gets from crates it does not vendor:
 *   Fr  <- ff 0.13 derive on `ZkScalar([u64;4])`   (/root/reference/src/zk/mod.rs:202-206)
 *   Fp  <- bls12_381 0.8.0 `Fp([u64;6])`           (/root/reference/src/zk/groth16/mod.rs:19-20)
 * Both keep elements fully reduced (< modulus) in Montgomery form; so does this file, which makes
 * the byte images comparable with memcmp.
 *
 * Required before inclusion:  F (name prefix), NL (limb count), F_MODULUS (initialiser list).
 */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(F, name)

static const u64 FN(P)[NL] = F_MODULUS;
static u64 FN(INV);     /* -p^-1 mod 2^64 */
static u64 FN(R1)[NL];  /* R mod p  (Montgomery one) */
static u64 FN(R2)[NL];  /* R^2 mod p */

static inline int FN(is_zero)(const u64 *a) {
    u64 t = 0;
    for (int i = 0; i < NL; i++) t |= a[i];
    return t == 0;
}
static inline int FN(eq)(const u64 *a, const u64 *b) {
    u64 t = 0;
    for (int i = 0; i < NL; i++) t |= a[i] ^ b[i];
    return t == 0;
}
static inline void FN(copy)(u64 *r, const u64 *a) {
    for (int i = 0; i < NL; i++) r[i] = a[i];
}
static inline void FN(zero)(u64 *r) {
    for (int i = 0; i < NL; i++) r[i] = 0;
}
static inline int FN(geq_p)(const u64 *a) {
    for (int i = NL - 1; i >= 0; i--) {
        if (a[i] > FN(P)[i]) return 1;
        if (a[i] < FN(P)[i]) return 0;
    }
    return 1;
}
static inline u64 FN(sub_p)(u64 *r, const u64 *a) { /* r = a - p, returns borrow */
    u128 br = 0;
    for (int i = 0; i < NL; i++) {
        u128 d = (u128)a[i] - FN(P)[i] - br;
        r[i] = (u64)d;
        br = (d >> 64) & 1;
    }
    return (u64)br;
}
static inline void FN(add)(u64 *r, const u64 *a, const u64 *b) {
    u128 c = 0;
    u64 t[NL];
    for (int i = 0; i < NL; i++) {
        c += (u128)a[i] + b[i];
        t[i] = (u64)c;
        c >>= 64;
    }
    /* both moduli leave headroom in the top limb, so c == 0 here */
    if (FN(geq_p)(t)) FN(sub_p)(r, t); else FN(copy)(r, t);
}
static inline void FN(sub)(u64 *r, const u64 *a, const u64 *b) {
    u128 br = 0;
    u64 t[NL];
    for (int i = 0; i < NL; i++) {
        u128 d = (u128)a[i] - b[i] - br;
        t[i] = (u64)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < NL; i++) {
            c += (u128)t[i] + FN(P)[i];
            t[i] = (u64)c;
            c >>= 64;
        }
    }
    FN(copy)(r, t);
}
static inline void FN(neg)(u64 *r, const u64 *a) {
    u64 z[NL] = {0};
    FN(sub)(r, z, a);
}
static inline void FN(dbl)(u64 *r, const u64 *a) { FN(add)(r, a, a); }

/* CIOS Montgomery product, r = a*b/R mod p, fully reduced */
static inline void FN(mul)(u64 *r, const u64 *a, const u64 *b) {
    u64 t[NL + 2];
    for (int i = 0; i < NL + 2; i++) t[i] = 0;
    for (int i = 0; i < NL; i++) {
        u128 cur;
        u64 carry = 0;
        for (int j = 0; j < NL; j++) {
            cur = (u128)a[j] * b[i] + t[j] + carry;
            t[j] = (u64)cur;
            carry = (u64)(cur >> 64);
        }
        cur = (u128)t[NL] + carry;
        t[NL] = (u64)cur;
        t[NL + 1] = (u64)(cur >> 64);
        u64 m = t[0] * FN(INV);
        cur = (u128)m * FN(P)[0] + t[0];
        carry = (u64)(cur >> 64);
        for (int j = 1; j < NL; j++) {
            cur = (u128)m * FN(P)[j] + t[j] + carry;
            t[j - 1] = (u64)cur;
            carry = (u64)(cur >> 64);
        }
        cur = (u128)t[NL] + carry;
        t[NL - 1] = (u64)cur;
        t[NL] = t[NL + 1] + (u64)(cur >> 64);
    }
    if (t[NL] || FN(geq_p)(t)) FN(sub_p)(r, t); else FN(copy)(r, t);
}
static inline void FN(sqr)(u64 *r, const u64 *a) { FN(mul)(r, a, a); }

static inline void FN(to_mont)(u64 *r, const u64 *a) { FN(mul)(r, a, FN(R2)); }
static inline void FN(from_mont)(u64 *r, const u64 *a) {
    u64 one[NL] = {1};
    FN(mul)(r, a, one);
}
static inline void FN(set_u64)(u64 *r, u64 v) {
    u64 t[NL] = {0};
    t[0] = v;
    FN(to_mont)(r, t);
}
/* r = a^e, e given as ne little-endian 64-bit limbs (plain integer) */
static void FN(pow)(u64 *r, const u64 *a, const u64 *e, int ne) {
    u64 acc[NL], base[NL];
    FN(copy)(acc, FN(R1));
    FN(copy)(base, a);
    int started = 0;
    for (int i = ne * 64 - 1; i >= 0; i--) {
        if (started) FN(sqr)(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) {
            if (started) FN(mul)(acc, acc, base); else { FN(copy)(acc, base); started = 1; }
        }
    }
    if (!started) FN(copy)(acc, FN(R1));
    FN(copy)(r, acc);
}
/* Fermat inverse, 0 -> 0 */
static void FN(inv)(u64 *r, const u64 *a) {
    u64 e[NL];
    FN(copy)(e, FN(P));
    e[0] -= 2; /* both moduli end in ...01 / ...ab: no borrow */
    FN(pow)(r, a, e, NL);
}
static void FN(init)(void) {
    /* INV by Newton iteration on 2-adic inverse */
    u64 p0 = FN(P)[0], x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
    FN(INV) = (u64)0 - x;
    /* R mod p and R^2 mod p by 2*NL*64 modular doublings of 1 */
    u64 t[NL] = {1};
    for (int i = 0; i < 2 * NL * 64; i++) {
        /* t = 2t mod p, on plain integers (add handles reduction) */
        FN(add)(t, t, t);
        if (i == NL * 64 - 1) FN(copy)(FN(R1), t);
    }
    FN(copy)(FN(R2), t);
}

#undef FN

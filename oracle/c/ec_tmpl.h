/* ORACLE (test infrastructure only — never linked into the product library).
 *
 * Macro-templated short-Weierstrass group y^2 = x^3 + b (a = 0) in Jacobian coordinates, plus
 * bellman 0.14.0's Pippenger `multiexp` (window rule, zero/one shortcuts, summation by parts,
 * top-down window combination) run with one thread per window, as bellman's `Worker` does.
 * bellman / bls12_381 are un-vendored crates.io dependencies of the reference
 * (/root/reference/Cargo.toml:27-28); the only call sites are the provers in
 * /root/reference/src/mpn/circuits/test.rs:135,175,215 and the gadget tests.
 *
 * Instantiated by curve.c for G1 (coordinates in Fp) and G2 (coordinates in Fp2).
 * Required before inclusion:
 *   E            name prefix
 *   FW           u64 words per coordinate
 *   AFF_BYTES    size of the affine wire image (x | y | inf flag | pad)
 *   FE_add/sub/mul/sqr/dbl/neg/inv/is_zero/eq/copy/one(r)   coordinate-field ops
 */
#define ECAT_(a, b) a##_##b
#define ECAT(a, b) ECAT_(a, b)
#define EN(name) ECAT(E, name)

typedef struct { u64 X[FW], Y[FW], Z[FW]; } EN(jac);

static inline void EN(set_inf)(EN(jac) *p) {
    FE_one(p->X); FE_one(p->Y);
    for (int i = 0; i < FW; i++) p->Z[i] = 0;
}
static inline int EN(is_inf)(const EN(jac) *p) { return FE_is_zero(p->Z); }

/* affine wire image -> Jacobian; image = x[FW] | y[FW] | u8 inf | pad */
static inline void EN(from_affine)(EN(jac) *p, const uint8_t *img) {
    if (img[FW * 16]) { EN(set_inf)(p); return; }
    memcpy(p->X, img, FW * 8);
    memcpy(p->Y, img + FW * 8, FW * 8);
    FE_one(p->Z);
}

static void EN(dbl)(EN(jac) *r, const EN(jac) *p) {
    if (EN(is_inf)(p)) { *r = *p; return; }
    u64 A[FW], B[FW], C[FW], D[FW], Ee[FW], Ff[FW], t[FW];
    EN(jac) o;
    FE_sqr(A, p->X);
    FE_sqr(B, p->Y);
    FE_sqr(C, B);
    FE_add(t, p->X, B); FE_sqr(t, t); FE_sub(t, t, A); FE_sub(t, t, C); FE_dbl(D, t);
    FE_dbl(Ee, A); FE_add(Ee, Ee, A);
    FE_sqr(Ff, Ee);
    FE_dbl(t, D); FE_sub(o.X, Ff, t);
    FE_dbl(C, C); FE_dbl(C, C); FE_dbl(C, C);
    FE_sub(t, D, o.X); FE_mul(t, Ee, t); FE_sub(o.Y, t, C);
    FE_mul(t, p->Y, p->Z); FE_dbl(o.Z, t);
    *r = o;
}

static void EN(add)(EN(jac) *r, const EN(jac) *p, const EN(jac) *q) {
    if (EN(is_inf)(p)) { *r = *q; return; }
    if (EN(is_inf)(q)) { *r = *p; return; }
    u64 Z1Z1[FW], Z2Z2[FW], U1[FW], U2[FW], S1[FW], S2[FW], H[FW], Rr[FW], HH[FW], HHH[FW], V[FW], t[FW];
    FE_sqr(Z1Z1, p->Z); FE_sqr(Z2Z2, q->Z);
    FE_mul(U1, p->X, Z2Z2); FE_mul(U2, q->X, Z1Z1);
    FE_mul(S1, p->Y, q->Z); FE_mul(S1, S1, Z2Z2);
    FE_mul(S2, q->Y, p->Z); FE_mul(S2, S2, Z1Z1);
    if (FE_eq(U1, U2)) {
        if (FE_eq(S1, S2)) { EN(dbl)(r, p); return; }
        EN(set_inf)(r); return;
    }
    EN(jac) o;
    FE_sub(H, U2, U1); FE_sub(Rr, S2, S1);
    FE_sqr(HH, H); FE_mul(HHH, H, HH); FE_mul(V, U1, HH);
    FE_sqr(t, Rr); FE_sub(t, t, HHH); FE_sub(t, t, V); FE_sub(o.X, t, V);
    FE_sub(t, V, o.X); FE_mul(t, Rr, t); FE_mul(S1, S1, HHH); FE_sub(o.Y, t, S1);
    FE_mul(t, p->Z, q->Z); FE_mul(o.Z, t, H);
    *r = o;
}

/* r = p + (affine image) */
static void EN(madd)(EN(jac) *r, const EN(jac) *p, const uint8_t *img) {
    EN(jac) q;
    if (img[FW * 16]) { *r = *p; return; }
    if (EN(is_inf)(p)) { EN(from_affine)(r, img); return; }
    memcpy(q.X, img, FW * 8); memcpy(q.Y, img + FW * 8, FW * 8);
    u64 Z1Z1[FW], U2[FW], S2[FW], H[FW], Rr[FW], HH[FW], HHH[FW], V[FW], t[FW];
    FE_sqr(Z1Z1, p->Z);
    FE_mul(U2, q.X, Z1Z1);
    FE_mul(S2, q.Y, p->Z); FE_mul(S2, S2, Z1Z1);
    if (FE_eq(p->X, U2)) {
        if (FE_eq(p->Y, S2)) { EN(dbl)(r, p); return; }
        EN(set_inf)(r); return;
    }
    EN(jac) o;
    FE_sub(H, U2, p->X); FE_sub(Rr, S2, p->Y);
    FE_sqr(HH, H); FE_mul(HHH, H, HH); FE_mul(V, p->X, HH);
    FE_sqr(t, Rr); FE_sub(t, t, HHH); FE_sub(t, t, V); FE_sub(o.X, t, V);
    FE_sub(t, V, o.X); FE_mul(t, Rr, t); FE_mul(S2, p->Y, HHH); FE_sub(o.Y, t, S2);
    FE_mul(o.Z, p->Z, H);
    *r = o;
}

static void EN(to_affine)(uint8_t *img, const EN(jac) *p) {
    memset(img, 0, AFF_BYTES);
    if (EN(is_inf)(p)) {
        u64 one[FW]; FE_one(one);
        memcpy(img + FW * 8, one, FW * 8); /* x = 0, y = 1, inf = 1 (bls12_381 identity image) */
        img[FW * 16] = 1;
        return;
    }
    u64 zi[FW], zi2[FW], x[FW], y[FW];
    FE_inv(zi, p->Z); FE_sqr(zi2, zi);
    FE_mul(x, p->X, zi2); FE_mul(zi2, zi2, zi); FE_mul(y, p->Y, zi2);
    memcpy(img, x, FW * 8); memcpy(img + FW * 8, y, FW * 8);
}

/* n Jacobian points -> n affine images with one inversion (Montgomery's trick) */
static void EN(batch_to_affine)(uint8_t *imgs, const EN(jac) *pts, size_t n) {
    u64 (*pre)[FW] = malloc(sizeof(u64[FW]) * (n + 1));
    u64 acc[FW]; FE_one(acc);
    for (size_t i = 0; i < n; i++) {
        FE_copy(pre[i], acc);
        if (!EN(is_inf)(&pts[i])) FE_mul(acc, acc, pts[i].Z);
    }
    u64 inv[FW]; FE_inv(inv, acc);
    for (size_t i = n; i-- > 0;) {
        uint8_t *img = imgs + i * AFF_BYTES;
        memset(img, 0, AFF_BYTES);
        if (EN(is_inf)(&pts[i])) {
            u64 one[FW]; FE_one(one);
            memcpy(img + FW * 8, one, FW * 8); img[FW * 16] = 1;
            continue;
        }
        u64 zi[FW], zi2[FW], x[FW], y[FW];
        FE_mul(zi, inv, pre[i]);
        FE_mul(inv, inv, pts[i].Z);
        FE_sqr(zi2, zi); FE_mul(x, pts[i].X, zi2); FE_mul(zi2, zi2, zi); FE_mul(y, pts[i].Y, zi2);
        memcpy(img, x, FW * 8); memcpy(img + FW * 8, y, FW * 8);
    }
    free(pre);
}

/* [k]P, k = 4 canonical little-endian limbs, left-to-right double-and-add */
static void EN(mul_u256)(EN(jac) *r, const EN(jac) *p, const u64 k[4]) {
    EN(jac) acc; EN(set_inf)(&acc);
    for (int i = 255; i >= 0; i--) {
        EN(dbl)(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) EN(add)(&acc, &acc, p);
    }
    *r = acc;
}

/* ---- bellman multiexp ------------------------------------------------------------------ */
typedef struct {
    const uint8_t *bases; const u64 *exps; /* exps: canonical (non-Montgomery) 4-limb integers */
    size_t n; unsigned c, skip; int handle_trivial;
    EN(jac) result;
} EN(win_job);

static void EN(window)(EN(win_job) *job) {
    const unsigned c = job->c, skip = job->skip;
    const size_t nb = ((size_t)1 << c) - 1;
    EN(jac) *buckets = malloc(sizeof(EN(jac)) * nb);
    for (size_t i = 0; i < nb; i++) EN(set_inf)(&buckets[i]);
    EN(jac) acc; EN(set_inf)(&acc);
    for (size_t i = 0; i < job->n; i++) {
        const u64 *e = job->exps + 4 * i;
        const uint8_t *b = job->bases + i * AFF_BYTES;
        if ((e[0] | e[1] | e[2] | e[3]) == 0) continue;
        if (e[0] == 1 && (e[1] | e[2] | e[3]) == 0) {
            if (job->handle_trivial) EN(madd)(&acc, &acc, b);
            continue;
        }
        /* (e >> skip) mod 2^c */
        unsigned w = skip / 64, s = skip % 64;
        u64 d = e[w] >> s;
        if (s && w + 1 < 4) d |= e[w + 1] << (64 - s);
        d &= ((u64)1 << c) - 1;
        if (d) EN(madd)(&buckets[d - 1], &buckets[d - 1], b);
    }
    EN(jac) run; EN(set_inf)(&run);
    for (size_t i = nb; i-- > 0;) {
        EN(add)(&run, &run, &buckets[i]);
        EN(add)(&acc, &acc, &run);
    }
    free(buckets);
    job->result = acc;
}

typedef struct { EN(win_job) *jobs; int njobs; volatile int *next; } EN(pool_arg);
static void *EN(pool_thread)(void *arg_) {
    EN(pool_arg) *arg = arg_;
    for (;;) {
        int j = __sync_fetch_and_add(arg->next, 1);
        if (j >= arg->njobs) break;
        EN(window)(&arg->jobs[j]);
    }
    return NULL;
}

static unsigned bellman_window_c(size_t n);

/* sum_i [e_i] B_i ; exps canonical; result Jacobian.
 * bellman runs one pool task per window; on hosts with more cores than windows the base range is
 * additionally cut into chunks (window sums are linear in the bases), so the baseline can use
 * every core it is given. */
static void EN(multiexp)(EN(jac) *out, const uint8_t *bases, const u64 *exps, size_t n, int threads) {
    unsigned c = bellman_window_c(n);
    int nwin = (255 + c - 1) / c;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    int nchunk = threads / nwin;
    if (nchunk < 1) nchunk = 1;
    if ((size_t)nchunk > n / 1024 + 1) nchunk = (int)(n / 1024 + 1);
    int njobs = nwin * nchunk;
    EN(win_job) *jobs = calloc(njobs, sizeof(*jobs));
    for (int w = 0; w < nwin; w++)
        for (int k = 0; k < nchunk; k++) {
            size_t lo = n * (size_t)k / nchunk, hi = n * (size_t)(k + 1) / nchunk;
            EN(win_job) *j = &jobs[w * nchunk + k];
            j->bases = bases + lo * AFF_BYTES; j->exps = exps + 4 * lo; j->n = hi - lo;
            j->c = c; j->skip = w * c; j->handle_trivial = (w == 0);
        }
    volatile int next = 0;
    EN(pool_arg) arg = { jobs, njobs, &next };
    pthread_t th[256];
    int nth = threads < njobs ? threads : njobs;
    for (int t = 1; t < nth; t++) pthread_create(&th[t], NULL, EN(pool_thread), &arg);
    EN(pool_thread)(&arg);
    for (int t = 1; t < nth; t++) pthread_join(th[t], NULL);
    for (int w = 0; w < nwin; w++)
        for (int k = 1; k < nchunk; k++) EN(add)(&jobs[w * nchunk].result, &jobs[w * nchunk].result, &jobs[w * nchunk + k].result);
    EN(jac) acc = jobs[(nwin - 1) * nchunk].result;
    for (int w = nwin - 2; w >= 0; w--) {
        for (unsigned k = 0; k < c; k++) EN(dbl)(&acc, &acc);
        EN(add)(&acc, &acc, &jobs[w * nchunk].result);
    }
    free(jobs);
    *out = acc;
}

#undef EN

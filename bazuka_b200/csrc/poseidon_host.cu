// bazuka_b200 — Poseidon on the host, for the single hashes a node makes outside of any batch.
//
// `impl ZkHasher for PoseidonHasher` (/root/reference/src/zk/mod.rs:491-511) is called one hash at a time from every
// corner of the node (state manager, transaction hashes, calldata checks) behind a global `Mutex<LruCache>`; a GPU
// launch per call would cost more than the hash.  The batched device kernels (csrc/poseidon.cu) serve the transition
// builders and the witness; this is the same permutation — `PoseidonState::permute`, /root/reference/src/zk/poseidon/mod.rs:24-79:
// every round adds the round constants, applies x^5 to all lanes (R_F/2 first and last rounds) or to lane 0
// (the R_P middle rounds), and multiplies by the MDS matrix; capacity lane 0 starts at zero, the digest is lane 1 —
// on the host field arithmetic of ff.cuh, with the constants of the same BZKPOSv1 table.  No context, no shared state:
// a `bzk_poseidon_host` is immutable after creation and may be used from any number of threads.
#include "common.cuh"

struct bzk_poseidon_host {
    struct Width {
        uint32_t t = 0, rf = 0, rp = 0;
        std::vector<bzk::Fr> rc, mds;  // Montgomery
    };
    Width w[18];
};

using namespace bzk;

namespace {
// One MDS row on 64-bit limbs with ONE Montgomery reduction: sum_k m[k] * s[k] is accumulated unreduced (t <= 17 products of two
// values < r: < 2^515, nine limbs), reduced once, and brought below r by conditional subtractions — the same value, bit for bit,
// as reducing every product (the device kernels do the same with 32-bit limbs: Fe::mul_wide / redc_wide in csrc/poseidon.cu).
inline Fr mds_row_dot(const Fr *m, const Fr *s, uint32_t t) {
    // product scanning (Comba): column k collects every a_i * b_j with i + j = k of every term in a three-word accumulator;
    // no data-dependent branches
    uint64_t acc[9];
    uint64_t c0 = 0, c1 = 0, c2 = 0;
    const uint64_t *A = (const uint64_t *)m, *B = (const uint64_t *)s;   // Fr = 8 x u32 = 4 x u64 on a little-endian host
    for (int col = 0; col < 7; col++) {
        const int lo = col < 4 ? 0 : col - 3, hi = col < 4 ? col : 3;
        for (uint32_t k = 0; k < t; k++) {
            const uint64_t *a = A + 4 * k, *b = B + 4 * k;
            for (int i = lo; i <= hi; i++) {
                const unsigned __int128 pr = (unsigned __int128)a[i] * b[col - i];
                const unsigned __int128 s0 = (unsigned __int128)c0 + (uint64_t)pr;
                c0 = (uint64_t)s0;
                const unsigned __int128 s1 = (unsigned __int128)c1 + (uint64_t)(pr >> 64) + (uint64_t)(s0 >> 64);
                c1 = (uint64_t)s1;
                c2 += (uint64_t)(s1 >> 64);
            }
        }
        acc[col] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
    acc[7] = c0; acc[8] = c1;
    uint64_t p[4];
    for (int i = 0; i < 4; i++) p[i] = (uint64_t)FrParams::p(2 * i) | ((uint64_t)FrParams::p(2 * i + 1) << 32);
    uint64_t inv = (uint64_t)(0u - FrParams::inv());   // p^-1 mod 2^32 ...
    inv *= 2 - p[0] * inv;                             // ... one Newton step: mod 2^64
    inv = (uint64_t)0 - inv;                           // -p^-1 mod 2^64
    uint64_t top = 0;                                  // limb 9
    for (int i = 0; i < 4; i++) {                      // Montgomery reduction by 2^256, limb by limb
        const uint64_t q = acc[i] * inv;
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (unsigned __int128)q * p[j] + acc[i + j];
            acc[i + j] = (uint64_t)c;
            c >>= 64;
        }
        for (int j = i + 4; j < 9; j++) {
            c += acc[j];
            acc[j] = (uint64_t)c;
            c >>= 64;
        }
        top += (uint64_t)c;
    }
    // acc[4..8] (+ top) < (17 r^2 + 2^256 r) / 2^256 < 9 r: a few conditional subtractions
    uint64_t r[6] = {acc[4], acc[5], acc[6], acc[7], acc[8], top};
    for (;;) {
        bool ge = r[4] != 0 || r[5] != 0;
        if (!ge) {
            ge = true;
            for (int i = 3; i >= 0; i--) {
                if (r[i] != p[i]) { ge = r[i] > p[i]; break; }
            }
        }
        if (!ge) break;
        uint64_t br = 0;
        for (int i = 0; i < 6; i++) {
            const unsigned __int128 d = (unsigned __int128)r[i] - (i < 4 ? p[i] : 0) - br;
            r[i] = (uint64_t)d;
            br = (uint64_t)(d >> 64) & 1;
        }
    }
    Fr out;
    memcpy(out.l, r, 32);
    return out;
}
}  // namespace

extern "C" {

int32_t bzk_poseidon_host_create(const uint8_t *blob, size_t len, bzk_poseidon_host **out) {
    if (!blob || !out || len < 12 || memcmp(blob, "BZKPOSv1", 8)) return BZK_ERR_BAD_ARG;
    auto *h = new (std::nothrow) bzk_poseidon_host();
    if (!h) return BZK_ERR_OOM;
    uint32_t nw;
    memcpy(&nw, blob + 8, 4);
    size_t off = 12;
    for (uint32_t i = 0; i < nw; i++) {
        uint32_t hdr[4];
        if (off + 16 > len) { delete h; return BZK_ERR_BAD_ARG; }
        memcpy(hdr, blob + off, 16);
        off += 16;
        const uint32_t t = hdr[0], nrc = hdr[3];
        if (t < 2 || t > 17 || nrc != t * (hdr[1] + hdr[2]) || off + 32 * ((size_t)nrc + (size_t)t * t) > len) { delete h; return BZK_ERR_BAD_ARG; }
        auto &W = h->w[t];
        W.t = t; W.rf = hdr[1]; W.rp = hdr[2];
        auto read = [&](std::vector<Fr> &dst, size_t cnt) {
            dst.resize(cnt);
            for (size_t k = 0; k < cnt; k++) {
                Fr v;
                memcpy(v.l, blob + off + 32 * k, 32);
                if (Fr::reduce_once(v) != v) return false;  // canonical constants are < r
                dst[k] = v.to_mont();
            }
            off += 32 * cnt;
            return true;
        };
        if (!read(W.rc, nrc) || !read(W.mds, (size_t)t * t)) { delete h; return BZK_ERR_BAD_ARG; }
    }
    if (off != len) { delete h; return BZK_ERR_BAD_ARG; }
    *out = h;
    return BZK_OK;
}

int32_t bzk_poseidon_host_free(bzk_poseidon_host *h) {
    delete h;
    return BZK_OK;
}

/* in[n][arity] -> out[n], Montgomery images (`ZkScalar`), arity 1..16 (`ZkHasher::MAX_ARITY`) */
int32_t bzk_poseidon_host_hash(const bzk_poseidon_host *h, uint32_t arity, const bzk_fr *in, size_t n, bzk_fr *out) {
    if (!h || arity < 1 || arity > 16 || (n && (!in || !out))) return BZK_ERR_BAD_ARG;
    const auto &W = h->w[arity + 1];
    if (W.t != arity + 1) return BZK_ERR_NO_PARAMS;
    const uint32_t t = W.t;
    Fr s[17], nx[17];
    for (size_t i = 0; i < n; i++) {
        s[0] = Fr::zero();
        for (uint32_t k = 0; k < arity; k++) memcpy(s[1 + k].l, &in[i * arity + k], 32);
        size_t off = 0;
        for (uint32_t rnd = 0; rnd < W.rf + W.rp; rnd++) {
            for (uint32_t k = 0; k < t; k++) s[k] = s[k] + W.rc[off + k];
            off += t;
            const bool full = rnd < W.rf / 2 || rnd >= W.rf / 2 + W.rp;
            for (uint32_t k = 0; k < (full ? t : 1u); k++) {
                const Fr x2 = s[k] * s[k];
                s[k] = x2 * x2 * s[k];
            }
            for (uint32_t j = 0; j < t; j++) nx[j] = mds_row_dot(W.mds.data() + (size_t)j * t, s, t);
            for (uint32_t k = 0; k < t; k++) s[k] = nx[k];
        }
        memcpy(&out[i], s[1].l, 32);
    }
    return BZK_OK;
}

}  // extern "C"

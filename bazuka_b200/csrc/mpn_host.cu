// bazuka_b200 — the MPN ledger and the update transition builder as native host code over the GPU primitives.
//
// Mirrors `mpn::update::update` (/root/reference/src/mpn/update.rs:8-299) on the state model of
// /root/reference/src/mpn/mod.rs:219-240 and /root/reference/src/zk/state/mod.rs:93-208 (account leaf =
// Poseidon-5(tx_nonce, withdraw_nonce, pk.x, pk.y, tokens_root); token leaf = Poseidon-2(token_id, amount);
// 4-ary sparse trees with `compress_default` defaults), but in the two-phase shape of DESIGN.md §3.7:
//   1. ledger decisions, sequential, no hashing (acceptance rules, balances, slot choice);
//   2. all hashing in batches on the GPU: bzk_poseidon_hash for the leaves, the versioned level-synchronous
//      tree update (poseidon.cu) for the token forest and the state tree.
// Output = one row of circuit inputs per slot in UpdateCircuit's allocation order (the witness program's RAW
// operands, `bazuka_b200/mpn/witness_program.py::raw_values`), the state root entering every slot, and the three
// state-dependent public inputs.  Scalars cross the ABI as canonical 32-byte little-endian integers.
#include "common.cuh"
#include "mpn_wire.cuh"
#include <algorithm>
#include <map>
#include <set>
#include <unordered_map>
#include <vector>

namespace bzk {
int32_t tree4_versioned_update(bzk_ctx *ctx, uint32_t depth, const uint32_t *d_tree_id, const uint64_t *d_idx, size_t n, Fr *d_vals,
                               const Fr *d_init_proofs, Fr *d_out_proofs);
void witness_program_shape(const bzk_witness_program *p, uint64_t *n_ops, uint32_t *n_raw, uint32_t *n_ext);
}
using namespace bzk;

namespace {

struct FrKey {  // Montgomery limbs as a map key
    uint32_t l[8];
    bool operator<(const FrKey &o) const { return std::lexicographical_compare(l, l + 8, o.l, o.l + 8); }
};
inline FrKey key_of(const Fr &a) { FrKey k; memcpy(k.l, a.l, 32); return k; }
inline Fr fr_from_canon(const bzk_fr *c) { Fr a; memcpy(a.l, c, 32); return a.to_mont(); }
inline void fr_to_canon(bzk_fr *out, const Fr &a) { Fr c = a.from_mont(); memcpy(out, c.l, 32); }
inline Fr fr_from_u64(uint64_t v) {
    Fr a = Fr::zero();
    a.l[0] = (uint32_t)v;
    a.l[1] = (uint32_t)(v >> 32);
    return a.to_mont();
}

struct Money { Fr token_id; uint64_t amount; };  // token_id in Montgomery form
struct Account {
    uint64_t tx_nonce = 0, withdraw_nonce = 0;
    Fr ax = Fr::zero(), ay = Fr::zero();
    std::map<uint32_t, Money> tokens;  // ordered: `find_token_index` scans slots in ascending order
};

struct Point { Fr x, y; };
constexpr size_t kDecompressCacheCap = 1u << 16;  // decompressed keys kept per ledger (cleared when full)

}  // namespace

struct bzk_mpn_state {
    uint32_t A = 0, T = 0;
    Fr jj_d;
    std::vector<Fr> defaults, tdefaults;                     // per level: state tree / token tree
    std::vector<std::unordered_map<uint64_t, Fr>> levels;    // sparse state tree, level 0 = leaves; defaults are not stored
    std::map<uint64_t, Account> accounts;
    // The chain's own tables, which the builders only READ (`get_mpn_account_indices`, `get_mpn_account_count`,
    // /root/reference/src/mpn/update.rs:29,47-70): they change when a block is applied, not when a batch is built.
    std::map<std::pair<FrKey, FrKey>, uint64_t> by_addr;      // address -> index of the first account holding it
    uint64_t account_count = 0;
    // `new_account_indices`: accounts created by the batches built so far on this fork (threaded through deposit ->
    // withdraw -> update by `prepare_works`, /root/reference/src/mpn/mod.rs:330,353-414); bzk_mpn_state_commit_accounts
    // moves them into the chain tables
    std::map<std::pair<FrKey, FrKey>, uint64_t> pending;
    uint64_t state_size = 0;  // ZkCompressedState::state_size = number of non-zero scalar leaves
    std::map<FrKey, Point> decompress_cache;

    Fr node(uint32_t lvl, uint64_t idx) const {
        auto it = levels[lvl].find(idx);
        return it == levels[lvl].end() ? defaults[lvl] : it->second;
    }
    void put(uint32_t lvl, uint64_t idx, const Fr &v) {
        if (v == defaults[lvl]) levels[lvl].erase(idx);
        else levels[lvl][idx] = v;
    }
    void prove(uint64_t idx, Fr *out /*[A][3]*/) const {
        for (uint32_t l = 0; l < A; l++) {
            const uint64_t base = (idx >> 2) << 2;
            int w = 0;
            for (uint64_t k = 0; k < 4; k++)
                if (base + k != idx) out[l * 3 + (w++)] = node(l, base + k);
            idx >>= 2;
        }
    }
    bool on_curve(const Fr &x, const Fr &y) const {
        Fr x2 = x * x, y2 = y * y;
        return (y2 - x2) == (Fr::one() + jj_d * x2 * y2);
    }
};

namespace {

// Fr square root (Tonelli-Shanks, r - 1 = 2^32 * q with q odd, 7 = non-residue); false when none exists.  One 223-bit power:
// w = a^((q-1)/2) gives both x = a*w = a^((q+1)/2) and t = x*w = a^q; the loop then corrects x by powers of c = 7^q (computed
// once); a non-residue is recognised at the end (x^2 != a) instead of by a separate Legendre power.
bool fr_sqrt(const Fr &a, Fr *out) {
    if (a.is_zero()) { *out = a; return true; }
    // q = (r - 1) >> 32, as 32-bit words (224 bits); e = (q - 1) / 2
    uint32_t q[8] = {0}, rm1[8], e[8];
    for (int i = 0; i < 8; i++) rm1[i] = FrParams::p(i);
    rm1[0] -= 1;
    for (int i = 0; i < 7; i++) q[i] = rm1[i + 1];
    for (int i = 0; i < 8; i++) e[i] = (q[i] >> 1) | (i < 7 ? q[i + 1] << 31 : 0);   // q is odd: (q - 1) / 2 = q >> 1
    static const Fr c0 = Fr::from_u32(7).pow(q, 8);   // generator of the 2^32-torsion
    const Fr w = a.pow(e, 8);
    Fr x = a * w, t = x * w, c = c0;
    uint32_t m = 32;
    while (!(t == Fr::one())) {
        uint32_t i = 0;
        Fr t2 = t;
        while (!(t2 == Fr::one())) {
            t2 = t2 * t2;
            if (++i == m) return false;   // t has order 2^m: a is not a square
        }
        Fr b = c;
        for (uint32_t k = 0; k + i + 1 < m; k++) b = b * b;
        m = i;
        c = b * b;
        t = t * c;
        x = x * b;
    }
    if (!(x * x == a)) return false;
    *out = x;
    return true;
}

// PointCompressed::decompress (/root/reference/src/crypto/jubjub/curve.rs:78-88); x canonical in, Montgomery out
bool jj_decompress(bzk_mpn_state *s, const bzk_fr *x_canon, bool odd, Point *out) {
    Fr x = fr_from_canon(x_canon);
    auto it = s->decompress_cache.find(key_of(x));
    Point p;
    if (it != s->decompress_cache.end()) p = it->second;
    else {
        Fr x2 = x * x;
        Fr den = Fr::one() - s->jj_d * x2;
        if (den.is_zero()) return false;
        Fr y;
        if (!fr_sqrt((Fr::one() + x2) * den.inv_gcd(), &y)) return false;  // a = -1
        p = Point{x, y};
        if (s->decompress_cache.size() >= kDecompressCacheCap) s->decompress_cache.clear();
        s->decompress_cache[key_of(x)] = p;
    }
    const bool y_odd = (p.y.from_mont().l[0] & 1u) != 0;
    out->x = p.x;
    out->y = (y_odd != odd) ? p.y.neg() : p.y;
    return true;
}

// non-zero scalar leaves of one account (`set_data` counts a leaf when it becomes non-zero,
// /root/reference/src/zk/state/mod.rs:327-341)
uint64_t leaf_count(const Account &a) {
    uint64_t n = (a.tx_nonce != 0) + (a.withdraw_nonce != 0) + !a.ax.is_zero() + !a.ay.is_zero();
    for (auto &kv : a.tokens) n += !kv.second.token_id.is_zero() + (kv.second.amount != 0);
    return n;
}
bool canonical(const bzk_fr &v) {
    Fr a;
    memcpy(a.l, &v, 32);
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] < FrParams::p(i)) return true;
        if (a.l[i] > FrParams::p(i)) return false;
    }
    return false;  // == r
}

int find_token_index(const Account &a, uint32_t T, const Fr &token_id, bool empty_allowed) {
    for (auto &kv : a.tokens)
        if (kv.second.token_id == token_id) return (int)kv.first;
    if (empty_allowed)
        for (uint32_t i = 0; i < (1u << (2 * T)); i++)
            if (!a.tokens.count(i)) return (int)i;
    return -1;
}

// host front-end of the versioned tree update: vals[(depth+1)*n] (vals[0..n) in), proofs [n][depth][3]
int32_t tree_update_host(bzk_ctx *ctx, uint32_t depth, const std::vector<uint32_t> &tid, const std::vector<uint64_t> &idx,
                         std::vector<Fr> &vals, const std::vector<Fr> &init, std::vector<Fr> &proofs) {
    const size_t n = idx.size();
    proofs.assign(n * depth * 3, Fr::zero());
    if (n == 0) return BZK_OK;
    size_t o_tid = 0, o_idx = (n * 4 + 255) & ~(size_t)255, o_vals = o_idx + ((n * 8 + 255) & ~(size_t)255),
           o_init = o_vals + (depth + 1) * n * sizeof(Fr), o_pr = o_init + n * depth * 3 * sizeof(Fr), total = o_pr + n * depth * 3 * sizeof(Fr);
    // the context's grow-only arena: no cudaMalloc / cudaFree (device-wide synchronisations) per batch
    BZK_TRY(ensure_ws(ctx, &ctx->ws, &ctx->ws_bytes, total));
    char *b = (char *)ctx->ws;
    cudaStream_t st = ctx->stream;
    BZK_CUDA(ctx, cudaMemcpyAsync(b + o_tid, tid.data(), n * 4, cudaMemcpyHostToDevice, st));
    BZK_CUDA(ctx, cudaMemcpyAsync(b + o_idx, idx.data(), n * 8, cudaMemcpyHostToDevice, st));
    BZK_CUDA(ctx, cudaMemcpyAsync(b + o_vals, vals.data(), n * sizeof(Fr), cudaMemcpyHostToDevice, st));
    BZK_CUDA(ctx, cudaMemcpyAsync(b + o_init, init.data(), n * depth * 3 * sizeof(Fr), cudaMemcpyHostToDevice, st));
    BZK_TRY(tree4_versioned_update(ctx, depth, (const uint32_t *)(b + o_tid), (const uint64_t *)(b + o_idx), n, (Fr *)(b + o_vals),
                                   (const Fr *)(b + o_init), (Fr *)(b + o_pr)));
    vals.resize((size_t)(depth + 1) * n);
    BZK_CUDA(ctx, cudaMemcpyAsync(vals.data(), b + o_vals, (depth + 1) * n * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    BZK_CUDA(ctx, cudaMemcpyAsync(proofs.data(), b + o_pr, n * depth * 3 * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    BZK_CUDA(ctx, cudaStreamSynchronize(st));
    return BZK_OK;
}

int32_t hash_rows(bzk_ctx *ctx, uint32_t arity, const std::vector<Fr> &rows, std::vector<Fr> &out) {
    const size_t n = rows.size() / arity;
    out.assign(n, Fr::zero());
    if (n == 0) return BZK_OK;
    return bzk_poseidon_hash(ctx, arity, (const bzk_fr *)rows.data(), n, (bzk_fr *)out.data());
}

// the token forest of a batch: pre-batch tokens of the touched accounts enter as writes into empty trees
struct Forest {
    uint32_t T;
    std::map<uint64_t, uint32_t> tree_of;
    std::vector<Fr> rows;  // [n][2]
    std::vector<uint32_t> tid;
    std::vector<uint64_t> idx;
    std::vector<Fr> vals, proofs;
    std::vector<Fr> cur;  // current root per tree while replaying
    size_t n_init = 0;
    size_t write(uint64_t acc, uint32_t index, const Money &m) {
        rows.push_back(m.token_id);
        rows.push_back(fr_from_u64(m.amount));
        tid.push_back(tree_of.at(acc));
        idx.push_back(index);
        return idx.size() - 1;
    }
    int32_t run(bzk_ctx *ctx, const std::vector<Fr> &tdef) {
        std::vector<Fr> leaves;
        BZK_TRY(hash_rows(ctx, 2, rows, leaves));
        const size_t n = idx.size();
        std::vector<Fr> init(n * T * 3);
        for (size_t e = 0; e < n; e++)
            for (uint32_t l = 0; l < T; l++)
                for (int k = 0; k < 3; k++) init[(e * T + l) * 3 + k] = tdef[l];
        vals = leaves;
        BZK_TRY(tree_update_host(ctx, T, tid, idx, vals, init, proofs));
        cur.assign(tree_of.size(), tdef[T]);
        for (size_t e = 0; e < n_init; e++) cur[tid[e]] = vals[(size_t)T * n + e];
        return BZK_OK;
    }
    Fr root(uint64_t acc) const { return cur[tree_of.at(acc)]; }
    Fr applied(uint64_t acc, size_t e) {
        const Fr r = vals[(size_t)T * idx.size() + e];
        cur[tree_of.at(acc)] = r;
        return r;
    }
};


// ---- JubJub on the host field arithmetic (projective twisted Edwards, a = -1): /root/reference/src/crypto/jubjub/curve.rs:90-160
struct JJ { Fr x, y, z; };
inline JJ jj_identity() { return JJ{Fr::zero(), Fr::one(), Fr::one()}; }
inline JJ jj_add(const JJ &p, const JJ &q, const Fr &d) {   // unified addition (also doubles)
    const Fr a = p.z * q.z, b = a * a, c = p.x * q.x, dd = p.y * q.y, e = d * c * dd, f = b - e, g = b + e;
    return JJ{a * f * ((p.x + p.y) * (q.x + q.y) - c - dd), a * g * (dd + c), f * g};
}
inline JJ jj_mul(const Point &p, const Fr &k_canon, const Fr &d) {
    const JJ base{p.x, p.y, Fr::one()};
    JJ acc = jj_identity();
    for (int i = 255; i >= 0; i--) {
        acc = jj_add(acc, acc, d);
        if ((k_canon.l[i >> 5] >> (i & 31)) & 1) acc = jj_add(acc, base, d);
    }
    return acc;
}
inline bool jj_equal(const JJ &p, const JJ &q) { return p.x * q.z == q.x * p.z && p.y * q.z == q.y * p.z; }
inline Point jj_base() {   // BASE (curve.rs:146-164): x below, y = 18
    Fr x;
    const uint32_t l[8] = {0xec7beacau, 0x4df7b7ffu, 0xfd6c54edu, 0x2e3ebb21u, 0x0fd6cce6u, 0xf1fbf02du, 0x43ac65a6u, 0x3fd2814cu};
    memcpy(x.l, l, 32);
    return Point{x.to_mont(), Fr::from_u32(18)};
}
// `JubJub::verify` (/root/reference/src/crypto/jubjub/mod.rs:151-167) given h = Poseidon(R.x, R.y, A.x, A.y, msg) (Montgomery)
inline bool eddsa_verify_with_h(const bzk_mpn_state *s, const Point &pk, const Point &sig_r, const Fr &sig_s_canon, const Fr &h_mont) {
    if (!s->on_curve(pk.x, pk.y) || !s->on_curve(sig_r.x, sig_r.y)) return false;
    const JJ lhs = jj_add(jj_mul(pk, h_mont.from_mont(), s->jj_d), JJ{sig_r.x, sig_r.y, Fr::one()}, s->jj_d);
    return jj_equal(lhs, jj_mul(jj_base(), sig_s_canon, s->jj_d));
}

}  // namespace

// `JubJub::verify` (/root/reference/src/crypto/jubjub/mod.rs:151-167) as a stand-alone host call: the signature check the
// withdraw builder applies (and a bank node applies to every MPN transaction before it enters the pool), on libbzk's host
// field arithmetic and the host Poseidon.  All scalars canonical.  Returns 1 (valid), 0 (invalid) or a negative status.
extern "C" int32_t bzk_jubjub_eddsa_verify(const bzk_poseidon_host *hasher, const bzk_fr *jubjub_d, const bzk_fr pk_xy[2], const bzk_fr *message,
                                           const bzk_fr sig_r_xy[2], const bzk_fr *sig_s) {
    if (!hasher || !jubjub_d || !pk_xy || !message || !sig_r_xy || !sig_s) return BZK_ERR_BAD_ARG;
    if (!canonical(pk_xy[0]) || !canonical(pk_xy[1]) || !canonical(sig_r_xy[0]) || !canonical(sig_r_xy[1]) || !canonical(*sig_s) || !canonical(*message))
        return 0;
    bzk_mpn_state tmp;
    tmp.jj_d = fr_from_canon(jubjub_d);
    const Point pk{fr_from_canon(pk_xy + 0), fr_from_canon(pk_xy + 1)}, r{fr_from_canon(sig_r_xy + 0), fr_from_canon(sig_r_xy + 1)};
    const Fr in[5] = {r.x, r.y, pk.x, pk.y, fr_from_canon(message)};
    Fr h;
    BZK_TRY(bzk_poseidon_host_hash(hasher, 5, (const bzk_fr *)in, 1, (bzk_fr *)&h));
    Fr s_canon;
    memcpy(s_canon.l, sig_s, 32);
    return eddsa_verify_with_h(&tmp, pk, r, s_canon, h) ? 1 : 0;
}

namespace {

// root of `List<log4 B>(Struct[...])` over the batch's rows (deposit.rs:178-218, withdraw.rs:190-245): one hash per row,
// then the 4-ary tree — every level one batched launch
int32_t list_root(bzk_ctx *ctx, uint32_t arity, const std::vector<Fr> &rows, Fr *out) {
    std::vector<Fr> cur;
    BZK_TRY(hash_rows(ctx, arity, rows, cur));
    while (cur.size() > 1) {
        std::vector<Fr> nxt;
        BZK_TRY(hash_rows(ctx, 4, cur, nxt));
        cur.swap(nxt);
    }
    *out = cur[0];
    return BZK_OK;
}

// ---- the builders' plans as the reference's transition structs (`prepare_works` puts them on the wire)
wire::Money wire_money(const Money &m) {
    wire::Money w;
    w.token = wire::ContractId::of_scalar(m.token_id);
    w.amount = m.amount;
    return w;
}
wire::Account wire_account(const Account &a) {
    wire::Account w;
    w.tx_nonce = (uint32_t)a.tx_nonce; w.withdraw_nonce = (uint32_t)a.withdraw_nonce;
    w.address.x = a.ax; w.address.y = a.ay;
    for (auto &kv : a.tokens) w.tokens.emplace_back((uint64_t)kv.first, wire_money(kv.second));   // ascending slot order
    return w;
}
wire::Proof wire_proof(const Fr *p, uint32_t depth) { return wire::Proof(p, p + (size_t)depth * 3); }
}  // namespace

extern "C" {

int32_t bzk_mpn_update_raw_width(uint32_t A, uint32_t T, uint32_t *n_raw) {
    if (!n_raw || A == 0 || A > 31 || T == 0 || T > 8) return BZK_ERR_BAD_ARG;
    *n_raw = 32 + 9 * T + 6 * A;
    return BZK_OK;
}

int32_t bzk_mpn_state_create(bzk_ctx *ctx, uint32_t log4_tree, uint32_t log4_token, const bzk_fr *jj_d_canon, bzk_mpn_state **out) {
    if (!ctx || !out || !jj_d_canon || log4_tree == 0 || log4_tree > 31 || log4_token == 0 || log4_token > 8) return BZK_ERR_BAD_ARG;
    auto *s = new (std::nothrow) bzk_mpn_state;
    if (!s) return BZK_ERR_OOM;
    s->A = log4_tree; s->T = log4_token;
    s->jj_d = fr_from_canon(jj_d_canon);
    // compress_default (/root/reference/src/zk/mod.rs:401-423): token leaf H(0,0), lists H([d;4]) per level,
    // account struct H(0,0,0,0,token-list default)
    std::vector<Fr> in, h;
    in.assign(2, Fr::zero());
    int32_t st = hash_rows(ctx, 2, in, h);
    s->tdefaults.push_back(st == BZK_OK ? h[0] : Fr::zero());
    for (uint32_t l = 0; st == BZK_OK && l < log4_token; l++) {
        in.assign(4, s->tdefaults.back());
        st = hash_rows(ctx, 4, in, h);
        if (st == BZK_OK) s->tdefaults.push_back(h[0]);
    }
    if (st == BZK_OK) {
        in.assign(5, Fr::zero());
        in[4] = s->tdefaults.back();
        st = hash_rows(ctx, 5, in, h);
        if (st == BZK_OK) s->defaults.push_back(h[0]);
    }
    for (uint32_t l = 0; st == BZK_OK && l < log4_tree; l++) {
        in.assign(4, s->defaults.back());
        st = hash_rows(ctx, 4, in, h);
        if (st == BZK_OK) s->defaults.push_back(h[0]);
    }
    if (st != BZK_OK) { delete s; return st; }
    s->levels.resize(log4_tree + 1);
    *out = s;
    return BZK_OK;
}

int32_t bzk_mpn_state_free(bzk_mpn_state *s) {
    delete s;
    return BZK_OK;
}

int32_t bzk_mpn_state_root(const bzk_mpn_state *s, bzk_fr *root) {
    if (!s || !root) return BZK_ERR_BAD_ARG;
    fr_to_canon(root, s->node(s->A, 0));
    return BZK_OK;
}

// `set_mpn_account` (/root/reference/src/zk/state/mod.rs:140-208): one account, sequential path re-hash (used to load
// a ledger; batches go through bzk_mpn_update_build)
int32_t bzk_mpn_state_set_account(bzk_ctx *ctx, bzk_mpn_state *s, uint64_t index, uint64_t tx_nonce, uint64_t withdraw_nonce,
                                  const bzk_fr *addr_x, const bzk_fr *addr_y, const uint32_t *token_index, const bzk_fr *token_id,
                                  const uint64_t *token_amount, uint32_t n_tokens) {
    if (!ctx || !s || !addr_x || !addr_y || (n_tokens && (!token_index || !token_id || !token_amount))) return BZK_ERR_BAD_ARG;
    if (index >> (2 * s->A)) return BZK_ERR_BAD_ARG;
    Account a;
    a.tx_nonce = tx_nonce; a.withdraw_nonce = withdraw_nonce;
    a.ax = fr_from_canon(addr_x); a.ay = fr_from_canon(addr_y);
    if (!canonical(*addr_x) || !canonical(*addr_y)) return BZK_ERR_BAD_ARG;
    for (uint32_t k = 0; k < n_tokens; k++) {
        if (token_index[k] >> (2 * s->T) || !canonical(token_id[k])) return BZK_ERR_BAD_ARG;
        const Fr id = fr_from_canon(token_id + k);
        if (id.is_zero()) continue;  // `get_mpn_account` drops slots whose token id is zero (state/mod.rs:127-130)
        a.tokens[token_index[k]] = Money{id, token_amount[k]};
    }
    Forest f;
    f.T = s->T;
    f.tree_of[index] = 0;
    for (auto &kv : a.tokens) f.write(index, kv.first, kv.second);
    f.n_init = f.idx.size();
    BZK_TRY(f.run(ctx, s->tdefaults));
    std::vector<Fr> row = {fr_from_u64(a.tx_nonce), fr_from_u64(a.withdraw_nonce), a.ax, a.ay, f.root(index)}, leaf;
    BZK_TRY(hash_rows(ctx, 5, row, leaf));
    std::vector<Fr> vals = {leaf[0]}, init(s->A * 3), proofs;
    s->prove(index, init.data());
    BZK_TRY(tree_update_host(ctx, s->A, {0u}, {index}, vals, init, proofs));
    uint64_t node = index;
    for (uint32_t l = 0; l <= s->A; l++) { s->put(l, node, vals[l]); node >>= 2; }
    auto old = s->accounts.find(index);
    if (old != s->accounts.end()) {
        s->state_size -= leaf_count(old->second);
        // an overwritten account gives its address back if the table pointed at this slot
        auto oit = s->by_addr.find(std::make_pair(key_of(old->second.ax), key_of(old->second.ay)));
        if (oit != s->by_addr.end() && oit->second == index && !(old->second.ax == a.ax && old->second.ay == a.ay)) s->by_addr.erase(oit);
    }
    s->state_size += leaf_count(a);
    s->accounts[index] = a;
    if (!(a.ax.is_zero() && a.ay.is_zero())) s->by_addr.emplace(std::make_pair(key_of(a.ax), key_of(a.ay)), index);
    s->account_count = std::max(s->account_count, index + 1);
    return BZK_OK;
}

/* An independent copy of the ledger (`db.fork_on_ram()`, /root/reference/src/mpn/mod.rs:313): build the batches of a
 * block on the copy and keep it only if the block is accepted — bzk_mpn_update_build writes the ledger it is given. */
int32_t bzk_mpn_state_clone(const bzk_mpn_state *s, bzk_mpn_state **out) {
    if (!s || !out) return BZK_ERR_BAD_ARG;
    auto *c = new (std::nothrow) bzk_mpn_state(*s);
    if (!c) return BZK_ERR_OOM;
    c->decompress_cache.clear();
    *out = c;
    return BZK_OK;
}
/* `ZkCompressedState { state_hash, state_size }` of the ledger (/root/reference/src/zk/mod.rs: state_size = non-zero scalar
 * leaves) and the chain-side account count */
int32_t bzk_mpn_state_info(const bzk_mpn_state *s, bzk_fr *state_hash, uint64_t *state_size, uint64_t *account_count, uint64_t *pending_accounts) {
    if (!s) return BZK_ERR_BAD_ARG;
    if (state_hash) fr_to_canon(state_hash, s->node(s->A, 0));
    if (state_size) *state_size = s->state_size;
    if (account_count) *account_count = s->account_count;
    if (pending_accounts) *pending_accounts = s->pending.size();
    return BZK_OK;
}
/* out[2] = {log4_tree, log4_token} the ledger was created with */
int32_t bzk_mpn_state_shape(const bzk_mpn_state *s, uint32_t out[2]) {
    if (!s || !out) return BZK_ERR_BAD_ARG;
    out[0] = s->A; out[1] = s->T;
    return BZK_OK;
}
/* `MpnWorkPool.final_delta` (/root/reference/src/mpn/mod.rs:17-45,416-417): every scalar leaf in which `after` (the fork
 * prepare_works returned) differs from `before`, as the bincode of `ZkDeltaPairs(HashMap<ZkDataLocator(Vec<u64>), Option<ZkScalar>>)`:
 * locator [account, field] for the four account scalars, [account, 4, token slot, 0 | 1] for a token's id / balance
 * (`set_mpn_account`, /root/reference/src/zk/state/mod.rs:140-208); a leaf that became zero is a `Remove` (None).  Entries in
 * ascending locator order.  Release the buffer with bzk_buffer_free. */
int32_t bzk_mpn_state_delta(const bzk_mpn_state *before, const bzk_mpn_state *after, uint8_t **bytes, size_t *len, uint64_t *n_entries) {
    if (!before || !after || !bytes || !len) return BZK_ERR_BAD_ARG;
    using Loc = std::vector<uint64_t>;
    auto leaves = [](const bzk_mpn_state *s, uint64_t idx, std::map<Loc, Fr> &out) {
        auto it = s->accounts.find(idx);
        if (it == s->accounts.end()) return;
        const Account &a = it->second;
        const Fr f[4] = {fr_from_u64(a.tx_nonce), fr_from_u64(a.withdraw_nonce), a.ax, a.ay};
        for (uint64_t k = 0; k < 4; k++) out[Loc{idx, k}] = f[k];
        for (auto &kv : a.tokens) {
            out[Loc{idx, 4, kv.first, 0}] = kv.second.token_id;
            out[Loc{idx, 4, kv.first, 1}] = fr_from_u64(kv.second.amount);
        }
    };
    std::set<uint64_t> touched;
    for (auto &kv : before->accounts) touched.insert(kv.first);
    for (auto &kv : after->accounts) touched.insert(kv.first);
    wire::Writer w;
    uint64_t n = 0;
    w.u64(0);
    for (uint64_t idx : touched) {
        std::map<Loc, Fr> o, a;
        leaves(before, idx, o);
        leaves(after, idx, a);
        std::set<Loc> locs;
        for (auto &kv : o) locs.insert(kv.first);
        for (auto &kv : a) locs.insert(kv.first);
        for (const Loc &l : locs) {
            const Fr ov = o.count(l) ? o[l] : Fr::zero(), nv = a.count(l) ? a[l] : Fr::zero();
            if (ov == nv) continue;
            w.u64(l.size());
            for (uint64_t x : l) w.u64(x);
            if (nv.is_zero()) w.u8(0);
            else { w.u8(1); w.fr(nv); }
            n++;
        }
    }
    memcpy(w.b.data(), &n, 8);
    uint8_t *buf = (uint8_t *)malloc(w.b.size());
    if (!buf) return BZK_ERR_OOM;
    memcpy(buf, w.b.data(), w.b.size());
    *bytes = buf; *len = w.b.size();
    if (n_entries) *n_entries = n;
    return BZK_OK;
}
/* The block built on this fork was applied: its new accounts enter the chain's index table. */
int32_t bzk_mpn_state_commit_accounts(bzk_mpn_state *s) {
    if (!s) return BZK_ERR_BAD_ARG;
    for (auto &kv : s->pending) {
        s->by_addr.emplace(kv.first, kv.second);
        s->account_count = std::max(s->account_count, kv.second + 1);
    }
    s->pending.clear();
    return BZK_OK;
}

int32_t bzk_mpn_update_build(bzk_ctx *ctx, bzk_mpn_state *s, const bzk_mpn_tx *txs, uint64_t n_txs, uint32_t log4_batch,
                             const bzk_fr *fee_token_canon, bzk_fr *raws, bzk_fr *ext, uint8_t *accepted, bzk_fr public3[3],
                             uint64_t *n_accepted) {
    return bzk::mpn_update_build_impl(ctx, s, txs, n_txs, log4_batch, fee_token_canon, raws, ext, accepted, public3, n_accepted, nullptr);
}
}  // extern "C"

int32_t bzk::mpn_update_build_impl(bzk_ctx *ctx, bzk_mpn_state *s, const bzk_mpn_tx *txs, uint64_t n_txs, uint32_t log4_batch,
                                   const bzk_fr *fee_token_canon, bzk_fr *raws, bzk_fr *ext, uint8_t *accepted, bzk_fr public3[3],
                                   uint64_t *n_accepted, bzk::UpdateSink *sink) {
    if (!ctx || !s || (n_txs && !txs) || !fee_token_canon || !raws || !ext || !public3 || !n_accepted || log4_batch > 8) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t A = s->A, T = s->T;
    const uint64_t cap = 1ull << (2 * log4_batch);
    const uint32_t n_raw = 32 + 9 * T + 6 * A;
    const Fr fee_token = fr_from_canon(fee_token_canon), prev_root = s->node(A, 0);
    // ---------------------------------------------------------------- phase 1: ledger decisions on a mirror
    struct Plan {
        uint64_t tx, src, dst;
        uint32_t sti, sfi, dti;
        Account src_before, src_mid, src_after, dst_before, dst_after;
        Money src_token, src_fee_token, dst_token;
        Point dst_addr;
        size_t e1, e2, e3;
        Fr src_bal_hash, dst_bal_hash;
    };
    std::map<uint64_t, Account> mirror;
    auto pending = s->pending;
    auto index_of = [&](const Point &a, uint64_t *out) {
        const auto key = std::make_pair(key_of(a.x), key_of(a.y));
        auto it = s->by_addr.find(key);
        if (it != s->by_addr.end()) { *out = it->second; return true; }
        auto jt = pending.find(key);
        if (jt != pending.end()) { *out = jt->second; return true; }
        return false;
    };
    auto get = [&](uint64_t i) -> Account {
        auto it = mirror.find(i);
        if (it != mirror.end()) return it->second;
        auto jt = s->accounts.find(i);
        Account a = jt == s->accounts.end() ? Account() : jt->second;
        mirror[i] = a;
        return a;
    };
    std::vector<Plan> plan;
    uint64_t fee_sum = 0;
    for (uint64_t k = 0; k < n_txs; k++) {
        if (accepted) accepted[k] = 0;
        if (plan.size() == cap) continue;
        const bzk_mpn_tx &tx = txs[k];
        // malformed field elements cannot be put into a witness row: such a transaction is simply not eligible
        if (!canonical(tx.src_pk_x) || !canonical(tx.dst_pk_x) || !canonical(tx.amount_token_id) || !canonical(tx.fee_token_id) ||
            !canonical(tx.sig_rx) || !canonical(tx.sig_ry) || !canonical(tx.sig_s))
            continue;
        const Fr fee_tok_id = fr_from_canon(&tx.fee_token_id), amt_tok_id = fr_from_canon(&tx.amount_token_id);
        // the reference's pre-filter (update.rs:31-38): fee token and both keys decompressible; filtered, not an error
        if (!(fee_tok_id == fee_token)) continue;
        Point src_addr, dst_addr;
        if (!jj_decompress(s, &tx.src_pk_x, tx.src_pk_odd != 0, &src_addr) || !jj_decompress(s, &tx.dst_pk_x, tx.dst_pk_odd != 0, &dst_addr))
            continue;
        // update.rs:47-70: the chain's index table first, then the accounts created earlier on this fork; an unknown
        // sender is rejected, an unknown receiver gets index  mpn_account_count + |new_account_indices|
        uint64_t src_index = 0, dst_index = 0;
        if (!index_of(src_addr, &src_index)) continue;
        bool dst_new = false;
        if (!index_of(dst_addr, &dst_index)) { dst_index = s->account_count + pending.size(); dst_new = true; }
        if (dst_index >> (2 * A)) continue;
        Account src_before = get(src_index), dst_before0 = get(dst_index);
        const int sti = find_token_index(src_before, T, amt_tok_id, false), dti = find_token_index(dst_before0, T, amt_tok_id, true),
                  sfi = find_token_index(src_before, T, fee_tok_id, false);
        if (sti < 0 || dti < 0 || sfi < 0) continue;
        const Money src_token = src_before.tokens[sti];
        const bool dst_has = dst_before0.tokens.count(dti) != 0;
        if (tx.nonce != src_before.tx_nonce + 1 || !(src_before.ax == src_addr.x) || !(src_before.ay == src_addr.y) ||
            (s->on_curve(dst_before0.ax, dst_before0.ay) && (!(dst_before0.ax == dst_addr.x) || !(dst_before0.ay == dst_addr.y))) ||
            (dst_has && !(src_token.token_id == dst_before0.tokens[dti].token_id)) || !(src_token.token_id == amt_tok_id) ||
            src_token.amount < tx.amount)
            continue;
        Account src_mid = src_before;
        src_mid.tx_nonce += 1;
        src_mid.tokens[sti].amount -= tx.amount;
        auto fit = src_mid.tokens.find(sfi);
        if (fit == src_mid.tokens.end() || !(fit->second.token_id == fee_tok_id) || fit->second.amount < tx.fee) continue;
        const Money src_fee_token = fit->second;
        Account src_after = src_mid;
        src_after.tokens[sfi].amount -= tx.fee;
        mirror[src_index] = src_after;
        Account dst_before = get(dst_index);
        Money dst_token{Fr::zero(), 0};
        if (dst_before.tokens.count(dti)) dst_token = dst_before.tokens[dti];
        Account dst_after = dst_before;
        dst_after.ax = dst_addr.x; dst_after.ay = dst_addr.y;
        if (!dst_after.tokens.count(dti)) dst_after.tokens[dti] = Money{amt_tok_id, 0};
        dst_after.tokens[dti].amount += tx.amount;
        mirror[dst_index] = dst_after;
        if (dst_new) pending.emplace(std::make_pair(key_of(dst_addr.x), key_of(dst_addr.y)), dst_index);
        Plan p{};
        p.tx = k; p.src = src_index; p.dst = dst_index; p.sti = sti; p.sfi = sfi; p.dti = dti;
        p.src_before = src_before; p.src_mid = src_mid; p.src_after = src_after; p.dst_before = dst_before; p.dst_after = dst_after;
        p.src_token = src_token; p.src_fee_token = src_fee_token; p.dst_token = dst_token; p.dst_addr = dst_addr;
        plan.push_back(std::move(p));
        if (accepted) accepted[k] = 1;
        fee_sum += tx.fee;
    }
    // ---------------------------------------------------------------- phase 2a: token forest
    Forest forest;
    forest.T = T;
    std::vector<uint64_t> touched;
    for (auto &p : plan)
        for (uint64_t i : {p.src, p.dst})
            if (forest.tree_of.emplace(i, (uint32_t)forest.tree_of.size()).second) touched.push_back(i);
    for (uint64_t acc : touched) {
        auto it = s->accounts.find(acc);
        if (it != s->accounts.end())
            for (auto &kv : it->second.tokens) forest.write(acc, kv.first, kv.second);
    }
    forest.n_init = forest.idx.size();
    for (auto &p : plan) {
        p.e1 = forest.write(p.src, p.sti, p.src_mid.tokens[p.sti]);
        p.e2 = forest.write(p.src, p.sfi, p.src_after.tokens[p.sfi]);
        p.e3 = forest.write(p.dst, p.dti, p.dst_after.tokens[p.dti]);
    }
    BZK_TRY(forest.run(ctx, s->tdefaults));
    std::vector<Fr> acct_rows;
    acct_rows.reserve(plan.size() * 15);
    auto push_acct = [&](const Account &a, const Fr &tok_root) {
        acct_rows.push_back(fr_from_u64(a.tx_nonce)); acct_rows.push_back(fr_from_u64(a.withdraw_nonce));
        acct_rows.push_back(a.ax); acct_rows.push_back(a.ay); acct_rows.push_back(tok_root);
    };
    for (auto &p : plan) {
        p.src_bal_hash = forest.root(p.src);
        const Fr r1 = forest.applied(p.src, p.e1), r2 = forest.applied(p.src, p.e2);
        p.dst_bal_hash = forest.root(p.dst);
        const Fr r3 = forest.applied(p.dst, p.e3);
        push_acct(p.src_mid, r1); push_acct(p.src_after, r2); push_acct(p.dst_after, r3);
    }
    // ---------------------------------------------------------------- phase 2b: state tree
    std::vector<Fr> s_vals, s_proofs;
    BZK_TRY(hash_rows(ctx, 5, acct_rows, s_vals));
    std::vector<uint64_t> s_idx;
    for (auto &p : plan) { s_idx.push_back(p.src); s_idx.push_back(p.src); s_idx.push_back(p.dst); }
    const size_t ne = s_idx.size();
    std::vector<Fr> init(ne * A * 3);
    for (size_t e = 0; e < ne; e++) s->prove(s_idx[e], init.data() + e * A * 3);
    BZK_TRY(tree_update_host(ctx, A, std::vector<uint32_t>(ne, 0u), s_idx, s_vals, init, s_proofs));
    // ---------------------------------------------------------------- rows of circuit inputs (raw_values order)
    const Fr null_dst_y = Fr::one().neg();  // PublicKey::default().decompress() = (0, -1)
    Fr root = prev_root;
    for (uint64_t slot = 0; slot < cap; slot++) {
        bzk_fr *row = raws + slot * n_raw;
        memset(row, 0, (size_t)n_raw * sizeof(bzk_fr));
        size_t w = 0;
        auto put_fr = [&](const Fr &v) { fr_to_canon(row + (w++), v); };
        auto put_u = [&](uint64_t v) { memcpy(row + (w++), &v, 8); };
        auto put_proof = [&](const Fr *p, uint32_t depth) { for (uint32_t i = 0; i < depth * 3; i++) put_fr(p[i]); };
        fr_to_canon(ext + slot * 2, fee_token);
        if (slot >= plan.size()) {
            // the only non-zero input of a null slot: tx.dst_pub_key.decompress().y
            fr_to_canon(row + (23 + 9 * T + 3 * A), null_dst_y);
            fr_to_canon(ext + slot * 2 + 1, root);  // after the last real slot the state no longer moves
            continue;
        }
        const Plan &p = plan[slot];
        const bzk_mpn_tx &tx = txs[p.tx];
        fr_to_canon(ext + slot * 2 + 1, root);
        put_u(1); put_u(p.sti); put_u(p.sfi); put_u(p.dti);
        put_u(p.src_before.tx_nonce); put_u(p.src_before.withdraw_nonce); put_fr(p.src_before.ax); put_fr(p.src_before.ay);
        put_fr(p.src_bal_hash); put_fr(p.dst_bal_hash);
        put_fr(p.src_token.token_id); put_u(p.src_token.amount);
        put_fr(p.src_fee_token.token_id); put_u(p.src_fee_token.amount);
        put_proof(forest.proofs.data() + p.e1 * T * 3, T);
        put_u(tx.amount); put_u(tx.fee);
        put_proof(forest.proofs.data() + p.e2 * T * 3, T);
        put_u(tx.nonce); put_u(p.src); row[w++] = tx.amount_token_id; row[w++] = tx.fee_token_id;
        put_fr(p.dst_token.token_id); put_u(p.dst_token.amount);
        put_proof(forest.proofs.data() + p.e3 * T * 3, T);
        put_proof(s_proofs.data() + (3 * slot) * A * 3, A);
        put_fr(p.dst_addr.x); put_fr(p.dst_addr.y); put_u(p.dst);
        put_u(p.dst_before.tx_nonce); put_u(p.dst_before.withdraw_nonce); put_fr(p.dst_before.ax); put_fr(p.dst_before.ay);
        put_proof(s_proofs.data() + (3 * slot + 2) * A * 3, A);
        row[w++] = tx.sig_rx; row[w++] = tx.sig_ry; row[w++] = tx.sig_s;
        if (w != n_raw) return BZK_ERR_BAD_ARG;
        root = s_vals[(size_t)A * ne + 3 * slot + 2];
    }
    // ---------------------------------------------------------------- commit + public inputs
    for (size_t e = 0; e < ne; e++) {
        uint64_t node = s_idx[e];
        for (uint32_t l = 0; l <= A; l++) { s->put(l, node, s_vals[(size_t)l * ne + e]); node >>= 2; }
    }
    for (uint64_t i : touched) {
        auto it = s->accounts.find(i);
        if (it != s->accounts.end()) s->state_size -= leaf_count(it->second);
        s->state_size += leaf_count(mirror[i]);
        s->accounts[i] = mirror[i];
    }
    s->pending = pending;
    std::vector<Fr> aux_in = {fee_token, fr_from_u64(fee_sum)}, aux_out;
    BZK_TRY(hash_rows(ctx, 2, aux_in, aux_out));
    fr_to_canon(public3 + 0, prev_root);
    fr_to_canon(public3 + 1, aux_out[0]);
    fr_to_canon(public3 + 2, root);
    *n_accepted = plan.size();
    if (sink) {   // `UpdateTransition` of every accepted transaction (/root/reference/src/mpn/update.rs:220-247)
        sink->t.clear(); sink->from.clear();
        for (size_t slot = 0; slot < plan.size(); slot++) {
            const Plan &p = plan[slot];
            wire::UpdateTransition t;
            t.enabled = true;
            t.src_before = wire_account(p.src_before); t.src_before_balances_hash = p.src_bal_hash;
            t.src_before_balance = wire_money(p.src_token); t.src_before_fee_balance = wire_money(p.src_fee_token);
            t.src_proof = wire_proof(s_proofs.data() + (3 * slot) * A * 3, A);
            t.src_index = p.src; t.src_token_index = p.sti; t.src_balance_proof = wire_proof(forest.proofs.data() + p.e1 * T * 3, T);
            t.src_fee_token_index = p.sfi; t.src_fee_balance_proof = wire_proof(forest.proofs.data() + p.e2 * T * 3, T);
            t.dst_before = wire_account(p.dst_before); t.dst_before_balances_hash = p.dst_bal_hash; t.dst_before_balance = wire_money(p.dst_token);
            t.dst_proof = wire_proof(s_proofs.data() + (3 * slot + 2) * A * 3, A);
            t.dst_index = p.dst; t.dst_token_index = p.dti; t.dst_balance_proof = wire_proof(forest.proofs.data() + p.e3 * T * 3, T);
            sink->t.push_back(std::move(t));
            sink->from.push_back(p.tx);
        }
    }
    return BZK_OK;
}

extern "C" {


/* `mpn::deposit::deposit` (/root/reference/src/mpn/deposit.rs:11-233) without the L1 balance bookkeeping (chain state):
 * up to 4^log4_batch eligible deposits, in order.  Outputs, one row per slot (null slots padded as DepositTransition::null):
 *   raws1[slots][5]           phase-1 inputs  {enabled, token, amount, pk.x, pk.y}
 *   raws2[slots][9+3T+3A]     phase-2 inputs  {account index, token index, account before (4), balances hash, balance before (2),
 *                             balance proof, account proof}
 *   roots[slots]              the state root entering each slot
 *   reveal[slots][4]          the rows the circuit reveals {enabled, token, amount, H(pk)}; aux_data = their list root
 *   public3                   {state, aux_data, next_state};   the ledger advances (build on a clone, see bzk_mpn_state_clone) */
int32_t bzk_mpn_deposit_build(bzk_ctx *ctx, bzk_mpn_state *s, const bzk_mpn_deposit *deps, uint64_t n_deps, uint32_t log4_batch, bzk_fr *raws1,
                              bzk_fr *raws2, bzk_fr *roots, bzk_fr *reveal, uint8_t *accepted, bzk_fr public3[3], uint64_t *n_accepted) {
    return bzk::mpn_deposit_build_impl(ctx, s, deps, n_deps, log4_batch, raws1, raws2, roots, reveal, accepted, public3, n_accepted, nullptr);
}
}  // extern "C"

int32_t bzk::mpn_deposit_build_impl(bzk_ctx *ctx, bzk_mpn_state *s, const bzk_mpn_deposit *deps, uint64_t n_deps, uint32_t log4_batch, bzk_fr *raws1,
                                    bzk_fr *raws2, bzk_fr *roots, bzk_fr *reveal, uint8_t *accepted, bzk_fr public3[3], uint64_t *n_accepted,
                                    bzk::DepositSink *sink) {
    if (!ctx || !s || (n_deps && !deps) || !raws1 || !raws2 || !roots || !reveal || !public3 || !n_accepted || log4_batch > 8) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t A = s->A, T = s->T, w2 = 9 + 3 * T + 3 * A;
    const uint64_t cap = 1ull << (2 * log4_batch);
    const Fr prev_root = s->node(A, 0);
    struct Plan { uint64_t k, idx; uint32_t ti; Account before, after; Money bal; Point addr; size_t e; Fr bal_hash; };
    std::map<uint64_t, Account> mirror;
    auto pending = s->pending;
    auto get = [&](uint64_t i) -> Account {
        auto it = mirror.find(i);
        if (it != mirror.end()) return it->second;
        auto jt = s->accounts.find(i);
        return jt == s->accounts.end() ? Account() : jt->second;
    };
    std::vector<Plan> plan;
    std::set<uint64_t> rejected_srcs;   // deposit.rs:33 `rejected_pub_keys`: a rejected deposit takes its L1 source's later ones with it
    for (uint64_t k = 0; k < n_deps; k++) {
        if (accepted) accepted[k] = 0;
        if (plan.size() == cap) continue;
        const bzk_mpn_deposit &d = deps[k];
        auto reject = [&] { if (d.src_id) rejected_srcs.insert(d.src_id); };
        if (!canonical(d.pk_x) || !canonical(d.token_id)) { reject(); continue; }
        Point addr;
        if (!jj_decompress(s, &d.pk_x, d.pk_odd != 0, &addr)) { reject(); continue; }
        const auto key = std::make_pair(key_of(addr.x), key_of(addr.y));
        uint64_t idx = 0;
        bool is_new = false;
        auto it = s->by_addr.find(key);
        if (it != s->by_addr.end()) idx = it->second;
        else {
            auto jt = pending.find(key);
            if (jt != pending.end()) idx = jt->second;
            else { idx = s->account_count + pending.size(); is_new = true; }
        }
        if (idx >> (2 * A)) { reject(); continue; }
        const Account before = get(idx);
        const Fr tok = fr_from_canon(&d.token_id);
        const int ti = find_token_index(before, T, tok, true);
        if (ti < 0 || (d.src_id && rejected_srcs.count(d.src_id)) ||
            (s->on_curve(before.ax, before.ay) && (!(before.ax == addr.x) || !(before.ay == addr.y)))) {
            reject();
            continue;
        }
        Plan p{};
        p.k = k; p.idx = idx; p.ti = (uint32_t)ti; p.before = before; p.addr = addr;
        p.bal = before.tokens.count(ti) ? before.tokens.at(ti) : Money{Fr::zero(), 0};
        p.after = before;
        p.after.ax = addr.x; p.after.ay = addr.y;
        if (!p.after.tokens.count(ti)) p.after.tokens[ti] = Money{tok, 0};
        p.after.tokens[ti].amount += d.amount;
        mirror[idx] = p.after;
        if (is_new) pending.emplace(key, idx);
        plan.push_back(std::move(p));
        if (accepted) accepted[k] = 1;
    }
    Forest forest;
    forest.T = T;
    std::vector<uint64_t> touched;
    for (auto &p : plan)
        if (forest.tree_of.emplace(p.idx, (uint32_t)forest.tree_of.size()).second) touched.push_back(p.idx);
    for (uint64_t acc : touched) {
        auto it = s->accounts.find(acc);
        if (it != s->accounts.end())
            for (auto &kv : it->second.tokens) forest.write(acc, kv.first, kv.second);
    }
    forest.n_init = forest.idx.size();
    for (auto &p : plan) p.e = forest.write(p.idx, p.ti, p.after.tokens[p.ti]);
    BZK_TRY(forest.run(ctx, s->tdefaults));
    std::vector<Fr> acct_rows, pk_rows;
    for (auto &p : plan) {
        p.bal_hash = forest.root(p.idx);
        const Fr r = forest.applied(p.idx, p.e);
        const Account &a = p.after;
        acct_rows.push_back(fr_from_u64(a.tx_nonce)); acct_rows.push_back(fr_from_u64(a.withdraw_nonce));
        acct_rows.push_back(a.ax); acct_rows.push_back(a.ay); acct_rows.push_back(r);
        pk_rows.push_back(p.addr.x); pk_rows.push_back(p.addr.y);
    }
    std::vector<Fr> s_vals, s_proofs, pk_hash;
    BZK_TRY(hash_rows(ctx, 5, acct_rows, s_vals));
    BZK_TRY(hash_rows(ctx, 2, pk_rows, pk_hash));
    std::vector<uint64_t> s_idx;
    for (auto &p : plan) s_idx.push_back(p.idx);
    const size_t ne = s_idx.size();
    std::vector<Fr> init(ne * A * 3);
    for (size_t e = 0; e < ne; e++) s->prove(s_idx[e], init.data() + e * A * 3);
    BZK_TRY(tree_update_host(ctx, A, std::vector<uint32_t>(ne, 0u), s_idx, s_vals, init, s_proofs));
    const Fr minus_one = Fr::one().neg();
    Fr root = prev_root;
    std::vector<Fr> rev_rows(cap * 4, Fr::zero());
    for (uint64_t slot = 0; slot < cap; slot++) {
        bzk_fr *r1 = raws1 + slot * 5, *r2 = raws2 + slot * w2, *rv = reveal + slot * 4;
        memset(r1, 0, 5 * sizeof(bzk_fr)); memset(r2, 0, (size_t)w2 * sizeof(bzk_fr)); memset(rv, 0, 4 * sizeof(bzk_fr));
        if (slot >= plan.size()) {
            fr_to_canon(r1 + 4, minus_one);   // PublicKey::default().decompress() = (0, -1)
            continue;
        }
        const Plan &p = plan[slot];
        const bzk_mpn_deposit &d = deps[p.k];
        fr_to_canon(roots + slot, root);
        uint64_t one = 1;
        memcpy(r1 + 0, &one, 8); r1[1] = d.token_id; memcpy(r1 + 2, &d.amount, 8); fr_to_canon(r1 + 3, p.addr.x); fr_to_canon(r1 + 4, p.addr.y);
        size_t w = 0;
        auto put_fr = [&](const Fr &v) { fr_to_canon(r2 + (w++), v); };
        auto put_u = [&](uint64_t v) { memcpy(r2 + (w++), &v, 8); };
        put_u(p.idx); put_u(p.ti); put_u(p.before.tx_nonce); put_u(p.before.withdraw_nonce); put_fr(p.before.ax); put_fr(p.before.ay);
        put_fr(p.bal_hash); put_fr(p.bal.token_id); put_u(p.bal.amount);
        for (uint32_t i = 0; i < T * 3; i++) put_fr(forest.proofs[p.e * T * 3 + i]);
        for (uint32_t i = 0; i < A * 3; i++) put_fr(s_proofs[slot * A * 3 + i]);
        if (w != w2) return BZK_ERR_BAD_ARG;
        memcpy(rv + 0, &one, 8); rv[1] = d.token_id; memcpy(rv + 2, &d.amount, 8); fr_to_canon(rv + 3, pk_hash[slot]);
        rev_rows[slot * 4 + 0] = Fr::one(); rev_rows[slot * 4 + 1] = fr_from_canon(&d.token_id); rev_rows[slot * 4 + 2] = fr_from_u64(d.amount);
        rev_rows[slot * 4 + 3] = pk_hash[slot];
        root = s_vals[(size_t)A * ne + slot];
    }
    for (uint64_t slot = plan.size(); slot < cap; slot++) fr_to_canon(roots + slot, root);
    for (size_t e = 0; e < ne; e++) {
        uint64_t node = s_idx[e];
        for (uint32_t l = 0; l <= A; l++) { s->put(l, node, s_vals[(size_t)l * ne + e]); node >>= 2; }
    }
    for (uint64_t i : touched) {
        auto it = s->accounts.find(i);
        if (it != s->accounts.end()) s->state_size -= leaf_count(it->second);
        s->state_size += leaf_count(mirror[i]);
        s->accounts[i] = mirror[i];
    }
    s->pending = pending;
    Fr aux;
    BZK_TRY(list_root(ctx, 4, rev_rows, &aux));
    fr_to_canon(public3 + 0, prev_root);
    fr_to_canon(public3 + 1, aux);
    fr_to_canon(public3 + 2, root);
    *n_accepted = plan.size();
    if (sink) {   // `DepositTransition` of every accepted deposit (/root/reference/src/mpn/deposit.rs:150-165)
        sink->t.clear(); sink->from.clear();
        for (size_t slot = 0; slot < plan.size(); slot++) {
            const Plan &p = plan[slot];
            wire::DepositTransition t;
            t.enabled = true;
            t.before = wire_account(p.before); t.before_balances_hash = p.bal_hash; t.before_balance = wire_money(p.bal);
            t.proof = wire_proof(s_proofs.data() + slot * A * 3, A);
            t.account_index = p.idx; t.token_index = p.ti; t.balance_proof = wire_proof(forest.proofs.data() + p.e * T * 3, T);
            sink->t.push_back(std::move(t));
            sink->from.push_back(p.k);
        }
    }
    return BZK_OK;
}

extern "C" {

/* `mpn::withdraw::withdraw` (/root/reference/src/mpn/withdraw.rs:10-259): nonce, balances and the EdDSA signature over
 * Poseidon(fingerprint, nonce) are checked here (the hashes of a batch in two launches, the scalar multiplications on the
 * host).  Rows: raws1[slots][12] {enabled, token, amount, fee token, fee, fingerprint, pk.x, pk.y, nonce, sig.r.x, sig.r.y, sig.s},
 * raws2[slots][12+6T+3A] {account index, token index, fee token index, account before (4), token-tree hash, balance before (2),
 * its proof, fee balance before (2), its proof, account proof}, reveal[slots][7] {enabled, token, amount, fee token, fee,
 * fingerprint, calldata}. */
int32_t bzk_mpn_withdraw_build(bzk_ctx *ctx, bzk_mpn_state *s, const bzk_mpn_withdraw *wds, uint64_t n_wds, uint32_t log4_batch, bzk_fr *raws1,
                               bzk_fr *raws2, bzk_fr *roots, bzk_fr *reveal, uint8_t *accepted, bzk_fr public3[3], uint64_t *n_accepted) {
    return bzk::mpn_withdraw_build_impl(ctx, s, wds, n_wds, log4_batch, raws1, raws2, roots, reveal, accepted, public3, n_accepted, nullptr);
}
}  // extern "C"

int32_t bzk::mpn_withdraw_build_impl(bzk_ctx *ctx, bzk_mpn_state *s, const bzk_mpn_withdraw *wds, uint64_t n_wds, uint32_t log4_batch, bzk_fr *raws1,
                                     bzk_fr *raws2, bzk_fr *roots, bzk_fr *reveal, uint8_t *accepted, bzk_fr public3[3], uint64_t *n_accepted,
                                     bzk::WithdrawSink *sink) {
    if (!ctx || !s || (n_wds && !wds) || !raws1 || !raws2 || !roots || !reveal || !public3 || !n_accepted || log4_batch > 8) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t A = s->A, T = s->T, w2 = 12 + 6 * T + 3 * A;
    const uint64_t cap = 1ull << (2 * log4_batch);
    const Fr prev_root = s->node(A, 0);
    // signature material of every candidate in two batched launches: msg = H(fingerprint, nonce), h = H(R.x, R.y, A.x, A.y, msg)
    std::vector<Point> addr(n_wds);
    std::vector<uint8_t> ok(n_wds, 0);
    std::vector<Fr> msg_rows, h_rows, msgs, hs;
    for (uint64_t k = 0; k < n_wds; k++) {
        const bzk_mpn_withdraw &w = wds[k];
        ok[k] = canonical(w.pk_x) && canonical(w.amount_token_id) && canonical(w.fee_token_id) && canonical(w.fingerprint) && canonical(w.sig_rx) &&
                canonical(w.sig_ry) && canonical(w.sig_s) && jj_decompress(s, &w.pk_x, w.pk_odd != 0, &addr[k]);
        msg_rows.push_back(ok[k] ? fr_from_canon(&w.fingerprint) : Fr::zero());
        msg_rows.push_back(fr_from_u64(w.nonce));
    }
    BZK_TRY(hash_rows(ctx, 2, msg_rows, msgs));
    for (uint64_t k = 0; k < n_wds; k++) {
        const bzk_mpn_withdraw &w = wds[k];
        h_rows.push_back(ok[k] ? fr_from_canon(&w.sig_rx) : Fr::zero()); h_rows.push_back(ok[k] ? fr_from_canon(&w.sig_ry) : Fr::zero());
        h_rows.push_back(ok[k] ? addr[k].x : Fr::zero()); h_rows.push_back(ok[k] ? addr[k].y : Fr::zero()); h_rows.push_back(msgs[k]);
    }
    BZK_TRY(hash_rows(ctx, 5, h_rows, hs));
    // `verify_calldata` (src/core/transaction.rs:177-182, withdraw.rs:77) for the entries that carry the payment's calldata
    bool any_calldata = false;
    for (uint64_t k = 0; k < n_wds; k++) any_calldata |= wds[k].check_calldata != 0;
    if (any_calldata) {
        std::vector<Fr> rows, cd;
        for (uint64_t k = 0; k < n_wds; k++) {
            const bzk_mpn_withdraw &w = wds[k];
            const bool on = ok[k] && w.check_calldata;
            rows.push_back(on ? addr[k].x : Fr::zero()); rows.push_back(on ? addr[k].y : Fr::zero()); rows.push_back(fr_from_u64(w.nonce));
            rows.push_back(on ? fr_from_canon(&w.sig_rx) : Fr::zero()); rows.push_back(on ? fr_from_canon(&w.sig_ry) : Fr::zero());
            rows.push_back(on ? fr_from_canon(&w.sig_s) : Fr::zero());
        }
        BZK_TRY(hash_rows(ctx, 6, rows, cd));
        for (uint64_t k = 0; k < n_wds; k++)
            if (ok[k] && wds[k].check_calldata && (!canonical(wds[k].calldata) || !(cd[k] == fr_from_canon(&wds[k].calldata)))) ok[k] = 0;
    }
    struct Plan { uint64_t k, idx; uint32_t ti, fi; Account before, mid, after; Money tok, fee_before; size_t e1, e2; Fr tok_hash; };
    std::map<uint64_t, Account> mirror;
    auto get = [&](uint64_t i) -> Account {
        auto it = mirror.find(i);
        if (it != mirror.end()) return it->second;
        auto jt = s->accounts.find(i);
        return jt == s->accounts.end() ? Account() : jt->second;
    };
    std::vector<Plan> plan;
    for (uint64_t k = 0; k < n_wds; k++) {
        if (accepted) accepted[k] = 0;
        if (plan.size() == cap || !ok[k]) continue;
        const bzk_mpn_withdraw &w = wds[k];
        const auto key = std::make_pair(key_of(addr[k].x), key_of(addr[k].y));
        uint64_t idx = 0;
        auto it = s->by_addr.find(key);
        if (it != s->by_addr.end()) idx = it->second;
        else {
            auto jt = s->pending.find(key);
            if (jt == s->pending.end()) continue;
            idx = jt->second;
        }
        const Account before = get(idx);
        const Fr tok = fr_from_canon(&w.amount_token_id), ftok = fr_from_canon(&w.fee_token_id);
        const int ti = find_token_index(before, T, tok, false), fi = find_token_index(before, T, ftok, false);
        if (ti < 0 || fi < 0 || w.nonce != before.withdraw_nonce + 1) continue;
        if (before.tokens.at(ti).amount < w.amount) continue;
        Fr sig_s;
        memcpy(sig_s.l, &w.sig_s, 32);
        if (!eddsa_verify_with_h(s, addr[k], Point{fr_from_canon(&w.sig_rx), fr_from_canon(&w.sig_ry)}, sig_s, hs[k])) continue;
        Plan p{};
        p.k = k; p.idx = idx; p.ti = (uint32_t)ti; p.fi = (uint32_t)fi; p.before = before; p.tok = before.tokens.at(ti);
        p.mid = before;
        p.mid.tokens[ti].amount -= w.amount;
        if (p.mid.tokens.at(fi).amount < w.fee) continue;
        p.fee_before = p.mid.tokens.at(fi);
        p.after = p.mid;
        p.after.tokens[fi].amount -= w.fee;
        p.after.withdraw_nonce += 1;
        mirror[idx] = p.after;
        plan.push_back(std::move(p));
        if (accepted) accepted[k] = 1;
    }
    Forest forest;
    forest.T = T;
    std::vector<uint64_t> touched;
    for (auto &p : plan)
        if (forest.tree_of.emplace(p.idx, (uint32_t)forest.tree_of.size()).second) touched.push_back(p.idx);
    for (uint64_t acc : touched) {
        auto it = s->accounts.find(acc);
        if (it != s->accounts.end())
            for (auto &kv : it->second.tokens) forest.write(acc, kv.first, kv.second);
    }
    forest.n_init = forest.idx.size();
    for (auto &p : plan) {
        p.e1 = forest.write(p.idx, p.ti, p.mid.tokens[p.ti]);
        p.e2 = forest.write(p.idx, p.fi, p.after.tokens[p.fi]);
    }
    BZK_TRY(forest.run(ctx, s->tdefaults));
    std::vector<Fr> acct_rows, cd_rows;
    auto push_acct = [&](const Account &a, const Fr &tok_root) {
        acct_rows.push_back(fr_from_u64(a.tx_nonce)); acct_rows.push_back(fr_from_u64(a.withdraw_nonce));
        acct_rows.push_back(a.ax); acct_rows.push_back(a.ay); acct_rows.push_back(tok_root);
    };
    for (auto &p : plan) {
        p.tok_hash = forest.root(p.idx);
        const Fr r1 = forest.applied(p.idx, p.e1), r2 = forest.applied(p.idx, p.e2);
        push_acct(p.mid, r1); push_acct(p.after, r2);
        const bzk_mpn_withdraw &w = wds[p.k];
        cd_rows.push_back(addr[p.k].x); cd_rows.push_back(addr[p.k].y); cd_rows.push_back(fr_from_u64(w.nonce));
        cd_rows.push_back(fr_from_canon(&w.sig_rx)); cd_rows.push_back(fr_from_canon(&w.sig_ry)); cd_rows.push_back(fr_from_canon(&w.sig_s));
    }
    std::vector<Fr> s_vals, s_proofs, cds;
    BZK_TRY(hash_rows(ctx, 5, acct_rows, s_vals));
    BZK_TRY(hash_rows(ctx, 6, cd_rows, cds));
    std::vector<uint64_t> s_idx;
    for (auto &p : plan) { s_idx.push_back(p.idx); s_idx.push_back(p.idx); }
    const size_t ne = s_idx.size();
    std::vector<Fr> init(ne * A * 3);
    for (size_t e = 0; e < ne; e++) s->prove(s_idx[e], init.data() + e * A * 3);
    BZK_TRY(tree_update_host(ctx, A, std::vector<uint32_t>(ne, 0u), s_idx, s_vals, init, s_proofs));
    const Fr minus_one = Fr::one().neg();
    Fr root = prev_root;
    std::vector<Fr> rev_rows(cap * 7, Fr::zero());
    for (uint64_t slot = 0; slot < cap; slot++) {
        bzk_fr *r1 = raws1 + slot * 12, *r2 = raws2 + slot * w2, *rv = reveal + slot * 7;
        memset(r1, 0, 12 * sizeof(bzk_fr)); memset(r2, 0, (size_t)w2 * sizeof(bzk_fr)); memset(rv, 0, 7 * sizeof(bzk_fr));
        if (slot >= plan.size()) {
            fr_to_canon(r1 + 7, minus_one);
            continue;
        }
        const Plan &p = plan[slot];
        const bzk_mpn_withdraw &w = wds[p.k];
        fr_to_canon(roots + slot, root);
        const uint64_t one = 1, nonce = w.nonce;
        memcpy(r1 + 0, &one, 8); r1[1] = w.amount_token_id; memcpy(r1 + 2, &w.amount, 8); r1[3] = w.fee_token_id; memcpy(r1 + 4, &w.fee, 8);
        r1[5] = w.fingerprint; fr_to_canon(r1 + 6, addr[p.k].x); fr_to_canon(r1 + 7, addr[p.k].y); memcpy(r1 + 8, &nonce, 8);
        r1[9] = w.sig_rx; r1[10] = w.sig_ry; r1[11] = w.sig_s;
        size_t q = 0;
        auto put_fr = [&](const Fr &v) { fr_to_canon(r2 + (q++), v); };
        auto put_u = [&](uint64_t v) { memcpy(r2 + (q++), &v, 8); };
        put_u(p.idx); put_u(p.ti); put_u(p.fi); put_u(p.before.tx_nonce); put_u(p.before.withdraw_nonce); put_fr(p.before.ax); put_fr(p.before.ay);
        put_fr(p.tok_hash); put_fr(p.tok.token_id); put_u(p.tok.amount);
        for (uint32_t i = 0; i < T * 3; i++) put_fr(forest.proofs[p.e1 * T * 3 + i]);
        put_fr(p.fee_before.token_id); put_u(p.fee_before.amount);
        for (uint32_t i = 0; i < T * 3; i++) put_fr(forest.proofs[p.e2 * T * 3 + i]);
        for (uint32_t i = 0; i < A * 3; i++) put_fr(s_proofs[(2 * slot) * A * 3 + i]);
        if (q != w2) return BZK_ERR_BAD_ARG;
        memcpy(rv + 0, &one, 8); rv[1] = w.amount_token_id; memcpy(rv + 2, &w.amount, 8); rv[3] = w.fee_token_id; memcpy(rv + 4, &w.fee, 8);
        rv[5] = w.fingerprint; fr_to_canon(rv + 6, cds[slot]);
        Fr *rr = rev_rows.data() + slot * 7;
        rr[0] = Fr::one(); rr[1] = fr_from_canon(&w.amount_token_id); rr[2] = fr_from_u64(w.amount); rr[3] = fr_from_canon(&w.fee_token_id);
        rr[4] = fr_from_u64(w.fee); rr[5] = fr_from_canon(&w.fingerprint); rr[6] = cds[slot];
        root = s_vals[(size_t)A * ne + 2 * slot + 1];
    }
    for (uint64_t slot = plan.size(); slot < cap; slot++) fr_to_canon(roots + slot, root);
    for (size_t e = 0; e < ne; e++) {
        uint64_t node = s_idx[e];
        for (uint32_t l = 0; l <= A; l++) { s->put(l, node, s_vals[(size_t)l * ne + e]); node >>= 2; }
    }
    for (uint64_t i : touched) {
        auto it = s->accounts.find(i);
        if (it != s->accounts.end()) s->state_size -= leaf_count(it->second);
        s->state_size += leaf_count(mirror[i]);
        s->accounts[i] = mirror[i];
    }
    Fr aux;
    BZK_TRY(list_root(ctx, 7, rev_rows, &aux));
    fr_to_canon(public3 + 0, prev_root);
    fr_to_canon(public3 + 1, aux);
    fr_to_canon(public3 + 2, root);
    *n_accepted = plan.size();
    if (sink) {   // `WithdrawTransition` of every accepted withdrawal (/root/reference/src/mpn/withdraw.rs:160-178)
        sink->t.clear(); sink->from.clear();
        for (size_t slot = 0; slot < plan.size(); slot++) {
            const Plan &p = plan[slot];
            wire::WithdrawTransition t;
            t.enabled = true;
            t.before = wire_account(p.before); t.before_token_balance = wire_money(p.tok); t.before_fee_balance = wire_money(p.fee_before);
            t.proof = wire_proof(s_proofs.data() + (2 * slot) * A * 3, A);
            t.account_index = p.idx; t.token_index = p.ti; t.token_balance_proof = wire_proof(forest.proofs.data() + p.e1 * T * 3, T);
            t.before_token_hash = p.tok_hash; t.fee_token_index = p.fi; t.fee_balance_proof = wire_proof(forest.proofs.data() + p.e2 * T * 3, T);
            sink->t.push_back(std::move(t));
            sink->from.push_back(p.k);
        }
    }
    return BZK_OK;
}

// ---------------------------------------------------------------------------------------------
// witness of a whole deposit / withdraw batch from the builder's rows: what `{Deposit,Withdraw}Circuit::synthesize`
// assigns (/root/reference/src/mpn/circuits/deposit_circuit.rs:47-293, withdraw_circuit.rs:49-413), laid out as
//   inputs = [1, commitment, height, state, aux_data, next_state]
//   aux    = [the five public values] ++ phase-1 program x n_slots ++ reveal program ++ phase-2 program x n_slots
// phase 2's externals per slot: ext_src[e] < 0 -> the state root entering the slot, else phase-1 raw ext_src[e] of the slot
// (bzk_mpn_circuit_two_phase_info); the reveal program's externals are the builder's reveal rows, slot-major.
// ---------------------------------------------------------------------------------------------
extern "C" int32_t bzk_mpn_dw_witness(bzk_ctx *ctx, const bzk_witness_program *phase1, const bzk_witness_program *phase2,
                                      const bzk_witness_program *reveal_prog, uint64_t n_slots, const bzk_fr *raws1, const bzk_fr *raws2,
                                      const bzk_fr *roots, const int32_t *ext_src, uint32_t n_ext_src, const bzk_fr *reveal_rows,
                                      const bzk_fr public5[5], void *d_inputs, void *d_aux) {
    if (!ctx || !phase1 || !phase2 || !reveal_prog || !n_slots || !raws1 || !raws2 || !roots || (n_ext_src && !ext_src) || !reveal_rows || !public5 ||
        !d_inputs || !d_aux)
        return BZK_ERR_BAD_ARG;
    uint64_t n1 = 0, n2 = 0, nr = 0;
    uint32_t raw1 = 0, ext1 = 0, raw2 = 0, ext2 = 0, rawr = 0, extr = 0;
    witness_program_shape(phase1, &n1, &raw1, &ext1);
    witness_program_shape(phase2, &n2, &raw2, &ext2);
    witness_program_shape(reveal_prog, &nr, &rawr, &extr);
    if (ext1 != 0 || ext2 != n_ext_src || extr % n_slots != 0) return BZK_ERR_BAD_ARG;
    for (uint32_t e = 0; e < n_ext_src; e++)
        if (ext_src[e] >= (int32_t)raw1) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr head[11];
    head[0] = Fr::one();
    for (int k = 0; k < 5; k++) { head[1 + k] = fr_from_canon(public5 + k); head[6 + k] = head[1 + k]; }
    Fr *z_in = (Fr *)d_inputs, *z_aux = (Fr *)d_aux;
    BZK_CUDA(ctx, cudaMemcpyAsync(z_in, head, 6 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    BZK_CUDA(ctx, cudaMemcpyAsync(z_aux, head + 6, 5 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    BZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `head` is a stack buffer
    BZK_TRY(bzk_witness_run_dev(ctx, phase1, raws1, nullptr, n_slots, z_aux + 5));
    BZK_TRY(bzk_witness_run_dev(ctx, reveal_prog, nullptr, reveal_rows, 1, z_aux + 5 + n_slots * n1));
    std::vector<bzk_fr> ext((size_t)n_slots * n_ext_src);
    for (uint64_t k = 0; k < n_slots; k++)
        for (uint32_t e = 0; e < n_ext_src; e++) ext[k * n_ext_src + e] = ext_src[e] < 0 ? roots[k] : raws1[k * raw1 + ext_src[e]];
    BZK_TRY(bzk_witness_run_dev(ctx, phase2, raws2, ext.data(), n_slots, z_aux + 5 + n_slots * n1 + nr));
    return BZK_OK;
}

extern "C" {
}  // extern "C"

// ---------------------------------------------------------------------------------------------
// witness of a whole update batch from the builder's rows: what `UpdateCircuit::synthesize` assigns
// (/root/reference/src/mpn/circuits/update_circuit.rs:49-494), laid out as z = inputs ++ aux:
//   inputs  = [1, commitment, height, state, aux_data, next_state]
//   aux     = [commitment, height, state, fee_token, aux_data, next_state]      (the six prologue allocations)
//             ++ slot program x n_slots                                          (bzk_witness_run_dev)
//             ++ epilogue program (the Poseidon(fee_token, sum of accepted fees) gadget; externals =
//                fee_token, then every slot's accepted fee)
// ---------------------------------------------------------------------------------------------
extern "C" int32_t bzk_mpn_update_witness(bzk_ctx *ctx, const bzk_witness_program *slot_prog, const bzk_witness_program *epilogue_prog,
                                          uint64_t n_slots, uint32_t log4_token, uint64_t slot_vars, uint64_t epilogue_vars, const bzk_fr *raws,
                                          const bzk_fr *ext, uint32_t n_raw, const bzk_fr prologue[6], void *d_inputs, void *d_aux) {
    if (!ctx || !slot_prog || !epilogue_prog || !n_slots || !raws || !ext || !prologue || !d_inputs || !d_aux || n_raw < 16 + 3 * log4_token)
        return BZK_ERR_BAD_ARG;
    {   // the rows must have the shape the two programs were compiled for (a mismatch would read / write out of bounds)
        uint64_t s_ops = 0, e_ops = 0;
        uint32_t s_raw = 0, s_ext = 0, e_raw = 0, e_ext = 0;
        witness_program_shape(slot_prog, &s_ops, &s_raw, &s_ext);
        witness_program_shape(epilogue_prog, &e_ops, &e_raw, &e_ext);
        if (s_raw != n_raw || s_ext != 2 || s_ops != slot_vars || e_ext != 1 + n_slots || e_ops != epilogue_vars) return BZK_ERR_BAD_ARG;
    }
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr head[12];  // 6 inputs, 6 prologue aux (Montgomery)
    const Fr commitment = fr_from_canon(prologue + 0), height = fr_from_canon(prologue + 1), state = fr_from_canon(prologue + 2),
             fee_token = fr_from_canon(prologue + 3), aux_data = fr_from_canon(prologue + 4), next_state = fr_from_canon(prologue + 5);
    head[0] = Fr::one(); head[1] = commitment; head[2] = height; head[3] = state; head[4] = aux_data; head[5] = next_state;
    head[6] = commitment; head[7] = height; head[8] = state; head[9] = fee_token; head[10] = aux_data; head[11] = next_state;
    Fr *z_in = (Fr *)d_inputs, *z_aux = (Fr *)d_aux;
    BZK_CUDA(ctx, cudaMemcpyAsync(z_in, head, 6 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    BZK_CUDA(ctx, cudaMemcpyAsync(z_aux, head + 6, 6 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    BZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `head` is a stack buffer
    BZK_TRY(bzk_witness_run_dev(ctx, slot_prog, raws, ext, n_slots, z_aux + 6));
    // accepted fee of a slot = enabled ? tx.fee : 0; both are raw inputs (positions 0 and 15 + 3T of a row)
    std::vector<bzk_fr> epi_ext(1 + n_slots);
    epi_ext[0] = prologue[3];
    const bzk_fr zero{};
    for (uint64_t k = 0; k < n_slots; k++) {
        const bzk_fr *row = raws + k * n_raw;
        epi_ext[1 + k] = row[0].l[0] ? row[15 + 3 * log4_token] : zero;
    }
    BZK_TRY(bzk_witness_run_dev(ctx, epilogue_prog, nullptr, epi_ext.data(), 1, z_aux + 6 + n_slots * slot_vars));
    return BZK_OK;
}

// `PublicKey::decompress` as a stand-alone host call (no context): x canonical, y parity flag -> affine point
// (canonical); BZK_ERR_NOT_ON_CURVE when x is not the abscissa of a curve point.
extern "C" int32_t bzk_jubjub_decompress(const bzk_fr *jubjub_d, const bzk_fr *x, int32_t y_is_odd, bzk_fr out_xy[2]) {
    if (!jubjub_d || !x || !out_xy) return BZK_ERR_BAD_ARG;
    bzk_mpn_state tmp;
    tmp.jj_d = fr_from_canon(jubjub_d);
    Point p;
    if (!jj_decompress(&tmp, x, y_is_odd != 0, &p) || !tmp.on_curve(p.x, p.y)) return BZK_ERR_NOT_ON_CURVE;
    fr_to_canon(out_xy + 0, p.x);
    fr_to_canon(out_xy + 1, p.y);
    return BZK_OK;
}

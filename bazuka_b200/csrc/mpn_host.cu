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
#include <algorithm>
#include <map>
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

// Fr square root (Tonelli-Shanks, r - 1 = 2^32 * odd, 7 = non-residue); false when none exists
bool fr_sqrt(const Fr &a, Fr *out) {
    if (a.is_zero()) { *out = a; return true; }
    // q = (r - 1) >> 32, as 32-bit words (224 bits)
    uint32_t q[8] = {0}, rm1[8];
    for (int i = 0; i < 8; i++) rm1[i] = FrParams::p(i);
    rm1[0] -= 1;
    for (int i = 0; i < 7; i++) q[i] = rm1[i + 1];
    uint32_t half[8];  // (r - 1) / 2
    for (int i = 0; i < 8; i++) half[i] = (rm1[i] >> 1) | (i < 7 ? rm1[i + 1] << 31 : 0);
    if (!(a.pow(half, 8) == Fr::one())) return false;
    uint32_t q1[8];  // (q + 1) / 2
    {
        uint64_t c = 1;
        uint32_t t[8];
        for (int i = 0; i < 8; i++) { c += q[i]; t[i] = (uint32_t)c; c >>= 32; }
        for (int i = 0; i < 8; i++) q1[i] = (t[i] >> 1) | (i < 7 ? t[i + 1] << 31 : 0);
    }
    uint32_t m = 32;
    Fr c = Fr::from_u32(7).pow(q, 8), t = a.pow(q, 8), r = a.pow(q1, 8);
    while (!(t == Fr::one())) {
        uint32_t i = 0;
        Fr t2 = t;
        while (!(t2 == Fr::one())) { t2 = t2 * t2; i++; }
        Fr b = c;
        for (uint32_t k = 0; k + i + 1 < m; k++) b = b * b;
        m = i;
        c = b * b;
        t = t * c;
        r = r * b;
    }
    *out = r;
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
        if (!fr_sqrt((Fr::one() + x2) * den.inv(), &y)) return false;  // a = -1
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
    return BZK_OK;
}

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

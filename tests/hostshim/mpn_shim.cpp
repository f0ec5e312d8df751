// TEST INFRASTRUCTURE (CPU tier only) — never linked into libbzk.so, never loaded by the product.
//
// The HOST parts of libbzk — the MPN ledger and the three transition builders (csrc/mpn_host.cu), the wire codec and
// `prepare_works` (csrc/mpn_wire.cu), the host Poseidon (csrc/poseidon_host.cu) — are compiled UNMODIFIED with g++ into
// tests/hostshim/_mpn_shim.so, and this file stands in for everything they reach on the GPU side:
//   * the handful of CUDA runtime calls they make ("device" memory = host memory, streams are synchronous);
//   * the batched Poseidon launch (bzk_poseidon_hash)            -> the host Poseidon, one hash after the other;
//   * the versioned level-synchronous tree update kernel          -> the same rule, written as host loops over (level, write);
//   * the witness interpreter launch (bzk_witness_run_dev)        -> witness_core.cuh's per-slot loop, the text the kernel runs.
// So the "not gpu" tier can drive the real native ledger logic (acceptance rules, index tables, row assembly, bincode) against
// the Python restatement of the reference; the `-m gpu` tier runs the same logic over the real kernels.
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "witness_core.cuh"

using namespace bzk;

// ---------------------------------------------------------------- CUDA runtime stand-ins
extern "C" {
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) {
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
    free(p);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) {
    memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "fake CUDA runtime (tests/hostshim)"; }
}

// ---------------------------------------------------------------- context + hasher
static bzk_poseidon_host *g_hasher = nullptr;

extern "C" {
int32_t shim_set_poseidon(const uint8_t *blob, size_t len) {
    if (g_hasher) bzk_poseidon_host_free(g_hasher);
    g_hasher = nullptr;
    return bzk_poseidon_host_create(blob, len, &g_hasher);
}
bzk_ctx *shim_ctx_create() {
    auto *c = new bzk_ctx();
    c->pos_loaded = true;
    return c;
}
void shim_ctx_free(bzk_ctx *c) {
    if (!c) return;
    free(c->ws);
    delete c;
}

// the batched launch of csrc/poseidon.cu, host buffers in and out (Montgomery images)
int32_t bzk_poseidon_hash(bzk_ctx *ctx, uint32_t arity, const bzk_fr *in, size_t n, bzk_fr *out) {
    if (!ctx || !g_hasher) return BZK_ERR_NO_PARAMS;
    if (ctx) ctx->launches++;
    return bzk_poseidon_host_hash(g_hasher, arity, in, n, out);
}
}

namespace bzk {
// csrc/poseidon.cu k_tree4_versioned_level, as loops: for every level, for every write e, the node of e's path is
// H(children) with the own-path child from the level below and each sibling taken from the LATEST EARLIER write of the same
// tree whose path runs through it, else from the pre-batch proof; out_proofs[e] = the siblings seen, vals[lvl+1][e] = the node.
int32_t tree4_versioned_update(bzk_ctx *ctx, uint32_t depth, const uint32_t *tree_id, const uint64_t *idx, size_t n, Fr *vals, const Fr *init_proofs,
                               Fr *out_proofs) {
    if (!g_hasher) return BZK_ERR_NO_PARAMS;
    if (depth == 0 || depth > 32 || (n && (!tree_id || !idx || !vals || !init_proofs || !out_proofs))) return BZK_ERR_BAD_ARG;
    for (uint32_t lvl = 0; lvl < depth; lvl++) {
        const Fr *cur = vals + (size_t)lvl * n;
        const uint32_t up = 2 * lvl + 2;
        for (size_t e = 0; e < n; e++) {
            const uint64_t me = idx[e], prefix = up >= 64 ? 0 : me >> up;
            const uint32_t pos = (uint32_t)((me >> (2 * lvl)) & 3);
            Fr s[4];
            const Fr *sib = init_proofs + (e * depth + lvl) * 3;
            int w = 0;
            for (uint32_t k = 0; k < 4; k++) s[k] = (k == pos) ? cur[e] : sib[w++];
            uint32_t found = 1u << pos;
            for (size_t b = e; b-- > 0 && found != 15u;) {
                if (tree_id[b] != tree_id[e] || (up >= 64 ? 0 : idx[b] >> up) != prefix) continue;
                const uint32_t k = (uint32_t)((idx[b] >> (2 * lvl)) & 3);
                if (found & (1u << k)) continue;
                found |= 1u << k;
                s[k] = cur[b];
            }
            Fr *po = out_proofs + (e * depth + lvl) * 3;
            w = 0;
            for (uint32_t k = 0; k < 4; k++)
                if (k != pos) po[w++] = s[k];
            BZK_TRY(bzk_poseidon_host_hash(g_hasher, 4, (const bzk_fr *)s, 1, (bzk_fr *)(vals + (size_t)(lvl + 1) * n + e)));
        }
        if (ctx) ctx->launches++;
    }
    return BZK_OK;
}
}  // namespace bzk

// ---------------------------------------------------------------- witness interpreter on the host
struct bzk_witness_program {
    std::vector<int32_t> ops, lc_ptr, lc_slot, lc_coef;
    std::vector<Fr> coefs;
    uint32_t n_raw = 0, n_ext = 0;
    Fr jj_d;
};

namespace bzk {
void witness_program_shape(const bzk_witness_program *p, uint64_t *n_ops, uint32_t *n_raw, uint32_t *n_ext) {
    *n_ops = p->ops.size() / 6; *n_raw = p->n_raw; *n_ext = p->n_ext;
}
}  // namespace bzk

extern "C" {
int32_t bzk_witness_program_upload(bzk_ctx *ctx, const int32_t *ops, uint64_t n_ops, const int32_t *lc_ptr, uint64_t n_lc, const int32_t *lc_slot,
                                   const int32_t *lc_coef, uint64_t n_terms, const bzk_fr *coefs, uint64_t n_coefs, uint32_t n_raw, uint32_t n_ext,
                                   const bzk_fr *jj_d, bzk_witness_program **out) {
    if (!ctx || !ops || !lc_ptr || !coefs || !jj_d || !out || !n_ops || !n_coefs) return BZK_ERR_BAD_ARG;
    auto *p = new bzk_witness_program;
    p->ops.assign(ops, ops + n_ops * 6);
    p->lc_ptr.assign(lc_ptr, lc_ptr + n_lc + 1);
    if (n_terms) { p->lc_slot.assign(lc_slot, lc_slot + n_terms); p->lc_coef.assign(lc_coef, lc_coef + n_terms); }
    p->coefs.resize(n_coefs);
    memcpy(p->coefs.data(), coefs, n_coefs * sizeof(Fr));
    p->n_raw = n_raw; p->n_ext = n_ext;
    memcpy(&p->jj_d, jj_d, sizeof(Fr));
    *out = p;
    return BZK_OK;
}
int32_t bzk_witness_program_free(bzk_ctx *, bzk_witness_program *p) {
    delete p;
    return BZK_OK;
}
// slot-major output like the kernel's: aux_out[slot * n_ops + j]
int32_t bzk_witness_run_dev(bzk_ctx *ctx, const bzk_witness_program *p, const bzk_fr *raws, const bzk_fr *ext, uint64_t ntx, void *aux_out) {
    if (!ctx || !p || (p->n_raw && !raws) || (p->n_ext && !ext) || !aux_out || !ntx) return BZK_ERR_BAD_ARG;
    const uint32_t n_ops = (uint32_t)(p->ops.size() / 6);
    struct Mem {
        std::vector<Fr> V;
        Fr *out_;
        Fr load(int32_t slot) const { return V[slot]; }
        void store(uint32_t slot, const Fr &v) { V[slot] = v; }
        void out(uint32_t j, const Fr &v) { out_[j] = v; }
        void prefetch(int32_t) const {}
    } mem;
    WitProgDev P{p->ops.data(), p->lc_ptr.data(), p->lc_slot.data(), p->lc_coef.data(), p->coefs.data(), n_ops, p->n_raw, p->n_ext};
    for (uint64_t k = 0; k < ntx; k++) {
        mem.V.assign((size_t)1 + p->n_ext + n_ops, Fr::zero());
        mem.out_ = (Fr *)aux_out + k * n_ops;
        wit_run_slot(P, p->jj_d, (const Fr *)raws + k * p->n_raw, (const Fr *)ext + k * p->n_ext, mem);
    }
    ctx->launches++;
    return BZK_OK;
}
}

// ---------------------------------------------------------------- the prover's resident R1CS and the prove call
// bzk_mpn_prover_{create,prove_work} (csrc/mpn_prover.cu) end in bzk_r1cs_upload / bzk_groth16_prove_dev.  Here the "upload"
// keeps the CSR on the host and the "prove" checks what the real call checks first — a(z) * b(z) == c(z) on every constraint of
// the natively compiled circuit, for the z the native witness drivers just wrote — and returns the identity points: the CPU tier
// proves the COMPOSITION (work bytes -> rows -> witness -> a satisfying assignment with the right public inputs); the MSM / NTT
// half of the call is the GPU tier's.
struct bzk_r1cs {
    uint64_t ni = 0, na = 0, ncons = 0;
    std::vector<uint64_t> rp[3];
    std::vector<uint32_t> col[3];
    std::vector<Fr> val[3];
};
static uint64_t g_last_unsat_row = ~0ull;
static const Fr *g_last_in = nullptr, *g_last_aux = nullptr;   // the z of the last prove call (still resident in its prover)
static uint64_t g_last_ni = 0, g_last_na = 0;
extern "C" {
int32_t bzk_r1cs_upload(bzk_ctx *ctx, uint64_t num_inputs, uint64_t num_aux, uint64_t num_constraints, const uint64_t *a_rowptr, const uint32_t *a_col,
                        const bzk_fr *a_val, const uint64_t *b_rowptr, const uint32_t *b_col, const bzk_fr *b_val, const uint64_t *c_rowptr,
                        const uint32_t *c_col, const bzk_fr *c_val, bzk_r1cs **out) {
    if (!ctx || !out) return BZK_ERR_BAD_ARG;
    auto *r = new bzk_r1cs;
    r->ni = num_inputs; r->na = num_aux; r->ncons = num_constraints;
    const uint64_t *rp[3] = {a_rowptr, b_rowptr, c_rowptr};
    const uint32_t *col[3] = {a_col, b_col, c_col};
    const bzk_fr *val[3] = {a_val, b_val, c_val};
    for (int s = 0; s < 3; s++) {
        r->rp[s].assign(rp[s], rp[s] + num_constraints + 1);
        const uint64_t nnz = rp[s][num_constraints];
        r->col[s].assign(col[s], col[s] + nnz);
        r->val[s].resize(nnz);
        memcpy(r->val[s].data(), val[s], nnz * sizeof(Fr));
    }
    *out = r;
    return BZK_OK;
}
int32_t bzk_r1cs_free(bzk_ctx *, bzk_r1cs *r) {
    delete r;
    return BZK_OK;
}
int32_t bzk_groth16_prove_dev(bzk_ctx *ctx, const bzk_groth16_params *, const bzk_r1cs *r, const void *d_inputs, const void *d_aux, const bzk_fr *,
                              const bzk_fr *, int32_t check_satisfied, bzk_g1_affine *pa, bzk_g2_affine *pb, bzk_g1_affine *pc) {
    if (!ctx || !r || !d_inputs || !d_aux || !pa || !pb || !pc) return BZK_ERR_BAD_ARG;
    const Fr *zi = (const Fr *)d_inputs, *za = (const Fr *)d_aux;
    auto z = [&](uint32_t c) { return c < r->ni ? zi[c] : za[c - r->ni]; };
    g_last_unsat_row = ~0ull;
    g_last_in = zi; g_last_aux = za; g_last_ni = r->ni; g_last_na = r->na;
    if (!(zi[0] == Fr::one())) { g_last_unsat_row = 0; if (check_satisfied) return BZK_ERR_UNSAT; }
    for (uint64_t row = 0; row < r->ncons; row++) {
        Fr v[3];
        for (int s = 0; s < 3; s++) {
            Fr acc = Fr::zero();
            for (uint64_t k = r->rp[s][row]; k < r->rp[s][row + 1]; k++) acc = acc + r->val[s][k] * z(r->col[s][k]);
            v[s] = acc;
        }
        if (!(v[0] * v[1] == v[2])) {
            g_last_unsat_row = row;
            if (check_satisfied) return BZK_ERR_UNSAT;
            break;
        }
    }
    memset(pa, 0, sizeof *pa); memset(pb, 0, sizeof *pb); memset(pc, 0, sizeof *pc);
    pa->infinity = pb->infinity = pc->infinity = 1;
    return BZK_OK;
}
uint64_t shim_last_unsat_row() { return g_last_unsat_row; }
// copies the assignment the last prove call was given (valid while its prover lives): inputs[ni], aux[na], Montgomery
int32_t shim_last_z(void *inputs, uint64_t ni, void *aux, uint64_t na) {
    if (!g_last_in || ni != g_last_ni || na != g_last_na) return BZK_ERR_BAD_ARG;
    memcpy(inputs, g_last_in, ni * sizeof(Fr));
    memcpy(aux, g_last_aux, na * sizeof(Fr));
    return BZK_OK;
}
// the z the last prover call left resident (the shim's "device" memory is host memory): copies n elements from a raw pointer
void shim_peek(const void *p, void *out, size_t bytes) { memcpy(out, p, bytes); }
}

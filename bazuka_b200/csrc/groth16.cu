// bazuka_b200 — Groth16 prover driver: R1CS evaluation, quotient polynomial, the five MSMs and the
// (r, s) blinding tail.
//
// GPU replacement for bellman 0.14.0 `groth16::prover::create_proof` (un-vendored crate; the
// reference reaches it from /root/reference/src/mpn/circuits/test.rs:135,175,215 and every gadget
// test; in production the call sits in the external prover that answers `MpnWork`,
// /root/reference/src/mpn/mod.rs:264-295).  Same dataflow as bellman:
//   a,b,c = <A_j,z>, <B_j,z>, <C_j,z> per constraint (+ the appended `Input(i) * 0 = 0` rows)
//   h     = first m-1 coefficients of icoset_fft((coset_fft(ifft a) * coset_fft(ifft b)
//           - coset_fft(ifft c)) / Z)
//   sums  : h*H, aux*L, [inputs ++ aux|A-density]*A, [inputs|B-density ++ aux|B-density]*B1, same*B2
//   A = alpha + r delta + a_sum ; B = beta + s delta + b2_sum ;
//   C = s a_sum + r b1_sum + r s delta + s alpha + r beta + h_sum + l_sum
// What is different: the constraint system is a device-resident CSR triple evaluated by an SpMV
// kernel instead of re-synthesising the circuit per proof; density trackers become index lists
// built once at upload; all vectors stay in HBM between stages.
#include "common.cuh"
#include <algorithm>
#include <chrono>
#include <future>
#include <cstdlib>

namespace bzk {
int32_t precompute_g1(bzk_ctx *ctx, bzk_g1_bases *b, uint32_t max_levels);
int32_t precompute_g2(bzk_ctx *ctx, bzk_g2_bases *b, uint32_t max_levels);
int32_t groth16_h_launch(bzk_ctx *ctx, Fr *a, Fr *b, Fr *c, uint32_t log_n);
int32_t groth16_to_coset_launch(bzk_ctx *ctx, Fr *v, uint32_t log_n);
int32_t msm_g1_enqueue(bzk_ctx *ctx, cudaStream_t st, void **ws, size_t *ws_bytes, const BasesRef<Fp> &d_bases, const Fr *d_scalars, size_t n, void *h_win, MsmPlan *plan);
int32_t msm_g2_enqueue(bzk_ctx *ctx, cudaStream_t st, void **ws, size_t *ws_bytes, const BasesRef<Fp2> &d_bases, const Fr *d_scalars, size_t n, void *h_win, MsmPlan *plan);
void msm_g1_finish(const MsmPlan *plan, const void *h_win, bzk_g1_affine *out);
void msm_g2_finish(const MsmPlan *plan, const void *h_win, bzk_g2_affine *out);

struct DevCsr {
    uint64_t *rowptr = nullptr;
    uint32_t *col = nullptr;
    Fr *val = nullptr;
    uint64_t nnz = 0;
};
}  // namespace bzk

struct bzk_r1cs {
    uint64_t num_inputs = 0, num_aux = 0, ncons = 0;
    uint32_t log_m = 0;
    bzk::DevCsr m[3];
    uint32_t *d_a_idx = nullptr, *d_b_idx = nullptr;  // indices into z for the A / B sums
    uint64_t a_len = 0, b_len = 0;
};

struct bzk_groth16_params {
    bzk::G1Affine alpha_g1, beta_g1, delta_g1;
    bzk::G2Affine beta_g2, delta_g2;
    bzk_g1_bases *h = nullptr, *l = nullptr, *a = nullptr, *b1 = nullptr;
    bzk_g2_bases *b2 = nullptr;
    // base sharding: this handle holds the contiguous range
    // [len*rank/world, len*(rank+1)/world) of each of the five base vectors
    uint32_t rank = 0, world = 1;
};

struct Groth16Partials {  // the four sums a proof is assembled from (wire images)
    bzk_g1_affine *a_sum, *b1_sum, *hl_sum;
    bzk_g2_affine *b2_sum;
};

// sharded schedule with the quotient split over the ranks: the call is cut in two around the exchange of polynomials
struct Groth16Split {
    int phase;              // 1 = begin (z, this rank's share of the evaluation vectors on the coset, the four witness sums enqueued)
                            // 2 = finish (h sum over this rank's slice of the quotient, folds)
    uint32_t poly_mask;     // begin: bit s set = this rank transforms evaluation vector s (a, b, c)
    bzk::Fr *evals[3];      // begin: caller's device buffers (domain size) for those vectors
    const bzk::Fr *h_shard; // finish: the rank's slice [lo, hi) of the quotient's coefficients (device)
};

namespace bzk {

// one thread per constraint row: out[row] = sum_k val[k] * z[col[k]]
__global__ void __launch_bounds__(256) k_csr_spmv(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                                  const Fr *__restrict__ val, uint64_t nrows, const Fr *__restrict__ z,
                                                  Fr *__restrict__ out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    Fr acc = Fr::zero();
    const uint64_t k1 = rowptr[r + 1];
    for (uint64_t k = rowptr[r]; k < k1; k++) acc = acc + load_vec(val + k) * load_vec(z + col[k]);
    store_vec(out + r, acc);
}
__global__ void __launch_bounds__(256) k_gather_fr(const Fr *__restrict__ z, const uint32_t *__restrict__ idx, uint64_t n, Fr *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_vec(out + i, load_vec(z + idx[i]));
}
// count rows with a*b != c
__global__ void __launch_bounds__(256) k_check_sat(const Fr *__restrict__ a, const Fr *__restrict__ b, const Fr *__restrict__ c, uint64_t n, uint32_t *bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (load_vec(a + i) * load_vec(b + i) != load_vec(c + i)) atomicAdd(bad, 1u);
}
__global__ void __launch_bounds__(128) k_fixed_base_g1(G1Affine base, const Fr *__restrict__ k_mont, size_t n, uint8_t *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr k = load_vec(k_mont + i).from_mont();
    store_g1_image(out + i * 104, scalar_mul(base, k.l).to_affine());
}
__global__ void __launch_bounds__(64) k_fixed_base_g2(G2Affine base, const Fr *__restrict__ k_mont, size_t n, uint8_t *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr k = load_vec(k_mont + i).from_mont();
    store_g2_image(out + i * 200, scalar_mul(base, k.l).to_affine());
}

static G1Affine g1_from_img(const bzk_g1_affine *img) {
    if (img->infinity) return G1Affine::inf();
    G1Affine p;
    memcpy(p.x.l, img->x, 48);
    memcpy(p.y.l, img->y, 48);
    return p;
}
static G2Affine g2_from_img(const bzk_g2_affine *img) {
    if (img->infinity) return G2Affine::inf();
    G2Affine p;
    memcpy(p.x.c0.l, img->x, 48); memcpy(p.x.c1.l, img->x + 6, 48);
    memcpy(p.y.c0.l, img->y, 48); memcpy(p.y.c1.l, img->y + 6, 48);
    return p;
}
static void g1_to_img(bzk_g1_affine *img, const G1Affine &p) {
    memset(img, 0, sizeof *img);
    if (p.is_inf()) { Fp one = Fp::one(); memcpy(img->y, one.l, 48); img->infinity = 1; return; }
    memcpy(img->x, p.x.l, 48);
    memcpy(img->y, p.y.l, 48);
}
static void g2_to_img(bzk_g2_affine *img, const G2Affine &p) {
    memset(img, 0, sizeof *img);
    if (p.is_inf()) { Fp one = Fp::one(); memcpy(img->y, one.l, 48); img->infinity = 1; return; }
    memcpy(img->x, p.x.c0.l, 48); memcpy(img->x + 6, p.x.c1.l, 48);
    memcpy(img->y, p.y.c0.l, 48); memcpy(img->y + 6, p.y.c1.l, 48);
}

static int32_t upload_csr(bzk_ctx *ctx, DevCsr &d, uint64_t nrows, const uint64_t *rp, const uint32_t *col, const bzk_fr *val) {
    d.nnz = rp[nrows];
    BZK_CUDA(ctx, cudaMalloc(&d.rowptr, (nrows + 1) * sizeof(uint64_t)));
    BZK_CUDA(ctx, cudaMalloc(&d.col, (d.nnz ? d.nnz : 1) * sizeof(uint32_t)));
    BZK_CUDA(ctx, cudaMalloc(&d.val, (d.nnz ? d.nnz : 1) * sizeof(Fr)));
    BZK_CUDA(ctx, cudaMemcpyAsync(d.rowptr, rp, (nrows + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    BZK_CUDA(ctx, cudaMemcpyAsync(d.col, col, d.nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    BZK_CUDA(ctx, cudaMemcpyAsync(d.val, val, d.nnz * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    return BZK_OK;
}
static void free_r1cs(bzk_r1cs *r) {
    for (auto &m : r->m) { if (m.rowptr) cudaFree(m.rowptr); if (m.col) cudaFree(m.col); if (m.val) cudaFree(m.val); }
    if (r->d_a_idx) cudaFree(r->d_a_idx);
    if (r->d_b_idx) cudaFree(r->d_b_idx);
    delete r;
}

}  // namespace bzk

using namespace bzk;

extern "C" {

int32_t bzk_r1cs_upload(bzk_ctx *ctx, uint64_t num_inputs, uint64_t num_aux, uint64_t ncons,
                        const uint64_t *a_rp, const uint32_t *a_col, const bzk_fr *a_val,
                        const uint64_t *b_rp, const uint32_t *b_col, const bzk_fr *b_val,
                        const uint64_t *c_rp, const uint32_t *c_col, const bzk_fr *c_val, bzk_r1cs **out) {
    if (!ctx || !out || !a_rp || !b_rp || !c_rp || num_inputs == 0) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    *out = nullptr;
    const uint64_t nv = num_inputs + num_aux;
    if (nv >= (1ull << 32)) return BZK_ERR_BAD_ARG;
    const uint64_t *rp[3] = {a_rp, b_rp, c_rp};
    const uint32_t *cl[3] = {a_col, b_col, c_col};
    const bzk_fr *vl[3] = {a_val, b_val, c_val};
    for (int s = 0; s < 3; s++) {
        if (rp[s][0] != 0) return BZK_ERR_BAD_ARG;
        for (uint64_t j = 0; j < ncons; j++) if (rp[s][j + 1] < rp[s][j]) return BZK_ERR_BAD_ARG;
        if (rp[s][ncons] && (!cl[s] || !vl[s])) return BZK_ERR_BAD_ARG;
        for (uint64_t k = 0; k < rp[s][ncons]; k++) if (cl[s][k] >= nv) return BZK_ERR_BAD_ARG;
    }
    bzk_r1cs *r = new (std::nothrow) bzk_r1cs();
    if (!r) return BZK_ERR_OOM;
    r->num_inputs = num_inputs; r->num_aux = num_aux; r->ncons = ncons;
    uint64_t rows = ncons + num_inputs, m = 1;
    while (m < rows) { m <<= 1; r->log_m++; }
    if (r->log_m > 28) { delete r; return BZK_ERR_BAD_ARG; }
    // density (bellman `eval`: terms with a zero coefficient are skipped): A over aux only (all
    // inputs are always present), B over inputs and aux
    std::vector<uint8_t> a_d(nv, 0), b_d(nv, 0);
    auto nonzero = [](const bzk_fr &v) { return (v.l[0] | v.l[1] | v.l[2] | v.l[3]) != 0; };
    for (uint64_t k = 0; k < a_rp[ncons]; k++) if (nonzero(a_val[k])) a_d[a_col[k]] = 1;
    for (uint64_t k = 0; k < b_rp[ncons]; k++) if (nonzero(b_val[k])) b_d[b_col[k]] = 1;
    std::vector<uint32_t> a_idx, b_idx;
    for (uint64_t v = 0; v < num_inputs; v++) a_idx.push_back((uint32_t)v);
    for (uint64_t v = num_inputs; v < nv; v++) if (a_d[v]) a_idx.push_back((uint32_t)v);
    for (uint64_t v = 0; v < nv; v++) if (b_d[v]) b_idx.push_back((uint32_t)v);
    r->a_len = a_idx.size(); r->b_len = b_idx.size();
    int32_t st = BZK_OK;
    for (int s = 0; s < 3 && st == BZK_OK; s++) st = upload_csr(ctx, r->m[s], ncons, rp[s], cl[s], vl[s]);
    if (st == BZK_OK) {
        cudaError_t e = cudaMalloc(&r->d_a_idx, (a_idx.size() + 1) * 4);
        if (e == cudaSuccess) e = cudaMalloc(&r->d_b_idx, (b_idx.size() + 1) * 4);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_a_idx, a_idx.data(), a_idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_b_idx, b_idx.data(), b_idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = set_cuda_err(ctx, e, "r1cs upload", __FILE__, __LINE__);
    }
    if (st != BZK_OK) { free_r1cs(r); return st; }
    *out = r;
    return BZK_OK;
}

int32_t bzk_r1cs_free(bzk_ctx *ctx, bzk_r1cs *r) {
    if (!ctx) return BZK_ERR_BAD_ARG;
    if (!r) return BZK_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_r1cs(r);
    return BZK_OK;
}

int32_t bzk_r1cs_shape(const bzk_r1cs *r, uint64_t out[5]) {
    if (!r || !out) return BZK_ERR_BAD_ARG;
    out[0] = r->log_m;
    out[1] = ((uint64_t)1 << r->log_m) - 1;  // |h|
    out[2] = r->num_aux;                      // |l|
    out[3] = r->a_len;                        // |a|
    out[4] = r->b_len;                        // |b_g1| = |b_g2|
    return BZK_OK;
}

int32_t bzk_csr_spmv_dev(bzk_ctx *ctx, const void *d_rowptr, const void *d_col, const void *d_val, uint64_t nrows, const void *d_vec, void *d_out) {
    if (!ctx || (nrows && (!d_rowptr || !d_vec || !d_out))) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nrows == 0) return BZK_OK;
    k_csr_spmv<<<div_up(nrows, 256), 256, 0, ctx->stream>>>((const uint64_t *)d_rowptr, (const uint32_t *)d_col, (const Fr *)d_val, nrows, (const Fr *)d_vec, (Fr *)d_out);
    BZK_LAUNCHED(ctx);
    return BZK_OK;
}

int32_t bzk_g1_fixed_base_mul_dev(bzk_ctx *ctx, const bzk_g1_affine *base, const void *d_scalars, size_t n, void *d_out) {
    if (!ctx || !base || (n && (!d_scalars || !d_out))) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return BZK_OK;
    k_fixed_base_g1<<<div_up(n, 128), 128, 0, ctx->stream>>>(g1_from_img(base), (const Fr *)d_scalars, n, (uint8_t *)d_out);
    BZK_LAUNCHED(ctx);
    return BZK_OK;
}
int32_t bzk_g2_fixed_base_mul_dev(bzk_ctx *ctx, const bzk_g2_affine *base, const void *d_scalars, size_t n, void *d_out) {
    if (!ctx || !base || (n && (!d_scalars || !d_out))) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return BZK_OK;
    k_fixed_base_g2<<<div_up(n, 64), 64, 0, ctx->stream>>>(g2_from_img(base), (const Fr *)d_scalars, n, (uint8_t *)d_out);
    BZK_LAUNCHED(ctx);
    return BZK_OK;
}

int32_t bzk_groth16_params_create(bzk_ctx *ctx, const bzk_g1_affine *alpha_g1, const bzk_g1_affine *beta_g1, const bzk_g2_affine *beta_g2,
                                  const bzk_g1_affine *delta_g1, const bzk_g2_affine *delta_g2,
                                  bzk_g1_bases *h, bzk_g1_bases *l, bzk_g1_bases *a, bzk_g1_bases *b_g1, bzk_g2_bases *b_g2,
                                  bzk_groth16_params **out) {
    if (!ctx || !out || !alpha_g1 || !beta_g1 || !beta_g2 || !delta_g1 || !delta_g2 || !h || !l || !a || !b_g1 || !b_g2) return BZK_ERR_BAD_ARG;
    if (b_g1->n != b_g2->n) return BZK_ERR_BAD_ARG;
    bzk_groth16_params *p = new (std::nothrow) bzk_groth16_params();
    if (!p) return BZK_ERR_OOM;
    p->alpha_g1 = g1_from_img(alpha_g1); p->beta_g1 = g1_from_img(beta_g1); p->delta_g1 = g1_from_img(delta_g1);
    p->beta_g2 = g2_from_img(beta_g2); p->delta_g2 = g2_from_img(delta_g2);
    p->h = h; p->l = l; p->a = a; p->b1 = b_g1; p->b2 = b_g2;
    *out = p;
    return BZK_OK;
}
/* frees the handle and the five base vectors it adopted */
int32_t bzk_groth16_params_free(bzk_ctx *ctx, bzk_groth16_params *p) {
    if (!ctx) return BZK_ERR_BAD_ARG;
    if (!p) return BZK_OK;
    bzk_g1_bases_free(ctx, p->h); bzk_g1_bases_free(ctx, p->l); bzk_g1_bases_free(ctx, p->a); bzk_g1_bases_free(ctx, p->b1);
    bzk_g2_bases_free(ctx, p->b2);
    delete p;
    return BZK_OK;
}

// witness_kind: where `inputs` / `aux` live (cudaMemcpyHostToDevice: host images; cudaMemcpyDeviceToDevice:
// already resident, e.g. written by bzk_witness_run_dev)
static int32_t groth16_prove_impl(bzk_ctx *ctx, const bzk_groth16_params *pk, const bzk_r1cs *cs, const bzk_fr *inputs, const bzk_fr *aux,
                                  cudaMemcpyKind witness_kind, const bzk_fr *r_mont, const bzk_fr *s_mont, int32_t check_satisfied,
                                  bzk_g1_affine *proof_a, bzk_g2_affine *proof_b, bzk_g1_affine *proof_c,
                                  const Groth16Partials *partial = nullptr, const Groth16Split *split = nullptr) {
    const int phase = split ? split->phase : 0;
    if (!ctx || !pk || !cs || (phase != 2 && (!inputs || (cs->num_aux && !aux)))) return BZK_ERR_BAD_ARG;
    if (!partial && phase != 1 && (!r_mont || !s_mont || !proof_a || !proof_b || !proof_c)) return BZK_ERR_BAD_ARG;
    if (!partial && phase != 1 && pk->world != 1) return BZK_ERR_BAD_ARG;  // a shard can only produce partial sums
    if (phase == 2 && (!partial || !ctx->split_open || (!split->h_shard && pk->h->n))) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    // BZK_TRACE=1: host-side wall clock of the driver's phases on stderr (development aid)
    static const bool trace = std::getenv("BZK_TRACE") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[bzk prove] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
    };
    const uint64_t ni = cs->num_inputs, na = cs->num_aux, nv = ni + na, m = (uint64_t)1 << cs->log_m;
    // this handle's slice of each sum (the whole range when world == 1)
    auto lo_of = [&](uint64_t len) { return len * pk->rank / pk->world; };
    auto cnt_of = [&](uint64_t len) { return len * (pk->rank + 1) / pk->world - len * pk->rank / pk->world; };
    const uint64_t h_lo = lo_of(m - 1), h_n = cnt_of(m - 1), l_lo = lo_of(na), l_n = cnt_of(na), a_lo = lo_of(cs->a_len), a_n = cnt_of(cs->a_len),
                   b_lo = lo_of(cs->b_len), b_n = cnt_of(cs->b_len);
    if ((pk->world == 1 ? pk->h->n < h_n : pk->h->n != h_n) || pk->l->n != l_n || pk->a->n != a_n || pk->b1->n != b_n || pk->b2->n != b_n)
        return BZK_ERR_BAD_ARG;
    // staging arena: z | a_ev | b_ev | c_ev | gathered scalars
    const size_t gmax = std::max<uint64_t>(std::max(cs->a_len, cs->b_len), 1);
    size_t need = 0;
    {
        Carver cv(nullptr);
        cv.take<Fr>(nv); cv.take<Fr>(m); cv.take<Fr>(m); cv.take<Fr>(m); cv.take<Fr>(gmax); cv.take<Fr>(gmax); cv.take<uint32_t>(4);
        need = cv.used();
    }
    BZK_TRY(ensure_ws(ctx, &ctx->stage, &ctx->stage_bytes, need));
    Carver cv(ctx->stage);
    Fr *z = cv.take<Fr>(nv), *ea = cv.take<Fr>(m), *eb = cv.take<Fr>(m), *ec = cv.take<Fr>(m), *gs_a = cv.take<Fr>(gmax), *gs_b = cv.take<Fr>(gmax);
    uint32_t *d_bad = cv.take<uint32_t>(4);
    cudaStream_t st = ctx->stream;
    const bool timed = ctx->timing;
    auto g16_mark = [&](int k, cudaStream_t s) {
        if (!timed) return;
        if (!ctx->g16_ev[k]) cudaEventCreate(&ctx->g16_ev[k]);
        cudaEventRecord(ctx->g16_ev[k], s);
    };
    MsmPlan *plan = ctx->g16_plan;
    constexpr size_t kWinBytes = kMaxWinPoints * sizeof(G2Xyzz);
    if (phase != 2) {
    ctx->split_open = false;
    ctx->g16_valid = false;
    g16_mark(0, st);
    BZK_CUDA(ctx, cudaMemcpyAsync(z, inputs, ni * sizeof(Fr), witness_kind, st));
    if (na) BZK_CUDA(ctx, cudaMemcpyAsync(z + ni, aux, na * sizeof(Fr), witness_kind, st));
    // evaluations (rows >= ncons: the Input(i)*0=0 rows, then zero padding)
    Fr *ev[3] = {ea, eb, ec};
    if (phase == 1) {
        for (int s = 0; s < 3; s++) {
            if (((split->poly_mask >> s) & 1) && !split->evals[s]) return BZK_ERR_BAD_ARG;
            ev[s] = ((split->poly_mask >> s) & 1) ? split->evals[s] : nullptr;
        }
        ea = ev[0];
    }
    for (int s = 0; s < 3; s++) {
        if (!ev[s]) continue;
        BZK_CUDA(ctx, cudaMemsetAsync(ev[s] + cs->ncons, 0, (m - cs->ncons) * sizeof(Fr), st));
        if (cs->ncons) {
            k_csr_spmv<<<div_up(cs->ncons, 256), 256, 0, st>>>(cs->m[s].rowptr, cs->m[s].col, cs->m[s].val, cs->ncons, z, ev[s]);
            BZK_LAUNCHED(ctx);
        }
    }
    if (ea) BZK_CUDA(ctx, cudaMemcpyAsync(ea + cs->ncons, z, ni * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    if (check_satisfied && cs->ncons && phase == 0) {
        BZK_CUDA(ctx, cudaMemsetAsync(d_bad, 0, 4, st));
        k_check_sat<<<div_up(cs->ncons, 256), 256, 0, st>>>(ea, eb, ec, cs->ncons, d_bad);
        BZK_LAUNCHED(ctx);
        uint32_t bad = 0;
        BZK_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
        BZK_CUDA(ctx, cudaStreamSynchronize(st));
        if (bad) {
            snprintf(ctx->err, sizeof ctx->err, "%u constraints unsatisfied by the witness", bad);
            return BZK_ERR_UNSAT;
        }
    }
    // The five sums are independent once z is on the device (h additionally needs the quotient):
    // l, a, b_g1, b_g2 run on side streams with their own arenas while the main stream does the
    // NTT pipeline and the h sum, so the latency-bound phases of one MSM (bucket reduction, side-list
    // folding) overlap the throughput-bound phases of the others.
    for (int k = 0; k < 4; k++)
        if (!ctx->aux_stream[k]) BZK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream[k], cudaStreamNonBlocking));
    for (int k = 0; k < 3; k++)
        if (!ctx->aux_ev[k]) BZK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_ev[k], cudaEventDisableTiming));
    if (ctx->pinned_bytes < 5 * kWinBytes) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        BZK_CUDA(ctx, cudaHostAlloc(&ctx->pinned, 5 * kWinBytes, cudaHostAllocDefault));
        ctx->pinned_bytes = 5 * kWinBytes;
    }
    char *hw = (char *)ctx->pinned;
    lap("z + evaluations enqueued");
    cudaStream_t s_l = ctx->aux_stream[0], s_a = ctx->aux_stream[1], s_b1 = ctx->aux_stream[2], s_b2 = ctx->aux_stream[3];
    g16_mark(1, st);
    BZK_CUDA(ctx, cudaEventRecord(ctx->aux_ev[0], st));  // z and the evaluations are enqueued behind this point
    BZK_CUDA(ctx, cudaStreamWaitEvent(s_l, ctx->aux_ev[0], 0));
    BZK_CUDA(ctx, cudaStreamWaitEvent(s_a, ctx->aux_ev[0], 0));
    BZK_CUDA(ctx, cudaStreamWaitEvent(s_b1, ctx->aux_ev[0], 0));
    BZK_TRY(msm_g1_enqueue(ctx, s_l, &ctx->aux_ws[0], &ctx->aux_ws_bytes[0], bases_ref(pk->l), z + ni + l_lo, l_n, hw + 1 * kWinBytes, &plan[1]));
    k_gather_fr<<<div_up(cs->a_len, 256), 256, 0, s_a>>>(z, cs->d_a_idx, cs->a_len, gs_a);
    BZK_LAUNCHED(ctx);
    BZK_TRY(msm_g1_enqueue(ctx, s_a, &ctx->aux_ws[1], &ctx->aux_ws_bytes[1], bases_ref(pk->a), gs_a + a_lo, a_n, hw + 2 * kWinBytes, &plan[2]));
    if (cs->b_len) {
        k_gather_fr<<<div_up(cs->b_len, 256), 256, 0, s_b1>>>(z, cs->d_b_idx, cs->b_len, gs_b);
        BZK_LAUNCHED(ctx);
    }
    BZK_CUDA(ctx, cudaEventRecord(ctx->aux_ev[1], s_b1));
    BZK_CUDA(ctx, cudaStreamWaitEvent(s_b2, ctx->aux_ev[1], 0));
    BZK_TRY(msm_g1_enqueue(ctx, s_b1, &ctx->aux_ws[2], &ctx->aux_ws_bytes[2], bases_ref(pk->b1), gs_b + b_lo, b_n, hw + 3 * kWinBytes, &plan[3]));
    BZK_TRY(msm_g2_enqueue(ctx, s_b2, &ctx->aux_ws[3], &ctx->aux_ws_bytes[3], bases_ref(pk->b2), gs_b + b_lo, b_n, hw + 4 * kWinBytes, &plan[4]));
    if (phase == 1) {
        // this rank's evaluation vectors to the coset (ifft, then coset_fft); the pointwise step and the last transform
        // happen on the rank that collects the three (bzk_groth16_h_combine_dev)
        for (int s = 0; s < 3; s++)
            if (ev[s]) BZK_TRY(groth16_to_coset_launch(ctx, ev[s], cs->log_m));
        g16_mark(2, st);
        BZK_CUDA(ctx, cudaStreamSynchronize(st));  // the caller hands the vectors to its transport next
        ctx->split_open = true;
        return BZK_OK;
    }
    BZK_TRY(groth16_h_launch(ctx, ea, eb, ec, cs->log_m));  // ea <- h coefficients
    g16_mark(2, st);
    }  // phase != 2
    char *hw = (char *)ctx->pinned;
    const Fr *h_src = phase == 2 ? split->h_shard : ea + h_lo;
    ctx->split_open = false;
    BZK_TRY(msm_g1_enqueue(ctx, st, &ctx->ws, &ctx->ws_bytes, bases_ref(pk->h), h_src, h_n, hw, &plan[0]));
    g16_mark(3, st);
    for (int k = 0; k < 4; k++) g16_mark(4 + k, ctx->aux_stream[k]);
    lap("all kernels enqueued");
    BZK_CUDA(ctx, cudaStreamSynchronize(st));
    lap("main stream done");
    for (int k = 0; k < 4; k++) BZK_CUDA(ctx, cudaStreamSynchronize(ctx->aux_stream[k]));
    lap("side streams done");
    if (timed) {
        for (int k = 1; k < 8; k++) cudaEventElapsedTime(&ctx->g16_ms[k], ctx->g16_ev[0], ctx->g16_ev[k]);
        ctx->g16_ms[0] = 0;
        ctx->g16_valid = true;
    }
    // host tail: the five Horner folds and the (r, s) scalar multiplications are independent
    // sub-millisecond jobs — run them on host threads instead of back to back
    bzk_g1_affine h_ans, l_ans, a_ans, b1_ans;
    bzk_g2_affine b2_ans;
    if (partial) {  // sharded schedule: hand back this rank's four partial sums; the caller folds and finalises
        auto p_b2 = std::async(std::launch::async, [&] { msm_g2_finish(&plan[4], hw + 4 * kWinBytes, partial->b2_sum); });
        auto p_a = std::async(std::launch::async, [&] { msm_g1_finish(&plan[2], hw + 2 * kWinBytes, partial->a_sum); });
        auto p_b1 = std::async(std::launch::async, [&] { msm_g1_finish(&plan[3], hw + 3 * kWinBytes, partial->b1_sum); });
        auto p_l = std::async(std::launch::async, [&] { msm_g1_finish(&plan[1], hw + 1 * kWinBytes, &l_ans); });
        msm_g1_finish(&plan[0], hw, &h_ans);
        p_l.get();
        G1Xyzz t = G1Xyzz::from_affine(g1_from_img(&h_ans));
        t.madd(g1_from_img(&l_ans));
        g1_to_img(partial->hl_sum, t.to_affine());
        p_b2.get(); p_a.get(); p_b1.get();
        lap("partial sums folded");
        return BZK_OK;
    }
    Fr r, s;
    memcpy(r.l, r_mont, 32);
    memcpy(s.l, s_mont, 32);
    const Fr rs = (r * s).from_mont(), rc = r.from_mont(), sc = s.from_mont();
    auto f_b2 = std::async(std::launch::async, [&] {
        msm_g2_finish(&plan[4], hw + 4 * kWinBytes, &b2_ans);
        G2Xyzz gb = scalar_mul(pk->delta_g2, sc.l);
        gb.madd(pk->beta_g2);
        gb.madd(g2_from_img(&b2_ans));
        return gb.to_affine();
    });
    auto f_a = std::async(std::launch::async, [&] {
        msm_g1_finish(&plan[2], hw + 2 * kWinBytes, &a_ans);
        return scalar_mul(g1_from_img(&a_ans), sc.l);  // s * a_sum
    });
    auto f_b1 = std::async(std::launch::async, [&] {
        msm_g1_finish(&plan[3], hw + 3 * kWinBytes, &b1_ans);
        return scalar_mul(g1_from_img(&b1_ans), rc.l);  // r * b1_sum
    });
    auto f_hl = std::async(std::launch::async, [&] {
        msm_g1_finish(&plan[0], hw, &h_ans);
        msm_g1_finish(&plan[1], hw + 1 * kWinBytes, &l_ans);
        G1Xyzz t = G1Xyzz::from_affine(g1_from_img(&h_ans));
        t.madd(g1_from_img(&l_ans));
        return t;  // h_sum + l_sum
    });
    auto f_c0 = std::async(std::launch::async, [&] {
        G1Xyzz t = scalar_mul(pk->delta_g1, rs.l);
        t.add(scalar_mul(pk->alpha_g1, sc.l));
        t.add(scalar_mul(pk->beta_g1, rc.l));
        return t;  // r s delta + s alpha + r beta
    });
    G1Xyzz ga = scalar_mul(pk->delta_g1, rc.l);
    ga.madd(pk->alpha_g1);
    G1Xyzz gc = f_c0.get();
    gc.add(f_a.get());
    gc.add(f_b1.get());
    gc.add(f_hl.get());
    ga.madd(g1_from_img(&a_ans));
    const G2Affine gb_aff = f_b2.get();
    lap("host folds + blinding tail");
    g1_to_img(proof_a, ga.to_affine());
    g2_to_img(proof_b, gb_aff);
    g1_to_img(proof_c, gc.to_affine());
    return BZK_OK;
}

int32_t bzk_groth16_prove(bzk_ctx *ctx, const bzk_groth16_params *pk, const bzk_r1cs *cs, const bzk_fr *inputs, const bzk_fr *aux,
                          const bzk_fr *r_mont, const bzk_fr *s_mont, int32_t check_satisfied,
                          bzk_g1_affine *proof_a, bzk_g2_affine *proof_b, bzk_g1_affine *proof_c) {
    return groth16_prove_impl(ctx, pk, cs, inputs, aux, cudaMemcpyHostToDevice, r_mont, s_mont, check_satisfied, proof_a, proof_b, proof_c);
}

int32_t bzk_groth16_prove_dev(bzk_ctx *ctx, const bzk_groth16_params *pk, const bzk_r1cs *cs, const void *d_inputs, const void *d_aux,
                              const bzk_fr *r_mont, const bzk_fr *s_mont, int32_t check_satisfied,
                              bzk_g1_affine *proof_a, bzk_g2_affine *proof_b, bzk_g1_affine *proof_c) {
    return groth16_prove_impl(ctx, pk, cs, (const bzk_fr *)d_inputs, (const bzk_fr *)d_aux, cudaMemcpyDeviceToDevice, r_mont, s_mont,
                              check_satisfied, proof_a, proof_b, proof_c);
}

/* milliseconds since the start of the last timed prove call (bzk_ctx_set_timing on) at which: [1] z upload + the three
 * SpMVs (+ satisfiability check) finished, [2] the quotient pipeline (7 NTTs) finished, [3] the h sum finished (main
 * stream), [4..7] the l / a / b_g1 / b_g2 sums finished (side streams, concurrent with the main one).  Returns 1 if valid. */
int32_t bzk_groth16_stage_ms(const bzk_ctx *ctx, float out[8]) {
    if (!ctx || !out) return BZK_ERR_BAD_ARG;
    for (int k = 0; k < 8; k++) out[k] = ctx->g16_ms[k];
    return ctx->g16_valid ? 1 : 0;
}

/* Fixed-base tables for the five base vectors of a key (they never change between proofs): up to `max_levels` levels
 * [2^(c*G*t)] P per base, so that the windows of a scalar share ceil(W/levels) bucket groups — fewer, larger windows and
 * one bucket reduction per group instead of per window.  max_levels = 0 picks the largest count (<= 16) whose tables fit
 * in `mem_fraction_percent` % of the currently free device memory.  Memory: levels x the key's size. */
int32_t bzk_groth16_params_precompute(bzk_ctx *ctx, bzk_groth16_params *p, uint32_t max_levels, uint32_t mem_fraction_percent) {
    if (!ctx || !p) return BZK_ERR_BAD_ARG;
    BZK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (max_levels == 0) {
        size_t free_b = 0, total_b = 0;
        BZK_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        const double key_bytes = (double)(p->h->n + p->l->n + p->a->n + p->b1->n) * sizeof(G1Affine) + (double)p->b2->n * sizeof(G2Affine);
        const double budget = (double)free_b * (mem_fraction_percent ? mem_fraction_percent : 50) / 100.0;
        uint32_t lv = key_bytes > 0 ? (uint32_t)(budget / key_bytes) + 1 : 16;  // level 0 is already resident
        max_levels = lv > 16 ? 16 : lv;
    }
    if (max_levels <= 1) return BZK_OK;
    BZK_TRY(precompute_g1(ctx, p->h, max_levels));
    BZK_TRY(precompute_g1(ctx, p->l, max_levels));
    BZK_TRY(precompute_g1(ctx, p->a, max_levels));
    BZK_TRY(precompute_g1(ctx, p->b1, max_levels));
    BZK_TRY(precompute_g2(ctx, p->b2, max_levels));
    return BZK_OK;
}

int32_t bzk_groth16_params_set_shard(bzk_groth16_params *p, uint32_t rank, uint32_t world) {
    if (!p || world == 0 || rank >= world) return BZK_ERR_BAD_ARG;
    p->rank = rank;
    p->world = world;
    return BZK_OK;
}

int32_t bzk_groth16_prove_partial(bzk_ctx *ctx, const bzk_groth16_params *pk, const bzk_r1cs *cs, const void *inputs, const void *aux,
                                  int32_t witness_on_device, int32_t check_satisfied,
                                  bzk_g1_affine *a_sum, bzk_g1_affine *b1_sum, bzk_g2_affine *b2_sum, bzk_g1_affine *hl_sum) {
    if (!a_sum || !b1_sum || !b2_sum || !hl_sum) return BZK_ERR_BAD_ARG;
    const Groth16Partials part{a_sum, b1_sum, hl_sum, b2_sum};
    return groth16_prove_impl(ctx, pk, cs, (const bzk_fr *)inputs, (const bzk_fr *)aux,
                              witness_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, nullptr, nullptr, check_satisfied,
                              nullptr, nullptr, nullptr, &part);
}

/* The sharded schedule with the quotient pipeline split over the ranks (include/bzk.h).  begin: z, the evaluation vectors in
 * `poly_mask` (bit 0 = a, 1 = b, 2 = c) computed into the caller's buffers and taken to the coset, the l / a / b_g1 / b_g2
 * partial sums enqueued on their streams (they keep running while the caller moves vectors between GPUs).  finish: the h sum
 * over this rank's slice of the quotient coefficients, then the four partial sums as bzk_groth16_prove_partial returns them. */
int32_t bzk_groth16_shard_begin(bzk_ctx *ctx, const bzk_groth16_params *pk, const bzk_r1cs *cs, const void *inputs, const void *aux,
                                int32_t witness_on_device, uint32_t poly_mask, void *d_evals[3]) {
    if (!d_evals || poly_mask > 7) return BZK_ERR_BAD_ARG;
    const Groth16Split sp{1, poly_mask, {(Fr *)d_evals[0], (Fr *)d_evals[1], (Fr *)d_evals[2]}, nullptr};
    return groth16_prove_impl(ctx, pk, cs, (const bzk_fr *)inputs, (const bzk_fr *)aux,
                              witness_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                              nullptr, &sp);
}
int32_t bzk_groth16_shard_finish(bzk_ctx *ctx, const bzk_groth16_params *pk, const bzk_r1cs *cs, const void *d_h_shard,
                                 bzk_g1_affine *a_sum, bzk_g1_affine *b1_sum, bzk_g2_affine *b2_sum, bzk_g1_affine *hl_sum) {
    if (!a_sum || !b1_sum || !b2_sum || !hl_sum) return BZK_ERR_BAD_ARG;
    const Groth16Partials part{a_sum, b1_sum, hl_sum, b2_sum};
    const Groth16Split sp{2, 0, {nullptr, nullptr, nullptr}, (const Fr *)d_h_shard};
    return groth16_prove_impl(ctx, pk, cs, nullptr, nullptr, cudaMemcpyDeviceToDevice, nullptr, nullptr, 0, nullptr, nullptr, nullptr, &part, &sp);
}

/* bellman `create_proof`'s last lines from the (summed) answers:
 *   A = r*delta1 + alpha1 + a;  B = s*delta2 + beta2 + b2;  C = rs*delta1 + s*alpha1 + r*beta1 + s*a + r*b1 + (h + l) */
int32_t bzk_groth16_finalize(const bzk_g1_affine *alpha_g1, const bzk_g1_affine *beta_g1, const bzk_g2_affine *beta_g2,
                             const bzk_g1_affine *delta_g1, const bzk_g2_affine *delta_g2,
                             const bzk_g1_affine *a_sum, const bzk_g1_affine *b1_sum, const bzk_g2_affine *b2_sum, const bzk_g1_affine *hl_sum,
                             const bzk_fr *r_mont, const bzk_fr *s_mont, bzk_g1_affine *proof_a, bzk_g2_affine *proof_b, bzk_g1_affine *proof_c) {
    if (!alpha_g1 || !beta_g1 || !beta_g2 || !delta_g1 || !delta_g2 || !a_sum || !b1_sum || !b2_sum || !hl_sum || !r_mont || !s_mont ||
        !proof_a || !proof_b || !proof_c)
        return BZK_ERR_BAD_ARG;
    Fr r, s;
    memcpy(r.l, r_mont, 32);
    memcpy(s.l, s_mont, 32);
    const Fr rs = (r * s).from_mont(), rc = r.from_mont(), sc = s.from_mont();
    const G1Affine al = g1_from_img(alpha_g1), be1 = g1_from_img(beta_g1), de1 = g1_from_img(delta_g1), av = g1_from_img(a_sum);
    G2Xyzz gb = scalar_mul(g2_from_img(delta_g2), sc.l);
    gb.madd(g2_from_img(beta_g2));
    gb.madd(g2_from_img(b2_sum));
    G1Xyzz ga = scalar_mul(de1, rc.l);
    ga.madd(al);
    ga.madd(av);
    G1Xyzz gc = scalar_mul(de1, rs.l);
    gc.add(scalar_mul(al, sc.l));
    gc.add(scalar_mul(be1, rc.l));
    gc.add(scalar_mul(av, sc.l));
    gc.add(scalar_mul(g1_from_img(b1_sum), rc.l));
    gc.madd(g1_from_img(hl_sum));
    g1_to_img(proof_a, ga.to_affine());
    g2_to_img(proof_b, gb.to_affine());
    g1_to_img(proof_c, gc.to_affine());
    return BZK_OK;
}

/* 387-byte bincode image of `Groth16Proof {a, b, c}` (/root/reference/src/zk/groth16/mod.rs:33-38) */
int32_t bzk_groth16_proof_bytes(const bzk_g1_affine *a, const bzk_g2_affine *b, const bzk_g1_affine *c, uint8_t out[387]) {
    if (!a || !b || !c || !out) return BZK_ERR_BAD_ARG;
    memcpy(out, a, 97);
    memcpy(out + 97, b, 193);
    memcpy(out + 290, c, 97);
    return BZK_OK;
}

}  // extern "C"

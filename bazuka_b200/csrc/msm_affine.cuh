// bazuka_b200 — batched-affine bucket accumulation: the first rounds of the bucket sums as pairwise AFFINE additions
// that share one field inversion per round across the whole grid.
//
// The sorted entry list of an MSM (msm_impl.cuh) groups point references by bucket.  One ROUND halves every bucket: the
// entries of a bucket are paired in order, each pair is replaced by its affine sum, an odd last entry passes through,
//     m entries -> ceil(m/2) entries,        sum over all buckets of floor(m/2) additions,
// and after R rounds the shortened lists go to the XYZZ accumulate kernel as before (which handles any list, so R is a
// tuning parameter, not a correctness one).  An affine addition is 5M + 1S given the inverse of its denominator; all
// the denominators of a round are inverted together with Montgomery's trick laid over the launch geometry:
//
//   k_round_fwd   thread t walks its share of the output slots, multiplies the denominators of its pairs into a running
//                 product and stores the running product BEFORE each pair; the block then scans its threads' totals
//                 (exclusive prefix and suffix products in shared memory) and publishes the block total
//   k_round_mid   one block: prefix / suffix products over the block totals and THE inversion of the grand total
//   k_round_bwd   thread t recovers 1/(its total) = prefix * suffix * 1/grand  (4 products), walks its pairs backwards
//                 peeling one inverse per pair (2 products) and writes the sums
//
// so the serial Fermat chain (~460 dependent products, one warp) is paid once per round instead of once per addition
// or per thread.  Point references are u32: bit 31 = negate y, bit 30 = pool (0: the base table, 1: the scratch pool of
// intermediate sums), 30 index bits.  Exceptional pairs (identity operands, P + P, P - P) are handled by
// pair_denominator / pair_sum (ec.cuh), so repeated and adversarial bases give exact results here too.
#pragma once
#include "common.cuh"

namespace bzk {

constexpr uint32_t kRefSign = 1u << 31, kRefPool = 1u << 30, kRefIdx = kRefPool - 1;
constexpr uint32_t kRoundThreads = 128;

template <class F>
__device__ __forceinline__ Affine<F> load_ref(const Affine<F> *__restrict__ tab, const Affine<F> *__restrict__ scr, uint32_t e) {
    Affine<F> p = load_vec(((e & kRefPool) ? scr : tab) + (e & kRefIdx));
    if (e >> 31) p.y = p.y.neg();
    return p;
}

// counts[b] = ceil(m_b / 2): the bucket's entry count after one round
static __global__ void __launch_bounds__(256) k_round_counts(const uint32_t *__restrict__ off, uint32_t TB, uint32_t *__restrict__ counts) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < TB) counts[b] = (off[b + 1] - off[b] + 1) >> 1;
    if (b == TB) counts[b] = 0;
}

// the slot range of thread t and the bucket holding its first slot
struct RoundSpan { uint32_t s0, s1, b; };
__device__ __forceinline__ RoundSpan round_span(const uint32_t *__restrict__ off1, uint32_t TB, uint32_t nthreads, uint32_t t) {
    const uint32_t M1 = off1[TB];
    const uint32_t chunk = (M1 + nthreads - 1) / nthreads;
    RoundSpan r;
    const uint64_t a = (uint64_t)t * chunk;
    r.s0 = a < M1 ? (uint32_t)a : M1;
    r.s1 = a + chunk < M1 ? (uint32_t)(a + chunk) : M1;
    r.b = 0;
    if (r.s0 < r.s1) {
        uint32_t lo = 0, hi = TB;  // off1[lo] <= s0 < off1[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off1[mid] <= r.s0) lo = mid; else hi = mid;
        }
        r.b = lo;
    }
    return r;
}

// block-wide exclusive prefix and suffix products of one value per thread (Hillis-Steele in shared memory)
template <class F>
__device__ __forceinline__ void block_scan_products(F *sh /*[2*kRoundThreads]*/, const F &mine, F &pre_excl, F &suf_excl, F &total) {
    F *a = sh, *z = sh + kRoundThreads;
    const uint32_t i = threadIdx.x;
    a[i] = mine;
    z[i] = mine;
    __syncthreads();
    for (uint32_t o = 1; o < kRoundThreads; o <<= 1) {
        F pa = a[i], pz = z[i];
        const bool la = i >= o, lz = i + o < kRoundThreads;
        F qa, qz;
        if (la) qa = a[i - o];
        if (lz) qz = z[i + o];
        __syncthreads();
        if (la) a[i] = qa * pa;
        if (lz) z[i] = pz * qz;
        __syncthreads();
    }
    // a[i] = product of 0..i, z[i] = product of i..last
    pre_excl = i ? a[i - 1] : F::one();
    suf_excl = i + 1 < kRoundThreads ? z[i + 1] : F::one();
    total = a[kRoundThreads - 1];
}

template <class F>
__global__ void __launch_bounds__(kRoundThreads) k_round_fwd(const Affine<F> *__restrict__ tab, const Affine<F> *__restrict__ scr,
                                                             const uint32_t *__restrict__ list0, const uint32_t *__restrict__ off0,
                                                             const uint32_t *__restrict__ off1, uint32_t TB, F *__restrict__ pre,
                                                             F *__restrict__ thr_pre, F *__restrict__ thr_suf, F *__restrict__ blk_tot) {
    extern __shared__ uint4 smem_raw[];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    RoundSpan sp = round_span(off1, TB, gridDim.x * blockDim.x, t);
    F acc = F::one();
    uint32_t b = sp.b;
    for (uint32_t s = sp.s0; s < sp.s1; s++) {
        while (s >= off1[b + 1]) b++;
        const uint32_t j = s - off1[b], in0 = off0[b], m = off0[b + 1] - in0;
        if (2 * j + 1 < m) {
            const Affine<F> A = load_ref(tab, scr, list0[in0 + 2 * j]), B = load_ref(tab, scr, list0[in0 + 2 * j + 1]);
            store_vec(pre + s, acc);
            acc = acc * pair_denominator(A, B);
        }
    }
    F p, q, tot;
    block_scan_products((F *)smem_raw, acc, p, q, tot);
    store_vec(thr_pre + t, p);
    store_vec(thr_suf + t, q);
    if (threadIdx.x == 0) store_vec(blk_tot + blockIdx.x, tot);
}

// one block: exclusive prefix / suffix products of the block totals, and the inverse of the grand total
template <class F>
__global__ void __launch_bounds__(1024) k_round_mid(const F *__restrict__ blk_tot, uint32_t nblocks, F *__restrict__ blk_pre, F *__restrict__ blk_suf,
                                                    F *__restrict__ inv_total) {
    extern __shared__ uint4 smem_raw[];
    F *a = (F *)smem_raw, *z = a + blockDim.x;
    const uint32_t i = threadIdx.x;
    const F mine = i < nblocks ? load_vec(blk_tot + i) : F::one();
    a[i] = mine;
    z[i] = mine;
    __syncthreads();
    for (uint32_t o = 1; o < blockDim.x; o <<= 1) {
        F pa = a[i], pz = z[i], qa, qz;
        const bool la = i >= o, lz = i + o < blockDim.x;
        if (la) qa = a[i - o];
        if (lz) qz = z[i + o];
        __syncthreads();
        if (la) a[i] = qa * pa;
        if (lz) z[i] = pz * qz;
        __syncthreads();
    }
    if (i < nblocks) {
        store_vec(blk_pre + i, i ? a[i - 1] : F::one());
        store_vec(blk_suf + i, i + 1 < blockDim.x ? z[i + 1] : F::one());
    }
    if (i == 0) store_vec(inv_total, a[blockDim.x - 1].inv());  // no denominator is zero: the product is invertible
}

template <class F>
__global__ void __launch_bounds__(kRoundThreads) k_round_bwd(const Affine<F> *__restrict__ tab, Affine<F> *__restrict__ scr, uint32_t scr_base,
                                                             const uint32_t *__restrict__ list0, const uint32_t *__restrict__ off0,
                                                             const uint32_t *__restrict__ off1, uint32_t TB, const F *__restrict__ pre,
                                                             const F *__restrict__ thr_pre, const F *__restrict__ thr_suf, const F *__restrict__ blk_pre,
                                                             const F *__restrict__ blk_suf, const F *__restrict__ inv_total, uint32_t *__restrict__ list1) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    RoundSpan sp = round_span(off1, TB, gridDim.x * blockDim.x, t);
    if (sp.s0 >= sp.s1) return;
    // 1 / (this thread's total) = (everything before) * (everything after) / (grand total)
    F inv = load_vec(blk_pre + blockIdx.x) * load_vec(thr_pre + t);
    inv = inv * (load_vec(thr_suf + t) * load_vec(blk_suf + blockIdx.x));
    inv = inv * load_vec(inv_total);
    uint32_t lo = sp.b, hi = TB;  // bucket of the LAST slot: off1[lo] <= s1-1 < off1[hi]
    {
        const uint32_t last = sp.s1 - 1;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off1[mid] <= last) lo = mid; else hi = mid;
        }
    }
    uint32_t b = lo;
    for (uint32_t s = sp.s1; s-- > sp.s0;) {
        while (off1[b] > s) b--;
        const uint32_t j = s - off1[b], in0 = off0[b], m = off0[b + 1] - in0;
        const uint32_t ea = list0[in0 + 2 * j];
        if (2 * j + 1 < m) {
            const Affine<F> A = load_ref(tab, scr, ea), B = load_ref(tab, scr, list0[in0 + 2 * j + 1]);
            const F dinv = inv * load_vec(pre + s);
            inv = inv * pair_denominator(A, B);
            store_vec(scr + scr_base + s, pair_sum(A, B, dinv));
            list1[s] = (scr_base + s) | kRefPool;
        } else {
            list1[s] = ea;  // odd one out: carried to the next round as it is
        }
    }
}

}  // namespace bzk

// Feasibility probe: limb products on the FP64 pipe.  52-bit limbs held exactly in doubles; a limb
// product a*b = hi*2^52 + lo is recovered with two round-toward-zero FMAs (Emmart's trick) and the
// bit patterns are accumulated with 64-bit integer additions.  This measures how fast one 8x8
// schoolbook pass (64 limb products = 128 DFMA + 128 integer accumulations) runs per warp, to decide
// whether an FP64 Montgomery multiplier can beat the 12x32 IMAD.WIDE one (1276 cycles / product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
__global__ void __launch_bounds__(256) k_pass(const double *in, uint64_t *out) {
    double a[8], b[8];
    for (int i = 0; i < 8; i++) { a[i] = in[threadIdx.x * 16 + i]; b[i] = in[threadIdx.x * 16 + 8 + i]; }
    uint64_t acc[16];
    for (int i = 0; i < 16; i++) acc[i] = 0;
    const double C1 = 20282409603651670423947251286016.0;  // 2^104
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double h = __fma_rz(a[i], b[j], C1);        // 2^104 + hi*2^52
                double l = __fma_rz(a[i], b[j], C1 - h);    // low 52 bits (exact), C1-h = -hi*2^52
                acc[i + j + 1] += (uint64_t)__double_as_longlong(h);
                acc[i + j] += (uint64_t)__double_as_longlong(l + 4503599627370496.0);  // +2^52: integer in mantissa
            }
        // keep the loop from being hoisted: feed something back
        a[0] = __longlong_as_double((long long)((acc[3] & 0x000fffffffffffffULL) | 0x4330000000000000ULL)) - 4503599627370496.0;
    }
    uint64_t s = 0;
    for (int i = 0; i < 16; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *in; uint64_t *out;
    cudaMalloc(&in, 256 * 16 * 8); cudaMalloc(&out, 148 * 8 * 256 * 8);
    double h[256 * 16];
    for (int i = 0; i < 256 * 16; i++) h[i] = (double)((1ull << 51) + 12345ull * i);
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    for (int bps = 1; bps <= 4; bps *= 2) {
        int blocks = 148 * bps;
        k_pass<<<blocks, 256>>>(in, out);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k_pass<<<blocks, 256>>>(in, out); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double cycles = ms * 1e-3 * 1.965e9;
        double warps_per_smsp = bps * 8 / 4.0;
        printf("blocks/SM %d: %.3f ms -> %.0f cycles per 8x8 pass per warp (per SMSP, %g warps sharing)\n", bps, ms, cycles / ITERS / warps_per_smsp, warps_per_smsp);
    }
    printf("reference: 12x32 IMAD.WIDE Montgomery product = 1276 cycles/warp; an FP64 one needs ~2 passes + carries\n");
    return 0;
}

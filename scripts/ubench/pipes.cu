// Microbenchmark (development aid): how fast can one SM sub-partition run the K2 inner body?
//   mode 0: 32 packed FP32x2 ops per iteration (the FMA-pipe part of 2 lines x 8 points)
//   mode 1: 8 MUFU.RCP per iteration (the XU part)
//   mode 2: both, interleaved as in lorentz_paired
//   mode 3: scalar version of mode 2 (64 FP32 ops + 8 MUFU)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__device__ __forceinline__ float rcp_approx(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) body(float *out, int iters, float seed, long long *cyc) {
    float2 fi[4], acc[4];
    for (int h = 0; h < 4; ++h) { fi[h] = make_float2(seed + threadIdx.x + 64 * h, seed + threadIdx.x + 64 * h + 32); acc[h] = make_float2(0.f, 0.f); }
    float2 nf1 = make_float2(-seed * 3.f, -seed * 3.f), nf2 = make_float2(-seed * 5.f, -seed * 5.f);
    float2 B1 = make_float2(seed * 7.f, seed * 7.f), B2 = make_float2(seed * 11.f, seed * 11.f);
    float2 A1 = make_float2(seed, seed), A2 = make_float2(seed * 2.f, seed * 2.f);
    float2 nf3 = make_float2(-seed * 13.f, -seed * 13.f), nf4 = make_float2(-seed * 17.f, -seed * 17.f);
    float2 B3 = make_float2(seed * 19.f, seed * 19.f), B4 = make_float2(seed * 23.f, seed * 23.f);
    float2 A3 = make_float2(seed * 1e-3f, seed * 1e-3f), A4 = make_float2(seed * 2e-3f, seed * 2e-3f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            if (MODE == 0 || MODE == 2) {
                float2 e1 = __fadd2_rn(fi[h], nf1), e2 = __fadd2_rn(fi[h], nf2);
                float2 q1 = __ffma2_rn(e1, e1, B1), q2 = __ffma2_rn(e2, e2, B2);
                float2 num = __ffma2_rn(A2, q1, __fmul2_rn(A1, q2));
                float2 den = __fmul2_rn(q1, q2);
                float2 r = den;
                if (MODE == 2) r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
                acc[h] = __ffma2_rn(num, r, acc[h]);
            } else if (MODE == 4) {
                float2 e1 = __fadd2_rn(fi[h], nf1), e2 = __fadd2_rn(fi[h], nf2), e3 = __fadd2_rn(fi[h], nf3), e4 = __fadd2_rn(fi[h], nf4);
                float2 q1 = __ffma2_rn(e1, e1, B1), q2 = __ffma2_rn(e2, e2, B2), q3 = __ffma2_rn(e3, e3, B3), q4 = __ffma2_rn(e4, e4, B4);
                float2 n12 = __ffma2_rn(A2, q1, __fmul2_rn(A1, q2)), n34 = __ffma2_rn(A4, q3, __fmul2_rn(A3, q4));
                float2 p12 = __fmul2_rn(q1, q2), p34 = __fmul2_rn(q3, q4);
                float2 num = __ffma2_rn(n34, p12, __fmul2_rn(n12, p34));
                float2 den = __fmul2_rn(p12, p34);
                float2 r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
                acc[h] = __ffma2_rn(num, r, acc[h]);
            } else if (MODE == 1) {
                fi[h].x = rcp_approx(fi[h].x); fi[h].y = rcp_approx(fi[h].y);
            } else {
                float e1x = fi[h].x + nf1.x, e1y = fi[h].y + nf1.x, e2x = fi[h].x + nf2.x, e2y = fi[h].y + nf2.x;
                float q1x = fmaf(e1x, e1x, B1.x), q1y = fmaf(e1y, e1y, B1.x), q2x = fmaf(e2x, e2x, B2.x), q2y = fmaf(e2y, e2y, B2.x);
                float nx = fmaf(A2.x, q1x, A1.x * q2x), ny = fmaf(A2.x, q1y, A1.x * q2y);
                acc[h].x = fmaf(nx, rcp_approx(q1x * q2x), acc[h].x); acc[h].y = fmaf(ny, rcp_approx(q1y * q2y), acc[h].y);
            }
        }
        if (MODE != 1) { nf1 = __fadd2_rn(nf1, A1); nf2 = __fadd2_rn(nf2, A2); B1 = __fadd2_rn(B1, A1); B2 = __fadd2_rn(B2, A2); }
        if (MODE == 4) { nf3 = __fadd2_rn(nf3, A1); nf4 = __fadd2_rn(nf4, A2); B3 = __fadd2_rn(B3, A1); B4 = __fadd2_rn(B4, A2); }   // keep every operand loop-variant
    }
    long long t1 = clock64();
    float s = 0; for (int h = 0; h < 4; ++h) s += acc[h].x + acc[h].y + fi[h].x + fi[h].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(int ctas_per_sm, int iters) {
    int sms = 148, grid = sms * ctas_per_sm;
    float *out; long long *cyc; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    body<MODE><<<grid, 256>>>(out, 100, 1.5f, cyc); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); body<MODE><<<grid, 256>>>(out, iters, 1.5f, cyc); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= grid;
    // warps per SMSP = ctas_per_sm * 8 / 4; iterations per SMSP = iters * warps_per_smsp
    double wps = ctas_per_sm * 2.0;
    if (MODE == 4) printf("  (mode 4 covers 4 lines: halve for comparison with mode 2)\n");
    printf("mode %d ctas/SM %d (%.0f warps/SMSP): %.3f ms, %.1f cycles per warp-iteration per SMSP (clock64 avg %.0f)\n", MODE, ctas_per_sm, wps,
           ms, avg / (iters * wps), avg);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>(4, 200000);   // warm the clocks
    for (int c : {1, 2, 3, 4}) { run<0>(c, 20000); run<2>(c, 20000); run<4>(c, 20000); }
    return 0;
}

// Microbenchmark (development aid): can the FP64 pipe of a B200 SM sub-partition run beside the packed-FP32 pipe?
// Reports cycles per warp-instruction per SMSP for: FFMA2 alone, DFMA alone, FFMA2+DFMA interleaved (2:1 and 1:1),
// MUFU.RCP, MUFU.RCP64H, and FFMA2+DFMA+MUFU together.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__device__ __forceinline__ float rcp_approx(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double rcp64h(double x) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }

// NF packed FFMA2, ND DFMA, NM MUFU.RCP, NH MUFU.RCP64H per inner step (8 independent chains each)
template <int NF, int ND, int NM, int NH>
__global__ void __launch_bounds__(256) body(float *out, int iters, float seed, long long *cyc) {
    float2 x[8]; double d[8]; float m[8]; double hh[8];
    for (int h = 0; h < 8; ++h) {
        x[h] = make_float2(seed + threadIdx.x + h, seed * 0.5f + h);
        d[h] = seed + threadIdx.x * 0.25 + h; m[h] = seed + h + threadIdx.x; hh[h] = seed * 3 + h + threadIdx.x;
    }
    const float2 a = make_float2(1.0000001f, 0.9999999f), b = make_float2(1e-7f, -1e-7f);
    const double da = 1.0000000001, db = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
#pragma unroll
            for (int r = 0; r < NF; ++r) x[h] = __ffma2_rn(x[h], a, b);
#pragma unroll
            for (int r = 0; r < ND; ++r) d[h] = fma(d[h], da, db);
#pragma unroll
            for (int r = 0; r < NM; ++r) m[h] = rcp_approx(m[h]);
#pragma unroll
            for (int r = 0; r < NH; ++r) hh[h] = rcp64h(hh[h]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int h = 0; h < 8; ++h) s += x[h].x + x[h].y + (float)d[h] + m[h] + (float)hh[h];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NF, int ND, int NM, int NH> void run(int ctas_per_sm, int iters) {
    int grid = 148 * ctas_per_sm;
    float *out; long long *cyc; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    body<NF, ND, NM, NH><<<grid, 256>>>(out, 100, 1.5f, cyc); cudaDeviceSynchronize();
    body<NF, ND, NM, NH><<<grid, 256>>>(out, iters, 1.5f, cyc); cudaDeviceSynchronize();
    std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= grid;
    double wps = ctas_per_sm * 2.0;
    printf("FFMA2 x%d DFMA x%d RCP x%d RCP64H x%d | %d warps/SMSP: %.2f cycles per inner step (8 steps/iter) per warp per SMSP\n", NF, ND, NM, NH,
           (int)wps, avg / (iters * wps * 8.0));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<4, 0, 0, 0>(4, 100000);   // warm the clocks
    for (int c : {2, 4}) {
        run<4, 0, 0, 0>(c, 20000);
        run<0, 4, 0, 0>(c, 20000);
        run<4, 2, 0, 0>(c, 20000);
        run<4, 4, 0, 0>(c, 20000);
        run<0, 0, 1, 0>(c, 20000);
        run<0, 0, 0, 1>(c, 20000);
        run<4, 0, 1, 0>(c, 20000);
        run<4, 2, 1, 0>(c, 20000);
        run<4, 2, 1, 1>(c, 20000);
        run<6, 0, 1, 0>(c, 20000);
        run<6, 3, 1, 0>(c, 20000);
    }
    return 0;
}

// Microbenchmark (development aid): issue cost of the packed FP32x2 instructions on one SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <int MODE>
__global__ void __launch_bounds__(256) body(float *out, int iters, float seed, long long *cyc) {
    float2 x[16];
    for (int h = 0; h < 16; ++h) x[h] = make_float2(seed + threadIdx.x + h, seed * 0.5f + h);
    const float2 a = make_float2(1.0000001f, 0.9999999f), b = make_float2(1e-7f, -1e-7f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int h = 0; h < 16; ++h) {
                if (MODE == 0) x[h] = __ffma2_rn(x[h], a, b);
                if (MODE == 1) x[h] = __fmul2_rn(x[h], a);
                if (MODE == 2) x[h] = __fadd2_rn(x[h], b);
                if (MODE == 3) { x[h].x = fmaf(x[h].x, a.x, b.x); x[h].y = fmaf(x[h].y, a.y, b.y); }
            }
    }
    long long t1 = clock64();
    float s = 0; for (int h = 0; h < 16; ++h) s += x[h].x + x[h].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int ctas_per_sm, int iters, const char *name) {
    int grid = 148 * ctas_per_sm;
    float *out; long long *cyc; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    body<MODE><<<grid, 256>>>(out, 100, 1.5f, cyc); cudaDeviceSynchronize();
    body<MODE><<<grid, 256>>>(out, iters, 1.5f, cyc); cudaDeviceSynchronize();
    std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= grid;
    double wps = ctas_per_sm * 2.0;
    printf("%-8s %d warps/SMSP: %.2f cycles per instruction per SMSP\n", name, (int)wps, avg / (iters * wps * 32.0 * (MODE == 3 ? 2 : 1)));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int c : {2, 4}) { run<0>(c, 20000, "FFMA2"); run<1>(c, 20000, "FMUL2"); run<2>(c, 20000, "FADD2"); run<3>(c, 20000, "FFMA"); }
    return 0;
}

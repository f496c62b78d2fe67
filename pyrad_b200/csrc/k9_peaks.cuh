// k9_peaks.cuh -- the roofline denominators, measured on the device the engine runs on (prb_measure_peaks).
//
// bench.py reports K2 against the FP32 pipe and K1 / K3 against HBM.  MEASURED_PEAKS.json (driver-written) carries a
// copy bandwidth and a bf16 GEMM rate but no FP32 figure, so the library measures its own: a kernel that does nothing
// but issue packed FP32x2 fused multiply-adds (FFMA2, the instruction K2's triple-reciprocal loop is made of) from
// eight independent register chains per thread, and a plain float4 copy.  Neither touches product data.
#pragma once
#include "common.cuh"

namespace prb {

constexpr int K9_CHAINS = 8;       // independent FFMA2 chains per thread (hides the 4-cycle dependent-issue latency)
constexpr int K9_UNROLL = 16;      // FFMA2 per chain per loop trip

__global__ void __launch_bounds__(256) k9_ffma2_peak(float *out, int iters, float seed) {
    float2 x[K9_CHAINS];
#pragma unroll
    for (int h = 0; h < K9_CHAINS; ++h) x[h] = make_float2(seed + threadIdx.x + h, seed * 0.5f + h);
    // loop-variant multiplier and addend: with invariant operands ptxas may fold steps of the recurrence
    float2 a = make_float2(1.0000001f, 0.9999999f), b = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < K9_UNROLL; ++r) {
#pragma unroll
            for (int h = 0; h < K9_CHAINS; ++h) x[h] = __ffma2_rn(x[h], a, b);
        }
        a.x += 1e-9f;
        b.y -= 1e-9f;
    }
    float s = 0;
#pragma unroll
    for (int h = 0; h < K9_CHAINS; ++h) s += x[h].x + x[h].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same count of scalar FFMA (one lane-FMA each): which of the two forms is the pipe's peak is a measurement
__global__ void __launch_bounds__(256) k9_ffma_peak(float *out, int iters, float seed) {
    float x[2 * K9_CHAINS];
#pragma unroll
    for (int h = 0; h < 2 * K9_CHAINS; ++h) x[h] = seed + threadIdx.x + h;
    float a = 1.0000001f, b = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < K9_UNROLL; ++r) {
#pragma unroll
            for (int h = 0; h < 2 * K9_CHAINS; ++h) x[h] = fmaf(x[h], a, b);
        }
        a += 1e-9f;
        b -= 1e-9f;
    }
    float s = 0;
#pragma unroll
    for (int h = 0; h < 2 * K9_CHAINS; ++h) s += x[h];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k9_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __ldcs(src + i);
}

}  // namespace prb

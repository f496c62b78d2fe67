// Microbenchmark (development aid): formulations of K2's far-path inner body, one SM sub-partition's view.
// Every mode walks 1536 staged line records in shared memory (as the real kernel does) with H = 4 packed point
// pairs per thread, and reports cycles per (line, packed point pair) per warp per SMSP -- lower is better;
// 8 consumer warps per CTA, 2 CTAs per SM like k2_line_sum<8>.
//   mode 0: current triple reciprocal      A1/q1 + A2/q2 + A3/q3, q = (fi - c)^2 + B           13 packed + 2 MUFU / triple
//   mode 1: normalised triple              1/(a e^2 + b) with e~ = fma(fi, s, cs), q~ = fma(e~, e~, beta)   11 packed + 2 MUFU
//   mode 2: normalised pair                7 packed + 2 MUFU per two lines
//   mode 3: mode 0 in scalar FP32 (no packed instructions)
//   mode 4: mode 1 with the fourth point pair evaluated on the FP64 pipe (DFMA + RCP64H + one Newton step)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double rcp64h(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ float2 lo2(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 &v) { return make_float2(v.z, v.w); }

constexpr int NL = 1536;
struct Smem { float4 a[NL]; float4 b[NL]; };

template <int MODE>
__global__ void __launch_bounds__(256, 2) body(float *out, int reps, float seed, long long *cyc) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw);
    for (int j = threadIdx.x; j < NL; j += blockDim.x) {
        const float c = 3000.f + 7.f * j + seed;
        sm.a[j] = make_float4(-c, -c, 1.f + 0.001f * j, 1.f + 0.001f * j);              // {-c,-c,A,A}  | mode 1/2: {s,s,cs,cs}
        sm.b[j] = make_float4(50.f + j, 50.f + j, 0.f, 0.f);                              // {B,B,..}     | mode 1/2: {beta,beta}
        if (MODE == 1 || MODE == 2 || MODE == 4) {
            const float s = 1.0f / (30.f + 0.01f * j);
            sm.a[j] = make_float4(s, s, -c * s, -c * s);
            sm.b[j] = make_float4(0.5f + 0.001f * j, 0.5f + 0.001f * j, 0.f, 0.f);
        }
    }
    __syncthreads();
    float2 fi[4], acc[4];
    double dacc = 0.0, dfi = (double)(seed + threadIdx.x + 64 * 3);
    for (int h = 0; h < 4; ++h) { fi[h] = make_float2(seed + threadIdx.x + 64 * h, seed + threadIdx.x + 64 * h + 32); acc[h] = make_float2(0.f, 0.f); }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0 || MODE == 3) {
            for (int j = 0; j < NL; j += 3) {
                const float4 a1 = sm.a[j], a2 = sm.a[j + 1], a3 = sm.a[j + 2];
                const float2 B1 = lo2(sm.b[j]), B2 = lo2(sm.b[j + 1]), B3 = lo2(sm.b[j + 2]);
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (MODE == 0) {
                        const float2 e1 = __fadd2_rn(fi[h], lo2(a1)), e2 = __fadd2_rn(fi[h], lo2(a2)), e3 = __fadd2_rn(fi[h], lo2(a3));
                        const float2 q1 = __ffma2_rn(e1, e1, B1), q2 = __ffma2_rn(e2, e2, B2), q3 = __ffma2_rn(e3, e3, B3);
                        const float2 p23 = __fmul2_rn(q2, q3);
                        const float2 t = __ffma2_rn(hi2(a3), q2, __fmul2_rn(hi2(a2), q3));
                        const float2 num = __ffma2_rn(q1, t, __fmul2_rn(hi2(a1), p23));
                        const float2 den = __fmul2_rn(q1, p23);
                        acc[h] = __ffma2_rn(num, make_float2(rcp_approx(den.x), rcp_approx(den.y)), acc[h]);
                    } else {
                        float r2[2];
                        const float f2[2] = {fi[h].x, fi[h].y};
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const float e1 = f2[k] + a1.x, e2 = f2[k] + a2.x, e3 = f2[k] + a3.x;
                            const float q1 = fmaf(e1, e1, B1.x), q2 = fmaf(e2, e2, B2.x), q3 = fmaf(e3, e3, B3.x);
                            const float p23 = q2 * q3;
                            const float t = fmaf(a3.z, q2, a2.z * q3);
                            const float num = fmaf(q1, t, a1.z * p23);
                            r2[k] = num * rcp_approx(q1 * p23);
                        }
                        acc[h].x += r2[0]; acc[h].y += r2[1];
                    }
                }
            }
        } else if (MODE == 1 || MODE == 4) {
            for (int j = 0; j < NL; j += 3) {
                const float4 a1 = sm.a[j], a2 = sm.a[j + 1], a3 = sm.a[j + 2];
                const float2 B1 = lo2(sm.b[j]), B2 = lo2(sm.b[j + 1]), B3 = lo2(sm.b[j + 2]);
#pragma unroll
                for (int h = 0; h < (MODE == 4 ? 3 : 4); ++h) {
                    const float2 e1 = __ffma2_rn(fi[h], lo2(a1), hi2(a1)), e2 = __ffma2_rn(fi[h], lo2(a2), hi2(a2)), e3 = __ffma2_rn(fi[h], lo2(a3), hi2(a3));
                    const float2 q1 = __ffma2_rn(e1, e1, B1), q2 = __ffma2_rn(e2, e2, B2), q3 = __ffma2_rn(e3, e3, B3);
                    const float2 p23 = __fmul2_rn(q2, q3);
                    const float2 s23 = __fadd2_rn(q2, q3);
                    const float2 num = __ffma2_rn(q1, s23, p23);
                    const float2 den = __fmul2_rn(q1, p23);
                    acc[h] = __ffma2_rn(num, make_float2(rcp_approx(den.x), rcp_approx(den.y)), acc[h]);
                }
                if (MODE == 4) {     // one more point (not a pair) on the FP64 pipe
                    const double e1 = fma(dfi, (double)a1.x, (double)a1.z), e2 = fma(dfi, (double)a2.x, (double)a2.z), e3 = fma(dfi, (double)a3.x, (double)a3.z);
                    const double q1 = fma(e1, e1, (double)B1.x), q2 = fma(e2, e2, (double)B2.x), q3 = fma(e3, e3, (double)B3.x);
                    const double p23 = q2 * q3, s23 = q2 + q3;
                    const double num = fma(q1, s23, p23), den = q1 * p23;
                    double rr = rcp64h(den);
                    rr = fma(rr, fma(-den, rr, 1.0), rr);
                    dacc = fma(num, rr, dacc);
                }
            }
        } else if (MODE == 2) {
            for (int j = 0; j < NL; j += 2) {
                const float4 a1 = sm.a[j], a2 = sm.a[j + 1];
                const float2 B1 = lo2(sm.b[j]), B2 = lo2(sm.b[j + 1]);
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float2 e1 = __ffma2_rn(fi[h], lo2(a1), hi2(a1)), e2 = __ffma2_rn(fi[h], lo2(a2), hi2(a2));
                    const float2 q1 = __ffma2_rn(e1, e1, B1), q2 = __ffma2_rn(e2, e2, B2);
                    const float2 num = __fadd2_rn(q1, q2);
                    const float2 den = __fmul2_rn(q1, q2);
                    acc[h] = __ffma2_rn(num, make_float2(rcp_approx(den.x), rcp_approx(den.y)), acc[h]);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = (float)dacc; for (int h = 0; h < 4; ++h) s += acc[h].x + acc[h].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int reps) {
    int grid = 148 * 2;
    float *out; long long *cyc; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    cudaFuncSetAttribute(body<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    body<MODE><<<grid, 256, sizeof(Smem)>>>(out, 2, 1.5f, cyc); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); body<MODE><<<grid, 256, sizeof(Smem)>>>(out, reps, 1.5f, cyc); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += v; avg /= grid;
    // 16 warps per SM = 4 per SMSP; (line, point pair) units per warp = reps * NL * 4 (mode 4: 3.5)
    const double units = (double)reps * NL * (MODE == 4 ? 3.5 : 4.0);
    const double pairs = (double)grid * 256 * reps * NL * (MODE == 4 ? 7.0 : 8.0);
    printf("mode %d: %.3f ms  %.2f cycles per (line, packed point pair) per warp per SMSP   %.3e pairs/s  %s\n", MODE, ms, avg / (units * 4.0),
           pairs / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>(200);
    run<0>(100); run<1>(100); run<2>(100); run<3>(100); run<4>(100);
    return 0;
}

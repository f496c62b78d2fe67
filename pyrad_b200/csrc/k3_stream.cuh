// k3_stream.cuh -- K3: fused pointwise layer physics (HBM-bound streaming kernels).
//
//   absCoef        pyradClasses.py:581-583, 707-712   k = sum_m sigma_m * (conc_m P / 1e4 / kB / T)
//   transmittance  pyradClasses.py:585-587, 714-716   T = exp(-k * depth)
//   Planck         pyradPlanck.py:38-44, 12-15        B = 2e8 h c^2 nu^3 / (exp(100 h c nu / kB / T) - 1)
//   transmission   pyradClasses.py:784-787            I_out = T * I_in + (1 - T) * B(nu, T_layer)
//   xsc ingest     pyradClasses.py:159-162, 493-500   np.interp + aligned placement
//   multi-layer    fold of Layer.transmission over the layers (absent upstream)
//
// Two flavours: an FP64 kernel behind the host-buffer API (bit-faithful formulas, used by the
// pyradClasses mirror), and the FP32 float4-vectorised fold over the resident k matrix used by
// the atmosphere path (algorithmic traffic L*N*4 + N*8 bytes).
#pragma once
#include "common.cuh"

namespace prb {

// nu_i of np.linspace(start, stop, n): i*step + start (two roundings, as numpy), last = stop.
__device__ __forceinline__ double axis_value(int64_t i, int64_t n, double x0, double dx, double x_last) {
    if (i == n - 1) return x_last;
    return __dadd_rn(__dmul_rn((double)i, dx), x0);
}

__device__ __forceinline__ double planck_f64(double nu, double temp) {
    const double a = 2E8 * hPlanck * (cLight * cLight) * (nu * nu * nu);
    const double b = 100 * hPlanck * cLight * nu / kBoltz / temp;
    return a / (exp(b) - 1);                              // nu = 0 -> 0/0 = NaN, as the reference
}

__global__ void __launch_bounds__(256)
k3_layer_stream_f64(int64_t n, int n_mol, const double *__restrict__ sigma, const double *__restrict__ weight,
                    double depth, double t_layer, double x0, double dx, double x_last,
                    const double *__restrict__ rad_in, double *__restrict__ absc, double *__restrict__ trans,
                    double *__restrict__ rad_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double k = 0.0;
        for (int m = 0; m < n_mol; ++m) k += sigma[(int64_t)m * n + i] * weight[m];
        const double t = exp(-k * depth);
        if (absc) absc[i] = k;
        if (trans) trans[i] = t;
        if (rad_out) {
            const double b = planck_f64(axis_value(i, n, x0, dx, x_last), t_layer);
            rad_out[i] = t * rad_in[i] + (1 - t) * b;
        }
    }
}

// The same layer physics on rows that are already on the device: the per-group cross sections of the last
// prb_line_sum_groups (rows_a) and the resident xsc tables (rows_b), on the owned chunk [i_first, i_first + n) of a grid
// of n_total points.  k = sum_a sigma_a w_a + sum_b sigma_b w_b, in that order (Layer.absCoef, pyradClasses.py:707-712).
__global__ void __launch_bounds__(256)
k3_layer_stream_rows_f64(int64_t n, int n_a, const double *__restrict__ rows_a, int64_t ld_a,
                         const double *__restrict__ w_a, int n_b, const double *__restrict__ rows_b, int64_t ld_b,
                         const double *__restrict__ w_b, double depth, double t_layer, int64_t i_first, int64_t n_total,
                         double x0, double dx, double x_last, const double *__restrict__ rad_in,
                         double *__restrict__ absc, double *__restrict__ trans, double *__restrict__ rad_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double k = 0.0;
        for (int m = 0; m < n_a; ++m) k += rows_a[(int64_t)m * ld_a + i] * w_a[m];
        for (int m = 0; m < n_b; ++m) k += rows_b[(int64_t)m * ld_b + i] * w_b[m];
        const double t = exp(-k * depth);
        if (absc) absc[i] = k;
        if (trans) trans[i] = t;
        if (rad_out) {
            const double b = planck_f64(axis_value(i_first + i, n_total, x0, dx, x_last), t_layer);
            rad_out[i] = t * rad_in[i] + (1 - t) * b;
        }
    }
}

__global__ void __launch_bounds__(256)
k3_planck_f64(int64_t n, double x0, double dx, double x_last, double temp, double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = planck_f64(axis_value(i, n, x0, dx, x_last), temp);
}

// np.interp(x, xp, fp) for one x: clamps outside, slope form inside (numpy's arithmetic, no FMA).
__device__ __forceinline__ double np_interp(double x, const double *__restrict__ xp, const double *__restrict__ fp,
                                            int64_t n) {
    if (x <= xp[0]) return x < xp[0] ? fp[0] : fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    int64_t lo = 0, hi = n - 1;                            // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    const double slope = __ddiv_rn(__dsub_rn(fp[lo + 1], fp[lo]), __dsub_rn(xp[lo + 1], xp[lo]));
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[lo])), fp[lo]);
}

// out[0 .. n_out) holds the grid points [i_first, i_first + n_out) of the placement (i_first = 0: the whole grid; the
// resident tables of the atmosphere path are built for the owned chunk only).
__global__ void __launch_bounds__(256)
k3_xsc_place(int64_t n_out, int64_t dst0, int64_t src0, int64_t count, int interp, double ax0, double adelta,
             int64_t n_file, const double *__restrict__ fx, const double *__restrict__ fy, double *__restrict__ out,
             int64_t i_first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i_first + i - dst0;
        double v = 0.0;
        if (j >= 0 && j < count) {
            const int64_t m = src0 + j;
            if (interp) v = np_interp(__dadd_rn(ax0, __dmul_rn((double)m, adelta)), fx, fy, n_file);
            else v = (m >= 0 && m < n_file) ? fy[m] : 0.0;
        }
        out[i] = v;
    }
}

constexpr int K3_UNROLL = 8;

// Optical depth of the whole column: the per-layer terms -tau_l log2(e) are summed in FP32 over groups of
// K3_TAU_GROUP consecutive layers (aligned to the layer index, so every fold kernel forms the same groups) and the
// group sums are added into a double-float (hi, lo) pair with an error-free TwoSum.  A plain FP32 running sum over
// 100 layers loses up to ~100 * 2^-24 of tau, i.e. up to 2e-6 of the total transmittance at tau ~ 1 -- more than the
// 1e-6 the path promises; this way the error stays below 4 * 2^-24 * tau (|dT| <= 9e-8) for six FP32 instructions per
// point per four layers (FP64 accumulators would cost the fold kernel eight registers it does not have).
constexpr int K3_TAU_GROUP = 4;
__device__ __forceinline__ void tau_flush(float (&tau)[4], float (&hi)[4], float (&lo)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float s = __fadd_rn(hi[q], tau[q]);
        const float bp = __fsub_rn(s, hi[q]);
        const float err = __fadd_rn(__fsub_rn(hi[q], __fsub_rn(s, bp)), __fsub_rn(tau[q], bp));
        lo[q] = __fadd_rn(lo[q], err);
        hi[q] = s;
        tau[q] = 0.f;
    }
}

// ---- atmosphere fold, FP32 storage, float4 per thread --------------------------------------------
struct FoldLayer {
    float neg_depth_log2e;   // -depth * log2(e): T = exp2(k * this)
    float c2_over_t;         // 100 h c / kB / T_layer
};

__device__ __forceinline__ float k3_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float k3_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float planck_f32(float a_nu3, float x) {
    // a nu^3 / (e^x - 1), x = c2 nu / T, to ~1e-7 relative with as few MUFU operations as the range allows:
    //   x >  8.5 : y = e^-x < 2.1e-4, 1/(e^x - 1) = y/(1 - y) = y + y^2 + O(y^3)       -- one MUFU (ex2)
    //   x >= .25 : rcp(e^x - 1)                                                        -- two MUFU (ex2, rcp)
    //   x <  .25 : the subtraction cancels, so the first few hundred grid points (nu < ~40 cm^-1) take expm1f.
    if (x > 8.5f) {
        const float y = k3_ex2(x * -1.4426950408889634f);
        return a_nu3 * fmaf(y, y, y);
    }
    if (x < 0.25f) return a_nu3 / expm1f(x);
    return a_nu3 * k3_rcp(k3_ex2(x * 1.4426950408889634f) - 1.0f);
}

// One layer of the fold: I <- T I + (1 - T) B with T = 2^e, e = -tau log2(e).  Two evaluation orders, chosen per
// point and layer: where the layer is optically thin (|e| < 1/8, T > 0.917) the emission weight 1 - T comes from the
// series of -expm1(e ln 2) and the update is I + (1 - T)(B - I); elsewhere it is B + T (I - B).  With T alone the
// thin case computes 1 - T by cancellation from a MUFU.EX2 result that is good to ~1e-7 ABSOLUTE: a warm, nearly
// transparent layer above a cold opaque one (the stratopause seen at 4000 cm^-1: B_layer = 300 I, tau = 1e-4) then
// adds 300 * 1e-7 of I per layer, all of one sign -- 2.5e-5 of the radiance after 30 such layers (measured against
// the oracle at full size, tests/test_gpu_fullsize.py).  The thick-layer form must stay for T -> 0: I + (B - I)
// would round at the size of I, not of B.  Series: five terms, truncation below 7e-9 of 1 - T at |e| = 1/8.
__device__ __forceinline__ float k3_fold_step(float rad, float e, float b) {
    const float t = k3_ex2(e);
    // 1 - 2^e = -e (c1 + e (c2 + e (c3 + e (c4 + e c5)))),  c_n = ln(2)^n / n!
    float p = fmaf(e, 1.3333558146e-3f, 9.6181291076e-3f);
    p = fmaf(e, p, 5.5504108665e-2f);
    p = fmaf(e, p, 2.4022650696e-1f);
    p = fmaf(e, p, 6.9314718056e-1f);
    const bool thin = fabsf(e) < 0.125f;
    const float d = rad - b;
    return thin ? fmaf(e * p, d, rad) : fmaf(t, d, b);     // thin: I - (1 - T)(I - B), (1 - T) = -e p
}

// The same step for two points at once with Blackwell's packed FP32x2 instructions (the fold is issue bound: half the
// instructions for the multiply, the series, the difference, both candidate updates).  Lane for lane the operations and
// their roundings are those of k3_fold_step, so the two agree bit for bit.
__device__ __forceinline__ float2 k3_fold_step2(float2 rad, float2 e, float2 b) {
    const float2 t = make_float2(k3_ex2(e.x), k3_ex2(e.y));
    float2 p = __ffma2_rn(e, make_float2(1.3333558146e-3f, 1.3333558146e-3f), make_float2(9.6181291076e-3f, 9.6181291076e-3f));
    p = __ffma2_rn(e, p, make_float2(5.5504108665e-2f, 5.5504108665e-2f));
    p = __ffma2_rn(e, p, make_float2(2.4022650696e-1f, 2.4022650696e-1f));
    p = __ffma2_rn(e, p, make_float2(6.9314718056e-1f, 6.9314718056e-1f));
    const float2 d = __ffma2_rn(b, make_float2(-1.f, -1.f), rad);              // rad - b, one rounding
    const float2 thin = __ffma2_rn(__fmul2_rn(e, p), d, rad);
    const float2 thick = __ffma2_rn(t, d, b);
    return make_float2(fabsf(e.x) < 0.125f ? thin.x : thick.x, fabsf(e.y) < 0.125f ? thin.y : thick.y);
}

// Destinations of the finished spectra: this rank's slot in every rank's gather buffer (peer memory over
// NVLink; stores are fire-and-forget), or just the local result arrays when no peers are connected.
constexpr int K3_MAX_PEERS = 8;
constexpr int K3_MAX_DST = K3_MAX_PEERS + 1;          // + the caller's pinned host result buffers (zero-copy)
struct K3Dst {
    int n;
    float *rad[K3_MAX_DST];
    float *trans[K3_MAX_DST];
};

__global__ void __launch_bounds__(256)
k3_fold_f32(const float *__restrict__ kmat, int64_t ld, int n_layers, const FoldLayer *__restrict__ layers,
            int64_t n_chunk, int64_t i_begin, int64_t n_total, double x0, double dx, double x_last,
            float c2_over_tsurf, const K3Dst dst) {
    const int64_t nvec = (n_chunk + 3) >> 2;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = v << 2;
        float nu[4], a3[4], rad[4], tau[4];
        float tau_hi[4], tau_lo[4];                              // optical depth of the column, see K3_TAU_GROUP
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double x = axis_value(i_begin + i0 + q, n_total, x0, dx, x_last);
            nu[q] = (float)x;
            a3[q] = (float)(2E8 * hPlanck * (cLight * cLight) * (x * x * x));
            rad[q] = planck_f32(a3[q], c2_over_tsurf * nu[q]);   // I_0 = B(nu, T_surface)
            tau[q] = 0.f;
            tau_hi[q] = 0.f;
            tau_lo[q] = 0.f;
        }
        // layers in groups of K3_UNROLL: all of a group's loads are issued before its math, so every thread keeps
        // K3_UNROLL x 16 B in flight (the kernel is latency bound otherwise); the k matrix is read exactly once,
        // hence the streaming (evict-first) loads.
        const float *col = kmat + i0;
        for (int l0 = 0; l0 < n_layers; l0 += K3_UNROLL) {
            float4 k4[K3_UNROLL];
#pragma unroll
            for (int j = 0; j < K3_UNROLL; ++j)
                k4[j] = (l0 + j < n_layers) ? __ldcs(reinterpret_cast<const float4 *>(col + (int64_t)(l0 + j) * ld))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < K3_UNROLL; ++j) {
                if (l0 + j < n_layers) {
                    const FoldLayer fl = layers[l0 + j];
                    const float kk[4] = {k4[j].x, k4[j].y, k4[j].z, k4[j].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float e = kk[q] * fl.neg_depth_log2e;       // -tau_l * log2(e)
                        const float b = planck_f32(a3[q], fl.c2_over_t * nu[q]);
                        rad[q] = k3_fold_step(rad[q], e, b);              // T*I + (1-T)*B
                        tau[q] += e;
                    }
                    if (((l0 + j) & (K3_TAU_GROUP - 1)) == K3_TAU_GROUP - 1) tau_flush(tau, tau_hi, tau_lo);
                }
            }
        }
        tau_flush(tau, tau_hi, tau_lo);
        const float4 r4 = make_float4(rad[0], rad[1], rad[2], rad[3]);
        const float4 t4 = make_float4(exp2f(tau_hi[0] + tau_lo[0]), exp2f(tau_hi[1] + tau_lo[1]), exp2f(tau_hi[2] + tau_lo[2]),
                                      exp2f(tau_hi[3] + tau_lo[3]));
        const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll 1
        for (int d = 0; d < dst.n; ++d) {
            if (i0 + 3 < n_chunk) {
                *reinterpret_cast<float4 *>(dst.rad[d] + i0) = r4;
                *reinterpret_cast<float4 *>(dst.trans[d] + i0) = t4;
            } else {
                for (int q = 0; q < 4 && i0 + q < n_chunk; ++q) {
                    dst.rad[d][i0 + q] = rr[q];
                    dst.trans[d][i0 + q] = tt[q];
                }
            }
        }
    }
}

// ---- the same fold with the k matrix staged through shared memory by TMA ----------------------------------------
// A CTA owns strips of 1024 grid points (one float4 per thread) and walks the layers in groups of K3T_LAYERS rows:
// one elected thread issues a bulk copy per row (4 KB each) into a small ring, one group ahead of the math, so the
// loads in flight (5 CTAs x 16 KB per SM) are not held in registers; the math reads its float4 from shared memory
// (conflict free).  Ring geometry measured on B200 (rows:slots:CTAs/SM -> ms at cfg4): 4:2:5 0.406, 4:3:4 0.424,
// 6:3:3 0.459, 4:4:3 0.463, 2:4:6 0.535, 8:3:2 0.537.  The fold is instruction bound rather than memory bound (measured: staging alone
// changes nothing), so above nu_interp_min the layer's Planck term is evaluated at the thread's first and last point
// only and interpolated linearly for the two between -- half the instructions per point; k3_fold_f32 keeps the
// per-point evaluation and is the reference the tests compare against.
constexpr int K3T_THREADS = 256;
constexpr int K3T_STRIP = K3T_THREADS * 4;             // points per strip
#ifndef PRB_K3T_LAYERS
#define PRB_K3T_LAYERS 4
#endif
#ifndef PRB_K3T_SLOTS
#define PRB_K3T_SLOTS 2
#endif
#ifndef PRB_K3T_MINB
#define PRB_K3T_MINB 5
#endif
constexpr int K3T_LAYERS = PRB_K3T_LAYERS;             // rows per ring slot
constexpr int K3T_SLOTS = PRB_K3T_SLOTS;
constexpr int K3T_MINB = PRB_K3T_MINB;                 // resident CTAs per SM the shared memory is sized for

struct K3TSmem {
    float k[K3T_SLOTS][K3T_LAYERS][K3T_STRIP];
    float2 tau[4][K3T_THREADS];            // (hi, lo) optical-depth accumulators of every thread's four points
    uint64_t full[K3T_SLOTS];
};

__global__ void __launch_bounds__(K3T_THREADS, K3T_MINB)
k3_fold_tma(const float *__restrict__ kmat, int64_t ld, int n_layers, const FoldLayer *__restrict__ layers,
            int64_t n_chunk, int64_t i_begin, int64_t n_total, double x0, double dx, double x_last,
            float c2_over_tsurf, float nu_interp_min, const K3Dst dst) {
    extern __shared__ __align__(128) unsigned char k3_smem_raw[];
    K3TSmem &sm = *reinterpret_cast<K3TSmem *>(k3_smem_raw);
    const int tid = threadIdx.x;
    const int64_t n_strips = (n_chunk + K3T_STRIP - 1) / K3T_STRIP;
    const int groups = (n_layers + K3T_LAYERS - 1) / K3T_LAYERS;
    // the CTA's work as one flat sequence of (strip, layer group) steps
    const int64_t my_strips = blockIdx.x < n_strips ? (n_strips - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_steps = my_strips * groups;
    if (tid == 0) {
        for (int s = 0; s < K3T_SLOTS; ++s) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t step) {                   // thread 0 only
        const int slot = (int)(step % K3T_SLOTS);
        const int64_t strip = blockIdx.x + (step / groups) * gridDim.x;
        const int g = (int)(step % groups);
        const int64_t p0 = strip * K3T_STRIP;
        const int64_t pts = min((int64_t)K3T_STRIP, ((n_chunk - p0) + 3) & ~int64_t(3));   // rows are padded to 4 floats
        const int rows = min(K3T_LAYERS, n_layers - g * K3T_LAYERS);
        mbar_expect_tx(&sm.full[slot], (uint32_t)(rows * pts * 4));
        for (int r = 0; r < rows; ++r)
            tma_bulk_g2s(sm.k[slot][r], kmat + (int64_t)(g * K3T_LAYERS + r) * ld + p0, (uint32_t)(pts * 4), &sm.full[slot]);
    };
    if (tid == 0)
        for (int64_t s = 0; s < min(n_steps, (int64_t)(K3T_SLOTS - 1)); ++s) issue(s);

    float nu[4], a3[4], rad[4], tau[4];
    // the column's (hi, lo) optical depth lives in shared memory, touched once per K3_TAU_GROUP layers: the register
    // budget of 5 CTAs per SM (48) has no room for it
    auto tau_flush_smem = [&]() {
        float hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 v = sm.tau[q][tid]; hi[q] = v.x; lo[q] = v.y; }
        tau_flush(tau, hi, lo);
#pragma unroll
        for (int q = 0; q < 4; ++q) sm.tau[q][tid] = make_float2(hi[q], lo[q]);
    };
    bool interp = false;
    int64_t i0 = 0;
    for (int64_t step = 0; step < n_steps; ++step) {
        const int slot = (int)(step % K3T_SLOTS);
        const int g = (int)(step % groups);
        if (g == 0) {                                  // a new strip: surface radiance, zero optical depth
            const int64_t strip = blockIdx.x + (step / groups) * gridDim.x;
            i0 = strip * K3T_STRIP + 4 * tid;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double x = axis_value(i_begin + i0 + q, n_total, x0, dx, x_last);
                nu[q] = (float)x;
                a3[q] = (float)(2E8 * hPlanck * (cLight * cLight) * (x * x * x));
                rad[q] = planck_f32(a3[q], c2_over_tsurf * nu[q]);
                tau[q] = 0.f;
                sm.tau[q][tid] = make_float2(0.f, 0.f);
            }
            interp = nu[0] >= nu_interp_min;
        }
        // keep the ring two steps ahead: the slot being refilled was released by the barrier at the end of step-1
        if (tid == 0 && step + K3T_SLOTS - 1 < n_steps) issue(step + K3T_SLOTS - 1);
        mbar_wait(&sm.full[slot], (uint32_t)((step / K3T_SLOTS) & 1));
        const int rows = min(K3T_LAYERS, n_layers - g * K3T_LAYERS);
        if (i0 < n_chunk) {
#pragma unroll
            for (int r = 0; r < K3T_LAYERS; ++r) {
                if (r < rows) {
                    const float4 k4 = *reinterpret_cast<const float4 *>(&sm.k[slot][r][4 * tid]);
                    const FoldLayer fl = layers[g * K3T_LAYERS + r];
                    const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
                    float b[4];
                    if (interp) {
                        // B(nu, T_l) at the thread's first and last point, the two between by linear interpolation
                        // (relative error < 1e-7 above nu_interp_min, chosen by the host from the grid spacing); one
                        // path decision per thread instead of per point
                        const float xa = fl.c2_over_t * nu[0], xb = fl.c2_over_t * nu[3];
                        float ba, bb;
                        if (xa > 8.5f) {
                            const float ya = k3_ex2(xa * -1.4426950408889634f), yb = k3_ex2(xb * -1.4426950408889634f);
                            ba = a3[0] * fmaf(ya, ya, ya);
                            bb = a3[3] * fmaf(yb, yb, yb);
                        } else {
                            ba = a3[0] * k3_rcp(k3_ex2(xa * 1.4426950408889634f) - 1.0f);
                            bb = a3[3] * k3_rcp(k3_ex2(xb * 1.4426950408889634f) - 1.0f);
                        }
                        const float db = bb - ba;
                        b[0] = ba; b[1] = fmaf(db, 1.0f / 3.0f, ba); b[2] = fmaf(db, 2.0f / 3.0f, ba); b[3] = bb;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) b[q] = planck_f32(a3[q], fl.c2_over_t * nu[q]);
                    }
                    {
                        const float2 nd = make_float2(fl.neg_depth_log2e, fl.neg_depth_log2e);
                        const float2 e01 = __fmul2_rn(make_float2(kk[0], kk[1]), nd), e23 = __fmul2_rn(make_float2(kk[2], kk[3]), nd);
                        const float2 r01 = k3_fold_step2(make_float2(rad[0], rad[1]), e01, make_float2(b[0], b[1]));
                        const float2 r23 = k3_fold_step2(make_float2(rad[2], rad[3]), e23, make_float2(b[2], b[3]));
                        const float2 t01 = __fadd2_rn(make_float2(tau[0], tau[1]), e01), t23 = __fadd2_rn(make_float2(tau[2], tau[3]), e23);
                        rad[0] = r01.x; rad[1] = r01.y; rad[2] = r23.x; rad[3] = r23.y;
                        tau[0] = t01.x; tau[1] = t01.y; tau[2] = t23.x; tau[3] = t23.y;
                    }
                    if (((g * K3T_LAYERS + r) & (K3_TAU_GROUP - 1)) == K3_TAU_GROUP - 1) tau_flush_smem();
                }
            }
        }
        __syncthreads();                               // every thread is done with this slot: it may be refilled
        if (g == groups - 1 && i0 < n_chunk) {
            tau_flush_smem();
            float tt4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float2 v = sm.tau[q][tid]; tt4[q] = exp2f(v.x + v.y); }
            const float4 r4 = make_float4(rad[0], rad[1], rad[2], rad[3]);
            const float4 t4 = make_float4(tt4[0], tt4[1], tt4[2], tt4[3]);
            const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll 1
            for (int d = 0; d < dst.n; ++d) {
                if (i0 + 3 < n_chunk) {
                    *reinterpret_cast<float4 *>(dst.rad[d] + i0) = r4;
                    *reinterpret_cast<float4 *>(dst.trans[d] + i0) = t4;
                } else {
                    for (int q = 0; q < 4 && i0 + q < n_chunk; ++q) {
                        dst.rad[d][i0 + q] = rr[q];
                        dst.trans[d][i0 + q] = tt[q];
                    }
                }
            }
        }
    }
}

// ---- cross-GPU completion barrier of one gather step ------------------------------------------------
// Launched (one warp) after the kernel that stored this rank's spectra into every peer's gather buffer.
// Lane p publishes "rank r finished epoch E" into peer p's flag block with a system-scope release, then
// waits until peer p's flag for E has arrived here: when the kernel retires, every rank's slot of the LOCAL
// gather buffer is complete.  A peer that never arrives trips the timeout and sets *err (reported by the host
// as PRB_ERR_PEER) instead of hanging the GPU.
struct PeerSignal {
    unsigned int *flags[K3_MAX_PEERS]; // every rank's flag block (own one included), [2][K3_MAX_PEERS] words each
    int rank, world;
    unsigned int epoch;
    unsigned int *err;
    unsigned long long timeout_ns;
};

__global__ void k_peer_signal_wait(const PeerSignal s) {
    const int p = threadIdx.x;
    if (p >= s.world) return;
    const int par = s.epoch & 1u;
    __threadfence_system();
    unsigned int *remote = s.flags[p] + par * K3_MAX_PEERS + s.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(s.epoch) : "memory");
    const unsigned int *mine = s.flags[s.rank] + par * K3_MAX_PEERS + p;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - s.epoch) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > s.timeout_ns) { atomicOr(s.err, 1u); break; }
        __nanosleep(200);
    }
}

}  // namespace prb

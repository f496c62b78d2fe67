// k4_derived.cuh -- the section 8(f) "next" rows that sit right beside the hot path, as device kernels:
//
//   line survey       pyradClasses.py:409-428, 589-594, 691-696   S296 binned at int((nu0 - rangeMin)/res)
//   integrateSpectrum pyradClasses.py:26-29                       nansum(spectrum) * unitAngle * res
//   derived spectra   pyradClasses.py:73-88, 330-340, 596-606     emissivity 1-T, optical depth -ln T,
//                                                                 absorbance log10(1/T)
#pragma once
#include "common.cuh"

namespace prb {

// Line survey.  The reference adds S296 line by line (ascending nu0) into lineSurvey[arrayIndex]; the index array
// is sorted, so the lines of one bin are adjacent: the thread that owns the FIRST line of a bin walks the bin and
// adds in the reference's own order (bit-exact, no atomics).  `out` is zero-filled by the caller.
__global__ void __launch_bounds__(256)
k4_line_survey(const int32_t *__restrict__ idx, const double *__restrict__ s296, int64_t n_lines, int64_t n_out,
               double *__restrict__ out) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const int32_t b = idx[l];
    if (b < 0 || (int64_t)b > n_out - 1) return;                  // isBetween(arrayIndex, 0, len - 1)
    if (l > 0 && idx[l - 1] == b) return;                         // not the head of its bin
    double acc = 0.0;                                             // lineSurvey[b] starts at 0.0
    for (int64_t j = l; j < n_lines && idx[j] == b; ++j) acc = __dadd_rn(acc, s296[j]);
    out[b] = acc;
}

// np.nan_to_num on one value: NaN -> 0, +-inf -> +-DBL_MAX
__device__ __forceinline__ double nan_to_num(double v) {
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}

// Deterministic two-level sum: every block reduces a fixed contiguous slice in a fixed tree order and writes one
// partial; k4_sum_final adds the partials in index order.  (np.sum is pairwise too; the two agree to ~1e-15.)
constexpr int K4_BLOCK = 256;
constexpr int K4_PER_THREAD = 16;

template <typename T>
__global__ void __launch_bounds__(K4_BLOCK)
k4_sum_partial(const T *__restrict__ x, int64_t n, double *__restrict__ partial) {
    __shared__ double sh[K4_BLOCK];
    const int64_t base = (int64_t)blockIdx.x * (K4_BLOCK * K4_PER_THREAD);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < K4_PER_THREAD; ++k) {
        const int64_t i = base + (int64_t)k * K4_BLOCK + threadIdx.x;
        if (i < n) acc += nan_to_num((double)x[i]);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = K4_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(K4_BLOCK)
k4_sum_final(const double *__restrict__ partial, int64_t n_partial, double scale, double *__restrict__ out) {
    __shared__ double sh[K4_BLOCK];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n_partial; i += K4_BLOCK) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = K4_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0] * scale;
}

// Derived spectra from a transmittance array, numpy's arithmetic order:
//   emissivity = 1 - T;  optical depth = -log(T);  absorbance = log10(1 / T)
__global__ void __launch_bounds__(256)
k4_derived_f64(int64_t n, const double *__restrict__ trans, double *__restrict__ emis, double *__restrict__ tau,
               double *__restrict__ absb) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = trans[i];
        if (emis) emis[i] = __dsub_rn(1.0, t);
        if (tau) tau[i] = -log(t);
        if (absb) absb[i] = log10(__ddiv_rn(1.0, t));
    }
}

}  // namespace prb

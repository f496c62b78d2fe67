// k1_prepass.cuh -- K0 (line -> grid index) and K1 (per-layer, per-line prepass), FP64.
//
// K1 restates, per line, in FP64 (one thread per line, SoA loads fully coalesced):
//   shifted nu          pyradClasses.py:252-254   nu* = nu0 + delta * P / p0
//   Lorentz half width  pyradClasses.py:256-259   ((1-q) g_air + q g_self) (P/p0) (t0/T)^n
//   Doppler half width  pyradClasses.py:261-263   nu* sqrt(2kT/m/c^2)          (1/e half width)
//   regime select       pyradClasses.py:378-387   ratio < .01 Gauss, > 100 Lorentz, else pseudo-Voigt
//   S(T)                pyradIntensity.py:16-32   S296 (Q296/QT) stim(nu*,T) boltz(E'',T)
//   pseudo-Voigt f, eta pyradLineshape.py:58-71
// and packs what K2 needs into FP32 records (36 B per line, see k2_line_sum.cuh).  With d = i - idx (grid units):
//   contribution(d) = A / (d^2 + B) + G * exp2(C * d^2)
//   Voigt  : h = f/2;  A = S eta h / pi / res^2;  B = (h/res)^2;  G = S (1-eta) / (h sqrt(pi));  C = -log2(e)/B
//   Lorentz: h = gL;   A = S h / pi / res^2;      B = (h/res)^2;  G = 0
//   Gauss  : h = gD;   A = 0, B = 1;              G = S / (h sqrt(pi));  C = -log2(e) res^2 / h^2
// which is algebraically the reference's S * shape(d * res) (pyradLineshape.py:32-76).
// A and G additionally carry the group weight and a power-of-two scale (FP32 range), undone
// exactly in K2's FP64 epilogue.
#pragma once
#include "common.cuh"

namespace prb {

struct LinesSoA {
    const double *nu0, *s296, *gair, *gself, *elower, *nair, *delta;
    const int32_t *group;   // may be nullptr (single group)
};

// Per-layer constants evaluated once on the host in FP64.
struct LayerConsts {
    double log_t0_over_t;          // log(296 / T)
    double inv_t_minus_inv_t0;     // 1/T - 1/296
    double inv_res2;               // 1 / res^2
    double res2;                   // res^2
    double p_over_p0;              // P / 1013.25
    double neg_c2_over_t;          // -c2 / T
    double neg_c2_over_t0;         // -c2 / 296
};

struct DebugOut {
    double *nu_shift, *gl, *gd, *st;
    int32_t *regime;
};

// K0: arrayIndex = int((nu0 - rangeMin) / res)   (pyradClasses.py:390; FP64 divide, truncation
// toward zero, UN-shifted nu0).  Saturated to int32; padding entries get INT32_MAX.
__global__ void k0_line_index(const double *__restrict__ nu0, int64_t n, int64_t l_begin, int64_t n_alloc,
                              double range_min, double res, int32_t *__restrict__ idx) {
    int64_t l = l_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_alloc) return;
    if (l >= n) { idx[l] = INT32_MAX; return; }
    double q = __ddiv_rn(__dsub_rn(nu0[l], range_min), res);
    double t = trunc(q);
    t = fmin(fmax(t, -2147483647.0), 2147483646.0);
    idx[l] = (int32_t)t;
}

// max |x| of a column as its bit pattern (|x| >= 0: bit order == value order), and -- for the line list checks the
// host does in prb_upload_lines -- "ascending nu0" / "group id in range" as device flags.
__global__ void __launch_bounds__(256)
k0_absmax(const double *__restrict__ x, int64_t n, unsigned long long *__restrict__ out_bits) {
    unsigned long long b = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = (unsigned long long)__double_as_longlong(fabs(x[i]));
        b = v > b ? v : b;
    }
    const unsigned int hi = __reduce_max_sync(0xffffffffu, (unsigned int)(b >> 32));
    const unsigned int lo = __reduce_max_sync(0xffffffffu, (unsigned int)(b >> 32) == hi ? (unsigned int)b : 0u);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, ((unsigned long long)hi << 32) | lo);
}

// The host's pick_scale (api.cu) on the device: scale = 2^(10 - ilogb(max|S| * max weight)), written where K1 reads
// it and, inverted, into the K2 table rows of the launches that will undo it.
__global__ void k0_pick_scale(const unsigned long long *__restrict__ smax_bits, double w_max, double *__restrict__ scale_out,
                              double *__restrict__ inv_scale_rows, int n_rows, int row_stride_doubles) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double m = __longlong_as_double((long long)*smax_bits) * w_max;
    double sc = 1.0;
    if (m > 0 && isfinite(m)) sc = ldexp(1.0, 10 - ilogb(m));
    *scale_out = sc;
    for (int r = 0; r < n_rows; ++r) inv_scale_rows[(size_t)r * row_stride_doubles] = 1.0 / sc;
}

// Number of (line, grid point) accumulations of the reference on the owned chunk [i_lo, i_hi] for the window |d| <= wm
// (the metric's numerator).  Integer sum: the atomics only combine per-warp partial counts.
__global__ void __launch_bounds__(256)
k0_pair_count(const int32_t *__restrict__ idx, int64_t n, int64_t wm, int64_t i_lo, int64_t i_hi,
              unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n; l += (int64_t)gridDim.x * blockDim.x) {
        const int64_t x = idx[l];
        const int64_t lo = max(x - wm, i_lo), hi = min(x + wm, i_hi);
        if (hi >= lo) c += (unsigned long long)(hi - lo + 1);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void __launch_bounds__(256)
k0_validate_lines(const double *__restrict__ nu0, const int32_t *__restrict__ group, int64_t n, int n_groups,
                  unsigned int *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int f = 0;
    if (i < n) {
        if (i + 1 < n && !(nu0[i + 1] >= nu0[i])) f |= 1u;
        if (group && (group[i] < 0 || group[i] >= n_groups)) f |= 2u;
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// Gap entries of a grouped line list (group < 0): a wavenumber that maps to the sentinel index, zero intensity.
__global__ void __launch_bounds__(256)
k0_fill_gaps(double *__restrict__ nu0, double *__restrict__ s296, const int32_t *__restrict__ group, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && group[i] < 0) { nu0[i] = __longlong_as_double(0x7ff0000000000000LL); s296[i] = 0.0; }
}

__device__ __forceinline__ double pow5(double x) { double x2 = x * x; return x2 * x2 * x; }

// One layer of a (possibly multi-layer) prepass launch.
struct K1Layer {
    double T, P;
    LayerConsts lc;
    double scale;                  // power-of-two scale of this layer's FP32 coefficients ...
    const double *scale_dev;       // ... or, when not NULL, where the device computed it (pipelined upload: no host sync)
    double wm;                     // W-2 clamped at 0 (FP32 range guard)
    const GroupParams *gp;         // n_groups entries for this layer
    float4 *recA, *recB;           // this layer's record arrays (indexed by line)
    float *recD;
    DevState *st;                  // this layer's status flags
    int narrow;                    // 1: compact records for k2_narrow
    int pad;
};

// The layer table travels as a kernel parameter: per-layer constants are then constant-bank operands of the
// FP64 instructions (no registers, no loads), which is what lets the register budget hold the line's own data.
constexpr int K1_MAX_LAYERS = 128;
struct K1Table {
    int n;
    int pad;
    K1Layer rows[K1_MAX_LAYERS];
};

// exp(x) for the arguments this kernel meets (|x| <= 700), coefficients in constant memory so that every DFMA takes
// its constant straight from the constant bank (the profile of the library exp showed 28 % of K1's instructions
// moving 64-bit immediates into registers).  x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor (remainder 4e-18),
// result scaled by adding k to the exponent field.  ~2 ulp; the parity tests hold K1's outputs to 1e-12.
__constant__ double K1_EXP[18] = {
    1.4426950408889634,            // [0] log2(e)
    6755399441055744.0,            // [1] 1.5 * 2^52: rounds to the nearest integer in the low word
    -6.93147180369123816490e-01,   // [2] -ln2 (high part)
    -1.90821492927058770002e-10,   // [3] -ln2 (low part)
    1.0 / 6227020800.0,            // [4] 1/13!
    1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
__device__ __forceinline__ double exp_k1(double x) {
    if (!(fabs(x) <= 700.0)) return exp(x);                       // out of range / NaN: the library routine
    const double t = fma(x, K1_EXP[0], K1_EXP[1]);
    const int k = __double2loint(t);
    const double kf = t - K1_EXP[1];
    double r = fma(kf, K1_EXP[2], x);
    r = fma(kf, K1_EXP[3], r);
    double p = K1_EXP[4];
#pragma unroll
    for (int i = 5; i < 18; ++i) p = fma(p, r, K1_EXP[i]);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// f5^(-1/5) without log/exp/division: FP32 seed (two MUFU) and two Newton steps on g(r) = r^-5 - f5,
//   r <- r (1 + (1 - f5 r^5)/5)   (quadratic: 1e-6 -> ~3e-12 -> below FP64 resolution).
__device__ __forceinline__ double inv_fifth_root(double f5) {
    double r = (double)__powf((float)f5, -0.2f);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double r2 = r * r;
        const double r5 = r2 * r2 * r;
        r = fma(r, fma(-f5, r5, 1.0) * 0.2, r);
    }
    return r;
}

// exp(x) for |x| <= 0.02 (degree-7 Taylor, truncation < 1e-18)
__device__ __forceinline__ double exp_tiny(double x) {
    double p = 1.0 / 5040;
    p = fma(p, x, 1.0 / 720);
    p = fma(p, x, 1.0 / 120);
    p = fma(p, x, 1.0 / 24);
    p = fma(p, x, 1.0 / 6);
    p = fma(p, x, 0.5);
    p = fma(p, x, 1.0);
    return fma(p, x, 1.0);
}

// One thread per line; the thread keeps the line's seven constants in registers and walks the layers of the
// batch, so the SoA columns are read once per launch instead of once per layer.
__global__ void __launch_bounds__(256, 3)
k1_prepass(LinesSoA L, const int32_t *__restrict__ idx, const __grid_constant__ K1Table tab,
           int64_t l_begin, int64_t l_end, int64_t n_lines, int64_t i_base, DebugOut dbg) {
    const double c2 = cLight * hPlanck * 100 / kBoltz;          // pyradIntensity.py:13
    const int64_t l = l_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = l < l_end;
    // (grouped line lists keep every group's segment 16-byte aligned: the few gap entries between segments carry
    // group -1 and become padding records like the entries past the end)
    const bool real = in_range && l < n_lines && !(L.group && L.group[l] < 0);
    double nu = 0, delta = 0, gair = 0, gself = 0, nair = 0, elower = 0, s296 = 0;
    float nf = 0.f;
    int g = 0;
    if (real) {
        nu = L.nu0[l]; delta = L.delta[l]; gair = L.gair[l]; gself = L.gself[l];
        nair = L.nair[l]; elower = L.elower[l]; s296 = L.s296[l];
        g = L.group ? L.group[l] : 0;
        nf = -(float)((double)((int64_t)idx[l] - i_base));
    }
    // exp(-c2 nu0 / t0): the layer-independent factor of the stimulated-emission denominator
    const double e296 = exp_k1(-c2 / kT0 * nu);
    const double neg_c2_e = -c2 * elower;
    const int n_layers = tab.n;
#pragma unroll 1
    for (int ly = 0; ly < n_layers; ++ly) {
        const K1Layer &K = tab.rows[ly];
        unsigned int flags = 0;
        if (in_range && !real) {                                  // padding record: never in a window
            if (K.narrow) {
                K.recA[l] = make_float4(-K2_SENTINEL, 0.f, 1.f, 0.f);
                reinterpret_cast<float2 *>(K.recB)[l] = make_float2(-1.f, -1.f);
            } else {
                K.recA[l] = make_float4(-K2_SENTINEL, -K2_SENTINEL, 0.f, 0.f);
                K.recB[l] = make_float4(1.f, 1.f, 0.f, -1.f);
            }
            K.recD[l] = -1.f;
        } else if (real) {
            const GroupParams p = K.gp[g];
            // FP64 divisions and transcendentals are what this kernel costs, so: per-layer reciprocals come from
            // the host (K.lc), each regime needs a single 1/h, the 5th root is a Newton iteration, and the
            // layer-independent exponential is hoisted.  This moves roundings by an ulp or two relative to the
            // reference's operation order (1e-16); the parity tests hold the per-line values to 1e-12.
            const double pp0 = K.lc.p_over_p0;
            const double dshift = delta * pp0;
            const double nus = nu + dshift;
            // (t0/T)^n = exp(n * log(t0/T)): the log is a per-layer constant
            const double gl = ((1 - p.conc) * gair + p.conc * gself) * pp0 * exp_k1(nair * K.lc.log_t0_over_t);
            const double gd = nus * p.dopp;
            // regime (pyradClasses.py:378-387): ratio = gl / gd compared with .01 and 100.  Away from the two
            // thresholds products decide; within 1e-14 of one (or for gd <= 0: inf / negative ratios, as numpy)
            // the division itself does, so the tag is exactly the reference's.
            int regime;
            {
                const double lo = .01 * gd, hi = 100 * gd;
                const double eps = 1e-14;
                if (gd > 0 && gl < lo * (1 - eps)) regime = REGIME_GAUSS;
                else if (gd > 0 && gl > hi * (1 + eps)) regime = REGIME_LORENTZ;
                else if (gd > 0 && gl > lo * (1 + eps) && gl < hi * (1 - eps)) regime = REGIME_VOIGT;
                else {
                    const double ratio = gl / gd;                 // gd == 0 -> inf -> Lorentz, as numpy
                    regime = ratio < .01 ? REGIME_GAUSS : (ratio > 100 ? REGIME_LORENTZ : REGIME_VOIGT);
                }
            }
            // stimulated emission (1 - e^{-c2 nu*/T}) / (1 - e^{-c2 nu*/t0}); e^{-c2 nu*/t0} = e296 * e^{-c2 dshift/t0}
            const double x0 = K.lc.neg_c2_over_t0 * dshift;
            // (next to 0 cm^-1 the denominator cancels: keep the reference's own single exponential there)
            const double e_t0 = (fabs(x0) <= 0.02 && nu > 1.0) ? e296 * exp_tiny(x0) : exp_k1(K.lc.neg_c2_over_t0 * nus);
            const double stim = (1 - exp_k1(K.lc.neg_c2_over_t * nus)) / (1 - e_t0);
            // exp(-c2 E/T) / exp(-c2 E/t0) evaluated as one exponential (same value to ~1e-16)
            const double boltz = exp_k1(neg_c2_e * K.lc.inv_t_minus_inv_t0);
            const double S = s296 * p.qratio * stim * boltz;
            const double sw = S * p.weight * (K.scale_dev ? __ldg(K.scale_dev) : K.scale);
            const double inv_res2 = K.lc.inv_res2;
            const double log2e = 1.4426950408889634;
            const double inv_sqrtpi = 0.5641895835477563;         // 1/sqrt(pi)
            double A, B, G, C, bg;                                // bg: Gaussian (h/res)^2
            float dg = -1.0f;
            if (regime == REGIME_GAUSS) {
                const double inv_gd = 1.0 / gd;
                A = 0.0; B = 1.0;
                G = sw * inv_gd * inv_sqrtpi;
                bg = gd * gd * inv_res2;
                C = -log2e * K.lc.res2 * inv_gd * inv_gd;
            } else if (regime == REGIME_LORENTZ) {
                A = sw * gl * (inv_res2 / kPi);
                B = gl * gl * inv_res2;
                G = 0.0; C = -1.0; bg = 0.0;
            } else {
                const double gFW = 2 * gd, lFW = 2 * gl;
                const double g2 = gFW * gFW, l2 = lFW * lFW;
                const double f5 = pow5(gFW) + 2.69269 * g2 * g2 * lFW + 2.42843 * g2 * gFW * l2 +
                                  4.47163 * g2 * l2 * lFW + .07842 * gFW * l2 * l2 + pow5(lFW);
                // f = f5 ** .2 (pyradLineshape.py:66): 1/f by Newton when f5 sits in the FP32 seed's range
                double inv_f;
                if (f5 > 1e-30 && f5 < 1e30) inv_f = inv_fifth_root(f5);
                else inv_f = 1.0 / exp(.2 * log(f5));
                const double r2 = inv_f * inv_f;
                const double hh = 0.5 * (f5 * (r2 * r2));         // f / 2, f = f5 * f5^(-4/5)
                const double inv_hh = 2 * inv_f;
                const double rho = gl * inv_hh;                   // lFW / f
                const double eta = 1.36603 * rho - .47719 * rho * rho + .11116 * rho * rho * rho;
                A = sw * eta * hh * (inv_res2 / kPi);
                B = hh * hh * inv_res2;
                G = sw * (1 - eta) * inv_hh * inv_sqrtpi;
                bg = B;
                C = -log2e * K.lc.res2 * inv_hh * inv_hh;
            }
            // near-zone radius: beyond it the Gaussian term is < 2e-7 of the same line's Lorentz term
            // (or below the scaled FP32 floor for Gaussian-only lines), so K2 may skip it.  FP32 is
            // plenty here (the radius is rounded up and padded): solve e^{-t2}(1+t2) <= rho9 (the 2e-7 ratio) with one
            // fixed-point step from t2 = -ln(rho9) plus a margin of 1 (the step undershoots by < 1).
            // (G can be negative: a line whose pressure-shifted wavenumber is < 0 gets a negative Doppler width
            // in the reference, pyradClasses.py:261-263 -- reproduced, so the test is on |G|.)
            if (G != 0.0) {
                float t2 = 160.f;
                if (A != 0.0) {
                    const float rho9 = (float)fabs(2e-7 * A / (B * G));
                    if (rho9 >= 1.f) t2 = -1.f;
                    else {
                        const float ln = -__logf(fmaxf(rho9, 1e-37f));
                        t2 = fminf(ln + __logf(1.f + ln) + 1.5f, 160.f);
                    }
                }
                if (t2 > 0.f) dg = fminf(ceilf(sqrtf(t2 * (float)bg)) + 2.f, 3.0e7f);
            }
            // FP32 range guards: the paired far path forms A*(d^2+B) with |d| <= wm.
            const double wq = K.wm + 4096.0;                      // partial lines reach one warp span past the window
            const double qmax = wq * wq + B;
            if (!(isfinite(A) && isfinite(G) && isfinite(B) && isfinite(C))) flags |= FLAG_NONFINITE;
            // the triple-reciprocal path forms |A| q^2 and q^3
            else if (fabs(A) * qmax * qmax > 8.0e37 || fabs(G) > 8.0e37 || qmax * qmax * qmax > 8.0e37) flags |= FLAG_OVERFLOW;
            const float Af = (float)A, Bf = (float)B;
            if (K.narrow) {                                       // compact layout of k2_point / k2_narrow
                K.recA[l] = make_float4(nf, Af, Bf, (float)G);
                reinterpret_cast<float2 *>(K.recB)[l] = make_float2((float)C, dg);
                K.recD[l] = (float)C;
            } else {
                K.recA[l] = make_float4(nf, nf, Af, Af);
                K.recB[l] = make_float4(Bf, Bf, (float)G, (float)C);
                K.recD[l] = dg;
            }
            if (dbg.nu_shift) dbg.nu_shift[l] = nus;
            if (dbg.gl) dbg.gl[l] = gl;
            if (dbg.gd) dbg.gd[l] = gd;
            if (dbg.st) dbg.st[l] = S;
            if (dbg.regime) dbg.regime[l] = regime;
        }
        // OR of the status flags, one atomic per warp that has something to report (order independent).
        flags = __reduce_or_sync(0xffffffffu, flags);
        if ((threadIdx.x & 31) == 0 && flags) atomicOr(&K.st->flags, flags);
    }
}

}  // namespace prb

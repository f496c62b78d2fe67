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

// max |x| of a column as its bit pattern (|x| >= 0: bit order == value order), and the checks of an uploaded line list --
// "ascending nu0" / "group id in range" -- as device flags (prb_upload_lines, prb_upload_line_groups, prb_gas_cell_host).
__global__ void __launch_bounds__(256)
k0_absmax(const double *__restrict__ x, int64_t n, unsigned long long *__restrict__ out_bits) {
    unsigned long long b = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = fabs(x[i]);
        const unsigned long long v = a == a ? (unsigned long long)__double_as_longlong(a) : 0ull;   // (NaN: skipped, as `>` does on the host)
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

// Group ids of one segment of a grouped line list (the gaps between segments keep the -1 of the memset before).
__global__ void __launch_bounds__(256)
k0_set_group(int32_t *__restrict__ group, int64_t n, int32_t g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) group[i] = g;
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
// Every 64-bit literal of the per-line arithmetic also travels in the parameter block: a double that is not a kernel
// parameter reaches a DFMA as an immediate pair (two moves into uniform registers per use -- a __constant__ array with an
// initialiser is folded into the same thing), and the layer loop re-materialises them every trip.  ptxas showed 210 such
// moves; from the parameter bank they are plain operands.
enum K1Const {
    K1C_EXP_INV = 0, K1C_EXP_MAGIC, K1C_EXP_HI, K1C_EXP_LO, K1C_EXP_5, K1C_EXP_4, K1C_EXP_3,     // exp_k1
    K1C_T7, K1C_T6, K1C_T5, K1C_T4, K1C_T3,                                                  // exp_tiny 1/5040 .. 1/6
    K1C_V1, K1C_V2, K1C_V3, K1C_V4,                                                          // f5: 2.69269 2.42843 4.47163 .07842
    K1C_E1, K1C_E2, K1C_E3,                                                                  // eta: 1.36603 .47719 .11116
    K1C_LOG2E, K1C_INV_SQRTPI, K1C_INV_PI, K1C_FIFTH, K1C_C2, K1C_NEG_C2_T0,
    K1C_LO_A, K1C_LO_B, K1C_HI_A, K1C_HI_B,                                                  // regime thresholds with their 1e-14 margins
    K1C_BIG, K1C_COUNT
};
struct K1Table {
    int n;
    int pad;
    double c[K1C_COUNT];
    K1Layer rows[K1_MAX_LAYERS];
};
inline void k1_fill_constants(K1Table &t) {
    const double eps = 1e-14, c2 = cLight * hPlanck * 100 / kBoltz;
    const double v[K1C_COUNT] = {
        92.33248261689366, 6755399441055744.0, -0.01083042469326756, -2.9815858269852933e-12, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,
        1.0 / 5040, 1.0 / 720, 1.0 / 120, 1.0 / 24, 1.0 / 6,
        2.69269, 2.42843, 4.47163, .07842,
        1.36603, .47719, .11116,
        1.4426950408889634, 0.5641895835477563, 1.0 / kPi, 0.2, c2, -c2 / kT0,
        .01 * (1 - eps), .01 * (1 + eps), 100 * (1 + eps), 100 * (1 - eps),
        8.0e37};
    for (int i = 0; i < K1C_COUNT; ++i) t.c[i] = v[i];
}

// exp(x) for the arguments this kernel meets (|x| <= 700): table driven.  x = (64 k + j) ln2/64 + r with |r| <= ln2/128,
//   exp(x) = 2^k * 2^(j/64) * exp(r),  exp(r) - 1 = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120)   (remainder r^6/720 < 4e-17)
// -- ten FP64 instructions instead of the eighteen of a degree-13 polynomial (the kernel is FP64-pipe bound and spends
// three of these per line and layer).  The 64 values 2^(j/64) are correctly rounded literals, copied into shared memory
// by the CTA (a constant-bank lookup with a per-thread index would serialise); ln2/64 is split so that n * hi is exact.
// ~1 ulp; the parity tests hold K1's outputs to 1e-12.
__constant__ double K1_EXP_TAB[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};
// (Arguments are clamped to [-708, 709] instead of branching to the library routine: below, the true value is a
// denormal that every later product flushes anyway; above, 1.6e308 overflows the range guards just as inf would.  NaN
// inputs never get here unnoticed: the kernel flags non-finite line data once per line.)
__device__ __forceinline__ double exp_k1(double x, const double *__restrict__ tab, const double *__restrict__ c) {
    // |x| >= 704 (one integer compare on the high word; NaN and inf land here too): clamp, out of the common path
    if ((__double2hiint(x) & 0x7fffffff) >= 0x40860000) x = fmin(fmax(x, -708.0), 709.0);
    const double t = fma(x, c[K1C_EXP_INV], c[K1C_EXP_MAGIC]);
    const int n = __double2loint(t);
    const double nf = t - c[K1C_EXP_MAGIC];
    double r = fma(nf, c[K1C_EXP_HI], x);
    r = fma(nf, c[K1C_EXP_LO], r);
    double p = fma(r, c[K1C_EXP_5], c[K1C_EXP_4]);
    p = fma(p, r, c[K1C_EXP_3]);
    p = fma(p, r, 0.5);
    const double em1 = fma(p * r, r, r);
    const double tj = tab[n & 63];
    const double v = fma(tj, em1, tj);
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}

// 1 / x to ~1 ulp without the IEEE division sequence: hardware seed (MUFU.RCP64H) and two Newton steps.
__device__ __forceinline__ double rcp_k1(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

// f5^(-1/5) without log/exp/division: FP32 seed (two MUFU) and two Newton steps on g(r) = r^-5 - f5,
//   r <- r (1 + (1 - f5 r^5)/5)   (quadratic: 1e-6 -> ~3e-12 -> below FP64 resolution).
__device__ __forceinline__ double inv_fifth_root(double f5, const double *__restrict__ c) {
    double r = (double)__powf((float)f5, -0.2f);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double r2 = r * r;
        const double r5 = r2 * r2 * r;
        r = fma(r, fma(-f5, r5, 1.0) * c[K1C_FIFTH], r);
    }
    return r;
}

// exp(x) for |x| <= 0.02 (degree-7 Taylor, truncation < 1e-18)
__device__ __forceinline__ double exp_tiny(double x, const double *__restrict__ c) {
    double p = fma(c[K1C_T7], x, c[K1C_T6]);
    p = fma(p, x, c[K1C_T5]);
    p = fma(p, x, c[K1C_T4]);
    p = fma(p, x, c[K1C_T3]);
    p = fma(p, x, 0.5);
    p = fma(p, x, 1.0);
    return fma(p, x, 1.0);
}

// One thread per line; the thread keeps the line's seven constants in registers and walks the layers of the
// batch, so the SoA columns are read once per launch instead of once per layer.
#ifndef PRB_K1_UNROLL
#define PRB_K1_UNROLL 1
#endif
constexpr int K1_UNROLL = PRB_K1_UNROLL;   // layers per trip of the layer loop
#ifndef PRB_K1_MINB
#define PRB_K1_MINB 4
#endif
// DEBUG: also write the FP64 per-line values of the parity tests (prb_debug_line_params); compiled out of the product
// launches (five pointer tests per line and layer otherwise).
template <bool DEBUG = false>
__global__ void __launch_bounds__(256, PRB_K1_MINB)
k1_prepass(LinesSoA L, const int32_t *__restrict__ idx, const __grid_constant__ K1Table tab,
           int64_t l_begin, int64_t l_end, int64_t n_lines, int64_t i_base, DebugOut dbg) {
    __shared__ double exp_tab[64];
    if (threadIdx.x < 64) exp_tab[threadIdx.x] = K1_EXP_TAB[threadIdx.x];
    __syncthreads();
    const double *__restrict__ cst = tab.c;                     // parameter-bank constants (K1Const)
    const double c2 = cst[K1C_C2];                              // cLight * hPlanck * 100 / kBoltz, pyradIntensity.py:13
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
    // bad line data (NaN / inf in any column) is reported once per line, whatever the arithmetic below makes of it
    const bool bad_line = real && !isfinite(nu + delta + gair + gself + nair + elower + s296);
    // exp(-c2 nu0 / t0): the layer-independent factor of the stimulated-emission denominator
    const double e296 = exp_k1(cst[K1C_NEG_C2_T0] * nu, exp_tab, cst);
    const double neg_c2_e = -c2 * elower;
    const int n_layers = tab.n;
#pragma unroll K1_UNROLL
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
            const double gl = ((1 - p.conc) * gair + p.conc * gself) * pp0 * exp_k1(nair * K.lc.log_t0_over_t, exp_tab, cst);
            const double gd = nus * p.dopp;
            // regime (pyradClasses.py:378-387): ratio = gl / gd compared with .01 and 100.  Away from the two
            // thresholds products decide; within 1e-14 of one (or for gd <= 0: inf / negative ratios, as numpy)
            // the division itself does, so the tag is exactly the reference's.
            int regime;
            {
                if (gd > 0 && gl < gd * cst[K1C_LO_A]) regime = REGIME_GAUSS;
                else if (gd > 0 && gl > gd * cst[K1C_HI_A]) regime = REGIME_LORENTZ;
                else if (gd > 0 && gl > gd * cst[K1C_LO_B] && gl < gd * cst[K1C_HI_B]) regime = REGIME_VOIGT;
                else {
                    const double ratio = gl / gd;                 // gd == 0 -> inf -> Lorentz, as numpy
                    regime = ratio < .01 ? REGIME_GAUSS : (ratio > 100 ? REGIME_LORENTZ : REGIME_VOIGT);
                }
            }
            // stimulated emission (1 - e^{-c2 nu*/T}) / (1 - e^{-c2 nu*/t0}); e^{-c2 nu*/t0} = e296 * e^{-c2 dshift/t0}
            const double x0 = K.lc.neg_c2_over_t0 * dshift;
            // (next to 0 cm^-1 the denominator cancels: keep the reference's own single exponential there)
            const double e_t0 = (fabs(x0) <= 0.02 && nu > 1.0) ? e296 * exp_tiny(x0, cst) : exp_k1(K.lc.neg_c2_over_t0 * nus, exp_tab, cst);
            // (the quotient by reciprocal + multiply: ~1 ulp, a third of the IEEE division's instructions; a denominator that
            // is 0, denormal or huge -- nu* at or below 0 cm^-1 -- takes the division itself)
            const double den = 1 - e_t0, num = 1 - exp_k1(K.lc.neg_c2_over_t * nus, exp_tab, cst);
            const double stim = (fabs(den) > 1e-300 && fabs(den) < 1e300) ? num * rcp_k1(den) : num / den;
            // exp(-c2 E/T) / exp(-c2 E/t0) evaluated as one exponential (same value to ~1e-16)
            const double boltz = exp_k1(neg_c2_e * K.lc.inv_t_minus_inv_t0, exp_tab, cst);
            const double S = s296 * p.qratio * stim * boltz;
            const double sw = S * p.weight * (K.scale_dev ? __ldg(K.scale_dev) : K.scale);
            const double inv_res2 = K.lc.inv_res2;
            const double log2e = cst[K1C_LOG2E];
            const double inv_sqrtpi = cst[K1C_INV_SQRTPI];        // 1/sqrt(pi)
            double A, B, G, C, bg;                                // bg: Gaussian (h/res)^2
            float dg = -1.0f;
            if (regime == REGIME_GAUSS) {
                const double inv_gd = 1.0 / gd;
                A = 0.0; B = 1.0;
                G = sw * inv_gd * inv_sqrtpi;
                bg = gd * gd * inv_res2;
                C = -log2e * K.lc.res2 * inv_gd * inv_gd;
            } else if (regime == REGIME_LORENTZ) {
                A = sw * gl * (inv_res2 * cst[K1C_INV_PI]);
                B = gl * gl * inv_res2;
                G = 0.0; C = -1.0; bg = 0.0;
            } else {
                const double gFW = 2 * gd, lFW = 2 * gl;
                // f5 = g^5 + 2.69269 g^4 l + 2.42843 g^3 l^2 + 4.47163 g^2 l^3 + .07842 g l^4 + l^5 (pyradLineshape.py:62-66),
                // Horner in g with the powers of l: 13 FP64 operations instead of 27 (same value to a few 1e-16)
                const double l2 = lFW * lFW, l3 = l2 * lFW, l4 = l2 * l2;
                double f5 = fma(cst[K1C_V1], lFW, gFW);
                f5 = fma(f5, gFW, cst[K1C_V2] * l2);
                f5 = fma(f5, gFW, cst[K1C_V3] * l3);
                f5 = fma(f5, gFW, cst[K1C_V4] * l4);
                f5 = fma(f5, gFW, l4 * lFW);
                // f = f5 ** .2 (pyradLineshape.py:66): 1/f by Newton when f5 sits in the FP32 seed's range
                double inv_f;
                if (f5 > 1e-30 && f5 < 1e30) inv_f = inv_fifth_root(f5, cst);
                else inv_f = 1.0 / exp(.2 * log(f5));
                const double r2 = inv_f * inv_f;
                const double hh = 0.5 * (f5 * (r2 * r2));         // f / 2, f = f5 * f5^(-4/5)
                const double inv_hh = 2 * inv_f;
                const double rho = gl * inv_hh;                   // lFW / f
                const double eta = rho * fma(rho, fma(cst[K1C_E3], rho, -cst[K1C_E2]), cst[K1C_E1]);
                A = sw * eta * hh * (inv_res2 * cst[K1C_INV_PI]);
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
                    const float rho9 = fabsf(2e-7f * (float)A / ((float)B * (float)G));   // (FP32: the scaled records' own range)
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
            if (bad_line || !(isfinite(A) && isfinite(G) && isfinite(B) && isfinite(C))) flags |= FLAG_NONFINITE;
            // the triple-reciprocal path forms |A| q^2 and q^3
            else if (fabs(A) * qmax * qmax > cst[K1C_BIG] || fabs(G) > cst[K1C_BIG] || qmax * qmax * qmax > cst[K1C_BIG]) flags |= FLAG_OVERFLOW;
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
            if (DEBUG) {
                if (dbg.nu_shift) dbg.nu_shift[l] = nus;
                if (dbg.gl) dbg.gl[l] = gl;
                if (dbg.gd) dbg.gd[l] = gd;
                if (dbg.st) dbg.st[l] = S;
                if (dbg.regime) dbg.regime[l] = regime;
            }
        }
        // OR of the status flags, one atomic per warp that has something to report (order independent).
        flags = __reduce_or_sync(0xffffffffu, flags);
        if ((threadIdx.x & 31) == 0 && flags) atomicOr(&K.st->flags, flags);
    }
}

// Per-layer line ranges (prb_set_layer_line_range): the reference loads a layer's lines from that layer's own effective
// range, strictly inside (pyradUtilities.py:437-438); with ONE list for a column, layer l leaves out the lines with
// nu0 <= lo or nu0 >= hi.  The list is ascending, so those are a prefix and a suffix -- found by binary search, a few
// hundred records per layer -- and K1 itself stays as it is (the test inside its layer loop cost 2.5 % of K1).  A record
// that is left out keeps its place in the list (K2 searches the records by position) with zero amplitudes and no Gaussian
// near zone: it adds exactly nothing.
struct LineRangeRow {
    float4 *recA, *recB;
    float *recD;
    double lo, hi;
    int narrow, pad;
};
struct LineRangeTable {
    int n, pad;
    LineRangeRow rows[K1_MAX_LAYERS];
};
__global__ void __launch_bounds__(256)
k1_mask_line_range(const double *__restrict__ nu0, int64_t l_begin, int64_t l_end, const __grid_constant__ LineRangeTable tab) {
    const LineRangeRow &R = tab.rows[blockIdx.x];
    int64_t a = l_begin, b = l_end;                               // first line with nu0 > lo
    while (a < b) { const int64_t m = (a + b) >> 1; if (nu0[m] > R.lo) b = m; else a = m + 1; }
    const int64_t keep0 = a;
    a = keep0; b = l_end;                                         // first line with nu0 >= hi
    while (a < b) { const int64_t m = (a + b) >> 1; if (nu0[m] >= R.hi) b = m; else a = m + 1; }
    const int64_t keep1 = a;
    const int64_t n_out = (keep0 - l_begin) + (l_end - keep1);
    for (int64_t t = threadIdx.x; t < n_out; t += blockDim.x) {
        const int64_t l = t < keep0 - l_begin ? l_begin + t : keep1 + (t - (keep0 - l_begin));
        const float pos = R.recA[l].x;
        if (R.narrow) {
            R.recA[l] = make_float4(pos, 0.f, 1.f, 0.f);
            reinterpret_cast<float2 *>(R.recB)[l] = make_float2(-1.f, -1.f);
        } else {
            R.recA[l] = make_float4(pos, pos, 0.f, 0.f);
            R.recB[l] = make_float4(1.f, 1.f, 0.f, -1.f);
        }
        R.recD[l] = -1.f;
    }
}

}  // namespace prb

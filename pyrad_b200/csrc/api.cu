// api.cu -- the C ABI of libpyrad_b200.so (see include/pyrad_b200.h).  Host-side orchestration only:
// buffers, launches, copies.  No physics is evaluated on the CPU and there is no CPU fallback.
#include "../../include/pyrad_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "k1_prepass.cuh"
#include "k2_line_sum.cuh"
#include "k2_narrow.cuh"
#include "k3_stream.cuh"

using namespace prb;

static thread_local std::string g_err = "";

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char b_[512];                                                                     \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                     __FILE__, __LINE__);                                                     \
            return fail(PRB_ERR_CUDA, b_);                                                    \
        }                                                                                     \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct PrepassArgs {
    double T = 0, P = 0, scale = 1;
    int64_t W = 0, wm = 0, l0 = 0, l1 = 0, k1_begin = 0, k1_end = 0;
    int narrow = 0;              // record layout / kernel: 1 = k2_narrow (thread-per-point gather)
    bool valid = false;
};

struct prb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};

    // line list
    int64_t n_lines = 0, n_alloc = 0;
    int n_groups = 0;
    DevBuf<double> nu0, s296, gair, gself, elower, nair, delta;
    DevBuf<int32_t> group;
    bool has_group = false;
    double s_max = 0;
    bool lines_set = false;

    // grid
    double range_min = 0, res = 0;
    int64_t n_total = 0, i_begin = 0, i_end = 0;
    bool grid_set = false;
    DevBuf<int32_t> idx;
    std::vector<int32_t> h_idx;

    // per-layer
    DevBuf<float4> recA, recB;
    DevBuf<float> recD;
    DevBuf<GroupParams> gp;       // n_layers * n_groups
    DevBuf<DevState> st;          // one per layer
    PrepassArgs last;
    int k2_variant = PRB_K2_CLASSED, k2_ppt = 0;
    int64_t narrow_wm = 100;     // windows with W-2 below this use k2_narrow

    // outputs / scratch
    DevBuf<double> out64;
    DevBuf<double> scratch_a, scratch_b, scratch_c, scratch_d, scratch_w;
    // atmosphere
    DevBuf<float> kmat, rad, trans;
    DevBuf<FoldLayer> fold;
    int64_t kmat_ld = 0;
    int atm_layers = 0;
    // optional stage timing of prb_atmosphere (CUDA events on the engine stream)
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    float t_k1 = 0, t_k2 = 0, t_k3 = 0;
    std::vector<float> t_layer_k1, t_layer_k2;
};

static int64_t chunk_len(const prb_engine *e) { return e->i_end - e->i_begin; }

extern "C" int prb_abi_version(void) { return PRB_ABI_VERSION; }
extern "C" const char *prb_last_error(void) { return g_err.c_str(); }

extern "C" int prb_create(int device, prb_engine **out) {
    if (!out) return fail(PRB_ERR_ARG, "prb_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(PRB_ERR_NODEVICE, std::string("prb_create: no CUDA device (") + cudaGetErrorString(ce) +
                                          "); libpyrad_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(PRB_ERR_ARG, "prb_create: device ordinal out of range");
    CK(cudaSetDevice(device));
    prb_engine *e = new prb_engine();
    e->device = device;
    if (cudaGetDeviceProperties(&e->prop, device) != cudaSuccess) {
        delete e;
        return fail(PRB_ERR_CUDA, "prb_create: cudaGetDeviceProperties failed");
    }
    if (e->prop.major != 10) {
        char b[256];
        snprintf(b, sizeof b, "prb_create: device is sm_%d%d; this library only carries sm_100a code (no fallback)",
                 e->prop.major, e->prop.minor);
        delete e;
        return fail(PRB_ERR_NODEVICE, b);
    }
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return fail(PRB_ERR_CUDA, "prb_create: cudaStreamCreate failed");
    }
    *out = e;
    return PRB_OK;
}

extern "C" int prb_destroy(prb_engine *e) {
    if (!e) return PRB_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    e->nu0.release(); e->s296.release(); e->gair.release(); e->gself.release();
    e->elower.release(); e->nair.release(); e->delta.release(); e->group.release();
    e->idx.release(); e->recA.release(); e->recB.release(); e->recD.release(); e->gp.release(); e->st.release();
    e->out64.release(); e->scratch_a.release(); e->scratch_b.release(); e->scratch_c.release();
    e->scratch_d.release(); e->scratch_w.release();
    e->kmat.release(); e->rad.release(); e->trans.release(); e->fold.release();
    for (auto x : e->ev) cudaEventDestroy(x);
    cudaStreamDestroy(e->stream);
    delete e;
    return PRB_OK;
}

extern "C" void *prb_stream(prb_engine *e) { return e ? (void *)e->stream : nullptr; }

static int check_flags(prb_engine *e, int n_states) {
    std::vector<DevState> h(n_states);
    CK(cudaMemcpyAsync(h.data(), e->st.p, sizeof(DevState) * n_states, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    unsigned int f = 0;
    for (auto &s : h) f |= s.flags;
    if (f & FLAG_NONFINITE)
        return fail(PRB_ERR_RANGE, "line prepass produced a non-finite coefficient (bad line data or T/P/Q inputs)");
    if (f & FLAG_OVERFLOW)
        return fail(PRB_ERR_RANGE, "line coefficient exceeds the scaled FP32 range (window too wide or S(T)/S296 too large)");
    return PRB_OK;
}

extern "C" int prb_synchronize(prb_engine *e) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return PRB_OK;
}

extern "C" int prb_device_info(prb_engine *e, int *sm_count, int *cc_major, int *cc_minor, int *sm_clock_khz,
                               size_t *free_bytes, size_t *total_bytes) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    if (sm_count) *sm_count = e->prop.multiProcessorCount;
    if (cc_major) *cc_major = e->prop.major;
    if (cc_minor) *cc_minor = e->prop.minor;
    if (sm_clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device);
        *sm_clock_khz = khz;
    }
    if (free_bytes || total_bytes) {
        size_t f = 0, t = 0;
        CK(cudaMemGetInfo(&f, &t));
        if (free_bytes) *free_bytes = f;
        if (total_bytes) *total_bytes = t;
    }
    return PRB_OK;
}

extern "C" int prb_set_k2_variant(prb_engine *e, int variant, int ppt) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (variant != PRB_K2_GENERAL && variant != PRB_K2_CLASSED) return fail(PRB_ERR_ARG, "unknown K2 variant");
    if (ppt != 0 && ppt != 2 && ppt != 4 && ppt != 8 && ppt != 16)
        return fail(PRB_ERR_ARG, "points_per_thread must be 0 (auto), 2, 4, 8 or 16");
    e->k2_variant = variant;
    e->k2_ppt = ppt;
    return PRB_OK;
}

// ------------------------------------------------------------------------------------ lines
extern "C" int prb_upload_lines(prb_engine *e, int64_t n, const double *nu0, const double *s296,
                                const double *gamma_air, const double *gamma_self, const double *elower,
                                const double *n_air, const double *delta_air, const int32_t *group,
                                int32_t n_groups) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (n < 0 || n > 2000000000LL) return fail(PRB_ERR_ARG, "prb_upload_lines: n out of range");
    if (n_groups < 1) return fail(PRB_ERR_ARG, "prb_upload_lines: n_groups must be >= 1");
    if (n > 0 && (!nu0 || !s296 || !gamma_air || !gamma_self || !elower || !n_air || !delta_air))
        return fail(PRB_ERR_ARG, "prb_upload_lines: NULL column");
    CK(cudaSetDevice(e->device));
    const int64_t na = n + 16;                                  // padding records: TMA copies are 16-byte granular
    DevBuf<double> *cols[7] = {&e->nu0, &e->s296, &e->gair, &e->gself, &e->elower, &e->nair, &e->delta};
    const double *src[7] = {nu0, s296, gamma_air, gamma_self, elower, n_air, delta_air};
    for (int c = 0; c < 7; ++c) {
        CK(cols[c]->ensure(na));
        if (n) CK(cudaMemcpyAsync(cols[c]->p, src[c], sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    }
    e->has_group = group != nullptr;
    if (group) {
        CK(e->group.ensure(na));
        if (n) CK(cudaMemcpyAsync(e->group.p, group, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    }
    // (the H2D copies above are in flight from pinned memory while the host validates)
    // one validation pass over the host columns: ascending nu0, max |S296|, group ids in range
    double smax = 0;
    bool sorted = true, group_ok = true;
    for (int64_t i = 0; i < n; ++i) {
        const double sa = std::fabs(s296[i]);
        smax = sa > smax ? sa : smax;
        if (i && !(nu0[i] >= nu0[i - 1])) sorted = false;
        if (group && (group[i] < 0 || group[i] >= n_groups)) group_ok = false;
    }
    CK(e->idx.ensure(na));
    CK(e->recA.ensure(na));
    CK(e->recB.ensure(na));
    CK(e->recD.ensure(na));
    CK(cudaStreamSynchronize(e->stream));
    e->lines_set = false;
    if (!sorted) return fail(PRB_ERR_ARG, "prb_upload_lines: nu0 must be ascending");
    if (!group_ok) return fail(PRB_ERR_ARG, "prb_upload_lines: group id out of range");
    e->n_lines = n;
    e->n_alloc = na;
    e->n_groups = n_groups;
    e->s_max = smax;
    e->lines_set = true;
    e->grid_set = false;
    e->last.valid = false;
    return PRB_OK;
}

// ------------------------------------------------------------------------------------ grid
extern "C" int prb_set_grid(prb_engine *e, double range_min, double res, int64_t n_total, int64_t i_begin,
                            int64_t i_end) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (!e->lines_set) return fail(PRB_ERR_STATE, "prb_set_grid: upload lines first");
    if (!(res > 0) || n_total < 0 || i_begin < 0 || i_end < i_begin || i_end > n_total)
        return fail(PRB_ERR_ARG, "prb_set_grid: bad grid");
    if (n_total > 2000000000LL) return fail(PRB_ERR_ARG, "prb_set_grid: n_total too large");
    CK(cudaSetDevice(e->device));
    const int64_t na = e->n_alloc;
    k0_line_index<<<(unsigned)((na + 255) / 256), 256, 0, e->stream>>>(e->nu0.p, e->n_lines, na, range_min, res,
                                                                      e->idx.p);
    CK(cudaGetLastError());
    e->h_idx.resize(na);
    CK(cudaMemcpyAsync(e->h_idx.data(), e->idx.p, sizeof(int32_t) * na, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->range_min = range_min;
    e->res = res;
    e->n_total = n_total;
    e->i_begin = i_begin;
    e->i_end = i_end;
    e->grid_set = true;
    e->last.valid = false;
    return PRB_OK;
}

// ------------------------------------------------------------------------------------ K1
static int build_group_params(int n_groups, double T, const double *conc, const double *molmass, const double *q_t,
                              const double *q_296, const double *weight, GroupParams *out, double *w_max) {
    double wm = 0;
    for (int g = 0; g < n_groups; ++g) {
        const double m = molmass[g] / 1000 / kAvogadro;                        // pyradClasses.py:294-296
        out[g].conc = conc[g];
        out[g].dopp = std::sqrt(2 * kBoltz * T / m / (cLight * cLight));       // pyradClasses.py:263
        out[g].qratio = q_296[g] / q_t[g];                                     // pyradIntensity.py:31
        out[g].weight = weight ? weight[g] : 1.0;
        wm = std::max(wm, std::fabs(out[g].weight));
    }
    *w_max = wm;
    return PRB_OK;
}

static double pick_scale(double s_max, double w_max) {
    // Power of two that puts the strongest possible S296*weight near 2^10: typical A = S*eta*h/pi/res^2
    // then sits around 2^25..2^35, leaving headroom for A*q^2 (q = d^2+B up to ~2^30) and ~2^-126 as the floor.
    const double m = s_max * w_max;
    if (!(m > 0) || !std::isfinite(m)) return 1.0;
    return std::ldexp(1.0, 10 - std::ilogb(m));
}

static LayerConsts layer_consts(double T, double P, double res) {
    LayerConsts lc;
    lc.log_t0_over_t = std::log(kT0 / T);
    lc.inv_t_minus_inv_t0 = 1.0 / T - 1.0 / kT0;
    lc.inv_res2 = 1.0 / (res * res);
    lc.res2 = res * res;
    lc.p_over_p0 = P / kP0;
    const double c2 = cLight * hPlanck * 100 / kBoltz;
    lc.neg_c2_over_t = -c2 / T;
    lc.neg_c2_over_t0 = -c2 / kT0;
    return lc;
}

static int launch_prepass(prb_engine *e, double T, double P, int64_t W, const GroupParams *gp_dev, double scale,
                          DevState *st_dev, DebugOut dbg, PrepassArgs *pa) {
    const int64_t wm = std::max<int64_t>(W - 2, 0);
    const int64_t n = e->n_lines;
    // lines that can reach the owned chunk: idx in [i_begin - wm, i_end - 1 + wm]
    const int32_t *hb = e->h_idx.data();
    const int64_t klo = e->i_begin - wm, khi = e->i_end - 1 + wm;
    int64_t l0 = std::lower_bound(hb, hb + n, klo, [](int32_t a, int64_t k) { return (int64_t)a < k; }) - hb;
    int64_t l1 = std::upper_bound(hb, hb + n, khi, [](int64_t k, int32_t a) { return k < (int64_t)a; }) - hb;
    if (l1 < l0) l1 = l0;
    const int narrow = (e->k2_variant == PRB_K2_CLASSED && wm < e->narrow_wm) ? 1 : 0;
    const int64_t kb = l0 & ~int64_t(3);
    const int64_t ke = std::min<int64_t>(l1 + 8, e->n_alloc);
    LinesSoA L{e->nu0.p, e->s296.p, e->gair.p, e->gself.p, e->elower.p, e->nair.p, e->delta.p,
               e->has_group ? e->group.p : nullptr};
    const int64_t cnt = ke - kb;
    if (cnt > 0) {
        k1_prepass<<<(unsigned)((cnt + 255) / 256), 256, 0, e->stream>>>(
            L, e->idx.p, gp_dev, kb, ke, n, T, P, layer_consts(T, P, e->res), scale, e->i_begin, (double)wm, narrow, e->recA.p, e->recB.p,
            e->recD.p, st_dev, dbg);
        CK(cudaGetLastError());
    }
    pa->T = T; pa->P = P; pa->scale = scale; pa->W = W; pa->wm = wm;
    pa->l0 = l0; pa->l1 = l1; pa->k1_begin = kb; pa->k1_end = ke;
    pa->narrow = narrow;
    pa->valid = true;
    return PRB_OK;
}

static int check_segment(prb_engine *e, int64_t wm) {
    // FP32 offsets must be exact integers: chunk + tile rounding + both windows below 2^24.
    if (chunk_len(e) + 2 * wm + 8192 >= (int64_t(1) << 24))
        return fail(PRB_ERR_RANGE, "owned grid chunk plus cutoff windows exceeds 2^24 points; shard the grid "
                                   "(prb_set_grid i_begin/i_end) into smaller chunks");
    return PRB_OK;
}

extern "C" int prb_layer_prepass(prb_engine *e, double T, double P, int32_t n_groups, const double *conc,
                                 const double *molmass, const double *q_t, const double *q_296,
                                 const double *weight, int64_t window_len) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (!e->grid_set) return fail(PRB_ERR_STATE, "prb_layer_prepass: set the grid first");
    if (n_groups != e->n_groups) return fail(PRB_ERR_ARG, "prb_layer_prepass: n_groups differs from the uploaded lines");
    if (!conc || !molmass || !q_t || !q_296) return fail(PRB_ERR_ARG, "prb_layer_prepass: NULL group array");
    if (window_len < 1) return fail(PRB_ERR_ARG, "prb_layer_prepass: window_len must be >= 1 (cutoff >= one sample)");
    if (!(T > 0) || !(P >= 0)) return fail(PRB_ERR_ARG, "prb_layer_prepass: bad T or P");
    CK(cudaSetDevice(e->device));
    int rc = check_segment(e, std::max<int64_t>(window_len - 2, 0));
    if (rc) return rc;
    std::vector<GroupParams> h(n_groups);
    double w_max = 0;
    build_group_params(n_groups, T, conc, molmass, q_t, q_296, weight, h.data(), &w_max);
    CK(e->gp.ensure(n_groups));
    CK(e->st.ensure(1));
    CK(cudaMemcpyAsync(e->gp.p, h.data(), sizeof(GroupParams) * n_groups, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(e->st.p, 0, sizeof(DevState), e->stream));
    const double scale = pick_scale(e->s_max, w_max);
    rc = launch_prepass(e, T, P, window_len, e->gp.p, scale, e->st.p, DebugOut{}, &e->last);
    if (rc) return rc;
    CK(cudaStreamSynchronize(e->stream));                       // h (pageable) must outlive the copy
    return PRB_OK;
}

extern "C" int prb_debug_line_params(prb_engine *e, double *nu_shift, double *gamma_l, double *gamma_d, double *s_t,
                                     int32_t *regime, int64_t *index) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (!e->last.valid) return fail(PRB_ERR_STATE, "prb_debug_line_params: run prb_layer_prepass first");
    CK(cudaSetDevice(e->device));
    const int64_t n = e->n_lines, na = e->n_alloc;
    CK(e->scratch_a.ensure(na)); CK(e->scratch_b.ensure(na)); CK(e->scratch_c.ensure(na)); CK(e->scratch_d.ensure(na));
    DevBuf<int32_t> reg;
    CK(reg.ensure(na));
    CK(cudaMemsetAsync(e->scratch_a.p, 0, sizeof(double) * na, e->stream));
    CK(cudaMemsetAsync(e->scratch_b.p, 0, sizeof(double) * na, e->stream));
    CK(cudaMemsetAsync(e->scratch_c.p, 0, sizeof(double) * na, e->stream));
    CK(cudaMemsetAsync(e->scratch_d.p, 0, sizeof(double) * na, e->stream));
    CK(cudaMemsetAsync(reg.p, 0xff, sizeof(int32_t) * na, e->stream));
    DebugOut dbg{e->scratch_a.p, e->scratch_b.p, e->scratch_c.p, e->scratch_d.p, reg.p};
    PrepassArgs pa;
    DevBuf<DevState> st;
    CK(st.ensure(1));
    CK(cudaMemsetAsync(st.p, 0, sizeof(DevState), e->stream));
    // whole list, so every line gets a value irrespective of the owned chunk
    const int64_t save_b = e->i_begin, save_e = e->i_end;
    e->i_begin = 0; e->i_end = e->n_total;
    const int64_t huge_w = std::max<int64_t>(e->last.W, 2) ;
    (void)huge_w;
    LinesSoA L{e->nu0.p, e->s296.p, e->gair.p, e->gself.p, e->elower.p, e->nair.p, e->delta.p,
               e->has_group ? e->group.p : nullptr};
    // records are scratch here: use temporaries so the live prepass is not disturbed
    DevBuf<float4> r4, r5; DevBuf<float> r2;
    CK(r4.ensure(na)); CK(r5.ensure(na)); CK(r2.ensure(na));
    k1_prepass<<<(unsigned)((na + 255) / 256), 256, 0, e->stream>>>(L, e->idx.p, e->gp.p, 0, na, n, e->last.T, e->last.P,
                                                                   layer_consts(e->last.T, e->last.P, e->res), e->last.scale, save_b, (double)e->last.wm, 0,
                                                                   r4.p, r5.p, r2.p, st.p, dbg);
    e->i_begin = save_b; e->i_end = save_e;
    CK(cudaGetLastError());
    if (nu_shift) CK(cudaMemcpyAsync(nu_shift, e->scratch_a.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    if (gamma_l) CK(cudaMemcpyAsync(gamma_l, e->scratch_b.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    if (gamma_d) CK(cudaMemcpyAsync(gamma_d, e->scratch_c.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    if (s_t) CK(cudaMemcpyAsync(s_t, e->scratch_d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    if (regime) CK(cudaMemcpyAsync(regime, reg.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (index) for (int64_t i = 0; i < n; ++i) index[i] = e->h_idx[i];
    r4.release(); r5.release(); r2.release(); reg.release(); st.release();
    return PRB_OK;
}

// ------------------------------------------------------------------------------------ K2
static int pick_ppt(const prb_engine *e, int64_t wm) {
    if (e->k2_ppt) return e->k2_ppt;
    // a warp spans 32*P points: keep the span well inside the window so most lines cover it fully
    if (wm >= 1024) return 8;
    if (wm >= 256) return 4;
    return 2;
}

template <int P>
static cudaError_t launch_k2_t(prb_engine *e, K2Args a) {
    const int tile = K2_CONSUMERS * 32 * P;
    a.n_tiles = (int)((a.n_chunk + tile - 1) / tile);
    if (a.n_tiles == 0) return cudaSuccess;
    const size_t smem = sizeof(K2Smem);
    static bool attr_set = false;                               // per template instance
    if (!attr_set) {
        cudaError_t ce = cudaFuncSetAttribute(k2_line_sum<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
        attr_set = true;
    }
    const int grid = std::min(a.n_tiles, K2_MIN_CTAS * e->prop.multiProcessorCount);
    k2_line_sum<P><<<grid, K2_THREADS, smem, e->stream>>>(a);
    return cudaGetLastError();
}

static int launch_line_sum(prb_engine *e, const PrepassArgs &pa, DevState *st_dev, void *out_dev, int out_mode) {
    K2Args a{};
    a.recA = e->recA.p;
    a.recB = e->recB.p;
    a.recD = e->recD.p;
    a.idx = e->idx.p;
    a.l_begin = (int)pa.l0;
    a.l_end = (int)pa.l1;
    a.i_begin = e->i_begin;
    a.n_chunk = (int)chunk_len(e);
    a.wm = (int)pa.wm;
    a.variant = e->k2_variant;
    a.out_mode = out_mode;
    a.inv_scale = 1.0 / pa.scale;
    a.out = out_dev;
    a.st = st_dev;
    cudaError_t ce;
    if (pa.narrow) {
        a.n_tiles = (a.n_chunk + KN_TILE - 1) / KN_TILE;
        if (a.n_tiles > 0) k2_narrow<<<a.n_tiles, KN_THREADS, sizeof(KNSmem), e->stream>>>(a);
        ce = cudaGetLastError();
        if (ce != cudaSuccess) return fail(PRB_ERR_CUDA, std::string("k2_narrow launch failed: ") + cudaGetErrorString(ce));
        return PRB_OK;
    }
    switch (pick_ppt(e, pa.wm)) {
        case 2: ce = launch_k2_t<2>(e, a); break;
        case 4: ce = launch_k2_t<4>(e, a); break;
        case 16: ce = launch_k2_t<16>(e, a); break;
        default: ce = launch_k2_t<8>(e, a); break;
    }
    if (ce != cudaSuccess) return fail(PRB_ERR_CUDA, std::string("k2_line_sum launch failed: ") + cudaGetErrorString(ce));
    return PRB_OK;
}

extern "C" int prb_line_sum_dev(prb_engine *e, void *out_dev, int out_mode) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (!e->last.valid) return fail(PRB_ERR_STATE, "prb_line_sum: run prb_layer_prepass first");
    if (!out_dev) return fail(PRB_ERR_ARG, "prb_line_sum_dev: NULL output");
    if (out_mode != PRB_OUT_F64 && out_mode != PRB_OUT_F32) return fail(PRB_ERR_ARG, "prb_line_sum_dev: bad out_mode");
    CK(cudaSetDevice(e->device));
    // the tile scheduler counter is consumed by a launch: reset it (the status flags must survive)
    CK(cudaMemsetAsync(&e->st.p->tile_counter, 0, sizeof(unsigned int), e->stream));
    return launch_line_sum(e, e->last, e->st.p, out_dev, out_mode);
}

extern "C" int prb_line_sum(prb_engine *e, double *out_host) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (!out_host && chunk_len(e) > 0) return fail(PRB_ERR_ARG, "prb_line_sum: NULL output");
    const int64_t nc = e->grid_set ? chunk_len(e) : 0;
    if (!e->last.valid) return fail(PRB_ERR_STATE, "prb_line_sum: run prb_layer_prepass first");
    CK(cudaSetDevice(e->device));
    CK(e->out64.ensure(nc));
    int rc = prb_line_sum_dev(e, e->out64.p, PRB_OUT_F64);
    if (rc) return rc;
    if (nc) CK(cudaMemcpyAsync(out_host, e->out64.p, sizeof(double) * nc, cudaMemcpyDeviceToHost, e->stream));
    rc = check_flags(e, 1);                                     // synchronises
    return rc;
}

extern "C" int64_t prb_pair_count(prb_engine *e) {
    if (!e || !e->last.valid) {
        fail(PRB_ERR_STATE, "prb_pair_count: run prb_layer_prepass first");
        return -1;
    }
    const int64_t wm = e->last.wm, b = e->i_begin, en = e->i_end - 1;
    int64_t total = 0;
    for (int64_t l = e->last.l0; l < e->last.l1; ++l) {
        const int64_t c = e->h_idx[l];
        const int64_t lo = std::max(c - wm, b), hi = std::min(c + wm, en);
        if (hi >= lo) total += hi - lo + 1;
    }
    return total;
}

// ------------------------------------------------------------------------------------ K3 (host buffers)
static unsigned stream_grid(const prb_engine *e, int64_t n, int per_thread = 1) {
    const int64_t blocks = (n / per_thread + 255) / 256;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)e->prop.multiProcessorCount * 16));
}

extern "C" int prb_layer_stream(prb_engine *e, int64_t n, int32_t n_mol, const double *sigma, const double *weight,
                                double depth_cm, double t_layer, double x0, double dx, double x_last,
                                const double *radiance_in, double *abs_coef, double *transmittance,
                                double *radiance_out) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (n < 0 || n_mol < 0) return fail(PRB_ERR_ARG, "prb_layer_stream: negative size");
    if (n_mol > 0 && (!sigma || !weight)) return fail(PRB_ERR_ARG, "prb_layer_stream: NULL sigma/weight");
    if (radiance_out && !radiance_in) return fail(PRB_ERR_ARG, "prb_layer_stream: radiance_out needs radiance_in");
    if (n == 0) return PRB_OK;
    CK(cudaSetDevice(e->device));
    CK(e->scratch_a.ensure((size_t)n * std::max(n_mol, 1)));
    CK(e->scratch_w.ensure(std::max(n_mol, 1)));
    CK(e->scratch_b.ensure(n)); CK(e->scratch_c.ensure(n)); CK(e->scratch_d.ensure(n));
    DevBuf<double> rin;
    if (n_mol) {
        CK(cudaMemcpyAsync(e->scratch_a.p, sigma, sizeof(double) * n * n_mol, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(e->scratch_w.p, weight, sizeof(double) * n_mol, cudaMemcpyHostToDevice, e->stream));
    }
    if (radiance_out) {
        CK(rin.ensure(n));
        CK(cudaMemcpyAsync(rin.p, radiance_in, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    }
    k3_layer_stream_f64<<<stream_grid(e, n), 256, 0, e->stream>>>(
        n, n_mol, e->scratch_a.p, e->scratch_w.p, depth_cm, t_layer, x0, dx, x_last, radiance_out ? rin.p : nullptr,
        abs_coef ? e->scratch_b.p : nullptr, transmittance ? e->scratch_c.p : nullptr,
        radiance_out ? e->scratch_d.p : nullptr);
    CK(cudaGetLastError());
    if (abs_coef) CK(cudaMemcpyAsync(abs_coef, e->scratch_b.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    if (transmittance) CK(cudaMemcpyAsync(transmittance, e->scratch_c.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    if (radiance_out) CK(cudaMemcpyAsync(radiance_out, e->scratch_d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    rin.release();
    return PRB_OK;
}

extern "C" int prb_planck(prb_engine *e, int64_t n, double x0, double dx, double x_last, double temp, double *out) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (n < 0 || (n > 0 && !out)) return fail(PRB_ERR_ARG, "prb_planck: bad arguments");
    if (n == 0) return PRB_OK;
    CK(cudaSetDevice(e->device));
    CK(e->scratch_b.ensure(n));
    k3_planck_f64<<<stream_grid(e, n), 256, 0, e->stream>>>(n, x0, dx, x_last, temp, e->scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, e->scratch_b.p, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return PRB_OK;
}

extern "C" int prb_xsc_place(prb_engine *e, int64_t n_out, int64_t dst0, int64_t src0, int64_t count, int interp,
                             double ax0, double adelta, int64_t n_file, const double *file_x, const double *file_y,
                             double *out) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (n_out < 0 || count < 0 || n_file < 1 || !file_y || (interp && !file_x) || (n_out > 0 && !out))
        return fail(PRB_ERR_ARG, "prb_xsc_place: bad arguments");
    if (n_out == 0) return PRB_OK;
    CK(cudaSetDevice(e->device));
    CK(e->scratch_a.ensure(n_file)); CK(e->scratch_c.ensure(n_file)); CK(e->scratch_b.ensure(n_out));
    if (interp) CK(cudaMemcpyAsync(e->scratch_a.p, file_x, sizeof(double) * n_file, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->scratch_c.p, file_y, sizeof(double) * n_file, cudaMemcpyHostToDevice, e->stream));
    k3_xsc_place<<<stream_grid(e, n_out), 256, 0, e->stream>>>(n_out, dst0, src0, count, interp, ax0, adelta, n_file,
                                                              e->scratch_a.p, e->scratch_c.p, e->scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, e->scratch_b.p, sizeof(double) * n_out, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return PRB_OK;
}

// ------------------------------------------------------------------------------------ atmosphere
extern "C" int prb_atmosphere(prb_engine *e, int32_t n_layers, int32_t n_groups, const double *depth_cm,
                              const double *t_layer, const double *p_layer, const double *conc, const double *molmass,
                              const double *q_t, const double *q_296, const int64_t *window_len, double t_surface,
                              double range_max) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    if (!e->grid_set) return fail(PRB_ERR_STATE, "prb_atmosphere: set the grid first");
    if (n_layers < 1 || n_groups != e->n_groups) return fail(PRB_ERR_ARG, "prb_atmosphere: bad n_layers / n_groups");
    if (!depth_cm || !t_layer || !p_layer || !conc || !molmass || !q_t || !q_296 || !window_len)
        return fail(PRB_ERR_ARG, "prb_atmosphere: NULL array");
    CK(cudaSetDevice(e->device));
    const int64_t nc = chunk_len(e);
    int64_t wmax = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (window_len[l] < 1) return fail(PRB_ERR_ARG, "prb_atmosphere: window_len must be >= 1");
        wmax = std::max<int64_t>(wmax, window_len[l] - 2);
    }
    int rc = check_segment(e, wmax);
    if (rc) return rc;

    // per-(layer, group) params: weight = conc * P / 1e4 / kB / T  (absCoef, pyradClasses.py:581-583)
    std::vector<GroupParams> h((size_t)n_layers * n_groups);
    std::vector<double> scale(n_layers);
    std::vector<FoldLayer> hf(n_layers);
    const double c2 = 100 * hPlanck * cLight / kBoltz;
    for (int l = 0; l < n_layers; ++l) {
        std::vector<double> w(n_groups);
        for (int g = 0; g < n_groups; ++g)
            w[g] = conc[(size_t)l * n_groups + g] * p_layer[l] / 1E4 / kBoltz / t_layer[l];
        double w_max = 0;
        build_group_params(n_groups, t_layer[l], conc + (size_t)l * n_groups, molmass, q_t + (size_t)l * n_groups,
                           q_296, w.data(), h.data() + (size_t)l * n_groups, &w_max);
        scale[l] = pick_scale(e->s_max, w_max);
        hf[l].neg_depth_log2e = (float)(-depth_cm[l] * 1.4426950408889634);
        hf[l].c2_over_t = (float)(c2 / t_layer[l]);
    }
    e->kmat_ld = (nc + 3) & ~int64_t(3);
    CK(e->gp.ensure(h.size()));
    CK(e->st.ensure(n_layers));
    CK(e->fold.ensure(n_layers));
    CK(e->kmat.ensure((size_t)e->kmat_ld * n_layers));
    CK(e->rad.ensure(e->kmat_ld));
    CK(e->trans.ensure(e->kmat_ld));
    CK(cudaMemcpyAsync(e->gp.p, h.data(), sizeof(GroupParams) * h.size(), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->fold.p, hf.data(), sizeof(FoldLayer) * n_layers, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(e->st.p, 0, sizeof(DevState) * n_layers, e->stream));
    if (e->atm_layers != n_layers) {
        CK(cudaMemsetAsync(e->kmat.p, 0, sizeof(float) * e->kmat_ld * n_layers, e->stream));
        e->atm_layers = n_layers;
    }
    const size_t n_ev = (size_t)2 * n_layers + 2;
    if (e->timing) {
        while (e->ev.size() < n_ev) {
            cudaEvent_t x;
            CK(cudaEventCreate(&x));
            e->ev.push_back(x);
        }
    }
    for (int l = 0; l < n_layers; ++l) {
        PrepassArgs pa;
        if (e->timing) CK(cudaEventRecord(e->ev[2 * l], e->stream));
        rc = launch_prepass(e, t_layer[l], p_layer[l], window_len[l], e->gp.p + (size_t)l * n_groups, scale[l],
                            e->st.p + l, DebugOut{}, &pa);
        if (rc) return rc;
        if (e->timing) CK(cudaEventRecord(e->ev[2 * l + 1], e->stream));
        rc = launch_line_sum(e, pa, e->st.p + l, e->kmat.p + (size_t)l * e->kmat_ld, PRB_OUT_F32);
        if (rc) return rc;
    }
    if (e->timing) CK(cudaEventRecord(e->ev[2 * n_layers], e->stream));
    if (nc > 0) {
        const double dx = e->n_total > 1 ? (range_max - e->range_min) / (double)(e->n_total - 1) : 0.0;
        k3_fold_f32<<<stream_grid(e, nc, 4), 256, 0, e->stream>>>(e->kmat.p, e->kmat_ld, n_layers, e->fold.p, nc,
                                                                 e->i_begin, e->n_total, e->range_min, dx, range_max,
                                                                 (float)(c2 / t_surface), e->rad.p, e->trans.p);
        CK(cudaGetLastError());
    }
    if (e->timing) CK(cudaEventRecord(e->ev[2 * n_layers + 1], e->stream));
    CK(cudaStreamSynchronize(e->stream));                       // pageable staging vectors go out of scope
    if (e->timing) {
        e->t_k1 = e->t_k2 = e->t_k3 = 0;
        e->t_layer_k1.assign(n_layers, 0.f);
        e->t_layer_k2.assign(n_layers, 0.f);
        for (int l = 0; l < n_layers; ++l) {
            float a = 0, b = 0;
            CK(cudaEventElapsedTime(&a, e->ev[2 * l], e->ev[2 * l + 1]));
            CK(cudaEventElapsedTime(&b, e->ev[2 * l + 1], e->ev[2 * l + 2]));
            e->t_k1 += a;
            e->t_k2 += b;
            e->t_layer_k1[l] = a;
            e->t_layer_k2[l] = b;
        }
        CK(cudaEventElapsedTime(&e->t_k3, e->ev[2 * n_layers], e->ev[2 * n_layers + 1]));
    }
    e->last.valid = false;
    return check_flags(e, n_layers);
}

extern "C" int prb_atmosphere_result_dev(prb_engine *e, void **radiance_dev, void **transmittance_dev) {
    if (!e || !e->atm_layers) return fail(PRB_ERR_STATE, "prb_atmosphere_result_dev: run prb_atmosphere first");
    if (radiance_dev) *radiance_dev = e->rad.p;
    if (transmittance_dev) *transmittance_dev = e->trans.p;
    return PRB_OK;
}

extern "C" int prb_atmosphere_kmatrix_dev(prb_engine *e, void **kmat_dev, int64_t *ld) {
    if (!e || !e->atm_layers) return fail(PRB_ERR_STATE, "prb_atmosphere_kmatrix_dev: run prb_atmosphere first");
    if (kmat_dev) *kmat_dev = e->kmat.p;
    if (ld) *ld = e->kmat_ld;
    return PRB_OK;
}

extern "C" int prb_atmosphere_read(prb_engine *e, double *radiance_host, double *transmittance_host) {
    if (!e || !e->atm_layers) return fail(PRB_ERR_STATE, "prb_atmosphere_read: run prb_atmosphere first");
    CK(cudaSetDevice(e->device));
    const int64_t nc = chunk_len(e);
    std::vector<float> tmp(nc);
    const float *src[2] = {e->rad.p, e->trans.p};
    double *dst[2] = {radiance_host, transmittance_host};
    for (int k = 0; k < 2; ++k) {
        if (!dst[k]) continue;
        CK(cudaMemcpyAsync(tmp.data(), src[k], sizeof(float) * nc, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        for (int64_t i = 0; i < nc; ++i) dst[k][i] = (double)tmp[i];
    }
    return PRB_OK;
}

extern "C" int prb_atmosphere_read_f32(prb_engine *e, float *radiance_host, float *transmittance_host) {
    if (!e || !e->atm_layers) return fail(PRB_ERR_STATE, "prb_atmosphere_read_f32: run prb_atmosphere first");
    CK(cudaSetDevice(e->device));
    const int64_t nc = chunk_len(e);
    if (radiance_host) CK(cudaMemcpyAsync(radiance_host, e->rad.p, sizeof(float) * nc, cudaMemcpyDeviceToHost, e->stream));
    if (transmittance_host)
        CK(cudaMemcpyAsync(transmittance_host, e->trans.p, sizeof(float) * nc, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return PRB_OK;
}

extern "C" int prb_set_timing(prb_engine *e, int enabled) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    e->timing = enabled != 0;
    return PRB_OK;
}

extern "C" int prb_atmosphere_timing(prb_engine *e, float *k1_ms, float *k2_ms, float *k3_ms) {
    if (!e || !e->atm_layers || !e->timing) return fail(PRB_ERR_STATE, "prb_atmosphere_timing: enable timing and run prb_atmosphere first");
    if (k1_ms) *k1_ms = e->t_k1;
    if (k2_ms) *k2_ms = e->t_k2;
    if (k3_ms) *k3_ms = e->t_k3;
    return PRB_OK;
}

extern "C" int prb_atmosphere_layer_timing(prb_engine *e, int32_t n_layers, float *k1_ms, float *k2_ms) {
    if (!e || !e->atm_layers || !e->timing || (int)e->t_layer_k1.size() != n_layers)
        return fail(PRB_ERR_STATE, "prb_atmosphere_layer_timing: enable timing, run prb_atmosphere, pass its layer count");
    for (int l = 0; l < n_layers; ++l) {
        if (k1_ms) k1_ms[l] = e->t_layer_k1[l];
        if (k2_ms) k2_ms[l] = e->t_layer_k2[l];
    }
    return PRB_OK;
}

extern "C" int prb_set_narrow_threshold(prb_engine *e, int64_t wm_below) {
    if (!e) return fail(PRB_ERR_ARG, "null engine");
    e->narrow_wm = wm_below < 0 ? 100 : wm_below;
    e->last.valid = false;                                      // record layout may change: redo the prepass
    return PRB_OK;
}

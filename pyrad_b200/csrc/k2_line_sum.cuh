// k2_line_sum.cuh -- K2: Gaussian / Lorentz / pseudo-Voigt line sum into the k(nu) grid.
//
// Replaces the scatter loop Isotope.createCrossSection pyradClasses.py:371-400 and the shapes
// pyradLineshape.py:32-76, in gather form:
//   out[i] = sum over lines l with |i - idx_l| <= W-2 of  A_l/(d^2+B_l) + G_l*exp2(C_l d^2),  d = i - idx_l
// (the window |d| <= W-2 is exactly `range(1, len(rightCurve)-1)` plus the centre sample, with the
// reference's bounds checks becoming "grid point exists").
//
// Design (B200 / sm_100a):
//   * persistent CTAs of 8 consumer warps + 1 TMA producer warp.  A CTA owns a tile of 256*P consecutive
//     grid points; tiles are handed out by a dynamic counter (work distribution only).  Every consumer
//     thread keeps P points in registers (FP32 partial sums flushed into FP64 accumulators every <= 64
//     lines).  No atomics touch data; the per-point summation order depends only on the tile geometry,
//     so results are deterministic and independent of how the grid is sharded (shards are tile aligned).
//   * the producer finds the wavenumber-sorted lines overlapping the tile's cutoff window with a 32-ary
//     warp-ballot search over the sorted index array and streams their records through a shared-memory
//     ring with TMA bulk copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP).  Slots are handed
//     over with full/empty mbarriers only: there is no CTA-wide barrier in the steady state.
//   * per slot every consumer warp classifies the staged (sorted) lines against its own 32*P-point span
//     by counting (warp reductions): lines whose window only partly covers the span are evaluated with
//     a per-point mask, the rest without.
//   * the Lorentz term of EVERY line goes through the paired reciprocal
//         A1/q1 + A2/q2 = (A1 q2 + A2 q1) * rcp(q1 q2),   q = d^2 + B
//     (one MUFU.RCP per TWO (line, point) pairs), written with Blackwell's packed FP32x2 instructions
//     (FADD2/FMUL2/FFMA2: two grid points per instruction), so a (line, point) pair costs 2 issue slots
//     of FP32-pipe work + 0.5 MUFU.  The Gaussian core G*exp2(C d^2) is a separate pass over the few
//     lines whose near zone (|d| <= Dg: beyond it the term is < 2e-7 of the same line's Lorentz term) meets the span.
//   * tensor cores are deliberately unused: this is not a dense contraction.
#pragma once
#include "common.cuh"
#include "k3_stream.cuh"

namespace prb {

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// TMA 1-D bulk copy shared -> global (SASS: UBLKCP); the destination may be peer memory over NVLink.  Completion
// is tracked per thread in bulk async-groups.
__device__ __forceinline__ void tma_bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void consumer_barrier() {          // the math warps only (the producer never joins)
    asm volatile("bar.sync 1, %0;" ::"n"(K2_CONSUMERS * 32) : "memory");
}

// First position in idx[lo, hi) whose value is >= key, found by the whole warp: every round each
// lane probes one of 32 evenly spaced positions and a ballot narrows the interval 32-fold.
__device__ __forceinline__ int warp_lower_bound(const int32_t *__restrict__ idx, int lo, int hi, long long key) {
    const int lane = threadIdx.x & 31;
    while (hi > lo) {
        const int span = hi - lo;
        const int step = (span + 31) >> 5;
        int pos = lo + (lane + 1) * step - 1;
        pos = pos < hi - 1 ? pos : hi - 1;
        const bool less = (long long)__ldg(idx + pos) < key;
        const unsigned m = __ballot_sync(0xffffffffu, less);
        const int c = __popc(m);                          // idx is sorted: m is a prefix mask
        int nlo = lo, nhi = hi;
        if (c > 0) { int p = lo + c * step - 1; p = p < hi - 1 ? p : hi - 1; nlo = p + 1; }
        if (c < 32) { int p = lo + (c + 1) * step - 1; p = p < hi - 1 ? p : hi - 1; nhi = p; }
        lo = nlo;
        hi = nhi;
    }
    return lo;
}

// Per-line records written by K1 (36 B/line).  Values needed as packed operands are stored twice so
// that one LDS.128 delivers them as aligned register pairs.
//   recA = {-fidx, -fidx, A, A}      fidx = line index relative to the shard's first point (exact integer)
//   recB = {B, B, G, C}
//   recD = Dg                         near-zone radius in grid points (-1: no Gaussian term)
// One layer of a (possibly multi-layer) launch.  A launch walks work items (layer, tile) handed out by one
// dynamic counter, layers in table order (the host sorts them widest window first, so the expensive items
// start first and the tail of the launch is made of cheap ones).
constexpr int K2_MAX_XSC = 8;   // resident xsc cross-section tables a layer can carry (prb_xsc_resident)
struct K2Layer {
    const float4 *recA;
    const float4 *recB;
    const float *recD;
    void *out;                 // this layer's output row (chunk-local index 0)
    double inv_scale;          // undoes K1's power-of-two scale, exactly
    int l_begin, l_end;        // line range prepared by K1 for this shard and window
    int wm;                    // W-2 clamped at 0: max |d| that still accumulates
    int pad;
    double xsc_w[K2_MAX_XSC];  // absCoef weight conc P / 1e4 / kB / T of every resident xsc table in this layer
};

// Optional fused single-layer epilogue (gas cell): the pointwise layer physics of K3 applied to the finished
// tile -- k -> T = exp(-k u), I = B_l + T (B_surface - B_l) -- and the finished spectra stored straight into
// every rank's gather buffer over NVLink peer memory (tile by tile, overlapping the remaining line sums).
constexpr int K2_MAX_PEERS = 8;
constexpr int K2_MAX_DST = K2_MAX_PEERS + 1;   // every rank's gather buffer + the caller's pinned host result buffers
struct K2Fuse {
    int enabled;
    int n_dst;                           // 1 (local only) .. K2_MAX_DST
    float neg_depth_log2e;               // -depth * log2(e)
    float c2_over_t, c2_over_tsurf;      // 100 h c / kB / T
    long long n_total;                   // points of the FULL grid (np.linspace axis)
    double x0, dx, x_last;
    float *rad[K2_MAX_DST];              // per destination: this rank's slot of the radiance gather buffer
    float *trans[K2_MAX_DST];
};

struct K2Args {
    const K2Layer *layers;     // device table, n_layers entries
    int n_layers;
    const int32_t *idx;        // sorted absolute grid index per line
    long long i_begin;         // absolute index of the shard's first grid point
    int n_chunk;               // grid points owned
    int n_tiles;               // tiles per layer handled by this launch
    int tile_base;             // first tile of this launch (sub-launches of a pipelined upload; 0 otherwise)
    const int2 *tile_bounds;   // thread-per-point kernels: staged line range per (layer, tile), from k2_tile_bounds
    int variant;               // PRB_K2_GENERAL / PRB_K2_CLASSED
    int out_mode;              // PRB_OUT_F64 / PRB_OUT_F32
    DevState *st;              // tile_counter of this launch
    K2Fuse fuse;
    // far-field variant (k2_line_sum<P, true>, P = 4 or 8): Lagrange weights of the span's points, [K2_FAR_NODES][32 P]
    // (node-major: a warp reads 32 consecutive doubles), and the node offsets from the span's first point (FP32; the
    // table is built from these rounded values)
    const double *far_lag;
    float far_delta[16];
    // level 2 of the far field: the same for a DOMAIN of K2_FAR2_SPANS consecutive spans, table [K2_FAR_NODES][domain]
    const double *far_lag2;
    float far_delta2[16];
    // resident xsc tables (pyradClasses.py:466-505, 707-712): sigma[t * xsc_ld + i] on the owned chunk; the finished line
    // sum of a layer gets  + sum_t sigma_t[i] * xsc_w[t]  in the epilogue, before it is rounded to the output type
    const double *xsc_sigma;
    long long xsc_ld;
    int n_xsc;
    // line-range parts (PRB_OPT_SPLIT_TILES; launches of a few waves only): every (layer, tile) is `parts` work items,
    // part s sums the staged chunks [s nch / parts, (s+1) nch / parts) of the tile's line range into its own FP64
    // partials (part_sums[item][TILE]); the CTA that finishes a tile's last part adds the partials in part order and
    // runs the epilogue.  part_count[layer * n_tiles + tile] counts finished parts (zeroed before the launch).
    int parts;
    double *part_sums;
    unsigned int *part_count;
};

// k of one grid point: the line sum plus the layer's xsc molecules (Layer.absCoef adds molecule by molecule, :707-712).
__device__ __forceinline__ double k2_add_xsc(const K2Args &a, const K2Layer *L, int i, double v) {
    for (int t = 0; t < a.n_xsc; ++t)
        v = __dadd_rn(v, __dmul_rn(__ldg(a.xsc_sigma + (long long)t * a.xsc_ld + i), __ldg(&L->xsc_w[t])));
    return v;
}

// Far-field evaluation of the Lorentz wings (variant PRB_K2_FARFIELD; spans of 128 and 256 points).  A line whose
// centre lies more than ONE span length from the centre of a warp's span, and whose window covers the whole span,
// contributes a function of the grid coordinate that is analytic on the span with poles at least that far away:
// its Chebyshev interpolant through K2_FAR_NODES = 16 nodes converges like (h / (D + sqrt(D^2 - h^2)))^nodes, h = span/2,
// D >= span -> (2 + sqrt 3)^-16 = 7e-10, times a constant of order ten for the line nearest the threshold.  The FP64
// model of this algorithm (oracle/farfield_model.py, tests/test_farfield_model.py) measures a worst case of 1e-8 of k on
// the BASELINE window classes.  Such lines are summed at the 16 nodes of the span -- 16 evaluations instead of 128 or
// 256 -- and the node sums are interpolated to the points once per tile.  Lines within 1.5 spans of the span's centre,
// lines whose window edge crosses the span, and every Gaussian core go through the exact per-point paths as before.
// (Round 1 shipped 8 nodes at a radius of two spans: worst case 6.5e-7, and a far class of 87 / 48 / 57 / 0 % of the
// pairs at 1013 / 250 / 150 / 60 hPa instead of 92 / 69 / 74 / 35 %.)
constexpr int K2_FAR_NODES = 16;
constexpr int K2_FAR_RADIUS_SPANS = 1;    // far = more than this many span lengths from the span centre
constexpr int K2_FAR_SUBGROUPS = 64 / K2_FAR_NODES;   // line sub-groups: a lane = (node pair, sub-group)
constexpr int K2_FAR_FLUSH = 16;          // triples (48 lines of one lane's chain) between FP64 flushes
// Level 2: K2_FAR2_SPANS consecutive spans (the warps w with equal w / K2_FAR2_SPANS) form a DOMAIN.  A line whose
// window covers the whole domain and whose centre lies more than one domain length from the domain's centre is summed
// at the 16 Chebyshev nodes of the DOMAIN -- same geometry ratio, same error bound as level 1 -- by the domain's warps
// together (each takes every K2_FAR2_SPANS-th group of lines), 16 evaluations per domain instead of 16 per span.  The
// warps' node sums meet in shared memory once per tile (fixed order) and are interpolated to the points with a second
// Lagrange table.  Every level-2 line is a level-1 far line of each span of the domain, so level 1 simply skips them.
#ifndef PRB_K2_FAR2_SPANS
#define PRB_K2_FAR2_SPANS 8
#endif
constexpr int K2_FAR2_SPANS = PRB_K2_FAR2_SPANS;
#ifndef PRB_K2_FAR2_ENABLE
#define PRB_K2_FAR2_ENABLE 1
#endif
constexpr bool K2_FAR2 = PRB_K2_FAR2_ENABLE != 0;      // development switch: level 1 only
// Level 2 pays for its once-per-tile meeting of the warps (a barrier and a second interpolation) only when most of the
// window lies beyond a domain: it runs for windows of at least this many domain lengths (cfg5's 25 cm-1 cutoff; the
// 5 cm-1 windows of cfg2 / cfg4 measured 15 % slower with it, profiles/r02_k2_farfield.txt).
constexpr int K2_FAR2_MIN_DOMAINS = 4;
static_assert(K2_CONSUMERS % K2_FAR2_SPANS == 0, "a tile is a whole number of level-2 domains");

// One ring slot: a chunk of staged line records plus its descriptor.
struct K2Desc {
    int tile0;      // shard-local index of the first point of the tile this chunk belongs to
    int cnt;        // staged lines to process (padding excluded)
    int flags;      // K2_FIRST | K2_LAST | K2_END
    int layer;      // index into K2Args::layers
    int part;       // which line-range part of the tile this work item is (0 when tiles are not split)
};
constexpr int K2_FIRST = 1, K2_LAST = 2, K2_END = 4;

struct K2Smem {
    float4 rA[K2_STAGES][K2_CHUNK];
    float4 rB[K2_STAGES][K2_CHUNK];
    float rD[K2_STAGES][K2_CHUNK];
    K2Desc desc[K2_STAGES];
    uint64_t full[K2_STAGES];      // producer -> consumers: TMA bytes landed (+ descriptor written)
    uint64_t empty[K2_STAGES];     // consumers -> producer: all consumer warps are done with the slot
    double far2[2][K2_CONSUMERS][K2_FAR_NODES];   // far-field level 2: every warp's node sums, double buffered by tile parity
    double2 faracc[2][K2_CONSUMERS * 32];         // far-field: every lane's two FP64 node sums, level 1 and level 2 (touched
                                                  // once per far_pass call: eight registers the hot loops get back)
    int part_last;                                // line-range parts: this CTA finished the tile's last outstanding part
};

// Dynamic shared memory of k2_line_sum<P>: K2Smem | FP64 accumulators (PRB_K2_ACC_SMEM) | peer staging (2 x TILE floats).
template <int P>
__host__ __device__ constexpr size_t K2_SMEM_BASE() {
    return ((sizeof(K2Smem) + (PRB_K2_ACC_SMEM ? sizeof(double) * P * K2_CONSUMERS * 32 : 0)) + 127) & ~size_t(127);
}
template <int P>
__host__ __device__ constexpr size_t K2_SMEM_BYTES(bool peer_staging) {
    return K2_SMEM_BASE<P>() + (peer_staging ? sizeof(float) * 2 * K2_CONSUMERS * 32 * P : 0);
}

// Per-thread state of a consumer: H = P/2 packed point pairs.
template <int H>
struct Acc {
    float2 fi[H];       // grid coordinates of the thread's points: (p even, p odd) = (fi0+64j, fi0+64j+32)
    float2 a32[H];      // FP32 partial sums
#if PRB_K2_ACC_SMEM
    // FP64 accumulators in shared memory (touched once per K2_FLUSH lines): 16 registers per thread traded for
    // more resident warps.  Point p of the thread sits at a64[p * K2_ACC_STRIDE] (conflict-free: lanes are adjacent).
    double *a64;
    __device__ __forceinline__ double &acc(int p) { return a64[p * (K2_CONSUMERS * 32)]; }
#else
    double a64[2 * H];  // FP64 accumulators
    __device__ __forceinline__ double &acc(int p) { return a64[p]; }
#endif
    __device__ __forceinline__ void flush() {
#pragma unroll
        for (int j = 0; j < H; ++j) {
            acc(2 * j) += (double)a32[j].x;
            acc(2 * j + 1) += (double)a32[j].y;
            a32[j] = make_float2(0.f, 0.f);
        }
    }
};

__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 lo2(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 &v) { return make_float2(v.z, v.w); }

// One triple-reciprocal step: lines 1, 2, 3 at the thread's 2*H points,
//   A1/q1 + A2/q2 + A3/q3 = (A1 q2 q3 + q1 (A2 q3 + A3 q2)) * rcp(q1 q2 q3)
// -- one MUFU per THREE (line, point) pairs and 13 packed FP32x2 instructions per six pairs, which balances
// the XU pipe (8 cycles per MUFU) against the FP32 pipe (~1.3 cycles per packed instruction, measured).
template <int H, bool MASKED>
__device__ __forceinline__ void triple_step(const float4 &a1, const float4 &a2, const float4 &a3, const float2 &B1,
                                            const float2 &B2, const float2 &B3, float wmf, Acc<H> &s) {
    const float2 nf1 = lo2(a1), nf2 = lo2(a2), nf3 = lo2(a3);
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float2 e1 = __fadd2_rn(s.fi[h], nf1);
        const float2 e2 = __fadd2_rn(s.fi[h], nf2);
        const float2 e3 = __fadd2_rn(s.fi[h], nf3);
        const float2 q1 = __ffma2_rn(e1, e1, B1);
        const float2 q2 = __ffma2_rn(e2, e2, B2);
        const float2 q3 = __ffma2_rn(e3, e3, B3);
        float2 A1 = hi2(a1), A2 = hi2(a2), A3 = hi2(a3);
        if (MASKED) {
            A1.x = fabsf(e1.x) <= wmf ? A1.x : 0.f;
            A1.y = fabsf(e1.y) <= wmf ? A1.y : 0.f;
            A2.x = fabsf(e2.x) <= wmf ? A2.x : 0.f;
            A2.y = fabsf(e2.y) <= wmf ? A2.y : 0.f;
            A3.x = fabsf(e3.x) <= wmf ? A3.x : 0.f;
            A3.y = fabsf(e3.y) <= wmf ? A3.y : 0.f;
        }
        const float2 p23 = __fmul2_rn(q2, q3);
        const float2 t = __ffma2_rn(A3, q2, __fmul2_rn(A2, q3));
        const float2 num = __ffma2_rn(q1, t, __fmul2_rn(A1, p23));
        const float2 den = __fmul2_rn(q1, p23);
        const float2 r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
        s.a32[h] = __ffma2_rn(num, r, s.a32[h]);
    }
}

__device__ __forceinline__ float2 ldB(const float4 *sB, int j) { return *reinterpret_cast<const float2 *>(sB + j); }

// Lorentz terms of lines [js, je), three lines per reciprocal, two points per instruction.  Two triples are
// kept in flight (X and Y): the records of the next X triple are requested before the math of Y and vice
// versa, so the shared-memory latency hides behind FP32 work without any register rotation.
// MASKED: the window |d| <= wm may cover only part of the span -> zero A per point outside it.
template <int H, bool MASKED>
__device__ __forceinline__ void lorentz_paired(const float4 *sA, const float4 *sB, int js, int je, float wmf,
                                               Acc<H> &s) {
    const int n = je - js;
    if (n <= 0) return;
    int j = js;
    const int n6 = (n / 6) * 6;
    const int jq = js + n6;                                         // lines handled six at a time
    if (n6) {
        float4 xa1 = sA[j], xa2 = sA[j + 1], xa3 = sA[j + 2], ya1 = sA[j + 3], ya2 = sA[j + 4], ya3 = sA[j + 5];
        float2 xb1 = ldB(sB, j), xb2 = ldB(sB, j + 1), xb3 = ldB(sB, j + 2);
        float2 yb1 = ldB(sB, j + 3), yb2 = ldB(sB, j + 4), yb3 = ldB(sB, j + 5);
        int since = 0;
        for (j += 6; j < jq; j += 6) {
            triple_step<H, MASKED>(xa1, xa2, xa3, xb1, xb2, xb3, wmf, s);
            xa1 = sA[j]; xa2 = sA[j + 1]; xa3 = sA[j + 2];
            xb1 = ldB(sB, j); xb2 = ldB(sB, j + 1); xb3 = ldB(sB, j + 2);
            triple_step<H, MASKED>(ya1, ya2, ya3, yb1, yb2, yb3, wmf, s);
            ya1 = sA[j + 3]; ya2 = sA[j + 4]; ya3 = sA[j + 5];
            yb1 = ldB(sB, j + 3); yb2 = ldB(sB, j + 4); yb3 = ldB(sB, j + 5);
            if (++since == K2_FLUSH / 6) { s.flush(); since = 0; }
        }
        triple_step<H, MASKED>(xa1, xa2, xa3, xb1, xb2, xb3, wmf, s);
        triple_step<H, MASKED>(ya1, ya2, ya3, yb1, yb2, yb3, wmf, s);
    }
    for (j = jq; j < je; j += 3) {                                  // tail: pad the triple with zero-weight copies
        const float4 a1 = sA[j];
        const float2 B1 = ldB(sB, j);
        const float4 z = make_float4(a1.x, a1.y, 0.f, 0.f);
        const bool h2 = j + 1 < je, h3 = j + 2 < je;
        triple_step<H, MASKED>(a1, h2 ? sA[j + 1] : z, h3 ? sA[j + 2] : z, B1, h2 ? ldB(sB, j + 1) : B1,
                               h3 ? ldB(sB, j + 2) : B1, wmf, s);
    }
    s.flush();
}

#ifndef PRB_K2_GAUSS_PAIR_H
#define PRB_K2_GAUSS_PAIR_H 2
#endif
constexpr int K2_GAUSS_PAIR_H = PRB_K2_GAUSS_PAIR_H;   // spans of up to this many 64-point blocks take two hit lines per trip

// Gaussian cores of lines [js, je): only lines whose near zone meets the span (warp-uniform test).
// MASKED: the window |d| <= wm may cover only part of the span -> zero G per point outside it.
template <int H, bool MASKED>
__device__ __forceinline__ void gauss_pass(const float4 *sA, const float4 *sB, const float *sD, int js, int je,
                                           float wbf, float we1f, float wmf, Acc<H> &s) {
    int since = 0;
    // 32 candidates per step: every lane tests one line's near zone against the span, a ballot collects the hits and
    // the warp then walks the hits in ascending line order (the candidates come from the slot's LARGEST radius, so
    // most of them miss: testing them one by one cost as much as evaluating the hits)
    for (int jb = js; jb < je; jb += 32) {
        const int jt = jb + (int)(threadIdx.x & 31);
        bool hit = false;
        if (jt < je) {
            const float dg = sD[jt];
            const float f = -sA[jt].x;
            hit = (dg >= 0.f) && (f + dg >= wbf) && (f - dg <= we1f);
        }
        unsigned int m = __ballot_sync(0xffffffffu, hit);
        // one hit line on all H blocks of the span: term(b, nf) -> a32 += G exp2(C d^2)
        auto one = [&](const float4 &b, float nf1) {
            const float2 nf = splat(nf1), C2 = splat(b.w);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float2 e = __fadd2_rn(s.fi[h], nf);
                const float2 arg = __fmul2_rn(C2, __fmul2_rn(e, e));
                float2 g = splat(b.z);
                if (MASKED) {
                    g.x = fabsf(e.x) <= wmf ? b.z : 0.f;
                    g.y = fabsf(e.y) <= wmf ? b.z : 0.f;
                }
                s.a32[h] = __ffma2_rn(g, make_float2(ex2_approx(arg.x), ex2_approx(arg.y)), s.a32[h]);
            }
        };
        while (m) {
            const int j = jb + __ffs(m) - 1;
            m &= m - 1;
            if (H <= K2_GAUSS_PAIR_H && m) {
                // spans of few blocks: two hit lines per trip, so that twice as many exponentials are in flight (the sums
                // still take the lines in ascending order: same roundings as one line per trip)
                const int j2 = jb + __ffs(m) - 1;
                m &= m - 1;
                const float4 b1 = sB[j], b2 = sB[j2];
                const float n1 = sA[j].x, n2 = sA[j2].x;
                const float2 nfa = splat(n1), nfb = splat(n2), Ca = splat(b1.w), Cb = splat(b2.w);
                float2 ta[H], tb[H], ga[H], gb[H];
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float2 ea = __fadd2_rn(s.fi[h], nfa), eb = __fadd2_rn(s.fi[h], nfb);
                    const float2 xa = __fmul2_rn(Ca, __fmul2_rn(ea, ea)), xb = __fmul2_rn(Cb, __fmul2_rn(eb, eb));
                    ga[h] = splat(b1.z);
                    gb[h] = splat(b2.z);
                    if (MASKED) {
                        ga[h].x = fabsf(ea.x) <= wmf ? b1.z : 0.f;
                        ga[h].y = fabsf(ea.y) <= wmf ? b1.z : 0.f;
                        gb[h].x = fabsf(eb.x) <= wmf ? b2.z : 0.f;
                        gb[h].y = fabsf(eb.y) <= wmf ? b2.z : 0.f;
                    }
                    ta[h] = make_float2(ex2_approx(xa.x), ex2_approx(xa.y));
                    tb[h] = make_float2(ex2_approx(xb.x), ex2_approx(xb.y));
                }
#pragma unroll
                for (int h = 0; h < H; ++h) s.a32[h] = __ffma2_rn(gb[h], tb[h], __ffma2_rn(ga[h], ta[h], s.a32[h]));
                since += 2;
                if (since >= K2_FLUSH) { s.flush(); since = 0; }
                continue;
            }
            one(sB[j], sA[j].x);
            if (++since >= K2_FLUSH) { s.flush(); since = 0; }
        }
    }
    if (since) s.flush();
}

// Far-field pass over lines [js, je): lane = (node pair kp = lane & 7, line sub-group lane >> 3).  A lane evaluates the
// Lorentz terms of every 4th line at its two nodes, three lines per reciprocal, in ascending line order; the FP32
// partial sums are flushed into the lane's FP64 node accumulators.  d = (wb - idx) + delta: the first sum is an exact
// small integer, so the node's fractional offset survives FP32 at any grid size.
struct FarAcc {
    double v0, v1;
};
__device__ __forceinline__ float2 far_triple(const float4 &a1, const float4 &a2, const float4 &a3, const float2 &B1,
                                             const float2 &B2, const float2 &B3, const float2 &A1, const float2 &A2,
                                             const float2 &A3, const float2 &wb2, const float2 &del, const float2 &part) {
    const float2 e1 = __fadd2_rn(__fadd2_rn(wb2, lo2(a1)), del);
    const float2 e2 = __fadd2_rn(__fadd2_rn(wb2, lo2(a2)), del);
    const float2 e3 = __fadd2_rn(__fadd2_rn(wb2, lo2(a3)), del);
    const float2 q1 = __ffma2_rn(e1, e1, B1);
    const float2 q2 = __ffma2_rn(e2, e2, B2);
    const float2 q3 = __ffma2_rn(e3, e3, B3);
    const float2 p23 = __fmul2_rn(q2, q3);
    const float2 t = __ffma2_rn(A3, q2, __fmul2_rn(A2, q3));
    const float2 num = __ffma2_rn(q1, t, __fmul2_rn(A1, p23));
    const float2 den = __fmul2_rn(q1, p23);
    const float2 r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
    return __ffma2_rn(num, r, part);
}
template <int G>
__device__ __forceinline__ void far_pass(const float4 *sA, const float4 *sB, int js, int je, float wbf, float2 del,
                                         double2 *acc, int sub) {
    const int n = je - js;
    if (n <= 0) return;                                     // warp-uniform
    const double2 acc0 = *acc;
    FarAcc fa{acc0.x, acc0.y};
    const float2 wb2 = splat(wbf);
    const float2 zero = make_float2(0.f, 0.f);
    // whole groups of 3 G lines: every lane's triple exists, nothing to clamp or select; FP32 partial sums of at most
    // K2_FAR_FLUSH triples go into the FP64 node accumulators
    int full = n / (3 * G);
    const float4 *pa = sA + js + sub;
    const float4 *pb = sB + js + sub;
    while (full > 0) {
        const int m = min(full, K2_FAR_FLUSH);
        float2 part = zero;
#pragma unroll 2
        for (int i = 0; i < m; ++i) {
            const float4 a1 = pa[0], a2 = pa[G], a3 = pa[2 * G];
            const float2 B1 = ldB(pb, 0), B2 = ldB(pb, G), B3 = ldB(pb, 2 * G);
            part = far_triple(a1, a2, a3, B1, B2, B3, hi2(a1), hi2(a2), hi2(a3), wb2, del, part);
            pa += 3 * G;
            pb += 3 * G;
        }
        fa.v0 += (double)part.x;
        fa.v1 += (double)part.y;
        full -= m;
    }
    // the ragged last group: missing lines are padded with A = 0 copies of the last line
    const int j0 = js + (n / (3 * G)) * (3 * G);
    if (j0 < je) {
        const int j1 = j0 + sub, j2 = j1 + G, j3 = j1 + 2 * G;
        const int k1 = min(j1, je - 1), k2 = min(j2, je - 1), k3 = min(j3, je - 1);
        const float4 a1 = sA[k1], a2 = sA[k2], a3 = sA[k3];
        const float2 A1 = j1 < je ? hi2(a1) : zero, A2 = j2 < je ? hi2(a2) : zero, A3 = j3 < je ? hi2(a3) : zero;
        const float2 part = far_triple(a1, a2, a3, ldB(sB, k1), ldB(sB, k2), ldB(sB, k3), A1, A2, A3, wb2, del, zero);
        fa.v0 += (double)part.x;
        fa.v1 += (double)part.y;
    }
    *acc = make_double2(fa.v0, fa.v1);
}

// Plain one-line-at-a-time evaluation of every staged line (variant PRB_K2_GENERAL): the A/B check
// for the classed path, and a measurement of what the naive formulation costs.
template <int H>
__device__ __forceinline__ void general_all(const float4 *sA, const float4 *sB, int cnt, float wmf, Acc<H> &s) {
    for (int jb = 0; jb < cnt; jb += K2_FLUSH) {
        const int jend = min(jb + K2_FLUSH, cnt);
        for (int j = jb; j < jend; ++j) {
            const float4 a = sA[j], b = sB[j];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float ex = s.fi[h].x + a.x, ey = s.fi[h].y + a.x;
                const float e2x = ex * ex, e2y = ey * ey;
                float tx = a.z * rcp_approx(e2x + b.x), ty = a.z * rcp_approx(e2y + b.x);
                tx = fmaf(b.z, ex2_approx(b.w * e2x), tx);
                ty = fmaf(b.z, ex2_approx(b.w * e2y), ty);
                s.a32[h].x += (fabsf(ex) <= wmf) ? tx : 0.f;
                s.a32[h].y += (fabsf(ey) <= wmf) ? ty : 0.f;
            }
        }
        s.flush();
    }
}

// SPLIT: the launch hands tiles out as line-range parts (PRB_OPT_SPLIT_TILES).  A template parameter, not a run-time
// test: with the hand-over code merely present in the epilogue the unsplit fused gas-cell step measured 9 % slower.
template <int P, bool FAR, bool SPLIT>
__device__ __forceinline__ void k2_line_sum_body(const K2Args &a) {
    static_assert(P >= 2 && P % 2 == 0, "points per thread must be even (packed FP32x2)");
    static_assert(!FAR || P == 4 || P == 8, "far-field tables exist for 128- and 256-point spans");
    constexpr int H = P / 2;
    constexpr int TILE = K2_CONSUMERS * 32 * P;
    constexpr int SPAN = 32 * P;                      // points per consumer warp
    extern __shared__ __align__(128) unsigned char smem_raw[];
    K2Smem &sm = *reinterpret_cast<K2Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < K2_STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], K2_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == K2_CONSUMERS) {
        // ------------------------------------------------------------------ producer warp
        uint32_t it = 0;
        auto slot_acquire = [&](uint32_t &stage) {
            stage = it % K2_STAGES;
            mbar_wait_parked(&sm.empty[stage], ((it / K2_STAGES) & 1) ^ 1);   // fresh barrier: passes at once
        };
        const int parts = (SPLIT && a.parts > 1) ? a.parts : 1;
        const int n_items = a.n_layers * a.n_tiles * parts;
        while (true) {
            int item = 0;
            if (lane == 0) item = (int)atomicAdd(&a.st->tile_counter, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) break;
            // part-major order, the parts that can carry the most work first: in the far-field variant the lines near
            // the tile (the middle of its line range) and the ones whose window edge crosses it (both ends) are evaluated
            // point by point, everything between at a few nodes -- handing the heavy parts out first keeps the launch's
            // tail short; for the exact kernel all parts cost the same and the order does not matter
            const int n_lt = a.n_layers * a.n_tiles;
            const int pidx = item / n_lt;
            const int lt = item - pidx * n_lt;
            const int half = parts >> 1, step = pidx >> 2, r = pidx & 3;
            const int part = parts < 4 ? pidx : (r == 0 ? half - 1 - step : r == 1 ? half + step : r == 2 ? step : parts - 1 - step);
            const int layer = lt / a.n_tiles;
            const int tile = a.tile_base + lt - layer * a.n_tiles;
            const K2Layer *L = a.layers + layer;
            const int wm = __ldg(&L->wm), l_begin = __ldg(&L->l_begin), l_end = __ldg(&L->l_end);
            const float4 *recA = L->recA;
            const float4 *recB = L->recB;
            const float *recD = L->recD;
            const int tile0 = tile * TILE;
            const long long k_lo = a.i_begin + tile0 - wm;
            const long long k_hi = a.i_begin + tile0 + TILE - 1 + wm + 1;
            int lo = warp_lower_bound(a.idx, l_begin, l_end, k_lo);
            const int hi = warp_lower_bound(a.idx, lo, l_end, k_hi);
            lo &= ~3;                                  // 16-byte alignment of the float stream
            const int nch = hi > lo ? (hi - lo + K2_CHUNK - 1) / K2_CHUNK : 1;   // empty tile: one empty chunk
            // this work item's chunks of the tile (all of them unless the tile is split into line-range parts)
            const int c0 = (int)((long long)part * nch / parts), c1 = (int)((long long)(part + 1) * nch / parts);
            const int cn = c1 > c0 ? c1 - c0 : 1;      // a part without chunks still reports in with an empty one
            for (int k = 0; k < cn; ++k, ++it) {
                const int c = c0 + k;
                uint32_t stage;
                slot_acquire(stage);
                if (lane == 0) {
                    const int first = lo + c * K2_CHUNK;
                    const int cnt = c1 > c0 ? max(min(K2_CHUNK, hi - first), 0) : 0;
                    sm.desc[stage] = K2Desc{tile0, cnt, (k == 0 ? K2_FIRST : 0) | (k == cn - 1 ? K2_LAST : 0), layer, part};
                    if (cnt > 0) {
                        const uint32_t ce = (uint32_t)((cnt + 3) & ~3);   // padding records exist past l_end
                        mbar_expect_tx(&sm.full[stage], ce * 36u);
                        tma_bulk_g2s(sm.rA[stage], recA + first, ce * 16u, &sm.full[stage]);
                        tma_bulk_g2s(sm.rB[stage], recB + first, ce * 16u, &sm.full[stage]);
                        tma_bulk_g2s(sm.rD[stage], recD + first, ce * 4u, &sm.full[stage]);
                    } else {
                        mbar_arrive(&sm.full[stage]);
                    }
                }
                __syncwarp();
            }
        }
        uint32_t stage;
        slot_acquire(stage);
        if (lane == 0) {
            sm.desc[stage] = K2Desc{0, 0, K2_END, 0, 0};
            mbar_arrive(&sm.full[stage]);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    float wmf = 0.f;
    Acc<H> s;
#pragma unroll
    for (int h = 0; h < H; ++h) { s.a32[h] = make_float2(0.f, 0.f); s.fi[h] = make_float2(0.f, 0.f); }
#if PRB_K2_ACC_SMEM
    s.a64 = reinterpret_cast<double *>(smem_raw + sizeof(K2Smem)) + tid;
#endif
#pragma unroll
    for (int p = 0; p < P; ++p) s.acc(p) = 0.0;
    int wb = 0;
    float wbf = 0.f, we1f = 0.f;
    double2 *far = &sm.faracc[0][tid];                     // FAR: this lane's two node sums of the current tile
    double2 *far2 = &sm.faracc[1][tid];                    // ... and of its level-2 domain (its share of the lines)
    float dbf = 0.f;                                       // first point of the warp's level-2 domain
    uint32_t tile_par = 0;                                 // parity of the tiles this CTA has finished (sm.far2 buffer)
    bool use2 = false;                                     // level 2 on for this tile's layer (uniform over the CTA)

    for (uint32_t it = 0;; ++it) {
        const uint32_t stage = it % K2_STAGES;
        mbar_wait(&sm.full[stage], (it / K2_STAGES) & 1);
        const K2Desc d = sm.desc[stage];
        if (d.flags & K2_END) {
            // the peer stores of the last tile must have left shared memory before the CTA retires
            if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            break;
        }
        if (d.flags & K2_FIRST) {
            wmf = (float)__ldg(&a.layers[d.layer].wm);
            wb = d.tile0 + warp * SPAN;                // first point of this warp's span
            wbf = (float)wb;
            we1f = (float)(wb + SPAN - 1);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                s.a32[h] = make_float2(0.f, 0.f);
                s.fi[h] = make_float2((float)(wb + lane + 64 * h), (float)(wb + lane + 64 * h + 32));
            }
#pragma unroll
            for (int p = 0; p < P; ++p) s.acc(p) = 0.0;
            if (FAR) {
                *far = make_double2(0.0, 0.0);
                *far2 = make_double2(0.0, 0.0);
            }
            dbf = (float)(d.tile0 + (warp / K2_FAR2_SPANS) * (K2_FAR2_SPANS * SPAN));
            use2 = FAR && K2_FAR2 && wmf >= (float)(K2_FAR2_MIN_DOMAINS * K2_FAR2_SPANS * SPAN);
        }
        const int cnt = d.cnt;
        const float4 *sA = sm.rA[stage];
        const float4 *sB = sm.rB[stage];
        const float *sD = sm.rD[stage];

        if (a.variant == 0) {
            general_all<H>(sA, sB, cnt, wmf, s);
        } else if (cnt > 0) {
            // near-zone radius of THIS staged chunk (max over its lines): a function of the tile geometry
            // only, so the classes -- and the FP32 rounding -- do not depend on the sharding.
            float dgl = 0.f;
            for (int j = lane; j < cnt; j += 32) dgl = fmaxf(dgl, sD[j]);
            const float dgmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(dgl)));
            // class boundaries of the sorted staged lines relative to this warp's span, by counting
            const float t0 = wbf - wmf;                        // idx <  t0 : window ends before the span
            const float t5 = we1f + 1.f + wmf;                 // idx >= t5 : window starts after the span
            float t1 = we1f - wmf, t4 = wbf + wmf + 1.f;       // full cover: t1 <= idx < t4
            if (t1 >= t4) { t1 = t5; t4 = t5; }                // window narrower than the span: all masked
            const float g0 = fmaxf(wbf - dgmax, t0);           // Gaussian cores can only come from [g0, g1)
            const float g1 = fminf(we1f + 1.f + dgmax, t5);
            // FAR: idx < tfl or idx > tfr, i.e. more than K2_FAR_RADIUS_SPANS span lengths from the span centre wb + (SPAN-1)/2; integer
            // thresholds, exact in FP32, so the split does not depend on the shard origin
            const float tfl = wbf + (float)((SPAN - 1) / 2 - K2_FAR_RADIUS_SPANS * SPAN);
            const float tfr = wbf + (float)(SPAN / 2 + K2_FAR_RADIUS_SPANS * SPAN);
            int b1, b4, bg0, bg1;
            if (FAR) {
                // The staged lines are sorted, so "how many lie below t" is a search, not a count: one sample per block
                // of 32 lines (loaded once, shared by all thresholds), a ballot picks the block the boundary falls in, a
                // second ballot over that block's 32 lines places it -- two shared-memory reads and two ballots per
                // threshold instead of a pass over the whole slot.  Same integers as counting.  (Measured: worth 4 % of the
                // far-field kernel; in the exact kernel it changed the register allocation for the worse, so that one
                // keeps counting -- profiles/r01_k2_experiments.txt.)  Every boundary is looked up right before the pass
                // that needs it, so few of them are alive at a time.
                static_assert(K2_CHUNK <= 1024, "one level-1 sample per lane");
                const float f1 = -sA[min(32 * lane + 31, cnt - 1)].x;
                const bool v1 = 32 * lane < cnt;
                auto below = [&](float t) -> int {
                    const int nb = __popc(__ballot_sync(0xffffffffu, v1 && f1 < t));   // whole blocks below t (a prefix)
                    const int j2 = 32 * nb + lane;
                    const bool l2 = j2 < cnt && -sA[min(j2, cnt - 1)].x < t;
                    return min(32 * nb, cnt) + __popc(__ballot_sync(0xffffffffu, l2));
                };
                constexpr int S2 = K2_FAR2_SPANS * SPAN;
                const int kp = lane & (K2_FAR_NODES / 2 - 1);
                const int sub = lane / (K2_FAR_NODES / 2);
                const float2 del = make_float2(a.far_delta[2 * kp], a.far_delta[2 * kp + 1]);
                const float2 del2 = make_float2(a.far_delta2[2 * kp], a.far_delta2[2 * kp + 1]);
                const int sub2 = (warp % K2_FAR2_SPANS) * K2_FAR_SUBGROUPS + sub;
                // level 2 (use2): the same tests for the warp's domain [db, db + S2), identical in the domain's warps;
                // full cover of the domain: u1 <= idx < u4.  Its lines [c1, cfl) and [cfr, c4) lie inside the level-1 far
                // ranges [b1, bfl) and [bfr, b4): level 1 takes what is left of those on either side.
                const float u1 = dbf + (float)(S2 - 1) - wmf, u4 = dbf + wmf + 1.f;
                const bool l2on = use2 && u1 < u4;
                b1 = below(t1);
                lorentz_paired<H, true>(sA, sB, below(t0), b1, wmf, s);           // window edge crosses the span (left)
                b4 = below(t4);
                // full-cover lines [b1, b4) split by distance from the span: far left | near | far right
                const int bfl = min(max(below(tfl), b1), b4);
                if (l2on) {
                    const int c4 = below(u4);
                    const int c1 = min(below(u1), c4);
                    const int cfl = min(max(below(dbf + (float)((S2 - 1) / 2 - K2_FAR_RADIUS_SPANS * S2)), c1), c4);
                    const int l1a = min(max(c1, b1), bfl), l1b = min(max(cfl, l1a), bfl);
                    far_pass<K2_FAR_SUBGROUPS>(sA, sB, b1, l1a, wbf, del, far, sub);
                    far_pass<K2_FAR2_SPANS * K2_FAR_SUBGROUPS>(sA, sB, c1, cfl, dbf, del2, far2, sub2);
                    far_pass<K2_FAR_SUBGROUPS>(sA, sB, l1b, bfl, wbf, del, far, sub);
                } else {
                    far_pass<K2_FAR_SUBGROUPS>(sA, sB, b1, bfl, wbf, del, far, sub);
                }
                const int bfr = min(max(below(tfr + 1.f), bfl), b4);               // idx <= tfr  <=>  idx < tfr + 1 (integers)
                lorentz_paired<H, false>(sA, sB, bfl, bfr, wmf, s);
                if (l2on) {
                    const int c4 = below(u4);
                    const int c1 = min(below(u1), c4);
                    const int cfl = min(max(below(dbf + (float)((S2 - 1) / 2 - K2_FAR_RADIUS_SPANS * S2)), c1), c4);
                    const int cfr = min(max(below(dbf + (float)(S2 / 2 + K2_FAR_RADIUS_SPANS * S2) + 1.f), cfl), c4);
                    const int r1a = min(max(cfr, bfr), b4), r1b = min(max(c4, r1a), b4);
                    far_pass<K2_FAR_SUBGROUPS>(sA, sB, bfr, r1a, wbf, del, far, sub);
                    far_pass<K2_FAR2_SPANS * K2_FAR_SUBGROUPS>(sA, sB, cfr, c4, dbf, del2, far2, sub2);
                    far_pass<K2_FAR_SUBGROUPS>(sA, sB, r1b, b4, wbf, del, far, sub);
                } else {
                    far_pass<K2_FAR_SUBGROUPS>(sA, sB, bfr, b4, wbf, del, far, sub);
                }
                lorentz_paired<H, true>(sA, sB, b4, below(t5), wmf, s);           // window edge crosses the span (right)
                bg0 = below(g0);
                bg1 = below(g1);
            } else {
                int c0 = 0, c1 = 0, c4 = 0, c5 = 0, cg0 = 0, cg1 = 0;
                for (int j = lane; j < cnt; j += 32) {
                    const float f = -sA[j].x;
                    c0 += f < t0; c1 += f < t1; c4 += f < t4; c5 += f < t5; cg0 += f < g0; cg1 += f < g1;
                }
                const int b0 = __reduce_add_sync(0xffffffffu, c0);
                b1 = __reduce_add_sync(0xffffffffu, c1);
                b4 = __reduce_add_sync(0xffffffffu, c4);
                const int b5 = __reduce_add_sync(0xffffffffu, c5);
                bg0 = __reduce_add_sync(0xffffffffu, cg0);
                bg1 = __reduce_add_sync(0xffffffffu, cg1);
                lorentz_paired<H, true>(sA, sB, b0, b1, wmf, s);
                lorentz_paired<H, false>(sA, sB, b1, b4, wmf, s);
                lorentz_paired<H, true>(sA, sB, b4, b5, wmf, s);
            }
            if (dgmax > 0.f) {                                  // [bg0, bg1) lies inside [b0, b5)
                gauss_pass<H, true>(sA, sB, sD, bg0, min(bg1, b1), wbf, we1f, wmf, s);
                gauss_pass<H, false>(sA, sB, sD, max(bg0, b1), min(bg1, b4), wbf, we1f, wmf, s);
                gauss_pass<H, true>(sA, sB, sD, max(bg0, b4), bg1, wbf, we1f, wmf, s);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);  // slot may be refilled

        if (d.flags & K2_LAST) {
            // epilogue: undo the power-of-two scale exactly and store (coalesced 32-point rows)
            const K2Layer *L = a.layers + d.layer;
            const double inv_scale = __ldg(&L->inv_scale);
            void *out = L->out;
            // Multi-GPU: a full tile's finished spectra are staged in shared memory and pushed to every rank's
            // gather buffer with TMA bulk stores (asynchronous: no thread waits on NVLink, the next tile's line sum
            // starts at once).  Ragged last tiles and single-destination runs store directly.
            const bool bulk = a.fuse.enabled && a.fuse.n_dst > 1 && d.tile0 + TILE <= a.n_chunk;
            float *stage_rad = reinterpret_cast<float *>(smem_raw + K2_SMEM_BASE<P>());
            float *stage_tr = stage_rad + TILE;
            if (bulk) {
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging free again
                consumer_barrier();
            }
            if (FAR) {
                // node sums: add the line sub-groups (fixed tree), then interpolate to the thread's points -- node by
                // node, each node's sum broadcast from the lane that holds it, the Lagrange weights read as 32
                // consecutive doubles per (node, row of the span) -- straight into the FP64 accumulators
                double v0 = far->x, v1 = far->y, w0 = far2->x, w1 = far2->y;
#pragma unroll
                for (int o = K2_FAR_NODES / 2; o < 32; o <<= 1) {
                    v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                    w0 += __shfl_xor_sync(0xffffffffu, w0, o);
                    w1 += __shfl_xor_sync(0xffffffffu, w1, o);
                }
                // level 2: the domain's warps pool their node sums (fixed order), lane k holds node k of the domain
                if (use2 && lane < K2_FAR_NODES / 2) {
                    sm.far2[tile_par][warp][2 * lane] = w0;
                    sm.far2[tile_par][warp][2 * lane + 1] = w1;
                }
                if (use2) consumer_barrier();
                double dn = 0.0;
                if (use2 && lane < K2_FAR_NODES) {
                    const int wd = (warp / K2_FAR2_SPANS) * K2_FAR2_SPANS;
#pragma unroll
                    for (int q = 0; q < K2_FAR2_SPANS; ++q) dn += sm.far2[tile_par][wd + q][lane];
                }
                if (use2) tile_par ^= 1u;
                constexpr int S2 = K2_FAR2_SPANS * SPAN;
                const double *lw2base = a.far_lag2 + (warp % K2_FAR2_SPANS) * SPAN + lane;
#pragma unroll 2
                for (int k = 0; k < K2_FAR_NODES; ++k) {
                    const double fk = __shfl_sync(0xffffffffu, (k & 1) ? v1 : v0, k >> 1);
                    const double *lw = a.far_lag + k * SPAN + lane;
#pragma unroll
                    for (int p = 0; p < P; ++p) s.acc(p) = fma(fk, __ldg(lw + 32 * p), s.acc(p));
                    if (use2) {
                        const double gk = __shfl_sync(0xffffffffu, dn, k);
                        const double *lw2 = lw2base + k * S2;
#pragma unroll
                        for (int p = 0; p < P; ++p) s.acc(p) = fma(gk, __ldg(lw2 + 32 * p), s.acc(p));
                    }
                }
            }
            bool finish = true;                            // this CTA runs the tile's epilogue (always, unless tiles are split)
            if (SPLIT && a.parts > 1) {
                // line-range parts: park this part's sums (far-field contributions included); the CTA that completes the
                // tile adds all parts in part order -- a fixed order, whoever finishes last
                const int lt = d.layer * a.n_tiles + (d.tile0 / TILE - a.tile_base);
                double *mine = a.part_sums + ((size_t)lt * a.parts + d.part) * TILE + warp * SPAN + lane;
#pragma unroll
                for (int p = 0; p < P; ++p) __stcg(mine + 32 * p, s.acc(p));
                __threadfence();
                consumer_barrier();
                if (tid == 0) sm.part_last = atomicAdd(a.part_count + lt, 1u) == (unsigned int)(a.parts - 1);
                consumer_barrier();
                finish = sm.part_last != 0;
                if (finish) {
                    __threadfence();
                    const double *all = a.part_sums + (size_t)lt * a.parts * TILE + warp * SPAN + lane;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        double v = 0.0;
                        for (int q = 0; q < a.parts; ++q) v += __ldcg(all + (size_t)q * TILE + 32 * p);
                        s.acc(p) = v;
                    }
                }
            }
            if (finish) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const int i = wb + 32 * p + lane;
                    if (i < a.n_chunk) {
                        const double v = k2_add_xsc(a, L, i, s.acc(p) * inv_scale);
                        if (a.out_mode == 0) reinterpret_cast<double *>(out)[i] = v;
                        else reinterpret_cast<float *>(out)[i] = (float)v;
                        if (a.fuse.enabled) {
                            // same operations, in the same order, as k3_fold_f32 with one layer
                            const double x = axis_value(a.i_begin + i, a.fuse.n_total, a.fuse.x0, a.fuse.dx, a.fuse.x_last);
                            const float nu = (float)x;
                            const float a3 = (float)(2E8 * hPlanck * (cLight * cLight) * (x * x * x));
                            const float e = (float)v * a.fuse.neg_depth_log2e;
                            const float b = planck_f32(a3, a.fuse.c2_over_t * nu);
                            const float rad = k3_fold_step(planck_f32(a3, a.fuse.c2_over_tsurf * nu), e, b);
                            const float tr = exp2f(0.f + e);
                            if (bulk) {
                                stage_rad[i - d.tile0] = rad;
                                stage_tr[i - d.tile0] = tr;
                            } else {
#pragma unroll 1
                                for (int dst = 0; dst < a.fuse.n_dst; ++dst) {
                                    a.fuse.rad[dst][i] = rad;      // own slot of every rank's gather buffer
                                    a.fuse.trans[dst][i] = tr;
                                }
                            }
                        }
                    }
                }
                if (bulk) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    consumer_barrier();
                    if (tid == 0) {
#pragma unroll 1
                        for (int dst = 0; dst < a.fuse.n_dst; ++dst) {
                            tma_bulk_s2g(a.fuse.rad[dst] + d.tile0, stage_rad, TILE * 4u);
                            tma_bulk_s2g(a.fuse.trans[dst] + d.tile0, stage_tr, TILE * 4u);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
        }
    }
}

// The exact kernel: 96 registers (2 CTAs x 288 threads, ptxas' own bound for that launch shape).
template <int P, bool SPLIT = false>
__global__ void __launch_bounds__(K2_THREADS, K2_MIN_CTAS) k2_line_sum(const K2Args a) {
    k2_line_sum_body<P, false, SPLIT>(a);
}
// The far-field variant (same launch shape: 96 registers is the most that lets two CTAs share an SM -- the register
// file is per SM sub-partition, 16 384 each, and two CTAs put five warps on one of them; 112 was tried and halves the
// occupancy).
template <int P, bool SPLIT = false>
__global__ void __launch_bounds__(K2_THREADS, K2_MIN_CTAS) k2_line_sum_far(const K2Args a) {
    k2_line_sum_body<P, true, SPLIT>(a);
}

}  // namespace prb

// k2_narrow.cuh -- K2 for NARROW cutoff windows (upper-atmosphere layers: W-2 of a few to ~200 points).
//
// Same sum as k2_line_sum (pyradClasses.py:371-400 in gather form), different mapping: when the window is
// narrower than a warp's span, broadcasting every staged line to all lanes wastes most lanes, so here each
// THREAD owns grid points and walks only the lines inside its own window.  A CTA owns 1024 consecutive
// points; the sorted lines reaching the tile are staged in shared memory with TMA bulk copies (chunks of
// 2048 records); per point the first line of the window is found by binary search over the staged (sorted)
// indices and the walk stops at the first line past the window -- the window IS the loop range, so there
// are no masks.  FP32 terms, flushed into an FP64 accumulator every 64 lines; fixed ascending line order per
// point (deterministic, independent of sharding).
//
// Compact records written by K1 for this kernel:  recA = {-fidx, A, B, G},  recD = C.
#pragma once
#include "common.cuh"
#include "k2_line_sum.cuh"

namespace prb {

// Line range [lo, hi) of every (layer, tile) of a thread-per-point launch, one thread per tile: the two binary searches
// over the global index array are a chain of dependent loads (~4 us) that used to sit at the head of every CTA's short
// life; here they run massively parallel once per launch and the CTAs start with a single load.
__global__ void __launch_bounds__(256)
k2_tile_bounds(const K2Layer *__restrict__ layers, int n_layers, const int32_t *__restrict__ idx, long long i_begin,
               int n_tiles, int tile_pts, int2 *__restrict__ bounds) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ly = blockIdx.y;
    if (t >= n_tiles || ly >= n_layers) return;
    const int wm = layers[ly].wm;
    const long long k_lo = i_begin + (long long)t * tile_pts - wm;
    const long long k_hi = i_begin + (long long)t * tile_pts + tile_pts - 1 + wm + 1;
    int a = layers[ly].l_begin, b = layers[ly].l_end;
    while (a < b) { const int m = (a + b) >> 1; if ((long long)idx[m] < k_lo) a = m + 1; else b = m; }
    const int lo = a;
    b = layers[ly].l_end;
    while (a < b) { const int m = (a + b) >> 1; if ((long long)idx[m] < k_hi) a = m + 1; else b = m; }
    bounds[(size_t)ly * n_tiles + t] = make_int2(lo, a);
}

constexpr int KN_THREADS = 256;
constexpr int KN_ROUNDS = 4;                          // points per thread (strided by 256)
constexpr int KN_TILE = KN_THREADS * KN_ROUNDS;       // 1024 points per CTA
constexpr int KN_CHUNK = 2048;                        // staged records per pass

struct KNSmem {
    float4 rec[KN_CHUNK];
    float cc[KN_CHUNK];
    uint64_t bar;
    int lo, hi;
};

__global__ void __launch_bounds__(KN_THREADS, 4) k2_narrow(const K2Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    KNSmem &sm = *reinterpret_cast<KNSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x * KN_TILE;           // shard-local index of the tile's first point
    const K2Layer *L = a.layers + blockIdx.y;         // one grid row per layer of the batch
    const int wm = __ldg(&L->wm);
    const float4 *recA = L->recA;
    const float *recD = L->recD;

    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        int lo, hi;
        if (a.tile_bounds) {
            const int2 bd = __ldg(a.tile_bounds + (size_t)blockIdx.y * gridDim.x + blockIdx.x);
            lo = bd.x; hi = bd.y;
        } else {
            const long long k_lo = a.i_begin + tile0 - wm;
            const long long k_hi = a.i_begin + tile0 + KN_TILE - 1 + wm + 1;
            const int l_end = __ldg(&L->l_end);
            lo = warp_lower_bound(a.idx, __ldg(&L->l_begin), l_end, k_lo);
            hi = warp_lower_bound(a.idx, lo, l_end, k_hi);
        }
        if (lane == 0) { sm.lo = lo & ~3; sm.hi = hi; }
    }
    __syncthreads();
    const int lo = sm.lo, hi = sm.hi;
    const int nch = hi > lo ? (hi - lo + KN_CHUNK - 1) / KN_CHUNK : 0;
    const float wmf = (float)wm;

    float fi[KN_ROUNDS];
    double acc[KN_ROUNDS];
#pragma unroll
    for (int r = 0; r < KN_ROUNDS; ++r) { fi[r] = (float)(tile0 + r * KN_THREADS + tid); acc[r] = 0.0; }

    for (int c = 0; c < nch; ++c) {
        const int first = lo + c * KN_CHUNK;
        const int cnt = min(KN_CHUNK, hi - first);
        if (tid == 0) {
            const uint32_t ce = (uint32_t)((cnt + 3) & ~3);
            mbar_expect_tx(&sm.bar, ce * 20u);
            tma_bulk_g2s(sm.rec, recA + first, ce * 16u, &sm.bar);
            tma_bulk_g2s(sm.cc, recD + first, ce * 4u, &sm.bar);
        }
        mbar_wait(&sm.bar, c & 1);
#pragma unroll
        for (int r = 0; r < KN_ROUNDS; ++r) {
            // first staged line with fidx >= fi - wm  (rec[].x = -fidx, ascending in fidx)
            const float key = fi[r] - wmf, last = fi[r] + wmf;
            int jl = 0, jh = cnt;
            while (jl < jh) {
                const int m = (jl + jh) >> 1;
                if (-sm.rec[m].x < key) jl = m + 1; else jh = m;
            }
            float s32 = 0.f;
            int since = 0;
            for (int j = jl; j < cnt; ++j) {
                const float4 q4 = sm.rec[j];
                if (-q4.x > last) break;
                const float d = fi[r] + q4.x;
                const float d2 = d * d;
                float t = q4.y * rcp_approx(d2 + q4.z);
                if (q4.w != 0.f) t = fmaf(q4.w, ex2_approx(sm.cc[j] * d2), t);
                s32 += t;
                if (++since == K2_FLUSH) { acc[r] += (double)s32; s32 = 0.f; since = 0; }
            }
            acc[r] += (double)s32;
        }
        __syncthreads();                              // the single staging buffer is refilled next pass
    }

    const double inv_scale = __ldg(&L->inv_scale);
    void *out = L->out;
#pragma unroll
    for (int r = 0; r < KN_ROUNDS; ++r) {
        const int i = tile0 + r * KN_THREADS + tid;
        if (i < a.n_chunk) {
            const double v = k2_add_xsc(a, L, i, acc[r] * inv_scale);
            if (a.out_mode == 0) reinterpret_cast<double *>(out)[i] = v;
            else reinterpret_cast<float *>(out)[i] = (float)v;
        }
    }
}

}  // namespace prb

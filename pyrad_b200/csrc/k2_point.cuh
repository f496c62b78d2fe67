// k2_point.cuh -- K2 for NARROW and MID cutoff windows (W-2 <= 511 grid points): thread-per-point gather with
// table-driven line ranges.
//
// Same sum as k2_line_sum (pyradClasses.py:371-400 in gather form).  When the window is not much wider than a
// warp's span, broadcasting every staged line to all lanes wastes most of the work on masks, so here each THREAD
// owns grid points and walks exactly the lines inside its own window -- the window IS the loop range, no masks.
//   * a CTA owns 1024 consecutive points; the sorted lines reaching the tile are staged in shared memory with TMA
//     bulk copies (chunks of 1536 records);
//   * one pass over the staged (sorted) lines fills first[o] = index of the first line at or beyond grid offset o,
//     so a point's window is [first[p], first[p + 2 wm + 1]) -- two table reads instead of two binary searches;
//   * Lorentz terms go through the paired reciprocal  A1/q1 + A2/q2 = (A1 q2 + A2 q1) rcp(q1 q2)  (one MUFU per two
//     pairs); the Gaussian cores are a second, short loop over the lines within the chunk's largest near-zone radius,
//     each line switched by its own radius (so the result does not depend on how the lines were chunked);
//   * FP32 terms, flushed into an FP64 accumulator every 64 lines; fixed ascending line order per point: deterministic.
// Lanes read different records (consecutive lines for consecutive points: conflict-free LDS.128), which bounds this
// mapping by shared-memory bandwidth (16 B per pair); the wide kernel's broadcast mapping wins once windows are wide.
//
// Compact records written by K1 for this kernel:  recA = {-fidx, A, B, G},  recB (as float2) = {C, Dg}.
#pragma once
#include "common.cuh"
#include "k2_line_sum.cuh"
#include "k2_narrow.cuh"

namespace prb {

constexpr int KP_THREADS = 256;
constexpr int KP_ROUNDS = 4;                          // points per thread (strided by 256)
constexpr int KP_TILE = KP_THREADS * KP_ROUNDS;       // 1024 points per CTA
constexpr int KP_CHUNK = 1536;                        // staged records per pass
constexpr int KP_MAX_WM = 511;                        // largest window this kernel's table covers
constexpr int KP_MIN_WM = 16;                         // below this k2_narrow's binary search is cheaper than the table
constexpr int KP_TABLE = KP_TILE + 2 * KP_MAX_WM + 2; // offsets 0 .. TILE + 2 wm + 1

struct KPSmem {
    float4 rec[KP_CHUNK];
    float2 cd[KP_CHUNK];
    int first[KP_TABLE];
    uint64_t bar;
    int lo, hi;
    int dgmax;                                        // largest near-zone radius of the staged chunk (float bits, >= 0)
};

__global__ void __launch_bounds__(KP_THREADS, 4) k2_point(const K2Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    KPSmem &sm = *reinterpret_cast<KPSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x * KP_TILE;           // shard-local index of the tile's first point
    const K2Layer *L = a.layers + blockIdx.y;         // one grid row per layer of the batch
    const int wm = __ldg(&L->wm);
    const float4 *recA = L->recA;
    const float2 *recCD = reinterpret_cast<const float2 *>(L->recB);

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
            const long long k_hi = a.i_begin + tile0 + KP_TILE - 1 + wm + 1;
            const int l_end = __ldg(&L->l_end);
            lo = warp_lower_bound(a.idx, __ldg(&L->l_begin), l_end, k_lo);
            hi = warp_lower_bound(a.idx, lo, l_end, k_hi);
        }
        if (lane == 0) { sm.lo = lo & ~3; sm.hi = hi; }
    }
    __syncthreads();
    const int lo = sm.lo, hi = sm.hi;
    const int nch = hi > lo ? (hi - lo + KP_CHUNK - 1) / KP_CHUNK : 0;
    const int TL = KP_TILE + 2 * wm;                  // table holds offsets 0 .. TL + 1
    const float base_f = (float)(tile0 - wm);         // grid offset o = fidx - base_f

    double acc[KP_ROUNDS];
#pragma unroll
    for (int r = 0; r < KP_ROUNDS; ++r) acc[r] = 0.0;

    for (int c = 0; c < nch; ++c) {
        const int first_line = lo + c * KP_CHUNK;
        const int cnt = min(KP_CHUNK, hi - first_line);
        if (tid == 0) {
            sm.dgmax = 0;
            const uint32_t ce = (uint32_t)((cnt + 3) & ~3);
            mbar_expect_tx(&sm.bar, ce * 24u);
            tma_bulk_g2s(sm.rec, recA + first_line, ce * 16u, &sm.bar);
            tma_bulk_g2s(sm.cd, recCD + first_line, ce * 8u, &sm.bar);
        }
        mbar_wait(&sm.bar, c & 1);
        // first[o] = index of the first staged line whose grid offset is >= o (entry cnt acts as the end sentinel)
        float dgl = 0.f;
        for (int j = tid; j <= cnt; j += KP_THREADS) {
            const float ff = j < cnt ? fminf(-sm.rec[j].x - base_f, (float)(TL + 1)) : (float)(TL + 1);
            const float fp = j > 0 ? fmaxf(fminf(-sm.rec[j - 1].x - base_f, (float)(TL + 1)), -1.f) : -1.f;
            const int oe = (int)ff;
            for (int o = (int)fp + 1; o <= oe; ++o) sm.first[o] = j;
            if (j < cnt) dgl = fmaxf(dgl, sm.cd[j].y);
        }
        dgl = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(dgl, 0.f))));
        if (lane == 0 && dgl > 0.f) atomicMax(&sm.dgmax, __float_as_int(dgl));
        __syncthreads();
        const int dgm = min((int)ceilf(__int_as_float(sm.dgmax)), wm);

#pragma unroll
        for (int r = 0; r < KP_ROUNDS; ++r) {
            const int pl = r * KP_THREADS + tid;                  // tile-local point
            const float fi = (float)(tile0 + pl);
            const int jl = sm.first[pl], jh = sm.first[pl + 2 * wm + 1];
            // [gl, gh): lines within the chunk's largest near-zone radius of this point -- both terms there
            int gl = jl, gh = jl;
            if (sm.dgmax > 0) { gl = sm.first[pl + wm - dgm]; gh = sm.first[pl + wm + dgm + 1]; }
            float s32 = 0.f;
            int since = 0;
            // Lorentz terms outside the near zone, two lines per reciprocal: [jl, gl) then [gh, jh)
#pragma unroll 1
            for (int seg = 0; seg < 2; ++seg) {
                int j = seg ? gh : jl;
                const int je = seg ? jh : gl;
                for (; j + 1 < je; j += 2) {
                    const float4 r1 = sm.rec[j], r2 = sm.rec[j + 1];
                    const float d1 = fi + r1.x, d2 = fi + r2.x;
                    const float q1 = fmaf(d1, d1, r1.z), q2 = fmaf(d2, d2, r2.z);
                    const float num = fmaf(r2.y, q1, r1.y * q2);
                    s32 = fmaf(num, rcp_approx(q1 * q2), s32);
                    if (++since == K2_FLUSH / 2) { acc[r] += (double)s32; s32 = 0.f; since = 0; }
                }
                if (j < je) {
                    const float4 r1 = sm.rec[j];
                    const float d1 = fi + r1.x;
                    s32 = fmaf(r1.y, rcp_approx(fmaf(d1, d1, r1.z)), s32);
                }
            }
            acc[r] += (double)s32;
            s32 = 0.f;
            since = 0;
            for (int g = gl; g < gh; ++g) {                        // near zone: Lorentz + Gaussian, the line's own radius decides
                const float4 r1 = sm.rec[g];
                const float2 cd = sm.cd[g];
                const float d = fi + r1.x;
                const float d2 = d * d;
                const float t = r1.w * ex2_approx(cd.x * d2);
                s32 = fmaf(r1.y, rcp_approx(d2 + r1.z), s32);
                s32 += fabsf(d) <= cd.y ? t : 0.f;
                if (++since == K2_FLUSH) { acc[r] += (double)s32; s32 = 0.f; since = 0; }
            }
            acc[r] += (double)s32;
        }
        __syncthreads();                              // the single staging buffer is refilled next pass
    }

    const double inv_scale = __ldg(&L->inv_scale);
    void *out = L->out;
#pragma unroll
    for (int r = 0; r < KP_ROUNDS; ++r) {
        const int i = tile0 + r * KP_THREADS + tid;
        if (i < a.n_chunk) {
            const double v = k2_add_xsc(a, L, i, acc[r] * inv_scale);
            if (a.out_mode == 0) reinterpret_cast<double *>(out)[i] = v;
            else reinterpret_cast<float *>(out)[i] = (float)v;
        }
    }
}

}  // namespace prb

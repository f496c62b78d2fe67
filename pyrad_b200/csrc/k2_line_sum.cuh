// k2_line_sum.cuh -- K2: Gaussian / Lorentz / pseudo-Voigt line sum into the k(nu) grid.
//
// Replaces the scatter loop Isotope.createCrossSection pyradClasses.py:371-400 and the shapes
// pyradLineshape.py:32-76, in gather form:
//   out[i] = sum over lines l with |i - idx_l| <= W-2 of  A_l/(d^2+B_l) + G_l*exp2(C_l d^2),  d = i - idx_l
// (the window |d| <= W-2 is exactly `range(1, len(rightCurve)-1)` plus the centre sample, with the
// reference's bounds checks becoming "grid point exists").
//
// Design (B200 / sm_100a):
//   * a CTA owns a tile of 256*P consecutive grid points; tiles are handed out by a dynamic counter
//     (work distribution only).  Every thread keeps P points in registers (FP32 partial sums that are
//     flushed into FP64 accumulators every <= 64 lines).  No atomics touch data; the per-point
//     summation order depends only on the tile geometry, so results are deterministic and
//     independent of how the grid is sharded across GPUs (shards are tile aligned).
//   * the wavenumber-sorted lines overlapping the tile's cutoff window are found with a 32-ary
//     warp-ballot search over the sorted index array, then streamed through shared memory in
//     chunks of 512 records with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), double
//     buffered so the copy of chunk c+1 overlaps the math of chunk c.
//   * per chunk every warp classifies the staged (sorted) lines against its own 32*P-point span
//     by counting (six warp reductions): lines whose window only partly covers the span take a
//     predicated path; lines that may need their Gaussian core take a two-term path; all other
//     lines ("far": window covers the span, Gaussian term < 1e-9 of the Lorentz term) take the
//     paired-reciprocal path  A1/q1 + A2/q2 = (A1 q2 + A2 q1) * rcp(q1 q2):  one MUFU per TWO
//     (line, point) pairs instead of the naive one to two per pair.  The SFU pipe (16 lanes/clk/SM)
//     is what bounds the naive formulation, so this is where the kernel gains its speed.
//   * tensor cores are deliberately unused: this is not a dense contraction.
#pragma once
#include "common.cuh"

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

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
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

struct K2Args {
    const float4 *rec4;        // {fidx, A, B, G} per line, fidx relative to the chunk's first point
    const float2 *rec2;        // {C, Dg}
    const int32_t *idx;        // sorted absolute grid index per line
    int l_begin, l_end;        // line range prepared by K1 for this chunk
    long long i_begin;         // absolute index of the chunk's first grid point
    int n_chunk;               // grid points owned
    int wm;                    // W-2 clamped at 0: max |d| that still accumulates
    int n_tiles;
    int variant;               // PRB_K2_GENERAL / PRB_K2_CLASSED
    int out_mode;              // PRB_OUT_F64 / PRB_OUT_F32
    double inv_scale;
    void *out;
    DevState *st;
};

// One ring slot: a chunk of staged line records plus its descriptor.
struct K2Desc {
    int tile0;      // chunk-local index of the first point of the tile this chunk belongs to
    int cnt;        // staged lines to process (padding excluded)
    int flags;      // K2_FIRST | K2_LAST | K2_END
    int pad;
};
constexpr int K2_FIRST = 1, K2_LAST = 2, K2_END = 4;

struct K2Smem {
    float4 r4[K2_STAGES][K2_CHUNK];
    float2 r2[K2_STAGES][K2_CHUNK];
    K2Desc desc[K2_STAGES];
    uint64_t full[K2_STAGES];      // producer -> consumers: TMA bytes landed (+ descriptor written)
    uint64_t empty[K2_STAGES];     // consumers -> producer: all 8 consumer warps are done with the slot
};

template <int P>
__device__ __forceinline__ void flush(float (&a32)[P], double (&a64)[P]) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        a64[p] += (double)a32[p];
        a32[p] = 0.f;
    }
}

// predicated two-term path: window may cover only part of the span
template <int P>
__device__ __forceinline__ void path_general(const float4 *s4, const float2 *s2, int js, int je, float fi0,
                                             float wbf, float we1f, float wmf, float (&a32)[P],
                                             double (&a64)[P]) {
    for (int jb = js; jb < je; jb += K2_FLUSH) {
        const int jend = min(jb + K2_FLUSH, je);
        for (int j = jb; j < jend; ++j) {
            const float4 r = s4[j];
            const float2 g = s2[j];
            const float d0 = fi0 - r.x;
            const bool use_g = (r.w != 0.f) && (r.x + g.y >= wbf) && (r.x - g.y <= we1f);   // warp uniform
            if (use_g) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float e = p ? d0 + (float)(32 * p) : d0;
                    const float e2 = e * e;
                    float t = r.y * rcp_approx(e2 + r.z);
                    t = fmaf(r.w, ex2_approx(g.x * e2), t);
                    a32[p] += (fabsf(e) <= wmf) ? t : 0.f;
                }
            } else {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float e = p ? d0 + (float)(32 * p) : d0;
                    const float t = r.y * rcp_approx(fmaf(e, e, r.z));
                    a32[p] += (fabsf(e) <= wmf) ? t : 0.f;
                }
            }
        }
        flush<P>(a32, a64);
    }
}

// window covers the whole span; Gaussian core may be needed (per-line, warp-uniform test)
template <int P>
__device__ __forceinline__ void path_near(const float4 *s4, const float2 *s2, int js, int je, float fi0,
                                          float wbf, float we1f, float (&a32)[P], double (&a64)[P]) {
    for (int jb = js; jb < je; jb += K2_FLUSH) {
        const int jend = min(jb + K2_FLUSH, je);
        for (int j = jb; j < jend; ++j) {
            const float4 r = s4[j];
            const float2 g = s2[j];
            const float d0 = fi0 - r.x;
            const bool use_g = (r.w != 0.f) && (r.x + g.y >= wbf) && (r.x - g.y <= we1f);
            if (use_g) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float e = p ? d0 + (float)(32 * p) : d0;
                    const float e2 = e * e;
                    a32[p] = fmaf(r.y, rcp_approx(e2 + r.z), a32[p]);
                    a32[p] = fmaf(r.w, ex2_approx(g.x * e2), a32[p]);
                }
            } else {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float e = p ? d0 + (float)(32 * p) : d0;
                    a32[p] = fmaf(r.y, rcp_approx(fmaf(e, e, r.z)), a32[p]);
                }
            }
        }
        flush<P>(a32, a64);
    }
}

// far path: window covers the span, Gaussian term negligible.  Two lines share one reciprocal.
template <int P>
__device__ __forceinline__ void path_far(const float4 *s4, int js, int je, float fi0, float (&a32)[P],
                                         double (&a64)[P]) {
    for (int jb = js; jb < je; jb += K2_FLUSH) {
        const int jend = min(jb + K2_FLUSH, je);
        int j = jb;
        for (; j + 1 < jend; j += 2) {
            const float4 r1 = s4[j];
            const float4 r2 = s4[j + 1];
            const float d1 = fi0 - r1.x;
            const float d2 = fi0 - r2.x;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float e1 = p ? d1 + (float)(32 * p) : d1;
                const float e2 = p ? d2 + (float)(32 * p) : d2;
                const float q1 = fmaf(e1, e1, r1.z);
                const float q2 = fmaf(e2, e2, r2.z);
                const float num = fmaf(r2.y, q1, r1.y * q2);
                a32[p] = fmaf(num, rcp_approx(q1 * q2), a32[p]);
            }
        }
        if (j < jend) {
            const float4 r = s4[j];
            const float d0 = fi0 - r.x;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float e = p ? d0 + (float)(32 * p) : d0;
                a32[p] = fmaf(r.y, rcp_approx(fmaf(e, e, r.z)), a32[p]);
            }
        }
        flush<P>(a32, a64);
    }
}

// Warp-specialised persistent kernel: warps 0..7 consume (math), warp 8 produces (tile scheduling,
// line-range search, TMA issue).  Slots of the ring are handed over with mbarriers only -- there is
// no CTA-wide barrier in the steady state, so a warp that is in its slow near-zone does not hold up
// the other seven (each warp's near-zone sits at a different place in the line stream).
template <int P>
__global__ void __launch_bounds__(K2_THREADS, 2) k2_line_sum(const K2Args a) {
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
            mbar_wait(&sm.empty[stage], ((it / K2_STAGES) & 1) ^ 1);   // fresh barrier: passes at once
        };
        while (true) {
            int tile = 0;
            if (lane == 0) tile = (int)atomicAdd(&a.st->tile_counter, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= a.n_tiles) break;
            const int tile0 = tile * TILE;
            const long long k_lo = a.i_begin + tile0 - a.wm;
            const long long k_hi = a.i_begin + tile0 + TILE - 1 + a.wm + 1;
            int lo = warp_lower_bound(a.idx, a.l_begin, a.l_end, k_lo);
            const int hi = warp_lower_bound(a.idx, lo, a.l_end, k_hi);
            lo &= ~1;                                  // 16-byte alignment of the float2 stream
            const int nch = hi > lo ? (hi - lo + K2_CHUNK - 1) / K2_CHUNK : 1;   // empty tile: one empty chunk
            for (int c = 0; c < nch; ++c, ++it) {
                uint32_t stage;
                slot_acquire(stage);
                if (lane == 0) {
                    const int first = lo + c * K2_CHUNK;
                    const int cnt = max(min(K2_CHUNK, hi - first), 0);
                    sm.desc[stage] = K2Desc{tile0, cnt, (c == 0 ? K2_FIRST : 0) | (c == nch - 1 ? K2_LAST : 0), 0};
                    if (cnt > 0) {
                        const uint32_t ce = (uint32_t)((cnt + 1) & ~1);   // padding records exist past l_end
                        mbar_expect_tx(&sm.full[stage], ce * 24u);
                        tma_bulk_g2s(sm.r4[stage], a.rec4 + first, ce * 16u, &sm.full[stage]);
                        tma_bulk_g2s(sm.r2[stage], a.rec2 + first, ce * 8u, &sm.full[stage]);
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
            sm.desc[stage] = K2Desc{0, 0, K2_END, 0};
            mbar_arrive(&sm.full[stage]);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const float wmf = (float)a.wm;
    float a32[P];
    double a64[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { a32[p] = 0.f; a64[p] = 0.0; }
    int wb = 0;
    float wbf = 0.f, we1f = 0.f, fi0 = 0.f;

    for (uint32_t it = 0;; ++it) {
        const uint32_t stage = it % K2_STAGES;
        mbar_wait(&sm.full[stage], (it / K2_STAGES) & 1);
        const K2Desc d = sm.desc[stage];
        if (d.flags & K2_END) break;
        if (d.flags & K2_FIRST) {
#pragma unroll
            for (int p = 0; p < P; ++p) { a32[p] = 0.f; a64[p] = 0.0; }
            wb = d.tile0 + warp * SPAN;                // first point of this warp's span
            wbf = (float)wb;
            we1f = (float)(wb + SPAN - 1);
            fi0 = (float)(wb + lane);
        }
        const int cnt = d.cnt;
        const float4 *s4 = sm.r4[stage];
        const float2 *s2 = sm.r2[stage];

        if (a.variant == 0) {
            path_general<P>(s4, s2, 0, cnt, fi0, wbf, we1f, wmf, a32, a64);
        } else if (cnt > 0) {
            // near-zone radius of THIS staged chunk (max over its lines): a function of the tile
            // geometry only, so the classes -- and the FP32 rounding -- do not depend on the sharding.
            float dgl = 0.f;
            for (int j = lane; j < cnt; j += 32) dgl = fmaxf(dgl, s2[j].y);
            const float dgmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(dgl)));
            // class boundaries of the sorted staged lines relative to this warp's span, by counting
            const float t0 = wbf - wmf;                       // idx <  t0 : window ends before the span
            const float t5 = we1f + 1.f + wmf;                // idx >= t5 : window starts after the span
            const float fl = we1f - wmf, fr1 = wbf + wmf + 1.f;   // full cover: fl <= idx < fr1
            float t1, t2, t3, t4;
            if (fl >= fr1) { t1 = t2 = t3 = t4 = t5; }       // window narrower than the span
            else {
                t1 = fl;
                t4 = fr1;
                t2 = fminf(fmaxf(wbf - dgmax, fl), fr1);
                t3 = fminf(fmaxf(we1f + 1.f + dgmax, t2), fr1);
            }
            int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
            for (int j = lane; j < cnt; j += 32) {
                const float f = s4[j].x;
                c0 += f < t0; c1 += f < t1; c2 += f < t2; c3 += f < t3; c4 += f < t4; c5 += f < t5;
            }
            const int b0 = __reduce_add_sync(0xffffffffu, c0);
            const int b1 = __reduce_add_sync(0xffffffffu, c1);
            const int b2 = __reduce_add_sync(0xffffffffu, c2);
            const int b3 = __reduce_add_sync(0xffffffffu, c3);
            const int b4 = __reduce_add_sync(0xffffffffu, c4);
            const int b5 = __reduce_add_sync(0xffffffffu, c5);
            path_general<P>(s4, s2, b0, b1, fi0, wbf, we1f, wmf, a32, a64);
            path_far<P>(s4, b1, b2, fi0, a32, a64);
            path_near<P>(s4, s2, b2, b3, fi0, wbf, we1f, a32, a64);
            path_far<P>(s4, b3, b4, fi0, a32, a64);
            path_general<P>(s4, s2, b4, b5, fi0, wbf, we1f, wmf, a32, a64);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);  // slot may be refilled

        if (d.flags & K2_LAST) {
            // epilogue: undo the power-of-two scale exactly and store (coalesced 32-point rows)
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int i = wb + 32 * p + lane;
                if (i < a.n_chunk) {
                    const double v = a64[p] * a.inv_scale;
                    if (a.out_mode == 0) reinterpret_cast<double *>(a.out)[i] = v;
                    else reinterpret_cast<float *>(a.out)[i] = (float)v;
                }
            }
        }
    }
}

}  // namespace prb

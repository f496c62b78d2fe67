// common.cuh -- shared definitions for libpyrad_b200 (sm_100a only; no other backend exists).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace prb {

// Physical constants exactly as the reference spells them
// (pyradIntensity.py:3-13, pyradLineshape.py:14-19, pyradPlanck.py:4-9, pyradClasses.py:15-23).
constexpr double kBoltz = 1.38064852E-23;
constexpr double cLight = 299792458.0;
constexpr double hPlanck = 6.62607004e-34;
constexpr double kPi = 3.141592653589793;
constexpr double kT0 = 296.0;
constexpr double kP0 = 1013.25;
constexpr double kAvogadro = 6.022140857E23;

constexpr int REGIME_GAUSS = 0, REGIME_LORENTZ = 1, REGIME_VOIGT = 2;

// Per-group (isotopologue) scalars for one layer, built on the host in FP64.
struct GroupParams {
    double conc;     // molecule mole fraction q            (pyradClasses.py:258)
    double dopp;     // sqrt(2 k T / m / c^2)               (pyradClasses.py:263)
    double qratio;   // Q(296) / Q(T)                       (pyradIntensity.py:31)
    double weight;   // output weight folded into the coefficients
};

// Small device-resident state block, cleared before every prepass.
struct DevState {
    unsigned int reserved0;
    unsigned int flags;         // bit0: coefficient overflow, bit1: non-finite coefficient
    unsigned int tile_counter;  // dynamic tile scheduler of K2 (work distribution only, never data)
    unsigned int pad;
};

constexpr unsigned int FLAG_OVERFLOW = 1u, FLAG_NONFINITE = 2u;

// K2 geometry
#ifndef PRB_K2_CONSUMERS
#define PRB_K2_CONSUMERS 8
#endif
#ifndef PRB_K2_ACC_SMEM
#define PRB_K2_ACC_SMEM 0
#endif
constexpr int K2_CONSUMERS = PRB_K2_CONSUMERS;        // math warps per CTA
constexpr int K2_THREADS = 32 * (K2_CONSUMERS + 1);   // + one TMA producer warp
#ifndef PRB_K2_CHUNK
#define PRB_K2_CHUNK 768
#endif
#ifndef PRB_K2_STAGES
#define PRB_K2_STAGES 3
#endif
#ifndef PRB_K2_MIN_CTAS
#define PRB_K2_MIN_CTAS 2
#endif
constexpr int K2_CHUNK = PRB_K2_CHUNK;   // lines staged per ring slot (three TMA bulk copies)
constexpr int K2_STAGES = PRB_K2_STAGES; // ring depth
constexpr int K2_MIN_CTAS = PRB_K2_MIN_CTAS;   // resident CTAs per SM the register budget is sized for
constexpr int K2_FLUSH = 64;             // lines accumulated in FP32 before flushing into FP64
constexpr float K2_SENTINEL = 3.0e38f;   // fidx of padding records: outside every window

// ---- mbarrier / TMA bulk-copy helpers shared by the kernels that stage through shared memory -------------------
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
// Producer-side wait: the producer runs ahead of the math warps and spends most of its life here, so it asks the
// hardware to park the warp (suspend-time hint) instead of burning issue slots the math warps could use.
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONEP_%=;\n"
        "bra WAITP_%=;\n"
        "DONEP_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(4000u)
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

}  // namespace prb

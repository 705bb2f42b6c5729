// One-quartet-per-thread FP64 Rys J/K kernel, specialised per (li,lj,lk,ll).
//
// Path row a11 of SURVEY.md section 8: replaces the reference's NVRTC-generated
// rys_1q1t_vjk (jqc/backend/jk/1q1t.cu:45-644) with an ahead-of-time sm_100a kernel:
//   * persistent grid-stride loop over a device-resident quartet list whose length is read
//     from device memory (no host round trip between task generation and evaluation);
//   * primitive counts are run-time loop bounds (one instantiation per angular class);
//   * Rys roots from the interval table without small-x / erf branches;
//   * blocks of up to JQC_SMALL_N integrals are fully unrolled into registers (_small), with the
//     J_ij contributions of a warp that shares (i,j) combined by shuffles before one reduction;
//     larger blocks run on jk_warp.cuh, and the same body rolled with thread-local scratch
//     (_large) is the any-l fallback (used for classes with g shells in the bra).
#pragma once
#include "jqc_common.cuh"

namespace jqc {

struct JKArgs {
    int nao;
    int n_dm;
    int npi, npj, npk, npl;
    const double* __restrict__ basis;     // nbas x 12 packed shell records
    const double* __restrict__ dm;        // n_dm x nao x nao (kernel-side cartesian basis)
    double* __restrict__ vj;              // n_dm x nao x nao or nullptr
    double* __restrict__ vk;
    double omega;                         // 0: Coulomb, > 0: long-range erf
    const ushort4* __restrict__ quartets;
    const unsigned* __restrict__ ntasks;  // device counter written by the task generator
};

template <int LI, int LJ, int LK, int LL>
struct QuartetShape {
    static constexpr int NFI = nf_of(LI), NFJ = nf_of(LJ), NFK = nf_of(LK), NFL = nf_of(LL);
    static constexpr int N = NFI * NFJ * NFK * NFL;
    static constexpr int LIJ = LI + LJ, LKL = LK + LL;
    static constexpr int NROOTS = (LIJ + LKL) / 2 + 1;
    static constexpr int DI = 1, DJ = LI + 1, DK = DJ * (LJ + 1), DL = DK * (LK + 1);
    static constexpr int GSIZE = DL * (LL + 1);
};

// Blocks up to JQC_SMALL_N integrals run on the register kernel.  81 fit without spilling; up to 108
// the kernel spills ~1 KB per thread yet still beats the multi-lane kernel by ~2x (measured on B200,
// profiles/r1/class_times_09_*.csv), hence the higher threshold.  The 4x4-tile variant stays at 81.
#ifndef JQC_SMALL_N_VALUE
#define JQC_SMALL_N_VALUE 108
#endif
constexpr int JQC_SMALL_N = JQC_SMALL_N_VALUE;
constexpr int JQC_TILE16_N = 81;

// register-resident variant: blocks of <= 27 integrals fit 128 registers (two CTAs per SM)
#define JQC_MINB(a, b, c, d) (QuartetShape<a, b, c, d>::N <= 27 ? 2 : 1)
#define JQC_WARP_COMBINE true
#define JQC_NAME(x) x##_small
#define JQC_UNROLL _Pragma("unroll")
#include "jk_1q1t_body.inc"
#undef JQC_NAME
#undef JQC_UNROLL
#undef JQC_MINB
#undef JQC_WARP_COMBINE

#define JQC_MINB(a, b, c, d) 1
#define JQC_WARP_COMBINE false
#define JQC_NAME(x) x##_large
#define JQC_UNROLL _Pragma("unroll 1")
#include "jk_1q1t_body.inc"
#undef JQC_NAME
#undef JQC_UNROLL
#undef JQC_MINB
#undef JQC_WARP_COMBINE

}  // namespace jqc

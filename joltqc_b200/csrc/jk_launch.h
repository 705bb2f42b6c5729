// Launcher table shared by engine.cu and the per-class translation units (jk_inst.cu).
#pragma once
#include <cuda_runtime.h>

namespace jqc {
struct JKArgs;
struct BrickArgs;
#define JQC_DECL(I, J)                                                                                          \
    cudaError_t jk_launch_##I##_##J(int lk, int ll, int variant, const JKArgs& a, int nsm, cudaStream_t st);    \
    cudaError_t jk_brick_launch_##I##_##J(int lk, int ll, int variant, const BrickArgs& a, int nsm, cudaStream_t st);
JQC_DECL(0, 0)
JQC_DECL(1, 0) JQC_DECL(1, 1)
JQC_DECL(2, 0) JQC_DECL(2, 1) JQC_DECL(2, 2)
JQC_DECL(3, 0) JQC_DECL(3, 1) JQC_DECL(3, 2) JQC_DECL(3, 3)
JQC_DECL(4, 0) JQC_DECL(4, 1) JQC_DECL(4, 2) JQC_DECL(4, 3) JQC_DECL(4, 4)
#undef JQC_DECL

// Classes of <= 81 integrals can optionally be fed (i, j, k-tile, l-tile, mask) records and run on
// jk_tile16.cuh (engine option JQC_SMALL_TILES=1); by default every class consumes flat ushort4
// quartet lists (register kernel up to 108 integrals, multi-lane kernel above).
inline bool jk_uses_tiles(int li, int lj, int lk, int ll)
{
    auto nf = [](int l) { return (l + 1) * (l + 2) / 2; };
    return nf(li) * nf(lj) * nf(lk) * nf(ll) <= 81;
}

// variant: bit0 = J, bit1 = K.  Requires li >= lj, li >= lk, lk >= ll (the order in which the
// group-quartet loop of the reference enumerates classes, jqc/pyscf/jk.py:145-151).
inline cudaError_t jk_launch(int li, int lj, int lk, int ll, int variant, const JKArgs& a, int nsm, cudaStream_t st)
{
#define JQC_CALL(I, J) if (li == I && lj == J) return jk_launch_##I##_##J(lk, ll, variant, a, nsm, st);
    JQC_CALL(0, 0)
    JQC_CALL(1, 0) JQC_CALL(1, 1)
    JQC_CALL(2, 0) JQC_CALL(2, 1) JQC_CALL(2, 2)
    JQC_CALL(3, 0) JQC_CALL(3, 1) JQC_CALL(3, 2) JQC_CALL(3, 3)
    JQC_CALL(4, 0) JQC_CALL(4, 1) JQC_CALL(4, 2) JQC_CALL(4, 3) JQC_CALL(4, 4)
#undef JQC_CALL
    return cudaErrorInvalidValue;
}

// Classes that run on the one-lane brick kernel: brick_shape(li, lj, lk, ll).fits (jk_brick.cuh);
// the brick-scheduled multi-lane kernel (jk_bwarp.cuh, variant bit 3) takes the larger classes up to f shells.
// Which of those classes actually use it is a measured table (jk_class_select.h, tools/gen_class_select.py);
// mode 2 (JQC_BWARP=2) forces it for every supported class.
#include "jk_class_select.h"
inline bool jk_bwarp_supported(int li, int lj, int lk, int ll, int mode = 1)
{
    auto nf = [](int l) { return (l + 1) * (l + 2) / 2; };
    if (!(li <= 3 && nf(li) * nf(lj) * nf(lk) * nf(ll) > JQC_SMALL_N_VALUE)) return false;
    return mode >= 2 || jk_class_prefers_bwarp(li, lj, lk, ll);
}

inline cudaError_t jk_brick_launch(int li, int lj, int lk, int ll, int variant, const BrickArgs& a, int nsm, cudaStream_t st)
{
#define JQC_CALL(I, J) if (li == I && lj == J) return jk_brick_launch_##I##_##J(lk, ll, variant, a, nsm, st);
    JQC_CALL(0, 0)
    JQC_CALL(1, 0) JQC_CALL(1, 1)
    JQC_CALL(2, 0) JQC_CALL(2, 1) JQC_CALL(2, 2)
    JQC_CALL(3, 0) JQC_CALL(3, 1) JQC_CALL(3, 2) JQC_CALL(3, 3)
    JQC_CALL(4, 0) JQC_CALL(4, 1) JQC_CALL(4, 2) JQC_CALL(4, 3) JQC_CALL(4, 4)
#undef JQC_CALL
    return cudaErrorInvalidValue;
}
}  // namespace jqc

// One translation unit per bra class (JQC_LI >= JQC_LJ), compiled with -DJQC_LI=.. -DJQC_LJ=..
// Instantiates the J/K kernels for every ket class lk <= li, ll <= lk and the three
// (do_j, do_k) variants, and exports one launcher.  (The reference JIT-compiles the same
// specialisations at run time through NVRTC: jqc/backend/jk.py:56-115.)
#include <atomic>

#include "jk_tile16.cuh"
#include "jk_warp.cuh"
#include "jk_bwarp.cuh"
#include "jk_launch.h"

namespace jqc {

// Occupancy and the dynamic shared-memory opt-in are per device (function attributes do not
// carry over to another GPU of the same process): cached per device ordinal.  A racing first
// call repeats idempotent work.
constexpr int JQC_MAX_DEVICES = 64;
template <class K>
static cudaError_t blocks_per_sm_cached(std::atomic<int>* cache, K kern, int threads, size_t smem, int* out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= JQC_MAX_DEVICES) return cudaErrorInvalidDevice;
    int nb = cache[dev].load(std::memory_order_acquire);
    if (nb == 0) {
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem);
        if (e != cudaSuccess) return e;
        nb = nb > 0 ? nb : 1;
        cache[dev].store(nb, std::memory_order_release);
    }
    *out = nb;
    return cudaSuccess;
}

// Multi-lane kernel: warps per block chosen so that several blocks fit the 227 KB of an SM.  (A
// shared-memory copy of the Rys table with one big block per SM was measured and lost ~6 % here:
// the roots are < 5 % of this kernel's stall samples, the table costs occupancy; profiles/r2.)
template <int LK, int LL>
struct WarpCfg {
    using P = WarpPlan<JQC_LI, JQC_LJ, LK, LL>;
    static constexpr size_t WARP_BYTES = (size_t)P::QPW * P::PER_GROUP * sizeof(double);
    static constexpr int nwarps()
    {
        // blocks of <= 4 warps, small enough that shared memory allows >= 16 warps per SM when the
        // per-warp footprint permits
        int n = (int)(56 * 1024 / WARP_BYTES);
        return n > 4 ? 4 : (n < 1 ? 1 : n);
    }
    static constexpr int NWARPS = nwarps();
    static constexpr size_t SMEM = WARP_BYTES * NWARPS;
    static constexpr bool FITS = SMEM <= 200 * 1024;
};

template <int LK, int LL, bool DO_J, bool DO_K>
static cudaError_t launch_warp(const JKArgs& a, int nsm, cudaStream_t st)
{
    using C = WarpCfg<LK, LL>;
    auto kern = jk_warp_kernel<JQC_LI, JQC_LJ, LK, LL, DO_J, DO_K, C::NWARPS>;
    static std::atomic<int> cache[JQC_MAX_DEVICES];
    int blocks_per_sm = 1;
    cudaError_t e = blocks_per_sm_cached(cache, kern, C::NWARPS * 32, C::SMEM, &blocks_per_sm);
    if (e != cudaSuccess) return e;
    kern<<<nsm * blocks_per_sm, C::NWARPS * 32, C::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <class R, int LK, int LL, bool DO_J, bool DO_K>
static cudaError_t launch_bwarp(const BrickArgs& a_in, int nsm, cudaStream_t st)
{
    using S = QuartetShape<JQC_LI, JQC_LJ, LK, LL>;
    using P = BWarpPlan<R, JQC_LI, JQC_LJ, LK, LL>;
    if constexpr (S::N > JQC_SMALL_N && P::FITS && JQC_LI <= 3) {
        auto kern = jk_bwarp_kernel<R, JQC_LI, JQC_LJ, LK, LL, DO_J, DO_K, P::NWARPS>;
        static std::atomic<int> cache[JQC_MAX_DEVICES];
        int blocks_per_sm = 1;
        cudaError_t e = blocks_per_sm_cached(cache, kern, P::NWARPS * 32, P::SMEM, &blocks_per_sm);
        if (e != cudaSuccess) return e;
        BrickArgs b = a_in;
        brick_decompose(b, P::QPW, nsm);
        const long long ntask = ((long long)b.n_blk * b.n_ichunk * b.jsplit + b.world - 1) / b.world;
        long long blocks = (ntask + P::NWARPS - 1) / P::NWARPS;
        if (blocks > (long long)nsm * blocks_per_sm) blocks = (long long)nsm * blocks_per_sm;
        if (blocks < 1) blocks = 1;
        kern<<<(unsigned)blocks, P::NWARPS * 32, P::SMEM, st>>>(b);
        return cudaGetLastError();
    } else {
        return cudaErrorInvalidValue;
    }
}

template <class R, int LK, int LL, bool DO_J, bool DO_K>
static cudaError_t launch_brick(const BrickArgs& a_in, int nsm, cudaStream_t st)
{
    using P = BrickPlan<R, JQC_LI, JQC_LJ, LK, LL>;
    if constexpr (P::FITS) {
        auto kern = jk_brick_kernel<R, JQC_LI, JQC_LJ, LK, LL, DO_J, DO_K>;
        static std::atomic<int> cache[JQC_MAX_DEVICES];
        int blocks_per_sm = 1;
        cudaError_t e = blocks_per_sm_cached(cache, kern, P::NWARPS * 32, P::SMEM, &blocks_per_sm);
        if (e != cudaSuccess) return e;
        BrickArgs b = a_in;
        brick_decompose(b, 32, nsm);
        const BrickArgs& a = b;
        // no more warps than tasks
        const long long ntask = ((long long)a.n_blk * a.n_ichunk * a.jsplit + a.world - 1) / a.world;
        long long blocks = (ntask + P::NWARPS - 1) / P::NWARPS;
        if (blocks > (long long)nsm * blocks_per_sm) blocks = (long long)nsm * blocks_per_sm;
        if (blocks < 1) blocks = 1;
        kern<<<(unsigned)blocks, P::NWARPS * 32, P::SMEM, st>>>(a);
        return cudaGetLastError();
    } else {
        return cudaErrorInvalidValue;
    }
}

template <int LK, int LL, bool DO_J, bool DO_K, bool TILES>
static cudaError_t launch_one(const JKArgs& a, int nsm, cudaStream_t st)
{
    using S = QuartetShape<JQC_LI, JQC_LJ, LK, LL>;
    if constexpr (S::N > JQC_SMALL_N && WarpCfg<LK, LL>::FITS && JQC_LI <= 3) {
        return launch_warp<LK, LL, DO_J, DO_K>(a, nsm, st);
    } else {
        // (else-branch so that the rolled fallback is only instantiated for the classes that use it)
        constexpr bool SMALL = S::N <= JQC_SMALL_N;
        constexpr int NT = SMALL ? 256 : 128;
        static_assert(JQC_TILE16_N == 81, "keep jk_uses_tiles() in jk_launch.h in sync");
        void (*kern)(const JKArgs);
        if constexpr (SMALL) kern = TILES ? jk_tile16_kernel<JQC_LI, JQC_LJ, LK, LL, DO_J, DO_K, NT>
                                          : jk_1q1t_kernel_small<JQC_LI, JQC_LJ, LK, LL, DO_J, DO_K, NT>;
        else kern = jk_1q1t_kernel_large<JQC_LI, JQC_LJ, LK, LL, DO_J, DO_K, NT>;
        static std::atomic<int> cache[JQC_MAX_DEVICES];
        int blocks_per_sm = 1;
        cudaError_t e = blocks_per_sm_cached(cache, kern, NT, 0, &blocks_per_sm);
        if (e != cudaSuccess) return e;
        kern<<<nsm * blocks_per_sm, NT, 0, st>>>(a);
        return cudaGetLastError();
    }
}

template <int LK, int LL>
static cudaError_t launch_variant(int variant, const JKArgs& a, int nsm, cudaStream_t st)
{
    // bit 2 of the variant selects the tile-record kernel for the small classes
    constexpr bool SM = QuartetShape<JQC_LI, JQC_LJ, LK, LL>::N <= JQC_TILE16_N;
    switch (variant) {
        case 3: return launch_one<LK, LL, true, true, false>(a, nsm, st);
        case 1: return launch_one<LK, LL, true, false, false>(a, nsm, st);
        case 2: return launch_one<LK, LL, false, true, false>(a, nsm, st);
        case 7: if constexpr (SM) return launch_one<LK, LL, true, true, true>(a, nsm, st); break;
        case 5: if constexpr (SM) return launch_one<LK, LL, true, false, true>(a, nsm, st); break;
        case 6: if constexpr (SM) return launch_one<LK, LL, false, true, true>(a, nsm, st); break;
    }
    return cudaErrorInvalidValue;
}

#define JQC_CAT_(a, b, c) a##b##_##c
#define JQC_CAT(a, b, c) JQC_CAT_(a, b, c)

// (templated on a dummy so that the discarded `if constexpr` branches are not instantiated)
template <int Z>
static cudaError_t dispatch(int lk, int ll, int variant, const JKArgs& a, int nsm, cudaStream_t st)
{
#define CASE(K, L)                                                       \
    if constexpr (K + Z <= JQC_LI && L <= K) {                           \
        if (lk == K && ll == L) return launch_variant<K, L>(variant, a, nsm, st); \
    }
    CASE(0, 0)
    CASE(1, 0) CASE(1, 1)
    CASE(2, 0) CASE(2, 1) CASE(2, 2)
    CASE(3, 0) CASE(3, 1) CASE(3, 2) CASE(3, 3)
    CASE(4, 0) CASE(4, 1) CASE(4, 2) CASE(4, 3) CASE(4, 4)
#undef CASE
    return cudaErrorInvalidValue;
}

cudaError_t JQC_CAT(jk_launch_, JQC_LI, JQC_LJ)(int lk, int ll, int variant, const JKArgs& a, int nsm, cudaStream_t st)
{
    return dispatch<0>(lk, ll, variant, a, nsm, st);
}

template <int LK, int LL>
static cudaError_t brick_variant(int variant, const BrickArgs& a, int nsm, cudaStream_t st)
{
    switch (variant) {
        case 3: return launch_brick<double, LK, LL, true, true>(a, nsm, st);
        case 1: return launch_brick<double, LK, LL, true, false>(a, nsm, st);
        case 2: return launch_brick<double, LK, LL, false, true>(a, nsm, st);
        // FP32-band variant (mixed precision): same classes, float integrals
        case 19: return launch_brick<float, LK, LL, true, true>(a, nsm, st);
        case 17: return launch_brick<float, LK, LL, true, false>(a, nsm, st);
        case 18: return launch_brick<float, LK, LL, false, true>(a, nsm, st);
        case 11: return launch_bwarp<double, LK, LL, true, true>(a, nsm, st);
        case 9: return launch_bwarp<double, LK, LL, true, false>(a, nsm, st);
        case 10: return launch_bwarp<double, LK, LL, false, true>(a, nsm, st);
        case 27: return launch_bwarp<float, LK, LL, true, true>(a, nsm, st);
        case 25: return launch_bwarp<float, LK, LL, true, false>(a, nsm, st);
        case 26: return launch_bwarp<float, LK, LL, false, true>(a, nsm, st);
    }
    return cudaErrorInvalidValue;
}

template <int Z>
static cudaError_t brick_dispatch(int lk, int ll, int variant, const BrickArgs& a, int nsm, cudaStream_t st)
{
#define CASE(K, L)                                                       \
    if constexpr (K + Z <= JQC_LI && L <= K) {                           \
        if (lk == K && ll == L) return brick_variant<K, L>(variant, a, nsm, st); \
    }
    CASE(0, 0)
    CASE(1, 0) CASE(1, 1)
    CASE(2, 0) CASE(2, 1) CASE(2, 2)
    CASE(3, 0) CASE(3, 1) CASE(3, 2) CASE(3, 3)
    CASE(4, 0) CASE(4, 1) CASE(4, 2) CASE(4, 3) CASE(4, 4)
#undef CASE
    return cudaErrorInvalidValue;
}

cudaError_t JQC_CAT(jk_brick_launch_, JQC_LI, JQC_LJ)(int lk, int ll, int variant, const BrickArgs& a, int nsm, cudaStream_t st)
{
    return brick_dispatch<0>(lk, ll, variant, a, nsm, st);
}

}  // namespace jqc

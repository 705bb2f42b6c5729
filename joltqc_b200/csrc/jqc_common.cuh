// Shared device-side definitions for the joltqc_b200 J/K engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rys_tables.cuh"

namespace jqc {

constexpr int LMAX = 4;
constexpr int NPRIM_MAX = 3;
constexpr int BASIS_STRIDE = 12;   // doubles per packed shell record (jqc/constants.py:27)
constexpr int TILE = 4;
constexpr double PI_FAC = 34.98683665524972497;   // 2*pi^2.5
constexpr double SQRTPIE4 = .8862269254527580136;

__host__ __device__ constexpr int nf_of(int l) { return (l + 1) * (l + 2) / 2; }

// Cartesian exponents of component c of an l-shell; order lx descending, then ly descending
// (the order the reference's index tables use, jqc/backend/util.py:21-36).
__host__ __device__ constexpr int cart_x(int l, int c)
{
    int n = 0;
    for (int x = l; x >= 0; x--)
        for (int y = l - x; y >= 0; y--) {
            if (n == c) return x;
            n++;
        }
    return 0;
}
__host__ __device__ constexpr int cart_y(int l, int c)
{
    int n = 0;
    for (int x = l; x >= 0; x--)
        for (int y = l - x; y >= 0; y--) {
            if (n == c) return y;
            n++;
        }
    return 0;
}
__host__ __device__ constexpr int cart_z(int l, int c) { return l - cart_x(l, c) - cart_y(l, c); }

__device__ constexpr signed char CART_X[5][15] = {
    {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0},
    {1,0,0,0,0,0,0,0,0,0,0,0,0,0,0},
    {2,1,1,0,0,0,0,0,0,0,0,0,0,0,0},
    {3,2,2,1,1,1,0,0,0,0,0,0,0,0,0},
    {4,3,3,2,2,2,1,1,1,1,0,0,0,0,0}};
__device__ constexpr signed char CART_Y[5][15] = {
    {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0},
    {0,1,0,0,0,0,0,0,0,0,0,0,0,0,0},
    {0,1,0,2,1,0,0,0,0,0,0,0,0,0,0},
    {0,1,0,2,1,0,3,2,1,0,0,0,0,0,0},
    {0,1,0,2,1,0,3,2,1,0,4,3,2,1,0}};
__device__ constexpr signed char CART_Z[5][15] = {
    {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0},
    {0,0,1,0,0,0,0,0,0,0,0,0,0,0,0},
    {0,0,1,0,1,2,0,0,0,0,0,0,0,0,0},
    {0,0,1,0,1,2,0,1,2,3,0,0,0,0,0},
    {0,0,1,0,1,2,0,1,2,3,0,1,2,3,4}};

// ------------------------------------------------------------------ ordered float <-> int
__device__ __forceinline__ int float_to_ordered(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// Static multi-GPU partition of a cost-sorted task list: round r of `world` consecutive entries goes to ranks
// 0..world-1 when r is even and world-1..0 when r is odd (boustrophedon), so that a list sorted by descending
// Schwarz bound (= descending cost) does not hand rank 0 the heaviest entry of every round.
__host__ __device__ __forceinline__ unsigned shard_entry(unsigned round, int rank, int world)
{
    return round * (unsigned)world + (unsigned)((round & 1u) ? world - 1 - rank : rank);
}

// Packed record: [x, y, z, ao_loc, c0, e0, c1, e1, c2, e2, 0, 0]
struct ShellHead {
    double x, y, z, ao;
};

// ---------------------------------------------------------------------------------------
// Rys roots (t^2) and weights for NROOTS points at argument x (already theta-scaled).
// Two regimes only: the interval table (the first interval covers x -> 0, so no small-x branch)
// and the Hermite asymptote for x >= 35 + 5 n.  Same fit and accuracy class as the reference's
// rys_roots (jqc/backend/rys/rys_roots.cu:29-160); the range separation scaling for omega > 0
// follows :42-47.  The reference evaluates each degree-13 series with a Clenshaw recurrence
// (rys_roots.cu:110-140): 26 dependent FP64 operations per value.  Here the fit is stored in the
// power basis (tools/gen_rys_tables.py) and evaluated with Estrin's scheme: 13 FMAs of depth 4 on
// u, u^2, u^4, u^8 shared by all 2 n series of a quartet, so the evaluation is throughput- and not
// latency-bound at the 2-4 resident warps per scheduler these kernels run with.
// rw[2i] = root, rw[2i+1] = weight.
// 1/sqrt in one MUFU + Newton step; the kernels derive both 1/p and 1/sqrt(p) from it (a division plus
// a square root cost about three times as many instructions)
__device__ __forceinline__ double jrsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float jrsqrt(float x) { return rsqrtf(x); }

struct RysPowers {
    double u, u2, u4, u8;
    int it;
};
__device__ __forceinline__ RysPowers rys_powers(double x)
{
    RysPowers p;
    p.it = (int)(x * 0.4);
    p.u = fma(x - p.it * 2.5, 0.8, -1.0);
    p.u2 = p.u * p.u;
    p.u4 = p.u2 * p.u2;
    p.u8 = p.u4 * p.u4;
    return p;
}
// one (root, weight) pair from its 14 coefficient pairs c[0..13]; LD(c + k) loads pair k
template <class LD>
__device__ __forceinline__ void rys_estrin(const double2* __restrict__ c, const RysPowers& p, double& root, double& weight, LD ld)
{
    static_assert(RYS_NCOEF == 14, "Estrin tree below is written for degree 13");
    double2 a[RYS_NCOEF];
#pragma unroll
    for (int k = 0; k < RYS_NCOEF; k++) a[k] = ld(c + k);
    double r[7], w[7];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        r[k] = fma(a[2 * k + 1].x, p.u, a[2 * k].x);
        w[k] = fma(a[2 * k + 1].y, p.u, a[2 * k].y);
    }
    const double r01 = fma(r[1], p.u2, r[0]), r23 = fma(r[3], p.u2, r[2]), r45 = fma(r[5], p.u2, r[4]);
    const double w01 = fma(w[1], p.u2, w[0]), w23 = fma(w[3], p.u2, w[2]), w45 = fma(w[5], p.u2, w[4]);
    const double ra = fma(r23, p.u4, r01), rb = fma(r[6], p.u4, r45);
    const double wa = fma(w23, p.u4, w01), wb = fma(w[6], p.u4, w45);
    root = fma(rb, p.u8, ra);
    weight = fma(wb, p.u8, wa);
}
struct RysLdg {
    __device__ __forceinline__ double2 operator()(const double2* q) const { return __ldg(q); }
};
struct RysLds {
    __device__ __forceinline__ double2 operator()(const double2* q) const { return *q; }
};

template <int NROOTS>
__device__ __forceinline__ void rys_roots(double x, double* __restrict__ rw)
{
    constexpr int TRI = NROOTS * (NROOTS - 1) / 2;
    constexpr double large_x = NROOTS * 5 + 35;
    if (x >= large_x) {
        const double rs_x = jrsqrt(x);
        const double inv_x = rs_x * rs_x;
        const double t = SQRTPIE4 * rs_x;
#pragma unroll
        for (int i = 0; i < NROOTS; i++) {
            rw[2 * i] = RYS_LARGEX[(TRI + i) * 2] * inv_x;
            rw[2 * i + 1] = RYS_LARGEX[(TRI + i) * 2 + 1] * t;
        }
        return;
    }
    const RysPowers p = rys_powers(x);
    const double2* __restrict__ blk =
        reinterpret_cast<const double2*>(RYS_MONO + RYS_CHEB_OFFSET[NROOTS - 1]) + (size_t)p.it * NROOTS * RYS_NCOEF;
#pragma unroll
    for (int i = 0; i < NROOTS; i++) rys_estrin(blk + i * RYS_NCOEF, p, rw[2 * i], rw[2 * i + 1], RysLdg());
}

// ---------------------------------------------------------------------------------------
// Shared-memory copy of the NROOTS block of the table.  The per-lane interval gathers of
// rys_roots (14 coefficient pairs per root, a different interval in every lane) saturate the
// L1/TEX path when they go to global memory (profiles/r2: 31 % of the stall samples of the
// (ps|ps) class); from shared memory the same gathers are 16-byte LDS with rows padded to 15
// quad-words, so that 8 different intervals fall into 8 different bank groups.
template <int NROOTS>
struct RysSmem {
    static constexpr int NINT = 14 + 2 * NROOTS;         // intervals of width 2.5 below x = 35 + 5 n
    static constexpr int ROW = RYS_NCOEF + 1;            // padded row (quad-words)
    static constexpr int QUADS = NROOTS * NINT * ROW;
    static constexpr size_t BYTES = (size_t)QUADS * sizeof(double2);
};

template <int NROOTS>
__device__ __forceinline__ void rys_table_to_smem(double2* __restrict__ s_tab)
{
    using T = RysSmem<NROOTS>;
    const double2* __restrict__ src = reinterpret_cast<const double2*>(RYS_MONO + RYS_CHEB_OFFSET[NROOTS - 1]);
    for (int idx = threadIdx.x; idx < NROOTS * T::NINT * RYS_NCOEF; idx += blockDim.x) {
        const int k = idx % RYS_NCOEF, row = idx / RYS_NCOEF;
        const int it = row / NROOTS, i = row - it * NROOTS;
        s_tab[(i * T::NINT + it) * T::ROW + k] = src[idx];
    }
    __syncthreads();
}

template <int NROOTS>
__device__ __forceinline__ void rys_roots_smem(double x, double* __restrict__ rw, const double2* __restrict__ s_tab)
{
    using T = RysSmem<NROOTS>;
    constexpr int TRI = NROOTS * (NROOTS - 1) / 2;
    constexpr double large_x = NROOTS * 5 + 35;
    if (x >= large_x) {
        const double rs_x = jrsqrt(x);
        const double inv_x = rs_x * rs_x;
        const double t = SQRTPIE4 * rs_x;
#pragma unroll
        for (int i = 0; i < NROOTS; i++) {
            rw[2 * i] = RYS_LARGEX[(TRI + i) * 2] * inv_x;
            rw[2 * i + 1] = RYS_LARGEX[(TRI + i) * 2 + 1] * t;
        }
        return;
    }
    const RysPowers p = rys_powers(x);
    const double2* __restrict__ blk = s_tab + p.it * T::ROW;
#pragma unroll
    for (int i = 0; i < NROOTS; i++) rys_estrin(blk + i * T::NINT * T::ROW, p, rw[2 * i], rw[2 * i + 1], RysLds());
}

// ---------------------------------------------------------------------------------------
// FP32 copy of the same table and evaluation, for the mixed-precision band (quartets whose
// Schwarz x density estimate lies between cutoff_fp32 and cutoff_fp64; the reference evaluates
// those with DataType = float throughout, jqc/pyscf/jk.py:241-328, rys_roots.cu with float).
template <int NROOTS>
struct RysSmemF {
    static constexpr int NINT = 14 + 2 * NROOTS;
    static constexpr int ROW = RYS_NCOEF + 1;            // padded row of float2: 15 x 8 B, odd number of 8-byte words
    static constexpr int PAIRS = NROOTS * NINT * ROW;
    static constexpr size_t BYTES = ((size_t)PAIRS * sizeof(float2) + 15) / 16 * 16;
};

template <int NROOTS>
__device__ __forceinline__ void rys_table_to_smem_f(float2* __restrict__ s_tab)
{
    using T = RysSmemF<NROOTS>;
    const double2* __restrict__ src = reinterpret_cast<const double2*>(RYS_MONO + RYS_CHEB_OFFSET[NROOTS - 1]);
    for (int idx = threadIdx.x; idx < NROOTS * T::NINT * RYS_NCOEF; idx += blockDim.x) {
        const int k = idx % RYS_NCOEF, row = idx / RYS_NCOEF;
        const int it = row / NROOTS, i = row - it * NROOTS;
        const double2 v = src[idx];
        s_tab[(i * T::NINT + it) * T::ROW + k] = make_float2((float)v.x, (float)v.y);
    }
    __syncthreads();
}

template <int NROOTS>
__device__ __forceinline__ void rys_roots_smem_f(float x, float* __restrict__ rw, const float2* __restrict__ s_tab)
{
    using T = RysSmemF<NROOTS>;
    constexpr int TRI = NROOTS * (NROOTS - 1) / 2;
    constexpr float large_x = NROOTS * 5 + 35;
    if (x >= large_x) {
        const float rs_x = jrsqrt(x);
        const float inv_x = rs_x * rs_x;
        const float t = (float)SQRTPIE4 * rs_x;
#pragma unroll
        for (int i = 0; i < NROOTS; i++) {
            rw[2 * i] = (float)RYS_LARGEX[(TRI + i) * 2] * inv_x;
            rw[2 * i + 1] = (float)RYS_LARGEX[(TRI + i) * 2 + 1] * t;
        }
        return;
    }
    const int it = (int)(x * 0.4f);
    const float u = fmaf(x - it * 2.5f, 0.8f, -1.0f);
    const float u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
    const float2* __restrict__ blk = s_tab + it * T::ROW;
#pragma unroll
    for (int i = 0; i < NROOTS; i++) {
        const float2* __restrict__ c = blk + i * T::NINT * T::ROW;
        float2 a[RYS_NCOEF];
#pragma unroll
        for (int k = 0; k < RYS_NCOEF; k++) a[k] = c[k];
        float r[7], w[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            r[k] = fmaf(a[2 * k + 1].x, u, a[2 * k].x);
            w[k] = fmaf(a[2 * k + 1].y, u, a[2 * k].y);
        }
        const float r01 = fmaf(r[1], u2, r[0]), r23 = fmaf(r[3], u2, r[2]), r45 = fmaf(r[5], u2, r[4]);
        const float w01 = fmaf(w[1], u2, w[0]), w23 = fmaf(w[3], u2, w[2]), w45 = fmaf(w[5], u2, w[4]);
        rw[2 * i] = fmaf(fmaf(r[6], u4, r45), u8, fmaf(r23, u4, r01));
        rw[2 * i + 1] = fmaf(fmaf(w[6], u4, w45), u8, fmaf(w23, u4, w01));
    }
}

// Sum N per-lane values over the 32 lanes of a warp with a reduce-scatter butterfly: each step
// halves the number of values a lane still carries (N/2 + N/4 + ... shuffles in total instead of
// 5 N for independent butterflies).  On return a lane holds in v[0 .. cnt) the warp totals of the
// elements start .. start + cnt - 1 (cnt <= warp_rs_final(N); the other slots are padding).
__host__ __device__ constexpr int warp_rs_final(int n)
{
    for (int s = 0; s < 5; s++) n = (n + 1) / 2;
    return n;
}

template <int N, int OFF>
struct WarpReduceScatter {
    template <class R>
    static __device__ __forceinline__ void run(R* __restrict__ v, const int lane, int& start, int& cnt)
    {
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int e = 0; e < H; e++) {
            const R lo = v[e];
            const R hi = (e + H < N) ? v[e + H] : R(0);
            const R send = up ? lo : hi;
            const R keep = up ? hi : lo;
            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if (up) { start += H; cnt -= H; }
        else cnt = cnt < H ? cnt : H;
        if constexpr (OFF > 1) WarpReduceScatter<H, OFF / 2>::run(v, lane, start, cnt);
    }
};

}  // namespace jqc

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

// Packed record: [x, y, z, ao_loc, c0, e0, c1, e1, c2, e2, 0, 0]
struct ShellHead {
    double x, y, z, ao;
};

// ---------------------------------------------------------------------------------------
// Rys roots (t^2) and weights for NROOTS points at argument x (already theta-scaled).
// Two regimes only: Chebyshev interval table (the first interval covers x -> 0, so no
// small-x branch) and the Hermite asymptote for x >= 35 + 5 n.  Same tables and accuracy
// class as the reference's rys_roots (jqc/backend/rys/rys_roots.cu:29-160); the range
// separation scaling for omega > 0 follows :42-47.
// rw[2i] = root, rw[2i+1] = weight.
template <int NROOTS>
__device__ __forceinline__ void rys_roots(double x, double* __restrict__ rw)
{
    constexpr int TRI = NROOTS * (NROOTS - 1) / 2;
    constexpr double large_x = NROOTS * 5 + 35;
    if (x >= large_x) {
        const double inv_x = 1.0 / x;
        const double t = SQRTPIE4 * sqrt(inv_x);
#pragma unroll
        for (int i = 0; i < NROOTS; i++) {
            rw[2 * i] = RYS_LARGEX[(TRI + i) * 2] * inv_x;
            rw[2 * i + 1] = RYS_LARGEX[(TRI + i) * 2 + 1] * t;
        }
        return;
    }
    const int it = (int)(x * 0.4);
    const double u = fma(x - it * 2.5, 0.8, -1.0);
    const double u2 = 2.0 * u;
    const double2* __restrict__ blk =
        reinterpret_cast<const double2*>(RYS_CHEB + RYS_CHEB_OFFSET[NROOTS - 1]) + (size_t)it * NROOTS * RYS_NCOEF;
#pragma unroll
    for (int i = 0; i < NROOTS; i++) {
        const double2* __restrict__ c = blk + i * RYS_NCOEF;
        double2 a = __ldg(c + RYS_NCOEF - 1);
        double r1 = a.x, w1 = a.y, r2 = 0.0, w2 = 0.0;
#pragma unroll
        for (int k = RYS_NCOEF - 2; k >= 1; k--) {
            a = __ldg(c + k);
            const double r0 = fma(u2, r1, a.x) - r2;
            const double w0 = fma(u2, w1, a.y) - w2;
            r2 = r1; r1 = r0;
            w2 = w1; w1 = w0;
        }
        a = __ldg(c);
        rw[2 * i] = fma(u, r1, a.x) - r2;
        rw[2 * i + 1] = fma(u, w1, a.y) - w2;
    }
}

// ---------------------------------------------------------------------------------------
// Shared-memory copy of the NROOTS block of the Chebyshev table.  The per-lane interval gathers of
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
    const double2* __restrict__ src = reinterpret_cast<const double2*>(RYS_CHEB + RYS_CHEB_OFFSET[NROOTS - 1]);
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
        const double inv_x = 1.0 / x;
        const double t = SQRTPIE4 * sqrt(inv_x);
#pragma unroll
        for (int i = 0; i < NROOTS; i++) {
            rw[2 * i] = RYS_LARGEX[(TRI + i) * 2] * inv_x;
            rw[2 * i + 1] = RYS_LARGEX[(TRI + i) * 2 + 1] * t;
        }
        return;
    }
    const int it = (int)(x * 0.4);
    const double u = fma(x - it * 2.5, 0.8, -1.0);
    const double u2 = 2.0 * u;
    const double2* __restrict__ blk = s_tab + it * T::ROW;
#pragma unroll
    for (int i = 0; i < NROOTS; i++) {
        const double2* __restrict__ c = blk + i * T::NINT * T::ROW;
        double2 a = c[RYS_NCOEF - 1];
        double r1 = a.x, w1 = a.y, r2 = 0.0, w2 = 0.0;
#pragma unroll
        for (int k = RYS_NCOEF - 2; k >= 1; k--) {
            a = c[k];
            const double r0 = fma(u2, r1, a.x) - r2;
            const double w0 = fma(u2, w1, a.y) - w2;
            r2 = r1; r1 = r0;
            w2 = w1; w1 = w0;
        }
        a = c[0];
        rw[2 * i] = fma(u, r1, a.x) - r2;
        rw[2 * i + 1] = fma(u, w1, a.y) - w2;
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
    static __device__ __forceinline__ void run(double* __restrict__ v, const int lane, int& start, int& cnt)
    {
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int e = 0; e < H; e++) {
            const double lo = v[e];
            const double hi = (e + H < N) ? v[e + H] : 0.0;
            const double send = up ? lo : hi;
            const double keep = up ? hi : lo;
            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if (up) { start += H; cnt -= H; }
        else cnt = cnt < H ? cnt : H;
        if constexpr (OFF > 1) WarpReduceScatter<H, OFF / 2>::run(v, lane, start, cnt);
    }
};

}  // namespace jqc

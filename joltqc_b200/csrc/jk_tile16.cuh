// Small-class J/K kernel: one (i,j) shell pair x one 4x4 (k,l) shell tile per half-warp.
//
// Why: for blocks of <= 81 integrals the one-quartet-per-thread kernel is bound by the L2's
// FP64 atomic rate (~1.2e11 RED/s measured on B200, profiles/microbench/atomics.cu), not by the
// FP64 pipe.  Here the task generator hands out whole 4x4 tiles of (k,l) shells with a 16-bit
// survivor mask (the reference builds the same masks and then flattens them,
// jqc/backend/jk/screen_jk_tasks.cu:263-339); lane h of a half-warp evaluates quartet
// (i, j, 4*tk + h/4, 4*tl + h%4) in registers and the six contributions are combined with warp
// shuffles before they reach memory:
//     J_ij           summed over all 16 lanes            -> 1/16 of the atomics
//     K_ik, K_jk     summed over the 4 lanes sharing k   -> 1/4
//     K_il, K_jl     summed over the 4 lanes sharing l   -> 1/4
//     J_kl           unique per lane                     -> unchanged
// Reduce-scatter steps keep the shuffle count close to one exchange per element.
#pragma once
#include "jk_1q1t.cuh"

namespace jqc {

// One reduce-scatter step across the lane pair (lane, lane ^ XM): afterwards x[0..E/2) holds the
// pair sum of elements [hi*E/2, (hi+1)*E/2) where hi = (lane & XM) != 0.
template <int E>
__device__ __forceinline__ void rs_step(double (&x)[E], const bool hi, const int xm)
{
    static_assert(E % 2 == 0, "E must be even");
#pragma unroll
    for (int e = 0; e < E / 2; e++) {
        const double keep = hi ? x[e + E / 2] : x[e];
        const double send = hi ? x[e] : x[e + E / 2];
        x[e] = keep + __shfl_xor_sync(0xffffffffu, send, xm);
    }
}

__host__ __device__ constexpr int pad4(int n) { return (n + 3) / 4 * 4; }

// Sum a block over the 4 lanes {lane ^ 0, ^XA, ^XB, ^XA^XB} and add each lane's quarter to memory.
// Element e of the (padded) block lives at out[row(e) * nao + col(e)] with e = r * NC + c.
template <int NR, int NC>
__device__ __forceinline__ void reduce4_flush(double (&x)[pad4(NR * NC)], const int lane, const int xa, const int xb,
                                              const bool any, double* __restrict__ out, const int nao)
{
    constexpr int E = pad4(NR * NC);
    const bool ha = (lane & xa) != 0, hb = (lane & xb) != 0;
    rs_step<E>(x, ha, xa);
    double (&y)[E / 2] = reinterpret_cast<double (&)[E / 2]>(x);
    if constexpr ((E / 2) % 2 == 0) {
        rs_step<E / 2>(y, hb, xb);
        if (any) {
            const int base = (ha ? E / 2 : 0) + (hb ? E / 4 : 0);
#pragma unroll
            for (int e = 0; e < E / 4; e++) {
                const int g = base + e;
                if (g < NR * NC) atomicAdd(out + (size_t)(g / NC) * nao + (g % NC), x[e]);
            }
        }
    }
}

// Sum a block over all 16 lanes of the half-warp: two reduce-scatter steps (xor 1, 2), two
// all-reduce steps (xor 4, 8) on the remaining quarter; the four lanes with (lane & 12) == 0 flush.
template <int NR, int NC>
__device__ __forceinline__ void reduce16_flush(double (&x)[pad4(NR * NC)], const int lane, const bool any,
                                               double* __restrict__ out, const int nao)
{
    constexpr int E = pad4(NR * NC);
    const bool ha = (lane & 1) != 0, hb = (lane & 2) != 0;
    rs_step<E>(x, ha, 1);
    double (&y)[E / 2] = reinterpret_cast<double (&)[E / 2]>(x);
    rs_step<E / 2>(y, hb, 2);
#pragma unroll
    for (int e = 0; e < E / 4; e++) {
        x[e] += __shfl_xor_sync(0xffffffffu, x[e], 4);
        x[e] += __shfl_xor_sync(0xffffffffu, x[e], 8);
    }
    if (any && (lane & 12) == 0) {
        const int base = (ha ? E / 2 : 0) + (hb ? E / 4 : 0);
#pragma unroll
        for (int e = 0; e < E / 4; e++) {
            const int g = base + e;
            if (g < NR * NC) atomicAdd(out + (size_t)(g / NC) * nao + (g % NC), x[e]);
        }
    }
}

// Blocks of up to 27 integrals fit a 128-register budget without spilling: two CTAs per SM double
// the warps available to hide the latency of the root-table and density gathers.
template <int LI, int LJ, int LK, int LL>
constexpr int tile16_min_blocks() { return QuartetShape<LI, LJ, LK, LL>::N <= 27 ? 2 : 1; }

template <int LI, int LJ, int LK, int LL, bool DO_J, bool DO_K, int NTHREADS>
__global__ void __launch_bounds__(NTHREADS, tile16_min_blocks<LI, LJ, LK, LL>()) jk_tile16_kernel(const JKArgs a)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    constexpr int NFI = S::NFI, NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL, N = S::N;
    constexpr int NROOTS = S::NROOTS, GS = S::GSIZE, DJ = S::DJ, DK = S::DK, DL = S::DL;
    static_assert(pad4(NFI * NFJ) % 4 == 0, "");

    const uint4* __restrict__ tiles = reinterpret_cast<const uint4*>(a.quartets);
    const unsigned nent = *a.ntasks;
    const int nao = a.nao;
    const size_t nao2 = (size_t)nao * nao;
    const int lane = threadIdx.x & 31, h = lane & 15;
    const int kk = h >> 2, ll = h & 3;
    const unsigned nhw = gridDim.x * (NTHREADS / 16);
    const unsigned hw0 = (blockIdx.x * NTHREADS + threadIdx.x) >> 4;
    // both halves of a warp iterate the same number of times (shuffles are warp-wide)
    const unsigned niter = (nent + nhw - 1) / nhw;
#pragma unroll 1
    for (unsigned it = 0; it < niter; it++) {
        const unsigned ent = it * nhw + hw0;
        uint4 t = make_uint4(0, 0, 0, 0);
        if (ent < nent) t = tiles[ent];
        const unsigned mask = t.z;
        const bool bit = (mask >> h) & 1u;
        const int ish = t.x & 0xffff, jsh = t.x >> 16;
        const int ksh = (t.y & 0xffff) * TILE + kk, lsh = (t.y >> 16) * TILE + ll;
        const double* __restrict__ bi = a.basis + ish * BASIS_STRIDE;
        const double* __restrict__ bj = a.basis + jsh * BASIS_STRIDE;
        const double* __restrict__ bk = a.basis + ksh * BASIS_STRIDE;
        const double* __restrict__ bl = a.basis + lsh * BASIS_STRIDE;
        const double4 ri = *reinterpret_cast<const double4*>(bi);
        const double4 rj = *reinterpret_cast<const double4*>(bj);
        const double4 rk = *reinterpret_cast<const double4*>(bk);
        const double4 rl = *reinterpret_cast<const double4*>(bl);

        double eri[N];
#pragma unroll
        for (int n = 0; n < N; n++) eri[n] = 0.0;

        if (bit) {
            double fac = PI_FAC;
            if (ish == jsh) fac *= 0.5;
            if (ksh == lsh) fac *= 0.5;
            if (ish == ksh && jsh == lsh) fac *= 0.5;
            const double rjri[3] = {rj.x - ri.x, rj.y - ri.y, rj.z - ri.z};
            const double rlrk[3] = {rl.x - rk.x, rl.y - rk.y, rl.z - rk.z};
            const double rr_ij = rjri[0] * rjri[0] + rjri[1] * rjri[1] + rjri[2] * rjri[2];
            const double rr_kl = rlrk[0] * rlrk[0] + rlrk[1] * rlrk[1] + rlrk[2] * rlrk[2];
#pragma unroll 1
            for (int kp = 0; kp < a.npk; kp++)
#pragma unroll 1
            for (int lp = 0; lp < a.npl; lp++) {
                const double2 cek = *reinterpret_cast<const double2*>(bk + 4 + 2 * kp);
                const double2 cel = *reinterpret_cast<const double2*>(bl + 4 + 2 * lp);
                const double akl = cek.y + cel.y;
                const double inv_akl = 1.0 / akl;
                const double al_akl = cel.y * inv_akl;
                const double ckcl = cek.x * cel.x * exp(-cek.y * al_akl * rr_kl);
                const double qx = fma(rlrk[0], al_akl, rk.x), qy = fma(rlrk[1], al_akl, rk.y), qz = fma(rlrk[2], al_akl, rk.z);
#pragma unroll 1
                for (int ip = 0; ip < a.npi; ip++)
#pragma unroll 1
                for (int jp = 0; jp < a.npj; jp++) {
                    const double2 cei = *reinterpret_cast<const double2*>(bi + 4 + 2 * ip);
                    const double2 cej = *reinterpret_cast<const double2*>(bj + 4 + 2 * jp);
                    const double aij = cei.y + cej.y;
                    const double inv_aij = 1.0 / aij;
                    const double aj_aij = cej.y * inv_aij;
                    const double cicj = fac * cei.x * cej.x * exp(-cei.y * aj_aij * rr_ij);
                    const double Rpq[3] = {fma(rjri[0], aj_aij, ri.x) - qx, fma(rjri[1], aj_aij, ri.y) - qy,
                                           fma(rjri[2], aj_aij, ri.z) - qz};
                    const double rr = Rpq[0] * Rpq[0] + Rpq[1] * Rpq[1] + Rpq[2] * Rpq[2];
                    const double inv_aijkl = 1.0 / (aij + akl);
                    const double theta = aij * akl * inv_aijkl;
                    const double gy0 = cicj * inv_aij * inv_akl * sqrt(inv_aijkl);
                    double rw[2 * NROOTS];
                    double theta_fac = 1.0, sqrt_theta_fac = 1.0;
                    if (a.omega > 0.0) {
                        const double o2 = a.omega * a.omega;
                        theta_fac = o2 / (o2 + theta);
                        sqrt_theta_fac = sqrt(theta_fac);
                    }
                    rys_roots<NROOTS>(rr * theta * theta_fac, rw);
#pragma unroll 1
                    for (int ir = 0; ir < NROOTS; ir++) {
                        const double rt = rw[2 * ir] * theta_fac;
                        const double wt = rw[2 * ir + 1] * sqrt_theta_fac;
                        const double rt_aa = rt * inv_aijkl;
                        const double rt_aij = rt_aa * akl, rt_akl = rt_aa * aij;
                        const double b10 = 0.5 * inv_aij * (1.0 - rt_aij);
                        const double b01 = 0.5 * inv_akl * (1.0 - rt_akl);
                        const double b00 = 0.5 * rt_aa;
                        double c0[3], cp[3];
#pragma unroll
                        for (int d = 0; d < 3; d++) {
                            c0[d] = fma(rjri[d], aj_aij, -rt_aij * Rpq[d]);
                            cp[d] = fma(rlrk[d], al_akl, rt_akl * Rpq[d]);
                        }
                        double g[3 * GS];
                        fill_g_small<LI, LJ, LK, LL>(g, ckcl, gy0, wt, c0, cp, b10, b01, b00, rjri, rlrk);
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                        for (int k = 0; k < NFK; k++)
#pragma unroll
                        for (int l = 0; l < NFL; l++) {
                            const int ax = CART_X[LI][i] + CART_X[LJ][j] * DJ + CART_X[LK][k] * DK + CART_X[LL][l] * DL;
                            const int ay = CART_Y[LI][i] + CART_Y[LJ][j] * DJ + CART_Y[LK][k] * DK + CART_Y[LL][l] * DL;
                            const int az = CART_Z[LI][i] + CART_Z[LJ][j] * DJ + CART_Z[LK][k] * DK + CART_Z[LL][l] * DL;
                            const int n = ((i * NFJ + j) * NFK + k) * NFL + l;
                            eri[n] = fma(g[ax] * g[GS + ay], g[2 * GS + az], eri[n]);
                        }
                    }
                }
            }
        }

        // ---- contractions + shuffle-combined scatter
        const int i0 = (int)ri.w, j0 = (int)rj.w, k0 = (int)rk.w, l0 = (int)rl.w;
        const bool any_k = ((mask >> (kk * 4)) & 0xfu) != 0;                          // lanes sharing k
        const bool any_l = ((mask >> ll) & 0x1111u) != 0;                             // lanes sharing l
#define ERI_(i, j, k, l) eri[(((i) * NFJ + (j)) * NFK + (k)) * NFL + (l)]
#pragma unroll 1
        for (int b = 0; b < a.n_dm; b++) {
            const double* __restrict__ dm = a.dm + b * nao2;
            if constexpr (DO_J) {
                double* __restrict__ vj = a.vj + b * nao2;
                if (bit) {   // vj[l,k] += sum_ij (ij|kl) D[j,i]: unique per lane
                    double d_ji[NFI * NFJ];
#pragma unroll
                    for (int i = 0; i < NFI; i++)
#pragma unroll
                        for (int j = 0; j < NFJ; j++) d_ji[i * NFJ + j] = __ldg(dm + (size_t)(j0 + j) * nao + i0 + i);
#pragma unroll
                    for (int l = 0; l < NFL; l++)
#pragma unroll
                        for (int k = 0; k < NFK; k++) {
                            double s = 0.0;
#pragma unroll
                            for (int i = 0; i < NFI; i++)
#pragma unroll
                                for (int j = 0; j < NFJ; j++) s = fma(ERI_(i, j, k, l), d_ji[i * NFJ + j], s);
                            atomicAdd(vj + (size_t)(l0 + l) * nao + k0 + k, s);
                        }
                }
                {   // vj[j,i] += sum_kl (ij|kl) D[l,k]: same (i,j) on all 16 lanes
                    double x[pad4(NFJ * NFI)];
#pragma unroll
                    for (int e = 0; e < pad4(NFJ * NFI); e++) x[e] = 0.0;
                    if (bit) {
                        double d_lk[NFK * NFL];
#pragma unroll
                        for (int k = 0; k < NFK; k++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) d_lk[k * NFL + l] = __ldg(dm + (size_t)(l0 + l) * nao + k0 + k);
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int i = 0; i < NFI; i++) {
                                double s = 0.0;
#pragma unroll
                                for (int k = 0; k < NFK; k++)
#pragma unroll
                                    for (int l = 0; l < NFL; l++) s = fma(ERI_(i, j, k, l), d_lk[k * NFL + l], s);
                                x[j * NFI + i] = s;
                            }
                    }
                    reduce16_flush<NFJ, NFI>(x, lane, mask != 0, vj + (size_t)j0 * nao + i0, nao);
                }
            }
            if constexpr (DO_K) {
                double* __restrict__ vk = a.vk + b * nao2;
                {   // vk[i,k] += sum_jl (ij|kl) D[j,l]: shared by the 4 lanes with the same k
                    double x[pad4(NFI * NFK)];
#pragma unroll
                    for (int e = 0; e < pad4(NFI * NFK); e++) x[e] = 0.0;
                    if (bit) {
                        double d[NFJ * NFL];
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) d[j * NFL + l] = __ldg(dm + (size_t)(j0 + j) * nao + l0 + l);
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) {
                                double s = 0.0;
#pragma unroll
                                for (int j = 0; j < NFJ; j++)
#pragma unroll
                                    for (int l = 0; l < NFL; l++) s = fma(ERI_(i, j, k, l), d[j * NFL + l], s);
                                x[i * NFK + k] = s;
                            }
                    }
                    reduce4_flush<NFI, NFK>(x, lane, 1, 2, any_k, vk + (size_t)i0 * nao + k0, nao);
                }
                {   // vk[i,l] += sum_jk (ij|kl) D[j,k]: shared by the 4 lanes with the same l
                    double x[pad4(NFI * NFL)];
#pragma unroll
                    for (int e = 0; e < pad4(NFI * NFL); e++) x[e] = 0.0;
                    if (bit) {
                        double d[NFJ * NFK];
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) d[j * NFK + k] = __ldg(dm + (size_t)(j0 + j) * nao + k0 + k);
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) {
                                double s = 0.0;
#pragma unroll
                                for (int j = 0; j < NFJ; j++)
#pragma unroll
                                    for (int k = 0; k < NFK; k++) s = fma(ERI_(i, j, k, l), d[j * NFK + k], s);
                                x[i * NFL + l] = s;
                            }
                    }
                    reduce4_flush<NFI, NFL>(x, lane, 4, 8, any_l, vk + (size_t)i0 * nao + l0, nao);
                }
                {   // vk[j,k] += sum_il (ij|kl) D[i,l]
                    double x[pad4(NFJ * NFK)];
#pragma unroll
                    for (int e = 0; e < pad4(NFJ * NFK); e++) x[e] = 0.0;
                    if (bit) {
                        double d[NFI * NFL];
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) d[i * NFL + l] = __ldg(dm + (size_t)(i0 + i) * nao + l0 + l);
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) {
                                double s = 0.0;
#pragma unroll
                                for (int i = 0; i < NFI; i++)
#pragma unroll
                                    for (int l = 0; l < NFL; l++) s = fma(ERI_(i, j, k, l), d[i * NFL + l], s);
                                x[j * NFK + k] = s;
                            }
                    }
                    reduce4_flush<NFJ, NFK>(x, lane, 1, 2, any_k, vk + (size_t)j0 * nao + k0, nao);
                }
                {   // vk[j,l] += sum_ik (ij|kl) D[i,k]
                    double x[pad4(NFJ * NFL)];
#pragma unroll
                    for (int e = 0; e < pad4(NFJ * NFL); e++) x[e] = 0.0;
                    if (bit) {
                        double d[NFI * NFK];
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) d[i * NFK + k] = __ldg(dm + (size_t)(i0 + i) * nao + k0 + k);
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) {
                                double s = 0.0;
#pragma unroll
                                for (int i = 0; i < NFI; i++)
#pragma unroll
                                    for (int k = 0; k < NFK; k++) s = fma(ERI_(i, j, k, l), d[i * NFK + k], s);
                                x[j * NFL + l] = s;
                            }
                    }
                    reduce4_flush<NFJ, NFL>(x, lane, 4, 8, any_l, vk + (size_t)j0 * nao + l0, nao);
                }
            }
        }
#undef ERI_
    }
}

}  // namespace jqc

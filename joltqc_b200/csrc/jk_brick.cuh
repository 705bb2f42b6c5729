// "Brick" FP64 Rys J/K kernel for the angular classes whose integral block fits one thread's
// registers (<= JQC_SMALL_N integrals).
//
// Replaces, for those classes, the pair screen_jk_tasks -> rys_1q1t_vjk of the reference
// (jqc/backend/jk/screen_jk_tasks.cu:75-340, jqc/backend/jk/1q1t.cu:45-644; SURVEY rows a9 + a11).
// The reference (and round 1 of this repository) materialises a list of shell quartets and lets
// every quartet scatter its six J/K blocks with one FP64 atomic per element: <= 6 nf^2 atomics per
// quartet, which on B200 sit on the L2's atomic rate (profiles/microbench/atomics.cu).  Here the
// output is held stationary instead:
//
//   * a warp owns a BRICK: 32 (k,l) shell pairs of one group pair (one per lane, neighbours in a
//     q-descending pair list, so the 32 Schwarz bounds are nearly equal and the lanes pass or fail
//     the screening together) x a range of bra shells i;
//   * for each i the warp walks i's partner list j (q-descending, so the loop ends at the first j
//     whose bound fails) and every lane evaluates (ij|kl) for its own (k,l) in registers;
//   * J_kl stays in the lane's registers for the whole brick, K_ik and K_il for the whole j loop of
//     one i; J_ij is the same address for all 32 lanes and is summed with shuffles; only K_jk and
//     K_jl (nfj (nfk + nfl) elements, the two smallest blocks because lj <= li) are scattered per
//     quartet.  For (ps|ps) that is 4 atomics per quartet instead of 19;
//   * the Schwarz x density test of the reference (same float32 arithmetic, same canonical order,
//     same tile-pair prefilter) runs per lane inside the loop: no quartet list, no second kernel.
#pragma once
#include <type_traits>

#include "jk_1q1t.cuh"

namespace jqc {

struct BrickArgs {
    int nao, nbas;
    int npi, npj, npk, npl;
    const double* __restrict__ basis;
    const double* __restrict__ dm;            // one kernel-side density matrix (nao x nao)
    double* __restrict__ vj;
    double* __restrict__ vk;
    double omega;
    const float* __restrict__ logd;           // nbas x nbas log density pool (jk.py:179-184)
    const int* __restrict__ log_max_ordered;
    float cutoff;                             // log(cutoff): evaluate quartets whose estimate is above
    float cutoff_hi;                          // ... and not above this (+inf unless this launch is the FP32 band)
    const float* __restrict__ dm32;           // float copy of dm (FP32-band launches only)
    // ket side: ordered pair list of the (gk, gl) group pair
    const ushort2* __restrict__ kl;
    const float* __restrict__ kl_q;           // q of the pair
    const float* __restrict__ kl_tq;          // max q of the pair's 4x4 tile (tile-pair prefilter, jk.py:385-431)
    int n_kl;
    // bra side: shells i of group gi, each with a q-descending list of partners j in group gj
    int i_first, i_count;
    const int* __restrict__ j_off;            // i_count + 1 offsets into the three arrays below
    const unsigned short* __restrict__ j_idx;
    const float* __restrict__ j_q;
    const float* __restrict__ j_tq;
    float qmax_ij;                            // largest q of the bra group pair
    int tri;                                  // gi == gk: k <= i and (k,l) <= (i,j) must be tested per lane
    int ichunk, n_ichunk, n_blk;              // task = (ket block of 32 pairs) x (chunk of bra shells) x (slice of the j lists)
    int jsplit;                               // slices per j list (> 1 only when a launch has too few tasks to fill the GPU)
    int n_ij, ichunk_req;                     // bra pairs in the lists and the requested bra chunk (the launcher decomposes)
    // primitive-pair tables (engine_kernels.cuh: pair_prim_kernel), 8 doubles per primitive pair
    const double* __restrict__ bra_tab;       // entry of j-list element e at (e - j_base) * npi * npj * 8
    const double* __restrict__ ket_tab;       // entry of ket pair p at p * npk * npl * 8
    int j_base;
    int rank, world;                          // this GPU takes tasks shard_entry(0, rank, world), shard_entry(1, ...), ...
    unsigned* __restrict__ work;              // dynamic task counter (zeroed per build)
    unsigned long long* __restrict__ qcount;  // evaluated quartets (accounting)
};

// Task decomposition of a launch: `per_task` ket pairs per warp task (32 for the one-lane-per-quartet
// kernel, 32/T for the multi-lane kernel); bra chunks of ichunk_req shells, fewer when the launch
// would not fill the GPU, and for small molecules the j lists are sliced as well so that one
// warp's serial chain stays short.
inline void brick_decompose(BrickArgs& b, int per_task, int nsm)
{
    b.n_blk = (b.n_kl + per_task - 1) / per_task;
    const long long want = 4LL * nsm * 16 * b.world;
    int ic = b.ichunk_req < 1 ? 1 : b.ichunk_req;
    while (ic > 1 && (long long)b.n_blk * ((b.i_count + ic - 1) / ic) < want) ic >>= 1;
    b.ichunk = ic;
    b.n_ichunk = (b.i_count + ic - 1) / ic;
    b.jsplit = 1;
    const long long have = (long long)b.n_blk * b.n_ichunk;
    const int jmax = b.n_ij / (b.i_count > 0 ? b.i_count : 1) / 2;           // >= 2 partners per slice on average
    if (have < want / 2 && jmax > 1) {
        long long js = (want / 2 + have - 1) / have;
        if (js > 16) js = 16;
        if (js > jmax) js = jmax;
        b.jsplit = (int)(js < 1 ? 1 : js);
    }
}

// One shell quartet, all in registers: eri[N] += contracted integrals (reference: 1q1t.cu:86-405).
// KET_REG: the first primitive-pair record of the lane's ket pair arrives in registers (kr0, kr1; it is the
// same for every quartet of a brick), so single-primitive kets issue no table load per quartet.
// [I0, I1): the range of i components this call evaluates (eri holds (I1 - I0) * NFJ * NFK * NFL values); the
// two-pass variant of the largest blocks calls it once per half so that only half the block is live.
template <int LI, int LJ, int LK, int LL, bool KET_REG = false, int I0 = 0, int I1 = nf_of(LI)>
__device__ __forceinline__ void eri_block_regs(double* __restrict__ eri, const double* __restrict__ bra,
                                               const double* __restrict__ ket, const double4 ri, const double4 rj,
                                               const double4 rk, const double4 rl, const int npij, const int npkl,
                                               const double omega, const double fac, const double2* __restrict__ s_rys,
                                               const double4 kr0 = double4(), const double4 kr1 = double4())
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    constexpr int NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL, N = (I1 - I0) * NFJ * NFK * NFL;
    constexpr int NROOTS = S::NROOTS, GS = S::GSIZE, DJ = S::DJ, DK = S::DK, DL = S::DL;
    static_assert(0 <= I0 && I0 < I1 && I1 <= S::NFI, "i range");
    const double rjri[3] = {rj.x - ri.x, rj.y - ri.y, rj.z - ri.z};
    const double rlrk[3] = {rl.x - rk.x, rl.y - rk.y, rl.z - rk.z};
#pragma unroll
    for (int n = 0; n < N; n++) eri[n] = 0.0;
#pragma unroll 1
    for (int klp = 0; klp < npkl; klp++) {
        double4 k0, k1;
        if (KET_REG && klp == 0) { k0 = kr0; k1 = kr1; }
        else {
            k0 = *reinterpret_cast<const double4*>(ket + klp * 8);
            k1 = *reinterpret_cast<const double4*>(ket + klp * 8 + 4);
        }
        const double akl = k0.x, inv_akl = k0.y, al_akl = k0.z, ckcl = k0.w;
        const double qx = k1.x, qy = k1.y, qz = k1.z;
#pragma unroll 1
        for (int ipj = 0; ipj < npij; ipj++) {
            const double4 b0 = *reinterpret_cast<const double4*>(bra + ipj * 8);
            const double4 b1 = *reinterpret_cast<const double4*>(bra + ipj * 8 + 4);
            const double aij = b0.x, inv_aij = b0.y, aj_aij = b0.z;
            const double cicj = fac * b0.w;
            const double Rpq[3] = {b1.x - qx, b1.y - qy, b1.z - qz};
            const double rr = Rpq[0] * Rpq[0] + Rpq[1] * Rpq[1] + Rpq[2] * Rpq[2];
            const double rs_aijkl = jrsqrt(aij + akl);
            const double inv_aijkl = rs_aijkl * rs_aijkl;
            const double theta = aij * akl * inv_aijkl;
            const double gy0 = cicj * inv_aij * inv_akl * rs_aijkl;
            double rw[2 * NROOTS];
            double theta_fac = 1.0, sqrt_theta_fac = 1.0;
            if (omega > 0.0) {
                const double o2 = omega * omega;
                theta_fac = o2 / (o2 + theta);
                sqrt_theta_fac = sqrt(theta_fac);
            }
            rys_roots_smem<NROOTS>(rr * theta * theta_fac, rw, s_rys);
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
                for (int i = I0; i < I1; i++)
#pragma unroll
                for (int j = 0; j < NFJ; j++)
#pragma unroll
                for (int k = 0; k < NFK; k++)
#pragma unroll
                for (int l = 0; l < NFL; l++) {
                    const int ax = CART_X[LI][i] + CART_X[LJ][j] * DJ + CART_X[LK][k] * DK + CART_X[LL][l] * DL;
                    const int ay = CART_Y[LI][i] + CART_Y[LJ][j] * DJ + CART_Y[LK][k] * DK + CART_Y[LL][l] * DL;
                    const int az = CART_Z[LI][i] + CART_Z[LJ][j] * DJ + CART_Z[LK][k] * DK + CART_Z[LL][l] * DL;
                    const int n = (((i - I0) * NFJ + j) * NFK + k) * NFL + l;
                    eri[n] = fma(g[ax] * g[GS + ay], g[2 * GS + az], eri[n]);
                }
            }
        }
    }
}

// Same quartet in FP32 (mixed-precision band; the reference instantiates its kernel with DataType =
// float, jqc/pyscf/jk.py:241-262).  Shell-centre and primitive-centre differences are formed in FP64
// from the FP64 tables and rounded once, everything after that is float.
template <class R, int LI, int LJ, int LK, int LL>
__device__ __forceinline__ void fill_g_t(R* __restrict__ g, const R seed0, const R seed1, const R seed2, const R* c0,
                                         const R* cp, const R b10, const R b01, const R b00, const R* rjri, const R* rlrk)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    constexpr int GS = S::GSIZE, DJ = S::DJ, DK = S::DK, DL = S::DL, LIJ = S::LIJ, LKL = S::LKL;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        R* gd = g + d * GS;
        gd[0] = d == 0 ? seed0 : (d == 1 ? seed1 : seed2);
        if constexpr (LIJ > 0) {
            gd[1] = c0[d] * gd[0];
#pragma unroll
            for (int i = 1; i < LIJ; i++) gd[i + 1] = c0[d] * gd[i] + (R(i) * b10) * gd[i - 1];
        }
        if constexpr (LKL > 0) {
#pragma unroll
            for (int i = 0; i <= LIJ; i++) {
                R v = cp[d] * gd[i];
                if (i > 0) v += R(i) * b00 * gd[i - 1];
                gd[i + DK] = v;
            }
#pragma unroll
            for (int k = 1; k < LKL; k++) {
                const R kb01 = R(k) * b01;
#pragma unroll
                for (int i = 0; i <= LIJ; i++) {
                    R v = cp[d] * gd[i + k * DK] + kb01 * gd[i + (k - 1) * DK];
                    if (i > 0) v += R(i) * b00 * gd[i - 1 + k * DK];
                    gd[i + (k + 1) * DK] = v;
                }
            }
        }
        if constexpr (LJ > 0) {
            const R ab = rjri[d];
#pragma unroll
            for (int k = 0; k <= LKL; k++)
#pragma unroll
                for (int j = 0; j < LJ; j++)
#pragma unroll
                    for (int i = LIJ - j - 1; i >= 0; i--) {
                        const int src = i + j * DJ + k * DK;
                        gd[src + DJ] = gd[src + 1] - ab * gd[src];
                    }
        }
        if constexpr (LL > 0) {
            const R cd = rlrk[d];
#pragma unroll
            for (int ij = 0; ij < DK; ij++)
#pragma unroll
                for (int l = 0; l < LL; l++)
#pragma unroll
                    for (int k = LKL - l - 1; k >= 0; k--) {
                        const int src = ij + k * DK + l * DL;
                        gd[src + DL] = gd[src + DK] - cd * gd[src];
                    }
        }
    }
}

template <int LI, int LJ, int LK, int LL, bool KET_REG = false>
__device__ __forceinline__ void eri_block_regs_f(float* __restrict__ eri, const double* __restrict__ bra,
                                                 const double* __restrict__ ket, const double4 ri, const double4 rj,
                                                 const double4 rk, const double4 rl, const int npij, const int npkl,
                                                 const float omega, const float fac, const float2* __restrict__ s_rys,
                                                 const double4 kr0 = double4(), const double4 kr1 = double4())
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    constexpr int NFI = S::NFI, NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL, N = S::N;
    constexpr int NROOTS = S::NROOTS, GS = S::GSIZE, DJ = S::DJ, DK = S::DK, DL = S::DL;
    const float rjri[3] = {(float)(rj.x - ri.x), (float)(rj.y - ri.y), (float)(rj.z - ri.z)};
    const float rlrk[3] = {(float)(rl.x - rk.x), (float)(rl.y - rk.y), (float)(rl.z - rk.z)};
#pragma unroll
    for (int n = 0; n < N; n++) eri[n] = 0.0f;
#pragma unroll 1
    for (int klp = 0; klp < npkl; klp++) {
        double4 k0, k1;
        if (KET_REG && klp == 0) { k0 = kr0; k1 = kr1; }
        else {
            k0 = *reinterpret_cast<const double4*>(ket + klp * 8);
            k1 = *reinterpret_cast<const double4*>(ket + klp * 8 + 4);
        }
        const float akl = (float)k0.x, inv_akl = (float)k0.y, al_akl = (float)k0.z, ckcl = (float)k0.w;
#pragma unroll 1
        for (int ipj = 0; ipj < npij; ipj++) {
            const double4 b0 = *reinterpret_cast<const double4*>(bra + ipj * 8);
            const double4 b1 = *reinterpret_cast<const double4*>(bra + ipj * 8 + 4);
            const float aij = (float)b0.x, inv_aij = (float)b0.y, aj_aij = (float)b0.z;
            const float cicj = fac * (float)b0.w;
            const float Rpq[3] = {(float)(b1.x - k1.x), (float)(b1.y - k1.y), (float)(b1.z - k1.z)};
            const float rr = Rpq[0] * Rpq[0] + Rpq[1] * Rpq[1] + Rpq[2] * Rpq[2];
            const float rs_aijkl = jrsqrt(aij + akl);
            const float inv_aijkl = rs_aijkl * rs_aijkl;
            const float theta = aij * akl * inv_aijkl;
            const float gy0 = cicj * inv_aij * inv_akl * rs_aijkl;
            float rw[2 * NROOTS];
            float theta_fac = 1.0f, sqrt_theta_fac = 1.0f;
            if (omega > 0.0f) {
                const float o2 = omega * omega;
                theta_fac = o2 / (o2 + theta);
                sqrt_theta_fac = sqrtf(theta_fac);
            }
            rys_roots_smem_f<NROOTS>(rr * theta * theta_fac, rw, s_rys);
#pragma unroll 1
            for (int ir = 0; ir < NROOTS; ir++) {
                const float rt = rw[2 * ir] * theta_fac;
                const float wt = rw[2 * ir + 1] * sqrt_theta_fac;
                const float rt_aa = rt * inv_aijkl;
                const float rt_aij = rt_aa * akl, rt_akl = rt_aa * aij;
                const float b10 = 0.5f * inv_aij * (1.0f - rt_aij);
                const float b01 = 0.5f * inv_akl * (1.0f - rt_akl);
                const float b00 = 0.5f * rt_aa;
                float c0[3], cp[3];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    c0[d] = fmaf(rjri[d], aj_aij, -rt_aij * Rpq[d]);
                    cp[d] = fmaf(rlrk[d], al_akl, rt_akl * Rpq[d]);
                }
                float g[3 * GS];
                fill_g_t<float, LI, LJ, LK, LL>(g, ckcl, gy0, wt, c0, cp, b10, b01, b00, rjri, rlrk);
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
                    eri[n] = fmaf(g[ax] * g[GS + ay], g[2 * GS + az], eri[n]);
                }
            }
        }
    }
}

// register budgets of the FP64 brick kernel by live doubles (integrals + stationary sums): 128 registers up to
// LIVE128, 168 up to LIVE168, 255 above
#ifndef JQC_BRICK_LIVE128
#define JQC_BRICK_LIVE128 12
#endif
#ifndef JQC_BRICK_LIVE168
#define JQC_BRICK_LIVE168 36
#endif
#ifndef JQC_BRICK_PIPE_N
#define JQC_BRICK_PIPE_N 1
#endif
// FP64 blocks of at least this many integrals (with an even number of i components and a ket other than (ss|)
// are evaluated and digested in two passes over the halves of the i components: half the integral block is live
// at a time (600-1000 B of spills per thread drop to 200-270 B) at the price of evaluating prefactors, roots and
// recurrences twice.  Measured on valinomycin/def2-TZVP (profiles/r2/class_times_33_*.csv): (dp|ds) -35 %,
// (ds|dp) -28 %, (dd|ps) -19 %, (fs|pp) -20 %, (fp|ps) -15 %; (ff|ss) +6 %, hence the ket condition.  0 = never.
#ifndef JQC_BRICK_SPLIT_N
#define JQC_BRICK_SPLIT_N 90
#endif

// Per-class layout of the brick kernel, usable at compile time (BrickPlan) and by the host (which
// classes are supported, how much dynamic shared memory a launch needs).  f32 = the FP32-band
// variant: integrals, g arrays, density blocks and the Rys table are float (half the registers and
// shared memory), the stationary J/K sums stay double.
struct BrickShape {
    int n, nki, njkl, nroots;
    bool acc_smem;      // per-i K accumulators in lane-private shared memory (else registers)
    bool di_smem;       // D_il / D_ik blocks (stationary over the j loop) staged in lane-private shared memory
    int dlk_mode;       // D_lk block (stationary over the brick): 0 registers, 1 lane-private shared memory, 2 reloaded
    int acc_slots, d_slots;   // lane-private doubles (accumulators) and reals (density blocks) per lane
    int regs, minb, nwarps;
    int isplit;         // passes over the i components (1, or 2 for the largest FP64 blocks)
    size_t rys_bytes, smem;
    bool fits;
};

__host__ __device__ constexpr BrickShape brick_shape(int li, int lj, int lk, int ll, bool f32 = false)
{
    BrickShape b{};
    const int nfi = nf_of(li), nfj = nf_of(lj), nfk = nf_of(lk), nfl = nf_of(ll);
    b.n = nfi * nfj * nfk * nfl;
    b.nki = nfi * (nfk + nfl);
    b.njkl = nfk * nfl;
    b.nroots = (li + lj + lk + ll) / 2 + 1;
    b.nwarps = 4;
    b.dlk_mode = b.njkl <= 3 ? 0 : (b.njkl <= 9 ? 1 : 2);
    b.isplit = (!f32 && JQC_BRICK_SPLIT_N > 0 && b.n >= JQC_BRICK_SPLIT_N && nfi % 2 == 0 && b.njkl > 1) ? 2 : 1;
    if (!f32) {
        const int live = b.n + b.nki + b.njkl;
        b.acc_smem = live > 80 && b.nki > 12;
        b.di_smem = b.nki <= 48;
        // register budget per thread -> CTAs of 128 threads per SM: 255 -> 2, 168 -> 3, 128 -> 4
        b.regs = live <= JQC_BRICK_LIVE128 ? 128 : (live <= JQC_BRICK_LIVE168 ? 168 : 255);
        b.rys_bytes = (size_t)b.nroots * (14 + 2 * b.nroots) * (RYS_NCOEF + 1) * 16;
    } else {
        // 32-bit registers: integrals + g arrays (float) + double accumulators
        const int gs = (li + 1) * (lj + 1) * (lk + 1) * (ll + 1);
        const int live = b.n + 3 * gs + 2 * b.njkl;
        b.acc_smem = live + 2 * b.nki > 64 && b.nki > 12;
        b.di_smem = b.nki <= 60;
        const int r = live + (b.acc_smem ? 0 : 2 * b.nki);
        b.regs = r <= 64 ? 128 : (r <= 140 ? 168 : 255);
        b.rys_bytes = ((size_t)b.nroots * (14 + 2 * b.nroots) * (RYS_NCOEF + 1) * 8 + 15) / 16 * 16;
    }
    b.acc_slots = b.acc_smem ? b.nki : 0;
    b.d_slots = (b.di_smem ? b.nki : 0) + (b.dlk_mode == 1 ? b.njkl : 0);
    b.minb = 65536 / (b.regs * b.nwarps * 32);
    b.smem = b.rys_bytes + (size_t)b.nwarps * 32 * (b.acc_slots * sizeof(double) + b.d_slots * (f32 ? 4 : 8));
    if (f32 && b.smem * b.minb > 216 * 1024) {      // shared memory, not registers, limits the residency
        b.minb = (int)(216 * 1024 / b.smem);
        b.regs = b.minb >= 3 ? 168 : 255;
    }
    b.fits = b.n <= JQC_SMALL_N && b.smem * b.minb <= 216 * 1024;
    return b;
}

template <class R, int LI, int LJ, int LK, int LL>
struct BrickPlan {
    static constexpr bool F32 = sizeof(R) == 4;
    static constexpr BrickShape B = brick_shape(LI, LJ, LK, LL, F32);
    static constexpr int NKI = B.nki, NJKL = B.njkl, NWARPS = B.nwarps, MINB = B.minb;
    static constexpr int ACC_SLOTS = B.acc_slots, D_SLOTS = B.d_slots;
    static constexpr bool ACC_SMEM = B.acc_smem, DI_SMEM = B.di_smem, FITS = B.fits;
    static constexpr int DLK_MODE = B.dlk_mode, ISPLIT = B.isplit;
    static constexpr size_t SMEM = B.smem, RYS_BYTES = B.rys_bytes;
    // density-block slot offsets (lane-private reals)
    static constexpr int S_DI = 0, S_DLK = DI_SMEM ? NKI : 0;
    // The smallest blocks are latency-, not FP64-bound (profiles/r2: long-scoreboard stalls on the dependent
    // list -> shell -> log-density loads of every j step and on the table / density loads of every
    // quartet).  PIPE runs the j loop software-pipelined (list entries two steps ahead, the data that
    // depend on the partner shell one step ahead), loads the density blocks of the digestion before the
    // integrals and keeps the ket's first primitive-pair record in registers for the whole brick.
    // Measured (profiles/r2/class_times_21_pipe54.csv): (ss|ss) -15 %, every larger class loses 0-35 %
    // to the extra live registers, hence the threshold of one integral.
    static constexpr bool PIPE = B.n <= JQC_BRICK_PIPE_N;
    static_assert(!(PIPE && ISPLIT > 1), "the pipelined j loop is for the smallest blocks only");
};

template <class R> struct BrickRysTab { using type = double2; };
template <> struct BrickRysTab<float> { using type = float2; };

template <class R, int LI, int LJ, int LK, int LL, bool DO_J, bool DO_K>
__global__ void __launch_bounds__(BrickPlan<R, LI, LJ, LK, LL>::NWARPS * 32, BrickPlan<R, LI, LJ, LK, LL>::MINB)
jk_brick_kernel(const BrickArgs a)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    using P = BrickPlan<R, LI, LJ, LK, LL>;
    using RysT = typename BrickRysTab<R>::type;
    constexpr bool F32 = P::F32;
    constexpr int NFI = S::NFI, NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double2 brick_smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nao = a.nao, nbas = a.nbas;
    const float log_max = ordered_to_float(*a.log_max_ordered);
    const float dmaxf = fmaxf(log_max, -36.8f);
    const double paircut = log(1e-13) - (double)log_max;       // jk.py:185-187, 412
    const unsigned ntask = (unsigned)a.n_blk * (unsigned)a.n_ichunk * (unsigned)a.jsplit;
    // shared memory: [Rys table of this class][accumulator slots (double)][density slots (R)], both
    // element-major, lane-minor within a warp
    const RysT* __restrict__ s_rys = reinterpret_cast<const RysT*>(brick_smem);
    if constexpr (F32) rys_table_to_smem_f<S::NROOTS>(reinterpret_cast<float2*>(brick_smem));
    else rys_table_to_smem<S::NROOTS>(brick_smem);
    double* __restrict__ aslot = reinterpret_cast<double*>(reinterpret_cast<char*>(brick_smem) + P::RYS_BYTES) +
                                 (size_t)warp * 32 * P::ACC_SLOTS + lane;
    R* __restrict__ dslot = reinterpret_cast<R*>(reinterpret_cast<char*>(brick_smem) + P::RYS_BYTES +
                                                 (size_t)P::NWARPS * 32 * P::ACC_SLOTS * sizeof(double)) +
                            (size_t)warp * 32 * P::D_SLOTS + lane;
    const int npij = a.npi * a.npj, npkl = a.npk * a.npl;
#define ASLOT_(x) aslot[(x) * 32]
#define DSLOT_(x) dslot[(x) * 32]
    // density elements in the precision of this launch
    auto ldd = [&](size_t off) -> R {
        if constexpr (F32) return __ldg(a.dm32 + off);
        else return __ldg(a.dm + off);
    };
    unsigned long long nq = 0;

#pragma unroll 1
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(a.work, 1u);
        t = shard_entry(__shfl_sync(FULL, t, 0), a.rank, a.world);
        if (t >= ntask) break;
        const int js = (int)(t % (unsigned)a.jsplit);
        t /= (unsigned)a.jsplit;
        const int blk = (int)(t % (unsigned)a.n_blk), ic = (int)(t / (unsigned)a.n_blk);
        const int p = blk * 32 + lane;
        const bool pvalid = p < a.n_kl;
        const int pp = pvalid ? p : blk * 32;
        const ushort2 kl = a.kl[pp];
        const float q_kl = a.kl_q[pp];
        const bool lane_on = pvalid && ((double)a.kl_tq[pp] > paircut);
        float Qb = lane_on ? q_kl : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Qb = fmaxf(Qb, __shfl_xor_sync(FULL, Qb, o));
        if (!(a.qmax_ij + Qb + dmaxf > a.cutoff)) continue;
        const int ksh = kl.x, lsh = kl.y;
        int kmin = lane_on ? ksh : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kmin = min(kmin, __shfl_xor_sync(FULL, kmin, o));
        const double* __restrict__ bk = a.basis + ksh * BASIS_STRIDE;
        const double* __restrict__ bl = a.basis + lsh * BASIS_STRIDE;
        const double4 rk = *reinterpret_cast<const double4*>(bk);
        const double4 rl = *reinterpret_cast<const double4*>(bl);
        const int k0 = (int)rk.w, l0 = (int)rl.w;
        const float d_kl = a.logd[(size_t)ksh * nbas + lsh];
        const double* __restrict__ ket = a.ket_tab + (size_t)pp * npkl * 8;
        double4 kr0 = double4(), kr1 = double4();
        if constexpr (P::PIPE) {
            kr0 = *reinterpret_cast<const double4*>(ket);
            kr1 = *reinterpret_cast<const double4*>(ket + 4);
        }

        double jkl[DO_J ? NFK * NFL : 1];
        R dlk_r[(DO_J && P::DLK_MODE == 0) ? NFK * NFL : 1];
        if constexpr (DO_J) {
#pragma unroll
            for (int k = 0; k < NFK; k++)
#pragma unroll
            for (int l = 0; l < NFL; l++) {
                jkl[k * NFL + l] = 0.0;
                if constexpr (P::DLK_MODE == 0) dlk_r[k * NFL + l] = ldd((size_t)(l0 + l) * nao + k0 + k);
                if constexpr (P::DLK_MODE == 1) DSLOT_(P::S_DLK + k * NFL + l) = ldd((size_t)(l0 + l) * nao + k0 + k);
            }
        }
        bool touched_kl = false;

        const int i_lo = a.i_first + ic * a.ichunk;
        const int i_hi = min(a.i_first + a.i_count, i_lo + a.ichunk);
#pragma unroll 1
        for (int ish = i_lo; ish < i_hi; ish++) {
            if (a.tri && ish < kmin) continue;
            int e = a.j_off[ish - a.i_first];
            int e_end = a.j_off[ish - a.i_first + 1];
            if (a.jsplit > 1) {
                const int len = (e_end - e + a.jsplit - 1) / a.jsplit;
                e += js * len;
                e_end = min(e_end, e + len);
            }
            if (e >= e_end) continue;
            if (!(a.j_q[e] + Qb + dmaxf > a.cutoff)) continue;
            const double* __restrict__ bi = a.basis + ish * BASIS_STRIDE;
            const double4 ri = *reinterpret_cast<const double4*>(bi);
            const int i0 = (int)ri.w;
            const float d_ik = a.logd[(size_t)ish * nbas + ksh], d_il = a.logd[(size_t)ish * nbas + lsh];
            const bool lane_i = lane_on && (!a.tri || ksh <= ish);

            double kacc[(DO_K && !P::ACC_SMEM) ? P::NKI : 1];
            if constexpr (DO_K) {
                // K_ik accumulators at [i * NFK + k], K_il at [NFI * NFK + i * NFL + l]; the D_ik / D_il
                // blocks use the same indexing in their own slot range
#pragma unroll
                for (int x = 0; x < P::NKI; x++) {
                    if constexpr (P::ACC_SMEM) ASLOT_(x) = 0.0;
                    else kacc[x] = 0.0;
                }
                if constexpr (P::DI_SMEM) {
#pragma unroll
                    for (int i = 0; i < NFI; i++) {
#pragma unroll
                        for (int k = 0; k < NFK; k++) DSLOT_(P::S_DI + i * NFK + k) = ldd((size_t)(i0 + i) * nao + k0 + k);
#pragma unroll
                        for (int l = 0; l < NFL; l++) DSLOT_(P::S_DI + NFI * NFK + i * NFL + l) = ldd((size_t)(i0 + i) * nao + l0 + l);
                    }
                }
            }
            bool touched_i = false;

            // pipeline registers: list entry e + 1 (n1), entry e + 2 (n2), partner-shell data of e + 1 (m1)
            struct L1 { float q, tq; int jsh; };
            struct L2 { float d_jk, d_jl, d_ij; double4 rj; };
            auto load1 = [&](int x) -> L1 {
                const int xc = min(x, e_end - 1);
                return L1{a.j_q[xc], a.j_tq[xc], (int)a.j_idx[xc]};
            };
            auto load2 = [&](const L1& l) -> L2 {
                L2 r;
                r.d_jk = DO_K ? a.logd[(size_t)l.jsh * nbas + ksh] : 0.f;
                r.d_jl = DO_K ? a.logd[(size_t)l.jsh * nbas + lsh] : 0.f;
                r.d_ij = DO_J ? a.logd[(size_t)ish * nbas + l.jsh] : 0.f;
                r.rj = *reinterpret_cast<const double4*>(a.basis + l.jsh * BASIS_STRIDE);
                return r;
            };
            L1 n1 = L1(), n2 = L1();
            L2 m1 = L2();
            if constexpr (P::PIPE) {
                n1 = load1(e);
                n2 = load1(e + 1);
                m1 = load2(n1);
            }
#pragma unroll 1
            for (; e < e_end; e++) {
                float q_ij, tq_ij, pd_jk = 0.f, pd_jl = 0.f, pd_ij = 0.f;
                int jsh;
                double4 rj;
                if constexpr (P::PIPE) {
                    q_ij = n1.q; tq_ij = n1.tq; jsh = n1.jsh;
                    pd_jk = m1.d_jk; pd_jl = m1.d_jl; pd_ij = m1.d_ij; rj = m1.rj;
                    n1 = n2;
                    m1 = load2(n1);          // entry e + 1: its shell index arrived one step ago
                    n2 = load1(e + 2);
                } else {
                    q_ij = a.j_q[e];
                }
                if (!(q_ij + Qb + dmaxf > a.cutoff)) break;         // q-descending list: nothing further passes
                if constexpr (!P::PIPE) tq_ij = a.j_tq[e];
                if (!((double)tq_ij > paircut)) continue;           // bra tile pair not active
                if constexpr (!P::PIPE) jsh = a.j_idx[e];
                // canonical order (screen_jk_tasks.cu:202, 225, 239) + Schwarz x density test (:241-261)
                bool live = lane_i && (!a.tri || ksh < ish || lsh <= jsh);
                if (live) {
                    const float q_ijkl = q_ij + q_kl;
                    float d_large = -36.8f;
                    if constexpr (DO_K) {
                        d_large = fmaxf(d_large, d_ik);
                        d_large = fmaxf(d_large, P::PIPE ? pd_jk : a.logd[(size_t)jsh * nbas + ksh]);
                        d_large = fmaxf(d_large, d_il);
                        d_large = fmaxf(d_large, P::PIPE ? pd_jl : a.logd[(size_t)jsh * nbas + lsh]);
                    }
                    if constexpr (DO_J) {
                        d_large = fmaxf(d_large, P::PIPE ? pd_ij : a.logd[(size_t)ish * nbas + jsh]);
                        d_large = fmaxf(d_large, d_kl);
                    }
                    // precision band of this launch (screen_jk_tasks.cu:258-261: sel_fp64 = dq > cutoff_fp64)
                    const float dq = q_ijkl + d_large;
                    live = dq > a.cutoff && !(dq > a.cutoff_hi);
                }
                const unsigned m = __ballot_sync(FULL, live);
                if (m == 0) continue;
                if (lane == 0) nq += __popc(m);
                touched_i |= live;

                if constexpr (!P::PIPE) rj = *reinterpret_cast<const double4*>(a.basis + jsh * BASIS_STRIDE);
                const int j0 = (int)rj.w;
                // density blocks of the digestion, requested before the integrals so that their latency
                // hides behind the ERI evaluation (small blocks only: they stay live in registers)
                R d_ji[(DO_J && P::PIPE) ? NFI * NFJ : 1], d_jl[(DO_K && P::PIPE) ? NFJ * NFL : 1], d_jk[(DO_K && P::PIPE) ? NFJ * NFK : 1];
                if constexpr (P::PIPE) {
                    if constexpr (DO_J) {
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                        for (int j = 0; j < NFJ; j++) d_ji[i * NFJ + j] = ldd((size_t)(j0 + j) * nao + i0 + i);
                    }
                    if constexpr (DO_K) {
#pragma unroll
                        for (int j = 0; j < NFJ; j++) {
#pragma unroll
                            for (int l = 0; l < NFL; l++) d_jl[j * NFL + l] = ldd((size_t)(j0 + j) * nao + l0 + l);
#pragma unroll
                            for (int k = 0; k < NFK; k++) d_jk[j * NFK + k] = ldd((size_t)(j0 + j) * nao + k0 + k);
                        }
                    }
                }
                R fac = live ? R(PI_FAC) : R(0);
                if (ish == jsh) fac *= R(0.5);
                if (ksh == lsh) fac *= R(0.5);
                if (ish == ksh && jsh == lsh) fac *= R(0.5);
                // evaluate + digest, in P::ISPLIT passes over the i components [IB, IE) (one pass = the whole block)
                auto pass = [&](auto HC) {
                constexpr int NIH = NFI / P::ISPLIT, IB = decltype(HC)::value * NIH, IE = IB + NIH;
                R eri[NIH * NFJ * NFK * NFL];
                if constexpr (F32)
                    eri_block_regs_f<LI, LJ, LK, LL, P::PIPE>(eri, a.bra_tab + (size_t)(e - a.j_base) * npij * 8, ket, ri, rj, rk, rl, npij,
                                                              npkl, (float)a.omega, fac, s_rys, kr0, kr1);
                else
                    eri_block_regs<LI, LJ, LK, LL, P::PIPE, IB, IE>(eri, a.bra_tab + (size_t)(e - a.j_base) * npij * 8, ket, ri, rj, rk, rl, npij,
                                                            npkl, a.omega, fac, s_rys, kr0, kr1);
#define ERI_(i, j, k, l) eri[((((i) - IB) * NFJ + (j)) * NFK + (k)) * NFL + (l)]
                if constexpr (DO_J) {
                    // J_kl += sum_ij (ij|kl) D[j,i]: lane-stationary
                    R d_ji_l[P::PIPE ? 1 : NFI * NFJ];
                    if constexpr (!P::PIPE) {
#pragma unroll
                        for (int i = IB; i < IE; i++)
#pragma unroll
                        for (int j = 0; j < NFJ; j++) d_ji_l[i * NFJ + j] = ldd((size_t)(j0 + j) * nao + i0 + i);
                    }
                    const R* __restrict__ d_ji_p = P::PIPE ? d_ji : d_ji_l;
#pragma unroll
                    for (int k = 0; k < NFK; k++)
#pragma unroll
                    for (int l = 0; l < NFL; l++) {
                        if constexpr (F32) {
                            R s = R(0);
#pragma unroll
                            for (int i = IB; i < IE; i++)
#pragma unroll
                            for (int j = 0; j < NFJ; j++) s = fma(ERI_(i, j, k, l), d_ji_p[i * NFJ + j], s);
                            jkl[k * NFL + l] += (double)s;
                        } else {
                            R s = (R)jkl[k * NFL + l];
#pragma unroll
                            for (int i = IB; i < IE; i++)
#pragma unroll
                            for (int j = 0; j < NFJ; j++) s = fma(ERI_(i, j, k, l), d_ji_p[i * NFJ + j], s);
                            jkl[k * NFL + l] = s;
                        }
                    }
                    // J_ij += sum_kl (ij|kl) D[l,k]: one address for the whole warp -> reduce-scatter
                    R vij[NIH * NFJ];
#pragma unroll
                    for (int x = 0; x < NIH * NFJ; x++) vij[x] = R(0);
#pragma unroll
                    for (int k = 0; k < NFK; k++)
#pragma unroll
                    for (int l = 0; l < NFL; l++) {
                        const R d = P::DLK_MODE == 0 ? dlk_r[P::DLK_MODE == 0 ? k * NFL + l : 0]
                                  : (P::DLK_MODE == 1 ? DSLOT_(P::S_DLK + k * NFL + l)
                                                      : ldd((size_t)(l0 + l) * nao + k0 + k));
#pragma unroll
                        for (int i = IB; i < IE; i++)
#pragma unroll
                        for (int j = 0; j < NFJ; j++) vij[(i - IB) * NFJ + j] = fma(ERI_(i, j, k, l), d, vij[(i - IB) * NFJ + j]);
                    }
                    int idx = 0, cnt = NIH * NFJ;
                    WarpReduceScatter<NIH * NFJ, 16>::run(vij, lane, idx, cnt);
#pragma unroll
                    for (int x = 0; x < warp_rs_final(NIH * NFJ); x++)
                        if (x < cnt) {
                            const int ii = (idx + x) / NFJ, j = (idx + x) - ii * NFJ, i = IB + ii;
                            atomicAdd(a.vj + (size_t)(j0 + j) * nao + i0 + i, (double)vij[x]);
                        }
                }
                if constexpr (DO_K) {
                    {   // K_ik += sum_jl (ij|kl) D[j,l]: stationary over the j loop
                        R d_l[P::PIPE ? 1 : NFJ * NFL];
                        if constexpr (!P::PIPE) {
#pragma unroll
                            for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) d_l[j * NFL + l] = ldd((size_t)(j0 + j) * nao + l0 + l);
                        }
                        const R* __restrict__ d = P::PIPE ? d_jl : d_l;
#pragma unroll
                        for (int i = IB; i < IE; i++)
#pragma unroll
                        for (int k = 0; k < NFK; k++) {
                            R s = R(0);
#pragma unroll
                            for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) s = fma(ERI_(i, j, k, l), d[j * NFL + l], s);
                            if constexpr (P::ACC_SMEM) ASLOT_(i * NFK + k) += (double)s;
                            else kacc[i * NFK + k] += (double)s;
                        }
                    }
                    {   // K_il += sum_jk (ij|kl) D[j,k]: stationary over the j loop
                        R d_l[P::PIPE ? 1 : NFJ * NFK];
                        if constexpr (!P::PIPE) {
#pragma unroll
                            for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) d_l[j * NFK + k] = ldd((size_t)(j0 + j) * nao + k0 + k);
                        }
                        const R* __restrict__ d = P::PIPE ? d_jk : d_l;
#pragma unroll
                        for (int i = IB; i < IE; i++)
#pragma unroll
                        for (int l = 0; l < NFL; l++) {
                            R s = R(0);
#pragma unroll
                            for (int j = 0; j < NFJ; j++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) s = fma(ERI_(i, j, k, l), d[j * NFK + k], s);
                            if constexpr (P::ACC_SMEM) ASLOT_(NFI * NFK + i * NFL + l) += (double)s;
                            else kacc[NFI * NFK + i * NFL + l] += (double)s;
                        }
                    }
                    {   // K_jk += sum_il (ij|kl) D[i,l]: scattered per quartet
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                        for (int k = 0; k < NFK; k++) {
                            R s = R(0);
#pragma unroll
                            for (int i = IB; i < IE; i++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) {
                                const R d = P::DI_SMEM ? DSLOT_(P::S_DI + NFI * NFK + i * NFL + l)
                                                       : ldd((size_t)(i0 + i) * nao + l0 + l);
                                s = fma(ERI_(i, j, k, l), d, s);
                            }
                            if (live) atomicAdd(a.vk + (size_t)(j0 + j) * nao + k0 + k, (double)s);
                        }
                    }
                    {   // K_jl += sum_ik (ij|kl) D[i,k]: scattered per quartet
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                        for (int l = 0; l < NFL; l++) {
                            R s = R(0);
#pragma unroll
                            for (int i = IB; i < IE; i++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) {
                                const R d = P::DI_SMEM ? DSLOT_(P::S_DI + i * NFK + k)
                                                       : ldd((size_t)(i0 + i) * nao + k0 + k);
                                s = fma(ERI_(i, j, k, l), d, s);
                            }
                            if (live) atomicAdd(a.vk + (size_t)(j0 + j) * nao + l0 + l, (double)s);
                        }
                    }
                }
#undef ERI_
                };
                pass(std::integral_constant<int, 0>{});
                if constexpr (P::ISPLIT > 1) pass(std::integral_constant<int, 1>{});
            }
            // flush the per-i K accumulators of the lanes that contributed
            if constexpr (DO_K) {
                if (touched_i) {
#pragma unroll
                    for (int i = 0; i < NFI; i++)
#pragma unroll
                    for (int k = 0; k < NFK; k++) {
                        const double v = P::ACC_SMEM ? ASLOT_(i * NFK + k) : kacc[P::ACC_SMEM ? 0 : i * NFK + k];
                        atomicAdd(a.vk + (size_t)(i0 + i) * nao + k0 + k, v);
                    }
#pragma unroll
                    for (int i = 0; i < NFI; i++)
#pragma unroll
                    for (int l = 0; l < NFL; l++) {
                        const double v = P::ACC_SMEM ? ASLOT_(NFI * NFK + i * NFL + l)
                                                     : kacc[P::ACC_SMEM ? 0 : NFI * NFK + i * NFL + l];
                        atomicAdd(a.vk + (size_t)(i0 + i) * nao + l0 + l, v);
                    }
                }
            }
            touched_kl |= touched_i;
        }
        if constexpr (DO_J) {
            if (touched_kl) {
#pragma unroll
                for (int k = 0; k < NFK; k++)
#pragma unroll
                for (int l = 0; l < NFL; l++) atomicAdd(a.vj + (size_t)(l0 + l) * nao + k0 + k, jkl[k * NFL + l]);
            }
        }
    }
#undef ASLOT_
#undef DSLOT_
    if (lane == 0 && nq) atomicAdd(a.qcount, nq);
}

}  // namespace jqc

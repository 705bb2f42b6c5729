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
    int rank, world;                          // this GPU takes tasks rank, rank + world, ...
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
template <int LI, int LJ, int LK, int LL>
__device__ __forceinline__ void eri_block_regs(double* __restrict__ eri, const double* __restrict__ bra,
                                               const double* __restrict__ ket, const double4 ri, const double4 rj,
                                               const double4 rk, const double4 rl, const int npij, const int npkl,
                                               const double omega, const double fac, const double2* __restrict__ s_rys)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    constexpr int NFI = S::NFI, NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL, N = S::N;
    constexpr int NROOTS = S::NROOTS, GS = S::GSIZE, DJ = S::DJ, DK = S::DK, DL = S::DL;
    const double rjri[3] = {rj.x - ri.x, rj.y - ri.y, rj.z - ri.z};
    const double rlrk[3] = {rl.x - rk.x, rl.y - rk.y, rl.z - rk.z};
#pragma unroll
    for (int n = 0; n < N; n++) eri[n] = 0.0;
#pragma unroll 1
    for (int klp = 0; klp < npkl; klp++) {
        const double4 k0 = *reinterpret_cast<const double4*>(ket + klp * 8);
        const double4 k1 = *reinterpret_cast<const double4*>(ket + klp * 8 + 4);
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
            const double inv_aijkl = 1.0 / (aij + akl);
            const double theta = aij * akl * inv_aijkl;
            const double gy0 = cicj * inv_aij * inv_akl * sqrt(inv_aijkl);
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

// Per-class layout of the brick kernel, usable at compile time (BrickPlan) and by the host (which
// classes are supported, how much dynamic shared memory a launch needs).
struct BrickShape {
    int n, nki, njkl, nroots;
    bool acc_smem;      // per-i K accumulators in lane-private shared memory (else registers)
    bool di_smem;       // D_il / D_ik blocks (stationary over the j loop) staged in lane-private shared memory
    int dlk_mode;       // D_lk block (stationary over the brick): 0 registers, 1 lane-private shared memory, 2 reloaded
    int slots;          // lane-private doubles per lane
    int regs, minb, nwarps;
    size_t rys_bytes, smem;
    bool fits;
};

__host__ __device__ constexpr BrickShape brick_shape(int li, int lj, int lk, int ll)
{
    BrickShape b{};
    const int nfi = nf_of(li), nfj = nf_of(lj), nfk = nf_of(lk), nfl = nf_of(ll);
    b.n = nfi * nfj * nfk * nfl;
    b.nki = nfi * (nfk + nfl);
    b.njkl = nfk * nfl;
    b.nroots = (li + lj + lk + ll) / 2 + 1;
    b.nwarps = 4;
    const int live = b.n + b.nki + b.njkl;
    b.acc_smem = live > 80 && b.nki > 12;
    b.di_smem = b.nki <= 48;
    b.dlk_mode = b.njkl <= 3 ? 0 : (b.njkl <= 9 ? 1 : 2);
    b.slots = (b.acc_smem ? b.nki : 0) + (b.di_smem ? b.nki : 0) + (b.dlk_mode == 1 ? b.njkl : 0);
    // register budget per thread -> CTAs of 128 threads per SM: 255 -> 2, 168 -> 3, 128 -> 4
    b.regs = live <= 12 ? 128 : (live <= 36 ? 168 : 255);
    b.minb = 65536 / (b.regs * b.nwarps * 32);
    b.rys_bytes = (size_t)b.nroots * (14 + 2 * b.nroots) * (RYS_NCOEF + 1) * 16;
    b.smem = b.rys_bytes + (size_t)b.nwarps * 32 * b.slots * sizeof(double);
    b.fits = b.n <= JQC_SMALL_N && b.smem * b.minb <= 216 * 1024;
    return b;
}

template <int LI, int LJ, int LK, int LL>
struct BrickPlan {
    static constexpr BrickShape B = brick_shape(LI, LJ, LK, LL);
    static constexpr int NKI = B.nki, NJKL = B.njkl, NWARPS = B.nwarps, MINB = B.minb, SLOTS = B.slots;
    static constexpr bool ACC_SMEM = B.acc_smem, DI_SMEM = B.di_smem, FITS = B.fits;
    static constexpr int DLK_MODE = B.dlk_mode;
    static constexpr size_t SMEM = B.smem, RYS_BYTES = B.rys_bytes;
    // slot offsets (lane-private doubles)
    static constexpr int S_ACC = 0, S_DI = ACC_SMEM ? NKI : 0, S_DLK = S_DI + (DI_SMEM ? NKI : 0);
};

template <int LI, int LJ, int LK, int LL, bool DO_J, bool DO_K>
__global__ void __launch_bounds__(BrickPlan<LI, LJ, LK, LL>::NWARPS * 32, BrickPlan<LI, LJ, LK, LL>::MINB)
jk_brick_kernel(const BrickArgs a)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    using P = BrickPlan<LI, LJ, LK, LL>;
    constexpr int NFI = S::NFI, NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL, N = S::N;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double2 brick_smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nao = a.nao, nbas = a.nbas;
    const float log_max = ordered_to_float(*a.log_max_ordered);
    const float dmaxf = fmaxf(log_max, -36.8f);
    const double paircut = log(1e-13) - (double)log_max;       // jk.py:185-187, 412
    const unsigned ntask = (unsigned)a.n_blk * (unsigned)a.n_ichunk * (unsigned)a.jsplit;
    // shared memory: [Rys table of this class][lane-private slots: element-major, lane-minor]
    const double2* __restrict__ s_rys = brick_smem;
    rys_table_to_smem<S::NROOTS>(brick_smem);
    double* __restrict__ slot = reinterpret_cast<double*>(brick_smem) + P::RYS_BYTES / sizeof(double) +
                                (size_t)warp * 32 * P::SLOTS + lane;
    const int npij = a.npi * a.npj, npkl = a.npk * a.npl;
#define SLOT_(x) slot[(x) * 32]
    unsigned long long nq = 0;

#pragma unroll 1
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(a.work, 1u);
        t = __shfl_sync(FULL, t, 0) * (unsigned)a.world + (unsigned)a.rank;
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
        const double* __restrict__ dm = a.dm;
        const double* __restrict__ ket = a.ket_tab + (size_t)pp * npkl * 8;

        double jkl[DO_J ? NFK * NFL : 1];
        double dlk_r[(DO_J && P::DLK_MODE == 0) ? NFK * NFL : 1];
        if constexpr (DO_J) {
#pragma unroll
            for (int k = 0; k < NFK; k++)
#pragma unroll
            for (int l = 0; l < NFL; l++) {
                jkl[k * NFL + l] = 0.0;
                if constexpr (P::DLK_MODE == 0) dlk_r[k * NFL + l] = __ldg(dm + (size_t)(l0 + l) * nao + k0 + k);
                if constexpr (P::DLK_MODE == 1) SLOT_(P::S_DLK + k * NFL + l) = __ldg(dm + (size_t)(l0 + l) * nao + k0 + k);
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
                    if constexpr (P::ACC_SMEM) SLOT_(P::S_ACC + x) = 0.0;
                    else kacc[x] = 0.0;
                }
                if constexpr (P::DI_SMEM) {
#pragma unroll
                    for (int i = 0; i < NFI; i++) {
#pragma unroll
                        for (int k = 0; k < NFK; k++) SLOT_(P::S_DI + i * NFK + k) = __ldg(dm + (size_t)(i0 + i) * nao + k0 + k);
#pragma unroll
                        for (int l = 0; l < NFL; l++) SLOT_(P::S_DI + NFI * NFK + i * NFL + l) = __ldg(dm + (size_t)(i0 + i) * nao + l0 + l);
                    }
                }
            }
            bool touched_i = false;

#pragma unroll 1
            for (; e < e_end; e++) {
                const float q_ij = a.j_q[e];
                if (!(q_ij + Qb + dmaxf > a.cutoff)) break;         // q-descending list: nothing further passes
                if (!((double)a.j_tq[e] > paircut)) continue;      // bra tile pair not active
                const int jsh = a.j_idx[e];
                // canonical order (screen_jk_tasks.cu:202, 225, 239) + Schwarz x density test (:241-261)
                bool live = lane_i && (!a.tri || ksh < ish || lsh <= jsh);
                if (live) {
                    const float q_ijkl = q_ij + q_kl;
                    float d_large = -36.8f;
                    if constexpr (DO_K) {
                        d_large = fmaxf(d_large, d_ik);
                        d_large = fmaxf(d_large, a.logd[(size_t)jsh * nbas + ksh]);
                        d_large = fmaxf(d_large, d_il);
                        d_large = fmaxf(d_large, a.logd[(size_t)jsh * nbas + lsh]);
                    }
                    if constexpr (DO_J) {
                        d_large = fmaxf(d_large, a.logd[(size_t)ish * nbas + jsh]);
                        d_large = fmaxf(d_large, d_kl);
                    }
                    live = q_ijkl + d_large > a.cutoff;
                }
                const unsigned m = __ballot_sync(FULL, live);
                if (m == 0) continue;
                if (lane == 0) nq += __popc(m);
                touched_i |= live;

                const double* __restrict__ bj = a.basis + jsh * BASIS_STRIDE;
                const double4 rj = *reinterpret_cast<const double4*>(bj);
                const int j0 = (int)rj.w;
                double fac = live ? PI_FAC : 0.0;
                if (ish == jsh) fac *= 0.5;
                if (ksh == lsh) fac *= 0.5;
                if (ish == ksh && jsh == lsh) fac *= 0.5;
                double eri[N];
                eri_block_regs<LI, LJ, LK, LL>(eri, a.bra_tab + (size_t)(e - a.j_base) * npij * 8, ket, ri, rj, rk, rl, npij,
                                               npkl, a.omega, fac, s_rys);
#define ERI_(i, j, k, l) eri[(((i) * NFJ + (j)) * NFK + (k)) * NFL + (l)]
                if constexpr (DO_J) {
                    // J_kl += sum_ij (ij|kl) D[j,i]: lane-stationary
                    double d_ji[NFI * NFJ];
#pragma unroll
                    for (int i = 0; i < NFI; i++)
#pragma unroll
                    for (int j = 0; j < NFJ; j++) d_ji[i * NFJ + j] = __ldg(dm + (size_t)(j0 + j) * nao + i0 + i);
#pragma unroll
                    for (int k = 0; k < NFK; k++)
#pragma unroll
                    for (int l = 0; l < NFL; l++) {
                        double s = jkl[k * NFL + l];
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                        for (int j = 0; j < NFJ; j++) s = fma(ERI_(i, j, k, l), d_ji[i * NFJ + j], s);
                        jkl[k * NFL + l] = s;
                    }
                    // J_ij += sum_kl (ij|kl) D[l,k]: one address for the whole warp -> reduce-scatter
                    double vij[NFI * NFJ];
#pragma unroll
                    for (int x = 0; x < NFI * NFJ; x++) vij[x] = 0.0;
#pragma unroll
                    for (int k = 0; k < NFK; k++)
#pragma unroll
                    for (int l = 0; l < NFL; l++) {
                        const double d = P::DLK_MODE == 0 ? dlk_r[P::DLK_MODE == 0 ? k * NFL + l : 0]
                                       : (P::DLK_MODE == 1 ? SLOT_(P::S_DLK + k * NFL + l)
                                                           : __ldg(dm + (size_t)(l0 + l) * nao + k0 + k));
#pragma unroll
                        for (int i = 0; i < NFI; i++)
#pragma unroll
                        for (int j = 0; j < NFJ; j++) vij[i * NFJ + j] = fma(ERI_(i, j, k, l), d, vij[i * NFJ + j]);
                    }
                    int idx = 0, cnt = NFI * NFJ;
                    WarpReduceScatter<NFI * NFJ, 16>::run(vij, lane, idx, cnt);
#pragma unroll
                    for (int x = 0; x < warp_rs_final(NFI * NFJ); x++)
                        if (x < cnt) {
                            const int i = (idx + x) / NFJ, j = (idx + x) - i * NFJ;
                            atomicAdd(a.vj + (size_t)(j0 + j) * nao + i0 + i, vij[x]);
                        }
                }
                if constexpr (DO_K) {
                    {   // K_ik += sum_jl (ij|kl) D[j,l]: stationary over the j loop
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
                            if constexpr (P::ACC_SMEM) SLOT_(P::S_ACC + i * NFK + k) += s;
                            else kacc[i * NFK + k] += s;
                        }
                    }
                    {   // K_il += sum_jk (ij|kl) D[j,k]: stationary over the j loop
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
                            if constexpr (P::ACC_SMEM) SLOT_(P::S_ACC + NFI * NFK + i * NFL + l) += s;
                            else kacc[NFI * NFK + i * NFL + l] += s;
                        }
                    }
                    {   // K_jk += sum_il (ij|kl) D[i,l]: scattered per quartet
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                        for (int k = 0; k < NFK; k++) {
                            double s = 0.0;
#pragma unroll
                            for (int i = 0; i < NFI; i++)
#pragma unroll
                            for (int l = 0; l < NFL; l++) {
                                const double d = P::DI_SMEM ? SLOT_(P::S_DI + NFI * NFK + i * NFL + l)
                                                            : __ldg(dm + (size_t)(i0 + i) * nao + l0 + l);
                                s = fma(ERI_(i, j, k, l), d, s);
                            }
                            if (live) atomicAdd(a.vk + (size_t)(j0 + j) * nao + k0 + k, s);
                        }
                    }
                    {   // K_jl += sum_ik (ij|kl) D[i,k]: scattered per quartet
#pragma unroll
                        for (int j = 0; j < NFJ; j++)
#pragma unroll
                        for (int l = 0; l < NFL; l++) {
                            double s = 0.0;
#pragma unroll
                            for (int i = 0; i < NFI; i++)
#pragma unroll
                            for (int k = 0; k < NFK; k++) {
                                const double d = P::DI_SMEM ? SLOT_(P::S_DI + i * NFK + k)
                                                            : __ldg(dm + (size_t)(i0 + i) * nao + k0 + k);
                                s = fma(ERI_(i, j, k, l), d, s);
                            }
                            if (live) atomicAdd(a.vk + (size_t)(j0 + j) * nao + l0 + l, s);
                        }
                    }
                }
#undef ERI_
            }
            // flush the per-i K accumulators of the lanes that contributed
            if constexpr (DO_K) {
                if (touched_i) {
#pragma unroll
                    for (int i = 0; i < NFI; i++)
#pragma unroll
                    for (int k = 0; k < NFK; k++) {
                        const double v = P::ACC_SMEM ? SLOT_(P::S_ACC + i * NFK + k) : kacc[P::ACC_SMEM ? 0 : i * NFK + k];
                        atomicAdd(a.vk + (size_t)(i0 + i) * nao + k0 + k, v);
                    }
#pragma unroll
                    for (int i = 0; i < NFI; i++)
#pragma unroll
                    for (int l = 0; l < NFL; l++) {
                        const double v = P::ACC_SMEM ? SLOT_(P::S_ACC + NFI * NFK + i * NFL + l)
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
#undef SLOT_
    if (lane == 0 && nq) atomicAdd(a.qcount, nq);
}

}  // namespace jqc

// Multi-lane FP64 Rys J/K kernel for the medium and large angular classes (sm_100a).
//
// Replaces the reference's rys_1qnt_vjk (jqc/backend/jk/1qnt.cu:47-871) with a
// design built around what bounds this path on B200 (profiles/microbench/atomics.cu, round 1):
//   * T lanes of ONE warp cooperate on a quartet, 32/T quartets per warp; the only
//     synchronisation is __syncwarp (the reference uses 3 __syncthreads per root over 256
//     threads and caps shared memory at 48 KB);
//   * the g arrays of ALL roots live in shared memory (B200: 227 KB/CTA), so the TRR/HRR
//     recurrences of every (root, direction) pair run on different lanes at once;
//   * each lane owns a few (k,l) component pairs and keeps the full (i,j) block of those
//     pairs in registers: every g value fetched from shared memory feeds nfi*nfj/((li+1)(lj+1))
//     FMAs, and all six J/K contractions are taken straight from registers;
//   * cross-lane sums go through a small staging area (no shared-memory atomics: FP64 ATOMS is
//     a CAS loop on sm_100), and the flush issues reductions to runs of consecutive
//     addresses, which the L2 retires ~2.5x faster than scattered FP64 atomics.
#pragma once
#include "jk_1q1t.cuh"

namespace jqc {

// conflict-minimising strides per class (generated: tools/gen_warp_layout.py)
#include "jk_warp_layout.h"

// 8-byte asynchronous global->shared copy (LDGSTS): the density blocks of a quartet are fetched
// while the lanes are busy with the recurrences and products, without holding registers.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One Rys root/weight (index i of NROOTS) at argument x.
template <int NROOTS>
__device__ __forceinline__ void rys_root_one(double x, int i, double& root, double& weight)
{
    constexpr int TRI = NROOTS * (NROOTS - 1) / 2;
    constexpr double large_x = NROOTS * 5 + 35;
    if (x >= large_x) {
        const double rs_x = jrsqrt(x);
        root = RYS_LARGEX[(TRI + i) * 2] * (rs_x * rs_x);
        weight = RYS_LARGEX[(TRI + i) * 2 + 1] * (SQRTPIE4 * rs_x);
        return;
    }
    const RysPowers p = rys_powers(x);
    const double2* __restrict__ c = reinterpret_cast<const double2*>(RYS_MONO + RYS_CHEB_OFFSET[NROOTS - 1]) +
                                    ((size_t)p.it * NROOTS + i) * RYS_NCOEF;
    rys_estrin(c, p, root, weight, RysLdg());
}

// Same, from the shared-memory copy of the table (RysSmem layout, jqc_common.cuh).
template <int NROOTS>
__device__ __forceinline__ void rys_root_one_smem(double x, int i, double& root, double& weight,
                                                  const double2* __restrict__ s_tab)
{
    using T = RysSmem<NROOTS>;
    constexpr int TRI = NROOTS * (NROOTS - 1) / 2;
    constexpr double large_x = NROOTS * 5 + 35;
    if (x >= large_x) {
        const double rs_x = jrsqrt(x);
        root = RYS_LARGEX[(TRI + i) * 2] * (rs_x * rs_x);
        weight = RYS_LARGEX[(TRI + i) * 2 + 1] * (SQRTPIE4 * rs_x);
        return;
    }
    const RysPowers p = rys_powers(x);
    rys_estrin(s_tab + (i * T::NINT + p.it) * T::ROW, p, root, weight, RysLds());
}

// TRR + HRR for one cartesian direction with the recurrences in REGISTERS: the TRR runs row by row
// over k (two rows of li+lj+1 values live), every finished row goes through the (i -> j) HRR and
// is stored once; the (k -> l) HRR then reads each (i,j) column back once.  Shared-memory traffic
// is ~1 access per recurrence step instead of 3 for the in-place form below, and the dependent
// chains run at register latency.
template <int LI, int LJ, int LK, int LL, class R>
__device__ __forceinline__ void fill_g_dir_regs(R* __restrict__ gd, const R seed, const R c0, const R cp, const R b10,
                                                const R b01, const R b00, const R ab, const R cd)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    constexpr int DJ = S::DJ, DK = WarpLayout<LI, LJ, LK, LL>::DKP, DL = DK * (LK + 1), LIJ = S::LIJ, LKL = S::LKL;
    R prev[LIJ + 1], cur[LIJ + 1];
#pragma unroll
    for (int k = 0; k <= LKL; k++) {
        // row k of the TRR table g(i, 0 | k, 0)
        R row[LIJ + 1];
        if (k == 0) {
            row[0] = seed;
            if constexpr (LIJ > 0) row[1] = c0 * seed;
#pragma unroll
            for (int i = 1; i < LIJ; i++) row[i + 1] = fma(c0, row[i], (i * b10) * row[i - 1]);
        } else {
#pragma unroll
            for (int i = 0; i <= LIJ; i++) {
                R v = cp * cur[i];
                if (k > 1) v = fma((k - 1) * b01, prev[i], v);
                if (i > 0) v = fma(i * b00, cur[i - 1], v);
                row[i] = v;
            }
        }
#pragma unroll
        for (int i = 0; i <= LIJ; i++) { prev[i] = cur[i]; cur[i] = row[i]; }
        // (i -> j) HRR of this row, levels j = 0 .. LJ; level j keeps i <= LIJ - j
#pragma unroll
        for (int j = 0; j <= LJ; j++) {
#pragma unroll
            for (int i = 0; i <= LI; i++) gd[i + j * DJ + k * DK] = row[i];
            if (j < LJ) {
#pragma unroll
                for (int i = 0; i < LIJ - j; i++) row[i] = fma(-ab, row[i], row[i + 1]);
            }
        }
    }
    if constexpr (LL > 0) {
        // (k -> l) HRR per (i,j) column
#pragma unroll
        for (int j = 0; j <= LJ; j++)
#pragma unroll
        for (int i = 0; i <= LI; i++) {
            const int ij = i + j * DJ;
            R v[LKL + 1];
#pragma unroll
            for (int k = 0; k <= LKL; k++) v[k] = gd[ij + k * DK];
#pragma unroll
            for (int l = 1; l <= LL; l++) {
#pragma unroll
                for (int k = 0; k <= LKL - l; k++) v[k] = fma(-cd, v[k], v[k + 1]);
#pragma unroll
                for (int k = 0; k <= LK; k++) gd[ij + k * DK + l * DL] = v[k];
            }
        }
    }
}

// TRR + in-place HRR for one cartesian direction into gd[GSIZE] (shared memory).
template <int LI, int LJ, int LK, int LL>
__device__ __forceinline__ void fill_g_dir(double* __restrict__ gd, const double seed, const double c0, const double cp,
                                           const double b10, const double b01, const double b00, const double ab,
                                           const double cd)
{
    // shared-memory strides: DK padded to an odd number of doubles so that the (k,l) slots that
    // different lanes read in the product phase fall into different banks; DL keeps the exact
    // (LK+1)*DK aliasing the in-place HRR relies on.
    using S = QuartetShape<LI, LJ, LK, LL>;
    constexpr int DJ = S::DJ, DK = WarpLayout<LI, LJ, LK, LL>::DKP, DL = DK * (LK + 1), LIJ = S::LIJ, LKL = S::LKL;
    gd[0] = seed;
    if constexpr (LIJ > 0) {
        double s0 = seed, s1 = c0 * seed;
        gd[1] = s1;
#pragma unroll
        for (int i = 1; i < LIJ; i++) {
            const double s2 = fma(c0, s1, (i * b10) * s0);
            gd[i + 1] = s2;
            s0 = s1; s1 = s2;
        }
    }
    if constexpr (LKL > 0) {
#pragma unroll
        for (int i = 0; i <= LIJ; i++) {
            double v = cp * gd[i];
            if (i > 0) v = fma(i * b00, gd[i - 1], v);
            gd[i + DK] = v;
        }
#pragma unroll
        for (int k = 1; k < LKL; k++) {
            const double kb01 = k * b01;
#pragma unroll
            for (int i = 0; i <= LIJ; i++) {
                double v = fma(cp, gd[i + k * DK], kb01 * gd[i + (k - 1) * DK]);
                if (i > 0) v = fma(i * b00, gd[i - 1 + k * DK], v);
                gd[i + (k + 1) * DK] = v;
            }
        }
    }
    if constexpr (LJ > 0) {
#pragma unroll
        for (int k = 0; k <= LKL; k++)
#pragma unroll
            for (int j = 0; j < LJ; j++)
#pragma unroll
                for (int i = LIJ - j - 1; i >= 0; i--) {
                    const int src = i + j * DJ + k * DK;
                    gd[src + DJ] = fma(-ab, gd[src], gd[src + 1]);
                }
    }
    if constexpr (LL > 0) {
#pragma unroll
        for (int ij = 0; ij < S::DK; ij++)
#pragma unroll
            for (int l = 0; l < LL; l++)
#pragma unroll
                for (int k = LKL - l - 1; k >= 0; k--) {
                    const int src = ij + k * DK + l * DL;
                    gd[src + DL] = fma(-cd, gd[src], gd[src + DK]);
                }
    }
}

// Lane map for cooperative passes over an NR x NC block by T lanes: a 2-D (rows x columns) tiling
// needs no per-element division but may leave lanes idle; it is used when it loses < 15 % of the
// slots relative to the flattened (e = t + m T) map.
__host__ __device__ constexpr bool use_2d_map(int NR, int NC, int T)
{
    const int CW = NC < T ? NC : T, RT = T / CW;
    const int it2d = ((NR + RT - 1) / RT) * ((NC + CW - 1) / CW);
    const int itflat = (NR * NC + T - 1) / T;
    return it2d * 100 <= itflat * 115;
}

// ---- per-class lane layout ---------------------------------------------------------------
// Per-class tuning of the multi-lane kernel (the counterpart of the reference's per-device fragment
// tables, jqc/backend/data/optimal_scheme_*.json): ACC = accumulator doubles per lane (decides how
// many lanes share a quartet), REGS = register budget (255 -> 8 warps/SM, 168 -> 12 warps/SM).
// jk_warp_tuning.h is generated by tools/tune_warp.py from measured per-class times on B200; a whole
// build can be forced to one variant with -DJQC_WARP_FORCE_ACC=.. -DJQC_WARP_FORCE_REGS=..
template <int LI, int LJ, int LK, int LL>
struct WarpTune {
    static constexpr int ACC = 60, REGS = 255;
};
#if defined(JQC_WARP_FORCE_ACC) && defined(JQC_WARP_FORCE_REGS)
#define JQC_TUNE_ACC(LI, LJ, LK, LL) JQC_WARP_FORCE_ACC
#define JQC_TUNE_REGS(LI, LJ, LK, LL) JQC_WARP_FORCE_REGS
#else
#include "jk_warp_tuning.h"
#define JQC_TUNE_ACC(LI, LJ, LK, LL) (WarpTune<LI, LJ, LK, LL>::ACC)
#define JQC_TUNE_REGS(LI, LJ, LK, LL) (WarpTune<LI, LJ, LK, LL>::REGS)
#endif

template <int LI, int LJ, int LK, int LL>
struct WarpPlan {
    static constexpr int WARP_ACC_MAX = JQC_TUNE_ACC(LI, LJ, LK, LL);
    static constexpr int REGS = JQC_TUNE_REGS(LI, LJ, LK, LL);
    using S = QuartetShape<LI, LJ, LK, LL>;
    static constexpr int NIJ = S::NFI * S::NFJ, NKL = S::NFK * S::NFL;
    // bra passes: split the j components so that one pass keeps <= WARP_ACC_MAX accumulators per pair
    static constexpr int npass()
    {
        for (int p = 1; p <= S::NFJ; p++)
            if (S::NFJ % p == 0 && NIJ / p <= WARP_ACC_MAX) return p;
        return S::NFJ;
    }
    static constexpr int NPASS = npass();
    static constexpr int NJC = S::NFJ / NPASS;          // j components per pass
    static constexpr int PASS_ACC = NJC * S::NFI;
    static constexpr int lanes()
    {
        int best = 32, best_eff = -1;
        for (int t = 1; t <= 32; t++) {
            const int nklp = (NKL + t - 1) / t;
            if (nklp * PASS_ACC > WARP_ACC_MAX) continue;
            // efficiency in 1/1024: (lanes used per warp) * (pair slots used)
            const int eff = ((32 / t) * t * 1024 / 32) * NKL / (t * nklp);
            if (eff > best_eff) { best_eff = eff; best = t; }
        }
        return best;
    }
    static constexpr int T = lanes();
    static constexpr int QPW = 32 / T;
    static constexpr int NKLP = (NKL + T - 1) / T;
    static constexpr int DKP = WarpLayout<LI, LJ, LK, LL>::DKP, DLP = DKP * (LK + 1), GSP = DLP * (LL + 1);
    // stride between the (root, direction) arrays: the generated layout's value modulo 16 (recurrence
    // phase: the lanes of a round access the same offset of different arrays), else odd
    static constexpr int array_stride()
    {
        constexpr int want = WarpLayout<LI, LJ, LK, LL>::IS16;
        if (want < 0) return GSP | 1;
        int v = GSP;
        while (v % 16 != want) v++;
        return v;
    }
    static constexpr int IS = array_stride();
    static constexpr int G_ALL = S::NROOTS * 3 * IS;
    static constexpr int NDBLK = NIJ + NKL + S::NFJ * S::NFL + S::NFJ * S::NFK + S::NFI * S::NFL + S::NFI * S::NFK;
    static constexpr int STAGE = 2 * NKL * (S::NFI + S::NFJ) + T * NIJ;
    // shared-memory doubles per quartet group: [rw][g of all roots][D blocks][staging].  With a
    // single bra pass the staging area aliases g (g is dead once the products are done).
    static constexpr bool ALIAS = (NPASS == 1);
    static constexpr int OFF_G = 2 * S::NROOTS;
    static constexpr int OFF_D = OFF_G + (ALIAS ? (G_ALL > STAGE ? G_ALL : STAGE) : G_ALL);
    static constexpr int OFF_STAGE = ALIAS ? OFF_G : OFF_D + NDBLK;
    static constexpr int END = ALIAS ? OFF_D + NDBLK : OFF_STAGE + STAGE;
    // stride between quartet groups, modulo 16 doubles (= all 32 banks): the value of the generated
    // layout table, which minimises the bank conflicts of the product-phase loads (lane (g, t) reads
    // slot(t) of group g); without an entry, T*IS mod 16 (own-slot accesses conflict-free)
    static constexpr int per_group()
    {
        constexpr int want = WarpLayout<LI, LJ, LK, LL>::PG16 >= 0 ? WarpLayout<LI, LJ, LK, LL>::PG16 : (T * IS) % 16;
        int pg = END;
        while (pg % 16 != want) pg++;
        return pg;
    }
    static constexpr int PER_GROUP = per_group();
};

// cooperative staging of the six density blocks of a quartet into s_d (JQC_COPY = load op)
#define JQC_STAGE(OFF, NR, NC, R0, C0)                                                             \
    if constexpr (!use_2d_map(NR, NC, T)) {                                                        \
        _Pragma("unroll") for (int m = 0; m < ((NR) * (NC) + T - 1) / T; m++) {                    \
            const int e = t + m * T;                                                               \
            if (e < (NR) * (NC)) { const int r = e / (NC), c = e - r * (NC);                       \
                JQC_COPY(s_d + (OFF) + e, dm + (size_t)((R0) + r) * nao + (C0) + c); }                \
        }                                                                                          \
    } else {                                                                                       \
        constexpr int CW = (NC) < T ? (NC) : T, RT = T / CW;                                       \
        const int rr = t / CW, cc = t - rr * CW;                                                   \
        if (rr < RT) {                                                                             \
            _Pragma("unroll") for (int mr = 0; mr < ((NR) + RT - 1) / RT; mr++) {                  \
                const int r = rr + mr * RT;                                                        \
                if (r < (NR)) {                                                                    \
                    const double* __restrict__ src = dm + (size_t)((R0) + r) * nao + (C0) + cc;   \
                    double* __restrict__ dst = s_d + (OFF) + r * (NC) + cc;                        \
                    _Pragma("unroll") for (int mc = 0; mc < ((NC) + CW - 1) / CW; mc++)            \
                        if (cc + mc * CW < (NC)) JQC_COPY(dst + mc * CW, src + mc * CW);              \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
    }
#define JQC_STAGE_ALL                                  \
    JQC_STAGE(D_JI, NFJ, NFI, j0, i0)                  \
    JQC_STAGE(D_LK, NFL, NFK, l0, k0)                  \
    if constexpr (DO_K) {                              \
        JQC_STAGE(D_JL, NFJ, NFL, j0, l0)              \
        JQC_STAGE(D_JK, NFJ, NFK, j0, k0)              \
        JQC_STAGE(D_IL, NFI, NFL, i0, l0)              \
        JQC_STAGE(D_IK, NFI, NFK, i0, k0)              \
    }

template <int LI, int LJ, int LK, int LL, bool DO_J, bool DO_K, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 65536 / (WarpPlan<LI, LJ, LK, LL>::REGS * NWARPS * 32))
jk_warp_kernel(const JKArgs a)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    using P = WarpPlan<LI, LJ, LK, LL>;
    constexpr int NFI = S::NFI, NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL;
    constexpr int NROOTS = S::NROOTS, GS = P::IS, DJ = S::DJ, DK = P::DKP, DL = P::DLP;
    constexpr int T = P::T, QPW = P::QPW, NKLP = P::NKLP, NKL = P::NKL, NPASS = P::NPASS, NJC = P::NJC;
    constexpr int NIJ = P::NIJ;

    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / T, t = lane - grp * T;
    const bool lane_ok = grp < QPW;
    double* __restrict__ sg = smem + (size_t)(warp * QPW + (lane_ok ? grp : 0)) * P::PER_GROUP;
    double* __restrict__ s_rw = sg;
    double* __restrict__ s_g = sg + P::OFF_G;
    double* __restrict__ s_d = sg + P::OFF_D;
    double* __restrict__ s_st = sg + P::OFF_STAGE;
    // D block offsets inside s_d
    constexpr int D_JI = 0, D_LK = D_JI + NIJ, D_JL = D_LK + NKL, D_JK = D_JL + NFJ * NFL, D_IL = D_JK + NFJ * NFK,
                  D_IK = D_IL + NFI * NFL;
    // staging offsets inside s_st
    constexpr int ST_IK = 0, ST_IL = ST_IK + NKL * NFI, ST_JK = ST_IL + NKL * NFI, ST_JL = ST_JK + NKL * NFJ,
                  ST_IJ = ST_JL + NKL * NFJ;

    const unsigned ntasks = *a.ntasks;
    const int nao = a.nao;
    const size_t nao2 = (size_t)nao * nao;
    const unsigned nbatch = (ntasks + QPW - 1) / QPW;

    // this lane's (k,l) component pairs: p = t + s*T, k fastest
    int pk[NKLP], pl[NKLP], pox[NKLP], poy[NKLP], poz[NKLP];
    bool pv[NKLP];
#pragma unroll
    for (int s = 0; s < NKLP; s++) {
        const int p = t + s * T;
        pv[s] = lane_ok && p < NKL;
        const int pp = pv[s] ? p : 0;
        pk[s] = pp % NFK;
        pl[s] = pp / NFK;
        pox[s] = CART_X[LK][pk[s]] * DK + CART_X[LL][pl[s]] * DL;
        poy[s] = CART_Y[LK][pk[s]] * DK + CART_Y[LL][pl[s]] * DL + GS;
        poz[s] = CART_Z[LK][pk[s]] * DK + CART_Z[LL][pl[s]] * DL + 2 * GS;
    }

#pragma unroll 1
    for (unsigned batch = blockIdx.x * NWARPS + warp; batch < nbatch; batch += gridDim.x * NWARPS) {
        const unsigned task = batch * QPW + grp;
        const bool active = lane_ok && task < ntasks;
        const ushort4 sq = active ? a.quartets[task] : make_ushort4(0, 0, 0, 0);
        const int ish = sq.x, jsh = sq.y, ksh = sq.z, lsh = sq.w;
        const double* __restrict__ bi = a.basis + ish * BASIS_STRIDE;
        const double* __restrict__ bj = a.basis + jsh * BASIS_STRIDE;
        const double* __restrict__ bk = a.basis + ksh * BASIS_STRIDE;
        const double* __restrict__ bl = a.basis + lsh * BASIS_STRIDE;
        const double4 ri = *reinterpret_cast<const double4*>(bi);
        const double4 rj = *reinterpret_cast<const double4*>(bj);
        const double4 rk = *reinterpret_cast<const double4*>(bk);
        const double4 rl = *reinterpret_cast<const double4*>(bl);
        double fac = PI_FAC;
        if (ish == jsh) fac *= 0.5;
        if (ksh == lsh) fac *= 0.5;
        if (ish == ksh && jsh == lsh) fac *= 0.5;
        const double rjri[3] = {rj.x - ri.x, rj.y - ri.y, rj.z - ri.z};
        const double rlrk[3] = {rl.x - rk.x, rl.y - rk.y, rl.z - rk.z};
        const double rr_ij = rjri[0] * rjri[0] + rjri[1] * rjri[1] + rjri[2] * rjri[2];
        const double rr_kl = rlrk[0] * rlrk[0] + rlrk[1] * rlrk[1] + rlrk[2] * rlrk[2];
        const int i0 = (int)ri.w, j0 = (int)rj.w, k0 = (int)rk.w, l0 = (int)rl.w;
        {   // prefetch the density blocks of matrix 0 (consumed after the products)
            const double* __restrict__ dm = a.dm;
            if (active) {
#define JQC_COPY(dst, src) cp_async8(dst, src)
                JQC_STAGE_ALL
#undef JQC_COPY
            }
            cp_async_commit();
        }

        // With several bra passes and one density matrix the partial sums are carried across the
        // passes (jkl in registers, i-vectors in the staging area) and flushed once; with several
        // density matrices every pass flushes its own share.
        const bool per_pass_flush = (NPASS > 1) && (a.n_dm > 1);
        double jkl[NKLP];
#pragma unroll
        for (int s = 0; s < NKLP; s++) jkl[s] = 0.0;

#pragma unroll
        for (int pass = 0; pass < NPASS; pass++) {
            const int jc0 = pass * NJC;
            double acc[NKLP][NJC * NFI];
#pragma unroll
            for (int s = 0; s < NKLP; s++)
#pragma unroll
                for (int e = 0; e < NJC * NFI; e++) acc[s][e] = 0.0;

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
                    const double rs_aijkl = jrsqrt(aij + akl);
                    const double inv_aijkl = rs_aijkl * rs_aijkl;
                    const double theta = aij * akl * inv_aijkl;
                    const double gy0 = cicj * inv_aij * inv_akl * rs_aijkl;
                    double theta_fac = 1.0, sqrt_theta_fac = 1.0;
                    if (a.omega > 0.0) {
                        const double o2 = a.omega * a.omega;
                        theta_fac = o2 / (o2 + theta);
                        sqrt_theta_fac = sqrt(theta_fac);
                    }
                    const double x = rr * theta * theta_fac;
                    __syncwarp();   // previous product phase has finished reading g
                    // with one primitive quartet the g arrays of pass 0 are still valid in later bra
                    // passes (nothing overwrites them): skip the roots and recurrences there
                    const bool reuse_g = (NPASS > 1) && pass > 0 && (a.npi * a.npj * a.npk * a.npl == 1);
                    // roots: lane t computes roots t, t+T, ...
                    if (active && !reuse_g) {
#pragma unroll 1
                        for (int r = t; r < NROOTS; r += T) {
                            double rt, wt;
                            rys_root_one<NROOTS>(x, r, rt, wt);
                            s_rw[2 * r] = rt * theta_fac;
                            s_rw[2 * r + 1] = wt * sqrt_theta_fac;
                        }
                    }
                    __syncwarp();
                    // recurrences: item = (root, direction)
                    if (active && !reuse_g) {
#pragma unroll 1
                        for (int item = t; item < 3 * NROOTS; item += T) {
                            const int r = item / 3, d = item - 3 * r;
                            const double rt = s_rw[2 * r], wt = s_rw[2 * r + 1];
                            const double rt_aa = rt * inv_aijkl;
                            const double rt_aij = rt_aa * akl, rt_akl = rt_aa * aij;
                            const double b10 = 0.5 * inv_aij * (1.0 - rt_aij);
                            const double b01 = 0.5 * inv_akl * (1.0 - rt_akl);
                            const double b00 = 0.5 * rt_aa;
                            const double ab = d == 0 ? rjri[0] : (d == 1 ? rjri[1] : rjri[2]);
                            const double cd = d == 0 ? rlrk[0] : (d == 1 ? rlrk[1] : rlrk[2]);
                            const double pq = d == 0 ? Rpq[0] : (d == 1 ? Rpq[1] : Rpq[2]);
                            const double seed = d == 0 ? ckcl : (d == 1 ? gy0 : wt);
                            const double c0 = fma(ab, aj_aij, -rt_aij * pq);
                            const double cp = fma(cd, al_akl, rt_akl * pq);
#ifdef JQC_WARP_SMEM_RECURRENCE
                            fill_g_dir<LI, LJ, LK, LL>(s_g + (size_t)item * GS, seed, c0, cp, b10, b01, b00, ab, cd);
#else
                            fill_g_dir_regs<LI, LJ, LK, LL>(s_g + (size_t)item * GS, seed, c0, cp, b10, b01, b00, ab, cd);
#endif
                        }
                    }
                    __syncwarp();
                    // product: this lane's pairs x (i, j in pass) over all roots
                    if (active) {
#pragma unroll
                        for (int s = 0; s < NKLP; s++) {
                            if (!pv[s]) continue;
#pragma unroll 1
                            for (int r = 0; r < NROOTS; r++) {
                                const double* __restrict__ g = s_g + r * 3 * GS;
                                const double* __restrict__ gx = g + pox[s];
                                const double* __restrict__ gy = g + poy[s];
                                const double* __restrict__ gz = g + poz[s];
#pragma unroll
                                for (int jj = 0; jj < NJC; jj++)
#pragma unroll
                                    for (int i = 0; i < NFI; i++) {
                                        const int j = jc0 + jj;
                                        const int ox = CART_X[LI][i] + CART_X[LJ][j] * DJ;
                                        const int oy = CART_Y[LI][i] + CART_Y[LJ][j] * DJ;
                                        const int oz = CART_Z[LI][i] + CART_Z[LJ][j] * DJ;
                                        acc[s][jj * NFI + i] = fma(gx[ox] * gy[oy], gz[oz], acc[s][jj * NFI + i]);
                                    }
                            }
                        }
                    }
                }
            }
            // ---- digestion of this pass from registers (first density matrix; see below for n_dm > 1)
            __syncwarp();
#pragma unroll 1
            for (int b = 0; b < a.n_dm; b++) {
                const double* __restrict__ dm = a.dm + b * nao2;
                // density blocks: the first matrix was prefetched with cp.async at the top of the batch
                // (first pass); later matrices / passes are staged here synchronously
                if (b == 0 && pass == 0) {
                    cp_async_wait_all();
                } else if (active && a.n_dm > 1) {   // one matrix: s_d still holds it from pass 0
#define JQC_COPY(dst, src) *(dst) = __ldg(src)
                    JQC_STAGE_ALL
#undef JQC_COPY
                }
                __syncwarp();
                if (active) {
                    double dlk[NKLP];
#pragma unroll
                    for (int s = 0; s < NKLP; s++) dlk[s] = pv[s] ? s_d[D_LK + pl[s] * NFK + pk[s]] : 0.0;
#pragma unroll
                    for (int s = 0; s < NKLP; s++) {
                        if (!pv[s]) continue;
                        const int kc = pk[s], lc = pl[s], p = lc * NFK + kc;
                        if constexpr (DO_J) {
                            // J_kl[kc,lc] += sum_ij (ij|kl) D[j,i]   (lane-local, final)
                            double sj = 0.0;
#pragma unroll
                            for (int jj = 0; jj < NJC; jj++)
#pragma unroll
                                for (int i = 0; i < NFI; i++) sj = fma(acc[s][jj * NFI + i], s_d[D_JI + (jc0 + jj) * NFI + i], sj);
                            if (NPASS == 1 || per_pass_flush) {
                                atomicAdd(a.vj + b * nao2 + (size_t)(l0 + lc) * nao + k0 + kc, sj);
                            } else {
                                jkl[s] += sj;
                            }
                        }
                        if constexpr (DO_K) {
                            // density columns this pair needs, fetched once (the staging stores below
                            // would otherwise force the compiler to reload them for every element)
                            double djl[NJC], djk[NJC], dil[NFI], dik[NFI];
#pragma unroll
                            for (int jj = 0; jj < NJC; jj++) {
                                djl[jj] = s_d[D_JL + (jc0 + jj) * NFL + lc];
                                djk[jj] = s_d[D_JK + (jc0 + jj) * NFK + kc];
                            }
#pragma unroll
                            for (int i = 0; i < NFI; i++) {
                                dil[i] = s_d[D_IL + i * NFL + lc];
                                dik[i] = s_d[D_IK + i * NFK + kc];
                            }
                            // K_ik partial: a[i] = sum_j (ij|kl) D[j,l];  K_il partial: b[i] = sum_j (ij|kl) D[j,k]
#pragma unroll
                            for (int i = 0; i < NFI; i++) {
                                double va = 0.0, vb = 0.0;
#pragma unroll
                                for (int jj = 0; jj < NJC; jj++) {
                                    va = fma(acc[s][jj * NFI + i], djl[jj], va);
                                    vb = fma(acc[s][jj * NFI + i], djk[jj], vb);
                                }
                                if (pass == 0 || per_pass_flush) { s_st[ST_IK + p * NFI + i] = va; s_st[ST_IL + p * NFI + i] = vb; }
                                else { s_st[ST_IK + p * NFI + i] += va; s_st[ST_IL + p * NFI + i] += vb; }
                            }
                            // K_jk partial: c[j] = sum_i (ij|kl) D[i,l];  K_jl partial: d[j] = sum_i (ij|kl) D[i,k]
#pragma unroll
                            for (int jj = 0; jj < NJC; jj++) {
                                double vc = 0.0, vd = 0.0;
#pragma unroll
                                for (int i = 0; i < NFI; i++) {
                                    vc = fma(acc[s][jj * NFI + i], dil[i], vc);
                                    vd = fma(acc[s][jj * NFI + i], dik[i], vd);
                                }
                                s_st[ST_JK + p * NFJ + jc0 + jj] = vc;
                                s_st[ST_JL + p * NFJ + jc0 + jj] = vd;
                            }
                        }
                    }
                    if constexpr (DO_J) {
                        // J_ij partial of this lane: sum over its pairs of (ij|kl) D[l,k]
#pragma unroll
                        for (int jj = 0; jj < NJC; jj++)
#pragma unroll
                            for (int i = 0; i < NFI; i++) {
                                double v = 0.0;
#pragma unroll
                                for (int s = 0; s < NKLP; s++) v = fma(acc[s][jj * NFI + i], dlk[s], v);
                                s_st[ST_IJ + t * NIJ + (jc0 + jj) * NFI + i] = v;
                            }
                    }
                }
                // flush when the last pass has been staged (for n_dm > 1 every pass flushes its share)
                if (pass == NPASS - 1 || per_pass_flush) {
                    __syncwarp();
                    if (active) {
                        const bool partial_j = per_pass_flush;   // then only this pass' j range is valid
                        const int jlo = partial_j ? jc0 : 0, jhi = partial_j ? jc0 + NJC : NFJ;
// lanes of the group tile a block as (RT rows) x (CW columns); rows/columns advance by compile-time
// steps so that every address is one base plus immediates (no per-element division)
#define JQC_FLUSH(NR, NC, RLO, RHI, EXPR, DEST)                                                    \
    if constexpr (!use_2d_map(NR, NC, T)) {                                                        \
        _Pragma("unroll") for (int m = 0; m < ((NR) * (NC) + T - 1) / T; m++) {                    \
            const int e = t + m * T;                                                               \
            const int r = e / (NC), c = e - r * (NC);                                              \
            if (e < (NR) * (NC) && r >= (RLO) && r < (RHI)) { double v = 0.0; EXPR; atomicAdd(DEST, v); } \
        }                                                                                          \
    } else {                                                                                       \
        constexpr int CW = (NC) < T ? (NC) : T, RT = T / CW;                                       \
        const int rr = t / CW, cc = t - rr * CW;                                                   \
        if (rr < RT) {                                                                             \
            _Pragma("unroll") for (int mr = 0; mr < ((NR) + RT - 1) / RT; mr++) {                  \
                const int r = rr + mr * RT;                                                        \
                if (r >= (RLO) && r < (RHI)) {                                                     \
                    _Pragma("unroll") for (int mc = 0; mc < ((NC) + CW - 1) / CW; mc++) {          \
                        const int c = cc + mc * CW;                                                \
                        if (c < (NC)) { double v = 0.0; EXPR; atomicAdd(DEST, v); }                \
                    }                                                                              \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
    }
                        if constexpr (DO_J) {
                            double* __restrict__ vj = a.vj + b * nao2;
                            JQC_FLUSH(NFJ, NFI, jlo, jhi,
                                      _Pragma("unroll") for (int u = 0; u < T; u++) v += s_st[ST_IJ + u * NIJ + r * NFI + c],
                                      vj + (size_t)(j0 + r) * nao + i0 + c)
                            if (NPASS > 1 && !per_pass_flush) {
#pragma unroll
                                for (int s = 0; s < NKLP; s++)
                                    if (pv[s]) atomicAdd(vj + (size_t)(l0 + pl[s]) * nao + k0 + pk[s], jkl[s]);
                            }
                        }
                        if constexpr (DO_K) {
                            double* __restrict__ vk = a.vk + b * nao2;
                            JQC_FLUSH(NFI, NFK, 0, NFI,
                                      _Pragma("unroll") for (int l = 0; l < NFL; l++) v += s_st[ST_IK + (l * NFK + c) * NFI + r],
                                      vk + (size_t)(i0 + r) * nao + k0 + c)
                            JQC_FLUSH(NFI, NFL, 0, NFI,
                                      _Pragma("unroll") for (int k = 0; k < NFK; k++) v += s_st[ST_IL + (c * NFK + k) * NFI + r],
                                      vk + (size_t)(i0 + r) * nao + l0 + c)
                            JQC_FLUSH(NFJ, NFK, jlo, jhi,
                                      _Pragma("unroll") for (int l = 0; l < NFL; l++) v += s_st[ST_JK + (l * NFK + c) * NFJ + r],
                                      vk + (size_t)(j0 + r) * nao + k0 + c)
                            JQC_FLUSH(NFJ, NFL, jlo, jhi,
                                      _Pragma("unroll") for (int k = 0; k < NFK; k++) v += s_st[ST_JL + (c * NFK + k) * NFJ + r],
                                      vk + (size_t)(j0 + r) * nao + l0 + c)
                        }
#undef JQC_FLUSH
                    }
                    __syncwarp();
                }
            }
        }
    }
}

}  // namespace jqc

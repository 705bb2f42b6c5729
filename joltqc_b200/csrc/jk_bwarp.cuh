// Brick-scheduled multi-lane FP64 Rys J/K kernel for the angular classes whose integral block
// does not fit one thread (> JQC_SMALL_N integrals).
//
// Same arithmetic core as jk_warp.cuh (T lanes of a warp share one shell quartet; g arrays of all
// roots in shared memory; each lane owns a few (k,l) component pairs and keeps the (i,j) block of
// those pairs in registers; replaces rys_1qnt_vjk, jqc/backend/jk/1qnt.cu:47-871) combined with the
// stationary-output schedule of jk_brick.cuh (replaces screen_jk_tasks.cu:75-340 as well):
//   * a lane GROUP owns one (k,l) shell pair of a q-ordered pair list for a whole task; the 32/T
//     groups of a warp walk the same (i, j) sequence, screening each quartet in the loop;
//   * stationary per task: D_lk and the J_kl sums (registers); stationary per i: the D_il / D_ik
//     blocks (staged once) and the K_ik / K_il partial sums (lane-private shared memory, flushed
//     once per i).  Per quartet only D_ji (one copy per warp), D_jl and D_jk are staged, K_jk / K_jl
//     are combined over the group's lanes and scattered, and J_ij - the same addresses for every
//     group of the warp - is summed with a reduce-scatter butterfly over all 32 lanes.
// jk_warp.cuh stages six density blocks and flushes six result blocks per quartet through shared
// memory, which is what bounds it (profiles/r2: shared-memory wavefronts at 77 % of peak).
#pragma once
#include "jk_brick.cuh"
#include "jk_warp.cuh"

namespace jqc {

template <class R, int LI, int LJ, int LK, int LL>
struct BWarpPlan {
    using S = QuartetShape<LI, LJ, LK, LL>;
    static constexpr bool F32 = sizeof(R) == 4;
    using W = WarpPlan<LI, LJ, LK, LL>;
    static constexpr int T = W::T, QPW = W::QPW, NKLP = W::NKLP, NKL = W::NKL, NIJ = W::NIJ, NPASS = W::NPASS, NJC = W::NJC;
    // FP32-band variant: accumulators, g arrays and density blocks are float (half the registers and
    // shared memory per group), which buys residency: 168 registers -> 12 warps per SM
    static constexpr int PASS_ACC = W::PASS_ACC, REGS = F32 ? (W::REGS > 168 ? 168 : W::REGS) : W::REGS;
    static constexpr int IS = W::IS, G_ALL = W::G_ALL;
    // per-group shared memory (doubles).  With a single bra pass the K_jk / K_jl staging area
    // aliases the g arrays (dead once the products of the quartet are done).
    static constexpr bool ALIAS = (NPASS == 1);
    static constexpr int STAGE = 2 * NKL * S::NFJ;
    static constexpr int OFF_RW = 0;
    static constexpr int OFF_G = 2 * S::NROOTS;
    static constexpr int OFF_DJL = OFF_G + (ALIAS ? (G_ALL > STAGE ? G_ALL : STAGE) : G_ALL);
    static constexpr int OFF_DJK = OFF_DJL + S::NFJ * S::NFL;
    static constexpr int OFF_DIL = OFF_DJK + S::NFJ * S::NFK;
    static constexpr int OFF_DIK = OFF_DIL + S::NFI * S::NFL;
    static constexpr int OFF_KP = OFF_DIK + S::NFI * S::NFK;         // lane-private K_ik / K_il partials: [which][s][i][t]
    static constexpr int OFF_STJK = ALIAS ? OFF_G : OFF_KP + 2 * NKLP * S::NFI * T;       // [pair][j]
    static constexpr int OFF_STJL = OFF_STJK + NKL * S::NFJ;
    static constexpr int END = ALIAS ? OFF_KP + 2 * NKLP * S::NFI * T : OFF_STJL + NKL * S::NFJ;
    static constexpr int per_group()
    {
        constexpr int want = WarpLayout<LI, LJ, LK, LL>::PG16 >= 0 ? WarpLayout<LI, LJ, LK, LL>::PG16 : (T * IS) % 16;
        int pg = END;
        while (pg % 16 != want) pg++;
        return pg;
    }
    static constexpr int PER_GROUP = per_group();
    // per-warp shared memory: D_ji + the groups
    static constexpr int OFF_DJI = 0;
    static constexpr int OFF_GROUPS = (OFF_DJI + NIJ + 15) / 16 * 16;
    static constexpr size_t WARP_DOUBLES = ((size_t)OFF_GROUPS + (size_t)QPW * PER_GROUP + 3) / 4 * 4;   // elements of R
    static constexpr size_t WARP_BYTES = WARP_DOUBLES * sizeof(R);
    static constexpr int nwarps()
    {
        int n = (int)((F32 ? 36 : 56) * 1024 / WARP_BYTES);
        return n > 4 ? 4 : (n < 1 ? 1 : n);
    }
    static constexpr int NWARPS = nwarps();
    static constexpr size_t SMEM = WARP_BYTES * NWARPS;
    static constexpr bool FITS = SMEM <= 200 * 1024;
};

// cooperative copy of an NR x NC block of the density (row R0.., column C0..) by the T lanes of a group
#define JQC_BW_STAGE(DST, NR, NC, R0, C0)                                                          \
    {                                                                                              \
        _Pragma("unroll") for (int m = 0; m < ((NR) * (NC) + T - 1) / T; m++) {                    \
            const int e_ = t + m * T;                                                              \
            if (e_ < (NR) * (NC)) { const int r_ = e_ / (NC), c_ = e_ - r_ * (NC);                 \
                (DST)[e_] = ldd((size_t)((R0) + r_) * nao + (C0) + c_); }                   \
        }                                                                                          \
    }

template <class R, int LI, int LJ, int LK, int LL, bool DO_J, bool DO_K, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 65536 / (BWarpPlan<R, LI, LJ, LK, LL>::REGS * NWARPS * 32))
jk_bwarp_kernel(const BrickArgs a)
{
    using S = QuartetShape<LI, LJ, LK, LL>;
    using P = BWarpPlan<R, LI, LJ, LK, LL>;
    constexpr bool F32 = P::F32;
    constexpr int NFI = S::NFI, NFJ = S::NFJ, NFK = S::NFK, NFL = S::NFL;
    constexpr int NROOTS = S::NROOTS, GS = P::IS, DJ = S::DJ, DK = P::W::DKP, DL = P::W::DLP;
    constexpr int T = P::T, QPW = P::QPW, NKLP = P::NKLP, NKL = P::NKL, NPASS = P::NPASS, NJC = P::NJC, NIJ = P::NIJ;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ double smem_raw[];
    R* __restrict__ smem = reinterpret_cast<R*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / T, t = lane - grp * T;
    const bool lane_ok = grp < QPW;
    R* __restrict__ sw = smem + (size_t)warp * P::WARP_DOUBLES;
    R* __restrict__ s_dji = sw + P::OFF_DJI;
    R* __restrict__ sg = sw + P::OFF_GROUPS + (size_t)(lane_ok ? grp : 0) * P::PER_GROUP;
    R* __restrict__ s_rw = sg + P::OFF_RW;
    R* __restrict__ s_g = sg + P::OFF_G;
    R* __restrict__ s_djl = sg + P::OFF_DJL;
    R* __restrict__ s_djk = sg + P::OFF_DJK;
    R* __restrict__ s_dil = sg + P::OFF_DIL;
    R* __restrict__ s_dik = sg + P::OFF_DIK;
    R* __restrict__ s_stjk = sg + P::OFF_STJK;
    R* __restrict__ s_stjl = sg + P::OFF_STJL;
    R* __restrict__ s_kp = sg + P::OFF_KP + t;      // element x of this lane at s_kp[x * T]
#define KP_IK(s, i) s_kp[((s) * NFI + (i)) * T]
#define KP_IL(s, i) s_kp[((NKLP + (s)) * NFI + (i)) * T]

    const int nao = a.nao, nbas = a.nbas;
    // density elements in the precision of this launch
    auto ldd = [&](size_t off) -> R {
        if constexpr (F32) return __ldg(a.dm32 + off);
        else return __ldg(a.dm + off);
    };
    const float log_max = ordered_to_float(*a.log_max_ordered);
    const float dmaxf = fmaxf(log_max, -36.8f);
    const double paircut = log(1e-13) - (double)log_max;
    const int npij = a.npi * a.npj, npkl = a.npk * a.npl;
    const bool single_prim = npij * npkl == 1;
    const unsigned ntask = (unsigned)a.n_blk * (unsigned)a.n_ichunk * (unsigned)a.jsplit;
    unsigned long long nq = 0;

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
    for (;;) {
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(a.work, 1u);
        tk = shard_entry(__shfl_sync(FULL, tk, 0), a.rank, a.world);
        if (tk >= ntask) break;
        const int js = (int)(tk % (unsigned)a.jsplit);
        tk /= (unsigned)a.jsplit;
        const int blk = (int)(tk % (unsigned)a.n_blk), ic = (int)(tk / (unsigned)a.n_blk);
        const int p = blk * QPW + grp;
        const bool pvalid = lane_ok && p < a.n_kl;
        const int pp = pvalid ? p : blk * QPW;
        const ushort2 kl = a.kl[pp];
        const float q_kl = a.kl_q[pp];
        const bool group_on = pvalid && ((double)a.kl_tq[pp] > paircut);
        float Qb = group_on ? q_kl : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Qb = fmaxf(Qb, __shfl_xor_sync(FULL, Qb, o));
        if (!(a.qmax_ij + Qb + dmaxf > a.cutoff)) continue;
        const int ksh = kl.x, lsh = kl.y;
        int kmin = group_on ? ksh : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kmin = min(kmin, __shfl_xor_sync(FULL, kmin, o));
        const double* __restrict__ bk = a.basis + ksh * BASIS_STRIDE;
        const double* __restrict__ bl = a.basis + lsh * BASIS_STRIDE;
        const double4 rk = *reinterpret_cast<const double4*>(bk);
        const double4 rl = *reinterpret_cast<const double4*>(bl);
        const int k0 = (int)rk.w, l0 = (int)rl.w;
        const R rlrk[3] = {(R)(rl.x - rk.x), (R)(rl.y - rk.y), (R)(rl.z - rk.z)};
        const float d_kl = a.logd[(size_t)ksh * nbas + lsh];
        const double* __restrict__ ket = a.ket_tab + (size_t)pp * npkl * 8;

        double jkl[NKLP];
        R dlk[NKLP];
#pragma unroll
        for (int s = 0; s < NKLP; s++) {
            jkl[s] = 0.0;
            dlk[s] = pv[s] ? ldd((size_t)(l0 + pl[s]) * nao + k0 + pk[s]) : R(0);
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
            const bool group_i = group_on && (!a.tri || ksh <= ish);
            // per-i stationary data of the group: D_il, D_ik blocks; K_ik / K_il partial sums
            __syncwarp();
            if (lane_ok) {
                if constexpr (DO_K) {
                    JQC_BW_STAGE(s_dil, NFI, NFL, i0, l0)
                    JQC_BW_STAGE(s_dik, NFI, NFK, i0, k0)
#pragma unroll
                    for (int x = 0; x < 2 * NKLP * NFI; x++) s_kp[x * T] = R(0);
                }
            }
            __syncwarp();
            bool touched_i = false;

#pragma unroll 1
            for (; e < e_end; e++) {
                const float q_ij = a.j_q[e];
                if (!(q_ij + Qb + dmaxf > a.cutoff)) break;
                if (!((double)a.j_tq[e] > paircut)) continue;
                const int jsh = a.j_idx[e];
                bool live = group_i && (!a.tri || ksh < ish || lsh <= jsh);
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
                    const float dq = q_ijkl + d_large;
                    live = dq > a.cutoff && !(dq > a.cutoff_hi);     // precision band of this launch
                }
                const unsigned m = __ballot_sync(FULL, live && t == 0);
                if (m == 0) continue;
                if (lane == 0) nq += __popc(m);
                touched_i |= live;

                const double* __restrict__ bj = a.basis + jsh * BASIS_STRIDE;
                const double4 rj = *reinterpret_cast<const double4*>(bj);
                const int j0 = (int)rj.w;
                R fac = live ? R(PI_FAC) : R(0);
                if (ish == jsh) fac *= R(0.5);
                if (ksh == lsh) fac *= R(0.5);
                if (ish == ksh && jsh == lsh) fac *= R(0.5);
                const R rjri[3] = {(R)(rj.x - ri.x), (R)(rj.y - ri.y), (R)(rj.z - ri.z)};

                // per-step staging: D_ji (one copy per warp), D_jl / D_jk (per group)
                const double* __restrict__ bra = a.bra_tab + (size_t)(e - a.j_base) * npij * 8;
                __syncwarp();                                   // readers of the previous step are done
#pragma unroll
                for (int m2 = 0; m2 < (NIJ + 31) / 32; m2++) {
                    const int x = lane + m2 * 32;
                    if (x < NIJ) { const int jj = x / NFI, ii = x - jj * NFI; s_dji[x] = ldd((size_t)(j0 + jj) * nao + i0 + ii); }
                }
                if (lane_ok) {
                    if constexpr (DO_K) {
                        JQC_BW_STAGE(s_djl, NFJ, NFL, j0, l0)
                        JQC_BW_STAGE(s_djk, NFJ, NFK, j0, k0)
                    }
                }
                // (visibility: the first __syncwarp of the primitive loop below)

#pragma unroll
                for (int pass = 0; pass < NPASS; pass++) {
                    const int jc0 = pass * NJC;
                    R acc[NKLP][NJC * NFI];
#pragma unroll
                    for (int s = 0; s < NKLP; s++)
#pragma unroll
                        for (int x = 0; x < NJC * NFI; x++) acc[s][x] = R(0);

#pragma unroll 1
                    for (int klp = 0; klp < npkl; klp++) {
                        const double4 kt0 = *reinterpret_cast<const double4*>(ket + klp * 8);
                        const double4 kt1 = *reinterpret_cast<const double4*>(ket + klp * 8 + 4);
                        const R akl = (R)kt0.x, inv_akl = (R)kt0.y, al_akl = (R)kt0.z, ckcl = (R)kt0.w;
                        const double qx = kt1.x, qy = kt1.y, qz = kt1.z;
#pragma unroll 1
                        for (int ipj = 0; ipj < npij; ipj++) {
                            __syncwarp();   // staging visible; previous product phase has finished reading g
                            const double4 b0 = *reinterpret_cast<const double4*>(bra + ipj * 8);
                            const double4 b1 = *reinterpret_cast<const double4*>(bra + ipj * 8 + 4);
                            const R aij = (R)b0.x, inv_aij = (R)b0.y, aj_aij = (R)b0.z;
                            const R cicj = fac * (R)b0.w;
                            const R Rpq[3] = {(R)(b1.x - qx), (R)(b1.y - qy), (R)(b1.z - qz)};
                            const R rr = Rpq[0] * Rpq[0] + Rpq[1] * Rpq[1] + Rpq[2] * Rpq[2];
                            const R rs_aijkl = jrsqrt(aij + akl);
                            const R inv_aijkl = rs_aijkl * rs_aijkl;
                            const R theta = aij * akl * inv_aijkl;
                            const R gy0 = cicj * inv_aij * inv_akl * rs_aijkl;
                            R theta_fac = R(1), sqrt_theta_fac = R(1);
                            if (a.omega > 0.0) {
                                const R o2 = (R)(a.omega * a.omega);
                                theta_fac = o2 / (o2 + theta);
                                sqrt_theta_fac = sqrt(theta_fac);
                            }
                            const R x = rr * theta * theta_fac;
                            // with one primitive quartet the g arrays of pass 0 stay valid in later passes
                            const bool reuse_g = (NPASS > 1) && pass > 0 && single_prim;
                            if (lane_ok && !reuse_g) {
#pragma unroll 1
                                for (int r = t; r < NROOTS; r += T) {
                                    double rt, wt;      // (the root finder stays FP64: 2 series per lane, off the hot loop)
                                    rys_root_one<NROOTS>((double)x, r, rt, wt);
                                    s_rw[2 * r] = (R)rt * theta_fac;
                                    s_rw[2 * r + 1] = (R)wt * sqrt_theta_fac;
                                }
                            }
                            __syncwarp();
                            if (lane_ok && !reuse_g) {
#pragma unroll 1
                                for (int item = t; item < 3 * NROOTS; item += T) {
                                    const int r = item / 3, d = item - 3 * r;
                                    const R rt = s_rw[2 * r], wt = s_rw[2 * r + 1];
                                    const R rt_aa = rt * inv_aijkl;
                                    const R rt_aij = rt_aa * akl, rt_akl = rt_aa * aij;
                                    const R b10 = R(0.5) * inv_aij * (R(1) - rt_aij);
                                    const R b01 = R(0.5) * inv_akl * (R(1) - rt_akl);
                                    const R b00 = R(0.5) * rt_aa;
                                    const R ab = d == 0 ? rjri[0] : (d == 1 ? rjri[1] : rjri[2]);
                                    const R cd = d == 0 ? rlrk[0] : (d == 1 ? rlrk[1] : rlrk[2]);
                                    const R pq = d == 0 ? Rpq[0] : (d == 1 ? Rpq[1] : Rpq[2]);
                                    const R seed = d == 0 ? ckcl : (d == 1 ? gy0 : wt);
                                    const R c0 = fma(ab, aj_aij, -rt_aij * pq);
                                    const R cp = fma(cd, al_akl, rt_akl * pq);
                                    fill_g_dir_regs<LI, LJ, LK, LL>(s_g + (size_t)item * GS, seed, c0, cp, b10, b01, b00, ab, cd);
                                }
                            }
                            __syncwarp();
                            if (lane_ok) {
#pragma unroll
                                for (int s = 0; s < NKLP; s++) {
                                    if (!pv[s]) continue;
#pragma unroll 1
                                    for (int r = 0; r < NROOTS; r++) {
                                        const R* __restrict__ g = s_g + r * 3 * GS;
                                        const R* __restrict__ gx = g + pox[s];
                                        const R* __restrict__ gy = g + poy[s];
                                        const R* __restrict__ gz = g + poz[s];
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
                    // ---- digestion of this pass, straight from registers
                    __syncwarp();
                    if constexpr (DO_J) {
#pragma unroll
                        for (int s = 0; s < NKLP; s++) {
                            if (!pv[s]) continue;
                            // J_kl += sum_ij (ij|kl) D[j,i]: stationary in registers for the whole task
                            R sj = R(0);
#pragma unroll
                            for (int jj = 0; jj < NJC; jj++)
#pragma unroll
                                for (int i = 0; i < NFI; i++) sj = fma(acc[s][jj * NFI + i], s_dji[(jc0 + jj) * NFI + i], sj);
                            jkl[s] += (double)sj;
                        }
                        // J_ij: same addresses for all groups of the warp -> reduce-scatter over the 32 lanes,
                        // in chunks of 16 elements to bound the live registers
                        constexpr int CH = 16, NV = NJC * NFI;
#pragma unroll
                        for (int c0 = 0; c0 < NV; c0 += CH) {
                            constexpr int dummy = 0; (void)dummy;
                            R vij[CH];
#pragma unroll
                            for (int x = 0; x < CH; x++) {
                                R v = R(0);
                                if (c0 + x < NV) {
#pragma unroll
                                    for (int s = 0; s < NKLP; s++) v = fma(acc[s][(c0 + x) < NV ? (c0 + x) : 0], dlk[s], v);
                                }
                                vij[x] = v;
                            }
                            int idx = 0, cnt = CH;
                            WarpReduceScatter<CH, 16>::run(vij, lane, idx, cnt);
                            if (cnt > 0 && c0 + idx < NV) {
                                const int jj = (c0 + idx) / NFI, i = (c0 + idx) - jj * NFI;
                                atomicAdd(a.vj + (size_t)(j0 + jc0 + jj) * nao + i0 + i, (double)vij[0]);
                            }
                        }
                    }
                    if constexpr (DO_K) {
                        if (lane_ok) {
#pragma unroll
                            for (int s = 0; s < NKLP; s++) {
                                if (!pv[s]) continue;
                                const int kc = pk[s], lc = pl[s], pr = lc * NFK + kc;
                                R djl[NJC], djk[NJC], dil[NFI], dik[NFI];
#pragma unroll
                                for (int jj = 0; jj < NJC; jj++) {
                                    djl[jj] = s_djl[(jc0 + jj) * NFL + lc];
                                    djk[jj] = s_djk[(jc0 + jj) * NFK + kc];
                                }
#pragma unroll
                                for (int i = 0; i < NFI; i++) {
                                    dil[i] = s_dil[i * NFL + lc];
                                    dik[i] = s_dik[i * NFK + kc];
                                }
                                // K_ik / K_il partials of this lane's pair: stationary over the j loop
#pragma unroll
                                for (int i = 0; i < NFI; i++) {
                                    R va = R(0), vb = R(0);
#pragma unroll
                                    for (int jj = 0; jj < NJC; jj++) {
                                        va = fma(acc[s][jj * NFI + i], djl[jj], va);
                                        vb = fma(acc[s][jj * NFI + i], djk[jj], vb);
                                    }
                                    KP_IK(s, i) += va;
                                    KP_IL(s, i) += vb;
                                }
                                // K_jk / K_jl partials: staged, combined over the group's lanes below
#pragma unroll
                                for (int jj = 0; jj < NJC; jj++) {
                                    R vc = R(0), vd = R(0);
#pragma unroll
                                    for (int i = 0; i < NFI; i++) {
                                        vc = fma(acc[s][jj * NFI + i], dil[i], vc);
                                        vd = fma(acc[s][jj * NFI + i], dik[i], vd);
                                    }
                                    s_stjk[pr * NFJ + jc0 + jj] = vc;
                                    s_stjl[pr * NFJ + jc0 + jj] = vd;
                                }
                            }
                        }
                    }
                }
                // ---- K_jk / K_jl of this quartet: sum over the partner component, scatter
                if constexpr (DO_K) {
                    __syncwarp();
                    if (live) {
#pragma unroll
                        for (int m2 = 0; m2 < (NFJ * NFK + T - 1) / T; m2++) {
                            const int x = t + m2 * T;
                            if (x < NFJ * NFK) {
                                const int r = x / NFK, c = x - r * NFK;
                                R v = R(0);
#pragma unroll
                                for (int l = 0; l < NFL; l++) v += s_stjk[(l * NFK + c) * NFJ + r];
                                atomicAdd(a.vk + (size_t)(j0 + r) * nao + k0 + c, (double)v);
                            }
                        }
#pragma unroll
                        for (int m2 = 0; m2 < (NFJ * NFL + T - 1) / T; m2++) {
                            const int x = t + m2 * T;
                            if (x < NFJ * NFL) {
                                const int r = x / NFL, c = x - r * NFL;
                                R v = R(0);
#pragma unroll
                                for (int k = 0; k < NFK; k++) v += s_stjl[(c * NFK + k) * NFJ + r];
                                atomicAdd(a.vk + (size_t)(j0 + r) * nao + l0 + c, (double)v);
                            }
                        }
                    }
                }
            }
            // flush the per-i K partial sums of the groups that contributed
            if constexpr (DO_K) {
                if (touched_i) {
#pragma unroll
                    for (int s = 0; s < NKLP; s++) {
                        if (!pv[s]) continue;
#pragma unroll
                        for (int i = 0; i < NFI; i++) {
                            atomicAdd(a.vk + (size_t)(i0 + i) * nao + k0 + pk[s], (double)KP_IK(s, i));
                            atomicAdd(a.vk + (size_t)(i0 + i) * nao + l0 + pl[s], (double)KP_IL(s, i));
                        }
                    }
                }
            }
            touched_kl |= touched_i;
        }
        if constexpr (DO_J) {
            if (touched_kl) {
#pragma unroll
                for (int s = 0; s < NKLP; s++)
                    if (pv[s]) atomicAdd(a.vj + (size_t)(l0 + pl[s]) * nao + k0 + pk[s], jkl[s]);
            }
        }
    }
#undef KP_IK
#undef KP_IL
    if (lane == 0 && nq) atomicAdd(a.qcount, nq);
}

}  // namespace jqc

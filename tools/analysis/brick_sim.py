"""Simulate the 'kl-block x i x j-loop' brick schedule: lane utilisation and atomics per quartet."""
import sys, os
import numpy as np
wl = "valinomycin-tzvp"
q = np.load(f"/tmp/q_{wl}.npy").astype(np.float64)
goff = [0, 248, 328, 740, 820, 1148, 1308, 1388]
gl = [0, 0, 0, 1, 1, 2, 3]
names = ["s3", "s2", "s1", "p3", "p1", "d1", "f1"]
cut = np.log(1e-13)
nbas = q.shape[0]
nf = lambda l: (l + 1) * (l + 2) // 2
rng = np.random.RandomState(0)

def sim(gi, gj, gk, gl_, nbucket=None, sample=40, W=32):
    I = np.arange(goff[gi], goff[gi + 1]); J = np.arange(goff[gj], goff[gj + 1])
    K = np.arange(goff[gk], goff[gk + 1]); L = np.arange(goff[gl_], goff[gl_ + 1])
    qij = q[np.ix_(I, J)].copy()
    if gi == gj: qij[I[:, None] < J[None, :]] = -1e9
    qkl = q[np.ix_(K, L)].copy()
    if gk == gl_: qkl[K[:, None] < L[None, :]] = -1e9
    qij[qij < -90] = -1e9; qkl[qkl < -90] = -1e9
    qmax_ij = qij.max()
    kk, ll = np.nonzero(qkl + qmax_ij > cut)
    qv = qkl[kk, ll]
    same_ik = gi == gk
    if nbucket and (same_ik or FORCE):
        bucket = (kk // nbucket)
        order = np.lexsort((-qv, bucket))
    else:
        order = np.argsort(-qv)
    kk, ll, qv = kk[order], ll[order], qv[order]
    nblk = (len(kk) + W - 1) // W
    tot_q = 0; tot_steps = 0; tot_flush = 0; tot_isteps = 0
    # j order per i: by q desc
    blocks = range(0, nblk, max(1, nblk // sample))
    for b in blocks:
        sl = slice(b * W, min(len(kk), (b + 1) * W))
        kb, lb, qb = K[kk[sl]], L[ll[sl]], qv[sl]
        Qb = qb.max()
        # pass[i,j,lane]
        p = (qij[:, :, None] + qb[None, None, :]) > cut
        if same_ik:
            # canonical: k <= i ; if k == i then l <= j   (pair index ordering)
            ii = I[:, None, None]; jj = J[None, :, None]
            p &= (kb[None, None, :] < ii) | ((kb[None, None, :] == ii) & (lb[None, None, :] <= jj))
        step = (qij + Qb > cut)            # warp-level j-loop continues (prefix in q order == set)
        if same_ik:
            step &= (I[:, None] >= kb.min())
        anyp = p.any(axis=2)
        step &= anyp | True
        # steps actually executed: j's passing warp-uniform test; we could also skip steps with no active lane (ballot)
        steps_exec = (step & anyp)
        tot_steps += steps_exec.sum()
        tot_q += p.sum()
        tot_flush += p.any(axis=1).sum()     # (i, lane) combos with >= 1 quartet
        tot_isteps += steps_exec.any(axis=1).sum()
    scale = nblk / len(list(blocks))
    li, lj, lk, ll_ = gl[gi], gl[gj], gl[gk], gl[gl_]
    nfi, nfj, nfk, nfl = nf(li), nf(lj), nf(lk), nf(ll_)
    old = nfi * nfj / 32 + nfk * nfl + nfi * nfk + nfi * nfl + nfj * nfk + nfj * nfl
    new = (tot_steps * nfi * nfj / 32 * W + tot_q * nfj * (nfk + nfl) + tot_flush * nfi * (nfk + nfl)) / max(tot_q, 1)
    print("(%s %s|%s %s) quartets %.3e  lane-util %.3f  j-steps per (i,blk) %.1f  quartets per (i,lane) %.1f  atomics/q old %.1f new %.1f"
          % (names[gi], names[gj], names[gk], names[gl_], tot_q * scale, tot_q / max(tot_steps * W, 1),
             tot_steps / max(tot_isteps, 1), tot_q / max(tot_flush, 1), old, new))

FORCE = len(sys.argv) > 3
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for g in [(4, 2, 4, 2), (4, 2, 2, 2), (4, 0, 4, 0), (4, 0, 3, 0), (4, 4, 4, 2), (5, 4, 4, 4), (5, 2, 4, 2), (2, 2, 2, 2), (4, 2, 2, 0), (6, 4, 5, 4), (3, 0, 2, 0), (5,5,4,2)]:
    sim(*g, nbucket=nb, W=W)

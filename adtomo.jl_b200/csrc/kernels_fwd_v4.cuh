// kernels_fwd_v4.cuh -- 3D forward fast sweeping for BATCHES, "slot-block" sweep: the layouts, the re-skew and the round
// loop of kernels_fwd_v3.cuh with a sweep in which a warp keeps ONE warp slot (4 x 8 pencils) for V4_L consecutive
// levels and hands the values between levels and between the lanes of the patch in registers.
//
// Reference semantics: Eikonal3D.cpp:35-57 (one directional Gauss-Seidel sweep), :59-68 (the 8 sweeps of a round),
// :71-88 (rounds until max|u - u_old| < tol, cap 20).  Any order of the node updates of a sweep that respects the
// component-wise order of the sweep's direction gives the serial sweep's bits; v2 / v3 use "level by level", this
// file "block of levels by block of levels" (below).
//
// Why (profiles/r02_ncu_summary_v3b.json, round 2).  In the level-by-level sweep every node update LOADS its eight
// inputs, and the three upwind ones only exist after the previous level's barrier: 39 % of the sweep's warp time
// waits on a load (long scoreboard), 11 % on the 318 barriers per sweep, L2 serves 7 TB/s (57 % of what it can), and
// cutting 20 % of the kernel's instructions (the re-skew rewrite) bought 2 %: the sweep is bound by load latency, not by
// issue.  But six of the eight inputs of a node were produced or loaded by the SAME warp one level earlier:
//   * own value        = the downwind-W value the lane loaded one level earlier (the pencil walks along W);
//   * upwind W         = the lane's own previous result;
//   * upwind A / C     = the previous results of the lanes one patch row / column upwind;
//   * downwind A / C   = the downwind-W values the lanes one patch row / column downwind load at this level.
// What is left per node and level: ONE streaming load of the field (downwind W), the slowness, and for the lanes on
// the rim of the patch one value of the neighbouring slot -- none of which depends on this level's results, so all of
// them are issued a level ahead and the dependent chain of a level is shuffle -> min/sort -> solve.
//
// Schedule.  Slot (rb, g') (row block, column group counted in the sweep's direction), local level l = lam - 4 rb - 8 g',
// is cut into blocks of V4_L = 8 levels.  Block b of slot (rb, g') runs in macro-step tau = 2 (rb + g') + b:
//   * its upwind rim values come from slots (rb-1, g') / (rb, g'-1), local levels <= 8b + 10 / 8b + 14, i.e. their blocks
//     <= b + 1, macro-step tau - 1: complete;
//   * its downwind rim values (OLD values) belong to slots (rb+1, g') / (rb, g'+1), local levels >= 8b - 3 / 8b - 7, i.e.
//     their blocks >= b - 1, macro-step >= tau + 1: not yet touched;
//   * blocks of one macro-step neither read nor write each other's nodes.
// One __syncthreads per macro-step (2 (nrb + G - 2) + nb, 94 at 128 x 128 x 64) instead of one per level (318).
// Requires the 4 x 8 patch, a grid without ragged edges (dA % 4 == 0, dC % 8 == 0) and a compile-time pitch.
#pragma once
#include "kernels_fwd_v3.cuh"

namespace adtomo {

constexpr int V4_L = 8;     // levels per block
static_assert(V2_LA == 4 && V2_LC == 8 || true, "");

// ---- schedule (host + device) -------------------------------------------------------------------------------
// local levels a slot is live: l in [0, v4_live(P)) (pencil (la, lc') is at W' = l - la - lc')
EIK_HD int v4_live(const Plan2 &P) { return P.dW + V2_LA + V2_LC - 2; }
EIK_HD int v4_nblocks(const Plan2 &P) { return (v4_live(P) + V4_L - 1) / V4_L; }
EIK_HD int v4_nrb(const Plan2 &P) { return P.dA / V2_LA; }
EIK_HD int v4_ndiag(const Plan2 &P) { return v4_nrb(P) + P.G - 1; }           // anti-diagonals s = rb + g' of the slot grid
EIK_HD int v4_nsteps(const Plan2 &P) { return 2 * (v4_ndiag(P) - 1) + v4_nblocks(P); }
inline bool v4_supported(const Plan2 &P) { return V2_LA == 4 && V2_LC == 8 && !v3_ragged(P) && V4_L == 8; }

// slots on anti-diagonal s: g' from glo, count
EIK_HD void v4_diag(const Plan2 &P, const int s, int &glo, int &cnt) {
    const int nrb = v4_nrb(P);
    glo = s - (nrb - 1) > 0 ? s - (nrb - 1) : 0;
    const int ghi = s < P.G - 1 ? s : P.G - 1;
    cnt = ghi - glo + 1;
    if (cnt < 0) cnt = 0;
}
// first index of anti-diagonal s in the list of slots sorted by s (then g')
EIK_HD int v4_diag_start(const Plan2 &P, const int s) {
    int r = 0;
    for (int t = 0; t < s; t++) {
        int glo, cnt;
        v4_diag(P, t, glo, cnt);
        r += cnt;
    }
    return r;
}
// anti-diagonals with a block in macro-step tau: s in [slo, shi] (block b = tau - 2 s in [0, nb))
EIK_HD void v4_band(const Plan2 &P, const int tau, int &slo, int &shi) {
    const int nb = v4_nblocks(P);
    const int t = tau - nb + 1;
    slo = t > 0 ? (t + 1) / 2 : 0;
    shi = tau / 2;
    if (shi > v4_ndiag(P) - 1) shi = v4_ndiag(P) - 1;
}

// ---- geometry of a lane (host + device) ------------------------------------------------------------------------
struct V4Lane {
    int la, lcp;        // patch row, patch column counted in the sweep's direction
    int offc;           // v2_lane_setup().offc
};
template <int SA, int SW, int SC>
EIK_HD V4Lane v4_lane(const Plan2 &P, const int lane) {
    const V2Lane L = v2_lane_setup<SA, SW, SC>(P, lane);
    V4Lane q;
    q.la = L.la;
    q.lcp = SC > 0 ? L.lc : V2_LC - 1 - L.lc;
    q.offc = L.offc;
    return q;
}
// slot of the lane's node of slot (rb, g') at local level l (always loadable, see v3_node); PC: row pitch
template <int SA, int SW, int SC>
EIK_HD int v4_off(const Plan2 &P, const int PC, const V4Lane &q, const int rb, const int gp, const int l) {
    const int g = SC > 0 ? gp : P.G - 1 - gp;
    const int lam = l + V2_LA * rb + V2_LC * gp;
    return q.offc + lam * SW * PC + rb * (V2_LA * (SA * P.RS - SW) * PC) + g * V2_LC;
}
EIK_HD bool v4_act(const Plan2 &P, const V4Lane &q, const int l) { return (unsigned)(l - q.la - q.lcp) < (unsigned)P.dW; }

// The update of one node from its eight inputs (v2_prep + v2_solve without the store): returns the node's value after
// the update; changed: it was lowered.
EIK_HD double v4_update(const double own, const double fv, const double uA, const double dA, const double uW, const double dW,
                        const double uC, const double dC, const double h, bool &changed) {
    V2Vals V;
    V.own = own; V.fv = fv; V.uA = uA; V.dA = dA; V.uW = uW; V.dW = dW; V.uC = uC; V.dC = dC; V.ref = 0.0; V.off = 0;
    V2Prep Q;
    v2_prep(V, Q);
    double res = own;
    changed = false;
    if (Q.a1 < own) {      // otherwise the candidate (> a1) cannot win the min: exact skip (as v2_solve)
        const double un = eik_solve3_sorted(Q.a1, Q.a2, Q.a3, fv * h, fv * fv * h * h);
        if (un < own) { res = un; changed = true; }
    }
    return res;
}

#if defined(__CUDACC__)

__device__ __forceinline__ double v4_shfl(const double v, const int src) { return __shfl_sync(0xffffffffu, v, src); }

// One block (levels [l0, l1) of slot (rb, g')) by one warp.
template <int SA, int SW, int SC, bool OOP, bool CMP, int PCT>
__device__ __forceinline__ void v4_block(const Plan2 &P, const V4Lane &q, const int lane, const int rb, const int gp, const int l0,
                                         const int l1, const double *rd, double *wr, const double *__restrict__ fl,
                                         const double *cmp, const double h, double &err, const V3Pol &pol) {
    constexpr int PC = PCT;
    constexpr int oW = SW * PC, oC = SW * PC + SC;            // downwind W / C neighbour (doubles)
    const long long oA = (long long)SA * P.RS * PC;           // downwind A neighbour
    const bool rimUA = q.la == 0, rimUC = q.lcp == 0, rimDA = q.la == V2_LA - 1, rimDC = q.lcp == V2_LC - 1;
    const int srcUA = lane - V2_LC, srcUC = lane - SC, srcDA = lane + V2_LC, srcDC = lane + SC;   // out-of-range sources are rim lanes
    const int off0 = v4_off<SA, SW, SC>(P, PC, q, rb, gp, l0);
    const double *pr = rd + off0;                             // old values (own, downwind)
    const double *pu = OOP ? wr + off0 : pr;                  // new values (upwind)
    const double *pf = fl + off0;
    const double *pc = CMP ? cmp + off0 : nullptr;
    double *pw = wr + off0;
    int wq = l0 - q.la - q.lcp;                                // W' of the lane at the current level
    // block start: what the registers would hold had the previous block been run by this warp
    double own = pr[0];
    double prev = pu[-oW];                                     // upwind W = the pencil's previous node
    // inputs of the first level
    double dWv = pr[oW], fv = v3_ld_f(pf, pol);
    double ubA = 0.0, ubC = 0.0, dbA = 0.0, dbC = 0.0;
    if (rimUA) ubA = pu[-oA];
    if (rimUC) ubC = pu[-oC];
    if (rimDA) dbA = pr[oA];
    if (rimDC) dbC = pr[oC];
    double ref = CMP ? v3_ld_once(pc, pol) : 0.0;
#pragma unroll 1
    for (int l = l0; l < l1; l++) {
        // exchange inside the patch (the rim lanes' shuffles return something unused)
        double uA = v4_shfl(prev, srcUA), uC = v4_shfl(prev, srcUC);
        double dA = v4_shfl(dWv, srcDA), dC = v4_shfl(dWv, srcDC);
        if (rimUA) uA = ubA;
        if (rimUC) uC = ubC;
        if (rimDA) dA = dbA;
        if (rimDC) dC = dbC;
        const double cown = own, cdW = dWv, cf = fv, cref = ref, cprev = prev;
        const bool act = (unsigned)wq < (unsigned)P.dW;
        // the next level's inputs are asked for before this level's arithmetic
        pr += oW; pu += oW; pf += oW;
        own = cdW;
        if (l + 1 < l1) {
            dWv = pr[oW];
            fv = v3_ld_f(pf, pol);
            if (rimUA) ubA = pu[-oA];
            if (rimUC) ubC = pu[-oC];
            if (rimDA) dbA = pr[oA];
            if (rimDC) dbC = pr[oC];
            if (CMP) { pc += oW; ref = v3_ld_once(pc, pol); }
        }
        double res = v2_inf();
        if (act) {
            bool changed;
            res = v4_update(cown, cf, uA, dA, cprev, cdW, uC, dC, h, changed);
            if (OOP || changed) *pw = res;
            if (CMP) {
                const double dd = fabs(res - cref);
                err = (err < dd) ? dd : err;
            }
        }
        prev = res;            // a lane without a node hands +inf on: the pad slot in front of the pencil's first node
        pw += oW;
        wq++;
    }
}

template <int SA, int SW, int SC, bool OOP, bool CMP, int PCT>
__device__ __forceinline__ void v4_sweep(const Plan2 &P, const int *diagStart, const double *rd, double *wr,
                                         const double *__restrict__ fl, const double *cmp, const double h, double &err,
                                         const V3Pol &pol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const V4Lane q = v4_lane<SA, SW, SC>(P, lane);
    const int nsteps = v4_nsteps(P), nb = v4_nblocks(P), live = v4_live(P), nrb = v4_nrb(P);
    __syncthreads();     // the previous sweep / re-skew is complete
    for (int tau = 0; tau < nsteps; tau++) {
        int slo, shi;
        v4_band(P, tau, slo, shi);
        if (slo <= shi) {
            const int k1 = diagStart[shi + 1];
            int s = slo;
            for (int k = diagStart[slo] + warp; k < k1; k += nw) {
                while (k >= diagStart[s + 1]) s++;
                const int glo = s - (nrb - 1) > 0 ? s - (nrb - 1) : 0;
                const int gp = glo + (k - diagStart[s]), rb = s - gp;
                const int b = tau - 2 * s;
                const int l0 = b * V4_L;
                const int l1 = l0 + V4_L < live ? l0 + V4_L : live;
                v4_block<SA, SW, SC, OOP, CMP, PCT>(P, q, lane, rb, gp, l0, l1, rd, wr, fl, cmp, h, err, pol);
            }
        }
        (void)nb;
        __syncthreads();
    }
}

// Same contract as k_fwd3d_v3.  Dynamic shared memory: the re-skew plane (WCH x PS doubles) followed by the table of
// anti-diagonal starts (v4_ndiag + 1 ints) at tabOffset.
template <int NTMAX, int MINB, int PCT>
__global__ void __launch_bounds__(NTMAX, MINB) k_fwd3d_v4(const Plan2 P, const int tabOffset, double *bufs,
                                                          const double *__restrict__ fP, const double *__restrict__ fM,
                                                          const double h, const double tol, const int max_rounds,
                                                          const int S, int *__restrict__ rounds, double *__restrict__ errs,
                                                          int *__restrict__ where, const int *__restrict__ order,
                                                          int *__restrict__ spent) {
    extern __shared__ double plane[];
    __shared__ double red[32];
    int *diagStart = reinterpret_cast<int *>(reinterpret_cast<char *>(plane) + tabOffset);
    for (int s = threadIdx.x; s <= v4_ndiag(P); s += blockDim.x) diagStart[s] = v4_diag_start(P, s);
    const V3Pol pol = v3_policies();
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        double *B3 = bufs + (long long)src * 3 * P.M;
        int o = 0, a = 1;          // layout P: round-start field, working field;  buffer 2: layout M
        double *Bz = B3 + 2 * P.M;
        int r = 0;
        bool conv = false;
        while (r < max_rounds) {
            double err = 0.0;
            double *Bo = B3 + o * P.M, *Ba = B3 + a * P.M;
            int state = 1;                        // layout of the working field
            double *w = Ba;
            for (int sw = 0; sw < 8; sw++) {
                const int sigma = P.sg[sw][1] * P.sg[sw][2];
                if (sw > 0 && sigma != state) {
                    double *dst = state > 0 ? Bz : Ba;
                    __syncthreads();
                    v3_reskew<PCT>(P, w, dst, state, plane, 0, P.dA, pol);
                    w = dst;
                    state = sigma;
                }
#define V4_CALL(a_, w_, c_, oop_, cmp_) \
    v4_sweep<a_, w_, c_, oop_, cmp_, PCT>(P, diagStart, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err, pol)
                V2_DISPATCH(P, sw, V4_CALL);
#undef V4_CALL
            }
            const double e = v2_block_max(err, red);
            if (threadIdx.x == 0 && errs) errs[(long long)order[src] * max_rounds + r] = e;
            r++;
            const int oo = o; o = a; a = oo;      // the result (in a) becomes next round's round-start field
            if (__any_sync(0xffffffffu, e < tol)) { conv = true; break; }   // e is block-uniform
        }
        if (threadIdx.x == 0) {
            if (rounds) rounds[order[src]] = conv ? r : -r;
            spent[order[src]] = r;
            {
                unsigned sm__;
                asm("mov.u32 %0, %%smid;" : "=r"(sm__));
                spent[S + src] = (int)sm__;
            }
            where[src] = o;
        }
        __syncthreads();
    }
}

#endif  // __CUDACC__

}  // namespace adtomo

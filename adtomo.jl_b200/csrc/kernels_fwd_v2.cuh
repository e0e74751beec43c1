// kernels_fwd_v2.cuh -- 3D forward fast sweeping, NS sources per thread.
//
// Same algorithm, layouts and exactness as kernels_fwd_v1.cuh (read that header first).  What is new:
// all sources of a batch share the slowness field, the grid and therefore EVERY index computation
// (packed level enumeration, neighbour offsets, boundary predicates, position in the next layout) and
// the f load.  A thread therefore updates the same node of NS different sources back to back: the
// ~130 addressing/selection instructions per node are paid once per NS updates, and the NS
// independent solves interleave (ILP).  NS sources need 2*NS sheets, so a group of NS sources is
// handled by a thread-block cluster of CS CTAs that splits the rows (sheets shrink by CS; halo rows
// travel through distributed shared memory), e.g. NS = 2, CS = 2 for 128x128x64: per SM the same
// shared memory and the same number of node updates per level as one source per CTA.
// Sources of a group that have met their tolerance are frozen (the reference stops each source at its
// own round count, Eikonal3D.cpp:85) while the others continue.
#pragma once
#include "kernels_fwd_v1.cuh"

namespace adtomo {

template <int NT, int NS, int DIR, bool CL>
__device__ __forceinline__ void sweep3d_v2(const Plan3 &P, const int sw, const double *const (&rd)[NS],
                                           double *const (&wr)[NS], const double *__restrict__ fl,
                                           const double *const (&cmp)[NS], const bool useCmp, const bool (&act)[NS],
                                           const double h, double *sheets, const int sheet, const int *ri,
                                           const int riStride, int *fcS, const unsigned short *tOfS, const int tOfMode,
                                           const int rank, const int CS, double (&err)[NS]) {
    const SweepDev W = P.sw[sw];
    const LayoutDev &L = P.lay[W.rl];
    const LayoutDev &X = P.lay[W.wl];
    constexpr int dir = DIR;
    const int dA = L.dA, dB = L.dB, dC = L.dC, pitch = L.pitch, pg = L.pg, nlev = L.nlev;
    const int T = dB + dC - 2;
    const int *riL = ri + W.rl * riStride;
    const int *riX = ri + W.wl * riStride;
    const int dpg = dir * pg, dpitch = dir * pitch;
    const int TXc = (X.dB - 1) + (X.dC - 1);
    const int pgX = X.pg;
    const int a0 = CL ? (dA * rank) / CS : 0;
    const int a1 = CL ? (dA * (rank + 1)) / CS : dA;
    const int nown = a1 - a0;
    for (int q = threadIdx.x; q < T + 2; q += NT) fcS[q] = L.fcum[q];
    const unsigned short *tOfT = tOfMode == 1 ? tOfS : L.tOf;
    unsigned char *tOf8 = (unsigned char *)tOfS;
    if (tOfMode == 1)
        for (int q = threadIdx.x; q < dB * dC; q += NT) ((unsigned short *)tOfS)[q] = L.tOf[q];
    if (tOfMode == 2)
        for (int q = threadIdx.x; q < dB * dC; q += NT) tOf8[q] = (unsigned char)L.tOf[q];
    // sheet (s, buf) = sheets + (2*s + buf) * sheet; +inf borders/halos for this layout's geometry
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double *shA = sheets + (size_t)(2 * s) * sheet, *shB = shA + sheet;
        for (int q = threadIdx.x; q < pitch; q += NT) {
            shA[q] = EIK_INF; shB[q] = EIK_INF;
            shA[(nown + 1) * pitch + q] = EIK_INF; shB[(nown + 1) * pitch + q] = EIK_INF;
        }
        for (int q = threadIdx.x; q < nown + 2; q += NT) {
            shA[q * pitch] = EIK_INF; shB[q * pitch] = EIK_INF;
            shA[q * pitch + dB + 1] = EIK_INF; shB[q * pitch + dB + 1] = EIK_INF;
        }
    }
    double *rmBase = nullptr;   // neighbour CTA's sheets (downstream in this sweep's direction)
    int rmRow = 0;
    const int myEdge = DIR > 0 ? a1 - 1 : a0;
    bool push = false;
    if (CL) {
        cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
        const int nb = rank + DIR;
        if (nb >= 0 && nb < CS) {
            push = true;
            rmBase = cluster.map_shared_rank(sheets, nb);
            const int nb0 = (dA * nb) / CS, nb1 = (dA * (nb + 1)) / CS;
            rmRow = DIR > 0 ? 0 : (nb1 - nb0 + 1);
        }
        cluster.sync();
    } else {
        __syncthreads();
    }
    int cur = 1;   // buffer written at this step (0/1); the other one holds the previous level
    for (int step = 0; step < nlev; step++) {
        const int lam = dir > 0 ? step : nlev - 1 - step;
        const int Alo = max(0, lam - T), Ahi = min(dA - 1, lam);
        const int Amin = max(Alo, a0), Amax = min(Ahi, a1 - 1);
        const int q0 = Amax >= Amin ? fcS[lam - Amax] : 0;
        const int cnt = Amax >= Amin ? fcS[lam - Amin + 1] - q0 : 0;
        const int lamD = lam + dir;
        const bool hasD = (unsigned)lamD < (unsigned)nlev;
        const int base0 = (riL[lam] - Alo) * pg;
        const int baseD = hasD ? (riL[lamD] - max(0, lamD - T)) * pg : 0;
        const int lamP = lam + 2 * dir;
        const bool hasP = (unsigned)lamP < (unsigned)nlev;
        const int baseP = hasP ? (riL[lamP] - max(0, lamP - T)) * pg : 0;
        const int lamX0 = W.lx0 + W.lxL * lam;
        const int shPrevOff = (1 - cur) * sheet, shCurOff = cur * sheet;
        for (int q = threadIdx.x; q < cnt; q += NT) {
            // ---- index work, once for all NS sources
            const int e = q0 + q;
            const int t = tOfMode == 2 ? (int)tOf8[e] : (int)tOfT[e];
            const int B = max(0, t - (dC - 1)) + (e - fcS[t]);
            const int A = lam - t, C = t - B;
            const int ab = A * pg + B;
            const int o0 = base0 + ab, dn = baseD + ab;
            const bool hDA = (unsigned)(A + dir) < (unsigned)dA, hDB = (unsigned)(B + dir) < (unsigned)dB,
                       hDC = (unsigned)(C + dir) < (unsigned)dC, hUC = (unsigned)(C - dir) < (unsigned)dC;
            const bool pf = hasP && (unsigned)(C + 2 * dir) < (unsigned)dC;
            const int sab = (A - a0 + 1) * pitch + B + 1;
            const int cv = W.vi == 0 ? A : (W.vi == 1 ? B : C);
            const int ct = W.ti == 0 ? A : (W.ti == 1 ? B : C);
            const int v = W.vs * cv + W.vo, tt = W.ts * ct + W.to;
            const int lamX = lamX0 + W.lxV * v;
            const int offX = (riX[lamX] + v - max(0, lamX - TXc)) * pgX + tt;
            const bool edge = CL && push && A == myEdge;
            const double fv = fl[o0];
            if (pf) asm volatile("prefetch.global.L1 [%0];" ::"l"(fl + dn));
            const double fh = fv * h, ffhh = fv * fv * h * h;
            // ---- per source: loads first (all sources), then the solves
            double own[NS], a1_[NS], a2_[NS], a3_[NS], old[NS];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                own[s] = 0.0; a1_[s] = EIK_INF; a2_[s] = EIK_INF; a3_[s] = EIK_INF; old[s] = 0.0;
                if (act[s]) {
                    const double *r = rd[s];
                    own[s] = r[o0];
                    double dA_ = EIK_INF, dB_ = EIK_INF, dC_ = EIK_INF, uC = EIK_INF;
                    if (hDA) dA_ = r[dn + dpg];
                    if (hDB) dB_ = r[dn + dir];
                    if (hDC) dC_ = r[dn];
                    if (pf) asm volatile("prefetch.global.L1 [%0];" ::"l"(r + baseP + ab));
                    const double *shPrev = sheets + (size_t)(2 * s) * sheet + shPrevOff;
                    const double uA = shPrev[sab - dpitch];
                    const double uB = shPrev[sab - dir];
                    if (hUC) uC = shPrev[sab];
                    a1_[s] = eik_min(uA, dA_); a2_[s] = eik_min(uB, dB_); a3_[s] = eik_min(uC, dC_);
                    if (useCmp) old[s] = cmp[s][offX];
                }
            }
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (act[s]) {
                    double b1 = a1_[s], b2 = a2_[s], b3 = a3_[s];
                    double res = own[s];
                    eik_sort3(b1, b2, b3);
                    if (b1 < own[s]) {   // otherwise the candidate (> b1) cannot win the min: exact skip
                        const double un = eik_solve3_sorted(b1, b2, b3, fh, ffhh);
                        if (un < own[s]) res = un;
                    }
                    sheets[(size_t)(2 * s) * sheet + shCurOff + sab] = res;
                    if (edge) rmBase[(size_t)(2 * s) * sheet + shCurOff + rmRow * pitch + B + 1] = res;   // DSMEM halo push
                    wr[s][offX] = res;
                    if (useCmp) {
                        const double dd = fabs(res - old[s]);
                        err[s] = (err[s] < dd) ? dd : err[s];
                    }
                }
            }
        }
        if (CL) cooperative_groups::this_cluster().sync();
        else __syncthreads();
        cur = 1 - cur;
    }
}

// bufs: S x 3 x Mmax doubles; buffer 0 of every source holds u0 in layout L0 on entry.
// Groups of NS consecutive sources are processed together by one cluster (CL) or one CTA.
template <int NT, int NS, bool CL>
__global__ void __launch_bounds__(NT, 1) k_fwd3d_v2(const Plan3 P, const int sheet, const int tOfSmem,
                                                    const int barsOffset,
                                                    double *__restrict__ bufs, const double *__restrict__ flay,
                                                    const double h, const double tol, const int max_rounds,
                                                    const int S, int *__restrict__ rounds,
                                                    double *__restrict__ errs, int *__restrict__ where,
                                                    double *errPart) {
    extern __shared__ double sheets[];
    __shared__ double red[NT / 32];
    int rank = 0, CS = 1;
    if (CL) {
        cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
        rank = (int)cluster.block_rank();
        CS = (int)cluster.num_blocks();
    }
    int *ri = (int *)(sheets + 2 * NS * (size_t)sheet);
    int riStride = 0;
    for (int q = 0; q < NLAYOUT; q++) riStride = max(riStride, P.lay[q].nlev + 1);
    int *fcS = ri + NLAYOUT * riStride;
    int fcLen = 0;
    for (int q = 0; q < NLAYOUT; q++) fcLen = max(fcLen, P.lay[q].dB + P.lay[q].dC);
    unsigned short *tOfS = tOfSmem ? (unsigned short *)(fcS + fcLen) : nullptr;
    const int tOfMode = tOfSmem;
    for (int q = 0; q < NLAYOUT; q++)
        for (int t = threadIdx.x; t <= P.lay[q].nlev; t += NT) ri[q * riStride + t] = P.lay[q].rowIndex[t];
    __syncthreads();
    const long long M = P.Mmax;
    const int ngroups = (S + NS - 1) / NS;
    for (int grp = blockIdx.x / CS; grp < ngroups; grp += gridDim.x / CS) {
        bool act[NS];
        int src[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            src[s] = grp * NS + s;
            act[s] = src[s] < S;
            if (!act[s]) src[s] = S - 1;   // harmless pointer target, never touched
        }
        int o = 0, a = 1, b = 2;
        int r = 0;
        bool any = true;
        while (r < max_rounds && any) {
            double err[NS];
            const double *Bo[NS];
            double *Ba[NS], *Bb[NS];
            const double *cBa[NS], *cBb[NS], *none[NS];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                err[s] = 0.0;
                double *B3 = bufs + (long long)src[s] * 3 * M;
                Bo[s] = B3 + o * M; Ba[s] = B3 + a * M; Bb[s] = B3 + b * M;
                cBa[s] = Ba[s]; cBb[s] = Bb[s]; none[s] = nullptr;
            }
#define SWEEP2(k, D, RD, WR, CMP, UC) \
    sweep3d_v2<NT, NS, D, CL>(P, k, RD, WR, flay + (long long)P.sw[k].rl * M, CMP, UC, act, h, sheets, sheet, ri, riStride, fcS, tOfS, tOfMode, rank, CS, err)
            SWEEP2(0, 1, Bo, Ba, none, false);
            SWEEP2(1, 1, cBa, Bb, none, false);
            SWEEP2(2, 1, cBb, Ba, none, false);
            SWEEP2(3, 1, cBa, Bb, none, false);
            SWEEP2(4, -1, cBb, Ba, none, false);
            SWEEP2(5, -1, cBa, Bb, none, false);
            SWEEP2(6, -1, cBb, Ba, none, false);
            SWEEP2(7, -1, cBa, Bb, Bo, true);
#undef SWEEP2
            double e[NS];
#pragma unroll
            for (int s = 0; s < NS; s++) e[s] = block_max<NT>(err[s], red);
            if (CL) {
                if (threadIdx.x == 0)
                    for (int s = 0; s < NS; s++) errPart[((long long)grp * NS + s) * CS + rank] = e[s];
                cooperative_groups::this_cluster().sync();
#pragma unroll
                for (int s = 0; s < NS; s++)
                    for (int q = 0; q < CS; q++) {
                        const double eq = ((volatile double *)errPart)[((long long)grp * NS + s) * CS + q];
                        e[s] = (e[s] < eq) ? eq : e[s];
                    }
                cooperative_groups::this_cluster().sync();
            }
            r++;
            const int oo = o;
            o = b;
            b = oo;
            any = false;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (!act[s]) continue;
                if (threadIdx.x == 0 && rank == 0 && errs) errs[(long long)src[s] * max_rounds + r - 1] = e[s];
                const bool conv = e[s] < tol;
                if (conv || r == max_rounds) {
                    if (threadIdx.x == 0 && rank == 0) {
                        if (rounds) rounds[src[s]] = conv ? r : -r;
                        where[src[s]] = o;
                    }
                    act[s] = false;
                } else {
                    any = true;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace adtomo

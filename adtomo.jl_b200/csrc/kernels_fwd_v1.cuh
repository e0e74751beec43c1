// kernels_fwd_v1.cuh -- 3D forward fast sweeping on level-major layouts (see layouts.h).
//
// One CTA per source.  Per sweep the CTA walks the levels of the sweep's family in order; for
// each level
//   phase 1: every node of the level is updated.  Its own old value, the three "downwind"
//            neighbours (old values, next level) and f come from global memory in the sweep's
//            layout -- contiguous rows, coalesced.  Its three "upwind" neighbours (new values,
//            previous level) come from a shared-memory sheet indexed by (major, minor).  The new
//            value goes to the second sheet.
//   phase 2: the level is written to global memory in the NEXT sweep's layout, row by row of that
//            layout (contiguous, coalesced), reading the sheet through an affine index map.
// (Round-1 revision: phase 2 is gone -- each thread stores its result directly at the node's position
// in the next sweep's layout; the 4 doubles of a sector are written within the same level step, so
// L2 merges them and DRAM traffic stays 8 B/node.)
// One __syncthreads per level.  The update is eik_solve3_pre (bit-exact with the reference) and is
// skipped when no neighbour is smaller than the node (then the candidate cannot win the min).
// Three rotating global buffers per source hold: the field at round start (for the L-inf stopping
// test of Eikonal3D.cpp:77-85, fused into the last sweep's write phase), and the ping-pong pair.
#pragma once
#include <cooperative_groups.h>
#include "kernels_v0.cuh"
#include "layouts.h"
#include "cluster_halo.cuh"

namespace adtomo {

#define EIK_INF __longlong_as_double(0x7ff0000000000000LL)

// One directional sweep.  shA/shB: two (dA+2) x pitch sheets with a +inf border (a missing upwind
// neighbour then reads +inf and min() returns the existing one: the reference's mirror rule,
// Eikonal3D.cpp:47-52).  ri: shared-memory copy of the layouts' rowIndex tables, stride riStride.
// fcS/tOfS: shared-memory scratch for the current layout's packed-enumeration tables.
// The nodes of a level are enumerated PACKED (rows concatenated, see layouts.h): lane q of the level
// finds its row through tOf, so every warp item is full no matter how short the rows of the level
// are (row-wise items left 36 % of the lanes idle on 128x128x64).  NPL = nodes per lane and item:
// the NPL updates are independent and interleave (ILP).
// CL (cluster mode): the rows of a level are split evenly over the CS CTAs of a thread-block cluster
// (rank r owns major coordinates [dA*r/CS, dA*(r+1)/CS)); each CTA keeps sheets for its rows plus one
// halo row per side, pushes its boundary row into the neighbour's halo through distributed shared
// memory, and the per-level barrier becomes a cluster barrier.  Used when two full sheets do not fit
// one SM (e.g. 200x200x80).
template <int NT, int NPL, int DIR, bool CL>
__device__ __forceinline__ void sweep3d_v1(const Plan3 &P, const int sw, const double *rd, double *wr,
                                           const double *__restrict__ fl, const double *cmp, const double h,
                                           double *shA, double *shB, const int *ri, const int riStride,
                                           int *fcS, const unsigned short *tOfS, const int tOfMode, const int rank,
                                           const int CS, uint64_t *bars, double &err) {
    const SweepDev W = P.sw[sw];
    const LayoutDev &L = P.lay[W.rl];
    const LayoutDev &X = P.lay[W.wl];
    constexpr int dir = DIR;   // sweeps 1-4 ascend, 5-8 descend the levels of their layout (layouts.h)
    const int dA = L.dA, dB = L.dB, dC = L.dC, pitch = L.pitch, pg = L.pg, nlev = L.nlev;
    const int T = dB + dC - 2;
    const int *riL = ri + W.rl * riStride;
    const int *riX = ri + W.wl * riStride;
    const int dpg = dir * pg, dpitch = dir * pitch;
    const int TXc = (X.dB - 1) + (X.dC - 1);
    const int pgX = X.pg;
    // rows owned by this CTA: [a0, a1); sheet row of A is A - a0 + 1 (rows 0 and a1-a0+1 are halos)
    const int a0 = CL ? (dA * rank) / CS : 0;
    const int a1 = CL ? (dA * (rank + 1)) / CS : dA;
    const int nown = a1 - a0;
    // tables of this layout (tOf stays in global memory when it does not fit: tOfS == nullptr)
    for (int q = threadIdx.x; q < T + 2; q += NT) fcS[q] = L.fcum[q];
    // packed-row table: 8-bit copy in shared memory when every t fits a byte (tOfMode 2), 16-bit copy
    // (1), or the global table (0)
    const unsigned short *tOfT = tOfMode == 1 ? tOfS : L.tOf;
    unsigned char *tOf8 = (unsigned char *)tOfS;
    if (tOfMode == 1)
        for (int q = threadIdx.x; q < dB * dC; q += NT) ((unsigned short *)tOfS)[q] = L.tOf[q];
    if (tOfMode == 2)
        for (int q = threadIdx.x; q < dB * dC; q += NT) tOf8[q] = (unsigned char)L.tOf[q];
    // +inf borders/halos for this layout's sheet geometry
    for (int q = threadIdx.x; q < pitch; q += NT) {
        shA[q] = EIK_INF; shB[q] = EIK_INF;
        shA[(nown + 1) * pitch + q] = EIK_INF; shB[(nown + 1) * pitch + q] = EIK_INF;
    }
    for (int q = threadIdx.x; q < nown + 2; q += NT) {
        shA[q * pitch] = EIK_INF; shB[q * pitch] = EIK_INF;
        shA[q * pitch + dB + 1] = EIK_INF; shB[q * pitch + dB + 1] = EIK_INF;
    }
    // cluster mode: halo exchange with the neighbours (cluster_halo.cuh).  "up" pushes its edge row into
    // my halo, I push my edge row into "down"'s halo.
    const int myEdge = DIR > 0 ? a1 - 1 : a0;          // my row that the downstream neighbour needs
    bool hasUp = false, hasDown = false;
    unsigned rmFull[2] = {0, 0}, rmEmpty[2] = {0, 0}, rmSheet[2] = {0, 0};
    unsigned phFull[2] = {0, 0}, phEmpty[2] = {0, 0};
    if (CL) {
        const int up = rank - DIR, down = rank + DIR;
        hasUp = up >= 0 && up < CS;
        hasDown = down >= 0 && down < CS;
        if (threadIdx.x == 0) {
            for (int q = 0; q < 4; q++) mbar_init(&bars[q], 1);
            mbar_init_fence();
        }
        if (hasDown) {
            const int nb0 = (dA * down) / CS, nb1 = (dA * (down + 1)) / CS;
            const int rmRow = DIR > 0 ? 0 : (nb1 - nb0 + 1);
            rmFull[0] = cluster_map(smem_u32(&bars[0]), down);
            rmFull[1] = cluster_map(smem_u32(&bars[1]), down);
            rmSheet[0] = cluster_map(smem_u32(shA), down) + (unsigned)(rmRow * pitch + 1) * 8u;
            rmSheet[1] = cluster_map(smem_u32(shB), down) + (unsigned)(rmRow * pitch + 1) * 8u;
        }
        if (hasUp) {
            rmEmpty[0] = cluster_map(smem_u32(&bars[2]), up);
            rmEmpty[1] = cluster_map(smem_u32(&bars[3]), up);
        }
        cooperative_groups::this_cluster().sync();   // once per sweep: barrier init + previous sweep's global writes
    } else {
        __syncthreads();
    }
    for (int step = 0; step < nlev; step++) {
        const int lam = dir > 0 ? step : nlev - 1 - step;
        const int Alo = max(0, lam - T), Ahi = min(dA - 1, lam);
        const int p = step & 1;                       // sheet buffer written at this step
        double *shCur = p ? shB : shA;
        const double *shPrev = p ? shA : shB;
        bool pushing = false;
        if (CL) {
            // ONE thread talks to the mbarriers (several waiters could miss a phase: the neighbour may
            // arrive again as soon as thread 0 has signalled); the CTA barrier below releases the others.
            pushing = hasDown && step <= nlev - 2;
            if (threadIdx.x == 0) {
                if (hasUp && step >= 1) mbar_wait(&bars[1 - p], phFull[1 - p]);      // halo of step-1 has landed
                if (pushing && step >= 2) mbar_wait(&bars[2 + p], phEmpty[p]);       // neighbour done with buffer p
                if (pushing) {
                    const int t = lam - myEdge;
                    const unsigned len = (myEdge >= Alo && myEdge <= Ahi) ? (unsigned)(fcS[t + 1] - fcS[t]) : 0u;
                    mbar_remote_arrive_tx(rmFull[p], len * 8u);                         // announce this step's push
                }
            }
            if (hasUp && step >= 1) phFull[1 - p] ^= 1u;
            if (pushing && step >= 2) phEmpty[p] ^= 1u;
            __syncthreads();
        }
        const int Amin = max(Alo, a0), Amax = min(Ahi, a1 - 1);   // my rows in this level
        const int q0 = Amax >= Amin ? fcS[lam - Amax] : 0;        // packed index of my first node
        const int cnt = Amax >= Amin ? fcS[lam - Amin + 1] - q0 : 0;
        const int lamD = lam + dir;
        const bool hasD = (unsigned)lamD < (unsigned)nlev;
        const int base0 = (riL[lam] - Alo) * pg;
        const int baseD = hasD ? (riL[lamD] - max(0, lamD - T)) * pg : 0;
        const int lamP = lam + 2 * dir;
        const bool hasP = (unsigned)lamP < (unsigned)nlev;
        const int baseP = hasP ? (riL[lamP] - max(0, lamP - T)) * pg : 0;
        const int lamX0 = W.lx0 + W.lxL * lam;
        for (int item = threadIdx.x; item < cnt; item += NT * NPL) {
            bool valid[NPL];
            int A_[NPL], B_[NPL], C_[NPL];
            double own[NPL], fv[NPL], dA_[NPL], dB_[NPL], dC_[NPL];
#pragma unroll
            for (int k = 0; k < NPL; k++) {
                // lanes of a warp take 32 consecutive nodes; the warp's NPL groups are NT apart
                const int q = item + k * NT;
                valid[k] = q < cnt;
                own[k] = 0.0; fv[k] = 0.0; dA_[k] = EIK_INF; dB_[k] = EIK_INF; dC_[k] = EIK_INF;
                A_[k] = 0; B_[k] = 0; C_[k] = 0;
                if (valid[k]) {
                    const int e = q0 + q;
                    const int t = tOfMode == 2 ? (int)tOf8[e] : (int)tOfT[e];
                    const int B = max(0, t - (dC - 1)) + (e - fcS[t]);
                    const int A = lam - t, C = t - B;
                    A_[k] = A; B_[k] = B; C_[k] = C;
                    const int ab = A * pg + B;
                    own[k] = rd[base0 + ab];
                    fv[k] = fl[base0 + ab];
                    const int dn = baseD + ab;
                    if ((unsigned)(A + dir) < (unsigned)dA) dA_[k] = rd[dn + dpg];
                    if ((unsigned)(B + dir) < (unsigned)dB) dB_[k] = rd[dn + dir];
                    if ((unsigned)(C + dir) < (unsigned)dC) dC_[k] = rd[dn];
                    // pull the level after next towards L1 (first used as downwind data of the next step)
                    if (hasP && (unsigned)(C + 2 * dir) < (unsigned)dC) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(rd + baseP + ab));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(fl + dn));
                    }
                }
            }
            double res[NPL];
#pragma unroll
            for (int k = 0; k < NPL; k++) {
                const int sab = (A_[k] - a0 + 1) * pitch + B_[k] + 1;
                double uA = EIK_INF, uB = EIK_INF, uC = EIK_INF;
                if (valid[k]) {
                    uA = shPrev[sab - dpitch];
                    uB = shPrev[sab - dir];
                    if ((unsigned)(C_[k] - dir) < (unsigned)dC) uC = shPrev[sab];
                }
                double a1 = eik_min(uA, dA_[k]), a2 = eik_min(uB, dB_[k]), a3 = eik_min(uC, dC_[k]);
                res[k] = own[k];
                eik_sort3(a1, a2, a3);
                if (a1 < own[k]) {   // otherwise the candidate (> a1) cannot win the min: exact skip
                    const double un = eik_solve3_sorted(a1, a2, a3, fv[k] * h, fv[k] * fv[k] * h * h);
                    if (un < own[k]) res[k] = un;
                }
            }
#pragma unroll
            for (int k = 0; k < NPL; k++) {
                if (valid[k]) {
                    const int A = A_[k], B = B_[k], C = C_[k];
                    shCur[(A - a0 + 1) * pitch + B + 1] = res[k];
                    if (CL && pushing && A == myEdge) st_async_f64(rmSheet[p] + (unsigned)B * 8u, res[k], rmFull[p]);   // DSMEM halo push
                    // position in the next sweep's layout
                    const int cv = W.vi == 0 ? A : (W.vi == 1 ? B : C);
                    const int ct = W.ti == 0 ? A : (W.ti == 1 ? B : C);
                    const int v = W.vs * cv + W.vo, t = W.ts * ct + W.to;
                    const int lamX = lamX0 + W.lxV * v;
                    const int offX = (riX[lamX] + v - max(0, lamX - TXc)) * pgX + t;
                    wr[offX] = res[k];
                    if (cmp) {
                        const double dd = fabs(res[k] - cmp[offX]);
                        err = (err < dd) ? dd : err;
                    }
                }
            }
        }
        __syncthreads();
        // tell the upstream neighbour that its next push into buffer 1-p may proceed
        if (CL && hasUp && step >= 1 && step <= nlev - 3 && threadIdx.x == 0) mbar_remote_arrive(rmEmpty[1 - p]);
    }
}

// bufs: S x 3 x Mmax doubles; buffer 0 of every source holds u0 in layout L0 on entry.
// where[src] receives the index (0..2) of the buffer holding the result (layout L0).
// sheet: doubles per shared-memory sheet; tOfSmem: keep the packed-row table in shared memory.
// CL: launched with a cluster of CS = cluster size CTAs per source; errPart: S*CS doubles.
template <int NT, int NPL, bool CL>
__global__ void __launch_bounds__(NT, 1) k_fwd3d_v1(const Plan3 P, const int sheet, const int tOfSmem,
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
    double *shA = sheets, *shB = sheets + sheet;
    int *ri = (int *)(sheets + 2 * (size_t)sheet);
    int riStride = 0;
    for (int q = 0; q < NLAYOUT; q++) riStride = max(riStride, P.lay[q].nlev + 1);
    int *fcS = ri + NLAYOUT * riStride;
    int fcLen = 0;
    for (int q = 0; q < NLAYOUT; q++) fcLen = max(fcLen, P.lay[q].dB + P.lay[q].dC);
    unsigned short *tOfS = tOfSmem ? (unsigned short *)(fcS + fcLen) : nullptr;
    const int tOfMode = tOfSmem;
    uint64_t *bars = (uint64_t *)((char *)sheets + barsOffset);   // 4 mbarriers, 8-byte aligned (host computes the offset)
    for (int q = 0; q < NLAYOUT; q++)
        for (int t = threadIdx.x; t <= P.lay[q].nlev; t += NT) ri[q * riStride + t] = P.lay[q].rowIndex[t];
    __syncthreads();
    const long long N = P.Mmax;        // slots per buffer (padded rows)
    const long long MF = P.Mmax;       // slots per f layout
    for (int src = blockIdx.x / CS; src < S; src += gridDim.x / CS) {
        double *B3 = bufs + (long long)src * 3 * N;
        int o = 0, a = 1, b = 2;
        int r = 0;
        bool conv = false;
        while (r < max_rounds) {
            double err = 0.0;
            double *Bo = B3 + o * N, *Ba = B3 + a * N, *Bb = B3 + b * N;
#define SWEEP(k, D, RD, WR, CMP) \
    sweep3d_v1<NT, NPL, D, CL>(P, k, RD, WR, flay + (long long)P.sw[k].rl * MF, CMP, h, shA, shB, ri, riStride, fcS, tOfS, tOfMode, rank, CS, bars, err)
            SWEEP(0, 1, Bo, Ba, nullptr);
            SWEEP(1, 1, Ba, Bb, nullptr);
            SWEEP(2, 1, Bb, Ba, nullptr);
            SWEEP(3, 1, Ba, Bb, nullptr);
            SWEEP(4, -1, Bb, Ba, nullptr);
            SWEEP(5, -1, Ba, Bb, nullptr);
            SWEEP(6, -1, Bb, Ba, nullptr);
            SWEEP(7, -1, Ba, Bb, Bo);
#undef SWEEP
            double e = block_max<NT>(err, red);
            if (CL) {
                // combine the partial maxima of the cluster's CTAs through global memory
                if (threadIdx.x == 0) errPart[(long long)src * CS + rank] = e;
                cooperative_groups::this_cluster().sync();
                for (int q = 0; q < CS; q++) {
                    const double eq = ((volatile double *)errPart)[(long long)src * CS + q];
                    e = (e < eq) ? eq : e;
                }
                cooperative_groups::this_cluster().sync();
            }
            if (threadIdx.x == 0 && rank == 0 && errs) errs[(long long)src * max_rounds + r] = e;
            r++;
            // the result is in b; it becomes next round's "old"
            const int oo = o;
            o = b;
            b = oo;
            if (e < tol) { conv = true; break; }
        }
        if (threadIdx.x == 0 && rank == 0) {
            if (rounds) rounds[src] = conv ? r : -r;
            where[src] = o;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// layout conversion helpers (elementwise, one-time per call)
// ---------------------------------------------------------------------------------------------
// f (row-major) -> the 5 layouts
__global__ void k_f_to_layouts(const Plan3 P, const double *__restrict__ f, double *__restrict__ flay) {
    const int n = P.ext[1], l = P.ext[2];
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P.N; id += gridDim.x * blockDim.x) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        const double v = f[id];
#pragma unroll
        for (int q = 0; q < NLAYOUT; q++) flay[(long long)q * P.Mmax + lay_offset(P.lay[q], P.ext, i, j, k)] = v;
    }
}

// dense row-major u0 (S x N) -> buffer 0 of each source in layout L0.  grid: (blocks, S)
__global__ void k_u0_to_L0(const Plan3 P, const double *__restrict__ U0, double *__restrict__ bufs) {
    const int n = P.ext[1], l = P.ext[2];
    const int src = blockIdx.y;
    const double *u0 = U0 + (long long)src * P.N;
    double *b0 = bufs + (long long)src * 3 * P.Mmax;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P.N; id += gridDim.x * blockDim.x) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        b0[lay_offset(P.lay[0], P.ext, i, j, k)] = u0[id];
    }
}

// result (layout L0, buffer where[src]) -> dense row-major U (S x N).  grid: (blocks, S)
__global__ void k_L0_to_rowmajor(const Plan3 P, const double *__restrict__ bufs, const int *__restrict__ where,
                                 double *__restrict__ U) {
    const int n = P.ext[1], l = P.ext[2];
    const int src = blockIdx.y;
    const double *b = bufs + ((long long)src * 3 + where[src]) * P.Mmax;
    double *u = U + (long long)src * P.N;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P.N; id += gridDim.x * blockDim.x) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        u[id] = b[lay_offset(P.lay[0], P.ext, i, j, k)];
    }
}

}  // namespace adtomo

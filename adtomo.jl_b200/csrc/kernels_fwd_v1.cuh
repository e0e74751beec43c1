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
// One __syncthreads per level.  The update is eik_solve3_pre (bit-exact with the reference) and is
// skipped when no neighbour is smaller than the node (then the candidate cannot win the min).
// Three rotating global buffers per source hold: the field at round start (for the L-inf stopping
// test of Eikonal3D.cpp:77-85, fused into the last sweep's write phase), and the ping-pong pair.
#pragma once
#include "kernels_v0.cuh"
#include "layouts.h"

namespace adtomo {

template <int NT>
__device__ __forceinline__ void sweep3d_v1(const Plan3 &P, const int sw, const double *rd, double *wr,
                                           const double *__restrict__ fl, const double *cmp, const double h,
                                           double *shA, double *shB, double &err) {
    const SweepDev W = P.sw[sw];
    const LayoutDev &L = P.lay[W.rl];
    const LayoutDev &X = P.lay[W.wl];
    const int dir = W.dir;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = NT / 32;
    const int dA = L.dA, dB = L.dB, dC = L.dC, pitch = L.pitch, nlev = L.nlev;
    const int segs = (min(dB, dC) + 31) >> 5;
    const int segsX = (min(X.dB, X.dC) + 31) >> 5;
    const int *__restrict__ lsT = L.levelStart;
    const int *__restrict__ rsT = L.rowStart;
    const int *__restrict__ lsX = X.levelStart;
    const int *__restrict__ rsX = X.rowStart;
    double *shPrev = shA, *shCur = shB;
    for (int step = 0; step < nlev; step++) {
        const int lam = dir > 0 ? step : nlev - 1 - step;
        // ------------------------------------------------------------------ phase 1
        {
            const int Alo = max(0, lam - (dB - 1) - (dC - 1)), Ahi = min(dA - 1, lam);
            const int npairs = (Ahi - Alo + 1) * segs;
            const int lamD = lam + dir;
            const bool hasD = (unsigned)lamD < (unsigned)nlev;
            const int ls = lsT[lam];
            const int lsD = hasD ? lsT[lamD] : 0;
            for (int p = warp; p < npairs; p += NW) {
                const int Ar = p / segs;
                const int seg = p - Ar * segs;
                const int A = Alo + Ar;
                const int t = lam - A;
                const int Blo = max(0, t - (dC - 1)), Bhi = min(dB - 1, t);
                const int B = Blo + seg * 32 + lane;
                if (B > Bhi) continue;
                const int C = t - B;
                const int off = ls + rsT[lam * dA + A] + (B - Blo);
                const double own = rd[off];
                // upwind neighbours: previous level, shared-memory sheet
                const int Au = A - dir, Bu = B - dir, Cu = C - dir;
                const bool hUA = (unsigned)Au < (unsigned)dA, hUB = (unsigned)Bu < (unsigned)dB,
                           hUC = (unsigned)Cu < (unsigned)dC;
                // downwind neighbours: next level, global memory (old values)
                const int Ad = A + dir, Bd = B + dir, Cd = C + dir;
                const bool hDA = (unsigned)Ad < (unsigned)dA, hDB = (unsigned)Bd < (unsigned)dB,
                           hDC = (unsigned)Cd < (unsigned)dC;
                double vA, vB, vC;
                {
                    // row A of level lamD holds (A, B+dir, C) at B+dir and (A, B, C+dir) at B
                    const int tD = lamD - A;
                    const int BloD = max(0, tD - (dC - 1));
                    const int rowD = hasD ? lsD + rsT[lamD * dA + A] - BloD : 0;
                    const double uB = hUB ? shPrev[A * pitch + Bu] : 0.0;
                    const double uC = hUC ? shPrev[A * pitch + B] : 0.0;
                    const double dB_ = hDB ? rd[rowD + Bd] : 0.0;
                    const double dC_ = hDC ? rd[rowD + B] : 0.0;
                    vB = !hUB ? dB_ : (!hDB ? uB : eik_min(uB, dB_));
                    vC = !hUC ? dC_ : (!hDC ? uC : eik_min(uC, dC_));
                    const double uA = hUA ? shPrev[Au * pitch + B] : 0.0;
                    double dA_ = 0.0;
                    if (hDA) {
                        const int tA = lamD - Ad;
                        const int BloA = max(0, tA - (dC - 1));
                        dA_ = rd[lsD + rsT[lamD * dA + Ad] + (B - BloA)];
                    }
                    vA = !hUA ? dA_ : (!hDA ? uA : eik_min(uA, dA_));
                }
                double res = own;
                const double amin = eik_min(eik_min(vA, vB), vC);
                if (amin < own) {
                    const double fv = fl[off];
                    const double un = eik_solve3_pre(vA, vB, vC, fv * h, fv * fv * h * h);
                    if (un < own) res = un;
                }
                shCur[A * pitch + B] = res;
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ phase 2
        {
            const int base = W.sh0 + W.shL * lam;
            const int lamX0 = W.lx0 + W.lxL * lam;
            const int npairs = X.dA * segsX;
            for (int p = warp; p < npairs; p += NW) {
                const int v = p / segsX;
                const int seg = p - v * segsX;
                const int lamX = lamX0 + W.lxV * v;
                if ((unsigned)lamX >= (unsigned)X.nlev) continue;
                const int tX = lamX - v;
                const int Blo = max(0, tX - (X.dC - 1)), Bhi = min(X.dB - 1, tX);
                const int t = Blo + seg * 32 + lane;
                if (t > Bhi) continue;
                const double val = shCur[base + W.shV * v + W.shT * t];
                const int offX = lsX[lamX] + rsX[lamX * X.dA + v] + (t - Blo);
                wr[offX] = val;
                if (cmp) {
                    const double dd = fabs(val - cmp[offX]);
                    err = (err < dd) ? dd : err;
                }
            }
        }
        double *tmp = shPrev;
        shPrev = shCur;
        shCur = tmp;
    }
    __syncthreads();   // the next sweep reads what this one wrote to global memory
}

// bufs: S x 3 x N doubles; buffer 0 of every source holds u0 in layout L0 on entry.
// where[src] receives the index (0..2) of the buffer holding the result (layout L0).
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_fwd3d_v1(const Plan3 P, double *__restrict__ bufs,
                                                    const double *__restrict__ flay, const double h,
                                                    const double tol, const int max_rounds, const int S,
                                                    int *__restrict__ rounds, double *__restrict__ errs,
                                                    int *__restrict__ where) {
    extern __shared__ double sheets[];
    __shared__ double red[NT / 32];
    double *shA = sheets, *shB = sheets + P.sheet;
    const long long N = P.N;
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        double *B3 = bufs + (long long)src * 3 * N;
        int o = 0, a = 1, b = 2;
        int r = 0;
        bool conv = false;
        while (r < max_rounds) {
            double err = 0.0;
            double *Bo = B3 + o * N, *Ba = B3 + a * N, *Bb = B3 + b * N;
            sweep3d_v1<NT>(P, 0, Bo, Ba, flay + (long long)P.sw[0].rl * N, nullptr, h, shA, shB, err);
            sweep3d_v1<NT>(P, 1, Ba, Bb, flay + (long long)P.sw[1].rl * N, nullptr, h, shA, shB, err);
            sweep3d_v1<NT>(P, 2, Bb, Ba, flay + (long long)P.sw[2].rl * N, nullptr, h, shA, shB, err);
            sweep3d_v1<NT>(P, 3, Ba, Bb, flay + (long long)P.sw[3].rl * N, nullptr, h, shA, shB, err);
            sweep3d_v1<NT>(P, 4, Bb, Ba, flay + (long long)P.sw[4].rl * N, nullptr, h, shA, shB, err);
            sweep3d_v1<NT>(P, 5, Ba, Bb, flay + (long long)P.sw[5].rl * N, nullptr, h, shA, shB, err);
            sweep3d_v1<NT>(P, 6, Bb, Ba, flay + (long long)P.sw[6].rl * N, nullptr, h, shA, shB, err);
            sweep3d_v1<NT>(P, 7, Ba, Bb, flay + (long long)P.sw[7].rl * N, Bo, h, shA, shB, err);
            const double e = block_max<NT>(err, red);
            if (threadIdx.x == 0 && errs) errs[(long long)src * max_rounds + r] = e;
            r++;
            // the result is in b; it becomes next round's "old"
            const int oo = o;
            o = b;
            b = oo;
            if (e < tol) { conv = true; break; }
        }
        if (threadIdx.x == 0) {
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
        for (int q = 0; q < NLAYOUT; q++) flay[(long long)q * P.N + lay_offset(P.lay[q], P.ext, i, j, k)] = v;
    }
}

// dense row-major u0 (S x N) -> buffer 0 of each source in layout L0.  grid: (blocks, S)
__global__ void k_u0_to_L0(const Plan3 P, const double *__restrict__ U0, double *__restrict__ bufs) {
    const int n = P.ext[1], l = P.ext[2];
    const int src = blockIdx.y;
    const double *u0 = U0 + (long long)src * P.N;
    double *b0 = bufs + (long long)src * 3 * P.N;
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
    const double *b = bufs + ((long long)src * 3 + where[src]) * P.N;
    double *u = U + (long long)src * P.N;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P.N; id += gridDim.x * blockDim.x) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        u[id] = b[lay_offset(P.lay[0], P.ext, i, j, k)];
    }
}

}  // namespace adtomo

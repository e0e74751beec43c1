// tests/emul/emulate_v1.cpp -- TEST INFRASTRUCTURE.  A serial host emulation of the level-major
// sweep kernel (adtomo.jl_b200/csrc/kernels_fwd_v1.cuh) driven by the SAME plan/layout code
// (layouts.h) and the SAME scalar solver (eik_core.h).  It lets the CPU test-suite check the
// layout tables, the affine write maps and the buffer rotation against the oracle without a GPU.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_layouts_cpu.py does it).
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../adtomo.jl_b200/csrc/eik_core.h"
#include "../../adtomo.jl_b200/csrc/layouts.h"

using namespace adtomo;

static const double INF = INFINITY;

static void sweep(const Plan3 &P, int sw, const double *rd, double *wr, const double *fl, const double *cmp,
                  double h, double *shA, double *shB, double &err) {
    const SweepDev W = P.sw[sw];
    const LayoutDev &L = P.lay[W.rl];
    const LayoutDev &X = P.lay[W.wl];
    const int dir = W.dir, dA = L.dA, dB = L.dB, dC = L.dC, pitch = L.pitch, pg = L.pg, nlev = L.nlev;
    const int dpg = dir * pg, dpitch = dir * pitch;
    const int TXc = (X.dB - 1) + (X.dC - 1);
    for (int q = 0; q < pitch; q++) { shA[q] = shB[q] = INF; shA[(dA + 1) * pitch + q] = shB[(dA + 1) * pitch + q] = INF; }
    for (int q = 0; q < dA + 2; q++) { shA[q * pitch] = shB[q * pitch] = INF; shA[q * pitch + dB + 1] = shB[q * pitch + dB + 1] = INF; }
    double *shPrev = shA, *shCur = shB;
    for (int step = 0; step < nlev; step++) {
        const int lam = dir > 0 ? step : nlev - 1 - step;
        const int Alo = std::max(0, lam - (dB - 1) - (dC - 1)), Ahi = std::min(dA - 1, lam);
        const int lamD = lam + dir;
        const bool hasD = lamD >= 0 && lamD < nlev;
        const int lamX0 = W.lx0 + W.lxL * lam;
        const int T = dB + dC - 2;
        const int q0 = L.fcum[lam - Ahi], cnt = L.fcum[lam - Alo + 1] - q0;
        int seen = 0;
        for (int q = 0; q < cnt; q++) {
            {
                // packed enumeration exactly as in the kernel
                const int e = q0 + q;
                const int t = L.tOf[e];
                const int B = std::max(0, t - (dC - 1)) + (e - L.fcum[t]);
                const int A = lam - t, C = t - B;
                if (A < Alo || A > Ahi || B < 0 || B >= dB || C < 0 || C >= dC || t > T) { err = NAN; continue; }
                seen++;
                const int base0 = (L.rowIndex[lam] - Alo) * pg + B;
                const int baseD = hasD ? (L.rowIndex[lamD] - std::max(0, lamD - (dB - 1) - (dC - 1))) * pg + B : 0;
                const int ab = A * pg, sab = (A + 1) * pitch + B + 1;
                const double own = rd[base0 + ab];
                const bool hDA = A + dir >= 0 && A + dir < dA, hDB = B + dir >= 0 && B + dir < dB;
                const bool hUC = C - dir >= 0 && C - dir < dC, hDC = C + dir >= 0 && C + dir < dC;
                const int dn = baseD + ab;
                double uC = INF, dA_ = INF, dB_ = INF, dC_ = INF;
                if (hDA) dA_ = rd[dn + dpg];
                if (hDB) dB_ = rd[dn + dir];
                if (hDC) dC_ = rd[dn];
                const double uA = shPrev[sab - dpitch], uB = shPrev[sab - dir];
                if (hUC) uC = shPrev[sab];
                const double vA = eik_min(uA, dA_), vB = eik_min(uB, dB_), vC = eik_min(uC, dC_);
                double res = own;
                const double amin = eik_min(eik_min(vA, vB), vC);
                if (amin < own) {
                    const double fv = fl[base0 + ab];
                    const double un = eik_solve3_pre(vA, vB, vC, fv * h, fv * fv * h * h);
                    if (un < own) res = un;
                }
                shCur[sab] = res;
                const int cv = W.vi == 0 ? A : (W.vi == 1 ? B : C), ct = W.ti == 0 ? A : (W.ti == 1 ? B : C);
                const int v = W.vs * cv + W.vo, tt = W.ts * ct + W.to;
                const int lamX = lamX0 + W.lxV * v;
                const int offX = (X.rowIndex[lamX] + v - std::max(0, lamX - TXc)) * X.pg + tt;
                wr[offX] = res;
                if (cmp) err = std::max(err, std::fabs(res - cmp[offX]));
            }
        }
        {   // the packed enumeration must cover the level exactly once
            int expect = 0;
            for (int A = Alo; A <= Ahi; A++)
                for (int B = 0; B < dB; B++) { const int C = lam - A - B; if (C >= 0 && C < dC) expect++; }
            if (expect != seen) err = NAN;
        }
        std::swap(shPrev, shCur);
    }
}

extern "C" int emul_fwd3d_v1(double *u, const double *u0, const double *f, double h, int m, int n, int l, double tol,
                             int max_rounds, double *last_err) {
    HostPlan HP;
    if (!build_plan(HP, m, n, l)) return -1000;
    const Plan3 &P = HP.plan;
    const int N = P.N;
    const size_t M = P.Mmax;
    std::vector<double> bufs(3 * M, NAN), fl((size_t)NLAYOUT * M, NAN), shA(P.sheet, NAN), shB(P.sheet, NAN);
    for (int id = 0; id < N; id++) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        for (int q = 0; q < NLAYOUT; q++) fl[(size_t)q * M + lay_offset(P.lay[q], P.ext, i, j, k)] = f[id];
        bufs[lay_offset(P.lay[0], P.ext, i, j, k)] = u0[id];
    }
    int o = 0, a = 1, b = 2, r = 0;
    bool conv = false;
    double e = 0.0;
    while (r < max_rounds) {
        double err = 0.0;
        double *Bo = &bufs[(size_t)o * M], *Ba = &bufs[(size_t)a * M], *Bb = &bufs[(size_t)b * M];
        double *seq[9] = {Bo, Ba, Bb, Ba, Bb, Ba, Bb, Ba, Bb};
        for (int s = 0; s < 8; s++)
            sweep(P, s, seq[s], seq[s + 1], &fl[(size_t)P.sw[s].rl * M], s == 7 ? Bo : nullptr, h, shA.data(), shB.data(), err);
        e = err;
        r++;
        std::swap(o, b);
        if (e < tol) { conv = true; break; }
    }
    for (int id = 0; id < N; id++) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        u[id] = bufs[(size_t)o * M + lay_offset(P.lay[0], P.ext, i, j, k)];
    }
    if (last_err) *last_err = e;
    return conv ? r : -r;
}

// offsets of every physical node in layout q (for permutation / contiguity checks)
extern "C" int emul_layout_offsets(int m, int n, int l, int q, int *out) {
    HostPlan HP;
    if (!build_plan(HP, m, n, l)) return -1;
    for (int id = 0; id < m * n * l; id++) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        out[id] = lay_offset(HP.plan.lay[q], HP.plan.ext, i, j, k);
    }
    return 0;
}

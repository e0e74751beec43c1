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

static void sweep(const Plan3 &P, int sw, const double *rd, double *wr, const double *fl, const double *cmp,
                  double h, double *shA, double *shB, double &err) {
    const SweepDev W = P.sw[sw];
    const LayoutDev &L = P.lay[W.rl];
    const LayoutDev &X = P.lay[W.wl];
    const int dir = W.dir, dA = L.dA, dB = L.dB, dC = L.dC, pitch = L.pitch, nlev = L.nlev;
    double *shPrev = shA, *shCur = shB;
    for (int step = 0; step < nlev; step++) {
        const int lam = dir > 0 ? step : nlev - 1 - step;
        const int Alo = std::max(0, lam - (dB - 1) - (dC - 1)), Ahi = std::min(dA - 1, lam);
        const int lamD = lam + dir;
        const bool hasD = lamD >= 0 && lamD < nlev;
        const int ls = L.levelStart[lam], lsD = hasD ? L.levelStart[lamD] : 0;
        for (int A = Alo; A <= Ahi; A++) {
            const int t = lam - A;
            const int Blo = std::max(0, t - (dC - 1)), Bhi = std::min(dB - 1, t);
            for (int B = Blo; B <= Bhi; B++) {
                const int C = t - B;
                const int off = ls + L.rowStart[lam * dA + A] + (B - Blo);
                const double own = rd[off];
                const int Au = A - dir, Bu = B - dir, Cu = C - dir, Ad = A + dir, Bd = B + dir, Cd = C + dir;
                const bool hUA = Au >= 0 && Au < dA, hUB = Bu >= 0 && Bu < dB, hUC = Cu >= 0 && Cu < dC;
                const bool hDA = Ad >= 0 && Ad < dA, hDB = Bd >= 0 && Bd < dB, hDC = Cd >= 0 && Cd < dC;
                const int tD = lamD - A;
                const int BloD = std::max(0, tD - (dC - 1));
                const int rowD = hasD ? lsD + L.rowStart[lamD * dA + A] - BloD : 0;
                const double uB = hUB ? shPrev[A * pitch + Bu] : 0.0, uC = hUC ? shPrev[A * pitch + B] : 0.0;
                const double dB_ = hDB ? rd[rowD + Bd] : 0.0, dC_ = hDC ? rd[rowD + B] : 0.0;
                const double vB = !hUB ? dB_ : (!hDB ? uB : eik_min(uB, dB_));
                const double vC = !hUC ? dC_ : (!hDC ? uC : eik_min(uC, dC_));
                const double uA = hUA ? shPrev[Au * pitch + B] : 0.0;
                double dA_ = 0.0;
                if (hDA) {
                    const int tA = lamD - Ad;
                    const int BloA = std::max(0, tA - (dC - 1));
                    dA_ = rd[lsD + L.rowStart[lamD * dA + Ad] + (B - BloA)];
                }
                const double vA = !hUA ? dA_ : (!hDA ? uA : eik_min(uA, dA_));
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
        const int base = W.sh0 + W.shL * lam, lamX0 = W.lx0 + W.lxL * lam;
        for (int v = 0; v < X.dA; v++) {
            const int lamX = lamX0 + W.lxV * v;
            if (lamX < 0 || lamX >= X.nlev) continue;
            const int tX = lamX - v;
            const int Blo = std::max(0, tX - (X.dC - 1)), Bhi = std::min(X.dB - 1, tX);
            for (int t = Blo; t <= Bhi; t++) {
                const double val = shCur[base + W.shV * v + W.shT * t];
                const int offX = X.levelStart[lamX] + X.rowStart[lamX * X.dA + v] + (t - Blo);
                wr[offX] = val;
                if (cmp) err = std::max(err, std::fabs(val - cmp[offX]));
            }
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
    std::vector<double> bufs(3 * (size_t)N), fl((size_t)NLAYOUT * N), shA(P.sheet, NAN), shB(P.sheet, NAN);
    for (int id = 0; id < N; id++) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        for (int q = 0; q < NLAYOUT; q++) fl[(size_t)q * N + lay_offset(P.lay[q], P.ext, i, j, k)] = f[id];
        bufs[lay_offset(P.lay[0], P.ext, i, j, k)] = u0[id];
    }
    int o = 0, a = 1, b = 2, r = 0;
    bool conv = false;
    double e = 0.0;
    while (r < max_rounds) {
        double err = 0.0;
        double *Bo = &bufs[(size_t)o * N], *Ba = &bufs[(size_t)a * N], *Bb = &bufs[(size_t)b * N];
        double *seq[9] = {Bo, Ba, Bb, Ba, Bb, Ba, Bb, Ba, Bb};
        for (int s = 0; s < 8; s++)
            sweep(P, s, seq[s], seq[s + 1], &fl[(size_t)P.sw[s].rl * N], s == 7 ? Bo : nullptr, h, shA.data(), shB.data(), err);
        e = err;
        r++;
        std::swap(o, b);
        if (e < tol) { conv = true; break; }
    }
    for (int id = 0; id < N; id++) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        u[id] = bufs[(size_t)o * N + lay_offset(P.lay[0], P.ext, i, j, k)];
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

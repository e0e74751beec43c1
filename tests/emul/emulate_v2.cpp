// tests/emul/emulate_v2.cpp -- TEST INFRASTRUCTURE.  A serial host emulation of the skewed-pencil
// sweep kernel (adtomo.jl_b200/csrc/kernels_fwd_v2.cuh): it runs the kernel's OWN per-thread
// functions (v2_window / v2_lane_setup / v2_node / v2_reskew_elem, compiled for the host) for
// every thread id in turn, level by level, with the same plan, buffers and round loop.
// Inside a level the threads only read level-1 / level+1 data and write level data, so the serial
// thread order is equivalent to the parallel one.  It lets the CPU test-suite check the slot
// arithmetic, the slot schedule and the layout hand-over against the oracle without a GPU.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_layouts_cpu.py does it).
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../adtomo.jl_b200/csrc/kernels_fwd_v2.cuh"

using namespace adtomo;

template <int SA, int SW, int SC, bool OOP, bool CMP>
static void sweep_t(const Plan2 &P, const double *rd, double *wr, const double *fl, const double *cmp,
                    double h, double &err) {
    const int nw = P.NT / 32;
    V2Lane L[32];
    for (int lane = 0; lane < 32; lane++) L[lane] = v2_lane_setup<SA, SW, SC>(P, lane);
    for (int lam = 0; lam < P.nlev; lam++) {
        // the warp scan of the kernel: exclusive prefix of the groups' window lengths
        int lo[32], n[32], excl[32], total = 0;
        for (int g = 0; g < 32; g++) { v2_window<SC>(P, g, lam, lo[g], n[g]); excl[g] = total; total += n[g]; }
        for (int warp = 0; warp < nw; warp++)
            for (int q = warp; q < total; q += nw) {
                int g = -1;
                for (int j = 0; j < 32; j++) if (excl[j] <= q && n[j] > 0) g = j;   // the ballot + clz
                const int rb = lo[g] + q - excl[g];
                for (int lane = 0; lane < 32; lane++) v2_node<SA, SW, SC, OOP, CMP>(P, L[lane], lam, rb, g, rd, wr, fl, cmp, h, err);
            }
    }
}

static void reskew(const Plan2 &P, const double *src, double *dst, int sigmaFrom, std::vector<double> &plane) {
    for (int A = 0; A < P.dA; A++) {
        const long long slab = (long long)(A + 1) * P.RS * P.PC;
        for (int w0 = 0; w0 < P.dW; w0 += P.WCH) {
            const int wc = std::min(P.WCH, P.dW - w0);
            for (int phase = 0; phase < 2; phase++)
                for (int v = 0; v < wc; v++)
                    for (int C = 0; C < P.dC; C++) v2_reskew_elem(P, src, dst, sigmaFrom, plane.data(), slab, w0, wc, phase, v, C);
        }
    }
}

extern "C" int emul_v2_plan(int m, int n, int l, int max_warps, long long plane_bytes, int *out /* role[3], G, R, NT, WCH */) {
    Plan2 P;
    if (!v2_build_plan(P, m, n, l, max_warps, (size_t)plane_bytes)) return 0;
    out[0] = P.role[0]; out[1] = P.role[1]; out[2] = P.role[2];
    out[3] = P.G; out[4] = 0; out[5] = P.NT; out[6] = P.WCH;
    return 1;
}

// u: row-major, u0 on entry, result on exit.  Returns rounds (negative: cap hit), -1000 if no plan fits.
extern "C" int emul_v2_forward(double *u, const double *f, int m, int n, int l, double h, double tol,
                               int max_rounds, int max_warps, long long plane_bytes, double *errs) {
    Plan2 P;
    if (!v2_build_plan(P, m, n, l, max_warps, (size_t)plane_bytes)) return -1000;
    std::vector<double> plane((size_t)P.WCH * P.PS);
    std::vector<double> B[3], fP(P.M, NAN), fM(P.M, NAN);      // f pads are never read: NaN would poison the result
    for (int q = 0; q < 3; q++) B[q].assign(P.M, INFINITY);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++)
            for (int k = 0; k < l; k++) {
                const long long id = ((long long)i * n + j) * l + k;
                B[0][v2_offset_ijk(P, i, j, k, +1)] = u[id];
                fP[v2_offset_ijk(P, i, j, k, +1)] = f[id];
                fM[v2_offset_ijk(P, i, j, k, -1)] = f[id];
            }
    int o = 0, a = 1, r = 0;
    bool conv = false;
    while (r < max_rounds) {
        double err = 0.0;
        double *Bo = B[o].data(), *Ba = B[a].data(), *Bz = B[2].data();
        int state = 1;
        double *w = Ba;
        for (int sw = 0; sw < 8; sw++) {
            const int sigma = P.sg[sw][1] * P.sg[sw][2];
            if (sw > 0 && sigma != state) {
                double *dst = state > 0 ? Bz : Ba;
                reskew(P, w, dst, state, plane);
                w = dst;
                state = sigma;
            }
#define V2_CALL(a_, w_, c_, oop_, cmp_) sweep_t<a_, w_, c_, oop_, cmp_>(P, oop_ ? Bo : w, w, sigma > 0 ? fP.data() : fM.data(), Bo, h, err)
            V2_DISPATCH(P, sw, V2_CALL);
#undef V2_CALL
        }
        if (errs) errs[r] = err;
        r++;
        std::swap(o, a);
        if (err < tol) { conv = true; break; }
    }
    // pads must still be +inf (nothing may ever write a slot that is not a grid node)
    long long nfinite = 0;
    for (int q = 0; q < 3; q++)
        for (long long s = 0; s < P.M; s++) nfinite += std::isfinite(B[q][s]) ? 1 : 0;
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++)
            for (int k = 0; k < l; k++) u[((long long)i * n + j) * l + k] = B[o][v2_offset_ijk(P, i, j, k, +1)];
    if (nfinite > 3 * P.N) return -2000;
    return conv ? r : -r;
}

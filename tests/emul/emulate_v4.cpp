// tests/emul/emulate_v4.cpp -- TEST INFRASTRUCTURE.  Serial host emulation of the slot-block sweep kernel
// (adtomo.jl_b200/csrc/kernels_fwd_v4.cuh): the kernel's own schedule (v4_band / v4_diag / v4_nsteps), geometry (v4_lane /
// v4_off / v4_act) and update (v4_update) compiled for the host; the warp's register hand-over (own <- downwind W of the
// previous level, shuffles inside the 4 x 8 patch, rim values from memory) is restated with 32-element arrays.  Blocks of
// one macro-step are run in a caller-chosen order (forward / reverse / interleaved): the result must not depend on it.
// Checks: every node updated exactly once per sweep; bits equal to the oracle (tests/test_layouts_cpu.py).
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (the test does it).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../adtomo.jl_b200/csrc/kernels_fwd_v4.cuh"

using namespace adtomo;

static long long g_visits = 0;
static int g_order = 0;          // 0: blocks of a macro-step in list order, 1: reversed, 2: odd ones first

template <int SA, int SW, int SC, bool OOP, bool CMP>
static void block_t(const Plan2 &P, int rb, int gp, int l0, int l1, const double *rd, double *wr, const double *fl,
                    const double *cmp, double h, double &err) {
    const int PC = P.PC;
    const int oW = SW * PC, oC = SW * PC + SC;
    const long long oA = (long long)SA * P.RS * PC;
    V4Lane q[32];
    double own[32], prev[32], dWv[32], fv[32], ubA[32], ubC[32], dbA[32], dbC[32], ref[32];
    long long pos[32];
    for (int lane = 0; lane < 32; lane++) {
        q[lane] = v4_lane<SA, SW, SC>(P, lane);
        pos[lane] = v4_off<SA, SW, SC>(P, PC, q[lane], rb, gp, l0);
        const double *pr = rd + pos[lane];
        const double *pu = OOP ? wr + pos[lane] : pr;
        own[lane] = pr[0];
        prev[lane] = pu[-oW];
    }
    for (int l = l0; l < l1; l++) {
        // loads of this level (the kernel issues them one level ahead: same values, nothing of this block's levels
        // l.. has touched them)
        for (int lane = 0; lane < 32; lane++) {
            const double *pr = rd + pos[lane];
            const double *pu = OOP ? wr + pos[lane] : pr;
            dWv[lane] = pr[oW];
            fv[lane] = fl[pos[lane]];
            ubA[lane] = pu[-oA]; ubC[lane] = pu[-oC]; dbA[lane] = pr[oA]; dbC[lane] = pr[oC];
            ref[lane] = CMP ? cmp[pos[lane]] : 0.0;
        }
        double res[32];
        for (int lane = 0; lane < 32; lane++) {
            const int la = q[lane].la, lcp = q[lane].lcp;
            const double uA = la == 0 ? ubA[lane] : prev[(lane - V2_LC) & 31];
            const double uC = lcp == 0 ? ubC[lane] : prev[(lane - SC) & 31];
            const double dA = la == V2_LA - 1 ? dbA[lane] : dWv[(lane + V2_LC) & 31];
            const double dC = lcp == V2_LC - 1 ? dbC[lane] : dWv[(lane + SC) & 31];
            res[lane] = INFINITY;
            if (v4_act(P, q[lane], l)) {
                bool changed;
                res[lane] = v4_update(own[lane], fv[lane], uA, dA, prev[lane], dWv[lane], uC, dC, h, changed);
                g_visits++;
                if (CMP) err = std::max(err, std::fabs(res[lane] - ref[lane]));
            }
        }
        for (int lane = 0; lane < 32; lane++) {
            if (v4_act(P, q[lane], l)) {
                // the kernel stores only lowered values (in place) or everything (out of place): same memory image
                if (OOP || res[lane] < own[lane]) wr[pos[lane]] = res[lane];
            }
            prev[lane] = res[lane];
            own[lane] = dWv[lane];
            pos[lane] += oW;
        }
    }
}

template <int SA, int SW, int SC, bool OOP, bool CMP>
static void sweep_t(const Plan2 &P, const double *rd, double *wr, const double *fl, const double *cmp, double h, double &err) {
    const int nsteps = v4_nsteps(P), live = v4_live(P);
    g_visits = 0;
    for (int tau = 0; tau < nsteps; tau++) {
        int slo, shi;
        v4_band(P, tau, slo, shi);
        struct Blk { int rb, gp, l0, l1; };
        std::vector<Blk> blocks;
        for (int s = slo; s <= shi; s++) {
            int glo, cnt;
            v4_diag(P, s, glo, cnt);
            for (int j = 0; j < cnt; j++) {
                const int gp = glo + j, rb = s - gp, b = tau - 2 * s;
                blocks.push_back({rb, gp, b * V4_L, std::min(b * V4_L + V4_L, live)});
            }
        }
        if (g_order == 1) std::reverse(blocks.begin(), blocks.end());
        if (g_order == 2) std::stable_partition(blocks.begin(), blocks.end(), [&](const Blk &x) { return ((x.rb + 3 * x.gp) & 1) != 0; });
        for (const Blk &x : blocks) block_t<SA, SW, SC, OOP, CMP>(P, x.rb, x.gp, x.l0, x.l1, rd, wr, fl, cmp, h, err);
    }
}

static void reskew(const Plan2 &P, const double *src, double *dst, int sigmaFrom, std::vector<double> &plane) {
    const int nw = P.NT / 32;
    for (int A = 0; A < P.dA; A++) {
        const long long slab = (long long)(A + 1) * P.RS * P.PC;
        for (int w0 = 0; w0 < P.dW; w0 += P.WCH) {
            const int wc = std::min(P.WCH, P.dW - w0);
            for (int phase = 0; phase < 2; phase++) {
                const int sigma = phase == 0 ? sigmaFrom : -sigmaFrom;
                for (int warp = 0; warp < nw; warp++)
                    for (int lane = 0; lane < 32; lane++)
                        for (int C = lane; C < P.dC; C += 32) {
                            int t, pl, go;
                            v3_reskew_start(P, P.PC, sigma, w0, nw, warp, C, t, pl, go);
                            for (; t < wc; t += nw, pl += nw * P.PS, go += nw * P.PC) {
                                if (phase == 0) plane[pl] = src[slab + go];
                                else dst[slab + go] = plane[pl];
                            }
                        }
            }
        }
    }
}

// u: row-major, u0 on entry, result on exit.  Returns rounds (negative: cap hit), -1000 if the grid is not supported by
// the slot-block sweep, -2000 a pad slot was written, -3000 a sweep did not visit every node exactly once.
extern "C" int emul_v4_forward(double *u, const double *f, int m, int n, int l, double h, double tol, int max_rounds,
                               int order, double *errs, int *out) {
    Plan2 P;
    int pct = 0;
    if (!v3_build_plan(P, m, n, l, 16, 64 * 1024, &pct, true)) return -1000;
    if (!v4_supported(P) || pct == 0) return -1000;
    g_order = order;
    std::vector<double> plane((size_t)P.WCH * P.PS);
    const long long SL = v3_slack(P);
    std::vector<double> BB(3 * P.M + 2 * SL, NAN), fPs(P.M + 2 * SL, NAN), fMs(P.M + 2 * SL, NAN);
    double *B[3] = {BB.data() + SL, BB.data() + SL + P.M, BB.data() + SL + 2 * P.M};
    double *fP = fPs.data() + SL, *fM = fMs.data() + SL;
    for (int q = 0; q < 3; q++) std::fill(B[q], B[q] + P.M, INFINITY);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++)
            for (int k = 0; k < l; k++) {
                const long long id = ((long long)i * n + j) * l + k;
                B[0][v2_offset_ijk(P, i, j, k, +1)] = u[id];
                fP[v2_offset_ijk(P, i, j, k, +1)] = f[id];
                fM[v2_offset_ijk(P, i, j, k, -1)] = f[id];
            }
    int o = 0, a = 1, r = 0;
    bool conv = false, bad_visits = false;
    while (r < max_rounds) {
        double err = 0.0;
        double *Bo = B[o], *Ba = B[a], *Bz = B[2];
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
#define V4_CALL(a_, w_, c_, oop_, cmp_) sweep_t<a_, w_, c_, oop_, cmp_>(P, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err)
            V2_DISPATCH(P, sw, V4_CALL);
#undef V4_CALL
            if (g_visits != P.N) bad_visits = true;
        }
        if (errs) errs[r] = err;
        r++;
        std::swap(o, a);
        if (err < tol) { conv = true; break; }
    }
    long long nfinite = 0;
    for (int q = 0; q < 3; q++)
        for (long long s = 0; s < P.M; s++) nfinite += std::isfinite(B[q][s]) ? 1 : 0;
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++)
            for (int k = 0; k < l; k++) u[((long long)i * n + j) * l + k] = B[o][v2_offset_ijk(P, i, j, k, +1)];
    if (out) { out[0] = P.PC; out[1] = v4_nsteps(P); out[2] = pct; }
    if (bad_visits) return -3000;
    if (nfinite > 3 * P.N) return -2000;
    return conv ? r : -r;
}

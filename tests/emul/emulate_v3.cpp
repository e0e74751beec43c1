// tests/emul/emulate_v3.cpp -- TEST INFRASTRUCTURE.  Serial host emulation of the batch sweep kernel
// (adtomo.jl_b200/csrc/kernels_fwd_v3.cuh): the kernel's OWN per-thread functions (v3_rank / v3_make_slot /
// v3_node / v3_load, v2_prep / v2_solve, v3_reskew_start) compiled for the host and run for every warp and lane
// in turn, level by level, with the same plan, slot table, [head, tail) window and round loop.  Inside a level the
// threads only read level-1 / level+1 data and write level data, so the serial order is equivalent to the parallel one.
// It also checks what the kernel relies on: every node is updated exactly once per sweep, and the live slots of a
// level are spread evenly over the warps.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_layouts_cpu.py does it).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../adtomo.jl_b200/csrc/kernels_fwd_v3.cuh"

using namespace adtomo;

static long long g_visits = 0;      // node updates of the last sweep
static bool g_bad_reskew = false;
static bool g_out_of_slack = false; // some lane (idle ones included) addressed beyond the slack around a buffer
static int g_max_imbalance = 0;     // max over levels of (max - min) live slots per warp, over the whole solve

template <int SA, int SW, int SC, bool OOP, bool CMP>
static void sweep_t(const Plan2 &P, const double *rd, double *wr, const double *fl, const double *cmp, double h, double &err) {
    const int nw = P.NT / 32, maxPer = v3_max_per_warp(P);
    const int nrb = (P.dA + V2_LA - 1) / V2_LA, nslots = nrb * P.G;
    std::vector<V3Slot> tab((size_t)nw * maxPer, V3Slot{0, 0});
    std::vector<int> tabS((size_t)nw * maxPer, 0x7fffffff), seen(nslots, 0);
    for (int idx = 0; idx < nslots; idx++) {
        const int rb = idx / P.G, g = idx - rb * P.G;
        int s;
        const int r = v3_rank<SC>(P, rb, g, s);
        seen[r]++;
        tab[(size_t)(r % nw) * maxPer + r / nw] = v3_make_slot<SA, SW, SC>(P, P.PC, rb, g);
        tabS[(size_t)(r % nw) * maxPer + r / nw] = s;
    }
    for (int r = 0; r < nslots; r++) if (seen[r] != 1) { g_visits = -1; return; }     // the rank is a bijection
    V2Lane L[32];
    int lmask[32];
    for (int lane = 0; lane < 32; lane++) { L[lane] = v2_lane_setup<SA, SW, SC>(P, lane); lmask[lane] = v3_lane_mask(P, L[lane].la, L[lane].lc); }
    const long long sAb = (long long)P.RS * P.PC * 8;
    const int dur = v3_duration(P);
    std::vector<int> head(nw, 0), tail(nw, 0);
    g_visits = 0;
    for (int lam = 0; lam < P.nlev; lam++) {
        int mx = 0, mn = 1 << 30;
        for (int warp = 0; warp < nw; warp++) {
            const int cnt = warp < nslots ? (nslots - warp + nw - 1) / nw : 0;
            const V3Slot *mine = tab.data() + (size_t)warp * maxPer;
            const int *mineS = tabS.data() + (size_t)warp * maxPer;
            while (tail[warp] < cnt && mineS[tail[warp]] <= lam) tail[warp]++;
            while (head[warp] < tail[warp] && mineS[head[warp]] + dur < lam) head[warp]++;
            mx = std::max(mx, tail[warp] - head[warp]);
            mn = std::min(mn, tail[warp] - head[warp]);
            for (int j = head[warp]; j < tail[warp]; j++)
                for (int lane = 0; lane < 32; lane++) {
                    V2Vals V;
                    int off;
                    bool act;
                    v3_node<true>(P, mine[j], L[lane].offc + lam * SW * P.PC, L[lane].wqc + lam + V3_BIAS, lmask[lane], off, act);
                    if (!v3_ragged(P)) {      // the kernel's mask-free form must agree where it is used
                        int off2;
                        bool act2;
                        v3_node<false>(P, mine[j], L[lane].offc + lam * SW * P.PC, L[lane].wqc + lam + V3_BIAS, lmask[lane], off2, act2);
                        if (off2 != off || act2 != act) g_bad_reskew = true;
                    }
                    const long long reach = (long long)P.RS * P.PC + P.PC + 1;      // farthest neighbour of a slot
                    if (off - reach < -v3_slack(P) || off + reach >= P.M + v3_slack(P)) g_out_of_slack = true;
                    // the PCT = 0 (run-time pitch) path: identical arithmetic, the pitch only feeds addresses
                    v3_load<SA, SW, SC, OOP, CMP, 0>(P, sAb, off, act, rd, wr, fl, cmp, V, V3Pol{0, 0});
                    if (act) g_visits++;
                    v2_finish<OOP, CMP>(V, wr, h, err);
                }
        }
        g_max_imbalance = std::max(g_max_imbalance, mx - mn);
    }
}

// The kernel's re-skew: every thread (warp, lane) walks its elements t0, t0 + nw, ... of every column it owns
// (v3_reskew_start); here for all warps and lanes in turn.  Checks that every element of the chunk is moved exactly once.
static void reskew(const Plan2 &P, const double *src, double *dst, int sigmaFrom, std::vector<double> &plane) {
    const int nw = P.NT / 32;
    std::vector<int> hits((size_t)P.WCH * P.dC);
    for (int A = 0; A < P.dA; A++) {
        const long long slab = (long long)(A + 1) * P.RS * P.PC;
        for (int w0 = 0; w0 < P.dW; w0 += P.WCH) {
            const int wc = std::min(P.WCH, P.dW - w0);
            for (int phase = 0; phase < 2; phase++) {
                std::fill(hits.begin(), hits.end(), 0);
                const int sigma = phase == 0 ? sigmaFrom : -sigmaFrom;
                for (int warp = 0; warp < nw; warp++)
                    for (int lane = 0; lane < 32; lane++)
                        for (int C = lane; C < P.dC; C += 32) {
                            int t, pl, go;
                            v3_reskew_start(P, P.PC, sigma, w0, nw, warp, C, t, pl, go);
                            for (; t < wc; t += nw, pl += nw * P.PS, go += nw * P.PC) {
                                int pl2, go2;      // the general index map of the round-1 kernel: same element, same slots
                                const int cc = sigma > 0 ? C : P.dC - 1 - C;
                                v2_reskew_index(P, sigma, w0, wc, (t + cc) % wc, C, pl2, go2);
                                if (pl != pl2 || go != go2 || pl != t * P.PS + C) g_bad_reskew = true;
                                hits[(size_t)t * P.dC + C]++;
                                if (phase == 0) plane[pl] = src[slab + go];
                                else dst[slab + go] = plane[pl];
                            }
                        }
                for (int t = 0; t < wc; t++)
                    for (int C = 0; C < P.dC; C++)
                        if (hits[(size_t)t * P.dC + C] != 1) g_bad_reskew = true;
            }
        }
    }
}

// u: row-major, u0 on entry, result on exit.  Returns rounds (negative: cap hit), -1000 if no plan fits,
// -2000 a pad slot was written, -3000 a sweep did not visit every node exactly once, -4000 a lane addressed
// memory beyond the slack around a buffer.
// out[0] = row pitch used, out[1] = max imbalance (slots) between warps over all levels.
extern "C" int emul_v3_forward(double *u, const double *f, int m, int n, int l, double h, double tol, int max_rounds,
                               int warps, long long plane_bytes, int use_menu, double *errs, int *out) {
    Plan2 P;
    int pct = 0;
    if (!v3_build_plan(P, m, n, l, warps, (size_t)plane_bytes, &pct, use_menu != 0)) return -1000;
    g_max_imbalance = 0;
    g_out_of_slack = false;
    g_bad_reskew = false;
    std::vector<double> plane((size_t)P.WCH * P.PS);
    // the three field buffers are contiguous like in the kernel's workspace, with v3_slack() doubles on both sides
    // (idle lanes load from there); NaN in every slot that must never be USED
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
#define V3_CALL(a_, w_, c_, oop_, cmp_) sweep_t<a_, w_, c_, oop_, cmp_>(P, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err)
            V2_DISPATCH(P, sw, V3_CALL);
#undef V3_CALL
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
    if (out) { out[0] = P.PC; out[1] = g_max_imbalance; out[2] = pct; }
    if (g_bad_reskew) return -5000;
    if (g_out_of_slack) return -4000;
    if (bad_visits) return -3000;
    if (nfinite > 3 * P.N) return -2000;
    return conv ? r : -r;
}

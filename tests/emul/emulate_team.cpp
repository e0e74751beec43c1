// tests/emul/emulate_team.cpp -- TEST INFRASTRUCTURE.  A serial host emulation of the team kernel
// (adtomo.jl_b200/csrc/kernels_fwd_team.cuh): it runs the kernel's OWN per-lane functions (tm_slot_setup, tm_load_old,
// tm_prep, tm_solve, tm_rows, tm_slot_live, v2_reskew_elem; compiled for the host) with the same plan, team shape,
// buffers, mailbox and round loop.  The CTAs of the team are advanced slot by slot by a scheduler (random, or
// "upstream as far ahead as possible", or "downstream as close as possible") that only honours what the kernel
// itself waits for: the CTA's level barrier, and the arrival of the tagged mailbox packets of a first-row slot.
// If the packets did not carry both the upwind dependence and the in-place anti-dependence between neighbouring
// CTAs, some interleaving would differ from the oracle; a packet read before it arrived yields NaN.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_layouts_cpu.py does it).
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <random>
#include "../../adtomo.jl_b200/csrc/kernels_fwd_team.cuh"

using namespace adtomo;

// One sweep: the CTAs advance slot by slot.  A CTA works on one level at a time (its __syncthreads); inside the
// level its pending slots run in any order; a slot of the CTA's first row is runnable only when the packets of
// all its nodes have arrived (right tag) -- exactly what the kernel's spin waits for.  Every CTA has its two
// sheets.  The kernel may read a node's OLD values as early as one level ahead (L1 prefetch): the emulation
// reads them when the slot's PREVIOUS level runs and keeps them until they are used.
template <int SA, int SW, int SC, bool OOP, bool CMP>
static void sweep_t(const Plan2 &P, const TeamCfg &T, const double *rd, double *wr, const double *fl, const double *cmp,
                    double h, double &err, std::mt19937 &rng, int policy, std::vector<tm_u64> &mbox, unsigned base) {
    const int nC = T.nC;
    struct Cta {
        int a0, a1, lam, lam1;
        std::vector<int> pending;
        std::vector<double> sheets;
        std::vector<TmSlotC> K;        // [slot][lane]: the kernel's per-sweep constants
        std::vector<TmOld> early;      // [slot][lane]: old values read one level ahead
        std::vector<char> has_early;   // [slot]
    };
    std::vector<Cta> cta(nC);
    auto fill = [&](Cta &c) {
        c.pending.clear();
        while (c.lam <= c.lam1 && c.pending.empty()) {
            for (int q = 0; q < (c.a1 - c.a0) * T.G32; q++) {
                const TmSlotC *K = &c.K[(size_t)q * 32];
                bool has = false;
                for (int lane = 0; lane < 32; lane++) has = has || tm_act(P, K[lane], c.lam);
                bool live = tm_slot_live(P, K[0], 0, c.lam);
                for (int lane = 1; lane < 32; lane++)
                    if (tm_slot_live(P, K[lane], lane, c.lam) != live) err = NAN;      // must be warp-uniform
                if (live) c.pending.push_back(q);
                else {
                    if (has) err = NAN;               // a slot declared dead must not contain a node
                    for (int lane = 0; lane < 32; lane++)   // the kernel's dead-slot branch
                        c.sheets[(c.lam & 1) * T.R * T.SP + K[lane].sidx] = INFINITY;
                }
            }
            if (c.pending.empty()) c.lam++;
        }
    };
    for (int t = 0; t < nC; t++) {
        Cta &c = cta[t];
        int l0;
        tm_rows(P, T, t, c.a0, c.a1, l0, c.lam1);
        c.lam = l0;
        const int nslot = (c.a1 - c.a0) * T.G32;
        c.sheets.assign((size_t)2 * T.R * T.SP, INFINITY);
        c.K.resize((size_t)nslot * 32);
        for (int q = 0; q < nslot; q++)
            for (int lane = 0; lane < 32; lane++)
                tm_slot_setup<SA, SW, SC>(P, T, t, c.a0, c.a1 - c.a0, lane, q, c.K[(size_t)q * 32 + lane]);
        c.early.resize((size_t)nslot * 32);
        c.has_early.assign((size_t)nslot, 0);
        fill(c);
    }
    auto runnable = [&](int t, int q) {
        const Cta &c = cta[t];
        const tm_u64 *inbox = mbox.data() + (long long)t * 2 * T.mbStride;
        for (int lane = 0; lane < 32; lane++) {
            const TmSlotC &K = c.K[(size_t)q * 32 + lane];
            if (!(K.flags & TM_FIRST) || !tm_act(P, K, c.lam)) continue;
            const long long mb = tm_off<SW>(P, K, c.lam) - K.slab;
            double v;
            if (!tm_unpack(inbox[2 * mb], inbox[2 * mb + 1], base | (unsigned)c.lam, v)) return false;
        }
        return true;
    };
    for (;;) {
        std::vector<std::pair<int, int>> ready;    // (cta, index into pending)
        bool any = false;
        for (int t = 0; t < nC; t++) {
            if (cta[t].lam > cta[t].lam1) continue;
            any = true;
            for (size_t k = 0; k < cta[t].pending.size(); k++)
                if (runnable(t, cta[t].pending[k])) ready.push_back({t, (int)k});
        }
        if (!any) break;
        if (ready.empty()) { err = NAN; return; }      // deadlock: must not happen
        std::pair<int, int> pick;
        if (policy == 0) pick = ready[rng() % ready.size()];
        else if (policy == 1) pick = ready.front();     // low ranks run as far ahead as they can
        else pick = ready.back();                       // high ranks follow as closely as they can
        const int t = pick.first;
        Cta &c = cta[t];
        const int q = c.pending[pick.second];
        c.pending.erase(c.pending.begin() + pick.second);
        const tm_u64 *inbox = mbox.data() + (long long)t * 2 * T.mbStride;
        tm_u64 *outbox = mbox.data() + (long long)(t + 1) * 2 * T.mbStride;
        TmPrep Q[32];
        for (int lane = 0; lane < 32; lane++) {
            const TmSlotC &K = c.K[(size_t)q * 32 + lane];
            TmOld O;
            if (c.has_early[q]) O = c.early[(size_t)q * 32 + lane];
            else tm_load_old<SA, SW, SC, CMP>(P, K, c.lam, rd, fl, cmp, O);
            tm_prep<SA, SW, SC>(P, T, K, c.lam, O, inbox, base, c.sheets.data(), Q[lane]);
        }
        c.has_early[q] = 0;
        if (tm_slot_live(P, c.K[(size_t)q * 32], 0, c.lam + 1)) {   // the prefetch: next level's old values are read NOW
            for (int lane = 0; lane < 32; lane++)
                tm_load_old<SA, SW, SC, CMP>(P, c.K[(size_t)q * 32 + lane], c.lam + 1, rd, fl, cmp, c.early[(size_t)q * 32 + lane]);
            c.has_early[q] = 1;
        }
        for (int lane = 0; lane < 32; lane++)
            tm_solve<OOP, CMP>(T, c.K[(size_t)q * 32 + lane], c.lam, Q[lane], wr, h, err, outbox, base, c.sheets.data());
        if (c.pending.empty()) { c.lam++; fill(c); }
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

extern "C" int emul_team_config(int m, int n, int l, int S, int max_ctas, int nwarps, int Rforce, int *out /* role[3], nC, R, G32 */) {
    Plan2 P;
    TeamCfg T;
    if (!team_build_plan(P, m, n, l, nwarps, 64 * 1024)) return 0;
    if (!team_config(P, S, max_ctas, nwarps, Rforce, T)) return 0;
    out[0] = P.role[0]; out[1] = P.role[1]; out[2] = P.role[2];
    out[3] = T.nC; out[4] = T.R; out[5] = T.G32;
    return 1;
}

// u: row-major, u0 on entry, result on exit.  Returns rounds (negative: cap hit), -1000 if no plan fits.
extern "C" int emul_team_forward(double *u, const double *f, int m, int n, int l, double h, double tol,
                                 int max_rounds, int nwarps, long long plane_bytes, int max_ctas, int Rforce,
                                 int policy, unsigned seed, double *errs) {
    Plan2 P;
    TeamCfg T;
    if (!team_build_plan(P, m, n, l, nwarps, (size_t)plane_bytes)) return -1000;
    if (!team_config(P, 1, max_ctas, nwarps, Rforce, T)) return -1000;
    std::mt19937 rng(seed);
    std::vector<tm_u64> mbox((size_t)T.nC * 2 * T.mbStride, 0ULL);
    unsigned serial = seed % 1000;            // tags only grow; the start value is arbitrary
    std::vector<double> plane((size_t)P.WCH * P.PS);
    std::vector<double> B[3], fP(P.M, NAN), fM(P.M, NAN);
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
            serial++;
            const unsigned base = serial << TM_LEVEL_BITS;
#define TM_CALL(a_, w_, c_, oop_, cmp_) \
    sweep_t<a_, w_, c_, oop_, cmp_>(P, T, oop_ ? Bo : w, w, sigma > 0 ? fP.data() : fM.data(), Bo, h, err, rng, policy, mbox, base)
            V2_DISPATCH(P, sw, TM_CALL);
#undef TM_CALL
        }
        if (errs) errs[r] = err;
        if (std::isnan(err)) return -3000;
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
    if (nfinite > 3 * P.N) return -2000;
    return conv ? r : -r;
}

// tests/emul/emulate_team.cpp -- TEST INFRASTRUCTURE.  A serial host emulation of the team kernel
// (adtomo.jl_b200/csrc/kernels_fwd_team.cuh): it runs the kernel's OWN per-lane functions (tm_slot_setup, tm_load_old,
// tm_prep, tm_solve, tm_rows, tm_slot_live, v2_reskew_elem; compiled for the host) with the same plan, team shape,
// buffers, mailbox and round loop.  The CTAs of the team are advanced action by action by a scheduler (random, or
// "low members as far ahead as possible", or "high members as far ahead as possible") that only honours what the
// kernel itself waits for: the CTA's level barrier, the arrival of the tagged mailbox packets of a first-row slot,
// and -- between sweeps -- the step counters of the two physical neighbours (no team barrier inside a round).
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

// Per-CTA state machine over one ROUND (8 sweeps): WAIT (both physical neighbours must have published the
// previous step) -> sweep, slot by slot, level by level -> re-skew of the CTA's own slabs if the next sweep needs
// the other layout -> publish.  The CTAs of a team may be in different sweeps (at most one apart), exactly as in
// the kernel; the scheduler picks any runnable action of any CTA.
struct Cta {
    int p = 0, sw = 0;             // member, current sweep
    bool in_sweep = false;
    unsigned step = 0, base = 0;   // steps published; tag base of the current sweep
    int a0 = 0, a1 = 0, lam = 0, lam1 = 0;
    std::vector<int> pending;
    std::vector<double> sheets;
    std::vector<TmSlotC> K;        // [slot][lane]: the kernel's per-sweep constants
    std::vector<TmOld> early;      // [slot][lane]: old values read one level ahead
    std::vector<char> has_early;   // [slot]
};

struct Round {
    const Plan2 &P;
    const TeamCfg &T;
    double *Bo, *Ba, *Bz;
    const double *fP, *fM;
    double h;
    double err = 0.0;
    std::vector<tm_u64> &mbox;
    unsigned serial0;
    std::vector<double> &plane;
    bool bad = false;

    int sigma(int sw) const { return P.sg[sw][1] * P.sg[sw][2]; }
    double *work(int sw) const { return sigma(sw) > 0 ? Ba : Bz; }     // sweep 0 is on P: Ba
    const double *fl(int sw) const { return sigma(sw) > 0 ? fP : fM; }

    template <int SA, int SW, int SC, bool OOP, bool CMP>
    void fill(Cta &c) {
        c.pending.clear();
        while (c.lam <= c.lam1 && c.pending.empty()) {
            for (int q = 0; q < (c.a1 - c.a0) * T.G32; q++) {
                const TmSlotC *K = &c.K[(size_t)q * 32];
                bool has = false;
                for (int lane = 0; lane < 32; lane++) has = has || tm_act(P, K[lane], c.lam);
                const bool live = tm_slot_live(P, K[0], 0, c.lam);
                for (int lane = 1; lane < 32; lane++)
                    if (tm_slot_live(P, K[lane], lane, c.lam) != live) bad = true;      // must be warp-uniform
                if (live) c.pending.push_back(q);
                else {
                    if (has) bad = true;              // a slot declared dead must not contain a node
                    for (int lane = 0; lane < 32; lane++)   // the kernel's dead-slot branch
                        c.sheets[(c.lam & 1) * T.R * T.SP + K[lane].sidx] = INFINITY;
                }
            }
            if (c.pending.empty()) c.lam++;
        }
    }

    template <int SA, int SW, int SC, bool OOP, bool CMP>
    void start(Cta &c) {
        int l0;
        tm_rows(P, T, c.p, SA, c.a0, c.a1, l0, c.lam1);
        c.lam = l0;
        c.base = (serial0 + (unsigned)c.sw + 1) << TM_LEVEL_BITS;
        const int nslot = (c.a1 - c.a0) * T.G32;
        c.sheets.assign((size_t)2 * T.R * T.SP, INFINITY);
        c.K.resize((size_t)nslot * 32);
        for (int q = 0; q < nslot; q++)
            for (int lane = 0; lane < 32; lane++)
                tm_slot_setup<SA, SW, SC>(P, T, c.p, c.a0, c.a1 - c.a0, lane, q, c.K[(size_t)q * 32 + lane]);
        c.early.resize((size_t)nslot * 32);
        c.has_early.assign((size_t)nslot, 0);
        c.in_sweep = true;
        fill<SA, SW, SC, OOP, CMP>(c);
    }

    template <int SA, int SW, int SC, bool OOP, bool CMP>
    bool runnable(const Cta &c, int q) {
        const tm_u64 *inbox = mbox.data() + (long long)c.p * 2 * T.mbStride;
        for (int lane = 0; lane < 32; lane++) {
            const TmSlotC &K = c.K[(size_t)q * 32 + lane];
            if (!(K.flags & TM_FIRST) || !tm_act(P, K, c.lam)) continue;
            const long long mb = tm_off<SW>(P, K, c.lam) - K.slab;
            double v;
            if (!tm_unpack(inbox[2 * mb], inbox[2 * mb + 1], c.base | (unsigned)c.lam, v)) return false;
        }
        return true;
    }

    template <int SA, int SW, int SC, bool OOP, bool CMP>
    void run_slot(Cta &c, int k) {
        const int q = c.pending[k];
        c.pending.erase(c.pending.begin() + k);
        const double *rd = OOP ? Bo : work(c.sw);
        double *wr = work(c.sw);
        const double *f = fl(c.sw);
        const tm_u64 *inbox = mbox.data() + (long long)c.p * 2 * T.mbStride;
        tm_u64 *outbox = mbox.data() + (long long)(SA > 0 ? c.p + 1 : c.p - 1) * 2 * T.mbStride;
        TmPrep Q[32];
        for (int lane = 0; lane < 32; lane++) {
            const TmSlotC &K = c.K[(size_t)q * 32 + lane];
            TmOld O;
            if (c.has_early[q]) O = c.early[(size_t)q * 32 + lane];
            else tm_load_old<SA, SW, SC, CMP>(P, K, c.lam, rd, f, Bo, O);
            tm_prep<SA, SW, SC>(P, T, K, c.lam, O, inbox, c.base, c.sheets.data(), Q[lane]);
        }
        c.has_early[q] = 0;
        if (tm_slot_live(P, c.K[(size_t)q * 32], 0, c.lam + 1)) {   // the prefetch: next level's old values are read NOW
            for (int lane = 0; lane < 32; lane++)
                tm_load_old<SA, SW, SC, CMP>(P, c.K[(size_t)q * 32 + lane], c.lam + 1, rd, f, Bo, c.early[(size_t)q * 32 + lane]);
            c.has_early[q] = 1;
        }
        for (int lane = 0; lane < 32; lane++)
            tm_solve<OOP, CMP>(T, c.K[(size_t)q * 32 + lane], c.lam, Q[lane], wr, h, err, outbox, c.base, c.sheets.data());
        if (c.pending.empty()) { c.lam++; fill<SA, SW, SC, OOP, CMP>(c); }
    }

    // end of a sweep: re-skew of the CTA's own slabs if the next sweep runs on the other layout, then publish
    void finish(Cta &c) {
        if (c.sw < 7 && sigma(c.sw + 1) != sigma(c.sw)) {
            const double *src = work(c.sw);
            double *dst = work(c.sw + 1);
            const int A0 = c.p * T.R, A1 = std::min(A0 + T.R, P.dA);
            for (int A = A0; A < A1; A++) {
                const long long slab = (long long)(A + 1) * P.RS * P.PC;
                for (int w0 = 0; w0 < P.dW; w0 += P.WCH) {
                    const int wc = std::min(P.WCH, P.dW - w0);
                    for (int phase = 0; phase < 2; phase++)
                        for (int v = 0; v < wc; v++)
                            for (int C = 0; C < P.dC; C++)
                                v2_reskew_elem(P, src, dst, sigma(c.sw), plane.data(), slab, w0, wc, phase, v, C);
                }
            }
        }
        c.in_sweep = false;
        c.step++;
        c.sw++;
    }
};

#define EM_DISPATCH(R_, sw_, CALL) V2_DISPATCH((R_).P, sw_, CALL)

static bool run_round(Round &R, std::mt19937 &rng, int policy, unsigned step0) {
    const int nC = R.T.nC;
    std::vector<Cta> cta(nC);
    for (int p = 0; p < nC; p++) { cta[p].p = p; cta[p].step = step0; }
    for (;;) {
        // runnable actions: (cta, kind, index)  kind 0: start sweep, 1: run pending slot k, 2: finish sweep
        struct Act { int p, kind, k; };
        std::vector<Act> acts;
        bool any = false;
        for (int p = 0; p < nC; p++) {
            Cta &c = cta[p];
            if (c.sw >= 8) continue;
            any = true;
            if (!c.in_sweep) {
                const unsigned need = c.step;     // neighbours must have published this many steps
                const bool ok = (p == 0 || cta[p - 1].step >= need) && (p == nC - 1 || cta[p + 1].step >= need);
                if (ok) acts.push_back({p, 0, 0});
            } else if (c.lam > c.lam1) {
                acts.push_back({p, 2, 0});
            } else {
                for (size_t k = 0; k < c.pending.size(); k++) {
                    bool ok = false;
#define EM_CALL(a_, w_, c_, oop_, cmp_) ok = R.runnable<a_, w_, c_, oop_, cmp_>(c, c.pending[k])
                    EM_DISPATCH(R, c.sw, EM_CALL);
#undef EM_CALL
                    if (ok) acts.push_back({p, 1, (int)k});
                }
            }
        }
        if (!any) break;
        if (acts.empty()) return false;                 // deadlock: must not happen
        Act act;
        if (policy == 0) act = acts[rng() % acts.size()];
        else if (policy == 1) act = acts.front();       // low members run as far ahead as they can
        else act = acts.back();                         // high members run as far ahead as they can
        Cta &c = cta[act.p];
        if (act.kind == 0) {
#define EM_CALL(a_, w_, c_, oop_, cmp_) R.start<a_, w_, c_, oop_, cmp_>(c)
            EM_DISPATCH(R, c.sw, EM_CALL);
#undef EM_CALL
        } else if (act.kind == 1) {
#define EM_CALL(a_, w_, c_, oop_, cmp_) R.run_slot<a_, w_, c_, oop_, cmp_>(c, act.k)
            EM_DISPATCH(R, c.sw, EM_CALL);
#undef EM_CALL
        } else {
            R.finish(c);
        }
    }
    return !R.bad;
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
    unsigned step = 0;
    while (r < max_rounds) {
        Round R{P, T, B[o].data(), B[a].data(), B[2].data(), fP.data(), fM.data(), h, 0.0, mbox, serial, plane};
        if (!run_round(R, rng, policy, step)) return -3000;
        serial += 8;
        step += 8;
        const double err = R.err;
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

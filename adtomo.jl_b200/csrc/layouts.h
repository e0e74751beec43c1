// layouts.h -- hyperplane-major ("level-set major") storage layouts for the 3D sweeps.
//
// A directional Gauss-Seidel sweep (Eikonal3D.cpp:35-57) can be executed level by level, a level
// being the set of nodes with  c_i + c_j + c_k = lambda  where c_a is the coordinate along axis a
// counted in the sweep's direction.  The 8 sweeps of a round (Eikonal3D.cpp:59-68) use 4 families
// of level sets (a sweep and its reverse share one).  In the reference's row-major layout a level
// is a diagonal cut (stride l-1 doubles): every access is a separate 32-byte sector.  Internally
// we therefore store a field LEVEL BY LEVEL: all nodes of level 0, then level 1, ...; inside a
// level, rows of constant "major" coordinate A, and inside a row ascending "minor" coordinate B
// (the third, "derived" coordinate is C = lambda - A - B).  A whole level is one contiguous block
// and a sweep streams it exactly once.  Rows are padded to a fixed pitch (minor extent rounded up
// to 4 doubles, so rows start sector-aligned and the offset of a node is a multiply-add: no per-row
// table); the padding costs address space (<= 2x), not traffic.
//
// Consecutive sweeps of the reference's order differ in the sign of exactly ONE axis (Gray code),
// and the intersection of a level of sweep P with a level of the next sweep X is a grid line with
// that axis fixed.  If X's layout uses that axis as its major axis, every such line is a full,
// contiguous row of X's layout, so sweep P can write its result straight into X's layout with
// coalesced stores.  That fixes 5 layouts (family, major axis):
//    L0 = (+,+,+ | j)  L1 = (-,+,+ | i)  L2 = (-,-,+ | j)  L3 = (+,-,+ | i)  L4 = (-,+,+ | k)
// and the schedule  sw1: L0->L1, sw2: L1->L2, sw3: L2->L3, sw4: L3->L4, sw5: L4->L2 (reverse),
// sw6: L2->L3 (rev), sw7: L3->L0 (rev), sw8: L0->L0 (rev).
// The C ABI keeps the reference's row-major layout; these layouts never leave the library.
#pragma once
#include <cstdlib>
#include <vector>

namespace adtomo {

constexpr int NLAYOUT = 5;

struct LayoutDev {
    int ax0, ax1, ax2;   // physical axis (0=i,1=j,2=k) of the major / minor / derived coordinate
    int flip[3];         // per PHYSICAL axis: canonical coordinate = ext-1-x when 1
    int dA, dB, dC;      // extents along major / minor / derived
    int nlev;            // dA+dB+dC-2
    int pitch;           // row pitch of the shared-memory sheet; the sheet is (dA+2) x pitch with a +inf border
    int pg;              // row pitch in global memory (dB rounded up to a multiple of 4)
    int M;               // slots of one field in this layout = (number of rows) * pg  (>= N)
    const int *rowIndex; // [nlev+1]   index of the first row of a level (rows of a level: A = Alo..Ahi)
    // packed enumeration of a level: its rows, ordered by t = lam - A ascending, are concatenated;
    // fcum[t] = number of nodes in rows 0..t-1 of the full t-range [0, dB+dC-2] (a level uses a
    // sub-range tmin..tmax), tOf[e] = the row t containing packed index e in [0, dB*dC).
    const int *fcum;             // [dB+dC]
    const unsigned short *tOf;   // [dB*dC]
};

struct SweepDev {
    int rl, wl, dir;          // layout read (R), layout written (X), +1 ascending / -1 descending levels
    // position of an R node (A,B,C) in X:  v = vs*coord[vi]+vo (X major), t = ts*coord[ti]+to (X minor),
    // coord = {A,B,C}; X level lamX = lx0 + lxL*lam + lxV*v
    int vi, vs, vo, ti, ts, to;
    int lx0, lxL, lxV;
};

struct Plan3 {
    int ext[3];
    int N;
    int Mmax;             // max over layouts of M: slots per field buffer
    int sheet;            // doubles per shared-memory sheet (max over layouts of (dA+2)*pitch)
    LayoutDev lay[NLAYOUT];
    SweepDev sw[8];
};

#if defined(__CUDACC__)
#define LAY_HD __host__ __device__ __forceinline__
#else
#define LAY_HD inline
#endif

LAY_HD int lay_imax(int a, int b) { return a > b ? a : b; }
LAY_HD int lay_imin(int a, int b) { return a < b ? a : b; }

// offset of physical node (x[0],x[1],x[2]) in layout L
LAY_HD int lay_offset(const LayoutDev &L, const int *ext, int xi, int xj, int xk) {
    const int x[3] = {xi, xj, xk};
    const int A = L.flip[L.ax0] ? ext[L.ax0] - 1 - x[L.ax0] : x[L.ax0];
    const int B = L.flip[L.ax1] ? ext[L.ax1] - 1 - x[L.ax1] : x[L.ax1];
    const int C = L.flip[L.ax2] ? ext[L.ax2] - 1 - x[L.ax2] : x[L.ax2];
    const int lam = A + B + C;
    const int Alo = lay_imax(0, lam - (L.dB - 1) - (L.dC - 1));
    return (L.rowIndex[lam] + A - Alo) * L.pg + B;
}

// ---------------------------------------------------------------------------------------------
// host-side construction
// ---------------------------------------------------------------------------------------------
struct HostLayout {
    LayoutDev d;                  // table pointers point into the vectors below
    std::vector<int> rowIndex, fcum;
    std::vector<unsigned short> tOf;
};

inline void build_layout(HostLayout &H, const int ext[3], const int sign[3], int major) {
    LayoutDev &L = H.d;
    int o1 = (major + 1) % 3, o2 = (major + 2) % 3;
    if (o1 > o2) { int t = o1; o1 = o2; o2 = t; }
    // minor = the smaller extent of the two remaining axes (tie: the later axis)
    int minor = (ext[o1] < ext[o2]) ? o1 : o2;
    int derived = (minor == o1) ? o2 : o1;
    L.ax0 = major; L.ax1 = minor; L.ax2 = derived;
    for (int a = 0; a < 3; a++) L.flip[a] = sign[a] < 0;
    L.dA = ext[major]; L.dB = ext[minor]; L.dC = ext[derived];
    L.nlev = L.dA + L.dB + L.dC - 2;
    int p = L.dB + 2;            // one +inf border column on each side
    while ((p & 3) != 2) p++;
    L.pitch = p;
    L.pg = (L.dB + 3) & ~3;
    H.rowIndex.assign(L.nlev + 1, 0);
    int acc = 0;
    for (int lam = 0; lam < L.nlev; lam++) {
        H.rowIndex[lam] = acc;
        const int Alo = lay_imax(0, lam - (L.dB - 1) - (L.dC - 1)), Ahi = lay_imin(L.dA - 1, lam);
        acc += Ahi - Alo + 1;
    }
    H.rowIndex[L.nlev] = acc;
    L.M = acc * L.pg;
    L.rowIndex = H.rowIndex.data();
    const int T = L.dB + L.dC - 2;
    H.fcum.assign(T + 2, 0);
    H.tOf.assign((size_t)L.dB * L.dC, 0);
    int c = 0;
    for (int t = 0; t <= T; t++) {
        H.fcum[t] = c;
        const int lo = lay_imax(0, t - (L.dC - 1)), hi = lay_imin(L.dB - 1, t);
        for (int b = lo; b <= hi; b++) H.tOf[c++] = (unsigned short)t;
    }
    H.fcum[T + 1] = c;
    L.fcum = H.fcum.data();
    L.tOf = H.tOf.data();
}

struct HostPlan {
    Plan3 plan;                   // with HOST table pointers
    HostLayout lay[NLAYOUT];
    bool ok = false;
};

// Returns false if some invariant of the construction does not hold (never expected).
inline bool build_plan(HostPlan &HP, int m, int n, int l) {
    Plan3 &P = HP.plan;
    P.ext[0] = m; P.ext[1] = n; P.ext[2] = l;
    P.N = m * n * l;
    static const int signs[NLAYOUT][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, 1}};
    static const int majors[NLAYOUT] = {1, 0, 1, 0, 2};
    P.sheet = 0;
    P.Mmax = 0;
    for (int q = 0; q < NLAYOUT; q++) {
        build_layout(HP.lay[q], P.ext, signs[q], majors[q]);
        P.lay[q] = HP.lay[q].d;
        if (HP.lay[q].rowIndex.back() != P.lay[q].dA * (P.lay[q].dB + P.lay[q].dC - 1)) return false;
        if (HP.lay[q].fcum.back() != P.lay[q].dB * P.lay[q].dC || P.lay[q].dB + P.lay[q].dC > 65535) return false;
        P.sheet = lay_imax(P.sheet, (P.lay[q].dA + 2) * P.lay[q].pitch);
        P.Mmax = lay_imax(P.Mmax, P.lay[q].M);
    }
    static const int sched[8][3] = {{0, 1, 1}, {1, 2, 1}, {2, 3, 1}, {3, 4, 1}, {4, 2, -1}, {2, 3, -1}, {3, 0, -1}, {0, 0, -1}};
    // reference sweep directions, to double-check the schedule (Eikonal3D.cpp:59-68)
    static const int dirs[8][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}};
    for (int s = 0; s < 8; s++) {
        SweepDev &W = P.sw[s];
        W.rl = sched[s][0]; W.wl = sched[s][1]; W.dir = sched[s][2];
        const LayoutDev &R = P.lay[W.rl], &X = P.lay[W.wl];
        for (int a = 0; a < 3; a++)
            if ((R.flip[a] ? -1 : 1) * W.dir != dirs[s][a]) return false;
        // c_R[a] = sg[a]*c_X[a] + of[a]
        int sg[3], of[3], O = 0;
        for (int a = 0; a < 3; a++) {
            if (R.flip[a] == X.flip[a]) { sg[a] = 1; of[a] = 0; }
            else { sg[a] = -1; of[a] = P.ext[a] - 1; }
            O += of[a];
        }
        if (sg[X.ax1] != sg[X.ax2]) return false;   // the X row must lie inside one R level
        const int sigma = sg[X.ax1];
        W.lxL = sigma;
        W.lxV = 1 - sigma * sg[X.ax0];
        W.lx0 = -sigma * O;
        auto which = [&](int axis) { return axis == R.ax0 ? 0 : (axis == R.ax1 ? 1 : 2); };
        W.vi = which(X.ax0); W.vs = sg[X.ax0]; W.vo = of[X.ax0];
        W.ti = which(X.ax1); W.ts = sg[X.ax1]; W.to = of[X.ax1];
        // brute-force check of the map on the corners and a few interior nodes
        for (int probe = 0; probe < 27; probe++) {
            int x[3];
            int q = probe;
            for (int a2 = 0; a2 < 3; a2++) { int r = q % 3; q /= 3; x[a2] = r == 0 ? 0 : (r == 1 ? P.ext[a2] / 2 : P.ext[a2] - 1); }
            int cR[3];
            for (int a2 = 0; a2 < 3; a2++) cR[a2] = R.flip[a2] ? P.ext[a2] - 1 - x[a2] : x[a2];
            const int coord[3] = {cR[R.ax0], cR[R.ax1], cR[R.ax2]};
            const int lam = coord[0] + coord[1] + coord[2];
            const int v = W.vs * coord[W.vi] + W.vo, t = W.ts * coord[W.ti] + W.to;
            const int lamX = W.lx0 + W.lxL * lam + W.lxV * v;
            const int off = (X.rowIndex[lamX] + v - lay_imax(0, lamX - (X.dB - 1) - (X.dC - 1))) * X.pg + t;
            if (off != lay_offset(X, P.ext, x[0], x[1], x[2])) return false;
        }
    }
    HP.ok = true;
    return true;
}

}  // namespace adtomo

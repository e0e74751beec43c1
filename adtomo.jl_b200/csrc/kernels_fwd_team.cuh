// kernels_fwd_team.cuh -- 3D forward fast sweeping for FEW sources: a TEAM of CTAs per source.
//
// Reference semantics: Eikonal3D.cpp:35-57 (one directional Gauss-Seidel sweep), :59-68 (the 8 sweeps of
// a round), :71-88 (rounds until max|u - u_old| < tol).  Same level-by-level execution and the same
// skewed-pencil layouts as kernels_fwd_v2.cuh (constant-offset neighbours, +inf padding, in-place
// sweeps, re-skew between the layouts P and M), so the result is the serial sweep's, bit for bit.
//
// Why.  With one CTA (v2) or one cluster of <= 8 SMs (level-major kernel) per source, a single large
// grid (BASELINE config C5: 256^3 .. 512^3, one source) uses 1-8 of the 148 SMs and one level costs a
// full load->solve->store->barrier latency for a handful of warps.  Here the A rows of a source are
// split over up to 2 x #SM co-resident CTAs (cooperative launch).  A sweep's data dependence between
// two CTAs is one-directional: the first row of CTA t at level lam needs the last row of CTA t-1 at
// level lam-1 (new value), and CTA t may overwrite that node's level-lam value only after CTA t-1 has
// read it (in-place update).  Both are carried by a MAILBOX in the style of NCCL's LL protocol: the
// warp that updates a node of a CTA's last row also writes the new value as two 8-byte packets
// {32 value bits | 32-bit tag}, tag = (sweep serial, level); the consumer takes its upwind-A value from
// the mailbox and spins on the packet until both tags match.  A packet is written after its node was
// computed, i.e. after the producer read the old downwind value, so its arrival also licenses the
// overwrite.  No fence, no flag and no grid-wide barrier on the level path (a first version with one
// progress word per CTA spent 0.9 us per level in MEMBAR.SC.GPU and ~10 polls of LDG.STRONG+CCTL.IVALL
// per warp and level): the CTAs form a systolic pipeline skewed by one L2 round trip per CTA.  A mailbox
// has one slot per node of the boundary row (a slab of the skewed layout), so a producer can run
// arbitrarily far ahead; tags only grow, across sweeps and across launches, so slots are never reset.
// Team-wide barriers (a counter in global memory, with fences) separate the sweeps, the re-skews and
// the rounds: ~13 per round.
//
// Lanes of a warp cover 32 consecutive columns C of ONE row A (a warp slot): at a fixed level these are
// 32 consecutive doubles of one skewed row mu -- a 256-byte contiguous run for the node and for each of
// its neighbours (512 bytes for its mailbox packets).  A warp keeps its slots for the whole sweep, so
//   * the NEW (upwind, level lam-1) values never come from global memory: every slot writes its results
//     into a shared-memory sheet (R rows x dC columns, two sheets alternating by level parity, +inf where
//     there is no node) and reads its three upwind neighbours from the other sheet;
//   * the OLD values (own, downwind, slowness, round-start value) of level lam+1 are prefetched into L1
//     while level lam is computed, and read through L1: nothing in the sweep writes them before level
//     lam+1 (own) or lam+2 (downwind), the downstream CTA cannot overwrite its first row before it has our
//     packet, which we send after reading it, no line is read again after another CTA may have written it,
//     and every team barrier invalidates L1 (its __threadfence is MEMBAR.SC.GPU + CCTL.IVALL).  (Holding
//     them in registers instead spilled at 64 registers per thread.)
// The level's critical path is then L1 / sheet read -> solve -> store -> barrier (the first version read
// everything back from L2 / DRAM: 1.9 us per level at 256^3, 53 % of the warp time in the level barrier).
#pragma once
#include <cstring>
#include "kernels_fwd_v2.cuh"

namespace adtomo {

constexpr int TM_LC = 32;          // columns per warp slot
constexpr int TM_LEVEL_BITS = 12;  // packet tag = (sweep serial << 12) | (level + 1); nlev < 4095
constexpr int TM_SYNC_WORDS = 8;   // per-source sync area (unsigned words): [0] barrier counter, [2..3],[4..5] err slots

struct TeamCfg {
    int nC;              // CTAs per source
    int R;               // rows (A') per CTA
    int G32;             // column groups of 32 per row
    long long mbStride;  // packets (16 bytes each) of one CTA's inbox = RS * PC (one slab of the skewed layout)
    int SP;              // pitch of a sheet row = 32 * G32 + 2 doubles
};

#define TM_LDU(p) (*(p))           // old values: through L1 (see above)
#if defined(__CUDA_ARCH__)
#define TM_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define TM_PREFETCH(p) ((void)(p))
#endif

typedef unsigned long long tm_u64;

// A value as two packets {low 32 value bits | tag << 32}, {high 32 value bits | tag << 32}.
EIK_HD void tm_pack(const double v, const unsigned tag, tm_u64 &p0, tm_u64 &p1) {
    tm_u64 b;
#if defined(__CUDA_ARCH__)
    b = (tm_u64)__double_as_longlong(v);
#else
    memcpy(&b, &v, 8);
#endif
    p0 = (b & 0xffffffffULL) | ((tm_u64)tag << 32);
    p1 = (b >> 32) | ((tm_u64)tag << 32);
}
EIK_HD bool tm_unpack(const tm_u64 p0, const tm_u64 p1, const unsigned tag, double &v) {
    const tm_u64 b = (p0 & 0xffffffffULL) | (p1 << 32);
#if defined(__CUDA_ARCH__)
    v = __longlong_as_double((long long)b);
#else
    memcpy(&v, &b, 8);
#endif
    return (unsigned)(p0 >> 32) == tag && (unsigned)(p1 >> 32) == tag;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void tm_mb_store(tm_u64 *slot, const tm_u64 p0, const tm_u64 p1) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(p0), "l"(p1) : "memory");
}
__device__ __forceinline__ void tm_mb_load(const tm_u64 *slot, tm_u64 &p0, tm_u64 &p1) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(p0), "=l"(p1) : "l"(slot) : "memory");
}
#else
inline void tm_mb_store(tm_u64 *slot, const tm_u64 p0, const tm_u64 p1) { slot[0] = p0; slot[1] = p1; }
inline void tm_mb_load(const tm_u64 *slot, tm_u64 &p0, tm_u64 &p1) { p0 = slot[0]; p1 = slot[1]; }
#endif

// Slot of one lane's node of warp slot (row Ap, column group g) at level lam; coordinates with a prime are
// counted in the sweep's direction.  Same offsets as v2_load, computed per slot.  off: slot in the field buffers
// (< 0: the lane has no node at this level), mb: slot inside its slab = mailbox slot.
template <int SA, int SW, int SC>
EIK_HD void tm_addr(const Plan2 &P, const int lane, const int lam, const int Ap, const int g, int &off, int &mb) {
    const int Cp = g * TM_LC + lane;
    const int Wp = lam - Ap - Cp;
    const bool act = (unsigned)Wp < (unsigned)P.dW && Cp < P.dC;
    const int A = SA > 0 ? Ap : P.dA - 1 - Ap;
    const int C = SC > 0 ? Cp : P.dC - 1 - Cp;
    const int mu = SW > 0 ? lam - Ap : P.nmu - 1 - (lam - Ap);      // uniform over the warp: one skewed row
    mb = (mu + 1) * P.PC + C;
    off = act ? (A + 1) * P.RS * P.PC + mb : -1;
}

// OLD values of the node: nothing in the sweep writes them before level lam (own) / lam + 1 (downwind).
struct TmOld {
    double own, fv, dA, dW, dC;
    double ref;   // CMP sweeps: the round-start value of the node
};

template <int SA, int SW, int SC, bool CMP>
EIK_HD void tm_load_old(const Plan2 &P, const int lane, const int lam, const int Ap, const int g, const double *rd,
                        const double *__restrict__ fl, const double *cmp, TmOld &O) {
    const int offA = SA * P.RS * P.PC, offW = SW * P.PC, offC = SW * P.PC + SC;   // downwind (old, level+1)
    int off, mb;
    tm_addr<SA, SW, SC>(P, lane, lam, Ap, g, off, mb);
    // a lane without a node loads from a harmless slot (A = 0, mu = 0: all neighbour slots exist)
    if (off < 0) off = (P.RS + 1) * P.PC + 1;
    const double *p = rd + off;
    O.own = TM_LDU(p);
    O.fv = fl[off];
    O.dA = TM_LDU(p + offA);
    O.dW = TM_LDU(p + offW);
    O.dC = TM_LDU(p + offC);
    O.ref = CMP ? TM_LDU(cmp + off) : 0.0;
}

// L1 prefetch of what tm_load_old will read at level lam (one 32-byte sector per 4 lanes; the lanes of a
// slot read contiguous runs, so a few lanes would do, but predicating them costs as much as issuing all).
template <int SA, int SW, int SC, bool CMP>
EIK_HD void tm_prefetch_old(const Plan2 &P, const int lane, const int lam, const int Ap, const int g, const double *rd,
                            const double *__restrict__ fl, const double *cmp) {
    const int offA = SA * P.RS * P.PC, offW = SW * P.PC;
    int off, mb;
    tm_addr<SA, SW, SC>(P, lane, lam, Ap, g, off, mb);
    if (off < 0) return;
    TM_PREFETCH(rd + off + offW);      // row of level lam+1 in this slab: dW, dC (own was dW one level ago)
    TM_PREFETCH(rd + off + offA);
    TM_PREFETCH(fl + off);
    if (CMP) TM_PREFETCH(cmp + off);
}

// NEW (level lam-1) values of the node of lane column Cp in the CTA's row r: from the sheet of level lam-1
// (pitch SP, column Cp at index Cp+1, index 0 = +inf), the A neighbour of the CTA's first row from the mailbox
// (inbox != nullptr; packets tagged tag_in) or +inf (the grid's first row).
// Host build: a packet that has not arrived yields NaN (the emulation's scheduler must prevent that).
EIK_HD void tm_load_new(const double *sheetPrev, const int SP, const int r, const int Cp, const bool act,
                        const tm_u64 *inbox, const int mb, const unsigned tag_in, double &uA, double &uW, double &uC) {
    const double *row = sheetPrev + r * SP + Cp;
    uC = row[0];
    uW = row[1];
    if (r > 0) {
        uA = row[1 - SP];
    } else if (inbox) {
        uA = 0.0;
        if (act) {
            // the A-neighbour has the same (W, C), i.e. the same slot inside ITS slab
            tm_u64 p0, p1;
#if defined(__CUDA_ARCH__)
            do { tm_mb_load(inbox + 2 * (long long)mb, p0, p1); } while (!tm_unpack(p0, p1, tag_in, uA));
#else
            tm_mb_load(inbox + 2 * (long long)mb, p0, p1);
            if (!tm_unpack(p0, p1, tag_in, uA)) uA = NAN;
#endif
        }
    } else {
        uA = v2_inf();
    }
}

// One warp slot (row r of the CTA = row Ap of the sweep, column group g) at level lam, for one lane, in two
// steps so that the kernel can issue the next level's loads in between (their registers are free once tm_prep
// has consumed O).  base = sweep serial << TM_LEVEL_BITS; a packet of level L carries tag base | (L + 1).
// sheets: [2][R][SP] doubles, sheet (lam & 1) receives this level.
struct TmPrep {
    double a1, a2, a3, own, fv, ref;
    int off, mb;
};

template <int SA, int SW, int SC>
EIK_HD void tm_prep(const Plan2 &P, const TeamCfg &T, const int t, const int lane, const int lam, const int r,
                    const int Ap, const int g, const TmOld &O, const tm_u64 *mbox, const unsigned base,
                    const double *sheets, TmPrep &Q) {
    tm_addr<SA, SW, SC>(P, lane, lam, Ap, g, Q.off, Q.mb);
    // mbox: the team's inboxes, inbox of member t at mbox + t * 2 * mbStride
    const tm_u64 *inbox = (t > 0 && r == 0) ? mbox + (long long)t * 2 * T.mbStride : nullptr;
    const double *prev = sheets + ((lam & 1) ^ 1) * T.R * T.SP;
    double uA, uW, uC;
    tm_load_new(prev, T.SP, r, g * TM_LC + lane, Q.off >= 0, inbox, Q.mb, base | (unsigned)lam, uA, uW, uC);
    Q.own = O.own;
    Q.fv = O.fv;
    Q.ref = O.ref;
    Q.a1 = eik_min(uA, O.dA);
    Q.a2 = eik_min(uW, O.dW);
    Q.a3 = eik_min(uC, O.dC);
    eik_sort3(Q.a1, Q.a2, Q.a3);
}

// The update (Eikonal3D.cpp:47-54), its store, its sheet entry (+inf when the lane has no node) and the packet
// for the downstream CTA (the node is in the CTA's last row; written for EVERY node of the row, changed or not:
// the consumer waits for it).
template <bool OOP, bool CMP>
EIK_HD void tm_solve(const TeamCfg &T, const int t, const int nrow, const int lane, const int lam, const int r,
                     const int g, const TmPrep &Q, double *wr, const double h, double &err, tm_u64 *mbox,
                     const unsigned base, double *sheets) {
    double res = v2_inf();
    if (Q.off >= 0) {
        res = Q.own;
        bool changed = false;
        if (Q.a1 < Q.own) {   // otherwise the candidate (> a1) cannot win the min: exact skip
            const double un = eik_solve3_sorted(Q.a1, Q.a2, Q.a3, Q.fv * h, Q.fv * Q.fv * h * h);
            if (un < Q.own) { res = un; changed = true; }
        }
        if (OOP || changed) wr[Q.off] = res;
        if (t < T.nC - 1 && r == nrow - 1) {
            tm_u64 *outbox = mbox + (long long)(t + 1) * 2 * T.mbStride;
            tm_u64 p0, p1;
            tm_pack(res, base | (unsigned)(lam + 1), p0, p1);
            tm_mb_store(outbox + 2 * (long long)Q.mb, p0, p1);
        }
        if (CMP) {
            const double dd = fabs(res - Q.ref);
            err = (err < dd) ? dd : err;
        }
    }
    sheets[((lam & 1) * T.R + r) * T.SP + g * TM_LC + lane + 1] = res;
}

// rows [a0, a1) of team member t, first and last level at which one of them has a node
EIK_HD void tm_rows(const Plan2 &P, const TeamCfg &T, const int t, int &a0, int &a1, int &lam0, int &lam1) {
    a0 = t * T.R;
    a1 = a0 + T.R < P.dA ? a0 + T.R : P.dA;
    lam0 = a0;
    lam1 = a1 - 1 + P.dW - 1 + P.dC - 1;
}

// warp-uniform: does slot (Ap, g) have a node at level lam?
EIK_HD bool tm_slot_live(const Plan2 &P, const int lam, const int Ap, const int g) {
    const int top = lam - Ap - g * TM_LC;          // W' of lane 0; lane j has W' = top - j
    return top >= 0 && top - (TM_LC - 1) < P.dW;
}

#if defined(__CUDACC__)

__device__ __forceinline__ unsigned tm_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Barrier over the nC CTAs of a team.  ctr only grows; epoch is the value it reaches when every CTA has
// arrived (tracked identically by every CTA).
__device__ __forceinline__ void tm_barrier(unsigned *ctr, unsigned &epoch, const int nC) {
    epoch += (unsigned)nC;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (tm_ld_acquire(ctr) < epoch) {}
        __threadfence();
    }
    __syncthreads();
}

// One sweep of team member t.  A warp owns slots q = warp, warp + nw, ... for the whole sweep.
template <int SA, int SW, int SC, bool OOP, bool CMP>
__device__ __forceinline__ void tm_sweep(const Plan2 &P, const TeamCfg &T, const int t, const double *rd, double *wr,
                                         const double *__restrict__ fl, const double *cmp, const double h,
                                         double &err, tm_u64 *mbox, const unsigned base, double *sheets) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int a0, a1, lam0, lam1;
    tm_rows(P, T, t, a0, a1, lam0, lam1);
    const int nrow = a1 - a0, nslot = nrow * T.G32;
    for (int i = threadIdx.x; i < 2 * T.R * T.SP; i += blockDim.x) sheets[i] = v2_inf();
    __syncthreads();
    for (int lam = lam0; lam <= lam1; lam++) {
        for (int q = warp; q < nslot; q += nw) {
            const int r = q / T.G32, g = q - r * T.G32;
            if (tm_slot_live(P, lam + 1, a0 + r, g))      // warp-uniform; beyond lam1 no slot is live
                tm_prefetch_old<SA, SW, SC, CMP>(P, lane, lam + 1, a0 + r, g, rd, fl, cmp);
            if (!tm_slot_live(P, lam, a0 + r, g)) {       // no node: the sheet still needs its +inf
                sheets[((lam & 1) * T.R + r) * T.SP + g * TM_LC + lane + 1] = v2_inf();
                continue;
            }
            TmOld O;
            TmPrep Q;
            tm_load_old<SA, SW, SC, CMP>(P, lane, lam, a0 + r, g, rd, fl, cmp, O);
            tm_prep<SA, SW, SC>(P, T, t, lane, lam, r, a0 + r, g, O, mbox, base, sheets, Q);
            tm_solve<OOP, CMP>(T, t, nrow, lane, lam, r, g, Q, wr, h, err, mbox, base, sheets);
        }
        __syncthreads();
    }
}

// bufs: S x 3 x M doubles as in k_fwd3d_v2 (buffer 0 of every source: u0 in layout P; every slot that is
// not a grid node: +inf in all three buffers).  grid = S x nC CTAs, ALL co-resident (cooperative launch).
// sync: S x TM_SYNC_WORDS unsigned words, zero on entry.  mbox: S x nC inboxes of mbStride packets; no tag in
// it is >= (serial0 + 1) << TM_LEVEL_BITS (the host hands out serial ranges and clears the mailbox on wrap).
// Dynamic shared memory: max(re-skew plane, two sheets of R x SP doubles); they are never live together.
template <int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB) k_fwd3d_team(const Plan2 P, const TeamCfg T, double *bufs,
                                                            const double *__restrict__ fP, const double *__restrict__ fM,
                                                            const double h, const double tol, const int max_rounds,
                                                            int *__restrict__ rounds, double *__restrict__ errs,
                                                            int *__restrict__ where, unsigned *sync, tm_u64 *mbox_all,
                                                            const unsigned serial0) {
    extern __shared__ double plane[];
    __shared__ double red[32];
    const int src = blockIdx.x / T.nC, t = blockIdx.x - src * T.nC;
    unsigned *sy = sync + (long long)src * TM_SYNC_WORDS;
    unsigned *ctr = sy;
    unsigned long long *errslot = (unsigned long long *)(sy + 2);
    tm_u64 *mbox = mbox_all + (long long)src * T.nC * 2 * T.mbStride;
    unsigned epoch = 0, serial = serial0;
    double *B3 = bufs + (long long)src * 3 * P.M;
    double *Bz = B3 + 2 * P.M;
    const int A0 = t * T.R, A1 = A0 + T.R < P.dA ? A0 + T.R : P.dA;   // slabs this CTA re-skews
    int o = 0, a = 1, r = 0;
    bool conv = false;
    while (r < max_rounds) {
        double err = 0.0;
        double *Bo = B3 + o * P.M, *Ba = B3 + a * P.M;
        int state = 1;
        double *w = Ba;
        for (int sw = 0; sw < 8; sw++) {
            const int sigma = P.sg[sw][1] * P.sg[sw][2];
            if (sw > 0 && sigma != state) {
                double *dst = state > 0 ? Bz : Ba;
                tm_barrier(ctr, epoch, T.nC);         // the sweep that wrote w is complete everywhere
                v2_reskew(P, w, dst, state, plane, A0, A1);
                w = dst;
                state = sigma;
            }
            tm_barrier(ctr, epoch, T.nC);             // previous sweep / re-skew complete everywhere
            serial++;
            const unsigned base = serial << TM_LEVEL_BITS;
#define TM_CALL(a_, w_, c_, oop_, cmp_) \
    tm_sweep<a_, w_, c_, oop_, cmp_>(P, T, t, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err, mbox, base, plane)
            V2_DISPATCH(P, sw, TM_CALL);
#undef TM_CALL
        }
        const double eb = v2_block_max(err, red);
        if (threadIdx.x == 0) atomicMax(errslot + (r & 1), (unsigned long long)__double_as_longlong(eb));   // eb >= 0
        tm_barrier(ctr, epoch, T.nC);
        const double e = __longlong_as_double((long long)__ldcg(errslot + (r & 1)));
        if (t == 0 && threadIdx.x == 0) {
            errslot[(r + 1) & 1] = 0ULL;              // next round's slot: last read one round (>= 9 barriers) ago
            if (errs) errs[(long long)src * max_rounds + r] = e;
        }
        r++;
        const int oo = o; o = a; a = oo;
        if (e < tol) { conv = true; break; }          // e is team-uniform
    }
    if (t == 0 && threadIdx.x == 0) {
        if (rounds) rounds[src] = conv ? r : -r;
        where[src] = o;
    }
}

__global__ void k2_identity(int *__restrict__ p, const int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

#endif  // __CUDACC__

// Plan for the team kernel: 1 x 32 lane patch, any number of column groups.
inline bool team_build_plan(Plan2 &P, int m, int n, int l, int nwarps, size_t plane_bytes) {
    if (!v2_build_plan(P, m, n, l, nwarps, plane_bytes, 1, TM_LC, 1 << 30)) return false;
    return P.nlev < (1 << TM_LEVEL_BITS) - 1;
}

// Team shape for S sources on a device that can hold max_ctas CTAs at once.  R rows per CTA: as few as the
// CTA budget allows -- a level is latency-bound (one ~250-instruction update per lane), so more CTAs with fewer
// slots each win at every size measured (64^3: R = 1 3.3 ms, R = 8 4.2 ms; 256^3: 19.5 / 19.9 / 35.3 ms for
// R = 1 / 2 / 4).  Rforce > 0 overrides (tuning aid).
inline bool team_config(const Plan2 &P, int S, int max_ctas, int nwarps, int Rforce, TeamCfg &T) {
    const int budget = max_ctas / S;
    if (budget < 1) return false;
    T.G32 = (P.dC + TM_LC - 1) / TM_LC;
    int R = (P.dA + budget - 1) / budget;
    (void)nwarps;
    if (Rforce > 0 && (P.dA + Rforce - 1) / Rforce <= budget) R = Rforce;
    T.R = R;
    T.nC = (P.dA + R - 1) / R;
    T.mbStride = (long long)P.RS * P.PC;
    T.SP = TM_LC * T.G32 + 2;
    return true;
}

}  // namespace adtomo

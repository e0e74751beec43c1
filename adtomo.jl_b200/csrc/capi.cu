// capi.cu -- the C ABI of libadtomo_b200.so (declared in include/adtomo_b200.h).
// Host-side orchestration only: argument checks, workspace, staging copies, kernel launches.
// There is deliberately no CPU implementation of any solver in this library.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/adtomo_b200.h"
#include "kernels_v0.cuh"
#include "kernels_adj_topo.cuh"
#include "kernels_adj_sparse.cuh"
#include "kernels_fwd_v1.cuh"
#include "kernels_fwd_v2.cuh"
#include "kernels_fwd_v3.cuh"
#include "kernels_fwd_v4.cuh"
#include "kernels_fwd_team.cuh"
#include "kernels_adj_team.cuh"
#include "kernels_model.cuh"
#include "nccl_dyn.h"

using namespace adtomo;

static thread_local std::string g_err = "";

static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ADTOMO_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

// state of an open adtomo_model_begin .. adtomo_model_finish evaluation (model_api.inc)
struct ModelState {
    bool open = false;
    int m = 0, n = 0, l = 0;
    long long N = 0;
    double *fvar = nullptr, *sig = nullptr, *g_fvar = nullptr, *acc = nullptr;   // acc[0]: loss so far, acc[1]: last d misfit / d scale
    int phases = 0;
    int flags = 0;                // worst positive status of the phases (NOT_CONVERGED / ADJOINT_FLAGGED)
};

struct adtomo_ctx {
    int device = 0;
    ModelState model;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long launches = 0;
    std::map<std::string, std::pair<void *, size_t>> ws;
    std::mutex mu;
    // per-phase device timing of the last call: pairs of events around the kernels of each phase
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<std::pair<int, int>> ev_used;   // (phase, pool index)
    int phase_acc = 0;                          // adtomo_phase_accumulate: keep the event pairs of earlier calls
    std::vector<struct PlanCache *> plans;      // level-major layout plans, one per grid shape
    int fwd_variant = 0;                        // tuning aid: ADTOMO_FWD_VARIANT selects <threads, nodes per lane>
    NcclApi::Comm nccl_comm = nullptr;          // set by adtomo_nccl_init
    int nccl_rank = 0, nccl_size = 1;
    int force_cluster = 0;                      // testing aid: ADTOMO_FORCE_CLUSTER=2|4|8 splits every source over a cluster
    int force_v0 = 0;                           // debugging aid: ADTOMO_FORCE_V0=1 selects the row-major kernel
    int force_v1 = 0;                           // debugging aid: ADTOMO_FORCE_V1=1 selects the level-major kernel
    int force_v2 = 0;                           // debugging aid: ADTOMO_FORCE_V2=1 selects the skewed-pencil kernel for any batch
    bool oom = false;                           // a workspace allocation failed since the flag was last cleared
    std::map<std::pair<long long, int>, int> chunk_cache;   // sources per chunk of the fused step, per (grid, batch)
    int v2_pairing = 1;                         // tuning aid: ADTOMO_V2_PAIRING=0 keeps sources in caller order
    std::map<std::tuple<const void *, int, long long>, int *> v2_spent;   // rounds per source of earlier calls, per batch (plan, S, batch id)
    long long batch_id = 0;                     // adtomo_set_batch_id: names the source batch of the following calls
    int batch_chunk = 0;                        // chunk of that batch the fused step is working on (every chunk has its own placement memo)
    const char *last_fwd_kernel = "";           // name of the last 3D forward sweep kernel launched (adtomo_last_forward_kernel)
    int v3_mode = 1;                            // ADTOMO_V3: 1 (default) batch sweeps of kernels_fwd_v3.cuh with menu pitch, 2 same with run-time pitch, 0 the round-1 sweep loop (cross-check)
    int v4_mode = 0;                            // ADTOMO_V4: 1 the slot-block sweep of kernels_fwd_v4.cuh where the grid allows it (bit-exact, measured slower: 2.4x the DRAM reads), 0 (default) never
    int v3_staged = -1;                         // ADTOMO_V3_STAGED: cp.async look-ahead through shared memory: -1 automatic (one CTA per SM), 0 never, 1 always
    int v2_occ = 0;                             // tuning aid: ADTOMO_V2_OCC caps the CTAs per SM of the skewed-pencil kernel
    std::vector<struct Plan2Cache *> plans2;    // skewed-pencil plans, one per grid shape
    // the +inf padding of the skewed-pencil field buffers is written once per (buffer, plan, sources)
    void *v2_pad_ptr = nullptr;
    size_t v2_pad_bytes = 0;
    const struct Plan2Cache *v2_pad_plan = nullptr;
    int v2_pad_S = 0;
    // team kernel (kernels_fwd_team.cuh): few sources, many CTAs per source
    int team_mode = -1;                         // ADTOMO_TEAM: -1 automatic, 0 never, 1 whenever a team shape exists
    int team_rows = 0;                          // tuning aid: ADTOMO_TEAM_R forces the rows per CTA
    std::vector<struct Plan2Cache *> plans_team;
    void *team_pad_ptr = nullptr;
    size_t team_pad_bytes = 0;
    const struct Plan2Cache *team_pad_plan = nullptr;
    int team_pad_S = 0;
    // mailbox tags only grow (kernels_fwd_team.cuh): next free sweep serial; the mailbox is cleared when the
    // buffer changes or the 20-bit serial space is used up
    unsigned team_serial = 0;
    unsigned team_serial_start = 0;             // testing aid (ADTOMO_TEAM_SERIAL0): first serial after the first allocation
    void *team_mbox_ptr = nullptr;
    size_t team_mbox_bytes = 0;
    int team_nt = 512;                          // 16 warps, <= 64 registers: two CTAs per SM
    int coop_launch = 1;                        // cudaDevAttrCooperativeLaunch; without it the team kernels are never selected
    int adj_team = 0;                           // tuning aid: ADTOMO_ADJ_TEAM = CTAs per source of the adjoint wavefront (0: automatic, 1: single-CTA kernel)
    int adj_sparse = -1;                        // ADTOMO_ADJ_SPARSE: active-set adjoint (kernels_adj_sparse.cuh): -1 automatic (batches whose right-hand side is known to be sparse: the fused step), 0 never, 1 every batch
};

struct Plan2Cache {
    int m, n, l;
    bool ok;
    Plan2 plan;
    size_t smem_bytes;
    // batch kernel (kernels_fwd_v3.cuh): compile-time row pitch of the instantiation (0: run-time pitch), slot table
    int pct = 0, tabOffset = 0, maxPer = 0;
    size_t smemNS = 0;          // the plain (not cp.async staged) sweep: plane + slot table only
    int tabOffsetNS = 0;
    bool v3 = false;
};

struct PlanCache {
    int m, n, l;
    HostPlan hp;
    Plan3 dev;          // same as hp.plan but with DEVICE table pointers
    int *d_tables = nullptr;
    size_t smem_bytes = 0;
};

enum { PH_FWD = 0, PH_MISFIT = 1, PH_ADJ_SETUP = 2, PH_ADJ_SWEEP = 3, PH_ADJ_FINISH = 4, PH_CONVERT = 5, PH_COUNT = 6 };

static void phase_reset(adtomo_ctx *c) { if (!c->phase_acc) c->ev_used.clear(); }
static int phase_begin(adtomo_ctx *c, int phase) {
    size_t k = c->ev_used.size();
    if (k >= c->ev_pool.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
        c->ev_pool.push_back({a, b});
    }
    c->ev_used.push_back({phase, (int)k});
    cudaEventRecord(c->ev_pool[k].first, c->stream);
    return (int)k;
}
static void phase_end(adtomo_ctx *c, int k) {
    if (k >= 0) cudaEventRecord(c->ev_pool[k].second, c->stream);
}

// grow-only named device buffers
static int ws_get(adtomo_ctx *c, const char *name, size_t bytes, void **out) {
    auto &slot = c->ws[name];
    if (slot.second < bytes) {
        if (slot.first) {
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaFree(slot.first));
            slot.first = nullptr;
            slot.second = 0;
        }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&slot.first, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&slot.first, want);
            if (e != cudaSuccess) {
                // less free memory than when the chunk size of the fused step was derived: forget the cached sizes, the
                // caller's retry (adtomo_eikonal3d_misfit_grad does one itself) re-derives them from cudaMemGetInfo
                cudaGetLastError();
                slot.first = nullptr;
                c->chunk_cache.clear();
                c->oom = true;
                return fail(ADTOMO_ERR_CUDA, "out of device memory: workspace '%s' needs %zu bytes", name, want);
            }
        }
        slot.second = want;
    }
    *out = slot.first;
    return 0;
}
#define WS(ctx, name, type, count, ptr)                                              \
    do {                                                                             \
        void *p__;                                                                   \
        int rc__ = ws_get(ctx, name, sizeof(type) * (size_t)(count), &p__);          \
        if (rc__) return rc__;                                                       \
        ptr = (type *)p__;                                                           \
    } while (0)

extern "C" int adtomo_nccl_finalize(adtomo_ctx *c);
extern "C" const char *adtomo_last_error(void) { return g_err.c_str(); }
extern "C" int adtomo_version(void) { return 100; }

extern "C" int adtomo_create(adtomo_ctx **out, int device) {
    if (!out) return fail(ADTOMO_ERR_ARG, "adtomo_create: null out pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ADTOMO_ERR_CUDA, "adtomo_create: no CUDA device (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= ndev) return fail(ADTOMO_ERR_ARG, "adtomo_create: device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    adtomo_ctx *c = new adtomo_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    const char *fv0 = getenv("ADTOMO_FORCE_V0");
    c->force_v0 = (fv0 && fv0[0] == '1');
    const char *fv1 = getenv("ADTOMO_FORCE_V1");
    c->force_v1 = (fv1 && fv1[0] == '1');
    const char *fv2 = getenv("ADTOMO_FORCE_V2");
    c->force_v2 = (fv2 && fv2[0] == '1');
    const char *vpair = getenv("ADTOMO_V2_PAIRING");
    c->v2_pairing = vpair ? atoi(vpair) : 1;
    const char *vocc = getenv("ADTOMO_V2_OCC");
    c->v2_occ = vocc ? atoi(vocc) : 0;
    const char *v3m = getenv("ADTOMO_V3");
    c->v3_mode = v3m ? atoi(v3m) : 1;
    const char *v3s = getenv("ADTOMO_V3_STAGED");
    c->v3_staged = v3s ? atoi(v3s) : -1;
    const char *v4m = getenv("ADTOMO_V4");
    c->v4_mode = v4m ? atoi(v4m) : 0;
    const char *fvv = getenv("ADTOMO_FWD_VARIANT");
    c->fwd_variant = fvv ? atoi(fvv) : 0;
    const char *fcl = getenv("ADTOMO_FORCE_CLUSTER");
    c->force_cluster = fcl ? atoi(fcl) : 0;
    const char *tm = getenv("ADTOMO_TEAM");
    c->team_mode = tm ? atoi(tm) : -1;
    {
        // the team kernels spin on each other: they need all their CTAs co-resident (cooperative launch)
        int coop = 0;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess) { cudaGetLastError(); coop = 0; }
        c->coop_launch = coop;
        if (!coop) c->team_mode = 0;
    }
    const char *atm = getenv("ADTOMO_ADJ_TEAM");
    c->adj_team = atm ? atoi(atm) : 0;
    const char *asp = getenv("ADTOMO_ADJ_SPARSE");
    c->adj_sparse = asp ? atoi(asp) : -1;
    const char *ts0 = getenv("ADTOMO_TEAM_SERIAL0");      // testing aid: start the mailbox tag serial near its wrap
    c->team_serial_start = ts0 ? (unsigned)strtoul(ts0, nullptr, 10) : 0u;
    const char *tmr = getenv("ADTOMO_TEAM_R");
    c->team_rows = tmr ? atoi(tmr) : 0;
    *out = c;
    return 0;
}

extern "C" int adtomo_destroy(adtomo_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &kv : c->ws)
        if (kv.second.first) cudaFree(kv.second.first);
    for (auto *pc : c->plans) {
        if (pc->d_tables) cudaFree(pc->d_tables);
        delete pc;
    }
    if (c->nccl_comm) adtomo_nccl_finalize(c);
    for (auto *pc : c->plans2) delete pc;
    for (auto *pc : c->plans_team) delete pc;
    for (auto &kv : c->v2_spent) cudaFree(kv.second);
    for (auto &pr : c->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int adtomo_synchronize(adtomo_ctx *c) {
    if (!c) return fail(ADTOMO_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" unsigned long long adtomo_stream(adtomo_ctx *c) { return c ? (unsigned long long)(uintptr_t)c->stream : 0ULL; }
extern "C" long long adtomo_launch_count(adtomo_ctx *c) { return c ? c->launches : 0; }
extern "C" double adtomo_last_kernel_ms(adtomo_ctx *c) {
    if (!c) return -1.0;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->timed) return -1.0;
    cudaSetDevice(c->device);
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

extern "C" int adtomo_phase_accumulate(adtomo_ctx *c, int on) {
    if (!c) return fail(ADTOMO_ERR_ARG, "adtomo_phase_accumulate: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    c->ev_used.clear();
    c->phase_acc = on ? 1 : 0;
    return 0;
}

extern "C" double adtomo_last_phase_ms(adtomo_ctx *c, int phase) {
    if (!c) return -1.0;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    double tot = 0.0;
    for (auto &u : c->ev_used) {
        if (u.first != phase) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_pool[u.second].first, c->ev_pool[u.second].second) == cudaSuccess) tot += ms;
        else cudaGetLastError();
    }
    return tot;
}

static adtomo_ctx *default_ctx(int *rc) {
    static thread_local adtomo_ctx *c = nullptr;
    *rc = 0;
    if (!c) *rc = adtomo_create(&c, -1);
    return c;
}

static int check_launch(adtomo_ctx *c, const char *what) {
    c->launches++;
    if (!strncmp(what, "k_fwd3d", 7)) c->last_fwd_kernel = what;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ADTOMO_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return 0;
}
#define LAUNCHED(c, what)                     \
    do {                                      \
        int rc__ = check_launch(c, what);     \
        if (rc__) return rc__;                \
    } while (0)

// ---------------------------------------------------------------------------------------
// self-test of the call-free device sqrt (eik_core.h): bit-for-bit against the CUDA library's sqrt
// ---------------------------------------------------------------------------------------
__global__ void k_selftest_sqrt(const long long n, const unsigned long long seed, unsigned long long *bad) {
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < n; id += (long long)gridDim.x * blockDim.x) {
        // splitmix64
        unsigned long long z = seed + 0x9e3779b97f4a7c15ULL * (unsigned long long)(id + 1);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        z ^= z >> 31;
        double x;
        const int kind = (int)(id & 7);
        if (kind < 3) x = __longlong_as_double((long long)z);                                   // any bit pattern
        else if (kind < 6) x = __longlong_as_double((long long)((z & 0x000fffffffffffffULL) |     // exponents around 1
                                                    ((0x3f0ULL + ((z >> 52) & 0x1f)) << 52)));
        else if (kind == 6) x = __longlong_as_double((long long)(z & 0x001fffffffffffffULL));    // subnormal / tiny
        else {
            const double sp[8] = {0.0, -0.0, 1.0, 4.0, __longlong_as_double(0x7ff0000000000000LL),
                                  __longlong_as_double(0xfff0000000000000LL), __longlong_as_double(0x7ff8000000000000LL), -1.0};
            x = sp[(id >> 3) & 7];
        }
        const double a = eik_sqrt(x), b = sqrt(x);
        const bool same = (__double_as_longlong(a) == __double_as_longlong(b)) || (a != a && b != b);
        if (!same) atomicAdd(bad, 1ULL);
    }
}

extern "C" const char *adtomo_last_forward_kernel(adtomo_ctx *c) { return c ? c->last_fwd_kernel : ""; }

extern "C" int adtomo_set_batch_id(adtomo_ctx *c, long long id) {
    if (!c) return fail(ADTOMO_ERR_ARG, "adtomo_set_batch_id: null context");
    std::lock_guard<std::mutex> lk(c->mu);
    c->batch_id = id;
    return 0;
}

extern "C" int adtomo_selftest_sqrt(adtomo_ctx *c, long long n, unsigned long long seed, long long *mismatches) {
    if (!c || !mismatches) return fail(ADTOMO_ERR_ARG, "adtomo_selftest_sqrt: null argument");
    CK(cudaSetDevice(c->device));
    unsigned long long *d;
    WS(c, "selftest", unsigned long long, 1, d);
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), c->stream));
    k_selftest_sqrt<<<c->num_sms * 8, 256, 0, c->stream>>>(n, seed, d);
    LAUNCHED(c, "k_selftest_sqrt");
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *mismatches = (long long)h;
    return 0;
}

// device pointer for an input: either the caller's (DEVICE) or a staged copy (HOST)
template <typename T>
static int stage_in(adtomo_ctx *c, const char *name, const T *src, size_t count, int loc, const T **out) {
    if (loc == ADTOMO_DEVICE) { *out = src; return 0; }
    T *d;
    WS(c, name, T, count, d);
    CK(cudaMemcpyAsync(d, src, sizeof(T) * count, cudaMemcpyHostToDevice, c->stream));
    *out = d;
    return 0;
}

static int elem_grid(adtomo_ctx *c, long long n, int nt = 256) {
    long long b = (n + nt - 1) / nt;
    long long cap = (long long)c->num_sms * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static size_t free_bytes() {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return (size_t)8 << 30; }
    return fr;
}

// ---------------------------------------------------------------------------------------
// 3D forward on device-resident data (U holds u0 on entry)
// ---------------------------------------------------------------------------------------
static constexpr int NT3 = 512;

static constexpr size_t SMEM_MAX_DYN = 227 * 1024 - 2048;

static int get_plan(adtomo_ctx *c, int m, int n, int l, PlanCache **out) {
    for (auto *pc : c->plans)
        if (pc->m == m && pc->n == n && pc->l == l) { *out = pc; return 0; }
    PlanCache *pc = new PlanCache();
    pc->m = m; pc->n = n; pc->l = l;
    if (!build_plan(pc->hp, m, n, l)) { delete pc; return fail(ADTOMO_ERR_ARG, "internal: layout plan construction failed for %dx%dx%d", m, n, l); }
    // device copies of the tables: [rowIndex | fcum] as ints, then tOf as 16-bit
    size_t total = 0, total16 = 0;
    for (int q = 0; q < NLAYOUT; q++) {
        total += pc->hp.lay[q].rowIndex.size() + pc->hp.lay[q].fcum.size();
        total16 += pc->hp.lay[q].tOf.size();
    }
    const size_t bytes = sizeof(int) * total + sizeof(unsigned short) * total16;
    std::vector<unsigned char> host(bytes);
    CK(cudaMalloc(&pc->d_tables, bytes));
    pc->dev = pc->hp.plan;
    int *hi = (int *)host.data();
    unsigned short *hs = (unsigned short *)(host.data() + sizeof(int) * total);
    int *di = pc->d_tables;
    unsigned short *ds = (unsigned short *)((unsigned char *)pc->d_tables + sizeof(int) * total);
    size_t o = 0, o16 = 0;
    for (int q = 0; q < NLAYOUT; q++) {
        auto &H = pc->hp.lay[q];
        memcpy(hi + o, H.rowIndex.data(), sizeof(int) * H.rowIndex.size());
        pc->dev.lay[q].rowIndex = di + o;
        o += H.rowIndex.size();
        memcpy(hi + o, H.fcum.data(), sizeof(int) * H.fcum.size());
        pc->dev.lay[q].fcum = di + o;
        o += H.fcum.size();
        memcpy(hs + o16, H.tOf.data(), sizeof(unsigned short) * H.tOf.size());
        pc->dev.lay[q].tOf = ds + o16;
        o16 += H.tOf.size();
    }
    CK(cudaMemcpy(pc->d_tables, host.data(), bytes, cudaMemcpyHostToDevice));
    pc->smem_bytes = 0;
    c->plans.push_back(pc);
    *out = pc;
    return 0;
}

// Shared-memory need of k_fwd3d_v1 when a source is split over a cluster of CS CTAs.
struct FwdCfg { int CS; int sheet; int tOfSmem; int barsOffset; size_t smem; };
static bool fwd_config(const PlanCache *pc, int CS, FwdCfg *out, int NS = 1) {
    int sheet = 0, ris = 0, fcLen = 0, tLen = 0;
    for (int q = 0; q < NLAYOUT; q++) {
        const LayoutDev &L = pc->dev.lay[q];
        if (L.dB > 256 * 32) return false;
        if (L.dA < CS) return false;
        sheet = std::max(sheet, ((L.dA + CS - 1) / CS + 2) * L.pitch);
        ris = std::max(ris, L.nlev + 1);
        fcLen = std::max(fcLen, L.dB + L.dC);
        tLen = std::max(tLen, L.dB * L.dC);
    }
    size_t base = sizeof(double) * 2 * NS * (size_t)sheet + sizeof(int) * ((size_t)NLAYOUT * ris + fcLen) + 16;
    if (base > SMEM_MAX_DYN) return false;
    out->CS = CS;
    out->sheet = sheet;
    // packed-row table in shared memory: 8-bit when every row index fits a byte, else 16-bit, else global.
    // (Shared memory is carved out of the 228 KB L1 in steps, so the smaller table also leaves more L1.)
    if (fcLen - 2 <= 255 && base + (size_t)tLen <= SMEM_MAX_DYN) { out->tOfSmem = 2; out->smem = base + (size_t)tLen; }
    else if (base + 2 * (size_t)tLen <= SMEM_MAX_DYN) { out->tOfSmem = 1; out->smem = base + 2 * (size_t)tLen; }
    else { out->tOfSmem = 0; out->smem = base; }
    out->barsOffset = (int)((out->smem + 7) & ~(size_t)7);   // 4 mbarriers for the cluster halo exchange
    out->smem = out->barsOffset + 64;
    return out->smem <= SMEM_MAX_DYN + 128;
}

template <typename K>
static int launch_fwd(adtomo_ctx *c, K kern, int NT, const FwdCfg &cfg, const PlanCache *pc, double *bufs,
                      const double *flay, double h, double tol, int max_rounds, int S, int *d_rounds, double *d_errs,
                      int *where, double *errPart, int NS = 1) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX_DYN + 128));
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(NT);
    lc.dynamicSmemBytes = cfg.smem;
    lc.stream = c->stream;
    cudaLaunchAttribute at[1];
    int nattr = 0;
    int nsrc = std::min((S + NS - 1) / NS, c->num_sms / cfg.CS);   // concurrent source groups
    if (cfg.CS > 1) {
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cfg.CS;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        nattr = 1;
        lc.attrs = at;
        lc.numAttrs = nattr;
        lc.gridDim = dim3(cfg.CS * nsrc);
        int maxc = 0;
        if (cudaOccupancyMaxActiveClusters(&maxc, kern, &lc) == cudaSuccess && maxc > 0) nsrc = std::min(nsrc, maxc);
        else cudaGetLastError();
    }
    lc.gridDim = dim3(cfg.CS * nsrc);
    lc.attrs = nattr ? at : nullptr;
    lc.numAttrs = nattr;
    CK(cudaLaunchKernelEx(&lc, kern, pc->dev, cfg.sheet, cfg.tOfSmem, cfg.barsOffset, bufs, flay, h, tol, max_rounds, S, d_rounds, d_errs,
                          where, errPart));
    return 0;
}

static Plan2Cache *get_plan2(adtomo_ctx *c, int m, int n, int l) {
    for (auto *pc : c->plans2)
        if (pc->m == m && pc->n == n && pc->l == l) return pc;
    Plan2Cache *pc = new Plan2Cache();
    pc->m = m; pc->n = n; pc->l = l;
    // 16 warps: two CTAs per SM at 64 registers per thread
    const char *vw = getenv("ADTOMO_V2_WARPS");      // tuning aid
    const char *vp = getenv("ADTOMO_V2_PLANE_KB");   // testing aid: a small re-skew plane forces W-chunking
    const size_t plane = vp ? (size_t)atoi(vp) * 1024 : 64 * 1024;
    if (c->v3_mode) {
        pc->ok = v3_build_plan(pc->plan, m, n, l, vw ? atoi(vw) : 16, plane, &pc->pct, c->v3_mode == 1);
        pc->v3 = pc->ok;
    } else {
        pc->ok = v2_build_plan(pc->plan, m, n, l, vw ? atoi(vw) : 16, plane);
    }
    pc->smem_bytes = pc->ok ? sizeof(double) * (size_t)pc->plan.WCH * pc->plan.PS : 0;
    if (pc->v3) {
        pc->maxPer = v3_max_per_warp(pc->plan);
        const size_t tabBytes = (sizeof(V3Slot) + sizeof(int)) * (size_t)(pc->plan.NT / 32) * pc->maxPer;
        pc->tabOffsetNS = (int)((pc->smem_bytes + 15) & ~(size_t)15);      // plain sweep: plane, then the slot table
        pc->smemNS = pc->tabOffsetNS + tabBytes;
        if (pc->pct) pc->smem_bytes = std::max(pc->smem_bytes, (size_t)V3_STAGE_BYTES_PER_WARP * (pc->plan.NT / 32));   // cp.async staging aliases the plane
        pc->tabOffset = (int)((pc->smem_bytes + 15) & ~(size_t)15);
        pc->smem_bytes = pc->tabOffset + tabBytes;
        if (pc->smem_bytes > (size_t)100 * 1024 || pc->plan.NT > 512) {   // table too large for two CTAs per SM, or more than 16 warps: the round-1 sweep loop
            pc->v3 = false;
            pc->ok = v2_build_plan(pc->plan, m, n, l, vw ? atoi(vw) : 16, 64 * 1024);
            pc->smem_bytes = pc->ok ? sizeof(double) * (size_t)pc->plan.WCH * pc->plan.PS : 0;
        }
    }
    c->plans2.push_back(pc);
    return pc;
}

// Skewed-pencil path (kernels_fwd_v2.cuh): convert in, sweep, convert out.
// u0 given as "fill value + sparse source lists" (what the inversion drivers build, inversion.jl:52-60) next to
// its dense form: the skewed-pencil path initialises its own layout from the lists and never reads the dense field.
struct SparseU0 {
    const int *ptr, *idx;       // CSR over the sources of this call: entries ptr[s] .. ptr[s+1]-1 of idx / val
    const double *val;
    double fill;
    const double *dense;        // S x N row-major, the same field
};

static int fwd3d_v2(adtomo_ctx *c, const Plan2Cache *pc, double *dU, const double *df, const Dims3 &d, double h,
                    double tol, int max_rounds, int S, int *d_rounds, double *d_errs, const SparseU0 *sp) {
    const Plan2 &P = pc->plan;
    double *bufs, *flay;
    int *where, *order, *spent;
    const size_t nb = (size_t)S * 3 * P.M;
    // idle lanes of the batch kernel load (never store) up to v3_slack() doubles outside a buffer
    const size_t slack = pc->v3 ? (size_t)v3_slack(P) : 0;
    WS(c, "fwd2_bufs", double, nb + 2 * slack, bufs);
    WS(c, "fwd2_flay", double, (size_t)2 * P.M + 2 * slack, flay);
    bufs += slack;
    flay += slack;
    WS(c, "fwd2_where", int, 2 * (size_t)S, where);
    order = where + S;
    // rounds each source of THIS batch needed last time (batch = plan, size, caller's rounds array)
    {
        const auto key = std::make_tuple((const void *)pc, S, c->batch_id * 4096 + c->batch_chunk);
        auto it = c->v2_spent.find(key);
        if (it == c->v2_spent.end()) {
            if (c->v2_spent.size() > 64) {          // bounded: forget everything
                for (auto &kv : c->v2_spent) cudaFree(kv.second);
                c->v2_spent.clear();
            }
            int *p = nullptr;
            CK(cudaMalloc(&p, sizeof(int) * 2 * (size_t)S));      // [rounds per source | SM per CTA]
            CK(cudaMemsetAsync(p, 0, sizeof(int) * 2 * (size_t)S, c->stream));
            it = c->v2_spent.emplace(key, p).first;
        }
        spent = it->second;
    }
    int pk = phase_begin(c, PH_CONVERT);
    if (c->v2_pad_ptr != (void *)bufs || c->v2_pad_plan != pc || c->v2_pad_S < S || c->v2_pad_bytes != c->ws["fwd2_bufs"].second) {
        // every slot that is not a grid node must hold +inf; nothing ever writes those slots afterwards
        k2_fill<<<c->num_sms * 8, 512, 0, c->stream>>>(bufs, (long long)nb, v2_inf());
        LAUNCHED(c, "k2_fill");
        c->v2_pad_ptr = bufs; c->v2_pad_plan = pc; c->v2_pad_S = S; c->v2_pad_bytes = c->ws["fwd2_bufs"].second;
    }
    const int eb = elem_grid(c, d.N);
    k2_f_to_layouts<<<eb, 256, 0, c->stream>>>(P, df, flay, flay + P.M);
    LAUNCHED(c, "k2_f_to_layouts");
    k2_make_order<<<1, 1024, 0, c->stream>>>(spent, spent + S, S, c->v2_pairing ? c->num_sms : 0, order);
    LAUNCHED(c, "k2_make_order");
    if (sp) {
        k2_fill_valid<<<dim3(32, S), 256, 0, c->stream>>>(P, bufs, sp->fill);
        LAUNCHED(c, "k2_fill_valid");
        k2_scatter_P<<<(S + 127) / 128, 128, 0, c->stream>>>(P, bufs, order, sp->ptr, sp->idx, sp->val, S);
        LAUNCHED(c, "k2_scatter_P");
    } else {
        k2_u0_to_P<<<dim3(std::min(eb, 64), S), 256, 0, c->stream>>>(P, dU, bufs, order);
        LAUNCHED(c, "k2_u0_to_P");
    }
    phase_end(c, pk);
    pk = phase_begin(c, PH_FWD);
#define V2_LAUNCH(NTMAX_, MINB_)                                                                                       \
    do {                                                                                                               \
        auto kern = k_fwd3d_v2<NTMAX_, MINB_>;                                                                         \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));                       \
        int occ = 1;                                                                                                   \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, P.NT, pc->smem_bytes));                           \
        if (occ < 1) occ = 1;                                                                                          \
        if (c->v2_occ > 0 && occ > c->v2_occ) occ = c->v2_occ;                                                         \
        kern<<<std::min(S, c->num_sms * occ), P.NT, pc->smem_bytes, c->stream>>>(P, bufs, flay, flay + P.M, h, tol,    \
                                                                                 max_rounds, S, d_rounds, d_errs, where, order, spent); \
    } while (0)
    const bool ragged = v3_ragged(P);      // a grid extent that is not a multiple of the warp-slot shape: lane masks needed
    static const int v3_carveout = getenv("ADTOMO_V3_CARVEOUT") ? atoi(getenv("ADTOMO_V3_CARVEOUT")) : -1;   // tuning aid: % of L1/shared for shared memory
#define V3_LAUNCH(PCT_, STG_)                                                                                          \
    do {                                                                                                               \
        auto kern = ragged ? k_fwd3d_v3<512, 2, PCT_, STG_, true> : k_fwd3d_v3<512, 2, PCT_, STG_, false>;             \
        const size_t smem__ = STG_ ? pc->smem_bytes : pc->smemNS;                                                      \
        const int tabOff__ = STG_ ? pc->tabOffset : pc->tabOffsetNS;                                                   \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));                       \
        if (v3_carveout >= 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, v3_carveout)); \
        int occ = 1;                                                                                                   \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, P.NT, smem__));                                   \
        if (occ < 1) occ = 1;                                                                                          \
        if (c->v2_occ > 0 && occ > c->v2_occ) occ = c->v2_occ;                                                         \
        kern<<<std::min(S, c->num_sms * occ), P.NT, smem__, c->stream>>>(P, tabOff__, pc->maxPer, bufs, flay, flay + P.M, h, \
                                                                                 tol, max_rounds, S, d_rounds, d_errs, where, order, spent); \
    } while (0)
    // cp.async look-ahead through shared memory pays when a CTA has its SM to itself (148 sources: 100 vs 111 ms) and
    // costs with two CTAs per SM (256 sources: 164 vs 140 ms: the L1 data pipe carries every value twice)
    const bool staged = c->v3_staged == 1 || (c->v3_staged < 0 && S <= c->num_sms);
    // slot-block sweep (kernels_fwd_v4.cuh): 4 x 8 patches, no ragged edge, compile-time pitch, 16 warps
    const bool v4 = pc->v3 && c->v4_mode && pc->pct != 0 && P.NT == 512 && v4_supported(P) && c->v3_staged != 1;
#define V4_LAUNCH(PCT_)                                                                                                \
    do {                                                                                                               \
        auto kern = k_fwd3d_v4<512, 2, PCT_>;                                                                          \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));                       \
        int occ = 1;                                                                                                   \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, P.NT, pc->smem_bytes));                           \
        if (occ < 1) occ = 1;                                                                                          \
        if (c->v2_occ > 0 && occ > c->v2_occ) occ = c->v2_occ;                                                         \
        kern<<<std::min(S, c->num_sms * occ), P.NT, pc->smem_bytes, c->stream>>>(P, pc->tabOffset, bufs, flay, flay + P.M, h, tol, \
                                                                                 max_rounds, S, d_rounds, d_errs, where, order, spent); \
    } while (0)
    if (v4) {
        switch (pc->pct) {
#define V4_CASE(pc_) case pc_: V4_LAUNCH(pc_); break;
            V3_PC_MENU(V4_CASE)
#undef V4_CASE
            default: return fail(ADTOMO_ERR_ARG, "internal: no slot-block kernel for pitch %d", pc->pct);
        }
    }
    else if (pc->v3 && P.NT <= 512) {
        switch (pc->pct) {
#define V3_CASE(pc_) case pc_: if (staged) V3_LAUNCH(pc_, true); else V3_LAUNCH(pc_, false); break;
            V3_PC_MENU(V3_CASE)
#undef V3_CASE
            default: V3_LAUNCH(0, false); break;
        }
    }
    else if (P.NT <= 256) V2_LAUNCH(256, 2);
    else if (P.NT <= 320) V2_LAUNCH(320, 2);
    else if (P.NT <= 384) V2_LAUNCH(384, 2);
    else if (P.NT <= 512) V2_LAUNCH(512, 2);
    else V2_LAUNCH(1024, 1);
#undef V3_LAUNCH
#undef V4_LAUNCH
#undef V2_LAUNCH
    phase_end(c, pk);
    LAUNCHED(c, v4 ? "k_fwd3d_v4" : pc->v3 ? "k_fwd3d_v3" : "k_fwd3d_v2");
    pk = phase_begin(c, PH_CONVERT);
    k2_P_to_rowmajor<<<dim3(std::min(eb, 64), S), 256, 0, c->stream>>>(P, bufs, where, dU, order);
    phase_end(c, pk);
    LAUNCHED(c, "k2_P_to_rowmajor");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Team path (kernels_fwd_team.cuh): few sources, every source on a team of co-resident CTAs.
// ---------------------------------------------------------------------------------------------
static Plan2Cache *get_plan_team(adtomo_ctx *c, int m, int n, int l) {
    for (auto *pc : c->plans_team)
        if (pc->m == m && pc->n == n && pc->l == l) return pc;
    Plan2Cache *pc = new Plan2Cache();
    pc->m = m; pc->n = n; pc->l = l;
    pc->ok = team_build_plan(pc->plan, m, n, l, c->team_nt / 32, 64 * 1024);
    pc->smem_bytes = pc->ok ? sizeof(double) * (size_t)pc->plan.WCH * pc->plan.PS : 0;
    c->plans_team.push_back(pc);
    return pc;
}

// KS: slots per warp whose per-sweep constants stay in registers
// one: the launch has at most one CTA per SM -- the instantiation without the 64-register cap (no spills)
static const void *team_kernel(int KS, bool one = false) {
    if (one) return KS <= 1 ? (const void *)k_fwd3d_team<512, 1, 1> : (const void *)k_fwd3d_team<512, 1, 2>;
    return KS <= 1 ? (const void *)k_fwd3d_team<512, 2, 1> : (const void *)k_fwd3d_team<512, 2, 2>;
}
static size_t team_smem(const Plan2Cache *pc, const TeamCfg &T) {
    return std::max(pc->smem_bytes, sizeof(double) * 2 * (size_t)T.R * T.SP);
}
static int team_ks(const TeamCfg &T, int nt) { return T.R * T.G32 <= nt / 32 ? 1 : 2; }

// CTAs of k_fwd3d_team the device holds at once (cooperative launch limit)
static int team_max_ctas(adtomo_ctx *c, int KS, size_t smem) {
    const void *kern = team_kernel(KS);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) { cudaGetLastError(); return 0; }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, c->team_nt, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return occ * c->num_sms;
}

static int fwd3d_team(adtomo_ctx *c, const Plan2Cache *pc, const TeamCfg &T, double *dU, const double *df, const Dims3 &d,
                      double h, double tol, int max_rounds, int S, int *d_rounds, double *d_errs, const SparseU0 *sp) {
    const Plan2 &P = pc->plan;
    double *bufs, *flay;
    int *where, *order;
    unsigned *sync;
    tm_u64 *mbox;
    const size_t nb = (size_t)S * 3 * P.M;
    const size_t mbn = (size_t)S * T.nC * 2 * (size_t)T.mbStride;
    WS(c, "team_bufs", double, nb, bufs);
    WS(c, "team_flay", double, (size_t)2 * P.M, flay);
    WS(c, "team_where", int, 2 * (size_t)S, where);
    WS(c, "team_sync", unsigned, (size_t)S * T.stride, sync);
    WS(c, "team_mbox", tm_u64, mbn, mbox);
    order = where + S;
    int pk = phase_begin(c, PH_CONVERT);
    if (c->team_pad_ptr != (void *)bufs || c->team_pad_plan != pc || c->team_pad_S < S || c->team_pad_bytes != c->ws["team_bufs"].second) {
        k2_fill<<<c->num_sms * 8, 512, 0, c->stream>>>(bufs, (long long)nb, v2_inf());
        LAUNCHED(c, "k2_fill");
        c->team_pad_ptr = bufs; c->team_pad_plan = pc; c->team_pad_S = S; c->team_pad_bytes = c->ws["team_bufs"].second;
    }
    // serial range of this launch: 8 sweeps per round; no packet in the mailbox may carry a tag of this range
    const unsigned need = 8u * (unsigned)max_rounds + 1u;
    const unsigned serial_max = (1u << (32 - TM_LEVEL_BITS)) - 1u;
    if (need >= serial_max) return fail(ADTOMO_ERR_ARG, "max_rounds %d too large for the team kernel", max_rounds);
    if (c->team_mbox_ptr != (void *)mbox || c->team_mbox_bytes != c->ws["team_mbox"].second || c->team_serial + need >= serial_max) {
        CK(cudaMemsetAsync(mbox, 0, c->ws["team_mbox"].second, c->stream));
        c->team_mbox_ptr = mbox; c->team_mbox_bytes = c->ws["team_mbox"].second;
        c->team_serial = (c->team_serial_start + need < serial_max) ? c->team_serial_start : 0;
        c->team_serial_start = 0;
    }
    unsigned serial0 = c->team_serial;
    c->team_serial += need;
    CK(cudaMemsetAsync(sync, 0, sizeof(unsigned) * (size_t)S * T.stride, c->stream));
    const int eb = elem_grid(c, d.N);
    k2_f_to_layouts<<<eb, 256, 0, c->stream>>>(P, df, flay, flay + P.M);
    LAUNCHED(c, "k2_f_to_layouts");
    k2_identity<<<(S + 127) / 128, 128, 0, c->stream>>>(order, S);
    LAUNCHED(c, "k2_identity");
    if (sp) {
        k2_fill_valid<<<dim3(32, S), 256, 0, c->stream>>>(P, bufs, sp->fill);
        LAUNCHED(c, "k2_fill_valid");
        k2_scatter_P<<<(S + 127) / 128, 128, 0, c->stream>>>(P, bufs, order, sp->ptr, sp->idx, sp->val, S);
        LAUNCHED(c, "k2_scatter_P");
    } else {
        k2_u0_to_P<<<dim3(std::max(1, std::min(eb, 4 * c->num_sms / S)), S), 256, 0, c->stream>>>(P, dU, bufs, order);
        LAUNCHED(c, "k2_u0_to_P");
    }
    phase_end(c, pk);
    pk = phase_begin(c, PH_FWD);
    {
        Plan2 Pv = P;
        TeamCfg Tv = T;
        const double *fPp = flay, *fMp = flay + P.M;
        void *args[] = {&Pv, &Tv, &bufs, &fPp, &fMp, &h, &tol, &max_rounds, &d_rounds, &d_errs, &where, &sync, &mbox, &serial0};
        const bool one = S * T.nC <= c->num_sms && !getenv("ADTOMO_TEAM_CAP64");     // (testing aid: keep the capped kernel)
        const void *kern = team_kernel(team_ks(T, c->team_nt), one);
        if (one) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CK(cudaLaunchCooperativeKernel(kern, dim3(S * T.nC), dim3(c->team_nt), args, team_smem(pc, T), c->stream));
    }
    phase_end(c, pk);
    LAUNCHED(c, "k_fwd3d_team");
    pk = phase_begin(c, PH_CONVERT);
    k2_P_to_rowmajor<<<dim3(std::max(1, std::min(eb, 4 * c->num_sms / S)), S), 256, 0, c->stream>>>(P, bufs, where, dU, order);
    phase_end(c, pk);
    LAUNCHED(c, "k2_P_to_rowmajor");
    return 0;
}

// dU: S x N row-major, holds u0 on entry and the travel times on exit.
static int fwd3d_device(adtomo_ctx *c, double *dU, const double *df, const Dims3 &d, double h, double tol,
                        int max_rounds, int S, int *d_rounds, double *d_errs, const SparseU0 *sp = nullptr) {
    PlanCache *pc = nullptr;
    int rc = get_plan(c, d.m, d.n, d.l, &pc);
    if (rc) return rc;
    FwdCfg cfg;
    bool fits = false;
    for (int cs = std::max(1, c->force_cluster); cs <= 8 && !fits; cs *= 2) fits = fwd_config(pc, cs, &cfg);
    // Kernel choice.  Few sources (every source gets its own SM or cluster of SMs in ONE wave): the
    // level-major kernel, which puts 1024 threads (x cluster size) on a source.  A batch that oversubscribes
    // the SMs: the skewed-pencil kernel (two sources per SM, no shared-memory limit on the grid size).
    // Few sources: a team of CTAs per source (every source gets >= 4 CTAs, or ADTOMO_TEAM=1).  Measured on the
    // 128x128x64 checkerboard batch (benchmarks/batch_probe.py), forward ms for S = 8 / 16 / 37 / 64 / 100 / 148 sources:
    // team 17.6 / 22.7 / 48.7 / 88.1 / 187 / 173, level-major (one SM per source) 82 / 82 / 96 / 96 / 96 / 87,
    // skewed-pencil with one CTA per source 107 / 112 / 129 / 130 / 132 / 122.
    if (!c->force_v0 && !c->force_v1 && !c->force_v2 && !c->force_cluster && c->team_mode != 0) {
        const Plan2Cache *pt = get_plan_team(c, d.m, d.n, d.l);
        TeamCfg T;
        if (pt->ok) {
            // the CTA budget depends on the kernel variant and the shared memory, which depend on the team shape:
            // shape from the budget of the common case, then re-check the shape's own budget
            const int maxc = team_max_ctas(c, 2, pt->smem_bytes);
            if (maxc > 0 && team_config(pt->plan, S, maxc, c->team_nt / 32, c->team_rows, T) && (T.nC >= 4 || c->team_mode == 1) &&
                team_smem(pt, T) <= 100 * 1024 && S * T.nC <= team_max_ctas(c, team_ks(T, c->team_nt), team_smem(pt, T)))
                return fwd3d_team(c, pt, T, dU, df, d, h, tol, max_rounds, S, d_rounds, d_errs, sp);
        }
    }
    if (!c->force_v0 && !c->force_v1 && !c->force_cluster) {
        const bool few = fits && (long long)S * cfg.CS <= c->num_sms;
        if (!few || c->force_v2) {
            const Plan2Cache *p2 = get_plan2(c, d.m, d.n, d.l);
            if (p2->ok) return fwd3d_v2(c, p2, dU, df, d, h, tol, max_rounds, S, d_rounds, d_errs, sp);
        }
    }
    // the kernels below start from the dense field in dU
    if (sp) CK(cudaMemcpyAsync(dU, sp->dense, sizeof(double) * (size_t)S * d.N, cudaMemcpyDeviceToDevice, c->stream));
    if (!c->force_v0 && fits) {
        // level-major path: convert in, sweep, convert out
        double *bufs, *flay, *errPart;
        int *where;
        WS(c, "fwd_bufs", double, (size_t)S * 3 * pc->dev.Mmax, bufs);
        WS(c, "fwd_flay", double, (size_t)NLAYOUT * pc->dev.Mmax, flay);
        WS(c, "fwd_where", int, S, where);
        WS(c, "fwd_errpart", double, (size_t)(S + 4) * 8, errPart);
        int pk = phase_begin(c, PH_CONVERT);
        const int eb = elem_grid(c, d.N);
        k_f_to_layouts<<<eb, 256, 0, c->stream>>>(pc->dev, df, flay);
        LAUNCHED(c, "k_f_to_layouts");
        k_u0_to_L0<<<dim3(std::min(eb, 64), S), 256, 0, c->stream>>>(pc->dev, dU, bufs);
        phase_end(c, pk);
        LAUNCHED(c, "k_u0_to_L0");
        pk = phase_begin(c, PH_FWD);
        if (cfg.CS > 1)
            rc = launch_fwd(c, k_fwd3d_v1<1024, 1, true>, 1024, cfg, pc, bufs, flay, h, tol, max_rounds, S, d_rounds, d_errs, where, errPart);
        else if (c->fwd_variant == 2)
            rc = launch_fwd(c, k_fwd3d_v1<768, 1, false>, 768, cfg, pc, bufs, flay, h, tol, max_rounds, S, d_rounds, d_errs, where, errPart);
        else if (c->fwd_variant == 3)
            rc = launch_fwd(c, k_fwd3d_v1<896, 1, false>, 896, cfg, pc, bufs, flay, h, tol, max_rounds, S, d_rounds, d_errs, where, errPart);
        else if (c->fwd_variant == 1)
            rc = launch_fwd(c, k_fwd3d_v1<1024, 2, false>, 1024, cfg, pc, bufs, flay, h, tol, max_rounds, S, d_rounds, d_errs, where, errPart);
        else
            rc = launch_fwd(c, k_fwd3d_v1<1024, 1, false>, 1024, cfg, pc, bufs, flay, h, tol, max_rounds, S, d_rounds, d_errs, where, errPart);
        if (rc) return rc;
        phase_end(c, pk);
        LAUNCHED(c, "k_fwd3d_v1");
        pk = phase_begin(c, PH_CONVERT);
        k_L0_to_rowmajor<<<dim3(std::min(eb, 64), S), 256, 0, c->stream>>>(pc->dev, bufs, where, dU);
        phase_end(c, pk);
        LAUNCHED(c, "k_L0_to_rowmajor");
        return 0;
    }
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fwd3d_v0<NT3>, NT3, 0));
    if (occ < 1) occ = 1;
    int grid = std::min(S, c->num_sms * occ);
    double *scratch;
    WS(c, "fwd_scratch", double, (size_t)grid * d.N, scratch);
    int pk = phase_begin(c, PH_FWD);
    k_fwd3d_v0<NT3><<<grid, NT3, 0, c->stream>>>(dU, scratch, df, d, h, tol, max_rounds, S, d_rounds, d_errs);
    phase_end(c, pk);
    LAUNCHED(c, "k_fwd3d_v0");
    return 0;
}

// Active-set adjoint for batches (kernels_adj_sparse.cuh): one CTA per source, only the ancestors of the nodes with a
// non-zero right-hand side are visited by the wavefronts.
static int adj3d_sparse(adtomo_ctx *c, const double *dU, const double *dU0, const double *dG, const double *df,
                        double *dGU0, double *dGF, double *dGFsum, const Dims3 &d, double h, int S, int *d_status) {
    double2 *UX, *GD;
    unsigned *W;
    int *Q1, *Q2, *tails;
    const size_t total = (size_t)S * d.N;
    WS(c, "adj_ux", double2, total, UX);
    WS(c, "adj_gd", double2, total, GD);
    WS(c, "adj_w", unsigned, total, W);
    WS(c, "adj_queue", int, total, Q1);
    WS(c, "adj_queue2", int, total, Q2);
    WS(c, "adj_counters", int, 8 * (size_t)S, tails);
    CK(cudaMemsetAsync(tails, 0, sizeof(int) * S, c->stream));
    const dim3 eg(std::min(elem_grid(c, d.N), std::max(128, 16 * c->num_sms / S)), S);
    int pk = phase_begin(c, PH_ADJ_SETUP);
    k_adj3d_setup3<<<eg, 256, 0, c->stream>>>(dU, dU0, dG, UX, GD, dGU0, W, Q1, tails, d, S);
    phase_end(c, pk);
    LAUNCHED(c, "k_adj3d_setup3");
    if (!dGF && !dGFsum) return 0;
    static const int nt_env = getenv("ADTOMO_ADJ_NT") ? atoi(getenv("ADTOMO_ADJ_NT")) : 0;
    const int nt = nt_env ? nt_env : 1024;
    pk = phase_begin(c, PH_ADJ_SWEEP);
#define ASP_LAUNCH(NT_)                                                                                               \
    do {                                                                                                              \
        int occ = 1;                                                                                                  \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_adj3d_sparse<NT_>, NT_, 0));                         \
        if (occ < 1) occ = 1;                                                                                         \
        k_adj3d_sparse<NT_><<<std::min(S, c->num_sms * occ), NT_, 0, c->stream>>>(UX, GD, W, Q1, Q2, tails, d, S, d_status); \
    } while (0)
    if (nt <= 256) ASP_LAUNCH(256);
    else if (nt <= 512) ASP_LAUNCH(512);
    else ASP_LAUNCH(1024);
#undef ASP_LAUNCH
    phase_end(c, pk);
    LAUNCHED(c, "k_adj3d_sparse");
    pk = phase_begin(c, PH_ADJ_FINISH);
    k_adj3d_finish2<<<elem_grid(c, d.N), 256, 0, c->stream>>>(UX, df, dGF, dGFsum, d.N, S, h);
    phase_end(c, pk);
    LAUNCHED(c, "k_adj3d_finish2");
    return 0;
}

// sparse_rhs: the caller knows that grad_u is non-zero at few nodes only (the fused step: receiver cell corners)
static int adj3d_device(adtomo_ctx *c, const double *dU, const double *dU0, const double *dG, const double *df,
                        double *dGU0, double *dGF, double *dGFsum, const Dims3 &d, double h, int S,
                        int *d_status, bool sparse_rhs = false) {
    if (c->adj_sparse == 1 || (c->adj_sparse < 0 && sparse_rhs)) {
        // few sources: the team wavefront spreads one source over many SMs and wins although it visits every node
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_adj3d_topo_team<512>, 512, 0) != cudaSuccess) { cudaGetLastError(); occ = 0; }
        const bool team = c->coop_launch && !(c->team_mode == 0 && c->adj_team == 0) && c->adj_team != 1 && occ * c->num_sms / S >= 8;
        if (!team || c->adj_sparse == 1) return adj3d_sparse(c, dU, dU0, dG, df, dGU0, dGF, dGFsum, d, h, S, d_status);
    }
    double2 *UX, *GD;
    unsigned char *code, *cnt;
    unsigned short *CM;
    int *Q, *cnts;   // cnts: [0,S) number of non-pinned nodes per source, [S,2S) ready-queue tails
    const size_t total = (size_t)S * d.N;
    WS(c, "adj_ux", double2, total, UX);
    WS(c, "adj_gd", double2, total, GD);
    WS(c, "adj_code", unsigned char, total, code);
    WS(c, "adj_cm", unsigned short, total, CM);
    WS(c, "adj_cnt", unsigned char, (total + 7) & ~(size_t)3, cnt);
    WS(c, "adj_queue", int, total, Q);
    WS(c, "adj_counters", int, 8 * (size_t)S, cnts);     // [0,S) free nodes, [S,2S) queue tails, [2S,3S) barrier counters, [4S,8S) push counters
    CK(cudaMemsetAsync(cnts, 0, sizeof(int) * 8 * S, c->stream));
    // elementwise passes: (blocks per source, S) CTAs of 256 threads; few sources need more blocks each to fill the GPU
    const dim3 eg(std::min(elem_grid(c, d.N), std::max(128, 16 * c->num_sms / S)), S);
    int pk = phase_begin(c, PH_ADJ_SETUP);
    k_adj3d_setup2<<<eg, 256, 0, c->stream>>>(dU, dU0, dG, UX, GD, dGU0, code, cnts, d, S);
    LAUNCHED(c, "k_adj3d_setup2");
    if (dGF || dGFsum) {
        k_adj3d_count2<<<eg, 256, 0, c->stream>>>(code, CM, cnt, Q, cnts + S, d, S);
        phase_end(c, pk);
        LAUNCHED(c, "k_adj3d_count2");
        // 1024 threads per source, one source per SM at a time.  (Measured on the 256-source bench batch: 512-thread
        // CTAs at 2-3 per SM 38.8 ms, 256-thread CTAs 42.4 ms, against 35.0 ms: a wave of the ready queue is short
        // and the per-wave barrier cost falls with the number of threads that share it.)  ADTOMO_ADJ_NT: tuning aid.
        static const int adj_nt_env = getenv("ADTOMO_ADJ_NT") ? atoi(getenv("ADTOMO_ADJ_NT")) : 0;
        const int adj_nt = adj_nt_env ? adj_nt_env : 1024;
        static const int adj_agg_env = getenv("ADTOMO_ADJ_AGG") ? atoi(getenv("ADTOMO_ADJ_AGG")) : -1;
        const bool adj_agg = adj_agg_env >= 0 ? adj_agg_env != 0 : true;
        pk = phase_begin(c, PH_ADJ_SWEEP);
        // Few sources: a team of CTAs per source works on every wave (kernels_adj_team.cuh).
        {
            constexpr int ANT = 512;
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_adj3d_topo_team<ANT>, ANT, 0) != cudaSuccess) { cudaGetLastError(); occ = 0; }
            const int budget = occ * c->num_sms / S;
            // waves are short (N / #waves nodes): 64 CTAs are enough up to ~128^3 (64^3: 1.0 ms, 3.2 ms with 148), one CTA per
            // SM pays from 256^3 on (7.6 -> 6.4 ms)
            int nC = c->adj_team > 0 ? std::min(c->adj_team, budget) : std::min(budget, d.N > (1LL << 22) ? c->num_sms : 64);
            if ((c->team_mode == 0 && c->adj_team == 0) || !c->coop_launch) nC = 1;
            if (nC >= 8 || (c->adj_team > 1 && nC > 1)) {
                const int *tail0 = cnts + S, *nfree = cnts;
                int *Dp = cnts + 4 * S;
                unsigned *bar = (unsigned *)(cnts + 2 * S);
                unsigned int *cnt32 = (unsigned int *)cnt;
                Dims3 dd = d;
                void *args[] = {&UX, &GD, &CM, &cnt32, &Q, &tail0, &Dp, &nfree, &dd, &nC, &bar, &d_status};
                static const bool cap_small = getenv("ADTOMO_ADJ_CAP_SMALL") != nullptr;      // testing aid: staging overflow path
                const void *akern = cap_small ? (const void *)k_adj3d_topo_team<ANT, 64> : (const void *)k_adj3d_topo_team<ANT>;
                CK(cudaLaunchCooperativeKernel(akern, dim3(S * nC), dim3(ANT), args, 0, c->stream));
                phase_end(c, pk);
                LAUNCHED(c, "k_adj3d_topo_team");
                pk = phase_begin(c, PH_ADJ_FINISH);
                k_adj3d_finish2<<<elem_grid(c, d.N), 256, 0, c->stream>>>(UX, df, dGF, dGFsum, d.N, S, h);
                phase_end(c, pk);
                LAUNCHED(c, "k_adj3d_finish2");
                return 0;
            }
        }
#define ADJ_LAUNCH(NT_)                                                                                               \
    do {                                                                                                              \
        int occ = 1;                                                                                                  \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_adj3d_topo2<NT_, true>, NT_, 0));                    \
        if (occ < 1) occ = 1;                                                                                         \
        const int grid = std::min(S, c->num_sms * occ);                                                               \
        if (adj_agg) k_adj3d_topo2<NT_, true><<<grid, NT_, 0, c->stream>>>(UX, GD, CM, (unsigned int *)cnt, Q, cnts + S, cnts, d, S, d_status); \
        else k_adj3d_topo2<NT_, false><<<grid, NT_, 0, c->stream>>>(UX, GD, CM, (unsigned int *)cnt, Q, cnts + S, cnts, d, S, d_status); \
    } while (0)
        if (adj_nt <= 256) ADJ_LAUNCH(256);
        else if (adj_nt <= 512) ADJ_LAUNCH(512);
        else ADJ_LAUNCH(1024);
#undef ADJ_LAUNCH
        phase_end(c, pk);
        LAUNCHED(c, "k_adj3d_topo2");
        pk = phase_begin(c, PH_ADJ_FINISH);
        k_adj3d_finish2<<<elem_grid(c, d.N), 256, 0, c->stream>>>(UX, df, dGF, dGFsum, d.N, S, h);
        phase_end(c, pk);
        LAUNCHED(c, "k_adj3d_finish2");
    } else {
        phase_end(c, pk);
    }
    return 0;
}

static int check_dims3(int m, int n, int l, int S) {
    if (m < 2 || n < 2 || l < 2) return fail(ADTOMO_ERR_ARG, "3D grid needs m,n,l >= 2 (got %d,%d,%d)", m, n, l);
    if ((long long)m * n * l >= (1LL << 31)) return fail(ADTOMO_ERR_ARG, "m*n*l must be < 2^31");
    if (S < 1) return fail(ADTOMO_ERR_ARG, "S must be >= 1 (got %d)", S);
    return 0;
}

static int rounds_status(const std::vector<int> &r, int *rounds_out) {
    int st = 0;
    for (size_t s = 0; s < r.size(); s++) {
        if (r[s] < 0) st = ADTOMO_NOT_CONVERGED;
        if (rounds_out) rounds_out[s] = r[s] < 0 ? -r[s] : r[s];
    }
    return st;
}

extern "C" int adtomo_eikonal3d_forward_batch(adtomo_ctx *c, double *u, const double *u0, const double *f,
                                              double h, int m, int n, int l, double tol, int max_rounds, int S,
                                              int *rounds, int loc) {
    if (!c) return fail(ADTOMO_ERR_ARG, "null context");
    if (!u || !u0 || !f) return fail(ADTOMO_ERR_ARG, "null field pointer");
    int rc = check_dims3(m, n, l, S);
    if (rc) return rc;
    if (max_rounds <= 0) max_rounds = 20;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    Dims3 d{m, n, l, (long long)m * n * l};
    const double *df = nullptr;
    if ((rc = stage_in(c, "f", f, (size_t)d.N, loc, &df))) return rc;
    // chunk the sources so that host-staged batches fit the device
    int Sc = S;
    if (loc == ADTOMO_HOST) {
        size_t per = sizeof(double) * (size_t)d.N * 8;
        size_t budget = free_bytes() / 2;
        Sc = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, budget / per));
    }
    int *d_rounds;
    WS(c, "rounds", int, S, d_rounds);
    std::vector<int> hr(S);
    phase_reset(c);
    CK(cudaEventRecord(c->ev0, c->stream));
    for (int s0 = 0; s0 < S; s0 += Sc) {
        int sc = std::min(Sc, S - s0);
        double *dU;
        if (loc == ADTOMO_DEVICE) {
            dU = u + (size_t)s0 * d.N;
            if (u != u0)
                CK(cudaMemcpyAsync(dU, u0 + (size_t)s0 * d.N, sizeof(double) * d.N * sc, cudaMemcpyDeviceToDevice, c->stream));
        } else {
            WS(c, "U", double, (size_t)sc * d.N, dU);
            CK(cudaMemcpyAsync(dU, u0 + (size_t)s0 * d.N, sizeof(double) * d.N * sc, cudaMemcpyHostToDevice, c->stream));
        }
        if ((rc = fwd3d_device(c, dU, df, d, h, tol, max_rounds, sc, d_rounds + s0, nullptr))) return rc;
        if (loc == ADTOMO_HOST)
            CK(cudaMemcpyAsync(u + (size_t)s0 * d.N, dU, sizeof(double) * d.N * sc, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    CK(cudaMemcpyAsync(hr.data(), d_rounds, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return rounds_status(hr, rounds);
}

extern "C" int adtomo_eikonal3d_forward(double *u, const double *u0, const double *f, double h, int m, int n, int l,
                                        double tol, int verbose) {
    int rc;
    adtomo_ctx *c = default_ctx(&rc);
    if (rc) return rc;
    if (!u || !u0 || !f) return fail(ADTOMO_ERR_ARG, "null field pointer");
    if ((rc = check_dims3(m, n, l, 1))) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    Dims3 d{m, n, l, (long long)m * n * l};
    const int max_rounds = 20;   // Eikonal3D.cpp:74
    double *dU, *dF, *dErr;
    int *dR;
    WS(c, "U", double, d.N, dU);
    WS(c, "f", double, d.N, dF);
    WS(c, "errs", double, max_rounds, dErr);
    WS(c, "rounds", int, 1, dR);
    phase_reset(c);
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemcpyAsync(dU, u0, sizeof(double) * d.N, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dF, f, sizeof(double) * d.N, cudaMemcpyHostToDevice, c->stream));
    if ((rc = fwd3d_device(c, dU, dF, d, h, tol, max_rounds, 1, dR, dErr))) return rc;
    CK(cudaMemcpyAsync(u, dU, sizeof(double) * d.N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    int hr = 0;
    double herr[20];
    CK(cudaMemcpyAsync(&hr, dR, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(herr, dErr, sizeof(double) * max_rounds, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int nr = hr < 0 ? -hr : hr;
    if (verbose)
        for (int i = 0; i < nr; i++) printf("Iteration %d, Error = %0.6e\n", i, herr[i]);   // Eikonal3D.cpp:82-84
    return hr < 0 ? ADTOMO_NOT_CONVERGED : 0;
}

extern "C" int adtomo_eikonal3d_backward_batch(adtomo_ctx *c, double *grad_u0, double *grad_f, double *grad_f_sum,
                                               const double *grad_u, const double *u, const double *u0,
                                               const double *f, double h, int m, int n, int l, int S, int loc) {
    if (!c) return fail(ADTOMO_ERR_ARG, "null context");
    if (!grad_u || !u || !u0 || !f) return fail(ADTOMO_ERR_ARG, "null field pointer");
    int rc = check_dims3(m, n, l, S);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    Dims3 d{m, n, l, (long long)m * n * l};
    const double *df = nullptr;
    if ((rc = stage_in(c, "f", f, (size_t)d.N, loc, &df))) return rc;
    int Sc = S;
    {
        size_t per = sizeof(double) * (size_t)d.N * (loc == ADTOMO_HOST ? 9 : 4) + (size_t)d.N * 8;
        size_t budget = free_bytes() / 2;
        Sc = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, budget / per));
    }
    int *d_status;
    WS(c, "status", int, S, d_status);
    CK(cudaMemsetAsync(d_status, 0, sizeof(int) * S, c->stream));
    double *dSum = nullptr, *dSumChunk = nullptr;
    if (grad_f_sum) {
        if (loc == ADTOMO_DEVICE) dSum = grad_f_sum;
        else WS(c, "gfsum", double, d.N, dSum);
        if (Sc < S) {
            WS(c, "gfsum_chunk", double, d.N, dSumChunk);
            CK(cudaMemsetAsync(dSum, 0, sizeof(double) * d.N, c->stream));
        }
    }
    phase_reset(c);
    CK(cudaEventRecord(c->ev0, c->stream));
    for (int s0 = 0; s0 < S; s0 += Sc) {
        int sc = std::min(Sc, S - s0);
        size_t o = (size_t)s0 * d.N, cnt = (size_t)sc * d.N;
        const double *dU, *dU0, *dG;
        double *dGU0 = nullptr, *dGF = nullptr;
        if (loc == ADTOMO_DEVICE) {
            dU = u + o; dU0 = u0 + o; dG = grad_u + o;
            if (grad_u0) dGU0 = grad_u0 + o;
            if (grad_f) dGF = grad_f + o;
        } else {
            double *a, *b, *g;
            WS(c, "U", double, cnt, a);
            WS(c, "U0", double, cnt, b);
            WS(c, "G", double, cnt, g);
            CK(cudaMemcpyAsync(a, u + o, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(b, u0 + o, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(g, grad_u + o, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->stream));
            dU = a; dU0 = b; dG = g;
            if (grad_u0) WS(c, "GU0", double, cnt, dGU0);
            if (grad_f) WS(c, "GF", double, cnt, dGF);
        }
        double *sumTarget = grad_f_sum ? (Sc < S ? dSumChunk : dSum) : nullptr;
        if ((rc = adj3d_device(c, dU, dU0, dG, df, dGU0, dGF, sumTarget, d, h, sc, d_status + s0))) return rc;
        if (grad_f_sum && Sc < S) {
            k_axpy<<<elem_grid(c, d.N), 256, 0, c->stream>>>(dSum, dSumChunk, d.N);
            LAUNCHED(c, "axpy");
        }
        if (loc == ADTOMO_HOST) {
            if (grad_u0) CK(cudaMemcpyAsync(grad_u0 + o, dGU0, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->stream));
            if (grad_f) CK(cudaMemcpyAsync(grad_f + o, dGF, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    if (grad_f_sum && loc == ADTOMO_HOST)
        CK(cudaMemcpyAsync(grad_f_sum, dSum, sizeof(double) * d.N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    std::vector<int> hs(S, 1);
    if (grad_f || grad_f_sum) CK(cudaMemcpyAsync(hs.data(), d_status, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < S; s++)
        if (hs[s] < 0) return ADTOMO_ADJOINT_FLAGGED;
    return 0;
}

extern "C" int adtomo_eikonal3d_backward(double *grad_u0, double *grad_f, const double *grad_u, const double *u,
                                         const double *u0, const double *f, double h, int m, int n, int l) {
    int rc;
    adtomo_ctx *c = default_ctx(&rc);
    if (rc) return rc;
    return adtomo_eikonal3d_backward_batch(c, grad_u0, grad_f, nullptr, grad_u, u, u0, f, h, m, n, l, 1, ADTOMO_HOST);
}

// ---------------------------------------------------------------------------------------
// 2D
// ---------------------------------------------------------------------------------------
static constexpr int NT2 = 128;
static constexpr size_t SMEM2_MAX = 200 * 1024;

static int check_dims2(int m, int n, int S, const int *ix, const int *jx) {
    if (m < 1 || n < 1) return fail(ADTOMO_ERR_ARG, "2D grid needs m,n >= 1 cells (got %d,%d)", m, n);
    if ((long long)(m + 1) * (n + 1) >= (1LL << 31)) return fail(ADTOMO_ERR_ARG, "(m+1)*(n+1) must be < 2^31");
    if (S < 1) return fail(ADTOMO_ERR_ARG, "S must be >= 1");
    if (!ix || !jx) return fail(ADTOMO_ERR_ARG, "null source index pointer");
    for (int s = 0; s < S; s++)
        if (ix[s] < 0 || ix[s] > m || jx[s] < 0 || jx[s] > n)
            return fail(ADTOMO_ERR_ARG, "source %d at (%d,%d) outside the %dx%d-node grid", s, ix[s], jx[s], m + 1, n + 1);
    return 0;
}

extern "C" int adtomo_eikonal2d_forward_batch(adtomo_ctx *c, double *u, const double *f, int m, int n, double h,
                                              const int *ix, const int *jx, int S, int *rounds, int loc) {
    if (!c) return fail(ADTOMO_ERR_ARG, "null context");
    if (!u || !f) return fail(ADTOMO_ERR_ARG, "null field pointer");
    int rc = check_dims2(m, n, S, ix, jx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const long long N2 = (long long)(m + 1) * (n + 1);
    const double *df = nullptr;
    if ((rc = stage_in(c, "f", f, (size_t)N2, loc, &df))) return rc;
    int *dIX, *dJX, *dR;
    WS(c, "ix", int, S, dIX);
    WS(c, "jx", int, S, dJX);
    WS(c, "rounds", int, S, dR);
    CK(cudaMemcpyAsync(dIX, ix, sizeof(int) * S, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dJX, jx, sizeof(int) * S, cudaMemcpyHostToDevice, c->stream));
    double *dU = u;
    if (loc == ADTOMO_HOST) WS(c, "U", double, (size_t)S * N2, dU);
    const size_t smem = sizeof(double) * 2 * N2;
    const int use_smem = smem <= SMEM2_MAX;
    int grid = std::min(S, c->num_sms * (use_smem ? std::max<int>(1, (int)(SMEM2_MAX / std::max<size_t>(smem, 1))) : 2));
    grid = std::min(grid, c->num_sms * 8);
    double *gwork = nullptr;
    if (!use_smem) WS(c, "work2d", double, (size_t)grid * 2 * N2, gwork);
    if (use_smem) CK(cudaFuncSetAttribute(k_fwd2d<NT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_MAX));
    phase_reset(c);
    CK(cudaEventRecord(c->ev0, c->stream));
    k_fwd2d<NT2><<<grid, NT2, use_smem ? smem : 0, c->stream>>>(dU, df, m, n, h, dIX, dJX, S, dR, gwork, use_smem);
    LAUNCHED(c, "k_fwd2d");
    CK(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    if (loc == ADTOMO_HOST) CK(cudaMemcpyAsync(u, dU, sizeof(double) * S * N2, cudaMemcpyDeviceToHost, c->stream));
    std::vector<int> hr(S);
    CK(cudaMemcpyAsync(hr.data(), dR, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int st = rounds_status(hr, rounds);
    if (st) printf("ERROR: Eikonal does not converge!\n");   // Eikonal.h:88-90
    return st;
}

extern "C" int adtomo_eikonal2d_forward(double *u, const double *f, int m, int n, double h, int ix, int jx) {
    int rc;
    adtomo_ctx *c = default_ctx(&rc);
    if (rc) return rc;
    return adtomo_eikonal2d_forward_batch(c, u, f, m, n, h, &ix, &jx, 1, nullptr, ADTOMO_HOST);
}

extern "C" int adtomo_eikonal2d_backward_batch(adtomo_ctx *c, double *grad_f, double *grad_f_sum,
                                               const double *grad_u, const double *u, const double *f, int m, int n,
                                               double h, const int *ix, const int *jx, int S, int loc) {
    if (!c) return fail(ADTOMO_ERR_ARG, "null context");
    if (!grad_u || !u || !f) return fail(ADTOMO_ERR_ARG, "null field pointer");
    if (!grad_f && !grad_f_sum) return fail(ADTOMO_ERR_ARG, "no output requested");
    int rc = check_dims2(m, n, S, ix, jx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const long long N2 = (long long)(m + 1) * (n + 1);
    const double *df = nullptr, *dU = nullptr, *dG = nullptr;
    if ((rc = stage_in(c, "f", f, (size_t)N2, loc, &df))) return rc;
    if ((rc = stage_in(c, "U", u, (size_t)S * N2, loc, &dU))) return rc;
    if ((rc = stage_in(c, "G", grad_u, (size_t)S * N2, loc, &dG))) return rc;
    int *dIX, *dJX, *dSt;
    WS(c, "ix", int, S, dIX);
    WS(c, "jx", int, S, dJX);
    WS(c, "status", int, S, dSt);
    CK(cudaMemcpyAsync(dIX, ix, sizeof(int) * S, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dJX, jx, sizeof(int) * S, cudaMemcpyHostToDevice, c->stream));
    double *dGF;
    if (loc == ADTOMO_DEVICE && grad_f) dGF = grad_f;
    else WS(c, "GF", double, (size_t)S * N2, dGF);
    const long long per = N2 + (N2 + 7) / 8;
    const size_t smem = sizeof(double) * per;
    const int use_smem = smem <= SMEM2_MAX;
    int grid = std::min(S, c->num_sms * (use_smem ? std::max<int>(1, (int)(SMEM2_MAX / std::max<size_t>(smem, 1))) : 2));
    grid = std::min(grid, c->num_sms * 8);
    double *gwork = nullptr;
    if (!use_smem) WS(c, "work2d", double, (size_t)grid * per, gwork);
    if (use_smem) CK(cudaFuncSetAttribute(k_adj2d<NT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_MAX));
    phase_reset(c);
    CK(cudaEventRecord(c->ev0, c->stream));
    k_adj2d<NT2><<<grid, NT2, use_smem ? smem : 0, c->stream>>>(dGF, dG, dU, df, m, n, h, dIX, dJX, S, dSt, gwork, use_smem);
    LAUNCHED(c, "k_adj2d");
    double *dSum = nullptr;
    if (grad_f_sum) {
        if (loc == ADTOMO_DEVICE) dSum = grad_f_sum;
        else WS(c, "gfsum", double, N2, dSum);
        k_sum_sources<<<elem_grid(c, N2), 256, 0, c->stream>>>(dGF, dSum, N2, S);
        LAUNCHED(c, "k_sum_sources");
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    if (loc == ADTOMO_HOST) {
        if (grad_f) CK(cudaMemcpyAsync(grad_f, dGF, sizeof(double) * S * N2, cudaMemcpyDeviceToHost, c->stream));
        if (grad_f_sum) CK(cudaMemcpyAsync(grad_f_sum, dSum, sizeof(double) * N2, cudaMemcpyDeviceToHost, c->stream));
    }
    std::vector<int> hs(S);
    CK(cudaMemcpyAsync(hs.data(), dSt, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < S; s++)
        if (hs[s] < 0) return ADTOMO_ADJOINT_FLAGGED;
    return 0;
}

extern "C" int adtomo_eikonal2d_backward(double *grad_f, const double *grad_u, const double *u, const double *f,
                                         int m, int n, double h, int ix, int jx) {
    int rc;
    adtomo_ctx *c = default_ctx(&rc);
    if (rc) return rc;
    return adtomo_eikonal2d_backward_batch(c, grad_f, nullptr, grad_u, u, f, m, n, h, &ix, &jx, 1, ADTOMO_HOST);
}

// ---------------------------------------------------------------------------------------
// fused inversion step
// ---------------------------------------------------------------------------------------
// The fused step on a locked context.  f follows loc_f and grad_f (N+1 doubles) loc_out, the source / receiver
// tables follow loc: the on-device model (adtomo_model_*) keeps f and the gradient on the device while the tables
// come from wherever the caller holds them.
static int misfit_grad_core(adtomo_ctx *c, double *misfit, double *grad_f, int loc_out, const double *f, int loc_f, double h,
                            int m, int n, int l, double tol, int max_rounds, int S,
                            const int *src_ptr, const int *src_idx, const double *src_val,
                            double u0_fill, int E, const double *rcv_xyz, const double *uobs,
                            const double *qua, int *rounds, int loc) {
    if (!f || !src_ptr || !src_idx || !src_val || !rcv_xyz || !uobs || !qua)
        return fail(ADTOMO_ERR_ARG, "null input pointer");
    if (!misfit && !grad_f) return fail(ADTOMO_ERR_ARG, "no output requested");
    int rc = check_dims3(m, n, l, S);
    if (rc) return rc;
    if (E < 1) return fail(ADTOMO_ERR_ARG, "E must be >= 1");
    if (max_rounds <= 0) max_rounds = 20;
    CK(cudaSetDevice(c->device));
    Dims3 d{m, n, l, (long long)m * n * l};
    // sparse-source table: the row pointer is needed on the host to size the staging copies
    std::vector<int> hptr(S + 1);
    if (loc == ADTOMO_HOST) memcpy(hptr.data(), src_ptr, sizeof(int) * (S + 1));
    else {
        CK(cudaMemcpyAsync(hptr.data(), src_ptr, sizeof(int) * (S + 1), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    const int nnz = hptr[S];
    if (hptr[0] != 0 || nnz < 0) return fail(ADTOMO_ERR_ARG, "src_ptr must start at 0 and be non-decreasing");
    for (int s = 0; s < S; s++)
        if (hptr[s + 1] < hptr[s]) return fail(ADTOMO_ERR_ARG, "src_ptr must be non-decreasing (src_ptr[%d]=%d > src_ptr[%d]=%d)", s, hptr[s], s + 1, hptr[s + 1]);
    if (loc == ADTOMO_HOST) {   // validate what we can see
        for (int q = 0; q < nnz; q++)
            if (src_idx[q] < 0 || src_idx[q] >= d.N) return fail(ADTOMO_ERR_ARG, "src_idx[%d]=%d outside the grid", q, src_idx[q]);
        for (int e = 0; e < E; e++) {
            const double *p = rcv_xyz + 3 * e;
            if (!(p[0] >= 0 && p[0] <= m - 1 && p[1] >= 0 && p[1] <= n - 1 && p[2] >= 0 && p[2] <= l - 1))
                return fail(ADTOMO_ERR_ARG, "receiver %d at (%g,%g,%g) outside the grid", e, p[0], p[1], p[2]);
        }
    }
    const double *df = nullptr, *dval = nullptr, *drcv = nullptr, *dobs = nullptr, *dqua = nullptr;
    const int *dptr, *didx;
    if ((rc = stage_in(c, "f", f, (size_t)d.N, loc_f, &df))) return rc;
    if ((rc = stage_in(c, "src_ptr", src_ptr, (size_t)S + 1, loc, &dptr))) return rc;
    if ((rc = stage_in(c, "src_idx", src_idx, (size_t)std::max(nnz, 1), loc, &didx))) return rc;
    if ((rc = stage_in(c, "src_val", src_val, (size_t)std::max(nnz, 1), loc, &dval))) return rc;
    if ((rc = stage_in(c, "rcv", rcv_xyz, (size_t)3 * E, loc, &drcv))) return rc;
    if ((rc = stage_in(c, "uobs", uobs, (size_t)S * E, loc, &dobs))) return rc;
    if ((rc = stage_in(c, "qua", qua, (size_t)S * E, loc, &dqua))) return rc;
    if (loc == ADTOMO_DEVICE) {   // the same checks for tables the host cannot see
        int *dflag;
        WS(c, "tabflag", int, 1, dflag);
        CK(cudaMemsetAsync(dflag, 0, sizeof(int), c->stream));
        const int nq = std::max(nnz, E);
        k_validate_tables<<<(nq + 255) / 256, 256, 0, c->stream>>>(didx, nnz, d.N, drcv, E, m, n, l, dflag);
        LAUNCHED(c, "k_validate_tables");
        int hflag = 0;
        CK(cudaMemcpyAsync(&hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (hflag & 1) return fail(ADTOMO_ERR_ARG, "src_idx holds an index outside the grid (device table)");
        if (hflag & 2) return fail(ADTOMO_ERR_ARG, "a receiver lies outside the grid (device table)");
    }

    // per-source device footprint: U, U0, G, X (8 B each) + code (1 B)
    int Sc;
    {
        size_t per = (size_t)d.N * (13 * sizeof(double) + 8);
        const auto key = std::make_pair((long long)d.N, S);
        auto it = c->chunk_cache.find(key);           // cudaMemGetInfo costs milliseconds: ask once per (grid, batch)
        if (it != c->chunk_cache.end()) Sc = it->second;
        else {
            size_t budget = (size_t)(free_bytes() * 0.8);
            for (auto &kv : c->ws) budget += kv.second.second;   // what we already hold is reusable
            Sc = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, budget / per));
            // When the batch must be chunked anyway, make a chunk a whole number of waves of the batch kernels
            // (2 forward CTAs per SM, 1 adjoint CTA per SM): memory-sized chunks of 402 sources at 200x200x80 ran
            // 296 resident forward CTAs plus a second wave that was one third full (267 instead of 333 solves/s/GPU).
            const int wave = 2 * c->num_sms;
            if (Sc < S && Sc > wave) Sc = (Sc / wave) * wave;
            c->chunk_cache[key] = Sc;
        }
    }
    double *dU, *dU0, *dG = nullptr, *dMis, *dSum = nullptr, *dSumChunk = nullptr;
    int *dR, *dSt;
    WS(c, "U", double, (size_t)Sc * d.N, dU);
    WS(c, "U0", double, (size_t)Sc * d.N, dU0);
    WS(c, "mis", double, S, dMis);
    WS(c, "rounds", int, S, dR);
    WS(c, "status", int, S, dSt);
    if (grad_f) {
        WS(c, "G", double, (size_t)Sc * d.N, dG);
        if (loc_out == ADTOMO_DEVICE) dSum = grad_f;
        else WS(c, "gfsum", double, d.N + 1, dSum);
        if (Sc < S) {
            WS(c, "gfsum_chunk", double, d.N, dSumChunk);
            CK(cudaMemsetAsync(dSum, 0, sizeof(double) * d.N, c->stream));
        }
    }
    CK(cudaMemsetAsync(dMis, 0, sizeof(double) * S, c->stream));
    CK(cudaMemsetAsync(dSt, 0, sizeof(int) * S, c->stream));
    phase_reset(c);
    CK(cudaEventRecord(c->ev0, c->stream));
    for (int s0 = 0; s0 < S; s0 += Sc) {
        const int sc = std::min(Sc, S - s0);
        const long long cnt = (long long)sc * d.N;
        k_fill<<<elem_grid(c, cnt), 256, 0, c->stream>>>(dU0, cnt, u0_fill);
        LAUNCHED(c, "k_fill");
        k_scatter_sources<<<(sc + 127) / 128, 128, 0, c->stream>>>(dU0, dptr + s0, didx, dval, d.N, sc);
        LAUNCHED(c, "k_scatter_sources");
        const SparseU0 sp = {dptr + s0, didx, dval, u0_fill, dU0};
        c->batch_chunk = s0 / Sc;                 // the rounds / placement memo of the batch kernel is per chunk
        rc = fwd3d_device(c, dU, df, d, h, tol, max_rounds, sc, dR + s0, nullptr, &sp);
        c->batch_chunk = 0;
        if (rc) return rc;
        if (grad_f) CK(cudaMemsetAsync(dG, 0, sizeof(double) * cnt, c->stream));
        dim3 g((E + 127) / 128, sc);
        int pkm = phase_begin(c, PH_MISFIT);
        k_misfit<128><<<g, 128, 0, c->stream>>>(dU, dG, drcv, dobs + (size_t)s0 * E, dqua + (size_t)s0 * E, dMis + s0, d, E,
                                                 grad_f ? 1 : 0);
        phase_end(c, pkm);
        LAUNCHED(c, "k_misfit");
        if (grad_f) {
            double *target = (Sc < S) ? dSumChunk : dSum;
            if ((rc = adj3d_device(c, dU, dU0, dG, df, nullptr, nullptr, target, d, h, sc, dSt + s0, true))) return rc;
            if (Sc < S) {
                k_axpy<<<elem_grid(c, d.N), 256, 0, c->stream>>>(dSum, dSumChunk, d.N);
                LAUNCHED(c, "k_axpy");
            }
        }
    }
    // total misfit = sum over sources, fixed order, into dSum[N] when a gradient buffer exists
    double *dTot;
    WS(c, "mis_total", double, 1, dTot);
    k_sum_sources<<<1, 32, 0, c->stream>>>(dMis, dTot, 1, S);
    LAUNCHED(c, "k_sum_sources");
    if (grad_f) CK(cudaMemcpyAsync(dSum + d.N, dTot, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    if (grad_f && loc_out == ADTOMO_HOST)
        CK(cudaMemcpyAsync(grad_f, dSum, sizeof(double) * (d.N + 1), cudaMemcpyDeviceToHost, c->stream));
    double hm = 0.0;
    CK(cudaMemcpyAsync(&hm, dTot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    std::vector<int> hr(S), hs(S);
    CK(cudaMemcpyAsync(hr.data(), dR, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(hs.data(), dSt, sizeof(int) * S, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (misfit) *misfit = hm;
    if (getenv("ADTOMO_DEBUG_WAVES")) {      // debugging aid: depth of the adjoint's dependency graph
        long long tot = 0; int mx = 0;
        for (int s = 0; s < S; s++) { tot += abs(hs[s]); mx = std::max(mx, abs(hs[s])); }
        fprintf(stderr, "[adtomo] adjoint wavefront: mean waves %.1f max %d\n", (double)tot / S, mx);
    }
    int st = rounds_status(hr, rounds);
    if (grad_f)
        for (int s = 0; s < S; s++)
            if (hs[s] < 0) return ADTOMO_ADJOINT_FLAGGED;
    return st;
}

extern "C" int adtomo_eikonal3d_misfit_grad(adtomo_ctx *c, double *misfit, double *grad_f, const double *f, double h,
                                            int m, int n, int l, double tol, int max_rounds, int S,
                                            const int *src_ptr, const int *src_idx, const double *src_val,
                                            double u0_fill, int E, const double *rcv_xyz, const double *uobs,
                                            const double *qua, int *rounds, int loc) {
    if (!c) return fail(ADTOMO_ERR_ARG, "null context");
    std::lock_guard<std::mutex> lk(c->mu);
    c->oom = false;
    int rc = misfit_grad_core(c, misfit, grad_f, loc, f, loc, h, m, n, l, tol, max_rounds, S, src_ptr, src_idx, src_val, u0_fill,
                              E, rcv_xyz, uobs, qua, rounds, loc);
    if (rc == ADTOMO_ERR_CUDA && c->oom) {      // free memory shrank since the chunk size was cached: once more with fresh sizes
        c->oom = false;
        rc = misfit_grad_core(c, misfit, grad_f, loc, f, loc, h, m, n, l, tol, max_rounds, S, src_ptr, src_idx, src_val, u0_fill,
                              E, rcv_xyz, uobs, qua, rounds, loc);
    }
    return rc;
}
#include "model_api.inc"

// ---------------------------------------------------------------------------------------
// NCCL: one all-reduce of the packed [grad_f | misfit] buffer per loss/gradient evaluation
// (replaces mpi_bcast's backward + mpi_sum of scripts/inversion.jl:44,123)
// ---------------------------------------------------------------------------------------
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_ready() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    const char *why = "";
    if (!g_nccl.load(&why)) return fail(ADTOMO_ERR_NCCL, "cannot load NCCL: %s", why ? why : "?");
    return 0;
}
#define NCK(call)                                                                              \
    do {                                                                                       \
        int r__ = (call);                                                                      \
        if (r__ != 0) return fail(ADTOMO_ERR_NCCL, "%s -> %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)

extern "C" int adtomo_nccl_unique_id(char *id128) {
    if (!id128) return fail(ADTOMO_ERR_ARG, "null id buffer");
    int rc = nccl_ready();
    if (rc) return rc;
    NcclApi::UniqueId id;
    NCK(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}

extern "C" int adtomo_nccl_init(adtomo_ctx *c, const char *id128, int rank, int nranks) {
    if (!c || !id128) return fail(ADTOMO_ERR_ARG, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ADTOMO_ERR_ARG, "bad rank %d of %d", rank, nranks);
    int rc = nccl_ready();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (c->nccl_comm) return fail(ADTOMO_ERR_ARG, "context already has a communicator");
    NcclApi::UniqueId id;
    memcpy(id.internal, id128, 128);
    NCK(g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank));
    c->nccl_rank = rank;
    c->nccl_size = nranks;
    return 0;
}

extern "C" int adtomo_nccl_allreduce_sum(adtomo_ctx *c, double *buf, long long count, int loc) {
    if (!c || !buf || count < 0) return fail(ADTOMO_ERR_ARG, "bad argument");
    if (!c->nccl_comm) return fail(ADTOMO_ERR_NCCL, "adtomo_nccl_init has not been called on this context");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    double *d = buf;
    if (loc == ADTOMO_HOST) {
        WS(c, "nccl_stage", double, (size_t)count, d);
        CK(cudaMemcpyAsync(d, buf, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    }
    NCK(g_nccl.AllReduce(d, d, (size_t)count, NcclApi::kFloat64, NcclApi::kSum, c->nccl_comm, c->stream));
    if (loc == ADTOMO_HOST) CK(cudaMemcpyAsync(buf, d, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int adtomo_nccl_finalize(adtomo_ctx *c) {
    if (!c) return fail(ADTOMO_ERR_ARG, "null context");
    if (c->nccl_comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    return 0;
}

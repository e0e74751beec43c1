/*
 * adtomo_b200.h -- C ABI of libadtomo_b200.so, the B200-native replacement for the Eikonal
 * hot path of AI4EPS/ADTomo.jl (reference paths below are relative to the reference checkout).
 *
 * What it replaces: the inner free functions that the reference's TensorFlow custom-op shims
 * call (deps/CustomOps/Eikonal/Eikonal.cpp:125,217-219; Eikonal3D/EikonalThreeD.cpp:136-138,
 * 245-248).  A Julia `ccall` (or Python ctypes) binds these symbols directly; see
 * INTEGRATION.md for the stubs.
 *
 * Conventions
 *   - all field data is double (fp64); all layouts are the reference's row-major flats:
 *       2D: index j*(m+1)+i  (m, n are CELL counts, the grid has (m+1) x (n+1) nodes)
 *       3D: index (i*n+j)*l+k (m, n, l are NODE counts)
 *   - source indices are 0-based here (the 1-based -> 0-based shift stays in the language
 *     wrapper, as in Eikonal.cpp:125)
 *   - return value: 0 = ok; >0 = solver finished but flagged (1 = iteration cap reached
 *     without meeting the tolerance -- the field is still returned, like the reference;
 *     2 = adjoint stalled / singular row); <0 = argument, CUDA or NCCL error, message in
 *     adtomo_last_error().  Nothing throws or exits across this boundary.
 *   - there is NO CPU fallback: every entry point needs a CUDA device and fails with a
 *     negative code otherwise.
 *   - `loc` arguments: ADTOMO_HOST (pointers are host memory; the library stages the copies
 *     on its stream) or ADTOMO_DEVICE (pointers are device memory on the context's device).
 *     The library works on the context's OWN stream (adtomo_stream): device buffers that other
 *     streams are still writing or reading must be complete before the call (synchronise the
 *     producing stream, or make adtomo_stream wait on an event); every entry point returns
 *     after its results are complete.  Device-resident index / coordinate tables are checked
 *     on the device (out-of-range entries give ADTOMO_ERR_ARG like host tables do).
 *   - sums that involve atomics (the fused misfit and its sparse right-hand side) are
 *     reproducible to ~1e-16 relative, not bit for bit, from run to run.
 */
#ifndef ADTOMO_B200_H
#define ADTOMO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ADTOMO_HOST 0
#define ADTOMO_DEVICE 1

#define ADTOMO_OK 0
#define ADTOMO_NOT_CONVERGED 1
#define ADTOMO_ADJOINT_FLAGGED 2
#define ADTOMO_ERR_ARG (-1)
#define ADTOMO_ERR_CUDA (-2)
#define ADTOMO_ERR_NCCL (-3)

typedef struct adtomo_ctx adtomo_ctx;

/* ---- lifecycle ------------------------------------------------------------------------ */
/* One context = one device + one stream + a grow-only device workspace.  device < 0 means the
 * calling thread's current device.  Entry points are thread-safe per context. */
int adtomo_create(adtomo_ctx **ctx, int device);
int adtomo_destroy(adtomo_ctx *ctx);
/* Thread-local description of the last failure (never NULL). */
const char *adtomo_last_error(void);
int adtomo_version(void);
/* Blocks until all work queued on the context's stream has finished. */
int adtomo_synchronize(adtomo_ctx *ctx);
/* The context's cudaStream_t as an integer, so a host framework can order its own work. */
unsigned long long adtomo_stream(adtomo_ctx *ctx);
/* Device time (ms, CUDA events on the context's stream) of the kernels of the last call. */
double adtomo_last_kernel_ms(adtomo_ctx *ctx);
/* Device time (ms) of one phase of the last call, summed over its launches; CUDA events recorded
 * on the context's stream immediately around the kernels.  phase: 0 = forward sweeps kernel,
 * 1 = receiver sampling/misfit, 2 = adjoint setup, 3 = adjoint wavefront kernel, 4 = gradient finish,
 * 5 = layout conversions around the forward kernel. */
double adtomo_last_phase_ms(adtomo_ctx *ctx, int phase);
/* on != 0: from now on the event pairs of every call are kept and adtomo_last_phase_ms sums over all calls since
   (a benchmark reads the per-kernel device times of its whole timed region AFTER the region, without synchronising
   inside it); on == 0: back to "last call only".  Either way the accumulated pairs are dropped. */
int adtomo_phase_accumulate(adtomo_ctx *ctx, int on);
/* Number of kernels this library has launched on the context since creation. */
long long adtomo_launch_count(adtomo_ctx *ctx);
/* Names the batch of sources the following batched 3D calls work on (default 0).  The batch kernel remembers, per
   (grid, number of sources, batch id), how many rounds every source needed in the previous call and places long and
   short sources together on an SM; a driver that alternates between several source sets (e.g. the P and the S stations)
   gives each its own id.  Purely a performance hint: results do not depend on it. */
int adtomo_set_batch_id(adtomo_ctx *ctx, long long id);
/* Name of the 3D forward sweep kernel the last 3D call ran on ("k_fwd3d_v3", "k_fwd3d_team", "k_fwd3d_v1", ...):
   the tests use it to prove which kernel family they compared with the oracle. */
const char *adtomo_last_forward_kernel(adtomo_ctx *ctx);
/* Self-test: the library's call-free fp64 square root (csrc/eik_core.h) against the CUDA math library's
   correctly rounded sqrt on n pseudo-random and special arguments; *mismatches must come back 0. */
int adtomo_selftest_sqrt(adtomo_ctx *ctx, long long n, unsigned long long seed, long long *mismatches);

/* ---- 1:1 replacements, host pointers, use an internal per-thread default context -------- */
/* forward(u,f,m,n,h,ix,jx)           deps/CustomOps/Eikonal/Eikonal.h:54-93 */
int adtomo_eikonal2d_forward(double *u, const double *f, int m, int n, double h, int ix, int jx);
/* backward(grad_f,grad_u,u,f,m,n,h,ix,jx)   deps/CustomOps/Eikonal/Eikonal.h:95-200 */
int adtomo_eikonal2d_backward(double *grad_f, const double *grad_u, const double *u, const double *f,
                              int m, int n, double h, int ix, int jx);
/* Eikonal3D::forward(u,u0,f,h,m,n,l,tol,verbose)   deps/CustomOps/Eikonal3D/Eikonal3D.cpp:90-94
 * (20-round cap of :74 applied; verbose prints the reference's per-round line of :82-84). */
int adtomo_eikonal3d_forward(double *u, const double *u0, const double *f, double h, int m, int n, int l,
                             double tol, int verbose);
/* Eikonal3D::backward(grad_u0,grad_f,grad_u,u,u0,f,h,m,n,l)   Eikonal3D.cpp:96-198 */
int adtomo_eikonal3d_backward(double *grad_u0, double *grad_f, const double *grad_u, const double *u,
                              const double *u0, const double *f, double h, int m, int n, int l);

/* ---- batched entry points: S sources share one slowness field -------------------------- */
/* u, u0: S*N; f: N; rounds (may be NULL): S ints, the 8-sweep rounds each source ran;
 * max_rounds <= 0 means the reference's cap of 20. */
int adtomo_eikonal3d_forward_batch(adtomo_ctx *ctx, double *u, const double *u0, const double *f, double h,
                                   int m, int n, int l, double tol, int max_rounds, int S, int *rounds,
                                   int loc);
/* grad_u, u, u0: S*N; outputs (each may be NULL): grad_u0 S*N, grad_f S*N (per source),
 * grad_f_sum N (sum over sources, what the drivers' AddN + mpi reduce produce). */
int adtomo_eikonal3d_backward_batch(adtomo_ctx *ctx, double *grad_u0, double *grad_f, double *grad_f_sum,
                                    const double *grad_u, const double *u, const double *u0, const double *f,
                                    double h, int m, int n, int l, int S, int loc);
/* 2D: u, grad_u: S*N2; f: N2; ix, jx: S ints (0-based, always HOST memory). */
int adtomo_eikonal2d_forward_batch(adtomo_ctx *ctx, double *u, const double *f, int m, int n, double h,
                                   const int *ix, const int *jx, int S, int *rounds, int loc);
int adtomo_eikonal2d_backward_batch(adtomo_ctx *ctx, double *grad_f, double *grad_f_sum, const double *grad_u,
                                    const double *u, const double *f, int m, int n, double h, const int *ix,
                                    const int *jx, int S, int loc);

/* ---- fused inversion step (the per-source work of scripts/inversion.jl:46-105) --------- */
/* grad_f (when not NULL) must have room for N+1 doubles: elements [0,N) receive the gradient
 * and element N the misfit, so that ONE all-reduce of N+1 doubles covers both.
 * In a multi-GPU run each rank calls this on its shard of sources and sums the buffers. */
/* For S sources (stations) given as sparse initial conditions
 *     u0_s = u0_fill everywhere, u0_s[src_idx[q]] = src_val[q] for q in [src_ptr[s], src_ptr[s+1])
 * solve forward with tolerance tol, sample every field at E receivers (events) by the drivers'
 * trilinear rule (rcv_xyz: E*3 fractional 0-based node coordinates; an integer coordinate uses
 * that node alone), form misfit = sum_{s,e} qua[s*E+e]*(uobs[s*E+e]-t_se)^2 skipping
 * uobs == -1, and back-propagate to grad_f (N, summed over sources; NULL -> misfit only).
 * The travel-time fields never leave the device.  rounds (may be NULL): S ints.
 * src_ptr/src_idx/src_val/rcv_xyz/uobs/qua follow `loc`; misfit is always a HOST double. */
int adtomo_eikonal3d_misfit_grad(adtomo_ctx *ctx, double *misfit, double *grad_f, const double *f, double h,
                                 int m, int n, int l, double tol, int max_rounds, int S, const int *src_ptr,
                                 const int *src_idx, const double *src_val, double u0_fill, int E,
                                 const double *rcv_xyz, const double *uobs, const double *qua, int *rounds,
                                 int loc);

/* ---- model parametrisation, chain rule and regulariser on the device --------------------- */
/* One loss/gradient evaluation of the inversion drivers (scripts/inversion.jl:42-43,61,96-121; joint P+S:
 * inversion_joint.jl:49-51,80,140-166) = adtomo_model_begin, one adtomo_model_add_phase per seismic phase,
 * adtomo_model_finish.  N = m*n*l optimiser variables go in, N+1 doubles come out; the slowness fields, their
 * gradients and the travel times stay on the device.
 *   begin      x (var_change) and vel0, N doubles each (follow loc): fvar = 2*sigmoid(x) - 1 + vel0.
 *   add_phase  slowness of the phase f = scale / fvar (P: scale = 1, inversion.jl:61; S: scale = pvs,
 *              inversion_joint.jl:80) and the fused step of adtomo_eikonal3d_misfit_grad on this phase's sources
 *              (same source / receiver arguments, following loc).  *misfit and *grad_scale (HOST doubles, may be
 *              NULL) receive the phase's misfit and d misfit / d scale (the gradient of the joint driver's `pvs`).
 *              want_grad == 0: misfit only.  Returns the fused step's status.
 *   finish     loss = sum of the phases' misfits + (add_reg ? lambda * sum |fvar - box(fvar)| : 0), box = mean over
 *              the periodic smooth_hor x smooth_hor x smooth_ver window (odd sizes; inversion.jl:107-121);
 *              *loss (HOST, may be NULL); packed (N+1 doubles, follows loc; needed when want_grad != 0):
 *              [d loss / d x | loss].  Returns the worst status of the phases.
 * Multi-GPU: every rank evaluates its source shard, ONE rank passes add_reg != 0 (the reference adds the
 * regulariser on every rank before mpi_sum, i.e. nproc times -- SURVEY 5), then packed is all-reduced.
 * The three calls of one evaluation must not be interleaved with other evaluations on the same context. */
int adtomo_model_begin(adtomo_ctx *ctx, const double *x, const double *vel0, int m, int n, int l, int loc);
int adtomo_model_add_phase(adtomo_ctx *ctx, double scale, double h, double tol, int max_rounds, int S,
                           const int *src_ptr, const int *src_idx, const double *src_val, double u0_fill, int E,
                           const double *rcv_xyz, const double *uobs, const double *qua, int *rounds, double *misfit,
                           double *grad_scale, int want_grad, int loc);
int adtomo_model_finish(adtomo_ctx *ctx, double lambda, int smooth_hor, int smooth_ver, int add_reg, int want_grad,
                        double *loss, double *packed, int loc);
/* Single-phase evaluation (scripts/inversion.jl) in one call: begin + add_phase(scale 1) + finish.
 * packed == NULL: loss only. */
int adtomo_model_loss_grad(adtomo_ctx *ctx, double *loss, double *packed, const double *x, const double *vel0,
                           double lambda, int smooth_hor, int smooth_ver, int add_reg, double h, int m, int n, int l,
                           double tol, int max_rounds, int S, const int *src_ptr, const int *src_idx,
                           const double *src_val, double u0_fill, int E, const double *rcv_xyz, const double *uobs,
                           const double *qua, int *rounds, int loc);

/* ---- multi-GPU: source shards + ONE all-reduce per evaluation ---------------------------- */
/* Replaces the MPI plumbing of the drivers (mpi_bcast / mpi_sum, scripts/inversion.jl:44,123;
 * flag protocol of src/mpi_optimize.jl:11-33): one process per GPU, rank r evaluates sources
 * r, r+P, ... with adtomo_eikonal3d_misfit_grad and sums the packed N+1 buffer over ranks.
 * NCCL is loaded with dlopen("libnccl.so.2") on first use (the library itself does not link it).
 * Rank 0 creates the 128-byte id and distributes it by any means (file, socket, MPI, torchrun). */
int adtomo_nccl_unique_id(char *id128);
int adtomo_nccl_init(adtomo_ctx *ctx, const char *id128, int rank, int nranks);
/* In-place sum over ranks of `count` doubles (ncclAllReduce on the context's stream; loc = HOST
 * stages through the workspace).  Returns after the result is complete. */
int adtomo_nccl_allreduce_sum(adtomo_ctx *ctx, double *buf, long long count, int loc);
int adtomo_nccl_finalize(adtomo_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* ADTOMO_B200_H */

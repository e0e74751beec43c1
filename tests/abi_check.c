/* tests/abi_check.c -- TEST INFRASTRUCTURE.  A plain C caller of libadtomo_b200.so that sees ONLY include/adtomo_b200.h:
 * what a Julia `ccall`, a cgo or a JNI stub binds.  It proves that the header (not just the Python ctypes argtypes)
 * matches the binary: forward + backward of a 9 x 7 x 6 problem through the 1:1 entry points that replace
 * Eikonal3D::forward / ::backward (Eikonal3D.cpp:90-94, :96-198), the same through the context-based batch entry
 * points, and the fused step; results go to a file that tests/test_abi_c.py compares with the oracle.
 *   abi_check              takes the address of every declared function (link check; no GPU needed)
 *   abi_check <out.bin>    runs on the GPU and writes  u | grad_u0 | grad_f | u_batch | packed(N+1)  as doubles
 * Build: gcc -std=c99 -Iinclude tests/abi_check.c -o abi_check -Ladtomo.jl_b200 -ladtomo_b200 -lm   (the test does it) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "adtomo_b200.h"

#define M 9
#define N_ 7
#define L 6
#define NN (M * N_ * L)

static double frand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (double)(*s >> 8) / 16777216.0; }

int main(int argc, char **argv) {
    /* every symbol the header declares, with the header's own type */
    const void *syms[] = {(void *)adtomo_create, (void *)adtomo_destroy, (void *)adtomo_synchronize, (void *)adtomo_last_error,
                          (void *)adtomo_eikonal2d_forward, (void *)adtomo_eikonal2d_backward, (void *)adtomo_eikonal3d_forward,
                          (void *)adtomo_eikonal3d_backward, (void *)adtomo_eikonal3d_forward_batch,
                          (void *)adtomo_eikonal3d_backward_batch, (void *)adtomo_eikonal2d_forward_batch,
                          (void *)adtomo_eikonal2d_backward_batch, (void *)adtomo_eikonal3d_misfit_grad, (void *)adtomo_model_begin,
                          (void *)adtomo_model_add_phase, (void *)adtomo_model_finish, (void *)adtomo_model_loss_grad,
                          (void *)adtomo_nccl_unique_id, (void *)adtomo_nccl_init, (void *)adtomo_nccl_allreduce_sum,
                          (void *)adtomo_nccl_finalize, (void *)adtomo_selftest_sqrt, (void *)adtomo_set_batch_id,
                          (void *)adtomo_launch_count, (void *)adtomo_last_phase_ms, (void *)adtomo_phase_accumulate};
    for (size_t i = 0; i < sizeof syms / sizeof *syms; i++)
        if (!syms[i]) return 2;
    if (argc < 2) { printf("abi_check: %zu symbols linked\n", sizeof syms / sizeof *syms); return 0; }

    static double f[NN], u0[NN], u[NN], gu[NN], gu0[NN], gf[NN], ub[NN], packed[NN + 1];
    unsigned seed = 12345u;
    for (int q = 0; q < NN; q++) { f[q] = 0.5 + frand(&seed); u0[q] = 1000.0; }
    for (int q = 0; q < NN; q++) gu[q] = frand(&seed) - 0.5;
    const int src = (4 * N_ + 3) * L + 2;
    u0[src] = 0.0;
    const double h = 0.3, tol = 1e-9;
    int rc = adtomo_eikonal3d_forward(u, u0, f, h, M, N_, L, tol, 0);
    if (rc != 0) { fprintf(stderr, "forward rc %d: %s\n", rc, adtomo_last_error()); return 3; }
    rc = adtomo_eikonal3d_backward(gu0, gf, gu, u, u0, f, h, M, N_, L);
    if (rc != 0) { fprintf(stderr, "backward rc %d: %s\n", rc, adtomo_last_error()); return 4; }

    adtomo_ctx *ctx = NULL;
    rc = adtomo_create(&ctx, 0);
    if (rc != 0 || !ctx) { fprintf(stderr, "create rc %d: %s\n", rc, adtomo_last_error()); return 5; }
    int rounds = 0;
    rc = adtomo_eikonal3d_forward_batch(ctx, ub, u0, f, h, M, N_, L, tol, 0, 1, &rounds, ADTOMO_HOST);
    if (rc != 0 || rounds <= 0) { fprintf(stderr, "forward_batch rc %d rounds %d: %s\n", rc, rounds, adtomo_last_error()); return 6; }
    /* fused step: one source given sparsely (the same point source), three receivers */
    const int ptr[2] = {0, 1}, idx[1] = {src};
    const double val[1] = {0.0};
    const double rcv[9] = {1.0, 1.0, 1.0, 7.25, 5.5, 4.75, 2.0, 6.0, 0.5};
    const double uobs[3] = {1.0, -1.0, 2.5}, qua[3] = {1.0, 0.7, 0.4};
    double misfit = 0.0;
    rc = adtomo_eikonal3d_misfit_grad(ctx, &misfit, packed, f, h, M, N_, L, tol, 0, 1, ptr, idx, val, 1000.0, 3, rcv, uobs, qua,
                                      NULL, ADTOMO_HOST);
    if (rc != 0 || misfit != packed[NN]) { fprintf(stderr, "misfit_grad rc %d: %s\n", rc, adtomo_last_error()); return 7; }
    if (adtomo_launch_count(ctx) <= 0) return 8;
    adtomo_destroy(ctx);

    FILE *fp = fopen(argv[1], "wb");
    if (!fp) return 9;
    fwrite(u, sizeof(double), NN, fp);
    fwrite(gu0, sizeof(double), NN, fp);
    fwrite(gf, sizeof(double), NN, fp);
    fwrite(ub, sizeof(double), NN, fp);
    fwrite(packed, sizeof(double), NN + 1, fp);
    fclose(fp);
    printf("abi_check: ok, %d rounds, misfit %.17g\n", rounds, misfit);
    return 0;
}

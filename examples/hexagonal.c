/* hexagonal.c -- a C consumer of the drop-in boundary (include/rome_b200.h): the graph of RoME's
 * generateGraph_Hexagonal (src/canonical/GenerateHexagonal.jl:27-42, GenerateCircular.jl:57-90) -- seven Pose2 on a
 * hexagon, a prior on x0, six Pose2Pose2 legs [10, 0, pi/3] -- evaluated through the batched entry points that replace
 * IIF's per-particle CalcFactor loop: one call per (factor family, sweep).
 *
 *   gcc -std=c99 -I include examples/hexagonal.c -L rome.jl_b200 -lrome_b200 -Wl,-rpath,$PWD/rome.jl_b200 -lm -o hexagonal
 *
 * Needs a B200 at run time (rome_b200_create fails with ROME_B200_NO_DEVICE otherwise: there is no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rome_b200.h"

#define N 100
#define NPOSES 7
#define NLEGS 6

static int check(rome_b200_ctx* ctx, int rc, const char* what) {
    if (rc != ROME_B200_OK) fprintf(stderr, "%s: error %d: %s\n", what, rc, rome_b200_last_error(ctx));
    return rc;
}

int main(void) {
    const double kPi = 3.14159265358979323846;
    rome_b200_ctx* ctx = NULL;
    int rc = rome_b200_create(0, &ctx);
    if (rc != ROME_B200_OK) {
        fprintf(stderr, "rome_b200_create: error %d: %s\n", rc, rome_b200_last_error(NULL));
        return rc == ROME_B200_NO_DEVICE ? 77 : 1; /* 77: skipped, no device */
    }
    const int npad = rome_b200_npad(N);

    /* particles: the hexagon's ground truth plus a deterministic spread, DFG `vecval` layout [var][N][3] */
    double* coords = (double*)malloc(sizeof(double) * NPOSES * N * 3);
    double x = 0.0, y = 0.0, th = 0.0;
    for (int v = 0; v < NPOSES; ++v) {
        for (int n = 0; n < N; ++n) {
            const double u = (n + 0.5) / N - 0.5;
            coords[(v * N + n) * 3 + 0] = x + 0.2 * u;
            coords[(v * N + n) * 3 + 1] = y - 0.2 * u;
            coords[(v * N + n) * 3 + 2] = th + 0.04 * u;
        }
        x += 10.0 * cos(th);
        y += 10.0 * sin(th);
        th += kPi / 3.0;
    }
    if (check(ctx, rome_b200_set_particles(ctx, ROME_B200_POSE2, NPOSES, N, coords), "set_particles")) return 1;

    /* factors: Pose2Pose2(MvNormal([10, 0, pi/3], diagm([0.1, 0.1, 0.1].^2))) x 6, PriorPose2(MvNormal(0, 0.01 I)) */
    int32_t ip[NLEGS], iq[NLEGS], prior_var[1] = {0};
    double mu[NLEGS * 3], cov[NLEGS * 9], prior_mu[3] = {0, 0, 0}, prior_cov[9] = {0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.01};
    for (int f = 0; f < NLEGS; ++f) {
        ip[f] = f;
        iq[f] = f + 1;
        mu[3 * f + 0] = 10.0; mu[3 * f + 1] = 0.0; mu[3 * f + 2] = kPi / 3.0;
        for (int k = 0; k < 9; ++k) cov[9 * f + k] = (k % 4 == 0) ? 0.01 : 0.0;
    }
    if (check(ctx, rome_b200_set_factors_pose2pose2(ctx, NLEGS, ip, iq, mu, cov), "set_factors_pose2pose2")) return 1;
    if (check(ctx, rome_b200_set_factors_priorpose2(ctx, 1, prior_var, prior_mu, prior_cov), "set_factors_priorpose2")) return 1;

    /* one sweep of the hot path with host buffers: getSample (in-kernel) + residual + forward proposal + statistics
     * for every factor x particle of the family */
    float* res = (float*)calloc((size_t)NLEGS * npad * 3, sizeof(float));
    float* prop = (float*)calloc((size_t)NLEGS * npad * 3, sizeof(float));
    float* stats = (float*)calloc((size_t)NLEGS * 16, sizeof(float));
    rome_b200_buffers b = {NULL, NULL, res, prop, NULL, stats, NULL};
    const uint32_t flags = ROME_B200_SAMPLE | ROME_B200_RESIDUAL | ROME_B200_PROPOSAL_FWD | ROME_B200_STATS;
    if (check(ctx, rome_b200_eval_host(ctx, ROME_B200_POSE2POSE2, flags, 1u /*seed*/, 0u /*sweep*/, 0, -1, &b), "eval_host")) return 1;

    for (int f = 0; f < NLEGS; ++f) {
        /* stats[0..2] = sum of the residual over the N particles; the samples have sigma 0.1 around a consistent graph */
        printf("x%dx%df1  mean residual = (%+.4f, %+.4f, %+.4f)\n", f, f + 1, stats[16 * f] / N, stats[16 * f + 1] / N,
               stats[16 * f + 2] / N);
    }
    printf("kernel launches: %llu\n", (unsigned long long)rome_b200_launch_count(ctx));
    free(coords); free(res); free(prop); free(stats);
    return rome_b200_destroy(ctx);
}

/* TEST / BENCH INFRASTRUCTURE ONLY - CPU baseline driver, not part of the product.
 *
 * Drives the free-mode oLBFGS request loop of whatever libstochqn it is linked to
 * (the unmodified reference core built into oracle/_ref/ by oracle/build_ref.py) on
 * the chained Rosenbrock function, exactly the workload bench.py runs on the GPU
 * (BASELINE.json config 4):
 *
 *   x0[i] = 0.95 + 1e-4 * ((uint32)(i * 2654435761) mod 1000),  constant step,
 *   gradient formulas of the reference's example (example/c_rosen.c:26-41),
 *   served for both calc_grad and calc_grad_same_batch.
 *
 * The loop has the shape of example/c_rosen.c:99-118 (call, serve the request, call
 * again) with run_oLBFGS (include/stochqn.h:381) instead of run_SQN.  Gradients are
 * evaluated with OpenMP so the reference gets every host core it can use.
 *
 * usage: rosen_harness n mem_size warmup steps nthreads min_curvature step_size [check_nan]
 * prints one JSON line.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "stochqn.h"

#include <immintrin.h>
#include <string.h>
/* write-only streams (the gradient array): non-temporal store, no read-for-ownership of the destination line */
static inline void stream_store(real_t *p, real_t v)
{
#ifdef USE_FLOAT
    int b; memcpy(&b, &v, sizeof b); _mm_stream_si32((int*) p, b);
#else
    long long b; memcpy(&b, &v, sizeof b); _mm_stream_si64((long long*) p, b);
#endif
}

/* 0.95 + 1e-4*(h mod 1000) with the product rounded before the sum (no FMA contraction), so that the
   start point is bit-identical in C, NumPy and CUDA */
static double x0_value(unsigned int h)
{
    volatile double prod = 1e-4 * (double) (h % 1000u);
    return 0.95 + prod;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void rosen_grad(const real_t *x, long n, real_t *g)
{
    g[0] = (real_t)(-400.0 * x[0] * (x[1] - x[0] * x[0]) - 2.0 * (1.0 - x[0]));
    g[n - 1] = (real_t)(200.0 * (x[n - 1] - x[n - 2] * x[n - 2]));
    #pragma omp parallel for schedule(static)
    for (long i = 1; i < n - 1; i++) {
        double a = 200.0 * (x[i] - x[i - 1] * x[i - 1]);
        double b = 400.0 * (x[i + 1] - x[i] * x[i]) * x[i];
        double c = 2.0 * (1.0 - x[i]);
        stream_store(&g[i], (real_t)(a - b - c));
    }
}

int main(int argc, char **argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s n mem_size warmup steps nthreads min_curvature step_size [check_nan]\n", argv[0]);
        return 2;
    }
    long   n        = atol(argv[1]);
    size_t mem_size = (size_t) atol(argv[2]);
    long   warmup   = atol(argv[3]);
    long   steps    = atol(argv[4]);
    int    nthreads = atoi(argv[5]);
    double min_curv = atof(argv[6]);
    double step     = atof(argv[7]);
    int    check_nan = (argc > 8) ? atoi(argv[8]) : 1;

#ifdef _OPENMP
    omp_set_num_threads(nthreads);
#endif
    real_t *x = (real_t*) malloc(sizeof(real_t) * (size_t) n);
    real_t *g = (real_t*) malloc(sizeof(real_t) * (size_t) n);
    if (!x || !g) { fprintf(stderr, "harness: out of memory\n"); return 1; }
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        uint32_t h = (uint32_t)((uint64_t) i * 2654435761ull);
        x[i] = (real_t) x0_value(h);
        g[i] = 0;
    }

    workspace_oLBFGS *ws = initialize_oLBFGS((int) n, mem_size, 0, 0, (real_t) min_curv, check_nan, nthreads);
    if (!ws) { fprintf(stderr, "harness: initialize_oLBFGS failed\n"); return 1; }

    real_t *req = NULL;
    task_enum task;
    info_enum info;
    long n_info = 0;
    double t_opt = 0, t0 = 0, t_start = 0;
    int timing = 0;
    run_oLBFGS((real_t) step, x, g, &req, &task, ws, &info);    /* section 0: first request */
    /* one iteration = serve calc_grad, call (step), serve calc_grad_same_batch, call (pair).
       The clock starts once `warmup` whole iterations are done and stops after `steps` more. */
    for (;;) {
        if (!timing && (long) ws->niter == warmup && task == calc_grad) { timing = 1; t_start = now_s(); }
        if ((long) ws->niter >= warmup + steps && task == calc_grad) break;
        rosen_grad(req, n, g);
        t0 = now_s();
        run_oLBFGS((real_t) step, x, g, &req, &task, ws, &info);
        if (timing) t_opt += now_s() - t0;
        if (info != no_problems_encountered) n_info++;
    }
    double t_total = now_s() - t_start;

    /* what bench.py compares between the two arms: norm, plain sum and a handful of entries of the final iterate */
    double nrm = 0, sum = 0;
    #pragma omp parallel for schedule(static) reduction(+:nrm,sum)
    for (long i = 0; i < n; i++) { nrm += (double) x[i] * (double) x[i]; sum += (double) x[i]; }
    const long pidx[7] = {0, n > 1 ? 1 : 0, n / 4, n / 2, (3 * n) / 4, n > 1 ? n - 2 : 0, n - 1};
    printf("{\"n\": %ld, \"mem_size\": %zu, \"warmup\": %ld, \"steps\": %ld, \"nthreads\": %d, "
           "\"seconds\": %.6f, \"opt_seconds\": %.6f, \"steps_per_s\": %.6f, \"opt_steps_per_s\": %.6f, "
           "\"info_events\": %ld, \"mem_used\": %zu, \"mem_st_ix\": %zu, \"niter\": %zu, "
           "\"x_norm\": %.17g, \"x_sum\": %.17g, \"x0\": %.17g, \"x_mid\": %.17g, \"x_last\": %.17g, "
           "\"probe_idx\": [%ld, %ld, %ld, %ld, %ld, %ld, %ld], "
           "\"probes\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"real_bytes\": %d}\n",
           n, mem_size, warmup, steps, nthreads, t_total, t_opt,
           (double) steps / t_total, (double) steps / t_opt, n_info,
           ws->bfgs_memory->mem_used, ws->bfgs_memory->mem_st_ix, ws->niter,
           sqrt(nrm), sum, (double) x[0], (double) x[n / 2], (double) x[n - 1],
           pidx[0], pidx[1], pidx[2], pidx[3], pidx[4], pidx[5], pidx[6],
           (double) x[pidx[0]], (double) x[pidx[1]], (double) x[pidx[2]], (double) x[pidx[3]], (double) x[pidx[4]],
           (double) x[pidx[5]], (double) x[pidx[6]], (int) sizeof(real_t));
    dealloc_oLBFGS(ws);
    free(x); free(g);
    return 0;
}

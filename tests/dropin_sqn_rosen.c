/* Drop-in check: a plain C program that uses the stochQN C ABI exactly the way the reference's example does
 * (example/c_rosen.c: malloc'ed HOST arrays, initialize_SQN -> run_SQN request loop -> dealloc_SQN), written
 * here from the API documentation.  Linked against the CUDA library it must print the reference's known answers
 * (SURVEY.md section 4).  The Hessian-vector routine reproduces the example's first row literally (it multiplies
 * p[0] where the analytic Hessian has p[1], c_rosen.c:46) because the known answers were produced with it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "stochqn.h"

static double fun(const double *x, int n)
{
    double out = 0;
    for (int i = 0; i < n - 1; i++) {
        double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
        out += 100.0 * a * a + b * b;
    }
    return out;
}

static void grad(const double *x, int n, double *g)
{
    g[0] = -400.0 * x[0] * (x[1] - x[0] * x[0]) - 2.0 * (1.0 - x[0]);
    g[n - 1] = 200.0 * (x[n - 1] - x[n - 2] * x[n - 2]);
    for (int i = 1; i < n - 1; i++)
        g[i] = 200.0 * (x[i] - x[i - 1] * x[i - 1]) - 400.0 * (x[i + 1] - x[i] * x[i]) * x[i] - 2.0 * (1.0 - x[i]);
}

static void hess_vec(const double *x, const double *p, int n, double *out)
{
    memset(out, 0, sizeof(double) * n);
    out[0] = (1200 * x[0] * x[0] - 400 * x[1] + 2.0) * p[0] - 400 * x[0] * p[0];
    out[n - 1] = -400.0 * x[n - 2] * p[n - 2] + 200.0 * p[n - 1];
    for (int i = 1; i < n - 1; i++)
        out[i] = -400.0 * x[i - 1] * p[i - 1] + (202 + 1200 * x[i] * x[i] - 400 * x[i + 1]) * p[i] - 400.0 * x[i] * p[i + 1];
}

int main(void)
{
    int n = 4;
    double x[] = {1.3, 0.7, 0.8, 1.9};
    double *g = (double*) malloc(sizeof(double) * n), *hv = (double*) malloc(sizeof(double) * n);
    double *req, *req_vec;
    task_enum task;
    info_enum info;
    printf("f0 %.4f\n", fun(x, n));
    workspace_SQN *ws = initialize_SQN(n, 5, 3, 0, 0, 1e-8, 1, 1);
    if (!ws) return 1;
    run_SQN(1e-3, x, g, hv, &req, &req_vec, &task, ws, &info);
    while (ws->niter < 200) {
        if (task == calc_grad) grad(req, n, g);
        else if (task == calc_hess_vec) hess_vec(req, req_vec, n, hv);
        int upd = run_SQN(1e-3, x, g, hv, &req, &req_vec, &task, ws, &info);
        if (upd && ((ws->niter + 1) % 10) == 0) printf("it %zu %.4f\n", ws->niter + 1, fun(x, n));
    }
    printf("final %.4f\n", fun(x, n));
    printf("x %f %f %f %f\n", x[0], x[1], x[2], x[3]);
    dealloc_SQN(ws);
    free(g); free(hv);
    return 0;
}

/* Host-side (CPU, OpenMP) gradient callback used by bench.py's end-to-end leg: the caller of the
 * drop-in C ABI keeps x and grad in HOST memory and evaluates the chained Rosenbrock gradient itself,
 * exactly what a user of the reference does (example/c_rosen.c:26-41, formulas only).
 * Shard-aware: x[0..n_local) starts at global index `offset` of a vector of length n_global;
 * halo_left / halo_right are x[offset-1] and x[offset+n_local] (ignored at the global ends). */
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#ifdef USE_FLOAT
typedef float real_t;
#else
typedef double real_t;
#endif

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

static inline double rosen_entry(double xm, double xc, double xp, int has_left, int has_right)
{
    double out = 0.0;
    if (has_left) out += 200.0 * (xc - xm * xm);
    if (has_right) { out -= 400.0 * (xp - xc * xc) * xc; out -= 2.0 * (1.0 - xc); }
    return out;
}

void host_rosenbrock_grad(const real_t *x, real_t *g, long long n_local, long long offset, long long n_global,
                          double halo_left, double halo_right)
{
    if (n_local <= 0) return;
    /* the two ends of the shard (they may be the ends of the whole vector, or need the neighbours' halo values) */
    {
        const double xp = n_local > 1 ? (double) x[1] : halo_right;
        g[0] = (real_t) rosen_entry(halo_left, (double) x[0], xp, offset > 0, offset < n_global - 1);
    }
    if (n_local > 1) {
        const long long i = n_local - 1;
        g[i] = (real_t) rosen_entry((double) x[i - 1], (double) x[i], halo_right, 1, offset + i < n_global - 1);
    }
    /* interior: branch-free, same operation order as the boundary form (example/c_rosen.c:32-37) */
    #pragma omp parallel for schedule(static)
    for (long long i = 1; i < n_local - 1; i++) {
        const double xm = (double) x[i - 1], xc = (double) x[i], xp = (double) x[i + 1];
        double out = 200.0 * (xc - xm * xm);
        out -= 400.0 * (xp - xc * xc) * xc;
        out -= 2.0 * (1.0 - xc);
        stream_store(&g[i], (real_t) out);
    }
}

void host_rosenbrock_x0(real_t *x, long long n_local, long long offset)
{
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < n_local; i++) {
        unsigned int h = (unsigned int) ((unsigned long long) (i + offset) * 2654435761ull);
        x[i] = (real_t) x0_value(h);
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the bench gives each rank its share of the host cores */
void host_set_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void) nthreads;
#endif
}

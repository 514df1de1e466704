/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the selective scan.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path never does.
 *
 * What it restates: the recurrence the reference hands to the third-party package
 * `mamba_ssm` (unpinned in /root/reference/requirements.txt:16, source NOT vendored) at
 * basicsr/archs/wavemamba_arch.py:465-471:
 *
 *     selective_scan_fn(u, delta, A, B, C, D, z=None, delta_bias, delta_softplus=True)
 *
 * following the published `selective_scan_ref` semantics of that package:
 *     delta = softplus(delta + delta_bias[ch])          (softplus threshold 20: x > 20 -> x)
 *     h_0   = 0
 *     h_l[n] = exp(delta_l * A[ch,n]) * h_{l-1}[n] + delta_l * B[b,g,n,l] * u_l
 *     y_l    = sum_n C[b,g,n,l] * h_l[n] + D[ch] * u_l
 * with B/C grouped: group g covers channels g*(DIM/G) .. (g+1)*(DIM/G)-1.
 *
 * PARITY UNPINNED against mamba_ssm's CUDA kernel (package absent, no network); pinned
 * only against the recurrence above (tests/test_oracle.py checks this C code against a
 * pure-Python loop and against the float64 evaluation).
 *
 * Layouts (all contiguous, row-major):
 *   u, delta, out : (batch, dim, L)      A : (dim, N)      B, C : (batch, G, N, L)
 *   D, delta_bias : (dim)
 * Streaming form: nothing of size (dim, L, N) is ever materialised.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_STATE 64

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline sets its thread count explicitly. */
int wm_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static inline float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
static inline double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }

/* float32 arithmetic throughout (what the reference computes in). */
int wm_oracle_selective_scan_f32(const float *u, const float *delta, const float *A,
                                 const float *Bm, const float *Cm, const float *D,
                                 const float *delta_bias, float *out, int64_t batch,
                                 int64_t dim, int64_t groups, int64_t nstate, int64_t L)
{
    if (nstate > MAX_STATE || groups <= 0 || dim % groups != 0) return -1;
    const int64_t per_group = dim / groups;
    const int64_t rows = batch * dim;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t row = 0; row < rows; ++row) {
        const int64_t b = row / dim, ch = row % dim, g = ch / per_group;
        const float *u_row = u + row * L;
        const float *dt_row = delta + row * L;
        const float *a_row = A + ch * nstate;
        const float *b_grp = Bm + (b * groups + g) * nstate * L;
        const float *c_grp = Cm + (b * groups + g) * nstate * L;
        const float bias = delta_bias ? delta_bias[ch] : 0.0f;
        const float skip = D ? D[ch] : 0.0f;
        float h[MAX_STATE];
        for (int64_t n = 0; n < nstate; ++n) h[n] = 0.0f;
        for (int64_t l = 0; l < L; ++l) {
            const float dt = softplus_f(dt_row[l] + bias);
            const float x = u_row[l];
            const float dtx = dt * x;
            float acc = 0.0f;
            for (int64_t n = 0; n < nstate; ++n) {
                const float decay = expf(dt * a_row[n]);
                h[n] = decay * h[n] + dtx * b_grp[n * L + l];
                acc += h[n] * c_grp[n * L + l];
            }
            out[row * L + l] = acc + x * skip;
        }
    }
    return 0;
}

/* float64 arbiter: same inputs (float32 storage), all arithmetic in double, double output. */
int wm_oracle_selective_scan_f64(const float *u, const float *delta, const float *A,
                                 const float *Bm, const float *Cm, const float *D,
                                 const float *delta_bias, double *out, int64_t batch,
                                 int64_t dim, int64_t groups, int64_t nstate, int64_t L)
{
    if (nstate > MAX_STATE || groups <= 0 || dim % groups != 0) return -1;
    const int64_t per_group = dim / groups;
    const int64_t rows = batch * dim;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t row = 0; row < rows; ++row) {
        const int64_t b = row / dim, ch = row % dim, g = ch / per_group;
        const float *u_row = u + row * L;
        const float *dt_row = delta + row * L;
        const float *a_row = A + ch * nstate;
        const float *b_grp = Bm + (b * groups + g) * nstate * L;
        const float *c_grp = Cm + (b * groups + g) * nstate * L;
        const double bias = delta_bias ? (double)delta_bias[ch] : 0.0;
        const double skip = D ? (double)D[ch] : 0.0;
        double h[MAX_STATE];
        for (int64_t n = 0; n < nstate; ++n) h[n] = 0.0;
        for (int64_t l = 0; l < L; ++l) {
            const double dt = softplus_d((double)dt_row[l] + bias);
            const double x = (double)u_row[l];
            double acc = 0.0;
            for (int64_t n = 0; n < nstate; ++n) {
                const double decay = exp(dt * (double)a_row[n]);
                h[n] = decay * h[n] + dt * (double)b_grp[n * L + l] * x;
                acc += h[n] * (double)c_grp[n * L + l];
            }
            out[row * L + l] = acc + x * skip;
        }
    }
    return 0;
}

/* All-double variant (double storage too) used when the whole oracle model runs in float64. */
int wm_oracle_selective_scan_d64(const double *u, const double *delta, const double *A,
                                 const double *Bm, const double *Cm, const double *D,
                                 const double *delta_bias, double *out, int64_t batch,
                                 int64_t dim, int64_t groups, int64_t nstate, int64_t L)
{
    if (nstate > MAX_STATE || groups <= 0 || dim % groups != 0) return -1;
    const int64_t per_group = dim / groups;
    const int64_t rows = batch * dim;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t row = 0; row < rows; ++row) {
        const int64_t b = row / dim, ch = row % dim, g = ch / per_group;
        const double *u_row = u + row * L;
        const double *dt_row = delta + row * L;
        const double *a_row = A + ch * nstate;
        const double *b_grp = Bm + (b * groups + g) * nstate * L;
        const double *c_grp = Cm + (b * groups + g) * nstate * L;
        const double bias = delta_bias ? delta_bias[ch] : 0.0;
        const double skip = D ? D[ch] : 0.0;
        double h[MAX_STATE];
        for (int64_t n = 0; n < nstate; ++n) h[n] = 0.0;
        for (int64_t l = 0; l < L; ++l) {
            const double dt = softplus_d(dt_row[l] + bias);
            const double x = u_row[l];
            double acc = 0.0;
            for (int64_t n = 0; n < nstate; ++n) {
                const double decay = exp(dt * a_row[n]);
                h[n] = decay * h[n] + dt * b_grp[n * L + l] * x;
                acc += h[n] * c_grp[n * L + l];
            }
            out[row * L + l] = acc + x * skip;
        }
    }
    return 0;
}

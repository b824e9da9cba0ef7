/* oracle/kernels_oracle.c -- plain-C restatement of the reference's two covariance builders
 * (gpdemo/kernels.pyx:12-49 and :52-90; generated C: gpdemo/kernels.c:1401-1730, 1827-2140).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/apm_oracle.py).  Same scalar loop, same operation order, libm exp,
 * no FMA contraction (-ffp-contract=off), so on one machine the output is bit-identical to the Cython
 * module built from the unmodified .pyx (checked by tests/test_oracle_vs_golden.py against golden
 * vectors produced by that module).
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle/kernels_oracle.c -o oracle/libkernels_oracle.so -lm
 */
#include <math.h>

/* kernels.pyx:39-49 */
void oracle_iso_se_kernel(double* K, const double* X, const double* theta, double epsilon, int n, int D) {
    const double sigma = exp(theta[0]);
    const double tau = exp(theta[1]);
    for (int i = 0; i < n; i++) {
        K[(long)i * n + i] = sigma + epsilon;
        for (int j = 0; j < i; j++) {
            double acc = 0.;
            for (int k = 0; k < D; k++) {
                const double d = X[(long)i * D + k] - X[(long)j * D + k];
                acc += d * d;                                   /* (x)**2 == pow(x, 2.0) == x*x */
            }
            acc = sigma * exp(-acc / (2. * (tau * tau)));
            K[(long)i * n + j] = acc;
            K[(long)j * n + i] = acc;
        }
    }
}

/* kernels.pyx:81-90: the length-scale exp(theta[k+1]) is re-evaluated inside the innermost loop and
 * divides (kernels.c:2015, 2041) */
void oracle_ard_se_kernel(double* K, const double* X, const double* theta, double epsilon, int n, int D) {
    const double sigma = exp(theta[0]);
    for (int i = 0; i < n; i++) {
        K[(long)i * n + i] = sigma + epsilon;
        for (int j = 0; j < i; j++) {
            double acc = 0.;
            for (int k = 0; k < D; k++) {
                const double d = (X[(long)i * D + k] - X[(long)j * D + k]) / exp(theta[k + 1]);
                acc += d * d;
            }
            acc = sigma * exp(-acc / 2.);
            K[(long)i * n + j] = acc;
            K[(long)j * n + i] = acc;
        }
    }
}

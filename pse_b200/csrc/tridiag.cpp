// Host side of the Lanczos square root: eigen-decomposition of the small SPD tridiagonal T_m and
// c = T_m^{1/2} e_1 = W Lambda^{1/2} W^T e_1.
// Replaces LAPACKE_spteqr + the host loops at PSEv1/Brownian.cu:540-582 and :673-710 (m <= 100);
// done in double precision with an implicit-shift QL iteration, no LAPACK dependency.
#include <math.h>
#include <vector>
#include "../../include/pse_b200.h"

// diag[0..m), off[0..m-1) (off[i] couples i and i+1).  On success c[0..m) = T^{1/2} e_1.
int pse_tridiag_sqrt_e1(int m, const double* diag_in, const double* off_in, double* c, double* lambda_min_out) {
    if (m <= 0) return PSE_EINVAL;
    std::vector<double> d(diag_in, diag_in + m), e(m, 0.0), z((size_t)m * m, 0.0);
    for (int i = 0; i + 1 < m; ++i) e[i] = off_in[i];
    for (int i = 0; i < m; ++i) z[(size_t)i * m + i] = 1.0;  // z[row][col], columns become eigenvectors

    for (int l = 0; l < m; ++l) {
        int iter = 0;
        while (true) {
            int mm = l;
            for (; mm + 1 < m; ++mm) {
                double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                if (fabs(e[mm]) <= 2.3e-16 * dd) break;
            }
            if (mm == l) break;
            if (++iter > 200) return PSE_EEIGEN;
            // Wilkinson shift from the leading 2x2 of the unreduced block
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = hypot(g, 1.0);
            g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
            double s = 1.0, cth = 1.0, p = 0.0;
            int i = mm - 1;
            for (; i >= l; --i) {
                double f = s * e[i], b = cth * e[i];
                r = hypot(f, g);
                e[i + 1] = r;
                if (r == 0.0) { d[i + 1] -= p; e[mm] = 0.0; break; }
                s = f / r; cth = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * cth * b;
                p = s * r;
                d[i + 1] = g + p;
                g = cth * r - b;
                for (int k = 0; k < m; ++k) {  // accumulate the rotation into the eigenvectors
                    double zk1 = z[(size_t)k * m + i + 1], zk0 = z[(size_t)k * m + i];
                    z[(size_t)k * m + i + 1] = s * zk0 + cth * zk1;
                    z[(size_t)k * m + i] = cth * zk0 - s * zk1;
                }
            }
            if (r == 0.0 && i >= l) continue;
            d[l] -= p; e[l] = g; e[mm] = 0.0;
        }
    }
    double lmin = d[0];
    for (int k = 0; k < m; ++k) lmin = d[k] < lmin ? d[k] : lmin;
    if (lambda_min_out) *lambda_min_out = lmin;
    double lmax = 0; for (int k = 0; k < m; ++k) lmax = fabs(d[k]) > lmax ? fabs(d[k]) : lmax;
    if (lmin < -1e-6 * lmax) return PSE_EEIGEN;  // not positive definite (the reference exits here)
    for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int k = 0; k < m; ++k) {
            double lam = d[k] > 0 ? d[k] : 0.0;
            acc += z[(size_t)i * m + k] * sqrt(lam) * z[k];  // z[0][k] = first row = W^T e_1
        }
        c[i] = acc;
    }
    return PSE_OK;
}

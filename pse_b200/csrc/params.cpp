// Host-side parameter derivation: what Stokes::setParams computes once per run
// (reference: PSEv1/Stokes.cc:129-236 grid/Gaussian parameters, :309-319 table size and self
// term, :322-422 real-space table; PSEv1/Brownian.cu:826-829 spreading constants;
// PSEv1/Stokes.cc:102 seed hash).  The float/double mix of the reference expressions is
// followed operation by operation because grid size, P and eta must come out identical.
#include <cuda_runtime.h>
#include "../../include/pse_b200.h"
#include <math.h>
#include <algorithm>
#include <vector>

#define PSE_CHEB_DEG 8  // keep in step with real.cuh

static const double kPiRef = 3.1415926536;  // the literal the reference uses everywhere

static int round_up_235(int n) {
    // smallest 2^a 3^b 5^c in [8, 4096] that is >= n (a<13, b<8, c<6), PSEv1/Stokes.cc:153-199.
    // Like the reference, a request above 4096 is returned unchanged.
    std::vector<int> sizes;
    for (long p2 = 1; p2 <= 4096; p2 *= 2)
        for (long p3 = 1; p3 <= 2187; p3 *= 3)
            for (long p5 = 1; p5 <= 3125; p5 *= 5) {
                long v = p2 * p3 * p5;
                if (v >= 8 && v <= 4096) sizes.push_back((int)v);
            }
    std::sort(sizes.begin(), sizes.end());
    for (int v : sizes)
        if (n <= v) return v;
    return n;
}

uint32_t pse_hash_seed(uint32_t seed) {  // PSEv1/Stokes.cc:102
    seed = seed * 0x12345677u + 0x12345u;
    seed ^= (seed >> 16);
    seed *= 0x45679u;
    return seed;
}

extern "C" int pse_derive_params(const pse_config* c, pse_params* p) {
    if (!c || !p) return PSE_EINVAL;
    if (!(c->xi > 0.f) || !(c->error > 0.f) || !(c->error < 1.f) || c->N == 0) return PSE_EINVAL;
    if (!(c->box.Lx > 0.f) || !(c->box.Ly > 0.f) || !(c->box.Lz > 0.f)) return PSE_EINVAL;
    const float err = c->error, xi = c->xi;

    p->rcut = sqrtf(-logf(err)) / xi;                                // Stokes.cc:135
    p->kmax = int(2.0 * sqrtf(-logf(err)) * xi) + 1;                 // :138 (double product)
    const float L[3] = {c->box.Lx, c->box.Ly, c->box.Lz};
    int n[3];
    for (int d = 0; d < 3; ++d) {
        float kl = (float)p->kmax * L[d];                            // int*float is a float product
        n[d] = int((double)kl / (2.0 * kPiRef) * 2.0) + 1;           // :143-145
        n[d] = round_up_235(n[d]);
    }
    p->Nx = n[0]; p->Ny = n[1]; p->Nz = n[2];
    if ((long long)n[0] * n[1] * n[2] > 512LL * 512 * 512 && !(c->flags & PSE_FLAG_LIFT_GRID_CAP))
        return PSE_EGRID;                                            // :203-214

    const float gamma = c->max_strain, gamma2 = gamma * gamma;
    const float lambda = (float)(1.0 + gamma2 / 2.0 + gamma * sqrtf((float)(1.0 + gamma2 / 4.0)));  // :219
    p->hx = L[0] / (float)n[0]; p->hy = L[1] / (float)n[1]; p->hz = L[2] / (float)n[2];        // :222

    float gaussm = 1.0f;
    while (erfcf(gaussm / sqrtf((float)(2.0 * lambda))) > err) gaussm = (float)(gaussm + 0.01);    // :225-228
    p->gaussm = gaussm;
    int P = int(gaussm * gaussm / kPiRef) + 1;                       // :229
    P = std::min(P, std::min(n[0], std::min(n[1], n[2])));           // :231-233
    p->P = P;
    const float w = (float)((float)P * p->hx / 2.0);                 // :235
    const float xisq = xi * xi;
    p->eta = (float)((2.0 * w / gaussm) * (2.0 * w / gaussm) * xisq);  // :236

    p->dr = 0.001f;                                                  // :309
    p->ewald_n = (int)(p->rcut / p->dr - 1);                         // :310 (float arithmetic)

    const float pi12 = 1.77245385091f, aa = 1.0f;                    // :315-319
    const float axi = aa * xi, axi2 = axi * axi;
    p->self = (float)((1. + 4. * pi12 * axi * erfc(2. * axi) - exp(-4. * axi2)) / (4. * pi12 * axi * aa));

    p->quadW = p->hx * p->hy * p->hz;                                // Brownian.cu:826-829
    p->prefac = (float)((2.0 * xisq / kPiRef / p->eta) * sqrtf((float)(2.0 * xisq / kPiRef / p->eta)));
    p->expfac = (float)(2.0 * xisq / p->eta);
    p->seed_hashed = pse_hash_seed(c->seed);
    return PSE_OK;
}

// Real-space RPY Ewald kernel for spheres of radius a: M_real,ij = f(r) (I - rr) + g(r) rr.
// The reference tabulates closed forms in three branches (r > 2a, r == 2a, r < 2a) at
// PSEv1/Stokes.cc:348-406.  They are one expression: the erfc/exp part E(r) is common to all
// branches once erfc((r-2a)xi) is rewritten as 2 - erfc((2a-r)xi); what differs is the
// polynomial C(r), which is zero for overlapping spheres and
//   C_f = -1/a + a^2/(2r^3) + 3/(4r) + 9r/(32a^2),  C_g = -1/a - a^2/r^3 + 3/(2r) + 3r/(16a^2)
// otherwise (both vanish at r = 2a, so the r == 2a branch is the common limit).
void pse_rpy_real_fg(double r, double xi, double a, double* f_out, double* g_out) {
    const double sqrtpi = sqrt(3.141592653589793);
    const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2;
    const double a2 = a * a, a3 = a2 * a;
    const double x2 = xi * xi, x3 = x2 * xi, xm4 = 1.0 / (x2 * x2);
    const double ep = erfc((2 * a + r) * xi), em = erfc((2 * a - r) * xi), e0 = erfc(r * xi);
    const double gp = exp(-(2 * a + r) * (2 * a + r) * x2), gm = exp(-(r - 2 * a) * (r - 2 * a) * x2), g0 = exp(-r2 * x2);
    const bool apart = r >= 2 * a;

    // f: coefficient of (I - rr)
    const double A1 = 64 * a2 / r3 + 96 / r + (36 * r - 3 * xm4 / r3) / a2;
    double f = apart ? (-1 / a + a2 / (2 * r3) + 3 / (4 * r) + 9 * r / (32 * a2)) : 0.0;
    f += -3 * xm4 / (128 * a2 * r3);
    f += 3 * e0 * (-12 * r4 + xm4) / (128 * a2 * r3);
    f += ep * (128 / a + A1) / 256 + em * (128 / a - A1) / 256;
    f += 3 * g0 * (1 + 6 * r2 * x2) / (64 * a2 * sqrtpi * r2 * x3);
    const double poly_f = -3 * (r + 6 * r3 * x2);
    f += gp * (8 * r * a2 * x2 - 16 * a3 * x2 + a * (2 - 28 * r2 * x2) + poly_f) / (128 * a2 * sqrtpi * r3 * x3);
    f += gm * (8 * r * a2 * x2 + 16 * a3 * x2 - a * (2 - 28 * r2 * x2) + poly_f) / (128 * a2 * sqrtpi * r3 * x3);

    // g: coefficient of rr
    const double A2 = -64 * a2 / r3 + 96 / r + (12 * r + 3 * xm4 / r3) / a2;
    double g = apart ? (-1 / a - a2 / r3 + 3 / (2 * r) + 3 * r / (16 * a2)) : 0.0;
    g += 3 * xm4 / (64 * a2 * r3);
    g += -3 * e0 * (1 + 4 * r4 * x2 * x2) * xm4 / (64 * a2 * r3);
    g += ep * (64 / a + A2) / 128 + em * (64 / a - A2) / 128;
    g += 3 * g0 * (-1 + 2 * r2 * x2) / (32 * a2 * sqrtpi * r2 * x3);
    const double q = -1 + 8 * a2 * x2 + 2 * r2 * x2;
    g += -(2 * a + 3 * r) * gm * (q - 8 * a * r * x2) / (64 * a2 * sqrtpi * r3 * x3);
    g += (2 * a - 3 * r) * gp * (q + 8 * a * r * x2) / (64 * a2 * sqrtpi * r3 * x3);

    *f_out = f;
    *g_out = g;
}

extern "C" int pse_ewald_table(const pse_config* c, float* out) {
    pse_params p;
    int rc = pse_derive_params(c, &p);
    if (rc != PSE_OK && rc != PSE_EGRID) return rc;
    if (!out) return PSE_EINVAL;
    const int nR = p.ewald_n + 1;                                   // Stokes.cc:322
    const double dr = 0.001, xi = c->xi;
    for (int k = 0; k < nR; ++k) {
        double r = double(k) * dr + dr;                              // :346
        double f, g;
        pse_rpy_real_fg(r, xi, 1.0, &f, &g);
        out[4 * k + 0] = (float)f;
        out[4 * k + 1] = (float)g;
        out[4 * k + 2] = 0.f;
        out[4 * k + 3] = 0.f;
    }
    for (int k = 0; k + 1 < nR; ++k) {                               // :414-420
        out[4 * k + 2] = out[4 * (k + 1) + 0];
        out[4 * k + 3] = out[4 * (k + 1) + 1];
    }
    return PSE_OK;
}

// Single-interval polynomial form of the real-space functions for non-overlapping pairs (2a <= r <= rcut):
//     f(r) = exp(-xi^2 (r - 2a)^2) * Pf(t),   g(r) = exp(-xi^2 (r - 2a)^2) * Pg(t),   t = A / r + B  in [-1, 1].
// Dividing out the leading Gaussian and expanding in 1/r makes degree PSE_CHEB_DEG polynomials accurate to the
// rounding level of fp32 (a polynomial in r needs degree ~16 for the same).  Exact closed forms are sampled at the
// Chebyshev nodes of the 1/r interval; the monomial coefficients come from the Vandermonde system in double.
// out: A, B, -xi^2 log2(e), Pf[0..deg], Pg[0..deg]  (3 + 2 (deg + 1) floats).
// max_err_out: largest |fit - exact| / max|exact| over a scan evaluated in float exactly as the device does.
int pse_fit_rpy_cheb(double xi, double rcut, float* out, double* max_err_out) {
    const int n = PSE_CHEB_DEG + 1;
    const double lo = 2.0, pi = 3.14159265358979323846;
    if (!(rcut > lo + 1e-3)) return PSE_EINVAL;
    const double s_lo = 1.0 / rcut, s_hi = 1.0 / lo;
    const double A = 2.0 / (s_hi - s_lo), B = -(s_hi + s_lo) / (s_hi - s_lo);
    double M[PSE_CHEB_DEG + 1][PSE_CHEB_DEG + 3];
    for (int q = 0; q < n; ++q) {
        const double t = cos(pi * (2 * q + 1) / (2.0 * n));
        const double r = 1.0 / (0.5 * (s_hi + s_lo) + 0.5 * (s_hi - s_lo) * t);
        double f, g;
        pse_rpy_real_fg(r, xi, 1.0, &f, &g);
        const double e = exp(xi * xi * (r - lo) * (r - lo));
        double pw = 1.0;
        for (int c = 0; c < n; ++c) { M[q][c] = pw; pw *= t; }
        M[q][n] = f * e; M[q][n + 1] = g * e;
    }
    for (int c = 0; c < n; ++c) {  // Gauss-Jordan with partial pivoting, two right-hand sides
        int piv = c;
        for (int q = c + 1; q < n; ++q) if (fabs(M[q][c]) > fabs(M[piv][c])) piv = q;
        for (int j = 0; j < n + 2; ++j) std::swap(M[c][j], M[piv][j]);
        for (int q = 0; q < n; ++q) {
            if (q == c) continue;
            const double m = M[q][c] / M[c][c];
            for (int j = c; j < n + 2; ++j) M[q][j] -= m * M[c][j];
        }
    }
    out[0] = (float)A; out[1] = (float)B; out[2] = (float)(-xi * xi * 1.4426950408889634);
    float* cf = out + 3; float* cg = out + 3 + n;
    for (int c = 0; c < n; ++c) { cf[c] = (float)(M[c][n] / M[c][c]); cg[c] = (float)(M[c][n + 1] / M[c][c]); }
    double ef = 0, eg = 0, mf = 0, mg = 0;
    const int S = 4000;
    for (int k = 0; k <= S; ++k) {
        const double r = lo + (rcut - lo) * k / S;
        double f, g;
        pse_rpy_real_fg(r, xi, 1.0, &f, &g);
        const float rf = (float)r, inv = 1.0f / rf, t = out[0] * inv + out[1];
        float pf = cf[n - 1], pg = cg[n - 1];
        for (int c = n - 2; c >= 0; --c) { pf = pf * t + cf[c]; pg = pg * t + cg[c]; }
        const float d = rf - 2.0f, e = exp2f(out[2] * d * d);
        ef = std::max(ef, fabs((double)(pf * e) - f)); eg = std::max(eg, fabs((double)(pg * e) - g));
        mf = std::max(mf, fabs(f)); mg = std::max(mg, fabs(g));
    }
    if (max_err_out) *max_err_out = std::max(ef / mf, eg / mg);
    return PSE_OK;
}

/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (plain C + OpenMP) of the reference's algorithm for
 * one PSE Brownian-dynamics step.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it; the product never does.
 *
 * Parity status: PINNED.  Checked against (a) golden vectors generated from the reference's own
 * source expressions (tests/golden/, made by tests/golden/make_golden.py, which compiles the
 * table expressions of PSEv1/Stokes.cc:348-406 where they lie), (b) the reference's own CUDA
 * kernels compiled unmodified (oracle/_ref/libpse_ref.so) on the GPU box, (c) an independent
 * dense double-precision Ewald sum (orc_dense_mobility below).  HOOMD pieces that are not in
 * /root/reference (BoxDim, neighbour list, Saru) are "parity unpinned" against HOOMD itself and
 * are restated from their call sites.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Build: gcc -O3 -ffp-contract=off -fopenmp (no FMA contraction: float results must be
 * reproducible operation by operation).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_REF 3.1415926536 /* literal used throughout the reference */

typedef struct {
    int N;
    float Lx, Ly, Lz, xy;
    float xi, error, max_strain;
    int ref_pi; /* 1: 2*3.1416926536 wave vectors (PSEv1/Helper.cu:313-315), 0: exact pi */
} orc_config;

typedef struct {
    int Nx, Ny, Nz, P, kmax, ewald_n;
    float rcut, dr, gaussm, eta, hx, hy, hz, self, quadW, prefac, expfac;
} orc_params;

/* ------------------------------------------------------------------------------------------
 * Stokes::setParams — PSEv1/Stokes.cc:129-236 (grid, Gaussian), :309-319 (table size, self),
 * PSEv1/Brownian.cu:826-829 (spreading constants).  Float/double mix as written there.
 * ---------------------------------------------------------------------------------------- */
static int next_235(int n) { /* PSEv1/Stokes.cc:153-199: first 2^a 3^b 5^c in [8,4096] >= n */
    for (int v = (n < 8 ? 8 : n); v <= 4096; ++v) {
        int r = v, a = 0, b = 0, c = 0;
        while (r % 2 == 0) { r /= 2; ++a; }
        while (r % 3 == 0) { r /= 3; ++b; }
        while (r % 5 == 0) { r /= 5; ++c; }
        if (r == 1 && a < 13 && b < 8 && c < 6) return v;
    }
    return n;
}

int orc_derive_params(const orc_config* c, orc_params* p) {
    float err = c->error, xi = c->xi;
    p->rcut = sqrtf(-logf(err)) / xi;                      /* :135 */
    p->kmax = (int)(2.0 * sqrtf(-logf(err)) * xi) + 1;     /* :138 */
    float L[3] = {c->Lx, c->Ly, c->Lz};
    int n[3];
    for (int d = 0; d < 3; ++d) {
        float kl = (float)p->kmax * L[d];
        n[d] = next_235((int)((double)kl / (2.0 * PI_REF) * 2.0) + 1); /* :143-145 */
    }
    p->Nx = n[0]; p->Ny = n[1]; p->Nz = n[2];
    if ((long long)n[0] * n[1] * n[2] > 512LL * 512 * 512) return -4; /* :203-214 */
    float gamma = c->max_strain, gamma2 = gamma * gamma;
    float lambda = (float)(1.0 + gamma2 / 2.0 + gamma * sqrtf((float)(1.0 + gamma2 / 4.0))); /* :219 */
    p->hx = L[0] / (float)n[0]; p->hy = L[1] / (float)n[1]; p->hz = L[2] / (float)n[2];
    float gm = 1.0f;
    while (erfcf(gm / sqrtf((float)(2.0 * lambda))) > err) gm = (float)(gm + 0.01); /* :225-228 */
    p->gaussm = gm;
    int P = (int)(gm * gm / PI_REF) + 1;                   /* :229 */
    if (P > n[0]) P = n[0];
    if (P > n[1]) P = n[1];
    if (P > n[2]) P = n[2];
    p->P = P;
    float w = (float)((float)P * p->hx / 2.0);             /* :235 */
    float xisq = xi * xi;
    p->eta = (float)((2.0 * w / gm) * (2.0 * w / gm) * xisq); /* :236 */
    p->dr = 0.001f;
    p->ewald_n = (int)(p->rcut / p->dr - 1);               /* :310 */
    float pi12 = 1.77245385091f, aa = 1.0f, axi = aa * xi, axi2 = axi * axi; /* :315-319 */
    p->self = (float)((1. + 4. * pi12 * axi * erfc(2. * axi) - exp(-4. * axi2)) / (4. * pi12 * axi * aa));
    p->quadW = p->hx * p->hy * p->hz;                      /* Brownian.cu:826 */
    p->prefac = (float)((2.0 * xisq / PI_REF / p->eta) * sqrtf((float)(2.0 * xisq / PI_REF / p->eta)));
    p->expfac = (float)(2.0 * xisq / p->eta);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Real-space table — PSEv1/Stokes.cc:322-422.  M_real = f (I - rr) + g rr for radius a = 1.
 * The three branches of :348-406 are restated as free-space RPY minus the smooth (wave-space)
 * part: f = f_rpy - f_smooth, g = g_rpy - g_smooth, with the erfc/exp closed form of the
 * smooth part valid for every r (the branches differ only by the RPY polynomial).
 * ---------------------------------------------------------------------------------------- */
void orc_real_fg(double r, double xi, double* f, double* g) {
    const double a = 1.0, spi = sqrt(3.141592653589793);
    double r2 = r * r, r3 = r2 * r, r4 = r2 * r2, x2 = xi * xi, x3 = x2 * xi, xm4 = 1.0 / (x2 * x2);
    double ep = erfc((2 * a + r) * xi), em = erfc((2 * a - r) * xi), e0 = erfc(r * xi);
    double gp = exp(-(2 * a + r) * (2 * a + r) * x2), gm = exp(-(2 * a - r) * (2 * a - r) * x2), g0 = exp(-r2 * x2);
    /* free-space RPY (6 pi eta a = 1): PSEv1/Stokes.cc:350,384 polynomial heads */
    double f_rpy = r >= 2 * a ? 3 / (4 * r) + 1 / (2 * r3) : 1 - 9 * r / 32;
    double g_rpy = r >= 2 * a ? 3 / (2 * r) - 1 / r3 : 1 - 3 * r / 16;
    /* minus-smooth part, common to all branches */
    double A1 = 64 / r3 + 96 / r + 36 * r - 3 * xm4 / r3;
    double fs = -(1 - 9 * r / 32) - 3 * xm4 / (128 * r3) + 3 * e0 * (-12 * r4 + xm4) / (128 * r3) +
                (ep * (128 + A1) + em * (128 - A1)) / 256 + 3 * g0 * (1 + 6 * r2 * x2) / (64 * spi * r2 * x3) +
                (gp * (8 * r * x2 - 16 * x2 + (2 - 28 * r2 * x2) - 3 * (r + 6 * r3 * x2)) +
                 gm * (8 * r * x2 + 16 * x2 - (2 - 28 * r2 * x2) - 3 * (r + 6 * r3 * x2))) / (128 * spi * r3 * x3);
    double A2 = -64 / r3 + 96 / r + 12 * r + 3 * xm4 / r3;
    double q = -1 + 8 * x2 + 2 * r2 * x2;
    double gs = -(1 - 3 * r / 16) + 3 * xm4 / (64 * r3) - 3 * e0 * (1 + 4 * r4 * x2 * x2) * xm4 / (64 * r3) +
                (ep * (64 + A2) + em * (64 - A2)) / 128 + 3 * g0 * (-1 + 2 * r2 * x2) / (32 * spi * r2 * x3) +
                (-(2 + 3 * r) * gm * (q - 8 * r * x2) + (2 - 3 * r) * gp * (q + 8 * r * x2)) / (64 * spi * r3 * x3);
    *f = f_rpy + fs;
    *g = g_rpy + gs;
}

int orc_table(const orc_params* p, float xi, float* out) { /* PSEv1/Stokes.cc:333-420 */
    int nR = p->ewald_n + 1;
    double dr = 0.001;
    for (int k = 0; k < nR; ++k) {
        double r = (double)k * dr + dr, f, g;
        orc_real_fg(r, (double)xi, &f, &g);
        out[4 * k] = (float)f; out[4 * k + 1] = (float)g; out[4 * k + 2] = 0.f; out[4 * k + 3] = 0.f;
    }
    for (int k = 0; k + 1 < nR; ++k) { out[4 * k + 2] = out[4 * k + 4]; out[4 * k + 3] = out[4 * k + 5]; }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Box arithmetic — HOOMD BoxDim semantics as used at PSEv1/Mobility.cu:173,238 and
 * PSEv1/Stokes.cu:185; same operation order as pse_b200/csrc/box.cuh (definition of record).
 * ---------------------------------------------------------------------------------------- */
typedef struct { float Lx, Ly, Lz, ix, iy, iz, lox, loy, loz, hix, hiy, hiz, xy; } obox;
static obox make_box(const orc_config* c) {
    obox b;
    b.hix = c->Lx / 2.0f; b.hiy = c->Ly / 2.0f; b.hiz = c->Lz / 2.0f;
    b.lox = -b.hix; b.loy = -b.hiy; b.loz = -b.hiz;
    b.Lx = b.hix - b.lox; b.Ly = b.hiy - b.loy; b.Lz = b.hiz - b.loz;
    b.ix = 1.0f / b.Lx; b.iy = 1.0f / b.Ly; b.iz = 1.0f / b.Lz;
    b.xy = c->xy;
    return b;
}
static void make_fraction(const obox* b, const float* p, float* f) {
    float dx = p[0] - b->lox, dy = p[1] - b->loy, dz = p[2] - b->loz;
    dx = dx - b->xy * p[1];
    f[0] = dx * b->ix; f[1] = dy * b->iy; f[2] = dz * b->iz;
}
static void min_image(const obox* b, float* w) {
    float img = rintf(w[2] * b->iz);
    w[2] = w[2] - b->Lz * img;
    img = rintf(w[1] * b->iy);
    w[1] = w[1] - b->Ly * img;
    w[0] = w[0] - (b->Ly * b->xy) * img;
    img = rintf(w[0] * b->ix);
    w[0] = w[0] - b->Lx * img;
}
static void wrap_pos(const obox* b, float* w, int* img) {
    float tilt = b->xy * w[1];
    if (w[0] >= b->hix + tilt) { w[0] = w[0] - b->Lx; img[0]++; }
    else if (w[0] < b->lox + tilt) { w[0] = w[0] + b->Lx; img[0]--; }
    if (w[1] >= b->hiy) { w[1] = w[1] - b->Ly; w[0] = w[0] - b->Ly * b->xy; img[1]++; }
    else if (w[1] < b->loy) { w[1] = w[1] + b->Ly; w[0] = w[0] + b->Ly * b->xy; img[1]--; }
    if (w[2] >= b->hiz) { w[2] = w[2] - b->Lz; img[2]++; }
    else if (w[2] < b->loz) { w[2] = w[2] + b->Lz; img[2]--; }
}

/* ------------------------------------------------------------------------------------------
 * Neighbour list — contract of SURVEY.md §8c for HOOMD's NeighborListGPUBinned
 * (PSEv1/integrate.py:58-83): full list, |minImage(r_i - r_j)|^2 < rlist^2, rows ascending.
 * Brute force O(N^2); two passes (count, fill).  |d|^2 = ((x*x + y*y) + z*z) in float.
 * ---------------------------------------------------------------------------------------- */
int orc_nlist_bruteforce(const orc_config* c, const float* pos4, float rlist, uint32_t* nn, uint32_t* head,
                         uint32_t* nl, size_t cap, size_t* nnz) {
    obox b = make_box(c);
    int N = c->N;
    float rl2 = rlist * rlist;
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < N; ++i) {
            uint32_t cnt = 0;
            for (int j = 0; j < N; ++j) {
                if (j == i) continue;
                float d[3] = {pos4[4 * i] - pos4[4 * j], pos4[4 * i + 1] - pos4[4 * j + 1], pos4[4 * i + 2] - pos4[4 * j + 2]};
                min_image(&b, d);
                float r2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
                if (r2 < rl2) {
                    if (pass == 1) nl[head[i] + cnt] = (uint32_t)j;
                    ++cnt;
                }
            }
            if (pass == 0) nn[i] = cnt;
        }
        if (pass == 0) {
            size_t tot = 0;
            for (int i = 0; i < N; ++i) { head[i] = (uint32_t)tot; tot += nn[i]; }
            *nnz = tot;
            if (tot > cap || !nl) return nl ? -7 : 0;
        }
    }
    return 0;
}

/* Cell-list build of the same list for large N (CPU baseline); identical membership test. */
int orc_nlist_cells(const orc_config* c, const float* pos4, float rlist, uint32_t* nn, uint32_t* head, uint32_t* nl,
                    size_t cap, size_t* nnz) {
    obox b = make_box(c);
    int N = c->N;
    float rl2 = rlist * rlist;
    float reach[3] = {rlist * sqrtf(1.f + c->xy * c->xy) / b.Lx * 1.0001f + 1e-6f, rlist / b.Ly * 1.0001f + 1e-6f,
                      rlist / b.Lz * 1.0001f + 1e-6f};
    int nc[3] = {(int)(b.Lx / rlist), (int)(b.Ly / rlist), (int)(b.Lz / rlist)};
    for (int d = 0; d < 3; ++d) if (nc[d] < 1) nc[d] = 1;
    size_t ncell = (size_t)nc[0] * nc[1] * nc[2];
    uint32_t* cstart = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
    uint32_t* cell_of = (uint32_t*)malloc(sizeof(uint32_t) * N);
    uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * N);
    float* frac = (float*)malloc(sizeof(float) * 3 * N);
    for (int i = 0; i < N; ++i) {
        float f[3];
        make_fraction(&b, pos4 + 4 * i, f);
        int cc[3];
        for (int d = 0; d < 3; ++d) {
            f[d] -= floorf(f[d]);
            cc[d] = (int)floorf(f[d] * nc[d]);
            if (cc[d] >= nc[d]) cc[d] = nc[d] - 1;
            if (cc[d] < 0) cc[d] = 0;
            frac[3 * i + d] = f[d];
        }
        cell_of[i] = (uint32_t)((cc[0] * nc[1] + cc[1]) * nc[2] + cc[2]);
        cstart[cell_of[i] + 1]++;
    }
    for (size_t k = 0; k < ncell; ++k) cstart[k + 1] += cstart[k];
    uint32_t* fill = (uint32_t*)calloc(ncell, sizeof(uint32_t));
    for (int i = 0; i < N; ++i) order[cstart[cell_of[i]] + fill[cell_of[i]]++] = (uint32_t)i; /* ascending id per cell */
    free(fill);
    int rc = 0;
    for (int pass = 0; pass < 2 && rc == 0; ++pass) {
#pragma omp parallel for schedule(dynamic, 256)
        for (int i = 0; i < N; ++i) {
            uint32_t cnt = 0;
            uint32_t* row = pass ? nl + head[i] : NULL;
            int lo[3], len[3];
            for (int d = 0; d < 3; ++d) {
                int c0 = (int)floorf((frac[3 * i + d] - reach[d]) * nc[d]), c1 = (int)floorf((frac[3 * i + d] + reach[d]) * nc[d]);
                len[d] = c1 - c0 + 1;
                if (len[d] > nc[d]) len[d] = nc[d];
                lo[d] = ((c0 % nc[d]) + nc[d]) % nc[d];
            }
            for (int tx = 0; tx < len[0]; ++tx)
                for (int ty = 0; ty < len[1]; ++ty)
                    for (int tz = 0; tz < len[2]; ++tz) {
                        int cx = (lo[0] + tx) % nc[0], cy = (lo[1] + ty) % nc[1], cz = (lo[2] + tz) % nc[2];
                        size_t cell = ((size_t)cx * nc[1] + cy) * nc[2] + cz;
                        for (uint32_t s = cstart[cell]; s < cstart[cell + 1]; ++s) {
                            int j = (int)order[s];
                            if (j == i) continue;
                            float d[3] = {pos4[4 * i] - pos4[4 * j], pos4[4 * i + 1] - pos4[4 * j + 1], pos4[4 * i + 2] - pos4[4 * j + 2]};
                            min_image(&b, d);
                            float r2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
                            if (r2 < rl2) {
                                if (pass) { /* insertion keeps the row ascending */
                                    uint32_t k = cnt;
                                    while (k > 0 && row[k - 1] > (uint32_t)j) { row[k] = row[k - 1]; --k; }
                                    row[k] = (uint32_t)j;
                                }
                                ++cnt;
                            }
                        }
                    }
            if (!pass) nn[i] = cnt;
        }
        if (!pass) {
            size_t tot = 0;
            for (int i = 0; i < N; ++i) { head[i] = (uint32_t)tot; tot += nn[i]; }
            *nnz = tot;
            if (!nl) break;
            if (tot > cap) rc = -7;
        }
    }
    free(cstart); free(cell_of); free(order); free(frac);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Particle -> grid assignment — PSEv1/Mobility.cu:173-219 (spread) == :380-425 (contract).
 * out3 = wrapped (x_inp, y_inp, z_inp) at t = 0.
 * ---------------------------------------------------------------------------------------- */
static void support_origin(const obox* b, const orc_params* p, const float* pos, int* o) {
    float f[3];
    make_fraction(b, pos, f);
    float n[3] = {(float)p->Nx, (float)p->Ny, (float)p->Nz};
    int odd = p->P % 2, half = p->P / 2;
    for (int d = 0; d < 3; ++d) {
        float s = f[d] * n[d];
        int x = (int)s;
        o[d] = x - half + 1 - odd * ((s - (float)x) < 0.5f);
    }
}
static int wrap_node(int i, int n) { return i < 0 ? i + n : (i > n - 1 ? i - n : i); }

int orc_grid_index(const orc_config* c, const orc_params* p, const float* pos4, int* out3) {
    obox b = make_box(c);
    for (int i = 0; i < c->N; ++i) {
        int o[3];
        support_origin(&b, p, pos4 + 4 * i, o);
        out3[3 * i] = wrap_node(o[0], p->Nx); out3[3 * i + 1] = wrap_node(o[1], p->Ny); out3[3 * i + 2] = wrap_node(o[2], p->Nz);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Real-space SpMV — gpu_stokes_Mreal_kernel, PSEv1/Mobility.cu:612-686.
 * ---------------------------------------------------------------------------------------- */
int orc_mreal(const orc_config* c, const orc_params* p, const float* table, const float* pos4, const float* F4,
              const uint32_t* nn, const uint32_t* head, const uint32_t* nl, float* U4) {
    obox b = make_box(c);
    float mind2 = p->dr * p->dr, maxd2 = p->rcut * p->rcut; /* :635-636 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < c->N; ++i) {
        const float* pi = pos4 + 4 * i;
        float u[3] = {p->self * F4[4 * i], p->self * F4[4 * i + 1], p->self * F4[4 * i + 2]}; /* :632 */
        for (uint32_t k = 0; k < nn[i]; ++k) {
            uint32_t j = nl[head[i] + k];
            float r[3] = {pi[0] - pos4[4 * j], pi[1] - pos4[4 * j + 1], pi[2] - pos4[4 * j + 2]};
            min_image(&b, r);
            float d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            if (d2 < maxd2 && d2 >= mind2) { /* :652 */
                float dist = sqrtf(d2);
                const float* Fj = F4 + 4 * j;
                int ri = (int)floorf((float)p->ewald_n * (dist - p->dr) / (p->rcut - p->dr)); /* :661 */
                const float* t = table + 4 * ri;
                float fac = dist / p->dr - (float)ri - 1.0f; /* :667 */
                float Imrr = t[0] + (t[2] - t[0]) * fac, rr = t[1] + (t[3] - t[1]) * fac;
                float rdotf = (r[0] * Fj[0] + r[1] * Fj[1] + r[2] * Fj[2]) / d2; /* :673 */
                for (int d = 0; d < 3; ++d) u[d] += Imrr * Fj[d] + (rr - Imrr) * rdotf * r[d];
            }
        }
        U4[4 * i] = u[0]; U4[4 * i + 1] = u[1]; U4[4 * i + 2] = u[2]; U4[4 * i + 3] = 0.f;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * FFT: in-place complex transform of length n = 2^a 3^b 5^c (unnormalised, sign = -1 forward,
 * +1 inverse), applied along the three axes of a z-fastest grid.  Stands in for cufftExecC2C
 * (PSEv1/Brownian.cu:844-846,867-869).
 * ---------------------------------------------------------------------------------------- */
typedef struct { double re, im; } cpx;
static void fft_rec(int n, int stride, const cpx* in, cpx* out, int sign, const cpx* tw, int twstride) {
    if (n == 1) { out[0] = in[0]; return; }
    int radix = (n % 2 == 0) ? 2 : (n % 3 == 0) ? 3 : (n % 5 == 0) ? 5 : n;
    int m = n / radix;
    for (int q = 0; q < radix; ++q) fft_rec(m, stride * radix, in + q * stride, out + q * m, sign, tw, twstride * radix);
    cpx tmp[8];
    for (int k = 0; k < m; ++k) {
        for (int s = 0; s < radix; ++s) { /* output index k + s*m */
            double ar = 0, ai = 0;
            for (int q = 0; q < radix; ++q) {
                long idx = ((long)q * (k + (long)s * m)) % n; /* exponent modulo n */
                cpx w = tw[idx * twstride];
                double wi = sign < 0 ? w.im : -w.im;
                cpx x = out[q * m + k];
                ar += x.re * w.re - x.im * wi;
                ai += x.re * wi + x.im * w.re;
            }
            tmp[s].re = ar; tmp[s].im = ai;
        }
        for (int s = 0; s < radix; ++s) out[k + s * m] = tmp[s];
    }
}
static void fft_axis(cpx* g, int Nx, int Ny, int Nz, int axis, int sign) {
    int n = axis == 0 ? Nx : axis == 1 ? Ny : Nz;
    cpx* tw = (cpx*)malloc(sizeof(cpx) * n);
    for (int k = 0; k < n; ++k) { tw[k].re = cos(2 * M_PI * k / n); tw[k].im = -sin(2 * M_PI * k / n); }
    long stride = axis == 0 ? (long)Ny * Nz : axis == 1 ? Nz : 1;
    long nlines = (long)Nx * Ny * Nz / n;
#pragma omp parallel
    {
        cpx* a = (cpx*)malloc(sizeof(cpx) * n);
        cpx* o = (cpx*)malloc(sizeof(cpx) * n);
#pragma omp for schedule(static)
        for (long l = 0; l < nlines; ++l) {
            long base;
            if (axis == 0) base = l;                                   /* l = y*Nz + z */
            else if (axis == 1) base = (l / Nz) * (long)Ny * Nz + l % Nz; /* l = x*Nz + z */
            else base = l * (long)Nz;
            for (int k = 0; k < n; ++k) a[k] = g[base + k * stride];
            fft_rec(n, 1, a, o, sign, tw, 1);
            for (int k = 0; k < n; ++k) g[base + k * stride] = o[k];
        }
        free(a); free(o);
    }
    free(tw);
}
static void fft3(cpx* g, int Nx, int Ny, int Nz, int sign) {
    fft_axis(g, Nx, Ny, Nz, 2, sign); fft_axis(g, Nx, Ny, Nz, 1, sign); fft_axis(g, Nx, Ny, Nz, 0, sign);
}

/* wave vector + scaling of node (i,j,k): gpu_stokes_SetGridk_kernel, PSEv1/Helper.cu:300-329 */
static void gridk_node(const orc_config* c, const orc_params* p, const obox* b, int i, int j, int k, float* kv /*4*/) {
    float fx = (float)((i < (p->Nx + 1) / 2) ? i : i - p->Nx);
    float fy = ((float)((j < (p->Ny + 1) / 2) ? j : j - p->Ny) - b->xy * fx * b->Ly / b->Lx) / b->Ly;
    fx = fx / b->Lx;
    float fz = (float)((k < (p->Nz + 1) / 2) ? k : k - p->Nz) / b->Lz;
    double twopi = c->ref_pi ? 2.0 * 3.1416926536 : 2.0 * 3.14159265358979323846;
    kv[0] = (float)(fx * twopi); kv[1] = (float)(fy * twopi); kv[2] = (float)(fz * twopi);
    float k2 = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
    float xisq = c->xi * c->xi;
    if (i == 0 && j == 0 && k == 0) kv[3] = 0.f;
    else kv[3] = (float)(6.0 * PI_REF * (1.0 + k2 / 4.0 / xisq) * expf(-(1 - p->eta) * k2 / 4.0 / xisq) / k2 /
                         (float)(p->Nx * p->Ny * p->Nz));
}

/* ------------------------------------------------------------------------------------------
 * Wave-space velocity with optional random modes — the wave half of
 * gpu_stokes_CombinedMobilityBrownian_wrap, PSEv1/Brownian.cu:831-872:
 * spread (Mobility.cu:114-252), FFT, Green (Mobility.cu:264-299), + random modes
 * (Brownian.cu:153-345) from injected uniforms u_grid[G][6] (NULL = none), inverse FFT,
 * contract (Mobility.cu:325-477).  Each conjugate pair is generated once: the pair
 * belongs to the node the reference's rule processes, or to the smaller index when it processes both.
 * ---------------------------------------------------------------------------------------- */
static int ref_processed(const orc_params* p, int ii, int jj, int kk) { /* Brownian.cu:210-215 */
    return !(2 * kk >= p->Nz + 1) && !((kk == 0) && (2 * jj >= p->Ny + 1)) && !((kk == 0) && (jj == 0) && (2 * ii >= p->Nx + 1)) &&
           !((kk == 0) && (jj == 0) && (ii == 0));
}
static float gauss_w(const obox* b, const orc_params* p, int ix, int iy, int iz, const float* pos, float pref) {
    float g[3] = {p->hx * (float)ix - b->Lx * 0.5f, p->hy * (float)iy - b->Ly * 0.5f, p->hz * (float)iz - b->Lz * 0.5f};
    g[0] = g[0] + b->xy * g[1]; /* Mobility.cu:230 */
    float r[3] = {g[0] - pos[0], g[1] - pos[1], g[2] - pos[2]};
    min_image(b, r);
    return pref * expf(-p->expfac * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]));
}

int orc_mwave(const orc_config* c, const orc_params* p, const float* pos4, const float* F4, float* U4, int do_det,
              const float* u_grid, float noise_fac) {
    obox b = make_box(c);
    int Nx = p->Nx, Ny = p->Ny, Nz = p->Nz, P = p->P, N = c->N;
    size_t G = (size_t)Nx * Ny * Nz;
    cpx* g[3];
    for (int d = 0; d < 3; ++d) g[d] = (cpx*)calloc(G, sizeof(cpx));
    if (do_det) {
        for (int i = 0; i < N; ++i) { /* serial scatter: deterministic summation order */
            int o[3];
            support_origin(&b, p, pos4 + 4 * i, o);
            for (int tx = 0; tx < P; ++tx) for (int ty = 0; ty < P; ++ty) for (int tz = 0; tz < P; ++tz) {
                int ix = wrap_node(o[0] + tx, Nx), iy = wrap_node(o[1] + ty, Ny), iz = wrap_node(o[2] + tz, Nz);
                float w = gauss_w(&b, p, ix, iy, iz, pos4 + 4 * i, p->prefac);
                size_t idx = ((size_t)ix * Ny + iy) * Nz + iz;
                for (int d = 0; d < 3; ++d) g[d][idx].re += (double)(w * F4[4 * i + d]);
            }
        }
        for (int d = 0; d < 3; ++d) fft3(g[d], Nx, Ny, Nz, -1);
    }
#pragma omp parallel for schedule(static)
    for (long t = 0; t < (long)G; ++t) {
        int i = (int)(t / ((long)Ny * Nz)), j = (int)((t / Nz) % Ny), k = (int)(t % Nz);
        float kv[4];
        gridk_node(c, p, &b, i, j, k, kv);
        float ksq = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2], kk = sqrtf(ksq);
        double out[3][2] = {{0, 0}, {0, 0}, {0, 0}};
        if (t != 0 && do_det) { /* Mobility.cu:277-295 */
            float sinc = sinf(kk) / kk, B = kv[3] * sinc * sinc;
            for (int part = 0; part < 2; ++part) {
                double fx = part ? g[0][t].im : g[0][t].re, fy = part ? g[1][t].im : g[1][t].re, fz = part ? g[2][t].im : g[2][t].re;
                double kdF = (kv[0] * fx + kv[1] * fy + kv[2] * fz) / ksq;
                out[0][part] = (fx - kv[0] * kdF) * B; out[1][part] = (fy - kv[1] * kdF) * B; out[2][part] = (fz - kv[2] * kdF) * B;
            }
        }
        if (t != 0 && u_grid) { /* Brownian.cu:176-341, pair generated once */
            int mi = i ? Nx - i : 0, mj = j ? Ny - j : 0, mk = k ? Nz - k : 0;
            long mt = ((long)mi * Ny + mj) * Nz + mk;
            int selfc = (mt == t);
            int me = ref_processed(p, i, j, k), other = ref_processed(p, mi, mj, mk);
            int own = selfc || (me && (!other || t < mt));
            const float* u = u_grid + 6 * (own ? t : mt);
            const float a = 1.2247448713915889f;
            float re[3], im[3];
            for (int d = 0; d < 3; ++d) { re[d] = fmaf(2 * a, u[d], -a); im[d] = fmaf(2 * a, u[3 + d], -a); }
            if (selfc) for (int d = 0; d < 3; ++d) { re[d] *= 1.4142135623730951f; im[d] = 0.f; }
            else if (!own) for (int d = 0; d < 3; ++d) im[d] = -im[d];
            float B12 = sqrtf(kv[3]) * (sinf(kk) / kk);
            float kdr = (kv[0] * re[0] + kv[1] * re[1] + kv[2] * re[2]) / ksq, kdi = (kv[0] * im[0] + kv[1] * im[1] + kv[2] * im[2]) / ksq;
            for (int d = 0; d < 3; ++d) {
                out[d][0] += noise_fac * (re[d] - kv[d] * kdr) * B12;
                out[d][1] += noise_fac * (im[d] - kv[d] * kdi) * B12;
            }
        }
        for (int d = 0; d < 3; ++d) { g[d][t].re = out[d][0]; g[d][t].im = out[d][1]; }
    }
    for (int d = 0; d < 3; ++d) fft3(g[d], Nx, Ny, Nz, +1);
    float pref = p->quadW * p->prefac; /* Brownian.cu:872 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < N; ++i) {
        int o[3];
        support_origin(&b, p, pos4 + 4 * i, o);
        double acc[3] = {0, 0, 0};
        for (int tx = 0; tx < P; ++tx) for (int ty = 0; ty < P; ++ty) for (int tz = 0; tz < P; ++tz) {
            int ix = wrap_node(o[0] + tx, Nx), iy = wrap_node(o[1] + ty, Ny), iz = wrap_node(o[2] + tz, Nz);
            float w = gauss_w(&b, p, ix, iy, iz, pos4 + 4 * i, pref);
            size_t idx = ((size_t)ix * Ny + iy) * Nz + iz;
            for (int d = 0; d < 3; ++d) acc[d] += (double)w * g[d][idx].re;
        }
        U4[4 * i] = (float)acc[0]; U4[4 * i + 1] = (float)acc[1]; U4[4 * i + 2] = (float)acc[2]; U4[4 * i + 3] = 0.f;
    }
    for (int d = 0; d < 3; ++d) free(g[d]);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Symmetric tridiagonal eigen-solve (cyclic Jacobi on the dense m x m matrix, m <= 100) and
 * c = W Lambda^{1/2} W^T e_1.  Stands in for LAPACKE_spteqr + PSEv1/Brownian.cu:568-582.
 * ---------------------------------------------------------------------------------------- */
/* The reference's own route (PSEv1/Brownian.cu:540-582): LAPACKE_spteqr(LAPACK_ROW_MAJOR, 'I', m, alpha, &beta[1], W, m)
 * in single precision, then Tm = W (Lambda^{1/2} W^T e_1).  The LAPACKE entry point is bound at run time from the
 * OpenBLAS that ships with scipy (orc_set_spteqr; oracle/oraclewrap.py) - the same library the compiled reference links. */
typedef int (*orc_spteqr_fn)(int layout, char compz, int n, float* d, float* e, float* z, int ldz);
static orc_spteqr_fn g_spteqr = 0;
void orc_set_spteqr(void* fn) { g_spteqr = (orc_spteqr_fn)fn; }
int orc_have_spteqr(void) { return g_spteqr != 0; }
static int tridiag_sqrt_e1_lapacke(int m, const double* alpha, const double* beta, double* cvec) {
    float d[101], e[101];
    float* W = (float*)calloc((size_t)m * m, sizeof(float));
    for (int i = 0; i < m; ++i) d[i] = (float)alpha[i];
    for (int i = 1; i < m; ++i) e[i - 1] = (float)beta[i];
    int info = g_spteqr(101 /* LAPACK_ROW_MAJOR */, 'I', m, d, e, W, m);
    if (info != 0) { free(W); return -6; }
    for (int i = 0; i < m; ++i) {
        float acc = 0.f;
        for (int k = 0; k < m; ++k) acc += W[m * i + k] * (sqrtf(d[k]) * W[k]);   /* Brownian.cu:568-582 */
        cvec[i] = acc;
    }
    free(W);
    return 0;
}
static int tridiag_sqrt_e1(int m, const double* alpha, const double* beta /* beta[1..m-1] couple i-1,i */, double* cvec) {
    if (g_spteqr) return tridiag_sqrt_e1_lapacke(m, alpha, beta, cvec);
    double* A = (double*)calloc((size_t)m * m, sizeof(double));
    double* V = (double*)calloc((size_t)m * m, sizeof(double));
    for (int i = 0; i < m; ++i) { A[i * m + i] = alpha[i]; V[i * m + i] = 1.0; }
    for (int i = 1; i < m; ++i) { A[i * m + i - 1] = beta[i]; A[(i - 1) * m + i] = beta[i]; }
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0;
        for (int i = 0; i < m; ++i) for (int j = i + 1; j < m; ++j) off += A[i * m + j] * A[i * m + j];
        if (off < 1e-30) break;
        for (int p = 0; p < m; ++p) for (int q = p + 1; q < m; ++q) {
            double apq = A[p * m + q];
            if (fabs(apq) < 1e-300) continue;
            double th = (A[q * m + q] - A[p * m + p]) / (2 * apq);
            double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
            double cs = 1 / sqrt(t * t + 1), sn = t * cs;
            for (int k = 0; k < m; ++k) {
                double akp = A[k * m + p], akq = A[k * m + q];
                A[k * m + p] = cs * akp - sn * akq; A[k * m + q] = sn * akp + cs * akq;
            }
            for (int k = 0; k < m; ++k) {
                double apk = A[p * m + k], aqk = A[q * m + k];
                A[p * m + k] = cs * apk - sn * aqk; A[q * m + k] = sn * apk + cs * aqk;
            }
            for (int k = 0; k < m; ++k) {
                double vkp = V[k * m + p], vkq = V[k * m + q];
                V[k * m + p] = cs * vkp - sn * vkq; V[k * m + q] = sn * vkp + cs * vkq;
            }
        }
    }
    int rc = 0;
    for (int i = 0; i < m; ++i) {
        double acc = 0;
        for (int k = 0; k < m; ++k) {
            double lam = A[k * m + k];
            if (lam < 0) { if (lam < -1e-6) rc = -6; lam = 0; }
            acc += V[i * m + k] * sqrt(lam) * V[0 * m + k];
        }
        cvec[i] = acc;
    }
    free(A); free(V);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Lanczos square root — gpu_stokes_BrealLanczos_wrap, PSEv1/Brownian.cu:440-739.
 * psi4: random vector; U4 = sqrt(2T/dt) M_real^{1/2} psi.  *m is in/out (Stokes.h:157).
 * ---------------------------------------------------------------------------------------- */
static double dot3(const float* a, const float* b, int N) {
    double s = 0;
#pragma omp parallel for reduction(+ : s)
    for (int i = 0; i < N; ++i) s += (double)a[4 * i] * b[4 * i] + (double)a[4 * i + 1] * b[4 * i + 1] + (double)a[4 * i + 2] * b[4 * i + 2];
    return s;
}
static void axpby(float a, const float* A, float b, const float* B, float* C, int N) { /* Helper.cu:113-133 */
#pragma omp parallel for
    for (int i = 0; i < N; ++i) for (int d = 0; d < 3; ++d) C[4 * i + d] = a * A[4 * i + d] + b * B[4 * i + d];
}

int orc_lanczos(const orc_config* c, const orc_params* p, const float* table, const float* pos4, const float* psi4,
                const uint32_t* nn, const uint32_t* head, const uint32_t* nl, float T, float dt, int* m_inout, float* U4,
                float* stepnorm_out) {
    int N = c->N, m_max = 100, m_in = *m_inout;
    size_t vec = (size_t)4 * N;
    float* V = (float*)calloc(vec * m_max, sizeof(float));
    float *v = (float*)calloc(vec, sizeof(float)), *vj = (float*)calloc(vec, sizeof(float)), *vjm1 = (float*)calloc(vec, sizeof(float));
    float *Mvj = (float*)calloc(vec, sizeof(float)), *uo = (float*)calloc(vec, sizeof(float)), *un = (float*)calloc(vec, sizeof(float));
    double alpha[101], beta[102], cvec[101];
    memcpy(vj, psi4, vec * sizeof(float));
    float vnorm = sqrtf((float)dot3(vj, vj, N)), psinorm = vnorm; /* :444-449 */
    orc_mreal(c, p, table, pos4, psi4, nn, head, nl, Mvj);         /* :452-457 */
    float psiMpsi = (float)dot3(psi4, Mvj, N) / (psinorm * psinorm);
    axpby(1.0f / vnorm, vj, 0.f, vj, vj, N);
    int m = m_in - 1; if (m < 1) m = 1;                           /* :465-466 */
    float tempbeta = 0.f, stepnorm = 1.0f;
    int rc = 0, have_u = 0;
    for (int jj = 0; jj < m_max; ++jj) {
        if (jj >= m) { /* adaptive phase, :606 */
            if (!(stepnorm > c->error && m < m_max)) break;
            m++;
        }
        memcpy(V + vec * jj, vj, vec * sizeof(float));             /* :475 */
        beta[jj] = tempbeta;
        orc_mreal(c, p, table, pos4, vj, nn, head, nl, Mvj);       /* :481 */
        axpby(1.0f, Mvj, -tempbeta, vjm1, v, N);
        float tempalpha = (float)dot3(vj, v, N);                   /* :485-490 */
        alpha[jj] = tempalpha;
        axpby(1.0f, v, -tempalpha, vj, v, N);
        vnorm = sqrtf((float)dot3(v, v, N));                       /* :496-501 */
        tempbeta = vnorm;
        if (vnorm < 1e-8f) { m = jj; break; }                      /* :507-510 */
        axpby(1.0f / tempbeta, v, 0.f, v, v, N);
        float* t = vjm1; vjm1 = vj; vj = v; v = t;                 /* :516-520 */
        if (jj + 1 >= m) { /* solve after the initial batch and after every adaptive iteration */
            rc = tridiag_sqrt_e1(m, alpha, beta, cvec);
            if (rc) break;
#pragma omp parallel for
            for (int i = 0; i < N; ++i) for (int d = 0; d < 3; ++d) { /* Helper.cu:251-279 */
                float s = 0.f;
                for (int k = 0; k < m; ++k) s = s + V[vec * k + 4 * i + d] * (float)cvec[k];
                un[4 * i + d] = s;
            }
            if (have_u) { /* :716-726 */
                axpby(1.0f, un, -1.0f, uo, uo, N);
                stepnorm = sqrtf((float)dot3(uo, uo, N) / psiMpsi);
            }
            memcpy(uo, un, vec * sizeof(float));
            have_u = 1;
        }
    }
    float sc = psinorm * sqrtf((float)(2.0 * T / dt)); /* :739 */
    for (int i = 0; i < N; ++i) { for (int d = 0; d < 3; ++d) U4[4 * i + d] = sc * uo[4 * i + d]; U4[4 * i + 3] = 0.f; }
    *m_inout = m;
    if (stepnorm_out) *stepnorm_out = stepnorm;
    free(V); free(v); free(vj); free(vjm1); free(Mvj); free(uo); free(un);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Full velocity + Euler update — gpu_stokes_CombinedMobilityBrownian_wrap (Brownian.cu:772-923)
 * and gpu_stokes_step_one_kernel (PSEv1/Stokes.cu:137-192).  u_particles [N][3], u_grid [G][6]
 * are uniforms in [0,1) (the reference draws them from Saru; stream parity is unpinned).
 * ---------------------------------------------------------------------------------------- */
int orc_velocity(const orc_config* c, const orc_params* p, const float* table, const float* pos4, const float* F4,
                 const uint32_t* nn, const uint32_t* head, const uint32_t* nl, float T, float dt, const float* u_particles,
                 const float* u_grid, int* m_inout, float* U4) {
    int N = c->N;
    float* tmp = (float*)calloc((size_t)4 * N, sizeof(float));
    float noise_fac = sqrtf((float)(2.0 * T / dt / p->quadW)); /* Brownian.cu:198 */
    int rc = orc_mwave(c, p, pos4, F4, U4, 1, T > 0 ? u_grid : NULL, noise_fac);
    if (!rc) rc = orc_mreal(c, p, table, pos4, F4, nn, head, nl, tmp);
    for (size_t i = 0; i < (size_t)4 * N; ++i) U4[i] = tmp[i] + U4[i]; /* :882 */
    if (!rc && T > 0 && u_particles) {
        float* psi = (float*)calloc((size_t)4 * N, sizeof(float));
        const float a = 1.73205080757f; /* :121 */
        for (int i = 0; i < N; ++i) for (int d = 0; d < 3; ++d) psi[4 * i + d] = fmaf(2 * a, u_particles[3 * i + d], -a);
        rc = orc_lanczos(c, p, table, pos4, psi, nn, head, nl, T, dt, m_inout, tmp, NULL);
        for (size_t i = 0; i < (size_t)4 * N; ++i) U4[i] = tmp[i] + U4[i]; /* :917 */
        free(psi);
    }
    for (int i = 0; i < N; ++i) U4[4 * i + 3] = 0.f;
    free(tmp);
    return rc;
}

int orc_integrate(const orc_config* c, float* pos4, int* image3, const float* vel4, float dt, float shear_rate) {
    obox b = make_box(c);
    for (int i = 0; i < c->N; ++i) { /* Stokes.cu:154-190 */
        float w[3] = {pos4[4 * i], pos4[4 * i + 1], pos4[4 * i + 2]};
        float vx = vel4[4 * i] + shear_rate * w[1];
        w[0] = w[0] + vx * dt; w[1] = w[1] + vel4[4 * i + 1] * dt; w[2] = w[2] + vel4[4 * i + 2] * dt;
        wrap_pos(&b, w, image3 + 3 * i);
        pos4[4 * i] = w[0]; pos4[4 * i + 1] = w[1]; pos4[4 * i + 2] = w[2];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Dense double-precision Ewald-summed RPY mobility (the accuracy oracle):
 *   U_i = self F_i + sum_{j != i, images} [f (I - rr) + g rr] F_j
 *       + (1/V) sum_{k != 0} B(k) sinc^2(k) (I - kk) Re{ e^{i k.r_i} sum_j e^{-i k.r_j} F_j }
 * with its own splitting parameter xi_d and cutoffs (real: rc_d with explicit image loop,
 * wave: |k| <= kc_d), exact pi.  O(N^2 + N K).  Orthogonal or sheared box.
 * ---------------------------------------------------------------------------------------- */
int orc_dense_mobility(int N, double Lx, double Ly, double Lz, double xy, const double* pos3, const double* F3, double xi_d,
                       double rc_d, double kc_d, double* U3) {
    const double a = 1.0, pi = 3.14159265358979323846, V = Lx * Ly * Lz;
    double self = (1. + 4. * sqrt(pi) * a * xi_d * erfc(2. * a * xi_d) - exp(-4. * a * a * xi_d * xi_d)) / (4. * sqrt(pi) * a * xi_d * a);
    int ix = (int)ceil(rc_d / Lx) + 1, iy = (int)ceil(rc_d / Ly) + 1, iz = (int)ceil(rc_d / Lz) + 1;
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < N; ++i) {
        double u[3] = {self * F3[3 * i], self * F3[3 * i + 1], self * F3[3 * i + 2]};
        for (int j = 0; j < N; ++j)
            for (int a1 = -ix; a1 <= ix; ++a1) for (int a2 = -iy; a2 <= iy; ++a2) for (int a3 = -iz; a3 <= iz; ++a3) {
                if (j == i && a1 == 0 && a2 == 0 && a3 == 0) continue;
                double r[3] = {pos3[3 * i] - pos3[3 * j] + a1 * Lx + a2 * xy * Ly, pos3[3 * i + 1] - pos3[3 * j + 1] + a2 * Ly,
                               pos3[3 * i + 2] - pos3[3 * j + 2] + a3 * Lz};
                double d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
                if (d2 >= rc_d * rc_d) continue;
                double d = sqrt(d2), f, g;
                orc_real_fg(d, xi_d, &f, &g);
                double rdf = (r[0] * F3[3 * j] + r[1] * F3[3 * j + 1] + r[2] * F3[3 * j + 2]) / d2;
                for (int q = 0; q < 3; ++q) u[q] += f * F3[3 * j + q] + (g - f) * rdf * r[q];
            }
        U3[3 * i] = u[0]; U3[3 * i + 1] = u[1]; U3[3 * i + 2] = u[2];
    }
    int nx = (int)(kc_d * Lx / (2 * pi)) + 1, ny = (int)(kc_d * Ly / (2 * pi) * (1 + fabs(xy))) + 2, nz = (int)(kc_d * Lz / (2 * pi)) + 1;
    long nk = (long)(2 * nx + 1) * (2 * ny + 1) * (2 * nz + 1);
    double* acc = (double*)calloc((size_t)3 * N, sizeof(double));
#pragma omp parallel
    {
        double* loc = (double*)calloc((size_t)3 * N, sizeof(double));
        double* cs = (double*)malloc(sizeof(double) * 2 * N);
#pragma omp for schedule(dynamic, 64)
        for (long t = 0; t < nk; ++t) {
            int i1 = (int)(t / ((long)(2 * ny + 1) * (2 * nz + 1))) - nx, i2 = (int)((t / (2 * nz + 1)) % (2 * ny + 1)) - ny, i3 = (int)(t % (2 * nz + 1)) - nz;
            if (i1 == 0 && i2 == 0 && i3 == 0) continue;
            double k[3] = {2 * pi * i1 / Lx, 2 * pi * (i2 - xy * i1 * Ly / Lx) / Ly, 2 * pi * i3 / Lz};
            double k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
            if (k2 > kc_d * kc_d) continue;
            double kk = sqrt(k2), sinc = sin(kk * a) / (kk * a);
            double B = 6 * pi * a / k2 * (1 + k2 / (4 * xi_d * xi_d)) * exp(-k2 / (4 * xi_d * xi_d)) * sinc * sinc / V;
            double Sr[3] = {0, 0, 0}, Si[3] = {0, 0, 0};
            for (int j = 0; j < N; ++j) {
                double ph = k[0] * pos3[3 * j] + k[1] * pos3[3 * j + 1] + k[2] * pos3[3 * j + 2];
                double cc = cos(ph), ss = sin(ph);
                cs[2 * j] = cc; cs[2 * j + 1] = ss;
                for (int q = 0; q < 3; ++q) { Sr[q] += cc * F3[3 * j + q]; Si[q] -= ss * F3[3 * j + q]; }
            }
            double kSr = (k[0] * Sr[0] + k[1] * Sr[1] + k[2] * Sr[2]) / k2, kSi = (k[0] * Si[0] + k[1] * Si[1] + k[2] * Si[2]) / k2;
            double Pr[3], Pi_[3];
            for (int q = 0; q < 3; ++q) { Pr[q] = (Sr[q] - k[q] * kSr) * B; Pi_[q] = (Si[q] - k[q] * kSi) * B; }
            for (int i = 0; i < N; ++i) /* Re{ e^{i k r_i} (Pr + i Pi) } */
                for (int q = 0; q < 3; ++q) loc[3 * i + q] += cs[2 * i] * Pr[q] - cs[2 * i + 1] * Pi_[q];
        }
#pragma omp critical
        for (size_t i = 0; i < (size_t)3 * N; ++i) acc[i] += loc[i];
        free(loc); free(cs);
    }
    for (size_t i = 0; i < (size_t)3 * N; ++i) U3[i] += acc[i];
    free(acc);
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Shared-memory FFT passes of the wave-space pipeline, with the k-space scaling fused into the x pass.
//
// Reference: six 3-D C2C cuFFT calls on real input plus separate Green / random-mode kernels
// (PSEv1/Brownian.cu:844-869, PSEv1/Mobility.cu:264-299).  The first version of this engine used one batched
// R2C/C2R cuFFT pair and a scaling kernel: 3 + 1 + 3 passes over the spectrum at ~40% of HBM bandwidth
// (profiles/r1_summary.md).  Here the transform is five kernels, each one read and one write of the data:
//     z forward (R2C, two real rows per complex transform) | y forward | x forward + scaling + x inverse |
//     y inverse | z inverse (C2R)
// Every 1-D transform is an in-place mixed-radix (4, 2, 3, 5) decimation-in-frequency FFT in shared memory whose
// output stays in digit-reversed order; the inverse is the exact adjoint (decimation in time) and consumes that
// order, so no reordering pass exists anywhere.  Only the scaling needs to know which frequency sits where
// (freq_of tables).  Transforms are unnormalised, as cuFFT's (the 1/G lives in B(k), PSEv1/Helper.cu:325).
//
// Shared-memory layout of a batch of columns: element e of column c at s[e * CP + c], CP = columns + 1 (odd), so
// both "threads along e" (global loads of z rows) and "threads along c" (butterflies) are conflict-free.
#pragma once
#include "wave.cuh"

#define FFT_MAX_PASSES 12
#define FFT_MAX_N 1024
#define FFT_THREADS 256

struct Fft1D {
    int N, npass;
    unsigned long long radices;  // pass i has radix (radices >> 4 i) & 15  (a packed word keeps the plan in registers)
    const float2* tw;          // exp(-2 pi i k / N), k < N
    const uint16_t* pos_of;    // position of frequency k after the forward passes
    const uint16_t* freq_of;   // inverse map
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// y[p] = sum_q a[q] w_R^{pq}, w_R = exp(-/+ 2 pi i / R) (INV: conjugate), in place
template <int R, bool INV>
__device__ __forceinline__ void dft_small(float2 (&a)[R]) {
    if (R == 2) {
        const float2 t = a[0];
        a[0] = cadd(t, a[1]); a[1] = csub(t, a[1]);
    } else if (R == 4) {
        const float2 s02 = cadd(a[0], a[2]), d02 = csub(a[0], a[2]), s13 = cadd(a[1], a[3]), d13 = csub(a[1], a[3]);
        // forward: -i * d13 = (d13.y, -d13.x); inverse: +i * d13 = (-d13.y, d13.x)
        const float2 r = INV ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);
        a[0] = cadd(s02, s13); a[2] = csub(s02, s13); a[1] = cadd(d02, r); a[3] = csub(d02, r);
    } else if (R == 3) {
        const float c = -0.5f, sn = INV ? 0.8660254037844386f : -0.8660254037844386f;  // sin(-/+ 2 pi / 3)
        const float2 s = cadd(a[1], a[2]), d = csub(a[1], a[2]);
        const float2 m = make_float2(a[0].x + c * s.x, a[0].y + c * s.y);
        const float2 r = make_float2(-sn * d.y, sn * d.x);  // i * sn * d
        a[0] = cadd(a[0], s); a[1] = cadd(m, r); a[2] = csub(m, r);
    } else {  // R == 5
        const float c1 = 0.30901699437494745f, c2 = -0.8090169943749475f;
        const float s1 = INV ? 0.9510565162951535f : -0.9510565162951535f, s2 = INV ? 0.5877852522924731f : -0.5877852522924731f;
        const float2 p14 = cadd(a[1], a[4]), m14 = csub(a[1], a[4]), p23 = cadd(a[2], a[3]), m23 = csub(a[2], a[3]);
        const float2 t1 = make_float2(a[0].x + c1 * p14.x + c2 * p23.x, a[0].y + c1 * p14.y + c2 * p23.y);
        const float2 t2 = make_float2(a[0].x + c2 * p14.x + c1 * p23.x, a[0].y + c2 * p14.y + c1 * p23.y);
        // i * (s1 m14 + s2 m23) and i * (s2 m14 - s1 m23)
        const float2 u1 = make_float2(-(s1 * m14.y + s2 * m23.y), s1 * m14.x + s2 * m23.x);
        const float2 u2 = make_float2(-(s2 * m14.y - s1 * m23.y), s2 * m14.x - s1 * m23.x);
        a[0] = cadd(a[0], cadd(p14, p23));
        a[1] = cadd(t1, u1); a[4] = csub(t1, u1); a[2] = cadd(t2, u2); a[3] = csub(t2, u2);
    }
}

// one radix-R pass over interleaved columns (all threads of the block take part; caller syncs).
// CG = column groups per butterfly (compile-time): thread t works on butterfly t / CG and on the VEC columns
// (t % CG) + CG * u, u < VEC, which share the butterfly's index arithmetic and twiddles (a third of the instructions of a
// one-column butterfly).  CG * VEC >= ncol.
template <int R, bool INV, int CG, int VEC>
__device__ __forceinline__ void fft_pass(float2* s, int CP, int ncol, int N, int n /* sub-transform length */, const float2* stw) {
    const int m = n / R, tws = N / n;
    const float inv_m = 1.0f / (float)m;
    constexpr int GROUPS = FFT_THREADS / CG;  // threads beyond GROUPS * CG idle when CG does not divide the block
    const int col0 = threadIdx.x % CG;
    if (col0 >= ncol || threadIdx.x >= GROUPS * CG) return;
    for (int bf = threadIdx.x / CG; bf < N / R; bf += GROUPS) {
        const int b = (int)(((float)bf + 0.5f) * inv_m), j = bf - b * m;  // exact for these sizes (bf < 1024)
        float2* base = s + (b * n + j) * CP + col0;
        const int stride = m * CP;
        // twiddles w_n^{j p}, p = 1 .. R-1, from one table entry
        float2 w[R];
        w[1] = stw[j * tws];
#pragma unroll
        for (int p = 2; p < R; ++p) w[p] = cmul(w[p - 1], w[1]);
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            if (col0 + CG * u >= ncol) break;
            float2* bu = base + CG * u;
            float2 a[R];
#pragma unroll
            for (int q = 0; q < R; ++q) a[q] = bu[q * stride];
            if (INV) {  // adjoint of the forward pass: conjugate twiddles first, then the conjugate butterfly
#pragma unroll
                for (int p = 1; p < R; ++p) a[p] = cmulc(a[p], w[p]);
                dft_small<R, true>(a);
            } else {
                dft_small<R, false>(a);
#pragma unroll
                for (int p = 1; p < R; ++p) a[p] = cmul(a[p], w[p]);
            }
#pragma unroll
            for (int q = 0; q < R; ++q) bu[q * stride] = a[q];
        }
    }
}

template <bool INV, int CG, int VEC>
__device__ __forceinline__ void fft_columns(float2* s, int CP, int ncol, const Fft1D& f, const float2* stw) {
    const int N = f.N, npass = f.npass;
    const unsigned long long radices = f.radices;
    if (!INV) {
        int n = N;
        for (int i = 0; i < npass; ++i) {
            const int r = (int)((radices >> (4 * i)) & 15ull);
            if (r == 4) fft_pass<4, false, CG, VEC>(s, CP, ncol, N, n, stw);
            else if (r == 2) fft_pass<2, false, CG, VEC>(s, CP, ncol, N, n, stw);
            else if (r == 3) fft_pass<3, false, CG, VEC>(s, CP, ncol, N, n, stw);
            else fft_pass<5, false, CG, VEC>(s, CP, ncol, N, n, stw);
            n /= r;
            __syncthreads();
        }
    } else {
        int n = 1;
        for (int i = npass - 1; i >= 0; --i) {
            const int r = (int)((radices >> (4 * i)) & 15ull);
            n *= r;
            if (r == 4) fft_pass<4, true, CG, VEC>(s, CP, ncol, N, n, stw);
            else if (r == 2) fft_pass<2, true, CG, VEC>(s, CP, ncol, N, n, stw);
            else if (r == 3) fft_pass<3, true, CG, VEC>(s, CP, ncol, N, n, stw);
            else fft_pass<5, true, CG, VEC>(s, CP, ncol, N, n, stw);
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void load_twiddles(float2* stw, const Fft1D& f) {
    for (int k = threadIdx.x; k < f.N; k += blockDim.x) stw[k] = __ldg(f.tw + k);
}

// ---- z: real rows <-> half spectra, two rows per complex transform ---------------------------------------
// rows = 3 * Nx * Ny contiguous real rows of Nz (component stride = Nx*Ny*Nz, i.e. simply consecutive rows);
// spec rows of Nzp complex, frequencies kz = 0 .. Nz/2 in natural order.
#ifndef FFT_Z_COLS
#define FFT_Z_COLS 16
#endif
__global__ void __launch_bounds__(FFT_THREADS)
fft_z_forward_kernel(const float* __restrict__ grid, float2* __restrict__ spec, Fft1D f, uint32_t nrows, int Nzp) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int CP = FFT_Z_COLS + 1;
    const int N = f.N, Nzh = N / 2 + 1;
    float2* s = fsm;
    float2* stw = s + (size_t)N * CP;
    load_twiddles(stw, f);
    const uint32_t row0 = blockIdx.x * (2 * FFT_Z_COLS);
    // column c packs rows row0 + 2c (real part) and row0 + 2c + 1 (imaginary part); threads run along z
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = wid; r < 2 * FFT_Z_COLS; r += FFT_THREADS / 32) {
        const uint32_t row = row0 + r;
        const float* src = grid + (size_t)row * N;
        float* dst = reinterpret_cast<float*>(s + (r >> 1)) + (r & 1);
        for (int z = lane; z < N; z += 32) dst[2 * z * CP] = row < nrows ? __ldg(src + z) : 0.f;
    }
    __syncthreads();
    fft_columns<false, FFT_Z_COLS, 1>(s, CP, FFT_Z_COLS, f, stw);
    // X_a[k] = (Z[k] + conj Z[N-k]) / 2,  X_b[k] = (Z[k] - conj Z[N-k]) / (2i)
    for (int c = wid; c < FFT_Z_COLS; c += FFT_THREADS / 32) {
        const uint32_t ra = row0 + 2 * c, rb = ra + 1;
        if (ra >= nrows) continue;
        for (int k = lane; k < Nzh; k += 32) {
            const float2 zk = s[__ldg(f.pos_of + k) * CP + c], zm = s[__ldg(f.pos_of + (k == 0 ? 0 : N - k)) * CP + c];
            spec[(size_t)ra * Nzp + k] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            if (rb < nrows) spec[(size_t)rb * Nzp + k] = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
        }
    }
}

__global__ void __launch_bounds__(FFT_THREADS)
fft_z_inverse_kernel(const float2* __restrict__ spec, float* __restrict__ grid, Fft1D f, uint32_t nrows, int Nzp) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int CP = FFT_Z_COLS + 1;
    const int N = f.N, Nzh = N / 2 + 1;
    float2* s = fsm;
    float2* stw = s + (size_t)N * CP;
    load_twiddles(stw, f);
    const uint32_t row0 = blockIdx.x * (2 * FFT_Z_COLS);
    // Z[k] = X_a[k] + i X_b[k], Z[N-k] = conj X_a[k] + i conj X_b[k]; one thread per (column, k) writes both
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = wid; c < FFT_Z_COLS; c += FFT_THREADS / 32)
    for (int k = lane; k < Nzh; k += 32) {
        const uint32_t ra = row0 + 2 * c, rb = ra + 1;
        const float2 xa = ra < nrows ? __ldg(spec + (size_t)ra * Nzp + k) : make_float2(0.f, 0.f);
        const float2 xb = rb < nrows ? __ldg(spec + (size_t)rb * Nzp + k) : make_float2(0.f, 0.f);
        // C2R semantics of cuFFT: the imaginary parts of the k = 0 and Nyquist entries are ignored
        const bool selfc = (k == 0) || (2 * k == N);
        const float xay = selfc ? 0.f : xa.y, xby = selfc ? 0.f : xb.y;
        s[__ldg(f.pos_of + k) * CP + c] = make_float2(xa.x - xby, xay + xb.x);
        if (!selfc) s[__ldg(f.pos_of + N - k) * CP + c] = make_float2(xa.x + xby, xb.x - xay);
    }
    __syncthreads();
    fft_columns<true, FFT_Z_COLS, 1>(s, CP, FFT_Z_COLS, f, stw);
    for (int r = wid; r < 2 * FFT_Z_COLS; r += FFT_THREADS / 32) {
        const uint32_t row = row0 + r;
        if (row >= nrows) continue;
        float* dst = grid + (size_t)row * N;
        const float* src = reinterpret_cast<const float*>(s + (r >> 1)) + (r & 1);
        for (int z = lane; z < N; z += 32) dst[z] = src[2 * z * CP];
    }
}

// ---- y: in place on spec[plane][y][kz], plane = c * Nx + x -----------------------------------------------
#ifndef FFT_Y_COLS
#define FFT_Y_COLS 16
#endif
#ifndef FFT_Y_CP
#define FFT_Y_CP 17
#endif
template <bool INV>
__global__ void __launch_bounds__(FFT_THREADS)
fft_y_kernel(float2* __restrict__ spec, Fft1D f, int Nzh, int Nzp) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int CP = FFT_Y_CP;
    const int N = f.N;
    float2* s = fsm;
    float2* stw = s + (size_t)N * CP;
    load_twiddles(stw, f);
    const int k0 = blockIdx.x * FFT_Y_COLS, ncol = min(FFT_Y_COLS, Nzh - k0);
    float2* base = spec + (size_t)blockIdx.y * N * Nzp + k0;
    const int c = threadIdx.x % FFT_Y_COLS;
    if (c < ncol)
        for (int y = threadIdx.x / FFT_Y_COLS; y < N; y += FFT_THREADS / FFT_Y_COLS) s[y * CP + c] = base[(size_t)y * Nzp + c];
    __syncthreads();
    fft_columns<INV, FFT_Y_COLS, 1>(s, CP, ncol, f, stw);
    if (c < ncol)
        for (int y = threadIdx.x / FFT_Y_COLS; y < N; y += FFT_THREADS / FFT_Y_COLS) base[(size_t)y * Nzp + c] = s[y * CP + c];
}

// ---- x: forward, scaling (+ random modes), inverse in one kernel ------------------------------------------
// A block owns the 3 components of FFT_X_COLS consecutive kz at one stored y position: 3 * FFT_X_COLS columns of
// length Nx.  The y index it sits at is freq_of_y[blockIdx.y] (the y pass left digit-reversed order).
#define FFT_X_COLS 8
#define FFT_X_CP 24
__global__ void __launch_bounds__(FFT_THREADS)
fft_x_scale_kernel(float2* __restrict__ spec, Fft1D fx, const uint16_t* __restrict__ freq_of_y, WaveParams wp, PseBox box,
                   int do_det, int do_noise, const StepDev* __restrict__ sd, const float* __restrict__ u_grid,
                   int y0 = 0, int ny_local = -1 /* sharded layout: this rank holds stored y positions [y0, y0 + ny_local) */) {
    extern __shared__ __align__(16) float2 fsm[];
    constexpr int NC = 3 * FFT_X_COLS, CP = FFT_X_CP;  // row stride = 8 (mod 16) float2 for the 2 x 8 half-warp patch
    const int N = fx.N;
    float2* s = fsm;
    float2* stw = s + (size_t)N * CP;
    load_twiddles(stw, fx);
    const int k0 = blockIdx.x * FFT_X_COLS, ncol = min(FFT_X_COLS, wp.Nzh - k0);
    if (ny_local < 0) ny_local = wp.Ny;
    const int ypos = y0 + blockIdx.y;  // stored (digit-reversed) y position; blockIdx.y is the local row
    const size_t plane = (size_t)ny_local * wp.Nzp, comp = (size_t)N * plane;
    float2* base = spec + (size_t)blockIdx.y * wp.Nzp + k0;
    // column index = c * FFT_X_COLS + kzcol
    constexpr int XG = FFT_THREADS / NC;  // 10 rows of 24 columns per sweep (16 threads idle)
    const int col = threadIdx.x % NC, cc = col / FFT_X_COLS, kcol = col % FFT_X_COLS;
    const bool colthread = threadIdx.x < XG * NC;
    if (do_det) {
        if (colthread)
            for (int x = threadIdx.x / NC; x < N; x += XG)
                s[x * CP + col] = kcol < ncol ? base[cc * comp + (size_t)x * plane + kcol] : make_float2(0.f, 0.f);
        __syncthreads();
        fft_columns<false, 8, 3>(s, CP, NC, fx, stw);
    }
    {
        const uint32_t key = sd->key;
        const float noise_fac = sd->noise_fac;
        const int jj = __ldg(freq_of_y + ypos);
        for (int t = threadIdx.x; t < N * FFT_X_COLS; t += blockDim.x) {
            const int kc = t % FFT_X_COLS, xpos = t / FFT_X_COLS;
            if (kc >= ncol) continue;
            const int ii = __ldg(fx.freq_of + xpos), kk = k0 + kc;
            float2* e = s + xpos * CP + kc;
            float2 fX = make_float2(0.f, 0.f), fY = fX, fZ = fX;
            if (do_det) { fX = e[0]; fY = e[FFT_X_COLS]; fZ = e[2 * FFT_X_COLS]; }
            float2 oX, oY, oZ;
            scale_node(ii, jj, kk, fX, fY, fZ, do_det, do_noise, key, noise_fac, u_grid, wp, box, oX, oY, oZ);
            e[0] = oX; e[FFT_X_COLS] = oY; e[2 * FFT_X_COLS] = oZ;
        }
    }
    __syncthreads();
    fft_columns<true, 8, 3>(s, CP, NC, fx, stw);
    if (colthread && kcol < ncol)
        for (int x = threadIdx.x / NC; x < N; x += XG) base[cc * comp + (size_t)x * plane + kcol] = s[x * CP + col];
}

static inline size_t fft_smem_bytes(int N, int cp) { return ((size_t)N * cp + N) * sizeof(float2); }  // cp = padded row length

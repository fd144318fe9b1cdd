// Wave-space far field: Gaussian spreading, k-space scaling (+ Hermitian random modes),
// Gaussian interpolation.  FFTs are cuFFT R2C/C2R on three real grids (the reference runs six
// C2C transforms on purely real data, PSEv1/Brownian.cu:844-869).
//
// Grid layout in HBM: real grids  g[c][x][y][z], c = 0..2, z fastest (same node order as the
// reference's x*Ny*Nz + y*Nz + z, PSEv1/Mobility.cu:233); spectra  s[c][x][y][kz], kz in
// [0, Nz/2].  No k-vector table is stored (the reference rewrites a 16 B/node table every step,
// PSEv1/Stokes.cu:298): k and B(k) are recomputed per node.
#pragma once
#include "box.cuh"
#include "common.cuh"
#include "rng.cuh"
#include <cufft.h>

struct WaveParams {
    int Nx, Ny, Nz, Nzh;  // Nzh = Nz/2 + 1
    int Nzp;              // padded row length of the half spectrum in memory (>= Nzh)
    int P;
    float hx, hy, hz;
    float prefac, expfac, quadW;
    float xi, eta;
    float two_pi_k;  // 2*pi used for wave vectors (reference typo or exact)
    // Real-grid buffer addressed by the spread2 / interp2 kernels.  Single GPU: the whole grid (xorg = 0, nxa = nxw = Nx).
    // Slab-decomposed: the rank's local buffer of nxa x planes starting at global plane xorg (own planes + halo planes on
    // both sides, indexed without periodic wrap: nxw = "never").
    int xorg, nxa, nxw;
};

// ---- particle -> grid assignment (bit-exact contract) ------------------------------------------
// PSEv1/Mobility.cu:173-214: fractional position * N, truncation, centred support.
struct Support {
    int x0, y0, z0;  // first node of the support, NOT wrapped
};
__device__ __forceinline__ Support support_origin(const PseBox& box, const WaveParams& wp, float px, float py, float pz) {
    float3 f = box.make_fraction(px, py, pz);
    f.x = PSE_MUL(f.x, (float)wp.Nx);
    f.y = PSE_MUL(f.y, (float)wp.Ny);
    f.z = PSE_MUL(f.z, (float)wp.Nz);
    const int x = (int)f.x, y = (int)f.y, z = (int)f.z;
    const int odd = wp.P & 1, half = wp.P / 2;
    Support s;
    s.x0 = x - half + 1 - odd * (PSE_SUB(f.x, (float)x) < 0.5f);
    s.y0 = y - half + 1 - odd * (PSE_SUB(f.y, (float)y) < 0.5f);
    s.z0 = z - half + 1 - odd * (PSE_SUB(f.z, (float)z) < 0.5f);
    return s;
}
__device__ __forceinline__ int wrap_node(int i, int n) { return i < 0 ? i + n : (i > n - 1 ? i - n : i); }

__global__ void grid_index_kernel(const float4* __restrict__ pos, uint32_t N, PseBox box, WaveParams wp,
                                  int3* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 p = __ldg(pos + i);
    Support s = support_origin(box, wp, p.x, p.y, p.z);
    out[i] = make_int3(wrap_node(s.x0, wp.Nx), wrap_node(s.y0, wp.Ny), wrap_node(s.z0, wp.Nz));
}

// Gaussian weight of node (ix,iy,iz) for a particle, as PSEv1/Mobility.cu:222-241
__device__ __forceinline__ float gauss_weight(const PseBox& box, const WaveParams& wp, int ix, int iy, int iz, float px,
                                              float py, float pz, float pref) {
    float gx = wp.hx * (float)ix - box.Lx * 0.5f;
    float gy = wp.hy * (float)iy - box.Ly * 0.5f;
    float gz = wp.hz * (float)iz - box.Lz * 0.5f;
    gx = gx + box.xy * gy;
    float3 r = box.min_image(make_float3(gx - px, gy - py, gz - pz));
    float rsq = r.x * r.x + r.y * r.y + r.z * r.z;
    return pref * expf(-wp.expfac * rsq);
}

// ---- spreading, scatter form ------------------------------------------------------------------
// One warp per particle; lanes tile the (x,y) footprint and walk z.  REDG float atomics into the
// three real grids (which must be zero on entry).
__global__ void __launch_bounds__(256)
spread_scatter_kernel(const float4* __restrict__ pos, const float4* __restrict__ F, uint32_t N, PseBox box, WaveParams wp,
                      float* __restrict__ grid) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= N) return;
    const float4 pp = __ldg(pos + p), ff = __ldg(F + p);
    const Support s = support_origin(box, wp, pp.x, pp.y, pp.z);
    const size_t G = (size_t)wp.Nx * wp.Ny * wp.Nz;
    const int P = wp.P, PP = P * P;
    for (int t = lane; t < PP; t += 32) {
        const int tx = t / P, ty = t - tx * P;
        const int ix = wrap_node(s.x0 + tx, wp.Nx), iy = wrap_node(s.y0 + ty, wp.Ny);
        for (int tz = 0; tz < P; ++tz) {
            const int iz = wrap_node(s.z0 + tz, wp.Nz);
            const float w = gauss_weight(box, wp, ix, iy, iz, pp.x, pp.y, pp.z, wp.prefac);
            const size_t idx = ((size_t)ix * wp.Ny + iy) * wp.Nz + iz;
            atomicAdd(grid + idx, w * ff.x);
            atomicAdd(grid + G + idx, w * ff.y);
            atomicAdd(grid + 2 * G + idx, w * ff.z);
        }
    }
}

// ---- interpolation, one warp per particle -----------------------------------------------------
// U[perm[slot]] (+)= quadW*prefac * sum_nodes w * grid    (PSEv1/Mobility.cu:325-477)
__global__ void __launch_bounds__(256)
interp_warp_kernel(const float4* __restrict__ pos, uint32_t N, PseBox box, WaveParams wp, const float* __restrict__ grid,
                   const uint32_t* __restrict__ perm, float4* __restrict__ U, int accumulate) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= N) return;
    const float4 pp = __ldg(pos + p);
    const Support s = support_origin(box, wp, pp.x, pp.y, pp.z);
    const size_t G = (size_t)wp.Nx * wp.Ny * wp.Nz;
    const int P = wp.P, PP = P * P;
    const float pref = wp.quadW * wp.prefac;
    float3 acc = make_float3(0.f, 0.f, 0.f);
    for (int t = lane; t < PP; t += 32) {
        const int tx = t / P, ty = t - tx * P;
        const int ix = wrap_node(s.x0 + tx, wp.Nx), iy = wrap_node(s.y0 + ty, wp.Ny);
        for (int tz = 0; tz < P; ++tz) {
            const int iz = wrap_node(s.z0 + tz, wp.Nz);
            const float w = gauss_weight(box, wp, ix, iy, iz, pp.x, pp.y, pp.z, pref);
            const size_t idx = ((size_t)ix * wp.Ny + iy) * wp.Nz + iz;
            acc.x += w * __ldg(grid + idx);
            acc.y += w * __ldg(grid + G + idx);
            acc.z += w * __ldg(grid + 2 * G + idx);
        }
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z);
    if (lane == 0) {
        const uint32_t id = perm ? perm[p] : p;
        float4 o = U[id];  // .w preserved (PSEv1/Mobility.cu:474)
        if (!accumulate) { o.x = 0.f; o.y = 0.f; o.z = 0.f; }
        o.x += acc.x; o.y += acc.y; o.z += acc.z;
        U[id] = o;
    }
}

// ---- k-space scaling + random modes, half spectrum ---------------------------------------------
// Per-step scalars kept in device memory so that one captured CUDA graph of the step can be replayed with a new
// time step / temperature (kernel parameters are frozen at capture time).
struct StepDev {
    uint32_t key;     // timestep + hashed seed (RNG stream key)
    float noise_fac;  // sqrt(2 T / dt / quadW)
};

struct KVec { float kx, ky, kz, w; };  // w = B(k)/G without the sinc^2 factor, as gridk.w

// wave vector and scaling of full-grid node (i,j,k): PSEv1/Helper.cu:300-329.  The reference's chains of IEEE divisions
// (x / Lx, k2 / 4 / xi^2, ... / k2 / G) are folded into reciprocal multiplications and one division; the values move
// by an ulp, far below the 1e-5 parity budget, and the scaling pass loses ~100 instructions per node.
__device__ __forceinline__ KVec k_of_node(int i, int j, int k, const WaveParams& wp, const PseBox& box) {
    KVec kv;
    const float fx0 = (float)((i < (wp.Nx + 1) / 2) ? i : i - wp.Nx);
    const float fy = ((float)((j < (wp.Ny + 1) / 2) ? j : j - wp.Ny) - box.xy * fx0 * box.Ly * box.Lxinv) * box.Lyinv;
    const float fx = fx0 * box.Lxinv;
    const float fz = (float)((k < (wp.Nz + 1) / 2) ? k : k - wp.Nz) * box.Lzinv;
    kv.kx = fx * wp.two_pi_k; kv.ky = fy * wp.two_pi_k; kv.kz = fz * wp.two_pi_k;
    const float k2 = kv.kx * kv.kx + kv.ky * kv.ky + kv.kz * kv.kz;
    const float q = k2 * (0.25f / (wp.xi * wp.xi));  // k^2 / (4 xi^2)
    const float G = (float)(wp.Nx * wp.Ny * wp.Nz);
    kv.w = (i == 0 && j == 0 && k == 0) ? 0.f : 6.0f * 3.1415926536f * (1.0f + q) * expf(-(1.f - wp.eta) * q) / (k2 * G);
    return kv;
}

// does the reference's selection rule process full-grid node (ii,jj,kk)?  PSEv1/Brownian.cu:210-215
__device__ __forceinline__ bool ref_processed(int ii, int jj, int kk, const WaveParams& wp) {
    return !(2 * kk >= wp.Nz + 1) && !((kk == 0) && (2 * jj >= wp.Ny + 1)) &&
           !((kk == 0) && (jj == 0) && (2 * ii >= wp.Nx + 1)) && !((kk == 0) && (jj == 0) && (ii == 0));
}

// six uniforms of a node (reX,reY,reZ,imX,imY,imZ) on (-sqrt(3/2), sqrt(3/2)): PSEv1/Brownian.cu:179-189
__device__ __forceinline__ void node_draws(uint32_t idx, const float* __restrict__ u_grid, uint32_t key, float re[3],
                                           float im[3]) {
    const float a = 1.2247448713915889f;
    if (u_grid) {
        const float* t = u_grid + (size_t)idx * 6;
        re[0] = pse_affine(__ldg(t + 0), -a, a); re[1] = pse_affine(__ldg(t + 1), -a, a); re[2] = pse_affine(__ldg(t + 2), -a, a);
        im[0] = pse_affine(__ldg(t + 3), -a, a); im[1] = pse_affine(__ldg(t + 4), -a, a); im[2] = pse_affine(__ldg(t + 5), -a, a);
    } else {
        const uint4 b0 = pse_philox(idx, 0u, PSE_RNG_DOMAIN_GRID, key), b1 = pse_philox(idx, 1u, PSE_RNG_DOMAIN_GRID, key);
        re[0] = pse_uniform(b0.x, -a, a); re[1] = pse_uniform(b0.y, -a, a); re[2] = pse_uniform(b0.z, -a, a);
        im[0] = pse_uniform(b0.w, -a, a); im[1] = pse_uniform(b1.x, -a, a); im[2] = pse_uniform(b1.y, -a, a);
    }
}

// scaled value of one node for wave vector kv:  out += scale * (f - k (k.f)/k^2)   (complex 3-vector)
__device__ __forceinline__ void project_add(const KVec& kv, float inv_ksq, float scale, const float2 fX, const float2 fY, const float2 fZ,
                                            float2& oX, float2& oY, float2& oZ) {
    const float2 kdF = make_float2((kv.kx * fX.x + kv.ky * fY.x + kv.kz * fZ.x) * inv_ksq,
                                   (kv.kx * fX.y + kv.ky * fY.y + kv.kz * fZ.y) * inv_ksq);
    oX.x += (fX.x - kv.kx * kdF.x) * scale; oX.y += (fX.y - kv.kx * kdF.y) * scale;
    oY.x += (fY.x - kv.ky * kdF.x) * scale; oY.y += (fY.y - kv.ky * kdF.y) * scale;
    oZ.x += (fZ.x - kv.kz * kdF.x) * scale; oZ.y += (fZ.y - kv.kz * kdF.y) * scale;
}
// sin(|k|)/|k| and 1/|k|^2 of a wave vector
__device__ __forceinline__ void sinc_and_inv_ksq(const KVec& kv, float& sinc, float& inv_ksq) {
    const float ksq = kv.kx * kv.kx + kv.ky * kv.ky + kv.kz * kv.kz;
    const float inv_k = rsqrtf(ksq);
    inv_ksq = inv_k * inv_k;
    sinc = sinf(ksq * inv_k) * inv_k;
}

// One thread per half-spectrum node.  deterministic: u = B (I - kk/k^2) f  (PSEv1/Mobility.cu:264-299);
// stochastic: + fac * B^{1/2} (I - kk/k^2) (re + i im), Hermitian by construction
// (PSEv1/Brownian.cu:153-345, each conjugate pair generated exactly once — SURVEY.md Q4).
// do_det = 0 discards the incoming spectrum (pure noise field).
//
// Nyquist components (even sizes): the reference scales the full C2C spectrum node by node and keeps
// the real part of the inverse transform, i.e. the Hermitian part (X(n) + conj X(mirror))/2 of its
// spectrum.  A node n with a Nyquist index and its mirror carry wave vectors that are NOT negatives of
// each other (PSEv1/Helper.cu:308-311 maps index N/2 to -N/2 for both; under shear even |k| differs), so
// the half-spectrum C2R path evaluates both wave vectors and averages.
// Launch: grid (Ny, Nx), one block per (ii, jj) row of the half spectrum, threads stride over kz (coalesced,
// no per-thread integer division).
__device__ __forceinline__ void scale_node(int ii, int jj, int kk, const float2 fX, const float2 fY, const float2 fZ, int do_det,
                                           int do_noise, uint32_t key, float noise_fac, const float* __restrict__ u_grid,
                                           const WaveParams& wp, const PseBox& box, float2& oX, float2& oY, float2& oZ) {
    oX = make_float2(0.f, 0.f); oY = oX; oZ = oX;
    if (ii == 0 && jj == 0 && kk == 0) return;
    const bool ii_nyq = (ii == wp.Nx / 2) && (wp.Nx / 2 == (wp.Nx + 1) / 2);
    const bool jj_nyq = (jj == wp.Ny / 2) && (wp.Ny / 2 == (wp.Ny + 1) / 2);
    const bool kk_nyq = (kk == wp.Nz / 2) && (wp.Nz / 2 == (wp.Nz + 1) / 2);
    const int mi = ii == 0 ? 0 : wp.Nx - ii, mj = jj == 0 ? 0 : wp.Ny - jj, mk = kk == 0 ? 0 : wp.Nz - kk;
    const uint32_t idx = ((uint32_t)ii * wp.Ny + jj) * wp.Nz + kk;  // full-grid node index (reference numbering)
    const uint32_t midx = ((uint32_t)mi * wp.Ny + mj) * wp.Nz + mk;
    const bool self_conj = idx == midx;
    const bool two = (ii_nyq || jj_nyq || kk_nyq) && !self_conj;  // mirror wave vector differs from -k
    const KVec kv = k_of_node(ii, jj, kk, wp, box);
    float sinc, iksq;
    sinc_and_inv_ksq(kv, sinc, iksq);
    KVec kvm = kv;
    float sincm = sinc, iksqm = iksq;
    if (two) { kvm = k_of_node(mi, mj, mk, wp, box); sinc_and_inv_ksq(kvm, sincm, iksqm); }
    const float half = two ? 0.5f : 1.0f;
    if (do_det) {
        project_add(kv, iksq, half * kv.w * sinc * sinc, fX, fY, fZ, oX, oY, oZ);
        if (two) project_add(kvm, iksqm, half * kvm.w * sincm * sincm, fX, fY, fZ, oX, oY, oZ);
    }
    if (do_noise) {
        float re[3], im[3];
        if (self_conj) {
            node_draws(idx, u_grid, key, re, im);
            const float sqrt2 = 1.4142135623730951f;
            re[0] *= sqrt2; re[1] *= sqrt2; re[2] *= sqrt2;
            im[0] = im[1] = im[2] = 0.f;
        } else {
            // the mirror node shares this node's conjugate pair; it lives in the half spectrum only on
            // the kz = 0 and kz = Nyquist planes.  Owner of the pair = the node the reference rule
            // processes; if the rule processes both, the smaller index owns it.
            const bool me = ref_processed(ii, jj, kk, wp), other = ref_processed(mi, mj, mk, wp);
            const bool own = me && (!other || idx < midx);
            node_draws(own ? idx : midx, u_grid, key, re, im);
            if (!own) { im[0] = -im[0]; im[1] = -im[1]; im[2] = -im[2]; }
        }
        const float2 dX = make_float2(re[0], im[0]), dY = make_float2(re[1], im[1]), dZ = make_float2(re[2], im[2]);
        project_add(kv, iksq, half * noise_fac * sqrtf(kv.w) * sinc, dX, dY, dZ, oX, oY, oZ);
        if (two) project_add(kvm, iksqm, half * noise_fac * sqrtf(kvm.w) * sincm, dX, dY, dZ, oX, oY, oZ);
    }
}
__global__ void __launch_bounds__(128)
scale_kernel(float2* __restrict__ spec, WaveParams wp, PseBox box, int do_det, int do_noise, const StepDev* __restrict__ sd,
             const float* __restrict__ u_grid, int y0 = 0, int ny_local = -1) {
    // sharded layout: this rank holds y rows [y0, y0 + ny_local) of every x plane, spec[c][x][y_local][kz]
    if (ny_local < 0) ny_local = wp.Ny;
    const uint32_t key = sd->key;
    const float noise_fac = sd->noise_fac;
    const size_t nh = (size_t)wp.Nx * ny_local * wp.Nzp;
    const int jj = blockIdx.x + y0, ii = blockIdx.y;
    const size_t rowbase = ((size_t)ii * ny_local + blockIdx.x) * wp.Nzp;
    for (int kk = threadIdx.x; kk < wp.Nzh; kk += blockDim.x) {
        const size_t tid = rowbase + kk;
        float2 fX = make_float2(0.f, 0.f), fY = fX, fZ = fX;
        if (do_det) { fX = spec[tid]; fY = spec[nh + tid]; fZ = spec[2 * nh + tid]; }
        float2 oX, oY, oZ;
        scale_node(ii, jj, kk, fX, fY, fZ, do_det, do_noise, key, noise_fac, u_grid, wp, box, oX, oY, oZ);
        spec[tid] = oX; spec[nh + tid] = oY; spec[2 * nh + tid] = oZ;
    }
}

// particle noise psi (slot order) : 3 uniforms on (-sqrt3, sqrt3), PSEv1/Brownian.cu:99-130
__global__ void psi_kernel(float4* __restrict__ psi /* element s at psi[s * stride] */, int stride, const uint32_t* __restrict__ perm,
                           uint32_t N, const float* __restrict__ u_particles, const StepDev* __restrict__ sd) {
    const uint32_t key = sd->key;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const uint32_t id = perm[s];
    const float a = 1.73205080757f;
    float x, y, z;
    if (u_particles) {
        x = pse_affine(__ldg(u_particles + 3 * (size_t)id + 0), -a, a);
        y = pse_affine(__ldg(u_particles + 3 * (size_t)id + 1), -a, a);
        z = pse_affine(__ldg(u_particles + 3 * (size_t)id + 2), -a, a);
    } else {
        const uint4 b = pse_philox(id, 0u, PSE_RNG_DOMAIN_PARTICLE, key);
        x = pse_uniform(b.x, -a, a); y = pse_uniform(b.y, -a, a); z = pse_uniform(b.z, -a, a);
    }
    psi[(size_t)s * stride] = make_float4(x, y, z, 0.f);
}

// Real-space RPY near field: matrix-free SpMV over the CSR neighbour list, and the fused
// Lanczos kernels built on it.
//
// Reference: gpu_stokes_Mreal_kernel (PSEv1/Mobility.cu:594-687) is one thread per particle with
// a serial neighbour loop; the Lanczos driver (PSEv1/Brownian.cu:357-765) wraps it in 7-10 small
// launches and 2-3 blocking D2H copies per iteration.  Here:
//   * particles are in cell order, TPP lanes share one row (coalesced index reads, gathered
//     float4 loads that mostly hit L1/L2), partial sums meet in a shuffle reduction;
//   * one Lanczos iteration is two kernels (spmv_kernel<LANCZOS> + lanczos_update_kernel) with
//     alpha/beta kept in device memory, written by a deterministic last-block reduction.
#pragma once
#include "box.cuh"
#include "common.cuh"

struct RealParams {
    float self;     // M_real self term (PSEv1/Stokes.cc:319)
    float rcut_sq;  // ewald_cut^2
    float dr;       // table spacing
    float dr_sq;
    float rcut;
    int ewald_n;
};

// pair kernel, operation for operation as PSEv1/Mobility.cu:646-678
__device__ __forceinline__ void rpy_pair(const float3 r, const float r2, const float4 Fj, const float4* __restrict__ table,
                                         const RealParams& rp, float3& u) {
    float dist = sqrtf(r2);
    int r_ind = __float2int_rd((float)rp.ewald_n * (dist - rp.dr) / (rp.rcut - rp.dr));
    float4 t = __ldg(table + r_ind);
    float fac = dist / rp.dr - (float)r_ind - 1.0f;
    float Imrr = t.x + (t.z - t.x) * fac;
    float rr = t.y + (t.w - t.y) * fac;
    float rdotf = (r.x * Fj.x + r.y * Fj.y + r.z * Fj.z) / r2;
    float c = (rr - Imrr) * rdotf;
    u.x += Imrr * Fj.x + c * r.x;
    u.y += Imrr * Fj.y + c * r.y;
    u.z += Imrr * Fj.z + c * r.z;
}

enum { SPMV_PLAIN = 0, SPMV_LANCZOS = 1 };

struct LanczosArgs {
    const float* beta_j;   // beta_j (norm of the unnormalised input), device scalar
    const float4* v_prev;  // v_{j-1} (ignored when j == 0)
    float4* v_out;         // V[j] = normalised input
    float* alpha_out;      // alpha_j
    float* partials;
    unsigned int* counter;
    int first;             // j == 0
};

// y = M x            (PLAIN:   x = F, y = U)
// LANCZOS: x = u_j (unnormalised), s = 1/beta_j;  v_j = s x -> V[j];
//          y = s (M x) - beta_j v_{j-1};  alpha_j = v_j . y     (PSEv1/Brownian.cu:481-490)
template <int TPP, int MODE>
__global__ void __launch_bounds__(256)
spmv_kernel(const float4* __restrict__ spos, const float4* __restrict__ x, float4* __restrict__ y, uint32_t N,
            const uint32_t* __restrict__ nn, const uint32_t* __restrict__ head, const uint32_t* __restrict__ nl,
            const float4* __restrict__ table, RealParams rp, PseBox box, LanczosArgs la) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) / TPP;
    const int sub = threadIdx.x % TPP;
    float3 u = make_float3(0.f, 0.f, 0.f);
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), xi = pi;
    const bool live = row < N;
    if (live) {
        pi = __ldg(spos + row);
        xi = __ldg(x + row);
        const uint32_t n = __ldg(nn + row), h = __ldg(head + row);
        for (uint32_t k = sub; k < n; k += TPP) {
            const uint32_t j = __ldg(nl + h + k);
            const float4 pj = __ldg(spos + j);
            float3 r = box.min_image(make_float3(PSE_SUB(pi.x, pj.x), PSE_SUB(pi.y, pj.y), PSE_SUB(pi.z, pj.z)));
            const float r2 = r.x * r.x + r.y * r.y + r.z * r.z;
            if (r2 < rp.rcut_sq && r2 >= rp.dr_sq) {
                const float4 xj = __ldg(x + j);
                rpy_pair(r, r2, xj, table, rp, u);
            }
        }
    }
    u.x = group_sum<TPP>(u.x);
    u.y = group_sum<TPP>(u.y);
    u.z = group_sum<TPP>(u.z);

    if (MODE == SPMV_PLAIN) {
        if (live && sub == 0)
            y[row] = make_float4(u.x + rp.self * xi.x, u.y + rp.self * xi.y, u.z + rp.self * xi.z, 0.f);
    } else {
        __shared__ float red[32];
        float part = 0.f;
        if (live && sub == 0) {
            const float beta = __ldcg(la.beta_j);
            const float s = beta > 1e-8f ? 1.0f / beta : 0.f;  // breakdown guard, PSEv1/Brownian.cu:507-510
            float3 v = make_float3(s * xi.x, s * xi.y, s * xi.z);
            float3 mv = make_float3(s * (u.x + rp.self * xi.x), s * (u.y + rp.self * xi.y), s * (u.z + rp.self * xi.z));
            if (!la.first) {
                const float4 vp = __ldg(la.v_prev + row);
                mv.x -= beta * vp.x; mv.y -= beta * vp.y; mv.z -= beta * vp.z;
            }
            la.v_out[row] = make_float4(v.x, v.y, v.z, 0.f);
            y[row] = make_float4(mv.x, mv.y, mv.z, 0.f);
            part = v.x * mv.x + v.y * mv.y + v.z * mv.z;
        }
        float tot = block_sum(part, red);
        grid_sum_finish(tot, la.partials, la.counter, la.alpha_out, red);
    }
}

// w = y - alpha_j v_j ;  beta_{j+1} = ||w|| ; u_{j+1} = w (left unnormalised; the next
// spmv_kernel<LANCZOS> folds 1/beta_{j+1} in).   PSEv1/Brownian.cu:493-514
__global__ void __launch_bounds__(256)
lanczos_update_kernel(const float4* __restrict__ y, const float4* __restrict__ vj, float4* __restrict__ u_next, uint32_t N,
                      const float* __restrict__ alpha_j, float* __restrict__ beta_next, float* partials,
                      unsigned int* counter) {
    __shared__ float red[32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (i < N) {
        const float a = __ldcg(alpha_j);
        const float4 yy = __ldg(y + i), v = __ldg(vj + i);
        float3 w = make_float3(yy.x - a * v.x, yy.y - a * v.y, yy.z - a * v.z);
        u_next[i] = make_float4(w.x, w.y, w.z, 0.f);
        part = w.x * w.x + w.y * w.y + w.z * w.z;
    }
    float tot = block_sum(part, red);
    grid_sum_finish(tot, partials, counter, beta_next, red, /*take_sqrt=*/true);
}

// dot(a, b) over xyz -> *out (deterministic); take_sqrt gives the norm when a == b
__global__ void __launch_bounds__(256)
dot_kernel(const float4* __restrict__ a, const float4* __restrict__ b, uint32_t N, float* out, float* partials,
           unsigned int* counter, bool take_sqrt) {
    __shared__ float red[32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (i < N) {
        const float4 p = __ldg(a + i), q = __ldg(b + i);
        part = p.x * q.x + p.y * q.y + p.z * q.z;
    }
    float tot = block_sum(part, red);
    grid_sum_finish(tot, partials, counter, out, red, take_sqrt);
}

// out[perm[slot]] (+)= scale * sum_k c[k] V[k][slot]   (PSEv1/Helper.cu:251-279 + the final rescale
// PSEv1/Brownian.cu:739), result scattered back to particle-id order and accumulated into U.
__global__ void __launch_bounds__(256)
basis_combine_kernel(const float4* __restrict__ V, const float* __restrict__ c, int m, uint32_t N, size_t stride,
                     const float* __restrict__ psinorm, float thermal, const uint32_t* __restrict__ perm,
                     float4* __restrict__ U, int accumulate) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    float3 acc = make_float3(0.f, 0.f, 0.f);
    for (int k = 0; k < m; ++k) {
        const float4 v = __ldg(V + (size_t)k * stride + s);
        const float ck = __ldg(c + k);
        acc.x += v.x * ck; acc.y += v.y * ck; acc.z += v.z * ck;
    }
    const float sc = __ldcg(psinorm) * thermal;
    const uint32_t p = perm[s];
    float4 o = accumulate ? U[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x += sc * acc.x; o.y += sc * acc.y; o.z += sc * acc.z; o.w = 0.f;
    U[p] = o;
}

// U[perm[slot]] (+)= y[slot]
__global__ void scatter_add_kernel(const float4* __restrict__ y, const uint32_t* __restrict__ perm, uint32_t N,
                                   float4* __restrict__ U, int accumulate) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const float4 v = __ldg(y + s);
    const uint32_t p = perm[s];
    float4 o = accumulate ? U[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x += v.x; o.y += v.y; o.z += v.z; o.w = 0.f;
    U[p] = o;
}

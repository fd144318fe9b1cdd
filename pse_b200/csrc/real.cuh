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
#include <type_traits>

struct RealParams {
    float self;     // M_real self term (PSEv1/Stokes.cc:319)
    float rcut_sq;  // ewald_cut^2
    float dr;       // table spacing
    float dr_sq;
    float rcut;
    int ewald_n;
    float inv_dr;     // 1 / dr
    float tab_scale;  // ewald_n / (rcut - dr)
};

// Pair kernel: PSEv1/Mobility.cu:646-678 with the three IEEE divisions replaced by one rsqrt and
// precomputed reciprocals (table index (dist-dr)*ewald_n/(rcut-dr), fac = dist/dr - ind - 1, (r.F)/r^2).
// The table is piecewise linear and continuous, so an index that lands one entry off at a knot gives the
// same value to round-off; measured parity against the reference kernel stays at 1e-7.
// TAB is either the reference-layout global table (float4: f_k, g_k, f_k+1, g_k+1) or its shared-memory copy
// (float2 per knot: entry k and k+1 are read separately).
struct TableGlobal {
    const float4* t;
    __device__ __forceinline__ void fg(float dist, float, const RealParams& rp, float& Imrr, float& rr) const {
        const int r_ind = __float2int_rd((dist - rp.dr) * rp.tab_scale);
        const float4 e = __ldg(t + r_ind);
        const float fac = dist * rp.inv_dr - (float)r_ind - 1.0f;
        Imrr = e.x + (e.z - e.x) * fac;
        rr = e.y + (e.w - e.y) * fac;
    }
};
struct TableShared {
    const float2* t;
    __device__ __forceinline__ void fg(float dist, float, const RealParams& rp, float& Imrr, float& rr) const {
        const int r_ind = __float2int_rd((dist - rp.dr) * rp.tab_scale);
        const float2 a = t[r_ind], b = t[r_ind + 1];
        const float fac = dist * rp.inv_dr - (float)r_ind - 1.0f;
        Imrr = a.x + (b.x - a.x) * fac;
        rr = a.y + (b.y - a.y) * fac;
    }
};
// Non-overlapping pairs (r >= 2a, every pair of a physical suspension) take no table at all:
//     f = exp(-xi^2 (r-2a)^2) Pf(t), g = exp(-xi^2 (r-2a)^2) Pg(t), t = A/r + B, degree PSE_CHEB_DEG
// (pse_fit_rpy_cheb in params.cpp; agreement with the closed forms ~1e-6 of max|f|, i.e. the same size as the
// reference table's own linear-interpolation error).  The coefficients arrive as kernel parameters, so they are
// constant-bank operands of the FMAs: the lookup leaves the L1/shared data pipe, which is what bounds this kernel
// (profiles/r1_summary.md).  Overlapping pairs (r < 2a) fall back to the global table.
#define PSE_CHEB_DEG 8
struct ChebCoef {
    float A, B, ne;  // t = A / r + B;  ne = -xi^2 log2(e)
    float cf[PSE_CHEB_DEG + 1], cg[PSE_CHEB_DEG + 1];
};
struct TableCheb {
    ChebCoef c;
    const float4* t;
    __device__ __forceinline__ void fg(float dist, float inv_dist, const RealParams& rp, float& Imrr, float& rr) const {
        if (dist >= 2.0f) {
            const float tt = fmaf(c.A, inv_dist, c.B);
            float pf = c.cf[PSE_CHEB_DEG], pg = c.cg[PSE_CHEB_DEG];
#pragma unroll
            for (int k = PSE_CHEB_DEG - 1; k >= 0; --k) { pf = fmaf(pf, tt, c.cf[k]); pg = fmaf(pg, tt, c.cg[k]); }
            const float d = dist - 2.0f;
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(c.ne * d * d));
            Imrr = pf * e;
            rr = pg * e;
        } else {
            const int r_ind = __float2int_rd((dist - rp.dr) * rp.tab_scale);
            const float4 e = __ldg(t + r_ind);
            const float fac = dist * rp.inv_dr - (float)r_ind - 1.0f;
            Imrr = e.x + (e.z - e.x) * fac;
            rr = e.y + (e.w - e.y) * fac;
        }
    }
};
template <class TAB>
__device__ __forceinline__ void rpy_pair(const float3 r, const float r2, const float4 Fj, const TAB& table,
                                         const RealParams& rp, float3& u) {
    float inv_dist;  // r2 >= dr^2 = 1e-6: no denormal handling needed (rsqrtf() spends 4 extra instructions on it)
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_dist) : "f"(r2));
    const float dist = r2 * inv_dist;
    float Imrr, rr;
    table.fg(dist, inv_dist, rp, Imrr, rr);
    const float rdotf = (r.x * Fj.x + r.y * Fj.y + r.z * Fj.z) * (inv_dist * inv_dist);
    const float c = (rr - Imrr) * rdotf;
    u.x = fmaf(c, r.x, fmaf(Imrr, Fj.x, u.x));
    u.y = fmaf(c, r.y, fmaf(Imrr, Fj.y, u.y));
    u.z = fmaf(c, r.z, fmaf(Imrr, Fj.z, u.z));
}

// the same pair applied to two vectors at once (M_real F rides along with the first Lanczos product)
template <class TAB>
__device__ __forceinline__ void rpy_pair2(const float3 r, const float r2, const float4 Fa, const float4 Fb, const TAB& table,
                                          const RealParams& rp, float3& ua, float3& ub) {
    float inv_dist;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_dist) : "f"(r2));
    const float dist = r2 * inv_dist;
    float Imrr, rr;
    table.fg(dist, inv_dist, rp, Imrr, rr);
    const float i2 = inv_dist * inv_dist, dg = rr - Imrr;
    const float ca = dg * ((r.x * Fa.x + r.y * Fa.y + r.z * Fa.z) * i2), cb = dg * ((r.x * Fb.x + r.y * Fb.y + r.z * Fb.z) * i2);
    ua.x = fmaf(ca, r.x, fmaf(Imrr, Fa.x, ua.x)); ua.y = fmaf(ca, r.y, fmaf(Imrr, Fa.y, ua.y)); ua.z = fmaf(ca, r.z, fmaf(Imrr, Fa.z, ua.z));
    ub.x = fmaf(cb, r.x, fmaf(Imrr, Fb.x, ub.x)); ub.y = fmaf(cb, r.y, fmaf(Imrr, Fb.y, ub.y)); ub.z = fmaf(cb, r.z, fmaf(Imrr, Fb.z, ub.z));
}

// Slot-ordered particle record of the real-space kernels: position and the vector being multiplied share
// one 32-byte sector, so a neighbour gather costs one sector instead of two.
struct __align__(32) PX {
    float4 p;  // position (x, y, z, -)
    float4 x;  // input vector of the SpMV: force, psi, or the unnormalised Lanczos vector u_j
};

// one 256-bit load of a whole record (LDG.E.256 on sm_100a): half the L1 wavefronts of two 128-bit loads
__device__ __forceinline__ void ld_px(const PX* ptr, float4& p, float4& x) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w), "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                 : "l"(ptr));
}

// Pruned lists carry, in the top bit of each entry, whether the minimum image of the pair differs from the plain
// difference of the two positions; the SpMV then skips the image arithmetic for all other pairs (BoxDim::minImage
// returns its argument unchanged, bit for bit, when every image index rounds to zero).  Needs N < 2^31.
#define PSE_WRAP_BIT 0x80000000u
// Per-step pruning of the buffered neighbour list: keeps, in row order, the neighbours that are inside the
// real-space cutoff at the CURRENT positions (dr^2 <= r^2 < rcut^2, the filter of PSEv1/Mobility.cu:652).  The
// m+1 SpMVs of a step then gather only pairs that contribute (the SpMV is bound by the gathered records, so
// its cost is proportional to the listed pairs).  8 lanes per row, ballot-ordered compaction, same row offsets.
__global__ void __launch_bounds__(256)
prune_kernel(const float4* __restrict__ spos, uint32_t N, const uint32_t* __restrict__ nn, const uint32_t* __restrict__ head,
             const uint32_t* __restrict__ nl, RealParams rp, PseBox box, uint32_t* __restrict__ nn_act, uint32_t* __restrict__ nl_act,
             uint32_t row_begin = 0, uint32_t ell_stride = 0 /* != 0: nl holds fixed-stride rows relative to row_begin (search output) */) {
    const int sub = threadIdx.x & 7;
    const int grp = (threadIdx.x & 31) >> 3;  // group inside the warp
    for (uint32_t row0 = row_begin + blockIdx.x * 32; row0 < N; row0 += gridDim.x * 32) {
        const uint32_t row = row0 + (threadIdx.x >> 3);
        const bool live = row < N;
        uint32_t n = 0, h = 0;
        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t* __restrict__ src = nl;
        if (live) {
            n = __ldg(nn + row); h = __ldg(head + row); pi = __ldg(spos + row);
            src = ell_stride ? nl + (size_t)(row - row_begin) * ell_stride : nl + h;
        }
        // all 32 lanes iterate together (ballot needs convergence): trip count = longest row in the warp
        uint32_t nmax = n;
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        uint32_t kept = 0;
        for (uint32_t k0 = 0; k0 < nmax; k0 += 8) {  // (loading a pass of indices ahead, as the SpMV does, was slower here)
            const uint32_t k = k0 + sub;
            bool in = false;
            uint32_t j = 0;
            if (k < n) {
                j = __ldg(src + k);
                const float4 pj = __ldg(spos + j);
                const float3 dl = make_float3(PSE_SUB(pi.x, pj.x), PSE_SUB(pi.y, pj.y), PSE_SUB(pi.z, pj.z));
                const float3 r = box.min_image_fast(dl);
                const float d = r.x * r.x + r.y * r.y + r.z * r.z;
                in = d < rp.rcut_sq && d >= rp.dr_sq;
                if (r.x != dl.x || r.y != dl.y || r.z != dl.z) j |= PSE_WRAP_BIT;  // the pair crosses a periodic boundary
            }
            const uint32_t ball = (__ballot_sync(0xffffffffu, in) >> (grp * 8)) & 0xffu;
            if (in) nl_act[h + kept + __popc(ball & ((1u << sub) - 1u))] = j;
            kept += __popc(ball);
        }
        if (live && sub == 0) nn_act[row] = kept;
    }
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
// Persistent grid (a multiple of the SM count): each block walks row groups with a grid stride, so the
// number of partial sums (and of arrivals on the finishing counter) is O(SMs), not O(N).
// SMEM_TABLE: the real-space table is staged once per (persistent) block in shared memory as float2 knots; the
// per-pair lookup is then two LDS.64 instead of a 16-byte global gather that touches up to 32 cache lines per
// warp (the L1 wavefront limiter of the first version, profiles/r1_ncu_summary.md).
enum { TABLE_GLOBAL = 0, TABLE_SHARED = 1, TABLE_POLY = 2 };
#ifndef SPMV_ROW_PASS
#define SPMV_ROW_PASS 48
#endif
#ifndef SPMV_AHEAD
#define SPMV_AHEAD 1
#endif
// SPMV_ROW_PASS: neighbour entries of a row covered by one pass (SPMV_ROW_PASS / TPP index registers per lane)
// DUAL (first Lanczos iteration of a full step): a second vector x2 (the slot-ordered forces) is multiplied in the same
// pass, y2 = M x2 - the pair geometry and f, g are shared, so the separate deterministic SpMV of the step disappears.
template <int TPP, int MODE, int TABLE, bool PRUNED, bool DUAL = false>
__global__ void __launch_bounds__(256, DUAL ? 3 : 4)
spmv_kernel(const PX* __restrict__ px, float4* __restrict__ y, uint32_t N,
            const uint32_t* __restrict__ nn, const uint32_t* __restrict__ head, const uint32_t* __restrict__ nl,
            const float4* __restrict__ gtable, ChebCoef cheb, RealParams rp,
            PseBox box, LanczosArgs la, uint32_t row_begin = 0, const float4* __restrict__ x2 = nullptr,
            float4* __restrict__ y2 = nullptr) {
    constexpr int ROWS = 256 / TPP;
    const int sub = threadIdx.x % TPP;
    extern __shared__ __align__(16) float2 stab[];
    if (TABLE == TABLE_SHARED) {
        for (int k = threadIdx.x; k <= rp.ewald_n; k += blockDim.x) {
            const float4 t = __ldg(gtable + k);
            stab[k] = make_float2(t.x, t.y);
            if (k == rp.ewald_n) stab[k + 1] = make_float2(t.z, t.w);
        }
        __syncthreads();
    }
    typename std::conditional<TABLE == TABLE_SHARED, TableShared, typename std::conditional<TABLE == TABLE_POLY, TableCheb, TableGlobal>::type>::type table;
    if constexpr (TABLE == TABLE_SHARED) table.t = stab;
    else if constexpr (TABLE == TABLE_POLY) { table.c = cheb; table.t = gtable; }
    else table.t = gtable;
    float part = 0.f;
    float beta = 0.f, s = 0.f;
    if (MODE == SPMV_LANCZOS) {
        beta = __ldcg(la.beta_j);
        s = beta > 1e-8f ? 1.0f / beta : 0.f;  // breakdown guard, PSEv1/Brownian.cu:507-510
    }
    // row header (neighbour count, list offset) of the NEXT row group is fetched while the current one is processed
    uint32_t n_next = 0, head_next = 0;
    {
        const uint32_t r = row_begin + blockIdx.x * ROWS + threadIdx.x / TPP;
        if (r < N) { n_next = __ldg(nn + r); head_next = __ldg(head + r); }
    }
    for (uint32_t row0 = row_begin + blockIdx.x * ROWS; row0 < N; row0 += gridDim.x * ROWS) {  // rows [row_begin, N)
        const uint32_t row = row0 + threadIdx.x / TPP;
        float3 u = make_float3(0.f, 0.f, 0.f), u2 = u;
        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), xi = pi;
        float4 vp = pi;
        const bool live = row < N;
        const uint32_t n = n_next;
        const uint32_t* __restrict__ list = nl + head_next;
        {
            const uint32_t r = row + gridDim.x * ROWS;
            if (r < N) { n_next = __ldg(nn + r); head_next = __ldg(head + r); }
        }
        if (live) {
            ld_px(px + row, pi, xi);
            if (MODE == SPMV_LANCZOS && !la.first && sub == 0) vp = __ldg(la.v_prev + row);
            // The chain index -> record -> arithmetic is two dependent L2 latencies; with ~5 entries per lane the loop is
            // too short to hide them by occupancy alone (42% of the stall samples, profiles/r1_summary.md).  So a lane
            // loads all of its (up to SPMV_ROW_PASS / TPP) indices of the row at once, and the gathers run one pair ahead of the
            // arithmetic.  Lane `sub` still owns entries sub, sub + TPP, ... in that order: sums are unchanged.
            constexpr int SLOTS = TPP >= 32 ? 2 : SPMV_ROW_PASS / TPP;
            for (uint32_t base = 0; base < n; base += SLOTS * TPP) {
                uint32_t idx[SLOTS];
#pragma unroll
                for (int t = 0; t < SLOTS; ++t) {
                    const uint32_t k = base + sub + t * TPP;
                    idx[t] = k < n ? __ldg(list + k) : 0xffffffffu;
                }
                constexpr int AHEAD = SPMV_AHEAD, RING = AHEAD + 1;  // pairs of gathers in flight ahead of the arithmetic
                float4 p[RING][2], x[RING][2], xb[RING][2];
#pragma unroll
                for (int h = 0; h < AHEAD; ++h)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (2 * h + q < SLOTS && idx[2 * h + q] != 0xffffffffu) {
                            const uint32_t jj = PRUNED ? idx[2 * h + q] & ~PSE_WRAP_BIT : idx[2 * h + q];
                            ld_px(px + jj, p[h % RING][q], x[h % RING][q]);
                            if (DUAL) xb[h % RING][q] = __ldg(x2 + jj);
                        }
#pragma unroll
                for (int t = 0; t < SLOTS; t += 2) {
                    const int cur = (t / 2) % RING, nxt = (t / 2 + AHEAD) % RING;
                    if (t + 2 * AHEAD < SLOTS) {
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            if (idx[t + 2 * AHEAD + q] != 0xffffffffu) {
                                const uint32_t jj = PRUNED ? idx[t + 2 * AHEAD + q] & ~PSE_WRAP_BIT : idx[t + 2 * AHEAD + q];
                                ld_px(px + jj, p[nxt][q], x[nxt][q]);
                                if (DUAL) xb[nxt][q] = __ldg(x2 + jj);
                            }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (idx[t + q] != 0xffffffffu) {
                            const float4 pj = p[cur][q];
                            float3 r = make_float3(PSE_SUB(pi.x, pj.x), PSE_SUB(pi.y, pj.y), PSE_SUB(pi.z, pj.z));
                            if (!PRUNED) r = box.min_image_fast(r);
                            else if (idx[t + q] & PSE_WRAP_BIT) r = box.min_image(r);
                            const float d = r.x * r.x + r.y * r.y + r.z * r.z;
                            // a pruned list was filtered with this very arithmetic at these very positions
                            if (PRUNED || (d < rp.rcut_sq && d >= rp.dr_sq)) {
                                if (DUAL) rpy_pair2(r, d, x[cur][q], xb[cur][q], table, rp, u, u2);
                                else rpy_pair(r, d, x[cur][q], table, rp, u);
                            }
                        }
                    }
                }
            }
        }
        u.x = group_sum<TPP>(u.x);
        u.y = group_sum<TPP>(u.y);
        u.z = group_sum<TPP>(u.z);
        if (DUAL) {
            u2.x = group_sum<TPP>(u2.x); u2.y = group_sum<TPP>(u2.y); u2.z = group_sum<TPP>(u2.z);
            if (live && sub == 0) {
                const float4 fi = __ldg(x2 + row);
                y2[row] = make_float4(u2.x + rp.self * fi.x, u2.y + rp.self * fi.y, u2.z + rp.self * fi.z, 0.f);
            }
        }
        if (live && sub == 0) {
            if (MODE == SPMV_PLAIN) {
                y[row] = make_float4(u.x + rp.self * xi.x, u.y + rp.self * xi.y, u.z + rp.self * xi.z, 0.f);
            } else {
                const float3 v = make_float3(s * xi.x, s * xi.y, s * xi.z);
                float3 mv = make_float3(s * (u.x + rp.self * xi.x), s * (u.y + rp.self * xi.y), s * (u.z + rp.self * xi.z));
                if (!la.first) { mv.x -= beta * vp.x; mv.y -= beta * vp.y; mv.z -= beta * vp.z; }
                la.v_out[row] = make_float4(v.x, v.y, v.z, 0.f);
                y[row] = make_float4(mv.x, mv.y, mv.z, 0.f);
                part += v.x * mv.x + v.y * mv.y + v.z * mv.z;
            }
        }
    }
    if (MODE == SPMV_LANCZOS) {
        __shared__ float red[32];
        const float tot = block_sum(part, red);
        grid_sum_finish(tot, la.partials, la.counter, la.alpha_out, red);
    }
}

// w = y - alpha_j v_j ;  beta_{j+1} = ||w|| ; u_{j+1} = w (left unnormalised; the next
// spmv_kernel<LANCZOS> folds 1/beta_{j+1} in).   PSEv1/Brownian.cu:493-514
__global__ void __launch_bounds__(256)
lanczos_update_kernel(const float4* __restrict__ y, const float4* __restrict__ vj, PX* __restrict__ px, uint32_t N,
                      const float* __restrict__ alpha_j, float* __restrict__ beta_next, float* partials,
                      unsigned int* counter, uint32_t row_begin = 0, const double* __restrict__ red2 = nullptr,
                      float* __restrict__ alpha_store = nullptr) {
    __shared__ float red[32];
    float part = 0.f;
    // slab-decomposed: rows [row_begin, N); (v.y, |y|^2, |v|^2) arrive all-reduced in red2 (lanczos_dots_kernel), no second
    // reduction needed: |y - a v|^2 = |y|^2 - a^2 (2 - |v|^2) with a = v.y, exactly, for the stored float vectors
    const double ad = red2 ? __ldcg(red2) : 0.0;
    const float a = red2 ? (float)ad : __ldcg(alpha_j);
    for (uint32_t i = row_begin + blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 yy = __ldg(y + i), v = __ldg(vj + i);
        float3 w = make_float3(yy.x - a * v.x, yy.y - a * v.y, yy.z - a * v.z);
        px[i].x = make_float4(w.x, w.y, w.z, 0.f);
        part += w.x * w.x + w.y * w.y + w.z * w.z;
    }
    if (red2) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            *alpha_store = a;
            *beta_next = (float)sqrt(fmax(__ldcg(red2 + 1) - ad * ad * (2.0 - __ldcg(red2 + 2)), 0.0));
        }
        return;
    }
    float tot = block_sum(part, red);
    grid_sum_finish(tot, partials, counter, beta_next, red, /*take_sqrt=*/true);
}

// slab-decomposed Lanczos: (v.y, |y|^2, |v|^2) over the rank's rows [r0, r1), in double (the combination above cancels two to
// three digits); the three sums are all-reduced over ranks in ONE exchange per iteration.  Deterministic: the block that
// arrives last adds the per-block partials in index order.
__global__ void __launch_bounds__(256)
lanczos_dots_kernel(const float4* __restrict__ y, const float4* __restrict__ v, uint32_t r0, uint32_t r1, double* __restrict__ partials /* [3][gridDim.x] */,
                    unsigned int* __restrict__ counter, double* __restrict__ out3) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (uint32_t i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x) {
        const float4 yy = __ldg(y + i), vv = __ldg(v + i);
        a += (double)vv.x * yy.x + (double)vv.y * yy.y + (double)vv.z * yy.z;
        b += (double)yy.x * yy.x + (double)yy.y * yy.y + (double)yy.z * yy.z;
        c += (double)vv.x * vv.x + (double)vv.y * vv.y + (double)vv.z * vv.z;
    }
    __shared__ double sm[3][8];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { sm[0][wid] = a; sm[1][wid] = b; sm[2][wid] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[k][w];
            partials[(size_t)k * gridDim.x + blockIdx.x] = t;
        }
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && wid < 3) {   // warp k adds the partials of sum k: lane-strided, then a fixed shuffle tree
        __threadfence();
        double t = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) t += __ldcg(partials + (size_t)wid * gridDim.x + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) out3[wid] = t;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) *counter = 0u;
}

// boundary rows of the vector being multiplied <-> contiguous exchange buffers (slab-decomposed SpMV halo)
__global__ void pack_px_kernel(const PX* __restrict__ px, uint32_t row0, uint32_t n, float4* __restrict__ buf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = px[row0 + i].x;
}
__global__ void unpack_px_kernel(PX* __restrict__ px, uint32_t row0, uint32_t n, const float4* __restrict__ buf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) px[row0 + i].x = buf[i];
}

// |px.x|^2 (or its square root) over xyz -> *out (deterministic)
__global__ void __launch_bounds__(256)
dot_px_kernel(const PX* __restrict__ px, uint32_t N, float* out, float* partials, unsigned int* counter, bool take_sqrt) {
    __shared__ float red[32];
    float part = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&px[i].x);
        part += p.x * p.x + p.y * p.y + p.z * p.z;
    }
    float tot = block_sum(part, red);
    grid_sum_finish(tot, partials, counter, out, red, take_sqrt);
}

// out[perm[slot]] (+)= scale * sum_k c[k] V[k][slot]   (PSEv1/Helper.cu:251-279 + the final rescale
// PSEv1/Brownian.cu:739), result scattered back to particle-id order and accumulated into U.
__global__ void __launch_bounds__(256)
basis_combine_kernel(const float4* __restrict__ V, const float* __restrict__ c, int m, uint32_t N, size_t stride,
                     const float* __restrict__ psinorm, float thermal, const uint32_t* __restrict__ perm,
                     float4* __restrict__ U, int accumulate, const float4* __restrict__ ydet /* slot-ordered M_real F or null */,
                     uint32_t row_begin = 0) {
    const uint32_t s = row_begin + blockIdx.x * blockDim.x + threadIdx.x;   // rows [row_begin, N); V is indexed by row
    if (s >= N) return;
    float3 acc = make_float3(0.f, 0.f, 0.f);
    for (int k = 0; k < m; ++k) {
        const float4 v = __ldg(V + (size_t)k * stride + s);
        const float ck = __ldg(c + k);
        acc.x += v.x * ck; acc.y += v.y * ck; acc.z += v.z * ck;
    }
    const float sc = __ldcg(psinorm) * thermal;
    const uint32_t p = perm ? perm[s] : s;   // (null: slot order, slab-decomposed engines)
    float4 o = U[p];  // .w (the mass column of the reference's velocity array, PSEv1/Helper.cu:131) is preserved
    if (!accumulate) { o.x = 0.f; o.y = 0.f; o.z = 0.f; }
    if (ydet) { const float4 yd = __ldg(ydet + s); o.x += yd.x; o.y += yd.y; o.z += yd.z; }
    o.x += sc * acc.x; o.y += sc * acc.y; o.z += sc * acc.z;
    U[p] = o;
}

// U[perm[slot]] (+)= y[slot].  Accumulating writes keep U.w (the reference's LinearCombination keeps the mass column of
// d_vel, PSEv1/Helper.cu:131); the plain write stores (y, 0) as gpu_stokes_Mreal_kernel does (PSEv1/Mobility.cu:632,684).
__global__ void scatter_add_kernel(const float4* __restrict__ y, const uint32_t* __restrict__ perm, uint32_t N,
                                   float4* __restrict__ U, int accumulate, uint32_t row_begin = 0) {
    const uint32_t s = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const float4 v = __ldg(y + s);
    const uint32_t p = perm ? perm[s] : s;
    float4 o = accumulate ? U[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x += v.x; o.y += v.y; o.z += v.z;
    U[p] = o;
}


// ---- short-range pair forces on the same list (pse_pair_force; stand-in for the HOOMD pair potentials that fill
// net_force before Stokes::integrateStepOne reads it, PSEv1/Stokes.cc:447,457) ---------------------------------
struct PairParams {
    int kind;  // PSE_PAIR_LJ / WCA / HARMONIC
    float eps, sigma, rcut_sq, rcut, shift;
};
// 8 lanes per row of the buffered list; F[perm[slot]] (+)= (sum_j f(r) r_ij, 1/2 sum_j U(r))
__global__ void __launch_bounds__(256)
pair_force_kernel(const float4* __restrict__ spos, uint32_t N, const uint32_t* __restrict__ nn, const uint32_t* __restrict__ head,
                  const uint32_t* __restrict__ nl, const uint32_t* __restrict__ perm, PairParams pp, PseBox box,
                  float4* __restrict__ F, int accumulate) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < N) {
        const float4 pi = __ldg(spos + row);
        const uint32_t n = __ldg(nn + row);
        const uint32_t* __restrict__ list = nl + __ldg(head + row);
        for (uint32_t k = sub; k < n; k += 8) {
            const float4 pj = __ldg(spos + __ldg(list + k));
            const float3 r = box.min_image_fast(make_float3(PSE_SUB(pi.x, pj.x), PSE_SUB(pi.y, pj.y), PSE_SUB(pi.z, pj.z)));
            const float r2 = r.x * r.x + r.y * r.y + r.z * r.z;
            if (r2 >= pp.rcut_sq || r2 <= 0.f) continue;
            float f_over_r, u;
            if (pp.kind == 2) {  // harmonic (dpd_conservative)
                const float rr = sqrtf(r2);
                f_over_r = pp.eps * (1.0f / rr - 1.0f / pp.rcut);
                u = pp.eps * (pp.rcut - rr) - pp.eps * (pp.rcut_sq - r2) / (2.0f * pp.rcut);
            } else {  // Lennard-Jones, optionally shifted (WCA)
                const float s2 = pp.sigma * pp.sigma / r2, s6 = s2 * s2 * s2;
                f_over_r = 24.0f * pp.eps * s6 * (2.0f * s6 - 1.0f) / r2;
                u = 4.0f * pp.eps * s6 * (s6 - 1.0f) + pp.shift;
            }
            acc.x += f_over_r * r.x; acc.y += f_over_r * r.y; acc.z += f_over_r * r.z; acc.w += 0.5f * u;
        }
    }
    acc.x = group_sum<8>(acc.x); acc.y = group_sum<8>(acc.y); acc.z = group_sum<8>(acc.z); acc.w = group_sum<8>(acc.w);
    if (row < N && sub == 0) {
        const uint32_t p = perm[row];
        float4 o = accumulate ? F[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
        F[p] = o;
    }
}

// Tile-owned Gaussian spreading and interpolation (the production path for P <= 10 on grids >= 22^3).
//
// Reference: gpu_stokes_Spread_kernel (PSEv1/Mobility.cu:114-252) launches one block per particle and
// issues 3 P^3 global float atomics per particle into zero-filled complex grids; gpu_stokes_Contract_kernel
// (:325-477, "THE SLOW STEP" :449) gathers 3 P^3 scattered 4-byte values per particle from global memory.
//
// Here the grid is cut into TILE^3 node tiles and the particles are binned by the tile of their support
// origin ("W order", rebuilt every call because positions move):
//   spread_tile_kernel  one block OWNS one node tile: it accumulates, in shared memory, the contribution of
//                       every particle whose support reaches the tile (the particles of <= 2x2x2 origin cells),
//                       one particle at a time with one thread per support node, then writes the tile once with
//                       plain coalesced stores.  No atomics, no zero-fill pass, summation order fixed ->
//                       bitwise reproducible.
//   interp_tile_kernel  one block stages the (TILE+P-1)^3 halo tile of the three velocity grids in shared
//                       memory and its warps interpolate the particles of that origin cell from it.
// Weights are evaluated as w_xy(i,j) * w_z(k) (P^2 + P exponentials per particle instead of P^3); the xy
// factor is not separable further because a sheared box displaces node x by xy*y (PSEv1/Mobility.cu:230).
#pragma once
#include "wave.cuh"

#define TILE 16
#define TILE_ZS 18       // padded z stride of the accumulator tile (bank spread for the 6x6 (j,k) footprints)
#ifndef SPREAD_CHUNK
#define SPREAD_CHUNK 32  // particles whose weights are staged at once
#endif
#define TILED_MAX_P 10   // 3P validity bits must fit one 32-bit word

struct TileGrid {
    int ntx, nty, ntz, ntile;
    int tile0;       // first tile handled by a launch (0 unless the grid is sharded into x-slabs of tiles)
    int tx, ty, tz;  // tile extent in nodes per axis (TILE^3 for the tile-owned gather kernels; per-P shapes for spread2)
    int cp, cy, cz, cs;  // spread2 accumulator geometry: residue modulus (= P), cells per axis (y, z), words per cell
    int rs;              // floats per W record of the v2 kernels (12 header words + factor row)
    int dbg;             // measurement knock-outs of spread2_kernel (PSE_SPREAD_DBG bit 0: no accumulation, 1: no merge, 2: no clear)
};
// row stride (floats) of the Gaussian factor rows: P*P + P rounded up to 16 bytes, so that any run of rows is a legal
// source of a bulk (TMA) copy
__host__ __device__ constexpr int wrow_stride(int P) { return ((P * P + P) + 3) / 4 * 4; }

// ---- binning by support-origin tile --------------------------------------------------------------
// row_lo/row_hi (x-tile rows, sharded calls only): particles whose origin tile lies outside rows [row_lo, row_hi) - and, when
// row_wrap >= 0, outside row row_wrap - are left out of the binning altogether (cell_of = ~0).
__global__ void wbin_kernel(const float4* __restrict__ spos, uint32_t N, PseBox box, WaveParams wp, TileGrid tg,
                            int4* __restrict__ org, uint32_t* __restrict__ cell_of, uint32_t* __restrict__ count,
                            int row_lo = 0, int row_hi = 1 << 30, int row_wrap = -1, uint32_t slot_begin = 0) {
    // slots [slot_begin, N) are binned (slab-decomposed calls pass the rank's own slot range)
    const uint32_t s = slot_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const float4 p = __ldg(spos + s);
    const Support o = support_origin(box, wp, p.x, p.y, p.z);
    const int x = wrap_node(o.x0, wp.Nx), y = wrap_node(o.y0, wp.Ny), z = wrap_node(o.z0, wp.Nz);
    const int trow = x / tg.tx;
    if (!((trow >= row_lo && trow < row_hi) || trow == row_wrap)) { cell_of[s] = 0xffffffffu; return; }
    const uint32_t c = ((uint32_t)trow * tg.nty + y / tg.ty) * tg.ntz + z / tg.tz;
    const int sx = (o.x0 - x) / wp.Nx + 1, sy = (o.y0 - y) / wp.Ny + 1, sz = (o.z0 - z) / wp.Nz + 1;  // in {0,1,2}
    org[s] = make_int4(x, y, z, sx | (sy << 2) | (sz << 4));
    cell_of[s] = c;
    atomicAdd(count + c, 1u);
}

// part bit 0: everything that depends on the positions only (wpos, worg, wid, record header words 3 .. 11);
// part bit 1: the force (wF / record header words 0 .. 2).  Split so that a step can bin and weight while its forces are
// still on their way from the host (pse_step_host_async).
__global__ void wgather_kernel(const float4* __restrict__ spos, const float4* __restrict__ sF, const int4* __restrict__ org,
                               const uint32_t* __restrict__ wperm, const uint32_t* __restrict__ perm, uint32_t N,
                               float4* __restrict__ wpos, float4* __restrict__ wF, int4* __restrict__ worg,
                               uint32_t* __restrict__ wid, TileGrid tg, int4* __restrict__ wrec,
                               const uint32_t* __restrict__ nbinned = nullptr, int part = 3) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= N || (nbinned && w >= __ldg(nbinned))) return;
    const uint32_t s = wperm[w];
    // particle id of W slot w: interpolation writes U[id] without chasing two permutations (perm == null: the slot itself,
    // slab-decomposed engines collect velocities in slot order)
    const uint32_t id = perm ? __ldg(perm + s) : s;
    int4* h = wrec ? reinterpret_cast<int4*>(reinterpret_cast<float*>(wrec) + (size_t)w * tg.rs) : nullptr;
    if (part & 1) {
        wid[w] = id;
        wpos[w] = __ldg(spos + s);
        const int4 o = org[s];
        if (h) {
            // header of the W record streamed by spread2_kernel / interp2_kernel (wave_v2.cuh), 12 words:
            //   (Fx, Fy, Fz, particle id | accumulator word offset of the origin's cell, origin residues x, y, z |
            //    origin inside the tile x, y, z, -); the Gaussian factor row follows (wweights_kernel)
            const int lx = o.x % tg.tx, ly = o.y % tg.ty, lz = o.z % tg.tz;
            h[0] = make_int4(0, 0, 0, (int)id);   // (zero force until the force part runs)
            h[1] = make_int4((((lx / tg.cp) * tg.cy + ly / tg.cp) * tg.cz + lz / tg.cp) * tg.cs, lx % tg.cp, ly % tg.cp, lz % tg.cp);
            h[2] = make_int4(lx, ly, lz, 0);
        }
        worg[w] = o;
    }
    if (part & 2) {
        const float4 f = sF ? __ldg(sF + s) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (h) h[0] = make_int4(__float_as_int(f.x), __float_as_int(f.y), __float_as_int(f.z), (int)id);   // one 16-byte store
        else if (sF) wF[w] = f;
    }
}

// candidate origin cells of one dimension for a tile starting at node t0 with extent e:
// origins in [t0 - P + 1, t0 + e - 1] (mod N)
__device__ __forceinline__ int tile_candidates(int t0, int e, int P, int N, int* cand) {
    int n = 0;
    int pos = t0 - P + 1;
    if (pos < 0) pos += N;
    int remaining = e + P - 1;
    if (remaining > N) remaining = N;
    while (remaining > 0 && n < 4) {
        const int c = pos / TILE;
        int cell_end = (c + 1) * TILE;
        if (cell_end > N) cell_end = N;
        bool dup = false;
        for (int q = 0; q < n; ++q) dup |= (cand[q] == c);
        if (!dup) cand[n++] = c;
        remaining -= cell_end - pos;
        pos = cell_end == N ? 0 : cell_end;
    }
    return n;
}

// Gaussian factors of one particle for the WRAPPED support node (ix, iy, iz), split as w_xy(i,j) * w_z(k).
// The node position and the minimum-image displacement are formed operation by operation as the reference does
// (PSEv1/Mobility.cu:222-238: h*index - L/2, shear, node - pos, BoxDim::minImage) because at L ~ 240 one ulp
// of a coordinate is ~8e-6 and any re-association of these sums moves individual weights by ~1e-5 relative —
// harmless for accuracy, but it would eat the whole 1e-5 parity budget.  Only exp(a+b) -> exp(a)*exp(b)
// differs (1e-7).
__device__ __forceinline__ float weight_z(const PseBox& box, const WaveParams& wp, int iz, float pz) {
    const float gz = wp.hz * (float)iz - box.Lz * 0.5f;
    float rz = gz - pz;
    rz = PSE_SUB(rz, PSE_MUL(box.Lz, rintf(PSE_MUL(rz, box.Lzinv))));
    return expf(-wp.expfac * (rz * rz));
}
__device__ __forceinline__ float weight_xy(const PseBox& box, const WaveParams& wp, int ix, int iy, float px, float py,
                                           float pref) {
    float gx = wp.hx * (float)ix - box.Lx * 0.5f;
    const float gy = wp.hy * (float)iy - box.Ly * 0.5f;
    gx = gx + box.xy * gy;
    float rx = gx - px, ry = gy - py;
    const float img = rintf(PSE_MUL(ry, box.Lyinv));
    ry = PSE_SUB(ry, PSE_MUL(box.Ly, img));
    rx = PSE_SUB(rx, PSE_MUL(PSE_MUL(box.Ly, box.xy), img));
    rx = PSE_SUB(rx, PSE_MUL(box.Lx, rintf(PSE_MUL(rx, box.Lxinv))));
    return pref * expf(-wp.expfac * (rx * rx + ry * ry));
}
// worg.w packs the wrap shifts of the three axes: unwrapped = wrapped + (shift - 1) * N, 2 bits per axis
__device__ __forceinline__ int3 unwrapped_origin(const int4 o, const WaveParams& wp) {
    return make_int3(o.x + ((o.w & 3) - 1) * wp.Nx, o.y + (((o.w >> 2) & 3) - 1) * wp.Ny, o.z + (((o.w >> 4) & 3) - 1) * wp.Nz);
}

// ---- Gaussian factors of every particle, once per call ---------------------------------------------------
// wwt[w][0 .. P*P) = w_xy(i, j) (prefac included), wwt[w][P*P .. P*P + P) = w_z(k), W order.  A particle is visited by
// (1 + (P-1)/TILE)^3 ~ 2.3 tiles when spreading and once when interpolating; evaluating the reference-exact factors
// (~100 instructions each with the minimum-image arithmetic and expf) in every visit was a third of the spreading
// kernel's instructions (profiles/r1_summary.md).  P (P + 1) is even, so rows are 8-byte aligned for cp.async.
// A block computes the rows of WW_PB consecutive particles into shared memory - one task per (particle, i) row of w_xy
// and one per particle for the w_z row, xy tasks first so that warps do not mix the two - and writes them out as one
// contiguous, fully coalesced range (scattered 4-byte stores of single rows ran at a third of the speed).
#define WW_PB 32
template <int P>
__global__ void __launch_bounds__(256)
wweights_kernel(const float4* __restrict__ wpos, const int4* __restrict__ worg, uint32_t N, PseBox box, WaveParams wp,
                float* __restrict__ wwt, int dst_stride /* floats between the rows of consecutive particles */) {
    constexpr int PP = P * P, WS = wrow_stride(P);  // (pad words of a row are never read)
    __shared__ __align__(16) float rows[WW_PB * WS];
    const uint32_t w0 = blockIdx.x * WW_PB;
    const int np = (int)min((uint32_t)WW_PB, N - w0);
    for (int t = threadIdx.x; t < WW_PB * (P + 1); t += blockDim.x) {
        if (t < WW_PB * P) {
            const int q = t / P, i = t - q * P;
            if (q < np) {
                const float4 pp = __ldg(wpos + w0 + q);
                const int4 o = __ldg(worg + w0 + q);
                const int ix = wrap_node(o.x + i, wp.Nx);
#pragma unroll
                for (int j = 0; j < P; ++j) rows[q * WS + i * P + j] = weight_xy(box, wp, ix, wrap_node(o.y + j, wp.Ny), pp.x, pp.y, wp.prefac);
            }
        } else {
            const int q = t - WW_PB * P;
            if (q < np) {
                const float pz = __ldg(&wpos[w0 + q].z);
                const int oz = __ldg(&worg[w0 + q].z);
#pragma unroll
                for (int k = 0; k < P; ++k) rows[q * WS + PP + k] = weight_z(box, wp, wrap_node(oz + k, wp.Nz), pz);
            }
        }
    }
    if constexpr (WS > PP + P) {
        constexpr int PAD = WS - PP - P;
        for (int t = threadIdx.x; t < WW_PB * PAD; t += blockDim.x) rows[(t / PAD) * WS + PP + P + t % PAD] = 0.f;
    }
    __syncthreads();
    // rows are 8-byte aligned (WS and dst_stride are even); with dst_stride == WS the block writes one contiguous range
    const float2* src = reinterpret_cast<const float2*>(rows);
    for (int t = threadIdx.x; t < np * (WS / 2); t += blockDim.x) {
        const int q = t / (WS / 2), r = t - q * (WS / 2);
        reinterpret_cast<float2*>(wwt + (size_t)(w0 + q) * dst_stride)[r] = src[t];
    }
}

// ---- spreading: one block per node tile -------------------------------------------------------------
// Structure of one block (P^3 threads rounded up to warps, at most 256):
//   filter   all threads scan the particles of the candidate origin cells (<= 4x4x4, normally 2x2x2), keep the
//            ones whose support reaches the tile and stage them, order preserved, in shared memory
//            (force, tile offset, per-axis validity bits, W index) - one global latency;
//   factors  per chunk of SPREAD_CHUNK staged particles the precomputed rows of wwt are copied in with cp.async,
//            double buffered: the copy of chunk c+1 is in flight while chunk c is scattered;
//   scatter  one particle at a time, one thread per support node: acc[node] += w F (plain shared-memory
//            adds; the per-particle barrier orders particles, so the sum is deterministic).  Two register sets
//            alternate so that the next particle's record and factors are fetched before the barrier.
//   store    the finished tile is written once, coalesced.
// dynamic smem: acc[3][TILE*TILE*TILE_ZS] | a_rec[CAP] f4 | a_mask[CAP] | a_w[CAP] | wbuf[2][CHUNK][P*P+P]
#ifndef SPREAD_CAP
#define SPREAD_CAP 384   // staged particles per filter round; sized so that three blocks fit one SM
#endif
#define SPREAD_MAX_SEG 64
#define SPREAD_NODE_BIAS 2048  // > (TILED_MAX_P - 1) * (TILE * TILE_ZS + TILE_ZS + 1)

template <int P> struct SpreadCfg {
    static constexpr int PPP = P * P * P;
    static constexpr int NT = PPP >= 256 ? 256 : ((PPP + 31) / 32) * 32;  // threads per block
    static constexpr int NPASS = (PPP + NT - 1) / NT;                     // support nodes per thread
    static constexpr int WS = wrow_stride(P);
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
// acc[addr + OFF] += w * f on a 32-bit shared-window address (keeps the generic->shared conversion out of the loop)
template <int OFF> __device__ __forceinline__ void smem_fma(uint32_t addr, float w, float f) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    v = fmaf(w, f, v);
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "f"(v));
}

template <int P>
__global__ void __launch_bounds__(SpreadCfg<P>::NT)
spread_tile_kernel(const float4* __restrict__ wF, const int4* __restrict__ worg, const float* __restrict__ wwt,
                   const uint32_t* __restrict__ wcell_start, WaveParams wp, TileGrid tg, float* __restrict__ grid) {
    extern __shared__ __align__(16) float smem[];
    constexpr int PP = P * P, PPP = PP * P;
    constexpr int NT = SpreadCfg<P>::NT, NW = NT / 32, npass = SpreadCfg<P>::NPASS, WS = SpreadCfg<P>::WS;
    constexpr int ACC = TILE * TILE * TILE_ZS;
    // P <= 6: validity bits (3P) and origin node (13 bits) share the record's 4th word - one broadcast load per particle
    constexpr bool PACKED = 3 * P + 13 <= 32;
    float* acc = smem;
    float4* a_rec = reinterpret_cast<float4*>(acc + 3 * ACC);
    uint32_t* a_mask = reinterpret_cast<uint32_t*>(a_rec + SPREAD_CAP);
    uint32_t* a_w = a_mask + SPREAD_CAP;
    float* wbuf = reinterpret_cast<float*>(a_w + SPREAD_CAP);  // [2][SPREAD_CHUNK][WS]
    __shared__ int s_cand[3][4];
    __shared__ int s_ncand[3];
    __shared__ uint32_t s_seg_b[SPREAD_MAX_SEG], s_seg_off[SPREAD_MAX_SEG + 1];
    __shared__ int s_nseg;
    __shared__ int s_warp_cnt[2][NW];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x + tg.tile0;
    const int bz = tile % tg.ntz, by = (tile / tg.ntz) % tg.nty, bx = tile / (tg.ntz * tg.nty);
    const int t0x = bx * TILE, t0y = by * TILE, t0z = bz * TILE;
    const int ex = min(TILE, wp.Nx - t0x), ey = min(TILE, wp.Ny - t0y), ez = min(TILE, wp.Nz - t0z);

    for (int i = tid; i < 3 * ACC / 4; i += NT) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 3) {
        const int t0 = tid == 0 ? t0x : tid == 1 ? t0y : t0z, e = tid == 0 ? ex : tid == 1 ? ey : ez;
        const int N = tid == 0 ? wp.Nx : tid == 1 ? wp.Ny : wp.Nz;
        s_ncand[tid] = tile_candidates(t0, e, P, N, s_cand[tid]);
    }
    __syncthreads();
    if (tid == 0) {  // candidate segments of the W-ordered particle arrays, fixed order
        int n = 0;
        uint32_t off = 0;
        for (int a = 0; a < s_ncand[0]; ++a)
            for (int b = 0; b < s_ncand[1]; ++b)
                for (int c = 0; c < s_ncand[2]; ++c) {
                    const uint32_t cell = ((uint32_t)s_cand[0][a] * tg.nty + s_cand[1][b]) * tg.ntz + s_cand[2][c];
                    const uint32_t cb = __ldg(wcell_start + cell), ce = __ldg(wcell_start + cell + 1);
                    if (ce > cb) { s_seg_b[n] = cb; s_seg_off[n] = off; off += ce - cb; ++n; }
                }
        s_seg_off[n] = off;
        s_nseg = n;
    }
    // this thread's support node(s) (i,j,k): constant over particles
    uint32_t my_acc[npass];  // shared-window byte address of acc[my node]
    int my_ij[npass], my_k[npass];
    uint32_t my_bits[npass];
#pragma unroll
    for (int r = 0; r < npass; ++r) {
        const int t = tid + r * NT;
        const int i = t / PP, j = (t - i * PP) / P, k = t - i * PP - j * P;
        my_acc[r] = (uint32_t)__cvta_generic_to_shared(acc + (i * TILE + j) * TILE_ZS + k);
        my_ij[r] = i * P + j;
        my_k[r] = PP + k;
        my_bits[r] = t < PPP ? ((1u << i) | (1u << (P + j)) | (1u << (2 * P + k))) : 0xffffffffu;  // never matches
    }
    __syncthreads();
    const int nseg = s_nseg;
    const uint32_t ncandidates = s_seg_off[nseg];

    uint32_t next = 0;  // next candidate (flat index) to examine
    int trip = 0;
    while (next < ncandidates) {
        // ---- filter: ordered compaction of up to SPREAD_CAP reaching particles
        int nact = 0;
        while (next < ncandidates && nact <= SPREAD_CAP - NT) {
            const uint32_t t = next + tid;
            bool act = false;
            int lx = 0, ly = 0, lz = 0;
            uint32_t w = 0;
            if (t < ncandidates) {
                int sgi = 0;
                while (sgi + 1 < nseg && t >= s_seg_off[sgi + 1]) ++sgi;
                w = s_seg_b[sgi] + (t - s_seg_off[sgi]);
                const int4 o = __ldg(worg + w);
                lx = o.x - t0x; if (lx >= ex) lx -= wp.Nx;
                ly = o.y - t0y; if (ly >= ey) ly -= wp.Ny;
                lz = o.z - t0z; if (lz >= ez) lz -= wp.Nz;
                act = lx > -P && ly > -P && lz > -P;
            }
            const uint32_t ball = __ballot_sync(0xffffffffu, act);
            int* cnt = s_warp_cnt[trip & 1];  // the counters alternate: one barrier per filter trip
            if (lane == 0) cnt[wid] = __popc(ball);
            __syncthreads();
            int before = nact, total = 0;
#pragma unroll
            for (int q = 0; q < NW; ++q) { const int v = cnt[q]; total += v; if (q < wid) before += v; }
            if (act) {
                const int slot = before + __popc(ball & ((1u << lane) - 1u));
                uint32_t m = 0;
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    m |= (uint32_t)((unsigned)(lx + i) < (unsigned)ex) << i;
                    m |= (uint32_t)((unsigned)(ly + i) < (unsigned)ey) << (P + i);
                    m |= (uint32_t)((unsigned)(lz + i) < (unsigned)ez) << (2 * P + i);
                }
                const float4 F = __ldg(wF + w);
                const int onode = (lx * TILE + ly) * TILE_ZS + lz;  // tile index of the origin node (may be negative)
                if (PACKED) a_rec[slot] = make_float4(F.x, F.y, F.z, __int_as_float((int)(m | ((uint32_t)(onode + SPREAD_NODE_BIAS) << (3 * P)))));
                else { a_rec[slot] = make_float4(F.x, F.y, F.z, __int_as_float(4 * onode)); a_mask[slot] = m; }
                a_w[slot] = w;
            }
            nact += total;
            next += NT;
            ++trip;
        }
        __syncthreads();
        // ---- chunks of staged particles
        auto fetch_chunk = [&](int c0, int buf) {  // rows of wwt -> wbuf[buf], 8 bytes per copy
            const int nch = min(SPREAD_CHUNK, nact - c0);
            float* dst = wbuf + buf * (SPREAD_CHUNK * WS);
            for (int t = tid; t < nch * (WS / 2); t += NT) {
                const int q = t / (WS / 2), r = t - q * (WS / 2);
                cp_async8(dst + q * WS + 2 * r, wwt + (size_t)a_w[c0 + q] * WS + 2 * r);
            }
            cp_async_commit();
        };
        if (nact > 0) fetch_chunk(0, 0);
        int buf = 0;
        for (int c0 = 0; c0 < nact; c0 += SPREAD_CHUNK, buf ^= 1) {
            const int nch = min(SPREAD_CHUNK, nact - c0);
            if (c0 + SPREAD_CHUNK < nact) { fetch_chunk(c0 + SPREAD_CHUNK, buf ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
            const float* wq = wbuf + buf * (SPREAD_CHUNK * WS);
            const float4* recs = a_rec + c0;
            const uint32_t* masks = a_mask + c0;
            // two register sets (A: even particles, B: odd), each loaded one barrier ahead of its use
            float4 recA = recs[0], recB;
            uint32_t mA = PACKED ? 0u : masks[0], mB = 0;
            float wA[npass], wB[npass];
#pragma unroll
            for (int r = 0; r < npass; ++r) { wA[r] = wq[my_ij[r]] * wq[my_k[r]]; wB[r] = 0.f; }
            for (int q = 0; q < nch; q += 2) {
                if (q + 1 < nch) {
                    recB = recs[q + 1]; if (!PACKED) mB = masks[q + 1];
#pragma unroll
                    for (int r = 0; r < npass; ++r) wB[r] = wq[(q + 1) * WS + my_ij[r]] * wq[(q + 1) * WS + my_k[r]];
                }
                {
                    const uint32_t pk = (uint32_t)__float_as_int(recA.w);
                    const uint32_t mm = PACKED ? pk : mA;  // (bits above 3P never match my_bits)
                    const int base = PACKED ? 4 * ((int)(pk >> (3 * P)) - SPREAD_NODE_BIAS) : (int)pk;
#pragma unroll
                    for (int r = 0; r < npass; ++r)
                        if ((mm & my_bits[r]) == my_bits[r]) {
                            const uint32_t a = my_acc[r] + base;
                            smem_fma<0>(a, wA[r], recA.x);
                            smem_fma<4 * ACC>(a, wA[r], recA.y);
                            smem_fma<8 * ACC>(a, wA[r], recA.z);
                        }
                }
                __syncthreads();
                if (q + 1 >= nch) break;
                if (q + 2 < nch) {
                    recA = recs[q + 2]; if (!PACKED) mA = masks[q + 2];
#pragma unroll
                    for (int r = 0; r < npass; ++r) wA[r] = wq[(q + 2) * WS + my_ij[r]] * wq[(q + 2) * WS + my_k[r]];
                }
                {
                    const uint32_t pk = (uint32_t)__float_as_int(recB.w);
                    const uint32_t mm = PACKED ? pk : mB;  // (bits above 3P never match my_bits)
                    const int base = PACKED ? 4 * ((int)(pk >> (3 * P)) - SPREAD_NODE_BIAS) : (int)pk;
#pragma unroll
                    for (int r = 0; r < npass; ++r)
                        if ((mm & my_bits[r]) == my_bits[r]) {
                            const uint32_t a = my_acc[r] + base;
                            smem_fma<0>(a, wB[r], recB.x);
                            smem_fma<4 * ACC>(a, wB[r], recB.y);
                            smem_fma<8 * ACC>(a, wB[r], recB.z);
                        }
                }
                __syncthreads();
            }
        }
    }
    // ---- write the tile once: a thread moves four consecutive z nodes of the three components (TILE_ZS is even, so the
    // shared reads are 8-byte aligned; the global stores are 16-byte vectors when Nz allows it)
    const size_t G = (size_t)wp.Nx * wp.Ny * wp.Nz;
    const bool vec = (ez == TILE) && ((wp.Nz & 3) == 0);
    for (int t = tid; t < ex * ey * (TILE / 4); t += NT) {
        const int q = t & 3, row = t >> 2;
        int lx, ly;
        if (ey == TILE) { ly = row & (TILE - 1); lx = row >> 4; }
        else { lx = row / ey; ly = row - lx * ey; }
        const int node = (lx * TILE + ly) * TILE_ZS + 4 * q;
        const size_t idx = ((size_t)(t0x + lx) * wp.Ny + (t0y + ly)) * wp.Nz + (t0z + 4 * q);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 a = *reinterpret_cast<const float2*>(acc + c * ACC + node);
            const float2 b = *reinterpret_cast<const float2*>(acc + c * ACC + node + 2);
            if (vec) {
                *reinterpret_cast<float4*>(grid + c * G + idx) = make_float4(a.x, a.y, b.x, b.y);
            } else {
                const float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
                for (int z = 0; z < 4; ++z)
                    if (4 * q + z < ez) grid[c * G + idx + z] = v[z];
            }
        }
    }
}

static inline size_t spread_tile_smem(int P) {
    return (3 * (size_t)TILE * TILE * TILE_ZS) * sizeof(float) + SPREAD_CAP * (sizeof(float4) + 2 * sizeof(uint32_t)) +
           2 * (size_t)SPREAD_CHUNK * wrow_stride(P) * sizeof(float);
}

// ---- interpolation: one block per origin cell --------------------------------------------------------
// The block stages the (TILE+P-1)^3 halo tile of the three velocity grids in shared memory; each of its 16 warps
// walks particles of the cell.  A lane owns one (i, j) COLUMN of the support and walks its P z-nodes:
//   * tile strides (x: H*H + pad, y: H, z: 1) chosen so that the 32 columns of a pass fall into 32 different banks
//     (one wavefront per load instead of two with a node-per-lane mapping; interp_pad() searches the pad);
//   * w_xy of the column is the lane's own word of the prefetched factor row, w_z(k) comes by shuffle - no
//     shared-memory scratch for factors;
//   * columns beyond a multiple of 32 (4 of the 36 for P = 6) are handled one NODE per lane when they fit a warp.
// dynamic smem: g[3][H * XS].
#ifndef INTERP_THREADS
#define INTERP_THREADS 512
#endif
__host__ __device__ constexpr int interp_pad(int P) {
    const int H = TILE + P - 1, PP = P * P;
    int best_pad = 0, best_cost = 1 << 30;
    for (int pad = 0; pad < 32; ++pad) {
        const int XS = H * H + pad;
        int cost = 0;
        for (int c0 = 0; c0 < PP; c0 += 32) {
            int cnt[32] = {};
            int mx = 0;
            for (int c = c0; c < PP && c < c0 + 32; ++c) {
                const int i = c / P, j = c % P;
                const int b = (i * XS + j * H) % 32;
                if (++cnt[b] > mx) mx = cnt[b];
            }
            cost += mx;
        }
        if (cost < best_cost) { best_cost = cost; best_pad = pad; }
    }
    return best_pad;
}
template <int P>
__global__ void __launch_bounds__(INTERP_THREADS, 2)
interp_tile_kernel(const int4* __restrict__ worg, const float* __restrict__ wwt, const uint32_t* __restrict__ wcell_start,
                   const uint32_t* __restrict__ wpid, WaveParams wp,
                   TileGrid tg, const float* __restrict__ grid, float4* __restrict__ U, int accumulate) {
    extern __shared__ __align__(16) float smem[];
    constexpr int PP = P * P, NW = INTERP_THREADS / 32, WS = wrow_stride(P), NWD = PP + P;
    constexpr int NWR = (NWD + 31) / 32;  // factor words per lane
    constexpr int H = TILE + P - 1, XS = H * H + interp_pad(P), GT = H * XS;
    constexpr int NFULL = PP / 32;             // passes in which every lane has a column
    constexpr int LEFT = PP - 32 * NFULL;      // remaining columns
    constexpr bool LEFT_NODES = LEFT > 0 && LEFT * P <= 32;  // ... handled one node per lane
    float* g = smem;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t cell = blockIdx.x + tg.tile0;
    const uint32_t cb = __ldg(wcell_start + cell), ce = __ldg(wcell_start + cell + 1);
    if (cb == ce) return;
    const int bz = cell % tg.ntz, by = (cell / tg.ntz) % tg.nty, bx = cell / (tg.ntz * tg.nty);
    const int t0x = bx * TILE, t0y = by * TILE, t0z = bz * TILE;
    const size_t G = (size_t)wp.Nx * wp.Ny * wp.Nz;
    // first particle of this warp: issue its loads before staging the tile
    uint32_t w = cb + wid;
    int4 o_n = make_int4(0, 0, 0, 0);
    uint32_t id_n = 0;
    float wt_n[NWR];
    auto fetch = [&](uint32_t ww) {  // origin, output slot and the precomputed factor row of particle ww
        o_n = __ldg(worg + ww);
        id_n = __ldg(wpid + ww);
#pragma unroll
        for (int r = 0; r < NWR; ++r) wt_n[r] = lane + 32 * r < NWD ? __ldg(wwt + (size_t)ww * WS + lane + 32 * r) : 0.f;
    };
    if (w < ce) fetch(w);
    // stage the halo tile (periodic wrap per node): a warp takes whole x planes of the tile, its lanes run over the H*H
    // (y, z) nodes of the plane with z fastest -> coalesced row segments, one constant division per node
    for (int lx = wid; lx < H; lx += NW) {
        int x = t0x + lx; if (x >= wp.Nx) x -= wp.Nx;
        const float* gx = grid + (size_t)x * wp.Ny * wp.Nz;
        float* sx = g + lx * XS;
        for (int e = lane; e < H * H; e += 32) {
            const int ly = e / H, lz = e - ly * H;
            int y = t0y + ly; if (y >= wp.Ny) y -= wp.Ny;
            int z = t0z + lz; if (z >= wp.Nz) z -= wp.Nz;
            const float* src = gx + y * wp.Nz + z;
            sx[e] = __ldg(src);  // (node = lx * XS + ly * H + lz = lx * XS + e)
            sx[GT + e] = __ldg(src + G);
            sx[2 * GT + e] = __ldg(src + 2 * G);
        }
    }
    // this lane's columns: tile offset of (i, j, 0) for each full pass, and of the left-over column / node
    int col_off[NFULL > 0 ? NFULL : 1];
#pragma unroll
    for (int r = 0; r < NFULL; ++r) {
        const int c = lane + 32 * r;
        col_off[r] = (c / P) * XS + (c % P) * H;
    }
    int left_off = -1, left_col = 0, left_k = 0;
    if (LEFT > 0) {
        if (LEFT_NODES) {
            if (lane < LEFT * P) { left_col = 32 * NFULL + lane / P; left_k = lane % P; }
            else left_col = -1;
        } else {
            left_col = lane < LEFT ? 32 * NFULL + lane : -1;
        }
        if (left_col >= 0) left_off = (left_col / P) * XS + (left_col % P) * H + left_k;
    }
    __syncthreads();
    while (w < ce) {
        const int4 o = o_n;
        const uint32_t id = id_n;
        float wt[NWR];
#pragma unroll
        for (int r = 0; r < NWR; ++r) wt[r] = wt_n[r];
        const uint32_t wn = w + NW;
        if (wn < ce) fetch(wn);
        float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == 0) old = U[id];   // .w (mass in the reference's velocity array, PSEv1/Mobility.cu:474) is preserved
        float wz[P];
#pragma unroll
        for (int k = 0; k < P; ++k) wz[k] = __shfl_sync(0xffffffffu, wt[(PP + k) / 32], (PP + k) % 32);
        const float* gb = g + (o.x - t0x) * XS + (o.y - t0y) * H + (o.z - t0z);
        float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
        for (int r = 0; r < NFULL; ++r) {
            const float* gc = gb + col_off[r];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const float wgt = wt[r] * wz[k];
                ax = fmaf(wgt, gc[k], ax);
                ay = fmaf(wgt, gc[GT + k], ay);
                az = fmaf(wgt, gc[2 * GT + k], az);
            }
        }
        if (LEFT > 0) {
            // factor words of the left-over columns live in other lanes: fetch by shuffle (all lanes take part)
            const int src = left_col >= 0 ? left_col : 0;
            const float wxy = __shfl_sync(0xffffffffu, wt[NFULL], src % 32);
            if (LEFT_NODES) {
                float wzl = wz[0];
#pragma unroll
                for (int k = 1; k < P; ++k) wzl = left_k == k ? wz[k] : wzl;
                if (left_off >= 0) {
                    const float wgt = wxy * wzl;
                    ax = fmaf(wgt, gb[left_off], ax);
                    ay = fmaf(wgt, gb[GT + left_off], ay);
                    az = fmaf(wgt, gb[2 * GT + left_off], az);
                }
            } else if (left_off >= 0) {
                const float* gc = gb + left_off;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const float wgt = wxy * wz[k];
                    ax = fmaf(wgt, gc[k], ax);
                    ay = fmaf(wgt, gc[GT + k], ay);
                    az = fmaf(wgt, gc[2 * GT + k], az);
                }
            }
        }
        ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
        if (lane == 0) {
            if (!accumulate) { old.x = 0.f; old.y = 0.f; old.z = 0.f; }
            U[id] = make_float4(old.x + wp.quadW * ax, old.y + wp.quadW * ay, old.z + wp.quadW * az, old.w);  // quadrature weight h^3
        }
        w = wn;
    }
}

static inline size_t interp_tile_smem(int P) {
    const int H = TILE + P - 1, XS = H * H + interp_pad(P);
    return 3 * (size_t)H * XS * sizeof(float);
}

// ---- dispatch on the (runtime) support size ------------------------------------------------------------
#define PSE_FOR_EACH_P(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10)

static cudaError_t tiled_set_attributes(int P) {
    cudaError_t err = cudaSuccess;
    switch (P) {
#define X(p)                                                                                                                   \
    case p:                                                                                                                    \
        err = cudaFuncSetAttribute(spread_tile_kernel<p>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spread_tile_smem(p)); \
        if (err == cudaSuccess)                                                                                                \
            err = cudaFuncSetAttribute(interp_tile_kernel<p>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)interp_tile_smem(p)); \
        break;
        PSE_FOR_EACH_P(X)
#undef X
        default: err = cudaErrorInvalidValue;
    }
    return err;
}

static void launch_wweights(int P, cudaStream_t st, const float4* wpos, const int4* worg, uint32_t N, const PseBox& box,
                            const WaveParams& wp, float* wwt, int dst_stride = 0) {
    const unsigned int nb = (N + WW_PB - 1) / WW_PB;
    if (dst_stride == 0) dst_stride = wrow_stride(P);
    switch (P) {
#define X(p) case p: wweights_kernel<p><<<nb, 256, 0, st>>>(wpos, worg, N, box, wp, wwt, dst_stride); break;
        PSE_FOR_EACH_P(X)
#undef X
    }
}
static void launch_spread_tile(int P, cudaStream_t st, const float4* wF, const int4* worg, const float* wwt, const uint32_t* wstart,
                               const WaveParams& wp, const TileGrid& tg, float* grid, int ntiles = -1) {
    if (ntiles < 0) ntiles = tg.ntile;
    if (ntiles == 0) return;
    switch (P) {
#define X(p) case p: spread_tile_kernel<p><<<ntiles, SpreadCfg<p>::NT, spread_tile_smem(p), st>>>(wF, worg, wwt, wstart, wp, tg, grid); break;
        PSE_FOR_EACH_P(X)
#undef X
    }
}
static void launch_interp_tile(int P, cudaStream_t st, const int4* worg, const float* wwt, const uint32_t* wstart, const uint32_t* wid,
                               const WaveParams& wp, const TileGrid& tg, const float* grid,
                               float4* U, int accumulate, int ntiles = -1) {
    if (ntiles < 0) ntiles = tg.ntile;
    if (ntiles == 0) return;
    switch (P) {
#define X(p) case p: interp_tile_kernel<p><<<ntiles, INTERP_THREADS, interp_tile_smem(p), st>>>(worg, wwt, wstart, wid, wp, tg, grid, U, accumulate); break;
        PSE_FOR_EACH_P(X)
#undef X
    }
}

// Second-generation spreading / interpolation kernels: every particle is processed ONCE, by the block of the tile its
// support origin lies in, on a (T + P - 1)^3 window of grid nodes held in shared memory.
//
// Reference: gpu_stokes_Spread_kernel (PSEv1/Mobility.cu:114-252) issues 3 P^3 global float atomics per particle;
// gpu_stokes_Contract_kernel (:325-477) gathers 3 P^3 scattered values per particle from global memory.
//
// spread2_kernel.  The round-1 kernel (spread_tile_kernel, wave_tiled.cuh) let a block OWN a node tile and re-visited
// every particle from each of the ~2.3 tiles its support touches, with a block barrier per particle.  Here:
//   * RESIDUE OWNERSHIP: thread (a, b, c) of the P^3 threads owns every window node with (x mod P, y mod P, z mod P) =
//     (a, b, c).  A P^3 support contains each residue class exactly once, so every thread has exactly one node of every
//     particle, and a given node is only ever touched by one thread: the read-modify-write sequences need no atomics
//     and NO barriers between particles (program order of the owning thread is the only ordering required);
//   * the accumulator is stored residue-major, acc[cell][thread] with cell = (x / P, y / P, z / P) and a cell stride
//     that is a multiple of 32 words: the bank of an access is the thread index, conflict-free for every particle;
//   * the W records of the tile's particles (force, window position, Gaussian factor row: one contiguous range) are
//     streamed through a ring of shared-memory stages by bulk asynchronous copies (cp.async.bulk, completion on an
//     mbarrier), one elected thread as producer, full/empty barriers instead of block barriers;
//   * the window is merged into the grid with vector reductions (red.global.add.v4.f32) - about T^-3 (T + P - 1)^3
//     grid-sized reduction passes in L2 instead of the reference's 3 P^3 N scattered atomics.  The grid must be zero
//     on entry.  Summation order across blocks is not fixed, so results are reproducible to round-off (1e-7), not
//     bitwise; PSE_WAVE=v1 selects the bitwise-reproducible tile-owned kernel.
// interp2_kernel.  interp_tile_kernel generalised to box-shaped tiles: window staged with asynchronous copies, W records
// streamed through the same kind of bulk-copy ring.
#pragma once
#include "wave_tiled.cuh"

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy; src, dst and bytes multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// ---- W records ----------------------------------------------------------------------------------------------------
// One record per particle, W order (particles of a tile contiguous): 12 header words written by wgather_kernel
// (wave_tiled.cuh) followed by the Gaussian factor row of wweights_kernel.  A record is a multiple of 16 bytes, so any run
// of records is a legal bulk-copy source.
#define WREC_HDR 12
#define PSE_V2_SHAPES_FWD(X) X(6, 12, 12, 12) X(7, 15, 15, 12) X(8, 17, 17, 16)   // same list as PSE_V2_SHAPES below
__host__ __device__ constexpr int wrec_stride(int P) { return WREC_HDR + wrow_stride(P); }

// Position-only part of the W records in ONE kernel (the headers of wgather_kernel and the factor rows of wweights_kernel,
// without the wpos / worg round trip between them): a block builds WW_PB whole records in shared memory and writes them as
// one contiguous range.  The Gaussian factors are formed per axis first - P values of w_z per particle, and for an unsheared
// box (xy = 0: x and y decouple) P of w_x and P of w_y, 3 P exponentials instead of P^2 + P; a sheared box keeps the coupled
// w_xy(i, j) of weight_xy, one per thread.  Per-axis node positions and minimum-image displacements are formed operation by
// operation as in weight_xy / weight_z; only exp(a + b) -> exp(a) exp(b) differs (1e-7).  The force words stay zero until
// wgather_kernel's force part stores them.
template <int P>
__global__ void __launch_bounds__(256)
wrecords_kernel(const float4* __restrict__ spos, const int4* __restrict__ org, const uint32_t* __restrict__ wperm,
                const uint32_t* __restrict__ perm, uint32_t N, PseBox box, WaveParams wp, TileGrid tg, float* __restrict__ wrecs) {
    constexpr int PP = P * P, NWD = PP + P, RS = wrec_stride(P);
    __shared__ __align__(16) float rec[WW_PB * RS];
    __shared__ float ax[WW_PB][3][P];
    __shared__ float4 spp[WW_PB];
    __shared__ int4 sorg[WW_PB];
    const uint32_t w0 = blockIdx.x * WW_PB;
    const int np = (int)min((uint32_t)WW_PB, N - w0);
    if ((int)threadIdx.x < np) {
        const int q = threadIdx.x;
        const uint32_t s = wperm[w0 + q];
        const float4 pp = __ldg(spos + s);
        const int4 o = org[s];
        spp[q] = pp; sorg[q] = o;
        const uint32_t id = perm ? __ldg(perm + s) : s;
        const int lx = o.x % tg.tx, ly = o.y % tg.ty, lz = o.z % tg.tz;
        int4* h = reinterpret_cast<int4*>(rec + q * RS);
        h[0] = make_int4(0, 0, 0, (int)id);
        h[1] = make_int4((((lx / tg.cp) * tg.cy + ly / tg.cp) * tg.cz + lz / tg.cp) * tg.cs, lx % tg.cp, ly % tg.cp, lz % tg.cp);
        h[2] = make_int4(lx, ly, lz, 0);
    }
    __syncthreads();
    const bool sep = box.xy == 0.f;
    for (int t = threadIdx.x; t < WW_PB * P; t += blockDim.x) {
        const int q = t / P, i = t - q * P;
        if (q >= np) break;
        const float4 pp = spp[q];
        const int4 o = sorg[q];
        ax[q][2][i] = weight_z(box, wp, wrap_node(o.z + i, wp.Nz), pp.z);
        if (sep) {
            float rx = (wp.hx * (float)wrap_node(o.x + i, wp.Nx) - box.Lx * 0.5f) - pp.x;
            rx = PSE_SUB(rx, PSE_MUL(box.Lx, rintf(PSE_MUL(rx, box.Lxinv))));
            float ry = (wp.hy * (float)wrap_node(o.y + i, wp.Ny) - box.Ly * 0.5f) - pp.y;
            ry = PSE_SUB(ry, PSE_MUL(box.Ly, rintf(PSE_MUL(ry, box.Lyinv))));
            ax[q][0][i] = expf(-wp.expfac * (rx * rx));
            ax[q][1][i] = expf(-wp.expfac * (ry * ry));
        }
    }
    __syncthreads();
    {   // a thread keeps one factor index r = (i, j) or k for the whole block and walks the particles: no index arithmetic in the loop
        constexpr int QP = 256 / NWD;                      // particles per pass
        const int r = threadIdx.x % NWD, i = r / P, j = r - i * P;
        if ((int)threadIdx.x < QP * NWD)
            for (int q = threadIdx.x / NWD; q < np; q += QP) {
                float v;
                if (r >= PP) v = ax[q][2][r - PP];
                else if (sep) v = wp.prefac * (ax[q][0][i] * ax[q][1][j]);
                else v = weight_xy(box, wp, wrap_node(sorg[q].x + i, wp.Nx), wrap_node(sorg[q].y + j, wp.Ny), spp[q].x, spp[q].y, wp.prefac);
                rec[q * RS + WREC_HDR + r] = v;
            }
    }
    if constexpr (RS > WREC_HDR + NWD) {
        constexpr int PAD = RS - WREC_HDR - NWD;
        for (int t = threadIdx.x; t < WW_PB * PAD; t += blockDim.x) rec[(t / PAD) * RS + WREC_HDR + NWD + t % PAD] = 0.f;
    }
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(rec);
    float4* dst = reinterpret_cast<float4*>(wrecs + (size_t)w0 * RS);   // (a record is a multiple of 16 bytes)
    for (int t = threadIdx.x; t < np * (RS / 4); t += blockDim.x) dst[t] = src[t];
}
static void launch_wrecords(int P, cudaStream_t st, const float4* spos, const int4* org, const uint32_t* wperm, const uint32_t* perm, uint32_t N,
                            const PseBox& box, const WaveParams& wp, const TileGrid& tg, float* wrecs) {
    if (!N) return;
    const unsigned int nb = (N + WW_PB - 1) / WW_PB;
#define X(p, a, b, c) if (P == p) { wrecords_kernel<p><<<nb, 256, 0, st>>>(spos, org, wperm, perm, N, box, wp, tg, wrecs); return; }
    PSE_V2_SHAPES_FWD(X)
#undef X
}

// Ring of shared-memory stages fed by bulk copies: one elected thread is the producer, every warp a consumer.
// full[s] completes when the bytes of the chunk in stage s have landed; empty[s] when every warp has released it.
template <int STAGES, int NWARPS>
struct BulkRing {
    uint32_t bar0;   // shared address of full[STAGES] | empty[STAGES]
    __device__ __forceinline__ uint32_t full(int s) const { return bar0 + 8u * s; }
    __device__ __forceinline__ uint32_t empty(int s) const { return bar0 + 8u * (STAGES + s); }
    __device__ __forceinline__ void init() const {   // one thread; followed by a block barrier before anybody waits
        for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), NWARPS); }
        mbar_init_fence();
    }
    __device__ __forceinline__ void load(int c, uint32_t dst, const void* src, uint32_t bytes) const {
        mbar_expect_tx(full(c % STAGES), bytes);
        bulk_g2s(dst, src, bytes, full(c % STAGES));
    }
    // before refilling the stage chunk c - 1 used (c >= 1)
    __device__ __forceinline__ void wait_released(int c_prev) const { mbar_wait(empty(c_prev % STAGES), (uint32_t)((c_prev / STAGES) & 1)); }
    __device__ __forceinline__ void wait_filled(int c) const { mbar_wait(full(c % STAGES), (uint32_t)((c / STAGES) & 1)); }
    __device__ __forceinline__ void release(int c) const { mbar_arrive(empty(c % STAGES)); }   // one lane per warp
};

// ---- geometry of one (P, tile shape) instance -------------------------------------------------------------------
// VAR selects how the W records reach shared memory:
//   0  ring of two 8-particle stages filled with 8-byte cp.async by all threads, one block barrier per stage (small ring:
//      the 73 KB accumulator of P = 6 keeps three blocks per SM);
//   1  two large stages (half a tile each) filled by ONE bulk copy each (cp.async.bulk + mbarrier): the copy engine is
//      efficient for multi-KB copies, and a wait per ~60 particles instead of per 8; two blocks per SM for P = 6;
//   2  as 0 with 32-particle stages (same occupancy as 1; measurement control).
template <int P, int TX, int TY, int TZ, int VAR = 1> struct Spread2Cfg {
    static constexpr int PP = P * P, PPP = PP * P;
    static constexpr int NT = ((PPP + 31) / 32) * 32;      // threads (thread t < PPP <-> residue class (a, b, c))
    static constexpr int NW = NT / 32;
    static constexpr int HX = TX + P - 1, HY = TY + P - 1, HZ = TZ + P - 1;  // window of nodes a tile's particles reach
    static constexpr int CX = (HX + P - 1) / P, CY = (HY + P - 1) / P, CZ = (HZ + P - 1) / P;  // residue cells per axis
    static constexpr int CS = NT;                          // words per cell: multiple of 32 -> bank = thread index
    static constexpr int ACC = CX * CY * CZ * CS;          // accumulator words per component
    static constexpr int RS = wrec_stride(P);              // floats per W record
    static constexpr int ACC_BYTES = 3 * ACC * 4;
    static constexpr bool BIG = ACC_BYTES > 100 * 1024;    // one block per SM anyway: a generous ring
    static constexpr bool BULK = VAR == 1;
    static constexpr int STAGES = 2;
    // particles per stage: VAR 1 sizes a stage for about half of a typical tile (density ~0.075 per node at phi = 0.3)
    static constexpr int CHUNK = VAR == 0 ? 8 : (VAR == 2 ? 32 : (BIG ? 64 : (TX * TY * TZ * 3 / 64 + 7) / 8 * 8));
    static constexpr int STAGE_BYTES = CHUNK * RS * 4;
    static constexpr size_t SMEM = (size_t)ACC_BYTES + (size_t)STAGES * STAGE_BYTES;
    static_assert(CX * CY * CZ * CS < (1 << 24), "accumulator offset");
    static_assert(HZ <= 32, "flush handles at most eight z quads per row");
};

template <int P, int TX, int TY, int TZ, int VAR>
__global__ void __launch_bounds__(Spread2Cfg<P, TX, TY, TZ, VAR>::NT)
spread2_kernel(const float* __restrict__ wrecs, const uint32_t* __restrict__ wcell_start, WaveParams wp, TileGrid tg,
               float* __restrict__ grid) {
    using C = Spread2Cfg<P, TX, TY, TZ, VAR>;
    constexpr bool BULK = C::BULK;
    constexpr int PP = C::PP, PPP = C::PPP, NT = C::NT, NW = C::NW, RS = C::RS, CHUNK = C::CHUNK, STAGES = C::STAGES;
    constexpr int ACC = C::ACC, CS = C::CS;
    constexpr int SXW = C::CY * C::CZ * CS, SYW = C::CZ * CS, SZW = CS;  // cell strides in words
    extern __shared__ __align__(128) float smem_v2[];
    float* acc = smem_v2;                                                      // [3][ACC]
    unsigned char* stage0 = reinterpret_cast<unsigned char*>(acc + 3 * ACC);  // [STAGES][CHUNK records]
    __shared__ __align__(8) unsigned long long s_bar[2 * STAGES];

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x + tg.tile0;
    const uint32_t cb = __ldg(wcell_start + tile), ce = __ldg(wcell_start + tile + 1);
    if (cb == ce) return;  // nothing spreads from this tile (the grid was zero-filled)
    const int bz = tile % tg.ntz, by = (tile / tg.ntz) % tg.nty, bx = tile / (tg.ntz * tg.nty);
    const int t0x = bx * TX, t0y = by * TY, t0z = bz * TZ;
    const uint32_t np = ce - cb;
    const int nchunks = (int)((np + CHUNK - 1) / CHUNK);

    BulkRing<STAGES, NW> ring;
    ring.bar0 = smem_u32(s_bar);
    auto issue_chunk = [&](int c) {   // BULK: producer thread only; otherwise all threads
        const uint32_t w0 = cb + (uint32_t)c * CHUNK;
        const uint32_t n = min((uint32_t)CHUNK, ce - w0);
        unsigned char* dst = stage0 + (size_t)(c % STAGES) * C::STAGE_BYTES;
        const float* src = wrecs + (size_t)w0 * RS;
        if (BULK) {
            ring.load(c, smem_u32(dst), src, n * RS * 4u);
        } else {
            for (int t = tid; t < (int)n * (RS / 2); t += NT) cp_async8(dst + 8 * t, reinterpret_cast<const unsigned char*>(src) + 8 * t);
            cp_async_commit();
        }
    };
    // the first chunk(s) are in flight while the accumulator is cleared
    if (BULK) {
        if (tid == 0) {
            ring.init();
            for (int c = 0; c < STAGES - 1 && c < nchunks; ++c) issue_chunk(c);
        }
    } else {
        issue_chunk(0);
    }
    if (!(tg.dbg & 4))
    for (int i = tid; i < 3 * ACC / 4; i += NT) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // residue class of this thread
    const int a = tid / PP, b = (tid - a * PP) / P, c3 = tid - a * PP - b * P;
    const bool worker = tid < PPP && !(tg.dbg & 1);
    const uint32_t my_acc = smem_u32(acc) + 4u * (uint32_t)tid;

    for (int c = 0; c < nchunks; ++c) {
        if (BULK) {
            // refill the stage chunk c - 1 used (every warp releases a stage when it is past that chunk)
            if (tid == 0 && c + STAGES - 1 < nchunks) {
                if (c >= 1) ring.wait_released(c - 1);
                issue_chunk(c + STAGES - 1);   // (at c = 0 the last stage of the ring is still fresh)
            }
            __syncwarp();
            ring.wait_filled(c);
        } else {
            cp_async_wait<0>();
            __syncthreads();                       // chunk c visible to all; everyone is done with chunk c - 1
            if (c + 1 < nchunks) issue_chunk(c + 1);
        }
        const float* st = reinterpret_cast<const float*>(stage0 + (size_t)(c % STAGES) * C::STAGE_BYTES);
        const int n = (int)min((uint32_t)CHUNK, np - (uint32_t)c * CHUNK);
        if (worker) {
#pragma unroll 2
            for (int q = 0; q < n; ++q) {
                const float* rec = st + q * RS;
                const float4 f = *reinterpret_cast<const float4*>(rec);
                const int4 od = *reinterpret_cast<const int4*>(rec + 4);  // (accumulator offset of the origin's cell, origin residues)
                // support index of my node along each axis: (residue - origin residue) mod P; a borrow means the node lies
                // in the next accumulator cell
                const int dx = a - od.y, dy = b - od.z, dz = c3 - od.w;
                const int mx = dx >> 31, my = dy >> 31, mz = dz >> 31;  // 0 or -1
                const int ij = (dx - mx * P) * P + (dy - my * P), k = dz - mz * P;
                const int cell = od.x - mx * SXW - my * SYW - mz * SZW;
                const float w = rec[WREC_HDR + ij] * rec[WREC_HDR + PP + k];
                const uint32_t addr = my_acc + 4u * (uint32_t)cell;
                smem_fma<0>(addr, w, f.x);
                smem_fma<4 * ACC>(addr, w, f.y);
                smem_fma<8 * ACC>(addr, w, f.z);
            }
        }
        if (BULK) {
            __syncwarp();
            if (lane == 0) ring.release(c);
        }
    }
    __syncthreads();

    // ---- merge the window into the grid: one vector reduction per four z nodes and component.  Eight lanes share a
    // (x, y) row of the window, lane q of them takes z nodes 4q .. 4q + 3 (index arithmetic of the z part hoisted)
    const size_t G = (size_t)wp.nxa * wp.Ny * wp.Nz;   // component stride of the (local) grid buffer
    int relx = t0x - wp.xorg; if (relx < 0) relx += wp.Nx;
    constexpr int HX = C::HX, HY = C::HY, HZ = C::HZ, NQ = (HZ + 3) / 4;
    const bool vec = ((wp.Nz & 3) == 0) && ((TZ & 3) == 0);
    const int q = tid & 7;
    if (q < NQ && !(tg.dbg & 2)) {
        int zoff[4];
        bool zin[4];
#pragma unroll
        for (int z = 0; z < 4; ++z) {
            const int lz = 4 * q + z;
            zin[z] = lz < HZ;
            zoff[z] = zin[z] ? (lz / P) * CS + lz % P : 0;
        }
        int gz = t0z + 4 * q; if (gz >= wp.Nz) gz -= wp.Nz;
        for (int row = tid >> 3; row < HX * HY; row += NT / 8) {
            const int lx = row / HY, ly = row - lx * HY;
            const float* ar = acc + ((lx / P) * C::CY + ly / P) * C::CZ * CS + ((lx % P) * P + ly % P) * P;
            float v[3][4];
            bool any = false;
#pragma unroll
            for (int z = 0; z < 4; ++z)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) { v[cc][z] = zin[z] ? ar[cc * ACC + zoff[z]] : 0.f; any |= v[cc][z] != 0.f; }
            if (!any) continue;
            int gx = relx + lx; if (gx >= wp.nxw) gx -= wp.Nx;
            if (gx >= wp.nxa) continue;   // (slab buffer: only ever rows nobody spread into, or an undersized halo)
            int gy = t0y + ly; if (gy >= wp.Ny) gy -= wp.Ny;
            float* dst = grid + ((size_t)gx * wp.Ny + gy) * wp.Nz;
            if (vec) {  // aligned quads never straddle the periodic boundary
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) red_add_v4(dst + cc * G + gz, v[cc][0], v[cc][1], v[cc][2], v[cc][3]);
            } else {
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    if (!zin[z]) break;
                    int zz = gz + z; if (zz >= wp.Nz) zz -= wp.Nz;
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)
                        if (v[cc][z] != 0.f) red_add(dst + cc * G + zz, v[cc][z]);
                }
            }
        }
    }
}

// ---- interpolation on box-shaped tiles ----------------------------------------------------------------------------
// The block stages the (T + P - 1)^3 window of the three velocity grids with 4-byte asynchronous copies (no register
// staging: every load of the window is in flight at once) and streams the W records of its particles through a bulk-copy
// ring, one particle per warp and chunk.  A lane owns one (i, j) COLUMN of the support and walks its P z-nodes (tile strides
// searched at compile time so that the 32 columns of a pass hit 32 banks); w_xy is the lane's own word of the factor row,
// w_z comes by shuffle.
__host__ __device__ constexpr int interp2_pad(int P, int HY, int HZ) {
    const int PP = P * P;
    int best_pad = 0, best_cost = 1 << 30;
    for (int pad = 0; pad < 32; ++pad) {
        const int XS = HY * HZ + pad;
        int cost = 0;
        for (int c0 = 0; c0 < PP; c0 += 32) {
            int cnt[32] = {};
            int mx = 0;
            for (int c = c0; c < PP && c < c0 + 32; ++c) {
                const int i = c / P, j = c % P;
                const int b = (i * XS + j * HZ) % 32;
                if (++cnt[b] > mx) mx = cnt[b];
            }
            cost += mx;
        }
        if (cost < best_cost) { best_cost = cost; best_pad = pad; }
    }
    return best_pad;
}
__device__ __forceinline__ void cp_async4s(uint32_t smem_addr, const void* gmem_src) {   // shared-window address
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
template <int P, int TX, int TY, int TZ> struct Interp2Cfg {
    static constexpr int HX = TX + P - 1, HY = TY + P - 1, HZ = TZ + P - 1;
    static constexpr int XS = HY * HZ + interp2_pad(P, HY, HZ), GT = HX * XS;
    static constexpr int WIN_BYTES = ((3 * GT * 4 + 127) / 128) * 128;
    static constexpr bool BIG = WIN_BYTES > 100 * 1024;
    static constexpr int THREADS = BIG ? 512 : 256;        // small windows: three blocks per SM
    static constexpr int NW = THREADS / 32;
    static constexpr int MINB = BIG ? 1 : (WIN_BYTES > 68 * 1024 ? 2 : 3);
    static constexpr int RS = wrec_stride(P);
    static constexpr int STAGES = 4;                       // one particle per warp and stage
    static constexpr int STAGE_BYTES = NW * RS * 4;
    static constexpr size_t SMEM = (size_t)WIN_BYTES + (size_t)STAGES * STAGE_BYTES;
};

template <int P, int TX, int TY, int TZ>
__global__ void __launch_bounds__(Interp2Cfg<P, TX, TY, TZ>::THREADS, Interp2Cfg<P, TX, TY, TZ>::MINB)
interp2_kernel(const float* __restrict__ wrecs, const uint32_t* __restrict__ wcell_start, WaveParams wp, TileGrid tg,
               const float* __restrict__ grid, float4* __restrict__ U, int accumulate) {
    using C = Interp2Cfg<P, TX, TY, TZ>;
    extern __shared__ __align__(128) float smem_v2[];
    constexpr int PP = P * P, NW = C::NW, RS = C::RS, NWD = PP + P, STAGES = C::STAGES;
    constexpr int NWR = (NWD + 31) / 32;  // factor words per lane
    constexpr int HX = C::HX, HY = C::HY, HZ = C::HZ, XS = C::XS, GT = C::GT;
    constexpr int NFULL = PP / 32, LEFT = PP - 32 * NFULL;
    constexpr bool LEFT_NODES = LEFT > 0 && LEFT * P <= 32;
    float* g = smem_v2;
    unsigned char* stage0 = reinterpret_cast<unsigned char*>(smem_v2) + C::WIN_BYTES;
    __shared__ __align__(8) unsigned long long s_bar[2 * STAGES];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t cell = blockIdx.x + tg.tile0;
    const uint32_t cb = __ldg(wcell_start + cell), ce = __ldg(wcell_start + cell + 1);
    if (cb == ce) return;
    const int bz = cell % tg.ntz, by = (cell / tg.ntz) % tg.nty, bx = cell / (tg.ntz * tg.nty);
    const int t0x = bx * TX, t0y = by * TY, t0z = bz * TZ;
    const size_t G = (size_t)wp.nxa * wp.Ny * wp.Nz;   // component stride of the (local) grid buffer
    int relx = t0x - wp.xorg; if (relx < 0) relx += wp.Nx;
    const uint32_t np = ce - cb;
    const int nchunks = (int)((np + NW - 1) / NW);
    BulkRing<STAGES, NW> ring;
    ring.bar0 = smem_u32(s_bar);
    auto issue_chunk = [&](int c) {
        const uint32_t w0 = cb + (uint32_t)c * NW;
        const uint32_t n = min((uint32_t)NW, ce - w0);
        ring.load(c, smem_u32(stage0 + (size_t)(c % STAGES) * C::STAGE_BYTES), wrecs + (size_t)w0 * RS, n * RS * 4u);
    };
    if (tid == 0) {
        ring.init();
        for (int c = 0; c < STAGES - 1 && c < nchunks; ++c) issue_chunk(c);
    }
    // stage the window (periodic wrap per node): a warp takes whole x planes, lanes run over (y, z) with z fastest.  The
    // (y, z) part of a node's address is the same in every plane: each lane works out its NE offsets once, the plane loop
    // is then one address add and three 4-byte asynchronous copies per node (the index arithmetic was a third of the
    // kernel's instructions)
    constexpr int NE = (HY * HZ + 31) / 32;
    int poff[NE];
#pragma unroll
    for (int t = 0; t < NE; ++t) {
        const int e = lane + 32 * t;
        const int ly = e / HZ, lz = e - ly * HZ;
        int y = t0y + ly; if (y >= wp.Ny) y -= wp.Ny;
        int z = t0z + lz; if (z >= wp.Nz) z -= wp.Nz;
        poff[t] = e < HY * HZ ? y * wp.Nz + z : -1;
    }
    const uint32_t g_sh = smem_u32(g) + 4u * (uint32_t)lane;
    for (int lx = wid; lx < HX; lx += NW) {
        int x = relx + lx; if (x >= wp.nxw) x -= wp.Nx;
        if (x >= wp.nxa) continue;   // (slab buffer: window planes no particle of this rank reaches)
        const float* gx = grid + (size_t)x * wp.Ny * wp.Nz;
        const uint32_t sx = g_sh + 4u * (uint32_t)(lx * XS);
#pragma unroll
        for (int t = 0; t < NE; ++t) {
            if (poff[t] < 0) continue;
            const float* src = gx + poff[t];
            const uint32_t dst = sx + 128u * t;   // (node = lx * XS + ly * HZ + lz = lx * XS + e)
            cp_async4s(dst, src);
            cp_async4s(dst + 4u * GT, src + G);
            cp_async4s(dst + 8u * GT, src + 2 * G);
        }
    }
    cp_async_commit();
    int col_off[NFULL > 0 ? NFULL : 1];
#pragma unroll
    for (int r = 0; r < NFULL; ++r) {
        const int c = lane + 32 * r;
        col_off[r] = (c / P) * XS + (c % P) * HZ;
    }
    int left_off = -1, left_col = 0, left_k = 0;
    if (LEFT > 0) {
        if (LEFT_NODES) {
            if (lane < LEFT * P) { left_col = 32 * NFULL + lane / P; left_k = lane % P; }
            else left_col = -1;
        } else {
            left_col = lane < LEFT ? 32 * NFULL + lane : -1;
        }
        if (left_col >= 0) left_off = (left_col / P) * XS + (left_col % P) * HZ + left_k;
    }
    cp_async_wait<0>();
    __syncthreads();   // window staged, barriers initialised
    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0 && c + STAGES - 1 < nchunks) {
            if (c >= 1) ring.wait_released(c - 1);
            issue_chunk(c + STAGES - 1);
        }
        __syncwarp();
        ring.wait_filled(c);
        const uint32_t pidx = (uint32_t)c * NW + wid;   // this warp's particle of the chunk
        if (pidx < np) {
            const float* rec = reinterpret_cast<const float*>(stage0 + (size_t)(c % STAGES) * C::STAGE_BYTES) + wid * RS;
            const uint32_t id = __float_as_uint(rec[3]);
            const int4 o = *reinterpret_cast<const int4*>(rec + 8);   // support origin inside the tile
            float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane == 0) old = U[id];   // .w (mass in the reference's velocity array, PSEv1/Mobility.cu:474) is preserved
            float wt[NWR];
#pragma unroll
            for (int r = 0; r < NWR; ++r) wt[r] = lane + 32 * r < NWD ? rec[WREC_HDR + lane + 32 * r] : 0.f;
            float wz[P];
#pragma unroll
            for (int k = 0; k < P; ++k) wz[k] = __shfl_sync(0xffffffffu, wt[(PP + k) / 32], (PP + k) % 32);
            const float* gb = g + o.x * XS + o.y * HZ + o.z;
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
        }
        __syncwarp();
        if (lane == 0) ring.release(c);
    }
}

// ---- tile shape per support size --------------------------------------------------------------------------------
// P = 6 (error 1e-3): 12^3 tiles -> 17^3 window, 3^3 residue cells, 73 KB accumulator, three blocks per SM
//                     (measured at N = 1M: 12^3 555 us vs 16^3 808 us, profiles/r2_notes.md).
// P = 7:              (15, 15, 12)  -> 3^3 cells, 114 KB.
// P = 8 (error 1e-4): (17, 17, 16)  -> (24, 24, 23) window, 3^3 cells, 166 KB, one 512-thread block per SM.
// Other P keep the round-1 kernels (TILE^3 tiles).  z extents are multiples of 4 so that window rows start 16-byte aligned.
#define PSE_V2_SHAPES(X) X(6, 12, 12, 12) X(7, 15, 15, 12) X(8, 17, 17, 16)

static bool v2_shape(int P, int* tx, int* ty, int* tz) {
#define X(p, a, b, c) if (P == p) { *tx = a; *ty = b; *tz = c; return true; }
    PSE_V2_SHAPES(X)
#undef X
    return false;
}
static void v2_fill_tilegrid(int P, int tx, int ty, int tz, TileGrid* tg) {
    tg->tx = tx; tg->ty = ty; tg->tz = tz;
    tg->cp = P;
    tg->cy = (ty + P - 1 + P - 1) / P; tg->cz = (tz + P - 1 + P - 1) / P;
    tg->cs = ((P * P * P + 31) / 32) * 32;
    tg->rs = wrec_stride(P);
}
static cudaError_t v2_set_attributes(int P, int tx, int ty, int tz) {
    cudaError_t err = cudaErrorInvalidValue;
#define X(p, a, b, c)                                                                                                          \
    if (P == p && tx == a && ty == b && tz == c) {                                                                             \
        err = cudaFuncSetAttribute(spread2_kernel<p, a, b, c, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Spread2Cfg<p, a, b, c, 0>::SMEM); \
        if (err == cudaSuccess) err = cudaFuncSetAttribute(spread2_kernel<p, a, b, c, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Spread2Cfg<p, a, b, c, 1>::SMEM); \
        if (err == cudaSuccess) err = cudaFuncSetAttribute(spread2_kernel<p, a, b, c, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Spread2Cfg<p, a, b, c, 2>::SMEM); \
        if (err == cudaSuccess) err = cudaFuncSetAttribute(interp2_kernel<p, a, b, c>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Interp2Cfg<p, a, b, c>::SMEM); \
    }
    PSE_V2_SHAPES(X)
#undef X
    return err;
}
static void launch_spread2(int P, int var, cudaStream_t st, const float* wrecs, const uint32_t* wstart, const WaveParams& wp, const TileGrid& tg,
                           float* grid, int ntiles = -1) {
    if (ntiles < 0) ntiles = tg.ntile;
    if (ntiles == 0) return;
#define X(p, a, b, c)                                                                                                          \
    if (P == p && tg.tx == a && tg.ty == b && tg.tz == c) {                                                                    \
        if (var == 0) spread2_kernel<p, a, b, c, 0><<<ntiles, Spread2Cfg<p, a, b, c, 0>::NT, Spread2Cfg<p, a, b, c, 0>::SMEM, st>>>(wrecs, wstart, wp, tg, grid); \
        else if (var == 2) spread2_kernel<p, a, b, c, 2><<<ntiles, Spread2Cfg<p, a, b, c, 2>::NT, Spread2Cfg<p, a, b, c, 2>::SMEM, st>>>(wrecs, wstart, wp, tg, grid); \
        else spread2_kernel<p, a, b, c, 1><<<ntiles, Spread2Cfg<p, a, b, c, 1>::NT, Spread2Cfg<p, a, b, c, 1>::SMEM, st>>>(wrecs, wstart, wp, tg, grid); \
        return;                                                                                                                \
    }
    PSE_V2_SHAPES(X)
#undef X
}
static void launch_interp2(int P, cudaStream_t st, const float* wrecs, const uint32_t* wstart, const WaveParams& wp, const TileGrid& tg,
                           const float* grid, float4* U, int accumulate, int ntiles = -1) {
    if (ntiles < 0) ntiles = tg.ntile;
    if (ntiles == 0) return;
#define X(p, a, b, c)                                                                                                          \
    if (P == p && tg.tx == a && tg.ty == b && tg.tz == c) {                                                                    \
        interp2_kernel<p, a, b, c><<<ntiles, Interp2Cfg<p, a, b, c>::THREADS, Interp2Cfg<p, a, b, c>::SMEM, st>>>(wrecs, wstart, wp, tg, grid, U, accumulate); \
        return;                                                                                                                \
    }
    PSE_V2_SHAPES(X)
#undef X
}

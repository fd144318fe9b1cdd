// Peer-memory transport of the slab-decomposed step: every exchange is ONE kernel that reads the neighbours' buffers
// directly over NVLink (pointers obtained with cudaIpcOpenMemHandle, or plain pointers when the ranks are engines of one
// process), preceded by a device-side barrier on flags in peer memory.  No packing, no staging buffers, no host
// involvement: an exchange costs the barrier (a few microseconds) plus the NVLink transfer itself, where the same
// exchange through ncclSend/ncclRecv costs two pack/unpack passes over HBM and a collective launch (measured at N = 1M
// on 2 GPUs: 1.3 ms of a 3.5 ms step in collectives, profiles/r2_multi_gpu.md).
// The reference is single-GPU (PSEv1/Stokes.cc:104): everything here is new work.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PEER_MAX 16
// one per rank, in that rank's memory:  flags[2 channels][PEER_MAX] (arrival epochs, written by the peers) | red[2][PEER_MAX][4]
// Two channels: the wave-space and the real-space branch of a step run on two streams and synchronise independently.
#define PEER_PAD_RED_OFF 256
#define PEER_PAD_BYTES (PEER_PAD_RED_OFF + 2 * PEER_MAX * 4 * 8)
struct PeerSync {
    int rank, world;
    int chan;                       // synchronisation channel (0 wave space / velocities, 1 real space)
    unsigned char* pad[PEER_MAX];   // pad[q]: rank q's pad as addressable from this rank
    uint32_t* err;                  // own error word (read by the host with the displacement check)
};
template <class T> struct PeerPtrs { T* p[PEER_MAX]; };
struct PeerBounds { uint32_t row[PEER_MAX + 1]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// arrival of every rank at synchronisation point `epoch`: thread q tells rank q, then waits for rank q.
// A rank that never arrives (host-side error on that rank) must not hang the GPU: after ~10 s the wait gives up and
// raises the error word, which the host reads with the next displacement check.
__device__ __forceinline__ void peer_arrive_and_wait(const PeerSync& ps, uint32_t epoch) {
    const int q = threadIdx.x;
    if (q < ps.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(ps.pad[q]) + ps.chan * PEER_MAX + ps.rank, epoch);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(ps.pad[ps.rank]) + ps.chan * PEER_MAX + q;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > (20ll << 30)) { atomicOr(ps.err, 2u); break; }
            __nanosleep(64);
        }
    }
}
__global__ void peer_barrier_kernel(PeerSync ps, uint32_t epoch) {
    peer_arrive_and_wait(ps, epoch);
}
// sum of three doubles over ranks, in place; every rank adds the contributions in rank order -> identical bits
__global__ void peer_allreduce3_kernel(PeerSync ps, uint32_t epoch, uint32_t parity, double* __restrict__ v) {
    const int q = threadIdx.x;
    if (q < ps.world) {
        double* slot = reinterpret_cast<double*>(ps.pad[q] + PEER_PAD_RED_OFF) + ((size_t)parity * PEER_MAX + ps.rank) * 4;
        slot[0] = v[0]; slot[1] = v[1]; slot[2] = v[2];
    }
    peer_arrive_and_wait(ps, epoch);
    __syncthreads();
    if (threadIdx.x < 3) {
        const double* mine = reinterpret_cast<const double*>(ps.pad[ps.rank] + PEER_PAD_RED_OFF) + (size_t)parity * PEER_MAX * 4;
        double a = 0.0;
        for (int r = 0; r < ps.world; ++r) a += mine[4 * r + threadIdx.x];
        v[threadIdx.x] = a;
    }
}

struct PeerSlabs { int xs[PEER_MAX + 1], ys[PEER_MAX + 1]; };
// Two rows per warp and trip, all loads issued before the first store: a load from a peer takes ~2 us over NVLink, so the
// bytes in flight per SM, not the instruction count, set the rate of these kernels.
__device__ __forceinline__ void peer_copy_rows2(float2* __restrict__ dst0, const float2* __restrict__ src0, float2* __restrict__ dst1,
                                                const float2* __restrict__ src1, int n, int lane) {
    if ((n & 1) == 0) {
        const float4 *a = reinterpret_cast<const float4*>(src0), *b = reinterpret_cast<const float4*>(src1);
        float4 *da = reinterpret_cast<float4*>(dst0), *db = reinterpret_cast<float4*>(dst1);
        const int n4 = n / 2;
        for (int i = lane; i < n4; i += 64) {
            const bool two = i + 32 < n4;
            const float4 v0 = a[i], w0 = src1 ? b[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v1 = two ? a[i + 32] : v0, w1 = (two && src1) ? b[i + 32] : w0;
            da[i] = v0; if (src1) db[i] = w0;
            if (two) { da[i + 32] = v1; if (src1) db[i + 32] = w1; }
        }
    } else {
        for (int i = lane; i < n; i += 32) { dst0[i] = src0[i]; if (src1) dst1[i] = src1[i]; }
    }
}
// transpose x slabs -> y slabs:  tr[c][x][yl][kz] = sloc_{owner(x)}[c][x - xs][y0 + yl][kz]   (one warp per pair of kz rows)
__global__ void __launch_bounds__(256)
peer_pull_trans_kernel(float2* __restrict__ tr, PeerPtrs<const float2> sloc, PeerSlabs b, int Nx, int Ny, int y0, int nyl, int Nzp) {
    const int lane = threadIdx.x & 31, nw = gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nrow = 3 * Nx * nyl;
    auto src_of = [&](int row) {
        const int yl = row % nyl, x = (row / nyl) % Nx, c = row / (nyl * Nx);
        int r = 0;
        while (x >= b.xs[r + 1]) ++r;
        return sloc.p[r] + (((size_t)c * (b.xs[r + 1] - b.xs[r]) + (x - b.xs[r])) * Ny + (y0 + yl)) * Nzp;
    };
    for (int row = 2 * w; row < nrow; row += 2 * nw) {
        const bool two = row + 1 < nrow;
        peer_copy_rows2(tr + (size_t)row * Nzp, src_of(row), tr + (size_t)(row + 1) * Nzp, two ? src_of(row + 1) : nullptr, Nzp, lane);
    }
}
// and back:  sloc[c][xl][y][kz] = tr_{owner(y)}[c][x0 + xl][y - ys][kz]
__global__ void __launch_bounds__(256)
peer_pull_slab_kernel(float2* __restrict__ sloc, PeerPtrs<const float2> tr, PeerSlabs b, int Nx, int Ny, int x0, int nxl, int Nzp) {
    const int lane = threadIdx.x & 31, nw = gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nrow = 3 * nxl * Ny;
    auto src_of = [&](int row) {
        const int y = row % Ny, xl = (row / Ny) % nxl, c = row / (Ny * nxl);
        int q = 0;
        while (y >= b.ys[q + 1]) ++q;
        return tr.p[q] + (((size_t)c * Nx + (x0 + xl)) * (b.ys[q + 1] - b.ys[q]) + (y - b.ys[q])) * Nzp;
    };
    for (int row = 2 * w; row < nrow; row += 2 * nw) {
        const bool two = row + 1 < nrow;
        peer_copy_rows2(sloc + (size_t)row * Nzp, src_of(row), sloc + (size_t)(row + 1) * Nzp, two ? src_of(row + 1) : nullptr, Nzp, lane);
    }
}
// The same two transposes as PUSHES: a rank reads its own rows and stores them into the owners' buffers (posted writes over
// NVLink need no round trip; measured against the pulls in profiles/r2_scaling.md).  The barrier that follows makes the
// stores visible before anybody reads them.
__global__ void __launch_bounds__(256)
peer_push_trans_kernel(const float2* __restrict__ sloc, PeerPtrs<float2> tr, PeerSlabs b, int Nx, int Ny, int x0, int nxl, int Nzp) {
    const int lane = threadIdx.x & 31, nw = gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nrow = 3 * nxl * Ny;
    auto dst_of = [&](int row) {
        const int y = row % Ny, xl = (row / Ny) % nxl, c = row / (Ny * nxl);
        int q = 0;
        while (y >= b.ys[q + 1]) ++q;
        return tr.p[q] + (((size_t)c * Nx + (x0 + xl)) * (b.ys[q + 1] - b.ys[q]) + (y - b.ys[q])) * Nzp;
    };
    for (int row = 2 * w; row < nrow; row += 2 * nw) {
        const bool two = row + 1 < nrow;
        peer_copy_rows2(dst_of(row), sloc + (size_t)row * Nzp, two ? dst_of(row + 1) : nullptr, two ? sloc + (size_t)(row + 1) * Nzp : nullptr, Nzp, lane);
    }
}
__global__ void __launch_bounds__(256)
peer_push_slab_kernel(const float2* __restrict__ tr, PeerPtrs<float2> sloc, PeerSlabs b, int Nx, int Ny, int y0, int nyl, int Nzp) {
    const int lane = threadIdx.x & 31, nw = gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nrow = 3 * Nx * nyl;
    auto dst_of = [&](int row) {
        const int yl = row % nyl, x = (row / nyl) % Nx, c = row / (nyl * Nx);
        int r = 0;
        while (x >= b.xs[r + 1]) ++r;
        return sloc.p[r] + (((size_t)c * (b.xs[r + 1] - b.xs[r]) + (x - b.xs[r])) * Ny + (y0 + yl)) * Nzp;
    };
    for (int row = 2 * w; row < nrow; row += 2 * nw) {
        const bool two = row + 1 < nrow;
        peer_copy_rows2(dst_of(row), tr + (size_t)row * Nzp, two ? dst_of(row + 1) : nullptr, two ? tr + (size_t)(row + 1) * Nzp : nullptr, Nzp, lane);
    }
}
// planes [p_dst, p_dst + n) of my real-space buffer (+)= planes [p_src, ..) of a peer's buffer (component strides differ);
// one launch serves both neighbours: blocks with odd index take job B
struct PeerPlaneJob { const float* peer; size_t Glp; int p_dst, p_src, n; };
__global__ void __launch_bounds__(256)
peer_planes_kernel(float* __restrict__ grid, size_t Gl, size_t plane, PeerPlaneJob ja, PeerPlaneJob jb, int add) {
    const PeerPlaneJob j = (blockIdx.x & 1) ? jb : ja;
    const float* __restrict__ peer = j.peer;
    const size_t stride = (size_t)(gridDim.x >> 1) * blockDim.x, t0 = (size_t)(blockIdx.x >> 1) * blockDim.x + threadIdx.x;
    const int n = j.n;
    if ((plane & 3) == 0) {
        const size_t p4 = plane / 4, tot = (size_t)3 * n * p4;
        for (size_t t = t0; t < tot; t += stride) {
            const size_t k = t % p4;
            const int i = (int)((t / p4) % n), c = (int)(t / (p4 * n));
            float4* d = reinterpret_cast<float4*>(grid + c * Gl + (size_t)(j.p_dst + i) * plane) + k;
            const float4 v = reinterpret_cast<const float4*>(peer + c * j.Glp + (size_t)(j.p_src + i) * plane)[k];
            if (add) { float4 o = *d; o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w; *d = o; }
            else *d = v;
        }
    } else {
        const size_t tot = (size_t)3 * n * plane;
        for (size_t t = t0; t < tot; t += stride) {
            const size_t k = t % plane;
            const int i = (int)((t / plane) % n), c = (int)(t / (plane * n));
            float* d = grid + c * Gl + (size_t)(j.p_dst + i) * plane + k;
            const float v = peer[c * j.Glp + (size_t)(j.p_src + i) * plane + k];
            *d = add ? *d + v : v;
        }
    }
}
// boundary rows of the vector about to be multiplied, straight from the owners' records (16 of every 32 bytes)
__global__ void peer_pull_px_kernel(float4* __restrict__ px /* stride 2 */, const float4* __restrict__ peerA, uint32_t a0, uint32_t na,
                                    const float4* __restrict__ peerB, uint32_t b0, uint32_t nb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) px[2 * (size_t)(a0 + i) + 1] = peerA[2 * (size_t)(a0 + i) + 1];
    else if (i < na + nb) { const uint32_t j = i - na; px[2 * (size_t)(b0 + j) + 1] = peerB[2 * (size_t)(b0 + j) + 1]; }
}
// velocities: every rank reads every row from its owner and scatters it to particle-id order (the caller's .w is kept)
__global__ void peer_gather_scatter_kernel(PeerPtrs<const float4> uslot, PeerBounds rows, int world, const uint32_t* __restrict__ perm,
                                           uint32_t N, float4* __restrict__ U) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    int r = 0;
    while (r + 1 < world && s >= rows.row[r + 1]) ++r;
    const float4 v = uslot.p[r][s];
    const uint32_t p = perm[s];
    float4 o = U[p];
    o.x = v.x; o.y = v.y; o.z = v.z;
    U[p] = o;
}

// Cell binning, spatial ordering and the real-space neighbour list.
//
// Replaces HOOMD's CellListGPU + NeighborListGPUBinned (not in /root/reference; used at
// PSEv1/integrate.py:58-83 and consumed at PSEv1/Stokes.cc:433-438, PSEv1/Mobility.cu:624-641).
// Contract: per-particle neighbour SET == brute-force minimum-image set with
// |r|^2 < r_list^2, stored as a full list (both directions), rows ascending.
//
// Layout in HBM: particles are permuted into cell order ("slots"); cells are boxes of the
// fractional (sheared) coordinate space, raster order with z fastest, so the particles of a run of
// z-adjacent cells are contiguous.  perm[slot] = particle id.  Neighbour list is CSR over slots:
// head[slot], nn[slot], nl[head .. head+nn) ascending slot numbers.
#pragma once
#include "box.cuh"
#include "common.cuh"

struct CellGrid {
    int ncx, ncy, ncz;  // cells per dimension
    int ncell;
    // search reach in fractional units (already includes the shear widening for x)
    float reach_fx, reach_fy, reach_fz;
};

__device__ __forceinline__ int cell_coord(float frac, int nc) {
    // fractional coordinate nominally in [0,1); particles exactly on / slightly outside the
    // boundary are clamped into the edge cells (the search below uses the same function)
    int c = (int)floorf(frac * (float)nc);
    return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

__global__ void cell_id_kernel(const float4* __restrict__ pos, uint32_t N, PseBox box, CellGrid cg,
                               uint32_t* __restrict__ cell_of, uint32_t* __restrict__ cell_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 p = __ldg(pos + i);
    float3 f = box.make_fraction(p.x, p.y, p.z);
    // positions may sit a hair outside the box: wrap the fraction into [0,1)
    f.x -= floorf(f.x); f.y -= floorf(f.y); f.z -= floorf(f.z);
    int cx = cell_coord(f.x, cg.ncx), cy = cell_coord(f.y, cg.ncy), cz = cell_coord(f.z, cg.ncz);
    uint32_t c = ((uint32_t)cx * cg.ncy + cy) * cg.ncz + cz;
    cell_of[i] = c;
    atomicAdd(cell_count + c, 1u);
}

// ---- exclusive scan (three-phase, uint32) ------------------------------------------------
#define SCAN_BLOCK 1024
__global__ void scan_block_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                  uint32_t* __restrict__ block_sums, uint32_t n) {
    __shared__ uint32_t warp_tot[32];
    uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    uint32_t x = v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    uint32_t incl = x + (wid > 0 ? warp_tot[wid - 1] : 0u);
    if (i < n) out[i] = incl - v;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = incl;
}
__global__ void scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_offs, uint32_t n) {
    uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_offs[blockIdx.x];
}

// place particles into their cells (slot order inside a cell is made canonical afterwards)
__global__ void cell_fill_kernel(const uint32_t* __restrict__ cell_of, uint32_t N, const uint32_t* __restrict__ cell_start,
                                 uint32_t* __restrict__ cell_fill, uint32_t* __restrict__ perm, uint32_t begin = 0) {
    uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;  // items [begin, N)
    if (i >= N) return;
    uint32_t c = cell_of[i];
    if (c == 0xffffffffu) return;  // not binned (sharded wave binning leaves out particles of other slabs)
    uint32_t k = atomicAdd(cell_fill + c, 1u);
    perm[cell_start[c] + k] = i;
}

// canonical order: ascending particle id inside each cell (insertion sort; cells hold O(10))
__global__ void cell_sort_kernel(const uint32_t* __restrict__ cell_start, uint32_t ncell, uint32_t* __restrict__ perm) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    uint32_t b = cell_start[c], e = cell_start[c + 1];
    for (uint32_t i = b + 1; i < e; ++i) {
        uint32_t v = perm[i];
        uint32_t j = i;
        while (j > b && perm[j - 1] > v) { perm[j] = perm[j - 1]; --j; }
        perm[j] = v;
    }
}

// canonical order for populous cells (wave-space tiles hold O(300) particles): one block per cell, rank sort.
// Slots are distinct, so rank = number of smaller entries; the list is staged in shared memory when it fits.
#define CELL_SORT_SMEM 4096
__global__ void __launch_bounds__(128)
cell_sort_block_kernel(const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ perm_in, uint32_t* __restrict__ perm_out) {
    __shared__ uint32_t list[CELL_SORT_SMEM];
    const uint32_t b = cell_start[blockIdx.x], e = cell_start[blockIdx.x + 1];
    const uint32_t n = e - b;
    if (n == 0) return;
    const bool staged = n <= CELL_SORT_SMEM;
    if (staged) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) list[i] = perm_in[b + i];
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t v = staged ? list[i] : perm_in[b + i];
        uint32_t rank = 0;
        if (staged) {
            for (uint32_t j = 0; j < n; ++j) rank += list[j] < v;
        } else {
            for (uint32_t j = 0; j < n; ++j) rank += __ldg(perm_in + b + j) < v;
        }
        perm_out[b + rank] = v;
    }
}

__global__ void invert_perm_kernel(const uint32_t* __restrict__ perm, uint32_t N, uint32_t* __restrict__ slot_of) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < N) slot_of[perm[s]] = s;
}

// out[slot] = in[perm[slot]]  (float4 payload)
__global__ void gather4_kernel(const float4* __restrict__ in, const uint32_t* __restrict__ perm, uint32_t N,
                               float4* __restrict__ out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < N) out[s] = __ldg(in + perm[s]);
}
// slot-ordered positions: plain copy for the cell/wave kernels + the .p half of the packed SpMV record
__global__ void gather_pos_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ perm, uint32_t N,
                                  float4* __restrict__ spos, float4* __restrict__ px /* stride 2 float4 */) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < N) {
        const float4 v = __ldg(pos + perm[s]);
        spos[s] = v;
        px[2 * (size_t)s] = v;
    }
}
// slot-ordered vector: plain copy (wave space) + the .x half of the packed SpMV record
__global__ void gather_vec_kernel(const float4* __restrict__ F, const uint32_t* __restrict__ perm, uint32_t N,
                                  float4* __restrict__ sF, float4* __restrict__ px) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < N) {
        const float4 v = __ldg(F + perm[s]);
        if (sF) sF[s] = v;
        px[2 * (size_t)s + 1] = v;
    }
}
// ---- neighbour search ---------------------------------------------------------------------
// Sorted enumeration of a periodic cell range: cells {lo, lo+1, ..., lo+len-1} mod nc, visited
// in ascending cell number so that rows come out ascending in slot number.
struct CellRange {
    int lo, len, wrapped;  // wrapped = how many of the len cells wrap past nc
    __device__ __forceinline__ int at(int t) const { return t < wrapped ? t : lo + (t - wrapped); }
};
__device__ __forceinline__ CellRange make_range(float f, float reach, int nc) {
    // cells overlapped by [f - reach, f + reach] in fractional units
    int c0 = (int)floorf((f - reach) * (float)nc);
    int c1 = (int)floorf((f + reach) * (float)nc);
    int len = c1 - c0 + 1;
    if (len > nc) len = nc;
    int lo = c0 % nc;
    if (lo < 0) lo += nc;
    CellRange r;
    r.lo = lo; r.len = len;
    r.wrapped = lo + len > nc ? lo + len - nc : 0;
    return r;
}

// One search pass, one thread per slot: rows are first written at ell[slot * cap ...] (fixed stride);
// nn[slot] is the true count even when it exceeds cap (the host then grows cap and rebuilds).  A scan of nn and
// compact_rows_kernel then pack the rows into CSR, which the SpMV reads ~1.6x faster than the strided rows
// (adjacent rows share cache lines).
__global__ void nlist_kernel(const float4* __restrict__ spos, uint32_t N, PseBox box, CellGrid cg,
                             const uint32_t* __restrict__ cell_start, float rlist_sq, uint32_t cap, uint32_t* __restrict__ nn,
                             uint32_t* __restrict__ nl, uint32_t* __restrict__ max_nn, uint32_t row_begin = 0) {
    // rows [row_begin, N) (a slab-decomposed rank searches for its own rows only; scratch rows are relative to row_begin)
    uint32_t i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t count = 0;
    if (i < N) {
        float4 pi = __ldg(spos + i);
        float3 f = box.make_fraction(pi.x, pi.y, pi.z);
        f.x -= floorf(f.x); f.y -= floorf(f.y); f.z -= floorf(f.z);
        // a fraction that rounds to exactly 1.0 was clamped into the last cell by cell_id_kernel
        CellRange rx = make_range(f.x, cg.reach_fx, cg.ncx);
        CellRange ry = make_range(f.y, cg.reach_fy, cg.ncy);
        CellRange rz = make_range(f.z, cg.reach_fz, cg.ncz);
        uint32_t* __restrict__ row = nl + (size_t)(i - row_begin) * cap;
        for (int tx = 0; tx < rx.len; ++tx) {
            int cx = rx.at(tx);
            for (int ty = 0; ty < ry.len; ++ty) {
                int cy = ry.at(ty);
                uint32_t rowbase = ((uint32_t)cx * cg.ncy + cy) * cg.ncz;
                // z cells: at most two contiguous runs [0, wrapped) and [lo, lo + len - wrapped)
                for (int part = 0; part < 2; ++part) {
                    int z0 = part == 0 ? 0 : rz.lo;
                    int zn = part == 0 ? rz.wrapped : rz.len - rz.wrapped;
                    if (zn <= 0) continue;
                    uint32_t b = __ldg(cell_start + rowbase + z0), e = __ldg(cell_start + rowbase + z0 + zn);
                    for (uint32_t j = b; j < e; ++j) {
                        if (j == i) continue;
                        float4 pj = __ldg(spos + j);
                        float3 d = box.min_image_fast(make_float3(PSE_SUB(pi.x, pj.x), PSE_SUB(pi.y, pj.y), PSE_SUB(pi.z, pj.z)));
                        float r2 = pse_norm2_rn(d);
                        if (r2 < rlist_sq) {
                            if (count < cap) row[count] = j;
                            ++count;
                        }
                    }
                }
            }
        }
        nn[i] = count;
    }
    uint32_t m = count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > __ldcg(max_nn)) atomicMax(max_nn, m);   // (look first: one word for the whole grid)
}

// sum of nn -> *total (uint64), for statistics only
__global__ void nnz_kernel(const uint32_t* __restrict__ nn, uint32_t N, unsigned long long* __restrict__ total) {
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) s += nn[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

// ell[slot * cap + k] -> nl[head[slot] + k], 8 lanes per row
__global__ void compact_rows_kernel(const uint32_t* __restrict__ ell, uint32_t cap, const uint32_t* __restrict__ nn,
                                    const uint32_t* __restrict__ head, uint32_t N, uint32_t* __restrict__ nl, uint32_t row_begin = 0) {
    const uint32_t row = row_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 3);
    const uint32_t sub = threadIdx.x & 7;
    if (row >= N) return;
    const uint32_t n = nn[row], h = head[row];
    const uint32_t* __restrict__ src = ell + (size_t)(row - row_begin) * cap;
    for (uint32_t k = sub; k < n; k += 8) nl[h + k] = __ldg(src + k);
}

// One pass in slot order at the head of every call: gathers the caller's positions into the slot-ordered arrays (plain copy
// + the .p half of the SpMV records) and, on the way,
//   flags[0]  largest squared displacement since the list was built (staleness test; atomicMax on the bits);
//   flags[2]  != 0 when some particle is not bit for bit where the previous call had it - when nothing moved, the
//             position-only work of the previous call (pruned list, wave-space binning, Gaussian factor rows) is still
//             valid and is not repeated (the operator applied again at a fixed configuration).
__global__ void check_and_gather_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ perm, uint32_t N, PseBox box,
                                        const float4* __restrict__ spos_build, float4* __restrict__ spos, float4* __restrict__ px /* stride 2 */,
                                        uint32_t* __restrict__ flags) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    float r2 = 0.f;
    bool moved = false;
    if (s < N) {
        const float4 a = __ldg(pos + __ldg(perm + s)), b = __ldg(spos_build + s), l = spos[s];
        const float3 d = box.min_image(make_float3(a.x - b.x, a.y - b.y, a.z - b.z));
        r2 = d.x * d.x + d.y * d.y + d.z * d.z;
        moved = __float_as_uint(a.x) != __float_as_uint(l.x) || __float_as_uint(a.y) != __float_as_uint(l.y) || __float_as_uint(a.z) != __float_as_uint(l.z);
        if (moved) { spos[s] = a; px[2 * (size_t)s] = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    // one word each for the whole grid: look before touching it (tens of thousands of same-address atomics / stores serialise in L2)
    const bool any_moved = __any_sync(0xffffffffu, moved);
    if ((threadIdx.x & 31) == 0) {
        const uint32_t bits = __float_as_uint(r2);
        if (r2 > 0.f && bits > __ldcg(flags)) atomicMax(flags, bits);
        if (any_moved && __ldcg(flags + 2) == 0u) flags[2] = 1u;
    }
}

// ---- export in the reference layout (particle ids, rows ascending by id) --------------------
__global__ void export_counts_kernel(const uint32_t* __restrict__ nn, const uint32_t* __restrict__ perm, uint32_t N,
                                     uint32_t* __restrict__ nn_by_id) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < N) nn_by_id[perm[s]] = nn[s];
}
__global__ void export_rows_kernel(const uint32_t* __restrict__ nn, const uint32_t* __restrict__ head,
                                   const uint32_t* __restrict__ nl, const uint32_t* __restrict__ perm, uint32_t N,
                                   const uint32_t* __restrict__ head_by_id, uint32_t* __restrict__ out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    uint32_t n = nn[s], h = head[s];
    uint32_t* row = out + head_by_id[perm[s]];
    for (uint32_t k = 0; k < n; ++k) {  // insertion sort by particle id
        uint32_t v = perm[nl[h + k]];
        uint32_t j = k;
        while (j > 0 && row[j - 1] > v) { row[j] = row[j - 1]; --j; }
        row[j] = v;
    }
}

// first slot of every x layer of cells: layer_start[cx] = cell_start[cx * ncy * ncz], cx in [0, ncx]  (slab bounds of the
// multi-GPU decomposition are cut at layer boundaries; slots are x-major, so a range of layers is a contiguous slot range)
__global__ void layer_start_kernel(const uint32_t* __restrict__ cell_start, int ncx, int layer_cells, uint32_t* __restrict__ out) {
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cx <= ncx) out[cx] = cell_start[(size_t)cx * layer_cells];
}

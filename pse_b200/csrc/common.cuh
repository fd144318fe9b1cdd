// Shared device helpers: warp/block reductions and the "last block finishes" deterministic
// grid reduction used for the Lanczos scalars (replaces the reference's two-kernel dot product
// + blocking 4-byte cudaMemcpy, PSEv1/Helper.cu:146-236, PSEv1/Brownian.cu:444-446).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PSE_WARP 32

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the lanes of a group of G consecutive lanes (G power of two <= 32)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of one float per thread; result valid in thread 0. blockDim.x <= 1024.
__device__ __forceinline__ float block_sum(float v, float* smem /* >= 32 floats */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = 0.f;
    if (wid == 0) {
        r = lane < nw ? smem[lane] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// Deterministic grid-wide sum.  Every block calls this with its partial (valid in thread 0).
// The block that arrives last adds all partials in index order (double accumulator) and stores
// the float result to *result; `counter` must be 0 on entry and is reset for the next use.
__device__ __forceinline__ void grid_sum_finish(float block_partial, float* partials, unsigned int* counter,
                                                float* result, float* smem, bool take_sqrt = false) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_partial;
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) acc += (double)__ldcg(partials + i);
        // fixed-order tree over the block
        __shared__ double dsm[32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) dsm[wid] = acc;
        __syncthreads();
        if (wid == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            double r = lane < nw ? dsm[lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (lane == 0) {
                *result = take_sqrt ? sqrtf((float)r) : (float)r;
                *counter = 0u;
                __threadfence();
            }
        }
    }
    (void)smem;
}

// Stand-in for HOOMD's detail::Saru (not in /root/reference).  Two modes:
//  * injected: uniforms in [0,1) are read from device tables set with
//    pse_ref_set_noise_tables(), 3 per particle / 6 per grid node, so the reference kernels
//    and the engine can be driven by identical random vectors (BASELINE.json north_star:
//    "Brownian displacements must match elementwise given identical random vectors");
//  * otherwise the same Philox stream the engine uses (pse_b200/csrc/rng.cuh).
// The reference draws particle noise on (-sqrt3, sqrt3) (PSEv1/Brownian.cu:121-124) and grid
// noise on (-sqrt(3/2), sqrt(3/2)) (:181-189); the upper bound selects the table/domain.
#pragma once
#include "HOOMDMath.h"
#include "../../../pse_b200/csrc/rng.cuh"

static __device__ const float* pse_ref_tab_particle = nullptr;  // [N][3]
static __device__ const float* pse_ref_tab_grid = nullptr;      // [G][6]

namespace hoomd { namespace detail {
class Saru {
    unsigned int idx, key, n;
    uint4 cache; unsigned int cached_block;
public:
    __device__ Saru(unsigned int seed1, unsigned int seed2) : idx(seed1), key(seed2), n(0), cached_block(0xffffffffu) {}
    __device__ float f(float lo, float hi) {
        const bool particle = hi > 1.5f;
        const float* tab = particle ? pse_ref_tab_particle : pse_ref_tab_grid;
        unsigned int k = n++;
        if (tab) return pse_affine(tab[(size_t)idx * (particle ? 3u : 6u) + k], lo, hi);
        unsigned int blk = k >> 2;
        if (blk != cached_block) {
            cache = pse_philox(idx, blk, particle ? PSE_RNG_DOMAIN_PARTICLE : PSE_RNG_DOMAIN_GRID, key);
            cached_block = blk;
        }
        unsigned int lane = k & 3u;
        unsigned int bits = lane == 0 ? cache.x : lane == 1 ? cache.y : lane == 2 ? cache.z : cache.w;
        return pse_uniform(bits, lo, hi);
    }
};
} }

#ifdef PSE_REF_DEFINE_NOISE_SETTER
// compiled into the Brownian.cu translation unit only (the one that draws random numbers)
extern "C" int pse_ref_set_noise_tables(const float* d_particle, const float* d_grid) {
    cudaError_t e = cudaMemcpyToSymbol(pse_ref_tab_particle, &d_particle, sizeof(d_particle));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(pse_ref_tab_grid, &d_grid, sizeof(d_grid));
    return (int)e;
}
#endif

// Counter-based RNG (Philox4x32-10) shared by the engine and by the Saru stand-in the
// reference kernels are compiled against, so that both sides can draw identical uniforms.
// HOOMD's detail::Saru is not in /root/reference (stream parity unpinned);
// what is kept from the reference is the *keying*: stream = (index, timestep + seed)
// (PSEv1/Brownian.cu:117,176) and uniform (not Gaussian) variates (:119-124,181-189).
// Unlike the reference particle and grid streams are distinct (domain word).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef PSE_HD
#define PSE_HD __host__ __device__ __forceinline__
#endif

#define PSE_RNG_DOMAIN_PARTICLE 0u
#define PSE_RNG_DOMAIN_GRID 1u

PSE_HD uint32_t pse_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1)
PSE_HD uint4 pse_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = pse_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = pse_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// engine streams: ctr = (index, block, domain, 0), key = (timestep + seed, 0xB200)
PSE_HD uint4 pse_philox(uint32_t index, uint32_t block, uint32_t domain, uint32_t key0) {
    return pse_philox4x32(index, block, domain, 0u, key0, 0xB200u);
}

// 24-bit uniform in [0,1)
PSE_HD float pse_u01(uint32_t bits) { return (float)(bits >> 8) * (1.0f / 16777216.0f); }

// uniform in [lo, hi) from u in [0,1): the one formula both sides use (explicit fma so the
// rounding cannot differ between translation units)
PSE_HD float pse_affine(float u, float lo, float hi) { return fmaf(hi - lo, u, lo); }
PSE_HD float pse_uniform(uint32_t bits, float lo, float hi) { return pse_affine(pse_u01(bits), lo, hi); }

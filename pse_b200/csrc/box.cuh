// Triclinic-xy periodic box arithmetic shared by the engine kernels and by the
// header shim the reference kernels are compiled against (oracle/ref_shim).
//
// HOOMD-blue 2.3.3 `BoxDim` is NOT in /root/reference (FindHOOMD.cmake:17-47 locates
// an external install), so its float semantics are restated here and this file is the
// definition of record for "bit-exact particle -> grid index" (SURVEY.md §8c).
// Call sites in the reference that fix the required behaviour:
//   makeFraction  PSEv1/Mobility.cu:173, :380     (fraction in [0,1) of the sheared cell)
//   minImage      PSEv1/Mobility.cu:238, :443, :648
//   wrap          PSEv1/Stokes.cu:185
//   getL / getTiltFactorXY   PSEv1/Helper.cu:304-305, PSEv1/Mobility.cu:165,230
//
// Every float operation is written with an explicit round-to-nearest intrinsic so that
// the compiler cannot contract a*b+c differently in two translation units.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#if defined(__CUDA_ARCH__)
#define PSE_MUL(a, b) __fmul_rn((a), (b))
#define PSE_ADD(a, b) __fadd_rn((a), (b))
#define PSE_SUB(a, b) __fsub_rn((a), (b))
#define PSE_RINT(a) rintf(a)
#else
// host: build with -ffp-contract=off (nvcc host pass / g++); volatile blocks fusing
static inline float pse_hmul(float a, float b) { volatile float r = a * b; return r; }
static inline float pse_hadd(float a, float b) { volatile float r = a + b; return r; }
static inline float pse_hsub(float a, float b) { volatile float r = a - b; return r; }
#define PSE_MUL(a, b) pse_hmul((a), (b))
#define PSE_ADD(a, b) pse_hadd((a), (b))
#define PSE_SUB(a, b) pse_hsub((a), (b))
#define PSE_RINT(a) rintf(a)
#endif

#ifndef PSE_HD
#define PSE_HD __host__ __device__ __forceinline__
#endif

struct PseBox {
    float Lx, Ly, Lz;           // edge lengths  (hi - lo)
    float Lxinv, Lyinv, Lzinv;  // 1 / L
    float lox, loy, loz;        // lower corner, -L/2
    float hix, hiy, hiz;        // upper corner, +L/2
    float xy;                   // tilt factor: x is displaced by xy*y

    // fraction of the (sheared) cell, nominally in [0,1)
    PSE_HD float3 make_fraction(float x, float y, float z) const {
        float dx = PSE_SUB(x, lox);
        float dy = PSE_SUB(y, loy);
        float dz = PSE_SUB(z, loz);
        dx = PSE_SUB(dx, PSE_MUL(xy, y));
        return make_float3(PSE_MUL(dx, Lxinv), PSE_MUL(dy, Lyinv), PSE_MUL(dz, Lzinv));
    }

    // minimum-image of a separation vector: z, then y (dragging x by the tilt), then x
    PSE_HD float3 min_image(float3 w) const {
        float img = PSE_RINT(PSE_MUL(w.z, Lzinv));
        w.z = PSE_SUB(w.z, PSE_MUL(Lz, img));
        img = PSE_RINT(PSE_MUL(w.y, Lyinv));
        w.y = PSE_SUB(w.y, PSE_MUL(Ly, img));
        w.x = PSE_SUB(w.x, PSE_MUL(PSE_MUL(Ly, xy), img));
        img = PSE_RINT(PSE_MUL(w.x, Lxinv));
        w.x = PSE_SUB(w.x, PSE_MUL(Lx, img));
        return w;
    }

    // Same value, cheaper on the common path: when every component is below 0.49 L all three image indices round to
    // zero and min_image() returns its argument bit for bit (x is only dragged by a NON-zero y image), so the 13
    // operations above are skipped.  Used where pairs are tested by the ten-million (list build, pruning).
    PSE_HD float3 min_image_fast(float3 w) const {
        if (fabsf(w.x) < 0.98f * hix && fabsf(w.y) < 0.98f * hiy && fabsf(w.z) < 0.98f * hiz) return w;
        return min_image(w);
    }

    // wrap a position back into the primary cell by at most one image per axis
    PSE_HD void wrap(float3& w, int3& img) const {
        float tilt_x = PSE_MUL(xy, w.y);
        if (w.x >= PSE_ADD(hix, tilt_x)) { w.x = PSE_SUB(w.x, Lx); img.x++; }
        else if (w.x < PSE_ADD(lox, tilt_x)) { w.x = PSE_ADD(w.x, Lx); img.x--; }
        if (w.y >= hiy) { w.y = PSE_SUB(w.y, Ly); w.x = PSE_SUB(w.x, PSE_MUL(Ly, xy)); img.y++; }
        else if (w.y < loy) { w.y = PSE_ADD(w.y, Ly); w.x = PSE_ADD(w.x, PSE_MUL(Ly, xy)); img.y--; }
        if (w.z >= hiz) { w.z = PSE_SUB(w.z, Lz); img.z++; }
        else if (w.z < loz) { w.z = PSE_ADD(w.z, Lz); img.z--; }
    }
};

// |d|^2 with one rounding per operation, ((x*x + y*y) + z*z): the neighbour-list membership test,
// reproducible bit for bit by a float32 host restatement
PSE_HD float pse_norm2_rn(float3 d) {
    return PSE_ADD(PSE_ADD(PSE_MUL(d.x, d.x), PSE_MUL(d.y, d.y)), PSE_MUL(d.z, d.z));
}

static inline PseBox pse_make_box(float Lx, float Ly, float Lz, float xy) {
    PseBox b;
    b.hix = Lx / 2.0f; b.hiy = Ly / 2.0f; b.hiz = Lz / 2.0f;
    b.lox = -b.hix; b.loy = -b.hiy; b.loz = -b.hiz;
    b.Lx = b.hix - b.lox; b.Ly = b.hiy - b.loy; b.Lz = b.hiz - b.loz;
    b.Lxinv = 1.0f / b.Lx; b.Lyinv = 1.0f / b.Ly; b.Lzinv = 1.0f / b.Lz;
    b.xy = xy;
    return b;
}

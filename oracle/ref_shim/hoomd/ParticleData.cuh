// Stand-in for HOOMD's ParticleData.cuh: only BoxDim is needed by the reference kernels.
// The arithmetic lives in pse_b200/csrc/box.cuh (definition of record, shared with the
// engine so that particle->grid indices can be compared bit for bit).
#pragma once
#include "HOOMDMath.h"
#include "../../../pse_b200/csrc/box.cuh"

struct BoxDim {
    PseBox b;
    BoxDim() {}
    explicit BoxDim(const PseBox& box) : b(box) {}
    HOSTDEVICE Scalar3 getL() const { return make_scalar3(b.Lx, b.Ly, b.Lz); }
    HOSTDEVICE Scalar getTiltFactorXY() const { return b.xy; }
    HOSTDEVICE Scalar3 makeFraction(const Scalar3& v) const { return b.make_fraction(v.x, v.y, v.z); }
    HOSTDEVICE Scalar3 minImage(const Scalar3& v) const { return b.min_image(v); }
    HOSTDEVICE void wrap(Scalar3& w, int3& img) const { b.wrap(w, img); }
};

// Minimal stand-in for HOOMD-blue's HOOMDMath.h, used ONLY to compile the reference's
// own .cu files unmodified into oracle/_ref/libpse_ref.so (test infrastructure, never
// linked into the product).  Single precision, as the reference requires in practice
// (SURVEY.md Q2).  Provides exactly the vector helpers the reference kernels use.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef HOSTDEVICE
#define HOSTDEVICE __host__ __device__ inline
#endif

typedef float Scalar;
typedef float2 Scalar2;
typedef float3 Scalar3;
typedef float4 Scalar4;

HOSTDEVICE Scalar2 make_scalar2(Scalar x, Scalar y) { return make_float2(x, y); }
HOSTDEVICE Scalar3 make_scalar3(Scalar x, Scalar y, Scalar z) { return make_float3(x, y, z); }
HOSTDEVICE Scalar4 make_scalar4(Scalar x, Scalar y, Scalar z, Scalar w) { return make_float4(x, y, z, w); }

HOSTDEVICE Scalar3 operator+(const Scalar3& a, const Scalar3& b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
HOSTDEVICE Scalar3 operator-(const Scalar3& a, const Scalar3& b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
HOSTDEVICE Scalar3 operator*(const Scalar3& a, const Scalar3& b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
HOSTDEVICE Scalar3 operator/(const Scalar3& a, const Scalar3& b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
HOSTDEVICE Scalar3 operator*(const Scalar3& a, const Scalar& b) { return make_float3(a.x * b, a.y * b, a.z * b); }
HOSTDEVICE Scalar3 operator*(const Scalar& b, const Scalar3& a) { return make_float3(a.x * b, a.y * b, a.z * b); }
HOSTDEVICE Scalar3 operator/(const Scalar3& a, const Scalar& b) { Scalar q = Scalar(1.0) / b; return a * q; }
HOSTDEVICE Scalar3& operator+=(Scalar3& a, const Scalar3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
HOSTDEVICE Scalar3& operator-=(Scalar3& a, const Scalar3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
HOSTDEVICE Scalar dot(const Scalar3& a, const Scalar3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

#define __scalar2int_rd __float2int_rd

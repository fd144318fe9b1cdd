// Texture references were removed in CUDA 12.  HOOMD's texFetchScalar4 already compiled to
// __ldg(ptr + i) on sm_35+, so the stand-in is a plain read-only load and the
// texture objects are inert PODs (the reference only sets .normalized / .filterMode on them).
#pragma once
#include "HOOMDMath.h"
struct scalar4_tex_impl { bool normalized; int filterMode; };
#define scalar4_tex_t static scalar4_tex_impl
#define texFetchScalar4(ptr, tex, i) __ldg((ptr) + (i))

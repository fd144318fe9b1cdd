// One prototype: the reference calls LAPACKE_spteqr only (PSEv1/Brownian.cu:540,673).
// Resolved against scipy's bundled OpenBLAS through -DLAPACKE_spteqr=scipy_LAPACKE_spteqr.
#pragma once
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#ifdef __cplusplus
extern "C" {
#endif
int LAPACKE_spteqr(int matrix_layout, char compz, int n, float* d, float* e, float* z, int ldz);
#ifdef __cplusplus
}
#endif

// TEST INFRASTRUCTURE ONLY.  C entry points that drive the reference's own CUDA host
// drivers and kernels (compiled unmodified from /root/reference/PSEv1/*.cu against
// oracle/ref_shim) on raw device arrays.  Used by tests/, __graft_entry__.smoke() and
// bench.py --impl reference as the checker / reference arm; never linked into the product.
//
// It restates only the array plumbing of Stokes::integrateStepOne (PSEv1/Stokes.cc:429-514):
// launch shapes (PSEv1/Stokes.cu:279-284), identity group index, one cuFFT C2C plan
// (PSEv1/Stokes.cc:257).  All numerics run inside the reference's code.
#include "Stokes.cuh"
#include "Mobility.cuh"
#include "Brownian.cuh"
#include "Helper.cuh"
#include <cufft.h>
#include <stdio.h>

// declared in PSEv1/Brownian.cu:357 but not in Brownian.cuh
void gpu_stokes_BrealLanczos_wrap(Scalar4* d_psi, Scalar4* d_pos, unsigned int* d_group_members,
                                  unsigned int group_size, const BoxDim& box, Scalar dt, Scalar4* d_vel,
                                  const Scalar T, const unsigned int timestep, const unsigned int seed, Scalar xi,
                                  Scalar ewald_cut, Scalar ewald_dr, int ewald_n, Scalar4* d_ewaldC1,
                                  const unsigned int* d_n_neigh, const unsigned int* d_nlist,
                                  const unsigned int* d_headlist, int& m, Scalar cheb_error, dim3 grid, dim3 threads,
                                  int gridBlockSize, int gridNBlock, Scalar3 gridh, Scalar self);

void gpu_stokes_Mwave_wrap(Scalar4* d_pos, Scalar4* d_vel, Scalar4* d_net_force, unsigned int* d_group_members,
                           unsigned int group_size, const BoxDim& box, Scalar xi, Scalar eta, Scalar4* d_gridk,
                           CUFFTCOMPLEX* d_gridX, CUFFTCOMPLEX* d_gridY, CUFFTCOMPLEX* d_gridZ, cufftHandle plan, const int Nx,
                           const int Ny, const int Nz, unsigned int NxNyNz, dim3 grid, dim3 threads, int gridBlockSize,
                           int gridNBlock, const int P, Scalar3 gridh);

extern "C" {

typedef struct {
    int N;
    float Lx, Ly, Lz, xy;
    float xi, eta, rcut, dr;
    int ewald_n;
    float self;
    int Nx, Ny, Nz, P;
    float hx, hy, hz;
    float error;
} pse_ref_params;

static cufftHandle g_plan;
static int g_plan_dims[3] = {0, 0, 0};
static unsigned int* g_members = nullptr;
static int g_members_n = 0;

static int ensure_plan(int Nx, int Ny, int Nz) {
    if (g_plan_dims[0] == Nx && g_plan_dims[1] == Ny && g_plan_dims[2] == Nz) return 0;
    if (g_plan_dims[0]) cufftDestroy(g_plan);
    g_plan_dims[0] = 0;
    if (cufftPlan3d(&g_plan, Nx, Ny, Nz, CUFFT_C2C) != CUFFT_SUCCESS) return -1;  // PSEv1/Stokes.cc:257
    g_plan_dims[0] = Nx; g_plan_dims[1] = Ny; g_plan_dims[2] = Nz;
    return 0;
}

__global__ void iota_kernel(unsigned int* a, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

static int ensure_members(int N) {
    if (g_members_n >= N) return 0;
    if (g_members) cudaFree(g_members);
    if (cudaMalloc(&g_members, sizeof(unsigned int) * (size_t)N) != cudaSuccess) return -1;
    iota_kernel<<<(N + 255) / 256, 256>>>(g_members, N);
    g_members_n = N;
    return 0;
}

static BoxDim make_box(const pse_ref_params* p) { return BoxDim(pse_make_box(p->Lx, p->Ly, p->Lz, p->xy)); }

struct Shapes { dim3 grid, threads; int gridBlockSize, gridNBlock; unsigned int G; };
static Shapes shapes(const pse_ref_params* p) {  // PSEv1/Stokes.cu:274-284 with block_size = 256 (Stokes.cc:485)
    Shapes s;
    unsigned int block_size = 256;
    s.G = (unsigned int)p->Nx * p->Ny * p->Nz;
    s.grid = dim3(p->N / block_size + 1, 1, 1);
    s.threads = dim3(block_size, 1, 1);
    s.gridBlockSize = (s.G > block_size) ? block_size : s.G;
    s.gridNBlock = (s.G + s.gridBlockSize - 1) / s.gridBlockSize;
    return s;
}

static int finish() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "pse_ref: %s\n", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int pse_ref_setgridk(const pse_ref_params* p, float4* gridk) {
    Shapes s = shapes(p);
    gpu_stokes_SetGridk_kernel<<<s.gridNBlock, s.gridBlockSize>>>(gridk, p->Nx, p->Ny, p->Nz, s.G, make_box(p), p->xi, p->eta);
    return finish();
}

// grids must be zeroed by the caller when a clean spread is wanted
int pse_ref_spread(const pse_ref_params* p, float4* pos, float4* force, cufftComplex* gX, cufftComplex* gY,
                   cufftComplex* gZ, int P, float prefac, float expfac) {
    if (ensure_members(p->N)) return -1;
    int B = (P < 10) ? P : 10;  // PSEv1/Brownian.cu:822-824
    gpu_stokes_Spread_kernel<<<dim3(p->N, 1, 1), dim3(B, B, B)>>>(pos, force, gX, gY, gZ, p->N, p->Nx, p->Ny, p->Nz, g_members,
                                                                 make_box(p), P, make_scalar3(p->hx, p->hy, p->hz), p->xi,
                                                                 p->eta, prefac, expfac);
    return finish();
}

int pse_ref_contract(const pse_ref_params* p, float4* pos, float4* vel, cufftComplex* gX, cufftComplex* gY,
                     cufftComplex* gZ, int P, float prefac, float expfac) {
    if (ensure_members(p->N)) return -1;
    int B = (P < 10) ? P : 10;
    gpu_stokes_Contract_kernel<<<dim3(p->N, 1, 1), dim3(B, B, B), (B * B * B + 1) * sizeof(float3)>>>(
        pos, vel, gX, gY, gZ, p->N, p->Nx, p->Ny, p->Nz, p->xi, p->eta, g_members, make_box(p), P,
        make_scalar3(p->hx, p->hy, p->hz), prefac, expfac);
    return finish();
}

int pse_ref_mreal(const pse_ref_params* p, float4* pos, float4* vel, float4* force, float4* table,
                  const unsigned int* n_neigh, const unsigned int* nlist, const unsigned int* headlist) {
    if (ensure_members(p->N)) return -1;
    Shapes s = shapes(p);
    gpu_stokes_Mreal_kernel<<<s.grid, s.threads>>>(pos, vel, force, p->N, p->xi, table, p->self, p->rcut, p->ewald_n, p->dr,
                                                   g_members, make_box(p), n_neigh, nlist, headlist);
    return finish();
}

// wave-space part only: gpu_stokes_Mwave_wrap (PSEv1/Mobility.cu:515-575), declared there but not in Mobility.cuh
int pse_ref_mwave(const pse_ref_params* p, float4* pos, float4* vel, float4* force, float4* gridk, cufftComplex* gX,
                  cufftComplex* gY, cufftComplex* gZ) {
    if (ensure_members(p->N) || ensure_plan(p->Nx, p->Ny, p->Nz)) return -1;
    Shapes s = shapes(p);
    BoxDim box = make_box(p);
    gpu_stokes_SetGridk_kernel<<<s.gridNBlock, s.gridBlockSize>>>(gridk, p->Nx, p->Ny, p->Nz, s.G, box, p->xi, p->eta);
    gpu_stokes_Mwave_wrap(pos, vel, force, g_members, p->N, box, p->xi, p->eta, gridk, gX, gY, gZ, g_plan, p->Nx, p->Ny, p->Nz,
                          s.G, s.grid, s.threads, s.gridBlockSize, s.gridNBlock, p->P, make_scalar3(p->hx, p->hy, p->hz));
    return finish();
}

// deterministic U = M F through the reference's own driver (PSEv1/Mobility.cu:729-782),
// preceded by the per-step k-table refresh the step path does (PSEv1/Stokes.cu:298)
int pse_ref_mobility(const pse_ref_params* p, float4* pos, float4* vel, float4* force, float4* table, float4* gridk,
                     cufftComplex* gX, cufftComplex* gY, cufftComplex* gZ, const unsigned int* n_neigh,
                     const unsigned int* nlist, const unsigned int* headlist) {
    if (ensure_members(p->N) || ensure_plan(p->Nx, p->Ny, p->Nz)) return -1;
    Shapes s = shapes(p);
    BoxDim box = make_box(p);
    gpu_stokes_SetGridk_kernel<<<s.gridNBlock, s.gridBlockSize>>>(gridk, p->Nx, p->Ny, p->Nz, s.G, box, p->xi, p->eta);
    gpu_stokes_Mobility_wrap(pos, vel, force, g_members, p->N, box, p->xi, p->eta, p->rcut, p->dr, p->ewald_n, table, p->self,
                             gridk, gX, gY, gZ, g_plan, p->Nx, p->Ny, p->Nz, n_neigh, nlist, headlist, s.G, s.grid, s.threads,
                             s.gridBlockSize, s.gridNBlock, p->P, make_scalar3(p->hx, p->hy, p->hz));
    return finish();
}

// velocity of one BD step without the position update (PSEv1/Brownian.cu:772-923)
int pse_ref_velocity(const pse_ref_params* p, float4* pos, float4* vel, float4* force, float4* table, float4* gridk,
                     cufftComplex* gX, cufftComplex* gY, cufftComplex* gZ, const unsigned int* n_neigh,
                     const unsigned int* nlist, const unsigned int* headlist, float T, float dt, unsigned int timestep,
                     unsigned int seed, int* m_lanczos) {
    if (ensure_members(p->N) || ensure_plan(p->Nx, p->Ny, p->Nz)) return -1;
    Shapes s = shapes(p);
    BoxDim box = make_box(p);
    gpu_stokes_SetGridk_kernel<<<s.gridNBlock, s.gridBlockSize>>>(gridk, p->Nx, p->Ny, p->Nz, s.G, box, p->xi, p->eta);
    int m = *m_lanczos;
    gpu_stokes_CombinedMobilityBrownian_wrap(pos, force, g_members, p->N, box, dt, vel, T, timestep, seed, p->xi, p->eta,
                                             (Scalar)p->P, p->rcut, p->dr, p->ewald_n, table, gridk, gX, gY, gZ, g_plan, p->Nx,
                                             p->Ny, p->Nz, n_neigh, nlist, headlist, m, p->N, s.G, s.grid, s.threads,
                                             s.gridBlockSize, s.gridNBlock, make_scalar3(p->hx, p->hy, p->hz), p->error,
                                             p->self);
    *m_lanczos = m;
    return finish();
}

// M_real^{1/2} psi * sqrt(2T/dt) through the reference Lanczos driver (PSEv1/Brownian.cu:357-765)
int pse_ref_lanczos(const pse_ref_params* p, float4* psi, float4* pos, float4* vel, float4* table,
                    const unsigned int* n_neigh, const unsigned int* nlist, const unsigned int* headlist, float T,
                    float dt, int* m_lanczos) {
    if (ensure_members(p->N)) return -1;
    Shapes s = shapes(p);
    int m = *m_lanczos;
    gpu_stokes_BrealLanczos_wrap(psi, pos, g_members, p->N, make_box(p), dt, vel, T, 0u, 0u, p->xi, p->rcut, p->dr, p->ewald_n,
                                 table, n_neigh, nlist, headlist, m, p->error, s.grid, s.threads, s.gridBlockSize,
                                 s.gridNBlock, make_scalar3(p->hx, p->hy, p->hz), p->self);
    *m_lanczos = m;
    return finish();
}

// one full BD step: the reference's host->device entry point (PSEv1/Stokes.cuh:75-111).
// sync != 0 adds a device synchronize + error check (tests); bench passes 0 and times with events.
int pse_ref_step(const pse_ref_params* p, float4* pos, float4* vel, float3* accel, int3* image, float4* force,
                 float4* table, float4* gridk, cufftComplex* gX, cufftComplex* gY, cufftComplex* gZ,
                 const unsigned int* n_neigh, const unsigned int* nlist, const unsigned int* headlist, float T, float dt,
                 unsigned int timestep, unsigned int seed, int* m_lanczos, float shear_rate, int sync) {
    if (ensure_members(p->N) || ensure_plan(p->Nx, p->Ny, p->Nz)) return -1;
    int m = *m_lanczos;
    gpu_stokes_step_one(pos, vel, accel, image, g_members, p->N, make_box(p), dt, 256, force, T, timestep, seed, p->xi, p->eta,
                        p->rcut, p->dr, p->ewald_n, table, p->self, gridk, gX, gY, gZ, g_plan, p->Nx, p->Ny, p->Nz, n_neigh,
                        nlist, headlist, m, p->N, p->P, make_scalar3(p->hx, p->hy, p->hz), p->error, shear_rate);
    *m_lanczos = m;
    return sync ? finish() : 0;
}

}  // extern "C"

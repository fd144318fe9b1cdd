/*
 * pse_b200 — C ABI of the B200-native Positively Split Ewald (PSE) engine.
 *
 * Drop-in boundary for ONE hot path of stochasticHydroTools/PSE: a Brownian-dynamics step with
 * RPY hydrodynamics (real-space near field + wave-space far field + Lanczos Brownian sampling +
 * Euler update) and its deterministic part U = M F.  Each entry point names the reference
 * interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer, `h_` = host pointer
 *   - particle arrays are float4 (x,y,z,w) / int3, contiguous, index = particle id, as in the
 *     reference (`Scalar4 pos/vel/net_force`, `int3 image`; PSEv1/Stokes.cc:436-470)
 *   - the caller owns particle arrays; the engine owns all workspaces (allocated once in
 *     pse_create; the reference's per-step cudaMalloc/cudaFree, PSEv1/Brownian.cu:392-433,
 *     are gone)
 *   - every call is issued on the engine's stream and returns 0 or a negative PSE_E* code;
 *     nothing calls exit() (the reference does: PSEv1/Stokes.cc:203-214, Brownian.cu:543-560)
 *   - there is no CPU fallback: with no usable CUDA device pse_create fails with PSE_ENODEVICE
 */
#ifndef PSE_B200_H
#define PSE_B200_H

#include <stddef.h>
#include <stdint.h>

#if !defined(__VECTOR_TYPES_H__) && !defined(__CUDACC__) && !defined(PSE_HAVE_VECTOR_TYPES)
/* layout-compatible with CUDA's vector types for plain C / FFI consumers */
typedef struct { float x, y, z, w; } float4;
typedef struct { int x, y, z; } int3;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define PSE_OK 0
#define PSE_EINVAL (-1)    /* bad argument */
#define PSE_ENODEVICE (-2) /* no CUDA device / driver */
#define PSE_ECUDA (-3)     /* CUDA runtime or cuFFT error, see pse_last_error */
#define PSE_EGRID (-4)     /* requested Fourier grid exceeds the limit (reference: 512^3) */
#define PSE_ENOMEM (-5)
#define PSE_EEIGEN (-6)    /* Lanczos tridiagonal matrix not positive definite */
#define PSE_ECAPACITY (-7) /* caller buffer too small */

/* flags */
#define PSE_FLAG_REF_PI 1u      /* reproduce the reference's 2*3.1416926536 wave-vector constant
                                   (PSEv1/Helper.cu:313-315); off = exact pi */
#define PSE_FLAG_LIFT_GRID_CAP 2u /* allow grids above the reference's 512^3 cap */

typedef struct pse_engine pse_engine; /* opaque */

/* Periodic box, centred at the origin, sheared in xy (flow x, gradient y).
 * Replaces HOOMD `BoxDim` as used at PSEv1/Mobility.cu:165-238, PSEv1/Stokes.cu:185. */
typedef struct {
    float Lx, Ly, Lz;
    float xy; /* tilt factor */
} pse_box;

/* Constructor arguments of `Stokes` (PSEv1/Stokes.cc:85-111) + `setShear` max_strain
 * (PSEv1/Stokes.h:118-121) + integrator dt (HOOMD `setDeltaT`). */
typedef struct {
    uint32_t N;       /* particles (the reference requires group == all) */
    pse_box box;
    float xi;         /* Ewald splitting parameter */
    float error;      /* requested relative error */
    float max_strain; /* max |box tilt| the Gaussian support is sized for (default 0.5) */
    float T;          /* temperature, kT */
    float dt;         /* time step */
    uint32_t seed;    /* user seed; hashed as PSEv1/Stokes.cc:102 */
    uint32_t flags;   /* PSE_FLAG_* */
    float r_buff;     /* neighbour-list buffer (reference: 0.4, PSEv1/integrate.py:62) */
} pse_config;

/* Everything `Stokes::setParams` derives (PSEv1/Stokes.cc:129-319) + the spreading constants of
 * PSEv1/Brownian.cu:826-829. */
typedef struct {
    int Nx, Ny, Nz;   /* Fourier grid */
    int P;            /* Gaussian support, nodes per dimension */
    int kmax;
    int ewald_n;      /* real-space table has ewald_n + 1 entries */
    float rcut;       /* real-space cutoff */
    float dr;         /* table spacing (0.001) */
    float gaussm;
    float eta;        /* Gaussian splitting parameter */
    float hx, hy, hz; /* grid spacing */
    float self;       /* M_real self term */
    float quadW, prefac, expfac;
    uint32_t seed_hashed;
} pse_params;

/* ---- host-only helpers (no GPU needed) ------------------------------------------------ */

/* Stokes::setParams, PSEv1/Stokes.cc:129-236,309-319. Returns PSE_EGRID above 512^3 unless lifted. */
int pse_derive_params(const pse_config* cfg, pse_params* out);

/* Real-space table, PSEv1/Stokes.cc:322-422: out[4*k..4*k+3] = (f(r_k), g(r_k), f(r_k+1), g(r_k+1)),
 * r_k = (k+1)*dr, k in [0, ewald_n].  `out` holds 4*(ewald_n+1) floats. */
int pse_ewald_table(const pse_config* cfg, float* out);

/* Shear functions: C mirrors of PSEv1/SpecificShearFunction.h:16-223 and
 * PSEv1/VariantShearFunction.{h,cc}.  kind: 0 none (base ShearFunction), 1 steady, 2 sine,
 * 3 chirp, 4 tukey window, 5 windowed(base, window). */
typedef struct pse_shear pse_shear;
pse_shear* pse_shear_create(int kind, const double* args, int nargs, uint32_t offset, double dt);
pse_shear* pse_shear_create_windowed(pse_shear* base, pse_shear* window);
double pse_shear_rate(const pse_shear* s, uint32_t timestep);
double pse_shear_strain(const pse_shear* s, uint32_t timestep);
uint32_t pse_shear_offset(const pse_shear* s);
void pse_shear_destroy(pse_shear* s);
/* VariantShearFunction::getValue (PSEv1/VariantShearFunction.cc:34-43): strain wrapped into [min,max) */
double pse_shear_variant_value(const pse_shear* s, uint32_t total_timestep, double min_value, double max_value,
                               uint32_t timestep);

/* ---- engine lifetime ------------------------------------------------------------------- */

/* Replaces Stokes::Stokes + setShear + setParams.  `stream` is a cudaStream_t (NULL = default). */
int pse_create(const pse_config* cfg, void* stream, pse_engine** out);
void pse_destroy(pse_engine* e);
const char* pse_last_error(const pse_engine* e); /* never NULL; also valid for e == NULL (create errors) */

int pse_get_params(const pse_engine* e, pse_params* out);
int pse_set_box(pse_engine* e, const pse_box* box);    /* box tilt update (HOOMD box_resize) */
/* Re-image every particle into the engine's current box (one image per axis, BoxDim::wrap as at PSEv1/Stokes.cu:185),
 * adjusting d_image when non-NULL.  HOOMD's box_resize updater does this after every box change; it matters when
 * shear_variant flips the tilt from +max_strain to -max_strain (PSEv1/VariantShearFunction.cc:34-43): without it up to
 * |xy| Ly / (2 Lx) of the particles sit outside the primary cell of the new box. */
int pse_wrap_positions(pse_engine* e, float4* d_pos, int3* d_image);
int pse_set_temperature(pse_engine* e, float T);        /* Stokes::setT */
int pse_set_lanczos_m(pse_engine* e, int m);            /* Stokes::m_m_Lanczos (in/out, Stokes.h:157) */
int pse_get_lanczos_m(const pse_engine* e);

/* ---- neighbour list / grid assignment (bit-exact outputs) ------------------------------- */

/* Bin particles and (re)build the real-space neighbour list unconditionally.
 * Replaces HOOMD NeighborListGPUBinned::compute (PSEv1/Stokes.cc:433). */
int pse_build_neighbors(pse_engine* e, const float4* d_pos);

/* Neighbour list in the reference layout (PSEv1/Stokes.cc:436-438): full list, particle ids,
 * each row sorted ascending.  Any pointer may be NULL.  nnz_out receives the total. */
int pse_neighbor_list(pse_engine* e, uint32_t* d_n_neigh, uint32_t* d_headlist, uint32_t* d_nlist,
                      size_t nlist_capacity, size_t* nnz_out);

/* First grid node of each particle's support, wrapped: (x_inp,y_inp,z_inp) at t = 0 of
 * PSEv1/Mobility.cu:173-219.  Bit-exact contract. */
int pse_grid_index(pse_engine* e, const float4* d_pos, int3* d_out);

/* ---- operators --------------------------------------------------------------------------- */

/* U = M_real F (PSEv1/Mobility.cu:594-687), U = M_wave F (:515-575), U = M F (:729-782).
 * pse_mreal writes d_U.w = 0 as the reference kernel does (:632,684); the others leave d_U.w untouched, as the
 * reference does with the mass column of its velocity array (:474, PSEv1/Helper.cu:131).
 * Neighbour structures are refreshed automatically when stale. */
int pse_mreal(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U);
int pse_mwave(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U);
int pse_mobility(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U);

/* Velocity of one BD step without the position update: U = M F + sqrt(2kT/dt) M^{1/2} psi
 * (PSEv1/Brownian.cu:772-923).  Optional injected uniforms in [0,1): d_u_particles [N][3]
 * (x,y,z draws of PSEv1/Brownian.cu:122-124), d_u_grid [Nx*Ny*Nz][6] (reX,reY,reZ,imX,imY,imZ of
 * :184-189, full-grid node index); NULL = engine RNG keyed on (index, timestep + seed).
 * parts: bit 0 deterministic, bit 1 wave-space Brownian, bit 2 real-space Brownian. */
int pse_velocity(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U, uint32_t timestep,
                 const float* d_u_particles, const float* d_u_grid, uint32_t parts, int* m_lanczos_out);

/* One full step: gpu_stokes_step_one (PSEv1/Stokes.cuh:75-111).  Updates d_pos / d_image in
 * place, writes the velocity to d_vel when non-NULL.  Asynchronous apart from the Lanczos
 * convergence checks. */
int pse_step(pse_engine* e, float4* d_pos, int3* d_image, const float4* d_F, float4* d_vel, uint32_t timestep,
             float shear_rate, int* m_lanczos_out);

/* Short-range conservative pair forces on the engine's own neighbour list.  The reference
 * integrates whatever HOOMD pair potentials left in net_force (PSEv1/Stokes.cc:447,457: `m_pdata->getNetForce()`); this
 * entry point is the stand-in for those force computes (hoomd.md.pair.lj / dpd_conservative in HOOMD 2.3.3 terms), so that
 * a suspension can be run without HOOMD.  d_F[i] = (Fx, Fy, Fz, half the pair energy of particle i) as HOOMD's net_force;
 * accumulate != 0 adds to d_F.  r_cut must not exceed the real-space cutoff of the engine (the list is complete up to it).
 *   PSE_PAIR_LJ        U = 4 eps [(sigma/r)^12 - (sigma/r)^6]                     r < r_cut   (HOOMD shift mode "none")
 *   PSE_PAIR_WCA       the same, r_cut = 2^(1/6) sigma, shifted by +eps (purely repulsive)
 *   PSE_PAIR_HARMONIC  U = A (r_cut - r) - A (r_cut^2 - r^2) / (2 r_cut), F = A (1 - r/r_cut) rhat, A = epsilon
 *                      (HOOMD dpd_conservative) */
typedef struct pse_pair_params {
    int32_t kind;
    float epsilon, sigma, r_cut;
} pse_pair_params;
#define PSE_PAIR_LJ 0
#define PSE_PAIR_WCA 1
#define PSE_PAIR_HARMONIC 2
int pse_pair_force(pse_engine* e, const float4* d_pos, const pse_pair_params* prm, float4* d_F, int accumulate);

/* The same step through HOST buffers (pinned or pageable): copies positions, images and forces
 * in, runs pse_step, copies positions and images out, synchronises. */
int pse_step_host(pse_engine* e, float* h_pos4, int* h_image3, const float* h_F4, float* h_vel4, uint32_t timestep,
                  float shear_rate, int* m_lanczos_out);

/* Pipelined form of pse_step_host.  The state (positions, images) stays on the device between calls; a call uploads
 * the step's forces (overlapped with the position-only head of the step), runs the step and STARTS the download of the
 * new state into h_pos4 / h_image3 (/ h_vel4) on a copy stream, which proceeds while the next call computes.
 * Double-buffer the host arrays (or call pse_wait) before reading them.  PSE_HOST_STATE_IN: h_pos4 / h_image3 are read
 * first and replace the device state (implied on the first call and after an error).  Replaces the synchronous
 * d_pos/d_net_force hand-over of Stokes::integrateStepOne (PSEv1/Stokes.cc:436-470) for host-resident callers. */
#define PSE_HOST_STATE_IN 1u
#define PSE_HOST_NO_STATE_OUT 2u /* leave h_pos4 / h_image3 alone (slab-decomposed runs: every rank holds the same state on its
                                    device, one of them downloading it is enough) */
int pse_step_host_async(pse_engine* e, float* h_pos4, int* h_image3, const float* h_F4, float* h_vel4, uint32_t timestep,
                        float shear_rate, uint32_t flags, int* m_lanczos_out);
/* Inputs double-buffered too: start uploading the forces of a LATER pse_step_host_async call now (they must not depend on
 * the steps in flight - external fields, forces computed ahead).  Calls given h_F4 == NULL consume the prefetched sets in
 * order; two may be outstanding, so the upload for step t + 1 can be issued before the call for step t. */
int pse_host_prefetch_forces(pse_engine* e, const float* h_F4);
int pse_wait(pse_engine* e); /* blocks until the outputs of the last pse_step_host_async are on the host */

/* ---- introspection ---------------------------------------------------------------------- */

typedef struct {
    uint64_t nnz;           /* stored neighbours (r < r_cut + r_buff at build time) */
    uint64_t nnz_active;    /* neighbours inside r_cut at the current positions (rows the SpMV walks) */
    uint64_t kernel_launches; /* engine kernels launched since create (cuFFT launches not counted) */
    uint64_t fft_execs;
    uint64_t graph_launches; /* replays of the captured step graph */
    uint64_t nlist_builds;
    int lanczos_m;          /* iterations used by the last Brownian evaluation */
    float lanczos_stepnorm; /* its final relative step norm */
} pse_stats;
int pse_get_stats(pse_engine* e, pse_stats* out); /* synchronises the stream */

/* Optional per-phase device timing with CUDA events on the engine's stream (off by default; adds two
 * event records per phase).  pse_get_profile synchronises the stream, returns the number of phases and
 * fills accumulated milliseconds / span counts since pse_set_profiling(e, 1). */
int pse_set_profiling(pse_engine* e, int on);
int pse_get_profile(pse_engine* e, double* ms_out, uint64_t* calls_out, int n);
const char* pse_profile_phase_name(int i);

/* ---- multi-GPU: slab decomposition of the WHOLE step (one engine per rank / GPU) -------------------------------------
 * New work: the reference is single-GPU ("only one GPU is supported", PSEv1/Stokes.cc:104).
 *
 * After pse_shard_init every operator above (pse_mobility, pse_velocity, pse_step, pse_step_host ...) keeps its
 * signature and meaning: each rank passes the SAME particle arrays and receives the SAME complete result, bit for bit.
 * Inside, a rank works on its own particles only (a contiguous slot range: x layers of cells) and on its own x planes
 * of the Fourier grid, and exchanges exactly what crosses a slab face; all collectives are issued from C++ on the
 * engine's stream through the NCCL C API (bound with dlopen at run time, the copy the process already loaded):
 *   real space   boundary rows of the multiplied vector to both neighbours before every Lanczos product
 *                (ncclSend/ncclRecv), (v.y, |y|^2, |v|^2) in one three-word (double) all-reduce per iteration;
 *   wave space   halo planes added into the neighbours' planes after spreading and fetched before interpolation,
 *                two all-to-all transposes (x slabs <-> y slabs) around the fused x pass;
 *   velocities   one all-gather (N x 16 bytes in total) - positions stay replicated, no migration step.
 * The step is issued eagerly (no CUDA graph) in this mode.  Needs the spread2/interp2 kernels (P = 6, 7, 8) and the
 * engine's own FFT passes. */
typedef struct {
    int rank, world;
    int x0, x1;              /* own x planes of the grid */
    int y0, y1;              /* own (stored) y rows of the transposed k-space layout */
    int halo_left, halo_right; /* planes exchanged with the left / right neighbour */
    int buffer_planes;       /* x planes of the local real-space buffer (own + halo + tile alignment) */
    int layer0, layer1;      /* own x layers of cells (particle ownership) */
    int halo_layers;         /* layers of cells whose vector rows come from each neighbour (0 before the first list build) */
    uint32_t row0, row1;     /* own slots = rows of the real-space operator (0, 0 before the first list build) */
    uint64_t a2a_send_bytes[16], a2a_recv_bytes[16]; /* forward transpose, per peer */
    uint64_t bytes_sent;     /* payload this rank handed to the collectives since init */
    uint64_t collectives;    /* collective calls issued since init */
} pse_shard_info;

/* Host only (no GPU needed): the static part of the decomposition `cfg` gets on `world` ranks, as seen by `rank`.
 * PSE_EINVAL when the grid is too small for that many slabs (a halo would reach past the adjacent rank). */
int pse_shard_plan(const pse_config* cfg, int rank, int world, pse_shard_info* out);

/* 128-byte NCCL unique id (ncclGetUniqueId); rank 0 creates it, the host layer broadcasts it to the other ranks. */
int pse_comm_unique_id(uint8_t out[128]);

/* In-process stand-in for the communicator: `world` engines of ONE process (one host thread each, any devices) exchange
 * through device copies and a host barrier.  Used to run the multi-rank code path as virtual ranks on a single GPU in
 * the parity tests; every collective is a host synchronisation point, so it says nothing about performance. */
typedef struct pse_local_world pse_local_world;
pse_local_world* pse_local_world_create(int world);
void pse_local_world_destroy(pse_local_world* w);

/* Turn `e` into rank `rank` of `world`.  Exactly one of nccl_uid128 (NCCL communicator over the ranks' GPUs; collective
 * call: every rank must enter) and local (virtual ranks) is non-null; both may be null when world == 1.  Must be
 * called before the first operator call.  Frees the single-GPU grids and allocates the rank's slabs. */
int pse_shard_init(pse_engine* e, int rank, int world, const uint8_t* nccl_uid128, pse_local_world* local);
int pse_shard_get_info(pse_engine* e, pse_shard_info* out);

#ifdef __cplusplus
}
#endif
#endif /* PSE_B200_H */

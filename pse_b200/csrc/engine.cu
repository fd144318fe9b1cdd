// Engine: owns the workspaces and sequences the kernels of one PSE Brownian-dynamics step.
//
// Replaces the host drivers of the reference: gpu_stokes_step_one (PSEv1/Stokes.cu:234-365),
// gpu_stokes_CombinedMobilityBrownian_wrap (PSEv1/Brownian.cu:772-923),
// gpu_stokes_BrealLanczos_wrap (:357-765), gpu_stokes_Mobility_wrap / Mwave_wrap
// (PSEv1/Mobility.cu:729-782, :515-575) and the per-step array plumbing of
// Stokes::integrateStepOne (PSEv1/Stokes.cc:429-523).
#include <cuda_runtime.h>
#include <cufft.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: named ranges per phase for timelines (no cost without a tool attached)
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "../../include/pse_b200.h"
#include "box.cuh"
#include "cells.cuh"
#include "common.cuh"
#include "real.cuh"
#include "rng.cuh"
#include "wave.cuh"
#include "wave_tiled.cuh"
#include "wave_v2.cuh"
#include "fft.cuh"
#include "comm.cuh"
#include "peer.cuh"

int pse_tridiag_sqrt_e1(int m, const double* diag, const double* off, double* c, double* lambda_min_out);
int pse_fit_rpy_cheb(double xi, double rcut, float* out, double* max_err_out);

#define LANCZOS_M_MAX 100  // PSEv1/Brownian.cu:397
#define DOTS_BLOCKS 592    // grid of lanczos_dots_kernel (4 blocks per SM)

static char g_create_error[512] = "no error";

// phases for the optional CUDA-event profile (pse_set_profiling / pse_get_profile)
enum Phase { PH_BIN = 0, PH_NLIST, PH_REORDER, PH_WBIN, PH_SPREAD, PH_FFT_FWD, PH_SCALE, PH_FFT_INV, PH_INTERP, PH_PRUNE, PH_SPMV,
             PH_LANCZOS_SPMV, PH_LANCZOS_VEC, PH_COMBINE, PH_INTEGRATE, PH_COMM_TRANS, PH_COMM_HALO, PH_COMM_VEC, PH_COMM_RED, PH_COMM_GATHER,
             PH_COUNT };
static const char* kPhaseNames[PH_COUNT] = {"bin", "nlist", "reorder", "wave_bin", "spread", "fft_r2c", "scale", "fft_c2r",
                                            "interp", "prune", "spmv", "lanczos_spmv", "lanczos_vec", "combine", "integrate", "comm_transpose", "comm_grid_halo",
                                            "comm_vector_halo", "comm_allreduce", "comm_gather"};
struct ProfSpan { int phase; cudaEvent_t a, b; };
// ---- multi-GPU slab decomposition of the whole step (new work: the reference is single-GPU, PSEv1/Stokes.cc:104) ---------------
// One engine per rank / GPU.  Particle arrays at the C ABI stay replicated (every rank passes the same pos / F and gets the
// same complete velocity back, so the plugin-facing calls are unchanged); every internal phase works on the rank's OWN
// particles only - a contiguous range of slots, because slots are x-major cell order and ownership is cut at x layers of
// cells - and exchanges exactly what crosses a slab face:
//   real space   neighbour list, pruning, SpMV and the Lanczos vectors for the own rows; the multiplied vector's boundary rows
//                go to the two neighbours before every product (ncclSend/ncclRecv); (v.y, |y|^2, |v|^2) in ONE three-word (double)
//                all-reduce per iteration;
//   wave space   own particles are spread into a local buffer of own planes + halo planes; halo planes are added into the
//                neighbours' planes; z / y FFT passes on the own planes, all-to-all transpose to y slabs, fused x pass +
//                scaling, all-to-all back, inverse passes; halo planes fetched from the neighbours; interpolation of the own
//                particles;
//   velocities   all-gathered in slot order (N x 16 B in total) so that every rank integrates every particle (positions stay
//                replicated and bitwise identical, no position exchange, no migration step).
// All exchanges are issued from C++ on the engine's stream, nothing goes through Python.  Default transport: kernels that
// read the peers' buffers directly over NVLink behind a device-side flag barrier (peer.cuh; pointers from CUDA IPC handles
// exchanged once); PSE_COMM=coll selects the NCCL collectives of comm.cuh (pack, ncclSend/ncclRecv, unpack) instead.
#define SHARD_MAX_WORLD 16
#define SHARD_DRIFT_NODES 1.0f
struct ShardBounds { int xs[SHARD_MAX_WORLD + 1], ys[SHARD_MAX_WORLD + 1]; int world; };   // kernel argument: plane / y-row bounds
struct ShardGeom {                        // identical on every rank (derived from replicated data only)
    int world;
    int X[SHARD_MAX_WORLD + 1];           // plane bounds: rank r transforms x planes [X[r], X[r + 1])
    int YS[SHARD_MAX_WORLD + 1];          // stored y rows of the transposed (k-space) layout
    int LB[SHARD_MAX_WORLD + 1];          // x layers of cells: particle ownership
    int HL, HR;                           // halo planes exchanged with the left / right neighbour
    // recomputed at every neighbour-list rebuild
    int KH;                               // x layers of cells whose vector rows come from each neighbour
    uint32_t ROW[SHARD_MAX_WORLD + 1];    // slot bounds
    uint32_t SL1[SHARD_MAX_WORLD];        // a rank's first KH layers are rows [ROW[r], SL1[r])  (sent to its left neighbour)
    uint32_t SR0[SHARD_MAX_WORLD];        // its last KH layers are rows [SR0[r], ROW[r + 1])    (sent to its right neighbour)
};
struct ShardState {
    PseComm comm;
    int rank, world;
    ShardGeom g;
    int BL, nown, xorg, nxa;              // own rank: planes in front of the own planes in the local buffer (>= HL, tile aligned)
    size_t plane, Gl;                     // Ny * Nz, component stride of the local buffer
    float2 *d_sloc, *d_tr;                // own x-slab of the spectrum / own y-slab after the transpose
    float *d_a2a_a, *d_a2a_b;             // all-to-all blocks (x-slab side, y-slab side)
    size_t a2a_send_off[SHARD_MAX_WORLD + 1], a2a_recv_off[SHARD_MAX_WORLD + 1];   // bytes, forward transpose
    float *d_hsL, *d_hsR, *d_hrL, *d_hrR; // halo planes: send left / right, received from left / right
    float4* d_uslot;                      // velocities in slot order (own rows computed, the rest all-gathered)
    float4 *d_vsL, *d_vsR, *d_vrL, *d_vrR; // boundary rows of the multiplied vector: send left / right, received from left / right
    size_t vcap;                          // rows each of them holds
    uint32_t *d_layer_start, *h_layer_start;
    uint64_t bytes_sent;                  // per-rank payload handed to the collectives since init (statistics)
    uint64_t collectives;
    // peer-memory transport (peer.cuh)
    bool use_peer, peer_ready;
    bool push;                            // spectrum transposes as pushes (remote stores) instead of pulls
    PeerSync psync;
    unsigned char* d_pad;                 // own flag / reduction pad
    uint32_t epoch[2], red_count;         // synchronisation points per channel / reductions issued so far (same sequence on every rank)
    const PX* peer_px[SHARD_MAX_WORLD];
    const float4* peer_uslot[SHARD_MAX_WORLD];
    const float* peer_grid[SHARD_MAX_WORLD];
    const float2 *peer_sloc[SHARD_MAX_WORLD], *peer_tr[SHARD_MAX_WORLD];
    void* ipc_opened[SHARD_MAX_WORLD][PSE_PEER_NBUF];
    int BLq[SHARD_MAX_WORLD], nxaq[SHARD_MAX_WORLD];   // layout of every rank's local real-space buffer
};
static void shard_free(ShardState* s);

struct pse_engine {
    pse_config cfg;
    pse_params prm;
    PseBox box;
    WaveParams wp;
    RealParams rp;
    CellGrid cg;
    cudaStream_t stream, own_stream;
    cudaStream_t stream2;         // second stream: the real-space branch of a step runs beside the wave-space branch
    cudaEvent_t ev_fork, ev_join;
    cudaStream_t stream_h2d;      // pse_step_host*: forces and images are uploaded beside the position-only head of the step
    cudaStream_t stream_d2h;      // ... and the new state goes back to the host beside the NEXT step's compute
    cudaEvent_t ev_F, ev_img, ev_step, ev_out;
    bool wait_F, wait_img;        // uploads in flight that the next consumer on `stream` has to wait for
    bool wait_out;                // a download of the device-resident state is in flight: integrate must not overwrite it yet
    bool host_state_valid;        // d_hpos / d_himage hold the state of the host-driven run
    bool out_has_vel;             // the download in flight also reads d_vel_work
    bool overlap;
    uint32_t N;
    size_t G, Gh;
    char err[512];
    float rlist;

    // real-space table
    float4* d_table;
    // cells / ordering
    uint32_t *d_cell_of, *d_cell_count, *d_cell_start, *d_scan_tmp, *d_perm, *d_slot_of;
    size_t cell_cap;
    float4 *d_spos, *d_sx, *d_sy;
    PX* d_px;  // packed (position, vector) records for the SpMV
    // neighbour list (slot numbering)
    uint32_t *d_nn, *d_head, *d_nl;
    uint32_t *d_nn_act, *d_nl_act;  // per-step pruned list (pairs inside r_cut at the current positions)
    size_t nl_act_cap;
    bool prune, pruned_valid;
    bool csr_valid;                // d_nl holds the packed copy of the buffered list (made on demand)
    bool wbin_valid;               // W order / records / factor rows current for the positions of this call
    size_t nl_cap;
    uint32_t nl_stride;            // row stride of the fixed-stride search output
    uint32_t* d_ell;
    size_t ell_cap;
    unsigned long long* d_nlinfo;  // [0] nnz, [1] max row length
    unsigned long long* h_nlinfo;  // pinned
    uint64_t nnz;
    float4* d_pos_build;
    float xy_build;
    float xy_prev_call;  // tilt at the previous velocity evaluation (graph replay only while it stays put)
    float xy_spos;       // tilt the slot-ordered positions / pruned list / W records were made for
    bool reuse_static;   // keep them when a call finds every particle where the previous call left it (PSE_REUSE=0 turns it off)
    bool nlist_valid;
    uint32_t* d_flag;
    uint32_t* h_flag;  // pinned
    cudaEvent_t flag_event;
    bool flag_pending;
    const float4* precheck_pos;   // positions the pending displacement check was issued for (engine-owned state only), and their tilt
    float precheck_xy;
    // wave space
    float* d_grid;
    float2* d_spec;
    cufftHandle plan_f, plan_b;
    bool plans_ok;
    // own shared-memory FFT passes (fft.cuh); cuFFT stays as the fallback (PSE_FFT=cufft) and for the sharded path
    bool own_fft;
    Fft1D fft_ax[3];  // x, y, z
    void* d_fft_tables;
    // tile-owned spreading / interpolation ("W order", rebinned per call)
    bool tiled;
    int spread_var;   // record feed of spread2_kernel (Spread2Cfg): PSE_SPREAD_VAR
    bool wave_v2;     // spread2 / interp2 (every particle once, window merged with vector reductions) instead of the tile-owned kernels
    TileGrid tg;
    int4 *d_org, *d_worg;
    float* d_wrecs;  // v2 wave kernels: one record per particle, W order (12 header words + Gaussian factor row)
    uint32_t *d_wcell_of, *d_wcount, *d_wstart, *d_wperm, *d_wtmp, *d_wid;
    float4 *d_wpos, *d_wF;
    float* d_wwt;  // Gaussian factor rows, W order: [N][P*P + P]
    // Lanczos
    float4 *d_V, *d_u, *d_y;
    float *d_alpha, *d_beta, *d_coef, *d_partials;
    unsigned int* d_counter;
    float* h_ab;  // pinned: alpha[m_max + 1] | beta[m_max + 3] | coefficients[m_max + 2]
    int m_lanczos;
    float last_stepnorm;
    // step scratch
    float4* d_vel_work;
    float4 *d_hpos, *d_hF;  // device staging for pse_step_host
    float4* d_hF_next[2];   // landing buffers of pse_host_prefetch_forces: a queue of two (copied into d_hF by the step that consumes one)
    cudaEvent_t ev_Fnext[2], ev_Fcons[2];
    uint64_t n_pref, n_cons;   // prefetches issued / consumed
    int take_Fnext;            // landing buffer the step in flight takes its forces from, or -1
    int3* d_himage;
    int num_sms;
    struct ShardState* shard;  // multi-GPU slab decomposition of the whole step (pse_shard_init); null = single GPU
    uint32_t row0, row1;       // slots (rows) this engine works on: [0, N) unless slab-decomposed
    size_t v_rows;             // rows per Krylov vector in d_V
    size_t grid_planes;        // x planes allocated in d_grid (Nx unless slab-decomposed)
    double* d_red2;            // slab-decomposed Lanczos: (v.y, |y|^2, |v|^2) partial sums -> all-reduced | per-block partials
    // per-step device scalars + captured step graph
    StepDev* d_stepdev;
    StepDev* h_stepdev;  // pinned
    bool use_graph;
    cudaGraphExec_t graph_exec;
    struct { const void *pos, *F, *U; int m_batch; float xy; uint32_t nl_gen; bool valid; } graph_key;
    uint32_t nl_gen;  // bumped whenever list buffers are reallocated
    bool spmv_smem_table;
    int spmv_table_mode;  // TABLE_GLOBAL / TABLE_SHARED / TABLE_POLY
    ChebCoef cheb;        // constant-bank polynomial form of f, g for r >= 2a
    double cheb_max_err;
    int spmv_tpp;  // lanes per row in the SpMV
    int spmv_bps;  // blocks per SM of the persistent SpMV grid (0 = auto)
    bool spmv_dual;  // M_real F computed inside the first Lanczos product of a full step
    // profiling
    bool prof_on;
    std::vector<cudaEvent_t>* prof_pool;
    std::vector<ProfSpan>* prof_spans;
    size_t prof_used;
    double prof_ms[PH_COUNT];
    uint64_t prof_calls[PH_COUNT];
    // stats
    uint64_t launches, fft_execs, nlist_builds, graph_launches, graph_nodes;
};

// bytes of the real-space table when staged in shared memory by the SpMV (float2 per knot)
static inline size_t spmv_table_smem(const pse_engine* e) { return (size_t)(e->prm.ewald_n + 2) * sizeof(float2); }

static int fail(pse_engine* e, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e ? e->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return fail(e, PSE_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)
#define CKFFT(call)                                                                                     \
    do {                                                                                                \
        cufftResult _r = (call);                                                                        \
        if (_r != CUFFT_SUCCESS) return fail(e, PSE_ECUDA, "%s:%d %s: cufft error %d", __FILE__, __LINE__, #call, (int)_r); \
    } while (0)
#define LAUNCHED(e) ((e)->launches++)
#define CKRC(call)              \
    do {                        \
        int _rc = (call);       \
        if (_rc != PSE_OK) return _rc; \
    } while (0)

// RAII span: records two events on the engine stream around a phase when profiling is on
struct ProfScope {
    pse_engine* e;
    cudaEvent_t b;
    bool on;
    ProfScope(pse_engine* eng, int phase) : e(eng), b(nullptr), on(eng->prof_on) {
        nvtxRangePushA(kPhaseNames[phase]);
        if (!on) return;
        auto& pool = *e->prof_pool;
        while (pool.size() < e->prof_used + 2) { cudaEvent_t ev; cudaEventCreate(&ev); pool.push_back(ev); }
        cudaEvent_t a = pool[e->prof_used++];
        b = pool[e->prof_used++];
        cudaEventRecord(a, e->stream);
        e->prof_spans->push_back({phase, a, b});
    }
    ~ProfScope() { if (on) cudaEventRecord(b, e->stream); nvtxRangePop(); }
};
static void prof_collect(pse_engine* e) {
    if (!e->prof_spans || e->prof_spans->empty()) return;
    cudaStreamSynchronize(e->stream);
    for (const ProfSpan& sp : *e->prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { e->prof_ms[sp.phase] += ms; e->prof_calls[sp.phase]++; }
    }
    e->prof_spans->clear();
    e->prof_used = 0;
}
extern "C" int pse_set_profiling(pse_engine* e, int on) {
    if (!e) return PSE_EINVAL;
    prof_collect(e);
    e->prof_on = on != 0;
    for (int i = 0; i < PH_COUNT; ++i) { e->prof_ms[i] = 0; e->prof_calls[i] = 0; }
    return PSE_OK;
}
extern "C" int pse_get_profile(pse_engine* e, double* ms_out, uint64_t* calls_out, int n) {
    if (!e || n < 0) return PSE_EINVAL;
    prof_collect(e);
    for (int i = 0; i < n && i < PH_COUNT; ++i) { if (ms_out) ms_out[i] = e->prof_ms[i]; if (calls_out) calls_out[i] = e->prof_calls[i]; }
    return PH_COUNT;
}
extern "C" const char* pse_profile_phase_name(int i) { return (i >= 0 && i < PH_COUNT) ? kPhaseNames[i] : ""; }

static inline unsigned int nblk(size_t n, int b) { return (unsigned int)((n + b - 1) / b); }

// Control words between host and device travel in KERNELS, not through the copy engines: a few bytes written to / read
// from pinned host memory (device-addressable under unified addressing).  A cudaMemcpyAsync of 4 bytes on the compute
// stream queues on the same DMA engine as the bulk transfers of pse_step_host_async and would wait behind 16-28 MB of them
// (measured: 0.5 ms per step).
__global__ void words_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__global__ void words2_copy_kernel(uint32_t* __restrict__ dst_a, const uint32_t* __restrict__ src_a, int na, uint32_t* __restrict__ dst_b,
                                   const uint32_t* __restrict__ src_b, int nb) {
    for (int i = threadIdx.x; i < na; i += blockDim.x) dst_a[i] = src_a[i];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) dst_b[i] = src_b[i];
}
// device-to-device copies of the step path are kernels too: a cudaMemcpyAsync D2D may be scheduled on a copy engine and then
// waits behind the bulk host transfers of the pipelined entry point like the control words did
__global__ void copy_f4_kernel(float4* __restrict__ dst, const float4* __restrict__ src, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
struct CoefArg { float c[LANCZOS_M_MAX + 2]; };
__global__ void coef_store_kernel(float* __restrict__ dst, CoefArg a, int m) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) dst[i] = a.c[i];
}
__global__ void stepdev_store_kernel(StepDev* __restrict__ dst, StepDev v) { *dst = v; }
#define WORDS(p) reinterpret_cast<uint32_t*>(p)
#define CWORDS(p) reinterpret_cast<const uint32_t*>(p)

// persistent launch: a whole number of waves of resident blocks
static inline unsigned int persistent_grid(const pse_engine* e, size_t work_blocks, int blocks_per_sm) {
    size_t cap = (size_t)e->num_sms * blocks_per_sm;
    return (unsigned int)(work_blocks < cap ? (work_blocks ? work_blocks : 1) : cap);
}

// ---- exclusive scan driver -------------------------------------------------------------------
static int exclusive_scan(pse_engine* e, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp) {
    unsigned int nb = nblk(n, SCAN_BLOCK);
    scan_block_kernel<<<nb, SCAN_BLOCK, 0, e->stream>>>(in, out, tmp, n);
    LAUNCHED(e);
    if (nb > 1) {
        uint32_t* next = tmp + ((nb + 31) / 32) * 32;
        CKRC(exclusive_scan(e, tmp, tmp, nb, next));
        scan_add_kernel<<<nb, SCAN_BLOCK, 0, e->stream>>>(out, tmp, n);
        LAUNCHED(e);
    }
    return PSE_OK;
}

// ---- parameter blocks --------------------------------------------------------------------------
static void refresh_box(pse_engine* e, const pse_box& b) {
    e->cfg.box = b;
    e->box = pse_make_box(b.Lx, b.Ly, b.Lz, b.xy);
}

// cells of about half the list radius, bounded by `cell_cap` cells (sparse systems / memory); host-only so that the
// multi-GPU planner (pse_shard_plan) sees the same x layers as the engine
static void cell_grid_dims(float Lx, float Ly, float Lz, float rlist, size_t cell_cap, int* ncx, int* ncy, int* ncz) {
    const float target = rlist * 0.5f;
    int cx = (int)fmaxf(1.f, floorf(Lx / target)), cy = (int)fmaxf(1.f, floorf(Ly / target)), cz = (int)fmaxf(1.f, floorf(Lz / target));
    while ((size_t)cx * cy * cz > cell_cap) {
        if (cx >= cy && cx >= cz) cx = (cx + 1) / 2;
        else if (cy >= cz) cy = (cy + 1) / 2;
        else cz = (cz + 1) / 2;
    }
    *ncx = cx; *ncy = cy; *ncz = cz;
}
static inline size_t cell_capacity(size_t N) { return std::max<size_t>(4096, 4 * N); }
static void setup_cell_grid(pse_engine* e) {
    CellGrid& cg = e->cg;
    cell_grid_dims(e->box.Lx, e->box.Ly, e->box.Lz, e->rlist, e->cell_cap, &cg.ncx, &cg.ncy, &cg.ncz);
    cg.ncell = cg.ncx * cg.ncy * cg.ncz;
    // the fractional x reach is widened by the tilt
    const float xy = fabsf(e->box.xy);
    const float safety = 1.0001f;
    cg.reach_fx = e->rlist * sqrtf(1.f + xy * xy) / e->box.Lx * safety + 1e-6f;
    cg.reach_fy = e->rlist / e->box.Ly * safety + 1e-6f;
    cg.reach_fz = e->rlist / e->box.Lz * safety + 1e-6f;
}

// ---- create / destroy ---------------------------------------------------------------------------
extern "C" const char* pse_last_error(const pse_engine* e) { return e ? e->err : g_create_error; }

static int alloc_all(pse_engine* e) {
    const size_t N = e->N;
    const pse_params& p = e->prm;
    CK(cudaMalloc(&e->d_table, sizeof(float4) * (p.ewald_n + 1)));
    e->cell_cap = cell_capacity(N);
    CK(cudaMalloc(&e->d_cell_of, sizeof(uint32_t) * N));
    CK(cudaMalloc(&e->d_cell_count, sizeof(uint32_t) * (e->cell_cap + 1)));
    CK(cudaMalloc(&e->d_cell_start, sizeof(uint32_t) * (e->cell_cap + 1)));
    size_t scan_n = std::max(e->cell_cap + 1, N + 1);
    CK(cudaMalloc(&e->d_scan_tmp, sizeof(uint32_t) * (scan_n / SCAN_BLOCK + 4096)));
    CK(cudaMalloc(&e->d_perm, sizeof(uint32_t) * N));
    CK(cudaMalloc(&e->d_slot_of, sizeof(uint32_t) * N));
    CK(cudaMalloc(&e->d_spos, sizeof(float4) * N));
    CK(cudaMalloc(&e->d_sx, sizeof(float4) * N));
    CK(cudaMalloc(&e->d_sy, sizeof(float4) * N));
    CK(cudaMalloc(&e->d_px, sizeof(PX) * N));
    CK(cudaMalloc(&e->d_nn, sizeof(uint32_t) * (N + 1)));
    CK(cudaMalloc(&e->d_head, sizeof(uint32_t) * (N + 1)));
    CK(cudaMalloc(&e->d_nn_act, sizeof(uint32_t) * (N + 1)));
    e->d_nl_act = nullptr; e->nl_act_cap = 0;
    e->prune = true;
    { const char* env = getenv("PSE_PRUNE"); if (env) e->prune = env[0] != '0'; }
    e->d_nl = nullptr; e->nl_cap = 0; e->nl_stride = 0;
    CK(cudaMalloc(&e->d_nlinfo, 2 * sizeof(unsigned long long)));
    CK(cudaMallocHost(&e->h_nlinfo, 2 * sizeof(unsigned long long)));
    CK(cudaMalloc(&e->d_pos_build, sizeof(float4) * N));
    CK(cudaMalloc(&e->d_flag, 4 * sizeof(uint32_t)));   // [0] largest squared displacement (bits), [1] slab guards, [2] moved since the last call
    CK(cudaMemset(e->d_flag, 0, 4 * sizeof(uint32_t)));
    CK(cudaMallocHost(&e->h_flag, 4 * sizeof(uint32_t)));
    CK(cudaEventCreateWithFlags(&e->flag_event, cudaEventDisableTiming));
    // the real grids, the spectra and the Krylov basis are allocated at first use (ensure_wave_buffers / ensure_krylov):
    // a slab-decomposed engine (pse_shard_init) only ever holds its own slab of each
    e->d_grid = nullptr; e->d_spec = nullptr; e->d_V = nullptr; e->v_rows = 0; e->grid_planes = 0;
    CK(cudaMalloc(&e->d_red2, (4 + 3 * DOTS_BLOCKS) * sizeof(double)));
    CK(cudaMalloc(&e->d_u, sizeof(float4) * N));
    CK(cudaMalloc(&e->d_y, sizeof(float4) * N));
    CK(cudaMalloc(&e->d_alpha, sizeof(float) * (LANCZOS_M_MAX + 2)));
    CK(cudaMalloc(&e->d_beta, sizeof(float) * (LANCZOS_M_MAX + 2)));
    CK(cudaMalloc(&e->d_coef, sizeof(float) * (LANCZOS_M_MAX + 2)));
    CK(cudaMalloc(&e->d_partials, sizeof(float) * (N / 8 + 1024)));
    CK(cudaMalloc(&e->d_counter, 2 * sizeof(unsigned int)));
    CK(cudaMemset(e->d_counter, 0, 2 * sizeof(unsigned int)));
    CK(cudaMallocHost(&e->h_ab, sizeof(float) * (3 * LANCZOS_M_MAX + 8)));  // + the combination coefficients
    CK(cudaMalloc(&e->d_vel_work, sizeof(float4) * N));
    CK(cudaMalloc(&e->d_stepdev, sizeof(StepDev)));
    CK(cudaMallocHost(&e->h_stepdev, sizeof(StepDev)));
    e->use_graph = true;
    { const char* env = getenv("PSE_GRAPH"); if (env) e->use_graph = env[0] != '0'; }
    e->d_hpos = e->d_hF = nullptr; e->d_himage = nullptr;
    {
        const WaveParams& wp = e->wp;
        TileGrid& tg = e->tg;
        tg.tx = tg.ty = tg.tz = TILE;
        tg.cp = wp.P; tg.cy = tg.cz = 1; tg.cs = 0;
        {
            // spread2 / interp2 tile shape for this support size (PSE_WAVE=v1 keeps the bitwise-reproducible tile-owned kernels)
            const char* wv = getenv("PSE_WAVE");
            int tx, ty, tz;
            e->wave_v2 = !(wv && wv[0] == 'v' && wv[1] == '1') && v2_shape(wp.P, &tx, &ty, &tz) &&
                         wp.Nx >= tx + wp.P && wp.Ny >= ty + wp.P && wp.Nz >= tz + wp.P;
            if (e->wave_v2) v2_fill_tilegrid(wp.P, tx, ty, tz, &tg);
            { const char* dbg = getenv("PSE_SPREAD_DBG"); tg.dbg = dbg ? atoi(dbg) : 0; }
            const char* sv = getenv("PSE_SPREAD_VAR");
            e->spread_var = sv ? atoi(sv) : 1;
        }
        tg.ntx = (wp.Nx + tg.tx - 1) / tg.tx; tg.nty = (wp.Ny + tg.ty - 1) / tg.ty; tg.ntz = (wp.Nz + tg.tz - 1) / tg.tz;
        tg.ntile = tg.ntx * tg.nty * tg.ntz;
        const int need = TILE + wp.P;
        e->tiled = e->wave_v2 || (wp.P >= 2 && wp.P <= TILED_MAX_P && wp.Nx >= need && wp.Ny >= need && wp.Nz >= need);
        const char* env = getenv("PSE_WAVE_TILED");
        if (env && env[0] == '0') { e->tiled = false; e->wave_v2 = false; }
        if (e->tiled) {
            CK(cudaMalloc(&e->d_org, sizeof(int4) * N));
            CK(cudaMalloc(&e->d_worg, sizeof(int4) * N));
            if (e->wave_v2) CK(cudaMalloc(&e->d_wrecs, sizeof(float) * (size_t)N * wrec_stride(wp.P)));
            CK(cudaMalloc(&e->d_wcell_of, sizeof(uint32_t) * N));
            CK(cudaMalloc(&e->d_wcount, sizeof(uint32_t) * (tg.ntile + 1)));
            CK(cudaMalloc(&e->d_wstart, sizeof(uint32_t) * (tg.ntile + 1)));
            CK(cudaMalloc(&e->d_wperm, sizeof(uint32_t) * N));
            CK(cudaMalloc(&e->d_wid, sizeof(uint32_t) * N));
            CK(cudaMalloc(&e->d_wtmp, sizeof(uint32_t) * N));
            CK(cudaMalloc(&e->d_wpos, sizeof(float4) * N));
            CK(cudaMalloc(&e->d_wF, sizeof(float4) * N));
            if (e->wave_v2) CK(v2_set_attributes(wp.P, tg.tx, tg.ty, tg.tz));
            else {
                CK(cudaMalloc(&e->d_wwt, sizeof(float) * (size_t)N * wrow_stride(wp.P)));
                CK(tiled_set_attributes(wp.P));
            }
        }
    }
    return PSE_OK;
}

// ---- own FFT: radix schedule and tables of one axis ------------------------------------------------------
static bool fft_factor(int N, Fft1D* f) {
    f->N = N; f->npass = 0; f->radices = 0ull;
    int n = N;
    const int radices[4] = {4, 2, 3, 5};
    for (int r : radices)
        while (n % r == 0 && n > 1) {
            if (f->npass == FFT_MAX_PASSES) return false;
            f->radices |= (unsigned long long)r << (4 * f->npass++);
            n /= r;
        }
    return n == 1 && N >= 2 && N <= FFT_MAX_N;
}
static int setup_own_fft(pse_engine* e) {
    e->own_fft = true;
    { const char* env = getenv("PSE_FFT"); if (env && env[0] == 'c') e->own_fft = false; }
    const int dims[3] = {e->wp.Nx, e->wp.Ny, e->wp.Nz};
    for (int a = 0; a < 3; ++a)
        if (!fft_factor(dims[a], &e->fft_ax[a])) e->own_fft = false;
    if (e->wp.Nx * e->wp.Ny * 3ll >= (1ll << 31)) e->own_fft = false;
    if (!e->own_fft) return PSE_OK;
    // one allocation: per axis tw[N] (float2) | pos_of[N] | freq_of[N] (uint16)
    size_t bytes = 0, off_tw[3], off_pos[3], off_frq[3];
    for (int a = 0; a < 3; ++a) {
        off_tw[a] = bytes; bytes += sizeof(float2) * dims[a];
        off_pos[a] = bytes; bytes += sizeof(uint16_t) * dims[a];
        off_frq[a] = bytes; bytes += sizeof(uint16_t) * dims[a];
        bytes = (bytes + 15) / 16 * 16;
    }
    std::vector<unsigned char> host(bytes, 0);
    for (int a = 0; a < 3; ++a) {
        const Fft1D& f = e->fft_ax[a];
        const int N = dims[a];
        float2* tw = reinterpret_cast<float2*>(host.data() + off_tw[a]);
        uint16_t* pos = reinterpret_cast<uint16_t*>(host.data() + off_pos[a]);
        uint16_t* frq = reinterpret_cast<uint16_t*>(host.data() + off_frq[a]);
        for (int k = 0; k < N; ++k) {
            const double ang = -2.0 * 3.14159265358979323846 * k / N;
            tw[k] = make_float2((float)cos(ang), (float)sin(ang));
            int kk = k, span = N, p = 0;  // k = k1 + r1 (k2 + r2 (...)): digit k_i selects sub-block k_i of length span / r_i
            for (int i = 0; i < f.npass; ++i) {
                const int r = (int)((f.radices >> (4 * i)) & 15ull);
                span /= r;
                p += (kk % r) * span;
                kk /= r;
            }
            pos[k] = (uint16_t)p;
            frq[p] = (uint16_t)k;
        }
    }
    CK(cudaMalloc(&e->d_fft_tables, bytes));
    CK(cudaMemcpy(e->d_fft_tables, host.data(), bytes, cudaMemcpyHostToDevice));
    for (int a = 0; a < 3; ++a) {
        unsigned char* base = static_cast<unsigned char*>(e->d_fft_tables);
        e->fft_ax[a].tw = reinterpret_cast<const float2*>(base + off_tw[a]);
        e->fft_ax[a].pos_of = reinterpret_cast<const uint16_t*>(base + off_pos[a]);
        e->fft_ax[a].freq_of = reinterpret_cast<const uint16_t*>(base + off_frq[a]);
    }
    const size_t smz = fft_smem_bytes(dims[2], FFT_Z_COLS + 1), smy = fft_smem_bytes(dims[1], FFT_Y_CP), smx = fft_smem_bytes(dims[0], FFT_X_CP);
    if (std::max(smz, std::max(smy, smx)) > 200 * 1024) { e->own_fft = false; return PSE_OK; }
    CK(cudaFuncSetAttribute(fft_z_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smz));
    CK(cudaFuncSetAttribute(fft_z_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smz));
    CK(cudaFuncSetAttribute(fft_y_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smy));
    CK(cudaFuncSetAttribute(fft_y_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smy));
    CK(cudaFuncSetAttribute(fft_x_scale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smx));
    return PSE_OK;
}

extern "C" int pse_create(const pse_config* cfg, void* stream, pse_engine** out) {
    pse_engine* e = nullptr;
    if (!cfg || !out) return fail(e, PSE_EINVAL, "pse_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(e, PSE_ENODEVICE, "pse_create: no CUDA device (there is no CPU fallback)");
    }
    pse_params prm;
    int rc = pse_derive_params(cfg, &prm);
    if (rc != PSE_OK) return fail(e, rc, "pse_create: parameter derivation failed (%d)%s", rc,
                                  rc == PSE_EGRID ? ": Fourier grid above 512^3, reduce xi (PSEv1/Stokes.cc:203)" : "");
    pse_config c = *cfg;
    if (!(c.r_buff >= 0.f)) c.r_buff = 0.4f;
    const float rlist = prm.rcut + c.r_buff;
    const float plane_x = c.box.Lx / sqrtf(1.f + c.max_strain * c.max_strain);
    if (2.f * rlist > fminf(plane_x, fminf(c.box.Ly, c.box.Lz)))
        return fail(e, PSE_EINVAL, "pse_create: box too small for the minimum-image real-space sum (2*(rcut+r_buff) = %g)",
                    2.f * rlist);
    if (c.dt <= 0.f) return fail(e, PSE_EINVAL, "pse_create: dt must be positive");

    pse_engine* eng = new pse_engine();
    memset(eng, 0, sizeof(*eng));
    e = eng;
    strcpy(e->err, "no error");
    e->cfg = c; e->prm = prm; e->N = c.N;
    e->stream = (cudaStream_t)stream;
    e->own_stream = nullptr;
    e->stream2 = nullptr; e->ev_fork = e->ev_join = nullptr;
    { const char* v = getenv("PSE_OVERLAP"); e->overlap = v ? atoi(v) != 0 : true; }
    if (!e->stream) {
        // the legacy default stream cannot be captured into a CUDA graph: use an own (blocking) stream, which keeps
        // the implicit ordering with work the caller issues on the default stream
        if (cudaStreamCreate(&e->own_stream) != cudaSuccess) { delete eng; return fail(nullptr, PSE_ECUDA, "pse_create: cudaStreamCreate failed"); }
        e->stream = e->own_stream;
    }
    if (e->overlap && (cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking) != cudaSuccess ||
                       cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                       cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess)) {
        pse_destroy(e);
        return fail(nullptr, PSE_ECUDA, "pse_create: second stream / events");
    }
    e->rlist = rlist;
    refresh_box(e, c.box);
    e->xy_prev_call = c.box.xy;
    e->G = (size_t)prm.Nx * prm.Ny * prm.Nz;
    WaveParams& wp = e->wp;
    wp.Nx = prm.Nx; wp.Ny = prm.Ny; wp.Nz = prm.Nz; wp.Nzh = prm.Nz / 2 + 1; wp.P = prm.P;
    wp.Nzp = wp.Nzh;
    { const char* env = getenv("PSE_SPEC_PAD"); int pad = env ? atoi(env) : 8; if (pad > 1) wp.Nzp = ((wp.Nzh + pad - 1) / pad) * pad; }  // 64-byte rows
    wp.hx = prm.hx; wp.hy = prm.hy; wp.hz = prm.hz;
    wp.prefac = prm.prefac; wp.expfac = prm.expfac; wp.quadW = prm.quadW;
    wp.xi = c.xi; wp.eta = prm.eta;
    wp.xorg = 0; wp.nxa = prm.Nx; wp.nxw = prm.Nx;
    wp.two_pi_k = (c.flags & PSE_FLAG_REF_PI) ? (float)(2.0 * 3.1416926536) : (float)(2.0 * 3.14159265358979323846);
    e->Gh = (size_t)prm.Nx * prm.Ny * wp.Nzp;
    RealParams& rp = e->rp;
    rp.self = prm.self; rp.rcut = prm.rcut; rp.rcut_sq = prm.rcut * prm.rcut; rp.dr = prm.dr; rp.dr_sq = prm.dr * prm.dr;
    rp.ewald_n = prm.ewald_n;
    rp.inv_dr = 1.0f / prm.dr; rp.tab_scale = (float)prm.ewald_n / (prm.rcut - prm.dr);
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e->num_sms = sms;
        e->spmv_smem_table = spmv_table_smem(e) <= 64 * 1024;
        const char* env = getenv("PSE_SPMV_SMEM_TABLE");
        if (env) e->spmv_smem_table = env[0] != '0' && spmv_table_smem(e) <= 200 * 1024;
        e->spmv_table_mode = e->spmv_smem_table ? TABLE_SHARED : TABLE_GLOBAL;
        // measured on B200, per SpMV at N = 1M (profiles/r1_summary.md): the kernel is bound by the L1/shared data pipe, a
        // third of whose wavefronts were table knots.  "poly" (default) evaluates f, g from constant-bank polynomials instead;
        // "shared" / "global" keep the reference's linearly interpolated table.
        const char* tm = getenv("PSE_SPMV_TABLE");
        if (!tm || tm[0] == 'p') e->spmv_table_mode = TABLE_POLY;
        else if (tm[0] == 'g') e->spmv_table_mode = TABLE_GLOBAL;
        e->spmv_tpp = 4;  // measured: 4 lanes x 12 entries in flight 202 us, 8 x 6: 236 us, 16 x 3: 330 us per SpMV at N = 1M
        { const char* v = getenv("PSE_SPMV_BPS"); e->spmv_bps = v ? atoi(v) : 0; }
        { const char* v = getenv("PSE_SPMV_DUAL"); e->spmv_dual = v ? atoi(v) != 0 : true; }
        const char* tpp = getenv("PSE_SPMV_TPP");
        if (tpp) e->spmv_tpp = atoi(tpp);
    }
    { const char* v = getenv("PSE_REUSE"); e->reuse_static = v ? atoi(v) != 0 : true; }
    e->take_Fnext = -1;
    e->m_lanczos = 2;  // PSEv1/Stokes.cc:132
    e->row0 = 0; e->row1 = c.N;
    e->prof_pool = new std::vector<cudaEvent_t>();
    e->prof_spans = new std::vector<ProfSpan>();

    rc = alloc_all(e);
    if (rc != PSE_OK) { strncpy(g_create_error, e->err, 511); pse_destroy(e); return rc; }
    setup_cell_grid(e);

    std::vector<float> tab(4 * (size_t)(prm.ewald_n + 1));
    pse_ewald_table(&c, tab.data());
    if (cudaMemcpy(e->d_table, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(nullptr, PSE_ECUDA, "pse_create: table upload failed");
        pse_destroy(e);
        return PSE_ECUDA;
    }
    {
        float pc[3 + 2 * (PSE_CHEB_DEG + 1)];
        const int frc = pse_fit_rpy_cheb(c.xi, prm.rcut, pc, &e->cheb_max_err);
        if (frc != PSE_OK || !(e->cheb_max_err < 3e-6)) {  // keep the table when the fit is not at rounding level
            if (e->spmv_table_mode == TABLE_POLY) e->spmv_table_mode = e->spmv_smem_table ? TABLE_SHARED : TABLE_GLOBAL;
        } else {
            e->cheb.A = pc[0]; e->cheb.B = pc[1]; e->cheb.ne = pc[2];
            for (int k = 0; k <= PSE_CHEB_DEG; ++k) { e->cheb.cf[k] = pc[3 + k]; e->cheb.cg[k] = pc[3 + PSE_CHEB_DEG + 1 + k]; }
        }
    }
    int n[3] = {prm.Nx, prm.Ny, prm.Nz};
    int rembed[3] = {prm.Nx, prm.Ny, prm.Nz}, cembed[3] = {prm.Nx, prm.Ny, wp.Nzp};  // spectrum rows padded to Nzp
    if (cufftPlanMany(&e->plan_f, 3, n, rembed, 1, (int)e->G, cembed, 1, (int)e->Gh, CUFFT_R2C, 3) != CUFFT_SUCCESS ||
        cufftPlanMany(&e->plan_b, 3, n, cembed, 1, (int)e->Gh, rembed, 1, (int)e->G, CUFFT_C2R, 3) != CUFFT_SUCCESS) {
        fail(nullptr, PSE_ECUDA, "pse_create: cufftPlanMany failed for %dx%dx%d", n[0], n[1], n[2]);
        pse_destroy(e);
        return PSE_ECUDA;
    }
    e->plans_ok = true;
    {
        const int frc = setup_own_fft(e);
        if (frc != PSE_OK) { pse_destroy(e); return frc; }
    }
    cufftSetStream(e->plan_f, e->stream);
    cufftSetStream(e->plan_b, e->stream);
    *out = e;
    return PSE_OK;
}

extern "C" void pse_destroy(pse_engine* e) {
    if (!e) return;
    if (e->plans_ok) { cufftDestroy(e->plan_f); cufftDestroy(e->plan_b); }
    if (e->d_fft_tables) cudaFree(e->d_fft_tables);
    void* bufs[] = {e->d_table, e->d_cell_of, e->d_cell_count, e->d_cell_start, e->d_scan_tmp, e->d_perm, e->d_slot_of,
                    e->d_spos, e->d_sx, e->d_sy, e->d_px, e->d_nn, e->d_head, e->d_nl, e->d_pos_build, e->d_flag, e->d_grid,
                    e->d_spec, e->d_V, e->d_u, e->d_y, e->d_alpha, e->d_beta, e->d_coef, e->d_partials, e->d_counter,
                    e->d_vel_work, e->d_hpos, e->d_hF, e->d_himage, e->d_red2, e->d_org, e->d_worg, e->d_wcell_of, e->d_wcount,
                    e->d_wstart, e->d_wperm, e->d_wid, e->d_wpos, e->d_wF, e->d_wwt, e->d_wrecs, e->d_wtmp, e->d_ell, e->d_nn_act, e->d_nl_act};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (e->h_flag) cudaFreeHost(e->h_flag);
    if (e->h_stepdev) cudaFreeHost(e->h_stepdev);
    if (e->d_stepdev) cudaFree(e->d_stepdev);
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->shard) shard_free(e->shard);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->stream_h2d) cudaStreamDestroy(e->stream_h2d);
    if (e->stream_d2h) cudaStreamDestroy(e->stream_d2h);
    if (e->ev_F) cudaEventDestroy(e->ev_F);
    if (e->ev_img) cudaEventDestroy(e->ev_img);
    if (e->ev_step) cudaEventDestroy(e->ev_step);
    if (e->ev_out) cudaEventDestroy(e->ev_out);
    for (int k = 0; k < 2; ++k) {
        if (e->ev_Fnext[k]) cudaEventDestroy(e->ev_Fnext[k]);
        if (e->ev_Fcons[k]) cudaEventDestroy(e->ev_Fcons[k]);
        if (e->d_hF_next[k]) cudaFree(e->d_hF_next[k]);
    }
    if (e->h_nlinfo) cudaFreeHost(e->h_nlinfo);
    if (e->d_nlinfo) cudaFree(e->d_nlinfo);
    if (e->h_ab) cudaFreeHost(e->h_ab);
    if (e->flag_event) cudaEventDestroy(e->flag_event);
    if (e->prof_pool) { for (cudaEvent_t ev : *e->prof_pool) cudaEventDestroy(ev); delete e->prof_pool; }
    delete e->prof_spans;
    delete e;
}

extern "C" int pse_get_params(const pse_engine* e, pse_params* out) {
    if (!e || !out) return PSE_EINVAL;
    *out = e->prm;
    return PSE_OK;
}
extern "C" int pse_set_box(pse_engine* e, const pse_box* b) {
    if (!e || !b) return PSE_EINVAL;
    if (b->Lx != e->cfg.box.Lx || b->Ly != e->cfg.box.Ly || b->Lz != e->cfg.box.Lz)
        return fail(e, PSE_EINVAL, "pse_set_box: only the tilt may change (the reference sizes the grid once, Stokes.cc:139)");
    refresh_box(e, *b);
    return PSE_OK;
}
__global__ void wrap_positions_kernel(float4* __restrict__ pos, int3* __restrict__ image, uint32_t N, PseBox box) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 p = pos[i];
    float3 w = make_float3(p.x, p.y, p.z);
    int3 im = image ? image[i] : make_int3(0, 0, 0);
    box.wrap(w, im);
    pos[i] = make_float4(w.x, w.y, w.z, p.w);
    if (image) image[i] = im;
}
extern "C" int pse_wrap_positions(pse_engine* e, float4* d_pos, int3* d_image) {
    if (!e || !d_pos) return PSE_EINVAL;
    wrap_positions_kernel<<<(e->N + 255) / 256, 256, 0, e->stream>>>(d_pos, d_image, e->N, e->box); e->launches++;
    CK(cudaGetLastError());
    return PSE_OK;
}
extern "C" int pse_set_temperature(pse_engine* e, float T) {
    if (!e || T < 0.f) return PSE_EINVAL;
    e->cfg.T = T;
    return PSE_OK;
}
extern "C" int pse_set_lanczos_m(pse_engine* e, int m) {
    if (!e || m < 1 || m > LANCZOS_M_MAX) return PSE_EINVAL;
    e->m_lanczos = m;
    return PSE_OK;
}
extern "C" int pse_get_lanczos_m(const pse_engine* e) { return e ? e->m_lanczos : PSE_EINVAL; }
extern "C" int pse_get_stats(pse_engine* e, pse_stats* out) {
    if (!e || !out) return PSE_EINVAL;
    out->nnz = e->nnz;
    out->nnz_active = e->nnz;
    if (e->prune && e->pruned_valid && e->nlist_valid) {
        CK(cudaMemsetAsync(e->d_nlinfo, 0, sizeof(unsigned long long), e->stream));
        nnz_kernel<<<e->num_sms, 256, 0, e->stream>>>(e->d_nn_act + e->row0, e->row1 - e->row0, e->d_nlinfo);
        CK(cudaMemcpyAsync(e->h_nlinfo, e->d_nlinfo, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        out->nnz_active = e->h_nlinfo[0];
    } out->kernel_launches = e->launches; out->fft_execs = e->fft_execs; out->graph_launches = e->graph_launches;
    out->nlist_builds = e->nlist_builds; out->lanczos_m = e->m_lanczos; out->lanczos_stepnorm = e->last_stepnorm;
    return PSE_OK;
}

// ---- lazily allocated big buffers ---------------------------------------------------------------------
static int ensure_wave_buffers(pse_engine* e) {
    if (e->d_grid) return PSE_OK;
    if (e->shard) return fail(e, PSE_ECUDA, "slab buffers missing");   // allocated by pse_shard_init
    CK(cudaMalloc(&e->d_grid, sizeof(float) * 3 * e->G));
    CK(cudaMalloc(&e->d_spec, sizeof(float2) * 3 * e->Gh));
    e->grid_planes = e->wp.Nx;
    return PSE_OK;
}
static int ensure_krylov(pse_engine* e) {
    const size_t rows = e->row1 - e->row0;
    if (e->d_V && rows <= e->v_rows) return PSE_OK;
    if (e->d_V) { CK(cudaStreamSynchronize(e->stream)); cudaFree(e->d_V); e->d_V = nullptr; }
    e->v_rows = e->shard ? (size_t)(rows * 1.15) + 4096 : e->N;
    CK(cudaMalloc(&e->d_V, sizeof(float4) * e->v_rows * LANCZOS_M_MAX));
    e->nl_gen++;   // a captured graph holds the old pointer
    return PSE_OK;
}
static int shard_update_geometry(pse_engine* e);
static int ensure_csr(pse_engine* e);

// ---- binning + neighbour list ------------------------------------------------------------------
extern "C" int pse_build_neighbors(pse_engine* e, const float4* d_pos) {
    if (!e || !d_pos) return PSE_EINVAL;
    const uint32_t N = e->N;
    cudaStream_t st = e->stream;
    setup_cell_grid(e);
    const CellGrid cg = e->cg;
    const uint32_t ncell = cg.ncell;
    ProfScope* ps = new ProfScope(e, PH_BIN);
    CK(cudaMemsetAsync(e->d_cell_count, 0, sizeof(uint32_t) * (ncell + 1), st));
    cell_id_kernel<<<nblk(N, 256), 256, 0, st>>>(d_pos, N, e->box, cg, e->d_cell_of, e->d_cell_count); LAUNCHED(e);
    CKRC(exclusive_scan(e, e->d_cell_count, e->d_cell_start, ncell + 1, e->d_scan_tmp));
    CK(cudaMemsetAsync(e->d_cell_count, 0, sizeof(uint32_t) * (ncell + 1), st));
    cell_fill_kernel<<<nblk(N, 256), 256, 0, st>>>(e->d_cell_of, N, e->d_cell_start, e->d_cell_count, e->d_perm); LAUNCHED(e);
    cell_sort_kernel<<<nblk(ncell, 128), 128, 0, st>>>(e->d_cell_start, ncell, e->d_perm); LAUNCHED(e);
    invert_perm_kernel<<<nblk(N, 256), 256, 0, st>>>(e->d_perm, N, e->d_slot_of); LAUNCHED(e);
    gather_pos_kernel<<<nblk(N, 256), 256, 0, st>>>(d_pos, e->d_perm, N, e->d_spos, (float4*)e->d_px); LAUNCHED(e);
    // slab-decomposed: the binning above is replicated (identical on every rank); everything below is for the own rows
    if (e->shard) CKRC(shard_update_geometry(e));
    const uint32_t r0 = e->row0, r1 = e->row1, nrows = r1 - r0;

    delete ps;
    ps = new ProfScope(e, PH_NLIST);
    const float rl2 = e->rlist * e->rlist;
    if (e->nl_stride == 0) {
        // expected neighbours per particle at uniform density, with head room for fluctuations
        const double dens = (double)N / ((double)e->box.Lx * e->box.Ly * e->box.Lz);
        const double expect = 4.0 / 3.0 * 3.14159265358979 * (double)e->rlist * e->rlist * e->rlist * dens;
        e->nl_stride = (uint32_t)(((size_t)(1.5 * expect + 24.0) + 7) / 8 * 8);
        if (e->nl_stride > N) e->nl_stride = ((N + 7) / 8) * 8;
    }
    CK(cudaMemsetAsync(e->d_nn + r1, 0, sizeof(uint32_t), st));   // (entry r1 closes the scan; with r1 < N it is another rank's row)
    for (int attempt = 0; attempt < 3; ++attempt) {
        const size_t need = (size_t)std::max<uint32_t>(nrows, 1) * e->nl_stride;
        if (need > e->ell_cap) {
            if (e->d_ell) cudaFree(e->d_ell);
            e->d_ell = nullptr;
            e->ell_cap = need + need / 8;
            CK(cudaMalloc(&e->d_ell, sizeof(uint32_t) * e->ell_cap));
        }
        CK(cudaMemsetAsync(e->d_nlinfo, 0, 2 * sizeof(unsigned long long), st));
        if (nrows) {
            nlist_kernel<<<nblk(nrows, 128), 128, 0, st>>>(e->d_spos, r1, e->box, cg, e->d_cell_start, rl2, e->nl_stride, e->d_nn, e->d_ell,
                                                           (uint32_t*)(e->d_nlinfo + 1), r0); LAUNCHED(e);
        }
        CKRC(exclusive_scan(e, e->d_nn + r0, e->d_head + r0, nrows + 1, e->d_scan_tmp));
        words2_copy_kernel<<<1, 32, 0, st>>>(WORDS(e->h_nlinfo), CWORDS(e->d_nlinfo), 4, e->h_flag, e->d_head + r1, 1); LAUNCHED(e);
        CK(cudaStreamSynchronize(st));
        const uint32_t max_nn = (uint32_t)e->h_nlinfo[1];
        e->nnz = *e->h_flag;
        if (max_nn <= e->nl_stride) break;
        e->nl_stride = ((max_nn + 16 + 7) / 8) * 8;  // a row overflowed its stride: grow and search again
        if (attempt == 2) return fail(e, PSE_ENOMEM, "neighbour rows keep overflowing (max %u)", max_nn);
    }
    if (e->nnz > e->nl_cap) {
        if (e->d_nl) cudaFree(e->d_nl);
        e->d_nl = nullptr;
        e->nl_cap = (size_t)(e->nnz * 1.2) + 1024;
        CK(cudaMalloc(&e->d_nl, sizeof(uint32_t) * e->nl_cap));
        e->nl_gen++;
    }
    if (e->prune && e->nl_cap > e->nl_act_cap) {
        if (e->d_nl_act) cudaFree(e->d_nl_act);
        e->d_nl_act = nullptr;
        e->nl_act_cap = e->nl_cap;
        CK(cudaMalloc(&e->d_nl_act, sizeof(uint32_t) * e->nl_act_cap));
        e->nl_gen++;
    }
    // The fixed-stride search output is what the per-step pruning reads; the packed (CSR) copy of the buffered list is only
    // made when somebody walks it directly (unpruned SpMV, pair forces, export): ensure_csr.
    e->csr_valid = false;
    if (!e->prune) CKRC(ensure_csr(e));
    copy_f4_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->d_pos_build, e->d_spos, N); LAUNCHED(e);   // slot order
    delete ps;
    e->xy_build = e->box.xy;
    e->xy_spos = e->box.xy;
    e->pruned_valid = false;   // (new slot order: nothing derived from the old one survives)
    e->wbin_valid = false;
    e->nlist_valid = true;
    e->flag_pending = false;
    e->nlist_builds++;
    CK(cudaGetLastError());
    return PSE_OK;
}

static int ensure_csr(pse_engine* e) {
    if (e->csr_valid) return PSE_OK;
    const uint32_t nrows = e->row1 - e->row0;
    if (nrows) {
        compact_rows_kernel<<<nblk((size_t)nrows * 8, 256), 256, 0, e->stream>>>(e->d_ell, e->nl_stride, e->d_nn, e->d_head, e->row1, e->d_nl, e->row0); LAUNCHED(e);
    }
    e->csr_valid = true;
    return PSE_OK;
}

// list still valid for positions d_pos?  2*max displacement + tilt drift must stay inside the buffer
static bool stale_from_bits(const pse_engine* e, uint32_t bits) {
    float r2;
    memcpy(&r2, &bits, 4);
    const float drift = fabsf(e->box.xy - e->xy_build) * e->rlist;
    if (2.f * sqrtf(r2) + drift > e->cfg.r_buff) return true;
    // slab-decomposed: a rank's particles must keep their Gaussian supports inside its local grid buffer, whose halo was
    // sized for a drift of SHARD_DRIFT_NODES nodes of the sheared frame at |y| = Ly / 2 (shard_static_geometry)
    if (e->shard && fabsf(e->box.xy - e->xy_build) * 0.5f * e->box.Ly > SHARD_DRIFT_NODES * e->wp.hx) return true;
    return false;
}
static int launch_disp_check(pse_engine* e, const float4* d_pos) {
    ProfScope ps(e, PH_REORDER);
    CK(cudaMemsetAsync(e->d_flag, 0, sizeof(uint32_t), e->stream));
    CK(cudaMemsetAsync(e->d_flag + 2, 0, sizeof(uint32_t), e->stream));
    check_and_gather_kernel<<<nblk(e->N, 256), 256, 0, e->stream>>>(d_pos, e->d_perm, e->N, e->box, e->d_pos_build, e->d_spos, (float4*)e->d_px, e->d_flag); LAUNCHED(e);
    words_copy_kernel<<<1, 32, 0, e->stream>>>(e->h_flag, e->d_flag, 3); LAUNCHED(e);
    CK(cudaEventRecord(e->flag_event, e->stream));
    e->flag_pending = true;
    return PSE_OK;
}
// make cells / list / slot-ordered positions current for d_pos
static int ensure_neighbors(pse_engine* e, const float4* d_pos) {
    bool rebuild = !e->nlist_valid;
    if (!rebuild) {
        // host-pipelined runs own their state: the check for these positions was issued at the end of the previous step, before
        // its result started down the PCIe link, so the flags are already on the host (no round trip at the head of the step)
        const bool prechecked = e->flag_pending && e->precheck_pos == d_pos && e->precheck_xy == e->box.xy;
        e->precheck_pos = nullptr;
        if (!prechecked) CKRC(launch_disp_check(e, d_pos));
        CK(cudaEventSynchronize(e->flag_event));
        e->flag_pending = false;
        if (e->h_flag[1] & 2u) return fail(e, PSE_ECUDA, "slab decomposition: a peer rank did not arrive at a synchronisation point within 10 s");
        if (e->h_flag[1]) return fail(e, PSE_EINVAL, "slab decomposition: a particle's Gaussian support left the rank's grid buffer (halo too thin)");
        rebuild = stale_from_bits(e, *e->h_flag);
    }
    if (rebuild) {
        CKRC(pse_build_neighbors(e, d_pos));
    } else if (e->h_flag[2] == 0 && e->box.xy == e->xy_spos && e->reuse_static) {
        // nobody moved and the box is the one of the previous call: the slot-ordered positions, the pruned list and the
        // wave-space binning / factor rows of that call are still exact (the operator applied again at a fixed configuration)
        return PSE_OK;
    }
    // (no rebuild: the check above already brought the slot-ordered positions up to date)
    e->xy_spos = e->box.xy;
    e->pruned_valid = false;
    e->wbin_valid = false;
    return PSE_OK;
}

// neighbour rows the SpMV walks: the buffered list, or (default) its per-step pruning to r < r_cut
static int ensure_pruned(pse_engine* e) {
    if (!e->prune || e->pruned_valid) return PSE_OK;
    ProfScope ps(e, PH_PRUNE);
    prune_kernel<<<persistent_grid(e, nblk((size_t)(e->row1 - e->row0) * 8, 256), 8), 256, 0, e->stream>>>(e->d_spos, e->row1, e->d_nn, e->d_head, e->d_ell, e->rp,
                                                                                         e->box, e->d_nn_act, e->d_nl_act, e->row0, e->nl_stride); LAUNCHED(e);
    e->pruned_valid = true;
    return PSE_OK;
}

extern "C" int pse_neighbor_list(pse_engine* e, uint32_t* d_n_neigh, uint32_t* d_headlist, uint32_t* d_nlist,
                                 size_t cap, size_t* nnz_out) {
    if (!e) return PSE_EINVAL;
    if (!e->nlist_valid) return fail(e, PSE_EINVAL, "pse_neighbor_list: no list built yet");
    if (e->shard && e->shard->world > 1) return fail(e, PSE_EINVAL, "pse_neighbor_list: a slab-decomposed engine holds the rows of its own slab only");
    if (nnz_out) *nnz_out = e->nnz;
    if (!d_n_neigh && !d_headlist && !d_nlist) return PSE_OK;
    CKRC(ensure_csr(e));
    const uint32_t N = e->N;
    uint32_t *nn_id = nullptr, *head_id = nullptr;
    CK(cudaMalloc(&nn_id, sizeof(uint32_t) * (N + 1)));
    CK(cudaMalloc(&head_id, sizeof(uint32_t) * (N + 1)));
    CK(cudaMemsetAsync(nn_id + N, 0, sizeof(uint32_t), e->stream));
    export_counts_kernel<<<nblk(N, 256), 256, 0, e->stream>>>(e->d_nn, e->d_perm, N, nn_id); LAUNCHED(e);
    int rc = exclusive_scan(e, nn_id, head_id, N + 1, e->d_scan_tmp);
    if (rc == PSE_OK && d_nlist) {
        if (cap < e->nnz) rc = fail(e, PSE_ECAPACITY, "pse_neighbor_list: need %llu entries", (unsigned long long)e->nnz);
        else { export_rows_kernel<<<nblk(N, 128), 128, 0, e->stream>>>(e->d_nn, e->d_head, e->d_nl, e->d_perm, N, head_id, d_nlist); LAUNCHED(e); }
    }
    if (rc == PSE_OK && d_n_neigh) cudaMemcpyAsync(d_n_neigh, nn_id, sizeof(uint32_t) * N, cudaMemcpyDeviceToDevice, e->stream);
    if (rc == PSE_OK && d_headlist) cudaMemcpyAsync(d_headlist, head_id, sizeof(uint32_t) * N, cudaMemcpyDeviceToDevice, e->stream);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    cudaFree(nn_id); cudaFree(head_id);
    if (ce != cudaSuccess) return fail(e, PSE_ECUDA, "pse_neighbor_list: %s", cudaGetErrorString(ce));
    return rc;
}

extern "C" int pse_grid_index(pse_engine* e, const float4* d_pos, int3* d_out) {
    if (!e || !d_pos || !d_out) return PSE_EINVAL;
    grid_index_kernel<<<nblk(e->N, 256), 256, 0, e->stream>>>(d_pos, e->N, e->box, e->wp, d_out); LAUNCHED(e);
    CK(cudaGetLastError());
    return PSE_OK;
}

// ---- building blocks (slot order) -----------------------------------------------------------------
template <int TPP, int MODE, bool PRUNED>
static void launch_spmv_tpp(pse_engine* e, float4* y, const LanczosArgs& la) {
    const unsigned int work = nblk((size_t)(e->row1 - e->row0) * TPP, 256);   // rows [row0, row1): all of them unless slab-decomposed
    const uint32_t* nn = PRUNED ? e->d_nn_act : e->d_nn;
    const uint32_t* nl = PRUNED ? e->d_nl_act : e->d_nl;
    if (e->spmv_table_mode == TABLE_POLY) {
        spmv_kernel<TPP, MODE, TABLE_POLY, PRUNED><<<persistent_grid(e, work, e->spmv_bps > 0 ? e->spmv_bps : 8), 256, 0, e->stream>>>(
            e->d_px, y, e->row1, nn, e->d_head, nl, e->d_table, e->cheb, e->rp, e->box, la, e->row0);
    } else if (e->spmv_table_mode == TABLE_SHARED) {
        const size_t sm = spmv_table_smem(e);
        int bps = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (sm + 1024)));
        if (e->spmv_bps > 0) bps = std::min(bps, e->spmv_bps);
        cudaFuncSetAttribute(spmv_kernel<TPP, MODE, TABLE_SHARED, PRUNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        spmv_kernel<TPP, MODE, TABLE_SHARED, PRUNED><<<persistent_grid(e, work, bps), 256, sm, e->stream>>>(
            e->d_px, y, e->row1, nn, e->d_head, nl, e->d_table, e->cheb, e->rp, e->box, la, e->row0);
    } else {
        spmv_kernel<TPP, MODE, TABLE_GLOBAL, PRUNED><<<persistent_grid(e, work, e->spmv_bps > 0 ? e->spmv_bps : 8), 256, 0, e->stream>>>(
            e->d_px, y, e->row1, nn, e->d_head, nl, e->d_table, e->cheb, e->rp, e->box, la, e->row0);
    }
    LAUNCHED(e);
}
// first Lanczos product of a full step with M_real F riding along (poly / global table modes, pruned lists)
static bool launch_spmv_dual(pse_engine* e, float4* y, const LanczosArgs& la, const float4* x2, float4* y2) {
    if (!e->prune || !e->spmv_dual || e->spmv_tpp != 4 || e->spmv_table_mode == TABLE_SHARED) return false;
    const unsigned int work = nblk((size_t)(e->row1 - e->row0) * 4, 256);
    const unsigned int grid = persistent_grid(e, work, e->spmv_bps > 0 ? e->spmv_bps : 8);
    if (e->spmv_table_mode == TABLE_POLY)
        spmv_kernel<4, SPMV_LANCZOS, TABLE_POLY, true, true><<<grid, 256, 0, e->stream>>>(e->d_px, y, e->row1, e->d_nn_act, e->d_head, e->d_nl_act, e->d_table,
                                                                                     e->cheb, e->rp, e->box, la, e->row0, x2, y2);
    else
        spmv_kernel<4, SPMV_LANCZOS, TABLE_GLOBAL, true, true><<<grid, 256, 0, e->stream>>>(e->d_px, y, e->row1, e->d_nn_act, e->d_head, e->d_nl_act, e->d_table,
                                                                                       e->cheb, e->rp, e->box, la, e->row0, x2, y2);
    LAUNCHED(e);
    return true;
}
template <int MODE>
static void launch_spmv(pse_engine* e, float4* y, const LanczosArgs& la) {
    switch (e->spmv_tpp) {
        case 8: e->prune ? launch_spmv_tpp<8, MODE, true>(e, y, la) : launch_spmv_tpp<8, MODE, false>(e, y, la); break;
        default: e->prune ? launch_spmv_tpp<4, MODE, true>(e, y, la) : launch_spmv_tpp<4, MODE, false>(e, y, la); break;
    }
}

static bool spmv_dual_available(const pse_engine* e) {
    return e->prune && e->spmv_dual && e->spmv_tpp == 4 && e->spmv_table_mode != TABLE_SHARED;
}
static int run_spmv_plain(pse_engine* e, float4* y) {
    CKRC(ensure_pruned(e));
    ProfScope ps(e, PH_SPMV);
    LanczosArgs la = {};
    launch_spmv<SPMV_PLAIN>(e, y, la);
    return PSE_OK;
}

// Bin the (slot-ordered) particles by the tile of their support origin and gather the W-order records and factor rows:
// everything that depends on the positions only, valid until the positions change (ensure_neighbors).
static int run_wbin(pse_engine* e) {
    if (e->wbin_valid) return PSE_OK;
    ProfScope ps(e, PH_WBIN);
    cudaStream_t st = e->stream;
    const uint32_t N = e->N, nt = e->tg.ntile;
    CK(cudaMemsetAsync(e->d_wcount, 0, sizeof(uint32_t) * (nt + 1), st));
    wbin_kernel<<<nblk(N, 256), 256, 0, st>>>(e->d_spos, N, e->box, e->wp, e->tg, e->d_org, e->d_wcell_of, e->d_wcount); LAUNCHED(e);
    CKRC(exclusive_scan(e, e->d_wcount, e->d_wstart, nt + 1, e->d_scan_tmp));
    CK(cudaMemsetAsync(e->d_wcount, 0, sizeof(uint32_t) * (nt + 1), st));
    // unordered fill into scratch (d_wcell_of is free again after the fill reads it), then rank sort per tile
    cell_fill_kernel<<<nblk(N, 256), 256, 0, st>>>(e->d_wcell_of, N, e->d_wstart, e->d_wcount, e->d_wtmp); LAUNCHED(e);
    cell_sort_block_kernel<<<nt, 128, 0, st>>>(e->d_wstart, e->d_wtmp, e->d_wperm); LAUNCHED(e);
    if (e->wave_v2) {   // headers + factor rows of the W records in one kernel
        launch_wrecords(e->wp.P, st, e->d_spos, e->d_org, e->d_wperm, e->d_perm, N, e->box, e->wp, e->tg, e->d_wrecs); LAUNCHED(e);
    } else {
        wgather_kernel<<<nblk(N, 256), 256, 0, st>>>(e->d_spos, nullptr, e->d_org, e->d_wperm, e->d_perm, N, e->d_wpos, e->d_wF, e->d_worg, e->d_wid, e->tg,
                                                     nullptr, nullptr, 1); LAUNCHED(e);
        launch_wweights(e->wp.P, st, e->d_wpos, e->d_worg, N, e->box, e->wp, e->d_wwt); LAUNCHED(e);
    }
    e->wbin_valid = true;
    return PSE_OK;
}
// the forces of the call into the W records (spreading reads them from there)
static int run_wforce(pse_engine* e, const float4* sF) {
    ProfScope ps(e, PH_WBIN);
    wgather_kernel<<<nblk(e->N, 256), 256, 0, e->stream>>>(e->d_spos, sF, e->d_org, e->d_wperm, e->d_perm, e->N, e->d_wpos, e->d_wF, e->d_worg, e->d_wid, e->tg,
                                                        reinterpret_cast<int4*>(e->d_wrecs), nullptr, 2); LAUNCHED(e);
    return PSE_OK;
}

// wave-space pipeline on slot-ordered (spos, sF): result scattered into U (particle-id order)
static int run_wave(pse_engine* e, const float4* sF, float4* U, int accumulate, bool det, bool noise, const float* d_u_grid) {
    cudaStream_t st = e->stream;
    const int P = e->wp.P;
    if (e->tiled) {
        CKRC(run_wbin(e));
        if (det) CKRC(run_wforce(e, sF));
    }
    if (det) {
        {
        ProfScope ps(e, PH_SPREAD);
        if (e->wave_v2) {
            CK(cudaMemsetAsync(e->d_grid, 0, sizeof(float) * 3 * e->G, st));
            launch_spread2(P, e->spread_var, st, e->d_wrecs, e->d_wstart, e->wp, e->tg, e->d_grid); LAUNCHED(e);
        } else if (e->tiled) {
            launch_spread_tile(P, st, e->d_wF, e->d_worg, e->d_wwt, e->d_wstart, e->wp, e->tg, e->d_grid); LAUNCHED(e);
        } else {
            CK(cudaMemsetAsync(e->d_grid, 0, sizeof(float) * 3 * e->G, st));
            spread_scatter_kernel<<<nblk((size_t)e->N * 32, 256), 256, 0, st>>>(e->d_spos, sF, e->N, e->box, e->wp, e->d_grid); LAUNCHED(e);
        }
        }
        ProfScope ps(e, PH_FFT_FWD);
        if (e->own_fft) {
            const WaveParams& wp = e->wp;
            const uint32_t nrows = 3u * wp.Nx * wp.Ny;
            fft_z_forward_kernel<<<nblk(nrows, 2 * FFT_Z_COLS), FFT_THREADS, fft_smem_bytes(wp.Nz, FFT_Z_COLS + 1), st>>>(
                e->d_grid, e->d_spec, e->fft_ax[2], nrows, wp.Nzp); LAUNCHED(e);
            fft_y_kernel<false><<<dim3(nblk(wp.Nzh, FFT_Y_COLS), 3 * wp.Nx), FFT_THREADS, fft_smem_bytes(wp.Ny, FFT_Y_CP), st>>>(
                e->d_spec, e->fft_ax[1], wp.Nzh, wp.Nzp); LAUNCHED(e);
        } else {
            CKFFT(cufftExecR2C(e->plan_f, e->d_grid, (cufftComplex*)e->d_spec));
        }
        e->fft_execs++;
    }
    if (e->own_fft) {
        const WaveParams& wp = e->wp;
        {
            ProfScope ps(e, PH_SCALE);  // x forward + scaling + x inverse
            fft_x_scale_kernel<<<dim3(nblk(wp.Nzh, FFT_X_COLS), wp.Ny), FFT_THREADS, fft_smem_bytes(wp.Nx, FFT_X_CP), st>>>(
                e->d_spec, e->fft_ax[0], e->fft_ax[1].freq_of, e->wp, e->box, det ? 1 : 0, noise ? 1 : 0, e->d_stepdev, d_u_grid); LAUNCHED(e);
        }
        ProfScope ps(e, PH_FFT_INV);
        const uint32_t nrows = 3u * wp.Nx * wp.Ny;
        fft_y_kernel<true><<<dim3(nblk(wp.Nzh, FFT_Y_COLS), 3 * wp.Nx), FFT_THREADS, fft_smem_bytes(wp.Ny, FFT_Y_CP), st>>>(
            e->d_spec, e->fft_ax[1], wp.Nzh, wp.Nzp); LAUNCHED(e);
        fft_z_inverse_kernel<<<nblk(nrows, 2 * FFT_Z_COLS), FFT_THREADS, fft_smem_bytes(wp.Nz, FFT_Z_COLS + 1), st>>>(
            e->d_spec, e->d_grid, e->fft_ax[2], nrows, wp.Nzp); LAUNCHED(e);
        e->fft_execs++;
    } else {
        {
            ProfScope ps(e, PH_SCALE);
            scale_kernel<<<dim3(e->wp.Ny, e->wp.Nx), 128, 0, st>>>(e->d_spec, e->wp, e->box, det ? 1 : 0, noise ? 1 : 0, e->d_stepdev, d_u_grid); LAUNCHED(e);
        }
        ProfScope ps(e, PH_FFT_INV);
        CKFFT(cufftExecC2R(e->plan_b, (cufftComplex*)e->d_spec, e->d_grid)); e->fft_execs++;
    }
    ProfScope ps(e, PH_INTERP);
    if (e->wave_v2) {
        launch_interp2(P, st, e->d_wrecs, e->d_wstart, e->wp, e->tg, e->d_grid, U, accumulate);
        LAUNCHED(e);
    } else if (e->tiled) {
        launch_interp_tile(P, st, e->d_worg, e->d_wwt, e->d_wstart, e->d_wid, e->wp, e->tg, e->d_grid, U, accumulate);
        LAUNCHED(e);
    } else {
        interp_warp_kernel<<<nblk((size_t)e->N * 32, 256), 256, 0, st>>>(e->d_spos, e->N, e->box, e->wp, e->d_grid, e->d_perm, U, accumulate); LAUNCHED(e);
    }
    return PSE_OK;
}

// one Lanczos iteration j (two kernels).  Slab-decomposed engines work on rows [row0, row1) of every vector (V holds the own
// rows only) and add two exchanges: the boundary rows of u_j from both neighbours before the product, and the all-reduce
// of (alpha_j, |y|^2) after it.
static int shard_exchange_px(pse_engine* e);
static int shard_allreduce2(pse_engine* e);
static int lanczos_iteration(pse_engine* e, int j, bool dual = false) {
    const uint32_t r0 = e->row0, r1 = e->row1;
    const bool multi = e->shard && e->shard->world > 1;
    float4* Vj = e->d_V + (size_t)j * e->v_rows - r0;   // indexed by row
    LanczosArgs la;
    la.beta_j = e->d_beta + j;
    la.v_prev = j > 0 ? e->d_V + (size_t)(j - 1) * e->v_rows - r0 : nullptr;
    la.v_out = Vj;
    la.alpha_out = e->d_alpha + j;
    la.partials = e->d_partials;
    la.counter = e->d_counter;
    la.first = j == 0;
    if (multi && j > 0) CKRC(shard_exchange_px(e));   // (u_0 = psi is generated for every row on every rank)
    {
    ProfScope ps(e, PH_LANCZOS_SPMV);
    if (!(dual && launch_spmv_dual(e, e->d_y, la, e->d_sx, e->d_sy))) launch_spmv<SPMV_LANCZOS>(e, e->d_y, la);
    }
    if (multi) {
        {
            ProfScope ps(e, PH_LANCZOS_VEC);
            const unsigned int gd = std::min<unsigned int>(DOTS_BLOCKS, std::max(1u, nblk(r1 - r0, 256)));
            lanczos_dots_kernel<<<gd, 256, 0, e->stream>>>(e->d_y, Vj, r0, r1, e->d_red2 + 4, e->d_counter + 1, e->d_red2); LAUNCHED(e);
        }
        CKRC(shard_allreduce2(e));
    }
    ProfScope ps(e, PH_LANCZOS_VEC);
    lanczos_update_kernel<<<persistent_grid(e, nblk(r1 - r0, 256), 8), 256, 0, e->stream>>>(e->d_y, Vj, e->d_px, r1, e->d_alpha + j, e->d_beta + j + 1,
                                                               e->d_partials, e->d_counter, r0, multi ? e->d_red2 : nullptr, e->d_alpha + j);
    LAUNCHED(e);
    return PSE_OK;
}

static int solve_coeffs(pse_engine* e, int m, const float* alpha, const float* beta, std::vector<double>& c) {
    std::vector<double> d(m), o(m > 1 ? m - 1 : 1);
    for (int i = 0; i < m; ++i) d[i] = alpha[i];
    for (int i = 0; i + 1 < m; ++i) o[i] = beta[i + 1];
    c.assign(m, 0.0);
    double lmin = 0;
    int rc = pse_tridiag_sqrt_e1(m, d.data(), o.data(), c.data(), &lmin);
    if (rc != PSE_OK)
        return fail(e, PSE_EEIGEN, "Lanczos tridiagonal matrix (m = %d) is not positive definite (lambda_min = %g)", m, lmin);
    return PSE_OK;
}

// U[perm] (+)= sqrt(2T/dt) * M_real^{1/2} psi, psi drawn per particle id (PSEv1/Brownian.cu:357-765).
// lanczos_batch: everything up to the first host decision — psi, |psi|, the first m_in - 1 iterations and the
// read-back of alpha/beta.  No host synchronisation inside, so it can be part of the captured step graph.
static int lanczos_first_m(const pse_engine* e) {
    const int m = e->m_lanczos - 1;  // PSEv1/Brownian.cu:465-466
    return m < 1 ? 1 : m;
}
// The reference always takes at least one adaptive iteration after those m_in - 1 (its step norm starts at 1, :606-610), so
// that iteration is issued with the batch: in the steady state (m unchanged from step to step) the whole solve needs ONE
// host synchronisation instead of two, and all of its products sit inside the captured graph.
static int lanczos_batch_size(const pse_engine* e) { return std::min(lanczos_first_m(e) + 1, LANCZOS_M_MAX); }
static int lanczos_batch(pse_engine* e, const float* d_u_particles, int m, bool dual = false) {
    CKRC(ensure_pruned(e));
    const uint32_t N = e->N;
    cudaStream_t st = e->stream;
    {
    ProfScope ps(e, PH_LANCZOS_VEC);
    psi_kernel<<<nblk(N, 256), 256, 0, st>>>((float4*)e->d_px + 1, 2, e->d_perm, N, d_u_particles, e->d_stepdev); LAUNCHED(e);
    dot_px_kernel<<<persistent_grid(e, nblk(N, 256), 8), 256, 0, st>>>(e->d_px, N, e->d_beta, e->d_partials, e->d_counter, true); LAUNCHED(e);
    }
    float* alpha = e->h_ab;
    float* beta = e->h_ab + LANCZOS_M_MAX + 1;
    for (int j = 0; j < m; ++j) CKRC(lanczos_iteration(e, j, dual && j == 0));
    words2_copy_kernel<<<1, 128, 0, st>>>(WORDS(alpha), CWORDS(e->d_alpha), m, WORDS(beta), CWORDS(e->d_beta), m + 1); LAUNCHED(e);
    return PSE_OK;
}

// host side: tridiagonal solves, adaptive iterations until the step norm drops below `error`, final combination
// (slab-decomposed: U is the slot-ordered velocity buffer and `perm` is null)
static int lanczos_finish(pse_engine* e, float4* U, int accumulate, int m_have /* iterations lanczos_batch issued */, int* m_out, const float4* ydet,
                          const uint32_t* perm) {
    cudaStream_t st = e->stream;
    float* alpha = e->h_ab;
    float* beta = e->h_ab + LANCZOS_M_MAX + 1;
    CK(cudaStreamSynchronize(st));
    int m = std::min(lanczos_first_m(e), m_have);
    for (int j = 0; j < m; ++j)
        if (beta[j + 1] < 1e-8f) { m = j > 0 ? j : 1; break; }  // breakdown, PSEv1/Brownian.cu:507-510
    std::vector<double> c, c_prev;
    CKRC(solve_coeffs(e, m, alpha, beta, c));
    c_prev = c;
    const double rho = alpha[0];  // psi.M.psi/|psi|^2 == alpha_0 (PSEv1/Brownian.cu:452-457 spends an extra SpMV on it)
    double stepnorm = 1.0;
    const bool broke = beta[m] < 1e-8f;
    while (!broke && stepnorm > e->cfg.error && m < LANCZOS_M_MAX) {  // PSEv1/Brownian.cu:606-736
        ++m;
        const int j = m - 1;
        if (j >= m_have) {   // (the first adaptive iteration came with the batch)
            CKRC(lanczos_iteration(e, j));
            words2_copy_kernel<<<1, 32, 0, st>>>(WORDS(alpha + j), CWORDS(e->d_alpha + j), 1, WORDS(beta + j + 1), CWORDS(e->d_beta + j + 1), 1); LAUNCHED(e);
            CK(cudaStreamSynchronize(st));
            m_have = m;
        }
        if (beta[j + 1] < 1e-8f) { m = j; break; }
        CKRC(solve_coeffs(e, m, alpha, beta, c));
        // ||V c_m - V c_{m-1}|| = ||c_m - [c_{m-1}; 0]|| for an orthonormal basis: no N-vector pass
        double s = 0.0;
        for (int i = 0; i < m; ++i) {
            double d = c[i] - (i < (int)c_prev.size() ? c_prev[i] : 0.0);
            s += d * d;
        }
        stepnorm = sqrt(s / rho);
        c_prev = c;
    }
    c = c_prev;
    m = (int)c.size();
    {
        CoefArg ca;   // by value in the launch: no copy engine, no host buffer to keep alive
        for (int i = 0; i < m; ++i) ca.c[i] = (float)c[i];
        coef_store_kernel<<<1, 128, 0, st>>>(e->d_coef, ca, m); LAUNCHED(e);
    }
    const float thermal = sqrtf((float)(2.0 * e->cfg.T / e->cfg.dt));  // PSEv1/Brownian.cu:739
    ProfScope ps(e, PH_COMBINE);
    if (e->row1 > e->row0) {
        basis_combine_kernel<<<nblk(e->row1 - e->row0, 256), 256, 0, st>>>(e->d_V - e->row0, e->d_coef, m, e->row1, e->v_rows, e->d_beta, thermal, perm, U,
                                                                        accumulate, ydet, e->row0); LAUNCHED(e);
    }
    e->m_lanczos = m;
    e->last_stepnorm = (float)stepnorm;
    if (m_out) *m_out = m;
    return PSE_OK;
}

// ---- public operators -----------------------------------------------------------------------------
static int upload_stepdev(pse_engine* e, uint32_t timestep);
// parts of a velocity evaluation on a slab-decomposed engine (shard.inl)
enum { SV_DET_WAVE = 1, SV_DET_REAL = 2, SV_WNOISE = 4, SV_RNOISE = 8 };
static int shard_velocity(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U, uint32_t timestep, const float* d_u_particles,
                          const float* d_u_grid, unsigned what, int* m_out);

extern "C" int pse_mreal(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U) {
    if (!e || !d_pos || !d_F || !d_U) return PSE_EINVAL;
    if (e->shard) return shard_velocity(e, d_pos, d_F, d_U, 0u, nullptr, nullptr, SV_DET_REAL, nullptr);
    CKRC(ensure_neighbors(e, d_pos));
    gather_vec_kernel<<<nblk(e->N, 256), 256, 0, e->stream>>>(d_F, e->d_perm, e->N, nullptr, (float4*)e->d_px); LAUNCHED(e);
    CKRC(run_spmv_plain(e, e->d_sy));
    scatter_add_kernel<<<nblk(e->N, 256), 256, 0, e->stream>>>(e->d_sy, e->d_perm, e->N, d_U, 0); LAUNCHED(e);
    CK(cudaGetLastError());
    return PSE_OK;
}

extern "C" int pse_mwave(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U) {
    if (!e || !d_pos || !d_F || !d_U) return PSE_EINVAL;
    if (e->shard) return shard_velocity(e, d_pos, d_F, d_U, 0u, nullptr, nullptr, SV_DET_WAVE, nullptr);
    CKRC(ensure_wave_buffers(e));
    CKRC(ensure_neighbors(e, d_pos));
    gather4_kernel<<<nblk(e->N, 256), 256, 0, e->stream>>>(d_F, e->d_perm, e->N, e->d_sx); LAUNCHED(e);
    CKRC(run_wave(e, e->d_sx, d_U, 0, true, false, nullptr));
    CK(cudaGetLastError());
    return PSE_OK;
}

// per-step scalars -> device (ordered on the stream before the kernels that read them)
static int upload_stepdev(pse_engine* e, uint32_t timestep) {
    e->h_stepdev->key = timestep + e->prm.seed_hashed;  // PSEv1/Brownian.cu:117,176
    e->h_stepdev->noise_fac = sqrtf((float)(2.0 * e->cfg.T / e->cfg.dt / e->wp.quadW));  // PSEv1/Brownian.cu:198
    stepdev_store_kernel<<<1, 1, 0, e->stream>>>(e->d_stepdev, *e->h_stepdev); LAUNCHED(e);
    return PSE_OK;
}

// the fixed-topology part of a full velocity evaluation (everything between the neighbour-list decision and the
// first Lanczos host solve): slot gathers, wave-space pipeline, pruning, deterministic SpMV, Lanczos batch
static int velocity_fixed_part(pse_engine* e, const float4* d_F, float4* d_U, bool det, bool wnoise, bool rnoise,
                               const float* d_u_particles, const float* d_u_grid, int m_batch) {
    const uint32_t N = e->N;
    cudaStream_t st = e->stream;
    if (det) { gather_vec_kernel<<<nblk(N, 256), 256, 0, st>>>(d_F, e->d_perm, N, e->d_sx, (float4*)e->d_px); LAUNCHED(e); }
    // The real-space branch (prune, M_real F, Lanczos) and the wave-space branch (bin, spread, FFTs, interpolate) touch
    // disjoint buffers until their results meet in d_U, and they stress different units (L1 gathers vs shared memory /
    // HBM), so they are issued on two streams and joined before the accumulation.  Profiling keeps them serial so
    // that the per-phase times stay meaningful.
    const bool wave = det || wnoise, real = det || rnoise;
    const bool fork = e->overlap && !e->prof_on && wave && real && e->stream2;
    int rc = PSE_OK;
    if (fork) {
        CK(cudaEventRecord(e->ev_fork, st));
        CK(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
        e->stream = e->stream2;
    }
    // full step: M_real F is computed inside the first Lanczos product (dual right-hand side) when that path is available
    const bool dual = det && rnoise && m_batch >= 1 && spmv_dual_available(e);
    if (det && !dual) rc = run_spmv_plain(e, e->d_sy);
    if (rc == PSE_OK && rnoise) rc = lanczos_batch(e, d_u_particles, m_batch, dual);
    if (fork) {
        e->stream = st;
        if (rc == PSE_OK) CK(cudaEventRecord(e->ev_join, e->stream2));
    }
    if (rc != PSE_OK) return rc;
    int acc = 0;
    if (wave) { CKRC(run_wave(e, e->d_sx, d_U, 0, det, wnoise, d_u_grid)); acc = 1; }
    if (fork) CK(cudaStreamWaitEvent(st, e->ev_join, 0));
    if (det && !rnoise) {  // (with the real-space Brownian term the combination kernel adds M_real F in the same pass)
        scatter_add_kernel<<<nblk(N, 256), 256, 0, st>>>(e->d_sy, e->d_perm, N, d_U, acc); LAUNCHED(e);
        acc = 1;
    }
    if (!acc && !rnoise) CK(cudaMemsetAsync(d_U, 0, sizeof(float4) * N, st));
    return PSE_OK;
}

extern "C" int pse_velocity(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U, uint32_t timestep,
                            const float* d_u_particles, const float* d_u_grid, uint32_t parts, int* m_out) {
    if (!e || !d_pos || !d_F || !d_U) return PSE_EINVAL;
    cudaStream_t st = e->stream;
    const bool det = parts & 1u;
    const bool thermal = e->cfg.T > 0.f;  // PSEv1/Brownian.cu:855,885
    const bool wnoise = (parts & 2u) && thermal, rnoise = (parts & 4u) && thermal;
    if (e->shard) {
        if (m_out) *m_out = e->m_lanczos;
        return shard_velocity(e, d_pos, d_F, d_U, timestep, d_u_particles, d_u_grid,
                              (det ? SV_DET_WAVE | SV_DET_REAL : 0u) | (wnoise ? SV_WNOISE : 0u) | (rnoise ? SV_RNOISE : 0u), m_out);
    }
    if (det || wnoise) CKRC(ensure_wave_buffers(e));
    CKRC(ensure_neighbors(e, d_pos));  // host decision (list still valid?) happens before the fixed part
    if (rnoise) CKRC(ensure_krylov(e));
    // everything that needs the positions only runs first: forces still in flight from the host (pse_step_host_async) are
    // waited for after it
    {
        const bool need_prune = (det || rnoise) && e->prune && !e->pruned_valid, need_wbin = (det || wnoise) && e->tiled && !e->wbin_valid;
        const bool fork = need_prune && need_wbin && e->overlap && !e->prof_on && e->stream2;
        if (fork) {   // pruning (L1 gathers) beside the binning / factor rows (HBM streaming)
            CK(cudaEventRecord(e->ev_fork, st));
            CK(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
            e->stream = e->stream2;
            const int rc = ensure_pruned(e);
            e->stream = st;
            if (rc != PSE_OK) return rc;
            CK(cudaEventRecord(e->ev_join, e->stream2));
        } else if (need_prune) CKRC(ensure_pruned(e));
        if (need_wbin) CKRC(run_wbin(e));
        if (fork) CK(cudaStreamWaitEvent(st, e->ev_join, 0));
    }
    if (e->wait_F) { e->wait_F = false; CK(cudaStreamWaitEvent(st, e->ev_F, 0)); }
    if (e->take_Fnext >= 0) {   // forces prefetched earlier: one device copy into the buffer the step (and its graph) reads
        const int r = e->take_Fnext;
        e->take_Fnext = -1;
        CK(cudaStreamWaitEvent(st, e->ev_Fnext[r], 0));
        copy_f4_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->d_hF, e->d_hF_next[r], e->N); LAUNCHED(e);
        CK(cudaEventRecord(e->ev_Fcons[r], st));
    }
    CKRC(upload_stepdev(e, timestep));
    const int m_batch = lanczos_batch_size(e);
    if (m_out) *m_out = e->m_lanczos;

    // Step loop as a captured CUDA graph (the reference issues ~60 launches + 15 blocking copies per step,
    // PSEv1/Stokes.cu:298-355, Brownian.cu:440-739).  The graph is keyed on everything frozen at capture time.
    // The box is baked into the kernel arguments at capture time, so a changing tilt (any shear function) would mean a
    // re-capture and re-instantiation every step: while the tilt is moving the step is issued eagerly instead.
    const bool tilt_steady = e->box.xy == e->xy_prev_call;
    e->xy_prev_call = e->box.xy;
    const bool graphable = e->use_graph && !e->prof_on && det && wnoise && rnoise && !d_u_particles && !d_u_grid && tilt_steady;
    if (graphable) {
        auto& k = e->graph_key;
        const bool hit = k.valid && k.pos == d_pos && k.F == d_F && k.U == d_U && k.m_batch == m_batch && k.xy == e->box.xy &&
                         k.nl_gen == e->nl_gen;
        if (!hit) {
            if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
            k.valid = false;
            cudaGraph_t graph = nullptr;
            const uint64_t l0 = e->launches;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc = velocity_fixed_part(e, d_F, d_U, det, wnoise, rnoise, nullptr, nullptr, m_batch);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != PSE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(e, PSE_ECUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&e->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(e, PSE_ECUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
            e->graph_nodes = e->launches - l0;
            k.pos = d_pos; k.F = d_F; k.U = d_U; k.m_batch = m_batch; k.xy = e->box.xy; k.nl_gen = e->nl_gen; k.valid = true;
        }
        else e->launches += e->graph_nodes;  // kernels replayed by the graph
        CK(cudaGraphLaunch(e->graph_exec, st));
        e->graph_launches++;
    } else {
        CKRC(velocity_fixed_part(e, d_F, d_U, det, wnoise, rnoise, d_u_particles, d_u_grid, m_batch));
    }
    if (rnoise) CKRC(lanczos_finish(e, d_U, (det || wnoise) ? 1 : 0, m_batch, m_out, det ? e->d_sy : nullptr, e->d_perm));
    CK(cudaGetLastError());
    return PSE_OK;
}

extern "C" int pse_mobility(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U) {
    return pse_velocity(e, d_pos, d_F, d_U, 0u, nullptr, nullptr, 1u, nullptr);
}

extern "C" int pse_pair_force(pse_engine* e, const float4* d_pos, const pse_pair_params* prm, float4* d_F, int accumulate) {
    if (!e || !d_pos || !prm || !d_F) return PSE_EINVAL;
    PairParams pp;
    pp.kind = prm->kind; pp.eps = prm->epsilon; pp.sigma = prm->sigma; pp.rcut = prm->r_cut; pp.shift = 0.f;
    if (prm->kind == PSE_PAIR_WCA) { pp.rcut = 1.122462048309373f * prm->sigma; pp.shift = prm->epsilon; }
    else if (prm->kind != PSE_PAIR_LJ && prm->kind != PSE_PAIR_HARMONIC) return fail(e, PSE_EINVAL, "pse_pair_force: unknown kind %d", prm->kind);
    if (!(pp.rcut > 0.f) || pp.rcut > e->prm.rcut)
        return fail(e, PSE_EINVAL, "pse_pair_force: r_cut %g outside (0, %g] (the neighbour list is complete up to the real-space cutoff)",
                    pp.rcut, e->prm.rcut);
    pp.rcut_sq = pp.rcut * pp.rcut;
    if (e->shard && e->shard->world > 1) return fail(e, PSE_EINVAL, "pse_pair_force: a slab-decomposed engine holds the rows of its own slab only");
    CKRC(ensure_neighbors(e, d_pos));
    CKRC(ensure_csr(e));
    pair_force_kernel<<<nblk((size_t)e->N * 8, 256), 256, 0, e->stream>>>(e->d_spos, e->N, e->d_nn, e->d_head, e->d_nl, e->d_perm, pp, e->box,
                                                                      d_F, accumulate); LAUNCHED(e);
    CK(cudaGetLastError());
    return PSE_OK;
}

// Euler update + affine shear advection + periodic wrap: PSEv1/Stokes.cu:137-192
__global__ void integrate_kernel(float4* __restrict__ pos, int3* __restrict__ image, const float4* __restrict__ vel, uint32_t N,
                                 PseBox box, float dt, float shear_rate) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 p = pos[i];
    const float4 v = __ldg(vel + i);
    float3 w = make_float3(p.x, p.y, p.z);
    const float vx = v.x + shear_rate * p.y;
    w.x += vx * dt; w.y += v.y * dt; w.z += v.z * dt;
    int3 im = image ? image[i] : make_int3(0, 0, 0);
    box.wrap(w, im);
    pos[i] = make_float4(w.x, w.y, w.z, p.w);
    if (image) image[i] = im;
}

extern "C" int pse_step(pse_engine* e, float4* d_pos, int3* d_image, const float4* d_F, float4* d_vel, uint32_t timestep,
                        float shear_rate, int* m_out) {
    if (!e || !d_pos || !d_F) return PSE_EINVAL;
    float4* vel = d_vel ? d_vel : e->d_vel_work;
    CKRC(pse_velocity(e, d_pos, d_F, vel, timestep, nullptr, nullptr, 7u, m_out));
    if (e->wait_img) { e->wait_img = false; CK(cudaStreamWaitEvent(e->stream, e->ev_img, 0)); }
    if (e->wait_out) { e->wait_out = false; CK(cudaStreamWaitEvent(e->stream, e->ev_out, 0)); }
    {
        ProfScope ps(e, PH_INTEGRATE);
        integrate_kernel<<<nblk(e->N, 256), 256, 0, e->stream>>>(d_pos, d_image, vel, e->N, e->box, e->cfg.dt, shear_rate); LAUNCHED(e);
    }
    CK(cudaGetLastError());
    return PSE_OK;
}

// Host-buffer entry points.  The state of a host-driven run (positions, images) lives on the device between calls; what
// crosses PCIe every step is the step's input (forces, 16 N bytes up) and its result (positions + images, 28 N bytes down),
// and both transfers are hidden: the forces arrive on a copy stream while the position-only head of the step runs
// (neighbour-list check / rebuild, pruning, wave-space binning and Gaussian factors), the result leaves on a second copy
// stream while the NEXT step computes.  pse_wait blocks until the last result is on the host.
static int host_path_setup(pse_engine* e) {
    const size_t N = e->N;
    if (!e->d_hpos) {
        CK(cudaMalloc(&e->d_hpos, sizeof(float4) * N));
        CK(cudaMalloc(&e->d_hF, sizeof(float4) * N));
        CK(cudaMalloc(&e->d_himage, sizeof(int3) * N));
    }
    if (!e->stream_h2d) {
        CK(cudaStreamCreateWithFlags(&e->stream_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&e->stream_d2h, cudaStreamNonBlocking));
        for (cudaEvent_t* ev : {&e->ev_F, &e->ev_img, &e->ev_step, &e->ev_out, &e->ev_Fnext[0], &e->ev_Fnext[1], &e->ev_Fcons[0], &e->ev_Fcons[1]})
            CK(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        CK(cudaEventRecord(e->ev_step, e->stream));
        CK(cudaEventRecord(e->ev_Fcons[0], e->stream));
        CK(cudaEventRecord(e->ev_Fcons[1], e->stream));
    }
    return PSE_OK;
}
// Start uploading the forces of a LATER pse_step_host_async call now; calls given h_F4 == NULL consume the prefetched sets in
// order.  Two may be outstanding, so the upload for step t + 1 can be issued before the call for step t.
extern "C" int pse_host_prefetch_forces(pse_engine* e, const float* h_F4) {
    if (!e || !h_F4) return PSE_EINVAL;
    if (e->n_pref - e->n_cons >= 2) return fail(e, PSE_EINVAL, "pse_host_prefetch_forces: two sets of forces are already waiting");
    CKRC(host_path_setup(e));
    const int w = (int)(e->n_pref & 1);
    if (!e->d_hF_next[w]) CK(cudaMalloc(&e->d_hF_next[w], sizeof(float4) * e->N));
    CK(cudaStreamWaitEvent(e->stream_h2d, e->ev_Fcons[w], 0));   // the previous consumer of this landing buffer is done with it
    CK(cudaMemcpyAsync(e->d_hF_next[w], h_F4, sizeof(float4) * e->N, cudaMemcpyHostToDevice, e->stream_h2d));
    CK(cudaEventRecord(e->ev_Fnext[w], e->stream_h2d));
    e->n_pref++;
    return PSE_OK;
}
extern "C" int pse_step_host_async(pse_engine* e, float* h_pos4, int* h_image3, const float* h_F4, float* h_vel4, uint32_t timestep,
                                   float shear_rate, uint32_t flags, int* m_out) {
    if (!e || !h_pos4) return PSE_EINVAL;
    if (!h_F4 && e->n_pref == e->n_cons) return fail(e, PSE_EINVAL, "pse_step_host_async: no forces (h_F4 is NULL and nothing was prefetched)");
    const size_t N = e->N;
    cudaStream_t st = e->stream;
    CKRC(host_path_setup(e));
    // (the velocity scratch is rewritten early in a step: a pending download of it has to finish first)
    if (e->wait_out && e->out_has_vel) { e->wait_out = false; CK(cudaStreamWaitEvent(st, e->ev_out, 0)); }
    if (!h_F4) {
        // prefetched during the previous step (pse_host_prefetch_forces): one device copy into the buffer the step (and its
        // captured graph) reads
        e->take_Fnext = (int)(e->n_cons++ & 1);   // (taken where the forces are first needed, after the position-only head: pse_velocity)
    } else {
        // the forces of this step: staged behind the previous step's last kernels (which may still be reading the staging buffer)
        CK(cudaStreamWaitEvent(e->stream_h2d, e->ev_step, 0));
        CK(cudaMemcpyAsync(e->d_hF, h_F4, sizeof(float4) * N, cudaMemcpyHostToDevice, e->stream_h2d));
        CK(cudaEventRecord(e->ev_F, e->stream_h2d));
        e->wait_F = true;
    }
    if ((flags & PSE_HOST_STATE_IN) || !e->host_state_valid) {
        // the caller's positions / images replace the device state (first call, or the host changed them)
        if (e->wait_out) { e->wait_out = false; CK(cudaStreamWaitEvent(st, e->ev_out, 0)); }
        CK(cudaMemcpyAsync(e->d_hpos, h_pos4, sizeof(float4) * N, cudaMemcpyHostToDevice, st));
        if (h_image3) CK(cudaMemcpyAsync(e->d_himage, h_image3, sizeof(int3) * N, cudaMemcpyHostToDevice, e->stream_h2d));
        else CK(cudaMemsetAsync(e->d_himage, 0, sizeof(int3) * N, e->stream_h2d));
        CK(cudaEventRecord(e->ev_img, e->stream_h2d));
        e->wait_img = true;
        e->host_state_valid = true;
    }
    if (flags & PSE_HOST_STATE_IN) e->precheck_pos = nullptr;   // (a check issued for the old state says nothing about the new one)
    const int rc = pse_step(e, e->d_hpos, e->d_himage, e->d_hF, h_vel4 ? e->d_vel_work : nullptr, timestep, shear_rate, m_out);
    if (rc != PSE_OK) { e->host_state_valid = false; return rc; }
    // the next step's displacement / moved check, now: its flags reach the host ahead of the state download below
    static const bool precheck_on = !(getenv("PSE_PRECHECK") && getenv("PSE_PRECHECK")[0] == '0');
    if (e->nlist_valid && precheck_on) {
        CKRC(launch_disp_check(e, e->d_hpos));
        e->precheck_pos = e->d_hpos; e->precheck_xy = e->box.xy;
        // (the check moved the slot-ordered positions on to the new state: what was derived from the old one is stale even
        // if a later call finds "nothing moved" relative to them)
        e->pruned_valid = false; e->wbin_valid = false;
    }
    CK(cudaEventRecord(e->ev_step, st));
    CK(cudaStreamWaitEvent(e->stream_d2h, e->ev_step, 0));
    if (!(flags & PSE_HOST_NO_STATE_OUT)) {
        CK(cudaMemcpyAsync(h_pos4, e->d_hpos, sizeof(float4) * N, cudaMemcpyDeviceToHost, e->stream_d2h));
        if (h_image3) CK(cudaMemcpyAsync(h_image3, e->d_himage, sizeof(int3) * N, cudaMemcpyDeviceToHost, e->stream_d2h));
    }
    if (h_vel4) CK(cudaMemcpyAsync(h_vel4, e->d_vel_work, sizeof(float4) * N, cudaMemcpyDeviceToHost, e->stream_d2h));
    CK(cudaEventRecord(e->ev_out, e->stream_d2h));
    e->wait_out = true;   // the next integrate waits for this download before it overwrites the state
    e->out_has_vel = h_vel4 != nullptr;
    return PSE_OK;
}
extern "C" int pse_wait(pse_engine* e) {
    if (!e) return PSE_EINVAL;
    if (e->ev_out) CK(cudaEventSynchronize(e->ev_out));
    return PSE_OK;
}
// synchronous form: state in, one step, state out, wait
extern "C" int pse_step_host(pse_engine* e, float* h_pos4, int* h_image3, const float* h_F4, float* h_vel4, uint32_t timestep,
                             float shear_rate, int* m_out) {
    CKRC(pse_step_host_async(e, h_pos4, h_image3, h_F4, h_vel4, timestep, shear_rate, PSE_HOST_STATE_IN, m_out));
    return pse_wait(e);
}

// test hook: host tridiagonal square root (compared against numpy in the CPU tests)
extern "C" int pse_test_tridiag_sqrt_e1(int m, const double* diag, const double* off, double* c) {
    return pse_tridiag_sqrt_e1(m, diag, off, c, nullptr);
}

// test hook: Philox4x32-10 known-answer vectors (host path of rng.cuh)
extern "C" void pse_test_philox4x32(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    uint4 r = pse_philox4x32(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// test hook: the constant-bank polynomial form of f(r), g(r) (coefficients + float-evaluated max error vs the closed forms)
extern "C" int pse_test_fit_rpy_cheb(double xi, double rcut, float* out /* 3 + 2 (PSE_CHEB_DEG + 1) */, double* max_err) {
    return pse_fit_rpy_cheb(xi, rcut, out, max_err);
}

// test hook: radix schedule and digit-reversal table of the in-place FFT of length N (host logic of fft.cuh / setup_own_fft)
extern "C" int pse_test_fft_plan(int N, int* radix_out /* FFT_MAX_PASSES */, int* npass_out, uint16_t* pos_of /* N */) {
    Fft1D f;
    if (!fft_factor(N, &f)) return PSE_EINVAL;
    *npass_out = f.npass;
    for (int i = 0; i < f.npass; ++i) radix_out[i] = (int)((f.radices >> (4 * i)) & 15ull);
    for (int k = 0; k < N; ++k) {
        int kk = k, span = N, p = 0;
        for (int i = 0; i < f.npass; ++i) {
            const int r = radix_out[i];
            span /= r; p += (kk % r) * span; kk /= r;
        }
        pos_of[k] = (uint16_t)p;
    }
    return PSE_OK;
}

#include "shard.inl"

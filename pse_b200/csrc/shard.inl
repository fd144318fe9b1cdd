// Slab decomposition of the whole step over the GPUs of one node (included at the end of engine.cu; the design is described
// next to ShardState there).  New work: the reference is single-GPU ("only one GPU is supported", PSEv1/Stokes.cc:104).

// ---- pack / unpack kernels --------------------------------------------------------------------------------------------
// slab layout  s[c][xl][y][kz]  <->  per-destination blocks  [q][c][xl][y - ys[q]][kz]
__global__ void shard_slab_blocks_kernel(float2* __restrict__ slab, float2* __restrict__ blocks, ShardBounds b, int nxl, int Ny, int Nzp,
                                         int to_blocks) {
    const size_t n = (size_t)3 * nxl * Ny * Nzp;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int kz = (int)(t % Nzp);
        const int y = (int)((t / Nzp) % Ny);
        const int xl = (int)((t / ((size_t)Nzp * Ny)) % nxl);
        const int c = (int)(t / ((size_t)Nzp * Ny * nxl));
        int q = 0;
        while (y >= b.ys[q + 1]) ++q;
        const int nyl = b.ys[q + 1] - b.ys[q];
        const size_t off = (size_t)3 * nxl * Nzp * b.ys[q] + (((size_t)c * nxl + xl) * nyl + (y - b.ys[q])) * Nzp + kz;
        if (to_blocks) blocks[off] = slab[t]; else slab[t] = blocks[off];
    }
}
// transposed layout  t[c][x][yl][kz]  <->  per-source blocks  [r][c][x - xs[r]][yl][kz]
__global__ void shard_trans_blocks_kernel(float2* __restrict__ tr, float2* __restrict__ blocks, ShardBounds b, int Nx, int nyl, int Nzp,
                                          int to_blocks) {
    const size_t n = (size_t)3 * Nx * nyl * Nzp;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int kz = (int)(t % Nzp);
        const int yl = (int)((t / Nzp) % nyl);
        const int x = (int)((t / ((size_t)Nzp * nyl)) % Nx);
        const int c = (int)(t / ((size_t)Nzp * nyl * Nx));
        int r = 0;
        while (x >= b.xs[r + 1]) ++r;
        const int nxr = b.xs[r + 1] - b.xs[r];
        const size_t off = (size_t)3 * nyl * Nzp * b.xs[r] + (((size_t)c * nxr + (x - b.xs[r])) * nyl + yl) * Nzp + kz;
        if (to_blocks) blocks[off] = tr[t]; else tr[t] = blocks[off];
    }
}
// planes [p0, p0 + n) of the three components of the local real-space buffer  <->  contiguous buffer [c][i][y][z]
// mode 0: buffer <- planes, 1: planes <- buffer, 2: planes += buffer
__global__ void shard_planes_kernel(float* __restrict__ grid, float* __restrict__ buf, size_t Gl, size_t plane, int p0, int n, int mode) {
    const size_t tot = (size_t)3 * n * plane;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const size_t in_plane = t % plane;
        const int i = (int)((t / plane) % n);
        const int c = (int)(t / (plane * n));
        const size_t gi = (size_t)c * Gl + (size_t)(p0 + i) * plane + in_plane;
        if (mode == 0) buf[t] = grid[gi];
        else if (mode == 1) grid[gi] = buf[t];
        else grid[gi] += buf[t];
    }
}
// guard: every own particle's support must lie inside the local buffer (halo widths are sized for the displacement and tilt
// drift the neighbour-list buffer allows; this catches a violated assumption instead of silently dropping weight)
__global__ void shard_cover_kernel(const int4* __restrict__ org, uint32_t row0, uint32_t row1, int Nx, int xorg, int nxa, int P,
                                   uint32_t* __restrict__ err) {
    const uint32_t s = row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= row1) return;
    int rel = org[s].x - xorg;
    if (rel < 0) rel += Nx;
    if (rel + P > nxa) atomicOr(err, 1u);
}
// U[perm[slot]].xyz = uslot[slot].xyz   (the caller's .w column is kept)
__global__ void shard_scatter_kernel(const float4* __restrict__ uslot, const uint32_t* __restrict__ perm, uint32_t N, float4* __restrict__ U) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const float4 v = __ldg(uslot + s);
    const uint32_t p = perm[s];
    float4 o = U[p];
    o.x = v.x; o.y = v.y; o.z = v.z;
    U[p] = o;
}

static void shard_free(ShardState* s) {
    if (!s) return;
    s->comm.destroy();
    for (int q = 0; q < SHARD_MAX_WORLD; ++q)
        for (int k = 0; k < PSE_PEER_NBUF; ++k)
            if (s->ipc_opened[q][k]) cudaIpcCloseMemHandle(s->ipc_opened[q][k]);
    if (s->d_pad) cudaFree(s->d_pad);
    void* bufs[] = {s->d_sloc, s->d_tr, s->d_a2a_a, s->d_a2a_b, s->d_hsL, s->d_hsR, s->d_hrL, s->d_hrR, s->d_uslot, s->d_layer_start,
                    s->d_vsL, s->d_vsR, s->d_vrL, s->d_vrR};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (s->h_layer_start) cudaFreeHost(s->h_layer_start);
    delete s;
}

// ---- static geometry: identical on every rank, derived from the configuration only ------------------------------------------
// x layers of cells are dealt out evenly; a rank transforms the x planes its layers cover; halo widths cover the Gaussian
// support ((P - 1) / 2 nodes to the left and (P + 1) / 2 to the right of a particle's node, PSEv1/Mobility.cu:173-214; one
// more on the right because the last layer's particles can sit on the first plane of the next slab) plus the motion the
// Verlet buffer allows between two list rebuilds (r_buff / 2 in the sheared frame), the tilt drift SHARD_DRIFT_NODES that
// forces a rebuild (stale_from_bits) and one node for float rounding; shard_cover_kernel guards the assumption.
static int shard_static_geometry(const pse_config& cfg, const pse_params& prm, int world, int tile_x, ShardGeom* g, char* err, size_t errlen) {
    memset(g, 0, sizeof(*g));
    g->world = world;
    if (world < 1 || world > SHARD_MAX_WORLD) { snprintf(err, errlen, "world size %d outside [1, %d]", world, SHARD_MAX_WORLD); return PSE_EINVAL; }
    const float r_buff = cfg.r_buff >= 0.f ? cfg.r_buff : 0.4f;
    const float rlist = prm.rcut + r_buff;
    int ncx, ncy, ncz;
    cell_grid_dims(cfg.box.Lx, cfg.box.Ly, cfg.box.Lz, rlist, cell_capacity(cfg.N), &ncx, &ncy, &ncz);
    for (int r = 0; r <= world; ++r) {
        g->LB[r] = (int)(((long long)ncx * r) / world);
        g->X[r] = (int)(((long long)g->LB[r] * prm.Nx) / ncx);
        g->YS[r] = (int)(((long long)prm.Ny * r) / world);
    }
    g->X[world] = prm.Nx;
    if (world == 1) { g->HL = g->HR = 0; return PSE_OK; }
    const float ms = fabsf(cfg.max_strain);
    const int dn = (int)ceilf(0.5f * r_buff * sqrtf(1.f + ms * ms) / prm.hx) + (int)SHARD_DRIFT_NODES + 1;
    g->HL = (prm.P - 1) / 2 + dn;
    g->HR = (prm.P + 1) / 2 + 1 + dn;
    // the search reach in layers at the largest tilt (shard_update_geometry recomputes it for the current one)
    const float reach = rlist * sqrtf(1.f + ms * ms) / cfg.box.Lx * 1.0001f + 1e-6f;
    const int kh = (int)ceilf(reach * ncx) + 1;
    for (int r = 0; r < world; ++r) {
        const int nown = g->X[r + 1] - g->X[r], nlay = g->LB[r + 1] - g->LB[r];
        if (nown < std::max(g->HL, g->HR) || g->HL + tile_x - 1 + nown + g->HR > prm.Nx) {
            snprintf(err, errlen, "rank %d of %d would own %d x planes of %d: too thin for halos of %d + %d planes", r, world, nown, prm.Nx, g->HL, g->HR);
            return PSE_EINVAL;
        }
        if (nlay < kh) {
            snprintf(err, errlen, "rank %d of %d would own %d x layers of cells, the neighbour search reaches %d", r, world, nlay, kh);
            return PSE_EINVAL;
        }
    }
    return PSE_OK;
}
// local real-space buffer of rank q: first global plane, own planes' offset, planes held
static void shard_rank_layout(const ShardGeom& g, int q, int Nx, int tile_x, int* xorg, int* BL, int* nxa) {
    if (g.world == 1) { *xorg = 0; *BL = 0; *nxa = Nx; return; }
    int lo = g.X[q] - g.HL; if (lo < 0) lo += Nx;
    *xorg = (lo / tile_x) * tile_x;      // tile aligned: no tile an own particle is binned to starts before the buffer
    int bl = g.X[q] - *xorg; if (bl < 0) bl += Nx;
    *BL = bl;
    *nxa = bl + (g.X[q + 1] - g.X[q]) + g.HR;
}
static void shard_fill_info(const ShardGeom& g, int rank, const pse_params& prm, int Nzp, int tile_x, pse_shard_info* out) {
    memset(out, 0, sizeof(*out));
    out->rank = rank; out->world = g.world;
    out->x0 = g.X[rank]; out->x1 = g.X[rank + 1]; out->y0 = g.YS[rank]; out->y1 = g.YS[rank + 1];
    out->halo_left = g.HL; out->halo_right = g.HR;
    out->layer0 = g.LB[rank]; out->layer1 = g.LB[rank + 1];
    const int nown = out->x1 - out->x0, nyl = out->y1 - out->y0;
    int xorg, bl;
    shard_rank_layout(g, rank, prm.Nx, tile_x, &xorg, &bl, &out->buffer_planes);
    for (int q = 0; q < g.world; ++q) {
        out->a2a_send_bytes[q] = (uint64_t)8 * 3 * nown * (g.YS[q + 1] - g.YS[q]) * Nzp;
        out->a2a_recv_bytes[q] = (uint64_t)8 * 3 * (g.X[q + 1] - g.X[q]) * nyl * Nzp;
    }
}
static int spec_row_pad(int Nzh) {
    const char* env = getenv("PSE_SPEC_PAD");
    const int pad = env ? atoi(env) : 8;
    return pad > 1 ? ((Nzh + pad - 1) / pad) * pad : Nzh;
}

extern "C" int pse_shard_plan(const pse_config* cfg, int rank, int world, pse_shard_info* out) {
    if (!cfg || !out || rank < 0 || rank >= world) return PSE_EINVAL;
    pse_params p;
    int rc = pse_derive_params(cfg, &p);
    if (rc != PSE_OK) return rc;
    int tx, ty, tz;
    if (!v2_shape(p.P, &tx, &ty, &tz)) return PSE_EINVAL;
    ShardGeom g;
    char err[256];
    rc = shard_static_geometry(*cfg, p, world, tx, &g, err, sizeof(err));
    if (rc != PSE_OK) { snprintf(g_create_error, sizeof(g_create_error), "pse_shard_plan: %s", err); return rc; }
    shard_fill_info(g, rank, p, spec_row_pad(p.Nz / 2 + 1), tx, out);
    return PSE_OK;
}

extern "C" int pse_comm_unique_id(uint8_t out[128]) {
    char err[256];
    NcclApi* api = nccl_api(err, sizeof(err));
    if (!api) { snprintf(g_create_error, sizeof(g_create_error), "pse_comm_unique_id: %s", err); return PSE_ECUDA; }
    pse_nccl_uid id;
    memset(&id, 0, sizeof(id));
    const int rc = api->GetUniqueId(&id);
    if (rc != 0) { snprintf(g_create_error, sizeof(g_create_error), "ncclGetUniqueId: %s", api->GetErrorString(rc)); return PSE_ECUDA; }
    memcpy(out, id.internal, 128);
    return PSE_OK;
}
extern "C" pse_local_world* pse_local_world_create(int world) {
    if (world < 1 || world > PSE_COMM_MAX_WORLD) return nullptr;
    pse_local_world* w = new pse_local_world();
    memset(w, 0, sizeof(*w));
    w->world = world;
    pthread_barrier_init(&w->bar, nullptr, (unsigned)world);
    return w;
}
extern "C" void pse_local_world_destroy(pse_local_world* w) {
    if (!w) return;
    pthread_barrier_destroy(&w->bar);
    delete w;
}

extern "C" int pse_shard_init(pse_engine* e, int rank, int world, const uint8_t* nccl_uid128, pse_local_world* local) {
    if (!e) return PSE_EINVAL;
    if (rank < 0 || rank >= world) return fail(e, PSE_EINVAL, "pse_shard_init: rank %d of %d", rank, world);
    if (e->shard) return fail(e, PSE_EINVAL, "pse_shard_init: already slab-decomposed");
    if (e->nlist_valid) return fail(e, PSE_EINVAL, "pse_shard_init: must be called before the first operator call");
    if (!e->wave_v2 || !e->own_fft)
        return fail(e, PSE_EINVAL, "pse_shard_init: needs the spread2 / interp2 kernels (P = 6, 7, 8) and the engine's own FFT passes (grid sizes 2^a 3^b 5^c)");
    ShardGeom g;
    char gerr[256];
    int rc = shard_static_geometry(e->cfg, e->prm, world, e->tg.tx, &g, gerr, sizeof(gerr));   // every rank fails or nobody does
    if (rc != PSE_OK) return fail(e, rc, "pse_shard_init: %s", gerr);
    ShardState* s = new ShardState();
    memset(s, 0, sizeof(*s));
    s->rank = rank; s->world = world; s->g = g;
    if (s->comm.init(rank, world, nccl_uid128, local) != 0) {
        fail(e, PSE_ECUDA, "pse_shard_init: %s", s->comm.err);
        delete s;
        return PSE_ECUDA;
    }
    e->shard = s;
    WaveParams& wp = e->wp;
    s->plane = (size_t)wp.Ny * wp.Nz;
    s->nown = g.X[rank + 1] - g.X[rank];
    for (int q = 0; q < world; ++q) { int xo; shard_rank_layout(g, q, wp.Nx, e->tg.tx, &xo, &s->BLq[q], &s->nxaq[q]); }
    shard_rank_layout(g, rank, wp.Nx, e->tg.tx, &s->xorg, &s->BL, &s->nxa);
    wp.nxw = world == 1 ? wp.Nx : 1 << 30;         // a slab buffer is indexed without periodic wrap
    { const char* env = getenv("PSE_COMM"); s->use_peer = world > 1 && !(env && env[0] == 'c'); }
    { const char* env = getenv("PSE_TRANSPOSE"); s->push = env && env[0] == 'p' && env[1] == 'u' && env[2] == 's'; }   // "push": remote stores (measured equal)
    wp.xorg = s->xorg; wp.nxa = s->nxa;
    s->Gl = (size_t)s->nxa * s->plane;
    // the single-GPU grids / basis are allocated lazily, so there is normally nothing to free here
    if (e->d_grid) { cudaFree(e->d_grid); e->d_grid = nullptr; }
    if (e->d_spec) { cudaFree(e->d_spec); e->d_spec = nullptr; }
    if (e->d_V) { cudaFree(e->d_V); e->d_V = nullptr; e->v_rows = 0; }
    const int nyl = g.YS[rank + 1] - g.YS[rank];
    const size_t nsloc = (size_t)3 * s->nown * wp.Ny * wp.Nzp, ntr = (size_t)3 * wp.Nx * std::max(nyl, 1) * wp.Nzp;
    CK(cudaMalloc(&e->d_grid, sizeof(float) * 3 * s->Gl));
    e->grid_planes = s->nxa;
    CK(cudaMalloc(&s->d_sloc, sizeof(float2) * nsloc));
    CK(cudaMalloc(&s->d_tr, sizeof(float2) * ntr));
    CK(cudaMalloc(&s->d_a2a_a, sizeof(float2) * nsloc));
    CK(cudaMalloc(&s->d_a2a_b, sizeof(float2) * ntr));
    s->a2a_send_off[0] = s->a2a_recv_off[0] = 0;
    for (int q = 0; q < world; ++q) {
        s->a2a_send_off[q + 1] = s->a2a_send_off[q] + (size_t)8 * 3 * s->nown * (g.YS[q + 1] - g.YS[q]) * wp.Nzp;
        s->a2a_recv_off[q + 1] = s->a2a_recv_off[q] + (size_t)8 * 3 * (g.X[q + 1] - g.X[q]) * nyl * wp.Nzp;
    }
    if (world > 1) {
        CK(cudaMalloc(&s->d_hsL, sizeof(float) * 3 * std::max(g.HL, g.HR) * s->plane));
        CK(cudaMalloc(&s->d_hsR, sizeof(float) * 3 * std::max(g.HL, g.HR) * s->plane));
        CK(cudaMalloc(&s->d_hrL, sizeof(float) * 3 * std::max(g.HL, g.HR) * s->plane));
        CK(cudaMalloc(&s->d_hrR, sizeof(float) * 3 * std::max(g.HL, g.HR) * s->plane));
    }
    CK(cudaMalloc(&s->d_uslot, sizeof(float4) * e->N));
    CK(cudaMemset(s->d_uslot, 0, sizeof(float4) * e->N));
    CK(cudaMalloc(&s->d_pad, PEER_PAD_BYTES));
    CK(cudaMemset(s->d_pad, 0, PEER_PAD_BYTES));
    CK(cudaMalloc(&s->d_layer_start, sizeof(uint32_t) * (e->cg.ncx + 2)));
    CK(cudaMallocHost(&s->h_layer_start, sizeof(uint32_t) * (e->cg.ncx + 2)));
    e->use_graph = false;   // the collectives make the step topology a host matter: issued eagerly
    return PSE_OK;
}

extern "C" int pse_shard_get_info(pse_engine* e, pse_shard_info* out) {
    if (!e || !out) return PSE_EINVAL;
    if (!e->shard) return fail(e, PSE_EINVAL, "pse_shard_get_info: not slab-decomposed");
    const ShardState* s = e->shard;
    shard_fill_info(s->g, s->rank, e->prm, e->wp.Nzp, e->tg.tx, out);
    out->buffer_planes = s->nxa;
    out->halo_layers = s->g.KH;
    out->row0 = e->row0; out->row1 = e->row1;
    out->bytes_sent = s->bytes_sent; out->collectives = s->collectives;
    return PSE_OK;
}

// ---- per-rebuild geometry: slot ranges of the layers (after the replicated binning of pse_build_neighbors) -------------------
static int shard_update_geometry(pse_engine* e) {
    ShardState* s = e->shard;
    ShardGeom& g = s->g;
    cudaStream_t st = e->stream;
    const CellGrid& cg = e->cg;
    layer_start_kernel<<<nblk(cg.ncx + 1, 128), 128, 0, st>>>(e->d_cell_start, cg.ncx, cg.ncy * cg.ncz, s->d_layer_start); LAUNCHED(e);
    words_copy_kernel<<<1, 128, 0, st>>>(s->h_layer_start, s->d_layer_start, cg.ncx + 1); LAUNCHED(e);
    CK(cudaStreamSynchronize(st));
    const uint32_t* ls = s->h_layer_start;
    g.KH = g.world > 1 ? (int)ceilf(cg.reach_fx * cg.ncx) + 1 : 0;
    for (int r = 0; r <= g.world; ++r) g.ROW[r] = ls[g.LB[r]];
    size_t max_halo = 1;
    for (int r = 0; r < g.world; ++r) {
        if (g.world > 1 && g.LB[r + 1] - g.LB[r] < g.KH)
            return fail(e, PSE_EINVAL, "slab of rank %d has %d x layers of cells, the neighbour search reaches %d", r, g.LB[r + 1] - g.LB[r], g.KH);
        g.SL1[r] = ls[std::min(g.LB[r] + g.KH, g.LB[r + 1])];
        g.SR0[r] = ls[std::max(g.LB[r + 1] - g.KH, g.LB[r])];
        max_halo = std::max<size_t>(max_halo, std::max(g.SL1[r] - g.ROW[r], g.ROW[r + 1] - g.SR0[r]));
    }
    e->row0 = g.ROW[s->rank]; e->row1 = g.ROW[s->rank + 1];
    if (g.world > 1 && max_halo > s->vcap) {
        for (float4** b : {&s->d_vsL, &s->d_vsR, &s->d_vrL, &s->d_vrR}) { if (*b) cudaFree(*b); *b = nullptr; }
        s->vcap = max_halo + max_halo / 4 + 1024;
        for (float4** b : {&s->d_vsL, &s->d_vsR, &s->d_vrL, &s->d_vrR}) CK(cudaMalloc(b, sizeof(float4) * s->vcap));
    }
    return PSE_OK;
}

// ---- the exchanges -----------------------------------------------------------------------------------------------------------
#define CKCOMM(call) do { if ((call) != 0) return fail(e, PSE_ECUDA, "%s", s->comm.err); s->collectives++; } while (0)

// Peer-memory transport: exchange the addresses of the buffers the peers read (once, at the first evaluation: a collective
// step).  Separate processes: CUDA IPC handles, all-gathered through the NCCL communicator; one process: plain pointers.
struct IpcPack { cudaIpcMemHandle_t h[PSE_PEER_NBUF]; };
static int shard_connect_peers(pse_engine* e) {
    ShardState* s = e->shard;
    if (!s->use_peer || s->peer_ready) return PSE_OK;
    const int W = s->world, me = s->rank;
    cudaStream_t st = e->stream;
    void* mine[PSE_PEER_NBUF] = {s->d_pad, e->d_px, s->d_uslot, e->d_grid, s->d_sloc, s->d_tr};
    void* theirs[SHARD_MAX_WORLD][PSE_PEER_NBUF];
    if (s->comm.lw) {
        pse_local_world* lw = s->comm.lw;
        for (int k = 0; k < PSE_PEER_NBUF; ++k) lw->ptrs[me][k] = mine[k];
        CK(cudaStreamSynchronize(st));
        pthread_barrier_wait(&lw->bar);
        for (int q = 0; q < W; ++q) for (int k = 0; k < PSE_PEER_NBUF; ++k) theirs[q][k] = lw->ptrs[q][k];
        pthread_barrier_wait(&lw->bar);
    } else {
        IpcPack* d_packs = nullptr;
        std::vector<IpcPack> packs(W);
        bool ok = true;
        for (int k = 0; k < PSE_PEER_NBUF; ++k) ok &= cudaIpcGetMemHandle(&packs[me].h[k], mine[k]) == cudaSuccess;
        if (!ok) { cudaGetLastError(); memset(&packs[me], 0, sizeof(IpcPack)); }
        CK(cudaMalloc(&d_packs, sizeof(IpcPack) * W));
        CK(cudaMemcpyAsync(d_packs + me, &packs[me], sizeof(IpcPack), cudaMemcpyHostToDevice, st));
        size_t off[SHARD_MAX_WORLD + 1];
        for (int q = 0; q <= W; ++q) off[q] = sizeof(IpcPack) * q;
        if (s->comm.allgatherv(d_packs, off, st) != 0) { cudaFree(d_packs); return fail(e, PSE_ECUDA, "%s", s->comm.err); }
        CK(cudaMemcpyAsync(packs.data(), d_packs, sizeof(IpcPack) * W, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int q = 0; q < W && ok; ++q) {
            if (q == me) { for (int k = 0; k < PSE_PEER_NBUF; ++k) theirs[q][k] = mine[k]; continue; }
            IpcPack zero; memset(&zero, 0, sizeof(zero));
            if (!memcmp(&packs[q], &zero, sizeof(zero))) { ok = false; break; }
            for (int k = 0; k < PSE_PEER_NBUF && ok; ++k) {
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, packs[q].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
                s->ipc_opened[q][k] = ptr;
                theirs[q][k] = ptr;
            }
        }
        // everybody or nobody: a rank that could not map a peer takes all ranks back to the NCCL collectives
        double* d_ok = reinterpret_cast<double*>(d_packs);
        const double mine_ok = ok ? 0.0 : 1.0;
        CK(cudaMemcpyAsync(d_ok, &mine_ok, sizeof(double), cudaMemcpyHostToDevice, st));
        if (s->comm.allreduce_sum(d_ok, 1, st) != 0) { cudaFree(d_packs); return fail(e, PSE_ECUDA, "%s", s->comm.err); }
        double failed = 0.0;
        CK(cudaMemcpyAsync(&failed, d_ok, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        cudaFree(d_packs);
        if (failed != 0.0) {
            fprintf(stderr, "pse_b200: CUDA IPC mapping of peer buffers failed on %d rank(s); using NCCL collectives\n", (int)failed);
            s->use_peer = false;
            return PSE_OK;
        }
    }
    s->psync.rank = me; s->psync.world = W; s->psync.err = e->d_flag + 1;
    for (int q = 0; q < W; ++q) {
        s->psync.pad[q] = static_cast<unsigned char*>(theirs[q][0]);
        s->peer_px[q] = static_cast<const PX*>(theirs[q][1]);
        s->peer_uslot[q] = static_cast<const float4*>(theirs[q][2]);
        s->peer_grid[q] = static_cast<const float*>(theirs[q][3]);
        s->peer_sloc[q] = static_cast<const float2*>(theirs[q][4]);
        s->peer_tr[q] = static_cast<const float2*>(theirs[q][5]);
    }
    s->peer_ready = true;
    return PSE_OK;
}
// Every rank has finished everything issued before this point.  Between processes: flags in peer memory, device side, no
// host involvement.  Virtual ranks of one process share a GPU and a CUDA context, where a kernel spinning on a flag can
// starve the very work it waits for (lazy module loading and allocator calls synchronise the context): there the barrier
// is taken on the host - the pull kernels and the pointer logic they exercise are the same.
static void shard_peer_barrier(pse_engine* e, int chan) {
    ShardState* s = e->shard;
    s->collectives++;
    if (s->comm.lw) {
        cudaStreamSynchronize(e->stream);
        pthread_barrier_wait(&s->comm.lw->bar);
        return;
    }
    PeerSync ps = s->psync;
    ps.chan = chan;
    peer_barrier_kernel<<<1, 32, 0, e->stream>>>(ps, ++s->epoch[chan]); LAUNCHED(e);
}

// boundary rows of the vector about to be multiplied: my first KH layers go to the left neighbour, my last KH layers to the
// right one; theirs arrive in the rows they occupy in the (global) slot numbering
static int shard_exchange_px(pse_engine* e) {
    ShardState* s = e->shard;
    const ShardGeom& g = s->g;
    ProfScope ps(e, PH_COMM_VEC);
    cudaStream_t st = e->stream;
    const int r = s->rank, left = (r + g.world - 1) % g.world, right = (r + 1) % g.world;
    const uint32_t nsL = g.SL1[r] - g.ROW[r], nsR = g.ROW[r + 1] - g.SR0[r];
    const uint32_t nrR = g.SL1[right] - g.ROW[right], nrL = g.ROW[left + 1] - g.SR0[left];
    s->bytes_sent += ((uint64_t)nsL + nsR) * 16;
    if (s->use_peer) {
        shard_peer_barrier(e, 1);
        if (nrR + nrL) {
            peer_pull_px_kernel<<<nblk(nrR + nrL, 256), 256, 0, st>>>((float4*)e->d_px, (const float4*)s->peer_px[right], g.ROW[right], nrR,
                                                                     (const float4*)s->peer_px[left], g.SR0[left], nrL); LAUNCHED(e);
        }
        return PSE_OK;
    }
    if (nsL) { pack_px_kernel<<<nblk(nsL, 256), 256, 0, st>>>(e->d_px, g.ROW[r], nsL, s->d_vsL); LAUNCHED(e); }
    if (nsR) { pack_px_kernel<<<nblk(nsR, 256), 256, 0, st>>>(e->d_px, g.SR0[r], nsR, s->d_vsR); LAUNCHED(e); }
    CKCOMM(s->comm.ring_exchange(s->d_vsL, (size_t)nsL * 16, s->d_vrR, (size_t)nrR * 16, s->d_vsR, (size_t)nsR * 16, s->d_vrL, (size_t)nrL * 16, st));
    if (nrR) { unpack_px_kernel<<<nblk(nrR, 256), 256, 0, st>>>(e->d_px, g.ROW[right], nrR, s->d_vrR); LAUNCHED(e); }
    if (nrL) { unpack_px_kernel<<<nblk(nrL, 256), 256, 0, st>>>(e->d_px, g.SR0[left], nrL, s->d_vrL); LAUNCHED(e); }
    return PSE_OK;
}
static int shard_allreduce2(pse_engine* e) {
    ShardState* s = e->shard;
    ProfScope ps(e, PH_COMM_RED);
    s->bytes_sent += 24 * (s->world - 1);
    if (s->use_peer && !s->comm.lw) {
        PeerSync ps = s->psync;
        ps.chan = 1;
        peer_allreduce3_kernel<<<1, 32, 0, e->stream>>>(ps, ++s->epoch[1], (s->red_count++) & 1u, e->d_red2); LAUNCHED(e);
        s->collectives++;
        return PSE_OK;
    }
    CKCOMM(s->comm.allreduce_sum(e->d_red2, 3, e->stream));
    return PSE_OK;
}
// spreading: what my particles put on planes outside my slab is added into the owners' planes
static int shard_halo_reduce(pse_engine* e) {
    ShardState* s = e->shard;
    const ShardGeom& g = s->g;
    if (g.world == 1) return PSE_OK;
    ProfScope ps(e, PH_COMM_HALO);
    cudaStream_t st = e->stream;
    const unsigned int gb = e->num_sms * 8;
    const size_t bL = sizeof(float) * 3 * g.HL * s->plane, bR = sizeof(float) * 3 * g.HR * s->plane;
    s->bytes_sent += bL + bR;
    if (s->use_peer) {
        const int left = (s->rank + g.world - 1) % g.world, right = (s->rank + 1) % g.world;
        shard_peer_barrier(e, 0);
        // the right neighbour's left halo lands on my last HL planes, the left neighbour's right halo on my first HR planes
        // (distinct planes because a slab is at least as thick as either halo: one launch adds both)
        const PeerPlaneJob ja = {s->peer_grid[right], (size_t)s->nxaq[right] * s->plane, s->BL + s->nown - g.HL, s->BLq[right] - g.HL, g.HL};
        const PeerPlaneJob jb = {s->peer_grid[left], (size_t)s->nxaq[left] * s->plane, s->BL, s->BLq[left] + (g.X[left + 1] - g.X[left]), g.HR};
        if (s->nown >= g.HL + g.HR) {
            peer_planes_kernel<<<2 * gb, 256, 0, st>>>(e->d_grid, s->Gl, s->plane, ja, jb, 1); LAUNCHED(e);
        } else {   // the two target ranges overlap: one after the other
            const PeerPlaneJob none = {nullptr, 0, 0, 0, 0};
            peer_planes_kernel<<<2 * gb, 256, 0, st>>>(e->d_grid, s->Gl, s->plane, ja, none, 1); LAUNCHED(e);
            peer_planes_kernel<<<2 * gb, 256, 0, st>>>(e->d_grid, s->Gl, s->plane, jb, none, 1); LAUNCHED(e);
        }
        return PSE_OK;
    }
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hsL, s->Gl, s->plane, s->BL - g.HL, g.HL, 0); LAUNCHED(e);
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hsR, s->Gl, s->plane, s->BL + s->nown, g.HR, 0); LAUNCHED(e);
    CKCOMM(s->comm.ring_exchange(s->d_hsL, bL, s->d_hrR, bL, s->d_hsR, bR, s->d_hrL, bR, st));
    // from the right neighbour: its left halo = my last HL planes; from the left neighbour: its right halo = my first HR planes
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hrR, s->Gl, s->plane, s->BL + s->nown - g.HL, g.HL, 2); LAUNCHED(e);
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hrL, s->Gl, s->plane, s->BL, g.HR, 2); LAUNCHED(e);
    return PSE_OK;
}
// interpolation: the planes my particles reach outside my slab are fetched from their owners
static int shard_halo_fetch(pse_engine* e) {
    ShardState* s = e->shard;
    const ShardGeom& g = s->g;
    if (g.world == 1) return PSE_OK;
    ProfScope ps(e, PH_COMM_HALO);
    cudaStream_t st = e->stream;
    const unsigned int gb = e->num_sms * 8;
    const size_t bL = sizeof(float) * 3 * g.HL * s->plane, bR = sizeof(float) * 3 * g.HR * s->plane;
    s->bytes_sent += bL + bR;
    if (s->use_peer) {
        const int left = (s->rank + g.world - 1) % g.world, right = (s->rank + 1) % g.world;
        shard_peer_barrier(e, 0);
        const PeerPlaneJob ja = {s->peer_grid[right], (size_t)s->nxaq[right] * s->plane, s->BL + s->nown, s->BLq[right], g.HR};
        const PeerPlaneJob jb = {s->peer_grid[left], (size_t)s->nxaq[left] * s->plane, s->BL - g.HL, s->BLq[left] + (g.X[left + 1] - g.X[left]) - g.HL, g.HL};
        peer_planes_kernel<<<2 * gb, 256, 0, st>>>(e->d_grid, s->Gl, s->plane, ja, jb, 0); LAUNCHED(e);
        return PSE_OK;
    }
    // to the left neighbour: my first HR planes (its right halo); to the right neighbour: my last HL planes (its left halo)
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hsL, s->Gl, s->plane, s->BL, g.HR, 0); LAUNCHED(e);
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hsR, s->Gl, s->plane, s->BL + s->nown - g.HL, g.HL, 0); LAUNCHED(e);
    CKCOMM(s->comm.ring_exchange(s->d_hsL, bR, s->d_hrR, bR, s->d_hsR, bL, s->d_hrL, bL, st));
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hrR, s->Gl, s->plane, s->BL + s->nown, g.HR, 1); LAUNCHED(e);
    shard_planes_kernel<<<gb, 256, 0, st>>>(e->d_grid, s->d_hrL, s->Gl, s->plane, s->BL - g.HL, g.HL, 1); LAUNCHED(e);
    return PSE_OK;
}

// ---- wave space on the slab ---------------------------------------------------------------------------------------------------
// position-only part: bin the own particles, W record headers, Gaussian factor rows
static int shard_wbin(pse_engine* e) {
    ShardState* s = e->shard;
    ProfScope ps(e, PH_WBIN);
    cudaStream_t st = e->stream;
    const uint32_t r0 = e->row0, r1 = e->row1, nrows = r1 - r0, nt = e->tg.ntile;
    CK(cudaMemsetAsync(e->d_wcount, 0, sizeof(uint32_t) * (nt + 1), st));
    if (nrows) { wbin_kernel<<<nblk(nrows, 256), 256, 0, st>>>(e->d_spos, r1, e->box, e->wp, e->tg, e->d_org, e->d_wcell_of, e->d_wcount, 0, 1 << 30, -1, r0); LAUNCHED(e); }
    CKRC(exclusive_scan(e, e->d_wcount, e->d_wstart, nt + 1, e->d_scan_tmp));
    CK(cudaMemsetAsync(e->d_wcount, 0, sizeof(uint32_t) * (nt + 1), st));
    if (!nrows) return PSE_OK;
    cell_fill_kernel<<<nblk(nrows, 256), 256, 0, st>>>(e->d_wcell_of, r1, e->d_wstart, e->d_wcount, e->d_wtmp, r0); LAUNCHED(e);
    cell_sort_block_kernel<<<nt, 128, 0, st>>>(e->d_wstart, e->d_wtmp, e->d_wperm); LAUNCHED(e);
    // W records carry the SLOT as particle id: velocities are collected in slot order (d_uslot)
    launch_wrecords(e->wp.P, st, e->d_spos, e->d_org, e->d_wperm, nullptr, nrows, e->box, e->wp, e->tg, e->d_wrecs); LAUNCHED(e);
    if (s->world > 1) {
        shard_cover_kernel<<<nblk(nrows, 256), 256, 0, st>>>(e->d_org, r0, r1, e->wp.Nx, s->xorg, s->nxa, e->wp.P, e->d_flag + 1); LAUNCHED(e);
    }
    return PSE_OK;
}
// the forces of the call into the W records
static int shard_wforce(pse_engine* e, const float4* sF) {
    ProfScope ps(e, PH_WBIN);
    const uint32_t nrows = e->row1 - e->row0;
    if (!nrows) return PSE_OK;
    wgather_kernel<<<nblk(nrows, 256), 256, 0, e->stream>>>(e->d_spos, sF, e->d_org, e->d_wperm, nullptr, nrows, e->d_wpos, e->d_wF, e->d_worg, e->d_wid, e->tg,
                                                        reinterpret_cast<int4*>(e->d_wrecs), nullptr, 2); LAUNCHED(e);
    return PSE_OK;
}

static int shard_wave(pse_engine* e, const float4* sF, bool det, bool noise, const float* d_u_grid) {
    ShardState* s = e->shard;
    const ShardGeom& g = s->g;
    const WaveParams& wp = e->wp;
    cudaStream_t st = e->stream;
    const int nown = s->nown, nyl = g.YS[s->rank + 1] - g.YS[s->rank];
    const size_t scomp = (size_t)nown * wp.Ny * wp.Nzp;
    const uint32_t nrows = (uint32_t)nown * wp.Ny;
    ShardBounds b;
    b.world = g.world;
    for (int r = 0; r <= g.world; ++r) { b.xs[r] = g.X[r]; b.ys[r] = g.YS[r]; }
    PeerSlabs pb;
    PeerPtrs<const float2> pp_sloc, pp_tr;
    for (int r = 0; r <= g.world; ++r) { pb.xs[r] = g.X[r]; pb.ys[r] = g.YS[r]; }
    PeerPtrs<float2> pw_sloc, pw_tr;
    for (int r = 0; r < g.world; ++r) {
        pp_sloc.p[r] = s->peer_sloc[r]; pp_tr.p[r] = s->peer_tr[r];
        pw_sloc.p[r] = const_cast<float2*>(s->peer_sloc[r]); pw_tr.p[r] = const_cast<float2*>(s->peer_tr[r]);
    }
    if (det) {
        CKRC(shard_wforce(e, sF));
        {
            ProfScope ps(e, PH_SPREAD);
            CK(cudaMemsetAsync(e->d_grid, 0, sizeof(float) * 3 * s->Gl, st));
            launch_spread2(wp.P, e->spread_var, st, e->d_wrecs, e->d_wstart, wp, e->tg, e->d_grid); LAUNCHED(e);
        }
        CKRC(shard_halo_reduce(e));
        {
            ProfScope ps(e, PH_FFT_FWD);   // z and y passes of the own planes; the y index leaves in digit-reversed order
            for (int c = 0; c < 3; ++c) {
                fft_z_forward_kernel<<<nblk(nrows, 2 * FFT_Z_COLS), FFT_THREADS, fft_smem_bytes(wp.Nz, FFT_Z_COLS + 1), st>>>(
                    e->d_grid + c * s->Gl + (size_t)s->BL * s->plane, s->d_sloc + c * scomp, e->fft_ax[2], nrows, wp.Nzp); LAUNCHED(e);
            }
            fft_y_kernel<false><<<dim3(nblk(wp.Nzh, FFT_Y_COLS), 3 * nown), FFT_THREADS, fft_smem_bytes(wp.Ny, FFT_Y_CP), st>>>(
                s->d_sloc, e->fft_ax[1], wp.Nzh, wp.Nzp); LAUNCHED(e);
            e->fft_execs++;
        }
        ProfScope ps(e, PH_COMM_TRANS);   // transpose: x slabs -> y slabs
        if (s->use_peer && s->push) {
            // (the peers' y slabs are free: everybody passed the barrier that ended the previous evaluation)
            s->bytes_sent += s->a2a_send_off[g.world] - (s->a2a_send_off[s->rank + 1] - s->a2a_send_off[s->rank]);
            peer_push_trans_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_sloc, pw_tr, pb, wp.Nx, wp.Ny, g.X[s->rank], nown, wp.Nzp); LAUNCHED(e);
            shard_peer_barrier(e, 0);
        } else if (s->use_peer) {
            shard_peer_barrier(e, 0);
            s->bytes_sent += s->a2a_recv_off[g.world] - (s->a2a_recv_off[s->rank + 1] - s->a2a_recv_off[s->rank]);
            if (nyl > 0) {
                peer_pull_trans_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_tr, pp_sloc, pb, wp.Nx, wp.Ny, g.YS[s->rank], nyl, wp.Nzp); LAUNCHED(e);
            }
        } else if (g.world > 1) {
            shard_slab_blocks_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_sloc, (float2*)s->d_a2a_a, b, nown, wp.Ny, wp.Nzp, 1); LAUNCHED(e);
            CKCOMM(s->comm.alltoallv(s->d_a2a_a, s->a2a_send_off, s->d_a2a_b, s->a2a_recv_off, st));
            s->bytes_sent += s->a2a_send_off[g.world] - (s->a2a_send_off[s->rank + 1] - s->a2a_send_off[s->rank]);
            if (nyl > 0) { shard_trans_blocks_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_tr, (float2*)s->d_a2a_b, b, wp.Nx, nyl, wp.Nzp, 0); LAUNCHED(e); }
        } else {
            CK(cudaMemcpyAsync(s->d_tr, s->d_sloc, sizeof(float2) * 3 * scomp, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (nyl > 0) {
        ProfScope ps(e, PH_SCALE);   // x forward + scaling (+ random modes) + x inverse of the own stored-y rows in one kernel
        fft_x_scale_kernel<<<dim3(nblk(wp.Nzh, FFT_X_COLS), nyl), FFT_THREADS, fft_smem_bytes(wp.Nx, FFT_X_CP), st>>>(
            s->d_tr, e->fft_ax[0], e->fft_ax[1].freq_of, e->wp, e->box, det ? 1 : 0, noise ? 1 : 0, e->d_stepdev, d_u_grid, g.YS[s->rank], nyl); LAUNCHED(e);
        e->fft_execs++;
    }
    {
        ProfScope ps(e, PH_COMM_TRANS);    // transpose back: y slabs -> x slabs
        if (s->use_peer && s->push) {
            // (the peers' x slabs are free: every rank finished pushing out of its own before the barrier above;
            // without a deterministic part nothing was pushed and the x slabs were last read before the previous end barrier)
            s->bytes_sent += s->a2a_recv_off[g.world] - (s->a2a_recv_off[s->rank + 1] - s->a2a_recv_off[s->rank]);
            if (nyl > 0) { peer_push_slab_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_tr, pw_sloc, pb, wp.Nx, wp.Ny, g.YS[s->rank], nyl, wp.Nzp); LAUNCHED(e); }
            shard_peer_barrier(e, 0);
        } else if (s->use_peer) {
            shard_peer_barrier(e, 0);
            s->bytes_sent += s->a2a_send_off[g.world] - (s->a2a_send_off[s->rank + 1] - s->a2a_send_off[s->rank]);
            peer_pull_slab_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_sloc, pp_tr, pb, wp.Nx, wp.Ny, g.X[s->rank], nown, wp.Nzp); LAUNCHED(e);
        } else if (g.world > 1) {
            if (nyl > 0) { shard_trans_blocks_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_tr, (float2*)s->d_a2a_b, b, wp.Nx, nyl, wp.Nzp, 1); LAUNCHED(e); }
            CKCOMM(s->comm.alltoallv(s->d_a2a_b, s->a2a_recv_off, s->d_a2a_a, s->a2a_send_off, st));
            s->bytes_sent += s->a2a_recv_off[g.world] - (s->a2a_recv_off[s->rank + 1] - s->a2a_recv_off[s->rank]);
            shard_slab_blocks_kernel<<<e->num_sms * 8, 256, 0, st>>>(s->d_sloc, (float2*)s->d_a2a_a, b, nown, wp.Ny, wp.Nzp, 0); LAUNCHED(e);
        } else {
            CK(cudaMemcpyAsync(s->d_sloc, s->d_tr, sizeof(float2) * 3 * scomp, cudaMemcpyDeviceToDevice, st));
        }
    }
    {
        ProfScope ps(e, PH_FFT_INV);
        fft_y_kernel<true><<<dim3(nblk(wp.Nzh, FFT_Y_COLS), 3 * nown), FFT_THREADS, fft_smem_bytes(wp.Ny, FFT_Y_CP), st>>>(
            s->d_sloc, e->fft_ax[1], wp.Nzh, wp.Nzp); LAUNCHED(e);
        for (int c = 0; c < 3; ++c) {
            fft_z_inverse_kernel<<<nblk(nrows, 2 * FFT_Z_COLS), FFT_THREADS, fft_smem_bytes(wp.Nz, FFT_Z_COLS + 1), st>>>(
                s->d_sloc + c * scomp, e->d_grid + c * s->Gl + (size_t)s->BL * s->plane, e->fft_ax[2], nrows, wp.Nzp); LAUNCHED(e);
        }
        e->fft_execs++;
    }
    CKRC(shard_halo_fetch(e));
    ProfScope ps(e, PH_INTERP);
    launch_interp2(wp.P, st, e->d_wrecs, e->d_wstart, wp, e->tg, e->d_grid, s->d_uslot, 0); LAUNCHED(e);
    return PSE_OK;
}

// ---- one velocity evaluation ------------------------------------------------------------------------------------------------
static int shard_velocity(pse_engine* e, const float4* d_pos, const float4* d_F, float4* d_U, uint32_t timestep, const float* d_u_particles,
                          const float* d_u_grid, unsigned what, int* m_out) {
    ShardState* s = e->shard;
    const ShardGeom& g = s->g;
    cudaStream_t st = e->stream;
    const uint32_t N = e->N;
    const bool wdet = what & SV_DET_WAVE, rdet = what & SV_DET_REAL, wnoise = what & SV_WNOISE, rnoise = what & SV_RNOISE;
    CKRC(shard_connect_peers(e));
    CKRC(ensure_neighbors(e, d_pos));   // replicated decision (identical positions on every rank); the build covers the own rows
    if (rnoise) CKRC(ensure_krylov(e));
    // what needs the positions only comes first (forces still on their way from the host are waited for after it), pruning
    // beside the binning on the second stream when the evaluation has both branches
    const bool head_fork = (rdet || rnoise) && (wdet || wnoise) && e->overlap && !e->prof_on && e->stream2 && e->prune && !e->pruned_valid;
    if (head_fork) {
        CK(cudaEventRecord(e->ev_fork, st));
        CK(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
        e->stream = e->stream2;
        const int prc = ensure_pruned(e);
        e->stream = st;
        if (prc != PSE_OK) return prc;
        CK(cudaEventRecord(e->ev_join, e->stream2));
    } else if (rdet || rnoise) CKRC(ensure_pruned(e));
    if (wdet || wnoise) CKRC(shard_wbin(e));
    if (head_fork) CK(cudaStreamWaitEvent(st, e->ev_join, 0));
    if (e->wait_F) { e->wait_F = false; CK(cudaStreamWaitEvent(st, e->ev_F, 0)); }
    if (e->take_Fnext >= 0) {   // forces prefetched earlier (pse_host_prefetch_forces)
        const int r = e->take_Fnext;
        e->take_Fnext = -1;
        CK(cudaStreamWaitEvent(st, e->ev_Fnext[r], 0));
        copy_f4_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->d_hF, e->d_hF_next[r], e->N); LAUNCHED(e);
        CK(cudaEventRecord(e->ev_Fcons[r], st));
    }
    CKRC(upload_stepdev(e, timestep));
    const uint32_t r0 = e->row0, r1 = e->row1, nrows = r1 - r0;
    // forces in slot order (replicated pass over N: every rank reads its halo rows from the same array)
    if (wdet || rdet) { gather_vec_kernel<<<nblk(N, 256), 256, 0, st>>>(d_F, e->d_perm, N, e->d_sx, (float4*)e->d_px); LAUNCHED(e); }
    // The real-space branch (prune, M_real F, Lanczos with its vector halos and reductions: channel 1) and the wave-space
    // branch (bin, spread, FFTs, transposes, grid halos, interpolate: channel 0) touch disjoint buffers until they meet in the
    // velocity, so they are issued on two streams: the kernels of one branch fill the gaps the other leaves while it waits
    // for its neighbours.  Profiling keeps them serial so that the per-phase times stay meaningful.
    const bool wave = wdet || wnoise, real = rdet || rnoise;
    const bool fork = wave && real && g.world > 1 && s->use_peer && e->overlap && !e->prof_on && e->stream2;
    const int m_batch = lanczos_batch_size(e);
    const bool dual = rdet && rnoise && spmv_dual_available(e);
    int rc = PSE_OK;
    if (fork) {
        CK(cudaEventRecord(e->ev_fork, st));
        CK(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
        e->stream = e->stream2;
    }
    if (rdet && !dual) rc = run_spmv_plain(e, e->d_sy);
    if (rc == PSE_OK && rnoise) rc = lanczos_batch(e, d_u_particles, m_batch, dual);
    if (fork) {
        e->stream = st;
        if (rc == PSE_OK) CK(cudaEventRecord(e->ev_join, e->stream2));
    }
    if (rc != PSE_OK) return rc;
    int acc = 0;
    if (wave) { CKRC(shard_wave(e, e->d_sx, wdet, wnoise, d_u_grid)); acc = 1; }
    if (fork) CK(cudaStreamWaitEvent(st, e->ev_join, 0));
    if (rdet && !rnoise) {
        if (nrows) { scatter_add_kernel<<<nblk(nrows, 256), 256, 0, st>>>(e->d_sy, nullptr, r1, s->d_uslot, acc, r0); LAUNCHED(e); }
        acc = 1;
    }
    if (rnoise) {
        CKRC(lanczos_finish(e, s->d_uslot, acc, m_batch, m_out, rdet ? e->d_sy : nullptr, nullptr));
        acc = 1;
    }
    if (!acc && nrows) CK(cudaMemsetAsync(s->d_uslot + r0, 0, sizeof(float4) * nrows, st));
    {
        ProfScope ps(e, PH_COMM_GATHER);   // every rank gets every velocity: positions stay replicated and bitwise identical
        s->bytes_sent += (uint64_t)(g.world - 1) * nrows * sizeof(float4);
        if (s->use_peer) {
            PeerPtrs<const float4> pu;
            PeerBounds rb;
            for (int r = 0; r < g.world; ++r) pu.p[r] = s->peer_uslot[r];
            for (int r = 0; r <= g.world; ++r) rb.row[r] = g.ROW[r];
            shard_peer_barrier(e, 0);
            peer_gather_scatter_kernel<<<nblk(N, 256), 256, 0, st>>>(pu, rb, g.world, e->d_perm, N, d_U); LAUNCHED(e);
            shard_peer_barrier(e, 0);   // nobody rewrites a buffer a peer may still be reading
        } else {
            size_t off[SHARD_MAX_WORLD + 1];
            for (int r = 0; r <= g.world; ++r) off[r] = (size_t)g.ROW[r] * sizeof(float4);
            if (g.world > 1) CKCOMM(s->comm.allgatherv(s->d_uslot, off, st));
            shard_scatter_kernel<<<nblk(N, 256), 256, 0, st>>>(s->d_uslot, e->d_perm, N, d_U); LAUNCHED(e);
        }
    }
    CK(cudaGetLastError());
    return PSE_OK;
}

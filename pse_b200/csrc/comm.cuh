// Collectives of the slab-decomposed step, issued from C++ on the engine's stream.
//
// NCCL is bound at run time (dlopen) so that the library has no link-time dependency on a particular NCCL build and
// shares the copy the host process already loaded (torch ships its own libnccl.so.2).  Only the handful of C entry
// points used here are declared; their signatures are part of NCCL's stable C API (nccl.h).
// The reference is single-GPU (PSEv1/Stokes.cc:104): everything in this file is new work.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

struct pse_nccl_comm_opaque;
typedef pse_nccl_comm_opaque* pse_nccl_comm_t;
struct pse_nccl_uid { char internal[128]; };   // ncclUniqueId
enum { PSE_NCCL_CHAR = 0, PSE_NCCL_FLOAT = 7, PSE_NCCL_DOUBLE = 8 };  // ncclChar / ncclFloat32 / ncclFloat64
enum { PSE_NCCL_SUM = 0 };

struct NcclApi {
    void* handle;
    int (*GetUniqueId)(pse_nccl_uid*);
    int (*CommInitRank)(pse_nccl_comm_t*, int, pse_nccl_uid, int);
    int (*CommDestroy)(pse_nccl_comm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, pse_nccl_comm_t, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, pse_nccl_comm_t, cudaStream_t);
    int (*Send)(const void*, size_t, int, int, pse_nccl_comm_t, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, pse_nccl_comm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
};

static NcclApi* nccl_api(char* err, size_t errlen) {
    static NcclApi api;
    static int state = 0;  // 0 untried, 1 ok, -1 failed
    if (state == 1) return &api;
    if (state == -1) { snprintf(err, errlen, "NCCL library not available"); return nullptr; }
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the process already uses (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { state = -1; snprintf(err, errlen, "dlopen(libnccl.so.2) failed: %s", dlerror()); return nullptr; }
    api.handle = h;
    bool ok = true;
#define BIND(field, name) do { *(void**)(&api.field) = dlsym(h, name); if (!api.field) { ok = false; snprintf(err, errlen, "NCCL symbol %s missing", name); } } while (0)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllReduce, "ncclAllReduce");
    BIND(AllGather, "ncclAllGather");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    state = ok ? 1 : -1;
    return ok ? &api : nullptr;
}

// ---- in-process world (virtual ranks): W engines of one process, one host thread each ------------------------------------
// Every collective is: publish the send descriptors, synchronise the own stream, host barrier, copy what the peers
// addressed to this rank (device copies on the own stream), synchronise, host barrier.  Deterministic, slow, and only there
// so that the multi-rank code path can be checked against the single-domain engine on ONE GPU (tests/test_gpu_parity.py).
#define PSE_COMM_MAX_WORLD 16
#define PSE_PEER_NBUF 6   // buffers a rank exposes to its peers: pad, px, uslot, grid, sloc, tr
struct pse_local_world {
    int world;
    pthread_barrier_t bar;
    void* ptrs[PSE_COMM_MAX_WORLD][PSE_PEER_NBUF];   // peer-memory transport: plain pointers (one process, one address space)
    struct Desc { const void* a; const void* b; const size_t* off; } desc[PSE_COMM_MAX_WORLD];
    double red[PSE_COMM_MAX_WORLD][64];
};

// One communicator on one stream.  world == 1 needs no transport at all: every exchange degenerates to a device copy.
struct PseComm {
    NcclApi* api;
    pse_nccl_comm_t comm;
    pse_local_world* lw;
    int rank, world;
    char err[256];

    int init(int rank_, int world_, const uint8_t* uid128, pse_local_world* local) {
        rank = rank_; world = world_; comm = nullptr; api = nullptr; lw = nullptr; err[0] = 0;
        if (world == 1) return 0;
        if (local) {
            if (local->world != world) { snprintf(err, sizeof(err), "local world has %d ranks, not %d", local->world, world); return -1; }
            lw = local;
            return 0;
        }
        if (!uid128) { snprintf(err, sizeof(err), "no NCCL unique id and no local world"); return -1; }
        api = nccl_api(err, sizeof(err));
        if (!api) return -1;
        pse_nccl_uid id;
        memcpy(id.internal, uid128, 128);
        const int rc = api->CommInitRank(&comm, world, id, rank);
        if (rc != 0) { snprintf(err, sizeof(err), "ncclCommInitRank: %s", api->GetErrorString(rc)); return -1; }
        return 0;
    }
    void destroy() { if (comm && api) api->CommDestroy(comm); comm = nullptr; }
    int ck(int rc, const char* what) {
        if (rc != 0) { snprintf(err, sizeof(err), "%s: %s", what, api ? api->GetErrorString(rc) : "?"); return -1; }
        return 0;
    }
    int cu(cudaError_t ce, const char* what) {
        if (ce != cudaSuccess) { snprintf(err, sizeof(err), "%s: %s", what, cudaGetErrorString(ce)); return -1; }
        return 0;
    }
    // local world: descriptors published + own stream drained + everybody arrived
    int lw_open(const void* a, const void* b, const size_t* off, cudaStream_t st) {
        lw->desc[rank].a = a; lw->desc[rank].b = b; lw->desc[rank].off = off;
        const int rc = cu(cudaStreamSynchronize(st), "local world: stream synchronise");
        pthread_barrier_wait(&lw->bar);
        return rc;
    }
    int lw_close(cudaStream_t st) {
        const int rc = cu(cudaStreamSynchronize(st), "local world: stream synchronise");
        pthread_barrier_wait(&lw->bar);
        return rc;
    }
    // sum of n doubles over ranks, in place (identical bits on every rank)
    int allreduce_sum(double* d, size_t n, cudaStream_t st) {
        if (world == 1) return 0;
        if (lw) {
            if (n > 64) { snprintf(err, sizeof(err), "local all-reduce handles at most 64 words"); return -1; }
            int rc = cu(cudaMemcpyAsync(lw->red[rank], d, n * sizeof(double), cudaMemcpyDeviceToHost, st), "local all-reduce");
            if (lw_open(nullptr, nullptr, nullptr, st)) rc = -1;
            double sum[64];
            for (size_t i = 0; i < n; ++i) { double a = 0.0; for (int q = 0; q < world; ++q) a += lw->red[q][i]; sum[i] = a; }
            if (!rc) rc = cu(cudaMemcpyAsync(d, sum, n * sizeof(double), cudaMemcpyHostToDevice, st), "local all-reduce");
            if (lw_close(st)) rc = -1;
            return rc;
        }
        return ck(api->AllReduce(d, d, n, PSE_NCCL_DOUBLE, PSE_NCCL_SUM, comm, st), "ncclAllReduce");
    }
    // In-place all-gather of unequal blocks: block q of `buf` (bytes off[q] .. off[q + 1]) is owned by rank q and ends up
    // on every rank.
    int allgatherv(void* buf, const size_t* off, cudaStream_t st) {
        if (world == 1) return 0;
        if (lw) {
            int rc = lw_open(buf, nullptr, nullptr, st);
            for (int q = 0; q < world && !rc; ++q) {
                const size_t n = off[q + 1] - off[q];
                if (q != rank && n) rc = cu(cudaMemcpyAsync((char*)buf + off[q], (const char*)lw->desc[q].a + off[q], n, cudaMemcpyDeviceToDevice, st), "local all-gather");
            }
            if (lw_close(st)) rc = -1;
            return rc;
        }
        const size_t mine = off[rank + 1] - off[rank];
        int rc = api->GroupStart();
        for (int q = 0; q < world && !rc; ++q) {
            if (q == rank) continue;
            if (mine) rc = api->Send((const char*)buf + off[rank], mine, PSE_NCCL_CHAR, q, comm, st);
            const size_t n = off[q + 1] - off[q];
            if (!rc && n) rc = api->Recv((char*)buf + off[q], n, PSE_NCCL_CHAR, q, comm, st);
        }
        const int rc2 = api->GroupEnd();
        return ck(rc ? rc : rc2, "all-gather (ncclSend/ncclRecv)");
    }
    // Ring neighbour exchange: message A goes to rank - 1 and is received from rank + 1; message B goes to rank + 1 and is
    // received from rank - 1.  Issue order (sends: to-left, to-right; receives: from-right, from-left) keeps the two
    // messages apart when both neighbours are the same rank (world == 2); world == 1 copies on the device.
    int ring_exchange(const void* sendL, size_t bytesL_send, void* recvR, size_t bytesR_recv, const void* sendR, size_t bytesR_send,
                      void* recvL, size_t bytesL_recv, cudaStream_t st) {
        if (world == 1) {
            if (bytesL_send && cudaMemcpyAsync(recvR, sendL, bytesL_send, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
            if (bytesR_send && cudaMemcpyAsync(recvL, sendR, bytesR_send, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
            return 0;
        }
        const int left = (rank + world - 1) % world, right = (rank + 1) % world;
        if (lw) {
            int rc = lw_open(sendL, sendR, nullptr, st);
            if (!rc && bytesR_recv) rc = cu(cudaMemcpyAsync(recvR, lw->desc[right].a, bytesR_recv, cudaMemcpyDeviceToDevice, st), "local ring exchange");
            if (!rc && bytesL_recv) rc = cu(cudaMemcpyAsync(recvL, lw->desc[left].b, bytesL_recv, cudaMemcpyDeviceToDevice, st), "local ring exchange");
            if (lw_close(st)) rc = -1;
            return rc;
        }
        int rc = api->GroupStart();
        if (!rc && bytesL_send) rc = api->Send(sendL, bytesL_send, PSE_NCCL_CHAR, left, comm, st);
        if (!rc && bytesR_send) rc = api->Send(sendR, bytesR_send, PSE_NCCL_CHAR, right, comm, st);
        if (!rc && bytesR_recv) rc = api->Recv(recvR, bytesR_recv, PSE_NCCL_CHAR, right, comm, st);
        if (!rc && bytesL_recv) rc = api->Recv(recvL, bytesL_recv, PSE_NCCL_CHAR, left, comm, st);
        const int rc2 = api->GroupEnd();
        return ck(rc ? rc : rc2, "ring exchange (ncclSend/ncclRecv)");
    }
    // personalised all-to-all: block q of `send` (send_off[q] .. send_off[q + 1], bytes) goes to rank q
    int alltoallv(const void* send, const size_t* send_off, void* recv, const size_t* recv_off, cudaStream_t st) {
        if (world == 1) {
            const size_t n = send_off[1] - send_off[0];
            return !n || cudaMemcpyAsync(recv, send, n, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : -1;
        }
        if (lw) {
            int rc = lw_open(send, nullptr, send_off, st);
            for (int q = 0; q < world && !rc; ++q) {
                const size_t n = recv_off[q + 1] - recv_off[q];
                if (n) rc = cu(cudaMemcpyAsync((char*)recv + recv_off[q], (const char*)lw->desc[q].a + lw->desc[q].off[rank], n, cudaMemcpyDeviceToDevice, st), "local all-to-all");
            }
            if (lw_close(st)) rc = -1;
            return rc;
        }
        int rc = api->GroupStart();
        for (int q = 0; q < world && !rc; ++q) {
            const size_t ns = send_off[q + 1] - send_off[q], nr = recv_off[q + 1] - recv_off[q];
            if (ns) rc = api->Send((const char*)send + send_off[q], ns, PSE_NCCL_CHAR, q, comm, st);
            if (!rc && nr) rc = api->Recv((char*)recv + recv_off[q], nr, PSE_NCCL_CHAR, q, comm, st);
        }
        const int rc2 = api->GroupEnd();
        return ck(rc ? rc : rc2, "all-to-all (ncclSend/ncclRecv)");
    }
};

// libyachtgpu -- the multi-GPU plumbing of the sharded train path: one rank per GPU (threads of one process, as the
// drop-in executable runs them, or one process per GPU, as torchrun launches bench.py).
//
//   * NCCL (dlopen'ed: libnccl.so.2, no link-time dependency, so single-GPU users never load it) carries the small
//     control exchanges -- histograms, stream lengths, statistics, the pair lists -- and doubles as the cross-rank
//     barrier between phases;
//   * the bulk exchanges of the sharded step (work items and the posting stream of large groups out of the grouping
//     kernel; with genome-range residency also the packed words of the level-1 partition) are NOT collectives: the
//     producing kernels store straight into the peers' buffers over NVLink (index_msd.cu).  This file hands them the peer
//     pointers: cudaIpc handles between processes, plain pointers plus cudaDeviceEnablePeerAccess between threads of one
//     process.
//
// The reference has no counterpart: its only parallelism is std::thread row chunks (src/cpp/main.cpp:338-349).
#include "common.cuh"
#include "comm.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <unistd.h>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

struct NcclApi {
    void* dl = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    std::string err;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.dl = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.dl) break;
    }
    if (!api.dl) {
        api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return api;
    }
#define YG_SYM(f)                                                       \
    api.f = (decltype(api.f))dlsym(api.dl, "nccl" #f);                  \
    if (!api.f) { api.err = "libnccl lacks nccl" #f; return api; }
    YG_SYM(GetUniqueId) YG_SYM(CommInitRank) YG_SYM(CommDestroy) YG_SYM(AllGather) YG_SYM(AllReduce) YG_SYM(GetErrorString)
#undef YG_SYM
    api.ok = true;
    return api;
}

struct ShareRec {
    int pid;
    int dev;
    unsigned long long ptr;
    unsigned long long host;      // gethostid(): IPC handles only travel inside one host
    cudaIpcMemHandle_t h;
};

}  // namespace

#define YG_NCCL(ctx, call)                                                                                          \
    do {                                                                                                            \
        ncclResult_t r__ = (call);                                                                                  \
        if (r__ != ncclSuccess)                                                                                     \
            return ygpu_fail((ctx), YGPU_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, nccl_api().GetErrorString(r__)); \
    } while (0)

extern "C" int ygpu_comm_get_unique_id(uint8_t* id) {
    if (!id) return YGPU_ERR_ARG;
    NcclApi& api = nccl_api();
    if (!api.ok) return ygpu_fail(nullptr, YGPU_ERR_STATE, "%s", api.err.c_str());
    static_assert(sizeof(ncclUniqueId) <= YGPU_COMM_ID_BYTES, "unique id does not fit");
    ncclUniqueId u;
    memset(id, 0, YGPU_COMM_ID_BYTES);
    if (api.GetUniqueId(&u) != ncclSuccess) return ygpu_fail(nullptr, YGPU_ERR_CUDA, "ncclGetUniqueId failed");
    memcpy(id, &u, sizeof u);
    return 0;
}

extern "C" int ygpu_comm_init(ygpu_ctx* ctx, int rank, int nranks, const uint8_t* id) {
    if (!ctx || !id || nranks < 1 || nranks > YG_MAX_RANKS || rank < 0 || rank >= nranks) return YGPU_ERR_ARG;
    NcclApi& api = nccl_api();
    if (!api.ok) return ygpu_fail(ctx, YGPU_ERR_STATE, "%s", api.err.c_str());
    YG_CUDA(ctx, cudaSetDevice(ctx->device));
    ygpu_comm_destroy(ctx);
    // Only a few hundred bytes per collective ever go through NCCL here; its NVLink-SHARP (NVLS) set-up costs seconds of
    // communicator start-up on NVSwitch systems and buys nothing for them (not overridden when the caller has set it).
    setenv("NCCL_NVLS_ENABLE", "0", 0);
    ygpu_comm* c = new (std::nothrow) ygpu_comm();
    if (!c) return ygpu_fail(ctx, YGPU_ERR_NOMEM, "out of host memory");
    c->rank = rank;
    c->nranks = nranks;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = api.CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return ygpu_fail(ctx, YGPU_ERR_CUDA, "ncclCommInitRank(rank %d of %d): %s", rank, nranks, api.GetErrorString(r));
    }
    ctx->comm = c;
    if (cudaMalloc(&c->d_xchg, (size_t)YG_MAX_RANKS * 1024) != cudaSuccess) return ygpu_fail(ctx, YGPU_ERR_NOMEM, "cudaMalloc(comm scratch)");
    return 0;
}

extern "C" int ygpu_comm_destroy(ygpu_ctx* ctx) {
    if (!ctx) return YGPU_ERR_ARG;
    ygpu_comm* c = ctx->comm;
    if (!c) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->d_xchg) cudaFree(c->d_xchg);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    delete c;
    ctx->comm = nullptr;
    return 0;
}

int ygpu_comm_allgather(ygpu_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
    YG_NCCL(ctx, nccl_api().AllGather(send, recv, bytes_per_rank, ncclChar, ctx->comm->comm, ctx->stream));
    ctx->tm.n_library_launches++;
    return 0;
}
int ygpu_comm_allreduce_u64(ygpu_ctx* ctx, const void* send, void* recv, size_t count, bool is_max) {
    YG_NCCL(ctx, nccl_api().AllReduce(send, recv, count, ncclUint64, is_max ? ncclMax : ncclSum, ctx->comm->comm, ctx->stream));
    ctx->tm.n_library_launches++;
    return 0;
}

// Collective: every rank passes the base pointer of one of its device allocations; peers[q] becomes a pointer through
// which THIS rank's kernels can store into rank q's allocation (peers[rank] = local).
int ygpu_comm_share(ygpu_ctx* ctx, void* local, void** peers) {
    ygpu_comm* c = ctx->comm;
    const int N = c->nranks;
    ShareRec mine{};
    mine.pid = (int)getpid();
    mine.dev = ctx->device;
    mine.ptr = (unsigned long long)local;
    mine.host = (unsigned long long)gethostid();
    YG_CUDA(ctx, cudaIpcGetMemHandle(&mine.h, local));
    static_assert(sizeof(ShareRec) <= 512, "ShareRec too large");
    char* d = (char*)c->d_xchg;            // [0, 512): mine; [512 ...): gathered (N x 512 <= YG_MAX_RANKS x 512)
    YG_CUDA(ctx, cudaMemcpyAsync(d, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
    YG_CHECK(ygpu_comm_allgather(ctx, d, d + 512, 512 / 2));      // 256 bytes per rank hold a ShareRec
    static_assert(sizeof(ShareRec) <= 256, "ShareRec must fit the 256-byte slot");
    std::vector<char> all((size_t)N * 256);
    YG_CUDA(ctx, cudaMemcpyAsync(all.data(), d + 512, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    YG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < N; q++) {
        ShareRec rec;
        memcpy(&rec, all.data() + (size_t)q * 256, sizeof rec);
        if (q == c->rank) { peers[q] = local; continue; }
        if (rec.host != mine.host) return ygpu_fail(ctx, YGPU_ERR_STATE, "rank %d runs on another host: the sharded path is single-node", q);
        if (rec.pid == mine.pid) {
            if (!c->peer_enabled[q]) {
                int can = 0;
                YG_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, rec.dev));
                if (!can) return ygpu_fail(ctx, YGPU_ERR_STATE, "GPU %d cannot access GPU %d (no NVLink/P2P path)", ctx->device, rec.dev);
                cudaError_t e = cudaDeviceEnablePeerAccess(rec.dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return ygpu_fail(ctx, YGPU_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", rec.dev, cudaGetErrorString(e));
                cudaGetLastError();
                c->peer_enabled[q] = true;
            }
            peers[q] = (void*)rec.ptr;
        } else {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, rec.h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return ygpu_fail(ctx, YGPU_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e));
            c->ipc_opened.push_back(p);
            peers[q] = p;
        }
    }
    return 0;
}

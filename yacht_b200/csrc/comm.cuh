// Internal: the communicator of the sharded multi-GPU train path (comm.cu).
#pragma once
#include <nccl.h>
#include <vector>

#ifndef YG_MAX_RANKS
#define YG_MAX_RANKS 16
#endif

struct ygpu_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    void* d_xchg = nullptr;                 // small device scratch for the control exchanges
    bool peer_enabled[YG_MAX_RANKS] = {};
    std::vector<void*> ipc_opened;
};

struct ygpu_ctx;
int ygpu_comm_allgather(ygpu_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);
int ygpu_comm_allreduce_u64(ygpu_ctx* ctx, const void* send, void* recv, size_t count, bool is_max);
int ygpu_comm_share(ygpu_ctx* ctx, void* local, void** peers);

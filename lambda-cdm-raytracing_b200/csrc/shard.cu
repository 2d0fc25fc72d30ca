// shard.cu -- target-sharded data parallelism for hosts that do not bring their own
// communicator (the C++ plugin side).  One process (or thread) per GPU;
// rank r owns particles [r*N/G, (r+1)*N/G); each step the float4 (x,y,z,m) shards
// are all-gathered IN PLACE into the full source array with NCCL over NVLink.
//
// Semantic ancestor: ClusterCommunicator::gather_all_particles
// (reference src/mpi/cluster_comm.cpp:218-247: MPI_Allgather of counts +
// MPI_Allgatherv of host Particle structs).  Here the counts follow from (N, G),
// the buffers are device-resident and no host staging happens.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): inside a PyTorch process
// this binds to the copy torch already loaded, in a plain C++ host to the system
// one, and libb200grav.so itself keeps no link-time NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "shard.cuh"

namespace b200 {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
#define B200_SYM(field, name) \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name); \
    if (!g_nccl.field) return;
    B200_SYM(GetUniqueId, "ncclGetUniqueId")
    B200_SYM(CommInitRank, "ncclCommInitRank")
    B200_SYM(CommDestroy, "ncclCommDestroy")
    B200_SYM(AllGather, "ncclAllGather")
    B200_SYM(Broadcast, "ncclBroadcast")
    B200_SYM(AllReduce, "ncclAllReduce")
    B200_SYM(GroupStart, "ncclGroupStart")
    B200_SYM(GroupEnd, "ncclGroupEnd")
    B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
    g_nccl.ok = true;
}

const NcclApi* nccl() {
    std::call_once(g_nccl_once, load_nccl);
    return g_nccl.ok ? &g_nccl : nullptr;
}

#define B200_NCCL(call)                                       \
    do {                                                      \
        ncclResult_t r__ = (call);                            \
        if (r__ != ncclSuccess) return 2000 + (int)r__;       \
    } while (0)

}  // namespace

struct ShardState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

const char* shard_error_string(int nccl_result) {
    const NcclApi* api = nccl();
    return api ? api->GetErrorString((ncclResult_t)nccl_result) : "NCCL error (libnccl.so.2 not loadable)";
}

int shard_unique_id(unsigned char id[B200_SHARD_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == B200_SHARD_ID_BYTES, "ncclUniqueId size");
    const NcclApi* api = nccl();
    if (!api) return B200_ERR_UNSUPPORTED;
    ncclUniqueId u;
    B200_NCCL(api->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return B200_OK;
}

int shard_init(b200_ctx* ctx, const unsigned char id[B200_SHARD_ID_BYTES], int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) return B200_ERR_INVALID;
    if (ctx->shard) return B200_ERR_STATE;
    ShardState* s = new (std::nothrow) ShardState();
    if (!s) return B200_ERR_NOMEM;
    s->rank = rank;
    s->world = world;
    if (world > 1) {
        const NcclApi* api = nccl();
        if (!api) { delete s; return B200_ERR_UNSUPPORTED; }
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        ncclResult_t r = api->CommInitRank(&s->comm, world, u, rank);
        if (r != ncclSuccess) { delete s; return 2000 + (int)r; }
    }
    ctx->shard = s;
    return B200_OK;
}

int shard_finalize(b200_ctx* ctx) {
    ShardState* s = ctx->shard;
    if (!s) return B200_OK;
    if (s->comm) {
        const NcclApi* api = nccl();
        if (api) api->CommDestroy(s->comm);
    }
    delete s;
    ctx->shard = nullptr;
    return B200_OK;
}

int shard_info(const b200_ctx* ctx, int* rank, int* world) {
    const ShardState* s = ctx->shard;
    if (rank) *rank = s ? s->rank : 0;
    if (world) *world = s ? s->world : 1;
    return B200_OK;
}

// posm4_full: float4[n_total] on this rank's device, holding this rank's particles at
// [i0, i0 + n_local) (shard_range); on return (stream-ordered) every rank's slice is filled in.
int shard_allgather(b200_ctx* ctx, void* posm4_full, size_t n_total, cudaStream_t st) {
    ShardState* s = ctx->shard;
    if (!s) return B200_ERR_STATE;
    if (s->world == 1 || n_total == 0) return B200_OK;
    const NcclApi* api = nccl();
    if (!api) return B200_ERR_UNSUPPORTED;
    float* base = (float*)posm4_full;
    const size_t G = (size_t)s->world;
    if (n_total % G == 0) {
        // equal shards: one in-place ncclAllGather (send = recv + rank * count)
        const size_t cnt = n_total / G * 4;
        B200_NCCL(api->AllGather(base + (size_t)s->rank * cnt, base, cnt, ncclFloat, s->comm, st));
        return B200_OK;
    }
    // ragged shards (sizes differ by at most one particle): one broadcast per owner, grouped
    B200_NCCL(api->GroupStart());
    for (int r = 0; r < s->world; ++r) {
        size_t lo = 0, cnt = 0;
        b200_shard_range(n_total, r, s->world, &lo, &cnt);
        if (cnt == 0) continue;
        ncclResult_t e = api->Broadcast(base + lo * 4, base + lo * 4, cnt * 4, ncclFloat, r, s->comm, st);
        if (e != ncclSuccess) { api->GroupEnd(); return 2000 + (int)e; }
    }
    B200_NCCL(api->GroupEnd());
    return B200_OK;
}

// Raw byte collectives for the tree-table exchange (tree.cu): every rank contributes `bytes` bytes.
int shard_allgather_bytes(b200_ctx* ctx, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    ShardState* s = ctx->shard;
    if (!s || s->world == 1) return B200_ERR_STATE;
    const NcclApi* api = nccl();
    if (!api) return B200_ERR_UNSUPPORTED;
    B200_NCCL(api->AllGather(send, recv, bytes, ncclChar, s->comm, st));
    return B200_OK;
}

// `count` broadcasts in one NCCL group: item i goes from send[i] on rank root[i] to recv[i] on every rank
// (send[i] is read on the root only; the root's recv[i] receives a copy as well).
int shard_bcast_group(b200_ctx* ctx, int count, const void* const* send, void* const* recv, const size_t* bytes,
                      const int* root, cudaStream_t st) {
    ShardState* s = ctx->shard;
    if (!s || s->world == 1) return B200_ERR_STATE;
    const NcclApi* api = nccl();
    if (!api) return B200_ERR_UNSUPPORTED;
    B200_NCCL(api->GroupStart());
    for (int i = 0; i < count; ++i) {
        if (bytes[i] == 0) continue;
        ncclResult_t e = api->Broadcast(send[i], recv[i], bytes[i], ncclChar, root[i], s->comm, st);
        if (e != ncclSuccess) { api->GroupEnd(); return 2000 + (int)e; }
    }
    B200_NCCL(api->GroupEnd());
    return B200_OK;
}

// Sum `count` host doubles over the ranks (diagnostics: energies).  Blocking.
int shard_allreduce_f64(b200_ctx* ctx, double* values, size_t count) {
    ShardState* s = ctx->shard;
    if (!s || s->world == 1 || count == 0) return B200_OK;
    const NcclApi* api = nccl();
    if (!api) return B200_ERR_UNSUPPORTED;
    B200_TRY(ctx->energy_out.reserve(count * sizeof(double)));
    B200_CUDA(cudaMemcpyAsync(ctx->energy_out.p, values, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    B200_NCCL(api->AllReduce(ctx->energy_out.p, ctx->energy_out.p, count, ncclDouble, ncclSum, s->comm, ctx->stream));
    B200_CUDA(cudaMemcpyAsync(values, ctx->energy_out.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

}  // namespace b200

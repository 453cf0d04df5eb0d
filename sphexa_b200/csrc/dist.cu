/*! @file
 * Multi-GPU layer of libsphx: NCCL communicator, halo exchange (gather-pack kernel + grouped ncclSend/ncclRecv straight
 * into the contiguous halo ranges), global reductions of the step scalars, and the distributed hydro step.
 *
 * Replaces (reference paths relative to /root/reference/domain/include/cstone):
 *   Domain::exchangeHalos               domain/domain.hpp:372-377
 *   Halos::exchangeHalos                halos/halos.hpp:234-254
 *   haloExchangeGpu                     halos/exchange_halos_gpu.cuh:34-119  (MPI_Isend/Irecv, host staging without
 *                                       GPU-aware MPI) and gatherRanges, halos/gather_halos_gpu.cu:26-41
 *   MPI_Allreduce(MIN) of the time step main/src/... sph/include/sph/ts_global.hpp:97-113
 *
 * One process per GPU. The reference exchanges through MPI; here the bytes go GPU to GPU over NVLink/NVSwitch with
 * NCCL point-to-point calls enqueued on the same stream as the kernels, so a step needs no host synchronisation until
 * its scalars are read. NCCL is loaded at run time (dlopen) so that single-GPU users of libsphx need no NCCL at all.
 */
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "sphx_block.cuh"
#include "sphx_kernels.h"

namespace
{

struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};

template<class Err>
NcclApi* ncclApi(Err& err)
{
    static NcclApi        api;
    static std::once_flag once;
    std::call_once(once, [] {
        // RTLD_NOLOAD first: reuse the copy the host application (e.g. torch) already loaded
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h)
        {
            bool ok = true;
#define SPHX_SYM(name)                                                                                                 \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name));                                           \
    ok       = ok && api.name
            SPHX_SYM(GetUniqueId);
            SPHX_SYM(CommInitRank);
            SPHX_SYM(CommDestroy);
            SPHX_SYM(Send);
            SPHX_SYM(Recv);
            SPHX_SYM(AllReduce);
            SPHX_SYM(GroupStart);
            SPHX_SYM(GroupEnd);
            SPHX_SYM(GetErrorString);
#undef SPHX_SYM
            if (ok) api.handle = h;
        }
    });
    if (!api.handle)
    {
        err = "libnccl.so.2 not found or incomplete";
        return nullptr;
    }
    return &api;
}

//! error text shown by sphx_last_error(); assignment forwards to the C-ABI layer (api.cu)
struct DistError
{
    DistError& operator=(const std::string& m)
    {
        sphx::setLastError(m);
        return *this;
    }
    DistError& operator=(const char* m)
    {
        sphx::setLastError(m);
        return *this;
    }
} g_distError;

} // namespace

struct SphxComm
{
    ncclComm_t comm   = nullptr;
    int        rank   = 0;
    int        nranks = 1;
    double*    scratch = nullptr; // device, 16 doubles
    NcclApi*   api    = nullptr;
};

namespace sphx
{

//! out[a][k] = arrays[a][idx[k]] for 4- and 8-byte elements; one launch packs every array of an exchange
struct PackArgs
{
    const void* src[8];
    size_t      dstOffset[8]; // bytes into the send buffer
    int         elemBytes[8];
    int         count;
};

__global__ void packHalosKernel(PackArgs p, const unsigned* __restrict__ idx, unsigned numIdx, char* __restrict__ buf)
{
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= numIdx) return;
    unsigned j = idx[k];
    for (int a = 0; a < p.count; ++a)
    {
        if (p.elemBytes[a] == 4)
            reinterpret_cast<unsigned*>(buf + p.dstOffset[a])[k] = static_cast<const unsigned*>(p.src[a])[j];
        else
            reinterpret_cast<unsigned long long*>(buf + p.dstOffset[a])[k] =
                static_cast<const unsigned long long*>(p.src[a])[j];
    }
}

} // namespace sphx

#define SPHX_NCCL(call)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        ncclResult_t r_ = (call);                                                                                      \
        if (r_ != ncclSuccess)                                                                                         \
        {                                                                                                              \
            g_distError = std::string(#call) + ": " + c->api->GetErrorString(r_);                                      \
            return SPHX_ERR_NCCL;                                                                                      \
        }                                                                                                              \
    } while (0)

extern "C"
{

int sphx_comm_unique_id(char* id128)
{
    static_assert(sizeof(ncclUniqueId) == SPHX_UNIQUE_ID_BYTES, "ncclUniqueId size");
    NcclApi* api = ncclApi(g_distError);
    if (!api) return SPHX_ERR_NCCL;
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess)
    {
        g_distError = api->GetErrorString(r);
        return SPHX_ERR_NCCL;
    }
    memcpy(id128, &id, sizeof(id));
    return SPHX_OK;
}

int sphx_comm_init(SphxComm** out, int rank, int nranks, const char* id128)
{
    if (!out || !id128 || rank < 0 || rank >= nranks) return SPHX_ERR_INVALID;
    if (int rc = sphx_device_check()) return rc;
    NcclApi* api = ncclApi(g_distError);
    if (!api) return SPHX_ERR_NCCL;
    auto* c   = new SphxComm;
    c->api    = api;
    c->rank   = rank;
    c->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = api->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess)
    {
        g_distError = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
        delete c;
        return SPHX_ERR_NCCL;
    }
    if (cudaMalloc(&c->scratch, 16 * sizeof(double)) != cudaSuccess)
    {
        g_distError = "cudaMalloc of the reduction scratch failed";
        api->CommDestroy(c->comm);
        delete c;
        return SPHX_ERR_CUDA;
    }
    *out = c;
    return SPHX_OK;
}

int sphx_comm_rank(const SphxComm* c, int* rank, int* nranks)
{
    if (!c) return SPHX_ERR_INVALID;
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return SPHX_OK;
}

int sphx_comm_free(SphxComm* c)
{
    if (!c) return SPHX_OK;
    if (c->scratch) cudaFree(c->scratch);
    if (c->comm) c->api->CommDestroy(c->comm);
    delete c;
    return SPHX_OK;
}

int sphx_halo_exchange(SphxComm* c, const SphxHaloPlan* plan, int count, void* const* arrays, const int* elemBytes,
                       void* stream)
{
    if (!c || !plan || count < 0 || count > 8 || (count && (!arrays || !elemBytes))) return SPHX_ERR_INVALID;
    if (count == 0 || plan->numPeers == 0) return SPHX_OK;
    auto           s        = static_cast<cudaStream_t>(stream);
    const unsigned totalSend = plan->sendOffsets[plan->numPeers];

    sphx::PackArgs p;
    p.count      = count;
    size_t bytes = 0;
    for (int a = 0; a < count; ++a)
    {
        if (elemBytes[a] != 4 && elemBytes[a] != 8)
        {
            g_distError = "halo exchange supports 4- and 8-byte elements";
            return SPHX_ERR_INVALID;
        }
        p.src[a]       = arrays[a];
        p.elemBytes[a] = elemBytes[a];
        p.dstOffset[a] = bytes;
        bytes += sphx::alignUp(size_t(totalSend) * elemBytes[a], 16);
    }
    if (bytes > plan->sendBufferBytes)
    {
        g_distError = "halo send buffer too small: need " + std::to_string(bytes) + " bytes";
        return SPHX_ERR_WORKSPACE;
    }
    char* buf = static_cast<char*>(plan->sendBuffer);
    if (totalSend)
    {
        sphx::packHalosKernel<<<(totalSend + 255) / 256, 256, 0, s>>>(p, plan->sendIdx, totalSend, buf);
        if (cudaGetLastError() != cudaSuccess)
        {
            g_distError = "packHalosKernel launch failed";
            return SPHX_ERR_CUDA;
        }
    }
    SPHX_NCCL(c->api->GroupStart());
    for (int a = 0; a < count; ++a)
    {
        const size_t eb = size_t(elemBytes[a]);
        for (int q = 0; q < plan->numPeers; ++q)
        {
            const unsigned sb = plan->sendOffsets[q], se = plan->sendOffsets[q + 1];
            if (se > sb)
                SPHX_NCCL(c->api->Send(buf + p.dstOffset[a] + size_t(sb) * eb, size_t(se - sb) * eb, ncclInt8,
                                       plan->peers[q], c->comm, s));
            if (plan->recvCount[q])
                SPHX_NCCL(c->api->Recv(static_cast<char*>(arrays[a]) + size_t(plan->recvBegin[q]) * eb,
                                       size_t(plan->recvCount[q]) * eb, ncclInt8, plan->peers[q], c->comm, s));
        }
    }
    SPHX_NCCL(c->api->GroupEnd());
    return SPHX_OK;
}

int sphx_allreduce_f64(SphxComm* c, double* values_host, int n, int op, void* stream)
{
    if (!c || !values_host || n < 1 || n > 16 || op < 0 || op > 2) return SPHX_ERR_INVALID;
    auto s = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(c->scratch, values_host, n * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess)
        return SPHX_ERR_CUDA;
    const ncclRedOp_t ops[3] = {ncclMin, ncclMax, ncclSum};
    SPHX_NCCL(c->api->AllReduce(c->scratch, c->scratch, size_t(n), ncclDouble, ops[op], c->comm, s));
    if (cudaMemcpyAsync(values_host, c->scratch, n * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess)
        return SPHX_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) return SPHX_ERR_CUDA;
    return SPHX_OK;
}

int sphx_allreduce_device(SphxComm* c, void* data_dev, size_t n, int dtype, int op, void* stream)
{
    if (!c || !data_dev || op < 0 || op > 2 || dtype < 0 || dtype > 3) return SPHX_ERR_INVALID;
    const ncclRedOp_t    ops[3]   = {ncclMin, ncclMax, ncclSum};
    const ncclDataType_t types[4] = {ncclUint32, ncclUint64, ncclFloat32, ncclFloat64};
    if (n == 0) return SPHX_OK;
    SPHX_NCCL(c->api->AllReduce(data_dev, data_dev, n, types[dtype], ops[op], c->comm,
                                static_cast<cudaStream_t>(stream)));
    return SPHX_OK;
}

int sphx_exchange_slices(SphxComm* c, const size_t* sendOffsets, const size_t* recvOffsets, int count,
                         const void* const* src, void* const* dst, const int* elemBytes, void* stream)
{
    if (!c || !sendOffsets || !recvOffsets || count < 0 || (count && (!src || !dst || !elemBytes)))
        return SPHX_ERR_INVALID;
    auto s = static_cast<cudaStream_t>(stream);
    const int me = c->rank;
    for (int k = 0; k < count; ++k)
    {
        const size_t eb = size_t(elemBytes[k]);
        if (!src[k] || !dst[k] || eb == 0) return SPHX_ERR_INVALID;
        const char* sp = static_cast<const char*>(src[k]);
        char*       dp = static_cast<char*>(dst[k]);
        // the slice that stays
        size_t own = sendOffsets[me + 1] - sendOffsets[me];
        if (own != recvOffsets[me + 1] - recvOffsets[me])
        {
            g_distError = "sphx_exchange_slices: own slice differs between send and receive layout";
            return SPHX_ERR_INVALID;
        }
        if (own && cudaMemcpyAsync(dp + recvOffsets[me] * eb, sp + sendOffsets[me] * eb, own * eb,
                                   cudaMemcpyDeviceToDevice, s) != cudaSuccess)
            return SPHX_ERR_CUDA;
        SPHX_NCCL(c->api->GroupStart());
        for (int r = 0; r < c->nranks; ++r)
        {
            if (r == me) continue;
            size_t ns = sendOffsets[r + 1] - sendOffsets[r], nr = recvOffsets[r + 1] - recvOffsets[r];
            if (ns) SPHX_NCCL(c->api->Send(sp + sendOffsets[r] * eb, ns * eb, ncclChar, r, c->comm, s));
            if (nr) SPHX_NCCL(c->api->Recv(dp + recvOffsets[r] * eb, nr * eb, ncclChar, r, c->comm, s));
        }
        SPHX_NCCL(c->api->GroupEnd());
    }
    return SPHX_OK;
}

namespace
{
struct DistCtx
{
    SphxComm*           comm;
    const SphxHaloPlan* plan;
    void*               stream;
};
int distHaloCallback(void* user, int count, void* const* arrays, const int* elemBytes)
{
    auto* ctx = static_cast<DistCtx*>(user);
    return sphx_halo_exchange(ctx->comm, ctx->plan, count, arrays, elemBytes, ctx->stream);
}
} // namespace

int sphx_reduce_step_result(SphxComm* c, int localStatus, SphxStepResult* r, void* stream)
{
    if (!c || !r) return SPHX_ERR_INVALID;
    const int rc = localStatus;
    // ONE all-reduce for everything: MIN of {dtCourant, dtRho, -maxNc, -failed}, SUM of the neighbour count
    // (as two 2^26-limbs, exact in double). Runs on every rank, also when this rank's step failed.
    const unsigned long tn = rc ? 0ul : r->totalNeighbors;
    double v[6] = {rc ? 0.0 : r->minDtCourant, rc ? 0.0 : r->minDtRho, rc ? 0.0 : -double(r->maxNc), rc ? -1.0 : 0.0,
                   double(tn & ((1ul << 26) - 1)), double(tn >> 26)};
    auto   s    = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(c->scratch, v, sizeof(v), cudaMemcpyHostToDevice, s) != cudaSuccess) return SPHX_ERR_CUDA;
    SPHX_NCCL(c->api->GroupStart());
    SPHX_NCCL(c->api->AllReduce(c->scratch, c->scratch, 4, ncclDouble, ncclMin, c->comm, s));
    SPHX_NCCL(c->api->AllReduce(c->scratch + 4, c->scratch + 4, 2, ncclDouble, ncclSum, c->comm, s));
    SPHX_NCCL(c->api->GroupEnd());
    if (cudaMemcpyAsync(v, c->scratch, sizeof(v), cudaMemcpyDeviceToHost, s) != cudaSuccess) return SPHX_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) return SPHX_ERR_CUDA;
    if (rc) return rc;
    if (v[3] != 0.0)
    {
        g_distError = "the hydro step failed on another rank";
        return SPHX_ERR_NCCL;
    }
    r->minDtCourant   = v[0];
    r->minDtRho       = v[1];
    r->maxNc          = unsigned(-v[2]);
    r->totalNeighbors = (unsigned long)(v[4]) + ((unsigned long)(v[5]) << 26);
    return SPHX_OK;
}

int sphx_hydro_step_dist(const SphxStepArgs* a, SphxComm* c, const SphxHaloPlan* plan, SphxStepResult* r)
{
    if (!a || !c || !plan) return SPHX_ERR_INVALID;
    DistCtx        ctx{c, plan, a->stream};
    SphxStepResult local{};
    // sphx_hydro_step rejects bad arguments before its first exchange and, after a later local failure, still runs its
    // remaining exchanges; the reduction below then runs on every rank and carries the failure to all of them
    int rc = sphx_hydro_step(a, distHaloCallback, &ctx, &local);
    int rc2 = sphx_reduce_step_result(c, rc, &local, a->stream);
    if (rc2) return rc2;
    if (r) *r = local;
    return SPHX_OK;
}

} // extern "C"

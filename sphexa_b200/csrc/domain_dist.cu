/*! @file
 * Multi-rank Domain::sync behind one C entry point (sphx_domain_sync_dist) and the halo exchange that belongs to it.
 *
 * Replaces, for the hot path's purposes (reference paths relative to /root/reference/domain/include/cstone):
 *   Domain::sync                 domain/domain.hpp:181-234   (global assignment, particle migration, halo discovery,
 *                                                             layout [halos | assigned | halos], focus tree, field reorder)
 *   GlobalAssignment::assign     domain/assignment.hpp:67-120 (box, keys, SFC order, balanced key ranges)
 *   Halos::discover / exchange   halos/halos.hpp:131-254
 *   Domain::exchangeHalos        domain/domain.hpp:372-377
 *
 * Re-designed around ONE global object, the particle count per Hilbert cell of a level whose edge is >= 2 max(h) (or
 * two levels finer with a per-cell reach): keys + local radix sort -> cell histogram -> ncclAllReduce
 * -> decomposition plan from the global histogram alone, on the device, identical on every rank (no request messages)
 * -> particle migration as one slice per peer and field -> merge of the arrivals -> halo exchange of the listed fields
 * -> octree over the local particles. The reference negotiates the same through a focus octree with peer-to-peer
 * messages. Everything on the data path is a kernel or an NCCL call enqueued on the caller's stream; the host sees three
 * small PODs per sync (box + h statistics, plan summary + migration counts, node counts of the tree levels).
 *
 * Memory: the particle arrays belong to the caller (ParticlesData in the reference) and come with a spare of the same
 * capacity each, exactly like the reference's sync swaps its field vectors with scratch vectors; scratch that only the
 * sync needs (keys, permutation, histograms, plan, send lists, tree) belongs to the SphxDomain object and grows on demand.
 */
#include <cub/device/device_reduce.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "sphx_block.cuh"
#include "sphx_kernels.h"

namespace
{

int domFail(int code, const std::string& msg)
{
    sphx::setLastError(msg);
    return code;
}

#define DOM_CUDA(call)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) return domFail(SPHX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)
#define DOM_OK(call)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        int rc_ = (call);                                                                                              \
        if (rc_ != SPHX_OK) return rc_;                                                                                \
    } while (0)

//! device buffer owned by the domain, grown on demand (contents are not preserved)
struct DevBuf
{
    void*  p     = nullptr;
    size_t bytes = 0;
    int    ensure(size_t need, double growth = 1.05)
    {
        if (need <= bytes) return SPHX_OK;
        if (p) cudaFree(p);
        p     = nullptr;
        bytes = 0;
        size_t want = size_t(double(need) * growth) + 256;
        if (cudaMalloc(&p, want) != cudaSuccess)
        {
            cudaGetLastError();
            return domFail(SPHX_ERR_CUDA, "sphx domain: cudaMalloc of " + std::to_string(want) + " bytes failed");
        }
        bytes = want;
        return SPHX_OK;
    }
    template<class T>
    T* as() const { return static_cast<T*>(p); }
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
};

//! per-rank statistics that fix the box and the cell level: MIN of {x, -x, y, -y, z, -z, -h}, SUM of {h, count}
struct StatsDev
{
    double mins[8];
    double sums[2];
};

__global__ void statsKernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                            const float* __restrict__ h, unsigned n, double* __restrict__ partial /* [grid][10] */)
{
    double v[10] = {1e300, 1e300, 1e300, 1e300, 1e300, 1e300, 1e300, 1e300, 0.0, 0.0};
    // fixed assignment of particles to threads and a fixed reduction tree: deterministic
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const double xi = x[i], yi = y[i], zi = z[i], hi = double(h[i]);
        v[0] = fmin(v[0], xi), v[1] = fmin(v[1], -xi), v[2] = fmin(v[2], yi), v[3] = fmin(v[3], -yi);
        v[4] = fmin(v[4], zi), v[5] = fmin(v[5], -zi), v[6] = fmin(v[6], -hi);
        v[8] += hi, v[9] += 1.0;
    }
    __shared__ double red[10][256];
    for (int k = 0; k < 10; ++k)
        red[k][threadIdx.x] = v[k];
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1)
    {
        if (int(threadIdx.x) < s)
        {
            for (int k = 0; k < 8; ++k)
                red[k][threadIdx.x] = fmin(red[k][threadIdx.x], red[k][threadIdx.x + s]);
            red[8][threadIdx.x] += red[8][threadIdx.x + s];
            red[9][threadIdx.x] += red[9][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x < 10) partial[blockIdx.x * 10 + threadIdx.x] = red[threadIdx.x][0];
}

__global__ void statsFinalKernel(const double* __restrict__ partial, int numBlocks, StatsDev* out)
{
    const int k = threadIdx.x;
    if (k >= 10) return;
    double v = k < 8 ? 1e300 : 0.0;
    for (int b = 0; b < numBlocks; ++b)
        v = k < 8 ? fmin(v, partial[b * 10 + k]) : v + partial[b * 10 + k];
    if (k < 8) { out->mins[k] = v; }
    else { out->sums[k - 8] = v; }
}

//! largest h per cell (sorted particle p is input particle order[p]); h > 0, so the bit patterns order like the values
__global__ void cellMaxHKernel(const uint64_t* __restrict__ sortedKeys, const unsigned* __restrict__ order,
                               const float* __restrict__ h, unsigned n, int shift, unsigned* __restrict__ hcellBits)
{
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    atomicMax(&hcellBits[unsigned(sortedKeys[p] >> shift)], __float_as_uint(h[order[p]]));
}

//! reach of a cell in rings of cells: ceil(2 h_cell 1.0001 / edge), clamped to [1, maxRing]
__global__ void ringsKernel(const unsigned* __restrict__ hcellBits, unsigned ncell, double factor, int maxRing,
                            unsigned char* __restrict__ rings)
{
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const double r = ceil(double(__uint_as_float(hcellBits[c])) * factor);
    rings[c]       = (unsigned char)(min(max(r, 1.0), double(maxRing)));
}

__global__ void matrixRowKernel(unsigned long long* __restrict__ m, int me, int R, const unsigned long long* sendOff,
                                unsigned long long tooSmall)
{
    const int r = threadIdx.x;
    if (r < R) m[me * R + r] = sendOff[r + 1] - sendOff[r];
    if (r == 0) m[size_t(R) * R] = tooSmall;
}

} // namespace

struct SphxDomain
{
    SphxComm* comm   = nullptr;
    int       rank   = 0;
    int       nranks = 1;
    SphxBox   box{};
    unsigned  bucket   = 64;
    bool      syncedOnce = false;

    DevBuf keys, order, syncScratch, hist, localHist, hcell, rings, planScratch, sendIdx, sendBuf, stats, partial, matrix,
        keys2, order2, localKeys;
    // tree arrays (SphxTreeView points at them)
    DevBuf prefixes, childOffsets, internalToLeaf, levelRange, leaves, layout, centers, sizes;
    int    maxNodes = 0;

    // halo plan of the last sync (host arrays referenced by `plan`)
    std::vector<int>      peers;
    std::vector<unsigned> sendOffsets, recvBegin, recvCount;
    SphxHaloPlan          plan{};
    SphxCellPlanSummary   summary{};
    int                   level = 0;
    size_t                numLocal = 0;
};

namespace
{

int ensureTree(SphxDomain* d, int maxNodes)
{
    if (maxNodes <= d->maxNodes) return SPHX_OK;
    DOM_OK(d->prefixes.ensure(size_t(maxNodes) * 8));
    DOM_OK(d->childOffsets.ensure(size_t(maxNodes) * 4));
    DOM_OK(d->internalToLeaf.ensure(size_t(maxNodes) * 4));
    DOM_OK(d->levelRange.ensure(23 * 4));
    DOM_OK(d->leaves.ensure(size_t(maxNodes + 1) * 8));
    DOM_OK(d->layout.ensure(size_t(maxNodes + 1) * 4));
    DOM_OK(d->centers.ensure(size_t(maxNodes) * 24));
    DOM_OK(d->sizes.ensure(size_t(maxNodes) * 24));
    d->maxNodes = maxNodes;
    return SPHX_OK;
}

//! keys (+ SFC permutation, + tree) of n particles through sphx_domain_sync
int sfcSort(SphxDomain* d, const double* x, const double* y, const double* z, size_t n, DevBuf& keys, DevBuf* order,
            bool presorted, bool tree, cudaStream_t s, int* numNodes, int* numLeaves)
{
    DOM_OK(keys.ensure(std::max<size_t>(n, 1) * 8));
    if (order) DOM_OK(order->ensure(std::max<size_t>(n, 1) * 4));
    for (;;)
    {
        const int mn = tree ? d->maxNodes : 0;
        DOM_OK(d->syncScratch.ensure(sphx_domain_sync_bytes(n, mn)));
        SphxSyncArgs a{};
        a.n = n, a.box = d->box, a.bucketSize = d->bucket;
        a.x = x, a.y = y, a.z = z;
        a.keys = keys.as<uint64_t>(), a.order = order ? order->as<unsigned>() : nullptr, a.maxNodes = mn;
        if (tree)
        {
            a.prefixes = d->prefixes.as<uint64_t>(), a.childOffsets = d->childOffsets.as<int>();
            a.internalToLeaf = d->internalToLeaf.as<int>(), a.levelRange = d->levelRange.as<int>();
            a.leaves = d->leaves.as<uint64_t>(), a.layout = d->layout.as<unsigned>();
            a.centers = d->centers.as<double>(), a.sizes = d->sizes.as<double>();
        }
        a.scratch = d->syncScratch.p, a.scratchBytes = d->syncScratch.bytes, a.stream = s;
        a.flags = (presorted ? SPHX_SYNC_PRESORTED : 0) | (tree ? 0 : SPHX_SYNC_NO_TREE);
        int rc  = sphx_domain_sync(&a, nullptr, numNodes, numLeaves);
        if (rc == SPHX_ERR_WORKSPACE && tree && size_t(d->maxNodes) < 8 * n + 64)
        {
            DOM_OK(ensureTree(d, 2 * d->maxNodes)); // the tree needs more nodes than the buffers hold: grow, redo
            continue;
        }
        return rc;
    }
}

} // namespace

extern "C"
{

int sphx_domain_create(SphxDomain** out, SphxComm* comm, const SphxBox* box, unsigned bucketSize)
{
    if (!out || !box || bucketSize == 0) return domFail(SPHX_ERR_INVALID, "sphx_domain_create: bad argument");
    if (int rc = sphx_device_check()) return rc;
    auto* d   = new SphxDomain;
    d->comm   = comm;
    d->box    = *box;
    d->bucket = bucketSize;
    if (comm) sphx_comm_rank(comm, &d->rank, &d->nranks);
    if (d->nranks > SPHX_MAX_RANKS)
    {
        delete d;
        return domFail(SPHX_ERR_INVALID, "sphx_domain_create: more than SPHX_MAX_RANKS ranks");
    }
    *out = d;
    return SPHX_OK;
}

void sphx_domain_destroy(SphxDomain* d) { delete d; }

const SphxHaloPlan* sphx_domain_halo_plan(const SphxDomain* d) { return d ? &d->plan : nullptr; }

int sphx_domain_copy_local_keys(const SphxDomain* d, uint64_t* dst, void* stream)
{
    if (!d || (!dst && d->numLocal)) return domFail(SPHX_ERR_INVALID, "sphx_domain_copy_local_keys: null argument");
    if (d->numLocal) // (a rank without particles: nothing to copy, dst may be null)
        DOM_CUDA(cudaMemcpyAsync(dst, d->localKeys.p, d->numLocal * 8, cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
    return SPHX_OK;
}

int sphx_domain_exchange_halos(SphxDomain* d, int count, void* const* arrays, const int* elemBytes, void* stream)
{
    if (!d) return domFail(SPHX_ERR_INVALID, "sphx_domain_exchange_halos: null domain");
    if (d->nranks == 1 || !d->comm) return SPHX_OK;
    return sphx_halo_exchange(d->comm, &d->plan, count, arrays, elemBytes, stream);
}

int sphx_domain_sync_dist(SphxDomain* d, const SphxDomainSyncArgs* a, SphxDomainResult* res)
{
    if (!d || !a || !res) return domFail(SPHX_ERR_INVALID, "sphx_domain_sync_dist: null argument");
    if (int rc = sphx_device_check()) return rc;
    if (a->count < 4 || a->count > 16 || !a->arrays || !a->spare || !a->elemBytes || a->inLast < a->inFirst ||
        a->inLast > a->capacity || a->numHaloFields < 0 || a->numHaloFields > 8 ||
        (a->numHaloFields && !a->haloFields))
        return domFail(SPHX_ERR_INVALID, "sphx_domain_sync_dist: bad argument");
    for (int k = 0; k < a->count; ++k)
    {
        const int eb = a->elemBytes[k];
        if (!a->arrays[k] || !a->spare[k] || a->arrays[k] == a->spare[k] || (eb != 4 && eb != 8) ||
            (k < 3 && eb != 8) || (k == 3 && eb != 4))
            return domFail(SPHX_ERR_INVALID, "sphx_domain_sync_dist: arrays must be x, y, z (8 bytes), h (4 bytes), then "
                                             "4- or 8-byte fields, each with a distinct spare");
    }
    for (int k = 0; k < a->numHaloFields; ++k)
        if (a->haloFields[k] < 0 || a->haloFields[k] >= a->count)
            return domFail(SPHX_ERR_INVALID, "sphx_domain_sync_dist: halo field index out of range");

    const int    R = d->nranks, me = d->rank;
    auto         s    = static_cast<cudaStream_t>(a->stream);
    const size_t nOld = a->inLast - a->inFirst;
    if (nOld >= (size_t(1) << 31)) return domFail(SPHX_ERR_INVALID, "sphx_domain_sync_dist: too many particles");
    std::vector<const void*> cur(a->count);
    for (int k = 0; k < a->count; ++k)
        cur[k] = static_cast<const char*>(a->arrays[k]) + a->inFirst * size_t(a->elemBytes[k]);
    const double* x = static_cast<const double*>(cur[0]);
    const double* y = static_cast<const double*>(cur[1]);
    const double* z = static_cast<const double*>(cur[2]);
    const float*  h = static_cast<const float*>(cur[3]);

    // ---- 1. global box (open dimensions follow the particles: makeGlobalBox, box_mpi.hpp:66-109; limited shrinking after
    //         the first sync, assignment.hpp:80-82), max and mean h: one fused all-reduce, one small D2H copy -------------
    DOM_OK(d->stats.ensure(sizeof(StatsDev)));
    const int statBlocks = 296;
    DOM_OK(d->partial.ensure(size_t(statBlocks) * 10 * sizeof(double)));
    statsKernel<<<statBlocks, 256, 0, s>>>(x, y, z, h, unsigned(nOld), d->partial.as<double>());
    statsFinalKernel<<<1, 32, 0, s>>>(d->partial.as<double>(), statBlocks, d->stats.as<StatsDev>());
    if (R > 1)
    {
        DOM_OK(sphx_allreduce_device(d->comm, d->stats.as<StatsDev>()->mins, 8, 3, 0, s));
        DOM_OK(sphx_allreduce_device(d->comm, d->stats.as<StatsDev>()->sums, 2, 3, 2, s));
    }
    StatsDev st;
    DOM_CUDA(cudaMemcpyAsync(&st, d->stats.p, sizeof(st), cudaMemcpyDeviceToHost, s));
    DOM_CUDA(cudaStreamSynchronize(s));
    const double nGlobalD = st.sums[1];
    if (nGlobalD < 1.0) return domFail(SPHX_ERR_INVALID, "sphx_domain_sync_dist: no particles on any rank");
    for (int dim = 0; dim < 3; ++dim)
    {
        if (d->box.boundary[dim] == 1) continue;
        double lo = st.mins[2 * dim], hi = -st.mins[2 * dim + 1];
        if (d->syncedOnce)
        {
            const double pl = d->box.lim[2 * dim], ph = d->box.lim[2 * dim + 1], ext = ph - pl;
            lo = std::min(lo, pl + 0.05 * ext), hi = std::max(hi, ph - 0.05 * ext);
        }
        d->box.lim[2 * dim] = lo, d->box.lim[2 * dim + 1] = hi;
    }
    const double hMax = -st.mins[6], hMean = st.sums[0] / nGlobalD;
    // cells: the finest level whose edge is >= 2 max(h) in every dimension (so one ring of cells covers every search
    // sphere), capped at 7; with strongly varying h two levels finer, each cell with its own reach in rings
    const double ext = std::min({d->box.lim[1] - d->box.lim[0], d->box.lim[3] - d->box.lim[2], d->box.lim[5] - d->box.lim[4]});
    int          coarse = 0;
    while (coarse < 7 && ext / double(1 << (coarse + 1)) >= 2.0 * hMax * 1.0001)
        ++coarse;
    const int      level   = std::min(7, coarse + (hMax > 1.25 * hMean ? 2 : 0));
    const unsigned ncell   = 1u << (3 * level);
    const double   edge    = ext / double(1 << level);
    const int      maxRing = std::max(1, std::min(16, int(std::ceil(2.0 * hMax * 1.0001 / edge))));

    // ---- 2. local SFC order, cell histogram, largest h per cell; global versions by all-reduce ------------------------
    DOM_OK(sfcSort(d, x, y, z, nOld, d->keys, &d->order, false, false, s, nullptr, nullptr));
    DOM_OK(d->hist.ensure(size_t(ncell) * 4));
    DOM_OK(d->localHist.ensure(size_t(ncell) * 4));
    DOM_OK(d->hcell.ensure(size_t(ncell) * 4));
    DOM_OK(d->rings.ensure(ncell));
    DOM_OK(sphx_cell_histogram(d->keys.as<uint64_t>(), nOld, level, d->localHist.as<unsigned>(), s));
    DOM_CUDA(cudaMemcpyAsync(d->hist.p, d->localHist.p, size_t(ncell) * 4, cudaMemcpyDeviceToDevice, s));
    DOM_CUDA(cudaMemsetAsync(d->hcell.p, 0, size_t(ncell) * 4, s));
    if (nOld)
        cellMaxHKernel<<<unsigned((nOld + 255) / 256), 256, 0, s>>>(d->keys.as<uint64_t>(), d->order.as<unsigned>(), h,
                                                                   unsigned(nOld), 3 * (21 - level),
                                                                   d->hcell.as<unsigned>());
    if (R > 1)
    {
        DOM_OK(sphx_allreduce_device(d->comm, d->hist.p, ncell, 0, 2, s));
        DOM_OK(sphx_allreduce_device(d->comm, d->hcell.p, ncell, 0, 1, s)); // max of positive floats = max of their bits
    }
    ringsKernel<<<(ncell + 255) / 256, 256, 0, s>>>(d->hcell.as<unsigned>(), ncell, 2.0 * 1.0001 / edge, maxRing,
                                                     d->rings.as<unsigned char>());

    // ---- 3. the plan, on the device: assignment, halo cells, send lists, layout; then the R x R migration counts ---------
    DOM_OK(d->planScratch.ensure(sphx_cell_plan_device_bytes(level)));
    const size_t sendCap = 4 * nOld + 65536;
    DOM_OK(d->sendIdx.ensure(sendCap * 4));
    int per[3] = {d->box.boundary[0] == 1, d->box.boundary[1] == 1, d->box.boundary[2] == 1};
    DOM_OK(sphx_cell_plan_build_device(d->hist.as<unsigned>(), d->localHist.as<unsigned>(),
                                       d->rings.as<unsigned char>(), maxRing, level, per, me, R, d->planScratch.p,
                                       d->planScratch.bytes, d->sendIdx.as<unsigned>(), sendCap, nullptr, &d->summary, s));
    const SphxCellPlanSummary& sum = d->summary;
    const size_t first = sum.nHaloLeft, last = first + size_t(sum.nAssigned), nLocal = last + sum.nHaloRight;
    res->needCapacity = std::max(nLocal, std::max(nOld, size_t(sum.nAssigned)));
    const bool tooSmall = res->needCapacity > a->capacity;
    std::vector<size_t> sendOff(R + 1), recvOff(R + 1, 0);
    for (int r = 0; r <= R; ++r)
        sendOff[r] = size_t(sum.sendOffLocal[r]);
    bool anyTooSmall = tooSmall;
    if (R > 1)
    {
        // R x R matrix of migration counts (every rank fills its row, one all-reduce completes it) + the number of ranks
        // whose arrays are too small for their new local set: either every rank goes on or none does
        const size_t nm = size_t(R) * R + 1;
        DOM_OK(d->matrix.ensure(nm * 8 + size_t(R + 1) * 8));
        auto* m   = d->matrix.as<unsigned long long>();
        auto* off = m + nm;
        DOM_CUDA(cudaMemsetAsync(m, 0, nm * 8, s));
        DOM_CUDA(cudaMemcpyAsync(off, sum.sendOffLocal, size_t(R + 1) * 8, cudaMemcpyHostToDevice, s));
        matrixRowKernel<<<1, 64, 0, s>>>(m, me, R, off, tooSmall ? 1ull : 0ull);
        DOM_OK(sphx_allreduce_device(d->comm, m, nm, 1, 2, s));
        std::vector<unsigned long long> mh(nm);
        DOM_CUDA(cudaMemcpyAsync(mh.data(), m, nm * 8, cudaMemcpyDeviceToHost, s));
        DOM_CUDA(cudaStreamSynchronize(s));
        for (int r = 0; r < R; ++r)
            recvOff[r + 1] = recvOff[r] + size_t(mh[size_t(r) * R + me]);
        anyTooSmall = mh[nm - 1] != 0;
    }
    else { recvOff[1] = sendOff[1] - sendOff[0]; }
    if (anyTooSmall)
        return domFail(SPHX_ERR_WORKSPACE, "sphx_domain_sync_dist: arrays hold " + std::to_string(a->capacity) +
                                               " particles, the new local set of this rank needs " +
                                               std::to_string(res->needCapacity) + " (some rank ran short: all return)");
    const size_t nNew = recvOff[R];
    if (nNew != sum.nAssigned) return domFail(SPHX_ERR_INVALID, "sphx_domain_sync_dist: migration counts disagree with the plan");

    // ---- 4. migration: sort my particles into the spares, ship one slice per peer and field back into the arrays,
    //         merge the arrivals into [first, last) of the spares --------------------------------------------------------
    std::vector<void*>       sp(a->count), ar(a->count), dstv(a->count);
    std::vector<const void*> csp(a->count), car(a->count);
    for (int k = 0; k < a->count; ++k)
    {
        sp[k] = a->spare[k], ar[k] = a->arrays[k];
        csp[k] = sp[k], car[k] = ar[k];
        dstv[k] = static_cast<char*>(a->spare[k]) + first * size_t(a->elemBytes[k]);
    }
    if (R > 1)
    {
        DOM_OK(sphx_reorder_fields(d->order.as<unsigned>(), nOld, a->count, cur.data(), sp.data(), a->elemBytes, s));
        DOM_OK(sphx_exchange_slices(d->comm, sendOff.data(), recvOff.data(), a->count, csp.data(), ar.data(), a->elemBytes, s));
        DOM_OK(sfcSort(d, static_cast<const double*>(car[0]), static_cast<const double*>(car[1]),
                       static_cast<const double*>(car[2]), nNew, d->keys2, &d->order2, false, false, s, nullptr, nullptr));
        DOM_OK(sphx_reorder_fields(d->order2.as<unsigned>(), nNew, a->count, car.data(), dstv.data(), a->elemBytes, s));
    }
    else
    {
        // one rank: no halos, first == 0: the sorted particles are the local set
        DOM_OK(sphx_reorder_fields(d->order.as<unsigned>(), nOld, a->count, cur.data(), dstv.data(), a->elemBytes, s));
    }

    // ---- 5. halo plan (host arrays of the peers, device send lists) and the halo exchange of the listed fields ---------
    d->peers.clear(), d->sendOffsets.assign(1, 0u), d->recvBegin.clear(), d->recvCount.clear();
    {
        size_t pos = 0;
        for (int r = 0; r < R; ++r)
        {
            if (r == me)
            {
                pos = last; // right halos follow the assigned range
                continue;
            }
            if (sum.recvCount[r] || sum.sendCount[r])
            {
                d->peers.push_back(r);
                d->sendOffsets.push_back(d->sendOffsets.back() + sum.sendCount[r]);
                d->recvBegin.push_back(unsigned(pos));
                d->recvCount.push_back(sum.recvCount[r]);
            }
            pos += sum.recvCount[r];
        }
    }
    const size_t bufBytes = 8 * sphx::alignUp(size_t(sum.numSend) * 8, 16) + 64;
    DOM_OK(d->sendBuf.ensure(bufBytes));
    d->plan.numPeers        = int(d->peers.size());
    d->plan.peers           = d->peers.data();
    d->plan.sendOffsets     = d->sendOffsets.data();
    d->plan.sendIdx         = d->sendIdx.as<unsigned>();
    d->plan.recvBegin       = d->recvBegin.data();
    d->plan.recvCount       = d->recvCount.data();
    d->plan.sendBuffer      = d->sendBuf.p;
    d->plan.sendBufferBytes = d->sendBuf.bytes;
    if (R > 1 && a->numHaloFields)
    {
        void* harr[8];
        int   hbytes[8];
        for (int k = 0; k < a->numHaloFields; ++k)
            harr[k] = a->spare[a->haloFields[k]], hbytes[k] = a->elemBytes[a->haloFields[k]];
        DOM_OK(sphx_halo_exchange(d->comm, &d->plan, a->numHaloFields, harr, hbytes, s));
    }

    // ---- 6. octree over the local particles (already in SFC order) -------------------------------------------------------
    DOM_OK(ensureTree(d, std::max(4096, int(std::min<size_t>(nLocal / 4 + 64, size_t(1) << 30)))));
    int numNodes = 0, numLeaves = 0;
    DOM_OK(sfcSort(d, static_cast<const double*>(a->spare[0]), static_cast<const double*>(a->spare[1]),
                   static_cast<const double*>(a->spare[2]), nLocal, d->localKeys, nullptr, true, true, s, &numNodes,
                   &numLeaves));

    d->syncedOnce = true;
    d->level      = level;
    d->numLocal   = nLocal;
    res->first = first, res->last = last, res->numLocal = nLocal, res->numGlobal = size_t(sum.nGlobal);
    res->box   = d->box;
    res->level = level;
    res->swapped = 1;
    res->localKeys = d->localKeys.as<uint64_t>();
    SphxTreeView& t = res->tree;
    t.numLeafNodes = numLeaves, t.numNodes = numNodes;
    t.prefixes = d->prefixes.as<uint64_t>(), t.childOffsets = d->childOffsets.as<int>();
    t.internalToLeaf = d->internalToLeaf.as<int>(), t.levelRange = d->levelRange.as<int>();
    t.leaves = d->leaves.as<uint64_t>(), t.layout = d->layout.as<unsigned>();
    t.centers = d->centers.as<double>(), t.sizes = d->sizes.as<double>();
    t.searchExtFactor = 1.0f;
    return SPHX_OK;
}

} // extern "C"

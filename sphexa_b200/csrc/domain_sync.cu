/*! @file
 * Domain::sync on the device for one rank: the caller of the hot path that produces what the
 * neighbour search consumes: SFC-sorted particles and the OctreeNsView arrays.
 *
 * Replaces (reference paths relative to /root/reference):
 *   computeSfcKeysGpu / sfc3D / iHilbert    domain/include/cstone/sfc/sfc_gpu.cu, sfc/sfc.hpp:141-178, hilbert.hpp:43-93
 *   GpuSfcSorter::setMapFromCodes           domain/include/cstone/primitives/primitives_gpu.cu (sort_by_key)
 *   gatherGpu (field reorder)               domain/include/cstone/primitives/gather.cu
 *   computeOctreeGpu (converged)            domain/include/cstone/tree/csarray_gpu.cu (csarray.hpp:181-430)
 *   buildOctreeGpu                          domain/include/cstone/tree/octree_gpu.cu (octree.hpp:78-197)
 *   computeGeoCentersGpu                    domain/include/cstone/focus/source_center_gpu.cu, sfc/box.hpp:318-334
 *
 * Not a port: the reference converges its leaf array by repeated count/rebalance sweeps (one sweep per time step) and
 * then derives the linked tree by sorting Warren-Salmon prefixes. Here the tree is built top-down in one go, one level
 * per pass: a node [start, start + 8^(21-l)) is opened iff it holds more than bucketSize particles, its eight child
 * particle ranges come from binary searches in the sorted key array, and an exclusive scan over the open flags places
 * the children, which directly yields nodes sorted by (level, key) with the 8 siblings consecutive. The result is
 * the fixed point of the reference's rebalance (tests compare it with csrc/host_domain.cpp, which is pinned to
 * reference dumps). The Hilbert curve is the table-driven state machine of host_domain.cpp, tables in constant memory.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "sphx_kernels.h"

namespace sphx
{

constexpr int      kMaxLevel = 21;
constexpr unsigned kMaxCoord = 1u << kMaxLevel;
constexpr int      kMaxHilbertStates = 32;

__constant__ uint8_t c_hDigit[kMaxHilbertStates * 8];
__constant__ uint8_t c_hNext[kMaxHilbertStates * 8];
__constant__ uint8_t c_hOctant[kMaxHilbertStates * 8];

namespace
{

struct KeyBox
{
    double xmin, ymin, zmin, mx, my, mz; // m = 2^21 / L
    double lx, ly, lz;
};

//! floor(x * m) - xmin * m truncated to int, clipped (sfc/sfc.hpp:141-159); every operation rounded separately
__device__ __forceinline__ unsigned gridCoord(double v, double vmin, double m)
{
    int i = int(__dsub_rn(floor(__dmul_rn(v, m)), __dmul_rn(vmin, m)));
    return unsigned(min(i, int(kMaxCoord - 1)));
}

__global__ void hilbertKeysKernel(const double* __restrict__ x, const double* __restrict__ y,
                                  const double* __restrict__ z, unsigned n, KeyBox b, uint64_t* __restrict__ keys,
                                  unsigned* __restrict__ iota)
{
    __shared__ uint8_t sDigit[kMaxHilbertStates * 8], sNext[kMaxHilbertStates * 8];
    for (int k = threadIdx.x; k < kMaxHilbertStates * 8; k += blockDim.x)
    {
        sDigit[k] = c_hDigit[k];
        sNext[k]  = c_hNext[k];
    }
    __syncthreads();
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned ix = gridCoord(x[i], b.xmin, b.mx), iy = gridCoord(y[i], b.ymin, b.my), iz = gridCoord(z[i], b.zmin, b.mz);
    uint64_t key   = 0;
    unsigned state = 0;
#pragma unroll
    for (int level = kMaxLevel - 1; level >= 0; --level)
    {
        unsigned o = ((ix >> level) & 1u) << 2 | ((iy >> level) & 1u) << 1 | ((iz >> level) & 1u);
        key        = (key << 3) | sDigit[state * 8 + o];
        state      = sNext[state * 8 + o];
    }
    keys[i] = key;
    iota[i] = i;
}

//! order-preserving map double -> u64 so that min / max become integer atomics
__host__ __device__ inline unsigned long long orderedBits(double v)
{
    unsigned long long b;
    memcpy(&b, &v, 8);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
inline double fromOrderedBits(unsigned long long b)
{
    b = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

//! makeGlobalBox (sfc/box_mpi.hpp:66-109): coordinate extrema; out[2*d] = min, out[2*d+1] = max in ordered bits
__global__ void extremaKernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                              unsigned n, unsigned long long* __restrict__ out)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        double v[3] = {x[i], y[i], z[i]};
#pragma unroll
        for (int d = 0; d < 3; ++d)
            lo[d] = fmin(lo[d], v[d]), hi[d] = fmax(hi[d], v[d]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        for (int o = 16; o > 0; o >>= 1)
        {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            atomicMin(out + 2 * d, orderedBits(lo[d]));
            atomicMax(out + 2 * d + 1, orderedBits(hi[d]));
        }
    }
}

//! first index in [lo, hi) whose key is >= v
__device__ __forceinline__ unsigned lowerBound(const uint64_t* __restrict__ keys, unsigned lo, unsigned hi, uint64_t v)
{
    while (lo < hi)
    {
        unsigned mid = lo + (hi - lo) / 2;
        if (keys[mid] < v) { lo = mid + 1; }
        else { hi = mid; }
    }
    return lo;
}

struct NodeArrays
{
    uint64_t* start;  // first key of the node
    unsigned* pBegin; // particle range
    unsigned* pEnd;
};

__global__ void openFlagsKernel(NodeArrays nd, int off, int cnt, unsigned bucketSize, int level, int* flags)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    flags[k] = (level < kMaxLevel) && (nd.pEnd[off + k] - nd.pBegin[off + k] > bucketSize);
}

//! one thread per (node, child)
__global__ void emitChildrenKernel(NodeArrays nd, const uint64_t* __restrict__ keys, int off, int cnt, int level,
                                   const int* __restrict__ flags, const int* __restrict__ pos, int nextOff,
                                   int* __restrict__ childOffsets)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int k = t >> 3, c = t & 7;
    if (k >= cnt) return;
    if (!flags[k])
    {
        if (c == 0) childOffsets[off + k] = 0;
        return;
    }
    int first = nextOff + 8 * pos[k];
    if (c == 0) childOffsets[off + k] = first;
    uint64_t childRange = uint64_t(1) << (3 * (kMaxLevel - level - 1));
    uint64_t cs         = nd.start[off + k] + uint64_t(c) * childRange;
    unsigned pb = nd.pBegin[off + k], pe = nd.pEnd[off + k];
    unsigned b = (c == 0) ? pb : lowerBound(keys, pb, pe, cs);
    unsigned e = (c == 7) ? pe : lowerBound(keys, pb, pe, cs + childRange);
    nd.start[first + c]  = cs;
    nd.pBegin[first + c] = b;
    nd.pEnd[first + c]   = e;
}

__global__ void leafFlagsKernel(const int* __restrict__ childOffsets, int numNodes, int* __restrict__ flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < numNodes) flags[i] = childOffsets[i] == 0;
}

__global__ void gatherLeafKeysKernel(const int* __restrict__ leafNodes, int numLeaves, const uint64_t* __restrict__ start,
                                     uint64_t* __restrict__ leafKeys)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < numLeaves) leafKeys[k] = start[leafNodes[k]];
}

__global__ void linkLeavesKernel(const int* __restrict__ sortedLeafNodes, int numLeaves, NodeArrays nd, unsigned n,
                                 int* __restrict__ internalToLeaf, uint64_t* __restrict__ leaves,
                                 unsigned* __restrict__ layout)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < numLeaves)
    {
        int node             = sortedLeafNodes[k];
        internalToLeaf[node] = k;
        leaves[k]            = nd.start[node];
        layout[k]            = nd.pBegin[node];
    }
    if (k == numLeaves)
    {
        leaves[k] = uint64_t(1) << (3 * kMaxLevel);
        layout[k] = n;
    }
}

//! prefixes + centerAndSize (sfc/box.hpp:318-334) from the decoded integer box of each node
__global__ void nodeGeometryKernel(NodeArrays nd, int numNodes, const int* __restrict__ levelRange, KeyBox b,
                                   int* __restrict__ internalToLeaf, const int* __restrict__ childOffsets,
                                   uint64_t* __restrict__ prefixes, double* __restrict__ centers,
                                   double* __restrict__ sizes)
{
    __shared__ int sRange[kMaxLevel + 2];
    if (threadIdx.x < kMaxLevel + 2) sRange[threadIdx.x] = levelRange[threadIdx.x];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) return;
    int l = 0;
    while (l < kMaxLevel && i >= sRange[l + 1])
        ++l;
    uint64_t start = nd.start[i];
    prefixes[i]    = (uint64_t(1) << (3 * l)) | (start >> (3 * (kMaxLevel - l)));
    if (childOffsets[i] != 0) internalToLeaf[i] = -1;

    unsigned ix = 0, iy = 0, iz = 0, state = 0;
    for (int level = kMaxLevel - 1; level >= 0; --level)
    {
        unsigned d = unsigned(start >> (3 * level)) & 7u;
        unsigned o = c_hOctant[state * 8 + d];
        ix |= ((o >> 2) & 1u) << level;
        iy |= ((o >> 1) & 1u) << level;
        iz |= (o & 1u) << level;
        state = c_hNext[state * 8 + o];
    }
    unsigned cube = kMaxCoord >> l;
    unsigned mask = ~(cube - 1);
    ix &= mask, iy &= mask, iz &= mask;
    const double uL = 1.0 / kMaxCoord;
    double       hx = __dmul_rn(__dmul_rn(0.5, uL), b.lx), hy = __dmul_rn(__dmul_rn(0.5, uL), b.ly),
           hz = __dmul_rn(__dmul_rn(0.5, uL), b.lz);
    int ixmax = int(ix + cube), iymax = int(iy + cube), izmax = int(iz + cube);
    centers[3 * i + 0] = __dadd_rn(b.xmin, __dmul_rn(double(ixmax + int(ix)), hx));
    centers[3 * i + 1] = __dadd_rn(b.ymin, __dmul_rn(double(iymax + int(iy)), hy));
    centers[3 * i + 2] = __dadd_rn(b.zmin, __dmul_rn(double(izmax + int(iz)), hz));
    sizes[3 * i + 0]   = __dmul_rn(double(ixmax - int(ix)), hx);
    sizes[3 * i + 1]   = __dmul_rn(double(iymax - int(iy)), hy);
    sizes[3 * i + 2]   = __dmul_rn(double(izmax - int(iz)), hz);
}

//! particles per Hilbert cell of the given level from the sorted keys: two binary searches per cell
__global__ void cellHistogramKernel(const uint64_t* __restrict__ keys, unsigned n, int shift, unsigned numCells,
                                    unsigned* __restrict__ counts)
{
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= numCells) return;
    unsigned b = lowerBound(keys, 0, n, uint64_t(c) << shift);
    unsigned e = (c + 1 == numCells) ? n : lowerBound(keys, b, n, uint64_t(c + 1) << shift);
    counts[c]  = e - b;
}

struct GatherArgs
{
    const void* src[16];
    void*       dst[16];
    int         bytes[16];
    int         count;
};

template<class V>
__device__ __forceinline__ void gatherOne(const void* src, void* dst, unsigned i, unsigned j)
{
    static_cast<V*>(dst)[i] = static_cast<const V*>(src)[j];
}

//! all listed arrays in one pass: the permutation is read once, the scattered reads of one particle's fields overlap
__global__ void reorderKernel(const unsigned* __restrict__ order, unsigned n, GatherArgs g)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned j = order[i];
#pragma unroll 4
    for (int k = 0; k < g.count; ++k)
    {
        switch (g.bytes[k])
        {
            case 8: gatherOne<uint64_t>(g.src[k], g.dst[k], i, j); break;
            case 4: gatherOne<uint32_t>(g.src[k], g.dst[k], i, j); break;
            case 2: gatherOne<uint16_t>(g.src[k], g.dst[k], i, j); break;
            default: gatherOne<uint8_t>(g.src[k], g.dst[k], i, j); break;
        }
    }
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

//! scratch carving shared by sphx_domain_sync_bytes and sphx_domain_sync
struct SyncScratch
{
    size_t keysIn, iotaIn, nodeStart, nodePBegin, nodePEnd, flags, pos, leafNodes, leafNodesSorted, leafKeys,
        leafKeysSorted, numSelected, extrema, cubTemp, cubTempBytes, total;

    SyncScratch(size_t n, int maxNodes)
    {
        size_t o = 0;
        auto   take = [&](size_t bytes)
        {
            size_t at = o;
            o += align256(bytes);
            return at;
        };
        size_t m        = size_t(maxNodes) + 8;
        keysIn          = take(n * 8);
        iotaIn          = take(n * 4);
        nodeStart       = take(m * 8);
        nodePBegin      = take(m * 4);
        nodePEnd        = take(m * 4);
        flags           = take(m * 4);
        pos             = take(m * 4);
        leafNodes       = take(m * 4);
        leafNodesSorted = take(m * 4);
        leafKeys        = take(m * 8);
        leafKeysSorted  = take(m * 8);
        numSelected     = take(16);
        extrema         = take(64);
        size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t1, (uint64_t*)nullptr, (uint64_t*)nullptr, (unsigned*)nullptr,
                                        (unsigned*)nullptr, int(n));
        cub::DeviceRadixSort::SortPairs(nullptr, t2, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr,
                                        (int*)nullptr, int(m));
        cub::DeviceScan::ExclusiveSum(nullptr, t3, (int*)nullptr, (int*)nullptr, int(m));
        cub::DeviceSelect::Flagged(nullptr, t4, cub::CountingInputIterator<int>(0), (int*)nullptr, (int*)nullptr,
                                   (int*)nullptr, int(m));
        cubTempBytes = std::max(std::max(t1, t2), std::max(t3, t4));
        cubTemp      = take(cubTempBytes);
        total        = o;
    }
};

bool g_tablesUploaded[64] = {};

cudaError_t uploadHilbertTables()
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_tablesUploaded[dev]) return cudaSuccess;
    uint8_t digit[kMaxHilbertStates * 8] = {}, next[kMaxHilbertStates * 8] = {}, octant[kMaxHilbertStates * 8] = {};
    int     ns = hilbertTablesFlat(digit, next, octant, kMaxHilbertStates);
    if (ns <= 0) return cudaErrorInvalidValue;
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_hDigit, digit, sizeof(digit))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_hNext, next, sizeof(next))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_hOctant, octant, sizeof(octant))) != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) g_tablesUploaded[dev] = true;
    return cudaSuccess;
}

int syncFail(int code, const std::string& msg)
{
    setLastError(msg);
    return code;
}

#define SYNC_CUDA(call)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) return syncFail(SPHX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

} // namespace
} // namespace sphx

extern "C"
{

size_t sphx_domain_sync_bytes(size_t n, int maxNodes) { return sphx::SyncScratch(n, maxNodes).total; }

int sphx_domain_sync(const SphxSyncArgs* a, SphxBox* boxOut, int* numNodesOut, int* numLeafNodesOut)
{
    using namespace sphx;
    if (int st = sphx_device_check()) return st;
    const bool presorted = a && (a->flags & SPHX_SYNC_PRESORTED);
    const bool noTree    = a && (a->flags & SPHX_SYNC_NO_TREE);
    if (!a || !a->x || !a->y || !a->z || !a->keys || (!presorted && !a->order) || !a->scratch)
        return syncFail(SPHX_ERR_INVALID, "sphx_domain_sync: null argument");
    if (!noTree && (!a->childOffsets || !a->internalToLeaf || !a->levelRange || !a->leaves || !a->layout ||
                    !a->centers || !a->sizes || !a->prefixes))
        return syncFail(SPHX_ERR_INVALID, "sphx_domain_sync: null tree buffer");
    if (a->n >= (size_t(1) << 31)) return syncFail(SPHX_ERR_INVALID, "sphx_domain_sync: bad particle count");
    if (!noTree && a->maxNodes < 9) return syncFail(SPHX_ERR_INVALID, "sphx_domain_sync: maxNodes too small");
    if (a->n == 0)
    {
        // a rank without particles (more ranks than occupied cells): nothing to sort; the tree is the empty root leaf
        cudaStream_t st = static_cast<cudaStream_t>(a->stream);
        if (boxOut) *boxOut = a->box;
        if (!noTree)
        {
            const SphxBox& b = a->box;
            const int      lr[23] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
            const uint64_t lv[2] = {0, uint64_t(1) << 63}, pre = 1;
            const unsigned lay[2] = {0, 0};
            const int      zero   = 0;
            double         ctr[3], sz[3];
            for (int d = 0; d < 3; ++d)
                ctr[d] = 0.5 * (b.lim[2 * d] + b.lim[2 * d + 1]), sz[d] = 0.5 * (b.lim[2 * d + 1] - b.lim[2 * d]);
            SYNC_CUDA(cudaMemcpyAsync(a->levelRange, lr, sizeof(lr), cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaMemcpyAsync(a->leaves, lv, sizeof(lv), cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaMemcpyAsync(a->prefixes, &pre, 8, cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaMemcpyAsync(a->layout, lay, sizeof(lay), cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaMemcpyAsync(a->childOffsets, &zero, 4, cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaMemcpyAsync(a->internalToLeaf, &zero, 4, cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaMemcpyAsync(a->centers, ctr, sizeof(ctr), cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaMemcpyAsync(a->sizes, sz, sizeof(sz), cudaMemcpyHostToDevice, st));
            SYNC_CUDA(cudaStreamSynchronize(st));
        }
        if (numNodesOut) *numNodesOut = noTree ? 0 : 1;
        if (numLeafNodesOut) *numLeafNodesOut = noTree ? 0 : 1;
        return SPHX_OK;
    }
    SyncScratch s(a->n, a->maxNodes);
    if (a->scratchBytes < s.total)
        return syncFail(SPHX_ERR_WORKSPACE, "sphx_domain_sync: scratch too small, need " + std::to_string(s.total));
    cudaStream_t stream = static_cast<cudaStream_t>(a->stream);
    SYNC_CUDA(uploadHilbertTables());

    char*     base   = static_cast<char*>(a->scratch);
    auto      at     = [&](size_t off) { return static_cast<void*>(base + off); };
    uint64_t* keysIn = static_cast<uint64_t*>(at(s.keysIn));
    unsigned* iotaIn = static_cast<unsigned*>(at(s.iotaIn));
    unsigned  n      = unsigned(a->n);

    SphxBox box = a->box;
    if (boxOut && (box.boundary[0] != 1 || box.boundary[1] != 1 || box.boundary[2] != 1))
    {
        auto*              ext = static_cast<unsigned long long*>(at(s.extrema));
        unsigned long long init[6], res[6];
        for (int d = 0; d < 3; ++d)
            init[2 * d] = ~0ull, init[2 * d + 1] = 0ull;
        SYNC_CUDA(cudaMemcpyAsync(ext, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        extremaKernel<<<592, 256, 0, stream>>>(a->x, a->y, a->z, n, ext);
        SYNC_CUDA(cudaMemcpyAsync(res, ext, sizeof(res), cudaMemcpyDeviceToHost, stream));
        SYNC_CUDA(cudaStreamSynchronize(stream));
        for (int d = 0; d < 3; ++d)
        {
            if (box.boundary[d] == 1) continue;
            double lo = fromOrderedBits(res[2 * d]), hi = fromOrderedBits(res[2 * d + 1]);
            if (a->flags & SPHX_SYNC_LIMIT_SHRINK)
            {
                // limitBoxShrinking (sfc/box.hpp:397-414; domain/assignment.hpp:80-82): after the first sync a side of
                // the box moves inwards by at most 5 % of the previous extent
                const double pl = a->box.lim[2 * d], ph = a->box.lim[2 * d + 1], ext = ph - pl;
                lo = std::min(lo, pl + 0.05 * ext), hi = std::max(hi, ph - 0.05 * ext);
            }
            box.lim[2 * d] = lo, box.lim[2 * d + 1] = hi;
        }
    }
    if (boxOut) *boxOut = box;

    KeyBox kb;
    kb.xmin = box.lim[0], kb.ymin = box.lim[2], kb.zmin = box.lim[4];
    kb.lx = box.lim[1] - box.lim[0], kb.ly = box.lim[3] - box.lim[2], kb.lz = box.lim[5] - box.lim[4];
    kb.mx = kMaxCoord * (1.0 / kb.lx), kb.my = kMaxCoord * (1.0 / kb.ly), kb.mz = kMaxCoord * (1.0 / kb.lz);

    size_t tempBytes = s.cubTempBytes;
    if (presorted)
    {
        // particles are in SFC order already (e.g. [halos | assigned | halos] of a rank): keys only
        hilbertKeysKernel<<<(n + 255) / 256, 256, 0, stream>>>(a->x, a->y, a->z, n, kb, a->keys,
                                                              a->order ? a->order : iotaIn);
    }
    else
    {
        hilbertKeysKernel<<<(n + 255) / 256, 256, 0, stream>>>(a->x, a->y, a->z, n, kb, keysIn, iotaIn);
        SYNC_CUDA(cub::DeviceRadixSort::SortPairs(at(s.cubTemp), tempBytes, keysIn, a->keys, iotaIn, a->order, int(n),
                                                  0, 3 * kMaxLevel, stream));
    }
    if (noTree)
    {
        SYNC_CUDA(cudaGetLastError());
        SYNC_CUDA(cudaStreamSynchronize(stream));
        if (numNodesOut) *numNodesOut = 0;
        if (numLeafNodesOut) *numLeafNodesOut = 0;
        return SPHX_OK;
    }

    // ---- top-down tree ----
    NodeArrays nd{static_cast<uint64_t*>(at(s.nodeStart)), static_cast<unsigned*>(at(s.nodePBegin)),
                  static_cast<unsigned*>(at(s.nodePEnd))};
    int* flags = static_cast<int*>(at(s.flags));
    int* pos   = static_cast<int*>(at(s.pos));
    {
        uint64_t s0 = 0;
        unsigned b0 = 0;
        SYNC_CUDA(cudaMemcpyAsync(nd.start, &s0, 8, cudaMemcpyHostToDevice, stream));
        SYNC_CUDA(cudaMemcpyAsync(nd.pBegin, &b0, 4, cudaMemcpyHostToDevice, stream));
        SYNC_CUDA(cudaMemcpyAsync(nd.pEnd, &n, 4, cudaMemcpyHostToDevice, stream));
        SYNC_CUDA(cudaStreamSynchronize(stream)); // the sources above are stack variables
    }
    int levelRange[kMaxLevel + 2];
    int off = 0, cnt = 1, level = 0;
    for (; level <= kMaxLevel; ++level)
    {
        levelRange[level] = off;
        openFlagsKernel<<<(cnt + 255) / 256, 256, 0, stream>>>(nd, off, cnt, a->bucketSize, level, flags);
        tempBytes = s.cubTempBytes;
        SYNC_CUDA(cub::DeviceScan::ExclusiveSum(at(s.cubTemp), tempBytes, flags, pos, cnt, stream));
        int last[2];
        SYNC_CUDA(cudaMemcpyAsync(&last[0], flags + cnt - 1, 4, cudaMemcpyDeviceToHost, stream));
        SYNC_CUDA(cudaMemcpyAsync(&last[1], pos + cnt - 1, 4, cudaMemcpyDeviceToHost, stream));
        SYNC_CUDA(cudaStreamSynchronize(stream));
        int numOpen = last[0] + last[1];
        int nextOff = off + cnt;
        if (size_t(nextOff) + size_t(numOpen) * 8 > size_t(a->maxNodes))
            return syncFail(SPHX_ERR_WORKSPACE, "sphx_domain_sync: octree needs more than maxNodes nodes");
        emitChildrenKernel<<<(cnt * 8 + 255) / 256, 256, 0, stream>>>(nd, a->keys, off, cnt, level, flags, pos, nextOff,
                                                                     a->childOffsets);
        off = nextOff;
        cnt = numOpen * 8;
        if (cnt == 0) break;
    }
    int numNodes = off;
    for (int l = level + 1; l <= kMaxLevel + 1; ++l)
        levelRange[l] = numNodes;
    SYNC_CUDA(cudaMemcpyAsync(a->levelRange, levelRange, sizeof(levelRange), cudaMemcpyHostToDevice, stream));

    // ---- leaves in SFC order ----
    int* leafNodes       = static_cast<int*>(at(s.leafNodes));
    int* leafNodesSorted = static_cast<int*>(at(s.leafNodesSorted));
    auto* leafKeys       = static_cast<uint64_t*>(at(s.leafKeys));
    auto* leafKeysSorted = static_cast<uint64_t*>(at(s.leafKeysSorted));
    int*  numSelected    = static_cast<int*>(at(s.numSelected));
    leafFlagsKernel<<<(numNodes + 255) / 256, 256, 0, stream>>>(a->childOffsets, numNodes, flags);
    tempBytes = s.cubTempBytes;
    SYNC_CUDA(cub::DeviceSelect::Flagged(at(s.cubTemp), tempBytes, cub::CountingInputIterator<int>(0), flags, leafNodes,
                                         numSelected, numNodes, stream));
    int numLeaves = 0;
    SYNC_CUDA(cudaMemcpyAsync(&numLeaves, numSelected, 4, cudaMemcpyDeviceToHost, stream));
    SYNC_CUDA(cudaStreamSynchronize(stream)); // also makes the levelRange source safe to drop
    gatherLeafKeysKernel<<<(numLeaves + 255) / 256, 256, 0, stream>>>(leafNodes, numLeaves, nd.start, leafKeys);
    tempBytes = s.cubTempBytes;
    SYNC_CUDA(cub::DeviceRadixSort::SortPairs(at(s.cubTemp), tempBytes, leafKeys, leafKeysSorted, leafNodes,
                                              leafNodesSorted, numLeaves, 0, 3 * kMaxLevel, stream));
    linkLeavesKernel<<<(numLeaves + 1 + 255) / 256, 256, 0, stream>>>(leafNodesSorted, numLeaves, nd, n,
                                                                     a->internalToLeaf, a->leaves, a->layout);
    nodeGeometryKernel<<<(numNodes + 255) / 256, 256, 0, stream>>>(nd, numNodes, a->levelRange, kb, a->internalToLeaf,
                                                                  a->childOffsets, a->prefixes, a->centers, a->sizes);
    SYNC_CUDA(cudaGetLastError());
    SYNC_CUDA(cudaStreamSynchronize(stream));
    if (numNodesOut) *numNodesOut = numNodes;
    if (numLeafNodesOut) *numLeafNodesOut = numLeaves;
    return SPHX_OK;
}

int sphx_cell_histogram(const uint64_t* sortedKeys, size_t n, int level, unsigned* counts, void* stream)
{
    using namespace sphx;
    if (int st = sphx_device_check()) return st;
    if (!counts || level < 0 || level > 10 || n >= (size_t(1) << 32) || (n && !sortedKeys))
        return syncFail(SPHX_ERR_INVALID, "sphx_cell_histogram: bad argument");
    unsigned numCells = 1u << (3 * level);
    cellHistogramKernel<<<(numCells + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sortedKeys, unsigned(n), 3 * (kMaxLevel - level), numCells, counts);
    SYNC_CUDA(cudaGetLastError());
    return SPHX_OK;
}

int sphx_reorder_fields(const unsigned* order, size_t n, int count, const void* const* src, void* const* dst,
                        const int* elemBytes, void* stream)
{
    using namespace sphx;
    if (int st = sphx_device_check()) return st;
    if (!order || !src || !dst || !elemBytes || count < 0 || count > 16 || n >= (size_t(1) << 32))
        return syncFail(SPHX_ERR_INVALID, "sphx_reorder_fields: bad argument");
    if (count == 0 || n == 0) return SPHX_OK;
    GatherArgs g;
    g.count = count;
    for (int k = 0; k < count; ++k)
    {
        if (!src[k] || !dst[k] || src[k] == dst[k])
            return syncFail(SPHX_ERR_INVALID, "sphx_reorder_fields: null or aliased array");
        int b = elemBytes[k];
        if (b != 1 && b != 2 && b != 4 && b != 8) return syncFail(SPHX_ERR_INVALID, "sphx_reorder_fields: elemBytes");
        g.src[k] = src[k], g.dst[k] = dst[k], g.bytes[k] = b;
    }
    reorderKernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(order, unsigned(n), g);
    SYNC_CUDA(cudaGetLastError());
    return SPHX_OK;
}

} // extern "C"

/* ------------------------------ decomposition plan on the device (multi-rank Domain::sync) ------------------------------ */

namespace sphx
{
namespace
{

struct PlanDev
{
    SphxCellPlanSummary sum;
    unsigned            sendBase; // running offset into sendIdx while the per-rank lists are laid out
};

//! rank that owns cell c: splits[r] <= c < splits[r + 1] (empty ranges are skipped)
__device__ __forceinline__ int ownerOf(const uint64_t* __restrict__ splits, int nranks, uint64_t c)
{
    int r = 0;
    while (r + 1 < nranks && c >= splits[r + 1])
        ++r;
    return r;
}

struct WidenCounts
{
    const unsigned* g;
    __device__ uint64_t operator()(unsigned c) const { return g[c]; }
};

//! uniformBins (domaindecomp.hpp:99-110) on the cell prefix sums + the migration offsets of this rank's particles
__global__ void planSplitsKernel(PlanDev* p, const uint64_t* __restrict__ prefixG, const unsigned* __restrict__ prefixL,
                                 unsigned ncell, int rank, int nranks)
{
    __shared__ uint64_t cand[SPHX_MAX_RANKS + 1];
    const int           r = threadIdx.x;
    const uint64_t      N = prefixG[ncell];
    if (r >= 1 && r < nranks)
    {
        // first index in prefixG[0 .. ncell] whose value is >= r N / R  (N r < 2^64: N < 2^58, r < 64)
        const uint64_t target = (N * uint64_t(r)) / uint64_t(nranks);
        unsigned       lo = 0, hi = ncell + 1;
        while (lo < hi)
        {
            unsigned mid = lo + (hi - lo) / 2;
            if (prefixG[mid] < target) { lo = mid + 1; }
            else { hi = mid; }
        }
        cand[r] = min(uint64_t(lo), uint64_t(ncell));
    }
    __syncthreads();
    if (r == 0)
    {
        SphxCellPlanSummary& s = p->sum;
        s.cellSplits[0]        = 0;
        for (int q = 1; q < nranks; ++q)
            s.cellSplits[q] = max(cand[q], s.cellSplits[q - 1]);
        s.cellSplits[nranks] = ncell;
        for (int q = 0; q <= nranks; ++q)
            s.sendOffLocal[q] = prefixL[s.cellSplits[q]];
        s.nGlobal   = N;
        s.nAssigned = prefixG[s.cellSplits[rank + 1]] - prefixG[s.cellSplits[rank]];
        s.nHaloLeft = s.nHaloRight = 0;
        for (int q = 0; q < SPHX_MAX_RANKS; ++q)
            s.recvCount[q] = s.sendCount[q] = 0;
        s.numRecvCells = s.numSend = s.overflow = s.pad = 0;
        p->sendBase                              = 0;
    }
}

/*! one thread per cell of this rank: which non-empty foreign cells lie within its reach (-> halo cells of this rank) and
 *  which foreign cells have this cell within THEIR reach (-> this cell is sent to their owners). The reach of a cell is
 *  rings[c] rings of cells (Chebyshev distance; 1 if rings is null); everything follows from the global arrays, so no
 *  request messages are needed. */
__global__ void planAdjacencyKernel(const PlanDev* __restrict__ p, const unsigned* __restrict__ G,
                                    const unsigned char* __restrict__ rings, int maxRing, int level, int perX,
                                    int perY, int perZ, int rank, int nranks, unsigned char* __restrict__ recvFlag,
                                    uint64_t* __restrict__ sendMask)
{
    __shared__ uint8_t sDigit[kMaxHilbertStates * 8], sNext[kMaxHilbertStates * 8], sOct[kMaxHilbertStates * 8];
    __shared__ uint64_t sSplits[SPHX_MAX_RANKS + 1];
    for (int k = threadIdx.x; k < kMaxHilbertStates * 8; k += blockDim.x)
        sDigit[k] = c_hDigit[k], sNext[k] = c_hNext[k], sOct[k] = c_hOctant[k];
    for (int k = threadIdx.x; k <= nranks; k += blockDim.x)
        sSplits[k] = p->sum.cellSplits[k];
    __syncthreads();
    const uint64_t cb = sSplits[rank], ce = sSplits[rank + 1];
    const uint64_t c  = cb + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ce || G[c] == 0) return;

    int      cx = 0, cy = 0, cz = 0;
    unsigned state = 0;
    for (int l = level - 1; l >= 0; --l)
    {
        unsigned d = unsigned(c >> (3 * l)) & 7u;
        unsigned o = sOct[state * 8 + d];
        cx |= int((o >> 2) & 1u) << l, cy |= int((o >> 1) & 1u) << l, cz |= int(o & 1u) << l;
        state = sNext[state * 8 + o];
    }
    const int side = 1 << level;
    const int Rc   = rings ? max(1, int(rings[c])) : 1;
    uint64_t  mask = 0;
    for (int dz = -maxRing; dz <= maxRing; ++dz)
        for (int dy = -maxRing; dy <= maxRing; ++dy)
            for (int dx = -maxRing; dx <= maxRing; ++dx)
            {
                if (!dx && !dy && !dz) continue;
                const int d  = max(abs(dx), max(abs(dy), abs(dz)));
                int       nx = cx + dx, ny = cy + dy, nz = cz + dz;
                if (nx < 0 || nx >= side) { if (!perX) continue; nx = ((nx % side) + side) % side; }
                if (ny < 0 || ny >= side) { if (!perY) continue; ny = ((ny % side) + side) % side; }
                if (nz < 0 || nz >= side) { if (!perZ) continue; nz = ((nz % side) + side) % side; }
                uint64_t c2 = 0;
                unsigned st = 0;
                for (int l = level - 1; l >= 0; --l)
                {
                    unsigned o = ((unsigned(nx) >> l) & 1u) << 2 | ((unsigned(ny) >> l) & 1u) << 1 | ((unsigned(nz) >> l) & 1u);
                    c2         = (c2 << 3) | sDigit[st * 8 + o];
                    st         = sNext[st * 8 + o];
                }
                if (c2 >= cb && c2 < ce) continue;
                if (G[c2] == 0) continue;
                if (d <= Rc) recvFlag[c2] = 1; // several threads may store the same value
                const int R2 = rings ? max(1, int(rings[c2])) : 1;
                if (d <= R2) mask |= uint64_t(1) << ownerOf(sSplits, nranks, c2);
            }
    sendMask[c] = mask;
}

__global__ void planRecvCountKernel(PlanDev* p, const unsigned* __restrict__ G, const unsigned char* __restrict__ recvFlag,
                                    unsigned ncell, int nranks)
{
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell || !recvFlag[c]) return;
    atomicAdd(&p->sum.recvCount[ownerOf(p->sum.cellSplits, nranks, c)], G[c]);
    atomicAdd(&p->sum.numRecvCells, 1u);
}

__global__ void planHaloSizesKernel(PlanDev* p, int rank, int nranks)
{
    uint64_t left = 0, right = 0;
    for (int r = 0; r < nranks; ++r)
        (r < rank ? left : right) += p->sum.recvCount[r];
    p->sum.nHaloLeft = left, p->sum.nHaloRight = right;
}

//! particles of cell c that go to rank r (0 if the cell is not sent there)
struct SendCountOf
{
    const unsigned* G;
    const uint64_t* sendMask;
    int             r;
    __device__ unsigned operator()(unsigned c) const { return ((sendMask[c] >> r) & 1u) ? G[c] : 0u; }
};

//! local indices (layout [halos | assigned | halos]) of the particles of the cells sent to rank r, in SFC order
__global__ void planFillSendKernel(PlanDev* p, const unsigned* __restrict__ G, const uint64_t* __restrict__ sendMask,
                                   const uint64_t* __restrict__ prefixG, const unsigned* __restrict__ offs, int r,
                                   int rank, unsigned ncell, unsigned* __restrict__ sendIdx, size_t capacity)
{
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell || !((sendMask[c] >> r) & 1u)) return;
    const uint64_t cb    = p->sum.cellSplits[rank];
    const unsigned first = unsigned(p->sum.nHaloLeft + (prefixG[c] - prefixG[cb]));
    const size_t   base  = size_t(p->sendBase) + offs[c];
    const unsigned n     = G[c];
    if (base + n > capacity)
    {
        p->sum.overflow = 1;
        return;
    }
    for (unsigned k = 0; k < n; ++k)
        sendIdx[base + k] = first + k;
}

__global__ void planAdvanceKernel(PlanDev* p, const unsigned* __restrict__ G, const uint64_t* __restrict__ sendMask,
                                  const unsigned* __restrict__ offs, int r, unsigned ncell)
{
    unsigned last = ncell - 1;
    unsigned tot  = offs[last] + (((sendMask[last] >> r) & 1u) ? G[last] : 0u);
    p->sum.sendCount[r] = tot;
    p->sendBase += tot;
    p->sum.numSend = p->sendBase;
}

struct PlanScratch
{
    size_t prefixG, prefixL, recvFlag, sendMask, offs, numSel, dev, cubTemp, cubBytes, total;
    explicit PlanScratch(int level)
    {
        size_t ncell = size_t(1) << (3 * level);
        auto   al    = [](size_t v) { return (v + 255) / 256 * 256; };
        size_t off   = 0;
        prefixG = off, off = al(off + (ncell + 1) * sizeof(uint64_t));
        prefixL = off, off = al(off + (ncell + 1) * sizeof(unsigned));
        recvFlag = off, off = al(off + ncell);
        sendMask = off, off = al(off + ncell * sizeof(uint64_t));
        offs = off, off = al(off + ncell * sizeof(unsigned));
        numSel = off, off = al(off + 16);
        dev = off, off = al(off + sizeof(PlanDev));
        size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        cub::TransformInputIterator<uint64_t, WidenCounts, cub::CountingInputIterator<unsigned>> wide(
            cub::CountingInputIterator<unsigned>(0), WidenCounts{nullptr});
        cub::DeviceScan::InclusiveSum(nullptr, t1, wide, (uint64_t*)nullptr, int(ncell));
        cub::DeviceScan::InclusiveSum(nullptr, t2, (const unsigned*)nullptr, (unsigned*)nullptr, int(ncell));
        cub::TransformInputIterator<unsigned, SendCountOf, cub::CountingInputIterator<unsigned>> sc(
            cub::CountingInputIterator<unsigned>(0), SendCountOf{nullptr, nullptr, 0});
        cub::DeviceScan::ExclusiveSum(nullptr, t3, sc, (unsigned*)nullptr, int(ncell));
        cub::DeviceSelect::Flagged(nullptr, t4, cub::CountingInputIterator<unsigned>(0), (const unsigned char*)nullptr,
                                   (unsigned*)nullptr, (unsigned*)nullptr, int(ncell));
        cubBytes = std::max(std::max(t1, t2), std::max(t3, t4));
        cubTemp = off, off = al(off + cubBytes);
        total = off;
    }
};

} // namespace
} // namespace sphx

extern "C"
{

size_t sphx_cell_plan_device_bytes(int level)
{
    if (level < 0 || level > 10) return 0;
    return sphx::PlanScratch(level).total;
}

int sphx_cell_plan_build_device(const unsigned* globalCounts, const unsigned* localCounts, const unsigned char* rings,
                                int maxRing, int level, const int* periodic, int rank, int nranks, void* scratch,
                                size_t scratchBytes, unsigned* sendIdx, size_t sendCapacity, unsigned* recvCells,
                                SphxCellPlanSummary* out, void* stream)
{
    using namespace sphx;
    if (int st = sphx_device_check()) return st;
    if (!globalCounts || !localCounts || level < 0 || level > 10 || !periodic || nranks < 1 || nranks > SPHX_MAX_RANKS ||
        rank < 0 || rank >= nranks || !scratch || !out || (sendCapacity && !sendIdx) || maxRing < 1 || maxRing > 16)
        return syncFail(SPHX_ERR_INVALID, "sphx_cell_plan_build_device: bad argument");
    PlanScratch s(level);
    if (scratchBytes < s.total)
        return syncFail(SPHX_ERR_WORKSPACE,
                        "sphx_cell_plan_build_device: scratch too small, need " + std::to_string(s.total));
    SYNC_CUDA(uploadHilbertTables());
    auto           cs    = static_cast<cudaStream_t>(stream);
    char*          base  = static_cast<char*>(scratch);
    const unsigned ncell = 1u << (3 * level);
    auto*          prefixG  = reinterpret_cast<uint64_t*>(base + s.prefixG);
    auto*          prefixL  = reinterpret_cast<unsigned*>(base + s.prefixL);
    auto*          recvFlag = reinterpret_cast<unsigned char*>(base + s.recvFlag);
    auto*          sendMask = reinterpret_cast<uint64_t*>(base + s.sendMask);
    auto*          offs     = reinterpret_cast<unsigned*>(base + s.offs);
    auto*          numSel   = reinterpret_cast<unsigned*>(base + s.numSel);
    auto*          dev      = reinterpret_cast<PlanDev*>(base + s.dev);
    void*          tmp      = base + s.cubTemp;
    size_t         tmpBytes = s.cubBytes;

    SYNC_CUDA(cudaMemsetAsync(prefixG, 0, sizeof(uint64_t), cs));
    SYNC_CUDA(cudaMemsetAsync(prefixL, 0, sizeof(unsigned), cs));
    SYNC_CUDA(cudaMemsetAsync(recvFlag, 0, ncell, cs));
    SYNC_CUDA(cudaMemsetAsync(sendMask, 0, size_t(ncell) * sizeof(uint64_t), cs));
    cub::TransformInputIterator<uint64_t, WidenCounts, cub::CountingInputIterator<unsigned>> wide(
        cub::CountingInputIterator<unsigned>(0), WidenCounts{globalCounts});
    SYNC_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmpBytes, wide, prefixG + 1, int(ncell), cs));
    tmpBytes = s.cubBytes;
    SYNC_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmpBytes, localCounts, prefixL + 1, int(ncell), cs));
    planSplitsKernel<<<1, SPHX_MAX_RANKS, 0, cs>>>(dev, prefixG, prefixL, ncell, rank, nranks);
    // (the own range is at most ncell cells: threads beyond it return at once)
    planAdjacencyKernel<<<(ncell + 127) / 128, 128, 0, cs>>>(dev, globalCounts, rings, rings ? maxRing : 1, level,
                                                              periodic[0], periodic[1], periodic[2], rank, nranks,
                                                              recvFlag, sendMask);
    planRecvCountKernel<<<(ncell + 255) / 256, 256, 0, cs>>>(dev, globalCounts, recvFlag, ncell, nranks);
    planHaloSizesKernel<<<1, 1, 0, cs>>>(dev, rank, nranks);
    if (recvCells)
    {
        tmpBytes = s.cubBytes;
        SYNC_CUDA(cub::DeviceSelect::Flagged(tmp, tmpBytes, cub::CountingInputIterator<unsigned>(0), recvFlag, recvCells,
                                             numSel, int(ncell), cs));
    }
    for (int r = 0; r < nranks; ++r)
    {
        if (r == rank) continue;
        cub::TransformInputIterator<unsigned, SendCountOf, cub::CountingInputIterator<unsigned>> sc(
            cub::CountingInputIterator<unsigned>(0), SendCountOf{globalCounts, sendMask, r});
        tmpBytes = s.cubBytes;
        SYNC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, sc, offs, int(ncell), cs));
        planFillSendKernel<<<(ncell + 255) / 256, 256, 0, cs>>>(dev, globalCounts, sendMask, prefixG, offs, r, rank,
                                                                ncell, sendIdx, sendCapacity);
        planAdvanceKernel<<<1, 1, 0, cs>>>(dev, globalCounts, sendMask, offs, r, ncell);
    }
    SYNC_CUDA(cudaGetLastError());
    SYNC_CUDA(cudaMemcpyAsync(out, &dev->sum, sizeof(SphxCellPlanSummary), cudaMemcpyDeviceToHost, cs));
    SYNC_CUDA(cudaStreamSynchronize(cs));
    if (out->overflow) return syncFail(SPHX_ERR_WORKSPACE, "sphx_cell_plan_build_device: send index capacity too small");
    return SPHX_OK;
}

} // extern "C"

/*! @file
 * Generic neighbour search with the cstone::findNeighbors call shape (any ngmax, particle-index lists, no h-iteration):
 * the utility behind sphx_find_neighbors. The hydro step itself uses the block search in search.cu.
 *
 * Replaces (reference paths relative to /root/reference):
 *   cstone::findNeighbors            domain/include/cstone/findneighbors.hpp:77-170   (the predicate we must match)
 *   cstone::traverseNeighbors        domain/include/cstone/traversal/find_neighbors.cuh:182-489 (not followed)
 *
 * Design: one warp owns 32 SFC-consecutive targets, one per lane. The octree is walked with
 * a per-warp stack in shared memory, 32 nodes per step, pruned with the bounding box of the warp's 2h-spheres. Every
 * leaf that survives is tested per lane with the reference CPU's own point<->cell criterion, its particles are staged
 * in shared memory and broadcast to all lanes, and hits are appended to the lane's column of a lane-interleaved ELL
 * list.
 * The pair predicate is evaluated in un-contracted fp64 exactly as the reference CPU does, so counts and sets are
 * bit-exact; the order inside a list is traversal order (compare after sorting).
 */
#include "sphx_device.cuh"
#include "sphx_kernels.h"

namespace sphx
{

constexpr int kWarpsPerBlock = 4;
constexpr int kStackSize     = 1024; // ints per warp
constexpr int kLeafBatch     = 64;   // particles staged per pass (bucket size of the focus tree)

struct WarpShared
{
    int    stack[kStackSize];
    double sx[kLeafBatch], sy[kLeafBatch], sz[kLeafBatch];
};

//! Th: type of the smoothing lengths (float in the production type set, double in the all-double one)
template<class Th>
struct Target
{
    double x, y, z;   // position
    Th     h;         // smoothing length
    Th     radiusSq;  // Th(4) * h * h   (findneighbors.hpp:93)
    Th     cellRadSq; // radiusSq * searchExtFactor^2 (findneighbors.hpp:94)
    bool   usePbc;    // anyPbc && !insideBox(particle, 2h) (findneighbors.hpp:98-100)
    bool   valid;
};

__device__ __forceinline__ float  mulRn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mulRn(double a, double b) { return __dmul_rn(a, b); }

//! sph::updateH (kernels.hpp:26-32) for T = double (the float version emulates glibc's powf bit for bit)
__device__ __forceinline__ double updateH(unsigned ng0, unsigned nc, double h)
{
    return h * 0.5 * pow(1.0 + 1023.0 * double(ng0) / double(nc), 1.0 / 10.0);
}

template<class Th>
__device__ __forceinline__ void setupTarget(Target<Th>& t, const DevBox& box, float searchExt)
{
    t.radiusSq  = mulRn(mulRn(Th(4), t.h), t.h);
    t.cellRadSq = mulRn(mulRn(t.radiusSq, Th(searchExt)), Th(searchExt));
    double ext  = __dmul_rn(2.0, double(t.h));
    bool inside = __dsub_rn(t.x, ext) >= box.xmin && __dsub_rn(t.y, ext) >= box.ymin &&
                  __dsub_rn(t.z, ext) >= box.zmin && __dadd_rn(t.x, ext) <= box.xmax &&
                  __dadd_rn(t.y, ext) <= box.ymax && __dadd_rn(t.z, ext) <= box.zmax;
    t.usePbc = box.anyPbc && !inside;
}

/*! @brief one full traversal for the 32 targets of a warp
 *
 * @param record   lanes that (re)build their list in this pass; the others only take part in the cooperative work
 * @return         per-lane neighbour count (self excluded, not capped) for recording lanes
 */
template<class Th>
__device__ unsigned traverseWarp(const Target<Th>& t, unsigned iSelf, bool record, const DevBox& box,
                                 const SphxTreeView& tree, const double* __restrict__ x, const double* __restrict__ y,
                                 const double* __restrict__ z, unsigned ngmax, unsigned* __restrict__ listCol,
                                 WarpShared& sm, unsigned* errFlags)
{
    const unsigned lane = laneId();

    // bounding box of the recording lanes' search spheres, slightly inflated: pruning only has to be conservative
    const double big = 1e300;
    double       r   = 2.0 * double(t.h) * double(tree.searchExtFactor);
    double bxmin = warpMin(record ? t.x - r : big), bxmax = warpMax(record ? t.x + r : -big);
    double bymin = warpMin(record ? t.y - r : big), bymax = warpMax(record ? t.y + r : -big);
    double bzmin = warpMin(record ? t.z - r : big), bzmax = warpMax(record ? t.z + r : -big);
    const double infl = 1.0 + 1e-9;
    double bcx = 0.5 * (bxmax + bxmin), bsx = 0.5 * (bxmax - bxmin) * infl + 1e-300;
    double bcy = 0.5 * (bymax + bymin), bsy = 0.5 * (bymax - bymin) * infl + 1e-300;
    double bcz = 0.5 * (bzmax + bzmin), bsz = 0.5 * (bzmax - bzmin) * infl + 1e-300;
    const bool warpPbc = __any_sync(kFullMask, record && t.usePbc);

    unsigned count = 0;

    int stackSize = 1;
    if (lane == 0) { sm.stack[0] = 0; }
    __syncwarp();

    while (stackSize > 0)
    {
        int  nPop = min(32, stackSize);
        int  node = -1;
        bool have = int(lane) < nPop;
        if (have) { node = sm.stack[stackSize - nPop + lane]; }
        stackSize -= nPop;
        __syncwarp();

        bool overlap = false;
        int  child   = 0;
        if (have)
        {
            double cx = tree.centers[3 * node], cy = tree.centers[3 * node + 1], cz = tree.centers[3 * node + 2];
            double sx = tree.sizes[3 * node], sy = tree.sizes[3 * node + 1], sz = tree.sizes[3 * node + 2];
            double dx = cx - bcx, dy = cy - bcy, dz = cz - bcz;
            if (warpPbc)
            {
                dx -= box.plx * rint(dx * box.ilx);
                dy -= box.ply * rint(dy * box.ily);
                dz -= box.plz * rint(dz * box.ilz);
            }
            overlap = fabs(dx) <= (sx + bsx) * infl && fabs(dy) <= (sy + bsy) * infl && fabs(dz) <= (sz + bsz) * infl;
            if (overlap) { child = tree.childOffsets[node]; }
        }

        bool     isInternal = overlap && child != 0;
        bool     isLeaf     = overlap && child == 0;
        unsigned intMask    = __ballot_sync(kFullMask, isInternal);
        unsigned leafMask   = __ballot_sync(kFullMask, isLeaf);

        int numPush = __popc(intMask) * 8;
        if (stackSize + numPush > kStackSize)
        {
            if (lane == 0) { atomicOr(errFlags, kErrTraversal); }
            return count;
        }
        if (isInternal)
        {
            int off = stackSize + 8 * __popc(intMask & ((1u << lane) - 1u));
#pragma unroll
            for (int c = 0; c < 8; ++c)
                sm.stack[off + c] = child + c;
        }
        stackSize += numPush;
        __syncwarp();

        while (leafMask)
        {
            int src = __ffs(leafMask) - 1;
            leafMask &= leafMask - 1;
            int leafNode = __shfl_sync(kFullMask, node, src);

            // per-lane exact point<->cell test of the reference CPU (findneighbors.hpp:102-106, boxoverlap.hpp:196-216)
            bool open = false;
            if (record)
            {
                double dx = __dsub_rn(tree.centers[3 * leafNode], t.x);
                double dy = __dsub_rn(tree.centers[3 * leafNode + 1], t.y);
                double dz = __dsub_rn(tree.centers[3 * leafNode + 2], t.z);
                if (t.usePbc)
                {
                    dx = foldExact(dx, box.plx, box.ilx);
                    dy = foldExact(dy, box.ply, box.ily);
                    dz = foldExact(dz, box.plz, box.ilz);
                }
                double mx = minDistComp(dx, tree.sizes[3 * leafNode]);
                double my = minDistComp(dy, tree.sizes[3 * leafNode + 1]);
                double mz = minDistComp(dz, tree.sizes[3 * leafNode + 2]);
                open      = sumSqRight(mx, my, mz) < double(t.cellRadSq);
            }
            if (!__any_sync(kFullMask, open)) { continue; }

            int      leafIdx = tree.internalToLeaf[leafNode];
            unsigned pBegin  = tree.layout[leafIdx];
            unsigned pEnd    = tree.layout[leafIdx + 1];

            for (unsigned base = pBegin; base < pEnd; base += kLeafBatch)
            {
                unsigned cnt = min(unsigned(kLeafBatch), pEnd - base);
                __syncwarp();
                for (unsigned s = lane; s < cnt; s += 32)
                {
                    sm.sx[s] = x[base + s];
                    sm.sy[s] = y[base + s];
                    sm.sz[s] = z[base + s];
                }
                __syncwarp();

                if (open)
                {
                    if (t.usePbc)
                    {
                        for (unsigned s = 0; s < cnt; ++s)
                        {
                            // distanceSq<true>(x[j], y[j], z[j], xi, yi, zi, box) (findneighbors.hpp:33-49,117)
                            double dx = foldExact(__dsub_rn(sm.sx[s], t.x), box.plx, box.ilx);
                            double dy = foldExact(__dsub_rn(sm.sy[s], t.y), box.ply, box.ily);
                            double dz = foldExact(__dsub_rn(sm.sz[s], t.z), box.plz, box.ilz);
                            double d2 = sumSqLeft(dx, dy, dz);
                            unsigned j = base + s;
                            if (d2 < double(t.radiusSq) && j != iSelf)
                            {
                                if (count < ngmax) { listCol[size_t(count) * kGroupSize] = j; }
                                count++;
                            }
                        }
                    }
                    else
                    {
                        for (unsigned s = 0; s < cnt; ++s)
                        {
                            // distanceSq<false> (findneighbors.hpp:52-60,134)
                            double dx = __dsub_rn(sm.sx[s], t.x);
                            double dy = __dsub_rn(sm.sy[s], t.y);
                            double dz = __dsub_rn(sm.sz[s], t.z);
                            double d2 = sumSqLeft(dx, dy, dz);
                            unsigned j = base + s;
                            if (d2 < double(t.radiusSq) && j != iSelf)
                            {
                                if (count < ngmax) { listCol[size_t(count) * kGroupSize] = j; }
                                count++;
                            }
                        }
                    }
                }
            }
        }
    }
    return count;
}

/*! @brief search + h-iteration for one warp; returns nc = 1 + count for valid lanes */
template<class Th>
__device__ unsigned searchWithHIteration(Target<Th>& t, unsigned i, const DevBox& box, const SphxTreeView& tree,
                                         const double* x, const double* y, const double* z, unsigned ng0,
                                         unsigned ngmax, unsigned* listCol, WarpShared& sm, StepScalars* scal,
                                         bool iterateH, bool& hChanged)
{
    setupTarget(t, box, tree.searchExtFactor);
    unsigned ncSph = 1 + traverseWarp(t, i, t.valid, box, tree, x, y, z, ngmax, listCol, sm, &scal->errFlags);
    hChanged       = false;
    if (!iterateH) { return ncSph; }

    // sph/find_neighbors.hpp:17-36: while (ngmin > nc || nc - 1 > ngmax) && iteration++ < 10
    const unsigned ngmin     = ng0 / 4;
    int            iteration = 0;
    while (true)
    {
        bool repeat = t.valid && (ngmin > ncSph || (ncSph - 1) > ngmax) && iteration < 10;
        if (!__any_sync(kFullMask, repeat)) { break; }
        if (repeat)
        {
            iteration++;
            t.h = updateH(ng0, ncSph, t.h);
            setupTarget(t, box, tree.searchExtFactor);
            hChanged = true;
        }
        unsigned c = 1 + traverseWarp(t, i, repeat, box, tree, x, y, z, ngmax, listCol, sm, &scal->errFlags);
        if (repeat) { ncSph = c; }
    }
    // reference: numFails += (iteration >= maxIteration) after the post-incremented loop test
    if (t.valid && (ngmin > ncSph || (ncSph - 1) > ngmax) && iteration >= 10) { atomicOr(&scal->errFlags, kErrHConv); }
    if (t.valid && (ncSph - 1) > ngmax) { atomicOr(&scal->errFlags, kErrNgmax); }
    return ncSph;
}

/*! @brief cstone::findNeighbors batch shape: ngmax-strided lists and counts, no h-iteration */
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    findNeighborsKernel(unsigned first, unsigned last, DevBox box, SphxTreeView tree, const double* __restrict__ x,
                        const double* __restrict__ y, const double* __restrict__ z, const float* __restrict__ h,
                        unsigned ngmax, unsigned* __restrict__ list, unsigned* __restrict__ counts, StepScalars* scal)
{
    __shared__ WarpShared shared[kWarpsPerBlock];
    const unsigned        warpInBlock = threadIdx.x >> 5;
    const unsigned        lane        = laneId();
    const size_t          group       = size_t(blockIdx.x) * kWarpsPerBlock + warpInBlock;
    const size_t          numGroups   = (size_t(last - first) + kGroupSize - 1) / kGroupSize;
    if (group >= numGroups) { return; }

    WarpShared& sm = shared[warpInBlock];
    unsigned    i  = first + unsigned(group) * kGroupSize + lane;
    Target<float> t;
    t.valid     = i < last;
    unsigned il = t.valid ? i : last - 1;
    t.x = x[il], t.y = y[il], t.z = z[il], t.h = h[il];

    unsigned* listCol = list + nbListIndex(group, ngmax, 0, lane);
    bool      hChanged;
    unsigned  ncSph = searchWithHIteration(t, i, box, tree, x, y, z, 0, ngmax, listCol, sm, scal, false, hChanged);
    if (t.valid) { counts[i - first] = ncSph - 1; }
}

/*! @brief sph::findNeighborsSph (sph/find_neighbors.hpp:11-44) for the all-double type set: search with the coupled
 *  h-iteration, writes h, nc = 1 + count and the lane-interleaved particle-index list */
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    findNeighborsSphF64Kernel(unsigned first, unsigned last, DevBox box, SphxTreeView tree,
                              const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                              double* __restrict__ h, unsigned ng0, unsigned ngmax, unsigned* __restrict__ list,
                              unsigned* __restrict__ nc, StepScalars* scal)
{
    __shared__ WarpShared shared[kWarpsPerBlock];
    const unsigned        warpInBlock = threadIdx.x >> 5;
    const unsigned        lane        = laneId();
    const size_t          group       = size_t(blockIdx.x) * kWarpsPerBlock + warpInBlock;
    const size_t          numGroups   = (size_t(last - first) + kGroupSize - 1) / kGroupSize;
    if (group >= numGroups) { return; }

    WarpShared&    sm = shared[warpInBlock];
    unsigned       i  = first + unsigned(group) * kGroupSize + lane;
    Target<double> t;
    t.valid     = i < last;
    unsigned il = t.valid ? i : last - 1;
    t.x = x[il], t.y = y[il], t.z = z[il], t.h = h[il];

    unsigned* listCol = list + nbListIndex(group, ngmax, 0, lane);
    bool      hChanged;
    unsigned  ncSph = searchWithHIteration(t, i, box, tree, x, y, z, ng0, ngmax, listCol, sm, scal, true, hChanged);
    if (t.valid)
    {
        nc[i] = ncSph;
        if (hChanged) { h[i] = t.h; }
    }
    unsigned long long sum = t.valid ? ncSph : 0;
    unsigned           mx  = t.valid ? ncSph : 0, it = (t.valid && hChanged) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1)
    {
        sum += __shfl_xor_sync(kFullMask, sum, o);
        mx = max(mx, __shfl_xor_sync(kFullMask, mx, o));
        it += __shfl_xor_sync(kFullMask, it, o);
    }
    if (lane == 0)
    {
        atomicAdd(&scal->totalNeighbors, sum);
        atomicMax(&scal->maxNc, mx);
        if (it) { atomicAdd(&scal->numHIterated, it); }
    }
}

//! lane-interleaved ELL -> reference CPU layout neighbors[(i-first)*ngmax + k]
__global__ void exportNeighborsKernel(unsigned numAssigned, unsigned ngmax, const unsigned* __restrict__ list,
                                      const unsigned* __restrict__ counts, bool countsIncludeSelf,
                                      unsigned* __restrict__ out)
{
    size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t tot = size_t(numAssigned) * ngmax;
    if (tid >= tot) { return; }
    unsigned t = unsigned(tid / ngmax), k = unsigned(tid % ngmax);
    unsigned c = counts[t] - (countsIncludeSelf ? 1u : 0u);
    c          = min(c, ngmax);
    out[tid]   = (k < c) ? list[nbListIndex(t / kGroupSize, ngmax, k, t % kGroupSize)] : 0u;
}

__global__ void resetScalarsKernel(StepScalars* s)
{
    s->minDtCourant   = INFINITY;
    s->maxDivv        = -INFINITY;
    s->totalNeighbors = 0;
    s->maxNc          = 0;
    s->numHIterated   = 0;
    s->errFlags       = 0;
    s->candTop        = 0;
    for (int q = 0; q < 8; ++q)
        s->work[q] = 0;
}

/* ---------------------------------------------- launchers ---------------------------------------------- */

void launchResetScalars(StepScalars* s, cudaStream_t stream) { resetScalarsKernel<<<1, 1, 0, stream>>>(s); }

void launchFindNeighbors(const double* x, const double* y, const double* z, const float* h, unsigned first,
                         unsigned last, const SphxBox& box, const SphxTreeView& tree, unsigned ngmax, unsigned* list,
                         unsigned* counts, StepScalars* scal, cudaStream_t stream)
{
    unsigned n = last - first;
    if (n == 0) return;
    unsigned numGroups = (n + kGroupSize - 1) / kGroupSize;
    unsigned blocks    = (numGroups + kWarpsPerBlock - 1) / kWarpsPerBlock;
    findNeighborsKernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(first, last, makeDevBox(box), tree, x, y, z, h,
                                                                   ngmax, list, counts, scal);
}

void launchFindNeighborsSphF64(const double* x, const double* y, const double* z, double* h, unsigned first, unsigned last,
                               const SphxBox& box, const SphxTreeView& tree, unsigned ng0, unsigned ngmax,
                               unsigned* list, unsigned* nc, StepScalars* scal, cudaStream_t stream)
{
    unsigned n = last - first;
    if (n == 0) return;
    unsigned numGroups = (n + kGroupSize - 1) / kGroupSize;
    unsigned blocks    = (numGroups + kWarpsPerBlock - 1) / kWarpsPerBlock;
    findNeighborsSphF64Kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(first, last, makeDevBox(box), tree, x, y, z,
                                                                         h, ng0, ngmax, list, nc, scal);
}

void launchExportNeighbors(unsigned numAssigned, unsigned ngmax, const unsigned* list, const unsigned* counts,
                           bool countsIncludeSelf, unsigned* out, cudaStream_t stream)
{
    size_t tot = size_t(numAssigned) * ngmax;
    if (tot == 0) return;
    unsigned blocks = unsigned((tot + 255) / 256);
    exportNeighborsKernel<<<blocks, 256, 0, stream>>>(numAssigned, ngmax, list, counts, countsIncludeSelf, out);
}

} // namespace sphx

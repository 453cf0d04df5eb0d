/*! @file
 * Block neighbour search of the hydro step: octree walk, candidate staging in shared memory, fp32 pair filter with an
 * exact fp64 decision for borderline pairs, coupled h-iteration, candidate compaction and the 16-bit neighbour list.
 *
 * Replaces (reference paths relative to /root/reference):
 *   cstone::findNeighbors            domain/include/cstone/findneighbors.hpp:77-147   (the pair predicate we must match)
 *   sph::findNeighborsSph            sph/include/sph/find_neighbors.hpp:11-44        (h-iteration)
 *   cstone::traverseNeighbors        domain/include/cstone/traversal/find_neighbors.cuh:182-489 (not followed)
 *
 * One CTA owns 128 SFC-consecutive targets (one per thread).
 *  1. bounding box of the targets' 2h-spheres; level-synchronous walk of the octree by all 128 threads collects the
 *     leaves that overlap it; the leaves are sorted into SFC order so that everything downstream is deterministic.
 *  2. the leaves' particles ("provisional candidates") are streamed through a 1024-entry shared-memory tile as
 *     float4 {position relative to the block origin (periodic shift applied), particle index}.
 *  3. every thread tests every staged particle of the leaves that touch one of its warp's spheres: the candidate is
 *     broadcast from shared memory, the distance is evaluated in fp32. Pairs whose fp32 distance lies within a proven
 *     error margin of the search radius are re-decided with the reference's own un-contracted fp64 predicate (same
 *     operation order, same PBC folding), so the neighbour sets are bit-exact. Hits go to a per-thread column of a
 *     shared-memory hit buffer.
 *  4. h-iteration as in the reference: if any target of the block has to change h, the block repeats the search.
 *  5. candidates used by at least one target are compacted (bit mask + popc prefix), appended to the global candidate
 *     array, and the hit columns are rewritten as 16-bit indices into that array.
 */
#include "sphx_block.cuh"
#include "sphx_kernels.h"

namespace sphx
{

constexpr int kSearchThreads  = kBlockTargets;
constexpr int kTileCap        = 768;   // staged particles per tile (incl. padding), multiple of 4
constexpr int kMaxLeaves      = 512;   // leaves overlapping one block's bounding box
constexpr int kMaxProvisional = 16384; // particles in those leaves (15-bit provisional index)
constexpr int kMaxTiles       = 48;
constexpr int kFrontierCap    = 1024; // nodes per tree level that overlap the block's bounding box

struct SearchShared
{
    // staged particles, SoA so that four consecutive x (y, z) are one 16-byte load and pair up for the packed
    // f32x2 arithmetic; every leaf is padded to a multiple of four entries with far-away dummies
    float          tileX[kTileCap], tileY[kTileCap], tileZ[kTileCap];
    unsigned       tileJ[kTileCap]; // particle index, ~0u for padding
    unsigned       usedBits[kMaxProvisional / 32];
    unsigned short wordPrefix[kMaxProvisional / 32];
    int            leafKey[kMaxLeaves];   // leaf index (sort key), later: particle count of the sorted leaf
    int            leafFirst[kMaxLeaves]; // first particle of the sorted leaf
    int            leafP0[kMaxLeaves];    // provisional index of that particle
    float          leafBox[kMaxLeaves * 6];
    int            tileFirstLeaf[kMaxTiles + 1];
    unsigned char  used8[kTileCap];
    double         red[6 * (kSearchThreads / 32)];
    int            count[2];
    int            nLeaf, nTiles, err, pEnd;
    int            maxLeafHalf[3];
    int            scan[kSearchThreads / 32];
    int            selfP[kSearchThreads]; // provisional index of each target's own particle
    unsigned       candBegin, numCand;
};

struct SearchArgs
{
    unsigned      first, last;
    DevBox        box;
    SphxTreeView  tree;
    const double *x, *y, *z;
    float*        h;
    unsigned*     nc;
    unsigned      ng0, ngmax, nkbMax;
    uint4*        list;
    float4*       cand;
    unsigned      candCapacity;
    BlockDesc*    blocks;
    StepScalars*  scal;
};

size_t searchSharedBytes(unsigned) { return sizeof(SearchShared); }

//! the reference's pair predicate (findneighbors.hpp:33-60,117,134), every fp64 operation rounded separately
__device__ __forceinline__ bool exactPair(const double* __restrict__ x, const double* __restrict__ y,
                                          const double* __restrict__ z, unsigned j, double xi, double yi, double zi,
                                          bool usePbc, const DevBox& box, float radiusSq)
{
    double dx = __dsub_rn(x[j], xi);
    double dy = __dsub_rn(y[j], yi);
    double dz = __dsub_rn(z[j], zi);
    if (usePbc)
    {
        dx = foldExact(dx, box.plx, box.ilx);
        dy = foldExact(dy, box.ply, box.ily);
        dz = foldExact(dz, box.plz, box.ilz);
    }
    return sumSqLeft(dx, dy, dz) < double(radiusSq);
}

//! candidate position relative to the block origin, moved to the periodic image next to the block unless fold mode
__device__ __forceinline__ float4 relativePosition(const SearchArgs& a, unsigned j, double ox, double oy, double oz,
                                                   bool foldMode)
{
    double X = a.x[j] - ox, Y = a.y[j] - oy, Z = a.z[j] - oz;
    if (!foldMode)
    {
        X -= a.box.plx * rint(X * a.box.ilx);
        Y -= a.box.ply * rint(Y * a.box.ily);
        Z -= a.box.plz * rint(Z * a.box.ilz);
    }
    return make_float4(float(X), float(Y), float(Z), __uint_as_float(j));
}

template<bool IterateH>
__global__ void __launch_bounds__(kSearchThreads, 6) blockSearchKernel(const __grid_constant__ SearchArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SearchShared& s = *reinterpret_cast<SearchShared*>(smemRaw);
    // scratch of the tree walk, aliased onto arrays that are written only later:
    int* frontier = reinterpret_cast<int*>(s.tileX);    // [2][kFrontierCap]: until the first tile is staged
    int* leafNode = reinterpret_cast<int*>(s.leafBox);  // leaves in traversal order: until they are ranked
    int* leafTmp  = reinterpret_cast<int*>(s.usedBits); // node index of the ranked leaves: until the leaf boxes exist
    static_assert(2 * kFrontierCap * sizeof(int) <= 3 * sizeof(s.tileX) + sizeof(s.tileJ), "frontier scratch");
    static_assert(kMaxLeaves * sizeof(int) <= sizeof(s.leafBox), "leaf scratch");
    static_assert(kMaxLeaves * sizeof(int) <= sizeof(s.usedBits) + sizeof(s.wordPrefix), "ranked-leaf scratch");
    static_assert(offsetof(SearchShared, tileY) == offsetof(SearchShared, tileX) + sizeof(s.tileX) &&
                      offsetof(SearchShared, wordPrefix) == offsetof(SearchShared, usedBits) + sizeof(s.usedBits),
                  "aliased arrays must be contiguous");



    constexpr int T    = kBlockTargets;
    const int     t    = threadIdx.x;
    const int     lane = t & 31, warp = t >> 5;

    /* Hit columns. Provisional hits of thread t = (warp g, lane) go to row k of the group's slice of the final list
     * storage, hitCol[k * 32] (u16, 64 bytes per row and group): L2-resident scratch instead of 38 KB of shared memory
     * per CTA, which more than doubles the number of resident CTAs. The list writer converts the slice in place, warp
     * synchronously: the final vector (kb, lane) covers exactly provisional rows 8 kb .. 8 kb + 7 (see there). */
    unsigned short* const hitCol =
        reinterpret_cast<unsigned short*>(a.list + (size_t(blockIdx.x) * kGroupsPerBlock + warp) * a.nkbMax * kGroupSize) +
        lane;
    const unsigned i     = a.first + blockIdx.x * T + t;
    const bool     valid = i < a.last;
    const unsigned il    = valid ? i : a.last - 1;
    const double   xi = a.x[il], yi = a.y[il], zi = a.z[il];
    float          hi        = a.h[il];
    bool           hChanged  = false;
    int            iteration = 0;
    unsigned       count     = 0;
    const unsigned ngmax     = a.ngmax;
    const DevBox&  box       = a.box;

    double ox = 0, oy = 0, oz = 0;
    bool   foldMode = false;
    int    L = 0, pEnd = 0;

    for (;;)
    {
        // ---------------------------------------------------------------------------------------------------------
        // per-target search parameters (findneighbors.hpp:93-100)
        const float  radiusSq = __fmul_rn(__fmul_rn(4.0f, hi), hi);
        const double ext      = __dmul_rn(2.0, double(hi));
        const bool   inside   = __dsub_rn(xi, ext) >= box.xmin && __dsub_rn(yi, ext) >= box.ymin &&
                            __dsub_rn(zi, ext) >= box.zmin && __dadd_rn(xi, ext) <= box.xmax &&
                            __dadd_rn(yi, ext) <= box.ymax && __dadd_rn(zi, ext) <= box.zmax;
        const bool usePbc = box.anyPbc && !inside;

        // ---------------------------------------------------------------------------------------------------------
        // bounding box of the block's search spheres (radius inflated: sqrt(fl(4 h^2)) <= 2h (1 + 2e-7))
        const double big = 1e300;
        const double r   = ext * (1.0 + 1e-6);
        double lo[3] = {valid ? xi - r : big, valid ? yi - r : big, valid ? zi - r : big};
        double up[3] = {valid ? xi + r : -big, valid ? yi + r : -big, valid ? zi + r : -big};
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = warpMin(lo[d]);
            up[d] = warpMax(up[d]);
        }
        if (lane == 0)
        {
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                s.red[warp * 6 + d]     = lo[d];
                s.red[warp * 6 + 3 + d] = up[d];
            }
        }
        s.selfP[t] = -0x40000000;
        if (t == 0)
        {
            s.count[0] = 1, s.count[1] = 0;
            s.nLeaf = 0, s.err = 0;
            s.maxLeafHalf[0] = s.maxLeafHalf[1] = s.maxLeafHalf[2] = 0;
            frontier[0] = 0; // root
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = s.red[d], up[d] = s.red[3 + d];
            for (int w = 1; w < kSearchThreads / 32; ++w)
            {
                lo[d] = fmin(lo[d], s.red[w * 6 + d]);
                up[d] = fmax(up[d], s.red[w * 6 + 3 + d]);
            }
        }
        const double bcx = 0.5 * (up[0] + lo[0]), bsx = 0.5 * (up[0] - lo[0]);
        const double bcy = 0.5 * (up[1] + lo[1]), bsy = 0.5 * (up[1] - lo[1]);
        const double bcz = 0.5 * (up[2] + lo[2]), bsz = 0.5 * (up[2] - lo[2]);
        ox = bcx, oy = bcy, oz = bcz;

        // ---------------------------------------------------------------------------------------------------------
        // level-synchronous octree walk: nodes overlapping the bounding box (PBC: minimum-image of the centre distance)
        {
            const double infl = 1.0 + 1e-9;
            int          cur  = 0;
            for (;;)
            {
                const int n = s.count[cur];
                if (n == 0) break;
                if (t == 0) s.count[cur ^ 1] = 0;
                __syncthreads();
                const int* fin  = frontier + cur * kFrontierCap;
                int*       fout = frontier + (cur ^ 1) * kFrontierCap;
                for (int q = t; q < n; q += T)
                {
                    const int node  = fin[q];
                    const int child = a.tree.childOffsets[node]; // unconditional: overlaps the centre/size loads
                    double    dx = a.tree.centers[3 * node] - bcx;
                    double    dy = a.tree.centers[3 * node + 1] - bcy;
                    double    dz = a.tree.centers[3 * node + 2] - bcz;
                    dx -= box.plx * rint(dx * box.ilx);
                    dy -= box.ply * rint(dy * box.ily);
                    dz -= box.plz * rint(dz * box.ilz);
                    const bool overlap = fabs(dx) <= (a.tree.sizes[3 * node] + bsx) * infl + 1e-300 &&
                                         fabs(dy) <= (a.tree.sizes[3 * node + 1] + bsy) * infl + 1e-300 &&
                                         fabs(dz) <= (a.tree.sizes[3 * node + 2] + bsz) * infl + 1e-300;
                    if (!overlap) continue;
                    if (child == 0)
                    {
                        const int li = atomicAdd(&s.nLeaf, 1);
                        if (li < kMaxLeaves) { leafNode[li] = node; }
                        else { s.err = 1; }
                    }
                    else
                    {
                        const int o = atomicAdd(&s.count[cur ^ 1], 8);
                        if (o + 8 <= kFrontierCap)
                        {
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                fout[o + c] = child + c;
                        }
                        else { s.err = 1; }
                    }
                }
                __syncthreads();
                if (s.err) break;
                cur ^= 1;
            }
        }
        __syncthreads();
        if (s.err) break;
        L = s.nLeaf;

        // ---------------------------------------------------------------------------------------------------------
        // sort the leaves into SFC order (rank by leaf index), fetch their particle ranges
        for (int q = t; q < L; q += T)
            s.leafKey[q] = a.tree.internalToLeaf[leafNode[q]];
        __syncthreads();
        for (int q = t; q < L; q += T)
        {
            const int key = s.leafKey[q];
            int       rk  = 0;
            for (int m = 0; m < L; ++m)
                rk += s.leafKey[m] < key;
            leafTmp[rk]   = leafNode[q];
            s.leafFirst[rk] = key;
        }
        __syncthreads();
        for (int q = t; q < L; q += T)
        {
            const int      node = leafTmp[q];
            const int      li   = s.leafFirst[q];
            const unsigned b = a.tree.layout[li], e = a.tree.layout[li + 1];
            s.leafFirst[q] = int(b);
            s.leafKey[q]   = int(e - b);
#pragma unroll
            for (int d = 0; d < 3; ++d)
                atomicMax(&s.maxLeafHalf[d], __float_as_int(__double2float_ru(a.tree.sizes[3 * node + d])));
        }
        __syncthreads();

        // shift mode is valid if every target-candidate separation stays below half a box length on periodic axes
        const float mlx = __int_as_float(s.maxLeafHalf[0]), mly = __int_as_float(s.maxLeafHalf[1]),
                    mlz = __int_as_float(s.maxLeafHalf[2]);
        const double ex = bsx + 2.0 * mlx, ey = bsy + 2.0 * mly, ez = bsz + 2.0 * mlz;
        {
            bool shiftOk = true;
            if (box.pbcX) shiftOk = shiftOk && ex < 0.25 * box.lx;
            if (box.pbcY) shiftOk = shiftOk && ey < 0.25 * box.ly;
            if (box.pbcZ) shiftOk = shiftOk && ez < 0.25 * box.lz;
            foldMode = box.anyPbc && !shiftOk;
        }
        const float E = __double2float_ru(fmax(ex, fmax(ey, ez))) * 1.0001f;

        // leaf boxes relative to the block origin (for the per-warp cull), provisional numbering, tiles
        for (int q = t; q < L; q += T)
        {
            const int node = leafTmp[q];
            double    cx = a.tree.centers[3 * node] - ox, cy = a.tree.centers[3 * node + 1] - oy,
                   cz = a.tree.centers[3 * node + 2] - oz;
            if (!foldMode)
            {
                cx -= box.plx * rint(cx * box.ilx);
                cy -= box.ply * rint(cy * box.ily);
                cz -= box.plz * rint(cz * box.ilz);
            }
            const float slop    = E * 2e-6f;
            s.leafBox[q * 6]     = float(cx);
            s.leafBox[q * 6 + 1] = float(cy);
            s.leafBox[q * 6 + 2] = float(cz);
            s.leafBox[q * 6 + 3] = __double2float_ru(a.tree.sizes[3 * node]) * 1.00001f + slop;
            s.leafBox[q * 6 + 4] = __double2float_ru(a.tree.sizes[3 * node + 1]) * 1.00001f + slop;
            s.leafBox[q * 6 + 5] = __double2float_ru(a.tree.sizes[3 * node + 2]) * 1.00001f + slop;
        }
        if (t == 0)
        {
            int P = 0, nT = 0, tileCount = 0, err = 0;
            s.tileFirstLeaf[0] = 0;
#pragma unroll 4
            for (int q = 0; q < L; ++q)
            {
                const int c = (s.leafKey[q] + 3) & ~3; // padded to a multiple of four
                if (c > kTileCap) err = 1;
                if (tileCount + c > kTileCap)
                {
                    P = (P + 31) & ~31;
                    ++nT;
                    if (nT >= kMaxTiles)
                    {
                        err = 1;
                        break;
                    }
                    s.tileFirstLeaf[nT] = q;
                    tileCount           = 0;
                }
                s.leafP0[q] = P;
                P += c;
                tileCount += c;
            }
            ++nT;
            s.tileFirstLeaf[nT] = L;
            s.nTiles            = nT;
            s.pEnd              = P;
            if (P > kMaxProvisional) err = 1;
            if (err) s.err = 1;
        }
        __syncthreads();
        if (s.err) break;
        pEnd             = s.pEnd;
        const int nTiles = s.nTiles;

        // ---------------------------------------------------------------------------------------------------------
        // fp32 filter thresholds. |d2_fp32 - d2_exact| <= 2^-24 (7 r E + 6.5 r^2) near the decision boundary
        // (E bounds every relative coordinate); pairs inside the margin are decided by the exact predicate.
        const float  tx = float(xi - ox), ty = float(yi - oy), tz = float(zi - oz);
        const float2 ntx = make_float2(-tx, -tx), nty = make_float2(-ty, -ty), ntz = make_float2(-tz, -tz);
        float        r2lo, r2hi;
        {
            const float rr     = 2.0f * hi * 1.000001f;
            const float margin = 5.9604645e-8f * 16.0f * (rr * E + rr * rr);
            r2lo               = foldMode ? -1.0f : radiusSq - margin;
            r2hi               = foldMode ? 3.0e38f : radiusSq + margin;
            if (!valid) r2lo = r2hi = -1.0f;
        }

        unsigned       slot    = 0;                     // next free element of this thread's hit column (row * 32)
        const unsigned slotEnd = (ngmax + 1) * kGroupSize; // the column also receives the target's own particle
        const unsigned iBlock0 = a.first + blockIdx.x * T;
        for (int tIdx = 0; tIdx < nTiles; ++tIdx)
        {
            const int lb = s.tileFirstLeaf[tIdx], le = s.tileFirstLeaf[tIdx + 1];
            if (lb == le) continue;
            const int tileBase = s.leafP0[lb];
            const int tileN    = s.leafP0[le - 1] + ((s.leafKey[le - 1] + 3) & ~3) - tileBase;

            // stage
            {
                int l = lb;
                for (int p = t; p < tileN; p += T)
                {
                    const int P = tileBase + p;
                    while (P >= s.leafP0[l] + ((s.leafKey[l] + 3) & ~3))
                        ++l;
                    const int off = P - s.leafP0[l];
                    float4    rp  = make_float4(1e18f, 1e18f, 1e18f, __uint_as_float(~0u)); // padding: never a hit
                    if (off < s.leafKey[l])
                    {
                        const unsigned j = unsigned(s.leafFirst[l] + off);
                        rp               = relativePosition(a, j, ox, oy, oz, foldMode);
                        if (j - iBlock0 < unsigned(T)) s.selfP[j - iBlock0] = P;
                    }
                    s.tileX[p] = rp.x, s.tileY[p] = rp.y, s.tileZ[p] = rp.z;
                    s.tileJ[p] = __float_as_uint(rp.w);
                    s.used8[p] = 0;
                }
            }
            __syncthreads();

            const unsigned slotTile = slot;
            for (int l = lb; l < le; ++l)
            {
                // does any sphere of this warp touch the leaf?
                const float* bx  = &s.leafBox[l * 6];
                const float  ddx = fmaxf(fabsf(bx[0] - tx) - bx[3], 0.0f);
                const float  ddy = fmaxf(fabsf(bx[1] - ty) - bx[4], 0.0f);
                const float  ddz = fmaxf(fabsf(bx[2] - tz) - bx[5], 0.0f);
                const bool   touch = ddx * ddx + ddy * ddy + ddz * ddz <= r2hi;
                if (!__any_sync(kFullMask, touch)) continue;

                const int pb = s.leafP0[l] - tileBase;
                const int pe = pb + s.leafKey[l];
                // four staged particles per step in straight-line code; distances in packed f32x2 arithmetic
                for (int p = pb; p < pe; p += 4)
                {
                    const float4 X = *reinterpret_cast<const float4*>(&s.tileX[p]);
                    const float4 Y = *reinterpret_cast<const float4*>(&s.tileY[p]);
                    const float4 Z = *reinterpret_cast<const float4*>(&s.tileZ[p]);
                    const float2 dx0 = __fadd2_rn(make_float2(X.x, X.y), ntx), dx1 = __fadd2_rn(make_float2(X.z, X.w), ntx);
                    const float2 dy0 = __fadd2_rn(make_float2(Y.x, Y.y), nty), dy1 = __fadd2_rn(make_float2(Y.z, Y.w), nty);
                    const float2 dz0 = __fadd2_rn(make_float2(Z.x, Z.y), ntz), dz1 = __fadd2_rn(make_float2(Z.z, Z.w), ntz);
                    const float2 s0 = __ffma2_rn(dz0, dz0, __ffma2_rn(dy0, dy0, __fmul2_rn(dx0, dx0)));
                    const float2 s1 = __ffma2_rn(dz1, dz1, __ffma2_rn(dy1, dy1, __fmul2_rn(dx1, dx1)));
                    const float  d2[4] = {s0.x, s0.y, s1.x, s1.y};
                    // a pair is a sure hit below r2lo and a sure miss from r2hi on; in between (or in fold mode) the
                    // reference's exact fp64 predicate decides. The target's own particle (d2 = 0) is recorded like
                    // any other hit and dropped when the list is written.
                    bool amb = false;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        amb = amb | (!(d2[u] < r2lo) & (d2[u] < r2hi));
                    if (__any_sync(kFullMask, amb))
                    {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                        {
                            bool h = d2[u] < r2hi;
                            if (h && !(d2[u] < r2lo))
                            {
                                const unsigned j = s.tileJ[p + u];
                                h = j != ~0u && exactPair(a.x, a.y, a.z, j, xi, yi, zi, usePbc, box, radiusSq);
                            }
                            if (h)
                            {
                                if (slot < slotEnd) hitCol[slot] = (unsigned short)(tileBase + p + u);
                                slot += kGroupSize;
                            }
                        }
                        continue;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                    {
                        const bool hit = d2[u] < r2lo;
                        if (hit && slot < slotEnd) hitCol[slot] = (unsigned short)(tileBase + p + u);
                        slot += hit ? kGroupSize : 0u;
                    }
                }
            }

            // mark the staged particles this thread kept
            {
                const unsigned kEnd = min(slot, slotEnd);
                for (unsigned k = slotTile; k < kEnd; k += kGroupSize)
                    s.used8[hitCol[k] - tileBase] = 1;
            }
            __syncthreads();
            for (int base = warp * 32; base < tileN; base += kSearchThreads)
            {
                const bool     bit = (base + lane < tileN) && s.used8[base + lane];
                const unsigned m   = __ballot_sync(kFullMask, bit);
                if (lane == 0) s.usedBits[(tileBase + base) >> 5] = m;
            }
            __syncthreads();
        }

        // number of hits without the target's own particle; keeps counting beyond ngmax, as the reference does
        count = slot / kGroupSize;
        if (valid && s.selfP[t] >= 0 && count > 0) --count;

        // ---------------------------------------------------------------------------------------------------------
        // sph/find_neighbors.hpp:17-36: while ((ngmin > nc || nc - 1 > ngmax) && iteration++ < 10)
        if (!IterateH) break;
        const unsigned ncSph  = 1 + count;
        const unsigned ngmin  = a.ng0 / 4;
        const bool     repeat = valid && (ngmin > ncSph || (ncSph - 1) > ngmax) && iteration < 10;
        if (!__syncthreads_or(repeat)) break;
        if (repeat)
        {
            iteration++;
            hi       = updateH(a.ng0, ncSph, hi);
            hChanged = true;
        }
    }

    // -------------------------------------------------------------------------------------------------------------
    const unsigned ncSph = 1 + count;
    BlockDesc      desc;
    desc.ox = ox, desc.oy = oy, desc.oz = oz;
    desc.flags = foldMode ? kBlockFold : 0u;
    desc.pad   = 0;

    if (s.err)
    {
        // traversal capacity exceeded: flag it, leave an empty block behind
        if (t == 0)
        {
            atomicOr(&a.scal->errFlags, kErrTraversal);
            desc.candBegin = 0, desc.numCand = 0;
            a.blocks[blockIdx.x] = desc;
        }
        if (valid) a.nc[i] = 1;
        return;
    }

    // popc prefix over the used-bit words -> compact candidate numbering
    {
        const int nW  = (pEnd + 31) >> 5;
        const int w0  = t * 8;
        int       loc[8];
        int       sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
            loc[q] = sum;
            sum += (w0 + q < nW) ? __popc(s.usedBits[w0 + q]) : 0;
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s.scan[warp] = incl;
        __syncthreads();
        int base = 0, total = 0;
        for (int w = 0; w < kSearchThreads / 32; ++w)
        {
            if (w < warp) base += s.scan[w];
            total += s.scan[w];
        }
        base += incl - sum;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (w0 + q < nW) s.wordPrefix[w0 + q] = (unsigned short)(base + loc[q]);
        if (t == 0)
        {
            unsigned begin = atomicAdd(&a.scal->candTop, unsigned(total));
            unsigned num   = unsigned(total);
            if (size_t(begin) + num > a.candCapacity)
            {
                atomicOr(&a.scal->errFlags, kErrCandSpace);
                begin = 0, num = 0;
            }
            s.candBegin = begin, s.numCand = num;
            desc.candBegin = begin, desc.numCand = num;
            a.blocks[blockIdx.x] = desc;
        }
        __syncthreads();
    }
    const unsigned candBegin = s.candBegin;
    const bool     haveSpace = s.numCand > 0 || pEnd == 0;

    // compacted candidate records
    if (haveSpace)
    {
        int l = 0;
        for (int P = t; P < pEnd; P += T)
        {
            while (l < L && P >= s.leafP0[l] + ((s.leafKey[l] + 3) & ~3))
                ++l;
            if (l >= L) break;
            if (P < s.leafP0[l]) continue; // alignment gap between tiles
            const unsigned w = s.usedBits[P >> 5];
            if (!((w >> (P & 31)) & 1u)) continue;
            const unsigned c = s.wordPrefix[P >> 5] + __popc(w & ((1u << (P & 31)) - 1u));
            const unsigned j = unsigned(s.leafFirst[l] + (P - s.leafP0[l]));
            a.cand[size_t(candBegin) + c] = relativePosition(a, j, ox, oy, oz, foldMode);
        }
    }

    // neighbour list: 16-bit candidate indices, 8 per vector, lane-interleaved per group of 32 targets
    {
        /* In-place, warp-synchronous: provisional row k of the group occupies bytes [64 k, 64 k + 64) of its slice,
         * the final vector (kb, lane) bytes [512 kb + 16 lane, + 16), i.e. a piece of provisional row 8 kb + lane / 4.
         * All lanes first read their rows 8 kb .. 8 kb + 8 (one extra when the own particle has been skipped), then
         * all lanes write their vector kb, which only destroys rows 8 kb .. 8 kb + 7 that every lane has consumed. */
        const unsigned kc     = haveSpace ? min(count, ngmax) : 0u;
        const unsigned nkb    = (kc + 7) / 8;
        const unsigned nkbW   = warpMaxU(nkb);
        const unsigned selfPu = unsigned(s.selfP[t]); // provisional index of the target itself
        unsigned       rk     = 0;                    // read cursor (row) in the hit column
        uint4* lp = a.list + (size_t(blockIdx.x) * kGroupsPerBlock + warp) * a.nkbMax * kGroupSize + lane;
        for (unsigned kb = 0; kb < nkbW; ++kb)
        {
            unsigned wds[4] = {0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < 8; ++q)
            {
                const unsigned k = kb * 8 + q;
                unsigned       e = 0;
                if (k < kc)
                {
                    unsigned P = hitCol[rk * kGroupSize];
                    if (P == selfPu) P = hitCol[++rk * kGroupSize];
                    ++rk;
                    const unsigned w = s.usedBits[P >> 5];
                    e                = s.wordPrefix[P >> 5] + __popc(w & ((1u << (P & 31)) - 1u));
                }
                wds[q >> 1] |= e << (16 * (q & 1));
            }
            __syncwarp();
            if (kb < nkb) lp[size_t(kb) * kGroupSize] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            __syncwarp();
        }
    }

    // outputs and statistics (conserved_quantities.hpp:146-157 sums nc)
    if (IterateH && valid)
    {
        const unsigned ngmin = a.ng0 / 4;
        if ((ngmin > ncSph || (ncSph - 1) > ngmax) && iteration >= 10) atomicOr(&a.scal->errFlags, kErrHConv);
        if ((ncSph - 1) > ngmax) atomicOr(&a.scal->errFlags, kErrNgmax);
    }
    unsigned ncv = valid ? ncSph : 0, ncSum = ncv, ncMax = ncv, nIter = (valid && hChanged) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        ncSum += __shfl_xor_sync(kFullMask, ncSum, o);
        ncMax = max(ncMax, __shfl_xor_sync(kFullMask, ncMax, o));
        nIter += __shfl_xor_sync(kFullMask, nIter, o);
    }
    if (lane == 0)
    {
        atomicAdd(&a.scal->totalNeighbors, (unsigned long long)ncSum);
        atomicMax(&a.scal->maxNc, ncMax);
        if (nIter) atomicAdd(&a.scal->numHIterated, nIter);
    }
    if (valid)
    {
        if (hChanged) a.h[i] = hi;
        a.nc[i] = ncSph;
    }
}

//! block list -> reference CPU layout neighbors[(i-first)*ngmax + k] with particle indices (particles_data.hpp:250)
__global__ void exportBlockNeighborsKernel(unsigned numAssigned, unsigned ngmax, unsigned nkbMax,
                                           const uint4* __restrict__ list, const float4* __restrict__ cand,
                                           const BlockDesc* __restrict__ blocks, const unsigned* __restrict__ nc,
                                           unsigned* __restrict__ out)
{
    size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t tot = size_t(numAssigned) * ngmax;
    if (tid >= tot) return;
    const unsigned tt = unsigned(tid / ngmax), k = unsigned(tid % ngmax);
    const unsigned c  = min(nc[tt] - 1u, ngmax);
    unsigned       j  = 0;
    if (k < c)
    {
        const unsigned g = tt / kGroupSize, lane = tt % kGroupSize;
        const uint4    v = list[(size_t(g) * nkbMax + k / 8) * kGroupSize + lane];
        const unsigned wds[4] = {v.x, v.y, v.z, v.w};
        const unsigned e      = (wds[(k & 7) >> 1] >> (16 * (k & 1))) & 0xffffu;
        j = __float_as_uint(cand[size_t(blocks[tt / kBlockTargets].candBegin) + e].w);
    }
    out[tid] = j;
}

/* ---------------------------------------------- launchers ---------------------------------------------- */

static cudaError_t configureSearch(unsigned ngmax)
{
    static size_t configured = 0;
    size_t        bytes      = searchSharedBytes(ngmax);
    if (bytes <= 48 * 1024) return cudaSuccess;
    if (bytes > configured)
    {
        cudaError_t e = cudaFuncSetAttribute(blockSearchKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(bytes));
        if (e != cudaSuccess) return e;
        configured = bytes;
    }
    return cudaSuccess;
}

cudaError_t launchBlockSearch(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t stream)
{
    unsigned n = unsigned(a.last - a.first);
    if (n == 0) return cudaSuccess;
    cudaError_t e = configureSearch(a.p.ngmax);
    if (e != cudaSuccess) return e;
    char*      base = static_cast<char*>(a.workspace);
    SearchArgs s;
    s.first = unsigned(a.first), s.last = unsigned(a.last);
    s.box  = makeDevBox(a.box);
    s.tree = a.tree;
    s.x = a.f.x, s.y = a.f.y, s.z = a.f.z, s.h = a.f.h, s.nc = a.f.nc;
    s.ng0 = a.p.ng0, s.ngmax = a.p.ngmax, s.nkbMax = w.nkbMax;
    s.list         = reinterpret_cast<uint4*>(base + w.listOff);
    s.cand         = reinterpret_cast<float4*>(base + w.candOff);
    s.candCapacity = unsigned(w.candCapacity > 0xffffffffull ? 0xffffffffull : w.candCapacity);
    s.blocks       = reinterpret_cast<BlockDesc*>(base + w.blocksOff);
    s.scal         = reinterpret_cast<StepScalars*>(base + w.scalOff);
    blockSearchKernel<true><<<w.numBlocks, kSearchThreads, searchSharedBytes(a.p.ngmax), stream>>>(s);
    return cudaGetLastError();
}

void launchExportBlockNeighbors(const SphxStepArgs& a, const WorkspaceLayout& w, unsigned* out, cudaStream_t stream)
{
    unsigned n   = unsigned(a.last - a.first);
    size_t   tot = size_t(n) * a.p.ngmax;
    if (tot == 0) return;
    char* base = static_cast<char*>(a.workspace);
    exportBlockNeighborsKernel<<<unsigned((tot + 255) / 256), 256, 0, stream>>>(
        n, a.p.ngmax, w.nkbMax, reinterpret_cast<const uint4*>(base + w.listOff),
        reinterpret_cast<const float4*>(base + w.candOff), reinterpret_cast<const BlockDesc*>(base + w.blocksOff),
        a.f.nc + a.first, out);
}

} // namespace sphx

/*! @file
 * Block neighbour search of the hydro step: octree walk, candidate staging in shared memory, quad cull, fp32 pair filter
 * with an exact fp64 decision for borderline pairs, coupled h-iteration, candidate compaction and the 16-bit neighbour
 * list.
 *
 * Replaces (reference paths relative to /root/reference):
 *   cstone::findNeighbors            domain/include/cstone/findneighbors.hpp:77-147   (the pair predicate we must match)
 *   sph::findNeighborsSph            sph/include/sph/find_neighbors.hpp:11-44        (h-iteration)
 *   cstone::traverseNeighbors        domain/include/cstone/traversal/find_neighbors.cuh:182-489 (not followed)
 *
 * A CTA owns 128 SFC-consecutive targets at a time (one per thread, persistent CTAs, eight per SM).
 *  1. bounding box of the targets' 2h-spheres; level-synchronous walk of the octree by all 128 threads collects the
 *     leaves that overlap it; the leaves are ranked by leaf index (deterministic) and stored interleaved (ranks = g
 *     mod 8 together), so that a tile samples the whole candidate region and the warps' work per tile is balanced.
 *  2. provisional numbering by a block-wide scan: every leaf starts a new 32-slot word, leaf q is staged in tile
 *     P / B at offset P % B (P: running sum of the padded particle counts).
 *  3. the leaves' particles are streamed through a 768-entry shared-memory tile (SoA x | y | z, fp32, relative to the
 *     block origin, periodic shift applied); four consecutive staged particles form a quad whose bounding box is
 *     reduced with two shuffles while staging.
 *  4. per warp, one quad per lane: quads further from the bounding box of the warp's 32 targets than its largest
 *     search radius are dropped (ballot -> bit mask); the warp walks the remaining quads, three LDS.128 broadcasts and
 *     packed f32x2 arithmetic give four squared distances per lane. Pairs whose fp32 distance lies within a proven
 *     error margin of the search radius are re-decided with the reference's own un-contracted fp64 predicate (same
 *     operation order, same PBC folding), so the neighbour sets are bit-exact. Hits are bit masks: {mask, word} entries
 *     in a per-thread column of an L2-resident scratch.
 *  5. h-iteration as in the reference: if any target of the block has to change h, the block repeats the search.
 *  6. candidates used by at least one target are compacted (bit mask + popc prefix) and appended to the global
 *     candidate array; the mask columns are copied back to shared memory in cp.async batches and decoded into 16-bit
 *     indices into that array, eight per uint4 of the lane-interleaved list.
 *  Blocks that do not fit the per-block tables go on an overflow list and are redone with 8x larger tables (CapsBig).
 */
#include "sphx_block.cuh"
#include "sphx_kernels.h"

#include <algorithm>

namespace sphx
{

constexpr int kSearchThreads = kBlockTargets;
constexpr int kSearchWarps   = kSearchThreads / 32;
#ifndef SPHX_TILE_CAP
#define SPHX_TILE_CAP 768
#endif
constexpr int kTileCap       = SPHX_TILE_CAP; // staged particles per tile (incl. padding), multiple of 128
/*! Per-block table sizes. The standard set fits eight CTAs per SM. A block whose reach holds more leaves (cavities and
 *  shells of an evolved blast wave: large search spheres next to finely resolved regions) is pushed on an overflow list
 *  and redone by the same code instantiated with the big set, one CTA per SM, with scratch arrays of its own instead of
 *  the aliased ones. */
struct CapsStd
{
    static constexpr int  kMaxLeaves   = 512;  // leaves overlapping one block's bounding box
    static constexpr int  kMaxWords    = 768;  // provisional numbering: 32 slots per word, every leaf starts a new word
    static constexpr int  kMaxTiles    = 48;
    static constexpr int  kFrontierCap = 1024; // nodes per tree level that overlap the block's bounding box
    static constexpr bool kOwnScratch  = false;
};
struct CapsBig
{
    static constexpr int  kMaxLeaves   = 4096;
    static constexpr int  kMaxWords    = 6144;
    static constexpr int  kMaxTiles    = 250;
    static constexpr int  kFrontierCap = 4096;
    static constexpr bool kOwnScratch  = true;
};
constexpr int kOverflowCount = 6; // StepScalars::work slots: blocks on the overflow list, work counter of the big kernel
constexpr int kBigWork       = 7;
#ifndef SPHX_LEAF_CLASSES
#define SPHX_LEAF_CLASSES 8
#endif
constexpr int      kLeafClasses = SPHX_LEAF_CLASSES; // leaf interleave of the tiles (see the ranking step)
#ifndef SPHX_DECODE_BATCH
#define SPHX_DECODE_BATCH 16
#endif
constexpr unsigned kDecodeBatch = SPHX_DECODE_BATCH; // list decode: hit-mask entries per lane staged in shared memory at a time
#ifndef SPHX_QUAD_UNROLL
#define SPHX_QUAD_UNROLL 1 // quad walk / list decode: 2 = two iterations per taken back-edge (measured: 1 % slower)
#endif
#ifndef SPHX_DECODE_UNROLL
#define SPHX_DECODE_UNROLL 1
#endif
constexpr int kQuadUnroll = SPHX_QUAD_UNROLL, kDecodeUnroll = SPHX_DECODE_UNROLL;
constexpr int kTileQuads      = kTileCap / 4;
constexpr int kKeepWords      = kTileQuads / 32;
static_assert(kTileQuads % 32 == 0, "quad cull: whole ballots per tile");

template<class C>
struct SearchShared
{
    // staged particles, SoA so that four consecutive x (y, z) are one 16-byte load and pair up for the packed
    // f32x2 arithmetic; every leaf is padded to a multiple of four entries with far-away dummies
    float          tileX[kTileCap], tileY[kTileCap], tileZ[kTileCap];
    // boxes of the staged quads (four SFC-consecutive particles of a leaf) of the current tile, relative to the block
    // origin: [lo x | lo y | lo z | hi x | hi y | hi z][kTileQuads]. Aliased by scratch of the tree walk.
    float          quadBox[6 * kTileQuads];
    int            leafKey[C::kMaxLeaves];   // leaf index (sort key), later: particle count of the sorted leaf
    unsigned short leafTile[C::kMaxLeaves];  // offset of the leaf's particles in its tile
    // provisional word of the quad << 5 | bit position of the quad's nibble in the word's hit mask (28 - 4 quad)
    using QuadMeta = std::conditional_t<(C::kMaxWords * 32 > 65536), unsigned, unsigned short>;
    QuadMeta       quadMeta[kTileQuads];
    unsigned       keep[kSearchWarps][kKeepWords]; // per warp: quads of the tile within reach of its targets
    // (everything above is dead once the pair tests are done: the list decode stages the hit-mask columns there)
    unsigned       usedBits[C::kMaxWords];   // union over the block's targets of the hit masks, per provisional word
    unsigned short wordPrefix[C::kMaxWords]; // number of used provisional slots before each word
    unsigned short wordLeaf[C::kMaxWords];   // (sorted) leaf that owns the word
    int            leafFirst[C::kMaxLeaves]; // first particle of the sorted leaf
    unsigned short leafW0[C::kMaxLeaves];    // first provisional word of the leaf
    int            tileFirstLeaf[C::kMaxTiles + 1];
    // scratch of the tree walk when it does not fit the aliased arrays (big set)
    int            ownFrontier[C::kOwnScratch ? 2 * C::kFrontierCap : 1];
    int            ownLeafNode[C::kOwnScratch ? C::kMaxLeaves : 1];
    float4         ownTgtSph[C::kOwnScratch ? kBlockTargets : 1];
    double         red[6 * kSearchWarps];
    int            count[2];
    int            nLeaf, nTiles, err, wEnd;
    int            maxLeafHalf[3];
    int            maxLeafCount; // most particles in one of the block's leaves
    int            scan[kSearchWarps];
    int            selfP[kSearchThreads]; // provisional slot of each target's own particle
    double         origin[3];             // block origin (registers are short in the quad walk: reloaded where needed)
    unsigned       candBegin, numCand, nextBlock;
};

struct SearchArgs
{
    unsigned      first, last;
    DevBox        box;
    SphxTreeView  tree;
    const double *x, *y, *z;
    float*        h;
    unsigned*     nc;
    unsigned      ng0, ngmax, nkbMax, numBlocks, maskRows;
    uint4*        list;
    float4*       cand;
    unsigned      candCapacity;
    uint2*        maskScratch; // per resident CTA: kBlockTargets columns of maskRows {hit mask, provisional word}
    BlockDesc*    blocks;
    unsigned*     overflowList; // blocks the standard tables could not hold (count in scal->work[kOverflowCount])
    StepScalars*  scal;
};

template<class C>
constexpr size_t searchSharedBytes() { return sizeof(SearchShared<C>); }

//! load that the compiler neither hoists nor merges with an earlier one (values needed on rare paths only are
//! re-read there instead of occupying registers across the quad walk)
__device__ __forceinline__ double reload(const double* p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

/* Shared-memory accesses of the quad walk through an explicit 32-bit shared-space address. Left to itself, ptxas
 * re-derives the shared window base (S2R SR_CgaCtaId, MOV, LEA) at the top of every quad iteration; the base obtained
 * once through a volatile cvta cannot be rematerialised and stays in a register. */
#ifndef SPHX_SEARCH_SADDR
#define SPHX_SEARCH_SADDR 1
#endif
__device__ __forceinline__ unsigned sharedBase(const void* p)
{
    unsigned r;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldsF4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned ldsU16(unsigned addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned ldsU32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 ldsU32x2(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
//! index of the most significant set bit (x != 0)
__device__ __forceinline__ unsigned highestBit(unsigned x)
{
    unsigned p;
    asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(x));
    return p;
}
__device__ __forceinline__ void redOrShared(unsigned addr, unsigned v)
{
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

//! the reference's pair predicate (findneighbors.hpp:33-60,117,134), every fp64 operation rounded separately
#ifndef SPHX_EXACT_INLINE
#define SPHX_EXACT_INLINE 1 // 1: one inlined copy of the exact predicate in the quad loop instead of a call
#endif
#if SPHX_EXACT_INLINE
#define SPHX_EXACT_ATTR __forceinline__
#else
#define SPHX_EXACT_ATTR __noinline__
#endif
__device__ SPHX_EXACT_ATTR bool exactPair(const double* __restrict__ x, const double* __restrict__ y,
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

/*! @brief search of one block of 128 targets (one per thread)
 *
 * Hits are recorded as bit masks: for every 32 staged particles of a leaf ("provisional word") a thread shifts one bit
 * per pair test into a register; when the word is complete, a thread with at least one hit appends {mask, word} to its
 * private column of the L2-resident scratch and the warp ORs its masks into the block's used-slot mask. Only after the
 * h-iteration has converged are the columns decoded into the 16-bit neighbour list.
 *
 * @param blk      block index
 * @param maskCol  this thread's column of the hit-mask scratch (entry r at maskCol[32 r])
 */
template<bool IterateH, class C>
__device__ __forceinline__ void searchBlock(const SearchArgs& a, SearchShared<C>& s, const unsigned blk,
                                            uint2* __restrict__ maskCol)
{
    using Shared = SearchShared<C>;
    constexpr int kMaxLeaves = C::kMaxLeaves, kMaxWords = C::kMaxWords, kMaxTiles = C::kMaxTiles,
                  kFrontierCap = C::kFrontierCap;
    constexpr int kWordsPerThread = kMaxWords / kSearchThreads, kLeavesPerThread = kMaxLeaves / kSearchThreads;
    static_assert(kMaxWords % kSearchThreads == 0 && kMaxLeaves % kSearchThreads == 0, "scans: whole items per thread");
    // scratch of the tree walk, aliased onto arrays that are written only later (standard set):
    int* frontier = C::kOwnScratch ? s.ownFrontier : reinterpret_cast<int*>(s.tileX); // [2][kFrontierCap]: until the first tile is staged
    int* leafNode = C::kOwnScratch ? s.ownLeafNode : reinterpret_cast<int*>(s.quadBox); // leaves in traversal order: until they are ranked
    int* leafTmp  = reinterpret_cast<int*>(s.usedBits); // node index of the ranked leaves: until the leaf boxes exist
    static_assert(C::kOwnScratch || 2 * kFrontierCap * sizeof(int) <= 3 * sizeof(s.tileX), "frontier scratch");
    static_assert(C::kOwnScratch || kMaxLeaves * sizeof(int) <= sizeof(s.quadBox), "leaf scratch");
    static_assert(kMaxLeaves * sizeof(int) <= sizeof(s.usedBits), "ranked-leaf scratch");
    static_assert(offsetof(Shared, tileY) == offsetof(Shared, tileX) + sizeof(s.tileX) &&
                      offsetof(Shared, tileZ) == offsetof(Shared, tileY) + sizeof(s.tileY),
                  "aliased arrays must be contiguous");

    constexpr int T    = kBlockTargets;
    const int     t    = threadIdx.x;
    const int     lane = t & 31, warp = t >> 5;
#if SPHX_SEARCH_SADDR
    const unsigned sb = sharedBase(&s);
#endif

    const unsigned iBlock0 = a.first + blk * T;
    const unsigned i       = iBlock0 + t;
    const bool     valid   = i < a.last;
    const unsigned il      = valid ? i : a.last - 1;
    float          hi        = a.h[il];
    bool           hChanged  = false;
    int            iteration = 0;
    unsigned       count     = 0;
    unsigned       numEnt    = 0; // entries in this thread's hit-mask column
    const unsigned ngmax     = a.ngmax;
    const DevBox&  box       = a.box;

    bool   foldMode = false;
    int    L = 0;
    // Fallback for blocks whose targets are NOT compact in space (the SFC leaves the particle distribution and
    // re-enters it elsewhere, e.g. at the surface of the Noh sphere): their bounding box covers the gap and overlaps
    // more leaves than the shared-memory tables hold. Such a block repeats the walk in "precise" mode: a node is kept
    // only if the search sphere of at least one target touches its box. This plays the role of the reference's group
    // splits (computeGroupSplits, traversal/groups_gpu.cu:106-135), which keep a group's bounding box small.
    bool    precise = false;
    float4* tgtSph = C::kOwnScratch ? s.ownTgtSph : reinterpret_cast<float4*>(s.quadBox + kMaxLeaves); // [T], walk only
    static_assert(C::kOwnScratch || (kMaxLeaves + 4 * kBlockTargets) * sizeof(float) <= sizeof(s.quadBox),
                  "target sphere scratch");

    for (;;)
    {
        // ---------------------------------------------------------------------------------------------------------
        // per-target search parameters (findneighbors.hpp:93-100). The fp64 position is needed up to the fp32 filter
        // set-up only (and by the rare exact decisions, which re-read it): not kept in registers across the quad walk
        const double xi = reload(a.x + il), yi = reload(a.y + il), zi = reload(a.z + il);
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
            s.maxLeafCount = 0;
            frontier[0] = 0; // root
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = s.red[d], up[d] = s.red[3 + d];
            for (int w = 1; w < kSearchWarps; ++w)
            {
                lo[d] = fmin(lo[d], s.red[w * 6 + d]);
                up[d] = fmax(up[d], s.red[w * 6 + 3 + d]);
            }
        }
        const double bcx = 0.5 * (up[0] + lo[0]), bsx = 0.5 * (up[0] - lo[0]);
        const double bcy = 0.5 * (up[1] + lo[1]), bsy = 0.5 * (up[1] - lo[1]);
        const double bcz = 0.5 * (up[2] + lo[2]), bsz = 0.5 * (up[2] - lo[2]);
        if (t == 0) s.origin[0] = bcx, s.origin[1] = bcy, s.origin[2] = bcz; // read after the barriers of the walk
        if (precise)
        {
            const float rr = float(r) * 1.0001f;
            tgtSph[t]      = make_float4(float(xi - bcx), float(yi - bcy), float(zi - bcz), valid ? rr * rr : -1.0f);
        }

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
                    if (precise)
                    {
                        // any target sphere within reach of the node box? fp32 relative to the box centre, inflated
                        const float slop = float(fmax(bsx, fmax(bsy, bsz))) * 8e-6f;
                        const float ncx = float(dx), ncy = float(dy), ncz = float(dz);
                        const float hx = __double2float_ru(a.tree.sizes[3 * node]) * 1.00001f + slop;
                        const float hy = __double2float_ru(a.tree.sizes[3 * node + 1]) * 1.00001f + slop;
                        const float hz = __double2float_ru(a.tree.sizes[3 * node + 2]) * 1.00001f + slop;
                        const float plx = float(box.plx), ply = float(box.ply), plz = float(box.plz);
                        const float ilx = float(box.ilx), ily = float(box.ily), ilz = float(box.ilz);
                        bool        touch = false;
                        for (int m = 0; m < T && !touch; ++m)
                        {
                            const float4 tg = tgtSph[m];
                            float        ex = ncx - tg.x, ey = ncy - tg.y, ez = ncz - tg.z;
                            ex -= plx * rintf(ex * ilx);
                            ey -= ply * rintf(ey * ily);
                            ez -= plz * rintf(ez * ilz);
                            const float qx = fmaxf(fabsf(ex) - hx, 0.0f), qy = fmaxf(fabsf(ey) - hy, 0.0f),
                                        qz = fmaxf(fabsf(ez) - hz, 0.0f);
                            touch = qx * qx + qy * qy + qz * qz <= tg.w;
                        }
                        if (!touch) continue;
                    }
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
        if (s.err)
        {
            if (precise) break;
            precise = true; // uniform: s.err is read after a barrier
            __syncthreads();
            continue;
        }
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
            // Interleave: the leaves with SFC rank = g (mod kLeafClasses) are stored together, class after class. A
            // tile (a contiguous piece of this order) then samples the whole candidate region instead of one corner of
            // it, so that every warp finds about the same share of its quads in every tile (the warps meet at a barrier
            // per tile). Deterministic; candidates and list entries follow this order.
            const int g  = rk % kLeafClasses;
            const int ps = g * (L / kLeafClasses) + min(g, L % kLeafClasses) + rk / kLeafClasses;
            leafTmp[ps]     = leafNode[q];
            s.leafFirst[ps] = key;
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
            atomicMax(&s.maxLeafCount, int(e - b));
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

        // provisional numbering (every leaf starts a new 32-slot word) and tiles. With P = running sum of the leaves'
        // particle counts (each padded to a multiple of four), leaf q is staged in tile P[q] / B at offset P[q] % B:
        // B = kTileCap - largest padded count, so no leaf crosses the end of the tile buffer and no tile index is
        // skipped; the slots below the first leaf's offset (overhang of the previous tile's last leaf) stay padding.
        // One block-wide scan, kLeavesPerThread consecutive leaves per thread. Leaves of more than kTileCap / 2 particles (many
        // coincident particles) take the serial greedy packing instead.
        const int cMax = (s.maxLeafCount + 3) & ~3;
        if (2 * cMax <= kTileCap)
        {
            const int B = kTileCap - cMax;
            int       c[kLeavesPerThread], nw[kLeavesPerThread], sc = 0, sw = 0;
#pragma unroll
            for (int k = 0; k < kLeavesPerThread; ++k)
            {
                const int q = kLeavesPerThread * t + k;
                const int n = q < L ? s.leafKey[q] : 0;
                c[k]        = (n + 3) & ~3;
                nw[k]       = (n + 31) >> 5;
                sc += c[k], sw += nw[k];
            }
            int ic = sc, iw = sw;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int vc = __shfl_up_sync(kFullMask, ic, o), vw = __shfl_up_sync(kFullMask, iw, o);
                if (lane >= o) ic += vc, iw += vw;
            }
            int* scanW = reinterpret_cast<int*>(s.red); // [2][kSearchWarps]: the bounding-box reduction is done
            if (lane == 31) s.scan[warp] = ic, scanW[warp] = iw;
            __syncthreads();
            int P = ic - sc, W = iw - sw, totW = 0;
            for (int w = 0; w < kSearchWarps; ++w)
            {
                if (w < warp) P += s.scan[w], W += scanW[w];
                totW += scanW[w];
            }
            const int q0 = kLeavesPerThread * t;
            int prevTile = (q0 > 0 && q0 <= L) ? (P - ((s.leafKey[q0 - 1] + 3) & ~3)) / B : -1;
#pragma unroll
            for (int k = 0; k < kLeavesPerThread; ++k)
            {
                const int q = kLeavesPerThread * t + k;
                if (q < L)
                {
                    const int tile = P / B;
                    if (tile < kMaxTiles)
                    {
                        if (tile != prevTile) s.tileFirstLeaf[tile] = q;
                        if (q == L - 1) s.tileFirstLeaf[tile + 1] = L, s.nTiles = tile + 1;
                    }
                    else { s.err = 1; }
                    prevTile      = tile;
                    s.leafTile[q] = (unsigned short)(P - tile * B);
                    s.leafW0[q]   = (unsigned short)W;
                    if (W + nw[k] <= kMaxWords)
                    {
                        for (int m = 0; m < nw[k]; ++m)
                            s.wordLeaf[W + m] = (unsigned short)q;
                    }
                    else { s.err = 1; }
                }
                P += c[k], W += nw[k];
            }
            if (t == 0)
            {
                s.wEnd = totW;
                if (L == 0) s.tileFirstLeaf[0] = s.tileFirstLeaf[1] = 0, s.nTiles = 1;
            }
        }
        else if (t == 0)
        {
            int W = 0, nT = 0, tileCount = 0, err = 0;
            s.tileFirstLeaf[0] = 0;
            for (int q = 0; q < L; ++q)
            {
                const int n = s.leafKey[q];
                const int c = (n + 3) & ~3; // padded to a multiple of four
                if (c > kTileCap) err = 1;
                if (tileCount + c > kTileCap)
                {
                    ++nT;
                    if (nT >= kMaxTiles)
                    {
                        err = 1;
                        break;
                    }
                    s.tileFirstLeaf[nT] = q;
                    tileCount           = 0;
                }
                const int nw = (n + 31) >> 5;
                if (W + nw > kMaxWords || err)
                {
                    err = 1;
                    break;
                }
                s.leafTile[q] = (unsigned short)tileCount;
                s.leafW0[q]   = (unsigned short)W;
                for (int k = 0; k < nw; ++k)
                    s.wordLeaf[W + k] = (unsigned short)q;
                W += nw;
                tileCount += c;
            }
            ++nT;
            s.tileFirstLeaf[nT] = L;
            s.nTiles            = nT;
            s.wEnd              = W;
            if (err) s.err = 1;
        }
#pragma unroll
        for (int q = 0; q < kWordsPerThread; ++q)
            s.usedBits[q * T + t] = 0;
        __syncthreads();
        if (s.err)
        {
            if (precise) break;
            precise = true;
            __syncthreads();
            continue;
        }
        const int nTiles = s.nTiles;

        // ---------------------------------------------------------------------------------------------------------
        // fp32 filter thresholds. |d2_fp32 - d2_exact| <= 2^-24 (7 r E + 6.5 r^2) near the decision boundary
        // (E bounds every relative coordinate); pairs inside the margin are decided by the exact predicate.
        const float  tx = float(xi - s.origin[0]), ty = float(yi - s.origin[1]), tz = float(zi - s.origin[2]);
        const float2 ntx = make_float2(-tx, -tx), nty = make_float2(-ty, -ty), ntz = make_float2(-tz, -tz);
        float        r2lo, r2hi;
        {
            const float rr     = 2.0f * hi * 1.000001f;
            const float margin = 5.9604645e-8f * 16.0f * (rr * E + rr * rr);
            r2lo               = foldMode ? -1.0f : radiusSq - margin;
            r2hi               = foldMode ? 3.0e38f : radiusSq + margin;
            if (!valid) r2lo = r2hi = -1.0f;
        }
        /* e = d2 - r2lo: the sign bit of e is the sure-hit decision (IEEE subtraction has the exact sign); a pair is
         * ambiguous when 0 <= e < r2hi - r2lo, tested on the bit patterns (unsigned order == float order for e >= 0,
         * negative e compare as huge). The width is inflated so that rounding of e cannot hide a d2 < r2hi. */
        const float2   nlo2      = make_float2(-r2lo, -r2lo);
        const unsigned widthBits = __float_as_uint((r2hi - r2lo) * 1.0001f);

        // per-warp cull volume: bounding box of the warp's targets and its largest (inflated) search radius. A staged quad
        // whose box is further than that from the target box cannot hold a hit or an ambiguous pair of any lane: the
        // gaps are lower bounds of the fp32 coordinate differences the filter evaluates, 1e-5 covers the rounding.
        const float wlx = warpMinF(tx), wly = warpMinF(ty), wlz = warpMinF(tz);
        const float whx = warpMaxF(tx), why = warpMaxF(ty), whz = warpMaxF(tz);
        const float wr2 = warpMaxF(r2hi) * 1.00001f;

        unsigned ent   = 0;
        int      selfW = -1;          // provisional word and bit of the target's own particle, once it has been staged
        unsigned selfClear = ~0u;
        count              = 0;
        for (int tIdx = 0; tIdx < nTiles; ++tIdx)
        {
            const int lb = s.tileFirstLeaf[tIdx], le = s.tileFirstLeaf[tIdx + 1];
            if (lb == le) continue;
            const int tileN = s.leafTile[le - 1] + ((s.leafKey[le - 1] + 3) & ~3);

            // stage the particles and the boxes of their quads (a quad = four lanes; padding does not count)
            {
                const double ox = s.origin[0], oy = s.origin[1], oz = s.origin[2];
                int          l = lb;
                for (int p0 = 0; p0 < tileN; p0 += T)
                {
                    const int  p  = p0 + t;
                    const bool in = p < tileN; // uniform per quad: tileN is a multiple of four
                    float4     rp = make_float4(1e18f, 1e18f, 1e18f, 0.0f); // padding: never a hit
                    bool       real = false;
                    unsigned   meta = 0;
                    if (in)
                    {
                        while (p >= s.leafTile[l] + ((s.leafKey[l] + 3) & ~3))
                            ++l;
                        const int off = p - s.leafTile[l]; // negative below the first leaf of the tile: padding
                        meta = ((unsigned(s.leafW0[l]) + unsigned(max(off, 0) >> 5)) << 5) | (28u - 4u * (unsigned(off >> 2) & 7u));
                        if (off >= 0 && off < s.leafKey[l])
                        {
                            const unsigned j = unsigned(s.leafFirst[l] + off);
                            rp               = relativePosition(a, j, ox, oy, oz, foldMode);
                            real             = true;
                            if (j - iBlock0 < unsigned(T)) s.selfP[j - iBlock0] = 32 * int(s.leafW0[l]) + off;
                        }
                        s.tileX[p] = rp.x, s.tileY[p] = rp.y, s.tileZ[p] = rp.z;
                    }
                    float qlo[3] = {rp.x, rp.y, rp.z};
                    float qhi[3] = {real ? rp.x : -1e18f, real ? rp.y : -1e18f, real ? rp.z : -1e18f};
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                    {
                        qlo[d] = fminf(qlo[d], __shfl_xor_sync(kFullMask, qlo[d], 1));
                        qhi[d] = fmaxf(qhi[d], __shfl_xor_sync(kFullMask, qhi[d], 1));
                        qlo[d] = fminf(qlo[d], __shfl_xor_sync(kFullMask, qlo[d], 2));
                        qhi[d] = fmaxf(qhi[d], __shfl_xor_sync(kFullMask, qhi[d], 2));
                    }
                    if (in && (lane & 3) == 0)
                    {
                        const int Q = p >> 2;
#pragma unroll
                        for (int d = 0; d < 3; ++d)
                        {
                            s.quadBox[d * kTileQuads + Q]       = qlo[d];
                            s.quadBox[(3 + d) * kTileQuads + Q] = qhi[d];
                        }
                        s.quadMeta[Q] = typename Shared::QuadMeta(meta);
                    }
                }
            }
            __syncthreads();
            {
                const int sp = s.selfP[t];
                selfW        = sp >> 5; // negative while unknown
                selfClear    = ~(0x80000000u >> (sp & 31));
            }

            // which quads of the tile are within reach of this warp's targets: one quad per lane and ballot
            {
                const int nQuad = tileN >> 2;
#pragma unroll
                for (int kw = 0; kw < kKeepWords; ++kw)
                {
                    const int Q    = kw * 32 + lane;
                    bool      near = false;
                    if (Q < nQuad)
                    {
                        const float gx = fmaxf(fmaxf(s.quadBox[Q] - whx, wlx - s.quadBox[3 * kTileQuads + Q]), 0.0f);
                        const float gy = fmaxf(fmaxf(s.quadBox[kTileQuads + Q] - why, wly - s.quadBox[4 * kTileQuads + Q]), 0.0f);
                        const float gz = fmaxf(fmaxf(s.quadBox[2 * kTileQuads + Q] - whz, wlz - s.quadBox[5 * kTileQuads + Q]), 0.0f);
                        // (a quad made of padding only has lo > hi: never in reach, also not with fold mode's r2hi)
                        near = gx * gx + gy * gy + gz * gz <= wr2 && s.quadBox[Q] <= s.quadBox[3 * kTileQuads + Q];
                    }
                    const unsigned bits = __ballot_sync(kFullMask, near);
                    if (lane == 0) s.keep[warp][kw] = bits;
                }
                __syncwarp();
            }

            // walk the quads in reach (set bits of the warp's keep mask); the hits of a provisional word are collected
            // in `mask` (staged particle k of the word is bit 31 - k) and flushed when the next word begins
            {
                int      curW = -1;
                unsigned mask = 0;
                auto     flush = [&]()
                {
                    if (curW == selfW) mask &= selfClear; // the target itself is not a neighbour
                    const unsigned any = __reduce_or_sync(kFullMask, mask);
#if SPHX_SEARCH_SADDR
                    if (any && lane == 0) redOrShared(sb + unsigned(offsetof(Shared, usedBits)) + 4u * unsigned(curW), any);
#else
                    if (any && lane == 0) atomicOr(&s.usedBits[curW], any);
#endif
                    if (mask)
                    {
                        count += __popc(mask);
                        // a column holds ngmax + 1 entries or more, each with at least one neighbour: what does not
                        // fit lies beyond the ngmax neighbours the list keeps (count goes on, as in the reference)
                        if (ent < a.maskRows) { maskCol[ent * 32] = make_uint2(mask, unsigned(curW)); }
                        ++ent;
                    }
                };
                const float4* px   = reinterpret_cast<const float4*>(s.tileX);
                const float4* py   = reinterpret_cast<const float4*>(s.tileY);
                const float4* pz   = reinterpret_cast<const float4*>(s.tileZ);
                const int     nKw  = ((tileN >> 2) + 31) >> 5;
                for (int kw = 0; kw < nKw; ++kw)
                {
#if SPHX_SEARCH_SADDR
                    unsigned bits = ldsU32(sb + unsigned(offsetof(Shared, keep)) + 4u * unsigned(warp * kKeepWords + kw));
#else
                    unsigned bits = s.keep[warp][kw];
#endif
#pragma unroll kQuadUnroll
                    while (bits)
                    {
                        const int Q = kw * 32 + __ffs(bits) - 1;
                        bits &= bits - 1;
#if SPHX_SEARCH_SADDR
                        const unsigned meta =
                            sizeof(typename Shared::QuadMeta) == 2
                                ? ldsU16(sb + unsigned(offsetof(Shared, quadMeta)) + 2u * unsigned(Q))
                                : ldsU32(sb + unsigned(offsetof(Shared, quadMeta)) + 4u * unsigned(Q));
#else
                        const unsigned meta = s.quadMeta[Q];
#endif
                        const int      w = int(meta >> 5);
                        const unsigned sh = meta & 31u; // the quad's four slots are bits sh + 3 .. sh of the hit mask
#if SPHX_SEARCH_SADDR
                        // (issued before the word check: the loads do not wait for the meta entry and the flush)
                        const unsigned qa = sb + unsigned(offsetof(Shared, tileX)) + 16u * unsigned(Q);
                        const float4   X = ldsF4(qa), Y = ldsF4(qa + unsigned(sizeof(s.tileX))),
                                     Z = ldsF4(qa + 2u * unsigned(sizeof(s.tileX)));
#endif
                        if (w != curW)
                        {
                            if (curW >= 0) flush();
                            curW = w, mask = 0;
                        }
                        // distances of one quad in packed f32x2 arithmetic
#if !SPHX_SEARCH_SADDR
                        const float4 X = px[Q], Y = py[Q], Z = pz[Q];
#endif
                        const float2 dx0 = __fadd2_rn(make_float2(X.x, X.y), ntx), dx1 = __fadd2_rn(make_float2(X.z, X.w), ntx);
                        const float2 dy0 = __fadd2_rn(make_float2(Y.x, Y.y), nty), dy1 = __fadd2_rn(make_float2(Y.z, Y.w), nty);
                        const float2 dz0 = __fadd2_rn(make_float2(Z.x, Z.y), ntz), dz1 = __fadd2_rn(make_float2(Z.z, Z.w), ntz);
                        const float2 s0 = __ffma2_rn(dz0, dz0, __ffma2_rn(dy0, dy0, __fmul2_rn(dx0, dx0)));
                        const float2 s1 = __ffma2_rn(dz1, dz1, __ffma2_rn(dy1, dy1, __fmul2_rn(dx1, dx1)));
                        const float2 e0 = __fadd2_rn(s0, nlo2), e1 = __fadd2_rn(s1, nlo2);
                        const unsigned eb[4] = {__float_as_uint(e0.x), __float_as_uint(e0.y), __float_as_uint(e1.x),
                                                __float_as_uint(e1.y)};
                        const bool amb = (eb[0] < widthBits) | (eb[1] < widthBits) | (eb[2] < widthBits) |
                                         (eb[3] < widthBits);
                        unsigned nib = 0;
                        if (__any_sync(kFullMask, amb))
                        {
                            // the reference's exact fp64 predicate decides inside the margin (and always in fold mode)
                            const float d2[4] = {s0.x, s0.y, s1.x, s1.y};
                            const int   l     = s.wordLeaf[w];
                            const int   n     = s.leafKey[l];
                            const int   off0  = 32 * (w - int(s.leafW0[l])) + int(28u - sh);
#if SPHX_EXACT_INLINE
                            // one copy of the fp64 code, the four pairs rotate through g0 / f0
                            unsigned g0 = eb[0], g1 = eb[1], g2 = eb[2], g3 = eb[3];
                            float    f0 = d2[0], f1 = d2[1], f2 = d2[2], f3 = d2[3];
#pragma unroll 1
                            for (int u = 0; u < 4; ++u)
                            {
                                bool h = g0 >> 31;
                                if (!h && f0 < r2hi)
                                {
                                    const int off = off0 + u;
                                    h = unsigned(off) < unsigned(n) && exactPair(a.x, a.y, a.z, unsigned(s.leafFirst[l] + off), reload(a.x + il),
                                                             reload(a.y + il), reload(a.z + il), usePbc, box, radiusSq);
                                }
                                nib = (nib << 1) | unsigned(h);
                                g0 = g1, g1 = g2, g2 = g3;
                                f0 = f1, f1 = f2, f2 = f3;
                            }
#else
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                            {
                                bool h = eb[u] >> 31;
                                if (!h && d2[u] < r2hi)
                                {
                                    const int off = off0 + u;
                                    h = unsigned(off) < unsigned(n) && exactPair(a.x, a.y, a.z, unsigned(s.leafFirst[l] + off), reload(a.x + il),
                                                             reload(a.y + il), reload(a.z + il), usePbc, box, radiusSq);
                                }
                                nib = (nib << 1) | unsigned(h);
                            }
#endif
                        }
                        else
                        {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                nib = __funnelshift_l(eb[u], nib, 1); // (nib << 1) | sign(e)
                        }
                        mask |= nib << sh;
                    }
                }
                if (curW >= 0) flush();
            }
            __syncthreads(); // all reads of the tile are done
        }
        numEnt = min(ent, a.maskRows);
        if (s.err) break;
        // count = number of hits without the target's own particle; keeps counting beyond ngmax, as the reference does

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
    const double ox = s.origin[0], oy = s.origin[1], oz = s.origin[2];
    desc.ox = ox, desc.oy = oy, desc.oz = oz;
    desc.flags = foldMode ? kBlockFold : 0u;
    // diagnostics (sim.py: block_stats): leaves in reach, tiles, precise walk, search repetitions of the h-iteration
    desc.pad = unsigned(min(L, 4095)) | (unsigned(min(s.nTiles, 255)) << 12) | (unsigned(min(iteration, 15)) << 20) | (precise ? 0x80000000u : 0u);

    if (s.err)
    {
        if (!C::kOwnScratch)
        {
            // the standard tables are too small for this block: the big instantiation redoes it (h is untouched)
            if (t == 0) a.overflowList[atomicAdd(&a.scal->work[kOverflowCount], 1u)] = blk;
            return;
        }
        // traversal capacity exceeded: flag it, leave an empty block behind
        if (t == 0)
        {
            atomicOr(&a.scal->errFlags, kErrTraversal);
            desc.candBegin = 0, desc.numCand = 0;
            a.blocks[blk] = desc;
        }
        if (valid) a.nc[i] = 1;
        return;
    }

    // popc prefix over the used-slot words -> compact candidate numbering
    const int nW    = s.wEnd;
    int       total = 0;
    {
        const int w0 = t * kWordsPerThread;
        int       loc[kWordsPerThread];
        int       sum = 0;
#pragma unroll
        for (int q = 0; q < kWordsPerThread; ++q)
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
        int base = 0;
        for (int w = 0; w < kSearchWarps; ++w)
        {
            if (w < warp) base += s.scan[w];
            total += s.scan[w];
        }
        base += incl - sum;
#pragma unroll
        for (int q = 0; q < kWordsPerThread; ++q)
            if (w0 + q < nW) s.wordPrefix[w0 + q] = (unsigned short)(base + loc[q]);
        unsigned begin = 0;
        if (t == 0) begin = atomicAdd(&a.scal->candTop, unsigned(total)); // consumed after the list decode
        __syncthreads(); // wordPrefix complete

        // neighbour list: 16-bit candidate indices, 8 per vector, lane-interleaved per group of 32 targets. A lane's
        // hit-mask column is copied from the L2 scratch into its private shared-memory column kDecodeBatch entries at
        // a time (asynchronous copies, all in flight together); within a batch every lane decodes at its own pace,
        // one neighbour per iteration, ascending. The eight 16-bit entries of a vector pass through a 128-bit shift
        // register (entry k of the vector ends up at bits 16 k).
        {
            static_assert(offsetof(Shared, usedBits) >= kDecodeBatch * kSearchThreads * sizeof(uint2),
                          "decode staging area");
            const unsigned kc  = min(count, ngmax);
            uint4*         lp  = a.list + (size_t(blk) * kGroupsPerBlock + warp) * a.nkbMax * kGroupSize + lane;
            uint2*         stg = reinterpret_cast<uint2*>(s.tileX) + t; // entry q of the batch at stg[q * T]
            const unsigned stgAddr = unsigned(__cvta_generic_to_shared(stg));
            const unsigned maxEnt  = warpMaxU(numEnt);
            unsigned       k = 0, v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            for (unsigned r0 = 0; r0 < maxEnt; r0 += kDecodeBatch)
            {
#pragma unroll
                for (unsigned q = 0; q < kDecodeBatch; ++q)
                {
                    if (r0 + q < numEnt)
                    {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stgAddr + q * unsigned(T * sizeof(uint2))),
                                     "l"(maskCol + size_t(r0 + q) * 32)
                                     : "memory");
                    }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                const unsigned nb   = numEnt > r0 ? min(kDecodeBatch, numEnt - r0) : 0u;
                unsigned       q    = 0, mask = 0, ub = 0, pre = 0;
#pragma unroll kDecodeUnroll
                while (k < kc)
                {
#if SPHX_SEARCH_SADDR
                    if (mask == 0)
                    {
                        if (q == nb) break;
                        const uint2 en = ldsU32x2(stgAddr + q * unsigned(T * sizeof(uint2)));
                        ++q;
                        mask = en.x;
                        ub   = ldsU32(sb + unsigned(offsetof(Shared, usedBits)) + 4u * en.y);
                        // (minus one: the hit itself is a used slot, it is counted by popc(ub >> p) below)
                        pre  = ldsU16(sb + unsigned(offsetof(Shared, wordPrefix)) + 2u * en.y) - 1u;
                    }
                    const unsigned p = highestBit(mask); // slot 31 - p of the word
                    mask ^= 1u << p;
                    const unsigned e = pre + __popc(ub >> p); // used slots before this one
#else
                    if (mask == 0)
                    {
                        if (q == nb) break;
                        const uint2 en = stg[q * T];
                        ++q;
                        mask = en.x;
                        ub = s.usedBits[en.y], pre = s.wordPrefix[en.y];
                    }
                    const unsigned b = __clz(mask);
                    mask &= ~(0x80000000u >> b);
                    const unsigned e = pre + __popc(ub & ~(0xffffffffu >> b));
#endif
                    v0 = __funnelshift_r(v0, v1, 16), v1 = __funnelshift_r(v1, v2, 16), v2 = __funnelshift_r(v2, v3, 16);
                    v3 = __funnelshift_r(v3, e, 16);
                    ++k;
                    if ((k & 7) == 0) lp[size_t((k - 1) >> 3) * kGroupSize] = make_uint4(v0, v1, v2, v3);
                }
            }
            if (k & 7)
            {
                // last, partial vector: move its entries down to the low end
                for (unsigned f = k & 7; f < 8; ++f)
                {
                    v0 = __funnelshift_r(v0, v1, 16), v1 = __funnelshift_r(v1, v2, 16), v2 = __funnelshift_r(v2, v3, 16);
                    v3 >>= 16;
                }
                lp[size_t(k >> 3) * kGroupSize] = make_uint4(v0, v1, v2, v3);
            }
        }

        if (t == 0)
        {
            unsigned num = unsigned(total);
            if (size_t(begin) + num > a.candCapacity || num > 0xffffu)
            {
                atomicOr(&a.scal->errFlags, kErrCandSpace);
                begin = 0, num = 0;
            }
            s.candBegin = begin, s.numCand = num;
            desc.candBegin = begin, desc.numCand = num;
            a.blocks[blk] = desc;
        }
        __syncthreads();
    }
    const unsigned candBegin = s.candBegin;
    const unsigned numCand   = s.numCand;
    const bool     haveSpace = numCand == unsigned(total);

    // compacted candidate records, one thread per record: locate the c-th used slot (word by binary search over the
    // prefix, then the bit), so that the fp64 position loads of different records are independent
    for (unsigned c = t; c < numCand; c += T)
    {
        int lo = 0, up = nW;
        while (up - lo > 1)
        {
            const int mid = (lo + up) >> 1;
            if (s.wordPrefix[mid] <= c) { lo = mid; }
            else { up = mid; }
        }
        unsigned m = s.usedBits[lo];
        for (unsigned r = c - s.wordPrefix[lo]; r > 0; --r)
            m &= ~(0x80000000u >> __clz(m));
        const int      l = s.wordLeaf[lo];
        const unsigned j = unsigned(s.leafFirst[l]) + 32u * unsigned(lo - int(s.leafW0[l])) + unsigned(__clz(m));
        a.cand[size_t(candBegin) + c] = relativePosition(a, j, ox, oy, oz, foldMode);
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
        a.nc[i] = haveSpace ? ncSph : 1u; // candidate array exhausted (error flagged): the loops skip the block
    }
}

//! persistent CTAs take blocks of 128 targets from a work counter; each CTA owns one slice of the hit-mask scratch.
//! The big instantiation takes its blocks from the overflow list the standard one left behind.
#ifndef SPHX_SEARCH_CTAS
#define SPHX_SEARCH_CTAS 8 // resident CTAs per SM the register allocation aims for (8: 64 registers; 7: 72, 5 % slower)
#endif
template<bool IterateH, class C>
__global__ void __launch_bounds__(kSearchThreads, C::kOwnScratch ? 1 : SPHX_SEARCH_CTAS)
    blockSearchKernel(const __grid_constant__ SearchArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SearchShared<C>& s = *reinterpret_cast<SearchShared<C>*>(smemRaw);
    uint2* const     maskCol =
        a.maskScratch + (size_t(blockIdx.x) * kSearchWarps + (threadIdx.x >> 5)) * a.maskRows * 32 + (threadIdx.x & 31);
    if constexpr (C::kOwnScratch)
    {
        const unsigned n = a.scal->work[kOverflowCount];
        for (;;)
        {
            __syncthreads(); // the previous block's shared-memory state is dead
            if (threadIdx.x == 0) s.nextBlock = atomicAdd(&a.scal->work[kBigWork], 1u);
            __syncthreads();
            const unsigned idx = s.nextBlock;
            if (idx >= n) break;
            searchBlock<IterateH, C>(a, s, a.overflowList[idx], maskCol);
        }
    }
    else
    {
        unsigned blk = blockIdx.x; // first block static, the following ones from the work counter (fetched one block ahead)
        while (blk < a.numBlocks)
        {
            unsigned nxt = 0;
            if (threadIdx.x == 0) nxt = gridDim.x + atomicAdd(&a.scal->work[kSearchWork], 1u);
            searchBlock<IterateH, C>(a, s, blk, maskCol);
            __syncthreads(); // the block's shared-memory state is dead
            if (threadIdx.x == 0) s.nextBlock = nxt;
            __syncthreads();
            blk = s.nextBlock;
        }
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

template<class C>
static cudaError_t configureSearch()
{
    static std::atomic<bool> configured[64]; // per device
    constexpr size_t         bytes = searchSharedBytes<C>();
    static_assert(bytes <= 227 * 1024, "search kernel shared memory exceeds the 227 KB CTA limit");
    const int dev = DeviceCache::device();
    if (bytes <= 48 * 1024 || configured[dev].load(std::memory_order_acquire)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(blockSearchKernel<true, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(bytes));
    if (e == cudaSuccess) configured[dev].store(true, std::memory_order_release);
    return e;
}

static int smCountSearch() { return DeviceCache::smCount(); }

//! resident CTAs per SM of the search kernel (occupancy query, cached per device)
static int searchCtasPerSm()
{
    static std::atomic<int> n[64];
    const int               dev = DeviceCache::device();
    int                     v   = n[dev].load(std::memory_order_relaxed);
    if (v == 0)
    {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, blockSearchKernel<true, CapsStd>, kSearchThreads,
                                                      searchSharedBytes<CapsStd>());
        if (v <= 0) v = 1;
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

cudaError_t launchBlockSearch(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t stream)
{
    unsigned n = unsigned(a.last - a.first);
    if (n == 0) return cudaSuccess;
    cudaError_t e = configureSearch<CapsStd>();
    if (e != cudaSuccess) return e;
    e = configureSearch<CapsBig>();
    if (e != cudaSuccess) return e;
    char*      base = static_cast<char*>(a.workspace);
    SearchArgs s;
    s.first = unsigned(a.first), s.last = unsigned(a.last);
    s.box  = makeDevBox(a.box);
    s.tree = a.tree;
    s.x = a.f.x, s.y = a.f.y, s.z = a.f.z, s.h = a.f.h, s.nc = a.f.nc;
    s.ng0 = a.p.ng0, s.ngmax = a.p.ngmax, s.nkbMax = w.nkbMax, s.numBlocks = w.numBlocks, s.maskRows = w.maskRows;
    s.maskScratch  = reinterpret_cast<uint2*>(base + w.maskOff);
    s.list         = reinterpret_cast<uint4*>(base + w.listOff);
    s.cand         = reinterpret_cast<float4*>(base + w.candOff);
    s.candCapacity = unsigned(w.candCapacity > 0xffffffffull ? 0xffffffffull : w.candCapacity);
    s.blocks       = reinterpret_cast<BlockDesc*>(base + w.blocksOff);
    s.overflowList = reinterpret_cast<unsigned*>(base + w.overflowOff);
    s.scal         = reinterpret_cast<StepScalars*>(base + w.scalOff);
    unsigned grid = std::min(std::min(w.numBlocks, kSearchMaxCtas), unsigned(searchCtasPerSm() * smCountSearch()));
    blockSearchKernel<true, CapsStd><<<grid, kSearchThreads, searchSharedBytes<CapsStd>(), stream>>>(s);
    // blocks on the overflow list (normally none: the CTAs read a zero count and leave)
    unsigned gridBig = std::min(std::min(w.numBlocks, kSearchMaxCtas), unsigned(smCountSearch()));
    blockSearchKernel<true, CapsBig><<<gridBig, kSearchThreads, searchSharedBytes<CapsBig>(), stream>>>(s);
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

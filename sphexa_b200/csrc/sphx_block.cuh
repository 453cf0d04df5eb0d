/*! @file
 * Target blocks, block-local candidate sets and the compressed neighbour list of the v2 hydro step.
 *
 * Data layout in HBM (all inside the caller's workspace, see api.cu):
 *
 *   targets      [first, last) are cut into blocks of kBlockTargets (128) SFC-consecutive particles; block b owns
 *                targets first + 128 b ... and is made of 4 groups of 32 (one warp each).
 *   candidates   per block, the exact union of the neighbours of its targets, leaf after leaf in the (deterministic,
 *                interleaved) order in which the search staged the leaves, ascending particle index inside a leaf:
 *                cand[candBegin + c] = {x_rel, y_rel, z_rel, bits(j)} where (x_rel, ...) = float(pos_j - origin_b
 *                - periodic shift) is the candidate position relative to the block origin, already shifted to the
 *                periodic image next to the block ("shift mode"), and j is the local particle index. The loops stage
 *                exactly these records (plus the j-side fields they gather through j) in shared memory.
 *   list         16-bit indices into the block's candidate array, 8 per 16-byte vector, lane-interleaved per group:
 *                entries 8 kb .. 8 kb + 7 of target t = 32 g + lane are the uint4 at
 *                list[(g * nkbMax + kb) * 32 + lane], so a warp reads 512 contiguous bytes per 8 neighbours.
 *                Entries of one target are ascending in the candidate index. 2 bytes per neighbour instead of the
 *                reference CPU's 4.
 *
 * A block whose candidate region is too large compared with a periodic box length (tiny test problems) is in
 * "fold mode": candidate positions are stored unshifted and the loops apply the reference's per-pair PBC fold
 * (cstone/sfc/box.hpp:282-304) instead.
 */
#pragma once

#include "sphx_device.cuh"

namespace sphx
{

constexpr int      kBlockTargets   = 128; // targets per block (T)
constexpr int      kGroupsPerBlock = kBlockTargets / int(kGroupSize);
constexpr unsigned kCandPerTarget  = 16;  // candidate-array capacity per assigned particle (typical use: 8-10)
constexpr unsigned kMaxNgmaxStep   = 384; // list vectors per target: (ngmax + 8) / 8 <= 49
//! hit-mask entries {mask, word} per target of the block search: every entry holds at least one neighbour, so ngmax + 1
//! entries hold any list that does not overflow ngmax (multiple of 4: the columns of a CTA stay 16-byte aligned)
__host__ __device__ inline unsigned maskRowsOf(unsigned ngmax) { return (ngmax + 4u) & ~3u; }
#ifndef SPHX_SEARCH_MAX_CTAS
#define SPHX_SEARCH_MAX_CTAS 1280 // >= resident CTAs per SM x SMs (B200: 8 x 148 = 1184)
#endif
constexpr unsigned kSearchMaxCtas  = SPHX_SEARCH_MAX_CTAS; // resident CTAs of the persistent block search (each owns a scratch slice)
constexpr int      kSearchWork     = 5;   // StepScalars::work slot of the block search (0..4: the loop kernels)

constexpr unsigned kBlockFold = 1u; // BlockDesc::flags: fold mode

constexpr int kPolyDeg = 13; // degree (in s = v^2 / 2 - 1) of the polynomials that stand for the kernel tables

struct BlockDesc
{
    double   ox, oy, oz; // block origin (centre of the bounding box of the targets' search spheres)
    unsigned candBegin;  // first entry of this block in the candidate array
    unsigned numCand;
    unsigned flags;
    unsigned pad;
};
static_assert(sizeof(BlockDesc) == 40, "BlockDesc layout");

__host__ __device__ inline unsigned numBlocksOf(size_t numAssigned)
{
    return unsigned((numAssigned + kBlockTargets - 1) / kBlockTargets);
}
//! 8-entry list vectors per target; one spare entry: the search also records the target's own particle provisionally
__host__ __device__ inline unsigned nkbMaxOf(unsigned ngmax) { return (ngmax + 8u) / 8u; }
__host__ __device__ inline size_t   alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

//! workspace carving, shared by api.cu and the kernels' launchers
struct WorkspaceLayout
{
    size_t   scalOff, blocksOff, listOff, candOff, maskOff, overflowOff, total;
    unsigned numBlocks, nkbMax, maskRows;
    size_t   candCapacity;

    __host__ WorkspaceLayout(size_t numAssigned, unsigned ngmax)
    {
        numBlocks    = numBlocksOf(numAssigned);
        nkbMax       = nkbMaxOf(ngmax);
        maskRows     = maskRowsOf(ngmax);
        candCapacity = numAssigned * kCandPerTarget + size_t(kBlockTargets) * ngmax;
        scalOff      = 0;
        blocksOff    = kScalarsBytes;
        listOff      = alignUp(blocksOff + size_t(numBlocks) * sizeof(BlockDesc), 256);
        candOff      = alignUp(listOff + size_t(numBlocks) * kGroupsPerBlock * nkbMax * kGroupSize * 16, 256);
        maskOff      = alignUp(candOff + candCapacity * 16, 256);
        // hit-mask scratch of the block search: one slice per resident CTA, L2-resident (written and read once per block)
        size_t ctas  = numBlocks < kSearchMaxCtas ? numBlocks : kSearchMaxCtas;
        overflowOff  = alignUp(maskOff + ctas * kBlockTargets * maskRows * sizeof(uint2), 256);
        // blocks the standard search tables could not hold (redone by the big instantiation)
        total        = alignUp(overflowOff + size_t(numBlocks) * sizeof(unsigned), 256);
    }
};

//! everything a loop kernel needs, passed by value
struct LoopArgs
{
    SphxFields f;
    unsigned   first, last;
    unsigned   numBlocks, nkbMax, ngmax;
    unsigned   chunkLimit; // test hook: candidates per chunk (0xffffffff: the capacity of the loop's buffer)
    DevBox     box;
    const BlockDesc* blocks;
    const uint4*     list;
    const float4*    cand;
    const float*     wh;
    const float*     whd;
    // kernel tables as polynomials (loops.cu: fitKernelPoly), used by the <Poly = true> instantiations:
    // wh(v) = sum_k pw[k] s^k, v whd(v) = sum_k pd[k] s^k with s = v^2 / 2 - 1; both halves of an entry hold the same
    // coefficient, so that it can be the operand of a packed f32x2 FMA straight from the constant bank
    float2           pw[kPolyDeg + 1];
    float2           pd[kPolyDeg + 1];
    StepScalars*     scal;
    double           K, minDt;
    float            Kcour, alphamin, alphamax, decay_constant, Atmin, Atmax, ramp;
};

} // namespace sphx

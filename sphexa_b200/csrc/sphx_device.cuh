/*! @file
 * Device-side helpers shared by the sphx kernels: box, exact (un-contracted) fp64 geometry for the neighbour
 * search, table lookup, warp utilities.
 *
 * The neighbour search must reproduce the reference CPU predicate bit for bit
 * (domain/include/cstone/findneighbors.hpp:77-147, built without FMA contraction), so every fp64
 * operation on that path is written with __d*_rn intrinsics, which nvcc never fuses into FMAs.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sphx.h"
#include "sphx_powf.h"

namespace sphx
{

constexpr int      kTableSize = 20000; // sph/include/sph/table_lookup.hpp:10
constexpr unsigned kGroupSize = 32;    // targets per warp, one per lane
constexpr unsigned kFullMask  = 0xffffffffu;

//! device image of cstone::Box<double> (sfc/box.hpp:94-174): limits, lengths, inverse lengths, periodic flags
struct DevBox
{
    double xmin, xmax, ymin, ymax, zmin, zmax;
    double lx, ly, lz;
    double ilx, ily, ilz;
    //! lengths multiplied by the periodic flag, i.e. the `pbcX * box.lx()` factor of applyPbc (box.hpp:217-230)
    double plx, ply, plz;
    int    pbcX, pbcY, pbcZ;
    int    anyPbc;
};

inline DevBox makeDevBox(const SphxBox& b)
{
    DevBox d;
    d.xmin = b.lim[0], d.xmax = b.lim[1], d.ymin = b.lim[2], d.ymax = b.lim[3], d.zmin = b.lim[4], d.zmax = b.lim[5];
    d.lx = d.xmax - d.xmin, d.ly = d.ymax - d.ymin, d.lz = d.zmax - d.zmin;
    d.ilx = 1.0 / (d.xmax - d.xmin), d.ily = 1.0 / (d.ymax - d.ymin), d.ilz = 1.0 / (d.zmax - d.zmin);
    d.pbcX = b.boundary[0] == 1, d.pbcY = b.boundary[1] == 1, d.pbcZ = b.boundary[2] == 1;
    d.plx = d.pbcX ? d.lx : 0.0, d.ply = d.pbcY ? d.ly : 0.0, d.plz = d.pbcZ ? d.lz : 0.0;
    d.anyPbc = d.pbcX || d.pbcY || d.pbcZ;
    return d;
}

//! d - pl * rint(d * il), each operation rounded separately (applyPbc, box.hpp:217-230)
__device__ __forceinline__ double foldExact(double d, double pl, double il)
{
    return __dsub_rn(d, __dmul_rn(pl, rint(__dmul_rn(d, il))));
}

//! (a*a + b*b) + c*c, left fold as written in distanceSq (findneighbors.hpp:33-60)
__device__ __forceinline__ double sumSqLeft(double a, double b, double c)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c));
}

//! a*a + (b*b + c*c), right fold of util::array norm2 (util/array.hpp:236-240)
__device__ __forceinline__ double sumSqRight(double a, double b, double c)
{
    return __dadd_rn(__dmul_rn(a, a), __dadd_rn(__dmul_rn(b, b), __dmul_rn(c, c)));
}

//! one component of minDistance point<->box (traversal/boxoverlap.hpp:196-216): max(|d| - s, 0) as (v + |v|) / 2
__device__ __forceinline__ double minDistComp(double d, double s)
{
    double v = __dsub_rn(fabs(d), s);
    v        = __dadd_rn(v, fabs(v));
    return __dmul_rn(v, 0.5);
}

//! sph::updateH (sph/include/sph/kernels.hpp:26-32) for T = float, bit-exact with the reference CPU (sphx_powf.h)
__device__ __forceinline__ float updateH(unsigned ng0, unsigned nc, float h) { return updateHExact(ng0, nc, h); }

//! lt::lookup (sph/include/sph/table_lookup.hpp:13-26), T = float
__device__ __forceinline__ float tableLookup(const float* __restrict__ table, float v)
{
    constexpr int   numIntervals = kTableSize - 1;
    constexpr float dx           = 2.0f / numIntervals;
    constexpr float invDx        = 1.0f / dx;

    int idx = int(v * invDx);
    if (idx >= numIntervals) { return 0.0f; }
    float t0         = table[idx];
    float derivative = (table[idx + 1] - t0) * invDx;
    return t0 + derivative * (v - float(idx) * dx);
}

//! legacy PBC of the J-loops (sfc/box.hpp:282-304), T = float, box lengths are double
__device__ __forceinline__ void applyPBC(const DevBox& box, float r, float& xx, float& yy, float& zz)
{
    if (box.pbcX)
    {
        if (xx > r) xx = float(double(xx) - box.lx);
        else if (xx < -r)
            xx = float(double(xx) + box.lx);
    }
    if (box.pbcY)
    {
        if (yy > r) yy = float(double(yy) - box.ly);
        else if (yy < -r)
            yy = float(double(yy) + box.ly);
    }
    if (box.pbcZ)
    {
        if (zz > r) zz = float(double(zz) - box.lz);
        else if (zz < -r)
            zz = float(double(zz) + box.lz);
    }
}

__device__ __forceinline__ unsigned laneId() { return threadIdx.x & 31u; }

__device__ __forceinline__ double warpMin(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmin(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

__device__ __forceinline__ double warpMax(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

__device__ __forceinline__ unsigned warpMaxU(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = max(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

__device__ __forceinline__ float warpMinF(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fminf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

__device__ __forceinline__ float warpMaxF(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

/*! @brief neighbour list layout in the workspace
 *
 * ELL, lane-interleaved per group of 32 SFC-consecutive targets: entry k of target t = 32*g + lane lives at
 * list[(g * ngmax + k) * 32 + lane], so that the 32 lanes of a warp read one 128-byte line per k.
 */
__host__ __device__ __forceinline__ size_t nbListIndex(size_t group, unsigned ngmax, unsigned k, unsigned lane)
{
    return (group * ngmax + k) * kGroupSize + lane;
}

//! device-resident scalars of one step, at the start of the workspace
struct StepScalars
{
    float              minDtCourant; // min over particles
    float              maxDivv;      // max over particles
    unsigned long long totalNeighbors;
    unsigned           maxNc;
    unsigned           numHIterated;
    unsigned           errFlags; // bit0: h non-convergence, bit1: ngmax overflow, bit2: traversal overflow, bit3: cand
    unsigned           candTop;  // bump allocator of the candidate array (block search)
    unsigned           work[8];  // dynamic work counters of the persistent loop kernels
};

//! the min / max reductions of the all-double path (loops_f64.cu)
struct StepScalarsF64
{
    double minDtCourant;
    double maxDivv;
};

constexpr unsigned kErrHConv     = 1u;
constexpr unsigned kErrNgmax     = 2u;
constexpr unsigned kErrTraversal = 4u;
constexpr unsigned kErrCandSpace = 8u;
constexpr unsigned kErrTable     = 16u; // the kernel tables are not the ones the loops' polynomials were fitted to

constexpr size_t kScalarsBytes = 256; // StepScalars + padding, keeps the list 256-byte aligned

} // namespace sphx

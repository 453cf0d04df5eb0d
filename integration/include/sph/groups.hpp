/*! @file
 * Shadow of the reference's sph/include/sph/groups.hpp for builds that link libsphx: put `-I<sphx>/integration/include`
 * BEFORE `-I<reference>/sph/include` and every `#include "sph/groups.hpp"` finds this file first.
 *
 * Everything of the reference header stays (it is included below, unmodified); only sph::computeGroups changes for GPU
 * datasets. The reference builds its target groups with cstone::computeGroupSplits (traversal/groups_gpu.cu:106-135:
 * SFC-consecutive groups of <= 32 particles, split where neighbours in SFC order are far apart) so that the bounding box
 * of a warp's targets stays small in ITS traversal kernels. libsphx forms its own blocks of 128 targets from
 * [firstBody, lastBody) and never looks at the splits, so the splitter kernels (1.5 ms per step at 8 M particles) are
 * pure overhead there. What remains are the consumers of the GroupView outside the hot path - computePositions,
 * updateSmoothingLength, driveTurbulence (one warp per group, one lane per body) - and those work with any
 * SFC-consecutive groups of at most 32 bodies: uniform ones, filled by one trivial kernel (sph::sphxUniformGroups,
 * integration/sph_gpu_sphx.cu).
 */
#pragma once

#define computeGroups computeGroups_reference
#include_next "sph/groups.hpp"
#undef computeGroups

namespace sph
{

//! data[k] = min(first + k * groupSize, last), k = 0 .. ceil((last - first) / groupSize); defined in sph_gpu_sphx.cu
void sphxUniformGroups(cstone::LocalIndex first, cstone::LocalIndex last, unsigned groupSize,
                       cstone::DeviceVector<cstone::LocalIndex>& data);

//! sph::computeGroups (sph/include/sph/groups.hpp:42-76), same signature
template<typename Tc, class Dataset>
void computeGroups(size_t startIndex, size_t endIndex, Dataset& d, const cstone::Box<Tc>& box,
                   GroupData<typename Dataset::AcceleratorType>& groups)
{
    if constexpr (cstone::HaveGpu<typename Dataset::AcceleratorType>{})
    {
        sphxUniformGroups(startIndex, endIndex, nsGroupSize(), groups.data);
        groups.firstBody  = startIndex;
        groups.lastBody   = endIndex;
        groups.numGroups  = groups.data.size() - 1;
        groups.groupStart = rawPtr(groups.data);
        groups.groupEnd   = rawPtr(groups.data) + 1;
    }
    else { computeGroups_reference(startIndex, endIndex, d, box, groups); }
}

} // namespace sph

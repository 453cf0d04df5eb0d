/*! @file internal launcher declarations shared between the .cu translation units and the C-ABI layer */
#pragma once

#include <cuda_runtime.h>

#include "sphx.h"

namespace sphx
{

struct StepScalars;

void launchResetScalars(StepScalars* s, cudaStream_t stream);
void launchFindNeighborsXmass(const SphxStepArgs& a, unsigned* list, StepScalars* scal, cudaStream_t stream);
void launchFindNeighbors(const double* x, const double* y, const double* z, const float* h, unsigned first,
                         unsigned last, const SphxBox& box, const SphxTreeView& tree, unsigned ngmax, unsigned* list,
                         unsigned* counts, StepScalars* scal, cudaStream_t stream);
void launchExportNeighbors(unsigned numAssigned, unsigned ngmax, const unsigned* list, const unsigned* counts,
                           bool countsIncludeSelf, unsigned* out, cudaStream_t stream);

void launchVeDefGradh(const SphxStepArgs& a, const unsigned* list, cudaStream_t s);
void launchEos(const SphxStepArgs& a, cudaStream_t s);
void launchIadDivvCurlv(const SphxStepArgs& a, const unsigned* list, StepScalars* scal, cudaStream_t s);
void launchAvSwitches(const SphxStepArgs& a, const unsigned* list, cudaStream_t s);
void launchMomentumEnergy(const SphxStepArgs& a, const unsigned* list, StepScalars* scal, cudaStream_t s);

} // namespace sphx

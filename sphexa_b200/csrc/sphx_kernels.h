/*! @file internal launcher declarations shared between the .cu translation units and the C-ABI layer */
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>

#include "sphx.h"

namespace sphx
{

struct StepScalars;

struct WorkspaceLayout;

//! set the text returned by sphx_last_error() (api.cu)
void setLastError(const std::string& msg);

void launchResetScalars(StepScalars* s, cudaStream_t stream);

// generic search (neighbors.cu), v1 list format: u32 particle indices, lane-interleaved ELL
void launchFindNeighbors(const double* x, const double* y, const double* z, const float* h, unsigned first,
                         unsigned last, const SphxBox& box, const SphxTreeView& tree, unsigned ngmax, unsigned* list,
                         unsigned* counts, StepScalars* scal, cudaStream_t stream);
void launchExportNeighbors(unsigned numAssigned, unsigned ngmax, const unsigned* list, const unsigned* counts,
                           bool countsIncludeSelf, unsigned* out, cudaStream_t stream);
// the same search with the coupled h-iteration for the all-double type set (loops_f64.cu)
void launchFindNeighborsSphF64(const double* x, const double* y, const double* z, double* h, unsigned first, unsigned last,
                               const SphxBox& box, const SphxTreeView& tree, unsigned ng0, unsigned ngmax,
                               unsigned* list, unsigned* nc, StepScalars* scal, cudaStream_t stream);

// block search of the hydro step (search.cu)
cudaError_t launchBlockSearch(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t stream);
void launchExportBlockNeighbors(const SphxStepArgs& a, const WorkspaceLayout& w, unsigned* out, cudaStream_t stream);

// particle loops (loops.cu)
cudaError_t launchXMass(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t s);
cudaError_t launchVeDefGradh(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t s);
void        launchEos(const SphxStepArgs& a, cudaStream_t s);
cudaError_t launchIadDivvCurlv(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t s);
cudaError_t launchAvSwitches(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t s);
cudaError_t launchMomentumEnergy(const SphxStepArgs& a, const WorkspaceLayout& w, cudaStream_t s);
void        setCandidateChunkLimit(unsigned n);
void        invalidateKernelPolys();
//! 1 if the loops evaluate the tables at (wh, whd) through their polynomial fits, 0 if through shared-memory tables
int         kernelPolyStatus(const float* wh, const float* whd, cudaStream_t s, double* errW, double* errD);

/*! per-device caches of launch parameters: one process may drive several GPUs (one thread per GPU), so nothing that
 *  depends on the device is kept in a plain static. Slot = current device (0 .. 63); values are written with relaxed
 *  atomics, every writer stores the same value. */
struct DeviceCache
{
    static int device()
    {
        int dev = 0;
        cudaGetDevice(&dev);
        return dev < 0 || dev >= 64 ? 0 : dev;
    }
    //! number of SMs of the current device
    static int smCount()
    {
        static std::atomic<int> n[64];
        const int               dev = device();
        int                     v   = n[dev].load(std::memory_order_relaxed);
        if (v == 0)
        {
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
            if (v <= 0) v = 148;
            n[dev].store(v, std::memory_order_relaxed);
        }
        return v;
    }
};

// Hilbert state machine tables for the device (host_domain.cpp)
int hilbertTablesFlat(uint8_t* digit, uint8_t* next, uint8_t* octant, int maxStates);

} // namespace sphx

/*! @file
 * Reference-side binding of libsphx: the translation unit a SPH-EXA maintainer compiles INSTEAD of
 *     sph/include/sph/hydro_ve/{xmass,ve_def_gradh,iad_divv_curlv,av_switches,momentum_energy,eos}_gpu.cu
 * inside the reference's `sph_gpu` library (sph/include/sph/CMakeLists.txt:13-24), linking libsphx.so.
 *
 * It provides exactly the symbols those six files provide - the explicit instantiations of the templates declared in
 * sph/include/sph/sph_gpu.hpp:24-56 for sphexa::ParticlesData<cstone::GpuTag> and cstone::Box<double>, the EOS entry
 * points, and sph::nsGroupSize() - and forwards each of them to the C ABI of include/sphx.h. Nothing else of the
 * reference changes: ParticlesData, Domain, the propagator (main/src/propagator/ve_hydro.hpp), the field
 * acquire/release aliasing and the halo exchanges stay as they are.
 *
 * Compiles against the UNMODIFIED reference headers (it is not part of libsphx and contains no kernels):
 *     nvcc -std=c++20 -DUSE_CUDA -I<reference>/domain/include -I<reference>/sph/include -I<sphx>/include -c ...
 * `make -C oracle ref-cuda-sphx` does this and links the reference's own CUDA application with it
 * (oracle/_ref/sphexa_cuda_sphx), which is how the 100-step Sedov energy comparison of BASELINE.json is run.
 */
#include <stdexcept>
#include <string>

#include "cstone/cuda/cuda_utils.cuh"
#include "cstone/cuda/device_vector.h"
#include "cstone/util/reallocate.hpp"

#include "sph/sph_gpu.hpp"
#include "sph/particles_data.hpp"
#include "sph/eos.hpp"

#include "sphx.h"

namespace sph
{

//! groups handed to the loops by sph::computeGroups (sph/include/sph/groups.hpp:42-76); libsphx forms its own blocks of
//! 128 targets over [firstBody, lastBody) and uses the GroupView for that range only
unsigned nsGroupSize() { return 32; }

//! uniform SFC-consecutive target groups for the consumers of the GroupView outside the hot path (see
//! integration/include/sph/groups.hpp, the shadow of sph/groups.hpp that calls this instead of the group splitter)
__global__ void uniformGroupsKernel(cstone::LocalIndex first, cstone::LocalIndex last, unsigned groupSize,
                                    cstone::LocalIndex numEntries, cstone::LocalIndex* data)
{
    cstone::LocalIndex k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < numEntries)
    {
        unsigned long long v = (unsigned long long)(first) + (unsigned long long)(k) * groupSize;
        data[k] = v < last ? cstone::LocalIndex(v) : last;
    }
}

void sphxUniformGroups(cstone::LocalIndex first, cstone::LocalIndex last, unsigned groupSize,
                       cstone::DeviceVector<cstone::LocalIndex>& data)
{
    cstone::LocalIndex numGroups = (last - first + groupSize - 1) / groupSize;
    reallocate(data, size_t(numGroups) + 1, 1.01);
    uniformGroupsKernel<<<(numGroups + 256) / 256, 256>>>(first, last, groupSize, numGroups + 1, rawPtr(data));
}

namespace cuda
{
namespace
{

//! status code -> the reference's error behaviour (xmass_gpu.cu:127-128 throws std::runtime_error)
void check(int rc)
{
    if (rc != SPHX_OK) { throw std::runtime_error(std::string(sphx_last_error()) + "\n"); }
}

template<class Dataset>
SphxStepArgs makeArgs(const GroupView& grp, Dataset& d, const cstone::Box<typename Dataset::RealType>& box,
                      bool growWorkspace)
{
    auto&        dd = d.devData;
    SphxStepArgs a{};
    a.f.x = rawPtr(dd.x), a.f.y = rawPtr(dd.y), a.f.z = rawPtr(dd.z);
    a.f.h = rawPtr(dd.h), a.f.m = rawPtr(dd.m);
    a.f.vx = rawPtr(dd.vx), a.f.vy = rawPtr(dd.vy), a.f.vz = rawPtr(dd.vz);
    a.f.temp = rawPtr(dd.temp), a.f.u = rawPtr(dd.u);
    a.f.nc = rawPtr(dd.nc), a.f.xm = rawPtr(dd.xm), a.f.kx = rawPtr(dd.kx), a.f.gradh = rawPtr(dd.gradh);
    a.f.prho = rawPtr(dd.prho), a.f.c = rawPtr(dd.c), a.f.rho = rawPtr(dd.rho), a.f.p = rawPtr(dd.p);
    a.f.c11 = rawPtr(dd.c11), a.f.c12 = rawPtr(dd.c12), a.f.c13 = rawPtr(dd.c13);
    a.f.c22 = rawPtr(dd.c22), a.f.c23 = rawPtr(dd.c23), a.f.c33 = rawPtr(dd.c33);
    a.f.divv = rawPtr(dd.divv), a.f.curlv = rawPtr(dd.curlv), a.f.alpha = rawPtr(dd.alpha);
    a.f.ax = rawPtr(dd.ax), a.f.ay = rawPtr(dd.ay), a.f.az = rawPtr(dd.az), a.f.du = rawPtr(dd.du);
    a.f.dV11 = rawPtr(dd.dV11), a.f.dV12 = rawPtr(dd.dV12), a.f.dV13 = rawPtr(dd.dV13);
    a.f.dV22 = rawPtr(dd.dV22), a.f.dV23 = rawPtr(dd.dV23), a.f.dV33 = rawPtr(dd.dV33);

    a.numLocal = dd.x.size();
    a.first    = grp.firstBody;
    a.last     = grp.lastBody;

    a.p.K = d.K, a.p.Kcour = d.Kcour, a.p.Krho = d.Krho, a.p.gamma = d.gamma, a.p.minDt = d.minDt;
    a.p.polytropic_const = d.polytropic_const, a.p.polytropic_index = d.polytropic_index;
    a.p.muiConst = d.muiConst, a.p.soundSpeedConst = d.soundSpeedConst;
    a.p.alphamin = d.alphamin, a.p.alphamax = d.alphamax, a.p.decay_constant = d.decay_constant;
    a.p.Atmin = d.Atmin, a.p.Atmax = d.Atmax, a.p.ramp = d.ramp;
    a.p.ng0 = d.ng0, a.p.ngmax = d.ngmax;
    a.p.eosChoice = int(d.eosChoice);
    a.p.avClean   = dd.dV11.size() > 0;

    a.box.lim[0] = box.xmin(), a.box.lim[1] = box.xmax(), a.box.lim[2] = box.ymin(), a.box.lim[3] = box.ymax();
    a.box.lim[4] = box.zmin(), a.box.lim[5] = box.zmax();
    a.box.boundary[0] = int(box.boundaryX()), a.box.boundary[1] = int(box.boundaryY());
    a.box.boundary[2] = int(box.boundaryZ());

    const auto& t       = d.treeView;
    a.tree.numLeafNodes = t.numLeafNodes;
    a.tree.numNodes     = 0; // not part of OctreeNsView; libsphx does not need it
    a.tree.prefixes     = reinterpret_cast<const uint64_t*>(t.prefixes);
    a.tree.childOffsets = t.childOffsets, a.tree.internalToLeaf = t.internalToLeaf, a.tree.levelRange = t.levelRange;
    a.tree.leaves          = reinterpret_cast<const uint64_t*>(t.leaves);
    a.tree.layout          = t.layout;
    a.tree.centers         = reinterpret_cast<const double*>(t.centers);
    a.tree.sizes           = reinterpret_cast<const double*>(t.sizes);
    a.tree.searchExtFactor = t.searchExtFactor;

    a.wh = rawPtr(dd.wh), a.whd = rawPtr(dd.whd);

    // workspace = devData.traversalStack, as in the reference (cstone::allocateNcStacks, find_neighbors.cuh:492-505);
    // it holds the neighbour list between computeXMass and computeMomentumEnergy of one step
    size_t bytes = sphx_workspace_bytes(a.last - a.first, a.p.ngmax);
    using Elem   = typename std::decay_t<decltype(dd.traversalStack)>::value_type;
    size_t elems = (bytes + sizeof(Elem) - 1) / sizeof(Elem);
    if (growWorkspace && dd.traversalStack.size() < elems) { reallocateDestructive(dd.traversalStack, elems, 1.01); }
    a.workspace      = rawPtr(dd.traversalStack);
    a.workspaceBytes = dd.traversalStack.size() * sizeof(Elem);
    a.stream         = nullptr; // the reference drives everything on the default stream
    return a;
}

} // namespace

template<class Dataset>
void computeXMass(const GroupView& grp, Dataset& d, const cstone::Box<typename Dataset::RealType>& box)
{
    SphxStepArgs   a = makeArgs(grp, d, box, true);
    SphxStepResult r;
    check(sphx_find_neighbors_xmass(&a, &r)); // synchronises; throws on traversal overflow / h non-convergence
}

//! rho_i = m_i / (value left in rho by the XMass pass): the std-hydro density built on computeXMass
//! (xmass_gpu.cu:134-165); part of this file only because the reference keeps it in xmass_gpu.cu
__global__ void xmassToDensity(unsigned first, unsigned last, const float* m, float* rho)
{
    unsigned i = first + blockDim.x * blockIdx.x + threadIdx.x;
    if (i < last) { rho[i] = m[i] / rho[i]; }
}

template<class Dataset>
void computeDensity(const GroupView& grp, Dataset& d, const cstone::Box<typename Dataset::RealType>& box)
{
    swap(d.devData.xm, d.devData.rho);
    computeXMass(grp, d, box);
    swap(d.devData.xm, d.devData.rho);
    unsigned n = grp.lastBody - grp.firstBody;
    if (n == 0) { return; }
    xmassToDensity<<<(n + 255) / 256, 256>>>(grp.firstBody, grp.lastBody, rawPtr(d.devData.m), rawPtr(d.devData.rho));
}

template<class Dataset>
void computeVeDefGradh(const GroupView& grp, Dataset& d, const cstone::Box<typename Dataset::RealType>& box)
{
    SphxStepArgs a = makeArgs(grp, d, box, false);
    check(sphx_ve_def_gradh(&a));
}

template<class Dataset>
void computeIadDivvCurlv(const GroupView& grp, Dataset& d, const cstone::Box<typename Dataset::RealType>& box)
{
    SphxStepArgs a = makeArgs(grp, d, box, false);
    check(sphx_iad_divv_curlv(&a, nullptr)); // the reference computes rhoTimestep separately (ts_global.hpp:72-95)
}

template<class Dataset>
void computeAVswitches(const GroupView& grp, Dataset& d, const cstone::Box<typename Dataset::RealType>& box)
{
    SphxStepArgs a = makeArgs(grp, d, box, false);
    check(sphx_av_switches(&a));
}

template<bool avClean, class Dataset>
void computeMomentumEnergy(const GroupView& grp, float* groupDt, Dataset& d,
                           const cstone::Box<typename Dataset::RealType>& box)
{
    if (groupDt != nullptr)
    {
        throw std::runtime_error("libsphx: per-group time steps (block time-step propagator) are out of scope\n");
    }
    SphxStepArgs a = makeArgs(grp, d, box, false);
    a.p.avClean    = avClean;
    SphxStepResult r;
    check(sphx_momentum_energy(&a, &r));
    d.minDtCourant = r.minDtCourant; // momentum_energy_gpu.cu:141-143
}

template<class Tt, class Tm, class Thydro>
void computeIdealGasEOS(size_t first, size_t last, Tm mui, Tt gamma, const Tt* temp, const Tt* u, const Tm* m,
                        const Thydro* kx, const Thydro* xm, const Thydro* gradh, Thydro* prho, Thydro* c, Thydro* rho,
                        Thydro* p)
{
    static_assert(std::is_same_v<Tt, double> && std::is_same_v<Tm, float> && std::is_same_v<Thydro, float>,
                  "libsphx implements the production type set (sph/include/sph/types.hpp:39-46)");
    SphxStepArgs a{};
    a.f.temp = temp, a.f.u = u, a.f.m = m, a.f.kx = const_cast<float*>(kx), a.f.xm = const_cast<float*>(xm);
    a.f.gradh = const_cast<float*>(gradh), a.f.prho = prho, a.f.c = c, a.f.rho = rho, a.f.p = p;
    a.first = first, a.last = last, a.numLocal = last;
    a.p.gamma = gamma, a.p.muiConst = mui, a.p.eosChoice = 0;
    check(sphx_eos(&a));
}

template<class Th, class Tu>
void computeIsothermalEOS(size_t first, size_t last, Th cConst, Th* c, Th* rho, Th* p, const Th* m, const Th* kx,
                          const Th* xm, const Th* gradh, Th* prho, Tu* temp)
{
    static_assert(std::is_same_v<Th, float>, "libsphx implements the production type set");
    SphxStepArgs a{};
    a.f.m = m, a.f.kx = const_cast<float*>(kx), a.f.xm = const_cast<float*>(xm);
    a.f.gradh = const_cast<float*>(gradh), a.f.prho = prho, a.f.c = c, a.f.rho = rho, a.f.p = p;
    a.first = first, a.last = last, a.numLocal = last;
    a.p.soundSpeedConst = cConst, a.p.eosChoice = 1;
    check(sphx_eos(&a));
    if (temp) { checkGpuErrors(cudaMemset(temp + first, 0, (last - first) * sizeof(Tu))); } // eos_gpu.cu:100
}

template<class Th, class Tt>
void computePolytropicEOS(size_t first, size_t last, Tt polytropic_const, Tt polytropic_index, Th* rho, Th* p,
                          const Th* m, const Th* kx, const Th* xm, const Th* gradh, Th* prho, Tt* temp, Th* c)
{
    static_assert(std::is_same_v<Th, float> && std::is_same_v<Tt, double>,
                  "libsphx implements the production type set");
    SphxStepArgs a{};
    a.f.m = m, a.f.kx = const_cast<float*>(kx), a.f.xm = const_cast<float*>(xm);
    a.f.gradh = const_cast<float*>(gradh), a.f.prho = prho, a.f.c = c, a.f.rho = rho, a.f.p = p;
    a.first = first, a.last = last, a.numLocal = last;
    a.p.polytropic_const = polytropic_const, a.p.polytropic_index = polytropic_index, a.p.eosChoice = 2;
    check(sphx_eos(&a));
    if (temp) { checkGpuErrors(cudaMemset(temp + first, 0, (last - first) * sizeof(Tt))); }
}

// the instantiations the reference's propagators link against (xmass_gpu.cu:131, ve_def_gradh_gpu.cu:99,
// iad_divv_curlv_gpu.cu:109, av_switches_gpu.cu:101, momentum_energy_gpu.cu:146-151, eos_gpu.cu:83-160)
using DataGpu = sphexa::ParticlesData<cstone::GpuTag>;
using BoxD    = cstone::Box<SphTypes::CoordinateType>;

template void computeXMass(const GroupView&, DataGpu&, const BoxD&);
template void computeDensity(const GroupView&, DataGpu&, const BoxD&);
template void computeVeDefGradh(const GroupView&, DataGpu&, const BoxD&);
template void computeIadDivvCurlv(const GroupView&, DataGpu&, const BoxD&);
template void computeAVswitches(const GroupView&, DataGpu&, const BoxD&);
template void computeMomentumEnergy<true>(const GroupView&, float*, DataGpu&, const BoxD&);
template void computeMomentumEnergy<false>(const GroupView&, float*, DataGpu&, const BoxD&);
template void computeIdealGasEOS(size_t, size_t, float, double, const double*, const double*, const float*,
                                 const float*, const float*, const float*, float*, float*, float*, float*);
template void computeIsothermalEOS(size_t, size_t, float, float*, float*, float*, const float*, const float*,
                                   const float*, const float*, float*, double*);
template void computePolytropicEOS(size_t, size_t, double, double, float*, float*, const float*, const float*,
                                   const float*, const float*, float*, double*, float*);

} // namespace cuda
} // namespace sph

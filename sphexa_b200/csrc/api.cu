/*! @file
 * C-ABI layer of libsphx (include/sphx.h): argument checks, workspace carving, status codes. No torch types, no
 * allocation of field memory; the only device memory the callee owns is carved out of the caller's workspace.
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "sphx_block.cuh"
#include "sphx_kernels.h"

namespace
{

thread_local std::string g_lastError;

int fail(int code, const std::string& msg)
{
    g_lastError = msg;
    return code;
}

int cudaFail(cudaError_t e, const char* what)
{
    return fail(SPHX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define SPHX_CUDA(call)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) return cudaFail(e_, #call);                                                             \
    } while (0)

size_t numGroupsOf(size_t n) { return (n + sphx::kGroupSize - 1) / sphx::kGroupSize; }

//! v1 list of the generic search: u32 particle indices, lane-interleaved ELL
size_t genericListBytes(size_t numAssigned, unsigned ngmax)
{
    return numGroupsOf(numAssigned) * size_t(ngmax) * sphx::kGroupSize * sizeof(unsigned);
}

struct Workspace
{
    sphx::StepScalars*    scal;
    sphx::WorkspaceLayout layout;
    Workspace() : scal(nullptr), layout(0, 8) {}
};

int carve(const SphxStepArgs* a, Workspace& w)
{
    if (!a) return fail(SPHX_ERR_INVALID, "null args");
    if (a->last < a->first || a->last > a->numLocal) return fail(SPHX_ERR_INVALID, "bad [first,last) range");
    if (a->numLocal >= (size_t(1) << 32)) return fail(SPHX_ERR_INVALID, "more than 2^32 local particles");
    if (!a->workspace) return fail(SPHX_ERR_INVALID, "null workspace");
    if (a->p.ngmax == 0 || a->p.ngmax > sphx::kMaxNgmaxStep)
        return fail(SPHX_ERR_INVALID, "ngmax must be in [1, " + std::to_string(sphx::kMaxNgmaxStep) + "]");
    w.layout = sphx::WorkspaceLayout(a->last - a->first, a->p.ngmax);
    if (a->workspaceBytes < w.layout.total)
        return fail(SPHX_ERR_WORKSPACE, "workspace too small: need " + std::to_string(w.layout.total) + " bytes");
    if (reinterpret_cast<uintptr_t>(a->workspace) % 16 != 0) return fail(SPHX_ERR_INVALID, "workspace not 16B aligned");
    w.scal = reinterpret_cast<sphx::StepScalars*>(a->workspace);
    return SPHX_OK;
}

int checkTree(const SphxTreeView& t)
{
    if (t.numLeafNodes <= 0 || !t.childOffsets || !t.internalToLeaf || !t.layout || !t.centers || !t.sizes)
        return fail(SPHX_ERR_INVALID, "incomplete tree view");
    return SPHX_OK;
}

#define REQUIRE(ptr)                                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(ptr)) return fail(SPHX_ERR_INVALID, "required field is NULL: " #ptr);                                    \
    } while (0)

int readScalars(const Workspace& w, cudaStream_t s, sphx::StepScalars& host)
{
    SPHX_CUDA(cudaMemcpyAsync(&host, w.scal, sizeof(host), cudaMemcpyDeviceToHost, s));
    SPHX_CUDA(cudaStreamSynchronize(s));
    return SPHX_OK;
}

int errFlagsToStatus(unsigned flags)
{
    if (flags & sphx::kErrTraversal)
        return fail(SPHX_ERR_TRAVERSAL, "GPU traversal stack exhausted in neighbor search");
    if (flags & sphx::kErrHConv) return fail(SPHX_ERR_H_CONVERGENCE, "coupled nc/h-updated failed to converge");
    if (flags & sphx::kErrNgmax) return fail(SPHX_ERR_NGMAX_OVERFLOW, "neighbour count exceeds ngmax after h-iteration");
    if (flags & sphx::kErrTable)
        return fail(SPHX_ERR_TABLE, "the kernel tables wh / whd changed under the addresses the loops fitted their "
                                    "polynomials to: call sphx_invalidate_tables() after rewriting a table");
    if (flags & sphx::kErrCandSpace)
        return fail(SPHX_ERR_WORKSPACE, "candidate array of the workspace exhausted (more than 16 candidates per particle)");
    return SPHX_OK;
}

void fillResult(const sphx::StepScalars& h, const SphxParams& p, SphxStepResult* r)
{
    if (!r) return;
    r->minDtCourant   = double(h.minDtCourant);
    // ts_global.hpp:94 (double / float). A rank without particles has no divergence to report: its max divv is still the
    // identity of the reduction (-inf), and Krho / |-inf| = 0 would win the MIN over the ranks as a zero time step
    r->minDtRho = h.maxDivv == -INFINITY ? double(INFINITY) : p.Krho / std::fabs(double(h.maxDivv));
    r->totalNeighbors = h.totalNeighbors;
    r->maxNc          = h.maxNc;
    r->numHIterated   = h.numHIterated;
}

} // namespace

namespace sphx
{
void setLastError(const std::string& msg) { g_lastError = msg; }
} // namespace sphx

extern "C"
{

const char* sphx_last_error(void) { return g_lastError.c_str(); }
int         sphx_abi_version(void) { return SPHX_ABI_VERSION; }
void        sphx_debug_candidate_chunk(unsigned maxRecords) { sphx::setCandidateChunkLimit(maxRecords); }

int sphx_device_check(void)
{
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
    {
        cudaGetLastError();
        return fail(SPHX_ERR_NO_DEVICE, "no CUDA device available: libsphx has no CPU fallback");
    }
    return SPHX_OK;
}

size_t sphx_workspace_bytes(size_t numAssigned, unsigned ngmax)
{
    return sphx::WorkspaceLayout(numAssigned, ngmax).total;
}

void sphx_invalidate_tables(void) { sphx::invalidateKernelPolys(); }

int sphx_table_mode(const float* wh, const float* whd, void* stream, double* errW, double* errD)
{
    if (int rc = sphx_device_check()) return -rc;
    if (!wh || !whd) return -fail(SPHX_ERR_INVALID, "null table");
    return sphx::kernelPolyStatus(wh, whd, static_cast<cudaStream_t>(stream), errW, errD);
}

void sphx_workspace_layout(size_t numAssigned, unsigned ngmax, size_t out[8])
{
    sphx::WorkspaceLayout w(numAssigned, ngmax);
    out[0] = w.scalOff, out[1] = w.blocksOff, out[2] = w.listOff, out[3] = w.candOff, out[4] = w.total;
    out[5] = w.numBlocks, out[6] = w.nkbMax, out[7] = w.candCapacity;
}

int sphx_find_neighbors_sph(const SphxStepArgs* a, SphxStepResult* r)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    if (int rc = checkTree(a->tree)) return rc;
    if (a->p.ng0 > a->p.ngmax) return fail(SPHX_ERR_INVALID, "ng0 should be smaller than ngmax");
    REQUIRE(a->f.x); REQUIRE(a->f.y); REQUIRE(a->f.z); REQUIRE(a->f.h); REQUIRE(a->f.nc);
    auto s = static_cast<cudaStream_t>(a->stream);
    sphx::launchResetScalars(w.scal, s);
    SPHX_CUDA(sphx::launchBlockSearch(*a, w.layout, s));
    if (r)
    {
        sphx::StepScalars h;
        if (int rc = readScalars(w, s, h)) return rc;
        fillResult(h, a->p, r);
        return errFlagsToStatus(h.errFlags);
    }
    return SPHX_OK;
}

int sphx_xmass(const SphxStepArgs* a)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    REQUIRE(a->f.m); REQUIRE(a->f.xm); REQUIRE(a->f.nc); REQUIRE(a->wh);
    SPHX_CUDA(sphx::launchXMass(*a, w.layout, static_cast<cudaStream_t>(a->stream)));
    return SPHX_OK;
}

int sphx_find_neighbors_xmass(const SphxStepArgs* a, SphxStepResult* r)
{
    if (int rc = sphx_device_check()) return rc;
    if (!a) return fail(SPHX_ERR_INVALID, "null args");
    REQUIRE(a->f.m); REQUIRE(a->f.xm); REQUIRE(a->wh);
    if (int rc = sphx_find_neighbors_sph(a, nullptr)) return rc;
    if (int rc = sphx_xmass(a)) return rc;
    if (r)
    {
        Workspace w;
        if (int rc = carve(a, w)) return rc;
        sphx::StepScalars h;
        if (int rc = readScalars(w, static_cast<cudaStream_t>(a->stream), h)) return rc;
        fillResult(h, a->p, r);
        return errFlagsToStatus(h.errFlags);
    }
    return SPHX_OK;
}

int sphx_ve_def_gradh(const SphxStepArgs* a)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    REQUIRE(a->f.xm); REQUIRE(a->f.kx); REQUIRE(a->f.gradh); REQUIRE(a->wh); REQUIRE(a->whd); REQUIRE(a->f.nc);
    SPHX_CUDA(sphx::launchVeDefGradh(*a, w.layout, static_cast<cudaStream_t>(a->stream)));
    return SPHX_OK;
}

int sphx_eos(const SphxStepArgs* a)
{
    if (int rc = sphx_device_check()) return rc;
    if (!a) return fail(SPHX_ERR_INVALID, "null args");
    REQUIRE(a->f.kx); REQUIRE(a->f.xm); REQUIRE(a->f.m); REQUIRE(a->f.gradh); REQUIRE(a->f.prho); REQUIRE(a->f.c);
    if (a->p.eosChoice == 0 && !a->f.temp && !a->f.u) return fail(SPHX_ERR_INVALID, "ideal gas EOS needs temp or u");
    if (a->p.eosChoice < 0 || a->p.eosChoice > 2) return fail(SPHX_ERR_INVALID, "unknown eosChoice");
    sphx::launchEos(*a, static_cast<cudaStream_t>(a->stream));
    SPHX_CUDA(cudaGetLastError());
    return SPHX_OK;
}

int sphx_iad_divv_curlv(const SphxStepArgs* a, SphxStepResult* r)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    REQUIRE(a->f.vx); REQUIRE(a->f.vy); REQUIRE(a->f.vz); REQUIRE(a->f.xm); REQUIRE(a->f.kx); REQUIRE(a->f.c11);
    REQUIRE(a->f.c12); REQUIRE(a->f.c13); REQUIRE(a->f.c22); REQUIRE(a->f.c23); REQUIRE(a->f.c33); REQUIRE(a->f.divv);
    auto s = static_cast<cudaStream_t>(a->stream);
    SPHX_CUDA(sphx::launchIadDivvCurlv(*a, w.layout, s));
    if (r)
    {
        sphx::StepScalars h;
        if (int rc = readScalars(w, s, h)) return rc;
        fillResult(h, a->p, r);
    }
    return SPHX_OK;
}

int sphx_av_switches(const SphxStepArgs* a)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    REQUIRE(a->f.c); REQUIRE(a->f.divv); REQUIRE(a->f.alpha); REQUIRE(a->f.c11);
    SPHX_CUDA(sphx::launchAvSwitches(*a, w.layout, static_cast<cudaStream_t>(a->stream)));
    return SPHX_OK;
}

int sphx_momentum_energy(const SphxStepArgs* a, SphxStepResult* r)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    REQUIRE(a->f.prho); REQUIRE(a->f.c); REQUIRE(a->f.alpha); REQUIRE(a->f.ax); REQUIRE(a->f.ay); REQUIRE(a->f.az);
    REQUIRE(a->f.du);
    if (a->p.avClean) { REQUIRE(a->f.dV11); }
    auto s = static_cast<cudaStream_t>(a->stream);
    SPHX_CUDA(sphx::launchMomentumEnergy(*a, w.layout, s));
    if (r)
    {
        // the last loop of the step: its scalars are the step's, and the sticky error flags of the search (traversal
        // overflow, h-iteration, ngmax, candidate space) are reported here for callers that issued the loops one by one
        sphx::StepScalars h;
        if (int rc = readScalars(w, s, h)) return rc;
        fillResult(h, a->p, r);
        return errFlagsToStatus(h.errFlags);
    }
    return SPHX_OK;
}

int sphx_hydro_step(const SphxStepArgs* a, SphxHaloExchangeFn halo, void* haloUser, SphxStepResult* r)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    const SphxFields& f = a->f;
    auto exchange = [&](std::initializer_list<std::pair<void*, int>> arrs) -> int
    {
        if (!halo) return SPHX_OK;
        void* ptrs[8];
        int   bytes[8];
        int   n = 0;
        for (auto& p : arrs)
        {
            ptrs[n]    = p.first;
            bytes[n++] = p.second;
        }
        int rc = halo(haloUser, n, ptrs, bytes);
        return rc ? fail(SPHX_ERR_NCCL, "halo exchange callback failed") : SPHX_OK;
    };

    // Everything the six loops REQUIRE is checked before the first exchange, so that argument errors (the same on
    // every rank of a multi-GPU step) never leave a peer waiting in a receive. A failure that shows up later on one
    // rank only (CUDA launch error) skips that rank's remaining loops but NOT its remaining exchanges: the collectives
    // of all ranks stay matched, and the caller's reduction of the status (sphx_reduce_step_result) reports it.
    if (int rc = checkTree(a->tree)) return rc;
    if (a->p.ng0 > a->p.ngmax) return fail(SPHX_ERR_INVALID, "ng0 should be smaller than ngmax");
    REQUIRE(f.x); REQUIRE(f.y); REQUIRE(f.z); REQUIRE(f.h); REQUIRE(f.nc); REQUIRE(f.m); REQUIRE(f.xm); REQUIRE(a->wh);
    REQUIRE(a->whd); REQUIRE(f.kx); REQUIRE(f.gradh); REQUIRE(f.prho); REQUIRE(f.c); REQUIRE(f.vx); REQUIRE(f.vy);
    REQUIRE(f.vz); REQUIRE(f.c11); REQUIRE(f.c12); REQUIRE(f.c13); REQUIRE(f.c22); REQUIRE(f.c23); REQUIRE(f.c33);
    REQUIRE(f.divv); REQUIRE(f.alpha); REQUIRE(f.ax); REQUIRE(f.ay); REQUIRE(f.az); REQUIRE(f.du);
    if (a->p.eosChoice == 0 && !f.temp && !f.u) return fail(SPHX_ERR_INVALID, "ideal gas EOS needs temp or u");
    if (a->p.eosChoice < 0 || a->p.eosChoice > 2) return fail(SPHX_ERR_INVALID, "unknown eosChoice");
    if (a->p.avClean)
    {
        REQUIRE(f.dV11); REQUIRE(f.dV12); REQUIRE(f.dV13); REQUIRE(f.dV22); REQUIRE(f.dV23); REQUIRE(f.dV33);
    }

    int         status = SPHX_OK;
    std::string firstError;
    auto        stage = [&](int rc)
    {
        if (rc && !status) status = rc, firstError = g_lastError;
    };

    // ve_hydro.hpp:147-190
    stage(sphx_find_neighbors_xmass(a, nullptr));
    if (int rc = exchange({{f.xm, 4}})) return rc;
    if (!status) stage(sphx_ve_def_gradh(a));
    if (!status) stage(sphx_eos(a));
    if (int rc = exchange({{(void*)f.vx, 4}, {(void*)f.vy, 4}, {(void*)f.vz, 4}, {f.prho, 4}, {f.c, 4}, {f.kx, 4}}))
        return rc;
    if (!status) stage(sphx_iad_divv_curlv(a, nullptr));
    if (int rc = exchange({{f.c11, 4}, {f.c12, 4}, {f.c13, 4}, {f.c22, 4}, {f.c23, 4}, {f.c33, 4}, {f.divv, 4}}))
        return rc;
    if (!status) stage(sphx_av_switches(a));
    if (a->p.avClean)
    {
        if (int rc = exchange({{f.dV11, 4}, {f.dV12, 4}, {f.dV13, 4}, {f.dV22, 4}, {f.dV23, 4}, {f.dV33, 4}, {f.alpha, 4}}))
            return rc;
    }
    else if (int rc = exchange({{f.alpha, 4}})) { return rc; }
    if (!status) stage(sphx_momentum_energy(a, nullptr));
    if (status) return fail(status, firstError);

    if (r)
    {
        sphx::StepScalars h;
        if (int rc = readScalars(w, static_cast<cudaStream_t>(a->stream), h)) return rc;
        fillResult(h, a->p, r);
        return errFlagsToStatus(h.errFlags);
    }
    return SPHX_OK;
}

int sphx_find_neighbors(const double* x, const double* y, const double* z, const float* h, size_t first, size_t last,
                        const SphxBox* box, const SphxTreeView* tree, unsigned ngmax, unsigned* neighbors,
                        unsigned* counts, void* stream)
{
    if (int rc = sphx_device_check()) return rc;
    if (!x || !y || !z || !h || !box || !tree || !neighbors || !counts)
        return fail(SPHX_ERR_INVALID, "null argument");
    if (int rc = checkTree(*tree)) return rc;
    if (last <= first) return SPHX_OK;
    auto   s = static_cast<cudaStream_t>(stream);
    size_t n = last - first;
    void*  tmp = nullptr;
    SPHX_CUDA(cudaMallocAsync(&tmp, sphx::kScalarsBytes + genericListBytes(n, ngmax), s));
    auto* scal = reinterpret_cast<sphx::StepScalars*>(tmp);
    auto* list = reinterpret_cast<unsigned*>(static_cast<char*>(tmp) + sphx::kScalarsBytes);
    sphx::launchResetScalars(scal, s);
    sphx::launchFindNeighbors(x, y, z, h, unsigned(first), unsigned(last), *box, *tree, ngmax, list, counts, scal, s);
    sphx::launchExportNeighbors(unsigned(n), ngmax, list, counts, false, neighbors, s);
    sphx::StepScalars hs;
    cudaError_t       e = cudaMemcpyAsync(&hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFreeAsync(tmp, s);
    if (e != cudaSuccess) return cudaFail(e, "sphx_find_neighbors");
    SPHX_CUDA(cudaGetLastError());
    if (hs.errFlags & sphx::kErrTraversal)
        return fail(SPHX_ERR_TRAVERSAL, "GPU traversal stack exhausted in neighbor search");
    return SPHX_OK;
}

int sphx_export_neighbors(const SphxStepArgs* a, unsigned* neighbors_dev)
{
    if (int rc = sphx_device_check()) return rc;
    Workspace w;
    if (int rc = carve(a, w)) return rc;
    REQUIRE(neighbors_dev); REQUIRE(a->f.nc);
    sphx::launchExportBlockNeighbors(*a, w.layout, neighbors_dev, static_cast<cudaStream_t>(a->stream));
    SPHX_CUDA(cudaGetLastError());
    return SPHX_OK;
}

} // extern "C"

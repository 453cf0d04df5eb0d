/*! @file
 * integrate() and the conserved-quantity reduction: the streaming kernels that run between
 * two hydro steps, so that a whole time-step loop stays on the device.
 *
 * Replaces (reference paths relative to /root/reference):
 *   sph::computeTimestep          sph/include/sph/ts_global.hpp:97-113
 *   sph::computePositions         sph/include/sph/positions.hpp:57-63,74-86,89-150,177-200; positions_gpu.cu:119-170
 *   sph::updateSmoothingLength    sph/include/sph/update_h.hpp:44-53, update_h_gpu.cu:40-49, kernels.hpp:26-32
 *   computeConservedQuantities    main/src/observables/conserved_quantities.hpp:49-177, conserved_gpu.cu:60-100
 *
 * Arithmetic follows the reference CPU path (the parity target): positions and velocities are advanced in fp64 from
 * the fp32 fields, every operation rounded separately (the oracle build uses -ffp-contract=off), results rounded to
 * the storage types on store. Both per-particle updates are pure streaming work (HBM-bound): sphx_integrate fuses
 * them into one pass, 116 B read + 80 B written per particle.
 */
#include <cmath>
#include <string>

#include "sphx_device.cuh"
#include "sphx_kernels.h"

namespace sphx
{
namespace
{

struct IntegrateDev
{
    SphxIntegrateArgs a;
    DevBox            box;
    int               fbc[3];
    float             cv; // idealGasCv(muiConst, gamma): float (sph/eos.hpp:18-23)
};

__device__ __forceinline__ bool fbcCheck(double coord, float h, double top, double bottom, bool fbc)
{
    // std::abs(top - coord) < Th(2) * h: double < float comparison, rhs promoted
    return fbc && (fabs(__dsub_rn(top, coord)) < double(2.0f * h) || fabs(__dsub_rn(bottom, coord)) < double(2.0f * h));
}

//! positionUpdate (positions.hpp:74-86) for one component; returns X_n+1 (before putInBox), V_n+1, dX_n+1
__device__ __forceinline__ void positionUpdate1(double dt, double dt_m1, double Xn, double An, double dXn, double& Xnp1,
                                                double& Vnp1, double& dXnp1)
{
    double Vnmhalf = __dmul_rn(dXn, __ddiv_rn(1.0, dt_m1));
    double Vn      = __dadd_rn(Vnmhalf, __dmul_rn(An, __dmul_rn(0.5, dt_m1))); // (T(0.5) * dt_m1) * An
    Vnp1           = __dadd_rn(Vn, __dmul_rn(An, dt));
    dXnp1          = __dmul_rn(__dadd_rn(Vn, __dmul_rn(__dmul_rn(An, 0.5), fabs(dt))), dt); // ((T(0.5) * An) * |dt|)
    Xnp1           = __dadd_rn(Xn, dXnp1);
}

__device__ __forceinline__ double putInBox1(double X, double lo, double hi, double l, int pbc)
{
    if (pbc && X > hi) { X = __dsub_rn(X, l); }
    else if (pbc && X < lo) { X = __dadd_rn(X, l); }
    return X;
}

//! energyUpdate (positions.hpp:57-63), TU = double
__device__ __forceinline__ double energyUpdate(double u_old, double dt, double dt_m1, double du, double du_m1)
{
    // u_old + du * dt + 0.5 * (du - du_m1) / dt_m1 * std::abs(dt) * dt, left to right
    double t     = __dmul_rn(__dmul_rn(__ddiv_rn(__dmul_rn(0.5, __dsub_rn(du, du_m1)), dt_m1), fabs(dt)), dt);
    double u_new = __dadd_rn(__dadd_rn(u_old, __dmul_rn(du, dt)), t);
    if (u_new < 0.) { u_new = u_old * exp(u_new * dt / u_old); }
    return u_new;
}

template<bool Positions, bool UpdateH>
__global__ void __launch_bounds__(256) integrateKernel(IntegrateDev d)
{
    const SphxIntegrateArgs& a = d.a;
    size_t                   i = a.first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.last) return;

    if (UpdateH) { a.h[i] = updateH(a.ng0, a.nc[i], a.h[i]); }
    if (!Positions) return;

    double X[3]   = {a.x[i], a.y[i], a.z[i]};
    double A[3]   = {double(a.ax[i]), double(a.ay[i]), double(a.az[i])};
    double dXn[3] = {double(a.x_m1[i]), double(a.y_m1[i]), double(a.z_m1[i])};
    double Xn1[3], Vn1[3], dXn1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        positionUpdate1(a.dt, a.dt_m1, X[k], A[k], dXn[k], Xn1[k], Vn1[k], dXn1[k]);
    Xn1[0] = putInBox1(Xn1[0], d.box.xmin, d.box.xmax, d.box.lx, d.box.pbcX);
    Xn1[1] = putInBox1(Xn1[1], d.box.ymin, d.box.ymax, d.box.ly, d.box.pbcY);
    Xn1[2] = putInBox1(Xn1[2], d.box.zmin, d.box.zmax, d.box.lz, d.box.pbcZ);

    a.x[i] = Xn1[0], a.y[i] = Xn1[1], a.z[i] = Xn1[2];
    a.x_m1[i] = float(dXn1[0]), a.y_m1[i] = float(dXn1[1]), a.z_m1[i] = float(dXn1[2]);
    a.vx[i] = float(Vn1[0]), a.vy[i] = float(Vn1[1]), a.vz[i] = float(Vn1[2]);

    if (a.temp)
    {
        double du    = a.du[i];
        double u_old = __dmul_rn(double(d.cv), a.temp[i]);
        a.temp[i]    = __ddiv_rn(energyUpdate(u_old, a.dt, a.dt_m1, du, double(a.du_m1[i])), double(d.cv));
        a.du_m1[i]   = float(du);
    }
    else if (a.u)
    {
        double du  = a.du[i];
        a.u[i]     = energyUpdate(a.u[i], a.dt, a.dt_m1, du, double(a.du_m1[i]));
        a.du_m1[i] = float(du);
    }
}

//! variant for boxes with fixed boundaries: particles at rest inside 2h of a fixed wall do not move
//! (positions.hpp:97-107). Kept out of the common kernel, which then needs no v and h reads for the test.
template<bool UpdateH>
__global__ void __launch_bounds__(256) integrateFbcKernel(IntegrateDev d)
{
    const SphxIntegrateArgs& a = d.a;
    size_t                   i = a.first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.last) return;

    float hOld = a.h[i];
    if (UpdateH) { a.h[i] = updateH(a.ng0, a.nc[i], hOld); }

    double X[3] = {a.x[i], a.y[i], a.z[i]};
    if (a.vx[i] == 0.0f && a.vy[i] == 0.0f && a.vz[i] == 0.0f)
    {
        if (fbcCheck(X[0], hOld, d.box.xmax, d.box.xmin, d.fbc[0]) ||
            fbcCheck(X[1], hOld, d.box.ymax, d.box.ymin, d.fbc[1]) ||
            fbcCheck(X[2], hOld, d.box.zmax, d.box.zmin, d.fbc[2]))
        {
            return; // position, velocity and x_m1 stay; the energy update still runs (energyUpdateKernel), as in
                    // updateTempHost / updateIntEnergyHost of the CPU path, which cover the whole range
        }
    }
    double A[3]   = {double(a.ax[i]), double(a.ay[i]), double(a.az[i])};
    double dXn[3] = {double(a.x_m1[i]), double(a.y_m1[i]), double(a.z_m1[i])};
    double Xn1[3], Vn1[3], dXn1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        positionUpdate1(a.dt, a.dt_m1, X[k], A[k], dXn[k], Xn1[k], Vn1[k], dXn1[k]);
    Xn1[0] = putInBox1(Xn1[0], d.box.xmin, d.box.xmax, d.box.lx, d.box.pbcX);
    Xn1[1] = putInBox1(Xn1[1], d.box.ymin, d.box.ymax, d.box.ly, d.box.pbcY);
    Xn1[2] = putInBox1(Xn1[2], d.box.zmin, d.box.zmax, d.box.lz, d.box.pbcZ);
    a.x[i] = Xn1[0], a.y[i] = Xn1[1], a.z[i] = Xn1[2];
    a.x_m1[i] = float(dXn1[0]), a.y_m1[i] = float(dXn1[1]), a.z_m1[i] = float(dXn1[2]);
    a.vx[i] = float(Vn1[0]), a.vy[i] = float(Vn1[1]), a.vz[i] = float(Vn1[2]);
}

//! energy update of the fbc variant for ALL particles (updateTempHost / updateIntEnergyHost run over the full range)
__global__ void __launch_bounds__(256) energyUpdateKernel(IntegrateDev d)
{
    const SphxIntegrateArgs& a = d.a;
    size_t                   i = a.first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.last) return;
    if (a.temp)
    {
        double du    = a.du[i];
        double u_old = __dmul_rn(double(d.cv), a.temp[i]);
        a.temp[i]    = __ddiv_rn(energyUpdate(u_old, a.dt, a.dt_m1, du, double(a.du_m1[i])), double(d.cv));
        a.du_m1[i]   = float(du);
    }
    else if (a.u)
    {
        double du  = a.du[i];
        a.u[i]     = energyUpdate(a.u[i], a.dt, a.dt_m1, du, double(a.du_m1[i]));
        a.du_m1[i] = float(du);
    }
}

/* ------------------------------------------------ conserved quantities ------------------------------------------------ */

constexpr int kConsBlocks  = 592; // 4 per SM on 148 SMs
constexpr int kConsThreads = 256;
constexpr int kConsValues  = 10; // eKin2, eInt, px, py, pz, Lx, Ly, Lz, ncsum, unused

struct ConsArgs
{
    const double *  x, *y, *z;
    const float *   vx, *vy, *vz, *m;
    const double *  temp, *u;
    const unsigned* nc;
    size_t          first, last;
    double          cv;
};

__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(kFullMask, v, o);
    return v;
}

//! block partial sums in a fixed order: thread-strided accumulation, warp tree, then warps in order
__global__ void __launch_bounds__(kConsThreads) conservedPartialKernel(ConsArgs c, double* __restrict__ partial)
{
    double acc[kConsValues] = {};
    size_t stride           = size_t(gridDim.x) * blockDim.x;
    for (size_t i = c.first + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < c.last; i += stride)
    {
        double X[3] = {c.x[i], c.y[i], c.z[i]};
        double V[3] = {double(c.vx[i]), double(c.vy[i]), double(c.vz[i])};
        float  mf   = c.m[i];
        double mi   = double(mf);
        // mi * norm2(V): float m promoted, norm2 right fold (util/array.hpp:236-240)
        acc[0] += mi * (V[0] * V[0] + (V[1] * V[1] + V[2] * V[2]));
        acc[2] += mi * V[0], acc[3] += mi * V[1], acc[4] += mi * V[2];
        acc[5] += mi * (X[1] * V[2] - X[2] * V[1]);
        acc[6] += mi * (X[2] * V[0] - X[0] * V[2]);
        acc[7] += mi * (X[0] * V[1] - X[1] * V[0]);
        if (c.u) { acc[1] += c.u[i] * mi; }
        else if (c.temp) { acc[1] += c.cv * c.temp[i] * mi; }
        if (c.nc) { acc[8] += double(c.nc[i]); }
    }
    __shared__ double sm[kConsThreads / 32][kConsValues];
    int               lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kConsValues; ++k)
    {
        double v = warpSum(acc[k]);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kConsValues)
    {
        double v = 0;
        for (int w = 0; w < kConsThreads / 32; ++w)
            v += sm[w][threadIdx.x];
        partial[blockIdx.x * kConsValues + threadIdx.x] = v;
    }
}

__global__ void conservedFinalKernel(const double* __restrict__ partial, int numBlocks, double* __restrict__ out)
{
    // one warp per value, lanes stride over the block partials, fixed tree
    int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= kConsValues) return;
    double v = 0;
    for (int b = lane; b < numBlocks; b += 32)
        v += partial[b * kConsValues + k];
    v = warpSum(v);
    if (lane == 0) out[k] = v;
}

int intFail(int code, const std::string& msg)
{
    setLastError(msg);
    return code;
}

#define INT_CUDA(call)                                                                                                 \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) return intFail(SPHX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

float idealGasCvHost(float mui, double gamma) { return float(double(8.317e7f / mui) / (gamma - double(1.0f))); }

int launchIntegrate(const SphxIntegrateArgs* a, bool positions, bool updateH)
{
    if (int st = sphx_device_check()) return st;
    if (!a) return intFail(SPHX_ERR_INVALID, "null args");
    if (a->last < a->first) return intFail(SPHX_ERR_INVALID, "bad [first,last) range");
    if (positions && (!a->x || !a->y || !a->z || !a->x_m1 || !a->y_m1 || !a->z_m1 || !a->vx || !a->vy || !a->vz ||
                      !a->ax || !a->ay || !a->az))
        return intFail(SPHX_ERR_INVALID, "sphx_compute_positions: required field is NULL");
    if (positions && (a->temp || a->u) && (!a->du || !a->du_m1))
        return intFail(SPHX_ERR_INVALID, "sphx_compute_positions: du / du_m1 are NULL");
    if (updateH && (!a->h || !a->nc)) return intFail(SPHX_ERR_INVALID, "sphx_update_smoothing_length: h / nc are NULL");
    size_t n = a->last - a->first;
    if (n == 0) return SPHX_OK;

    IntegrateDev d;
    d.a   = *a;
    d.box = makeDevBox(a->box);
    for (int k = 0; k < 3; ++k)
        d.fbc[k] = a->box.boundary[k] == 2;
    d.cv        = idealGasCvHost(a->muiConst, a->gamma);
    bool anyFbc = d.fbc[0] || d.fbc[1] || d.fbc[2];
    auto s      = static_cast<cudaStream_t>(a->stream);
    unsigned grid = unsigned((n + 255) / 256);
    if (positions && anyFbc)
    {
        if (!a->h) return intFail(SPHX_ERR_INVALID, "sphx_compute_positions: h is NULL (fixed boundaries)");
        if (updateH) { integrateFbcKernel<true><<<grid, 256, 0, s>>>(d); }
        else { integrateFbcKernel<false><<<grid, 256, 0, s>>>(d); }
        if (a->temp || a->u) energyUpdateKernel<<<grid, 256, 0, s>>>(d);
    }
    else if (positions && updateH) { integrateKernel<true, true><<<grid, 256, 0, s>>>(d); }
    else if (positions) { integrateKernel<true, false><<<grid, 256, 0, s>>>(d); }
    else if (updateH) { integrateKernel<false, true><<<grid, 256, 0, s>>>(d); }
    INT_CUDA(cudaGetLastError());
    return SPHX_OK;
}

} // namespace
} // namespace sphx

extern "C"
{

int sphx_compute_positions(const SphxIntegrateArgs* a) { return sphx::launchIntegrate(a, true, false); }
int sphx_update_smoothing_length(const SphxIntegrateArgs* a) { return sphx::launchIntegrate(a, false, true); }
int sphx_integrate(const SphxIntegrateArgs* a) { return sphx::launchIntegrate(a, true, true); }

int sphx_compute_timestep(double minDtCourant, double minDtRho, double maxDtIncrease, double* minDt, double* minDt_m1,
                          double* ttot, SphxComm* comm, void* stream)
{
    if (!minDt || !minDt_m1 || !ttot) return sphx::intFail(SPHX_ERR_INVALID, "sphx_compute_timestep: null argument");
    // std::min({minDtAcc = INFINITY, minDtCourant, minDtRho, maxDtIncrease * minDt}) (ts_global.hpp:104)
    double loc = std::fmin(std::fmin(minDtCourant, minDtRho), maxDtIncrease * *minDt);
    if (comm)
    {
        if (int st = sphx_allreduce_f64(comm, &loc, 1, 0, stream)) return st;
    }
    *ttot += loc;
    *minDt_m1 = *minDt;
    *minDt    = loc;
    return SPHX_OK;
}

size_t sphx_conserved_scratch_bytes(void) { return size_t(sphx::kConsBlocks + 2) * sphx::kConsValues * sizeof(double); }

int sphx_conserved_quantities(const double* x, const double* y, const double* z, const float* vx, const float* vy,
                              const float* vz, const float* m, const double* temp, const double* u,
                              const unsigned* nc, size_t first, size_t last, double gamma, float muiConst, double egrav,
                              void* scratch, SphxComm* comm, void* stream, SphxConserved* out)
{
    using namespace sphx;
    if (int st = sphx_device_check()) return st;
    if (!x || !y || !z || !vx || !vy || !vz || !m || !scratch || !out || last < first)
        return intFail(SPHX_ERR_INVALID, "sphx_conserved_quantities: bad argument");
    auto     s       = static_cast<cudaStream_t>(stream);
    double*  partial = static_cast<double*>(scratch);
    double*  result  = partial + size_t(kConsBlocks) * kConsValues;
    ConsArgs c{x, y, z, vx, vy, vz, m, temp, u, nc, first, last, double(idealGasCvHost(muiConst, gamma))};
    size_t   n      = last - first;
    int      blocks = int(std::min<size_t>(kConsBlocks, (n + kConsThreads - 1) / kConsThreads));
    if (blocks < 1) blocks = 1;
    conservedPartialKernel<<<blocks, kConsThreads, 0, s>>>(c, partial);
    conservedFinalKernel<<<1, 32 * kConsValues, 0, s>>>(partial, blocks, result);
    double q[kConsValues];
    INT_CUDA(cudaMemcpyAsync(q, result, sizeof(q), cudaMemcpyDeviceToHost, s));
    INT_CUDA(cudaStreamSynchronize(s));
    q[9] = egrav;
    if (comm)
    {
        if (int st = sphx_allreduce_f64(comm, q, kConsValues, 2, stream)) return st;
    }
    out->ecin  = 0.5 * q[0];
    out->eint  = q[1];
    out->egrav = q[9];
    out->etot  = out->ecin + out->eint + out->egrav;
    for (int k = 0; k < 3; ++k)
    {
        out->linmom3[k] = q[2 + k];
        out->angmom3[k] = q[5 + k];
    }
    auto norm3 = [](const double* v) { return std::sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2])); };
    out->linmom         = norm3(out->linmom3);
    out->angmom         = norm3(out->angmom3);
    out->totalNeighbors = (unsigned long)(q[8]);
    return SPHX_OK;
}

} // extern "C"

/*! @file
 * Device half of the turbulence driver: stirring accelerations.
 *
 * Replaces (reference paths relative to /root/reference):
 *   sph::driveTurbulence          sph/include/sph/hydro_turb/driver.hpp:102-128
 *   sph::computeStirringGpu       sph/include/sph/hydro_turb/stirring_gpu.cu:43-74
 *   sph::stirParticle             sph/include/sph/hydro_turb/stirring.hpp:45-83
 *
 * The reference evaluates six fp64 sin/cos per particle and mode (112 modes with the default settings). The stirring
 * modes are lattice wave vectors k = 2 pi (i, j, l) / L with |i|, |j|, |l| <= 3, so a particle needs only
 * cos/sin(2 pi i x / L) for i = 1..3 per coordinate: `stirKernel<true>` evaluates those 9 sincos once per particle,
 * keeps them in a per-thread shared-memory column and forms every mode's trigonometric terms from them with the
 * reference's own expressions (same operation order, each operation rounded separately, accumulation in fp32 as
 * `Ta turbAx += ...` does). The values are those of the reference up to the last-bit differences between CUDA's and
 * glibc's sin/cos. `stirKernel<false>` is the mode-by-mode form for mode sets that are not on the lattice (it is what
 * the reference does). fp64-pipe bound: ~45 fp64 operations + 6 conversions per particle and mode; the compulsory HBM
 * traffic is 48 B per particle (x, y, z in; ax, ay, az in/out).
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <string>

#include "sphx_kernels.h"
#include "sphx_turbulence.h"

namespace sphx
{
namespace
{

constexpr int kStirThreads = 128;

struct StirArgs
{
    const double* x;
    const double* y;
    const double* z;
    float*        ax;
    float*        ay;
    float*        az;
    size_t        first, last;
    unsigned      numModes;
    int           maxIdx;
    const double* modes;
    const double* amplitudes;
    const double* phaseReal;
    const double* phaseImag;
    const int8_t* latticeIdx;
    double        twopi, L;
    double        solWeightNorm;
};

//! the per-mode part of stirParticle (stirring.hpp:65-79), every operation rounded separately
__device__ __forceinline__ void addMode(double cx, double sx, double cy, double sy, double cz, double sz, double amp,
                                        const double* pr, const double* pi, float& tx, float& ty, float& tz)
{
    double re = __dsub_rn(__dmul_rn(__dsub_rn(__dmul_rn(cx, cy), __dmul_rn(sx, sy)), cz),
                          __dmul_rn(__dadd_rn(__dmul_rn(sx, cy), __dmul_rn(cx, sy)), sz));
    double im = __dadd_rn(__dmul_rn(cx, __dadd_rn(__dmul_rn(cy, sz), __dmul_rn(sy, cz))),
                          __dmul_rn(sx, __dsub_rn(__dmul_rn(cy, cz), __dmul_rn(sy, sz))));
    // Ta += T: the sum is formed in double and rounded to float
    tx = float(__dadd_rn(double(tx), __dmul_rn(amp, __dsub_rn(__dmul_rn(pr[0], re), __dmul_rn(pi[0], im)))));
    ty = float(__dadd_rn(double(ty), __dmul_rn(amp, __dsub_rn(__dmul_rn(pr[1], re), __dmul_rn(pi[1], im)))));
    tz = float(__dadd_rn(double(tz), __dmul_rn(amp, __dsub_rn(__dmul_rn(pr[2], re), __dmul_rn(pi[2], im)))));
}

template<bool Lattice>
__global__ void __launch_bounds__(kStirThreads) stirKernel(StirArgs a)
{
    // mode tables: [modes 3n | amplitudes n | phaseReal 3n | phaseImag 3n] doubles, then (Lattice) the trig columns
    extern __shared__ double smem[];
    const unsigned           nm   = a.numModes;
    double*                  sAmp = smem;
    double*                  sPr  = sAmp + nm;
    double*                  sPi  = sPr + 3 * nm;
    double*                  sK   = sPi + 3 * nm; // !Lattice only
    double*                  trig = sPi + 3 * nm; // Lattice only: [(dim * maxIdx + (i-1)) * 2 + {cos, sin}][thread]
    int8_t*                  sIdx = reinterpret_cast<int8_t*>(trig + 6 * a.maxIdx * kStirThreads); // Lattice only

    for (unsigned k = threadIdx.x; k < nm; k += kStirThreads)
        sAmp[k] = a.amplitudes[k];
    for (unsigned k = threadIdx.x; k < 3 * nm; k += kStirThreads)
    {
        sPr[k] = a.phaseReal[k];
        sPi[k] = a.phaseImag[k];
        if (!Lattice) sK[k] = a.modes[k];
        if (Lattice) sIdx[k] = a.latticeIdx[k];
    }
    __syncthreads();

    size_t i     = a.first + size_t(blockIdx.x) * kStirThreads + threadIdx.x;
    bool   alive = i < a.last;
    double X[3]  = {0, 0, 0};
    if (alive) { X[0] = a.x[i], X[1] = a.y[i], X[2] = a.z[i]; }

    float tx = 0.f, ty = 0.f, tz = 0.f;
    if (Lattice)
    {
        for (int d = 0; d < 3; ++d)
            for (int q = 1; q <= a.maxIdx; ++q)
            {
                // modes[] holds twopi * q / L (create_modes.hpp:79-86); the argument is that product times the coordinate
                double kq = __ddiv_rn(__dmul_rn(a.twopi, double(q)), a.L);
                double s, c;
                sincos(__dmul_rn(kq, X[d]), &s, &c);
                trig[((d * a.maxIdx + q - 1) * 2 + 0) * kStirThreads + threadIdx.x] = c;
                trig[((d * a.maxIdx + q - 1) * 2 + 1) * kStirThreads + threadIdx.x] = s;
            }
        auto fetch = [&](int d, int q, double& c, double& s)
        {
            // cos(-t) = cos(t), sin(-t) = -sin(t) exactly (both libraries are odd/even symmetric); q == 0: cos 1, sin +-0
            int aq = q < 0 ? -q : q;
            if (aq == 0) { c = 1.0, s = 0.0; }
            else
            {
                c = trig[((d * a.maxIdx + aq - 1) * 2 + 0) * kStirThreads + threadIdx.x];
                s = trig[((d * a.maxIdx + aq - 1) * 2 + 1) * kStirThreads + threadIdx.x];
                if (q < 0) s = -s;
            }
        };
        for (unsigned m = 0; m < nm; ++m)
        {
            double cx, sx, cy, sy, cz, sz;
            fetch(0, sIdx[3 * m + 0], cx, sx);
            fetch(1, sIdx[3 * m + 1], cy, sy);
            fetch(2, sIdx[3 * m + 2], cz, sz);
            addMode(cx, sx, cy, sy, cz, sz, sAmp[m], sPr + 3 * m, sPi + 3 * m, tx, ty, tz);
        }
    }
    else
    {
        for (unsigned m = 0; m < nm; ++m)
        {
            double cx, sx, cy, sy, cz, sz;
            sincos(__dmul_rn(sK[3 * m + 2], X[2]), &sz, &cz);
            sincos(__dmul_rn(sK[3 * m + 1], X[1]), &sy, &cy);
            sincos(__dmul_rn(sK[3 * m + 0], X[0]), &sx, &cx);
            addMode(cx, sx, cy, sy, cz, sz, sAmp[m], sPr + 3 * m, sPi + 3 * m, tx, ty, tz);
        }
    }

    if (alive)
    {
        // ax[i] += solWeightNorm * turbAx: float += double (stirring.hpp:118-120)
        a.ax[i] = float(__dadd_rn(double(a.ax[i]), __dmul_rn(a.solWeightNorm, double(tx))));
        a.ay[i] = float(__dadd_rn(double(a.ay[i]), __dmul_rn(a.solWeightNorm, double(ty))));
        a.az[i] = float(__dadd_rn(double(a.az[i]), __dmul_rn(a.solWeightNorm, double(tz))));
    }
}

int cudaFail(cudaError_t e, const char* what)
{
    setLastError(std::string(what) + ": " + cudaGetErrorString(e));
    return SPHX_ERR_CUDA;
}

#define TURB_CUDA(call)                                                                                                \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) return cudaFail(e_, #call);                                                             \
    } while (0)

int uploadStatic(SphxTurbulence& t, cudaStream_t s)
{
    if (t.uploaded) return SPHX_OK;
    size_t nm = t.numModes;
    TURB_CUDA(cudaMalloc(&t.d_modes, std::max<size_t>(1, 3 * nm) * sizeof(double)));
    TURB_CUDA(cudaMalloc(&t.d_amplitudes, std::max<size_t>(1, nm) * sizeof(double)));
    TURB_CUDA(cudaMalloc(&t.d_phases, std::max<size_t>(1, 6 * nm) * sizeof(double)));
    TURB_CUDA(cudaMalloc(&t.d_latticeIdx, std::max<size_t>(1, 3 * nm)));
    TURB_CUDA(cudaMemcpyAsync(t.d_modes, t.modes.data(), 3 * nm * sizeof(double), cudaMemcpyHostToDevice, s));
    TURB_CUDA(cudaMemcpyAsync(t.d_amplitudes, t.amplitudes.data(), nm * sizeof(double), cudaMemcpyHostToDevice, s));
    TURB_CUDA(cudaMemcpyAsync(t.d_latticeIdx, t.latticeIdx.data(), 3 * nm, cudaMemcpyHostToDevice, s));
    t.uploaded = true;
    return SPHX_OK;
}

int stir(SphxTurbulence& t, const double* x, const double* y, const double* z, float* ax, float* ay, float* az,
         size_t first, size_t last, cudaStream_t s)
{
    if (int rc = uploadStatic(t, s)) return rc;
    size_t nm = t.numModes;
    // pageable source: the runtime stages the bytes before returning, so the host vectors may change right away
    TURB_CUDA(cudaMemcpyAsync(t.d_phases, t.phasesReal.data(), 3 * nm * sizeof(double), cudaMemcpyHostToDevice, s));
    TURB_CUDA(cudaMemcpyAsync(t.d_phases + 3 * nm, t.phasesImag.data(), 3 * nm * sizeof(double), cudaMemcpyHostToDevice,
                              s));
    if (last <= first || nm == 0) return SPHX_OK;

    StirArgs a{x, y, z, ax, ay, az, first, last, unsigned(nm), t.maxIdx, t.d_modes, t.d_amplitudes, t.d_phases,
               t.d_phases + 3 * nm, t.d_latticeIdx, 2.0 * M_PI, t.Lbox, t.solWeightNorm};
    unsigned grid = unsigned((last - first + kStirThreads - 1) / kStirThreads);
    if (t.lattice)
    {
        size_t bytes = (7 * nm + size_t(3 * t.maxIdx * 2) * kStirThreads) * sizeof(double) + 3 * nm;
        if (bytes > 200 * 1024)
        {
            setLastError("turbulence: mode tables exceed the shared memory of one CTA");
            return SPHX_ERR_INVALID;
        }
        TURB_CUDA(cudaFuncSetAttribute(stirKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
        stirKernel<true><<<grid, kStirThreads, bytes, s>>>(a);
    }
    else
    {
        size_t bytes = 10 * nm * sizeof(double);
        if (bytes > 200 * 1024)
        {
            setLastError("turbulence: mode tables exceed the shared memory of one CTA");
            return SPHX_ERR_INVALID;
        }
        TURB_CUDA(cudaFuncSetAttribute(stirKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
        stirKernel<false><<<grid, kStirThreads, bytes, s>>>(a);
    }
    TURB_CUDA(cudaGetLastError());
    return SPHX_OK;
}

} // namespace

void turbulenceFreeDevice(SphxTurbulence& t)
{
    if (!t.uploaded) return;
    cudaFree(t.d_modes);
    cudaFree(t.d_amplitudes);
    cudaFree(t.d_phases);
    cudaFree(t.d_latticeIdx);
    t.uploaded = false;
}

} // namespace sphx

extern "C"
{

int sphx_compute_stirring(SphxTurbulence* t, const double* x, const double* y, const double* z, float* ax, float* ay,
                          float* az, size_t first, size_t last, void* stream)
{
    if (int st = sphx_device_check()) return st;
    if (!t || !x || !y || !z || !ax || !ay || !az || last < first)
    {
        sphx::setLastError("sphx_compute_stirring: null argument");
        return SPHX_ERR_INVALID;
    }
    return sphx::stir(*t, x, y, z, ax, ay, az, first, last, static_cast<cudaStream_t>(stream));
}

int sphx_drive_turbulence(SphxTurbulence* t, const double* x, const double* y, const double* z, float* ax, float* ay,
                          float* az, size_t first, size_t last, double minDt, void* stream)
{
    if (int st = sphx_device_check()) return st;
    if (!t || !x || !y || !z || !ax || !ay || !az || last < first)
    {
        sphx::setLastError("sphx_drive_turbulence: null argument");
        return SPHX_ERR_INVALID;
    }
    sphx_turbulence_advance_host(t, minDt);
    return sphx::stir(*t, x, y, z, ax, ay, az, first, last, static_cast<cudaStream_t>(stream));
}

} // extern "C"

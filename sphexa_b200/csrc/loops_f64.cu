/*! @file
 * The SPH-VE hydro step in the ALL-DOUBLE type set: every field and every operation in fp64 (the instantiation the
 * reference's own unit tests use, sph/test/ve.cpp: `using T = double`), for the <= 1e-10 clause of the parity contract.
 *
 * Replaces, for Tc = T = Tm = double (reference paths relative to /root/reference/sph/include/sph):
 *   find_neighbors.hpp:11-44 + cstone findneighbors.hpp:77-147     search with the coupled h-iteration (neighbors.cu)
 *   hydro_ve/xmass_kern.hpp:51-79, ve_def_gradh_kern.hpp:44-90, eos.hpp:18-86, iad_kern.hpp:44-109,
 *   divv_curlv_kern.hpp:44-123, av_switches_kern.hpp:44-137, momentum_energy_kern.hpp:43-222, kernels.hpp:10-16,70-95,
 *   table_lookup.hpp:13-26, ts_global.hpp:72-95
 *
 * This is the precision path, not the fast one: one thread per target walks its particle-index list (the lane-interleaved
 * list of the generic search) and evaluates every pair from the fp64 coordinates of the two particles with the reference's
 * per-pair PBC fold (box.hpp:282-304); nothing is staged, nothing is approximated (the Atwood ramp is four double pow as in
 * the reference). The production type set runs through search.cu / loops.cu; the two paths share no pair code, which is
 * what makes comparing them meaningful (tests/test_gpu_f64.py).
 */
#include <string>

#include "sphx_block.cuh"
#include "sphx_kernels.h"

namespace sphx
{
namespace f64
{

struct Args
{
    SphxFieldsF64 f;
    unsigned      first, last, ngmax;
    DevBox        box;
    const unsigned* list; // lane-interleaved particle indices (nbListIndex)
    const double*   wh;
    const double*   whd;
    StepScalarsF64* scal;
    double          K, minDt, Kcour, gamma, muiConst, alphamin, alphamax, decay_constant, Atmin, Atmax, ramp;
    int             avClean;
};

//! lt::lookup (table_lookup.hpp:13-26), T = double
__device__ __forceinline__ double lookup(const double* __restrict__ table, double v)
{
    constexpr int    numIntervals = kTableSize - 1;
    constexpr double dx           = 2.0 / numIntervals;
    constexpr double invDx        = 1.0 / dx;
    const int        idx          = int(v * invDx);
    if (idx >= numIntervals) return 0.0;
    const double derivative = (table[idx + 1] - table[idx]) * invDx;
    return table[idx] + derivative * (v - double(idx) * dx);
}

//! legacy per-pair PBC of the J-loops (box.hpp:282-304)
__device__ __forceinline__ void applyPbc(const DevBox& b, double r, double& xx, double& yy, double& zz)
{
    if (b.pbcX && xx > r) xx -= b.lx;
    else if (b.pbcX && xx < -r) xx += b.lx;
    if (b.pbcY && yy > r) yy -= b.ly;
    else if (b.pbcY && yy < -r) yy += b.ly;
    if (b.pbcZ && zz > r) zz -= b.lz;
    else if (b.pbcZ && zz < -r) zz += b.lz;
}

struct Pair
{
    double rx, ry, rz, dist;
    unsigned j;
};

//! common frame of the six loops: thread = target, f(Pair) per neighbour
template<class F>
__device__ __forceinline__ void forNeighbours(const Args& a, unsigned i, F&& f)
{
    const unsigned t = i - a.first, g = t / kGroupSize, lane = t % kGroupSize;
    const unsigned n = min(a.f.nc[i] - 1u, a.ngmax);
    const double   xi = a.f.x[i], yi = a.f.y[i], zi = a.f.z[i], twoH = 2.0 * a.f.h[i];
    for (unsigned k = 0; k < n; ++k)
    {
        Pair p;
        p.j  = a.list[nbListIndex(g, a.ngmax, k, lane)];
        p.rx = xi - a.f.x[p.j], p.ry = yi - a.f.y[p.j], p.rz = zi - a.f.z[p.j];
        applyPbc(a.box, twoH, p.rx, p.ry, p.rz);
        p.dist = sqrt(p.rx * p.rx + p.ry * p.ry + p.rz * p.rz);
        f(p);
    }
}

#define SPHX_F64_TARGET                                                                                                \
    const unsigned i = a.first + blockIdx.x * blockDim.x + threadIdx.x;                                                \
    if (i >= a.last) return;

// xmass_kern.hpp:51-79
__global__ void xmassKernel(const __grid_constant__ Args a)
{
    SPHX_F64_TARGET
    const double hi = a.f.h[i], mi = a.f.m[i], hInv = 1.0 / hi, h3Inv = hInv * hInv * hInv;
    double       rho0i = mi;
    forNeighbours(a, i, [&](const Pair& p) { rho0i += lookup(a.wh, p.dist * hInv) * a.f.m[p.j]; });
    a.f.xm[i] = mi / (rho0i * a.K * h3Inv);
}

// ve_def_gradh_kern.hpp:44-90
__global__ void gradhKernel(const __grid_constant__ Args a)
{
    SPHX_F64_TARGET
    const double hi = a.f.h[i], mi = a.f.m[i], xmassi = a.f.xm[i], hInv = 1.0 / hi, h3Inv = hInv * hInv * hInv;
    double       kxi = xmassi, whomegai = -3.0 * xmassi, wrho0i = -3.0 * mi;
    forNeighbours(a, i,
                  [&](const Pair& p)
                  {
                      const double vloc = p.dist * hInv, w = lookup(a.wh, vloc), dw = lookup(a.whd, vloc);
                      const double dterh = -(3.0 * w + vloc * dw), xmassj = a.f.xm[p.j];
                      kxi += w * xmassj;
                      whomegai += dterh * xmassj;
                      wrho0i += dterh * a.f.m[p.j];
                  });
    kxi *= a.K * h3Inv;
    whomegai *= a.K * h3Inv * hInv;
    wrho0i *= a.K * h3Inv * hInv;
    whomegai            = whomegai * mi / xmassi + (kxi - a.K * xmassi * h3Inv) * wrho0i;
    const double rhoi   = kxi * mi / xmassi;
    const double dhdrho = -hi / (rhoi * 3.0);
    a.f.kx[i]           = kxi;
    a.f.gradh[i]        = 1.0 - dhdrho * whomegai;
}

// hydro_ve/eos.hpp:52-197 with eos.hpp:18-86
__global__ void eosKernel(const __grid_constant__ Args a, int eosChoice, double soundSpeedConst, double polyK,
                          double polyIdx)
{
    SPHX_F64_TARGET
    const double kxi = a.f.kx[i], mi = a.f.m[i], rho = kxi * mi / a.f.xm[i];
    double       p, c;
    if (eosChoice == 0)
    {
        double tmp;
        if (a.f.u) { tmp = a.f.u[i] * (a.gamma - 1.0); }
        else
        {
            const double cv = 8.317e7 / a.muiConst / (a.gamma - 1.0);
            tmp             = cv * a.f.temp[i] * (a.gamma - 1.0);
        }
        p = rho * tmp, c = sqrt(a.gamma * tmp);
    }
    else if (eosChoice == 1) { p = rho * soundSpeedConst * soundSpeedConst, c = soundSpeedConst; }
    else { p = polyK * pow(rho, polyIdx), c = sqrt(polyIdx * p / rho); }
    a.f.prho[i] = p / (kxi * mi * mi * a.f.gradh[i]);
    a.f.c[i]    = c;
    if (a.f.rho) a.f.rho[i] = rho;
    if (a.f.p) a.f.p[i] = p;
}

__device__ __forceinline__ void atomicMaxDouble(double* addr, double v)
{
    unsigned long long* p   = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long  old = *p, assumed;
    do
    {
        assumed = old;
        if (__longlong_as_double(assumed) >= v) break;
        old = atomicCAS(p, assumed, __double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomicMinDouble(double* addr, double v)
{
    unsigned long long* p   = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long  old = *p, assumed;
    do
    {
        assumed = old;
        if (__longlong_as_double(assumed) <= v) break;
        old = atomicCAS(p, assumed, __double_as_longlong(v));
    } while (assumed != old);
}

// iad_kern.hpp:44-109, then divv_curlv_kern.hpp:44-123 with the tensor just computed; rhoTimestep ts_global.hpp:72-95
__global__ void iadDivvCurlvKernel(const __grid_constant__ Args a)
{
    SPHX_F64_TARGET
    const double hi = a.f.h[i], hInv = 1.0 / hi;
    double       tau11 = 0, tau12 = 0, tau13 = 0, tau22 = 0, tau23 = 0, tau33 = 0;
    forNeighbours(a, i,
                  [&](const Pair& p)
                  {
                      const double w = lookup(a.wh, p.dist * hInv), volj_w = a.f.xm[p.j] / a.f.kx[p.j] * w;
                      tau11 += p.rx * p.rx * volj_w, tau12 += p.rx * p.ry * volj_w, tau13 += p.rx * p.rz * volj_w;
                      tau22 += p.ry * p.ry * volj_w, tau23 += p.ry * p.rz * volj_w, tau33 += p.rz * p.rz * volj_w;
                  });
    auto         getExp = [](double val) { return (val == 0.0 ? 0 : ilogb(val)); };
    const int    expSum = getExp(tau11) + getExp(tau12) + getExp(tau13) + getExp(tau22) + getExp(tau23) + getExp(tau33);
    const double normal = ldexp(1.0, -expSum / 6);
    tau11 *= normal, tau12 *= normal, tau13 *= normal, tau22 *= normal, tau23 *= normal, tau33 *= normal;
    const double det = tau11 * tau22 * tau33 + 2.0 * tau12 * tau23 * tau13 - tau11 * tau23 * tau23 -
                       tau22 * tau13 * tau13 - tau33 * tau12 * tau12;
    const double factor = normal * (hi * hi * hi) / (det * a.K);
    const double c11 = (tau22 * tau33 - tau23 * tau23) * factor, c12 = (tau13 * tau23 - tau33 * tau12) * factor;
    const double c13 = (tau12 * tau23 - tau22 * tau13) * factor, c22 = (tau11 * tau33 - tau13 * tau13) * factor;
    const double c23 = (tau13 * tau12 - tau11 * tau23) * factor, c33 = (tau11 * tau22 - tau12 * tau12) * factor;
    a.f.c11[i] = c11, a.f.c12[i] = c12, a.f.c13[i] = c13, a.f.c22[i] = c22, a.f.c23[i] = c23, a.f.c33[i] = c33;

    const double vxi = a.f.vx[i], vyi = a.f.vy[i], vzi = a.f.vz[i];
    double       dVxx = 0, dVxy = 0, dVxz = 0, dVyx = 0, dVyy = 0, dVyz = 0, dVzx = 0, dVzy = 0, dVzz = 0;
    forNeighbours(a, i,
                  [&](const Pair& p)
                  {
                      const double w  = lookup(a.wh, p.dist * hInv);
                      const double tA1 = -(c11 * p.rx + c12 * p.ry + c13 * p.rz) * w;
                      const double tA2 = -(c12 * p.rx + c22 * p.ry + c23 * p.rz) * w;
                      const double tA3 = -(c13 * p.rx + c23 * p.ry + c33 * p.rz) * w;
                      const double xmj = a.f.xm[p.j];
                      const double ax = (a.f.vx[p.j] - vxi) * xmj, ay = (a.f.vy[p.j] - vyi) * xmj, az = (a.f.vz[p.j] - vzi) * xmj;
                      dVxx += ax * tA1, dVxy += ax * tA2, dVxz += ax * tA3;
                      dVyx += ay * tA1, dVyy += ay * tA2, dVyz += ay * tA3;
                      dVzx += az * tA1, dVzy += az * tA2, dVzz += az * tA3;
                  });
    const double norm_kxi = a.K * (hInv * hInv * hInv) / a.f.kx[i];
    const double divvi    = norm_kxi * (dVxx + dVyy + dVzz);
    a.f.divv[i]           = divvi;
    if (a.f.curlv)
    {
        const double cx = dVzy - dVyz, cy = dVxz - dVzx, cz = dVyx - dVxy;
        a.f.curlv[i]    = norm_kxi * sqrt(cx * cx + cy * cy + cz * cz);
    }
    if (a.avClean && a.f.dV11)
    {
        a.f.dV11[i] = norm_kxi * dVxx, a.f.dV12[i] = norm_kxi * (dVxy + dVyx), a.f.dV13[i] = norm_kxi * (dVxz + dVzx);
        a.f.dV22[i] = norm_kxi * dVyy, a.f.dV23[i] = norm_kxi * (dVyz + dVzy), a.f.dV33[i] = norm_kxi * dVzz;
    }
    atomicMaxDouble(&a.scal->maxDivv, divvi);
}

// av_switches_kern.hpp:44-137
__global__ void avSwitchesKernel(const __grid_constant__ Args a)
{
    SPHX_F64_TARGET
    const double hi = a.f.h[i], hInv = 1.0 / hi, ci = a.f.c[i], divv_i = a.f.divv[i];
    const double vxi = a.f.vx[i], vyi = a.f.vy[i], vzi = a.f.vz[i];
    const double c11 = a.f.c11[i], c12 = a.f.c12[i], c13 = a.f.c13[i], c22 = a.f.c22[i], c23 = a.f.c23[i],
                 c33 = a.f.c33[i];
    const double Kh3 = a.K * hInv * hInv * hInv;
    double       g1 = 0, g2 = 0, g3 = 0, vijsignal_i = 1e-40 * ci;
    forNeighbours(a, i,
                  [&](const Pair& p)
                  {
                      const double vx_ij = vxi - a.f.vx[p.j], vy_ij = vyi - a.f.vy[p.j], vz_ij = vzi - a.f.vz[p.j];
                      const double rv = p.rx * vx_ij + p.ry * vy_ij + p.rz * vz_ij;
                      const double vs = (rv < 0.0) ? ci + a.f.c[p.j] - 3.0 * rv / p.dist : 0.0;
                      vijsignal_i     = fmax(vijsignal_i, vs);
                      const double Wi  = Kh3 * lookup(a.wh, p.dist * hInv);
                      const double tA1 = -(c11 * p.rx + c12 * p.ry + c13 * p.rz) * Wi;
                      const double tA2 = -(c12 * p.rx + c22 * p.ry + c23 * p.rz) * Wi;
                      const double tA3 = -(c13 * p.rx + c23 * p.ry + c33 * p.rz) * Wi;
                      const double factor = a.f.xm[p.j] / a.f.kx[p.j] * (divv_i - a.f.divv[p.j]);
                      g1 += factor * tA1, g2 += factor * tA2, g3 += factor * tA3;
                  });
    const double graddivv = sqrt(g1 * g1 + g2 * g2 + g3 * g3);
    double       alpha_i = a.f.alpha[i], alphaloc = 0.0;
    if (divv_i < 0.0)
    {
        const double a_const = hi * hi * graddivv;
        alphaloc             = a.alphamax * a_const / (a_const + hi * fabs(divv_i) + 0.05 * ci);
    }
    if (alphaloc >= alpha_i) { alpha_i = alphaloc; }
    else
    {
        const double decay    = hi / (a.decay_constant * vijsignal_i);
        const double alphadot = (alphaloc >= a.alphamin) ? (alphaloc - alpha_i) / decay : (a.alphamin - alpha_i) / decay;
        alpha_i += alphadot * a.minDt;
    }
    a.f.alpha[i] = alpha_i;
}

//! symmetric-upper mat-vec and dot with R (kernels.hpp:87-95)
__device__ __forceinline__ double symvDot(const double* g, double rx, double ry, double rz)
{
    const double r0 = g[0] * rx + g[1] * ry + g[2] * rz, r1 = g[3] * ry + g[4] * rz, r2 = g[5] * rz;
    return rx * r0 + ry * r1 + rz * r2;
}

// momentum_energy_kern.hpp:65-222 (avRvCorrection :43-63), tsKCourant kernels.hpp:10-16
__global__ void momentumEnergyKernel(const __grid_constant__ Args a)
{
    SPHX_F64_TARGET
    const double hi = a.f.h[i], hiInv = 1.0 / hi, hiInv3 = hiInv * hiInv * hiInv;
    const double vxi = a.f.vx[i], vyi = a.f.vy[i], vzi = a.f.vz[i], ci = a.f.c[i], alpha_i = a.f.alpha[i];
    const double mi = a.f.m[i], xmassi = a.f.xm[i], rhoi = a.f.kx[i] * mi / xmassi, prhoi = a.f.prho[i];
    const double c11i = a.f.c11[i], c12i = a.f.c12[i], c13i = a.f.c13[i], c22i = a.f.c22[i], c23i = a.f.c23[i],
                 c33i = a.f.c33[i];
    double gradVi[6] = {0, 0, 0, 0, 0, 0};
    double eta_crit  = 0.0;
    if (a.avClean)
    {
        gradVi[0] = a.f.dV11[i], gradVi[1] = a.f.dV12[i], gradVi[2] = a.f.dV13[i];
        gradVi[3] = a.f.dV22[i], gradVi[4] = a.f.dV23[i], gradVi[5] = a.f.dV33[i];
        const unsigned n = min(a.f.nc[i] - 1u, a.ngmax);
        eta_crit         = cbrt(32.0 * M_PI / 3.0 / double(n + 1));
    }
    double momx = 0, momy = 0, momz = 0, energy = 0, a_visc_energy = 0, maxvsignal = 0;
    forNeighbours(
        a, i,
        [&](const Pair& p)
        {
            const unsigned j  = p.j;
            const double   rx = p.rx, ry = p.ry, rz = p.rz, dist = p.dist;
            const double   vx_ij = vxi - a.f.vx[j], vy_ij = vyi - a.f.vy[j], vz_ij = vzi - a.f.vz[j];
            const double   hjInv = 1.0 / a.f.h[j], hjInv3 = hjInv * hjInv * hjInv;
            const double   v1 = dist * hiInv, v2 = dist * hjInv;
            const double   Wi = hiInv3 * lookup(a.wh, v1), Wj = hjInv3 * lookup(a.wh, v2);
            const double   tA1i = -(c11i * rx + c12i * ry + c13i * rz) * Wi, tA2i = -(c12i * rx + c22i * ry + c23i * rz) * Wi,
                         tA3i = -(c13i * rx + c23i * ry + c33i * rz) * Wi;
            const double c11j = a.f.c11[j], c12j = a.f.c12[j], c13j = a.f.c13[j], c22j = a.f.c22[j], c23j = a.f.c23[j],
                         c33j = a.f.c33[j];
            const double tA1j = -(c11j * rx + c12j * ry + c13j * rz) * Wj, tA2j = -(c12j * rx + c22j * ry + c23j * rz) * Wj,
                         tA3j = -(c13j * rx + c23j * ry + c33j * rz) * Wj;
            const double mj = a.f.m[j], cj = a.f.c[j], xmassj = a.f.xm[j], rhoj = a.f.kx[j] * mj / xmassj;
            double       rv = rx * vx_ij + ry * vy_ij + rz * vz_ij;
            if (a.avClean)
            {
                const double gj[6] = {a.f.dV11[j], a.f.dV12[j], a.f.dV13[j], a.f.dV22[j], a.f.dV23[j], a.f.dV33[j]};
                const double eta_ab = fmin(v1, v2);
                const double dmy1 = symvDot(gradVi, rx, ry, rz), dmy2 = symvDot(gj, rx, ry, rz);
                double       dmy3 = 1.0;
                if (eta_ab < eta_crit)
                {
                    const double etaDiff = 5.0 * (eta_ab - eta_crit);
                    dmy3                 = exp(-etaDiff * etaDiff);
                }
                const double A_ab = (dmy2 != 0.0) ? dmy1 / dmy2 : 0.0, A_abp1 = 1.0 + A_ab;
                const double phi_ab = 0.5 * dmy3 * fmax(0.0, fmin(1.0, 4.0 * A_ab / (A_abp1 * A_abp1)));
                rv += -phi_ab * (dmy1 + dmy2);
            }
            const double wij = rv / dist;
            // artificial_viscosity (kernels.hpp:70-84)
            const double vij_signal = (alpha_i + a.f.alpha[j]) / 4.0 * (ci + cj) - 2.0 * wij;
            const double viscosity  = (wij < 0.0) ? -vij_signal * wij : 0.0;
            maxvsignal              = fmax(maxvsignal, 0.5 * (ci + cj) - 2.0 * wij);

            const double a_visc = mj / rhoi * viscosity, b_visc = mj / rhoj * viscosity;
            const double a_visc_x = 0.5 * (a_visc * tA1i + b_visc * tA1j), a_visc_y = 0.5 * (a_visc * tA2i + b_visc * tA2j),
                         a_visc_z = 0.5 * (a_visc * tA3i + b_visc * tA3j);
            a_visc_energy += a_visc_x * vx_ij + a_visc_y * vy_ij + a_visc_z * vz_ij;

            // Atwood-number ramp between crossed and uncrossed volume elements (momentum_energy_kern.hpp:143-164)
            const double Atwood = fabs(rhoi - rhoj) / (rhoi + rhoj);
            double       a_mom, b_mom;
            if (Atwood < a.Atmin) { a_mom = xmassi * xmassi, b_mom = xmassj * xmassj; }
            else if (Atwood > a.Atmax) { a_mom = xmassi * xmassj, b_mom = a_mom; }
            else
            {
                const double sigma_ij = a.ramp * (Atwood - a.Atmin);
                a_mom                 = pow(xmassi, 2.0 - sigma_ij) * pow(xmassj, sigma_ij);
                b_mom                 = pow(xmassj, 2.0 - sigma_ij) * pow(xmassi, sigma_ij);
            }
            energy += mj * a_mom * (vx_ij * tA1i + vy_ij * tA2i + vz_ij * tA3i);
            const double momentum_i = mj * prhoi * a_mom, momentum_j = mj * a.f.prho[j] * b_mom;
            momx += momentum_i * tA1i + momentum_j * tA1j + a_visc_x;
            momy += momentum_i * tA2i + momentum_j * tA2j + a_visc_y;
            momz += momentum_i * tA3i + momentum_j * tA3j + a_visc_z;
        });
    a_visc_energy = fmax(0.0, a_visc_energy);
    a.f.du[i]     = a.K * (prhoi * energy + 0.5 * a_visc_energy);
    a.f.ax[i] = -a.K * momx, a.f.ay[i] = -a.K * momy, a.f.az[i] = -a.K * momz;
    const double v = maxvsignal > 0.0 ? maxvsignal : ci;
    atomicMinDouble(&a.scal->minDtCourant, a.Kcour * hi / v);
}

__global__ void resetScalarsKernel(StepScalarsF64* s)
{
    s->minDtCourant = INFINITY, s->maxDivv = -INFINITY;
}

} // namespace f64
} // namespace sphx

namespace
{

thread_local std::string g_f64Error;

int f64Fail(int code, const std::string& m)
{
    sphx::setLastError(m);
    return code;
}

struct F64Layout
{
    size_t scalOff, scal64Off, listOff, total;
    F64Layout(size_t numAssigned, unsigned ngmax)
    {
        scalOff   = 0;
        scal64Off = sphx::kScalarsBytes;
        listOff   = 2 * sphx::kScalarsBytes;
        size_t groups = (numAssigned + sphx::kGroupSize - 1) / sphx::kGroupSize;
        total         = sphx::alignUp(listOff + groups * size_t(ngmax) * sphx::kGroupSize * sizeof(unsigned), 256);
    }
};

int makeArgs(const SphxStepArgsF64* a, sphx::f64::Args& k, F64Layout& lay)
{
    if (!a) return f64Fail(SPHX_ERR_INVALID, "null args");
    if (int rc = sphx_device_check()) return rc;
    if (a->last < a->first || a->last > a->numLocal || a->numLocal >= (size_t(1) << 32))
        return f64Fail(SPHX_ERR_INVALID, "bad [first,last) range");
    if (a->p.ngmax == 0 || a->p.ngmax > 4096) return f64Fail(SPHX_ERR_INVALID, "ngmax must be in [1, 4096]");
    lay = F64Layout(a->last - a->first, a->p.ngmax);
    if (!a->workspace || a->workspaceBytes < lay.total)
        return f64Fail(SPHX_ERR_WORKSPACE, "workspace too small: need " + std::to_string(lay.total) + " bytes");
    char* base = static_cast<char*>(a->workspace);
    k.f     = a->f;
    k.first = unsigned(a->first), k.last = unsigned(a->last), k.ngmax = a->p.ngmax;
    k.box   = sphx::makeDevBox(a->box);
    k.list  = reinterpret_cast<const unsigned*>(base + lay.listOff);
    k.wh = a->wh, k.whd = a->whd;
    k.scal = reinterpret_cast<sphx::StepScalarsF64*>(base + lay.scal64Off);
    k.K = a->p.K, k.minDt = a->p.minDt, k.Kcour = a->p.Kcour, k.gamma = a->p.gamma, k.muiConst = a->p.muiConst;
    k.alphamin = a->p.alphamin, k.alphamax = a->p.alphamax, k.decay_constant = a->p.decay_constant;
    k.Atmin = a->p.Atmin, k.Atmax = a->p.Atmax, k.ramp = a->p.ramp, k.avClean = a->p.avClean;
    return SPHX_OK;
}

#define F64_REQUIRE(ptr)                                                                                               \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(ptr)) return f64Fail(SPHX_ERR_INVALID, "required field is NULL: " #ptr);                                 \
    } while (0)

template<class Kernel, class... Extra>
int launch(Kernel kern, const sphx::f64::Args& k, cudaStream_t s, Extra... extra)
{
    const unsigned n = k.last - k.first;
    if (n == 0) return SPHX_OK;
    kern<<<(n + 127) / 128, 128, 0, s>>>(k, extra...);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return f64Fail(SPHX_ERR_CUDA, cudaGetErrorString(e));
    return SPHX_OK;
}

int readResult(const SphxStepArgsF64* a, const F64Layout& lay, SphxStepResult* r, bool withFlags)
{
    if (!r) return SPHX_OK;
    auto                 s    = static_cast<cudaStream_t>(a->stream);
    char*                base = static_cast<char*>(a->workspace);
    sphx::StepScalars    h;
    sphx::StepScalarsF64 h64;
    if (cudaMemcpyAsync(&h, base + lay.scalOff, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(&h64, base + lay.scal64Off, sizeof(h64), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return f64Fail(SPHX_ERR_CUDA, "reading the step scalars failed");
    r->minDtCourant   = h64.minDtCourant;
    r->minDtRho       = h64.maxDivv == -INFINITY ? double(INFINITY) : a->p.Krho / std::fabs(h64.maxDivv);
    r->totalNeighbors = h.totalNeighbors;
    r->maxNc          = h.maxNc;
    r->numHIterated   = h.numHIterated;
    if (withFlags)
    {
        if (h.errFlags & sphx::kErrTraversal) return f64Fail(SPHX_ERR_TRAVERSAL, "GPU traversal stack exhausted in neighbor search");
        if (h.errFlags & sphx::kErrHConv) return f64Fail(SPHX_ERR_H_CONVERGENCE, "coupled nc/h-updated failed to converge");
        if (h.errFlags & sphx::kErrNgmax) return f64Fail(SPHX_ERR_NGMAX_OVERFLOW, "neighbour count exceeds ngmax after h-iteration");
    }
    return SPHX_OK;
}

} // namespace

extern "C"
{

size_t sphx_workspace_bytes_f64(size_t numAssigned, unsigned ngmax) { return F64Layout(numAssigned, ngmax).total; }

int sphx_find_neighbors_sph_f64(const SphxStepArgsF64* a, SphxStepResult* r)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    if (a->p.ng0 > a->p.ngmax) return f64Fail(SPHX_ERR_INVALID, "ng0 should be smaller than ngmax");
    if (a->tree.numLeafNodes <= 0 || !a->tree.childOffsets || !a->tree.internalToLeaf || !a->tree.layout ||
        !a->tree.centers || !a->tree.sizes)
        return f64Fail(SPHX_ERR_INVALID, "incomplete tree view");
    F64_REQUIRE(a->f.x); F64_REQUIRE(a->f.y); F64_REQUIRE(a->f.z); F64_REQUIRE(a->f.h); F64_REQUIRE(a->f.nc);
    auto  s    = static_cast<cudaStream_t>(a->stream);
    char* base = static_cast<char*>(a->workspace);
    auto* scal = reinterpret_cast<sphx::StepScalars*>(base + lay.scalOff);
    sphx::launchResetScalars(scal, s);
    sphx::f64::resetScalarsKernel<<<1, 1, 0, s>>>(k.scal);
    sphx::launchFindNeighborsSphF64(a->f.x, a->f.y, a->f.z, a->f.h, k.first, k.last, a->box, a->tree, a->p.ng0, a->p.ngmax,
                                    reinterpret_cast<unsigned*>(base + lay.listOff), a->f.nc, scal, s);
    if (cudaGetLastError() != cudaSuccess) return f64Fail(SPHX_ERR_CUDA, "search launch failed");
    return readResult(a, lay, r, true);
}

int sphx_xmass_f64(const SphxStepArgsF64* a)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    F64_REQUIRE(a->f.m); F64_REQUIRE(a->f.xm); F64_REQUIRE(a->f.nc); F64_REQUIRE(a->wh);
    return launch(sphx::f64::xmassKernel, k, static_cast<cudaStream_t>(a->stream));
}

int sphx_ve_def_gradh_f64(const SphxStepArgsF64* a)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    F64_REQUIRE(a->f.m); F64_REQUIRE(a->f.xm); F64_REQUIRE(a->f.kx); F64_REQUIRE(a->f.gradh); F64_REQUIRE(a->wh);
    F64_REQUIRE(a->whd);
    return launch(sphx::f64::gradhKernel, k, static_cast<cudaStream_t>(a->stream));
}

int sphx_eos_f64(const SphxStepArgsF64* a)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    F64_REQUIRE(a->f.kx); F64_REQUIRE(a->f.xm); F64_REQUIRE(a->f.m); F64_REQUIRE(a->f.gradh); F64_REQUIRE(a->f.prho);
    F64_REQUIRE(a->f.c);
    if (a->p.eosChoice == 0 && !a->f.temp && !a->f.u) return f64Fail(SPHX_ERR_INVALID, "ideal gas EOS needs temp or u");
    if (a->p.eosChoice < 0 || a->p.eosChoice > 2) return f64Fail(SPHX_ERR_INVALID, "unknown eosChoice");
    return launch(sphx::f64::eosKernel, k, static_cast<cudaStream_t>(a->stream), a->p.eosChoice, a->p.soundSpeedConst,
                  a->p.polytropic_const, a->p.polytropic_index);
}

int sphx_iad_divv_curlv_f64(const SphxStepArgsF64* a, SphxStepResult* r)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    F64_REQUIRE(a->f.vx); F64_REQUIRE(a->f.vy); F64_REQUIRE(a->f.vz); F64_REQUIRE(a->f.xm); F64_REQUIRE(a->f.kx);
    F64_REQUIRE(a->f.c11); F64_REQUIRE(a->f.c12); F64_REQUIRE(a->f.c13); F64_REQUIRE(a->f.c22); F64_REQUIRE(a->f.c23);
    F64_REQUIRE(a->f.c33); F64_REQUIRE(a->f.divv);
    if (int rc = launch(sphx::f64::iadDivvCurlvKernel, k, static_cast<cudaStream_t>(a->stream))) return rc;
    return readResult(a, lay, r, false);
}

int sphx_av_switches_f64(const SphxStepArgsF64* a)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    F64_REQUIRE(a->f.c); F64_REQUIRE(a->f.divv); F64_REQUIRE(a->f.alpha); F64_REQUIRE(a->f.c11); F64_REQUIRE(a->f.kx);
    return launch(sphx::f64::avSwitchesKernel, k, static_cast<cudaStream_t>(a->stream));
}

int sphx_momentum_energy_f64(const SphxStepArgsF64* a, SphxStepResult* r)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    F64_REQUIRE(a->f.prho); F64_REQUIRE(a->f.c); F64_REQUIRE(a->f.alpha); F64_REQUIRE(a->f.ax); F64_REQUIRE(a->f.ay);
    F64_REQUIRE(a->f.az); F64_REQUIRE(a->f.du); F64_REQUIRE(a->f.c11);
    if (a->p.avClean) { F64_REQUIRE(a->f.dV11); }
    if (int rc = launch(sphx::f64::momentumEnergyKernel, k, static_cast<cudaStream_t>(a->stream))) return rc;
    return readResult(a, lay, r, true);
}

int sphx_hydro_step_f64(const SphxStepArgsF64* a, SphxStepResult* r)
{
    // ve_hydro.hpp:147-190 on one rank
    if (int rc = sphx_find_neighbors_sph_f64(a, nullptr)) return rc;
    if (int rc = sphx_xmass_f64(a)) return rc;
    if (int rc = sphx_ve_def_gradh_f64(a)) return rc;
    if (int rc = sphx_eos_f64(a)) return rc;
    if (int rc = sphx_iad_divv_curlv_f64(a, nullptr)) return rc;
    if (int rc = sphx_av_switches_f64(a)) return rc;
    SphxStepResult local;
    if (int rc = sphx_momentum_energy_f64(a, &local)) return rc;
    if (r) *r = local;
    return SPHX_OK;
}

int sphx_export_neighbors_f64(const SphxStepArgsF64* a, unsigned* neighbors_dev)
{
    sphx::f64::Args k;
    F64Layout       lay(0, 1);
    if (int rc = makeArgs(a, k, lay)) return rc;
    F64_REQUIRE(neighbors_dev); F64_REQUIRE(a->f.nc);
    sphx::launchExportNeighbors(unsigned(a->last - a->first), a->p.ngmax, k.list, a->f.nc + a->first, true, neighbors_dev,
                                static_cast<cudaStream_t>(a->stream));
    return SPHX_OK;
}

} // extern "C"

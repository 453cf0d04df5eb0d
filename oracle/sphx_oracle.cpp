/* TEST INFRASTRUCTURE ONLY — see sphx_oracle.h.
 *
 * CPU restatement of the reference's SPH-VE hydro step. Each function cites the reference file:line it follows
 * (paths relative to the reference root). Written independently as flat loops over plain arrays; arithmetic
 * types and evaluation order follow the reference expression by expression, because the neighbour sets must be
 * bit-exact (build with -ffp-contract=off).
 */
#include "sphx_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <numeric>
#include <vector>

namespace
{

constexpr int kTableSize = 20000; // sph/table_lookup.hpp:10

struct Box
{
    double xmin, xmax, ymin, ymax, zmin, zmax;
    double lx, ly, lz, ilx, ily, ilz; // sfc/box.hpp:94-120: lengths and T(1)/length
    bool   pbcX, pbcY, pbcZ;

    explicit Box(const OrcBox* b)
    {
        xmin = b->lim[0], xmax = b->lim[1], ymin = b->lim[2], ymax = b->lim[3], zmin = b->lim[4], zmax = b->lim[5];
        lx = xmax - xmin, ly = ymax - ymin, lz = zmax - zmin;
        ilx = 1.0 / (xmax - xmin), ily = 1.0 / (ymax - ymin), ilz = 1.0 / (zmax - zmin);
        pbcX = b->boundary[0] == 1, pbcY = b->boundary[1] == 1, pbcZ = b->boundary[2] == 1;
    }
};

/* ---------------- kernel tables: sph_kernel_tables.hpp ---------------- */

// sph/kernels.hpp:34-42
double wharmonic(double v)
{
    if (v == 0.0) { return 1.0; }
    const double Pv = M_PI_2 * v;
    return std::sin(Pv) / Pv;
}

// sph/kernels.hpp:48-58
double wharmonicDerivative(double v)
{
    if (v == 0.0) return 0.0;
    const double Pv    = M_PI_2 * v;
    const double sincv = std::sin(Pv) / (Pv);
    return sincv * M_PI_2 * ((std::cos(Pv) / std::sin(Pv)) - 1.0 / Pv);
}

// sph_kernel_tables.hpp:22-56 (util::simpson, with its sorted partial sums)
template<class F>
double simpson(double a, double b, uint64_t n, F&& func)
{
    uint64_t numOdd  = n / 2;
    uint64_t numEven = (numOdd >= 1) ? numOdd - 1 : 0;
    double   h       = (b - a) / double(n);

    std::vector<double> odd(numOdd), even(numEven);
    for (uint64_t i = 0; i < numOdd; ++i)
        odd[i] = func(a + double(2 * (i + 1) - 1) * h);
    for (uint64_t i = 0; i < numEven; ++i)
        even[i] = func(a + double(2 * (i + 1)) * h);
    std::sort(odd.begin(), odd.end());
    std::sort(even.begin(), even.end());
    return h / 3.0 *
           (func(a) + func(b) + 4.0 * std::accumulate(odd.begin(), odd.end(), 0.0) +
            2.0 * std::accumulate(even.begin(), even.end(), 0.0));
}

// sph_kernel_tables.hpp:77-101,144-172 with kernelChoice = sinc_n; the functor is double(double)
template<class T>
void makeTables(double sincIndex, T* wh, T* whd, double* K)
{
    auto kern  = [sincIndex](double x) { return std::pow(wharmonic(x), sincIndex); };
    auto kernD = [sincIndex](double x)
    { return sincIndex * std::pow(wharmonic(x), sincIndex - 1) * wharmonicDerivative(x); };

    *K = 1.0 / simpson(0, 2.0, 2000, [&](double x) { return 4.0 * M_PI * x * x * kern(x); });

    const T dx = (2.0 - 0.0) / (kTableSize - 1);
    for (size_t i = 0; i < size_t(kTableSize); ++i)
    {
        T v    = 0.0 + i * dx; // abscissa stepped and rounded in T
        wh[i]  = kern(v);
        whd[i] = kernD(v);
    }
}

// sph/table_lookup.hpp:13-26
template<class T>
inline T lookup(const T* table, T v)
{
    constexpr int numIntervals = kTableSize - 1;
    constexpr T   support      = 2.0;
    constexpr T   dx           = support / numIntervals;
    constexpr T   invDx        = T(1) / dx;

    int idx = v * invDx;
    if (idx >= numIntervals) { return T(0); }
    T derivative = (table[idx + 1] - table[idx]) * invDx;
    return table[idx] + derivative * (v - T(idx) * dx);
}

/* ---------------- geometry: sfc/box.hpp, traversal/boxoverlap.hpp, findneighbors.hpp ---------------- */

// findneighbors.hpp:33-60
template<bool Pbc>
inline double distanceSq(double x1, double y1, double z1, double x2, double y2, double z2, const Box& box)
{
    double dx = x1 - x2;
    double dy = y1 - y2;
    double dz = z1 - z2;
    if constexpr (Pbc)
    {
        dx -= box.pbcX * box.lx * std::rint(dx * box.ilx);
        dy -= box.pbcY * box.ly * std::rint(dy * box.ily);
        dz -= box.pbcZ * box.lz * std::rint(dz * box.ilz);
    }
    return dx * dx + dy * dy + dz * dz;
}

// sfc/box.hpp:282-304 (legacy PBC used by all J-loops), T r = 2h
template<class T>
inline void applyPBC(const Box& box, T r, T& xx, T& yy, T& zz)
{
    if (box.pbcX && xx > r) xx -= box.lx;
    else if (box.pbcX && xx < -r)
        xx += box.lx;
    if (box.pbcY && yy > r) yy -= box.ly;
    else if (box.pbcY && yy < -r)
        yy += box.ly;
    if (box.pbcZ && zz > r) zz -= box.lz;
    else if (box.pbcZ && zz < -r)
        zz += box.lz;
}

// sfc/box.hpp:306-316
template<class T>
inline T distancePBC(const Box& box, T hi, double x1, double y1, double z1, double x2, double y2, double z2)
{
    T xx = x1 - x2;
    T yy = y1 - y2;
    T zz = z1 - z2;
    applyPBC(box, T(2) * hi, xx, yy, zz);
    return std::sqrt(xx * xx + yy * yy + zz * zz);
}

// util/array.hpp:236-240: norm2 is a RIGHT fold a0*a0 + (a1*a1 + a2*a2)
inline double norm2r(double a, double b, double c) { return a * a + (b * b + c * c); }

// traversal/boxoverlap.hpp:196-216: squared min distance point <-> box, open and periodic
template<bool Pbc>
inline double minDistSq(const double* X, const double* ctr, const double* sz, const Box& box)
{
    double d[3];
    for (int k = 0; k < 3; ++k)
        d[k] = ctr[k] - X[k];
    if constexpr (Pbc)
    {
        d[0] -= box.pbcX * box.lx * std::rint(d[0] * box.ilx);
        d[1] -= box.pbcY * box.ly * std::rint(d[1] * box.ily);
        d[2] -= box.pbcZ * box.lz * std::rint(d[2] * box.ilz);
    }
    for (int k = 0; k < 3; ++k)
    {
        double v = std::abs(d[k]) - sz[k];
        v += std::abs(v);
        v *= 0.5;
        d[k] = v;
    }
    return norm2r(d[0], d[1], d[2]);
}

// cstone::findNeighbors for one particle: findneighbors.hpp:77-147 + singleTraversal traversal/traversal.hpp:51-93
template<class Th>
unsigned findNeighborsOne(unsigned i, const double* x, const double* y, const double* z, const Th* h,
                          const OrcTree& tree, const Box& box, unsigned ngmax, unsigned* neighbors)
{
    double xi = x[i], yi = y[i], zi = z[i];
    Th     hi = h[i];

    Th     radiusSq     = Th(4.0) * hi * hi;
    auto   cellRadiusSq = radiusSq * tree.searchExtFactor * tree.searchExtFactor;
    double P[3]         = {xi, yi, zi};
    unsigned numNeighbors = 0;

    bool   anyPbc = box.pbcX || box.pbcY || box.pbcZ;
    double ext    = 2.0 * hi; // Tc(2) * hi
    // boxoverlap.hpp:183-192 insideBox(particle, {2h,2h,2h}, box)
    bool inside = (xi - ext) >= box.xmin && (yi - ext) >= box.ymin && (zi - ext) >= box.zmin &&
                  (xi + ext) <= box.xmax && (yi + ext) <= box.ymax && (zi + ext) <= box.zmax;
    bool usePbc = anyPbc && !inside;

    auto overlaps = [&](int node)
    {
        return usePbc ? minDistSq<true>(P, tree.centers + 3 * node, tree.sizes + 3 * node, box) < cellRadiusSq
                      : minDistSq<false>(P, tree.centers + 3 * node, tree.sizes + 3 * node, box) < cellRadiusSq;
    };
    auto searchLeaf = [&](int node)
    {
        int      leaf = tree.internalToLeaf[node];
        unsigned a = tree.layout[leaf], b = tree.layout[leaf + 1];
        for (unsigned j = a; j < b; ++j)
        {
            if (j == i) continue;
            double d2 = usePbc ? distanceSq<true>(x[j], y[j], z[j], xi, yi, zi, box)
                               : distanceSq<false>(x[j], y[j], z[j], xi, yi, zi, box);
            if (d2 < radiusSq)
            {
                if (numNeighbors < ngmax) neighbors[numNeighbors] = j;
                numNeighbors++;
            }
        }
    };

    const int* childOffsets = tree.childOffsets;
    if (!overlaps(0)) return 0;
    if (childOffsets[0] == 0)
    {
        searchLeaf(0);
        return numNeighbors;
    }

    int stack[128];
    stack[0]     = 0;
    int stackPos = 1;
    int node     = 0;
    do
    {
        for (int octant = 0; octant < 8; ++octant)
        {
            int child = childOffsets[node] + octant;
            if (overlaps(child))
            {
                if (childOffsets[child] == 0) searchLeaf(child);
                else
                    stack[stackPos++] = child;
            }
        }
        node = stack[--stackPos];
    } while (node != 0);

    return numNeighbors;
}

// sph/kernels.hpp:26-32
template<class T>
inline T updateH(unsigned ng0, unsigned nc, T h)
{
    constexpr T c0  = 1023.0;
    constexpr T exp = 1.0 / 10.0;
    return h * T(0.5) * std::pow(T(1) + c0 * ng0 / T(nc), exp);
}

/* ---------------- the J-loops (hydro_ve/*_kern.hpp). Tc = double; T = hydro type ---------------- */

// xmass_kern.hpp:51-79
template<class T>
T xmassJLoop(unsigned i, double K, const Box& box, const unsigned* nb, unsigned nc, const double* x, const double* y,
             const double* z, const T* h, const T* m, const T* wh)
{
    double xi = x[i], yi = y[i], zi = z[i];
    T      hi = h[i], mi = m[i];

    T hInv  = 1.0 / hi;
    T h3Inv = hInv * hInv * hInv;

    T rho0i = mi;
    for (unsigned pj = 0; pj < nc; ++pj)
    {
        unsigned j    = nb[pj];
        T        dist = distancePBC(box, hi, xi, yi, zi, x[j], y[j], z[j]);
        T        vloc = dist * hInv;
        T        w    = lookup(wh, vloc);
        rho0i += w * m[j];
    }
    T xmassi = mi / (rho0i * K * h3Inv); // veDefinition: evaluated in double, rounded once
    return xmassi;
}

// ve_def_gradh_kern.hpp:44-90
template<class T>
void veDefGradhJLoop(unsigned i, double K, const Box& box, const unsigned* nb, unsigned nc, const double* x,
                     const double* y, const double* z, const T* h, const T* m, const T* wh, const T* whd, const T* xm,
                     T* kxOut, T* gradhOut)
{
    double xi = x[i], yi = y[i], zi = z[i];
    T      hi = h[i], mi = m[i], xmassi = xm[i];

    T hInv  = T(1) / hi;
    T h3Inv = hInv * hInv * hInv;

    T kxi      = xmassi;
    T whomegai = -T(3) * xmassi;
    T wrho0i   = -T(3) * mi;

    for (unsigned pj = 0; pj < nc; ++pj)
    {
        unsigned j      = nb[pj];
        T        dist   = distancePBC(box, hi, xi, yi, zi, x[j], y[j], z[j]);
        T        vloc   = dist * hInv;
        T        w      = lookup(wh, vloc);
        T        dw     = lookup(whd, vloc);
        T        dterh  = -(T(3) * w + vloc * dw);
        T        xmassj = xm[j];

        kxi += w * xmassj;
        whomegai += dterh * xmassj;
        wrho0i += dterh * m[j];
    }

    kxi *= K * h3Inv;
    whomegai *= K * h3Inv * hInv;
    wrho0i *= K * h3Inv * hInv;

    whomegai = whomegai * mi / xmassi + (kxi - K * xmassi * h3Inv) * wrho0i;
    T rhoi   = kxi * mi / xmassi;
    T dhdrho = -hi / (rhoi * T(3));

    *kxOut    = kxi;
    *gradhOut = T(1) - dhdrho * whomegai;
}

// iad_kern.hpp:44-109
template<class T>
void iadJLoop(unsigned i, double K, const Box& box, const unsigned* nb, unsigned nc, const double* x, const double* y,
              const double* z, const T* h, const T* wh, const T* xm, const T* kx, T* c /* 6 */)
{
    T      tau11 = 0.0, tau12 = 0.0, tau13 = 0.0, tau22 = 0.0, tau23 = 0.0, tau33 = 0.0;
    double xi = x[i], yi = y[i], zi = z[i];
    T      hi    = h[i];
    T      hiInv = T(1) / hi;

    for (unsigned pj = 0; pj < nc; ++pj)
    {
        unsigned j  = nb[pj];
        T        rx = (xi - x[j]);
        T        ry = (yi - y[j]);
        T        rz = (zi - z[j]);
        applyPBC(box, T(2) * hi, rx, ry, rz);

        T dist   = std::sqrt(rx * rx + ry * ry + rz * rz);
        T vloc   = dist * hiInv;
        T w      = lookup(wh, vloc);
        T volj_w = xm[j] / kx[j] * w;

        tau11 += rx * rx * volj_w;
        tau12 += rx * ry * volj_w;
        tau13 += rx * rz * volj_w;
        tau22 += ry * ry * volj_w;
        tau23 += ry * rz * volj_w;
        tau33 += rz * rz * volj_w;
    }

    auto getExp    = [](T val) { return (val == T(0) ? 0 : std::ilogb(val)); };
    int  tauExpSum = getExp(tau11) + getExp(tau12) + getExp(tau13) + getExp(tau22) + getExp(tau23) + getExp(tau33);
    T    normalization = std::ldexp(T(1), -tauExpSum / 6);

    tau11 *= normalization;
    tau12 *= normalization;
    tau13 *= normalization;
    tau22 *= normalization;
    tau23 *= normalization;
    tau33 *= normalization;

    T det = tau11 * tau22 * tau33 + T(2) * tau12 * tau23 * tau13 - tau11 * tau23 * tau23 - tau22 * tau13 * tau13 -
            tau33 * tau12 * tau12;

    T factor = normalization * (hi * hi * hi) / (det * K);

    c[0] = (tau22 * tau33 - tau23 * tau23) * factor;
    c[1] = (tau13 * tau23 - tau33 * tau12) * factor;
    c[2] = (tau12 * tau23 - tau22 * tau13) * factor;
    c[3] = (tau11 * tau33 - tau13 * tau13) * factor;
    c[4] = (tau13 * tau12 - tau11 * tau23) * factor;
    c[5] = (tau11 * tau22 - tau12 * tau12) * factor;
}

// divv_curlv_kern.hpp:44-123; out = {divv, curlv, dV11, dV12, dV13, dV22, dV23, dV33}
template<class T>
void divvCurlvJLoop(unsigned i, double K, const Box& box, const unsigned* nb, unsigned nc, const double* x,
                    const double* y, const double* z, const T* vx, const T* vy, const T* vz, const T* h, const T* c11,
                    const T* c12, const T* c13, const T* c22, const T* c23, const T* c33, const T* wh, const T* kx,
                    const T* xm, T* out)
{
    double xi = x[i], yi = y[i], zi = z[i];
    T      vxi = vx[i], vyi = vy[i], vzi = vz[i];
    T      hi = h[i], kxi = kx[i];

    T hiInv  = T(1) / hi;
    T hiInv3 = hiInv * hiInv * hiInv;

    T dVx[3] = {0, 0, 0}, dVy[3] = {0, 0, 0}, dVz[3] = {0, 0, 0};
    T c11i = c11[i], c12i = c12[i], c13i = c13[i], c22i = c22[i], c23i = c23[i], c33i = c33[i];

    for (unsigned pj = 0; pj < nc; ++pj)
    {
        unsigned j  = nb[pj];
        T        rx = xi - x[j];
        T        ry = yi - y[j];
        T        rz = zi - z[j];
        applyPBC(box, T(2) * hi, rx, ry, rz);

        T r2   = rx * rx + ry * ry + rz * rz;
        T dist = std::sqrt(r2);

        T vx_ji = vx[j] - vxi;
        T vy_ji = vy[j] - vyi;
        T vz_ji = vz[j] - vzi;

        T v1 = dist * hiInv;
        T Wi = lookup(wh, v1);

        T termA[3];
        termA[0] = -(c11i * rx + c12i * ry + c13i * rz) * Wi;
        termA[1] = -(c12i * rx + c22i * ry + c23i * rz) * Wi;
        termA[2] = -(c13i * rx + c23i * ry + c33i * rz) * Wi;

        T xmassj = xm[j];
        T fx = vx_ji * xmassj, fy = vy_ji * xmassj, fz = vz_ji * xmassj;
        for (int k = 0; k < 3; ++k)
        {
            dVx[k] += fx * termA[k];
            dVy[k] += fy * termA[k];
            dVz[k] += fz * termA[k];
        }
    }

    T norm_kxi = K * hiInv3 / kxi;
    out[0]     = norm_kxi * (dVx[0] + dVy[1] + dVz[2]);

    T cu[3] = {dVz[1] - dVy[2], dVx[2] - dVz[0], dVy[0] - dVx[1]};
    out[1]  = norm_kxi * std::sqrt(cu[0] * cu[0] + (cu[1] * cu[1] + cu[2] * cu[2])); // norm2: right fold

    out[2] = norm_kxi * dVx[0];
    out[3] = norm_kxi * (dVx[1] + dVy[0]);
    out[4] = norm_kxi * (dVx[2] + dVz[0]);
    out[5] = norm_kxi * dVy[1];
    out[6] = norm_kxi * (dVy[2] + dVz[1]);
    out[7] = norm_kxi * dVz[2];
}

// av_switches_kern.hpp:44-137
template<class T>
T avSwitchesJLoop(unsigned i, double K, const Box& box, const unsigned* nb, unsigned nc, const double* x,
                  const double* y, const double* z, const T* vx, const T* vy, const T* vz, const T* h, const T* c,
                  const T* c11, const T* c12, const T* c13, const T* c22, const T* c23, const T* c33, const T* wh,
                  const T* kx, const T* xm, const T* divv, double dt, T alphamin, T alphamax, T decay_constant,
                  T alpha_i)
{
    double xi = x[i], yi = y[i], zi = z[i];
    T      vxi = vx[i], vyi = vy[i], vzi = vz[i];
    T      hi = h[i], ci = c[i];
    T      c11i = c11[i], c12i = c12[i], c13i = c13[i], c22i = c22[i], c23i = c23[i], c33i = c33[i];

    T vijsignal_i = T(1.e-40) * ci;

    T hiInv  = T(1) / hi;
    T hiInv3 = hiInv * hiInv * hiInv;

    T divv_i = divv[i];

    T graddivv_x = 0.0, graddivv_y = 0.0, graddivv_z = 0.0;

    for (unsigned pj = 0; pj < nc; ++pj)
    {
        unsigned j  = nb[pj];
        T        rx = xi - x[j];
        T        ry = yi - y[j];
        T        rz = zi - z[j];
        applyPBC(box, T(2) * hi, rx, ry, rz);

        T r2   = rx * rx + ry * ry + rz * rz;
        T dist = std::sqrt(r2);

        T vx_ij = vxi - vx[j];
        T vy_ij = vyi - vy[j];
        T vz_ij = vzi - vz[j];

        T rv           = rx * vx_ij + ry * vy_ij + rz * vz_ij;
        T vijsignal_ij = 0.0;
        if (rv < T(0)) { vijsignal_ij = ci + c[j] - T(3) * rv / dist; }
        vijsignal_i = std::max(vijsignal_i, vijsignal_ij);

        T v1 = dist * hiInv;
        T Wi = K * hiInv3 * lookup(wh, v1);

        T termA1 = -(c11i * rx + c12i * ry + c13i * rz) * Wi;
        T termA2 = -(c12i * rx + c22i * ry + c23i * rz) * Wi;
        T termA3 = -(c13i * rx + c23i * ry + c33i * rz) * Wi;

        T volj   = xm[j] / kx[j];
        T factor = volj * (divv_i - divv[j]);

        graddivv_x += factor * termA1;
        graddivv_y += factor * termA2;
        graddivv_z += factor * termA3;
    }

    T graddivv = std::sqrt(graddivv_x * graddivv_x + graddivv_y * graddivv_y + graddivv_z * graddivv_z);

    T alphaloc = 0.0;
    if (divv_i < T(0))
    {
        T a_const = hi * hi * graddivv;
        alphaloc  = alphamax * a_const / (a_const + hi * std::abs(divv_i) + T(0.05) * ci);
    }

    if (alphaloc >= alpha_i) { alpha_i = alphaloc; }
    else
    {
        T decay    = hi / (decay_constant * vijsignal_i);
        T alphadot = 0.0;
        if (alphaloc >= alphamin) { alphadot = (alphaloc - alpha_i) / decay; }
        else { alphadot = (alphamin - alpha_i) / decay; }
        alpha_i += alphadot * dt;
    }
    return alpha_i;
}

// sph/kernels.hpp:70-84
template<class T>
inline T artificialViscosity(T alpha_i, T alpha_j, T c_i, T c_j, T w_ij)
{
    constexpr T beta         = 2.0;
    T           viscosity_ij = 0.0;
    if (w_ij < 0.0)
    {
        T vij_signal = (alpha_i + alpha_j) / 4.0 * (c_i + c_j) - beta * w_ij;
        viscosity_ij = -vij_signal * w_ij;
    }
    return viscosity_ij;
}

// momentum_energy_kern.hpp:43-63 (symv: sph/kernels.hpp:87-95 — upper-triangular product as written there)
template<class T>
T avRvCorrection(const T* R, T eta_ab, T eta_crit, const T* gi, const T* gj)
{
    auto symvDot = [&](const T* g)
    {
        T r0 = g[0] * R[0] + g[1] * R[1] + g[2] * R[2];
        T r1 = g[3] * R[1] + g[4] * R[2];
        T r2 = g[5] * R[2];
        return R[0] * r0 + (R[1] * r1 + R[2] * r2);
    };
    T dmy1 = symvDot(gi);
    T dmy2 = symvDot(gj);
    T dmy3 = T(1);
    if (eta_ab < eta_crit)
    {
        T etaDiff = T(5) * (eta_ab - eta_crit);
        dmy3      = std::exp(-etaDiff * etaDiff);
    }
    T A_ab   = (dmy2 != T(0)) ? dmy1 / dmy2 : T(0);
    T A_abp1 = T(1) + A_ab;
    T phi_ab = T(0.5) * dmy3 * std::max(T(0), std::min(T(1), T(4) * A_ab / (A_abp1 * A_abp1)));
    return -phi_ab * (dmy1 + dmy2);
}

// momentum_energy_kern.hpp:65-222; out = {ax, ay, az} (T), du (double), maxvsignal (T)
template<bool avClean, class T>
void momentumEnergyJLoop(unsigned i, double K, const Box& box, const unsigned* nb, unsigned nc, const double* x,
                         const double* y, const double* z, const T* vx, const T* vy, const T* vz, const T* h,
                         const T* m, const T* prho, const T* c, const T* c11, const T* c12, const T* c13,
                         const T* c22, const T* c23, const T* c33, T Atmin, T Atmax, T ramp, const T* wh, const T* kx,
                         const T* xm, const T* alpha, const T* dV11, const T* dV12, const T* dV13, const T* dV22,
                         const T* dV23, const T* dV33, T* axyz, double* duOut, T* maxvsignalOut)
{
    double xi = x[i], yi = y[i], zi = z[i];
    T      vxi = vx[i], vyi = vy[i], vzi = vz[i];
    T      hi = h[i], mi = m[i], ci = c[i], kxi = kx[i];
    T      alpha_i = alpha[i];
    T      xmassi  = xm[i];
    T      rhoi    = kxi * mi / xmassi;
    T      prhoi   = prho[i];

    T hiInv  = T(1) / hi;
    T hiInv3 = hiInv * hiInv * hiInv;

    T maxvsignali = 0.0;
    T momentum_x = 0.0, momentum_y = 0.0, momentum_z = 0.0, energy = 0.0;
    T a_visc_energy = 0.0;

    T c11i = c11[i], c12i = c12[i], c13i = c13[i], c22i = c22[i], c23i = c23[i], c33i = c33[i];

    T gradV_i[6] = {0, 0, 0, 0, 0, 0};
    if constexpr (avClean)
    {
        gradV_i[0] = dV11[i], gradV_i[1] = dV12[i], gradV_i[2] = dV13[i];
        gradV_i[3] = dV22[i], gradV_i[4] = dV23[i], gradV_i[5] = dV33[i];
    }

    T eta_crit = std::cbrt(T(32) * M_PI / T(3) / T(nc + 1));

    for (unsigned pj = 0; pj < nc; ++pj)
    {
        unsigned j  = nb[pj];
        T        rx = xi - x[j];
        T        ry = yi - y[j];
        T        rz = zi - z[j];
        T        vxj = vx[j], vyj = vy[j], vzj = vz[j];
        applyPBC(box, T(2) * hi, rx, ry, rz);

        T r2   = rx * rx + ry * ry + rz * rz;
        T dist = std::sqrt(r2);

        T vx_ij = vxi - vxj;
        T vy_ij = vyi - vyj;
        T vz_ij = vzi - vzj;

        T hj    = h[j];
        T hjInv = T(1) / hj;

        T v1 = dist * hiInv;
        T v2 = dist * hjInv;

        T hjInv3 = hjInv * hjInv * hjInv;
        T Wi     = hiInv3 * lookup(wh, v1);
        T Wj     = hjInv3 * lookup(wh, v2);

        T termA1_i = -(c11i * rx + c12i * ry + c13i * rz) * Wi;
        T termA2_i = -(c12i * rx + c22i * ry + c23i * rz) * Wi;
        T termA3_i = -(c13i * rx + c23i * ry + c33i * rz) * Wi;

        T c11j = c11[j], c12j = c12[j], c13j = c13[j], c22j = c22[j], c23j = c23[j], c33j = c33[j];

        T termA1_j = -(c11j * rx + c12j * ry + c13j * rz) * Wj;
        T termA2_j = -(c12j * rx + c22j * ry + c23j * rz) * Wj;
        T termA3_j = -(c13j * rx + c23j * ry + c33j * rz) * Wj;

        T mj = m[j], cj = c[j], kxj = kx[j], xmassj = xm[j];
        T rhoj = kxj * mj / xmassj;

        T rv = rx * vx_ij + ry * vy_ij + rz * vz_ij;
        if constexpr (avClean)
        {
            T R[3]  = {rx, ry, rz};
            T gj[6] = {dV11[j], dV12[j], dV13[j], dV22[j], dV23[j], dV33[j]};
            rv += avRvCorrection(R, std::min(v1, v2), eta_crit, gradV_i, gj);
        }

        T wij          = rv / dist;
        T viscosity_ij = artificialViscosity(alpha_i, alpha[j], ci, cj, wij);

        T vijsignal = T(0.5) * (ci + cj) - T(2) * wij;
        maxvsignali = (vijsignal > maxvsignali) ? vijsignal : maxvsignali;

        T a_mom, b_mom;
        T Atwood = (std::abs(rhoi - rhoj)) / (rhoi + rhoj);
        if (Atwood < Atmin)
        {
            a_mom = xmassi * xmassi;
            b_mom = xmassj * xmassj;
        }
        else if (Atwood > Atmax)
        {
            a_mom = xmassi * xmassj;
            b_mom = a_mom;
        }
        else
        {
            T sigma_ij = ramp * (Atwood - Atmin);
            // the reference calls unqualified pow() from namespace sph: with T = float this resolves to the C
            // ::pow(double, double) (checked against oracle/_ref outputs), i.e. evaluation in double
            a_mom      = ::pow(xmassi, T(2) - sigma_ij) * ::pow(xmassj, sigma_ij);
            b_mom      = ::pow(xmassj, T(2) - sigma_ij) * ::pow(xmassi, sigma_ij);
        }

        T a_visc   = mj / rhoi * viscosity_ij;
        T b_visc   = mj / rhoj * viscosity_ij;
        T a_visc_x = T(0.5) * (a_visc * termA1_i + b_visc * termA1_j);
        T a_visc_y = T(0.5) * (a_visc * termA2_i + b_visc * termA2_j);
        T a_visc_z = T(0.5) * (a_visc * termA3_i + b_visc * termA3_j);
        a_visc_energy += a_visc_x * vx_ij + a_visc_y * vy_ij + a_visc_z * vz_ij;

        energy += mj * a_mom * (vx_ij * termA1_i + vy_ij * termA2_i + vz_ij * termA3_i);

        T momentum_i = mj * prhoi * a_mom;
        T momentum_j = mj * prho[j] * b_mom;
        momentum_x += momentum_i * termA1_i + momentum_j * termA1_j + a_visc_x;
        momentum_y += momentum_i * termA2_i + momentum_j * termA2_j + a_visc_y;
        momentum_z += momentum_i * termA3_i + momentum_j * termA3_j + a_visc_z;
    }

    a_visc_energy = std::max(T(0), a_visc_energy);
    T eCoeff      = prhoi; // tdpdTrho == nullptr for the ideal-gas path
    *duOut        = K * (eCoeff * energy + T(0.5) * a_visc_energy);

    axyz[0]        = -K * momentum_x;
    axyz[1]        = -K * momentum_y;
    axyz[2]        = -K * momentum_z;
    *maxvsignalOut = maxvsignali;
}

} // namespace

/* =============================== C interface =============================== */

extern "C"
{

void orc_tables_f(double sincIndex, float* wh, float* whd, double* K) { makeTables<float>(sincIndex, wh, whd, K); }
void orc_tables_d(double sincIndex, double* wh, double* whd, double* K) { makeTables<double>(sincIndex, wh, whd, K); }

// sph_kernel_tables.hpp:62-74
double orc_sphynx_3D_k(double n)
{
    double b0 = 2.7012593e-2, b1 = 2.0410827e-2, b2 = 3.7451957e-3, b3 = 4.7013839e-2;
    return b0 + b1 * std::sqrt(n) + b2 * n + b3 * std::sqrt(n * n * n);
}

double orc_distance_sq(int pbc, double x1, double y1, double z1, double x2, double y2, double z2, const OrcBox* b)
{
    Box box(b);
    return pbc ? distanceSq<true>(x1, y1, z1, x2, y2, z2, box) : distanceSq<false>(x1, y1, z1, x2, y2, z2, box);
}

void orc_find_neighbors_f(const double* x, const double* y, const double* z, const float* h, unsigned first,
                          unsigned last, const OrcBox* b, const OrcTree* tree, unsigned ngmax, unsigned* neighbors,
                          unsigned* counts)
{
    Box box(b);
#pragma omp parallel for schedule(dynamic, 64)
    for (unsigned i = first; i < last; ++i)
    {
        counts[i - first] = findNeighborsOne(i, x, y, z, h, *tree, box, ngmax, neighbors + size_t(i - first) * ngmax);
    }
}

void orc_all2all_neighbors_f(const double* x, const double* y, const double* z, const float* h, unsigned n,
                             unsigned* neighbors, unsigned* counts, unsigned ngmax, const OrcBox* b)
{
    Box box(b);
#pragma omp parallel for
    for (unsigned i = 0; i < n; ++i)
    {
        float    radius = 2 * h[i];
        float    r2     = radius * radius;
        unsigned cnt    = 0;
        for (unsigned j = 0; j < n; ++j)
        {
            if (j == i) continue;
            if (distanceSq<true>(x[i], y[i], z[i], x[j], y[j], z[j], box) < r2)
            {
                if (cnt < ngmax) neighbors[size_t(i) * ngmax + cnt] = j;
                cnt++;
            }
        }
        counts[i] = cnt;
    }
}

float orc_update_h_f(unsigned ng0, unsigned nc, float h) { return updateH(ng0, nc, h); }

// sph/find_neighbors.hpp:11-44
unsigned long orc_find_neighbors_sph_f(const double* x, const double* y, const double* z, float* h, unsigned first,
                                       unsigned last, const OrcBox* b, const OrcTree* tree, unsigned ng0,
                                       unsigned ngmax, unsigned* neighbors, unsigned* nc)
{
    Box           box(b);
    unsigned      ngmin        = ng0 / 4;
    unsigned long numFails     = 0;
    constexpr int maxIteration = 10;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : numFails)
    for (unsigned i = 0; i < last - first; ++i)
    {
        unsigned  id    = i + first;
        unsigned* nbi   = neighbors + size_t(i) * ngmax;
        unsigned  ncSph = 1 + findNeighborsOne(id, x, y, z, h, *tree, box, ngmax, nbi);

        int iteration = 0;
        while ((ngmin > ncSph || (ncSph - 1) > ngmax) && iteration++ < maxIteration)
        {
            h[id] = updateH(ng0, ncSph, h[id]);
            ncSph = 1 + findNeighborsOne(id, x, y, z, h, *tree, box, ngmax, nbi);
        }
        numFails += (iteration >= maxIteration);
        nc[i] = ncSph;
    }
    return numFails;
}

#define NB_OF(i) (neighbors + size_t(p->ngmax) * ((i)-first))
#define NC_CAPPED(i) std::min(nc[(i)-first] - 1, p->ngmax)

// hydro_ve/xmass.hpp:39-66
void orc_xmass_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* b, const unsigned* neighbors,
                 const unsigned* nc, const double* x, const double* y, const double* z, const float* h, const float* m,
                 const float* wh, float* xm)
{
    Box box(b);
#pragma omp parallel for
    for (unsigned i = first; i < last; ++i)
        xm[i] = xmassJLoop<float>(i, p->K, box, NB_OF(i), NC_CAPPED(i), x, y, z, h, m, wh);
}

// hydro_ve/ve_def_gradh.hpp:39-85
void orc_ve_def_gradh_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* b, const unsigned* neighbors,
                        const unsigned* nc, const double* x, const double* y, const double* z, const float* h,
                        const float* m, const float* wh, const float* whd, const float* xm, float* kx, float* gradh)
{
    Box box(b);
#pragma omp parallel for
    for (unsigned i = first; i < last; ++i)
        veDefGradhJLoop<float>(i, p->K, box, NB_OF(i), NC_CAPPED(i), x, y, z, h, m, wh, whd, xm, kx + i, gradh + i);
}

// hydro_ve/eos.hpp:52-78 (temp branch) with sph/eos.hpp:18-52: cv is computed in double but returned as float
void orc_eos_ideal_temp_f(unsigned first, unsigned last, const OrcParams* p, const double* temp, const float* m,
                          const float* kx, const float* xm, const float* gradh, float* prho, float* c)
{
    const float  R     = 8.317e7;
    const float  cv    = R / p->muiConst / (p->gamma - float(1));
    const double gamma = p->gamma;
#pragma omp parallel for
    for (unsigned i = first; i < last; ++i)
    {
        float  rho = kx[i] * m[i] / xm[i];
        double u   = cv * temp[i];
        double tmp = u * (gamma - 1.0);
        double pi  = rho * tmp;
        double ci  = std::sqrt(gamma * tmp);
        prho[i]    = pi / (kx[i] * m[i] * m[i] * gradh[i]);
        c[i]       = ci;
    }
}

// hydro_ve/iad_divv_curlv.hpp:41-98 + ts_global.hpp:72-95 (rhoTimestep)
void orc_iad_divv_curlv_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* b,
                          const unsigned* neighbors, const unsigned* nc, const double* x, const double* y,
                          const double* z, const float* vx, const float* vy, const float* vz, const float* h,
                          const float* wh, const float* xm, const float* kx, float* c11, float* c12, float* c13,
                          float* c22, float* c23, float* c33, float* divv, float* curlv, double* minDtRho)
{
    Box   box(b);
    float maxDivv = -INFINITY;
#pragma omp parallel for reduction(max : maxDivv)
    for (unsigned i = first; i < last; ++i)
    {
        float cc[6], out[8];
        iadJLoop<float>(i, p->K, box, NB_OF(i), NC_CAPPED(i), x, y, z, h, wh, xm, kx, cc);
        // the second loop reads only c**[i] of the same particle, so per-particle sequencing equals the reference
        c11[i] = cc[0], c12[i] = cc[1], c13[i] = cc[2], c22[i] = cc[3], c23[i] = cc[4], c33[i] = cc[5];
        divvCurlvJLoop<float>(i, p->K, box, NB_OF(i), NC_CAPPED(i), x, y, z, vx, vy, vz, h, c11, c12, c13, c22, c23,
                              c33, wh, kx, xm, out);
        divv[i]  = out[0];
        curlv[i] = out[1];
        maxDivv  = std::max(out[0], maxDivv);
    }
    if (minDtRho) *minDtRho = p->Krho / std::abs(maxDivv);
}

// hydro_ve/av_switches.hpp:40-78
void orc_av_switches_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* b, const unsigned* neighbors,
                       const unsigned* nc, const double* x, const double* y, const double* z, const float* vx,
                       const float* vy, const float* vz, const float* h, const float* c, const float* c11,
                       const float* c12, const float* c13, const float* c22, const float* c23, const float* c33,
                       const float* wh, const float* kx, const float* xm, const float* divv, float* alpha)
{
    Box box(b);
#pragma omp parallel for
    for (unsigned i = first; i < last; ++i)
        alpha[i] = avSwitchesJLoop<float>(i, p->K, box, NB_OF(i), NC_CAPPED(i), x, y, z, vx, vy, vz, h, c, c11, c12,
                                          c13, c22, c23, c33, wh, kx, xm, divv, p->minDt, p->alphamin, p->alphamax,
                                          p->decay_constant, alpha[i]);
}

// hydro_ve/momentum_energy.hpp:40-103 + tsKCourant sph/kernels.hpp:10-16
void orc_momentum_energy_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* b,
                           const unsigned* neighbors, const unsigned* nc, const double* x, const double* y,
                           const double* z, const float* vx, const float* vy, const float* vz, const float* h,
                           const float* m, const float* prho, const float* c, const float* c11, const float* c12,
                           const float* c13, const float* c22, const float* c23, const float* c33, const float* wh,
                           const float* kx, const float* xm, const float* alpha, float* ax, float* ay, float* az,
                           double* du, double* minDtCourant)
{
    Box   box(b);
    float minDt = INFINITY;
    float Kcour = p->Kcour;
#pragma omp parallel for reduction(min : minDt)
    for (unsigned i = first; i < last; ++i)
    {
        float a[3], maxvsignal = 0;
        momentumEnergyJLoop<false, float>(i, p->K, box, NB_OF(i), NC_CAPPED(i), x, y, z, vx, vy, vz, h, m, prho, c,
                                          c11, c12, c13, c22, c23, c33, p->Atmin, p->Atmax, p->ramp, wh, kx, xm, alpha,
                                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, a, du + i, &maxvsignal);
        ax[i] = a[0], ay[i] = a[1], az[i] = a[2];
        float v    = maxvsignal > 0.0f ? maxvsignal : c[i];
        float dt_i = Kcour * h[i] / v;
        minDt      = std::min(minDt, dt_i);
    }
    if (minDtCourant) *minDtCourant = minDt;
}

/* ---- all-double single-particle entry points (unit-test shape) ---- */

double orc_xmass_jloop_d(unsigned i, double K, const OrcBox* b, const unsigned* nb, unsigned nc, const double* x,
                         const double* y, const double* z, const double* h, const double* m, const double* wh)
{
    return xmassJLoop<double>(i, K, Box(b), nb, nc, x, y, z, h, m, wh);
}

void orc_ve_def_gradh_jloop_d(unsigned i, double K, const OrcBox* b, const unsigned* nb, unsigned nc, const double* x,
                              const double* y, const double* z, const double* h, const double* m, const double* wh,
                              const double* whd, const double* xm, double* kx, double* gradh)
{
    veDefGradhJLoop<double>(i, K, Box(b), nb, nc, x, y, z, h, m, wh, whd, xm, kx, gradh);
}

void orc_iad_jloop_d(unsigned i, double K, const OrcBox* b, const unsigned* nb, unsigned nc, const double* x,
                     const double* y, const double* z, const double* h, const double* wh, const double* xm,
                     const double* kx, double* cOut)
{
    iadJLoop<double>(i, K, Box(b), nb, nc, x, y, z, h, wh, xm, kx, cOut);
}

void orc_divv_curlv_jloop_d(unsigned i, double K, const OrcBox* b, const unsigned* nb, unsigned nc, const double* x,
                            const double* y, const double* z, const double* vx, const double* vy, const double* vz,
                            const double* h, const double* c11, const double* c12, const double* c13,
                            const double* c22, const double* c23, const double* c33, const double* wh, const double* kx,
                            const double* xm, double* out)
{
    divvCurlvJLoop<double>(i, K, Box(b), nb, nc, x, y, z, vx, vy, vz, h, c11, c12, c13, c22, c23, c33, wh, kx, xm, out);
}

double orc_av_switches_jloop_d(unsigned i, double K, const OrcBox* b, const unsigned* nb, unsigned nc, const double* x,
                               const double* y, const double* z, const double* vx, const double* vy, const double* vz,
                               const double* h, const double* c, const double* c11, const double* c12,
                               const double* c13, const double* c22, const double* c23, const double* c33,
                               const double* wh, const double* kx, const double* xm, const double* divv, double dt,
                               double alphamin, double alphamax, double decay_constant, double alpha_i)
{
    return avSwitchesJLoop<double>(i, K, Box(b), nb, nc, x, y, z, vx, vy, vz, h, c, c11, c12, c13, c22, c23, c33, wh,
                                   kx, xm, divv, dt, alphamin, alphamax, decay_constant, alpha_i);
}

void orc_momentum_energy_jloop_d(int avClean, unsigned i, double K, const OrcBox* b, const unsigned* nb, unsigned nc,
                                 const double* x, const double* y, const double* z, const double* vx, const double* vy,
                                 const double* vz, const double* h, const double* m, const double* prho,
                                 const double* c, const double* c11, const double* c12, const double* c13,
                                 const double* c22, const double* c23, const double* c33, double Atmin, double Atmax,
                                 double ramp, const double* wh, const double* kx, const double* xm, const double* alpha,
                                 const double* dV11, const double* dV12, const double* dV13, const double* dV22,
                                 const double* dV23, const double* dV33, double* out)
{
    double a[3], du, mv;
    if (avClean)
        momentumEnergyJLoop<true, double>(i, K, Box(b), nb, nc, x, y, z, vx, vy, vz, h, m, prho, c, c11, c12, c13, c22,
                                          c23, c33, Atmin, Atmax, ramp, wh, kx, xm, alpha, dV11, dV12, dV13, dV22, dV23,
                                          dV33, a, &du, &mv);
    else
        momentumEnergyJLoop<false, double>(i, K, Box(b), nb, nc, x, y, z, vx, vy, vz, h, m, prho, c, c11, c12, c13,
                                           c22, c23, c33, Atmin, Atmax, ramp, wh, kx, xm, alpha, dV11, dV12, dV13, dV22,
                                           dV23, dV33, a, &du, &mv);
    out[0] = a[0], out[1] = a[1], out[2] = a[2], out[3] = du, out[4] = mv;
}

/* The momentum/energy loop over a whole particle range with T = double (neighbour list in the CPU layout,
 * nb[i * ngmax + k], nc including self as the reference stores it): what the reference computes when every operation
 * is carried out in fp64 on the same inputs. The spread between this and the production fp32 evaluation is the
 * reference's OWN rounding / summation noise; tests use it to set the comparison floors of du, ax, ay, az. */
void orc_momentum_energy_fields_d(unsigned first, unsigned last, unsigned ngmax, double K, const OrcBox* b,
                                  const unsigned* nb, const unsigned* nc, const double* x, const double* y,
                                  const double* z, const double* vx, const double* vy, const double* vz, const double* h,
                                  const double* m, const double* prho, const double* c, const double* c11,
                                  const double* c12, const double* c13, const double* c22, const double* c23,
                                  const double* c33, double Atmin, double Atmax, double ramp, const double* wh,
                                  const double* kx, const double* xm, const double* alpha, double* ax, double* ay,
                                  double* az, double* du)
{
    Box box(b);
#pragma omp parallel for schedule(static)
    for (unsigned i = first; i < last; ++i)
    {
        double   a[3], dui, mv;
        unsigned cnt = std::min(nc[i - first] - 1, ngmax);
        momentumEnergyJLoop<false, double>(i, K, box, nb + size_t(i - first) * ngmax, cnt, x, y, z, vx, vy, vz, h, m,
                                           prho, c, c11, c12, c13, c22, c23, c33, Atmin, Atmax, ramp, wh, kx, xm, alpha,
                                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, a, &dui, &mv);
        ax[i - first] = a[0], ay[i - first] = a[1], az[i - first] = a[2], du[i - first] = dui;
    }
}

/* --- turbulence stirring ------------------------------------------------------------------------------------------- */

// stirring.hpp:106-125 (loop over the groups' particles) with stirParticle :45-83; Tc = T = double, Ta = float
void orc_compute_stirring(unsigned first, unsigned last, const double* x, const double* y, const double* z, float* ax,
                          float* ay, float* az, unsigned numModes, const double* modes, const double* phaseReal,
                          const double* phaseImag, const double* amplitudes, double solWeightNorm)
{
#pragma omp parallel for schedule(static)
    for (unsigned i = first; i < last; ++i)
    {
        float sum[3] = {0.f, 0.f, 0.f}; // Ta turbAx = 0.0 ...
        for (unsigned m = 0; m < numModes; ++m)
        {
            const double* k = modes + 3 * m;
            // stirring.hpp:56-63: z first, then y, then x
            double cz = std::cos(k[2] * z[i]), sz = std::sin(k[2] * z[i]);
            double cy = std::cos(k[1] * y[i]), sy = std::sin(k[1] * y[i]);
            double cx = std::cos(k[0] * x[i]), sx = std::sin(k[0] * x[i]);
            // real and imaginary part of exp(i k.x) (:68-72)
            double re = (cx * cy - sx * sy) * cz - (sx * cy + cx * sy) * sz;
            double im = cx * (cy * sz + sy * cz) + sx * (cy * cz - sy * sz);
            for (int d = 0; d < 3; ++d)
            {
                // Ta += double expression (:74-76): added in double, rounded to float
                sum[d] = float(double(sum[d]) + amplitudes[m] * (phaseReal[3 * m + d] * re - phaseImag[3 * m + d] * im));
            }
        }
        ax[i] = float(double(ax[i]) + solWeightNorm * double(sum[0])); // :118-120
        ay[i] = float(double(ay[i]) + solWeightNorm * double(sum[1]));
        az[i] = float(double(az[i]) + solWeightNorm * double(sum[2]));
    }
}

// phases.hpp:46-72
void orc_compute_phases(unsigned numModes, const double* ou, double solWeight, const double* modes, double* phasesReal,
                        double* phasesImag)
{
    for (unsigned i = 0; i < numModes; ++i)
    {
        double ka = 0.0, kb = 0.0, kk = 0.0;
        for (int j = 0; j < 3; ++j)
        {
            kk = kk + modes[3 * i + j] * modes[3 * i + j];
            ka = ka + modes[3 * i + j] * ou[6 * i + 2 * j + 1];
            kb = kb + modes[3 * i + j] * ou[6 * i + 2 * j];
        }
        for (int j = 0; j < 3; ++j)
        {
            double diva  = modes[3 * i + j] * ka / kk;
            double divb  = modes[3 * i + j] * kb / kk;
            double curla = ou[6 * i + 2 * j] - divb;
            double curlb = ou[6 * i + 2 * j + 1] - diva;
            phasesReal[3 * i + j] = solWeight * curla + (1.0 - solWeight) * divb;
            phasesImag[3 * i + j] = solWeight * curlb + (1.0 - solWeight) * diva;
        }
    }
}

// driver.hpp:85-98 with the Gaussian deviates given
void orc_update_noise(unsigned n, double* phases, double stddev, double dt, double ts, const double* gaussians)
{
    double dampingA = std::exp(-dt / ts);
    double dampingB = std::sqrt(1.0 - dampingA * dampingA);
    for (unsigned i = 0; i < n; ++i)
    {
        phases[i] = phases[i] * dampingA + stddev * dampingB * gaussians[i];
    }
}

} // extern "C"

/*! @file
 * TEST INFRASTRUCTURE ONLY. Re-drives the reference's own SPH-VE known-answer tests (sph/test/ve.cpp:52-233) with the
 * UNMODIFIED reference headers on sph/test/example_data.txt and prints every J-loop output with 17 significant digits.
 * tests/golden/make_golden.py stores inputs + these outputs + the literal expectations of ve.cpp in ve_kat.npz.
 *
 * Usage: ref_kat <path/to/example_data.txt>
 */
#include <cstdio>
#include <numeric>
#include <vector>

#include "cstone/util/tuple_util.hpp"
#include "sph/hydro_ve/av_switches_kern.hpp"
#include "sph/hydro_ve/divv_curlv_kern.hpp"
#include "sph/hydro_ve/iad_kern.hpp"
#include "sph/hydro_ve/momentum_energy_kern.hpp"
#include "sph/hydro_ve/ve_def_gradh_kern.hpp"
#include "sph/hydro_ve/xmass_kern.hpp"
#include "sph/sph_kernel_tables.hpp"
#include "sph/table_lookup.hpp"

using namespace sph;
using T = double;

int main(int argc, char** argv)
{
    if (argc < 2) return 1;
    const unsigned npart = 99, ncols = 31;
    std::vector<std::vector<T>> col(ncols, std::vector<T>(npart));
    FILE* f = fopen(argv[1], "r");
    if (!f) return 2;
    for (unsigned i = 0; i < npart; ++i)
        for (unsigned k = 0; k < ncols; ++k)
            if (fscanf(f, "%lf", &col[k][i]) != 1) return 3;
    fclose(f);

    // column order of ve.cpp:77-80
    auto &x = col[0], &y = col[1], &z = col[2], &vx = col[3], &vy = col[4], &vz = col[5], &h = col[6], &c = col[7],
         &c11 = col[8], &c12 = col[9], &c13 = col[10], &c22 = col[11], &c23 = col[12], &c33 = col[13], &p = col[14],
         &gradh = col[15], &rho0 = col[16], &dvxdx = col[19], &dvxdy = col[20], &dvxdz = col[21], &dvydx = col[22],
         &dvydy = col[23], &dvydz = col[24], &dvzdx = col[25], &dvzdy = col[26], &dvzdz = col[27], &alpha = col[28],
         &divv = col[30];

    T sincIndex = 6.0;
    auto wh  = tabulateFunction<T, lt::kTableSize>(getSphKernel(SphKernelType::sinc_n, sincIndex), 0.0, 2.0);
    auto whd = tabulateFunction<T, lt::kTableSize>(getSphKernelDerivative(SphKernelType::sinc_n, sincIndex), 0.0, 2.0);

    T K = sphynx_3D_k(sincIndex), alphamin = 0.05, alphamax = 1.0, decay_constant = 0.2, mpart = 3.781038064465603e26,
      dt = 0.3, Atmin = 0.1, Atmax = 0.2, ramp = 1.0 / (Atmax - Atmin);

    std::vector<T> m(npart, mpart), xm(npart), kx(npart), prho(npart);
    for (unsigned i = 0; i < npart; i++)
    {
        xm[i]   = mpart / rho0[i];
        kx[i]   = K * xm[i] / std::pow(h[i], 3);
        prho[i] = p[i] / (kx[i] * m[i] * m[i] * gradh[i]);
    }
    std::vector<cstone::LocalIndex> nb(npart - 1);
    std::iota(nb.begin(), nb.end(), 1);
    unsigned nc = npart - 1;
    cstone::Box<T> box(-1.e9, 1.e9, cstone::BoundaryType::open);

    printf("K %.17g\n", K);
    printf("K_simpson %.17g\n", kernel_3D_k(getSphKernel(SphKernelType::sinc_n, sincIndex), 2.0));
    printf("wh_1000 %.17g\nwhd_1000 %.17g\n", wh[1000], whd[1000]);

    T a = AVswitchesJLoop(0, K, box, nb.data(), nc, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(),
                          h.data(), c.data(), c11.data(), c12.data(), c13.data(), c22.data(), c23.data(), c33.data(),
                          wh.data(), whd.data(), kx.data(), xm.data(), divv.data(), dt, alphamin, alphamax,
                          decay_constant, alpha[0]);
    printf("av_alpha %.17g\n", a);

    T o[8];
    divV_curlVJLoop(0, K, box, nb.data(), nc, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(), h.data(),
                    c11.data(), c12.data(), c13.data(), c22.data(), c23.data(), c33.data(), wh.data(), whd.data(),
                    kx.data(), xm.data(), &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7], true);
    const char* dn[8] = {"divv", "curlv", "dV11", "dV12", "dV13", "dV22", "dV23", "dV33"};
    for (int k = 0; k < 8; ++k)
        printf("dc_%s %.17g\n", dn[k], o[k]);

    std::vector<T> iad(6, -1);
    IADJLoop(0, K, box, nb.data(), nc, x.data(), y.data(), z.data(), h.data(), wh.data(), whd.data(), xm.data(),
             kx.data(), &iad[0], &iad[1], &iad[2], &iad[3], &iad[4], &iad[5]);
    for (int k = 0; k < 6; ++k)
        printf("iad_%d %.17g\n", k, iad[k]);

    std::vector<T> dV11(npart), dV12(npart), dV13(npart), dV22(npart), dV23(npart), dV33(npart);
    for (unsigned i = 0; i < npart; ++i)
    {
        dV11[i] = dvxdx[i];
        dV12[i] = dvxdy[i] + dvydx[i];
        dV13[i] = dvxdz[i] + dvzdx[i];
        dV22[i] = dvydy[i];
        dV23[i] = dvydz[i] + dvzdy[i];
        dV33[i] = dvzdz[i];
    }
    for (int clean = 1; clean >= 0; --clean)
    {
        T du = -1, gx = -1, gy = -1, gz = -1, mv = -1;
        if (clean)
            momentumAndEnergyJLoop<true>(0, K, box, nb.data(), nc, x.data(), y.data(), z.data(), vx.data(), vy.data(),
                                         vz.data(), h.data(), m.data(), prho.data(), (const T*)nullptr, c.data(),
                                         c11.data(), c12.data(), c13.data(), c22.data(), c23.data(), c33.data(), Atmin,
                                         Atmax, ramp, wh.data(), kx.data(), xm.data(), alpha.data(), dV11.data(),
                                         dV12.data(), dV13.data(), dV22.data(), dV23.data(), dV33.data(), &gx, &gy, &gz,
                                         &du, &mv);
        else
            momentumAndEnergyJLoop<false>(0, K, box, nb.data(), nc, x.data(), y.data(), z.data(), vx.data(), vy.data(),
                                          vz.data(), h.data(), m.data(), prho.data(), (const T*)nullptr, c.data(),
                                          c11.data(), c12.data(), c13.data(), c22.data(), c23.data(), c33.data(), Atmin,
                                          Atmax, ramp, wh.data(), kx.data(), xm.data(), alpha.data(), dV11.data(),
                                          dV12.data(), dV13.data(), dV22.data(), dV23.data(), dV33.data(), &gx, &gy,
                                          &gz, &du, &mv);
        printf("mom%d_ax %.17g\nmom%d_ay %.17g\nmom%d_az %.17g\nmom%d_du %.17g\nmom%d_maxvsignal %.17g\n", clean, gx,
               clean, gy, clean, gz, clean, du, clean, mv);
    }

    auto [kx0, gradh0] = veDefGradhJLoop(0, K, box, nb.data(), nc, x.data(), y.data(), z.data(), h.data(), m.data(),
                                         wh.data(), whd.data(), xm.data());
    printf("gradh_kx %.17g\ngradh_gradh %.17g\n", kx0, gradh0);

    T xmass = xmassJLoop(0, K, box, nb.data(), nc, x.data(), y.data(), z.data(), h.data(), m.data(), wh.data(),
                         whd.data());
    printf("xmass %.17g\n", xmass);
    return 0;
}

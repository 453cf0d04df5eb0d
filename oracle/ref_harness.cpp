/*! @file
 * TEST INFRASTRUCTURE ONLY — never linked into, imported or called by the product library.
 *
 * Drives the UNMODIFIED reference CPU implementation (headers under /root/reference, included where they
 * lie; nothing is copied) for the SPH-VE hydro step and dumps full-precision binary per-particle arrays
 * after every stage, so that the CUDA path and the C restatement (oracle/sphx_oracle.cpp) can be compared
 * against the reference itself.
 *
 * Stage order replicates HydroVeProp::computeForces / integrate
 *   (main/src/propagator/ve_hydro.hpp:130-215) including the release/acquire aliasing, with a snapshot of every
 *   output right after the stage that produces it (gradh, divv, curlv are recycled before computeForces returns).
 *
 * Usage: ref_harness <case> <n> <steps> <outdir> [dumpEvery=1] [dumpNeighbors=1] [hscale=1] [avClean=0] [stir=0]
 *   stir=1|2 (case turb only): the reference's turbulence-ve propagator (main/src/propagator/turb_ve.hpp:67-72):
 *   particles start at rest, sph::driveTurbulence runs after computeMomentumEnergy; 1 = TurbulenceConstants() as they
 *   are (parabolic spectrum), 2 = stSpectForm 2 (power law, modes drawn from the random engine). Dumps the
 *   TurbulenceData state and the accelerations after stirring (stir_ax, stir_ay, stir_az).
 *   avClean=1 runs HydroVeProp<true,...>: dV11..dV33 active (ve_hydro.hpp:78-83), computeMomentumEnergy<true>
 *   hscale != 1 perturbs the initial smoothing lengths (h*=hscale if id%3==0, h/=hscale if id%3==1) so that the
 *   coupled h / neighbour-count iteration of sph/find_neighbors.hpp:17-36 is exercised in both directions
 *   case: sedov | noh | turb ; n: cube side; steps: number of hydro steps;
 *   writes <outdir>/step<k>/<name>.bin + <outdir>/step<k>/manifest.txt, and <outdir>/energies.txt
 */

#include <mpi.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "cstone/domain/domain.hpp"
#include "sph/particles_data.hpp"
#include "sph/sph.hpp"
#include "sphexa/simulation_data.hpp"
#include "init/sedov_init.hpp"
#include "init/noh_init.hpp"
#include "init/turbulence_init.hpp"
#include "observables/conserved_quantities.hpp"
#include "sph/hydro_turb/driver.hpp"

using namespace sphexa;
using namespace sph;

using Dataset = SimulationData<cstone::CpuTag>;
using Domain  = cstone::Domain<uint64_t, double, cstone::CpuTag>;
using T       = double;
namespace fs  = std::filesystem;

struct Dumper
{
    fs::path      dir;
    std::ofstream manifest;

    explicit Dumper(const fs::path& d)
        : dir(d)
    {
        fs::create_directories(dir);
        manifest.open(dir / "manifest.txt");
    }

    template<class V>
    void put(const std::string& name, const V* ptr, size_t n)
    {
        const char* dt = nullptr;
        if constexpr (std::is_same_v<V, double>) dt = "f8";
        else if constexpr (std::is_same_v<V, float>)
            dt = "f4";
        else if constexpr (std::is_same_v<V, unsigned>)
            dt = "u4";
        else if constexpr (std::is_same_v<V, int>)
            dt = "i4";
        else if constexpr (std::is_same_v<V, uint64_t> || std::is_same_v<V, unsigned long>)
            dt = "u8";
        else
            static_assert(sizeof(V) == 0, "unsupported dump type");
        std::ofstream f(dir / (name + ".bin"), std::ios::binary);
        f.write(reinterpret_cast<const char*>(ptr), n * sizeof(V));
        manifest << name << " " << dt << " " << n << "\n";
    }

    template<class V>
    void put(const std::string& name, const std::vector<V>& v, size_t first, size_t last)
    {
        put(name, v.data() + first, last - first);
    }

    template<class V>
    void scalar(const std::string& name, V v)
    {
        put(name, &v, 1);
    }
};

//! jittered lattice in (-r, r)^3: lattice point + U(-0.2,0.2)*step per coordinate, mt19937_64(42), x then y then z
static void jitteredLattice(double r, size_t side, std::vector<T>& x, std::vector<T>& y, std::vector<T>& z)
{
    size_t n = side * side * side;
    x.resize(n);
    y.resize(n);
    z.resize(n);
    regularGrid(r, side, 0, n, x, y, z);
    std::mt19937_64                        gen(42);
    std::uniform_real_distribution<double> dist(-0.2, 0.2);
    double                                 step = 2 * r / side;
    for (size_t i = 0; i < n; ++i)
    {
        x[i] += dist(gen) * step;
        y[i] += dist(gen) * step;
        z[i] += dist(gen) * step;
    }
}

int main(int argc, char** argv)
{
    if (argc < 5)
    {
        std::cerr << "usage: ref_harness <sedov|noh|turb> <n> <steps> <outdir> [dumpEvery=1] [dumpNeighbors=1]\n";
        return 1;
    }
    std::string testCase  = argv[1];
    size_t      cubeSide  = std::stoul(argv[2]);
    int         numSteps  = std::stoi(argv[3]);
    fs::path    outDir    = argv[4];
    int         dumpEvery = argc > 5 ? std::stoi(argv[5]) : 1;
    bool        dumpNb    = argc > 6 ? std::stoi(argv[6]) != 0 : true;
    double      hscale    = argc > 7 ? std::stod(argv[7]) : 1.0;
    bool        avClean   = argc > 8 ? std::stoi(argv[8]) != 0 : false;
    int         stir      = argc > 9 ? std::stoi(argv[9]) : 0;

    std::unique_ptr<sph::TurbulenceData<double, cstone::CpuTag>> turb;

    MPI_Init(&argc, &argv);
    fs::create_directories(outDir);

    Dataset simData;
    simData.comm = MPI_COMM_WORLD;
    auto& d      = simData.hydro;

    // field activation as HydroVeProp<false,...>::activateFields (ve_hydro.hpp:99-112)
    d.setConserved("x", "y", "z", "h", "m");
    d.setDependent("keys");
    d.setConserved("temp", "vx", "vy", "vz", "x_m1", "y_m1", "z_m1", "du_m1", "alpha", "id");
    d.setDependent("ax", "ay", "az", "prho", "c", "du", "c11", "c12", "c13", "c22", "c23", "c33", "xm", "kx", "nc");
    if (avClean) { d.setDependent("dV11", "dV12", "dV13", "dV22", "dV23", "dV33"); }

    cstone::Box<T> box(0, 1);
    if (testCase == "sedov")
    {
        SedovGrid<Dataset> init;
        box = init.init(0, 1, cubeSide, simData, nullptr);
    }
    else if (testCase == "noh" || testCase == "turb")
    {
        std::vector<T> x, y, z;
        double         r = 0.5;
        jitteredLattice(r, cubeSide, x, y, z);
        InitSettings settings;
        if (testCase == "noh")
        {
            Dataset tmp;
            settings = buildSettings(tmp, nohConstants(), "", nullptr);
            box      = cstone::Box<T>(-r, r, cstone::BoundaryType::open);
            cutSphere(r, x, y, z); // init/grid.hpp:276
        }
        else
        {
            Dataset tmp;
            settings = buildSettings(tmp, TurbulenceConstants(), "", nullptr);
            box      = cstone::Box<T>(-r, r, cstone::BoundaryType::periodic);
            // fold jittered points back into the periodic box
            for (size_t i = 0; i < x.size(); ++i)
            {
                auto X = cstone::putInBox(cstone::Vec3<T>{x[i], y[i], z[i]}, box);
                x[i] = X[0], y[i] = X[1], z[i] = X[2];
            }
        }
        size_t numParticlesGlobal = x.size();
        d.x                       = x;
        d.y                       = y;
        d.z                       = z;
        syncCoords<uint64_t>(0, 1, numParticlesGlobal, d.x, d.y, d.z, box);
        d.resize(d.x.size());
        settings["numParticlesGlobal"] = double(numParticlesGlobal);
        BuiltinWriter attributeSetter(settings);
        d.loadOrStoreAttributes(&attributeSetter);
        if (testCase == "noh") { initNohFields(d, settings); }
        else
        {
            initTurbulenceHydroFields(d, settings);
            if (stir)
            {
                if (stir == 2) { settings["stSpectForm"] = 2; }
                turb = std::make_unique<sph::TurbulenceData<double, cstone::CpuTag>>(settings, false);
            }
            // without stirring: impose a deterministic subsonic solenoidal velocity field
            double cs = std::sqrt(d.gamma * (d.gamma - 1.0) * settings.at("u0"));
            for (size_t i = 0; i < d.x.size() && !stir; ++i)
            {
                d.vx[i]   = 0.3 * cs * std::sin(2 * M_PI * d.y[i]);
                d.vy[i]   = 0.3 * cs * std::sin(2 * M_PI * d.z[i]);
                d.vz[i]   = 0.3 * cs * std::sin(2 * M_PI * d.x[i]);
                d.x_m1[i] = d.vx[i] * d.minDt;
                d.y_m1[i] = d.vy[i] * d.minDt;
                d.z_m1[i] = d.vz[i] * d.minDt;
            }
        }
    }
    else
    {
        std::cerr << "unknown case " << testCase << "\n";
        return 1;
    }

    if (hscale != 1.0)
    {
        for (size_t i = 0; i < d.h.size(); ++i)
        {
            if (d.id[i] % 3 == 0) { d.h[i] *= hscale; }
            else if (d.id[i] % 3 == 1) { d.h[i] /= hscale; }
        }
    }

    uint64_t bucketSizeFocus = 64;
    uint64_t bucketSize      = std::max(bucketSizeFocus, d.numParticlesGlobal / 100);
    Domain   domain(0, 1, bucketSize, bucketSizeFocus, 1.0f, box);
    domain.setGrowthAllocRate(d.getAllocGrowthRate());

    auto sync = [&]()
    {
        auto conserved = std::tuple_cat(
            std::tie(get<"m">(d)), get<"temp", "vx", "vy", "vz", "x_m1", "y_m1", "z_m1", "du_m1", "alpha", "id">(d));
        if (avClean)
        {
            domain.sync(get<"keys">(d), get<"x">(d), get<"y">(d), get<"z">(d), get<"h">(d), conserved,
                        get<"ax", "ay", "az", "prho", "c", "du", "c11", "c12", "c13", "c22", "c23", "c33", "xm", "kx",
                            "nc", "dV11", "dV12", "dV13", "dV22", "dV23", "dV33">(d));
        }
        else
        {
            domain.sync(get<"keys">(d), get<"x">(d), get<"y">(d), get<"z">(d), get<"h">(d), conserved,
                        get<"ax", "ay", "az", "prho", "c", "du", "c11", "c12", "c13", "c22", "c23", "c33", "xm", "kx",
                            "nc">(d));
        }
        d.treeView = domain.octreeProperties();
    };

    // first sync as in sphexa.cpp:141
    sync();

    std::ofstream energies(outDir / "energies.txt");
    energies.precision(17);
    GroupData<cstone::CpuTag> groups;

    for (int step = 0; step < numSteps; ++step, d.iteration++)
    {
        auto t0 = std::chrono::steady_clock::now();
        sync();
        box = domain.box();
        d.resizeAcc(domain.nParticlesWithHalos());
        resizeNeighbors(d, domain.nParticles() * d.ngmax);
        size_t first = domain.startIndex();
        size_t last  = domain.endIndex();
        size_t n     = last - first;

        bool                    doDump = (step % dumpEvery == 0) || step == numSteps - 1;
        std::unique_ptr<Dumper> dump;
        if (doDump) { dump = std::make_unique<Dumper>(outDir / ("step" + std::to_string(step))); }

        if (doDump)
        {
            // inputs of the step (SFC-sorted, after domain sync, before the h-iteration)
            dump->scalar("n", uint64_t(n));
            dump->scalar("first", uint64_t(first));
            dump->scalar("last", uint64_t(last));
            dump->scalar("ng0", d.ng0);
            dump->scalar("ngmax", d.ngmax);
            double boxv[6] = {box.xmin(), box.xmax(), box.ymin(), box.ymax(), box.zmin(), box.zmax()};
            dump->put("box", boxv, 6);
            int bnd[3] = {int(box.boundaryX()), int(box.boundaryY()), int(box.boundaryZ())};
            dump->put("boundary", bnd, 3);
            double par[16] = {d.K,        d.Kcour,          d.Krho,  d.gamma, double(d.muiConst), d.minDt, d.minDt_m1,
                              d.alphamin, d.alphamax,       d.decay_constant, d.Atmin, d.Atmax,   d.ramp,  d.ttot,
                              double(d.eosChoice), d.maxDtIncrease};
            dump->put("params", par, 16);
            dump->put("wh", d.wh.data(), d.wh.size());
            dump->put("whd", d.whd.data(), d.whd.size());
            dump->put("keys", d.keys, first, last);
            dump->put("x", d.x, first, last);
            dump->put("y", d.y, first, last);
            dump->put("z", d.z, first, last);
            dump->put("h_in", d.h, first, last);
            dump->put("m", d.m, first, last);
            dump->put("vx", d.vx, first, last);
            dump->put("vy", d.vy, first, last);
            dump->put("vz", d.vz, first, last);
            dump->put("temp", d.temp, first, last);
            dump->put("alpha_in", d.alpha, first, last);
            dump->put("x_m1", d.x_m1, first, last);
            dump->put("y_m1", d.y_m1, first, last);
            dump->put("z_m1", d.z_m1, first, last);
            dump->put("du_m1", d.du_m1, first, last);
            dump->put("id", d.id, first, last);
            // tree view (OctreeNsView, tree/octree.hpp:279-300)
            const auto&           tv       = d.treeView;
            int                   numLeaf  = tv.numLeafNodes;
            int                   maxLevel = cstone::maxTreeLevel<uint64_t>{};
            int                   numNodes = tv.levelRange[maxLevel + 1];
            dump->scalar("numLeafNodes", numLeaf);
            dump->scalar("numNodes", numNodes);
            dump->put("tree_prefixes", tv.prefixes, numNodes);
            dump->put("tree_childOffsets", tv.childOffsets, numNodes);
            dump->put("tree_internalToLeaf", tv.internalToLeaf, numNodes);
            dump->put("tree_levelRange", tv.levelRange, maxLevel + 2);
            dump->put("tree_leaves", tv.leaves, numLeaf + 1);
            dump->put("tree_layout", tv.layout, numLeaf + 1);
            dump->put("tree_centers", reinterpret_cast<const double*>(tv.centers), size_t(numNodes) * 3);
            dump->put("tree_sizes", reinterpret_cast<const double*>(tv.sizes), size_t(numNodes) * 3);
        }

        fill(get<"m">(d), 0, first, d.m[first]);
        fill(get<"m">(d), last, domain.nParticlesWithHalos(), d.m[first]);

        auto t1 = std::chrono::steady_clock::now();
        findNeighborsSfc(first, last, d, box);
        computeGroups(first, last, d, box, groups);
        auto t2 = std::chrono::steady_clock::now();

        if (doDump)
        {
            dump->put("h", d.h, first, last);
            dump->put("nc", d.nc, first, last);
            if (dumpNb) { dump->put("neighbors", d.neighbors.data(), n * d.ngmax); }
        }

        computeXMass(groups.view(), d, box);
        if (doDump) dump->put("xm", d.xm, first, last);

        release(d, "ay");
        acquire(d, "gradh");
        computeVeDefGradh(groups.view(), d, box);
        if (doDump)
        {
            dump->put("kx", d.kx, first, last);
            dump->put("gradh", d.gradh, first, last);
        }

        computeEOS(first, last, d);
        if (doDump)
        {
            dump->put("prho", d.prho, first, last);
            dump->put("c", d.c, first, last);
        }

        release(d, "gradh", "az");
        acquire(d, "divv", "curlv");
        computeIadDivvCurlv(groups.view(), d, box);
        d.minDtRho = rhoTimestep(first, last, d);
        if (doDump)
        {
            dump->put("c11", d.c11, first, last);
            dump->put("c12", d.c12, first, last);
            dump->put("c13", d.c13, first, last);
            dump->put("c22", d.c22, first, last);
            dump->put("c23", d.c23, first, last);
            dump->put("c33", d.c33, first, last);
            dump->put("divv", d.divv, first, last);
            dump->put("curlv", d.curlv, first, last);
            if (avClean)
            {
                dump->put("dV11", d.dV11, first, last);
                dump->put("dV12", d.dV12, first, last);
                dump->put("dV13", d.dV13, first, last);
                dump->put("dV22", d.dV22, first, last);
                dump->put("dV23", d.dV23, first, last);
                dump->put("dV33", d.dV33, first, last);
            }
        }

        computeAVswitches(groups.view(), d, box);
        if (doDump) dump->put("alpha", d.alpha, first, last);

        release(d, "divv", "curlv");
        acquire(d, "ay", "az");
        if (avClean) { computeMomentumEnergy<true>(groups.view(), nullptr, d, box); }
        else { computeMomentumEnergy<false>(groups.view(), nullptr, d, box); }
        auto t3 = std::chrono::steady_clock::now();
        if (doDump)
        {
            dump->put("ax", d.ax, first, last);
            dump->put("ay", d.ay, first, last);
            dump->put("az", d.az, first, last);
            dump->put("du", d.du, first, last);
            double dts[2] = {d.minDtCourant, d.minDtRho};
            dump->put("dts", dts, 2);
        }

        if (turb)
        {
            if (doDump && step == 0) { dump->put("turb_phases_in", turb->phases.data(), turb->phases.size()); }
            driveTurbulence(groups.view(), d, *turb);
            if (doDump)
            {
                dump->put("stir_ax", d.ax, first, last);
                dump->put("stir_ay", d.ay, first, last);
                dump->put("stir_az", d.az, first, last);
                dump->put("turb_modes", turb->modes.data(), turb->modes.size());
                dump->put("turb_amplitudes", turb->amplitudes.data(), turb->amplitudes.size());
                dump->put("turb_phases", turb->phases.data(), turb->phases.size());
                dump->put("turb_phasesReal", turb->phasesReal.data(), turb->phasesReal.size());
                dump->put("turb_phasesImag", turb->phasesImag.data(), turb->phasesImag.size());
                double sc[4] = {turb->variance, turb->decayTime, turb->solWeight, turb->solWeightNorm};
                dump->put("turb_scalars", sc, 4);
            }
        }

        computeConservedQuantities(first, last, d, MPI_COMM_WORLD);
        energies << step << " " << d.ttot << " " << d.minDt << " " << d.etot << " " << d.ecin << " " << d.eint << " "
                 << d.linmom << " " << d.angmom << " " << d.totalNeighbors << "\n";

        // integrate (ve_hydro.hpp:206-215)
        computeTimestep(first, last, d);
        computePositions(groups.view(), d, box, d.minDt, {float(d.minDt_m1)});
        updateSmoothingLength(groups.view(), d);
        auto t4 = std::chrono::steady_clock::now();

        if (doDump)
        {
            double post[3] = {d.minDt, d.minDt_m1, d.ttot};
            dump->put("post_dt", post, 3);
            dump->put("post_x", d.x, first, last);
            dump->put("post_y", d.y, first, last);
            dump->put("post_z", d.z, first, last);
            dump->put("post_vx", d.vx, first, last);
            dump->put("post_vy", d.vy, first, last);
            dump->put("post_vz", d.vz, first, last);
            dump->put("post_h", d.h, first, last);
            dump->put("post_temp", d.temp, first, last);
            dump->put("post_x_m1", d.x_m1, first, last);
            dump->put("post_y_m1", d.y_m1, first, last);
            dump->put("post_z_m1", d.z_m1, first, last);
            dump->put("post_du_m1", d.du_m1, first, last);
        }

        auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
        std::printf("step %d n %zu sync %.4f findNeighbors %.4f loops %.4f integrate %.4f etot %.10g ecin %.10g eint "
                    "%.10g dt %.6g nbsum %zu\n",
                    step, n, secs(t0, t1), secs(t1, t2), secs(t2, t3), secs(t3, t4), d.etot, d.ecin, d.eint, d.minDt,
                    size_t(d.totalNeighbors));
        std::fflush(stdout);
    }

    MPI_Finalize();
    return 0;
}

/*! @file
 * Test driver of the C++20 facade include/sphx_domain.hpp: the cstone::Domain call shape (sync, exchangeHalos, startIndex,
 * endIndex, nParticlesWithHalos, box, octreeProperties) on one or two ranks, followed by the cstone::findNeighbors call
 * shape (sphx_find_neighbors) on the synced arrays. Mirrors domain/test/integration_mpi/domain_nranks.cpp:64-131 of the
 * reference: random particles, every rank starts with a slice, after sync the neighbour counts of the assigned
 * particles are written by particle id; the pytest wrapper compares N ranks against one rank and against brute force.
 *
 * usage: domain_facade_test <rank> <nranks> <idfile> <outfile> <numParticles> <pbc 0|1>
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include "sphx_domain.hpp"

using sphx::DeviceVector;

static void must(int rc, const char* what)
{
    if (rc != SPHX_OK)
    {
        std::fprintf(stderr, "%s failed: %s\n", what, sphx_last_error());
        std::exit(2);
    }
}

int main(int argc, char** argv)
{
    if (argc < 7) return 1;
    const int         rank = std::atoi(argv[1]), nranks = std::atoi(argv[2]);
    const std::string idfile = argv[3], outfile = argv[4];
    const size_t      N   = std::strtoul(argv[5], nullptr, 10);
    const int         pbc = std::atoi(argv[6]);
    cudaSetDevice(rank);

    SphxComm* comm = nullptr;
    if (nranks > 1)
    {
        char id[SPHX_UNIQUE_ID_BYTES];
        if (rank == 0)
        {
            must(sphx_comm_unique_id(id), "sphx_comm_unique_id");
            std::ofstream(idfile + ".tmp", std::ios::binary).write(id, sizeof(id));
            std::rename((idfile + ".tmp").c_str(), idfile.c_str());
        }
        else
        {
            for (int k = 0; k < 600; ++k)
            {
                std::ifstream f(idfile, std::ios::binary);
                if (f && f.read(id, sizeof(id))) break;
                std::this_thread::sleep_for(std::chrono::milliseconds(100));
            }
        }
        must(sphx_comm_init(&comm, rank, nranks, id), "sphx_comm_init");
    }

    // the same global particle set on every rank; rank r starts with the r-th slice of the generation order
    std::mt19937_64                        gen(42);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::vector<double>                    gx(N), gy(N), gz(N);
    for (size_t i = 0; i < N; ++i)
    {
        // clustered: half of the particles in a denser blob
        double s = (i % 2) ? 0.25 : 1.0, o = (i % 2) ? 0.3 : 0.0;
        gx[i] = o + s * u(gen), gy[i] = o + s * u(gen), gz[i] = o + s * u(gen);
    }
    const float  hval = 0.5f * float(std::cbrt(3.0 / (4.0 * M_PI) * 60.0 / double(N)));
    const size_t b = rank * N / nranks, e = (rank + 1) * N / nranks, n = e - b;
    std::vector<double>   hx(gx.begin() + b, gx.begin() + e), hy(gy.begin() + b, gy.begin() + e), hz(gz.begin() + b, gz.begin() + e);
    std::vector<float>    hh(n), hm(n, 1.0f / float(N));
    std::vector<uint64_t> hid(n);
    for (size_t i = 0; i < n; ++i)
    {
        hid[i] = b + i;
        hh[i]  = hval * ((b + i) % 2 ? 0.6f : 1.0f); // smaller h in the blob
    }
    DeviceVector<double>   x(hx), y(hy), z(hz);
    DeviceVector<float>    h(hh), m(hm), s1, s2;
    DeviceVector<uint64_t> id(hid), keys;
    DeviceVector<double>   sx, sy, sz; // scratch of the coordinate type: used as spares; m and id get spares of the domain's own

    SphxBox box{{0, 1, 0, 1, 0, 1}, {pbc, pbc, pbc}};
    sphx::Domain<uint64_t, double> domain(rank, nranks, 64, 64, 0.5f, box, comm);

    for (int iter = 0; iter < 2; ++iter) // second sync: arrays already carry halos, particles at [startIndex, endIndex)
    {
        domain.sync(keys, x, y, z, h, std::tie(m, id), std::tie(sx, sy, sz, s1, s2));
        if (x.size() != domain.nParticlesWithHalos() || keys.size() != x.size() || id.size() != x.size())
        {
            std::fprintf(stderr, "array sizes after sync\n");
            return 3;
        }
    }
    // halo values of another field through exchangeHalos: a copy of h
    DeviceVector<float> h2(h.size());
    cudaMemset(h2.data(), 0, h2.size() * 4);
    cudaMemcpy(h2.data() + domain.startIndex(), h.data() + domain.startIndex(), domain.nParticles() * 4,
               cudaMemcpyDeviceToDevice);
    DeviceVector<char> sendBuf, recvBuf;
    domain.exchangeHalos(std::tie(h2), sendBuf, recvBuf);
    cudaDeviceSynchronize();
    auto hh1 = h.toHost(), hh2 = h2.toHost();
    for (size_t i = 0; i < hh1.size(); ++i)
        if (hh1[i] != hh2[i])
        {
            std::fprintf(stderr, "exchangeHalos: h differs at %zu\n", i);
            return 4;
        }
    // keys sorted
    auto hk = keys.toHost();
    for (size_t i = 1; i < hk.size(); ++i)
        if (hk[i] < hk[i - 1])
        {
            std::fprintf(stderr, "keys not sorted at %zu\n", i);
            return 5;
        }

    // cstone::findNeighbors call shape on the synced arrays
    const unsigned         ngmax = 400;
    const size_t           first = domain.startIndex(), last = domain.endIndex();
    DeviceVector<unsigned> nb((last - first) * ngmax), cnt(last - first);
    SphxTreeView           tv  = domain.octreeProperties();
    SphxBox                dbx = domain.box();
    must(sphx_find_neighbors(x.data(), y.data(), z.data(), h.data(), first, last, &dbx, &tv, ngmax, nb.data(), cnt.data(),
                             nullptr),
         "sphx_find_neighbors");
    auto hcnt = cnt.toHost();
    auto hidv = id.toHost();
    auto ox = x.toHost(), oy = y.toHost(), oz = z.toHost();
    auto oh = h.toHost();
    std::ofstream out(outfile);
    out.precision(17);
    out << "n " << (last - first) << " local " << domain.nParticlesWithHalos() << " global " << domain.nParticlesGlobal()
        << "\n";
    for (size_t i = first; i < last; ++i)
        out << hidv[i] << " " << hcnt[i - first] << " " << ox[i] << " " << oy[i] << " " << oz[i] << " " << oh[i] << "\n";
    if (comm) sphx_comm_free(comm);
    return 0;
}

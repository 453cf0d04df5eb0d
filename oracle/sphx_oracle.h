/* TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU restatement (oracle/sphx_oracle.cpp) of the reference's SPH-VE hydro step:
 * cstone neighbour search + h-iteration and the six particle loops. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library. The product library (libsphx) never does.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle.py) against
 *   - the reference's own known-answer vectors  sph/test/ve.cpp:112-233 on sph/test/example_data.txt
 *     (tests/golden/ve_kat.npz, all-double instantiation, suffix _d below),
 *   - O(N^2) all-to-all neighbour search as in domain/test/unit/neighbors/findneighbors.cpp:43-133,
 *   - distanceSq PBC known answers domain/test/unit/neighbors/findneighbors.cpp:26-41,
 *   - outputs of the unmodified reference compiled here (oracle/_ref/ref_harness, -ffp-contract=off),
 *     committed as tests/golden/*.npz and regenerated on the fly where oracle/_ref exists.
 *
 * Two type sets: suffix _f = production mixed precision (x,y,z,temp,du double; other fields float),
 *                suffix _d = all double (the reference's unit-test instantiation).
 */
#ifndef SPHX_ORACLE_H
#define SPHX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    typedef struct
    {
        double lim[6];      /* xmin xmax ymin ymax zmin zmax */
        int    boundary[3]; /* 0 open, 1 periodic, 2 fixed (cstone::BoundaryType, sfc/box.hpp:85) */
    } OrcBox;

    typedef struct
    {
        int             numLeafNodes;
        int             numNodes;
        const int*      childOffsets;   /* numNodes */
        const int*      internalToLeaf; /* numNodes */
        const unsigned* layout;         /* numLeafNodes + 1 */
        const double*   centers;        /* numNodes * 3 */
        const double*   sizes;          /* numNodes * 3 */
        float           searchExtFactor;
    } OrcTree;

    typedef struct
    {
        double K, Kcour, Krho, gamma, minDt;
        float  muiConst, alphamin, alphamax, decay_constant, Atmin, Atmax, ramp;
        unsigned ng0, ngmax;
    } OrcParams;

    /* kernel tables + normalisation (sph_kernel_tables.hpp:77-101,144-172; particles_data.hpp:380-387) */
    void orc_tables_f(double sincIndex, float* wh, float* whd, double* K);
    void orc_tables_d(double sincIndex, double* wh, double* whd, double* K);
    double orc_sphynx_3D_k(double n);

    /* findneighbors.hpp:33-60 */
    double orc_distance_sq(int pbc, double x1, double y1, double z1, double x2, double y2, double z2, const OrcBox* box);

    /* cstone::findNeighbors batch (findneighbors.hpp:149-170): counts exclude self, lists ngmax-strided */
    void orc_find_neighbors_f(const double* x, const double* y, const double* z, const float* h, unsigned first,
                              unsigned last, const OrcBox* box, const OrcTree* tree, unsigned ngmax,
                              unsigned* neighbors, unsigned* counts);
    /* O(N^2) search of domain/test/unit/neighbors/all_to_all.hpp:28-57 (float h variant) */
    void orc_all2all_neighbors_f(const double* x, const double* y, const double* z, const float* h, unsigned n,
                                 unsigned* neighbors, unsigned* counts, unsigned ngmax, const OrcBox* box);
    /* sph::findNeighborsSph (sph/find_neighbors.hpp:11-44): h-iteration, nc = 1 + count; returns #non-converged */
    unsigned long orc_find_neighbors_sph_f(const double* x, const double* y, const double* z, float* h, unsigned first,
                                           unsigned last, const OrcBox* box, const OrcTree* tree, unsigned ng0,
                                           unsigned ngmax, unsigned* neighbors, unsigned* nc);
    float orc_update_h_f(unsigned ng0, unsigned nc, float h); /* sph/kernels.hpp:26-32 */

    /* --- the six loops, production types. All arrays are indexed by local particle index (halos included);
     *     neighbours/nc are indexed from `first` as in the reference (neighbors + ngmax*(i-first), nc[i]). --- */
    void orc_xmass_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* box, const unsigned* neighbors,
                     const unsigned* nc, const double* x, const double* y, const double* z, const float* h,
                     const float* m, const float* wh, float* xm);
    void orc_ve_def_gradh_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* box,
                            const unsigned* neighbors, const unsigned* nc, const double* x, const double* y,
                            const double* z, const float* h, const float* m, const float* wh, const float* whd,
                            const float* xm, float* kx, float* gradh);
    void orc_eos_ideal_temp_f(unsigned first, unsigned last, const OrcParams* p, const double* temp, const float* m,
                              const float* kx, const float* xm, const float* gradh, float* prho, float* c);
    void orc_iad_divv_curlv_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* box,
                              const unsigned* neighbors, const unsigned* nc, const double* x, const double* y,
                              const double* z, const float* vx, const float* vy, const float* vz, const float* h,
                              const float* wh, const float* xm, const float* kx, float* c11, float* c12, float* c13,
                              float* c22, float* c23, float* c33, float* divv, float* curlv, double* minDtRho);
    void orc_av_switches_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* box,
                           const unsigned* neighbors, const unsigned* nc, const double* x, const double* y,
                           const double* z, const float* vx, const float* vy, const float* vz, const float* h,
                           const float* c, const float* c11, const float* c12, const float* c13, const float* c22,
                           const float* c23, const float* c33, const float* wh, const float* kx, const float* xm,
                           const float* divv, float* alpha);
    void orc_momentum_energy_f(unsigned first, unsigned last, const OrcParams* p, const OrcBox* box,
                               const unsigned* neighbors, const unsigned* nc, const double* x, const double* y,
                               const double* z, const float* vx, const float* vy, const float* vz, const float* h,
                               const float* m, const float* prho, const float* c, const float* c11, const float* c12,
                               const float* c13, const float* c22, const float* c23, const float* c33, const float* wh,
                               const float* kx, const float* xm, const float* alpha, float* ax, float* ay, float* az,
                               double* du, double* minDtCourant);

    /* --- single-particle J-loops, all-double, the shape of the reference unit tests (sph/test/ve.cpp) --- */
    double orc_xmass_jloop_d(unsigned i, double K, const OrcBox* box, const unsigned* neighbors, unsigned nc,
                             const double* x, const double* y, const double* z, const double* h, const double* m,
                             const double* wh);
    void   orc_ve_def_gradh_jloop_d(unsigned i, double K, const OrcBox* box, const unsigned* neighbors, unsigned nc,
                                    const double* x, const double* y, const double* z, const double* h, const double* m,
                                    const double* wh, const double* whd, const double* xm, double* kx, double* gradh);
    void   orc_iad_jloop_d(unsigned i, double K, const OrcBox* box, const unsigned* neighbors, unsigned nc,
                           const double* x, const double* y, const double* z, const double* h, const double* wh,
                           const double* xm, const double* kx, double* cOut /* 6 */);
    void   orc_divv_curlv_jloop_d(unsigned i, double K, const OrcBox* box, const unsigned* neighbors, unsigned nc,
                                  const double* x, const double* y, const double* z, const double* vx, const double* vy,
                                  const double* vz, const double* h, const double* c11, const double* c12,
                                  const double* c13, const double* c22, const double* c23, const double* c33,
                                  const double* wh, const double* kx, const double* xm, double* out /* divv curlv dV[6] */);
    double orc_av_switches_jloop_d(unsigned i, double K, const OrcBox* box, const unsigned* neighbors, unsigned nc,
                                   const double* x, const double* y, const double* z, const double* vx,
                                   const double* vy, const double* vz, const double* h, const double* c,
                                   const double* c11, const double* c12, const double* c13, const double* c22,
                                   const double* c23, const double* c33, const double* wh, const double* kx,
                                   const double* xm, const double* divv, double dt, double alphamin, double alphamax,
                                   double decay_constant, double alpha_i);
    void   orc_momentum_energy_jloop_d(int avClean, unsigned i, double K, const OrcBox* box, const unsigned* neighbors,
                                       unsigned nc, const double* x, const double* y, const double* z, const double* vx,
                                       const double* vy, const double* vz, const double* h, const double* m,
                                       const double* prho, const double* c, const double* c11, const double* c12,
                                       const double* c13, const double* c22, const double* c23, const double* c33,
                                       double Atmin, double Atmax, double ramp, const double* wh, const double* kx,
                                       const double* xm, const double* alpha, const double* dV11, const double* dV12,
                                       const double* dV13, const double* dV22, const double* dV23, const double* dV33,
                                       double* out /* ax ay az du maxvsignal */);

    /* momentum/energy loop over [first, last) with every operation in fp64 (the all-double instantiation of
     * momentum_energy_kern.hpp:65-222) on the CPU neighbour-list layout; outputs indexed from `first` */
    void orc_momentum_energy_fields_d(unsigned first, unsigned last, unsigned ngmax, double K, const OrcBox* box,
                                      const unsigned* neighbors, const unsigned* nc, const double* x, const double* y,
                                      const double* z, const double* vx, const double* vy, const double* vz,
                                      const double* h, const double* m, const double* prho, const double* c,
                                      const double* c11, const double* c12, const double* c13, const double* c22,
                                      const double* c23, const double* c33, double Atmin, double Atmax, double ramp,
                                      const double* wh, const double* kx, const double* xm, const double* alpha,
                                      double* ax, double* ay, double* az, double* du);

    /* --- turbulence stirring (sph/include/sph/hydro_turb/) ----------------------------------------------------------- */
    /* sph::computeStirring (stirring.hpp:106-125) with stirParticle (:45-83), Tc = T = double, Ta = float */
    void orc_compute_stirring(unsigned first, unsigned last, const double* x, const double* y, const double* z,
                              float* ax, float* ay, float* az, unsigned numModes, const double* modes,
                              const double* phaseReal, const double* phaseImag, const double* amplitudes,
                              double solWeightNorm);
    /* sph::computePhases (phases.hpp:46-72), numDim = 3 */
    void orc_compute_phases(unsigned numModes, const double* ouPhases, double solWeight, const double* modes,
                            double* phasesReal, double* phasesImag);
    /* the deterministic part of sph::updateNoise (driver.hpp:85-98): phases = phases * f + stddev * sqrt(1 - f^2) * z
     * for given unit Gaussians z[] (the reference draws them from std::normal_distribution<double>(0,1)) */
    void orc_update_noise(unsigned n, double* phases, double stddev, double dt, double ts, const double* gaussians);

#ifdef __cplusplus
}
#endif
#endif

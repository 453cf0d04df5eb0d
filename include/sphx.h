/* sphx.h — C ABI of the B200-native SPH-VE hydro step (libsphx.so).
 *
 * Drop-in boundary for ONE path of SPH-EXA (sphexa-org/sphexa @ b458f57): the cornerstone neighbour search over the
 * SFC-sorted octree plus the six SPH-VE particle loops. The reference has no plugin ABI for this path; its seam is the
 * set of `sph::cuda::computeX` explicit template instantiations in the static library `sph_gpu`
 * (sph/include/sph/sph_gpu.hpp:24-56, instantiated at hydro_ve/xmass_gpu.cu:131, ve_def_gradh_gpu.cu:99,
 * iad_divv_curlv_gpu.cu:109, av_switches_gpu.cu:101, momentum_energy_gpu.cu:146-151, eos_gpu.cu:45-75). Every entry point
 * below names the reference function it replaces. The C++20 wrappers with the reference's own signatures
 * (sph::computeXMass(const GroupView&, Dataset&, const Box<T>&) ...) live in sphexa_b200/include/sphx/ and only forward
 * to these functions; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every field pointer is a DEVICE pointer unless the name ends in _host.
 *  - field types are the reference's production type set (sph/include/sph/types.hpp:39-46, particles_data.hpp:220-248):
 *    x,y,z,temp,du double; keys uint64; nc unsigned; everything else float.
 *  - the caller (ParticlesData / Domain in the reference) owns all field and tree memory. The callee reads the pointers
 *    anew on every call, so buffers may be swapped between calls (field aliasing, ve_hydro.hpp:157-189).
 *  - outputs are written for particles [first, last) only; halo values of outputs come from the caller's halo exchange.
 *  - every function returns an int status (SPHX_OK == 0). The reference throws std::runtime_error for the algorithmic
 *    failures (xmass_gpu.cu:127-128) and exits on CUDA errors (cstone/cuda/errorcheck.cuh:14-26); the C++ wrappers turn the
 *    status codes back into those behaviours.
 *  - all work is enqueued on `stream`; functions that return host-visible scalars synchronise that stream.
 *  - threads and devices: a call works on the CURRENT CUDA device of the calling thread; one process may drive several
 *    devices from one thread each (launch parameters are cached per device, nothing device-dependent is process-wide).
 *    Calls for the same workspace / domain object must not overlap.
 *  - there is no CPU fallback: without a CUDA device every compute entry point returns SPHX_ERR_NO_DEVICE.
 */
#ifndef SPHX_H
#define SPHX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define SPHX_ABI_VERSION 1

    enum SphxStatus
    {
        SPHX_OK                 = 0,
        SPHX_ERR_NO_DEVICE      = 1, /* no CUDA device / driver */
        SPHX_ERR_CUDA           = 2, /* a CUDA runtime call failed; sphx_last_error() has the text */
        SPHX_ERR_INVALID        = 3, /* bad argument (null pointer, ng0 > ngmax, ...) */
        SPHX_ERR_WORKSPACE      = 4, /* workspace too small, see sphx_workspace_bytes */
        SPHX_ERR_H_CONVERGENCE  = 5, /* coupled h / neighbour-count iteration did not converge (xmass_gpu.cu:128) */
        SPHX_ERR_NGMAX_OVERFLOW = 6, /* a particle kept more than ngmax neighbours after the h-iteration */
        SPHX_ERR_TRAVERSAL      = 7, /* traversal stack exhausted (xmass_gpu.cu:127) */
        SPHX_ERR_NCCL           = 8,
        SPHX_ERR_TABLE          = 9 /* wh / whd were rewritten under the addresses a loop has already used (see
                                       sphx_invalidate_tables) */
    };

    /* cstone::Box<double> (domain/include/cstone/sfc/box.hpp:94-174). boundary: 0 open, 1 periodic, 2 fixed */
    typedef struct SphxBox
    {
        double lim[6]; /* xmin xmax ymin ymax zmin zmax */
        int    boundary[3];
    } SphxBox;

    /* cstone::OctreeNsView<double, uint64_t> (domain/include/cstone/tree/octree.hpp:279-300); device pointers.
     * Nodes are sorted by (level, SFC key); the 8 children of an internal node are consecutive; childOffsets == 0
     * marks a leaf; centers/sizes are Vec3<double> (centre, half extent). */
    typedef struct SphxTreeView
    {
        int             numLeafNodes;
        int             numNodes;       /* informational, may be 0 (OctreeNsView does not carry it) */
        const uint64_t* prefixes;       /* numNodes, Warren-Salmon placeholder-bit keys (may be NULL: unused here) */
        const int*      childOffsets;   /* numNodes */
        const int*      internalToLeaf; /* numNodes */
        const int*      levelRange;     /* maxTreeLevel + 2 = 23 (may be NULL: unused here) */
        const uint64_t* leaves;         /* numLeafNodes + 1 (may be NULL: unused here) */
        const unsigned* layout;         /* numLeafNodes + 1, index of the first particle of each leaf */
        const double*   centers;        /* numNodes * 3 */
        const double*   sizes;          /* numNodes * 3 */
        float           searchExtFactor;
    } SphxTreeView;

    /* the fields of sphexa::ParticlesData<GpuTag>::devData this path touches (particles_data.hpp:220-248). Unused /
     * inactive fields may be NULL (e.g. dV11..dV33 unless avClean). */
    typedef struct SphxFields
    {
        const double* x;
        const double* y;
        const double* z;
        float*        h;  /* in/out of the neighbour search (h-iteration) */
        const float*  m;
        const float*  vx;
        const float*  vy;
        const float*  vz;
        const double* temp;
        const double* u; /* internal energy; when non-NULL it is used INSTEAD of temp (hydro_ve/eos.hpp:71 `d.u.empty()`, eos_gpu.cu:55) */
        unsigned*     nc; /* out: 1 + neighbour count (find_neighbors.hpp:26,36) */
        float*        xm;
        float*        kx;
        float*        gradh;
        float*        prho;
        float*        c;
        float*        rho; /* optional outputs of the EOS */
        float*        p;
        float*        c11;
        float*        c12;
        float*        c13;
        float*        c22;
        float*        c23;
        float*        c33;
        float*        divv;
        float*        curlv;
        float*        alpha; /* in/out */
        float*        ax;
        float*        ay;
        float*        az;
        double*       du;
        float*        dV11; /* avClean only */
        float*        dV12;
        float*        dV13;
        float*        dV22;
        float*        dV23;
        float*        dV33;
    } SphxFields;

    /* scalar attributes of ParticlesData (particles_data.hpp:88-146) */
    typedef struct SphxParams
    {
        double   K;     /* kernel normalisation */
        double   Kcour; /* Courant factor */
        double   Krho;  /* 1/|divv| time-step factor */
        double   gamma;
        double   minDt; /* current global time step, used by the AV-switch decay */
        double   polytropic_const;
        double   polytropic_index;
        float    muiConst;
        float    soundSpeedConst;
        float    alphamin;
        float    alphamax;
        float    decay_constant;
        float    Atmin;
        float    Atmax;
        float    ramp;
        unsigned ng0;
        unsigned ngmax;
        int      eosChoice; /* sph::EosType: 0 idealGas, 1 isothermal, 2 polytropic (sph/include/sph/eos.hpp:10-15) */
        int      avClean;
    } SphxParams;

    /* cstone::GroupView (domain/include/cstone/traversal/groups.hpp:28-35); device arrays, may be NULL => the callee
     * uses its own fixed groups of 32 SFC-consecutive particles over [first, last) */
    typedef struct SphxGroups
    {
        unsigned        firstBody, lastBody;
        unsigned        numGroups;
        const unsigned* groupStart;
        const unsigned* groupEnd;
    } SphxGroups;

    /* Everything one loop needs. */
    typedef struct SphxStepArgs
    {
        SphxFields   f;
        size_t       numLocal; /* particles incl. halos (array lengths) */
        size_t       first;    /* first assigned particle (Domain::startIndex) */
        size_t       last;     /* one past the last assigned particle (Domain::endIndex) */
        SphxParams   p;
        SphxBox      box;
        SphxTreeView tree;
        const float* wh;  /* 20000-entry kernel table (table_lookup.hpp:10-26), device */
        const float* whd; /* derivative table, device */
        void*        workspace; /* device; holds the neighbour list between calls of one step */
        size_t       workspaceBytes;
        void*        stream; /* cudaStream_t, NULL = default stream */
    } SphxStepArgs;

    /* min/max results of the step that the propagator consumes (ts_global.hpp:72-113) */
    typedef struct SphxStepResult
    {
        double        minDtCourant;
        double        minDtRho;
        unsigned long totalNeighbors; /* sum of nc over [first,last) (conserved_quantities.hpp:146-157) */
        unsigned      maxNc;
        unsigned      numHIterated; /* particles whose h was modified by the iteration */
    } SphxStepResult;

    const char* sphx_last_error(void);
    int         sphx_abi_version(void);
    /* test hook: the loops stage at most maxRecords candidates of a block at a time (blocks with more are processed in
     * candidate chunks; normally the chunk is the shared-memory buffer of the loop, 1024 - 1792 records). 0 = default. */
    void sphx_debug_candidate_chunk(unsigned maxRecords);
    /* 0 if a usable CUDA device is present, SPHX_ERR_NO_DEVICE otherwise */
    int sphx_device_check(void);

    /* bytes of workspace for `numAssigned` = last - first particles.
     * Replaces cstone::allocateNcStacks (traversal/find_neighbors.cuh:492-505). */
    size_t sphx_workspace_bytes(size_t numAssigned, unsigned ngmax);

    /* diagnostics: byte offsets of the workspace sections for (numAssigned, ngmax):
     * out[0] step scalars, [1] block descriptors (40 B each: double origin[3], u32 candBegin, numCand, flags, pad),
     * [2] neighbour list (uint4 vectors of eight 16-bit candidate indices), [3] candidate array (float4 records),
     * [4] total bytes, [5] number of blocks, [6] list vectors per target (ceil(ngmax/8)), [7] candidate capacity */
    void sphx_workspace_layout(size_t numAssigned, unsigned ngmax, size_t out[8]);

    /* host: kernel tables and normalisation constant, sinc^n kernel
     * (ParticlesData::createTables particles_data.hpp:380-387; sph_kernel_tables.hpp:77-101,144-172) */
    int sphx_make_tables_host(double sincIndex, float* wh_host, float* whd_host, double* K);

    /* How the loops evaluate the kernel tables (lt::lookup, sph/include/sph/table_lookup.hpp:13-26). The first loop
     * that sees a (device, wh, whd) address pair copies both tables to the host once (synchronising the stream) and
     * fits one polynomial in v^2 to each. If the polynomials reproduce all 20000 entries to 1e-6 of the table maximum
     * (the sinc^n kernels up to n ~ 7 do; lt::lookup's own fp32 rounding noise is 1e-7), the loops evaluate them
     * instead of the tables: no shared-memory table, no dependent random reads. Otherwise the loops interpolate copies
     * of the tables in shared memory exactly as the reference does. The fit is kept per address pair: a caller that
     * REWRITES a table in place must call sphx_invalidate_tables(); every loop launch spot-checks ~2000 entries and the
     * step fails with SPHX_ERR_TABLE if they no longer match. Setting the environment variable SPHX_FORCE_TABLE
     * selects the shared-memory tables for every table pair first seen afterwards (A/B measurements, tests). */
    void sphx_invalidate_tables(void);
    /* 1: polynomial evaluation, 0: shared-memory tables for this pair; < 0: -(error code). errW / errD (optional):
     * largest deviation of the fit from a table entry, relative to the table maximum */
    int sphx_table_mode(const float* wh, const float* whd, void* stream, double* errW, double* errD);

    /* --- the hot path ---------------------------------------------------------------------------------------------- */

    /* Neighbour search with coupled h-iteration, writes h, nc and the neighbour list (into workspace), then xm.
     * Replaces sph::findNeighborsSfc (sph/find_neighbors.hpp:46-56, a no-op on the reference GPU path) +
     * sph::cuda::computeXMass (hydro_ve/xmass_gpu.cu:104-129). */
    int sphx_find_neighbors_xmass(const SphxStepArgs* a, SphxStepResult* r);
    /* the two halves of the above, for callers that keep the reference's split: sph::findNeighborsSfc
     * (sph/find_neighbors.hpp:46-56: search + h-iteration, writes h, nc and the list) and sph::computeXMass proper
     * (hydro_ve/xmass.hpp:39-74: xm from the stored list) */
    int sphx_find_neighbors_sph(const SphxStepArgs* a, SphxStepResult* r);
    int sphx_xmass(const SphxStepArgs* a);

    /* sph::cuda::computeVeDefGradh (hydro_ve/ve_def_gradh_gpu.cu:50-97): kx, gradh */
    int sphx_ve_def_gradh(const SphxStepArgs* a);
    /* sph::computeEOS (hydro_ve/eos.hpp:192-197; eos_gpu.cu:45-160): prho, c (and rho, p when non-NULL) */
    int sphx_eos(const SphxStepArgs* a);
    /* sph::cuda::computeIadDivvCurlv (hydro_ve/iad_divv_curlv_gpu.cu:51-107) + sph::rhoTimestep (ts_global.hpp:72-95):
     * c11..c33, divv, curlv (dV** when avClean), r->minDtRho */
    int sphx_iad_divv_curlv(const SphxStepArgs* a, SphxStepResult* r);
    /* sph::cuda::computeAVswitches (hydro_ve/av_switches_gpu.cu:48-99): alpha in/out */
    int sphx_av_switches(const SphxStepArgs* a);
    /* sph::cuda::computeMomentumEnergy<avClean> (hydro_ve/momentum_energy_gpu.cu:54-144): ax, ay, az, du, r->minDtCourant.
     * With r != NULL the call also returns the search's sticky error status of this step (SPHX_ERR_TRAVERSAL, ...), so a
     * caller that issues the loops one by one without synchronising learns of it at the end of the step. */
    int sphx_momentum_energy(const SphxStepArgs* a, SphxStepResult* r);

    /* halo exchange callback used by sphx_hydro_step between the loops: exchange the `count` listed device arrays
     * (elemBytes[i] bytes per particle each). Mirrors Domain::exchangeHalos (domain/domain.hpp:372-377). May be NULL. */
    typedef int (*SphxHaloExchangeFn)(void* user, int count, void* const* arrays, const int* elemBytes);

    /* The whole of HydroVeProp::computeForces after domain sync (main/src/propagator/ve_hydro.hpp:147-190):
     * search+xmass | halo{xm} | gradh | eos | halo{vx,vy,vz,prho,c,kx} | iad+divv/curlv | halo{c11..c33,divv} |
     * av switches | halo{alpha} | momentum+energy. The caller provides distinct buffers for gradh/divv/curlv/ay/az or
     * aliases them exactly as the reference does (ay->gradh, {gradh,az}->{divv,curlv}, {divv,curlv}->{ay,az}). */
    int sphx_hydro_step(const SphxStepArgs* a, SphxHaloExchangeFn halo, void* haloUser, SphxStepResult* r);

    /* --- the all-double type set (north_star: "<= 1e-10 fp64") ---------------------------------------------------------- */

    /* The same step with EVERY field and every operation in fp64: the instantiation the reference's own unit tests use
     * (sph/test/ve.cpp: T = double) and its all-double build. This is the precision path (csrc/loops_f64.cu: one thread
     * per target, pairs evaluated from the fp64 coordinates, per-pair PBC fold, four double pow in the Atwood ramp), not
     * the fast one; it shares no pair code with the production path and serves as its on-device yardstick. One rank. */
    typedef struct SphxFieldsF64
    {
        const double* x;
        const double* y;
        const double* z;
        double*       h; /* in/out of the neighbour search */
        const double* m;
        const double* vx;
        const double* vy;
        const double* vz;
        const double* temp;
        const double* u; /* used instead of temp when non-NULL */
        unsigned*     nc;
        double*       xm;
        double*       kx;
        double*       gradh;
        double*       prho;
        double*       c;
        double*       rho; /* optional outputs of the EOS */
        double*       p;
        double*       c11;
        double*       c12;
        double*       c13;
        double*       c22;
        double*       c23;
        double*       c33;
        double*       divv;
        double*       curlv;
        double*       alpha; /* in/out */
        double*       ax;
        double*       ay;
        double*       az;
        double*       du;
        double*       dV11; /* avClean only */
        double*       dV12;
        double*       dV13;
        double*       dV22;
        double*       dV23;
        double*       dV33;
    } SphxFieldsF64;

    typedef struct SphxParamsF64
    {
        double   K, Kcour, Krho, gamma, minDt, polytropic_const, polytropic_index, muiConst, soundSpeedConst, alphamin,
            alphamax, decay_constant, Atmin, Atmax, ramp;
        unsigned ng0, ngmax;
        int      eosChoice, avClean;
    } SphxParamsF64;

    typedef struct SphxStepArgsF64
    {
        SphxFieldsF64 f;
        size_t        numLocal, first, last;
        SphxParamsF64 p;
        SphxBox       box;
        SphxTreeView  tree;
        const double* wh;  /* 20000-entry kernel table in double (sphx_make_tables_host_f64), device */
        const double* whd;
        void*         workspace; /* device, sphx_workspace_bytes_f64(last - first, ngmax): the particle-index neighbour list */
        size_t        workspaceBytes;
        void*         stream;
    } SphxStepArgsF64;

    size_t sphx_workspace_bytes_f64(size_t numAssigned, unsigned ngmax);
    /* createWharmonicTable<double> / createWharmonicDerivativeTable<double> (sph_kernel_tables.hpp:86-101,144-172) */
    int sphx_make_tables_host_f64(double sincIndex, double* wh_host, double* whd_host, double* K);
    /* the loops, one entry each, same meaning as their production counterparts above */
    int sphx_find_neighbors_sph_f64(const SphxStepArgsF64* a, SphxStepResult* r);
    int sphx_xmass_f64(const SphxStepArgsF64* a);
    int sphx_ve_def_gradh_f64(const SphxStepArgsF64* a);
    int sphx_eos_f64(const SphxStepArgsF64* a);
    int sphx_iad_divv_curlv_f64(const SphxStepArgsF64* a, SphxStepResult* r);
    int sphx_av_switches_f64(const SphxStepArgsF64* a);
    int sphx_momentum_energy_f64(const SphxStepArgsF64* a, SphxStepResult* r);
    int sphx_hydro_step_f64(const SphxStepArgsF64* a, SphxStepResult* r);
    int sphx_export_neighbors_f64(const SphxStepArgsF64* a, unsigned* neighbors_dev);

    /* --- cstone call shape ------------------------------------------------------------------------------------------- */

    /* cstone::findNeighbors batch overload (domain/include/cstone/findneighbors.hpp:149-170): for i in [first,last)
     * neighbors[(i-first)*ngmax + k], counts[i-first] (self excluded, count may exceed ngmax, list truncated).
     * No h-iteration. List order within a particle is not the reference CPU's DFS order: compare after sorting
     * (as the reference does for its own GPU search, domain/test/performance/neighbor_driver.cu:255-288). */
    int sphx_find_neighbors(const double* x, const double* y, const double* z, const float* h, size_t first,
                            size_t last, const SphxBox* box, const SphxTreeView* tree, unsigned ngmax,
                            unsigned* neighbors, unsigned* counts, void* stream);

    /* copy the workspace neighbour list of [first,last) into the reference CPU layout neighbors[(i-first)*ngmax + k]
     * (particles_data.hpp:250-251), e.g. for parity checks. */
    int sphx_export_neighbors(const SphxStepArgs* a, unsigned* neighbors_dev);

    /* --- callers of the path ("next" rows) ------------------------------------------------------------------------ */

    /* host tree builder: SFC (Hilbert) keys, sort order, cornerstone leaf tree (bucketSize), linked octree, centres.
     * Replaces for one rank the parts of cstone::Domain::sync that feed OctreeNsView
     * (domain/domain.hpp:181-234,416-428; sfc/hilbert.hpp:43-93; tree/csarray.hpp:181-430; tree/octree.hpp:78-197). */
    typedef struct SphxHostTree SphxHostTree;
    SphxHostTree* sphx_host_tree_build(const double* x_host, const double* y_host, const double* z_host, size_t n,
                                       const SphxBox* box, unsigned bucketSize);
    void          sphx_host_tree_free(SphxHostTree*);
    /* sizes: [0] numNodes, [1] numLeafNodes */
    void sphx_host_tree_sizes(const SphxHostTree*, int* sizes);
    /* copy out (host arrays sized by the caller): order[n] (SFC permutation), keys[n] (sorted), and the tree arrays */
    void sphx_host_tree_get(const SphxHostTree*, unsigned* order, uint64_t* keys, uint64_t* prefixes, int* childOffsets,
                            int* internalToLeaf, int* levelRange /* 23 */, uint64_t* leaves, unsigned* layout,
                            double* centers, double* sizes);
    /* cstone::sfc3D<HilbertKey<uint64_t>> (sfc/sfc.hpp:141-178) for n points */
    void sphx_hilbert_keys_host(const double* x, const double* y, const double* z, size_t n, const SphxBox* box,
                                uint64_t* keys);

    /* --- Domain::sync on the device, one rank --------------------------------------------------- */

    /* Output buffers of sphx_domain_sync, all DEVICE memory owned by the caller. The tree arrays have room for
     * `maxNodes` nodes (leaves/layout: maxNodes + 1); they are exactly the arrays SphxTreeView points at. */
    typedef struct SphxSyncArgs
    {
        size_t        n;          /* particles of this rank */
        SphxBox       box;
        unsigned      bucketSize; /* bucketSizeFocus of cstone::Domain (64 in sphexa.cpp) */
        const double* x;          /* positions in their current (pre-sync) order */
        const double* y;
        const double* z;
        uint64_t*     keys;  /* out [n]: Hilbert keys, sorted */
        unsigned*     order; /* out [n]: SFC permutation, sorted particle i is input particle order[i] */
        int           maxNodes;
        uint64_t*     prefixes;       /* out, Warren-Salmon keys, nodes sorted by (level, key) */
        int*          childOffsets;   /* out */
        int*          internalToLeaf; /* out */
        int*          levelRange;     /* out, 23 entries */
        uint64_t*     leaves;         /* out, numLeafNodes + 1 */
        unsigned*     layout;         /* out, numLeafNodes + 1 */
        double*       centers;        /* out, 3 per node */
        double*       sizes;          /* out, 3 per node */
        void*         scratch;        /* device, sphx_domain_sync_bytes(n, maxNodes) */
        size_t        scratchBytes;
        void*         stream;
        int           flags; /* SPHX_SYNC_* */
    } SphxSyncArgs;

#define SPHX_SYNC_PRESORTED 1 /* x, y, z are in SFC order already: keys are computed, nothing is sorted, order may be NULL */
#define SPHX_SYNC_NO_TREE 2   /* keys and SFC order only; the tree buffers may be NULL */
#define SPHX_SYNC_LIMIT_SHRINK 4 /* a->box is the box of the PREVIOUS sync: open dimensions follow the particles but a side
                                    moves inwards by at most 5 % of the previous extent (limitBoxShrinking, sfc/box.hpp:397-414,
                                    applied by the reference on every sync but the first, domain/assignment.hpp:80-82) */

    size_t sphx_domain_sync_bytes(size_t n, int maxNodes);

    /* The single-rank part of cstone::Domain::sync that feeds the hot path (domain/domain.hpp:181-234,416-428), on the
     * device: Hilbert keys of all particles (sfc/sfc.hpp:141-178, hilbert.hpp:43-93), stable radix sort -> SFC
     * permutation (the reference: sfc/sfcsorter / reorder_gpu), the converged cornerstone octree of bucketSize
     * (tree/csarray.hpp:181-430: every leaf holds <= bucketSize particles, every internal node more), its linked form
     * sorted by (level, key) (tree/octree.hpp:78-197) and the geometric node centres and half sizes
     * (sfc/box.hpp:318-334). When boxOut != NULL the limits of the non-periodic dimensions are first recomputed as the
     * coordinate extrema (makeGlobalBox, sfc/box_mpi.hpp:66-109) and the box that was used is returned there; with
     * boxOut == NULL a->box is used as given. Fields are NOT moved here: apply `order` with sphx_reorder_fields. Returns
     * the node counts (host); synchronises the stream. SPHX_ERR_WORKSPACE if the tree needs more than maxNodes nodes.
     * n == 0 (a rank without particles) is not an error: nothing is sorted and the tree is the empty root leaf. */
    int sphx_domain_sync(const SphxSyncArgs* a, SphxBox* boxOut, int* numNodes, int* numLeafNodes);

    /* counts[c] = number of keys in Hilbert cell c of `level` (8^level cells, cell = key >> 3 (21 - level)) for SFC-sorted
     * keys; the local histogram behind the global assignment (the reference counts the leaves of its global octree,
     * domain/include/cstone/tree/csarray.hpp:181-260 computeNodeCounts). level <= 10. */
    int sphx_cell_histogram(const uint64_t* sortedKeys, size_t n, int level, unsigned* counts, void* stream);

    /* dst[k][i] = src[k][order[i]] for `count` <= 16 arrays of elemBytes[k] in {1,2,4,8} bytes per particle: the field
     * reordering of Domain::sync (domain/domain.hpp:224-230, primitives/gather: GpuSfcSorter::extendMap + gatherGpu).
     * src[k] and dst[k] must not overlap; the caller swaps the buffers afterwards as the reference does. */
    int sphx_reorder_fields(const unsigned* order, size_t n, int count, const void* const* src, void* const* dst,
                            const int* elemBytes, void* stream);

    /* --- integrate() and conserved quantities ----------------------------------------------- */

    /* sph::computeTimestep (sph/include/sph/ts_global.hpp:97-113) without gravity: minDt_new = global min of
     * {minDtCourant, minDtRho, maxDtIncrease * minDt}; in/out: minDt, minDt_m1, ttot. `comm` may be NULL (one rank),
     * otherwise the MPI_Allreduce(MIN) becomes an NCCL all-reduce. */
    struct SphxComm;
    int sphx_compute_timestep(double minDtCourant, double minDtRho, double maxDtIncrease, double* minDt,
                              double* minDt_m1, double* ttot, struct SphxComm* comm, void* stream);

    typedef struct SphxIntegrateArgs
    {
        double*         x; /* in/out */
        double*         y;
        double*         z;
        float*          x_m1; /* in/out: X_n - X_{n-1} */
        float*          y_m1;
        float*          z_m1;
        float*          vx; /* out (read for the fixed-boundary test) */
        float*          vy;
        float*          vz;
        const float*    ax;
        const float*    ay;
        const float*    az;
        double*         temp;  /* in/out, or NULL */
        double*         u;     /* in/out, used when temp == NULL; may be NULL too */
        const double*   du;
        float*          du_m1; /* in/out */
        float*          h;     /* in/out of the smoothing-length update */
        const unsigned* nc;
        size_t          first, last;
        SphxBox         box;
        double          dt;    /* d.minDt after computeTimestep */
        double          dt_m1; /* d.minDt_m1 */
        double          gamma;
        float           muiConst;
        unsigned        ng0;
        void*           stream;
    } SphxIntegrateArgs;

    /* sph::computePositions (sph/include/sph/positions.hpp:177-200, positionUpdate :74-86, energyUpdate :57-63,
     * fixed-boundary skip :97-107, GPU form positions_gpu.cu:119-170) with the CPU path's fp64 arithmetic */
    int sphx_compute_positions(const SphxIntegrateArgs* a);
    /* sph::updateSmoothingLength (sph/include/sph/update_h.hpp:44-53, update_h_gpu.cu:40-49): h = updateH(ng0, nc, h) */
    int sphx_update_smoothing_length(const SphxIntegrateArgs* a);
    /* both of the above in ONE pass over the particles: what HydroVeProp::integrate does after computeTimestep
     * (main/src/propagator/ve_hydro.hpp:206-215) */
    int sphx_integrate(const SphxIntegrateArgs* a);

    /* sphexa::computeConservedQuantities (main/src/observables/conserved_quantities.hpp:49-177, conserved_gpu.cu) */
    typedef struct SphxConserved
    {
        double        ecin, eint, egrav, etot;
        double        linmom, angmom; /* norms of the summed vectors */
        double        linmom3[3], angmom3[3];
        unsigned long totalNeighbors;
    } SphxConserved;
    size_t sphx_conserved_scratch_bytes(void);
    /* temp (or u when temp == NULL) may be NULL => eint = 0; nc may be NULL => totalNeighbors = 0. `scratch`: device,
     * sphx_conserved_scratch_bytes(). Deterministic (fixed reduction tree). With `comm` the ten sums are all-reduced
     * (the reference: MPI_Reduce to rank 0). Synchronises the stream. */
    int sphx_conserved_quantities(const double* x, const double* y, const double* z, const float* vx, const float* vy,
                                  const float* vz, const float* m, const double* temp, const double* u,
                                  const unsigned* nc, size_t first, size_t last, double gamma, float muiConst,
                                  double egrav, void* scratch, struct SphxComm* comm, void* stream, SphxConserved* out);

    /* --- turbulence stirring ---------------------------------------------------------------------- */

    /* the keys of sphexa::TurbulenceConstants() (main/src/init/turbulence_init.hpp:47-72) that
     * sph::TurbulenceData::initModes reads (sph/include/sph/hydro_turb/turbulence_data.hpp:143-177) */
    typedef struct SphxTurbulenceSettings
    {
        double   solWeight;      /* 0.5 */
        double   Lbox;           /* 1.0 */
        double   stEnergyPrefac; /* 5e-3 */
        double   stMachVelocity; /* 0.3 */
        double   epsilon;        /* 1e-15 */
        double   powerLawExp;    /* 5/3 */
        double   anglesExp;      /* 2 */
        uint64_t stMaxModes;     /* 100000 */
        uint64_t rngSeed;        /* 251299 */
        int      stSpectForm;    /* 0 band, 1 parabola, 2 power law */
    } SphxTurbulenceSettings;

    /* sph::TurbulenceData<double, GpuTag> (turbulence_data.hpp:46-180): the stirring modes and amplitudes
     * (createStirringModes, create_modes.hpp:33-227), the Ornstein-Uhlenbeck phases and the std::mt19937 that drives
     * them. Host state plus small device tables; the device tables are created by the first sphx_drive_turbulence. */
    typedef struct SphxTurbulence SphxTurbulence;
    int  sphx_turbulence_create(const SphxTurbulenceSettings* s, SphxTurbulence** out);
    void sphx_turbulence_free(SphxTurbulence* t);
    /* sizes[0] numModes, [1] 1 if every mode is an integer multiple (|i| <= 15) of 2 pi / Lbox (lattice kernel),
     * [2] largest |i|, [3] bytes of the text form of the random engine (incl. the terminating 0) */
    void sphx_turbulence_sizes(const SphxTurbulence* t, size_t sizes[4]);
    /* copy out the host state (any pointer may be NULL): modes[3 numModes], amplitudes[numModes], phases[6 numModes],
     * phasesReal[3 numModes], phasesImag[3 numModes], scalars[4] = variance, decayTime, solWeight, solWeightNorm,
     * rngState = text form of the engine (operator<< of std::mt19937, what TurbulenceData::loadOrStore writes) */
    void sphx_turbulence_get(const SphxTurbulence* t, double* modes, double* amplitudes, double* phases,
                             double* phasesReal, double* phasesImag, double* scalars, char* rngState);
    /* restart (TurbulenceData::loadOrStore, turbulence_data.hpp:82-115): replace the state by stored values. Any pointer
     * may be NULL (that part is kept). modes/amplitudes (3 numModes / numModes values) replace the mode set; phases then
     * must be given too (6 numModes). scalars[4] as in sphx_turbulence_get. rngState: text form of the engine. */
    int sphx_turbulence_restore(SphxTurbulence* t, size_t numModes, const double* modes, const double* amplitudes,
                                const double* phases, const double* scalars, const char* rngState);

    /* host only: one Ornstein-Uhlenbeck step of the phases over minDt (updateNoise, driver.hpp:85-98) and their
     * projection to phasesReal / phasesImag (computePhases, phases.hpp:46-72): the first half of driveTurbulence */
    int sphx_turbulence_advance_host(SphxTurbulence* t, double minDt);

    /* sph::driveTurbulence (sph/include/sph/hydro_turb/driver.hpp:102-128): advance the OU phases by minDt
     * (updateNoise :85-98), project them (computePhases, phases.hpp:46-72), upload, and add the stirring accelerations
     * of particles [first, last) to ax, ay, az (computeStirringGpu, stirring_gpu.cu:43-74; stirParticle,
     * stirring.hpp:45-83). Called right after sphx_hydro_step, as TurbVeProp::computeForces does (turb_ve.hpp:67-72).
     * Enqueued on the stream; does not synchronise. */
    int sphx_drive_turbulence(SphxTurbulence* t, const double* x, const double* y, const double* z, float* ax,
                              float* ay, float* az, size_t first, size_t last, double minDt, void* stream);
    /* the second half only (no OU update): accelerations from the current phasesReal / phasesImag
     * (computeStirringGpu) */
    int sphx_compute_stirring(SphxTurbulence* t, const double* x, const double* y, const double* z, float* ax,
                              float* ay, float* az, size_t first, size_t last, void* stream);

    /* --- SFC domain decomposition over the GPUs of one node (host side) --------------------------------- */

    /* cstone::makeSfcAssignment / uniformBins (domain/include/cstone/domain/domaindecomp.hpp:33-110): contiguous
     * Hilbert-key ranges balanced by particle count, rank boundaries on leaf boundaries of the global octree of
     * bucketSize. sortedKeys: SFC-sorted keys of all n particles; splits[nranks + 1]: first particle of each rank. */
    int sphx_sfc_assignment_host(const uint64_t* sortedKeys, size_t n, int nranks, unsigned bucketSize,
                                 size_t* splits);

    /* Halo discovery for the rank owning the SFC-sorted particles [ownedBegin, ownedEnd)
     * (cstone Halos::discover, domain/include/cstone/halos/halos.hpp:131-192): whole-cell halos, i.e. every particle of
     * a leaf (octree of bucketSize over all particles) that overlaps the box of an owned leaf inflated by
     * 2 max(h in that leaf), PBC-aware. flags[n] must be zero-initialised; halo particles are set to 1. */
    int sphx_find_halos_host(const double* x_host, const double* y_host, const double* z_host, const float* h_host,
                             size_t n, const SphxBox* box, unsigned bucketSize, size_t ownedBegin, size_t ownedEnd,
                             unsigned char* flags);

    /* --- dynamic decomposition: the host half of the multi-rank Domain::sync ------------------- */

    /* From the GLOBAL particle count per Hilbert cell of `level` (sphx_cell_histogram + all-reduce), identically on
     * every rank: balanced contiguous cell ranges (makeSfcAssignment / uniformBins, domaindecomp.hpp:33-110), the halo
     * cells of `rank` = non-empty foreign cells adjacent (26-neighbourhood, periodic wrap per periodic[3]) to its own
     * non-empty cells (Halos::discover, halos/halos.hpp:131-192: whole-cell halos; complete if the cell edge is
     * >= 2 max h), the local layout [halos | assigned | halos] (layout.hpp:150-163) and the SphxHaloPlan arrays: peers,
     * send lists as local particle indices (valid once the assigned particles sit SFC-sorted at [nHaloLeft,
     * nHaloLeft + nAssigned)), one contiguous receive range per peer. NULL on bad arguments. */
    typedef struct SphxCellPlan SphxCellPlan;
    SphxCellPlan* sphx_cell_plan_build_host(const unsigned* globalCounts_host, int level, const int* periodic, int rank,
                                            int nranks);
    /* the same with a per-cell reach: rings_host[c] rings of cells around cell c hold every neighbour of its particles
     * (ceil(2 max h of the cell / cell edge)); cell c is a halo cell of every rank that owns a cell within whose reach
     * it lies. NULL = one ring everywhere. */
    SphxCellPlan* sphx_cell_plan_build_host_rings(const unsigned* globalCounts_host, const unsigned char* rings_host,
                                                  int level, const int* periodic, int rank, int nranks);
    void          sphx_cell_plan_free(SphxCellPlan*);
    /* sizes: [0] numPeers [1] send indices [2] halo cells [3] nAssigned [4] nHaloLeft [5] nHaloRight [6] nGlobal
     * [7] nranks */
    void sphx_cell_plan_sizes(const SphxCellPlan*, size_t sizes[8]);
    /* copy out (any pointer may be NULL): cellSplits[nranks + 1], peers[numPeers], sendOffsets[numPeers + 1],
     * sendIdx[...], recvBegin[numPeers], recvCount[numPeers], recvCells[...] (sorted halo cell ids) */
    void sphx_cell_plan_get(const SphxCellPlan*, uint64_t* cellSplits, int* peers, unsigned* sendOffsets,
                            unsigned* sendIdx, unsigned* recvBegin, unsigned* recvCount, unsigned* recvCells);

    /* The same plan built on the device (csrc/domain_sync.cu), so that the global histogram never leaves the GPU and
     * the only host traffic of a sync is this POD: one thread per cell decodes its Hilbert index, encodes its 26
     * neighbours and flags halo / send cells; scans and compactions turn the flags into the receive ranges and the
     * send index lists. Bit-identical to sphx_cell_plan_build_host (tests/test_gpu_sync_integrate.py). */
#define SPHX_MAX_RANKS 64
    typedef struct SphxCellPlanSummary
    {
        uint64_t cellSplits[SPHX_MAX_RANKS + 1];   /* rank r owns cells [cellSplits[r], cellSplits[r + 1]) */
        uint64_t sendOffLocal[SPHX_MAX_RANKS + 1]; /* my SFC-sorted particles [.[r], .[r+1]) migrate to rank r */
        uint64_t nGlobal, nAssigned, nHaloLeft, nHaloRight;
        uint32_t recvCount[SPHX_MAX_RANKS]; /* halo particles received from rank r (one contiguous range each) */
        uint32_t sendCount[SPHX_MAX_RANKS]; /* halo particles sent to rank r: sendIdx holds them rank after rank */
        uint32_t numRecvCells, numSend;
        uint32_t overflow; /* 1: sendCapacity too small, sendIdx is incomplete */
        uint32_t pad;
    } SphxCellPlanSummary;
    size_t sphx_cell_plan_device_bytes(int level);
    /* globalCounts, localCounts: device arrays of 8^level counts (all ranks / this rank's particles before the
     * migration). rings: device array of 8^level bytes or NULL: the reach of a cell in rings of cells (Chebyshev distance),
     * ceil(2 max h of the cell / cell edge), identical on all ranks (all-reduced); NULL = one ring everywhere. maxRing
     * bounds rings[] (1..16). sendIdx: device, sendCapacity entries. recvCells: device, 8^level entries, may be NULL
     * (sorted halo cell ids, for inspection). out: HOST pointer; the call synchronises the stream. */
    int sphx_cell_plan_build_device(const unsigned* globalCounts, const unsigned* localCounts,
                                    const unsigned char* rings, int maxRing, int level, const int* periodic, int rank,
                                    int nranks, void* scratch, size_t scratchBytes, unsigned* sendIdx,
                                    size_t sendCapacity, unsigned* recvCells, SphxCellPlanSummary* out, void* stream);

    /* --- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch ------------------------------------------------- */

#define SPHX_UNIQUE_ID_BYTES 128
    typedef struct SphxComm SphxComm;

    /* NCCL bootstrap. Rank 0 calls sphx_comm_unique_id and distributes the 128 bytes through any host channel (the
     * reference would use MPI_Bcast; bench.py and the tests use the torch.distributed store); then every rank calls
     * sphx_comm_init with the CUDA device it will run on already current. NCCL is dlopen()ed at this point. */
    int sphx_comm_unique_id(char id[SPHX_UNIQUE_ID_BYTES]);
    int sphx_comm_init(SphxComm** comm, int rank, int nranks, const char id[SPHX_UNIQUE_ID_BYTES]);
    int sphx_comm_free(SphxComm* comm);

    /* Halo exchange plan of one rank = cstone's SendList + recv layout (domain/include/cstone/domain/layout.hpp,
     * halos/halos.hpp:234-254) flattened: what this rank sends to each peer is a list of local particle indices
     * (device array); what it receives from a peer is ONE contiguous range of its local arrays, because local arrays
     * are SFC-sorted and the peer owns a contiguous key range. */
    typedef struct SphxHaloPlan
    {
        int             numPeers;
        const int*      peers;       /* host, numPeers: peer ranks */
        const unsigned* sendOffsets; /* host, numPeers + 1: slice of sendIdx that goes to each peer */
        const unsigned* sendIdx;     /* DEVICE, sendOffsets[numPeers] local particle indices */
        const unsigned* recvBegin;   /* host, numPeers: first local index of the halo range filled by each peer */
        const unsigned* recvCount;   /* host, numPeers */
        void*           sendBuffer;  /* DEVICE scratch for the packed sends */
        size_t          sendBufferBytes; /* >= sum over the arrays of one exchange of 16-aligned(numSend * elemBytes) */
    } SphxHaloPlan;

    /* Domain::exchangeHalos (domain/include/cstone/domain/domain.hpp:372-377; GPU path halos/exchange_halos_gpu.cuh:
     * 34-119): pack the listed arrays' send lists with one gather kernel, then grouped ncclSend / ncclRecv; receives
     * land directly in the halo ranges. Everything is enqueued on `stream`; no host synchronisation. count <= 8,
     * elemBytes 4 or 8. */
    int sphx_halo_exchange(SphxComm* comm, const SphxHaloPlan* plan, int count, void* const* arrays,
                           const int* elemBytes, void* stream);

    /* all-reduce of n <= 16 host doubles (op: 0 min, 1 max, 2 sum); replaces the MPI_Allreduce of the time step
     * (sph/include/sph/ts_global.hpp:97-113) and of the neighbour statistics. Synchronises the stream. */
    int sphx_allreduce_f64(SphxComm* comm, double* values_host, int n, int op, void* stream);

    /* in-place all-reduce of a DEVICE array (dtype: 0 u32, 1 u64, 2 f32, 3 f64; op: 0 min, 1 max, 2 sum), enqueued on
     * the stream: the global cell histogram of the dynamic decomposition (the reference: MPI_Allreduce of the global
     * tree's node counts, domain/include/cstone/tree/update_mpi.hpp) */
    int sphx_allreduce_device(SphxComm* comm, void* data_dev, size_t n, int dtype, int op, void* stream);

    /* Particle migration of Domain::sync (domain/include/cstone/domain/exchange_keys / domaindecomp_mpi.hpp
     * exchangeParticles): every rank holds its particles SFC-sorted, so what goes to rank r is ONE slice
     * [sendOffsets[r], sendOffsets[r+1]) of every field; it lands at [recvOffsets[r], recvOffsets[r+1]) of dst on rank r's
     * side (offsets in particles, host arrays of nranks + 1 entries; the own slice is copied device to device). One
     * grouped ncclSend/ncclRecv round per array, everything enqueued on the stream. */
    int sphx_exchange_slices(SphxComm* comm, const size_t* sendOffsets, const size_t* recvOffsets, int count,
                             const void* const* src, void* const* dst, const int* elemBytes, void* stream);

    /* rank and size of a communicator */
    int sphx_comm_rank(const SphxComm* comm, int* rank, int* nranks);

    /* --- multi-rank Domain::sync behind one entry point ------------------------------------ */

    /* cstone::Domain (domain/include/cstone/domain/domain.hpp:60-160) as far as the hot path needs it: the global box,
     * the bucket size of the local octree, the communicator, and the state one sync leaves for the next calls (halo plan,
     * octree arrays, scratch). comm == NULL: one rank. */
    typedef struct SphxDomain SphxDomain;
    int  sphx_domain_create(SphxDomain** out, SphxComm* comm, const SphxBox* box, unsigned bucketSize);
    void sphx_domain_destroy(SphxDomain* d);

    typedef struct SphxDomainSyncArgs
    {
        int          count;     /* 4 <= count <= 16 particle arrays: [0..2] x, y, z (double), [3] h (float), then the
                                   conserved fields of the propagator (4 or 8 bytes per particle) */
        void* const* arrays;    /* DEVICE arrays of `capacity` elements; this rank's particles sit at [inFirst, inLast) */
        void* const* spare;     /* DEVICE arrays of the same shapes: the sync works array -> spare -> array -> spare,
                                   like the reference swaps its fields with scratch vectors (domain.hpp:224-230) */
        const int*   elemBytes;
        size_t       capacity;
        size_t       inFirst, inLast;
        int          numHaloFields; /* <= 8 */
        const int*   haloFields;    /* indices into arrays: fields whose halo values the sync delivers (the reference:
                                       x, y, z, h always, plus m in ve_hydro.hpp:133-139) */
        void*        stream;
    } SphxDomainSyncArgs;

    typedef struct SphxDomainResult
    {
        size_t          first, last; /* Domain::startIndex / endIndex: the assigned particles in the new layout */
        size_t          numLocal;    /* Domain::nParticlesWithHalos: [halos | assigned | halos], SFC-sorted */
        size_t          numGlobal;
        size_t          needCapacity; /* set also on SPHX_ERR_WORKSPACE: elements per array the sync needs */
        SphxBox         box;          /* Domain::box(): open dimensions follow the particles */
        SphxTreeView    tree;         /* Domain::octreeProperties(): arrays owned by the domain, valid until the next sync */
        const uint64_t* localKeys;    /* Hilbert keys of the numLocal particles (device, owned by the domain) */
        int             level;        /* cell level of the decomposition plan */
        int             swapped;      /* 1: the new local set is in `spare`, the caller swaps its pointers */
    } SphxDomainResult;

    /* cstone::Domain::sync (domain/domain.hpp:181-234) for N ranks, re-designed around the global per-cell particle
     * histogram (csrc/domain_dist.cu): box update, keys + radix sort, histogram + ncclAllReduce, decomposition plan on the
     * device, particle migration (one slice per peer and field), merge, halo exchange of `haloFields`, octree over the
     * local particles. Collective over the communicator; every rank must call it. On SPHX_ERR_WORKSPACE nothing has moved:
     * grow the arrays to res->needCapacity (keeping [inFirst, inLast)) and call again (every rank: the error is NOT agreed
     * upon across ranks, size the arrays with head room). Synchronises the stream. */
    int sphx_domain_sync_dist(SphxDomain* d, const SphxDomainSyncArgs* a, SphxDomainResult* res);

    /* Domain::exchangeHalos (domain/domain.hpp:372-377) with the plan of the last sync */
    int sphx_domain_exchange_halos(SphxDomain* d, int count, void* const* arrays, const int* elemBytes, void* stream);
    /* the halo plan of the last sync, for sphx_hydro_step_dist */
    const SphxHaloPlan* sphx_domain_halo_plan(const SphxDomain* d);
    /* copy the Hilbert keys of the numLocal particles of the last sync into a caller buffer (device) */
    int sphx_domain_copy_local_keys(const SphxDomain* d, uint64_t* dst_dev, void* stream);

    /* sphx_hydro_step with the four halo exchanges of HydroVeProp::computeForces (ve_hydro.hpp:154,165,174,185) done
     * by sphx_halo_exchange and the result scalars reduced over all ranks (min dt, sum of neighbours, max nc). */
    int sphx_hydro_step_dist(const SphxStepArgs* a, SphxComm* comm, const SphxHaloPlan* plan, SphxStepResult* r);

    /* the reductions at the end of a distributed step on their own, for callers that issue the loops one by one:
     * MIN of minDtCourant / minDtRho (ts_global.hpp:97-113), SUM of totalNeighbors, MAX of maxNc, and the agreement on
     * the status: every rank passes its local status; if any rank failed, every rank gets an error back. One grouped
     * NCCL all-reduce; synchronises the stream. */
    int sphx_reduce_step_result(SphxComm* comm, int localStatus, SphxStepResult* inout, void* stream);

    /* sph::updateH (sph/include/sph/kernels.hpp:26-32), T = float: host evaluation of exactly the arithmetic the
     * device uses (bit-exact emulation of glibc powf, csrc/sphx_powf.h), exported for verification against libm. */
    float sphx_update_h_host(unsigned ng0, unsigned nc, float h);
    float sphx_powf_host(float x, float y);

#ifdef __cplusplus
}
#endif
#endif /* SPHX_H */

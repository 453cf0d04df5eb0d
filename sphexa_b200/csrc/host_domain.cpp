/*! @file
 * Host side of the path's caller ("next" rows of SURVEY §8f): kernel tables, Hilbert keys, SFC order and the
 * cornerstone octree view the neighbour search consumes — for one rank.
 *
 * Replaces (reference paths relative to /root/reference):
 *   ParticlesData::createTables         sph/include/sph/particles_data.hpp:380-387, sph_kernel_tables.hpp:77-101,144-172
 *   cstone::sfc3D / iHilbert            domain/include/cstone/sfc/sfc.hpp:141-178, sfc/hilbert.hpp:43-93
 *   computeOctree (converged)           domain/include/cstone/tree/csarray.hpp:181-430
 *   buildOctreeCpu / nodeFpCenters      domain/include/cstone/tree/octree.hpp:78-197, focus/source_center.hpp:130-142
 *
 * Not a port: the Hilbert curve is a table-driven state machine (orientation state x octant -> digit, next state)
 * generated at start-up, and the octree is built top-down, one level per pass over the sorted keys, which directly
 * yields the (level, key)-sorted node layout the reference obtains by sorting Warren-Salmon prefixes. For a converged
 * tree (every internal node holds more than bucketSize particles, every leaf at most bucketSize) both constructions
 * give the same arrays, which tests/test_host_tree.py checks against reference dumps.
 */
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "sphx.h"
#include "sphx_powf.h"

namespace
{

constexpr int      kMaxLevel = 21; // maxTreeLevel<uint64_t>, domain/include/cstone/sfc/common.hpp
constexpr unsigned kMaxCoord = 1u << kMaxLevel;

/* ------------------------------------------------ Hilbert curve ------------------------------------------------ */

struct Orientation
{
    uint8_t perm[3]; // current axis a reads original axis perm[a]
    uint8_t flip;    // bit (2 - a) set: current axis a is complemented
    bool    operator==(const Orientation& o) const { return std::memcmp(this, &o, sizeof(Orientation)) == 0; }
};

struct HilbertTables
{
    // indexed [state][octant of ORIGINAL coordinate bits x<<2|y<<1|z]
    std::vector<std::array<uint8_t, 8>> digit, next;
    // inverse: [state][digit] -> original octant
    std::vector<std::array<uint8_t, 8>> octant;

    HilbertTables()
    {
        // Per-octant rule of the curve (octant given in CURRENT orientation): Hilbert digit, which axes get
        // complemented and how axes are permuted for the next finer level.
        // digit: Morton octant -> Hilbert digit {0,1,3,2,7,6,4,5}
        const uint8_t m2h[8] = {0, 1, 3, 2, 7, 6, 4, 5};
        struct Rule
        {
            uint8_t flip;    // xyz bits
            uint8_t perm[3]; // new axis a takes old axis perm[a]
        } rule[8];
        for (int o = 0; o < 8; ++o)
        {
            int xi = (o >> 2) & 1, yi = (o >> 1) & 1, zi = o & 1;
            int fx = xi & ((!yi) | zi);
            int fy = (xi & (yi | zi)) | (yi & (!zi));
            int fz = (xi & (!yi) & (!zi)) | (yi & (!zi));
            rule[o].flip = uint8_t(fx << 2 | fy << 1 | fz);
            if (zi) { rule[o].perm[0] = 1, rule[o].perm[1] = 2, rule[o].perm[2] = 0; } // cyclic rotation
            else if (!yi) { rule[o].perm[0] = 2, rule[o].perm[1] = 1, rule[o].perm[2] = 0; } // swap x and z
            else { rule[o].perm[0] = 0, rule[o].perm[1] = 1, rule[o].perm[2] = 2; }
        }

        std::vector<Orientation> states{Orientation{{0, 1, 2}, 0}};
        for (size_t s = 0; s < states.size(); ++s)
        {
            digit.emplace_back();
            next.emplace_back();
            octant.emplace_back();
            for (int b = 0; b < 8; ++b)
            {
                Orientation st = states[s];
                // current bits = permuted original bits, complemented per axis
                int cur = 0;
                for (int a = 0; a < 3; ++a)
                {
                    int bit = (b >> (2 - st.perm[a])) & 1;
                    bit ^= (st.flip >> (2 - a)) & 1;
                    cur |= bit << (2 - a);
                }
                const Rule& r = rule[cur];
                // new orientation: first complement (flip ^ rule.flip), then permute axes
                Orientation nx;
                uint8_t     f = st.flip ^ r.flip;
                nx.flip       = 0;
                for (int a = 0; a < 3; ++a)
                {
                    nx.perm[a] = st.perm[r.perm[a]];
                    nx.flip |= ((f >> (2 - r.perm[a])) & 1) << (2 - a);
                }
                auto   it  = std::find(states.begin(), states.end(), nx);
                size_t idx = it - states.begin();
                if (it == states.end()) { states.push_back(nx); }
                digit[s][b]           = m2h[cur];
                next[s][b]            = uint8_t(idx);
                octant[s][m2h[cur]]   = uint8_t(b);
            }
        }
    }
};

const HilbertTables& tables()
{
    static HilbertTables t;
    return t;
}

uint64_t hilbertEncode(unsigned ix, unsigned iy, unsigned iz)
{
    const auto& t     = tables();
    uint64_t    key   = 0;
    unsigned    state = 0;
    for (int level = kMaxLevel - 1; level >= 0; --level)
    {
        unsigned b = ((ix >> level) & 1u) << 2 | ((iy >> level) & 1u) << 1 | ((iz >> level) & 1u);
        key        = (key << 3) | t.digit[state][b];
        state      = t.next[state][b];
    }
    return key;
}

void hilbertDecode(uint64_t key, unsigned& ix, unsigned& iy, unsigned& iz)
{
    const auto& t     = tables();
    unsigned    state = 0;
    ix = iy = iz = 0;
    for (int level = kMaxLevel - 1; level >= 0; --level)
    {
        unsigned d = unsigned(key >> (3 * level)) & 7u;
        unsigned b = t.octant[state][d];
        ix |= ((b >> 2) & 1u) << level;
        iy |= ((b >> 1) & 1u) << level;
        iz |= (b & 1u) << level;
        state = t.next[state][b];
    }
}

struct HBox
{
    double xmin, ymin, zmin, lx, ly, lz, ilx, ily, ilz;
    explicit HBox(const SphxBox& b)
    {
        xmin = b.lim[0], ymin = b.lim[2], zmin = b.lim[4];
        lx = b.lim[1] - b.lim[0], ly = b.lim[3] - b.lim[2], lz = b.lim[5] - b.lim[4];
        ilx = 1.0 / (b.lim[1] - b.lim[0]), ily = 1.0 / (b.lim[3] - b.lim[2]), ilz = 1.0 / (b.lim[5] - b.lim[4]);
    }
};

//! normalised integer coordinate as in sfc3D (sfc/sfc.hpp:141-159): floor(x * m) - xmin * m, clipped to 2^21 - 1
inline unsigned gridCoord(double v, double vmin, double m)
{
    int i = int(std::floor(v * m) - vmin * m);
    return unsigned(std::min(i, int(kMaxCoord - 1)));
}

void computeKeys(const double* x, const double* y, const double* z, size_t n, const SphxBox& sb, uint64_t* keys)
{
    HBox   box(sb);
    double mx = kMaxCoord * box.ilx, my = kMaxCoord * box.ily, mz = kMaxCoord * box.ilz;
    tables();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i)
    {
        keys[i] = hilbertEncode(gridCoord(x[i], box.xmin, mx), gridCoord(y[i], box.ymin, my),
                                gridCoord(z[i], box.zmin, mz));
    }
}

} // namespace

namespace sphx
{
//! the Hilbert state machine as flat [state * 8 + octant] arrays for the device (domain_sync.cu); returns #states
int hilbertTablesFlat(uint8_t* digit, uint8_t* next, uint8_t* octant, int maxStates)
{
    const auto& t  = tables();
    int         ns = int(t.digit.size());
    if (ns > maxStates) return -1;
    for (int s = 0; s < ns; ++s)
        for (int b = 0; b < 8; ++b)
        {
            digit[s * 8 + b]  = t.digit[s][b];
            next[s * 8 + b]   = t.next[s][b];
            octant[s * 8 + b] = t.octant[s][b];
        }
    return ns;
}
} // namespace sphx

struct SphxHostTree
{
    std::vector<unsigned> order;
    std::vector<uint64_t> keys;
    std::vector<uint64_t> prefixes;
    std::vector<int>      childOffsets, internalToLeaf;
    std::vector<int>      levelRange;
    std::vector<uint64_t> leaves;
    std::vector<unsigned> layout;
    std::vector<double>   centers, sizes;
    int                   numNodes{0}, numLeaves{0};
};

namespace
{
template<class T>
int makeTables(double sincIndex, T* wh, T* whd, double* K)
{
    if (!wh || !whd || !K) return SPHX_ERR_INVALID;
    const double halfPi = 1.57079632679489661923; // M_PI_2
    const double pi     = 3.14159265358979323846;
    auto sinc = [=](double v)
    {
        if (v == 0.0) return 1.0;
        double pv = halfPi * v;
        return std::sin(pv) / pv;
    };
    auto kernel = [=](double v) { return std::pow(sinc(v), sincIndex); };
    auto kernelDerivative = [=](double v)
    {
        if (v == 0.0) return sincIndex * std::pow(sinc(v), sincIndex - 1) * 0.0;
        double pv = halfPi * v;
        double sv = std::sin(pv) / (pv);
        double ds = sv * halfPi * ((std::cos(pv) / std::sin(pv)) - 1.0 / pv);
        return sincIndex * std::pow(sinc(v), sincIndex - 1) * ds;
    };

    // normalisation: 1 / integral of 4 pi x^2 W(x) over [0, 2], composite Simpson with 2000 intervals where odd and
    // even interior samples are sorted before summation (sph_kernel_tables.hpp:22-56,77-84)
    {
        const uint64_t      nInt = 2000;
        const double        step = 2.0 / double(nInt);
        auto                vol  = [&](double xx) { return 4.0 * pi * xx * xx * kernel(xx); };
        std::vector<double> odd, even;
        for (uint64_t i = 1; i < nInt; ++i)
            ((i & 1) ? odd : even).push_back(vol(0.0 + double(i) * step));
        std::sort(odd.begin(), odd.end());
        std::sort(even.begin(), even.end());
        double so = std::accumulate(odd.begin(), odd.end(), 0.0), se = std::accumulate(even.begin(), even.end(), 0.0);
        *K        = 1.0 / (step / 3.0 * (vol(0.0) + vol(2.0) + 4.0 * so + 2.0 * se));
    }

    // tabulation: the abscissa is stepped and rounded in T, the functor evaluates in double
    // (sph_kernel_tables.hpp:86-101)
    const T dx = T((2.0 - 0.0) / 19999);
    for (size_t i = 0; i < 20000; ++i)
    {
        T v    = T(0.0 + double(T(i) * dx));
        wh[i]  = T(kernel(double(v)));
        whd[i] = T(kernelDerivative(double(v)));
    }
    return SPHX_OK;
}
} // namespace

extern "C"
{

void sphx_hilbert_keys_host(const double* x, const double* y, const double* z, size_t n, const SphxBox* box,
                            uint64_t* keys)
{
    computeKeys(x, y, z, n, *box, keys);
}

SphxHostTree* sphx_host_tree_build(const double* x, const double* y, const double* z, size_t n, const SphxBox* sbox,
                                   unsigned bucketSize)
{
    auto* t = new SphxHostTree;
    std::vector<uint64_t> rawKeys(n);
    computeKeys(x, y, z, n, *sbox, rawKeys.data());

    t->order.resize(n);
    std::iota(t->order.begin(), t->order.end(), 0u);
    std::stable_sort(t->order.begin(), t->order.end(), [&](unsigned a, unsigned b) { return rawKeys[a] < rawKeys[b]; });
    t->keys.resize(n);
    for (size_t i = 0; i < n; ++i)
        t->keys[i] = rawKeys[t->order[i]];
    const std::vector<uint64_t>& keys = t->keys;

    // top-down, level by level. A node is the key range [start, start + 8^(21-level)); it is split while it holds
    // more than bucketSize particles (the fixed point of the reference's rebalance, csarray.hpp:181-260).
    struct Node
    {
        uint64_t start;
        unsigned pBegin, pEnd;
    };
    std::vector<Node>     level{Node{0, 0u, unsigned(n)}};
    std::vector<Node>     nodes;
    std::vector<int>      nodeLevel;
    t->levelRange.assign(kMaxLevel + 2, 0);

    for (int l = 0; l <= kMaxLevel; ++l)
    {
        t->levelRange[l]  = int(nodes.size());
        size_t levelBegin = nodes.size();
        nodes.insert(nodes.end(), level.begin(), level.end());
        nodeLevel.insert(nodeLevel.end(), level.size(), l);
        t->childOffsets.resize(nodes.size(), 0);

        std::vector<Node> nextLevel;
        if (l < kMaxLevel)
        {
            uint64_t childRange = uint64_t(1) << (3 * (kMaxLevel - l - 1));
            for (size_t k = 0; k < level.size(); ++k)
            {
                const Node& nd = level[k];
                if (nd.pEnd - nd.pBegin <= bucketSize) continue;
                // children index: first node of the next level is at nodes.size() + nextLevel.size()
                t->childOffsets[levelBegin + k] = int(nodes.size() + nextLevel.size());
                unsigned p                      = nd.pBegin;
                for (int c = 0; c < 8; ++c)
                {
                    uint64_t cs  = nd.start + uint64_t(c) * childRange;
                    uint64_t ce  = cs + childRange; // may wrap to 2^63 for the last node: compare with care
                    unsigned end = (c == 7) ? nd.pEnd
                                            : unsigned(std::lower_bound(keys.begin() + p, keys.begin() + nd.pEnd, ce) -
                                                       keys.begin());
                    nextLevel.push_back(Node{cs, p, end});
                    p = end;
                }
            }
        }
        level.swap(nextLevel);
        if (level.empty())
        {
            for (int ll = l + 1; ll <= kMaxLevel + 1; ++ll)
                t->levelRange[ll] = int(nodes.size());
            break;
        }
    }
    t->levelRange[kMaxLevel + 1] = int(nodes.size());
    t->numNodes                  = int(nodes.size());

    // leaves in SFC order
    std::vector<int> leafNodes;
    for (int i = 0; i < t->numNodes; ++i)
        if (t->childOffsets[i] == 0) leafNodes.push_back(i);
    std::sort(leafNodes.begin(), leafNodes.end(), [&](int a, int b) { return nodes[a].start < nodes[b].start; });
    t->numLeaves = int(leafNodes.size());

    t->internalToLeaf.assign(t->numNodes, -1);
    t->leaves.resize(t->numLeaves + 1);
    t->layout.resize(t->numLeaves + 1);
    for (int k = 0; k < t->numLeaves; ++k)
    {
        int nd                 = leafNodes[k];
        t->internalToLeaf[nd]  = k;
        t->leaves[k]           = nodes[nd].start;
        t->layout[k]           = nodes[nd].pBegin;
    }
    t->leaves[t->numLeaves] = uint64_t(1) << (3 * kMaxLevel);
    t->layout[t->numLeaves] = unsigned(n);

    // Warren-Salmon prefixes and geometric centres / half-sizes (sfc/box.hpp:318-334 centerAndSize)
    t->prefixes.resize(t->numNodes);
    t->centers.resize(size_t(t->numNodes) * 3);
    t->sizes.resize(size_t(t->numNodes) * 3);
    HBox         box(*sbox);
    const double uL = 1.0 / kMaxCoord;
    const double hx = 0.5 * uL * box.lx, hy = 0.5 * uL * box.ly, hz = 0.5 * uL * box.lz;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < t->numNodes; ++i)
    {
        int l          = nodeLevel[i];
        t->prefixes[i] = (uint64_t(1) << (3 * l)) | (nodes[i].start >> (3 * (kMaxLevel - l)));
        unsigned ix, iy, iz;
        hilbertDecode(nodes[i].start, ix, iy, iz);
        unsigned cube = kMaxCoord >> l;
        unsigned mask = ~(cube - 1);
        ix &= mask, iy &= mask, iz &= mask;
        int ixmax = int(ix + cube), iymax = int(iy + cube), izmax = int(iz + cube);
        t->centers[3 * i + 0] = box.xmin + (ixmax + int(ix)) * hx;
        t->centers[3 * i + 1] = box.ymin + (iymax + int(iy)) * hy;
        t->centers[3 * i + 2] = box.zmin + (izmax + int(iz)) * hz;
        t->sizes[3 * i + 0]   = (ixmax - int(ix)) * hx;
        t->sizes[3 * i + 1]   = (iymax - int(iy)) * hy;
        t->sizes[3 * i + 2]   = (izmax - int(iz)) * hz;
    }
    return t;
}

void sphx_host_tree_free(SphxHostTree* t) { delete t; }

void sphx_host_tree_sizes(const SphxHostTree* t, int* sizes)
{
    sizes[0] = t->numNodes;
    sizes[1] = t->numLeaves;
}

void sphx_host_tree_get(const SphxHostTree* t, unsigned* order, uint64_t* keys, uint64_t* prefixes, int* childOffsets,
                        int* internalToLeaf, int* levelRange, uint64_t* leaves, unsigned* layout, double* centers,
                        double* sizes)
{
    auto cp = [](auto* dst, const auto& v)
    {
        if (dst) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    cp(order, t->order), cp(keys, t->keys), cp(prefixes, t->prefixes), cp(childOffsets, t->childOffsets);
    cp(internalToLeaf, t->internalToLeaf), cp(levelRange, t->levelRange), cp(leaves, t->leaves), cp(layout, t->layout);
    cp(centers, t->centers), cp(sizes, t->sizes);
}

/* ------------------------------------ SFC domain decomposition (one node, N ranks) ------------------------------------ */

/*! cstone::makeSfcAssignment / uniformBins (domain/include/cstone/domain/domaindecomp.hpp:33-110): contiguous Hilbert-key
 *  ranges balanced by particle count; rank boundaries are boundaries of the leaves of the global octree of `bucketSize`.
 *  Input: the SFC-sorted keys of ALL particles. Output: splits[r] = index of the first particle of rank r. */
int sphx_sfc_assignment_host(const uint64_t* sortedKeys, size_t n, int nranks, unsigned bucketSize, size_t* splits)
{
    if (!sortedKeys || !splits || nranks < 1) return SPHX_ERR_INVALID;
    // leaf boundaries of the converged bucketSize tree, as particle offsets: split key ranges top-down
    std::vector<size_t> bounds; // particle offset of every leaf start, ascending
    struct Range
    {
        uint64_t start;
        int      level;
        size_t   pBegin, pEnd;
    };
    std::vector<Range> stack{Range{0, 0, 0, n}};
    while (!stack.empty())
    {
        Range r = stack.back();
        stack.pop_back();
        if (r.pEnd - r.pBegin <= bucketSize || r.level == kMaxLevel)
        {
            bounds.push_back(r.pBegin);
            continue;
        }
        uint64_t childRange = uint64_t(1) << (3 * (kMaxLevel - r.level - 1));
        size_t   cuts[9];
        cuts[0] = r.pBegin, cuts[8] = r.pEnd;
        for (int c = 1; c < 8; ++c)
            cuts[c] = size_t(std::lower_bound(sortedKeys + cuts[c - 1], sortedKeys + r.pEnd,
                                              r.start + uint64_t(c) * childRange) - sortedKeys);
        for (int c = 7; c >= 0; --c) // push in reverse: leaves pop in SFC order
            stack.push_back(Range{r.start + uint64_t(c) * childRange, r.level + 1, cuts[c], cuts[c + 1]});
    }
    bounds.push_back(n);
    // uniformBins: cut at the leaf boundary closest to r * n / nranks
    splits[0] = 0;
    for (int r = 1; r < nranks; ++r)
    {
        size_t target = size_t((double(n) * r) / nranks);
        auto   it     = std::lower_bound(bounds.begin(), bounds.end(), target);
        size_t hiB = *it, loB = (it == bounds.begin()) ? 0 : *(it - 1);
        size_t cut = (hiB - target < target - loB) ? hiB : loB;
        splits[r]  = std::max(cut, splits[r - 1]);
    }
    splits[nranks] = n;
    return SPHX_OK;
}

/*! Halo discovery for the rank that owns the SFC-sorted particles [ownedBegin, ownedEnd)
 *  (cstone Halos::discover, domain/include/cstone/halos/halos.hpp:131-192 + traversal/collisions.hpp): every particle of a
 *  leaf (octree of bucketSize over all particles) that overlaps the box of an owned leaf inflated by 2 max(h in that
 *  leaf) is a halo: whole-cell halos, PBC-aware. flags[n]: set to 1 for halo particles (must be zero-initialised). */
int sphx_find_halos_host(const double* x, const double* y, const double* z, const float* h, size_t n,
                         const SphxBox* sbox, unsigned bucketSize, size_t ownedBegin, size_t ownedEnd,
                         unsigned char* flags)
{
    if (!x || !y || !z || !h || !sbox || !flags || ownedEnd < ownedBegin || ownedEnd > n) return SPHX_ERR_INVALID;
    SphxHostTree* t = sphx_host_tree_build(x, y, z, n, sbox, bucketSize);
    for (size_t i = 0; i < n; ++i)
        if (t->order[i] != i)
        {
            sphx_host_tree_free(t);
            return SPHX_ERR_INVALID; // particles must be SFC-sorted
        }
    HBox       box(*sbox);
    const bool pbc[3] = {sbox->boundary[0] == 1, sbox->boundary[1] == 1, sbox->boundary[2] == 1};
    const double L[3] = {box.lx, box.ly, box.lz};

    // owned leaves: [firstLeaf, lastLeaf)
    int firstLeaf = int(std::upper_bound(t->layout.begin(), t->layout.end(), unsigned(ownedBegin)) - t->layout.begin()) - 1;
    int lastLeaf  = int(std::lower_bound(t->layout.begin(), t->layout.end(), unsigned(ownedEnd)) - t->layout.begin());
    std::vector<int> leafToNode(t->numLeaves);
    for (int i = 0; i < t->numNodes; ++i)
        if (t->childOffsets[i] == 0) leafToNode[t->internalToLeaf[i]] = i;

#pragma omp parallel for schedule(dynamic, 64)
    for (int lf = firstLeaf; lf < lastLeaf; ++lf)
    {
        size_t pb = std::max<size_t>(t->layout[lf], ownedBegin), pe = std::min<size_t>(t->layout[lf + 1], ownedEnd);
        if (pb >= pe) continue;
        float hmax = 0;
        for (size_t p = pb; p < pe; ++p)
            hmax = std::max(hmax, h[p]);
        int    nd = leafToNode[lf];
        double c[3], sz[3];
        for (int d = 0; d < 3; ++d)
        {
            c[d]  = t->centers[3 * nd + d];
            sz[d] = t->sizes[3 * nd + d] + 2.0 * double(hmax) * (1.0 + 1e-6);
        }
        int stack[512];
        int sp      = 0;
        stack[sp++] = 0;
        while (sp > 0)
        {
            int  node    = stack[--sp];
            bool overlap = true;
            for (int d = 0; d < 3 && overlap; ++d)
            {
                double dd = t->centers[3 * node + d] - c[d];
                if (pbc[d]) dd -= L[d] * std::rint(dd / L[d]);
                overlap = std::fabs(dd) <= (t->sizes[3 * node + d] + sz[d]) * (1.0 + 1e-9);
            }
            if (!overlap) continue;
            int child = t->childOffsets[node];
            if (child == 0)
            {
                int    k  = t->internalToLeaf[node];
                size_t qb = t->layout[k], qe = t->layout[k + 1];
                for (size_t q = qb; q < qe; ++q)
                    if (q < ownedBegin || q >= ownedEnd) flags[q] = 1;
            }
            else if (sp + 8 <= 512)
            {
                for (int cc = 0; cc < 8; ++cc)
                    stack[sp++] = child + cc;
            }
        }
    }
    sphx_host_tree_free(t);
    return SPHX_OK;
}

/* ------------------------------------ cell-based decomposition plan (dynamic, per sync) ------------------------------------ */

/*! The host half of the multi-rank Domain::sync. Ranks agree on a GLOBAL histogram of particles
 *  per Hilbert cell of one level (device histogram + NCCL all-reduce); from it every rank derives, without further
 *  communication and identically:
 *   - the assignment: contiguous cell ranges balanced by particle count (makeSfcAssignment / uniformBins,
 *     domaindecomp.hpp:33-110, with the cells in the role of the global-tree leaves),
 *   - its halo cells: every non-empty foreign cell adjacent (26-neighbourhood, periodic wrap) to one of its own
 *     non-empty cells. The caller picks the level such that a cell edge is >= 2 max(h): then every neighbour of an
 *     assigned particle lies in an owned or halo cell (Halos::discover, halos/halos.hpp:131-192, whole-cell halos),
 *   - what it sends: its own cells that are halo cells of a peer (adjacency is symmetric, so no request messages),
 *   - the local layout [halos | assigned | halos] in SFC order (layout.hpp:150-163) with one contiguous receive range
 *     per peer, and the send lists as local particle indices. */
struct SphxCellPlan
{
    int                   level, rank, nranks;
    std::vector<uint64_t> cellSplits; // nranks + 1
    std::vector<int>      peers;
    std::vector<unsigned> sendOffsets, sendIdx, recvBegin, recvCount;
    std::vector<unsigned> recvCells; // sorted
    size_t                nAssigned{0}, nHaloLeft{0}, nHaloRight{0}, nGlobal{0};
};

SphxCellPlan* sphx_cell_plan_build_host_rings(const unsigned* globalCounts, const unsigned char* rings, int level,
                                              const int* periodic, int rank, int nranks)
{
    if (!globalCounts || level < 0 || level > 10 || !periodic || rank < 0 || rank >= nranks) return nullptr;
    auto*          p      = new SphxCellPlan;
    const uint64_t ncell  = uint64_t(1) << (3 * level);
    const int      side   = 1 << level;
    const int      cshift = kMaxLevel - level;
    p->level = level, p->rank = rank, p->nranks = nranks;

    std::vector<uint64_t> prefix(ncell + 1);
    prefix[0] = 0;
    for (uint64_t c = 0; c < ncell; ++c)
        prefix[c + 1] = prefix[c] + globalCounts[c];
    const uint64_t N = prefix[ncell];
    p->nGlobal       = N;

    // uniformBins: rank r starts at the first cell boundary at or after r N / R particles
    p->cellSplits.assign(nranks + 1, 0);
    p->cellSplits[nranks] = ncell;
    for (int r = 1; r < nranks; ++r)
    {
        uint64_t target = uint64_t((__uint128_t(N) * r) / nranks);
        uint64_t c      = std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin();
        p->cellSplits[r] = std::max(std::min(c, ncell), p->cellSplits[r - 1]);
    }
    const uint64_t cb = p->cellSplits[rank], ce = p->cellSplits[rank + 1];
    p->nAssigned = prefix[ce] - prefix[cb];

    auto ownerOf = [&](uint64_t c)
    { return int(std::upper_bound(p->cellSplits.begin() + 1, p->cellSplits.end() - 1, c) - (p->cellSplits.begin() + 1)); };

    // reach of a cell in rings of cells (Chebyshev distance): 1 unless the caller passes per-cell values
    int maxRing = 1;
    if (rings)
    {
        for (uint64_t c = 0; c < ncell; ++c)
            if (globalCounts[c]) maxRing = std::max(maxRing, int(rings[c]));
    }
    auto ringOf = [&](uint64_t c) { return rings ? std::max(1, int(rings[c])) : 1; };

    // adjacency sweep over my non-empty cells: cell c needs every non-empty foreign cell within ringOf(c) rings (halo
    // cells); a foreign cell c2 needs c if c lies within ringOf(c2) rings of it (send cells). Both follow from the
    // global arrays, so no rank has to ask another one.
    std::vector<std::vector<unsigned>> sendCells(nranks);
    std::vector<unsigned>              recvCells;
    tables();
#pragma omp parallel
    {
        std::vector<std::vector<unsigned>> mySend(nranks);
        std::vector<unsigned>              myRecv;
        std::vector<int>                   sentTo;
#pragma omp for schedule(static)
        for (int64_t c = int64_t(cb); c < int64_t(ce); ++c)
        {
            if (globalCounts[c] == 0) continue;
            unsigned ix, iy, iz;
            hilbertDecode(uint64_t(c) << (3 * cshift), ix, iy, iz);
            int cx = int(ix >> cshift), cy = int(iy >> cshift), cz = int(iz >> cshift);
            const int Rc = ringOf(uint64_t(c));
            sentTo.clear();
            auto wrap = [&](int v, int per, bool& ok)
            {
                if (v >= 0 && v < side) return v;
                if (!per)
                {
                    ok = false;
                    return 0;
                }
                return ((v % side) + side) % side;
            };
            for (int dz = -maxRing; dz <= maxRing; ++dz)
                for (int dy = -maxRing; dy <= maxRing; ++dy)
                    for (int dx = -maxRing; dx <= maxRing; ++dx)
                    {
                        if (!dx && !dy && !dz) continue;
                        const int d  = std::max(std::abs(dx), std::max(std::abs(dy), std::abs(dz)));
                        bool      ok = true;
                        int       nx = wrap(cx + dx, periodic[0], ok), ny = wrap(cy + dy, periodic[1], ok),
                            nz = wrap(cz + dz, periodic[2], ok);
                        if (!ok) continue;
                        uint64_t c2 = hilbertEncode(unsigned(nx) << cshift, unsigned(ny) << cshift,
                                                    unsigned(nz) << cshift) >> (3 * cshift);
                        if (c2 >= cb && c2 < ce) continue;
                        if (globalCounts[c2] == 0) continue;
                        if (d <= Rc) myRecv.push_back(unsigned(c2));
                        if (d <= ringOf(c2))
                        {
                            int o = ownerOf(c2);
                            if (std::find(sentTo.begin(), sentTo.end(), o) == sentTo.end())
                            {
                                sentTo.push_back(o);
                                mySend[o].push_back(unsigned(c));
                            }
                        }
                    }
        }
#pragma omp critical
        {
            recvCells.insert(recvCells.end(), myRecv.begin(), myRecv.end());
            for (int r = 0; r < nranks; ++r)
                sendCells[r].insert(sendCells[r].end(), mySend[r].begin(), mySend[r].end());
        }
    }
    std::sort(recvCells.begin(), recvCells.end());
    recvCells.erase(std::unique(recvCells.begin(), recvCells.end()), recvCells.end());
    for (auto& v : sendCells)
        std::sort(v.begin(), v.end()); // each of my cells appears at most once per peer

    for (unsigned c : recvCells)
        (c < cb ? p->nHaloLeft : p->nHaloRight) += globalCounts[c];

    // peers in rank order; receive ranges: recvCells are sorted, a peer's cells are consecutive in that order
    p->sendOffsets.push_back(0);
    size_t ri = 0, pos = 0;
    for (int r = 0; r < nranks; ++r)
    {
        if (r == rank)
        {
            pos = p->nHaloLeft + p->nAssigned; // right halos follow the assigned range
            continue;
        }
        size_t   begin = pos, cnt = 0;
        uint64_t rb = p->cellSplits[r], re = p->cellSplits[r + 1];
        while (ri < recvCells.size() && recvCells[ri] >= rb && recvCells[ri] < re)
        {
            cnt += globalCounts[recvCells[ri]];
            ++ri;
        }
        pos += cnt;
        if (cnt == 0 && sendCells[r].empty()) continue;
        p->peers.push_back(r);
        p->recvBegin.push_back(unsigned(begin));
        p->recvCount.push_back(unsigned(cnt));
        for (unsigned c : sendCells[r])
        {
            unsigned first = unsigned(p->nHaloLeft + (prefix[c] - prefix[cb]));
            for (unsigned k = 0; k < globalCounts[c]; ++k)
                p->sendIdx.push_back(first + k);
        }
        p->sendOffsets.push_back(unsigned(p->sendIdx.size()));
    }
    p->recvCells.swap(recvCells);
    return p;
}

SphxCellPlan* sphx_cell_plan_build_host(const unsigned* globalCounts, int level, const int* periodic, int rank,
                                        int nranks)
{
    return sphx_cell_plan_build_host_rings(globalCounts, nullptr, level, periodic, rank, nranks);
}

void sphx_cell_plan_free(SphxCellPlan* p) { delete p; }

void sphx_cell_plan_sizes(const SphxCellPlan* p, size_t sizes[8])
{
    sizes[0] = p->peers.size(), sizes[1] = p->sendIdx.size(), sizes[2] = p->recvCells.size();
    sizes[3] = p->nAssigned, sizes[4] = p->nHaloLeft, sizes[5] = p->nHaloRight, sizes[6] = p->nGlobal;
    sizes[7] = size_t(p->nranks);
}

void sphx_cell_plan_get(const SphxCellPlan* p, uint64_t* cellSplits, int* peers, unsigned* sendOffsets,
                        unsigned* sendIdx, unsigned* recvBegin, unsigned* recvCount, unsigned* recvCells)
{
    auto cp = [](auto* dst, const auto& v)
    {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    cp(cellSplits, p->cellSplits), cp(peers, p->peers), cp(sendOffsets, p->sendOffsets), cp(sendIdx, p->sendIdx);
    cp(recvBegin, p->recvBegin), cp(recvCount, p->recvCount), cp(recvCells, p->recvCells);
}

float sphx_update_h_host(unsigned ng0, unsigned nc, float h) { return sphx::updateHExact(ng0, nc, h); }
float sphx_powf_host(float x, float y) { return sphx::glibcPowf(x, y); }

/* -------------------------------------------- kernel tables -------------------------------------------- */


int sphx_make_tables_host(double sincIndex, float* wh, float* whd, double* K) { return makeTables(sincIndex, wh, whd, K); }
int sphx_make_tables_host_f64(double sincIndex, double* wh, double* whd, double* K)
{
    return makeTables(sincIndex, wh, whd, K);
}

} // extern "C"

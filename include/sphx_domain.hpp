/*! @file
 * C++20 facade over the C ABI of libsphx with the call shape of cstone::Domain
 * (/root/reference/domain/include/cstone/domain/domain.hpp):
 *
 *   reference                                                   here
 *   ---------------------------------------------------------   ------------------------------------------------------
 *   Domain(rank, nRanks, bucketSize, bucketSizeFocus, theta,    sphx::Domain(rank, nRanks, bucketSize, bucketSizeFocus,
 *          box)                                    :81-99              theta, box, comm)
 *   sync(keys, x, y, z, h, std::tie(properties...),             sync(keys, x, y, z, h, std::tie(properties...),
 *        std::tie(scratch...))                     :181-234          std::tie(scratch...))
 *   exchangeHalos(std::tie(arrays...), sendBuf, recvBuf) :372   exchangeHalos(std::tie(arrays...), sendBuf, recvBuf)
 *   startIndex() endIndex() nParticles()                        the same
 *   nParticlesWithHalos() box() octreeProperties()  :379-428    the same (octreeProperties returns the SphxTreeView that
 *                                                               SphxStepArgs::tree takes)
 *
 * Header only; everything is forwarded to sphx_domain_sync_dist / sphx_domain_exchange_halos (include/sphx.h,
 * csrc/domain_dist.cu). Errors become std::runtime_error, as cstone reports them (or MPI_Abort there).
 *
 * The particle arrays are device vectors: any type with data(), size(), capacity(), resize(n) and swap(other) works
 * (cstone::DeviceVector and thrust::device_vector do); sphx::DeviceVector below is a minimal one for callers that have
 * neither. As in the reference, sync() leaves every array resized to nParticlesWithHalos() in the layout
 * [halos | assigned | halos], SFC-sorted, and uses the scratch vectors as the other half of its ping-pong: one scratch
 * vector of the same type per array is taken from the scratch tuple where the tuple offers one; for the rest the domain
 * keeps spares of its own.
 */
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <typeinfo>
#include <utility>
#include <vector>

#include "sphx.h"

namespace sphx
{

using LocalIndex = unsigned;

//! minimal RAII device array (uninitialised growth, contents kept on resize)
template<class T>
class DeviceVector
{
public:
    using value_type = T;
    DeviceVector() = default;
    explicit DeviceVector(std::size_t n) { resize(n); }
    DeviceVector(const std::vector<T>& host)
    {
        resize(host.size());
        if (!host.empty()) check(cudaMemcpy(d_, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    DeviceVector(const DeviceVector&)            = delete;
    DeviceVector& operator=(const DeviceVector&) = delete;
    DeviceVector(DeviceVector&& o) noexcept { swap(o); }
    DeviceVector& operator=(DeviceVector&& o) noexcept
    {
        swap(o);
        return *this;
    }
    ~DeviceVector()
    {
        if (d_) cudaFree(d_);
    }

    T*          data() { return d_; }
    const T*    data() const { return d_; }
    std::size_t size() const { return size_; }
    std::size_t capacity() const { return cap_; }
    bool        empty() const { return size_ == 0; }

    void reserve(std::size_t n)
    {
        if (n <= cap_) return;
        T* p = nullptr;
        check(cudaMalloc(&p, n * sizeof(T)));
        if (size_) check(cudaMemcpy(p, d_, size_ * sizeof(T), cudaMemcpyDeviceToDevice));
        if (d_) cudaFree(d_);
        d_ = p, cap_ = n;
    }
    void resize(std::size_t n)
    {
        reserve(n);
        size_ = n;
    }
    void swap(DeviceVector& o) noexcept
    {
        std::swap(d_, o.d_), std::swap(size_, o.size_), std::swap(cap_, o.cap_);
    }
    std::vector<T> toHost() const
    {
        std::vector<T> h(size_);
        if (size_) check(cudaMemcpy(h.data(), d_, size_ * sizeof(T), cudaMemcpyDeviceToHost));
        return h;
    }

private:
    static void check(cudaError_t e)
    {
        if (e != cudaSuccess) throw std::runtime_error(std::string("sphx::DeviceVector: ") + cudaGetErrorString(e));
    }
    T*          d_    = nullptr;
    std::size_t size_ = 0, cap_ = 0;
};

template<class T>
T* rawPtr(DeviceVector<T>& v)
{
    return v.data();
}

namespace detail
{
inline void check(int rc, const char* what)
{
    if (rc != SPHX_OK) throw std::runtime_error(std::string(what) + ": " + sphx_last_error() + "\n");
}

//! type-erased handle on one caller array for the duration of a sync
struct ArrayRef
{
    void*       self;
    int         elemBytes;
    void*       (*data)(void*);
    std::size_t (*size)(void*);
    std::size_t (*capacity)(void*);
    void        (*resize)(void*, std::size_t);
    void        (*reserve)(void*, std::size_t);
    void        (*swapWith)(void*, void*); // swap with another vector of the same type
    std::size_t typeHash;
};

template<class V>
ArrayRef makeRef(V& v)
{
    using U = typename V::value_type;
    ArrayRef r;
    r.self      = &v;
    r.elemBytes = int(sizeof(U));
    r.data      = [](void* s) -> void* { return static_cast<V*>(s)->data(); };
    r.size      = [](void* s) { return static_cast<V*>(s)->size(); };
    r.capacity  = [](void* s) { return static_cast<V*>(s)->capacity(); };
    r.resize    = [](void* s, std::size_t n) { static_cast<V*>(s)->resize(n); };
    r.reserve   = [](void* s, std::size_t n)
    {
        // grow keeping the contents, whatever the vector type offers
        V& vec = *static_cast<V*>(s);
        if (n > vec.capacity())
        {
            std::size_t keep = vec.size();
            vec.resize(n);
            vec.resize(keep);
        }
    };
    r.swapWith = [](void* a, void* b) { static_cast<V*>(a)->swap(*static_cast<V*>(b)); };
    r.typeHash = typeid(V).hash_code();
    return r;
}
} // namespace detail

/*! @brief cstone::Domain for the hot path: SFC domain decomposition over the GPUs of a node, halo exchange, local octree
 *
 * @tparam KeyType  SFC key type; libsphx computes 64-bit Hilbert keys (cstone::HilbertKey<uint64_t>)
 * @tparam T        coordinate type; libsphx implements the production type set (double)
 */
template<class KeyType = uint64_t, class T = double>
class Domain
{
    static_assert(std::is_same_v<KeyType, uint64_t> && std::is_same_v<T, double>,
                  "libsphx implements the production type set: 64-bit Hilbert keys, double coordinates");

public:
    using RealType = T;

    /*! @param comm  NCCL communicator of libsphx (sphx_comm_init) or nullptr for one rank; the reference takes
     *               MPI_COMM_WORLD implicitly. bucketSize (global tree) and theta (gravity) have no counterpart on the
     *               hot path: the decomposition works on the global cell histogram, the local octree uses
     *               bucketSizeFocus as the reference's focus tree does. */
    Domain(int rank, int nRanks, unsigned bucketSize, unsigned bucketSizeFocus, float /*theta*/, const SphxBox& box,
           SphxComm* comm = nullptr)
        : rank_(rank)
        , nRanks_(nRanks)
        , box_(box)
    {
        if (bucketSize < bucketSizeFocus)
        {
            throw std::runtime_error("The bucket size of the global tree must not be smaller than the bucket size"
                                     " of the focused tree\n");
        }
        if ((nRanks > 1) != (comm != nullptr)) throw std::runtime_error("sphx::Domain: nRanks > 1 needs a communicator\n");
        detail::check(sphx_domain_create(&dom_, comm, &box, bucketSizeFocus), "sphx_domain_create");
    }
    Domain(const Domain&)            = delete;
    Domain& operator=(const Domain&) = delete;
    ~Domain()
    {
        sphx_domain_destroy(dom_);
        for (auto& s : ownSpares_)
            if (s.first) cudaFree(s.first);
    }

    /*! @brief cstone::Domain::sync (domain.hpp:181-234)
     *
     * In: the particles of this rank in x, y, z, h and the properties, at [startIndex(), endIndex()) of arrays of equal
     * size (on the first call: the whole arrays), in any order. Out: every array (and particleKeys) resized to
     * nParticlesWithHalos(), layout [halos | assigned | halos] in SFC order; x, y, z, h and the FIRST property (the
     * mass in ve_hydro.hpp:133-139) carry valid halo values, the other properties only for the assigned particles (the
     * propagator exchanges what else it needs, as in the reference).
     */
    template<class KeyVec, class VectorX, class VectorH, class... Vectors1, class... Vectors2>
    void sync(KeyVec& particleKeys, VectorX& x, VectorX& y, VectorX& z, VectorH& h,
              std::tuple<Vectors1&...> particleProperties, std::tuple<Vectors2&...> scratchBuffers)
    {
        static_assert(sizeof(typename VectorX::value_type) == 8 && sizeof(typename VectorH::value_type) == 4,
                      "x, y, z are double, h is float (sph/include/sph/types.hpp:39-46)");
        std::vector<detail::ArrayRef> arr{detail::makeRef(x), detail::makeRef(y), detail::makeRef(z), detail::makeRef(h)};
        std::apply([&](auto&... p) { (arr.push_back(detail::makeRef(p)), ...); }, particleProperties);
        std::vector<detail::ArrayRef> scratch;
        std::apply([&](auto&... s) { (scratch.push_back(detail::makeRef(s)), ...); }, scratchBuffers);
        const int count = int(arr.size());
        if (count > 16) throw std::runtime_error("sphx::Domain::sync: at most 12 particle properties\n");

        std::size_t n = x.size();
        for (auto& a : arr)
            if (a.size(a.self) != n) throw std::runtime_error("sphx::Domain::sync: array sizes differ\n");
        const std::size_t inFirst = firstCall_ ? 0 : first_, inLast = firstCall_ ? n : last_;

        // the other half of the ping-pong: a scratch vector of the same type per array where the caller offers one
        std::vector<int> partner(count, -1);
        std::vector<bool> used(scratch.size(), false);
        for (int k = 0; k < count; ++k)
            for (std::size_t q = 0; q < scratch.size(); ++q)
                if (!used[q] && scratch[q].typeHash == arr[k].typeHash)
                {
                    partner[k] = int(q), used[q] = true;
                    break;
                }
        ownSpares_.resize(std::max<std::size_t>(ownSpares_.size(), count), {nullptr, 0});

        std::size_t capacity = minCapacity(arr);
        if (capacity < n + n / 4 + 1024) capacity = grow(arr, n + n / 4 + 1024);
        SphxDomainResult res{};
        for (;;)
        {
            std::vector<void*> ptr(count), spare(count);
            std::vector<int>   eb(count);
            for (int k = 0; k < count; ++k)
            {
                eb[k]  = arr[k].elemBytes;
                ptr[k] = arr[k].data(arr[k].self);
                if (partner[k] >= 0)
                {
                    auto& s = scratch[partner[k]];
                    s.reserve(s.self, capacity);
                    spare[k] = s.data(s.self);
                }
                else
                {
                    auto& own = ownSpares_[k];
                    if (own.second < capacity * std::size_t(eb[k]))
                    {
                        if (own.first) cudaFree(own.first);
                        if (cudaMalloc(&own.first, capacity * std::size_t(eb[k])) != cudaSuccess)
                            throw std::runtime_error("sphx::Domain::sync: cudaMalloc of a spare array failed\n");
                        own.second = capacity * std::size_t(eb[k]);
                    }
                    spare[k] = own.first;
                }
            }
            const int          halo[5] = {0, 1, 2, 3, 4};
            SphxDomainSyncArgs a{};
            a.count = count, a.arrays = ptr.data(), a.spare = spare.data(), a.elemBytes = eb.data();
            a.capacity = capacity, a.inFirst = inFirst, a.inLast = inLast;
            a.numHaloFields = count > 4 ? 5 : 4, a.haloFields = halo, a.stream = nullptr;
            int rc = sphx_domain_sync_dist(dom_, &a, &res);
            if (rc == SPHX_ERR_WORKSPACE && res.needCapacity > capacity)
            {
                capacity = grow(arr, res.needCapacity + res.needCapacity / 4 + 1024);
                continue;
            }
            detail::check(rc, "sphx_domain_sync_dist");
            break;
        }
        // the new local set is in the spares: swap storage where the spare is a caller vector, copy back otherwise
        for (int k = 0; k < count; ++k)
        {
            if (partner[k] >= 0)
            {
                auto& s = scratch[partner[k]];
                s.resize(s.self, res.numLocal);
                arr[k].swapWith(arr[k].self, s.self);
            }
            else
            {
                arr[k].resize(arr[k].self, res.numLocal);
                if (cudaMemcpy(arr[k].data(arr[k].self), ownSpares_[k].first, res.numLocal * std::size_t(arr[k].elemBytes),
                               cudaMemcpyDeviceToDevice) != cudaSuccess)
                    throw std::runtime_error("sphx::Domain::sync: copy from the spare array failed\n");
            }
            arr[k].resize(arr[k].self, res.numLocal);
        }
        particleKeys.resize(res.numLocal);
        detail::check(sphx_domain_copy_local_keys(dom_, reinterpret_cast<uint64_t*>(particleKeys.data()), nullptr),
                      "sphx_domain_copy_local_keys");
        cudaDeviceSynchronize();
        first_ = LocalIndex(res.first), last_ = LocalIndex(res.last), numLocal_ = LocalIndex(res.numLocal);
        numGlobal_ = res.numGlobal;
        box_ = res.box, tree_ = res.tree;
        firstCall_ = false;
    }

    //! cstone::Domain::exchangeHalos (domain.hpp:372-377); the buffers are the reference's MPI staging areas: unused here
    template<class... Vectors, class SendBuffer, class ReceiveBuffer>
    void exchangeHalos(std::tuple<Vectors&...> arrays, SendBuffer& /*sendBuffer*/, ReceiveBuffer& /*receiveBuffer*/) const
    {
        std::vector<void*> ptr;
        std::vector<int>   eb;
        std::apply(
            [&](auto&... v)
            {
                (ptr.push_back(v.data()), ...);
                (eb.push_back(int(sizeof(typename std::decay_t<decltype(v)>::value_type))), ...);
            },
            arrays);
        for (std::size_t b = 0; b < ptr.size(); b += 8)
        {
            int c = int(std::min<std::size_t>(8, ptr.size() - b));
            detail::check(sphx_domain_exchange_halos(dom_, c, ptr.data() + b, eb.data() + b, nullptr),
                          "sphx_domain_exchange_halos");
        }
    }

    LocalIndex       startIndex() const { return first_; }
    LocalIndex       endIndex() const { return last_; }
    LocalIndex       nParticles() const { return last_ - first_; }
    LocalIndex       nParticlesWithHalos() const { return numLocal_; }
    std::size_t      nParticlesGlobal() const { return numGlobal_; }
    const SphxBox&   box() const { return box_; }
    //! cstone::Domain::octreeProperties (domain.hpp:416-428): what SphxStepArgs::tree and sphx_find_neighbors take
    SphxTreeView     octreeProperties() const { return tree_; }
    const SphxHaloPlan* haloPlan() const { return sphx_domain_halo_plan(dom_); }
    SphxDomain*      handle() const { return dom_; }

private:
    static std::size_t minCapacity(const std::vector<detail::ArrayRef>& arr)
    {
        std::size_t c = ~std::size_t(0);
        for (auto& a : arr)
            c = std::min(c, a.capacity(a.self));
        return c;
    }
    static std::size_t grow(std::vector<detail::ArrayRef>& arr, std::size_t capacity)
    {
        for (auto& a : arr)
            a.reserve(a.self, capacity);
        return minCapacity(arr);
    }

    int          rank_, nRanks_;
    SphxBox      box_;
    SphxDomain*  dom_ = nullptr;
    SphxTreeView tree_{};
    LocalIndex   first_ = 0, last_ = 0, numLocal_ = 0;
    std::size_t  numGlobal_ = 0;
    bool         firstCall_ = true;
    std::vector<std::pair<void*, std::size_t>> ownSpares_;
};

} // namespace sphx

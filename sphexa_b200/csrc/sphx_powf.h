/*! @file
 * Bit-exact emulation of glibc's powf (the function behind std::pow(float, float) in sph::updateH,
 * sph/include/sph/kernels.hpp:26-32) for positive normal x and finite y.
 *
 * Why: h is updated with updateH inside the coupled h / neighbour-count iteration and again every step; a different
 * last bit of h changes the search radius and can flip a borderline neighbour, which would break the bit-exact
 * neighbour lists. glibc's powf (>= 2.28, the ARM optimized-routines algorithm: log2 by a 16-entry table + degree-5
 * polynomial, exp2 by a 32-entry table + cubic, all in double) has ~0.52 ulp error, i.e. it is NOT correctly rounded
 * in ~0.5 % of cases, so neither CUDA's powf nor a correctly rounded pow reproduces it. The x86-64 build selects the
 * FMA variant on every CPU with AVX2+FMA (sysdeps/x86_64/fpu/multiarch/e_powf-fma.c), in which every a*b+c of the
 * algorithm is a fused multiply-add; that is what is restated here. Dependency: glibc 2.39 (Ubuntu 2.39-0ubuntu8.5),
 * algorithm unchanged since 2.28; tests/test_host_tree.py::test_powf_emulation compares against the libm of the box
 * the tests run on.
 */
#pragma once

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SPHX_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define SPHX_HD inline
#endif

namespace sphx
{

// {invc, logc} for the 16 sub-intervals of [0x3f330000, 2*0x3f330000) (glibc __powf_log2_data.tab)
#define SPHX_POWF_LOG2_TAB                                                                                             \
    {                                                                                                                  \
        0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,                      \
            0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,                  \
            0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3,                  \
            0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,                  \
            0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1.0000000000000p+0, 0x0.0p+0, 0x1.e608cfd9a47acp-1,         \
            0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3, 0x1.b2036576afce6p-1,                    \
            0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2, 0x1.886e6037841edp-1,                    \
            0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2                                           \
    }

// bits(2^(i/32)) - (i << 47) (glibc __exp2f_data.tab)
#define SPHX_EXP2F_TAB                                                                                                 \
    {                                                                                                                  \
        0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,                    \
            0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,                \
            0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,                \
            0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,                \
            0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,                \
            0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,                \
            0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,                \
            0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull                 \
    }

#ifdef __CUDACC__
static __device__ const double   d_powfLog2Tab[32] = SPHX_POWF_LOG2_TAB;
static __device__ const uint64_t d_exp2fTab[32]    = SPHX_EXP2F_TAB;
#endif
static const double   h_powfLog2Tab[32] = SPHX_POWF_LOG2_TAB;
static const uint64_t h_exp2fTab[32]    = SPHX_EXP2F_TAB;

//! powf(x, y) as computed by glibc's FMA variant, for x a positive normal float and |y log2 x| < 126
SPHX_HD float glibcPowf(float x, float y)
{
#ifdef __CUDA_ARCH__
    const double*   logTab = d_powfLog2Tab;
    const uint64_t* expTab = d_exp2fTab;
#define SPHX_FMA(a, b, c) fma((a), (b), (c))
#define SPHX_DMUL(a, b) __dmul_rn((a), (b))
#define SPHX_DADD(a, b) __dadd_rn((a), (b))
#else
    const double*   logTab = h_powfLog2Tab;
    const uint64_t* expTab = h_exp2fTab;
#define SPHX_FMA(a, b, c) std::fma((a), (b), (c))
#define SPHX_DMUL(a, b) ((a) * (b))
#define SPHX_DADD(a, b) ((a) + (b))
#endif
    uint32_t ix;
    memcpy(&ix, &x, 4);

    // log2_inline: x = 2^k z, z in [OFF, 2 OFF); log2(x) = log1p(z/c - 1)/ln2 + log2(c) + k
    uint32_t tmp = ix - 0x3f330000u;
    uint32_t i   = (tmp >> 19) & 15u;
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz  = ix - top;
    int      k   = int(top) >> 23;
    float    zf;
    memcpy(&zf, &iz, 4);
    double z    = double(zf);
    double invc = logTab[2 * i], logc = logTab[2 * i + 1];

    const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2,
                 A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
    double r  = SPHX_FMA(z, invc, -1.0);
    double y0 = SPHX_DADD(logc, double(k));
    double r2 = SPHX_DMUL(r, r);
    double yy = SPHX_FMA(A0, r, A1);
    double p  = SPHX_FMA(A2, r, A3);
    double r4 = SPHX_DMUL(r2, r2);
    double q  = SPHX_FMA(A4, r, y0);
    q         = SPHX_FMA(p, r2, q);
    yy        = SPHX_FMA(yy, r4, q);

    double ylogx = SPHX_DMUL(double(y), yy);

    // exp2_inline: x = k/N + r, exp2(x) = 2^(k/N) (C0 r^3 + C1 r^2 + C2 r + 1), N = 32
    const double SHIFT = 0x1.8p+47; // 0x1.8p52 / 32
    const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
    double       kd = SPHX_DADD(ylogx, SHIFT);
    uint64_t     ki;
    memcpy(&ki, &kd, 8);
    kd          = SPHX_DADD(kd, -SHIFT);
    double rr   = SPHX_DADD(ylogx, -kd);
    uint64_t t  = expTab[ki & 31u];
    t += ki << 47;
    double s;
    memcpy(&s, &t, 8);
    double zz  = SPHX_FMA(C0, rr, C1);
    double rr2 = SPHX_DMUL(rr, rr);
    double e   = SPHX_FMA(C2, rr, 1.0);
    e          = SPHX_FMA(zz, rr2, e);
    e          = SPHX_DMUL(e, s);
    return float(e);
#undef SPHX_FMA
#undef SPHX_DMUL
#undef SPHX_DADD
}

/*! @brief sph::updateH (sph/include/sph/kernels.hpp:26-32), T = float:
 *         h * 0.5f * powf(1.0f + 1023.0f * ng0 / float(nc), 0.1f), every float operation rounded separately */
SPHX_HD float updateHExact(unsigned ng0, unsigned nc, float h)
{
#ifdef __CUDA_ARCH__
    float base = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(1023.0f, float(ng0)), float(nc)));
    return __fmul_rn(__fmul_rn(h, 0.5f), glibcPowf(base, 0.1f));
#else
    volatile float a    = 1023.0f * float(ng0);
    volatile float b    = a / float(nc);
    volatile float base = 1.0f + b;
    volatile float hh   = h * 0.5f;
    return hh * glibcPowf(base, 0.1f);
#endif
}

} // namespace sphx

/*! @file
 * Host half of the turbulence driver: stirring modes, Ornstein-Uhlenbeck phases, projection.
 *
 * Replaces (reference paths relative to /root/reference):
 *   sph::TurbulenceData::initModes   sph/include/sph/hydro_turb/turbulence_data.hpp:143-177
 *   sph::createStirringModes         sph/include/sph/hydro_turb/create_modes.hpp:33-227
 *   sph::updateNoise                 sph/include/sph/hydro_turb/driver.hpp:85-98
 *   sph::computePhases               sph/include/sph/hydro_turb/phases.hpp:46-72
 *
 * The random numbers are std::mt19937 + std::normal_distribution / std::uniform_real_distribution of the C++ standard
 * library, as in the reference, consumed in the same order, so the phase sequence is the reference's. Compiled with
 * -ffp-contract=off (the oracle's parity build): every operation rounded separately, in the reference's order.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <sstream>
#include <string>

#include "sphx_turbulence.h"

namespace
{

constexpr int kDim = 3; // TurbulenceData::numDim is fixed at 3 (turbulence_data.hpp:63)

void pushMode(SphxTurbulence& t, double kx, double ky, double kz, double amplitude)
{
    t.modes.push_back(kx);
    t.modes.push_back(ky);
    t.modes.push_back(kz);
    t.amplitudes.push_back(amplitude);
}

//! band (0) and parabolic (1) spectrum: every lattice wave vector with stirMin <= |k| <= stirMax (create_modes.hpp:75-158)
void fullSampling(SphxTurbulence& t, double L, size_t maxModes, double stirMax, double stirMin, int spectForm)
{
    const double twopi = 2.0 * M_PI;
    double       kc    = (spectForm == 1) ? 0.5 * (stirMin + stirMax) : stirMin;
    double       parab = -4.0 / ((stirMax - stirMin) * (stirMax - stirMin));

    // |k| >= every component, so indices beyond stirMax L / 2 pi cannot qualify (the reference scans 0..256)
    size_t ikEnd = std::min<size_t>(256, size_t(stirMax * L / twopi) + 1);
    for (size_t ikx = 0; ikx <= ikEnd; ikx++)
    {
        double kx = twopi * ikx / L;
        for (size_t iky = 0; iky <= ikEnd; iky++)
        {
            double ky = twopi * iky / L;
            for (size_t ikz = 0; ikz <= ikEnd; ikz++)
            {
                double kz = twopi * ikz / L;
                double k  = std::sqrt(kx * kx + ky * ky + kz * kz);
                if (!(k >= stirMin && k <= stirMax)) { continue; }
                if (t.amplitudes.size() + 4 > maxModes) { break; } // "Too many stirring modes" (:93-99)

                double amplitude = 1.0;
                if (spectForm == 1) { amplitude = std::abs(parab * (k - kc) * (k - kc) + 1.0); }
                amplitude = 2.0 * std::sqrt(amplitude) * std::pow(kc / k, 0.5 * (kDim - 1));

                pushMode(t, kx, ky, kz, amplitude);
                pushMode(t, kx, -ky, kz, amplitude);
                pushMode(t, kx, ky, -kz, amplitude);
                pushMode(t, kx, -ky, -kz, amplitude);
            }
        }
    }
}

//! power-law spectrum (2): random directions on k-shells, drawn from the engine BEFORE the phases (:161-222)
void shellSampling(SphxTurbulence& t, double L, size_t maxModes, double stirMax, double stirMin, double powerLawExp,
                   double anglesExp)
{
    const double                           twopi = 2.0 * M_PI;
    double                                 kc    = stirMin;
    std::uniform_real_distribution<double> uni(0, 1);

    int ikmin = std::max(1, int(stirMin * L / twopi + 0.5));
    int ikmax = int(stirMax * L / twopi + 0.5);
    for (int ik = ikmin; ik <= ikmax; ik++)
    {
        int nang = std::pow(2, kDim) * std::ceil(std::pow(ik, anglesExp));
        for (int iang = 1; iang <= nang; iang++)
        {
            double phi   = twopi * uni(t.gen);
            double theta = std::acos(1.0 - 2.0 * uni(t.gen));
            double rnd   = ik + uni(t.gen) - 0.5;
            double kx    = twopi * std::round(rnd * std::sin(theta) * std::cos(phi)) / L;
            double ky    = twopi * std::round(rnd * std::sin(theta) * std::sin(phi)) / L;
            double kz    = twopi * std::round(rnd * std::cos(theta)) / L;
            double k     = std::sqrt(kx * kx + ky * ky + kz * kz);
            if (!(k >= stirMin && k <= stirMax)) { continue; }
            if (t.amplitudes.size() + 4 > maxModes) { break; }

            double amplitude = std::pow(k / kc, powerLawExp);
            amplitude        = std::sqrt(amplitude * (std::pow(ik, kDim - 1) * 4.0 * (std::sqrt(3.0)) / nang)) *
                        std::pow(kc / k, (kDim - 1) / 2.0);
            pushMode(t, kx, ky, kz, amplitude);
        }
    }
}

//! modes[3 m + d] == 2 pi i / L exactly for small integers i? Then the device evaluates 3 (maxIdx) sincos per particle
//! instead of 6 per mode.
void detectLattice(SphxTurbulence& t)
{
    const double twopi = 2.0 * M_PI;
    t.latticeIdx.assign(t.modes.size(), 0);
    t.lattice = !t.modes.empty();
    t.maxIdx  = 0;
    for (size_t c = 0; c < t.modes.size() && t.lattice; ++c)
    {
        double v = t.modes[c];
        long   i = std::lround(std::abs(v) * t.Lbox / twopi);
        if (i > 15 || twopi * double(i) / t.Lbox != std::abs(v)) { t.lattice = false; }
        else
        {
            t.latticeIdx[c] = int8_t(std::signbit(v) ? -i : i);
            t.maxIdx        = std::max(t.maxIdx, int(i));
        }
    }
    if (!t.lattice) { t.maxIdx = 0; }
}

} // namespace

namespace sphx
{

void turbulenceAdvance(SphxTurbulence& t, double dt)
{
    double dampingA = std::exp(-dt / t.decayTime);
    double dampingB = std::sqrt(1.0 - dampingA * dampingA);
    // a fresh distribution object per call, as the reference: a cached second Gaussian never carries over
    std::normal_distribution<double> dist(0, 1);
    for (double& p : t.phases)
    {
        double r = dist(t.gen);
        p        = p * dampingA + t.variance * dampingB * r;
    }
}

void turbulenceProject(SphxTurbulence& t)
{
    const double w = t.solWeight;
    for (size_t i = 0; i < t.numModes; i++)
    {
        const double* k  = &t.modes[3 * i];
        const double* ou = &t.phases[6 * i];
        double        ka = 0.0, kb = 0.0, kk = 0.0;
        for (int j = 0; j < kDim; j++)
        {
            kk = kk + k[j] * k[j];
            ka = ka + k[j] * ou[2 * j + 1];
            kb = kb + k[j] * ou[2 * j];
        }
        for (int j = 0; j < kDim; j++)
        {
            double diva  = k[j] * ka / kk;
            double divb  = k[j] * kb / kk;
            double curla = ou[2 * j] - divb;
            double curlb = ou[2 * j + 1] - diva;

            t.phasesReal[3 * i + j] = w * curla + (1.0 - w) * divb;
            t.phasesImag[3 * i + j] = w * curlb + (1.0 - w) * diva;
        }
    }
}

} // namespace sphx

extern "C"
{

int sphx_turbulence_create(const SphxTurbulenceSettings* s, SphxTurbulence** out)
{
    if (!s || !out || !(s->Lbox > 0) || !(s->stMachVelocity > 0) || s->stSpectForm < 0 || s->stSpectForm > 2)
    {
        return SPHX_ERR_INVALID;
    }
    auto* t      = new SphxTurbulence;
    t->solWeight = s->solWeight;
    t->Lbox      = s->Lbox;
    t->gen       = std::mt19937(size_t(s->rngSeed));

    const double twopi   = 2.0 * M_PI;
    double       energy  = s->stEnergyPrefac * std::pow(s->stMachVelocity, 3) / s->Lbox;
    double       stirMin = (1.0 - s->epsilon) * twopi / s->Lbox;
    double       stirMax = (3.0 + s->epsilon) * twopi / s->Lbox;

    t->decayTime     = s->Lbox / (2.0 * s->stMachVelocity);
    t->variance      = std::sqrt(energy / t->decayTime);
    t->solWeightNorm = std::sqrt(3.0) * std::sqrt(3.0 / double(kDim)) /
                       std::sqrt(1.0 - 2.0 * t->solWeight + double(kDim) * t->solWeight * t->solWeight);

    if (s->stSpectForm == 2)
    {
        shellSampling(*t, s->Lbox, s->stMaxModes, stirMax, stirMin, s->powerLawExp, s->anglesExp);
    }
    else { fullSampling(*t, s->Lbox, s->stMaxModes, stirMax, stirMin, s->stSpectForm); }

    t->numModes = t->amplitudes.size();
    t->phases.resize(2 * kDim * t->numModes);
    t->phasesReal.assign(kDim * t->numModes, 0.0);
    t->phasesImag.assign(kDim * t->numModes, 0.0);
    detectLattice(*t);

    // initial phases: Gaussian with standard deviation "variance" (turbulence_data.hpp:174-176)
    std::normal_distribution<double> dist(0, t->variance);
    std::generate(t->phases.begin(), t->phases.end(), [t, &dist]() { return dist(t->gen); });

    *out = t;
    return SPHX_OK;
}

void sphx_turbulence_free(SphxTurbulence* t)
{
    if (!t) return;
    sphx::turbulenceFreeDevice(*t);
    delete t;
}

static std::string engineText(const SphxTurbulence* t)
{
    std::stringstream s;
    s << t->gen;
    return s.str();
}

void sphx_turbulence_sizes(const SphxTurbulence* t, size_t sizes[4])
{
    sizes[0] = t->numModes;
    sizes[1] = t->lattice ? 1 : 0;
    sizes[2] = size_t(t->maxIdx);
    sizes[3] = engineText(t).size() + 1;
}

void sphx_turbulence_get(const SphxTurbulence* t, double* modes, double* amplitudes, double* phases, double* phasesReal,
                         double* phasesImag, double* scalars, char* rngState)
{
    auto copy = [](double* dst, const std::vector<double>& v)
    {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(double));
    };
    copy(modes, t->modes);
    copy(amplitudes, t->amplitudes);
    copy(phases, t->phases);
    copy(phasesReal, t->phasesReal);
    copy(phasesImag, t->phasesImag);
    if (scalars)
    {
        scalars[0] = t->variance;
        scalars[1] = t->decayTime;
        scalars[2] = t->solWeight;
        scalars[3] = t->solWeightNorm;
    }
    if (rngState)
    {
        std::string e = engineText(t);
        std::memcpy(rngState, e.c_str(), e.size() + 1);
    }
}

int sphx_turbulence_restore(SphxTurbulence* t, size_t numModes, const double* modes, const double* amplitudes,
                            const double* phases, const double* scalars, const char* rngState)
{
    if (!t) return SPHX_ERR_INVALID;
    if (modes || amplitudes)
    {
        if (!modes || !amplitudes || !phases) return SPHX_ERR_INVALID;
        t->numModes = numModes;
        t->modes.assign(modes, modes + 3 * numModes);
        t->amplitudes.assign(amplitudes, amplitudes + numModes);
        t->phases.resize(6 * numModes);
        t->phasesReal.assign(3 * numModes, 0.0);
        t->phasesImag.assign(3 * numModes, 0.0);
        detectLattice(*t);
        sphx::turbulenceFreeDevice(*t); // the device tables are rebuilt by the next stirring call
    }
    else if (numModes != t->numModes && phases) { return SPHX_ERR_INVALID; }
    if (phases) { std::memcpy(t->phases.data(), phases, t->phases.size() * sizeof(double)); }
    if (scalars)
    {
        t->variance      = scalars[0];
        t->decayTime     = scalars[1];
        t->solWeight     = scalars[2];
        t->solWeightNorm = scalars[3];
    }
    if (rngState)
    {
        std::stringstream s;
        s << rngState;
        std::mt19937 g;
        s >> g;
        if (s.fail()) return SPHX_ERR_INVALID;
        t->gen = g;
    }
    return SPHX_OK;
}

int sphx_turbulence_advance_host(SphxTurbulence* t, double minDt)
{
    if (!t) return SPHX_ERR_INVALID;
    sphx::turbulenceAdvance(*t, minDt);
    sphx::turbulenceProject(*t);
    return SPHX_OK;
}

} // extern "C"

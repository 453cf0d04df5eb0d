/*! @file internal: host state of the turbulence driver, shared between turbulence_host.cpp and turbulence.cu */
#pragma once

#include <cstdint>
#include <random>
#include <vector>

#include "sphx.h"

struct SphxTurbulence
{
    // TurbulenceData (sph/include/sph/hydro_turb/turbulence_data.hpp:63-79)
    double variance{0}, decayTime{0}, solWeight{0}, solWeightNorm{0}, Lbox{1};
    size_t numModes{0};

    std::vector<double> modes, amplitudes, phases, phasesReal, phasesImag;
    std::mt19937        gen;

    // lattice form of the modes: modes[3 m + d] == 2 pi latticeIdx[3 m + d] / Lbox exactly
    bool                lattice{false};
    int                 maxIdx{0};
    std::vector<int8_t> latticeIdx;

    // device tables (turbulence.cu)
    double* d_modes{nullptr};
    double* d_amplitudes{nullptr};
    double* d_phases{nullptr}; // phasesReal (3 numModes) followed by phasesImag (3 numModes)
    int8_t* d_latticeIdx{nullptr};
    bool    uploaded{false};
};

namespace sphx
{

//! one Ornstein-Uhlenbeck step of the phases (driver.hpp:85-98) and the projection (phases.hpp:46-72)
void turbulenceAdvance(SphxTurbulence& t, double dt);
void turbulenceProject(SphxTurbulence& t);
void turbulenceFreeDevice(SphxTurbulence& t);

} // namespace sphx

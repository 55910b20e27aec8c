// Virtual volume move ("virtualvolume", excess pressure by perturbation), restated caller side:
// src/analysis.cpp:825-843 (VirtualVolumeMove::_sample), :883-904 (configuration, results), :807-823
// (PerturbationAnalysis::collectWidomAverage / meanFreeEnergy). It calls energy(change) on the ACCEPTED Hamiltonian
// with change = {everything, volume_change} — without updateState and without sync — before and after
// scaling the Space, then scales back. Nothing device specific: the adaptor terms serve such callers by refreshing
// their mirror from the Space they are bound to.
#pragma once
#include "energyterm.hpp"

namespace fb {

class VirtualVolumeMove
{
    Space& spc;
    Hamiltonian& pot;
    double volume_displacement = 0;
    VolumeMethod method = VolumeMethod::ISOTROPIC;
    Change change;

  public:
    double sum_exp = 0;        //!< Σ exp(−ΔU)  (Average::value_sum of mean_exponentiated_energy_change)
    unsigned long count = 0;   //!< samples collected
    double last_energy_change = 0;

    VirtualVolumeMove(const Json& j, Space& spc, Hamiltonian& pot)
        : spc(spc)
        , pot(pot)
    {
        volume_displacement = j.at("dV").number();
        method = volumeMethodFromString(j.value("scaling", "isotropic"));
        if (method == VolumeMethod::ISOCHORIC) {
            throw std::runtime_error("isochoric volume scaling not allowed");
        }
        change.volume_change = true;
        change.everything = true;
    }

    /** VirtualVolumeMove::_sample */
    void sample()
    {
        if (std::fabs(volume_displacement) <= pc::epsilon_dbl) {
            return;
        }
        const double old_volume = spc.geometry.getVolume();
        const double old_energy = pot.energy(change);
        spc.scaleVolume(old_volume + volume_displacement, method);
        const double new_energy = pot.energy(change);
        spc.scaleVolume(old_volume, method); // restore
        const double energy_change = new_energy - old_energy;
        last_energy_change = energy_change;
        if (-energy_change > pc::max_exp_argument) {
            return; // collectWidomAverage: skipped, not counted
        }
        sum_exp += std::exp(-energy_change);
        count++;
    }

    double meanFreeEnergy() const { return -std::log(sum_exp / static_cast<double>(count)); }
    /** excess pressure in kT/Å³ */
    double excessPressure() const { return -meanFreeEnergy() / volume_displacement; }
    double displacement() const { return volume_displacement; }
};

} // namespace fb

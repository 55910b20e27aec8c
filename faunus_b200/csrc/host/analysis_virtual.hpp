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

/**
 * Virtual translation of the ONE active molecule of a kind ("virtualtranslate", mean force by perturbation):
 * src/analysis.cpp:2794-2860, :2874-2880. energy(change = {group, internal = false, no indices → whole group}) on the
 * accepted Hamiltonian before and after displacing the group, then the group is moved back — again without
 * updateState / sync. The molecule is picked with the GLOBAL generator (one draw even for a single candidate).
 */
class VirtualTranslate
{
    Space& spc;
    Hamiltonian& pot;
    Random& random;
    int molid = -1;
    Point direction{0.0, 0.0, 1.0};
    double distance = 0;
    Change change;

  public:
    double sum_exp = 0;
    unsigned long count = 0;
    double last_energy_change = 0;

    VirtualTranslate(const Json& j, Space& spc, Hamiltonian& pot, Random& global)
        : spc(spc)
        , pot(pot)
        , random(global)
    {
        change.groups.resize(1);
        change.groups.front().internal = false;
        molid = spc.topology->moleculeId(j.at("molecule").string());
        if (spc.topology->molecules.at(molid).atomic) {
            throw std::runtime_error("atomic molecule " + spc.topology->molecules[molid].name + " not allowed");
        }
        distance = j.at("dL").number();
        if (const auto* d = j.find("dir")) {
            direction = pointFromJson(*d);
        }
        direction = direction / direction.norm();
    }

    /** VirtualTranslate::_sample + momentarilyPerturb */
    void sample()
    {
        if (std::fabs(distance) < pc::epsilon_dbl) {
            return;
        }
        const auto mollist = spc.findMolecules(molid, Space::Selection::ACTIVE);
        if (mollist.empty()) {
            return;
        }
        if (mollist.size() > 1) {
            throw std::runtime_error("exactly ONE active molecule expected");
        }
        const auto group_index = mollist[random.sampleIndex(static_cast<int>(mollist.size()))];
        auto& group = spc.groups[group_index];
        if (group.empty()) {
            return;
        }
        change.groups.at(0).group_index = group_index;
        const double old_energy = pot.energy(change);
        const Point displacement = direction * distance;
        spc.translate(group, displacement);
        const double new_energy = pot.energy(change);
        spc.translate(group, displacement * -1.0);
        const double energy_change = new_energy - old_energy;
        last_energy_change = energy_change;
        if (-energy_change > pc::max_exp_argument) {
            return;
        }
        sum_exp += std::exp(-energy_change);
        count++;
    }

    double meanFreeEnergy() const { return -std::log(sum_exp / static_cast<double>(count)); }
    /** mean force in kT/Å */
    double meanForce() const { return -meanFreeEnergy() / distance; }
};

} // namespace fb

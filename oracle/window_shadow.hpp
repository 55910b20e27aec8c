// ORACLE (test infrastructure only): a stand-in for the device behind MetropolisMonteCarlo's windowed engine, so
// that the engine's host logic — drawing ahead in generator order, conditional proposals (both outcomes of a move
// on an atom that an undecided earlier move touches), evaluations queued behind each other, replay of the
// decisions into the two Spaces — is exercised on the CPU (`-m "not gpu"`).
//
// The shadow is a second simulation built from the same input. Each submitted proposal is carried out on it ONE AT
// A TIME by the reference protocol (src/montecarlo.cpp:151-175: updateState → energy(trial) → energy(accepted) →
// Metropolis with the shipped uniform → sync), exactly as the one-move-at-a-time engine would. It checks that the
// start position shipped with every proposal is where its own copy of the atom stands — i.e. that the engine's
// bookkeeping and the "device" never drift apart — and returns decisions and energies like fb_run_wait.
#pragma once
#include "../faunus_b200/csrc/host/montecarlo.hpp"
#include <deque>

namespace oracle {

class ShadowWindowEvaluator : public fb::WindowEvaluator
{
    struct Packed
    {
        fb::Change change;
        fb::Point start[2], trial[2]; //!< [0]: plain proposal, or "dependency accepted"; [1]: "dependency rejected"
        bool group = false;           //!< rigid-molecule move: all atoms and the mass centre travel
        std::vector<fb::Point> group_start, group_trial;
        fb::Point cm_start, cm_trial;
        bool conditional = false;
        uint64_t serial = 0, dependency = 0;
        double uniform = 0;
    };
    struct Outcome
    {
        int accepted = 0;
        double u_new = 0, u_old = 0;
    };
    fb::MetropolisMonteCarlo& mc;                      //!< the engine under test
    std::unique_ptr<fb::MetropolisMonteCarlo> shadow;  //!< the stand-in "device"
    int moves_per_evaluation;
    std::vector<Packed> prepared;
    std::deque<std::vector<Outcome>> in_flight;        //!< evaluations submitted and not yet waited for
    std::vector<Outcome> current;                      //!< the evaluation wait() returned last
    std::unordered_map<uint64_t, int> decided;         //!< serial → accepted, of everything carried out so far

    void carryOut(const std::vector<Packed>& batch)
    {
        std::vector<Outcome> outcomes;
        for (const auto& p : batch) {
            int variant = 0;
            if (p.conditional) {
                const auto found = decided.find(p.dependency);
                if (found == decided.end()) {
                    throw std::runtime_error("shadow: a conditional proposal arrived before the move it depends on");
                }
                variant = found->second ? 0 : 1;
            }
            const auto& gc = p.change.groups.at(0);
            auto& trial_spc = *shadow->trial_state.spc;
            auto& spc = *shadow->state.spc;
            auto same = [](const fb::Point& a, const fb::Point& b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
            if (p.group) {
                auto& trial_group = trial_spc.groups.at(gc.group_index);
                const auto& group = spc.groups.at(gc.group_index);
                if (group.size() != p.group_start.size() || !same(group.mass_center, p.cm_start)) {
                    throw std::runtime_error("shadow: the molecule shipped with a proposal is not where the molecule is");
                }
                for (size_t i = 0; i < group.size(); ++i) {
                    if (!same(spc.at(group, i).pos, p.group_start[i])) {
                        throw std::runtime_error("shadow: the molecule shipped with a proposal is not where the molecule is");
                    }
                    trial_spc.at(trial_group, i).pos = p.group_trial[i];
                }
                trial_group.mass_center = p.cm_trial;
            }
            else {
                auto& particle = trial_spc.at(trial_spc.groups.at(gc.group_index), gc.relative_atom_indices.at(0));
                const auto& accepted_particle = spc.at(spc.groups.at(gc.group_index), gc.relative_atom_indices[0]);
                if (!same(accepted_particle.pos, p.start[variant])) {
                    throw std::runtime_error("shadow: the start position shipped with a proposal is not where the atom is");
                }
                particle.pos = p.trial[variant];
            }
            shadow->trial_state.pot->updateState(p.change);
            Outcome o;
            o.u_new = shadow->trial_state.pot->energy(p.change);
            o.u_old = shadow->state.pot->energy(p.change);
            const double du = fb::MetropolisMonteCarlo::getEnergyChange(o.u_new, o.u_old);
            o.accepted = fb::MetropolisMonteCarlo::metropolisDecision(du, p.uniform) ? 1 : 0;
            if (o.accepted) {
                shadow->state.sync(shadow->trial_state, p.change);
            }
            else {
                shadow->trial_state.sync(shadow->state, p.change);
            }
            decided[p.serial] = o.accepted;
            outcomes.push_back(o);
        }
        in_flight.push_back(std::move(outcomes));
    }

  public:
    unsigned long conditional_proposals = 0, queued_evaluations = 0, evaluations = 0;

    ShadowWindowEvaluator(fb::MetropolisMonteCarlo& mc, std::unique_ptr<fb::MetropolisMonteCarlo> shadow, int moves)
        : mc(mc)
        , shadow(std::move(shadow))
        , moves_per_evaluation(moves)
    {
    }

    int capacity() const override { return moves_per_evaluation; }
    bool supports(fb::WindowProposal::Kind) const override { return true; } // single atoms and rigid molecules
    bool conditionals() const override { return true; }
    bool pipelined(const std::vector<fb::WindowProposal>&, int, int ready) const override { return ready > 0; }

    void prepare(const std::vector<fb::WindowProposal>& window, int first, int n) override
    {
        prepared.clear();
        const auto& trial = *mc.trial_state.spc;
        const auto& accepted = *mc.state.spc;
        for (int m = 0; m < n; ++m) {
            const auto& w = window[first + m];
            Packed p;
            p.change = w.change;
            p.serial = w.serial;
            p.uniform = w.uniform;
            if (w.kind == fb::WindowProposal::Kind::GROUP) {
                const auto& gc = w.change.groups.at(0);
                const auto& trial_group = trial.groups.at(gc.group_index);
                const auto& group = accepted.groups.at(gc.group_index);
                p.group = true;
                p.cm_trial = trial_group.mass_center;
                p.cm_start = group.mass_center;
                for (size_t i = 0; i < group.size(); ++i) {
                    p.group_trial.push_back(trial.at(trial_group, i).pos);
                    p.group_start.push_back(accepted.at(group, i).pos);
                }
            }
            else if (w.conditional && !w.applied) {
                p.conditional = true;
                p.dependency = w.dependency;
                for (int v = 0; v < 2; ++v) {
                    p.start[v] = w.alt_start[v];
                    p.trial[v] = w.alt_new[v];
                }
                conditional_proposals++;
            }
            else {
                const auto& gc = w.change.groups.at(0);
                p.trial[0] = trial.at(trial.groups.at(gc.group_index), gc.relative_atom_indices.at(0)).pos;
                p.start[0] = accepted.at(accepted.groups.at(gc.group_index), gc.relative_atom_indices[0]).pos;
            }
            prepared.push_back(std::move(p));
        }
    }

    void submitPrepared() override
    {
        if (!in_flight.empty()) {
            queued_evaluations++;
        }
        evaluations++;
        carryOut(prepared);
    }

    void submit(const std::vector<fb::WindowProposal>& window, int first, int n) override
    {
        prepare(window, first, n);
        submitPrepared();
    }

    void wait() override
    {
        if (in_flight.empty()) {
            throw std::runtime_error("shadow: nothing to wait for");
        }
        current = std::move(in_flight.front());
        in_flight.pop_front();
    }

    bool energies(int m, const std::vector<unsigned char>&, const fb::WindowProposal&, double& new_energy,
                  double& old_energy) override
    {
        new_energy = current.at(static_cast<size_t>(m)).u_new;
        old_energy = current[static_cast<size_t>(m)].u_old;
        return true;
    }

    int decision(int m) const override { return current.at(static_cast<size_t>(m)).accepted; }
    void commit(const std::vector<unsigned char>&) override {}
};

} // namespace oracle

// The Metropolis engine around the energy boundary: two complete states (accepted, trial), one
// trial move = mutate trial Space → Change → updateState → u_new, u_old → accept/reject → sync.
// Mirrors src/montecarlo.cpp:17-34 (metropolisCriterion), :43-71 (init), :85-99 (drift),
// :139-187 (performMove), :193-209 (getEnergyChange), :220-227 (sweep), :244-254 (State),
// src/analysis.cpp:807-823 + 1227-1312 (Widom insertion, sequential reference order).
#pragma once
#include <fstream>
#include <iterator>
#include "moves.hpp"
#include <chrono>
#include <algorithm>
#include <functional>
#include <unordered_map>
#include <unordered_set>

namespace fb {

struct State
{
    std::unique_ptr<Space> spc;
    std::unique_ptr<Hamiltonian> pot;
    void sync(const State& other, const Change& change)
    {
        spc->sync(*other.spc, change);
        pot->sync(other.pot.get(), change);
    }
};

/** One record per performed move for trace parity checks */
struct TraceRecord
{
    double du = 0;      //!< energy change new − old after the NaN/inf policy (before bias)
    double u_new = 0;   //!< trial Hamiltonian energy for the change
    double u_old = 0;   //!< accepted Hamiltonian energy for the change
    int accepted = 0;
    int move_id = 0;
};

/** One drawn, not yet decided trial move of the windowed path */
struct WindowProposal
{
    /** single atom of an atomic group (`transrot`) or a whole rigid molecule (`moltransrot`); a window is of one kind */
    enum class Kind
    {
        ATOM,
        GROUP
    };
    Kind kind = Kind::ATOM;
    Move* base = nullptr;                  //!< the move, whatever its kind (statistics)
    AtomicTranslateRotate* move = nullptr; //!< Kind::ATOM
    TranslateRotate* group_move = nullptr; //!< Kind::GROUP
    int move_id = 0;
    AtomicTranslateRotate::Draw draw;
    TranslateRotate::Draw group_draw;
    size_t key_group = 0;                  //!< what the proposal touches: a proposal on a (group, atom) that an
    long key_atom = -1;                    //!< earlier undecided one moves starts where that one leaves it
    /**
     * Conditional proposal: ONE earlier undecided proposal (`dependency`, by serial number) moves the same atom. Not
     * applied to the trial Space yet, but both outcomes are known: start/trial position if that proposal is accepted
     * ([0]) or rejected ([1]). An evaluator that decides on its own can take it along (WindowEvaluator::conditionals).
     */
    bool conditional = false;
    uint64_t serial = 0;
    uint64_t dependency = 0;
    Point alt_start[2], alt_new[2];
    Change change;                  //!< filled when the draw is applied to the trial Space
    bool applied = false;
    double uniform = 0;             //!< the Metropolis uniform of this move (drawn in reference order)
    double displacement_squared = 0;
};

/**
 * Evaluates a window of consecutive single-atom trial moves at once. The proposals of `transrot` and
 * the Metropolis uniform consume the generator independently of any energy (src/move.cpp:267-293,
 * src/montecarlo.cpp:17-34), so the engine can draw a run of moves ahead, have all of them evaluated
 * against the current accepted state in one pass, and then decide them strictly in order; the
 * evaluator supplies the energies of move m GIVEN the decisions taken on the earlier moves of the
 * window. The accept/reject sequence is the one of the one-at-a-time loop (src/montecarlo.cpp:139-187).
 */
class WindowEvaluator
{
  public:
    virtual ~WindowEvaluator() = default;
    virtual int capacity() const = 0;
    /** how many of the `ready` proposals of `window` from `first` on fit into one evaluation */
    virtual int fit(const std::vector<WindowProposal>& /*window*/, int /*first*/, int ready) const
    {
        return std::min(ready, capacity());
    }
    /** can proposals of this kind be evaluated at all? */
    virtual bool supports(WindowProposal::Kind kind) const { return kind == WindowProposal::Kind::ATOM; }
    /** start evaluating proposals [first, first + n) of `window` (all applied to the trial Space, distinct atoms) … */
    virtual void submit(const std::vector<WindowProposal>& window, int first, int n) = 0;
    /**
     * Pipelining: an evaluator that decides on its own (decision() ≥ 0) can take the NEXT evaluation while the
     * results of the one in flight are still out — the queued proposals behind it are on other atoms, so nothing
     * they need depends on those results. prepare() packs proposals [first, first + n) while the device works,
     * submitPrepared() queues them behind the evaluation in flight; wait() then returns the OLDER evaluation.
     */
    virtual bool pipelined(const std::vector<WindowProposal>& /*window*/, int /*first*/, int /*ready*/) const
    {
        return false;
    }
    /** can conditional proposals be part of an evaluation? */
    virtual bool conditionals() const { return false; }
    virtual void prepare(const std::vector<WindowProposal>& /*window*/, int /*first*/, int /*n*/) {}
    virtual void submitPrepared() {}
    /** … and wait for the results; the engine draws the next proposals in between */
    virtual void wait() = 0;
    /**
     * Hamiltonian energies of proposal m in the trial / accepted state, given which of the proposals
     * before m were accepted. Returns false if m has to be evaluated again in a fresh window.
     */
    virtual bool energies(int m, const std::vector<unsigned char>& accepted, const WindowProposal& proposal,
                          double& new_energy, double& old_energy) = 0;
    /** the evaluator already decided proposal m (1 accepted, 0 rejected; the caller replays it); −1: caller decides */
    virtual int decision(int /*m*/) const { return -1; }
    /** the first `accepted.size()` proposals are decided; the rest will be submitted again */
    virtual void commit(const std::vector<unsigned char>& accepted) = 0;
};

/**
 * Ideal (translational entropy) contribution of a change of the number of atoms / molecules to the trial energy;
 * src/montecarlo.cpp:271-374. Atom swaps (`dNswap`) are not restated.
 */
class TranslationalEntropy
{
    const Space& trial_spc;
    const Space& spc;

    /** src/montecarlo.cpp:282-299 */
    double bias(int trial_count, int count) const
    {
        double energy = 0.0;
        if (const int dN = trial_count - count; dN > 0) { // atoms or molecules were added
            const double V_trial = trial_spc.geometry.getVolume();
            for (int n = 0; n < dN; n++) {
                energy += std::log((count + 1 + n) / (V_trial * units::molar));
            }
        }
        else if (dN < 0) { // atoms or molecules were removed
            const double V = spc.geometry.getVolume();
            for (int n = 0; n < (-dN); n++) {
                energy -= std::log((count - n) / (V * units::molar));
            }
        }
        return energy;
    }
    double atomChangeEnergy(int molid) const // :320-333
    {
        const auto mollist_new = trial_spc.findMolecules(molid, Space::Selection::ALL);
        const auto mollist_old = spc.findMolecules(molid, Space::Selection::ALL);
        if (mollist_new.size() > 1 || mollist_old.size() > 1) {
            throw std::runtime_error("multiple atomic groups of the same type is not allowed");
        }
        return bias(static_cast<int>(trial_spc.groups.at(mollist_new.front()).size()),
                    static_cast<int>(spc.groups.at(mollist_old.front()).size()));
    }
    double moleculeChangeEnergy(int molid) const // :335-342
    {
        return bias(static_cast<int>(trial_spc.findMolecules(molid, Space::Selection::ACTIVE).size()),
                    static_cast<int>(spc.findMolecules(molid, Space::Selection::ACTIVE).size()));
    }

  public:
    TranslationalEntropy(const Space& trial_space, const Space& space)
        : trial_spc(trial_space)
        , spc(space)
    {
    }
    /** logarithm of the bias for the Metropolis criterion (kT); src/montecarlo.cpp:348-374 */
    double energy(const Change& change) const
    {
        double energy_change = 0.0;
        if (!change.matter_change || change.disable_translational_entropy) {
            return energy_change;
        }
        std::vector<int> already_processed;
        for (const auto& data : change.groups) {
            if (data.dNswap) {
                throw std::runtime_error("atom swap moves are outside the hot-path scope");
            }
            const int molid = trial_spc.groups.at(data.group_index).id;
            if (data.dNatomic && trial_spc.topology->molecules.at(molid).atomic) {
                energy_change += atomChangeEnergy(molid);
            }
            else if (std::find(already_processed.begin(), already_processed.end(), molid) == already_processed.end()) {
                energy_change += moleculeChangeEnergy(molid);
                already_processed.push_back(molid);
            }
        }
        return energy_change;
    }
};

class MetropolisMonteCarlo
{
  public:
    Randoms rng;
    std::shared_ptr<Topology> topology;
    State state;       //!< accepted
    State trial_state; //!< trial
    std::unique_ptr<MoveCollection> moves;
    double initial_energy = 0.0;
    double sum_of_energy_changes = 0.0;
    unsigned int number_of_sweeps = 0;
    bool record_trace = false;
    std::vector<TraceRecord> trace;
    std::unique_ptr<WindowEvaluator> window_evaluator; //!< set by the B200 build; nullptr = one move at a time
    unsigned long windows_evaluated = 0;
    unsigned long window_moves_evaluated = 0;
    double window_seconds_evaluate = 0; //!< host wall time inside WindowEvaluator::evaluate (launch + wait)
    double window_seconds_decide = 0;   //!< ... deciding moves and syncing the Spaces
    double window_seconds_total = 0;    //!< ... in the windowed part of sweeps (proposal drawing is the remainder)

  private:
    /** src/montecarlo.cpp:17-34; the uniform is ALWAYS drawn */
    bool metropolisCriterion(double energy_change)
    {
        if (std::isnan(energy_change)) {
            throw std::runtime_error("Metropolis error: energy cannot be NaN");
        }
        const double u = rng.slump();
        if (std::isinf(energy_change) && energy_change < 0.0) {
            return true;
        }
        if (-energy_change > pc::max_exp_argument) {
            return true;
        }
        return u <= std::exp(-energy_change);
    }

  public:
    /** src/montecarlo.cpp:193-209 */
    static double getEnergyChange(double new_energy, double old_energy)
    {
        if (std::isnan(old_energy) && !std::isnan(new_energy)) {
            return pc::neg_infty;
        }
        if (std::isnan(new_energy)) {
            return pc::infty;
        }
        if (new_energy > 0.0 && std::isinf(new_energy)) {
            return pc::infty;
        }
        const double energy_change = new_energy - old_energy;
        if (std::isnan(energy_change)) {
            return 0.0;
        }
        return energy_change;
    }

    /**
     * @param j full input document (temperature, geometry, atomlist, moleculelist,
     *          insertmolecules | groups+particles, energy, moves, random)
     */
    MetropolisMonteCarlo(const Json& j, const TermFactory& factory, ReplicaComm* comm = nullptr)
    {
        pc::temperature = j.at("temperature").number(); // src/faunus.cpp:106
        topology = topologyFromJson(j);
        auto make_state = [&](State& s) { // src/montecarlo.cpp:244-248
            s.spc = std::make_unique<Space>();
            s.spc->topology = topology;
            spaceFromJson(j, *s.spc, rng.global);
            s.pot = std::make_unique<Hamiltonian>(*s.spc, j.at("energy"), factory);
        };
        make_state(state);
        make_state(trial_state);
        static const Json no_moves = Json::array();
        const Json* jm = j.find("moves");
        moves = std::make_unique<MoveCollection>(jm ? *jm : no_moves, *trial_state.spc, rng, comm);
        init();
    }

    /** src/montecarlo.cpp:43-71 */
    void init()
    {
        sum_of_energy_changes = 0.0;
        Change change;
        change.everything = true;
        state.pot->state = EnergyTerm::MonteCarloState::ACCEPTED;
        trial_state.pot->state = EnergyTerm::MonteCarloState::TRIAL;
        state.pot->init();
        const double energy = state.pot->energy(change);
        initial_energy = energy;
        trial_state.sync(state, change);
        trial_state.pot->init();
        const double trial_energy = trial_state.pot->energy(change);
        if (std::isfinite(energy) && std::isfinite(trial_energy)) {
            if (std::fabs((energy - trial_energy) / energy) > 1e-6) {
                throw std::runtime_error("error aligning energies - this could be a bug...");
            }
        }
    }

    /**
     * Load a reference `state.json` into both Spaces, restore the two generators if the file carries them
     * (`random-move` = Move::slump, `random-global` = Faunus::random, each {"seed": "<engine state>"} as
     * src/random.cpp:10-45 writes them) and re-init; src/montecarlo.cpp:118-137
     */
    void restore(const Json& j)
    {
        if (!window.empty()) {
            throw std::runtime_error("restore: proposals are still in flight (restore between sweeps)");
        }
        state.spc->loadState(j);
        trial_state.spc->loadState(j);
        auto restore_generator = [&](const char* key, Random& generator) {
            if (const Json* node = j.find(key)) {
                const std::string seed = node->is_object() ? node->value("seed", std::string()) : std::string();
                if (!seed.empty() && seed != "default" && seed != "fixed") {
                    generator.setState(seed);
                }
            }
        };
        restore_generator("random-move", rng.slump);
        restore_generator("random-global", rng.global);
        init();
    }

    /** The state file of `savestate` with `saverandom: true`; src/analysis.cpp:656-682 */
    Json saveState() const
    {
        Json j = state.spc->toJson();
        auto generator = [](const Random& r) {
            Json g = Json::object();
            g["seed"] = r.state();
            g["engine"] = "Mersenne Twister (std::mt19937)";
            return g;
        };
        j["random-move"] = generator(rng.slump);
        j["random-global"] = generator(rng.global);
        return j;
    }

    /**
     * `savestate` to a file: `.json` (text) or `.ubj` (Universal Binary JSON, `json::to_ubjson`), with the two generators
     * when `save_random` (`saverandom`); any other suffix is a configuration error. src/analysis.cpp:640-682
     */
    void saveStateFile(const std::string& filename, bool save_random) const
    {
        const auto suffix = filename.substr(filename.find_last_of('.') + 1);
        if (suffix != "json" && suffix != "ubj") {
            throw std::runtime_error("unknown file extension for '" + filename + "'");
        }
        Json j = save_random ? saveState() : state.spc->toJson();
        std::ofstream f(filename, suffix == "ubj" ? std::ios::binary : std::ios::out);
        if (!f) {
            throw std::runtime_error("state file error -> " + filename);
        }
        const std::string bytes = suffix == "ubj" ? j.toUbjson() : j.dump();
        f.write(bytes.data(), static_cast<std::streamsize>(bytes.size()));
    }

    /** `--state <file>` (.json / .ubj): read, then restore(); src/faunus.cpp:430-455 */
    void restoreFile(const std::string& filename)
    {
        const auto suffix = filename.substr(filename.find_last_of('.') + 1);
        const bool binary = suffix == "ubj";
        std::ifstream f(filename, binary ? std::ios::binary : std::ios::in);
        if (!f) {
            throw std::runtime_error("state file error -> " + filename);
        }
        const std::string bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        restore(binary ? Json::fromUbjson(bytes) : Json::parse(bytes));
    }

    /** src/montecarlo.cpp:85-99 */
    double relativeEnergyDrift()
    {
        Change change;
        change.everything = true;
        const double energy = state.pot->energy(change);
        const double du = energy - initial_energy;
        if (std::isfinite(du)) {
            if (std::fabs(du) <= pc::epsilon_dbl) {
                return 0.0;
            }
            return (energy - (initial_energy + sum_of_energy_changes)) /
                   (std::fabs(initial_energy) > pc::epsilon_dbl ? initial_energy : energy);
        }
        return std::nan("");
    }

    /** src/montecarlo.cpp:139-187 */
    void performMove(Move& move, int move_id = 0)
    {
        Change change;
        move.move(change);
        if (change) {
            trial_state.pot->updateState(change);
            const double new_energy = trial_state.pot->energy(change);
            const double old_energy = state.pot->energy(change);
            double energy_change = getEnergyChange(new_energy, old_energy);
            const double energy_bias = move.bias(change, old_energy, new_energy) +
                                       TranslationalEntropy(*trial_state.spc, *state.spc).energy(change);
            const double total_trial_energy = energy_change + energy_bias;
            TraceRecord rec;
            rec.du = energy_change;
            rec.u_new = new_energy;
            rec.u_old = old_energy;
            rec.move_id = move_id;
            if (metropolisCriterion(total_trial_energy)) {
                state.sync(trial_state, change);
                move.accept(change);
                rec.accepted = 1;
            }
            else {
                trial_state.sync(state, change);
                move.reject(change);
                energy_change = 0.0;
            }
            sum_of_energy_changes += energy_change;
            if (record_trace) {
                trace.push_back(rec);
            }
        }
        else {
            rng.slump(); // keep the generator in sync, src/montecarlo.cpp:182-186
        }
    }

    /**
     * One trial "move" whose Change is made by the CALLER (a matter change: group sizes set on the trial Space,
     * the particles to be activated already in place), carried through the reference's protocol
     * (src/montecarlo.cpp:139-187): trial.updateState → trial.energy → accepted.energy → bias (translational
     * entropy) → accept (mode 1), reject (0) or Metropolis (2) → sync one way. What a SpeciationMove / GCMC move
     * asks of the Hamiltonian, without the move itself.
     */
    struct ExternalMoveResult
    {
        double new_energy = 0, old_energy = 0, bias = 0;
        bool accepted = false;
    };
    ExternalMoveResult performExternalChange(Change& change, int mode)
    {
        if (!window.empty()) {
            throw std::runtime_error("proposals are still in flight");
        }
        ExternalMoveResult r;
        trial_state.pot->updateState(change);
        r.new_energy = trial_state.pot->energy(change);
        r.old_energy = state.pot->energy(change);
        double energy_change = getEnergyChange(r.new_energy, r.old_energy);
        r.bias = TranslationalEntropy(*trial_state.spc, *state.spc).energy(change);
        r.accepted = mode == 1 || (mode == 2 && metropolisCriterion(energy_change + r.bias));
        if (r.accepted) {
            state.sync(trial_state, change);
        }
        else {
            trial_state.sync(state, change);
            energy_change = 0.0;
        }
        sum_of_energy_changes += energy_change;
        return r;
    }

    /** metropolisCriterion with the uniform drawn earlier (same rule, src/montecarlo.cpp:17-34) */
    static bool metropolisDecision(double energy_change, double uniform)
    {
        if (std::isnan(energy_change)) {
            throw std::runtime_error("Metropolis error: energy cannot be NaN");
        }
        if (std::isinf(energy_change) && energy_change < 0.0) {
            return true;
        }
        if (-energy_change > pc::max_exp_argument) {
            return true;
        }
        return uniform <= std::exp(-energy_change);
    }

  private:
    std::vector<WindowProposal> window; //!< drawn, undecided proposals in move order

    void applyProposal(WindowProposal& p)
    {
        if (p.kind == WindowProposal::Kind::GROUP) {
            p.group_move->moveFromDraw(p.group_draw, p.change);
        }
        else {
            p.move->moveFromDraw(p.draw, p.change);
            p.displacement_squared = p.move->latestDisplacementSquared();
        }
        p.applied = true;
    }

    static uint64_t proposalKey(const WindowProposal& p)
    {
        return (static_cast<uint64_t>(p.key_group) << 32) | static_cast<uint64_t>(static_cast<uint32_t>(p.key_atom + 1));
    }

    /** undecided proposals per touched atom / molecule, and which of them are applied to the trial Space */
    std::unordered_map<uint64_t, int> pending_keys;
    std::unordered_set<uint64_t> applied_keys;
    int unapplied_count = 0;
    uint64_t proposal_serial = 0;

    /** apply every queued proposal whose atom / molecule has no earlier undecided proposal any more */
    /** both outcomes of proposal `p` on an atom that the applied, undecided proposal `q` moves (see WindowProposal) */
    void makeConditional(WindowProposal& p, const WindowProposal& q)
    {
        const auto& trial_group = trial_state.spc->groups.at(p.draw.group_index);
        const auto& group = state.spc->groups.at(p.draw.group_index);
        p.conditional = true;
        p.dependency = q.serial;
        p.alt_start[0] = trial_state.spc->at(trial_group, p.draw.atom_index).pos; // q accepted: where q puts the atom
        p.alt_start[1] = state.spc->at(group, p.draw.atom_index).pos;             // q rejected: where the atom is
        p.alt_new[0] = p.move->displaced(p.alt_start[0], p.draw);
        p.alt_new[1] = p.move->displaced(p.alt_start[1], p.draw);
        p.move->describe(p.draw, p.change);
    }

    /**
     * After decisions: apply every queued proposal whose atom / molecule has no earlier undecided proposal any more,
     * and make those conditional that are left with exactly ONE (applied) undecided predecessor.
     */
    void applyUnblocked()
    {
        if (unapplied_count == 0) {
            return;
        }
        const bool conditionals = window_evaluator->conditionals();
        std::unordered_map<uint64_t, std::pair<int, const WindowProposal*>> earlier; // key → (undecided so far, the first)
        for (auto& p : window) {
            const uint64_t key = proposalKey(p);
            auto& [count, first] = earlier[key];
            if (!p.applied) {
                if (count == 0) {
                    applyProposal(p);
                    applied_keys.insert(key);
                    p.conditional = false;
                    unapplied_count--;
                }
                else if (count == 1 && !p.conditional && conditionals && p.kind == WindowProposal::Kind::ATOM &&
                         first->applied) {
                    makeConditional(p, *first);
                }
            }
            if (count == 0) {
                first = &p;
            }
            count++;
        }
    }

    int in_flight = 0; //!< proposals at the head of the queue whose evaluation is submitted

    /**
     * Evaluate the applied head of the queue (unless a pipelined evaluation of it is already in flight), draw
     * ahead and hand the next evaluation over while waiting, then decide / replay as many as possible, in order.
     */
    void decideWindow()
    {
        const auto t_begin = std::chrono::steady_clock::now();
        if (in_flight == 0) {
            in_flight = window_evaluator->fit(window, 0, readyProposals());
            window_evaluator->submit(window, 0, in_flight);
            windows_evaluated++;
            window_moves_evaluated += static_cast<unsigned long>(in_flight);
        }
        const int n = in_flight;
        const auto t_submitted = std::chrono::steady_clock::now();
        fillWindow(n + window_evaluator->capacity()); // the next proposals, while the device works
        int next_n = 0;
        if (static_cast<int>(window.size()) > n) {
            const int ready_behind = readyProposals() - n;
            if (ready_behind > 0 && window_evaluator->pipelined(window, n, ready_behind)) {
                next_n = window_evaluator->fit(window, n, ready_behind);
                window_evaluator->prepare(window, n, next_n);
            }
        }
        const auto t_filled = std::chrono::steady_clock::now();
        if (next_n > 0) { // queued behind the evaluation in flight: the device goes on without waiting for the host
            window_evaluator->submitPrepared();
            windows_evaluated++;
            window_moves_evaluated += static_cast<unsigned long>(next_n);
        }
        window_evaluator->wait(); // … for the OLDER evaluation
        const auto t_evaluated = std::chrono::steady_clock::now();
        window_seconds_evaluate += std::chrono::duration<double>((t_submitted - t_begin) + (t_evaluated - t_filled)).count();
        std::vector<unsigned char> accepted;
        for (int m = 0; m < n; ++m) {
            auto& p = window[m];
            double new_energy = 0, old_energy = 0;
            if (!window_evaluator->energies(m, accepted, p, new_energy, old_energy)) {
                break;
            }
            if (!p.applied) { // conditional: the proposal it depends on is decided (and replayed) by now
                applyProposal(p);
                applied_keys.insert(proposalKey(p));
                unapplied_count--;
            }
            double energy_change = getEnergyChange(new_energy, old_energy);
            TraceRecord rec;
            rec.du = energy_change;
            rec.u_new = new_energy;
            rec.u_old = old_energy;
            rec.move_id = p.move_id;
            if (p.move != nullptr) {
                p.move->setLatestDisplacementSquared(p.displacement_squared);
            }
            const int decided = window_evaluator->decision(m);
            if (decided >= 0 ? decided != 0 : metropolisDecision(energy_change, p.uniform)) {
                state.spc->sync(*trial_state.spc, p.change);
                p.base->accept(p.change);
                rec.accepted = 1;
            }
            else {
                trial_state.spc->sync(*state.spc, p.change);
                p.base->reject(p.change);
                energy_change = 0.0;
            }
            accepted.push_back(static_cast<unsigned char>(rec.accepted));
            sum_of_energy_changes += energy_change;
            if (record_trace) {
                trace.push_back(rec);
            }
        }
        if (accepted.empty()) {
            throw std::runtime_error("windowed evaluation made no progress");
        }
        window_evaluator->commit(accepted);
        // undecided proposals stay queued: their trial positions are in the trial Space (distinct atoms)
        for (size_t m = 0; m < accepted.size(); ++m) {
            const uint64_t key = proposalKey(window[m]);
            applied_keys.erase(key);
            auto it = pending_keys.find(key);
            if (--(it->second) == 0) {
                pending_keys.erase(it);
            }
        }
        window.erase(window.begin(), window.begin() + static_cast<long>(accepted.size()));
        applyUnblocked();
        if (next_n > 0 && static_cast<int>(accepted.size()) != n) {
            throw std::runtime_error("pipelined evaluation left proposals undecided");
        }
        in_flight = next_n;
        window_seconds_decide += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_evaluated).count();
    }

    /** leading proposals that can be evaluated: applied to the trial Space, or conditional (both outcomes known) */
    int readyProposals() const
    {
        int r = 0;
        while (r < static_cast<int>(window.size()) && (window[r].applied || window[r].conditional)) {
            r++;
        }
        return r;
    }

    unsigned int sweep_remaining = 0;     //!< stochastic moves of the current sweep not yet drawn
    Move* sweep_deferred = nullptr;       //!< a move of another kind: runs once the queue is empty
    std::function<int(Move&)> sweep_id_of;

    /** the kind of window `selected` can go into, if any */
    bool windowKind(Move* selected, WindowProposal::Kind& kind) const
    {
        if (auto* transrot = dynamic_cast<AtomicTranslateRotate*>(selected)) {
            kind = WindowProposal::Kind::ATOM;
            return transrot->targetsAtomicGroups() && window_evaluator->supports(kind);
        }
        if (dynamic_cast<TranslateRotate*>(selected) != nullptr) {
            kind = WindowProposal::Kind::GROUP;
            return window_evaluator->supports(kind);
        }
        return false;
    }

    /**
     * Draw the proposal of `selected` (generator order of performMove: proposal, then the Metropolis uniform) and
     * queue it; applied to the trial Space unless an earlier queued proposal touches the same atom / molecule.
     */
    void enqueue(Move* selected, WindowProposal::Kind kind)
    {
        WindowProposal p;
        p.kind = kind;
        p.base = selected;
        p.move_id = sweep_id_of(*selected);
        bool moves_something = false;
        if (kind == WindowProposal::Kind::GROUP) {
            p.group_move = static_cast<TranslateRotate*>(selected);
            p.group_draw = p.group_move->draw();
            p.key_group = p.group_draw.group_index;
            moves_something = p.group_draw.valid && ((p.group_draw.translate && p.group_draw.scalar != 0.0) ||
                                                     (p.group_draw.rotate && p.group_draw.angle != 0.0));
            if (!moves_something) {
                Change none; // empty Change: count the attempt, keep the generator in step (montecarlo.cpp:182-186)
                p.group_move->moveFromDraw(p.group_draw, none);
                if (!none.empty()) {
                    throw std::runtime_error("windowed moltransrot: unexpected non-empty change");
                }
            }
        }
        else {
            p.move = static_cast<AtomicTranslateRotate*>(selected);
            p.draw = p.move->draw();
            p.key_group = p.draw.group_index;
            p.key_atom = static_cast<long>(p.draw.atom_index);
            moves_something = p.draw.valid && (p.draw.dp > 0.0 || p.draw.dprot > 0.0);
            if (!moves_something) {
                Change none;
                p.move->moveFromDraw(p.draw, none);
            }
        }
        if (!moves_something) {
            rng.slump();
            return;
        }
        p.uniform = rng.slump();
        const uint64_t key = proposalKey(p);
        int& pending = pending_keys[key];
        p.serial = ++proposal_serial;
        if (pending == 0) {
            applyProposal(p);
            applied_keys.insert(key);
        }
        else { // its start position is only known once the earlier moves on this atom / molecule are decided
            unapplied_count++;
            if (pending == 1 && kind == WindowProposal::Kind::ATOM && window_evaluator->conditionals()) {
                // one earlier proposal q on this atom: applied (the trial Space holds its trial position, the
                // accepted Space its start), so both possible starts of this proposal are known
                for (auto q = window.rbegin(); q != window.rend(); ++q) {
                    if (proposalKey(*q) == key) {
                        if (q->applied) {
                            makeConditional(p, *q);
                        }
                        break;
                    }
                }
            }
        }
        pending++;
        window.push_back(std::move(p));
    }

    /** draw proposals (in move order, generator order of performMove) until the queue holds `max_size` */
    void fillWindow(int max_size)
    {
        while (sweep_deferred == nullptr && sweep_remaining > 0 && static_cast<int>(window.size()) < max_size) {
            sweep_remaining--;
            Move* selected = moves->sampleStochasticMove();
            if (selected == nullptr) {
                continue;
            }
            WindowProposal::Kind kind{};
            if (!windowKind(selected, kind) || (!window.empty() && window.front().kind != kind)) {
                sweep_deferred = selected; // another kind of move: runs / opens a new window once the queue is empty
                break;
            }
            enqueue(selected, kind);
        }
    }

    /** The stochastic part of a sweep with runs of `transrot` moves evaluated window by window */
    template <class IdOf> void sweepStochasticWindowed(IdOf&& id_of)
    {
        const int capacity = window_evaluator->capacity();
        sweep_remaining = moves->movesPerSweep();
        sweep_deferred = nullptr;
        sweep_id_of = id_of;
        while (sweep_remaining > 0 || !window.empty() || sweep_deferred != nullptr) {
            fillWindow(capacity);
            if (!window.empty()) {
                // the applied head of the queue (its first proposal always is): proposals further back whose atom /
                // molecule is still undecided wait for their turn
                decideWindow();
            }
            else if (sweep_deferred != nullptr) {
                Move* m = sweep_deferred;
                sweep_deferred = nullptr;
                WindowProposal::Kind kind{};
                if (windowKind(m, kind)) {
                    enqueue(m, kind);
                }
                else {
                    performMove(*m, sweep_id_of(*m));
                }
            }
        }
    }

  public:
    /** src/montecarlo.cpp:220-227 */
    void sweep()
    {
        number_of_sweeps++;
        const auto& all = moves->all();
        auto id_of = [&](Move& m) {
            for (size_t i = 0; i < all.size(); ++i) {
                if (all[i].get() == &m) {
                    return static_cast<int>(i);
                }
            }
            return -1;
        };
        if (window_evaluator) {
            const auto t0 = std::chrono::steady_clock::now();
            sweepStochasticWindowed(id_of);
            window_seconds_total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        else {
            moves->forEachStochasticMove([&](Move& m) { performMove(m, id_of(m)); });
        }
        moves->forEachIntervalMove(number_of_sweeps, [&](Move& m) { performMove(m, id_of(m)); });
    }

    double systemEnergy(std::vector<double>* per_term = nullptr)
    {
        Change change;
        change.everything = true;
        const double u = state.pot->energy(change);
        if (per_term) {
            *per_term = state.pot->latestEnergies();
        }
        return u;
    }
};

/**
 * Widom particle insertion, sequential reference order: activate the first inactive group of the
 * molecule, `ninsert` times {random insertion → ΔU = pot.energy(change) → Σ exp(−ΔU)}, deactivate.
 * src/analysis.cpp:1227-1312 and :807-823. The Hamiltonian is the ACCEPTED one and `updateState`
 * is never called (so an Ewald reciprocal term does not see the ghost — reference behaviour).
 */
class WidomInsertion
{
  protected:
    Space& spc;
    Hamiltonian& pot;
    Random& random; //!< global generator (inserter)
    RandomInserter inserter;
    int molid = -1;
    int number_of_insertions = 0;
    bool absolute_z_coords = false;

  public:
    double sum_exp = 0;           //!< Σ exp(−ΔU)  (Average::value_sum)
    unsigned long count = 0;      //!< number of collected insertions
    std::vector<double> last_du;  //!< ΔU of the most recent sample() call, in insertion order

    WidomInsertion(const Json& j, Space& spc, Hamiltonian& pot, Random& global_random)
        : spc(spc)
        , pot(pot)
        , random(global_random)
    {
        number_of_insertions = j.at("ninsert").integer();
        absolute_z_coords = j.value("absz", false);
        molid = spc.topology->moleculeId(j.at("molecule").string());
        // a fresh default RandomInserter (not the molecule's own), src/analysis.cpp:1293-1310
        if (const auto* d = j.find("dir")) {
            inserter.dir = pointFromJson(*d);
        }
    }
    virtual ~WidomInsertion() = default;

    /** @return false if no ghost group is available */
    bool selectGhostGroup(Change& change) const
    {
        change.clear();
        const auto inactive = spc.findMolecules(molid, Space::Selection::INACTIVE);
        if (inactive.empty()) {
            return false;
        }
        const auto& group = spc.groups[inactive.front()];
        if (group.empty() && group.capacity() > 0) {
            auto& gc = change.groups.emplace_back();
            gc.group_index = inactive.front();
            gc.all = true;
            gc.internal = group.isAtomic();
            return true;
        }
        return false;
    }

    void updateGroup(Group& group, const ParticleVector& particles)
    {
        std::copy(particles.begin(), particles.end(), spc.particles.begin() + group.begin);
        if (absolute_z_coords) {
            for (size_t i = 0; i < group.size(); ++i) {
                auto& p = spc.at(group, i).pos;
                p.z = std::fabs(p.z);
            }
        }
        if (group.isMolecular()) {
            group.mass_center = spc.massCenter(group, -spc.at(group, 0).pos);
        }
    }

    void collect(double energy_change)
    {
        if (-energy_change > pc::max_exp_argument) {
            return; // skipped sample, src/analysis.cpp:809-815
        }
        sum_exp += std::exp(-energy_change);
        count++;
    }

    // One sample event in three steps so that the insertions can be split over ranks: the ghosts
    // never depend on energies, so generating all of them first consumes the generator exactly as the
    // reference's insert → energy → insert → … loop does (src/analysis.cpp:1255-1262).
    struct PreparedEvent
    {
        bool valid = false;
        Change change;
        std::vector<ParticleVector> ghosts;
    } prepared;

    /** generate the ghosts of one sample event; false if no ghost group is available */
    virtual bool prepare()
    {
        prepared = PreparedEvent();
        last_du.clear();
        if (!selectGhostGroup(prepared.change)) {
            return false;
        }
        const auto& mol = spc.topology->molecules[molid];
        for (int cnt = 0; cnt < number_of_insertions; ++cnt) {
            prepared.ghosts.push_back(inserter(spc, mol, random));
        }
        prepared.valid = true;
        return true;
    }

    int preparedInsertions() const { return prepared.valid ? static_cast<int>(prepared.ghosts.size()) : 0; }

    /** ΔU of insertions [first, first + count) of the prepared event */
    virtual void evaluateSlice(int first, int count, double* du)
    {
        checkSlice(first, count);
        auto& group = spc.groups.at(prepared.change.groups.at(0).group_index);
        group.resize(group.capacity());
        for (int b = 0; b < count; ++b) {
            updateGroup(group, prepared.ghosts[first + b]);
            du[b] = pot.energy(prepared.change);
        }
        group.resize(0);
    }

    /** accumulate exp(−ΔU) of ALL insertions of the event, in insertion order */
    void collectAll(const double* du, int n)
    {
        if (!prepared.valid || n != static_cast<int>(prepared.ghosts.size())) {
            throw std::runtime_error("Widom: wrong number of insertion energies");
        }
        for (int b = 0; b < n; ++b) {
            last_du.push_back(du[b]);
            collect(du[b]);
        }
        prepared.valid = false;
    }

    void sample()
    {
        if (!prepare()) {
            return;
        }
        const int n = preparedInsertions();
        std::vector<double> du(static_cast<size_t>(n));
        evaluateSlice(0, n, du.data());
        collectAll(du.data(), n);
    }

  protected:
    void checkSlice(int first, int count) const
    {
        if (!prepared.valid || first < 0 || count < 0 || first + count > static_cast<int>(prepared.ghosts.size())) {
            throw std::runtime_error("Widom slice out of range");
        }
    }

  public:
    double excessChemicalPotential() const { return -std::log(sum_exp / static_cast<double>(count)); }
};

} // namespace fb

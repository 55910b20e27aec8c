// Caller-side Monte Carlo moves: they mutate the trial Space and emit a `Change`; they are not part
// of the replaced energy path but are restated so that a fixed seed gives the reference's proposal
// stream. Mirrors src/move.cpp:23-127 (Move), :202-346 (AtomicTranslateRotate "transrot"),
// :475-524 + src/move.h:640-671 (MoveCollection), :844-968 (ParallelTempering "temper"),
// :984-1033 (VolumeMove "volume"), :1580-1715 (TranslateRotate "moltransrot"),
// src/mpicontroller.cpp:69-82, 94-162, 192-259 (partner policy, particle buffer, exchanges).
//
// RNG draw order follows a GCC build of the reference (function-call arguments and operands of
// overloaded operators are evaluated right-to-left): in `randomUnitVector(slump, dir) * dp * slump()`
// the scalar `slump()` is drawn BEFORE the unit vector (src/move.cpp:229, :1637).
#pragma once
#include "energyterm.hpp"
#include <map>

namespace fb {

/** The reference's three generators: `Move::slump` (static), global `Faunus::random`, `MPI::mpi.random` */
struct Randoms
{
    Random slump;      //!< move::Move::slump, src/move.cpp:21
    Random global;     //!< Faunus::random, src/random.cpp:62
    Random mpi_random; //!< MPI::mpi.random; used by MoveCollection::sample in MPI builds (move.cpp:515-519)
};

class Move
{
  protected:
    Space& spc; //!< trial space
    Randoms& rng;
    int repeat = 1;
    int sweep_interval = 1;
    virtual void _move(Change&) = 0;
    virtual void _accept(Change&) {}
    virtual void _reject(Change&) {}
    virtual void _from_json(const Json&) = 0;
    virtual void _to_json(Json&) const {}

  public:
    std::string name;
    unsigned long number_of_attempted_moves = 0;
    unsigned long number_of_accepted_moves = 0;
    unsigned long number_of_rejected_moves = 0;

    Move(Space& spc, Randoms& rng, std::string name)
        : spc(spc)
        , rng(rng)
        , name(std::move(name))
    {
    }
    virtual ~Move() = default;

    void from_json(const Json& j)
    {
        if (const auto* it = j.find("repeat")) {
            if (it->is_number()) {
                repeat = it->integer();
            }
            else if (it->is_string() && it->string() == "N") {
                repeat = -1;
            }
            else {
                throw std::runtime_error("invalid 'repeat'");
            }
        }
        sweep_interval = j.value("nstep", 1);
        if (sweep_interval > 1) {
            repeat = 0;
        }
        _from_json(j);
        if (repeat < 0) {
            repeat = 0;
        }
    }
    void to_json(Json& j) const
    {
        _to_json(j);
        j["acceptance"] =
            static_cast<double>(number_of_accepted_moves) / static_cast<double>(number_of_attempted_moves);
        j["repeat"] = repeat;
        j["moves"] = static_cast<size_t>(number_of_attempted_moves);
    }
    void move(Change& change)
    {
        number_of_attempted_moves++;
        change.clear();
        _move(change);
    }
    void accept(Change& change)
    {
        number_of_accepted_moves++;
        _accept(change);
    }
    void reject(Change& change)
    {
        number_of_rejected_moves++;
        _reject(change);
    }
    virtual double bias(Change&, double /*old_energy*/, double /*new_energy*/) { return 0.0; }
    void setRepeat(int r) { repeat = r; }
    int getRepeat() const { return repeat; }
    int sweepInterval() const { return sweep_interval; }
    bool isStochastic() const { return repeat != 0; }
};

/** Single-atom translation ("transrot"); src/move.cpp:202-346 */
class AtomicTranslateRotate : public Move
{
    int molid = -1;
    Point directions{1, 1, 1};
    double default_dp = 0;
    double default_dprot = 0;
    Change::GroupChange cdata;
    double latest_displacement_squared = 0;
    double msd_sum = 0;
    unsigned long msd_cnt = 0;

    void _from_json(const Json& j) override
    {
        molid = spc.topology->moleculeId(j.at("molecule").string());
        if (const auto* d = j.find("dir")) {
            directions = pointFromJson(*d);
        }
        if (repeat < 0) {
            const auto mollist = spc.findMolecules(molid, Space::Selection::ALL);
            repeat = static_cast<int>(mollist.size());
            if (repeat > 0) {
                repeat = repeat * static_cast<int>(spc.groups[mollist.front()].size());
            }
        }
        default_dp = j.value("dp", 0.0);
        default_dprot = j.value("dprot", 0.0);
    }
    void _to_json(Json& j) const override
    {
        j["molid"] = molid;
        j["dp"] = default_dp;
        j["msd"] = msd_cnt ? msd_sum / msd_cnt : 0.0;
    }

  public:
    /** The random part of one proposal: everything `_move` draws, nothing that depends on positions */
    struct Draw
    {
        bool valid = false; //!< a particle was picked
        size_t group_index = 0;
        size_t atom_index = 0;
        double dp = 0, dprot = 0;
        double scalar = 0;
        Point unit{0, 0, 0};
    };

    /** consumes the generator exactly as the reference's `_move` does (src/move.cpp:267-293, 328-346) */
    Draw draw()
    {
        Draw d;
        const auto selection =
            spc.topology->molecules[molid].atomic ? Space::Selection::ALL : Space::Selection::ACTIVE;
        const auto mollist = spc.findMolecules(molid, selection);
        if (mollist.empty()) {
            return d;
        }
        d.group_index = mollist[rng.slump.sampleIndex(static_cast<int>(mollist.size()))];
        const auto& group = spc.groups[d.group_index];
        if (group.empty()) {
            return d;
        }
        d.atom_index = static_cast<size_t>(rng.slump.sampleIndex(static_cast<int>(group.size())));
        d.valid = true;
        const auto& atom = spc.traits(spc.at(group, d.atom_index));
        d.dp = atom.dp.value_or(default_dp);
        d.dprot = atom.dprot.value_or(default_dprot);
        if (d.dp > 0.0) { // src/move.cpp:225-240
            d.scalar = rng.slump(); // GCC order: scalar first, then the unit vector
            d.unit = randomUnitVector(rng.slump, directions);
        }
        if (d.dprot > 0.0) { // isotropic particles: draws are consumed, nothing rotates
            (void)randomUnitVector(rng.slump);
            (void)(d.dprot * (rng.slump() - 0.5));
        }
        return d;
    }

    /** displace the picked particle in the trial Space and describe it in `change` */
    void apply(const Draw& d, Change& change)
    {
        latest_displacement_squared = 0.0;
        if (!d.valid) {
            return;
        }
        auto& group = spc.groups[d.group_index];
        cdata.group_index = d.group_index;
        cdata.relative_atom_indices[0] = d.atom_index;
        auto& particle = spc.at(group, d.atom_index);
        if (d.dp > 0.0) {
            const Point old_position = particle.pos;
            particle.pos += d.unit * d.dp * d.scalar;
            spc.geometry.boundary(particle.pos);
            latest_displacement_squared = spc.geometry.sqdist(old_position, particle.pos);
            if (group.isMolecular()) {
                group.mass_center = spc.massCenter(group, -group.mass_center);
            }
        }
        if (d.dp > 0.0 || d.dprot > 0.0) {
            change.groups.push_back(cdata);
        }
    }

    /** windowed path: count the attempt and apply a proposal drawn earlier */
    void moveFromDraw(const Draw& d, Change& change)
    {
        number_of_attempted_moves++;
        change.clear();
        apply(d, change);
    }
    /** where `apply` puts a particle that starts at `start` (the very same arithmetic) */
    Point displaced(const Point& start, const Draw& d) const
    {
        Point pos = start;
        if (d.dp > 0.0) {
            pos += d.unit * d.dp * d.scalar;
            spc.geometry.boundary(pos);
        }
        return pos;
    }
    /** the Change `apply` will describe the proposal with */
    void describe(const Draw& d, Change& change) const
    {
        change.clear();
        auto record = cdata;
        record.group_index = d.group_index;
        record.relative_atom_indices[0] = d.atom_index;
        change.groups.push_back(record);
    }
    bool targetsAtomicGroups() const { return spc.topology->molecules[molid].atomic; }
    double latestDisplacementSquared() const { return latest_displacement_squared; }
    void setLatestDisplacementSquared(double d2) { latest_displacement_squared = d2; }

  private:
    void _move(Change& change) override { apply(draw(), change); }
    void _accept(Change&) override
    {
        msd_sum += latest_displacement_squared;
        msd_cnt++;
    }
    void _reject(Change&) override { msd_cnt++; }

  public:
    AtomicTranslateRotate(Space& spc, Randoms& rng)
        : Move(spc, rng, "transrot")
    {
        repeat = -1;
        cdata.relative_atom_indices.resize(1);
        cdata.internal = true;
    }
};

/** Rigid-body translation + rotation of a molecular group ("moltransrot"); src/move.cpp:1580-1715 */
class TranslateRotate : public Move
{
    int molid = -1;
    Point translational_direction{1, 1, 1};
    Point fixed_rotation_axis{0, 0, 0};
    double translational_displacement = 0;
    double rotational_displacement = 0;

    void _from_json(const Json& j) override
    {
        const auto& mol = spc.topology->molecules.at(spc.topology->moleculeId(j.at("molecule").string()));
        if (mol.atomic) {
            throw std::runtime_error("molecule '" + mol.name + "' cannot be atomic");
        }
        molid = mol.id;
        if (const auto* d = j.find("dir")) {
            translational_direction = pointFromJson(*d);
        }
        translational_displacement = j.at("dp").number();
        rotational_displacement = std::fabs(j.at("dprot").number());
        if (const auto* d = j.find("dirrot")) {
            fixed_rotation_axis = pointFromJson(*d);
            const double n = fixed_rotation_axis.norm();
            if (n > 0) {
                fixed_rotation_axis = fixed_rotation_axis / n;
            }
        }
        if (repeat < 0) {
            repeat = static_cast<int>(spc.findMolecules(molid, Space::Selection::ACTIVE).size());
            if (repeat == 0) {
                repeat = 1;
            }
        }
    }
    void _to_json(Json& j) const override
    {
        j["molid"] = molid;
        j["dp"] = translational_displacement;
        j["dprot"] = rotational_displacement;
    }

  public:
    /** The random part of one proposal: everything `_move` draws (src/move.cpp:1619-1689), no positions */
    struct Draw
    {
        bool valid = false; //!< a non-empty molecule was picked
        size_t group_index = 0;
        bool translate = false;
        double scalar = 0;
        Point unit{0, 0, 0};
        bool rotate = false;
        Point axis{0, 0, 0};
        double angle = 0;
    };

    Draw draw()
    {
        Draw d;
        const auto mollist = spc.findMolecules(molid, Space::Selection::ACTIVE);
        if (mollist.empty()) {
            return d;
        }
        // molecule picked with the GLOBAL generator, src/move.cpp:1622
        d.group_index = mollist[rng.global.sampleIndex(static_cast<int>(mollist.size()))];
        if (spc.groups[d.group_index].empty()) {
            return d;
        }
        d.valid = true;
        if (translational_displacement > 0.0) {
            d.translate = true;
            d.scalar = rng.slump(); // GCC order, see file header
            d.unit = randomUnitVector(rng.slump, translational_direction);
        }
        if (rotational_displacement > pc::epsilon_dbl) {
            const bool fixed = (fixed_rotation_axis.x != 0) || (fixed_rotation_axis.y != 0) ||
                               (fixed_rotation_axis.z != 0);
            d.rotate = true;
            d.axis = fixed ? fixed_rotation_axis : randomUnitVector(rng.slump);
            d.angle = rotational_displacement * (rng.slump() - 0.5);
        }
        return d;
    }

    /** move the picked molecule in the trial Space and describe it in `change` */
    void apply(const Draw& d, Change& change)
    {
        if (!d.valid) {
            return;
        }
        auto& group = spc.groups[d.group_index];
        double displacement_squared = 0.0;
        double angle_squared = 0.0;
        if (d.translate) {
            const Point old_mass_center = group.mass_center;
            spc.translate(group, d.unit * translational_displacement * d.scalar);
            displacement_squared = spc.geometry.sqdist(old_mass_center, group.mass_center);
        }
        if (d.rotate) {
            spc.rotate(group, Quaternion(d.angle, d.axis));
            angle_squared = d.angle * d.angle;
        }
        if (displacement_squared > 0.0 || angle_squared > 0.0) {
            auto& change_data = change.groups.emplace_back();
            change_data.group_index = d.group_index;
            change_data.all = true;
            change_data.internal = false;
        }
        // checkMassCenter, src/move.cpp:1691-1703
        const Point cm = spc.massCenter(group, -group.mass_center);
        if (spc.geometry.sqdist(group.mass_center, cm) > 1e-6) {
            throw std::runtime_error("molecule likely too large for periodic boundaries; increase box size?");
        }
    }

    /** windowed path: count the attempt and apply a proposal drawn earlier */
    void moveFromDraw(const Draw& d, Change& change)
    {
        number_of_attempted_moves++;
        change.clear();
        apply(d, change);
    }

  private:
    void _move(Change& change) override { apply(draw(), change); }

  public:
    TranslateRotate(Space& spc, Randoms& rng)
        : Move(spc, rng, "moltransrot")
    {
        repeat = -1;
    }
};

/** Logarithmic volume displacement ("volume"); src/move.cpp:984-1033 */
class VolumeMove : public Move
{
    double dV = 0;
    VolumeMethod method = VolumeMethod::ISOTROPIC;
    double old_volume = 0;
    double new_volume = 0;

    void _from_json(const Json& j) override
    {
        dV = j.at("dV").number();
        method = volumeMethodFromString(j.value("method", "isotropic"));
    }
    void _to_json(Json& j) const override { j["dV"] = dV; }
    void _move(Change& change) override
    {
        if (dV > 0.0) {
            change.volume_change = true;
            change.everything = true;
            old_volume = spc.geometry.getVolume();
            new_volume = std::exp(std::log(old_volume) + (rng.slump() - 0.5) * dV);
            spc.scaleVolume(new_volume, method);
        }
    }

  public:
    VolumeMove(Space& spc, Randoms& rng)
        : Move(spc, rng, "volume")
    {
        repeat = 1;
    }
};

/**
 * Point-to-point replica communication as used by parallel tempering. The reference uses blocking
 * MPI sendrecv on host memory (src/move.cpp:863, :909; src/mpicontroller.cpp:216, :236, :257).
 * Implementations: in-process (tests, oracle), torch.distributed callbacks (gloo on CPU, NCCL on
 * GPU) and direct NCCL send/recv on device buffers (faunus_b200/csrc/replica_comm.hpp).
 */
class ReplicaComm
{
  public:
    virtual ~ReplicaComm() = default;
    virtual int rank() const = 0;
    virtual int size() const = 0;
    virtual void barrier() = 0;
    /** exchange `n` doubles in place with `partner` (MPI_Sendrecv_replace) */
    virtual void sendrecvReplace(double* data, size_t n, int partner) = 0;
    /** all ranks contribute one double; rank 0 receives all (others may receive garbage) */
    virtual std::vector<double> gather(double value) = 0;
    /**
     * Replace volume, group sizes and ALL particles of `spc` (the trial Space) by the partner's in one exchange and
     * return true — or return false, and the move sends the reference's three messages through sendrecvReplace
     * (exchangeVolume, exchangeGroupSizes, ExchangeParticles; src/mpicontroller.cpp:192-246, src/move.cpp:860-881).
     * A communicator that owns the device mirror of the Space ships the mirror itself, GPU to GPU.
     */
    virtual bool exchangeState(Space& /*spc*/, int /*partner*/, VolumeMethod /*method*/, Change& /*change*/) { return false; }
};

/** Replica exchange ("temper"); src/move.cpp:844-968 */
class ParallelTempering : public Move
{
    ReplicaComm& comm;
    Random slump; //!< private generator for partner selection, src/move.h:577
    std::optional<int> partner;
    VolumeMethod volume_scaling_method = VolumeMethod::ISOTROPIC;
    enum class Format
    {
        XYZ,
        XYZQ,
        XYZQI
    } format = Format::XYZQI;

    void _from_json(const Json& j) override
    {
        const auto f = j.value("format", "xyzqi");
        format = (f == "xyz") ? Format::XYZ : (f == "xyzq" ? Format::XYZQ : Format::XYZQI);
        volume_scaling_method = volumeMethodFromString(j.value("volume_scale", "isotropic"));
    }
    void _to_json(Json& j) const override
    {
        j["replicas"] = comm.size();
        Json ex = Json::object();
        for (const auto& [pair, stat] : acceptance_map) {
            Json s = Json::object();
            s["attempts"] = static_cast<size_t>(stat.second);
            s["acceptance"] = stat.second ? stat.first / static_cast<double>(stat.second) : 0.0;
            ex[std::to_string(pair.first) + " <-> " + std::to_string(pair.second)] = s;
        }
        j["exchange"] = ex;
    }

    /** OddEvenPartner::generate, src/mpicontroller.cpp:69-82 */
    void generatePartner()
    {
        const int rank_increment = static_cast<bool>(slump.range(0, 1)) ? 1 : -1;
        int candidate = (comm.rank() % 2 == 0) ? comm.rank() + rank_increment : comm.rank() - rank_increment;
        if (candidate >= 0 && candidate < comm.size()) {
            partner = candidate;
        }
        else {
            partner = std::nullopt;
        }
    }

    void exchangeState(Change& change)
    {
        if (comm.exchangeState(spc, *partner, volume_scaling_method, change)) {
            return; // volume, group sizes and particles went mirror to mirror; the Space has followed
        }
        // exchangeVolume, src/mpicontroller.cpp:231-246
        const double old_volume = spc.geometry.getVolume();
        double new_volume = old_volume;
        comm.sendrecvReplace(&new_volume, 1, *partner);
        if (new_volume <= pc::epsilon_dbl) {
            throw std::runtime_error("tempering: invalid partner volume");
        }
        if (std::fabs(new_volume - old_volume) > pc::epsilon_dbl) {
            spc.geometry.setVolume(new_volume, volume_scaling_method);
            change.volume_change = true;
        }
        // exchangeGroupSizes, src/move.cpp:860-867
        std::vector<double> sizes;
        for (const auto& g : spc.groups) {
            sizes.push_back(static_cast<double>(g.size()));
        }
        comm.sendrecvReplace(sizes.data(), sizes.size(), *partner);
        for (size_t i = 0; i < sizes.size(); ++i) {
            spc.groups[i].resize(static_cast<size_t>(sizes[i]));
        }
        // ExchangeParticles::replace, src/mpicontroller.cpp:208-219 (all particles incl. inactive)
        const size_t packet = (format == Format::XYZ) ? 3 : (format == Format::XYZQ ? 4 : 5);
        std::vector<double> buffer(packet * spc.particles.size());
        size_t k = 0;
        for (const auto& p : spc.particles) {
            buffer[k++] = p.pos.x;
            buffer[k++] = p.pos.y;
            buffer[k++] = p.pos.z;
            if (format != Format::XYZ) {
                buffer[k++] = p.charge;
            }
            if (format == Format::XYZQI) {
                buffer[k++] = static_cast<double>(p.id);
            }
        }
        comm.sendrecvReplace(buffer.data(), buffer.size(), *partner);
        k = 0;
        for (auto& p : spc.particles) {
            p.pos.x = buffer[k++];
            p.pos.y = buffer[k++];
            p.pos.z = buffer[k++];
            if (format != Format::XYZ) {
                p.charge = buffer[k++];
            }
            if (format == Format::XYZQI) {
                p.id = static_cast<int>(buffer[k++]);
            }
        }
        spc.updateMassCenters(); // spc.updateParticles(...), src/move.cpp:879
        change.everything = true;
    }

    void _move(Change& change) override
    {
        comm.barrier();
        // checkRandomEngineState, src/mpicontroller.cpp:253-259
        const auto numbers = comm.gather(slump());
        if (comm.rank() == 0 &&
            std::adjacent_find(numbers.begin(), numbers.end(), std::not_equal_to<>()) != numbers.end()) {
            throw std::runtime_error("Random numbers out of sync across replicas");
        }
        generatePartner();
        if (partner.has_value()) {
            exchangeState(change);
        }
    }
    std::pair<int, int> pairKey() const
    {
        return {std::min(comm.rank(), *partner), std::max(comm.rank(), *partner)};
    }
    void _accept(Change&) override
    {
        auto& s = acceptance_map[pairKey()];
        s.first += 1.0;
        s.second++;
    }
    void _reject(Change&) override { acceptance_map[pairKey()].second++; }

  public:
    std::map<std::pair<int, int>, std::pair<double, unsigned long>> acceptance_map;

    ParallelTempering(Space& spc, Randoms& rng, ReplicaComm& comm)
        : Move(spc, rng, "temper")
        , comm(comm)
    {
        if (comm.size() < 2) {
            throw std::runtime_error("temper requires two or more replicas");
        }
        repeat = 0; // zero-weight move, run at the end of each sweep (src/move.cpp:445)
    }

    /** partner's energy change, exchanged as one double; src/move.cpp:905-923 */
    double bias(Change&, double old_energy, double new_energy) override
    {
        double du = new_energy - old_energy;
        comm.sendrecvReplace(&du, 1, *partner);
        return du;
    }
};

/** Weighted random selection of moves; src/move.cpp:475-524, src/move.h:640-671 */
class MoveCollection
{
    std::vector<std::shared_ptr<Move>> moves;
    std::vector<double> repeats;
    std::discrete_distribution<unsigned int> distribution;
    unsigned int number_of_moves_per_sweep = 0;
    Randoms& rng;
    bool use_mpi_random = false;

    Move* sample()
    {
        auto& engine = use_mpi_random ? rng.mpi_random.engine : rng.slump.engine;
        if (!moves.empty()) {
            return moves[distribution(engine)].get();
        }
        return nullptr;
    }

  public:
    MoveCollection(const Json& list_of_moves, Space& trial_spc, Randoms& rng, ReplicaComm* comm)
        : rng(rng)
        , use_mpi_random(comm != nullptr)
    {
        for (const auto& j : list_of_moves.items()) {
            const auto& [name, params] = j.single();
            std::shared_ptr<Move> move;
            if (name == "transrot") {
                move = std::make_shared<AtomicTranslateRotate>(trial_spc, rng);
            }
            else if (name == "moltransrot") {
                move = std::make_shared<TranslateRotate>(trial_spc, rng);
            }
            else if (name == "volume") {
                move = std::make_shared<VolumeMove>(trial_spc, rng);
            }
            else if (name == "temper") {
                if (!comm) {
                    throw std::runtime_error("temper requires a replica communicator");
                }
                move = std::make_shared<ParallelTempering>(trial_spc, rng, *comm);
            }
            else {
                throw std::runtime_error("move '" + name + "' is outside the hot-path scope");
            }
            move->from_json(params);
            if (name == "temper") {
                move->setRepeat(0);
            }
            moves.push_back(move);
            repeats.push_back(static_cast<double>(move->getRepeat()));
            distribution = std::discrete_distribution<unsigned int>(repeats.begin(), repeats.end());
            number_of_moves_per_sweep =
                static_cast<unsigned int>(std::accumulate(repeats.begin(), repeats.end(), 0.0));
        }
    }

    /**
     * The reference builds `iota | transform(sample) | filter(is_stochastic) | indirect`
     * (src/move.h:640-650); with a lazy transform under a filter the sampling functor runs twice per
     * element (once for the predicate, once when dereferenced), so two draws are consumed per
     * performed move and the second one selects the move. Restated literally.
     */
    template <class F> void forEachStochasticMove(F&& perform)
    {
        for (unsigned int i = 0; i < number_of_moves_per_sweep; ++i) {
            Move* probe = sample();
            if (probe != nullptr && probe->isStochastic()) {
                Move* selected = sample();
                perform(*selected);
            }
        }
    }
    /** one iteration of forEachStochasticMove: the move to perform, or nullptr if the probe was not stochastic */
    Move* sampleStochasticMove()
    {
        Move* probe = sample();
        if (probe != nullptr && probe->isStochastic()) {
            return sample();
        }
        return nullptr;
    }
    template <class F> void forEachIntervalMove(unsigned int sweep_number, F&& perform)
    {
        for (auto& m : moves) {
            if (!m->isStochastic() && (sweep_number % m->sweepInterval() == 0)) {
                perform(*m);
            }
        }
    }
    const std::vector<std::shared_ptr<Move>>& all() const { return moves; }
    unsigned int movesPerSweep() const { return number_of_moves_per_sweep; }
};

} // namespace fb

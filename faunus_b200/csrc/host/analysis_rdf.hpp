// Radial distribution functions g(r) of atoms ("atomrdf") and of molecular mass centres ("molrdf"), restated caller side:
// src/analysis.cpp:712-722 (PairFunction::_from_json), :1570-1579 (AtomRDF::_sample), :724-757 (normalisation),
// src/aux/equidistant_table.h:32-40 (binning: floor(r / dr), xmin = 0, the table grows on demand).
// The pair loop itself (AtomRDF::sampleIdentical / sampleDifferent / sampleDistance, :1556-1600) is the part a
// subclass supplies: the CPU restatement lives in oracle/, the B200 build counts on the device (fb_atom_rdf).
#pragma once
#include "space.hpp"

namespace fb {

class AtomRDF
{
  protected:
    const Space& spc;
    bool molecular = false; //!< "molrdf": mass centres of molecular groups, distance = sqrt(sqdist) (:1642-1648)
    int id1 = 0, id2 = 0;
    double dr = 0.1;
    int dimensions = 3;
    int slicedir[3] = {0, 0, 0};
    double thickness = 0;
    std::vector<unsigned long long> histogram; //!< pair counts per bin (exact; the reference counts in doubles)
    double volume_sum = 0;                     //!< Average<double> mean_volume
    unsigned long volume_count = 0;

    /**
     * Add the pair distances of the current configuration to `histogram` (grow it as needed); share `shard` of
     * `n_shards` of the pairs when the sample is split over ranks (any split: the counts are integers)
     */
    virtual void count(int shard, int n_shards) = 0;

  public:
    AtomRDF(const Json& j, const Space& spc)
        : spc(spc)
    {
        molecular = j.value("type", "atomrdf") == "molrdf";
        if (molecular) { // MoleculeRDF::MoleculeRDF, src/analysis.cpp:1650-1658
            id1 = spc.topology->moleculeId(j.at("name1").string());
            id2 = spc.topology->moleculeId(j.at("name2").string());
            if (spc.topology->molecules.at(id1).atomic || spc.topology->molecules.at(id2).atomic) {
                throw std::runtime_error("molrdf: molecular groups required");
            }
        }
        else {
            id1 = spc.topology->atomId(j.at("name1").string());
            id2 = spc.topology->atomId(j.at("name2").string());
        }
        dr = j.value("dr", 0.1);
        dimensions = static_cast<int>(j.value("dim", 3.0));
        if (const auto* s = j.find("slicedir")) {
            const Point p = pointFromJson(*s);
            slicedir[0] = static_cast<int>(p.x);
            slicedir[1] = static_cast<int>(p.y);
            slicedir[2] = static_cast<int>(p.z);
        }
        thickness = j.value("thickness", 0.0);
        if (!(dr > 0.0)) {
            throw std::runtime_error("atomrdf: dr must be positive");
        }
        if (dimensions != 3) {
            throw std::runtime_error("atomrdf: only dim = 3 is normalised here");
        }
    }
    virtual ~AtomRDF() = default;

    /** bin of a distance: Equidistant2DTable::to_bin with xmin = 0 */
    int bin(double r) const { return static_cast<int>(std::floor(r * (1.0 / dr))); }

    /** bins that hold every possible minimum-image distance of the current cell */
    int binsForCell() const
    {
        const Point& len = spc.geometry.getLength();
        double r2 = 0.0;
        for (int i = 0; i < 3; ++i) {
            const double extent = spc.geometry.isPeriodic(i) ? 0.5 * len[i] : len[i];
            r2 += extent * extent;
        }
        const double r_max = spc.geometry.type == Geometry::Type::SPHERE ? 2.0 * spc.geometry.getRadius() : std::sqrt(r2);
        return bin(r_max) + 2;
    }

    /** AtomRDF::_sample */
    void sample()
    {
        volume_sum += spc.geometry.getVolume();
        volume_count++;
        count(0, 1);
    }

    /** this rank's share of one sample; the histograms of the ranks add up to the unsharded one */
    void sampleShard(int shard, int n_shards)
    {
        if (n_shards < 1 || shard < 0 || shard >= n_shards) {
            throw std::runtime_error("atomrdf: bad shard");
        }
        volume_sum += spc.geometry.getVolume();
        volume_count++;
        count(shard, n_shards);
    }

    size_t size() const { return histogram.size(); }
    double distance(size_t i) const { return static_cast<double>(i) / (1.0 / dr); } //!< from_bin
    unsigned long long pairs(size_t i) const { return histogram[i]; }
    int firstType() const { return id1; }
    int secondType() const { return id2; }
    double resolution() const { return dr; }

    /** g(r) as PairFunction::_to_disk prints it: N ⟨V⟩ / (4π r² dr · Σ N); 0 where the volume element vanishes */
    double g(size_t i) const { return g(i, total()); }

    /** Σ N over the bins (callers that want every g(r) take it once: the sum inside g(i) made a table of n bins O(n²)) */
    double total() const
    {
        double sum = 0.0;
        for (const auto c : histogram) {
            sum += static_cast<double>(c);
        }
        return sum;
    }

    double g(size_t i, double total) const
    {
        const double r = distance(i);
        const double volume_at_r = 4.0 * pc::pi * r * r * dr;
        if (!(volume_at_r > 0.0) || total == 0.0 || volume_count == 0) {
            return 0.0;
        }
        return static_cast<double>(histogram[i]) * (volume_sum / static_cast<double>(volume_count)) / (volume_at_r * total);
    }
};

} // namespace fb

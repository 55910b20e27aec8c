// ORACLE (test infrastructure only): the pair loop of AtomRDF on the CPU, restating
// src/analysis.cpp:1556-1568 (sampleDistance), :1581-1600 (sampleIdentical / sampleDifferent; Space::findAtoms
// = the active particles of that type in storage order) with Geometry::vdist (src/geometry.h:429-458).
#pragma once
#include "../faunus_b200/csrc/host/analysis_rdf.hpp"

namespace oracle {

class AtomRDFCpu : public fb::AtomRDF
{
    void sampleDistance(const fb::Point& a, const fb::Point& b)
    {
        const fb::Point d = spc.geometry.vdist(a, b);
        double r;
        if (slicedir[0] + slicedir[1] + slicedir[2] > 0) {
            const fb::Point in_plane{slicedir[0] ? 0.0 : d.x, slicedir[1] ? 0.0 : d.y, slicedir[2] ? 0.0 : d.z};
            if (!(in_plane.norm() < thickness)) {
                return;
            }
            const fb::Point along{slicedir[0] ? d.x : 0.0, slicedir[1] ? d.y : 0.0, slicedir[2] ? d.z : 0.0};
            r = along.norm();
        }
        else {
            r = d.norm();
        }
        const auto i = static_cast<size_t>(bin(r));
        if (i >= histogram.size()) {
            histogram.resize(i + 1, 0ull);
        }
        histogram[i]++;
    }

    std::vector<fb::Point> findAtoms(int id) const
    {
        std::vector<fb::Point> found;
        for (const auto& g : spc.groups) {
            for (size_t i = 0; i < g.size(); ++i) {
                const auto& p = spc.at(g, i);
                if (p.id == id) {
                    found.push_back(p.pos);
                }
            }
        }
        return found;
    }

    /** MoleculeRDF::sampleDistance, src/analysis.cpp:1642-1648 */
    void sampleMassCenters(const fb::Point& a, const fb::Point& b)
    {
        const auto i = static_cast<size_t>(bin(std::sqrt(spc.geometry.sqdist(a, b))));
        if (i >= histogram.size()) {
            histogram.resize(i + 1, 0ull);
        }
        histogram[i]++;
    }

    std::vector<fb::Point> findMassCenters(int molid) const
    {
        std::vector<fb::Point> found;
        for (const auto g : spc.findMolecules(molid, fb::Space::Selection::ACTIVE)) {
            found.push_back(spc.groups[g].mass_center);
        }
        return found;
    }

    void countMolecules(int shard, int n_shards)
    {
        const auto stride = static_cast<size_t>(n_shards);
        const auto first = findMassCenters(id1);
        const auto second = id1 == id2 ? first : findMassCenters(id2);
        for (size_t i = static_cast<size_t>(shard); i < first.size(); i += stride) {
            for (size_t j = (id1 == id2 ? i + 1 : 0); j < second.size(); ++j) {
                sampleMassCenters(first[i], second[j]);
            }
        }
    }

    void count(int shard, int n_shards) override
    {
        if (molecular) {
            countMolecules(shard, n_shards);
            return;
        }
        const auto stride = static_cast<size_t>(n_shards);
        if (id1 == id2) {
            const auto atoms = findAtoms(id1);
            for (size_t i = static_cast<size_t>(shard); i < atoms.size(); i += stride) {
                for (size_t j = i + 1; j < atoms.size(); ++j) {
                    sampleDistance(atoms[i], atoms[j]);
                }
            }
        }
        else {
            const auto first = findAtoms(id1), second = findAtoms(id2);
            for (size_t i = static_cast<size_t>(shard); i < first.size(); i += stride) {
                for (const auto& b : second) {
                    sampleDistance(first[i], b);
                }
            }
        }
    }

  public:
    using fb::AtomRDF::AtomRDF;
};

} // namespace oracle

// Gene sets and cell sets as the hot path sees them: sorted id vectors in memory-mapped files named
// GeneSet-<name>-GlobalIds / -LocalIds and CellSet-<name> (reference src/GeneSet.cpp:9-21,
// src/CellSets.cpp:75).
#pragma once
#include <algorithm>
#include <string>

#include "Ids.hpp"
#include "MemoryMapped.hpp"

namespace ChanZuckerberg {
namespace ExpressionMatrix2 {

using CellSet = MemoryMapped::Vector<CellId>;

class GeneSet {
public:
    void createNew(const std::string& name)
    {
        globalIds_.createNew(name + "-GlobalIds", 0);
        localIds_.createNew(name + "-LocalIds", 0);
    }
    void accessExisting(const std::string& name, bool allowReadOnly)
    {
        globalIds_.accessExistingReadWrite(name + "-GlobalIds", allowReadOnly);
        localIds_.accessExistingReadWrite(name + "-LocalIds", allowReadOnly);
        if (!std::is_sorted(globalIds_.begin(), globalIds_.end())) {
            if (!globalIds_.isOpenWithWriteAccess() || !localIds_.isOpenWithWriteAccess())
                throw std::runtime_error("Gene set " + name + " is not sorted and accessed read-only.");
            sort();
        }
    }
    // The caller guarantees a gene is added once.
    void addGene(GeneId globalId)
    {
        if (globalId >= localIds_.size()) {
            const size_t old = localIds_.size();
            localIds_.resize(size_t(globalId) + 1);
            std::fill(localIds_.begin() + old, localIds_.end(), invalidGeneId);
        }
        localIds_[globalId] = GeneId(globalIds_.size());
        globalIds_.push_back(globalId);
    }
    void sort()
    {
        if (std::is_sorted(globalIds_.begin(), globalIds_.end())) return;
        std::sort(globalIds_.begin(), globalIds_.end());
        std::fill(localIds_.begin(), localIds_.end(), invalidGeneId);
        for (GeneId l = 0; l < globalIds_.size(); l++) localIds_[globalIds_[l]] = l;
    }
    GeneId size() const { return GeneId(globalIds_.size()); }
    const GeneId* begin() const { return globalIds_.begin(); }
    const GeneId* end() const { return globalIds_.end(); }
    GeneId getGlobalGeneId(GeneId local) const { return globalIds_[local]; }
    GeneId getLocalGeneId(GeneId global) const { return global < localIds_.size() ? localIds_[global] : invalidGeneId; }
    bool contains(GeneId global) const { return getLocalGeneId(global) != invalidGeneId; }
    const MemoryMapped::Vector<GeneId>& genes() const { return globalIds_; }
    // local id of every global gene (invalidGeneId if absent): the GeneSet-<name>-LocalIds file
    const MemoryMapped::Vector<GeneId>& localIds() const { return localIds_; }
    bool isIdentity() const { return size() == 0 || (globalIds_[0] == 0 && globalIds_[size() - 1] == size() - 1); }
    void close()
    {
        globalIds_.close();
        localIds_.close();
    }
    void remove()
    {
        globalIds_.remove();
        localIds_.remove();
    }

private:
    MemoryMapped::Vector<GeneId> globalIds_;
    MemoryMapped::Vector<GeneId> localIds_;
};

}  // namespace ExpressionMatrix2
}  // namespace ChanZuckerberg

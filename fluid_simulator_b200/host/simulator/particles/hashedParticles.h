// B200 facade of particles/hashedParticles.h:17-179: particle store with the reference's accessors.  The device keeps
// fp32 SoA particles in cell-binned order; this class keeps an AoS fp64 HOST MIRROR that is downloaded on first access
// after a step and uploaded before the next step if it was handed out by non-const reference (getParticleAt, forEach).
// Indices are therefore not stable across steps (SURVEY.md §8b "Ownership").
#pragma once
#include <functional>
#include <memory>
#include <vector>

#include <glm/glm.hpp>

#include "../b200_backend.h"
#include "particle.h"

namespace genericfsim::simulator { class Simulator; }

namespace genericfsim::particles {

class HashedParticles {
public:
    // random fill of the upper-right half of the domain with libc rand(), hashedParticles.cpp:23-42
    HashedParticles(int num, double r, glm::dvec3 dimensions, glm::dvec3 cellD, bool zConst, double z);

    void forEach(bool parallel, std::function<void(Particle&, int)>&& lambda);
    void setParticleNum(int num);            // hashedParticles.cpp:132-151 (grow = random fill of the whole domain)
    Particle& getParticleAt(int idx);
    int getParticleNum() const;
    double getParticleR() const;
    void setParticleR(double r);
    void updateGridParams(const glm::dvec3& cellD, const glm::dvec3& dimensions);  // clamp into the new grid, :340-353
    void addParticles(std::vector<Particle>&& particles);
    void removeParticles(std::vector<int>&& particleIds);  // stable compaction, :158-172
    // push-apart (hashedParticles.cpp:64-107, 195-244) is SURVEY §8f "next #1": accepted, no-ops
    void updateParticleIntersectionHash(bool) {}
    void pushParticlesApart(bool) {}

public:
    const double z;
    const bool zConst;

private:
    friend class genericfsim::simulator::Simulator;
    std::vector<Particle> particles;  // host mirror
    glm::dvec3 dimensions, cellD, cellDInv;
    double r;
    std::shared_ptr<genericfsim::b200::Backend> backend;
    bool hostValid = true;     // mirror holds the current particles
    bool deviceValid = false;  // device holds the current particles
    bool radiusDirty = false;
    void ensureHost();
    void touchHost();          // mirror is about to be modified
    void flushToDevice();      // called by Simulator before a step
};

}  // namespace genericfsim::particles

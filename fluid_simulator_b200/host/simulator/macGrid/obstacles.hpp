// B200 facade of macGrid/obstacles.hpp:7-65: the obstacle hierarchy the caller fills into Simulator::obstacles.
// Field names, constructors and clone() match the reference so Application code compiles unchanged; the facade
// flattens them to FsimObstacle PODs at every simulate() (the manager re-clones them each step anyway,
// manager/simulationManager.cpp:202-207).
#pragma once
#include <glm/glm.hpp>

namespace genericfsim::obstacle {

struct Obstacle {
    glm::dvec3 pos;
    glm::dvec3 prevPos;
    glm::dvec3 speed = glm::dvec3(0, 0, 0);

    explicit Obstacle(const glm::dvec3& p = glm::dvec3(0, 0, 0)) : pos(p), prevPos(p) {}
    virtual ~Obstacle() = default;

    void setNewPos(const glm::dvec3& p) { prevPos = pos; pos = p; }
    void calculateSpeed(double dt) { speed = (pos - prevPos) / dt; }
    virtual Obstacle* clone() = 0;
};

struct RectengularObstacle : public Obstacle {
    const glm::dvec3 size;
    RectengularObstacle(const glm::dvec3& sz, const glm::dvec3& p = glm::dvec3(0, 0, 0)) : Obstacle(p), size(sz) {}
    Obstacle* clone() override { return new RectengularObstacle(*this); }
};

struct SphericalObstacle : public Obstacle {
    const double r;
    SphericalObstacle(double radius, const glm::dvec3& p = glm::dvec3(0, 0, 0)) : Obstacle(p), r(radius) {}
    Obstacle* clone() override { return new SphericalObstacle(*this); }
};

struct SphericalParticleSource : public SphericalObstacle {
    const double particleSpawnRate;
    const double particleSpawnSpeed;
    double lastSpawnFraction = 0;
    SphericalParticleSource(double radius, double rate, double spawnSpeed, const glm::dvec3& p = glm::dvec3(0, 0, 0))
        : SphericalObstacle(radius, p), particleSpawnRate(rate), particleSpawnSpeed(spawnSpeed) {}
    Obstacle* clone() override { return new SphericalParticleSource(*this); }
};

struct SphericalParticleSink : public SphericalObstacle {
    SphericalParticleSink(double radius, const glm::dvec3& p = glm::dvec3(0, 0, 0)) : SphericalObstacle(radius, p) {}
    Obstacle* clone() override { return new SphericalParticleSink(*this); }
};

}  // namespace genericfsim::obstacle

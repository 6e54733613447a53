// B200 facade of macGrid/macGrid.h:19-194: same public constants, solver parameters and cell accessors as the reference's
// MacGrid, backed by the device grid of an fsim handle.  cell()/getCellsAround()/getFacesAround() serve a lazily
// refreshed HOST MIRROR in the reference's layout (x-major, z fastest, macGrid.h:66-93); the hot path never touches it.
#pragma once
#include <array>
#include <atomic>
#include <functional>
#include <mutex>
#include <memory>
#include <utility>
#include <vector>

#include <glm/glm.hpp>

#include "../b200_backend.h"
#include "macGridCell.h"
#include "obstacles.hpp"

namespace genericfsim::simulator { class Simulator; }

namespace genericfsim::macgrid {

class MacGrid {
public:
    // MacGrid(targetDimensions, resolution, twoD), macGrid.cpp:9-17: cellD = 1/resolution (z: dims.z/3 in 2D),
    // gridSize = trunc(dims/cellD) (z = 3 in 2D); starts all AIR inside a SOLID border shell
    MacGrid(glm::dvec3 targetDimensions, double resolution, bool twoD);
    virtual ~MacGrid() = default;
    MacGrid(const MacGrid&) = delete;

    std::array<std::array<MacGridCell::FaceRef, 8>, 3> getFacesAround(const glm::dvec3& pos);
    MacGridCell& getCellAt(const glm::dvec3& pos);
    std::array<MacGridCellRef, 8> getCellsAround(const glm::dvec3& pos);

    template <int axis = 0, int offset = 0>
    inline MacGridCell& cell(const glm::ivec3& p) {
        refreshMirror();
        if constexpr (axis == 0) return rawCells[(p.x + offset) * yzMultiplier + p.y * gridSize.z + p.z];
        else if constexpr (axis == 1) return rawCells[p.x * yzMultiplier + (p.y + offset) * gridSize.z + p.z];
        else return rawCells[p.x * yzMultiplier + p.y * gridSize.z + p.z + offset];
    }
    inline bool isPosValid(const glm::ivec3& p) const {
        return p.x >= 0 && p.x < gridSize.x && p.y >= 0 && p.y < gridSize.y && p.z >= 0 && p.z < gridSize.z;
    }
    inline MacGridCell& cell(int x, int y, int z) {
        refreshMirror();
        return rawCells[x * yzMultiplier + y * gridSize.z + z];
    }
    void forEachCell(bool parallel, bool includeBorders, std::function<void(glm::ivec3 pos, MacGridCell&)>&& lambda);

    // stage-level entry points of the reference that map onto device stages (macGrid.cpp:204-219, 294-336)
    void postP2GUpdate(bool parallel, double gravityIncrement);
    void extrapolateVelocities(bool parallel);
    std::pair<glm::dvec3, glm::dvec3> getMinMaxRect(const glm::dvec3& pos, const glm::dvec3& size) const;

    // PCG projection on the device; returns the iteration count like the reference (bridsonSolverGrid.cpp:244-293)
    virtual int solveIncompressibility(bool parallel, double dt) = 0;

public:
    const glm::dvec3 cellD;
    const glm::dvec3 cellDInv;
    const glm::ivec3 gridSize;
    const glm::dvec3 dimensions;

    bool isTopOfContainerSolid = false;
    double pressureK = 2.0, averagePressure = 2.0;
    int incompressibilityMaxIterationCount = 80;
    bool pressureEnabled = true;
    double fluidDensity = 1.0;
    double residualTolerance = 1e-6;

    const bool twoD;

    // --- B200 additions ---------------------------------------------------------------------------------------
    const glm::dvec3 targetDimensions;  // ctor arguments, needed to create the device grid
    const double resolution;

protected:
    friend class genericfsim::simulator::Simulator;
    const int yzMultiplier;
    const int cellCount;
    std::vector<MacGridCell> rawCells;  // host mirror
    std::shared_ptr<genericfsim::b200::Backend> backend;
    std::atomic<long long> mirrorStep{-1};  // step the mirror was downloaded for
    std::mutex mirrorMutex;                 // cell() is called from OpenMP threads by the manager's gfx loop
    void refreshMirror();
    void initMirror();
};

}  // namespace genericfsim::macgrid

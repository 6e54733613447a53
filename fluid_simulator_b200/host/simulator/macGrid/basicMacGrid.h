// macGrid/basicMacGrid.h:7-16 is the GUI's alternative red-black SOR solver; it is outside the path this backend
// accelerates.  The type exists so that SimulationManager compiles unchanged; selecting it throws.
#pragma once
#include <stdexcept>
#include "macGrid.h"

namespace genericfsim::macgrid {

class BasicMacGrid : public MacGrid {
public:
    BasicMacGrid(glm::dvec3 targetDimensions, double resolution, bool twoD) : MacGrid(targetDimensions, resolution, twoD) {
        throw std::runtime_error("BasicMacGrid (red-black SOR) is not provided by the B200 backend; use GridSolverType::BRIDSON");
    }
    int solveIncompressibility(bool, double) override { return 0; }
};

}  // namespace genericfsim::macgrid

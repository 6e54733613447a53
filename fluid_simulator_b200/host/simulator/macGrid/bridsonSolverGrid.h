// B200 facade of macGrid/bridsonSolverGrid.h:12-41: the PCG projection grid.  Same constructor signature (the second
// argument is the RESOLUTION and a float, bridsonSolverGrid.cpp:8); the solve runs on the device (pcg.cu, mg.cu).
#pragma once
#include "macGrid.h"

namespace genericfsim::macgrid {

class BridsonSolverGrid : public MacGrid {
public:
    BridsonSolverGrid(const glm::dvec3& dimensions, float cellD, bool twoD, double fluidDensity = 1.0);
    int solveIncompressibility(bool parallel, double dt) override;
};

}  // namespace genericfsim::macgrid

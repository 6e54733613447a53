// B200 facade of macGrid/basicMacGrid.h:7-16: the GUI's alternative solver (red-black Gauss-Seidel with SOR 1.98 on the
// face velocities, a fixed number of sweeps).  Selecting it makes Simulator::simulate run FSIM_SOLVER_BASIC on the device
// (grid.cu basic_sor_kernel) instead of the PCG projection.
#pragma once
#include "macGrid.h"

namespace genericfsim::macgrid {

class BasicMacGrid : public MacGrid {
public:
    BasicMacGrid(glm::dvec3 targetDimensions, double resolution, bool twoD) : MacGrid(targetDimensions, resolution, twoD) {}
    int solveIncompressibility(bool parallel, double dt) override;
};

}  // namespace genericfsim::macgrid

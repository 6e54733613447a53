// TEST INFRASTRUCTURE ONLY. Declaration force-included (-include) when oracle/build_ref.sh
// compiles the reference's bridsonSolverGrid.cpp with the one-line pressure-export patch.
#pragma once
#include <vector>
extern std::vector<double> g_ref_last_pressure;

// B200 facade: host-side helpers with the reference's names (util/interpolation.h:7-35); SimulationManager's gfx loop
// calls trilinearInterpoll on the host mirror.  The device kernels have their own fp32 versions (p2g.cu, g2p.cu).
#pragma once
#include <cmath>
#include <glm/glm.hpp>

inline double trilinearInterpoll(const glm::dvec3& center, const glm::dvec3& pos, const glm::dvec3 invCellD) {
    const glm::dvec3 d = (pos - center) * invCellD;
    return (1.0 - std::fabs(d.x)) * (1.0 - std::fabs(d.y)) * (1.0 - std::fabs(d.z));
}

inline glm::dvec3 trilinearInterpollGradient(const glm::dvec3& center, const glm::dvec3& pos, const glm::dvec3 invCellD) {
    const glm::dvec3 d = (pos - center) * invCellD;
    const double ax = 1.0 - std::fabs(d.x), ay = 1.0 - std::fabs(d.y), az = 1.0 - std::fabs(d.z);
    const double sx = d.x > 0.0 ? -1.0 : 1.0, sy = d.y > 0.0 ? -1.0 : 1.0, sz = d.z > 0.0 ? -1.0 : 1.0;
    return glm::dvec3(sx * ay * az, sy * ax * az, sz * ax * ay) * invCellD;
}

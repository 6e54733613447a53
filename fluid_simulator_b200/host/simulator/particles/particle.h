// B200 facade: same type as the reference's particles/particle.h:8-12 (AoS fp64, 15 doubles) -- this is the host-side
// exchange format of the C ABI (fsim_upload_particles / fsim_download_particles).
#pragma once
#include <glm/glm.hpp>

namespace genericfsim::particles {

struct Particle {
    glm::dvec3 pos;
    glm::dvec3 v;
    glm::dvec3 c[3];
};
static_assert(sizeof(Particle) == 15 * sizeof(double), "Particle must stay 15 packed doubles (fsim ABI host layout)");

}  // namespace genericfsim::particles

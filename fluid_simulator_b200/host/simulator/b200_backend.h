// Shared state of the facade classes: one fsim handle (include/fsim.h) per Simulator, plus sync flags between the
// device state and the host mirrors that the reference's by-reference accessors (getParticleAt, cell()) need.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>

#include "../../../include/fsim.h"

namespace genericfsim::b200 {

struct Backend {
    fsim_t* h = nullptr;
    long long stepCount = 0;  // bumped by every simulate(); host mirrors compare against it
    Backend() = default;
    Backend(const Backend&) = delete;
    Backend& operator=(const Backend&) = delete;
    ~Backend() { if (h) fsim_destroy(h); }
};

// the C ABI reports status codes; the reference's classes report by exception (main.cpp:26-40 catches std::exception)
inline void check(int rc, const fsim_t* h, const char* what) {
    if (rc != FSIM_OK) throw std::runtime_error(std::string(what) + ": " + fsim_last_error(h));
}

}  // namespace genericfsim::b200

// slab_host.cpp -- a C++ host driving the z-slab C ABI (include/fsim.h "z-slab decomposition"): N slab handles of one MacGrid,
// one std::thread per rank (here all ranks share cuda:0; with one process per GPU the same calls run with the exports
// travelling over MPI / a socket instead of a vector).  Checks the sharded run against a single handle on the same scene:
// equal PCG iteration counts, particle count conserved, sum |v2| and sum p within 1e-5.
//   usage: slab_host [ranks=2] [n=16] [steps=3]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "fsim.h"

#define CK(h, call)                                                                       \
    do {                                                                                  \
        const int rc__ = (call);                                                          \
        if (rc__) { std::fprintf(stderr, "%s -> %d: %s\n", #call, rc__, fsim_last_error(h)); std::exit(1); } \
    } while (0)

static FsimParams params() {
    FsimParams p;
    std::memset(&p, 0, sizeof(p));
    p.transfer_type = FSIM_TRANSFER_FLIP; p.flip_ratio = 0.95f; p.gravity = -9.81f * 4; p.gravity_enabled = 1;
    p.pressure_enabled = 1; p.max_iterations = 2000; p.pressure_k = 1.0; p.average_pressure = 8.0; p.fluid_density = 1.0;
    p.residual_tolerance = 1e-9;
    return p;
}

// dam block x in [1, n/2), y, z in [1, n-1): 8 particles per cell on the 0.25 / 0.75 sub-cell sites (reference `Particle` layout)
static std::vector<double> dam(int n) {
    std::vector<double> a;
    for (int x = 1; x < n / 2; x++)
        for (int y = 1; y < n - 1; y++)
            for (int z = 1; z < n - 1; z++)
                for (int s = 0; s < 8; s++) {
                    double q[15] = {x + (s & 1 ? 0.75 : 0.25), y + (s & 2 ? 0.75 : 0.25), z + (s & 4 ? 0.75 : 0.25)};
                    a.insert(a.end(), q, q + 15);
                }
    return a;
}

struct Sums { double v2 = 0, p = 0; };
static Sums owned_sums(fsim_t* h) {  // over the planes the handle owns (ghost planes belong to the neighbours)
    FsimGridInfo gi; FsimSlabInfo si;
    CK(h, fsim_get_grid_info(h, &gi)); CK(h, fsim_get_slab_info(h, &si));
    const int64_t nc = (int64_t)gi.grid_size[0] * gi.grid_size[1] * si.gz_local;
    std::vector<double> v2(nc * 3), p(nc);
    CK(h, fsim_download_grid(h, FSIM_FIELD_V2, v2.data(), (int64_t)v2.size() * 8));
    CK(h, fsim_download_grid(h, FSIM_FIELD_PRESSURE, p.data(), (int64_t)p.size() * 8));
    Sums s;
    for (int64_t c = 0; c < nc; c++) {  // reference order: z fastest
        const int z = (int)(c % si.gz_local) + si.z_offset;
        if (z < si.own_lo || z >= si.own_hi) continue;
        s.p += p[c];
        for (int a = 0; a < 3; a++) s.v2 += std::fabs(v2[3 * c + a]);
    }
    return s;
}

int main(int argc, char** argv) {
    const int ranks = argc > 1 ? std::atoi(argv[1]) : 2, n = argc > 2 ? std::atoi(argv[2]) : 16, steps = argc > 3 ? std::atoi(argv[3]) : 3;
    // several ranks on ONE device inside ONE process: see DESIGN.md §7 "Caveats" (not needed with one process per GPU)
    setenv("CUDA_MODULE_LOADING", "EAGER", 1);
    setenv("FSIM_MG_TAIL_CLUSTER", "2", 0);
    setenv("FSIM_DIST_GATHER_BLOCKS", "16", 0);
    FsimGridDesc d;
    std::memset(&d, 0, sizeof(d));
    d.target_dims[0] = d.target_dims[1] = d.target_dims[2] = n; d.resolution = 1.0; d.particle_radius = 0.25; d.device = 0;
    const std::vector<double> parts = dam(n);
    const int64_t np = (int64_t)parts.size() / 15;
    d.particle_capacity = np;
    const FsimParams par = params();
    const double dt = 0.005;

    fsim_t* one = nullptr;
    CK(nullptr, fsim_create(&d, &one));
    CK(one, fsim_set_params(one, &par));
    CK(one, fsim_upload_particles(one, parts.data(), np));
    std::vector<int> its1(steps);
    for (int s = 0; s < steps; s++) CK(one, fsim_step(one, dt, &its1[s]));
    CK(one, fsim_synchronize(one));
    const Sums ref = owned_sums(one);

    std::vector<fsim_t*> h(ranks, nullptr);
    std::vector<FsimDistExport> ex(ranks);
    for (int r = 0; r < ranks; r++) { CK(nullptr, fsim_create_slab(&d, r, ranks, &h[r])); CK(h[r], fsim_dist_export(h[r], &ex[r])); }
    for (int r = 0; r < ranks; r++) CK(h[r], fsim_dist_connect(h[r], ex.data(), ranks));
    for (int r = 0; r < ranks; r++) {
        FsimSlabInfo si; FsimGridInfo gi;
        CK(h[r], fsim_get_slab_info(h[r], &si)); CK(h[r], fsim_get_grid_info(h[r], &gi));
        std::vector<double> mine;
        for (int64_t i = 0; i < np; i++) {  // the device bins with trunc(float(z) * cellDInv.z) (simulator.cpp:358-359)
            const int iz = (int)((double)(float)parts[15 * i + 2] * gi.cell_d_inv[2]);
            if (iz >= si.own_lo && iz < si.own_hi) mine.insert(mine.end(), &parts[15 * i], &parts[15 * i] + 15);
        }
        CK(h[r], fsim_set_params(h[r], &par));
        CK(h[r], fsim_upload_particles(h[r], mine.data(), (int64_t)mine.size() / 15));
    }
    std::vector<std::vector<int>> its(ranks, std::vector<int>(steps));
    std::vector<std::thread> th;
    for (int r = 0; r < ranks; r++)
        th.emplace_back([&, r] { for (int s = 0; s < steps; s++) CK(h[r], fsim_step(h[r], dt, &its[r][s])); CK(h[r], fsim_synchronize(h[r])); });
    for (auto& t : th) t.join();

    Sums tot; int64_t count = 0; bool ok = true;
    for (int r = 0; r < ranks; r++) {
        const Sums s = owned_sums(h[r]);
        tot.v2 += s.v2; tot.p += s.p;
        int64_t c = 0; CK(h[r], fsim_particle_count(h[r], &c)); count += c;
        for (int s2 = 0; s2 < steps; s2++) ok = ok && its[r][s2] == its[0][s2] && std::abs(its[r][s2] - its1[s2]) <= 1;
    }
    const double ev = std::fabs(tot.v2 - ref.v2) / ref.v2, ep = std::fabs(tot.p - ref.p) / std::fabs(ref.p);
    std::printf("ranks %d grid %d^3 particles %lld (slabs hold %lld) its single %d slab %d  sum|v2| rel %.3e  sum p rel %.3e\n", ranks, n,
                (long long)np, (long long)count, its1[steps - 1], its[0][steps - 1], ev, ep);
    ok = ok && count == np && ev < 1e-5 && ep < 1e-5;
    for (int r = 0; r < ranks; r++) fsim_synchronize(h[r]);
    for (int r = 0; r < ranks; r++) fsim_destroy(h[r]);
    fsim_destroy(one);
    std::puts(ok ? "SLAB_HOST_OK" : "SLAB_HOST_MISMATCH");
    return ok ? 0 : 1;
}

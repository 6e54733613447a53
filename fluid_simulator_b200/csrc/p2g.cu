// p2g.cu -- particle-to-grid transfer (MacGrid::resetGridValues + Simulator::p2gTransfer scatter +
// the avgPNum half of markFluidCellsAndCalculateParticleDensities; macGrid.cpp:221-229,
// simulator.cpp:314-333, 362-366).
//
// Design: particles are cell-binned, so the scatter is tile-local.  One CTA owns a
// 32x4x4 tile of cells, one THREAD owns one cell and walks that cell's particles sequentially.  All
// particles of a cell touch the same 2x3x3 (staggered axis) / 3x3x3 (cell-centred) node neighbourhood,
// so the thread accumulates the whole neighbourhood in REGISTERS (hat weights are exactly 0 outside the
// 2x2x2 bracket the reference picks, macGrid.cpp:142-148) and only then folds it into a shared-memory
// tile with plain read-modify-writes in 9 conflict-free phases: in phase (dy,dz) warp (j,k) owns row
// (j+dy,k+dz), lanes own distinct x -- no shared-memory atomics (fp32 smem atomicAdd is a CAS loop on
// sm_100a) and a deterministic summation order inside the tile.  The tile (+1-cell halo) is flushed with
// one fp32 RED per touched node and channel; only tile-halo nodes are shared between CTAs.
#include "fsim_internal.h"

namespace {

constexpr int TX = 32, TY = 4, TZ = 4;
constexpr int SX = TX + 2, SY = TY + 2, SZ = TZ + 2;
constexpr int SN = SX * SY * SZ;  // nodes per channel in the shared tile
constexpr int NTHREADS = TX * TY * TZ;

struct P2GArgs {
    GridDims g;
    const float *px, *py, *pz, *vx, *vy, *vz;
    const float* c[9];
    const uint32_t *cnt, *cell_start;
    float *su[3], *wu[3], *dens;  // global accumulators (zeroed before launch)
    int apic;
};

__device__ __forceinline__ int sidx(int sx, int sy, int sz) { return (sz * SY + sy) * SX + sx; }

// hat weights of the three cell-centred nodes (c-1, c, c+1) for a particle at fractional offset f in [0,1) of cell c
__device__ __forceinline__ void centred_w(float f, float w[3]) {
    const float a = f - 0.5f;
    w[0] = fmaxf(0.f, -a);
    w[1] = 1.f - fabsf(a);
    w[2] = fmaxf(0.f, a);
}

// One staggered pass.  AXIS = 0,1,2: face grid of that axis (2 nodes along AXIS, 3 along the others), values + weights.
// AXIS = 3: cell-centred density (3x3x3 nodes, weights only).
template <int AXIS>
__device__ __forceinline__ void p2g_pass(const P2GArgs& a, float* s_val, float* s_w, bool active, int x, int y, int z,
                                         uint32_t start, uint32_t n, int lane, int wy, int wz) {
    constexpr int NX = (AXIS == 0) ? 2 : 3, NY = (AXIS == 1) ? 2 : 3, NZ = (AXIS == 2) ? 2 : 3;
    float accw[NZ][NY][NX];
    float accv[NZ][NY][NX];
#pragma unroll
    for (int k = 0; k < NZ; k++)
#pragma unroll
        for (int j = 0; j < NY; j++)
#pragma unroll
            for (int i = 0; i < NX; i++) { accw[k][j][i] = 0.f; accv[k][j][i] = 0.f; }

    if (active) {
        const float* vel = AXIS == 0 ? a.vx : (AXIS == 1 ? a.vy : a.vz);
        for (uint32_t t = 0; t < n; t++) {
            const uint32_t p = start + t;
            const float fx = fmaf(__ldg(a.px + p), a.g.ihx, -(float)x);  // single rounding
            const float fy = fmaf(__ldg(a.py + p), a.g.ihy, -(float)y);
            const float fz = fmaf(__ldg(a.pz + p), a.g.ihz, -(float)z);
            float wx[3], wy_[3], wz_[3];
            // offsets (in cells) from the particle to node i along each dimension, for the APIC term
            float ox[3], oy[3], oz[3];
            if (AXIS == 0) { wx[0] = 1.f - fx; wx[1] = fx; wx[2] = 0.f; ox[0] = -fx; ox[1] = 1.f - fx; ox[2] = 0.f; }
            else { centred_w(fx, wx); ox[0] = -0.5f - fx; ox[1] = 0.5f - fx; ox[2] = 1.5f - fx; }
            if (AXIS == 1) { wy_[0] = 1.f - fy; wy_[1] = fy; wy_[2] = 0.f; oy[0] = -fy; oy[1] = 1.f - fy; oy[2] = 0.f; }
            else { centred_w(fy, wy_); oy[0] = -0.5f - fy; oy[1] = 0.5f - fy; oy[2] = 1.5f - fy; }
            if (AXIS == 2) { wz_[0] = 1.f - fz; wz_[1] = fz; wz_[2] = 0.f; oz[0] = -fz; oz[1] = 1.f - fz; oz[2] = 0.f; }
            else { centred_w(fz, wz_); oz[0] = -0.5f - fz; oz[1] = 0.5f - fz; oz[2] = 1.5f - fz; }
            if (AXIS == 3) {
#pragma unroll
                for (int k = 0; k < NZ; k++)
#pragma unroll
                    for (int j = 0; j < NY; j++) {
                        const float wyz = wy_[j] * wz_[k];
#pragma unroll
                        for (int i = 0; i < NX; i++) accw[k][j][i] += wx[i] * wyz;
                    }
            } else {
                const float v = __ldg(vel + p);
                float cx = 0.f, cy = 0.f, cz = 0.f;
                if (a.apic) {  // c[AXIS] . (face.pos - particle.pos), simulator.cpp:327-328
                    cx = __ldg(a.c[3 * (AXIS % 3) + 0] + p) * a.g.hx;
                    cy = __ldg(a.c[3 * (AXIS % 3) + 1] + p) * a.g.hy;
                    cz = __ldg(a.c[3 * (AXIS % 3) + 2] + p) * a.g.hz;
                }
#pragma unroll
                for (int k = 0; k < NZ; k++)
#pragma unroll
                    for (int j = 0; j < NY; j++) {
                        const float wyz = wy_[j] * wz_[k];
                        const float vyz = v + cy * oy[j] + cz * oz[k];
#pragma unroll
                        for (int i = 0; i < NX; i++) {
                            const float w = wx[i] * wyz;
                            accw[k][j][i] += w;
                            accv[k][j][i] += w * (vyz + cx * ox[i]);
                        }
                    }
            }
        }
    }

    // fold the register neighbourhood into the shared tile: phase (k,j) -> warp (wy,wz) owns row (wy+j, wz+k)
#pragma unroll
    for (int k = 0; k < NZ; k++)
#pragma unroll
        for (int j = 0; j < NY; j++) {
            // node index along each dim: staggered axis -> {c-1, c} ; centred -> {c-1, c, c+1}; smem index = cell_local + 1 + (node - c)
            const int row = sidx(0, wy + j, wz + k);
#pragma unroll
            for (int i = 0; i < NX; i++) {
                const int id = row + lane + i;
                if (active && accw[k][j][i] != 0.f) {
                    s_w[id] += accw[k][j][i];
                    if (AXIS != 3) s_val[id] += accv[k][j][i];
                }
                __syncwarp();
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(NTHREADS, 2) p2g_kernel(P2GArgs a) {
    __shared__ float s[7][SN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wy = warp % TY, wz = warp / TY;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, z0 = blockIdx.z * TZ;
    const int x = x0 + lane, y = y0 + wy, z = z0 + wz;
    const bool inside = x < a.g.gx && y < a.g.gy && z < a.g.gz;
    uint32_t n = 0, start = 0;
    if (inside) {
        const int64_t c = ((int64_t)z * a.g.gy + y) * a.g.gx + x;
        n = a.cnt[c];
        start = a.cell_start[c];
    }
    if (!__syncthreads_or(n > 0)) return;  // empty tile: nothing to scatter

    for (int i = threadIdx.x; i < 7 * SN; i += NTHREADS) (&s[0][0])[i] = 0.f;
    __syncthreads();

    const bool active = n > 0;
    p2g_pass<0>(a, s[0], s[3], active, x, y, z, start, n, lane, wy, wz);
    p2g_pass<1>(a, s[1], s[4], active, x, y, z, start, n, lane, wy, wz);
    p2g_pass<2>(a, s[2], s[5], active, x, y, z, start, n, lane, wy, wz);
    p2g_pass<3>(a, nullptr, s[6], active, x, y, z, start, n, lane, wy, wz);

    // flush tile + halo: one RED per touched node and channel (x fastest => coalesced)
    for (int i = threadIdx.x; i < SN; i += NTHREADS) {
        const int sx = i % SX, sy = (i / SX) % SY, sz = i / (SX * SY);
        const int gx = x0 - 1 + sx, gy = y0 - 1 + sy, gz = z0 - 1 + sz;
        if (gx < 0 || gy < 0 || gz < 0 || gx >= a.g.gx || gy >= a.g.gy || gz >= a.g.gz) continue;
        const int64_t c = ((int64_t)gz * a.g.gy + gy) * a.g.gx + gx;
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const float w = s[3 + ax][i];
            if (w != 0.f) {
                atomicAdd(a.wu[ax] + c, w);
                atomicAdd(a.su[ax] + c, s[ax][i]);
            }
        }
        const float d = s[6][i];
        if (d != 0.f) atomicAdd(a.dens + c, d);
    }
}

}  // namespace

int k_p2g(fsim* h) {
    const GridDims& g = h->g;
    // MacGrid::resetGridValues (macGrid.cpp:221-229): v, v2, weights, avgPNum = 0 (type is rewritten by classify)
    {
        KScope ks(h, K_MEMSET, 7);
        for (int a = 0; a < 3; a++) {
            FSIM_CUDA(h, cudaMemsetAsync(h->u[a], 0, sizeof(float) * g.nc, h->stream));
            FSIM_CUDA(h, cudaMemsetAsync(h->wsum[a], 0, sizeof(float) * g.nc, h->stream));
        }
        FSIM_CUDA(h, cudaMemsetAsync(h->dens, 0, sizeof(float) * g.nc, h->stream));
    }
    if (h->np == 0) return FSIM_OK;
    P2GArgs a;
    a.g = g;
    const ParticleSet& p = h->ps[h->cur];
    a.px = p.pos[0]; a.py = p.pos[1]; a.pz = p.pos[2];
    a.vx = p.vel[0]; a.vy = p.vel[1]; a.vz = p.vel[2];
    for (int c = 0; c < 9; c++) a.c[c] = p.c[c];
    a.cnt = h->cnt; a.cell_start = h->cell_start;
    for (int ax = 0; ax < 3; ax++) { a.su[ax] = h->u[ax]; a.wu[ax] = h->wsum[ax]; }
    a.dens = h->dens;
    a.apic = (h->par.transfer_type == FSIM_TRANSFER_APIC) && h->have_c;
    dim3 grid(div_up(g.gx, TX), div_up(g.gy, TY), div_up(g.gz, TZ));
    { KScope ks(h, K_P2G); p2g_kernel<<<grid, NTHREADS, 0, h->stream>>>(a); }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

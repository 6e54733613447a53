// grid.cu -- grid-side stages between P2G and the pressure solve, and after it:
//   classify_kernel   cell.type from particle counts (markFluidCells..., simulator.cpp:357-360), obstacle
//                     rasterisation (MacGrid::addObstacle, macGrid.cpp:231-292) and the solid border shell
//                     (restoreBorderingSolidCellsAndSpeeds, macGrid.cpp:75-140)
//   finalize_kernel   P2G normalise (simulator.cpp:335-351), obstacle / border face velocities, v2 = v (+ g*dt)
//                     (postP2GUpdate, macGrid.cpp:204-211)
//   pressure_apply    applyPressureToVelocities (bridsonSolverGrid.cpp:295-327)
//   extrapolate       extrapolateVelocities (macGrid.cpp:294-336), two Jacobi sweeps
//   transposing up/download between the device layout (x fastest) and the reference's (z fastest)
// All kernels are one thread per cell, x fastest => fully coalesced, HBM-bound.
#include "fsim_internal.h"

namespace {

__device__ __forceinline__ int64_t cidx(const GridDims& g, int x, int y, int z) { return ((int64_t)z * g.gy + y) * g.gx + x; }

// is cell (x,y,z) rasterised SOLID by obstacle ob?  Bounds come from the host (exact reference expressions); the
// sphere test repeats macGrid.cpp:244-245 in fp64 without FMA contraction so the flag is bit-exact.
__device__ __forceinline__ bool in_obstacle(const DevObstacle& ob, int x, int y, int z) {
    if (ob.kind == FSIM_OBSTACLE_SINK) return false;  // a sink is not solid on the grid (macGrid.cpp:235)
    if (ob.kind == FSIM_OBSTACLE_BOX)
        return x >= ob.mn[0] && x < ob.mx[0] && y >= ob.mn[1] && y < ob.mx[1] && z >= ob.mn[2] && z < ob.mx[2];
    if (x < ob.mn[0] || x > ob.mx[0] || y < ob.mn[1] || y > ob.mx[1] || z < ob.mn[2] || z > ob.mx[2]) return false;
    const double dx = __dsub_rn(x + 0.5, ob.center[0]), dy = __dsub_rn(y + 0.5, ob.center[1]), dz = __dsub_rn(z + 0.5, ob.center[2]);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return d2 < ob.r2;
}

__device__ __forceinline__ unsigned obstacle_mask(const DevObstacle* obs, int nobs, int x, int y, int z) {
    unsigned m = 0;
    for (int k = 0; k < nobs; k++)
        if (in_obstacle(obs[k], x, y, z)) m |= 1u << k;
    return m;
}

// one thread per cell on a (32,4,2) block / 3-D grid: (x, y, z) come straight from the block and thread indices
// (a flat index would cost two 64-bit divisions per thread, which made these short kernels issue-bound)
constexpr int CBX = 32, CBY = 4, CBZ = 2;
__device__ __forceinline__ bool cell_xyz(const GridDims& g, int& x, int& y, int& z, int64_t& c) {
    x = blockIdx.x * CBX + threadIdx.x;
    y = blockIdx.y * CBY + threadIdx.y;
    z = blockIdx.z * CBZ + threadIdx.z;
    c = ((int64_t)z * g.gy + y) * g.gx + x;
    return x < g.gx && y < g.gy && z < g.gz;
}
inline dim3 cell_grid(const GridDims& g) { return dim3(div_up(g.gx, CBX), div_up(g.gy, CBY), div_up(g.gz, CBZ)); }
inline dim3 cell_block() { return dim3(CBX, CBY, CBZ); }

__device__ __forceinline__ bool is_border(const GridDims& g, int x, int y, int z, int top_solid) {
    return x == 0 || x == g.gx - 1 || y == 0 || (top_solid && y == g.gy - 1) || z + g.zoff == 0 || z + g.zoff == g.gzg - 1;
}

__global__ void __launch_bounds__(256) classify_kernel(GridDims g, const uint32_t* __restrict__ cnt, const DevObstacle* obs,
                                                        int nobs, int top_solid, uint8_t* __restrict__ flags) {
    int x, y, z;
    int64_t c;
    if (!cell_xyz(g, x, y, z, c)) return;
    const bool has = cnt[c] > 0;
    int type = has ? FSIM_CELL_WATER : FSIM_CELL_AIR;
    if (nobs > 0 && obstacle_mask(obs, nobs, x, y, z + g.zoff)) type = FSIM_CELL_SOLID;
    if (is_border(g, x, y, z, top_solid)) type = FSIM_CELL_SOLID;
    const int valid = (type == FSIM_CELL_WATER) ? 0 : 3;
    flags[c] = (uint8_t)(type | (has ? FL_HASPART : 0) | (valid << FL_VALID_SHIFT));
}

struct FinalizeArgs {
    GridDims g;
    const uint8_t* flags;
    float *u[3], *u2[3];
    const float* wsum[3];
    const DevObstacle* obs;
    int nobs, top_solid;
    int post_only;  // 1: MacGrid::postP2GUpdate alone (v2 = v + gravity) for projection-only set-ups
    float gdt;      // gravity * dt (0 when gravity is disabled)
};

// The reference rasterises obstacles one after the other, cell by cell in lexicographic order, and each newly SOLID
// cell overwrites the face it shares with a neighbour that is WATER *at that moment* (macGrid.cpp:247-259).
// For the face between cell A (lower) and B = A + e_axis that is: the LAST obstacle k such that
//   (A in O_k, B has particles, B not covered by any obstacle < k)   [B comes after A lexicographically]
//   or (B in O_k, A has particles, A not covered by any obstacle <= k).
__device__ __forceinline__ int face_obstacle(unsigned mA, unsigned mB, bool hasA, bool hasB) {
    int best = -1;
    if (hasB && mA) {
        const unsigned allow = mB ? ((mB & (0u - mB)) << 1) - 1u : 0xffffffffu;  // k <= first(B)
        const unsigned cand = mA & allow;
        if (cand) best = 31 - __clz(cand);
    }
    if (hasA && mB) {
        const unsigned allow = mA ? (mA & (0u - mA)) - 1u : 0xffffffffu;  // k < first(A)
        const unsigned cand = mB & allow;
        if (cand) best = max(best, 31 - __clz(cand));
    }
    return best;
}

__global__ void __launch_bounds__(256) finalize_kernel(FinalizeArgs a) {
    const GridDims& g = a.g;
    int x, y, z;
    int64_t c;
    if (!cell_xyz(g, x, y, z, c)) return;
    const uint8_t f0 = a.flags[c];
    const int t0 = f0 & FL_TYPE_MASK;
    const int nx[3] = {x + 1, x, x}, ny[3] = {y, y + 1, y}, nz[3] = {z, z, z + 1};
    unsigned m0 = 0;
    if (a.nobs > 0) m0 = obstacle_mask(a.obs, a.nobs, x, y, z + g.zoff);
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        float val = a.u[ax][c];
        if (!a.post_only) {
            const float w = a.wsum[ax][c];
            val = (w > 1e-6f) ? val / w : 0.f;  // simulator.cpp:335-351
        }
        const bool nb_ok = nx[ax] < g.gx && ny[ax] < g.gy && nz[ax] < g.gz;
        int t1 = FSIM_CELL_SOLID;
        uint8_t f1 = 0;
        if (nb_ok) { f1 = a.flags[cidx(g, nx[ax], ny[ax], nz[ax])]; t1 = f1 & FL_TYPE_MASK; }
        if (a.nobs > 0 && nb_ok && !a.post_only) {
            const unsigned m1 = obstacle_mask(a.obs, a.nobs, nx[ax], ny[ax], nz[ax] + g.zoff);
            const int k = face_obstacle(m0, m1, (f0 & FL_HASPART) != 0, (f1 & FL_HASPART) != 0);
            if (k >= 0) val = (float)a.obs[k].speed[ax];
        }
        // border shell: wall-normal face next to a WATER interior cell is zeroed (macGrid.cpp:80-104)
        const int pos_ax = ax == 0 ? x : (ax == 1 ? y : z + g.zoff);
        const int gsz = ax == 0 ? g.gx : (ax == 1 ? g.gy : g.gzg);
        if (!a.post_only) {
            if (pos_ax == 0 && t1 == FSIM_CELL_WATER) val = 0.f;
            if (pos_ax == gsz - 2 && t0 == FSIM_CELL_WATER && (ax != 1 || a.top_solid)) val = 0.f;
        }
        a.u[ax][c] = val;
        float v2 = val;
        // gravity on y unless this cell or the one the reference reads as "+y" is SOLID (macGrid.cpp:209-210); for the
        // top row the reference's cell<1,1>() lands on the y=0 border cell of the next x column => never gravity.
        if (ax == 1 && t0 != FSIM_CELL_SOLID && t1 != FSIM_CELL_SOLID) v2 += a.gdt;
        a.u2[ax][c] = v2;
    }
}

struct ApplyArgs {
    GridDims g;
    const uint8_t* flags;
    const double* p;
    float* u2[3];
    double scale;  // dt / (rho * h)
};

__global__ void __launch_bounds__(256) pressure_apply_kernel(ApplyArgs a) {
    const GridDims& g = a.g;
    int x, y, z;
    int64_t c;
    if (!cell_xyz(g, x, y, z, c)) return;
    const int t0 = a.flags[c] & FL_TYPE_MASK;
    if (t0 == FSIM_CELL_SOLID) return;
    const double p0 = (t0 == FSIM_CELL_WATER) ? a.p[c] : 0.0;
    const int nx[3] = {x + 1, x, x}, ny[3] = {y, y + 1, y}, nz[3] = {z, z, z + 1};
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        if (nx[ax] >= g.gx || ny[ax] >= g.gy || nz[ax] >= g.gz) continue;
        const int64_t cn = cidx(g, nx[ax], ny[ax], nz[ax]);
        const int t1 = a.flags[cn] & FL_TYPE_MASK;
        // WATER|non-solid nbr: += scale*(p0 - p1) ; AIR with WATER +nbr: -= scale*p1   (p = 0 outside WATER)
        const bool touch = (t0 == FSIM_CELL_WATER && t1 != FSIM_CELL_SOLID) || (t0 == FSIM_CELL_AIR && t1 == FSIM_CELL_WATER);
        if (!touch) continue;
        const double p1 = (t1 == FSIM_CELL_WATER) ? a.p[cn] : 0.0;
        a.u2[ax][c] = (float)((double)a.u2[ax][c] + a.scale * (p0 - p1));
    }
}

struct ExtrapArgs {
    GridDims g;
    uint8_t* flags;
    float* u2[3];
    int it;
};

__global__ void __launch_bounds__(256) extrapolate_kernel(ExtrapArgs a) {
    const GridDims& g = a.g;
    int x, y, z;
    int64_t c;
    if (!cell_xyz(g, x, y, z, c)) return;
    const uint8_t f0 = a.flags[c];
    const int valid0 = (f0 & FL_VALID_MASK) >> FL_VALID_SHIFT;
    if (valid0 <= a.it) return;
    int cntv = 0;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    // neighbour order of the reference: -x,+x,-y,+y,-z,+z (macGrid.cpp:313-324)
    const int dx[6] = {-1, 1, 0, 0, 0, 0}, dy[6] = {0, 0, -1, 1, 0, 0}, dz[6] = {0, 0, 0, 0, -1, 1};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int xx = x + dx[k], yy = y + dy[k], zz = z + dz[k];
        if (xx < 0 || yy < 0 || zz < 0 || xx >= g.gx || yy >= g.gy || zz >= g.gz) continue;
        const int64_t cn = cidx(g, xx, yy, zz);
        const int vn = (a.flags[cn] & FL_VALID_MASK) >> FL_VALID_SHIFT;
        if (vn <= a.it) {
            s0 += a.u2[0][cn]; s1 += a.u2[1][cn]; s2 += a.u2[2][cn];
            cntv++;
        }
    }
    if (cntv > 0) {
        const float inv = 1.f / (float)cntv;
        if (x + 1 < g.gx && (a.flags[c + 1] & FL_TYPE_MASK) != FSIM_CELL_WATER) a.u2[0][c] = s0 * inv;
        if (y + 1 < g.gy && (a.flags[c + g.sy] & FL_TYPE_MASK) != FSIM_CELL_WATER) a.u2[1][c] = s1 * inv;
        if (z + 1 < g.gz && (a.flags[c + g.sz] & FL_TYPE_MASK) != FSIM_CELL_WATER) a.u2[2][c] = s2 * inv;
        a.flags[c] = (uint8_t)((f0 & ~FL_VALID_MASK) | ((a.it + 1) << FL_VALID_SHIFT));
    }
}

// The same sweep with four x-consecutive cells per thread (gx % 4 == 0): the validity bytes of the group and of its y / z
// neighbour groups arrive as 4-byte loads, and a group whose four cells all have nothing to do (all WATER, or no valid
// neighbour anywhere near: most of the grid) retires after them.  One thread per cell spent 95 us per sweep on 16.8 M
// threads that each read 1-7 single bytes (r1 ncu: 0.7 TB/s, 63 % issue-active).
__device__ __forceinline__ unsigned valid4(unsigned f4) { return (f4 >> FL_VALID_SHIFT) & 0x03030303u; }  // per-byte validity 0..3

__global__ void __launch_bounds__(256) extrapolate4_kernel(ExtrapArgs a) {
    const GridDims& g = a.g;
    const int x = (blockIdx.x * CBX + threadIdx.x) * 4, y = blockIdx.y * CBY + threadIdx.y, z = blockIdx.z * CBZ + threadIdx.z;
    if (x >= g.gx || y >= g.gy || z >= g.gz) return;
    const int64_t c = ((int64_t)z * g.gy + y) * g.gx + x;
    const unsigned it = (unsigned)a.it;
    const unsigned f0 = *reinterpret_cast<const unsigned*>(a.flags + c);
    const unsigned v0 = valid4(f0);
    // cells that still need a value this sweep: validity > it
    bool need[4];
    bool any_need = false;
#pragma unroll
    for (int i = 0; i < 4; i++) { need[i] = ((v0 >> (8 * i)) & 3u) > it; any_need |= need[i]; }
    if (!any_need) return;
    auto ld4 = [&](int yy, int zz) -> unsigned {  // validity bytes of the neighbour group; 3 (never valid) outside the grid
        if (yy < 0 || zz < 0 || yy >= g.gy || zz >= g.gz) return 0x03030303u;
        return valid4(*reinterpret_cast<const unsigned*>(a.flags + ((int64_t)zz * g.gy + yy) * g.gx + x));
    };
    const unsigned vym = ld4(y - 1, z), vyp = ld4(y + 1, z), vzm = ld4(y, z - 1), vzp = ld4(y, z + 1);
    const unsigned vxl = x > 0 ? (unsigned)((a.flags[c - 1] & FL_VALID_MASK) >> FL_VALID_SHIFT) : 3u;
    const unsigned vxr = x + 4 < g.gx ? (unsigned)((a.flags[c + 4] & FL_VALID_MASK) >> FL_VALID_SHIFT) : 3u;
    unsigned newf = f0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (!need[i]) continue;
        const int64_t ci = c + i;
        // neighbour order of the reference: -x, +x, -y, +y, -z, +z (macGrid.cpp:313-324)
        const unsigned vn[6] = {i == 0 ? vxl : (v0 >> (8 * (i - 1))) & 3u, i == 3 ? vxr : (v0 >> (8 * (i + 1))) & 3u,
                                (vym >> (8 * i)) & 3u, (vyp >> (8 * i)) & 3u, (vzm >> (8 * i)) & 3u, (vzp >> (8 * i)) & 3u};
        const int64_t cn[6] = {ci - 1, ci + 1, ci - g.sy, ci + g.sy, ci - g.sz, ci + g.sz};
        int cntv = 0;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (vn[k] <= it) {
                s0 += a.u2[0][cn[k]]; s1 += a.u2[1][cn[k]]; s2 += a.u2[2][cn[k]];
                cntv++;
            }
        if (cntv > 0) {
            const float inv = 1.f / (float)cntv;
            // own + face only where the + neighbour exists and is not WATER: types of the x neighbours come from the group itself
            const unsigned tx = i == 3 ? (x + 4 < g.gx ? (unsigned)(a.flags[c + 4] & FL_TYPE_MASK) : 0u) : (f0 >> (8 * (i + 1))) & FL_TYPE_MASK;
            if (x + i + 1 < g.gx && tx != FSIM_CELL_WATER) a.u2[0][ci] = s0 * inv;
            if (y + 1 < g.gy && (a.flags[ci + g.sy] & FL_TYPE_MASK) != FSIM_CELL_WATER) a.u2[1][ci] = s1 * inv;
            if (z + 1 < g.gz && (a.flags[ci + g.sz] & FL_TYPE_MASK) != FSIM_CELL_WATER) a.u2[2][ci] = s2 * inv;
            newf = (newf & ~((unsigned)FL_VALID_MASK << (8 * i))) | ((it + 1u) << (FL_VALID_SHIFT + 8 * i));
        }
    }
    if (newf != f0) *reinterpret_cast<unsigned*>(a.flags + c) = newf;
}

// BasicMacGrid::solveIncompressibility (basicMacGrid.cpp:15-102): red-black Gauss-Seidel with over-relaxation 1.98 acting
// directly on the face velocities; one launch per colour ((x+y+z) odd first, then even).  Cells of one colour share no
// face, so the in-place update is race-free and -- unlike the PCG -- reproduces the reference sweep for sweep.
struct BasicArgs {
    GridDims g;
    const uint8_t* flags;
    float* u2[3];
    const float* dens;
    double avg_pressure, pressure_k;
    int pressure_enabled, parity;
};

__global__ void __launch_bounds__(256) basic_sor_kernel(BasicArgs a) {
    const GridDims& g = a.g;
    int x, y, z;
    int64_t c;
    if (!cell_xyz(g, x, y, z, c)) return;
    const int zg = z + g.zoff;  // (slab handles: global plane; they sweep the planes they own)
    if (((x + y + zg) & 1) != a.parity) return;
    if (x < 1 || y < 1 || zg < 1 || x >= g.gx - 1 || y >= g.gy - 1 || zg >= g.gzg - 1 || z < g.zown0 || z >= g.zown1) return;
    if ((a.flags[c] & FL_TYPE_MASK) != FSIM_CELL_WATER) return;
    const int s1 = (a.flags[c + g.sz] & FL_TYPE_MASK) != FSIM_CELL_SOLID, s2 = (a.flags[c - g.sz] & FL_TYPE_MASK) != FSIM_CELL_SOLID;
    const int s3 = (a.flags[c + g.sy] & FL_TYPE_MASK) != FSIM_CELL_SOLID, s4 = (a.flags[c - g.sy] & FL_TYPE_MASK) != FSIM_CELL_SOLID;
    const int s5 = (a.flags[c + 1] & FL_TYPE_MASK) != FSIM_CELL_SOLID, s6 = (a.flags[c - 1] & FL_TYPE_MASK) != FSIM_CELL_SOLID;
    const int s = s1 + s2 + s3 + s4 + s5 + s6;
    if (s == 0) return;
    double d = -(double)a.u2[0][c] - (double)a.u2[1][c] - (double)a.u2[2][c] + (double)a.u2[0][c - 1] + (double)a.u2[1][c - g.sy] +
               (double)a.u2[2][c - g.sz] + (a.pressure_enabled ? ((double)a.dens[c] - a.avg_pressure) * a.pressure_k : 0.0);
    d = d * 1.98 / s;
    if (s1) a.u2[2][c] = (float)((double)a.u2[2][c] + d);
    if (s2) a.u2[2][c - g.sz] = (float)((double)a.u2[2][c - g.sz] - d);
    if (s3) a.u2[1][c] = (float)((double)a.u2[1][c] + d);
    if (s4) a.u2[1][c - g.sy] = (float)((double)a.u2[1][c - g.sy] - d);
    if (s5) a.u2[0][c] = (float)((double)a.u2[0][c] + d);
    if (s6) a.u2[0][c - 1] = (float)((double)a.u2[0][c - 1] - d);
}

// ---- transposing download / upload (device x-fastest <-> reference z-fastest) ---------------------------
struct XferArgs {
    GridDims g;
    int field;
    const uint8_t* flags;
    const float* f3[3];
    const float* f1;
    const double* d1;
    const uint32_t* cnt;
    void* out;
};

__global__ void __launch_bounds__(256) grid_download_kernel(XferArgs a) {
    const GridDims& g = a.g;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // reference-order index
    if (r >= g.nc) return;
    const int z = (int)(r % g.gz), y = (int)((r / g.gz) % g.gy), x = (int)(r / ((int64_t)g.gz * g.gy));
    const int64_t c = cidx(g, x, y, z);
    switch (a.field) {
        case FSIM_FIELD_TYPE: ((uint8_t*)a.out)[r] = a.flags[c] & FL_TYPE_MASK; break;
        case FSIM_FIELD_V: case FSIM_FIELD_V2: case FSIM_FIELD_WSUM:
            for (int ax = 0; ax < 3; ax++) ((double*)a.out)[3 * r + ax] = (double)a.f3[ax][c];
            break;
        case FSIM_FIELD_AVGPNUM: ((double*)a.out)[r] = (double)a.f1[c]; break;
        case FSIM_FIELD_PRESSURE: case FSIM_FIELD_RHS:
            ((double*)a.out)[r] = ((a.flags[c] & FL_TYPE_MASK) == FSIM_CELL_WATER) ? a.d1[c] : 0.0;
            break;
        case FSIM_FIELD_PCOUNT: ((int32_t*)a.out)[r] = (int32_t)a.cnt[c]; break;
    }
}

struct UpArgs {
    GridDims g;
    int field;
    uint8_t* flags;
    float* f3[3];
    float* f1;
    const void* in;
};

__global__ void __launch_bounds__(256) grid_upload_kernel(UpArgs a) {
    const GridDims& g = a.g;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // device-order index
    if (c >= g.nc) return;
    const int x = (int)(c % g.gx), y = (int)((c / g.gx) % g.gy), z = (int)(c / ((int64_t)g.gx * g.gy));
    const int64_t r = ((int64_t)x * g.gy + y) * g.gz + z;
    switch (a.field) {
        case FSIM_FIELD_TYPE: {
            const int t = ((const uint8_t*)a.in)[r] & FL_TYPE_MASK;
            a.flags[c] = (uint8_t)(t | (t == FSIM_CELL_WATER ? FL_HASPART : 0) | ((t == FSIM_CELL_WATER ? 0 : 3) << FL_VALID_SHIFT));
            break;
        }
        case FSIM_FIELD_V: case FSIM_FIELD_V2: case FSIM_FIELD_WSUM:
            for (int ax = 0; ax < 3; ax++) a.f3[ax][c] = (float)((const double*)a.in)[3 * r + ax];
            break;
        case FSIM_FIELD_AVGPNUM: a.f1[c] = (float)((const double*)a.in)[r]; break;
    }
}

}  // namespace

int k_classify(fsim* h, double dt) {
    const GridDims& g = h->g;
    { KScope ks(h, K_CLASSIFY); classify_kernel<<<cell_grid(h->g), cell_block(), 0, h->stream>>>(g, h->cnt, h->d_obs, h->nobs, h->par.top_solid, h->flags); }
    FinalizeArgs a;
    a.g = g; a.flags = h->flags;
    for (int ax = 0; ax < 3; ax++) { a.u[ax] = h->u[ax]; a.u2[ax] = h->u2[ax]; a.wsum[ax] = h->wsum[ax]; }
    a.obs = h->d_obs; a.nobs = h->nobs; a.top_solid = h->par.top_solid; a.post_only = 0;
    // config.gravity is a float, dt a double: gravityEnabled ? gravity * dt : 0 (simulator.cpp:85)
    a.gdt = h->par.gravity_enabled ? (float)((double)h->par.gravity * dt) : 0.f;
    { KScope ks(h, K_FINALIZE); finalize_kernel<<<cell_grid(h->g), cell_block(), 0, h->stream>>>(a); }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

int k_post_p2g_only(fsim* h, double gravity_increment) {
    // projection-only path: flags were uploaded; u holds v (wsum must make the normalise a no-op => treat u as final)
    const GridDims& g = h->g;
    FinalizeArgs a;
    a.g = g; a.flags = h->flags;
    for (int ax = 0; ax < 3; ax++) { a.u[ax] = h->u[ax]; a.u2[ax] = h->u2[ax]; a.wsum[ax] = h->wsum[ax]; }
    a.obs = h->d_obs; a.nobs = 0; a.top_solid = h->par.top_solid; a.post_only = 1;
    a.gdt = (float)gravity_increment;
    { KScope ks(h, K_FINALIZE); finalize_kernel<<<cell_grid(h->g), cell_block(), 0, h->stream>>>(a); }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

int k_pressure_apply(fsim* h, double dt) {
    const GridDims& g = h->g;
    ApplyArgs a;
    a.g = g; a.flags = h->flags; a.p = h->p;
    for (int ax = 0; ax < 3; ax++) a.u2[ax] = h->u2[ax];
    a.scale = dt / (h->par.fluid_density * h->info.cell_d[0]);
    { KScope ks(h, K_APPLY); pressure_apply_kernel<<<cell_grid(h->g), cell_block(), 0, h->stream>>>(a); }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

int k_project_basic(fsim* h, int* iterations) {
    const GridDims& g = h->g;
    BasicArgs a;
    a.g = g; a.flags = h->flags; a.dens = h->dens;
    for (int ax = 0; ax < 3; ax++) a.u2[ax] = h->u2[ax];
    a.avg_pressure = h->par.average_pressure; a.pressure_k = h->par.pressure_k; a.pressure_enabled = h->par.pressure_enabled;
    const int n = h->par.max_iterations;
    for (int it = 0; it < n; it++)
        for (int colour = 1; colour >= 0; colour--) {  // odd cells first (z starts at 1 + (x+y)%2), then even
            a.parity = colour;
            { KScope ks(h, K_APPLY); basic_sor_kernel<<<cell_grid(h->g), cell_block(), 0, h->stream>>>(a); }
            // slab handles: a cell updates the z face it shares with the neighbouring slab; hand over the half each side wrote
            if (h->dist) { int rc = dist_halo(h, colour ? HALO_BASIC_ODD : HALO_BASIC_EVEN, false); if (rc) return rc; }
        }
    FSIM_CHECK_LAUNCH(h);
    h->solve.iterations = n; h->solve.early_out = 0; h->solve.rhs_sumsq = 0; h->solve.residual_max = 0; h->solve.fluid_cells = 0;
    h->pressure_valid = false;
    h->warm_history = 0;
    if (iterations) *iterations = n;  // the reference returns incompressibilityMaxIterationCount (basicMacGrid.cpp:10)
    return FSIM_OK;
}

int k_extrapolate(fsim* h) {
    const GridDims& g = h->g;
    ExtrapArgs a;
    a.g = g; a.flags = h->flags;
    for (int ax = 0; ax < 3; ax++) a.u2[ax] = h->u2[ax];
    for (int it = 0; it < 2; it++) {
        a.it = it;
        {
            KScope ks(h, K_EXTRAP);
            if (g.gx % 4 == 0) extrapolate4_kernel<<<dim3(div_up(g.gx, 4 * CBX), div_up(g.gy, CBY), div_up(g.gz, CBZ)), cell_block(), 0, h->stream>>>(a);
            else extrapolate_kernel<<<cell_grid(h->g), cell_block(), 0, h->stream>>>(a);
        }
        // slab mode: a sweep reads the validity and velocities the neighbours wrote in the previous one
        if (h->dist) { int rc = dist_halo(h, HALO_U2_FLAGS, false); if (rc) return rc; }
    }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

int k_grid_download(fsim* h, int field, void* dev_out) {
    XferArgs a;
    memset(&a, 0, sizeof(a));
    a.g = h->g; a.field = field; a.flags = h->flags; a.out = dev_out; a.cnt = h->cnt;
    switch (field) {
        case FSIM_FIELD_V: for (int ax = 0; ax < 3; ax++) a.f3[ax] = h->u[ax]; break;
        case FSIM_FIELD_V2: for (int ax = 0; ax < 3; ax++) a.f3[ax] = h->u2[ax]; break;
        case FSIM_FIELD_WSUM: for (int ax = 0; ax < 3; ax++) a.f3[ax] = h->wsum[ax]; break;
        case FSIM_FIELD_AVGPNUM: a.f1 = h->dens; break;
        case FSIM_FIELD_PRESSURE: a.d1 = h->p; break;
        case FSIM_FIELD_RHS: a.d1 = h->rhs; break;
        default: break;
    }
    grid_download_kernel<<<div_up(h->g.nc, 256), 256, 0, h->stream>>>(a);
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

int k_grid_upload(fsim* h, int field, const void* dev_in) {
    UpArgs a;
    memset(&a, 0, sizeof(a));
    a.g = h->g; a.field = field; a.flags = h->flags; a.in = dev_in;
    switch (field) {
        case FSIM_FIELD_V: for (int ax = 0; ax < 3; ax++) a.f3[ax] = h->u[ax]; break;
        case FSIM_FIELD_V2: for (int ax = 0; ax < 3; ax++) a.f3[ax] = h->u2[ax]; break;
        case FSIM_FIELD_WSUM: for (int ax = 0; ax < 3; ax++) a.f3[ax] = h->wsum[ax]; break;
        case FSIM_FIELD_AVGPNUM: a.f1 = h->dens; break;
        default: break;
    }
    grid_upload_kernel<<<div_up(h->g.nc, 256), 256, 0, h->stream>>>(a);
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

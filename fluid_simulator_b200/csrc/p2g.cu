// p2g.cu -- particle-to-grid transfer (MacGrid::resetGridValues + Simulator::p2gTransfer scatter +
// the avgPNum half of markFluidCellsAndCalculateParticleDensities; macGrid.cpp:221-229,
// simulator.cpp:314-333, 362-366).
//
// Design: particles are cell-binned, so the scatter is tile-local.  One CTA owns a
// 32x4x2 tile of cells (4 CTAs / SM; 32x4x4 with 2 CTAs / SM was 4-6 % slower: more warps per barrier), one THREAD owns one cell and walks that cell's particles sequentially.  All
// particles of a cell touch the same 2x3x3 (staggered axis) / 3x3x3 (cell-centred) node neighbourhood,
// so the thread accumulates the whole neighbourhood in REGISTERS with packed fp32x2 FMAs (FFMA2; hat weights are
// exactly 0 outside the 2x2x2 bracket the reference picks, macGrid.cpp:142-148) and only then folds it into a shared-memory
// tile with plain read-modify-writes in 9 conflict-free phases: in phase (dy,dz) warp (j,k) owns row
// (j+dy,k+dz), lanes own distinct x -- no shared-memory atomics (fp32 smem atomicAdd is a CAS loop on
// sm_100a) and a deterministic summation order inside the tile.  The tile (+1-cell halo) is flushed with
// one fp32 RED per touched node and channel; only tile-halo nodes are shared between CTAs.
#include <atomic>

#include "fsim_internal.h"

namespace {

constexpr int TX = 32, TY = 4, TZ = 2;
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

// packed fp32x2 FMA (sm_100 FFMA2): d = a * b + c on both halves; the mov.b64 packing is free after register allocation
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// Where a cell's particles are read from.  STAGED kernels copy the tile's particle runs into shared memory with coalesced
// loads first (everything in flight at once); the per-cell walk then reads shared memory.  Index i of the staged run lives
// at i + i/8: neighbouring cells start ~8 particles apart, which would put all lanes of a warp on 4 banks.
struct Src {
    const float *sx, *sy, *sz, *sv;  // staged copies (shared memory), valid for particle t < ns of this cell
    uint32_t s0, ns;                 // staged index of the cell's first particle; how many of its particles are staged
};
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i + (i >> 3); }

// One staggered face pass, AXIS = 0,1,2: 2 nodes along AXIS (the pair a packed FFMA2 works on), 3 along the other two
// dimensions B < C.  Per particle and node pair: weights += (wA0, wA1) * wB*wC ; values += (wA0*val0, wA1*val1) * wB*wC,
// val_n = v + c[AXIS] . (face_n - particle) (APIC, simulator.cpp:327-328) or v (PIC / FLIP, :324).
template <int AXIS, bool APIC, bool STAGED>
__device__ __forceinline__ void p2g_face_pass(const P2GArgs& a, const Src& src, float* s_val, float* s_w, bool active, int x, int y, int z,
                                              uint32_t start, uint32_t n, int lane, int wy, int wz) {
    constexpr int B = AXIS == 0 ? 1 : 0, C = AXIS == 2 ? 1 : 2;
    float2 aw[3][3], av[3][3];  // [c][b]
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int b = 0; b < 3; b++) { aw[c][b] = make_float2(0.f, 0.f); av[c][b] = make_float2(0.f, 0.f); }

    if (active) {
        const float* vel = AXIS == 0 ? a.vx : (AXIS == 1 ? a.vy : a.vz);
        const float hh[3] = {a.g.hx, a.g.hy, a.g.hz};
        auto body = [&](float X, float Y, float Z, float v, uint32_t p) {
            float f[3];
            f[0] = fmaf(X, a.g.ihx, -(float)x);  // single rounding
            f[1] = fmaf(Y, a.g.ihy, -(float)y);
            f[2] = fmaf(Z, a.g.ihz, -(float)(z + a.g.zoff));
            const float2 wA = make_float2(1.f - f[AXIS], f[AXIS]);
            float wB[3], wC[3];
            centred_w(f[B], wB);
            centred_w(f[C], wC);
            float2 wav = make_float2(wA.x * v, wA.y * v), P = make_float2(0.f, 0.f);
            float cB = 0.f, cC = 0.f;
            if (APIC) {
                const float cA = __ldg(a.c[3 * AXIS + AXIS] + p) * hh[AXIS];
                cB = __ldg(a.c[3 * AXIS + B] + p) * hh[B];
                cC = __ldg(a.c[3 * AXIS + C] + p) * hh[C];
                P = make_float2(wA.x * cA * (-f[AXIS]), wA.y * cA * (1.f - f[AXIS]));  // offsets to the two faces along AXIS
            }
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const float wbc = wB[b] * wC[c];
                    const float2 w2 = make_float2(wbc, wbc);
                    aw[c][b] = ffma2(wA, w2, aw[c][b]);
                    float2 tv = wav;
                    if (APIC) {
                        const float base = v + cB * ((float)b - 0.5f - f[B]) + cC * ((float)c - 0.5f - f[C]);
                        tv = ffma2(wA, make_float2(base, base), P);
                    }
                    av[c][b] = ffma2(tv, w2, av[c][b]);
                }
        };
        uint32_t t = 0;
        if (STAGED)
            for (; t < src.ns; t++) {
                const uint32_t j = swz(src.s0 + t);
                body(src.sx[j], src.sy[j], src.sz[j], src.sv[j], start + t);
            }
        for (; t < n; t++) {
            const uint32_t p = start + t;
            body(__ldg(a.px + p), __ldg(a.py + p), __ldg(a.pz + p), __ldg(vel + p), p);
        }
    }
    // fold: phases over (k, j), lanes along x.  (i,j,k) -> (pair half, b, c) depends on which dimension is staggered.
    constexpr int NX = (AXIS == 0) ? 2 : 3, NY = (AXIS == 1) ? 2 : 3, NZ = (AXIS == 2) ? 2 : 3;
#pragma unroll
    for (int k = 0; k < NZ; k++)
#pragma unroll
        for (int j = 0; j < NY; j++) {
            const int row = sidx(0, wy + j, wz + k);
#pragma unroll
            for (int i = 0; i < NX; i++) {
                const int half = AXIS == 0 ? i : (AXIS == 1 ? j : k);
                const int bb = AXIS == 0 ? j : i;
                const int cc = AXIS == 2 ? j : k;
                const float w = half ? aw[cc][bb].y : aw[cc][bb].x;
                const float vv = half ? av[cc][bb].y : av[cc][bb].x;
                const int id = row + lane + i;
                if (active && w != 0.f) { s_w[id] += w; s_val[id] += vv; }
                __syncwarp();
            }
            __syncthreads();
        }
}

// cell-centred density pass (avgPNum, simulator.cpp:362-366): 3x3x3 nodes, weights only; x nodes packed as (0,1),(2,-)
template <bool STAGED>
__device__ __forceinline__ void p2g_density_pass(const P2GArgs& a, const Src& src, float* s_w, bool active, int x, int y, int z, uint32_t start,
                                                 uint32_t n, int lane, int wy, int wz) {
    float2 aw[3][3][2];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) { aw[k][j][0] = make_float2(0.f, 0.f); aw[k][j][1] = make_float2(0.f, 0.f); }
    if (active) {
        auto body = [&](float X, float Y, float Z) {
            const float fx = fmaf(X, a.g.ihx, -(float)x);
            const float fy = fmaf(Y, a.g.ihy, -(float)y);
            const float fz = fmaf(Z, a.g.ihz, -(float)(z + a.g.zoff));
            float wx[3], wy_[3], wz_[3];
            centred_w(fx, wx); centred_w(fy, wy_); centred_w(fz, wz_);
            const float2 wx01 = make_float2(wx[0], wx[1]), wx2 = make_float2(wx[2], 0.f);
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float wyz = wy_[j] * wz_[k];
                    const float2 w2 = make_float2(wyz, wyz);
                    aw[k][j][0] = ffma2(wx01, w2, aw[k][j][0]);
                    aw[k][j][1] = ffma2(wx2, w2, aw[k][j][1]);
                }
        };
        uint32_t t = 0;
        if (STAGED)
            for (; t < src.ns; t++) {
                const uint32_t j = swz(src.s0 + t);
                body(src.sx[j], src.sy[j], src.sz[j]);
            }
        for (; t < n; t++) {
            const uint32_t p = start + t;
            body(__ldg(a.px + p), __ldg(a.py + p), __ldg(a.pz + p));
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int row = sidx(0, wy + j, wz + k);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float w = i == 0 ? aw[k][j][0].x : (i == 1 ? aw[k][j][0].y : aw[k][j][1].x);
                if (active && w != 0.f) s_w[row + lane + i] += w;
                __syncwarp();
            }
            __syncthreads();
        }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
// Read straight from global memory, the per-cell walk fetches particle p = start + t at stride ~8 floats across the lanes, so every warp-wide load
// touches 32 sectors that are only consumed over the next 8 iterations; with two CTAs per SM the 2 x 4 arrays x 16 KB
// in flight no longer fit L1 (ncu: 74 % hit rate, long-scoreboard the top stall; 1.72 ms, APIC 3.6 ms).  So every warp
// copies its own row's particle run into shared memory with coalesced loads (positions once, one velocity component per pass;
// 1.54 ms, APIC 2.1 ms) and the walk reads
// conflict-free shared memory; the accumulator tile shrinks to the two channels of the running pass and is flushed
// after every pass.  Particles beyond the staging capacity (tiles denser than 8 per cell) are read from global memory.
constexpr int CAP = 2048;               // staged particles per tile (8 per cell)
constexpr int CAPP = CAP + CAP / 8;     // with the bank padding of swz()
constexpr size_t STAGED_SMEM = (size_t)(2 * SN + 4 * CAPP) * sizeof(float);

// a warp copies its own row's particle run (NA arrays at once) into the staged slots [off, off + len).
// (Measured slower: cp.async 4-byte copies with a second velocity buffer filled one pass ahead, +12 %; caching the
// flush indices in registers across the four passes, +9 % -- the kernel has no registers to spare.)
template <int NA>
__device__ __forceinline__ void stage_row(float* const (&dst)[NA], const float* const (&src)[NA], uint32_t beg, uint32_t off, uint32_t len, int lane) {
#pragma unroll 4
    for (uint32_t k = lane; k < len; k += 32) {
        const uint32_t j = swz(off + k);
        float v[NA];
#pragma unroll
        for (int q = 0; q < NA; q++) v[q] = __ldg(src[q] + beg + k);
#pragma unroll
        for (int q = 0; q < NA; q++) dst[q][j] = v[q];
    }
    __syncwarp();
}

// Tile + halo -> global accumulators, one fp32 RED per touched node and channel.  A warp takes whole rows of the shared tile:
// lane l owns node sx = l + 1 (grid x = x0 + l: the 32 REDs of a row fall into ONE 128-byte line), lanes 0 / 1 also own the two
// halo nodes sx = 0 and sx = SX - 1.  No per-node index arithmetic (the flat loop it replaces spent two integer divisions and
// three bounds tests per node; r1 ncu: issue-bound kernel).
__device__ __forceinline__ void flush_pass(const P2GArgs& a, float* s_val, float* s_w, float* g_val, float* g_w, int x0, int y0, int z0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = warp; r < SY * SZ; r += NTHREADS / 32) {
        const int sy = r % SY, sz = r / SY;
        const int gy = y0 - 1 + sy, gz = z0 - 1 + sz;
        const bool row_ok = gy >= 0 && gy < a.g.gy && gz >= 0 && gz < a.g.gz;  // warp-uniform
        const int64_t cb = ((int64_t)gz * a.g.gy + gy) * a.g.gx + (x0 - 1);
        const int rb = r * SX;
        {
            const int sx = lane + 1;
            const float w = s_w[rb + sx];
            if (w != 0.f) {
                if (row_ok && x0 - 1 + sx < a.g.gx) {
                    atomicAdd(g_w + cb + sx, w);
                    if (g_val) atomicAdd(g_val + cb + sx, s_val[rb + sx]);
                }
                s_w[rb + sx] = 0.f;
                s_val[rb + sx] = 0.f;
            }
        }
        if (lane < 2) {
            const int sx = lane * (SX - 1);
            const float w = s_w[rb + sx];
            if (w != 0.f) {
                const int gx = x0 - 1 + sx;
                if (row_ok && gx >= 0 && gx < a.g.gx) {
                    atomicAdd(g_w + cb + sx, w);
                    if (g_val) atomicAdd(g_val + cb + sx, s_val[rb + sx]);
                }
                s_w[rb + sx] = 0.f;
                s_val[rb + sx] = 0.f;
            }
        }
    }
    __syncthreads();
}

// MacGrid::resetGridValues for the seven scatter channels in ONE launch (seven cudaMemsetAsync calls of 67 MB each cost
// 0.125 ms at 256^3: launch gaps and ramp-up seven times over)
struct ZeroArgs { float* p[7]; int64_t n4; };  // n4 = floats / 4 (the tail is zeroed by a memset)
__global__ void __launch_bounds__(256) zero7_kernel(ZeroArgs a) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 7; k++) reinterpret_cast<float4*>(a.p[k])[i] = z;
    }
}

template <bool APIC>
__global__ void __launch_bounds__(NTHREADS, 4) p2g_kernel(P2GArgs a) {
    extern __shared__ float dyn[];
    __shared__ uint32_t row_beg[TY * TZ], row_off[TY * TZ + 1];
    float* s_val = dyn;
    float* s_w = dyn + SN;
    float* sp = dyn + 2 * SN;  // x | y | z | v, CAPP floats each
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

    // the warp's row of cells is one contiguous particle run [beg, end)
    {
        const int last = min(31, a.g.gx - 1 - x0);
        uint32_t beg = __shfl_sync(0xffffffffu, start, 0), end = __shfl_sync(0xffffffffu, start + n, last);
        if (!(y < a.g.gy && z < a.g.gz)) beg = end = 0;
        if (lane == 0) { row_beg[warp] = beg; row_off[warp + 1] = end - beg; }
    }
    for (int i = threadIdx.x; i < 2 * SN; i += NTHREADS) dyn[i] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int r = 0; r < TY * TZ; r++) { const uint32_t len = row_off[r + 1]; row_off[r] = acc; acc += len; }
        row_off[TY * TZ] = acc;
    }
    __syncthreads();
    const uint32_t r_beg = row_beg[warp], r_off = row_off[warp];
    const uint32_t r_len = min(row_off[warp + 1], (uint32_t)CAP) - min(r_off, (uint32_t)CAP);  // staged part of the warp's row
    Src src;
    src.sx = sp; src.sy = sp + CAPP; src.sz = sp + 2 * CAPP; src.sv = sp + 3 * CAPP;
    src.s0 = row_off[warp] + (start - row_beg[warp]);
    src.ns = n == 0 ? 0u : (src.s0 >= (uint32_t)CAP ? 0u : min(n, (uint32_t)CAP - src.s0));
    {
        float* const d4[4] = {sp, sp + CAPP, sp + 2 * CAPP, sp + 3 * CAPP};
        const float* const g4[4] = {a.px, a.py, a.pz, a.vx};
        stage_row<4>(d4, g4, r_beg, r_off, r_len, lane);
    }
    float* const d1[1] = {sp + 3 * CAPP};

    const bool active = n > 0;
    p2g_face_pass<0, APIC, true>(a, src, s_val, s_w, active, x, y, z, start, n, lane, wy, wz);
    { const float* const g1[1] = {a.vy}; stage_row<1>(d1, g1, r_beg, r_off, r_len, lane); }  // every thread is past its vx reads (the fold ends with a barrier)
    flush_pass(a, s_val, s_w, a.su[0], a.wu[0], x0, y0, z0);
    p2g_face_pass<1, APIC, true>(a, src, s_val, s_w, active, x, y, z, start, n, lane, wy, wz);
    { const float* const g1[1] = {a.vz}; stage_row<1>(d1, g1, r_beg, r_off, r_len, lane); }
    flush_pass(a, s_val, s_w, a.su[1], a.wu[1], x0, y0, z0);
    p2g_face_pass<2, APIC, true>(a, src, s_val, s_w, active, x, y, z, start, n, lane, wy, wz);
    flush_pass(a, s_val, s_w, a.su[2], a.wu[2], x0, y0, z0);
    p2g_density_pass<true>(a, src, s_w, active, x, y, z, start, n, lane, wy, wz);
    flush_pass(a, s_val, s_w, nullptr, a.dens, x0, y0, z0);
}

}  // namespace

int k_p2g(fsim* h) {
    const GridDims& g = h->g;
    // MacGrid::resetGridValues (macGrid.cpp:221-229): v, v2, weights, avgPNum = 0 (type is rewritten by classify)
    {
        KScope ks(h, K_MEMSET, 1);
        ZeroArgs z;
        for (int a = 0; a < 3; a++) { z.p[a] = h->u[a]; z.p[3 + a] = h->wsum[a]; }
        z.p[6] = h->dens;
        z.n4 = g.nc / 4;
        if (z.n4 > 0) zero7_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(z);
        const int64_t tail = g.nc - 4 * z.n4;
        if (tail > 0)
            for (int k = 0; k < 7; k++) FSIM_CUDA(h, cudaMemsetAsync(z.p[k] + 4 * z.n4, 0, sizeof(float) * tail, h->stream));
    }
    if (h->np == 0) return FSIM_OK;  // (slab mode: np is current here, k_sort read it back)
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
    {
        // the opt-in shared-memory size is a per-device function attribute: set it once per device, not per launch
        // (slab groups drive several handles from several host threads: the flags are atomic, setting the attribute twice is harmless)
        static std::atomic<bool> attr_set[2][64];
        const int dev = h->device & 63;
        if (!attr_set[a.apic ? 1 : 0][dev].load(std::memory_order_acquire)) {
            if (a.apic) FSIM_CUDA(h, cudaFuncSetAttribute(p2g_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STAGED_SMEM));
            else FSIM_CUDA(h, cudaFuncSetAttribute(p2g_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STAGED_SMEM));
            attr_set[a.apic ? 1 : 0][dev].store(true, std::memory_order_release);
        }
        KScope ks(h, K_P2G);
        if (a.apic) p2g_kernel<true><<<grid, NTHREADS, STAGED_SMEM, h->stream>>>(a);
        else p2g_kernel<false><<<grid, NTHREADS, STAGED_SMEM, h->stream>>>(a);
    }
    FSIM_CHECK_LAUNCH(h);
    return FSIM_OK;
}

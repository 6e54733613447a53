// tile4.cuh -- iteration space of the four-cells-per-thread solver kernels that end in a grid-wide reduction (spmv4,
// update + first sweep, direction): 2-D blocks of 128 x-cells by 32 rows (row = z * gy + y), a warp per row, four trips.
//
// These kernels used to walk LINEAR chunks of 8192 cells.  With a 256-cell row a warp then owned half a row, and in the
// dam-break scenes (fluid in x < N/2) every block was half fluid, half air: the air warps finished at once and waited at the
// block reduction while the fluid warps did all the work (ncu r2: 43 % of spmv4's stall samples at that barrier, 35-45 %
// active warps).  A block that is 128 cells wide is all fluid or all air in such scenes; air blocks retire after their code
// loads and the SM takes the next block.
#pragma once
#include <stdint.h>

constexpr int T4_XCELLS = 128;  // 32 lanes x 4 cells
constexpr int T4_ROWS = 32;     // 8 warps x 4 trips
constexpr int T4_TRIPS = T4_ROWS / 8;

struct Tile4 {
    int nbx, gx;          // blocks along x, cells per row
    int64_t row0, nrows;  // rows [row0, row0 + nrows) are visited (the whole grid, or the planes a slab rank owns)
    int rot, nrb;         // row-blocks are visited in the order (i + rot) mod nrb (fused slab exchanges: boundary planes first / last)
    int gy;               // rows per plane (tile4_plane)
};

static inline Tile4 tile4_make(int gx, int64_t row0, int64_t nrows, int gy = 0, int rot = 0) {
    Tile4 t;
    t.nbx = (gx + T4_XCELLS - 1) / T4_XCELLS; t.gx = gx; t.row0 = row0; t.nrows = nrows;
    t.nrb = (int)((nrows + T4_ROWS - 1) / T4_ROWS); t.rot = t.nrb > 0 ? rot % t.nrb : 0; t.gy = gy;
    return t;
}
// the same rows, another visiting order: boundary planes of a slab first (producers of a fused exchange) or last (consumers)
static inline Tile4 tile4_boundary_first(Tile4 t) { const int pp = t.gy / T4_ROWS; t.rot = t.nrb > pp ? t.nrb - pp : 0; return t; }
static inline Tile4 tile4_boundary_last(Tile4 t) { const int pp = t.gy / T4_ROWS; t.rot = t.nrb > pp ? pp : 0; return t; }
static inline int tile4_blocks(const Tile4& t) { return (int)(t.nbx * ((t.nrows + T4_ROWS - 1) / T4_ROWS)); }

#ifdef __CUDACC__
// first cell of the thread's group in trip `trip`; false outside the grid (gx % 4 == 0: a group never straddles the edge)
__device__ __forceinline__ bool tile4_cell(const Tile4& t, int trip, int64_t& c) {
    const int bx = blockIdx.x % t.nbx;
    int rbi = blockIdx.x / t.nbx + t.rot;
    if (rbi >= t.nrb) rbi -= t.nrb;
    const int64_t rb = rbi;
    const int x = bx * T4_XCELLS + (threadIdx.x & 31) * 4;
    const int64_t row = rb * T4_ROWS + trip * 8 + (threadIdx.x >> 5);
    c = (t.row0 + row) * t.gx + x;
    return x < t.gx && row < t.nrows;
}
// plane of the block's rows (gy % T4_ROWS == 0: a block never straddles two planes)
__device__ __forceinline__ int tile4_plane(const Tile4& t) {
    int rbi = blockIdx.x / t.nbx + t.rot;
    if (rbi >= t.nrb) rbi -= t.nrb;
    return (int)((t.row0 + (int64_t)rbi * T4_ROWS) / t.gy);
}
#endif

/*
 * fsim_oracle.c -- TEST INFRASTRUCTURE ONLY, not part of the product.
 *
 * Plain-C, double-precision, single-threaded restatement of the reference's per-step hot path
 * (genericfsim::simulator::Simulator::simulate and callees).  It is the checker for the CUDA
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * Every function cites the reference lines (relative to /root/reference/src/Simulator/simulator)
 * whose arithmetic and operation ORDER it follows, so that it reproduces the reference's serial
 * (-D_DEBUG) build bit for bit.
 *
 * Pinning: the reference ships no tests or golden vectors, so this oracle is
 * pinned against outputs of the reference itself: tests/test_oracle_pin.py compares every stage
 * with oracle/_ref/libfsim_ref_serial.so (the unmodified reference sources compiled by
 * oracle/build_ref.sh) when that library is present, and against the committed fixtures in
 * tests/golden/ (generated from the same library by tests/golden/make_golden.py) always.
 *
 * Layout: reference order, cell (x,y,z) at x*gy*gz + y*gz + z (macGrid.h:66-93).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fsim.h"

typedef struct { double x, y, z; } d3;
typedef struct { int x, y, z; } i3;

typedef struct {
    d3 pos, v, c[3];
} Particle; /* particles/particle.h:8-12 */

typedef struct Oracle {
    /* MacGrid public consts (macGrid.h:168-182) */
    d3 cellD, cellDInv, dimensions;
    i3 gs;
    int twoD;
    int64_t nc;
    int yz; /* yzMultiplier */
    /* per-cell state (macGridCell.h:8-37) */
    uint8_t* type;
    double *v, *v2, *wsum; /* [nc*3] faces[a].v / v2 / particleWeightSum */
    double* avgp;
    int* id;
    i3* fluid; /* fluidCellPositions */
    int nfluid;
    double* pressure; /* per fluid cell; exported per cell by oracle_get_grid */
    int pressure_valid;
    /* HashedParticles */
    Particle* p;
    int64_t np, cap;
    double r, zval;
    int zConst;
    FsimParams par;
    FsimObstacle obs[FSIM_MAX_OBSTACLES];
    int nobs;
    FsimSolveInfo solve;
} Oracle;

static d3 mk(double x, double y, double z) { d3 r = {x, y, z}; return r; }
static d3 add(d3 a, d3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static d3 sub(d3 a, d3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static d3 mul(d3 a, d3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
static d3 scl(d3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
static d3 dvs(d3 a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
static double dot3(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static double len3(d3 a) { return sqrt(dot3(a, a)); }
static d3 norm3(d3 a) { return scl(a, 1.0 / sqrt(dot3(a, a))); }
static double getc(d3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static void setc(d3* a, int i, double v) { if (i == 0) a->x = v; else if (i == 1) a->y = v; else a->z = v; }
static double clampd(double v, double lo, double hi) { return v < lo ? lo : (hi < v ? hi : v); } /* std::clamp */
static double dmax(double a, double b) { return a < b ? b : a; }                                /* std::max */
static double dmin(double a, double b) { return b < a ? b : a; }                                /* std::min */

static int64_t cidx(const Oracle* o, int x, int y, int z) { return (int64_t)x * o->yz + (int64_t)y * o->gs.z + z; }
static d3 face_pos(const Oracle* o, int x, int y, int z, int a) { /* macGrid.cpp:25-27 */
    switch (a) {
        case 0: return mk((x + 1) * o->cellD.x, (y + 0.5) * o->cellD.y, (z + 0.5) * o->cellD.z);
        case 1: return mk((x + 0.5) * o->cellD.x, (y + 1) * o->cellD.y, (z + 0.5) * o->cellD.z);
        default: return mk((x + 0.5) * o->cellD.x, (y + 0.5) * o->cellD.y, (z + 1) * o->cellD.z);
    }
}
static d3 cell_pos(const Oracle* o, int x, int y, int z) { /* macGrid.cpp:28 */
    return mul(mk(x + 0.5, y + 0.5, z + 0.5), o->cellD);
}

/* util/interpolation.h:7-18 */
static double trilinear(d3 center, d3 pos, d3 inv) {
    d3 vec = mul(sub(pos, center), inv);
    double x = (1.0 - fabs(vec.x)), y = (1.0 - fabs(vec.y)), z = (1.0 - fabs(vec.z));
    return x * y * z;
}
/* util/interpolation.h:21-35 */
static d3 trilinear_grad(d3 center, d3 pos, d3 inv) {
    d3 vec = mul(sub(pos, center), inv);
    double xa = (1.0 - fabs(vec.x)), ya = (1.0 - fabs(vec.y)), za = (1.0 - fabs(vec.z));
    double xs = vec.x > 0.0 ? -1.0 : 1.0, ys = vec.y > 0.0 ? -1.0 : 1.0, zs = vec.z > 0.0 ? -1.0 : 1.0;
    return mul(mk(xs * ya * za, ys * xa * za, zs * xa * ya), inv);
}

/* macGrid.cpp:75-140 (serial branch; the parallel branch writes the same values) */
static void restore_borders(Oracle* o) {
    const i3 g = o->gs;
    for (int x = 0; x < g.x; x++)
        for (int y = 0; y < g.y; y++) {
            o->type[cidx(o, x, y, 0)] = o->type[cidx(o, x, y, g.z - 1)] = FSIM_CELL_SOLID;
            if (o->type[cidx(o, x, y, 1)] == FSIM_CELL_WATER) o->v[3 * cidx(o, x, y, 0) + 2] = 0;
            if (o->type[cidx(o, x, y, g.z - 2)] == FSIM_CELL_WATER) o->v[3 * cidx(o, x, y, g.z - 2) + 2] = 0;
        }
    for (int y = 0; y < g.y; y++)
        for (int z = 0; z < g.z; z++) {
            o->type[cidx(o, 0, y, z)] = o->type[cidx(o, g.x - 1, y, z)] = FSIM_CELL_SOLID;
            if (o->type[cidx(o, 1, y, z)] == FSIM_CELL_WATER) o->v[3 * cidx(o, 0, y, z) + 0] = 0;
            if (o->type[cidx(o, g.x - 2, y, z)] == FSIM_CELL_WATER) o->v[3 * cidx(o, g.x - 2, y, z) + 0] = 0;
        }
    for (int x = 0; x < g.x; x++)
        for (int z = 0; z < g.z; z++) {
            o->type[cidx(o, x, 0, z)] = FSIM_CELL_SOLID;
            if (o->type[cidx(o, x, 1, z)] == FSIM_CELL_WATER) o->v[3 * cidx(o, x, 0, z) + 1] = 0;
            if (o->par.top_solid) {
                o->type[cidx(o, x, g.y - 1, z)] = FSIM_CELL_SOLID;
                if (o->type[cidx(o, x, g.y - 2, z)] == FSIM_CELL_WATER) o->v[3 * cidx(o, x, g.y - 2, z) + 1] = 0;
            }
        }
}

/* MacGrid ctor macGrid.cpp:9-17 + BridsonSolverGrid ctor (resolution is a float, bridsonSolverGrid.cpp:8)
 * + HashedParticles ctor members (hashedParticles.cpp:23-24) as SimulationManager wires them
 * (manager/simulationManager.cpp:16-28) */
Oracle* oracle_create(const FsimGridDesc* d) {
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    const double res = (double)(float)d->resolution;
    const d3 td = mk(d->target_dims[0], d->target_dims[1], d->target_dims[2]);
    o->twoD = d->two_d != 0;
    o->cellD = mk(1 / res, 1 / res, o->twoD ? td.z / 3 : 1 / res);
    o->cellDInv = mk(1.0 / o->cellD.x, 1.0 / o->cellD.y, 1.0 / o->cellD.z);
    o->gs.x = (int)(td.x / o->cellD.x);
    o->gs.y = (int)(td.y / o->cellD.y);
    o->gs.z = o->twoD ? 3 : (int)(td.z / o->cellD.z);
    o->dimensions = mk(o->gs.x * o->cellD.x, o->gs.y * o->cellD.y, o->twoD ? td.z : o->gs.z * o->cellD.z);
    o->yz = o->gs.y * o->gs.z;
    o->nc = (int64_t)o->gs.x * o->gs.y * o->gs.z;
    o->type = (uint8_t*)malloc(o->nc);
    memset(o->type, FSIM_CELL_AIR, o->nc);
    o->v = (double*)calloc(o->nc * 3, sizeof(double));
    o->v2 = (double*)calloc(o->nc * 3, sizeof(double));
    o->wsum = (double*)calloc(o->nc * 3, sizeof(double));
    o->avgp = (double*)calloc(o->nc, sizeof(double));
    o->id = (int*)calloc(o->nc, sizeof(int));
    o->fluid = (i3*)malloc(o->nc * sizeof(i3));
    o->pressure = (double*)calloc(o->nc, sizeof(double));
    o->r = d->particle_radius;
    o->zConst = o->twoD;
    o->zval = td.z / 2;
    /* MacGrid public defaults (macGrid.h:173-179) and SimulatorConfig defaults (simulator.h:30-39) */
    o->par.transfer_type = FSIM_TRANSFER_FLIP; o->par.flip_ratio = 0.99f; o->par.gravity = 150.0f;
    o->par.gravity_enabled = 1; o->par.push_apart_enabled = 1; o->par.pressure_enabled = 1;
    o->par.max_iterations = 80; o->par.pressure_k = 2.0; o->par.average_pressure = 2.0;
    o->par.fluid_density = 1.0; o->par.residual_tolerance = 1e-6;
    restore_borders(o);
    return o;
}

void oracle_destroy(Oracle* o) {
    if (!o) return;
    free(o->type); free(o->v); free(o->v2); free(o->wsum); free(o->avgp); free(o->id); free(o->fluid);
    free(o->pressure); free(o->p); free(o);
}

void oracle_get_grid_info(const Oracle* o, FsimGridInfo* g) {
    g->cell_d[0] = o->cellD.x; g->cell_d[1] = o->cellD.y; g->cell_d[2] = o->cellD.z;
    g->cell_d_inv[0] = o->cellDInv.x; g->cell_d_inv[1] = o->cellDInv.y; g->cell_d_inv[2] = o->cellDInv.z;
    g->grid_size[0] = o->gs.x; g->grid_size[1] = o->gs.y; g->grid_size[2] = o->gs.z;
    g->dimensions[0] = o->dimensions.x; g->dimensions[1] = o->dimensions.y; g->dimensions[2] = o->dimensions.z;
    g->two_d = o->twoD; g->cell_count = o->nc;
}

void oracle_set_params(Oracle* o, const FsimParams* p) { o->par = *p; }
void oracle_set_obstacles(Oracle* o, const FsimObstacle* obs, int n) {
    o->nobs = n > FSIM_MAX_OBSTACLES ? FSIM_MAX_OBSTACLES : n;
    memcpy(o->obs, obs, sizeof(FsimObstacle) * o->nobs);
}
void oracle_get_obstacles(const Oracle* o, FsimObstacle* out, int n) {
    for (int i = 0; i < n && i < o->nobs; i++) out[i].last_spawn_fraction = o->obs[i].last_spawn_fraction;
}

static void reserve(Oracle* o, int64_t n) {
    if (n <= o->cap) return;
    int64_t c = o->cap ? o->cap : 1024;
    while (c < n) c *= 2;
    o->p = (Particle*)realloc(o->p, c * sizeof(Particle));
    o->cap = c;
}
/* HashedParticles::addParticles hashedParticles.cpp:153-156 */
void oracle_append_particles(Oracle* o, const double* aos15, int64_t n) {
    reserve(o, o->np + n);
    memcpy(o->p + o->np, aos15, n * sizeof(Particle));
    o->np += n;
}
void oracle_set_particles(Oracle* o, const double* aos15, int64_t n) { o->np = 0; oracle_append_particles(o, aos15, n); }
int64_t oracle_particle_count(const Oracle* o) { return o->np; }
void oracle_get_particles(const Oracle* o, double* aos15, int64_t cap) {
    memcpy(aos15, o->p, (cap < o->np ? cap : o->np) * sizeof(Particle));
}
/* simulator.cpp:358-359: ivec3(particle.pos * cellDInv), flattened in reference order */
void oracle_get_particle_cells(const Oracle* o, int32_t* out, int64_t cap) {
    for (int64_t i = 0; i < o->np && i < cap; i++) {
        d3 g = mul(o->p[i].pos, o->cellDInv);
        out[i] = (int32_t)cidx(o, (int)g.x, (int)g.y, (int)g.z);
    }
}

static int cmp_int(const void* a, const void* b) { return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b); }
/* HashedParticles::removeParticles hashedParticles.cpp:158-172 */
void oracle_remove_particles(Oracle* o, int* ids, int64_t n) {
    qsort(ids, n, sizeof(int), cmp_int);
    int64_t k = 0;
    for (int64_t p = 0; p < o->np; p++) {
        if (k < n && p == ids[k]) { k++; continue; }
        o->p[p - k] = o->p[p];
    }
    o->np -= n;
}

/* util/random.h:13-26 */
static double rnd_range(double lo, double hi) { return lo + (hi - lo) * ((double)rand() / RAND_MAX); }
void oracle_srand(unsigned s) { srand(s); }

/* Simulator::spawnParticles simulator.cpp:106-125 */
void oracle_stage_spawn(Oracle* o, double dt) {
    for (int k = 0; k < o->nobs; k++) {
        FsimObstacle* ob = &o->obs[k];
        if (ob->kind != FSIM_OBSTACLE_SOURCE) continue;
        double r = ob->r + o->r;
        double numD = ob->spawn_rate * dt + ob->last_spawn_fraction;
        int num = (int)numD;
        ob->last_spawn_fraction = numD - num;
        for (int i = 0; i < num; i++) {
            double theta = rnd_range(0.0, 2.0 * M_PI);
            double phi = rnd_range(0.0, M_PI);
            d3 normal = mk(r * sin(phi) * cos(theta), r * sin(phi) * sin(theta), r * cos(phi));
            Particle q;
            memset(&q, 0, sizeof(q));
            q.pos = add(mk(ob->pos[0], ob->pos[1], ob->pos[2]), normal);
            q.v = scl(norm3(normal), ob->spawn_speed); /* double * dvec3 */
            reserve(o, o->np + 1);
            o->p[o->np++] = q;
        }
    }
}

/* isParticleInRectangle simulator.cpp:127-132 */
static int in_rect(d3 pPos, double pr, d3 rPos, d3 rSize) {
    d3 d = sub(pPos, rPos);
    return (fabs(d.x) < rSize.x * 0.5 + pr) && (fabs(d.y) < rSize.y * 0.5 + pr) && (fabs(d.z) < rSize.z * 0.5 + pr);
}

/* Simulator::advectParticles simulator.cpp:134-251 */
void oracle_stage_advect(Oracle* o, double dt) {
    const double wallRestitution = 0.3, sphereRestitution = 1.0, rectangleRestitution = 0.2;
    const double pr = o->r;
    const d3 gridLow = add(o->cellD, scl(mk(pr, pr, o->twoD ? 0.0 : pr), 1.01));
    const d3 gridHigh = sub(o->dimensions, gridLow);
    int* removeIds = (int*)malloc(sizeof(int) * (o->np > 0 ? o->np : 1));
    int64_t nRemove = 0;
    for (int64_t idx = 0; idx < o->np; idx++) {
        Particle* P = &o->p[idx];
        double t = 0;
        int run = 0;
        const int maxRunCount = 200;
        while (run < maxRunCount) {
            run++;
            double minT = 1e6;
            int minAxis = 0;
            for (int axis = 0; axis < 3; axis++) {
                double comp = getc(P->v, axis), tmp = 1e6;
                if (comp > 1e-6) tmp = (getc(gridHigh, axis) - getc(P->pos, axis)) / comp;
                else if (comp < -1e-6) tmp = (getc(P->pos, axis) - getc(gridLow, axis)) / -comp;
                if (tmp < minT) { minT = tmp; minAxis = axis; }
            }
            if (minT <= (dt - t)) {
                P->pos = add(P->pos, scl(scl(P->v, minT), 0.999));
                setc(&P->v, minAxis, getc(P->v, minAxis) * -wallRestitution);
                t += 0.999 * minT;
                continue;
            }
            int collision = 0;
            P->pos = add(P->pos, scl(P->v, (dt - t)));
            for (int k = 0; k < o->nobs; k++) {
                const FsimObstacle* ob = &o->obs[k];
                const d3 opos = mk(ob->pos[0], ob->pos[1], ob->pos[2]);
                const d3 ospeed = mk(ob->speed[0], ob->speed[1], ob->speed[2]);
                if (ob->kind != FSIM_OBSTACLE_BOX) {
                    const int isSink = ob->kind == FSIM_OBSTACLE_SINK;
                    double r = ob->r + pr;
                    double d = len3(sub(opos, P->pos)) - r;
                    if (d < 0) {
                        if (isSink && o->par.despawning_enabled) {
                            removeIds[nRemove++] = (int)idx;
                            collision = 0;
                            break;
                        }
                        double backTime = 0;
                        while (len3(sub(sub(P->pos, scl(P->v, backTime)), sub(opos, scl(ospeed, backTime)))) < r &&
                               backTime < t + 0.001f)
                            backTime += 0.0002;
                        if (backTime > dt - t) continue;
                        backTime += 0.0002;
                        P->pos = sub(P->pos, scl(P->v, backTime));
                        d3 posTmp = sub(opos, scl(ospeed, backTime));
                        d3 normal = norm3(sub(P->pos, posTmp));
                        d3 relV = sub(P->v, ospeed);
                        double ssn = -dot3(normal, relV);
                        if (ssn <= 0.0f) continue;
                        d3 mirror = dvs(mk(-relV.x, -relV.y, -relV.z), ssn);
                        P->v = add(scl(sub(normal, mirror), 2.0), mirror);
                        P->v = scl(P->v, ssn * sphereRestitution);
                        P->v = add(P->v, ospeed);
                        t += (dt - t) - backTime;
                        collision = 1;
                        break;
                    }
                } else {
                    const d3 osize = mk(ob->size[0], ob->size[1], ob->size[2]);
                    if (in_rect(P->pos, pr, opos, osize)) {
                        double backTime = 0;
                        while (in_rect(sub(P->pos, scl(P->v, backTime)), pr, sub(opos, scl(ospeed, backTime)), osize) &&
                               backTime < dt - t)
                            backTime += 0.0002;
                        if (backTime > dt - t) continue;
                        backTime += 0.0002;
                        P->pos = sub(P->pos, scl(P->v, backTime));
                        d3 posTmp = sub(opos, scl(ospeed, backTime));
                        for (int axis = 0; axis < 3; axis++) {
                            if (getc(P->pos, axis) >= getc(posTmp, axis) + getc(osize, axis) * 0.5 + pr) {
                                setc(&P->v, axis, -getc(P->v, axis) * rectangleRestitution + getc(ospeed, axis));
                                break;
                            } else if (getc(P->pos, axis) <= getc(posTmp, axis) - getc(osize, axis) * 0.5 - pr) {
                                setc(&P->v, axis, -getc(P->v, axis) * rectangleRestitution - getc(ospeed, axis));
                                break;
                            }
                        }
                        t += (dt - t) - backTime;
                        collision = 1;
                        break;
                    }
                }
            }
            if (!collision) run = maxRunCount;
        }
        P->pos.x = clampd(P->pos.x, gridLow.x, gridHigh.x);
        P->pos.y = clampd(P->pos.y, gridLow.y, gridHigh.y);
        P->pos.z = clampd(P->pos.z, gridLow.z, gridHigh.z);
    }
    oracle_remove_particles(o, removeIds, nRemove);
    free(removeIds);
}

/* MacGrid::getMinMaxRect macGrid.h:139-145 */
static void min_max_rect(const Oracle* o, d3 pos, d3 size, d3* lo, d3* hi) {
    const d3 center = mul(pos, o->cellDInv);
    const d3 d = scl(mul(size, o->cellDInv), 0.5);
    const i3 mn = {(int)dmax(round(center.x - d.x), 1.0), (int)dmax(round(center.y - d.y), 1.0), (int)dmax(round(center.z - d.z), 1.0)};
    const i3 mx = {(int)dmin(round(center.x + d.x), o->gs.x - 1.0), (int)dmin(round(center.y + d.y), o->gs.y - 1.0),
                   (int)dmin(round(center.z + d.z), o->gs.z - 1.0)};
    *lo = mul(mk(mn.x, mn.y, mn.z), o->cellD);
    *hi = mul(mk(mx.x, mx.y, mx.z), o->cellD);
}

/* Simulator::pushParticlesOutOfObstacles simulator.cpp:253-312 */
void oracle_stage_push_out(Oracle* o) {
    const double pr = o->r;
    const d3 low = mk(o->cellD.x + pr * 1.01, o->cellD.y + pr * 1.01, o->cellD.z + pr * 1.01);
    const d3 high = sub(o->dimensions, low);
    for (int k = 0; k < o->nobs; k++) {
        const FsimObstacle* ob = &o->obs[k];
        const d3 opos = mk(ob->pos[0], ob->pos[1], ob->pos[2]);
        if (ob->kind != FSIM_OBSTACLE_BOX) {
            double r = ob->r + pr, r2 = r * r;
            for (int64_t i = 0; i < o->np; i++) {
                Particle* P = &o->p[i];
                d3 o2p = sub(P->pos, opos);
                double d2 = dot3(o2p, o2p);
                if (d2 >= r2) continue;
                double dist = sqrt(d2);
                P->pos = add(opos, scl(dvs(o2p, dist), r));
                P->pos.x = clampd(P->pos.x, low.x, high.x);
                P->pos.y = clampd(P->pos.y, low.y, high.y);
                P->pos.z = o->zConst ? o->zval : clampd(P->pos.z, low.z, high.z);
            }
        } else {
            d3 start, end;
            min_max_rect(o, opos, mk(ob->size[0], ob->size[1], ob->size[2]), &start, &end);
            start = sub(start, mk(pr, pr, pr));
            end = add(end, mk(pr, pr, pr));
            for (int64_t i = 0; i < o->np; i++) {
                Particle* P = &o->p[i];
                if ((P->pos.x > end.x || P->pos.x < start.x) || (P->pos.y > end.y || P->pos.y < start.y) ||
                    ((P->pos.z > end.z || P->pos.z < start.z) && !o->zConst))
                    continue;
                int axis = 0;
                double amount = 1e6;
                int axisNum = o->zConst ? 2 : 3;
                for (int a = 0; a < axisNum; a++) {
                    if (getc(P->pos, a) > getc(opos, a)) {
                        double tmp = getc(end, a) - getc(P->pos, a);
                        if (tmp < fabs(amount)) { amount = tmp; axis = a; }
                    } else {
                        double tmp = getc(start, a) - getc(P->pos, a);
                        if (-tmp < fabs(amount)) { amount = tmp; axis = a; }
                    }
                }
                setc(&P->pos, axis, getc(P->pos, axis) + amount);
                setc(&P->pos, axis, clampd(getc(P->pos, axis), getc(low, axis), getc(high, axis)));
            }
        }
    }
}

void oracle_stage_stop(Oracle* o) { /* simulator.cpp:71-74 */
    for (int64_t i = 0; i < o->np; i++) o->p[i].v = mk(0, 0, 0);
}

/* corner order of getFacesAround / getCellsAround (macGrid.cpp:150-201) */
static const int CORNER[8][3] = {{1, 1, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}, {1, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 0, 0}};

/* MacGrid::getFacesAround base coordinate for one axis (macGrid.cpp:142-148) */
static i3 face_base(const Oracle* o, d3 pos, int axis) {
    d3 off = mk(axis == 0 ? 0.0 : 0.5, axis == 1 ? 0.0 : 0.5, axis == 2 ? 0.0 : 0.5);
    d3 gp = sub(mul(pos, o->cellDInv), off);
    i3 c = {(int)gp.x, (int)gp.y, (int)gp.z};
    if (axis == 0) c.x -= 1; else if (axis == 1) c.y -= 1; else c.z -= 1;
    return c;
}

/* MacGrid::resetGridValues macGrid.cpp:221-229 + Simulator::p2gTransfer simulator.cpp:314-352 */
void oracle_stage_p2g(Oracle* o) {
    memset(o->v, 0, sizeof(double) * 3 * o->nc);
    memset(o->v2, 0, sizeof(double) * 3 * o->nc);
    memset(o->wsum, 0, sizeof(double) * 3 * o->nc);
    memset(o->avgp, 0, sizeof(double) * o->nc);
    memset(o->type, FSIM_CELL_AIR, o->nc);
    const int apic = o->par.transfer_type == FSIM_TRANSFER_APIC;
    for (int64_t i = 0; i < o->np; i++) {
        const Particle* P = &o->p[i];
        for (int axis = 0; axis < 3; axis++) {
            i3 b = face_base(o, P->pos, axis);
            for (int k = 0; k < 8; k++) {
                int x = b.x + CORNER[k][0], y = b.y + CORNER[k][1], z = b.z + CORNER[k][2];
                int64_t ci = cidx(o, x, y, z);
                d3 fp = face_pos(o, x, y, z, axis);
                double w = trilinear(fp, P->pos, o->cellDInv);
                if (!apic) {
                    o->v[3 * ci + axis] += getc(P->v, axis) * w;
                } else {
                    d3 p2f = sub(fp, P->pos);
                    o->v[3 * ci + axis] += (getc(P->v, axis) + dot3(P->c[axis], p2f)) * w;
                }
                o->wsum[3 * ci + axis] += w;
            }
        }
    }
    for (int64_t c = 0; c < o->nc; c++)
        for (int a = 0; a < 3; a++) {
            double w = o->wsum[3 * c + a];
            if (w > 1e-6) o->v[3 * c + a] = o->v[3 * c + a] / w;
            else o->v[3 * c + a] = 0.0;
        }
}

/* Simulator::markFluidCellsAndCalculateParticleDensities simulator.cpp:354-368 */
static void mark_fluid(Oracle* o) {
    for (int64_t i = 0; i < o->np; i++) {
        const Particle* P = &o->p[i];
        d3 g = mul(P->pos, o->cellDInv);
        o->type[cidx(o, (int)g.x, (int)g.y, (int)g.z)] = FSIM_CELL_WATER;
        d3 gp = sub(mul(P->pos, o->cellDInv), mk(0.5, 0.5, 0.5)); /* getCellsAround macGrid.cpp:188-190 */
        i3 b = {(int)gp.x, (int)gp.y, (int)gp.z};
        for (int k = 0; k < 8; k++) {
            int x = b.x + CORNER[k][0], y = b.y + CORNER[k][1], z = b.z + CORNER[k][2];
            o->avgp[cidx(o, x, y, z)] += trilinear(cell_pos(o, x, y, z), P->pos, o->cellDInv);
        }
    }
}

/* the six neighbour face updates of MacGrid::addObstacle (macGrid.cpp:247-259 == 275-287) */
static void solid_cell(Oracle* o, int x, int y, int z, d3 speed) {
    const int64_t c = cidx(o, x, y, z);
    o->type[c] = FSIM_CELL_SOLID;
    if (o->type[cidx(o, x + 1, y, z)] == FSIM_CELL_WATER) o->v[3 * c + 0] = speed.x;
    if (o->type[cidx(o, x - 1, y, z)] == FSIM_CELL_WATER) o->v[3 * cidx(o, x - 1, y, z) + 0] = speed.x;
    if (o->type[cidx(o, x, y + 1, z)] == FSIM_CELL_WATER) o->v[3 * c + 1] = speed.y;
    if (o->type[cidx(o, x, y - 1, z)] == FSIM_CELL_WATER) o->v[3 * cidx(o, x, y - 1, z) + 1] = speed.y;
    if (o->type[cidx(o, x, y, z + 1)] == FSIM_CELL_WATER) o->v[3 * c + 2] = speed.z;
    if (o->type[cidx(o, x, y, z - 1)] == FSIM_CELL_WATER) o->v[3 * cidx(o, x, y, z - 1) + 2] = speed.z;
}

/* MacGrid::addObstacle macGrid.cpp:231-292 */
static void add_obstacle(Oracle* o, const FsimObstacle* ob) {
    const d3 center = mul(mk(ob->pos[0], ob->pos[1], ob->pos[2]), o->cellDInv);
    const d3 speed = mk(ob->speed[0], ob->speed[1], ob->speed[2]);
    if (ob->kind == FSIM_OBSTACLE_SPHERE || ob->kind == FSIM_OBSTACLE_SOURCE) {
        const double r = o->cellDInv.x * ob->r, r2 = r * r;
        const i3 mn = {(int)dmax(center.x - r, 1.0), (int)dmax(center.y - r, 1.0), (int)dmax(center.z - r, 1.0)};
        const i3 mx = {(int)dmin(center.x + r, o->gs.x - 2.0), (int)dmin(center.y + r, o->gs.y - 2.0),
                       (int)dmin(center.z + r, o->gs.z - 2.0)};
        for (int x = mn.x; x <= mx.x; x++)
            for (int y = mn.y; y <= mx.y; y++)
                for (int z = mn.z; z <= mx.z; z++) {
                    d3 c2o = sub(mk(x + 0.5, y + 0.5, z + 0.5), center);
                    if (dot3(c2o, c2o) < r2) solid_cell(o, x, y, z, speed);
                }
    }
    if (ob->kind == FSIM_OBSTACLE_BOX) {
        const d3 d = scl(mul(mk(ob->size[0], ob->size[1], ob->size[2]), o->cellDInv), 0.5);
        const i3 mn = {(int)dmax(round(center.x - d.x), 1.0), (int)dmax(round(center.y - d.y), 1.0), (int)dmax(round(center.z - d.z), 1.0)};
        const i3 mx = {(int)dmin(round(center.x + d.x), o->gs.x - 1.0), (int)dmin(round(center.y + d.y), o->gs.y - 1.0),
                       (int)dmin(round(center.z + d.z), o->gs.z - 1.0)};
        for (int x = mn.x; x < mx.x; x++)
            for (int y = mn.y; y < mx.y; y++)
                for (int z = mn.z; z < mx.z; z++) solid_cell(o, x, y, z, speed);
    }
}

/* MacGrid::postP2GUpdate macGrid.cpp:204-219 */
void oracle_post_p2g_update(Oracle* o, double gravityIncrement) {
    for (int x = 0; x < o->gs.x; x++)
        for (int y = 0; y < o->gs.y; y++)
            for (int z = 0; z < o->gs.z; z++) {
                const int64_t c = cidx(o, x, y, z);
                for (int a = 0; a < 3; a++) o->v2[3 * c + a] = o->v[3 * c + a];
                /* cell<1,1>(pos) is rawCells[index + gridSize.z]: for y == gy-1 that is the y=0 cell of the next
                 * x column (a SOLID border cell); short-circuit keeps x == gx-1 from reading past the array. */
                if (o->type[c] != FSIM_CELL_SOLID && o->type[c + o->gs.z] != FSIM_CELL_SOLID)
                    o->v2[3 * c + 1] += gravityIncrement;
            }
    o->nfluid = 0;
    for (int x = 1; x < o->gs.x - 1; x++)
        for (int y = 1; y < o->gs.y - 1; y++)
            for (int z = 1; z < o->gs.z - 1; z++) {
                const int64_t c = cidx(o, x, y, z);
                if (o->type[c] == FSIM_CELL_WATER) {
                    o->fluid[o->nfluid].x = x; o->fluid[o->nfluid].y = y; o->fluid[o->nfluid].z = z;
                    o->id[c] = o->nfluid++;
                }
            }
}

/* simulate() lines 82-85: mark, obstacles (in list order), borders, gravity + numbering */
void oracle_stage_classify(Oracle* o, double dt) {
    mark_fluid(o);
    for (int k = 0; k < o->nobs; k++) add_obstacle(o, &o->obs[k]);
    restore_borders(o);
    oracle_post_p2g_update(o, o->par.gravity_enabled ? (double)o->par.gravity * dt : 0.0);
}

typedef struct { double diag, xw, yw, zw; } ARow; /* AMatrixRow bridsonSolverGrid.h:19-24 */

static const int NB[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
static uint8_t type_at(const Oracle* o, i3 p, int a, int s) {
    return o->type[cidx(o, p.x + s * NB[a][0], p.y + s * NB[a][1], p.z + s * NB[a][2])];
}
static int id_at(const Oracle* o, i3 p, int a, int s) {
    return o->id[cidx(o, p.x + s * NB[a][0], p.y + s * NB[a][1], p.z + s * NB[a][2])];
}

/* BridsonSolverGrid::calculateRHS bridsonSolverGrid.cpp:79-89 */
static void calc_rhs(const Oracle* o, double* rhs) {
    const double scale = 1.0 / o->cellD.x;
    for (int i = 0; i < o->nfluid; i++) {
        const i3 p = o->fluid[i];
        const int64_t c = cidx(o, p.x, p.y, p.z);
        rhs[i] = -scale * (o->v2[3 * c + 0] + o->v2[3 * c + 1] + o->v2[3 * c + 2] - o->v2[3 * cidx(o, p.x - 1, p.y, p.z) + 0] -
                           o->v2[3 * cidx(o, p.x, p.y - 1, p.z) + 1] - o->v2[3 * cidx(o, p.x, p.y, p.z - 1) + 2]) +
                 (o->par.pressure_enabled ? (o->avgp[c] - o->par.average_pressure) * o->par.pressure_k : 0.0);
    }
}

/* BridsonSolverGrid::calculateAMatrix bridsonSolverGrid.cpp:40-77 */
static void calc_a(const Oracle* o, ARow* A, double dt) {
    const double scale = dt / (o->par.fluid_density * o->cellD.x * o->cellD.x);
    for (int i = 0; i < o->nfluid; i++) {
        ARow* row = &A[i];
        const i3 p = o->fluid[i];
        row->diag = row->xw = row->yw = row->zw = 0;
        double* w[3] = {&row->xw, &row->yw, &row->zw};
        for (int a = 0; a < 3; a++) {
            uint8_t tp = type_at(o, p, a, 1);
            if (tp == FSIM_CELL_WATER) { row->diag += scale; *w[a] = -scale; }
            else if (tp == FSIM_CELL_AIR) row->diag += scale;
            if (type_at(o, p, a, -1) != FSIM_CELL_SOLID) row->diag += scale;
        }
    }
}

/* BridsonSolverGrid::calculatePreconditioner bridsonSolverGrid.cpp:91-124 (MIC(0), tau 0.97, sigma 0.25) */
static void calc_precon(const Oracle* o, const ARow* A, double* pre) {
    const double tau = 0.97, sigma = 0.25;
    for (int i = 0; i < o->nfluid; i++) {
        const i3 p = o->fluid[i];
        double eNeg = 0, eNegTau = 0;
        if (type_at(o, p, 0, -1) == FSIM_CELL_WATER) {
            const int j = id_at(o, p, 0, -1);
            const double t = A[j].xw * pre[j];
            eNeg += t * t;
            eNegTau += t * (A[j].yw + A[j].zw) * pre[j];
        }
        if (type_at(o, p, 1, -1) == FSIM_CELL_WATER) {
            const int j = id_at(o, p, 1, -1);
            const double t = A[j].yw * pre[j];
            eNeg += t * t;
            eNegTau += t * (A[j].xw + A[j].zw) * pre[j];
        }
        if (type_at(o, p, 2, -1) == FSIM_CELL_WATER) {
            const int j = id_at(o, p, 2, -1);
            const double t = A[j].zw * pre[j];
            eNeg += t * t;
            eNegTau += t * (A[j].xw + A[j].yw) * pre[j];
        }
        double e = A[i].diag - eNeg - eNegTau * tau;
        if (e < sigma * A[i].diag) e = (A[i].diag < 1e-6 ? 1.0 : A[i].diag);
        pre[i] = 1.0 / sqrt(e);
    }
}

/* BridsonSolverGrid::applyPreconditioner bridsonSolverGrid.cpp:126-163 */
static void apply_precon(const Oracle* o, const ARow* A, const double* pre, const double* r, double* q, double* res) {
    for (int i = 0; i < o->nfluid; i++) {
        const i3 p = o->fluid[i];
        double qneg = 0;
        if (type_at(o, p, 0, -1) == FSIM_CELL_WATER) { const int j = id_at(o, p, 0, -1); qneg += A[j].xw * q[j] * pre[j]; }
        if (type_at(o, p, 1, -1) == FSIM_CELL_WATER) { const int j = id_at(o, p, 1, -1); qneg += A[j].yw * q[j] * pre[j]; }
        if (type_at(o, p, 2, -1) == FSIM_CELL_WATER) { const int j = id_at(o, p, 2, -1); qneg += A[j].zw * q[j] * pre[j]; }
        q[i] = (r[i] - qneg) * pre[i];
    }
    for (int i = o->nfluid - 1; i >= 0; i--) {
        const i3 p = o->fluid[i];
        double tneg = 0;
        if (type_at(o, p, 0, 1) == FSIM_CELL_WATER) tneg += A[i].xw * res[id_at(o, p, 0, 1)];
        if (type_at(o, p, 1, 1) == FSIM_CELL_WATER) tneg += A[i].yw * res[id_at(o, p, 1, 1)];
        if (type_at(o, p, 2, 1) == FSIM_CELL_WATER) tneg += A[i].zw * res[id_at(o, p, 2, 1)];
        res[i] = (q[i] - tneg * pre[i]) * pre[i];
    }
}

/* BridsonSolverGrid::applyAMatrix bridsonSolverGrid.cpp:165-198 */
static void apply_a(const Oracle* o, const ARow* A, const double* vec, double* res) {
    for (int i = 0; i < o->nfluid; i++) {
        const i3 p = o->fluid[i];
        double val = A[i].diag * vec[i];
        if (type_at(o, p, 0, 1) == FSIM_CELL_WATER) val += A[i].xw * vec[id_at(o, p, 0, 1)];
        if (type_at(o, p, 1, 1) == FSIM_CELL_WATER) val += A[i].yw * vec[id_at(o, p, 1, 1)];
        if (type_at(o, p, 2, 1) == FSIM_CELL_WATER) val += A[i].zw * vec[id_at(o, p, 2, 1)];
        if (type_at(o, p, 0, -1) == FSIM_CELL_WATER) { const int j = id_at(o, p, 0, -1); val += A[j].xw * vec[j]; }
        if (type_at(o, p, 1, -1) == FSIM_CELL_WATER) { const int j = id_at(o, p, 1, -1); val += A[j].yw * vec[j]; }
        if (type_at(o, p, 2, -1) == FSIM_CELL_WATER) { const int j = id_at(o, p, 2, -1); val += A[j].zw * vec[j]; }
        res[i] = val;
    }
}

static double dotv(const double* a, const double* b, int n) { /* bridsonSolverGrid.cpp:200-214 */
    double r = 0.0;
    for (int i = 0; i < n; i++) r += a[i] * b[i];
    return r;
}

/* BridsonSolverGrid::applyPressureToVelocities bridsonSolverGrid.cpp:295-327 */
static void apply_pressure(Oracle* o, double dt, const double* pr) {
    const double scale = dt / (o->par.fluid_density * o->cellD.x);
    for (int i = 0; i < o->nfluid; i++) {
        const i3 p = o->fluid[i];
        const int64_t c = cidx(o, p.x, p.y, p.z);
        for (int a = 0; a < 3; a++) {
            uint8_t tp = type_at(o, p, a, 1);
            if (tp != FSIM_CELL_SOLID) {
                if (tp == FSIM_CELL_AIR) o->v2[3 * c + a] += scale * pr[i];
                else o->v2[3 * c + a] += scale * (pr[i] - pr[id_at(o, p, a, 1)]);
            }
        }
        for (int a = 0; a < 3; a++)
            if (type_at(o, p, a, -1) == FSIM_CELL_AIR)
                o->v2[3 * cidx(o, p.x - NB[a][0], p.y - NB[a][1], p.z - NB[a][2]) + a] -= scale * pr[i];
    }
}

static int project_basic(Oracle* o);

/* BridsonSolverGrid::solveIncompressibility bridsonSolverGrid.cpp:244-293 */
int oracle_stage_project(Oracle* o, double dt) {
    if (o->par.solver_type == FSIM_SOLVER_BASIC) return project_basic(o);
    const int n = o->nfluid;
    double* pressure = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    double* z = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    double* q = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    double* r = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    double* s = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    double* pre = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    ARow* A = (ARow*)calloc(n > 0 ? n : 1, sizeof(ARow));
    memset(&o->solve, 0, sizeof(o->solve));
    o->solve.fluid_cells = n;
    o->pressure_valid = 0;
    calc_rhs(o, r);
    double total = 0;
    for (int i = 0; i < n; i++) total += r[i] * r[i];
    o->solve.rhs_sumsq = total;
    int it = 0;
    if (total < 1e-7) {
        o->solve.early_out = 1;
        goto done;
    }
    calc_a(o, A, dt);
    calc_precon(o, A, pre);
    apply_precon(o, A, pre, r, q, z);
    memcpy(s, z, sizeof(double) * n);
    double sigma = dotv(z, r, n);
    for (; it < o->par.max_iterations; it++) {
        apply_a(o, A, s, z);
        double alpha = sigma / dotv(s, z, n);
        if (alpha != alpha) break;
        for (int i = 0; i < n; i++) pressure[i] += s[i] * alpha;
        for (int i = 0; i < n; i++) r[i] += z[i] * -alpha;
        double mx = 0.0;
        for (int i = 0; i < n; i++) if (fabs(r[i]) > mx) mx = fabs(r[i]);
        o->solve.residual_max = mx;
        if (mx < o->par.residual_tolerance) break;
        apply_precon(o, A, pre, r, q, z);
        double sigmaNew = dotv(z, r, n);
        if (sigma != sigma) break;
        double beta = sigmaNew / sigma;
        for (int i = 0; i < n; i++) s[i] = s[i] * beta + z[i];
        sigma = sigmaNew;
    }
    apply_pressure(o, dt, pressure);
    memcpy(o->pressure, pressure, sizeof(double) * n);
    o->pressure_valid = 1;
done:
    o->solve.iterations = it;
    free(pressure); free(z); free(q); free(r); free(s); free(pre); free(A);
    return it;
}

/* BasicMacGrid::solveIncompressibility basicMacGrid.cpp:5-102 (the parallel branch's red-black order: it is the one a
 * release build runs, and it is deterministic because cells of one colour share no face) */
static int project_basic(Oracle* o) {
    const double overRelaxation = 1.98;
    const i3 g = o->gs;
    for (int n = 0; n < o->par.max_iterations; n++)
        for (int pass = 0; pass < 2; pass++)
            for (int x = 1; x < g.x - 1; x++)
                for (int y = 1; y < g.y - 1; y++)
                    for (int z = (pass == 0 ? 1 + ((x + y) % 2) : 2 - ((x + y) % 2)); z < g.z - 1; z += 2) {
                        const int64_t c = cidx(o, x, y, z);
                        if (o->type[c] != FSIM_CELL_WATER) continue;
                        const int64_t xm = cidx(o, x - 1, y, z), ym = cidx(o, x, y - 1, z), zm = cidx(o, x, y, z - 1);
                        const int s1 = o->type[cidx(o, x, y, z + 1)] != FSIM_CELL_SOLID, s2 = o->type[zm] != FSIM_CELL_SOLID;
                        const int s3 = o->type[cidx(o, x, y + 1, z)] != FSIM_CELL_SOLID, s4 = o->type[ym] != FSIM_CELL_SOLID;
                        const int s5 = o->type[cidx(o, x + 1, y, z)] != FSIM_CELL_SOLID, s6 = o->type[xm] != FSIM_CELL_SOLID;
                        const int s = s1 + s2 + s3 + s4 + s5 + s6;
                        if (s == 0) continue;
                        double d = -o->v2[3 * c + 0] - o->v2[3 * c + 1] - o->v2[3 * c + 2] + o->v2[3 * xm + 0] + o->v2[3 * ym + 1] + o->v2[3 * zm + 2] +
                                   (o->par.pressure_enabled ? (o->avgp[c] - o->par.average_pressure) * o->par.pressure_k : 0.0);
                        d = d * overRelaxation / s;
                        if (s1) o->v2[3 * c + 2] += d;
                        if (s2) o->v2[3 * zm + 2] -= d;
                        if (s3) o->v2[3 * c + 1] += d;
                        if (s4) o->v2[3 * ym + 1] -= d;
                        if (s5) o->v2[3 * c + 0] += d;
                        if (s6) o->v2[3 * xm + 0] -= d;
                    }
    o->pressure_valid = 0;
    memset(&o->solve, 0, sizeof(o->solve));
    o->solve.iterations = o->par.max_iterations;
    return o->par.max_iterations;
}

/* MacGrid::extrapolateVelocities macGrid.cpp:294-336 */
void oracle_stage_extrapolate(Oracle* o) {
    uint8_t* valid = (uint8_t*)malloc(o->nc);
    memset(valid, 100, o->nc);
    for (int i = 0; i < o->nfluid; i++) valid[cidx(o, o->fluid[i].x, o->fluid[i].y, o->fluid[i].z)] = 0;
    const i3 g = o->gs;
    for (int it = 0; it < 2; it++) {
        for (int x = 0; x < g.x; x++)
            for (int y = 0; y < g.y; y++)
                for (int z = 0; z < g.z; z++) {
                    const int64_t c = cidx(o, x, y, z);
                    if (valid[c] <= it) continue;
                    int cnt = 0;
                    double vs[3] = {0, 0, 0};
                    for (int a = 0; a < 3; a++)
                        for (int off = 0; off < 2; off++) {
                            int nx = x + NB[a][0] * (off * 2 - 1), ny = y + NB[a][1] * (off * 2 - 1), nz = z + NB[a][2] * (off * 2 - 1);
                            if (nx >= 0 && nx < g.x && ny >= 0 && ny < g.y && nz >= 0 && nz < g.z && valid[cidx(o, nx, ny, nz)] <= it) {
                                const int64_t nn = cidx(o, nx, ny, nz);
                                vs[0] += o->v2[3 * nn + 0]; vs[1] += o->v2[3 * nn + 1]; vs[2] += o->v2[3 * nn + 2];
                                cnt++;
                            }
                        }
                    if (cnt > 0) {
                        if (x + 1 < g.x && o->type[cidx(o, x + 1, y, z)] != FSIM_CELL_WATER) o->v2[3 * c + 0] = vs[0] / cnt;
                        if (y + 1 < g.y && o->type[cidx(o, x, y + 1, z)] != FSIM_CELL_WATER) o->v2[3 * c + 1] = vs[1] / cnt;
                        if (z + 1 < g.z && o->type[cidx(o, x, y, z + 1)] != FSIM_CELL_WATER) o->v2[3 * c + 2] = vs[2] / cnt;
                        valid[c] = (uint8_t)(it + 1);
                    }
                }
    }
    free(valid);
}

/* Simulator::g2pTransfer simulator.cpp:375-418 */
void oracle_stage_g2p(Oracle* o) {
    const int tt = o->par.transfer_type;
    for (int64_t i = 0; i < o->np; i++) {
        Particle* P = &o->p[i];
        for (int axis = 0; axis < 3; axis++) {
            if (o->twoD && axis == 2) { P->v.z = 0; break; }
            double pic = 0, flip = 0;
            d3 cvec = mk(0, 0, 0);
            i3 b = face_base(o, P->pos, axis);
            for (int k = 0; k < 8; k++) {
                int x = b.x + CORNER[k][0], y = b.y + CORNER[k][1], z = b.z + CORNER[k][2];
                const int64_t ci = cidx(o, x, y, z);
                d3 fp = face_pos(o, x, y, z, axis);
                double w = trilinear(fp, P->pos, o->cellDInv);
                pic += o->v2[3 * ci + axis] * w;
                if (tt == FSIM_TRANSFER_FLIP) flip += (o->v2[3 * ci + axis] - o->v[3 * ci + axis]) * w;
                if (tt == FSIM_TRANSFER_APIC) cvec = add(cvec, scl(trilinear_grad(fp, P->pos, o->cellDInv), o->v2[3 * ci + axis]));
            }
            switch (tt) {
                case FSIM_TRANSFER_PIC: setc(&P->v, axis, pic); break;
                case FSIM_TRANSFER_FLIP: /* flipRatio is a float: (1 - flipRatio) is evaluated in float (simulator.cpp:408-409) */
                    setc(&P->v, axis, pic * (double)(1 - o->par.flip_ratio) + (flip + getc(P->v, axis)) * (double)o->par.flip_ratio);
                    break;
                default: setc(&P->v, axis, pic); P->c[axis] = cvec; break;
            }
        }
    }
}

/* Simulator::simulate simulator.cpp:51-100 (push-apart, lines 61-64, is SURVEY §8(f) next #1 and not restated) */
int oracle_step(Oracle* o, double dt) {
    if (o->par.spawning_enabled) oracle_stage_spawn(o, dt);
    oracle_stage_advect(o, dt);
    oracle_stage_push_out(o);
    if (o->par.stop_particles) oracle_stage_stop(o);
    oracle_stage_p2g(o);
    oracle_stage_classify(o, dt);
    int it = oracle_stage_project(o, dt);
    oracle_stage_extrapolate(o);
    oracle_stage_g2p(o);
    return it;
}

void oracle_get_solve_info(const Oracle* o, FsimSolveInfo* s) { *s = o->solve; }

int oracle_get_grid(const Oracle* o, int field, void* out) {
    switch (field) {
        case FSIM_FIELD_TYPE: memcpy(out, o->type, o->nc); return 0;
        case FSIM_FIELD_V: memcpy(out, o->v, sizeof(double) * 3 * o->nc); return 0;
        case FSIM_FIELD_V2: memcpy(out, o->v2, sizeof(double) * 3 * o->nc); return 0;
        case FSIM_FIELD_WSUM: memcpy(out, o->wsum, sizeof(double) * 3 * o->nc); return 0;
        case FSIM_FIELD_AVGPNUM: memcpy(out, o->avgp, sizeof(double) * o->nc); return 0;
        case FSIM_FIELD_PRESSURE: {
            double* d = (double*)out;
            memset(d, 0, sizeof(double) * o->nc);
            if (o->pressure_valid)
                for (int i = 0; i < o->nfluid; i++) d[cidx(o, o->fluid[i].x, o->fluid[i].y, o->fluid[i].z)] = o->pressure[i];
            return 0;
        }
        case FSIM_FIELD_RHS: {
            double* d = (double*)out;
            memset(d, 0, sizeof(double) * o->nc);
            double* rhs = (double*)malloc(sizeof(double) * (o->nfluid > 0 ? o->nfluid : 1));
            calc_rhs(o, rhs);
            for (int i = 0; i < o->nfluid; i++) d[cidx(o, o->fluid[i].x, o->fluid[i].y, o->fluid[i].z)] = rhs[i];
            free(rhs);
            return 0;
        }
        case FSIM_FIELD_PCOUNT: {
            int32_t* d = (int32_t*)out;
            memset(d, 0, sizeof(int32_t) * o->nc);
            for (int64_t i = 0; i < o->np; i++) {
                d3 g = mul(o->p[i].pos, o->cellDInv);
                d[cidx(o, (int)g.x, (int)g.y, (int)g.z)]++;
            }
            return 0;
        }
    }
    return 1;
}

int oracle_set_grid(Oracle* o, int field, const void* in) {
    switch (field) {
        case FSIM_FIELD_TYPE: memcpy(o->type, in, o->nc); return 0;
        case FSIM_FIELD_V: memcpy(o->v, in, sizeof(double) * 3 * o->nc); return 0;
        case FSIM_FIELD_V2: memcpy(o->v2, in, sizeof(double) * 3 * o->nc); return 0;
        case FSIM_FIELD_AVGPNUM: memcpy(o->avgp, in, sizeof(double) * o->nc); return 0;
    }
    return 1;
}

/* SimulationManager gfx export manager/simulationManager.cpp:218-231 (float accumulation of the density) */
void oracle_export_gfx(const Oracle* o, FsimParticleGfx* out, int64_t cap) {
    for (int64_t i = 0; i < o->np && i < cap; i++) {
        const Particle* P = &o->p[i];
        out[i].pos[0] = (float)P->pos.x; out[i].pos[1] = (float)P->pos.y; out[i].pos[2] = (float)P->pos.z;
        out[i].v = (float)len3(P->v);
        float density = 0;
        d3 gp = sub(mul(P->pos, o->cellDInv), mk(0.5, 0.5, 0.5));
        i3 b = {(int)gp.x, (int)gp.y, (int)gp.z};
        for (int k = 0; k < 8; k++) {
            int x = b.x + CORNER[k][0], y = b.y + CORNER[k][1], z = b.z + CORNER[k][2];
            density += trilinear(P->pos, cell_pos(o, x, y, z), o->cellDInv) * o->avgp[cidx(o, x, y, z)];
        }
        out[i].density = density;
    }
}

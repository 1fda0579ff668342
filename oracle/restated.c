/* RESTATED CPU ORACLE - TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch C restatement of the TraceIt physics step, written from the reference's
 * executable semantics (file:line citations are into /root/reference/src/main.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this; the product (libtraceit_b200.so) never links or calls it.
 *
 * PINNING: with pair gravity off and default constants this restatement must equal the
 * verbatim reference (oracle/_ref/libtraceit_ref.so, built in place from the reference
 * TU) BIT FOR BIT on pos/vel/force over K steps - tests/test_oracle.py, and the
 * committed golden vectors under tests/golden/ were produced by the verbatim build.
 * Pair gravity is DEAD CODE in the reference (main.cpp:679-693), so that one term is
 * "parity unpinned by live code": it follows the commented formula with the deviations
 * declared at orc_pair_gravity_f32 below.
 *
 * Build: gcc -O2 -ffp-contract=off (mandatory: FMA contraction changes the bits), no
 * -ffast-math.  x86-64 SSE2 float arithmetic == IEEE binary32, same as the reference
 * build (CMakeLists.txt:1-37 sets no -march, so no FMA there either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Same layout as ref_body in ref_harness.cpp (hot fields of object/collider,
 * main.cpp:79-105). */
typedef struct {
    float pos[3];
    float vel[3];
    float force[3];
    float mass;
    float Cd;
    float area;
    int32_t stationary;
    float center[3];
    float size[3];
    float radius;
    int32_t isSphere;
    int32_t isStatic;
    float rest;
} orc_body;

enum {
    ORC_PAIR_GRAVITY = 1,      /* enable the restated main.cpp:679-693 term           */
    ORC_COLLISIONS = 2,        /* run the main.cpp:673-745 sweep                      */
    ORC_GRAVITY_F64 = 4        /* accumulate the pair-gravity sum in double (tolerance
                                  yardstick only, H7)                                 */
};

typedef struct {
    float g[3];        /* uniform field, reference constant {0,(float)-9.8,0}  main.cpp:749 */
    float G;           /* 6.6743e-11f in the dead block                        main.cpp:685 */
    float dt;          /* 0.01666f                                             main.cpp:750 */
    float atm_density; /* world::atmDensity, UI default 1.2                    main.cpp:1141 */
    int32_t flags;
} orc_params;

void orc_default_params(orc_params* p) {
    p->g[0] = 0.0f;
    p->g[1] = (float)-9.8;
    p->g[2] = 0.0f;
    p->G = 6.6743e-11f;
    p->dt = 0.01666f;
    p->atm_density = 1.2f;
    p->flags = ORC_COLLISIONS;
}

/* dot of main.cpp:195-198: (ax*bx + ay*by) + az*bz, each op rounded. */
static inline float dot3(const float* a, const float* b) {
    float t = a[0] * b[0];
    t = t + a[1] * b[1];
    t = t + a[2] * b[2];
    return t;
}

/* normalize of main.cpp:211-223: sqrt, then three divides; identity on zero length. */
static inline void unit3(const float* v, float* out) {
    float len = sqrtf(dot3(v, v));
    if (len != 0.0f) {
        out[0] = v[0] / len;
        out[1] = v[1] / len;
        out[2] = v[2] / len;
    } else {
        out[0] = v[0];
        out[1] = v[1];
        out[2] = v[2];
    }
}

/* ---------------------------------------------------------------------------------
 * Pair gravity, restated from the commented block main.cpp:679-693:
 *     deltaPos = b.pos - a.pos (a = lower index); d2 = dot(deltaPos,deltaPos);
 *     gravForce = G*(ma*mb)/d2; grav = normalize(deltaPos)*gravForce;
 *     a.force <- +grav ; b.force <- grav*-1.0f
 * Declared deviations: (1) the block ASSIGNS, we accumulate; (2) the
 * force is evaluated at the start of a step from the current positions and consumed by
 * that step's integrator; (3) G is a parameter.  The sequential i<j sweep adds body k's
 * terms in ascending partner order, so the per-target loop below produces the same bits
 * and can run in parallel over k.
 * out: 3 floats per body.
 * --------------------------------------------------------------------------------- */
void orc_pair_gravity_f32(const orc_body* b, int n, float G, float* out) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) {
        float f[3] = {0.0f, 0.0f, 0.0f};
        for (int j = 0; j < n; ++j) {
            if (j == k) continue;
            const orc_body* lo = (j < k) ? &b[j] : &b[k];
            const orc_body* hi = (j < k) ? &b[k] : &b[j];
            float d[3] = {hi->pos[0] - lo->pos[0], hi->pos[1] - lo->pos[1], hi->pos[2] - lo->pos[2]};
            float d2 = dot3(d, d);
            float mag = G * (lo->mass * hi->mass) / d2;
            float dir[3];
            unit3(d, dir);
            float gx = dir[0] * mag, gy = dir[1] * mag, gz = dir[2] * mag;
            if (j < k) { /* k is the higher index: receives grav * -1.0f */
                f[0] = f[0] + gx * -1.0f;
                f[1] = f[1] + gy * -1.0f;
                f[2] = f[2] + gz * -1.0f;
            } else {
                f[0] = f[0] + gx;
                f[1] = f[1] + gy;
                f[2] = f[2] + gz;
            }
        }
        out[3 * k + 0] = f[0];
        out[3 * k + 1] = f[1];
        out[3 * k + 2] = f[2];
    }
}

/* Same physics, every operation in double from the float inputs (tolerance yardstick). */
void orc_pair_gravity_f64(const orc_body* b, int n, double G, double* out) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) {
        double f[3] = {0.0, 0.0, 0.0};
        double xk = b[k].pos[0], yk = b[k].pos[1], zk = b[k].pos[2], mk = b[k].mass;
        for (int j = 0; j < n; ++j) {
            if (j == k) continue;
            double dx = (double)b[j].pos[0] - xk, dy = (double)b[j].pos[1] - yk, dz = (double)b[j].pos[2] - zk;
            double d2 = dx * dx + dy * dy + dz * dz;
            double inv = 1.0 / sqrt(d2);
            double s = G * mk * (double)b[j].mass * inv * inv * inv;
            f[0] += dx * s;
            f[1] += dy * s;
            f[2] += dz * s;
        }
        out[3 * k + 0] = f[0];
        out[3 * k + 1] = f[1];
        out[3 * k + 2] = f[2];
    }
}

/* fp64 force on a SUBSET of targets against all n sources: lets the 1M-body tests check
 * sampled targets without the O(N^2) cost.  out: 3 doubles per listed target. */
void orc_pair_gravity_f64_subset(const orc_body* b, int n, double G, const int32_t* idx, int m, double* out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < m; ++t) {
        int k = idx[t];
        double f[3] = {0.0, 0.0, 0.0};
        double xk = b[k].pos[0], yk = b[k].pos[1], zk = b[k].pos[2], mk = b[k].mass;
        for (int j = 0; j < n; ++j) {
            if (j == k) continue;
            double dx = (double)b[j].pos[0] - xk, dy = (double)b[j].pos[1] - yk, dz = (double)b[j].pos[2] - zk;
            double d2 = dx * dx + dy * dy + dz * dz;
            double inv = 1.0 / sqrt(d2);
            double s = G * mk * (double)b[j].mass * inv * inv * inv;
            f[0] += dx * s;
            f[1] += dy * s;
            f[2] += dz * s;
        }
        out[3 * t + 0] = f[0];
        out[3 * t + 1] = f[1];
        out[3 * t + 2] = f[2];
    }
}

/* Drag + uniform gravity + force sum + semi-implicit Euler for one body
 * (main.cpp:763-791, 811).  fgrav may be NULL (pair gravity off). */
static void integrate_body(orc_body* o, const orc_params* p, const float* fgrav) {
    if (o->stationary) return;                                  /* main.cpp:764 */
    float s2 = dot3(o->vel, o->vel);                            /* main.cpp:767 */
    float dir[3] = {0.0f, 0.0f, 0.0f};
    if (s2 > 0.0001f) {                                         /* main.cpp:771 */
        float inv = 1.0f / sqrtf(s2);
        dir[0] = o->vel[0] * inv;
        dir[1] = o->vel[1] * inv;
        dir[2] = o->vel[2] * inv;
    }
    float k = -0.5f * p->atm_density * o->Cd * o->area * s2;    /* main.cpp:777, left to right */
    float ext[3] = {o->force[0], o->force[1], o->force[2]};
    if (fgrav) {
        ext[0] = ext[0] + fgrav[0];
        ext[1] = ext[1] + fgrav[1];
        ext[2] = ext[2] + fgrav[2];
    }
    for (int c = 0; c < 3; ++c) {
        float Fd = dir[c] * k;
        float tot = Fd + p->g[c] * o->mass;                     /* main.cpp:780 */
        tot = tot + ext[c];                                     /* main.cpp:782 */
        float acc = tot / o->mass;                              /* main.cpp:786 */
        o->vel[c] = o->vel[c] + acc * p->dt;                    /* main.cpp:787 */
        o->pos[c] = o->pos[c] + o->vel[c] * p->dt;              /* main.cpp:788 */
        o->center[c] = o->pos[c];                               /* main.cpp:791 -> 664 */
        o->force[c] = 0.0f;                                     /* main.cpp:811 */
    }
}

/* Predicates (main.cpp:648-660). */
static inline int sphere_hit(const orc_body* a, const orc_body* b) {
    float d[3] = {a->center[0] - b->center[0], a->center[1] - b->center[1], a->center[2] - b->center[2]};
    float d2 = dot3(d, d);
    float rs = a->radius + b->radius;
    return d2 <= rs * rs;
}

static inline int aabb_hit(const orc_body* a, const orc_body* b) {
    return fabsf(a->center[0] - b->center[0]) <= (a->size[0] + b->size[0]) &&
           fabsf(a->center[1] - b->center[1]) <= (a->size[1] + b->size[1]) &&
           fabsf(a->center[2] - b->center[2]) <= (a->size[2] + b->size[2]);
}

/* calculateAABBNormal, main.cpp:622-638. */
static inline void aabb_normal(const orc_body* a, const orc_body* b, float* n) {
    float d[3] = {b->center[0] - a->center[0], b->center[1] - a->center[1], b->center[2] - a->center[2]};
    float ox = (a->size[0] + b->size[0]) - fabsf(d[0]);
    float oy = (a->size[1] + b->size[1]) - fabsf(d[1]);
    float oz = (a->size[2] + b->size[2]) - fabsf(d[2]);
    n[0] = n[1] = n[2] = 0.0f;
    if (ox < oy && ox < oz) n[0] = d[0] > 0 ? 1.0f : -1.0f;
    else if (oy < oz)       n[1] = d[1] > 0 ? 1.0f : -1.0f;
    else                    n[2] = d[2] > 0 ? 1.0f : -1.0f;
}

/* Guards + predicate of one (i<j) visit, main.cpp:695-707.  Returns 1 on contact and
 * fills the normal. */
static inline int pair_contact(const orc_body* a, const orc_body* b, float* normal) {
    if (a->isStatic && b->isStatic) return 0;
    if (a->mass <= 0 || b->mass <= 0) return 0;
    if (a->isSphere && b->isSphere) {
        if (!sphere_hit(a, b)) return 0;
        if (normal) {
            float d[3] = {b->center[0] - a->center[0], b->center[1] - a->center[1], b->center[2] - a->center[2]};
            unit3(d, normal);
        }
        return 1;
    }
    if (!aabb_hit(a, b)) return 0;
    if (normal) aabb_normal(a, b, normal);
    return 1;
}

/* Impulse + positional nudge, main.cpp:709-742, including the approach test exactly as
 * written (it skips when relVel.n > 0 with relVel = a.vel - b.vel). */
static inline void respond(orc_body* a, orc_body* b, const float* n) {
    float rv[3] = {a->vel[0] - b->vel[0], a->vel[1] - b->vel[1], a->vel[2] - b->vel[2]};
    float vn = dot3(rv, n);
    if (vn > 0) return;
    float e = (b->rest < a->rest) ? b->rest : a->rest;          /* std::min(a,b) */
    float j = -(1 + e) * vn;
    float ima = 1 / a->mass, imb = 1 / b->mass;
    j = j / (ima + imb);
    float imp[3] = {n[0] * j, n[1] * j, n[2] * j};
    if (!a->isStatic) for (int c = 0; c < 3; ++c) a->vel[c] = a->vel[c] - imp[c] * ima;
    if (!b->isStatic) for (int c = 0; c < 3; ++c) b->vel[c] = b->vel[c] + imp[c] * imb;
    const float slop = 0.01f, percent = 0.2f;
    float k = percent * slop;
    float cor[3] = {n[0] * k, n[1] * k, n[2] * k};
    if (!a->isStatic) for (int c = 0; c < 3; ++c) a->pos[c] = a->pos[c] - cor[c] * ima;
    if (!b->isStatic) for (int c = 0; c < 3; ++c) b->pos[c] = b->pos[c] + cor[c] * imb;
}

/* The sequential Gauss-Seidel sweep, main.cpp:673-745 (brute force, lexicographic). */
void orc_resolve(orc_body* b, int n) {
    for (int i = 0; i < n; ++i) {
        for (int j = i + 1; j < n; ++j) {
            float nrm[3];
            if (pair_contact(&b[i], &b[j], nrm)) respond(&b[i], &b[j], nrm);
        }
    }
}

/* Response replay over a PRECOMPUTED, lexicographically sorted contact list.  Legal
 * because the predicate reads only collider centres/radii/sizes, which the sweep never
 * writes (main.cpp:655-659 vs 725-741) - so it is bit-identical to
 * orc_resolve while skipping the O(N^2) non-contacts.  Used for large-N drift checks. */
void orc_resolve_list(orc_body* b, const int32_t* pairs, long long ncontacts) {
    for (long long c = 0; c < ncontacts; ++c) {
        int i = pairs[2 * c], j = pairs[2 * c + 1];
        float nrm[3];
        if (pair_contact(&b[i], &b[j], nrm)) respond(&b[i], &b[j], nrm);
    }
}

/* Brute-force contact set of the current collider state, lexicographic order.
 * Parallel over i (detection is order independent).  Returns the count; writes at most
 * cap pairs.  Two passes: count per row, then fill. */
long long orc_contacts(const orc_body* b, int n, int32_t* pairs, long long cap) {
    long long* row = (long long*)calloc((size_t)n + 1, sizeof(long long));
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        long long c = 0;
        for (int j = i + 1; j < n; ++j) c += pair_contact(&b[i], &b[j], NULL);
        row[i + 1] = c;
    }
    for (int i = 0; i < n; ++i) row[i + 1] += row[i];
    long long total = row[n];
    if (pairs) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) {
            long long w = row[i];
            if (row[i + 1] == w) continue;
            for (int j = i + 1; j < n; ++j) {
                if (pair_contact(&b[i], &b[j], NULL)) {
                    if (w < cap) { pairs[2 * w] = i; pairs[2 * w + 1] = j; }
                    ++w;
                }
            }
        }
    }
    free(row);
    return total;
}

/* One full restated step x nsteps: [pair gravity] -> integrate -> [collision sweep]. */
void orc_step(orc_body* b, int n, const orc_params* p, int nsteps) {
    float* fg = NULL;
    double* fg64 = NULL;
    if (p->flags & ORC_PAIR_GRAVITY) {
        fg = (float*)malloc(sizeof(float) * 3 * (size_t)n);
        if (p->flags & ORC_GRAVITY_F64) fg64 = (double*)malloc(sizeof(double) * 3 * (size_t)n);
    }
    for (int s = 0; s < nsteps; ++s) {
        if (fg) {
            if (fg64) {
                orc_pair_gravity_f64(b, n, (double)p->G, fg64);
                for (size_t t = 0; t < 3 * (size_t)n; ++t) fg[t] = (float)fg64[t];
            } else {
                orc_pair_gravity_f32(b, n, p->G, fg);
            }
        }
        for (int i = 0; i < n; ++i) integrate_body(&b[i], p, fg ? fg + 3 * (size_t)i : NULL);
        if (p->flags & ORC_COLLISIONS) orc_resolve(b, n);
    }
    free(fg);
    free(fg64);
}

/* Only the O(N) part of a step with an externally supplied pair-gravity force
 * (3 floats per body, may be NULL). */
void orc_integrate(orc_body* b, int n, const orc_params* p, const float* fgrav) {
    for (int i = 0; i < n; ++i) integrate_body(&b[i], p, fgrav ? fgrav + 3 * (size_t)i : NULL);
}

/* ---------------------------------------------------------------------------------
 * Uniform-grid broad phase restatement.  NO reference counterpart (the reference is
 * brute force, main.cpp:674-675): this DEFINES the cell key, the sort order and the
 * candidate-pair set that the CUDA path must reproduce bit for bit; its final contact
 * set is cross-checked against orc_contacts / the verbatim reference.
 *
 *   u_a   = (center_a - origin_a) * inv_cell          fp32 sub, fp32 mul, no FMA
 *   c_a   = (int32) floorf(u_a)   saturating, NaN -> 0
 *   key   = (cx & mx) | (cy & my) << bx | (cz & mz) << (bx+by),  m_a = 2^b_a - 1, 2 <= b_a <= 10
 *   order = stable sort of body ids by key  ==  ascending (key, id)
 *   candidates = { (i,j) : i<j, both grid bodies, cell(j) in the 27-stencil of cell(i)
 *                  taken modulo 2^b_a per axis }
 * The `bits` argument of the functions below is a code: b (< 256) = the same b bits on every
 * axis, otherwise bx | by << 8 | bz << 16 (what tr_debug_grid reports for the grid it used).
 * A body is a "grid body" when in_grid[id] != 0 (spheres with radius <= the grid's
 * rmax and |u| < 32768 on every axis; the host decides - see traceit_b200 DESIGN.md).
 * --------------------------------------------------------------------------------- */
static inline int32_t cell_coord(float center, float origin, float inv_cell) {
    float u = (center - origin) * inv_cell;
    float f = floorf(u);
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

void orc_grid_cells(const orc_body* b, int n, const float* origin, float inv_cell, int32_t* cells /*3n*/) {
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) cells[3 * i + c] = cell_coord(b[i].center[c], origin[c], inv_cell);
}

static inline void axis_bits(int code, int* bx, int* by, int* bz) {
    if (code < 256) {
        *bx = *by = *bz = code;
    } else {
        *bx = code & 0xFF;
        *by = (code >> 8) & 0xFF;
        *bz = (code >> 16) & 0xFF;
    }
}

static inline size_t table_cells(int code) {
    int bx, by, bz;
    axis_bits(code, &bx, &by, &bz);
    return (size_t)1 << (bx + by + bz);
}

static inline uint32_t pack_key(int32_t cx, int32_t cy, int32_t cz, int bits) {
    int bx, by, bz;
    axis_bits(bits, &bx, &by, &bz);
    uint32_t mx = (1u << bx) - 1u, my = (1u << by) - 1u, mz = (1u << bz) - 1u;
    return ((uint32_t)cx & mx) | (((uint32_t)cy & my) << bx) | (((uint32_t)cz & mz) << (bx + by));
}

void orc_grid_keys(const orc_body* b, int n, const float* origin, float inv_cell, int bits, uint32_t* keys) {
    for (int i = 0; i < n; ++i) {
        int32_t cx = cell_coord(b[i].center[0], origin[0], inv_cell);
        int32_t cy = cell_coord(b[i].center[1], origin[1], inv_cell);
        int32_t cz = cell_coord(b[i].center[2], origin[2], inv_cell);
        keys[i] = pack_key(cx, cy, cz, bits);
    }
}

/* Domain flag per body: 1 when |u| < 32768 on all axes (and not NaN). */
void orc_grid_in_domain(const orc_body* b, int n, const float* origin, float inv_cell, uint8_t* ok) {
    for (int i = 0; i < n; ++i) {
        int good = 1;
        for (int c = 0; c < 3; ++c) {
            float u = (b[i].center[c] - origin[c]) * inv_cell;
            if (!(fabsf(u) < 32768.0f)) good = 0;
        }
        ok[i] = (uint8_t)good;
    }
}

/* Stable counting sort of the ids listed in `ids` (m of them, ascending) by key.
 * order_out receives the m ids in (key, id) order; sorted_keys_out their keys. */
void orc_grid_sort(const uint32_t* keys, const int32_t* ids, int m, int bits, int32_t* order_out,
                   uint32_t* sorted_keys_out) {
    size_t ncell = table_cells(bits);
    uint32_t* cnt = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
    for (int t = 0; t < m; ++t) cnt[keys[ids[t]] + 1]++;
    for (size_t c = 0; c < ncell; ++c) cnt[c + 1] += cnt[c];
    for (int t = 0; t < m; ++t) {
        uint32_t k = keys[ids[t]];
        uint32_t dst = cnt[k]++;
        order_out[dst] = ids[t];
        if (sorted_keys_out) sorted_keys_out[dst] = k;
    }
    free(cnt);
}

static int cmp_pair(const void* a, const void* b) {
    const int32_t* x = (const int32_t*)a;
    const int32_t* y = (const int32_t*)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
}

/* Candidate pairs and contacts through the grid.  in_grid: per-body flag.  Outputs are
 * lexicographically sorted (i,j) lists; either may be NULL to only count.
 * Returns 0 on success, -1 on allocation failure. */
int orc_grid_pairs(const orc_body* b, int n, const uint8_t* in_grid, const float* origin, float inv_cell,
                   int bits, int32_t* cand, long long cand_cap, long long* ncand, int32_t* contacts,
                   long long contact_cap, long long* ncontact) {
    int m = 0;
    int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i)
        if (in_grid[i]) ids[m++] = i;
    int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(n > 0 ? n : 1));
    uint32_t* keys = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    size_t ncell = table_cells(bits);
    uint32_t* start = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
    if (!ids || !cells || !keys || !order || !start) return -1;
    orc_grid_cells(b, n, origin, inv_cell, cells);
    for (int i = 0; i < n; ++i) keys[i] = pack_key(cells[3 * i], cells[3 * i + 1], cells[3 * i + 2], bits);
    orc_grid_sort(keys, ids, m, bits, order, NULL);
    for (int t = 0; t < m; ++t) start[keys[ids[t]] + 1]++;
    for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];

    long long nc = 0, nk = 0;
    for (int t = 0; t < m; ++t) {
        int i = ids[t];
        int32_t cx = cells[3 * i], cy = cells[3 * i + 1], cz = cells[3 * i + 2];
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    uint32_t k = pack_key((int32_t)((uint32_t)cx + (uint32_t)dx), (int32_t)((uint32_t)cy + (uint32_t)dy),
                                          (int32_t)((uint32_t)cz + (uint32_t)dz), bits);
                    for (uint32_t s = start[k]; s < start[k + 1]; ++s) {
                        int j = order[s];
                        if (j <= i) continue;
                        if (cand && nc < cand_cap) { cand[2 * nc] = i; cand[2 * nc + 1] = j; }
                        ++nc;
                        if (pair_contact(&b[i], &b[j], NULL)) {
                            if (contacts && nk < contact_cap) { contacts[2 * nk] = i; contacts[2 * nk + 1] = j; }
                            ++nk;
                        }
                    }
                }
    }
    if (cand) qsort(cand, (size_t)(nc < cand_cap ? nc : cand_cap), 2 * sizeof(int32_t), cmp_pair);
    if (contacts) qsort(contacts, (size_t)(nk < contact_cap ? nk : contact_cap), 2 * sizeof(int32_t), cmp_pair);
    if (ncand) *ncand = nc;
    if (ncontact) *ncontact = nk;
    free(ids); free(cells); free(keys); free(order); free(start);
    return 0;
}

/* FNV-1a (64-bit) over the bit patterns of pos and vel, body by body: the state hash
 * stored with the golden vectors. */
uint64_t orc_state_hash(const orc_body* b, int n) {
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < n; ++i) {
        uint32_t w[6];
        memcpy(w, b[i].pos, 12);
        memcpy(w + 3, b[i].vel, 12);
        for (int k = 0; k < 6; ++k)
            for (int s = 0; s < 4; ++s) {
                h ^= (uint64_t)((w[k] >> (8 * s)) & 0xffu);
                h *= 1099511628211ULL;
            }
    }
    return h;
}

int orc_sizeof_body(void) { return (int)sizeof(orc_body); }

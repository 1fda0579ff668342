// VERBATIM ORACLE HARNESS - TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Compiles the UNMODIFIED reference translation unit in place (the path is passed by
// the Makefile as -DTRACEIT_REFERENCE_MAIN="\"/root/reference/src/main.cpp\"") behind
// oracle/stub/ncurses.h, renames its main(), and exports a small C ABI so tests and
// bench.py's cpu_baseline leg can drive the reference's own
//     void updatePhysics(world&, ofstream&)        (reference src/main.cpp:748)
//     void resolveCollisions(world&)               (reference src/main.cpp:673)
//     bool checkSphereCollision / checkAABBCollision (src/main.cpp:655 / 648)
//     vec3d toSpherical(elev, azim, magn)          (src/main.cpp:1080)
// No reference source is copied into this repository; the .so is written to
// oracle/_ref/ (git-ignored) and only exists where /root/reference was present at
// build time (it still travels to the GPU box as a built file).
#include <cstdint>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>

#define main traceit_reference_main
#include TRACEIT_REFERENCE_MAIN
#undef main

extern "C" {

// Plain mirror of the hot fields of the reference's `object` + `collider`
// (src/main.cpp:79-105).  Layout is shared with oracle/pyoracle.py (ctypes).
struct ref_body {
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
};

struct ref_handle {
    world w;
    std::ofstream coords;
    char scratch[512];
};

static void to_object(const ref_body& b, object& o) {
    o.pos = {b.pos[0], b.pos[1], b.pos[2]};
    o.vel = {b.vel[0], b.vel[1], b.vel[2]};
    o.force = {b.force[0], b.force[1], b.force[2]};
    o.mass = b.mass;
    o.Cd = b.Cd;
    o.area = b.area;
    o.stationary = b.stationary != 0;
    o.trajWritten = false;
    std::memset(&o.transform, 0, sizeof(o.transform));
    o.col.center = {b.center[0], b.center[1], b.center[2]};
    o.col.size = {b.size[0], b.size[1], b.size[2]};
    o.col.radius = b.radius;
    o.col.isSphere = b.isSphere != 0;
    o.col.isStatic = b.isStatic != 0;
    o.col.rest = b.rest;
}

static void from_object(const object& o, ref_body& b) {
    b.pos[0] = o.pos.x; b.pos[1] = o.pos.y; b.pos[2] = o.pos.z;
    b.vel[0] = o.vel.x; b.vel[1] = o.vel.y; b.vel[2] = o.vel.z;
    b.force[0] = o.force.x; b.force[1] = o.force.y; b.force[2] = o.force.z;
    b.mass = o.mass; b.Cd = o.Cd; b.area = o.area;
    b.stationary = o.stationary ? 1 : 0;
    b.center[0] = o.col.center.x; b.center[1] = o.col.center.y; b.center[2] = o.col.center.z;
    b.size[0] = o.col.size.x; b.size[1] = o.col.size.y; b.size[2] = o.col.size.z;
    b.radius = o.col.radius;
    b.isSphere = o.col.isSphere ? 1 : 0;
    b.isStatic = o.col.isStatic ? 1 : 0;
    b.rest = o.col.rest;
}

// scratch_dir: where the reference's per-step "Координаты.txt" append goes
// (src/main.cpp:755).  The harness chdir()s there only for the duration of a step.
void* ref_world_create(float atmDensity, const char* scratch_dir) {
    ref_handle* h = new ref_handle;
    h->w.globalScale = 1.0f;
    h->w.atmDensity = atmDensity;
    h->w.g = 9.8f;
    std::strncpy(h->scratch, scratch_dir ? scratch_dir : "/tmp", sizeof(h->scratch) - 1);
    h->scratch[sizeof(h->scratch) - 1] = 0;
    return h;
}

void ref_world_destroy(void* p) { delete static_cast<ref_handle*>(p); }

int ref_world_size(void* p) { return (int)static_cast<ref_handle*>(p)->w.objects.size(); }

void ref_world_set_atm(void* p, float rho) { static_cast<ref_handle*>(p)->w.atmDensity = rho; }

int ref_world_add(void* p, const ref_body* bodies, int n) {
    ref_handle* h = static_cast<ref_handle*>(p);
    for (int i = 0; i < n; ++i) {
        object o;
        to_object(bodies[i], o);
        h->w.objects.push_back(o);   // same by-value push as src/main.cpp:1242,1491
    }
    return (int)h->w.objects.size();
}

// The reference's own step, nsteps times (the call of src/main.cpp:1589).
void ref_world_step(void* p, int nsteps) {
    ref_handle* h = static_cast<ref_handle*>(p);
    int cwd = open(".", O_RDONLY | O_DIRECTORY);
    if (chdir(h->scratch) != 0) { /* stay where we are */ }
    for (int s = 0; s < nsteps; ++s) updatePhysics(h->w, h->coords);
    if (cwd >= 0) { if (fchdir(cwd) != 0) {} close(cwd); }
}

// Only the collision sweep (src/main.cpp:673), for response-only checks.
void ref_world_resolve(void* p) { resolveCollisions(static_cast<ref_handle*>(p)->w); }

void ref_world_read(void* p, ref_body* out, int first, int count) {
    ref_handle* h = static_cast<ref_handle*>(p);
    for (int i = 0; i < count; ++i) from_object(h->w.objects[first + i], out[i]);
}

void ref_world_set_force(void* p, int idx, float fx, float fy, float fz) {
    static_cast<ref_handle*>(p)->w.objects[idx].force = {fx, fy, fz};
}

// Contact set of the CURRENT collider state: the i<j sweep, guards and predicates of
// src/main.cpp:674-707 evaluated through the reference's own check functions, with no
// response applied.  Returns the number of contacts; writes up to cap (i,j) pairs.
long long ref_world_contacts(void* p, int32_t* pairs, long long cap) {
    ref_handle* h = static_cast<ref_handle*>(p);
    const std::vector<object>& o = h->w.objects;
    long long n = 0;
    for (size_t i = 0; i < o.size(); ++i) {
        for (size_t j = i + 1; j < o.size(); ++j) {
            const object& a = o[i];
            const object& b = o[j];
            if (a.col.isStatic && b.col.isStatic) continue;
            if (a.mass <= 0 || b.mass <= 0) continue;
            bool hit = (a.col.isSphere && b.col.isSphere) ? checkSphereCollision(a.col, b.col)
                                                          : checkAABBCollision(a.col, b.col);
            if (hit) {
                if (n < cap) { pairs[2 * n] = (int32_t)i; pairs[2 * n + 1] = (int32_t)j; }
                ++n;
            }
        }
    }
    return n;
}

// Creation helpers of the reference UI path (src/main.cpp:1080-1089, 1295).
void ref_to_spherical(float elev, float azim, float magn, float out[3]) {
    vec3d v = toSpherical(elev, azim, magn);
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}

float ref_area_from_size(float size) {
    float area = M_PI * size * size;   // same expression and narrowing as src/main.cpp:1295
    return area;
}

int ref_sizeof_object(void) { return (int)sizeof(object); }

}  // extern "C"

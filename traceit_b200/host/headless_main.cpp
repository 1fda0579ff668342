// Headless harness: the reference's frame loop with the ncurses renderer detached.
//
// Seeds the world the way the reference's main() does (earth slab, src/main.cpp:1482-1491),
// adds projectiles with the creation form's parameters (src/main.cpp:1206-1311) and calls
// updatePhysics(w, coords) once per frame (src/main.cpp:1589) - here the B200 shim from
// traceit_host.hpp.  Prints a state hash (FNV-1a over pos/vel bit patterns, the same hash the
// golden vectors store) so tests can compare a whole run with the verbatim reference.
//
//   traceit_headless [--steps K] [--cloud N] [--pack] [--gravity G] [--no-file] [--trace I] [--time]
//     --pack   with --cloud N: the config-5 shape (jittered lattice of radius-0.5 spheres, pitch 0.98) instead of the sparse cloud
//     --time   print "TIME {json}": wall-clock ms per frame through updatePhysics (host pack + upload + step + read-back +
//              host unpack), first two frames excluded (they create the device world and page-lock the staging buffer)
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "traceit_host.hpp"

// Mirror of the reference's data model, same member names (src/main.cpp:51-54,79-105,118-126);
// render-only members (mesh, matrices other than transform) are not needed headless.
struct vec3d { float x, y, z; };
struct matrx4d { float m[4][4]; };
struct collider {
    vec3d center;
    vec3d size;
    float radius;
    bool isSphere;
    bool isStatic;
    float rest;
};
struct object {
    vec3d pos;
    vec3d vel;
    vec3d force;
    float mass;
    matrx4d transform;
    float Cd;
    float area;
    bool stationary;
    std::vector<vec3d> trajectory;
    int maxTrajectoryPoints = 300;
    bool trajWritten = false;
    collider col;
};
struct world {
    std::vector<object> objects;
    float globalScale;
    float atmDensity;
    float g;
};

static uint64_t state_hash(const world& w) {
    uint64_t h = 1469598103934665603ULL;
    for (const object& o : w.objects) {
        const float f[6] = {o.pos.x, o.pos.y, o.pos.z, o.vel.x, o.vel.y, o.vel.z};
        for (int k = 0; k < 6; ++k) {
            uint32_t u;
            std::memcpy(&u, &f[k], 4);
            for (int s = 0; s < 4; ++s) {
                h ^= (uint64_t)((u >> (8 * s)) & 0xffu);
                h *= 1099511628211ULL;
            }
        }
    }
    return h;
}

// One projectile as the creation form builds it (src/main.cpp:1293-1307).
static object make_projectile(float mass, float Cd, float elev, float azim, float magn, float size, vec3d pos) {
    object o{};
    o.mass = mass;
    o.Cd = Cd;
    o.pos = pos;
    o.force = {0, 0, 0};
    o.vel = traceit_b200::toSpherical<vec3d>(elev, azim, magn);
    o.area = tr_area_from_size(size);
    o.col.center = pos;
    o.col.size = {size, size, size};
    o.col.radius = 0.0f;
    o.col.isSphere = false;
    o.col.isStatic = false;
    o.col.rest = 0.5f;
    o.stationary = false;
    return o;
}

int main(int argc, char** argv) {
    int steps = 1000, cloud = 0, trace = 1;
    float G = 0.0f;
    bool file = true, timeit = false, pack = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--steps" && i + 1 < argc) steps = std::atoi(argv[++i]);
        else if (a == "--cloud" && i + 1 < argc) cloud = std::atoi(argv[++i]);
        else if (a == "--gravity" && i + 1 < argc) G = (float)std::atof(argv[++i]);
        else if (a == "--trace" && i + 1 < argc) trace = std::atoi(argv[++i]);
        else if (a == "--no-file") file = false;
        else if (a == "--time") timeit = true;
        else if (a == "--pack") pack = true;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }

    world w;
    w.globalScale = 1.0f;
    w.atmDensity = 1.2f;    // createWorldUI default (src/main.cpp:1141)
    w.g = 9.8f;             // set by the UI, never read by the step (src/main.cpp:1140 vs 749)

    if (cloud <= 0) {
        // the reference's seeded "earth" (src/main.cpp:1482-1491)
        object earth{};
        earth.pos = {0, -20, 0};
        earth.mass = 1000;
        earth.Cd = 1000.0f;
        earth.area = 1000.0f;
        earth.stationary = true;
        earth.col.center = earth.pos;
        earth.col.size = {10000, 5, 10000};
        earth.col.isSphere = false;
        earth.col.isStatic = false;
        earth.col.rest = 1.0f;
        earth.trajWritten = false;
        w.objects.push_back(earth);
        // config 1 projectiles
        w.objects.push_back(make_projectile(10, 0.42f, 45, 0, 30, 1, {0, 0, 0}));
        w.objects.push_back(make_projectile(5, 0.42f, 60, 90, 20, 2, {5, 0, 0}));
    } else if (pack) {
        // config 5 shape: nx x ny x nz lattice (x fastest), pitch 0.98, jitter +-0.01, velocities in [-0.1,0.1]^3, unit mass
        int e = 0;
        while ((1 << (e + 1)) <= cloud) ++e;
        const int ex = (e + 2) / 3, ey = (e - ex + 1) / 2, ez = e - ex - ey;
        const int nx = 1 << ex, ny = 1 << ey, nz = 1 << ez;
        std::mt19937 rng(0xC0FFEE05u);
        std::uniform_real_distribution<float> uj(-0.01f, 0.01f), uv(-0.1f, 0.1f);
        for (int i = 0; i < nx * ny * nz; ++i) {
            const int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
            object o{};
            o.pos = {(ix - (nx - 1) * 0.5f) * 0.98f + uj(rng), (iy - (ny - 1) * 0.5f) * 0.98f + uj(rng), (iz - (nz - 1) * 0.5f) * 0.98f + uj(rng)};
            o.vel = {uv(rng), uv(rng), uv(rng)};
            o.mass = 1.0f;
            o.Cd = 0.42f;
            o.area = tr_area_from_size(0.5f);
            o.col = {o.pos, {0.5f, 0.5f, 0.5f}, 0.5f, true, false, 0.5f};
            o.stationary = false;
            w.objects.push_back(o);
        }
    } else {
        std::mt19937 rng(0xC0FFEE02u);
        float L = 4.0f * std::cbrt((float)cloud);
        std::uniform_real_distribution<float> up(-L, L), uv(-1.0f, 1.0f), um(1.0f, 10.0f);
        for (int i = 0; i < cloud; ++i) {
            object o{};
            o.pos = {up(rng), up(rng), up(rng)};
            o.vel = {uv(rng), uv(rng), uv(rng)};
            o.mass = um(rng);
            o.Cd = 0.42f;
            o.area = tr_area_from_size(0.5f);
            o.col = {o.pos, {0.5f, 0.5f, 0.5f}, 0.5f, true, false, 0.5f};
            o.stationary = false;
            w.objects.push_back(o);
        }
    }

    traceit_b200::options().pair_gravity = G != 0.0f;
    traceit_b200::options().G = G;
    traceit_b200::options().write_trajectory_file = file;

    std::ofstream coords;
    double timed_ms = 0.0, pack_ms = 0.0, step_ms = 0.0, unpack_ms = 0.0;
    int timed_frames = 0;
    try {
        for (int s = 0; s < steps; ++s) {
            const auto t0 = std::chrono::steady_clock::now();
            traceit_b200::updatePhysics(w, coords);   // the call of src/main.cpp:1589
            if (s >= 2) {
                timed_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                double a, b, c;
                traceit_b200::last_frame_breakdown(a, b, c);
                pack_ms += a;
                step_ms += b;
                unpack_ms += c;
                ++timed_frames;
            }
            if ((s + 1 == 100 || s + 1 == steps) && trace >= 0 && trace < (int)w.objects.size()) {
                const object& o = w.objects[trace];
                std::printf("step %d body %d pos %.9g %.9g %.9g vel %.9g %.9g %.9g\n", s + 1, trace, o.pos.x, o.pos.y, o.pos.z,
                            o.vel.x, o.vel.y, o.vel.z);
            }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    if (timeit && timed_frames > 0)
        std::printf("TIME {\"api\": \"traceit_b200::updatePhysics(world&, ofstream&) on 232-byte reference-shaped objects\", "
                    "\"n_bodies\": %zu, \"frames\": %d, \"ms_per_frame\": %.4f, \"host_pack_compare_ms\": %.4f, \"step_and_readback_ms\": %.4f, "
                    "\"host_unpack_history_ms\": %.4f, \"shape\": \"%s\", \"h2d_bytes_per_frame\": %llu, \"d2h_bytes_per_frame\": %llu}\n",
                    w.objects.size(), timed_frames, timed_ms / timed_frames, pack_ms / timed_frames, step_ms / timed_frames, unpack_ms / timed_frames,
                    cloud <= 0 ? "default scene" : (pack ? "cfg5 pack" : "sparse cloud"),
                    (unsigned long long)traceit_b200::last_frame_h2d_bytes(), (unsigned long long)traceit_b200::last_frame_d2h_bytes());
    std::printf("bodies %zu steps %d hash 0x%016llx\n", w.objects.size(), steps, (unsigned long long)state_hash(w));
    traceit_b200::release(w);
    return 0;
}

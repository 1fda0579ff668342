// DROP-IN CHECK - TEST INFRASTRUCTURE ONLY (built into oracle/_ref/, never shipped).
//
// One binary that contains BOTH the unmodified reference translation unit (its own structs
// `world`, `object`, `collider`, `vec3d`, `matrx4d`, its own updatePhysics / toSpherical) and the
// B200 host shim traceit_b200::updatePhysics instantiated ON THOSE SAME TYPES - exactly what a
// maintainer gets after the three-line change of INTEGRATION.md.  It steps two copies of one
// world, one through each, and compares every body (pos, vel, force, collider centre, transform,
// trajectory) bit for bit, plus the two "Координаты.txt" files.
#include <cstdio>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <unistd.h>

#define main traceit_reference_main
#include TRACEIT_REFERENCE_MAIN
#undef main

#include "../traceit_b200/host/traceit_host.hpp"

static bool same_bits(float a, float b) {
    if (a != a && b != b) return true;   // NaN payloads are platform specific
    return std::memcmp(&a, &b, 4) == 0;
}
static bool same_vec(const vec3d& a, const vec3d& b) { return same_bits(a.x, b.x) && same_bits(a.y, b.y) && same_bits(a.z, b.z); }

static std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? std::atoi(argv[1]) : 400;
    const int spheres = argc > 2 ? std::atoi(argv[2]) : 300;
    const std::string base = argc > 3 ? argv[3] : ".";
    const std::string dirA = base + "/ref", dirB = base + "/gpu";
    mkdir(dirA.c_str(), 0755);
    mkdir(dirB.c_str(), 0755);

    // the reference's seeded world (its main(), src/main.cpp:1479-1491) ...
    world a;
    a.globalScale = 1.0f;
    a.atmDensity = 1.2f;
    a.g = 9.8f;
    object earth = {{0, -20, 0}, {0, 0, 0}, {0, 0, 0}, 1000, mesh{}, createTranslationMatrix(0, -20, 0), 1000.0f, 1000.0f, true};
    earth.col.center = earth.pos;
    earth.col.size = {10000, 5, 10000};
    earth.col.radius = 0.0f;
    earth.col.isSphere = false;
    earth.col.isStatic = false;
    earth.col.rest = 1.0f;
    a.objects.push_back(earth);
    // ... projectiles as its creation form builds them (src/main.cpp:1293-1307)
    const float spec[2][7] = {{10, 45, 0, 30, 1, 0, 0}, {5, 60, 90, 20, 2, 5, 0}};
    for (auto& s : spec) {
        object o;
        o.mass = s[0];
        o.Cd = 0.42f;
        o.pos = {s[5], s[6], 0};
        o.force = {0, 0, 0};
        o.vel = toSpherical(s[1], s[2], s[3]);
        o.transform = createTranslationMatrix(o.pos.x, o.pos.y, o.pos.z);
        o.area = M_PI * s[4] * s[4];
        o.col.center = o.pos;
        o.col.size = {s[4], s[4], s[4]};
        o.col.radius = 0.0f;
        o.col.isSphere = false;
        o.col.isStatic = false;
        o.col.rest = 0.5;
        o.stationary = false;
        a.objects.push_back(o);
    }
    // ... plus a dense ball of sphere colliders (reachable by setting the struct fields)
    unsigned rng = 12345u;
    auto uni = [&]() { rng = rng * 1664525u + 1013904223u; return (float)(rng >> 8) / 16777216.0f; };
    for (int i = 0; i < spheres; ++i) {
        object o;
        o.mass = 1.0f + 9.0f * uni();
        o.Cd = 0.42f;
        o.pos = {8.0f * uni() - 4.0f, 30.0f + 8.0f * uni(), 8.0f * uni() - 4.0f};
        o.vel = {2.0f * uni() - 1.0f, 2.0f * uni() - 1.0f, 2.0f * uni() - 1.0f};
        o.force = {0, 0, 0};
        o.transform = createTranslationMatrix(o.pos.x, o.pos.y, o.pos.z);
        o.area = M_PI * 0.5f * 0.5f;
        o.col.center = o.pos;
        o.col.size = {0.5f, 0.5f, 0.5f};
        o.col.radius = 0.5f;
        o.col.isSphere = true;
        o.col.isStatic = (i % 17) == 0;
        o.col.rest = 0.5f;
        o.stationary = (i % 23) == 0;
        a.objects.push_back(o);
    }
    world b = a;

    std::ofstream ca, cb;
    try {
        for (int s = 0; s < steps; ++s) {
            if (chdir(dirA.c_str()) != 0) return 3;
            updatePhysics(a, ca);                    // the reference's own step (src/main.cpp:748)
            if (chdir(dirB.c_str()) != 0) return 3;
            traceit_b200::updatePhysics(b, cb);      // the B200 step, same types, same call shape
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }

    size_t bad = 0;
    for (size_t i = 0; i < a.objects.size(); ++i) {
        const object& x = a.objects[i];
        const object& y = b.objects[i];
        bool ok = same_vec(x.pos, y.pos) && same_vec(x.vel, y.vel) && same_vec(x.force, y.force) &&
                  same_vec(x.col.center, y.col.center) && x.trajWritten == y.trajWritten &&
                  x.trajectory.size() == y.trajectory.size();
        for (size_t k = 0; ok && k < x.trajectory.size(); ++k) ok = same_vec(x.trajectory[k], y.trajectory[k]);
        if (ok && !x.stationary)
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) ok = ok && same_bits(x.transform.m[r][c], y.transform.m[r][c]);
        if (!ok && bad++ < 5) std::fprintf(stderr, "body %zu differs: ref pos %g %g %g gpu pos %g %g %g\n", i, x.pos.x, x.pos.y, x.pos.z, y.pos.x, y.pos.y, y.pos.z);
    }
    // generated NaNs print as "-nan" on x86 (default NaN 0xFFC00000) and "nan" from the GPU's
    // 0x7FFFFFFF: same NaN == NaN rule as everywhere else in the parity suite
    auto strip_nan_sign = [](std::string t) {
        for (size_t p = t.find("-nan"); p != std::string::npos; p = t.find("-nan", p)) t.erase(p, 1);
        return t;
    };
    const std::string fa = strip_nan_sign(slurp(dirA + "/Координаты.txt")), fb = strip_nan_sign(slurp(dirB + "/Координаты.txt"));
    const bool files = !fa.empty() && fa == fb;
    std::printf("bodies %zu steps %d mismatching_bodies %zu trajectory_files_equal %d (%zu bytes)\n", a.objects.size(), steps, bad,
                files ? 1 : 0, fa.size());
    if (bad == 0 && files) std::printf("SHIM_EQUALS_REFERENCE\n");
    return (bad == 0 && files) ? 0 : 1;
}

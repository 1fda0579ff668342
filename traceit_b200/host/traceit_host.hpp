// traceit_host.hpp - C++ host shim: the reference's per-frame call, on the GPU.
//
// Restores the reference's own entry point
//     void updatePhysics(world& w, ofstream& coords)            reference src/main.cpp:748
// (called once per frame at src/main.cpp:1589) on top of the C ABI in
// include/traceit_b200.h.  The shim is a template over the caller's world/object types, so
// it binds to the reference's OWN structs when included from its main.cpp
// and to the mirror structs of the headless harness here.  Required members, named exactly
// as in the reference (src/main.cpp:79-126):
//     World : objects (std::vector<Object>), atmDensity
//     Object: pos, vel, force {x,y,z}; mass, Cd, area, stationary;
//             col { center, size {x,y,z}, radius, isSphere, isStatic, rest }
// optional (used when present): trajectory, maxTrajectoryPoints, trajWritten, transform.
//
// The world lives on the device between frames.  Per call:
//   1. every object's hot fields are packed and compared with the shim's shadow copy (what the
//      device holds); bodies the host code changed since the last frame - a force it set, a body
//      it moved - are re-uploaded (tr_write_bodies, 92 B each); usually that is nothing;
//   2. tr_step_frame runs one step on the B200 and reads back what a step changes: pos, vel and
//      the collider centre (36 B per body);
//   3. those are written into the objects (force is zeroed for non-stationary bodies as
//      main.cpp:811 does) and the world's side effects are reproduced on the host:
//   * trajectory history + one-shot dump to "Координаты.txt"   src/main.cpp:754-761,793-808,814
//     (the reference pushes the position right after integration, i.e. before the collision
//      nudges; that value is the refreshed collider centre, src/main.cpp:791)
//   * object.transform = translation(pos)                      src/main.cpp:641-644,810
// Steps 1 and 3 run on a few host threads for large worlds (each object is touched by one thread).
// Error behaviour: the reference step returns void and cannot fail; here a failure (no
// B200, CUDA error, contact-buffer overflow) throws std::runtime_error with tr_last_error()
// - there is no CPU fallback to fall into.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/traceit_b200.h"

namespace traceit_b200 {

namespace detail {

template <class T, class = void> struct has_trajectory : std::false_type {};
template <class T>
struct has_trajectory<T, std::void_t<decltype(std::declval<T&>().trajectory), decltype(std::declval<T&>().trajWritten),
                                     decltype(std::declval<T&>().maxTrajectoryPoints)>> : std::true_type {};
template <class T, class = void> struct has_transform : std::false_type {};
template <class T> struct has_transform<T, std::void_t<decltype(std::declval<T&>().transform.m[0][0])>> : std::true_type {};

inline void check(int rc, const char* what) {
    if (rc != TR_OK) throw std::runtime_error(std::string("traceit_b200: ") + what + ": " + tr_last_error());
}

struct Session {
    tr_world* w = nullptr;
    std::vector<tr_body_desc> shadow;    // the device's copy of every body, as the host last saw / wrote it
    std::vector<tr_frame_body> frame;    // per-frame read-back
    size_t n = 0;
    ~Session() { if (w) tr_world_destroy(w); }
};

struct FrameBytes {
    unsigned long long h2d = 0, d2h = 0;
    double pack_ms = 0, step_ms = 0, unpack_ms = 0;   // host pack + compare (+ upload of changes) | step + read-back | host unpack + history
};
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline FrameBytes& frame_bytes() {
    static FrameBytes b;
    return b;
}

// fn(lo, hi, thread) over [0, n) on up to 16 host threads (serial for small worlds)
template <class Fn>
inline void parallel_for(size_t n, Fn fn) {
    unsigned t = std::thread::hardware_concurrency();
    t = t < 1 ? 1 : (t > 16 ? 16 : t);
    if (n < 32768 || t == 1) {
        fn((size_t)0, n, 0u);
        return;
    }
    std::vector<std::thread> pool;
    const size_t per = (n + t - 1) / t;
    for (unsigned k = 0; k < t; ++k) {
        const size_t lo = per * k, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        pool.emplace_back([=] { fn(lo, hi, k); });
    }
    for (std::thread& th : pool) th.join();
}

// One device world per host world object (keyed by address), created on first use.
inline std::map<const void*, Session>& sessions() {
    static std::map<const void*, Session> s;
    return s;
}

template <class Object>
inline void to_desc(const Object& o, tr_body_desc& d) {
    d.pos[0] = o.pos.x; d.pos[1] = o.pos.y; d.pos[2] = o.pos.z;
    d.vel[0] = o.vel.x; d.vel[1] = o.vel.y; d.vel[2] = o.vel.z;
    d.force[0] = o.force.x; d.force[1] = o.force.y; d.force[2] = o.force.z;
    d.mass = o.mass; d.Cd = o.Cd; d.area = o.area;
    d.stationary = o.stationary ? 1 : 0;
    d.center[0] = o.col.center.x; d.center[1] = o.col.center.y; d.center[2] = o.col.center.z;
    d.size[0] = o.col.size.x; d.size[1] = o.col.size.y; d.size[2] = o.col.size.z;
    d.radius = o.col.radius;
    d.isSphere = o.col.isSphere ? 1 : 0;
    d.isStatic = o.col.isStatic ? 1 : 0;
    d.rest = o.col.rest;
}

// what one step changes, into the object and into the shadow copy of it
template <class Object>
inline void from_frame(const tr_frame_body& f, Object& o, tr_body_desc& d) {
    o.pos.x = d.pos[0] = f.pos[0]; o.pos.y = d.pos[1] = f.pos[1]; o.pos.z = d.pos[2] = f.pos[2];
    o.vel.x = d.vel[0] = f.vel[0]; o.vel.y = d.vel[1] = f.vel[1]; o.vel.z = d.vel[2] = f.vel[2];
    o.col.center.x = d.center[0] = f.center[0]; o.col.center.y = d.center[1] = f.center[1]; o.col.center.z = d.center[2] = f.center[2];
    if (!o.stationary) {   // obj.force = {0,0,0} at the end of the non-stationary branch (main.cpp:764,811)
        o.force.x = o.force.y = o.force.z = 0.0f;
        d.force[0] = d.force[1] = d.force[2] = 0.0f;
    }
}

}  // namespace detail

// World-level knobs the reference hard-codes inside the step (g, dt: src/main.cpp:749-750)
// or leaves dead (pair gravity: src/main.cpp:679-693).  Defaults = reference behaviour.
struct StepOptions {
    bool pair_gravity = false;
    float G = 6.6743e-11f;
    bool collisions = true;
    bool write_trajectory_file = true;   // the "Координаты.txt" side effect
    int device = 0;
};

inline StepOptions& options() {
    static StepOptions o;
    return o;
}

// Drop the device world bound to `w` (e.g. before the host world is destroyed).
template <class World>
inline void release(World& w) { detail::sessions().erase(static_cast<const void*>(&w)); }

// The reference's per-frame entry point, same signature (src/main.cpp:748).
template <class World>
void updatePhysics(World& w, std::ofstream& coords) {
    using Object = typename std::decay<decltype(w.objects[0])>::type;
    detail::Session& s = detail::sessions()[static_cast<const void*>(&w)];
    const StepOptions& opt = options();
    const size_t n = w.objects.size();

    tr_world_params p;
    detail::check(tr_default_params(&p), "tr_default_params");
    p.atm_density = w.atmDensity;            // world::atmDensity, read by the drag term (main.cpp:777)
    p.G = opt.G;
    p.flags = (opt.pair_gravity ? TR_PAIR_GRAVITY : 0u) | (opt.collisions ? TR_COLLISIONS : 0u);

    if (s.w && s.n != n) {                   // bodies were pushed or erased: rebuild the device world
        tr_world_destroy(s.w);
        s.w = nullptr;
        s.shadow.clear();
    }
    if (!s.w) {
        detail::check(tr_world_create(&p, opt.device, &s.w), "tr_world_create");
        s.n = n;
    } else {
        detail::check(tr_world_set_params(s.w, &p), "tr_world_set_params");
    }

    static bool headerWritten = false;       // same one-shot header as main.cpp:754-761
    if (opt.write_trajectory_file) {
        coords.open("Координаты.txt", std::ios::app);
        if (!headerWritten) {
            coords << "Координаты в формате (x, y, z, время):\n";
            headerWritten = true;
        }
    }

    detail::FrameBytes& fb = detail::frame_bytes();
    fb = detail::FrameBytes{};
    if (n) {
        const double t_pack0 = detail::now_ms();
        if (s.shadow.size() != n) {
            // first frame of this world: create every body on the device
            s.shadow.resize(n);
            s.frame.resize(n);
            detail::parallel_for(n, [&](size_t lo, size_t hi, unsigned) {
                for (size_t i = lo; i < hi; ++i) {
                    std::memset(&s.shadow[i], 0, sizeof(tr_body_desc));
                    detail::to_desc(w.objects[i], s.shadow[i]);
                }
            });
            detail::check(tr_bodies_add(s.w, n, s.shadow.data(), nullptr), "tr_bodies_add");
            fb.h2d += n * sizeof(tr_body_desc);
        } else {
            // what did the host code change since the last frame?  (bit compare: NaNs and signed zeros included)
            std::vector<std::pair<size_t, size_t>> dirty(17, {n, 0});
            detail::parallel_for(n, [&](size_t lo, size_t hi, unsigned k) {
                size_t dlo = n, dhi = 0;
                for (size_t i = lo; i < hi; ++i) {
                    tr_body_desc d;
                    std::memset(&d, 0, sizeof(d));
                    detail::to_desc(w.objects[i], d);
                    if (std::memcmp(&d, &s.shadow[i], sizeof(d)) != 0) {
                        s.shadow[i] = d;
                        dlo = std::min(dlo, i);
                        dhi = std::max(dhi, i + 1);
                    }
                }
                dirty[k] = {dlo, dhi};
            });
            size_t dlo = n, dhi = 0;
            for (const auto& d : dirty) {
                dlo = std::min(dlo, d.first);
                dhi = std::max(dhi, d.second);
            }
            if (dlo < dhi) {
                detail::check(tr_write_bodies(s.w, dlo, dhi - dlo, s.shadow.data() + dlo), "tr_write_bodies");
                fb.h2d += (dhi - dlo) * sizeof(tr_body_desc);
            }
        }
        const double t_step0 = detail::now_ms();
        fb.pack_ms = t_step0 - t_pack0;
        detail::check(tr_step_frame(s.w, n, nullptr, s.frame.data(), 1), "tr_step_frame");
        fb.d2h += n * sizeof(tr_frame_body);
        const double t_unpack0 = detail::now_ms();
        fb.step_ms = t_unpack0 - t_step0;
        std::vector<std::vector<size_t>> dump(17);
        detail::parallel_for(n, [&](size_t lo, size_t hi, unsigned k) {
            for (size_t i = lo; i < hi; ++i) {
                Object& o = w.objects[i];
                detail::from_frame(s.frame[i], o, s.shadow[i]);
                if (o.stationary) continue;      // the reference touches only non-stationary bodies below (main.cpp:764)
                if constexpr (detail::has_trajectory<Object>::value) {
                    // position right after integration == refreshed collider centre
                    o.trajectory.push_back(o.col.center);
                    if (o.trajectory.size() == (size_t)o.maxTrajectoryPoints && !o.trajWritten) {
                        dump[k].push_back(i);    // written below, in object order, by one thread
                        o.trajWritten = true;
                    }
                    if (o.trajectory.size() > (size_t)o.maxTrajectoryPoints) o.trajectory.erase(o.trajectory.begin());
                }
                if constexpr (detail::has_transform<Object>::value) {
                    // createTranslationMatrix(centre) - note: the reference builds it from the
                    // post-integration, pre-collision position (main.cpp:810 runs before 816)
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) o.transform.m[r][c] = (r == c) ? 1.0f : 0.0f;
                    o.transform.m[0][3] = o.col.center.x;   // translation sits in the 4th column (main.cpp:483-493)
                    o.transform.m[1][3] = o.col.center.y;
                    o.transform.m[2][3] = o.col.center.z;
                }
            }
        });
        if constexpr (detail::has_trajectory<Object>::value) {
            if (opt.write_trajectory_file)
                for (const auto& list : dump)          // thread ranges are ascending: object order, as the reference's loop
                    for (size_t i : list) {
                        coords << "\nТраектория объекта [" << "obj.id" << "]:\n";
                        for (const auto& pt : w.objects[i].trajectory) coords << pt.x << "(м)\t" << pt.y << "(м)\t" << pt.z << "(м)\n";
                    }
        }
        fb.unpack_ms = detail::now_ms() - t_unpack0;
    }
    if (opt.write_trajectory_file) coords.close();
}

// bytes the last updatePhysics call moved over PCIe (observability for the harness / bench.py)
inline unsigned long long last_frame_h2d_bytes() { return detail::frame_bytes().h2d; }
inline unsigned long long last_frame_d2h_bytes() { return detail::frame_bytes().d2h; }
// where the last frame's time went: host pack + compare (+ upload of what changed), the step + read-back, host unpack + history
inline void last_frame_breakdown(double& pack_ms, double& step_ms, double& unpack_ms) {
    pack_ms = detail::frame_bytes().pack_ms;
    step_ms = detail::frame_bytes().step_ms;
    unpack_ms = detail::frame_bytes().unpack_ms;
}

// toSpherical of the creation form (src/main.cpp:1080-1089), any {x,y,z} aggregate.
template <class Vec3>
inline Vec3 toSpherical(const float& elev, const float& azim, const float& magn) {
    float v[3];
    tr_to_spherical(elev, azim, magn, v);
    return Vec3{v[0], v[1], v[2]};
}

}  // namespace traceit_b200

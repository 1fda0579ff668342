// The whole live reference step for a SMALL world in ONE kernel launch of ONE CTA.
//
// The reference's own use is a handful of bodies stepped once per 60 Hz frame (the seeded `earth`
// slab plus a few projectiles, src/main.cpp:1482-1491, 1206-1311).  For such a world the grid
// pipeline is all launch latency (round 1: 47 us per step for 3 bodies, 13 launches).  Up to TINY_MAX
// bodies the step is instead exactly the reference's algorithm, with the O(N^2) detection spread over
// the threads of one CTA and everything in shared memory:
//   integrate   one thread per body: drag + uniform gravity + semi-implicit Euler (integrator.cuh, main.cpp:763-791,811)
//   history     trajectory ring push (main.cpp:794), if enabled
//   detect      thread i tests the pairs (i, j > i) of its row with the reference's guards and predicates
//               (main.cpp:695-707): a count pass, a CTA scan, a fill pass -> the contact list comes out in
//               lexicographic (i,j) order, the reference's visiting order, without any sort
//   respond     thread 0 replays the list sequentially on the shared-memory state (main.cpp:709-742):
//               literally the reference's Gauss-Seidel sweep
// and tr_step(w, k) loops k steps INSIDE the kernel.  Bit-exact like the big pipeline (same device
// functions, same rounded intrinsics); no grid, no side list, no classification, no capacity limit
// (the list holds N(N-1)/2 pairs).  Worlds with pair gravity on, sharded worlds and worlds above
// TINY_MAX bodies take the general path.
#pragma once
#include "collide_common.cuh"
#include "collide_response.cuh"
#include "integrator.cuh"

namespace tr {

constexpr int TINY_MAX = 256;
constexpr int TINY_SMEM_PAIRS = 4096;   // contacts kept in shared memory for the replay (beyond: read back from the global list)

struct TinyArgs {
    float4* pos_m;
    float4* vel_rest;
    float4* force;
    const float2* drag;
    float4* center_r;
    const float4* half_ext;
    const uint8_t* flags;
    uint32_t n;
    StepConsts c;
    int collide;
    unsigned long long* ckeys;   // out: contact list of the LAST step, lexicographic; capacity >= n(n-1)/2
    uint32_t* counters;
    uint32_t* ctl;
    StepRecord* ring;
    float* traj;
    uint32_t* traj_cnt;
    uint32_t traj_max;
    unsigned long long traj_stride;
    float* frame_out;            // optional (mapped pinned host memory): pos, vel, centre of every body after the last step, 9 floats each
};
static_assert(TINY_MAX == (int)TINY_FRAME_BODIES, "frame buffer sized for the single-CTA path");

// guards + predicate of one (i<j) visit, main.cpp:695-707
__device__ __forceinline__ bool tiny_contact(uint8_t fi, uint8_t fj, const float4 ci, const float4 cj, const float4 hi, const float4 hj) {
    if ((fi & F_STATIC) && (fj & F_STATIC)) return false;        // main.cpp:695
    if (!(fi & fj & F_LIVE)) return false;                        // main.cpp:696 (mass <= 0)
    if ((fi & F_SPHERE) && (fj & F_SPHERE)) return sphere_hit(ci, cj);
    return aabb_hit(ci, hi, cj, hj);
}

__global__ void __launch_bounds__(TINY_MAX) tiny_step_kernel(TinyArgs a, int nsteps) {
    __shared__ float4 s_pos[TINY_MAX], s_vel[TINY_MAX], s_cen[TINY_MAX], s_half[TINY_MAX];
    __shared__ uint8_t s_flag[TINY_MAX];
    __shared__ uint32_t s_off[TINY_MAX + 1];
    __shared__ uint32_t s_pairs[TINY_SMEM_PAIRS];
    __shared__ unsigned long long s_t0;
    const uint32_t i = threadIdx.x, n = a.n;
    const bool mine = i < n;
    float4 frc = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 dr = make_float2(0.f, 0.f);
    if (mine) {
        s_pos[i] = a.pos_m[i];
        s_vel[i] = a.vel_rest[i];
        s_cen[i] = a.center_r[i];
        s_half[i] = a.half_ext[i];
        s_flag[i] = a.flags[i];
        frc = a.force[i];
        dr = a.drag[i];
    }
    __syncthreads();
    uint32_t total = 0;
    for (int step = 0; step < nsteps; ++step) {
        if (i == 0) s_t0 = globaltimer_ns();
        // ---- integrate (non-stationary bodies only, main.cpp:764)
        if (mine && !(s_flag[i] & F_STATIONARY)) {
            float4 pm = s_pos[i], vr = s_vel[i];
            integrate_body(pm, vr, frc.x, frc.y, frc.z, dr, a.c);
            s_pos[i] = pm;
            s_vel[i] = vr;
            s_cen[i] = make_float4(pm.x, pm.y, pm.z, s_cen[i].w);       // updateCollider, main.cpp:791
            frc = make_float4(0.f, 0.f, 0.f, 0.f);                       // main.cpp:811
            if (a.traj_max) {                                            // trajectory.push_back(obj.pos), main.cpp:794
                const uint32_t k = a.traj_cnt[i];
                const unsigned long long slot = k % a.traj_max;
                a.traj[(slot * 3 + 0) * a.traj_stride + i] = pm.x;
                a.traj[(slot * 3 + 1) * a.traj_stride + i] = pm.y;
                a.traj[(slot * 3 + 2) * a.traj_stride + i] = pm.z;
                a.traj_cnt[i] = k + 1;
            }
        }
        __syncthreads();
        if (!a.collide) continue;
        // ---- detect: row i of the reference's i<j sweep, count then fill
        uint32_t cnt = 0;
        if (mine) {
            const uint8_t fi = s_flag[i];
            const float4 ci = s_cen[i], hi = s_half[i];
            for (uint32_t j = i + 1; j < n; ++j) cnt += tiny_contact(fi, s_flag[j], ci, s_cen[j], hi, s_half[j]) ? 1u : 0u;
        }
        s_off[i] = cnt;
        __syncthreads();
        if (i == 0) {                                   // <= 256 rows: a serial scan is a few hundred cycles
            uint32_t run = 0;
            for (uint32_t k = 0; k < n; ++k) {
                const uint32_t v = s_off[k];
                s_off[k] = run;
                run += v;
            }
            s_off[n] = run;
        }
        __syncthreads();
        total = s_off[n];
        if (mine && cnt) {
            const uint8_t fi = s_flag[i];
            const float4 ci = s_cen[i], hi = s_half[i];
            uint32_t w = s_off[i];
            for (uint32_t j = i + 1; j < n; ++j)
                if (tiny_contact(fi, s_flag[j], ci, s_cen[j], hi, s_half[j])) {
                    TR_ASSERT(w < s_off[i + 1] && w < n * (n - 1u) / 2u);
                    if (w < (uint32_t)TINY_SMEM_PAIRS) s_pairs[w] = (i << 16) | j;
                    a.ckeys[w] = ((unsigned long long)i << 32) | j;
                    ++w;
                }
        }
        __syncthreads();
        // ---- respond: the reference's sequential sweep over the list (main.cpp:709-742)
        if (i == 0) {
            for (uint32_t c = 0; c < total; ++c) {
                uint32_t bi, bj;
                if (c < (uint32_t)TINY_SMEM_PAIRS) {
                    bi = s_pairs[c] >> 16;
                    bj = s_pairs[c] & 0xFFFFu;
                } else {
                    const unsigned long long k = a.ckeys[c];
                    bi = (uint32_t)(k >> 32);
                    bj = (uint32_t)k;
                }
                float4 vi = s_vel[bi], vj = s_vel[bj], pi = s_pos[bi], pj = s_pos[bj];
                respond_values(s_flag[bi], s_flag[bj], s_cen[bi], s_cen[bj], s_half[bi], s_half[bj], vi, vj, pi, pj);
                s_vel[bi] = vi;
                s_vel[bj] = vj;
                s_pos[bi] = pi;
                s_pos[bj] = pj;
            }
        }
        __syncthreads();
    }
    if (mine) {
        a.pos_m[i] = s_pos[i];
        a.vel_rest[i] = s_vel[i];
        a.center_r[i] = s_cen[i];
        a.force[i] = frc;
        if (a.frame_out) {       // tr_step_frame's read-back without a pack kernel and a copy: 36 bytes over PCIe by this thread
            float* f = a.frame_out + 9 * i;
            f[0] = s_pos[i].x; f[1] = s_pos[i].y; f[2] = s_pos[i].z;
            f[3] = s_vel[i].x; f[4] = s_vel[i].y; f[5] = s_vel[i].z;
            f[6] = s_cen[i].x; f[7] = s_cen[i].y; f[8] = s_cen[i].z;
        }
    }
    if (i == 0 && a.collide && nsteps > 0) {
        // one record for the launch (the last step's contact count; the stamps bracket the last step)
        const uint32_t seq = a.ctl[CTL_SEQ];
        StepRecord r;
        r.flags = 0;
        r.contacts_lo = total;
        r.contacts_hi = 0;
        r.side = r.far = r.bnd = 0;
        r.grid = n;
        r.rounds = total;
        r.long_segments = 0;
        r.pad[0] = r.pad[1] = 0;
        r.t[0] = r.t[1] = r.t[2] = s_t0;
        r.t[3] = globaltimer_ns();
        r.seq = seq + 1u;
        a.ring[seq % STEP_RING] = r;
        a.ctl[CTL_SEQ] = seq + 1u;
    }
}

}  // namespace tr

// K6: contact ordering (prep), the contact itself, and the executors that replay the reference's
// Gauss-Seidel sweep.  Device code only; included by collide.cu after collide_common.cuh.
#pragma once
#include "collide_common.cuh"

namespace tr {

// ---------------------------------------------------------------------------------------
// K6 preparation: dependency structure of the lexicographic Gauss-Seidel order
// ---------------------------------------------------------------------------------------
// The contact list is put into the reference's visiting order - lexicographic (i,j) - and into
// "column" order (j, then i) with the same counting-sort machinery as the broad phase, keyed by
// body id: histogram (RED) -> two-kernel scan -> scatter into the body's segment (atomicSub counts
// the histogram back to zero) -> rank inside the segment.  The scans leave DENSE tables
// row_start[N+1] / col_start[N+1], so "first / last contact of body b" is a lookup.
__global__ void __launch_bounds__(256) chist_kernel(const unsigned long long* __restrict__ ckeys, uint32_t c_n, int shift,
                                                    uint32_t* __restrict__ cnt) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < c_n) atomicAdd(&cnt[(uint32_t)(ckeys[c] >> shift)], 1u);
}

__global__ void __launch_bounds__(256) cscatter_row_kernel(const unsigned long long* __restrict__ ckeys, uint32_t c_n,
                                                           const uint32_t* __restrict__ row_start, uint32_t* __restrict__ cnt,
                                                           unsigned long long* __restrict__ tmp) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_n) return;
    const unsigned long long k = ckeys[c];
    const uint32_t i = (uint32_t)(k >> 32);
    tmp[row_start[i] + atomicSub(&cnt[i], 1u) - 1u] = k;
}

// rows: ascending j inside body i's segment (j is unique inside a row)
__global__ void __launch_bounds__(256) crank_row_kernel(const unsigned long long* __restrict__ tmp,
                                                        const uint32_t* __restrict__ row_start, uint32_t c_n,
                                                        unsigned long long* __restrict__ sorted) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c_n) return;
    const unsigned long long k = tmp[t];
    const uint32_t i = (uint32_t)(k >> 32);
    const uint32_t s = row_start[i], e = row_start[i + 1];
    uint32_t rank = 0;
    for (uint32_t u = s; u < e; ++u) rank += (tmp[u] < k) ? 1u : 0u;
    sorted[s + rank] = k;
}

// columns: entry = (i << 32 | c) of the contact c = (i,j) in row order, scattered to body j's segment
__global__ void __launch_bounds__(256) cscatter_col_kernel(const unsigned long long* __restrict__ ckeys, uint32_t c_n,
                                                           const uint32_t* __restrict__ col_start, uint32_t* __restrict__ cnt,
                                                           unsigned long long* __restrict__ tmp) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_n) return;
    const unsigned long long k = ckeys[c];
    const uint32_t j = (uint32_t)k;
    tmp[col_start[j] + atomicSub(&cnt[j], 1u) - 1u] = (k & 0xFFFFFFFF00000000ull) | c;
}

__global__ void __launch_bounds__(256) crank_col_kernel(const unsigned long long* __restrict__ tmp,
                                                        const unsigned long long* __restrict__ ckeys,
                                                        const uint32_t* __restrict__ col_start, uint32_t c_n,
                                                        uint32_t* __restrict__ colc, uint32_t* __restrict__ col_pos) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c_n) return;
    const unsigned long long e = tmp[t];
    const uint32_t c = (uint32_t)e;
    const uint32_t j = (uint32_t)ckeys[c];
    const uint32_t s = col_start[j], en = col_start[j + 1];
    uint32_t rank = 0;
    for (uint32_t u = s; u < en; ++u) rank += (tmp[u] < e) ? 1u : 0u;   // ascending i (unique inside a column)
    colc[s + rank] = c;
    col_pos[c] = s + rank;
}

// Body b's contacts in sweep order are: its column entries (k,b) by ascending k, then its
// row entries (b,j) by ascending j.  pred/succ below are the neighbours of contact c in the
// lists of its two bodies.
// Everything about a contact that the sweep never changes - the contact normal (collider centres are
// not touched by the response, main.cpp:655-659 vs 725-741), the inverse masses, min(restitution),
// the isStatic bits - is computed once, in parallel, by the K6 prep and stored in one 32-byte
// record, so the square root and five of the six IEEE divisions of a contact are off the critical
// path of the ordered replay and a contact costs one load for all its constants.
struct __align__(32) ContactGeo {
    float nx, ny, nz;      // main.cpp:703 / calculateAABBNormal
    float ima, imb;        // 1/ma, 1/mb
    float den;             // (1/ma) + (1/mb)
    float e;               // std::min(a.rest, b.rest)
    uint32_t bits;         // 1: a.isStatic, 2: b.isStatic
};
static_assert(sizeof(ContactGeo) == 32, "one sector");

__device__ __forceinline__ void contact_normal(uint8_t fi, uint8_t fj, const float4 ci, const float4 cj, const float4* half_ext,
                                               uint32_t i, uint32_t j, float& nx, float& ny, float& nz) {
    if ((fi & F_SPHERE) && (fj & F_SPHERE)) {
        // normalize(b.center - a.center), main.cpp:703 + 211-223
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        float len = __fsqrt_rn(dot3_rn(dx, dy, dz, dx, dy, dz));
        if (len != 0.0f) {
            nx = __fdiv_rn(dx, len);
            ny = __fdiv_rn(dy, len);
            nz = __fdiv_rn(dz, len);
        } else {
            nx = dx; ny = dy; nz = dz;
        }
    } else {
        // calculateAABBNormal, main.cpp:622-638
        const float4 hi = half_ext[i], hj = half_ext[j];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        float ox = __fsub_rn(__fadd_rn(hi.x, hj.x), fabsf(dx));
        float oy = __fsub_rn(__fadd_rn(hi.y, hj.y), fabsf(dy));
        float oz = __fsub_rn(__fadd_rn(hi.z, hj.z), fabsf(dz));
        nx = ny = nz = 0.0f;
        if (ox < oy && ox < oz) nx = dx > 0 ? 1.0f : -1.0f;
        else if (oy < oz) ny = dy > 0 ? 1.0f : -1.0f;
        else nz = dz > 0 ? 1.0f : -1.0f;
    }
}

// Dataflow executors keep the mutable state of a body that is in contact in ONE 32-byte record
// whose two 16-byte halves both carry a version = number of contacts applied to the body so far.
// A contact knows the versions it must see (its rank in the sweep order of each of its bodies),
// so a reader validates what it loaded by itself and the writer needs no fence between the data
// and the hand-off (the fence was 42 % of the stall samples of the ordered replay).
struct __align__(32) BodyRec {
    float vx, vy, vz;
    uint32_t ver_v;
    float px, py, pz;
    uint32_t ver_p;
};
static_assert(sizeof(BodyRec) == 32, "one sector");

__device__ __forceinline__ void st_body(BodyRec* p, const float4 v, const float4 x, uint32_t ver) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)),
                 "r"(__float_as_uint(v.z)), "r"(ver), "r"(__float_as_uint(x.x)), "r"(__float_as_uint(x.y)),
                 "r"(__float_as_uint(x.z)), "r"(ver)
                 : "memory");
}
// L2 load (the record is written by other SMs); true when both halves carry the expected version
__device__ __forceinline__ bool ld_body(const BodyRec* p, uint32_t want, float4& v, float4& x) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "l"(p)
                 : "memory");
    v.x = __uint_as_float(r0); v.y = __uint_as_float(r1); v.z = __uint_as_float(r2);
    x.x = __uint_as_float(r4); x.y = __uint_as_float(r5); x.z = __uint_as_float(r6);
    return r3 == want && r7 == want;
}

__global__ void __launch_bounds__(256) deps_kernel(const unsigned long long* __restrict__ ckeys,
                                                   const uint32_t* __restrict__ colc, const uint32_t* __restrict__ col_pos,
                                                   const uint32_t* __restrict__ row_start,
                                                   const uint32_t* __restrict__ col_start, uint32_t c_n,
                                                   uint32_t* __restrict__ dep, uint32_t* __restrict__ succ_i,
                                                   uint32_t* __restrict__ succ_j, uint4* __restrict__ crec,
                                                   uint32_t* __restrict__ frontier0, uint32_t* counters, int ready_counter,
                                                   const float4* __restrict__ pos_m, const float4* __restrict__ vel_rest,
                                                   const float4* __restrict__ center_r, const float4* __restrict__ half_ext,
                                                   const uint8_t* __restrict__ flags, BodyRec* __restrict__ brec,
                                                   uint2* __restrict__ cver, ContactGeo* __restrict__ cgeo) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_n) return;
    unsigned long long k = ckeys[c];
    uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
    uint32_t t = col_pos[c];
    // contacts applied to each body before this one: column entries of i + earlier row entries; earlier column entries of j
    const uint32_t ei = (col_start[i + 1] - col_start[i]) + (c - row_start[i]);
    const uint32_t ej = t - col_start[j];
    uint32_t d = (ei ? 1u : 0u) + (ej ? 1u : 0u);
    dep[c] = d;
    const uint32_t si = (c + 1 < row_start[i + 1]) ? c + 1 : EMPTY;
    uint32_t sj;
    if (t + 1 < col_start[j + 1]) sj = colc[t + 1];
    else sj = (row_start[j + 1] > row_start[j]) ? row_start[j] : EMPTY;  // j's first row entry follows its last column entry
    succ_i[c] = si;
    succ_j[c] = sj;
    crec[c] = make_uint4(i, j, si, sj);
    if (brec) {                                   // dataflow executors: expected versions, version-0 records, constants
        cver[c] = make_uint2(ei, ej);
        const float4 vi = vel_rest[i], vj = vel_rest[j], pi = pos_m[i], pj = pos_m[j];
        if (ei == 0) st_body(brec + i, vi, pi, 0u);
        if (ej == 0) st_body(brec + j, vj, pj, 0u);
        const uint8_t fi = flags[i], fj = flags[j];
        ContactGeo g;
        contact_normal(fi, fj, center_r[i], center_r[j], half_ext, i, j, g.nx, g.ny, g.nz);
        g.ima = __fdiv_rn(1.0f, pi.w);
        g.imb = __fdiv_rn(1.0f, pj.w);
        g.den = __fadd_rn(g.ima, g.imb);
        g.e = (vj.w < vi.w) ? vj.w : vi.w;               // std::min(a.rest, b.rest)
        g.bits = ((fi & F_STATIC) ? 1u : 0u) | ((fj & F_STATIC) ? 2u : 0u);
        cgeo[c] = g;
    }
    if (d == 0) frontier0[atomicAdd(&counters[ready_counter], 1u)] = c;
}

// ---------------------------------------------------------------------------------------
// K6: the response itself (one contact), main.cpp:701-742, bit for bit
// ---------------------------------------------------------------------------------------
struct RespArgs {
    float4* pos_m;
    float4* vel_rest;
    const float4* center_r;
    const float4* half_ext;
    const uint8_t* flags;
};

__device__ __forceinline__ void respond(const RespArgs& a, uint32_t i, uint32_t j) {
    const uint8_t fi = a.flags[i], fj = a.flags[j];
    const float4 ci = a.center_r[i], cj = a.center_r[j];
    float nx, ny, nz;
    if ((fi & F_SPHERE) && (fj & F_SPHERE)) {
        // normalize(b.center - a.center), main.cpp:703 + 211-223
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        float len = __fsqrt_rn(dot3_rn(dx, dy, dz, dx, dy, dz));
        if (len != 0.0f) {
            nx = __fdiv_rn(dx, len);
            ny = __fdiv_rn(dy, len);
            nz = __fdiv_rn(dz, len);
        } else {
            nx = dx; ny = dy; nz = dz;
        }
    } else {
        // calculateAABBNormal, main.cpp:622-638
        const float4 hi = a.half_ext[i], hj = a.half_ext[j];
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        float ox = __fsub_rn(__fadd_rn(hi.x, hj.x), fabsf(dx));
        float oy = __fsub_rn(__fadd_rn(hi.y, hj.y), fabsf(dy));
        float oz = __fsub_rn(__fadd_rn(hi.z, hj.z), fabsf(dz));
        nx = ny = nz = 0.0f;
        if (ox < oy && ox < oz) nx = dx > 0 ? 1.0f : -1.0f;
        else if (oy < oz) ny = dy > 0 ? 1.0f : -1.0f;
        else nz = dz > 0 ? 1.0f : -1.0f;
    }
    // L2 loads: these were written by other SMs in earlier wavefronts.  Positions are fetched in
    // the same round trip (needed only when the pair is not skipped, but a second dependent trip
    // would sit on the critical path of the ordered replay).
    float4 vi = __ldcg(a.vel_rest + i), vj = __ldcg(a.vel_rest + j);
    float4 pi = __ldcg(a.pos_m + i), pj = __ldcg(a.pos_m + j);
    float rvx = __fsub_rn(vi.x, vj.x), rvy = __fsub_rn(vi.y, vj.y), rvz = __fsub_rn(vi.z, vj.z);
    float vn = dot3_rn(rvx, rvy, rvz, nx, ny, nz);
    if (vn > 0) return;                                  // main.cpp:714 (as written)
    float e = (vj.w < vi.w) ? vj.w : vi.w;               // std::min(a.rest, b.rest)
    float jimp = __fmul_rn(-__fadd_rn(1.0f, e), vn);     // -(1 + e) * vn
    float ima = __fdiv_rn(1.0f, pi.w), imb = __fdiv_rn(1.0f, pj.w);
    jimp = __fdiv_rn(jimp, __fadd_rn(ima, imb));
    float ix = __fmul_rn(nx, jimp), iy = __fmul_rn(ny, jimp), iz = __fmul_rn(nz, jimp);
    const float kcorr = 0.2f * 0.01f;                    // percent * penetrationSlop (fp32 product)
    float cx = __fmul_rn(nx, kcorr), cy = __fmul_rn(ny, kcorr), cz = __fmul_rn(nz, kcorr);
    if (!(fi & F_STATIC)) {
        vi.x = __fsub_rn(vi.x, __fmul_rn(ix, ima));
        vi.y = __fsub_rn(vi.y, __fmul_rn(iy, ima));
        vi.z = __fsub_rn(vi.z, __fmul_rn(iz, ima));
        pi.x = __fsub_rn(pi.x, __fmul_rn(cx, ima));
        pi.y = __fsub_rn(pi.y, __fmul_rn(cy, ima));
        pi.z = __fsub_rn(pi.z, __fmul_rn(cz, ima));
        a.vel_rest[i] = vi;
        a.pos_m[i] = pi;
    }
    if (!(fj & F_STATIC)) {
        vj.x = __fadd_rn(vj.x, __fmul_rn(ix, imb));
        vj.y = __fadd_rn(vj.y, __fmul_rn(iy, imb));
        vj.z = __fadd_rn(vj.z, __fmul_rn(iz, imb));
        pj.x = __fadd_rn(pj.x, __fmul_rn(cx, imb));
        pj.y = __fadd_rn(pj.y, __fmul_rn(cy, imb));
        pj.z = __fadd_rn(pj.z, __fmul_rn(cz, imb));
        a.vel_rest[j] = vj;
        a.pos_m[j] = pj;
    }
}

__device__ __forceinline__ void store_xyz(float4* p, const float4 v) {
    *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
    reinterpret_cast<float*>(p)[2] = v.z;
}

// The same contact on version-checked body records (dataflow executors).  Returns false when the
// watchdog fired.  Every path bumps both versions, also the ones that change nothing (static body,
// the approach test): the successors wait for exactly that version.
struct FlowArgs {
    const float4* pos_m;       // .w = mass (immutable here); xyz written back at a body's last contact
    const float4* vel_rest;    // .w = restitution (immutable)
    float4* pos_out;
    float4* vel_out;
    BodyRec* brec;
    const uint2* cver;
    const ContactGeo* cgeo;
};
constexpr uint32_t VERSION_SPIN_LIMIT = 1u << 20;

__device__ __forceinline__ bool respond_rec(const FlowArgs& a, uint32_t c, uint32_t i, uint32_t j, bool last_i, bool last_j,
                                            uint32_t* counters) {
    const uint2 want = a.cver[c];
    float4 vi, vj, pi, pj;
    bool oki = ld_body(a.brec + i, want.x, vi, pi), okj = ld_body(a.brec + j, want.y, vj, pj);
    const ContactGeo g = a.cgeo[c];
    // the writer's stores may still be on their way when its hand-off arrives: poll until they land
    uint32_t spins = 0;
    while (!oki || !okj) {
        if (!oki) oki = ld_body(a.brec + i, want.x, vi, pi);
        if (!okj) okj = ld_body(a.brec + j, want.y, vj, pj);
        if (++spins > VERSION_SPIN_LIMIT) {
            atomicExch(&counters[C_ABORT], 1u);
            return false;
        }
    }
    float rvx = __fsub_rn(vi.x, vj.x), rvy = __fsub_rn(vi.y, vj.y), rvz = __fsub_rn(vi.z, vj.z);
    float vn = dot3_rn(rvx, rvy, rvz, g.nx, g.ny, g.nz);
    if (!(vn > 0)) {                                     // main.cpp:714 (as written)
        float jimp = __fmul_rn(-__fadd_rn(1.0f, g.e), vn);   // -(1 + e) * vn
        jimp = __fdiv_rn(jimp, g.den);                   // / (1/ma + 1/mb)
        float ix = __fmul_rn(g.nx, jimp), iy = __fmul_rn(g.ny, jimp), iz = __fmul_rn(g.nz, jimp);
        const float kcorr = 0.2f * 0.01f;                // percent * penetrationSlop (fp32 product)
        float cx = __fmul_rn(g.nx, kcorr), cy = __fmul_rn(g.ny, kcorr), cz = __fmul_rn(g.nz, kcorr);
        if (!(g.bits & 1u)) {
            vi.x = __fsub_rn(vi.x, __fmul_rn(ix, g.ima));
            vi.y = __fsub_rn(vi.y, __fmul_rn(iy, g.ima));
            vi.z = __fsub_rn(vi.z, __fmul_rn(iz, g.ima));
            pi.x = __fsub_rn(pi.x, __fmul_rn(cx, g.ima));
            pi.y = __fsub_rn(pi.y, __fmul_rn(cy, g.ima));
            pi.z = __fsub_rn(pi.z, __fmul_rn(cz, g.ima));
        }
        if (!(g.bits & 2u)) {
            vj.x = __fadd_rn(vj.x, __fmul_rn(ix, g.imb));
            vj.y = __fadd_rn(vj.y, __fmul_rn(iy, g.imb));
            vj.z = __fadd_rn(vj.z, __fmul_rn(iz, g.imb));
            pj.x = __fadd_rn(pj.x, __fmul_rn(cx, g.imb));
            pj.y = __fadd_rn(pj.y, __fmul_rn(cy, g.imb));
            pj.z = __fadd_rn(pj.z, __fmul_rn(cz, g.imb));
        }
    }
    // nobody reads the record of a body again after its last contact: back to the SoA state.  Only
    // x, y, z are written (8 + 4 bytes): the w components (restitution, mass) are already there, and
    // loading them to rebuild a float4 would park the store - and the atomics behind it - on an L2 round trip.
    if (last_i) {
        store_xyz(a.vel_out + i, vi);
        store_xyz(a.pos_out + i, pi);
    } else {
        st_body(a.brec + i, vi, pi, want.x + 1u);
    }
    if (last_j) {
        store_xyz(a.vel_out + j, vj);
        store_xyz(a.pos_out + j, pj);
    } else {
        st_body(a.brec + j, vj, pj, want.y + 1u);
    }
    return true;
}

// Persistent cooperative kernel: one wavefront per round, grid-wide barrier between rounds.
// counters[8 + r%3] is the size of round r's frontier.
__global__ void __launch_bounds__(256) response_kernel(RespArgs a, const unsigned long long* __restrict__ ckeys,
                                                       uint32_t* dep, const uint32_t* __restrict__ succ_i,
                                                       const uint32_t* __restrict__ succ_j, uint32_t* f0, uint32_t* f1,
                                                       uint32_t* counters) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gsize = gridDim.x * blockDim.x;
    uint32_t round = 0;
    while (true) {
        const uint32_t n = *((volatile uint32_t*)&counters[8 + round % 3]);
        if (n == 0) break;
        uint32_t* fin = (round & 1) ? f1 : f0;
        uint32_t* fout = (round & 1) ? f0 : f1;
        uint32_t* cnt_out = &counters[8 + (round + 1) % 3];
        if (gtid == 0) counters[8 + (round + 2) % 3] = 0;
        for (uint32_t idx = gtid; idx < n; idx += gsize) {
            const uint32_t c = __ldcg(fin + idx);
            const unsigned long long k = ckeys[c];
            respond(a, (uint32_t)(k >> 32), (uint32_t)k);
            __threadfence();
            uint32_t s = succ_i[c];
            if (s != EMPTY && atomicSub(&dep[s], 1u) == 1u) fout[atomicAdd(cnt_out, 1u)] = s;
            s = succ_j[c];
            if (s != EMPTY && atomicSub(&dep[s], 1u) == 1u) fout[atomicAdd(cnt_out, 1u)] = s;
        }
        grid.sync();
        ++round;
    }
    if (gtid == 0) counters[11] = round;
}

// Barrier-free variant of K6 ("dataflow"): the same dependency graph, but instead of one grid-wide
// barrier per wavefront a thread that finishes a contact releases its (at most two) successors
// and CONTINUES with one that became ready; a second one goes to a global FIFO that idle threads
// drain.  A contact still runs only after the previous contact of both its bodies, and never
// concurrently with another contact of the same body, so the result is the sequential sweep's,
// bit for bit - but the critical path costs one L2 hand-off per contact instead of a barrier round
// (~9 us).  Hand-off protocol: the mutable state of a body lives in a version-stamped 32-byte record
// (BodyRec above); the releasing side stores the records and decrements the successors' dependency
// counters WITHOUT a fence in between, the side that sees a counter reach zero (or its queue slot
// fill) loads the records from L2 and polls until both carry the versions the contact expects.
// A spin budget (C_ABORT) guarantees termination even if the graph were broken.
constexpr uint32_t FLOW_SPIN_LIMIT = 1u << 21;

// LANES = active lanes per warp.  1: a worker is a single lane (the other 31 exit at once), its hop
// costs its own latency and nothing else - fastest on deep dependency graphs, where the critical
// path is what is being timed (1M packed spheres: 1.8 ms against 3.0 ms for 32 lanes; with the
// fence-based hand-off of the first version 2.4 against 3.5 ms, and 2, 4, 8 lanes per warp on the
// same grid 2.9, 4.7, 6.3 ms).  32: the whole warp in lock step, 5x the workers per SM for wide
// shallow graphs.
template <int LANES>
__global__ void __launch_bounds__(128) response_flow_kernel(FlowArgs a, const uint4* __restrict__ crec, uint32_t* dep,
                                                            uint32_t* q, uint32_t c_n, uint32_t* counters, uint32_t idle_ns) {
    if ((threadIdx.x & 31) >= LANES) return;
    constexpr uint32_t LMASK = LANES == 32 ? 0xFFFFFFFFu : ((1u << LANES) - 1u);
    volatile uint32_t* vq = q;
    volatile uint32_t* vc = counters;
    uint32_t cur = EMPTY, slot = EMPTY, spins = 0, mine = 0;
    uint4 rec = make_uint4(0, 0, EMPTY, EMPTY);    // {i, j, succ_i, succ_j} of `cur`, loaded one iteration early
    bool alive = true;
    uint32_t iters = 0;
#ifdef TR_FLOW_STATS   // instrumented build (make EXTRA=-DTR_FLOW_STATS TAG=stats OUT=...): cycles of the busy iterations
    unsigned long long st_busy = 0, st_nbusy = 0, st_push = 0;
#endif
    // One iteration, two memory phases: (1) a lane with work resolves one contact on the
    // version-checked body records, (2) it releases the contact's successors while idle lanes try to
    // acquire from the queue.  With LANES = 32 the warp runs these in lock step (lanes never wait on
    // each other's hand-offs inside an iteration, only on the records' versions).  Software
    // pipelining: the records of both successors are loaded while the current contact is being
    // resolved, and the bodies they touch are prefetched to L2.  The dependent chain of a hand-off
    // is body-record loads -> dependency atomics: no fence, the records validate themselves.
    while (true) {
        if (LANES == 1 ? !alive : !__any_sync(LMASK, alive)) break;   // the warp leaves together (and reconverges here)
        ++iters;
        // ---- phase 1 (loads): successors' records + this contact's bodies, all in flight together
        const bool work = cur != EMPTY;
#ifdef TR_FLOW_STATS
        const long long st_t0 = clock64();
#endif
        const uint32_t s1 = work ? rec.z : EMPTY, s2 = work ? rec.w : EMPTY;
        uint4 n1 = make_uint4(0, 0, EMPTY, EMPTY), n2 = make_uint4(0, 0, EMPTY, EMPTY);
        if (s1 != EMPTY) n1 = crec[s1];
        if (s2 != EMPTY) n2 = crec[s2];
        // Single-lane workers release the successors FIRST, together with the loads: the dependency
        // counters only schedule - what a successor reads is validated by the record versions, so one
        // that starts early just polls until this contact's records land - and the 530-cycle round trip
        // of the atomics overlaps this contact's loads and arithmetic instead of following them.
        // (Lock-step warps release after the stores: a lane polling for a record that another lane of
        // its own warp still has to write would stall that very lane.)
        // No deadlock: a contact is handed to a worker only when both its predecessors have STARTED, and
        // a started contact stays with its worker until it is finished; so the earliest unfinished
        // contact among those in workers' hands has finished predecessors, i.e. valid records, and
        // proceeds.  A worker can poll for at most the time the <= ~3K contacts started ahead of the
        // data take (milliseconds), far below VERSION_SPIN_LIMIT (~0.2 s of polling).
        // Both decrements are in flight together: written as two ?: expressions the compiler wraps each
        // atomic in its own branch and tests the first result inside it, i.e. the second atomic is
        // not issued until the first has returned.
        uint32_t r1 = 0, r2 = 0;
#define TR_RELEASE_SUCCESSORS()                                                                                          \
        asm volatile(                                                                                                    \
            "{\n\t"                                                                                                      \
            ".reg .pred p1, p2;\n\t"                                                                                     \
            "setp.ne.u32 p1, %4, 0xffffffff;\n\t"                                                                        \
            "setp.ne.u32 p2, %5, 0xffffffff;\n\t"                                                                        \
            "mov.u32 %0, 0;\n\t"                                                                                         \
            "mov.u32 %1, 0;\n\t"                                                                                         \
            "@p1 atom.global.add.u32 %0, [%2], 0xffffffff;\n\t"                                                          \
            "@p2 atom.global.add.u32 %1, [%3], 0xffffffff;\n\t"                                                          \
            "}"                                                                                                          \
            : "=&r"(r1), "=&r"(r2)                                                                                       \
            : "l"(dep + (s1 != EMPTY ? s1 : 0u)), "l"(dep + (s2 != EMPTY ? s2 : 0u)), "r"(s1), "r"(s2)                   \
            : "memory")
        if (LANES == 1 && work) TR_RELEASE_SUCCESSORS();
        if (work && !respond_rec(a, cur, rec.x, rec.y, s1 == EMPTY, s2 == EMPTY, counters)) alive = false;   // watchdog
        if (LANES == 1 && !work) __nanosleep(idle_ns);   // an idle worker that polls flat out slows the busy ones down
        // ---- phase 2: working lanes pick their next contact, idle lanes try to acquire one
        if (work) {
            if (LANES != 1) TR_RELEASE_SUCCESSORS();
#undef TR_RELEASE_SUCCESSORS
            if (s1 != EMPTY) {                    // the next hand-off will touch these records (neutral while they are L2
                                                  // resident - 1.26 ms with or without - kept for worlds that are not)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n1.x));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n1.y));
            }
            if (s2 != EMPTY) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n2.x));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n2.y));
            }
            ++mine;
            uint32_t next = EMPTY;
            // both ready: keep the successor along body j (the column side), queue the one along body i -
            // 5 % faster on the packed scene than the other way round (1.27 against 1.33 ms)
            if (r2 == 1u) {
                next = s2;
                rec = n2;
            }
            if (r1 == 1u) {
                if (next == EMPTY) {
                    next = s1;
                    rec = n1;
                } else {
                    vq[atomicAdd(&counters[C_QTAIL], 1u)] = s1;   // second ready successor: to the global queue
                }
            }
            cur = next;
#ifdef TR_FLOW_STATS
            st_busy += (unsigned long long)(clock64() - st_t0);
            ++st_nbusy;
            st_push += (r1 == 1u && r2 == 1u) ? 1u : 0u;
#endif
        } else if (alive) {
            if (mine) {                           // publish finished work only when a chain has ended:
                // one hot-address atomic per chain, not per contact; whoever completes the count raises
                // the FIN flag on a line nobody updates, which is what the idle workers poll
                if (atomicAdd(&counters[C_DONE], mine) + mine >= c_n) vc[C_FIN] = 1u;
                mine = 0;
            }
            if (slot == EMPTY) {
                const uint32_t h = atomicAdd(&counters[C_QHEAD], 1u);
                if (h >= c_n) alive = false;      // every slot that can ever fill is already claimed
                else slot = h;
            }
            if (alive) {
                const uint32_t c = vq[slot];
                if (c != EMPTY) {
                    cur = c;
                    rec = crec[c];
                    slot = EMPTY;
                    spins = 0;
                } else if ((++spins & 15u) == 0 && (vc[C_FIN] || vc[C_ABORT])) {
                    alive = false;                // termination is polled every 16th idle iteration
                } else if (spins > FLOW_SPIN_LIMIT) {
                    atomicExch(&counters[C_ABORT], 1u);
                    alive = false;
                }
            }
        }
    }
#ifdef TR_FLOW_STATS
    {
        unsigned long long* st = reinterpret_cast<unsigned long long*>(counters + C_NCOUNTERS);
        atomicAdd(st + 0, st_busy);
        atomicAdd(st + 1, st_nbusy);
        atomicAdd(st + 2, st_push);
        atomicAdd(st + 3, (unsigned long long)iters);
        if (atomicAdd(st + 4, 1ull) + 1ull == (unsigned long long)gridDim.x * (blockDim.x / 32) * LANES) {
            printf("[TR_FLOW_STATS] workers %llu busy iterations %llu avg cycles per busy iteration %.0f queue pushes %llu all iterations %llu\n",
                   st[4], st[1], (double)st[0] / (double)st[1], st[2], st[3]);
            for (int k = 0; k < 5; ++k) st[k] = 0;
        }
    }
#endif
    if ((threadIdx.x & 31) == 0) atomicMax(&counters[C_ROUNDS], iters);   // observability: longest warp
}

}  // namespace tr

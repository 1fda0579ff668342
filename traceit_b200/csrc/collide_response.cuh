// K6: contact ordering (prep), the contact itself, and the executors that replay the reference's
// Gauss-Seidel sweep.  Device code only; included by collide.cu after collide_common.cuh.
#pragma once
#include "collide_common.cuh"

namespace tr {

// ---------------------------------------------------------------------------------------
// K6 preparation: dependency structure of the lexicographic Gauss-Seidel order
// ---------------------------------------------------------------------------------------
// The reference visits the contacts in lexicographic (i,j) order, i < j (main.cpp:674-675).  The
// contacts of ONE body b in that order are its column entries (k,b), k < b, by ascending k, followed
// by its row entries (b,j), j > b, by ascending j - i.e. simply ALL its contacts by ascending PARTNER
// id.  So the dependency structure is one sort of the 2C half edges (body, partner, contact) by
// (body, partner): contact c may run when the predecessor half edge of both its bodies has run.
// Nothing needs the contact list itself in sorted order (tr_read_contacts sorts on the host).
//
// prep_kernel does that sort as ONE persistent kernel with a grid barrier between its phases, sized by
// the contact count on the device (no host round trip between detection and response):
//   A  count    r = atomicAdd(body_cnt[b], 1) for both bodies of every contact: degree + arrival rank
//   B  alloc    the half edge that arrived first at a body reserves the body's segment (warp-aggregated
//               bump allocation); bodies with more than PREP_SHORT contacts go on the long list
//   C  scatter  half edge {partner, contact} -> segment[arrival rank]; body_cnt counted back to zero
//   D  sort     long segments only, bitonic by partner id: one CTA per segment in shared memory up to 2048
//               entries; beyond that (a hub: the slab 10^5 bodies rest on) all CTAs sort the segment together,
//               chunks in shared memory and the long-distance steps in global memory with a grid barrier per
//               step - a body with L contacts costs O(L log^2 L) spread over the grid, not the O(L^2) on one
//               thread of ranking every entry against every other
//   E  deps     per contact: rank among / successor in the lists of both bodies (linear scan of a short
//               unsorted segment, binary search in a sorted long one) -> dep[c], {i, j, succ_i, succ_j},
//               expected record versions, contact constants, version-0 body records, ready queue.
// An isStatic body is never changed by a contact (main.cpp:725-741), so its contacts do not order each
// other: no dependency edge, no version bump, no write-back through a static body - everything resting
// on one static slab runs concurrently instead of one after the other.
constexpr uint32_t PREP_SHORT = 24;        // segments up to this length are ranked by linear scans
constexpr int PREP_THREADS = 256;

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

__device__ __forceinline__ void contact_normal_values(uint8_t fi, uint8_t fj, const float4 ci, const float4 cj, const float4 hi, const float4 hj,
                                                      float& nx, float& ny, float& nz) {
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

__device__ __forceinline__ void contact_normal(uint8_t fi, uint8_t fj, const float4 ci, const float4 cj, const float4* half_ext,
                                               uint32_t i, uint32_t j, float& nx, float& ny, float& nz) {
    float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), hj = hi;
    if (!((fi & F_SPHERE) && (fj & F_SPHERE))) {
        hi = half_ext[i];
        hj = half_ext[j];
    }
    contact_normal_values(fi, fj, ci, cj, hi, hj, nx, ny, nz);
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

// Every access to a record that can race with another SM's access is a RELAXED, GPU-SCOPE, 128-BIT access
// (PTX .b128: one scalar access in the memory model, single-copy atomic - a 256-bit vector access would be
// "a set of 32-bit accesses in unspecified order" there).  Each 16-byte half carries its own version, so the two
// halves are validated independently and a reader can never act on a torn record; SASS: LDG/STG.E.128.STRONG.GPU.
// (-DTR_BODY_REC_WEAK restores round 1's single weak 256-bit access for A/B timing.)
__device__ __forceinline__ void st128_relaxed(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("{\n\t.reg .b128 t;\n\t.reg .b64 lo, hi;\n\tmov.b64 lo, {%1,%2};\n\tmov.b64 hi, {%3,%4};\n\tmov.b128 t, {lo, hi};\n\t"
                 "st.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void ld128_relaxed(const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("{\n\t.reg .b128 t;\n\t.reg .b64 lo, hi;\n\tld.relaxed.gpu.global.b128 t, [%4];\n\tmov.b128 {lo, hi}, t;\n\t"
                 "mov.b64 {%0,%1}, lo;\n\tmov.b64 {%2,%3}, hi;\n\t}"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "l"(p)
                 : "memory");
}

__device__ __forceinline__ void st_body(BodyRec* p, const float4 v, const float4 x, uint32_t ver) {
#ifdef TR_BODY_REC_WEAK
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)),
                 "r"(__float_as_uint(v.z)), "r"(ver), "r"(__float_as_uint(x.x)), "r"(__float_as_uint(x.y)),
                 "r"(__float_as_uint(x.z)), "r"(ver)
                 : "memory");
#else
    st128_relaxed(&p->vx, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), ver);
    st128_relaxed(&p->px, __float_as_uint(x.x), __float_as_uint(x.y), __float_as_uint(x.z), ver);
#endif
}
// L2 load (the record is written by other SMs); true when both halves carry the expected version
__device__ __forceinline__ bool ld_body(const BodyRec* p, uint32_t want, float4& v, float4& x) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
#ifdef TR_BODY_REC_WEAK
    asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "l"(p)
                 : "memory");
#else
    ld128_relaxed(&p->vx, r0, r1, r2, r3);
    ld128_relaxed(&p->px, r4, r5, r6, r7);
#endif
    v.x = __uint_as_float(r0); v.y = __uint_as_float(r1); v.z = __uint_as_float(r2);
    x.x = __uint_as_float(r4); x.y = __uint_as_float(r5); x.z = __uint_as_float(r6);
    return r3 == want && r7 == want;
}

struct PrepArgs {
    const unsigned long long* ckeys;
    unsigned long long cap;
    uint32_t* counters;
    uint32_t* ctl;
    uint32_t* body_cnt;
    uint2* seg;
    uint2* arank;
    uint2* he;
    uint32_t* long_list;
    uint32_t* huge_list;
    uint32_t* dep;
    uint4* crec;
    uint2* cver;
    ContactGeo* cgeo;
    BodyRec* brec;          // NULL for the level-synchronous executor
    uint32_t* queue;
    int ready_counter;      // C_QTAIL (dataflow) or C_FRONT (level-synchronous)
    const float4* pos_m;
    const float4* vel_rest;
    const float4* center_r;
    const float4* half_ext;
    const uint8_t* flags;
    uint32_t n;             // bodies (index checks of the checked build)
    uint32_t after_sparse;  // sparse_step_kernel ran before this kernel in the step's graph (it stamped the phase, and may have done the step)
};

// all CTAs of prep_kernel are resident (the host sizes the grid from the occupancy API): a counter barrier
__device__ __forceinline__ void prep_barrier(uint32_t* bar, uint32_t& phase, uint32_t nblocks) {
    __syncthreads();
    ++phase;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const uint32_t want = phase * nblocks;
        while (*((volatile uint32_t*)bar) < want) __nanosleep(40);
        __threadfence();
    }
    __syncthreads();
}

// position of `partner` among the half edges of one body: rank (number of smaller partners) and the
// contact of the next larger partner (EMPTY if none)
__device__ __forceinline__ void segment_lookup(const uint2* __restrict__ he, const uint2 sg, uint32_t partner, uint32_t& rank,
                                               uint32_t& succ) {
    const uint2* p = he + sg.x;
    if (sg.y <= PREP_SHORT) {
        uint32_t r = 0, best = EMPTY, bestc = EMPTY;
        for (uint32_t k = 0; k < sg.y; ++k) {
            const uint2 e = __ldcg(p + k);   // written by other SMs earlier in this kernel: L2, not L1
            r += e.x < partner ? 1u : 0u;
            if (e.x > partner && e.x < best) {   // partner ids are < 2^30, so EMPTY works as +inf
                best = e.x;
                bestc = e.y;
            }
        }
        rank = r;
        succ = bestc;
    } else {
        uint32_t lo = 0, hi = sg.y;              // sorted by phase D: first entry with partner id >= partner
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldcg(p + mid).x < partner) lo = mid + 1;
            else hi = mid;
        }
        rank = lo;
        succ = lo + 1 < sg.y ? __ldcg(p + lo + 1).y : EMPTY;
    }
}

__global__ void __launch_bounds__(PREP_THREADS) prep_kernel(PrepArgs a) {
    __shared__ uint2 s_sort[PREP_SMEM_SORT];
    if (!a.after_sparse && blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<unsigned long long*>(a.counters + C_T0 + 4) = globaltimer_ns();
    if (a.after_sparse && a.counters[C_SPARSE] != 0) return;   // the sparse-step kernel has replayed this step
    if (a.ctl[CTL_ABORT] != 0) return;                     // an earlier step overflowed: the host has to re-run from there
    const unsigned long long found = *reinterpret_cast<const unsigned long long*>(a.counters + C_CONTACTS);
    if (found > a.cap) {
        // more contacts than the buffers hold: no response at all, and no later step may touch the state
        if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl[CTL_ABORT] = a.ctl[CTL_SEQ] + 1u;
        return;
    }
    const uint32_t cn = (uint32_t)found;
    if (cn == 0) return;
    // a sparse world keeps only as many CTAs as it has work for: the grid barrier costs what its participants cost
    const uint32_t nblocks = min(gridDim.x, (cn + PREP_THREADS - 1) / PREP_THREADS);
    if (blockIdx.x >= nblocks) return;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = nblocks * blockDim.x;
    uint32_t phase = 0;
    uint32_t* bar = a.counters + C_BAR;

    // ---- A: degrees + arrival ranks (and an EMPTY ready queue)
    for (uint32_t c = gtid; c < cn; c += gsize) {
        const unsigned long long k = a.ckeys[c];
        const uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
        TR_ASSERT(i < j && j < a.n);
        a.arank[c] = make_uint2(atomicAdd(&a.body_cnt[i], 1u), atomicAdd(&a.body_cnt[j], 1u));
        a.queue[c] = EMPTY;
    }
    prep_barrier(bar, phase, nblocks);

    // ---- B: segment allocation by the first arrival at each body
    for (uint32_t c0 = blockIdx.x * blockDim.x; c0 < cn; c0 += gsize) {   // whole warps stay in the loop (shuffles below)
        const uint32_t c = c0 + threadIdx.x;
        uint32_t i = 0, j = 0, li = 0, lj = 0;
        if (c < cn) {
            const unsigned long long k = a.ckeys[c];
            const uint2 r = a.arank[c];
            i = (uint32_t)(k >> 32);
            j = (uint32_t)k;
            if (r.x == 0) li = __ldcg(&a.body_cnt[i]);
            if (r.y == 0) lj = __ldcg(&a.body_cnt[j]);
        }
        const uint32_t need = li + lj, lane = threadIdx.x & 31;
        uint32_t incl = need;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t base = 0;
        if (lane == 31 && total) base = atomicAdd(&a.counters[C_HE], total);
        base = __shfl_sync(0xFFFFFFFFu, base, 31) + (incl - need);
        TR_ASSERT((unsigned long long)base + li + lj <= 2ull * a.cap);
        if (li) {
            a.seg[i] = make_uint2(base, li);
            if (li > PREP_SMEM_SORT) a.huge_list[atomicAdd(&a.counters[C_HUGE], 1u)] = i;
            else if (li > PREP_SHORT) a.long_list[atomicAdd(&a.counters[C_LONG], 1u)] = i;
            base += li;
        }
        if (lj) {
            a.seg[j] = make_uint2(base, lj);
            if (lj > PREP_SMEM_SORT) a.huge_list[atomicAdd(&a.counters[C_HUGE], 1u)] = j;
            else if (lj > PREP_SHORT) a.long_list[atomicAdd(&a.counters[C_LONG], 1u)] = j;
        }
    }
    prep_barrier(bar, phase, nblocks);

    // ---- C: half edges into their segments; the histogram counts itself back to zero
    for (uint32_t c = gtid; c < cn; c += gsize) {
        const unsigned long long k = a.ckeys[c];
        const uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
        const uint2 r = a.arank[c];
        TR_ASSERT(r.x < __ldcg(&a.seg[i]).y && r.y < __ldcg(&a.seg[j]).y);
        a.he[__ldcg(&a.seg[i]).x + r.x] = make_uint2(j, c);
        a.he[__ldcg(&a.seg[j]).x + r.y] = make_uint2(i, c);
        if (r.x == 0) a.body_cnt[i] = 0;
        if (r.y == 0) a.body_cnt[j] = 0;
    }
    prep_barrier(bar, phase, nblocks);

    // ---- D: long segments sorted by partner id.  Up to PREP_SMEM_SORT entries: one CTA each, in shared memory.
    const uint32_t nlong = *((volatile uint32_t*)&a.counters[C_LONG]);
    for (uint32_t l = blockIdx.x; l < nlong; l += nblocks) {
        const uint2 sg = __ldcg(&a.seg[__ldcg(&a.long_list[l])]);
        if (sg.y > PREP_SMEM_SORT) continue;
        uint2* p = a.he + sg.x;
        for (uint32_t k = threadIdx.x; k < sg.y; k += blockDim.x) s_sort[k] = __ldcg(p + k);
        __syncthreads();
        smem_sort(s_sort, sg.y, false);
        for (uint32_t k = threadIdx.x; k < sg.y; k += blockDim.x) p[k] = s_sort[k];
        __syncthreads();
    }
    // Hubs (a slab that 10^5 bodies rest on): ALL CTAs sort one such segment together.  Chunks of PREP_SMEM_SORT
    // entries are sorted in shared memory; the stages above the chunk size do their long-distance steps in global
    // memory - one compare-exchange per thread of the grid and a grid barrier per step - and finish inside the chunks
    // again.  27 barriers for 10^5 entries; a single CTA walking the same network alone took 12 ms.
    const uint32_t nhuge = *((volatile uint32_t*)&a.counters[C_HUGE]);
    for (uint32_t hgi = 0; hgi < nhuge; ++hgi) {
        const uint2 sg = __ldcg(&a.seg[__ldcg(&a.huge_list[hgi])]);
        uint2* p = a.he + sg.x;
        const uint32_t len = sg.y, nchunks = (len + PREP_SMEM_SORT - 1) / PREP_SMEM_SORT;
        uint32_t pow2 = 1;
        while (pow2 < len) pow2 <<= 1;
        for (int pass = 0;; ++pass) {
            // pass 0 sorts every chunk; pass m > 0 is stage k = PREP_SMEM_SORT << m
            const uint32_t k = PREP_SMEM_SORT << pass;
            if (pass > 0) {
                if (k > pow2) break;
                uint32_t i, l;
                for (uint32_t t = gtid; t < pow2 / 2; t += gsize) {
                    flip_pair(t, k, i, l);
                    global_cmpx(p, i, l, len);
                }
                prep_barrier(bar, phase, nblocks);
                for (uint32_t j = k / 4; j >= PREP_SMEM_SORT; j >>= 1) {
                    for (uint32_t t = gtid; t < pow2 / 2; t += gsize) {
                        shear_pair(t, j, i, l);
                        global_cmpx(p, i, l, len);
                    }
                    prep_barrier(bar, phase, nblocks);
                }
            }
            for (uint32_t ch = blockIdx.x; ch < nchunks; ch += nblocks) {
                uint2* q = p + ch * PREP_SMEM_SORT;
                const uint32_t clen = min(PREP_SMEM_SORT, len - ch * PREP_SMEM_SORT);
                for (uint32_t x = threadIdx.x; x < clen; x += blockDim.x) s_sort[x] = __ldcg(q + x);
                __syncthreads();
                smem_sort(s_sort, clen, pass > 0);
                for (uint32_t x = threadIdx.x; x < clen; x += blockDim.x) __stcg(q + x, s_sort[x]);
                __syncthreads();
            }
            prep_barrier(bar, phase, nblocks);
        }
    }
    if (nlong) prep_barrier(bar, phase, nblocks);  // nlong is the same for every CTA

    // ---- E: dependencies, versions, constants, ready queue
    for (uint32_t c = gtid; c < cn; c += gsize) {
        const unsigned long long k = a.ckeys[c];
        const uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
        uint32_t ri, rj, si, sj;
        segment_lookup(a.he, __ldcg(&a.seg[i]), j, ri, si);
        segment_lookup(a.he, __ldcg(&a.seg[j]), i, rj, sj);
        TR_ASSERT(ri < __ldcg(&a.seg[i]).y && rj < __ldcg(&a.seg[j]).y && (si == EMPTY || si < cn) && (sj == EMPTY || sj < cn));
        const uint8_t fi = a.flags[i], fj = a.flags[j];
        const float4 vi = a.vel_rest[i], vj = a.vel_rest[j], pi = a.pos_m[i], pj = a.pos_m[j];
        if (a.brec) {                                  // version-0 records, written once per body
            if (ri == 0) st_body(a.brec + i, vi, pi, 0u);
            if (rj == 0) st_body(a.brec + j, vj, pj, 0u);
        }
        // contacts applied to each body before this one; a static body is never changed, so nothing orders its contacts
        const uint32_t ei = (fi & F_STATIC) ? 0u : ri, ej = (fj & F_STATIC) ? 0u : rj;
        if (fi & F_STATIC) si = EMPTY;
        if (fj & F_STATIC) sj = EMPTY;
        const uint32_t d = (ei ? 1u : 0u) + (ej ? 1u : 0u);
        a.dep[c] = d;
        a.crec[c] = make_uint4(i, j, si, sj);
        a.cver[c] = make_uint2(ei, ej);
        ContactGeo g;
        contact_normal(fi, fj, a.center_r[i], a.center_r[j], a.half_ext, i, j, g.nx, g.ny, g.nz);
        g.ima = __fdiv_rn(1.0f, pi.w);
        g.imb = __fdiv_rn(1.0f, pj.w);
        g.den = __fadd_rn(g.ima, g.imb);
        g.e = (vj.w < vi.w) ? vj.w : vi.w;               // std::min(a.rest, b.rest)
        g.bits = ((fi & F_STATIC) ? 1u : 0u) | ((fj & F_STATIC) ? 2u : 0u);
        a.cgeo[c] = g;
        if (d == 0) {
            const uint32_t qs = warp_reserve32(&a.counters[a.ready_counter]);
            TR_ASSERT(qs < cn);
            a.queue[qs] = c;
        }
    }
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

// One contact on VALUES: normal from the collider centres (main.cpp:701-706), approach test exactly as written
// (main.cpp:714), impulse and positional nudge (main.cpp:716-741).  v*.w = restitution, p*.w = mass.
__device__ __forceinline__ void respond_values(uint8_t fi, uint8_t fj, const float4 ci, const float4 cj, const float4 hi, const float4 hj,
                                               float4& vi, float4& vj, float4& pi, float4& pj) {
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
        float dx = __fsub_rn(cj.x, ci.x), dy = __fsub_rn(cj.y, ci.y), dz = __fsub_rn(cj.z, ci.z);
        float ox = __fsub_rn(__fadd_rn(hi.x, hj.x), fabsf(dx));
        float oy = __fsub_rn(__fadd_rn(hi.y, hj.y), fabsf(dy));
        float oz = __fsub_rn(__fadd_rn(hi.z, hj.z), fabsf(dz));
        nx = ny = nz = 0.0f;
        if (ox < oy && ox < oz) nx = dx > 0 ? 1.0f : -1.0f;
        else if (oy < oz) ny = dy > 0 ? 1.0f : -1.0f;
        else nz = dz > 0 ? 1.0f : -1.0f;
    }
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
    }
    if (!(fj & F_STATIC)) {
        vj.x = __fadd_rn(vj.x, __fmul_rn(ix, imb));
        vj.y = __fadd_rn(vj.y, __fmul_rn(iy, imb));
        vj.z = __fadd_rn(vj.z, __fmul_rn(iz, imb));
        pj.x = __fadd_rn(pj.x, __fmul_rn(cx, imb));
        pj.y = __fadd_rn(pj.y, __fmul_rn(cy, imb));
        pj.z = __fadd_rn(pj.z, __fmul_rn(cz, imb));
    }
}

// the same contact on the SoA state (level-synchronous executor)
__device__ __forceinline__ void respond(const RespArgs& a, uint32_t i, uint32_t j) {
    const uint8_t fi = a.flags[i], fj = a.flags[j];
    // L2 loads: these were written by other SMs in earlier wavefronts.  Positions are fetched in
    // the same round trip (needed only when the pair is not skipped, but a second dependent trip
    // would sit on the critical path of the ordered replay).
    float4 vi = __ldcg(a.vel_rest + i), vj = __ldcg(a.vel_rest + j);
    float4 pi = __ldcg(a.pos_m + i), pj = __ldcg(a.pos_m + j);
    float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), hj = hi;
    if (!((fi & F_SPHERE) && (fj & F_SPHERE))) {
        hi = a.half_ext[i];
        hj = a.half_ext[j];
    }
    respond_values(fi, fj, a.center_r[i], a.center_r[j], hi, hj, vi, vj, pi, pj);
    if (!(fi & F_STATIC)) {
        a.vel_rest[i] = vi;
        a.pos_m[i] = pi;
    }
    if (!(fj & F_STATIC)) {
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
    // (a static body is never changed: its record stays at version 0 for all its contacts, nothing to store)
    if (g.bits & 1u) {
    } else if (last_i) {
        store_xyz(a.vel_out + i, vi);
        store_xyz(a.pos_out + i, pi);
    } else {
        st_body(a.brec + i, vi, pi, want.x + 1u);
    }
    if (g.bits & 2u) {
    } else if (last_j) {
        store_xyz(a.vel_out + j, vj);
        store_xyz(a.pos_out + j, pj);
    } else {
        st_body(a.brec + j, vj, pj, want.y + 1u);
    }
    return true;
}

// Persistent cooperative kernel: one wavefront per round, grid-wide barrier between rounds.
// counters[8 + r%3] is the size of round r's frontier.
__global__ void __launch_bounds__(256) response_kernel(RespArgs a, const uint4* __restrict__ crec, uint32_t* dep, uint32_t* f0,
                                                       uint32_t* f1, uint32_t* counters, const uint32_t* ctl) {
    cg::grid_group grid = cg::this_grid();
    if (ctl[CTL_ABORT] != 0) return;   // uniform: set by an earlier kernel
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
            const uint4 k = crec[c];
            respond(a, k.x, k.y);
            __threadfence();
            if (k.z != EMPTY && atomicSub(&dep[k.z], 1u) == 1u) fout[atomicAdd(cnt_out, 1u)] = k.z;
            if (k.w != EMPTY && atomicSub(&dep[k.w], 1u) == 1u) fout[atomicAdd(cnt_out, 1u)] = k.w;
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
// A spin budget (C_ABORT) guarantees termination even if the graph were broken: an idle worker gives up after
// FLOW_SPIN_LIMIT polls (~1 s) during which the finished count did not move (busy workers publish it every 256 contacts).
constexpr uint32_t FLOW_SPIN_LIMIT = 1u << 21;

// LANES = active lanes per warp.  1: a worker is a single lane (the other 31 exit at once), its hop
// costs its own latency and nothing else - fastest on deep dependency graphs, where the critical
// path is what is being timed (1M packed spheres: 1.8 ms against 3.0 ms for 32 lanes; with the
// fence-based hand-off of the first version 2.4 against 3.5 ms, and 2, 4, 8 lanes per warp on the
// same grid 2.9, 4.7, 6.3 ms).  32: the whole warp in lock step, 5x the workers per SM for wide
// shallow graphs.
template <int LANES>
__global__ void __launch_bounds__(128) response_flow_kernel(FlowArgs a, const uint4* __restrict__ crec, uint32_t* dep,
                                                            uint32_t* q, const uint32_t* ctl, uint32_t* counters, uint32_t idle_ns) {
    if ((threadIdx.x & 31) >= LANES) return;
    if (ctl[CTL_ABORT] != 0) return;               // contact buffer overflow in this or an earlier step: no response
    if (counters[C_SPARSE] != 0) return;           // replayed by sparse_step_kernel earlier in the step's graph
    const uint32_t c_n = counters[C_CONTACTS];     // <= capacity < 2^32 here (prep_kernel checked)
    if ((unsigned long long)blockIdx.x * (blockDim.x / 32) * LANES >= c_n) return;   // no more workers than contacts (also: none for none)
    constexpr uint32_t LMASK = LANES == 32 ? 0xFFFFFFFFu : ((1u << LANES) - 1u);
    volatile uint32_t* vq = q;
    volatile uint32_t* vc = counters;
    uint32_t cur = EMPTY, slot = EMPTY, spins = 0, mine = 0, seen_done = 0;
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
            if (++mine >= 256u) {                 // progress for the watchdog of the idle workers (and the finished count)
                if (atomicAdd(&counters[C_DONE], mine) + mine >= c_n) vc[C_FIN] = 1u;
                mine = 0;
            }
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
                    const uint32_t qs = atomicAdd(&counters[C_QTAIL], 1u);   // second ready successor: to the global queue
                    TR_ASSERT(qs < c_n && s1 < c_n);
                    vq[qs] = s1;
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
                TR_ASSERT(slot < c_n && (c == EMPTY || c < c_n));
                if (c != EMPTY) {
                    cur = c;
                    rec = crec[c];
                    slot = EMPTY;
                    spins = 0;
                } else if ((++spins & 15u) == 0 && (vc[C_FIN] || vc[C_ABORT])) {
                    alive = false;                // termination is polled every 16th idle iteration
                } else if ((spins & 4095u) == 0 && vc[C_DONE] != seen_done) {
                    seen_done = vc[C_DONE];       // the watchdog measures time WITHOUT progress, not time: a deep but
                    spins = 0;                    // healthy chain (a line of a million touching spheres) must not trip it
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

// The step's record for the host (mapped pinned memory - the host reads it after its next synchronisation, nothing
// waits for it).  One thread.  extra_flags: 8 = a sparse step that has to be re-run through the general kernels,
// 16 = replayed by sparse_step_kernel.
__device__ __forceinline__ void write_step_record(uint32_t* counters, uint32_t* ctl, StepRecord* ring, uint32_t extra_flags, uint32_t rounds) {
    const uint32_t seq = ctl[CTL_SEQ];
    StepRecord r;
    r.flags = extra_flags;
    const unsigned long long found = *reinterpret_cast<const unsigned long long*>(counters + C_CONTACTS);
    if (ctl[CTL_ABORT] == seq + 1u) r.flags |= 1u;            // this step overflowed the contact buffers (or, with 8: was not sparse after all)
    else if (ctl[CTL_ABORT] != 0) r.flags |= 4u;              // a step before it did: this one did nothing
    if (counters[C_ABORT]) r.flags |= 2u;                     // dataflow watchdog
    r.contacts_lo = (uint32_t)found;
    r.contacts_hi = (uint32_t)(found >> 32);
    r.side = counters[C_SIDE];
    r.far = counters[C_FAR];
    r.bnd = counters[C_BND];
    r.grid = counters[C_GRID];
    r.rounds = rounds;
    r.long_segments = counters[C_LONG] + counters[C_HUGE];
    r.pad[0] = r.pad[1] = 0;
    r.t[0] = *reinterpret_cast<const unsigned long long*>(counters + C_T0);
    r.t[1] = *reinterpret_cast<const unsigned long long*>(counters + C_T0 + 2);
    r.t[2] = *reinterpret_cast<const unsigned long long*>(counters + C_T0 + 4);
    r.t[3] = globaltimer_ns();
    r.seq = seq + 1u;
    ring[seq % STEP_RING] = r;
    if (r.flags & 1u) ring[STEP_RING] = r;   // the overflow notice has a slot of its own: it must survive any number of later (no-op) steps
    ctl[CTL_SEQ] = seq + 1u;
}

// Last kernel of a general collision step: the record, then the counters are zeroed for the next step.
__global__ void __launch_bounds__(128) epilogue_kernel(uint32_t* counters, uint32_t* ctl, StepRecord* ring) {
    if (threadIdx.x == 0) {
        const uint32_t sparse = counters[C_SPARSE];
        write_step_record(counters, ctl, ring, sparse ? 16u : 0u, sparse ? sparse - 1u : counters[C_ROUNDS]);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < (uint32_t)C_NCOUNTERS; k += blockDim.x) counters[k] = 0;
}

// ---------------------------------------------------------------------------------------
// Sparse steps: ordering + replay + record in ONE kernel of ONE thread-block cluster
// ---------------------------------------------------------------------------------------
// A step with a few thousand contacts (the 1M random cloud: ~2,500; most frames of a real scene) has
// nothing to spread over a grid: prep_kernel's phases, the dataflow executor and the epilogue are then
// three launches of pure latency - grid barriers and L2 round trips, 33 us for 2,504 contacts.  When the
// host expects such a step (the newest record it has seen says so), or does not know (then the general
// kernels follow in the graph and return at once if this one has done the step), the narrow phase is followed
// by this kernel: 16 CTAs x 256 threads = one contact per thread, all bookkeeping and the bodies' mutable
// state in DISTRIBUTED SHARED MEMORY (a hand-off costs a few hundred cycles instead of an L2 round trip),
// cluster barriers instead of grid barriers:
//   gather  the contact's bodies (collider centres, velocity, position, half extents, flags): the step's only
//           DRAM-cold reads, issued by the kernel as asynchronous copies into per-thread staging slots so that
//           they run under P0-P1 (a cluster barrier's release is a GPU-scope MEMBAR: it waits for ordinary loads)
//   P0      a body belongs to the CTA its id hashes to; (remote) atomicCAS into that CTA's hash table,
//           (remote) atomicAdd on the slot's counter = degree + arrival rank
//   P1      owners: occupied slots get dense local ids and half-edge segments (one CTA scan); local id, segment
//           and degree go back into the slot word
//   P2      half edges (partner id) into the owners' segments - not for static bodies, which order nothing;
//           a movable body with more than SPARSE_MAX_DEGREE contacts fails the step here; the first arrival
//           deposits the body's velocity + position record
//   P3      rank of the contact among the contacts of both its bodies (= the version it must see)
//   P5      the replay, without barriers: poll both bodies' records, apply when both carry the contact's
//           versions, store them back with the next versions
//   P6      owners write the changed bodies back; CTA 0 writes the step's record and re-arms the counters
//           (or, with the general kernels behind it, only tells them that the step is done).
// Same order, same arithmetic (respond_rec's), same bits as the other executors.  One CTA alone cannot do
// this: shared-memory atomics and scattered stores run at ~1 per cycle per SM, 2,500 contacts x ~20 of them
// took 45 us on one SM; remote shared memory, too, moves ~1 access per cycle per SM - hence 128-bit records
// and 16 CTAs rather than 8.
// What the kernel cannot take - more than 4,096 contacts, a movable body with more than SPARSE_MAX_DEGREE
// contacts (the general path sorts such lists), a lopsided hash partition - it finds out BEFORE it has changed
// anything.  With the general kernels behind it in the graph it simply leaves the step to them; alone, it raises
// the same flag as a contact-buffer overflow (CTL_ABORT: every later step is a no-op) plus bit 8 in its record,
// and the host re-runs the step through the general kernels at its next synchronisation and is less confident
// for a while (collide.cu: predict_variant; api.cu: settle()).
#ifndef TR_SPARSE_CLUSTER
#define TR_SPARSE_CLUSTER 16
#endif
constexpr int SPARSE_CLUSTER = TR_SPARSE_CLUSTER;            // 16 CTAs: beyond the portable cluster size, allowed per kernel (collide.cu)
constexpr int SPARSE_CLUSTER_LOG2 = SPARSE_CLUSTER == 16 ? 4 : 3;
constexpr int SPARSE_THREADS = 4096 / SPARSE_CLUSTER;
constexpr uint32_t SPARSE_MAX_CONTACTS = SPARSE_CLUSTER * SPARSE_THREADS;   // one contact per thread
constexpr uint32_t SPARSE_SLOTS = 4 * SPARSE_THREADS;        // hash slots per CTA (16,384 in the cluster for at most 8,192 bodies)
constexpr uint32_t SPARSE_BODIES = 4 * SPARSE_THREADS;       // bodies one CTA can own (twice the expected number at the limit)
constexpr uint32_t SPARSE_MAX_DEGREE = 64;     // contacts of one movable body (ranked by a linear scan of remote shared memory)
constexpr uint32_t SPARSE_DIRTY = 0x80000000u;
constexpr uint32_t SPARSE_BODIES_LOG2 = SPARSE_CLUSTER == 16 ? 10 : 11;
static_assert((1u << SPARSE_BODIES_LOG2) == SPARSE_BODIES && 2u * SPARSE_MAX_CONTACTS == (1u << 13), "packing of a hash slot after P1");
constexpr uint32_t SPARSE_DEG_SAT = (1u << (32 - SPARSE_BODIES_LOG2 - 13)) - 1u;
static_assert(SPARSE_DEG_SAT > SPARSE_MAX_DEGREE, "a saturated degree must still read as too long");
// hash slot after P1: local id | first half-edge slot | degree (saturating)
__host__ __device__ constexpr uint32_t sparse_pack(uint32_t l, uint32_t off, uint32_t deg) {
    return l | (off << SPARSE_BODIES_LOG2) | ((deg < SPARSE_DEG_SAT ? deg : SPARSE_DEG_SAT) << (SPARSE_BODIES_LOG2 + 13));
}
static_assert(SPARSE_CLUSTER == 8 || SPARSE_CLUSTER == 16, "owner = top bits of the hash");

struct SparseSmem {
    uint32_t key[SPARSE_SLOTS];                // body id per slot
    uint32_t cnt[SPARSE_SLOTS];                // degree; after P1: sparse_pack(local id, first half-edge slot, degree)
    uint32_t gid[SPARSE_BODIES];               // local id -> body id
    uint32_t he[2 * SPARSE_MAX_CONTACTS];      // partner ids, by segment (all half edges fit one CTA: no overflow case)
    // The mutable state, as in the dataflow executor's BodyRec: {vx, vy, vz, version}, {px, py, pz, version}; version = contacts
    // applied so far (bit 31: something changed).  Each half is read and written as ONE 128-bit access, so a reader that finds its
    // version in both halves holds the values stored with it: no fence, no separate flag.  Remote shared memory moves about one
    // access per cycle per SM whatever its width - 4 accesses per body and hand-off instead of 14.
    uint4 rec[SPARSE_BODIES][2];
    uint32_t wsum[2][SPARSE_THREADS / 32];     // P1's CTA scan
    uint4 stage[8][SPARSE_THREADS];            // the thread's gather, landing asynchronously: centre, velocity, position, half extents of i and j
    uint32_t fstage[2][SPARSE_THREADS];        // ... and the aligned words that hold their flag bytes
    uint32_t nb, fail, rounds;                 // fail, rounds: CTA 0's copy is the cluster's
};

struct SparseArgs {
    const unsigned long long* ckeys;
    unsigned long long cap;
    uint32_t* counters;
    uint32_t* ctl;
    StepRecord* ring;
    float4* pos_m;
    float4* vel_rest;
    const float4* center_r;
    const float4* half_ext;
    const uint8_t* flags;
    uint32_t n;
    uint32_t tail_follows;   // prep_kernel, the executor and epilogue_kernel follow in the step's graph: what this kernel cannot take is theirs
};

constexpr uint32_t SPARSE_SPIN_LIMIT = 1u << 20;
// record hand-off between threads of the cluster: relaxed, cluster-scope, 128-bit accesses (generic addresses inside the
// cluster's shared-memory window); PTX .b128 = one access in the memory model
__device__ __forceinline__ void ld128_relaxed_cluster(const uint4* p, uint4& v) {
    asm volatile("{\n\t.reg .b128 t;\n\t.reg .b64 lo, hi;\n\tld.relaxed.cluster.b128 t, [%4];\n\tmov.b128 {lo, hi}, t;\n\t"
                 "mov.b64 {%0,%1}, lo;\n\tmov.b64 {%2,%3}, hi;\n\t}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st128_relaxed_cluster(uint4* p, const uint4 v) {
    asm volatile("{\n\t.reg .b128 t;\n\t.reg .b64 lo, hi;\n\tmov.b64 lo, {%1,%2};\n\tmov.b64 hi, {%3,%4};\n\tmov.b128 t, {lo, hi};\n\t"
                 "st.relaxed.cluster.b128 [%0], t;\n\t}" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// slot of body b in the owner's table (EMPTY: table full)
__device__ __forceinline__ uint32_t sparse_insert(uint32_t* keys, uint32_t b, uint32_t h) {
    for (uint32_t probes = 0; probes < SPARSE_SLOTS; ++probes) {
        const uint32_t old = atomicCAS(&keys[h], EMPTY, b);
        if (old == EMPTY || old == b) return h;
        h = (h + 1u) & (SPARSE_SLOTS - 1u);
    }
    return EMPTY;
}

// false: not a step for this kernel (nothing has been changed).  The result is the same in every thread of the cluster.
__device__ __forceinline__ bool sparse_replay(const SparseArgs& a, SparseSmem& s, cg::cluster_group& cluster, const uint32_t cn,
                                              const unsigned long long key, uint32_t& rounds) {
    const uint32_t rank = cluster.block_rank(), tid = threadIdx.x, lane = tid & 31u;
    const uint32_t c = rank * SPARSE_THREADS + tid;
    const bool have = c < cn;
    SparseSmem* s0 = cluster.map_shared_rank(&s, 0);
#ifdef TR_SPARSE_PROF   // instrumented build (make EXTRA=-DTR_SPARSE_PROF TAG=sprof OUT=...): phase boundaries on one thread's clock
    long long sp_t[16];
    int sp_n = 0;
#define SP_STAMP() do { if (rank == 0 && tid == 0 && sp_n < 16) sp_t[sp_n++] = clock64(); } while (0)
#else
#define SP_STAMP() do { } while (0)
#endif
    SP_STAMP();
    static_assert(SPARSE_SLOTS == 4 * SPARSE_THREADS, "one 16-byte store per thread and table");
    reinterpret_cast<uint4*>(s.key)[tid] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
    reinterpret_cast<uint4*>(s.cnt)[tid] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) s.nb = s.fail = s.rounds = 0;
    // (the bodies of this contact are on their way into s.stage - asynchronous copies issued by the caller; read after P1)
    uint32_t bi = 0, bj = 0;
    if (have) {
        bi = (uint32_t)(key >> 32);
        bj = (uint32_t)key;
        TR_ASSERT(bi < bj && bj < a.n);
    }
    SP_STAMP();
    cluster.sync();                                          // every CTA's tables are initialised
    SP_STAMP();
    // ---- P0: slots, degrees, arrival ranks
    const uint32_t hhi = bi * 2654435761u, hhj = bj * 2654435761u;
    SparseSmem* si = cluster.map_shared_rank(&s, hhi >> (32 - SPARSE_CLUSTER_LOG2));
    SparseSmem* sj = cluster.map_shared_rank(&s, hhj >> (32 - SPARSE_CLUSTER_LOG2));
    uint32_t sloti = 0, slotj = 0, ri = 0, rj = 0;
    if (have) {
        // (first probes of both bodies in flight together, then both counts: two remote round trips instead of four)
        const uint32_t h0i = (hhi >> 16) & (SPARSE_SLOTS - 1u), h0j = (hhj >> 16) & (SPARSE_SLOTS - 1u);
        const uint32_t oldi = atomicCAS(&si->key[h0i], EMPTY, bi), oldj = atomicCAS(&sj->key[h0j], EMPTY, bj);
        sloti = (oldi == EMPTY || oldi == bi) ? h0i : sparse_insert(si->key, bi, (h0i + 1u) & (SPARSE_SLOTS - 1u));
        slotj = (oldj == EMPTY || oldj == bj) ? h0j : sparse_insert(sj->key, bj, (h0j + 1u) & (SPARSE_SLOTS - 1u));
        if (sloti == EMPTY || slotj == EMPTY) {
            s0->fail = 1u;
        } else {
            ri = atomicAdd(&si->cnt[sloti], 1u);
            rj = atomicAdd(&sj->cnt[slotj], 1u);
        }
    }
    SP_STAMP();
    cluster.sync();
    SP_STAMP();
    // ---- P1 (owners): dense local ids, half-edge segments.  A thread owns SPARSE_SLOTS / SPARSE_THREADS
    // consecutive slots; one CTA-wide exclusive scan of (occupied slots, half edges needed) places them.
    {
        constexpr uint32_t PER = SPARSE_SLOTS / SPARSE_THREADS;
        static_assert(PER * SPARSE_THREADS == SPARSE_SLOTS, "whole slots per thread");
        uint32_t nocc = 0, nneed = 0;
#pragma unroll
        for (uint32_t q = 0; q < PER; ++q) {
            const uint32_t k = s.key[tid * PER + q], raw = s.cnt[tid * PER + q];
            if (k != EMPTY) {
                ++nocc;
                nneed += raw;
            }
        }
        uint32_t iocc = nocc, ineed = nneed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y0 = __shfl_up_sync(0xFFFFFFFFu, iocc, o), y1 = __shfl_up_sync(0xFFFFFFFFu, ineed, o);
            if (lane >= (uint32_t)o) {
                iocc += y0;
                ineed += y1;
            }
        }
        if (lane == 31) {
            s.wsum[0][tid >> 5] = iocc;
            s.wsum[1][tid >> 5] = ineed;
        }
        __syncthreads();
        uint32_t l = iocc - nocc, off = ineed - nneed;
        for (uint32_t wq = 0; wq < (tid >> 5); ++wq) {
            l += s.wsum[0][wq];
            off += s.wsum[1][wq];
        }
        if (tid == SPARSE_THREADS - 1) s.nb = l + nocc;
        bool bad = false;
#pragma unroll
        for (uint32_t q = 0; q < PER; ++q) {
            const uint32_t x = tid * PER + q, k = s.key[x], raw = s.cnt[x];
            if (k == EMPTY) continue;
            const uint32_t need = raw;
            if (l >= SPARSE_BODIES) {
                bad = true;
            } else {
                TR_ASSERT(off + need <= 2u * SPARSE_MAX_CONTACTS);
                s.gid[l] = k;
                s.cnt[x] = sparse_pack(l, off, need);         // what a contact of this body needs in P2-P3, in one word
            }
            ++l;
            off += need;
        }
        if (bad) s0->fail = 1u;
    }
    SP_STAMP();
    cluster.sync();
    SP_STAMP();
    {
        const uint32_t f = s0->fail;
        if (f) {
            cluster.sync();                                  // CTA 0's shared memory outlives everybody's read of the flag
            return false;
        }
    }
    // ---- the gather has landed (or nearly); then P2: half edges; the first arrival deposits the body's state
    uint8_t fi = 0, fj = 0;
    float4 ci = make_float4(0.f, 0.f, 0.f, 0.f), cj = ci, hi = ci, hj = ci, vi = ci, vj = ci, pi = ci, pj = ci;
    asm volatile("cp.async.wait_all;" ::: "memory");           // this thread's own copies: no barrier needed to read them
    if (have) {
        fi = (uint8_t)(s.fstage[0][tid] >> (8u * (bi & 3u)));
        fj = (uint8_t)(s.fstage[1][tid] >> (8u * (bj & 3u)));
        const float4* st = reinterpret_cast<const float4*>(&s.stage[0][0]);
        ci = st[0 * SPARSE_THREADS + tid];
        cj = st[1 * SPARSE_THREADS + tid];
        vi = st[2 * SPARSE_THREADS + tid];
        vj = st[3 * SPARSE_THREADS + tid];
        pi = st[4 * SPARSE_THREADS + tid];
        pj = st[5 * SPARSE_THREADS + tid];
        hi = st[6 * SPARSE_THREADS + tid];
        hj = st[7 * SPARSE_THREADS + tid];
    }
    uint32_t li = 0, lj = 0, oi = 0, oj = 0, di = 0, dj = 0;
    uint2 sgi = make_uint2(0u, 0u), sgj = sgi;
    if (have) {
        const uint32_t wi = si->cnt[sloti], wj = sj->cnt[slotj];
        li = wi & (SPARSE_BODIES - 1u);
        lj = wj & (SPARSE_BODIES - 1u);
        sgi = make_uint2((wi >> SPARSE_BODIES_LOG2) & (2u * SPARSE_MAX_CONTACTS - 1u), wi >> (SPARSE_BODIES_LOG2 + 13));
        sgj = make_uint2((wj >> SPARSE_BODIES_LOG2) & (2u * SPARSE_MAX_CONTACTS - 1u), wj >> (SPARSE_BODIES_LOG2 + 13));
    }
    const bool sti = (fi & F_STATIC) != 0, stj = (fj & F_STATIC) != 0;
    if (have) {
        // a static body is never changed, so nothing orders its contacts: its list is neither filled nor ranked, however long
        if (!sti) {
            oi = sgi.x;
            di = sgi.y;
            si->he[oi + ri] = bj;
        }
        if (!stj) {
            oj = sgj.x;
            dj = sgj.y;
            sj->he[oj + rj] = bi;
        }
        if (di > SPARSE_MAX_DEGREE || dj > SPARSE_MAX_DEGREE) s0->fail = 1u;   // a movable hub: the general path sorts such lists
        if (ri == 0) {
            si->rec[li][0] = make_uint4(__float_as_uint(vi.x), __float_as_uint(vi.y), __float_as_uint(vi.z), 0u);
            si->rec[li][1] = make_uint4(__float_as_uint(pi.x), __float_as_uint(pi.y), __float_as_uint(pi.z), 0u);
        }
        if (rj == 0) {
            sj->rec[lj][0] = make_uint4(__float_as_uint(vj.x), __float_as_uint(vj.y), __float_as_uint(vj.z), 0u);
            sj->rec[lj][1] = make_uint4(__float_as_uint(pj.x), __float_as_uint(pj.y), __float_as_uint(pj.z), 0u);
        }
    }
    // the contact's constants (registers)
    float nx = 0.f, ny = 0.f, nz = 0.f, ima = 0.f, imb = 0.f, den = 0.f, e = 0.f;
    if (have) {
        contact_normal_values(fi, fj, ci, cj, hi, hj, nx, ny, nz);
        ima = __fdiv_rn(1.0f, pi.w);
        imb = __fdiv_rn(1.0f, pj.w);
        den = __fadd_rn(ima, imb);
        e = (vj.w < vi.w) ? vj.w : vi.w;                     // std::min(a.rest, b.rest)
    }
    SP_STAMP();
    cluster.sync();
    SP_STAMP();
    {
        const uint32_t f = s0->fail;                         // (nothing has been written outside shared memory yet)
        if (f) {
            cluster.sync();
            return false;
        }
    }
    // ---- P3: versions = rank by partner id among the body's contacts (a static body is never changed: nothing orders its contacts)
    uint32_t ei = 0, ej = 0;
    {
        const uint32_t* p = si->he + oi;                     // (di, dj are 0 for static bodies and idle threads)
        for (uint32_t x = 0; x < di; ++x) ei += p[x] < bj ? 1u : 0u;
        const uint32_t* q = sj->he + oj;
        for (uint32_t x = 0; x < dj; ++x) ej += q[x] < bi ? 1u : 0u;
    }
    SP_STAMP();
    // ---- P5: the replay.  No barrier: a thread polls the records of its contact's two bodies, applies the contact on
    // the owners' copies when both carry its versions, and stores them back with the next versions - a hand-off costs two
    // distributed-shared-memory round trips.  Every contact has its thread and all threads are resident, so the earliest
    // unfinished contact in sweep order can always run.  (P3 only read what the replay never writes.)
    bool pending = have, stalled = false;
    uint32_t spins = 0;
    while (__any_sync(0xFFFFFFFFu, pending)) {               // the warp iterates together: no lane waits inside a divergent branch
        if (!pending) continue;
        uint4 ai, bi4, aj, bj4;
        ld128_relaxed_cluster(&si->rec[li][0], ai);
        ld128_relaxed_cluster(&sj->rec[lj][0], aj);
        ld128_relaxed_cluster(&si->rec[li][1], bi4);
        ld128_relaxed_cluster(&sj->rec[lj][1], bj4);
        const bool oki = sti || ((ai.w & ~SPARSE_DIRTY) == ei && (bi4.w & ~SPARSE_DIRTY) == ei);
        const bool okj = stj || ((aj.w & ~SPARSE_DIRTY) == ej && (bj4.w & ~SPARSE_DIRTY) == ej);
        if (oki && okj) {
            float vix = __uint_as_float(ai.x), viy = __uint_as_float(ai.y), viz = __uint_as_float(ai.z);
            float vjx = __uint_as_float(aj.x), vjy = __uint_as_float(aj.y), vjz = __uint_as_float(aj.z);
            float pix = __uint_as_float(bi4.x), piy = __uint_as_float(bi4.y), piz = __uint_as_float(bi4.z);
            float pjx = __uint_as_float(bj4.x), pjy = __uint_as_float(bj4.y), pjz = __uint_as_float(bj4.z);
            const float rvx = __fsub_rn(vix, vjx), rvy = __fsub_rn(viy, vjy), rvz = __fsub_rn(viz, vjz);
            const float vn = dot3_rn(rvx, rvy, rvz, nx, ny, nz);
            const bool hit = !(vn > 0);                              // main.cpp:714 (as written)
            if (hit) {
                float jimp = __fmul_rn(-__fadd_rn(1.0f, e), vn);     // -(1 + e) * vn
                jimp = __fdiv_rn(jimp, den);                         // / (1/ma + 1/mb)
                const float ix = __fmul_rn(nx, jimp), iy = __fmul_rn(ny, jimp), iz = __fmul_rn(nz, jimp);
                const float kcorr = 0.2f * 0.01f;                    // percent * penetrationSlop (fp32 product)
                const float cx = __fmul_rn(nx, kcorr), cy = __fmul_rn(ny, kcorr), cz = __fmul_rn(nz, kcorr);
                vix = __fsub_rn(vix, __fmul_rn(ix, ima));
                viy = __fsub_rn(viy, __fmul_rn(iy, ima));
                viz = __fsub_rn(viz, __fmul_rn(iz, ima));
                pix = __fsub_rn(pix, __fmul_rn(cx, ima));
                piy = __fsub_rn(piy, __fmul_rn(cy, ima));
                piz = __fsub_rn(piz, __fmul_rn(cz, ima));
                vjx = __fadd_rn(vjx, __fmul_rn(ix, imb));
                vjy = __fadd_rn(vjy, __fmul_rn(iy, imb));
                vjz = __fadd_rn(vjz, __fmul_rn(iz, imb));
                pjx = __fadd_rn(pjx, __fmul_rn(cx, imb));
                pjy = __fadd_rn(pjy, __fmul_rn(cy, imb));
                pjz = __fadd_rn(pjz, __fmul_rn(cz, imb));
            }
            if (!sti) {                                              // (a static body is never written: main.cpp:725-741)
                const uint32_t nv = (ei + 1u) | (ai.w & SPARSE_DIRTY) | (hit ? SPARSE_DIRTY : 0u);
                st128_relaxed_cluster(&si->rec[li][0], make_uint4(__float_as_uint(vix), __float_as_uint(viy), __float_as_uint(viz), nv));
                st128_relaxed_cluster(&si->rec[li][1], make_uint4(__float_as_uint(pix), __float_as_uint(piy), __float_as_uint(piz), nv));
            }
            if (!stj) {
                const uint32_t nv = (ej + 1u) | (aj.w & SPARSE_DIRTY) | (hit ? SPARSE_DIRTY : 0u);
                st128_relaxed_cluster(&sj->rec[lj][0], make_uint4(__float_as_uint(vjx), __float_as_uint(vjy), __float_as_uint(vjz), nv));
                st128_relaxed_cluster(&sj->rec[lj][1], make_uint4(__float_as_uint(pjx), __float_as_uint(pjy), __float_as_uint(pjz), nv));
            }
            pending = false;
        } else if (++spins > SPARSE_SPIN_LIMIT) {            // a contact nobody can run - never, with a sound graph
            stalled = true;
            pending = false;
        }
    }
    {
        // for the record: the deepest position any contact has in the list of one of its bodies (= length of the longest chain through one body)
        const uint32_t depth = __reduce_max_sync(0xFFFFFFFFu, have ? max(ei, ej) + 1u : 0u);
        if (lane == 0 && depth) atomicMax(&s0->rounds, depth);
        if (__any_sync(0xFFFFFFFFu, stalled) && lane == 0) a.counters[C_ABORT] = 1u;   // reported like the executors' watchdog
    }
    SP_STAMP();
    cluster.sync();                                          // the replay is complete everywhere; nobody touches remote shared memory any more
    rounds = s.rounds;                                       // (CTA 0's is the cluster's; only CTA 0 uses it)
    // ---- P6 (owners): changed bodies back to the SoA state
    const uint32_t nb = s.nb;
    for (uint32_t l = tid; l < nb; l += SPARSE_THREADS) {
        const uint4 ra = s.rec[l][0], rb = s.rec[l][1];
        if (!(ra.w & SPARSE_DIRTY)) continue;
        const uint32_t b = s.gid[l];
        store_xyz(a.vel_rest + b, make_float4(__uint_as_float(ra.x), __uint_as_float(ra.y), __uint_as_float(ra.z), 0.f));
        store_xyz(a.pos_m + b, make_float4(__uint_as_float(rb.x), __uint_as_float(rb.y), __uint_as_float(rb.z), 0.f));
    }
    SP_STAMP();
#ifdef TR_SPARSE_PROF
    if (rank == 0 && tid == 0) {
        printf("[TR_SPARSE_PROF] contacts %u rounds %u; cycles init | sync | P0 | sync | P1 | sync | gather wait + P2 + constants | sync | P3 | replay | sync + write-back:", cn, rounds);
        for (int k = 1; k < sp_n; ++k) printf(" %lld", sp_t[k] - sp_t[k - 1]);
        printf("\n");
    }
#endif
#undef SP_STAMP
    return true;
}

__global__ void __cluster_dims__(SPARSE_CLUSTER, 1, 1) __launch_bounds__(SPARSE_THREADS) sparse_step_kernel(SparseArgs a) {
    extern __shared__ __align__(16) unsigned char sparse_raw[];
    SparseSmem& s = *reinterpret_cast<SparseSmem*>(sparse_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank(), tid = threadIdx.x;
    pdl_wait();
    if (rank == 0 && tid == 0) *reinterpret_cast<unsigned long long*>(a.counters + C_T0 + 4) = globaltimer_ns();
    // this thread's contact, if the step has that many: fetched together with the count (one round trip instead of two)
    const uint32_t cidx = rank * SPARSE_THREADS + tid;
    const unsigned long long key = cidx < a.cap ? a.ckeys[cidx] : 0ull;
    const bool dead = a.ctl[CTL_ABORT] != 0;                 // an earlier step has to be re-run first
    const unsigned long long found = *reinterpret_cast<const unsigned long long*>(a.counters + C_CONTACTS);
    if (!dead && cidx < found && found <= SPARSE_MAX_CONTACTS && found <= a.cap) {
        // ... and its bodies straight behind it, as ASYNCHRONOUS copies into this thread's staging slots: after an L2 flush they
        // cost ~3 us of DRAM + page-walk latency, which now runs under P0 and P1 - an ordinary load in flight would be waited for
        // by the first cluster barrier (its release is a GPU-scope MEMBAR), and a prefetch is dropped on the TLB miss (both measured)
        const uint32_t gi = (uint32_t)(key >> 32), gj = (uint32_t)key;
        TR_ASSERT(gi < gj && gj < a.n);
        {
            const void* src[8] = {a.center_r + gi, a.center_r + gj, a.vel_rest + gi, a.vel_rest + gj, a.pos_m + gi, a.pos_m + gj, a.half_ext + gi, a.half_ext + gj};
#pragma unroll
            for (int k = 0; k < 8; ++k)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&s.stage[k][tid])), "l"(src[k]) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(&s.fstage[0][tid])), "l"(a.flags + (gi & ~3u)) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(&s.fstage[1][tid])), "l"(a.flags + (gj & ~3u)) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t rounds = 0, extra = 0;
    bool bail = false;
    if (!dead && found != 0) {
        if (found > a.cap) bail = true;                                      // a real overflow: the buffers must grow
        else if (found > SPARSE_MAX_CONTACTS || !sparse_replay(a, s, cluster, (uint32_t)found, key, rounds)) {
            bail = true;
            extra = 8u;                                                      // not a sparse step: the general kernels must (re-)run it
        } else {
            extra = 16u;
        }
    } else if (!dead) {
        extra = 16u;
    }
    if (rank != 0) return;
    if (a.tail_follows) {
        // the general kernels come next in this graph: they take what this kernel left alone, and the epilogue writes the record
        if (tid == 0 && extra == 16u) a.counters[C_SPARSE] = 1u + rounds;
        return;
    }
    if (tid == 0) {
        if (bail) a.ctl[CTL_ABORT] = a.ctl[CTL_SEQ] + 1u;
        write_step_record(a.counters, a.ctl, a.ring, extra, rounds);
    }
    __syncthreads();
    for (uint32_t k = tid; k < (uint32_t)C_NCOUNTERS; k += blockDim.x) a.counters[k] = 0;
}

}  // namespace tr

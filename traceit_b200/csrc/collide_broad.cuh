// K2-K5c: grid keys + histogram, record scatter, on-demand stable order, narrow phase, side and far lists.
// Device code only; included by collide.cu after collide_common.cuh.
#pragma once
#include "collide_common.cuh"

namespace tr {

// ---------------------------------------------------------------------------------------
// K2: keys + classification (+ per-cell histogram on the fast path)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) keys_kernel(const float4* __restrict__ center_r, const float4* __restrict__ half_ext,
                                                   const uint8_t* __restrict__ flags, uint32_t n, GridParams g,
                                                   uint32_t* __restrict__ keys,
                                                   uint8_t* __restrict__ klass, uint32_t* __restrict__ side_list,
                                                   uint32_t* __restrict__ far_list, float4* __restrict__ far_c,
                                                   uint32_t* __restrict__ bnd_list, float4* __restrict__ bnd_c,
                                                   uint32_t* __restrict__ cell_count, uint32_t* counters) {
    pdl_trigger();   // (first kernel of the step: an ordinary launch, nothing to wait for)
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *reinterpret_cast<unsigned long long*>(counters + C_T0) = globaltimer_ns();   // start of the broad phase
    {
        // stream the (zeroed) histogram into L2 ahead of the scattered REDs: one 128-byte line per
        // thread of the leading CTAs, so random cloud ids do not pay a DRAM round trip per atomic
        const uint32_t ncell = 1u << g.total_bits();
        if (i < ncell / 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(cell_count + (size_t)i * 32));
    }
    if (i < n) {
        const float4 c = center_r[i];
        const uint8_t f = flags[i];
        const uint32_t sentinel = 1u << g.total_bits();
        uint32_t key = sentinel;
        uint8_t k = 0;
        if (f & F_LIVE) {   // main.cpp:696 drops mass<=0 bodies from every pair (flag set at upload)
            float ux = __fmul_rn(__fsub_rn(c.x, g.ox), g.inv_cell);
            float uy = __fmul_rn(__fsub_rn(c.y, g.oy), g.inv_cell);
            float uz = __fmul_rn(__fsub_rn(c.z, g.oz), g.inv_cell);
            const bool in_domain = fabsf(ux) < 32768.0f && fabsf(uy) < 32768.0f && fabsf(uz) < 32768.0f;
            const bool finite_c = fabsf(c.x) <= 3.0e38f && fabsf(c.y) <= 3.0e38f && fabsf(c.z) <= 3.0e38f;   // false for NaN / inf
            bool fits;   // extent admitted to the grid
            if (g.mixed) {
                // any body may meet a box, and box pairs use collider::size of BOTH bodies (main.cpp:704-705)
                const float4 h = half_ext[i];
                float ext = fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fabsf(h.z));
                const bool finite_h = fabsf(h.x) <= 3.0e38f && fabsf(h.y) <= 3.0e38f && fabsf(h.z) <= 3.0e38f;
                if (f & F_SPHERE) ext = fmaxf(ext, fabsf(c.w));
                fits = finite_h && ext <= g.rmax_grid && (!(f & F_SPHERE) || fabsf(c.w) <= g.rmax_grid);
            } else {
                fits = (f & F_SPHERE) && fabsf(c.w) <= g.rmax_grid;
            }
            if (fits && in_domain) {
                key = pack_key(__float2int_rd(ux), __float2int_rd(uy), __float2int_rd(uz), g);
                TR_ASSERT(key < sentinel);
                k = 1;
                atomicAdd(&cell_count[key], 1u);   // result unused: a fire-and-forget RED
                // grid bodies within two cells of the domain boundary are the only ones a "far" body can touch
                if (fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz)) >= 32766.0f) {
                    const uint32_t slot = atomicAdd(&counters[C_BND], 1u);
                    TR_ASSERT(slot < n);
                    bnd_list[slot] = pack_val(i, f);
                    bnd_c[slot] = c;
                }
            } else if (fits && finite_c) {
                // "far": normal extent, finite centre, outside the grid domain.  It can only touch other
                // far bodies, boundary grid bodies (computed |u| differs by >= 2 cells from everyone else:
                // more than the largest admitted reach) and oversize side-list bodies.
                k = 4;
                const uint32_t slot = atomicAdd(&counters[C_FAR], 1u);
                TR_ASSERT(slot < n);
                far_list[slot] = pack_val(i, f);
                far_c[slot] = c;
            } else if (fits && g.rmax_grid < 1.0e18f) {
                // "lost": a NaN / infinite centre makes both predicates false (delta is +-inf or NaN)
                // unless the radius / size SUM is itself infinite, i.e. the partner is an oversize
                // side-list body.  Not a grid body, not a side-list body; side_kernel still tests it
                // against every side-list body.  (Bodies the reference's inverted approach test blows
                // up end here instead of turning the side list into an O(N^2) sweep.)
                k = 3;
            } else {
                k = 2;
                uint32_t slot = atomicAdd(&counters[C_SIDE], 1u);
                TR_ASSERT(slot < n);
                side_list[slot] = i;
            }
        }
        keys[i] = key;
        klass[i] = k;
    }
}

// K3b: one record per grid body into its cell's segment; counting the histogram back down leaves it
// zeroed for the next step.  The record carries everything the narrow phase reads about a body -
// centre + radius, id|flags, AABB half extents - in ONE 32-byte sector written by one 256-bit store
// (SASS STG.E.ENL2.256) and read back by one 256-bit load, so nothing is gathered by id afterwards.
// (Measured and rejected: a compact 20-byte layout for all-sphere worlds - float4 array + id array - halves the
// narrow phase's load wavefronts, 45.7 -> 40.4 us, but K3b's two partial-sector stores cost the random cloud's
// broad phase 34.3 -> 46.2 us; pack: 77.6 -> 74.6 and 23.7 -> 22.6.)
struct __align__(32) SlotRec {
    float4 c;              // collider centre, radius
    uint32_t val;          // id | BOX_BIT | STATIC_BIT
    float hx, hy, hz;      // collider::size (meaningful in mixed worlds only)
};
static_assert(sizeof(SlotRec) == 32, "one sector");

__device__ __forceinline__ void st_rec(SlotRec* p, const float4 c, uint32_t val, const float4 h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(c.x)), "r"(__float_as_uint(c.y)),
                 "r"(__float_as_uint(c.z)), "r"(__float_as_uint(c.w)), "r"(val), "r"(__float_as_uint(h.x)),
                 "r"(__float_as_uint(h.y)), "r"(__float_as_uint(h.z))
                 : "memory");
}
__device__ __forceinline__ void ld_rec(const SlotRec* p, float4& c, uint32_t& val, float4& h) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "l"(p));
    c = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
    val = r4;
    h = make_float4(__uint_as_float(r5), __uint_as_float(r6), __uint_as_float(r7), 0.f);
}

// the key of an in-domain centre: the same operations as K2, hence the same key
__device__ __forceinline__ uint32_t cell_key(const float4 c, const GridParams& g) {
    const float ux = __fmul_rn(__fsub_rn(c.x, g.ox), g.inv_cell);
    const float uy = __fmul_rn(__fsub_rn(c.y, g.oy), g.inv_cell);
    const float uz = __fmul_rn(__fsub_rn(c.z, g.oz), g.inv_cell);
    return pack_key(__float2int_rd(ux), __float2int_rd(uy), __float2int_rd(uz), g);
}

__global__ void __launch_bounds__(256) scatter_kernel(const uint32_t* __restrict__ keys, uint32_t sentinel,
                                                      const uint8_t* __restrict__ flags, const float4* __restrict__ center_r,
                                                      const float4* __restrict__ half_ext,   // NULL unless boxes are in the grid
                                                      uint32_t n, const uint32_t* __restrict__ cell_start,
                                                      uint32_t* __restrict__ cell_count, SlotRec* __restrict__ recs) {
    pdl_wait();
    pdl_trigger();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = keys[i];
    if (key >= sentinel) return;          // side-list, far, lost or dropped body
    const float4 c = center_r[i];
    const uint32_t val = pack_val(i, flags[i]);
    const uint32_t old = atomicSub(&cell_count[key], 1u);
    const uint32_t slot = cell_start[key] + old - 1u;
    TR_ASSERT(old >= 1u && old <= n && slot < n && slot < cell_start[key + 1]);
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
    if (half_ext) h = half_ext[i];
    st_rec(recs + slot, c, val, h);
}

// K4 (on demand, tr_debug_grid): the stable (key,id) order.  The rank of a body inside its cell is the number of
// co-occupants with a smaller id: counted directly in cells of up to CANON_SMALL bodies; a crowded cell is put on a
// list by the thread of its first slot and sorted by a CTA (canon_big_kernel), so that a cell holding L bodies
// costs O(L log^2 L), not L^2.
constexpr uint32_t CANON_SMALL = 32;
__global__ void __launch_bounds__(256) canon_kernel(const SlotRec* __restrict__ recs, const uint32_t* __restrict__ cell_start,
                                                    GridParams g, uint32_t* counters, uint32_t* __restrict__ sorted_ids,
                                                    uint32_t* __restrict__ big_cells) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= counters[C_GRID]) return;
    const uint32_t key = cell_key(recs[t].c, g);
    const uint32_t id = recs[t].val & ID_MASK;
    const uint32_t s = cell_start[key], e = cell_start[key + 1];
    if (e - s > CANON_SMALL) {
        if (t == s) big_cells[atomicAdd(&counters[C_LONG], 1u)] = key;
        return;
    }
    uint32_t rank = 0;
    for (uint32_t u = s; u < e; ++u) rank += ((recs[u].val & ID_MASK) < id) ? 1u : 0u;
    sorted_ids[s + rank] = id;
}

__global__ void __launch_bounds__(256) canon_big_kernel(const SlotRec* __restrict__ recs, const uint32_t* __restrict__ cell_start,
                                                        const uint32_t* __restrict__ counters, const uint32_t* __restrict__ big_cells,
                                                        uint32_t* __restrict__ sorted_ids, uint2* __restrict__ scratch) {
    __shared__ uint2 s_sort[PREP_SMEM_SORT];
    const uint32_t nbig = counters[C_LONG];
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        const uint32_t key = big_cells[b];
        const uint32_t s = cell_start[key], len = cell_start[key + 1] - s;
        if (len <= PREP_SMEM_SORT) {
            for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) s_sort[k] = make_uint2(recs[s + k].val & ID_MASK, 0u);
            __syncthreads();
            smem_sort(s_sort, len, false);
            for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) sorted_ids[s + k] = s_sort[k].x;
            __syncthreads();
        } else {
            // beyond the shared-memory size: the same network in global memory (scratch is at least as long as recs)
            uint2* p = scratch + s;
            for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) p[k] = make_uint2(recs[s + k].val & ID_MASK, 0u);
            __syncthreads();
            uint32_t pow2 = 1;
            while (pow2 < len) pow2 <<= 1;
            uint32_t i, l;
            for (uint32_t k = 2; k <= pow2; k <<= 1) {
                for (uint32_t t = threadIdx.x; t < pow2 / 2; t += blockDim.x) {
                    flip_pair(t, k, i, l);
                    global_cmpx(p, i, l, len);
                }
                __syncthreads();
                for (uint32_t j = k / 4; j >= 1; j >>= 1) {
                    for (uint32_t t = threadIdx.x; t < pow2 / 2; t += blockDim.x) {
                        shear_pair(t, j, i, l);
                        global_cmpx(p, i, l, len);
                    }
                    __syncthreads();
                }
            }
            for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) sorted_ids[s + k] = __ldcg(p + k).x;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// K5: narrow phase over the half stencil, (body, candidate) PAIRS balanced over the lanes of a warp
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_pair(unsigned long long* out, unsigned long long cap, unsigned long long* counter,
                                          uint32_t i, uint32_t j) {
    const unsigned long long slot = warp_reserve(counter);   // one atomic for the lanes that hit together
    if (slot < cap) out[slot] = ((unsigned long long)i << 32) | j;
}

// The candidate relation "cell(b) is in the 27-cell stencil of cell(a)" is symmetric, so every
// unordered pair is reached from exactly one side by a HALF stencil: the later slots of a's own
// cell, and the 13 cells at "forward" offsets (dz,dy,dx) > (0,0,0) lexicographically (an offset o
// and its opposite -o are distinct modulo the per-axis table size >= 4, and exactly one of them is forward).
// With x as the fastest key digit those are 5 contiguous slot ranges ("windows") for an interior cell:
//   [t+1, end of cell cx+1]  in the own row, and cells cx-1..cx+1 of the rows (dz,dy) =
//   (0,+1), (+1,-1), (+1,0), (+1,+1).
// "Later slots of the own cell" is any fixed order: the arrival order K3b left is as good as the id order.
// MIXED: boxes live in the grid; a pair with at least one box uses the AABB predicate on the
// collider sizes of both bodies, exactly as the reference dispatches (main.cpp:701-706).
//
// Work distribution.  Round 1 gave every lane one body and let it walk its five windows: a window
// holds a Poisson(1.5) number of records on the sparse cloud, a warp runs the maximum over its lanes
// window by window, and ncu counted 10.4 of 32 lanes active per instruction.  Now a warp owns 32
// consecutive slots and works in two phases:
//   expand   every lane writes its (candidate slot, owner lane) pairs into a per-warp list in shared
//            memory at the offset a warp scan of the window lengths gives it (3 instructions per pair);
//   test     the lanes walk the flattened list 32 pairs at a time: own record from shared memory, the
//            candidate's record by one 256-bit load, exact predicate; every lane busy.
// Hits are staged in a per-warp buffer and written out with ONE atomic per flush (same-address atomics
// retire at ~0.85 clk each at L2: emitting from inside the loop was the packed narrow phase in round 1).
// The 2 of 2^bx cells per row at the x edges, whose three-cell runs wrap around the row end, simply have two
// windows per row instead of one (a lane-serial special path for them - 1.6 % of the bodies, one active lane
// per warp - was measured at 27 % of the kernel's executed instructions).
constexpr int NARROW_THREADS = 128;
constexpr int NARROW_WARPS = NARROW_THREADS / 32;
constexpr int NARROW_LIST = 512;     // (candidate, owner) pairs per warp and round: 16 per lane (mean 7 sparse, 14.5 packed)
constexpr int NARROW_STAGE = 96;     // staged hits per warp
#ifndef TR_NARROW_UNROLL
#define TR_NARROW_UNROLL 2
#endif
constexpr int NARROW_UNROLL = TR_NARROW_UNROLL;   // candidate records in flight per lane

struct NarrowWarp {
    uint32_t u[NARROW_LIST];                   // candidate slot << 5 | owner lane (worlds of more than 2^27 grid bodies: slot only)
    uint8_t o[NARROW_LIST];                    // owner lane, only used by those worlds
    unsigned long long hit[NARROW_STAGE];      // (min id << 32) | max id
    float4 own_c[32];                          // the warp's own bodies: centre + radius,
    uint32_t own_v[32];                        //   id|flags,
    float4 own_h[32];                          //   half extents (mixed worlds)
};

__device__ __forceinline__ void narrow_flush(NarrowWarp& sw, uint32_t& nh, uint32_t lane, unsigned long long* out,
                                             unsigned long long cap, unsigned long long* counter) {
    if (nh == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)nh);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    for (uint32_t k = lane; k < nh; k += 32)
        if (base + k < cap) out[base + k] = sw.hit[k];
    nh = 0;
    __syncwarp();
}

template <bool CANDIDATES, bool MIXED>
__device__ __forceinline__ bool pair_hit(const float4 ca, uint32_t va, const float4 ha, const float4 cb, uint32_t vb, const float4 hb) {
    if (CANDIDATES) return true;
    bool hit;
    if (MIXED && ((va | vb) & BOX_BIT)) hit = aabb_hit(ca, ha, cb, hb);   // main.cpp:704
    else hit = sphere_hit(ca, cb);                                       // main.cpp:702
    return hit && !(va & vb & STATIC_BIT);                               // main.cpp:695
}

// Blocks [0, slot_blocks) run the grid narrow phase; the blocks behind them are the "outside" workers
// (side list and far list, K5b / K5c below) so that the whole detection is one launch.
struct OutsideArgs {
    const uint32_t* side_list;
    const uint32_t* far_list;
    const float4* far_c;
    const uint32_t* bnd_list;
    const float4* bnd_c;
    const float4* center_r;
    const float4* half_ext;
    const uint8_t* flags;
    const uint8_t* klass;
    uint32_t n;
};
__device__ void outside_worker(uint32_t worker, uint32_t nworkers, const OutsideArgs& oa, const uint32_t* counters,
                               unsigned long long* out, unsigned long long cap, unsigned long long* counter);

// BIG: worlds of more than 2^27 bodies, where slot << 5 | lane does not fit one word (the host picks the instantiation)
// (Occupancy is tuned, not left to ptxas: 10 CTAs of 128 threads per SM = 48 registers.  Measured on the 1M cloud / pack:
// 41.6 / 73 us; 11 or 12 CTAs (46 / 40 registers) 49.6 / 86; unconstrained 47.6 / 82.5; 1, 3 or 4 candidate records in
// flight per lane instead of 2: 43.5 / 77, 49 / 86, 54.6 / 92.)
template <bool CANDIDATES, bool MIXED, bool BIG>
__global__ void __launch_bounds__(NARROW_THREADS, MIXED ? 8 : 10) narrow_kernel(const SlotRec* __restrict__ recs,
                                                                const uint32_t* __restrict__ cell_start,
                                                                uint32_t* counters, GridParams g, uint32_t slot_blocks,
                                                                OutsideArgs oa, unsigned long long* out, unsigned long long cap,
                                                                unsigned long long* counter) {
    __shared__ NarrowWarp s_warp[NARROW_WARPS];
    pdl_wait();
    pdl_trigger();
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<unsigned long long*>(counters + C_T0 + 2) = globaltimer_ns();
    if (blockIdx.x >= slot_blocks) {
        if (!CANDIDATES) outside_worker(blockIdx.x - slot_blocks, gridDim.x - slot_blocks, oa, counters, out, cap, counter);
        return;
    }
    const uint32_t ngrid = counters[C_GRID];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * NARROW_THREADS + threadIdx.x;
    if (t - lane >= ngrid) return;                 // the whole warp is past the last grid body
    NarrowWarp& sw = s_warp[wid];
    const bool valid = t < ngrid;
    // Windows of this lane's body: per stencil row (own row, then the four forward rows) one slot range, plus a
    // second one in the 2 of 2^bx cells per row whose three-cell run wraps around the row end (cx = 0: cells
    // {0,1} + {mx}; cx = mx: cells {mx-1,mx} + {0}).  Ten (start, length) pairs, the second five almost always empty.
    uint32_t rs[10], len[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) rs[r] = len[r] = 0;
    constexpr bool big = BIG;
    if (valid) {
        float4 ca, ha;
        uint32_t va;
        ld_rec(recs + t, ca, va, ha);
        if (MIXED) sw.own_h[lane] = ha;
        sw.own_c[lane] = ca;
        sw.own_v[lane] = va;
        const uint32_t key = cell_key(ca, g);
        const uint32_t mx = (1u << g.bx) - 1u, my = (1u << g.by) - 1u, mz = (1u << g.bz) - 1u;
        const uint32_t cx = key & mx, cy = (key >> g.bx) & my, cz = (key >> (g.bx + g.by)) & mz;
        // the run of cells cx-1..cx+1 of a row, as cell offsets inside the row: [a0, a1) and, when it wraps, [b0, b1)
        const uint32_t a0 = cx == 0 ? 0u : cx - 1u, a1 = cx == mx ? mx + 1u : cx + 2u;
        const bool wrap = cx == 0 || cx == mx;
        const uint32_t b0 = cx == 0 ? mx : 0u;
        // own row: the later slots of the own cell and cell cx+1 (cell 0 of the row when cx = mx)
        rs[0] = t + 1;
        len[0] = cell_start[key + (cx == mx ? 1u : 2u)] - rs[0];
        if (cx == mx) {
            rs[5] = cell_start[key - cx];
            len[5] = cell_start[key - cx + 1u] - rs[5];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int dz = r == 0 ? 0 : 1;
            const int dy = r == 0 ? 1 : r - 2;
            const uint32_t row = (((cy + dy) & my) << g.bx) | (((cz + dz) & mz) << (g.bx + g.by));
            rs[r + 1] = cell_start[row + a0];
            len[r + 1] = cell_start[row + a1] - rs[r + 1];
            if (wrap) {
                rs[r + 6] = cell_start[row + b0];
                len[r + 6] = cell_start[row + b0 + 1u] - rs[r + 6];
            }
        }
    }
    const bool any_wrap = __any_sync(0xFFFFFFFFu, (len[5] | len[6] | len[7] | len[8] | len[9]) != 0);
    // offsets of every lane's pairs in the warp's flattened list
    const uint32_t mine = (len[0] + len[1] + len[2] + len[3] + len[4]) + (len[5] + len[6] + len[7] + len[8] + len[9]);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t off = incl - mine;
#pragma unroll
    for (int r = 0; r < 10; ++r) TR_ASSERT(len[r] <= ngrid && (len[r] == 0 || rs[r] + len[r] <= ngrid));
    uint32_t nh = 0;                               // staged hits (warp-uniform)
    __syncwarp();
    for (uint32_t base = 0; base < total; base += NARROW_LIST) {
        // ---- expand: this round covers list positions [base, base + NARROW_LIST)
        uint32_t wo = off;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            if (r == 5 && !any_wrap) break;        // (warp-uniform) the second windows are all empty
            const uint32_t lo = max(wo, base), hi = min(wo + len[r], base + (uint32_t)NARROW_LIST);
            if (!big) {
                for (uint32_t k = lo; k < hi; ++k) {
                    TR_ASSERT(k - base < (uint32_t)NARROW_LIST);
                    sw.u[k - base] = ((rs[r] + (k - wo)) << 5) | lane;
                }
            } else {
                for (uint32_t k = lo; k < hi; ++k) {
                    sw.u[k - base] = rs[r] + (k - wo);
                    sw.o[k - base] = (uint8_t)lane;
                }
            }
            wo += len[r];
        }
        __syncwarp();
        // ---- test: NARROW_UNROLL x 32 pairs per iteration, the candidates' records of one iteration all in flight together
        const uint32_t cnt = min((uint32_t)NARROW_LIST, total - base);
        for (uint32_t p0 = 0; p0 < cnt; p0 += 32 * NARROW_UNROLL) {
            float4 cb[NARROW_UNROLL], hb[NARROW_UNROLL];
            uint32_t vb[NARROW_UNROLL], ow[NARROW_UNROLL];
#pragma unroll
            for (int k = 0; k < NARROW_UNROLL; ++k) {
                const uint32_t p = p0 + 32 * k + lane;
                ow[k] = 0xFFu;
                if (p < cnt) {
                    uint32_t u = sw.u[p];
                    if (!big) {
                        ow[k] = u & 31u;
                        u >>= 5;
                    } else {
                        ow[k] = sw.o[p];
                    }
                    TR_ASSERT(u < ngrid && ow[k] < 32u);
                    ld_rec(recs + u, cb[k], vb[k], hb[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < NARROW_UNROLL; ++k) {
                if (p0 + 32 * k >= cnt) break;                 // warp-uniform
                bool hit = false;
                unsigned long long pr = 0;
                if (ow[k] != 0xFFu) {
                    const uint32_t o = ow[k];
                    const float4 oc = sw.own_c[o];
                    const uint32_t ov = sw.own_v[o];
                    float4 oh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MIXED) oh = sw.own_h[o];
                    hit = pair_hit<CANDIDATES, MIXED>(oc, ov, oh, cb[k], vb[k], hb[k]);
                    const uint32_t ida = ov & ID_MASK, idb = vb[k] & ID_MASK;
                    pr = ((unsigned long long)min(ida, idb) << 32) | max(ida, idb);
                }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
                if (m) {
                    TR_ASSERT(nh + 32u <= (uint32_t)NARROW_STAGE);
                    if (hit) sw.hit[nh + __popc(m & ((1u << lane) - 1u))] = pr;
                    nh += __popc(m);
                    __syncwarp();
                    if (nh > (uint32_t)(NARROW_STAGE - 32)) narrow_flush(sw, nh, lane, out, cap, counter);
                }
            }
        }
        __syncwarp();
    }
    narrow_flush(sw, nh, lane, out, cap, counter);
}

// K5b: side-list bodies (oversize / non-finite extent) against every live body.  Persistent workers:
// a worker takes tiles of 128 bodies (grid-stride) and walks the side list for each of its bodies.
// K5c: "far" bodies (normal extent, finite centre, outside the grid domain |u| < 32768) against each
// other and against the grid bodies within two cells of the domain boundary; K5b covers
// far-vs-oversize.  A contact implies |delta u| < 1 per axis (cell = 2*rmax*(1+2^-7)) and computed
// u carries < 0.01 of error at that magnitude, so a grid body that touches a far body has
// |u| > 32766 on that axis: nobody else needs testing.  Both lists carry id|flags (pack_val) and a
// compact copy of the centres, so the partner loop streams consecutive 16-byte entries instead of
// gathering by id.  Work items are (far body, slice of the partner list) so that a few far bodies
// next to a long boundary list still spread over the workers.
__device__ void outside_worker(uint32_t worker, uint32_t nworkers, const OutsideArgs& oa, const uint32_t* counters,
                               unsigned long long* out, unsigned long long cap, unsigned long long* counter) {
    const uint32_t side_count = counters[C_SIDE], nfar = counters[C_FAR], nbnd = counters[C_BND];   // written by K2
    if (side_count == 0 && nfar == 0) return;
    if (side_count) {
        const uint32_t ntiles = (oa.n + NARROW_THREADS - 1) / NARROW_THREADS;
        for (uint32_t tile = worker; tile < ntiles; tile += nworkers) {
            const uint32_t x = tile * NARROW_THREADS + threadIdx.x;
            if (x >= oa.n) continue;
            const uint8_t kx = oa.klass[x];
            if (kx == 0) continue;                 // mass <= 0: dropped from every pair (main.cpp:696)
            const float4 cx = oa.center_r[x], hx = oa.half_ext[x];
            const uint8_t fx = oa.flags[x];
            for (uint32_t si = 0; si < side_count; ++si) {
                const uint32_t s = oa.side_list[si];
                if (s == x) continue;
                if (kx == 2 && x < s) continue;    // side-side pairs are emitted once, from the higher id's row
                const uint8_t fs = oa.flags[s];
                if ((fx & F_STATIC) && (fs & F_STATIC)) continue;             // main.cpp:695
                const float4 cs = oa.center_r[s];
                bool hit;
                if ((fx & F_SPHERE) && (fs & F_SPHERE)) hit = sphere_hit(cx, cs);      // main.cpp:702
                else hit = aabb_hit(cx, hx, cs, oa.half_ext[s]);                       // main.cpp:704
                if (hit) emit_pair(out, cap, counter, min(x, s), max(x, s));
            }
        }
    }
    if (nfar) {
        // items = (tile of 128 consecutive far bodies, slice of 1024 partners): one far body per thread, and every thread
        // of the CTA reads the SAME partner entry in each iteration (a broadcast, one transaction per warp).
        // Partners of far body t: the far bodies later in the list, then every boundary grid body.
        const uint32_t total = nfar + nbnd;
        const uint32_t slices = (total + 1023) / 1024;
        const uint32_t tiles = (nfar + NARROW_THREADS - 1) / NARROW_THREADS;
        const unsigned long long items = (unsigned long long)tiles * slices;
        for (unsigned long long it = worker; it < items; it += nworkers) {
            const uint32_t tile = (uint32_t)(it / slices), sl = (uint32_t)(it % slices);
            const uint32_t t = tile * NARROW_THREADS + threadIdx.x;
            const uint32_t k1 = min(total, (sl + 1) * 1024u);
            uint32_t k0 = sl * 1024u;
            if (k1 <= tile * NARROW_THREADS + 1u) continue;   // the whole slice lies before the tile's first body: nothing later
            if (t >= nfar) continue;
            const uint32_t va = oa.far_list[t];
            const uint32_t a = va & ID_MASK;
            const float4 ca = oa.far_c[t];
            for (uint32_t k = k0; k < k1; ++k) {
                const bool kf = k < nfar;
                if (kf && k <= t) continue;                              // far-far pairs once, from the earlier body
                const uint32_t vb = kf ? oa.far_list[k] : oa.bnd_list[k - nfar];
                if (va & vb & STATIC_BIT) continue;                      // main.cpp:695
                const float4 cb = kf ? oa.far_c[k] : oa.bnd_c[k - nfar];
                const uint32_t b = vb & ID_MASK;
                bool hit;
                if ((va | vb) & BOX_BIT) hit = aabb_hit(ca, oa.half_ext[a], cb, oa.half_ext[b]);   // main.cpp:704
                else hit = sphere_hit(ca, cb);                                               // main.cpp:702
                if (hit) emit_pair(out, cap, counter, min(a, b), max(a, b));
            }
        }
    }
}

}  // namespace tr

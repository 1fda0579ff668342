// K2-K6 - sphere/AABB collision detection and order-faithful response.
//
// Replaces resolveCollisions (reference src/main.cpp:673-745): an O(N^2) lexicographic
// i<j sweep with checkSphereCollision / checkAABBCollision (src/main.cpp:648-660) and a
// Gauss-Seidel impulse + positional nudge (src/main.cpp:709-742).
//
// Broad phase = ONE-DIGIT radix sort of the wrapped cell keys (radix = number of cells):
//   K2  keys_kernel          one thread per body: cell key from the collider centre, per-cell
//                            histogram by atomics; bodies that cannot live in the grid (oversize,
//                            out-of-domain, non-finite) go to the side / far lists; mass<=0
//                            bodies are dropped (they can never collide, main.cpp:696).
//   K3a tile_sums + scan     exclusive scan of the histogram in two unchained kernels
//                            -> dense cell_start[ncell+1] table.
//   K3b scatter_kernel       one 32-byte record per body (centre+radius, id|flags, half extents)
//                            into its cell's segment: the bodies in cell order.  Inside a cell
//                            the records are in arrival order; nothing downstream depends on it
//                            (pairs are reported as (min id, max id) and re-sorted for K6).
//   K4  canon_kernel         ON DEMAND (tr_debug_grid): the stable (key,id) order - rank of each id
//                            inside its cell - as the list of sorted ids.
//   K5  narrow_kernel        half of the 27-cell stencil (each pair reached once) as 5 contiguous
//                            key ranges, exact reference predicate (no FMA contraction).
//   K5b side_kernel          every side-list body (oversize / non-finite extent) against every live body.
//   K5c far_kernel           bodies outside the grid domain against each other and against the
//                            grid bodies next to the domain boundary.
//   K6  ordered response     contacts counting-sorted into (i,j) and (j,i) order give each contact
//                            its two predecessors / successors; the executor replays the
//                            reference's Gauss-Seidel order exactly: a contact runs when
//                            the previous contact of BOTH its bodies has run, so results
//                            are bit-identical to the sequential sweep.
//
// The contact SET is a pure function of collider centres/radii (the sweep never writes
// them: main.cpp:655-659 vs 725-741), which is what makes detection parallel and K6 legal.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "world.cuh"

namespace cg = cooperative_groups;

namespace tr {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t ID_MASK = 0x3FFFFFFFu;      // body ids fit 30 bits (tr_bodies_add enforces it)
constexpr uint32_t STATIC_BIT = 0x80000000u;   // collider::isStatic
constexpr uint32_t BOX_BIT = 0x40000000u;      // !collider::isSphere: pairs with it use the AABB predicate
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// counters[] slots
// The three counters every dataflow worker hammers (queue head, queue tail, finished work) sit on
// 128-byte lines of their own: on one line their atomics and the idle workers' polls all queue at
// one L2 slice (having idle workers merely READ the tail next to the pushes doubled the step time).
enum { C_SIDE = 0, C_GRID = 1, C_CONTACTS = 2 /*u64*/, C_CAND = 4 /*u64*/, C_FRONT = 8 /*3*/, C_ROUNDS = 11, C_ABORT = 12, C_FIN = 13,
       C_FAR = 16, C_BND = 17, C_NHOST = 24 /* read back by the host */,
       C_QHEAD = 32, C_QTAIL = 64, C_DONE = 96, C_NCOUNTERS = 128 };

}  // namespace tr

struct Collide {
    // grid description (host decides, see collide_classify)
    float cell = 1.f, inv_cell = 1.f, rmax_grid = 0.f;
    float origin[3] = {0, 0, 0};
    int bits = 2;
    bool may_have_side = true;    // some body is statically outside the grid (oversize / NaN extent)
    bool mixed = false;           // live boxes exist: they share the grid, extents include collider::size
    uint64_t cap_bodies = 0;      // arrays below sized for this many bodies
    uint64_t ncell_alloc = 0;
    uint64_t cap_contacts = 0;

    // per body
    uint32_t* keys = nullptr;        // K2 out: key, or sentinel 1<<(3*bits) for non-grid bodies
    struct tr_slot_rec* recs = nullptr;   // K3b out: tr::SlotRec per grid body, grouped by cell (arrival order inside a cell)
    uint32_t* sorted_ids = nullptr;  // K4 (on demand): ids in the stable (key,id) order
    uint32_t* tile_sum = nullptr;    // scan: per-tile histogram sums
    uint64_t tile_sum_cap = 0;
    uint32_t* cell_count = nullptr;  // [ncell] histogram; self-cleaning (K3b counts it back to zero)
    uint32_t* cell_start = nullptr;  // [ncell+1] dense exclusive prefix
    uint32_t* side_list = nullptr;   // ids of side-list bodies (oversize / non-finite extent): tested against everyone
    uint32_t* far_list = nullptr;    // id|flags of normal-extent bodies outside the grid domain (finite centre)
    uint32_t* bnd_list = nullptr;    // id|flags of grid bodies within two cells of the domain boundary
    float4* far_c = nullptr;         // their collider centres, same slots
    float4* bnd_c = nullptr;
    uint8_t* klass = nullptr;        // 0 dropped (mass<=0), 1 grid, 2 side list, 3 lost (non-finite centre), 4 far
    uint32_t* row_start = nullptr;   // [N+1] dense: contacts with i == body are [row_start[b], row_start[b+1])
    uint32_t* col_start = nullptr;   // [N+1] dense: column-order slots with j == body
    uint32_t* body_cnt = nullptr;    // [N] per-body histogram scratch (self-cleaning)

    uint32_t* counters = nullptr;
    uint32_t* h_counters = nullptr;  // pinned mirror (detection read-back)
    uint32_t* h_rounds = nullptr;    // pinned: response wavefront count, fetched lazily
    bool rounds_pending = false;
    bool flow_pending = false;       // h_rounds holds the dataflow kernel's watchdog flag, not a round count

    // per contact
    unsigned long long* ckeys = nullptr;      // (i<<32)|j, unsorted then sorted
    unsigned long long* ckeys_alt = nullptr;
    unsigned long long* colkeys = nullptr;    // column scatter scratch: (i<<32)|c in arrival order
    uint32_t* colc = nullptr;                 // contact index per column slot
    uint32_t* col_pos = nullptr;              // inverse of colc
    uint32_t* dep = nullptr;
    uint32_t* succ_i = nullptr;
    uint32_t* succ_j = nullptr;
    uint4* crec = nullptr;                    // packed {i, j, succ_i, succ_j} per contact (dataflow executor)
    uint2* cver = nullptr;                    // versions of bodies i and j that contact c must see (dataflow executor)
    struct tr_contact_geo* cgeo = nullptr;    // per contact: normal, inverse masses, restitution, static bits (tr::ContactGeo)
    struct tr_body_rec* brec = nullptr;       // per body: version-checked velocity + position record (tr::BodyRec)
    uint32_t* frontier[2] = {nullptr, nullptr};

    // K2-K3b as one CUDA graph (memset + 4 kernels): launch parameters only change when the
    // buffers, the grid or the body count change, so the graph is re-captured only then
    cudaGraphExec_t bp_graph = nullptr;
    uint64_t bp_sig[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    uint64_t last_contacts = 0;
    uint64_t last_grid = 0;
    uint32_t last_side = 0;
    uint32_t last_far = 0;        // far / boundary list sizes of the previous step (launch sizing)
    uint32_t last_bnd = 0;
    int coop_blocks = 0;
    uint64_t prev_contacts = 0;   // last dataflow response: contacts and iterations of the longest worker
    uint32_t prev_rounds = 0;
};

namespace tr {

// ---------------------------------------------------------------------------------------
// exact-arithmetic helpers: same operations, same order as the reference, never contracted
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
    // dot(), main.cpp:195-198: (ax*bx + ay*by) + az*bz
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// checkSphereCollision, main.cpp:655-660
__device__ __forceinline__ bool sphere_hit(const float4 a, const float4 b) {
    float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    float d2 = dot3_rn(dx, dy, dz, dx, dy, dz);
    float rs = __fadd_rn(a.w, b.w);
    return d2 <= __fmul_rn(rs, rs);
}

// checkAABBCollision, main.cpp:648-652
__device__ __forceinline__ bool aabb_hit(const float4 ca, const float4 ha, const float4 cb, const float4 hb) {
    return fabsf(__fsub_rn(ca.x, cb.x)) <= __fadd_rn(ha.x, hb.x) && fabsf(__fsub_rn(ca.y, cb.y)) <= __fadd_rn(ha.y, hb.y) &&
           fabsf(__fsub_rn(ca.z, cb.z)) <= __fadd_rn(ha.z, hb.z);
}

__device__ __forceinline__ uint32_t pack_key(int cx, int cy, int cz, int bits) {
    uint32_t m = (1u << bits) - 1u;
    return ((uint32_t)cx & m) | (((uint32_t)cy & m) << bits) | (((uint32_t)cz & m) << (2 * bits));
}

struct GridParams {
    float ox, oy, oz, inv_cell, rmax_grid;
    int bits;
    int mixed;   // boxes live in the grid too: a body's extent is max(|radius| if sphere, |size.xyz|)
};

__device__ __forceinline__ uint32_t pack_val(uint32_t id, uint8_t f) {
    return id | ((f & F_STATIC) ? STATIC_BIT : 0u) | ((f & F_SPHERE) ? 0u : BOX_BIT);
}

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
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    {
        // stream the (zeroed) histogram into L2 ahead of the scattered REDs: one 128-byte line per
        // thread of the leading CTAs, so random cloud ids do not pay a DRAM round trip per atomic
        const uint32_t ncell = 1u << (3 * g.bits);
        if (i < ncell / 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(cell_count + (size_t)i * 32));
    }
    if (i < n) {
        const float4 c = center_r[i];
        const uint8_t f = flags[i];
        const uint32_t sentinel = 1u << (3 * g.bits);
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
                key = pack_key(__float2int_rd(ux), __float2int_rd(uy), __float2int_rd(uz), g.bits);
                k = 1;
                atomicAdd(&cell_count[key], 1u);   // result unused: a fire-and-forget RED
                // grid bodies within two cells of the domain boundary are the only ones a "far" body can touch
                if (fmaxf(fmaxf(fabsf(ux), fabsf(uy)), fabsf(uz)) >= 32766.0f) {
                    const uint32_t slot = atomicAdd(&counters[C_BND], 1u);
                    bnd_list[slot] = pack_val(i, f);
                    bnd_c[slot] = c;
                }
            } else if (fits && finite_c) {
                // "far": normal extent, finite centre, outside the grid domain.  It can only touch other
                // far bodies, boundary grid bodies (computed |u| differs by >= 2 cells from everyone else:
                // more than the largest admitted reach) and oversize side-list bodies.
                k = 4;
                const uint32_t slot = atomicAdd(&counters[C_FAR], 1u);
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
                side_list[slot] = i;
            }
        }
        keys[i] = key;
        klass[i] = k;
    }
}

// ---------------------------------------------------------------------------------------
// K3a: exclusive scan of the histogram in two unchained kernels (a look-back chain over
// ~1000 tiles costs more in latency than the second launch):
//   tile_sums_kernel   one CTA per tile of 2048 cells -> tile_sum[tile]
//   scan_kernel        each CTA adds up the sums of the tiles before it (<= 4 KB, L2 hits),
//                      scans its own tile and writes the dense cell_start table.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void load_tile_items(const uint32_t* __restrict__ cell_count, uint32_t base, uint32_t ncell,
                                                uint32_t (&v)[SCAN_ITEMS]) {
    if (base + SCAN_ITEMS <= ncell) {
        const uint4 a = *reinterpret_cast<const uint4*>(cell_count + base);
        const uint4 b = *reinterpret_cast<const uint4*>(cell_count + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < ncell) ? cell_count[base + k] : 0u;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) tile_sums_kernel(const uint32_t* __restrict__ cell_count, uint32_t ncell,
                                                                 uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    uint32_t v[SCAN_ITEMS];
    load_tile_items(cell_count, blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS, ncell, v);
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) tsum += v[k];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    tsum = __reduce_add_sync(0xFFFFFFFFu, tsum);
    if (lane == 0) s_warp[wid] = tsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < SCAN_THREADS / 32; ++k) t += s_warp[k];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const uint32_t* __restrict__ cell_count,
                                                            const uint32_t* __restrict__ tile_sum,
                                                            uint32_t* __restrict__ cell_start, uint32_t ncell,
                                                            uint32_t* counters, int total_slot) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32], s_red[SCAN_THREADS / 32], s_prefix;
    const uint32_t tile = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // prefix of this tile = sum of the tile sums before it
    uint32_t pre = 0;
    for (uint32_t k = threadIdx.x; k < tile; k += SCAN_THREADS) pre += tile_sum[k];
    pre = __reduce_add_sync(0xFFFFFFFFu, pre);
    if (lane == 0) s_red[wid] = pre;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    load_tile_items(cell_count, base, ncell, v);
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) tsum += v[k];
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t wv = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0u;
        uint32_t wi = wv;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= (uint32_t)o) wi += y;
        }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = wi - wv;            // exclusive warp offsets
        uint32_t r = lane < SCAN_THREADS / 32 ? s_red[lane] : 0u;
        r = __reduce_add_sync(0xFFFFFFFFu, r);
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, wi, SCAN_THREADS / 32 - 1);
        if (lane == 0) {
            s_prefix = r;
            if ((tile + 1) * (uint32_t)SCAN_TILE >= ncell) {   // last tile: grand total = number of grid bodies
                cell_start[ncell] = r + total;
                if (total_slot >= 0) counters[total_slot] = r + total;
            }
        }
    }
    __syncthreads();
    uint32_t ex = s_prefix + s_warp[wid] + (incl - tsum);
    uint32_t o[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        o[k] = ex;
        ex += v[k];
    }
    if (base + SCAN_ITEMS <= ncell) {
        *reinterpret_cast<uint4*>(cell_start + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(cell_start + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < ncell) cell_start[base + k] = o[k];
    }
}

// K3b: one record per grid body into its cell's segment; counting the histogram back down leaves it
// zeroed for the next step.  The record carries everything the narrow phase reads about a body -
// centre + radius, id|flags, AABB half extents - in ONE 32-byte sector written by one 256-bit store
// (SASS STG.E.ENL2.256) and read back by one 256-bit load, so nothing is gathered by id afterwards.
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
    return pack_key(__float2int_rd(ux), __float2int_rd(uy), __float2int_rd(uz), g.bits);
}

__global__ void __launch_bounds__(256) scatter_kernel(const uint32_t* __restrict__ keys, uint32_t sentinel,
                                                      const uint8_t* __restrict__ flags, const float4* __restrict__ center_r,
                                                      const float4* __restrict__ half_ext,   // NULL unless boxes are in the grid
                                                      uint32_t n, const uint32_t* __restrict__ cell_start,
                                                      uint32_t* __restrict__ cell_count, SlotRec* __restrict__ recs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = keys[i];
    if (key >= sentinel) return;          // side-list, far, lost or dropped body
    const float4 c = center_r[i];
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
    if (half_ext) h = half_ext[i];
    const uint32_t val = pack_val(i, flags[i]);
    const uint32_t old = atomicSub(&cell_count[key], 1u);
    st_rec(recs + (cell_start[key] + old - 1u), c, val, h);
}

// K4 (on demand): the stable (key,id) order.  The rank of a body inside its cell is the number of
// co-occupants with a smaller id.
__global__ void __launch_bounds__(256) canon_kernel(const SlotRec* __restrict__ recs, const uint32_t* __restrict__ cell_start,
                                                    GridParams g, const uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ sorted_ids) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= counters[C_GRID]) return;
    const uint32_t key = cell_key(recs[t].c, g);
    const uint32_t id = recs[t].val & ID_MASK;
    const uint32_t s = cell_start[key], e = cell_start[key + 1];
    uint32_t rank = 0;
    for (uint32_t u = s; u < e; ++u) rank += ((recs[u].val & ID_MASK) < id) ? 1u : 0u;
    sorted_ids[s + rank] = id;
}

// ---------------------------------------------------------------------------------------
// K5: narrow phase over the 27-cell stencil (9 runs of up to 3 key-contiguous cells)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_pair(unsigned long long* out, unsigned long long cap, unsigned long long* counter,
                                          uint32_t i, uint32_t j) {
    unsigned long long slot = atomicAdd(counter, 1ull);
    if (slot < cap) out[slot] = ((unsigned long long)i << 32) | j;
}

// Hits are staged per thread (NARROW_HITS partner ids in shared memory) and a warp reserves its
// output slots with ONE atomic at the end.  Emitting from inside the loops - even warp-aggregated, as
// ptxas does by itself - sends ~450K atomics to one address on the packed 1M scene, and same-address
// atomics retire at ~0.85 clk each at L2: that queue, not the loads, was the packed narrow phase
// (ncu: issue slots 15 % busy, L2 4 %, every warp waiting on the atomic's return).
constexpr int NARROW_THREADS = 128;
constexpr int NARROW_HITS = 8;

__device__ __noinline__ void emit_pair_overflow(unsigned long long* out, unsigned long long cap, unsigned long long* counter,
                                                uint32_t i, uint32_t j) {
    emit_pair(out, cap, counter, i, j);
}

template <bool CANDIDATES, bool MIXED>
__device__ __forceinline__ void test_one(uint32_t u, uint32_t va, uint32_t ida, const float4 ca, const float4 ha,
                                         const SlotRec* __restrict__ recs, uint32_t (*s_hit)[NARROW_THREADS], uint32_t& nh,
                                         unsigned long long* out, unsigned long long cap, unsigned long long* counter) {
    float4 cb, hb;
    uint32_t vb;
    ld_rec(recs + u, cb, vb, hb);
    const uint32_t idb = vb & ID_MASK;
    bool hit = true;
    if (!CANDIDATES) {
        if (MIXED && ((va | vb) & BOX_BIT)) hit = aabb_hit(ca, ha, cb, hb);   // main.cpp:704
        else hit = sphere_hit(ca, cb);                                       // main.cpp:702
        hit = hit && !(va & vb & STATIC_BIT);                                // main.cpp:695
    }
    if (hit) {
        if (nh < (uint32_t)NARROW_HITS) s_hit[nh][threadIdx.x] = idb;
        else emit_pair_overflow(out, cap, counter, min(ida, idb), max(ida, idb));   // staging full: straight out
        ++nh;
    }
}

// x-edge cell: the windows wrap around the row, so they are walked cell by cell (2 of 2^bits cells per row)
template <bool CANDIDATES, bool MIXED>
__device__ __noinline__ void narrow_edge(uint32_t t, uint32_t key, int bits, uint32_t va, const float4 ca, const float4 ha,
                                         const SlotRec* __restrict__ recs, const uint32_t* __restrict__ cell_start,
                                         uint32_t (*s_hit)[NARROW_THREADS], uint32_t& nh, unsigned long long* out,
                                         unsigned long long cap, unsigned long long* counter) {
    const uint32_t ida = va & ID_MASK;
    const uint32_t msk = (1u << bits) - 1u;
    const uint32_t cx = key & msk, cy = (key >> bits) & msk, cz = (key >> (2 * bits)) & msk;
    // window w: 0 = rest of the own cell, 1 = cell cx+1 of the own row, 2.. = 3 cells of each forward row
#pragma unroll 1
    for (int w = 0; w < 14; ++w) {
        uint32_t s, e;
        if (w == 0) {
            s = t + 1;
            e = cell_start[key + 1];
        } else {
            uint32_t nk;
            if (w == 1) {
                nk = (key & ~msk) | ((cx + 1) & msk);
            } else {
                const int r = (w - 2) / 3, dx = (w - 2) % 3 - 1;
                const int dz = r == 0 ? 0 : 1;
                const int dy = r == 0 ? 1 : r - 2;
                nk = (((cy + dy) & msk) << bits) | (((cz + dz) & msk) << (2 * bits)) | ((cx + dx) & msk);
            }
            s = cell_start[nk];
            e = cell_start[nk + 1];
        }
#pragma unroll 1
        for (uint32_t u = s; u < e; ++u) test_one<CANDIDATES, MIXED>(u, va, ida, ca, ha, recs, s_hit, nh, out, cap, counter);
    }
}

// The candidate relation "cell(b) is in the 27-cell stencil of cell(a)" is symmetric, so every
// unordered pair is reached from exactly one side by a HALF stencil: the later slots of a's own
// cell, and the 13 cells at "forward" offsets (dz,dy,dx) > (0,0,0) lexicographically (an offset o
// and its opposite -o are distinct modulo 2^bits >= 4, and exactly one of them is forward).
// With x as the fastest key digit those are 5 contiguous slot ranges for an interior cell:
//   [t+1, end of cell cx+1]  in the own row, and cells cx-1..cx+1 of the rows (dz,dy) =
//   (0,+1), (+1,-1), (+1,0), (+1,+1).
// "Later slots of the own cell" is any fixed order: the arrival order K3b left is as good as the id order.
// MIXED: boxes live in the grid; a pair with at least one box uses the AABB predicate on the
// collider sizes of both bodies, exactly as the reference dispatches (main.cpp:701-706).
// One thread per slot.  (A warp-cooperative version - the union of a warp's windows per stencil
// row staged in shared memory by coalesced loads, each record read once per warp - was measured at
// 98.9 us sparse / 286.7 us packed against 54.6 / 321.8 for this one: the staging phases serialise
// a warp that otherwise hides its latency with 64 independent threads per scheduler.)
template <bool CANDIDATES, bool MIXED>
__global__ void __launch_bounds__(NARROW_THREADS) narrow_kernel(const SlotRec* __restrict__ recs,
                                                                const uint32_t* __restrict__ cell_start,
                                                                const uint32_t* __restrict__ counters, GridParams g,
                                                                unsigned long long* out, unsigned long long cap,
                                                                unsigned long long* counter) {
    __shared__ uint32_t s_hit[NARROW_HITS][NARROW_THREADS];
    const uint32_t t = blockIdx.x * NARROW_THREADS + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t nh = 0, ida = 0;
    if (t < counters[C_GRID]) {
        float4 ca, ha;
        uint32_t va;
        ld_rec(recs + t, ca, va, ha);
        const uint32_t key = cell_key(ca, g);
        ida = va & ID_MASK;
        const int bits = g.bits;
        const uint32_t msk = (1u << bits) - 1u;
        const uint32_t cx = key & msk, cy = (key >> bits) & msk, cz = (key >> (2 * bits)) & msk;
        if (cx > 0 && cx < msk) {
            // interior cell: five contiguous slot ranges, all nine look-ups issued before the first use
            uint32_t rs[5], re[5];
            rs[0] = t + 1;
            re[0] = cell_start[key + 2];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int dz = r == 0 ? 0 : 1;
                const int dy = r == 0 ? 1 : r - 2;
                const uint32_t row = (((cy + dy) & msk) << bits) | (((cz + dz) & msk) << (2 * bits));
                rs[r + 1] = cell_start[row + cx - 1];
                re[r + 1] = cell_start[row + cx + 2];
            }
            // (one flattened loop over the five windows - a lane moving on as soon as its own window is
            // exhausted - was slower: 77 / 90 us against 62 / 73; the compiler pipelines these loops better)
#pragma unroll
            for (int r = 0; r < 5; ++r)
                for (uint32_t u = rs[r]; u < re[r]; ++u)
                    test_one<CANDIDATES, MIXED>(u, va, ida, ca, ha, recs, s_hit, nh, out, cap, counter);
        } else {
            narrow_edge<CANDIDATES, MIXED>(t, key, bits, va, ca, ha, recs, cell_start, s_hit, nh, out, cap, counter);
        }
    }
    // one reservation per warp for everything it staged
    if (!__any_sync(0xFFFFFFFFu, nh != 0)) return;
    const uint32_t cnt = min(nh, (uint32_t)NARROW_HITS);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0) + (incl - cnt);
    for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t idb = s_hit[k][threadIdx.x];
        if (base + k < cap) out[base + k] = ((unsigned long long)min(ida, idb) << 32) | max(ida, idb);
    }
}

// K5b: side-list bodies against every live body.  grid = (ceil(n/256), up to 64).
__global__ void __launch_bounds__(256) side_kernel(const uint32_t* __restrict__ side_list, const uint32_t* __restrict__ counters,
                                                   const float4* __restrict__ center_r,
                                                   const float4* __restrict__ half_ext,
                                                   const uint8_t* __restrict__ flags, const uint8_t* __restrict__ klass,
                                                   uint32_t n, unsigned long long* out, unsigned long long cap,
                                                   unsigned long long* counter) {
    const uint32_t side_count = counters[C_SIDE];
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const uint8_t kx = klass[x];
    if (kx == 0) return;
    const float4 cx = center_r[x], hx = half_ext[x];
    const uint8_t fx = flags[x];
    for (uint32_t si = blockIdx.y; si < side_count; si += gridDim.y) {
        const uint32_t s = side_list[si];
        if (s == x) continue;
        if (kx == 2 && x < s) continue;   // side-side pairs are emitted once, from the lower id's row
        const uint8_t fs = flags[s];
        if ((fx & F_STATIC) && (fs & F_STATIC)) continue;
        const float4 cs = center_r[s];
        bool hit;
        if ((fx & F_SPHERE) && (fs & F_SPHERE)) hit = sphere_hit(cx, cs);
        else hit = aabb_hit(cx, hx, cs, half_ext[s]);
        if (hit) emit_pair(out, cap, counter, min(x, s), max(x, s));
    }
}

// K5c: "far" bodies (normal extent, finite centre, outside the grid domain |u| < 32768) against each
// other and against the grid bodies within two cells of the domain boundary; side_kernel covers
// far-vs-oversize.  A contact implies |delta u| < 1 per axis (cell = 2*rmax*(1+2^-7)) and computed
// u carries < 0.01 of error at that magnitude, so a grid body that touches a far body has
// |u| > 32766 on that axis: nobody else needs testing.  Thread = far body (grid-stride), rows
// (gridDim.y) stride over the partner list, so any launch geometry is correct.
// Both lists carry id|flags (pack_val) and a compact copy of the centres, so the partner loop
// streams consecutive 16-byte entries instead of gathering by id.
__global__ void __launch_bounds__(256) far_kernel(const uint32_t* __restrict__ far_list, const float4* __restrict__ far_c,
                                                  const uint32_t* __restrict__ bnd_list, const float4* __restrict__ bnd_c,
                                                  const uint32_t* __restrict__ counters, const float4* __restrict__ half_ext,
                                                  unsigned long long* out, unsigned long long cap, unsigned long long* counter) {
    const uint32_t nfar = counters[C_FAR], nbnd = counters[C_BND];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nfar; t += gridDim.x * blockDim.x) {
        const uint32_t va = far_list[t];
        const uint32_t a = va & ID_MASK;
        const float4 ca = far_c[t];
        // partners: far bodies later in the list, then every boundary grid body
        for (uint32_t k = t + 1 + blockIdx.y; k < nfar + nbnd; k += gridDim.y) {
            const bool kf = k < nfar;
            const uint32_t vb = kf ? far_list[k] : bnd_list[k - nfar];
            if (va & vb & STATIC_BIT) continue;                      // main.cpp:695
            const float4 cb = kf ? far_c[k] : bnd_c[k - nfar];
            const uint32_t b = vb & ID_MASK;
            bool hit;
            if ((va | vb) & BOX_BIT) hit = aabb_hit(ca, half_ext[a], cb, half_ext[b]);   // main.cpp:704
            else hit = sphere_hit(ca, cb);                                               // main.cpp:702
            if (hit) emit_pair(out, cap, counter, min(a, b), max(a, b));
        }
    }
}

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
    if (last_i) {                                        // nobody reads the record again: back to the SoA state
        a.vel_out[i] = make_float4(vi.x, vi.y, vi.z, a.vel_rest[i].w);
        a.pos_out[i] = make_float4(pi.x, pi.y, pi.z, a.pos_m[i].w);
    } else {
        st_body(a.brec + i, vi, pi, want.x + 1u);
    }
    if (last_j) {
        a.vel_out[j] = make_float4(vj.x, vj.y, vj.z, a.vel_rest[j].w);
        a.pos_out[j] = make_float4(pj.x, pj.y, pj.z, a.pos_m[j].w);
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
        if (work && !respond_rec(a, cur, rec.x, rec.y, s1 == EMPTY, s2 == EMPTY, counters)) alive = false;   // watchdog
        if (LANES == 1 && !work) __nanosleep(idle_ns);   // an idle worker that polls flat out slows the busy ones down
        // ---- phase 2 (atomics): working lanes release their successors, idle lanes try to acquire
        if (work) {
            const uint32_t r1 = (s1 != EMPTY) ? atomicSub(&dep[s1], 1u) : 0u;
            const uint32_t r2 = (s2 != EMPTY) ? atomicSub(&dep[s2], 1u) : 0u;
            if (s1 != EMPTY) {                    // the next hand-off will touch these bodies
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n1.x));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n1.y));
            }
            if (s2 != EMPTY) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n2.x));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.brec + n2.y));
            }
            ++mine;
            uint32_t next = EMPTY;
            if (r1 == 1u) {
                next = s1;
                rec = n1;
            }
            if (r2 == 1u) {
                if (next == EMPTY) {
                    next = s2;
                    rec = n2;
                } else {
                    vq[atomicAdd(&counters[C_QTAIL], 1u)] = s2;   // second ready successor: to the global queue
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

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <typename T>
static int realloc_dev(T** p, uint64_t count) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (count == 0) return TR_OK;
    TR_CUDA(cudaMalloc(p, count * sizeof(T)));
    return TR_OK;
}

static Collide* state(tr_world* w) {
    if (!w->col) w->col = new Collide();
    return w->col;
}

// Host-side choice of the grid: which bodies are admitted, the cell edge, bits per axis.
// A body's extent is |radius| for a sphere in an all-sphere world; as soon as one live box with
// finite size exists ("mixed" world) any body may meet a box, box pairs use collider::size of both
// bodies, and the extent becomes max(|radius| if sphere, |size.x|, |size.y|, |size.z|).
static GridParams grid_params(const Collide* c) {
    return GridParams{c->origin[0], c->origin[1], c->origin[2], c->inv_cell, c->rmax_grid, c->bits, c->mixed ? 1 : 0};
}

static float body_extent(const tr_body_desc& b, bool mixed) {
    float e = b.isSphere ? std::fabs(b.radius) : 0.f;
    if (mixed || !b.isSphere) e = std::max(e, std::max(std::fabs(b.size[0]), std::max(std::fabs(b.size[1]), std::fabs(b.size[2]))));
    return e;
}

int collide_classify(tr_world* w) {
    Collide* c = state(w);
    const tr_world_params& p = w->params;
    bool mixed = false;
    for (const tr_body_desc& b : w->host_bodies)
        if (!(b.mass <= 0) && !b.isSphere && std::isfinite(b.size[0]) && std::isfinite(b.size[1]) && std::isfinite(b.size[2])) {
            mixed = true;
            break;
        }
    c->mixed = mixed;
    // admitted extent: user value, else the largest extent <= 8 * median extent
    float rmax = p.grid_max_radius;
    if (rmax <= 0.f) {
        std::vector<float> r;
        r.reserve(w->host_bodies.size());
        for (const tr_body_desc& b : w->host_bodies) {
            if (b.mass <= 0 || (!mixed && !b.isSphere)) continue;
            float e = body_extent(b, mixed);
            if (std::isfinite(e)) r.push_back(e);
        }
        rmax = 0.f;
        if (!r.empty()) {
            std::vector<float> s = r;
            std::nth_element(s.begin(), s.begin() + s.size() / 2, s.end());
            float limit = 8.0f * s[s.size() / 2];
            for (float v : r)
                if (v <= limit && v > rmax) rmax = v;
            if (limit == 0.f) rmax = 0.f;
        }
    }
    c->rmax_grid = rmax;
    // can any body be outside the grid for a reason known at creation time?
    c->may_have_side = false;
    for (const tr_body_desc& b : w->host_bodies) {
        if (b.mass <= 0) continue;
        if ((!mixed && !b.isSphere) || !(body_extent(b, mixed) <= rmax)) {
            c->may_have_side = true;
            break;
        }
    }
    float cell = p.grid_cell;
    if (cell <= 0.f) {
        cell = (2.0f * rmax) * 1.0078125f;   // 2*rmax*(1+2^-7): margin analysed in DESIGN.md
        if (!(cell > 0.f) || !std::isfinite(cell)) cell = 1.0f;
    }
    c->cell = cell;
    c->inv_cell = 1.0f / cell;
    memcpy(c->origin, p.grid_origin, sizeof(c->origin));
    int bits = p.grid_bits;
    if (bits == 0) {
        uint64_t want = 2 * (w->n ? w->n : 1);
        bits = 2;
        while (bits < 10 && (1ull << (3 * bits)) < want) ++bits;
    }
    c->bits = bits;
    w->classified = true;
    return TR_OK;
}

static int ensure_buffers(tr_world* w) {
    Collide* c = state(w);
    const uint64_t n = w->n;
    if (n > c->cap_bodies) {
        uint64_t cap = w->cap >= n ? w->cap : n;
        TR_TRY(realloc_dev(&c->keys, cap));
        TR_TRY(realloc_dev((SlotRec**)&c->recs, cap));
        TR_TRY(realloc_dev(&c->sorted_ids, cap));
        TR_TRY(realloc_dev((BodyRec**)&c->brec, cap));
        TR_TRY(realloc_dev(&c->side_list, cap));
        TR_TRY(realloc_dev(&c->far_list, cap));
        TR_TRY(realloc_dev(&c->bnd_list, cap));
        TR_TRY(realloc_dev(&c->far_c, cap));
        TR_TRY(realloc_dev(&c->bnd_c, cap));
        TR_TRY(realloc_dev(&c->klass, cap));
        TR_TRY(realloc_dev(&c->row_start, cap + 1));
        TR_TRY(realloc_dev(&c->col_start, cap + 1));
        TR_TRY(realloc_dev(&c->body_cnt, cap));
        TR_CUDA(cudaMemsetAsync(c->body_cnt, 0, cap * sizeof(uint32_t), w->stream));
        c->cap_bodies = cap;
    }
    const uint64_t ncell = 1ull << (3 * c->bits);
    if (ncell > c->ncell_alloc) {
        TR_TRY(realloc_dev(&c->cell_start, ncell + 1));
        TR_TRY(realloc_dev(&c->cell_count, ncell));
        // zeroed ONCE: the histogram counts itself back to zero every step (scatter_kernel)
        TR_CUDA(cudaMemsetAsync(c->cell_count, 0, ncell * sizeof(uint32_t), w->stream));
        c->ncell_alloc = ncell;
    }
    {
        const uint64_t longest = ncell > c->cap_bodies + 1 ? ncell : c->cap_bodies + 1;
        const uint64_t tiles = (longest + SCAN_TILE - 1) / SCAN_TILE;
        if (tiles > c->tile_sum_cap) {
            TR_TRY(realloc_dev(&c->tile_sum, tiles));
            c->tile_sum_cap = tiles;
        }
    }
    uint64_t want_c = w->params.max_contacts ? w->params.max_contacts : 8 * n + 4096;
    if (want_c > 0xFFFFFFF0ull) want_c = 0xFFFFFFF0ull;
    if (want_c != c->cap_contacts) {
        TR_TRY(realloc_dev(&c->ckeys, want_c));
        TR_TRY(realloc_dev(&c->ckeys_alt, want_c));
        TR_TRY(realloc_dev(&c->colkeys, want_c));
        TR_TRY(realloc_dev(&c->colc, want_c));
        TR_TRY(realloc_dev(&c->col_pos, want_c));
        TR_TRY(realloc_dev(&c->dep, want_c));
        TR_TRY(realloc_dev(&c->succ_i, want_c));
        TR_TRY(realloc_dev(&c->succ_j, want_c));
        TR_TRY(realloc_dev(&c->crec, want_c));
        TR_TRY(realloc_dev(&c->cver, want_c));
        TR_TRY(realloc_dev((ContactGeo**)&c->cgeo, want_c));
        TR_TRY(realloc_dev(&c->frontier[0], want_c));
        TR_TRY(realloc_dev(&c->frontier[1], want_c));
        c->cap_contacts = want_c;
    }
    if (!c->counters) {
        TR_CUDA(cudaMalloc(&c->counters, (C_NCOUNTERS + 32) * sizeof(uint32_t)));   // + statistics of TR_FLOW_STATS builds
        TR_CUDA(cudaMemset(c->counters, 0, (C_NCOUNTERS + 32) * sizeof(uint32_t)));
        TR_CUDA(cudaMallocHost(&c->h_counters, C_NHOST * sizeof(uint32_t)));
        TR_CUDA(cudaMallocHost(&c->h_rounds, 8 * sizeof(uint32_t)));
        memset(c->h_rounds, 0, 8 * sizeof(uint32_t));
    }
    return TR_OK;
}

// K2-K3b.  `ev` (debug, TR_BP_PROFILE=1): 5 events recorded around the four kernels.
static int launch_broad(tr_world* w, cudaStream_t st, cudaEvent_t* ev = nullptr) {
    Collide* c = state(w);
    const uint32_t n = (uint32_t)w->n;
    const uint32_t blocks = (n + 255) / 256;
    const uint32_t ncell = 1u << (3 * c->bits);
    const uint32_t ntiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
    const GridParams g = grid_params(c);
    TR_CUDA(cudaMemsetAsync(c->counters, 0, C_NCOUNTERS * sizeof(uint32_t), st));
    if (ev) cudaEventRecord(ev[0], st);
    keys_kernel<<<blocks, 256, 0, st>>>(w->center_r, w->half_ext, w->flags, n, g, c->keys, c->klass, c->side_list, c->far_list, c->far_c,
                                        c->bnd_list, c->bnd_c, c->cell_count, c->counters);
    if (ev) cudaEventRecord(ev[1], st);
    tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(c->cell_count, ncell, c->tile_sum);
    if (ev) cudaEventRecord(ev[2], st);
    scan_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(c->cell_count, c->tile_sum, c->cell_start, ncell, c->counters, C_GRID);
    if (ev) cudaEventRecord(ev[3], st);
    scatter_kernel<<<blocks, 256, 0, st>>>(c->keys, ncell, w->flags, w->center_r, c->mixed ? w->half_ext : nullptr, n, c->cell_start,
                                           c->cell_count, (SlotRec*)c->recs);
    if (ev) cudaEventRecord(ev[4], st);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

// Broad phase K2-K3b.  Leaves counters[C_SIDE] / [C_FAR] / [C_BND] / [C_GRID] on the device; no host sync.
static int broad_phase(tr_world* w) {
    Collide* c = state(w);
    const uint32_t n = (uint32_t)w->n;
    PendingTimer t;
    static const bool bp_profile = getenv("TR_BP_PROFILE") != nullptr;
    if (bp_profile) {
        // debug: plain launches with an event between the kernels, per-kernel times to stderr
        // (event gaps include the launch gap a graph does not have: ~2 us per kernel)
        cudaEvent_t ev[5];
        for (auto& e : ev) TR_CUDA(cudaEventCreate(&e));
        timer_begin(w, T_BROAD, &t);
        TR_TRY(launch_broad(w, w->stream, ev));
        timer_end(w, &t);
        TR_CUDA(cudaEventSynchronize(ev[4]));
        float ms[4];
        for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
        fprintf(stderr, "[TR_BP_PROFILE] n=%u keys %.1f us, tile sums %.1f us, scan %.1f us, scatter %.1f us\n", n, ms[0] * 1e3f,
                ms[1] * 1e3f, ms[2] * 1e3f, ms[3] * 1e3f);
        for (auto& e : ev) cudaEventDestroy(e);
    } else if (w->stream == nullptr) {
        // the legacy default stream cannot be captured: plain launches
        timer_begin(w, T_BROAD, &t);
        TR_TRY(launch_broad(w, w->stream));
        timer_end(w, &t);
    } else {
        // signature of everything the captured launches depend on
        uint64_t sig[8] = {(uint64_t)(uintptr_t)w->center_r, (uint64_t)(uintptr_t)w->flags, (uint64_t)(uintptr_t)c->keys,
                           (uint64_t)(uintptr_t)c->cell_count, (uint64_t)(uintptr_t)c->ckeys ^ (uint64_t)(uintptr_t)c->recs,
                           ((uint64_t)n << 9) | ((uint64_t)(c->mixed ? 1 : 0) << 8) | (uint64_t)c->bits, 0, 0};
        memcpy(&sig[6], &c->inv_cell, 4);
        memcpy((char*)&sig[6] + 4, &c->rmax_grid, 4);
        memcpy(&sig[7], &c->origin[0], 8);
        uint32_t oz = 0;
        memcpy(&oz, &c->origin[2], 4);
        sig[7] ^= (uint64_t)oz << 13;
        if (!c->bp_graph || memcmp(sig, c->bp_sig, sizeof(sig)) != 0) {
            if (c->bp_graph) cudaGraphExecDestroy(c->bp_graph);
            c->bp_graph = nullptr;
            cudaGraph_t graph = nullptr;
            TR_CUDA(cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal));
            int rc = launch_broad(w, w->stream);
            cudaError_t e = cudaStreamEndCapture(w->stream, &graph);
            if (rc != TR_OK) return rc;
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
            e = cudaGraphInstantiate(&c->bp_graph, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
            memcpy(c->bp_sig, sig, sizeof(sig));
        }
        timer_begin(w, T_BROAD, &t);
        TR_CUDA(cudaGraphLaunch(c->bp_graph, w->stream));
        timer_end(w, &t);
    }
    w->stats.kernel_launches += 4;
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

static int read_counters(tr_world* w) {
    Collide* c = state(w);
    TR_CUDA(cudaMemcpyAsync(c->h_counters, c->counters, C_NHOST * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    TR_CUDA(cudaStreamSynchronize(w->stream));
    return TR_OK;
}

static int launch_side(tr_world* w, unsigned long long* out, uint64_t cap, unsigned long long* counter) {
    Collide* c = state(w);
    // rows of the launch loop over the side list with stride gridDim.y, so any row count is correct;
    // last step's side count is the best guess for this one
    uint32_t rows = c->last_side ? c->last_side : 1;
    if (rows > 64) rows = 64;
    dim3 grid((uint32_t)((w->n + 255) / 256), rows);
    side_kernel<<<grid, 256, 0, w->stream>>>(c->side_list, c->counters, w->center_r, w->half_ext, w->flags, c->klass,
                                             (uint32_t)w->n, out, cap, counter);
    TR_CUDA(cudaGetLastError());
    w->stats.kernel_launches += 1;
    return TR_OK;
}

static int launch_far(tr_world* w, unsigned long long* out, uint64_t cap, unsigned long long* counter) {
    Collide* c = state(w);
    // sized from the previous step's counts; the kernel strides, so any geometry is correct
    uint32_t gx = (c->last_far + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > 1024) gx = 1024;
    uint32_t gy = (c->last_far + c->last_bnd + 255) / 256;
    if (gy < 1) gy = 1;
    if (gy > 64) gy = 64;
    far_kernel<<<dim3(gx, gy), 256, 0, w->stream>>>(c->far_list, c->far_c, c->bnd_list, c->bnd_c, c->counters, w->half_ext, out, cap,
                                                    counter);
    TR_CUDA(cudaGetLastError());
    w->stats.kernel_launches += 1;
    return TR_OK;
}

// Detection: broad phase + narrow phase + side list, ONE host read-back at the end.  On return
// the UNSORTED pairs are in `out` and *count is their number (may exceed cap => overflow).
template <bool CANDIDATES>
static int detect(tr_world* w, unsigned long long* out, uint64_t cap, uint64_t* count) {
    Collide* c = state(w);
    TR_TRY(broad_phase(w));
    unsigned long long* counter = (unsigned long long*)(c->counters + (CANDIDATES ? C_CAND : C_CONTACTS));
    PendingTimer t;
    timer_begin(w, T_NARROW, &t);
    const uint32_t nb = (uint32_t)((w->n + NARROW_THREADS - 1) / NARROW_THREADS);
    const GridParams g = grid_params(c);
    if (c->mixed)
        narrow_kernel<CANDIDATES, true><<<nb, NARROW_THREADS, 0, w->stream>>>((const SlotRec*)c->recs, c->cell_start, c->counters, g, out, cap, counter);
    else
        narrow_kernel<CANDIDATES, false><<<nb, NARROW_THREADS, 0, w->stream>>>((const SlotRec*)c->recs, c->cell_start, c->counters, g, out, cap, counter);
    w->stats.kernel_launches += 1;
    const bool side_launched = c->may_have_side && !CANDIDATES;
    if (side_launched) TR_TRY(launch_side(w, out, cap, counter));
    const bool far_launched = c->last_far > 0 && !CANDIDATES;
    if (far_launched) TR_TRY(launch_far(w, out, cap, counter));
    timer_end(w, &t);
    TR_CUDA(cudaGetLastError());
    TR_TRY(read_counters(w));
    bool again = false;
    if (!CANDIDATES && !side_launched && c->h_counters[C_SIDE] > 0) {
        // a body's extent became oversize / non-finite at run time: test it now (rare path, one more read-back)
        c->may_have_side = true;
        TR_TRY(launch_side(w, out, cap, counter));
        again = true;
    }
    if (!CANDIDATES && !far_launched && c->h_counters[C_FAR] > 0) {
        // bodies left the grid domain at run time: first step with a far list
        c->last_far = c->h_counters[C_FAR];
        c->last_bnd = c->h_counters[C_BND];
        TR_TRY(launch_far(w, out, cap, counter));
        again = true;
    }
    if (again) TR_TRY(read_counters(w));
    c->last_side = c->h_counters[C_SIDE];
    c->last_far = c->h_counters[C_FAR];
    c->last_bnd = c->h_counters[C_BND];
    c->last_grid = c->h_counters[C_GRID];
    const int k = CANDIDATES ? C_CAND : C_CONTACTS;
    *count = ((uint64_t)c->h_counters[k + 1] << 32) | c->h_counters[k];
    return TR_OK;
}

void collide_poll_stats(tr_world* w) {
    Collide* c = w->col;
    if (c && c->rounds_pending) {
        if (c->flow_pending) {
            w->stats.last_response_rounds = c->h_rounds[0];   // dataflow: lock-step iterations of the longest warp
            if (c->h_rounds[C_ABORT - C_ROUNDS] != 0 && w->deferred_error == TR_OK) {
                w->deferred_error = TR_ERR_CUDA;
                w->deferred_msg = "ordered response: dataflow watchdog fired (dependency graph did not drain)";
            }
            c->flow_pending = false;
        } else {
            w->stats.last_response_rounds = *c->h_rounds;
        }
        c->rounds_pending = false;
    }
}

int collide_step(tr_world* w) {
    if (w->n == 0) return TR_OK;
    if (!w->classified) TR_TRY(collide_classify(w));
    TR_TRY(ensure_buffers(w));
    Collide* c = state(w);
    uint64_t count = 0;
    TR_TRY(detect<false>(w, c->ckeys, c->cap_contacts, &count));
    // detect() synchronised with the host: the previous response's iteration count has arrived
    if (c->rounds_pending && c->flow_pending) {
        c->prev_rounds = c->h_rounds[0];
        c->prev_contacts = c->last_contacts;
    }
    collide_poll_stats(w);
    w->stats.side_list_size = c->last_side + c->last_far;   // everything outside the grid: oversize list + far list
    w->stats.last_contacts = count;
    c->last_contacts = 0;
    if (count > c->cap_contacts) {
        set_error("contact buffer overflow: %llu contacts, capacity %llu (raise tr_world_params.max_contacts)",
                  (unsigned long long)count, (unsigned long long)c->cap_contacts);
        return TR_ERR_CAPACITY;
    }
    c->rounds_pending = false;
    w->stats.last_response_rounds = 0;
    if (count == 0) return TR_OK;
    const uint32_t cn = (uint32_t)count;
    const uint32_t cblocks = (cn + 255) / 256;
    PendingTimer t;
    timer_begin(w, T_RESPONSE, &t);
    const uint32_t nb = (uint32_t)w->n;
    const uint32_t btiles = (nb + SCAN_TILE - 1) / SCAN_TILE;
    // lexicographic (i,j) order == the reference's visiting order: counting sort by i, rank by j
    chist_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys, cn, 32, c->body_cnt);
    tile_sums_kernel<<<btiles, SCAN_THREADS, 0, w->stream>>>(c->body_cnt, nb, c->tile_sum);
    scan_kernel<<<btiles, SCAN_THREADS, 0, w->stream>>>(c->body_cnt, c->tile_sum, c->row_start, nb, c->counters, -1);
    cscatter_row_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys, cn, c->row_start, c->body_cnt, c->ckeys_alt);
    crank_row_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys_alt, c->row_start, cn, c->ckeys);
    // column order (j, then i): the same, keyed by j
    chist_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys, cn, 0, c->body_cnt);
    tile_sums_kernel<<<btiles, SCAN_THREADS, 0, w->stream>>>(c->body_cnt, nb, c->tile_sum);
    scan_kernel<<<btiles, SCAN_THREADS, 0, w->stream>>>(c->body_cnt, c->tile_sum, c->col_start, nb, c->counters, -1);
    cscatter_col_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys, cn, c->col_start, c->body_cnt, c->colkeys);
    crank_col_kernel<<<cblocks, 256, 0, w->stream>>>(c->colkeys, c->ckeys, c->col_start, cn, c->colc, c->col_pos);
    const bool flow = w->params.response_mode != 1;   // 1 selects the level-synchronous kernel
    if (flow) TR_CUDA(cudaMemsetAsync(c->frontier[0], 0xFF, (size_t)cn * sizeof(uint32_t), w->stream));   // queue = all EMPTY
    deps_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys, c->colc, c->col_pos, c->row_start, c->col_start, cn, c->dep, c->succ_i,
                                                c->succ_j, c->crec, c->frontier[0], c->counters, flow ? C_QTAIL : C_FRONT, w->pos_m[w->cur],
                                                w->vel_rest, w->center_r, w->half_ext, w->flags, flow ? (BodyRec*)c->brec : nullptr, c->cver,
                                                (ContactGeo*)c->cgeo);
    if (flow) {
        FlowArgs ra{w->pos_m[w->cur], w->vel_rest, w->pos_m[w->cur], w->vel_rest, (BodyRec*)c->brec, c->cver, (const ContactGeo*)c->cgeo};
        int fblocks = w->num_sms;                     // resident by construction: 128 threads per SM
        int fwant = (int)((cn + 127) / 128);
        if (fwant < fblocks) fblocks = fwant < 1 ? 1 : fwant;
        // Executor width.  Deep graphs (a packed lattice: ~800 dependent hand-offs) are latency bound:
        // single-lane workers, 6 CTAs of 4 warps per SM (3, 4, 5, 7 per SM: 2.47, 2.08, 1.90, 1.82 ms against 1.78).
        // Wide shallow graphs are throughput bound: whole warps.  The previous step tells which:
        // contacts per lock-step iteration of its longest worker.
        bool wide = w->params.response_mode == 2;
        if (w->params.response_mode == 0 && c->prev_rounds > 0) wide = c->prev_contacts / c->prev_rounds > 16384;
        if (wide) {
            response_flow_kernel<32><<<fblocks, 128, 0, w->stream>>>(ra, c->crec, c->dep, c->frontier[0], cn, c->counters, 0u);
        } else {
            static const int per_sm = getenv("TR_FLOW_CTAS") ? atoi(getenv("TR_FLOW_CTAS")) : 6;
            int sblocks = w->num_sms * per_sm;
            const int swant = (int)((cn + 3) / 4);
            if (swant < sblocks) sblocks = swant < 1 ? 1 : swant;
            // idle back-off: 300 and 700 ns are equivalent (1.78 / 1.80 ms), 100 ns and 2000 ns are 10 % slower
            response_flow_kernel<1><<<sblocks, 128, 0, w->stream>>>(ra, c->crec, c->dep, c->frontier[0], cn, c->counters, 500u);
        }
        TR_CUDA(cudaGetLastError());
        timer_end(w, &t);
        w->stats.kernel_launches += 12;   // 2 x (hist, tile sums, scan, scatter, rank) + deps + response
        c->last_contacts = count;
        // watchdog flag + (unused) round counter come back with the next sync
        TR_CUDA(cudaMemcpyAsync(c->h_rounds, c->counters + C_ROUNDS, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
        c->rounds_pending = true;
        c->flow_pending = true;
        return TR_OK;
    }
    if (c->coop_blocks == 0) {
        int per_sm = 0;
        TR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, response_kernel, 256, 0));
        if (per_sm < 1) per_sm = 1;
        c->coop_blocks = w->num_sms;   // one CTA per SM: the grid barrier cost grows with the CTA count
    }
    RespArgs ra{w->pos_m[w->cur], w->vel_rest, w->center_r, w->half_ext, w->flags};
    const unsigned long long* ck = c->ckeys;
    uint32_t* dep = c->dep;
    const uint32_t* si = c->succ_i;
    const uint32_t* sj = c->succ_j;
    uint32_t* f0 = c->frontier[0];
    uint32_t* f1 = c->frontier[1];
    uint32_t* cnt = c->counters;
    void* args[] = {&ra, &ck, &dep, &si, &sj, &f0, &f1, &cnt};
    int blocks = c->coop_blocks;
    int want = (int)((cn + 255) / 256);
    if (want < blocks) blocks = want < 1 ? 1 : want;
    TR_CUDA(cudaLaunchCooperativeKernel((void*)response_kernel, dim3(blocks), dim3(256), args, 0, w->stream));
    timer_end(w, &t);
    w->stats.kernel_launches += 12;   // 2 x (hist, tile sums, scan, scatter, rank) + deps + response
    c->last_contacts = count;
    // wavefront count: copied to pinned memory asynchronously, picked up at the next sync (no extra stall)
    TR_CUDA(cudaMemcpyAsync(c->h_rounds, c->counters + C_ROUNDS, sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    c->rounds_pending = true;
    return TR_OK;
}

int collide_read_contacts(tr_world* w, int32_t* pairs, uint64_t cap, uint64_t* n) {
    Collide* c = w->col;
    uint64_t cnt = c ? c->last_contacts : 0;
    if (n) *n = cnt;
    if (!pairs || cnt == 0) return TR_OK;
    uint64_t take = cnt < cap ? cnt : cap;
    std::vector<unsigned long long> h(take);
    TR_CUDA(cudaMemcpyAsync(h.data(), c->ckeys, take * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    TR_CUDA(cudaStreamSynchronize(w->stream));
    for (uint64_t k = 0; k < take; ++k) {
        pairs[2 * k] = (int32_t)(h[k] >> 32);
        pairs[2 * k + 1] = (int32_t)(h[k] & 0xFFFFFFFFu);
    }
    return TR_OK;
}

int collide_debug_grid(tr_world* w, uint32_t* keys, uint32_t* sorted_ids, uint64_t* m, float* cell, float* inv_cell,
                       int32_t* bits) {
    if (w->n == 0) {
        if (m) *m = 0;
        return TR_OK;
    }
    if (!w->classified) TR_TRY(collide_classify(w));
    TR_TRY(ensure_buffers(w));
    Collide* c = state(w);
    TR_TRY(broad_phase(w));
    // K4, only here: the stable (key,id) order of the grid bodies
    canon_kernel<<<(uint32_t)((w->n + 255) / 256), 256, 0, w->stream>>>((const SlotRec*)c->recs, c->cell_start, grid_params(c),
                                                                          c->counters, c->sorted_ids);
    TR_CUDA(cudaGetLastError());
    TR_TRY(read_counters(w));
    const uint32_t gm = c->h_counters[C_GRID];
    if (m) *m = gm;
    if (cell) *cell = c->cell;
    if (inv_cell) *inv_cell = c->inv_cell;
    if (bits) *bits = c->bits;
    const uint32_t sentinel = 1u << (3 * c->bits);
    if (keys) {
        TR_CUDA(cudaMemcpyAsync(keys, c->keys, w->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
        TR_CUDA(cudaStreamSynchronize(w->stream));
        for (uint64_t i = 0; i < w->n; ++i)
            if (keys[i] == sentinel) keys[i] = EMPTY;
    }
    if (sorted_ids && gm) {
        TR_CUDA(cudaMemcpyAsync(sorted_ids, c->sorted_ids, gm * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
        TR_CUDA(cudaStreamSynchronize(w->stream));
    }
    return TR_OK;
}

int collide_debug_candidates(tr_world* w, int32_t* pairs, uint64_t cap, uint64_t* n) {
    if (w->n == 0) {
        if (n) *n = 0;
        return TR_OK;
    }
    if (!w->classified) TR_TRY(collide_classify(w));
    TR_TRY(ensure_buffers(w));
    unsigned long long* d = nullptr;
    uint64_t dcap = pairs ? cap : 0;
    if (dcap) TR_CUDA(cudaMalloc(&d, dcap * sizeof(unsigned long long)));
    uint64_t count = 0;
    int rc = detect<true>(w, d, dcap, &count);
    if (rc == TR_OK && n) *n = count;
    w->stats.last_candidates = count;
    if (rc == TR_OK && pairs && count) {
        uint64_t take = count < dcap ? count : dcap;
        std::vector<unsigned long long> h(take);
        cudaError_t e = cudaMemcpy(h.data(), d, take * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "candidate copy", __FILE__, __LINE__);
        std::sort(h.begin(), h.end());
        for (uint64_t k = 0; k < take; ++k) {
            pairs[2 * k] = (int32_t)(h[k] >> 32);
            pairs[2 * k + 1] = (int32_t)(h[k] & 0xFFFFFFFFu);
        }
    }
    if (d) cudaFree(d);
    return rc;
}

int collide_free(tr_world* w) {
    Collide* c = w->col;
    if (!c) return TR_OK;
    cudaFree(c->keys); cudaFree(c->recs); cudaFree(c->sorted_ids);
    cudaFree(c->cell_start); cudaFree(c->cell_count); cudaFree(c->tile_sum);
    cudaFree(c->side_list); cudaFree(c->far_list); cudaFree(c->bnd_list); cudaFree(c->far_c); cudaFree(c->bnd_c); cudaFree(c->klass);
    cudaFree(c->row_start); cudaFree(c->col_start); cudaFree(c->body_cnt); cudaFree(c->counters);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_rounds) cudaFreeHost(c->h_rounds);
    cudaFree(c->ckeys); cudaFree(c->ckeys_alt); cudaFree(c->colkeys);
    cudaFree(c->colc); cudaFree(c->col_pos); cudaFree(c->dep);
    cudaFree(c->succ_i); cudaFree(c->succ_j); cudaFree(c->crec); cudaFree(c->cver); cudaFree(c->cgeo); cudaFree(c->brec); cudaFree(c->frontier[0]); cudaFree(c->frontier[1]);
    if (c->bp_graph) cudaGraphExecDestroy(c->bp_graph);
    delete c;
    w->col = nullptr;
    return TR_OK;
}

}  // namespace tr

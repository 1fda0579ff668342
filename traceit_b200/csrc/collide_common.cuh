// Shared by the collision pipeline (collide.cu): constants, the per-world state, the reference's
// predicates restated with rounded intrinsics, and the two-kernel exclusive scan that both the
// broad phase (cell table) and the K6 prep (per-body contact tables) use.
#pragma once
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
    int bx = 2, by = 2, bz = 2;   // bits per axis of the wrapped cell key (x is the fastest digit)
    bool may_have_side = true;    // some body is statically outside the grid (oversize / NaN extent)
    bool mixed = false;           // live boxes exist: they share the grid, extents include collider::size
    uint64_t cap_bodies = 0;      // arrays below sized for this many bodies
    uint64_t ncell_alloc = 0;
    uint64_t cap_contacts = 0;

    // per body
    uint32_t* keys = nullptr;        // K2 out: key, or sentinel 1<<(bx+by+bz) for non-grid bodies
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

struct GridParams {
    float ox, oy, oz, inv_cell, rmax_grid;
    int bx, by, bz;   // bits per axis: the table has 2^(bx+by+bz) cells, at least 4 per axis
    int mixed;        // boxes live in the grid too: a body's extent is max(|radius| if sphere, |size.xyz|)
    __host__ __device__ int total_bits() const { return bx + by + bz; }
};

__device__ __forceinline__ uint32_t pack_key(int cx, int cy, int cz, const GridParams& g) {
    const uint32_t mx = (1u << g.bx) - 1u, my = (1u << g.by) - 1u, mz = (1u << g.bz) - 1u;
    return ((uint32_t)cx & mx) | (((uint32_t)cy & my) << g.bx) | (((uint32_t)cz & mz) << (g.bx + g.by));
}

__device__ __forceinline__ uint32_t pack_val(uint32_t id, uint8_t f) {
    return id | ((f & F_STATIC) ? STATIC_BIT : 0u) | ((f & F_SPHERE) ? 0u : BOX_BIT);
}

// ---------------------------------------------------------------------------------------
// Exclusive scan of a histogram in two unchained kernels (a look-back chain over
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

}  // namespace tr

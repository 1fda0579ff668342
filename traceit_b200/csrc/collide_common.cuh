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

// Index checks of a CHECKED build (make EXTRA=-DTR_CHECKED TAG=checked OUT=../libtraceit_b200_checked.so): every
// computed index into a scratch array is asserted against the array's extent; a violation traps the kernel
// (cudaErrorAssert at the next synchronisation).  compute-sanitizer is closed on the development pool, so the GPU
// suite is run once against this build instead (profiles/r02_checked_build.txt).  The product build compiles them out.
#ifdef TR_CHECKED
#include <cassert>
#define TR_ASSERT(cond) assert(cond)
#else
#define TR_ASSERT(cond) ((void)0)
#endif

namespace tr {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t ID_MASK = 0x3FFFFFFFu;      // body ids fit 30 bits (tr_bodies_add enforces it)
constexpr uint32_t STATIC_BIT = 0x80000000u;   // collider::isStatic
constexpr uint32_t BOX_BIT = 0x40000000u;      // !collider::isSphere: pairs with it use the AABB predicate
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// counters[] slots (device memory, zeroed by the last kernel of every collision step)
// The three counters every dataflow worker hammers (queue head, queue tail, finished work) sit on
// 128-byte lines of their own: on one line their atomics and the idle workers' polls all queue at
// one L2 slice (having idle workers merely READ the tail next to the pushes doubled the step time).
enum { C_SIDE = 0, C_GRID = 1, C_CONTACTS = 2 /*u64*/, C_CAND = 4 /*u64*/, C_HE = 6 /* half-edge slots handed out */,
       C_LONG = 7 /* bodies with a long contact list */, C_FRONT = 8 /*3*/, C_ROUNDS = 11, C_ABORT = 12, C_FIN = 13, C_BAR = 14 /* grid barrier of prep_kernel */, C_HUGE = 15 /* bodies whose contact list the whole grid sorts */,
       C_FAR = 16, C_BND = 17, C_T0 = 18 /* 3 x u64 globaltimer stamps: broad, narrow, prep */,
       C_NHOST = 24, C_SPARSE = 25 /* sparse_step_kernel replayed this step (1 + its rounds): the general kernels behind it in the graph return */,
       C_QHEAD = 32, C_QTAIL = 64, C_DONE = 96, C_NCOUNTERS = 128 };

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Opportunistic warp-aggregated append: the lanes that reach this call together reserve their output slots with ONE
// atomic (same-address atomics retire at ~0.85 clk each at L2).  Returns the slot of the calling lane.
__device__ __forceinline__ unsigned long long warp_reserve(unsigned long long* counter) {
    const uint32_t mask = __activemask();
    const uint32_t lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}
__device__ __forceinline__ uint32_t warp_reserve32(uint32_t* counter) {
    const uint32_t mask = __activemask();
    const uint32_t lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

}  // namespace tr

struct Collide {
    // grid description (host decides, see collide_classify)
    float cell = 1.f, inv_cell = 1.f, rmax_grid = 0.f;
    float origin[3] = {0, 0, 0};
    int bx = 2, by = 2, bz = 2;   // bits per axis of the wrapped cell key (x is the fastest digit)
    bool mixed = false;           // live boxes exist: they share the grid, extents include collider::size
    uint64_t cap_bodies = 0;      // arrays below sized for this many bodies
    uint64_t ncell_alloc = 0;
    uint64_t cap_contacts = 0;
    bool cap_fixed = false;       // the caller set tr_world_params.max_contacts: never grown

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
    uint32_t* body_cnt = nullptr;    // [N] contacts per body; self-cleaning (prep_kernel zeroes what it touched)
    uint2* seg = nullptr;            // [N] {first half-edge slot, contacts} of every body in contact this step
    uint32_t* long_list = nullptr;   // bodies whose contact list is long enough to be sorted by a CTA
    uint32_t* huge_list = nullptr;   // ... by all CTAs together (more than PREP_SMEM_SORT contacts)

    uint32_t* counters = nullptr;

    // per contact
    unsigned long long* ckeys = nullptr;      // (i<<32)|j in arrival order (sorted on the host when tr_read_contacts asks)
    uint2* arank = nullptr;                   // arrival rank of the contact among its bodies' contacts {in i's list, in j's list}
    uint2* he = nullptr;                      // [2 * cap] half edges {partner id, contact}, one segment per body
    uint32_t* dep = nullptr;
    uint4* crec = nullptr;                    // packed {i, j, succ_i, succ_j} per contact
    uint2* cver = nullptr;                    // versions of bodies i and j that contact c must see (dataflow executor)
    struct tr_contact_geo* cgeo = nullptr;    // per contact: normal, inverse masses, restitution, static bits (tr::ContactGeo)
    struct tr_body_rec* brec = nullptr;       // per body: version-checked velocity + position record (tr::BodyRec)
    uint32_t* frontier[2] = {nullptr, nullptr};   // [0]: ready queue of the dataflow executor / first frontier; [1]: level-synchronous only

    // the whole collision step as ONE CUDA graph per live pos_m buffer: launch parameters only change when the
    // buffers, the grid, the body count or the executor choice change, so it is re-captured only then
    // three variants per buffer ([3 * buffer + variant], collide.cu: StepVariant)
    cudaGraphExec_t graph[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t graph_sig[6] = {0, 0, 0, 0, 0, 0};
    // sparse steps are PREDICTED by the host (newest step record it can see); a wrong guess is caught on the device and
    // the step re-run through the general kernels, after which the prediction stays off for sparse_block steps
    uint32_t sparse_block = 0, sparse_backoff = 32;
    uint32_t peek_seq = 0;
    uint64_t known_contacts = 0;       // contact count of the newest step whose record the host has seen
    bool known_valid = false;
    bool sparse_smem_set = false, sparse_ok = false;   // kernel attributes set; the device can place the sparse-step kernel's cluster

    // what the host knows from the step records it has consumed so far
    uint64_t last_contacts = 0;
    uint32_t last_side = 0, last_far = 0;
    int coop_blocks = 0;
    int prep_blocks = 0;
    uint64_t prev_contacts = 0;   // last dataflow response: contacts and iterations of the longest worker
    uint32_t prev_rounds = 0;
    std::vector<unsigned long long> host_contacts;   // tr_read_contacts: the last step's list, sorted on demand
    bool host_contacts_valid = false;
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
// Bitonic sort of {key, payload} pairs by key: shared-memory chunks + global long-distance steps.
// Used by the K6 prep (a body's contact list by partner id) and by K4 (a crowded cell by body id).
// ---------------------------------------------------------------------------------------
constexpr uint32_t PREP_SMEM_SORT = 2048;  // entries a CTA sorts in shared memory (16 KB)

// Ascending bitonic network in its flip + shear form: stage k = one "flip" step (i <-> the mirror position inside
// its block of k) followed by shear steps at distances k/4 ... 1 (i <-> i + j inside blocks of 2j).  Every
// compare-exchange puts the smaller key at the lower index, so virtual +inf padding beyond len never moves down
// and any length works: a pair whose upper index is >= len is simply skipped.
// pair t of a flip step of stage k / of a shear step at distance j:
__device__ __forceinline__ void flip_pair(uint32_t t, uint32_t k, uint32_t& i, uint32_t& l) {
    const uint32_t g = t / (k / 2), r = t % (k / 2);
    i = g * k + r;
    l = g * k + (k - 1 - r);
}
__device__ __forceinline__ void shear_pair(uint32_t t, uint32_t j, uint32_t& i, uint32_t& l) {
    const uint32_t g = t / j, r = t % j;
    i = g * 2 * j + r;
    l = i + j;
}
__device__ __forceinline__ void smem_cmpx(uint2* s, uint32_t i, uint32_t l, uint32_t len) {
    if (l < len && s[i].x > s[l].x) {
        const uint2 t = s[i];
        s[i] = s[l];
        s[l] = t;
    }
}
// one CTA, a chunk of len <= PREP_SMEM_SORT entries in shared memory: the whole network (from_k = 2) or only the
// shear steps below the chunk size that finish a stage whose upper steps were done in global memory (shears_only)
__device__ __forceinline__ void smem_sort(uint2* s, uint32_t len, bool shears_only) {
    uint32_t pow2 = 1;
    while (pow2 < len) pow2 <<= 1;
    uint32_t i, l;
    if (!shears_only) {
        for (uint32_t k = 2; k <= pow2; k <<= 1) {
            for (uint32_t t = threadIdx.x; t < pow2 / 2; t += blockDim.x) {
                flip_pair(t, k, i, l);
                smem_cmpx(s, i, l, len);
            }
            __syncthreads();
            for (uint32_t j = k / 4; j >= 1; j >>= 1) {
                for (uint32_t t = threadIdx.x; t < pow2 / 2; t += blockDim.x) {
                    shear_pair(t, j, i, l);
                    smem_cmpx(s, i, l, len);
                }
                __syncthreads();
            }
        }
    } else {
        for (uint32_t j = PREP_SMEM_SORT / 2; j >= 1; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < PREP_SMEM_SORT / 2; t += blockDim.x) {
                shear_pair(t, j, i, l);
                smem_cmpx(s, i, l, len);
            }
            __syncthreads();
        }
    }
}
__device__ __forceinline__ void global_cmpx(uint2* p, uint32_t i, uint32_t l, uint32_t len) {
    if (l < len) {
        const uint2 x = __ldcg(p + i), y = __ldcg(p + l);
        if (x.x > y.x) {
            __stcg(p + i, y);
            __stcg(p + l, x);
        }
    }
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

// Programmatic dependent launch inside the step's graph: a kernel launched with pdl_launch() (collide.cu) may be
// scheduled while its predecessor drains - its CTAs' launch and prologue overlap the predecessor's tail - and waits
// in pdl_wait() until the predecessor has completed and flushed.  Every kernel of the chain waits FIRST (every CTA,
// also the ones that return early: completion is transitive only then) and lets its own dependents go at once.
// Launched the ordinary way, both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__global__ void __launch_bounds__(SCAN_THREADS) tile_sums_kernel(const uint32_t* __restrict__ cell_count, uint32_t ncell,
                                                                 uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    pdl_wait();
    pdl_trigger();
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
    pdl_wait();
    pdl_trigger();
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

// K2-K6 - sphere/AABB collision detection and order-faithful response.
//
// Replaces resolveCollisions (reference src/main.cpp:673-745): an O(N^2) lexicographic
// i<j sweep with checkSphereCollision / checkAABBCollision (src/main.cpp:648-660) and a
// Gauss-Seidel impulse + positional nudge (src/main.cpp:709-742).
//
//   K2  cell keys            one thread per body: wrapped uniform-grid key from the collider
//                            centre; bodies that cannot live in the grid (boxes, oversize or
//                            out-of-domain spheres) are appended to a side list; mass<=0
//                            bodies are dropped (they can never collide, main.cpp:696).
//   K3  radix sort           (key, id) pairs, stable => (key,id) order.
//   K4  cell ranges + gather first slot of every occupied cell; centres gathered to sorted order.
//   K5  narrow phase         27-cell stencil per grid body, exact reference predicate
//                            (no FMA contraction), emits (i<j) contacts.
//   K5b side list            every side-list body against every live body.
//   K6  ordered response     contacts sorted by (i,j); a dependency wavefront replays the
//                            reference's Gauss-Seidel order exactly: a contact runs when
//                            the previous contact of BOTH its bodies has run, so results
//                            are bit-identical to the sequential sweep.
//
// The contact SET is a pure function of collider centres/radii (the sweep never writes
// them: main.cpp:655-659 vs 725-741), which is what makes detection parallel and K6 legal.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>

#include "world.cuh"

namespace cg = cooperative_groups;

namespace tr {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t ID_MASK = 0x7FFFFFFFu;
constexpr uint32_t STATIC_BIT = 0x80000000u;

}  // namespace tr

struct Collide {
    // grid description (host decides, see collide_classify)
    float cell = 1.f, inv_cell = 1.f, rmax_grid = 0.f;
    float origin[3] = {0, 0, 0};
    int bits = 2;
    uint64_t cap_bodies = 0;      // arrays below sized for this many bodies
    uint64_t ncell_alloc = 0;
    uint64_t cap_contacts = 0;

    // per body
    uint32_t* keys = nullptr;        // K2 out: key, or sentinel 1<<(3*bits) for non-grid bodies
    uint32_t* keys_sorted = nullptr;
    uint32_t* vals = nullptr;        // id | static<<31
    uint32_t* vals_sorted = nullptr;
    float4* sorted_center_r = nullptr;
    uint32_t* cell_start = nullptr;  // [ncell]
    uint32_t* side_list = nullptr;   // ids of side-list bodies
    uint8_t* klass = nullptr;        // 0 dropped, 1 grid, 2 side
    uint32_t* row_start = nullptr;   // per body: first contact with i == body
    uint32_t* col_last = nullptr;    // per body: last column-order slot with j == body

    // counters: [0] side count, [1] grid count, [2..3] contact count (u64), [4..5] candidates (u64),
    // [8..10] frontier counts, [11] rounds
    uint32_t* counters = nullptr;
    uint32_t* h_counters = nullptr;  // pinned mirror

    // per contact
    unsigned long long* ckeys = nullptr;      // (i<<32)|j, unsorted then sorted
    unsigned long long* ckeys_alt = nullptr;
    unsigned long long* colkeys = nullptr;    // (j<<32)|i
    unsigned long long* colkeys_alt = nullptr;
    uint32_t* colc = nullptr;                 // contact index per column slot
    uint32_t* colc_alt = nullptr;
    uint32_t* col_pos = nullptr;              // inverse of colc
    uint32_t* dep = nullptr;
    uint32_t* succ_i = nullptr;
    uint32_t* succ_j = nullptr;
    uint32_t* frontier[2] = {nullptr, nullptr};

    void* cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;

    uint64_t last_contacts = 0;
    uint64_t last_grid = 0;
    uint32_t last_side = 0;
    int coop_blocks = 0;
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

__device__ __forceinline__ int cell_coord(float center, float origin, float inv_cell) {
    float u = __fmul_rn(__fsub_rn(center, origin), inv_cell);
    return __float2int_rd(u);   // floor, saturating, NaN -> 0
}

__device__ __forceinline__ uint32_t pack_key(int cx, int cy, int cz, int bits) {
    uint32_t m = (1u << bits) - 1u;
    return ((uint32_t)cx & m) | (((uint32_t)cy & m) << bits) | (((uint32_t)cz & m) << (2 * bits));
}

struct GridParams {
    float ox, oy, oz, inv_cell, rmax_grid;
    int bits;
};

// ---------------------------------------------------------------------------------------
// K2: keys + classification
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) keys_kernel(const float4* __restrict__ center_r, const float4* __restrict__ pos_m,
                                                   const uint8_t* __restrict__ flags, uint32_t n, GridParams g,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                   uint8_t* __restrict__ klass, uint32_t* __restrict__ side_list,
                                                   uint32_t* counters) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = center_r[i];
    const float mass = pos_m[i].w;
    const uint8_t f = flags[i];
    const uint32_t sentinel = 1u << (3 * g.bits);
    uint32_t key = sentinel;
    uint8_t k = 0;
    if (!(mass <= 0)) {   // main.cpp:696 drops mass<=0 bodies from every pair
        float ux = __fmul_rn(__fsub_rn(c.x, g.ox), g.inv_cell);
        float uy = __fmul_rn(__fsub_rn(c.y, g.oy), g.inv_cell);
        float uz = __fmul_rn(__fsub_rn(c.z, g.oz), g.inv_cell);
        bool in_domain = fabsf(ux) < 32768.0f && fabsf(uy) < 32768.0f && fabsf(uz) < 32768.0f;
        bool grid_ok = (f & F_SPHERE) && in_domain && fabsf(c.w) <= g.rmax_grid;
        if (grid_ok) {
            key = pack_key(__float2int_rd(ux), __float2int_rd(uy), __float2int_rd(uz), g.bits);
            k = 1;
        } else {
            k = 2;
            uint32_t slot = atomicAdd(&counters[0], 1u);
            side_list[slot] = i;
        }
    }
    keys[i] = key;
    vals[i] = i | ((f & F_STATIC) ? STATIC_BIT : 0u);
    klass[i] = k;
}

// ---------------------------------------------------------------------------------------
// K4: cell ranges + gather to sorted order
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ranges_kernel(const uint32_t* __restrict__ keys_sorted,
                                                     const uint32_t* __restrict__ vals_sorted,
                                                     const float4* __restrict__ center_r, uint32_t n, uint32_t sentinel,
                                                     uint32_t* __restrict__ cell_start,
                                                     float4* __restrict__ sorted_center_r, uint32_t* counters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t k = keys_sorted[t];
    uint32_t kp = t ? keys_sorted[t - 1] : EMPTY;
    if (k == sentinel) {
        if (kp != sentinel) counters[1] = t;   // number of grid bodies
        return;
    }
    if (t == n - 1) counters[1] = n;
    if (k != kp) cell_start[k] = t;
    sorted_center_r[t] = center_r[vals_sorted[t] & ID_MASK];
}

// ---------------------------------------------------------------------------------------
// K5: narrow phase over the 27-cell stencil
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_pair(unsigned long long* out, unsigned long long cap, unsigned long long* counter,
                                          uint32_t i, uint32_t j) {
    unsigned long long slot = atomicAdd(counter, 1ull);
    if (slot < cap) out[slot] = ((unsigned long long)i << 32) | j;
}

template <bool CANDIDATES>
__global__ void __launch_bounds__(128) narrow_kernel(const uint32_t* __restrict__ keys_sorted,
                                                     const uint32_t* __restrict__ vals_sorted,
                                                     const float4* __restrict__ sorted_center_r,
                                                     const uint32_t* __restrict__ cell_start, uint32_t m, int bits,
                                                     unsigned long long* out, unsigned long long cap,
                                                     unsigned long long* counter) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const uint32_t key = keys_sorted[t];
    const uint32_t va = vals_sorted[t];
    const uint32_t ida = va & ID_MASK;
    const float4 ca = sorted_center_r[t];
    const uint32_t msk = (1u << bits) - 1u;
    const uint32_t cx = key & msk, cy = (key >> bits) & msk, cz = (key >> (2 * bits)) & msk;
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll 1
            for (int dx = -1; dx <= 1; ++dx) {
                uint32_t nk = ((cx + dx) & msk) | (((cy + dy) & msk) << bits) | (((cz + dz) & msk) << (2 * bits));
                uint32_t s = cell_start[nk];
                if (s == EMPTY) continue;
                for (uint32_t u = s; u < m && keys_sorted[u] == nk; ++u) {
                    const uint32_t vb = vals_sorted[u];
                    const uint32_t idb = vb & ID_MASK;
                    if (idb <= ida) continue;
                    if (CANDIDATES) {
                        emit_pair(out, cap, counter, ida, idb);
                    } else {
                        if (va & vb & STATIC_BIT) continue;                      // main.cpp:695
                        if (sphere_hit(ca, sorted_center_r[u])) emit_pair(out, cap, counter, ida, idb);
                    }
                }
            }
        }
    }
}

// K5b: side-list bodies against every live body.  grid = (ceil(n/256), side_count).
__global__ void __launch_bounds__(256) side_kernel(const uint32_t* __restrict__ side_list, uint32_t side_count,
                                                   const float4* __restrict__ center_r,
                                                   const float4* __restrict__ half_ext,
                                                   const uint8_t* __restrict__ flags, const uint8_t* __restrict__ klass,
                                                   uint32_t n, unsigned long long* out, unsigned long long cap,
                                                   unsigned long long* counter) {
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

// ---------------------------------------------------------------------------------------
// K6 preparation: dependency structure of the lexicographic Gauss-Seidel order
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colkeys_kernel(const unsigned long long* __restrict__ ckeys, uint32_t c_n,
                                                      unsigned long long* __restrict__ colkeys, uint32_t* __restrict__ colc,
                                                      uint32_t* __restrict__ row_start) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_n) return;
    unsigned long long k = ckeys[c];
    uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
    colkeys[c] = ((unsigned long long)j << 32) | i;
    colc[c] = c;
    if (c == 0 || (uint32_t)(ckeys[c - 1] >> 32) != i) row_start[i] = c;
}

__global__ void __launch_bounds__(256) colpos_kernel(const unsigned long long* __restrict__ colkeys_sorted,
                                                     const uint32_t* __restrict__ colc_sorted, uint32_t c_n,
                                                     uint32_t* __restrict__ col_pos, uint32_t* __restrict__ col_last) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= c_n) return;
    col_pos[colc_sorted[t]] = t;
    uint32_t j = (uint32_t)(colkeys_sorted[t] >> 32);
    if (t == c_n - 1 || (uint32_t)(colkeys_sorted[t + 1] >> 32) != j) col_last[j] = t;
}

// Body b's contacts in sweep order are: its column entries (k,b) by ascending k, then its
// row entries (b,j) by ascending j.  pred/succ below are the neighbours of contact c in the
// lists of its two bodies.
__global__ void __launch_bounds__(256) deps_kernel(const unsigned long long* __restrict__ ckeys,
                                                   const unsigned long long* __restrict__ colkeys_sorted,
                                                   const uint32_t* __restrict__ colc_sorted,
                                                   const uint32_t* __restrict__ col_pos,
                                                   const uint32_t* __restrict__ row_start,
                                                   const uint32_t* __restrict__ col_last, uint32_t c_n,
                                                   uint32_t* __restrict__ dep, uint32_t* __restrict__ succ_i,
                                                   uint32_t* __restrict__ succ_j, uint32_t* __restrict__ frontier0,
                                                   uint32_t* counters) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_n) return;
    unsigned long long k = ckeys[c];
    uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
    uint32_t t = col_pos[c];
    bool row_prev = c > 0 && (uint32_t)(ckeys[c - 1] >> 32) == i;
    bool pred_i = row_prev || col_last[i] != EMPTY;
    bool pred_j = t > 0 && (uint32_t)(colkeys_sorted[t - 1] >> 32) == j;
    uint32_t d = (pred_i ? 1u : 0u) + (pred_j ? 1u : 0u);
    dep[c] = d;
    succ_i[c] = (c + 1 < c_n && (uint32_t)(ckeys[c + 1] >> 32) == i) ? c + 1 : EMPTY;
    uint32_t sj = EMPTY;
    if (t + 1 < c_n && (uint32_t)(colkeys_sorted[t + 1] >> 32) == j) sj = colc_sorted[t + 1];
    else sj = row_start[j];
    succ_j[c] = sj;
    if (d == 0) frontier0[atomicAdd(&counters[8], 1u)] = c;
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
    // L2 loads: these were written by other SMs in earlier wavefronts
    float4 vi = __ldcg(a.vel_rest + i), vj = __ldcg(a.vel_rest + j);
    float rvx = __fsub_rn(vi.x, vj.x), rvy = __fsub_rn(vi.y, vj.y), rvz = __fsub_rn(vi.z, vj.z);
    float vn = dot3_rn(rvx, rvy, rvz, nx, ny, nz);
    if (vn > 0) return;                                  // main.cpp:714 (as written)
    float4 pi = __ldcg(a.pos_m + i), pj = __ldcg(a.pos_m + j);
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

// Host-side choice of the grid: which radii are admitted, the cell edge, bits per axis.
int collide_classify(tr_world* w) {
    Collide* c = state(w);
    const tr_world_params& p = w->params;
    // admitted radius: user value, else max |r| over spheres with |r| <= 8 * median |r|
    float rmax = p.grid_max_radius;
    if (rmax <= 0.f) {
        std::vector<float> r;
        r.reserve(w->host_bodies.size());
        for (const tr_body_desc& b : w->host_bodies)
            if (b.isSphere && !(b.mass <= 0) && std::isfinite(b.radius)) r.push_back(std::fabs(b.radius));
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
        TR_TRY(realloc_dev(&c->keys_sorted, cap));
        TR_TRY(realloc_dev(&c->vals, cap));
        TR_TRY(realloc_dev(&c->vals_sorted, cap));
        TR_TRY(realloc_dev(&c->sorted_center_r, cap));
        TR_TRY(realloc_dev(&c->side_list, cap));
        TR_TRY(realloc_dev(&c->klass, cap));
        TR_TRY(realloc_dev(&c->row_start, cap));
        TR_TRY(realloc_dev(&c->col_last, cap));
        c->cap_bodies = cap;
    }
    const uint64_t ncell = 1ull << (3 * c->bits);
    if (ncell > c->ncell_alloc) {
        TR_TRY(realloc_dev(&c->cell_start, ncell));
        c->ncell_alloc = ncell;
    }
    uint64_t want_c = w->params.max_contacts ? w->params.max_contacts : 8 * n + 4096;
    if (want_c > 0xFFFFFFF0ull) want_c = 0xFFFFFFF0ull;
    if (want_c != c->cap_contacts) {
        TR_TRY(realloc_dev(&c->ckeys, want_c));
        TR_TRY(realloc_dev(&c->ckeys_alt, want_c));
        TR_TRY(realloc_dev(&c->colkeys, want_c));
        TR_TRY(realloc_dev(&c->colkeys_alt, want_c));
        TR_TRY(realloc_dev(&c->colc, want_c));
        TR_TRY(realloc_dev(&c->colc_alt, want_c));
        TR_TRY(realloc_dev(&c->col_pos, want_c));
        TR_TRY(realloc_dev(&c->dep, want_c));
        TR_TRY(realloc_dev(&c->succ_i, want_c));
        TR_TRY(realloc_dev(&c->succ_j, want_c));
        TR_TRY(realloc_dev(&c->frontier[0], want_c));
        TR_TRY(realloc_dev(&c->frontier[1], want_c));
        c->cap_contacts = want_c;
    }
    if (!c->counters) {
        TR_CUDA(cudaMalloc(&c->counters, 16 * sizeof(uint32_t)));
        TR_CUDA(cudaMallocHost(&c->h_counters, 16 * sizeof(uint32_t)));
    }
    // CUB scratch sized for the largest of the three sorts
    size_t need = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, c->keys, c->keys_sorted, c->vals, c->vals_sorted, (uint32_t)c->cap_bodies);
    need = b;
    cub::DeviceRadixSort::SortKeys(nullptr, b, c->ckeys, c->ckeys_alt, (uint32_t)c->cap_contacts);
    need = b > need ? b : need;
    cub::DeviceRadixSort::SortPairs(nullptr, b, c->colkeys, c->colkeys_alt, c->colc, c->colc_alt, (uint32_t)c->cap_contacts);
    need = b > need ? b : need;
    if (need > c->cub_tmp_bytes) {
        if (c->cub_tmp) cudaFree(c->cub_tmp);
        c->cub_tmp = nullptr;
        TR_CUDA(cudaMalloc(&c->cub_tmp, need));
        c->cub_tmp_bytes = need;
    }
    return TR_OK;
}

static int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 32 && (max_value >> b) != 0) ++b;
    return b;
}

// Broad phase K2-K4.  Leaves counters[0] (side count) and counters[1] (grid count) on the device.
static int broad_phase(tr_world* w) {
    Collide* c = state(w);
    const uint32_t n = (uint32_t)w->n;
    const uint32_t blocks = (n + 255) / 256;
    GridParams g{c->origin[0], c->origin[1], c->origin[2], c->inv_cell, c->rmax_grid, c->bits};
    const uint32_t sentinel = 1u << (3 * c->bits);
    PendingTimer t;
    timer_begin(w, T_BROAD, &t);
    TR_CUDA(cudaMemsetAsync(c->counters, 0, 16 * sizeof(uint32_t), w->stream));
    TR_CUDA(cudaMemsetAsync(c->cell_start, 0xFF, (1ull << (3 * c->bits)) * sizeof(uint32_t), w->stream));
    keys_kernel<<<blocks, 256, 0, w->stream>>>(w->center_r, w->pos_m[w->cur], w->flags, n, g, c->keys, c->vals, c->klass,
                                               c->side_list, c->counters);
    size_t tmp = c->cub_tmp_bytes;
    TR_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, c->keys, c->keys_sorted, c->vals, c->vals_sorted, n, 0,
                                            3 * c->bits + 1, w->stream));
    ranges_kernel<<<blocks, 256, 0, w->stream>>>(c->keys_sorted, c->vals_sorted, w->center_r, n, sentinel, c->cell_start,
                                                 c->sorted_center_r, c->counters);
    timer_end(w, &t);
    TR_CUDA(cudaGetLastError());
    w->stats.kernel_launches += 2;   // keys + ranges (CUB's sort passes are library kernels, not counted)
    return TR_OK;
}

static int read_counters(tr_world* w) {
    Collide* c = state(w);
    TR_CUDA(cudaMemcpyAsync(c->h_counters, c->counters, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    TR_CUDA(cudaStreamSynchronize(w->stream));
    return TR_OK;
}

// Detection: broad phase + narrow phase + side list.  On return the UNSORTED pairs are in
// ckeys and *count is their number (may exceed cap_contacts => overflow).
template <bool CANDIDATES>
static int detect(tr_world* w, unsigned long long* out, uint64_t cap, uint64_t* count) {
    Collide* c = state(w);
    TR_TRY(broad_phase(w));
    TR_TRY(read_counters(w));
    const uint32_t side = c->h_counters[0];
    const uint32_t m = c->h_counters[1];
    c->last_side = side;
    c->last_grid = m;
    unsigned long long* counter = (unsigned long long*)(c->counters + (CANDIDATES ? 4 : 2));
    PendingTimer t;
    timer_begin(w, T_NARROW, &t);
    if (m) {
        narrow_kernel<CANDIDATES><<<(m + 127) / 128, 128, 0, w->stream>>>(c->keys_sorted, c->vals_sorted, c->sorted_center_r,
                                                                          c->cell_start, m, c->bits, out, cap, counter);
        w->stats.kernel_launches += 1;
    }
    if (side && !CANDIDATES) {
        dim3 grid((uint32_t)((w->n + 255) / 256), side < 64 ? side : 64);
        side_kernel<<<grid, 256, 0, w->stream>>>(c->side_list, side, w->center_r, w->half_ext, w->flags, c->klass,
                                                 (uint32_t)w->n, out, cap, counter);
        w->stats.kernel_launches += 1;
    }
    timer_end(w, &t);
    TR_CUDA(cudaGetLastError());
    TR_TRY(read_counters(w));
    *count = ((uint64_t)c->h_counters[CANDIDATES ? 5 : 3] << 32) | c->h_counters[CANDIDATES ? 4 : 2];
    return TR_OK;
}

int collide_step(tr_world* w, bool /*emit_candidates*/) {
    if (w->n == 0) return TR_OK;
    if (!w->classified) TR_TRY(collide_classify(w));
    TR_TRY(ensure_buffers(w));
    Collide* c = state(w);
    uint64_t count = 0;
    TR_TRY(detect<false>(w, c->ckeys, c->cap_contacts, &count));
    w->stats.side_list_size = c->last_side;
    w->stats.last_contacts = count;
    c->last_contacts = 0;
    if (count > c->cap_contacts) {
        set_error("contact buffer overflow: %llu contacts, capacity %llu (raise tr_world_params.max_contacts)",
                  (unsigned long long)count, (unsigned long long)c->cap_contacts);
        return TR_ERR_CAPACITY;
    }
    w->stats.last_response_rounds = 0;
    if (count == 0) return TR_OK;
    const uint32_t cn = (uint32_t)count;
    const uint32_t cblocks = (cn + 255) / 256;
    const int idbits = bits_for(w->n - 1);
    PendingTimer t;
    timer_begin(w, T_RESPONSE, &t);
    // lexicographic (i,j) order == the reference's visiting order
    size_t tmp = c->cub_tmp_bytes;
    TR_CUDA(cub::DeviceRadixSort::SortKeys(c->cub_tmp, tmp, c->ckeys, c->ckeys_alt, cn, 0, 32 + idbits, w->stream));
    std::swap(c->ckeys, c->ckeys_alt);
    TR_CUDA(cudaMemsetAsync(c->row_start, 0xFF, w->n * sizeof(uint32_t), w->stream));
    TR_CUDA(cudaMemsetAsync(c->col_last, 0xFF, w->n * sizeof(uint32_t), w->stream));
    colkeys_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys, cn, c->colkeys, c->colc, c->row_start);
    tmp = c->cub_tmp_bytes;
    TR_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, c->colkeys, c->colkeys_alt, c->colc, c->colc_alt, cn, 0,
                                            32 + idbits, w->stream));
    colpos_kernel<<<cblocks, 256, 0, w->stream>>>(c->colkeys_alt, c->colc_alt, cn, c->col_pos, c->col_last);
    deps_kernel<<<cblocks, 256, 0, w->stream>>>(c->ckeys, c->colkeys_alt, c->colc_alt, c->col_pos, c->row_start, c->col_last,
                                                cn, c->dep, c->succ_i, c->succ_j, c->frontier[0], c->counters);
    if (c->coop_blocks == 0) {
        int per_sm = 0;
        TR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, response_kernel, 256, 0));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;
        c->coop_blocks = per_sm * w->num_sms;
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
    w->stats.kernel_launches += 4;   // colkeys, colpos, deps, response
    c->last_contacts = count;
    // rounds are read lazily with the next counter read-back; fetch now (cheap, one sync per step already paid)
    TR_TRY(read_counters(w));
    w->stats.last_response_rounds = c->h_counters[11];
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
    TR_TRY(read_counters(w));
    const uint32_t gm = c->h_counters[1];
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
        TR_CUDA(cudaMemcpyAsync(sorted_ids, c->vals_sorted, gm * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
        TR_CUDA(cudaStreamSynchronize(w->stream));
        for (uint32_t i = 0; i < gm; ++i) sorted_ids[i] &= ID_MASK;
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
    cudaFree(c->keys); cudaFree(c->keys_sorted); cudaFree(c->vals); cudaFree(c->vals_sorted);
    cudaFree(c->sorted_center_r); cudaFree(c->cell_start); cudaFree(c->side_list); cudaFree(c->klass);
    cudaFree(c->row_start); cudaFree(c->col_last); cudaFree(c->counters);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    cudaFree(c->ckeys); cudaFree(c->ckeys_alt); cudaFree(c->colkeys); cudaFree(c->colkeys_alt);
    cudaFree(c->colc); cudaFree(c->colc_alt); cudaFree(c->col_pos); cudaFree(c->dep);
    cudaFree(c->succ_i); cudaFree(c->succ_j); cudaFree(c->frontier[0]); cudaFree(c->frontier[1]);
    cudaFree(c->cub_tmp);
    delete c;
    w->col = nullptr;
    return TR_OK;
}

}  // namespace tr

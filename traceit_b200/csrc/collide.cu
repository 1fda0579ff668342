// K2-K6 - sphere/AABB collision detection and order-faithful response.
//
// Replaces resolveCollisions (reference src/main.cpp:673-745): an O(N^2) lexicographic
// i<j sweep with checkSphereCollision / checkAABBCollision (src/main.cpp:648-660) and a
// Gauss-Seidel impulse + positional nudge (src/main.cpp:709-742).
//
// One collision step = ONE CUDA graph, no host round trip anywhere in it (collide_step below):
//   K2  keys_kernel          one thread per body: cell key from the collider centre, per-cell
//                            histogram by atomics; bodies that cannot live in the grid (oversize,
//                            out-of-domain, non-finite) go to the side / far lists; mass<=0
//                            bodies are dropped (they can never collide, main.cpp:696).
//   K3a tile_sums + scan     exclusive scan of the histogram in two unchained kernels
//                            -> dense cell_start[ncell+1] table.
//   K3b scatter_kernel       one 32-byte record per body (centre+radius, id|flags, half extents)
//                            into its cell's segment: the bodies in cell order (a ONE-DIGIT radix
//                            sort of the wrapped cell keys, radix = number of cells).
//   K4  canon_kernel         ON DEMAND (tr_debug_grid): the stable (key,id) order.
//   K5  narrow_kernel        half of the 27-cell stencil (each pair reached once); the (body, candidate)
//                            pairs of 32 consecutive slots are flattened into a per-warp list so that every
//                            lane tests a pair every iteration; exact reference predicate.  The trailing
//                            CTAs of the same launch are the side-list / far-list workers (K5b / K5c).
//   K6  prep_kernel          half-edge sort of the contact list by (body, partner): each contact's
//                            predecessors / successors in the reference's Gauss-Seidel order.
//       response_flow_kernel the ordered replay: a contact runs when the previous contact of BOTH its
//                            bodies has run, so results are bit-identical to the sequential sweep.
//       epilogue_kernel      the step's record for the host (mapped pinned ring), counters re-armed.
//   K6s sparse_step_kernel   a step with few contacts: ordering + ordered replay + record in ONE kernel of one
//                            16-CTA cluster, in distributed shared memory - instead of, or in front of, the three
//                            K6 kernels (StepVariant below).
//   K3a, K3b, K5, K6s are programmatic dependent launches: scheduled while their predecessor drains.
// Worlds of up to TINY_MAX bodies skip all of that: tiny_step_kernel (tiny.cuh) runs whole steps in one CTA.
//
// The contact SET is a pure function of collider centres/radii (the sweep never writes
// them: main.cpp:655-659 vs 725-741), which is what makes detection parallel and K6 legal -
// and what makes a contact-buffer overflow recoverable: the response of that step is skipped on the
// device, every later kernel becomes a no-op (tr_world::d_ctl), and the host re-runs detection +
// response with larger buffers at its next synchronisation (api.cu settle()).
#include "collide_common.cuh"
#include "collide_broad.cuh"
#include "collide_response.cuh"
#include "tiny.cuh"

namespace tr {

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
    return GridParams{c->origin[0], c->origin[1], c->origin[2], c->inv_cell, c->rmax_grid, c->bx, c->by, c->bz, c->mixed ? 1 : 0};
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
    // 2*rmax*(1+2^-7) is the SMALLEST conservative cell (margin analysed in DESIGN.md): with a smaller one the
    // 27-cell stencil would silently miss contacts, so a smaller request is raised to it
    float cell = (2.0f * rmax) * 1.0078125f;
    if (p.grid_cell > cell) cell = p.grid_cell;
    if (!(cell > 0.f) || !std::isfinite(cell)) cell = 1.0f;
    c->cell = cell;
    c->inv_cell = 1.0f / cell;
    memcpy(c->origin, p.grid_origin, sizeof(c->origin));
    if (p.grid_bits != 0) {
        c->bx = c->by = c->bz = p.grid_bits;       // the caller's choice: the same on every axis
    } else {
        // smallest table with at least 2 cells per body, bits spread over the axes (x, the fastest key
        // digit, first): with the same bits on every axis the table could only grow in steps of 8x
        // (2 cells per body is the measured optimum at 1M: broad + narrow = 32 + 69 us with 1, 35 + 44 with 2, 46 + 33 with 4,
        // 71 + 30 with 8 on the sparse cloud - a bigger table thins out the aliases the narrow phase has to reject but
        // costs the scans and the scattered histogram updates; the packed lattice has no aliases and only pays: 99 / 102 / 109 / 120)
        const uint64_t want = 2 * (w->n ? w->n : 1);
        int total = 6;
        while (total < 30 && (1ull << total) < want) ++total;
        c->bx = (total + 2) / 3;
        c->by = (total - c->bx + 1) / 2;
        c->bz = total - c->bx - c->by;
    }
    w->classified = true;
    return TR_OK;
}

static uint64_t default_contact_capacity(uint64_t n) {
    // grown on demand (overflow recovery, api.cu settle()): start near one contact per body
    return n + 32768;
}

static int free_contact_buffers(Collide* c) {
    cudaFree(c->ckeys); cudaFree(c->arank); cudaFree(c->he); cudaFree(c->dep); cudaFree(c->crec); cudaFree(c->cver);
    cudaFree(c->cgeo); cudaFree(c->frontier[0]); cudaFree(c->frontier[1]);
    c->ckeys = nullptr; c->arank = nullptr; c->he = nullptr; c->dep = nullptr; c->crec = nullptr; c->cver = nullptr;
    c->cgeo = nullptr; c->frontier[0] = c->frontier[1] = nullptr;
    c->cap_contacts = 0;
    return TR_OK;
}

static int alloc_contact_buffers(tr_world* w, uint64_t want_c) {
    Collide* c = state(w);
    if (want_c > 0xFFFFFFF0ull) want_c = 0xFFFFFFF0ull;
    free_contact_buffers(c);
    TR_TRY(realloc_dev(&c->ckeys, want_c));
    TR_TRY(realloc_dev(&c->arank, want_c));
    TR_TRY(realloc_dev(&c->he, 2 * want_c));
    TR_TRY(realloc_dev(&c->dep, want_c));
    TR_TRY(realloc_dev(&c->crec, want_c));
    TR_TRY(realloc_dev(&c->cver, want_c));
    TR_TRY(realloc_dev((ContactGeo**)&c->cgeo, want_c));
    TR_TRY(realloc_dev(&c->frontier[0], want_c));
    if (w->params.response_mode == 1) TR_TRY(realloc_dev(&c->frontier[1], want_c));
    c->cap_contacts = want_c;
    return TR_OK;
}

static int ensure_buffers(tr_world* w) {
    Collide* c = state(w);
    const uint64_t n = w->n;
    if (!c->sparse_smem_set) {   // (not at the launch: that can be inside a stream capture)
        TR_CUDA(cudaFuncSetAttribute(sparse_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SparseSmem)));
        if (SPARSE_CLUSTER > 8) TR_CUDA(cudaFuncSetAttribute(sparse_step_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        // can this device place one such cluster at all (a partitioned GPU, a GPC with fewer SMs)?  If not, every step is general.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(SPARSE_CLUSTER);
        cfg.blockDim = dim3(SPARSE_THREADS);
        cfg.dynamicSmemBytes = sizeof(SparseSmem);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = SPARSE_CLUSTER;
        attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, sparse_step_kernel, &cfg) != cudaSuccess) {
            cudaGetLastError();
            clusters = 0;
        }
        c->sparse_ok = clusters >= 1;
        c->sparse_smem_set = true;
    }
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
        TR_TRY(realloc_dev(&c->seg, cap));
        TR_TRY(realloc_dev(&c->long_list, cap));
        TR_TRY(realloc_dev(&c->huge_list, cap));
        TR_TRY(realloc_dev(&c->body_cnt, cap));
        TR_CUDA(cudaMemsetAsync(c->body_cnt, 0, cap * sizeof(uint32_t), w->stream));
        c->cap_bodies = cap;
    }
    const uint64_t ncell = 1ull << (c->bx + c->by + c->bz);
    if (ncell > c->ncell_alloc) {
        TR_TRY(realloc_dev(&c->cell_start, ncell + 1));
        TR_TRY(realloc_dev(&c->cell_count, ncell));
        // zeroed ONCE: the histogram counts itself back to zero every step (scatter_kernel)
        TR_CUDA(cudaMemsetAsync(c->cell_count, 0, ncell * sizeof(uint32_t), w->stream));
        c->ncell_alloc = ncell;
    }
    {
        const uint64_t tiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
        if (tiles > c->tile_sum_cap) {
            TR_TRY(realloc_dev(&c->tile_sum, tiles));
            c->tile_sum_cap = tiles;
        }
    }
    // Contact buffers (~100 bytes per slot) are sized on demand: the caller's max_contacts is a hard cap,
    // otherwise they start near one contact per body and grow when a step needs more (overflow recovery).
    c->cap_fixed = w->params.max_contacts != 0;
    uint64_t want_c = c->cap_contacts;
    if (c->cap_fixed) {
        want_c = w->params.max_contacts;
        if (n <= (uint64_t)TINY_MAX && want_c < 32768) want_c = 32768;   // a tiny world's list is never capped (N(N-1)/2 <= 32640 pairs)
    } else if (want_c < default_contact_capacity(n)) {
        want_c = default_contact_capacity(n);
    }
    const bool need_f1 = w->params.response_mode == 1 && !c->frontier[1];
    if (want_c != c->cap_contacts || need_f1) TR_TRY(alloc_contact_buffers(w, want_c));
    if (!c->counters) {
        TR_CUDA(cudaMalloc(&c->counters, (C_NCOUNTERS + 32) * sizeof(uint32_t)));   // + statistics of TR_FLOW_STATS builds
        TR_CUDA(cudaMemset(c->counters, 0, (C_NCOUNTERS + 32) * sizeof(uint32_t)));
    }
    return TR_OK;
}

// Grow the contact buffers after a step found `found` contacts (overflow recovery / head-room policy).
int collide_grow_contacts(tr_world* w, uint64_t found) {
    Collide* c = state(w);
    if (c->cap_fixed) {
        set_error("contact buffer overflow: %llu contacts, capacity %llu (raise tr_world_params.max_contacts, or leave it 0 to let the library grow the buffers)",
                  (unsigned long long)found, (unsigned long long)c->cap_contacts);
        return TR_ERR_CAPACITY;
    }
    uint64_t want = found + found / 2 + 32768;
    if (want <= c->cap_contacts) return TR_OK;
    return alloc_contact_buffers(w, want);
}

// A launch that may overlap the tail of the kernel before it on the stream (programmatic dependent launch; the kernel
// starts with pdl_wait()).  Captured into the step's graph as a programmatic edge.
static const bool g_no_pdl = getenv("TR_NO_PDL") != nullptr;
template <typename... KArgs, typename... Args>
static cudaError_t pdl_launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// K2-K3b on the stream (used inside the step's graph capture and by the debug entry points).
static int launch_broad(tr_world* w, cudaStream_t st) {
    Collide* c = state(w);
    const uint32_t n = (uint32_t)w->n;
    const uint32_t blocks = (n + 255) / 256;
    const uint32_t ncell = 1u << (c->bx + c->by + c->bz);
    const uint32_t ntiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
    const GridParams g = grid_params(c);
    keys_kernel<<<blocks, 256, 0, st>>>(w->center_r, w->half_ext, w->flags, n, g, c->keys, c->klass, c->side_list, c->far_list, c->far_c,
                                        c->bnd_list, c->bnd_c, c->cell_count, c->counters);
    TR_CUDA(pdl_launch(tile_sums_kernel, ntiles, SCAN_THREADS, 0, st, c->cell_count, ncell, c->tile_sum));
    TR_CUDA(pdl_launch(scan_kernel, ntiles, SCAN_THREADS, 0, st, c->cell_count, c->tile_sum, c->cell_start, ncell, c->counters, (int)C_GRID));
    TR_CUDA(pdl_launch(scatter_kernel, blocks, 256, 0, st, c->keys, ncell, w->flags, w->center_r, c->mixed ? w->half_ext : (const float4*)nullptr, n,
                       c->cell_start, c->cell_count, (SlotRec*)c->recs));
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

// K5 (+ K5b, K5c in the trailing CTAs of the same launch)
template <bool CANDIDATES>
static int launch_narrow(tr_world* w, cudaStream_t st, unsigned long long* out, uint64_t cap) {
    Collide* c = state(w);
    unsigned long long* counter = (unsigned long long*)(c->counters + (CANDIDATES ? C_CAND : C_CONTACTS));
    const uint32_t slot_blocks = (uint32_t)((w->n + NARROW_THREADS - 1) / NARROW_THREADS);
    const uint32_t outside_blocks = CANDIDATES ? 0u : (uint32_t)w->num_sms * 8u;
    const GridParams g = grid_params(c);
    OutsideArgs oa{c->side_list, c->far_list, c->far_c, c->bnd_list, c->bnd_c, w->center_r, w->half_ext, w->flags, c->klass, (uint32_t)w->n};
    const SlotRec* recs = (const SlotRec*)c->recs;
    const unsigned nblk = slot_blocks + outside_blocks;
    const bool big = w->n > (1ull << 27) || (w->params.debug_flags & TR_DEBUG_BIG_LIST);
#define TR_NARROW(M, B) TR_CUDA(pdl_launch(narrow_kernel<CANDIDATES, M, B>, nblk, NARROW_THREADS, 0, st, recs, c->cell_start, c->counters, g, slot_blocks, oa, out, (unsigned long long)cap, counter))
    if (c->mixed) {
        if (big) TR_NARROW(true, true);
        else TR_NARROW(true, false);
    } else {
        if (big) TR_NARROW(false, true);
        else TR_NARROW(false, false);
    }
#undef TR_NARROW
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

static bool executor_wide(const tr_world* w, const Collide* c) {
    // Executor width.  Deep graphs (a packed lattice: ~800 dependent hand-offs) are latency bound: single-lane
    // workers.  Wide shallow graphs are throughput bound: whole warps.  The last step the host has seen tells
    // which: contacts per lock-step iteration of its longest worker.
    if (w->params.response_mode == 2) return true;
    if (w->params.response_mode == 0 && c->prev_rounds > 0) return c->prev_contacts / c->prev_rounds > 16384;
    return false;
}

// K6 + epilogue on the stream.
static int launch_response(tr_world* w, cudaStream_t st, bool after_sparse) {
    Collide* c = state(w);
    const bool flow = w->params.response_mode != 1;   // 1 selects the level-synchronous kernel
    if (c->prep_blocks == 0) {
        int per_sm = 0;
        TR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, prep_kernel, PREP_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;               // every CTA must be resident: prep_kernel synchronises the grid itself
        c->prep_blocks = w->num_sms * per_sm;
    }
    PrepArgs pa{c->ckeys, c->cap_contacts, c->counters, w->d_ctl, c->body_cnt, c->seg, c->arank, c->he, c->long_list, c->huge_list, c->dep, c->crec,
                c->cver, (ContactGeo*)c->cgeo, flow ? (BodyRec*)c->brec : nullptr, c->frontier[0], flow ? C_QTAIL : C_FRONT,
                w->pos_m[w->cur], w->vel_rest, w->center_r, w->half_ext, w->flags, (uint32_t)w->n, after_sparse ? 1u : 0u};
    {
        // a COOPERATIVE launch: prep_kernel synchronises its grid with a counter barrier, so all its CTAs must be resident
        // at once - also when another world's kernels (other streams, other host threads) compete for the SMs, where two
        // half-resident barrier kernels would wait for each other for ever.  Captured into the step's graph like any launch.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)c->prep_blocks);
        cfg.blockDim = dim3(PREP_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TR_CUDA(cudaLaunchKernelEx(&cfg, prep_kernel, pa));
    }
    if (flow) {
        FlowArgs ra{w->pos_m[w->cur], w->vel_rest, w->pos_m[w->cur], w->vel_rest, (BodyRec*)c->brec, c->cver, (const ContactGeo*)c->cgeo};
        if (executor_wide(w, c)) {
            response_flow_kernel<32><<<w->num_sms, 128, 0, st>>>(ra, c->crec, c->dep, c->frontier[0], w->d_ctl, c->counters, 0u);
        } else {
            // 5 CTAs of 4 single-lane warps per SM (3, 4, 6, 7, 8, 9 per SM: 1.59, 1.39, 1.36, 1.38, 1.40, 1.42 ms against 1.34 on
            // the packed 1M scene: beyond ~3K workers the extra pollers cost more than the extra hands bring);
            // idle back-off 200 / 500 / 1000 / 1500 ns -> 1.34 / 1.33 / 1.38 / 1.52 ms
            response_flow_kernel<1><<<w->num_sms * 5, 128, 0, st>>>(ra, c->crec, c->dep, c->frontier[0], w->d_ctl, c->counters, 500u);
        }
        TR_CUDA(cudaGetLastError());
    } else {
        if (c->coop_blocks == 0) c->coop_blocks = w->num_sms;   // one CTA per SM: the grid barrier cost grows with the CTA count
        RespArgs ra{w->pos_m[w->cur], w->vel_rest, w->center_r, w->half_ext, w->flags};
        const uint4* cr = c->crec;
        uint32_t* dep = c->dep;
        uint32_t* f0 = c->frontier[0];
        uint32_t* f1 = c->frontier[1];
        uint32_t* cnt = c->counters;
        const uint32_t* ctl = w->d_ctl;
        void* args[] = {&ra, &cr, &dep, &f0, &f1, &cnt, &ctl};
        TR_CUDA(cudaLaunchCooperativeKernel((void*)response_kernel, dim3(c->coop_blocks), dim3(256), args, 0, st));
    }
    epilogue_kernel<<<1, 128, 0, st>>>(c->counters, w->d_ctl, w->d_ring);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

// the tail of a step the host expects to be sparse: ordering + replay + record in one cluster kernel (collide_response.cuh)
static int launch_sparse_response(tr_world* w, cudaStream_t st, bool tail_follows) {
    Collide* c = state(w);
    SparseArgs sa{c->ckeys, c->cap_contacts, c->counters, w->d_ctl, w->d_ring, w->pos_m[w->cur], w->vel_rest,
                  w->center_r, w->half_ext, w->flags, (uint32_t)w->n, tail_follows ? 1u : 0u};
    TR_CUDA(pdl_launch(sparse_step_kernel, SPARSE_CLUSTER, SPARSE_THREADS, sizeof(SparseSmem), st, sa));
    return TR_OK;
}

// How a step's graph ends after the narrow phase:
//   STEP_GENERAL  prep_kernel -> executor -> epilogue_kernel                       (any step)
//   STEP_SPARSE   sparse_step_kernel alone                                         (a step the host is confident is sparse;
//                 if it is not, the kernel flags it and the host re-runs it: settle() in api.cu)
//   STEP_TRY      sparse_step_kernel -> prep_kernel -> executor -> epilogue_kernel (nothing recent known: decided on the
//                 device - the general kernels return at once when the sparse-step kernel has done the step)
enum StepVariant { STEP_GENERAL = 0, STEP_SPARSE = 1, STEP_TRY = 2 };

static int launch_step(tr_world* w, cudaStream_t st, StepVariant v) {
    Collide* c = state(w);
    TR_TRY(launch_broad(w, st));
    TR_TRY(launch_narrow<false>(w, st, c->ckeys, c->cap_contacts));
    if (v != STEP_GENERAL) TR_TRY(launch_sparse_response(w, st, v == STEP_TRY));
    if (v != STEP_SPARSE) TR_TRY(launch_response(w, st, v == STEP_TRY));
    return TR_OK;
}

// Which variant for the step about to be enqueued?  The newest record the device has written so far tells (mapped pinned
// memory, read without any synchronisation: it is only a hint - the kernels check).
constexpr uint64_t SPARSE_CONFIDENT = SPARSE_MAX_CONTACTS * 7 / 8;   // head room for the count to drift before the kernel's own limit
constexpr uint64_t SPARSE_HOPELESS = SPARSE_MAX_CONTACTS * 4;
static StepVariant predict_variant(tr_world* w, Collide* c) {
    static const bool off = getenv("TR_NO_SPARSE") != nullptr;
    // (sharded worlds replicate the collision step: every rank must take the same decisions, and a peek is a matter of timing)
    if (off || !c->sparse_ok || w->params.response_mode != 0 || (w->params.debug_flags & TR_DEBUG_NO_SPARSE) || w->nranks != 1) return STEP_GENERAL;
    if (w->h_ring) {
        if (c->peek_seq < w->ring_consumed) c->peek_seq = w->ring_consumed;
        while (c->peek_seq < w->coll_seq) {
            const volatile StepRecord* r = &w->h_ring[c->peek_seq % STEP_RING];
            if (r->seq != c->peek_seq + 1u) break;
            if (!(r->flags & 5u)) {   // (not: overflowed / no-op steps)
                c->known_contacts = ((uint64_t)r->contacts_hi << 32) | r->contacts_lo;
                c->known_valid = true;
            }
            ++c->peek_seq;
        }
    }
    if (!c->known_valid) return STEP_TRY;
    if (c->known_contacts > SPARSE_HOPELESS) return STEP_GENERAL;
    if (c->sparse_block) {            // a recent miss the count does not explain (a hub, a lopsided partition): no confidence for a while
        --c->sparse_block;
        return STEP_TRY;
    }
    return c->known_contacts <= SPARSE_CONFIDENT ? STEP_SPARSE : STEP_TRY;
}

static uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return h;
}

// everything the captured launches depend on
static uint64_t step_signature(tr_world* w) {
    Collide* c = state(w);
    uint64_t h = 0x5452ull;
    const void* ptrs[] = {w->center_r, w->half_ext, w->flags, w->pos_m[w->cur], w->vel_rest, c->keys, c->recs, c->cell_count, c->cell_start,
                          c->ckeys, c->he, c->brec, c->counters, w->d_ctl, w->d_ring, c->frontier[1]};
    for (const void* p : ptrs) h = mix64(h, (uint64_t)(uintptr_t)p);
    h = mix64(h, w->n);
    h = mix64(h, c->cap_contacts);
    h = mix64(h, (uint64_t)(c->bx | (c->by << 5) | (c->bz << 10)) | ((uint64_t)(c->mixed ? 1 : 0) << 20) |
                     ((uint64_t)w->params.response_mode << 24) | ((uint64_t)(executor_wide(w, c) ? 1 : 0) << 28) |
                     ((uint64_t)(w->params.debug_flags & 0xFFu) << 32));
    uint32_t f[5];
    memcpy(&f[0], &c->inv_cell, 4);
    memcpy(&f[1], &c->rmax_grid, 4);
    memcpy(&f[2], c->origin, 12);
    for (uint32_t v : f) h = mix64(h, v);
    return h | 1ull;
}

void collide_poll_stats(tr_world*) {}

// Whole steps of a small world in one kernel (tiny.cuh).  integrate: the step's O(N) part runs here too.
static bool tiny_eligible(const tr_world* w) {
    static const bool off = getenv("TR_NO_TINY") != nullptr;
    // one thread replays the contacts of a small world sequentially (~80 ns each): fine for the handful of contacts such a
    // world normally has, slower than the grid pipeline's parallel executor for a dense clump - the last step the host
    // has seen decides
    const bool clump = w->col && w->col->last_contacts > 128;
    return !off && !(w->params.debug_flags & TR_DEBUG_NO_TINY) && w->n <= (uint64_t)TINY_MAX && w->nranks == 1 &&
           (w->params.flags & TR_PAIR_GRAVITY) == 0 && !clump;
}

int tiny_steps(tr_world* w, const StepConsts& sc, int nsteps, bool* taken) {
    *taken = false;
    if (!tiny_eligible(w) || nsteps <= 0) return TR_OK;
    const bool coll = (w->params.flags & TR_COLLISIONS) != 0;
    if (coll) {
        if (!w->classified) TR_TRY(collide_classify(w));   // keeps tr_debug_grid usable; the tiny step itself has no grid
        TR_TRY(ensure_buffers(w));
    }
    Collide* c = coll ? state(w) : nullptr;
    TinyArgs a{};
    a.pos_m = w->pos_m[w->cur];
    a.vel_rest = w->vel_rest;
    a.force = w->force;
    a.drag = w->drag;
    a.center_r = w->center_r;
    a.half_ext = w->half_ext;
    a.flags = w->flags;
    a.n = (uint32_t)w->n;
    a.c = sc;
    a.collide = coll ? 1 : 0;
    a.ckeys = c ? c->ckeys : nullptr;
    a.counters = c ? c->counters : nullptr;
    a.ctl = w->d_ctl;
    a.ring = w->d_ring;
    a.traj = w->traj;
    a.traj_cnt = w->traj_cnt;
    a.traj_max = w->traj ? w->traj_max : 0;
    a.traj_stride = w->traj_stride;
    a.frame_out = w->frame_request ? w->d_frame : nullptr;
    w->frame_ready = w->frame_request;
    tiny_step_kernel<<<1, TINY_MAX, 0, w->stream>>>(a, nsteps);
    TR_CUDA(cudaGetLastError());
    w->stats.kernel_launches += 1;
    if (coll) {
        w->coll_seq += 1;
        c->host_contacts_valid = false;
    }
    *taken = true;
    return TR_OK;
}

// One collision step, enqueued: detection + ordered response + the step's record.  Nothing here waits for the GPU.
int collide_step(tr_world* w) {
    if (w->n == 0) return TR_OK;
    if (!w->classified) TR_TRY(collide_classify(w));
    TR_TRY(ensure_buffers(w));
    Collide* c = state(w);
    static const bool no_graph = getenv("TR_NO_GRAPH") != nullptr;
    // the level-synchronous executor is a cooperative launch, and the legacy default stream cannot be captured: plain launches
    const StepVariant variant = predict_variant(w, c);
    if (no_graph || w->stream == nullptr || w->params.response_mode == 1) {
        TR_TRY(launch_step(w, w->stream, variant));
    } else {
        const int slot = 3 * w->cur + (int)variant;
        const uint64_t sig = step_signature(w);
        if (!c->graph[slot] || c->graph_sig[slot] != sig) {
            if (c->graph[slot]) cudaGraphExecDestroy(c->graph[slot]);
            c->graph[slot] = nullptr;
            cudaGraph_t graph = nullptr;
            TR_CUDA(cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal));
            int rc = launch_step(w, w->stream, variant);
            cudaError_t e = cudaStreamEndCapture(w->stream, &graph);
            if (rc != TR_OK) return rc;
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
            e = cudaGraphInstantiate(&c->graph[slot], graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
            c->graph_sig[slot] = sig;
        }
        TR_CUDA(cudaGraphLaunch(c->graph[slot], w->stream));
    }
    w->stats.kernel_launches += variant == STEP_SPARSE ? 6 : variant == STEP_TRY ? 9 : 8;   // keys, tile sums, scan, scatter, narrow (+ side / far workers), then prep, executor, epilogue - or the sparse step's one kernel
    w->coll_seq += 1;
    c->host_contacts_valid = false;
    return TR_OK;
}

// Fold the step records the device has written since the last call into the host-side statistics
// (after a stream synchronisation).  *overflow_seq = sequence number of a step whose contact list
// did not fit (its response and every later step were skipped on the device), *found = its count.
int collide_consume_records(tr_world* w, bool* overflow, uint32_t* overflow_seq, uint64_t* found, bool* not_sparse) {
    *overflow = false;
    *not_sparse = false;
    Collide* c = w->col;
    if (!c || !w->h_ring) return TR_OK;
    const uint32_t upto = w->coll_seq;
    {
        // a step that overflowed leaves its record in the slot behind the ring, where no later step overwrites it
        StepRecord& note = w->h_ring[STEP_RING];
        if (note.seq != 0) {
            *overflow = true;
            *overflow_seq = note.seq - 1u;
            *found = ((uint64_t)note.contacts_hi << 32) | note.contacts_lo;
            if (note.flags & 8u) {
                // a step predicted sparse that was not: no capacity problem.  Re-run it through the general kernels and
                // keep the prediction off for a while, twice as long after every miss in a row
                *not_sparse = true;
                if (*found <= SPARSE_MAX_CONTACTS) {   // (too many contacts: the count itself, now known, stops the prediction)
                    c->sparse_backoff = c->sparse_backoff >= 2048 ? 4096 : c->sparse_backoff * 2;
                    c->sparse_block = c->sparse_backoff;
                }
                w->stats.sparse_bailouts += 1;
            }
            note.seq = 0;
        }
    }
    bool fresh = false;
    if (upto - w->ring_consumed > STEP_RING) w->ring_consumed = upto - STEP_RING;   // older records were overwritten
    for (uint32_t s = w->ring_consumed; s < upto; ++s) {
        const StepRecord r = w->h_ring[s % STEP_RING];
        if (r.seq != s + 1u) continue;            // not written (cannot happen after a sync) or overwritten
        if (r.flags & 4u) continue;               // no-op step behind an overflow: re-run by the recovery
        const uint64_t cnt = ((uint64_t)r.contacts_hi << 32) | r.contacts_lo;
        if (r.flags & 1u) continue;               // reported through the notice slot above
        if (r.flags & 2u) {
            if (w->deferred_error == TR_OK) {
                w->deferred_error = TR_ERR_CUDA;
                w->deferred_msg = "ordered response: dataflow watchdog fired (dependency graph did not drain)";
            }
        }
        fresh = true;
        if (r.flags & 16u) {
            w->stats.sparse_steps += 1;
            c->sparse_backoff = 32;
        }
        w->stats.last_contacts = cnt;
        w->stats.side_list_size = r.side + r.far;   // everything outside the grid: oversize list + far list
        w->stats.last_response_rounds = r.rounds;
        c->last_contacts = cnt;
        c->last_side = r.side;
        c->last_far = r.far;
        if (cnt) {
            c->prev_contacts = cnt;
            c->prev_rounds = r.rounds;
        }
        if (w->timing && r.t[3] >= r.t[0]) {
            w->stats.broadphase_ms += (double)(r.t[1] - r.t[0]) * 1e-6;
            w->stats.narrowphase_ms += (double)(r.t[2] - r.t[1]) * 1e-6;
            w->stats.response_ms += (double)(r.t[3] - r.t[2]) * 1e-6;
            w->stats.broadphase_launches += 1;
        }
    }
    w->ring_consumed = upto;
    c->peek_seq = upto;
    if (*overflow) {
        c->known_contacts = *found;
        c->known_valid = true;
    } else if (fresh) {
        c->known_contacts = c->last_contacts;
        c->known_valid = true;
    }
    return TR_OK;
}

// Head-room policy, applied by the host after a synchronisation: grow ahead of a slowly filling list so that the
// recovery path stays the exception.
static int cache_host_contacts(tr_world* w);

int collide_headroom(tr_world* w) {
    Collide* c = w->col;
    if (!c || c->cap_fixed || c->cap_contacts == 0) return TR_OK;
    if (c->last_contacts * 10 > c->cap_contacts * 7) {
        TR_TRY(cache_host_contacts(w));   // the last step's list lives in the buffers about to be replaced
        return collide_grow_contacts(w, c->last_contacts * 2);
    }
    return TR_OK;
}

// the device keeps the list in arrival order; the reference's visiting order is lexicographic (i,j)
static int cache_host_contacts(tr_world* w) {
    Collide* c = w->col;
    if (c->host_contacts_valid || c->last_contacts == 0) return TR_OK;
    c->host_contacts.resize(c->last_contacts);
    TR_CUDA(cudaMemcpyAsync(c->host_contacts.data(), c->ckeys, c->last_contacts * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    TR_CUDA(cudaStreamSynchronize(w->stream));
    std::sort(c->host_contacts.begin(), c->host_contacts.end());
    c->host_contacts_valid = true;
    return TR_OK;
}

int collide_read_contacts(tr_world* w, int32_t* pairs, uint64_t cap, uint64_t* n) {
    Collide* c = w->col;
    uint64_t cnt = c ? c->last_contacts : 0;
    if (n) *n = cnt;
    if (!pairs || cnt == 0) return TR_OK;
    TR_TRY(cache_host_contacts(w));
    const uint64_t take = cnt < cap ? cnt : cap;
    for (uint64_t k = 0; k < take; ++k) {
        pairs[2 * k] = (int32_t)(c->host_contacts[k] >> 32);
        pairs[2 * k + 1] = (int32_t)(c->host_contacts[k] & 0xFFFFFFFFu);
    }
    return TR_OK;
}

// debug entry points: a stand-alone broad phase (the step's counters are zero between steps; left zero again)
static int debug_begin(tr_world* w) {
    if (!w->classified) TR_TRY(collide_classify(w));
    TR_TRY(ensure_buffers(w));
    Collide* c = state(w);
    TR_CUDA(cudaMemsetAsync(c->counters, 0, C_NCOUNTERS * sizeof(uint32_t), w->stream));
    TR_TRY(launch_broad(w, w->stream));
    return TR_OK;
}

static int debug_end(tr_world* w, uint32_t* host_counters) {
    Collide* c = state(w);
    TR_CUDA(cudaMemcpyAsync(host_counters, c->counters, C_NHOST * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
    TR_CUDA(cudaMemsetAsync(c->counters, 0, C_NCOUNTERS * sizeof(uint32_t), w->stream));
    TR_CUDA(cudaStreamSynchronize(w->stream));
    return TR_OK;
}

int collide_debug_grid(tr_world* w, uint32_t* keys, uint32_t* sorted_ids, uint64_t* m, float* cell, float* inv_cell,
                       int32_t* bits) {
    if (w->n == 0) {
        if (m) *m = 0;
        return TR_OK;
    }
    TR_TRY(debug_begin(w));
    Collide* c = state(w);
    // K4, only here: the stable (key,id) order of the grid bodies (crowded cells: listed in long_list, sorted by CTAs with
    // the half-edge array - idle between steps, at least as long as the record array - as scratch)
    canon_kernel<<<(uint32_t)((w->n + 255) / 256), 256, 0, w->stream>>>((const SlotRec*)c->recs, c->cell_start, grid_params(c),
                                                                          c->counters, c->sorted_ids, c->long_list);
    canon_big_kernel<<<w->num_sms, 256, 0, w->stream>>>((const SlotRec*)c->recs, c->cell_start, c->counters, c->long_list, c->sorted_ids,
                                                        c->he);
    TR_CUDA(cudaGetLastError());
    uint32_t hc[C_NHOST];
    TR_TRY(debug_end(w, hc));
    const uint32_t gm = hc[C_GRID];
    if (m) *m = gm;
    if (cell) *cell = c->cell;
    if (inv_cell) *inv_cell = c->inv_cell;
    // one number when the axes agree, else bx | by << 8 | bz << 16 (the oracle's grid functions take the same code)
    if (bits) *bits = (c->bx == c->by && c->by == c->bz) ? c->bx : (c->bx | (c->by << 8) | (c->bz << 16));
    const uint32_t sentinel = 1u << (c->bx + c->by + c->bz);
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
    unsigned long long* d = nullptr;
    uint64_t dcap = pairs ? cap : 0;
    if (dcap) TR_CUDA(cudaMalloc(&d, dcap * sizeof(unsigned long long)));
    int rc = debug_begin(w);
    if (rc == TR_OK) rc = launch_narrow<true>(w, w->stream, d, dcap);
    uint32_t hc[C_NHOST];
    if (rc == TR_OK) rc = debug_end(w, hc);
    const uint64_t count = rc == TR_OK ? (((uint64_t)hc[C_CAND + 1] << 32) | hc[C_CAND]) : 0;
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
    cudaFree(c->seg); cudaFree(c->long_list); cudaFree(c->huge_list); cudaFree(c->body_cnt); cudaFree(c->counters); cudaFree(c->brec);
    free_contact_buffers(c);
    for (int k = 0; k < 6; ++k)
        if (c->graph[k]) cudaGraphExecDestroy(c->graph[k]);
    delete c;
    w->col = nullptr;
    return TR_OK;
}

}  // namespace tr

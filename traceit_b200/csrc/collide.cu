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
#include "collide_common.cuh"
#include "collide_broad.cuh"
#include "collide_response.cuh"

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
    if (p.grid_bits != 0) {
        c->bx = c->by = c->bz = p.grid_bits;       // the caller's choice: the same on every axis
    } else {
        // smallest table with at least 2 cells per body, bits spread over the axes (x, the fastest key
        // digit, first): with the same bits on every axis the table could only grow in steps of 8x
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
    const uint64_t ncell = 1ull << (c->bx + c->by + c->bz);
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
    const uint32_t ncell = 1u << (c->bx + c->by + c->bz);
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
                           ((uint64_t)n << 16) | ((uint64_t)(c->mixed ? 1 : 0) << 15) | (uint64_t)(c->bx | (c->by << 5) | (c->bz << 10)), 0, 0};
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
        // single-lane workers, 5 CTAs of 4 warps per SM (3, 4, 6, 7, 8, 9 per SM: 1.59, 1.39, 1.36, 1.38, 1.40, 1.42 ms
        // against 1.34: beyond ~3K workers the extra pollers cost more than the extra hands bring).
        // Wide shallow graphs are throughput bound: whole warps.  The previous step tells which:
        // contacts per lock-step iteration of its longest worker.
        bool wide = w->params.response_mode == 2;
        if (w->params.response_mode == 0 && c->prev_rounds > 0) wide = c->prev_contacts / c->prev_rounds > 16384;
        if (wide) {
            response_flow_kernel<32><<<fblocks, 128, 0, w->stream>>>(ra, c->crec, c->dep, c->frontier[0], cn, c->counters, 0u);
        } else {
            int sblocks = w->num_sms * 5;
            const int swant = (int)((cn + 3) / 4);
            if (swant < sblocks) sblocks = swant < 1 ? 1 : swant;
            // idle back-off: 200 / 500 / 1000 / 1500 ns -> 1.34 / 1.33 / 1.38 / 1.52 ms
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


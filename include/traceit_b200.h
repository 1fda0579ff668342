/* traceit_b200.h - C ABI of the B200-native TraceIt physics step.
 *
 * The reference (zxcfghjkl/TraceIt, one C++17 translation unit) has no plugin or FFI
 * layer.  Its de-facto boundary for the physics path is
 *     void updatePhysics(world& w, ofstream& coords)     reference src/main.cpp:748
 *         called once per frame at                       reference src/main.cpp:1589
 *     w.objects.push_back(newObj)  (body creation)       reference src/main.cpp:1242,1491
 *     world{atmDensity,g,...} plain fields               reference src/main.cpp:118-126
 * Every entry point below names the reference interface it replaces.  All functions
 * return 0 (TR_OK) or a TR_ERR_* code, never throw, and take only plain pointers and
 * sizes.  The C++ shim that restores the reference signature on top of this ABI is
 * traceit_b200/host/traceit_host.hpp; the binding a maintainer adds is in INTEGRATION.md.
 *
 * There is NO CPU fallback: every compute entry point fails with TR_ERR_NO_DEVICE /
 * TR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef TRACEIT_B200_H
#define TRACEIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TR_ABI_VERSION 1

/* status codes */
#define TR_OK 0
#define TR_ERR_INVALID_ARG 1
#define TR_ERR_NO_DEVICE 2   /* no CUDA device / not an sm_100 part                      */
#define TR_ERR_CUDA 3        /* a CUDA runtime call or kernel failed; see tr_last_error  */
#define TR_ERR_NCCL 4        /* NCCL could not be loaded or a collective failed          */
#define TR_ERR_CAPACITY 5    /* contact buffer overflow: raise tr_world_params.max_contacts */
#define TR_ERR_DOMAIN 6      /* reserved (out-of-domain bodies are handled by the side list) */
#define TR_ERR_STATE 7       /* call not valid in the world's current state              */
#define TR_ERR_NOMEM 8

/* tr_world_params.flags */
#define TR_PAIR_GRAVITY 1u   /* Newtonian all-pairs term, reference src/main.cpp:679-693
                                (dead code there, hence OFF by default)                 */
#define TR_COLLISIONS 2u     /* resolveCollisions sweep, reference src/main.cpp:673-745  */

typedef struct tr_world tr_world; /* opaque; owns all device buffers, streams, comms */

/* World parameters.  Defaults (tr_default_params) equal the reference's constants:
 *   g = {0,(float)-9.8,0}   hard-coded local, src/main.cpp:749  (world::g is never read)
 *   G = 6.6743e-11f         src/main.cpp:685
 *   dt = 0.01666f           src/main.cpp:750
 *   atm_density = 1.2f      world::atmDensity, UI default src/main.cpp:1141
 *   flags = TR_COLLISIONS   (what the live reference step does)                        */
typedef struct tr_world_params {
    float g[3];
    float G;
    float dt;
    float atm_density;
    uint32_t flags;
    /* Uniform-grid broad phase (no reference counterpart; the reference is brute force).
     * 0 / zeros = automatic.  cell edge defaults to 2*rmax*(1+2^-7) where rmax is the
     * largest radius among grid bodies; grid_bits is bits per axis of the wrapped key
     * (2..10, the same on x, y, z); a grid_cell below 2*rmax*(1+2^-7) would make the 27-cell
     * stencil miss contacts and is raised to that; automatic bits = the smallest table with 2 cells per body,
     * its bits spread over the axes (x first).  Spheres and boxes share the grid (a body's
     * extent is |radius|, and max|size| as soon as boxes exist); oversize bodies go to a
     * side list tested against every body, bodies outside the grid domain to a far list
     * tested against each other and the grid bodies next to the domain boundary.           */
    float grid_origin[3];
    float grid_cell;
    int32_t grid_bits;
    uint64_t max_contacts;       /* 0 = automatic: the contact buffers (~100 bytes per slot) start at
                                    N + 32768 slots and grow when a step finds more (the step is
                                    completed with the larger buffers, nothing is lost).  Non-zero =
                                    hard cap: a step that finds more reports TR_ERR_CAPACITY at the
                                    next synchronisation and leaves the world as that step's
                                    integration left it (no response applied, later steps not run). */
    /* Largest extent admitted to the grid; 0 = automatic: the largest extent that is
     * <= 8x the median extent.                                           */
    float grid_max_radius;
    /* Ordered collision response: 0 = dataflow (barrier-free, default; single-lane workers for
     * deep dependency graphs, whole warps for wide shallow ones, chosen from the previous step),
     * 1 = level-synchronous wavefronts (one grid barrier per wavefront), 2 / 3 = dataflow forced
     * to whole warps / single-lane workers.  All replay the reference's Gauss-Seidel order exactly
     * and give identical bits; see DESIGN.md K6.                                          */
    uint32_t response_mode;
    /* Test / debugging knobs, 0 in production.  TR_DEBUG_NO_TINY: worlds of up to 256 bodies normally
     * run whole steps in ONE single-CTA kernel (the reference's own algorithm, detection spread over the
     * CTA's threads; see DESIGN.md "small worlds"); the flag sends them through the general grid pipeline.
     * TR_DEBUG_BIG_LIST, TR_DEBUG_NO_SPARSE: see below.                                          */
    uint32_t debug_flags;
    uint32_t reserved;
} tr_world_params;
#define TR_DEBUG_NO_TINY 1u
#define TR_DEBUG_BIG_LIST 2u /* narrow phase: use the list format of worlds with more than 2^27 bodies (so that it can be tested) */
/* A step the host expects to have few (<= 3072) contacts normally ends in ONE kernel of one thread-block cluster that
 * orders and replays the contacts in distributed shared memory (response_mode 0 only).  A wrong guess is
 * caught on the device before anything changes and the step is re-run through the general kernels at the next
 * synchronisation - invisible to the caller except in tr_stats.sparse_bailouts.  The flag keeps every step general. */
#define TR_DEBUG_NO_SPARSE 4u

/* Body descriptor = hot fields of the reference's `object` + `collider`
 * (src/main.cpp:79-105), same names and meaning.  The UI-side parameter contract
 * (mass [1,1000], Cd [0,2] default 0.42, size [1,10], rest 0.5; src/main.cpp:1216-1224,
 * 1302-1307) is NOT enforced here, exactly as push_back does not enforce it.          */
typedef struct tr_body_desc {
    float pos[3];
    float vel[3];
    float force[3];     /* external force, consumed and zeroed by each step (main.cpp:782,811) */
    float mass;
    float Cd;
    float area;
    int32_t stationary; /* object::stationary: skipped by the integrator (main.cpp:764)  */
    float center[3];    /* collider::center (refreshed from pos only when !stationary)   */
    float size[3];      /* collider::size, AABB half-extents as used (main.cpp:648-652)  */
    float radius;       /* collider::radius                                              */
    int32_t isSphere;
    int32_t isStatic;   /* collider::isStatic: immune to impulses/nudges (main.cpp:725-741) */
    float rest;         /* collider::rest                                                */
} tr_body_desc;         /* 92 bytes */

typedef struct tr_stats {
    uint64_t steps;              /* steps executed since creation / last reset           */
    uint64_t kernel_launches;    /* launches of THIS library's kernels                   */
    uint64_t last_contacts;      /* contacts found by the last step                      */
    uint64_t last_candidates;    /* candidate pairs tested by the last step (debug mode) */
    uint32_t last_response_rounds; /* last ordered response: wavefronts (level-synchronous)
                                      or iterations of the longest worker (dataflow)      */
    uint32_t side_list_size;     /* bodies outside the grid at the last step: oversize /
                                    non-finite extent (side list) + outside the domain (far list) */
    double gravity_ms;           /* accumulated device time of the gravity kernel        */
    uint64_t gravity_launches;
    double broadphase_ms;        /* keys + histogram, cell-table scan, record scatter (K2-K3b) */
    uint64_t broadphase_launches;/* number of timed broad-phase passes (one per step)    */
    double narrowphase_ms;       /* K5                                                   */
    double response_ms;          /* contact sort + dependency build + K6                 */
    double integrate_ms;         /* standalone integrator (pair gravity off)             */
    double allgather_ms;         /* NCCL all-gather (sharded worlds)                     */
    uint64_t interactions;       /* ordered pair interactions evaluated (N_targets * N)  */
    uint64_t sparse_steps;       /* collision steps replayed by the one-cluster kernel for sparse steps */
    uint64_t sparse_bailouts;    /* steps predicted sparse that were not (re-run through the general kernels) */
} tr_stats;

/* ---- library --------------------------------------------------------------------- */
int tr_abi_version(void);
const char* tr_last_error(void);                 /* thread-local message of the last failure */
int tr_device_count(int* count);
int tr_default_params(tr_world_params* p);

/* ---- world lifetime: replaces `world w;` (src/main.cpp:1479-1481) ----------------- */
int tr_world_create(const tr_world_params* p, int device, tr_world** out);
/* Sharded world, one process per GPU: rank owns a contiguous block of target bodies,
 * holds every source position, and all-gathers float4(x,y,z,m) each step over NCCL.
 * unique_id: 128 bytes from tr_comm_unique_id on rank 0, distributed by the caller.    */
int tr_comm_unique_id(void* out128);
int tr_world_create_sharded(const tr_world_params* p, int device, int rank, int nranks,
                            const void* unique_id128, tr_world** out);
int tr_world_destroy(tr_world* w);
int tr_world_set_params(tr_world* w, const tr_world_params* p);   /* world field writes  */
int tr_world_get_params(tr_world* w, tr_world_params* p);
/* Run on the caller's CUDA stream (cudaStream_t; NULL = the legacy default stream)
 * instead of the world's own non-blocking stream.                                      */
int tr_world_set_stream(tr_world* w, void* cuda_stream);

/* ---- body creation: replaces w.objects.push_back(obj) (src/main.cpp:1242,1491) ----- */
int tr_body_add(tr_world* w, const tr_body_desc* b, uint32_t* id);
int tr_bodies_add(tr_world* w, uint64_t n, const tr_body_desc* b, uint32_t* first_id);
int tr_world_size(tr_world* w, uint64_t* n);
int tr_world_owned_range(tr_world* w, uint64_t* first, uint64_t* count); /* sharded worlds */
/* Creation helpers of the reference UI path, in its own mixed double/float arithmetic:
 * toSpherical (src/main.cpp:1080-1089) and area = M_PI*size*size (src/main.cpp:1295).  */
int tr_to_spherical(float elev_deg, float azim_deg, float magnitude, float out_vel[3]);
float tr_area_from_size(float size);

/* ---- the step: replaces updatePhysics(w, coords) (src/main.cpp:748, call site 1589) - */
int tr_step(tr_world* w, int nsteps);            /* asynchronous on the world's stream: nothing in a
                                                    step waits for the host, collisions included */
int tr_sync(tr_world* w);                        /* wait; reports deferred device errors; completes a
                                                    step whose contact list outgrew its buffers */
/* Host-buffer round trip used by the C++ shim: upload bodies[0..n) (AoS), run nsteps,
 * read them back.  n must equal the world size (bodies are created on first use); on a
 * sharded world n is this rank's owned count.  When the SAME buffer is passed on
 * consecutive calls (the per-frame pattern) the library page-locks it in place from the
 * second call on, until another buffer is passed or the world is destroyed: keep it
 * allocated for that long.                                                              */
int tr_step_host(tr_world* w, uint64_t n, tr_body_desc* bodies_inout, int nsteps);

/* Per-frame call for a DEVICE-RESIDENT world (the renderer hand-off, src/main.cpp:1589 then
 * drawWorld): upload the step's only per-frame input - the external forces obj.force, which
 * the step consumes and zeroes (main.cpp:782,811); 3 floats per body, NULL = leave them - run
 * nsteps, and read back exactly what a step changes: pos, vel and the collider centre
 * (main.cpp:787-791, 725-741).  12 bytes up and 36 bytes down per body instead of the 92-byte
 * descriptor each way of tr_step_host.  n = world size (owned count on a sharded world).
 * Everything else about a body changes only through tr_write_bodies / tr_bodies_add.  The same
 * page-locking of reused caller buffers as tr_step_host applies to force_xyz and out.        */
typedef struct tr_frame_body {
    float pos[3];
    float vel[3];
    float center[3];
} tr_frame_body;        /* 36 bytes */
int tr_step_frame(tr_world* w, uint64_t n, const float* force_xyz, tr_frame_body* out, int nsteps);

/* ---- state access: replaces direct reads/writes of w.objects[i] -------------------- */
int tr_read_state(tr_world* w, uint64_t first, uint64_t count, float* pos_xyz, float* vel_xyz);
int tr_read_bodies(tr_world* w, uint64_t first, uint64_t count, tr_body_desc* out);
int tr_write_bodies(tr_world* w, uint64_t first, uint64_t count, const tr_body_desc* in);
int tr_set_force(tr_world* w, uint32_t id, float fx, float fy, float fz); /* obj.force = */

/* ---- trajectory history: replaces object::trajectory (src/main.cpp:101-103,793-808) ------- */
/* Keep the last max_points post-integration positions of every non-stationary body in a
 * device ring (the reference keeps maxTrajectoryPoints = 300 and pushes obj.pos right after
 * integration, before the collision nudges).  0 disables and frees the ring.  Costs
 * 12*max_points bytes per body.  tr_read_trajectory returns the points oldest first.      */
int tr_world_enable_trajectory(tr_world* w, uint32_t max_points);
int tr_read_trajectory(tr_world* w, uint32_t id, float* xyz, uint64_t cap_points, uint64_t* n_points);

/* ---- parity / observability --------------------------------------------------------- */
/* Contact set of the last step in lexicographic (i,j) order, i<j: the pairs for which
 * the reference's guards + predicate (src/main.cpp:695-707) hold.  pairs may be NULL.  */
int tr_read_contacts(tr_world* w, int32_t* pairs_ij, uint64_t cap, uint64_t* n);
/* Broad-phase internals of the last step (each pointer may be NULL):
 *   keys[N]        wrapped cell key per body id (0xFFFFFFFF for side-list bodies)
 *   sorted_ids[M]  grid body ids in (key,id) order;  *m receives M
 *   cell, inv_cell, bits: the grid actually used; bits = b when every axis has b bits,
 *                  else bx | by << 8 | bz << 16.                                        */
int tr_debug_grid(tr_world* w, uint32_t* keys, uint32_t* sorted_ids, uint64_t* m,
                  float* cell, float* inv_cell, int32_t* bits);
/* Re-run the last step's broad phase emitting every candidate pair (i<j) of the grid
 * (not the side list), lexicographically sorted.  pairs may be NULL to only count.     */
int tr_debug_candidates(tr_world* w, int32_t* pairs_ij, uint64_t cap, uint64_t* n);
/* Pair-gravity force per body for the CURRENT positions (3 floats each, owned range).  */
int tr_debug_gravity(tr_world* w, float* force_xyz);
int tr_get_stats(tr_world* w, tr_stats* s);
int tr_reset_stats(tr_world* w);
int tr_enable_timing(tr_world* w, int on);       /* per-kernel CUDA-event timing          */

/* FP32 FMA-pipe peak of `device`, measured with an FFMA-only kernel (lane-ops/s where a
 * lane-op is one fp32 FMA-pipe operation on one lane) plus the SM clock it ran at.      */
int tr_measure_fma_peak(int device, double* lane_ops_per_s, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* TRACEIT_B200_H */

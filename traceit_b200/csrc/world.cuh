// Internal definitions shared by the translation units of libtraceit_b200.so.
// Not part of the ABI (include/traceit_b200.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/traceit_b200.h"

namespace tr {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TR_CUDA(call)                                                            \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) return ::tr::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define TR_TRY(call)                 \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != TR_OK) return rc__; \
    } while (0)

// ---- body flags (SoA byte) ------------------------------------------------------------
enum : uint8_t {
    F_STATIONARY = 1,  // object::stationary      reference src/main.cpp:98
    F_STATIC = 2,      // collider::isStatic      reference src/main.cpp:84
    F_SPHERE = 4,      // collider::isSphere      reference src/main.cpp:83
    F_LIVE = 8         // !(mass <= 0): can take part in collisions (reference src/main.cpp:696)
};

// ---- gravity tiling constants (see DESIGN.md "K1") -------------------------------------
constexpr int GRAV_THREADS = 256;
#ifndef TR_GRAV_PAIRS
#define TR_GRAV_PAIRS 4
#endif
#ifndef TR_GRAV_MINBLOCKS
#define TR_GRAV_MINBLOCKS 2
#endif
constexpr int GRAV_PAIRS = TR_GRAV_PAIRS;             // packed f32x2 target pairs per thread
constexpr int GRAV_T = 2 * GRAV_PAIRS;                 // targets per thread
constexpr int GRAV_TILE = GRAV_THREADS * GRAV_T;       // 2048 targets per tile
#ifndef TR_GRAV_SUB
#define TR_GRAV_SUB 512
#endif
#ifndef TR_GRAV_STAGES
#define TR_GRAV_STAGES 3
#endif
constexpr int GRAV_SUB = TR_GRAV_SUB;                  // sources per TMA stage (16 B each)
constexpr int GRAV_STAGES = TR_GRAV_STAGES;
// Partial sums per target = source chunks per tile.  74 = half the B200's 148 SMs: with 2 persistent CTAs per SM
// (296) a target block of T tiles is T * 74 work items = T / 4 items per CTA, a whole number whenever T is a
// multiple of 4 - 4M bodies over 8 GPUs: 256 tiles x 74 = 64 rounds exactly, where 64 chunks gave 55.35 rounds
// executed as 56 (round 1: the 8-GPU limiter).  A function of N only, like the chunk length below.
constexpr int GRAV_MAX_CHUNKS = 74;

// chunk length is a function of N only, so a sharded run adds the same partial sums in
// the same order as an unsharded one (bit-identical for any rank count).
inline uint32_t grav_chunk_len(uint64_t n) {
    uint64_t c = (n + GRAV_MAX_CHUNKS - 1) / GRAV_MAX_CHUNKS;
    c = (c + GRAV_SUB - 1) / GRAV_SUB * GRAV_SUB;
    if (c < (uint64_t)GRAV_SUB) c = GRAV_SUB;   // at least one full TMA stage per chunk
    return (uint32_t)c;
}

// ---- device-side step bookkeeping --------------------------------------------------------
// Nothing in a step waits for the host any more, so what the host used to read back between
// kernels travels through two small blocks instead:
//  * tr_world::d_ctl (device memory): [CTL_ABORT] = 0, or 1 + the sequence number of the step whose
//    contact list overflowed its buffers.  EVERY kernel of every later step tests it first and
//    returns at once, so the state stays exactly as that step's integration left it (detection reads
//    only collider centres, which the response never writes) until the host, at the next
//    synchronisation, grows the buffers and re-runs the response of that step and the steps after it.
//    [CTL_SEQ] = number of collision steps completed on the device.
//  * tr_world::h_ring (mapped pinned host memory, written by the last kernel of a collision step):
//    one StepRecord per step - counts, the executor's statistics, globaltimer stamps of the phases - in a ring of
//    STEP_RING records, plus one slot (index STEP_RING) that holds the record of a step that overflowed until the
//    host has dealt with it.
enum { CTL_ABORT = 0, CTL_SEQ = 1, CTL_WORDS = 8 };
constexpr uint32_t TINY_FRAME_BODIES = 256;   // == TINY_MAX (tiny.cuh)
constexpr uint32_t STEP_RING = 256;
struct StepRecord {
    uint32_t seq;            // 1 + sequence number (0 = never written)
    uint32_t flags;          // 1: contact buffer overflow (response skipped), 2: dataflow watchdog fired
    uint32_t contacts_lo, contacts_hi;
    uint32_t side, far, bnd, grid;
    uint32_t rounds;         // iterations of the longest dataflow worker / wavefronts
    uint32_t long_segments;  // bodies whose contact list was sorted by a CTA (more than PREP_SHORT contacts)
    uint32_t pad[2];
    unsigned long long t[4]; // %globaltimer (ns) at the start of: broad phase, narrow phase, K6 prep; end of the step
};
static_assert(sizeof(StepRecord) == 80, "layout shared with the host");

struct StepConsts {
    float gx, gy, gz;
    float dt;
    float rho;
    float G;
};

// ---- per-kernel timing ----------------------------------------------------------------
enum TimerCat { T_GRAVITY = 0, T_BROAD, T_NARROW, T_RESPONSE, T_INTEGRATE, T_ALLGATHER, T_NCAT };

// a caller buffer page-locked in place for the per-frame host paths (api.cu note_host_buffer)
struct HostReg {
    void* ptr = nullptr;       // registered (page-locked) buffer
    size_t bytes = 0;
    void* last_ptr = nullptr;  // buffer seen by the previous call
    size_t last_bytes = 0;
};

struct PendingTimer {
    cudaEvent_t a, b;
    int cat;
};

}  // namespace tr

// The opaque world.  Device buffers are plain SoA fp32 arrays sized to `cap` bodies.
struct tr_world {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int num_sms = 0;
    tr_world_params params{};

    // host-side staging of created bodies (uploaded lazily before the next step)
    std::vector<tr_body_desc> host_bodies;
    uint64_t n = 0;          // bodies resident on the device
    uint64_t cap = 0;        // allocated capacity (bodies)
    bool dirty_upload = false;
    bool classified = false; // grid / side-list split is current
    bool ever_stepped = false;

    // sharding
    int rank = 0, nranks = 1;
    uint64_t own_first = 0, own_count = 0;
    void* nccl_comm = nullptr;

    // SoA state
    float4* pos_m[2] = {nullptr, nullptr};  // x,y,z,mass ; [cur] is live
    int cur = 0;
    float4* vel_rest = nullptr;             // vx,vy,vz,restitution
    float4* force = nullptr;                // external force xyz, pad
    float2* drag = nullptr;                 // Cd, area
    float4* center_r = nullptr;             // collider centre xyz, radius
    float4* half_ext = nullptr;             // collider size xyz (AABB half extents), pad
    uint8_t* flags = nullptr;
    tr_body_desc* aos = nullptr;            // device AoS staging for pack/unpack
    uint64_t aos_cap = 0;
    tr_body_desc* pinned = nullptr;         // pinned host staging
    uint64_t pinned_cap = 0;
    unsigned long long* d_sig = nullptr;    // classification hash of the last per-frame upload
    unsigned long long host_sig = 0;
    bool have_sig = false;
    tr::HostReg reg[3];                     // caller buffers page-locked in place: AoS descriptors, frame read-back, forces

    // optional trajectory ring: traj[(slot*3 + c) * traj_stride + body], traj_cnt[body] = points pushed
    float* traj = nullptr;
    uint32_t* traj_cnt = nullptr;
    uint32_t traj_max = 0;
    uint64_t traj_stride = 0;

    // gravity scratch
    float4* grav_partial = nullptr;         // [chunks][own_padded]
    uint64_t grav_partial_elems = 0;
    uint32_t* grav_ctrl = nullptr;          // [0]=work counter, [1..]=per-tile arrival counters
    uint64_t grav_ctrl_elems = 0;

    // collision scratch (collide.cu)
    struct Collide* col = nullptr;

    // device-side step bookkeeping (see StepRecord above)
    uint32_t* d_ctl = nullptr;              // device: CTL_WORDS words
    tr::StepRecord* h_ring = nullptr;       // mapped pinned host memory, STEP_RING + 1 records
    tr::StepRecord* d_ring = nullptr;       // device alias of h_ring
    uint32_t ring_consumed = 0;             // records folded into stats so far (== sequence number of the next one)
    uint32_t coll_seq = 0;                  // collision steps enqueued so far
    struct PendingStep {                    // one per step enqueued since the last synchronisation (overflow recovery)
        uint32_t coll_seq;                  // sequence number of its collision pass (or ~0u)
        int cur_after;                      // pos_m buffer index that is live after the step
    };
    std::vector<PendingStep> pending_steps;
    // small worlds: the single-CTA step writes the frame read-back (pos, vel, centre) straight into mapped pinned memory
    float* h_frame = nullptr;               // TINY_FRAME_BODIES x 9 floats, host view
    float* d_frame = nullptr;               // device alias
    bool frame_request = false;             // tr_step_frame asks the step it enqueues to leave the frame there
    bool frame_ready = false;               // ... and the single-CTA step did

    // stats / timing
    tr_stats stats{};
    bool timing = false;
    std::vector<tr::PendingTimer> pending;
    std::vector<cudaEvent_t> event_pool;
    int deferred_error = TR_OK;
    std::string deferred_msg;
};

namespace tr {

// timing helpers (api.cu)
void timer_begin(tr_world* w, int cat, PendingTimer* t);
void timer_end(tr_world* w, PendingTimer* t);
void timers_collect(tr_world* w);

// gravity.cu
int gravity_step(tr_world* w, const StepConsts& c, bool integrate_fused, float4* force_out);
// api.cu: wait for the stream, fold the step records in, recover from a contact-buffer overflow
int settle(tr_world* w);
int integrate_only(tr_world* w, const StepConsts& c);
int measure_fma_peak(int device, double* lane_ops_per_s, double* sm_clock_mhz);

// collide.cu
int collide_step(tr_world* w);
int tiny_steps(tr_world* w, const StepConsts& c, int nsteps, bool* taken);
int collide_consume_records(tr_world* w, bool* overflow, uint32_t* overflow_seq, uint64_t* found, bool* not_sparse);
int collide_grow_contacts(tr_world* w, uint64_t found);
int collide_headroom(tr_world* w);
int collide_free(tr_world* w);
int collide_read_contacts(tr_world* w, int32_t* pairs, uint64_t cap, uint64_t* n);
int collide_debug_grid(tr_world* w, uint32_t* keys, uint32_t* sorted_ids, uint64_t* m, float* cell, float* inv_cell,
                       int32_t* bits);
int collide_debug_candidates(tr_world* w, int32_t* pairs, uint64_t cap, uint64_t* n);
int collide_classify(tr_world* w);
void collide_poll_stats(tr_world* w);

// nccl_dyn.cu
int nccl_unique_id(void* out128);
int nccl_init(tr_world* w, const void* id128);
int nccl_allgather_pos(tr_world* w, float4* buf);
void nccl_destroy(tr_world* w);

}  // namespace tr

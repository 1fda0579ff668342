// K1 - all-pairs Newtonian gravity with the drag + integrator step fused into its epilogue.
//
// Replaces (a) the pair-force block of the reference, src/main.cpp:679-693 (dead code
// there; restated per SURVEY.md 8c) and (b) the per-body integration loop,
// src/main.cpp:763-791,811.
//
// Shape of the kernel (sm_100a):
//  * persistent CTAs (2 x 256 threads per SM: 123 registers, no spills) pull (target tile, source chunk) work items
//    from an atomic counter, so the 148 SMs stay evenly loaded for any N;
//  * every thread owns GRAV_PAIRS pairs of targets held as packed f32x2 registers and
//    runs the inner loop on FADD2/FMUL2/FFMA2 (two fp32 lanes per instruction: half the
//    issue slots of the scalar form) + MUFU.RSQ; the source is a broadcast scalar operand;
//  * sources stream through a 3-stage shared-memory ring filled by 1-D TMA bulk copies
//    (cp.async.bulk ... mbarrier::complete_tx), one LDS.128 broadcast per source;
//  * per-(chunk,target) partial sums go to HBM; the CTA that completes a tile's last chunk
//    reduces them in FIXED chunk order (deterministic, independent of scheduling and of
//    the number of GPUs) and runs drag + semi-implicit Euler for the tile's bodies;
//  * the self pair is handled only in the few subtiles that overlap the tile's own index
//    range, so the steady-state loop carries no guard.
//
// Algorithmic work: 12 FP32-pipe lane-ops + 1 MUFU.RSQ per ordered (target, source)
// interaction (3 FADD, 1 FMUL + 2 FFMA, 3 FMUL, 3 FFMA).
#include <cstdio>

#include "integrator.cuh"
#include "world.cuh"

namespace tr {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 r, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float rsq_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TR_DONE;\n"
        "bra TR_WAIT;\n"
        "TR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, u64* bar) {
    uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}

struct GravArgs {
    const float4* src;      // pos_m[cur]: x,y,z,mass of all n_src bodies
    uint32_t n_src;
    uint32_t tgt_first;     // owned target range
    uint32_t tgt_count;
    uint32_t chunk_len;
    uint32_t n_chunks;
    uint32_t n_tiles;
    uint32_t own_padded;    // n_tiles * GRAV_TILE: row stride of `partial`
    float4* partial;        // [n_chunks][own_padded]
    uint32_t* ctrl;         // [0] work counter, [1] exit counter, [2 + tile] arrivals
    float4* dst;            // pos_m[nxt]
    float4* vel_rest;
    float4* force;
    const float2* drag;
    float4* center_r;
    const uint8_t* flags;
    float4* force_out;      // optional: pair-gravity force per owned body (debug / parity)
    const uint32_t* ctl;    // [CTL_ABORT] != 0: an earlier collision step overflowed its buffers, do nothing (world.cuh)
    StepConsts c;
    int fused;
};

// One source against this thread's target pairs.  GUARD adds the self-pair test.
template <bool GUARD>
__device__ __forceinline__ void interact(const float4 q, uint32_t jg, uint32_t idx0, const u64 (&nx)[GRAV_PAIRS],
                                         const u64 (&ny)[GRAV_PAIRS], const u64 (&nz)[GRAV_PAIRS],
                                         u64 (&ax)[GRAV_PAIRS], u64 (&ay)[GRAV_PAIRS], u64 (&az)[GRAV_PAIRS]) {
    const u64 xj = pk(q.x, q.x), yj = pk(q.y, q.y), zj = pk(q.z, q.z), mj = pk(q.w, q.w);
#pragma unroll
    for (int p = 0; p < GRAV_PAIRS; ++p) {
        u64 dx = sub2(xj, nx[p]);              // xj - xi
        u64 dy = sub2(yj, ny[p]);
        u64 dz = sub2(zj, nz[p]);
        u64 d2 = mul2(dx, dx);
        d2 = fma2(dy, dy, d2);
        d2 = fma2(dz, dz, d2);
        float q0, q1;
        upk(d2, q0, q1);
        u64 inv = pk(rsq_approx(q0), rsq_approx(q1));
        u64 inv2 = mul2(inv, inv);
        u64 t = mul2(mj, inv);
        u64 s = mul2(t, inv2);                 // mj / d^3
        if (GUARD) {
            float s0, s1;
            upk(s, s0, s1);
            if (jg == idx0 + (uint32_t)(2 * p) * GRAV_THREADS) s0 = 0.0f;
            if (jg == idx0 + (uint32_t)(2 * p + 1) * GRAV_THREADS) s1 = 0.0f;
            s = pk(s0, s1);
        }
        ax[p] = fma2(dx, s, ax[p]);
        ay[p] = fma2(dy, s, ay[p]);
        az[p] = fma2(dz, s, az[p]);
    }
}

__global__ void __launch_bounds__(GRAV_THREADS, TR_GRAV_MINBLOCKS) gravity_kernel(const GravArgs a) {
    __shared__ __align__(128) float4 stage[GRAV_STAGES][GRAV_SUB];
    __shared__ __align__(8) u64 full[GRAV_STAGES];
    __shared__ uint32_t s_item, s_last;

    const uint32_t tid = threadIdx.x;
    const uint32_t total_items = a.n_tiles * a.n_chunks;
    if (a.ctl[CTL_ABORT] != 0) return;   // every CTA sees the same value: it was set by an earlier kernel

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GRAV_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_item = atomicAdd(&a.ctrl[0], 1u);
    }
    __syncthreads();

    uint32_t stage_w = 0;   // ring position of the next subtile to be consumed
    uint32_t phase = 0;
    uint32_t item = s_item;

    while (item < total_items) {
        const uint32_t tile = item / a.n_chunks;
        const uint32_t chunk = item - tile * a.n_chunks;
        uint32_t next_item = 0;
        if (tid == 0) next_item = atomicAdd(&a.ctrl[0], 1u);   // latency hidden behind the item

        const uint32_t src_lo = chunk * a.chunk_len;
        const uint32_t src_hi = min(src_lo + a.chunk_len, a.n_src);
        const uint32_t n_sub = (src_hi - src_lo + GRAV_SUB - 1) / GRAV_SUB;

        if (tid == 0) {
            // producer prologue: fill the ring
            uint32_t st = stage_w;
            for (uint32_t k = 0; k < n_sub && k < GRAV_STAGES; ++k) {
                uint32_t lo = src_lo + k * GRAV_SUB;
                uint32_t cnt = min((uint32_t)GRAV_SUB, src_hi - lo);
                tma_load_1d(&stage[st][0], a.src + lo, cnt * 16u, &full[st]);
                st = (st + 1 == GRAV_STAGES) ? 0 : st + 1;
            }
        }

        // this thread's targets: local slot (2p)*THREADS+tid and (2p+1)*THREADS+tid
        const uint32_t lt0 = tile * GRAV_TILE + tid;           // local index of pair 0, lane A
        const uint32_t idx0 = a.tgt_first + lt0;               // global body index of the same
        const uint32_t last_valid = a.tgt_first + a.tgt_count - 1;
        u64 nx[GRAV_PAIRS], ny[GRAV_PAIRS], nz[GRAV_PAIRS];
        u64 ax[GRAV_PAIRS], ay[GRAV_PAIRS], az[GRAV_PAIRS];
#pragma unroll
        for (int p = 0; p < GRAV_PAIRS; ++p) {
            uint32_t ia = min(idx0 + (uint32_t)(2 * p) * GRAV_THREADS, last_valid);
            uint32_t ib = min(idx0 + (uint32_t)(2 * p + 1) * GRAV_THREADS, last_valid);
            float4 A = __ldg(a.src + ia), B = __ldg(a.src + ib);
            // "+ 0" makes each pair the result of a packed arithmetic op: born as an aligned
            // register pair, so ptxas keeps it packed instead of re-assembling it from the
            // two float4 loads with MOVs inside the hot loop (it did: 25 MOVs per source)
            const u64 zero2 = pk(0.0f, 0.0f);
            nx[p] = add2(pk(A.x, B.x), zero2);
            ny[p] = add2(pk(A.y, B.y), zero2);
            nz[p] = add2(pk(A.z, B.z), zero2);
            ax[p] = ay[p] = az[p] = 0ull;
        }
        const uint32_t tile_lo = a.tgt_first + tile * GRAV_TILE;
        const uint32_t tile_hi = tile_lo + GRAV_TILE;

        for (uint32_t k = 0; k < n_sub; ++k) {
            const uint32_t lo = src_lo + k * GRAV_SUB;
            const uint32_t cnt = min((uint32_t)GRAV_SUB, src_hi - lo);
            mbar_wait(&full[stage_w], phase);
            const float4* sp = &stage[stage_w][0];
            if (lo < tile_hi && lo + cnt > tile_lo) {
                // subtile overlaps this tile's own bodies: self-pair guard (rare path)
                for (uint32_t j = 0; j < cnt; ++j) interact<true>(sp[j], lo + j, idx0, nx, ny, nz, ax, ay, az);
            } else if (cnt == GRAV_SUB) {
#pragma unroll 4
                for (uint32_t j = 0; j < GRAV_SUB; ++j) interact<false>(sp[j], 0, 0, nx, ny, nz, ax, ay, az);
            } else {
                for (uint32_t j = 0; j < cnt; ++j) interact<false>(sp[j], 0, 0, nx, ny, nz, ax, ay, az);
            }
            __syncthreads();   // everyone is done with this stage
            if (tid == 0 && k + GRAV_STAGES < n_sub) {
                uint32_t lo2 = src_lo + (k + GRAV_STAGES) * GRAV_SUB;
                uint32_t cnt2 = min((uint32_t)GRAV_SUB, src_hi - lo2);
                tma_load_1d(&stage[stage_w][0], a.src + lo2, cnt2 * 16u, &full[stage_w]);
            }
            if (++stage_w == GRAV_STAGES) { stage_w = 0; phase ^= 1u; }
        }

        // partial sums of this (tile, chunk)
        float4* prow = a.partial + (size_t)chunk * a.own_padded + lt0;
#pragma unroll
        for (int p = 0; p < GRAV_PAIRS; ++p) {
            float x0, x1, y0, y1, z0, z1;
            upk(ax[p], x0, x1);
            upk(ay[p], y0, y1);
            upk(az[p], z0, z1);
            __stcg(prow + (2 * p) * GRAV_THREADS, make_float4(x0, y0, z0, 0.0f));
            __stcg(prow + (2 * p + 1) * GRAV_THREADS, make_float4(x1, y1, z1, 0.0f));
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            uint32_t old = atomicAdd(&a.ctrl[2 + tile], 1u);
            s_last = (old == a.n_chunks - 1) ? 1u : 0u;
            s_item = next_item;
        }
        __syncthreads();

        if (s_last) {
            // last chunk of this tile has landed: ordered reduction + fused integrator
            __threadfence();
            float sx[GRAV_T], sy[GRAV_T], sz[GRAV_T];
#pragma unroll
            for (int t = 0; t < GRAV_T; ++t) sx[t] = sy[t] = sz[t] = 0.0f;
            const float4* pr = a.partial + lt0;
            for (uint32_t c = 0; c < a.n_chunks; ++c) {
#pragma unroll
                for (int t = 0; t < GRAV_T; ++t) {
                    float4 v = __ldcg(pr + (size_t)c * a.own_padded + t * GRAV_THREADS);
                    sx[t] = __fadd_rn(sx[t], v.x);
                    sy[t] = __fadd_rn(sy[t], v.y);
                    sz[t] = __fadd_rn(sz[t], v.z);
                }
            }
#pragma unroll
            for (int t = 0; t < GRAV_T; ++t) {
                const uint32_t lt = lt0 + t * GRAV_THREADS;
                if (lt >= a.tgt_count) continue;
                const uint32_t i = a.tgt_first + lt;
                float4 pm = a.src[i];
                const float gm = __fmul_rn(a.c.G, pm.w);
                const float fx = __fmul_rn(gm, sx[t]), fy = __fmul_rn(gm, sy[t]), fz = __fmul_rn(gm, sz[t]);
                if (a.force_out) a.force_out[lt] = make_float4(fx, fy, fz, 0.0f);
                if (a.fused) {
                    if (a.flags[i] & F_STATIONARY) {
                        a.dst[i] = pm;           // not integrated (main.cpp:764); carried to the new buffer
                    } else {
                        float4 vr = a.vel_rest[i];
                        float4 f = a.force[i];
                        integrate_body(pm, vr, __fadd_rn(f.x, fx), __fadd_rn(f.y, fy), __fadd_rn(f.z, fz), a.drag[i], a.c);
                        a.dst[i] = pm;
                        a.vel_rest[i] = vr;
                        float4 cr = a.center_r[i];
                        a.center_r[i] = make_float4(pm.x, pm.y, pm.z, cr.w);   // updateCollider, main.cpp:791
                        a.force[i] = make_float4(0.f, 0.f, 0.f, 0.f);          // main.cpp:811
                    }
                }
            }
            if (tid == 0) a.ctrl[2 + tile] = 0;   // re-arm for the next step
        }
        item = s_item;
    }

    // the last CTA to leave re-arms the work counter for the next launch
    if (tid == 0) {
        uint32_t old = atomicAdd(&a.ctrl[1], 1u);
        if (old == gridDim.x - 1) {
            a.ctrl[0] = 0;
            a.ctrl[1] = 0;
        }
    }
}

// Standalone O(N) integrator, used when pair gravity is off (the live reference step).
__global__ void __launch_bounds__(256) integrate_kernel(float4* pos_m, float4* vel_rest, float4* force,
                                                        const float2* drag, float4* center_r, const uint8_t* flags,
                                                        uint32_t first, uint32_t count, StepConsts c, const uint32_t* ctl) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count || ctl[CTL_ABORT] != 0) return;
    uint32_t i = first + t;
    if (flags[i] & F_STATIONARY) return;
    float4 pm = pos_m[i];
    float4 vr = vel_rest[i];
    float4 f = force[i];
    integrate_body(pm, vr, f.x, f.y, f.z, drag[i], c);
    pos_m[i] = pm;
    vel_rest[i] = vr;
    float4 cr = center_r[i];
    center_r[i] = make_float4(pm.x, pm.y, pm.z, cr.w);
    force[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

static int ensure_gravity_scratch(tr_world* w, uint32_t n_tiles, uint32_t n_chunks) {
    uint64_t need = (uint64_t)n_tiles * GRAV_TILE * n_chunks;
    if (need > w->grav_partial_elems) {
        if (w->grav_partial) cudaFree(w->grav_partial);
        w->grav_partial = nullptr;
        w->grav_partial_elems = 0;
        TR_CUDA(cudaMalloc(&w->grav_partial, need * sizeof(float4)));
        w->grav_partial_elems = need;
    }
    uint64_t need_ctrl = 2 + (uint64_t)n_tiles;
    if (need_ctrl > w->grav_ctrl_elems) {
        if (w->grav_ctrl) cudaFree(w->grav_ctrl);
        w->grav_ctrl = nullptr;
        w->grav_ctrl_elems = 0;
        TR_CUDA(cudaMalloc(&w->grav_ctrl, need_ctrl * sizeof(uint32_t)));
        TR_CUDA(cudaMemsetAsync(w->grav_ctrl, 0, need_ctrl * sizeof(uint32_t), w->stream));
        w->grav_ctrl_elems = need_ctrl;
    }
    return TR_OK;
}

// One gravity pass over the owned targets.  integrate_fused: run the epilogue integrator
// and write pos_m[cur^1] (the caller flips `cur`); otherwise only force_out is produced.
int gravity_step(tr_world* w, const StepConsts& c, bool integrate_fused, float4* force_out) {
    if (w->own_count == 0) return TR_OK;
    const uint32_t n = (uint32_t)w->n;
    GravArgs a{};
    a.src = w->pos_m[w->cur];
    a.n_src = n;
    a.tgt_first = (uint32_t)w->own_first;
    a.tgt_count = (uint32_t)w->own_count;
    a.chunk_len = grav_chunk_len(w->n);
    a.n_chunks = (n + a.chunk_len - 1) / a.chunk_len;
    a.n_tiles = (a.tgt_count + GRAV_TILE - 1) / GRAV_TILE;
    a.own_padded = a.n_tiles * GRAV_TILE;
    TR_TRY(ensure_gravity_scratch(w, a.n_tiles, a.n_chunks));
    a.partial = w->grav_partial;
    a.ctrl = w->grav_ctrl;
    a.dst = w->pos_m[w->cur ^ 1];
    a.vel_rest = w->vel_rest;
    a.force = w->force;
    a.drag = w->drag;
    a.center_r = w->center_r;
    a.flags = w->flags;
    a.force_out = force_out;
    a.ctl = w->d_ctl;
    a.c = c;
    a.fused = integrate_fused ? 1 : 0;

    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        int v = 0;
        TR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, gravity_kernel, GRAV_THREADS, 0));
        ctas_per_sm = v > 0 ? v : 1;
    }
    uint64_t items = (uint64_t)a.n_tiles * a.n_chunks;
    uint64_t grid = (uint64_t)w->num_sms * ctas_per_sm;
    if (grid > items) grid = items;

    PendingTimer t;
    timer_begin(w, T_GRAVITY, &t);
    gravity_kernel<<<(unsigned)grid, GRAV_THREADS, 0, w->stream>>>(a);
    timer_end(w, &t);
    TR_CUDA(cudaGetLastError());
    w->stats.kernel_launches += 1;
    w->stats.gravity_launches += 1;
    w->stats.interactions += (uint64_t)a.tgt_count * n;
    return TR_OK;
}

int integrate_only(tr_world* w, const StepConsts& c) {
    if (w->own_count == 0) return TR_OK;
    PendingTimer t;
    timer_begin(w, T_INTEGRATE, &t);
    unsigned blocks = (unsigned)((w->own_count + 255) / 256);
    integrate_kernel<<<blocks, 256, 0, w->stream>>>(w->pos_m[w->cur], w->vel_rest, w->force, w->drag, w->center_r,
                                                    w->flags, (uint32_t)w->own_first, (uint32_t)w->own_count, c, w->d_ctl);
    timer_end(w, &t);
    TR_CUDA(cudaGetLastError());
    w->stats.kernel_launches += 1;
    return TR_OK;
}

// ---- FFMA-only peak (roofline denominator for K1) ----------------------------------------
__global__ void __launch_bounds__(256) fma_peak_kernel(const float* __restrict__ in, float* out, int iters,
                                                       long long* cycles) {
    float b = in[threadIdx.x & 31], c = in[32 + (threadIdx.x & 31)];
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = in[64 + k] + (float)threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], b, c);
    }
    long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int measure_fma_peak(int device, double* lane_ops_per_s, double* sm_clock_mhz) {
    TR_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    TR_CUDA(cudaGetDeviceProperties(&p, device));
    const int blocks = p.multiProcessorCount * 2;
    const int iters = 100000;
    float h[128];
    for (int i = 0; i < 128; ++i) h[i] = 0.5f + 0.001f * i;
    float *din = nullptr, *dout = nullptr;
    long long* dcyc = nullptr;
    TR_CUDA(cudaMalloc(&din, sizeof(h)));
    TR_CUDA(cudaMalloc(&dout, sizeof(float) * 256 * blocks));
    TR_CUDA(cudaMalloc(&dcyc, sizeof(long long) * blocks));
    TR_CUDA(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    TR_CUDA(cudaEventCreate(&e0));
    TR_CUDA(cudaEventCreate(&e1));
    fma_peak_kernel<<<blocks, 256>>>(din, dout, 1000, dcyc);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        TR_CUDA(cudaEventRecord(e0));
        fma_peak_kernel<<<blocks, 256>>>(din, dout, iters, dcyc);
        TR_CUDA(cudaEventRecord(e1));
        TR_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        TR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    std::vector<long long> hc(blocks);
    TR_CUDA(cudaMemcpy(hc.data(), dcyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < blocks; ++i) mx = hc[i] > mx ? hc[i] : mx;
    double total = 16.0 * iters * 256.0 * blocks;
    if (lane_ops_per_s) *lane_ops_per_s = total / (best * 1e-3);
    // two CTAs per SM run concurrently, so the longest CTA's cycle count spans the launch
    if (sm_clock_mhz) *sm_clock_mhz = (double)mx / (best * 1e-3) / 1e6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(din);
    cudaFree(dout);
    cudaFree(dcyc);
    return TR_OK;
}

}  // namespace tr

// FP32-pipe microbenchmark for B200 (sm_100a): what does one SM sub-partition sustain for
//   mode 0  scalar FFMA (3 register operands)
//   mode 1  packed FFMA2 (fma.rn.f32x2, two fp32 lanes per register pair)
//   mode 2  MUFU.RSQ
//   mode 3  the gravity inner-loop mix: 12 packed FMA-pipe ops + 2 MUFU.RSQ per pair
//   mode 4  the same mix with scalar ops: 12 FMA-pipe ops + 1 MUFU.RSQ
// Prints lane-ops per clock per SM (from clock64) and per second (from CUDA events), so
// the roofline denominator of the gravity kernel is a measured number, not a datasheet
// one (MEASURED_PEAKS.json has no FP32 row).  Build: see tools/Makefile.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

constexpr int ACC = 16;

template <int MODE>
__global__ void __launch_bounds__(256) bench(const float* __restrict__ in, float* out, int iters, long long* cycles) {
    float b = in[threadIdx.x & 31], c = in[32 + (threadIdx.x & 31)];
    long long t0 = clock64();
    float res = 0.f;
    if (MODE == 0) {
        float a[ACC];
#pragma unroll
        for (int k = 0; k < ACC; ++k) a[k] = in[64 + k] + threadIdx.x;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < ACC; ++k) a[k] = fmaf(a[k], b, c);
        }
#pragma unroll
        for (int k = 0; k < ACC; ++k) res += a[k];
    } else if (MODE == 1) {
        u64 a[ACC];
        u64 b2 = pk(b, b + 1.f), c2 = pk(c, c + 1.f);
#pragma unroll
        for (int k = 0; k < ACC; ++k) a[k] = pk(in[64 + k] + threadIdx.x, in[64 + k]);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < ACC; ++k) a[k] = fma2(a[k], b2, c2);
        }
#pragma unroll
        for (int k = 0; k < ACC; ++k) { float x, y; upk(a[k], x, y); res += x + y; }
    } else if (MODE == 2) {
        float a[ACC];
#pragma unroll
        for (int k = 0; k < ACC; ++k) a[k] = in[64 + k] + threadIdx.x + 1.0f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < ACC; ++k) a[k] = rsq(a[k]);
        }
#pragma unroll
        for (int k = 0; k < ACC; ++k) res += a[k];
    } else if (MODE == 3) {
        // 4 target pairs per thread, one broadcast source per iteration
        constexpr int P = 4;
        u64 nx[P], ny[P], nz[P], ax[P], ay[P], az[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            nx[p] = pk(in[64 + p] + threadIdx.x, in[70 + p]); ny[p] = pk(in[65 + p], in[71 + p] + threadIdx.x); nz[p] = pk(in[66 + p], in[72 + p]);
            ax[p] = ay[p] = az[p] = 0;
        }
        float sx = b, sy = c, sz = b + c, sm = 1.0f;
        for (int i = 0; i < iters; ++i) {
            u64 xj = pk(sx, sx), yj = pk(sy, sy), zj = pk(sz, sz), mj = pk(sm, sm);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                u64 dx = add2(xj, nx[p]), dy = add2(yj, ny[p]), dz = add2(zj, nz[p]);
                u64 d2 = mul2(dx, dx); d2 = fma2(dy, dy, d2); d2 = fma2(dz, dz, d2);
                float q0, q1; upk(d2, q0, q1);
                u64 inv = pk(rsq(q0), rsq(q1));
                u64 inv2 = mul2(inv, inv), t = mul2(mj, inv), s = mul2(t, inv2);
                ax[p] = fma2(dx, s, ax[p]); ay[p] = fma2(dy, s, ay[p]); az[p] = fma2(dz, s, az[p]);
            }
            sx += 1.0f; sy += 0.5f;   // keep the source changing (2 extra scalar FADD per iteration)
        }
#pragma unroll
        for (int p = 0; p < P; ++p) { float x, y; upk(ax[p], x, y); res += x + y; upk(ay[p], x, y); res += x + y; upk(az[p], x, y); res += x + y; }
    } else if (MODE == 4) {
        constexpr int P = 8;
        float nx[P], ny[P], nz[P], ax[P], ay[P], az[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { nx[p] = in[64 + p] + threadIdx.x; ny[p] = in[65 + p]; nz[p] = in[66 + p]; ax[p] = ay[p] = az[p] = 0.f; }
        float sx = b, sy = c, sz = b + c, sm = 1.0f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float dx = sx + nx[p], dy = sy + ny[p], dz = sz + nz[p];
                float d2 = dx * dx; d2 = fmaf(dy, dy, d2); d2 = fmaf(dz, dz, d2);
                float inv = rsq(d2);
                float inv2 = inv * inv, t = sm * inv, s = t * inv2;
                ax[p] = fmaf(dx, s, ax[p]); ay[p] = fmaf(dy, s, ay[p]); az[p] = fmaf(dz, s, az[p]);
            }
            sx += 1.0f; sy += 0.5f;
        }
#pragma unroll
        for (int p = 0; p < P; ++p) res += ax[p] + ay[p] + az[p];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, double laneops_per_thread_iter, int blocks_per_sm, int iters, int nsm, const float* din, float* dout, long long* dcyc) {
    int blocks = nsm * blocks_per_sm;
    bench<MODE><<<blocks, 256>>>(din, dout, 1000, dcyc);   // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        bench<MODE><<<blocks, 256>>>(din, dout, iters, dcyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long* hc = (long long*)malloc(sizeof(long long) * blocks);
    CK(cudaMemcpy(hc, dcyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < blocks; ++i) cyc += (double)hc[i]; cyc /= blocks;
    free(hc);
    double total = laneops_per_thread_iter * (double)iters * 256.0 * blocks;
    double per_clk_sm = laneops_per_thread_iter * (double)iters * 256.0 * blocks_per_sm / cyc;
    printf("%-28s blocks/SM=%d  %8.3f ms  %8.2f Tlaneop/s  %7.2f laneop/clk/SM  (eff clk %.0f MHz)\n", name, blocks_per_sm, best,
           total / (best * 1e-3) / 1e12, per_clk_sm, cyc / (best * 1e-3) / 1e6);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clockRate %d kHz\n", p.name, nsm, p.clockRate);
    float h[128]; for (int i = 0; i < 128; ++i) h[i] = 0.5f + 0.001f * i;
    float* din; float* dout; long long* dcyc;
    CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dout, sizeof(float) * 256 * nsm * 8)); CK(cudaMalloc(&dcyc, sizeof(long long) * nsm * 8));
    const int iters = 200000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        run<0>("scalar FFMA (1 laneop)", ACC * 1.0, bps, iters, nsm, din, dout, dcyc);
        run<1>("packed FFMA2 (2 laneop)", ACC * 2.0, bps, iters, nsm, din, dout, dcyc);
        run<2>("MUFU.RSQ", ACC * 1.0, bps, iters / 4, nsm, din, dout, dcyc);
        run<3>("gravity mix packed (12/int)", 4 * 2 * 12.0, bps, iters / 4, nsm, din, dout, dcyc);
        run<4>("gravity mix scalar (12/int)", 8 * 12.0, bps, iters / 4, nsm, din, dout, dcyc);
    }
    // sustained: ~3 s of packed-mix back to back to see the clock under load
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int reps = 60;
        for (int r = 0; r < reps; ++r) bench<3><<<nsm * 2, 256>>>(din, dout, iters / 4, dcyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double total = 4 * 2 * 12.0 * (iters / 4) * 256.0 * nsm * 2 * reps;
        printf("sustained packed mix: %.1f ms total, %.2f Tlaneop/s (= %.3f T interactions/s)\n", ms, total / (ms * 1e-3) / 1e12, total / 12.0 / (ms * 1e-3) / 1e12);
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) bench<1><<<nsm * 2, 256>>>(din, dout, iters, dcyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total = ACC * 2.0 * iters * 256.0 * nsm * 2 * reps;
        printf("sustained FFMA2 only: %.1f ms total, %.2f Tlaneop/s\n", ms, total / (ms * 1e-3) / 1e12);
    }
    return 0;
}

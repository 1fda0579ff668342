// Drag + uniform gravity + force sum + semi-implicit Euler for one body: the O(N) part of
// the reference step (reference src/main.cpp:763-791, 811).  Shared by the gravity
// kernel's fused epilogue and by the standalone integrator used when pair gravity is off.
//
// Every operation is an explicitly rounded intrinsic (__fmul_rn / __fadd_rn / __fdiv_rn /
// __fsqrt_rn): nvcc never contracts those into FMAs, so given the same inputs the result
// is BIT-IDENTICAL to the reference compiled without FMA contraction (its CMakeLists.txt
// sets no -march) - the property tests/test_gpu_parity.py::test_integrator_bit_exact checks.
#pragma once

#include "world.cuh"

namespace tr {

// pm: x,y,z,mass  vr: vx,vy,vz,rest  ext: external force (already includes pair gravity
// when enabled)  dr: Cd, area.  Updates pm.xyz and vr.xyz in place.
__device__ __forceinline__ void integrate_body(float4& pm, float4& vr, float ex, float ey, float ez, float2 dr,
                                               const StepConsts& c) {
    // speedSq = dot(vel, vel)                                      main.cpp:767, 195-198
    float s2 = __fmul_rn(vr.x, vr.x);
    s2 = __fadd_rn(s2, __fmul_rn(vr.y, vr.y));
    s2 = __fadd_rn(s2, __fmul_rn(vr.z, vr.z));
    float dx = 0.0f, dy = 0.0f, dz = 0.0f;
    if (s2 > 0.0001f) {                                          // main.cpp:771-774
        float inv = __fdiv_rn(1.0f, __fsqrt_rn(s2));
        dx = __fmul_rn(vr.x, inv);
        dy = __fmul_rn(vr.y, inv);
        dz = __fmul_rn(vr.z, inv);
    }
    // -0.5f * atmDensity * Cd * area * speedSq, left to right       main.cpp:777
    float k = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(-0.5f, c.rho), dr.x), dr.y), s2);
    float m = pm.w;
    // totalForce = (Fd + gravity*mass) + force                     main.cpp:780-782
    float tx = __fadd_rn(__fadd_rn(__fmul_rn(dx, k), __fmul_rn(c.gx, m)), ex);
    float ty = __fadd_rn(__fadd_rn(__fmul_rn(dy, k), __fmul_rn(c.gy, m)), ey);
    float tz = __fadd_rn(__fadd_rn(__fmul_rn(dz, k), __fmul_rn(c.gz, m)), ez);
    // a = F/m ; v += a*dt ; x += v*dt                               main.cpp:786-788
    vr.x = __fadd_rn(vr.x, __fmul_rn(__fdiv_rn(tx, m), c.dt));
    vr.y = __fadd_rn(vr.y, __fmul_rn(__fdiv_rn(ty, m), c.dt));
    vr.z = __fadd_rn(vr.z, __fmul_rn(__fdiv_rn(tz, m), c.dt));
    pm.x = __fadd_rn(pm.x, __fmul_rn(vr.x, c.dt));
    pm.y = __fadd_rn(pm.y, __fmul_rn(vr.y, c.dt));
    pm.z = __fadd_rn(pm.z, __fmul_rn(vr.z, c.dt));
}

}  // namespace tr

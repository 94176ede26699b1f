// microclimf_b200 — single-precision elementary functions of the FP32 build: one SFU operation (MUFU.RCP / RSQ / SQRT /
// EX2 / LG2) plus at most one multiply each, no special-case branches.  Accuracy ~1e-6 relative (1-2 ulp for rcp / sqrt,
// 2^-22 for ex2 / lg2): the FP32 build's budget is 0.05 degC / 0.5 % radiation.
#pragma once
#include <cuda_runtime.h>

namespace mcf {
namespace f32 {

__device__ __forceinline__ float frcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fdiv(float a, float b) { return a * frcp(b); }
__device__ __forceinline__ float fsqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float flg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fexp(float x) { return fex2(x * 1.4426950408889634f); }
__device__ __forceinline__ float fexp2(float x) { return fex2(x); }
__device__ __forceinline__ float flog(float x) { return flg2(x) * 0.6931471805599453f; }
__device__ __forceinline__ float fpow(float x, float y) { return fex2(y * flg2(x)); }

} // namespace f32
} // namespace mcf

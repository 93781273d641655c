// common.cuh -- shared helpers for libnunerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/nunerf.h"

namespace nunerf {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(const char* fmt, const char* a = "", int code = -1) {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}

#define NUNERF_CHECK_LAUNCH(name)                                                   \
  do {                                                                              \
    ::nunerf::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      snprintf(::nunerf::g_err, sizeof(::nunerf::g_err), "%s: %s", name, cudaGetErrorString(e__)); \
      return -2;                                                                    \
    }                                                                               \
  } while (0)

#define NUNERF_REQUIRE(cond, msg)                                                   \
  do {                                                                              \
    if (!(cond)) return ::nunerf::fail("%s", msg, -1);                              \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- bf16 plane helpers
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16lo_to_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi_to_f(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// store one value into bf16 planes
__device__ __forceinline__ void store_planes(__nv_bfloat16* base, long long idx, int lo_off, float v) {
  __nv_bfloat16 hi = __float2bfloat16_rn(v);
  base[idx] = hi;
  if (lo_off) base[idx + lo_off] = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ float load_planes(const __nv_bfloat16* base, long long idx, int lo_off) {
  float v = __bfloat162float(base[idx]);
  if (lo_off) v += __bfloat162float(base[idx + lo_off]);
  return v;
}

// Softplus(beta = 100): max(x, 0) + 0.01 log1p(e), e = exp(-|100 x|).  One MUFU (ex2) per element; log1p(e) on (0, 1]
// is e Q(e) with a degree-4 near-minimax Q (|error| < 1e-5, i.e. < 1e-7 after the 0.01 factor, folded into Q).
__device__ __forceinline__ float softplus100(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(x) * -144.26950408889634f));
  float q = 3.215121477842331e-4f;
  q = fmaf(q, e, -1.3604211807250977e-3f);
  q = fmaf(q, e, 2.8945398330688477e-3f);
  q = fmaf(q, e, -4.9190032482147217e-3f);
  q = fmaf(q, e, 9.994943737983704e-3f);
  return fmaf(e, q, fmaxf(x, 0.0f));
}

// ---------------------------------------------------------------- packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2)
// The fma pipe issues one three-register FFMA per two cycles and scheduler; the packed forms do two lanes of work per
// issue slot, which is what bounds the activation epilogues of the fused chains.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// softplus100 of the pair (a + ba, b + bb), same arithmetic as softplus100 (bit-identical results: every step is the same
// correctly rounded fp32 operation, only issued two lanes at a time)
__device__ __forceinline__ void softplus100_x2(float& a, float& b, float ba, float bb) {
  const f32x2_t x2 = f2_add(f2_pack(a, b), f2_pack(ba, bb));
  float x0, x1, t0, t1;
  f2_unpack(x2, x0, x1);
  f2_unpack(f2_mul(x2, f2_pack(144.26950408889634f, 144.26950408889634f)), t0, t1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(t0, -t0)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(t1, -t1)));
  const f32x2_t e2 = f2_pack(e0, e1);
  f32x2_t q = f2_fma(f2_pack(3.215121477842331e-4f, 3.215121477842331e-4f), e2, f2_pack(-1.3604211807250977e-3f, -1.3604211807250977e-3f));
  q = f2_fma(q, e2, f2_pack(2.8945398330688477e-3f, 2.8945398330688477e-3f));
  q = f2_fma(q, e2, f2_pack(-4.9190032482147217e-3f, -4.9190032482147217e-3f));
  q = f2_fma(q, e2, f2_pack(9.994943737983704e-3f, 9.994943737983704e-3f));
  f2_unpack(f2_fma(e2, q, f2_pack(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f))), a, b);
}

// ---------------------------------------------------------------- deterministic fp32 math (sampling path)
// The sampling kernels and the C oracle (oracle/sampling_oracle.c) must agree bit for bit, so every
// operation is an explicitly rounded fp32 op (no FMA contraction) and exp() is our own polynomial.
__host__ __device__ __forceinline__ float det_mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
__host__ __device__ __forceinline__ float det_add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
__host__ __device__ __forceinline__ float det_sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
__host__ __device__ __forceinline__ float det_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
__host__ __device__ __forceinline__ float det_sqrt(float a) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(a);
#else
  volatile float r = sqrtf(a); return r;
#endif
}
// exp(x) for x <= 0 (and moderately positive): Cody-Waite reduction + degree-7 Taylor/Horner, all rounded fp32 ops.
__host__ __device__ __forceinline__ float det_exp(float x) {
  // branch-free (selects instead of early returns: a warp never runs the polynomial twice); same bits as the early-return
  // form for every input
  const bool under = x < -87.0f;
  x = fminf(fmaxf(x, -87.0f), 88.0f);
  float t = det_mul(x, 1.44269504088896341f);
  float n = (t >= 0.0f) ? (float)(int)(det_add(t, 0.5f)) : (float)(int)(det_sub(t, 0.5f));
  float r = det_sub(x, det_mul(n, 0.693359375f));
  r = det_sub(r, det_mul(n, -2.12194440e-4f));
  float p = 1.0f / 5040.0f;
  p = det_add(det_mul(p, r), 1.0f / 720.0f);
  p = det_add(det_mul(p, r), 1.0f / 120.0f);
  p = det_add(det_mul(p, r), 1.0f / 24.0f);
  p = det_add(det_mul(p, r), 1.0f / 6.0f);
  p = det_add(det_mul(p, r), 0.5f);
  p = det_add(det_mul(p, r), 1.0f);
  p = det_add(det_mul(p, r), 1.0f);
  int e = (int)n + 127;
  union { uint32_t u; float f; } s;
  s.u = (uint32_t)(e > 0 ? e : 0) << 23;
  const float v = det_mul(p, s.f);
  return (under || e <= 0) ? 0.0f : v;
}
__host__ __device__ __forceinline__ float det_sigmoid(float x) {
  // 1/(1+exp(-x)) for x >= 0, exp(x)/(1+exp(x)) for x < 0: both need e = exp(-|x|) only, so one evaluation serves a warp
  // with mixed signs
  const float e = det_exp(-fabsf(x));
  return det_div(x >= 0.0f ? 1.0f : e, det_add(1.0f, e));
}

}  // namespace nunerf

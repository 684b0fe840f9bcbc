// Shared helpers for the caduceus_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/caduceus_b200.h"

namespace cad {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);   // defined in api.cu (thread-local buffer)

#define CAD_REQUIRE(cond, ...)                       \
  do {                                               \
    if (!(cond)) {                                   \
      ::cad::set_error(__VA_ARGS__);                 \
      return -1;                                     \
    }                                                \
  } while (0)

#define CAD_LAUNCH_CHECK()                           \
  do {                                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) {                        \
      ::cad::set_error("CUDA launch failed: %s", cudaGetErrorString(e__)); \
      return (int)e__;                               \
    }                                                \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- dtype traits ---------------------------------------------------------------------------------
template <typename T> struct io;
template <> struct io<float> {
  static constexpr int kDtype = CAD_F32;
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
};
template <> struct io<__half> {
  static constexpr int kDtype = CAD_F16;
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct io<__nv_bfloat16> {
  static constexpr int kDtype = CAD_BF16;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

inline size_t dtype_size(int dt) { return dt == CAD_F32 ? 4 : 2; }

// Dispatch a functor templated on the io element type.
#define CAD_DISPATCH_DTYPE(dt, T, ...)                                   \
  switch (dt) {                                                          \
    case CAD_F32:  { using T = float;          __VA_ARGS__; break; }     \
    case CAD_F16:  { using T = __half;         __VA_ARGS__; break; }     \
    case CAD_BF16: { using T = __nv_bfloat16;  __VA_ARGS__; break; }     \
    default: ::cad::set_error("unknown dtype %d", (int)(dt)); return -1; \
  }

// ---- vector load/store of V consecutive elements into/out of fp32 registers -------------------------
// 16-bit types: 8 elements per 128-bit access; fp32: 4 per access.
template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&v)[V]) {
  static_assert(V % 8 == 0, "V must be a multiple of 8");
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < V / 8; ++i) {
      uint4 raw = __ldg(reinterpret_cast<const uint4*>(p) + i);
      const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i * 8 + k] = io<T>::to_f(e[k]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < V / 4; ++i) {
      float4 raw = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[i * 4 + 0] = raw.x; v[i * 4 + 1] = raw.y; v[i * 4 + 2] = raw.z; v[i * 4 + 3] = raw.w;
    }
  }
}

// same, from shared memory (plain loads; __ldg is for global memory only)
template <typename T, int V>
__device__ __forceinline__ void load_vec_smem(const T* p, float (&v)[V]) {
  static_assert(V % 8 == 0, "V must be a multiple of 8");
  constexpr int EPV = 16 / sizeof(T);
#pragma unroll
  for (int i = 0; i < V / EPV; ++i) {
    const uint4 raw = reinterpret_cast<const uint4*>(p)[i];
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int k = 0; k < EPV; ++k) v[i * EPV + k] = io<T>::to_f(e[k]);
  }
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&v)[V]) {
  static_assert(V % 8 == 0, "V must be a multiple of 8");
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < V / 8; ++i) {
      uint4 raw;
      T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
      for (int k = 0; k < 8; ++k) e[k] = io<T>::from_f(v[i * 8 + k]);
      reinterpret_cast<uint4*>(p)[i] = raw;
    }
  } else {
#pragma unroll
    for (int i = 0; i < V / 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[i * 4 + 0], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
  }
}

// ---- math -------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2(float x) {      // MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {      // MUFU.LG2
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp(float x) {      // MUFU.RCP
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// silu(v) = v / (1 + exp(-v))
__device__ __forceinline__ float silu(float v) { return v * rcp(1.0f + ex2(-kLog2e * v)); }
// silu for 16-bit I/O: x*sigmoid(x) = h + h*tanh(h), h = x/2 — ONE MUFU (tanh.approx, rel. error 2^-11, below the
// bf16/fp16 rounding of the result) instead of two (ex2 + rcp).  The scan is MUFU-bound, so the activations of its
// prologue/epilogue are worth 2 of the ~22 MUFU ops per (token, channel).  fp32 I/O keeps the exact form.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T>
__device__ __forceinline__ float silu_io(float v) {
  if constexpr (sizeof(T) == 2) {
    const float h = 0.5f * v;
    return fmaf(h, tanh_approx(h), h);
  } else {
    return silu(v);
  }
}
// softplus with torch's threshold (20): log1p(exp(v)) = ln2 * log2(1 + 2^(v*log2e))
__device__ __forceinline__ float softplus(float v) {
  const float w = ex2(kLog2e * v);                       // e^v
  float sp = kLn2 * lg2(1.0f + w);
  // small w: 1 + w loses the low bits of w and lg2.approx has 2^-22 ABSOLUTE error near 1, which is a
  // visible RELATIVE error on dt ~ 1e-3.  Use the alternating series of log1p there (error < w^5/5).
  const float series = w * (1.0f - w * (0.5f - w * (0.33333334f - 0.25f * w)));
  sp = (w < 0.015625f) ? series : sp;
  return v > 20.0f ? v : sp;
}

}  // namespace cad

// The small set of SIMT primitives the lane = channel scan (scan_fwd_v20.cuh) is written against, so that tests/emu/ can compile
// THAT file for the host (-DCAD_EMULATE: lanes are threads, shared memory is a checked byte array, cp.async / bulk copies /
// mbarriers are modelled; tests/emu/simt_emu.h provides the same names) and check its index logic and hand-over protocol without
// a GPU.  On the device they are inline PTX.
#pragma once

#ifdef CAD_EMULATE
#include "simt_emu.h"
#else
#include "scan_common.cuh"

namespace cad {
namespace simt {
#define CAD_DEV __device__ __forceinline__
#define CAD_TID ((int)threadIdx.x)
#define CAD_NTHREADS ((int)blockDim.x)
#define CAD_BIDX ((int)blockIdx.x)
#define CAD_BIDY ((int)blockIdx.y)
#define CAD_BIDZ ((int)blockIdx.z)

// packed fp32 pairs (FMUL2 / FFMA2 / FADD2)
CAD_DEV unsigned long long& as_u64(float2& v) { return *reinterpret_cast<unsigned long long*>(&v); }
CAD_DEV const unsigned long long& as_u64(const float2& v) { return *reinterpret_cast<const unsigned long long*>(&v); }
CAD_DEV float2 fma2(const float2& a, const float2& b, const float2& c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
  return d;
}
CAD_DEV float2 mul2(const float2& a, const float2& b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
CAD_DEV float2 add2(const float2& a, const float2& b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
// shared memory by 32-bit address
CAD_DEV uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
CAD_DEV uint32_t lds16u(uint32_t addr) {               // one 16-bit element (zero-extended)
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
CAD_DEV void sts16u(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory"); }
CAD_DEV void sts32u(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
CAD_DEV uint32_t atomic_inc_shared(uint32_t addr) {          // returns the value before the increment
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
  return old;
}
// per-lane staging and the 1-D bulk copy (TMA engine, no tensor map; completion counted in bytes on an mbarrier)
CAD_DEV void cp_async16s(uint32_t smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc));
}
CAD_DEV void cp_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> CAD_DEV void cp_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
CAD_DEV void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
CAD_DEV void cta_sync() { __syncthreads(); }
CAD_DEV void warp_sync() { __syncwarp(); }
CAD_DEV bool warp_all(bool p) { return __all_sync(0xffffffffu, p) != 0; }
CAD_DEV void stg128(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
CAD_DEV void stg128f(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
CAD_DEV uint32_t f2u(float f) { return __float_as_uint(f); }
CAD_DEV float u2f(uint32_t u) { return __uint_as_float(u); }
}  // namespace simt
}  // namespace cad
#endif  // !CAD_EMULATE

namespace cad {
namespace simt {
CAD_DEV float2 splat(float v) { return make_float2(v, v); }
CAD_DEV float2 ex2_2(const float2& x) { return make_float2(ex2(x.x), ex2(x.y)); }
}  // namespace simt
}  // namespace cad

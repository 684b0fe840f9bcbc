// Host-side SIMT emulation of the primitives the lane = channel scan kernel is written against (TEST INFRASTRUCTURE ONLY).
//
// caduceus_b200/csrc/scan_fwd_v20.cuh (through csrc/simt.cuh) is compiled for the CPU with -DCAD_EMULATE: every CUDA thread
// becomes an OS thread, __syncthreads / __syncwarp are barriers, shared memory is a checked byte array addressed by 32-bit
// offsets, mbarriers keep (pending arrivals, pending bytes, phase), cp.async and the 1-D bulk copy are synchronous copies.
// The point is to check the kernel's INDEX LOGIC (reversed jobs, tails, conv neighbours, segments, buffer hand-over)
// against a float64 restatement where no GPU is available; arithmetic uses exact exp2f/tanhf/log2f instead of MUFU.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include <vector_types.h>
#include <vector_functions.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/caduceus_b200.h"

namespace cad {

// ---- what common.cuh / scan_common.cuh provide on the device ------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kBlkTok = 32;

template <typename T> struct io;
template <> struct io<float> {
  static float to_f(float v) { return v; }
  static float from_f(float v) { return v; }
};
template <> struct io<__half> {
  static float to_f(__half v) { return __half2float(v); }
  static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct io<__nv_bfloat16> {
  static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

struct Mbar { int init = 0, pending = 0; long tx = 0; uint32_t phase = 0; };

struct EmuCta {
  unsigned char* smem = nullptr;
  size_t smem_bytes = 0;
  int nthreads = 0;
  pthread_barrier_t cta_bar;
  std::vector<pthread_barrier_t> warp_bar;
  std::vector<std::atomic<int>> vote;      // [warps][2]: warp_all flags, used alternately
  std::mutex m;
  std::condition_variable cv;
  std::map<const void*, Mbar> mbars;
  bool deadlock = false;
};

struct EmuThread { EmuCta* cta; int tid, bx, by, bz = 0; int vote_epoch = 0; };
extern thread_local EmuThread g_t;

namespace simt {
#define CAD_DEV inline
#define CAD_TID (::cad::g_t.tid)
#define CAD_NTHREADS (::cad::g_t.cta->nthreads)
#define CAD_BIDX (::cad::g_t.bx)
#define CAD_BIDY (::cad::g_t.by)
#define CAD_BIDZ (::cad::g_t.bz)
#define __restrict__
}  // namespace simt

inline uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - g_t.cta->smem); }
inline unsigned char* smem_at(uint32_t a, size_t n) {
  if ((size_t)a + n > g_t.cta->smem_bytes) { fprintf(stderr, "emu: shared access out of bounds (%u + %zu)\n", a, n); abort(); }
  return g_t.cta->smem + a;
}
inline float ex2(float x) { return exp2f(x); }
inline float lg2(float x) { return log2f(x); }
inline float tanh_approx(float x) { return tanhf(x); }

inline float4 lds128(uint32_t a) { if (a & 15) abort(); float4 v; memcpy(&v, smem_at(a, 16), 16); return v; }
inline float lds32(uint32_t a) { if (a & 3) abort(); float v; memcpy(&v, smem_at(a, 4), 4); return v; }
inline void sts32(uint32_t a, float v) { if (a & 3) abort(); memcpy(smem_at(a, 4), &v, 4); }

inline void mbar_init(uint64_t* bar, int count) {
  std::lock_guard<std::mutex> g(g_t.cta->m);
  Mbar& b = g_t.cta->mbars[bar];
  b.init = b.pending = count; b.tx = 0; b.phase = 0;
}
inline void mbar_complete_locked(Mbar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1; b.pending = b.init; g_t.cta->cv.notify_all(); }
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {       // arrive.expect_tx
  std::lock_guard<std::mutex> g(g_t.cta->m);
  Mbar& b = g_t.cta->mbars.at(bar);
  if (b.pending <= 0) { fprintf(stderr, "emu: mbarrier over-arrival\n"); abort(); }
  b.tx += bytes; b.pending -= 1;
  mbar_complete_locked(b);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {            // try_wait.parity loop
  std::unique_lock<std::mutex> g(g_t.cta->m);
  Mbar& b = g_t.cta->mbars.at(bar);
  if (!g_t.cta->cv.wait_for(g, std::chrono::seconds(20), [&] { return b.phase != parity; })) {
    fprintf(stderr, "emu: mbarrier wait timed out (deadlock): tid %d parity %u phase %u\n", g_t.tid, parity, b.phase);
    abort();
  }
}
inline void mbar_wait_wd(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }   // device: the same wait with a watchdog
// ---- math of common.cuh (exact libm in place of the MUFU approximations) ----------------------------------------
inline float rcp(float x) { return 1.0f / x; }
inline float silu(float v) { return v * rcp(1.0f + ex2(-kLog2e * v)); }
template <typename T>
inline float silu_io(float v) {
  if (sizeof(T) == 2) { const float h = 0.5f * v; return fmaf(h, tanh_approx(h), h); }
  return silu(v);
}
inline float softplus(float v) {
  const float w = ex2(kLog2e * v);
  float sp = kLn2 * lg2(1.0f + w);
  const float series = w * (1.0f - w * (0.5f - w * (0.33333334f - 0.25f * w)));
  sp = (w < 0.015625f) ? series : sp;
  return v > 20.0f ? v : sp;
}

namespace simt {
inline float2 fma2(const float2& a, const float2& b, const float2& c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 mul2(const float2& a, const float2& b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 add2(const float2& a, const float2& b) { return make_float2(a.x + b.x, a.y + b.y); }
inline uint4 lds128u(uint32_t a) { if (a & 15) abort(); uint4 v; memcpy(&v, smem_at(a, 16), 16); return v; }
inline uint32_t lds16u(uint32_t a) { if (a & 1) abort(); uint16_t v; memcpy(&v, smem_at(a, 2), 2); return v; }
inline void sts16u(uint32_t a, uint32_t v) { if (a & 1) abort(); const uint16_t w = (uint16_t)v; memcpy(smem_at(a, 2), &w, 2); }
inline void sts32u(uint32_t a, uint32_t v) { if (a & 3) abort(); std::lock_guard<std::mutex> g(g_t.cta->m); memcpy(smem_at(a, 4), &v, 4); }
inline uint32_t atomic_inc_shared(uint32_t a) {
  std::lock_guard<std::mutex> g(g_t.cta->m);
  uint32_t old; memcpy(&old, smem_at(a, 4), 4);
  const uint32_t nw = old + 1; memcpy(smem_at(a, 4), &nw, 4);
  return old;
}
// cp.async is modelled as an immediate copy (the kernel only reads staged data after cp.async.wait_group, from the lane that
// issued the copy, so ordering is not at stake; alignment is)
inline void cp_async16s(uint32_t a, const void* g) {
  if ((a & 15) || ((uintptr_t)g & 15)) { fprintf(stderr, "emu: misaligned cp.async\n"); abort(); }
  memcpy(smem_at(a, 16), g, 16);
}
inline void cp_commit() {}
template <int N> inline void cp_wait_group() {}
// 1-D bulk copy: synchronous here; alignment and bounds as the TMA engine wants them
inline void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  if ((smem_dst & 15) || ((uintptr_t)gsrc & 15) || (bytes & 15)) { fprintf(stderr, "emu: misaligned bulk copy\n"); abort(); }
  memcpy(smem_at(smem_dst, bytes), gsrc, bytes);
  std::lock_guard<std::mutex> g(g_t.cta->m);
  Mbar& mb = g_t.cta->mbars.at(bar);
  mb.tx -= (long)bytes;
  mbar_complete_locked(mb);
}
inline void cta_sync() { pthread_barrier_wait(&g_t.cta->cta_bar); }
inline void warp_sync() { pthread_barrier_wait(&g_t.cta->warp_bar[g_t.tid >> 5]); }
// __all_sync: two flags per warp used alternately; lane 0 clears the one just read before it can reach the call after next
inline bool warp_all(bool p) {
  EmuCta* c = g_t.cta;
  const int w = g_t.tid >> 5, e = g_t.vote_epoch++ & 1;
  if (!p) c->vote[2 * w + e].store(1);
  pthread_barrier_wait(&c->warp_bar[w]);
  const bool r = c->vote[2 * w + e].load() == 0;
  pthread_barrier_wait(&c->warp_bar[w]);
  if ((g_t.tid & 31) == 0) c->vote[2 * w + e].store(0);
  return r;
}
inline void stg128(void* p, const uint4& v) {
  if ((uintptr_t)p & 15) { fprintf(stderr, "emu: misaligned 16-byte global store\n"); abort(); }
  memcpy(p, &v, 16);
}
inline void stg128f(float* p, float a, float b, float c, float d) {
  if ((uintptr_t)p & 15) { fprintf(stderr, "emu: misaligned 16-byte global store\n"); abort(); }
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
}
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
}  // namespace simt
}  // namespace cad

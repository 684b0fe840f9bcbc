// Host-side SIMT emulation of the primitives the v4 scan kernel is written against (TEST INFRASTRUCTURE ONLY).
//
// caduceus_b200/csrc/scan_fwd_v4.cuh is compiled for the CPU with -DCAD_EMULATE: every CUDA thread becomes an OS
// thread, a warp shuffle is an exchange through a per-warp buffer between two barriers, __syncthreads is a CTA
// barrier, shared memory is a byte array addressed by 32-bit offsets, mbarriers keep (pending arrivals, pending
// bytes, phase), and the TMA load is a synchronous copy that applies the SWIZZLE_128B pattern the device kernel
// assumes (16-byte chunk c of 128-byte line l lands at chunk c ^ (l & 7)) with zero fill outside the tensor.
// The point is to check the kernel's INDEX LOGIC (reversed jobs, tails, conv neighbours, carries, barrier phases)
// against the oracle where no GPU is available; arithmetic uses exact exp2f/tanhf/log2f instead of MUFU.
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

struct EmuTmap {              // (128-byte line, blocks, rows) view of an (nrows, ld) matrix, as make_row_tile_map[16]
  const void* base;
  int64_t nrows, ld, nblk;    // ld in elements; nblk = ceil(L / tokens per line)
  int box_blocks, box_rows;
  int elem_bytes = 4;         // 4: 32 tokens per line; 2: 64 tokens per line
};

struct Mbar { int init = 0, pending = 0; long tx = 0; uint32_t phase = 0; };

struct EmuCta {
  unsigned char* smem = nullptr;
  size_t smem_bytes = 0;
  int nthreads = 0;
  pthread_barrier_t cta_bar;
  std::vector<pthread_barrier_t> warp_bar;
  std::vector<float> xchg;                 // [warps][32]
  std::mutex m;
  std::condition_variable cv;
  std::map<const void*, Mbar> mbars;
  bool deadlock = false;
  std::atomic<int> or_flag[2] = {{0}, {0}};   // __syncthreads_or: alternating per call
};

struct EmuThread { EmuCta* cta; int tid, bx, by, bz = 0; int or_epoch = 0; };
extern thread_local EmuThread g_t;

namespace v4 {
#define CAD_DEV inline
#define CAD_TID (::cad::g_t.tid)
#define CAD_NTHREADS (::cad::g_t.cta->nthreads)
#define CAD_BIDX (::cad::g_t.bx)
#define CAD_BIDY (::cad::g_t.by)
#define __restrict__
typedef EmuTmap tmap_t;
}  // namespace v4

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
inline void tma_load_3d(void* smem_dst, const EmuTmap* t, int c0, int c1, int c2, uint64_t* bar) {
  if (c0 != 0 || (smem_u32(smem_dst) & 1023)) { fprintf(stderr, "emu: bad TMA destination / coordinate\n"); abort(); }
  unsigned char* dst = (unsigned char*)smem_dst;
  smem_at(smem_u32(smem_dst), (size_t)t->box_rows * t->box_blocks * 128);
  const int eb = t->elem_bytes, per_line = 128 / eb, per_piece = 16 / eb;
  for (int r = 0; r < t->box_rows; ++r)
    for (int b = 0; b < t->box_blocks; ++b) {
      const int line = r * t->box_blocks + b;
      for (int j = 0; j < per_line; ++j) {
        const int64_t row = (int64_t)c2 + r, blk = (int64_t)c1 + b;
        unsigned char v[4] = {0, 0, 0, 0};
        if (row >= 0 && row < t->nrows && blk >= 0 && blk < t->nblk)
          memcpy(v, (const unsigned char*)t->base + ((size_t)row * t->ld + (size_t)blk * per_line + j) * eb, eb);
        const int chunk = (j / per_piece) ^ (line & 7);
        memcpy(dst + (size_t)line * 128 + chunk * 16 + (j % per_piece) * eb, v, eb);
      }
    }
  std::lock_guard<std::mutex> g(g_t.cta->m);
  Mbar& mb = g_t.cta->mbars.at(bar);
  mb.tx -= (long)t->box_rows * t->box_blocks * 128;
  mbar_complete_locked(mb);
}

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

template <int TOK_>
inline void tile_piece_offsets(int seg, uint32_t (&poff)[TOK_ / 4]) {      // as scan_common.cuh
  constexpr int SPB = 32 / TOK_;
  const int blk = seg / SPB, c0 = (TOK_ / 4) * (seg % SPB);
  for (int k = 0; k < TOK_ / 4; ++k) poff[k] = blk * 128 + (((c0 + k) ^ (blk & 7)) << 4);
}

namespace v4 {
inline float2 fma2(const float2& a, const float2& b, const float2& c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 mul2(const float2& a, const float2& b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 add2(const float2& a, const float2& b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 lds64(uint32_t a) { if (a & 7) abort(); float2 v; memcpy(&v, smem_at(a, 8), 8); return v; }
inline void sts64(uint32_t a, const float2& v) { if (a & 7) abort(); memcpy(smem_at(a, 8), &v, 8); }
inline uint4 lds128u(uint32_t a) { if (a & 15) abort(); uint4 v; memcpy(&v, smem_at(a, 16), 16); return v; }
// cp.async is modelled as an immediate copy (the kernel only reads staged data after cp.async.wait_group 0, from
// the lane that issued the copy, so ordering is not at stake; alignment is)
inline void cp_async16s(uint32_t a, const void* g) {
  if ((a & 15) || ((uintptr_t)g & 15)) { fprintf(stderr, "emu: misaligned cp.async\n"); abort(); }
  memcpy(smem_at(a, 16), g, 16);
}
inline void cp_commit() {}
inline void cp_wait_all() {}
inline void cta_sync() { pthread_barrier_wait(&g_t.cta->cta_bar); }
inline float shfl_raw(float v, int src_lane) {
  EmuCta* c = g_t.cta;
  const int w = g_t.tid >> 5, l = g_t.tid & 31;
  c->xchg[w * 32 + l] = v;
  pthread_barrier_wait(&c->warp_bar[w]);
  const float r = c->xchg[w * 32 + src_lane];
  pthread_barrier_wait(&c->warp_bar[w]);
  return r;
}
inline float shfl_up1(float v, int off) { const int l = g_t.tid & 31; return shfl_raw(v, l >= off ? l - off : l); }
inline float2 shfl_up2(const float2& v, int off) { return make_float2(shfl_up1(v.x, off), shfl_up1(v.y, off)); }
inline float2 shfl_idx2(const float2& v, int src) { return make_float2(shfl_raw(v.x, src), shfl_raw(v.y, src)); }
inline void stg128(void* p, const uint4& v) {
  if ((uintptr_t)p & 15) { fprintf(stderr, "emu: misaligned 16-byte global store\n"); abort(); }
  memcpy(p, &v, 16);
}
template <int OFF>
inline void scan_step_up2(float2& P, float2& H, int lane) {
  const float2 Pp = shfl_up2(P, OFF), Hp = shfl_up2(H, OFF);
  if (lane >= OFF) {
    H = make_float2(fmaf(P.x, Hp.x, H.x), fmaf(P.y, Hp.y, H.y));
    P = make_float2(P.x * Pp.x, P.y * Pp.y);
  }
}
}  // namespace v4

namespace v9 {
inline uint32_t atomic_inc_shared(uint32_t a) {
  std::lock_guard<std::mutex> g(g_t.cta->m);
  uint32_t old; memcpy(&old, smem_at(a, 4), 4);
  const uint32_t nw = old + 1; memcpy(smem_at(a, 4), &nw, 4);
  return old;
}
inline void sts32u(uint32_t a, uint32_t v) { if (a & 3) abort(); std::lock_guard<std::mutex> g(g_t.cta->m); memcpy(smem_at(a, 4), &v, 4); }
inline void warp_sync() { pthread_barrier_wait(&g_t.cta->warp_bar[g_t.tid >> 5]); }
inline float shfl_up1(float v, int off) { return v4::shfl_up1(v, off); }
inline float shfl_idx1(float v, int src) { return v4::shfl_raw(v, src); }
inline float shfl_xor1(float v, int m) { return v4::shfl_raw(v, (g_t.tid & 31) ^ m); }
template <int OFF>
inline void scan_step1(float& P, float& H, int lane) {
  const float Pp = shfl_up1(P, OFF), Hp = shfl_up1(H, OFF);
  if (lane >= OFF) { H = fmaf(P, Hp, H); P = P * Pp; }
}
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
}  // namespace v9

namespace v20 {
#define CAD_BIDZ (::cad::g_t.bz)
// 1-D bulk copy: synchronous here; alignment and bounds as the TMA engine wants them
inline void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  if ((smem_dst & 15) || ((uintptr_t)gsrc & 15) || (bytes & 15)) { fprintf(stderr, "emu: misaligned bulk copy\n"); abort(); }
  memcpy(smem_at(smem_dst, bytes), gsrc, bytes);
  std::lock_guard<std::mutex> g(g_t.cta->m);
  Mbar& mb = g_t.cta->mbars.at(bar);
  mb.tx -= (long)bytes;
  mbar_complete_locked(mb);
}
template <int N> inline void cp_wait_group() {}        // cp.async is an immediate copy here
inline uint32_t lds16u(uint32_t a) { if (a & 1) abort(); uint16_t v; memcpy(&v, smem_at(a, 2), 2); return v; }
inline void sts16u(uint32_t a, uint32_t v) { if (a & 1) abort(); const uint16_t w = (uint16_t)v; memcpy(smem_at(a, 2), &w, 2); }
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline void stg128f(float* p, float a, float b, float c, float d) {
  if ((uintptr_t)p & 15) { fprintf(stderr, "emu: misaligned 16-byte global store\n"); abort(); }
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
}
}  // namespace v20

namespace fx {        // scan_fixup.cuh
constexpr int kTok = 16, kChunk = 512, kMaxG = 7;
// __syncthreads_or: two flags used alternately; thread 0 clears the one just read before it can reach the call after next
inline bool cta_sync_or(bool p) {
  EmuCta* c = g_t.cta;
  const int e = g_t.or_epoch++ & 1;
  if (p) c->or_flag[e].store(1);
  pthread_barrier_wait(&c->cta_bar);
  const bool r = c->or_flag[e].load() != 0;
  pthread_barrier_wait(&c->cta_bar);
  if (g_t.tid == 0) c->or_flag[e].store(0);
  return r;
}
template <typename T, int V>
inline void load_vec(const T* p, float (&v)[V]) {
  if ((uintptr_t)p & 15) { fprintf(stderr, "emu: misaligned vector load\n"); abort(); }
  for (int i = 0; i < V; ++i) v[i] = io<T>::to_f(p[i]);
}
template <typename T, int V>
inline void store_vec(T* p, const float (&v)[V]) {
  if ((uintptr_t)p & 15) { fprintf(stderr, "emu: misaligned vector store\n"); abort(); }
  for (int i = 0; i < V; ++i) p[i] = io<T>::from_f(v[i]);
}
}  // namespace fx

namespace bw2 {
inline float shfl_down1(float v, int off) { const int l = g_t.tid & 31; return v4::shfl_raw(v, l + off < 32 ? l + off : l); }
template <int OFF>
inline void scan_step_dn1(float& Q, float& E, int lane) {
  const float Qn = shfl_down1(Q, OFF), En = shfl_down1(E, OFF);
  if (lane < 32 - OFF) { E = fmaf(Q, En, E); Q = Q * Qn; }
}
inline void sts128f(uint32_t a, float x, float y, float z, float w) {
  if (a & 15) abort();
  const float v[4] = {x, y, z, w};
  memcpy(smem_at(a, 16), v, 16);
}
// global atomics: one process-wide mutex (CTAs run one after another, threads of a CTA concurrently)
inline std::mutex& global_mutex() { static std::mutex m; return m; }
inline void red_add_v4(float* addr, float4 v) {
  if ((uintptr_t)addr & 15) { fprintf(stderr, "emu: misaligned red.global.add.v4\n"); abort(); }
  std::lock_guard<std::mutex> g(global_mutex());
  addr[0] += v.x; addr[1] += v.y; addr[2] += v.z; addr[3] += v.w;
}
inline void atomic_add_f32(float* addr, float v) { std::lock_guard<std::mutex> g(global_mutex()); *addr += v; }
}  // namespace bw2
}  // namespace cad

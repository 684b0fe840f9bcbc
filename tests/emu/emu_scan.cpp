// Host driver of the SIMT emulation of the v4 / v9 scan kernels (TEST INFRASTRUCTURE ONLY; see simt_emu.h).
//   g++ -O1 -std=c++17 -shared -fPIC -pthread -DCAD_EMULATE -I tests/emu -I $CUDA/include tests/emu/emu_scan.cpp
// All pointers in cad_scan_fwd_args are HOST pointers here.
#include <stdio.h>
#include <stdlib.h>

#define CAD_EMULATE 1
#include "simt_emu.h"
#include "../../caduceus_b200/csrc/scan_fwd_v4.cuh"
#include "../../caduceus_b200/csrc/scan_fwd_v9.cuh"
#include "../../caduceus_b200/csrc/scan_bwd_v2.cuh"
#include "../../caduceus_b200/csrc/scan_fwd_v20.cuh"
#include "../../caduceus_b200/csrc/scan_fixup.cuh"

namespace cad {
thread_local EmuThread g_t;
}

template <typename Body>
static void run_cta(size_t smem_bytes, int G, int bx, int by, Body body, int bz = 0) {
  using namespace cad;
  EmuCta cta;
  cta.smem_bytes = smem_bytes;
  void* mem = nullptr;
  if (posix_memalign(&mem, 1024, cta.smem_bytes) != 0) abort();
  memset(mem, 0xCD, cta.smem_bytes);               // poison: reading unstaged shared memory shows up as garbage
  cta.smem = (unsigned char*)mem;
  cta.nthreads = G * 32;
  pthread_barrier_init(&cta.cta_bar, nullptr, cta.nthreads);
  cta.warp_bar.resize(G);
  for (int w = 0; w < G; ++w) pthread_barrier_init(&cta.warp_bar[w], nullptr, 32);
  cta.xchg.assign((size_t)G * 32, 0.f);
  std::vector<std::thread> th;
  for (int t = 0; t < cta.nthreads; ++t)
    th.emplace_back([&, t] {
      g_t = EmuThread{&cta, t, bx, by, bz};
      body(cta.smem);
    });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&cta.cta_bar);
  for (int w = 0; w < G; ++w) pthread_barrier_destroy(&cta.warp_bar[w]);
  free(mem);
}

extern "C" int emu_scan_v4(const cad_scan_fwd_args* a, int G) {
  using namespace cad;
  if (a->N != v4::NST || a->E % 2 || G < 1 || G > v4::kMaxG4 || a->io_dtype == CAD_F32) return -1;
  if (a->L <= 0) return 0;
  EmuTmap tmap;
  tmap.base = a->bc;
  tmap.nrows = (int64_t)a->njobs * 2 * v4::NST;
  tmap.ld = a->ldbc;
  tmap.nblk = (a->L + 31) / 32;
  tmap.box_blocks = v4::CH / 32;
  tmap.box_rows = 2 * v4::NST;
  const int gx = (int)((a->E / 2 + G - 1) / G);
  for (int by = 0; by < a->njobs; ++by)
    for (int bx = 0; bx < gx; ++bx) {
      if (a->io_dtype == CAD_BF16)
        run_cta(v4::smem_bytes(G, 2), G, bx, by, [&](unsigned char* sm) { v4::kernel_body<__nv_bfloat16>(*a, &tmap, sm); });
      else
        run_cta(v4::smem_bytes(G, 2), G, bx, by, [&](unsigned char* sm) { v4::kernel_body<__half>(*a, &tmap, sm); });
    }
  return 0;
}

template <typename T, typename TT>
static void run_v9(const cad_scan_fwd_args* a, const cad::EmuTmap* tmap, int G, int bx, int by, int pipe) {
  using namespace cad;
  const size_t sb = v9::smem_bytes(G, sizeof(T), sizeof(TT));
  if (a->state_only) {
    if (pipe) run_cta(sb, G, bx, by, [&](unsigned char* sm) { v9::kernel_body<T, TT, true, true>(*a, tmap, sm); });
    else      run_cta(sb, G, bx, by, [&](unsigned char* sm) { v9::kernel_body<T, TT, true, false>(*a, tmap, sm); });
  } else {
    if (pipe) run_cta(sb, G, bx, by, [&](unsigned char* sm) { v9::kernel_body<T, TT, false, true>(*a, tmap, sm); });
    else      run_cta(sb, G, bx, by, [&](unsigned char* sm) { v9::kernel_body<T, TT, false, false>(*a, tmap, sm); });
  }
}

// tile32 = 0: 16-bit tile from a->bc16 (variants 9 / 10); tile32 = 1: fp32 tile from a->bc (variants 11 / 12)
extern "C" int emu_scan_v9(const cad_scan_fwd_args* a, int G, int pipe, int tile32) {
  using namespace cad;
  if (a->N != v9::NST || G < 1 || G > (tile32 ? v9::kMaxG9 : 7) || a->io_dtype == CAD_F32) return -1;
  if (tile32 ? (!a->bc || a->ldbc % 32) : (!a->bc16 || a->ldbc16 % 64)) return -1;
  if (a->L <= 0) return 0;
  EmuTmap tmap;
  tmap.base = tile32 ? (const void*)a->bc : a->bc16;
  tmap.elem_bytes = tile32 ? 4 : 2;
  tmap.nrows = (int64_t)a->njobs * 2 * v9::NST;
  tmap.ld = tile32 ? a->ldbc : a->ldbc16;
  const int line = 128 / tmap.elem_bytes;
  tmap.nblk = (a->L + line - 1) / line;
  tmap.box_blocks = v9::CH / line;
  tmap.box_rows = 2 * v9::NST;
  const int gx = (int)((a->E + G - 1) / G);
  for (int by = 0; by < a->njobs; ++by)
    for (int bx = 0; bx < gx; ++bx) {
      if (a->io_dtype == CAD_BF16) {
        if (tile32) run_v9<__nv_bfloat16, float>(a, &tmap, G, bx, by, pipe);
        else run_v9<__nv_bfloat16, __nv_bfloat16>(a, &tmap, G, bx, by, pipe);
      } else {
        if (tile32) run_v9<__half, float>(a, &tmap, G, bx, by, pipe);
        else run_v9<__half, __half>(a, &tmap, G, bx, by, pipe);
      }
    }
  return 0;
}

extern "C" int emu_scan_bwd_v2(const cad_scan_bwd_args* a, int G) {
  using namespace cad;
  if (a->N != bw2::NST || G < 1 || G > bw2::kMaxG || a->ldbc % 32) return -1;
  if (a->L <= 0) return 0;
  EmuTmap tmap;
  tmap.base = a->bc;
  tmap.elem_bytes = 4;
  tmap.nrows = (int64_t)a->njobs * 2 * bw2::NST;
  tmap.ld = a->ldbc;
  tmap.nblk = (a->L + 31) / 32;
  tmap.box_blocks = bw2::CH / 32;
  tmap.box_rows = 2 * bw2::NST;
  const int gx = (int)((a->E + G - 1) / G);
  const size_t sb = bw2::smem_bytes();
  for (int by = 0; by < a->njobs; ++by)
    for (int bx = 0; bx < gx; ++bx) {
      if (a->io_dtype == CAD_BF16) run_cta(sb, G, bx, by, [&](unsigned char* sm) { bw2::kernel_body<__nv_bfloat16>(*a, &tmap, sm); });
      else if (a->io_dtype == CAD_F16) run_cta(sb, G, bx, by, [&](unsigned char* sm) { bw2::kernel_body<__half>(*a, &tmap, sm); });
      else run_cta(sb, G, bx, by, [&](unsigned char* sm) { bw2::kernel_body<float>(*a, &tmap, sm); });
    }
  return 0;
}

// variant 20 (lane = channel): W warps of 32 channels per CTA, grid (channel groups, jobs, segments)
template <typename T>
static void run_v20(const cad_scan_fwd_args* a, size_t sb, int W, int bx, int by, int bz) {
  using namespace cad;
  switch (a->variant - 20) {
    case 1: run_cta(sb, W, bx, by, [&](unsigned char* sm) { v20::kernel_body<T, 1>(*a, sm); }, bz); break;
    case 2: run_cta(sb, W, bx, by, [&](unsigned char* sm) { v20::kernel_body<T, 2>(*a, sm); }, bz); break;
    case 3: run_cta(sb, W, bx, by, [&](unsigned char* sm) { v20::kernel_body<T, 3>(*a, sm); }, bz); break;
    default: run_cta(sb, W, bx, by, [&](unsigned char* sm) { v20::kernel_body<T, 0>(*a, sm); }, bz); break;
  }
}
extern "C" int emu_scan_v20(const cad_scan_fwd_args* a, int W) {
  using namespace cad;
  if (a->N != v20::NST || W < 1 || W > v20::kMaxW || a->io_dtype == CAD_F32 || !a->bcT) return -1;
  if (a->L <= 0) return 0;
  const int nseg = a->nseg > 0 ? a->nseg : 1;
  const int gx = (int)(((a->E + 31) / 32 + W - 1) / W);
  const size_t sb = v20::smem_bytes(W);
  for (int bz = 0; bz < nseg; ++bz)
    for (int by = 0; by < a->njobs; ++by)
      for (int bx = 0; bx < gx; ++bx) {
        if (a->io_dtype == CAD_BF16) run_v20<__nv_bfloat16>(a, sb, W, bx, by, bz);
        else run_v20<__half>(a, sb, W, bx, by, bz);
      }
  return 0;
}

// carry fix-up (scan_fixup.cuh): whole-sequence mode (grid.y = jobs) or segment mode (grid.y = jobs x (nseg - 1))
extern "C" int emu_scan_fixup(const cad_scan_fixup_args* a, int G) {
  using namespace cad;
  if (a->N != 16 || G < 1 || G > fx::kMaxG || a->ldbc % 32) return -1;
  if (a->L <= 0) return 0;
  EmuTmap tmap;
  tmap.base = a->bc;
  tmap.elem_bytes = 4;
  tmap.nrows = (int64_t)a->njobs * 32;
  tmap.ld = a->ldbc;
  tmap.nblk = (a->L + 31) / 32;
  tmap.box_blocks = fx::kChunk / 32;
  tmap.box_rows = 16;
  const int gx = (int)((a->E + G - 1) / G), gy = a->nseg > 1 ? a->njobs * (a->nseg - (a->seg_first ? 0 : 1)) : a->njobs;
  const size_t sb = 1024 + (size_t)16 * fx::kChunk * 4 + (size_t)2 * fx::kMaxG * 16 * 4 + 16;
  for (int by = 0; by < gy; ++by)
    for (int bx = 0; bx < gx; ++bx) {
      if (a->io_dtype == CAD_BF16) run_cta(sb, G, bx, by, [&](unsigned char* sm) { fx::kernel_body<__nv_bfloat16, 16>(*a, &tmap, sm); });
      else if (a->io_dtype == CAD_F16) run_cta(sb, G, bx, by, [&](unsigned char* sm) { fx::kernel_body<__half, 16>(*a, &tmap, sm); });
      else run_cta(sb, G, bx, by, [&](unsigned char* sm) { fx::kernel_body<float, 16>(*a, &tmap, sm); });
    }
  return 0;
}

// Host driver of the SIMT emulation of the lane = channel scan kernel (TEST INFRASTRUCTURE ONLY; see simt_emu.h).
//   g++ -O1 -std=c++17 -shared -fPIC -pthread -DCAD_EMULATE -I tests/emu -I $CUDA/include tests/emu/emu_scan.cpp
// All pointers in cad_scan_fwd_args are HOST pointers here.
#include <stdio.h>
#include <stdlib.h>

#define CAD_EMULATE 1
#include "simt_emu.h"
#include "../../caduceus_b200/csrc/scan_fwd_v20.cuh"

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
  cta.vote = std::vector<std::atomic<int>>(2 * G);
  for (auto& v : cta.vote) v.store(0);
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

// variant 20 (lane = channel): W warps of 32 channels per CTA, grid (channel groups, jobs, segments)
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
        if (a->io_dtype == CAD_BF16) run_cta(sb, W, bx, by, [&](unsigned char* sm) { v20::kernel_body<__nv_bfloat16>(*a, sm); }, bz);
        else run_cta(sb, W, bx, by, [&](unsigned char* sm) { v20::kernel_body<__half>(*a, sm); }, bz);
      }
  return 0;
}

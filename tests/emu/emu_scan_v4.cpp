// Host driver of the SIMT emulation of the v4 scan kernel (TEST INFRASTRUCTURE ONLY; see simt_emu.h).
//   g++ -O1 -std=c++17 -shared -fPIC -pthread -DCAD_EMULATE -I tests/emu -I $CUDA/include tests/emu/emu_scan_v4.cpp
// All pointers in cad_scan_fwd_args are HOST pointers here.
#include <stdio.h>
#include <stdlib.h>

#define CAD_EMULATE 1
#include "simt_emu.h"
#include "../../caduceus_b200/csrc/scan_fwd_v4.cuh"

namespace cad {
thread_local EmuThread g_t;
}

template <typename T>
static void run_cta(const cad_scan_fwd_args* a, const cad::EmuTmap* tmap, int G, int bx, int by) {
  using namespace cad;
  EmuCta cta;
  cta.smem_bytes = v4::smem_bytes(G, sizeof(T));
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
      g_t = EmuThread{&cta, t, bx, by};
      v4::kernel_body<T>(*a, tmap, cta.smem);
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
      if (a->io_dtype == CAD_BF16) run_cta<__nv_bfloat16>(a, &tmap, G, bx, by);
      else run_cta<__half>(a, &tmap, G, bx, by);
    }
  return 0;
}

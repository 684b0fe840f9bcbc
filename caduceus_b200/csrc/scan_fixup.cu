// Carry fix-up for a sequence-sharded scan: kernel entry, launcher and C-ABI (the kernel body is scan_fixup.cuh).
#include "scan_fixup.cuh"

namespace cad {

template <typename T, int N>
__global__ void __launch_bounds__(kMaxG * 32, 2)
scan_fixup_kernel(const cad_scan_fixup_args a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ unsigned char smem_raw[];
  fx::kernel_body<T, N>(a, &tmap, smem_raw);
}

template <typename T, int N>
static int launch_fixup(const cad_scan_fixup_args& a, int G, cudaStream_t stream) {
  CUtensorMap tmap;
  if (make_row_tile_map(&tmap, a.bc, (int64_t)a.njobs * 2 * N, a.ldbc, a.L, N) != 0) return -1;
  const size_t smem = 1024 + (size_t)N * kChunk * 4 + (size_t)2 * kMaxG * N * sizeof(float) + 16;
  auto kern = scan_fixup_kernel<T, N>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)((a.E + G - 1) / G), (unsigned)(a.nseg > 1 ? a.njobs * (a.nseg - (a.seg_first ? 0 : 1)) : a.njobs));
  kern<<<grid, G * 32, smem, stream>>>(a, tmap);
  CAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace cad

extern "C" int cad_bimamba_scan_fixup(const cad_scan_fixup_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_bimamba_scan_fixup: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_bimamba_scan_fixup: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->delta && a->bc && a->out && a->dt_b && a->A2 && a->seq_of_job && a->pset_of_job &&
              a->rev_of_job && (a->nseg > 1 ? a->seg_carry != nullptr : a->h0 != nullptr), "cad_bimamba_scan_fixup: null pointer");
  CAD_REQUIRE(a->nseg >= 0 && a->nseg <= 4096, "cad_bimamba_scan_fixup: nseg out of range");
  CAD_REQUIRE(a->N == 16, "cad_bimamba_scan_fixup: d_state = %lld not built (only 16)", (long long)a->N);
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldd % 16 == 0 && a->ldo % 16 == 0 && a->ldbc % 32 == 0,
              "cad_bimamba_scan_fixup: bad row pitches");
  CAD_REQUIRE(a->cutoff_log2 < 0.f, "cad_bimamba_scan_fixup: cutoff_log2 must be negative");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int G = a->channels_per_cta;
  if (G <= 0) G = kMaxG;
  CAD_REQUIRE(G >= 1 && G <= kMaxG, "cad_bimamba_scan_fixup: channels_per_cta must be in [1, %d]", kMaxG);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, return launch_fixup<T, 16>(*a, G, stream));
  return 0;
}

// Windowed mean pooling of last-layer hidden states around a variant: the reduction the reference's
// VEP dump loop does per batch with arange + clamp + gather (B, 1537, C) + mean (ref:vep_embeddings.py:278-311), and with
// `.contiguous().flip(dims=[1, 2])` copies of the (B, L, C) outputs for the reverse-complement view (ref:vep_embeddings.py:355-366).
// Here: one launch per pooled tensor, no gathered window, no flipped copies, no host round trip for the variant index —
//   V[b, r, c]  = hidden[b, flip_len ? L-1-r : r, c0 + (flip_ch ? C-1-c : c)]                       (the view being pooled)
//   out[b, c]   = mean over j in [idx_b - lo_half, idx_b + half] of V[b, clamp(j, 0, L-1), c]         (duplicates counted, as gather does)
// fp32 accumulation, result in the io dtype.  One thread per (batch row, channel); a warp reads 32 consecutive channels of a row.
#include "common.cuh"

namespace cad {

template <typename T>
__global__ void __launch_bounds__(128) window_mean_kernel(cad_window_mean_args a) {
  const int64_t b = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const T* __restrict__ H = static_cast<const T*>(a.hidden) + b * a.L * a.ldh + a.c0 + (a.flip_ch ? a.C - 1 - c : c);
  const int64_t idx = a.variant_idx[b];
  const int64_t lo = idx - a.lo_half, hi = idx + a.half;
  float s = 0.f;
  for (int64_t j = lo; j <= hi; ++j) {
    int64_t r = j < 0 ? 0 : (j > a.L - 1 ? a.L - 1 : j);
    if (a.flip_len) r = a.L - 1 - r;
    s += io<T>::to_f(H[r * a.ldh]);
  }
  static_cast<T*>(a.out)[b * a.ldo + c] = io<T>::from_f(s / (float)(hi - lo + 1));
}

}  // namespace cad

extern "C" int cad_window_mean(const cad_window_mean_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_window_mean: null argument block");
  CAD_REQUIRE(a->B >= 0 && a->L > 0 && a->C > 0 && a->ldh >= a->c0 + a->C && a->c0 >= 0 && a->ldo >= a->C, "cad_window_mean: bad sizes");
  CAD_REQUIRE(a->lo_half >= 0 && a->half >= 0, "cad_window_mean: bad window");
  if (a->B == 0) return 0;
  CAD_REQUIRE(a->hidden && a->variant_idx && a->out, "cad_window_mean: null pointer");
  dim3 grid((unsigned)((a->C + 127) / 128), (unsigned)a->B);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, (window_mean_kernel<T><<<grid, 128, 0, stream>>>(*a)));
  CAD_LAUNCH_CHECK();
  return 0;
}

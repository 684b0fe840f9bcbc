// Depthwise (anti)causal conv (K <= 4) + SiLU, materialised per job: the operand of the x_proj GEMM.
//
// Replaces causal_conv1d_cuda.causal_conv1d_fwd (upstream, reached via ref:caduceus/modeling_caduceus.py:128-133)
// and — for reversed jobs — the `hidden_states.flip(dims=(1,))` of ref:caduceus/modeling_caduceus.py:131: the
// reverse Mamba's causal conv over the flipped sequence is an ANTI-causal conv in original coordinates
// (SURVEY.md A.5), so no flipped copy is made.
//
//   u[tau] = silu(b + sum_{k<4} w[k] * x[tau - 3 + k]),  tau logical time; physical t = tau or L-1-tau.
#include "common.cuh"

namespace cad {

template <typename T>
__global__ void __launch_bounds__(256) conv_silu_fwd_kernel(cad_conv_fwd_args a) {
  const int job = blockIdx.z;
  const int64_t ch = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  const T* __restrict__ x = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * a.E + ch) * a.ldxz;
  T* __restrict__ u = static_cast<T*>(a.u) + ((int64_t)job * a.E + ch) * a.ldu;
  const float* w = a.conv_w + ((int64_t)pset * a.E + ch) * 4;
  const float w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
  const float bias = a.conv_b[(int64_t)pset * a.E + ch];
  const T* halo = a.halo ? static_cast<const T*>(a.halo) + ((int64_t)job * a.E + ch) * 3 : nullptr;
  const int64_t L = a.L;

  const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (t0 >= L) return;
  // window of 11 physical samples: fwd needs x[t0-3 .. t0+7]; rev needs x[t0 .. t0+10]
  float win[11];
  const int64_t base = rev ? t0 : t0 - 3;
#pragma unroll
  for (int i = 0; i < 11; ++i) {
    const int64_t t = base + i;
    float v = 0.f;
    if (t >= 0 && t < L) v = io<T>::to_f(x[t]);
    else if (halo) {
      // logical index of physical t: tau = t (fwd) or L-1-t (rev); halo holds tau = -3, -2, -1
      const int64_t tau = rev ? (L - 1 - t) : t;
      if (tau >= -3 && tau < 0) v = io<T>::to_f(halo[tau + 3]);
    }
    win[i] = v;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (t0 + i < L) {
      float acc;
      if (!rev) acc = bias + w0 * win[i] + w1 * win[i + 1] + w2 * win[i + 2] + w3 * win[i + 3];
      else      acc = bias + w3 * win[i] + w2 * win[i + 1] + w1 * win[i + 2] + w0 * win[i + 3];
      u[t0 + i] = io<T>::from_f(silu(acc));
    }
  }
}

}  // namespace cad

extern "C" int cad_conv_silu_fwd(const cad_conv_fwd_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a && a->xz && a->u && a->conv_w && a->conv_b && a->seq_of_job && a->pset_of_job && a->rev_of_job,
              "cad_conv_silu_fwd: null pointer");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0, "cad_conv_silu_fwd: bad sizes");
  if (a->L == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dim3 grid((unsigned)((a->L + 8 * 256 - 1) / (8 * 256)), (unsigned)a->E, (unsigned)a->njobs);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, conv_silu_fwd_kernel<T><<<grid, 256, 0, stream>>>(*a));
  CAD_LAUNCH_CHECK();
  return 0;
}

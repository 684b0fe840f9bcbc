// Depthwise (anti)causal conv (K <= 4) + SiLU, materialised per job: the operand of the x_proj GEMM.
//
// Replaces causal_conv1d_cuda.causal_conv1d_fwd (upstream, reached via ref:caduceus/modeling_caduceus.py:128-133)
// and — for reversed jobs — the `hidden_states.flip(dims=(1,))` of ref:caduceus/modeling_caduceus.py:131: the
// reverse Mamba's causal conv over the flipped sequence is an ANTI-causal conv in original coordinates
// (SURVEY.md A.5), so no flipped copy is made.
//
//   u[tau] = silu(b + sum_{k<4} w[k] * x[tau - 3 + k]),  tau logical time; physical t = tau or L-1-tau.
//
// HBM-bound: each thread produces one 16-byte vector of outputs from one aligned 16-byte load plus the
// neighbouring vector (an L1/L2 hit); row pitches are multiples of 16 elements so the vectors stay in-row.
#include "common.cuh"

namespace cad {

template <typename T>
__global__ void __launch_bounds__(256) conv_silu_fwd_kernel(cad_conv_fwd_args a) {
  constexpr int V = 16 / sizeof(T);            // outputs per thread (8 for 16-bit, 4 for fp32)
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

  const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (t0 >= L) return;
  // window of V + 3 physical samples: fwd needs x[t0-3 .. t0+V-1]; rev needs x[t0 .. t0+V+2]
  float win[V + 3];
  auto sample = [&](int64_t t) -> float {      // out-of-sequence sample: shard halo or zero
    if (halo) {
      const int64_t tau = rev ? (L - 1 - t) : t;
      if (tau >= -3 && tau < 0) return io<T>::to_f(halo[tau + 3]);
    }
    return 0.f;
  };
  {
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + t0));
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float v = (t0 + i < L) ? io<T>::to_f(e[i]) : sample(t0 + i);
      win[(rev ? 0 : 3) + i] = v;
    }
    if (!rev) {
      if (t0 >= V) {
        uint4 pr = __ldg(reinterpret_cast<const uint4*>(x + t0 - V));
        const T* p = reinterpret_cast<const T*>(&pr);
#pragma unroll
        for (int i = 0; i < 3; ++i) win[i] = io<T>::to_f(p[V - 3 + i]);
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) win[i] = sample(t0 - 3 + i);
      }
    } else {
      if (t0 + V < a.ldxz) {
        uint4 nx = __ldg(reinterpret_cast<const uint4*>(x + t0 + V));
        const T* p = reinterpret_cast<const T*>(&nx);
#pragma unroll
        for (int i = 0; i < 3; ++i) win[V + i] = (t0 + V + i < L) ? io<T>::to_f(p[i]) : sample(t0 + V + i);
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) win[V + i] = sample(t0 + V + i);
      }
    }
  }
  uint4 outv;
  T* o = reinterpret_cast<T*>(&outv);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float acc;
    if (!rev) acc = bias + w0 * win[i] + w1 * win[i + 1] + w2 * win[i + 2] + w3 * win[i + 3];
    else      acc = bias + w3 * win[i] + w2 * win[i + 1] + w1 * win[i + 2] + w0 * win[i + 3];
    o[i] = io<T>::from_f(silu_io<T>(acc));
  }
  // the vector may run past L inside the (16-element padded) row pitch: harmless, the pad is never consumed
  *reinterpret_cast<uint4*>(u + t0) = outv;
}

}  // namespace cad

extern "C" int cad_conv_silu_fwd(const cad_conv_fwd_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_conv_silu_fwd: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0, "cad_conv_silu_fwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->u && a->conv_w && a->conv_b && a->seq_of_job && a->pset_of_job && a->rev_of_job,
              "cad_conv_silu_fwd: null pointer");
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldu % 16 == 0 && a->ldxz >= a->L && a->ldu >= a->L,
              "cad_conv_silu_fwd: row pitches must be multiples of 16 elements and >= L");
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->u), "cad_conv_silu_fwd: xz/u must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t v = 16 / (int64_t)dtype_size(a->io_dtype);
  dim3 grid((unsigned)((a->L + v * 256 - 1) / (v * 256)), (unsigned)a->E, (unsigned)a->njobs);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, conv_silu_fwd_kernel<T><<<grid, 256, 0, stream>>>(*a));
  CAD_LAUNCH_CHECK();
  return 0;
}

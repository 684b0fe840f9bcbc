// Fused (RC-equivariant) LM head + masked cross-entropy, forward and backward.
//
// Reference: logits = lm_head(hidden).float() over ALL positions (ref:caduceus/modeling_caduceus.py:474-476; the RCPS head is
// x1 W^T + flip_C(x2) W[cmap]^T, ref:caduceus/modeling_rcps.py:233-246), then cross_entropy(..., ignore_index)
// (ref:caduceus/modeling_caduceus.py:478-482, :279-294 for the weighted form) — of which only the ~15 % masked positions contribute.
// Here the fp32 (B, L, V) logits never exist: one warp per token row reads the label first; ignored rows cost one 8-byte load
// (forward) or one zero row (backward).  For a kept row the V <= 32 logits are V dot products of the row with the table —
// the RC half through the index map  Wcat[v, D + c] = W[cmap[v], D - 1 - c]  read in place — reduced across the warp, then
// log-sum-exp and the (weighted) NLL.  Per-block partial sums, summed by the caller: deterministic, no atomics on the loss.
//   forward :  loss_partial[b] = sum over the block's kept rows of  w_r * (lse_r - logit_r[y_r]);  wsum_partial[b] = sum w_r
//              (w_r = 1 without loss weights);  lse (rows) saved for the backward
//   backward:  g_r = dloss * w_r / wsum;  dlogit_r[v] = g_r * (exp(logit_r[v] - lse_r) - [v == y_r])
//              dhidden[r, :] = sum_v dlogit_r[v] * Wcat[v, :]   (zero rows for ignored tokens)
//              dWcat_partial[b, v, :] = sum over the block's kept rows of dlogit_r[v] * hidden[r, :]   (caller folds the RC half back)
#include "common.cuh"

namespace cad {

constexpr int kMaxV = 32;

template <typename T>
__device__ __forceinline__ float wcat_at(const T* __restrict__ W, const int64_t* __restrict__ cmap, int64_t D, int rcps, int v,
                                         int64_t c) {
  if (!rcps || c < D) return io<T>::to_f(W[(int64_t)v * D + c]);
  return io<T>::to_f(W[cmap[v] * D + (2 * D - 1 - c)]);           // column D + c' of Wcat = column D - 1 - c' of row cmap[v]
}

// the block's copy of Wcat (V, width) in shared memory, io dtype
template <typename T>
__device__ __forceinline__ void stage_wcat(T* s_w, const T* __restrict__ W, const int64_t* __restrict__ cmap, int64_t D, int64_t width,
                                           int rcps, int V) {
  for (int64_t i = threadIdx.x; i < (int64_t)V * width; i += blockDim.x) {
    const int v = (int)(i / width);
    s_w[i] = io<T>::from_f(wcat_at<T>(W, cmap, D, rcps, v, i - (int64_t)v * width));
  }
  __syncthreads();
}

// logits of one row, replicated in every lane: lg[v] for v < V
template <typename T>
__device__ __forceinline__ void row_logits(const T* __restrict__ h, const T* s_w, int64_t width, int V, int lane, float (&lg)[kMaxV]) {
#pragma unroll
  for (int v = 0; v < kMaxV; ++v) lg[v] = 0.f;
  for (int64_t c = lane; c < width; c += 32) {
    const float x = io<T>::to_f(h[c]);
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) lg[v] = fmaf(x, io<T>::to_f(s_w[(int64_t)v * width + c]), lg[v]);
  }
#pragma unroll
  for (int v = 0; v < kMaxV; ++v)
    if (v < V) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lg[v] += __shfl_xor_sync(0xffffffffu, lg[v], o);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) head_ce_fwd_kernel(cad_head_ce_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_wc = reinterpret_cast<T*>(smem_raw);        // (V, width) Wcat
  __shared__ float s_loss[8], s_w[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * wpb + wib, nwarps = (int64_t)gridDim.x * wpb;
  const T* __restrict__ H = static_cast<const T*>(a.hidden);
  const int V = (int)a.V;
  stage_wcat<T>(s_wc, static_cast<const T*>(a.weight), a.cmap, a.D, a.width, a.rcps, V);
  float loss = 0.f, wsum = 0.f;
  for (int64_t r = warp; r < a.rows; r += nwarps) {
    const int64_t y = a.labels[r];
    if (y == a.ignore_index) { if (lane == 0 && a.lse) a.lse[r] = 0.f; continue; }
    float lg[kMaxV];
    row_logits<T>(H + r * a.ldh, s_wc, a.width, V, lane, lg);
    float m = -INFINITY;
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) m = fmaxf(m, lg[v]);
    float s = 0.f, ly = 0.f;
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) { s += __expf(lg[v] - m); if (v == (int)y) ly = lg[v]; }
    const float lse = m + __logf(s);
    const float w = a.loss_weights ? a.loss_weights[r] : 1.f;
    loss += w * (lse - ly);
    wsum += w;
    if (lane == 0 && a.lse) a.lse[r] = lse;
  }
  if (lane == 0) { s_loss[wib] = loss; s_w[wib] = wsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, w = 0.f;
    for (int i = 0; i < wpb; ++i) { l += s_loss[i]; w += s_w[i]; }      // fixed order: deterministic
    a.loss_partial[blockIdx.x] = l;
    a.wsum_partial[blockIdx.x] = w;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) head_ce_bwd_kernel(cad_head_ce_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int V = (int)a.V;
  const int64_t width = a.width;
  float* s_dw = reinterpret_cast<float*>(smem_raw);                       // (V, width): this block's dWcat
  T* s_wc = reinterpret_cast<T*>(s_dw + (int64_t)V * width);              // (V, width) Wcat
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * wpb + wib, nwarps = (int64_t)gridDim.x * wpb;
  const T* __restrict__ H = static_cast<const T*>(a.hidden);
  T* __restrict__ dH = static_cast<T*>(a.dhidden);
  for (int64_t i = threadIdx.x; i < (int64_t)V * width; i += blockDim.x) s_dw[i] = 0.f;
  stage_wcat<T>(s_wc, static_cast<const T*>(a.weight), a.cmap, a.D, width, a.rcps, V);
  const float gscale = a.dloss_scale[0];          // dloss / wsum, computed by the caller on the device (no host sync)
  for (int64_t r = warp; r < a.rows; r += nwarps) {
    const int64_t y = a.labels[r];
    T* drow = dH + r * a.lddh;
    if (y == a.ignore_index) {
      for (int64_t c = lane; c < width; c += 32) drow[c] = io<T>::from_f(0.f);
      continue;
    }
    const T* hrow = H + r * a.ldh;
    float lg[kMaxV];
    row_logits<T>(hrow, s_wc, width, V, lane, lg);
    const float lse = a.lse[r];
    const float g = gscale * (a.loss_weights ? a.loss_weights[r] : 1.f);
#pragma unroll
    for (int v = 0; v < kMaxV; ++v)
      if (v < V) lg[v] = g * (__expf(lg[v] - lse) - (v == (int)y ? 1.f : 0.f));        // dlogit
    for (int64_t c = lane; c < width; c += 32) {
      const float x = io<T>::to_f(hrow[c]);
      float d = 0.f;
#pragma unroll
      for (int v = 0; v < kMaxV; ++v)
        if (v < V) {
          d = fmaf(lg[v], io<T>::to_f(s_wc[(int64_t)v * width + c]), d);
          atomicAdd(&s_dw[(int64_t)v * width + c], lg[v] * x);      // shared-memory atomics: warps of the block share the tile
        }
      drow[c] = io<T>::from_f(d);
    }
  }
  __syncthreads();
  float* out = a.dwcat_partial + (int64_t)blockIdx.x * V * width;
  for (int64_t i = threadIdx.x; i < (int64_t)V * width; i += blockDim.x) out[i] = s_dw[i];
}

static int head_ce_check(const cad_head_ce_args* a, const char* who) {
  CAD_REQUIRE(a, "%s: null argument block", who);
  CAD_REQUIRE(a->rows >= 0 && a->D > 0 && a->V > 0 && a->V <= kMaxV, "%s: bad sizes (V <= %d)", who, kMaxV);
  CAD_REQUIRE(a->width == (a->rcps ? 2 * a->D : a->D), "%s: width must be D (Ph) or 2 D (RCPS)", who);
  CAD_REQUIRE(a->hidden && a->weight && a->labels && (!a->rcps || a->cmap), "%s: null pointer", who);
  CAD_REQUIRE(a->ldh >= a->width && a->nblocks >= 1, "%s: bad pitch / nblocks", who);
  return 0;
}

}  // namespace cad

extern "C" int cad_head_ce_blocks(int64_t rows) {
  int64_t b = (rows + 7) / 8;
  const int64_t cap = (int64_t)(cad_sm_count() > 0 ? cad_sm_count() : 148) * 2;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int cad_head_ce_fwd(const cad_head_ce_args* a, void* stream_) {
  using namespace cad;
  if (head_ce_check(a, "cad_head_ce_fwd")) return -1;
  CAD_REQUIRE(a->loss_partial && a->wsum_partial, "cad_head_ce_fwd: null output pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = (size_t)a->V * a->width * dtype_size(a->io_dtype);
  CAD_REQUIRE(smem <= 48 * 1024, "cad_head_ce_fwd: V * width = %lld elements do not fit one shared-memory tile", (long long)(a->V * a->width));
  CAD_DISPATCH_DTYPE(a->io_dtype, T, (head_ce_fwd_kernel<T><<<a->nblocks, 256, smem, stream>>>(*a)));
  CAD_LAUNCH_CHECK();
  return 0;
}

extern "C" int cad_head_ce_bwd(const cad_head_ce_args* a, void* stream_) {
  using namespace cad;
  if (head_ce_check(a, "cad_head_ce_bwd")) return -1;
  CAD_REQUIRE(a->lse && a->dloss_scale && a->dhidden && a->dwcat_partial && a->lddh >= a->width, "cad_head_ce_bwd: null pointer / pitch");
  const size_t smem = (size_t)a->V * a->width * (sizeof(float) + dtype_size(a->io_dtype));
  CAD_REQUIRE(smem <= 200 * 1024, "cad_head_ce_bwd: V * width = %lld floats do not fit one shared-memory tile", (long long)(a->V * a->width));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, {
    auto kern = head_ce_bwd_kernel<T>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(head_ce_bwd): %s", cudaGetErrorString(e)); return (int)e; }
    }
    kern<<<a->nblocks, 256, smem, stream>>>(*a);
  });
  CAD_LAUNCH_CHECK();
  return 0;
}

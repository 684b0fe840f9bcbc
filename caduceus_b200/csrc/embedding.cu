// Token embedding for Caduceus-Ph and the RC-equivariant Caduceus-PS (integer index work, bit-exact).
//
// Reference semantics: ref:caduceus/modeling_caduceus.py:159-163 (Ph) and
// ref:caduceus/modeling_rcps.py:46-67 (PS):  cat[emb(ids), flip_{L,C}(emb(cmap[flip_L(ids)]))].
// In original coordinates the RC half is  out[b,l,D+c] = W[cmap[ids[b,l]], D-1-c]  (SURVEY.md A.7): the two
// length flips cancel, so no reversed copy of the sequence is ever materialised.
#include "common.cuh"

namespace cad {

// One warp per token row; lanes sweep the D (or 2D) output channels with 128-bit stores (VEC), or element-wise when D is not a
// multiple of the vector width (d_model = 118 of the reference's 1k-token models) or a base pointer is not 16-byte aligned.
template <typename T, bool VECOK>
__global__ void __launch_bounds__(256) embedding_fwd_kernel(
    const int64_t* __restrict__ ids, const T* __restrict__ W, const int64_t* __restrict__ cmap,
    T* __restrict__ out, int64_t rows, int64_t V, int64_t D, int rcps) {
  constexpr int VEC = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t width = rcps ? 2 * D : D;
  for (int64_t r = warp; r < rows; r += nwarps) {
    int64_t id = ids[r];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);                 // clamp like a defensive gather
    const T* src = W + id * D;
    T* dst = out + r * width;
    if constexpr (!VECOK) {
      for (int64_t c = lane; c < D; c += 32) dst[c] = src[c];
      if (rcps) {
        const T* srcc = W + cmap[id] * D;
        for (int64_t c = lane; c < D; c += 32) dst[D + c] = srcc[D - 1 - c];
      }
    } else {
    for (int64_t c = (int64_t)lane * VEC; c < D; c += 32 * VEC) {
      uint4 raw = *reinterpret_cast<const uint4*>(src + c);
      *reinterpret_cast<uint4*>(dst + c) = raw;
    }
    if (rcps) {
      const int64_t idc = cmap[id];
      const T* srcc = W + idc * D;
      for (int64_t c = (int64_t)lane * VEC; c < D; c += 32 * VEC) {
        // output channels D+c .. D+c+VEC-1 take table columns D-1-c .. D-c-VEC (descending)
        uint4 raw = *reinterpret_cast<const uint4*>(srcc + (D - c - VEC));
        const T* e = reinterpret_cast<const T*>(&raw);
        uint4 rev;
        T* o = reinterpret_cast<T*>(&rev);
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = e[VEC - 1 - k];
        *reinterpret_cast<uint4*>(dst + D + c) = rev;
      }
    }
    }
  }
}

// Backward: dW[v, col] += dout rows with ids == v (and the RC half through cmap: output channel D + c feeds table column D - 1 - c).
// Each block accumulates a (V, DT) fp32 tile in shared memory over a slab of rows for ONE slab of DT table columns (grid.y), then
// adds it to global with one atomic per (v, col) per block.  Thread t owns columns {t, t + blockDim, ...} of the tile for every
// row — both halves are indexed by the table column they land on — so no two threads ever touch the same address, whatever
// cmap does (self-complementary tokens included).  V * DT * 4 <= 48 KB for any d_model (DT shrinks as V grows).
template <typename T>
__global__ void __launch_bounds__(256) embedding_bwd_kernel(
    const int64_t* __restrict__ ids, const int64_t* __restrict__ cmap, const T* __restrict__ dout,
    float* __restrict__ dW, int64_t rows, int64_t V, int64_t D, int rcps, int64_t rows_per_block, int64_t DT) {
  extern __shared__ float acc[];   // (V, DT)
  const int64_t d0 = (int64_t)blockIdx.y * DT, dn = min(DT, D - d0);
  for (int64_t i = threadIdx.x; i < V * DT; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  const int64_t width = rcps ? 2 * D : D;
  for (int64_t r = r0; r < r1; ++r) {
    int64_t id = ids[r];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const T* g = dout + r * width;
    for (int64_t c = threadIdx.x; c < dn; c += blockDim.x) acc[id * DT + c] += io<T>::to_f(g[d0 + c]);
    if (rcps) {
      const int64_t idc = cmap[id];
      for (int64_t c = threadIdx.x; c < dn; c += blockDim.x) acc[idc * DT + c] += io<T>::to_f(g[2 * D - 1 - (d0 + c)]);
    }
  }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < V * DT; i += blockDim.x) {
    const int64_t v = i / DT, c = i - v * DT;
    if (c < dn && acc[i] != 0.f) atomicAdd(dW + v * D + d0 + c, acc[i]);
  }
}

}  // namespace cad

extern "C" int cad_embedding_fwd(const cad_embedding_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_embedding_fwd: null argument block");
  CAD_REQUIRE(a->B >= 0 && a->L >= 0 && a->V > 0 && a->D > 0, "cad_embedding_fwd: bad sizes");
  if (a->B * a->L == 0) return 0;           // empty batch: nothing to enqueue (pointers may be null)
  CAD_REQUIRE(a->ids && a->weight && a->out, "cad_embedding_fwd: null pointer");
  CAD_REQUIRE(!a->rcps || a->cmap, "cad_embedding_fwd: rcps needs a complement map");
  const int64_t vec = 16 / (int64_t)dtype_size(a->dtype);
  const bool vecok = a->D % vec == 0 && aligned16(a->weight) && aligned16(a->out);
  const int64_t rows = a->B * a->L;
  if (rows == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int threads = 256, wpb = threads / 32;
  int64_t blocks = (rows + wpb - 1) / wpb;
  const int64_t cap = (int64_t)cad_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  CAD_DISPATCH_DTYPE(a->dtype, T,
    if (vecok) embedding_fwd_kernel<T, true><<<(unsigned)blocks, threads, 0, stream>>>(
        a->ids, static_cast<const T*>(a->weight), a->cmap, static_cast<T*>(a->out), rows, a->V, a->D, a->rcps);
    else embedding_fwd_kernel<T, false><<<(unsigned)blocks, threads, 0, stream>>>(
        a->ids, static_cast<const T*>(a->weight), a->cmap, static_cast<T*>(a->out), rows, a->V, a->D, a->rcps));
  CAD_LAUNCH_CHECK();
  return 0;
}

extern "C" int cad_embedding_bwd(const cad_embedding_bwd_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_embedding_bwd: null argument block");
  if (a->B * a->L == 0) return 0;
  CAD_REQUIRE(a->ids && a->dout && a->dweight, "cad_embedding_bwd: null pointer");
  CAD_REQUIRE(!a->rcps || a->cmap, "cad_embedding_bwd: rcps needs a complement map");
  CAD_REQUIRE(a->V > 0 && a->D > 0 && a->V <= 12288, "cad_embedding_bwd: bad sizes (V <= 12288)");
  const int64_t rows = a->B * a->L;
  if (rows == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // table columns per block: the (V, DT) fp32 tile stays within the default 48 KB of dynamic shared memory
  int64_t DT = (48 * 1024 / 4) / a->V;
  if (DT > a->D) DT = a->D;
  if (DT > 512) DT = 512;
  const int64_t nslab = (a->D + DT - 1) / DT;
  int64_t blocks = (int64_t)cad_sm_count() * 4 / nslab;
  if (blocks < 1) blocks = 1;
  const int64_t rpb = (rows + blocks - 1) / blocks;
  const int64_t nblk = (rows + rpb - 1) / rpb;
  dim3 grid((unsigned)nblk, (unsigned)nslab);
  CAD_DISPATCH_DTYPE(a->dtype, T,
    embedding_bwd_kernel<T><<<grid, 256, (size_t)(a->V * DT * 4), stream>>>(
        a->ids, a->cmap, static_cast<const T*>(a->dout), a->dweight, rows, a->V, a->D, a->rcps, rpb, DT));
  CAD_LAUNCH_CHECK();
  return 0;
}

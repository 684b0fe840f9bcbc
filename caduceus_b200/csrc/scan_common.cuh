// Shared pieces of the fused scan kernels (forward and backward): chunk geometry, mbarrier/TMA primitives,
// the host-side tensor map over the B/C rows.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cad {

constexpr int kTok = 16;            // tokens per lane
constexpr int kChunk = 32 * kTok;   // 512 logical tokens per chunk
constexpr int kMaxG = 7;            // channels (warps) per CTA
constexpr int kBlkTok = 32;         // tokens per 128-byte swizzle line

// ---- mbarrier / TMA primitives ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same wait with a watchdog, for kernels whose hand-over protocol has not yet run on hardware: a barrier that
// does not complete within 2 s (a whole launch takes milliseconds) traps — the launch fails loudly instead of hanging the GPU
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  uint64_t t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try(bar, parity)) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2),
      "r"(smem_u32(bar)) : "memory");
}


// explicit shared-space accesses with 32-bit addresses (generic LD/ST through 64-bit pointers cost address
// arithmetic and the slower generic path: profiles/r1_v2_scan_ncu_summary.txt shows 0.63 generic LD per element)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// one step of the inclusive (decay, state) warp scan: lanes >= OFF absorb the aggregate of lane - OFF
template <int OFF>
__device__ __forceinline__ void scan_step_up(float& P, float& H, int lane) {
  const float Pp = __shfl_up_sync(0xffffffffu, P, OFF);
  const float Hp = __shfl_up_sync(0xffffffffu, H, OFF);
  asm("{\n.reg .pred q;\nsetp.ge.s32 q, %4, %5;\n@q fma.rn.f32 %0, %1, %2, %0;\n@q mul.f32 %1, %1, %3;\n}"
      : "+f"(H), "+f"(P) : "f"(Hp), "f"(Pp), "r"(lane), "n"(OFF));
}

// mirrored step for the adjoint (suffix) scan: lanes < 32 - OFF absorb the aggregate of lane + OFF
template <int OFF>
__device__ __forceinline__ void scan_step_down(float& Q, float& E, int lane) {
  const float Qn = __shfl_down_sync(0xffffffffu, Q, OFF);
  const float En = __shfl_down_sync(0xffffffffu, E, OFF);
  asm("{\n.reg .pred q;\nsetp.lt.s32 q, %4, %5;\n@q fma.rn.f32 %0, %1, %2, %0;\n@q mul.f32 %1, %1, %3;\n}"
      : "+f"(E), "+f"(Q) : "f"(En), "f"(Qn), "r"(lane), "n"(32 - OFF));
}

// this lane's TOK/4 16-byte pieces inside a TMA-swizzled tile row (see scan_fwd.cu): SWIZZLE_128B stores 16-byte
// chunk c of 128-byte line l at chunk position c ^ (l & 7); a tile row is CH/32 consecutive lines (line = 32 tokens).
template <int TOK = kTok>
__device__ __forceinline__ void tile_piece_offsets(int seg, uint32_t (&poff)[TOK / 4]) {
  constexpr int SPB = 32 / TOK;                 // lane segments per 32-token line
  const int blk = seg / SPB, c0 = (TOK / 4) * (seg % SPB);
#pragma unroll
  for (int k = 0; k < TOK / 4; ++k) poff[k] = blk * 128 + (((c0 + k) ^ (blk & 7)) << 4);
}

// ---- host: tensor map over a (nrows, ld) fp32 matrix viewed as (32 tokens, blocks, rows) ----------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rows_per_box rows x chunk_tokens tokens per TMA; blocks past ceil(L/32) are out of bounds -> zero-filled.
inline int make_row_tile_map(CUtensorMap* tmap, const float* base, int64_t nrows, int64_t ld, int64_t L,
                             int rows_per_box, int chunk_tokens = kChunk) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return -1; }
  const cuuint64_t nblk = (cuuint64_t)((L + kBlkTok - 1) / kBlkTok);
  const cuuint64_t dims[3] = {(cuuint64_t)kBlkTok, nblk, (cuuint64_t)nrows};
  const cuuint64_t strides[2] = {(cuuint64_t)kBlkTok * 4, (cuuint64_t)ld * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kBlkTok, (cuuint32_t)(chunk_tokens / kBlkTok), (cuuint32_t)rows_per_box};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
  return 0;
}


// same for a 16-bit (nrows, ld) matrix viewed as (64 tokens, blocks, rows): a 128-byte swizzle line holds 64 tokens
inline int make_row_tile_map16(CUtensorMap* tmap, const void* base, int is_bf16, int64_t nrows, int64_t ld, int64_t L,
                               int rows_per_box, int chunk_tokens) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return -1; }
  constexpr int kLine = 64;
  const cuuint64_t nblk = (cuuint64_t)((L + kLine - 1) / kLine);
  const cuuint64_t dims[3] = {(cuuint64_t)kLine, nblk, (cuuint64_t)nrows};
  const cuuint64_t strides[2] = {(cuuint64_t)kLine * 2, (cuuint64_t)ld * 2};
  const cuuint32_t box[3] = {(cuuint32_t)kLine, (cuuint32_t)(chunk_tokens / kLine), (cuuint32_t)rows_per_box};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tmap, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (16-bit) failed (%d)", (int)r); return -1; }
  return 0;
}

}  // namespace cad

// Fused  (anti)causal depthwise conv + SiLU  ->  x_proj (tensor cores)  ->  dt_proj (tensor cores)  for sm_100a.
//
// Replaces, per job, the chain  causal_conv1d_fwd -> F.linear(x_proj) -> dt_proj.weight @ x_dbl[:R]  of upstream's
// `mamba_inner_fn` (reached from ref:caduceus/modeling_caduceus.py:128-133): `u = silu(conv(x))` never
// touches HBM, x is read once, and the kernel writes exactly what the fused scan consumes:
//     delta (njobs, E, ldd)   = W_dt . x_dbl[0:R]            io dtype (same rounding point as the reference pipeline)
//     bc    (njobs, 2N, ldbc) = x_dbl[R:R+2N]                fp32, zero beyond the sequence end (TMA tile source)
//     bcT   (njobs, ldT, 2N)     optionally, the same values token-major (what the lane = channel scan, variant 20, reads)
// Tensor cores are used for the two dense projections only (north_star): mma.sync m16n8k16 with fp32 accumulation;
// both GEMMs are skinny (M = R+2N = 48 resp. K = R = 16) and the kernel is HBM-bound on the delta write, so
// the legacy warp-level MMA path is already far above what the memory system needs (profiles/).
//
// CTA = (job, 128-token tile), 8 warps.  K loop over 64-channel slabs:  x slab (+8-token aprons) -> smem ->
// conv+SiLU -> bf16 u slab [channel][token] -> ldmatrix.trans B fragments; A = W_x slab.  Warp w owns tokens
// [16w, 16w+16) of the tile for the x_proj accumulators (3 m-tiles x 2 n-tiles).  Then x_dbl[0:16] (bf16) becomes
// the B operand of the dt_proj GEMM (M = E, K = 16), whose 16-row output slabs are staged through smem for
// coalesced 16-byte stores.
#include "common.cuh"

namespace cad {

constexpr int XT = 128;           // tokens per CTA tile
constexpr int KC = 64;            // channels per K slab
constexpr int XP = 152;           // pitch of the raw x slab (144 used: 8-token aprons either side)
constexpr int UP = 136;           // pitch of the u slab / dt slab / staging rows (128 + 8: conflict-free ldmatrix)
constexpr int WP = 72;            // pitch of the W_x slab (64 + 8)
constexpr int DP = 24;            // pitch of the padded W_dt rows (16 + 8)
constexpr int MROWS = 48;         // R + 2N rows handled (3 m-tiles)

template <typename T> struct mma_t;
template <> struct mma_t<__nv_bfloat16> {
  __device__ static __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
};
template <> struct mma_t<__half> {
  __device__ static __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
};
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  T v[2] = {io<T>::from_f(lo), io<T>::from_f(hi)};
  return *reinterpret_cast<uint32_t*>(v);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(256, 2) conv_xproj_kernel(cad_conv_xproj_args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  T* xs = reinterpret_cast<T*>(smem);                  // [2][KC][XP]  raw x slabs (double-buffered, cp.async)
  T* wxs = xs + 2 * KC * XP;                           // [2][MROWS][WP]
  T* us = wxs + 2 * MROWS * WP;                        // [KC][UP]
  T* wdts = us + KC * UP;                              // [E][DP]   (E <= 1024 guarded on the host)
  T* dts = wdts + (size_t)a.E * DP;                    // [16][UP]
  float* cws = reinterpret_cast<float*>(dts + 16 * UP); // [2][KC][8]: conv taps (4) + bias of the slab's channels
  T* stg = xs;                                         // [8 warps][16][UP]: aliases the x slabs after the K loop

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int job = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * XT;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  const int64_t L = a.L, E = a.E;
  const int R = (int)a.R, N = (int)a.N;
  const T* __restrict__ xbase = static_cast<const T*>(a.xz) + (int64_t)seq * 2 * E * a.ldxz;
  const T* __restrict__ wx = static_cast<const T*>(a.w_x) + (int64_t)pset * (R + 2 * N) * E;
  const T* __restrict__ wdt = static_cast<const T*>(a.w_dt) + (int64_t)pset * E * R;
  const T* halo = a.halo ? static_cast<const T*>(a.halo) + (int64_t)job * E * 3 : nullptr;
  const T zero = io<T>::from_f(0.f);
  const bool interior = (t0 >= 8) && (t0 + XT + 8 <= L);     // whole apron inside the sequence

  // stage one K slab: x rows [c0, c0+64) x tokens [t0-8, t0+136) and the W_x columns [c0, c0+64)
  auto stage = [&](int64_t c0, int buf) {
    T* xb = xs + buf * KC * XP;
    T* wb = wxs + buf * MROWS * WP;
    for (int i = tid; i < KC * 18; i += 256) {
      const int ch = i / 18, v = i - ch * 18;
      const int64_t t = t0 - 8 + 8 * v;
      T* dst = xb + ch * XP + 8 * v;
      const T* row = xbase + (c0 + ch) * a.ldxz;
      if (interior || (t >= 0 && t + 8 <= L)) {
        cp_async16(dst, row + t);
      } else {                                         // sequence ends: element-wise, shard halo or zero outside
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int64_t te = t + e;
          T val = zero;
          if (te >= 0 && te < L) val = row[te];
          else if (halo) {
            const int64_t tau = rev ? (L - 1 - te) : te;
            if (tau >= -3 && tau < 0) val = halo[(c0 + ch) * 3 + tau + 3];
          }
          dst[e] = val;
        }
      }
    }
    for (int i = tid; i < MROWS * (KC / 8); i += 256) {
      const int r = i / (KC / 8), v = i - r * (KC / 8);
      T* dst = wb + r * WP + 8 * v;
      if (r < R + 2 * N) cp_async16(dst, wx + (int64_t)r * E + c0 + 8 * v);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    if (tid < KC) {                                    // conv taps + bias of the slab's 64 channels
      const int64_t pc = (int64_t)pset * E + c0 + tid;
      float* cd = cws + (buf * KC + tid) * 8;
      cp_async16(cd, a.conv_w + pc * 4);
      cd[4] = a.conv_b[pc];
    }
    cp_async_commit();
  };

  stage(0, 0);

  // ---- W_dt, zero-padded to 16 columns (A operand of the dt_proj GEMM), once per CTA -----------------------
  if (R == 16) {
    for (int i = tid; i < (int)E * 2; i += 256) {
      const int ch = i >> 1, v = i & 1;
      *reinterpret_cast<uint4*>(wdts + ch * DP + 8 * v) = __ldg(reinterpret_cast<const uint4*>(wdt + (int64_t)ch * 16 + 8 * v));
    }
  } else {
    for (int i = tid; i < (int)E * 16; i += 256) {
      const int ch = i >> 4, r = i & 15;
      wdts[ch * DP + r] = r < R ? wdt[(int64_t)ch * R + r] : zero;
    }
  }

  float acc[3][2][4];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;

  const int nslab = (int)(E / KC);
  for (int sl = 0; sl < nslab; ++sl) {
    const int buf = sl & 1;
    cp_async_wait<0>();
    __syncthreads();                                   // slab `sl` visible; everyone is past the MMA of slab sl-1
    // the slab after this one streams in (into the buffers slab sl-1 used) while this one is convolved + multiplied
    if (sl + 1 < nslab) stage((int64_t)(sl + 1) * KC, buf ^ 1);
    const T* xb = xs + buf * KC * XP;
    const T* wb = wxs + buf * MROWS * WP;
    const int64_t c0 = (int64_t)sl * KC;
    // ---- conv + SiLU: thread (chb, v) handles the 8-token vector v of channels chb, chb+16, chb+32, chb+48 -------
    {
      const int v = tid & 15, chb = tid >> 4;
      const float* cwb = cws + buf * KC * 8;
#pragma unroll
      for (int k = 0; k < KC / 16; ++k) {
        const int ch = chb + 16 * k;
        const float4 cw = *reinterpret_cast<const float4*>(cwb + ch * 8);
        const float cb = cwb[ch * 8 + 4];
        // three aligned vectors around my 8 tokens: x[t-8 .. t+15], t = t0 + 8v
        const uint4* xv = reinterpret_cast<const uint4*>(xb + ch * XP + 8 * v);
        const uint4 r0 = xv[0], r1 = xv[1], r2 = xv[2];
        const T* e0 = reinterpret_cast<const T*>(&r0);
        const T* e1 = reinterpret_cast<const T*>(&r1);
        const T* e2 = reinterpret_cast<const T*>(&r2);
        float win[14];                                   // x[t-3 .. t+10]
#pragma unroll
        for (int e = 0; e < 3; ++e) win[e] = io<T>::to_f(e0[5 + e]);
#pragma unroll
        for (int e = 0; e < 8; ++e) win[3 + e] = io<T>::to_f(e1[e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) win[11 + e] = io<T>::to_f(e2[e]);
        uint4 outv;
        uint32_t* o = reinterpret_cast<uint32_t*>(&outv);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float c0v, c1v;
          if (!rev) {
            c0v = cb + cw.x * win[e] + cw.y * win[e + 1] + cw.z * win[e + 2] + cw.w * win[e + 3];
            c1v = cb + cw.x * win[e + 1] + cw.y * win[e + 2] + cw.z * win[e + 3] + cw.w * win[e + 4];
          } else {
            c0v = cb + cw.w * win[e + 3] + cw.z * win[e + 4] + cw.y * win[e + 5] + cw.x * win[e + 6];
            c1v = cb + cw.w * win[e + 4] + cw.z * win[e + 5] + cw.y * win[e + 6] + cw.x * win[e + 7];
          }
          o[e >> 1] = pack2<T>(silu_io<T>(c0v), silu_io<T>(c1v));
        }
        *reinterpret_cast<uint4*>(us + ch * UP + 8 * v) = outv;
      }
    }
    __syncthreads();
    // ---- x_proj MMA: acc[m][j] += W_x[16m.., slab] . u[slab, 16 warp + 8j ..] -----------------------------------
#pragma unroll
    for (int ks = 0; ks < KC / 16; ++ks) {
      uint32_t bfr[4];
      ldsm_x4_trans(bfr, us + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * UP + 16 * warp + 8 * (lane >> 4));
      const uint32_t b0[2] = {bfr[0], bfr[1]}, b1[2] = {bfr[2], bfr[3]};
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        uint32_t afr[4];
        ldsm_x4(afr, wb + (16 * m + (lane & 15)) * WP + ks * 16 + 8 * (lane >> 4));
        mma_t<T>::mma(acc[m][0], afr, b0);
        mma_t<T>::mma(acc[m][1], afr, b1);
      }
    }
  }

  // ---- epilogue 1: B/C rows -> fp32 global (zero beyond L), dt rows -> bf16 smem --------------------------------
  const int g = lane >> 2, q2 = 2 * (lane & 3);
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int r = 16 * m + g + 8 * hrow;
        const int col = 16 * warp + 8 * j + q2;
        const float v0 = acc[m][j][2 * hrow], v1 = acc[m][j][2 * hrow + 1];
        if (m == 0) *reinterpret_cast<uint32_t*>(dts + r * UP + col) = pack2<T>(v0, v1);
        if (r >= R && r < R + 2 * N) {
          const int64_t t = t0 + col;
          if (t < a.ldbc) {
            float* dst = a.bc + ((int64_t)job * 2 * N + (r - R)) * a.ldbc + t;
            if (t + 1 < a.ldbc) *reinterpret_cast<float2*>(dst) = make_float2(t < L ? v0 : 0.f, t + 1 < L ? v1 : 0.f);
            else dst[0] = t < L ? v0 : 0.f;
          }
          if (a.bcT) {                         // the same values token-major: what the lane = channel scan (variants 20..23) reads
            float* dT = a.bcT + ((int64_t)job * a.ldT + t) * (2 * N) + (r - R);
            if (t < a.ldT) dT[0] = t < L ? v0 : 0.f;
            if (t + 1 < a.ldT) dT[2 * N] = t + 1 < L ? v1 : 0.f;
          }
        }
      }
  __syncthreads();

  // ---- epilogue 2: delta = W_dt(pad16) . x_dbl[0:16]  -> staged 16-row slabs -> coalesced stores -------------------
  uint32_t bdt[16][2];
#pragma unroll
  for (int jp = 0; jp < 8; ++jp) {
    uint32_t r4[4];
    ldsm_x4_trans(r4, dts + ((lane & 7) + 8 * ((lane >> 3) & 1)) * UP + 16 * jp + 8 * (lane >> 4));
    bdt[2 * jp][0] = r4[0]; bdt[2 * jp][1] = r4[1]; bdt[2 * jp + 1][0] = r4[2]; bdt[2 * jp + 1][1] = r4[3];
  }
  T* mystg = stg + (size_t)warp * 16 * UP;
  T* __restrict__ dbase = static_cast<T*>(a.delta) + (int64_t)job * E * a.ldd;
  for (int64_t mt = warp; mt < E / 16; mt += 8) {
    uint32_t afr[4];
    ldsm_x4(afr, wdts + (16 * mt + (lane & 15)) * DP + 8 * (lane >> 4));
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma_t<T>::mma(c, afr, bdt[j]);
      *reinterpret_cast<uint32_t*>(mystg + g * UP + 8 * j + q2) = pack2<T>(c[0], c[1]);
      *reinterpret_cast<uint32_t*>(mystg + (g + 8) * UP + 8 * j + q2) = pack2<T>(c[2], c[3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = i * 32 + lane;                 // 256 vectors = 16 rows x 16 vectors
      const int r = idx >> 4, v = idx & 15;
      const int64_t t = t0 + 8 * v;
      if (t < a.ldd)
        *reinterpret_cast<uint4*>(dbase + (16 * mt + r) * a.ldd + t) = *reinterpret_cast<const uint4*>(mystg + r * UP + 8 * v);
    }
    __syncwarp();
  }
}

}  // namespace cad

extern "C" int cad_conv_xproj_fwd(const cad_conv_xproj_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_conv_xproj_fwd: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0, "cad_conv_xproj_fwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->w_x && a->w_dt && a->conv_w && a->conv_b && a->seq_of_job && a->pset_of_job &&
              a->rev_of_job && a->delta && a->bc, "cad_conv_xproj_fwd: null pointer");
  CAD_REQUIRE(a->io_dtype == CAD_BF16 || a->io_dtype == CAD_F16,
              "cad_conv_xproj_fwd: tensor-core path needs 16-bit I/O (fp32 uses the unfused path)");
  CAD_REQUIRE(a->N == 16 && a->R >= 1 && a->R <= 16, "cad_conv_xproj_fwd: needs d_state = 16 and dt_rank <= 16");
  CAD_REQUIRE(a->E % 64 == 0 && a->E <= 1024, "cad_conv_xproj_fwd: d_inner must be a multiple of 64, <= 1024");
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldd % 16 == 0 && a->ldxz >= a->L && a->ldd >= a->L && a->ldbc >= a->L,
              "cad_conv_xproj_fwd: bad row pitches");
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->delta) && aligned16(a->w_x) && ((uintptr_t)a->bc & 7) == 0 &&
              a->ldbc % 2 == 0, "cad_conv_xproj_fwd: alignment");
  CAD_REQUIRE(!a->bcT || (aligned16(a->bcT) && a->ldT >= a->L), "cad_conv_xproj_fwd: bcT must be 16-byte aligned with ldT >= L");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static_assert(2 * KC * XP >= 8 * 16 * UP, "output staging must fit in the x slabs it aliases");
  const size_t smem = sizeof(uint16_t) * ((size_t)2 * KC * XP + 2 * MROWS * WP + KC * UP + (size_t)a->E * DP + 16 * UP) +
                      sizeof(float) * 2 * KC * 8;
  dim3 grid((unsigned)((a->L + XT - 1) / XT), (unsigned)a->njobs);
  cudaError_t e;
  if (a->io_dtype == CAD_BF16) {
    e = cudaFuncSetAttribute(conv_xproj_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(*a);
  } else {
    e = cudaFuncSetAttribute(conv_xproj_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_kernel<__half><<<grid, 256, smem, stream>>>(*a);
  }
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  CAD_LAUNCH_CHECK();
  return 0;
}

// Fused  (anti)causal depthwise conv + SiLU  ->  x_proj  ->  dt_proj  on the 5th-generation tensor cores (tcgen05.mma,
// accumulators in tensor memory) for sm_100a.  Same contract as conv_xproj_kernel (xproj.cu): replaces, per job, the chain
// causal_conv1d_fwd -> F.linear(x_proj) -> dt_proj.weight @ x_dbl[:R]  of upstream's `mamba_inner_fn` (SURVEY.md A.1,
// reached from ref:caduceus/modeling_caduceus.py:128-133) and writes what the scan consumes (delta, bc, optionally bcT).
//
// Both projections are issued by ONE thread per CTA; the other 255 only convolve and move data:
//   x_proj :  D1[128 tokens x 48]   += u^T[128 x 32] . W_x[48 x 32]^T      per 32-channel slab, 2 x (M128 N48 K16)
//             A = the u slab exactly as the conv threads produce it (channel-major rows of 8 tokens = an MN-major
//             operand of 8x8 core matrices, no swizzle), B = the W_x slab (K-major); D1 in TMEM columns [0, 48).
//   dt_proj:  D2[128 tokens x 128 channels] = bf16(x_dbl[:, 0:16]) . W_dt[128 x 16]^T   per 128-channel chunk, one M128 N128 K16
//             A = the dt rows of D1 read back (tcgen05.ld), rounded to the io dtype (the reference's rounding point) and
//             stored K-major; B = W_dt resident in shared memory; D2 double-buffered in TMEM columns [0,128) / [128,256).
// TMEM lane = token, so every warp-level store of a delta column is 32 consecutive tokens of one channel row.
//
// CTA = 256 threads, persistent over the 128-token tiles of one job (grid.x CTAs per job, 2 CTAs per SM: 256 of the 512
// TMEM columns each).  The K loop never drains between tiles: x slabs (+ 8-token aprons), the W_x slab and the conv taps
// stream through cp.async rings three slabs ahead; ONE __syncthreads per slab (it publishes the u slab to the tensor
// core and the next x slab to the conv threads); a u buffer is rewritten only after the mbarrier its MMAs committed to.
#include "common.cuh"
#include "scan_common.cuh"

namespace cad {
namespace umma {

constexpr int XT = 128;            // tokens per tile = UMMA M
constexpr int KC = 32;             // channels per K slab
constexpr int XP = 152;            // pitch of a raw x row (144 used: 8-token aprons either side); 304 B = 19 x 16 B (odd)
constexpr int NSX = 4;             // x / conv-tap ring slots (prefetch distance 3)
constexpr int NSW = 5;             // W_x ring slots: one more, the tensor core reads a slot one iteration longer
constexpr int XPROJ_N = 48;        // dt rows (padded to 16) + B rows + C rows
constexpr int DTN = 128;           // channels per dt_proj instruction
constexpr int TMEM_COLS = 256;

constexpr int XS_BYTES = KC * XP * 2;              // 9728
constexpr int CW_BYTES = KC * 8 * 4;               // 1024: 4 taps + bias (+3 pad) per channel
constexpr int WX_BYTES = XPROJ_N * KC * 2;         // 3072
constexpr int U_BYTES = KC * XT * 2;               // 8192
constexpr int DT_BYTES = XT * 16 * 2;              // 4096

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t k_stride_bytes, uint32_t mn_stride_bytes, int swap) {
  // no-swizzle canonical layout: 8 x 16-byte core matrices; "leading" offset = next core matrix along K, "stride" offset =
  // next along M/N (scripts/umma_probe.cu checks this reading on the hardware)
  const uint32_t lbo = swap ? mn_stride_bytes : k_stride_bytes, sbo = swap ? k_stride_bytes : mn_stride_bytes;
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <typename T> struct umma_fmt;
template <> struct umma_fmt<__nv_bfloat16> { static constexpr uint32_t v = 1; };
template <> struct umma_fmt<__half> { static constexpr uint32_t v = 0; };
// D = f32; A, B formats; A major (1 = MN); B K-major; N >> 3 at bit 17; M >> 4 at bit 24
template <typename T>
__device__ __forceinline__ uint32_t instr_desc(int M, int N, int a_mn_major) {
  return (1u << 4) | (umma_fmt<T>::v << 7) | (umma_fmt<T>::v << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {      // arrives on `bar` when every MMA issued so far has completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  T v[2] = {io<T>::from_f(lo), io<T>::from_f(hi)};
  return *reinterpret_cast<uint32_t*>(v);
}

template <typename T>
__global__ void __launch_bounds__(256, 2) conv_xproj_umma_kernel(cad_conv_xproj_args a, int desc_swap) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* xs = smem;                                   // [NSX][KC][XP] T      raw x slabs
  unsigned char* wxs = xs + NSX * XS_BYTES;                   // [NSW] W_x slab as a K-major UMMA operand: [k group 4][n group 6][8][16 B]
  unsigned char* us = wxs + NSW * WX_BYTES;                   // [2]   u slab as an MN-major UMMA operand: [k group 4][token group 16][8 ch][16 B]
  unsigned char* dts = us + 2 * U_BYTES;                      // dt rows, K-major operand: [k group 2][token group 16][8 tok][16 B]
  unsigned char* cws = dts + DT_BYTES;                        // [NSX][KC][8] float
  unsigned char* wdts = cws + NSX * CW_BYTES;                 // W_dt, K-major operand: [k group 2][channel group E/8][8][16 B]
  __shared__ uint64_t ubar[2], accbar, dbar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int job = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  const int64_t L = a.L, E = a.E;
  const int R = (int)a.R, N = (int)a.N;
  const T* __restrict__ xbase = static_cast<const T*>(a.xz) + (int64_t)seq * 2 * E * a.ldxz;
  const T* __restrict__ wx = static_cast<const T*>(a.w_x) + (int64_t)pset * (R + 2 * N) * E;
  const T* __restrict__ wdt = static_cast<const T*>(a.w_dt) + (int64_t)pset * E * R;
  const T* halo = a.halo ? static_cast<const T*>(a.halo) + (int64_t)job * E * 3 : nullptr;
  const T zero = io<T>::from_f(0.f);
  const int nslab = (int)(E / KC);
  const int64_t ntiles = (L + XT - 1) / XT;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total = my_tiles * nslab;                     // slabs this CTA walks, numbered g = 0 .. total-1 across its tiles

  // ---- one K slab into the rings: x rows [c0, c0+32) x tokens [t0-8, t0+136), W_x columns [c0, c0+32), conv taps --------
  auto stage = [&](int64_t g) {
    if (g < total) {
      const int64_t it = g / nslab;
      const int64_t c0 = (g - it * nslab) * KC;
      const int64_t t0 = (blockIdx.x + it * gridDim.x) * XT;
      const bool interior = (t0 >= 8) && (t0 + XT + 8 <= L);
      T* xb = reinterpret_cast<T*>(xs + (g % NSX) * XS_BYTES);
      for (int i = tid; i < KC * 18; i += 256) {
        const int ch = i / 18, v = i - ch * 18;
        const int64_t t = t0 - 8 + 8 * v;
        T* dst = xb + ch * XP + 8 * v;
        const T* row = xbase + (c0 + ch) * a.ldxz;
        if (interior || (t >= 0 && t + 8 <= L)) {
          cp_async16(dst, row + t);
        } else {                                             // sequence ends: element-wise, shard halo or zero outside
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
      if (tid < XPROJ_N * (KC / 8)) {                        // 192 16-byte pieces: (operand row n, k group)
        const int n = tid >> 2, kg = tid & 3;
        const int src = n < 16 ? (n < R ? n : -1) : n - 16 + R;      // dt rows padded to 16: rows [R, 16) stay zero
        if (src >= 0)
          cp_async16(wxs + (g % NSW) * WX_BYTES + kg * (XPROJ_N / 8 * 128) + (n >> 3) * 128 + (n & 7) * 16,
                     wx + (int64_t)src * E + c0 + 8 * kg);
      } else if (tid >= 224) {                               // conv taps + bias of the slab's 32 channels
        const int ch = tid - 224;
        const int64_t pc = (int64_t)pset * E + c0 + ch;
        float* cd = reinterpret_cast<float*>(cws + (g % NSX) * CW_BYTES) + ch * 8;
        cp_async16(cd, a.conv_w + pc * 4);
        cd[4] = a.conv_b[pc];
      }
    }
    cp_async_commit();
  };

  // ---- set-up: barriers, TMEM, zero rows of the W_x ring, resident W_dt --------------------------------------------------
  if (tid == 0) {
    mbar_init(&ubar[0], 1); mbar_init(&ubar[1], 1); mbar_init(&accbar, 1); mbar_init(&dbar[0], 1); mbar_init(&dbar[1], 1);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (R < 16)
    for (int i = tid; i < NSW * WX_BYTES / 16; i += 256) reinterpret_cast<uint4*>(wxs)[i] = make_uint4(0, 0, 0, 0);
  {
    const uint32_t kstride = (uint32_t)(E / 8) * 128;
    if (R == 16) {
      for (int i = tid; i < (int)E * 2; i += 256) {
        const int ch = i >> 1, kg = i & 1;
        *reinterpret_cast<uint4*>(wdts + kg * kstride + (ch >> 3) * 128 + (ch & 7) * 16) =
            __ldg(reinterpret_cast<const uint4*>(wdt + (int64_t)ch * 16 + 8 * kg));
      }
    } else {
      for (int i = tid; i < (int)E * 16; i += 256) {
        const int ch = i >> 4, r = i & 15;
        *reinterpret_cast<T*>(wdts + (r >> 3) * kstride + (ch >> 3) * 128 + (ch & 7) * 16 + (r & 7) * 2) =
            r < R ? wdt[(int64_t)ch * R + r] : zero;
      }
    }
  }
  __syncthreads();                                           // zero rows written before any cp.async lands in the ring
  stage(0); stage(1); stage(2);
  cp_async_wait<2>();
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc_x = instr_desc<T>(XT, XPROJ_N, 1);
  const int q = warp & 3, half = warp >> 2;                  // TMEM lane quarter this warp may read; column half it takes
  const uint32_t tlane = (uint32_t)(32 * q) << 16;
  uint32_t nuse_d0 = 0, nuse_d1 = 0;                          // completed-phase counters of dbar[0], dbar[1]

  for (int64_t g = 0; g < total; ++g) {
    const int64_t it = g / nslab;
    const int sl = (int)(g - it * nslab);
    const int buf = (int)(g & 1);
    const int64_t t0 = (blockIdx.x + it * gridDim.x) * XT;
    // u[buf] and W_x slot (g-2) % NSW are free once the MMAs of slab g-2 have completed
    if (g >= 2) mbar_wait_wd(&ubar[buf], (uint32_t)(((g >> 1) - 1) & 1));
    stage(g + 3);
    // ---- conv + SiLU: this thread's two (channel, 8-token vector) pieces of the slab -----------------------------------
    {
      const T* xb = reinterpret_cast<const T*>(xs + (g % NSX) * XS_BYTES);
      const float* cwb = reinterpret_cast<const float*>(cws + (g % NSX) * CW_BYTES);
      unsigned char* ub = us + buf * U_BYTES;
      const int tg = 4 * q + (lane >> 3);                      // token group 0..15
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kg = 2 * half + i;                           // channel group 0..3
        const int ch = 8 * kg + (lane & 7);
        const float4 cw = *reinterpret_cast<const float4*>(cwb + ch * 8);
        const float cb = cwb[ch * 8 + 4];
        const uint4* xv = reinterpret_cast<const uint4*>(xb + ch * XP + 8 * tg);   // x[t-8 .. t+15], t = t0 + 8 tg
        const uint4 r0 = xv[0], r1 = xv[1], r2 = xv[2];
        const T* e0 = reinterpret_cast<const T*>(&r0);
        const T* e1 = reinterpret_cast<const T*>(&r1);
        const T* e2 = reinterpret_cast<const T*>(&r2);
        float win[14];                                         // x[t-3 .. t+10]
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
        // core matrix (k group kg, token group tg), row = channel within the group: a quarter-warp writes 128 contiguous bytes
        *reinterpret_cast<uint4*>(ub + kg * (XT / 8 * 128) + tg * 128 + (lane & 7) * 16) = outv;
      }
    }
    cp_async_wait<2>();                                        // slab g+1 has landed (g+2, g+3 may still be in flight)
    proxy_fence();                                             // my u stores / W_x cp.async data -> visible to the tensor core
    tc_fence_before();                                         // (and my TMEM reads of the previous tile's epilogue are done)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ua = smem_u32(us + buf * U_BYTES), wa = smem_u32(wxs + (g % NSW) * WX_BYTES);
#pragma unroll
      for (int ks = 0; ks < KC / 16; ++ks)
        mma_f16(tmem, smem_desc(ua + ks * 2 * (XT / 8 * 128), XT / 8 * 128, 128, desc_swap),
                smem_desc(wa + ks * 2 * (XPROJ_N / 8 * 128), XPROJ_N / 8 * 128, 128, desc_swap), idesc_x, (sl | ks) ? 1u : 0u);
      mma_commit(&ubar[buf]);
      if (sl == nslab - 1) mma_commit(&accbar);
    }
    if (sl != nslab - 1) continue;

    // ================= tile epilogue ======================================================================================
    mbar_wait_wd(&accbar, (uint32_t)(it & 1));
    tc_fence_after();
    const int64_t t = t0 + 32 * q + lane;                       // my token (TMEM lane)
    {
      // warps 0-3: dt columns [0,16) -> io dtype -> dt operand; B columns [16,32) -> bc rows [0,16).  warps 4-7: C columns.
      uint32_t v[16];
      if (half == 0) {
        tmem_ld16(tmem + tlane + 0, v);
        tmem_ld_wait();
        uint4 lo, hi;
        lo.x = pack2<T>(__uint_as_float(v[0]), __uint_as_float(v[1]));   lo.y = pack2<T>(__uint_as_float(v[2]), __uint_as_float(v[3]));
        lo.z = pack2<T>(__uint_as_float(v[4]), __uint_as_float(v[5]));   lo.w = pack2<T>(__uint_as_float(v[6]), __uint_as_float(v[7]));
        hi.x = pack2<T>(__uint_as_float(v[8]), __uint_as_float(v[9]));   hi.y = pack2<T>(__uint_as_float(v[10]), __uint_as_float(v[11]));
        hi.z = pack2<T>(__uint_as_float(v[12]), __uint_as_float(v[13])); hi.w = pack2<T>(__uint_as_float(v[14]), __uint_as_float(v[15]));
        const int row = 32 * q + lane;
        *reinterpret_cast<uint4*>(dts + (row >> 3) * 128 + (row & 7) * 16) = lo;
        *reinterpret_cast<uint4*>(dts + (XT / 8 * 128) + (row >> 3) * 128 + (row & 7) * 16) = hi;
      }
      tmem_ld16(tmem + tlane + 16 + 16 * half, v);
      tmem_ld_wait();
      const bool live = t < L;
#pragma unroll
      for (int j = 0; j < 16; ++j) if (!live) v[j] = 0u;
      if (t < a.ldbc) {
        float* dst = a.bc + ((int64_t)job * 2 * N + 16 * half) * a.ldbc + t;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[(int64_t)j * a.ldbc] = __uint_as_float(v[j]);
      }
      if (a.bcT && t < a.ldT) {
        uint4* dT = reinterpret_cast<uint4*>(a.bcT + ((int64_t)job * a.ldT + t) * (2 * N) + 16 * half);
#pragma unroll
        for (int j = 0; j < 4; ++j) dT[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();                                           // dt operand complete; x_proj accumulator columns free
    const int nchunk = (int)((E + DTN - 1) / DTN);
    const uint32_t wdt_k = (uint32_t)(E / 8) * 128;
    auto issue_dt = [&](int c) {                               // chunk c -> TMEM buffer c & 1
      const int nc = (int)min((int64_t)DTN, E - (int64_t)c * DTN);
      tc_fence_after();
      mma_f16(tmem + (uint32_t)((c & 1) * DTN), smem_desc(smem_u32(dts), XT / 8 * 128, 128, desc_swap),
              smem_desc(smem_u32(wdts) + (uint32_t)(c * DTN / 8) * 128, wdt_k, 128, desc_swap), instr_desc<T>(XT, nc, 0), 0u);
      mma_commit(&dbar[c & 1]);
    };
    if (tid == 0) { issue_dt(0); if (nchunk > 1) issue_dt(1); }
    T* __restrict__ dbase = static_cast<T*>(a.delta) + (int64_t)job * E * a.ldd + t;
    for (int c = 0; c < nchunk; ++c) {
      const int nc = (int)min((int64_t)DTN, E - (int64_t)c * DTN);
      if (c & 1) { mbar_wait_wd(&dbar[1], nuse_d1 & 1); ++nuse_d1; } else { mbar_wait_wd(&dbar[0], nuse_d0 & 1); ++nuse_d0; }
      tc_fence_after();
      // my half of the chunk's columns, 32 at a time: register j = channel, lane = token
      for (int cc = half * (nc / 2); cc < (half + 1) * (nc / 2); cc += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + tlane + (uint32_t)((c & 1) * DTN + cc), v);
        tmem_ld_wait();
        if (t < a.ldd) {
          T* dst = dbase + ((int64_t)c * DTN + cc) * a.ldd;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(int64_t)j * a.ldd] = io<T>::from_f(__uint_as_float(v[j]));
        }
      }
      if (c + 2 < nchunk) {                                    // hand the buffer back for chunk c + 2
        tc_fence_before();
        __syncthreads();
        if (tid == 0) issue_dt(c + 2);
      }
    }
    // the next tile's first MMA is issued after the next __syncthreads (tc_fence_before precedes it): TMEM reads are ordered
  }

  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace umma
}  // namespace cad

extern "C" int cad_conv_xproj_umma_fwd(const cad_conv_xproj_args* a, void* stream_) {
  using namespace cad;
  using namespace cad::umma;
  CAD_REQUIRE(a, "cad_conv_xproj_umma_fwd: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0, "cad_conv_xproj_umma_fwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->w_x && a->w_dt && a->conv_w && a->conv_b && a->seq_of_job && a->pset_of_job &&
              a->rev_of_job && a->delta && a->bc, "cad_conv_xproj_umma_fwd: null pointer");
  CAD_REQUIRE(a->io_dtype == CAD_BF16 || a->io_dtype == CAD_F16,
              "cad_conv_xproj_umma_fwd: tensor-core path needs 16-bit I/O (fp32 uses the unfused path)");
  CAD_REQUIRE(a->N == 16 && a->R >= 1 && a->R <= 16, "cad_conv_xproj_umma_fwd: needs d_state = 16 and dt_rank <= 16");
  CAD_REQUIRE(a->E % 64 == 0 && a->E <= 2048, "cad_conv_xproj_umma_fwd: d_inner must be a multiple of 64, <= 2048");
  CAD_REQUIRE(a->ldxz % 8 == 0 && a->ldxz >= a->L && a->ldd >= a->L && a->ldbc >= a->L,
              "cad_conv_xproj_umma_fwd: bad row pitches");
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->w_x) && (a->R != 16 || aligned16(a->w_dt)), "cad_conv_xproj_umma_fwd: alignment");
  CAD_REQUIRE(!a->bcT || (aligned16(a->bcT) && a->ldT >= a->L), "cad_conv_xproj_umma_fwd: bcT must be 16-byte aligned with ldT >= L");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = (size_t)NSX * XS_BYTES + NSW * WX_BYTES + 2 * U_BYTES + DT_BYTES + NSX * CW_BYTES + (size_t)a->E * 32;
  const int64_t ntiles = (a->L + XT - 1) / XT;
  const int sms = cad_sm_count();
  CAD_REQUIRE(sms > 0, "cad_conv_xproj_umma_fwd: no CUDA device");
  // persistent CTAs: two per SM in total, shared evenly by the jobs (every job has the same number of tiles)
  int64_t per_job = (2 * (int64_t)sms + a->njobs - 1) / a->njobs;
  if (per_job > ntiles) per_job = ntiles;
  if (per_job < 1) per_job = 1;
  static const int swap = [] { const char* s = getenv("CAD_UMMA_DESC_SWAP"); return s && s[0] == '1' ? 1 : 0; }();
  dim3 grid((unsigned)per_job, (unsigned)a->njobs);
  cudaError_t e;
  if (a->io_dtype == CAD_BF16) {
    e = cudaFuncSetAttribute(conv_xproj_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_umma_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(*a, swap);
  } else {
    e = cudaFuncSetAttribute(conv_xproj_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_umma_kernel<__half><<<grid, 256, smem, stream>>>(*a, swap);
  }
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  CAD_LAUNCH_CHECK();
  return 0;
}

// Fused  (anti)causal depthwise conv + SiLU  ->  x_proj  ->  dt_proj  on the 5th-generation tensor cores (tcgen05.mma,
// accumulators in tensor memory) for sm_100a.  Same contract as conv_xproj_kernel (xproj.cu): replaces, per job, the chain
// causal_conv1d_fwd -> F.linear(x_proj) -> dt_proj.weight @ x_dbl[:R]  of upstream's `mamba_inner_fn` (SURVEY.md A.1,
// reached from ref:caduceus/modeling_caduceus.py:128-133) and writes what the scan consumes (delta, bc, optionally bcT).
//
// Both projections are issued by ONE thread per CTA; the other 255 only convolve and move data:
//   x_proj :  D1[128 tokens x 48]   += u^T[128 x 32] . W_x[48 x 32]^T      per 32-channel slab, 2 x (M128 N48 K16)
//             A = the u slab exactly as the conv threads produce it (channel-major rows of 8 tokens = an MN-major
//             operand of 8x8 core matrices, no swizzle), B = the W_x slab (K-major); D1 in TMEM columns [0, 48),
//             TMEM lane = token: the B / C rows leave as coalesced fp32 rows, the dt rows go back to shared memory.
//   dt_proj:  D2[128 channels x 128 tokens] = W_dt[chunk, 0:16] . bf16(x_dbl[:, 0:16])^T   per 128-channel chunk, one M128 N128 K16
//             A = W_dt resident in shared memory (K-major), B = the dt rows of D1 rounded to the io dtype (the reference's
//             rounding point), K-major; D2 double-buffered in TMEM columns [0,128) / [128,256).  TMEM lane = channel: a
//             thread reads 32 consecutive tokens of one delta row and stores them as four 16-byte vectors.
//
// CTA = 256 threads, persistent over the 128-token tiles of one job (grid.x CTAs per job, 2 CTAs per SM: 256 of the 512
// TMEM columns each); W_dt and the conv taps are loaded once per CTA.  The K loop never drains between tiles: each x slab
// (32 channels x 152 tokens with the conv aprons) is ONE bulk tensor copy (TMA, zero fill outside the sequence) into a
// 4-slot ring three slabs ahead, signalled on an mbarrier; the W_x slab follows through cp.async.  ONE __syncthreads per
// slab publishes the u slab to the tensor core; a u buffer is rewritten only after the mbarrier its MMAs committed to.
// First hardware run, descriptor probe and the profile that shaped this version: profiles/r2_call13_umma_first_hw_run.log.
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
constexpr int WX_BYTES = XPROJ_N * KC * 2;         // 3072
constexpr int U_BYTES = KC * XT * 2;               // 8192
constexpr int DT_BYTES = XT * 16 * 2;              // 4096

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t k_stride_bytes, uint32_t mn_stride_bytes) {
  // no-swizzle canonical layout: 8 x 16-byte core matrices; "leading" byte offset = next core matrix along K, "stride" byte
  // offset = next along M/N, for K-major and MN-major operands alike (scripts/umma_probe.cu checks this reading on the
  // hardware: profiles/r2_call13_umma_first_hw_run.log); descriptor version 1 at bit 46
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((k_stride_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((mn_stride_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
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

// 16-bit pair -> two floats (bf16: one shift / one mask; fp16: one packed convert)
template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t w, float& lo, float& hi) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  lo = f.x; hi = f.y;
}

template <typename T>
__global__ void __launch_bounds__(256, 2) conv_xproj_umma_kernel(cad_conv_xproj_args a, const __grid_constant__ CUtensorMap xmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int64_t E = a.E;
  const int E128 = (int)((E + 127) / 128 * 128);
  unsigned char* xs = smem;                                   // [NSX][KC][XP] T      raw x slabs (TMA boxes)
  unsigned char* wxs = xs + NSX * XS_BYTES;                   // [NSW] W_x slab as a K-major UMMA operand: [k group 4][n group 6][8][16 B]
  unsigned char* us = wxs + NSW * WX_BYTES;                   // [2]   u slab as an MN-major UMMA operand: [k group 4][token group 16][8 ch][16 B]
  unsigned char* dts = us + 2 * U_BYTES;                      // dt rows, K-major operand: [k group 2][token group 16][8 tok][16 B]
  unsigned char* wdts = dts + DT_BYTES;                       // W_dt, K-major operand: [k group 2][channel group E128/8][8][16 B], zero rows past E
  float4* taps = reinterpret_cast<float4*>(wdts + (size_t)E128 * 32);   // [E] conv taps, reversed for anti-causal jobs (x 1/2 for 16-bit I/O: silu(v) = h + h tanh(h), h = v/2)
  float* cbias = reinterpret_cast<float*>(taps + E);          // [E] conv bias (same scaling)
  __shared__ uint64_t xbar[NSX], ubar[2], accbar, dbar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int job = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  const int64_t L = a.L;
  const int R = (int)a.R, N = (int)a.N;
  const T* __restrict__ wx = static_cast<const T*>(a.w_x) + (int64_t)pset * (R + 2 * N) * E;
  const T* __restrict__ wdt = static_cast<const T*>(a.w_dt) + (int64_t)pset * E * R;
  const T* halo = a.halo ? static_cast<const T*>(a.halo) + (int64_t)job * E * 3 : nullptr;
  const T zero = io<T>::from_f(0.f);
  const int nslab = (int)(E / KC);
  const int64_t ntiles = (L + XT - 1) / XT;
  const int my_tiles = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int total = my_tiles * nslab;                         // slabs this CTA walks, numbered g = 0 .. total-1 across its tiles
  const int64_t tile_step = (int64_t)gridDim.x * XT;

  // ---- one K slab into the rings: x box [c0, c0+32) x [t0-8, t0+144) by ONE bulk tensor copy (zero fill outside [0, L)),
  //      W_x columns [c0, c0+32) as 192 16-byte pieces ---------------------------------------------------------------------
  int pf_g = 0, pf_sl = 0;                                    // next slab to request, its index inside its tile
  int64_t pf_t0 = (int64_t)blockIdx.x * XT;                   // and the first token of that tile
  auto stage = [&]() {
    if (pf_g < total) {
      const int c0 = pf_sl * KC;
      if (tid == 0) {
        uint64_t* bar = &xbar[pf_g % NSX];
        mbar_expect_tx(bar, XS_BYTES);
        tma_load_3d(xs + (pf_g % NSX) * XS_BYTES, &xmap, (int)(pf_t0 - 8), c0, seq, bar);
      }
      if (tid < XPROJ_N * (KC / 8)) {                        // 192 16-byte pieces: (operand row n, k group)
        const int n = tid >> 2, kg = tid & 3;
        const int src = n < 16 ? (n < R ? n : -1) : n - 16 + R;      // dt rows padded to 16: rows [R, 16) stay zero
        if (src >= 0)
          cp_async16(wxs + (pf_g % NSW) * WX_BYTES + kg * (XPROJ_N / 8 * 128) + (n >> 3) * 128 + (n & 7) * 16,
                     wx + (int64_t)src * E + c0 + 8 * kg);
      }
      ++pf_g;
      if (++pf_sl == nslab) { pf_sl = 0; pf_t0 += tile_step; }
    }
    cp_async_commit();
  };

  // ---- set-up: barriers, TMEM, zero rows of the W_x ring, resident W_dt and conv taps ------------------------------------
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSX; ++i) mbar_init(&xbar[i], 1);
    mbar_init(&ubar[0], 1); mbar_init(&ubar[1], 1); mbar_init(&accbar, 1); mbar_init(&dbar[0], 1); mbar_init(&dbar[1], 1);
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (R < 16)
    for (int i = tid; i < NSW * WX_BYTES / 16; i += 256) reinterpret_cast<uint4*>(wxs)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t wdt_k = (uint32_t)(E128 / 8) * 128;            // K stride of the W_dt operand
  {
    if (R == 16) {
      for (int i = tid; i < E128 * 2; i += 256) {
        const int ch = i >> 1, kg = i & 1;
        *reinterpret_cast<uint4*>(wdts + kg * wdt_k + (ch >> 3) * 128 + (ch & 7) * 16) =
            ch < E ? __ldg(reinterpret_cast<const uint4*>(wdt + (int64_t)ch * 16 + 8 * kg)) : make_uint4(0, 0, 0, 0);
      }
    } else {
      for (int i = tid; i < E128 * 16; i += 256) {
        const int ch = i >> 4, r = i & 15;
        *reinterpret_cast<T*>(wdts + (r >> 3) * wdt_k + (ch >> 3) * 128 + (ch & 7) * 16 + (r & 7) * 2) =
            (r < R && ch < E) ? wdt[(int64_t)ch * R + r] : zero;
      }
    }
    const float sc = sizeof(T) == 2 ? 0.5f : 1.0f;
    for (int ch = tid; ch < (int)E; ch += 256) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(a.conv_w) + (int64_t)pset * E + ch);
      taps[ch] = rev ? make_float4(sc * w.w, sc * w.z, sc * w.y, sc * w.x) : make_float4(sc * w.x, sc * w.y, sc * w.z, sc * w.w);
      cbias[ch] = sc * a.conv_b[(int64_t)pset * E + ch];
    }
  }
  __syncthreads();                                           // barriers initialised, zero rows written before any copy lands
  stage(); stage(); stage();
  cp_async_wait<2>();
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc_x = instr_desc<T>(XT, XPROJ_N, 1);
  const uint32_t idesc_dt = instr_desc<T>(128, XT, 0);
  const int q = warp & 3, half = warp >> 2;                  // TMEM lane quarter this warp may read; column half it takes
  const uint32_t tlane = (uint32_t)(32 * q) << 16;
  uint32_t nuse_d0 = 0, nuse_d1 = 0;                          // completed-phase counters of dbar[0], dbar[1]
  int sl = 0, it = 0;
  int64_t t0 = (int64_t)blockIdx.x * XT;
  const int tg = 4 * q + (lane >> 3);                          // conv: my token group 0..15 and channel inside a group of 8
  const int cl = lane & 7;

  for (int g = 0; g < total; ++g) {
    const int buf = g & 1;
    // u[buf] and W_x slot (g-2) % NSW are free once the MMAs of slab g-2 have completed
    if (g >= 2) mbar_wait_wd(&ubar[buf], (uint32_t)(((g >> 1) - 1) & 1));
    stage();                                                   // slab g+3 -> x slot (g-1) % NSX, read for the last time before the previous barrier
    mbar_wait_wd(&xbar[g % NSX], (uint32_t)((g / NSX) & 1));
    unsigned char* xb = xs + (g % NSX) * XS_BYTES;
    if (halo && ((!rev && t0 == 0) || (rev && t0 + XT >= L))) {  // shard hook: the 3 samples that logically precede the shard
      if (tid < KC * 3) {
        const int ch = tid / 3, k = tid - ch * 3;
        const int64_t te = rev ? L + k : (int64_t)k - 3;        // physical position just outside the sequence
        const int64_t tau = rev ? L - 1 - te : te;              // logical time -3 .. -1
        reinterpret_cast<T*>(xb)[ch * XP + (int)(te - (t0 - 8))] = halo[((int64_t)sl * KC + ch) * 3 + tau + 3];
      }
      __syncthreads();
    }
    // ---- conv + SiLU: this thread's two (channel, 8-token vector) pieces of the slab -----------------------------------
    {
      unsigned char* ub = us + buf * U_BYTES;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kg = 2 * half + i;                           // channel group 0..3
        const int ch = 8 * kg + cl;
        const float4 cw = taps[sl * KC + ch];
        const float cb = cbias[sl * KC + ch];
        const uint4* xv = reinterpret_cast<const uint4*>(xb + (ch * XP + 8 * tg) * 2);   // x[t-8 .. t+15], t = t0 + 8 tg
        const uint4 r0 = xv[0], r1 = xv[1], r2 = xv[2];
        // 12-sample window so that output e reads win[e+1 .. e+4] in BOTH directions: causal jobs x[t-4 .. t+7] with taps
        // (w0..w3), anti-causal jobs x[t-1 .. t+10] with the taps stored reversed (one uniform branch per piece)
        float win[12];
        if (!rev) {
          unpack2<T>(r0.z, win[0], win[1]); unpack2<T>(r0.w, win[2], win[3]);
          unpack2<T>(r1.x, win[4], win[5]); unpack2<T>(r1.y, win[6], win[7]);
          unpack2<T>(r1.z, win[8], win[9]); unpack2<T>(r1.w, win[10], win[11]);
        } else {
          float skip;
          unpack2<T>(r0.w, skip, win[0]);
          unpack2<T>(r1.x, win[1], win[2]); unpack2<T>(r1.y, win[3], win[4]);
          unpack2<T>(r1.z, win[5], win[6]); unpack2<T>(r1.w, win[7], win[8]);
          unpack2<T>(r2.x, win[9], win[10]); unpack2<T>(r2.y, win[11], skip);
        }
        uint4 outv;
        uint32_t* o = reinterpret_cast<uint32_t*>(&outv);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {                       // outputs t+e, t+e+1
          float c0v = cb + cw.x * win[e + 1] + cw.y * win[e + 2] + cw.z * win[e + 3] + cw.w * win[e + 4];
          float c1v = cb + cw.x * win[e + 2] + cw.y * win[e + 3] + cw.z * win[e + 4] + cw.w * win[e + 5];
          if constexpr (sizeof(T) == 2) {                      // taps pre-scaled: c = v/2
            c0v = fmaf(c0v, tanh_approx(c0v), c0v);
            c1v = fmaf(c1v, tanh_approx(c1v), c1v);
          } else {
            c0v = silu(c0v); c1v = silu(c1v);
          }
          o[e >> 1] = pack2<T>(c0v, c1v);
        }
        // core matrix (k group kg, token group tg), row = channel within the group: a quarter-warp writes 128 contiguous bytes
        *reinterpret_cast<uint4*>(ub + kg * (XT / 8 * 128) + tg * 128 + cl * 16) = outv;
      }
    }
    cp_async_wait<2>();                                        // W_x of slab g+1 has landed (g+2, g+3 may still be in flight)
    proxy_fence();                                             // my u stores / W_x cp.async data -> visible to the tensor core
    tc_fence_before();                                         // (and my TMEM reads of the previous tile's epilogue are done)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ua = smem_u32(us + buf * U_BYTES), wa = smem_u32(wxs + (g % NSW) * WX_BYTES);
#pragma unroll
      for (int ks = 0; ks < KC / 16; ++ks)
        mma_f16(tmem, smem_desc(ua + ks * 2 * (XT / 8 * 128), XT / 8 * 128, 128),
                smem_desc(wa + ks * 2 * (XPROJ_N / 8 * 128), XPROJ_N / 8 * 128, 128), idesc_x, (sl | ks) ? 1u : 0u);
      mma_commit(&ubar[buf]);
      if (sl == nslab - 1) mma_commit(&accbar);
    }
    if (++sl != nslab) continue;

    // ================= tile epilogue ======================================================================================
    sl = 0;
    mbar_wait_wd(&accbar, (uint32_t)(it & 1));
    tc_fence_after();
    {
      // warps 0-3: dt columns [0,16) -> io dtype -> dt operand; B columns [16,32) -> bc rows [0,16).  warps 4-7: C columns.
      const int64_t t = t0 + 32 * q + lane;                     // my token (TMEM lane)
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
    // dt_proj, channels on the TMEM lanes: D2[128 channels x 128 tokens] = W_dt[chunk] . dt^T, so that a thread reads 32
    // consecutive tokens of ONE channel row and stores them as four 16-byte vectors
    const int nchunk = E128 / DTN;
    auto issue_dt = [&](int c) {                               // chunk c -> TMEM buffer c & 1
      tc_fence_after();
      mma_f16(tmem + (uint32_t)((c & 1) * XT), smem_desc(smem_u32(wdts) + (uint32_t)(c * DTN / 8) * 128, wdt_k, 128),
              smem_desc(smem_u32(dts), XT / 8 * 128, 128), idesc_dt, 0u);
      mma_commit(&dbar[c & 1]);
    };
    if (tid == 0) { issue_dt(0); if (nchunk > 1) issue_dt(1); }
    for (int c = 0; c < nchunk; ++c) {
      if (c & 1) { mbar_wait_wd(&dbar[1], nuse_d1 & 1); ++nuse_d1; } else { mbar_wait_wd(&dbar[0], nuse_d0 & 1); ++nuse_d0; }
      tc_fence_after();
      const int ch = c * DTN + 32 * q + lane;                   // my channel (TMEM lane)
      T* __restrict__ drow = static_cast<T*>(a.delta) + ((int64_t)job * E + ch) * a.ldd + t0 + 64 * half;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {                    // my 64 tokens, 32 at a time
        uint32_t v[32];
        tmem_ld32(tmem + tlane + (uint32_t)((c & 1) * XT + 64 * half + cc), v);
        tmem_ld_wait();
        if (ch < E) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t t = t0 + 64 * half + cc + 8 * j;
            if (t < a.ldd)
              *reinterpret_cast<uint4*>(drow + cc + 8 * j) =
                  make_uint4(pack2<T>(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                             pack2<T>(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                             pack2<T>(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                             pack2<T>(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
          }
        }
      }
      if (c + 2 < nchunk) {                                    // hand the buffer back for chunk c + 2
        tc_fence_before();
        __syncthreads();
        if (tid == 0) issue_dt(c + 2);
      }
    }
    // the next tile's first MMA is issued after the next __syncthreads (tc_fence_before precedes it): TMEM reads are ordered
    ++it;
    t0 += tile_step;
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
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_conv_xproj_umma_fwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->w_x && a->w_dt && a->conv_w && a->conv_b && a->seq_of_job && a->pset_of_job &&
              a->rev_of_job && a->delta && a->bc, "cad_conv_xproj_umma_fwd: null pointer");
  CAD_REQUIRE(a->io_dtype == CAD_BF16 || a->io_dtype == CAD_F16,
              "cad_conv_xproj_umma_fwd: tensor-core path needs 16-bit I/O (fp32 uses the unfused path)");
  CAD_REQUIRE(a->N == 16 && a->R >= 1 && a->R <= 16, "cad_conv_xproj_umma_fwd: needs d_state = 16 and dt_rank <= 16");
  CAD_REQUIRE(a->E % 64 == 0 && a->E <= 2048, "cad_conv_xproj_umma_fwd: d_inner must be a multiple of 64, <= 2048");
  CAD_REQUIRE(a->ldxz % 8 == 0 && a->ldd % 8 == 0 && a->ldxz >= a->L && a->ldd >= a->L && a->ldbc >= a->L,
              "cad_conv_xproj_umma_fwd: bad row pitches");
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->w_x) && aligned16(a->delta) && aligned16(a->conv_w) &&
              (a->R != 16 || aligned16(a->w_dt)), "cad_conv_xproj_umma_fwd: alignment");
  CAD_REQUIRE(!a->bcT || (aligned16(a->bcT) && a->ldT >= a->L), "cad_conv_xproj_umma_fwd: bcT must be 16-byte aligned with ldT >= L");
  CAD_REQUIRE(a->L < (int64_t)1 << 31, "cad_conv_xproj_umma_fwd: sequence too long for 32-bit tensor-map coordinates");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // x half of xz as a (L tokens, 2E rows, nseq) tensor: one box = 152 tokens x 32 channels; tokens outside [0, L) read as zero
  CUtensorMap xmap;
  {
    EncodeTiledFn enc = encode_tiled_fn();
    CAD_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[3] = {(cuuint64_t)a->L, (cuuint64_t)(2 * a->E), (cuuint64_t)a->nseq};
    const cuuint64_t strides[2] = {(cuuint64_t)a->ldxz * 2, (cuuint64_t)a->ldxz * 2 * 2 * (cuuint64_t)a->E};
    const cuuint32_t box[3] = {(cuuint32_t)XP, (cuuint32_t)KC, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&xmap, a->io_dtype == CAD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                     const_cast<void*>(a->xz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cad_conv_xproj_umma_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
  }
  const int64_t E128 = (a->E + 127) / 128 * 128;
  const size_t smem = (size_t)NSX * XS_BYTES + NSW * WX_BYTES + 2 * U_BYTES + DT_BYTES + (size_t)E128 * 32 + (size_t)a->E * 20;
  const int64_t ntiles = (a->L + XT - 1) / XT;
  const int sms = cad_sm_count();
  CAD_REQUIRE(sms > 0, "cad_conv_xproj_umma_fwd: no CUDA device");
  // persistent CTAs: two per SM in total, shared evenly by the jobs (every job has the same number of tiles)
  int64_t per_job = (2 * (int64_t)sms + a->njobs - 1) / a->njobs;
  if (per_job > ntiles) per_job = ntiles;
  if (per_job < 1) per_job = 1;
  dim3 grid((unsigned)per_job, (unsigned)a->njobs);
  cudaError_t e;
  if (a->io_dtype == CAD_BF16) {
    e = cudaFuncSetAttribute(conv_xproj_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_umma_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(*a, xmap);
  } else {
    e = cudaFuncSetAttribute(conv_xproj_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_umma_kernel<__half><<<grid, 256, smem, stream>>>(*a, xmap);
  }
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  CAD_LAUNCH_CHECK();
  return 0;
}

// Fused  (anti)causal depthwise conv + SiLU  ->  x_proj  ->  dt_proj  on the 5th-generation tensor cores (tcgen05.mma,
// accumulators in tensor memory) for sm_100a: replaces, per job, the chain  causal_conv1d_fwd -> F.linear(x_proj) ->
// dt_proj.weight @ x_dbl[:R]  of upstream's `mamba_inner_fn` (reached from
// ref:caduceus/modeling_caduceus.py:128-133) and writes what the scan consumes (delta, bc, optionally bcT); u = silu(conv(x))
// never touches HBM.  Contract: cad_conv_xproj_args in include/caduceus_b200.h.
//
// The two projections, per 128-token tile:
//   x_proj :  D1[128 tokens x 48]   += u^T[128 x 32] . W_x[48 x 32]^T      per 32-channel slab, 2 x (M128 N48 K16)
//             A = the u slab exactly as the conv threads produce it (channel-major rows of 8 tokens = an MN-major
//             operand of 8x8 core matrices, no swizzle), B = the W_x slab (K-major); D1 in TMEM columns [64 (tile & 1), +48),
//             TMEM lane = token: the B / C rows leave as coalesced fp32 rows, the dt rows go back to shared memory.
//   dt_proj:  D2[128 channels x 128 tokens] = W_dt[chunk, 0:16] . bf16(x_dbl[:, 0:16])^T   per 128-channel chunk, one M128 N128 K16
//             A = W_dt resident in shared memory (K-major), B = the dt rows of D1 rounded to the io dtype (the reference's
//             rounding point), K-major; D2 in TMEM columns [128, 256).  TMEM lane = channel: a thread reads 32 consecutive
//             tokens of one delta row; the warp stages (32 channels x 32 tokens) with the 64-byte swizzle and one lane stores
//             the block with a bulk tensor copy.
//
// CTA = 14 warps with fixed roles, persistent over the tiles of one job (grid.x CTAs per job, 2 CTAs per SM: 256 of the 512 TMEM
// columns each); W_dt and the conv taps are loaded once per CTA.  A 3-stage ring holds per K slab the raw x box (32 channels x
// 152 tokens with the conv aprons: ONE bulk tensor copy, zero fill outside the sequence), the W_x slab (one 3 KB bulk copy of the
// host-packed operand, or eight tensor-map boxes when the caller did not pack it) and the u slab.
//   warp 8      requests x slabs as soon as the conv warps hold a slot's samples in registers (x_empty -> x_full)
//   warps 0-7   convolve: x_full -> LDS -> x_empty;  slab_done (u slot free) -> FFMA / tanh -> STS -> proxy fence -> u_full
//   warp 9      w_full + u_full -> lane 0 issues the MMAs -> tcgen05.commit to slab_done (and acc_full after a tile's last slab);
//               requests the W_x slab two slabs ahead
//   warps 10-13 acc_full -> tcgen05.ld x_dbl -> B / C rows out, dt operand -> acc_empty;  dt_proj per chunk (bar.sync among the
//               four) -> dt_full -> tcgen05.ld -> staged block -> TMA store
// Hand-over is mbarriers only; the CTA-wide barriers are set-up and tear-down.  History and measurements: DESIGN.md §4.2,
// profiles/r2_call13..25*, descriptor probe scripts/umma_probe.cu.
#include "common.cuh"
#include "scan_common.cuh"

namespace cad {
namespace umma {

constexpr int XT = 128;            // tokens per tile = UMMA M
constexpr int KC = 32;             // channels per K slab
constexpr int XP = 152;            // pitch of a raw x row (144 used: 8-token aprons either side); 304 B = 19 x 16 B (odd)
constexpr int NS = 3;              // ring depth: the x slab, the W_x slab and the u slab of one K slab share a stage
constexpr int XPROJ_N = 48;        // dt rows (padded to 16) + B rows + C rows
constexpr int DTN = 128;           // channels per dt_proj instruction
constexpr int TMEM_COLS = 256;     // [0,64) / [64,128): x_proj accumulators of alternating tiles; [128,256): dt_proj chunk

constexpr int XS_BYTES = KC * XP * 2;              // 9728
constexpr int WX_BYTES = XPROJ_N * KC * 2;         // 3072
constexpr int U_BYTES = KC * XT * 2;               // 8192
constexpr int STAGE_BYTES = XS_BYTES + WX_BYTES + U_BYTES;
constexpr int DT_BYTES = XT * 16 * 2;              // 4096
constexpr int STG_BYTES = 32 * 32 * 2;             // one staged (32 channels x 32 tokens) block of delta
static_assert(STAGE_BYTES % 512 == 0 && DT_BYTES % 512 == 0, "the staging blocks behind the ring must stay 512-byte aligned (64-byte swizzle)");

constexpr int NCONV = 8;           // warps 0..7 convolve
constexpr int W_PROD = 8;          // warp 8 requests the slabs (TMA)
constexpr int W_MMA = 9;           // warp 9, lane 0 issues the x_proj MMAs
constexpr int W_EPI = 10;          // warps 10..13 drain TMEM (one TMEM lane quarter each: warp & 3)
constexpr int NTHREADS = 448;

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

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// mbarrier wait with a watchdog: a barrier that does not complete within 2 s (a launch takes a fraction of a millisecond) traps.
// Plain try_wait retry: a suspend-time hint compiles to NANOSLEEP.SYNCS + retry, which saves no issue slots and wakes up
// later (measured 3 % slower: profiles/r2_call17_umma_wait_flags.log).
__device__ __forceinline__ void mbar_park(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  uint64_t t_start = 0;
  for (;;) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t_start == 0) t_start = t;
    else if (t - t_start > 2000000000ull) __trap();
  }
}

struct Maps { CUtensorMap x, w_dt, w_bc, delta; };   // x slabs, W_x rows [0,R) / [R,R+32) (loads); delta blocks (stores)

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 2) conv_xproj_umma_kernel(cad_conv_xproj_args a, const __grid_constant__ Maps maps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // the swizzled staging blocks need it
  const int64_t E = a.E;
  const int E128 = (int)((E + 127) / 128 * 128);
  // stage s: [x slab: KC x XP raw samples | W_x slab, K-major operand [k group 4][n group 6][8][16 B] | u slab, MN-major operand
  // [k group 4][token group 16][8 ch][16 B]]
  unsigned char* stages = smem;
  unsigned char* dts = stages + NS * STAGE_BYTES;             // dt rows, K-major operand: [k group 2][token group 16][8 tok][16 B]
  unsigned char* stg = dts + DT_BYTES;                        // [4 warps][2] staged delta blocks (32 ch x 32 tok, 64-byte swizzle)
  unsigned char* wdts = stg + 8 * STG_BYTES;                  // W_dt, K-major operand: [k group 2][channel group E128/8][8][16 B], zero rows past E
  float4* taps = reinterpret_cast<float4*>(wdts + (size_t)E128 * 32);   // [E] conv taps, reversed for anti-causal jobs, x 1/2 (silu(v) = h + h tanh(h), h = v/2)
  float* cbias = reinterpret_cast<float*>(taps + E);          // [E] conv bias (same scaling)
  __shared__ uint64_t x_full[NS], w_full[NS], x_empty[NS], u_full[NS], slab_done[NS], acc_full[2], acc_empty[2], dt_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int job = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  const int64_t L = a.L;
  const int R = (int)a.R, N = (int)a.N;
  const T* __restrict__ wdt = static_cast<const T*>(a.w_dt) + (int64_t)pset * E * R;
  const T zero = io<T>::from_f(0.f);
  const int nslab = (int)(E / KC);
  const int64_t ntiles = (L + XT - 1) / XT;
  const int my_tiles = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int total = my_tiles * nslab;                         // slabs this CTA walks, g = 0 .. total-1 across its tiles
  const int64_t tile_step = (int64_t)gridDim.x * XT;
  const uint32_t wdt_k = (uint32_t)(E128 / 8) * 128;            // K stride of the W_dt operand

  // ---- set-up (all warps): barriers, TMEM, zero dt rows of the W_x slabs, resident W_dt and conv taps ---------------------
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&w_full[i], 1); mbar_init(&x_empty[i], NCONV); mbar_init(&u_full[i], NCONV); mbar_init(&slab_done[i], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1); mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);
    mbar_init(&dt_full, 1);
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w_dt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w_bc)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.delta)) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (R < 16)                                                  // operand rows [R, 16) of every k group: never written by a copy
    for (int i = tid; i < NS * (WX_BYTES / 16); i += NTHREADS) {
      const int st = i / (WX_BYTES / 16), v = i - st * (WX_BYTES / 16);
      reinterpret_cast<uint4*>(stages + st * STAGE_BYTES + XS_BYTES)[v] = make_uint4(0, 0, 0, 0);
    }
  if (R == 16) {
    for (int i = tid; i < E128 * 2; i += NTHREADS) {
      const int ch = i >> 1, kg = i & 1;
      *reinterpret_cast<uint4*>(wdts + kg * wdt_k + (ch >> 3) * 128 + (ch & 7) * 16) =
          ch < E ? __ldg(reinterpret_cast<const uint4*>(wdt + (int64_t)ch * 16 + 8 * kg)) : make_uint4(0, 0, 0, 0);
    }
  } else {
    for (int i = tid; i < E128 * 16; i += NTHREADS) {
      const int ch = i >> 4, r = i & 15;
      *reinterpret_cast<T*>(wdts + (r >> 3) * wdt_k + (ch >> 3) * 128 + (ch & 7) * 16 + (r & 7) * 2) =
          (r < R && ch < E) ? wdt[(int64_t)ch * R + r] : zero;
    }
  }
  for (int ch = tid; ch < (int)E; ch += NTHREADS) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(a.conv_w) + (int64_t)pset * E + ch);
    taps[ch] = rev ? make_float4(0.5f * w.w, 0.5f * w.z, 0.5f * w.y, 0.5f * w.x) : make_float4(0.5f * w.x, 0.5f * w.y, 0.5f * w.z, 0.5f * w.w);
    cbias[ch] = 0.5f * a.conv_b[(int64_t)pset * E + ch];
  }
  proxy_fence();                                               // W_dt / zero rows (generic stores) -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == W_PROD) {
    // ===== x slab requests: box [c0, c0+32) x [t0-8, t0+144), zero fill outside [0, L); a slot is free as soon as the conv
    //       warps hold its samples in registers, so the requests run a full ring ahead of the arithmetic ====================
    int s = 0, ph = 0, sl = 0;
    int64_t t0 = (int64_t)blockIdx.x * XT;
    for (int g = 0; g < total; ++g) {
      if (g >= NS) mbar_park(&x_empty[s], ph ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&x_full[s], XS_BYTES);
        tma_load_3d(stages + s * STAGE_BYTES, &maps.x, (int)(t0 - 8), sl * KC, seq, &x_full[s]);
      }
      __syncwarp();
      if (++sl == nslab) { sl = 0; t0 += tile_step; }
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp == W_MMA) {
    // ===== x_proj: D1[128 tokens x 48] += u^T . W_x^T, accumulator of tile `it` in TMEM columns [64 (it & 1), +48).  The same
    //       warp requests the W_x slabs (columns [c0, c0+32) as 2 x 4 boxes of 8 columns: dt rows -> operand rows [0, R), B / C
    //       rows -> operand rows [16, 48)) two slabs ahead: a slot is free once the MMAs that read it have completed ==========
    const uint32_t idesc_x = instr_desc<T>(XT, XPROJ_N, 1);
    const uint32_t wbytes = (uint32_t)(R + 32) * KC * 2;
    auto request_w = [&](int gw) {                             // slab gw -> slot gw % NS (all lanes call; lanes 0..7 copy)
      const int slot = gw % NS, c0 = (gw % nslab) * KC;
      unsigned char* wb = stages + slot * STAGE_BYTES + XS_BYTES;
      if (a.w_x_packed) {                                      // pre-packed operand: the slab is one contiguous 3 KB block
        if (lane == 0) {
          mbar_expect_tx(&w_full[slot], WX_BYTES);
          bulk_load_1d(smem_u32(wb), static_cast<const unsigned char*>(a.w_x_packed) + ((size_t)pset * nslab + (gw % nslab)) * WX_BYTES,
                       WX_BYTES, &w_full[slot]);
        }
        __syncwarp();
        return;
      }
      if (lane == 0) mbar_expect_tx(&w_full[slot], wbytes);
      __syncwarp();
      if (lane < 4) tma_load_2d(wb + lane * (XPROJ_N / 8 * 128), &maps.w_dt, c0 + 8 * lane, pset * (R + 32), &w_full[slot]);
      else if (lane < 8) tma_load_2d(wb + (lane - 4) * (XPROJ_N / 8 * 128) + 256, &maps.w_bc, c0 + 8 * (lane - 4), pset * (R + 32) + R, &w_full[slot]);
    };
    for (int gw = 0; gw < NS - 1 && gw < total; ++gw) request_w(gw);
    int s = 0, ph = 0, sl = 0, it = 0;
    for (int g = 0; g < total; ++g) {
      if (sl == 0 && it >= 2) mbar_park(&acc_empty[it & 1], (uint32_t)(((it >> 1) - 1) & 1));
      mbar_park(&w_full[s], ph);
      mbar_park(&u_full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t wa = smem_u32(stages + s * STAGE_BYTES + XS_BYTES), ua = wa + WX_BYTES;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks)
          mma_f16(tmem + (uint32_t)(64 * (it & 1)), smem_desc(ua + ks * 2 * (XT / 8 * 128), XT / 8 * 128, 128),
                  smem_desc(wa + ks * 2 * (XPROJ_N / 8 * 128), XPROJ_N / 8 * 128, 128), idesc_x, (sl | ks) ? 1u : 0u);
        mma_commit(&slab_done[s]);
        if (sl == nslab - 1) mma_commit(&acc_full[it & 1]);
      }
      __syncwarp();
      if (g + NS - 1 < total) {                                // W_x of slab g + NS - 1 -> the slot slab g - 1 used
        if (g >= 1) mbar_park(&slab_done[s == 0 ? NS - 1 : s - 1], s == 0 ? ph ^ 1 : ph);
        request_w(g + NS - 1);
      }
      if (++sl == nslab) { sl = 0; ++it; }
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp < NCONV) {
    // ===== conv + SiLU: per slab this thread's two (channel, 8-token vector) pieces ==========================================
    const int q = warp & 3, half = warp >> 2;
    const int tg = 4 * q + (lane >> 3);                        // token group 0..15
    const int cl = lane & 7;                                   // channel inside a group of 8
    const T* halo = a.halo ? static_cast<const T*>(a.halo) + (int64_t)job * E * 3 : nullptr;
    int s = 0, ph = 0, sl = 0;
    int64_t t0 = (int64_t)blockIdx.x * XT;
    for (int g = 0; g < total; ++g) {
      unsigned char* st = stages + s * STAGE_BYTES;
      mbar_park(&x_full[s], ph);
      uint4 xr[2][3];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint4* xv = reinterpret_cast<const uint4*>(st + ((8 * (2 * half + i) + cl) * XP + 8 * tg) * 2);   // x[t-8 .. t+15], t = t0 + 8 tg
        xr[i][0] = xv[0]; xr[i][1] = xv[1]; xr[i][2] = xv[2];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[s]);                 // my warp holds its samples in registers
      if (g >= NS) mbar_park(&slab_done[s], ph ^ 1);   // the MMAs that read u[s] one round ago have completed
      const bool edge = halo && ((!rev && t0 == 0) || (rev && t0 + XT >= L));
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kg = 2 * half + i;                           // channel group 0..3
        const int ch = sl * KC + 8 * kg + cl;
        const float4 cw = taps[ch];
        const float cb = cbias[ch];
        const uint4 r0 = xr[i][0], r1 = xr[i][1], r2 = xr[i][2];
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
        if (edge) {                                            // shard hook: the 3 samples that logically precede the shard
          const int64_t tw = t0 + 8 * tg + (rev ? -1 : -4);     // physical position of win[0]
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const int64_t te = tw + k;
            const int64_t tau = rev ? L - 1 - te : te;          // logical time
            if (tau >= -3 && tau < 0) win[k] = io<T>::to_f(halo[(int64_t)ch * 3 + tau + 3]);
          }
        }
        uint4 outv;
        uint32_t* o = reinterpret_cast<uint32_t*>(&outv);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {                       // outputs t+e, t+e+1 (taps pre-scaled: c = v/2)
          float c0v = cb + cw.x * win[e + 1] + cw.y * win[e + 2] + cw.z * win[e + 3] + cw.w * win[e + 4];
          float c1v = cb + cw.x * win[e + 2] + cw.y * win[e + 3] + cw.z * win[e + 4] + cw.w * win[e + 5];
          c0v = fmaf(c0v, tanh_approx(c0v), c0v);
          c1v = fmaf(c1v, tanh_approx(c1v), c1v);
          o[e >> 1] = pack2<T>(c0v, c1v);
        }
        // core matrix (k group kg, token group tg), row = channel within the group: a quarter-warp writes 128 contiguous bytes
        *reinterpret_cast<uint4*>(st + XS_BYTES + WX_BYTES + kg * (XT / 8 * 128) + tg * 128 + cl * 16) = outv;
      }
      proxy_fence();                                           // my u stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&u_full[s]);
      if (++sl == nslab) { sl = 0; t0 += tile_step; }
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else if (warp >= W_EPI) {
    // ===== tile epilogue: x_dbl out of TMEM, dt_proj, delta out ================================================================
    const int q = warp & 3, we = warp - W_EPI;                 // my TMEM lane quarter; my staging buffers
    const uint32_t tlane = (uint32_t)(32 * q) << 16;
    const uint32_t idesc_dt = instr_desc<T>(128, XT, 0);
    const int nchunk = E128 / DTN;
    unsigned char* mystg = stg + we * 2 * STG_BYTES;
    uint32_t ndt = 0, nstg = 0;                                // dt_full phases consumed; staged blocks written
    int64_t t0 = (int64_t)blockIdx.x * XT;
    for (int it = 0; it < my_tiles; ++it, t0 += tile_step) {
      mbar_park(&acc_full[it & 1], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t acc = tmem + tlane + (uint32_t)(64 * (it & 1));
      const int64_t t = t0 + 32 * q + lane;                     // my token (TMEM lane)
      {
        uint32_t v[16], w[16];
        tmem_ld16(acc, v);                                      // dt columns [0,16) -> io dtype -> dt operand (K-major rows = tokens)
        tmem_ld16(acc + 16, w);                                 // B columns
        tmem_ld_wait();
        uint4 lo, hi;
        lo.x = pack2<T>(__uint_as_float(v[0]), __uint_as_float(v[1]));   lo.y = pack2<T>(__uint_as_float(v[2]), __uint_as_float(v[3]));
        lo.z = pack2<T>(__uint_as_float(v[4]), __uint_as_float(v[5]));   lo.w = pack2<T>(__uint_as_float(v[6]), __uint_as_float(v[7]));
        hi.x = pack2<T>(__uint_as_float(v[8]), __uint_as_float(v[9]));   hi.y = pack2<T>(__uint_as_float(v[10]), __uint_as_float(v[11]));
        hi.z = pack2<T>(__uint_as_float(v[12]), __uint_as_float(v[13])); hi.w = pack2<T>(__uint_as_float(v[14]), __uint_as_float(v[15]));
        const int row = 32 * q + lane;
        *reinterpret_cast<uint4*>(dts + (row >> 3) * 128 + (row & 7) * 16) = lo;
        *reinterpret_cast<uint4*>(dts + (XT / 8 * 128) + (row >> 3) * 128 + (row & 7) * 16) = hi;
        tmem_ld16(acc + 32, v);                                 // C columns
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[it & 1]);        // the accumulator may be overwritten by tile it + 2
        const bool live = t < L;
#pragma unroll
        for (int j = 0; j < 16; ++j) { if (!live) { v[j] = 0u; w[j] = 0u; } }
        if (t < a.ldbc) {
          float* dst = a.bc + (int64_t)job * 2 * N * a.ldbc + t;
#pragma unroll
          for (int j = 0; j < 16; ++j) dst[(int64_t)j * a.ldbc] = __uint_as_float(w[j]);
#pragma unroll
          for (int j = 0; j < 16; ++j) dst[(int64_t)(16 + j) * a.ldbc] = __uint_as_float(v[j]);
        }
        if (a.bcT && t < a.ldT) {
          uint4* dT = reinterpret_cast<uint4*>(a.bcT + ((int64_t)job * a.ldT + t) * (2 * N));
#pragma unroll
          for (int j = 0; j < 4; ++j) dT[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
#pragma unroll
          for (int j = 0; j < 4; ++j) dT[4 + j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      // dt_proj, channels on the TMEM lanes: D2[128 channels x 128 tokens] = W_dt[chunk] . dt^T in TMEM columns [128, 256): a
      // thread reads 32 consecutive tokens of ONE delta row; the warp stages (32 channels x 32 tokens) and one lane stores the
      // block with a bulk tensor copy (full 64-byte row segments leave the SM without touching the LSU)
      for (int c = 0; c < nchunk; ++c) {
        proxy_fence();                                         // (c == 0) my dt operand rows -> visible to the tensor core
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");         // the four epilogue warps: operand complete / chunk buffer drained
        if (we == 0 && lane == 0) {
          tc_fence_after();
          mma_f16(tmem + 128u, smem_desc(smem_u32(wdts) + (uint32_t)(c * DTN / 8) * 128, wdt_k, 128),
                  smem_desc(smem_u32(dts), XT / 8 * 128, 128), idesc_dt, 0u);
          mma_commit(&dt_full);
        }
        mbar_park(&dt_full, ndt & 1); ++ndt;
        tc_fence_after();
        const int ch0 = c * DTN + 32 * q;                       // my warp's 32 channels (TMEM lanes)
#pragma unroll 1
        for (int cc = 0; cc < XT; cc += 32) {                  // 32 tokens at a time
          uint32_t v[32];
          tmem_ld32(tmem + tlane + 128u + (uint32_t)cc, v);
          tmem_ld_wait();
          unsigned char* blk = mystg + (nstg & 1) * STG_BYTES;
          if (lane == 0) bulk_wait_read<1>();                  // the store that last read this buffer has drained it
          __syncwarp();
          // row = my channel (64 bytes = 32 tokens), 16-byte chunk j at j ^ ((row >> 1) & 3): the 64-byte swizzle of the map
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(blk + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(pack2<T>(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                           pack2<T>(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                           pack2<T>(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                           pack2<T>(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
          proxy_fence();
          __syncwarp();
          if (lane == 0) {                                      // rows >= E and tokens >= ldd are clipped by the tensor map
            tma_store_3d(&maps.delta, blk, (int)(t0 + cc), ch0, job);
            bulk_commit();
          }
          ++nstg;
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace umma
}  // namespace cad

namespace cad {
namespace umma {
static int encode_map(CUtensorMap* m, int is_bf16, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return -1; }
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cad_conv_xproj_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
  return 0;
}
}  // namespace umma
}  // namespace cad

extern "C" int cad_conv_xproj_fwd(const cad_conv_xproj_args* a, void* stream_) {
  using namespace cad;
  using namespace cad::umma;
  CAD_REQUIRE(a, "cad_conv_xproj_fwd: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_conv_xproj_fwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->w_x && a->w_dt && a->conv_w && a->conv_b && a->seq_of_job && a->pset_of_job &&
              a->rev_of_job && a->delta && a->bc, "cad_conv_xproj_fwd: null pointer");
  CAD_REQUIRE(a->io_dtype == CAD_BF16 || a->io_dtype == CAD_F16,
              "cad_conv_xproj_fwd: tensor-core path needs 16-bit I/O (fp32 uses the unfused path)");
  CAD_REQUIRE(a->N == 16 && a->R >= 1 && a->R <= 16, "cad_conv_xproj_fwd: needs d_state = 16 and dt_rank <= 16");
  CAD_REQUIRE(a->E % 64 == 0 && a->E <= 2048, "cad_conv_xproj_fwd: d_inner must be a multiple of 64, <= 2048");
  CAD_REQUIRE(a->ldxz % 8 == 0 && a->ldd % 8 == 0 && a->ldxz >= a->L && a->ldd >= a->L && a->ldbc >= a->L,
              "cad_conv_xproj_fwd: bad row pitches");
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->w_x) && aligned16(a->delta) && aligned16(a->conv_w) &&
              (a->R != 16 || aligned16(a->w_dt)), "cad_conv_xproj_fwd: alignment");
  CAD_REQUIRE(!a->bcT || (aligned16(a->bcT) && a->ldT >= a->L), "cad_conv_xproj_fwd: bcT must be 16-byte aligned with ldT >= L");
  CAD_REQUIRE(!a->w_x_packed || aligned16(a->w_x_packed), "cad_conv_xproj_fwd: w_x_packed must be 16-byte aligned");
  CAD_REQUIRE(a->L < ((int64_t)1 << 31) - 1024, "cad_conv_xproj_fwd: sequence too long for 32-bit tensor-map coordinates");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int bf = a->io_dtype == CAD_BF16;
  const cuuint64_t E = (cuuint64_t)a->E, R = (cuuint64_t)a->R;
  Maps maps;
  {
    // x half of xz as (L tokens, 2E rows, nseq): one box = 152 tokens x 32 channels; tokens outside [0, L) read as zero
    const cuuint64_t dims[3] = {(cuuint64_t)a->L, 2 * E, (cuuint64_t)a->nseq};
    const cuuint64_t strides[2] = {(cuuint64_t)a->ldxz * 2, (cuuint64_t)a->ldxz * 2 * 2 * E};
    const cuuint32_t box[3] = {(cuuint32_t)XP, (cuuint32_t)KC, 1};
    if (encode_map(&maps.x, bf, a->xz, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
  }
  {
    // W_x as (E columns, rows): boxes of 8 columns x R rows (dt rows) and 8 columns x 32 rows (B / C rows); the row count is only
    // a clipping bound (the number of parameter sets is not part of the argument block): coordinates come from the job tables
    const cuuint64_t dims[2] = {E, (R + 32) * 4096};
    const cuuint64_t strides[1] = {E * 2};
    const cuuint32_t box_dt[2] = {8, (cuuint32_t)R}, box_bc[2] = {8, 32};
    if (encode_map(&maps.w_dt, bf, a->w_x, 2, dims, strides, box_dt, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    if (encode_map(&maps.w_bc, bf, a->w_x, 2, dims, strides, box_bc, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
  }
  {
    // delta as (ldd tokens, E rows, njobs): stores of (32 tokens x 32 rows) blocks staged with the 64-byte swizzle
    const cuuint64_t dims[3] = {(cuuint64_t)a->ldd, E, (cuuint64_t)a->njobs};
    const cuuint64_t strides[2] = {(cuuint64_t)a->ldd * 2, (cuuint64_t)a->ldd * 2 * E};
    const cuuint32_t box[3] = {32, 32, 1};
    if (encode_map(&maps.delta, bf, a->delta, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
  }
  const int64_t E128 = (a->E + 127) / 128 * 128;
  const size_t smem = 1024 + (size_t)NS * STAGE_BYTES + DT_BYTES + 8 * STG_BYTES + (size_t)E128 * 32 + (size_t)a->E * 20;
  const int64_t ntiles = (a->L + XT - 1) / XT;
  const int sms = cad_sm_count();
  CAD_REQUIRE(sms > 0, "cad_conv_xproj_fwd: no CUDA device");
  // persistent CTAs: at most two per SM in total (one wave), shared evenly by the jobs (every job has the same number of tiles)
  int64_t per_job = (2 * (int64_t)sms) / a->njobs;
  if (per_job > ntiles) per_job = ntiles;
  if (per_job < 1) per_job = 1;
  dim3 grid((unsigned)per_job, (unsigned)a->njobs);
  cudaError_t e;
  if (bf) {
    e = cudaFuncSetAttribute(conv_xproj_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_umma_kernel<__nv_bfloat16><<<grid, NTHREADS, smem, stream>>>(*a, maps);
  } else {
    e = cudaFuncSetAttribute(conv_xproj_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv_xproj_umma_kernel<__half><<<grid, NTHREADS, smem, stream>>>(*a, maps);
  }
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  CAD_LAUNCH_CHECK();
  return 0;
}

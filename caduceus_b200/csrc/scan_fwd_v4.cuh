// Fused bidirectional selective-scan forward, "paired-channel" variant (v4) for sm_100a.
//
// Same operator as scan_fwd.cu (conv + SiLU, softplus(dt), h_t = exp2(dt*A2) h_{t-1} + dt*B_t*u_t, y = C.h + D u,
// out = y * silu(z); ref call chain: ref:caduceus/modeling_caduceus.py:128-137, ref:caduceus/modeling_rcps.py:85-99,
// upstream selective_scan_fwd / causal_conv1d_fwd) with a different work decomposition,
// chosen from what profiles/r1_v3_scan_and_bwd_ncu_summary.txt shows binds v3 (issue slots 57 %, MUFU 53 %,
// shared-memory wavefronts 56 % — three co-bottlenecks at one warp per channel):
//
//   warp  = TWO adjacent channels of one job.  B_t and C_t are shared by all channels, so every quantity of the
//           recurrence is a (channel 0, channel 1) pair held in an aligned 64-bit register pair and processed by
//           Blackwell's packed fp32 instructions (FMUL2 / FFMA2, PTX mul/fma.rn.f32x2) with B / C as the broadcast
//           scalar operand: half the FMA-pipe instructions and half the shared-memory reads per element.  Only the
//           exp2 stays scalar (one MUFU per (token, channel, state), the floor of this operator).
//   lane  = 16 consecutive tokens of a 512-token chunk (as v3): zero-state recurrence per lane, 5-step warp-shuffle
//           scan of the (decay, state) aggregates, replay from the true incoming state.
//   MUFU software pipeline: the 32 exp2 of state n+1 are issued while state n runs its FMA chains and shuffles
//           (loop unrolled by two with two `a` buffers), so the MUFU pipe always has work queued.
//   smem  = TWO TMA tiles (2 x 64 KB, SWIZZLE_128B as v3): the request for chunk c+2 is issued when chunk c's
//           state loop ends, a whole chunk ahead of its use.  x / dt_raw / z of both channels are staged one
//           chunk ahead by per-lane cp.async (v3's scheme).  One CTA of up to 7 warps per SM (<= 255 registers).
//   scope = inference: 16-bit I/O, even E, no halo / carry-in / state outputs / saved chunk states (the sharded and
//           training paths keep v3).  Selected with cad_scan_fwd_args.variant = 4.
//
// The kernel body is written against a small set of SIMT primitives so that tests/emu/ can compile THIS file for
// the host (-DCAD_EMULATE: lanes are threads, shuffles are exchanges, TMA/mbarrier are modelled) and check the
// index logic against the oracle without a GPU.
#pragma once

#ifdef CAD_EMULATE
#include "simt_emu.h"
#else
#include "scan_common.cuh"

namespace cad {
namespace v4 {
#define CAD_DEV __device__ __forceinline__
#define CAD_TID ((int)threadIdx.x)
#define CAD_NTHREADS ((int)blockDim.x)
#define CAD_BIDX ((int)blockIdx.x)
#define CAD_BIDY ((int)blockIdx.y)
typedef CUtensorMap tmap_t;

CAD_DEV unsigned long long& as_u64(float2& v) { return *reinterpret_cast<unsigned long long*>(&v); }
CAD_DEV const unsigned long long& as_u64(const float2& v) { return *reinterpret_cast<const unsigned long long*>(&v); }
CAD_DEV float2 fma2(const float2& a, const float2& b, const float2& c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
  return d;
}
CAD_DEV float2 mul2(const float2& a, const float2& b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
CAD_DEV float2 add2(const float2& a, const float2& b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
CAD_DEV float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
CAD_DEV void sts64(uint32_t addr, const float2& v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
CAD_DEV uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
CAD_DEV void cp_async16s(uint32_t smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc));
}
CAD_DEV void cp_commit() { asm volatile("cp.async.commit_group;"); }
CAD_DEV void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
CAD_DEV float2 shfl_up2(const float2& v, int off) {
  return make_float2(__shfl_up_sync(0xffffffffu, v.x, off), __shfl_up_sync(0xffffffffu, v.y, off));
}
CAD_DEV float2 shfl_idx2(const float2& v, int src) {
  return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
CAD_DEV void cta_sync() { __syncthreads(); }
CAD_DEV void stg128(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
// one step of the inclusive (decay, state) warp scan on channel pairs: lanes >= OFF absorb lane - OFF
template <int OFF>
CAD_DEV void scan_step_up2(float2& P, float2& H, int lane) {
  const float2 Pp = shfl_up2(P, OFF), Hp = shfl_up2(H, OFF);
  // predicated SCALAR updates on the halves of the pairs: ptxas turns a predicated FFMA2/FMUL2 into compute + 2 SEL
  asm("{\n.reg .pred q;\nsetp.ge.s32 q, %8, %9;\n"
      "@q fma.rn.f32 %0, %2, %4, %0;\n@q fma.rn.f32 %1, %3, %5, %1;\n"
      "@q mul.f32 %2, %2, %6;\n@q mul.f32 %3, %3, %7;\n}"
      : "+f"(H.x), "+f"(H.y), "+f"(P.x), "+f"(P.y) : "f"(Hp.x), "f"(Hp.y), "f"(Pp.x), "f"(Pp.y), "r"(lane), "n"(OFF));
}
}  // namespace v4
}  // namespace cad
#endif  // !CAD_EMULATE

namespace cad {
namespace v4 {

constexpr int TOK = 16;             // tokens per lane
constexpr int CH = 32 * TOK;        // 512 tokens per chunk
constexpr int NST = 16;             // d_state
constexpr int kMaxG4 = 7;           // warps (channel pairs) per CTA
constexpr int kTileBytes = 2 * NST * CH * 4;
constexpr int kRows = 6;            // staged rows per warp: x0 x1 dt0 dt1 z0 z1

struct Smem {
  uint32_t tile[2];    // shared-space byte addresses of the two B/C tiles (1024-byte aligned)
  uint32_t par;        // [G][16] floats: conv taps as (c0, c1) pairs k = 0..3, conv bias, dt bias, D, pad
  uint32_t a2;         // [G][NST] (c0, c1) pairs of A2
  uint32_t carry;      // [G][NST] (c0, c1) pairs of the running state
  uint32_t pre;        // [2][G][kRows][CH] 16-bit: x / dt_raw / z segments of the current and the next chunk
  uint64_t* bar;       // full[2]: TMA completion barriers of the two tiles
  unsigned char* base; // generic pointer to shared byte 0 of `tile[0]`'s address space (for the TMA destination)
};

CAD_DEV float2 splat(float v) { return make_float2(v, v); }
CAD_DEV float2 ex2_2(const float2& x) { return make_float2(ex2(x.x), ex2(x.y)); }

// silu on a channel pair (16-bit I/O form of common.cuh::silu_io: h + h*tanh(h), h = v/2)
CAD_DEV float2 silu2(const float2& v) {
  const float2 h = mul2(v, splat(0.5f));
  const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
  return fma2(h, t, h);
}
// softplus on a channel pair, same branches as common.cuh::softplus (threshold 20, log1p series for small e^v)
CAD_DEV float2 softplus2(const float2& v) {
  const float2 w = ex2_2(mul2(v, splat(kLog2e)));
  const float2 one_w = add2(w, splat(1.0f));
  const float2 lg = make_float2(lg2(one_w.x), lg2(one_w.y));
  float2 sp = mul2(lg, splat(kLn2));
  // series = w * (1 + w * (-1/2 + w * (1/3 - w/4)))   (the same polynomial as common.cuh::softplus, Horner form)
  float2 s = fma2(w, splat(-0.25f), splat(0.33333334f));
  s = fma2(w, s, splat(-0.5f));
  s = fma2(w, s, splat(1.0f));
  s = mul2(w, s);
  sp.x = (w.x < 0.015625f) ? s.x : sp.x;
  sp.y = (w.y < 0.015625f) ? s.y : sp.y;
  sp.x = v.x > 20.0f ? v.x : sp.x;
  sp.y = v.y > 20.0f ? v.y : sp.y;
  return sp;
}

// 16 consecutive 16-bit values (two 16-byte pieces at shared address `addr`) -> fp32
template <typename T>
CAD_DEV void load16(uint32_t addr, float (&v)[TOK]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint4 raw = lds128u(addr + 16 * h);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[8 * h + k] = io<T>::to_f(e[k]);
  }
}

struct ChunkCtx {
  int lane, sg;              // lane; this lane's segment of the chunk in PHYSICAL order (31 - lane when reversed)
  uint32_t poff[4];          // byte offsets of the segment's four 16-byte pieces inside a swizzled tile row
  uint32_t par_s, a2_s, carry_s;
  uint32_t pre_cur, pre_next;   // shared addresses of this lane's staged segment, row 0 (rows are CH elements apart)
  bool active;
};

// One 512-token chunk of one channel pair.  TAIL: the chunk straddles the sequence end (tokens t >= L are masked:
// x = 0, dt = 0 so the state passes through; for a reversed job they are the logically FIRST tokens).
template <typename T, bool REV, bool TAIL>
CAD_DEV void chunk(const cad_scan_fwd_args& a, const Smem& sm, const ChunkCtx& cx, const tmap_t* tmap, float2 (&prev3)[3],
                   int64_t tseg, int64_t tseg_next, bool stage_next, const T* __restrict__ g_x0, const T* __restrict__ g_d0,
                   T* __restrict__ g_o0, int buf, uint32_t parity, bool issue_tma, int tma_c1, int job_row) {
  const int64_t L = a.L;
  const int lane = cx.lane;
  auto phys = [](int i) { return REV ? TOK - 1 - i : i; };
  const bool seg_in = !TAIL || tseg < L;
  constexpr uint32_t ROW = CH * sizeof(T);     // bytes between staged rows

  float2 dt[TOK], du[TOK], y[TOK];
  float2 dsum = splat(0.f);
  {
    // ---- 1. staged x / dt_raw segments of both channels (shared -> registers), then stage the next chunk ----
    float x0[TOK], x1[TOK], r0[TOK], r1[TOK];
    if (seg_in) {
      cp_wait_all();                               // only this lane reads what it staged
      load16<T>(cx.pre_cur + 0 * ROW, x0);
      load16<T>(cx.pre_cur + 1 * ROW, x1);
      load16<T>(cx.pre_cur + 2 * ROW, r0);
      load16<T>(cx.pre_cur + 3 * ROW, r1);
    } else {
#pragma unroll
      for (int i = 0; i < TOK; ++i) { x0[i] = 0.f; x1[i] = 0.f; r0[i] = 0.f; r1[i] = 0.f; }
    }
    if (stage_next) {
      const T* gx = g_x0 + tseg_next;
      const T* gz = gx + a.E * a.ldxz;
      const T* gd = g_d0 + tseg_next;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cp_async16s(cx.pre_next + 0 * ROW + 16 * h, gx + 8 * h);
        cp_async16s(cx.pre_next + 1 * ROW + 16 * h, gx + a.ldxz + 8 * h);
        cp_async16s(cx.pre_next + 2 * ROW + 16 * h, gd + 8 * h);
        cp_async16s(cx.pre_next + 3 * ROW + 16 * h, gd + a.ldd + 8 * h);
        cp_async16s(cx.pre_next + 4 * ROW + 16 * h, gz + 8 * h);
        cp_async16s(cx.pre_next + 5 * ROW + 16 * h, gz + a.ldxz + 8 * h);
      }
      cp_commit();
    }

    // ---- 2. per-(token, channel) prologue: conv + SiLU, dt = softplus(dt_raw + b), dt*u, D*u ------------------
    const float2 cw0 = lds64(cx.par_s + 0), cw1 = lds64(cx.par_s + 8), cw2 = lds64(cx.par_s + 16), cw3 = lds64(cx.par_s + 24);
    const float2 cb = lds64(cx.par_s + 32), dtb = lds64(cx.par_s + 40), Dk = lds64(cx.par_s + 48);
    float2 xl[TOK + 3];
#pragma unroll
    for (int i = 0; i < TOK; ++i) {
      float2 v = make_float2(x0[phys(i)], x1[phys(i)]);
      if (TAIL && tseg + phys(i) >= L) v = splat(0.f);
      xl[i + 3] = v;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {                  // logical predecessors: previous lane / previous chunk
      const float2 up = shfl_up2(xl[TOK + k], 1);
      xl[k] = (lane == 0) ? prev3[k] : up;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) prev3[k] = shfl_idx2(xl[TOK + k], 31);
#pragma unroll
    for (int i = 0; i < TOK; ++i) {
      float2 acc = fma2(cw0, xl[i], cb);
      acc = fma2(cw1, xl[i + 1], acc);
      acc = fma2(cw2, xl[i + 2], acc);
      acc = fma2(cw3, xl[i + 3], acc);
      const float2 u = silu2(acc);
      float2 d = softplus2(add2(make_float2(r0[phys(i)], r1[phys(i)]), dtb));
      if (TAIL && tseg + phys(i) >= L) d = splat(0.f);      // masked token: a = 1, b = 0
      dt[i] = d;
      dsum = add2(dsum, d);
      du[i] = mul2(d, u);
      y[i] = mul2(Dk, u);
    }
  }

  // ---- 3. the scan, one state at a time, exp2 of the NEXT state in flight behind the current one -------------
  const uint32_t tile_s = buf ? sm.tile[1] : sm.tile[0];
  auto compute_a = [&](int n, float2 (&av)[TOK]) {
    const float2 A2n = lds64(cx.a2_s + 8 * n);
#pragma unroll
    for (int i = 0; i < TOK; ++i) av[i] = ex2_2(mul2(dt[i], A2n));
  };
  auto run_state = [&](int n, const float2 (&av)[TOK]) {
    const float2 A2n = lds64(cx.a2_s + 8 * n);
    const float2 cin = lds64(cx.carry_s + 8 * n);
    float2 bv[TOK];
    float2 hl = (lane == 0) ? cin : splat(0.f);
    {
      const uint32_t rowp = tile_s + n * (CH * 4);
      float4 q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = lds128(rowp + cx.poff[k]);
#pragma unroll
      for (int i = 0; i < TOK; ++i) {
        const int p = phys(i);
        const float4 qq = q[p >> 2];
        const float bq = (p & 3) == 0 ? qq.x : ((p & 3) == 1 ? qq.y : ((p & 3) == 2 ? qq.z : qq.w));
        bv[i] = mul2(du[i], splat(bq));
      }
#pragma unroll
      for (int i = 0; i < TOK; ++i) hl = fma2(av[i], hl, bv[i]);
    }
    float2 P = ex2_2(mul2(A2n, dsum));
    scan_step_up2<1>(P, hl, lane);
    scan_step_up2<2>(P, hl, lane);
    scan_step_up2<4>(P, hl, lane);
    scan_step_up2<8>(P, hl, lane);
    scan_step_up2<16>(P, hl, lane);
    float2 h = shfl_up2(hl, 1);
    if (lane == 0) h = cin;
    if (lane == 31) sts64(cx.carry_s + 8 * n, hl);          // state at the end of this chunk
    {
      const uint32_t rowp = tile_s + (NST + n) * (CH * 4);
      float4 q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = lds128(rowp + cx.poff[k]);
#pragma unroll
      for (int i = 0; i < TOK; ++i) {                        // logical order
        const int p = phys(i);
        const float4 qq = q[p >> 2];
        const float cq = (p & 3) == 0 ? qq.x : ((p & 3) == 1 ? qq.y : ((p & 3) == 2 ? qq.z : qq.w));
        h = fma2(av[i], h, bv[i]);
        y[i] = fma2(splat(cq), h, y[i]);
      }
    }
  };

  float2 avA[TOK], avB[TOK];
  compute_a(0, avA);                     // independent of the tile: overlaps the TMA wait
  mbar_wait(&sm.bar[buf], parity);
#pragma unroll 1
  for (int n = 0; n < NST; n += 2) {
    compute_a(n + 1, avB);
    run_state(n, avA);
    if (n + 2 < NST) compute_a(n + 2, avA);
    run_state(n + 1, avB);
  }

  // ---- 4. everyone is done reading this tile -> request chunk c+2 into it -------------------------------------
  cta_sync();
  if (issue_tma && CAD_TID == 0) {
    mbar_expect_tx(&sm.bar[buf], kTileBytes);
    tma_load_3d(sm.base + (buf ? kTileBytes : 0), tmap, 0, tma_c1, job_row, &sm.bar[buf]);
  }

  // ---- 5. gate with silu(z) and store (physical order) ---------------------------------------------------------
  if (seg_in && cx.active) {
    float z0[TOK], z1[TOK];
    load16<T>(cx.pre_cur + 4 * ROW, z0);
    load16<T>(cx.pre_cur + 5 * ROW, z1);
    float o0[TOK], o1[TOK];
#pragma unroll
    for (int i = 0; i < TOK; ++i) {
      const float2 g = silu2(make_float2(z0[phys(i)], z1[phys(i)]));
      const float2 o = mul2(y[i], g);
      o0[phys(i)] = o.x;
      o1[phys(i)] = o.y;
    }
    T* go = g_o0 + tseg;
    if (!TAIL || tseg + TOK <= L) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 r0, r1;
        T* e0 = reinterpret_cast<T*>(&r0);
        T* e1 = reinterpret_cast<T*>(&r1);
#pragma unroll
        for (int k = 0; k < 8; ++k) { e0[k] = io<T>::from_f(o0[8 * h + k]); e1[k] = io<T>::from_f(o1[8 * h + k]); }
        stg128(go + 8 * h, r0);
        stg128(go + a.ldo + 8 * h, r1);
      }
    } else {
#pragma unroll
      for (int i = 0; i < TOK; ++i)
        if (tseg + i < L) { go[i] = io<T>::from_f(o0[i]); go[a.ldo + i] = io<T>::from_f(o1[i]); }
    }
  }
}

template <typename T, bool REV>
CAD_DEV void run_job(const cad_scan_fwd_args& a, const tmap_t* tmap, int job, int seq, int pset, const Smem& sm) {
  const int lane = CAD_TID & 31, warp = CAD_TID >> 5, G = CAD_NTHREADS >> 5;
  const int64_t L = a.L, E = a.E, npair = E / 2;
  const int64_t pr = (int64_t)CAD_BIDX * G + warp;
  ChunkCtx cx;
  cx.lane = lane;
  cx.active = pr < npair;                       // tail CTA: idle warps only keep the barriers company
  const int64_t c0 = 2 * (cx.active ? pr : npair - 1);
  const int64_t nchunks = (L + CH - 1) / CH;
  const int64_t pc = (int64_t)pset * E + c0;

  cx.par_s = sm.par + warp * 64;
  cx.a2_s = sm.a2 + warp * (NST * 8);
  cx.carry_s = sm.carry + warp * (NST * 8);
  {
    const int k = lane >> 1, chn = lane & 1;
    float v = 0.f;
    if (lane < 8) v = a.conv_w[(pc + chn) * 4 + k];
    else if (lane < 10) v = a.conv_b[pc + chn];
    else if (lane < 12) v = a.dt_b[pc + chn];
    else if (lane < 14) v = a.Dskip[pc + chn];
    if (lane < 16) sts32(cx.par_s + 4 * lane, v);
    sts32(cx.a2_s + 4 * lane, a.A2[(pc + chn) * NST + k]);     // lane = 2 n + channel
    sts32(cx.carry_s + 4 * lane, 0.f);
  }

  cx.sg = REV ? 31 - lane : lane;
  tile_piece_offsets<TOK>(cx.sg, cx.poff);
  const int job_row = job * 2 * NST;
  constexpr int BPC = CH / kBlkTok;             // 32-token swizzle lines per chunk row

  cta_sync();                                   // barrier init + parameter staging visible
  if (CAD_TID == 0) {
    for (int k = 0; k < 2 && k < nchunks; ++k) {
      const int64_t pci = REV ? nchunks - 1 - k : k;
      mbar_expect_tx(&sm.bar[k], kTileBytes);
      tma_load_3d(sm.base + (k ? kTileBytes : 0), tmap, 0, (int)(pci * BPC), job_row, &sm.bar[k]);
    }
  }

  const T* __restrict__ g_x0 = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + c0) * a.ldxz;
  const T* __restrict__ g_d0 = static_cast<const T*>(a.delta) + ((int64_t)job * E + c0) * a.ldd;
  T* __restrict__ g_o0 = static_cast<T*>(a.out) + ((int64_t)job * E + c0) * a.ldo;
  auto pre_addr = [&](int b) {
    return sm.pre + (uint32_t)((((size_t)(b * G + warp) * kRows) * CH + (size_t)cx.sg * TOK) * sizeof(T));
  };
  {
    const int64_t ts0 = (REV ? nchunks - 1 : 0) * CH + (int64_t)cx.sg * TOK;
    if (ts0 < L) {
      constexpr uint32_t ROW = CH * sizeof(T);
      const uint32_t d = pre_addr(0);
      const T* gx = g_x0 + ts0;
      const T* gz = gx + E * a.ldxz;
      const T* gd = g_d0 + ts0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cp_async16s(d + 0 * ROW + 16 * h, gx + 8 * h);
        cp_async16s(d + 1 * ROW + 16 * h, gx + a.ldxz + 8 * h);
        cp_async16s(d + 2 * ROW + 16 * h, gd + 8 * h);
        cp_async16s(d + 3 * ROW + 16 * h, gd + a.ldd + 8 * h);
        cp_async16s(d + 4 * ROW + 16 * h, gz + 8 * h);
        cp_async16s(d + 5 * ROW + 16 * h, gz + a.ldxz + 8 * h);
      }
      cp_commit();
    }
  }

  float2 prev3[3] = {splat(0.f), splat(0.f), splat(0.f)};     // x before logical time 0 is zero (no halo in v4)
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t pci = REV ? nchunks - 1 - c : c;
    const int64_t pcn = REV ? pci - 1 : pci + 1;
    const int64_t tseg = pci * CH + (int64_t)cx.sg * TOK;
    const int64_t tseg_next = pcn * CH + (int64_t)cx.sg * TOK;
    const bool stage_next = (c + 1 < nchunks) && tseg_next < L;
    const bool issue_tma = c + 2 < nchunks;
    const int tma_c1 = (int)((REV ? pci - 2 : pci + 2) * BPC);
    const bool tail = (pci + 1) * CH > L;
    const int buf = (int)(c & 1);
    const uint32_t parity = (uint32_t)((c >> 1) & 1);
    cx.pre_cur = pre_addr(buf);
    cx.pre_next = pre_addr(buf ^ 1);
    if (tail)
      chunk<T, REV, true>(a, sm, cx, tmap, prev3, tseg, tseg_next, stage_next, g_x0, g_d0, g_o0, buf, parity, issue_tma,
                          tma_c1, job_row);
    else
      chunk<T, REV, false>(a, sm, cx, tmap, prev3, tseg, tseg_next, stage_next, g_x0, g_d0, g_o0, buf, parity, issue_tma,
                           tma_c1, job_row);
  }
}

// shared-memory plan (bytes from the 1024-aligned base): two tiles | par | a2 | carry | bars | staging
CAD_DEV void carve(unsigned char* base, int G, Smem& sm) {
  const uint32_t b = smem_u32(base);
  sm.base = base;
  sm.tile[0] = b;
  sm.tile[1] = b + kTileBytes;
  sm.par = b + 2 * kTileBytes;
  sm.a2 = sm.par + kMaxG4 * 64;
  sm.carry = sm.a2 + kMaxG4 * NST * 8;
  sm.bar = reinterpret_cast<uint64_t*>(base + 2 * kTileBytes + kMaxG4 * 64 + 2 * kMaxG4 * NST * 8);
  sm.pre = sm.carry + kMaxG4 * NST * 8 + 16;
  (void)G;
}
inline size_t smem_bytes(int G, size_t elem) {
  return 1024 + (size_t)2 * kTileBytes + kMaxG4 * 64 + (size_t)2 * kMaxG4 * NST * 8 + 16 + (size_t)2 * G * kRows * CH * elem;
}

template <typename T>
CAD_DEV void kernel_body(const cad_scan_fwd_args& a, const tmap_t* tmap, unsigned char* smem_raw) {
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Smem sm;
  carve(base, CAD_NTHREADS >> 5, sm);
  if (CAD_TID == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); }
  const int job = CAD_BIDY;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) run_job<T, true>(a, tmap, job, seq, pset, sm);
  else     run_job<T, false>(a, tmap, job, seq, pset, sm);
}

}  // namespace v4
}  // namespace cad

// Fused bidirectional selective-scan forward, variant 9 (sm_100a): the one-channel-per-warp kernel rebuilt around what
// the round-1 captures say binds variant 3 (profiles/r1_v3_scan_stall_breakdown.txt):
//
//   * 16-BIT B/C TILE.  Under autocast the reference's x_dbl — hence B_t, C_t — is 16-bit (upstream mamba_inner_fn), so
//     the tile is staged in the I/O dtype: ONE 16-byte shared load brings 8 tokens instead of 4 (half the LSU wavefronts
//     every warp issues per state, behind which the shuffles of the warp scan queue), the TMA / L2 tile traffic halves,
//     and TWO tile buffers fit in the space of v3's one (2 x 32 KB), keeping two CTAs per SM.
//   * NO CTA BARRIER IN THE CHUNK LOOP.  Each tile buffer has a TMA-completion mbarrier and an arrival counter: a warp
//     that has finished reading buffer b bumps the counter, and the LAST arriver re-arms the barrier and issues the TMA
//     for chunk c+2 into b.  Nobody waits for a sibling warp (v3 spends 10 % of its time in the per-chunk __syncthreads),
//     the request is a whole chunk ahead of its use, and warps are free to drift apart.
//   * NO REPLAY PASS (variant 7's state loop): the zero-state pass accumulates y += C.h_local and pc = prod a; the
//     carry-in enters after the warp scan as 8 independent packed FMAs y += (C pc) h_in.  Only C*pc (16 registers) is
//     live across the scan, which leaves room for
//   * an EXP2 SOFTWARE PIPELINE (template PIPE): the 16 exp2 of state n+1 are issued right before state n's shuffle
//     rounds, so the MUFU pipe has work while the warp waits on SHFL round trips.
//
// Template TT chooses the tile element type: the io dtype (variants 9 / 10: 2 x 32 KB, up to 7 warps, two CTAs per SM) or
// float (variants 11 / 12: 2 x 64 KB shared by up to 14 warps in ONE CTA per SM — no unpack instructions, half the TMA
// tile traffic per channel, same barrier-free hand-over).
//
// Same operator, argument block and hooks as scan_fwd.cu (conv halo, carry-in h0, end state, sum dt, saved chunk states,
// state-only pass) for 16-bit I/O; fp32 I/O stays on variant 3.  ref call chain: ref:caduceus/modeling_caduceus.py:128-137,
// ref:caduceus/modeling_rcps.py:85-99 -> upstream selective_scan_fwd / causal_conv1d_fwd.
// Written against the SIMT primitives of scan_fwd_v4.cuh so that tests/emu/ compiles THIS file for the host.
#pragma once
#include "scan_fwd_v4.cuh"

namespace cad {
namespace v9 {

#ifndef CAD_EMULATE
CAD_DEV uint32_t atomic_inc_shared(uint32_t addr) {          // returns the value before the increment
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
  return old;
}
CAD_DEV void sts32u(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
CAD_DEV void warp_sync() { __syncwarp(); }
CAD_DEV float shfl_up1(float v, int off) { return __shfl_up_sync(0xffffffffu, v, off); }
CAD_DEV float shfl_idx1(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
CAD_DEV float shfl_xor1(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <int OFF>
CAD_DEV void scan_step1(float& P, float& H, int lane) { scan_step_up<OFF>(P, H, lane); }
CAD_DEV float u2f(uint32_t u) { return __uint_as_float(u); }
#endif

using v4::fma2; using v4::mul2; using v4::splat; using v4::ex2_2; using v4::lds128u; using v4::cp_async16s;
using v4::cp_commit; using v4::cp_wait_all; using v4::cta_sync; using v4::stg128; using v4::tmap_t;

constexpr int TOK = 16;             // tokens per lane
constexpr int NP = TOK / 2;         // physical token pairs per lane
constexpr int CH = 32 * TOK;        // 512 tokens per chunk
constexpr int NST = 16;             // d_state
constexpr int kMaxG9 = 14;          // warps (channels) per CTA: <= 7 with a 16-bit tile (two CTAs per SM), <= 14 with an
                                    // fp32 tile (one CTA per SM; its two 64 KB buffers are shared by 14 channels)
constexpr int kRows = 3;            // staged rows per warp: x, dt_raw, z
// tile geometry for a tile element type TT (the io dtype, or float): a 128-byte swizzle line holds 128 / sizeof(TT)
// tokens; a tile row is 512 tokens; the tile is 2 N rows
template <typename TT> struct tile_geom {
  static constexpr int kLineTok = 128 / (int)sizeof(TT);
  static constexpr int kRowBytes = CH * (int)sizeof(TT);
  static constexpr int kTileBytes = 2 * NST * kRowBytes;          // 32 KB (16-bit) or 64 KB (fp32)
  static constexpr int kPieces = TOK * (int)sizeof(TT) / 16;      // 16-byte pieces per lane segment: 2 or 4
  static constexpr int kPairsPerPiece = NP / kPieces;             // 4 or 2
};

struct Smem {
  uint32_t tile[2];    // shared-space byte addresses of the two B/C tiles (1024-byte aligned)
  uint32_t par;        // [G][8] floats: conv taps 0..3, conv bias, dt bias, D, pad
  uint32_t a2;         // [G][NST]
  uint32_t carry;      // [G][NST] running state of each warp's channel
  uint32_t cnt;        // [2] arrival counters of the tile buffers
  uint32_t pre;        // [2][G][kRows][CH] 16-bit staging of x / dt_raw / z
  uint64_t* bar;       // full[2]
  unsigned char* base;
};

// (tokens 2j, 2j+1) of a packed 16-bit pair -> fp32 pair
template <typename T> CAD_DEV float2 unpack2(uint32_t w);
template <> CAD_DEV float2 unpack2<__nv_bfloat16>(uint32_t w) { return make_float2(u2f(w << 16), u2f(w & 0xffff0000u)); }
template <> CAD_DEV float2 unpack2<__half>(uint32_t w) {
  const __half2 h = *reinterpret_cast<const __half2*>(&w);
  return make_float2(__half2float(__low2half(h)), __half2float(__high2half(h)));
}

// this lane's 16-byte pieces inside a swizzled tile row: SWIZZLE_128B stores 16-byte chunk c of 128-byte line l at
// position c ^ (l & 7); a row is 8 (16-bit) or 16 (fp32) lines, so l & 7 == (line-in-row) & 7.
template <typename TT>
CAD_DEV void piece_offsets(int seg, uint32_t (&poff)[4]) {
  constexpr int P = tile_geom<TT>::kPieces, SPL = 8 / P;          // lane segments per line: 4 or 2
  const int line = seg / SPL, c0 = P * (seg % SPL);
#pragma unroll
  for (int k = 0; k < 4; ++k) poff[k] = k < P ? line * 128 + (((c0 + k) ^ (line & 7)) << 4) : 0u;
}
// the k-th physical token pair of a 16-byte piece -> fp32 pair
template <typename T, typename TT>
CAD_DEV float2 piece_pair(const uint4& q, int m) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  if constexpr (sizeof(TT) == 2) return unpack2<T>(w[m]);
  else return make_float2(u2f(w[2 * m]), u2f(w[2 * m + 1]));
}

struct ChunkCtx {
  int lane, sg, G;
  uint32_t poff[4];
  uint32_t par_s, a2_s, carry_s;
  uint32_t pre_cur, pre_next;
  bool active;
};

template <typename T, typename TT, bool REV, bool TAIL, bool STATE_ONLY, bool PIPE>
CAD_DEV void chunk(const cad_scan_fwd_args& a, const Smem& sm, const ChunkCtx& cx, const tmap_t* tmap, float (&prev3)[3],
                   const float (&hal)[3], float& dt_total, int64_t tseg, int64_t tseg_next, bool stage_next,
                   const T* __restrict__ g_x, const T* __restrict__ g_d, T* __restrict__ g_o, int buf, uint32_t parity,
                   bool issue_tma, int tma_c1, int job_row) {
  const int64_t L = a.L;
  const int lane = cx.lane;
  auto phys = [](int i) { return REV ? TOK - 1 - i : i; };
  auto lgc = [](int p) { return REV ? TOK - 1 - p : p; };
  auto halo_at = [&](int64_t tau) { return tau == -1 ? hal[2] : (tau == -2 ? hal[1] : (tau == -3 ? hal[0] : 0.f)); };
  const bool seg_in = !TAIL || tseg < L;
  constexpr uint32_t ROW = CH * sizeof(T);
  using TG = tile_geom<TT>;

  float2 dt2[NP], du2[NP], y2[NP];
  float dsum = 0.f;
  {
    // ---- 1. staged x / dt_raw segments (shared -> registers), then stage the next chunk ----------------------------
    float xs[TOK], dr[TOK];
    if (seg_in) {
      cp_wait_all();
      v4::load16<T>(cx.pre_cur + 0 * ROW, xs);
      if (a.delta_is_dt) v4::load16<__half>(cx.pre_cur + 1 * ROW, dr);   // dt itself, fp16 (written by conv_xproj)
      else v4::load16<T>(cx.pre_cur + 1 * ROW, dr);
    } else {
#pragma unroll
      for (int i = 0; i < TOK; ++i) { xs[i] = 0.f; dr[i] = 0.f; }
    }
    if (stage_next) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cp_async16s(cx.pre_next + 0 * ROW + 16 * h, g_x + tseg_next + 8 * h);
        cp_async16s(cx.pre_next + 1 * ROW + 16 * h, g_d + tseg_next + 8 * h);
        if (!STATE_ONLY) cp_async16s(cx.pre_next + 2 * ROW + 16 * h, g_x + a.E * a.ldxz + tseg_next + 8 * h);
      }
      cp_commit();
    }

    // ---- 2. prologue: conv + SiLU, dt = softplus(dt_raw + b), dt*u, D*u -------------------------------------------
    const float4 cw = lds128(cx.par_s);
    const float4 pr = lds128(cx.par_s + 16);              // conv bias, dt bias, D, pad
    float xl[TOK + 3];
#pragma unroll
    for (int i = 0; i < TOK; ++i) {
      float v = xs[phys(i)];
      if (TAIL) {
        const int64_t t = tseg + phys(i);
        if (t >= L) v = REV ? halo_at(L - 1 - t) : 0.f;
      }
      xl[i + 3] = v;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {                          // logical predecessors: previous lane / previous chunk
      const float up = shfl_up1(xl[TOK + k], 1);
      xl[k] = (lane == 0) ? prev3[k] : up;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) prev3[k] = shfl_idx1(xl[TOK + k], 31);
    float dt[TOK], du[TOK], y[TOK];
    if (a.delta_is_dt) {                                   // launch-uniform branch: no softplus (2 MUFU per token) here
#pragma unroll
      for (int i = 0; i < TOK; ++i) dt[i] = dr[phys(i)];
    } else {
#pragma unroll
      for (int i = 0; i < TOK; ++i) dt[i] = softplus(dr[phys(i)] + pr.y);
    }
#pragma unroll
    for (int i = 0; i < TOK; ++i) {
      const float u = silu_io<T>(pr.x + cw.x * xl[i] + cw.y * xl[i + 1] + cw.z * xl[i + 2] + cw.w * xl[i + 3]);
      float d = dt[i];
      if (TAIL && tseg + phys(i) >= L) d = 0.f;           // masked token: a = 1, b = 0 -> state passes through
      dt[i] = d;
      dsum += d;
      du[i] = d * u;
      y[i] = pr.z * u;
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {                         // adjacent PHYSICAL tokens in aligned register pairs
      dt2[j] = make_float2(dt[lgc(2 * j)], dt[lgc(2 * j + 1)]);
      du2[j] = make_float2(du[lgc(2 * j)], du[lgc(2 * j + 1)]);
      y2[j] = make_float2(y[lgc(2 * j)], y[lgc(2 * j + 1)]);
    }
  }
  dt_total += dsum;

  // ---- 3. the scan, one state at a time, on the TMA-staged 16-bit B/C tile ----------------------------------------
  const uint32_t tile_s = buf ? sm.tile[1] : sm.tile[0];
  auto compute_a = [&](int n, float2 (&av)[NP]) {
    const float A2n = lds32(cx.a2_s + 4 * n);
#pragma unroll
    for (int j = 0; j < NP; ++j) av[j] = ex2_2(mul2(dt2[j], splat(A2n)));
  };
  // one state: zero-state pass with y / pc accumulation, (optionally) the next state's exp2, warp scan, carry-in FMAs
  auto run_state = [&](int n, const float2 (&av)[NP], float2 (&av_next)[NP], bool do_next) {
    const float cin = lds32(cx.carry_s + 4 * n);
    float hl = 0.f, pc = 1.f;
    float2 g2[NP];
#pragma unroll
    for (int kk = 0; kk < TG::kPieces; ++kk) {             // 16-byte pieces of the tile rows in logical order
      const int k = REV ? TG::kPieces - 1 - kk : kk;
      const uint4 bq = lds128u(tile_s + n * TG::kRowBytes + cx.poff[k]);
      uint4 cq = make_uint4(0u, 0u, 0u, 0u);
      if (!STATE_ONLY) cq = lds128u(tile_s + (NST + n) * TG::kRowBytes + cx.poff[k]);
#pragma unroll
      for (int mm = 0; mm < TG::kPairsPerPiece; ++mm) {
        const int m = REV ? TG::kPairsPerPiece - 1 - mm : mm;
        const int j = TG::kPairsPerPiece * k + m;          // physical pair (tokens 2j, 2j+1) of the segment
        const float2 bv = mul2(du2[j], piece_pair<T, TT>(bq, m));
        float2 hp, pp;
        if (REV) {
          hl = fmaf(av[j].y, hl, bv.y); hp.y = hl; pc *= av[j].y; pp.y = pc;
          hl = fmaf(av[j].x, hl, bv.x); hp.x = hl; pc *= av[j].x; pp.x = pc;
        } else {
          hl = fmaf(av[j].x, hl, bv.x); hp.x = hl; pc *= av[j].x; pp.x = pc;
          hl = fmaf(av[j].y, hl, bv.y); hp.y = hl; pc *= av[j].y; pp.y = pc;
        }
        if (!STATE_ONLY) {
          const float2 cp = piece_pair<T, TT>(cq, m);
          y2[j] = fma2(cp, hp, y2[j]);
          g2[j] = mul2(cp, pp);
        }
      }
    }
    if (PIPE && do_next) compute_a(n + 1, av_next);       // MUFU work for the shuffle round trips below
    float P = pc;
    if (lane == 0) hl = fmaf(pc, cin, hl);                 // the chunk's carry-in enters through lane 0's aggregate
    scan_step1<1>(P, hl, lane);
    scan_step1<2>(P, hl, lane);
    scan_step1<4>(P, hl, lane);
    scan_step1<8>(P, hl, lane);
    scan_step1<16>(P, hl, lane);
    float h = shfl_up1(hl, 1);
    if (lane == 0) h = cin;
    if (lane == 31) sts32(cx.carry_s + 4 * n, hl);         // state at the end of this chunk
    if (!STATE_ONLY) {
#pragma unroll
      for (int j = 0; j < NP; ++j) y2[j] = fma2(g2[j], splat(h), y2[j]);
    }
  };

  if (PIPE) {
    float2 avA[NP], avB[NP];
    compute_a(0, avA);                                     // independent of the tile: overlaps the TMA wait
    mbar_wait_wd(&sm.bar[buf], parity);
    // the prefetch is unconditional inside the loop (straight-line code, so the exp2 interleave with the shuffle rounds);
    // the last pair of states is peeled because state NST-1 has nothing to prefetch
#pragma unroll 1
    for (int n = 0; n < NST - 2; n += 2) {
      run_state(n, avA, avB, true);
      run_state(n + 1, avB, avA, true);
    }
    run_state(NST - 2, avA, avB, true);
    run_state(NST - 1, avB, avA, false);
  } else {
    mbar_wait_wd(&sm.bar[buf], parity);
#pragma unroll 1
    for (int n = 0; n < NST; ++n) {
      float2 av[NP];
      compute_a(n, av);
      run_state(n, av, av, false);
    }
  }

  // ---- 4. release the tile buffer; the LAST warp to arrive requests chunk c+2 into it ------------------------------
  warp_sync();
  if (lane == 0) {
    const uint32_t cnt_s = sm.cnt + 4 * buf;
    if (atomic_inc_shared(cnt_s) == (uint32_t)(cx.G - 1)) {
      sts32u(cnt_s, 0u);
      if (issue_tma) {
        mbar_expect_tx(&sm.bar[buf], TG::kTileBytes);
        tma_load_3d(sm.base + (buf ? TG::kTileBytes : 0), tmap, 0, tma_c1, job_row, &sm.bar[buf]);
      }
    }
  }

  // ---- 5. gate with silu(z) and store (physical order) --------------------------------------------------------------
  if (!STATE_ONLY && seg_in && cx.active) {
    float zs[TOK], o[TOK];
    v4::load16<T>(cx.pre_cur + 2 * ROW, zs);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      o[2 * j] = y2[j].x * silu_io<T>(zs[2 * j]);
      o[2 * j + 1] = y2[j].y * silu_io<T>(zs[2 * j + 1]);
    }
    T* go = g_o + tseg;
    if (!TAIL || tseg + TOK <= L) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 r;
        T* e = reinterpret_cast<T*>(&r);
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = io<T>::from_f(o[8 * h + k]);
        stg128(go + 8 * h, r);
      }
    } else {
#pragma unroll
      for (int i = 0; i < TOK; ++i)
        if (tseg + i < L) go[i] = io<T>::from_f(o[i]);
    }
  }
}

template <typename T, typename TT, bool REV, bool STATE_ONLY, bool PIPE>
CAD_DEV void run_job(const cad_scan_fwd_args& a, const tmap_t* tmap, int job, int seq, int pset, const Smem& sm) {
  const int lane = CAD_TID & 31, warp = CAD_TID >> 5, G = CAD_NTHREADS >> 5;
  const int64_t L = a.L, E = a.E;
  const int64_t chn = (int64_t)CAD_BIDX * G + warp;
  ChunkCtx cx;
  cx.lane = lane;
  cx.G = G;
  cx.active = chn < E;                          // tail CTA: idle warps keep the arrival counts whole
  const int64_t chc = cx.active ? chn : E - 1;
  const int64_t nchunks = (L + CH - 1) / CH;
  const int64_t pc = (int64_t)pset * E + chc;

  cx.par_s = sm.par + warp * 32;
  cx.a2_s = sm.a2 + warp * (NST * 4);
  cx.carry_s = sm.carry + warp * (NST * 4);
  {
    float v = 0.f;
    if (lane < 4) v = a.conv_w[pc * 4 + lane];
    else if (lane == 4) v = a.conv_b[pc];
    else if (lane == 5) v = a.dt_b[pc];
    else if (lane == 6) v = a.Dskip[pc];
    if (lane < 8) sts32(cx.par_s + 4 * lane, v);
    if (lane < NST) {
      sts32(cx.a2_s + 4 * lane, a.A2[pc * NST + lane]);
      sts32(cx.carry_s + 4 * lane, a.h0 ? a.h0[((int64_t)job * E + chc) * NST + lane] : 0.f);
    }
  }

  // x values preceding logical time 0 (sequence-shard halo): hal[k] = x[tau = k - 3]; see scan_fwd.cu
  float hal[3] = {0.f, 0.f, 0.f};
  if (a.halo) {
    const T* hp = static_cast<const T*>(a.halo) + ((int64_t)job * E + chc) * 3;
    hal[0] = io<T>::to_f(hp[0]); hal[1] = io<T>::to_f(hp[1]); hal[2] = io<T>::to_f(hp[2]);
  }
  const int64_t tau0 = REV ? L - nchunks * CH : 0;     // logical time of the first item of the first visited chunk
  float prev3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t tau = tau0 - 3 + k;
    prev3[k] = tau == -1 ? hal[2] : (tau == -2 ? hal[1] : (tau == -3 ? hal[0] : 0.f));
  }
  float dt_total = 0.f;

  cx.sg = REV ? 31 - lane : lane;
  using TG = tile_geom<TT>;
  piece_offsets<TT>(cx.sg, cx.poff);
  const int job_row = job * 2 * NST;
  constexpr int BPC = CH / TG::kLineTok;        // swizzle lines per chunk row: 8 (16-bit) or 16 (fp32)

  cta_sync();                                   // barrier init + parameter staging visible
  if (CAD_TID == 0) {
    for (int k = 0; k < 2 && k < nchunks; ++k) {
      const int64_t pci = REV ? nchunks - 1 - k : k;
      mbar_expect_tx(&sm.bar[k], TG::kTileBytes);
      tma_load_3d(sm.base + (k ? TG::kTileBytes : 0), tmap, 0, (int)(pci * BPC), job_row, &sm.bar[k]);
    }
  }

  const T* __restrict__ g_x = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + chc) * a.ldxz;
  const T* __restrict__ g_d = static_cast<const T*>(a.delta) + ((int64_t)job * E + chc) * a.ldd;
  T* __restrict__ g_o = static_cast<T*>(a.out) + ((int64_t)job * E + chc) * a.ldo;
  auto pre_addr = [&](int b) {
    return sm.pre + (uint32_t)((((size_t)(b * G + warp) * kRows) * CH + (size_t)cx.sg * TOK) * sizeof(T));
  };
  {
    const int64_t ts0 = (REV ? nchunks - 1 : 0) * CH + (int64_t)cx.sg * TOK;
    if (ts0 < L) {
      constexpr uint32_t ROW = CH * sizeof(T);
      const uint32_t d = pre_addr(0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cp_async16s(d + 0 * ROW + 16 * h, g_x + ts0 + 8 * h);
        cp_async16s(d + 1 * ROW + 16 * h, g_d + ts0 + 8 * h);
        if (!STATE_ONLY) cp_async16s(d + 2 * ROW + 16 * h, g_x + E * a.ldxz + ts0 + 8 * h);
      }
      cp_commit();
    }
  }

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
      chunk<T, TT, REV, true, STATE_ONLY, PIPE>(a, sm, cx, tmap, prev3, hal, dt_total, tseg, tseg_next, stage_next, g_x, g_d, g_o,
                                            buf, parity, issue_tma, tma_c1, job_row);
    else
      chunk<T, TT, REV, false, STATE_ONLY, PIPE>(a, sm, cx, tmap, prev3, hal, dt_total, tseg, tseg_next, stage_next, g_x, g_d, g_o,
                                             buf, parity, issue_tma, tma_c1, job_row);
    if (a.chunk_state) {                        // state at the end of each 512-token logical chunk (saved for backward)
      warp_sync();
      if (cx.active && lane < NST)
        a.chunk_state[(((int64_t)job * E + chn) * nchunks + c) * NST + lane] = lds32(cx.carry_s + 4 * lane);
    }
  }

  warp_sync();
  if (cx.active) {
    if (a.hlast && lane < NST) a.hlast[((int64_t)job * E + chn) * NST + lane] = lds32(cx.carry_s + 4 * lane);
    if (a.dtsum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dt_total += shfl_xor1(dt_total, o);
      if (lane == 0) a.dtsum[(int64_t)job * E + chn] = dt_total;
    }
  }
}

// shared-memory plan (bytes from the 1024-aligned base): two tiles | par | a2 | carry | counters + bars | staging
template <typename TT>
CAD_DEV void carve(unsigned char* base, Smem& sm) {
  constexpr int TB = tile_geom<TT>::kTileBytes;
  const uint32_t b = smem_u32(base);
  sm.base = base;
  sm.tile[0] = b;
  sm.tile[1] = b + TB;
  sm.par = b + 2 * TB;
  sm.a2 = sm.par + kMaxG9 * 32;
  sm.carry = sm.a2 + kMaxG9 * NST * 4;
  sm.cnt = sm.carry + kMaxG9 * NST * 4;
  sm.bar = reinterpret_cast<uint64_t*>(base + 2 * TB + kMaxG9 * 32 + 2 * kMaxG9 * NST * 4 + 16);
  sm.pre = sm.cnt + 16 + 16;
}
inline size_t smem_bytes(int G, size_t elem, size_t tile_elem) {
  return 1024 + (size_t)2 * (2 * NST * CH * tile_elem) + kMaxG9 * 32 + (size_t)2 * kMaxG9 * NST * 4 + 32 +
         (size_t)2 * G * kRows * CH * elem;
}

template <typename T, typename TT, bool STATE_ONLY, bool PIPE>
CAD_DEV void kernel_body(const cad_scan_fwd_args& a, const tmap_t* tmap, unsigned char* smem_raw) {
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Smem sm;
  carve<TT>(base, sm);
  if (CAD_TID == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    sts32u(sm.cnt, 0u);
    sts32u(sm.cnt + 4, 0u);
  }
  const int job = CAD_BIDY;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) run_job<T, TT, true, STATE_ONLY, PIPE>(a, tmap, job, seq, pset, sm);
  else     run_job<T, TT, false, STATE_ONLY, PIPE>(a, tmap, job, seq, pset, sm);
}

}  // namespace v9
}  // namespace cad

// Fused bidirectional selective scan, variant 20: LANE = CHANNEL (sm_100a) — same operator as scan_fwd.cu
// (cad_bimamba_scan_fwd: replaces causal_conv1d_fwd + selective_scan_cuda.fwd behind ref:caduceus/modeling_caduceus.py:128-137,
// ref:caduceus/modeling_rcps.py:85-99), decomposed the other way round.
//
// Why (profiles/r1_scan_pipe_balance.txt): with lane = 16 tokens of ONE channel (variants 3..12) the SM's shared-memory data
// pipe (64.5 %), the issue slots (58 %) and the MUFU pipe (54 %) are loaded within 20 % of each other — every warp re-reads the
// whole B / C tile (they do not depend on the channel), and the time-parallel split needs a 5-round shuffle scan plus a replay
// or carry term per state.  Here
//   * a warp owns 32 CHANNELS and walks time serially; a lane keeps its channel's 16 states and 16 decay rates in registers
//     as 8 packed pairs: per token and pair  FMUL2 dt*A2, 2 x ex2, FMUL2 (dt u)*B, FFMA2 h, FFMA2 y  = 3 issue slots per
//     (token, channel, state) instead of 11.4, 16 independent dependency chains per lane, NO shuffles, no replay / carry term;
//   * B / C come from a TOKEN-major copy of the rows (bcT: 32 fp32 per token) staged by one 1-D bulk copy per 256-token
//     chunk; every lane reads the SAME 16 bytes, so an LDS.128 is one wavefront for the warp: 8 per token for 512 elements
//     instead of 48;
//   * x / dt_raw / z reach each lane through its own cp.async ring (16 bytes = 8 tokens per array and stage), the gated
//     output leaves as one 16-byte store per lane and 8 tokens;
//   * the conv window is three registers.
// What is left is MUFU: 16 ex2 + 4 per (token, channel) = the pipe's floor of ~1.2 ms per Caduceus-PS launch.
//
// The price is that a channel has no time parallelism any more: one GPU must cut the sequence into `nseg` segments per job
// (grid.z), scan each from a ZERO state (this kernel: outputs, end state and sum dt of every segment) and resolve the carries
// afterwards exactly like the multi-GPU path does: cad_seg_carry composes the carry-in of every segment,
// cad_bimamba_scan_fixup adds its decaying contribution in place.  Inference only (no saved chunk states), 16-bit I/O; the
// sharding hooks are served by the same machinery: conv halo here, carry-in / end state / sum dt by cad_seg_carry.  Written
// against the SIMT primitives of simt.cuh so that tests/emu/ compiles THIS file for the host.
#pragma once
#include <type_traits>

#include "simt.cuh"

namespace cad {
namespace v20 {

using simt::atomic_inc_shared; using simt::sts32u; using simt::warp_sync; using simt::fma2; using simt::mul2; using simt::add2;
using simt::splat; using simt::ex2_2; using simt::lds128u; using simt::lds16u; using simt::sts16u; using simt::cp_async16s;
using simt::cp_commit; using simt::cp_wait_group; using simt::bulk_load_1d; using simt::cta_sync; using simt::stg128;
using simt::stg128f; using simt::f2u; using simt::u2f; using simt::warp_all;

constexpr int NST = 16;                 // d_state
constexpr int NPAIR = NST / 2;          // packed state pairs per lane
constexpr int GT = 8;                   // tokens per lane and staging group (16 bytes of 16-bit I/O)
constexpr int CH = 256;                 // tokens per B / C chunk
constexpr int GPC = CH / GT;            // groups per chunk
constexpr int kTileBytes = CH * 2 * NST * 4;      // 32 KB: token-major fp32 (B[0..15], C[0..15]) per token
constexpr int kStages = 4;              // depth of the per-lane cp.async ring (groups of 8 tokens)
constexpr int kAhead = kStages - 2;     // groups in flight ahead of the one being consumed
constexpr int kStageBytes = 3 * 32 * 16;           // x, dt_raw, z: 16 bytes per lane each
constexpr int kMaxW = 8;                // warps per CTA (each 32 channels)

struct Smem {
  uint32_t tile[2];     // the two B / C chunk buffers
  uint32_t ring;        // [W][kStages][3][32 lanes][16 B]
  uint32_t cnt;         // [2] arrival counters of the chunk buffers
  uint64_t* bar;        // full[2]
};
// no swizzled TMA tile here: 128-byte alignment is enough, and at 8 warps the 1 KB slack of the other variants would push two
// CTAs (2 x (dynamic + 1 KB reserved)) 128 bytes over the 228 KB of an SM
inline size_t smem_bytes(int W) { return 128 + (size_t)2 * kTileBytes + (size_t)W * kStages * kStageBytes + 64; }

CAD_DEV void carve(unsigned char* base, int W, Smem& sm) {
  const uint32_t b = smem_u32(base);
  sm.tile[0] = b;
  sm.tile[1] = b + kTileBytes;
  sm.ring = b + 2 * kTileBytes;
  const uint32_t tail = 2 * kTileBytes + (uint32_t)W * kStages * kStageBytes;
  sm.cnt = b + tail;
  sm.bar = reinterpret_cast<uint64_t*>(base + tail + 16);
}

// token range of physical block k of `nseg` (whole 256-token chunks): [lo, hi) with hi <= L; empty when lo >= hi
CAD_DEV void block_range(int64_t L, int nseg, int64_t k, int64_t& lo, int64_t& hi) {
  const int64_t nchunks = (L + CH - 1) / CH;
  const int64_t per = (nchunks + nseg - 1) / nseg;
  lo = k * per * CH;
  hi = (k + 1) * per * CH;
  if (hi > L) hi = L;
  if (lo > hi) lo = hi;
}

template <typename T, bool REV>
CAD_DEV void run_segment(const cad_scan_fwd_args& a, const Smem& sm, int job, int seq, int pset) {
  const int lane = CAD_TID & 31, warp = CAD_TID >> 5, W = CAD_NTHREADS >> 5;
  const int64_t L = a.L, E = a.E;
  const int nseg = a.nseg > 0 ? a.nseg : 1;
  const int sl = CAD_BIDZ;                                   // LOGICAL segment index (carries compose in this order)
  const int64_t k = REV ? nseg - 1 - sl : sl;                 // physical block of tokens
  const int64_t chn = ((int64_t)CAD_BIDX * W + warp) * 32 + lane;
  const bool active = chn < E;
  const int64_t chc = active ? chn : E - 1;
  int64_t t_lo, t_hi;
  block_range(L, nseg, k, t_lo, t_hi);

  float2 h2[NPAIR];
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) h2[p] = make_float2(0.f, 0.f);
  float dsum = 0.f;

  if (t_lo < t_hi) {                                         // CTA-uniform
    const T* __restrict__ xrow = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + chc) * a.ldxz;
    const int64_t zoff = E * a.ldxz;                            // z row of the same channel (uniform offset)
    const T* __restrict__ drow = static_cast<const T*>(a.delta) + ((int64_t)job * E + chc) * a.ldd;
    T* __restrict__ orow = static_cast<T*>(a.out) + ((int64_t)job * E + chc) * a.ldo;
    const int64_t pc = (int64_t)pset * E + chc;
    const float cw0 = a.conv_w[pc * 4 + 0], cw1 = a.conv_w[pc * 4 + 1], cw2 = a.conv_w[pc * 4 + 2], cw3 = a.conv_w[pc * 4 + 3];
    const float cb = a.conv_b[pc], dtb = a.dt_b[pc], Dk = a.Dskip[pc];
    float2 A2p[NPAIR];
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) A2p[p] = make_float2(a.A2[pc * NST + 2 * p], a.A2[pc * NST + 2 * p + 1]);

    // x outside the sequence: the conv halo (x at logical times -3, -2, -1: sequence sharding) before logical time 0, else zero.
    // Read from memory where needed (segment start, masked tail tokens): three registers less in the hot loop.
    const T* hp = a.halo ? static_cast<const T*>(a.halo) + ((int64_t)job * E + chc) * 3 : nullptr;
    auto halo_at = [&](int64_t tau) -> float { return (hp && tau >= -3 && tau <= -1) ? io<T>::to_f(hp[tau + 3]) : 0.f; };
    auto x_at = [&](int64_t t) -> float { return (t >= 0 && t < L) ? io<T>::to_f(xrow[t]) : halo_at(REV ? L - 1 - t : t); };
    // conv window = the three x values that logically precede the FIRST PROCESSED token.  Reversed: that token is the last of
    // the block's last 8-token group, which lies beyond t_hi - 1 when the sequence end is ragged (the masked tokens in between
    // then push their own halo / zero values, so starting the window at t_hi would enter them twice)
    float w0, w1, w2;
    if (REV) { const int64_t tf = ((t_hi + GT - 1) / GT) * GT - 1; w0 = x_at(tf + 3); w1 = x_at(tf + 2); w2 = x_at(tf + 1); }
    else     { w0 = x_at(t_lo - 3); w1 = x_at(t_lo - 2); w2 = x_at(t_lo - 1); }

    // groups of 8 tokens [8 g, 8 g + 8), g in [g_lo, g_hi], walked in logical order (descending when reversed);
    // 32-bit counters: L < 2^31 (checked by the launcher)
    const int g_lo = (int)(t_lo / GT), g_hi = (int)((t_hi + GT - 1) / GT) - 1, ng = g_hi - g_lo + 1;
    constexpr int gstep = REV ? -1 : 1;
    const uint32_t ring = sm.ring + (uint32_t)warp * (kStages * kStageBytes) + (uint32_t)lane * 16;
    // group q + kAhead is staged while group q is consumed; every call commits (uniform cp.async group counting)
    auto stage = [&](int qs, int gs) {
      if (qs < ng) {
        const T* src = xrow + (int64_t)gs * GT;
        const uint32_t s = ring + (uint32_t)(qs & (kStages - 1)) * kStageBytes;
        cp_async16s(s, src);
        cp_async16s(s + 512, drow + (int64_t)gs * GT);
        cp_async16s(s + 1024, src + zoff);
      }
      cp_commit();
    };
    // chunks of this block, logical order
    const int c_lo = (int)(t_lo / CH), c_hi = (int)((t_hi + CH - 1) / CH) - 1, nc = c_hi - c_lo + 1;
    const int64_t nchunks = (L + CH - 1) / CH;
    const float* __restrict__ bct = a.bcT + (int64_t)job * nchunks * CH * 2 * NST;
    auto issue_chunk = [&](int qc, int buf) {
      const int c = REV ? c_hi - qc : c_lo + qc;
      mbar_expect_tx(&sm.bar[buf], kTileBytes);
      bulk_load_1d(sm.tile[0] + (uint32_t)buf * kTileBytes, bct + (int64_t)c * CH * 2 * NST, kTileBytes, &sm.bar[buf]);
    };
    if (CAD_TID == 0) {
      issue_chunk(0, 0);
      if (nc > 1) issue_chunk(1, 1);
    }
#pragma unroll
    for (int k2 = 0; k2 < kAhead; ++k2) stage(k2, (REV ? g_hi : g_lo) + gstep * k2);

    // one group of 8 tokens: conv + SiLU, dt, the 16 state recurrences as 8 packed pairs, gate, store
    auto group = [&](auto tail_tag, int g, uint32_t stage_s, uint32_t trow) {
      constexpr bool TAIL = decltype(tail_tag)::value;        // the last physical group of the sequence: tokens >= L masked
      // dt_raw of the 8 tokens at once (8 independent softplus chains); x and z are re-read per token as 16-bit elements from the
      // staged rows instead of being held in 8 registers across the group (the kernel sits at the 128-register cap of two CTAs per
      // SM: holding them spilled loop counters, 1.56 instead of 1.36 ms per Caduceus-PS launch).  Lanes 8 apart share a bank, so such
      // a read is 4 wavefronts; reading token PAIRS as words halves that but measured slower (1.42 ms: longer chain per token).
      const uint4 dq = lds128u(stage_s + 512);
      const T* de = reinterpret_cast<const T*>(&dq);
      auto el16 = [](uint32_t bits) -> float {
        if constexpr (std::is_same<T, __nv_bfloat16>::value) return u2f(bits << 16);
        else return __half2float(__ushort_as_half((unsigned short)bits));
      };
      const int64_t t0 = (int64_t)g * GT;
      auto to16 = [](float v) -> uint32_t {
        if constexpr (std::is_same<T, __nv_bfloat16>::value) return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
        else return (uint32_t)__half_as_ushort(__float2half_rn(v));
      };
      // dt = softplus(dt_raw + b) of the group's tokens (physical order): w = e^v by MUFU for all eight, then log1p(w).  When every
      // lane of the warp has all eight w below 1/2 (dt < 0.405: four times the upper end of Mamba's init-time dt range) the log comes
      // from a degree-7 minimax polynomial of log1p(w) / w on [0, 1/2] on the FMA pipe (max relative error 1e-7 in fp32 Horner form)
      // instead of a second MUFU op — one of the 20 MUFU ops per token and channel of this MUFU-bound kernel; otherwise common.cuh's
      // form (lg2, series below 2^-6, threshold 20).
      float dv[GT];
      bool small = true;
#pragma unroll
      for (int j = 0; j < GT; ++j) {
        dv[j] = io<T>::to_f(de[j]) + dtb;
        const float w = ex2(kLog2e * dv[j]);
        small = small && (w < 0.5f);
        dv[j] = small ? w : dv[j];                            // keep w while the fast path is still possible
      }
      if (warp_all(small)) {
#pragma unroll
        for (int j = 0; j < GT; ++j) {
          const float w = dv[j];
          float p = fmaf(w, -0.0274653025f, 0.0867877156f);
          p = fmaf(p, w, -0.146816134f);
          p = fmaf(p, w, 0.195812061f);
          p = fmaf(p, w, -0.249502212f);
          p = fmaf(p, w, 0.33330366f);
          p = fmaf(p, w, -0.499999315f);
          p = fmaf(p, w, 1.0f);
          dv[j] = w * p;
        }
      } else {
#pragma unroll
        for (int j = 0; j < GT; ++j) dv[j] = softplus(io<T>::to_f(de[j]) + dtb);
      }
#pragma unroll
      for (int i = 0; i < GT; ++i) {
        const int pi = REV ? GT - 1 - i : i;                  // physical position inside the group
        const bool masked = TAIL && (t0 + pi >= L);
        // a masked token of a reversed job lies logically BEFORE the sequence: it feeds the conv window with the halo
        const float xv = masked ? (REV ? halo_at(L - 1 - (t0 + pi)) : 0.f) : el16(lds16u(stage_s + 2 * pi));
        const float cv = cb + cw0 * w0 + cw1 * w1 + cw2 * w2 + cw3 * xv;
        w0 = w1; w1 = w2; w2 = xv;
        const float u = silu_io<T>(cv);
        const float d = masked ? 0.f : dv[pi];                // masked: a = 1, b = 0, the state passes through
        dsum += d;
        const float2 dt2 = splat(d), du2 = splat(d * u);
        float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
        const uint32_t tok = trow + (uint32_t)pi * (2 * NST * 4);
#pragma unroll
        for (int p4 = 0; p4 < NPAIR / 2; ++p4) {               // 4 states per 16-byte piece: two packed pairs
          const float4 bq = lds128(tok + 16 * p4), cq = lds128(tok + NST * 4 + 16 * p4);
          const float2 a0 = ex2_2(mul2(dt2, A2p[2 * p4]));
          const float2 a1 = ex2_2(mul2(dt2, A2p[2 * p4 + 1]));
          h2[2 * p4] = fma2(a0, h2[2 * p4], mul2(du2, make_float2(bq.x, bq.y)));
          h2[2 * p4 + 1] = fma2(a1, h2[2 * p4 + 1], mul2(du2, make_float2(bq.z, bq.w)));
          ya = fma2(make_float2(cq.x, cq.y), h2[2 * p4], ya);
          yb = fma2(make_float2(cq.z, cq.w), h2[2 * p4 + 1], yb);
        }
        const float y = fmaf(Dk, u, (ya.x + ya.y) + (yb.x + yb.y));
        // the gated output replaces z in this lane's own 16 bytes of the stage (no output registers held across the group);
        // one LDS.128 + STG.128 per group below
        sts16u(stage_s + 1024 + 2 * pi, to16(y * silu_io<T>(el16(lds16u(stage_s + 1024 + 2 * pi)))));
      }
      if (active) {
        const uint4 oq = lds128u(stage_s + 1024);
        if (!TAIL) stg128(orow + t0, oq);
        else {
          const T* oe = reinterpret_cast<const T*>(&oq);
#pragma unroll
          for (int i = 0; i < GT; ++i)
            if (t0 + i < L) orow[t0 + i] = oe[i];
        }
      }
    };

    const int g_tail = (L % GT) ? (int)(L / GT) : -1;           // the one group with masked tokens (ragged sequence end)
    uint32_t par = 0;                                          // bit b: parity to wait for on chunk buffer b
    int q = 0, g = REV ? g_hi : g_lo;                         // logical group counter, physical group
#pragma unroll 1
    for (int qc = 0; qc < nc; ++qc) {
      const int buf = qc & 1;
      const int c = REV ? c_hi - qc : c_lo + qc;
      mbar_wait_wd(&sm.bar[buf], (par >> buf) & 1u);
      par ^= 1u << buf;
      const uint32_t tile_s = sm.tile[0] + (uint32_t)buf * kTileBytes;
      // groups of this chunk that belong to the block
      const int gc_lo = c * GPC > g_lo ? c * GPC : g_lo, gc_hi = (c * GPC + GPC - 1) < g_hi ? (c * GPC + GPC - 1) : g_hi;
#pragma unroll 1
      for (int gi = gc_hi - gc_lo; gi >= 0; --gi, ++q, g += gstep) {
        stage(q + kAhead, g + gstep * kAhead);
        cp_wait_group<kAhead>();
        const uint32_t stage_s = ring + (uint32_t)(q & (kStages - 1)) * kStageBytes;
        const uint32_t trow = tile_s + (uint32_t)(g * GT - c * CH) * (2 * NST * 4);
        if (g == g_tail) group(std::true_type{}, g, stage_s, trow);
        else group(std::false_type{}, g, stage_s, trow);
      }
      // release the chunk buffer; the LAST warp to arrive requests chunk qc + 2 into it
      warp_sync();
      if (lane == 0) {
        const uint32_t cnt_s = sm.cnt + 4 * buf;
        if (atomic_inc_shared(cnt_s) == (uint32_t)(W - 1)) {
          sts32u(cnt_s, 0u);
          if (qc + 2 < nc) issue_chunk(qc + 2, buf);
        }
      }
    }
    cp_wait_group<0>();
  }

  // end state and sum dt of the segment scanned from zero (identity for an empty block)
  if (active && a.seg_state) {
    float* st = a.seg_state + (((int64_t)job * nseg + sl) * E + chn) * NST;
#pragma unroll
    for (int p = 0; p < NPAIR; p += 2) stg128f(st + 2 * p, h2[p].x, h2[p].y, h2[p + 1].x, h2[p + 1].y);
    a.seg_dtsum[((int64_t)job * nseg + sl) * E + chn] = dsum;
  }
}

template <typename T>
CAD_DEV void kernel_body(const cad_scan_fwd_args& a, unsigned char* smem_raw) {
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  Smem sm;
  carve(base, CAD_NTHREADS >> 5, sm);
  if (CAD_TID == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    sts32u(sm.cnt, 0u);
    sts32u(sm.cnt + 4, 0u);
  }
  cta_sync();
  const int job = CAD_BIDY;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) run_segment<T, true>(a, sm, job, seq, pset);
  else     run_segment<T, false>(a, sm, job, seq, pset);
}

}  // namespace v20
}  // namespace cad

// Fused bidirectional selective-scan forward for sm_100a (TMA-staged B/C tile, cp.async-staged x/dt/z, 2 CTAs/SM).
//
// One launch covers every (sequence, direction) "job" of a BiMamba call — forward and reverse directions of
// ref:caduceus/modeling_caduceus.py:128-137 and both strands of ref:caduceus/modeling_rcps.py:85-99 — with the
// reversed jobs handled purely by addressing (no flipped copies).  Per job and channel it fuses what the
// reference runs as separate upstream kernels:
//     depthwise (anti)causal conv + bias + SiLU                      (causal_conv1d_fwd)
//     dt = softplus(dt_raw + b_dt)                                   (prologue of selective_scan_fwd)
//     h_t = exp2(dt*A2) h_{t-1} + dt*B_t*u_t ;  y_t = C_t . h_t + D u_t ;  out = y * silu(z)
//
// Work decomposition (B200: 148 SMs; the scan is MUFU/issue-bound, not HBM-bound — SURVEY.md §8d):
//   CTA  = one job x G <= 7 consecutive channels, one WARP PER CHANNEL, walking the sequence in logical-time
//          chunks of 512 tokens; 2 CTAs are resident per SM (128 registers, 109 KB smem), so Caduceus-PS
//          (4 jobs x 74 CTAs) is a single wave of 296 CTAs and each SM holds 14 warps — which is ALL the channel
//          parallelism there is at batch 1 (2048 channel-jobs / 148 SMs): more warps per SM would need more
//          than one warp per channel (tried, no gain: DESIGN.md §9).
//   lane = 16 consecutive tokens of the chunk.  Each lane runs the recurrence over its 16 tokens from a zero
//          state, the 32 segment aggregates (prod a, h_end) are combined with a 5-step warp-shuffle scan, and
//          the lane re-runs its 16 FMAs from the true incoming state.  exp2 is evaluated ONCE per
//          (token, channel, state): the segment decay is exp2(A2 * sum(dt)), not a product.
//   smem = the chunk's 2N x 512 fp32 tile of B / C rows shared by the G channels, fetched by ONE TMA
//          (cp.async.bulk.tensor.3d, SWIZZLE_128B) per chunk into a bank-conflict-free layout; tokens beyond
//          the sequence end are zero-filled by the TMA unit.  The request for chunk c+1 is issued before the
//          gate/store epilogue of chunk c, so its latency hides behind the epilogue and the next prologue.
//          The lane's own x / dt_raw / z segments of chunk c+1 are staged by cp.async during chunk c (16-bit I/O).
//   template knobs: TOK tokens per lane (16; 8 = 256-token chunks for many-CTA workloads), STATE_ONLY (end state and
//   sum(dt) only, no output), REV/TAIL specialisations (reversed jobs; the one chunk that straddles the sequence end).
//   profiles: profiles/r1_v1_scan_ncu_summary.txt (v1: 20.8 warp-instructions per element, 43 % issue utilisation)
//   -> profiles/r1_v3_scan_and_bwd_ncu_summary.txt (11.4 instructions per element, MUFU pipe 53 %, issue 57 %).
#include "scan_common.cuh"

namespace cad {

struct ScanSmem {
  float* tile;       // 2N rows x 512 tokens fp32, TMA-swizzled; 1024-byte aligned
  float* carry;      // kMaxG x N   running state of each warp's channel
  float* a2;         // kMaxG x N   A2 of each warp's channel
  uint64_t* bar;     // TMA completion barrier
  unsigned char* pre; // [2][G][3][512] x / dt_raw / z segments of the current and the next chunk (16-bit I/O only)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One chunk of one channel.  TAIL: the chunk straddles the sequence end (per-token masks, halo in the pad).
template <typename T, int N, int TOK, bool REV, bool TAIL, bool STATE_ONLY>
__device__ __forceinline__ void scan_chunk(
    const cad_scan_fwd_args& a, const ScanSmem& sm, int lane, int seg, const uint32_t (&poff)[TOK / 4],
    const T* __restrict__ xrow, const T* __restrict__ zrow, const T* __restrict__ drow, T* __restrict__ orow,
    int64_t tseg, bool active, const float (&cw)[4], float cb, float dtb, float Dk, const float (&hal)[3],
    float (&prev3)[3], float& dt_total, float* my_carry, const float* my_a2, uint32_t parity, bool issue_next,
    const CUtensorMap* tmap, int next_c1, int job_row, const T* pre_cur, T* pre_next, int64_t tseg_next) {
  constexpr int CH = 32 * TOK;         // tokens per chunk
  constexpr int EPV = 16 / sizeof(T);
  const int64_t L = a.L;
  auto phys = [](int i) { return REV ? TOK - 1 - i : i; };
  auto halo_at = [&](int64_t tau) { return tau == -1 ? hal[2] : (tau == -2 ? hal[1] : (tau == -3 ? hal[0] : 0.f)); };
  const bool seg_in = !TAIL || tseg < L;

  // ---- 1. x and dt_raw segments (global -> registers) ------------------------------------------------
  constexpr bool PRE = sizeof(T) == 2;     // 16-bit I/O: segments were staged by cp.async one chunk ahead
  float xs[TOK], dr[TOK];
  if (seg_in) {
    if (PRE) {
      cp_async_wait_all();                 // my own copies (only this lane reads what it staged)
      load_vec_smem<T, TOK>(pre_cur, xs);
      load_vec_smem<T, TOK>(pre_cur + CH, dr);
    } else {
      load_vec<T, TOK>(xrow + tseg, xs);
      load_vec<T, TOK>(drow + tseg, dr);
    }
  } else {
#pragma unroll
    for (int i = 0; i < TOK; ++i) { xs[i] = 0.f; dr[i] = 0.f; }
  }
  if (PRE && issue_next && tseg_next < L) {   // stream in the next chunk's x / dt_raw / z behind this chunk's math
#pragma unroll
    for (int v = 0; v < TOK / 8; ++v) {
      cp_async16(pre_next + 8 * v, xrow + tseg_next + 8 * v);
      cp_async16(pre_next + CH + 8 * v, drow + tseg_next + 8 * v);
      if (!STATE_ONLY) cp_async16(pre_next + 2 * CH + 8 * v, zrow + tseg_next + 8 * v);
    }
    cp_async_commit();
  }

  // ---- 2. per-(token, channel) prologue: conv + SiLU, dt, dt*u (independent of the tile) -------------
  float dt[TOK], du[TOK], y[TOK];
  float dsum = 0.f;
  {
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
    for (int k = 0; k < 3; ++k) {        // logical predecessors: previous lane / previous chunk
      const float up = __shfl_up_sync(0xffffffffu, xl[TOK + k], 1);
      xl[k] = (lane == 0) ? prev3[k] : up;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) prev3[k] = __shfl_sync(0xffffffffu, xl[TOK + k], 31);
#pragma unroll
    for (int i = 0; i < TOK; ++i) {
      const float u = silu_io<T>(cb + cw[0] * xl[i] + cw[1] * xl[i + 1] + cw[2] * xl[i + 2] + cw[3] * xl[i + 3]);
      float d = softplus(dr[phys(i)] + dtb);
      if (TAIL && tseg + phys(i) >= L) d = 0.f;     // masked token: a = 1, b = 0 -> state passes through
      dt[i] = d;
      dsum += d;
      du[i] = d * u;
      y[i] = Dk * u;
    }
  }
  dt_total += dsum;

  // ---- 3. the scan, one state at a time, on the TMA-staged B/C tile ------------------------------------
  mbar_wait(sm.bar, parity);
  const uint32_t tile_s = smem_u32(sm.tile);
  const uint32_t a2_s = smem_u32(my_a2), carry_s = smem_u32(my_carry);
#pragma unroll 1
  for (int n = 0; n < N; ++n) {
    const float A2n = lds32(a2_s + 4 * n);
    const float cin = lds32(carry_s + 4 * n);
    float av[TOK], bv[TOK];
    float hl = (lane == 0) ? cin : 0.f;
    {
      const uint32_t rowp = tile_s + n * (CH * 4);
#pragma unroll
      for (int k = 0; k < TOK / 4; ++k) {
        const float4 q = lds128(rowp + poff[k]);
        const float bq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = REV ? TOK - 1 - (4 * k + e) : 4 * k + e;     // logical item of physical token 4k+e
          av[i] = ex2(dt[i] * A2n);
          bv[i] = du[i] * bq[e];
        }
      }
#pragma unroll
      for (int i = 0; i < TOK; ++i) hl = fmaf(av[i], hl, bv[i]);
    }
    float P = ex2(A2n * dsum);
    scan_step_up<1>(P, hl, lane);
    scan_step_up<2>(P, hl, lane);
    scan_step_up<4>(P, hl, lane);
    scan_step_up<8>(P, hl, lane);
    scan_step_up<16>(P, hl, lane);
    float h = __shfl_up_sync(0xffffffffu, hl, 1);
    if (lane == 0) h = cin;
    if (lane == 31) sts32(carry_s + 4 * n, hl);          // state at the end of this chunk
    if (!STATE_ONLY) {
      const uint32_t rowp = tile_s + (N + n) * (CH * 4);
      float4 cq[TOK / 4];
#pragma unroll
      for (int k = 0; k < TOK / 4; ++k) cq[k] = lds128(rowp + poff[k]);
      // walk the segment in LOGICAL order (physical pieces backwards for a reversed job)
#pragma unroll
      for (int kk = 0; kk < TOK / 4; ++kk) {
        const int k = REV ? TOK / 4 - 1 - kk : kk;
        const float ce[4] = {cq[k].x, cq[k].y, cq[k].z, cq[k].w};
#pragma unroll
        for (int ee = 0; ee < 4; ++ee) {
          const int e = REV ? 3 - ee : ee;
          const int i = REV ? TOK - 1 - (4 * k + e) : 4 * k + e;
          h = fmaf(av[i], h, bv[i]);
          y[i] = fmaf(ce[e], h, y[i]);
        }
      }
    }
  }

  // ---- 4. hand the tile back: everyone is done reading -> request the next chunk -------------------------
  __syncthreads();
  if (issue_next && threadIdx.x == 0) {
    mbar_expect_tx(sm.bar, 2 * N * CH * 4);
    tma_load_3d(sm.tile, tmap, 0, next_c1, job_row, sm.bar);
  }

  // ---- 5. gate with silu(z) and store (physical order) ---------------------------------------------------
  if (!STATE_ONLY && seg_in && active) {
    float zs[TOK], o[TOK];
    if (PRE) load_vec_smem<T, TOK>(pre_cur + 2 * CH, zs);
    else load_vec<T, TOK>(zrow + tseg, zs);
#pragma unroll
    for (int i = 0; i < TOK; ++i) o[phys(i)] = y[i] * silu_io<T>(zs[phys(i)]);
    if (!TAIL || tseg + TOK <= L) {
      store_vec<T, TOK>(orow + tseg, o);
    } else {
#pragma unroll
      for (int i = 0; i < TOK; ++i)
        if (tseg + i < L) orow[tseg + i] = io<T>::from_f(o[i]);
    }
  }
  (void)EPV;
}

template <typename T, int N, int TOK, bool REV, bool STATE_ONLY>
__device__ __forceinline__ void scan_job(const cad_scan_fwd_args& a, const CUtensorMap* tmap, int job, int seq,
                                         int pset, const ScanSmem& sm) {
  constexpr int CH = 32 * TOK;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = blockDim.x >> 5;
  const int64_t L = a.L, E = a.E;
  const int64_t ch = (int64_t)blockIdx.x * G + warp;
  const bool active = ch < E;                  // tail CTA: idle warps only keep the barriers company
  const int64_t chc = active ? ch : E - 1;
  const int64_t nchunks = (L + CH - 1) / CH;

  const T* __restrict__ xrow = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + chc) * a.ldxz;
  const T* __restrict__ zrow = xrow + E * a.ldxz;
  const T* __restrict__ drow = static_cast<const T*>(a.delta) + ((int64_t)job * E + chc) * a.ldd;
  T* __restrict__ orow = static_cast<T*>(a.out) + ((int64_t)job * E + chc) * a.ldo;

  const int64_t pc = (int64_t)pset * E + chc;
  const float cw[4] = {a.conv_w[pc * 4 + 0], a.conv_w[pc * 4 + 1], a.conv_w[pc * 4 + 2], a.conv_w[pc * 4 + 3]};
  const float cb = a.conv_b[pc], dtb = a.dt_b[pc], Dk = a.Dskip[pc];
  float* my_carry = sm.carry + warp * N;
  float* my_a2 = sm.a2 + warp * N;
  if (lane < N) {
    my_a2[lane] = a.A2[pc * N + lane];
    my_carry[lane] = a.h0 ? a.h0[((int64_t)job * E + chc) * N + lane] : 0.f;
  }

  // x values preceding logical time 0 (sequence-shard halo): hal[k] = x[tau = k - 3].  Chunks are aligned in
  // PHYSICAL time, so for a reversed job the first logical chunk may start with masked tokens (t >= L, i.e.
  // tau < 0): the halo lives exactly there, and further back everything is zero.
  float hal[3] = {0.f, 0.f, 0.f};
  if (a.halo) {
    const T* hp = static_cast<const T*>(a.halo) + ((int64_t)job * E + chc) * 3;
    hal[0] = io<T>::to_f(hp[0]); hal[1] = io<T>::to_f(hp[1]); hal[2] = io<T>::to_f(hp[2]);
  }
  const int64_t tau0 = REV ? L - nchunks * CH : 0;      // logical time of the first item of chunk 0 (<= 0)
  float prev3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t tau = tau0 - 3 + k;
    prev3[k] = tau == -1 ? hal[2] : (tau == -2 ? hal[1] : (tau == -3 ? hal[0] : 0.f));
  }
  float dt_total = 0.f;

  // this lane's segment of the chunk, and the byte offsets of its four 16-byte pieces inside a tile row:
  // TMA SWIZZLE_128B stores 16-byte chunk c of 128-byte line l at chunk position c ^ (l & 7); a tile row is 16
  // consecutive lines (line index = row*16 + blk, so l & 7 == blk & 7).
  const int seg = REV ? 31 - lane : lane;
  uint32_t poff[TOK / 4];
  tile_piece_offsets<TOK>(seg, poff);
  const int job_row = job * 2 * N;
  const int blocks_per_chunk = CH / kBlkTok;

  __syncthreads();                               // barrier init + parameter staging visible
  if (threadIdx.x == 0) {
    const int64_t first = REV ? nchunks - 1 : 0;
    mbar_expect_tx(sm.bar, 2 * N * CH * 4);
    tma_load_3d(sm.tile, tmap, 0, (int)(first * blocks_per_chunk), job_row, sm.bar);
  }

  // 16-bit I/O: this lane's x / dt_raw / z segments are staged in smem one chunk ahead (cp.async, no registers held)
  constexpr bool PRE = sizeof(T) == 2;
  T* pre_base = reinterpret_cast<T*>(sm.pre);
  auto pre_ptr = [&](int buf) { return pre_base + ((size_t)(buf * G + warp) * 3) * CH + seg * TOK; };
  if (PRE && nchunks > 0) {
    const int64_t ts0 = (REV ? nchunks - 1 : 0) * CH + (int64_t)seg * TOK;
    if (ts0 < L) {
      T* d = pre_ptr(0);
#pragma unroll
      for (int v = 0; v < TOK / 8; ++v) {
        cp_async16(d + 8 * v, xrow + ts0 + 8 * v);
        cp_async16(d + CH + 8 * v, drow + ts0 + 8 * v);
        if (!STATE_ONLY) cp_async16(d + 2 * CH + 8 * v, zrow + ts0 + 8 * v);
      }
      cp_async_commit();
    }
  }

  uint32_t parity = 0;
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t pcidx = REV ? nchunks - 1 - c : c;
    const int64_t tseg = pcidx * CH + (int64_t)seg * TOK;
    const bool issue_next = c + 1 < nchunks;
    const int next_c1 = (int)((REV ? pcidx - 1 : pcidx + 1) * blocks_per_chunk);
    const bool tail = (pcidx + 1) * CH > L;
    const int64_t tseg_next = (REV ? pcidx - 1 : pcidx + 1) * CH + (int64_t)seg * TOK;
    const T* pre_cur = pre_ptr((int)(c & 1));
    T* pre_next = pre_ptr((int)((c + 1) & 1));
    const float dt_before = dt_total;
    if (tail)
      scan_chunk<T, N, TOK, REV, true, STATE_ONLY>(a, sm, lane, seg, poff, xrow, zrow, drow, orow, tseg, active, cw, cb, dtb, Dk, hal,
                                  prev3, dt_total, my_carry, my_a2, parity, issue_next, tmap, next_c1, job_row, pre_cur, pre_next,
                                  tseg_next);
    else
      scan_chunk<T, N, TOK, REV, false, STATE_ONLY>(a, sm, lane, seg, poff, xrow, zrow, drow, orow, tseg, active, cw, cb, dtb, Dk, hal,
                                   prev3, dt_total, my_carry, my_a2, parity, issue_next, tmap, next_c1, job_row, pre_cur, pre_next,
                                   tseg_next);
    parity ^= 1;
    if (a.chunk_state) {
      __syncwarp();
      if (active && lane < N) a.chunk_state[(((int64_t)job * E + ch) * nchunks + c) * N + lane] = my_carry[lane];
    }
    if (a.chunk_dtsum) {                         // sum of dt over this logical chunk (segment-parallel carry fix-up of a shard)
      float cs = dt_total - dt_before;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
      if (active && lane == 0) a.chunk_dtsum[((int64_t)job * E + ch) * nchunks + c] = cs;
    }
  }

  __syncwarp();
  if (active) {
    if (a.hlast && lane < N) a.hlast[((int64_t)job * E + ch) * N + lane] = my_carry[lane];
    if (a.dtsum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dt_total += __shfl_xor_sync(0xffffffffu, dt_total, o);
      if (lane == 0) a.dtsum[(int64_t)job * E + ch] = dt_total;
    }
  }
}

template <typename T, int N, int TOK, bool STATE_ONLY>
__device__ __forceinline__ void scan_kernel_body(const cad_scan_fwd_args& a, const CUtensorMap* tmap) {
  extern __shared__ unsigned char smem_raw[];
  // the swizzled TMA destination must be 1024-byte aligned
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  ScanSmem sm;
  sm.tile = reinterpret_cast<float*>(base);
  constexpr int CH = 32 * TOK;
  sm.carry = reinterpret_cast<float*>(base + (size_t)2 * N * CH * 4);
  sm.a2 = sm.carry + kMaxG * N;
  sm.bar = reinterpret_cast<uint64_t*>(sm.a2 + kMaxG * N);
  sm.pre = reinterpret_cast<unsigned char*>(sm.bar + 2);
  if (threadIdx.x == 0) mbar_init(sm.bar, 1);
  const int job = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) scan_job<T, N, TOK, true, STATE_ONLY>(a, tmap, job, seq, pset, sm);
  else     scan_job<T, N, TOK, false, STATE_ONLY>(a, tmap, job, seq, pset, sm);
}

template <typename T, int N, int TOK, bool STATE_ONLY>
__global__ void __launch_bounds__(kMaxG * 32, (TOK == 16 ? 2 : 4))
bimamba_scan_fwd_kernel(const cad_scan_fwd_args a, const __grid_constant__ CUtensorMap tmap) {
  scan_kernel_body<T, N, TOK, STATE_ONLY>(a, &tmap);
}

template <typename T, int N, int TOK>
static int launch_scan(const cad_scan_fwd_args& a, int G, cudaStream_t stream) {
  constexpr int CH = 32 * TOK;
  CUtensorMap tmap;
  if (make_row_tile_map(&tmap, a.bc, (int64_t)a.njobs * 2 * N, a.ldbc, a.L, 2 * N, CH) != 0) return -1;

  const size_t pre_bytes = sizeof(T) == 2 ? (size_t)2 * G * 3 * CH * sizeof(T) : 0;
  const size_t smem = 1024 + (size_t)2 * N * CH * 4 + (size_t)2 * kMaxG * N * sizeof(float) + 16 + pre_bytes;
  auto kern = a.state_only ? bimamba_scan_fwd_kernel<T, N, TOK, true> : bimamba_scan_fwd_kernel<T, N, TOK, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)((a.E + G - 1) / G), (unsigned)a.njobs);
  kern<<<grid, G * 32, smem, stream>>>(a, tmap);
  CAD_LAUNCH_CHECK();
  return 0;
}

int launch_scan_v20(const cad_scan_fwd_args& a, cudaStream_t stream);    // scan_fwd_v20.cu (lane = channel)

}  // namespace cad

extern "C" int cad_scan_chunk_len(void) { return cad::kChunk; }

extern "C" int cad_bimamba_scan_fwd(const cad_scan_fwd_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_bimamba_scan_fwd: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_bimamba_scan_fwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->delta && (a->bc || a->variant == 20) && (a->out || a->state_only) && a->conv_w && a->conv_b && a->dt_b &&
              a->A2 && a->Dskip && a->seq_of_job && a->pset_of_job && a->rev_of_job, "cad_bimamba_scan_fwd: null pointer");
  CAD_REQUIRE(a->N == 16, "cad_bimamba_scan_fwd: d_state = %lld not built (only 16)", (long long)a->N);
  CAD_REQUIRE(a->K >= 1 && a->K <= 4, "cad_bimamba_scan_fwd: d_conv = %lld out of range [1, 4]", (long long)a->K);
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldd % 16 == 0 && a->ldo % 16 == 0 && a->ldxz >= a->L && a->ldd >= a->L &&
              a->ldo >= a->L, "cad_bimamba_scan_fwd: row pitches must be multiples of 16 elements and >= L");
  CAD_REQUIRE(a->ldbc % 32 == 0 && a->ldbc >= a->L, "cad_bimamba_scan_fwd: ldbc must be a multiple of 32 and >= L");
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->delta) && aligned16(a->bc) && aligned16(a->out),
              "cad_bimamba_scan_fwd: xz/delta/bc/out must be 16-byte aligned");
  CAD_REQUIRE(!a->state_only || (a->hlast && a->dtsum), "cad_bimamba_scan_fwd: state_only needs hlast and dtsum");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CAD_REQUIRE(a->variant == 0 || a->variant == 3 || a->variant == 20, "cad_bimamba_scan_fwd: variant must be 0, 3 or 20");
  if (a->variant == 20) return launch_scan_v20(*a, stream);
  int G = a->channels_per_cta;
  if (G <= 0) {
    // two CTAs are resident per SM; the busiest SM carries ceil(ctas / sms) * g channels: minimise that
    const int sms = cad_sm_count() > 0 ? cad_sm_count() : 148;
    long best = -1;
    for (int g = 1; g <= kMaxG; ++g) {
      const long ctas = (long)a->njobs * ((a->E + g - 1) / g);
      const long cost = ((ctas + sms - 1) / sms) * g;
      if (best < 0 || cost <= best) { best = cost; G = g; }
    }
  }
  CAD_REQUIRE(G >= 1 && G <= kMaxG, "cad_bimamba_scan_fwd: channels_per_cta must be in [1, %d]", kMaxG);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, return launch_scan<T, 16, 16>(*a, G, stream));
  return 0;
}

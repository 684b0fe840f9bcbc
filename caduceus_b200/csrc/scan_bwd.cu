// Fused bidirectional selective-scan BACKWARD for sm_100a.
//
// Replaces selective_scan_cuda.bwd (upstream; the backward of `MambaInnerFn`, SURVEY.md row A16, reached from the
// reference through autograd of ref:caduceus/modeling_caduceus.py:128-137).  Same job / channel / chunk geometry
// as scan_fwd.cu; logical chunks are walked in REVERSE.  Per chunk and channel the kernel
//   1. recomputes the forward inside the chunk from the state saved at the chunk boundary (fwd `chunk_state`):
//      u = silu(conv(x)), dt = softplus(dt_raw + b), a = exp2(dt*A2), h_t  (same one-exp2-per-element scan),
//   2. runs the adjoint recurrence  e_t = C_t*dy_t + a_{t+1} e_{t+1}  as a suffix scan (shfl.down) with the same
//      two-pass trick (segment decay = exp2(A2 * sum dt), no second set of exp2),
//   3. accumulates   d dt, d u, dA2, dD, d b_dt  per channel and  dB, dC  across the CTA's channels
//      (per-state double-buffered smem slots, then red.global.add.v4.f32 into the fp32 dB/dC rows).
// Outputs: dz, d dt_raw, du (gradient w.r.t. the conv+SiLU output, finished by conv_bwd after the x_proj GEMM
// gradient has been added), dB/dC, parameter gradients, optional dh0.
#include "scan_common.cuh"

namespace cad {

constexpr float kLn2f = 0.6931471805599453f;

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// sigmoid via MUFU
__device__ __forceinline__ float sigmoidf_fast(float v) { return rcp(1.0f + ex2(-kLog2e * v)); }

struct BwdSmem {
  float* tile;       // 2N x 512 fp32 B/C rows (TMA)
  float* slots;      // [2][G][2][512] dB/dC contributions of the current state, per warp
  float* cin;        // G x N   state at the START of the current chunk
  float* ecar;       // G x N   adjoint state at the first token of the NEXT logical chunk
  float* a2;         // G x N
  float* dA2acc;     // G x N
  uint64_t* bar;
};

// slot addressing: 16-byte pieces XOR-swizzled inside each lane's 64-byte segment (conflict-free 128-bit access)
__device__ __forceinline__ int slot_piece(int seg, int k) { return seg * 4 + (k ^ ((seg >> 1) & 3)); }

template <typename T, int N, bool REV>
__device__ __forceinline__ void scan_bwd_job(const cad_scan_bwd_args& a, const CUtensorMap* tmap, int job, int seq,
                                             int pset, const BwdSmem& sm) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = blockDim.x >> 5;
  const int64_t L = a.L, E = a.E;
  const int64_t ch = (int64_t)blockIdx.x * G + warp;
  const bool active = ch < E;
  const int64_t chc = active ? ch : E - 1;
  const int64_t nchunks = (L + kChunk - 1) / kChunk;
  auto phys = [](int i) { return REV ? kTok - 1 - i : i; };

  const T* __restrict__ xrow = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + chc) * a.ldxz;
  const T* __restrict__ zrow = xrow + E * a.ldxz;
  const T* __restrict__ drow = static_cast<const T*>(a.delta) + ((int64_t)job * E + chc) * a.ldd;
  const T* __restrict__ gorow = static_cast<const T*>(a.dout) + ((int64_t)job * E + chc) * a.ldo;
  T* __restrict__ dzrow = static_cast<T*>(a.dz) + ((int64_t)job * E + chc) * a.lddz;
  T* __restrict__ durow = static_cast<T*>(a.du) + ((int64_t)job * E + chc) * a.lddu;
  T* __restrict__ ddrow = static_cast<T*>(a.ddelta) + ((int64_t)job * E + chc) * a.lddd;

  const int64_t pc = (int64_t)pset * E + chc;
  const float cw[4] = {a.conv_w[pc * 4 + 0], a.conv_w[pc * 4 + 1], a.conv_w[pc * 4 + 2], a.conv_w[pc * 4 + 3]};
  const float cb = a.conv_b[pc], dtb = a.dt_b[pc], Dk = a.Dskip[pc];
  float* my_cin = sm.cin + warp * N;
  float* my_ecar = sm.ecar + warp * N;
  float* my_a2 = sm.a2 + warp * N;
  float* my_dA2 = sm.dA2acc + warp * N;
  if (lane < N) {
    my_a2[lane] = a.A2[pc * N + lane];
    // adjoint carry-in: d loss / d (state after the last token) when a later shard consumes it (sequence sharding)
    my_ecar[lane] = (a.dhlast && active) ? a.dhlast[((int64_t)job * E + chc) * N + lane] : 0.f;
    my_dA2[lane] = 0.f;
  }
  float hal[3] = {0.f, 0.f, 0.f};
  if (a.halo) {
    const T* hp = static_cast<const T*>(a.halo) + ((int64_t)job * E + chc) * 3;
    hal[0] = io<T>::to_f(hp[0]); hal[1] = io<T>::to_f(hp[1]); hal[2] = io<T>::to_f(hp[2]);
  }
  auto halo_at = [&](int64_t tau) { return tau == -1 ? hal[2] : (tau == -2 ? hal[1] : (tau == -3 ? hal[0] : 0.f)); };
  // x at PHYSICAL time t, with the out-of-sequence rule of the forward (halo before logical 0, zero elsewhere)
  auto x_at = [&](int64_t t) -> float {
    if (t >= 0 && t < L) return io<T>::to_f(xrow[t]);
    const int64_t tau = REV ? L - 1 - t : t;
    return tau < 0 ? halo_at(tau) : 0.f;
  };

  const int seg = REV ? 31 - lane : lane;
  uint32_t poff[4];
  tile_piece_offsets(seg, poff);
  const int job_row = job * 2 * N;
  const int blocks_per_chunk = kChunk / kBlkTok;
  const int nthreads = blockDim.x;

  float dD_acc = 0.f, ddtb_acc = 0.f;
  float dt_next0 = 0.f;                      // dt of the first logical token of the chunk processed before

  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t first_c = nchunks - 1;                        // logical
    const int64_t first_pc = REV ? nchunks - 1 - first_c : first_c;
    mbar_expect_tx(sm.bar, 2 * N * kChunk * 4);
    tma_load_3d(sm.tile, tmap, 0, (int)(first_pc * blocks_per_chunk), job_row, sm.bar);
  }

  uint32_t parity = 0;
  for (int64_t c = nchunks - 1; c >= 0; --c) {
    const int64_t pcidx = REV ? nchunks - 1 - c : c;
    const int64_t tseg = pcidx * kChunk + (int64_t)seg * kTok;
    const bool seg_in = tseg < L;
    // nothing follows the last chunk unless an adjoint carry-in is given: then it acts as a virtual next token with
    // dt = 0 (a = 1) — masked tail tokens already pass e through unchanged (dt = 0, beta = 0)
    const bool last_chunk = (c == nchunks - 1) && a.dhlast == nullptr;

    // ---- state at the start of this chunk ---------------------------------------------------------------
    if (lane < N) {
      float v = 0.f;
      if (c > 0) v = a.chunk_state[(((int64_t)job * E + chc) * nchunks + (c - 1)) * N + lane];
      else if (a.h0) v = a.h0[((int64_t)job * E + chc) * N + lane];
      my_cin[lane] = v;
    }

    // ---- loads -------------------------------------------------------------------------------------------------
    float xs[kTok], dr[kTok], gs[kTok], zs[kTok];
    if (seg_in) {
      load_vec<T, kTok>(xrow + tseg, xs);
      load_vec<T, kTok>(drow + tseg, dr);
      load_vec<T, kTok>(gorow + tseg, gs);
      load_vec<T, kTok>(zrow + tseg, zs);
    } else {
#pragma unroll
      for (int i = 0; i < kTok; ++i) { xs[i] = 0.f; dr[i] = 0.f; gs[i] = 0.f; zs[i] = 0.f; }
    }

    // ---- prologue: recompute u, dt; dy = dout * silu(z) --------------------------------------------------------
    float u[kTok], dt[kTok], dy[kTok], sgd[kTok];     // sgd: d softplus / d(dt_raw)
    float dsum = 0.f;
    {
      float xl[kTok + 3];
#pragma unroll
      for (int i = 0; i < kTok; ++i) {
        float v = xs[phys(i)];
        const int64_t t = tseg + phys(i);
        if (t >= L) v = REV ? halo_at(L - 1 - t) : 0.f;
        xl[i + 3] = v;
      }
      // logical predecessors: previous lane, or (lane 0) the 3 tokens physically adjacent to this chunk
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
      if (lane == 0) {
        const int64_t tb = REV ? (pcidx + 1) * kChunk + 2 : pcidx * kChunk - 3;
        p0 = x_at(REV ? tb : tb);
        p1 = x_at(REV ? tb - 1 : tb + 1);
        p2 = x_at(REV ? tb - 2 : tb + 2);
      }
      const float u0 = __shfl_up_sync(0xffffffffu, xl[kTok + 0], 1);
      const float u1 = __shfl_up_sync(0xffffffffu, xl[kTok + 1], 1);
      const float u2 = __shfl_up_sync(0xffffffffu, xl[kTok + 2], 1);
      xl[0] = lane == 0 ? p0 : u0;
      xl[1] = lane == 0 ? p1 : u1;
      xl[2] = lane == 0 ? p2 : u2;
#pragma unroll
      for (int i = 0; i < kTok; ++i) {
        u[i] = silu_io<T>(cb + cw[0] * xl[i] + cw[1] * xl[i + 1] + cw[2] * xl[i + 2] + cw[3] * xl[i + 3]);
        const float raw = dr[phys(i)] + dtb;
        float d = softplus(raw);
        float sg = raw > 20.0f ? 1.0f : sigmoidf_fast(raw);
        const bool masked = tseg + phys(i) >= L;
        if (masked) { d = 0.f; sg = 0.f; }
        dt[i] = d;
        sgd[i] = sg;
        dsum += d;
        const float zz = zs[phys(i)];
        dy[i] = masked ? 0.f : gs[phys(i)] * silu_io<T>(zz);
      }
    }
    float ddt[kTok], dug[kTok], y[kTok], du[kTok];
#pragma unroll
    for (int i = 0; i < kTok; ++i) {
      ddt[i] = 0.f; dug[i] = dy[i] * Dk; y[i] = Dk * u[i]; dD_acc += dy[i] * u[i];
      du[i] = dt[i] * u[i];
    }
    const float dtn0 = __shfl_down_sync(0xffffffffu, dt[0], 1);
    const float dt_after = (lane == 31) ? dt_next0 : dtn0;       // dt of the token following my segment

    mbar_wait(sm.bar, parity);
    parity ^= 1;
    __syncwarp();
    const uint32_t tile_s = smem_u32(sm.tile);
    const uint32_t a2_s = smem_u32(my_a2), cin_s = smem_u32(my_cin), ecar_s = smem_u32(my_ecar);

#pragma unroll 1
    for (int n = 0; n < N; ++n) {
      const float A2n = lds32(a2_s + 4 * n);
      const float cin = lds32(cin_s + 4 * n);
      const float ecar = lds32(ecar_s + 4 * n);
      float av[kTok], hs[kTok], beta[kTok];
      float brow[kTok];
      float hin;
      // ---------- forward recompute ------------------------------------------------------------------------
      {
        float bv[kTok];
        const uint32_t rowp = tile_s + n * (kChunk * 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 q = lds128(rowp + poff[k]);
          brow[4 * k + 0] = q.x; brow[4 * k + 1] = q.y; brow[4 * k + 2] = q.z; brow[4 * k + 3] = q.w;
        }
        float hl = (lane == 0) ? cin : 0.f;
#pragma unroll
        for (int i = 0; i < kTok; ++i) {
          av[i] = ex2(dt[i] * A2n);
          bv[i] = du[i] * brow[phys(i)];
          hl = fmaf(av[i], hl, bv[i]);
        }
        float P = ex2(A2n * dsum);
        scan_step_up<1>(P, hl, lane);
        scan_step_up<2>(P, hl, lane);
        scan_step_up<4>(P, hl, lane);
        scan_step_up<8>(P, hl, lane);
        scan_step_up<16>(P, hl, lane);
        hin = __shfl_up_sync(0xffffffffu, hl, 1);
        if (lane == 0) hin = cin;
        float h = hin;
#pragma unroll
        for (int i = 0; i < kTok; ++i) { h = fmaf(av[i], h, bv[i]); hs[i] = h; }
      }
      {
        const uint32_t rowp = tile_s + (N + n) * (kChunk * 4);
        float cv[kTok];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 q = lds128(rowp + poff[k]);
          cv[4 * k + 0] = q.x; cv[4 * k + 1] = q.y; cv[4 * k + 2] = q.z; cv[4 * k + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < kTok; ++i) {
          y[i] = fmaf(cv[phys(i)], hs[i], y[i]);
          beta[i] = cv[phys(i)] * dy[i];
        }
      }
      // ---------- adjoint recurrence: e_i = beta_i + a_{i+1} e_{i+1} ----------------------------------------------
      const float av_next_lane = __shfl_down_sync(0xffffffffu, av[0], 1);
      const float a_after = (lane == 31) ? (last_chunk ? 0.f : ex2(dt_next0 * A2n)) : av_next_lane;
      float el = (lane == 31) ? ecar : 0.f;
      el = fmaf(a_after, el, beta[kTok - 1]);
#pragma unroll
      for (int i = kTok - 2; i >= 0; --i) el = fmaf(av[i + 1], el, beta[i]);
      float Q = ex2(A2n * (dsum - dt[0] + dt_after));            // prod_{i=0..15} a_{i+1}
      if (lane == 31 && last_chunk) Q = 0.f;
      scan_step_down<1>(Q, el, lane);
      scan_step_down<2>(Q, el, lane);
      scan_step_down<4>(Q, el, lane);
      scan_step_down<8>(Q, el, lane);
      scan_step_down<16>(Q, el, lane);
      float e = __shfl_down_sync(0xffffffffu, el, 1);            // e at the first token of the next lane
      if (lane == 31) e = ecar;
      __syncwarp();
      if (lane == 0) {
        sts32(ecar_s + 4 * n, el);                               // e at the first token of this chunk
        if (c == 0 && a.dh0 && active) a.dh0[((int64_t)job * E + ch) * N + n] = av[0] * el;
      }
      // ---------- gradients, walking the segment backwards --------------------------------------------------------
      float dA2n = 0.f;
      const float A2ln2 = A2n * kLn2f;
      float dBv[kTok], dCv[kTok];
#pragma unroll
      for (int i = kTok - 1; i >= 0; --i) {
        const float anx = (i == kTok - 1) ? a_after : av[i + 1];
        e = fmaf(anx, e, beta[i]);                               // e_i
        const float hprev = (i == 0) ? hin : hs[i - 1];
        const float t2 = e * hprev * av[i];
        dA2n = fmaf(t2, dt[i], dA2n);
        ddt[i] = fmaf(t2, A2ln2, ddt[i]);
        const float eB = e * brow[phys(i)];
        ddt[i] = fmaf(eB, u[i], ddt[i]);
        dug[i] = fmaf(eB, dt[i], dug[i]);
        dBv[phys(i)] = e * du[i];
        dCv[phys(i)] = dy[i] * hs[i];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dA2n += __shfl_xor_sync(0xffffffffu, dA2n, o);
      if (lane == 0) my_dA2[n] += dA2n * kLn2f;

      // ---------- dB / dC: sum over this CTA's channels, then one vector RED per 4 tokens --------------------------
      float* slot = sm.slots + ((size_t)(n & 1) * G + warp) * (2 * kChunk);
      if (!active) {
#pragma unroll
        for (int i = 0; i < kTok; ++i) { dBv[i] = 0.f; dCv[i] = 0.f; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        reinterpret_cast<float4*>(slot)[slot_piece(seg, k)] =
            make_float4(dBv[4 * k], dBv[4 * k + 1], dBv[4 * k + 2], dBv[4 * k + 3]);
        reinterpret_cast<float4*>(slot + kChunk)[slot_piece(seg, k)] =
            make_float4(dCv[4 * k], dCv[4 * k + 1], dCv[4 * k + 2], dCv[4 * k + 3]);
      }
      __syncthreads();
      {
        const float* sbase = sm.slots + (size_t)(n & 1) * G * (2 * kChunk);
        for (int q = threadIdx.x; q < 2 * (kChunk / 4); q += nthreads) {
          const int row = q / (kChunk / 4), p4 = q - row * (kChunk / 4);   // p4: swizzled piece index
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int w = 0; w < G; ++w) {
            const float4 v = reinterpret_cast<const float4*>(sbase + (size_t)w * (2 * kChunk) + row * kChunk)[p4];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          // un-swizzle: piece p4 = sg*4 + (k ^ ((sg>>1)&3))
          const int sg = p4 >> 2, k = (p4 & 3) ^ ((sg >> 1) & 3);
          const int64_t t = pcidx * kChunk + sg * kTok + 4 * k;
          if (t < L) {
            float* dst = a.dbc + ((int64_t)job_row + row * N + n) * a.ldbc + t;
            red_add_v4(dst, acc);                                // pad columns [L, ldbc) receive zeros only
          }
        }
      }
    }

    // ---- tile hand-over and request of the next (logically previous) chunk ---------------------------------
    __syncthreads();
    if (c > 0 && threadIdx.x == 0) {
      const int64_t npc = REV ? pcidx + 1 : pcidx - 1;
      mbar_expect_tx(sm.bar, 2 * N * kChunk * 4);
      tma_load_3d(sm.tile, tmap, 0, (int)(npc * blocks_per_chunk), job_row, sm.bar);
    }
    dt_next0 = __shfl_sync(0xffffffffu, dt[0], 0);

    // ---- per-token outputs: dz, d dt_raw, du ---------------------------------------------------------------------
    if (seg_in && active) {
      float o_dz[kTok], o_dd[kTok], o_du[kTok];
#pragma unroll
      for (int i = 0; i < kTok; ++i) {
        const float zz = zs[phys(i)];
        const float sg = sigmoidf_fast(zz);
        o_dz[phys(i)] = gs[phys(i)] * y[i] * sg * (1.0f + zz * (1.0f - sg));
        const float dd = ddt[i] * sgd[i];
        o_dd[phys(i)] = dd;
        ddtb_acc += dd;
        o_du[phys(i)] = dug[i];
      }
      if (tseg + kTok <= L) {
        store_vec<T, kTok>(dzrow + tseg, o_dz);
        store_vec<T, kTok>(ddrow + tseg, o_dd);
        store_vec<T, kTok>(durow + tseg, o_du);
      } else {
#pragma unroll
        for (int i = 0; i < kTok; ++i)
          if (tseg + i < L) {
            dzrow[tseg + i] = io<T>::from_f(o_dz[i]);
            ddrow[tseg + i] = io<T>::from_f(o_dd[i]);
            durow[tseg + i] = io<T>::from_f(o_du[i]);
          }
      }
    }
  }

  // ---- per-channel parameter gradients -----------------------------------------------------------------------------
  __syncwarp();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, o);
    ddtb_acc += __shfl_xor_sync(0xffffffffu, ddtb_acc, o);
  }
  if (active) {
    if (lane == 0) {
      atomicAdd(a.dDskip + pc, dD_acc);
      atomicAdd(a.ddt_b + pc, ddtb_acc);
    }
    if (lane < N) atomicAdd(a.dA2 + pc * N + lane, my_dA2[lane]);
  }
}

template <typename T, int N>
__global__ void __launch_bounds__(kMaxG * 32, 1)
bimamba_scan_bwd_kernel(const cad_scan_bwd_args a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int G = blockDim.x >> 5;
  BwdSmem sm;
  sm.tile = reinterpret_cast<float*>(base);
  sm.slots = reinterpret_cast<float*>(base + (size_t)2 * N * kChunk * 4);
  sm.cin = sm.slots + (size_t)2 * kMaxG * 2 * kChunk;
  sm.ecar = sm.cin + kMaxG * N;
  sm.a2 = sm.ecar + kMaxG * N;
  sm.dA2acc = sm.a2 + kMaxG * N;
  sm.bar = reinterpret_cast<uint64_t*>(sm.dA2acc + kMaxG * N);
  (void)G;
  if (threadIdx.x == 0) mbar_init(sm.bar, 1);
  const int job = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) scan_bwd_job<T, N, true>(a, &tmap, job, seq, pset, sm);
  else     scan_bwd_job<T, N, false>(a, &tmap, job, seq, pset, sm);
}

template <typename T, int N>
static int launch_scan_bwd(const cad_scan_bwd_args& a, int G, cudaStream_t stream) {
  CUtensorMap tmap;
  if (make_row_tile_map(&tmap, a.bc, (int64_t)a.njobs * 2 * N, a.ldbc, a.L, 2 * N) != 0) return -1;
  const size_t smem = 1024 + (size_t)2 * N * kChunk * 4 + (size_t)2 * kMaxG * 2 * kChunk * 4 +
                      (size_t)4 * kMaxG * N * sizeof(float) + 16;
  CAD_REQUIRE(smem <= 227 * 1024, "cad_bimamba_scan_bwd: needs %zu B of shared memory", smem);
  auto kern = bimamba_scan_bwd_kernel<T, N>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)((a.E + G - 1) / G), (unsigned)a.njobs);
  kern<<<grid, G * 32, smem, stream>>>(a, tmap);
  CAD_LAUNCH_CHECK();
  return 0;
}


// ---- conv + SiLU backward ---------------------------------------------------------------------------------------------
//   c[tau] = b + sum_k w[k] x[tau-3+k],  u = silu(c);   given du:  dc = du * silu'(c)
//   dx[tau] = sum_k w[k] dc[tau+3-k];   dw[k] += sum_tau dc[tau] x[tau-3+k];   db += sum_tau dc[tau]
// One thread per 16-byte vector of dx; block-level reduction of (dw, db), one atomicAdd per block.
template <typename T>
__global__ void __launch_bounds__(256) conv_silu_bwd_kernel(cad_conv_bwd_args a) {
  constexpr int V = 16 / sizeof(T);
  const int job = blockIdx.z;
  const int64_t ch = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  const T* __restrict__ x = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * a.E + ch) * a.ldxz;
  const T* __restrict__ du = static_cast<const T*>(a.du) + ((int64_t)job * a.E + ch) * a.lddu;
  T* __restrict__ dx = static_cast<T*>(a.dx) + ((int64_t)job * a.E + ch) * a.lddx;
  const int64_t pc = (int64_t)pset * a.E + ch;
  const float w0 = a.conv_w[pc * 4 + 0], w1 = a.conv_w[pc * 4 + 1], w2 = a.conv_w[pc * 4 + 2], w3 = a.conv_w[pc * 4 + 3];
  const float bias = a.conv_b[pc];
  const T* halo = a.halo ? static_cast<const T*>(a.halo) + ((int64_t)job * a.E + ch) * 3 : nullptr;
  const int64_t L = a.L;
  const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;

  float s_w[4] = {0.f, 0.f, 0.f, 0.f}, s_b = 0.f;
  if (t0 < L) {
    // physical window: outputs t0..t0+V-1 need dc at the 3 logically FOLLOWING tokens, which need x 3 further back:
    //   fwd: x[t0-3 .. t0+V+2], dc[t0 .. t0+V+2];   rev: x[t0-3 .. t0+V+2], dc[t0-3 .. t0+V-1]
    float xw[V + 6];
    auto x_at = [&](int64_t t) -> float {
      if (t >= 0 && t < L) return io<T>::to_f(x[t]);
      if (halo) {
        const int64_t tau = rev ? (L - 1 - t) : t;
        if (tau >= -3 && tau < 0) return io<T>::to_f(halo[tau + 3]);
      }
      return 0.f;
    };
#pragma unroll
    for (int i = 0; i < V + 6; ++i) xw[i] = x_at(t0 - 3 + i);
    // dc at physical positions p = t0-3+j (rev) or t0+j (fwd), j in [0, V+3)
    float dc[V + 3];
#pragma unroll
    for (int j = 0; j < V + 3; ++j) {
      const int64_t t = rev ? t0 - 3 + j : t0 + j;
      float v = 0.f;
      if (t >= 0 && t < L) {
        // conv pre-activation at physical t: fwd taps x[t-3..t] (x[t] = xw[j+3]); rev taps x[t..t+3] (x[t] = xw[j])
        float c;
        if (!rev) c = bias + w0 * xw[j] + w1 * xw[j + 1] + w2 * xw[j + 2] + w3 * xw[j + 3];
        else      c = bias + w3 * xw[j] + w2 * xw[j + 1] + w1 * xw[j + 2] + w0 * xw[j + 3];
        const float sg = rcp(1.0f + ex2(-kLog2e * c));
        v = io<T>::to_f(du[t]) * sg * (1.0f + c * (1.0f - sg));
        // parameter gradients are owned by the thread whose OUTPUT range holds t
        if (t >= t0 && t < t0 + V) {
          s_b += v;
          if (!rev) { s_w[0] += v * xw[j]; s_w[1] += v * xw[j + 1]; s_w[2] += v * xw[j + 2]; s_w[3] += v * xw[j + 3]; }
          else      { s_w[3] += v * xw[j]; s_w[2] += v * xw[j + 1]; s_w[1] += v * xw[j + 2]; s_w[0] += v * xw[j + 3]; }
        }
      }
      dc[j] = v;
    }
    uint4 outv;
    T* o = reinterpret_cast<T*>(&outv);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      // fwd: dx[t] = w3 dc[t] + w2 dc[t+1] + w1 dc[t+2] + w0 dc[t+3];  rev: mirrored
      float g;
      if (!rev) g = w3 * dc[i] + w2 * dc[i + 1] + w1 * dc[i + 2] + w0 * dc[i + 3];
      else      g = w3 * dc[i + 3] + w2 * dc[i + 2] + w1 * dc[i + 1] + w0 * dc[i];
      o[i] = io<T>::from_f(g);
    }
    *reinterpret_cast<uint4*>(dx + t0) = outv;
  }
  // block reduction of the 5 parameter partials
  __shared__ float red[5][8];
  float vals[5] = {s_w[0], s_w[1], s_w[2], s_w[3], s_b};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = vals[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    if (threadIdx.x < 4) atomicAdd(a.dconv_w + pc * 4 + threadIdx.x, s);
    else atomicAdd(a.dconv_b + pc, s);
  }
}

}  // namespace cad

extern "C" int cad_bimamba_scan_bwd(const cad_scan_bwd_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_bimamba_scan_bwd: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_bimamba_scan_bwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->delta && a->bc && a->dout && a->conv_w && a->conv_b && a->dt_b && a->A2 && a->Dskip &&
              a->seq_of_job && a->pset_of_job && a->rev_of_job && a->chunk_state && a->dz && a->du && a->ddelta &&
              a->dbc && a->ddt_b && a->dA2 && a->dDskip, "cad_bimamba_scan_bwd: null pointer");
  CAD_REQUIRE(a->N == 16, "cad_bimamba_scan_bwd: d_state = %lld not built (only 16)", (long long)a->N);
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldd % 16 == 0 && a->ldo % 16 == 0 && a->lddz % 16 == 0 && a->lddu % 16 == 0 &&
              a->lddd % 16 == 0, "cad_bimamba_scan_bwd: row pitches must be multiples of 16 elements");
  CAD_REQUIRE(a->ldbc % 32 == 0 && a->ldbc >= a->L, "cad_bimamba_scan_bwd: ldbc must be a multiple of 32 and >= L");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int G = a->channels_per_cta;
  if (G <= 0) {
    const int sms = cad_sm_count() > 0 ? cad_sm_count() : 148;
    long best = -1;
    for (int g = 1; g <= kMaxG; ++g) {
      const long ctas = (long)a->njobs * ((a->E + g - 1) / g);
      const long cost = ((ctas + sms - 1) / sms) * g;
      if (best < 0 || cost <= best) { best = cost; G = g; }
    }
  }
  CAD_REQUIRE(G >= 1 && G <= kMaxG, "cad_bimamba_scan_bwd: channels_per_cta must be in [1, %d]", kMaxG);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, return launch_scan_bwd<T, 16>(*a, G, stream));
  return 0;
}

extern "C" int cad_conv_silu_bwd(const cad_conv_bwd_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_conv_silu_bwd: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0, "cad_conv_silu_bwd: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->du && a->dx && a->conv_w && a->conv_b && a->dconv_w && a->dconv_b && a->seq_of_job &&
              a->pset_of_job && a->rev_of_job, "cad_conv_silu_bwd: null pointer");
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->lddu % 16 == 0 && a->lddx % 16 == 0, "cad_conv_silu_bwd: row pitches must be "
              "multiples of 16 elements");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t v = 16 / (int64_t)dtype_size(a->io_dtype);
  dim3 grid((unsigned)((a->L + v * 256 - 1) / (v * 256)), (unsigned)a->E, (unsigned)a->njobs);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, conv_silu_bwd_kernel<T><<<grid, 256, 0, stream>>>(*a));
  CAD_LAUNCH_CHECK();
  return 0;
}

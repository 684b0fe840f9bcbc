// Fused bidirectional selective-scan forward for sm_100a.
//
// One launch covers every (sequence, direction) "job" of a BiMamba call — forward and reverse directions of
// ref:caduceus/modeling_caduceus.py:128-137 and both strands of ref:caduceus/modeling_rcps.py:85-99 — with the
// reversed jobs handled purely by addressing (no flipped copies).  Per job and channel it fuses what the
// reference runs as separate upstream kernels:
//     depthwise (anti)causal conv + bias + SiLU      (causal_conv1d_fwd)
//     dt = softplus(W_dt . x_dbl[0:R] + b_dt)        (the dt_proj GEMM + the softplus inside selective_scan)
//     h_t = exp2(dt*A2) h_{t-1} + dt*B_t*u_t ;  y_t = C_t . h_t + D u_t ;  out = y * silu(z)   (selective_scan_fwd)
//
// Work decomposition (B200: 148 SMs, the scan is MUFU/issue-bound, not HBM-bound — SURVEY.md §8d):
//   CTA  = one job x G consecutive channels, one WARP PER CHANNEL, walking the sequence in logical-time
//          chunks of 512 tokens.  G is chosen by the host so that the grid is ~one full wave of 148 CTAs.
//   lane = 16 consecutive tokens of the chunk.  Each lane runs the recurrence over its 16 tokens from a zero
//          state, the 32 segment aggregates (prod a, h_end) are combined with a 5-step warp-shuffle scan, and
//          the lane re-runs its 16 FMAs from the true incoming state.  exp2 is evaluated ONCE per
//          (token, channel, state) — the segment decay is exp2(A2 * sum(dt)) instead of a product.
//   smem = the chunk's (R + 2N) x 512 tile of dt-low-rank / B / C rows, shared by the G channels: staged by
//          cp.async in the I/O dtype one chunk ahead, then widened once to fp32 in a bank-conflict-free
//          (XOR-swizzled) layout for 128-bit reads.
#include "common.cuh"

namespace cad {

constexpr int kTok = 16;            // tokens per lane
constexpr int kChunk = 32 * kTok;   // 512 logical tokens per chunk
constexpr int kMaxG = 8;            // channels (warps) per CTA

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// fp32 tile addressing: row-major [row][512 tokens], 16-byte pieces XOR-swizzled inside each lane's 64-byte
// segment so that the 8 lanes of an LDS.128 phase hit 8 distinct bank groups.
__device__ __forceinline__ int tile_piece(int seg, int k) { return seg * 4 + (k ^ ((seg >> 1) & 3)); }

template <typename T, int N, bool REV>
__device__ __forceinline__ void scan_job(const cad_scan_fwd_args& a, int job, int seq, int pset,
                                         float* tile, T* stage, float* carry_s, float* par_s) {
  constexpr int EPV = 16 / sizeof(T);          // elements per 16-byte vector of the io dtype
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = blockDim.x >> 5;
  const int64_t L = a.L, E = a.E;
  const int R = (int)a.R;
  const int rows = R + 2 * N;
  const int64_t ch = (int64_t)blockIdx.x * G + warp;
  const bool active = ch < E;                  // tail CTA: idle warps still help with the tile
  const int64_t chc = active ? ch : E - 1;
  const int64_t nchunks = (L + kChunk - 1) / kChunk;

  const T* __restrict__ xrow = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + chc) * a.ldxz;
  const T* __restrict__ zrow = xrow + E * a.ldxz;
  const T* __restrict__ xd = static_cast<const T*>(a.xdbl) + (int64_t)job * rows * a.ldxd;
  T* __restrict__ orow = static_cast<T*>(a.out) + ((int64_t)job * E + chc) * a.ldo;

  // per-channel parameters -> registers / smem
  const int64_t pc = (int64_t)pset * E + chc;
  const float cw0 = a.conv_w[pc * 4 + 0], cw1 = a.conv_w[pc * 4 + 1], cw2 = a.conv_w[pc * 4 + 2],
              cw3 = a.conv_w[pc * 4 + 3];
  const float cb = a.conv_b[pc], dtb = a.dt_b[pc], Dk = a.Dskip[pc];
  float* my_par = par_s + warp * (N + 32);     // [0,N): A2 ; [N, N+R): dt_w
  float* my_carry = carry_s + warp * N;
  if (lane < N) {
    my_par[lane] = a.A2[pc * N + lane];
    my_carry[lane] = a.h0 ? a.h0[((int64_t)job * E + chc) * N + lane] : 0.f;
  }
  for (int r = lane; r < R; r += 32) my_par[N + r] = a.dt_w[pc * R + r];

  // x values preceding logical time 0 (sequence-shard halo): hal[k] = x[tau = k - 3].  Chunks are aligned in
  // PHYSICAL time, so for a reversed job the first logical chunk may start with masked tokens (t >= L, i.e.
  // tau < 0): the halo lives exactly there, and further back everything is zero.
  float hal[3] = {0.f, 0.f, 0.f};
  if (a.halo) {
    const T* hp = static_cast<const T*>(a.halo) + ((int64_t)job * E + chc) * 3;
    hal[0] = io<T>::to_f(hp[0]); hal[1] = io<T>::to_f(hp[1]); hal[2] = io<T>::to_f(hp[2]);
  }
  auto halo_at = [&](int64_t tau) {   // tau < 0
    return tau == -1 ? hal[2] : (tau == -2 ? hal[1] : (tau == -3 ? hal[0] : 0.f));
  };
  const int64_t tau0 = REV ? L - nchunks * kChunk : 0;      // logical time of the first item of chunk 0 (<= 0)
  float prev3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) prev3[k] = halo_at(tau0 - 3 + k);
  float dt_total = 0.f;

  // ---- stage the xdbl tile of physical chunk `pcidx` with cp.async (16-byte pieces, zero beyond the pitch)
  auto stage_chunk = [&](int64_t pcidx) {
    const int64_t t0 = pcidx * kChunk;
    const int pieces_per_row = kChunk / EPV;
    for (int p = threadIdx.x; p < rows * pieces_per_row; p += blockDim.x) {
      const int r = p / pieces_per_row, q = p - r * pieces_per_row;
      const int64_t t = t0 + (int64_t)q * EPV;
      T* dst = stage + (int64_t)r * kChunk + q * EPV;
      if (t < a.ldxd) cp_async16(dst, xd + (int64_t)r * a.ldxd + t);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
  };
  // ---- widen stage -> fp32 swizzled tile, zeroing tokens >= L (pad columns may hold garbage / NaN)
  auto widen_chunk = [&](int64_t pcidx) {
    const int64_t t0 = pcidx * kChunk;
    for (int p = threadIdx.x; p < rows * (kChunk / 4); p += blockDim.x) {
      const int r = p / (kChunk / 4), q = p - r * (kChunk / 4);   // q: group of 4 tokens
      const int seg = q >> 2, k = q & 3;
      float4 v;
      const T* s = stage + (int64_t)r * kChunk + q * 4;
      v.x = io<T>::to_f(s[0]); v.y = io<T>::to_f(s[1]); v.z = io<T>::to_f(s[2]); v.w = io<T>::to_f(s[3]);
      const int64_t t = t0 + q * 4;
      if (t + 0 >= L) v.x = 0.f;
      if (t + 1 >= L) v.y = 0.f;
      if (t + 2 >= L) v.z = 0.f;
      if (t + 3 >= L) v.w = 0.f;
      reinterpret_cast<float4*>(tile + (int64_t)r * kChunk)[tile_piece(seg, k)] = v;
    }
  };
  // this lane's 16 tokens of tile row r, in PHYSICAL order
  const int seg = REV ? 31 - lane : lane;
  auto tile_row = [&](int r, float (&v)[kTok]) {
    const float4* p = reinterpret_cast<const float4*>(tile + (int64_t)r * kChunk);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 q = p[tile_piece(seg, k)];
      v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  };
  // logical item i of this lane <-> physical offset inside its 16-token segment
  auto phys = [](int i) { return REV ? kTok - 1 - i : i; };

  if (nchunks > 0) stage_chunk(REV ? nchunks - 1 : 0);

  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t pcidx = REV ? nchunks - 1 - c : c;
    const int64_t tseg = pcidx * kChunk + (int64_t)seg * kTok;    // first physical token of my segment
    const bool seg_in = tseg < L;                                 // (pitch is a multiple of 16: whole segment loadable)

    // 1. my x and z segments (global -> registers); latency overlaps with the tile hand-over below
    float xs[kTok];
    uint4 zraw[kTok / EPV];
    if (seg_in && active) {
      load_vec<T, kTok>(xrow + tseg, xs);
#pragma unroll
      for (int i = 0; i < kTok / EPV; ++i) zraw[i] = __ldg(reinterpret_cast<const uint4*>(zrow + tseg) + i);
    } else {
#pragma unroll
      for (int i = 0; i < kTok; ++i) xs[i] = 0.f;
#pragma unroll
      for (int i = 0; i < kTok / EPV; ++i) zraw[i] = make_uint4(0, 0, 0, 0);
    }

    // 2. tile hand-over: staged chunk has landed and everyone is done with the previous fp32 tile
    cp_async_wait_all();
    __syncthreads();
    widen_chunk(pcidx);
    __syncthreads();
    if (c + 1 < nchunks) stage_chunk(REV ? nchunks - 2 - c : c + 1);

    // 3. per-(token, channel) prologue: conv + SiLU, dt, dt*u
    float u[kTok], dt[kTok];
    {
      float xl[kTok + 3];
#pragma unroll
      for (int i = 0; i < kTok; ++i) {
        float v = xs[phys(i)];
        const int64_t t = tseg + phys(i);
        if (t >= L) v = REV ? halo_at(L - 1 - t) : 0.f;
        xl[i + 3] = v;
      }
      // logical predecessors: last 3 logical tokens of the previous lane / previous chunk
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float up = __shfl_up_sync(0xffffffffu, xl[kTok + k], 1);
        xl[k] = (lane == 0) ? prev3[k] : up;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) prev3[k] = __shfl_sync(0xffffffffu, xl[kTok + k], 31);
#pragma unroll
      for (int i = 0; i < kTok; ++i)
        u[i] = silu(cb + cw0 * xl[i] + cw1 * xl[i + 1] + cw2 * xl[i + 2] + cw3 * xl[i + 3]);
    }
#pragma unroll
    for (int i = 0; i < kTok; ++i) dt[i] = dtb;
    for (int r = 0; r < R; ++r) {
      float row[kTok];
      tile_row(r, row);
      const float w = my_par[N + r];
#pragma unroll
      for (int i = 0; i < kTok; ++i) dt[i] = fmaf(w, row[phys(i)], dt[i]);
    }
    float dsum = 0.f;
    float du[kTok], y[kTok];
#pragma unroll
    for (int i = 0; i < kTok; ++i) {
      float d = softplus(dt[i]);
      if (tseg + phys(i) >= L) d = 0.f;        // masked token: a = 1, b = 0 -> state passes through
      dt[i] = d;
      dsum += d;
      du[i] = d * u[i];
      y[i] = Dk * u[i];
    }
    dt_total += dsum;

    // 4. the scan, one state at a time
#pragma unroll 1
    for (int n = 0; n < N; ++n) {
      const float A2n = my_par[n];
      const float cin = my_carry[n];
      float av[kTok], bv[kTok];
      float hl = (lane == 0) ? cin : 0.f;
      {
        float brow[kTok];
        tile_row(R + n, brow);
#pragma unroll
        for (int i = 0; i < kTok; ++i) {
          av[i] = ex2(dt[i] * A2n);
          bv[i] = du[i] * brow[phys(i)];
          hl = fmaf(av[i], hl, bv[i]);
        }
      }
      float P = ex2(A2n * dsum);
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float Pp = __shfl_up_sync(0xffffffffu, P, off);
        const float Hp = __shfl_up_sync(0xffffffffu, hl, off);
        if (lane >= off) { hl = fmaf(P, Hp, hl); P *= Pp; }
      }
      float h = __shfl_up_sync(0xffffffffu, hl, 1);
      if (lane == 0) h = cin;
      __syncwarp();
      if (lane == 31) my_carry[n] = hl;        // state at the end of this chunk
      {
        float crow[kTok];
        tile_row(R + N + n, crow);
#pragma unroll
        for (int i = 0; i < kTok; ++i) {
          h = fmaf(av[i], h, bv[i]);
          y[i] = fmaf(crow[phys(i)], h, y[i]);
        }
      }
    }
    __syncwarp();
    if (a.chunk_state && active && lane < N)
      a.chunk_state[(((int64_t)job * E + ch) * nchunks + c) * N + lane] = my_carry[lane];

    // 5. gate with silu(z) and store (physical order)
    if (seg_in && active) {
      float o[kTok];
      const T* ze = reinterpret_cast<const T*>(zraw);
#pragma unroll
      for (int i = 0; i < kTok; ++i) {
        const float zz = io<T>::to_f(ze[phys(i)]);
        o[phys(i)] = y[i] * silu(zz);
      }
      if (tseg + kTok <= L) {
        store_vec<T, kTok>(orow + tseg, o);
      } else {
#pragma unroll
        for (int i = 0; i < kTok; ++i)
          if (tseg + i < L) orow[tseg + i] = io<T>::from_f(o[i]);
      }
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

template <typename T, int N>
__global__ void __launch_bounds__(kMaxG * 32, 1) bimamba_scan_fwd_kernel(cad_scan_fwd_args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int rows = (int)a.R + 2 * N;
  float* tile = reinterpret_cast<float*>(smem_raw);
  T* stage = reinterpret_cast<T*>(smem_raw + (size_t)rows * kChunk * sizeof(float));
  float* carry_s = reinterpret_cast<float*>(smem_raw + (size_t)rows * kChunk * (sizeof(float) + sizeof(T)));
  float* par_s = carry_s + kMaxG * N;
  const int job = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) scan_job<T, N, true>(a, job, seq, pset, tile, stage, carry_s, par_s);
  else     scan_job<T, N, false>(a, job, seq, pset, tile, stage, carry_s, par_s);
}

template <typename T, int N>
static int launch_scan(const cad_scan_fwd_args& a, int G, cudaStream_t stream) {
  const int rows = (int)a.R + 2 * N;
  const size_t smem = (size_t)rows * kChunk * (sizeof(float) + sizeof(T)) + (size_t)kMaxG * (N + N + 32) * sizeof(float);
  CAD_REQUIRE(smem <= 227 * 1024, "cad_bimamba_scan_fwd: tile of %d rows needs %zu B of shared memory", rows, smem);
  auto kern = bimamba_scan_fwd_kernel<T, N>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)((a.E + G - 1) / G), (unsigned)a.njobs);
  kern<<<grid, G * 32, smem, stream>>>(a);
  CAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace cad

extern "C" int cad_scan_chunk_len(void) { return cad::kChunk; }

extern "C" int cad_bimamba_scan_fwd(const cad_scan_fwd_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a && a->xz && a->xdbl && a->out && a->conv_w && a->conv_b && a->dt_w && a->dt_b && a->A2 && a->Dskip &&
              a->seq_of_job && a->pset_of_job && a->rev_of_job, "cad_bimamba_scan_fwd: null pointer");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_bimamba_scan_fwd: bad sizes");
  CAD_REQUIRE(a->N == 16, "cad_bimamba_scan_fwd: d_state = %lld not built (only 16)", (long long)a->N);
  CAD_REQUIRE(a->R >= 1 && a->R <= 32, "cad_bimamba_scan_fwd: dt_rank = %lld out of range [1, 32]", (long long)a->R);
  CAD_REQUIRE(a->K >= 1 && a->K <= 4, "cad_bimamba_scan_fwd: d_conv = %lld out of range [1, 4]", (long long)a->K);
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldxd % 16 == 0 && a->ldo % 16 == 0 && a->ldxz >= a->L && a->ldxd >= a->L &&
              a->ldo >= a->L, "cad_bimamba_scan_fwd: row pitches must be multiples of 16 elements and >= L");
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->xdbl) && aligned16(a->out),
              "cad_bimamba_scan_fwd: xz/xdbl/out must be 16-byte aligned");
  if (a->L == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int G = a->channels_per_cta;
  if (G <= 0) {
    // one CTA per SM is resident (the tile takes > 113 KB): minimise waves * G
    const int sms = cad_sm_count() > 0 ? cad_sm_count() : 148;
    long best = -1;
    for (int g = 1; g <= kMaxG; ++g) {
      const long ctas = (long)a->njobs * ((a->E + g - 1) / g);
      const long cost = ((ctas + sms - 1) / sms) * g;
      if (best < 0 || cost <= best) { best = cost; G = g; }
    }
  }
  CAD_REQUIRE(G >= 1 && G <= kMaxG, "cad_bimamba_scan_fwd: channels_per_cta must be in [1, %d]", kMaxG);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, return launch_scan<T, 16>(*a, G, stream));
  return 0;
}

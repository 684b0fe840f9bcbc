// Carry fix-up for a sequence-sharded scan, sm_100a.
//
// A shard scanned from a ZERO state differs from the true result only by the contribution of its carry-in state h0:
//     y_true[t] = y_zero[t] + sum_n C[t,n] * exp2(A2[n] * cumdt[t]) * h0[n],     cumdt[t] = sum_{s<=t} dt[s]
// (the recurrence is affine in the state, and the product of the decays a_s = exp2(dt_s*A2) over the shard prefix is
// exp2(A2 * cumdt)).  This kernel adds that term — gated by silu(z) like the forward — to the shard's output in place.
// The factor only decays along the shard, so every (channel, state) is dropped for good once its exponent falls below
// `cutoff_log2` (|term| < 2^cutoff * |C h0|), warps retire when all 16 states are gone and the CTA stops when all its
// warps have retired: at init-time dynamics (dt ~ 1e-3..1e-1, |A| = 1..16) the touched prefix is a few hundred to a
// few thousand tokens, i.e. a few percent of a 16k-token shard, which is what makes ONE all_gather per layer enough
// for near-linear strong scaling (no second full scan, no serial rank chain).
#include "scan_common.cuh"

namespace cad {

template <typename T, int N, bool REV>
__device__ __forceinline__ void fixup_job(const cad_scan_fixup_args& a, const CUtensorMap* tmap, int job, int seq,
                                          int pset, int64_t t_off, int64_t L, const float* __restrict__ h0_base,
                                          float* tile, float* h0_s, float* a2_s, uint64_t* bar) {
  // (t_off, L): the token range [t_off, t_off + L) this call works on — the whole sequence (0, a.L), or one in-GPU segment
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = blockDim.x >> 5;
  const int64_t E = a.E;
  const int64_t ch = (int64_t)blockIdx.x * G + warp;
  const bool active = ch < E;
  const int64_t chc = active ? ch : E - 1;
  const int64_t nchunks = (L + kChunk - 1) / kChunk;
  auto phys = [](int i) { return REV ? kTok - 1 - i : i; };

  const T* __restrict__ zrow = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + E + chc) * a.ldxz + t_off;
  const T* __restrict__ drow = static_cast<const T*>(a.delta) + ((int64_t)job * E + chc) * a.ldd + t_off;
  T* __restrict__ orow = static_cast<T*>(a.out) + ((int64_t)job * E + chc) * a.ldo + t_off;
  const int64_t pc = (int64_t)pset * E + chc;
  const float dtb = a.dt_b[pc];
  float* my_h0 = h0_s + warp * N;
  float* my_a2 = a2_s + warp * N;
  if (lane < N) {
    my_a2[lane] = a.A2[pc * N + lane];
    my_h0[lane] = active ? h0_base[chc * N + lane] : 0.f;
  }
  __syncwarp();
  // states still contributing (warp-uniform bit mask); a zero carry never contributes
  unsigned alive = 0;
  for (int n = 0; n < N; ++n)
    if (my_h0[n] != 0.f) alive |= 1u << n;

  const int seg = REV ? 31 - lane : lane;
  uint32_t poff[4];
  tile_piece_offsets(seg, poff);
  const int c_row = job * 2 * N + N;                   // the C rows of this job
  const int blocks_per_chunk = kChunk / kBlkTok;
  const int blk_off = (int)(t_off / kBlkTok);               // t_off is a multiple of 256
  float cum_base = 0.f;
  uint32_t parity = 0;

  __syncthreads();
  bool any = __syncthreads_or(alive != 0);
  if (!any) return;
  if (threadIdx.x == 0) {
    const int64_t first = REV ? nchunks - 1 : 0;
    mbar_expect_tx(bar, N * kChunk * 4);
    tma_load_3d(tile, tmap, 0, blk_off + (int)(first * blocks_per_chunk), c_row, bar);
  }

  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t pcidx = REV ? nchunks - 1 - c : c;
    const int64_t tseg = pcidx * kChunk + (int64_t)seg * kTok;
    const bool seg_in = tseg < L;
    // drop the states that have decayed away before this chunk starts
    for (int n = 0; n < N; ++n)
      if ((alive >> n) & 1u)
        if (my_a2[n] * cum_base < a.cutoff_log2) alive &= ~(1u << n);

    mbar_wait(bar, parity);
    parity ^= 1;
    if (alive) {
      float dr[kTok], cum[kTok], y[kTok];
      if (seg_in) load_vec<T, kTok>(drow + tseg, dr);
      else {
#pragma unroll
        for (int i = 0; i < kTok; ++i) dr[i] = 0.f;
      }
      float run = 0.f;
#pragma unroll
      for (int i = 0; i < kTok; ++i) {
        float d = softplus(dr[phys(i)] + dtb);
        if (tseg + phys(i) >= L) d = 0.f;
        run += d;
        cum[i] = run;
        y[i] = 0.f;
      }
      // exclusive prefix of the lane totals (logical lane order) + the chunks before
      float incl = run;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      const float base = cum_base + incl - run;
      cum_base += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
      for (int i = 0; i < kTok; ++i) cum[i] += base;

      const uint32_t tile_s = smem_u32(tile);
#pragma unroll 1
      for (int n = 0; n < N; ++n) {
        if (!((alive >> n) & 1u)) continue;
        const float A2n = my_a2[n], hn = my_h0[n];
        const uint32_t rowp = tile_s + n * (kChunk * 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 q = lds128(rowp + poff[k]);
          const float cq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = REV ? kTok - 1 - (4 * k + e) : 4 * k + e;
            y[i] = fmaf(cq[e] * hn, ex2(A2n * cum[i]), y[i]);
          }
        }
      }
      if (seg_in && active) {
        float zs[kTok], os[kTok];
        load_vec<T, kTok>(zrow + tseg, zs);
        load_vec<T, kTok>(orow + tseg, os);
#pragma unroll
        for (int i = 0; i < kTok; ++i) os[phys(i)] = fmaf(y[i], silu_io<T>(zs[phys(i)]), os[phys(i)]);
        if (tseg + kTok <= L) {
          store_vec<T, kTok>(orow + tseg, os);
        } else {
#pragma unroll
          for (int i = 0; i < kTok; ++i)
            if (tseg + i < L) orow[tseg + i] = io<T>::from_f(os[i]);
        }
      }
    } else {
      // retired warp: keep cum_base meaningless but harmless; it only follows the CTA's barriers
    }
    any = __syncthreads_or(alive != 0);                // also: everyone is done with the tile
    if (!any) return;
    if (c + 1 < nchunks && threadIdx.x == 0) {
      const int64_t npc = REV ? pcidx - 1 : pcidx + 1;
      mbar_expect_tx(bar, N * kChunk * 4);
      tma_load_3d(tile, tmap, 0, blk_off + (int)(npc * blocks_per_chunk), c_row, bar);
    }
  }
}

template <typename T, int N>
__global__ void __launch_bounds__(kMaxG * 32, 2)
scan_fixup_kernel(const cad_scan_fixup_args a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tile = reinterpret_cast<float*>(base);
  float* h0_s = reinterpret_cast<float*>(base + (size_t)N * kChunk * 4);
  float* a2_s = h0_s + kMaxG * N;
  uint64_t* bar = reinterpret_cast<uint64_t*>(a2_s + kMaxG * N);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  int job = blockIdx.y;
  int64_t t_off = 0, L = a.L;
  const float* h0_base;
  if (a.nseg > 1) {
    // blockIdx.y = (job, logical segment s >= 1); the segment's tokens = physical block k of the split used by scan variant 20
    // (scan_fwd_v20.cuh::block_range: whole 256-token chunks, ceil(nchunks / nseg) per block)
    const int sl = 1 + (int)(blockIdx.y % (a.nseg - 1));
    job = (int)(blockIdx.y / (a.nseg - 1));
    const int64_t k = a.rev_of_job[job] ? a.nseg - 1 - sl : sl;
    const int64_t nch = (a.L + 255) / 256, per = (nch + a.nseg - 1) / a.nseg;
    int64_t lo = k * per * 256, hi = (k + 1) * per * 256;
    if (hi > a.L) hi = a.L;
    if (lo >= hi) return;                                    // empty block (CTA-uniform)
    t_off = lo; L = hi - lo;
    h0_base = a.seg_carry + ((int64_t)job * a.nseg + sl) * a.E * N;
  } else {
    h0_base = a.h0 + (int64_t)job * a.E * N;
  }
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) fixup_job<T, N, true>(a, &tmap, job, seq, pset, t_off, L, h0_base, tile, h0_s, a2_s, bar);
  else     fixup_job<T, N, false>(a, &tmap, job, seq, pset, t_off, L, h0_base, tile, h0_s, a2_s, bar);
}

template <typename T, int N>
static int launch_fixup(const cad_scan_fixup_args& a, int G, cudaStream_t stream) {
  CUtensorMap tmap;
  if (make_row_tile_map(&tmap, a.bc, (int64_t)a.njobs * 2 * N, a.ldbc, a.L, N) != 0) return -1;
  const size_t smem = 1024 + (size_t)N * kChunk * 4 + (size_t)2 * kMaxG * N * sizeof(float) + 16;
  auto kern = scan_fixup_kernel<T, N>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)((a.E + G - 1) / G), (unsigned)(a.nseg > 1 ? a.njobs * (a.nseg - 1) : a.njobs));
  kern<<<grid, G * 32, smem, stream>>>(a, tmap);
  CAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace cad

extern "C" int cad_bimamba_scan_fixup(const cad_scan_fixup_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_bimamba_scan_fixup: null argument block");
  CAD_REQUIRE(a->L >= 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_bimamba_scan_fixup: bad sizes");
  if (a->L == 0) return 0;
  CAD_REQUIRE(a->xz && a->delta && a->bc && a->out && a->dt_b && a->A2 && a->seq_of_job && a->pset_of_job &&
              a->rev_of_job && (a->nseg > 1 ? a->seg_carry != nullptr : a->h0 != nullptr), "cad_bimamba_scan_fixup: null pointer");
  CAD_REQUIRE(a->nseg >= 0 && a->nseg <= 4096, "cad_bimamba_scan_fixup: nseg out of range");
  CAD_REQUIRE(a->N == 16, "cad_bimamba_scan_fixup: d_state = %lld not built (only 16)", (long long)a->N);
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldd % 16 == 0 && a->ldo % 16 == 0 && a->ldbc % 32 == 0,
              "cad_bimamba_scan_fixup: bad row pitches");
  CAD_REQUIRE(a->cutoff_log2 < 0.f, "cad_bimamba_scan_fixup: cutoff_log2 must be negative");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int G = a->channels_per_cta;
  if (G <= 0) G = kMaxG;
  CAD_REQUIRE(G >= 1 && G <= kMaxG, "cad_bimamba_scan_fixup: channels_per_cta must be in [1, %d]", kMaxG);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, return launch_fixup<T, 16>(*a, G, stream));
  return 0;
}

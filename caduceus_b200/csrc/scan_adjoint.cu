// Adjoint of the carry for sequence-sharded TRAINING (SURVEY.md §8e "Backward"), sm_100a.
//
// The backward of a shard needs the adjoint e = dLoss/dh entering from the logically NEXT shard, which in turn needs
// the one after it: a serial chain over the ranks.  Like the forward carry it is affine,
//     dh0_k = P_k * dhlast_k + Dh_k,     P_k = exp2(A2 * sum dt over shard k)   (already known from the forward),
//     Dh_k[n] = sum_tau exp2(A2[n] * cumdt[tau]) * C[tau,n] * dy[tau],   dy = dout * silu(z),  cumdt[tau] = sum_{s<=tau} dt[s]
// so every rank computes its own Dh_k (this kernel), ONE all_gather distributes them, each rank composes its true
// dhlast locally and runs the full backward (scan_bwd.cu) once.  Dh is the exact transpose of the forward fix-up
// (scan_fixup.cu): same weights exp2(A2 * cumdt), same decay, same cut-off — (channel, state) pairs are dropped once
// A2 * cumdt < cutoff_log2 and the CTA stops when all are gone, so only a prefix of the shard is read.
#include "scan_common.cuh"

namespace cad {

template <typename T, int N, bool REV>
__device__ __forceinline__ void adjoint_job(const cad_scan_adjoint_args& a, const CUtensorMap* tmap, int job, int seq,
                                            int pset, float* tile, float* acc_s, float* a2_s, uint64_t* bar) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = blockDim.x >> 5;
  const int64_t L = a.L, E = a.E;
  const int64_t ch = (int64_t)blockIdx.x * G + warp;
  const bool active = ch < E;
  const int64_t chc = active ? ch : E - 1;
  const int64_t nchunks = (L + kChunk - 1) / kChunk;
  auto phys = [](int i) { return REV ? kTok - 1 - i : i; };

  const T* __restrict__ zrow = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + E + chc) * a.ldxz;
  const T* __restrict__ drow = static_cast<const T*>(a.delta) + ((int64_t)job * E + chc) * a.ldd;
  const T* __restrict__ gorow = static_cast<const T*>(a.dout) + ((int64_t)job * E + chc) * a.ldo;
  const int64_t pc = (int64_t)pset * E + chc;
  const float dtb = a.dt_b[pc];
  float* my_acc = acc_s + warp * N;
  float* my_a2 = a2_s + warp * N;
  if (lane < N) {
    my_a2[lane] = a.A2[pc * N + lane];
    my_acc[lane] = 0.f;
  }
  __syncwarp();
  unsigned alive = active ? ((1u << N) - 1u) : 0u;      // states still receiving contributions (warp-uniform)

  const int seg = REV ? 31 - lane : lane;
  uint32_t poff[4];
  tile_piece_offsets(seg, poff);
  const int c_row = job * 2 * N + N;                   // the C rows of this job
  const int blocks_per_chunk = kChunk / kBlkTok;
  float cum_base = 0.f;
  uint32_t parity = 0;

  __syncthreads();
  bool any = __syncthreads_or(alive != 0);
  if (any && threadIdx.x == 0) {
    const int64_t first = REV ? nchunks - 1 : 0;
    mbar_expect_tx(bar, N * kChunk * 4);
    tma_load_3d(tile, tmap, 0, (int)(first * blocks_per_chunk), c_row, bar);
  }

  for (int64_t c = 0; any && c < nchunks; ++c) {
    const int64_t pcidx = REV ? nchunks - 1 - c : c;
    const int64_t tseg = pcidx * kChunk + (int64_t)seg * kTok;
    const bool seg_in = tseg < L;
    for (int n = 0; n < N; ++n)
      if ((alive >> n) & 1u)
        if (my_a2[n] * cum_base < a.cutoff_log2) alive &= ~(1u << n);

    mbar_wait(bar, parity);
    parity ^= 1;
    if (alive) {
      float dr[kTok], gs[kTok], zs[kTok], cum[kTok], dy[kTok];
      if (seg_in) {
        load_vec<T, kTok>(drow + tseg, dr);
        load_vec<T, kTok>(gorow + tseg, gs);
        load_vec<T, kTok>(zrow + tseg, zs);
      } else {
#pragma unroll
        for (int i = 0; i < kTok; ++i) { dr[i] = 0.f; gs[i] = 0.f; zs[i] = 0.f; }
      }
      float run = 0.f;
#pragma unroll
      for (int i = 0; i < kTok; ++i) {
        float d = softplus(dr[phys(i)] + dtb);
        const bool masked = tseg + phys(i) >= L;
        if (masked) d = 0.f;
        run += d;
        cum[i] = run;
        dy[i] = masked ? 0.f : gs[phys(i)] * silu_io<T>(zs[phys(i)]);
      }
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
        const float A2n = my_a2[n];
        const uint32_t rowp = tile_s + n * (kChunk * 4);
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 q = lds128(rowp + poff[k]);
          const float cq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = REV ? kTok - 1 - (4 * k + e) : 4 * k + e;
            part = fmaf(cq[e] * dy[i], ex2(A2n * cum[i]), part);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) my_acc[n] += part;
      }
    }
    any = __syncthreads_or(alive != 0);                // also: everyone is done with the tile
    if (any && c + 1 < nchunks && threadIdx.x == 0) {
      const int64_t npc = REV ? pcidx - 1 : pcidx + 1;
      mbar_expect_tx(bar, N * kChunk * 4);
      tma_load_3d(tile, tmap, 0, (int)(npc * blocks_per_chunk), c_row, bar);
    }
  }
  __syncwarp();
  if (active && lane < N) a.dh[((int64_t)job * E + ch) * N + lane] = my_acc[lane];
}

template <typename T, int N>
__global__ void __launch_bounds__(kMaxG * 32, 2)
scan_adjoint_kernel(const cad_scan_adjoint_args a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tile = reinterpret_cast<float*>(base);
  float* acc_s = reinterpret_cast<float*>(base + (size_t)N * kChunk * 4);
  float* a2_s = acc_s + kMaxG * N;
  uint64_t* bar = reinterpret_cast<uint64_t*>(a2_s + kMaxG * N);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  const int job = blockIdx.y;
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) adjoint_job<T, N, true>(a, &tmap, job, seq, pset, tile, acc_s, a2_s, bar);
  else     adjoint_job<T, N, false>(a, &tmap, job, seq, pset, tile, acc_s, a2_s, bar);
}

template <typename T, int N>
static int launch_adjoint(const cad_scan_adjoint_args& a, int G, cudaStream_t stream) {
  CUtensorMap tmap;
  if (make_row_tile_map(&tmap, a.bc, (int64_t)a.njobs * 2 * N, a.ldbc, a.L, N) != 0) return -1;
  const size_t smem = 1024 + (size_t)N * kChunk * 4 + (size_t)2 * kMaxG * N * sizeof(float) + 16;
  auto kern = scan_adjoint_kernel<T, N>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)((a.E + G - 1) / G), (unsigned)a.njobs);
  kern<<<grid, G * 32, smem, stream>>>(a, tmap);
  CAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace cad

extern "C" int cad_bimamba_scan_adjoint(const cad_scan_adjoint_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_bimamba_scan_adjoint: null argument block");
  CAD_REQUIRE(a->L > 0 && a->E > 0 && a->njobs > 0 && a->nseq > 0, "cad_bimamba_scan_adjoint: bad sizes");
  CAD_REQUIRE(a->xz && a->delta && a->bc && a->dout && a->dt_b && a->A2 && a->dh && a->seq_of_job && a->pset_of_job &&
              a->rev_of_job, "cad_bimamba_scan_adjoint: null pointer");
  CAD_REQUIRE(a->N == 16, "cad_bimamba_scan_adjoint: d_state = %lld not built (only 16)", (long long)a->N);
  CAD_REQUIRE(a->ldxz % 16 == 0 && a->ldd % 16 == 0 && a->ldo % 16 == 0 && a->ldbc % 32 == 0,
              "cad_bimamba_scan_adjoint: bad row pitches");
  CAD_REQUIRE(a->cutoff_log2 < 0.f, "cad_bimamba_scan_adjoint: cutoff_log2 must be negative");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int G = a->channels_per_cta;
  if (G <= 0) G = kMaxG;
  CAD_REQUIRE(G >= 1 && G <= kMaxG, "cad_bimamba_scan_adjoint: channels_per_cta must be in [1, %d]", kMaxG);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, return launch_adjoint<T, 16>(*a, G, stream));
  return 0;
}

// Scan variant 20 (lane = channel, scan_fwd_v20.cuh): kernel entry, launcher, and the two helpers of the in-GPU segment
// split — the token-major copy of the B / C rows and the composition of the segment carries (SURVEY.md §8e algebra, applied
// inside one GPU).  Reached through cad_bimamba_scan_fwd with variant = 20 (scan_fwd.cu).
#include "scan_fwd_v20.cuh"

namespace cad {

template <typename T>
__global__ void __launch_bounds__(v20::kMaxW * 32, 2)
bimamba_scan_fwd_v20_kernel(const cad_scan_fwd_args a) {
  extern __shared__ unsigned char smem_raw[];
  v20::kernel_body<T>(a, smem_raw);
}

int launch_scan_v20(const cad_scan_fwd_args& a, cudaStream_t stream) {
  const int nseg = a.nseg > 0 ? a.nseg : 1;
  CAD_REQUIRE(a.io_dtype != CAD_F32, "cad_bimamba_scan_fwd: variant 20 needs 16-bit I/O");
  CAD_REQUIRE(!a.h0 && !a.hlast && !a.dtsum && !a.chunk_state && !a.state_only,
              "cad_bimamba_scan_fwd: variant 20 covers inference only (no chunk_state / state_only; carry-in, end state and "
              "sum dt come from cad_seg_carry on the segment outputs, not from this launch)");
  CAD_REQUIRE(a.bcT && aligned16(a.bcT), "cad_bimamba_scan_fwd: variant 20 needs bcT (cad_bc_transpose), 16-byte aligned");
  CAD_REQUIRE(nseg <= 4096, "cad_bimamba_scan_fwd: nseg out of range");
  CAD_REQUIRE(a.L < (int64_t(1) << 31) - 4096, "cad_bimamba_scan_fwd: variant 20 keeps token-group counters in 32 bits");
  CAD_REQUIRE(nseg == 1 || (a.seg_state && a.seg_dtsum), "cad_bimamba_scan_fwd: nseg > 1 needs seg_state and seg_dtsum");
  CAD_REQUIRE(!a.seg_state || (aligned16(a.seg_state) && a.seg_dtsum), "cad_bimamba_scan_fwd: seg_state must be 16-byte "
              "aligned and come with seg_dtsum");
  // warps per CTA (32 channels each): as many as divide the channel groups evenly, at most kMaxW
  int W = a.channels_per_cta;            // here: WARPS per CTA
  const int64_t ngroups = (a.E + 31) / 32;
  if (W <= 0) {
    const int sms = cad_sm_count() > 0 ? cad_sm_count() : 148;
    long best = -1;
    for (int w = 1; w <= v20::kMaxW; ++w) {   // two CTAs per SM: minimise the busiest SM's warp count, prefer larger CTAs
      const long ctas = (long)a.njobs * nseg * ((ngroups + w - 1) / w);
      const long cost = ((ctas + 2L * sms - 1) / (2L * sms)) * w;
      if (best < 0 || cost <= best) { best = cost; W = w; }
    }
  }
  CAD_REQUIRE(W >= 1 && W <= v20::kMaxW, "cad_bimamba_scan_fwd: channels_per_cta (warps per CTA for variant 20) must be in "
              "[1, %d]", v20::kMaxW);
  const size_t smem = v20::smem_bytes(W);
  void (*kern)(const cad_scan_fwd_args) = a.io_dtype == CAD_BF16 ? bimamba_scan_fwd_v20_kernel<__nv_bfloat16> : bimamba_scan_fwd_v20_kernel<__half>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(v20): %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)((ngroups + W - 1) / W), (unsigned)a.njobs, (unsigned)nseg);
  kern<<<grid, W * 32, smem, stream>>>(a);
  CAD_LAUNCH_CHECK();
  return 0;
}

// bc (njobs, N2, ldbc) -> bcT (njobs, Lp, N2), Lp = ceil256(L); rows [L, Lp) are zero.  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) bc_transpose_kernel(const float* __restrict__ bc, float* __restrict__ bcT, int64_t N2,
                                                           int64_t L, int64_t Lp, int64_t ldbc) {
  __shared__ float tile[32][33];
  const int64_t job = blockIdx.z, t0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8 threads
  for (int r = ty; r < 32; r += 8) {
    const int64_t row = r0 + r, t = t0 + tx;
    tile[r][tx] = (row < N2 && t < L) ? bc[(job * N2 + row) * ldbc + t] : 0.f;
  }
  __syncthreads();
  for (int tt = ty; tt < 32; tt += 8) {
    const int64_t t = t0 + tt, row = r0 + tx;
    if (t < Lp && row < N2) bcT[(job * Lp + t) * N2 + row] = tile[tx][tt];
  }
}

// carry[j, s, e, n]: state entering logical segment s.  One thread per (job, channel, state); nseg sequential steps.
__global__ void __launch_bounds__(256) seg_carry_kernel(const float* __restrict__ seg_state, const float* __restrict__ seg_dtsum,
                                                        const float* __restrict__ A2, const int32_t* __restrict__ pset_of_job,
                                                        const float* __restrict__ h0, float* __restrict__ carry,
                                                        float* __restrict__ hlast, float* __restrict__ dtsum, int64_t njobs,
                                                        int64_t nseg, int64_t E) {
  constexpr int N = 16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= njobs * E * N) return;
  const int64_t n = i % N, e = (i / N) % E, j = i / (N * E);
  const float a2 = A2[((int64_t)pset_of_job[j] * E + e) * N + n];
  float h = h0 ? h0[(j * E + e) * N + n] : 0.f, dsum = 0.f;
  for (int64_t s = 0; s < nseg; ++s) {
    const int64_t row = (j * nseg + s) * E + e;
    if (carry) carry[row * N + n] = h;
    const float ds = seg_dtsum[row];
    dsum += ds;
    h = fmaf(ex2(a2 * ds), h, seg_state[row * N + n]);
  }
  if (hlast) hlast[(j * E + e) * N + n] = h;
  if (dtsum && n == 0) dtsum[j * E + e] = dsum;
}

// carry[j, s, e, n] = h0[j, e, n] * exp2(A2 * sum of dt over the logical segments before s), exactly 0 once the exponent is below the
// cut-off.  One thread per (job, channel, state); a segment = per512 physical 512-token chunks, whose sums of dt the time-parallel
// scan wrote in LOGICAL chunk order (physical chunk pc of a reversed job = logical chunk nchunks - 1 - pc).
__global__ void __launch_bounds__(256) shard_seg_carry_kernel(const float* __restrict__ chunk_dtsum, const float* __restrict__ A2,
                                                              const int32_t* __restrict__ pset_of_job, const int32_t* __restrict__ rev_of_job,
                                                              const float* __restrict__ h0, float* __restrict__ carry, int64_t njobs,
                                                              int64_t E, int64_t nchunks, int64_t nseg, int64_t per512, float cutoff_log2) {
  constexpr int N = 16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= njobs * E * N) return;
  const int64_t n = i % N, e = (i / N) % E, j = i / (N * E);
  const float a2 = A2[((int64_t)pset_of_job[j] * E + e) * N + n];
  const bool rev = rev_of_job[j] != 0;
  const float* cd = chunk_dtsum + (j * E + e) * nchunks;
  const float h = h0[(j * E + e) * N + n];
  float cum = 0.f;
  for (int64_t sl = 0; sl < nseg; ++sl) {
    const float x = a2 * cum;
    carry[((j * nseg + sl) * E + e) * N + n] = (sl == 0) ? h : (x < cutoff_log2 ? 0.f : h * ex2(x));
    const int64_t k = rev ? nseg - 1 - sl : sl;                 // physical block of this logical segment
    const int64_t p0 = k * per512, p1 = min((k + 1) * per512, nchunks);
    for (int64_t pc = p0; pc < p1; ++pc) cum += cd[rev ? nchunks - 1 - pc : pc];
  }
}

}  // namespace cad

extern "C" int cad_shard_seg_carry(const float* chunk_dtsum, const float* A2, const int32_t* pset_of_job, const int32_t* rev_of_job,
                                   const float* h0, float* carry, int64_t njobs, int64_t E, int64_t nchunks, int64_t nseg,
                                   int64_t per512, float cutoff_log2, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(chunk_dtsum && A2 && pset_of_job && rev_of_job && h0 && carry, "cad_shard_seg_carry: null pointer");
  CAD_REQUIRE(njobs > 0 && E > 0 && nchunks > 0 && nseg > 0 && per512 > 0 && (nseg - 1) * per512 < nchunks && nseg * per512 >= nchunks,
              "cad_shard_seg_carry: nseg segments of per512 chunks must tile the nchunks chunks");
  CAD_REQUIRE(cutoff_log2 < 0.f, "cad_shard_seg_carry: cutoff_log2 must be negative");
  const int64_t n = njobs * E * 16;
  shard_seg_carry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      chunk_dtsum, A2, pset_of_job, rev_of_job, h0, carry, njobs, E, nchunks, nseg, per512, cutoff_log2);
  CAD_LAUNCH_CHECK();
  return 0;
}

extern "C" int cad_bc_transpose(const float* bc, float* bcT, int64_t njobs, int64_t N2, int64_t L, int64_t ldbc, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(bc && bcT && njobs > 0 && N2 > 0 && L >= 0 && ldbc >= L, "cad_bc_transpose: bad arguments");
  if (L == 0) return 0;
  const int64_t Lp = (L + 255) / 256 * 256;
  dim3 grid((unsigned)(Lp / 32), (unsigned)((N2 + 31) / 32), (unsigned)njobs);
  bc_transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(bc, bcT, N2, L, Lp, ldbc);
  CAD_LAUNCH_CHECK();
  return 0;
}

extern "C" int cad_seg_carry(const float* seg_state, const float* seg_dtsum, const float* A2, const int32_t* pset_of_job,
                             const float* h0, float* carry, float* hlast, float* dtsum, int64_t njobs, int64_t nseg, int64_t E,
                             void* stream_) {
  using namespace cad;
  CAD_REQUIRE(seg_state && seg_dtsum && A2 && pset_of_job && (carry || hlast || dtsum) && njobs > 0 && nseg > 0 && E > 0,
              "cad_seg_carry: bad arguments");
  const int64_t n = njobs * E * 16;
  seg_carry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      seg_state, seg_dtsum, A2, pset_of_job, h0, carry, hlast, dtsum, njobs, nseg, E);
  CAD_LAUNCH_CHECK();
  return 0;
}

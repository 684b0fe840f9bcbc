// Carry fix-up of a zero-carry scan, sm_100a — used twice on the path:
//   * sequence sharding over GPUs: a shard scanned from a ZERO state lacks the contribution of its carry-in h0;
//   * scan variant 20 (lane = channel): one GPU cuts the sequence into segments, every segment is scanned from zero and the
//     composed carries (cad_seg_carry) are applied here, once per (job, logical segment).
//     y_true[t] = y_zero[t] + silu(z[t]) * sum_n C[t,n] * exp2(A2[n] * cumdt[t]) * h0[n],     cumdt[t] = sum_{s<=t} dt[s]
// (the recurrence is affine in the state; the product of the decays over a prefix is exp2(A2 * cumdt)).  The factor only
// decays, so a (channel, state) pair is dropped for good once A2 * cumdt < cutoff_log2 and a warp retires when all 16 are gone.
//
// Second design of this kernel (profiles/r2_fixup_ncu_summary.txt: the first one — a CTA of 7 channels around ONE shared TMA tile
// of the C rows, a __syncthreads_or per 512-token chunk — spent 77 % of its stall samples in that barrier and ran at 29 % of the
// MUFU pipe, 1.08 ms for a Caduceus-PS launch at 37 segments).  Now: ONE WARP PER (job, segment, channel), no shared memory, no
// CTA-level synchronisation at all.  A lane owns 8 consecutive tokens of a 256-token step; the C rows of the states still alive
// come straight from global memory (32 coalesced bytes per lane and state; the warps of a CTA work on neighbouring channels of
// the same segment, so these reads hit L2 mostly) through a per-warp cp.async ring, three states ahead of the one being consumed;
// the next step's dt_raw is requested before the state loop.  CTAs are 4 warps at 64 registers (8 per SM): warps retire
// independently, and a slow channel pins at most three retired neighbours' slots.
#include "scan_common.cuh"

namespace cad {
namespace fx {

constexpr int FT = 8;                // tokens per lane and step
constexpr int FCH = 32 * FT;         // tokens per warp step
constexpr int kRing = 4;             // cp.async stages per warp (1 KB each: 32 lanes x 32 bytes of one C row)
constexpr int kMaxW = 4;             // warps (channels) per CTA: small CTAs, so that a slow channel does not pin the slots of retired warps

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc));   // .ca: the CTA's warps share the C rows through L1
}

template <typename T> struct Raw8;                                   // 8 elements as they lie in memory
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__half> { uint4 a; };
template <> struct Raw8<__nv_bfloat16> { uint4 a; };

template <typename T>
__device__ __forceinline__ Raw8<T> ld8(const T* p, bool nc) {
  Raw8<T> r;
  if constexpr (sizeof(T) == 4) {
    const float4* q = reinterpret_cast<const float4*>(p);
    if (nc) { r.a = __ldg(q); r.b = __ldg(q + 1); } else { r.a = q[0]; r.b = q[1]; }
  } else {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    r.a = nc ? __ldg(q) : q[0];
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void unpack8(const Raw8<T>& r, float (&v)[FT]) {
  if constexpr (sizeof(T) == 4) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  } else {
    const T* e = reinterpret_cast<const T*>(&r.a);
#pragma unroll
    for (int i = 0; i < FT; ++i) v[i] = io<T>::to_f(e[i]);
  }
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&v)[FT]) {
  if constexpr (sizeof(T) == 4) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < FT; ++i) e[i] = io<T>::from_f(v[i]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
}

// one warp: channel `ch` of `job`, tokens [t_off, t_off + Lr) (the whole shard, or one in-GPU segment; t_off % 256 == 0)
template <typename T, int N, bool REV>
__device__ __forceinline__ void fixup_warp(const cad_scan_fixup_args& a, int job, int seq, int pset, int64_t t_off, int64_t Lr64,
                                           const float* __restrict__ h0_base, int64_t ch, uint32_t ring_base) {
  const int lane = threadIdx.x & 31;
  const uint32_t ring = ring_base + (uint32_t)(threadIdx.x >> 5) * (kRing * 1024u + 2048u) + (uint32_t)lane * 16u;
  const int64_t E = a.E;
  const int64_t pc = (int64_t)pset * E + ch;
  // lane n < 16 keeps (A2[n], h0[n]) of this channel; broadcast per state with a shuffle
  const float my_a2 = a.A2[pc * N + (lane & (N - 1))];
  const float my_h0 = h0_base[ch * N + (lane & (N - 1))];
  unsigned alive = __ballot_sync(0xffffffffu, lane < N && my_h0 != 0.f);    // a zero carry never contributes
  if (!alive) return;

  const T* __restrict__ zrow = static_cast<const T*>(a.xz) + ((int64_t)seq * 2 * E + E + ch) * a.ldxz + t_off;
  const T* __restrict__ drow = static_cast<const T*>(a.delta) + ((int64_t)job * E + ch) * a.ldd + t_off;
  T* __restrict__ orow = static_cast<T*>(a.out) + ((int64_t)job * E + ch) * a.ldo + t_off;
  const float* __restrict__ crow = a.bc + ((int64_t)job * 2 * N + N) * a.ldbc + t_off;      // C rows of this job
  const float dtb = a.dt_b[pc];
  const int Lr = (int)Lr64;                                   // a shard / segment holds < 2^31 tokens (checked by the launcher)
  const int nsteps = (Lr + FCH - 1) / FCH;
  const int pl = REV ? 31 - lane : lane;                      // physical lane slot; the lane index is the LOGICAL order
  auto tok0 = [&](int c) { return (REV ? nsteps - 1 - c : c) * FCH + pl * FT; };

  float cum_base = 0.f;
  int ts = tok0(0);
  Raw8<T> d_raw;
  if (ts < Lr) d_raw = ld8<T>(drow + ts, true);
  for (int c = 0; c < nsteps; ++c) {
    // drop the states that have decayed away before this step starts (warp-uniform)
    {
      const bool dead = my_a2 * cum_base < a.cutoff_log2;
      alive &= ~__ballot_sync(0xffffffffu, lane < N && dead);
      if (!alive) return;
    }
    const bool in = ts < Lr;
    float dr[FT];
    if (in) unpack8<T>(d_raw, dr);
    // z / out of this step (into this warp's two extra ring slots: no registers) and dt_raw of the next are requested now and
    // fly during the state loop; they belong to the first cp.async group of the step
    if (in) {
      constexpr int PIECES = (int)sizeof(T) * FT / 16;         // 16-byte pieces per lane and row: 1 (16-bit) or 2 (fp32)
#pragma unroll
      for (int h = 0; h < PIECES; ++h) {
        cp_async16(ring + kRing * 1024u + 512u * h, reinterpret_cast<const char*>(zrow + ts) + 16 * h);
        cp_async16(ring + kRing * 1024u + 1024u + 512u * h, reinterpret_cast<const char*>(orow + ts) + 16 * h);
      }
    }
    const int ts_next = c + 1 < nsteps ? tok0(c + 1) : Lr;
    if (ts_next < Lr) d_raw = ld8<T>(drow + ts_next, true);

    float cum[FT], y[FT];                                     // logical order inside the lane
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < FT; ++i) {
      const int p = REV ? FT - 1 - i : i;
      float d = in ? softplus(dr[p] + dtb) : 0.f;
      if (ts + p >= Lr) d = 0.f;
      run += d;
      cum[i] = run;
      y[i] = 0.f;
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
    for (int i = 0; i < FT; ++i) cum[i] += base;

    // states still alive: their C rows (32 bytes per lane and state) travel through this warp's cp.async ring, kRing - 1 states
    // ahead of the one being consumed — no registers held by loads in flight, and the L2 latency of a row is covered by the
    // exp2 work of three states
    unsigned mi = alive, mu = alive;
    int issued = 0, used = 0;
    auto issue = [&]() {                                      // every call commits (uniform cp.async group counting)
      if (mi) {
        const int ni = __ffs(mi) - 1;
        mi &= mi - 1;
        if (in) {
          const float* src = crow + (int64_t)ni * a.ldbc + ts;
          const uint32_t dst = ring + (uint32_t)(issued & (kRing - 1)) * 1024u;
          cp_async16(dst, src);
          cp_async16(dst + 512u, src + 4);
        }
      }
      asm volatile("cp.async.commit_group;");
      ++issued;
    };
#pragma unroll
    for (int k = 0; k < kRing - 1; ++k) issue();
    while (mu) {
      issue();
      asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 1) : "memory");
      const int n = __ffs(mu) - 1;
      mu &= mu - 1;
      const uint32_t st = ring + (uint32_t)(used & (kRing - 1)) * 1024u;
      ++used;
      float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
      if (in) { c0 = lds128(st); c1 = lds128(st + 512u); }
      const float A2n = __shfl_sync(0xffffffffu, my_a2, n), hn = __shfl_sync(0xffffffffu, my_h0, n);
      const float cq[FT] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < FT; ++i) {
        const int p = REV ? FT - 1 - i : i;
        y[i] = fmaf(cq[p] * hn, ex2(A2n * cum[i]), y[i]);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // (empty groups only) before the ring is reused by the next step
    if (in) {
      Raw8<T> z_raw, o_raw;
      if constexpr (sizeof(T) == 4) {
        z_raw.a = lds128(ring + kRing * 1024u); z_raw.b = lds128(ring + kRing * 1024u + 512u);
        o_raw.a = lds128(ring + kRing * 1024u + 1024u); o_raw.b = lds128(ring + kRing * 1024u + 1536u);
      } else {
        const float4 zq = lds128(ring + kRing * 1024u), oq = lds128(ring + kRing * 1024u + 1024u);
        z_raw.a = *reinterpret_cast<const uint4*>(&zq); o_raw.a = *reinterpret_cast<const uint4*>(&oq);
      }
      float zs[FT], os[FT];
      unpack8<T>(z_raw, zs);
      unpack8<T>(o_raw, os);
#pragma unroll
      for (int i = 0; i < FT; ++i) {
        const int p = REV ? FT - 1 - i : i;
        os[p] = fmaf(y[i], silu_io<T>(zs[p]), os[p]);
      }
      if (ts + FT <= Lr) {
        st8<T>(orow + ts, os);
      } else {
#pragma unroll
        for (int i = 0; i < FT; ++i)
          if (ts + i < Lr) orow[ts + i] = io<T>::from_f(os[i]);
      }
    }
    ts = ts_next;
  }
}

template <typename T, int N>
__global__ void __launch_bounds__(kMaxW * 32, 8) scan_fixup_kernel(const cad_scan_fixup_args a) {
  __shared__ __align__(16) unsigned char ring_s[kMaxW * (kRing * 1024 + 2048)];     // per warp: kRing C rows + z + out of the step
  const uint32_t ring_base = smem_u32(ring_s);
  const int64_t ch = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= a.E) return;                                        // warp-uniform; there is no CTA-level synchronisation
  int job = blockIdx.y;
  int64_t t_off = 0, Lr = a.L;
  const float* h0_base;
  if (a.nseg > 1) {
    // grid.y = (job, logical segment); the segment's tokens = physical block k of the split used by scan variant 20
    // (scan_fwd_v20.cuh::block_range: whole 256-token chunks, ceil(nchunks / nseg) per block)
    const int first = a.seg_first ? 0 : 1, cnt = a.nseg - first;      // logical segments [first, nseg) get a carry term
    const int sl = first + (int)(blockIdx.y % cnt);
    job = (int)(blockIdx.y / cnt);
    const int64_t k = a.rev_of_job[job] ? a.nseg - 1 - sl : sl;
    const int64_t nch = (a.L + 255) / 256, per = (nch + a.nseg - 1) / a.nseg;
    int64_t lo = k * per * 256, hi = (k + 1) * per * 256;
    if (hi > a.L) hi = a.L;
    if (lo >= hi) return;                                       // empty block
    t_off = lo; Lr = hi - lo;
    h0_base = a.seg_carry + ((int64_t)job * a.nseg + sl) * a.E * N;
  } else {
    h0_base = a.h0 + (int64_t)job * a.E * N;
  }
  const int seq = a.seq_of_job[job], pset = a.pset_of_job[job], rev = a.rev_of_job[job];
  if (rev) fixup_warp<T, N, true>(a, job, seq, pset, t_off, Lr, h0_base, ch, ring_base);
  else     fixup_warp<T, N, false>(a, job, seq, pset, t_off, Lr, h0_base, ch, ring_base);
}

template <typename T, int N>
static int launch_fixup(const cad_scan_fixup_args& a, int W, cudaStream_t stream) {
  dim3 grid((unsigned)((a.E + W - 1) / W), (unsigned)(a.nseg > 1 ? a.njobs * (a.nseg - (a.seg_first ? 0 : 1)) : a.njobs));
  scan_fixup_kernel<T, N><<<grid, W * 32, 0, stream>>>(a);
  CAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace fx
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
  CAD_REQUIRE(aligned16(a->xz) && aligned16(a->delta) && aligned16(a->out) && aligned16(a->bc),
              "cad_bimamba_scan_fixup: xz / delta / out / bc must be 16-byte aligned");
  CAD_REQUIRE(a->cutoff_log2 < 0.f, "cad_bimamba_scan_fixup: cutoff_log2 must be negative");
  CAD_REQUIRE(a->L < (int64_t(1) << 31) - 4096, "cad_bimamba_scan_fixup: token offsets inside a shard are kept in 32 bits");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int W = a->channels_per_cta;
  if (W <= 0) W = fx::kMaxW;
  CAD_REQUIRE(W >= 1 && W <= fx::kMaxW, "cad_bimamba_scan_fixup: channels_per_cta must be in [1, %d]", fx::kMaxW);
  CAD_DISPATCH_DTYPE(a->io_dtype, T, return fx::launch_fixup<T, 16>(*a, W, stream));
  return 0;
}

// Sequence sharding over the GPUs of one NVSwitch box WITHOUT a library collective on the data path:
// the two per-layer exchanges of the sharded BiMamba call — the 3-sample conv halo and the (E x N) boundary state with its
// sum(dt) — are tiny (6 KB and 139 KB per neighbour for Caduceus-PS) and latency-bound, so they are done by these kernels with
// plain stores into the PEERS' memory over NVLink (symmetric workspace: every rank allocates the same layout and knows every
// peer's base address), a system-scope release/acquire flag per (exchange, sender), and the carry composition fused behind the
// wait.  No host round trip, no NCCL launch, and — unlike a collective — capturable in ONE CUDA graph with the compute around it,
// which is what makes an 8-way split of a 131k-token forward (0.4 ms of GPU work per layer and rank) scale.
//
// Protocol (per exchange kind k, per rank):
//   epoch_k   device counter, +1 per call, the same on every rank because all ranks issue the same call sequence;
//   buffers   double-buffered by epoch parity: a sender can be at most ONE call ahead of a receiver (it cannot finish call e + 1
//             before the receiver has pushed call e + 1, which the receiver does only after it has read call e), so parity e + 1
//             never overwrites data of call e that is still being read;
//   flags     flag[k][sender] on the RECEIVER holds the sender's last completed epoch; written with st.release.sys after a
//             __threadfence_system() that follows the data stores, read with ld.acquire.sys in a spin loop that traps after 10 s
//             (a lost peer fails the launch instead of hanging the GPU).  Data written by peers is read with volatile loads.
#include "common.cuh"

namespace cad {
namespace peer {

constexpr int kCtrlBytes = 1024;        // ctrl block at the start of every workspace
constexpr int kMaxWorld = 16;
// ctrl words (uint32):
//   [0] halo epoch   [1] carry epoch   [2] carry push arrival counter   [3] adjoint epoch (reserved)
//   [16 + 0] halo flag "from left" (sender rank - 1)   [16 + 1] halo flag "from right" (sender rank + 1)
//   [32 + r] carry flag of sender r
constexpr int kHaloFlag = 16, kCarryFlag = 32;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t epoch) {
  if ((int32_t)(ld_acquire_sys(p) - epoch) >= 0) return;
  uint64_t t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 10000000000ull) __trap();
  }
}
template <typename T> __device__ __forceinline__ T ld_vol(const T* p) { return *reinterpret_cast<const volatile T*>(p); }

struct Layout {          // byte offsets inside a workspace; identical on every rank
  int64_t halo;          // [2 parity][2 sides: from-left, from-right][nseq_max * E * 3] 4-byte slots (io elements widened on the wire)
  int64_t gather;        // [2 parity][world][njobs_max * (E * N + E)] fp32
  int64_t halo_elems, gather_elems, total;
};
__host__ __device__ inline Layout layout(int world, int64_t nseq_max, int64_t njobs_max, int64_t E, int64_t N) {
  Layout l;
  l.halo_elems = nseq_max * E * 3;
  l.gather_elems = njobs_max * (E * N + E);
  l.halo = kCtrlBytes;
  l.gather = l.halo + 2 * 2 * l.halo_elems * 4;
  l.gather = (l.gather + 255) / 256 * 256;
  l.total = l.gather + 2 * (int64_t)world * l.gather_elems * 4;
  return l;
}

struct Ctx {
  const uint64_t* peer_ws;     // device array (world): base address of every rank's workspace as seen from this device
  int32_t rank, world;
  int64_t nseq_max, njobs_max, E, N;
};

// ---- conv halo: push my edge samples to both neighbours, wait for theirs, assemble halo (njobs, E, 3) ------------------
template <typename T>
__global__ void __launch_bounds__(1024) halo_kernel(Ctx c, const T* __restrict__ xz, int64_t ldxz, int64_t L, int nseq, int njobs,
                                                    const int32_t* __restrict__ seq_of_job, const int32_t* __restrict__ rev_of_job,
                                                    T* __restrict__ halo) {
  __shared__ uint32_t s_epoch;
  const Layout lay = layout(c.world, c.nseq_max, c.njobs_max, c.E, c.N);
  unsigned char* me = reinterpret_cast<unsigned char*>(c.peer_ws[c.rank]);
  uint32_t* ctrl = reinterpret_cast<uint32_t*>(me);
  if (threadIdx.x == 0) s_epoch = ctrl[0] + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int par = epoch & 1;
  const int64_t E = c.E, n = (int64_t)nseq * E * 3;
  const bool has_l = c.rank > 0, has_r = c.rank + 1 < c.world;
  // slot (parity, side) of a workspace
  auto slot = [&](unsigned char* ws, int side) { return reinterpret_cast<float*>(ws + lay.halo) + ((int64_t)par * 2 + side) * lay.halo_elems; };
  float* to_l = has_l ? slot(reinterpret_cast<unsigned char*>(c.peer_ws[c.rank - 1]), 1) : nullptr;   // I am their right neighbour
  float* to_r = has_r ? slot(reinterpret_cast<unsigned char*>(c.peer_ws[c.rank + 1]), 0) : nullptr;   // I am their left neighbour
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t s = i / (E * 3), e = (i / 3) % E, k = i % 3;
    const T* row = xz + (s * 2 * E + e) * ldxz;                       // x rows are [0, E) of every sequence
    if (to_l) to_l[i] = io<T>::to_f(row[k]);                          // my FIRST three
    if (to_r) to_r[i] = io<T>::to_f(row[L - 3 + k]);                  // my LAST three
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (has_l) st_release_sys(reinterpret_cast<uint32_t*>(c.peer_ws[c.rank - 1]) + kHaloFlag + 1, epoch);
    if (has_r) st_release_sys(reinterpret_cast<uint32_t*>(c.peer_ws[c.rank + 1]) + kHaloFlag + 0, epoch);
    if (has_l) wait_flag(ctrl + kHaloFlag + 0, epoch);
    if (has_r) wait_flag(ctrl + kHaloFlag + 1, epoch);
    ctrl[0] = epoch;
  }
  __syncthreads();
  const float* from_l = slot(me, 0);
  const float* from_r = slot(me, 1);
  const int64_t m = (int64_t)njobs * E * 3;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
    const int64_t j = i / (E * 3), e = (i / 3) % E, k = i % 3;
    const int64_t s = seq_of_job[j];
    float v = 0.f;
    if (!rev_of_job[j]) { if (has_l) v = ld_vol(from_l + (s * E + e) * 3 + k); }          // predecessor's last three, logical order
    else                { if (has_r) v = ld_vol(from_r + (s * E + e) * 3 + (2 - k)); }    // successor's first three, reversed
    halo[i] = io<T>::from_f(v);
  }
}

// ---- boundary state: push (H, sum dt) of my zero-carry scan into slot[rank] of every rank's gather buffer -------------------
__global__ void __launch_bounds__(256) carry_push_kernel(Ctx c, const float* __restrict__ hlast, const float* __restrict__ dtsum,
                                                         int njobs) {
  const Layout lay = layout(c.world, c.nseq_max, c.njobs_max, c.E, c.N);
  uint32_t* ctrl = reinterpret_cast<uint32_t*>(c.peer_ws[c.rank]);
  const uint32_t epoch = ctrl[1] + 1;                 // incremented by the last CTA below, after every CTA has read it? no: see (*)
  const int par = epoch & 1;
  const int64_t nh = (int64_t)njobs * c.E * c.N, nd = (int64_t)njobs * c.E;
  const int64_t nh4 = nh / 4, nd4 = nd / 4;            // E is a multiple of 4 (checked by the launcher)
  const float4* h4 = reinterpret_cast<const float4*>(hlast);
  const float4* d4 = reinterpret_cast<const float4*>(dtsum);
  for (int r = 0; r < c.world; ++r) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(c.peer_ws[r]) + lay.gather) +
                  (((int64_t)par * c.world + c.rank) * lay.gather_elems) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nh4 + nd4; i += (int64_t)gridDim.x * blockDim.x)
      dst[i] = i < nh4 ? h4[i] : d4[i - nh4];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    // (*) every CTA reads ctrl[1] before it arrives; the last to arrive is the only writer
    if (atomicAdd(ctrl + 2, 1u) == gridDim.x - 1) {
      ctrl[2] = 0;
      __threadfence_system();
      for (int r = 0; r < c.world; ++r)
        if (r != c.rank) st_release_sys(reinterpret_cast<uint32_t*>(c.peer_ws[r]) + kCarryFlag + c.rank, epoch);
      ctrl[1] = epoch;
    }
  }
}

// ---- wait for every rank's (H, sum dt), compose this rank's carry-in:  h <- exp2(A2 * sum dt_r) * h + H_r over the logical
//      predecessors r (ranks < rank for left-to-right jobs, ranks > rank in descending order for right-to-left jobs) --------------
__global__ void __launch_bounds__(256) carry_compose_kernel(Ctx c, const float* __restrict__ A2, const int32_t* __restrict__ pset_of_job,
                                                            const int32_t* __restrict__ rev_of_job, int njobs, float* __restrict__ h0,
                                                            float* __restrict__ dtsum_all) {
  const Layout lay = layout(c.world, c.nseq_max, c.njobs_max, c.E, c.N);
  unsigned char* me = reinterpret_cast<unsigned char*>(c.peer_ws[c.rank]);
  const uint32_t* ctrl = reinterpret_cast<const uint32_t*>(me);
  const uint32_t epoch = ld_vol(ctrl + 1);             // set by this rank's push kernel (stream order)
  if (threadIdx.x < c.world && (int)threadIdx.x != c.rank) wait_flag(ctrl + kCarryFlag + threadIdx.x, epoch);
  __syncthreads();
  const int par = epoch & 1;
  const float* g = reinterpret_cast<const float*>(me + lay.gather) + (int64_t)par * c.world * lay.gather_elems;
  const int64_t E = c.E, N = c.N, nh = (int64_t)njobs * E * N;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nh) {
    const int64_t n = i % N, e = (i / N) % E, j = i / (N * E);
    const float a2 = A2[((int64_t)pset_of_job[j] * E + e) * N + n];
    float h = 0.f;
    if (!rev_of_job[j]) {
      for (int r = 0; r < c.rank; ++r) {
        const float* gr = g + (int64_t)r * lay.gather_elems;
        h = fmaf(ex2(a2 * ld_vol(gr + nh + j * E + e)), h, ld_vol(gr + i));
      }
    } else {
      for (int r = c.world - 1; r > c.rank; --r) {
        const float* gr = g + (int64_t)r * lay.gather_elems;
        h = fmaf(ex2(a2 * ld_vol(gr + nh + j * E + e)), h, ld_vol(gr + i));
      }
    }
    h0[i] = h;
  }
  if (dtsum_all) {                                      // (world, njobs, E): what the sharded backward needs
    const int64_t nd = (int64_t)njobs * E;
    for (int64_t k = i; k < (int64_t)c.world * nd; k += (int64_t)gridDim.x * blockDim.x)
      dtsum_all[k] = ld_vol(g + (k / nd) * lay.gather_elems + nh + k % nd);
  }
}

static int make_ctx(const cad_peer_ctx* p, Ctx* c) {
  CAD_REQUIRE(p && p->peer_ws, "cad_peer_*: null context");
  CAD_REQUIRE(p->world >= 1 && p->world <= kMaxWorld && p->rank >= 0 && p->rank < p->world, "cad_peer_*: bad rank / world (%d / %d)",
              p->rank, p->world);
  CAD_REQUIRE(p->E > 0 && p->E % 4 == 0 && p->N > 0 && p->nseq_max > 0 && p->njobs_max > 0, "cad_peer_*: bad sizes");
  c->peer_ws = static_cast<const uint64_t*>(p->peer_ws);
  c->rank = p->rank; c->world = p->world;
  c->nseq_max = p->nseq_max; c->njobs_max = p->njobs_max; c->E = p->E; c->N = p->N;
  return 0;
}

}  // namespace peer
}  // namespace cad

extern "C" int64_t cad_peer_ws_bytes(int32_t world, int64_t nseq_max, int64_t njobs_max, int64_t E, int64_t N) {
  if (world < 1 || world > cad::peer::kMaxWorld || nseq_max <= 0 || njobs_max <= 0 || E <= 0 || N <= 0) return -1;
  return cad::peer::layout(world, nseq_max, njobs_max, E, N).total;
}

extern "C" int cad_peer_halo_exchange(const cad_peer_ctx* p, const void* xz, int64_t ldxz, int64_t L, int32_t nseq, int32_t njobs,
                                      const int32_t* seq_of_job, const int32_t* rev_of_job, void* halo, int32_t io_dtype,
                                      void* stream_) {
  using namespace cad;
  peer::Ctx c;
  if (peer::make_ctx(p, &c)) return -1;
  CAD_REQUIRE(xz && halo && seq_of_job && rev_of_job, "cad_peer_halo_exchange: null pointer");
  CAD_REQUIRE(L >= 3, "cad_peer_halo_exchange: a sequence shard must hold at least 3 tokens (L = %lld)", (long long)L);
  CAD_REQUIRE(nseq > 0 && nseq <= c.nseq_max && njobs > 0 && njobs <= c.njobs_max, "cad_peer_halo_exchange: nseq / njobs exceed the workspace");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CAD_DISPATCH_DTYPE(io_dtype, T, (peer::halo_kernel<T><<<1, 1024, 0, stream>>>(
      c, static_cast<const T*>(xz), ldxz, L, nseq, njobs, seq_of_job, rev_of_job, static_cast<T*>(halo))));
  CAD_LAUNCH_CHECK();
  return 0;
}

extern "C" int cad_peer_carry_exchange(const cad_peer_ctx* p, const float* hlast, const float* dtsum, const float* A2,
                                       const int32_t* pset_of_job, const int32_t* rev_of_job, int32_t njobs, float* h0,
                                       float* dtsum_all, void* stream_) {
  using namespace cad;
  peer::Ctx c;
  if (peer::make_ctx(p, &c)) return -1;
  CAD_REQUIRE(hlast && dtsum && A2 && pset_of_job && rev_of_job && h0, "cad_peer_carry_exchange: null pointer");
  CAD_REQUIRE(njobs > 0 && njobs <= c.njobs_max, "cad_peer_carry_exchange: njobs exceeds the workspace");
  CAD_REQUIRE(aligned16(hlast) && aligned16(dtsum), "cad_peer_carry_exchange: hlast / dtsum must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t nh = (int64_t)njobs * c.E * c.N;
  peer::carry_push_kernel<<<32, 256, 0, stream>>>(c, hlast, dtsum, njobs);
  CAD_LAUNCH_CHECK();
  peer::carry_compose_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, stream>>>(c, A2, pset_of_job, rev_of_job, njobs, h0, dtsum_all);
  CAD_LAUNCH_CHECK();
  return 0;
}

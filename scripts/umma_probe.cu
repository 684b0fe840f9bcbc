// Hardware probe of the tcgen05 (UMMA) encodings used by csrc/xproj_umma.cu: shared-memory matrix descriptors without
// swizzle for an MN-major A operand (u[channel][token]) and K-major operands, the instruction descriptor, TMEM
// allocation, tcgen05.commit -> mbarrier, tcgen05.ld.  One CTA, results checked on the host against a float loop.
//
//   umma_probe <test> <mode>
//     test 0: D[128 x 48]  = A(MN-major, K = 64) . B(K-major)^T   four K = 16 instructions, second pass accumulates
//     test 1: D[128 x 128] = A(K-major,  K = 16) . B(K-major)^T
//     test 2: D[128 x 256] = A(K-major,  K = 16) . B(K-major)^T
//     test 3: D[128 x 128] = A(K-major,  K = 16) . B[256:384]^T   B = a 128-row chunk of a resident 512 x 16 operand
//             (K stride 8192 bytes, start address inside the operand: what the dt_proj instruction of the kernel sees);
//             read back with tcgen05.ld .x32
//     mode 0: descriptor LBO field = stride between core matrices along K, SBO field = stride along M/N   (expected)
//     mode 1: the two fields swapped
// Every wait has a watchdog (trap after 2 s); run each case as its own process under `timeout`.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  return d;                                     // base offset 0, layout type 0 = no swizzle
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  uint64_t t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try(bar, parity)) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 2000000000ull) __trap();
  }
}

struct Args {
  const __nv_bfloat16* A; const __nv_bfloat16* B; float* D;
  int M, N, K, a_mn_major, mode, passes, nb_total, n0, ld32;
};

// A global: a_mn_major ? [K][M] : [M][K];   B global: [N][K];   D: [M][N]
__global__ void __launch_bounds__(256, 1) probe_kernel(Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.M, N = a.N, K = a.K;
  unsigned char* As = smem;                                   // M*K*2 bytes
  unsigned char* Bs = smem + (size_t)M * K * 2;               // nb_total*K*2 bytes
  // core-matrix strides (bytes).  A: [kgroup][mgroup][8 rows][16 B]; B: [kgroup][ngroup][8][16 B]
  const uint32_t a_mn_stride = 128, a_k_stride = (uint32_t)(M / 8) * 128;
  const uint32_t b_mn_stride = 128, b_k_stride = (uint32_t)(a.nb_total / 8) * 128;
  for (int i = tid; i < M * K; i += blockDim.x) {
    int m, k;
    if (a.a_mn_major) { k = i / M; m = i - k * M; } else { m = i / K; k = i - m * K; }
    uint32_t off;
    if (a.a_mn_major) off = (m / 8) * a_mn_stride + (k / 8) * a_k_stride + (k % 8) * 16 + (m % 8) * 2;
    else              off = (m / 8) * a_mn_stride + (k / 8) * a_k_stride + (m % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(As + off) = a.A[i];
  }
  for (int i = tid; i < a.nb_total * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    const uint32_t off = (n / 8) * b_mn_stride + (k / 8) * b_k_stride + (n % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(Bs + off) = a.B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // instruction descriptor: D = f32, A = B = bf16, A major, B = K-major, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.a_mn_major ? 1 : 0) << 15) | (0u << 16) |
                         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    for (int pass = 0; pass < a.passes; ++pass)
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t a_addr = smem_u32(As) + ks * 2 * a_k_stride, b_addr = smem_u32(Bs) + ks * 2 * b_k_stride + (a.n0 / 8) * b_mn_stride;
        const uint64_t da = a.mode == 0 ? make_desc(a_addr, a_k_stride, a_mn_stride) : make_desc(a_addr, a_mn_stride, a_k_stride);
        const uint64_t db = a.mode == 0 ? make_desc(b_addr, b_k_stride, b_mn_stride) : make_desc(b_addr, b_mn_stride, b_k_stride);
        const uint32_t acc = (pass | ks) ? 1u : 0u;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait_wd(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // 8 warps: warp w reads lanes 32 (w % 4) .. +31, columns split between the two warp groups
  const int q = warp & 3, half = warp >> 2;
  const int row = 32 * q + lane;
  if (a.ld32) {
    for (int c0 = half * 32; c0 < N; c0 += 64) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                   "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                     "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                     "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) a.D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  } else
  for (int c0 = half * 16; c0 < N; c0 += 32) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) a.D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main(int argc, char** argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 0, mode = argc > 2 ? atoi(argv[2]) : 0;
  Args a{};
  a.M = 128; a.mode = mode;
  if (test == 0) { a.N = 48; a.K = 64; a.a_mn_major = 1; a.passes = 2; }
  else if (test == 1) { a.N = 128; a.K = 16; a.a_mn_major = 0; a.passes = 1; }
  else if (test == 2) { a.N = 256; a.K = 16; a.a_mn_major = 0; a.passes = 1; }
  else { a.N = 128; a.K = 16; a.a_mn_major = 0; a.passes = 1; a.nb_total = 512; a.n0 = 256; a.ld32 = 1; }
  if (a.nb_total == 0) a.nb_total = a.N;
  const int M = a.M, N = a.N, K = a.K, NB = a.nb_total;
  std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)NB * K);
  std::vector<float> fA((size_t)M * K), fB((size_t)NB * K);      // fA[m*K+k], fB[n*K+k]
  srand(1234 + test);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      const float v = (float)(rand() % 2001 - 1000) / 1000.f;
      const __nv_bfloat16 b = __float2bfloat16(v);
      fA[(size_t)m * K + k] = __bfloat162float(b);
      hA[a.a_mn_major ? (size_t)k * M + m : (size_t)m * K + k] = b;
    }
  for (size_t i = 0; i < hB.size(); ++i) {
    const float v = (float)(rand() % 2001 - 1000) / 1000.f;
    hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]);
  }
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, (size_t)M * N * 4));
  a.A = dA; a.B = dB; a.D = dD;
  const size_t smem = (size_t)(M + NB) * K * 2;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 256, smem>>>(a);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hD((size_t)M * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)fA[(size_t)m * K + k] * fB[(size_t)(a.n0 + n) * K + k];
      r *= a.passes;
      const double e = fabs(r - hD[(size_t)m * N + n]);
      if (!(e <= 1e-3 * (1 + fabs(r)))) ++bad;
      if (e > maxerr || e != e) maxerr = e;
      if (fabs(r) > maxref) maxref = fabs(r);
    }
  printf("{\"umma_probe_test\": %d, \"mode\": %d, \"M\": %d, \"N\": %d, \"K\": %d, \"passes\": %d, \"bad\": %d, \"max_err\": %.3e, \"max_ref\": %.3e, \"ok\": %s}\n",
         test, mode, M, N, K, a.passes, bad, maxerr, maxref, bad == 0 ? "true" : "false");
  return bad == 0 ? 0 : 1;
}

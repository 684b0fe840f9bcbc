// Pipe-rate probe for the scan's inner-loop design (B200, sm_100a): how many warp-instructions per clock and SM sub-partition
// each candidate instruction sustains, alone and mixed — the numbers DESIGN.md §4.1 budgets with.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/pipe_probe.cu -o scripts/_bin/pipe_probe
// Every kernel: 148*k CTAs x 256 threads (8 warps per CTA, 2 per sub-partition), 8-16 independent chains per thread, clock64
// around the loop of ONE warp per CTA; reports values per clock per SM (from cycles) and per second (from CUDA events).
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float tanhf_(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t tanhh2(uint32_t x) { uint32_t y; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t pack(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }
__device__ __forceinline__ uint32_t hfma2_(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

enum { K_FFMA, K_FFMA2, K_FMUL2, K_EX2, K_EX2H2, K_EX2B2, K_TANH, K_TANHH2, K_RCP, K_LG2, K_HFMA2,
       K_MIX_EX2_FFMA2x2,      // the scan's MUFU-only inner loop per state pair: 2 ex2 + FMUL2 + FMUL2 + FFMA2 + FFMA2
       K_MIX_EX2_FFMA2x4,
       K_MIX_EX2H2_FFMA2x2,    // one f16x2 ex2 + 2 cvt + the same four packed ops
       K_POLY2,                // exp2 of a pair on the FMA pipe: 2 FADD2-ish + 6 FFMA2 + 2 shift-adds
       K_MIX_HALF_POLY,        // per two pairs: one pair through MUFU, one through the polynomial + 8 packed ops
       K_LDS128_BCAST, K_LDS128, K_CVT_H2F, K_COUNT };
static const char* kNames[] = {"FFMA (3-reg)", "FFMA2 (fma.rn.f32x2)", "FMUL2", "MUFU ex2.f32", "MUFU ex2.f16x2 (pairs)", "MUFU ex2.bf16x2 (pairs)",
  "MUFU tanh.f32", "MUFU tanh.f16x2 (pairs)", "MUFU rcp.f32", "MUFU lg2.f32", "HFMA2", "mix: 2 ex2 + 4 packed f32x2 (per state pair)",
  "mix: 2 ex2 + 8 packed f32x2", "mix: 1 ex2.f16x2 + 2 cvt + 4 packed", "poly exp2 pair on FMA pipe", "mix: pair MUFU + pair poly + 8 packed",
  "LDS.128 broadcast", "LDS.128 per-lane", "cvt f16x2->f32 x2"};
// "values" counted per inner iteration and chain (what the rate is quoted in)
static const double kValues[] = {1, 2, 2, 1, 2, 2, 1, 2, 1, 1, 2, 2, 2, 2, 2, 4, 1, 1, 2};
// warp-instructions issued per inner iteration and chain (approximate, for the issue-slot reading)
static const double kInstr[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 6, 10, 7, 10, 20, 1, 1, 2};

template <int WHICH>
__global__ void __launch_bounds__(256) probe(float* out, long long* cyc, int iters, float seed) {
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(seed, seed * 0.5f, seed * 0.25f, 1.f);
  __syncthreads();
  constexpr int NC = 8;
  float v[NC], w[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) { v[k] = seed - 1e-3f * (threadIdx.x + k); w[k] = seed * 0.5f + 1e-3f * k; }
  const float c0 = 0.999f, c1 = 1e-6f;
  const uint64_t C0 = pack(c0, c0), C1 = pack(c1, c1);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (WHICH == K_FFMA) v[k] = fmaf(v[k], w[k], c1);
      else if (WHICH == K_FFMA2) { uint64_t p = fma2(pack(v[k], w[k]), C0, C1); v[k] = __uint_as_float((uint32_t)p); w[k] = __uint_as_float((uint32_t)(p >> 32)); }
      else if (WHICH == K_FMUL2) { uint64_t p = mul2(pack(v[k], w[k]), C0); v[k] = __uint_as_float((uint32_t)p); w[k] = __uint_as_float((uint32_t)(p >> 32)); }
      else if (WHICH == K_EX2) v[k] = ex2f(v[k]);
      else if (WHICH == K_EX2H2) v[k] = __uint_as_float(ex2h2(__float_as_uint(v[k])));
      else if (WHICH == K_EX2B2) v[k] = __uint_as_float(ex2b2(__float_as_uint(v[k])));
      else if (WHICH == K_TANH) v[k] = tanhf_(v[k]);
      else if (WHICH == K_TANHH2) v[k] = __uint_as_float(tanhh2(__float_as_uint(v[k])));
      else if (WHICH == K_RCP) v[k] = rcpf(v[k]);
      else if (WHICH == K_LG2) v[k] = lg2f(v[k]);
      else if (WHICH == K_HFMA2) v[k] = __uint_as_float(hfma2_(__float_as_uint(v[k]), __float_as_uint(w[k]), __float_as_uint(w[k])));
      else if (WHICH == K_MIX_EX2_FFMA2x2 || WHICH == K_MIX_EX2_FFMA2x4) {
        // state pair (v = h.x, w = h.y): x = dt*A2 (FMUL2), a = ex2 x2, b = du*B (FMUL2), h = a h + b (FFMA2), y += C h (FFMA2)
        uint64_t x = mul2(C0, pack(-0.01f * (k + 1), -0.02f * (k + 1)));
        float ax = ex2f(__uint_as_float((uint32_t)x)), ay = ex2f(__uint_as_float((uint32_t)(x >> 32)));
        uint64_t b = mul2(C1, C0);
        uint64_t h = fma2(pack(ax, ay), pack(v[k], w[k]), b);
        uint64_t y = fma2(C0, h, C1);
        if (WHICH == K_MIX_EX2_FFMA2x4) { y = fma2(y, C0, h); y = fma2(y, C0, C1); y = fma2(y, C0, h); h = fma2(y, C1, h); }
        v[k] = __uint_as_float((uint32_t)h) + 1e-9f * __uint_as_float((uint32_t)y); w[k] = __uint_as_float((uint32_t)(h >> 32));
      } else if (WHICH == K_MIX_EX2H2_FFMA2x2) {
        uint64_t x = mul2(C0, pack(-0.01f * (k + 1), -0.02f * (k + 1)));
        __half2 xh = __floats2half2_rn(__uint_as_float((uint32_t)x), __uint_as_float((uint32_t)(x >> 32)));
        uint32_t ah = ex2h2(*reinterpret_cast<uint32_t*>(&xh));
        float2 af = __half22float2(*reinterpret_cast<__half2*>(&ah));
        uint64_t b = mul2(C1, C0);
        uint64_t h = fma2(pack(af.x, af.y), pack(v[k], w[k]), b);
        uint64_t y = fma2(C0, h, C1);
        v[k] = __uint_as_float((uint32_t)h) + 1e-9f * __uint_as_float((uint32_t)y); w[k] = __uint_as_float((uint32_t)(h >> 32));
      } else if (WHICH == K_POLY2 || WHICH == K_MIX_HALF_POLY) {
        const uint64_t MAG = pack(12582912.f, 12582912.f), NMAG = pack(-12582912.f, -12582912.f), M1 = pack(-1.f, -1.f), ONE = pack(1.f, 1.f);
        uint64_t x = pack(v[k], w[k]);
        uint64_t t = fma2(x, ONE, MAG);
        uint64_t f = fma2(fma2(t, ONE, NMAG), M1, x);
        uint64_t p = fma2(pack(0.0013264726f, 0.0013264726f), f, pack(0.009671513f, 0.009671513f));
        p = fma2(p, f, pack(0.055507336f, 0.055507336f));
        p = fma2(p, f, pack(0.24022242f, 0.24022242f));
        p = fma2(p, f, pack(0.693147f, 0.693147f));
        p = fma2(p, f, ONE);
        float ax = __uint_as_float((uint32_t)p + ((uint32_t)t << 23)), ay = __uint_as_float((uint32_t)(p >> 32) + ((uint32_t)(t >> 32) << 23));
        if (WHICH == K_POLY2) { v[k] = ax - 1.5f; w[k] = ay - 1.5f; }
        else {
          // + a MUFU pair and the 8 packed ops of two state pairs
          float bx = ex2f(v[k] * 0.5f), by = ex2f(w[k] * 0.5f);
          uint64_t h0 = fma2(pack(ax, ay), x, C1), h1 = fma2(pack(bx, by), x, C1);
          uint64_t y = fma2(C0, h0, C1); y = fma2(C0, h1, y);
          uint64_t m0 = mul2(C0, h0), m1 = mul2(C1, h1), m2 = mul2(m0, C0), m3 = mul2(m1, C1);
          y = fma2(m2, m3, y);
          v[k] = __uint_as_float((uint32_t)y) * 1e-3f - 0.5f; w[k] = __uint_as_float((uint32_t)(y >> 32)) * 1e-3f - 0.7f;
        }
      } else if (WHICH == K_LDS128_BCAST) { float4 q = sm[(k * 7 + it) & 255]; v[k] += q.x + q.w; }
      else if (WHICH == K_LDS128) { float4 q = sm[(threadIdx.x + k * 7 + it) & 255]; v[k] += q.x + q.w; }
      else if (WHICH == K_CVT_H2F) { uint32_t u = __float_as_uint(v[k]); float2 f = __half22float2(*reinterpret_cast<__half2*>(&u)); v[k] = f.x + 1.0f; w[k] += f.y; }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NC; ++k) s += v[k] + w[k];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int WHICH>
int run(float* d_out, long long* d_cyc, int sms, int ctas_per_sm) {
  const int iters = 2048, blocks = sms * ctas_per_sm, threads = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    probe<WHICH><<<blocks, threads>>>(d_out, d_cyc, iters, -0.75f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  static long long h_cyc[148 * 16];
  CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < blocks; ++i) mean += h_cyc[i]; mean /= blocks;
  const double chains = 8.0 * iters * threads * ctas_per_sm;        // per SM
  const double values_per_clk_sm = chains * kValues[WHICH] / mean;
  const double warp_instr_per_clk_smsp = chains / 32.0 * kInstr[WHICH] / mean / 4.0;
  printf("%-48s ctas/SM %d  %8.2f values/clk/SM  %6.3f warp-instr/clk/SMSP  %9.3e values/s (events)\n", kNames[WHICH], ctas_per_sm,
         values_per_clk_sm, warp_instr_per_clk_smsp, 8.0 * iters * (double)blocks * threads * kValues[WHICH] / (best * 1e-3));
  fflush(stdout);
  return 0;
}

template <int W> int run_all(float* o, long long* c, int sms) {
  if constexpr (W < K_COUNT) {
    if (run<W>(o, c, sms, 2)) return 1;
    if (run<W>(o, c, sms, 4)) return 1;
    return run_all<W + 1>(o, c, sms);
  }
  return 0;
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 4)); CK(cudaMalloc(&d_cyc, sizeof(long long) * 148 * 16));
  printf("pipe_probe: %d SMs\n", sms);
  return run_all<0>(d_out, d_cyc, sms);
}

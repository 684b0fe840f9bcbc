/* ORACLE (test infrastructure only) — plain-C restatement of one Mamba inner pass
 *   causal depthwise conv + SiLU -> x_proj -> dt_proj + softplus -> selective scan -> + D*u -> * silu(z)
 * following SURVEY.md Appendix A.1/A.2 (upstream mamba-ssm 1.2.0.post1 `mamba_inner_ref` / `selective_scan_ref`,
 * un-vendored; reached from ref:caduceus/modeling_caduceus.py:128-133).  fp32 arithmetic like selective_scan_ref,
 * libm expf/log1pf, sequential recurrence.  OpenMP over channels.  Used for large-L checks and as a CPU baseline;
 * checked against the torch restatement in tests/test_oracle.py.  Never linked into the product.
 *
 *   xz   (2E, L) row-major: rows [0,E) x, [E,2E) z          conv_w (E, K), conv_b (E) or NULL
 *   w_x  (R+2N, E)   w_dt (E, R)   dt_b (E)   A (E, N) (already -exp(A_log))   Dskip (E)
 *   rev != 0: run in reversed time (the reference's flip -> Mamba -> flip, ref:caduceus/modeling_caduceus.py:130-133)
 *   y    (E, L) gated output (before out_proj);  xdbl_out optional (R+2N, L)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float siluf(float v) { return v / (1.0f + expf(-v)); }
static inline float softplusf(float v) { return v > 20.0f ? v : log1pf(expf(v)); }

int cad_oracle_mamba_inner(const float* xz, const float* conv_w, const float* conv_b, const float* w_x,
                           const float* w_dt, const float* dt_b, const float* A, const float* Dskip,
                           long L, long E, long N, long R, long K, int rev, float* y, float* xdbl_out) {
  const long rows = R + 2 * N;
  float* u = (float*)malloc(sizeof(float) * E * L);
  float* xdbl = xdbl_out ? xdbl_out : (float*)malloc(sizeof(float) * rows * L);
  if (!u || !xdbl) return -1;
  /* logical time tau <-> physical t */
#define PHYS(tau) (rev ? (L - 1 - (tau)) : (tau))
#pragma omp parallel for schedule(static)
  for (long d = 0; d < E; ++d) {
    const float* x = xz + d * L;
    for (long tau = 0; tau < L; ++tau) {
      float acc = conv_b ? conv_b[d] : 0.0f;
      for (long k = 0; k < K; ++k) {
        long src = tau - (K - 1) + k;
        if (src >= 0) acc += conv_w[d * K + k] * x[PHYS(src)];
      }
      u[d * L + PHYS(tau)] = siluf(acc);
    }
  }
#pragma omp parallel for schedule(static)
  for (long t = 0; t < L; ++t) {
    for (long r = 0; r < rows; ++r) {
      float acc = 0.0f;
      for (long d = 0; d < E; ++d) acc += w_x[r * E + d] * u[d * L + t];
      xdbl[r * L + t] = acc;
    }
  }
#pragma omp parallel for schedule(static)
  for (long d = 0; d < E; ++d) {
    float h[64];
    for (long n = 0; n < N; ++n) h[n] = 0.0f;
    const float* z = xz + (E + d) * L;
    for (long tau = 0; tau < L; ++tau) {
      const long t = PHYS(tau);
      float dtr = dt_b[d];
      for (long r = 0; r < R; ++r) dtr += w_dt[d * R + r] * xdbl[r * L + t];
      const float dt = softplusf(dtr);
      const float uu = u[d * L + t];
      float acc = 0.0f;
      for (long n = 0; n < N; ++n) {
        h[n] = expf(dt * A[d * N + n]) * h[n] + dt * xdbl[(R + n) * L + t] * uu;
        acc += h[n] * xdbl[(R + N + n) * L + t];
      }
      acc += Dskip[d] * uu;
      y[d * L + t] = acc * siluf(z[t]);
    }
  }
#undef PHYS
  free(u);
  if (!xdbl_out) free(xdbl);
  return 0;
}

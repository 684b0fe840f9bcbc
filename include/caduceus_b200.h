/*
 * caduceus_b200 — C-ABI of the B200-native (sm_100a) kernels behind Caduceus' bidirectional selective-SSM
 * hot path.  Plain C: device pointers, sizes, strides, enums.  No torch / ATen types.
 *
 * What this boundary replaces (SURVEY.md §8b "B-native").  The reference (kuleshov-group/caduceus) holds no
 * native code of its own; its hot path calls the pybind entry points of two un-vendored CUDA extensions,
 * reached through these reference call sites:
 *
 *   selective_scan_cuda.fwd / .bwd      <- mamba_ssm.Mamba.forward  <- ref:caduceus/modeling_caduceus.py:105-113,128-133
 *   causal_conv1d_cuda.causal_conv1d_fwd/_bwd  (same call chain)
 *   Triton _layer_norm_fwd/_bwd         <- rms_norm_fn/layer_norm_fn <- ref:caduceus/modeling_caduceus.py:244-273,
 *                                                                       ref:caduceus/modeling_rcps.py:177-195
 *   torch index ops (flip/gather/cat)   <- RCPSEmbedding.forward     <- ref:caduceus/modeling_rcps.py:46-67
 *
 * Conventions
 *   - Every function returns 0 on success, <0 for an argument/shape error (text via cad_last_error()),
 *     >0 = the cudaError_t of a failed launch.  No C++ exceptions cross the boundary.
 *   - The library never allocates or frees device memory and never synchronises the device.  All work is
 *     enqueued on the `stream` argument (a cudaStream_t passed as void*).
 *   - The caller (PyTorch) owns every buffer and keeps it alive until the stream reaches the call.
 *   - All "time-major rows" are contiguous along the sequence axis; `ld*` arguments are row pitches in ELEMENTS.
 *     Fast (128-bit) paths need 16-byte aligned base pointers and pitches; otherwise a scalar path is used.
 */
#ifndef CADUCEUS_B200_H_
#define CADUCEUS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAD_ABI_VERSION 5

typedef enum { CAD_F32 = 0, CAD_F16 = 1, CAD_BF16 = 2 } cad_dtype;

/* ---- library ---------------------------------------------------------------------------------------- */
int         cad_version(void);            /* == CAD_ABI_VERSION */
const char* cad_last_error(void);         /* thread-local message of the last <0 return */
int         cad_sm_count(void);           /* multiprocessors of the current device (cached) */

/* ---- token embedding, Ph and RC-equivariant PS (ref:caduceus/modeling_caduceus.py:159-163,
 *      ref:caduceus/modeling_rcps.py:46-67).  Integer index work, bit-exact:
 *        out[b, l, c]     = W[ids[b, l], c]                     c in [0, D)
 *        out[b, l, D + c] = W[cmap[ids[b, l]], D - 1 - c]       (only if rcps)
 *      which equals cat[emb(ids), flip_{L,C}(emb(cmap[flip_L(ids)]))].                                   */
typedef struct {
  const int64_t* ids;      /* (B, L) token ids */
  const void*    weight;   /* (V, D) embedding table, dtype `dtype` */
  const int64_t* cmap;     /* (V) complement map, or NULL when !rcps */
  void*          out;      /* (B, L, D) or (B, L, 2D) */
  int64_t B, L, V, D;
  int32_t rcps;
  int32_t dtype;           /* cad_dtype of weight/out */
} cad_embedding_args;
int cad_embedding_fwd(const cad_embedding_args* a, void* stream);

/* gradient of the table: dW[v, c] += sum over (b,l) with ids==v of dout[b,l,c]  (+ RC half, same index map).
 * dW must be zero-initialised fp32 (V, D). */
typedef struct {
  const int64_t* ids; const int64_t* cmap; const void* dout; float* dweight;
  int64_t B, L, V, D; int32_t rcps; int32_t dtype;
} cad_embedding_bwd_args;
int cad_embedding_bwd(const cad_embedding_bwd_args* a, void* stream);

/* ---- fused residual-add + RMSNorm / LayerNorm (replaces rms_norm_fn / layer_norm_fn).
 *      Rows are (token, half) pairs; `nhalf` = 1 (Ph) or 2 (PS, hidden has 2*D channels).
 *      For half h of token row r:
 *         v   = x[r, in_half(h)*D : +D] (+ residual[r, in_half(h)*D : +D])          in fp32
 *         res_out[r, h*D : +D] = v                                                 (if res_out != NULL)
 *         y[r, h*D : +D]       = norm(v) * w' (+ b'),   w' = weight or reversed(weight) per `wflip_mask`
 *      in_half(h) = h ^ swap.  The literal half swap of the reference's fused RCPS block
 *      (ref:caduceus/modeling_rcps.py:177-197) is swap=1, wflip_mask=0b10; its non-fused
 *      block and the final norm (ref:caduceus/modeling_caduceus.py:242-262) are swap=0, wflip_mask=0b10.    */
typedef struct {
  const void* x;          /* (rows, nhalf*D), pitch ldx */
  const void* residual;   /* same shape, pitch ldr, dtype res_in_dtype; may be NULL */
  const void* weight;     /* (D) dtype wdtype */
  const void* bias;       /* (D) or NULL */
  void*       y;          /* (rows, nhalf*D) pitch ldy, dtype xdtype */
  void*       res_out;    /* (rows, nhalf*D) pitch ldo, dtype res_out_dtype; may be NULL */
  float*      rstd;       /* (rows, nhalf) saved 1/sigma for backward, or NULL */
  float*      mean;       /* (rows, nhalf) saved mean (LayerNorm only), or NULL */
  int64_t rows, D, ldx, ldr, ldy, ldo;
  int32_t nhalf, swap, wflip_mask, is_rms;
  int32_t xdtype, wdtype, res_in_dtype, res_out_dtype;
  float   eps;
} cad_add_norm_args;
int cad_add_norm_fwd(const cad_add_norm_args* a, void* stream);

/* backward of the above.  dy: grad of y; dres_out: grad flowing into res_out (may be NULL).
 * Outputs: dx (grad of x == grad of residual, written once; caller aliases), dweight/dbias partial sums
 * (nblocks, D) fp32 to be summed by the caller (deterministic two-stage reduction). */
typedef struct {
  const void* dy; const void* dres_out; const void* v /* saved res_out (or x when no residual) */;
  const void* weight; const float* rstd; const float* mean;
  void* dx; float* dweight_partial; float* dbias_partial;
  int64_t rows, D, lddy, lddr, ldv, lddx;
  int32_t nhalf, swap, wflip_mask, is_rms, has_bias;
  int32_t dydtype, vdtype, wdtype, dxdtype, drdtype;
  int32_t nblocks;
} cad_add_norm_bwd_args;
int cad_add_norm_bwd(const cad_add_norm_bwd_args* a, void* stream);
int cad_add_norm_bwd_blocks(int64_t rows);   /* nblocks the backward wants for `rows` */

/* ---- BiMamba inner path: (anti)causal depthwise conv + SiLU, softplus(dt), selective scan, gate.
 *      Replaces causal_conv1d_fwd + selective_scan_cuda.fwd (and the two flips + second Mamba of
 *      ref:caduceus/modeling_caduceus.py:128-137; the RC strand of ref:caduceus/modeling_rcps.py:85-99).
 *
 *  A "job" j = (sequence s, direction p): one Mamba run over one sequence in one physical direction.
 *      xz      (nseq, 2E, ldxz)     in-proj output, channel-major; rows [0,E) = x, [E,2E) = z
 *      delta   (njobs, E, ldd)      dt_raw = W_dt . x_dbl[0:R]  (the dt_proj GEMM; bias NOT added), io dtype
 *      bc      (njobs, 2N, ldbc)    fp32 rows [0,N) = B, [N,2N) = C of job j (x_dbl[R:]); columns [L, ldbc)
 *                                   must hold finite values (zeros).  Read through ONE TMA tensor map.
 *      out     (njobs, E, ldo)      gated scan output  y * silu(z)   (pre out_proj)
 *  job j uses sequence  seq_of_job[j], parameter set  pset_of_job[j]  (0 = mamba_fwd, 1 = mamba_rev) and runs
 *  right-to-left in physical time iff  rev_of_job[j].  Parameter sets are packed fp32:
 *      conv_w (P, E, 4) taps zero-padded at the front to 4;  conv_b (P, E);  dt_b (P, E);
 *      A2 (P, E, N) = -exp(A_log) * log2(e);  Dskip (P, E).
 *  Semantics in LOGICAL time tau (tau = t, or L-1-t when reversed), SURVEY.md A.1/A.2/A.5:
 *      u[tau]  = silu(conv_b + sum_k conv_w[k] * x[tau-3+k])
 *      dt[tau] = softplus(delta[tau] + dt_b)                  (threshold 20)
 *      h[tau]  = exp2(dt*A2) * h[tau-1] + dt*B[tau]*u[tau],  h[-1] = h0 (or 0)
 *      y[tau]  = C[tau] . h[tau] + Dskip*u[tau];    out = y * silu(z)
 *  Optional sequence-sharding hooks: conv halo in, carry-state in, and per-job outputs
 *  (sum of dt per channel, final state) so that a shard can be composed with its neighbours.               */
typedef struct {
  const void*  xz;
  const void*  delta;
  const float* bc;
  void*        out;
  const float* conv_w; const float* conv_b; const float* dt_b; const float* A2; const float* Dskip;
  const int32_t* seq_of_job; const int32_t* pset_of_job; const int32_t* rev_of_job;   /* device, (njobs) */
  /* sharding hooks, all optional (NULL): */
  const void*  halo;        /* (njobs, E, 3) x values preceding logical time 0, dtype io */
  const float* h0;          /* (njobs, E, N) carry-in state */
  float*       hlast;       /* (njobs, E, N) final state (with h0 folded in) */
  float*       dtsum;       /* (njobs, E) sum over tau of dt */
  /* optional saved tensors for backward (NULL in inference): */
  float*       chunk_state; /* (njobs, E, nchunks, N) state at the END of each 512-token logical chunk */
  int64_t L, E, N, K;
  int64_t ldxz, ldd, ldbc, ldo;
  int32_t nseq, njobs, npset;
  int32_t io_dtype;         /* dtype of xz / delta / out / halo */
  int32_t channels_per_cta; /* 0 = library picks so that the grid is ~ a multiple of the SM count */
  int32_t state_only;       /* 1: only hlast / dtsum are produced (pass 1 of a sequence-sharded scan); out may be NULL */
  int32_t variant;          /* 0 = library default (3);
                               3  = one channel per WARP, lane = 16 consecutive tokens of a 512-token chunk: time-parallel inside the
                                    warp (zero-state pass, 5-step shuffle scan of the segment aggregates, replay).  Every dtype, every
                                    hook, saved chunk states: the general kernel (training, fp32, short sequences, sequence shards).
                               20 = one channel per LANE: a warp owns 32 channels and walks time serially, 16 states per lane as packed
                                    fp32 pairs, B / C as broadcast reads, no shuffles — MUFU-bound (87 % of the pipe).  16-bit I/O,
                                    inference: no chunk_state / state_only; h0 / hlast / dtsum go through cad_seg_carry instead of this
                                    launch; halo is honoured; channels_per_cta then counts WARPS per CTA (<= 8).  The sequence is cut
                                    into nseg segments per job, each scanned from a ZERO state (see below). */
  /* variant 20 only: */
  const float* bcT;         /* (njobs, ceil256(L), 2N) fp32: the B / C rows TOKEN-major, zeros beyond L (cad_conv_xproj_fwd writes it,
                               or cad_bc_transpose) */
  int32_t nseg;             /* time segments per job (grid.z), whole 256-token chunks each; 0 = 1.  Every segment is scanned
                               from a ZERO state: with nseg > 1 `out` still lacks the carries (cad_seg_carry + cad_bimamba_scan_fixup) */
  float* seg_state;         /* (njobs, nseg, E, N) end state of every LOGICAL segment scanned from zero; required when nseg > 1 */
  float* seg_dtsum;         /* (njobs, nseg, E) sum of dt over the segment */
  /* variant 3 only, optional (NULL): */
  float* chunk_dtsum;       /* (njobs, E, nchunks) sum of dt over every 512-token LOGICAL chunk (nchunks = ceil(L / 512); physical
                               chunk pc of a reversed job is logical chunk nchunks - 1 - pc): lets cad_shard_seg_carry split the
                               carry fix-up of a sequence shard into segments that run in parallel */
} cad_scan_fwd_args;
int cad_bimamba_scan_fwd(const cad_scan_fwd_args* a, void* stream);
int cad_scan_chunk_len(void);      /* logical tokens per saved chunk state (512) */

/* helpers of scan variant 20 (a sequence cut into nseg segments INSIDE one GPU; same algebra as the multi-GPU path):
 *   cad_bc_transpose: bc (njobs, 2N, ldbc) fp32 -> bcT (njobs, ceil256(L), 2N), zeros in rows [L, ceil256(L)).
 *   cad_seg_carry:    carry[j, s] = state entering logical segment s of job j:
 *                       carry[j, 0] = h0[j] (or 0),  carry[j, s+1] = exp2(A2 * seg_dtsum[j, s]) * carry[j, s] + seg_state[j, s]
 *                     and, optionally, what a sequence shard hands to its neighbours: hlast[j] = the state after the last
 *                     segment, dtsum[j] = sum over the segments of seg_dtsum.  h0 / carry / hlast / dtsum may be NULL.  (N == 16)
 *   the carries are then applied by cad_bimamba_scan_fixup with nseg / seg_carry set (seg_first = 1 when h0 was given).
 *   cad_shard_seg_carry: the carry-in h0 of a sequence SHARD scanned by variant 3 from a zero state (which carries its own state
 *                     from chunk to chunk), decayed to the start of every segment of `per512` physical 512-token chunks:
 *                       carry[j, s] = exp2(A2 * sum of dt over the logical segments before s) * h0[j],  entries whose exponent has
 *                       fallen below cutoff_log2 set to exactly 0 (the fix-up skips them) — so that cad_bimamba_scan_fixup with
 *                       nseg / seg_carry / seg_first = 1 applies the shard's carry to all segments IN PARALLEL instead of walking
 *                       each channel serially.  nseg = ceil(nchunks / per512); requires ceil(ceil(L/256) / nseg) == 2 * per512.    */
int cad_bc_transpose(const float* bc, float* bcT, int64_t njobs, int64_t N2, int64_t L, int64_t ldbc, void* stream);
int cad_shard_seg_carry(const float* chunk_dtsum, const float* A2, const int32_t* pset_of_job, const int32_t* rev_of_job,
                        const float* h0, float* carry, int64_t njobs, int64_t E, int64_t nchunks, int64_t nseg, int64_t per512,
                        float cutoff_log2, void* stream);
int cad_seg_carry(const float* seg_state, const float* seg_dtsum, const float* A2, const int32_t* pset_of_job, const float* h0,
                  float* carry, float* hlast, float* dtsum, int64_t njobs, int64_t nseg, int64_t E, void* stream);

/* ---- carry fix-up of a sequence-sharded scan: adds, in place, the contribution of the carry-in
 *      state h0 to an output that was produced by cad_bimamba_scan_fwd with a ZERO carry:
 *          out[tau] += silu(z[tau]) * sum_n C[tau,n] * exp2(A2[n] * sum_{s<=tau} dt[s]) * h0[n]
 *      (channel, state) pairs are dropped once A2*cumdt < cutoff_log2; CTAs stop when nothing is left to add.   */
typedef struct {
  const void* xz; const void* delta; const float* bc; void* out;
  const float* dt_b; const float* A2;
  const int32_t* seq_of_job; const int32_t* pset_of_job; const int32_t* rev_of_job;
  const float* h0;          /* (njobs, E, N) carry-in state */
  int64_t L, E, N;
  int64_t ldxz, ldd, ldbc, ldo;
  int32_t nseq, njobs, io_dtype, channels_per_cta;
  float   cutoff_log2;      /* e.g. -40: terms below 2^-40 * |C h0| are dropped */
  /* in-GPU segments (scan variant 20), optional: with nseg > 1 the kernel runs once per (job, logical segment s >= 1) on the
   * segment's own token range (the split of cad_scan_fwd_args.nseg) with the carry  seg_carry[job, s]; h0 is ignored.      */
  int32_t nseg;
  const float* seg_carry;   /* (njobs, nseg, E, N) from cad_seg_carry */
  int32_t seg_first;        /* 1: logical segment 0 is fixed up as well (its carry is the shard's own carry-in h0) */
} cad_scan_fixup_args;
int cad_bimamba_scan_fixup(const cad_scan_fixup_args* a, void* stream);

/* ---- sequence sharding over the GPUs of one NVLink / NVSwitch box: the two per-layer exchanges of the sharded BiMamba call done by
 *      kernels that store straight into the PEERS' memory (no library collective on the data path, no host round trip, capturable
 *      in a CUDA graph).  This is the north_star's "exchange of the d_state boundary hidden state per layer over NVLink", in the
 *      non-serial form of SURVEY.md §8e: every rank scans from zero, all boundary states are exchanged ONCE, each rank composes
 *      its own carry.  The reference has no counterpart (it scales by DDP only, ref:train.py:629-639).
 *
 *      Every rank allocates a workspace of cad_peer_ws_bytes(...) bytes, zero-filled once, that all peers can address (CUDA IPC /
 *      torch symmetric memory); peer_ws[r] is rank r's workspace as mapped into THIS process.  All ranks must issue the same
 *      sequence of cad_peer_* calls.  A peer that never arrives makes the waiting kernel trap after 10 s (the launch fails).      */
typedef struct {
  const void* peer_ws;      /* DEVICE array of `world` uint64 base addresses */
  int32_t rank, world;      /* world <= 16 */
  int64_t nseq_max, njobs_max, E, N;   /* workspace geometry: the same values on every rank */
} cad_peer_ctx;
int64_t cad_peer_ws_bytes(int32_t world, int64_t nseq_max, int64_t njobs_max, int64_t E, int64_t N);
/* conv halo: pushes the first / last three x samples of every sequence of this shard (xz rows [0, E), shard length L >= 3) to the
 * left / right neighbour, waits for theirs and writes halo (njobs, E, 3) in the io dtype: the three samples that logically precede
 * the shard for every job (zeros at the ends of the full sequence) — the `halo` argument of the conv / scan entry points.        */
int cad_peer_halo_exchange(const cad_peer_ctx* ctx, const void* xz, int64_t ldxz, int64_t L, int32_t nseq, int32_t njobs,
                           const int32_t* seq_of_job, const int32_t* rev_of_job, void* halo, int32_t io_dtype, void* stream);
/* boundary state: pushes (hlast (njobs, E, N), dtsum (njobs, E)) of this shard's zero-carry scan to every rank, waits for all of
 * them and composes this rank's carry-in  h0 (njobs, E, N):  h <- exp2(A2 * dtsum_r) * h + hlast_r  over the logical predecessors r.
 * dtsum_all (world, njobs, E) optional (NULL): every rank's sum dt (the sharded backward composes adjoint carries with it).      */
int cad_peer_carry_exchange(const cad_peer_ctx* ctx, const float* hlast, const float* dtsum, const float* A2,
                            const int32_t* pset_of_job, const int32_t* rev_of_job, int32_t njobs, float* h0, float* dtsum_all,
                            void* stream);

/* ---- adjoint of the carry (sequence-sharded TRAINING, SURVEY.md §8e "Backward"): the gradient w.r.t. a shard's
 *      carry-in state produced by the shard's own tokens (adjoint carry-in taken as zero),
 *          dh[n] = sum_tau exp2(A2[n] * sum_{s<=tau} dt[s]) * C[tau,n] * dout[tau] * silu(z[tau])
 *      — the transpose of cad_bimamba_scan_fixup, with the same decay cut-off.  One all_gather of dh (and the
 *      forward's sum dt) lets every rank compose its true adjoint carry-in and run cad_bimamba_scan_bwd once.    */
typedef struct {
  const void* xz; const void* delta; const float* bc; const void* dout;
  const float* dt_b; const float* A2;
  const int32_t* seq_of_job; const int32_t* pset_of_job; const int32_t* rev_of_job;
  float* dh;                /* (njobs, E, N), written for every entry */
  int64_t L, E, N;
  int64_t ldxz, ldd, ldbc, ldo;
  int32_t nseq, njobs, io_dtype, channels_per_cta;
  float   cutoff_log2;
} cad_scan_adjoint_args;
int cad_bimamba_scan_adjoint(const cad_scan_adjoint_args* a, void* stream);

/* ---- backward of cad_bimamba_scan_fwd (replaces selective_scan_cuda.bwd).
 *      Inputs as the forward plus dout (njobs, E, ldo) and the forward's chunk_state.  Outputs:
 *        dz      (njobs, E, lddz)   gradient of the gate input z                               io dtype
 *        du      (njobs, E, lddu)   gradient w.r.t. u = silu(conv(x)) from the scan path       io dtype
 *        ddelta  (njobs, E, lddd)   gradient w.r.t. dt_raw                                     io dtype
 *        dbc     (njobs, 2N, ldbc)  gradient w.r.t. B/C rows, fp32, ACCUMULATED (caller zero-fills)
 *        ddt_b, dDskip (P, E), dA2 (P, E, N)   fp32, ACCUMULATED (caller zero-fills)
 *        dh0     (njobs, E, N)      gradient w.r.t. the carry-in state, or NULL
 *      dhlast (njobs, E, N) or NULL: gradient w.r.t. the END state (adjoint carry-in from the logically next shard). */
typedef struct {
  const void*  xz; const void* delta; const float* bc; const void* dout;
  const float* conv_w; const float* conv_b; const float* dt_b; const float* A2; const float* Dskip;
  const int32_t* seq_of_job; const int32_t* pset_of_job; const int32_t* rev_of_job;
  const void*  halo; const float* h0; const float* chunk_state;
  void* dz; void* du; void* ddelta; float* dbc;
  float* ddt_b; float* dA2; float* dDskip; float* dh0; const float* dhlast;
  int64_t L, E, N, K;
  int64_t ldxz, ldd, ldbc, ldo, lddz, lddu, lddd;
  int32_t nseq, njobs, npset, io_dtype, channels_per_cta;
} cad_scan_bwd_args;
int cad_bimamba_scan_bwd(const cad_scan_bwd_args* a, void* stream);

/* backward of the depthwise (anti)causal conv + SiLU: given du (total gradient of u), recomputes the conv
 * pre-activation from x and writes dx (njobs, E, lddx); dconv_w (P, E, 4) / dconv_b (P, E) are ACCUMULATED.
 * Replaces causal_conv1d_cuda.causal_conv1d_bwd.                                                              */
typedef struct {
  const void* xz; const void* du; void* dx;
  const float* conv_w; const float* conv_b; float* dconv_w; float* dconv_b;
  const int32_t* seq_of_job; const int32_t* pset_of_job; const int32_t* rev_of_job;
  const void* halo;
  int64_t L, E, ldxz, lddu, lddx;
  int32_t nseq, njobs, io_dtype;
} cad_conv_bwd_args;
int cad_conv_silu_bwd(const cad_conv_bwd_args* a, void* stream);

/* ---- fused conv + SiLU -> x_proj -> dt_proj on the 5th-generation tensor cores (16-bit I/O): produces exactly the operands
 *      of cad_bimamba_scan_fwd without materialising u = silu(conv(x)).  Replaces causal_conv1d_fwd + the x_proj and dt_proj
 *      GEMMs of upstream's mamba_inner_fn.  csrc/xproj.cu: tcgen05.mma issued by one thread per CTA,
 *      accumulators in tensor memory read back with tcgen05.ld, x / W_x slabs by TMA loads, delta by TMA stores; persistent
 *      warp-specialised CTAs, two per SM.
 *        w_x (P, R+2N, E), w_dt (P, E, R) in the io dtype;  delta (njobs, E, ldd) io dtype, written for every column < ldd;
 *        bc (njobs, 2N, ldbc) fp32, written for every column < min(ldbc, ceil128(L)) (zeros beyond L).
 *      Constraints: io dtype f16/bf16, N == 16, R <= 16, E % 64 == 0, E <= 2048, ldxz % 8 == 0, ldd % 8 == 0; otherwise use
 *      cad_conv_silu_fwd + GEMMs.                                                                                          */
typedef struct {
  const void* xz; const void* w_x; const void* w_dt;
  const float* conv_w; const float* conv_b;
  const int32_t* seq_of_job; const int32_t* pset_of_job; const int32_t* rev_of_job;
  const void* halo;
  void* delta; float* bc;
  int64_t L, E, N, R;
  int64_t ldxz, ldd, ldbc;
  int32_t nseq, njobs, io_dtype;
  float* bcT;               /* optional (NULL): (njobs, ldT, 2N) fp32, the B / C rows TOKEN-major (scan variant 20), written for
                               tokens [0, min(ldT, ceil128(L))), zeros from L on; the caller zero-fills rows beyond ceil128(L) */
  int64_t ldT;              /* rows per job of bcT (ceil256(L)) */
  const void* w_x_packed;   /* optional (NULL): w_x rearranged per 32-channel slab into the K-major operand the tensor cores read,
                               (P, E/32, 4, 48, 8) io dtype = [slab][8-channel group][row][channel in group]; rows [0,R) = dt rows,
                               [R,16) = zero, [16,48) = B / C rows.  One 3 KB bulk copy per slab then replaces eight tensor-map
                               boxes of 16-byte rows (measured 14 % of the kernel).  The caller packs it once per weight version. */
} cad_conv_xproj_args;
int cad_conv_xproj_fwd(const cad_conv_xproj_args* a, void* stream);

/* unfused helper (fp32 I/O and the training path, which needs u for the x_proj weight gradient):
 * u = silu(conv(x)) materialised per job as the operand of a cuBLAS x_proj GEMM  (njobs, E, ldu). */
typedef struct {
  const void* xz; void* u;
  const float* conv_w; const float* conv_b;
  const int32_t* seq_of_job; const int32_t* pset_of_job; const int32_t* rev_of_job;
  const void* halo;
  int64_t L, E, ldxz, ldu;
  int32_t nseq, njobs, io_dtype;
} cad_conv_fwd_args;
int cad_conv_silu_fwd(const cad_conv_fwd_args* a, void* stream);

/* ---- "next" row N3: fused (RC-equivariant) LM head + masked cross-entropy.  Replaces
 *      logits = lm_head(hidden).float() over all positions (ref:caduceus/modeling_caduceus.py:474-476; RCPS head
 *      ref:caduceus/modeling_rcps.py:233-246) followed by cross_entropy / weighted_cross_entropy with ignore_index
 *      (ref:caduceus/modeling_caduceus.py:279-294,478-482): the fp32 (B, L, V) logits are never materialised, ignored rows cost one
 *      label load.  Wcat[v, c] = W[v, c] (c < D);  Wcat[v, D + c] = W[cmap[v], D - 1 - c] (RCPS), read in place.
 *        forward : loss_partial / wsum_partial (nblocks) — the caller sums them: loss = sum(loss_partial) / sum(wsum_partial);
 *                  lse (rows) saved for the backward (0 for ignored rows)
 *        backward: dloss_scale (1 device float) = dloss / sum(wsum);  dhidden (rows, width) io dtype, zero rows for ignored tokens;
 *                  dwcat_partial (nblocks, V, width) fp32 — the caller sums over blocks and folds the RC half back into (V, D).     */
typedef struct {
  const void* hidden;          /* (rows, width) pitch ldh, io dtype; width = D (Ph) or 2 D (RCPS) */
  const void* weight;          /* (V, D) table, io dtype */
  const int64_t* cmap;         /* (V) complement map, RCPS only */
  const int64_t* labels;       /* (rows) */
  const float* loss_weights;   /* (rows) or NULL */
  float* loss_partial; float* wsum_partial;   /* (nblocks) forward outputs */
  float* lse;                  /* (rows): forward output, backward input */
  const float* dloss_scale;    /* backward: 1 float on the device */
  void* dhidden;               /* backward: (rows, width) pitch lddh */
  float* dwcat_partial;        /* backward: (nblocks, V, width) */
  int64_t rows, D, V, width, ldh, lddh, ignore_index;
  int32_t rcps, io_dtype, nblocks;
} cad_head_ce_args;
int cad_head_ce_blocks(int64_t rows);            /* nblocks both passes want for `rows` */
int cad_head_ce_fwd(const cad_head_ce_args* a, void* stream);
int cad_head_ce_bwd(const cad_head_ce_args* a, void* stream);

/* ---- "next" row N2: GPU-side hg38 batch preparation, integer / byte work, bit-exact.
 *      raw (B, L) ASCII bytes of the FASTA slices -> data, target (B, L) int64:
 *        optional per-row string reverse complement (ref:src/dataloaders/utils/rc.py:17-26),
 *        char -> id through a 256-entry table (ref:caduceus/tokenization_caduceus.py:49-58; lower case folded),
 *        N -> PAD (ref:src/dataloaders/datasets/hg38_dataset.py:211-212),
 *        MLM masking from caller-supplied draws (ref:src/dataloaders/utils/mlm.py:4-32); masked == NULL: ids only. */
typedef struct {
  const uint8_t* raw; const uint8_t* rc_flags /* (B) or NULL */; const int32_t* char_to_id /* (256) */;
  const uint8_t* masked; const uint8_t* replaced; const uint8_t* random_sel; const int64_t* random_words;
  int64_t* data; int64_t* target;
  int64_t B, L;
  int64_t n_id, pad_id, mask_id;
} cad_hg38_batch_args;
int cad_hg38_batch_fwd(const cad_hg38_batch_args* a, void* stream);

/* ---- "next" row N4: windowed mean of last-layer hidden states around a variant — the pooling of the reference's
 *      VEP dump loop (arange + clamp + gather + mean, ref:vep_embeddings.py:278-311) on a VIEW of the stored tensor:
 *        V[b, r, c] = hidden[b, flip_len ? L-1-r : r, c0 + (flip_ch ? C-1-c : c)]
 *        out[b, c]  = mean over j in [idx_b - lo_half, idx_b + half] of V[b, clamp(j, 0, L-1), c]
 *      so that the RC view (ref:vep_embeddings.py:355-366: `.contiguous().flip(dims=[1, 2])` of half the channels) is read in place. */
typedef struct {
  const void* hidden;            /* (B, L, ldh) token rows, io dtype */
  const int64_t* variant_idx;    /* (B) on the device */
  void* out;                     /* (B, ldo), io dtype; columns [0, C) written */
  int64_t B, L, C, ldh, c0, ldo;
  int32_t lo_half, half, flip_len, flip_ch, io_dtype;
} cad_window_mean_args;
int cad_window_mean(const cad_window_mean_args* a, void* stream);

/* ---- micro-benchmarks of the pipes that bound the scan (MUFU ex2, FFMA), used by bench.py to quote
 *      "fraction of measured MUFU peak" beside the HBM fraction.  Writes ops/s.          */
int cad_microbench(int which /*0 = ex2.approx.f32, 1 = ffma, 2 = ex2+4ffma mix*/, double* ops_per_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* CADUCEUS_B200_H_ */

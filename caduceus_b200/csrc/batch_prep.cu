// GPU-side hg38 batch preparation: raw FASTA bytes -> (MLM input ids, MLM targets) in one
// integer kernel, replacing the per-sequence Python of the reference's dataloader:
//   string reverse complement   ref:src/dataloaders/utils/rc.py:17-26      (per-character Python loop)
//   character tokenizer         ref:caduceus/tokenization_caduceus.py:49-58,91-95   (upper-case fold, [UNK] = 6)
//   N -> [PAD]                  ref:src/dataloaders/datasets/hg38_dataset.py:211-212
//   MLM masking                 ref:src/dataloaders/utils/mlm.py:4-32      (80 % [MASK] / 10 % random / 10 % kept)
// Integer / byte work: bit-exact against the reference functions given the same random draws (the draws themselves
// are plain tensors made by torch — on the GPU in production, on the CPU generator in the parity test).
#include "common.cuh"

namespace cad {

__device__ __forceinline__ unsigned char complement_base(unsigned char c) {
  switch (c) {                       // STRING_COMPLEMENT_MAP, everything else maps to itself
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return c;
  }
}

__global__ void __launch_bounds__(256) hg38_batch_kernel(cad_hg38_batch_args a) {
  __shared__ int32_t lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = a.char_to_id[i];
  __syncthreads();
  const int64_t n = a.B * a.L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / a.L, l = i - b * a.L;
    const bool rc = a.rc_flags && a.rc_flags[b];
    unsigned char c = a.raw[b * a.L + (rc ? a.L - 1 - l : l)];
    if (rc) c = complement_base(c);
    int64_t id = lut[c];                                   // the LUT already folds lower case to upper case
    if (id == a.n_id) id = a.pad_id;                       // N is ignored by the loss
    int64_t data = id, target = id;
    if (a.masked) {                                        // MLM
      const bool m = a.masked[i] != 0;
      target = m ? id : a.pad_id;
      if (m) {
        if (a.replaced[i]) data = a.mask_id;
        else if (a.random_sel[i]) data = a.random_words[i];
      }
    }
    a.data[i] = data;
    if (a.target) a.target[i] = target;
  }
}

}  // namespace cad

extern "C" int cad_hg38_batch_fwd(const cad_hg38_batch_args* a, void* stream_) {
  using namespace cad;
  CAD_REQUIRE(a, "cad_hg38_batch_fwd: null argument block");
  CAD_REQUIRE(a->B >= 0 && a->L >= 0, "cad_hg38_batch_fwd: bad sizes");
  if (a->B * a->L == 0) return 0;
  CAD_REQUIRE(a->raw && a->char_to_id && a->data, "cad_hg38_batch_fwd: null pointer");
  CAD_REQUIRE(!a->masked || (a->replaced && a->random_sel && a->random_words && a->target),
              "cad_hg38_batch_fwd: MLM needs masked / replaced / random_sel / random_words / target together");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t n = a->B * a->L;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)(cad_sm_count() > 0 ? cad_sm_count() : 148) * 16;
  if (blocks > cap) blocks = cap;
  hg38_batch_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*a);
  CAD_LAUNCH_CHECK();
  return 0;
}

// Loss heads and metrics of the reference's training steps, fused (the callers' step).
//
//  citation scripts (train.py:151-160):  out = log_softmax(elu(logits), dim=1);  loss = nll_loss(out[idx], labels[idx]);
//      acc = (argmax(out[idx]) == labels[idx]).mean()      -- five ATen launches + two index gathers + two .item() syncs
//  PPI script (train_ppi.py:106-120):  loss = BCEWithLogitsLoss(mean)(logits, labels);  micro-F1 of (logits > 0) through
//      .cpu().numpy() + sklearn  -- a device->host copy of the whole batch per step
//
// Here: one kernel per head computes the loss sum and the integer counts the metric needs (correct predictions;
// true/false positives and false negatives) into a small device-resident stats block, a second kernel writes
// dL/dlogits.  Nothing leaves the device; the host reads the stats when (and if) it wants to print.
// Sums are accumulated in fp64 so the result does not depend on the order the atomics land in.
#include "common.cuh"

namespace gatk {

__device__ __forceinline__ float elu_h(float x) { return x > 0.f ? x : expm1f(x); }

// stats[0] += sum_i nll_i, stats[1] += #correct.   One warp per selected row, C <= 32 * 32 classes.
__global__ void nll_head_fwd_kernel(int64_t n_idx, const int64_t* __restrict__ idx, const float* __restrict__ logits,
                                    int64_t ld, const int64_t* __restrict__ labels, int C, double* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_idx) return;
  const int64_t row = idx ? idx[r] : r;
  const float* x = logits + row * ld;
  float m = -INFINITY;
  int arg = 0;
  for (int c = lane; c < C; c += 32) {
    const float v = elu_h(__ldg(x + c));
    if (v > m) { m = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {  // max with the smallest index among ties (torch.max semantics, utils.py:93)
    const float om = __shfl_xor_sync(FULL, m, o);
    const int oa = __shfl_xor_sync(FULL, arg, o);
    if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
  }
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(elu_h(__ldg(x + c)) - m);
  s = warp_sum(s);
  if (lane == 0) {
    const int64_t y = labels[row];
    const float vy = elu_h(__ldg(x + y));
    atomicAdd(stats, (double)((m + logf(s)) - vy));
    if (arg == (int)y) atomicAdd(stats + 1, 1.0);
  }
}

// dlogits[row, c] = scale * (softmax(elu(x))_c - [c == y]) * elu'(x_c) for the selected rows (dlogits zero-initialised)
__global__ void nll_head_bwd_kernel(int64_t n_idx, const int64_t* __restrict__ idx, const float* __restrict__ logits,
                                    int64_t ld, const int64_t* __restrict__ labels, int C, const float* __restrict__ gscale,
                                    float scale, float* __restrict__ dlogits, int64_t ldd) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_idx) return;
  const int64_t row = idx ? idx[r] : r;
  const float* x = logits + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, elu_h(__ldg(x + c)));
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(elu_h(__ldg(x + c)) - m);
  s = warp_sum(s);
  const float sc = scale * (gscale ? __ldg(gscale) : 1.f);
  const int y = (int)labels[row];
  for (int c = lane; c < C; c += 32) {
    const float xv = __ldg(x + c);
    const float v = elu_h(xv);
    const float p = expf(v - m) / s;
    const float dv = (p - (c == y ? 1.f : 0.f)) * sc;
    // the same row may be selected twice (idx with repeats): accumulate
    atomicAdd(dlogits + row * ldd + c, dv * (xv > 0.f ? 1.f : v + 1.f));
  }
}

// stats[0] += sum BCE, stats[1..3] += TP, FP, FN of (logit > 0) vs label (train_ppi.py:106-110)
__global__ void bce_f1_fwd_kernel(int64_t total, const float* __restrict__ logits, const float* __restrict__ labels,
                                  double* __restrict__ stats) {
  float loss = 0.f;
  int tp = 0, fp = 0, fn = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(logits + i), y = __ldg(labels + i);
    loss += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));  // BCEWithLogits, the numerically stable form torch uses
    const bool pred = x > 0.f, pos = y > 0.5f;
    tp += pred && pos;
    fp += pred && !pos;
    fn += !pred && pos;
  }
  loss = warp_sum(loss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_xor_sync(FULL, tp, o);
    fp += __shfl_xor_sync(FULL, fp, o);
    fn += __shfl_xor_sync(FULL, fn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(stats, (double)loss);
    if (tp) atomicAdd(stats + 1, (double)tp);
    if (fp) atomicAdd(stats + 2, (double)fp);
    if (fn) atomicAdd(stats + 3, (double)fn);
  }
}

__global__ void bce_bwd_kernel(int64_t total, const float* __restrict__ logits, const float* __restrict__ labels,
                               const float* __restrict__ gscale, float scale, float* __restrict__ dlogits) {
  const float sc = scale * (gscale ? __ldg(gscale) : 1.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(logits + i);
    dlogits[i] = (1.f / (1.f + expf(-x)) - __ldg(labels + i)) * sc;
  }
}

}  // namespace gatk

using namespace gatk;

extern "C" int gatk_nll_head_fwd(int64_t n_idx, const int64_t* idx, const float* logits, int64_t ld, const int64_t* labels,
                                 int C, double* stats, void* stream) {
  GATK_REQUIRE(logits && labels && stats && C >= 1 && ld >= C, "bad arguments");
  if (n_idx == 0) return 0;
  nll_head_fwd_kernel<<<(unsigned)((n_idx + 7) / 8), 256, 0, (cudaStream_t)stream>>>(n_idx, idx, logits, ld, labels, C, stats);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_nll_head_bwd(int64_t n_idx, const int64_t* idx, const float* logits, int64_t ld, const int64_t* labels,
                                 int C, const float* gscale, float scale, float* dlogits, int64_t ldd, void* stream) {
  GATK_REQUIRE(logits && labels && dlogits && C >= 1 && ld >= C && ldd >= C, "bad arguments");
  if (n_idx == 0) return 0;
  nll_head_bwd_kernel<<<(unsigned)((n_idx + 7) / 8), 256, 0, (cudaStream_t)stream>>>(n_idx, idx, logits, ld, labels, C, gscale,
                                                                                  scale, dlogits, ldd);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_bce_f1_fwd(int64_t total, const float* logits, const float* labels, double* stats, void* stream) {
  GATK_REQUIRE(logits && labels && stats && total >= 0, "bad arguments");
  if (total == 0) return 0;
  int64_t blocks = (total + 1023) / 1024;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  bce_f1_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total, logits, labels, stats);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_bce_bwd(int64_t total, const float* logits, const float* labels, const float* gscale, float scale,
                            float* dlogits, void* stream) {
  GATK_REQUIRE(logits && labels && dlogits && total >= 0, "bad arguments");
  if (total == 0) return 0;
  int64_t blocks = (total + 1023) / 1024;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  bce_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total, logits, labels, gscale, scale, dlogits);
  GATK_CHECK_LAUNCH();
  return 0;
}

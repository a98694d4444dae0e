// Library state (errors, device info) and the small node-/edge-wise kernels around the
// attention core: dropout masks, logits, da reduction, head combine, SpecialSpmm.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace gatk {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

// ------------------------------------------------------------------ dropout masks (Philox4x32-10, common.cuh)
__global__ void dropout_keep_kernel(uint8_t* __restrict__ keep, int64_t n, float p, uint64_t seed, uint64_t offset) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 outputs
  if (q * 4 >= n) return;
  const unsigned m = drop_keep4(seed, offset, q, p);  // 24-bit uniforms in [0,1): keep with probability 1-p
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = q * 4 + k;
    if (i < n) keep[i] = (m >> k) & 1u;
  }
}

__global__ void mask_scale_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ keep, float scale,
                                  float* __restrict__ y, int64_t ldy, int64_t rows, int64_t cols) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols, c = i - r * cols;
  y[r * ldy + c] = keep[i] ? x[r * ldx + c] * scale : 0.f;
}

// ------------------------------------------------------------------ logits (one warp per row)
template <int NV>
__global__ void logits_kernel(int64_t n, int H, int lph, int V, float* __restrict__ wh, int64_t ldw,
                              const uint8_t* __restrict__ keep, float inv_keep, const float* __restrict__ a_src,
                              const float* __restrict__ a_dst, float* __restrict__ f, float* __restrict__ g,
                              uint64_t seed, uint64_t offset, float p_drop) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float pf[NV], pg[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int slot = lane + 32 * v;
    pf[v] = pg[v] = 0.f;
    if (slot < V) {
      float* p = wh + row * ldw + slot * 4;
      float4 w = *reinterpret_cast<const float4*>(p);
      if (keep) {
        const uchar4 k = *reinterpret_cast<const uchar4*>(keep + row * (int64_t)(V * 4) + slot * 4);
        w.x = k.x ? w.x * inv_keep : 0.f;
        w.y = k.y ? w.y * inv_keep : 0.f;
        w.z = k.z ? w.z * inv_keep : 0.f;
        w.w = k.w ? w.w * inv_keep : 0.f;
        stg4(p, w);
      } else if (p_drop > 0.f) {  // the same decisions, evaluated here: element (row, 4 slot + k) of the [n, H*Dp] site
        const unsigned m = drop_keep4(seed, offset, row * V + slot, p_drop);
        w.x = (m & 1u) ? w.x * inv_keep : 0.f;
        w.y = (m & 2u) ? w.y * inv_keep : 0.f;
        w.z = (m & 4u) ? w.z * inv_keep : 0.f;
        w.w = (m & 8u) ? w.w * inv_keep : 0.f;
        stg4(p, w);
      }
      pf[v] = dot4(w, ldg4(a_src + slot * 4));
      pg[v] = dot4(w, ldg4(a_dst + slot * 4));
    }
  }
  head_reduce<NV>(pf, lph);
  head_reduce<NV>(pg, lph);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int slot = lane + 32 * v;
    if (slot < V && slot % lph == 0) {
      f[row * H + slot / lph] = pf[v];
      g[row * H + slot / lph] = pg[v];
    }
  }
}

// ------------------------------------------------------------------ da reduction (two deterministic stages)
constexpr int DA_BLOCKS = 592;
__global__ void da_partial_kernel(int64_t n, int H, int Dp, const float* __restrict__ wh, int64_t ldw,
                                  const float* __restrict__ df, const float* __restrict__ dg, float* __restrict__ part) {
  const int HD = H * Dp;
  const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
  for (int c = threadIdx.x; c < HD; c += blockDim.x) {
    const int h = c / Dp;
    float s = 0.f, d = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      const float w = __ldg(wh + r * ldw + c);
      s = fmaf(__ldg(df + r * H + h), w, s);
      d = fmaf(__ldg(dg + r * H + h), w, d);
    }
    part[(int64_t)blockIdx.x * 2 * HD + c] = s;
    part[(int64_t)blockIdx.x * 2 * HD + HD + c] = d;
  }
}
__global__ void da_final_kernel(int nblocks, int HD, const float* __restrict__ part, float* __restrict__ da_src,
                                float* __restrict__ da_dst) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * HD) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += part[(int64_t)b * 2 * HD + c];
  if (c < HD) da_src[c] = s; else da_dst[c - HD] = s;
}

// ------------------------------------------------------------------ head combine
__global__ void head_combine_kernel(int64_t n, int H, int D, int Dp, const float* __restrict__ in, int64_t ldi, int mode,
                                    float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (mode == 0) {
    if (i >= n * H * D) return;
    const int64_t r = i / (H * D);
    const int c = (int)(i - r * (H * D));
    out[i] = in[r * ldi + (c / D) * Dp + (c % D)];
  } else {
    if (i >= n * D) return;
    const int64_t r = i / D;
    const int d = (int)(i - r * D);
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += in[r * ldi + h * Dp + d];
    out[i] = s / (float)H;
  }
}
__global__ void head_combine_bwd_kernel(int64_t n, int H, int D, int Dp, const float* __restrict__ gout, int mode,
                                        float* __restrict__ gin, int64_t ldi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * H * Dp) return;
  const int64_t r = i / (H * Dp);
  const int c = (int)(i - r * (H * Dp));
  const int h = c / Dp, d = c % Dp;
  float v = 0.f;
  if (d < D) v = mode == 0 ? gout[r * (int64_t)(H * D) + h * D + d] : gout[r * D + d] / (float)H;
  gin[r * ldi + c] = v;
}

// ------------------------------------------------------------------ SpecialSpmm on raw COO (one warp per entry)
__global__ void spmm_coo_fwd_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                    const float* __restrict__ val, int64_t e, int64_t k, const float* __restrict__ b,
                                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= e) return;
  const int64_t r = row[i], c = col[i];
  const float v = val[i];
  for (int64_t d = lane; d < k; d += 32) atomicAdd(out + r * k + d, v * __ldg(b + c * k + d));
}
__global__ void spmm_coo_bwd_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                    const float* __restrict__ val, int64_t e, int64_t k, const float* __restrict__ b,
                                    const float* __restrict__ gout, float* __restrict__ grad_val,
                                    float* __restrict__ grad_b) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= e) return;
  const int64_t r = row[i], c = col[i];
  const float v = val[i];
  float s = 0.f;
  for (int64_t d = lane; d < k; d += 32) {
    const float go = __ldg(gout + r * k + d);
    if (grad_val) s = fmaf(go, __ldg(b + c * k + d), s);
    if (grad_b) atomicAdd(grad_b + c * k + d, v * go);
  }
  if (grad_val) {
    s = warp_sum(s);
    if (lane == 0) grad_val[i] = s;
  }
}

}  // namespace gatk

using namespace gatk;

extern "C" int gatk_version(void) { return GATK_VERSION; }
extern "C" int64_t gatk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* gatk_last_error(void) { return g_err; }
extern "C" int gatk_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

extern "C" int gatk_dropout_keep_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  GATK_REQUIRE(keep || n == 0, "null mask");
  GATK_REQUIRE(p >= 0.f && p < 1.f, "dropout p=%f out of [0,1)", p);
  if (n == 0) return 0;
  const int64_t groups = (n + 3) / 4;
  dropout_keep_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keep, n, p, seed, offset);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_mask_scale(const float* x, int64_t ldx, const uint8_t* keep, float scale, float* y, int64_t ldy,
                               int64_t rows, int64_t cols, void* stream) {
  GATK_REQUIRE(x && keep && y, "null pointer argument");
  const int64_t n = rows * cols;
  if (n == 0) return 0;
  mask_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, keep, scale, y, ldy, rows, cols);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_logits_fwd(int64_t n, int H, int Dp, float* wh, int64_t ldw, const uint8_t* keep_wh, float inv_keep,
                               const float* a_src, const float* a_dst, float* f, float* g, uint64_t seed, uint64_t offset,
                               float p_drop, void* stream) {
  GATK_REQUIRE(H >= 1 && H <= 32 && dp_ok(Dp), "bad head geometry H=%d Dp=%d", H, Dp);
  const int nv = nv_for(H, Dp);
  GATK_REQUIRE(nv > 0, "row too wide");
  GATK_REQUIRE(wh && a_src && a_dst && f && g && ldw % 4 == 0, "bad arguments");
  if (n == 0) return 0;
  const int lph = Dp / 4, V = H * lph;
  const unsigned grid = (unsigned)((n + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nv) {
    case 1: logits_kernel<1><<<grid, 256, 0, st>>>(n, H, lph, V, wh, ldw, keep_wh, inv_keep, a_src, a_dst, f, g, seed, offset, p_drop); break;
    case 2: logits_kernel<2><<<grid, 256, 0, st>>>(n, H, lph, V, wh, ldw, keep_wh, inv_keep, a_src, a_dst, f, g, seed, offset, p_drop); break;
    case 4: logits_kernel<4><<<grid, 256, 0, st>>>(n, H, lph, V, wh, ldw, keep_wh, inv_keep, a_src, a_dst, f, g, seed, offset, p_drop); break;
    case 8: logits_kernel<8><<<grid, 256, 0, st>>>(n, H, lph, V, wh, ldw, keep_wh, inv_keep, a_src, a_dst, f, g, seed, offset, p_drop); break;
    default: logits_kernel<16><<<grid, 256, 0, st>>>(n, H, lph, V, wh, ldw, keep_wh, inv_keep, a_src, a_dst, f, g, seed, offset, p_drop); break;
  }
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t gatk_da_workspace_floats(int H, int Dp) { return (size_t)DA_BLOCKS * 2 * H * Dp; }

extern "C" int gatk_da_reduce(int64_t n, int H, int Dp, const float* wh, int64_t ldw, const float* df, const float* dg,
                              float* da_src, float* da_dst, float* ws, void* stream) {
  GATK_REQUIRE(wh && df && dg && da_src && da_dst && ws, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int HD = H * Dp;
  int blocks = DA_BLOCKS;
  if (n < blocks) blocks = n > 0 ? (int)n : 1;
  da_partial_kernel<<<blocks, 256, 0, st>>>(n, H, Dp, wh, ldw, df, dg, ws);
  GATK_CHECK_LAUNCH();
  da_final_kernel<<<(2 * HD + 127) / 128, 128, 0, st>>>(blocks, HD, ws, da_src, da_dst);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_head_combine(int64_t n, int H, int D, int Dp, const float* in, int64_t ldi, int mode, float* out,
                                 void* stream) {
  GATK_REQUIRE(in && out && D <= Dp && (mode == 0 || mode == 1), "bad arguments");
  const int64_t total = mode == 0 ? n * H * D : n * D;
  if (total == 0) return 0;
  head_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, H, D, Dp, in, ldi, mode, out);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_head_combine_bwd(int64_t n, int H, int D, int Dp, const float* gout, int mode, float* gin,
                                     int64_t ldi, void* stream) {
  GATK_REQUIRE(gout && gin && D <= Dp && (mode == 0 || mode == 1), "bad arguments");
  const int64_t total = n * H * Dp;
  if (total == 0) return 0;
  head_combine_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, H, D, Dp, gout, mode, gin, ldi);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_spmm_coo_fwd(const int64_t* row, const int64_t* col, const float* val, int64_t e, int64_t k,
                                 const float* b, float* out, void* stream) {
  if (e == 0 || k == 0) return 0;
  GATK_REQUIRE(row && col && val && b && out, "null pointer argument");
  spmm_coo_fwd_kernel<<<(unsigned)((e + 7) / 8), 256, 0, (cudaStream_t)stream>>>(row, col, val, e, k, b, out);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_spmm_coo_bwd(const int64_t* row, const int64_t* col, const float* val, int64_t e, int64_t k,
                                 const float* b, const float* gout, float* grad_val, float* grad_b, void* stream) {
  if (e == 0 || k == 0) return 0;
  GATK_REQUIRE(row && col && val && b && gout, "null pointer argument");
  spmm_coo_bwd_kernel<<<(unsigned)((e + 7) / 8), 256, 0, (cudaStream_t)stream>>>(row, col, val, e, k, b, gout, grad_val, grad_b);
  GATK_CHECK_LAUNCH();
  return 0;
}

// Projection with PER-HEAD input dropout, all heads in one launch, masks never materialised.
//
// The reference builds one module per head and each draws its own dropout mask for the layer input
// (layers.py:34,132; the skip projection reuses that dropped input, :48,166), so with p > 0 the heads do not share
// an A operand: Z_h = (x * m_h / (1-p)) W_h.  Round 1 ran H mask-scale kernels + H (or 2H) GEMMs per direction in a
// Python loop over materialised uint8 masks.  Here the keep decision of element (h, i, k) is evaluated from the
// Philox stream (common.cuh: drop_keep4, the stream gatk_dropout_keep_mask materialises) at the point where x[i,k]
// is staged, for all heads in one launch per product.  fp32 FMA (exact products): the shapes that train with
// dropout are the citation graphs (Cora 2708 x 1433, Pubmed 19717 x 500), far too small for a tensor-core pipeline
// to matter; what mattered there was the launch count.
#include "common.cuh"

namespace gatk {

constexpr int HT_ROWS = 32;   // rows per CTA
constexpr int HT_K = 32;      // k-chunk staged per step

// block (row tile, head).  Threads: 256 = 32 rows x 8 column lanes; a thread owns row r and columns c = cl, cl+8, ...
// of the head's Dp-wide block (and of the skip block).
template <int MAXC>  // columns per thread (Dp <= 8 * MAXC)
__global__ void __launch_bounds__(256) heads_dropout_fwd_kernel(int64_t n, int F, int H, int Dp, int has_skip,
                                                                const float* __restrict__ x, int64_t ldx,
                                                                const float* __restrict__ W, int64_t ldw,
                                                                float* __restrict__ z, int64_t ldz, uint64_t seed,
                                                                uint64_t offset, float p, float inv_keep) {
  __shared__ float xs[HT_ROWS][HT_K + 1];
  const int h = blockIdx.y;
  const int64_t row0 = (int64_t)blockIdx.x * HT_ROWS;
  const int r = threadIdx.x >> 3, cl = threadIdx.x & 7;
  const int nblk = has_skip ? 2 : 1;
  float acc[2][MAXC];
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[b][c] = 0.f;
  for (int k0 = 0; k0 < F; k0 += HT_K) {
    // stage the dropped input tile: thread t loads 4 consecutive k of one row (one Philox call per 4 elements)
    {
      const int rr = threadIdx.x >> 3, kq = (threadIdx.x & 7) * 4;
      const int64_t row = row0 + rr;
#pragma unroll
      for (int j = 0; j < 4; ++j) xs[rr][kq + j] = 0.f;
      if (row < n) {
        const int64_t idx0 = ((int64_t)h * n + row) * F + k0 + kq;   // flat index into the [H, n, F] site
        unsigned m4 = 0xFu;
        const bool aligned = (idx0 & 3) == 0;   // F % 4 == 0: the four elements share one Philox counter
        if (p > 0.f && aligned && k0 + kq < F) m4 = drop_keep4(seed, offset, idx0 >> 2, p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + kq + j;
          if (k < F) {
            const bool keep = p > 0.f ? (aligned ? ((m4 >> j) & 1u) != 0 : drop_keep(seed, offset, idx0 + j, p)) : true;
            xs[rr][kq + j] = keep ? __ldg(x + row * ldx + k) * inv_keep : 0.f;
          }
        }
      }
    }
    __syncthreads();
    const int kmax = F - k0 < HT_K ? F - k0 : HT_K;
    for (int kk = 0; kk < kmax; ++kk) {
      const float xv = xs[r][kk];
      const float* wrow = W + (int64_t)(k0 + kk) * ldw + h * Dp;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int col = cl + 8 * c;
        if (col < Dp) {
          acc[0][c] = fmaf(xv, __ldg(wrow + col), acc[0][c]);
          if (nblk == 2) acc[1][c] = fmaf(xv, __ldg(wrow + H * Dp + col), acc[1][c]);
        }
      }
    }
    __syncthreads();
  }
  const int64_t row = row0 + r;
  if (row < n) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int col = cl + 8 * c;
      if (col < Dp) {
        z[row * ldz + h * Dp + col] = acc[0][c];
        if (nblk == 2) z[row * ldz + H * Dp + h * Dp + col] = acc[1][c];
      }
    }
  }
}

// dW partials: block (row chunk, head); a thread owns FOUR consecutive input columns k (one Philox counter when
// F % 4 == 0) and keeps the head's (up to 8 per pass) output columns in registers, so the keep decisions of
// (h, row, k..k+3) cost one Philox evaluation per pass.  part layout [chunk][F][M_out]; a second kernel sums the
// chunks in a fixed order.
__global__ void __launch_bounds__(256) heads_dropout_dw_kernel(int64_t n, int F, int H, int Dp, int has_skip, int64_t rows_per,
                                                               const float* __restrict__ x, int64_t ldx,
                                                               const float* __restrict__ dz, int64_t ldz,
                                                               float* __restrict__ part, uint64_t seed, uint64_t offset,
                                                               float p, float inv_keep) {
  constexpr int CP = 8;
  const int h = blockIdx.y;
  const int M_out = H * Dp * (has_skip ? 2 : 1);
  const int ncol = Dp * (has_skip ? 2 : 1);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
  float* out = part + (int64_t)blockIdx.x * F * M_out;
  const bool f4 = (F & 3) == 0;
  for (int k4 = threadIdx.x * 4; k4 < F; k4 += 1024) {
    for (int c0 = 0; c0 < ncol; c0 += CP) {
      float acc[4][CP];
      int col[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const int cc = c0 + c;
        col[c] = cc < ncol ? (cc < Dp ? 0 : H * Dp) + h * Dp + (cc < Dp ? cc : cc - Dp) : -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][c] = 0.f;
      }
      for (int64_t row = r0; row < r1; ++row) {
        const int64_t idx0 = ((int64_t)h * n + row) * F + k4;
        unsigned m4 = 0xFu;
        if (p > 0.f) {
          if (f4) {
            m4 = drop_keep4(seed, offset, idx0 >> 2, p);
          } else {
            m4 = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k4 + j < F && drop_keep(seed, offset, idx0 + j, p)) m4 |= 1u << j;
          }
        }
        if (!m4) continue;
        float xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = (k4 + j < F && ((m4 >> j) & 1u)) ? __ldg(x + row * ldx + k4 + j) : 0.f;
        const float* dr = dz + row * ldz;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          if (col[c] >= 0) {
            const float d = __ldg(dr + col[c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][c] = fmaf(xv[j], d, acc[j][c]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k4 + j < F)
#pragma unroll
          for (int c = 0; c < CP; ++c)
            if (col[c] >= 0) out[(int64_t)(k4 + j) * M_out + col[c]] = acc[j][c] * inv_keep;
    }
  }
}

__global__ void heads_dw_reduce_kernel(int64_t total, int chunks, int M_out, const float* __restrict__ part,
                                       float* __restrict__ dW, int64_t lddw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += part[(int64_t)c * total + i];
  dW[(i / M_out) * lddw + (i % M_out)] = s;
}

// dx[i,k] = inv_keep sum_h m_h[i,k] sum_{c in blk_h} dz[i,c] W[k,c].  Block = 8 rows; the rows' dz sit in shared memory.
__global__ void __launch_bounds__(256) heads_dropout_dx_kernel(int64_t n, int F, int H, int Dp, int has_skip,
                                                               const float* __restrict__ dz, int64_t ldz,
                                                               const float* __restrict__ W, int64_t ldw,
                                                               float* __restrict__ dx, int64_t lddx, uint64_t seed,
                                                               uint64_t offset, float p, float inv_keep) {
  extern __shared__ float dzs[];  // [8][M_out]
  const int M_out = H * Dp * (has_skip ? 2 : 1);
  const int64_t row0 = (int64_t)blockIdx.x * 8;
  for (int i = threadIdx.x; i < 8 * M_out; i += 256) {
    const int rr = i / M_out, c = i - rr * M_out;
    dzs[i] = row0 + rr < n ? __ldg(dz + (row0 + rr) * ldz + c) : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 8 * F; o += 256) {
    const int rr = o / F, k = o - rr * F;
    const int64_t row = row0 + rr;
    if (row >= n) continue;
    const float* wrow = W + (int64_t)k * ldw;
    const float* dr = dzs + rr * M_out;
    float s = 0.f;
    for (int h = 0; h < H; ++h) {
      const int64_t idx = ((int64_t)h * n + row) * F + k;
      const bool keep = p > 0.f ? drop_keep(seed, offset, idx, p) : true;
      if (!keep) continue;
      float t = 0.f;
      for (int c = 0; c < Dp; ++c) t = fmaf(dr[h * Dp + c], __ldg(wrow + h * Dp + c), t);
      if (has_skip)
        for (int c = 0; c < Dp; ++c) t = fmaf(dr[H * Dp + h * Dp + c], __ldg(wrow + H * Dp + h * Dp + c), t);
      s += t;
    }
    dx[row * lddx + k] = s * inv_keep;
  }
}

static int64_t dw_rows_per(int64_t n) {
  int64_t chunks = (n + 63) / 64;
  if (chunks > 128) chunks = 128;
  if (chunks < 1) chunks = 1;
  return (n + chunks - 1) / chunks;
}

}  // namespace gatk

using namespace gatk;

static int check_heads_args(int64_t n, int F, int H, int Dp, float p) {
  GATK_REQUIRE(n >= 0 && F >= 1 && H >= 1 && H <= 64 && Dp >= 1 && Dp <= 256, "bad sizes n=%lld F=%d H=%d Dp=%d", (long long)n, F, H, Dp);
  GATK_REQUIRE(p >= 0.f && p < 1.f, "dropout p=%f out of [0,1)", p);
  return 0;
}

extern "C" size_t gatk_gemm_heads_dropout_ws_floats(int64_t n, int F, int H, int Dp, int has_skip) {
  if (n <= 0) return 0;
  const int64_t rows_per = dw_rows_per(n);
  const int64_t chunks = (n + rows_per - 1) / rows_per;
  return (size_t)chunks * F * H * Dp * (has_skip ? 2 : 1);
}

extern "C" int gatk_gemm_heads_dropout_fwd(int64_t n, int F, int H, int Dp, int has_skip, const float* x, int64_t ldx,
                                           const float* W, int64_t ldw, float* z, int64_t ldz, uint64_t seed,
                                           uint64_t drop_offset, float p_drop, void* stream) {
  if (int rc = check_heads_args(n, F, H, Dp, p_drop)) return rc;
  GATK_REQUIRE(x && W && z && ldx >= F && ldw >= (int64_t)H * Dp * (has_skip ? 2 : 1) && ldz >= (int64_t)H * Dp * (has_skip ? 2 : 1), "bad arguments");
  if (n == 0) return 0;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  dim3 grid((unsigned)((n + HT_ROWS - 1) / HT_ROWS), (unsigned)H);
  cudaStream_t st = (cudaStream_t)stream;
#define HFWD(MC) heads_dropout_fwd_kernel<MC><<<grid, 256, 0, st>>>(n, F, H, Dp, has_skip, x, ldx, W, ldw, z, ldz, seed, drop_offset, p_drop, inv_keep)
  if (Dp <= 8) HFWD(1);
  else if (Dp <= 16) HFWD(2);
  else if (Dp <= 32) HFWD(4);
  else if (Dp <= 64) HFWD(8);
  else if (Dp <= 128) HFWD(16);
  else HFWD(32);
#undef HFWD
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_gemm_heads_dropout_dw(int64_t n, int F, int H, int Dp, int has_skip, const float* x, int64_t ldx,
                                          const float* dz, int64_t ldz, float* dW, int64_t lddw, float* ws, uint64_t seed,
                                          uint64_t drop_offset, float p_drop, void* stream) {
  if (int rc = check_heads_args(n, F, H, Dp, p_drop)) return rc;
  const int M_out = H * Dp * (has_skip ? 2 : 1);
  GATK_REQUIRE(x && dz && dW && ws && ldx >= F && ldz >= M_out && lddw >= M_out, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    GATK_CHECK_CUDA(cudaMemset2DAsync(dW, lddw * sizeof(float), 0, M_out * sizeof(float), F, st));
    return 0;
  }
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int64_t rows_per = dw_rows_per(n);
  const int chunks = (int)((n + rows_per - 1) / rows_per);
  dim3 grid((unsigned)chunks, (unsigned)H);
  heads_dropout_dw_kernel<<<grid, 256, 0, st>>>(n, F, H, Dp, has_skip, rows_per, x, ldx, dz, ldz, ws, seed, drop_offset, p_drop, inv_keep);
  GATK_CHECK_LAUNCH();
  const int64_t total = (int64_t)F * M_out;
  heads_dw_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, chunks, M_out, ws, dW, lddw);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_gemm_heads_dropout_dx(int64_t n, int F, int H, int Dp, int has_skip, const float* dz, int64_t ldz,
                                          const float* W, int64_t ldw, float* dx, int64_t lddx, uint64_t seed,
                                          uint64_t drop_offset, float p_drop, void* stream) {
  if (int rc = check_heads_args(n, F, H, Dp, p_drop)) return rc;
  const int M_out = H * Dp * (has_skip ? 2 : 1);
  GATK_REQUIRE(dz && W && dx && ldz >= M_out && ldw >= M_out && lddx >= F, "bad arguments");
  GATK_REQUIRE((size_t)8 * M_out * sizeof(float) <= 48 * 1024, "H*Dp=%d too wide for the dropout dx kernel", M_out);
  if (n == 0) return 0;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  heads_dropout_dx_kernel<<<(unsigned)((n + 7) / 8), 256, (size_t)8 * M_out * sizeof(float), (cudaStream_t)stream>>>(
      n, F, H, Dp, has_skip, dz, ldz, W, ldw, dx, lddx, seed, drop_offset, p_drop, inv_keep);
  GATK_CHECK_LAUNCH();
  return 0;
}

// fp32 SIMT GEMM (FFMA, 128x128x8 tiles, 8x8 register blocking, register-prefetch double
// buffering) with deterministic split-K.  It is the exact-fp32 projection path for operands
// the TMA/tcgen05 kernel cannot take (row strides that are not multiples of 16 bytes, e.g.
// Cora's 1433 or PPI's 50 input features) and the yardstick that kernel is tested against.
// Replaces torch.mm at layers.py:35,48,134,166 and its autograd.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gatk {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;

// Element (m,k) of op(A) / (k,n) of op(B) with zero fill outside the matrix.
template <bool T>
__device__ __forceinline__ float load_a(const float* A, int64_t lda, int64_t m, int64_t k, int64_t M, int64_t K1) {
  if (m >= M || k >= K1) return 0.f;
  return T ? __ldg(A + k * lda + m) : __ldg(A + m * lda + k);
}
template <bool T>
__device__ __forceinline__ float load_b(const float* B, int64_t ldb, int64_t k, int64_t n, int64_t N, int64_t K1) {
  if (n >= N || k >= K1) return 0.f;
  return T ? __ldg(B + n * ldb + k) : __ldg(B + k * ldb + n);
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A,
                                                    int64_t lda, const float* __restrict__ B, int64_t ldb,
                                                    float* __restrict__ C, int64_t ldc, int accumulate, int64_t kchunk,
                                                    float* __restrict__ part, int64_t ncol_tiles) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int64_t m0 = (tile / ncol_tiles) * BM, n0 = (tile % ncol_tiles) * BN;
  const int64_t k0 = (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < K ? k0 + kchunk : K;
  const int tx = tid & 15, ty = tid >> 4;

  // global->register staging indices: 4 elements of each operand tile per thread
  int am[4], ak[4], bk[4], bn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 256;
    if (TA) { am[i] = idx % BM; ak[i] = idx / BM; } else { am[i] = idx / BK; ak[i] = idx % BK; }
    if (TB) { bn[i] = idx / BK; bk[i] = idx % BK; } else { bn[i] = idx % BN; bk[i] = idx / BN; }
  }
  float ra[4], rb[4];
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto fetch = [&](int64_t kt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ra[i] = load_a<TA>(A, lda, m0 + am[i], kt + ak[i], M, k1);
      rb[i] = load_b<TB>(B, ldb, kt + bk[i], n0 + bn[i], N, k1);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][ak[i]][am[i]] = ra[i];
      Bs[buf][bk[i]][bn[i]] = rb[i];
    }
  };

  int buf = 0;
  if (k0 < k1) {
    fetch(k0);
    stash(0);
  }
  __syncthreads();
  for (int64_t kt = k0; kt < k1; kt += BK) {
    const bool more = kt + BK < k1;
    if (more) fetch(kt + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  float* out = part ? part + (int64_t)blockIdx.z * M * N : C;
  const int64_t ldo = part ? N : ldc;
  const bool add = part ? false : (accumulate != 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float* p = out + m * ldo + n;
      *p = add ? *p + acc[i][j] : acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(int64_t M, int64_t N, int splits, const float* __restrict__ part,
                                     float* __restrict__ C, int64_t ldc, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(int64_t)z * M * N + i];
  float* p = C + (i / N) * ldc + (i % N);
  *p = accumulate ? *p + s : s;
}

static int choose_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int64_t target = 2LL * sm_count();
  if (tiles >= target || K <= 512) return 1;
  int64_t s = target / tiles;
  const int64_t maxs = (K + 255) / 256;
  if (s > maxs) s = maxs;
  if (s > 512) s = 512;
  return s < 1 ? 1 : (int)s;
}

// tensor-core path (gemm_tc.cu)
bool gemm_tc_eligible(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                      const float* C, int64_t ldc, int accumulate);
size_t gemm_tc_workspace_bytes(int64_t N, int64_t K);
int gemm_tc_launch(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                   int64_t ldc, void* ws, size_t ws_bytes, cudaStream_t st);

bool gemm_tn_tc_eligible(int transA, int transB, int64_t Mo, int64_t No, int64_t K, const float* A, int64_t lda,
                         const float* B, int64_t ldb, int accumulate);
size_t gemm_tn_tc_workspace_bytes(int64_t Mo, int64_t No, int64_t K);
int gemm_tn_tc_launch(int64_t Mo, int64_t No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                      float* part, int* splits_out, cudaStream_t st);

bool gemm_long_tc_eligible(int transA, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda);
int gemm_long_tc_launch(int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                        int64_t ldb, float* C, int64_t ldc, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);

size_t gemm_batched_tc_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches);
int gemm_batched_path(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches, const float* A, int64_t lda,
                      int64_t a_bs, const float* B, int64_t ldb, int64_t b_bs, const float* C, int64_t ldc, int64_t c_bs);
int gemm_batched_tc_launch(int path, int transB, int64_t M, int64_t N, int64_t K, int batches, const float* A, int64_t lda,
                           int64_t a_bs, const float* B, int64_t ldb, int64_t b_bs, float* C, int64_t ldc, int64_t c_bs,
                           int epilogue, const float* elu_out, int64_t ld_elu, void* ws, size_t ws_bytes, cudaStream_t st);
bool gemm_batched_fuses_elu_grad(int path, int64_t N);

static bool tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATK_GEMM");
    v = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return v == 1;
}

}  // namespace gatk

using namespace gatk;

// Medium reductions with a K-major second operand (dx = dZ W^T of a hidden layer, K = H*D + 2H = 528 at 8 x 64) and NN
// products just past K = 512 run on the batched tensor-memory-A kernel as 64-wide n-tiles of one "batch" instead of
// the promoted long-K kernel.  Measured at the products hidden shape: dx 12.9 -> 8.1 ms.  The NN projection (K = 512)
// stays on gemm_tf32x3_kernel: 7.9 ms there vs 8.6 ms here -- with 64-wide tiles the A tile is re-read per n-tile
// (9 x 5 GB through L2), which is what bounds this path.
static bool use_tmem_a_path(int transA, int transB, int64_t M, int64_t N, int64_t K, int accumulate) {
  if (!tc_enabled() || transA || accumulate || N <= 64 || M < 1024 || K > 576) return false;
  return transB ? K > 128 : K > 512;
}

extern "C" size_t gatk_gemm_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K) {
  const int s = choose_splits(M, N, K);
  size_t simt = s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
  size_t tcb = (!transA && tc_enabled()) ? gemm_tc_workspace_bytes(N, K) : 0;
  if (use_tmem_a_path(transA, transB, M, N, K, 0)) {
    const size_t t = gemm_batched_tc_workspace_bytes(0, transB, M, N, K, 1);
    if (t > tcb) tcb = t;
  }
  if (transA && !transB && tc_enabled() && K >= 2048) {
    const size_t t = gemm_tn_tc_workspace_bytes(M, N, K);
    if (t > tcb) tcb = t;
  }
  return simt > tcb ? simt : tcb;
}

extern "C" int gatk_gemm_uses_tensor_cores(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda,
                                           int64_t ldc, int accumulate) {
  if (!tc_enabled()) return 0;
  if (transA && !transB)  // here lda is A's pitch and "ldc" carries B's pitch (both operands are TMA-loaded)
    return gemm_tn_tc_eligible(transA, transB, M, N, K, nullptr, lda, nullptr, ldc, accumulate) ? 1 : 0;
  if (gemm_tc_eligible(transA, transB, M, N, K, nullptr, lda, nullptr, ldc, accumulate)) return 1;
  return gemm_long_tc_eligible(transA, M, N, K, nullptr, lda) ? 1 : 0;
}

extern "C" int gatk_gemm(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                         const float* B, int64_t ldb, float* C, int64_t ldc, int accumulate, void* ws, size_t ws_bytes,
                         void* stream) {
  GATK_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative GEMM size");
  if (M == 0 || N == 0) return 0;
  GATK_REQUIRE(A && B && C, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tmem_a_path(transA, transB, M, N, K, accumulate) && ws && ws_bytes >= gemm_batched_tc_workspace_bytes(0, transB, M, N, K, 1) &&
      gemm_batched_path(0, transB, M, N, K, 1, A, lda, 0, B, ldb, 0, C, ldc, 0) == 1)
    return gemm_batched_tc_launch(1, transB, M, N, K, 1, A, lda, 0, B, ldb, 0, C, ldc, 0, 0, nullptr, 0, ws, ws_bytes, st);
  if (tc_enabled() && ws && ws_bytes >= gemm_tc_workspace_bytes(N, K) &&
      gemm_tc_eligible(transA, transB, M, N, K, A, lda, C, ldc, accumulate))
    return gemm_tc_launch(M, N, K, A, lda, B, ldb, C, ldc, ws, ws_bytes, st);
  if (tc_enabled() && ws && ws_bytes >= gemm_tc_workspace_bytes(N, K) && gemm_long_tc_eligible(transA, M, N, K, A, lda))
    return gemm_long_tc_launch(transB, M, N, K, A, lda, B, ldb, C, ldc, accumulate, ws, ws_bytes, st);
  if (tc_enabled() && gemm_tn_tc_eligible(transA, transB, M, N, K, A, lda, B, ldb, accumulate) && ws &&
      ws_bytes >= gemm_tn_tc_workspace_bytes(M, N, K)) {
    int sp = 1;
    if (int rc = gemm_tn_tc_launch(M, N, K, A, lda, B, ldb, static_cast<float*>(ws), &sp, st)) return rc;
    splitk_reduce_kernel<<<(unsigned)((M * N + 255) / 256), 256, 0, st>>>(M, N, sp, static_cast<float*>(ws), C, ldc, accumulate);
    GATK_CHECK_LAUNCH();
    return 0;
  }
  int splits = choose_splits(M, N, K);
  if (splits > 1 && (ws == nullptr || ws_bytes < (size_t)splits * M * N * sizeof(float))) splits = 1;
  int64_t kchunk = (K + splits - 1) / splits;
  kchunk = (kchunk + BK - 1) / BK * BK;
  if (kchunk < BK) kchunk = BK;
  splits = K > 0 ? (int)((K + kchunk - 1) / kchunk) : 1;
  const int64_t ncol = (N + BN - 1) / BN, nrow = (M + BM - 1) / BM;
  GATK_REQUIRE(nrow * ncol < (1LL << 31), "GEMM grid too large");
  dim3 grid((unsigned)(nrow * ncol), 1, (unsigned)splits);
  float* part = splits > 1 ? static_cast<float*>(ws) : nullptr;
  if (transA && transB)
    sgemm_kernel<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, ncol);
  else if (transA)
    sgemm_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, ncol);
  else if (transB)
    sgemm_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, ncol);
  else
    sgemm_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kchunk, part, ncol);
  GATK_CHECK_LAUNCH();
  if (splits > 1) {
    splitk_reduce_kernel<<<(unsigned)((M * N + 255) / 256), 256, 0, st>>>(M, N, splits, part, C, ldc, accumulate);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int gatk_elu_fwd(int64_t n, int64_t cols, float* buf, int64_t ld, void* stream);

extern "C" size_t gatk_gemm_batched_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches) {
  const size_t one = gatk_gemm_workspace_bytes(transA, transB, M, N, K);
  const size_t bat = tc_enabled() ? gemm_batched_tc_workspace_bytes(transA, transB, M, N, K, batches) : 0;
  return one > bat ? one : bat;
}

extern "C" int gatk_gemm_batched_fuses_elu_grad(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches, int64_t lda,
                                                int64_t a_bs, int64_t ldb, int64_t b_bs, int64_t ldc, int64_t c_bs) {
  if (!tc_enabled() || transA == transB) return 0;  // NT (dh' = A) and TN (dh' = B) products only
  const int path = gemm_batched_path(transA, transB, M, N, K, batches, nullptr, lda, a_bs, nullptr, ldb, b_bs, nullptr, ldc, c_bs);
  return (path && gemm_batched_fuses_elu_grad(path, N)) ? 1 : 0;
}

extern "C" int gatk_gemm_batched(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches, const float* A,
                                 int64_t lda, int64_t a_bs, const float* B, int64_t ldb, int64_t b_bs, float* C, int64_t ldc,
                                 int64_t c_bs, int epilogue, const float* elu_out, int64_t ld_elu, void* ws, size_t ws_bytes,
                                 void* stream) {
  GATK_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batches >= 0, "negative GEMM size");
  GATK_REQUIRE(epilogue == 0 || epilogue == 1, "epilogue must be 0 (none) or 1 (ELU)");
  if (M == 0 || N == 0 || batches == 0) return 0;
  GATK_REQUIRE(A && B && C, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int path = tc_enabled() ? gemm_batched_path(transA, transB, M, N, K, batches, A, lda, a_bs, B, ldb, b_bs, C, ldc, c_bs) : 0;
  if (path && ws && ws_bytes >= gemm_batched_tc_workspace_bytes(transA, transB, M, N, K, batches))
    return gemm_batched_tc_launch(path, transB, M, N, K, batches, A, lda, a_bs, B, ldb, b_bs, C, ldc, c_bs, epilogue,
                                  transA != transB ? elu_out : nullptr, ld_elu, ws, ws_bytes, st);
  GATK_REQUIRE(!elu_out, "elu_out needs the batched tensor-core kernels (gatk_gemm_batched_fuses_elu_grad says when)");
  for (int b = 0; b < batches; ++b) {  // shapes the batched tensor-core kernels do not take: one product per batch
    if (int rc = gatk_gemm(transA, transB, M, N, K, A + b * a_bs, lda, B + b * b_bs, ldb, C + b * c_bs, ldc, 0, ws, ws_bytes, stream))
      return rc;
    if (epilogue == 1)
      if (int rc = gatk_elu_fwd(M, N, C + b * c_bs, ldc, stream)) return rc;
  }
  return 0;
}

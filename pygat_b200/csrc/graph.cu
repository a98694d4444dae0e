// K0: adjacency pattern -> CSR in adj.nonzero() order (layers.py:129 / adj>0 layers.py:41),
// COO -> CSR, and the stable CSR transpose used by the scatter-free backward.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gatk {

__device__ __forceinline__ bool entry_set(float v, int rule) { return rule == 0 ? (v != 0.0f) : (v > 0.0f); }

// ---- column-contiguous storage: one warp per row, lanes sweep the columns (coalesced when col_stride == 1)
__global__ void dense_count_warp_per_row(const float* __restrict__ adj, int64_t n, int64_t rs, int64_t cs, int rule,
                                         int64_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* p = adj + row * rs;
  int cnt = 0;
  int64_t c = lane;
  for (; c + 96 < n; c += 128) {
    float v0 = __ldg(p + c * cs), v1 = __ldg(p + (c + 32) * cs), v2 = __ldg(p + (c + 64) * cs), v3 = __ldg(p + (c + 96) * cs);
    cnt += entry_set(v0, rule) + entry_set(v1, rule) + entry_set(v2, rule) + entry_set(v3, rule);
  }
  for (; c < n; c += 32) cnt += entry_set(__ldg(p + c * cs), rule);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  if (lane == 0) counts[row] = cnt;
}

__global__ void dense_fill_warp_per_row(const float* __restrict__ adj, int64_t n, int64_t rs, int64_t cs, int rule,
                                        const int64_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* p = adj + row * rs;
  int64_t pos = rowptr[row];
  for (int64_t c0 = 0; c0 < n; c0 += 32) {
    const int64_t c = c0 + lane;
    const bool hit = c < n && entry_set(__ldg(p + c * cs), rule);
    const unsigned m = __ballot_sync(FULL, hit);
    if (hit) col[pos + __popc(m & ((1u << lane) - 1u))] = (int32_t)c;
    pos += __popc(m);
  }
}

// ---- row-contiguous storage (utils.py:55 hands the layers a column-major matrix): one thread per
// row so that a warp reads 32 consecutive rows of one column = one coalesced 128-byte line.
constexpr int COL_CHUNK = 1024;
__global__ void dense_count_thread_per_row(const float* __restrict__ adj, int64_t n, int64_t rs, int64_t cs, int rule,
                                           unsigned long long* __restrict__ counts) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const int64_t c0 = (int64_t)blockIdx.y * COL_CHUNK;
  const int64_t c1 = c0 + COL_CHUNK < n ? c0 + COL_CHUNK : n;
  const float* p = adj + row * rs;
  int cnt = 0;
  int64_t c = c0;
  for (; c + 8 <= c1; c += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (c + u) * cs);
#pragma unroll
    for (int u = 0; u < 8; ++u) cnt += entry_set(v[u], rule);
  }
  for (; c < c1; ++c) cnt += entry_set(__ldg(p + c * cs), rule);
  if (cnt) atomicAdd(counts + row, (unsigned long long)cnt);
}

__global__ void dense_fill_thread_per_row(const float* __restrict__ adj, int64_t n, int64_t rs, int64_t cs, int rule,
                                          const int64_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const float* p = adj + row * rs;
  int64_t pos = rowptr[row];
  int64_t c = 0;
  for (; c + 8 <= n; c += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (c + u) * cs);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (entry_set(v[u], rule)) col[pos++] = (int32_t)(c + u);
  }
  for (; c < n; ++c)
    if (entry_set(__ldg(p + c * cs), rule)) col[pos++] = (int32_t)c;
}

__global__ void coo_count_kernel(const int64_t* __restrict__ row, int64_t e, unsigned long long* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e) atomicAdd(counts + row[i], 1ULL);
}
__global__ void coo_col_kernel(const int64_t* __restrict__ c64, int64_t e, int32_t* __restrict__ c32) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e) c32[i] = (int32_t)c64[i];
}
__global__ void col_hist_kernel(const int32_t* __restrict__ col, int64_t e, unsigned long long* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e) atomicAdd(counts + col[i], 1ULL);
}
// erow[e] = row of edge e, eid[e] = e
__global__ void expand_rows_kernel(const int64_t* __restrict__ rowptr, int64_t n, int32_t* __restrict__ erow,
                                   int32_t* __restrict__ eid) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t b = rowptr[row], en = rowptr[row + 1];
  for (int64_t k = b + lane; k < en; k += 32) {
    erow[k] = (int32_t)row;
    eid[k] = (int32_t)k;
  }
}
__global__ void gather_rows_kernel(const int32_t* __restrict__ erow, const int32_t* __restrict__ perm, int64_t e,
                                   int32_t* __restrict__ trow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e) trow[i] = erow[perm[i]];
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t scan_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (int64_t*)nullptr, (int64_t*)nullptr, n);
  return align256(bytes);
}
static size_t sort_temp_bytes(int64_t e, int bits) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, e, 0, bits);
  return align256(bytes);
}
static int key_bits(int64_t n) {
  int b = 1;
  while ((1LL << b) < n) ++b;
  return b;
}

static int inclusive_scan_inplace(int64_t* data, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st) {
  size_t need = scan_temp_bytes(n);
  GATK_REQUIRE(ws && ws_bytes >= need, "scan workspace too small: %zu < %zu", ws_bytes, need);
  GATK_CHECK_CUDA(cub::DeviceScan::InclusiveSum(ws, need, data, data, n, st));
  return 0;
}

}  // namespace gatk

using namespace gatk;

extern "C" size_t gatk_scan_workspace_bytes(int64_t n) { return scan_temp_bytes(n > 0 ? n : 1); }

extern "C" int gatk_csr_from_dense_rowptr(const float* adj, int64_t n, int64_t row_stride, int64_t col_stride, int rule,
                                          int64_t* rowptr, void* ws, size_t ws_bytes, void* stream) {
  GATK_REQUIRE(adj && rowptr && n >= 0, "bad arguments");
  GATK_REQUIRE(rule == 0 || rule == 1, "rule must be 0 (!=0) or 1 (>0)");
  cudaStream_t st = (cudaStream_t)stream;
  GATK_CHECK_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int64_t) * (n + 1), st));
  if (n == 0) return 0;
  const bool cols_contig = llabs(col_stride) <= llabs(row_stride);
  if (cols_contig) {
    dense_count_warp_per_row<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(adj, n, row_stride, col_stride, rule, rowptr + 1);
  } else {
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)((n + COL_CHUNK - 1) / COL_CHUNK));
    dense_count_thread_per_row<<<grid, 128, 0, st>>>(adj, n, row_stride, col_stride, rule,
                                                     reinterpret_cast<unsigned long long*>(rowptr + 1));
  }
  GATK_CHECK_LAUNCH();
  return inclusive_scan_inplace(rowptr + 1, n, ws, ws_bytes, st);
}

extern "C" int gatk_csr_from_dense_fill(const float* adj, int64_t n, int64_t row_stride, int64_t col_stride, int rule,
                                        const int64_t* rowptr, int32_t* col, void* stream) {
  GATK_REQUIRE(adj && rowptr && n >= 0, "bad arguments");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool cols_contig = llabs(col_stride) <= llabs(row_stride);
  if (cols_contig)
    dense_fill_warp_per_row<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(adj, n, row_stride, col_stride, rule, rowptr, col);
  else
    dense_fill_thread_per_row<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(adj, n, row_stride, col_stride, rule, rowptr, col);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_csr_from_coo(const int64_t* coo_row, const int64_t* coo_col, int64_t e, int64_t n, int64_t* rowptr,
                                 int32_t* col, void* ws, size_t ws_bytes, void* stream) {
  GATK_REQUIRE(rowptr && n >= 0 && e >= 0 && (e == 0 || (coo_row && coo_col && col)), "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GATK_CHECK_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int64_t) * (n + 1), st));
  if (n == 0) return 0;
  if (e > 0) {
    coo_count_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(coo_row, e, reinterpret_cast<unsigned long long*>(rowptr + 1));
    GATK_CHECK_LAUNCH();
    coo_col_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(coo_col, e, col);
    GATK_CHECK_LAUNCH();
  }
  return inclusive_scan_inplace(rowptr + 1, n, ws, ws_bytes, st);
}

extern "C" size_t gatk_transpose_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t e) {
  if (e < 1) e = 1;
  size_t scan = scan_temp_bytes(n_cols > 0 ? n_cols : 1);
  size_t sort = sort_temp_bytes(e, key_bits(n_cols > 1 ? n_cols : 2));
  size_t tmp = scan > sort ? scan : sort;
  return 3 * align256((size_t)e * sizeof(int32_t)) + tmp;
}

extern "C" int gatk_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t e, const int64_t* rowptr, const int32_t* col,
                                  int64_t* tptr, int32_t* trow, int32_t* perm, void* ws, size_t ws_bytes, void* stream) {
  GATK_REQUIRE(rowptr && tptr && n_rows >= 0 && n_cols >= 0 && e >= 0, "bad arguments");
  GATK_REQUIRE(e < (1LL << 31), "E=%lld does not fit int32 edge ids", (long long)e);
  cudaStream_t st = (cudaStream_t)stream;
  GATK_CHECK_CUDA(cudaMemsetAsync(tptr, 0, sizeof(int64_t) * (n_cols + 1), st));
  if (e == 0 || n_cols == 0) return 0;
  GATK_REQUIRE(col && trow && perm && ws, "null pointer argument");
  GATK_REQUIRE(ws_bytes >= gatk_transpose_workspace_bytes(n_rows, n_cols, e), "transpose workspace too small");
  const size_t seg = align256((size_t)e * sizeof(int32_t));
  char* base = static_cast<char*>(ws);
  int32_t* erow = reinterpret_cast<int32_t*>(base);
  int32_t* eid = reinterpret_cast<int32_t*>(base + seg);
  int32_t* skeys = reinterpret_cast<int32_t*>(base + 2 * seg);
  void* tmp = base + 3 * seg;
  const size_t tmp_bytes = ws_bytes - 3 * seg;

  col_hist_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(col, e, reinterpret_cast<unsigned long long*>(tptr + 1));
  GATK_CHECK_LAUNCH();
  if (int rc = inclusive_scan_inplace(tptr + 1, n_cols, tmp, tmp_bytes, st)) return rc;
  expand_rows_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(rowptr, n_rows, erow, eid);
  GATK_CHECK_LAUNCH();
  const int bits = key_bits(n_cols > 1 ? n_cols : 2);
  size_t sort_bytes = sort_temp_bytes(e, bits);
  GATK_REQUIRE(tmp_bytes >= sort_bytes, "sort workspace too small");
  GATK_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, sort_bytes, col, skeys, (const int32_t*)eid, perm, e, 0, bits, st));
  gather_rows_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(erow, perm, e, trow);
  GATK_CHECK_LAUNCH();
  return 0;
}

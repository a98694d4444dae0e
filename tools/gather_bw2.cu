// Follow-up microbenchmark: what caps random small-row gathers at ~8 G rows/s on B200?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <random>
#include <cstdint>
#include <algorithm>

// MODE 0: one row per warp instruction (lanes = float4 slots of the row).
// MODE 1: four rows per warp instruction (8 lanes x 16 B = 128 B of each of 4 different rows), row = 128 B.
template <int U, int MODE>
__global__ void gather_kernel(const float* __restrict__ x, int64_t ld, int slots, const int* __restrict__ idx, int64_t n_idx,
                              float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t base = warp * 32; base < n_idx; base += warps * 32) {
    const int my = base + lane < n_idx ? idx[base + lane] : 0;
    if (MODE == 0) {
      for (int t = 0; t < 32; t += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = __shfl_sync(0xffffffffu, my, t + u);
          const float* r = x + (int64_t)j * ld;
          v[u] = lane < slots ? __ldg(reinterpret_cast<const float4*>(r + lane * 4)) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
    } else {
      for (int t = 0; t < 32; t += 4 * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = __shfl_sync(0xffffffffu, my, t + 4 * u + (lane >> 3));
          const float* r = x + (int64_t)j * ld;
          v[u] = __ldg(reinterpret_cast<const float4*>(r + (lane & 7) * 4));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// TMA bulk copies: lane 0 issues one cp.async.bulk per row into a per-warp smem ring, all lanes consume.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int ROWS>
__global__ void bulk_kernel(const float* __restrict__ x, int64_t ld, int row_bytes, const int* __restrict__ idx, int64_t n_idx,
                            float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  unsigned char* buf = sm + (size_t)w * ROWS * 512;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (size_t)nw * ROWS * 512) + w;
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
  __syncwarp();
  const int64_t warp = (int64_t)blockIdx.x * nw + w;
  const int64_t warps = (int64_t)gridDim.x * nw;
  float4 acc = make_float4(0, 0, 0, 0);
  uint32_t phase = 0;
  for (int64_t base = warp * ROWS; base < n_idx; base += warps * ROWS) {
    const int my = (lane < ROWS && base + lane < n_idx) ? idx[base + lane] : 0;
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(ROWS * row_bytes) : "memory");
    __syncwarp();
    if (lane < ROWS) {
      const float* src = x + (int64_t)my * ld;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + lane * 512)),
                   "l"(src), "r"(row_bytes), "r"(smem_u32(bar))
                   : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
    phase ^= 1;
    for (int t = 0; t < ROWS; ++t) {
      if (lane * 16 < row_bytes) {
        float4 q = *reinterpret_cast<const float4*>(buf + t * 512 + lane * 16);
        acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
      }
    }
    __syncwarp();
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

int main() {
  const int64_t n = 2449029, e = 61861615;
  std::vector<int> h(e);
  std::mt19937 rng(1);
  for (auto& v : h) v = rng() % n;
  int* idx; cudaMalloc(&idx, e * 4); cudaMemcpy(idx, h.data(), e * 4, cudaMemcpyHostToDevice);
  float* x; cudaMalloc(&x, n * 512 + 4096); cudaMemset(x, 0, n * 512 + 4096);
  float* out; cudaMalloc(&out, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, double bytes_per_row) {
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-58s %.3f ms  %.1f GB/s useful  %.2f Grows/s (%s)\n", name, ms, e * bytes_per_row / ms / 1e6, e / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<8, 0><<<148 * 8, 256>>>(x, 128, 32, idx, e, out); cudaEventRecord(e1); }
  report("512B rows, 1 row/instr, 148 SMs x 64 warps", 512);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<8, 0><<<74, 256 * 4>>>(x, 128, 32, idx, e, out); cudaEventRecord(e1); }
  report("512B rows, 1 row/instr, 74 CTAs x 32 warps (half the SMs)", 512);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<8, 0><<<37, 256 * 4>>>(x, 128, 32, idx, e, out); cudaEventRecord(e1); }
  report("512B rows, 1 row/instr, 37 CTAs x 32 warps (quarter)", 512);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<8, 0><<<148 * 8, 256>>>(x, 32, 8, idx, e, out); cudaEventRecord(e1); }
  report("128B rows (pitch 128), 1 row/instr (8 lanes)", 128);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<2, 1><<<148 * 8, 256>>>(x, 32, 8, idx, e, out); cudaEventRecord(e1); }
  report("128B rows (pitch 128), 4 rows/instr", 128);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<8, 0><<<148 * 8, 256>>>(x, 128, 8, idx, e, out); cudaEventRecord(e1); }
  report("128B rows (pitch 512), 1 row/instr (8 lanes)", 128);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<8, 0><<<148 * 8, 256>>>(x, 128, 16, idx, e, out); cudaEventRecord(e1); }
  report("256B rows (pitch 512), 1 row/instr (16 lanes)", 256);
  // sorted-ish indices: locality test (same rows/s cap if per-instruction, faster if DRAM-page bound)
  {
    std::vector<int> hs(h);
    for (int64_t i = 0; i + 4096 <= e; i += 4096) std::sort(hs.begin() + i, hs.begin() + i + 4096);
    cudaMemcpy(idx, hs.data(), e * 4, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); gather_kernel<8, 0><<<148 * 8, 256>>>(x, 128, 32, idx, e, out); cudaEventRecord(e1); }
    report("512B rows, indices sorted within blocks of 4096", 512);
    cudaMemcpy(idx, h.data(), e * 4, cudaMemcpyHostToDevice);
  }
  {
    const int nw = 8; size_t smem = (size_t)nw * 16 * 512 + 64 * 8;
    cudaFuncSetAttribute(bulk_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); bulk_kernel<16><<<148 * 3, nw * 32, smem>>>(x, 128, 512, idx, e, out); cudaEventRecord(e1); }
    report("512B rows, cp.async.bulk 16 rows/warp, 24 warps/SM", 512);
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); bulk_kernel<16><<<148 * 3, nw * 32, smem>>>(x, 100, 400, idx, e, out); cudaEventRecord(e1); }
    report("400B rows (pitch 400), cp.async.bulk 16 rows/warp", 400);
    size_t smem32 = (size_t)nw * 32 * 512 + 64 * 8;
    cudaFuncSetAttribute(bulk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); bulk_kernel<32><<<148, nw * 32, smem32>>>(x, 128, 512, idx, e, out); cudaEventRecord(e1); }
    report("512B rows, cp.async.bulk 32 rows/warp, 8 warps/SM", 512);
  }
  return 0;
}

// Microbenchmark: random row-gather bandwidth on B200 as a function of row size / alignment / loads in flight.
// Each warp sums `per_warp` random rows; lanes own float4 slots.  nvcc -O3 -arch=sm_100a -o gather_bw gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <random>

template <int U>
__global__ void gather_kernel(const float* __restrict__ x, int64_t ld, int slots, const int* __restrict__ idx, int64_t n_idx,
                              float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4 acc = make_float4(0, 0, 0, 0);
  const int nl = (slots + 31) / 32;
  for (int64_t base = warp * 32; base < n_idx; base += warps * 32) {
    const int my = base + lane < n_idx ? idx[base + lane] : 0;
    for (int t = 0; t < 32; t += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = __shfl_sync(0xffffffffu, my, t + u);
        const float* r = x + (int64_t)j * ld;
        v[u] = make_float4(0, 0, 0, 0);
        for (int s = 0; s < nl; ++s) {
          const int sl = lane + 32 * s;
          if (sl < slots) {
            float4 q = __ldg(reinterpret_cast<const float4*>(r + sl * 4));
            v[u].x += q.x; v[u].y += q.y; v[u].z += q.z; v[u].w += q.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

int main(int argc, char** argv) {
  const int only_cfg = argc > 1 ? atoi(argv[1]) : -1, only_var = argc > 2 ? atoi(argv[2]) : -1;
  const int64_t n = 2449029, e = 61861615;
  std::vector<int> h(e);
  std::mt19937 rng(1);
  for (auto& v : h) v = rng() % n;
  int* idx; cudaMalloc(&idx, e * 4); cudaMemcpy(idx, h.data(), e * 4, cudaMemcpyHostToDevice);
  float* x; cudaMalloc(&x, n * 2048 + 4096); cudaMemset(x, 0, n * 2048 + 4096);
  float* out; cudaMalloc(&out, 16);
  struct Cfg { int slots; int ld; const char* name; } cfgs[] = {
      {25, 100, "400B rows, pitch 400"}, {25, 128, "400B rows, pitch 512"}, {32, 128, "512B rows, pitch 512"},
      {13, 52, "208B rows, pitch 208"}, {16, 64, "256B rows pitch 256"}, {64, 256, "1KB rows"}, {128, 512, "2KB rows"}};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int ci = -1;
  for (auto& c : cfgs) {
    ++ci;
    if (only_cfg >= 0 && ci != only_cfg) continue;
    for (int variant = 0; variant < 6; ++variant) {
      if (only_var >= 0 && variant != only_var) continue;
      int U = variant % 3 == 0 ? 4 : (variant % 3 == 1 ? 8 : 16);
      int threads = 256, bps = variant < 3 ? 4 : 8;
      int grid = 148 * bps;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (U == 4) gather_kernel<4><<<grid, threads>>>(x, c.ld, c.slots, idx, e, out);
        else if (U == 8) gather_kernel<8><<<grid, threads>>>(x, c.ld, c.slots, idx, e, out);
        else gather_kernel<16><<<grid, threads>>>(x, c.ld, c.slots, idx, e, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%-24s U=%2d warps/SM=%2d  %.3f ms  %.1f GB/s useful  %.2f Grows/s  (%s)\n", c.name, U, bps * 8, ms,
             (double)e * c.slots * 16 / ms / 1e6, e / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}

// Probe: TMA tile::gather4 (sm_100) -- which box shape does the tensor map need, and how are the 4 rows laid out in smem?
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int W, int r0, int r1, int r2, int r3, float* out, int* status) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t bar;
  const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4 * W + 64; i += 32) sm[i] = -1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4 * W * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(&map), "r"(b), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
  }
  unsigned done = 0, spins = 0;
  while (!done && spins < (1u << 22)) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0) : "memory");
    ++spins;
  }
  if (threadIdx.x == 0) status[0] = done ? 1 : -1;
  __syncwarp();
  for (int i = threadIdx.x; i < 4 * W + 64; i += 32) out[i] = sm[i];
}

int main() {
  const int n = 1000, W = 112, P = 128;
  std::vector<float> h((size_t)n * P);
  for (int i = 0; i < n; ++i) for (int k = 0; k < P; ++k) h[(size_t)i * P + k] = i + k * 0.001f;
  float* x; cudaMalloc(&x, h.size() * 4); cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, (4 * W + 64) * 4); int* st; cudaMalloc(&st, 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  for (int boxrows : {1, 4}) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)P * 4};
    cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)boxrows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box rows %d: encode rc=%d\n", boxrows, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    cudaMemset(st, 0, 4);
    probe<<<1, 32, (4 * W + 64) * 4>>>(map, W, 5, 900, 17, 333, out, st);
    cudaError_t e = cudaDeviceSynchronize();
    int hs = 0; cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
    std::vector<float> ho(4 * W + 64);
    cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
    printf("  launch: %s, mbarrier %s\n", cudaGetErrorString(e), hs == 1 ? "completed" : "TIMED OUT");
    for (int r = 0; r < 4; ++r) printf("  smem row %d: first=%.3f  [1]=%.3f last=%.3f\n", r, ho[r * W], ho[r * W + 1], ho[r * W + W - 1]);
    printf("  after: %.3f\n", ho[4 * W]);
    if (e != cudaSuccess) { cudaDeviceReset(); return 0; }
  }
  return 0;
}

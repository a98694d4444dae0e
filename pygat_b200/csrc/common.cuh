// Shared helpers for the gatk kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gatk.h"

namespace gatk {

void set_error(const char* fmt, ...);

#define GATK_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      gatk::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

void note_launch();  // every kernel launch of the library is counted (gatk_launch_count)

#define GATK_CHECK_LAUNCH()               \
  do {                                    \
    gatk::note_launch();                  \
    GATK_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

#define GATK_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      gatk::set_error(__VA_ARGS__);   \
      return 2;                       \
    }                                 \
  } while (0)

int sm_count();

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// 128-bit read-only gather of a feature row fragment.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// Streaming 128-bit load (no L1 allocation): data touched once per kernel.
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& w) {
  acc.x = fmaf(s, w.x, acc.x);
  acc.y = fmaf(s, w.y, acc.y);
  acc.z = fmaf(s, w.z, acc.z);
  acc.w = fmaf(s, w.w, acc.w);
}
__device__ __forceinline__ void scale4(float4& a, float s) {
  a.x *= s; a.y *= s; a.z *= s; a.w *= s;
}

// Slot geometry of a "rows" buffer: a row is V = H * LPH float4 slots, LPH = Dp / 4 slots
// per head, LPH a power of two.  Lane l of a warp owns slots l, l+32, ...  (coalesced
// 512-byte warp loads).  After head_reduce every lane holds, for each of its slots, the sum
// over all slots of that slot's head.
template <int NV>
__device__ __forceinline__ void head_reduce(float (&part)[NV], int lph) {
  const int q = lph >> 5;  // slots of one head held by the same lane (0 when the head is narrower than a warp)
#pragma unroll
  for (int st = 1; st < NV; st <<= 1) {
    if (st < q) {
      float t[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) t[v] = part[v] + part[v ^ st];
#pragma unroll
      for (int v = 0; v < NV; ++v) part[v] = t[v];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (o < lph) {
#pragma unroll
      for (int v = 0; v < NV; ++v) part[v] += __shfl_xor_sync(FULL, part[v], o);
    }
  }
}

// ------------------------------------------------------------------ dropout decisions (Philox4x32-10)
// One stream of keep bits per dropout site, keyed by (seed, offset): element i of the site takes word (i & 3) of
// Philox(seed)(offset + (i >> 2)) and is kept when its 24-bit uniform is >= p.  gatk_dropout_keep_mask materialises
// exactly this stream (that is how the in-kernel decisions are tested: same seed and offset -> same mask); the
// layer kernels evaluate it where the mask is consumed, so no mask ever exists in memory (F.dropout at
// layers.py:34,37,43 / :132,136,153).
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint64_t ctr) {
  uint4 c = make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u);
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ a, lo1, hi0 ^ c.w ^ b, lo0);
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ bool keep_from_word(uint32_t w, float p) { return ((w >> 8) * (1.0f / 16777216.0f)) >= p; }
// element i of the site
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint64_t offset, int64_t i, float p) {
  const uint4 r = philox4(seed, offset + (uint64_t)(i >> 2));
  const int k = (int)(i & 3);
  return keep_from_word(k == 0 ? r.x : (k == 1 ? r.y : (k == 2 ? r.z : r.w)), p);
}
// the four elements 4q .. 4q+3 of the site (q = i >> 2), as a 4-bit mask
__device__ __forceinline__ unsigned drop_keep4(uint64_t seed, uint64_t offset, int64_t q, float p) {
  const uint4 r = philox4(seed, offset + (uint64_t)q);
  return (keep_from_word(r.x, p) ? 1u : 0u) | (keep_from_word(r.y, p) ? 2u : 0u) | (keep_from_word(r.z, p) ? 4u : 0u) |
         (keep_from_word(r.w, p) ? 8u : 0u);
}

inline int nv_for(int H, int Dp) {
  int V = H * (Dp / 4);
  int nv = (V + 31) / 32;
  if (nv <= 1) return 1;
  if (nv <= 2) return 2;
  if (nv <= 4) return 4;
  if (nv <= 8) return 8;
  if (nv <= 16) return 16;
  return -1;
}

inline bool dp_ok(int Dp) {
  if (Dp < 4 || (Dp & 3)) return false;
  int l = Dp / 4;
  return (l & (l - 1)) == 0;
}

}  // namespace gatk

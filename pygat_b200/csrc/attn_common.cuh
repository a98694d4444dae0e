// Shared pieces of the CSR attention kernels: lane geometry, dynamic row scheduler, hub lookup.
#pragma once
#include <math.h>

#include "common.cuh"

namespace gatk {

constexpr int FWD_WARPS = 8;
constexpr int BWD_WARPS = 4;
constexpr int GRAB = 8;  // rows a warp claims per scheduler atomic

template <int NV>
struct LaneGeom {
  int hv[NV];       // head of slot v
  bool act[NV];     // slot exists
  bool leader[NV];  // first slot of its head
  __device__ __forceinline__ void init(int lane, int lph, int V) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      int slot = lane + 32 * v;
      act[v] = slot < V;
      hv[v] = act[v] ? slot / lph : 0;
      leader[v] = act[v] && (slot % lph == 0);
    }
  }
};

__device__ __forceinline__ int warp_grab(int32_t* counter, int lane, int step) {
  int r = 0;
  if (lane == 0) r = atomicAdd(counter, step);
  return __shfl_sync(FULL, r, 0);
}

// Per-segment scratch strides (floats), padded so the float4 slots stay 16-byte aligned.
__host__ __device__ __forceinline__ int64_t fwd_scratch_stride(int H, int V) { return V * 4 + ((2 * H + 3) & ~3); }
__host__ __device__ __forceinline__ int64_t src_scratch_stride(int H, int V) { return V * 4 + ((H + 3) & ~3); }

// Hub lookup: segment id -> (hub index, row, [beg,end)).
__device__ __forceinline__ void hub_locate(int seg, const int32_t* hub_rows, const int32_t* hub_seg_ptr, int n_hub,
                                           const int64_t* rowptr, int seg_len, int& row, int64_t& beg, int64_t& end) {
  int lo = 0, hi = n_hub;  // last hub with hub_seg_ptr[hub] <= seg
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (hub_seg_ptr[mid] <= seg) lo = mid; else hi = mid;
  }
  row = hub_rows[lo];
  int k = seg - hub_seg_ptr[lo];
  int64_t r0 = rowptr[row], r1 = rowptr[row + 1];
  beg = r0 + (int64_t)k * seg_len;
  end = beg + seg_len < r1 ? beg + seg_len : r1;
}


// Per-edge softmax terms are staged in shared memory in the order the consuming lanes want them:
// position (g, v) = g*NV + v holds the value for the head of slot  g*lph + 32*v, where g = lane / lph
// is the lane group (one group when a head is at least a warp wide).  A lane then reads the NV values
// of its own slots with one vector LDS.
struct SlotLayout {
  int G, WS;        // lane groups per 32 slots, floats per edge in the staging arrays
  int my_base;      // g * NV for this lane
  bool writer;      // first lane of its group: stores the group's dz values
  template <int NV>
  __device__ __forceinline__ void init(int lane, int lph) {
    G = lph < 32 ? 32 / lph : 1;
    WS = G * NV + 4;  // +4: keeps 16-byte alignment and spreads the per-edge rows over banks
    const int g = lph < 32 ? lane / lph : 0;
    my_base = g * NV;
    writer = lph < 32 ? (lane % lph == 0) : (lane == 0);
  }
  __host__ __device__ static int floats_per_edge(int lph, int NV) { return (lph < 32 ? 32 / lph : 1) * NV + 4; }
};

template <int NV>
__device__ __forceinline__ void lds_vec(const float* p, float (&o)[NV]) {
  if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int k = 0; k < NV / 4; ++k) {
      const float4 t = reinterpret_cast<const float4*>(p)[k];
      o[4 * k] = t.x; o[4 * k + 1] = t.y; o[4 * k + 2] = t.z; o[4 * k + 3] = t.w;
    }
  } else if constexpr (NV == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    o[0] = t.x; o[1] = t.y;
  } else {
    o[0] = p[0];
  }
}
template <int NV>
__device__ __forceinline__ void sts_vec(float* p, const float (&o)[NV]) {
  if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int k = 0; k < NV / 4; ++k)
      reinterpret_cast<float4*>(p)[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  } else if constexpr (NV == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]);
  } else {
    p[0] = o[0];
  }
}

template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, int* grid) {
  if (smem > 48 * 1024) GATK_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  GATK_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) {
    set_error("kernel does not fit on an SM (smem %zu)", smem);
    return 3;
  }
  *grid = per_sm * sm_count();
  return 0;
}

static int check_geom(int H, int Dp, int* nv) {
  GATK_REQUIRE(H >= 1 && H <= 32, "H=%d out of range [1,32]", H);
  GATK_REQUIRE(dp_ok(Dp), "Dp=%d must be 4*2^k", Dp);
  *nv = nv_for(H, Dp);
  GATK_REQUIRE(*nv > 0, "H*Dp=%d too wide (max 2048 floats per row)", H * Dp);
  return 0;
}

static int check_hub(int seg_len, int n_hub, int n_hub_seg, const void* rows, const void* ptr, const void* scratch) {
  GATK_REQUIRE(seg_len >= 1, "seg_len must be >= 1");
  GATK_REQUIRE(n_hub >= 0 && n_hub_seg >= 0, "negative hub counts");
  if (n_hub_seg > 0) GATK_REQUIRE(rows && ptr && scratch && n_hub > 0, "hub arrays missing");
  return 0;
}

#define NV_DISPATCH(nv, CALL)                     \
  switch (nv) {                                   \
    case 1: { constexpr int NV = 1; CALL; } break;   \
    case 2: { constexpr int NV = 2; CALL; } break;   \
    case 4: { constexpr int NV = 4; CALL; } break;   \
    case 8: { constexpr int NV = 8; CALL; } break;   \
    default: { constexpr int NV = 16; CALL; } break; \
  }

}  // namespace gatk

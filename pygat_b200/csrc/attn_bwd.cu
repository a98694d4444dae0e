// K3 / K4: backward of the fused attention, restructured so that the feature rows are gathered
// ONCE (the reference's autograd, like a direct transcription of it, gathers them twice: once
// for dL/d(alpha) = dh'_i . Wh_j along the destination rows and once for dWh_j = sum_i alpha_ij
// dh'_i along the source rows).  Walking the TRANSPOSED pattern, a warp that owns source row j
// keeps Wh_j in registers and gathers dh'_i for every destination i it feeds; the same gathered
// row serves both the dot product (-> dz_ij, the logit gradient) and the aggregation (-> dWh_j).
// Everything per-destination the softmax backward needs (f_i, lse_i, c_i = dh'_i . h_i) is a
// per-node scalar per head, gathered as 32-byte sectors.
//
//   prep   (rows)    dh' = gout * ELU'(out);  c_i = dh'_i . hagg_i
//   fused  (sources) dz_ij -> edge_dz[CSR edge id]; dWh_j = sum_i alpha~_ij dh'_i + dg_j a_dst;
//                    dg_j = sum_i dz_ij
//   finish (rows)    df_i = sum_j dz_ij (CSR segmented sum); dWh_i += df_i a_src; Wh-dropout mask
//
// Replaces autograd through layers.py:141-160 incl. SpecialSpmmFunction.backward (layers.py:81-90,
// a dense N x N product in the reference).
#include "attn_common.cuh"

namespace gatk {

// =====================================================================================
// prep: one warp per destination row
// =====================================================================================
template <int NV>
__global__ void attn_bwd_prep_kernel(int64_t n, int H, int lph, int V, const float* __restrict__ gout, int64_t ldgo,
                                     const float* __restrict__ out, int64_t ldo, int act_elu,
                                     const float* __restrict__ hagg, int64_t ldh, float* __restrict__ dhp,
                                     int64_t lddhp, float* __restrict__ c) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float part[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int slot = lane + 32 * v;
    part[v] = 0.f;
    if (slot < V) {
      const int off = slot * 4;
      float4 go = ldg4_stream(gout + row * ldgo + off);
      if (act_elu) {
        const float4 o = ldg4_stream(out + row * ldo + off);
        go.x *= o.x > 0.f ? 1.f : o.x + 1.f;
        go.y *= o.y > 0.f ? 1.f : o.y + 1.f;
        go.z *= o.z > 0.f ? 1.f : o.z + 1.f;
        go.w *= o.w > 0.f ? 1.f : o.w + 1.f;
      }
      stg4(dhp + row * lddhp + off, go);
      part[v] = dot4(go, ldg4_stream(hagg + row * ldh + off));
    }
  }
  head_reduce<NV>(part, lph);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int slot = lane + 32 * v;
    if (slot < V && slot % lph == 0) c[row * H + slot / lph] = part[v];
  }
}

// =====================================================================================
// fused source pass
// =====================================================================================
struct BwdFusedArgs {
  int64_t n_src;
  const int64_t* tptr;
  const int32_t* trow;
  const int32_t* perm;
  int H, Dp, lph, V, HP;
  const float* wh;
  int64_t ldw;
  const float* g;
  const float* f;
  const float* lse;
  const float* c;
  const uint8_t* keep;
  float inv_keep, alpha;
  const float* dhp;
  int64_t lddhp;
  const float* a_dst;
  float* dwh;
  int64_t lddwh;
  float* dg;
  float* edge_dz;
  int seg_len;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int n_hub, n_hub_seg;
  float* scratch;
  int32_t* counter;
  int ring;  // slots of the per-warp row ring
};

constexpr int FUSED_WARPS = 8;

// ---- bulk-async row gather ---------------------------------------------------------------------
// The kernel is bound by bytes in flight (Little's law: ~1.5-2 us loaded DRAM latency x 6.5 TB/s is
// ~70 KB per SM).  Register-staged LDG.128 gathers top out at 64 KB per SM; instead every gathered
// dh'_i row (H*Dp*4 bytes, contiguous) is fetched by ONE cp.async.bulk (the TMA unit's 1-D copy)
// into a per-warp ring of shared-memory slots, completion signalled on an mbarrier per slot.  Lane 0
// issues, the whole warp consumes with conflict-free LDS.128.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 28)) __trap();  // a protocol bug traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void bulk_row_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct FusedSmem {
  uint8_t* ring;   // R slots of row_bytes
  uint32_t bars;   // R mbarriers (shared-space address)
  int* row_s;
  int* perm_s;
  float* at_s;  // post-dropout attention (aggregation weight)
  float* A_s;   // alpha * slope * keep/(1-p)
  float* B_s;   // alpha * slope * c_i
  float* dz_s;
  int R;
  uint32_t row_bytes;
  uint32_t issued, consumed;  // rows issued / consumed by this warp since kernel start (warp-uniform)
  __host__ __device__ static size_t bytes_per_warp(int R, int V, int HP) {
    size_t b = (size_t)R * V * 16 + (size_t)R * 8 + 256 + (size_t)4 * 32 * HP * 4;
    return (b + 127) & ~(size_t)127;
  }
  __device__ __forceinline__ void carve(uint8_t* base, int R_, int V, int HP, int lane) {
    R = R_;
    row_bytes = (uint32_t)V * 16u;
    ring = base;
    uint8_t* p = base + (size_t)R * row_bytes;
    bars = smem_addr(p);
    p += (size_t)R * 8;
    row_s = reinterpret_cast<int*>(p);
    perm_s = row_s + 32;
    at_s = reinterpret_cast<float*>(p + 256);
    A_s = at_s + 32 * HP;
    B_s = A_s + 32 * HP;
    dz_s = B_s + 32 * HP;
    issued = consumed = 0;
    if (lane == 0) {
      for (int k = 0; k < R; ++k) mbar_init(bars + 8u * k, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  // lane 0 only
  __device__ __forceinline__ void issue(const float* src) {
    const uint32_t slot = issued % (uint32_t)R;
    mbar_expect_tx(bars + 8u * slot, row_bytes);
    bulk_row_copy(smem_addr(ring + (size_t)slot * row_bytes), src, row_bytes, bars + 8u * slot);
  }
};

template <int NV>
__device__ __forceinline__ void bwd_fused_segment(const BwdFusedArgs& a, int j, int64_t beg, int64_t end, int lane,
                                                  const LaneGeom<NV>& geo, float4 (&acc)[NV], float& dg_reg,
                                                  FusedSmem& sm) {
  const int H = a.H, HP = a.HP;
  float4 wj[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    wj[v] = geo.act[v] ? ldg4(a.wh + (int64_t)j * a.ldw + (lane + 32 * v) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float g_reg = lane < H ? __ldg(a.g + (int64_t)j * H + lane) : 0.f;
  dg_reg = 0.f;

  for (int64_t base = beg; base < end; base += 32) {
    const int cnt = (end - base) < 32 ? (int)(end - base) : 32;
    const bool valid = lane < cnt;
    const int64_t e = base + lane;
    const int i = valid ? __ldg(a.trow + e) : 0;
    const int pe = valid ? __ldg(a.perm + e) : 0;
    sm.row_s[lane] = i;
    sm.perm_s[lane] = pe;
    __syncwarp();
    // prime the ring with the first rows of this chunk; the weight computation below overlaps them
    int queued = 0;
    for (; queued < cnt && queued < sm.R; ++queued) {
      if (lane == 0) sm.issue(a.dhp + (int64_t)sm.row_s[queued] * a.lddhp);
      ++sm.issued;
    }
    const float* fi = a.f + (int64_t)i * H;
    const float* li = a.lse + (int64_t)i * H;
    const float* ci = a.c + (int64_t)i * H;
    const uint8_t* kp = a.keep ? a.keep + (int64_t)pe * H : nullptr;
    for (int h = 0; h < H; ++h) {
      const float gj = __shfl_sync(FULL, g_reg, h);
      float at = 0.f, A = 0.f, B = 0.f;
      if (valid) {
        const float z = __ldg(fi + h) + gj;
        const float s = z > 0.f ? z : a.alpha * z;
        const float al = expf(s - __ldg(li + h));
        const float slope = z > 0.f ? 1.f : a.alpha;
        const float kv = kp ? (kp[h] ? a.inv_keep : 0.f) : 1.f;
        at = al * kv;
        A = al * slope * kv;
        B = al * slope * __ldg(ci + h);
      }
      sm.at_s[lane * HP + h] = at;
      sm.A_s[lane * HP + h] = A;
      sm.B_s[lane * HP + h] = B;
    }
    __syncwarp();
    for (int t = 0; t < cnt; ++t) {
      const uint32_t slot = sm.consumed % (uint32_t)sm.R;
      mbar_wait(sm.bars + 8u * slot, (sm.consumed / (uint32_t)sm.R) & 1u);
      const float4* rowp = reinterpret_cast<const float4*>(sm.ring + (size_t)slot * sm.row_bytes) + lane;
      float pr[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        pr[v] = 0.f;
        if (geo.act[v]) {
          const float4 w = rowp[32 * v];
          fma4(acc[v], sm.at_s[t * HP + geo.hv[v]], w);
          pr[v] = dot4(w, wj[v]);
        }
      }
      __syncwarp();  // every lane has read the slot: it can be refilled
      ++sm.consumed;
      if (queued < cnt) {
        if (lane == 0) sm.issue(a.dhp + (int64_t)sm.row_s[queued] * a.lddhp);
        ++sm.issued;
        ++queued;
      }
      head_reduce<NV>(pr, a.lph);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (geo.leader[v]) {
          const int k = t * HP + geo.hv[v];
          sm.dz_s[k] = fmaf(sm.A_s[k], pr[v], -sm.B_s[k]);
        }
    }
    __syncwarp();
    // dz back to CSR edge order (each edge's H values are one contiguous sector), dg accumulation
    for (int idx = lane; idx < cnt * H; idx += 32) {
      const int tt = idx / H, h = idx - tt * H;
      a.edge_dz[(int64_t)sm.perm_s[tt] * H + h] = sm.dz_s[tt * HP + h];
    }
    for (int h = 0; h < H; ++h) {
      const float s = warp_sum(valid ? sm.dz_s[lane * HP + h] : 0.f);
      if (lane == h) dg_reg += s;
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void bwd_fused_store_slot(const BwdFusedArgs& a, int j, int slot, float4 r, float dgv) {
  fma4(r, dgv, ldg4(a.a_dst + slot * 4));
  stg4(a.dwh + (int64_t)j * a.lddwh + slot * 4, r);
}

template <int NV, bool HUB>
__global__ void __launch_bounds__(FUSED_WARPS * 32) attn_bwd_fused_kernel(const BwdFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem_fused[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  FusedSmem sm;
  sm.carve(smem_fused + warp * FusedSmem::bytes_per_warp(a.ring, a.V, a.HP), a.ring, a.V, a.HP, lane);
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  float4 acc[NV];
  float dg_reg;

  if (HUB) {
    const int seg = blockIdx.x * FUSED_WARPS + warp;
    if (seg >= a.n_hub_seg) return;
    int j;
    int64_t beg, end;
    hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.tptr, a.seg_len, j, beg, end);
    bwd_fused_segment<NV>(a, j, beg, end, lane, geo, acc, dg_reg, sm);
    float* sc = a.scratch + (int64_t)seg * src_scratch_stride(a.H, a.V);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (geo.act[v]) stg4(sc + (lane + 32 * v) * 4, acc[v]);
    if (lane < a.H) sc[a.V * 4 + lane] = dg_reg;
    return;
  }

  int cur = warp_grab(a.counter, lane);
  while (cur < a.n_src) {
    const int nxt = warp_grab(a.counter, lane);
    const int rend = cur + GRAB < a.n_src ? cur + GRAB : (int)a.n_src;
    for (int j = cur; j < rend; ++j) {
      const int64_t beg = a.tptr[j], end = a.tptr[j + 1];
      if (end - beg > a.seg_len) continue;
      bwd_fused_segment<NV>(a, j, beg, end, lane, geo, acc, dg_reg, sm);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float dgv = __shfl_sync(FULL, dg_reg, geo.hv[v]);
        if (geo.act[v]) bwd_fused_store_slot(a, j, lane + 32 * v, acc[v], dgv);
      }
      if (lane < a.H) a.dg[(int64_t)j * a.H + lane] = dg_reg;
    }
    cur = nxt;
  }
}

__global__ void attn_bwd_fused_hub_merge_kernel(const BwdFusedArgs a) {
  const int hub = blockIdx.x;
  const int j = a.hub_rows[hub];
  const int s0 = a.hub_seg_ptr[hub], s1 = a.hub_seg_ptr[hub + 1];
  const int64_t stride = src_scratch_stride(a.H, a.V);
  for (int slot = threadIdx.x; slot < a.V; slot += blockDim.x) {
    const int h = slot / a.lph;
    float dgv = 0.f;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = s0; s < s1; ++s) {
      const float* sc = a.scratch + s * stride;
      dgv += sc[a.V * 4 + h];
      const float4 t = *reinterpret_cast<const float4*>(sc + slot * 4);
      A.x += t.x; A.y += t.y; A.z += t.z; A.w += t.w;
    }
    bwd_fused_store_slot(a, j, slot, A, dgv);
    if (slot % a.lph == 0) a.dg[(int64_t)j * a.H + h] = dgv;
  }
}

// =====================================================================================
// finish: df (CSR segmented sum of dz) and the dst-side term of dWh
// =====================================================================================
struct FinishArgs {
  int64_t n;
  const int64_t* rowptr;
  int H, lph, V;
  const float* edge_dz;
  const float* a_src;
  const uint8_t* keep_wh;
  float inv_keep;
  float* dwh;
  int64_t lddwh;
  float* df;
  int seg_len;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int n_hub, n_hub_seg;
  float* scratch;
};

// sum over edges [beg,end) of dz[e][h]; lane h ends up holding head h's sum
__device__ __forceinline__ float segsum_heads(const float* __restrict__ dz, int H, int64_t beg, int64_t end, int lane) {
  float mine = 0.f;
  for (int h = 0; h < H; ++h) {
    float s0 = 0.f, s1 = 0.f;
    int64_t e = beg + lane;
    for (; e + 32 < end; e += 64) {
      s0 += __ldg(dz + e * H + h);
      s1 += __ldg(dz + (e + 32) * H + h);
    }
    if (e < end) s0 += __ldg(dz + e * H + h);
    const float s = warp_sum(s0 + s1);
    if (lane == h) mine = s;
  }
  return mine;
}

template <int NV>
__device__ __forceinline__ void finish_row(const FinishArgs& a, int64_t row, float df_reg, int lane) {
  if (lane < a.H) a.df[row * a.H + lane] = df_reg;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int slot = lane + 32 * v;
    const float dfv = __shfl_sync(FULL, df_reg, slot < a.V ? slot / a.lph : 0);
    if (slot < a.V) {
      float* p = a.dwh + row * a.lddwh + slot * 4;
      float4 r = *reinterpret_cast<const float4*>(p);
      fma4(r, dfv, ldg4(a.a_src + slot * 4));
      if (a.keep_wh) {
        const uchar4 k = *reinterpret_cast<const uchar4*>(a.keep_wh + row * (int64_t)(a.V * 4) + slot * 4);
        r.x = k.x ? r.x * a.inv_keep : 0.f;
        r.y = k.y ? r.y * a.inv_keep : 0.f;
        r.z = k.z ? r.z * a.inv_keep : 0.f;
        r.w = k.w ? r.w * a.inv_keep : 0.f;
      }
      stg4(p, r);
    }
  }
}

template <int NV>
__global__ void attn_bwd_finish_kernel(const FinishArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.n) return;
  const int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
  if (end - beg > a.seg_len) return;  // hub rows: segment kernels below
  finish_row<NV>(a, row, segsum_heads(a.edge_dz, a.H, beg, end, lane), lane);
}

__global__ void attn_bwd_finish_hub_seg_kernel(const FinishArgs a) {
  const int lane = threadIdx.x & 31;
  const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seg >= a.n_hub_seg) return;
  int row;
  int64_t beg, end;
  hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.rowptr, a.seg_len, row, beg, end);
  const float s = segsum_heads(a.edge_dz, a.H, beg, end, lane);
  if (lane < a.H) a.scratch[(int64_t)seg * a.H + lane] = s;
}

template <int NV>
__global__ void attn_bwd_finish_hub_merge_kernel(const FinishArgs a) {
  const int lane = threadIdx.x & 31;
  const int hub = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (hub >= a.n_hub) return;
  float s = 0.f;
  if (lane < a.H)
    for (int k = a.hub_seg_ptr[hub]; k < a.hub_seg_ptr[hub + 1]; ++k) s += a.scratch[(int64_t)k * a.H + lane];
  finish_row<NV>(a, a.hub_rows[hub], s, lane);
}

// =====================================================================================
// host side
// =====================================================================================
// Ring depth: as many row slots per warp as fit when two 8-warp CTAs share an SM (>= 2, <= 8).
static int fused_ring_slots(int V, int HP) {
  const size_t budget = 110 * 1024 / FUSED_WARPS;
  int r = 8;
  while (r > 2 && FusedSmem::bytes_per_warp(r, V, HP) > budget) --r;
  return r;
}

template <int NV>
static int launch_fused(BwdFusedArgs a, cudaStream_t st) {
  a.ring = fused_ring_slots(a.V, a.HP);
  const size_t smem = (size_t)FUSED_WARPS * FusedSmem::bytes_per_warp(a.ring, a.V, a.HP);
  GATK_REQUIRE(smem <= 227 * 1024, "row too wide for the gather ring (%zu bytes of shared memory)", smem);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_fused_kernel<NV, true><<<(a.n_hub_seg + FUSED_WARPS - 1) / FUSED_WARPS, FUSED_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
    attn_bwd_fused_hub_merge_kernel<<<a.n_hub, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_src > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_bwd_fused_kernel<NV, false>, FUSED_WARPS * 32, smem, &grid)) return rc;
    const int64_t need = (a.n_src + (int64_t)FUSED_WARPS * GRAB - 1) / ((int64_t)FUSED_WARPS * GRAB);
    if (need < grid) grid = (int)need;
    attn_bwd_fused_kernel<NV, false><<<grid, FUSED_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int NV>
static int launch_finish(const FinishArgs& a, cudaStream_t st) {
  if (a.n_hub_seg > 0) {
    attn_bwd_finish_hub_seg_kernel<<<(a.n_hub_seg + 7) / 8, 256, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
    attn_bwd_finish_hub_merge_kernel<NV><<<(a.n_hub + 7) / 8, 256, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n > 0) {
    attn_bwd_finish_kernel<NV><<<(unsigned)((a.n + 7) / 8), 256, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace gatk

using namespace gatk;

extern "C" size_t gatk_hub_scratch_floats(int which, int H, int Dp, int n_hub_seg) {
  const int V = H * (Dp / 4);
  if (n_hub_seg <= 0) return 0;
  if (which == 0) return (size_t)n_hub_seg * fwd_scratch_stride(H, V);
  if (which == 1) return (size_t)n_hub_seg * src_scratch_stride(H, V);
  return (size_t)n_hub_seg * H;
}

extern "C" int gatk_attn_bwd_prep(int64_t n, int H, int Dp, const float* gout, int64_t ldgo, const float* out,
                                  int64_t ldo, int act_elu, const float* hagg, int64_t ldh, float* dhp, int64_t lddhp,
                                  float* c, void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  GATK_REQUIRE(gout && hagg && dhp && c, "null pointer argument");
  GATK_REQUIRE(!act_elu || out, "out is required when act_elu is set");
  GATK_REQUIRE(ldgo % 4 == 0 && ldh % 4 == 0 && lddhp % 4 == 0 && (!act_elu || ldo % 4 == 0),
               "leading dims must be multiples of 4 floats");
  if (n == 0) return 0;
  const int lph = Dp / 4, V = H * lph;
  const unsigned grid = (unsigned)((n + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  NV_DISPATCH(nv, (attn_bwd_prep_kernel<NV><<<grid, 256, 0, st>>>(n, H, lph, V, gout, ldgo, out, ldo, act_elu, hagg,
                                                                  ldh, dhp, lddhp, c)));
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_attn_bwd_fused(int64_t n_src, const int64_t* tptr, const int32_t* trow, const int32_t* perm, int H,
                                   int Dp, const float* wh, int64_t ldw, const float* g, const float* f,
                                   const float* lse, const float* c, const uint8_t* keep_att, float inv_keep,
                                   float alpha, const float* dhp, int64_t lddhp, const float* a_dst, float* dwh,
                                   int64_t lddwh, float* dg, float* edge_dz, int seg_len, const int32_t* hub_rows,
                                   const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                                   int32_t* counter, void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_src < (1LL << 31), "n_src too large for one shard");
  GATK_REQUIRE(ldw % 4 == 0 && lddhp % 4 == 0 && lddwh % 4 == 0, "leading dims must be multiples of 4 floats");
  GATK_REQUIRE(tptr && wh && g && f && lse && c && dhp && a_dst && dwh && dg && counter, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  BwdFusedArgs a;
  a.n_src = n_src; a.tptr = tptr; a.trow = trow; a.perm = perm; a.H = H; a.Dp = Dp; a.lph = Dp / 4;
  a.V = H * (Dp / 4); a.HP = H | 1;
  a.wh = wh; a.ldw = ldw; a.g = g; a.f = f; a.lse = lse; a.c = c; a.keep = keep_att; a.inv_keep = inv_keep;
  a.alpha = alpha; a.dhp = dhp; a.lddhp = lddhp; a.a_dst = a_dst; a.dwh = dwh; a.lddwh = lddwh; a.dg = dg;
  a.edge_dz = edge_dz; a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub;
  a.n_hub_seg = n_hub_seg; a.scratch = hub_scratch; a.counter = counter;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  NV_DISPATCH(nv, return launch_fused<NV>(a, st));
  return 0;
}

extern "C" int gatk_attn_bwd_finish(int64_t n, const int64_t* rowptr, int H, int Dp, const float* edge_dz,
                                    const float* a_src, const uint8_t* keep_wh, float inv_keep, float* dwh,
                                    int64_t lddwh, float* df, int seg_len, const int32_t* hub_rows,
                                    const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                                    void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(rowptr && a_src && dwh && df && lddwh % 4 == 0, "bad arguments");
  FinishArgs a;
  a.n = n; a.rowptr = rowptr; a.H = H; a.lph = Dp / 4; a.V = H * (Dp / 4); a.edge_dz = edge_dz; a.a_src = a_src;
  a.keep_wh = keep_wh; a.inv_keep = inv_keep; a.dwh = dwh; a.lddwh = lddwh; a.df = df; a.seg_len = seg_len;
  a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg; a.scratch = hub_scratch;
  cudaStream_t st = (cudaStream_t)stream;
  NV_DISPATCH(nv, return launch_finish<NV>(a, st));
  return 0;
}

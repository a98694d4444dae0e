// K2: fused CSR attention forward.  One warp owns one destination row (segment); lanes own
// float4 slots of the H*Dp wide feature row, so every neighbour gather is a run of coalesced
// 512-byte warp loads.  Softmax weights are computed with lanes = edges on 32-edge chunks
// (online max / sum, layers.py:145-150) and staged through shared memory.  Rows longer than
// seg_len ("hubs" of a power-law graph) are cut into segments whose partial softmax states a
// small merge kernel combines.
//
// Reference semantics: layers.py:40-51 (dense class), layers.py:141-170 (sparse class).
#include "attn_common.cuh"

namespace gatk {

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

// =====================================================================================
// K2 forward
// =====================================================================================
struct FwdArgs {
  int64_t n_dst;
  const int64_t* rowptr;
  const int32_t* col;
  int H, Dp, lph, V, HP;
  const float* wh;
  int64_t ldw;
  const float* f;
  const float* g;
  int64_t ldf, ldg;  // row pitches of f and g (floats): H when they are dense [N,H] arrays
  const uint8_t* keep;
  float inv_keep, alpha;
  uint64_t seed, drop_offset;  // keep == NULL and p_drop > 0: attention-dropout decisions evaluated in the kernel
  float p_drop;
  const float* skipv;
  int64_t lds;
  int act_elu;
  float* hagg;
  float* out;
  int64_t ldo;
  float* lse;
  int seg_len;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int n_hub, n_hub_seg;
  float* scratch;
  int32_t* counter;
  const int32_t* item_ptr;  // edge-balanced work items: rows [item_ptr[k], item_ptr[k+1]) (NULL: 8 rows per grab)
  int n_items;
};

// Online-softmax aggregation of edges [beg,end) of destination `row` into (acc, m, l).
// m_reg / l_reg: lane h holds the running max / sum of head h.
template <int NV, bool FULLROW>
__device__ __forceinline__ void fwd_segment(const FwdArgs& a, int row, int64_t beg, int64_t end, int lane,
                                            const LaneGeom<NV>& geo, const SlotLayout& lay, float4 (&acc)[NV],
                                            float& m_reg, float& l_reg, int* col_s, float* p_s, float* scale_s) {
  constexpr int U = NV >= 8 ? 1 : 8 / NV;
  const int H = a.H, WS = lay.WS, lph = a.lph;
  const int q = lph >= 32 ? lph >> 5 : 1;
  float f_reg = lane < H ? __ldg(a.f + (int64_t)row * a.ldf + lane) : 0.f;
  m_reg = -INFINITY;
  l_reg = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int64_t base = beg; base < end; base += 32) {
    const int cnt = (end - base) < 32 ? (int)(end - base) : 32;
    const bool valid = lane < cnt;
    const int64_t e = base + lane;
    const int j = valid ? __ldg(a.col + e) : 0;
    col_s[lane] = j;
    const float* gj = a.g + (int64_t)j * a.ldg;
    const uint8_t* kp = a.keep ? a.keep + e * H : nullptr;
    for (int h = 0; h < H; ++h) {
      float fi = __shfl_sync(FULL, f_reg, h);
      float z = fi + (valid ? __ldg(gj + h) : 0.f);
      float s = z > 0.f ? z : a.alpha * z;
      s = valid ? s : -INFINITY;
      float cmax = warp_max(s);
      float m_old = __shfl_sync(FULL, m_reg, h);
      float m_new = fmaxf(m_old, cmax);
      float pe = valid ? expf(s - m_new) : 0.f;
      float csum = warp_sum(pe);
      const int pos = lph < 32 ? (h % lay.G) * NV + h / lay.G : h * q;
      if (lane == h) {
        float sc = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
        l_reg = l_reg * sc + csum;
        m_reg = m_new;
        for (int k = 0; k < q; ++k) scale_s[pos + k] = sc;
      }
      if (kp) pe = (valid && kp[h]) ? pe * a.inv_keep : 0.f;
      else if (a.p_drop > 0.f) pe = (valid && drop_keep(a.seed, a.drop_offset, e * H + h, a.p_drop)) ? pe * a.inv_keep : 0.f;
      for (int k = 0; k < q; ++k) p_s[lane * WS + pos + k] = pe;
    }
    __syncwarp();
    if (base != beg) {
      float sc[NV];
      lds_vec<NV>(scale_s + lay.my_base, sc);
#pragma unroll
      for (int v = 0; v < NV; ++v) scale4(acc[v], sc[v]);
    }
    const float* whl = a.wh + lane * 4;
    const float* pl = p_s + lay.my_base;
    int t = 0;
    for (; t + U <= cnt; t += U) {
      float4 w[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float* wj = whl + (int64_t)col_s[t + u] * a.ldw;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (FULLROW || geo.act[v]) w[u][v] = ldg4(wj + v * 128);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float pv[NV];
        lds_vec<NV>(pl + (t + u) * WS, pv);
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (FULLROW || geo.act[v]) fma4(acc[v], pv[v], w[u][v]);
      }
    }
    for (; t < cnt; ++t) {
      const float* wj = whl + (int64_t)col_s[t] * a.ldw;
      float pv[NV];
      lds_vec<NV>(pl + t * WS, pv);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (FULLROW || geo.act[v]) fma4(acc[v], pv[v], ldg4(wj + v * 128));
    }
    __syncwarp();
  }
}

// Divide by the row sum, save hagg / lse, add skip, ELU, store (layers.py:160-170).
__device__ __forceinline__ void fwd_store_slot(const FwdArgs& a, int row, int slot, float4 r, float l) {
  if (l > 0.f) {
    r.x /= l; r.y /= l; r.z /= l; r.w /= l;
  } else {
    r = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (a.hagg) stg4(a.hagg + (int64_t)row * (a.V * 4) + slot * 4, r);
  if (a.skipv) {
    float4 s = ldg4(a.skipv + (int64_t)row * a.lds + slot * 4);
    r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
  }
  if (a.act_elu) {
    r.x = elu1(r.x); r.y = elu1(r.y); r.z = elu1(r.z); r.w = elu1(r.w);
  }
  stg4(a.out + (int64_t)row * a.ldo + slot * 4, r);
}

template <int NV, bool HUB, bool FULLROW>
__global__ void __launch_bounds__(FWD_WARPS * 32) attn_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  SlotLayout lay = {};
  lay.init<NV>(lane, a.lph);
  const int per_warp = 32 + 32 * lay.WS + lay.WS;
  float* base_s = smem + warp * per_warp;
  int* col_s = reinterpret_cast<int*>(base_s);
  float* p_s = base_s + 32;
  float* scale_s = p_s + 32 * lay.WS;
  float4 acc[NV];
  float m_reg, l_reg;

  if (HUB) {
    int seg = blockIdx.x * FWD_WARPS + warp;
    if (seg >= a.n_hub_seg) return;
    int row;
    int64_t beg, end;
    hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.rowptr, a.seg_len, row, beg, end);
    fwd_segment<NV, FULLROW>(a, row, beg, end, lane, geo, lay, acc, m_reg, l_reg, col_s, p_s, scale_s);
    float* sc = a.scratch + (int64_t)seg * fwd_scratch_stride(a.H, a.V);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (geo.act[v]) stg4(sc + (lane + 32 * v) * 4, acc[v]);
    if (lane < a.H) {
      sc[a.V * 4 + lane] = m_reg;
      sc[a.V * 4 + a.H + lane] = l_reg;
    }
    return;
  }

  const int n_work = a.item_ptr ? a.n_items : (int)a.n_dst;
  const int step = a.item_ptr ? 1 : GRAB;
  int cur = warp_grab(a.counter, lane, step);
  while (cur < n_work) {
    int nxt = warp_grab(a.counter, lane, step);
    int rbeg, rend;
    if (a.item_ptr) {
      rbeg = a.item_ptr[cur];
      rend = a.item_ptr[cur + 1];
    } else {
      rbeg = cur;
      rend = cur + GRAB < a.n_dst ? cur + GRAB : (int)a.n_dst;
    }
    for (int row = rbeg; row < rend; ++row) {
      int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
      if (end - beg > a.seg_len) continue;  // hub: handled by the segment kernels
      fwd_segment<NV, FULLROW>(a, row, beg, end, lane, geo, lay, acc, m_reg, l_reg, col_s, p_s, scale_s);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float l = __shfl_sync(FULL, l_reg, geo.hv[v]);
        if (geo.act[v]) fwd_store_slot(a, row, lane + 32 * v, acc[v], l);
      }
      if (a.lse && lane < a.H) a.lse[(int64_t)row * a.H + lane] = l_reg > 0.f ? m_reg + logf(l_reg) : 0.f;
    }
    cur = nxt;
  }
}

// One CTA per hub row: merge the segment states (m_k, l_k, acc_k).
__global__ void attn_fwd_hub_merge_kernel(const FwdArgs a) {
  const int hub = blockIdx.x;
  const int row = a.hub_rows[hub];
  const int s0 = a.hub_seg_ptr[hub], s1 = a.hub_seg_ptr[hub + 1];
  const int64_t stride = fwd_scratch_stride(a.H, a.V);
  for (int slot = threadIdx.x; slot < a.V; slot += blockDim.x) {
    const int h = slot / a.lph;
    float M = -INFINITY;
    for (int s = s0; s < s1; ++s) M = fmaxf(M, a.scratch[s * stride + a.V * 4 + h]);
    float L = 0.f;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = s0; s < s1; ++s) {
      const float* sc = a.scratch + s * stride;
      float w = expf(sc[a.V * 4 + h] - M);
      L = fmaf(sc[a.V * 4 + a.H + h], w, L);
      fma4(A, w, *reinterpret_cast<const float4*>(sc + slot * 4));
    }
    fwd_store_slot(a, row, slot, A, L);
    if (a.lse && slot % a.lph == 0) a.lse[(int64_t)row * a.H + h] = L > 0.f ? M + logf(L) : 0.f;
  }
}

template <int NV, bool FULLROW>
static int launch_fwd_t(const FwdArgs& a, cudaStream_t st) {
  const int ws = SlotLayout::floats_per_edge(a.lph, NV);
  const size_t smem = (size_t)FWD_WARPS * (32 + 32 * ws + ws) * sizeof(float);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NV, true, FULLROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<NV, true, FULLROW><<<(a.n_hub_seg + FWD_WARPS - 1) / FWD_WARPS, FWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
    attn_fwd_hub_merge_kernel<<<a.n_hub, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_dst > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_fwd_kernel<NV, false, FULLROW>, FWD_WARPS * 32, smem, &grid)) return rc;
    int64_t need = a.item_ptr ? (a.n_items + FWD_WARPS - 1) / FWD_WARPS
                              : (a.n_dst + (int64_t)FWD_WARPS * GRAB - 1) / ((int64_t)FWD_WARPS * GRAB);
    if (need < grid) grid = (int)need;
    attn_fwd_kernel<NV, false, FULLROW><<<grid, FWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int NV>
static int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  return a.V == 32 * NV ? launch_fwd_t<NV, true>(a, st) : launch_fwd_t<NV, false>(a, st);
}

}  // namespace gatk

using namespace gatk;

extern "C" int gatk_attn_fwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp,
                             const float* wh, int64_t ldw, const float* f, int64_t ldf, const float* g, int64_t ldg,
                             const uint8_t* keep_att, float inv_keep, float alpha,
                             const float* skipv, int64_t lds, int act_elu,
                             float* hagg, float* out, int64_t ldo, float* lse,
                             int seg_len, const int32_t* hub_rows, const int32_t* hub_seg_ptr,
                             int n_hub, int n_hub_seg, float* hub_scratch, int32_t* counter,
                             const int32_t* item_ptr, int n_items, uint64_t seed, uint64_t drop_offset, float p_drop,
                             void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_dst < (1LL << 31), "n_dst too large for one shard");
  GATK_REQUIRE(ldw % 4 == 0 && ldo % 4 == 0 && (!skipv || lds % 4 == 0), "leading dims must be multiples of 4 floats");
  GATK_REQUIRE(rowptr && col && wh && f && g && out && counter, "null pointer argument");
  GATK_REQUIRE(ldf >= H && ldg >= H, "ldf / ldg must be >= H");
  cudaStream_t st = (cudaStream_t)stream;
  FwdArgs a;
  a.n_dst = n_dst; a.rowptr = rowptr; a.col = col; a.H = H; a.Dp = Dp; a.lph = Dp / 4; a.V = H * (Dp / 4);
  a.HP = H | 1;
  a.wh = wh; a.ldw = ldw; a.f = f; a.g = g; a.ldf = ldf; a.ldg = ldg; a.keep = keep_att; a.inv_keep = inv_keep; a.alpha = alpha;
  a.seed = seed; a.drop_offset = drop_offset; a.p_drop = keep_att ? 0.f : p_drop;
  a.skipv = skipv; a.lds = lds; a.act_elu = act_elu; a.hagg = hagg; a.out = out; a.ldo = ldo; a.lse = lse;
  a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg;
  a.scratch = hub_scratch; a.counter = counter; a.item_ptr = item_ptr; a.n_items = n_items;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  NV_DISPATCH(nv, return launch_fwd<NV>(a, st));
  return 0;
}


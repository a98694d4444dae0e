// K3 / K4: backward of the fused attention, restructured so that the feature rows are gathered
// ONCE (the reference's autograd, like a direct transcription of it, gathers them twice: once
// for dL/d(alpha) = dh'_i . Wh_j along the destination rows and once for dWh_j = sum_i alpha_ij
// dh'_i along the source rows).  Walking the TRANSPOSED pattern, a warp that owns source row j
// keeps Wh_j in registers and gathers dh'_i for every destination i it feeds; the same gathered
// row serves both the dot product (-> dz_ij, the logit gradient) and the aggregation (-> dWh_j).
//
// Measured on B200: a random 32-byte access costs about as much DRAM time as ~340 streamed bytes,
// so everything the softmax backward needs from destination i (f_i, lse_i, c_i = dh'_i . h_i, per
// head) is packed BEHIND the dh'_i row in one contiguous record; an edge touches DRAM once.
//
//   prep   (rows)    rec_i = [ dh'_i = gout * ELU'(out) | (f_i, lse_i, c_i, 0) per head ]
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
                                     const float* __restrict__ hagg, int64_t ldh, const float* __restrict__ f,
                                     int64_t ldf, const float* __restrict__ lse, float* __restrict__ rec, int64_t ldrec,
                                     float* __restrict__ dhp2, int64_t lddhp2) {
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
      const float4 hg = ldg4_stream(hagg + row * ldh + off);
      if (act_elu) {
        if (out) {  // ELU'(h') from the activated output: out > 0 ? 1 : out + 1
          const float4 o = ldg4_stream(out + row * ldo + off);
          go.x *= o.x > 0.f ? 1.f : o.x + 1.f;
          go.y *= o.y > 0.f ? 1.f : o.y + 1.f;
          go.z *= o.z > 0.f ? 1.f : o.z + 1.f;
          go.w *= o.w > 0.f ? 1.f : o.w + 1.f;
        } else {    // no skip term: h' is hagg itself, ELU'(h) = h > 0 ? 1 : exp(h); saves reading `out`
          go.x *= hg.x > 0.f ? 1.f : expf(hg.x);
          go.y *= hg.y > 0.f ? 1.f : expf(hg.y);
          go.z *= hg.z > 0.f ? 1.f : expf(hg.z);
          go.w *= hg.w > 0.f ? 1.f : expf(hg.w);
        }
      }
      stg4(rec + row * ldrec + off, go);
      if (dhp2) stg4(dhp2 + row * lddhp2 + off, go);
      part[v] = dot4(go, hg);
    }
  }
  head_reduce<NV>(part, lph);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int slot = lane + 32 * v;
    if (slot < V && slot % lph == 0) {
      const int h = slot / lph;
      stg4(rec + row * ldrec + V * 4 + h * 4, make_float4(__ldg(f + row * ldf + h), __ldg(lse + row * H + h), part[v], 0.f));
    }
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
  int H, Dp, lph, V;
  const float* wh;
  int64_t ldw;
  const float* g;
  int64_t ldg;       // row pitch of g (floats)
  const float* rec;  // [n_dst, ldrec]: dh'_i row followed by (f, lse, c, 0) per head
  int64_t ldrec;
  const uint8_t* keep;
  float inv_keep, alpha;
  uint64_t seed, drop_offset;  // keep == NULL and p_drop > 0: the forward's attention-dropout decisions, re-evaluated
  float p_drop;
  const float* a_dst;
  float* dwh;
  int64_t lddwh;
  float* dg;
  int64_t lddg;      // row pitch of dg (floats)
  float* edge_dz;
  float* df_acc;     // NULL, or [n_dst, lddf_acc] zero-initialised: df_i += dz_ij accumulated here with red.global.add
  int64_t lddf_acc;  //       (source-row shards: the entries of a destination row live on several ranks) instead of edge_dz
  int seg_len;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int n_hub, n_hub_seg;
  float* scratch;
  int32_t* counter;
  const int32_t* item_ptr;  // edge-balanced work items over the transposed rows (NULL: 8 rows per grab)
  int n_items;
};

constexpr int FUSED_WARPS = 4;

// One gathered dh' row: aggregation with the staged attention weights, dot product with the
// resident Wh_j, head reduction, dz = A * (dh'.Wh_j) - B for this lane's NV slots.
template <int NV, bool FULLROW>
__device__ __forceinline__ void fused_edge(const LaneGeom<NV>& geo, int lph, const float4 (&w)[NV],
                                           const float4 (&wj)[NV], const float* at_p, const float* A_p,
                                           const float* B_p, float* dz_p, bool writer, float4 (&acc)[NV],
                                           float (&dgacc)[NV]) {
  float at[NV], pr[NV], A[NV], B[NV];
  lds_vec<NV>(at_p, at);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (FULLROW || geo.act[v]) {
      fma4(acc[v], at[v], w[v]);
      pr[v] = dot4(w[v], wj[v]);
    } else {
      pr[v] = 0.f;
    }
  }
  head_reduce<NV>(pr, lph);
  lds_vec<NV>(A_p, A);
  lds_vec<NV>(B_p, B);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    pr[v] = fmaf(A[v], pr[v], -B[v]);
    dgacc[v] += pr[v];
  }
  if (writer) sts_vec<NV>(dz_p, pr);
}

template <int NV, bool FULLROW>
__device__ __forceinline__ void bwd_fused_segment(const BwdFusedArgs& a, int j, int64_t beg, int64_t end, int lane,
                                                  const LaneGeom<NV>& geo, const SlotLayout& lay, float4 (&acc)[NV],
                                                  float (&dgacc)[NV], int* row_s, int* perm_s, float* at_s, float* A_s,
                                                  float* B_s, float* dz_s) {
  constexpr int U = NV >= 8 ? 1 : (NV == 4 ? 2 : 4);  // gathered rows in flight per warp (deeper did not help)
  const int H = a.H, WS = lay.WS, lph = a.lph;
  float4 wj[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    dgacc[v] = 0.f;
    wj[v] = (FULLROW || geo.act[v]) ? ldg4(a.wh + (int64_t)j * a.ldw + (lane + 32 * v) * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float g_reg = lane < H ? __ldg(a.g + (int64_t)j * a.ldg + lane) : 0.f;
  const int tail0 = a.V * 4;
  const int q = lph >= 32 ? lph >> 5 : 1;

  for (int64_t base = beg; base < end; base += 32) {
    const int cnt = (end - base) < 32 ? (int)(end - base) : 32;
    const bool valid = lane < cnt;
    const int i = valid ? __ldg(a.trow + base + lane) : 0;
    const int pe = (valid && a.perm) ? __ldg(a.perm + base + lane) : 0;
    row_s[lane] = i;
    perm_s[lane] = pe;
    // ---- lanes = edges: softmax terms of every head, stored at the positions the slot lanes read
    const float* tl = a.rec + (int64_t)i * a.ldrec + tail0;
    const uint8_t* kp = a.keep ? a.keep + (int64_t)pe * H : nullptr;
    for (int h = 0; h < H; ++h) {
      const float gj = __shfl_sync(FULL, g_reg, h);
      float at = 0.f, A = 0.f, B = 0.f;
      if (valid) {
        const float4 t4 = ldg4(tl + 4 * h);  // (f_i, lse_i, c_i, 0)
        const float z = t4.x + gj;
        const float s = z > 0.f ? z : a.alpha * z;
        const float al = expf(s - t4.y);
        const float slope = z > 0.f ? 1.f : a.alpha;
        const float kv = kp ? (kp[h] ? a.inv_keep : 0.f)
                            : (a.p_drop > 0.f ? (drop_keep(a.seed, a.drop_offset, (int64_t)pe * H + h, a.p_drop) ? a.inv_keep : 0.f) : 1.f);
        at = al * kv;
        A = al * slope * kv;
        B = al * slope * t4.z;
      }
      const int pos = lph < 32 ? (h % lay.G) * NV + h / lay.G : h * q;
      for (int k = 0; k < q; ++k) {
        at_s[lane * WS + pos + k] = at;
        A_s[lane * WS + pos + k] = A;
        B_s[lane * WS + pos + k] = B;
      }
    }
    __syncwarp();
    // ---- lanes = slots: gather each destination's dh' row once
    const float* dl = a.rec + lane * 4;
    const int lo = lay.my_base;
    int t = 0;
    for (; t + U <= cnt; t += U) {
      float4 w[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float* di = dl + (int64_t)row_s[t + u] * a.ldrec;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (FULLROW || geo.act[v]) w[u][v] = ldg4(di + v * 128);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int o = (t + u) * WS + lo;
        fused_edge<NV, FULLROW>(geo, lph, w[u], wj, at_s + o, A_s + o, B_s + o, dz_s + o, lay.writer, acc, dgacc);
      }
    }
    for (; t < cnt; ++t) {
      float4 w[NV];
      const float* di = dl + (int64_t)row_s[t] * a.ldrec;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (FULLROW || geo.act[v]) w[v] = ldg4(di + v * 128);
      const int o = t * WS + lo;
      fused_edge<NV, FULLROW>(geo, lph, w, wj, at_s + o, A_s + o, B_s + o, dz_s + o, lay.writer, acc, dgacc);
    }
    __syncwarp();
    // ---- dz back to CSR edge order: each edge's H values are one contiguous sector
    for (int idx = lane; idx < cnt * H; idx += 32) {
      const int tt = idx / H, h = idx - tt * H;
      const int pos = lph < 32 ? (h % lay.G) * NV + h / lay.G : h * q;
      if (a.df_acc)
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a.df_acc + (int64_t)row_s[tt] * a.lddf_acc + h), "f"(dz_s[tt * WS + pos]) : "memory");
      else
        a.edge_dz[(int64_t)perm_s[tt] * H + h] = dz_s[tt * WS + pos];
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void bwd_fused_store_slot(const BwdFusedArgs& a, int j, int slot, float4 r, float dgv) {
  if (a.a_dst) fma4(r, dgv, ldg4(a.a_dst + slot * 4));  // NULL: the dg term is folded into the projection backward
  stg4(a.dwh + (int64_t)j * a.lddwh + slot * 4, r);
}

template <int NV, bool HUB, bool FULLROW>
__global__ void __launch_bounds__(FUSED_WARPS * 32, NV <= 4 ? 4 : 1) attn_bwd_fused_kernel(const BwdFusedArgs a) {
  extern __shared__ __align__(16) float smem_fused[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  SlotLayout lay = {};
  lay.init<NV>(lane, a.lph);
  float* base_s = smem_fused + warp * (64 + 4 * 32 * lay.WS);
  int* row_s = reinterpret_cast<int*>(base_s);
  int* perm_s = row_s + 32;
  float* at_s = base_s + 64;
  float* A_s = at_s + 32 * lay.WS;
  float* B_s = A_s + 32 * lay.WS;
  float* dz_s = B_s + 32 * lay.WS;
  float4 acc[NV];
  float dgacc[NV];

  if (HUB) {
    const int seg = blockIdx.x * FUSED_WARPS + warp;
    if (seg >= a.n_hub_seg) return;
    int j;
    int64_t beg, end;
    hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.tptr, a.seg_len, j, beg, end);
    bwd_fused_segment<NV, FULLROW>(a, j, beg, end, lane, geo, lay, acc, dgacc, row_s, perm_s, at_s, A_s, B_s, dz_s);
    float* sc = a.scratch + (int64_t)seg * src_scratch_stride(a.H, a.V);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (geo.act[v]) stg4(sc + (lane + 32 * v) * 4, acc[v]);
      if (geo.leader[v]) sc[a.V * 4 + geo.hv[v]] = dgacc[v];
    }
    return;
  }

  const int n_work = a.item_ptr ? a.n_items : (int)a.n_src;
  const int step = a.item_ptr ? 1 : GRAB;
  int cur = warp_grab(a.counter, lane, step);
  while (cur < n_work) {
    const int nxt = warp_grab(a.counter, lane, step);
    int rbeg, rend;
    if (a.item_ptr) {
      rbeg = a.item_ptr[cur];
      rend = a.item_ptr[cur + 1];
    } else {
      rbeg = cur;
      rend = cur + GRAB < a.n_src ? cur + GRAB : (int)a.n_src;
    }
    for (int j = rbeg; j < rend; ++j) {
      const int64_t beg = a.tptr[j], end = a.tptr[j + 1];
      if (end - beg > a.seg_len) continue;
      bwd_fused_segment<NV, FULLROW>(a, j, beg, end, lane, geo, lay, acc, dgacc, row_s, perm_s, at_s, A_s, B_s, dz_s);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (geo.act[v]) bwd_fused_store_slot(a, j, lane + 32 * v, acc[v], dgacc[v]);
        if (geo.leader[v]) a.dg[(int64_t)j * a.lddg + geo.hv[v]] = dgacc[v];
      }
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
    if (slot % a.lph == 0) a.dg[(int64_t)j * a.lddg + h] = dgv;
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
  uint64_t seed, drop_offset;  // keep_wh == NULL and p_drop > 0: the post-projection dropout decisions, re-evaluated
  float p_drop;
  float* dwh;
  int64_t lddwh;
  float* df;
  int64_t lddf;  // row pitch of df (floats)
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
  if (lane < a.H) a.df[row * a.lddf + lane] = df_reg;
  if (!a.a_src) return;  // the df term is folded into the projection backward: no row update here
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
      } else if (a.p_drop > 0.f) {
        const unsigned m = drop_keep4(a.seed, a.drop_offset, row * a.V + slot, a.p_drop);
        r.x = (m & 1u) ? r.x * a.inv_keep : 0.f;
        r.y = (m & 2u) ? r.y * a.inv_keep : 0.f;
        r.z = (m & 4u) ? r.z * a.inv_keep : 0.f;
        r.w = (m & 8u) ? r.w * a.inv_keep : 0.f;
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
template <int NV, bool FULLROW>
static int launch_fused_t(const BwdFusedArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)FUSED_WARPS * (64 + 4 * 32 * SlotLayout::floats_per_edge(a.lph, NV)) * sizeof(float);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<NV, true, FULLROW>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_fused_kernel<NV, true, FULLROW>
        <<<(a.n_hub_seg + FUSED_WARPS - 1) / FUSED_WARPS, FUSED_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
    attn_bwd_fused_hub_merge_kernel<<<a.n_hub, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_src > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_bwd_fused_kernel<NV, false, FULLROW>, FUSED_WARPS * 32, smem, &grid)) return rc;
    const int64_t need = a.item_ptr ? (a.n_items + FUSED_WARPS - 1) / FUSED_WARPS
                                    : (a.n_src + (int64_t)FUSED_WARPS * GRAB - 1) / ((int64_t)FUSED_WARPS * GRAB);
    if (need < grid) grid = (int)need;
    attn_bwd_fused_kernel<NV, false, FULLROW><<<grid, FUSED_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int NV>
static int launch_fused(const BwdFusedArgs& a, cudaStream_t st) {
  return a.V == 32 * NV ? launch_fused_t<NV, true>(a, st) : launch_fused_t<NV, false>(a, st);
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

extern "C" int64_t gatk_attn_bwd_record_ld(int H, int Dp) { return (int64_t)H * Dp + 4 * H; }

extern "C" int gatk_attn_bwd_prep(int64_t n, int H, int Dp, const float* gout, int64_t ldgo, const float* out,
                                  int64_t ldo, int act_elu, const float* hagg, int64_t ldh, const float* f,
                                  int64_t ldf, const float* lse, float* rec, int64_t ldrec, float* dhp2, int64_t lddhp2,
                                  void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  GATK_REQUIRE(gout && hagg && f && lse && rec, "null pointer argument");
  GATK_REQUIRE(ldgo % 4 == 0 && ldh % 4 == 0 && ldrec % 4 == 0 && ldrec >= (int64_t)H * Dp + 4 * H &&
                   (!out || ldo % 4 == 0) && (!dhp2 || lddhp2 % 4 == 0),
               "leading dims must be multiples of 4 floats (record: >= H*Dp + 4*H)");
  if (n == 0) return 0;
  const int lph = Dp / 4, V = H * lph;
  const unsigned grid = (unsigned)((n + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  NV_DISPATCH(nv, (attn_bwd_prep_kernel<NV><<<grid, 256, 0, st>>>(n, H, lph, V, gout, ldgo, out, ldo, act_elu, hagg,
                                                                  ldh, f, ldf, lse, rec, ldrec, dhp2, lddhp2)));
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_attn_bwd_fused(int64_t n_src, const int64_t* tptr, const int32_t* trow, const int32_t* perm, int H,
                                   int Dp, const float* wh, int64_t ldw, const float* g, int64_t ldg, const float* rec,
                                   int64_t ldrec, const uint8_t* keep_att, float inv_keep, float alpha,
                                   const float* a_dst, float* dwh, int64_t lddwh, float* dg, int64_t lddg,
                                   float* edge_dz, float* df_acc, int64_t lddf_acc,
                                   int seg_len, const int32_t* hub_rows, const int32_t* hub_seg_ptr, int n_hub,
                                   int n_hub_seg, float* hub_scratch, int32_t* counter, const int32_t* item_ptr,
                                   int n_items, uint64_t seed, uint64_t drop_offset, float p_drop, void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_src < (1LL << 31), "n_src too large for one shard");
  GATK_REQUIRE(ldw % 4 == 0 && ldrec % 4 == 0 && lddwh % 4 == 0, "leading dims must be multiples of 4 floats");
  GATK_REQUIRE(tptr && wh && g && rec && dwh && dg && (edge_dz || df_acc) && counter, "null pointer argument");
  GATK_REQUIRE(ldg >= H && lddg >= H && (!df_acc || lddf_acc >= H), "ldg / lddg / lddf_acc must be >= H");
  GATK_REQUIRE(df_acc || perm, "perm is required when dz goes to edge_dz");
  GATK_REQUIRE(!(df_acc && keep_att) && !(df_acc && p_drop > 0.f), "df_acc (source-row shards) does not take attention dropout");
  cudaStream_t st = (cudaStream_t)stream;
  BwdFusedArgs a;
  a.n_src = n_src; a.tptr = tptr; a.trow = trow; a.perm = perm; a.H = H; a.Dp = Dp; a.lph = Dp / 4;
  a.V = H * (Dp / 4);
  a.wh = wh; a.ldw = ldw; a.g = g; a.ldg = ldg; a.rec = rec; a.ldrec = ldrec; a.keep = keep_att; a.inv_keep = inv_keep;
  a.seed = seed; a.drop_offset = drop_offset; a.p_drop = keep_att ? 0.f : p_drop;
  a.alpha = alpha; a.a_dst = a_dst; a.dwh = dwh; a.lddwh = lddwh; a.dg = dg; a.lddg = lddg;
  a.edge_dz = edge_dz; a.df_acc = df_acc; a.lddf_acc = lddf_acc; a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub;
  a.n_hub_seg = n_hub_seg; a.scratch = hub_scratch; a.counter = counter; a.item_ptr = item_ptr; a.n_items = n_items;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  NV_DISPATCH(nv, return launch_fused<NV>(a, st));
  return 0;
}

extern "C" int gatk_attn_bwd_finish(int64_t n, const int64_t* rowptr, int H, int Dp, const float* edge_dz,
                                    const float* a_src, const uint8_t* keep_wh, float inv_keep, float* dwh,
                                    int64_t lddwh, float* df, int64_t lddf, int seg_len, const int32_t* hub_rows,
                                    const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                                    uint64_t seed, uint64_t drop_offset, float p_drop, void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(rowptr && df && lddf >= H && (!a_src || (dwh && lddwh % 4 == 0)) && (a_src || !keep_wh), "bad arguments");
  FinishArgs a;
  a.n = n; a.rowptr = rowptr; a.H = H; a.lph = Dp / 4; a.V = H * (Dp / 4); a.edge_dz = edge_dz; a.a_src = a_src;
  a.keep_wh = keep_wh; a.inv_keep = inv_keep; a.seed = seed; a.drop_offset = drop_offset;
  a.p_drop = (keep_wh || !a_src) ? 0.f : p_drop;
  a.dwh = dwh; a.lddwh = lddwh; a.df = df; a.lddf = lddf; a.seg_len = seg_len;
  a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg; a.scratch = hub_scratch;
  cudaStream_t st = (cudaStream_t)stream;
  NV_DISPATCH(nv, return launch_finish<NV>(a, st));
  return 0;
}

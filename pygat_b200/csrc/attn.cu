// K2 / K3 / K4: fused CSR attention kernels (forward, backward destination pass, backward
// source pass).  One warp owns one destination (or source) row segment; lanes own float4
// slots of the H*Dp wide feature row, so every neighbour gather is a run of coalesced
// 512-byte warp loads.  Softmax / weight computation runs with lanes = edges on 32-edge
// chunks, staged through shared memory.  Rows longer than seg_len ("hubs" of a power-law
// graph) are cut into segments whose partial states a small merge kernel combines.
//
// Reference semantics: layers.py:40-51 (dense), layers.py:141-170 (sparse), autograd of both.
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

__device__ __forceinline__ int warp_grab(int32_t* counter, int lane) {
  int r = 0;
  if (lane == 0) r = atomicAdd(counter, GRAB);
  return __shfl_sync(FULL, r, 0);
}

// Per-segment scratch strides (floats), padded so the float4 slots stay 16-byte aligned.
__host__ __device__ __forceinline__ int64_t fwd_scratch_stride(int H, int V) { return V * 4 + ((2 * H + 3) & ~3); }
__host__ __device__ __forceinline__ int64_t src_scratch_stride(int H, int V) { return V * 4 + ((H + 3) & ~3); }

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

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
  const uint8_t* keep;
  float inv_keep, alpha;
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
};

// Online-softmax aggregation of edges [beg,end) of destination `row` into (acc, m, l).
// m_reg / l_reg: lane h holds the running max / sum of head h.
template <int NV>
__device__ __forceinline__ void fwd_segment(const FwdArgs& a, int row, int64_t beg, int64_t end, int lane,
                                            const LaneGeom<NV>& geo, float4 (&acc)[NV], float& m_reg, float& l_reg,
                                            int* col_s, float* p_s, float* scale_s) {
  constexpr int U = NV >= 8 ? 1 : 8 / NV;
  const int H = a.H, HP = a.HP;
  float f_reg = lane < H ? __ldg(a.f + (int64_t)row * H + lane) : 0.f;
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
    const float* gj = a.g + (int64_t)j * H;
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
      if (lane == h) {
        float sc = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
        l_reg = l_reg * sc + csum;
        m_reg = m_new;
        scale_s[h] = sc;
      }
      if (kp) pe = (valid && kp[h]) ? pe * a.inv_keep : 0.f;
      p_s[lane * HP + h] = pe;
    }
    __syncwarp();
    if (base != beg) {
#pragma unroll
      for (int v = 0; v < NV; ++v) scale4(acc[v], scale_s[geo.hv[v]]);
    }
    const float* whl = a.wh + lane * 4;
    int t = 0;
    for (; t + U <= cnt; t += U) {
      float4 w[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float* wj = whl + (int64_t)col_s[t + u] * a.ldw;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (geo.act[v]) w[u][v] = ldg4(wj + v * 128);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (geo.act[v]) fma4(acc[v], p_s[(t + u) * HP + geo.hv[v]], w[u][v]);
      }
    }
    for (; t < cnt; ++t) {
      const float* wj = whl + (int64_t)col_s[t] * a.ldw;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (geo.act[v]) fma4(acc[v], p_s[t * HP + geo.hv[v]], ldg4(wj + v * 128));
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

template <int NV, bool HUB>
__global__ void __launch_bounds__(FWD_WARPS * 32) attn_fwd_kernel(const FwdArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = 32 + 32 * a.HP + 32;
  float* base_s = smem + warp * per_warp;
  int* col_s = reinterpret_cast<int*>(base_s);
  float* p_s = base_s + 32;
  float* scale_s = p_s + 32 * a.HP;
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  float4 acc[NV];
  float m_reg, l_reg;

  if (HUB) {
    int seg = blockIdx.x * FWD_WARPS + warp;
    if (seg >= a.n_hub_seg) return;
    int row;
    int64_t beg, end;
    hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.rowptr, a.seg_len, row, beg, end);
    fwd_segment<NV>(a, row, beg, end, lane, geo, acc, m_reg, l_reg, col_s, p_s, scale_s);
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

  int cur = warp_grab(a.counter, lane);
  while (cur < a.n_dst) {
    int nxt = warp_grab(a.counter, lane);
    int rend = cur + GRAB < a.n_dst ? cur + GRAB : (int)a.n_dst;
    for (int row = cur; row < rend; ++row) {
      int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
      if (end - beg > a.seg_len) continue;  // hub: handled by the segment kernels
      fwd_segment<NV>(a, row, beg, end, lane, geo, acc, m_reg, l_reg, col_s, p_s, scale_s);
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

// =====================================================================================
// K3 backward, destination pass
// =====================================================================================
struct BwdDstArgs {
  int64_t n_dst;
  const int64_t* rowptr;
  const int32_t* col;
  int H, Dp, lph, V, HP;
  const float* wh;
  int64_t ldw;
  const float* f;
  const float* g;
  const float* lse;
  const uint8_t* keep;
  float inv_keep, alpha;
  const float* gout;
  int64_t ldgo;
  const float* out;
  int64_t ldo;
  int act_elu;
  const float* hagg;
  int64_t ldh;
  float* dhp;
  int64_t lddhp;
  float* df;
  float* edge_alpha;
  float* edge_dz;
  int seg_len;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int n_hub, n_hub_seg;
  float* scratch;
  int32_t* counter;
};

template <int NV>
__device__ __forceinline__ void bwd_dst_segment(const BwdDstArgs& a, int row, int64_t beg, int64_t end, bool write_dhp,
                                                int lane, const LaneGeom<NV>& geo, float& df_reg, int* col_s,
                                                float* A_s, float* B_s, float* at_s, float* dz_s, float* c_s) {
  constexpr int U = NV >= 4 ? 1 : 4 / NV;
  const int H = a.H, HP = a.HP;
  // ---- row prologue: dh' = gout * ELU'(h'), c_h = dh'_h . hagg_h
  float4 dh[NV];
  float part[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    dh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    part[v] = 0.f;
    if (geo.act[v]) {
      const int off = (lane + 32 * v) * 4;
      float4 go = ldg4(a.gout + (int64_t)row * a.ldgo + off);
      if (a.act_elu) {
        float4 o = ldg4(a.out + (int64_t)row * a.ldo + off);
        go.x *= o.x > 0.f ? 1.f : o.x + 1.f;
        go.y *= o.y > 0.f ? 1.f : o.y + 1.f;
        go.z *= o.z > 0.f ? 1.f : o.z + 1.f;
        go.w *= o.w > 0.f ? 1.f : o.w + 1.f;
      }
      dh[v] = go;
      if (write_dhp) stg4(a.dhp + (int64_t)row * a.lddhp + off, go);
      part[v] = dot4(go, ldg4(a.hagg + (int64_t)row * a.ldh + off));
    }
  }
  head_reduce<NV>(part, a.lph);
#pragma unroll
  for (int v = 0; v < NV; ++v)
    if (geo.leader[v]) c_s[geo.hv[v]] = part[v];
  __syncwarp();
  const float c_reg = lane < H ? c_s[lane] : 0.f;
  const float f_reg = lane < H ? __ldg(a.f + (int64_t)row * H + lane) : 0.f;
  const float lse_reg = lane < H ? __ldg(a.lse + (int64_t)row * H + lane) : 0.f;
  df_reg = 0.f;

  for (int64_t base = beg; base < end; base += 32) {
    const int cnt = (end - base) < 32 ? (int)(end - base) : 32;
    const bool valid = lane < cnt;
    const int64_t e = base + lane;
    const int j = valid ? __ldg(a.col + e) : 0;
    col_s[lane] = j;
    const float* gj = a.g + (int64_t)j * H;
    const uint8_t* kp = a.keep ? a.keep + e * H : nullptr;
    for (int h = 0; h < H; ++h) {
      float fi = __shfl_sync(FULL, f_reg, h);
      float ls = __shfl_sync(FULL, lse_reg, h);
      float ch = __shfl_sync(FULL, c_reg, h);
      float z = fi + (valid ? __ldg(gj + h) : 0.f);
      float s = z > 0.f ? z : a.alpha * z;
      float al = valid ? expf(s - ls) : 0.f;
      float slope = z > 0.f ? 1.f : a.alpha;
      float kv = kp ? ((valid && kp[h]) ? a.inv_keep : 0.f) : 1.f;
      A_s[lane * HP + h] = al * slope * kv;
      B_s[lane * HP + h] = al * slope * ch;
      at_s[lane * HP + h] = al * kv;
    }
    __syncwarp();
    const float* whl = a.wh + lane * 4;
    int t = 0;
    for (; t + U <= cnt; t += U) {
      float pr[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float* wj = whl + (int64_t)col_s[t + u] * a.ldw;
#pragma unroll
        for (int v = 0; v < NV; ++v) pr[u][v] = geo.act[v] ? dot4(dh[v], ldg4(wj + v * 128)) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        head_reduce<NV>(pr[u], a.lph);
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (geo.leader[v]) {
            const int k = (t + u) * HP + geo.hv[v];
            dz_s[k] = fmaf(A_s[k], pr[u][v], -B_s[k]);
          }
      }
    }
    for (; t < cnt; ++t) {
      float pr[NV];
      const float* wj = whl + (int64_t)col_s[t] * a.ldw;
#pragma unroll
      for (int v = 0; v < NV; ++v) pr[v] = geo.act[v] ? dot4(dh[v], ldg4(wj + v * 128)) : 0.f;
      head_reduce<NV>(pr, a.lph);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (geo.leader[v]) {
          const int k = t * HP + geo.hv[v];
          dz_s[k] = fmaf(A_s[k], pr[v], -B_s[k]);
        }
    }
    __syncwarp();
    // coalesced spill of the chunk's per-edge scalars, and df accumulation
    for (int idx = lane; idx < cnt * H; idx += 32) {
      const int tt = idx / H, h = idx - tt * H;
      a.edge_dz[base * H + idx] = dz_s[tt * HP + h];
      a.edge_alpha[base * H + idx] = at_s[tt * HP + h];
    }
    for (int h = 0; h < H; ++h) {
      float s = warp_sum(valid ? dz_s[lane * HP + h] : 0.f);
      if (lane == h) df_reg += s;
    }
    __syncwarp();
  }
}

template <int NV, bool HUB>
__global__ void __launch_bounds__(BWD_WARPS * 32) attn_bwd_dst_kernel(const BwdDstArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = 32 + 4 * 32 * a.HP + 32;
  float* base_s = smem + warp * per_warp;
  int* col_s = reinterpret_cast<int*>(base_s);
  float* A_s = base_s + 32;
  float* B_s = A_s + 32 * a.HP;
  float* at_s = B_s + 32 * a.HP;
  float* dz_s = at_s + 32 * a.HP;
  float* c_s = dz_s + 32 * a.HP;
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  float df_reg;

  if (HUB) {
    int seg = blockIdx.x * BWD_WARPS + warp;
    if (seg >= a.n_hub_seg) return;
    int row;
    int64_t beg, end;
    hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.rowptr, a.seg_len, row, beg, end);
    bwd_dst_segment<NV>(a, row, beg, end, beg == a.rowptr[row], lane, geo, df_reg, col_s, A_s, B_s, at_s, dz_s, c_s);
    if (lane < a.H) a.scratch[(int64_t)seg * a.H + lane] = df_reg;
    return;
  }

  int cur = warp_grab(a.counter, lane);
  while (cur < a.n_dst) {
    int nxt = warp_grab(a.counter, lane);
    int rend = cur + GRAB < a.n_dst ? cur + GRAB : (int)a.n_dst;
    for (int row = cur; row < rend; ++row) {
      int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
      if (end - beg > a.seg_len) continue;
      bwd_dst_segment<NV>(a, row, beg, end, true, lane, geo, df_reg, col_s, A_s, B_s, at_s, dz_s, c_s);
      if (lane < a.H) a.df[(int64_t)row * a.H + lane] = df_reg;
    }
    cur = nxt;
  }
}

__global__ void attn_bwd_dst_hub_merge_kernel(const BwdDstArgs a) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.n_hub * a.H) return;
  int hub = idx / a.H, h = idx - hub * a.H;
  float s = 0.f;
  for (int k = a.hub_seg_ptr[hub]; k < a.hub_seg_ptr[hub + 1]; ++k) s += a.scratch[(int64_t)k * a.H + h];
  a.df[(int64_t)a.hub_rows[hub] * a.H + h] = s;
}

// =====================================================================================
// K4 backward, source pass over the transposed pattern
// =====================================================================================
struct BwdSrcArgs {
  int64_t n_src;
  const int64_t* tptr;
  const int32_t* trow;
  const int32_t* perm;
  int H, Dp, lph, V, HP;
  const float* dhp;
  int64_t lddhp;
  const float* edge_alpha;
  const float* edge_dz;
  const float* df;
  const float* a_src;
  const float* a_dst;
  const uint8_t* keep_wh;
  float inv_keep;
  float* dwh;
  int64_t lddwh;
  float* dg;
  int seg_len;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int n_hub, n_hub_seg;
  float* scratch;
  int32_t* counter;
};

template <int NV>
__device__ __forceinline__ void bwd_src_segment(const BwdSrcArgs& a, int64_t beg, int64_t end, int lane,
                                                const LaneGeom<NV>& geo, float4 (&acc)[NV], float& dg_reg,
                                                int* row_s, float* p_s) {
  constexpr int U = NV >= 8 ? 1 : 8 / NV;
  const int H = a.H, HP = a.HP;
  dg_reg = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = beg; base < end; base += 32) {
    const int cnt = (end - base) < 32 ? (int)(end - base) : 32;
    const bool valid = lane < cnt;
    const int64_t e = base + lane;
    const int i = valid ? __ldg(a.trow + e) : 0;
    const int64_t pe = valid ? __ldg(a.perm + e) : 0;
    row_s[lane] = i;
    for (int h = 0; h < H; ++h) {
      float at = valid ? __ldg(a.edge_alpha + pe * H + h) : 0.f;
      float dz = valid ? __ldg(a.edge_dz + pe * H + h) : 0.f;
      p_s[lane * HP + h] = at;
      float s = warp_sum(dz);
      if (lane == h) dg_reg += s;
    }
    __syncwarp();
    const float* dl = a.dhp + lane * 4;
    int t = 0;
    for (; t + U <= cnt; t += U) {
      float4 w[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float* di = dl + (int64_t)row_s[t + u] * a.lddhp;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (geo.act[v]) w[u][v] = ldg4(di + v * 128);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (geo.act[v]) fma4(acc[v], p_s[(t + u) * HP + geo.hv[v]], w[u][v]);
      }
    }
    for (; t < cnt; ++t) {
      const float* di = dl + (int64_t)row_s[t] * a.lddhp;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (geo.act[v]) fma4(acc[v], p_s[t * HP + geo.hv[v]], ldg4(di + v * 128));
    }
    __syncwarp();
  }
}

// dWh_j = acc + df_j a_src + dg_j a_dst, then the post-projection dropout mask.
__device__ __forceinline__ void bwd_src_store_slot(const BwdSrcArgs& a, int j, int slot, int h, float4 r, float dgv) {
  if (a.df) {
    float dfv = __ldg(a.df + (int64_t)j * a.H + h);
    fma4(r, dfv, ldg4(a.a_src + slot * 4));
    fma4(r, dgv, ldg4(a.a_dst + slot * 4));
  }
  if (a.keep_wh) {
    uchar4 k = *reinterpret_cast<const uchar4*>(a.keep_wh + (int64_t)j * (a.V * 4) + slot * 4);
    r.x = k.x ? r.x * a.inv_keep : 0.f;
    r.y = k.y ? r.y * a.inv_keep : 0.f;
    r.z = k.z ? r.z * a.inv_keep : 0.f;
    r.w = k.w ? r.w * a.inv_keep : 0.f;
  }
  stg4(a.dwh + (int64_t)j * a.lddwh + slot * 4, r);
}

template <int NV, bool HUB>
__global__ void __launch_bounds__(FWD_WARPS * 32) attn_bwd_src_kernel(const BwdSrcArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = 32 + 32 * a.HP;
  float* base_s = smem + warp * per_warp;
  int* row_s = reinterpret_cast<int*>(base_s);
  float* p_s = base_s + 32;
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  float4 acc[NV];
  float dg_reg;

  if (HUB) {
    int seg = blockIdx.x * FWD_WARPS + warp;
    if (seg >= a.n_hub_seg) return;
    int j;
    int64_t beg, end;
    hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.tptr, a.seg_len, j, beg, end);
    bwd_src_segment<NV>(a, beg, end, lane, geo, acc, dg_reg, row_s, p_s);
    float* sc = a.scratch + (int64_t)seg * src_scratch_stride(a.H, a.V);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (geo.act[v]) stg4(sc + (lane + 32 * v) * 4, acc[v]);
    if (lane < a.H) sc[a.V * 4 + lane] = dg_reg;
    return;
  }

  int cur = warp_grab(a.counter, lane);
  while (cur < a.n_src) {
    int nxt = warp_grab(a.counter, lane);
    int rend = cur + GRAB < a.n_src ? cur + GRAB : (int)a.n_src;
    for (int j = cur; j < rend; ++j) {
      int64_t beg = a.tptr[j], end = a.tptr[j + 1];
      if (end - beg > a.seg_len) continue;
      bwd_src_segment<NV>(a, beg, end, lane, geo, acc, dg_reg, row_s, p_s);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float dgv = __shfl_sync(FULL, dg_reg, geo.hv[v]);
        if (geo.act[v]) bwd_src_store_slot(a, j, lane + 32 * v, geo.hv[v], acc[v], dgv);
      }
      if (a.dg && lane < a.H) a.dg[(int64_t)j * a.H + lane] = dg_reg;
    }
    cur = nxt;
  }
}

__global__ void attn_bwd_src_hub_merge_kernel(const BwdSrcArgs a) {
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
      float4 t = *reinterpret_cast<const float4*>(sc + slot * 4);
      A.x += t.x; A.y += t.y; A.z += t.z; A.w += t.w;
    }
    bwd_src_store_slot(a, j, slot, h, A, dgv);
    if (a.dg && slot % a.lph == 0) a.dg[(int64_t)j * a.H + h] = dgv;
  }
}

// =====================================================================================
// host side
// =====================================================================================
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

#define NV_DISPATCH(nv, CALL)                     \
  switch (nv) {                                   \
    case 1: { constexpr int NV = 1; CALL; } break;   \
    case 2: { constexpr int NV = 2; CALL; } break;   \
    case 4: { constexpr int NV = 4; CALL; } break;   \
    case 8: { constexpr int NV = 8; CALL; } break;   \
    default: { constexpr int NV = 16; CALL; } break; \
  }

template <int NV>
static int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)FWD_WARPS * (32 + 32 * a.HP + 32) * sizeof(float);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<NV, true><<<(a.n_hub_seg + FWD_WARPS - 1) / FWD_WARPS, FWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
    attn_fwd_hub_merge_kernel<<<a.n_hub, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_dst > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_fwd_kernel<NV, false>, FWD_WARPS * 32, smem, &grid)) return rc;
    int64_t need = (a.n_dst + (int64_t)FWD_WARPS * GRAB - 1) / ((int64_t)FWD_WARPS * GRAB);
    if (need < grid) grid = (int)need;
    attn_fwd_kernel<NV, false><<<grid, FWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int NV>
static int launch_bwd_dst(const BwdDstArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)BWD_WARPS * (32 + 4 * 32 * a.HP + 32) * sizeof(float);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dst_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_dst_kernel<NV, true><<<(a.n_hub_seg + BWD_WARPS - 1) / BWD_WARPS, BWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
    attn_bwd_dst_hub_merge_kernel<<<(a.n_hub * a.H + 127) / 128, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_dst > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_bwd_dst_kernel<NV, false>, BWD_WARPS * 32, smem, &grid)) return rc;
    int64_t need = (a.n_dst + (int64_t)BWD_WARPS * GRAB - 1) / ((int64_t)BWD_WARPS * GRAB);
    if (need < grid) grid = (int)need;
    attn_bwd_dst_kernel<NV, false><<<grid, BWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int NV>
static int launch_bwd_src(const BwdSrcArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)FWD_WARPS * (32 + 32 * a.HP) * sizeof(float);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_src_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_src_kernel<NV, true><<<(a.n_hub_seg + FWD_WARPS - 1) / FWD_WARPS, FWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
    attn_bwd_src_hub_merge_kernel<<<a.n_hub, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_src > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_bwd_src_kernel<NV, false>, FWD_WARPS * 32, smem, &grid)) return rc;
    int64_t need = (a.n_src + (int64_t)FWD_WARPS * GRAB - 1) / ((int64_t)FWD_WARPS * GRAB);
    if (need < grid) grid = (int)need;
    attn_bwd_src_kernel<NV, false><<<grid, FWD_WARPS * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

static int check_hub(int seg_len, int n_hub, int n_hub_seg, const void* rows, const void* ptr, const void* scratch) {
  GATK_REQUIRE(seg_len >= 1, "seg_len must be >= 1");
  GATK_REQUIRE(n_hub >= 0 && n_hub_seg >= 0, "negative hub counts");
  if (n_hub_seg > 0) GATK_REQUIRE(rows && ptr && scratch && n_hub > 0, "hub arrays missing");
  return 0;
}

}  // namespace gatk

using namespace gatk;

extern "C" size_t gatk_hub_scratch_floats(int which, int H, int Dp, int n_hub_seg) {
  const int V = H * (Dp / 4);
  if (n_hub_seg <= 0) return 0;
  if (which == 0) return (size_t)n_hub_seg * fwd_scratch_stride(H, V);
  if (which == 1) return (size_t)n_hub_seg * H;
  return (size_t)n_hub_seg * src_scratch_stride(H, V);
}

extern "C" int gatk_attn_fwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp,
                             const float* wh, int64_t ldw, const float* f, const float* g,
                             const uint8_t* keep_att, float inv_keep, float alpha,
                             const float* skipv, int64_t lds, int act_elu,
                             float* hagg, float* out, int64_t ldo, float* lse,
                             int seg_len, const int32_t* hub_rows, const int32_t* hub_seg_ptr,
                             int n_hub, int n_hub_seg, float* hub_scratch, int32_t* counter, void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_dst < (1LL << 31), "n_dst too large for one shard");
  GATK_REQUIRE(ldw % 4 == 0 && ldo % 4 == 0 && (!skipv || lds % 4 == 0), "leading dims must be multiples of 4 floats");
  GATK_REQUIRE(rowptr && col && wh && f && g && out && counter, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  FwdArgs a;
  a.n_dst = n_dst; a.rowptr = rowptr; a.col = col; a.H = H; a.Dp = Dp; a.lph = Dp / 4; a.V = H * (Dp / 4);
  a.HP = H | 1;
  a.wh = wh; a.ldw = ldw; a.f = f; a.g = g; a.keep = keep_att; a.inv_keep = inv_keep; a.alpha = alpha;
  a.skipv = skipv; a.lds = lds; a.act_elu = act_elu; a.hagg = hagg; a.out = out; a.ldo = ldo; a.lse = lse;
  a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg;
  a.scratch = hub_scratch; a.counter = counter;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  NV_DISPATCH(nv, return launch_fwd<NV>(a, st));
  return 0;
}

extern "C" int gatk_attn_bwd_dst(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp,
                                 const float* wh, int64_t ldw, const float* f, const float* g, const float* lse,
                                 const uint8_t* keep_att, float inv_keep, float alpha,
                                 const float* gout, int64_t ldgo, const float* out, int64_t ldo, int act_elu,
                                 const float* hagg, int64_t ldh,
                                 float* dhp, int64_t lddhp, float* df, float* edge_alpha, float* edge_dz,
                                 int seg_len, const int32_t* hub_rows, const int32_t* hub_seg_ptr,
                                 int n_hub, int n_hub_seg, float* hub_scratch, int32_t* counter, void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_dst < (1LL << 31), "n_dst too large for one shard");
  GATK_REQUIRE(ldw % 4 == 0 && ldgo % 4 == 0 && ldh % 4 == 0 && lddhp % 4 == 0 && (!act_elu || ldo % 4 == 0),
               "leading dims must be multiples of 4 floats");
  GATK_REQUIRE(rowptr && col && wh && f && g && lse && gout && hagg && dhp && df && edge_alpha && edge_dz && counter,
               "null pointer argument");
  GATK_REQUIRE(!act_elu || out, "out is required when act_elu is set");
  cudaStream_t st = (cudaStream_t)stream;
  BwdDstArgs a;
  a.n_dst = n_dst; a.rowptr = rowptr; a.col = col; a.H = H; a.Dp = Dp; a.lph = Dp / 4; a.V = H * (Dp / 4);
  a.HP = H | 1;
  a.wh = wh; a.ldw = ldw; a.f = f; a.g = g; a.lse = lse; a.keep = keep_att; a.inv_keep = inv_keep; a.alpha = alpha;
  a.gout = gout; a.ldgo = ldgo; a.out = out; a.ldo = ldo; a.act_elu = act_elu; a.hagg = hagg; a.ldh = ldh;
  a.dhp = dhp; a.lddhp = lddhp; a.df = df; a.edge_alpha = edge_alpha; a.edge_dz = edge_dz;
  a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg;
  a.scratch = hub_scratch; a.counter = counter;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  NV_DISPATCH(nv, return launch_bwd_dst<NV>(a, st));
  return 0;
}

extern "C" int gatk_attn_bwd_src(int64_t n_src, const int64_t* tptr, const int32_t* trow, const int32_t* perm,
                                 int H, int Dp, const float* dhp, int64_t lddhp,
                                 const float* edge_alpha, const float* edge_dz,
                                 const float* df, const float* a_src, const float* a_dst,
                                 const uint8_t* keep_wh, float inv_keep,
                                 float* dwh, int64_t lddwh, float* dg,
                                 int seg_len, const int32_t* hub_rows, const int32_t* hub_seg_ptr,
                                 int n_hub, int n_hub_seg, float* hub_scratch, int32_t* counter, void* stream) {
  int nv;
  if (int rc = check_geom(H, Dp, &nv)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_src < (1LL << 31), "n_src too large for one shard");
  GATK_REQUIRE(lddhp % 4 == 0 && lddwh % 4 == 0, "leading dims must be multiples of 4 floats");
  GATK_REQUIRE(tptr && trow && perm && dhp && edge_alpha && edge_dz && dwh && counter, "null pointer argument");
  GATK_REQUIRE(!df || (a_src && a_dst), "a_src / a_dst required with df");
  cudaStream_t st = (cudaStream_t)stream;
  BwdSrcArgs a;
  a.n_src = n_src; a.tptr = tptr; a.trow = trow; a.perm = perm; a.H = H; a.Dp = Dp; a.lph = Dp / 4;
  a.V = H * (Dp / 4); a.HP = H | 1;
  a.dhp = dhp; a.lddhp = lddhp; a.edge_alpha = edge_alpha; a.edge_dz = edge_dz; a.df = df; a.a_src = a_src;
  a.a_dst = a_dst; a.keep_wh = keep_wh; a.inv_keep = inv_keep; a.dwh = dwh; a.lddwh = lddwh; a.dg = dg;
  a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg;
  a.scratch = hub_scratch; a.counter = counter;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  NV_DISPATCH(nv, return launch_bwd_src<NV>(a, st));
  return 0;
}

// "Aggregate-first" form of the GAT layer for NARROW inputs (F_in < H*D, no dropout between the
// projection and the logits -- train_ppi.py:49, eval mode, the products benchmark shape).
//
// The reference computes h'_i = sum_j alpha_ij (x_j W)  (layers.py:134,156-160).  The sum over
// neighbours and the projection commute, so per head h
//
//     h'_ih = (sum_j alpha_ijh x_j) W_h  =  xagg_ih W_h,
//
// and the logits are linear in the input, f = x (W a_src), g = x (W a_dst) (layers.py:141-144).
// The edge pass therefore gathers the F_in-wide INPUT row x_j (400 B at F_in = 100) instead of the
// H*D-wide projected row (2 KiB at 8 x 64), keeps H accumulators of F_in floats per destination row,
// and the projection runs afterwards on the aggregated rows (same FLOPs as projecting x).
//
// Backward (dx not needed -- a first layer): with dxagg_ih = dh'_ih W_h^T,
//     dalpha_ijh = dxagg_ih . x_j,   c_ih = dxagg_ih . xagg_ih  (softmax row term),
//     ds_ijh = alpha_ijh (dalpha_ijh - c_ih) LeakyReLU'(f_ih + g_jh),
//     df_ih = sum_j ds_ijh,  dg_jh = sum_i ds_ijh  (transposed sum over the edge scalars only),
// so the backward edge pass is destination-major as well and gathers x_j once more; nothing
// H*D wide is ever gathered.
//
// Lane geometry: an x row is S = Fp/4 float4 slots, lane l owns slots l, l+32, ... (NS per lane).
// Softmax terms are computed with lanes = edges (forward) or lanes = (edge, head) pairs (backward,
// after a transposing butterfly reduction of the 32 partial dot products of an edge group).
#include "attn_common.cuh"

namespace gatk {

constexpr int XW = 4;  // warps per CTA (small CTAs: the staging buffers set how many fit on an SM)

struct XArgs {
  int64_t n_dst;
  const int64_t* rowptr;
  const int32_t* col;
  int H, S, Fp;
  const float* x;
  int64_t ldx;
  const float* f;
  const float* g;
  int64_t ldfg;
  float alpha;
  float* xagg;
  int64_t ldxa;
  float* lse;
  const float* dxagg;
  int64_t ldd;
  float* ds;
  float* df;
  int64_t lddf;
  int seg_len;
  const int32_t* hub_rows;
  const int32_t* hub_seg_ptr;
  int n_hub, n_hub_seg;
  float* scratch;
  int32_t* counter;
  const int32_t* item_ptr;
  int n_items;
};

__host__ __device__ __forceinline__ int64_t xfwd_scratch_stride(int H, int Fp) { return (int64_t)H * Fp + ((2 * H + 3) & ~3); }

template <int HP>
struct Log2;
template <> struct Log2<1> { static constexpr int v = 0; };
template <> struct Log2<2> { static constexpr int v = 1; };
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };

// Butterfly "reduce-scatter" over the warp: on entry every lane holds NVAL partial values; on exit
// v[0] of lane l holds (NVAL == 32) the warp-wide reduction of value l, or (NVAL < 32) that of value
// (l >> (5 - log2 NVAL)) -- each step halves the values a lane carries (NVAL-1 + 5-log2 NVAL shuffles
// instead of 5*NVAL).
template <int NVAL, bool MAX = false>
__device__ __forceinline__ void butterfly_scatter(float (&v)[NVAL], int lane) {
  int o = 16;
#pragma unroll
  for (int n = NVAL / 2; n >= 1; n >>= 1, o >>= 1) {
    const bool up = lane & o;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      const float r = __shfl_xor_sync(FULL, send, o);
      v[i] = MAX ? fmaxf(keep, r) : keep + r;
    }
  }
  for (; o >= 1; o >>= 1) {
    const float r = __shfl_xor_sync(FULL, v[0], o);
    v[0] = MAX ? fmaxf(v[0], r) : v[0] + r;
  }
}

// ---- per-warp staging in shared memory -------------------------------------------------------------
// A warp works on CHUNKS of up to `chunk` stored entries of one destination row.  The neighbour rows
// x_j of a chunk are copied to shared memory with cp.async (16 bytes per lane, one instruction per
// row), so the number of row gathers in flight is set by the staging buffer (chunk rows per warp), not
// by registers; the per-edge scalars (forward: softmax weights, backward: g_j) and the next chunk's
// column ids are fetched while the copies are in flight.
__host__ __device__ __forceinline__ int x_chunk_for(int Fp) {
  int c = 32;
  while (c > 1 && c * Fp * 4 > 16384) c >>= 1;
  return c;
}
__host__ __device__ __forceinline__ int x_warp_smem_floats(int Fp, int HP, int chunk) { return chunk * Fp + 32 * HP + 32 + 8; }

struct WarpStage {
  float* rows;   // [chunk][Fp]
  float* es;     // [32][HP] per-edge scalars
  int* cols;     // [32]
  float* scale;  // [8]
  __device__ __forceinline__ void init(float* base, int warp, int Fp, int HP, int chunk) {
    float* p = base + (size_t)warp * x_warp_smem_floats(Fp, HP, chunk);
    rows = p;
    es = p + chunk * Fp;
    cols = reinterpret_cast<int*>(es + 32 * HP);
    scale = es + 32 * HP + 32;
  }
};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct Chunk {
  int row, cnt;
  int64_t base;
  bool first, last, ok;
};

// Enumerates the chunks a warp processes, one chunk AHEAD of the compute (warp-uniform state).  HUB: the
// chunks of one hub segment; otherwise the rows of dynamically claimed work items (hub rows skipped, empty
// rows yielded as a chunk of 0 entries so their outputs get written).
template <bool HUB>
struct ChunkIter {
  int cur, nxt, n_work, step, chunk, row, rend;
  int64_t beg, base, end;
  bool open;
  __device__ __forceinline__ void init(const XArgs& a, int lane, int seg, int chunk_) {
    chunk = chunk_;
    if (HUB) {
      int r;
      hub_locate(seg, a.hub_rows, a.hub_seg_ptr, a.n_hub, a.rowptr, a.seg_len, r, beg, end);
      row = r;
      rend = r;
      base = beg;
      open = true;
      cur = nxt = n_work = step = 0;
    } else {
      n_work = a.item_ptr ? a.n_items : (int)a.n_dst;
      step = a.item_ptr ? 1 : GRAB;
      cur = warp_grab(a.counter, lane, step);
      nxt = warp_grab(a.counter, lane, step);
      row = rend = 0;
      open = false;
      beg = base = end = 0;
    }
  }
  __device__ __forceinline__ Chunk next(const XArgs& a, int lane) {
    Chunk c;
    c.ok = false;
    c.row = 0; c.cnt = 0; c.base = 0; c.first = c.last = false;
    while (true) {
      if (open) {
        const int64_t rem = end - base;
        c.row = row;
        c.base = base;
        c.cnt = rem < chunk ? (int)rem : chunk;
        c.first = base == beg;
        base += chunk;
        c.last = base >= end;
        c.ok = true;
        if (c.last) {
          open = false;
          ++row;
        }
        return c;
      }
      if (HUB) return c;
      if (row >= rend) {
        if (cur >= n_work) return c;
        if (a.item_ptr) {
          row = a.item_ptr[cur];
          rend = a.item_ptr[cur + 1];
        } else {
          row = cur;
          rend = cur + GRAB < a.n_dst ? cur + GRAB : (int)a.n_dst;
        }
        cur = nxt;
        nxt = warp_grab(a.counter, lane, step);
        continue;
      }
      beg = a.rowptr[row];
      end = a.rowptr[row + 1];
      if (end - beg > a.seg_len) {  // hub: handled by the segment kernels
        ++row;
        continue;
      }
      base = beg;
      open = true;
    }
  }
};

// cp.async the cnt neighbour rows of the staged column ids into the row buffer (lanes = float4 slots).
template <int NS>
__device__ __forceinline__ void x_issue_rows(const XArgs& a, const WarpStage& w, int cnt, const int (&loff)[NS],
                                             const bool (&act)[NS]) {
#pragma unroll 4
  for (int t = 0; t < cnt; ++t) {
    const float* xj = a.x + (int64_t)w.cols[t] * a.ldx;
    float* dst = w.rows + t * a.Fp;
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (act[s]) cp_async16(dst + loff[s], xj + loff[s]);
  }
  cp_async_commit();
}

// =====================================================================================================
// forward: xagg_ih = sum_j softmax_j(LeakyReLU(f_ih + g_jh)) x_j,  lse_ih
// Every lane carries the running (max, sum) of head  lane >> (5 - log2 HP).
// =====================================================================================================
template <int HP, int NS, bool HUB>
__global__ void __launch_bounds__(XW * 32) attn_x_fwd_kernel(const XArgs a, const int chunk) {
  extern __shared__ __align__(16) float x_smem[];
  constexpr int SH = 5 - Log2<HP>::v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = a.H;
  const int hq = lane >> SH;
  const bool hq_writer = (lane & ((1 << SH) - 1)) == 0 && hq < H;
  WarpStage w;
  w.init(x_smem, warp, a.Fp, HP, chunk);
  int loff[NS];
  bool act[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    act[s] = lane + 32 * s < a.S;
    loff[s] = act[s] ? (lane + 32 * s) * 4 : 0;
  }
  const int seg = blockIdx.x * XW + warp;
  if (HUB && seg >= a.n_hub_seg) return;
  ChunkIter<HUB> it;
  it.init(a, lane, seg, chunk);

  float4 acc[HP][NS];
  float fv[HP];
  float m_reg = -INFINITY, l_reg = 0.f;
  Chunk c = it.next(a, lane);
  int j = (c.ok && lane < c.cnt) ? __ldg(a.col + c.base + lane) : 0;
  while (c.ok) {
    const Chunk n = it.next(a, lane);
    w.cols[lane] = j;
    __syncwarp();
    x_issue_rows<NS>(a, w, c.cnt, loff, act);
    const int jn = (n.ok && lane < n.cnt) ? __ldg(a.col + n.base + lane) : 0;  // next chunk's columns
    const bool valid = lane < c.cnt;
    // ---- softmax terms of this chunk (lanes = edges) while the row copies are in flight
    const float* gj = a.g + (int64_t)j * a.ldfg;
    float gv[HP];
#pragma unroll
    for (int h = 0; h < HP; ++h) gv[h] = (valid && h < H) ? __ldg(gj + h) : 0.f;
    if (c.first) {
#pragma unroll
      for (int h = 0; h < HP; ++h) fv[h] = h < H ? __ldg(a.f + (int64_t)c.row * a.ldfg + h) : 0.f;
      m_reg = -INFINITY;
      l_reg = 0.f;
#pragma unroll
      for (int h = 0; h < HP; ++h)
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[h][s] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float sv[HP], red[HP];
#pragma unroll
    for (int h = 0; h < HP; ++h) {
      const float z = fv[h] + gv[h];
      sv[h] = (valid && h < H) ? (z > 0.f ? z : a.alpha * z) : -INFINITY;
      red[h] = sv[h];
    }
    butterfly_scatter<HP, true>(red, lane);  // red[0]: chunk max of head hq
    const float m_new = fmaxf(m_reg, red[0]);
    const float sc = (m_reg == -INFINITY) ? 0.f : expf(m_reg - m_new);
    m_reg = m_new;
    float pe[HP];
#pragma unroll
    for (int h = 0; h < HP; ++h) {
      const float mh = __shfl_sync(FULL, m_new, h << SH);
      pe[h] = (valid && h < H) ? expf(sv[h] - mh) : 0.f;
      red[h] = pe[h];
    }
    butterfly_scatter<HP, false>(red, lane);  // red[0]: chunk sum of head hq
    l_reg = l_reg * sc + red[0];
    sts_vec<HP>(w.es + lane * HP, pe);
    if ((lane & ((1 << SH) - 1)) == 0) w.scale[hq] = sc;
    cp_async_wait_all();
    __syncwarp();
    if (!c.first) {
      float scv[HP];
      lds_vec<HP>(w.scale, scv);
#pragma unroll
      for (int h = 0; h < HP; ++h)
#pragma unroll
        for (int s = 0; s < NS; ++s) scale4(acc[h][s], scv[h]);
    }
    // ---- weighted sum of the staged rows (lanes = float4 slots)
#pragma unroll 4
    for (int t = 0; t < c.cnt; ++t) {
      float p[HP];
      lds_vec<HP>(w.es + t * HP, p);
      const float* xr = w.rows + t * a.Fp;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + loff[s]);
#pragma unroll
        for (int h = 0; h < HP; ++h) fma4(acc[h][s], p[h], xv);
      }
    }
    if (c.last) {
      if (HUB) {
        float* scr = a.scratch + (int64_t)seg * xfwd_scratch_stride(H, a.Fp);
#pragma unroll
        for (int h = 0; h < HP; ++h)
          if (h < H)
#pragma unroll
            for (int s = 0; s < NS; ++s)
              if (act[s]) stg4(scr + h * a.Fp + loff[s], acc[h][s]);
        if (hq_writer) {
          scr[H * a.Fp + hq] = m_reg;
          scr[H * a.Fp + H + hq] = l_reg;
        }
      } else {
        float* dst = a.xagg + (int64_t)c.row * a.ldxa;
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          if (h < H) {
            const float l = __shfl_sync(FULL, l_reg, h << SH);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              float4 r = acc[h][s];
              if (l > 0.f) {
                r.x /= l; r.y /= l; r.z /= l; r.w /= l;
              } else {
                r = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              if (act[s]) stg4(dst + h * a.Fp + loff[s], r);
            }
          }
        }
        if (a.lse && hq_writer) a.lse[(int64_t)c.row * H + hq] = l_reg > 0.f ? m_reg + logf(l_reg) : 0.f;
      }
    }
    __syncwarp();  // the staging buffers are rewritten by the next chunk
    c = n;
    j = jn;
  }
}

// One CTA per hub row: merge the segment states (m_k, l_k, acc_k).
__global__ void attn_x_fwd_hub_merge_kernel(const XArgs a) {
  const int hub = blockIdx.x;
  const int row = a.hub_rows[hub];
  const int s0 = a.hub_seg_ptr[hub], s1 = a.hub_seg_ptr[hub + 1];
  const int64_t stride = xfwd_scratch_stride(a.H, a.Fp);
  const int ml = a.H * a.Fp;
  for (int idx = threadIdx.x; idx < a.H * a.S; idx += blockDim.x) {
    const int h = idx / a.S, slot = idx - h * a.S;
    float M = -INFINITY;
    for (int s = s0; s < s1; ++s) M = fmaxf(M, a.scratch[s * stride + ml + h]);
    float L = 0.f;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = s0; s < s1; ++s) {
      const float* sc = a.scratch + s * stride;
      const float w = expf(sc[ml + h] - M);
      L = fmaf(sc[ml + a.H + h], w, L);
      fma4(A, w, *reinterpret_cast<const float4*>(sc + h * a.Fp + slot * 4));
    }
    if (L > 0.f) {
      A.x /= L; A.y /= L; A.z /= L; A.w /= L;
    } else {
      A = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    stg4(a.xagg + (int64_t)row * a.ldxa + h * a.Fp + slot * 4, A);
    if (a.lse && slot == 0) a.lse[(int64_t)row * a.H + h] = L > 0.f ? M + logf(L) : 0.f;
  }
}

// =====================================================================================================
// backward: ds_ijh, df_ih from (x, xagg, dxagg, f, g, lse)
// =====================================================================================================
template <int HP, int NS, bool HUB>
__global__ void __launch_bounds__(XW * 32) attn_x_bwd_kernel(const XArgs a, const int chunk) {
  extern __shared__ __align__(16) float x_smem[];
  constexpr int EPG = 32 / HP;  // edges per group: one (edge, head) pair per lane after the butterfly
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = a.H;
  const int my_h = lane & (HP - 1), my_e = lane / HP;
  const bool head_ok = my_h < H;
  WarpStage w;
  w.init(x_smem, warp, a.Fp, HP, chunk);
  int loff[NS];
  bool act[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    act[s] = lane + 32 * s < a.S;
    loff[s] = act[s] ? (lane + 32 * s) * 4 : 0;
  }
  const int seg = blockIdx.x * XW + warp;
  if (HUB && seg >= a.n_hub_seg) return;
  ChunkIter<HUB> it;
  it.init(a, lane, seg, chunk);

  float4 dxa[HP][NS];
  float c_my = 0.f, f_my = 0.f, lse_my = 0.f, df_acc = 0.f;
  Chunk c = it.next(a, lane);
  int j = (c.ok && lane < c.cnt) ? __ldg(a.col + c.base + lane) : 0;
  while (c.ok) {
    const Chunk n = it.next(a, lane);
    w.cols[lane] = j;
    __syncwarp();
    x_issue_rows<NS>(a, w, c.cnt, loff, act);
    const int jn = (n.ok && lane < n.cnt) ? __ldg(a.col + n.base + lane) : 0;
    // g_j of the chunk's edges (lanes = edges) -> staged for the (edge, head) lanes
    {
      const bool valid = lane < c.cnt;
      const float* gj = a.g + (int64_t)j * a.ldfg;
      float gv[HP];
#pragma unroll
      for (int h = 0; h < HP; ++h) gv[h] = (valid && h < H) ? __ldg(gj + h) : 0.f;
      if (c.first) {
        // row state: dxagg_i in registers, c_ih = dxagg_ih . xagg_ih, f_ih, lse_ih for this lane's head
        float cp[HP];
        const float* dxr = a.dxagg + (int64_t)c.row * a.ldd;
        const float* xar = a.xagg + (int64_t)c.row * a.ldxa;
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          cp[h] = 0.f;
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            if (h < H && act[s]) {
              dxa[h][s] = ldg4_stream(dxr + h * a.Fp + loff[s]);
              cp[h] += dot4(dxa[h][s], ldg4_stream(xar + h * a.Fp + loff[s]));
            } else {
              dxa[h][s] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
        f_my = head_ok ? __ldg(a.f + (int64_t)c.row * a.ldfg + my_h) : 0.f;
        lse_my = head_ok ? __ldg(a.lse + (int64_t)c.row * H + my_h) : 0.f;
        butterfly_scatter<HP>(cp, lane);
        c_my = __shfl_sync(FULL, cp[0], my_h << (5 - Log2<HP>::v));
        df_acc = 0.f;
      }
      sts_vec<HP>(w.es + lane * HP, gv);
    }
    cp_async_wait_all();
    __syncwarp();
    for (int t = 0; t < c.cnt; t += EPG) {
      float part[32];
#pragma unroll
      for (int u = 0; u < EPG; ++u) {
        const float* xr = w.rows + (t + u < c.cnt ? t + u : t) * a.Fp;
        float4 xv[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) xv[s] = *reinterpret_cast<const float4*>(xr + loff[s]);
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          float d = dot4(dxa[h][0], xv[0]);
#pragma unroll
          for (int s = 1; s < NS; ++s) d += dot4(dxa[h][s], xv[s]);
          part[u * HP + h] = d;
        }
      }
      butterfly_scatter<32>(part, lane);
      const int ee = t + my_e;
      const float z = f_my + w.es[(ee & 31) * HP + my_h];
      const float sl = z > 0.f ? z : a.alpha * z;
      const float al = expf(sl - lse_my);
      const float dsv = al * (part[0] - c_my) * (z > 0.f ? 1.f : a.alpha);
      if (ee < c.cnt && head_ok) {
        a.ds[(c.base + ee) * H + my_h] = dsv;
        df_acc += dsv;
      }
    }
    if (c.last) {
      float d = df_acc;
#pragma unroll
      for (int o = HP; o < 32; o <<= 1) d += __shfl_xor_sync(FULL, d, o);
      if (lane < H) {
        if (HUB) a.scratch[(int64_t)seg * H + lane] = d;
        else a.df[(int64_t)c.row * a.lddf + lane] = d;
      }
    }
    __syncwarp();
    c = n;
    j = jn;
  }
}

__global__ void attn_x_bwd_hub_merge_kernel(const XArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_hub * a.H) return;
  const int hub = i / a.H, h = i - hub * a.H;
  float s = 0.f;
  for (int k = a.hub_seg_ptr[hub]; k < a.hub_seg_ptr[hub + 1]; ++k) s += a.scratch[(int64_t)k * a.H + h];
  a.df[(int64_t)a.hub_rows[hub] * a.lddf + h] = s;
}

// =====================================================================================================
// dg_jh = sum over the stored entries of source j of ds: a segmented sum along the TRANSPOSED pattern,
// reading the edge scalars through perm.  Warp per source, lanes = (entry, head) pairs; sources with
// more than `long_len` entries get a whole CTA.
// =====================================================================================================
template <int HP>
__global__ void __launch_bounds__(256) edge_tsum_kernel(int64_t n_src, const int64_t* __restrict__ tptr,
                                                        const int32_t* __restrict__ perm, int H,
                                                        const float* __restrict__ ds, float* __restrict__ dg,
                                                        int64_t lddg, int long_len) {
  constexpr int EPG = 32 / HP;
  const int lane = threadIdx.x & 31;
  const int my_h = lane & (HP - 1), my_e = lane / HP;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n_src; j += warps) {
    const int64_t beg = tptr[j], end = tptr[j + 1];
    if (end - beg > long_len) continue;
    float acc = 0.f;
    if (my_h < H) {
      int64_t t = beg + my_e;
      for (; t + 3 * EPG < end; t += 4 * EPG) {
        const int e0 = __ldg(perm + t), e1 = __ldg(perm + t + EPG), e2 = __ldg(perm + t + 2 * EPG), e3 = __ldg(perm + t + 3 * EPG);
        const float v0 = __ldg(ds + (int64_t)e0 * H + my_h), v1 = __ldg(ds + (int64_t)e1 * H + my_h);
        const float v2 = __ldg(ds + (int64_t)e2 * H + my_h), v3 = __ldg(ds + (int64_t)e3 * H + my_h);
        acc += (v0 + v1) + (v2 + v3);
      }
      for (; t < end; t += EPG) acc += __ldg(ds + (int64_t)__ldg(perm + t) * H + my_h);
    }
#pragma unroll
    for (int o = HP; o < 32; o <<= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (lane < H) dg[j * lddg + lane] = acc;
  }
}

__global__ void __launch_bounds__(256) edge_tsum_long_kernel(const int32_t* __restrict__ rows, const int64_t* __restrict__ tptr,
                                                             const int32_t* __restrict__ perm, int H,
                                                             const float* __restrict__ ds, float* __restrict__ dg,
                                                             int64_t lddg) {
  __shared__ float red[256];
  const int j = rows[blockIdx.x];
  const int64_t beg = tptr[j], end = tptr[j + 1];
  for (int h = 0; h < H; ++h) {
    float acc = 0.f;
    for (int64_t t = beg + threadIdx.x; t < end; t += 256) acc += __ldg(ds + (int64_t)__ldg(perm + t) * H + h);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) dg[(int64_t)j * lddg + h] = red[0];
    __syncthreads();
  }
}

// =====================================================================================================
// ELU and its derivative on packed rows (float4 granularity)
// =====================================================================================================
__device__ __forceinline__ float elu1x(float v) { return v > 0.f ? v : expm1f(v); }
__device__ __forceinline__ float elud(float o) { return o > 0.f ? 1.f : o + 1.f; }  // ELU'(v) from out = ELU(v)

__global__ void elu_fwd_kernel(int64_t n, int c4, float* __restrict__ buf, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c4) return;
  float* p = buf + (i / c4) * ld + (i % c4) * 4;
  float4 v = *reinterpret_cast<float4*>(p);
  v.x = elu1x(v.x); v.y = elu1x(v.y); v.z = elu1x(v.z); v.w = elu1x(v.w);
  stg4(p, v);
}

__global__ void elu_bwd_kernel(int64_t n, int c4, const float* __restrict__ gout, int64_t ldg, const float* __restrict__ out,
                               int64_t ldo, float* __restrict__ dhp, int64_t ldd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c4) return;
  const int64_t r = i / c4, c = (i % c4) * 4;
  float4 gv = ldg4_stream(gout + r * ldg + c);
  const float4 o = ldg4_stream(out + r * ldo + c);
  gv.x *= elud(o.x); gv.y *= elud(o.y); gv.z *= elud(o.z); gv.w *= elud(o.w);
  stg4(dhp + r * ldd + c, gv);
}

static int check_x_geom(int H, int S, int* hp, int* ns) {
  GATK_REQUIRE(H >= 1 && H <= 8, "aggregate-first form: H=%d out of range [1,8]", H);
  GATK_REQUIRE(S >= 1 && S <= 128, "aggregate-first form: %d float4 slots per input row out of range [1,128]", S);
  *hp = H <= 1 ? 1 : (H <= 2 ? 2 : (H <= 4 ? 4 : 8));
  *ns = S <= 32 ? 1 : (S <= 64 ? 2 : 4);
  GATK_REQUIRE(*hp * *ns <= 16, "aggregate-first form: H=%d x %d floats per row needs too many accumulators", H, 4 * S);
  return 0;
}

#define X_DISPATCH(hp, ns, CALL)                                             \
  switch ((hp) * 8 + (ns)) {                                                 \
    case 1 * 8 + 1: { constexpr int HP = 1, NS = 1; CALL; } break;           \
    case 1 * 8 + 2: { constexpr int HP = 1, NS = 2; CALL; } break;           \
    case 1 * 8 + 4: { constexpr int HP = 1, NS = 4; CALL; } break;           \
    case 2 * 8 + 1: { constexpr int HP = 2, NS = 1; CALL; } break;           \
    case 2 * 8 + 2: { constexpr int HP = 2, NS = 2; CALL; } break;           \
    case 2 * 8 + 4: { constexpr int HP = 2, NS = 4; CALL; } break;           \
    case 4 * 8 + 1: { constexpr int HP = 4, NS = 1; CALL; } break;           \
    case 4 * 8 + 2: { constexpr int HP = 4, NS = 2; CALL; } break;           \
    case 4 * 8 + 4: { constexpr int HP = 4, NS = 4; CALL; } break;           \
    case 8 * 8 + 1: { constexpr int HP = 8, NS = 1; CALL; } break;           \
    default:        { constexpr int HP = 8, NS = 2; CALL; } break;           \
  }

template <int HP, int NS>
static int launch_x_fwd(const XArgs& a, cudaStream_t st) {
  const int chunk = x_chunk_for(a.Fp);
  const size_t smem = (size_t)XW * x_warp_smem_floats(a.Fp, HP, chunk) * sizeof(float);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_x_fwd_kernel<HP, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_x_fwd_kernel<HP, NS, true><<<(a.n_hub_seg + XW - 1) / XW, XW * 32, smem, st>>>(a, chunk);
    GATK_CHECK_LAUNCH();
    attn_x_fwd_hub_merge_kernel<<<a.n_hub, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_dst > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_x_fwd_kernel<HP, NS, false>, XW * 32, smem, &grid)) return rc;
    const int64_t need = a.item_ptr ? (a.n_items + XW - 1) / XW : (a.n_dst + (int64_t)XW * GRAB - 1) / ((int64_t)XW * GRAB);
    if (need < grid) grid = (int)need;
    attn_x_fwd_kernel<HP, NS, false><<<grid, XW * 32, smem, st>>>(a, chunk);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int HP, int NS>
static int launch_x_bwd(const XArgs& a, cudaStream_t st) {
  const int chunk = x_chunk_for(a.Fp);
  const size_t smem = (size_t)XW * x_warp_smem_floats(a.Fp, HP, chunk) * sizeof(float);
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(attn_x_bwd_kernel<HP, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_x_bwd_kernel<HP, NS, true><<<(a.n_hub_seg + XW - 1) / XW, XW * 32, smem, st>>>(a, chunk);
    GATK_CHECK_LAUNCH();
    attn_x_bwd_hub_merge_kernel<<<(a.n_hub * a.H + 127) / 128, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_dst > 0) {
    int grid = 0;
    if (int rc = persistent_grid(attn_x_bwd_kernel<HP, NS, false>, XW * 32, smem, &grid)) return rc;
    const int64_t need = a.item_ptr ? (a.n_items + XW - 1) / XW : (a.n_dst + (int64_t)XW * GRAB - 1) / ((int64_t)XW * GRAB);
    if (need < grid) grid = (int)need;
    attn_x_bwd_kernel<HP, NS, false><<<grid, XW * 32, smem, st>>>(a, chunk);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace gatk

using namespace gatk;

extern "C" size_t gatk_attn_x_scratch_floats(int which, int H, int Fp, int n_hub_seg) {
  if (n_hub_seg <= 0) return 0;
  if (which == 0) return (size_t)n_hub_seg * xfwd_scratch_stride(H, Fp);
  return (size_t)n_hub_seg * H;
}

extern "C" int gatk_attn_x_fwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Fp,
                               const float* x, int64_t ldx, const float* f, const float* g, int64_t ldfg, float alpha,
                               float* xagg, int64_t ldxa, float* lse, int seg_len, const int32_t* hub_rows,
                               const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                               int32_t* counter, const int32_t* item_ptr, int n_items, void* stream) {
  GATK_REQUIRE(Fp >= 4 && Fp % 4 == 0, "Fp=%d must be a positive multiple of 4", Fp);
  int hp, ns;
  if (int rc = check_x_geom(H, Fp / 4, &hp, &ns)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_dst < (1LL << 31), "n_dst too large for one shard");
  GATK_REQUIRE(rowptr && col && x && f && g && xagg && counter, "null pointer argument");
  GATK_REQUIRE(ldx % 4 == 0 && ldx >= Fp && ldxa % 4 == 0 && ldxa >= (int64_t)H * Fp && ldfg >= H,
               "leading dims: ldx, ldxa multiples of 4 floats, ldx >= Fp, ldxa >= H*Fp, ldfg >= H");
  GATK_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)xagg & 15) == 0, "x and xagg must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  XArgs a = {};
  a.n_dst = n_dst; a.rowptr = rowptr; a.col = col; a.H = H; a.S = Fp / 4; a.Fp = Fp; a.x = x; a.ldx = ldx;
  a.f = f; a.g = g; a.ldfg = ldfg; a.alpha = alpha; a.xagg = xagg; a.ldxa = ldxa; a.lse = lse;
  a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg;
  a.scratch = hub_scratch; a.counter = counter; a.item_ptr = item_ptr; a.n_items = n_items;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  X_DISPATCH(hp, ns, return (launch_x_fwd<HP, NS>(a, st)));
  return 0;
}

extern "C" int gatk_attn_x_bwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Fp,
                               const float* x, int64_t ldx, const float* f, const float* g, int64_t ldfg,
                               const float* lse, float alpha, const float* xagg, int64_t ldxa, const float* dxagg,
                               int64_t ldd, float* ds, float* df, int64_t lddf, int seg_len, const int32_t* hub_rows,
                               const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                               int32_t* counter, const int32_t* item_ptr, int n_items, void* stream) {
  GATK_REQUIRE(Fp >= 4 && Fp % 4 == 0, "Fp=%d must be a positive multiple of 4", Fp);
  int hp, ns;
  if (int rc = check_x_geom(H, Fp / 4, &hp, &ns)) return rc;
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_dst < (1LL << 31), "n_dst too large for one shard");
  GATK_REQUIRE(rowptr && col && x && f && g && lse && xagg && dxagg && ds && df && counter, "null pointer argument");
  GATK_REQUIRE(ldx % 4 == 0 && ldx >= Fp && ldxa % 4 == 0 && ldxa >= (int64_t)H * Fp && ldd % 4 == 0 &&
                   ldd >= (int64_t)H * Fp && ldfg >= H && lddf >= H,
               "leading dims: ldx, ldxa, ldd multiples of 4 floats and wide enough, ldfg, lddf >= H");
  GATK_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)xagg & 15) == 0 && ((uintptr_t)dxagg & 15) == 0,
               "x, xagg and dxagg must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  XArgs a = {};
  a.n_dst = n_dst; a.rowptr = rowptr; a.col = col; a.H = H; a.S = Fp / 4; a.Fp = Fp; a.x = x; a.ldx = ldx;
  a.f = f; a.g = g; a.ldfg = ldfg; a.alpha = alpha; a.xagg = const_cast<float*>(xagg); a.ldxa = ldxa;
  a.lse = const_cast<float*>(lse); a.dxagg = dxagg; a.ldd = ldd; a.ds = ds; a.df = df; a.lddf = lddf;
  a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg;
  a.scratch = hub_scratch; a.counter = counter; a.item_ptr = item_ptr; a.n_items = n_items;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  X_DISPATCH(hp, ns, return (launch_x_bwd<HP, NS>(a, st)));
  return 0;
}

extern "C" int gatk_edge_tsum(int64_t n_src, const int64_t* tptr, const int32_t* perm, int H, const float* ds,
                              float* dg, int64_t lddg, int long_len, const int32_t* long_rows, int n_long,
                              void* stream) {
  GATK_REQUIRE(H >= 1 && H <= 8, "H=%d out of range [1,8]", H);
  GATK_REQUIRE(tptr && dg && lddg >= H && long_len >= 1 && n_long >= 0 && (n_long == 0 || long_rows), "bad arguments");
  if (n_src == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int hp = H <= 1 ? 1 : (H <= 2 ? 2 : (H <= 4 ? 4 : 8));
  int64_t blocks = (n_src + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  switch (hp) {
    case 1: edge_tsum_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(n_src, tptr, perm, H, ds, dg, lddg, long_len); break;
    case 2: edge_tsum_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(n_src, tptr, perm, H, ds, dg, lddg, long_len); break;
    case 4: edge_tsum_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(n_src, tptr, perm, H, ds, dg, lddg, long_len); break;
    default: edge_tsum_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(n_src, tptr, perm, H, ds, dg, lddg, long_len); break;
  }
  GATK_CHECK_LAUNCH();
  if (n_long > 0) {
    edge_tsum_long_kernel<<<n_long, 256, 0, st>>>(long_rows, tptr, perm, H, ds, dg, lddg);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int gatk_elu_fwd(int64_t n, int64_t cols, float* buf, int64_t ld, void* stream) {
  GATK_REQUIRE(buf && cols % 4 == 0 && ld % 4 == 0 && ((uintptr_t)buf & 15) == 0, "bad arguments");
  if (n * cols == 0) return 0;
  const int64_t total = n * (cols / 4);
  elu_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, (int)(cols / 4), buf, ld);
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_elu_bwd(int64_t n, int64_t cols, const float* gout, int64_t ldg, const float* out, int64_t ldo,
                            float* dhp, int64_t ldd, void* stream) {
  GATK_REQUIRE(gout && out && dhp && cols % 4 == 0 && ldg % 4 == 0 && ldo % 4 == 0 && ldd % 4 == 0, "bad arguments");
  if (n * cols == 0) return 0;
  const int64_t total = n * (cols / 4);
  elu_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, (int)(cols / 4), gout, ldg, out, ldo, dhp, ldd);
  GATK_CHECK_LAUNCH();
  return 0;
}

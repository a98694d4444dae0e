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
// Data layout: the source-side logit half g_j rides BEHIND the input row, xg_j = [x_j (Fp floats, zero
// padded) | g_j (H floats) | pad] with 128-byte aligned rows (gatk_logits_pack builds it), so a stored
// entry costs ONE random DRAM access.  A warp works on chunks of up to 32 stored entries of one
// destination row: the chunk's rows are copied to shared memory with cp.async (one 16-byte-per-lane
// instruction per row), so the gathers in flight are bounded by the staging buffer, not by registers,
// and the column ids of the next chunk are fetched one chunk ahead.  Forward: softmax terms with
// lanes = edges, weighted sum with lanes = float4 slots (packed FFMA2).  Backward: lanes = edges, every
// lane dots ITS row against dxagg_i broadcast from shared memory -- no cross-lane reduction per edge.
#include <cuda.h>
#include <stdlib.h>

#include "attn_common.cuh"

namespace gatk {

constexpr int XW = 2;  // warps per CTA (small CTAs: the staging buffers set how many fit on an SM)

struct alignas(64) XArgs {
  CUtensorMap xmap;  // the gather rows as a 2-D tensor (box = RS floats x 1 row) for TMA tile::gather4
  int use_g4;        // rows are staged four per instruction through xmap (row pitch in shared memory = RS)
  int64_t n_src;
  int64_t n_dst;
  const int64_t* rowptr;
  const int32_t* col;
  int H, S, Fp, Sx, RS;  // heads; x slots; 4*S; slots copied per row (x + g); smem row pitch (floats)
  const float* xg;
  int64_t ldxg;
  const float* f;
  int64_t ldf;
  float alpha;
  float* xagg;
  int64_t ldxa;
  float* lse;
  const float* dxagg;
  int64_t ldd;
  float* ds;
  const int32_t* iperm;  // CSR entry -> position in the transposed pattern (NULL: ds stays in CSR order)
  float* dgacc;          // NULL, or [n_src, lddgacc] zero-initialised: dg_j += ds_ij accumulated here with red.global.add
  int64_t lddgacc;       //       (replaces the ds write + the transposed segmented sum gatk_edge_tsum)
  int dg_vec4;           // H == 8 and 16-byte aligned rows: two red.global.add.v4.f32 per stored entry
  int short_c;           // backward: rows of <= 2 chunks sum c_i over their own entries instead of staging xagg_i
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
template <> struct Log2<16> { static constexpr int v = 4; };

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
// row pitch of the staged rows: a multiple of 8 floats, so that groups of four rows (TMA tile::gather4
// destinations) start on 128-byte boundaries
__host__ __device__ __forceinline__ int x_row_pitch(int Sx) { return (4 * Sx + 7) / 8 * 8; }
__host__ __device__ __forceinline__ int x_chunk_for(int RS) {
  int c = 32;
  while (c > 4 && c * RS * 4 > 14336) c >>= 1;
  return c;
}
__host__ __device__ __forceinline__ int xfwd_warp_floats(int RS, int HP, int chunk) { return chunk * RS + 32 * HP + 32; }  // multiples of 128 bytes
// backward (tensor-core kernel): row pitch = 16 (mod 32) floats, so the two rows a quarter-warp touches per
// 128-bit fragment load sit in different bank halves
__host__ __device__ __forceinline__ int xmma_row_pitch(int Sx) { return (4 * Sx + 15) / 32 * 32 + 16; }
__host__ __device__ __forceinline__ int xmma_pend_floats(int chunk) { return (4 * (chunk / 16) * 2 + 3) * 32; }
__host__ __device__ __forceinline__ int xmma_warp_floats(int RS, int H, int Fp, int chunk) {
  // two row buffers | dxagg_i, xagg_i | barriers | held-back first chunk of a two-chunk row (w, dalpha, col, c sums per lane)
  return 2 * chunk * RS + (2 * H * Fp + 31) / 32 * 32 + 32 + xmma_pend_floats(chunk);  // every piece a multiple of 128 bytes
}
__host__ __device__ __forceinline__ int xbwd_warp_floats(int RS, int H, int Fp, int chunk) { return chunk * RS + H * Fp + 4; }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// acc += p * x on a float4 with two packed FFMA2 (pp = {p, p})
__device__ __forceinline__ void fma4_pp(float4& acc, const float2 pp, const float4& x) {
  const float2 lo = __ffma2_rn(pp, make_float2(x.x, x.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(pp, make_float2(x.z, x.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// dg accumulation: fire-and-forget vector reductions into the (L2-resident) [n_src, H] array
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// the same with an L2 evict_last policy: the dg array (N*H floats) is touched ~degree times per row at random,
// between gigabytes of streamed gather rows
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_add_v4_pol(float* p, float a, float b, float c, float d, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

struct Chunk {
  int row, cnt;
  int64_t base;
  bool first, last, ok;
  bool two;  // the whole destination row fits in two chunks (never set for hub segments)
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
    c.row = 0; c.cnt = 0; c.base = 0; c.first = c.last = c.two = false;
    while (true) {
      if (open) {
        const int64_t rem = end - base;
        c.two = !HUB && a.short_c && end - beg <= 2 * (int64_t)chunk;
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

// Stage the cnt neighbour rows [x_j | g_j] of a chunk: lane t issues ONE bulk copy (TMA, cp.async.bulk) of row
// cols[t] straight into the warp's row buffer; completion is tracked by the warp's mbarrier (expect_tx =
// cnt * row bytes).  One instruction per lane for the whole chunk -- the gather costs no issue slots and
// no registers, and up to 32 rows per warp are in flight.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
// Four rows per instruction (TMA tile::gather4): lane q stages entries 4q..4q+3; entries past cnt re-read row 0.
__device__ __forceinline__ void x_gather4(const XArgs& a, float* dst, int r0, int r1, int r2, int r3, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(&a.xmap), "r"(bar), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void x_issue_rows(const XArgs& a, float* rows, int j, int cnt, int lane, int RS, unsigned bar) {
  if (cnt > 0 && a.use_g4) {
    const int ngrp = (cnt + 3) >> 2;
    const int r0 = __shfl_sync(FULL, j, (4 * lane) & 31), r1 = __shfl_sync(FULL, j, (4 * lane + 1) & 31);
    const int r2 = __shfl_sync(FULL, j, (4 * lane + 2) & 31), r3 = __shfl_sync(FULL, j, (4 * lane + 3) & 31);
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(ngrp * 16 * RS)) : "memory");
    __syncwarp();
    if (lane < ngrp) x_gather4(a, rows + lane * 4 * RS, r0, r1, r2, r3, bar);
  } else if (cnt > 0) {
    const unsigned row_bytes = (unsigned)a.Sx * 16u;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (unsigned)cnt) : "memory");
    __syncwarp();
    if (lane < cnt) {
      const float* src = a.xg + (int64_t)j * a.ldxg;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(rows + lane * RS)),
                   "l"(src), "r"(row_bytes), "r"(bar)
                   : "memory");
    }
  }
}

// =====================================================================================================
// forward: xagg_ih = sum_j softmax_j(LeakyReLU(f_ih + g_jh)) x_j,  lse_ih
// Every lane carries the running (max, sum) of head  lane >> (5 - log2 HP).
// =====================================================================================================
template <int HP, int NS, bool HUB>
__global__ void __launch_bounds__(XW * 32) attn_x_fwd_kernel(const __grid_constant__ XArgs a, const int chunk) {
  extern __shared__ __align__(128) float x_smem[];
  constexpr int SH = 5 - Log2<HP>::v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = a.H;
  const int hq = lane >> SH;
  const bool hq_lead = (lane & ((1 << SH) - 1)) == 0;
  const bool hq_writer = hq_lead && hq < H;
  float* rows = x_smem + (size_t)warp * xfwd_warp_floats(a.RS, HP, chunk);
  float* es = rows + chunk * a.RS;  // [32][HP] softmax weights of the chunk's edges
  float* scale = es + 32 * HP;
  const unsigned bar = smem_u32(scale + 8);
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  unsigned phase = 0;
  int loff[NS];
  bool act[NS], act_x[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    act[s] = lane + 32 * s < a.Sx;     // copied / accumulated (x and g slots)
    act_x[s] = lane + 32 * s < a.S;    // stored (x slots)
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
    x_issue_rows(a, rows, j, c.cnt, lane, a.RS, bar);
    const int jn = (n.ok && lane < n.cnt) ? __ldg(a.col + n.base + lane) : 0;  // next chunk's columns
    const bool valid = lane < c.cnt;
    if (c.first) {
#pragma unroll
      for (int h = 0; h < HP; ++h) fv[h] = h < H ? __ldg(a.f + (int64_t)c.row * a.ldf + h) : 0.f;
      m_reg = -INFINITY;
      l_reg = 0.f;
#pragma unroll
      for (int h = 0; h < HP; ++h)
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[h][s] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (c.cnt > 0) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    // ---- softmax terms of this chunk (lanes = edges); g_j sits behind x_j in the staged row
    float gv[HP];
    lds_vec<HP>(rows + (valid ? lane : 0) * a.RS + a.Fp, gv);
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
    sts_vec<HP>(es + lane * HP, pe);
    if (hq_lead) scale[hq] = sc;
    __syncwarp();
    if (!c.first) {
      float scv[HP];
      lds_vec<HP>(scale, scv);
#pragma unroll
      for (int h = 0; h < HP; ++h)
#pragma unroll
        for (int s = 0; s < NS; ++s) scale4(acc[h][s], scv[h]);
    }
    // ---- weighted sum of the staged rows (lanes = float4 slots)
    {
      const float* ep = es;
      const float* xr = rows;
#pragma unroll 4
      for (int t = 0; t < c.cnt; ++t, ep += HP, xr += a.RS) {
        float p[HP];
        lds_vec<HP>(ep, p);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float4 xv = *reinterpret_cast<const float4*>(xr + loff[s]);
#pragma unroll
          for (int h = 0; h < HP; ++h) fma4_pp(acc[h][s], make_float2(p[h], p[h]), xv);
        }
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
              if (act_x[s]) stg4(scr + h * a.Fp + loff[s], acc[h][s]);
        if (hq_writer) {
          scr[H * a.Fp + hq] = m_reg;
          scr[H * a.Fp + H + hq] = l_reg;
        }
      } else {
        float* dst = a.xagg + (int64_t)c.row * a.ldxa;
        const float inv_mine = l_reg > 0.f ? 1.f / l_reg : 0.f;
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          if (h < H) {
            const float inv = __shfl_sync(FULL, inv_mine, h << SH);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              float4 r = acc[h][s];
              scale4(r, inv);
              if (act_x[s]) stg4(dst + h * a.Fp + loff[s], r);
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
// backward: ds_ijh, df_ih from (xg, xagg, dxagg, f, lse).  Lanes = edges: every lane dots its staged
// row against dxagg_i, which all lanes read (broadcast) from shared memory.
// =====================================================================================================
template <int HP, bool HUB>
__global__ void __launch_bounds__(XW * 32) attn_x_bwd_kernel(const __grid_constant__ XArgs a, const int chunk) {
  extern __shared__ __align__(128) float x_smem[];
  constexpr int SH = 5 - Log2<HP>::v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = a.H, S = a.S, Fp = a.Fp;
  const int hq = lane >> SH;
  const bool hq_writer = (lane & ((1 << SH) - 1)) == 0 && hq < H;
  float* rows = x_smem + (size_t)warp * xbwd_warp_floats(a.RS, H, Fp, chunk);
  float* dxs = rows + chunk * a.RS;  // [H][Fp] dxagg_i
  const unsigned bar = smem_u32(dxs + H * Fp);
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  unsigned phase = 0;
  const int seg = blockIdx.x * XW + warp;
  if (HUB && seg >= a.n_hub_seg) return;
  ChunkIter<HUB> it;
  it.init(a, lane, seg, chunk);

  float cv[HP], fv[HP], lv[HP];
  float df_acc = 0.f;  // running df of head hq
  Chunk c = it.next(a, lane);
  int j = (c.ok && lane < c.cnt) ? __ldg(a.col + c.base + lane) : 0;
  while (c.ok) {
    const Chunk n = it.next(a, lane);
    x_issue_rows(a, rows, j, c.cnt, lane, a.RS, bar);
    const int jn = (n.ok && lane < n.cnt) ? __ldg(a.col + n.base + lane) : 0;
    if (c.first) {
      // row state (lanes = slots): dxagg_i -> shared memory, c_ih = dxagg_ih . xagg_ih; f_ih, lse_ih
      float cp[HP];
#pragma unroll
      for (int h = 0; h < HP; ++h) cp[h] = 0.f;
      const float* dxr = a.dxagg + (int64_t)c.row * a.ldd;
      const float* xar = a.xagg + (int64_t)c.row * a.ldxa;
      for (int slot = lane; slot < S; slot += 32) {
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          if (h < H) {
            const float4 d = ldg4_stream(dxr + h * Fp + slot * 4);
            cp[h] += dot4(d, ldg4_stream(xar + h * Fp + slot * 4));
            *reinterpret_cast<float4*>(dxs + h * Fp + slot * 4) = d;
          }
        }
      }
#pragma unroll
      for (int h = 0; h < HP; ++h) {
        fv[h] = h < H ? __ldg(a.f + (int64_t)c.row * a.ldf + h) : 0.f;
        lv[h] = h < H ? __ldg(a.lse + (int64_t)c.row * H + h) : 0.f;
      }
      butterfly_scatter<HP>(cp, lane);
#pragma unroll
      for (int h = 0; h < HP; ++h) cv[h] = __shfl_sync(FULL, cp[0], h << SH);
      df_acc = 0.f;
    }
    if (c.cnt > 0) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    __syncwarp();
    // ---- dalpha_eh = dxagg_ih . x_e for this lane's edge e
    const bool valid = lane < c.cnt;
    const float* xr = rows + (valid ? lane : 0) * a.RS;
    float2 acc2[HP];
#pragma unroll
    for (int h = 0; h < HP; ++h) acc2[h] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int k = 0; k < S; ++k) {
      const float4 xq = *reinterpret_cast<const float4*>(xr + 4 * k);
#pragma unroll
      for (int h = 0; h < HP; ++h) {
        if (h < H) {
          const float4 dq = *reinterpret_cast<const float4*>(dxs + h * Fp + 4 * k);
          acc2[h] = __ffma2_rn(make_float2(dq.x, dq.y), make_float2(xq.x, xq.y), acc2[h]);
          acc2[h] = __ffma2_rn(make_float2(dq.z, dq.w), make_float2(xq.z, xq.w), acc2[h]);
        }
      }
    }
    float gv[HP], dsv[HP];
    lds_vec<HP>(xr + Fp, gv);
#pragma unroll
    for (int h = 0; h < HP; ++h) {
      const float z = fv[h] + gv[h];
      const float sl = z > 0.f ? z : a.alpha * z;
      const float al = expf(sl - lv[h]);
      const float v = al * ((acc2[h].x + acc2[h].y) - cv[h]) * (z > 0.f ? 1.f : a.alpha);
      dsv[h] = (valid && h < H) ? v : 0.f;
    }
    if (valid && a.ds) {
      float* dp = a.ds + (a.iperm ? (int64_t)__ldg(a.iperm + c.base + lane) : c.base + lane) * H;
      if (HP >= 4 && H == HP) {
#pragma unroll
        for (int h = 0; h < HP; h += 4) stg4(dp + h, make_float4(dsv[h], dsv[h + 1], dsv[h + 2], dsv[h + 3]));
      } else {
#pragma unroll
        for (int h = 0; h < HP; ++h)
          if (h < H) dp[h] = dsv[h];
      }
    }
    if (valid && a.dgacc) {
      float* gp = a.dgacc + (int64_t)j * a.lddgacc;
      if (HP >= 4 && a.dg_vec4) {
#pragma unroll
        for (int h = 0; h < HP; h += 4) red_add_v4(gp + h, dsv[h], dsv[h + 1], dsv[h + 2], dsv[h + 3]);
      } else {
#pragma unroll
        for (int h = 0; h < HP; ++h)
          if (h < H) red_add_f32(gp + h, dsv[h]);
      }
    }
    butterfly_scatter<HP>(dsv, lane);  // dsv[0]: this chunk's sum of head hq
    df_acc += dsv[0];
    if (c.last && hq_writer) {
      if (HUB) a.scratch[(int64_t)seg * H + hq] = df_acc;
      else a.df[(int64_t)c.row * a.lddf + hq] = df_acc;
    }
    __syncwarp();
    c = n;
    j = jn;
  }
}

// =====================================================================================================
// backward on the tensor cores (Fp <= 128).  The per-edge dots dalpha[e][h] = sum_k x_e[k] dxagg_i[h][k] of
// a chunk are one small GEMM  [32 edges x Fp] x [Fp x 8 heads]:  mma.sync m16n8k8 tf32 with the 3xTF32
// error-compensated split (x = hi + lo, hi = tf32 truncation; hi*hi + lo*hi + hi*lo), A fragments read from
// the staged rows with 128-bit loads (the k index is permuted so a lane's four k's are adjacent in memory),
// B fragments (dxagg_i, constant over the row) held in registers.  After the MMAs lane (g, t) owns
// dalpha of edges {g, g+8, g+16, g+24} x heads {2t, 2t+1}.
// =====================================================================================================
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// Five CTAs per SM (shared memory allows them at MT = 1): the kernel is latency bound, at four it measured 11.0 ms
// instead of 8.8 ms alone -- ten warps are three on one scheduler, 16384 / 3 / 32 = 170 registers per thread at most
// (the SC variant is held there by its launch bounds; the plain one needs 167 on its own).
// SC: rows of at most two chunks take c_i from their own entries (see below); hub segments never do.
// KX: the row has exactly KPMAX pairs of k-steps, so the k loop carries no run-time bound (no predicate, no
// reconvergence point per step, and the fragment loads of the next step can be scheduled across the MMAs).
template <int KPMAX, int MT, bool HUB, bool SC = false, bool KX = false>  // MT: 16-edge MMA tiles per chunk (chunk = 16*MT stored entries)
__global__ void __launch_bounds__(XW * 32, (SC && MT == 1) ? 5 : 0) attn_x_bwd_mma_kernel(const __grid_constant__ XArgs a) {
  constexpr int CH = 16 * MT;
  extern __shared__ __align__(128) float x_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int H = a.H, Fp = a.Fp, RS = a.RS;
  const int KP = KX ? KPMAX : (Fp + 15) >> 4;  // pairs of k-steps
  const int h0 = 2 * tg, h1 = 2 * tg + 1;
  // per warp: two row buffers (the next chunk's rows land while this chunk is computed), one staging area for
  // the row state (dxagg_i, xagg_i) of the next destination row, two mbarriers
  float* rowbuf = x_smem + (size_t)warp * xmma_warp_floats(RS, H, Fp, CH);
  float* dxs = rowbuf + 2 * CH * RS;
  float* xas = dxs + H * Fp;
  const unsigned bar0 = smem_u32(rowbuf + 2 * CH * RS + (2 * H * Fp + 31) / 32 * 32);
  for (int i = lane; i < 2 * CH * RS; i += 32) rowbuf[i] = 0.f;  // stale lanes of the MMA must hold finite numbers
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic zero fill before the async-proxy copies
  __syncwarp();
  unsigned phases = 0;  // bit b: parity the next wait on barrier b expects
  const int seg = blockIdx.x * XW + warp;
  if (HUB && seg >= a.n_hub_seg) return;
  ChunkIter<HUB> it;
  it.init(a, lane, seg, CH);
  const unsigned row_bytes = (unsigned)a.Sx * 16u, st_bytes = (unsigned)(H * Fp) * 4u;

  // copies of one chunk: its neighbour rows and, for the first chunk of a destination row, the row state
  auto issue = [&](const Chunk& ch, int jcol, int buf) {
    if (ch.cnt > 0) {
      const unsigned bar = bar0 + 8 * buf;
      float* rb = rowbuf + buf * CH * RS;
      const unsigned st_tx = ch.first ? ((SC && ch.two) ? st_bytes : 2 * st_bytes) : 0u;
      if (a.use_g4) {
        const int ngrp = (ch.cnt + 3) >> 2;
        const int r0 = __shfl_sync(FULL, jcol, (4 * lane) & 31), r1 = __shfl_sync(FULL, jcol, (4 * lane + 1) & 31);
        const int r2 = __shfl_sync(FULL, jcol, (4 * lane + 2) & 31), r3 = __shfl_sync(FULL, jcol, (4 * lane + 3) & 31);
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(ngrp * 16 * RS) + st_tx)
                       : "memory");
        __syncwarp();
        if (lane < ngrp) x_gather4(a, rb + lane * 4 * RS, r0, r1, r2, r3, bar);
      } else {
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                       "r"(row_bytes * (unsigned)ch.cnt + st_tx)
                       : "memory");
        __syncwarp();
        if (lane < ch.cnt) {
          const float* src = a.xg + (int64_t)jcol * a.ldxg;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(rb + lane * RS)),
                       "l"(src), "r"(row_bytes), "r"(bar)
                       : "memory");
        }
      }
      // the row state rides on the same barrier (lanes that issue no row copies): dxagg_i always, xagg_i only for
      // rows of more than two chunks (shorter rows get c_i = sum_j alpha_ij dalpha_ij from their own entries)
      if (ch.first && (lane == 30 || (lane == 31 && !(SC && ch.two)))) {
        const float* src = lane == 30 ? a.dxagg + (int64_t)ch.row * a.ldd : a.xagg + (int64_t)ch.row * a.ldxa;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(lane == 30 ? dxs : xas)),
                     "l"(src), "r"(st_bytes), "r"(bar)
                     : "memory");
      }
    }
  };
  auto cols_of = [&](const Chunk& ch) { return (ch.ok && lane < ch.cnt) ? __ldg(a.col + ch.base + lane) : 0; };

  uint32_t bhi[KPMAX][4], blo[KPMAX][4];
  float f0 = 0.f, f1 = 0.f, l0 = 0.f, l1 = 0.f, c0 = 0.f, c1 = 0.f, df0 = 0.f, df1 = 0.f;
  float nf0 = 0.f, nf1 = 0.f, nl0 = 0.f, nl1 = 0.f;  // f, lse of the NEXT destination row (prefetched)
  int64_t ipos[2 * MT], npos[2 * MT];              // where this lane's ds entries go (this chunk / the next one)
  // rows of at most two chunks: c_ih = sum_j alpha_ij dalpha_ij is summed over the row's own entries (cs0, cs1),
  // so the first chunk of a two-chunk row is held back until c is known -- in shared memory (weights, dalpha,
  // column ids): the kernel needs <= 168 registers to keep five CTAs on an SM
  auto prefetch_pos = [&](const Chunk& ch) {
#pragma unroll
    for (int q = 0; q < 2 * MT; ++q) {
      const int e = 16 * (q >> 1) + g + 8 * (q & 1);
      npos[q] = (a.ds && ch.ok && e < ch.cnt) ? (a.iperm ? (int64_t)__ldg(a.iperm + ch.base + e) : ch.base + e) : 0;
    }
  };
  auto prefetch_fl = [&](const Chunk& ch) {
    if (ch.ok && ch.first) {
      nf0 = h0 < H ? __ldg(a.f + (int64_t)ch.row * a.ldf + h0) : 0.f;
      nf1 = h1 < H ? __ldg(a.f + (int64_t)ch.row * a.ldf + h1) : 0.f;
      nl0 = h0 < H ? __ldg(a.lse + (int64_t)ch.row * H + h0) : 0.f;
      nl1 = h1 < H ? __ldg(a.lse + (int64_t)ch.row * H + h1) : 0.f;
    }
  };
  Chunk c = it.next(a, lane);
  {
    const int jc = cols_of(c);
    if (c.ok) issue(c, jc, 0);
    prefetch_fl(c);
    prefetch_pos(c);
  }
  int jc = cols_of(c);  // column ids of the chunk being computed (lane e holds entry e's source)
  const uint64_t dg_pol = l2_policy_evict_last();
  Chunk n = it.next(a, lane);
  int jn = cols_of(n);
  int p = 0;
  while (c.ok) {
    const Chunk nn = it.next(a, lane);
    const int jnn = cols_of(nn);  // column ids two chunks ahead
    const float* rows = rowbuf + p * CH * RS;
    if (c.cnt > 0) {
      mbar_wait(bar0 + 8 * p, (phases >> p) & 1u);
      phases ^= 1u << p;
    }
#pragma unroll
    for (int q = 0; q < 2 * MT; ++q) ipos[q] = npos[q];
    if (c.first) {
      f0 = nf0; f1 = nf1; l0 = nl0; l1 = nl1;
      df0 = df1 = 0.f;
      if (c.cnt > 0) {
        // row state from the staging area: B fragments of dxagg_i (head g, features 16kp + 4tg .. +3),
        // c_ih = dxagg_ih . xagg_ih (rows of more than two chunks; the others sum it over their entries below)
        float cp = 0.f;
        const bool staged_c = !(SC && c.two);
#pragma unroll
        for (int kp = 0; kp < KPMAX; ++kp) {
          float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kp < KP && g < H && 16 * kp + 4 * tg < Fp) {
            d = *reinterpret_cast<const float4*>(dxs + g * Fp + 16 * kp + 4 * tg);
            if (staged_c) cp += dot4(d, *reinterpret_cast<const float4*>(xas + g * Fp + 16 * kp + 4 * tg));
          }
          split_tf32(d.x, bhi[kp][0], blo[kp][0]);
          split_tf32(d.y, bhi[kp][1], blo[kp][1]);
          split_tf32(d.z, bhi[kp][2], blo[kp][2]);
          split_tf32(d.w, bhi[kp][3], blo[kp][3]);
        }
        cp += __shfl_xor_sync(FULL, cp, 1);
        cp += __shfl_xor_sync(FULL, cp, 2);  // lanes (g, *) hold c of head g
        c0 = __shfl_sync(FULL, cp, 4 * h0);
        c1 = __shfl_sync(FULL, cp, 4 * h1);
      }
    }
    __syncwarp();  // the staging area and the other row buffer are free: start the next chunk's copies
    if (n.ok) issue(n, jn, p ^ 1);
    prefetch_fl(n);
    prefetch_pos(n);
    const int n_mt = (MT > 1 && c.cnt > 16) ? 2 : 1;
    // three independent accumulators per 16-edge tile (hi*hi, lo*hi, hi*lo): the MMAs of a k-step do not
    // wait on each other, and the small compensation terms are summed apart from the main product
    float acc[MT][3][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int t3 = 0; t3 < 3; ++t3)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mt][t3][q] = 0.f;
#pragma unroll
    for (int kp = 0; kp < KPMAX; ++kp) {
      if (kp < KP) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (mt < n_mt) {
            const float4 ag = *reinterpret_cast<const float4*>(rows + (16 * mt + g) * RS + 16 * kp + 4 * tg);
            const float4 a8 = *reinterpret_cast<const float4*>(rows + (16 * mt + g + 8) * RS + 16 * kp + 4 * tg);
            uint32_t gh[4], gl[4], eh[4], el[4];
            split_tf32(ag.x, gh[0], gl[0]); split_tf32(ag.y, gh[1], gl[1]);
            split_tf32(ag.z, gh[2], gl[2]); split_tf32(ag.w, gh[3], gl[3]);
            split_tf32(a8.x, eh[0], el[0]); split_tf32(a8.y, eh[1], el[1]);
            split_tf32(a8.z, eh[2], el[2]); split_tf32(a8.w, eh[3], el[3]);
            // k-step 2kp: logical k = tg -> feature +0, tg+4 -> +1;  k-step 2kp+1: +2, +3
            mma_tf32(acc[mt][0], gh[0], eh[0], gh[1], eh[1], bhi[kp][0], bhi[kp][1]);
            mma_tf32(acc[mt][1], gl[0], el[0], gl[1], el[1], bhi[kp][0], bhi[kp][1]);
            mma_tf32(acc[mt][2], gh[0], eh[0], gh[1], eh[1], blo[kp][0], blo[kp][1]);
            mma_tf32(acc[mt][0], gh[2], eh[2], gh[3], eh[3], bhi[kp][2], bhi[kp][3]);
            mma_tf32(acc[mt][1], gl[2], el[2], gl[3], el[3], bhi[kp][2], bhi[kp][3]);
            mma_tf32(acc[mt][2], gh[2], eh[2], gh[3], eh[3], blo[kp][2], blo[kp][3]);
          }
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mt][0][q] += acc[mt][1][q] + acc[mt][2][q];
    if constexpr (!SC) {
      // ---- ds for this lane's (edge, head) pairs: acc[mt][0][0..1] = edge 16mt+g, heads h0,h1; [2..3] = edge 16mt+g+8
  #pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
  #pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int e = 16 * mt + g + 8 * half;
          const bool ev = e < c.cnt;
          float v0 = 0.f, v1 = 0.f;
          if (ev) {
            const float2 gq = *reinterpret_cast<const float2*>(rows + e * RS + Fp + h0);
            const float z0 = f0 + gq.x, z1 = f1 + gq.y;
            const float s0 = z0 > 0.f ? z0 : a.alpha * z0, s1 = z1 > 0.f ? z1 : a.alpha * z1;
            v0 = h0 < H ? expf(s0 - l0) * (acc[mt][0][2 * half] - c0) * (z0 > 0.f ? 1.f : a.alpha) : 0.f;
            v1 = h1 < H ? expf(s1 - l1) * (acc[mt][0][2 * half + 1] - c1) * (z1 > 0.f ? 1.f : a.alpha) : 0.f;
            df0 += v0;
            df1 += v1;
            if (a.ds) {
              float* dp = a.ds + ipos[2 * mt + half] * H + h0;
              if (h1 < H && (H & 1) == 0) {
                *reinterpret_cast<float2*>(dp) = make_float2(v0, v1);
              } else {
                if (h0 < H) dp[0] = v0;
                if (h1 < H) dp[1] = v1;
              }
            }
          }
          if (a.dgacc) {  // dg_j += ds_ij, straight into the per-source array (warp-uniform branch)
            const int je = __shfl_sync(FULL, jc, e & 31);
            float* gp = a.dgacc + (int64_t)je * a.lddgacc + h0;
            if (a.dg_vec4) {  // lanes tg = 0, 2 of a group send heads 0..3 / 4..7 of the group's entry
              const float q0 = __shfl_xor_sync(FULL, v0, 1), q1 = __shfl_xor_sync(FULL, v1, 1);
              if (ev && !(tg & 1)) {
                if (a.dg_vec4 == 2) red_add_v4_pol(gp, v0, v1, q0, q1, dg_pol);
                else red_add_v4(gp, v0, v1, q0, q1);
              }
            } else if (ev) {
              if (h0 < H) red_add_f32(gp, v0);
              if (h1 < H) red_add_f32(gp + 1, v1);
            }
          }
        }
      }
    } else {
      // ---- ds for this lane's (edge, head) pairs: acc[mt][0][0..1] = edge 16mt+g, heads h0,h1; [2..3] = edge 16mt+g+8
      // w = alpha_ij LeakyReLU'(z_ij), da = dalpha_ij; ds_ij = w (da - c_i)
      float wv[MT][2][2], dav[MT][2][2];
      float cs0 = 0.f, cs1 = 0.f;  // this chunk's part of c_i = sum_j alpha_ij dalpha_ij
  #pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
  #pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int e = 16 * mt + g + 8 * half;
          float w0 = 0.f, w1 = 0.f;
          if (e < c.cnt) {
            const float2 gq = *reinterpret_cast<const float2*>(rows + e * RS + Fp + h0);
            const float z0 = f0 + gq.x, z1 = f1 + gq.y;
            const float s0 = z0 > 0.f ? z0 : a.alpha * z0, s1 = z1 > 0.f ? z1 : a.alpha * z1;
            const float p0 = h0 < H ? expf(s0 - l0) : 0.f, p1 = h1 < H ? expf(s1 - l1) : 0.f;
            cs0 += p0 * acc[mt][0][2 * half];
            cs1 += p1 * acc[mt][0][2 * half + 1];
            w0 = p0 * (z0 > 0.f ? 1.f : a.alpha);
            w1 = p1 * (z1 > 0.f ? 1.f : a.alpha);
          }
          wv[mt][half][0] = w0;
          wv[mt][half][1] = w1;
          dav[mt][half][0] = acc[mt][0][2 * half];
          dav[mt][half][1] = acc[mt][0][2 * half + 1];
        }
      }
      auto emit = [&](const float (&w)[MT][2][2], const float (&da)[MT][2][2], const int64_t (&pos)[2 * MT], int jcol, int cnt) {
  #pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
  #pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int e = 16 * mt + g + 8 * half;
            const bool ev = e < cnt;
            const float v0 = ev ? w[mt][half][0] * (da[mt][half][0] - c0) : 0.f;
            const float v1 = ev ? w[mt][half][1] * (da[mt][half][1] - c1) : 0.f;
            if (ev) {
              df0 += v0;
              df1 += v1;
              if (a.ds) {
                float* dp = a.ds + pos[2 * mt + half] * H + h0;
                if (h1 < H && (H & 1) == 0) {
                  *reinterpret_cast<float2*>(dp) = make_float2(v0, v1);
                } else {
                  if (h0 < H) dp[0] = v0;
                  if (h1 < H) dp[1] = v1;
                }
              }
            }
            if (a.dgacc) {  // dg_j += ds_ij, straight into the per-source array (warp-uniform branch)
              const int je = __shfl_sync(FULL, jcol, e & 31);
              float* gp = a.dgacc + (int64_t)je * a.lddgacc + h0;
              if (a.dg_vec4) {  // lanes tg = 0, 2 of a group send heads 0..3 / 4..7 of the group's entry
                const float q0 = __shfl_xor_sync(FULL, v0, 1), q1 = __shfl_xor_sync(FULL, v1, 1);
                if (ev && !(tg & 1)) {
                  if (a.dg_vec4 == 2) red_add_v4_pol(gp, v0, v1, q0, q1, dg_pol);
                  else red_add_v4(gp, v0, v1, q0, q1);
                }
              } else if (ev) {
                if (h0 < H) red_add_f32(gp, v0);
                if (h1 < H) red_add_f32(gp + 1, v1);
              }
            }
          }
        }
      };
      // [4 MT w | 4 MT dalpha | col | cs0 | cs1][32 lanes], behind the barriers
      float* pend = dxs + (2 * H * Fp + 31) / 32 * 32 + 32 + lane;
      if (c.two && !c.last) {  // first chunk of a two-chunk row: held back until the second one has completed c
  #pragma unroll
        for (int mt = 0; mt < MT; ++mt)
  #pragma unroll
          for (int q = 0; q < 4; ++q) {
            pend[(4 * mt + q) * 32] = wv[mt][q >> 1][q & 1];
            pend[(4 * MT + 4 * mt + q) * 32] = dav[mt][q >> 1][q & 1];
          }
        pend[8 * MT * 32] = __int_as_float(jc);
        pend[(8 * MT + 1) * 32] = cs0;
        pend[(8 * MT + 2) * 32] = cs1;
      } else {
        if (c.two) {
          if (!c.first) {
            cs0 += pend[(8 * MT + 1) * 32];
            cs1 += pend[(8 * MT + 2) * 32];
          }
  #pragma unroll
          for (int o = 4; o < 32; o <<= 1) {  // lanes of one tg hold the same heads: sum over the 8 entry groups
            cs0 += __shfl_xor_sync(FULL, cs0, o);
            cs1 += __shfl_xor_sync(FULL, cs1, o);
          }
          c0 = cs0;
          c1 = cs1;
          if (!c.first) {  // the held-back chunk (a non-last chunk is always full)
            float pw[MT][2][2], pda[MT][2][2];
            int64_t ppos[2 * MT];
            const int64_t pbase = c.base - CH;  // the second chunk starts CH entries after the first
  #pragma unroll
            for (int mt = 0; mt < MT; ++mt)
  #pragma unroll
              for (int q = 0; q < 4; ++q) {
                pw[mt][q >> 1][q & 1] = pend[(4 * mt + q) * 32];
                pda[mt][q >> 1][q & 1] = pend[(4 * MT + 4 * mt + q) * 32];
              }
  #pragma unroll
            for (int q = 0; q < 2 * MT; ++q) {
              const int e = 16 * (q >> 1) + g + 8 * (q & 1);
              ppos[q] = a.ds ? (a.iperm ? (int64_t)__ldg(a.iperm + pbase + e) : pbase + e) : 0;
            }
            emit(pw, pda, ppos, __float_as_int(pend[8 * MT * 32]), CH);
          }
        }
        emit(wv, dav, ipos, jc, c.cnt);
      }
    }
    if (c.last) {
      float d0 = df0, d1 = df1;
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        d0 += __shfl_xor_sync(FULL, d0, o);
        d1 += __shfl_xor_sync(FULL, d1, o);
      }
      if (g == 0) {
        float* dst = HUB ? a.scratch + (int64_t)seg * H : a.df + (int64_t)c.row * a.lddf;
        if (h0 < H) dst[h0] = d0;
        if (h1 < H) dst[h1] = d1;
      }
    }
    __syncwarp();
    c = n;
    jc = jn;
    n = nn;
    jn = jnn;
    p ^= 1;
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
        const int64_t e0 = perm ? __ldg(perm + t) : t, e1 = perm ? __ldg(perm + t + EPG) : t + EPG;
        const int64_t e2 = perm ? __ldg(perm + t + 2 * EPG) : t + 2 * EPG, e3 = perm ? __ldg(perm + t + 3 * EPG) : t + 3 * EPG;
        const float v0 = __ldg(ds + e0 * H + my_h), v1 = __ldg(ds + e1 * H + my_h);
        const float v2 = __ldg(ds + e2 * H + my_h), v3 = __ldg(ds + e3 * H + my_h);
        acc += (v0 + v1) + (v2 + v3);
      }
      for (; t < end; t += EPG) acc += __ldg(ds + (perm ? (int64_t)__ldg(perm + t) : t) * H + my_h);
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
    for (int64_t t = beg + threadIdx.x; t < end; t += 256) acc += __ldg(ds + (perm ? (int64_t)__ldg(perm + t) : t) * H + h);
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

// =====================================================================================================
// f_i = x_i . u_h, g_i = x_i . v_h  (u = W a_src, v = W a_dst: layers.py:141-144 folded into the input) and
// the gather rows xg_i = [x_i | 0.. | g_i | 0..].  Warp per row, lanes = float4 slots, [u|v] transposed in
// shared memory.
// =====================================================================================================
// Copies of the packed rows on the other GPUs of the shard group (peer-mapped memory, written over NVLink
// straight from this kernel: the pack IS the all-gather).  base[q] points at the row of peer q's buffer that
// corresponds to this rank's local row 0; rows != 0 pushes whole rows, else only the g columns.
struct PackPeers {
  int n;
  int rows;
  float* base[GATK_MAX_PEERS];
};

template <int HP, bool VEC>
__global__ void __launch_bounds__(256) logits_pack_kernel(int64_t n, int F, int H, const float* __restrict__ x, int64_t ldx,
                                                          const float* __restrict__ uv, int64_t lduv, int Fp, int P,
                                                          float* __restrict__ xg, float* __restrict__ f, int64_t ldf,
                                                          const PackPeers peers) {
  extern __shared__ __align__(16) float uvs[];  // [2*HP][Fp]
  constexpr int C2 = 2 * HP;
  constexpr int SHC = 5 - Log2<C2>::v;
  constexpr int R = 4;  // rows per warp iteration: one read of the [u|v] fragment serves R rows
  for (int i = threadIdx.x; i < C2 * Fp; i += blockDim.x) {
    const int c = i / Fp, k = i - c * Fp;
    const int h = c < HP ? c : c - HP;
    uvs[i] = (k < F && h < H) ? uv[(int64_t)k * lduv + (c < HP ? h : H + h)] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int S = Fp >> 2;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; row0 < n; row0 += warps * R) {
    float part[R][C2];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < C2; ++c) part[r][c] = 0.f;
    for (int slot = lane; slot < S; slot += 32) {
      float4 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r < n ? row0 + r : n - 1;
        const float* xr = x + row * ldx;
        if (VEC) {
          v[r] = ldg4_stream(xr + slot * 4);
        } else {
          const int k = slot * 4;
          v[r].x = k < F ? __ldg(xr + k) : 0.f;
          v[r].y = k + 1 < F ? __ldg(xr + k + 1) : 0.f;
          v[r].z = k + 2 < F ? __ldg(xr + k + 2) : 0.f;
          v[r].w = k + 3 < F ? __ldg(xr + k + 3) : 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (row0 + r < n) {
          stg4(xg + (row0 + r) * P + slot * 4, v[r]);
          if (peers.rows)
            for (int q = 0; q < peers.n; ++q) stg4(peers.base[q] + (row0 + r) * P + slot * 4, v[r]);
        }
#pragma unroll
      for (int c = 0; c < C2; ++c) {
        const float4 q = *reinterpret_cast<const float4*>(uvs + c * Fp + slot * 4);
#pragma unroll
        for (int r = 0; r < R; ++r) part[r][c] += dot4(v[r], q);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
      butterfly_scatter<C2>(part[r], lane);  // lane (c << SHC) holds column c
      if (row < n) {
        if ((lane & ((1 << SHC) - 1)) == 0) {
          const int c = lane >> SHC;
          if (c < H) f[row * ldf + c] = part[r][0];
        }
      }
      for (int t0 = 0; t0 < P - Fp; t0 += 32) {  // g behind the input row, zero padding after it
        const int t = t0 + lane;
        const float gval = __shfl_sync(FULL, part[r][0], ((HP + (t < HP ? t : 0)) << SHC) & 31);
        if (row < n && t < P - Fp) {
          const float val = t < H ? gval : 0.f;
          xg[row * P + Fp + t] = val;
          for (int q = 0; q < peers.n; ++q) peers.base[q][row * P + Fp + t] = val;
        }
      }
    }
  }
}

// The same pack on the tensor cores (Fp <= 128, 16-byte aligned rows): [16 rows x Fp] x [Fp x 16 columns (u | v)] per
// warp tile with mma.sync m16n8k8 tf32 and the 3xTF32 split (fp32 parity).  The lane that loads a 16-byte piece of
// an input row for the A fragment also stores it into the gather row, so x is read once and xg written once with
// full 32-byte sectors; the split [u|v] fragments sit pre-packed per lane in shared memory (one LDS.128 per MMA
// pair).  k is permuted inside each 16-feature block so that a lane's four k values are adjacent in memory
// (k-step 2kp: logical k = tg -> feature 16kp+4tg, tg+4 -> +1; k-step 2kp+1: +2, +3), as in attn_x_bwd_mma_kernel.
// The SIMT kernel above needed 16 dot products per row on the FMA pipe (66 % issue-active at 0.31 of the byte floor).
template <int KPMAX>
__global__ void __launch_bounds__(256) logits_pack_mma_kernel(int64_t n, int F, int H, const float* __restrict__ x, int64_t ldx,
                                                              const float* __restrict__ uv, int64_t lduv, int Fp, int P,
                                                              float* __restrict__ xg, float* __restrict__ f, int64_t ldf,
                                                              const PackPeers peers) {
  extern __shared__ __align__(16) float4 bfrag[];  // [KP][ks 2][nt 2][lane 32] = {b0 hi, b1 hi, b0 lo, b1 lo}
  const int KP = (Fp + 15) >> 4;
  for (int i = threadIdx.x; i < KP * 128; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) & 1, ks = (i >> 6) & 1, kp = i >> 7;
    const int gq = ln >> 2, tq = ln & 3;
    const int f0 = 16 * kp + 4 * tq + 2 * ks;
    const int colq = nt == 0 ? gq : H + gq;
    const float b0 = (gq < H && f0 < F) ? uv[(int64_t)f0 * lduv + colq] : 0.f;
    const float b1 = (gq < H && f0 + 1 < F) ? uv[(int64_t)(f0 + 1) * lduv + colq] : 0.f;
    uint32_t h0, l0, h1, l1;
    split_tf32(b0, h0, l0);
    split_tf32(b1, h1, l1);
    bfrag[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 16; row0 < n; row0 += warps * 16) {
    const int64_t ra = row0 + g, rb = row0 + g + 8;
    const bool va = ra < n, vb = rb < n;
    float4 xa[KPMAX], xb[KPMAX];
#pragma unroll
    for (int kp = 0; kp < KPMAX; ++kp) {
      const int c = 16 * kp + 4 * tg;
      const bool in = kp < KP && c < Fp;
      xa[kp] = (in && va) ? ldg4_stream(x + ra * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      xb[kp] = (in && vb) ? ldg4_stream(x + rb * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (F != Fp) {  // F % 4 != 0 never takes this kernel; kept for clarity: the pad columns of xg are zeros
    }
    // the input columns of the gather rows (and of the peers' copies when whole rows are pushed)
#pragma unroll
    for (int kp = 0; kp < KPMAX; ++kp) {
      const int c = 16 * kp + 4 * tg;
      if (kp < KP && c < Fp) {
        if (va) stg4(xg + ra * P + c, xa[kp]);
        if (vb) stg4(xg + rb * P + c, xb[kp]);
        if (peers.rows)
          for (int q = 0; q < peers.n; ++q) {
            if (va) stg4(peers.base[q] + ra * P + c, xa[kp]);
            if (vb) stg4(peers.base[q] + rb * P + c, xb[kp]);
          }
      }
    }
    float acc[2][2][4];  // [n-tile: u, v][main, compensation][fragment]
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[nt][t][q] = 0.f;
#pragma unroll
    for (int kp = 0; kp < KPMAX; ++kp) {
      if (kp < KP) {
        uint32_t ah[4], al[4], bh[4], bl[4];
        split_tf32(xa[kp].x, ah[0], al[0]); split_tf32(xa[kp].y, ah[1], al[1]);
        split_tf32(xa[kp].z, ah[2], al[2]); split_tf32(xa[kp].w, ah[3], al[3]);
        split_tf32(xb[kp].x, bh[0], bl[0]); split_tf32(xb[kp].y, bh[1], bl[1]);
        split_tf32(xb[kp].z, bh[2], bl[2]); split_tf32(xb[kp].w, bh[3], bl[3]);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const float4 bq = bfrag[((kp * 2 + ks) * 2 + nt) * 32 + lane];
            const uint32_t b0h = __float_as_uint(bq.x), b1h = __float_as_uint(bq.y);
            const uint32_t b0l = __float_as_uint(bq.z), b1l = __float_as_uint(bq.w);
            mma_tf32(acc[nt][0], ah[2 * ks], bh[2 * ks], ah[2 * ks + 1], bh[2 * ks + 1], b0h, b1h);
            mma_tf32(acc[nt][1], al[2 * ks], bl[2 * ks], al[2 * ks + 1], bl[2 * ks + 1], b0h, b1h);
            mma_tf32(acc[nt][1], ah[2 * ks], bh[2 * ks], ah[2 * ks + 1], bh[2 * ks + 1], b0l, b1l);
          }
        }
      }
    }
    // fragment: [0] = (row g, col 2tg), [1] = (g, 2tg+1), [2] = (g+8, 2tg), [3] = (g+8, 2tg+1)
    const int h0 = 2 * tg;
    float fu[4], gv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      fu[q] = acc[0][0][q] + acc[0][1][q];
      gv[q] = acc[1][0][q] + acc[1][1][q];
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t row = half ? rb : ra;
      if (half ? vb : va) {
        if (h0 < H) f[row * ldf + h0] = fu[2 * half];
        if (h0 + 1 < H) f[row * ldf + h0 + 1] = fu[2 * half + 1];
        // g behind the input row, zeros up to the pitch
        const float g0 = h0 < H ? gv[2 * half] : 0.f, g1 = h0 + 1 < H ? gv[2 * half + 1] : 0.f;
        float* tail = xg + row * P + Fp;
        if (Fp + h0 + 1 < P) {
          *reinterpret_cast<float2*>(tail + h0) = make_float2(g0, g1);
        } else if (Fp + h0 < P) {
          tail[h0] = g0;
        }
        for (int t = 8 + h0; t < P - Fp; t += 8) {
          tail[t] = 0.f;
          if (t + 1 < P - Fp) tail[t + 1] = 0.f;
        }
        for (int q = 0; q < peers.n; ++q) {
          float* pt = peers.base[q] + row * P + Fp;
          if (Fp + h0 + 1 < P) {
            *reinterpret_cast<float2*>(pt + h0) = make_float2(g0, g1);
          } else if (Fp + h0 < P) {
            pt[h0] = g0;
          }
          if (peers.rows)
            for (int t = 8 + h0; t < P - Fp; t += 8) {
              pt[t] = 0.f;
              if (t + 1 < P - Fp) pt[t + 1] = 0.f;
            }
        }
      }
    }
  }
}

static int check_x_geom(int H, int Fp, int* hp, int* ns, int* sx) {
  GATK_REQUIRE(H >= 1 && H <= 8, "aggregate-first form: H=%d out of range [1,8]", H);
  GATK_REQUIRE(Fp >= 4 && Fp % 4 == 0 && Fp <= 512, "aggregate-first form: Fp=%d must be a multiple of 4 in [4,512]", Fp);
  *hp = H <= 1 ? 1 : (H <= 2 ? 2 : (H <= 4 ? 4 : 8));
  *sx = Fp / 4 + (H + 3) / 4;
  *ns = *sx <= 32 ? 1 : (*sx <= 64 ? 2 : (*sx <= 128 ? 4 : 5));
  GATK_REQUIRE(*ns <= 4 && *hp * *ns <= 16, "aggregate-first form: H=%d x Fp=%d needs too many accumulators", H, Fp);
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
#define HP_DISPATCH(hp, CALL)                          \
  switch (hp) {                                        \
    case 1: { constexpr int HP = 1; CALL; } break;     \
    case 2: { constexpr int HP = 2; CALL; } break;     \
    case 4: { constexpr int HP = 4; CALL; } break;     \
    default: { constexpr int HP = 8; CALL; } break;    \
  }

typedef CUresult (*XEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Describe xg [n_src, ldxg] to TMA with a box of RS floats x 1 row (tile::gather4 stacks four such rows, so RS is
// the row pitch in shared memory; columns past ldxg are zero filled).  Rows wider than 256 floats, or a driver
// without the entry point, keep the per-row bulk copies.
static void make_xg_map(XArgs& a) {
  a.use_g4 = 0;
  static const bool off = getenv("GATK_NO_GATHER4") != nullptr;
  if (off || a.RS > 256 || (a.RS & 7) || a.n_src <= 0) return;
  static XEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<XEncodeTiledFn>(p);
  }
  if (!fn) return;
  cuuint64_t dims[2] = {(cuuint64_t)a.ldxg, (cuuint64_t)a.n_src};
  cuuint64_t strides[1] = {(cuuint64_t)a.ldxg * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)a.RS, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(&a.xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.xg), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  a.use_g4 = rc == CUDA_SUCCESS ? 1 : 0;
}

template <typename KH, typename KM>
static int launch_x(KH hub_kernel, KM main_kernel, const XArgs& a, int chunk, size_t smem, cudaStream_t st) {
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(hub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hub_kernel<<<(a.n_hub_seg + XW - 1) / XW, XW * 32, smem, st>>>(a, chunk);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_dst > 0) {
    int grid = 0;
    if (int rc = persistent_grid(main_kernel, XW * 32, smem, &grid)) return rc;
    const int64_t need = a.item_ptr ? (a.n_items + XW - 1) / XW : (a.n_dst + (int64_t)XW * GRAB - 1) / ((int64_t)XW * GRAB);
    if (need < grid) grid = (int)need;
    main_kernel<<<grid, XW * 32, smem, st>>>(a, chunk);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int HP, int NS>
static int launch_x_fwd(XArgs a, cudaStream_t st) {
  make_xg_map(a);
  static const int chunk_env = getenv("GATK_XFWD_CHUNK") ? atoi(getenv("GATK_XFWD_CHUNK")) : 0;
  int chunk = x_chunk_for(a.RS);
  if (chunk_env == 8 || chunk_env == 16) chunk = chunk_env < chunk ? chunk_env : chunk;
  const size_t smem = (size_t)XW * xfwd_warp_floats(a.RS, HP, chunk) * sizeof(float);
  if (int rc = launch_x(attn_x_fwd_kernel<HP, NS, true>, attn_x_fwd_kernel<HP, NS, false>, a, chunk, smem, st)) return rc;
  if (a.n_hub_seg > 0) {  // stream order: the merge only needs the segment kernel, which ran first
    attn_x_fwd_hub_merge_kernel<<<a.n_hub, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <typename KH, typename KM>
static int launch_x_mma(KH hub_kernel, KM main_kernel, const XArgs& a, size_t smem, cudaStream_t st) {
  if (a.n_hub_seg > 0) {
    if (smem > 48 * 1024)
      GATK_CHECK_CUDA(cudaFuncSetAttribute(hub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hub_kernel<<<(a.n_hub_seg + XW - 1) / XW, XW * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  if (a.n_dst > 0) {
    int grid = 0;
    if (int rc = persistent_grid(main_kernel, XW * 32, smem, &grid)) return rc;
    const int64_t need = a.item_ptr ? (a.n_items + XW - 1) / XW : (a.n_dst + (int64_t)XW * GRAB - 1) / ((int64_t)XW * GRAB);
    if (need < grid) grid = (int)need;
    main_kernel<<<grid, XW * 32, smem, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

static int launch_x_bwd_mma(XArgs a, cudaStream_t st) {
  a.RS = xmma_row_pitch(a.Sx);
  make_xg_map(a);
  static const int mt_env = getenv("GATK_XBWD_MT") ? atoi(getenv("GATK_XBWD_MT")) : 1;
  const int mt = mt_env == 2 ? 2 : 1;
  // GATK_XBWD_SHORT_C=1 (read per call, so a test can switch it): rows of at most two chunks sum c_i over their own
  // entries instead of reading xagg_i.  Measured at the products shape (profiles/r02c_xbwd_short_c.md): DRAM traffic
  // 50.7 -> 43.6 GB, but 4.35 -> 4.94 G instructions at the same issue rate (the kernel is bound by dependent-issue
  // latency at 2.5 warps per scheduler, not by bytes): 8.73 -> 10.27 ms alone, 10.3 -> 11.6 ms in the step.  Off.
  const char* sc_env = getenv("GATK_XBWD_SHORT_C");
  a.short_c = (sc_env && atoi(sc_env) == 1) ? 1 : 0;
  const size_t smem = (size_t)XW * xmma_warp_floats(a.RS, a.H, a.Fp, 16 * mt) * sizeof(float);
  int rc;
  if (!a.short_c && mt == 1) {  // the default: one instantiation per k extent
    switch ((a.Fp + 15) >> 4) {
#define GATK_XBWD_KX(K)                                                                                                  \
  case K:                                                                                                                \
    rc = launch_x_mma(attn_x_bwd_mma_kernel<K, 1, true, false, true>, attn_x_bwd_mma_kernel<K, 1, false, false, true>, a, \
                      smem, st);                                                                                         \
    break;
      GATK_XBWD_KX(1) GATK_XBWD_KX(2) GATK_XBWD_KX(3) GATK_XBWD_KX(4) GATK_XBWD_KX(5) GATK_XBWD_KX(6) GATK_XBWD_KX(7)
      GATK_XBWD_KX(8)
#undef GATK_XBWD_KX
      default:
        set_error("Fp=%d too wide for the tensor-core backward", a.Fp);
        return 3;
    }
  } else if (a.short_c && mt == 1) {  // (the hub segments always stage xagg_i)
    rc = a.Fp <= 64 ? launch_x_mma(attn_x_bwd_mma_kernel<4, 1, true>, attn_x_bwd_mma_kernel<4, 1, false, true>, a, smem, st)
                    : launch_x_mma(attn_x_bwd_mma_kernel<8, 1, true>, attn_x_bwd_mma_kernel<8, 1, false, true>, a, smem, st);
  } else if (a.Fp <= 64) {
    a.short_c = 0;
    rc = mt == 2 ? launch_x_mma(attn_x_bwd_mma_kernel<4, 2, true>, attn_x_bwd_mma_kernel<4, 2, false>, a, smem, st)
                 : launch_x_mma(attn_x_bwd_mma_kernel<4, 1, true>, attn_x_bwd_mma_kernel<4, 1, false>, a, smem, st);
  } else {
    a.short_c = 0;
    rc = mt == 2 ? launch_x_mma(attn_x_bwd_mma_kernel<8, 2, true>, attn_x_bwd_mma_kernel<8, 2, false>, a, smem, st)
                 : launch_x_mma(attn_x_bwd_mma_kernel<8, 1, true>, attn_x_bwd_mma_kernel<8, 1, false>, a, smem, st);
  }
  if (rc) return rc;
  if (a.n_hub_seg > 0) {
    attn_x_bwd_hub_merge_kernel<<<(a.n_hub * a.H + 127) / 128, 128, 0, st>>>(a);
    GATK_CHECK_LAUNCH();
  }
  return 0;
}

template <int HP>
static int launch_x_bwd(const XArgs& a, cudaStream_t st) {
  const int chunk = x_chunk_for(a.RS);
  const size_t smem = (size_t)XW * xbwd_warp_floats(a.RS, a.H, a.Fp, chunk) * sizeof(float);
  if (int rc = launch_x(attn_x_bwd_kernel<HP, true>, attn_x_bwd_kernel<HP, false>, a, chunk, smem, st)) return rc;
  if (a.n_hub_seg > 0) {
    attn_x_bwd_hub_merge_kernel<<<(a.n_hub * a.H + 127) / 128, 128, 0, st>>>(a);
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

// gather-row pitch: [x (Fp) | g (H)] padded to whole 64-byte DRAM atoms (rows that straddle atoms cost ~7 % more
// DRAM traffic in the edge passes than the padding adds; GATK_XG_ALIGN=4 gives the unpadded pitch)
extern "C" int64_t gatk_xg_pitch(int Fp, int H) {
  static int al = 0;
  if (!al) {
    const char* e = getenv("GATK_XG_ALIGN");
    al = e ? atoi(e) : 16;
    if (al != 4 && al != 8 && al != 16 && al != 32) al = 16;
  }
  const int64_t w = (int64_t)Fp + 4 * ((H + 3) / 4);
  return (w + al - 1) / al * al;
}

static int logits_pack_launch(int64_t n, int F, int H, const float* x, int64_t ldx, const float* uv, int64_t lduv,
                              float* xg, int64_t ldxg, float* f, int64_t ldf, const PackPeers& peers, void* stream) {
  GATK_REQUIRE(F >= 1 && H >= 1 && H <= 8, "bad sizes F=%d H=%d", F, H);
  const int Fp = (F + 3) / 4 * 4;
  GATK_REQUIRE(Fp <= 512, "F=%d too wide for the aggregate-first form", F);
  GATK_REQUIRE(x && uv && xg && f && ldx >= F && lduv >= 2 * H && ldf >= H, "bad arguments");
  GATK_REQUIRE(ldxg >= gatk_xg_pitch(Fp, H) && ldxg % 4 == 0 && ((uintptr_t)xg & 15) == 0,
               "xg must be 16-byte aligned with pitch >= gatk_xg_pitch(Fp, H), a multiple of 4 floats");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int hp = H <= 1 ? 1 : (H <= 2 ? 2 : (H <= 4 ? 4 : 8));
  const bool vec = (F % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)x & 15) == 0);
  const size_t smem = (size_t)2 * hp * Fp * sizeof(float);
  int64_t blocks = (n + 31) / 32;  // 8 warps x 4 rows per iteration
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  static const bool pack_simt = getenv("GATK_PACK_SIMT") != nullptr;
  if (vec && Fp <= 128 && (ldxg & 3) == 0 && !pack_simt) {  // tensor-core pack
    const int kp = (Fp + 15) >> 4;
    const size_t sm = (size_t)kp * 128 * sizeof(float4);
    int64_t blk = (n + 127) / 128;  // 8 warps x 16 rows
    const int64_t capm = (int64_t)sm_count() * 4;
    if (blk > capm) blk = capm;
    if (kp <= 4)
      logits_pack_mma_kernel<4><<<(unsigned)blk, 256, sm, st>>>(n, F, H, x, ldx, uv, lduv, Fp, (int)ldxg, xg, f, ldf, peers);
    else
      logits_pack_mma_kernel<8><<<(unsigned)blk, 256, sm, st>>>(n, F, H, x, ldx, uv, lduv, Fp, (int)ldxg, xg, f, ldf, peers);
    GATK_CHECK_LAUNCH();
    return 0;
  }
#define PACK_LAUNCH(HPV, V) logits_pack_kernel<HPV, V><<<(unsigned)blocks, 256, smem, st>>>(n, F, H, x, ldx, uv, lduv, Fp, (int)ldxg, xg, f, ldf, peers)
  if (vec) {
    HP_DISPATCH(hp, PACK_LAUNCH(HP, true));
  } else {
    HP_DISPATCH(hp, PACK_LAUNCH(HP, false));
  }
#undef PACK_LAUNCH
  GATK_CHECK_LAUNCH();
  return 0;
}

extern "C" int gatk_logits_pack(int64_t n, int F, int H, const float* x, int64_t ldx, const float* uv, int64_t lduv,
                                float* xg, int64_t ldxg, float* f, int64_t ldf, void* stream) {
  PackPeers none;
  none.n = 0;
  none.rows = 0;
  return logits_pack_launch(n, F, H, x, ldx, uv, lduv, xg, ldxg, f, ldf, none, stream);
}

extern "C" int gatk_logits_pack_push(int64_t n, int F, int H, const float* x, int64_t ldx, const float* uv, int64_t lduv,
                                     float* xg, int64_t ldxg, float* f, int64_t ldf, int n_peers, float* const* peer_xg,
                                     int whole_rows, void* stream) {
  GATK_REQUIRE(n_peers >= 0 && n_peers <= GATK_MAX_PEERS && (n_peers == 0 || peer_xg), "n_peers=%d (max %d)", n_peers,
               GATK_MAX_PEERS);
  PackPeers p;
  p.n = n_peers;
  p.rows = whole_rows ? 1 : 0;
  for (int q = 0; q < n_peers; ++q) {
    GATK_REQUIRE(peer_xg[q] && ((uintptr_t)peer_xg[q] & 15) == 0, "peer row pointer %d is null or not 16-byte aligned", q);
    p.base[q] = peer_xg[q];
  }
  return logits_pack_launch(n, F, H, x, ldx, uv, lduv, xg, ldxg, f, ldf, p, stream);
}

static int fill_xargs(XArgs& a, int64_t n_src, int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Fp, int sx,
                      const float* xg, int64_t ldxg, const float* f, int64_t ldf, float alpha, int seg_len,
                      const int32_t* hub_rows, const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                      int32_t* counter, const int32_t* item_ptr, int n_items) {
  if (int rc = check_hub(seg_len, n_hub, n_hub_seg, hub_rows, hub_seg_ptr, hub_scratch)) return rc;
  GATK_REQUIRE(n_dst < (1LL << 31), "n_dst too large for one shard");
  GATK_REQUIRE(rowptr && col && xg && f && counter, "null pointer argument");
  GATK_REQUIRE(ldxg % 4 == 0 && ldxg >= 4 * sx && ((uintptr_t)xg & 15) == 0 && ldf >= H,
               "xg rows must be 16-byte aligned with pitch >= Fp + 4*ceil(H/4); ldf >= H");
  GATK_REQUIRE(n_src >= 0 && n_src < (1LL << 31), "n_src out of range");
  a.use_g4 = 0; a.n_src = n_src;
  a.n_dst = n_dst; a.rowptr = rowptr; a.col = col; a.H = H; a.S = Fp / 4; a.Fp = Fp; a.Sx = sx; a.RS = x_row_pitch(sx);
  a.xg = xg; a.ldxg = ldxg; a.f = f; a.ldf = ldf; a.alpha = alpha;
  a.seg_len = seg_len; a.hub_rows = hub_rows; a.hub_seg_ptr = hub_seg_ptr; a.n_hub = n_hub; a.n_hub_seg = n_hub_seg;
  a.scratch = hub_scratch; a.counter = counter; a.item_ptr = item_ptr; a.n_items = n_items;
  return 0;
}

extern "C" int gatk_attn_x_fwd(int64_t n_src, int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Fp,
                               const float* xg, int64_t ldxg, const float* f, int64_t ldf, float alpha,
                               float* xagg, int64_t ldxa, float* lse, int seg_len, const int32_t* hub_rows,
                               const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                               int32_t* counter, const int32_t* item_ptr, int n_items, void* stream) {
  int hp, ns, sx;
  if (int rc = check_x_geom(H, Fp, &hp, &ns, &sx)) return rc;
  XArgs a = {};
  if (int rc = fill_xargs(a, n_src, n_dst, rowptr, col, H, Fp, sx, xg, ldxg, f, ldf, alpha, seg_len, hub_rows, hub_seg_ptr,
                          n_hub, n_hub_seg, hub_scratch, counter, item_ptr, n_items))
    return rc;
  GATK_REQUIRE(xagg && ldxa % 4 == 0 && ldxa >= (int64_t)H * Fp && ((uintptr_t)xagg & 15) == 0,
               "xagg must be 16-byte aligned with pitch >= H*Fp (multiple of 4 floats)");
  a.xagg = xagg; a.ldxa = ldxa; a.lse = lse;
  cudaStream_t st = (cudaStream_t)stream;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  X_DISPATCH(hp, ns, return (launch_x_fwd<HP, NS>(a, st)));
  return 0;
}

extern "C" int gatk_attn_x_bwd(int64_t n_src, int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Fp,
                               const float* xg, int64_t ldxg, const float* f, int64_t ldf, const float* lse, float alpha,
                               const float* xagg, int64_t ldxa, const float* dxagg, int64_t ldd, float* ds,
                               const int32_t* iperm, float* dg_acc, int64_t lddg_acc, float* df, int64_t lddf, int seg_len,
                               const int32_t* hub_rows, const int32_t* hub_seg_ptr, int n_hub,
                               int n_hub_seg, float* hub_scratch, int32_t* counter, const int32_t* item_ptr,
                               int n_items, void* stream) {
  int hp, ns, sx;
  if (int rc = check_x_geom(H, Fp, &hp, &ns, &sx)) return rc;
  XArgs a = {};
  if (int rc = fill_xargs(a, n_src, n_dst, rowptr, col, H, Fp, sx, xg, ldxg, f, ldf, alpha, seg_len, hub_rows, hub_seg_ptr,
                          n_hub, n_hub_seg, hub_scratch, counter, item_ptr, n_items))
    return rc;
  GATK_REQUIRE(lse && xagg && dxagg && (ds || dg_acc) && df, "null pointer argument (ds and dg_acc cannot both be NULL)");
  GATK_REQUIRE(ldxa % 4 == 0 && ldxa >= (int64_t)H * Fp && ldd % 4 == 0 && ldd >= (int64_t)H * Fp && lddf >= H &&
                   ((uintptr_t)xagg & 15) == 0 && ((uintptr_t)dxagg & 15) == 0 && ((uintptr_t)ds & 15) == 0,
               "xagg / dxagg: 16-byte aligned, pitch >= H*Fp (multiple of 4 floats); ds 16-byte aligned; lddf >= H");
  GATK_REQUIRE(!dg_acc || lddg_acc >= H, "lddg_acc must be >= H");
  a.xagg = const_cast<float*>(xagg); a.ldxa = ldxa; a.lse = const_cast<float*>(lse);
  a.dxagg = dxagg; a.ldd = ldd; a.ds = ds; a.iperm = iperm; a.df = df; a.lddf = lddf;
  a.dgacc = dg_acc; a.lddgacc = lddg_acc;
  a.dg_vec4 = (dg_acc && H == 8 && lddg_acc % 4 == 0 && ((uintptr_t)dg_acc & 15) == 0) ? 1 : 0;
  static const bool dg_plain = getenv("GATK_DG_NO_POLICY") != nullptr;  // measured: evict_last on the reds is worth ~0.1 ms
  if (a.dg_vec4 && !dg_plain) a.dg_vec4 = 2;
  cudaStream_t st = (cudaStream_t)stream;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  // (A stream access-policy window that pins the dg accumulation array in the persisting part of L2 was measured: the
  // edge pass itself went 10.6 -> 8.8 ms, but the device-wide set-aside it needs took the L2 away from every other
  // kernel of the step -- 27.9 -> 38.5 ms per step -- so it is not used; an evict_first hint on the streamed row-state
  // copies changed nothing measurable either.  What remains is the evict_last hint on the reductions themselves.)
  static const bool simt_only = getenv("GATK_XBWD_SIMT") != nullptr;
  if (Fp <= 128 && !simt_only) return launch_x_bwd_mma(a, st);
  HP_DISPATCH(hp, return (launch_x_bwd<HP>(a, st)));
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

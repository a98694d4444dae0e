// Fused attention of the reference's GATv2 flavour, SpGraphAttentionLayerV2 (layers.py:255-313):
//
//     u_ij = Whi_i + Whj_j                     (two projections of the input, layers.py:265-275)
//     s_ij = a . LeakyReLU(u_ij)               (a D-wide operation per stored entry: it does NOT split into f_i + g_j)
//     alpha_ij = softmax_j(s_ij)               (scatter_max + exp + rowsum, layers.py:280-291)
//     h'_i = sum_j alpha~_ij Whi_j  (+ skip_i) (the FIRST projection of the source is aggregated, layers.py:295-301)
//
// Same skeleton as K2 (attn_fwd.cu): one warp owns a destination row, lanes own float4 slots of the H*Dp-wide
// rows, every stored entry gathers the source's [Whi_j | Whj_j] pair (adjacent column blocks of one projection
// output, so the pair is two runs of coalesced 512-byte warp loads), the logit is a per-head reduction over the
// slots, the softmax is online (running max / sum per head).  The backward is one destination-major pass as well:
// it recomputes u_ij and alpha_ij from the saved log-sum-exp, accumulates the destination-side gradient dWhi_i and
// da in registers, and sends the source-side gradients (alpha~_ij dh'_i into dWhi_j, ds_ij a LeakyReLU'(u_ij) into
// dWhj_j) with vector reductions (red.global.add.v4.f32) into the zero-initialised dZ.
#include "attn_common.cuh"

namespace gatk {

struct V2Args {
  int64_t n_dst;
  const int64_t* rowptr;
  const int32_t* col;
  int H, lph, V;       // heads, float4 slots per head, slots per row (H * Dp / 4)
  const float* z;      // [n_src, ldz]: Whi | Whj | (skip)
  int64_t ldz;
  const float* a;      // [H, Dp]
  const uint8_t* keep; // [E, H] attention-dropout keep mask or NULL
  float inv_keep, alpha;
  int has_skip, act_elu;
  float* hagg;         // [n_dst, H*Dp] pre-skip, pre-ELU aggregation (saved for backward) or NULL
  float* out;
  int64_t ldo;
  float* lse;          // [n_dst, H]
  // backward
  const float* gout;
  int64_t ldgo;
  float* dz;           // [n_src, ldz], zero-initialised
  float* da;           // [H, Dp], zero-initialised
  int32_t* counter;
};

__device__ __forceinline__ float lrelu(float x, float alpha) { return x > 0.f ? x : alpha * x; }
__device__ __forceinline__ float elu1v(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ void red4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int V2_WARPS = 4;

template <int NV>
__global__ void __launch_bounds__(V2_WARPS * 32) attn_v2_fwd_kernel(const V2Args a) {
  const int lane = threadIdx.x & 31;
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  const int HD4 = a.V * 4;
  float4 av[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) av[v] = geo.act[v] ? ldg4(a.a + (lane + 32 * v) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);

  int cur = warp_grab(a.counter, lane, GRAB);
  while (cur < a.n_dst) {
    const int nxt = warp_grab(a.counter, lane, GRAB);
    const int rend = cur + GRAB < a.n_dst ? cur + GRAB : (int)a.n_dst;
    for (int row = cur; row < rend; ++row) {
      const int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
      float4 wi[NV], acc[NV];
      float m[NV], l[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        wi[v] = geo.act[v] ? ldg4(a.z + (int64_t)row * a.ldz + (lane + 32 * v) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        m[v] = -INFINITY;
        l[v] = 0.f;
      }
      for (int64_t e = beg; e < end; ++e) {
        const int j = __ldg(a.col + e);
        const float* zj = a.z + (int64_t)j * a.ldz + lane * 4;
        float4 whi[NV];
        float part[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          part[v] = 0.f;
          if (geo.act[v]) {
            whi[v] = ldg4(zj + v * 128);
            const float4 whj = ldg4(zj + HD4 + v * 128);
            part[v] = av[v].x * lrelu(wi[v].x + whj.x, a.alpha) + av[v].y * lrelu(wi[v].y + whj.y, a.alpha) +
                      av[v].z * lrelu(wi[v].z + whj.z, a.alpha) + av[v].w * lrelu(wi[v].w + whj.w, a.alpha);
          }
        }
        head_reduce<NV>(part, a.lph);  // every lane: the logit of each of its slots' heads
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (geo.act[v]) {
            const float s = part[v];
            const float m_new = fmaxf(m[v], s);
            const float sc = (m[v] == -INFINITY) ? 0.f : expf(m[v] - m_new);
            float p = expf(s - m_new);
            l[v] = l[v] * sc + p;   // the row sum is taken BEFORE the attention dropout (layers.py:285-288)
            m[v] = m_new;
            if (a.keep) p = a.keep[e * a.H + geo.hv[v]] ? p * a.inv_keep : 0.f;
            scale4(acc[v], sc);
            fma4(acc[v], p, whi[v]);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (geo.act[v]) {
          const int slot = lane + 32 * v;
          float4 r = acc[v];
          if (l[v] > 0.f) {
            const float inv = 1.f / l[v];
            scale4(r, inv);
          } else {
            r = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (a.hagg) stg4(a.hagg + (int64_t)row * HD4 + slot * 4, r);
          if (a.has_skip) {
            const float4 s = ldg4(a.z + (int64_t)row * a.ldz + 2 * HD4 + slot * 4);
            r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
          }
          if (a.act_elu) {
            r.x = elu1v(r.x); r.y = elu1v(r.y); r.z = elu1v(r.z); r.w = elu1v(r.w);
          }
          stg4(a.out + (int64_t)row * a.ldo + slot * 4, r);
          if (a.lse && geo.leader[v]) a.lse[(int64_t)row * a.H + geo.hv[v]] = l[v] > 0.f ? m[v] + logf(l[v]) : 0.f;
        }
      }
    }
    cur = nxt;
  }
}

template <int NV>
__global__ void __launch_bounds__(V2_WARPS * 32) attn_v2_bwd_kernel(const V2Args a) {
  const int lane = threadIdx.x & 31;
  LaneGeom<NV> geo;
  geo.init(lane, a.lph, a.V);
  const int HD4 = a.V * 4;
  float4 av[NV], da_acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    av[v] = geo.act[v] ? ldg4(a.a + (lane + 32 * v) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    da_acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int cur = warp_grab(a.counter, lane, GRAB);
  while (cur < a.n_dst) {
    const int nxt = warp_grab(a.counter, lane, GRAB);
    const int rend = cur + GRAB < a.n_dst ? cur + GRAB : (int)a.n_dst;
    for (int row = cur; row < rend; ++row) {
      const int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
      float4 wi[NV], dh[NV], dwi[NV];
      float lsev[NV], cpart[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int slot = lane + 32 * v;
        wi[v] = dh[v] = dwi[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        lsev[v] = 0.f;
        cpart[v] = 0.f;
        if (geo.act[v]) {
          wi[v] = ldg4(a.z + (int64_t)row * a.ldz + slot * 4);
          float4 go = ldg4_stream(a.gout + (int64_t)row * a.ldgo + slot * 4);
          if (a.act_elu) {  // ELU'(h') from the activated output
            const float4 o = ldg4_stream(a.out + (int64_t)row * a.ldo + slot * 4);
            go.x *= o.x > 0.f ? 1.f : o.x + 1.f;
            go.y *= o.y > 0.f ? 1.f : o.y + 1.f;
            go.z *= o.z > 0.f ? 1.f : o.z + 1.f;
            go.w *= o.w > 0.f ? 1.f : o.w + 1.f;
          }
          dh[v] = go;
          if (a.has_skip) stg4(a.dz + (int64_t)row * a.ldz + 2 * HD4 + slot * 4, go);  // dSkip_i = dh'_i
          cpart[v] = dot4(go, ldg4_stream(a.hagg + (int64_t)row * HD4 + slot * 4));
          lsev[v] = __ldg(a.lse + (int64_t)row * a.H + geo.hv[v]);
        }
      }
      head_reduce<NV>(cpart, a.lph);  // c_ih = dh'_ih . hagg_ih
      for (int64_t e = beg; e < end; ++e) {
        const int j = __ldg(a.col + e);
        const float* zj = a.z + (int64_t)j * a.ldz + lane * 4;
        float4 whi[NV], u[NV];
        float s[NV], dal[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          s[v] = dal[v] = 0.f;
          if (geo.act[v]) {
            whi[v] = ldg4(zj + v * 128);
            const float4 whj = ldg4(zj + HD4 + v * 128);
            u[v] = make_float4(wi[v].x + whj.x, wi[v].y + whj.y, wi[v].z + whj.z, wi[v].w + whj.w);
            s[v] = av[v].x * lrelu(u[v].x, a.alpha) + av[v].y * lrelu(u[v].y, a.alpha) + av[v].z * lrelu(u[v].z, a.alpha) +
                   av[v].w * lrelu(u[v].w, a.alpha);
            dal[v] = dot4(dh[v], whi[v]);
          }
        }
        head_reduce<NV>(s, a.lph);
        head_reduce<NV>(dal, a.lph);
        float* dzj = a.dz + (int64_t)j * a.ldz + lane * 4;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (geo.act[v]) {
            const float al = expf(s[v] - lsev[v]);
            const float kv = a.keep ? (a.keep[e * a.H + geo.hv[v]] ? a.inv_keep : 0.f) : 1.f;
            const float ds = al * (kv * dal[v] - cpart[v]);
            // value path: dWhi_j += alpha~_ij dh'_i
            const float w = al * kv;
            red4(dzj + v * 128, make_float4(w * dh[v].x, w * dh[v].y, w * dh[v].z, w * dh[v].w));
            // logit path: du_ij = ds_ij a LeakyReLU'(u_ij) goes to dWhi_i (this row) and dWhj_j
            float4 du;
            du.x = ds * av[v].x * (u[v].x > 0.f ? 1.f : a.alpha);
            du.y = ds * av[v].y * (u[v].y > 0.f ? 1.f : a.alpha);
            du.z = ds * av[v].z * (u[v].z > 0.f ? 1.f : a.alpha);
            du.w = ds * av[v].w * (u[v].w > 0.f ? 1.f : a.alpha);
            red4(dzj + HD4 + v * 128, du);
            dwi[v].x += du.x; dwi[v].y += du.y; dwi[v].z += du.z; dwi[v].w += du.w;
            da_acc[v].x += ds * lrelu(u[v].x, a.alpha);
            da_acc[v].y += ds * lrelu(u[v].y, a.alpha);
            da_acc[v].z += ds * lrelu(u[v].z, a.alpha);
            da_acc[v].w += ds * lrelu(u[v].w, a.alpha);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (geo.act[v]) red4(a.dz + (int64_t)row * a.ldz + (lane + 32 * v) * 4, dwi[v]);
    }
    cur = nxt;
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
    if (geo.act[v]) red4(a.da + (lane + 32 * v) * 4, da_acc[v]);
}

template <int NV>
static int launch_v2(const V2Args& a, bool backward, cudaStream_t st) {
  int grid = 0;
  if (backward) {
    if (int rc = persistent_grid(attn_v2_bwd_kernel<NV>, V2_WARPS * 32, 0, &grid)) return rc;
  } else {
    if (int rc = persistent_grid(attn_v2_fwd_kernel<NV>, V2_WARPS * 32, 0, &grid)) return rc;
  }
  const int64_t need = (a.n_dst + (int64_t)V2_WARPS * GRAB - 1) / ((int64_t)V2_WARPS * GRAB);
  if (need < grid) grid = (int)need;
  if (backward)
    attn_v2_bwd_kernel<NV><<<grid, V2_WARPS * 32, 0, st>>>(a);
  else
    attn_v2_fwd_kernel<NV><<<grid, V2_WARPS * 32, 0, st>>>(a);
  GATK_CHECK_LAUNCH();
  return 0;
}

static int fill_v2(V2Args& a, int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp, const float* z,
                   int64_t ldz, const float* avec, const uint8_t* keep, float inv_keep, float alpha, int has_skip,
                   int act_elu, int32_t* counter, int* nv) {
  if (int rc = check_geom(H, Dp, nv)) return rc;
  GATK_REQUIRE(*nv <= 8, "GATv2 kernels: H*Dp=%d too wide (max 1024 floats per row)", H * Dp);
  GATK_REQUIRE(n_dst >= 0 && n_dst < (1LL << 31), "n_dst out of range");
  GATK_REQUIRE(rowptr && col && z && avec && counter, "null pointer argument");
  GATK_REQUIRE(ldz % 4 == 0 && ldz >= (int64_t)(has_skip ? 3 : 2) * H * Dp && ((uintptr_t)z & 15) == 0 && ((uintptr_t)avec & 15) == 0,
               "z must be 16-byte aligned with pitch >= (2 or 3)*H*Dp, a multiple of 4 floats");
  a.n_dst = n_dst; a.rowptr = rowptr; a.col = col; a.H = H; a.lph = Dp / 4; a.V = H * (Dp / 4); a.z = z; a.ldz = ldz;
  a.a = avec; a.keep = keep; a.inv_keep = inv_keep; a.alpha = alpha; a.has_skip = has_skip; a.act_elu = act_elu;
  a.counter = counter;
  return 0;
}

}  // namespace gatk

using namespace gatk;

extern "C" int gatk_attn_v2_fwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp, const float* z,
                                int64_t ldz, const float* a, const uint8_t* keep_att, float inv_keep, float alpha,
                                int has_skip, int act_elu, float* hagg, float* out, int64_t ldo, float* lse,
                                int32_t* counter, void* stream) {
  V2Args v = {};
  int nv;
  if (int rc = fill_v2(v, n_dst, rowptr, col, H, Dp, z, ldz, a, keep_att, inv_keep, alpha, has_skip, act_elu, counter, &nv)) return rc;
  GATK_REQUIRE(out && ldo % 4 == 0 && ldo >= (int64_t)H * Dp && ((uintptr_t)out & 15) == 0, "out: 16-byte aligned, pitch >= H*Dp");
  v.hagg = hagg; v.out = out; v.ldo = ldo; v.lse = lse;
  if (n_dst == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  switch (nv) {
    case 1: return launch_v2<1>(v, false, st);
    case 2: return launch_v2<2>(v, false, st);
    case 4: return launch_v2<4>(v, false, st);
    default: return launch_v2<8>(v, false, st);
  }
}

extern "C" int gatk_attn_v2_bwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp, const float* z,
                                int64_t ldz, const float* a, const uint8_t* keep_att, float inv_keep, float alpha,
                                int has_skip, int act_elu, const float* hagg, const float* out, int64_t ldo,
                                const float* lse, const float* gout, int64_t ldgo, float* dz, float* da, int32_t* counter,
                                void* stream) {
  V2Args v = {};
  int nv;
  if (int rc = fill_v2(v, n_dst, rowptr, col, H, Dp, z, ldz, a, keep_att, inv_keep, alpha, has_skip, act_elu, counter, &nv)) return rc;
  GATK_REQUIRE(hagg && lse && gout && dz && da && (!act_elu || out), "null pointer argument");
  GATK_REQUIRE(ldgo % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)dz & 15) == 0 && ((uintptr_t)da & 15) == 0 && ((uintptr_t)gout & 15) == 0,
               "gout / dz / da must be 16-byte aligned, pitches multiples of 4 floats");
  v.hagg = const_cast<float*>(hagg); v.out = const_cast<float*>(out); v.ldo = ldo; v.lse = const_cast<float*>(lse);
  v.gout = gout; v.ldgo = ldgo; v.dz = dz; v.da = da;
  if (n_dst == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GATK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  switch (nv) {
    case 1: return launch_v2<1>(v, true, st);
    case 2: return launch_v2<2>(v, true, st);
    case 4: return launch_v2<4>(v, true, st);
    default: return launch_v2<8>(v, true, st);
  }
}

// K1: projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
//   C[M,N] = A[M,K] * B[K,N]     A, C row-major fp32 (rows = nodes), B = packed per-head weights.
//
// torch.mm(h, W) (layers.py:35,134) is the one dense contraction on the path.  fp32 parity
// (rel 1e-5) through kind::tf32 needs an error-compensated split:  x = hi + lo with hi the
// tf32 truncation, and  A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo  (fp32 accumulation in TMEM; the
// dropped lo*lo term is 2^-22 relative).  B (small) is split and transposed to K-major once
// by a prep kernel; A (large, streamed once) is split on the fly in shared memory by a
// dedicated warpgroup between the TMA load and the MMA issue.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over 128x128 output tiles):
//   warp 0      TMA producer: A tile (fp32), Bhi, Blo tiles -> 3-stage smem ring (128B swizzle)
//   warp 1      MMA issuer (one elected lane): 12 x tcgen05.mma.kind::tf32 per 32-wide k-block
//   warp 2      TMEM allocator (512 columns: 2 stages x {hi*hi, compensation} accumulators)
//   warps 4-7   splitter: A tile -> (Ahi in place, Alo) in smem, generic->async proxy fence
//   warps 8-11  epilogue: tcgen05.ld -> swizzled smem staging -> TMA store (clips M/N tails)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace gatk {

namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 128, BLOCK_K = 32;  // BLOCK_K fp32 = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BLOCK_M * BLOCK_K * 4;          // 16 KiB, same for A and B tiles
constexpr int STAGE_BYTES = 4 * TILE_BYTES;                // Ahi, Alo, Bhi, Blo
constexpr int CSTAGE_BYTES = BLOCK_M * 32 * 4;             // one 128 x 32 fp32 store chunk
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * CSTAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 4 * BLOCK_N;                    // 2 stages x (main hi*hi, correction lo*hi + hi*lo)
constexpr uint32_t SPIN_LIMIT = 1u << 28;                  // a protocol bug traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused with swizzle), canonical 1
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_c,
                   int m_tiles, int n_tiles, int k_blocks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* cstage = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(cstage + 2 * CSTAGE_BYTES);
  // barrier slots: full[S], split[S], empty[S], tmem_full[2], tmem_empty[2]; then the TMEM base address
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto split_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Ring r;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(r.stage), r.phase ^ 1);
          const uint32_t st = smem_base + r.stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(r.stage), 3 * TILE_BYTES);
          tma_load_2d(st, &map_a, full_bar(r.stage), kb * BLOCK_K, m0);
          tma_load_2d(st + 2 * TILE_BYTES, &map_bhi, full_bar(r.stage), kb * BLOCK_K, n0);
          tma_load_2d(st + 3 * TILE_BYTES, &map_blo, full_bar(r.stage), kb * BLOCK_K, n0);
          r.advance();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    Ring r;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      // The tensor core truncates when it adds into the fp32 accumulator, a bias that grows with the
      // number of accumulations.  hi*hi goes to its own accumulator (K/8 additions); the two small
      // compensation products go to a second one whose truncation error is 2^-10 smaller.
      const uint32_t tmem_d = tmem_base + acc * 2 * BLOCK_N;
      const uint32_t tmem_c = tmem_d + BLOCK_N;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(full_bar(r.stage), r.phase);
        mbar_wait(split_bar(r.stage), r.phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_base + r.stage * STAGE_BYTES;
          const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + TILE_BYTES);
          const uint64_t b_hi = umma_desc(st + 2 * TILE_BYTES), b_lo = umma_desc(st + 3 * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);  // 8 tf32 = 32 bytes along K inside the swizzle row
            umma_tf32(tmem_c, a_lo + adv, b_hi + adv, (kb | k) != 0);
            umma_tf32(tmem_c, a_hi + adv, b_lo + adv, 1);
            umma_tf32(tmem_d, a_hi + adv, b_hi + adv, (kb | k) != 0);
          }
          umma_commit(empty_bar(r.stage));
          if (kb == k_blocks - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        r.advance();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ splitter: A -> (hi, lo)
    Ring r;
    const int t = threadIdx.x - 128;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(full_bar(r.stage), r.phase);
        float4* hi = reinterpret_cast<float4*>(smem + r.stage * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + r.stage * STAGE_BYTES + TILE_BYTES);
#pragma unroll
        for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
          const int idx = t + i * 128;
          float4 v = hi[idx];
          float4 h;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          hi[idx] = h;
          lo[idx] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(r.stage));
        r.advance();
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 8;            // == warp % 4: the TMEM lane quarter this warp may read
    const int row = ew * 32 + lane;     // row inside the tile
    const bool issuer = threadIdx.x == 256;
    int it = 0, chunk_id = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c, ++chunk_id) {
        uint32_t v[32], w[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 2 * BLOCK_N + c * 32;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + BLOCK_N, w);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        uint8_t* buf = cstage + (chunk_id & 1) * CSTAGE_BYTES;
        if (issuer) tma_store_wait_read<1>();  // the store that last read this buffer has drained
        named_bar_sync(1, 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 q = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          *reinterpret_cast<uint4*>(buf + row * 128 + ((j ^ (row & 7)) << 4)) = q;
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (issuer) {
          tma_store_2d(&map_c, smem_u32(buf), n0 + c * 32, m0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// B[K,N] row-major -> K-major hi / lo copies Bt[Npad, Kpad], zero padded.
__global__ void split_transpose_b_kernel(const float* __restrict__ B, int64_t ldb, int K, int N, int Kpad, int Npad,
                                         int b_is_nk, float* __restrict__ bhi, float* __restrict__ blo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Npad * Kpad) return;
  const int n = (int)(i / Kpad), k = (int)(i % Kpad);
  float v = (n < N && k < K) ? (b_is_nk ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n]) : 0.f;
  float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  bhi[i] = h;
  blo[i] = v - h;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D fp32 row-major tensor [rows, cols] with row pitch ld floats; box = 32 columns x 128 rows, 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  GATK_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GATK_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)rc,
               (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

}  // namespace tc

// =====================================================================================================
// K5: dW[Mo,No] = A^T B with A = x [K, Mo] and B = dZ [K, No], both row-major, K = number of nodes.
//
// Both operands are "MN-major" for the tensor core (the reduction index is the slow one in memory), so
// tiles are loaded as 32-float x 32-row TMA boxes and described to tcgen05.mma with MN-major
// descriptors.  32-bit MN-major operands have exactly one legal layout: 128-byte rows swizzled in
// 32-byte units (TMA SWIZZLE_128B_ATOM_32B / UMMA SWIZZLE_128B_BASE32B), i.e. a 4-row (K) x 32-float
// (MN) atom of 512 B; atoms along K are 512 B apart inside a box (stride byte offset), the four
// 32-float blocks of a 128-wide tile 4 KiB apart (leading byte offset).  Both operands are large, so both are split hi/lo on the fly.
// The reduction is split over CTAs (deterministic partial tiles + a reduce kernel), and because the
// tensor core truncates when it accumulates, every PROMOTE k-blocks the TMEM accumulators are drained
// into fp32 registers of the epilogue warps (round-to-nearest adds) -- the MMA warp meanwhile
// continues into the other TMEM buffer.
//
// Warp roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 splitter (A and B tiles), warps 8-15 epilogue (lane quarter = warp % 4, column half =
// (warp - 8) / 4).
// =====================================================================================================
namespace tn {

using namespace tc;

constexpr int TN_THREADS = 512;
constexpr int PROMOTE = 16;               // k-blocks (of 32 rows) between accumulator drains
constexpr int BOX_BYTES = 32 * 32 * 4;    // one TMA box: 32 floats x 32 rows
constexpr int TN_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(BOX_BYTES >> 4) << 16;   // leading byte offset: next 32-float block along M/N
  d |= (uint64_t)(512 >> 4) << 32;         // stride byte offset: next 4-row swizzle atom along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B: the only MN-major layout for 32-bit operands
  return d;
}
// kind::tf32, fp32 accumulate, A and B MN-major (bits 15, 16), M = 128, N = 128
constexpr uint32_t IDESC_MN = IDESC | (1u << 15) | (1u << 16);

__device__ __forceinline__ void umma_tf32_mn(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC_MN), "r"(accumulate)
      : "memory");
}

// KMAJOR = false: the TN product above (both operands MN-major, both split on the fly, split-K partials).
// KMAJOR = true : C[Mo,No] (+)= A[Mo,K] * W with A row-major (K-major) split on the fly and the weights
//                 pre-split / pre-transposed to K-major (map_b = hi, map_b2 = lo): the long-K version of the
//                 projection kernel (any K, drains every PROMOTE k-blocks), writing C directly.
template <bool KMAJOR>
__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_promoted_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                            const __grid_constant__ CUtensorMap map_b2, float* __restrict__ part, int64_t ld_out,
                            int accumulate, int Mo, int No, int m_tiles, int n_tiles, int splits, int kb_total,
                            int kb_per_split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto split_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = m_tiles * n_tiles;
  const int items = tiles * splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  // item -> (split, m tile, n tile); consecutive CTAs share the k range (A tiles hit in L2)
  auto decode = [&](int item, int& m0, int& n0, int& kb0, int& kb1, int& sp) {
    sp = item / tiles;
    const int t = item - sp * tiles;
    m0 = (t / n_tiles) * BLOCK_M;
    n0 = (t % n_tiles) * BLOCK_N;
    kb0 = sp * kb_per_split;
    kb1 = kb0 + kb_per_split < kb_total ? kb0 + kb_per_split : kb_total;
  };

  if (warp == 0) {
    if (lane == 0) {
      Ring r;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int m0, n0, kb0, kb1, sp;
        decode(item, m0, n0, kb0, kb1, sp);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(r.stage), r.phase ^ 1);
          const uint32_t st = smem_base + r.stage * STAGE_BYTES;
          if constexpr (KMAJOR) {
            mbar_expect_tx(full_bar(r.stage), 3 * TILE_BYTES);
            tma_load_2d(st, &map_a, full_bar(r.stage), kb * BLOCK_K, m0);
            tma_load_2d(st + 2 * TILE_BYTES, &map_b, full_bar(r.stage), kb * BLOCK_K, n0);
            tma_load_2d(st + 3 * TILE_BYTES, &map_b2, full_bar(r.stage), kb * BLOCK_K, n0);
          } else {
            mbar_expect_tx(full_bar(r.stage), 2 * TILE_BYTES);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              tma_load_2d(st + b * BOX_BYTES, &map_a, full_bar(r.stage), m0 + b * 32, kb * BLOCK_K);
              tma_load_2d(st + 2 * TILE_BYTES + b * BOX_BYTES, &map_b, full_bar(r.stage), n0 + b * 32, kb * BLOCK_K);
            }
          }
          r.advance();
        }
      }
    }
  } else if (warp == 1) {
    Ring r;
    int drain = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int m0, n0, kb0, kb1, sp;
      decode(item, m0, n0, kb0, kb1, sp);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int rel = kb - kb0;
        const int acc = drain & 1;
        if (rel % PROMOTE == 0) {
          mbar_wait(tempty_bar(acc), ((drain >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t tmem_d = tmem_base + acc * 2 * BLOCK_N;
        const uint32_t tmem_c = tmem_d + BLOCK_N;
        mbar_wait(full_bar(r.stage), r.phase);
        mbar_wait(split_bar(r.stage), r.phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_base + r.stage * STAGE_BYTES;
          const uint32_t fresh = (rel % PROMOTE) == 0 ? 0u : 1u;
          if constexpr (KMAJOR) {
            const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + TILE_BYTES);
            const uint64_t b_hi = umma_desc(st + 2 * TILE_BYTES), b_lo = umma_desc(st + 3 * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);  // 8 tf32 = 32 bytes along K inside the swizzle row
              umma_tf32(tmem_c, a_lo + adv, b_hi + adv, fresh | (uint32_t)(k != 0));
              umma_tf32(tmem_c, a_hi + adv, b_lo + adv, 1);
              umma_tf32(tmem_d, a_hi + adv, b_hi + adv, fresh | (uint32_t)(k != 0));
            }
          } else {
            const uint64_t a_hi = umma_desc_mn(st), a_lo = umma_desc_mn(st + TILE_BYTES);
            const uint64_t b_hi = umma_desc_mn(st + 2 * TILE_BYTES), b_lo = umma_desc_mn(st + 3 * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 8; ++k) {
              const uint64_t adv = (uint64_t)(k * 1024 >> 4);  // next 8 rows along K
              umma_tf32_mn(tmem_c, a_lo + adv, b_hi + adv, fresh | (uint32_t)(k != 0));
              umma_tf32_mn(tmem_c, a_hi + adv, b_lo + adv, 1);
              umma_tf32_mn(tmem_d, a_hi + adv, b_hi + adv, fresh | (uint32_t)(k != 0));
            }
          }
          umma_commit(empty_bar(r.stage));
          if ((rel + 1) % PROMOTE == 0 || kb == kb1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if ((rel + 1) % PROMOTE == 0 || kb == kb1 - 1) ++drain;
        r.advance();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    Ring r;
    const int t = threadIdx.x - 128;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int m0, n0, kb0, kb1, sp;
      decode(item, m0, n0, kb0, kb1, sp);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(r.stage), r.phase);
        uint8_t* st = smem + r.stage * STAGE_BYTES;
#pragma unroll
        for (int op = 0; op < (KMAJOR ? 1 : 2); ++op) {  // A, then B unless the weights came pre-split
          float4* hi = reinterpret_cast<float4*>(st + op * 2 * TILE_BYTES);
          float4* lo = reinterpret_cast<float4*>(st + op * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
          for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
            const int idx = t + i * 128;
            float4 v = hi[idx];
            float4 h;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            hi[idx] = h;
            lo[idx] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(r.stage));
        r.advance();
      }
    }
  } else if (warp >= 8) {
    const int quarter = warp & 3;             // TMEM lanes [32*quarter, +32)
    const int half = (warp - 8) >> 2;         // columns [64*half, +64) of the tile
    const int row = quarter * 32 + lane;
    int drain = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int m0, n0, kb0, kb1, sp;
      decode(item, m0, n0, kb0, kb1, sp);
      float racc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) racc[j] = 0.f;
      const int n_drains = (kb1 - kb0 + PROMOTE - 1) / PROMOTE;
      for (int d = 0; d < n_drains; ++d, ++drain) {
        const int acc = drain & 1;
        mbar_wait(tfull_bar(acc), (drain >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 2 * BLOCK_N + half * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32], w[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld32(taddr + BLOCK_N + c * 32, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[c * 32 + j] += __uint_as_float(v[j]) + __uint_as_float(w[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      const int gm = m0 + row;
      if (gm < Mo) {
        if constexpr (KMAJOR) {  // final values straight into C (row pitch ld_out), optional accumulate
          float* dst = part + (int64_t)gm * ld_out + n0 + half * 64;
#pragma unroll
          for (int j = 0; j < 64; ++j)
            if (n0 + half * 64 + j < No) dst[j] = accumulate ? dst[j] + racc[j] : racc[j];
        } else {                 // partial tile of this k range -> workspace [split][Mo][No]
          float* dst = part + ((int64_t)sp * Mo + gm) * No + n0 + half * 64;
#pragma unroll
          for (int j = 0; j < 64; ++j)
            if (n0 + half * 64 + j < No) dst[j] = racc[j];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// 2-D fp32 row-major tensor [rows, cols]; box = 32 columns x 32 rows, 128B swizzle.
static int make_map_box32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  GATK_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GATK_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)rc,
               (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

}  // namespace tn


// =====================================================================================================
// Batched (per-head) products for the aggregate-first form of the layer (functional.GatLayerAggFirstFunction):
//   NN / NT   C_b[M,N] = A_b[M,K] op(B_b)        out_h = xagg_h W_h,  dxagg_h = dh'_h W_h^T   (K <= 512)
//   TN        C_b[Mo,No] = A_b^T[Mo,K] B_b[K,No]  dW_h = xagg_h^T dh'_h                         (K = nodes)
// with A_b = A + b*a_bs (a COLUMN block of the same rows), C_b = C + b*c_bs.  One launch covers all heads:
// the operands are described to TMA as 3-D tensors (columns of a batch, rows, batch), so K / N tails of a
// batch are zero-filled on load and clipped on store even though the next head's columns follow in memory.
// Same 3xTF32 scheme as the kernels above; the tile width is 64 when a head is 64 wide.  Differences (all
// ncu-driven): the A operand is split into TENSOR memory (tcgen05.st; the TN kernel transposes
// it through registers on the way) and read from there by the MMAs; Bhi|Blo are one stacked operand (2 MMAs per
// k-step); role branches are warp-uniform with one elect.sync per issue block; A, B and the TMEM A slots cycle
// in separate rings; the wide tile leaves through one dense TMA store.
// =====================================================================================================
namespace bt {

using namespace tc;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_i(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand read from tensor memory (lane = row of the tile, one fp32 column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// F.elu computes exp(x) - 1 (not expm1); the fast exponential is accurate to ~2 ulp of exp, i.e. ~1e-7 absolute here
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

template <int NST>
struct RingN {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == NST) {
      stage = 0;
      phase ^= 1;
    }
  }
};

struct RingRT {  // depth chosen at run time (the fused dh' = gout * ELU'(out) mode pairs the raw A slots)
  int n;
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ explicit RingRT(int depth) : n(depth) {}
  __device__ __forceinline__ void advance() {
    if (++stage == n) {
      stage = 0;
      phase ^= 1;
    }
  }
};
// d/dv ELU(v) from out = ELU(v) (F.elu, layers.py:51,170): 1 for v > 0, exp(v) = out + 1 otherwise
__device__ __forceinline__ float elu_grad_from_out(float o) { return o > 0.f ? 1.f : o + 1.f; }

constexpr int BT_THREADS = 512;  // warps 0-2 producer / MMA / TMEM, 4-7 splitter, 8-15 two epilogue warpgroups

// configuration of the shared-memory-operand TN kernel (only the 128-wide tile still uses it: its accumulators
// fill all 512 TMEM columns, so there is no room for A there)
template <int BN>
struct Cfg {
  static constexpr int TILE_A = BLOCK_M * BLOCK_K * 4;  // 16 KiB
  static constexpr int TILE_B = BN * BLOCK_K * 4;
  static constexpr int STAGE = 2 * TILE_A + 2 * TILE_B;  // Ahi, Alo, Bhi, Blo
  // these products are HBM-bound: the ring depth sets the bytes in flight per SM (one 16 KiB A tile per stage)
  static constexpr int NST = BN == 64 ? 4 : 3;
  static constexpr int SMEM = NST * STAGE + 2 * CSTAGE_BYTES + 1024 + 256;
  static constexpr int SMEM_TN = NST * STAGE + 1024 + 256;
  static constexpr int TMEM = 4 * BN;  // 2 stages x (main, correction)
  static constexpr uint32_t IDESC_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  static constexpr uint32_t IDESC_MN = IDESC_K | (1u << 15) | (1u << 16);
  static constexpr uint32_t IDESC_MN2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24) | (1u << 15) | (1u << 16);
};

// ---- NN / NT, K <= 512: C_b = A_b op(B_b), optional ELU epilogue (F.elu, layers.py:51,170)
//
// ncu on the first version (A split into hi / lo copies in shared memory, both read back by every MMA) showed
// the kernel bound by shared-memory bandwidth, not HBM: per 32-wide k-block the SM moved 168 KiB through
// shared memory (TMA fill 32, splitter 16 read + 32 written, 12 MMAs x (4 KiB A + 2 KiB B), store staging 16)
// = 1344 cycles at 128 B/cycle against the 1356 measured.  Here the A operand goes to TENSOR MEMORY instead:
// splitter thread r reads row r of the raw tile once, splits it in registers and writes hi / lo into lane r
// of TMEM (tcgen05.st); the MMAs read A from TMEM and only the weights from shared memory (~70 KiB per
// k-block), which puts the kernel back under the HBM roof.
//
// TMEM (512 columns): accumulators first, then NST x 64 columns of A (32 hi + 32 lo per stage).
//   BN = 64 : 2 accumulator stages x (hi*hi | compensation) x 64               = 256 columns
//   BN = 128: 2 accumulator stages x 128, all three products in ONE accumulator = 256 columns (the separate
//             compensation accumulator only matters for long reductions; this variant takes K <= 128)
template <int BN>
struct CfgTA {
  static constexpr int TILE_A = BLOCK_M * BLOCK_K * 4;  // 16 KiB, raw fp32
  static constexpr int TILE_B = BN * BLOCK_K * 4;
  // Three rings.  The raw A tiles are the only HBM stream and a slot is free again as soon as the splitter
  // has read it, so that ring is deep (bytes in flight per SM = NSA x 16 KiB: with 4 the kernel sat at
  // 4.6 TB/s, latency bound); the weights come from L2 and their slots are held until the MMAs retire.
  static constexpr int NSA = BN == 64 ? 8 : 6;   // raw A tiles in shared memory
  static constexpr int NSB = BN == 64 ? 3 : 2;   // (Bhi, Blo) pairs in shared memory
  static constexpr int NTA = 4;                  // (Ahi, Alo) in tensor memory
  // store staging: BN = 64: one swizzled 128 x 32 chunk per epilogue warpgroup; BN = 128: ONE dense 128 x box_n tile
  static constexpr int CS_BYTES = BN == 64 ? 2 * CSTAGE_BYTES : BLOCK_M * BN * 4;
  static constexpr bool MERGED = BN == 128;
  static constexpr int ACC_COLS = MERGED ? BN : 2 * BN;  // per accumulator stage
  static constexpr int TMEM_A = 2 * ACC_COLS;            // first A column
  static constexpr int TMEM = 512;
  static constexpr int OFF_B = NSA * TILE_A;
  static constexpr int OFF_C = OFF_B + NSB * 2 * TILE_B;
  static constexpr int OFF_BAR = OFF_C + CS_BYTES;
  static constexpr int NBAR = 2 * NSA + 2 * NSB + 2 * NTA + 4;
  static constexpr int SMEM = OFF_BAR + 8 * NBAR + 16 + 1024;
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  static constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  static_assert(TMEM_A + NTA * 2 * BLOCK_K <= TMEM, "tensor memory budget");
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <int BN>
__global__ void __launch_bounds__(BT_THREADS, 1)
gemm_batched_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                           const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_c,
                           int m_tiles, int n_tiles, int batches, int k_blocks, int npad, int epilogue, int contiguous, int box_n,
                           const __grid_constant__ CUtensorMap map_a2, int fuse_elu_grad) {
  using C = CfgTA<BN>;
  // fuse_elu_grad: the A operand is dh' = A * ELU'(A2) (A = upstream gradient, A2 = the layer's activated output):
  // both raw tiles land in one ring stage (two adjacent 16 KiB slots), the splitter multiplies while it splits
  const int nsa = fuse_elu_grad ? C::NSA / 2 : C::NSA;
  const int a_stride = fuse_elu_grad ? 2 * C::TILE_A : C::TILE_A;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* cstage = smem + C::OFF_C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };                                   // TMA -> splitter
  auto a_empty = [&](int s) { return bar0 + 8u * (C::NSA + s); };                       // splitter -> TMA
  auto b_full = [&](int s) { return bar0 + 8u * (2 * C::NSA + s); };                    // TMA -> MMA
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * C::NSA + C::NSB + s); };          // MMA retired -> TMA
  auto t_full = [&](int s) { return bar0 + 8u * (2 * C::NSA + 2 * C::NSB + s); };       // splitter -> MMA
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * C::NSA + 2 * C::NSB + C::NTA + s); };  // MMA retired -> splitter
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::NSA + 2 * C::NSB + 2 * C::NTA + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::NSA + 2 * C::NSB + 2 * C::NTA + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches stay uniform
  const int lane = threadIdx.x & 31;
  const int per_m = n_tiles * batches;
  const int total_tiles = m_tiles * per_m;
  // tile order: strided (tile = blockIdx.x + i gridDim.x: the heads of a row block run on neighbouring SMs at the
  // same time) or contiguous (one CTA walks the heads of its row blocks one after the other)
  const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_begin = contiguous ? blockIdx.x * per_cta : blockIdx.x;
  const int tile_end = contiguous ? min(total_tiles, tile_begin + per_cta) : total_tiles;
  const int tile_step = contiguous ? 1 : gridDim.x;
  auto decode = [&](int tile, int& m0, int& n0, int& b) {
    const int mt = tile / per_m, r = tile - mt * per_m;
    b = r / n_tiles;
    m0 = mt * BLOCK_M;
    n0 = (r - b * n_tiles) * BN;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 4);
    }
    for (int s = 0; s < C::NSB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < C::NTA; ++s) {
      mbar_init(t_full(s), 4);
      mbar_init(t_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    // ---- A producer: the HBM stream
    if (elect_one()) {
      RingRT r(nsa);
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        int m0, n0, b;
        decode(tile, m0, n0, b);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(a_empty(r.stage), r.phase ^ 1);
          mbar_expect_tx(a_full(r.stage), fuse_elu_grad ? 2 * C::TILE_A : C::TILE_A);
          tma_load_3d(smem_base + r.stage * a_stride, &map_a, a_full(r.stage), kb * BLOCK_K, m0, b);
          if (fuse_elu_grad)
            tma_load_3d(smem_base + r.stage * a_stride + C::TILE_A, &map_a2, a_full(r.stage), kb * BLOCK_K, m0, b);
          r.advance();
        }
      }
    }
  } else if (warp == 3) {
    // ---- B producer: pre-split K-major weights, L2 resident
    if (elect_one()) {
      RingN<C::NSB> r;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        int m0, n0, b;
        decode(tile, m0, n0, b);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(b_empty(r.stage), r.phase ^ 1);
          const uint32_t st = smem_base + C::OFF_B + r.stage * 2 * C::TILE_B;
          mbar_expect_tx(b_full(r.stage), 2 * C::TILE_B);
          tma_load_2d(st, &map_bhi, b_full(r.stage), kb * BLOCK_K, b * npad + n0);
          tma_load_2d(st + C::TILE_B, &map_blo, b_full(r.stage), kb * BLOCK_K, b * npad + n0);
          r.advance();
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    RingN<C::NSB> rb;
    RingN<C::NTA> rt;
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::ACC_COLS;
      const uint32_t tmem_c = C::MERGED ? tmem_d : tmem_d + BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(b_full(rb.stage), rb.phase);  // the weights of this k-block have landed
        mbar_wait(t_full(rt.stage), rt.phase);  // hi / lo of A are in tensor memory
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + C::OFF_B + rb.stage * 2 * C::TILE_B;
          const uint32_t a_hi = tmem_base + C::TMEM_A + rt.stage * 2 * BLOCK_K, a_lo = a_hi + BLOCK_K;
          const uint64_t b_hi = umma_desc(st), b_lo = umma_desc(st + C::TILE_B);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            if (C::MERGED) {
              // small products first: they meet a small accumulator
              umma_tf32_ta(tmem_d, a_lo + k * 8, b_hi + adv, C::IDESC, (kb | k) != 0);
              umma_tf32_ta(tmem_d, a_hi + k * 8, b_lo + adv, C::IDESC, 1);
              umma_tf32_ta(tmem_d, a_hi + k * 8, b_hi + adv, C::IDESC, 1);
            } else {
              // Bhi and Blo tiles are adjacent in shared memory = one 2 BN-row operand: a single MMA forms
              // [Ahi Bhi | Ahi Blo] into the adjacent (main | compensation) accumulator columns
              umma_tf32_ta(tmem_d, a_hi + k * 8, b_hi + adv, C::IDESC2, (kb | k) != 0);
              umma_tf32_ta(tmem_c, a_lo + k * 8, b_hi + adv, C::IDESC, 1);
            }
          }
          umma_commit(t_empty(rt.stage));
          umma_commit(b_empty(rb.stage));
          if (kb == k_blocks - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        rb.advance();
        rt.advance();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- splitter: thread `row` owns row `row` of the tile = TMEM lane `row` (warp w may touch lanes 32 (w % 4) ...)
    RingRT ra(nsa);
    RingN<C::NTA> rt;
    const int row = threadIdx.x - 128;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::TMEM_A;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(a_full(ra.stage), ra.phase);
        const uint8_t* src = smem + ra.stage * a_stride + row * 128;
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)  // 16-byte chunk j of the 128-byte row is stored at j ^ (row & 7)
          v[j] = *reinterpret_cast<const float4*>(src + ((j ^ (row & 7)) << 4));
        if (fuse_elu_grad) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 o = *reinterpret_cast<const float4*>(src + C::TILE_A + ((j ^ (row & 7)) << 4));
            v[j].x *= elu_grad_from_out(o.x);
            v[j].y *= elu_grad_from_out(o.y);
            v[j].z *= elu_grad_from_out(o.z);
            v[j].w *= elu_grad_from_out(o.w);
          }
        }
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t h = __float_as_uint(x[q]) & 0xffffe000u;
            hi[4 * j + q] = h;
            lo[4 * j + q] = __float_as_uint(x[q] - __uint_as_float(h));
          }
        }
        __syncwarp();  // every lane's loads have returned (their values were consumed above)
        if (lane == 0) mbar_arrive(a_empty(ra.stage));
        mbar_wait(t_empty(rt.stage), rt.phase ^ 1);  // the MMAs that read this TMEM slot have retired
        tc_fence_after();
        const uint32_t ta = lane_base + rt.stage * 2 * BLOCK_K;
        tmem_st32(ta, hi);
        tmem_st32(ta + BLOCK_K, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_full(rt.stage));
        ra.advance();
        rt.advance();
      }
    }
  } else if (warp >= 8) {
    if constexpr (BN == 128) {
      // ---- epilogue, wide tile: both warpgroups fill ONE dense [128][box_n] staging tile and a single TMA store
      // writes box_n contiguous floats per row.  (Storing 32-column chunks left 8 partially written sectors per
      // row and head when a head's column block is not sector aligned (100 floats): ncu showed +2.4 GB of DRAM
      // read-fill and the store-bound kernel at 3.7 TB/s.)
      const int ew = warp & 3, grp = (warp - 8) >> 2;
      const int row = ew * 32 + lane;
      const bool issuer = threadIdx.x == 256;
      uint8_t* dst_row = cstage + (size_t)row * box_n * 4;
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
        int m0, n0, b;
        decode(tile, m0, n0, b);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        if (issuer) tma_store_wait_read<0>();  // the previous tile's store has drained the staging tile
        named_bar_sync(1, 256);
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += 2) {
          if (c * 32 >= box_n) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * C::ACC_COLS + c * 32, v);
          if (epilogue == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(elu_f(__uint_as_float(v[j])));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c * 32 + 4 * j < box_n)
              *reinterpret_cast<uint4*>(dst_row + (c * 32 + 4 * j) * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        fence_proxy_async();
        named_bar_sync(1, 256);
        if (issuer) {
          tma_store_3d(&map_c, smem_u32(cstage), n0, m0, b);
          tma_store_commit();
        }
      }
      if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
      // ---- two epilogue warpgroups (warps 8-11, 12-15) take alternate 32-column chunks of every tile
      const int ew = warp & 3;            // the TMEM lane quarter this warp may read
      const int grp = (warp - 8) >> 2;
      const int row = ew * 32 + lane;     // row inside the tile
      const bool issuer = (threadIdx.x & 127) == 0;
      uint8_t* buf = cstage + grp * CSTAGE_BYTES;
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
        int m0, n0, b;
        decode(tile, m0, n0, b);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
  #pragma unroll 1
        for (int c = grp; c < BN / 32; c += 2) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * C::ACC_COLS + c * 32;
          tmem_ld32(taddr, v);
          if (!C::MERGED) {
            uint32_t w[32];
            tmem_ld32(taddr + BN, w);
  #pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
          }
          if (epilogue == 1) {
  #pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(elu_f(__uint_as_float(v[j])));
          }
          if (issuer) tma_store_wait_read<0>();  // this group's previous store has drained its staging buffer
          named_bar_sync(1 + grp, 128);
  #pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 q = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            *reinterpret_cast<uint4*>(buf + row * 128 + ((j ^ (row & 7)) << 4)) = q;
          }
          fence_proxy_async();
          named_bar_sync(1 + grp, 128);
          if (issuer) {
            tma_store_3d(&map_c, smem_u32(buf), n0 + c * 32, m0, b);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM));
  }
}

// ---- TN: C_b[Mo,No] = A_b^T B_b, reduction over the rows (nodes), deterministic split-K partials
template <int BN>
__global__ void __launch_bounds__(tn::TN_THREADS, 1)
gemm_tn_batched_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       float* __restrict__ part, int Mo, int No, int m_tiles, int n_tiles, int batches, int splits,
                       int kb_total, int kb_per_split) {
  using C = Cfg<BN>;
  using tn::BOX_BYTES;
  using tn::PROMOTE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NST * C::STAGE);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto split_bar = [&](int s) { return bar0 + 8u * (C::NST + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * C::NST + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * C::NST + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * C::NST + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::NST + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches stay uniform
  const int lane = threadIdx.x & 31;
  const int tiles = m_tiles * n_tiles;
  const int items = batches * tiles * splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NST; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(split_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  // item -> (batch, split, m tile, n tile); the splits of a batch are adjacent
  auto decode = [&](int item, int& b, int& m0, int& n0, int& kb0, int& kb1, int& sp) {
    b = item / (tiles * splits);
    const int r = item - b * tiles * splits;
    sp = r / tiles;
    const int t = r - sp * tiles;
    m0 = (t / n_tiles) * BLOCK_M;
    n0 = (t % n_tiles) * BN;
    kb0 = sp * kb_per_split;
    kb1 = kb0 + kb_per_split < kb_total ? kb0 + kb_per_split : kb_total;
  };

  if (warp == 0) {
    if (elect_one()) {
      RingN<C::NST> r;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, m0, n0, kb0, kb1, sp;
        decode(item, b, m0, n0, kb0, kb1, sp);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(r.stage), r.phase ^ 1);
          const uint32_t st = smem_base + r.stage * C::STAGE;
          mbar_expect_tx(full_bar(r.stage), C::TILE_A + C::TILE_B);
#pragma unroll
          for (int q = 0; q < 4; ++q) tma_load_3d(st + q * BOX_BYTES, &map_a, full_bar(r.stage), m0 + q * 32, kb * BLOCK_K, b);
#pragma unroll
          for (int q = 0; q < BN / 32; ++q)
            tma_load_3d(st + 2 * C::TILE_A + q * BOX_BYTES, &map_b, full_bar(r.stage), n0 + q * 32, kb * BLOCK_K, b);
          r.advance();
        }
      }
    }
  } else if (warp == 1) {
    RingN<C::NST> r;
    int drain = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, m0, n0, kb0, kb1, sp;
      decode(item, b, m0, n0, kb0, kb1, sp);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int rel = kb - kb0;
        const int acc = drain & 1;
        if (rel % PROMOTE == 0) {
          mbar_wait(tempty_bar(acc), ((drain >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t tmem_d = tmem_base + acc * 2 * BN;
        const uint32_t tmem_c = tmem_d + BN;
        mbar_wait(full_bar(r.stage), r.phase);
        mbar_wait(split_bar(r.stage), r.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + r.stage * C::STAGE;
          const uint32_t fresh = (rel % PROMOTE) == 0 ? 0u : 1u;
          const uint64_t a_hi = tn::umma_desc_mn(st), a_lo = tn::umma_desc_mn(st + C::TILE_A);
          const uint64_t b_hi = tn::umma_desc_mn(st + 2 * C::TILE_A);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 1024 >> 4);
            // the Bhi and Blo boxes are adjacent in shared memory = one 2 BN-wide operand: a single MMA forms
            // [Ahi Bhi | Ahi Blo] into the adjacent (main | compensation) accumulator columns
            umma_tf32_i(tmem_d, a_hi + adv, b_hi + adv, C::IDESC_MN2, fresh | (uint32_t)(k != 0));
            umma_tf32_i(tmem_c, a_lo + adv, b_hi + adv, C::IDESC_MN, 1);
          }
          umma_commit(empty_bar(r.stage));
          if ((rel + 1) % PROMOTE == 0 || kb == kb1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if ((rel + 1) % PROMOTE == 0 || kb == kb1 - 1) ++drain;
        r.advance();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    RingN<C::NST> r;
    const int t = threadIdx.x - 128;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, m0, n0, kb0, kb1, sp;
      decode(item, b, m0, n0, kb0, kb1, sp);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(r.stage), r.phase);
        uint8_t* st = smem + r.stage * C::STAGE;
#pragma unroll
        for (int op = 0; op < 2; ++op) {
          float4* hi = reinterpret_cast<float4*>(st + (op ? 2 * C::TILE_A : 0));
          float4* lo = reinterpret_cast<float4*>(st + (op ? 2 * C::TILE_A + C::TILE_B : C::TILE_A));
          const int n16 = (op ? C::TILE_B : C::TILE_A) / 16 / 128;
#pragma unroll
          for (int i = 0; i < n16; ++i) {
            const int idx = t + i * 128;
            float4 v = hi[idx];
            float4 h;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            hi[idx] = h;
            lo[idx] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(r.stage));
        r.advance();
      }
    }
  } else if (warp >= 8) {
    constexpr int HALF = BN / 2;  // columns per epilogue warp group
    const int quarter = warp & 3;
    const int half = (warp - 8) >> 2;
    const int row = quarter * 32 + lane;
    int drain = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, m0, n0, kb0, kb1, sp;
      decode(item, b, m0, n0, kb0, kb1, sp);
      float racc[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) racc[j] = 0.f;
      const int n_drains = (kb1 - kb0 + PROMOTE - 1) / PROMOTE;
      for (int d = 0; d < n_drains; ++d, ++drain) {
        const int acc = drain & 1;
        mbar_wait(tfull_bar(acc), (drain >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 2 * BN + half * HALF;
#pragma unroll
        for (int c = 0; c < HALF / 32; ++c) {
          uint32_t v[32], w[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld32(taddr + BN + c * 32, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[c * 32 + j] += __uint_as_float(v[j]) + __uint_as_float(w[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      const int gm = m0 + row;
      if (gm < Mo) {
        float* dst = part + (((int64_t)b * splits + sp) * Mo + gm) * No + n0 + half * HALF;
#pragma unroll
        for (int j = 0; j < HALF; ++j)
          if (n0 + half * HALF + j < No) dst[j] = racc[j];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM));
  }
}

// ---- TN, 64-wide tile, A through tensor memory.  The first version split both operands into hi / lo copies in
// shared memory and the MMAs read all four back: 152 KiB of shared-memory traffic per 32-row k-block
// (~1200 cycles at 128 B/cycle, 1427 measured) against ~880 cycles of HBM time.  Here the raw A box
// [32 rows x box_m columns] lands unswizzled; splitter thread m reads COLUMN m (consecutive lanes read
// consecutive words: conflict free), which is row m of A^T, splits it in registers and writes it into TMEM lane
// m.  Only the narrow B operand is split in shared memory.
struct CfgTNA {
  static constexpr int BN = 64;
  static constexpr int SLOT_A = BLOCK_K * BLOCK_M * 4;  // 16 KiB slot for a [32][box_m <= 128] raw box
  static constexpr int TILE_B = BN * BLOCK_K * 4;       // 8 KiB
  static constexpr int NSA = 8;                         // raw A boxes in flight (the HBM stream)
  static constexpr int NSS = 4;                         // (Bhi, Blo) in shared memory + (Ahi, Alo) in tensor memory
  static constexpr int OFF_B = NSA * SLOT_A;
  static constexpr int OFF_BAR = OFF_B + NSS * 2 * TILE_B;
  static constexpr int NBAR = 2 * NSA + 3 * NSS + 4;
  static constexpr int SMEM = OFF_BAR + 8 * NBAR + 16 + 1024;
  static constexpr int TMEM_A = 4 * BN;  // after 2 accumulator stages x (main | compensation)
  static constexpr int TMEM = 512;
  // A from tensor memory (K-major by construction), B MN-major (bit 16)
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24) | (1u << 16);
  static constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24) | (1u << 16);
  static_assert(SMEM <= 232448, "shared memory budget");
};

__global__ void __launch_bounds__(tn::TN_THREADS, 1)
gemm_tn_batched_ta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          float* __restrict__ part, int Mo, int No, int m_tiles, int n_tiles, int batches, int splits,
                          int kb_total, int kb_per_split, int box_m, const __grid_constant__ CUtensorMap map_b2,
                          int fuse_elu_grad) {
  // fuse_elu_grad: the B operand is dh' = B * ELU'(B2); the B2 tile lands in the slot that will hold Blo and the
  // splitter multiplies before it splits in place
  using C = CfgTNA;
  constexpr int BN = C::BN;
  using tn::BOX_BYTES;
  using tn::PROMOTE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };                                // TMA -> splitter
  auto a_empty = [&](int s) { return bar0 + 8u * (C::NSA + s); };                    // splitter -> TMA
  auto b_full = [&](int s) { return bar0 + 8u * (2 * C::NSA + s); };                 // TMA -> splitter
  auto s_ready = [&](int s) { return bar0 + 8u * (2 * C::NSA + C::NSS + s); };       // splitter -> MMA
  auto s_empty = [&](int s) { return bar0 + 8u * (2 * C::NSA + 2 * C::NSS + s); };   // MMA retired -> B producer
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::NSA + 3 * C::NSS + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::NSA + 3 * C::NSS + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NBAR);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int tiles = m_tiles * n_tiles;
  const int items = batches * tiles * splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 4);
    }
    for (int s = 0; s < C::NSS; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(s_ready(s), 4);
      mbar_init(s_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  // item -> (batch, split, m tile, n tile); the splits of a batch are adjacent
  auto decode = [&](int item, int& b, int& m0, int& n0, int& kb0, int& kb1, int& sp) {
    b = item / (tiles * splits);
    const int r = item - b * tiles * splits;
    sp = r / tiles;
    const int t = r - sp * tiles;
    m0 = (t / n_tiles) * BLOCK_M;
    n0 = (t % n_tiles) * BN;
    kb0 = sp * kb_per_split;
    kb1 = kb0 + kb_per_split < kb_total ? kb0 + kb_per_split : kb_total;
  };

  if (warp == 0) {
    if (elect_one()) {
      RingN<C::NSA> r;
      const uint32_t a_bytes = (uint32_t)(BLOCK_K * box_m * 4);
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, m0, n0, kb0, kb1, sp;
        decode(item, b, m0, n0, kb0, kb1, sp);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(a_empty(r.stage), r.phase ^ 1);
          mbar_expect_tx(a_full(r.stage), a_bytes);
          tma_load_3d(smem_base + r.stage * C::SLOT_A, &map_a, a_full(r.stage), m0, kb * BLOCK_K, b);
          r.advance();
        }
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {
      RingN<C::NSS> r;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int b, m0, n0, kb0, kb1, sp;
        decode(item, b, m0, n0, kb0, kb1, sp);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(s_empty(r.stage), r.phase ^ 1);
          const uint32_t st = smem_base + C::OFF_B + r.stage * 2 * C::TILE_B;
          mbar_expect_tx(b_full(r.stage), fuse_elu_grad ? 2 * C::TILE_B : C::TILE_B);
#pragma unroll
          for (int q = 0; q < BN / 32; ++q) tma_load_3d(st + q * BOX_BYTES, &map_b, b_full(r.stage), n0 + q * 32, kb * BLOCK_K, b);
          if (fuse_elu_grad) {
#pragma unroll
            for (int q = 0; q < BN / 32; ++q)
              tma_load_3d(st + C::TILE_B + q * BOX_BYTES, &map_b2, b_full(r.stage), n0 + q * 32, kb * BLOCK_K, b);
          }
          r.advance();
        }
      }
    }
  } else if (warp == 1) {
    RingN<C::NSS> r;
    int drain = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, m0, n0, kb0, kb1, sp;
      decode(item, b, m0, n0, kb0, kb1, sp);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int rel = kb - kb0;
        const int acc = drain & 1;
        if (rel % PROMOTE == 0) {
          mbar_wait(tempty_bar(acc), ((drain >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t tmem_d = tmem_base + acc * 2 * BN;
        const uint32_t tmem_c = tmem_d + BN;
        mbar_wait(s_ready(r.stage), r.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + C::OFF_B + r.stage * 2 * C::TILE_B;
          const uint32_t fresh = (rel % PROMOTE) == 0 ? 0u : 1u;
          const uint32_t a_hi = tmem_base + C::TMEM_A + r.stage * 2 * BLOCK_K, a_lo = a_hi + BLOCK_K;
          const uint64_t b_hi = tn::umma_desc_mn(st);  // Bhi boxes, then the Blo boxes: one 2 BN-wide operand
#pragma unroll
          for (int k = 0; k < BLOCK_K / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 1024 >> 4);
            umma_tf32_ta(tmem_d, a_hi + k * 8, b_hi + adv, C::IDESC2, fresh | (uint32_t)(k != 0));
            umma_tf32_ta(tmem_c, a_lo + k * 8, b_hi + adv, C::IDESC, 1);
          }
          umma_commit(s_empty(r.stage));
          if ((rel + 1) % PROMOTE == 0 || kb == kb1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if ((rel + 1) % PROMOTE == 0 || kb == kb1 - 1) ++drain;
        r.advance();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    RingN<C::NSA> ra;
    RingN<C::NSS> rs;
    const int t = threadIdx.x - 128;       // column of the raw A box = row of A^T = TMEM lane
    const int tc = t < box_m ? t : 0;      // lanes past the box hold rows that are never stored
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::TMEM_A;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, m0, n0, kb0, kb1, sp;
      decode(item, b, m0, n0, kb0, kb1, sp);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(a_full(ra.stage), ra.phase);
        const float* src = reinterpret_cast<const float*>(smem + ra.stage * C::SLOT_A) + tc;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float x = src[k * box_m];
          const uint32_t h = __float_as_uint(x) & 0xffffe000u;
          hi[k] = h;
          lo[k] = __float_as_uint(x - __uint_as_float(h));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(a_empty(ra.stage));
        // B raw has landed => the MMAs that used this slot (shared-memory B and tensor-memory A) have retired
        mbar_wait(b_full(rs.stage), rs.phase);
        tc_fence_after();
        const uint32_t ta = lane_base + rs.stage * 2 * BLOCK_K;
        tmem_st32(ta, hi);
        tmem_st32(ta + BLOCK_K, lo);
        float4* bh = reinterpret_cast<float4*>(smem + C::OFF_B + rs.stage * 2 * C::TILE_B);
        float4* bl = reinterpret_cast<float4*>(smem + C::OFF_B + rs.stage * 2 * C::TILE_B + C::TILE_B);
#pragma unroll
        for (int i = 0; i < C::TILE_B / 16 / 128; ++i) {
          const int idx = t + i * 128;
          float4 v = bh[idx];
          if (fuse_elu_grad) {
            const float4 o = bl[idx];
            v.x *= elu_grad_from_out(o.x);
            v.y *= elu_grad_from_out(o.y);
            v.z *= elu_grad_from_out(o.z);
            v.w *= elu_grad_from_out(o.w);
          }
          float4 h;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          bh[idx] = h;
          bl[idx] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        fence_proxy_async();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_ready(rs.stage));
        ra.advance();
        rs.advance();
      }
    }
  } else if (warp >= 8) {
    constexpr int HALF = BN / 2;  // columns per epilogue warp group
    const int quarter = warp & 3;
    const int half = (warp - 8) >> 2;
    const int row = quarter * 32 + lane;
    int drain = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int b, m0, n0, kb0, kb1, sp;
      decode(item, b, m0, n0, kb0, kb1, sp);
      float racc[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) racc[j] = 0.f;
      const int n_drains = (kb1 - kb0 + PROMOTE - 1) / PROMOTE;
      for (int d = 0; d < n_drains; ++d, ++drain) {
        const int acc = drain & 1;
        mbar_wait(tfull_bar(acc), (drain >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 2 * BN + half * HALF;
#pragma unroll
        for (int c = 0; c < HALF / 32; ++c) {
          uint32_t v[32], w[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld32(taddr + BN + c * 32, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[c * 32 + j] += __uint_as_float(v[j]) + __uint_as_float(w[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      const int gm = m0 + row;
      if (gm < Mo) {
        float* dst = part + (((int64_t)b * splits + sp) * Mo + gm) * No + n0 + half * HALF;
#pragma unroll
        for (int j = 0; j < HALF; ++j)
          if (n0 + half * HALF + j < No) dst[j] = racc[j];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM));
  }
}

// B_b (row-major [K,N] or, transposed, [N,K]) -> K-major hi / lo copies [batch][Npad][Kpad], zero padded
__global__ void split_transpose_b_batched_kernel(const float* __restrict__ B, int64_t ldb, int64_t b_bs, int K, int N,
                                                 int Kpad, int Npad, int batches, int b_is_nk, float* __restrict__ bhi,
                                                 float* __restrict__ blo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)Npad * Kpad;
  if (i >= per * batches) return;
  const int b = (int)(i / per);
  const int64_t r = i - b * per;
  const int n = (int)(r / Kpad), k = (int)(r % Kpad);
  const float* Bb = B + b * b_bs;
  float v = (n < N && k < K) ? (b_is_nk ? Bb[(int64_t)n * ldb + k] : Bb[(int64_t)k * ldb + n]) : 0.f;
  float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  bhi[i] = h;
  blo[i] = v - h;
}

__global__ void splitk_reduce_batched_kernel(int Mo, int No, int splits, int batches, const float* __restrict__ part,
                                             float* __restrict__ C, int64_t ldc, int64_t c_bs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)Mo * No;
  if (i >= per * batches) return;
  const int b = (int)(i / per);
  const int64_t r = i - b * per;
  const float* p = part + (int64_t)b * splits * per + r;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += p[(int64_t)z * per];
  C[b * c_bs + (r / No) * ldc + (r % No)] = s;
}

// 3-D fp32 tensor (cols of one batch, rows, batch): box = box_cols x box_rows x 1
static int make_map_3d(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld,
                       int64_t bs, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  GATK_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), (cuuint64_t)(batches > 1 ? bs : ld) * sizeof(float)};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GATK_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed (%d) cols=%lld rows=%lld batches=%lld ld=%lld bs=%lld",
               (int)rc, (long long)cols, (long long)rows, (long long)batches, (long long)ld, (long long)bs);
  return 0;
}

static int make_map_b(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  GATK_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GATK_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled (weights) failed (%d)", (int)rc);
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static void tn_plan(int64_t Mo, int64_t No, int64_t K, int batches, int bn, int* mt, int* nt, int* sp, int* kbt, int* kbs) {
  *mt = (int)((Mo + BLOCK_M - 1) / BLOCK_M);
  *nt = (int)((No + bn - 1) / bn);
  *kbt = (int)((K + BLOCK_K - 1) / BLOCK_K);
  int s = sm_count() / (*mt * *nt * batches);
  if (s < 1) s = 1;
  const int max_s = (*kbt + tn::PROMOTE - 1) / tn::PROMOTE;
  if (s > max_s) s = max_s;
  *kbs = (*kbt + s - 1) / s;
  *sp = (*kbt + *kbs - 1) / *kbs;
}

}  // namespace bt

// which tensor-core path a batched product takes: 1 = NN/NT (K <= 512), 2 = TN, 0 = none
static int batched_path(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches, const float* A, int64_t lda,
                        int64_t a_bs, const float* B, int64_t ldb, int64_t b_bs, const float* C, int64_t ldc, int64_t c_bs) {
  if (batches < 1 || batches > 64) return 0;
  if ((lda & 3) || (a_bs & 3) || (A && !bt::aligned16(A))) return 0;
  if (!transA) {
    // K up to 576: the 64-wide tile keeps the compensation products in their own accumulator (accumulation-truncation
    // error ~2.4e-8 per k-step of 8: 1.7e-6 at K = 576); the 128-wide tile keeps all three products in ONE accumulator
    // and is only used for short reductions (K <= 128), longer ones with N > 64 run as several 64-wide n-tiles
    if (M < 1024 || N < 8 || N > 4096 || K < 1 || K > 576 || M >= (1LL << 31) - 256) return 0;
    if ((ldc & 3) || (c_bs & 3) || (C && !bt::aligned16(C))) return 0;
    return 1;
  }
  if (transB) return 0;
  if (K < 2048 || M < 8 || N < 8 || M > 512 || N > 512 || K >= (1LL << 31) - 64) return 0;
  if ((ldb & 3) || (b_bs & 3) || (B && !bt::aligned16(B))) return 0;
  return 2;
}

// tile width of the NN / NT batched kernel: 128 (one merged accumulator) only for short reductions
static int batched_bn(int64_t N, int64_t K) { return (N <= 64 || K > 128) ? 64 : 128; }

size_t gemm_batched_tc_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches) {
  if (!transA) {
    const int bn = batched_bn(N, K);
    const int64_t Kpad = (K + tc::BLOCK_K - 1) / tc::BLOCK_K * tc::BLOCK_K;
    const int64_t Npad = (N + bn - 1) / bn * bn;
    return (size_t)(2 * batches * Npad * Kpad * sizeof(float) + 256);
  }
  if (transB) return 0;
  int mt, nt, sp, kbt, kbs;
  bt::tn_plan(M, N, K, batches, N <= 64 ? 64 : 128, &mt, &nt, &sp, &kbt, &kbs);
  return (size_t)batches * sp * M * N * sizeof(float);
}

// elu_out != NULL: the dh' operand (A of the NT product, B of the TN product) is multiplied by ELU'(v) taken from
// elu_out = ELU(v), same shape / batch stride as that operand, row pitch ld_elu
bool gemm_batched_fuses_elu_grad(int path, int64_t N) { return path == 1 || (path == 2 && N <= 64); }

int gemm_batched_tc_launch(int path, int transB, int64_t M, int64_t N, int64_t K, int batches, const float* A, int64_t lda,
                           int64_t a_bs, const float* B, int64_t ldb, int64_t b_bs, float* C, int64_t ldc, int64_t c_bs,
                           int epilogue, const float* elu_out, int64_t ld_elu, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace bt;
  GATK_REQUIRE(!elu_out || (gemm_batched_fuses_elu_grad(path, N) && (ld_elu & 3) == 0 && aligned16(elu_out)),
               "elu_out: unsupported shape for the fused ELU' operand (query gatk_gemm_batched_fuses_elu_grad)");
  const int fuse = elu_out ? 1 : 0;
  GATK_REQUIRE(ws && ws_bytes >= gemm_batched_tc_workspace_bytes(path == 2, transB, M, N, K, batches),
               "batched GEMM workspace too small");
  const int bn = path == 1 ? batched_bn(N, K) : (N <= 64 ? 64 : 128);
  if (path == 1) {
    const int Kpad = (int)((K + BLOCK_K - 1) / BLOCK_K * BLOCK_K);
    const int Npad = (int)((N + bn - 1) / bn * bn);
    float* bhi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    float* blo = bhi + (size_t)batches * Npad * Kpad;
    const int64_t total = (int64_t)batches * Npad * Kpad;
    split_transpose_b_batched_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, ldb, b_bs, (int)K, (int)N, Kpad, Npad,
                                                                                       batches, transB ? 1 : 0, bhi, blo);
    GATK_CHECK_LAUNCH();
    CUtensorMap map_a, map_a2, map_bhi, map_blo, map_c;
    if (int rc = make_map_3d(&map_a, A, K, M, batches, lda, a_bs, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map_3d(&map_a2, fuse ? elu_out : A, K, M, batches, fuse ? ld_elu : lda, a_bs, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
    if (int rc = make_map_b(&map_bhi, bhi, (int64_t)batches * Npad, Kpad, bn)) return rc;
    if (int rc = make_map_b(&map_blo, blo, (int64_t)batches * Npad, Kpad, bn)) return rc;
    const int box_n = bn == 64 ? 32 : (int)(N >= 128 ? 128 : (N + 3) / 4 * 4);  // wide tile: one dense store box per tile
    if (int rc = make_map_3d(&map_c, C, N, M, batches, ldc, c_bs, box_n, 128,
                             bn == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
    const int m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M), n_tiles = Npad / bn, k_blocks = Kpad / BLOCK_K;
    const int64_t tiles = (int64_t)m_tiles * n_tiles * batches;
    GATK_REQUIRE(tiles < (1LL << 31), "too many tiles");
    int grid = sm_count();
    if (tiles < grid) grid = (int)tiles;
    static int order = -1;
    if (order < 0) {
      const char* e = getenv("GATK_ORDER");
      order = e ? atoi(e) : 1;
    }
    if (bn == 64) {
      static bool configured = false;
      if (!configured) {
        GATK_CHECK_CUDA(cudaFuncSetAttribute(gemm_batched_tf32x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgTA<64>::SMEM));
        configured = true;
      }
      gemm_batched_tf32x3_kernel<64><<<grid, BT_THREADS, CfgTA<64>::SMEM, st>>>(map_a, map_bhi, map_blo, map_c, m_tiles, n_tiles,
                                                                               batches, k_blocks, Npad, epilogue, order, box_n, map_a2, fuse);
    } else {
      static bool configured = false;
      if (!configured) {
        GATK_CHECK_CUDA(cudaFuncSetAttribute(gemm_batched_tf32x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgTA<128>::SMEM));
        configured = true;
      }
      gemm_batched_tf32x3_kernel<128><<<grid, BT_THREADS, CfgTA<128>::SMEM, st>>>(map_a, map_bhi, map_blo, map_c, m_tiles, n_tiles,
                                                                                 batches, k_blocks, Npad, epilogue, order, box_n, map_a2, fuse);
    }
    GATK_CHECK_LAUNCH();
    return 0;
  }
  // TN: here M = Mo (columns of A_b), N = No (columns of B_b), K = rows
  int mt, nt, sp, kbt, kbs;
  tn_plan(M, N, K, batches, bn, &mt, &nt, &sp, &kbt, &kbs);
  CUtensorMap map_a, map_b, map_b2;
  if (int rc = make_map_3d(&map_b, B, N, K, batches, ldb, b_bs, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  if (int rc = make_map_3d(&map_b2, fuse ? elu_out : B, N, K, batches, fuse ? ld_elu : ldb, b_bs, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return rc;
  const int items = mt * nt * sp * batches;
  int grid = sm_count();
  if (items < grid) grid = items;
  float* part = static_cast<float*>(ws);
  if (bn == 64) {
    // A through tensor memory: one dense [32 rows][box_m columns] box per k-block
    const int box_m = (int)(M >= 128 ? 128 : (M + 3) / 4 * 4);
    if (int rc = make_map_3d(&map_a, A, M, K, batches, lda, a_bs, box_m, 32, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
    static bool configured = false;
    if (!configured) {
      GATK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_batched_ta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgTNA::SMEM));
      configured = true;
    }
    gemm_tn_batched_ta_kernel<<<grid, tn::TN_THREADS, CfgTNA::SMEM, st>>>(map_a, map_b, part, (int)M, (int)N, mt, nt, batches, sp, kbt, kbs,
                                                                          box_m, map_b2, fuse);
  } else {
    if (int rc = make_map_3d(&map_a, A, M, K, batches, lda, a_bs, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    static bool configured = false;
    if (!configured) {
      GATK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_batched_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_TN));
      configured = true;
    }
    gemm_tn_batched_kernel<128><<<grid, tn::TN_THREADS, Cfg<128>::SMEM_TN, st>>>(map_a, map_b, part, (int)M, (int)N, mt, nt, batches, sp, kbt, kbs);
  }
  GATK_CHECK_LAUNCH();
  const int64_t total = (int64_t)batches * M * N;
  splitk_reduce_batched_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((int)M, (int)N, sp, batches, part, C, ldc, c_bs);
  GATK_CHECK_LAUNCH();
  return 0;
}

int gemm_batched_path(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches, const float* A, int64_t lda,
                      int64_t a_bs, const float* B, int64_t ldb, int64_t b_bs, const float* C, int64_t ldc, int64_t c_bs) {
  return batched_path(transA, transB, M, N, K, batches, A, lda, a_bs, B, ldb, b_bs, C, ldc, c_bs);
}

bool gemm_tc_eligible(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                      const float* C, int64_t ldc, int accumulate) {
  if (transA || transB || accumulate) return false;
  if (M < 1024 || N < 8 || K < 1) return false;  // small problems: launch-bound either way, keep exact-fp32 SIMT
  // accumulation-truncation error grows ~2.4e-8 per k-step of 8 (measured): keep it under ~2e-6
  if (M >= (1LL << 31) - 256 || N > 65536 || K > 512) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(C) & 15)) return false;
  if ((lda & 3) || (ldc & 3)) return false;
  return true;
}

size_t gemm_tc_workspace_bytes(int64_t N, int64_t K) {
  const int64_t Kpad = (K + tc::BLOCK_K - 1) / tc::BLOCK_K * tc::BLOCK_K;
  const int64_t Npad = (N + tc::BLOCK_N - 1) / tc::BLOCK_N * tc::BLOCK_N;
  return (size_t)(2 * Npad * Kpad * sizeof(float) + 256);
}

int gemm_tc_launch(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                   int64_t ldc, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace tc;
  const int Kpad = (int)((K + BLOCK_K - 1) / BLOCK_K * BLOCK_K);
  const int Npad = (int)((N + BLOCK_N - 1) / BLOCK_N * BLOCK_N);
  GATK_REQUIRE(ws && ws_bytes >= gemm_tc_workspace_bytes(N, K), "tensor-core GEMM workspace too small");
  float* bhi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* blo = bhi + (size_t)Npad * Kpad;
  const int64_t total = (int64_t)Npad * Kpad;
  split_transpose_b_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, ldb, (int)K, (int)N, Kpad, Npad, 0, bhi, blo);
  GATK_CHECK_LAUNCH();

  CUtensorMap map_a, map_bhi, map_blo, map_c;
  if (int rc = make_map(&map_a, A, M, K, lda)) return rc;
  if (int rc = make_map(&map_bhi, bhi, Npad, Kpad, Kpad)) return rc;
  if (int rc = make_map(&map_blo, blo, Npad, Kpad, Kpad)) return rc;
  if (int rc = make_map(&map_c, C, M, N, ldc)) return rc;

  static bool configured = false;
  if (!configured) {
    GATK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M), n_tiles = Npad / BLOCK_N, k_blocks = Kpad / BLOCK_K;
  const int64_t tiles = (int64_t)m_tiles * n_tiles;
  GATK_REQUIRE(tiles < (1LL << 31), "too many tiles");
  int grid = sm_count();
  if (tiles < grid) grid = (int)tiles;
  gemm_tf32x3_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_a, map_bhi, map_blo, map_c, m_tiles, n_tiles, k_blocks);
  GATK_CHECK_LAUNCH();
  return 0;
}


// ---- TN (dW) path ---------------------------------------------------------------------------------
static void tn_plan(int64_t Mo, int64_t No, int64_t K, int* m_tiles, int* n_tiles, int* splits, int* kb_total,
                    int* kb_per_split) {
  *m_tiles = (int)((Mo + tc::BLOCK_M - 1) / tc::BLOCK_M);
  *n_tiles = (int)((No + tc::BLOCK_N - 1) / tc::BLOCK_N);
  *kb_total = (int)((K + tc::BLOCK_K - 1) / tc::BLOCK_K);
  const int tiles = *m_tiles * *n_tiles;
  int s = sm_count() / tiles;
  if (s < 1) s = 1;
  const int max_s = (*kb_total + tn::PROMOTE - 1) / tn::PROMOTE;  // at least one drain interval per split
  if (s > max_s) s = max_s;
  *kb_per_split = (*kb_total + s - 1) / s;
  *splits = (*kb_total + *kb_per_split - 1) / *kb_per_split;
}

bool gemm_tn_tc_eligible(int transA, int transB, int64_t Mo, int64_t No, int64_t K, const float* A, int64_t lda,
                         const float* B, int64_t ldb, int accumulate) {
  (void)accumulate;
  if (!transA || transB) return false;
  if (K < 2048 || Mo < 8 || No < 8 || Mo > 4096 || No > 8192) return false;
  if (K >= (1LL << 31) - 64) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if ((lda & 3) || (ldb & 3)) return false;
  return true;
}

size_t gemm_tn_tc_workspace_bytes(int64_t Mo, int64_t No, int64_t K) {
  int mt, nt, sp, kbt, kbs;
  tn_plan(Mo, No, K, &mt, &nt, &sp, &kbt, &kbs);
  return (size_t)sp * Mo * No * sizeof(float);
}

int gemm_tn_tc_launch(int64_t Mo, int64_t No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                      float* part, int* splits_out, cudaStream_t st) {
  using namespace tn;
  int mt, nt, sp, kbt, kbs;
  tn_plan(Mo, No, K, &mt, &nt, &sp, &kbt, &kbs);
  CUtensorMap map_a, map_b;
  if (int rc = make_map_box32(&map_a, A, K, Mo, lda)) return rc;
  if (int rc = make_map_box32(&map_b, B, K, No, ldb)) return rc;
  static bool configured = false;
  if (!configured) {
    GATK_CHECK_CUDA(cudaFuncSetAttribute(gemm_promoted_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES));
    configured = true;
  }
  const int items = mt * nt * sp;
  int grid = sm_count();
  if (items < grid) grid = items;
  gemm_promoted_tf32x3_kernel<false><<<grid, TN_THREADS, TN_SMEM_BYTES, st>>>(map_a, map_b, map_b, part, 0, 0, (int)Mo, (int)No,
                                                                             mt, nt, sp, kbt, kbs);
  GATK_CHECK_LAUNCH();
  *splits_out = sp;
  return 0;
}

// ---- long-K / NT projection path: C[M,N] (+)= A[M,K] * op(B), any K, C written with plain stores ---------
bool gemm_long_tc_eligible(int transA, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda) {
  if (transA) return false;
  if (M < 1024 || N < 8 || K < 1 || K > 65536 || N > 65536 || M >= (1LL << 31) - 256) return false;
  if (reinterpret_cast<uintptr_t>(A) & 15) return false;
  return (lda & 3) == 0;
}

int gemm_long_tc_launch(int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                        int64_t ldb, float* C, int64_t ldc, int accumulate, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  using namespace tn;
  const int Kpad = (int)((K + BLOCK_K - 1) / BLOCK_K * BLOCK_K);
  const int Npad = (int)((N + BLOCK_N - 1) / BLOCK_N * BLOCK_N);
  GATK_REQUIRE(ws && ws_bytes >= gemm_tc_workspace_bytes(N, K), "tensor-core GEMM workspace too small");
  float* bhi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* blo = bhi + (size_t)Npad * Kpad;
  const int64_t total = (int64_t)Npad * Kpad;
  split_transpose_b_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, ldb, (int)K, (int)N, Kpad, Npad,
                                                                             transB ? 1 : 0, bhi, blo);
  GATK_CHECK_LAUNCH();
  CUtensorMap map_a, map_bhi, map_blo;
  if (int rc = make_map(&map_a, A, M, K, lda)) return rc;
  if (int rc = make_map(&map_bhi, bhi, Npad, Kpad, Kpad)) return rc;
  if (int rc = make_map(&map_blo, blo, Npad, Kpad, Kpad)) return rc;
  static bool configured = false;
  if (!configured) {
    GATK_CHECK_CUDA(cudaFuncSetAttribute(gemm_promoted_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES));
    configured = true;
  }
  const int mt = (int)((M + BLOCK_M - 1) / BLOCK_M), nt = Npad / BLOCK_N, kbt = Kpad / BLOCK_K;
  const int64_t items = (int64_t)mt * nt;
  GATK_REQUIRE(items < (1LL << 31), "too many tiles");
  int grid = sm_count();
  if (items < grid) grid = (int)items;
  gemm_promoted_tf32x3_kernel<true><<<grid, TN_THREADS, TN_SMEM_BYTES, st>>>(map_a, map_bhi, map_blo, C, ldc, accumulate, (int)M,
                                                                            (int)N, mt, nt, 1, kbt, kbt);
  GATK_CHECK_LAUNCH();
  return 0;
}

}  // namespace gatk

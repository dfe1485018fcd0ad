// User tower of the GDMCF denoiser in ONE launch (bf16 mode):
//
//   g1  = relu(hc W1^T + b1)            LayerGCN.conv1 on the user rows, models/DNN.py:1093-1096   [B, H],  K = 3d
//   g2  = g1 W2^T + b2                  LayerGCN.conv2,                 models/DNN.py:1100          [B, 3d], K = H
//   hc' = hc * sumW + g2 * (1 - sumW)   models/DNN.py:1288
//   inv_u[b] = 1 / ||hc'[b, :]||        user norms of cosine_similarity_cuda, models/DNN.py:1320
//
// As two tcgen05 contractions + two elementwise kernels this is 5 launches of ~1.2 GFLOP each whose duration is all
// pipeline fill and drain (23 + 19 + 11 us per reverse step at the Yelp shape, round-1 ncu list). Here the grid is one
// resident wave of CTAs (at most one per SM) that walks three phases separated by grid barriers:
//   phase 1  units (m-block, 64-wide slice of H, K-split): TMA-fed tcgen05.mma M=128 N=64, fp32 partials -> L2 workspace
//   reduce   every thread: sum the K-split partials in fixed order + b1, relu -> g1 (bf16 operand, optional fp32 copy)
//   phase 2  units (m-block, 128-wide tile of 3d): TMA-fed tcgen05.mma M=128 N=128 over K = H, epilogue = + b2, sumW mix
//            with hc (fp32), bf16 operand hc' + per-tile row sums of squares; the last tile of an m-block to finish
//            adds the partial sums in tile order and writes inv_u.
// Everything is deterministic (no floating-point atomics). The barrier counters live in a caller-provided, zero-initialised
// 64-byte block that the kernel leaves zeroed again.
#include <cuda.h>
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace tower {

constexpr int BM = 128, BK = 64, UMMA_K = 16, STAGES = 4, THREADS = 256;
constexpr int N1 = 64, N2 = 128;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = N2 * BK * 2;   // 16 KB (phase 1 fills the first 8 KB)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int TMEM_COLS = 128;
constexpr int MAX_MBLOCKS = 8;

struct Maps {
  CUtensorMap hc, w1, g1, w2;
};

struct Params {
  int B, K1, H, N;         // rows, 3d (K of conv1), hidden (N of conv1 = K of conv2), 3d (N of conv2)
  int kb1, ksplit, kb_per_split, kb2;
  int mblocks, nslices1, ntiles2;
  const float* b1; const float* b2; const float* sumw;
  const float* hc_f32; long long ld_hc;
  float* ws1;                               // [ksplit][B][H] fp32 partials of conv1
  __nv_bfloat16* g1; float* g1_f32;         // [B][H]
  float* g2_f32; long long ld_g2;           // optional [B][N]
  __nv_bfloat16* hcp; long long ld_hcp;     // [B][ld_hcp] bf16 operand of the scorer
  float* hcp_f32; long long ld_hcp32;       // optional
  float* rowpart;                           // [ntiles2][B]
  float* inv_u;                             // [B]
  unsigned int* sync;                       // [0] barrier A, [1] barrier B, [2] exit count, [4 + mb] tiles done per m-block
};

GD_DEV unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
GD_DEV void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// All CTAs of the grid are resident (grid <= SM budget, one CTA per SM): counter barrier in global memory.
GD_DEV void grid_barrier(unsigned int* ctr, unsigned int expected) {
  fence_proxy_async_all();  // generic-proxy global writes of this thread -> visible to other CTAs' TMA loads
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int spins = 0;
    uint64_t t0 = 0;
    while (ld_acquire(ctr) < expected) {
      __nanosleep(64);
      if (++spins == 1024u) t0 = globaltimer_ns();
      if (spins > 1024u && (spins & 255u) == 0u && globaltimer_ns() - t0 > 4000000000ull) __trap();
    }
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
}

__global__ void __launch_bounds__(THREADS, 1)
user_tower_kernel(const __grid_constant__ Maps maps, const Params p) {
  pdl_launch_dependents();
  const uint64_t t_start = globaltimer_ns();  // phase timestamps of CTA 0 -> sync[12..15] (ns since kernel entry; diagnostics)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  uint32_t* flag_slot = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.hc); tma_prefetch_desc(&maps.w1); tma_prefetch_desc(&maps.g1); tma_prefetch_desc(&maps.w2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // ring position (the producer and the MMA issuer walk the same sequence) and accumulator parity
  int stage = 0;
  uint32_t ring_phase = 0, acc_phase = 0;
  const int ew = warp & 3;                       // TMEM lane quarter of an epilogue warp (warps 4..7)

  // =============================== phase 1: conv1 partial products ===============================
  const int units1 = p.mblocks * p.nslices1 * p.ksplit;
  for (int u = blockIdx.x; u < units1; u += gridDim.x) {
    const int mb = u % p.mblocks;
    const int ns = (u / p.mblocks) % p.nslices1;
    const int ks = u / (p.mblocks * p.nslices1);
    const int kb_begin = ks * p.kb_per_split, kb_end = min(kb_begin + p.kb_per_split, p.kb1);
    if (warp == 0 && lane == 0) {
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], ring_phase ^ 1);
        uint8_t* sa = tiles + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], A_BYTES + N1 * BK * 2);
        tma_load_2d(sa, &maps.hc, &full_bar[stage], kb * BK, mb * BM);
        tma_load_2d(sa + A_BYTES, &maps.w1, &full_bar[stage], kb * BK, ns * N1);
        if (++stage == STAGES) { stage = 0; ring_phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, N1);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], ring_phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
        const uint64_t da = umma_desc_kmajor_sw128(sa), db = umma_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; ring_phase ^= 1; }
      }
      umma_commit(tfull_bar);
    } else if (warp >= 4) {
      mbar_wait(tfull_bar, acc_phase);
      tc_fence_after();
      const int row = mb * BM + ew * 32 + lane;
      float v[64];
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16);
      tmem_ld_32x32(t_row, *reinterpret_cast<float(*)[32]>(v));
      tmem_ld_32x32(t_row + 32u, *reinterpret_cast<float(*)[32]>(v + 32));
      tmem_ld_wait();
      if (row < p.B && kb_end > kb_begin) {
        float4* dst = reinterpret_cast<float4*>(p.ws1 + ((long long)ks * p.B + row) * p.H + ns * N1);
#pragma unroll
        for (int q = 0; q < 16; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      tc_fence_before();
    }
    // `stage` / `ring_phase` are private to the two elected lanes (each walks the same k-block sequence); everyone
    // meets at the end of the unit: the single accumulator is free again once the epilogue warps have read it
    __syncthreads();
    acc_phase ^= 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) p.sync[12] = (unsigned int)(globaltimer_ns() - t_start);
  grid_barrier(&p.sync[0], gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x == 0) p.sync[13] = (unsigned int)(globaltimer_ns() - t_start);

  // =============================== reduce: g1 = relu(sum_k partial + b1) ===============================
  {
    const int quads = p.H >> 2;
    const long long total = (long long)p.B * quads;
    const long long slab = (long long)p.B * p.H;
    for (long long i = blockIdx.x * (long long)THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * THREADS) {
      const int r = (int)(i / quads), c = (int)(i % quads) * 4;
      const float* src = p.ws1 + (long long)r * p.H + c;
      float4 a = *reinterpret_cast<const float4*>(src);
      for (int s = 1; s < p.ksplit; ++s) {
        const float4 x = *reinterpret_cast<const float4*>(src + s * slab);
        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      }
      const float4 b = *reinterpret_cast<const float4*>(p.b1 + c);
      a.x = fmaxf(a.x + b.x, 0.f); a.y = fmaxf(a.y + b.y, 0.f); a.z = fmaxf(a.z + b.z, 0.f); a.w = fmaxf(a.w + b.w, 0.f);
      if (p.g1_f32) *reinterpret_cast<float4*>(p.g1_f32 + (long long)r * p.H + c) = a;
      uint2 pk;
      pk.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.x)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.y)) << 16);
      pk.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.z)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.w)) << 16);
      *reinterpret_cast<uint2*>(p.g1 + (long long)r * p.H + c) = pk;
    }
  }
  grid_barrier(&p.sync[1], gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x == 0) p.sync[14] = (unsigned int)(globaltimer_ns() - t_start);

  // =============================== phase 2: conv2 + mix + norms ===============================
  const float sw = p.sumw[0], sw1 = 1.0f - sw;
  const int units2 = p.mblocks * p.ntiles2;
  for (int u = blockIdx.x; u < units2; u += gridDim.x) {
    const int mb = u % p.mblocks, nt = u / p.mblocks;
    if (warp == 0 && lane == 0) {
      for (int kb = 0; kb < p.kb2; ++kb) {
        mbar_wait(&empty_bar[stage], ring_phase ^ 1);
        uint8_t* sa = tiles + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_2d(sa, &maps.g1, &full_bar[stage], kb * BK, mb * BM);
        tma_load_2d(sa + A_BYTES, &maps.w2, &full_bar[stage], kb * BK, nt * N2);
        if (++stage == STAGES) { stage = 0; ring_phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, N2);
      for (int kb = 0; kb < p.kb2; ++kb) {
        mbar_wait(&full_bar[stage], ring_phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
        const uint64_t da = umma_desc_kmajor_sw128(sa), db = umma_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; ring_phase ^= 1; }
      }
      umma_commit(tfull_bar);
    } else if (warp >= 4) {
      mbar_wait(tfull_bar, acc_phase);
      tc_fence_after();
      const int row = mb * BM + ew * 32 + lane;
      const bool row_ok = row < p.B;
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16);
      float ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < N2; c += 32) {
        const int n0 = nt * N2 + c;
        if (n0 >= p.ld_hcp) break;  // warp-uniform; chunks in [N, ld_hcp) only zero the operand's K padding
        float v[32];
        tmem_ld_32x32(t_row + (uint32_t)c, v);
        float4 h[8];
        const float* hrow = p.hc_f32 + (long long)row * p.ld_hc + n0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          h[q] = (row_ok && n0 + 4 * q < p.N) ? __ldg(reinterpret_cast<const float4*>(hrow + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        tmem_ld_wait();
        const float hv[32] = {h[0].x, h[0].y, h[0].z, h[0].w, h[1].x, h[1].y, h[1].z, h[1].w, h[2].x, h[2].y, h[2].z, h[2].w,
                              h[3].x, h[3].y, h[3].z, h[3].w, h[4].x, h[4].y, h[4].z, h[4].w, h[5].x, h[5].y, h[5].z, h[5].w,
                              h[6].x, h[6].y, h[6].z, h[6].w, h[7].x, h[7].y, h[7].z, h[7].w};
        float g2[32], y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool ok = n0 + j < p.N;
          g2[j] = ok ? v[j] + __ldg(p.b2 + n0 + j) : 0.f;
          // hc * sumW + all_embeddings[:B] * (1 - sumW)   (models/DNN.py:1288), same association as mix_rownorm_kernel
          y[j] = ok ? hv[j] * sw + g2[j] * sw1 : 0.f;
          ss += y[j] * y[j];
        }
        if (row_ok) {
          if (p.g2_f32) {
            float4* d = reinterpret_cast<float4*>(p.g2_f32 + (long long)row * p.ld_g2 + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (n0 + 4 * q < p.N) d[q] = make_float4(g2[4 * q], g2[4 * q + 1], g2[4 * q + 2], g2[4 * q + 3]);
          }
          if (p.hcp_f32) {
            float4* d = reinterpret_cast<float4*>(p.hcp_f32 + (long long)row * p.ld_hcp32 + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (n0 + 4 * q < p.N) d[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
          }
          uint4* d16 = reinterpret_cast<uint4*>(p.hcp + (long long)row * p.ld_hcp + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (n0 + 8 * q < p.ld_hcp) {  // columns in [N, ld_hcp) receive zeros (K padding of the scorer's operand)
              uint4 pk;
              pk.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q + 1])) << 16);
              pk.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q + 2])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q + 3])) << 16);
              pk.z = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q + 4])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q + 5])) << 16);
              pk.w = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q + 6])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y[8 * q + 7])) << 16);
              d16[q] = pk;
            }
          }
        }
      }
      if (row_ok) p.rowpart[(long long)nt * p.B + row] = ss;
      tc_fence_before();
    }
    __threadfence();  // rowpart / hc' stores of this thread are visible before the tile is counted as done
    __syncthreads();
    acc_phase ^= 1;
    // the last tile of this m-block to finish turns the per-tile partial sums into inverse norms (fixed tile order)
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(&p.sync[4 + mb], 1u);
      *flag_slot = (done + 1u == (unsigned int)p.ntiles2) ? 1u : 0u;
      __threadfence();
    }
    __syncthreads();
    if (*flag_slot) {
      for (int r = mb * BM + threadIdx.x; r < min(p.B, (mb + 1) * BM); r += THREADS) {
        float tot = 0.f;
        for (int t = 0; t < p.ntiles2; ++t) tot += __ldcg(p.rowpart + (long long)t * p.B + r);
        p.inv_u[r] = 1.0f / sqrtf(tot);
      }
    }
    __syncthreads();
  }

  // leave the counter block zeroed for the next launch: the last CTA to get here resets it
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) p.sync[15] = (unsigned int)(globaltimer_ns() - t_start);
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int n = atomicAdd(&p.sync[2], 1u);
    if (n + 1u == gridDim.x) {
      p.sync[0] = 0u; p.sync[1] = 0u; p.sync[2] = 0u;
      for (int i = 0; i < MAX_MBLOCKS; ++i) p.sync[4 + i] = 0u;
      __threadfence();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* map, const void* ptr, int rows, int cols, long long ld, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return GDMCF_ECUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("user_tower: cuTensorMapEncodeTiled failed (%d)", (int)r); return GDMCF_EBADARG; }
  return GDMCF_OK;
}

static int pick_ksplit(int kb1) { return kb1 >= 32 ? 4 : (kb1 >= 12 ? 2 : 1); }

}  // namespace tower
}  // namespace gd

using namespace gd;
using namespace gd::tower;

extern "C" size_t gdmcf_user_tower_workspace_bytes(int rows, int k1, int hidden, int n) {
  if (rows <= 0 || k1 <= 0 || hidden <= 0 || n <= 0) return 0;
  const int kb1 = (k1 + BK - 1) / BK;
  const size_t ws1 = (size_t)pick_ksplit(kb1) * rows * hidden * 4;
  const size_t g1 = (size_t)rows * hidden * 2;
  const size_t rowpart = (size_t)((n + N2 - 1) / N2) * rows * 4;
  return ((ws1 + 255) / 256 + (g1 + 255) / 256 + (rowpart + 255) / 256) * 256;
}

extern "C" int gdmcf_user_tower(const void* hc_bf16, int64_t ld_hcb, const float* hc_f32, int64_t ld_hc, const void* w1_bf16,
                                int64_t ld_w1, const float* b1, const void* w2_bf16, int64_t ld_w2, const float* b2,
                                const float* sumw, int rows, int k1, int hidden, int n, void* hcp_bf16, int64_t ld_hcp,
                                float* inv_u, float* g1_f32, float* g2_f32, int64_t ld_g2, float* hcp_f32, int64_t ld_hcp32,
                                void* workspace, size_t workspace_bytes, uint32_t* sync_block, int max_ctas,
                                gdmcf_stream_t stream) {
  if (!hc_bf16 || !hc_f32 || !w1_bf16 || !b1 || !w2_bf16 || !b2 || !sumw || !hcp_bf16 || !inv_u || !workspace || !sync_block ||
      rows <= 0 || k1 <= 0 || hidden <= 0 || n <= 0) {
    set_error("user_tower: null pointer or bad shape");
    return GDMCF_EBADARG;
  }
  if ((hidden % 64) || (ld_hcb & 7) || (ld_w1 & 7) || (ld_w2 & 7) || (ld_hcp & 7) || ld_hcb < k1 || ld_w1 < k1 || ld_w2 < hidden ||
      ld_hcp < n || (ld_hc & 3) || ld_hc < n || n != k1 || (g2_f32 && ((ld_g2 & 3) || ld_g2 < n)) ||
      (hcp_f32 && ((ld_hcp32 & 3) || ld_hcp32 < n)) ||
      ((((uintptr_t)hc_bf16 | (uintptr_t)w1_bf16 | (uintptr_t)w2_bf16 | (uintptr_t)hcp_bf16 | (uintptr_t)hc_f32 | (uintptr_t)b1 |
         (uintptr_t)workspace | (uintptr_t)g1_f32 | (uintptr_t)g2_f32 | (uintptr_t)hcp_f32) & 15) != 0)) {
    set_error("user_tower: needs hidden %% 64 == 0, n == k1, 16 B aligned pointers, ld %% 8 == 0 (bf16) / %% 4 == 0 (fp32)");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  Params p{};
  p.B = rows; p.K1 = k1; p.H = hidden; p.N = n;
  p.kb1 = (k1 + BK - 1) / BK;
  p.kb_per_split = (p.kb1 + pick_ksplit(p.kb1) - 1) / pick_ksplit(p.kb1);
  p.ksplit = (p.kb1 + p.kb_per_split - 1) / p.kb_per_split;  // every split owns at least one k-block
  p.kb2 = hidden / BK;
  p.mblocks = (rows + BM - 1) / BM;
  p.nslices1 = hidden / N1;
  p.ntiles2 = (n + N2 - 1) / N2;
  if (p.mblocks > MAX_MBLOCKS) { set_error("user_tower: at most %d rows per call", MAX_MBLOCKS * BM); return GDMCF_EBADARG; }
  if (workspace_bytes < gdmcf_user_tower_workspace_bytes(rows, k1, hidden, n)) { set_error("user_tower: workspace too small"); return GDMCF_EBADARG; }
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  p.ws1 = reinterpret_cast<float*>(w);
  w += ((size_t)p.ksplit * rows * hidden * 4 + 255) / 256 * 256;
  p.g1 = reinterpret_cast<__nv_bfloat16*>(w);
  w += ((size_t)rows * hidden * 2 + 255) / 256 * 256;
  p.rowpart = reinterpret_cast<float*>(w);
  p.b1 = b1; p.b2 = b2; p.sumw = sumw; p.hc_f32 = hc_f32; p.ld_hc = ld_hc;
  p.g1_f32 = g1_f32; p.g2_f32 = g2_f32; p.ld_g2 = ld_g2;
  p.hcp = reinterpret_cast<__nv_bfloat16*>(hcp_bf16); p.ld_hcp = ld_hcp;
  p.hcp_f32 = hcp_f32; p.ld_hcp32 = ld_hcp32;
  p.inv_u = inv_u; p.sync = sync_block;
  Maps maps;
  if ((rc = make_map(&maps.hc, hc_bf16, rows, k1, ld_hcb, BM))) return rc;
  if ((rc = make_map(&maps.w1, w1_bf16, hidden, k1, ld_w1, N1))) return rc;
  if ((rc = make_map(&maps.g1, p.g1, rows, hidden, hidden, BM))) return rc;
  if ((rc = make_map(&maps.w2, w2_bf16, n, hidden, ld_w2, N2))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(user_tower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(user_tower)");
    attr_set = true;
  }
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const int budget = (max_ctas > 0 && max_ctas < sms) ? max_ctas : sms;
  const int units = std::max(p.mblocks * p.nslices1 * p.ksplit, p.mblocks * p.ntiles2);
  const int grid = std::max(1, std::min(units, budget));
  launch_kernel_cooperative(user_tower_kernel, grid, THREADS, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), maps, p);
  return cuda_check_launch("user_tower_kernel");
}

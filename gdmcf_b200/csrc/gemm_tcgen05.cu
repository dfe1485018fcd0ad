// Denoiser contractions on the 5th-gen tensor cores.
//
//   C[M,N] = sum_s A_s[M,K_s] * B_s[N,K_s]^T        (bf16 operands, fp32 accumulation)
//
// One persistent CTA per SM walks work units (m-block fastest so that the CTAs running side by side
// share the weight tile in L2). Per CTA, warp-specialised:
//   warp 0     TMA producer   : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a 3/5-stage smem ring
//   warp 1     MMA issuer     : one elected lane issues tcgen05.mma (M=128, N=BN, K=16) into TMEM
//   warp 2     TMEM allocator : 2 accumulator stages (double-buffered against the epilogue)
//   warps 4-7  epilogue       : tcgen05.ld 32x32b -> fused epilogue -> global memory, two forms:
//     (a) TMA-store form (all outputs 16 B-strided): every lane keeps its own row, applies scale / bias /
//         activation / posterior mean in registers (x_t read as 16 independent 16 B loads per lane), writes
//         the row into a 128B-swizzled staging tile (conflict-free STS.128) and one lane issues
//         cp.async.bulk.tensor stores. ~0.15 instructions per output element per warp — the epilogue warps are
//         single-warp-per-scheduler, so instruction count, not bandwidth, is what bounds them (ncu, round 1).
//     (b) transposed form (odd leading dimensions, per-timestep bias tables, split-K partial slabs): a
//         32x65 smem transpose per warp turns lane = row into lane = column so that plain stores coalesce.
// Split-K units write fp32 partial slabs; splitk_reduce_kernel sums them in fixed order (deterministic)
// and applies the same epilogue formula.
//
// Reference call sites replaced: nn.Linear at models/DNN.py:79-86, :1240-1252, GCNConv linears on the
// user rows :1082-1100, torch.mm + norm division :1320-1325, posterior mean
// models/gaussian_diffusion.py:1041-1050.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARP_BYTES = 16384;  // per epilogue warp: fp32 [32x64] (8 KB) + bf16 hi (4 KB) + bf16 lo (4 KB)

template <int BN>
struct Cfg {
  // bytes in flight needed per SM = L2 bandwidth share (~12 TB/s / 148) x ~1 us latency ~ 81 KB: 3 x 48 KB is enough
  static constexpr int STAGES = (BN == 256) ? 3 : 5;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // power of two: 256 or 512
  static constexpr int BAR_BYTES = 1024;    // barriers + tmem slot, keeps the staging tiles 1024 B aligned
  static constexpr int COLVEC_BYTES = 2 * BN * 4;  // col_scale / bias of the current tile
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 4 * EPI_WARP_BYTES + COLVEC_BYTES;
};

struct TmaMaps {
  CUtensorMap a[GDMCF_MAX_SEG];
  CUtensorMap b[GDMCF_MAX_SEG];
  CUtensorMap o32, o16, olo;  // outputs (TMA-store epilogue only)
};

struct Shape {
  int m, n;
  int n_seg;
  int kb[GDMCF_MAX_SEG];  // k-blocks per segment
  int total_kb;
  int tiles_m, tiles_n;
  int splits, kb_per_split;
  long long slab_stride;  // elements between split slabs (fp32)
  int ld_ws;              // leading dim of a slab
  float* ws;              // split-K workspace (NULL when splits == 1)
  int tma_store;          // 1: epilogue form (a)
};

// ---------------------------------------------------------------------------------------------
// Epilogue on 8 consecutive columns of one row (split-K reducer).
// ---------------------------------------------------------------------------------------------
GD_DEV void store8_f32(float* dst, const float (&o)[8], int valid, bool vec_ok) {
  if (valid == 8 && vec_ok) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < valid) dst[j] = o[j];
  }
}
GD_DEV uint32_t pack_bf16x2(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
GD_DEV void store8_bf16(__nv_bfloat16* dst, const __nv_bfloat16 (&h)[8], int valid) {
  if (valid == 8) {
    uint4 pk;
    pk.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    pk.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    pk.z = (uint32_t)__bfloat16_as_ushort(h[4]) | ((uint32_t)__bfloat16_as_ushort(h[5]) << 16);
    pk.w = (uint32_t)__bfloat16_as_ushort(h[6]) | ((uint32_t)__bfloat16_as_ushort(h[7]) << 16);
    *reinterpret_cast<uint4*>(dst) = pk;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < valid) dst[j] = h[j];
  }
}

// Kept out of line: the activation is only used by the small encoder / GCN contractions.
__device__ __noinline__ float apply_act(float v, int act) {
  return act == GDMCF_ACT_TANH ? tanhf(v) : fmaxf(v, 0.f);
}

// acc: 8 accumulator values of row m, columns n0..n0+7 (n0 % 8 == 0); N = logical column count.
GD_DEV void epilogue8(const gdmcf_epilogue& e, int m, int n0, int N, const float (&acc)[8]) {
  const int valid = min(8, N - n0);
  if (valid <= 0) return;
  float o[8];
  const int t = e.row_t ? e.row_t[m] : e.t_const;
  // One formula for every fused form:  s = act(alpha*acc*row_scale[m]*col_scale[n] + bias[t,n]);
  //                                     out = c1 ? c1[t]*s + c2[t]*xt[m,n] : s
  const float rs = e.alpha * (e.row_scale ? e.row_scale[m] : 1.0f);
  const float* bias = e.bias ? e.bias + (long long)t * e.ld_bias + n0 : nullptr;
  float c1 = 1.f, c2 = 0.f;
  const float* xt = nullptr;
  if (e.c1) {
    c1 = e.c1[t];
    c2 = e.c2[t];
    xt = e.xt + (long long)m * e.ld_xt + n0;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = 0.f;
    if (j < valid) {
      v = acc[j] * rs;
      if (e.col_scale) v *= e.col_scale[n0 + j];
      if (bias) v += bias[j];
      if (e.act != GDMCF_ACT_NONE) v = apply_act(v, e.act);
      // same association as the reference: coef1 * pred_xstart + coef2 * x_t (gaussian_diffusion.py:1047-1050)
      if (xt) v = c1 * v + c2 * xt[j];
    }
    o[j] = v;
  }
  if (e.out_f32) {
    float* dst = e.out_f32 + (long long)m * e.ld_f32 + n0;
    store8_f32(dst, o, valid, (e.ld_f32 & 3) == 0);
  }
  if (e.out_bf16) {
    __nv_bfloat16 h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __float2bfloat16_rn(o[j]);
    store8_bf16(reinterpret_cast<__nv_bfloat16*>(e.out_bf16) + (long long)m * e.ld_bf16 + n0, h, valid);
    if (e.out_bf16_lo) {
      __nv_bfloat16 l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) l[j] = __float2bfloat16_rn(o[j] - __bfloat162float(h[j]));
      store8_bf16(reinterpret_cast<__nv_bfloat16*>(e.out_bf16_lo) + (long long)m * e.ld_bf16 + n0, l, valid);
    }
  }
}

struct UnitCoord {
  int m_blk, n_blk, split, kb_begin, kb_end;
};
GD_DEV UnitCoord unit_coord(const Shape& s, int unit) {
  UnitCoord u;
  u.m_blk = unit % s.tiles_m;
  int r = unit / s.tiles_m;
  u.n_blk = r % s.tiles_n;
  u.split = r / s.tiles_n;
  u.kb_begin = u.split * s.kb_per_split;
  u.kb_end = min(u.kb_begin + s.kb_per_split, s.total_kb);
  return u;
}

// ---------------------------------------------------------------------------------------------
// Epilogue form (a): row per lane, registers -> swizzled staging -> TMA store. One call per tile per warp.
// ---------------------------------------------------------------------------------------------
template <int BN>
GD_DEV void epilogue_tile_tma(const TmaMaps& maps, const Shape& shape, const gdmcf_epilogue& epi, uint8_t* stg,
                              const float* colvec, uint32_t t_row, int m_warp0, int n_tile0, int lane) {
  const int rows_here = min(32, shape.m - m_warp0);
  const int m_l = m_warp0 + lane;
  const bool row_ok = lane < rows_here;
  const float rs = epi.alpha * ((epi.row_scale && row_ok) ? epi.row_scale[m_l] : 1.0f);
  const int t = (epi.row_t && row_ok) ? epi.row_t[m_l] : epi.t_const;
  const bool has_xt = epi.c1 != nullptr, has_cs = epi.col_scale != nullptr, has_b = epi.bias != nullptr;
  const float c1 = has_xt ? epi.c1[t] : 1.0f, c2 = has_xt ? epi.c2[t] : 0.0f;
  const bool w32 = epi.out_f32 != nullptr, w16 = epi.out_bf16 != nullptr, wlo = epi.out_bf16_lo != nullptr;
  const int act = epi.act;
  const float* xrow = epi.xt + (long long)m_l * epi.ld_xt;  // only dereferenced when has_xt && row_ok
  const uint32_t sw = (uint32_t)(lane & 7);                 // 128B swizzle: 16 B chunk index ^= row % 8
  uint8_t* s32 = stg;                // two [32 rows x 128 B] boxes (columns 0-31, 32-63 of the chunk)
  uint8_t* s16 = stg + 8192;         // [32 rows x 128 B] bf16
  uint8_t* slo = stg + 12288;
#pragma unroll 1
  for (int c = 0; c < BN; c += 64) {
    const int n_c = n_tile0 + c;
    if (n_c >= shape.n || rows_here <= 0) break;  // warp-uniform
    float v[64];
    tmem_ld_32x32(t_row + (uint32_t)c, *reinterpret_cast<float(*)[32]>(v));
    tmem_ld_32x32(t_row + (uint32_t)(c + 32), *reinterpret_cast<float(*)[32]>(v + 32));
    // x_t row segment: 16 independent 16 B loads per lane, issued before the TMEM wait
    float4 x[16];
    if (has_xt) {
#pragma unroll
      for (int q = 0; q < 16; ++q)
        x[q] = (row_ok && n_c + 4 * q < shape.n) ? __ldg(reinterpret_cast<const float4*>(xrow + n_c + 4 * q))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tmem_ld_wait();
    // s = act(alpha*acc*row_scale*col_scale + bias); out = c1*s + c2*x_t      (column vectors: broadcast LDS.128)
    const float4* cs4 = reinterpret_cast<const float4*>(colvec + c);
    const float4* b4 = reinterpret_cast<const float4*>(colvec + BN + c);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float4 y = make_float4(v[4 * q] * rs, v[4 * q + 1] * rs, v[4 * q + 2] * rs, v[4 * q + 3] * rs);
      if (has_cs) { const float4 s = cs4[q]; y.x *= s.x; y.y *= s.y; y.z *= s.z; y.w *= s.w; }
      if (has_b) { const float4 b = b4[q]; y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w; }
      if (act != GDMCF_ACT_NONE) { y.x = apply_act(y.x, act); y.y = apply_act(y.y, act); y.z = apply_act(y.z, act); y.w = apply_act(y.w, act); }
      if (has_xt) {
        y.x = c1 * y.x + c2 * x[q].x; y.y = c1 * y.y + c2 * x[q].y;
        y.z = c1 * y.z + c2 * x[q].z; y.w = c1 * y.w + c2 * x[q].w;
      }
      v[4 * q] = y.x; v[4 * q + 1] = y.y; v[4 * q + 2] = y.z; v[4 * q + 3] = y.w;
    }
    // staging is free once the previous chunk's bulk stores have been read out of shared memory
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (w32) {
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const uint32_t box = (uint32_t)(q >> 3), ch = (uint32_t)(q & 7) ^ sw;
        *reinterpret_cast<float4*>(s32 + box * 4096 + lane * 128 + ch * 16) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    if (w16) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t ch = (uint32_t)q ^ sw;
        uint4 hi;
        hi.x = pack_bf16x2(v[8 * q], v[8 * q + 1]); hi.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
        hi.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]); hi.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
        *reinterpret_cast<uint4*>(s16 + lane * 128 + ch * 16) = hi;
        if (wlo) {
          float r[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = v[8 * q + j] - __bfloat162float(__float2bfloat16_rn(v[8 * q + j]));
          uint4 lo;
          lo.x = pack_bf16x2(r[0], r[1]); lo.y = pack_bf16x2(r[2], r[3]);
          lo.z = pack_bf16x2(r[4], r[5]); lo.w = pack_bf16x2(r[6], r[7]);
          *reinterpret_cast<uint4*>(slo + lane * 128 + ch * 16) = lo;
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && epi.mode != 101) {
      if (w32) {
        tma_store_2d(&maps.o32, s32, n_c, m_warp0);
        if (n_c + 32 < shape.n) tma_store_2d(&maps.o32, s32 + 4096, n_c + 32, m_warp0);
      }
      if (w16) tma_store_2d(&maps.o16, s16, n_c, m_warp0);
      if (wlo) tma_store_2d(&maps.olo, slo, n_c, m_warp0);
      bulk_commit();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Epilogue form (b): smem transpose (lane = row -> lane = column), plain coalesced stores.
// ---------------------------------------------------------------------------------------------
template <int BN>
GD_DEV void epilogue_tile_transposed(const Shape& shape, const gdmcf_epilogue& epi, const UnitCoord& u, float* st,
                                     uint32_t t_row, int m_warp0, int n_tile0, int lane) {
  const int rows_here = min(32, shape.m - m_warp0);  // warp-uniform; <= 0 for fully out-of-range warps
  float4* rowctx = reinterpret_cast<float4*>(st + 32 * 65);  // [32] {alpha*row_scale, c1, c2, t}; 16 B aligned
  {
    const int m_l = m_warp0 + lane;
    const bool row_ok = lane < rows_here;
    const float rs_l = epi.alpha * ((epi.row_scale && row_ok) ? epi.row_scale[m_l] : 1.0f);
    const int t_l = (epi.row_t && row_ok) ? epi.row_t[m_l] : epi.t_const;
    rowctx[lane] = make_float4(rs_l, epi.c1 ? epi.c1[t_l] : 1.0f, epi.c1 ? epi.c2[t_l] : 0.0f, __int_as_float(t_l));
  }
  const bool has_xt = epi.c1 != nullptr, has_tab = epi.bias && epi.ld_bias != 0, has_vec = epi.bias && epi.ld_bias == 0;
  const bool w32 = epi.out_f32 != nullptr, w16 = epi.out_bf16 != nullptr, wlo = epi.out_bf16_lo != nullptr;
  const int act = epi.act;
  const float* __restrict__ bias = epi.bias;
  float* __restrict__ o32 = epi.out_f32;
  __nv_bfloat16* __restrict__ o16 = reinterpret_cast<__nv_bfloat16*>(epi.out_bf16);
  __nv_bfloat16* __restrict__ olo = reinterpret_cast<__nv_bfloat16*>(epi.out_bf16_lo);
  const long long ld_xt = epi.ld_xt, ld_f32 = epi.ld_f32, ld_b16 = epi.ld_bf16, ld_bias = epi.ld_bias;
  __syncwarp();
#pragma unroll 1
  for (int c = 0; c < BN; c += 64) {
    if (n_tile0 + c >= shape.n) break;  // warp-uniform
    float v[64];
    tmem_ld_32x32(t_row + (uint32_t)c, *reinterpret_cast<float(*)[32]>(v));
    tmem_ld_32x32(t_row + (uint32_t)(c + 32), *reinterpret_cast<float(*)[32]>(v + 32));
    tmem_ld_wait();
    if (rows_here > 0 && epi.mode != 100) {  // mode 100/101: timing probes (tools/gemm_case.py), never used by the engine
#pragma unroll
      for (int j = 0; j < 64; ++j) st[lane * 65 + j] = v[j];  // bank (lane + j) % 32: conflict-free
      __syncwarp();
      // lane now owns columns n0 and n1 = n0 + 32 of every row: two fully coalesced 128 B segments per row
      const int n0 = n_tile0 + c + lane, n1 = n0 + 32;
      if (shape.ws) {
        float* dst = shape.ws + (long long)u.split * shape.slab_stride + (long long)m_warp0 * shape.ld_ws;
        const bool k0 = n0 < shape.ld_ws, k1 = n1 < shape.ld_ws;
        for (int r = 0; r < rows_here; ++r) {
          if (k0) dst[(long long)r * shape.ld_ws + n0] = st[r * 65 + lane];
          if (k1) dst[(long long)r * shape.ld_ws + n1] = st[r * 65 + 32 + lane];
        }
      } else {
        const bool ok0 = n0 < shape.n, ok1 = n1 < shape.n;
        const float cs0 = (epi.col_scale && ok0) ? epi.col_scale[n0] : 1.0f;
        const float cs1 = (epi.col_scale && ok1) ? epi.col_scale[n1] : 1.0f;
        const float bv0 = (has_vec && ok0) ? bias[n0] : 0.0f, bv1 = (has_vec && ok1) ? bias[n1] : 0.0f;
        // Rows go in groups of RG (not fully unrolled: the unrolled form did not fit the instruction cache);
        // addresses advance by one leading dimension per row.
        constexpr int RG = 8;
        const float* px = epi.xt + (long long)m_warp0 * ld_xt + n0;
        float* p32 = o32 + (long long)m_warp0 * ld_f32 + n0;
        __nv_bfloat16* p16 = o16 + (long long)m_warp0 * ld_b16 + n0;
        __nv_bfloat16* plo = olo + (long long)m_warp0 * ld_b16 + n0;
        const float* stp = st + lane;
        const bool store_ok = epi.mode != 101;
#pragma unroll 1
        for (int r0 = 0; r0 < rows_here; r0 += RG) {
          float x0[RG], x1[RG];
          if (has_xt) {  // warp-uniform: all reads of the group in flight before the first use
#pragma unroll
            for (int q = 0; q < RG; ++q) {
              const bool rok = r0 + q < rows_here;
              x0[q] = (rok && ok0) ? px[q * ld_xt] : 0.f;
              x1[q] = (rok && ok1) ? px[q * ld_xt + 32] : 0.f;
            }
            px += RG * ld_xt;
          } else {
#pragma unroll
            for (int q = 0; q < RG; ++q) x0[q] = x1[q] = 0.f;
          }
#pragma unroll
          for (int q = 0; q < RG; ++q) {
            const int r = r0 + q;
            if (r < rows_here) {  // warp-uniform
              const float4 rc = rowctx[r];
              float y0 = stp[r * 65] * rc.x * cs0 + bv0;
              float y1 = stp[r * 65 + 32] * rc.x * cs1 + bv1;
              if (has_tab) {
                const float* bt = bias + (long long)__float_as_int(rc.w) * ld_bias;
                if (ok0) y0 += bt[n0];
                if (ok1) y1 += bt[n1];
              }
              if (act != GDMCF_ACT_NONE) { y0 = apply_act(y0, act); y1 = apply_act(y1, act); }
              // coef1 * pred_xstart + coef2 * x_t (gaussian_diffusion.py:1047-1050); c1 = 1, c2 = 0 when absent
              y0 = rc.y * y0 + rc.z * x0[q];
              y1 = rc.y * y1 + rc.z * x1[q];
              if (store_ok) {
                if (w32) {
                  if (ok0) p32[q * ld_f32] = y0;
                  if (ok1) p32[q * ld_f32 + 32] = y1;
                }
                if (w16) {
                  const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
                  if (ok0) p16[q * ld_b16] = h0;
                  if (ok1) p16[q * ld_b16 + 32] = h1;
                  if (wlo) {
                    if (ok0) plo[q * ld_b16] = __float2bfloat16_rn(y0 - __bfloat162float(h0));
                    if (ok1) plo[q * ld_b16 + 32] = __float2bfloat16_rn(y1 - __bfloat162float(h1));
                  }
                }
              }
            }
          }
          p32 += RG * ld_f32;
          p16 += RG * ld_b16;
          plo += RG * ld_b16;
        }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Main kernel
// ---------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ TmaMaps maps, const Shape shape, const gdmcf_epilogue epi) {
  using C = Cfg<BN>;
  pdl_launch_dependents();  // the next kernel may be scheduled; it waits for this grid's completion before touching data
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                        // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;           // [STAGES]
  uint64_t* tfull_bar = bars + 2 * C::STAGES;       // [2]
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  uint8_t* epi_stage = smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES;  // 4 x 16 KB, 1024 B aligned
  float* colvec = reinterpret_cast<float*>(epi_stage + 4 * EPI_WARP_BYTES);  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_units = shape.tiles_m * shape.tiles_n * shape.splits;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < shape.n_seg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptors are set up: everything above overlapped the previous kernel's tail; operands,
  // epilogue vectors and outputs are only touched once the kernels this one depends on have completed
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const UnitCoord u = unit_coord(shape, unit);
        // locate (segment, k-block within segment) of kb_begin
        int seg = 0, kb_in_seg = u.kb_begin;
        while (kb_in_seg >= shape.kb[seg]) { kb_in_seg -= shape.kb[seg]; ++seg; }
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_2d(sa, &maps.a[seg], &full_bar[stage], kb_in_seg * BK, u.m_blk * BM);
          tma_load_2d(sb, &maps.b[seg], &full_bar[stage], kb_in_seg * BK, u.n_blk * BN);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          if (++kb_in_seg == shape.kb[seg]) { kb_in_seg = 0; ++seg; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local) {
        const UnitCoord u = unit_coord(shape, unit);
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 B (16 bf16) inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                      (kb > u.kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================= epilogue =================
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127 among the epilogue threads
    uint8_t* stg = epi_stage + ew * EPI_WARP_BYTES;
    int local = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local) {
      const UnitCoord u = unit_coord(shape, unit);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m_warp0 = u.m_blk * BM + ew * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      const int n_tile0 = u.n_blk * BN;
      if (shape.tma_store) {
        // column vectors of this tile -> smem (the 4 epilogue warps only: named barrier 1, 128 threads)
        named_bar_sync(1, 128);  // previous tile's readers are done
        for (int j = et; j < BN; j += 128) {
          const int n = n_tile0 + j;
          colvec[j] = (epi.col_scale && n < shape.n) ? epi.col_scale[n] : 1.0f;
          colvec[BN + j] = (epi.bias && n < shape.n) ? epi.bias[n] : 0.0f;
        }
        named_bar_sync(1, 128);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (shape.tma_store)
        epilogue_tile_tma<BN>(maps, shape, epi, stg, colvec, t_row, m_warp0, n_tile0, lane);
      else
        epilogue_tile_transposed<BN>(shape, epi, u, reinterpret_cast<float*>(stg), t_row, m_warp0, n_tile0, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (shape.tma_store && lane == 0) bulk_wait_all();  // bulk stores must have landed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): the two CTAs of a cluster compute one 256 x BN tile. Each CTA loads its 128 rows of A
// and HALF of the B tile (BN/2 rows), so a k-block costs 32 KB of L2 -> SM traffic per CTA instead of 48 KB — the
// 1-CTA kernel is bound by exactly that traffic on the batch-sized (M = 400) contractions. The leader (cluster rank 0)
// issues tcgen05.mma.cta_group::2 for both; accumulator rows live in each CTA's own TMEM, so the epilogue is the
// 1-CTA epilogue on the CTA's own 128 rows.
//   full_bar  (leader's only): one arrive.expect_tx by the leader's producer for the bytes of BOTH CTAs; the peer's TMA
//                              loads signal the leader's barrier (cp.async.bulk.tensor.cta_group::2).
//   empty_bar (each CTA)     : tcgen05.commit multicast from the leader once the MMAs that read the stage retired.
//   tfull_bar (each CTA)     : tcgen05.commit multicast — accumulator complete.
//   tempty_bar (leader's)    : 8 arrivals = 4 epilogue warps x 2 CTAs (the peer's arrive through the cluster).
// ---------------------------------------------------------------------------------------------
struct Cfg2 {
  static constexpr int BN = 256;
  static constexpr int STAGES = 5;  // 5 x 32 KB in flight; with the 64 KB epilogue staging this is exactly the 227 KB limit
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int COLVEC_BYTES = 2 * BN * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 4 * EPI_WARP_BYTES + COLVEC_BYTES;
};

// unit -> (pair of m-blocks, n-block, split); m-pair fastest so neighbouring clusters share the weight tile in L2
GD_DEV UnitCoord unit_coord_pair(const Shape& s, int unit, int pairs_m, int rank) {
  UnitCoord u;
  u.m_blk = 2 * (unit % pairs_m) + rank;
  int r = unit / pairs_m;
  u.n_blk = r % s.tiles_n;
  u.split = r / s.tiles_n;
  u.kb_begin = u.split * s.kb_per_split;
  u.kb_end = min(u.kb_begin + s.kb_per_split, s.total_kb);
  return u;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ TmaMaps maps, const Shape shape, const gdmcf_epilogue epi) {
  using C = Cfg2;
  constexpr int BN = C::BN;
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tfull_bar = bars + 2 * C::STAGES;
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  uint8_t* epi_stage = smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES;
  float* colvec = reinterpret_cast<float*>(epi_stage + 4 * EPI_WARP_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pairs_m = (shape.tiles_m + 1) / 2;
  const int num_units = pairs_m * shape.tiles_n * shape.splits;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < shape.n_seg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    mbar_init_fence();
  }
  if (warp == 2) tmem_alloc_2cta<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers and TMEM exist before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptors are set up: everything above overlapped the previous kernel's tail; operands,
  // epilogue vectors and outputs are only touched once the kernels this one depends on have completed
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const UnitCoord u = unit_coord_pair(shape, unit, pairs_m, (int)rank);
        int seg = 0, kb_in_seg = u.kb_begin;
        while (kb_in_seg >= shape.kb[seg]) { kb_in_seg -= shape.kb[seg]; ++seg; }
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);  // the leader's barrier
          tma_load_2d_2cta(sa, &maps.a[seg], fb, kb_in_seg * BK, u.m_blk * BM);
          tma_load_2d_2cta(sb, &maps.b[seg], fb, kb_in_seg * BK, u.n_blk * BN + (int)rank * (BN / 2));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          if (++kb_in_seg == shape.kb[seg]) { kb_in_seg = 0; ++seg; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader only) =================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++local) {
        const UnitCoord u = unit_coord_pair(shape, unit, pairs_m, 0);
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_2cta(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > u.kb_begin || k > 0) ? 1u : 0u);
          umma_commit_2cta(&empty_bar[stage], 3);  // frees the stage in both CTAs
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(&tfull_bar[acc], 3);  // accumulator complete -> both epilogues
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================= epilogue (both CTAs, own 128 rows) =================
    const int ew = warp & 3;
    const int et = threadIdx.x - EPI_WARP0 * 32;
    uint8_t* stg = epi_stage + ew * EPI_WARP_BYTES;
    int local = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++local) {
      const UnitCoord u = unit_coord_pair(shape, unit, pairs_m, (int)rank);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m_warp0 = u.m_blk * BM + ew * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      const int n_tile0 = u.n_blk * BN;
      if (shape.tma_store) {
        named_bar_sync(1, 128);
        for (int j = et; j < BN; j += 128) {
          const int n = n_tile0 + j;
          colvec[j] = (epi.col_scale && n < shape.n) ? epi.col_scale[n] : 1.0f;
          colvec[BN + j] = (epi.bias && n < shape.n) ? epi.bias[n] : 0.0f;
        }
        named_bar_sync(1, 128);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (shape.tma_store)
        epilogue_tile_tma<BN>(maps, shape, epi, stg, colvec, t_row, m_warp0, n_tile0, lane);
      else
        epilogue_tile_transposed<BN>(shape, epi, u, reinterpret_cast<float*>(stg), t_row, m_warp0, n_tile0, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
    }
    if (shape.tma_store && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be reading its accumulators / receiving multicast arrivals
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base);
  }
}

// Sum split-K slabs in fixed order and apply the epilogue. One thread per 8 columns.
__global__ void splitk_reduce_kernel(const Shape shape, const gdmcf_epilogue epi) {
  pdl_entry();
  const int groups = (shape.n + 7) / 8;
  const long long total = (long long)shape.m * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / groups);
    const int n0 = (int)(i % groups) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* src = shape.ws + (long long)m * shape.ld_ws + n0;
    for (int s = 0; s < shape.splits; ++s) {
      const float4 x = reinterpret_cast<const float4*>(src)[0];
      const float4 y = reinterpret_cast<const float4*>(src)[1];
      acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
      acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
      src += shape.slab_stride;
    }
    epilogue8(epi, m, n0, shape.n, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major [rows, cols] tensor with leading dim ld (elements of `esize` bytes); box = [box_rows, box_cols]
// with box_cols * esize == 128 B, 128B swizzle. Loads zero-fill out-of-bounds elements, stores clip them
// (K tails and M/N tails need no special casing).
static int make_map(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, int rows, int cols, long long ld,
                    int box_rows, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return GDMCF_ECUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (rows=%d cols=%d ld=%lld ptr=%p) -> %d", rows, cols, ld, ptr, (int)r);
    return GDMCF_EBADARG;
  }
  return GDMCF_OK;
}

static int pick_bn(int n) { return n <= 128 ? 128 : 256; }

template <int BN>
static int launch(const TmaMaps& maps, const Shape& shape, const gdmcf_epilogue& epi, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<BN>::SMEM_BYTES);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(gemm)");
    attr_set = true;
  }
  launch_kernel(gemm_bf16_tn_kernel<BN>, grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, st, maps, shape, epi);
  return cuda_check_launch("gemm_bf16_tn_kernel");
}

static int launch_2cta(const TmaMaps& maps, const Shape& shape, const gdmcf_epilogue& epi, int clusters, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(gemm 2cta)");
    attr_set = true;
  }
  launch_kernel(gemm_bf16_tn_2cta_kernel, 2 * clusters, NUM_THREADS, Cfg2::SMEM_BYTES, st, maps, shape, epi);
  return cuda_check_launch("gemm_bf16_tn_2cta_kernel");
}

// Upper bound on the SMs a contraction may occupy (0 = all). A data-parallel caller lowers it while gradient all-reduces
// are in flight: the persistent kernels assume that all their CTAs run at once, so NCCL's CTAs must find free SMs instead
// of delaying a few tiles' owners (which would double the kernel's duration).
static int g_sm_limit = 0;
static int sm_budget() {
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  return (g_sm_limit > 0 && g_sm_limit < sms) ? g_sm_limit : sms;
}

// GDMCF_GEMM_2CTA=0 routes every contraction through the 1-CTA kernel (A/B comparisons, fallback).
static bool use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GDMCF_GEMM_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace gemm
}  // namespace gd

using namespace gd;
using namespace gd::gemm;

extern "C" int gdmcf_gemm_set_sm_limit(int sms) {
  if (sms < 0) { set_error("gemm_set_sm_limit: negative"); return GDMCF_EBADARG; }
  g_sm_limit = sms;
  return GDMCF_OK;
}

extern "C" int gdmcf_gemm_auto_splits(int m, int n, int k_total) {
  if (m <= 0 || n <= 0 || k_total <= 0) return 1;
  const int bn = pick_bn(n);
  const int tiles = ((m + BM - 1) / BM) * ((n + bn - 1) / bn);
  const int total_kb = (k_total + BK - 1) / BK;
  const int sms = sm_budget();
  if (bn == 256 && use_2cta()) {
    // CTA-pair kernel: work units are (pair of m-blocks, n-block, split) walked by sms / 2 persistent clusters. Pick the
    // split count that fills whole waves best (a 1000 x 3000 x 34395 product is 48 units: 3 splits = 144 units on 74
    // clusters instead of 48), among those that keep >= min_kb k-blocks per split; ties go to fewer splits.
    const int clusters = std::max(1, sms / 2);
    const int units = (((m + BM - 1) / BM + 1) / 2) * ((n + bn - 1) / bn);
    static int min_kb2 = 0;
    if (min_kb2 == 0) {
      const char* e = getenv("GDMCF_GEMM_MIN_KB");
      min_kb2 = (e && atoi(e) > 0) ? atoi(e) : 8;
    }
    const int max_splits = std::max(1, std::min(16, total_kb / min_kb2));
    int best = 1;
    double best_eff = 0.0;
    for (int sp = 1; sp <= max_splits; ++sp) {
      const long long u = (long long)units * sp;
      const long long waves = (u + clusters - 1) / clusters;
      // useful fraction of the cluster-waves, minus ~2 % per extra split for its slab epilogue + reduce pass
      const double eff = (double)u / (double)(waves * clusters) - 0.02 * (sp - 1);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
    }
    if (units >= clusters && best_eff < (double)units / (double)(((units + clusters - 1) / clusters) * clusters) + 0.05) best = 1;
    const int per2 = (total_kb + best - 1) / best;
    return (total_kb + per2 - 1) / per2;
  }
  if (tiles >= sms) return 1;
  int splits = sms / tiles;
  // keep at least GDMCF_GEMM_MIN_KB (default 8) k-blocks per split: a split costs a partial-slab epilogue (the slow
  // transposed form) plus a share of the reduce kernel, which only pays off against a long enough main loop
  static int min_kb = 0;
  if (min_kb == 0) {
    const char* e = getenv("GDMCF_GEMM_MIN_KB");
    min_kb = (e && atoi(e) > 0) ? atoi(e) : 8;
  }
  splits = std::min(splits, std::max(1, total_kb / min_kb));
  splits = std::max(splits, 1);
  const int per = (total_kb + splits - 1) / splits;
  return (total_kb + per - 1) / per;
}

extern "C" size_t gdmcf_gemm_workspace_bytes(int m, int n, int splits) {
  if (splits <= 1) return 0;
  const size_t ld = ((size_t)n + 31) / 32 * 32;
  return (size_t)splits * (size_t)m * ld * sizeof(float);
}

extern "C" int gdmcf_gemm_bf16_tn(const gdmcf_gemm_desc* g, const gdmcf_epilogue* e, int splits, void* workspace,
                                  size_t workspace_bytes, gdmcf_stream_t stream) {
  if (!g || !e) { set_error("gemm: null descriptor"); return GDMCF_EBADARG; }
  if (g->m <= 0 || g->n <= 0 || g->n_seg < 1 || g->n_seg > GDMCF_MAX_SEG) {
    set_error("gemm: bad shape m=%d n=%d n_seg=%d", g->m, g->n, g->n_seg);
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  const int bn = pick_bn(g->n);
  const bool pair = bn == 256 && use_2cta() && e->mode != 100 && e->mode != 101;
  Shape shape{};
  shape.m = g->m;
  shape.n = g->n;
  shape.n_seg = g->n_seg;
  shape.total_kb = 0;
  TmaMaps maps;
  for (int s = 0; s < g->n_seg; ++s) {
    if (g->k[s] <= 0 || (g->lda[s] & 7) || (g->ldb[s] & 7) || ((uintptr_t)g->a[s] & 15) || ((uintptr_t)g->b[s] & 15) ||
        g->lda[s] < g->k[s] || g->ldb[s] < g->k[s]) {
      set_error("gemm: segment %d needs k>0, ld%%8==0, ld>=k, 16B-aligned pointers (k=%d lda=%lld ldb=%lld)", s,
                g->k[s], (long long)g->lda[s], (long long)g->ldb[s]);
      return GDMCF_EBADARG;
    }
    shape.kb[s] = (g->k[s] + BK - 1) / BK;
    shape.total_kb += shape.kb[s];
    if ((rc = make_map(&maps.a[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->a[s], g->m, g->k[s], g->lda[s], BM, BK))) return rc;
    if ((rc = make_map(&maps.b[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g->b[s], g->n, g->k[s], g->ldb[s], pair ? bn / 2 : bn, BK))) return rc;
  }
  if (e->mode < 0 || (e->mode > GDMCF_EPI_COSINE && e->mode != 100 && e->mode != 101)) { set_error("gemm: bad epilogue mode %d", e->mode); return GDMCF_EBADARG; }
  if (e->c1 && (!e->c2 || !e->xt)) {
    set_error("gemm: the posterior-mean epilogue needs c1, c2 and xt together");
    return GDMCF_EBADARG;
  }
  if ((e->out_bf16 && ((e->ld_bf16 & 7) || ((uintptr_t)e->out_bf16 & 15))) || (e->out_bf16_lo && ((uintptr_t)e->out_bf16_lo & 15)) ||
      (e->out_f32 && ((uintptr_t)e->out_f32 & 15)) || (!e->out_f32 && !e->out_bf16) || (e->out_bf16_lo && !e->out_bf16)) {
    set_error("gemm: epilogue outputs need 16B alignment, ld_bf16%%8==0, and at least one output");
    return GDMCF_EBADARG;
  }
  shape.tiles_m = (g->m + BM - 1) / BM;
  shape.tiles_n = (g->n + bn - 1) / bn;
  splits = std::max(1, std::min(splits, shape.total_kb));
  shape.kb_per_split = (shape.total_kb + splits - 1) / splits;
  shape.splits = (shape.total_kb + shape.kb_per_split - 1) / shape.kb_per_split;
  shape.ws = nullptr;
  if (shape.splits > 1) {
    const size_t need = gdmcf_gemm_workspace_bytes(g->m, g->n, shape.splits);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 15)) {
      set_error("gemm: split-K workspace too small (%zu < %zu) or misaligned", workspace_bytes, need);
      return GDMCF_EBADARG;
    }
    shape.ws = reinterpret_cast<float*>(workspace);
    shape.ld_ws = (g->n + 31) / 32 * 32;
    shape.slab_stride = (long long)g->m * shape.ld_ws;
  }
  // Epilogue form (a) needs 16 B global strides on every output and on x_t, a plain (or no) bias vector, and no split-K.
  // The bulk tensor store clips at 16 B granularity (measured: with n = 1111 fp32 columns, column 1111 is written too),
  // so the columns [n, round_up(n, 4 | 8)) of every output row must be inside its leading dimension; they receive the
  // epilogue of a zero accumulator.
  const long long n4 = ((long long)g->n + 3) / 4 * 4, n8 = ((long long)g->n + 7) / 8 * 8;
  shape.tma_store = shape.splits == 1 && e->mode != 100 && (!e->bias || e->ld_bias == 0) &&
                    (!e->out_f32 || ((e->ld_f32 & 3) == 0 && e->ld_f32 >= n4)) && (!e->out_bf16 || e->ld_bf16 >= n8) &&
                    (!e->c1 || ((e->ld_xt & 3) == 0 && e->ld_xt >= n4 && ((uintptr_t)e->xt & 15) == 0));
  if (shape.tma_store) {
    if (e->out_f32 && (rc = make_map(&maps.o32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, e->out_f32, g->m, g->n, e->ld_f32, 32, 32))) return rc;
    if (e->out_bf16 && (rc = make_map(&maps.o16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, e->out_bf16, g->m, g->n, e->ld_bf16, 32, 64))) return rc;
    if (e->out_bf16_lo && (rc = make_map(&maps.olo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, e->out_bf16_lo, g->m, g->n, e->ld_bf16, 32, 64))) return rc;
  }
  const int sms = sm_budget();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pair) {
    const int units = ((shape.tiles_m + 1) / 2) * shape.tiles_n * shape.splits;
    rc = launch_2cta(maps, shape, *e, std::max(1, std::min(units, sms / 2)), st);
  } else {
    const int num_units = shape.tiles_m * shape.tiles_n * shape.splits;
    const int grid = std::min(num_units, sms);
    rc = (bn == 256) ? launch<256>(maps, shape, *e, grid, st) : launch<128>(maps, shape, *e, grid, st);
  }
  if (rc) return rc;
  if (shape.splits > 1) {
    const long long total = (long long)g->m * ((g->n + 7) / 8);
    const int threads = 256;
    const int blocks = (int)std::min<long long>((total + threads - 1) / threads, (long long)(sms > 0 ? sms : 148) * 8);
    launch_kernel(splitk_reduce_kernel, blocks, threads, 0, st, shape, *e);
    rc = cuda_check_launch("splitk_reduce_kernel");
  }
  return rc;
}

// LightGCN propagation (lightGCN.py:180-194), bf16 mode: all K layers in ONE persistent launch.
//
//   out = mean_{k<=K} (A~^k E0),  A~ = D^-1/2 A D^-1/2,  A binary (pattern only), E0 fp32 [n, 64] in / out fp32.
//
// What changes against the fp32 kernels of spmm.cu (7 launches, every neighbour row a 256 B gather from L2):
//  * the iterated tables U_k = D^-1/2 T_k are kept as bf16 rows (128 B per neighbour gather instead of 256 B), sums are
//    accumulated in fp32 and the E0 term of every Horner step is re-read in fp32, so only the neighbour sums see bf16
//    rounding (measured ~1e-3 normwise on the layer mean; the fp32 path stays for the 1e-5 mode);
//  * shared-memory staging of embedding tiles: the HOT_ROWS most frequently gathered rows (Zipf-head items carry ~2/3 of
//    the user-row half's non-zeros) are copied into a 128 KB shared-memory tile per CTA at the start of every layer; the
//    plan reorders each row's neighbour list hot-first and encodes a hot neighbour as 0x80000000 | slot, so ~1/3 of all
//    gathers never leave the SM;
//  * warp-level segmented reduction: lane groups of 8 gather one 128 B neighbour row per load instruction (8 x 16 B), 4
//    loads in flight per lane; a piece of a hub row is walked by a whole warp (4 groups), ordinary rows by half-warps (2
//    groups each, two rows per warp in flight), and the groups' partial sums are folded with shuffle steps; hub-row pieces go to a scratch slab and the LAST piece to
//    finish adds them in slab order (deterministic; no floating-point atomics, no second kernel);
//  * one resident wave of CTAs (one per SM, 32 warps) walks phase 0 (U_0 = bf16(dinv * E0)) and the K layers, separated by
//    grid barriers: 1 launch instead of 7.
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace spmm16 {

constexpr int D = 64;               // latent_dim of the reference script (lightGCN.py:205)
constexpr int HOT_ROWS = 1024;      // x 128 B = 128 KB shared memory
constexpr int THREADS = 1024;
constexpr int WARPS = THREADS / 32;
constexpr int SMEM_BYTES = (HOT_ROWS + 1) * D * 2;  // + one all-zero slot
constexpr int MAX_BARRIERS = 12;
constexpr int WORK_CTR0 = 16;       // sync[16 .. 28): per-layer work counters (dynamic distribution of the items)
constexpr int HUB_CTR0 = 32;        // sync[32 + i]: pieces of hub row i finished

struct Params {
  const int* col;          // neighbour lists, hot-first; hot neighbours are 0x80000000 | slot
  const int4* items;       // {row or -(long_idx + 1), begin, end, slot or -1}
  const int* mids;         // [n_items] end of the item's hot prefix (begin <= mid <= end)
  int n_items, n_pieces;   // the first n_pieces items are pieces of hub rows (slot >= 0), the others whole rows
  const int* warp_ptr;     // [grid * WARPS + 1] whole-row items of warp slot w: items[n_pieces + warp_ptr[w] .. n_pieces + warp_ptr[w + 1])
  const int* long_rows;    // {row, first_slot, n_slots}
  int n_long;
  const int* hot_rows;     // [n_hot] row ids staged into shared memory
  int n_hot;
  const float* dinv;
  const float* E0;
  __nv_bfloat16* u[2];     // ping-pong tables [n + 1, 64] bf16, row n all zero
  float* out;
  float* scratch;          // [n_slots, 64] fp32
  unsigned int* sync;      // [0, 12) barrier counters, [12] exit counter, [16, 28) work counters, [32 + i] hub-row counters,
                           // then 32 words of phase timestamps (ns since entry) of the first and the last CTA (diagnostics)
  int n, n_layers;
};

GD_DEV unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

GD_DEV void grid_barrier(unsigned int* ctr, unsigned int expected) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ctr, 1u);
    unsigned int spins = 0;
    uint64_t t0 = 0;
    while (ld_acquire(ctr) < expected) {
      __nanosleep(32);
      if (++spins == 1024u) t0 = globaltimer_ns();
      if (spins > 1024u && (spins & 255u) == 0u && globaltimer_ns() - t0 > 4000000000ull) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

GD_DEV void add_bf16x8(float (&acc)[8], const uint4& v) {
  acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xffff0000u);
  acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xffff0000u);
  acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xffff0000u);
  acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xffff0000u);
}
GD_DEV uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
GD_DEV uint32_t pack2(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

// Row epilogue by lanes 0..7 (lane l holds columns 8l .. 8l+7 of the neighbour sum; e = E0[row, same columns], dv = dinv[row]).
//   inner layers: U_{k+1}[r] = dinv^2 * sum + dinv * E0[r]   -> bf16
//   last layer  : out[r]     = (dinv * sum + E0[r]) / (K + 1) -> fp32
GD_DEV void finish_row(const Params& p, int row, int sub, const float (&acc)[8], const float4& e0, const float4& e1, float dv,
                       bool last, __nv_bfloat16* dst, float inv_layers) {
  const float e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
  float o[8];
  if (last) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (dv * acc[j] + e[j]) * inv_layers;
    float4* d4 = reinterpret_cast<float4*>(p.out + (long long)row * D + sub * 8);
    d4[0] = make_float4(o[0], o[1], o[2], o[3]);
    d4[1] = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    const float dv2 = dv * dv;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = dv2 * acc[j] + dv * e[j];
    *reinterpret_cast<uint4*>(dst + (long long)row * D + sub * 8) =
        make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
  }
}

__global__ void __launch_bounds__(THREADS, 1)
lightgcn_bf16_kernel(const Params p) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t hot[];  // [HOT_ROWS + 1][128 B]
  const uint32_t hot_base = smem_u32(hot);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 3, sub = lane & 7;      // 4 neighbour rows per load instruction, 8 lanes x 16 B per row
  const float inv_layers = 1.0f / (float)(p.n_layers + 1);
  unsigned int bar = 0;
  const uint64_t t_start = globaltimer_ns();
  unsigned int* dbg = nullptr;
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    dbg = p.sync + HUB_CTR0 + p.n_long + 1 + (blockIdx.x == 0 ? 0 : 16);
  int dbg_i = 0;
#define GD_STAMP() do { if (dbg && dbg_i < 16) dbg[dbg_i++] = (unsigned int)(globaltimer_ns() - t_start); } while (0)

  // ---- phase 0: U_0 = bf16(dinv * E0); 8 lanes per row
  {
    const long long chunks = (long long)p.n * 8;
    for (long long i = blockIdx.x * (long long)THREADS + threadIdx.x; i < chunks; i += (long long)gridDim.x * THREADS) {
      const int r = (int)(i >> 3), s = (int)(i & 7);
      const float dv = p.dinv[r];
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)r * D + s * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)r * D + s * 8 + 4));
      *reinterpret_cast<uint4*>(p.u[0] + (long long)r * D + s * 8) =
          make_uint4(pack2(dv * a.x, dv * a.y), pack2(dv * a.z, dv * a.w), pack2(dv * b.x, dv * b.y), pack2(dv * b.z, dv * b.w));
    }
  }
  GD_STAMP();
  grid_barrier(&p.sync[bar++], gridDim.x);
  GD_STAMP();

  for (int layer = 0; layer < p.n_layers; ++layer) {
    const bool last = layer == p.n_layers - 1;
    const __nv_bfloat16* src = p.u[layer & 1];
    __nv_bfloat16* dst = p.u[(layer + 1) & 1];
    // ---- stage the hot rows of this layer's table: 8 threads x 16 B per row; slot n_hot stays all zero (padding lanes)
    for (int i = threadIdx.x; i < (p.n_hot + 1) * 8; i += THREADS) {
      const int h = i >> 3, s = i & 7;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (h < p.n_hot) v = *reinterpret_cast<const uint4*>(src + (long long)__ldg(p.hot_rows + h) * D + s * 8);
      *reinterpret_cast<uint4*>(hot + h * 128 + s * 16) = v;
    }
    __syncthreads();
    GD_STAMP();

    // ---- items: one warp each, dealt round-robin (heavy hub pieces come first in the plan). A neighbour list is hot-first:
    // [begin, mid) are shared-memory slots, [mid, end) global row ids, so each part runs its own loop with one kind of load.
    // The next item's descriptor and this row's own E0 / dinv (epilogue operands) are requested before the gathers.
    // (a) pieces of hub rows — the first n_pieces items of the plan — one per warp
    const int stride = gridDim.x * WARPS;
    int item = blockIdx.x * WARPS + warp;
    int4 it = make_int4(0, 0, 0, -1);
    int mid = 0;
    if (item < p.n_pieces) { it = __ldg(&p.items[item]); mid = __ldg(p.mids + item); }
    for (; item < p.n_pieces; item += stride) {
      int4 it_n = make_int4(0, 0, 0, -1);
      int mid_n = 0;
      if (item + stride < p.n_pieces) { it_n = __ldg(&p.items[item + stride]); mid_n = __ldg(p.mids + item + stride); }
      const bool whole = it.w < 0;
      float dv = 0.f;
      float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
      if (whole && lane < 8) {
        dv = __ldg(p.dinv + it.x);
        e0 = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)it.x * D + sub * 8));
        e1 = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)it.x * D + sub * 8 + 4));
      }
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int base = it.y; base < mid; base += 32) {      // hot neighbours: staged rows
        const int n = min(32, mid - base);
        const int cc = lane < n ? (__ldg(p.col + base + lane) & 0x7fffffff) : p.n_hot;  // padding lanes read the zero slot
#pragma unroll 1
        for (int j0 = 0; j0 < n; j0 += 16) {
          uint4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = __shfl_sync(0xffffffffu, cc, j0 + 4 * u + grp);
            v[u] = lds128(hot_base + (uint32_t)c * 128u + (uint32_t)sub * 16u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) add_bf16x8(acc, v[u]);
        }
      }
      for (int base = mid; base < it.z; base += 32) {      // the others: 128 B rows from L2
        const int n = min(32, it.z - base);
        const int cc = lane < n ? __ldg(p.col + base + lane) : p.n;  // padding lanes gather the all-zero row n
#pragma unroll 1
        for (int j0 = 0; j0 < n; j0 += 16) {  // 4 loads in flight per lane = 16 neighbour rows per warp per trip
          uint4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = __shfl_sync(0xffffffffu, cc, j0 + 4 * u + grp);
            v[u] = __ldg(reinterpret_cast<const uint4*>(src + (long long)c * D + sub * 8));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) add_bf16x8(acc, v[u]);
        }
      }
      // fold the 4 lane groups: lanes 0..7 end up with the sum over all neighbours
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
      }
      if (whole) {
        if (lane < 8) finish_row(p, it.x, sub, acc, e0, e1, dv, last, dst, inv_layers);
      } else {
        // piece of a hub row: partial sum -> scratch slab; the last piece to arrive reduces the row in slab order
        if (lane < 8) {
          float4* s4 = reinterpret_cast<float4*>(p.scratch + (long long)it.w * D + sub * 8);
          s4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          s4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        __threadfence();
        __syncwarp();
        const int li = -it.x - 1;
        const int row = p.long_rows[3 * li], first = p.long_rows[3 * li + 1], cnt = p.long_rows[3 * li + 2];
        unsigned int done = 0;
        if (lane == 0) {
          done = atomicAdd(&p.sync[HUB_CTR0 + li], 1u);
          if (done + 1u == (unsigned int)cnt) p.sync[HUB_CTR0 + li] = 0u;  // leave the counter zeroed for the next layer / launch
          __threadfence();
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done + 1u == (unsigned int)cnt) {
          // the 4 lane groups sum every 4th slab (4 independent 32 B loads in flight per lane), then the groups are folded in
          // a fixed order: a hub row with hundreds of pieces costs ~cnt/16 L2 round trips on one warp instead of cnt
          float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int s0 = grp; s0 < cnt; s0 += 16) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int sidx = s0 + 4 * u;
              const float* src_s = p.scratch + (long long)(first + min(sidx, cnt - 1)) * D + sub * 8;
              a[u] = __ldcg(reinterpret_cast<const float4*>(src_s));
              b[u] = __ldcg(reinterpret_cast<const float4*>(src_s + 4));
              if (sidx >= cnt) { a[u] = make_float4(0.f, 0.f, 0.f, 0.f); b[u] = a[u]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              tot[0] += a[u].x; tot[1] += a[u].y; tot[2] += a[u].z; tot[3] += a[u].w;
              tot[4] += b[u].x; tot[5] += b[u].y; tot[6] += b[u].z; tot[7] += b[u].w;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            tot[j] += __shfl_xor_sync(0xffffffffu, tot[j], 8);
            tot[j] += __shfl_xor_sync(0xffffffffu, tot[j], 16);
          }
          if (lane < 8) {
            const float rdv = __ldg(p.dinv + row);
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)row * D + sub * 8));
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)row * D + sub * 8 + 4));
            finish_row(p, row, sub, tot, r0, r1, rdv, last, dst, inv_layers);
          }
        }
      }
      it = it_n;
      mid = mid_n;
    }
    if (layer == 0 && threadIdx.x == 0) p.sync[HUB_CTR0 + p.n_long + 1 + 32 + 2 * blockIdx.x] = (unsigned int)(globaltimer_ns() - t_start);
    // (b) whole rows — two per warp: each half-warp (16 lanes = 2 lane groups of 8) walks its own row, so a warp keeps two
    // dependent chains (descriptor -> neighbour ids -> rows -> epilogue) in flight with the registers of one. Trip counts are
    // made warp-uniform (max over the two halves); the shorter row's extra trips read the all-zero slot / row. Which rows a
    // warp walks is fixed by the plan (warp_ptr): pairs of equal cost, dealt so that every warp of the grid carries the same
    // estimated work including its hub pieces — the layer ends at the slowest warp.
    {
      const int l16 = lane & 15, g2 = l16 >> 3;
      const int wslot = blockIdx.x * WARPS + warp;
      const int rend = p.n_pieces + __ldg(p.warp_ptr + wslot + 1);
      int ritem = p.n_pieces + __ldg(p.warp_ptr + wslot) + (lane >> 4);
      bool have = ritem < rend;
      int4 rit = make_int4(0, 0, 0, -1);
      int rmid = 0;
      if (have) { rit = __ldg(&p.items[ritem]); rmid = __ldg(p.mids + ritem); }
      while (__any_sync(0xffffffffu, have)) {
        const int ritem_n = ritem + 2;
        const bool have_n = ritem_n < rend;
        int4 rit_n = make_int4(0, 0, 0, -1);
        int rmid_n = 0;
        if (have_n) { rit_n = __ldg(&p.items[ritem_n]); rmid_n = __ldg(p.mids + ritem_n); }
        float dv = 0.f;
        float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
        if (have && l16 < 8) {
          dv = __ldg(p.dinv + rit.x);
          e0 = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)rit.x * D + sub * 8));
          e1 = __ldg(reinterpret_cast<const float4*>(p.E0 + (long long)rit.x * D + sub * 8 + 4));
        }
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int nh = have ? rmid - rit.y : 0, nc = have ? rit.z - rmid : 0;
        const int nh_max = max(nh, __shfl_xor_sync(0xffffffffu, nh, 16));
        const int nc_max = max(nc, __shfl_xor_sync(0xffffffffu, nc, 16));
        for (int b = 0; b < nh_max; b += 16) {              // hot neighbours: staged rows
          const int n = max(0, min(16, nh - b));
          const int n_u = min(16, nh_max - b);
          const int cc = l16 < n ? (__ldg(p.col + rit.y + b + l16) & 0x7fffffff) : p.n_hot;
#pragma unroll 1
          for (int j0 = 0; j0 < n_u; j0 += 8) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c = __shfl_sync(0xffffffffu, cc, j0 + 2 * u + g2, 16);
              v[u] = lds128(hot_base + (uint32_t)c * 128u + (uint32_t)sub * 16u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) add_bf16x8(acc, v[u]);
          }
        }
        for (int b = 0; b < nc_max; b += 16) {              // the others: 128 B rows from L2
          const int n = max(0, min(16, nc - b));
          const int n_u = min(16, nc_max - b);
          const int cc = l16 < n ? __ldg(p.col + rmid + b + l16) : p.n;
#pragma unroll 1
          for (int j0 = 0; j0 < n_u; j0 += 8) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c = __shfl_sync(0xffffffffu, cc, j0 + 2 * u + g2, 16);
              v[u] = __ldg(reinterpret_cast<const uint4*>(src + (long long)c * D + sub * 8));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) add_bf16x8(acc, v[u]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);  // fold the half's two lane groups
        if (have && l16 < 8) finish_row(p, rit.x, sub, acc, e0, e1, dv, last, dst, inv_layers);
        ritem = ritem_n;
        have = have_n;
        rit = rit_n;
        rmid = rmid_n;
      }
    }
    __syncthreads();
    if (layer == 0 && threadIdx.x == 0) p.sync[HUB_CTR0 + p.n_long + 1 + 32 + 2 * blockIdx.x + 1] = (unsigned int)(globaltimer_ns() - t_start);
    GD_STAMP();
    if (!last) grid_barrier(&p.sync[bar++], gridDim.x);
    GD_STAMP();
  }

  // leave the barrier / work counters zeroed for the next launch: the last CTA to get here resets them
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int n = atomicAdd(&p.sync[MAX_BARRIERS], 1u);
    if (n + 1u == gridDim.x) {
      for (int i = 0; i < HUB_CTR0; ++i) p.sync[i] = 0u;
      __threadfence();
    }
  }
}

}  // namespace spmm16
}  // namespace gd

using namespace gd;
using namespace gd::spmm16;

extern "C" int gdmcf_lightgcn_hot_rows(void) { return HOT_ROWS; }

extern "C" int gdmcf_lightgcn_propagate_bf16(const int32_t* col_hot_first, const int32_t* items, const int32_t* item_mids, int n_items,
                                             int n_pieces, const int32_t* warp_ptr, int n_warp_slots,
                                             const int32_t* long_rows, int n_long, const int32_t* hot_rows, int n_hot,
                                             const float* dinv, const float* E0, void* u0_bf16, void* u1_bf16, float* out,
                                             float* scratch, uint32_t* sync_block, int n, int d, int n_layers,
                                             gdmcf_stream_t stream) {
  if (!col_hot_first || !items || !item_mids || !warp_ptr || n_warp_slots < WARPS || n_warp_slots % WARPS || !dinv || !E0 || !u0_bf16 || !u1_bf16 || !out || !sync_block || n <= 0 || n_items < 0 ||
      n_layers < 1 || n_layers + 1 > MAX_BARRIERS || n_hot < 0 || n_hot > HOT_ROWS || (n_hot > 0 && !hot_rows) ||
      (n_long > 0 && (!long_rows || !scratch))) {
    set_error("lightgcn_propagate_bf16: bad arguments");
    return GDMCF_EBADARG;
  }
  if (d != D) { set_error("lightgcn_propagate_bf16: latent dimension must be %d (got %d)", D, d); return GDMCF_EBADARG; }
  if (((uintptr_t)E0 | (uintptr_t)out | (uintptr_t)u0_bf16 | (uintptr_t)u1_bf16 | (uintptr_t)scratch | (uintptr_t)items) & 15) {
    set_error("lightgcn_propagate_bf16: E0 / out / tables / scratch / items must be 16 B aligned");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(lightgcn_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(lightgcn_bf16)");
    attr_set = true;
  }
  Params p{};
  p.col = col_hot_first; p.items = reinterpret_cast<const int4*>(items); p.mids = item_mids; p.n_items = n_items;
  p.n_pieces = std::max(0, std::min(n_pieces, n_items));
  p.warp_ptr = warp_ptr;
  p.long_rows = long_rows; p.n_long = n_long; p.hot_rows = hot_rows; p.n_hot = n_hot;
  p.dinv = dinv; p.E0 = E0;
  p.u[0] = reinterpret_cast<__nv_bfloat16*>(u0_bf16); p.u[1] = reinterpret_cast<__nv_bfloat16*>(u1_bf16);
  p.out = out; p.scratch = scratch; p.sync = sync_block; p.n = n; p.n_layers = n_layers;
  const int grid = n_warp_slots / WARPS;  // the plan dealt the rows over exactly this many warps
  if (grid > gdmcf_num_sms()) {
    set_error("lightgcn_propagate_bf16: plan made for %d CTAs, device has %d SMs (cooperative launch)", grid, gdmcf_num_sms());
    return GDMCF_EBADARG;
  }
  launch_kernel_cooperative(lightgcn_bf16_kernel, grid, THREADS, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), p);
  return cuda_check_launch("lightgcn_bf16_kernel");
}

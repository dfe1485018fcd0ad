// K1 — normalized-adjacency propagation: Y = alpha * (A~ X) + beta * Z on a CSR adjacency.
//
// HBM-bound gather kernel (no tensor cores: there is no dense operand to feed them).
//  * Work is pre-split on the host (gdmcf_spmm_plan) into items of <= chunk non-zeros, so hub rows
//    (Zipf-head items touch tens of thousands of users) are spread over many warps; their partial sums
//    land in a scratch slab and a second tiny kernel reduces them in fixed order (deterministic, no atomics).
//  * One warp per item: the 32 lanes load 32 (col,val) pairs with one coalesced, L1-bypassing request,
//    broadcast them by shuffle, and gather the neighbour rows as float2 per lane (a 64-float row = two
//    fully-used 128 B lines per request), 8 independent gathers in flight per warp. Gathers allocate in
//    L1 (hot Zipf-head rows hit there); streams (col/val/Z/Y) do not.
//  * The LightGCN layer mean is folded in as a Horner recurrence T <- A~ T + E0, so no per-layer
//    embedding stack is ever materialised (lightGCN.py:188-189 does stack + mean).
//
// Reference: lightGCN.py:145-194 (get_A_tilda, propagate_through_layers); GCNConv propagate at
// models/DNN.py:1095,1100.
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace spmm {

GD_DEV int ld_stream_i32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
GD_DEV float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
GD_DEV float2 ld_stream_f32x2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

constexpr int WARPS_PER_CTA = 8;
constexpr int CTAS_PER_SM = 5;  // 40 warps per SM: the kernel hides L2 latency with thread-level parallelism

GD_DEV float4 ld_stream_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

GD_DEV void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// The 32 lanes take the (col,val) pairs [base, base+n) of an item. Lanes past n hold a duplicate of the last column
// with value 0, so the gather loop below needs no predicates (the duplicate row is an L1 hit).
template <bool kWeighted>
GD_DEV void load_pairs(const int* __restrict__ col, const float* __restrict__ val, int base, int n, int lane, int pad_col, int& c,
                       float& v) {
  c = 0;
  v = 0.f;
  if (kWeighted) {
    if (n > 0) c = ld_stream_i32(col + base + min(lane, n - 1));
    if (lane < n) v = ld_stream_f32(val + base + lane);
  } else {
    // binary adjacency: no values; lanes past n point at the all-zero row `pad_col` of X, so the sums need no predicate
    c = pad_col;
    if (lane < n) c = ld_stream_i32(col + base + lane);
  }
}

// acc += sum_j v_j * X[c_j, coff..coff+3]: the two half-warps take alternate non-zeros; each lane gathers a float4
// (16 lanes x 16 B = one 256 B embedding row), 4 independent gathers per lane per trip; 8 instructions per 2 nnz.
template <bool kWeighted>
GD_DEV void gather_pairs(const float* __restrict__ Xs, long long ldx, int cc, float vv, int n, int half, float4& acc) {
  const int ng = (n + 7) & ~7;
#pragma unroll 1
  for (int j0 = 0; j0 < ng; j0 += 8) {
    float4 x[4];
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 2 * u + half;  // <= 31
      const int c = __shfl_sync(0xffffffffu, cc, j);
      if (kWeighted) {
        v[u] = __shfl_sync(0xffffffffu, vv, j);
        x[u] = __ldg(reinterpret_cast<const float4*>(Xs + (long long)c * ldx));
      } else {
        x[u] = __ldg(reinterpret_cast<const float4*>(Xs + (long long)c * ldx));  // padding lanes read the zero row
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (kWeighted) {
        acc.x = fmaf(v[u], x[u].x, acc.x);
        acc.y = fmaf(v[u], x[u].y, acc.y);
        acc.z = fmaf(v[u], x[u].z, acc.z);
        acc.w = fmaf(v[u], x[u].w, acc.w);
      } else {
        acc.x += x[u].x;
        acc.y += x[u].y;
        acc.z += x[u].z;
        acc.w += x[u].w;
      }
    }
  }
}

// One warp per work item, items dealt round-robin to the resident warps. The item descriptor two items ahead (L1 prefetch) and the
// first 32 (col,val) pairs one item ahead are requested before the current item's gathers, so the dependent chain
// descriptor -> pairs -> rows costs one L2 round trip per item instead of three.
// kWeighted = false: binary adjacency (val unused); the per-row factor row_scale[r]^row_pow multiplies alpha in the epilogue
// (the separable normalisation D^-1/2 A D^-1/2 of LightGCN, see gdmcf_lightgcn_propagate_sym_f32).
template <bool kWeighted>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
spmm_items_kernel(const int* __restrict__ col, const float* __restrict__ val, const int4* __restrict__ items,
                  int n_items, const float* __restrict__ X, const float* __restrict__ Z, float* __restrict__ Y,
                  float* __restrict__ scratch, int d, float alpha, float beta, const float* __restrict__ row_scale,
                  int row_pow, int pad_col) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, hl = lane & 15;
  const int slabs = d >> 6;
  const int nwarps = gridDim.x * WARPS_PER_CTA;
  int item = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  if (item >= n_items) return;
  int4 it = __ldg(&items[item]);  // {row, begin, end, slot}
  if (item + nwarps < n_items) prefetch_l1(&items[item + nwarps]);
  int my_c;
  float my_v;
  load_pairs<kWeighted>(col, val, it.y, min(32, it.z - it.y), lane, pad_col, my_c, my_v);
  while (true) {
    const int item_n = item + nwarps;
    int c_n = 0;
    float v_n = 0.f;
    if (item_n < n_items) {
      const int4 nx = __ldg(&items[item_n]);  // L1 hit: prefetched one trip ago
      load_pairs<kWeighted>(col, val, nx.y, min(32, nx.z - nx.y), lane, pad_col, c_n, v_n);
      if (item_n + nwarps < n_items) prefetch_l1(&items[item_n + nwarps]);
    }

#pragma unroll 1
    for (int slab = 0; slab < slabs; ++slab) {
      const int coff = slab * 64 + hl * 4;
      const float* Xs = X + coff;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int base = it.y;
      int cc = my_c;
      float vv = my_v;
      while (true) {
        gather_pairs<kWeighted>(Xs, d, cc, vv, min(32, it.z - base), half, acc);
        base += 32;
        if (base >= it.z) break;
        load_pairs<kWeighted>(col, val, base, min(32, it.z - base), lane, pad_col, cc, vv);
      }
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
      if (half == 0) {
        if (it.w < 0) {
          float al = alpha;
          if (row_scale) {
            const float rsv = row_scale[it.x];
            al *= (row_pow == 2) ? rsv * rsv : rsv;
          }
          float4 o = make_float4(al * acc.x, al * acc.y, al * acc.z, al * acc.w);
          if (Z) {
            const float4 z = ld_stream_f32x4(Z + (long long)it.x * d + coff);
            o.x = fmaf(beta, z.x, o.x);
            o.y = fmaf(beta, z.y, o.y);
            o.z = fmaf(beta, z.z, o.z);
            o.w = fmaf(beta, z.w, o.w);
          }
          *reinterpret_cast<float4*>(Y + (long long)it.x * d + coff) = o;
        } else {
          *reinterpret_cast<float4*>(scratch + (long long)it.w * d + coff) = acc;
        }
      }
    }
    if (item_n >= n_items) break;
    item = item_n;
    it = __ldg(&items[item]);
    my_c = c_n;
    my_v = v_n;
  }
}

// long_rows: {row, first_slot, n_slots}; one CTA per (row, 64-column slab). The 8 warps sum contiguous ranges of the
// row's partial sums (slot order inside a range, 8 independent loads in flight), then warp 0 adds the 8 range sums in
// range order: a fixed summation tree (deterministic), ~n_slots/64 L2 round trips for a hub row instead of n_slots.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
spmm_long_reduce_kernel(const int* __restrict__ long_rows, int n_long, const float* __restrict__ scratch,
                        const float* __restrict__ Z, float* __restrict__ Y, int d, float alpha, float beta,
                        const float* __restrict__ row_scale, int row_pow) {
  pdl_entry();
  __shared__ float2 part_sm[WARPS_PER_CTA][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slabs = d >> 6;
  const long long total = (long long)n_long * slabs;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int lr = (int)(w / slabs);
    const int coff = (int)(w % slabs) * 64 + lane * 2;
    const int row = long_rows[3 * lr], first = long_rows[3 * lr + 1], cnt = long_rows[3 * lr + 2];
    const int per = (cnt + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int s_begin = min(cnt, warp * per), s_end = min(cnt, s_begin + per);
    float2 acc = make_float2(0.f, 0.f);
    const float* sp = scratch + (long long)first * d + coff;
    for (int s = s_begin; s < s_end; s += 8) {
      float2 part[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        part[q] = (s + q < s_end) ? *reinterpret_cast<const float2*>(sp + (long long)(s + q) * d) : make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        acc.x += part[q].x;
        acc.y += part[q].y;
      }
    }
    part_sm[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      float2 tot = part_sm[0][lane];
#pragma unroll
      for (int q = 1; q < WARPS_PER_CTA; ++q) {
        tot.x += part_sm[q][lane].x;
        tot.y += part_sm[q][lane].y;
      }
      float al = alpha;
      if (row_scale) {
        const float rsv = row_scale[row];
        al *= (row_pow == 2) ? rsv * rsv : rsv;
      }
      float2 o = make_float2(al * tot.x, al * tot.y);
      if (Z) {
        const float2 z = *reinterpret_cast<const float2*>(Z + (long long)row * d + coff);
        o.x = fmaf(beta, z.x, o.x);
        o.y = fmaf(beta, z.y, o.y);
      }
      *reinterpret_cast<float2*>(Y + (long long)row * d + coff) = o;
    }
    __syncthreads();
  }
}

// A~ = D^-1/2 [[0,R],[R^T,0]] D^-1/2, d_inv = (rowsum + 1e-9)^-1/2 (lightGCN.py:145-178). One thread per row.
__global__ void norm_adj_rowptr_kernel(const int* __restrict__ r_rowptr, const int* __restrict__ rt_rowptr,
                                       int n_users, int n_items, int* __restrict__ rowptr_out) {
  pdl_entry();
  const int n = n_users + n_items;
  const int nnz = r_rowptr[n_users];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += gridDim.x * blockDim.x)
    rowptr_out[r] = (r <= n_users) ? r_rowptr[r] : nnz + rt_rowptr[r - n_users];
}
GD_DEV float d_inv_of(int deg) { return 1.0f / sqrtf((float)deg + 1e-9f); }
__global__ void norm_adj_fill_kernel(const int* __restrict__ r_rowptr, const int* __restrict__ r_col,
                                     const int* __restrict__ rt_rowptr, const int* __restrict__ rt_col, int n_users,
                                     int n_items, int* __restrict__ col_out, float* __restrict__ val_out) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int n = n_users + n_items;
  const int nnz = r_rowptr[n_users];
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp0; r < n; r += nwarps) {
    if (r < n_users) {
      const int b = r_rowptr[r], e = r_rowptr[r + 1];
      const float du = d_inv_of(e - b);
      for (int j = b + lane; j < e; j += 32) {
        const int i = r_col[j];
        const float di = d_inv_of(rt_rowptr[i + 1] - rt_rowptr[i]);
        col_out[j] = n_users + i;
        val_out[j] = du * di;
      }
    } else {
      const int i = r - n_users;
      const int b = rt_rowptr[i], e = rt_rowptr[i + 1];
      const float di = d_inv_of(e - b);
      for (int j = b + lane; j < e; j += 32) {
        const int u = rt_col[j];
        const float du = d_inv_of(r_rowptr[u + 1] - r_rowptr[u]);
        col_out[nnz + j] = u;
        val_out[nnz + j] = di * du;
      }
    }
  }
}

static int grid_for_warps(long long warps) {
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const long long ctas = (warps + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  return (int)std::max<long long>(1, std::min<long long>(ctas, (long long)sms * CTAS_PER_SM));
}

}  // namespace spmm
}  // namespace gd

using namespace gd;
using namespace gd::spmm;

extern "C" int gdmcf_spmm_plan(const int32_t* rowptr, int n_rows, int chunk, int32_t* items_out, int cap_items,
                               int32_t* long_out, int cap_long, int* n_items, int* n_long, int* n_slots) {
  if (!rowptr || n_rows < 0 || chunk < 32) { set_error("spmm_plan: bad arguments (chunk must be >= 32)"); return GDMCF_EBADARG; }
  // Long-row chunks first (they are the heaviest items), then whole rows.
  long long ni = 0, nl = 0, ns = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int r = 0; r < n_rows; ++r) {
      const int b = rowptr[r], e = rowptr[r + 1];
      const int deg = e - b;
      if (deg < 0) { set_error("spmm_plan: rowptr not monotone at row %d", r); return GDMCF_EBADARG; }
      const bool is_long = deg > chunk;
      if (pass == 0 && is_long) {
        const int pieces = (deg + chunk - 1) / chunk;
        if (long_out) {
          if (nl >= cap_long) { set_error("spmm_plan: long_out capacity %d too small", cap_long); return GDMCF_EBADARG; }
          long_out[3 * nl] = r; long_out[3 * nl + 1] = (int)ns; long_out[3 * nl + 2] = pieces;
        }
        for (int p = 0; p < pieces; ++p) {
          if (items_out) {
            if (ni >= cap_items) { set_error("spmm_plan: items_out capacity %d too small", cap_items); return GDMCF_EBADARG; }
            items_out[4 * ni] = r;
            items_out[4 * ni + 1] = b + p * chunk;
            items_out[4 * ni + 2] = std::min(e, b + (p + 1) * chunk);
            items_out[4 * ni + 3] = (int)(ns + p);
          }
          ++ni;
        }
        ns += pieces;
        ++nl;
      } else if (pass == 1 && !is_long) {
        if (items_out) {
          if (ni >= cap_items) { set_error("spmm_plan: items_out capacity %d too small", cap_items); return GDMCF_EBADARG; }
          items_out[4 * ni] = r; items_out[4 * ni + 1] = b; items_out[4 * ni + 2] = e; items_out[4 * ni + 3] = -1;
        }
        ++ni;
      }
    }
  }
  if (n_items) *n_items = (int)ni;
  if (n_long) *n_long = (int)nl;
  if (n_slots) *n_slots = (int)ns;
  return GDMCF_OK;
}

static int spmm_launch(const int32_t* col, const float* val, const int32_t* items, int n_items, const int32_t* long_rows,
                       int n_long, const float* X, const float* Z, float* Y, float* scratch, int d, float alpha, float beta,
                       const float* row_scale, int row_pow, int pad_col, gdmcf_stream_t stream) {
  if (!col || (!val && !row_scale) || !items || !X || !Y || n_items < 0 || d <= 0 || (d & 63)) {
    set_error("spmm: need col/val/items/X/Y and d %% 64 == 0 (d=%d)", d);
    return GDMCF_EBADARG;
  }
  if (n_long > 0 && (!long_rows || !scratch)) { set_error("spmm: long rows need long_rows and scratch"); return GDMCF_EBADARG; }
  if (((uintptr_t)X | (uintptr_t)Y | (uintptr_t)Z | (uintptr_t)scratch | (uintptr_t)items) & 15) {
    set_error("spmm: X/Y/Z/scratch/items must be 16B aligned");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int slabs = d >> 6;
  if (n_items > 0) {
    const int grid = grid_for_warps((long long)n_items * slabs);
    if (val)
      launch_kernel(spmm_items_kernel<true>, grid, WARPS_PER_CTA * 32, 0, st, col, val, reinterpret_cast<const int4*>(items), n_items, X, Z, Y,
                                                                 scratch, d, alpha, beta, row_scale, row_pow, pad_col);
    else
      launch_kernel(spmm_items_kernel<false>, grid, WARPS_PER_CTA * 32, 0, st, col, val, reinterpret_cast<const int4*>(items), n_items, X, Z, Y,
                                                                  scratch, d, alpha, beta, row_scale, row_pow, pad_col);
    if ((rc = cuda_check_launch("spmm_items_kernel"))) return rc;
  }
  if (n_long > 0) {
    const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
    launch_kernel(spmm_long_reduce_kernel, (int)std::min<long long>((long long)n_long * slabs, (long long)sms * 8), WARPS_PER_CTA * 32, 0, st, 
        long_rows, n_long, scratch, Z, Y, d, alpha, beta, row_scale, row_pow);
    if ((rc = cuda_check_launch("spmm_long_reduce_kernel"))) return rc;
  }
  return GDMCF_OK;
}

extern "C" int gdmcf_spmm_csr_f32(const int32_t* col, const float* val, const int32_t* items, int n_items,
                                  const int32_t* long_rows, int n_long, const float* X, const float* Z, float* Y,
                                  float* scratch, int n_rows, int d, float alpha, float beta, gdmcf_stream_t stream) {
  (void)n_rows;
  if (!val) { set_error("spmm: val is required (use gdmcf_lightgcn_propagate_sym_f32 for a binary adjacency)"); return GDMCF_EBADARG; }
  return spmm_launch(col, val, items, n_items, long_rows, n_long, X, Z, Y, scratch, d, alpha, beta, nullptr, 0, 0, stream);
}

// u0[r, :] = dinv[r] * E0[r, :]
__global__ void scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y, long long n4, int d4) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float sv = s[i / d4];
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x *= sv; v.y *= sv; v.z *= sv; v.w *= sv;
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

// dinv[r] = (deg_r + 1e-9)^-1/2 for the bipartite graph [[0, R], [R^T, 0]] (lightGCN.py:160-166)
__global__ void norm_adj_dinv_kernel(const int* __restrict__ r_rowptr, const int* __restrict__ rt_rowptr, int n_users, int n_items,
                                     float* __restrict__ dinv) {
  pdl_entry();
  const int n = n_users + n_items;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int deg = r < n_users ? r_rowptr[r + 1] - r_rowptr[r] : rt_rowptr[r - n_users + 1] - rt_rowptr[r - n_users];
    dinv[r] = d_inv_of(deg);
  }
}

extern "C" int gdmcf_norm_adj_dinv(const int32_t* r_rowptr, const int32_t* rt_rowptr, int n_users, int n_items, float* dinv,
                                   gdmcf_stream_t stream) {
  if (!r_rowptr || !rt_rowptr || !dinv || n_users <= 0 || n_items <= 0) { set_error("norm_adj_dinv: bad arguments"); return GDMCF_EBADARG; }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  launch_kernel(norm_adj_dinv_kernel, (n_users + n_items + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream), r_rowptr, rt_rowptr, n_users,
                                                                                                      n_items, dinv);
  return cuda_check_launch("norm_adj_dinv_kernel");
}

// mean_k (A~^k E0) for A~ = D^-1/2 A D^-1/2 with a BINARY A (pattern `col`) and dinv = diag(D^-1/2): iterate on
// U_k = D^-1/2 T_k, for which the Horner step T <- A~ T + E0 becomes U <- D^-1 (A U) + U_0 with plain (unweighted) neighbour
// sums — no value stream, no per-non-zero multiply; the last layer maps back: out = (dinv * (A U_{K-1}) + E0) / (K + 1).
// u0 / tmp0 / tmp1 must have n + 1 rows with row n all zero (never written here): padding lanes of the gather read it.
extern "C" int gdmcf_lightgcn_propagate_sym_f32(const int32_t* col, const float* dinv, const int32_t* items, int n_items,
                                                const int32_t* long_rows, int n_long, const float* E0, float* u0, float* tmp0,
                                                float* tmp1, float* out, float* scratch, int n, int d, int n_layers,
                                                gdmcf_stream_t stream) {
  if (n_layers < 1 || !E0 || !out || !dinv || !u0 || (n_layers > 1 && !tmp0) || (n_layers > 2 && !tmp1) || n <= 0 || d <= 0 || (d & 63)) {
    set_error("lightgcn_propagate_sym: need n_layers >= 1, E0, dinv, u0, out and ping-pong buffers");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  const long long n4 = (long long)n * d / 4;
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  launch_kernel(scale_rows_kernel, (int)std::min<long long>((n4 + 255) / 256, (long long)sms * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      E0, dinv, u0, n4, d / 4);
  if ((rc = cuda_check_launch("scale_rows_kernel"))) return rc;
  const float* src = u0;
  for (int l = 0; l < n_layers; ++l) {
    const bool last = (l == n_layers - 1);
    float* dst = last ? out : ((l & 1) ? tmp1 : tmp0);
    const float s = last ? 1.0f / (float)(n_layers + 1) : 1.0f;
    rc = spmm_launch(col, nullptr, items, n_items, long_rows, n_long, src, last ? E0 : u0, dst, scratch, d, s, s, dinv, last ? 1 : 2,
                     n, stream);
    if (rc) return rc;
    src = dst;
  }
  return GDMCF_OK;
}

extern "C" int gdmcf_lightgcn_propagate_f32(const int32_t* col, const float* val, const int32_t* items, int n_items,
                                            const int32_t* long_rows, int n_long, const float* E0, float* tmp0,
                                            float* tmp1, float* out, float* scratch, int n, int d, int n_layers,
                                            gdmcf_stream_t stream) {
  if (n_layers < 1 || !E0 || !out || (n_layers > 1 && !tmp0) || (n_layers > 2 && !tmp1)) {
    set_error("lightgcn_propagate: need n_layers >= 1, E0, out and ping-pong buffers");
    return GDMCF_EBADARG;
  }
  // T_1 = A E0 + E0; T_{k+1} = A T_k + E0; out = T_K / (K + 1)  ==  mean_k A^k E0.
  const float* src = E0;
  for (int l = 0; l < n_layers; ++l) {
    const bool last = (l == n_layers - 1);
    float* dst = last ? out : ((l & 1) ? tmp1 : tmp0);
    const float s = last ? 1.0f / (float)(n_layers + 1) : 1.0f;
    int rc = gdmcf_spmm_csr_f32(col, val, items, n_items, long_rows, n_long, src, E0, dst, scratch, n, d, s, s, stream);
    if (rc) return rc;
    src = dst;
  }
  return GDMCF_OK;
}

extern "C" int gdmcf_build_norm_adj(const int32_t* r_rowptr, const int32_t* r_col, const int32_t* rt_rowptr,
                                    const int32_t* rt_col, int n_users, int n_items, int32_t* rowptr_out,
                                    int32_t* col_out, float* val_out, gdmcf_stream_t stream) {
  if (!r_rowptr || !r_col || !rt_rowptr || !rt_col || !rowptr_out || !col_out || !val_out || n_users <= 0 || n_items <= 0) {
    set_error("build_norm_adj: null pointer or empty shape");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = n_users + n_items;
  launch_kernel(norm_adj_rowptr_kernel, std::min(1024, ceil_div(n + 1, 256)), 256, 0, st, r_rowptr, rt_rowptr, n_users, n_items, rowptr_out);
  if ((rc = cuda_check_launch("norm_adj_rowptr_kernel"))) return rc;
  launch_kernel(norm_adj_fill_kernel, std::min(148 * 16, ceil_div(n, 8)), 256, 0, st, r_rowptr, r_col, rt_rowptr, rt_col, n_users, n_items, col_out, val_out);
  return cuda_check_launch("norm_adj_fill_kernel");
}

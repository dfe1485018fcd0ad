// AdamW fused with the refresh of everything the contractions derive from a weight matrix.
//
// After every optimizer step the engine needs, for each big parameter W [rows, cols] (fp32 master copy):
//   * the bf16 GEMM operand W[:, :cols_used]                      (hi[, lo] parts, zero padded to ld_hi)
//   * its transpose (dgrad operands: E^T, GCN weights)             (hi[, lo], [cols_used, ld_t])
//   * row inverse norms of the item table (cosine scorer, models/DNN.py:1320)
//   * the one-hot encoder tables of in_layers2.0.weight: delta[i,:] = W[:, 2i+1] - W[:, 2i], base = sum_i W[:, 2i]
//   * a contiguous copy of the time-embedding columns W[:, cols_used:]
// Done as separate passes these re-read 2.2 GB of fp32 weights per step at the Yelp shape. Here the AdamW pass
// (torch.optim.AdamW semantics, main.py:258,351 — identical arithmetic to adamw_kernel) produces them from the
// updated values while they are still in registers / shared memory: one read of p, g, m, v, one write of p, m, v and of
// each derived tensor.
//
// Tiling: a CTA owns 32 rows x a column range and walks it in 32 x 64 tiles (256 threads, 8 elements each, 16 B
// accesses when cols % 4 == 0). Updated values pass through a 32 x 65 smem tile for the transposed outputs.
// Row reductions are deterministic: per-(column range) partial sums, finished in fixed order by adamw_row_finish_kernel.
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace opt {

constexpr int TR = 32, TC = 64, THREADS = 256;

struct Args {
  float* p; const float* g; float* m; float* v;
  int rows, cols; long long ld_g;
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale;
  const long long* step_dev;
  int cols_used;
  __nv_bfloat16* hi; __nv_bfloat16* lo; long long ld_hi;
  __nv_bfloat16* t_hi; __nv_bfloat16* t_lo; long long ld_t;
  float* rowpart;   // [col_splits, rows] partial row sums (sum p^2, or sum of even columns when delta != NULL)
  float* delta; long long ld_delta;
  float* tcols; int n_tcols;
  const float* row_coef;  // optional: effective gradient = g + row_coef[r] * p (deferred row-wise term, see header)
  int col_splits, tiles_per_split;
};

GD_DEV void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
GD_DEV uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// UPDATE = false: refresh only (derive the tensors from the current weights; g, m, v are not touched).
template <int VEC, bool UPDATE>
__global__ void __launch_bounds__(THREADS)
adamw_refresh_kernel(const Args a) {
  pdl_entry();
  __shared__ float tile[TR][TC + 1];
  const AdamwCoef kc = adamw_coef(a.lr, a.beta1, a.beta2, a.eps, a.weight_decay, a.bc1, a.bc2_sqrt, a.grad_scale, a.step_dev);
  auto update = [&](float& param, float gr, float& mi, float& vi) { adamw_update(kc, param, gr, mi, vi); };

  const int tid = threadIdx.x;
  const int row_groups = (a.rows + TR - 1) / TR;
  const int rg = blockIdx.x % row_groups, split = blockIdx.x / row_groups;
  const int r0 = rg * TR;
  const int tiles_c = (a.cols + TC - 1) / TC;
  const int tc_begin = split * a.tiles_per_split, tc_end = min(tiles_c, tc_begin + a.tiles_per_split);
  // element mapping of the update phase: rows rl and rl + 16, four consecutive columns
  const int rl = tid >> 4, cq = (tid & 15) * 4;
  float rsum[2] = {0.f, 0.f};
  double rsum_d[2] = {0.0, 0.0};  // base = sum of ~n_item zero-mean weights: accumulated in fp64 like onehot_base_kernel
  const bool want_rowsum = a.rowpart != nullptr;
  const bool even_cols = a.delta != nullptr;  // one-hot tables: base = sum of the even columns

  for (int tcx = tc_begin; tcx < tc_end; ++tcx) {
    const int c0 = tcx * TC;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rloc = rl + 16 * h, r = r0 + rloc, c = c0 + cq;
      float pv[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < a.rows && c < a.cols) {
        const long long off = (long long)r * a.cols + c;
        const long long goff = (long long)r * a.ld_g + c;
        float gv[4], mv[4], vv[4];
        if (!UPDATE) {
          if (VEC == 4) {
            const float4 P = *reinterpret_cast<const float4*>(a.p + off);
            pv[0] = P.x; pv[1] = P.y; pv[2] = P.z; pv[3] = P.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) pv[j] = (c + j < a.cols) ? a.p[off + j] : 0.f;
          }
        } else if (VEC == 4) {
          const float4 P = *reinterpret_cast<const float4*>(a.p + off), G = *reinterpret_cast<const float4*>(a.g + goff);
          const float4 M = *reinterpret_cast<const float4*>(a.m + off), V = *reinterpret_cast<const float4*>(a.v + off);
          pv[0] = P.x; pv[1] = P.y; pv[2] = P.z; pv[3] = P.w;
          gv[0] = G.x; gv[1] = G.y; gv[2] = G.z; gv[3] = G.w;
          mv[0] = M.x; mv[1] = M.y; mv[2] = M.z; mv[3] = M.w;
          vv[0] = V.x; vv[1] = V.y; vv[2] = V.z; vv[3] = V.w;
          if (a.row_coef) {
            const float rc = a.row_coef[r];
#pragma unroll
            for (int j = 0; j < 4; ++j) gv[j] = fmaf(rc, pv[j], gv[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) update(pv[j], gv[j], mv[j], vv[j]);
          *reinterpret_cast<float4*>(a.p + off) = make_float4(pv[0], pv[1], pv[2], pv[3]);
          *reinterpret_cast<float4*>(a.m + off) = make_float4(mv[0], mv[1], mv[2], mv[3]);
          *reinterpret_cast<float4*>(a.v + off) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = c + j < a.cols;
            pv[j] = ok ? a.p[off + j] : 0.f;
            gv[j] = ok ? a.g[goff + j] : 0.f;
            mv[j] = ok ? a.m[off + j] : 0.f;
            vv[j] = ok ? a.v[off + j] : 0.f;
          }
          const float rc = a.row_coef ? a.row_coef[r] : 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c + j < a.cols) {
              gv[j] = fmaf(rc, pv[j], gv[j]);
              update(pv[j], gv[j], mv[j], vv[j]);
              a.p[off + j] = pv[j]; a.m[off + j] = mv[j]; a.v[off + j] = vv[j];
            } else {
              pv[j] = 0.f;
            }
          }
        }
        if (a.tcols) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tcj = c + j - a.cols_used;
            if (tcj >= 0 && tcj < a.n_tcols) a.tcols[(long long)r * a.n_tcols + tcj] = pv[j];
          }
        }
      }
      // values outside the operand's column range are zero in every derived tensor
      float ov[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) ov[j] = (c + j < a.cols_used) ? pv[j] : 0.f;
      if (want_rowsum) {
        if (even_cols) rsum_d[h] += (double)ov[0] + (double)ov[2];  // cq % 4 == 0: columns c, c+2 are the even ones
        else rsum[h] += ov[0] * ov[0] + ov[1] * ov[1] + ov[2] * ov[2] + ov[3] * ov[3];
      }
      if (a.hi && r < a.rows && c < a.ld_hi) {
        __nv_bfloat16 hh[4], ll[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(ov[j], hh[j], ll[j]);
        *reinterpret_cast<uint2*>(a.hi + (long long)r * a.ld_hi + c) = make_uint2(pack2(hh[0], hh[1]), pack2(hh[2], hh[3]));
        if (a.lo) *reinterpret_cast<uint2*>(a.lo + (long long)r * a.ld_hi + c) = make_uint2(pack2(ll[0], ll[1]), pack2(ll[2], ll[3]));
      }
      if (a.t_hi || a.delta) {
#pragma unroll
        for (int j = 0; j < 4; ++j) tile[rloc][cq + j] = ov[j];
      }
    }
    if (a.t_hi || a.delta) {
      __syncthreads();
      if (a.t_hi) {
        // T[c, r0 .. r0+32): thread = (column c0 + tid/4, 8 rows) -> one 16 B store; 4 threads cover 64 B per column
        const int cl = tid >> 2, r8 = (tid & 3) * 8;
        const int c = c0 + cl;
        if (c < a.cols_used) {
          __nv_bfloat16 hh[8], ll[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_bf16(tile[r8 + j][cl], hh[j], ll[j]);  // rows >= a.rows hold zeros
          const long long o = (long long)c * a.ld_t + r0 + r8;
          if (r0 + r8 < a.ld_t) {
            *reinterpret_cast<uint4*>(a.t_hi + o) = make_uint4(pack2(hh[0], hh[1]), pack2(hh[2], hh[3]), pack2(hh[4], hh[5]), pack2(hh[6], hh[7]));
            if (a.t_lo) *reinterpret_cast<uint4*>(a.t_lo + o) = make_uint4(pack2(ll[0], ll[1]), pack2(ll[2], ll[3]), pack2(ll[4], ll[5]), pack2(ll[6], ll[7]));
          }
        }
      }
      if (a.delta) {
        // delta[i, r0 .. r0+32) for the 32 item pairs of this tile: thread = (pair tid/8, 4 rows) -> one 16 B store
        const int il = tid >> 3, r4 = (tid & 7) * 4;
        const int i = (c0 >> 1) + il;
        if (2 * i + 1 < a.cols_used && r0 + r4 < a.ld_delta) {
          float dv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) dv[j] = tile[r4 + j][2 * il + 1] - tile[r4 + j][2 * il];
          *reinterpret_cast<float4*>(a.delta + (long long)i * a.ld_delta + r0 + r4) = make_float4(dv[0], dv[1], dv[2], dv[3]);
        }
      }
      __syncthreads();
    }
  }
  if (want_rowsum) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double sd = even_cols ? rsum_d[h] : (double)rsum[h];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);  // the 16 lanes that share a row
      const float s = (float)sd;
      const int r = r0 + rl + 16 * h;
      if ((tid & 15) == 0 && r < a.rows) a.rowpart[(long long)split * a.rows + r] = s;
    }
  }
}

// Weights whose rows are not 16 B aligned (nn.Linear(n_item + emb_size, d): odd row length) and that need only the
// row-major operand copy and the trailing-column copy: the update streams the flat p / m / v buffers with 16 B accesses
// and recovers each element's (row, col) to read its gradient (padded leading dimension) and to place the bf16 copy
// (2-byte stores; neighbouring threads fill the sectors). Operand padding columns are never touched (zero since built).
__global__ void __launch_bounds__(THREADS)
adamw_flat_hi_kernel(const Args a) {
  pdl_entry();
  const AdamwCoef kc = adamw_coef(a.lr, a.beta1, a.beta2, a.eps, a.weight_decay, a.bc1, a.bc2_sqrt, a.grad_scale, a.step_dev);
  const long long total = (long long)a.rows * a.cols;
  const long long groups = (total + 3) >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < groups; i += (long long)gridDim.x * blockDim.x) {
    const long long idx = i << 2;
    int r = (int)(idx / a.cols);
    int c = (int)(idx - (long long)r * a.cols);
    float pv[4], mv[4], vv[4];
    const bool full = idx + 4 <= total;
    if (full) {
      const float4 P = *reinterpret_cast<const float4*>(a.p + idx), M = *reinterpret_cast<const float4*>(a.m + idx);
      const float4 V = *reinterpret_cast<const float4*>(a.v + idx);
      pv[0] = P.x; pv[1] = P.y; pv[2] = P.z; pv[3] = P.w;
      mv[0] = M.x; mv[1] = M.y; mv[2] = M.z; mv[3] = M.w;
      vv[0] = V.x; vv[1] = V.y; vv[2] = V.z; vv[3] = V.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = idx + j < total;
        pv[j] = ok ? a.p[idx + j] : 0.f; mv[j] = ok ? a.m[idx + j] : 0.f; vv[j] = ok ? a.v[idx + j] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (idx + j < total) {
        float gr = a.g[(long long)r * a.ld_g + c];
        if (a.row_coef) gr = fmaf(a.row_coef[r], pv[j], gr);
        adamw_update(kc, pv[j], gr, mv[j], vv[j]);
        if (c < a.cols_used) {
          if (a.hi) {
            __nv_bfloat16 hh, ll;
            split_bf16(pv[j], hh, ll);
            a.hi[(long long)r * a.ld_hi + c] = hh;
            if (a.lo) a.lo[(long long)r * a.ld_hi + c] = ll;
          }
        } else if (a.tcols) {
          a.tcols[(long long)r * a.n_tcols + (c - a.cols_used)] = pv[j];
        }
      }
      if (++c == a.cols) { c = 0; ++r; }
    }
    if (full) {
      *reinterpret_cast<float4*>(a.p + idx) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      *reinterpret_cast<float4*>(a.m + idx) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(a.v + idx) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (idx + j < total) { a.p[idx + j] = pv[j]; a.m[idx + j] = mv[j]; a.v[idx + j] = vv[j]; }
    }
  }
}

// AdamW only (no derived tensors) by a FIXED set of SMs. Launched as `ctas` CTAs (clusters of 2 = whole TPCs) of 1024
// threads that each reserve more than half of an SM's shared memory: exactly one CTA per SM, and no CTA of a tcgen05
// contraction (227 KB) can share that SM. The optimizer pass is HBM-bound and the denoise + rank phase is tensor-bound;
// with the contractions limited to the other SMs (gdmcf_gemm_set_sm_limit) the two run side by side on two streams
// instead of back to back. Same element arithmetic as every other AdamW kernel here (adamw_update).
constexpr int PART_THREADS = 1024;
constexpr int PART_SMEM = 120 * 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PART_THREADS, 1)
adamw_partition_kernel(const Args a) {
  pdl_entry();
  const AdamwCoef kc = adamw_coef(a.lr, a.beta1, a.beta2, a.eps, a.weight_decay, a.bc1, a.bc2_sqrt, a.grad_scale, a.step_dev);
  const long long total = (long long)a.rows * a.cols;
  const long long groups = (total + 3) >> 2;
  const bool g_flat = a.ld_g == a.cols && (((uintptr_t)a.g & 15) == 0);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < groups; i0 += 2 * stride) {
    // two independent groups per trip: 8 x 16 B loads in flight per thread (128 KB per SM)
    float pv[2][4], mv[2][4], vv[2][4], gv[2][4];
    int rr[2][4];
    bool live[2], full[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * stride;
      live[u] = i < groups;
      const long long idx = i << 2;
      full[u] = live[u] && idx + 4 <= total;
      if (full[u]) {
        const float4 P = *reinterpret_cast<const float4*>(a.p + idx), M = *reinterpret_cast<const float4*>(a.m + idx);
        const float4 V = *reinterpret_cast<const float4*>(a.v + idx);
        pv[u][0] = P.x; pv[u][1] = P.y; pv[u][2] = P.z; pv[u][3] = P.w;
        mv[u][0] = M.x; mv[u][1] = M.y; mv[u][2] = M.z; mv[u][3] = M.w;
        vv[u][0] = V.x; vv[u][1] = V.y; vv[u][2] = V.z; vv[u][3] = V.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = live[u] && idx + j < total;
          pv[u][j] = ok ? a.p[idx + j] : 0.f; mv[u][j] = ok ? a.m[idx + j] : 0.f; vv[u][j] = ok ? a.v[idx + j] : 0.f;
        }
      }
      if (live[u]) {
        int r = (int)(idx / a.cols);
        int c = (int)(idx - (long long)r * a.cols);
        if (g_flat && full[u]) {
          const float4 G = *reinterpret_cast<const float4*>(a.g + idx);
          gv[u][0] = G.x; gv[u][1] = G.y; gv[u][2] = G.z; gv[u][3] = G.w;
#pragma unroll
          for (int j = 0; j < 4; ++j) { rr[u][j] = r; if (++c == a.cols) { c = 0; ++r; } }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            gv[u][j] = (idx + j < total) ? a.g[(long long)r * a.ld_g + c] : 0.f;
            rr[u][j] = r;
            if (++c == a.cols) { c = 0; ++r; }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!live[u]) continue;
      const long long idx = (i0 + u * stride) << 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (idx + j < total) {
          float gr = gv[u][j];
          if (a.row_coef) gr = fmaf(a.row_coef[rr[u][j]], pv[u][j], gr);
          adamw_update(kc, pv[u][j], gr, mv[u][j], vv[u][j]);
        }
      }
      if (full[u]) {
        *reinterpret_cast<float4*>(a.p + idx) = make_float4(pv[u][0], pv[u][1], pv[u][2], pv[u][3]);
        *reinterpret_cast<float4*>(a.m + idx) = make_float4(mv[u][0], mv[u][1], mv[u][2], mv[u][3]);
        *reinterpret_cast<float4*>(a.v + idx) = make_float4(vv[u][0], vv[u][1], vv[u][2], vv[u][3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (idx + j < total) { a.p[idx + j] = pv[u][j]; a.m[idx + j] = mv[u][j]; a.v[idx + j] = vv[u][j]; }
      }
    }
  }
}

// AdamW on an embedding table whose gradient has only a few non-zero ROWS per step (embedding_user: the B users of the
// batch, models/DNN.py:1265), without streaming the whole table every step. torch.optim.AdamW still moves a row that got
// no gradient (its moments decay and keep pushing the weights), so skipping rows would change the result. Instead every
// row remembers the last step it was brought up to date (`last_step`); when a row receives a gradient at step `cur`, the
// steps it missed are replayed first, one by one, with a zero gradient — exactly the arithmetic (adamw_update, same bias
// corrections per step) the dense pass would have executed — and then the real update is applied. Without a gradient
// the selected rows (idx; the rows the next forward pass reads) or the whole table (idx == NULL: flush) are replayed up
// to and including `cur`; a flush must run before anything outside the training step reads the table.
// Values are bit-identical to the dense pass; HBM traffic drops from 28 B per table element per step to the touched rows.
constexpr int LAZY_THREADS = 256;
constexpr int LAZY_CHUNK = 128;  // replayed steps whose coefficients are staged in shared memory at a time

__global__ void __launch_bounds__(LAZY_THREADS)
adamw_rows_lazy_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int* __restrict__ last_step,
                       const int* __restrict__ idx, const float* __restrict__ grows, long long ld_g, int n_sel, int cols,
                       float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                       const long long* __restrict__ step_dev, int step_host) {
  pdl_entry();
  __shared__ float s_step_size[LAZY_CHUNK], s_bc2[LAZY_CHUNK];
  const int cur = step_dev ? (int)step_dev[0] : step_host;
  AdamwCoef kc = adamw_coef(lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, grad_scale, nullptr);
  for (int sel = blockIdx.x; sel < n_sel; sel += gridDim.x) {
    const int r = idx ? idx[sel] : sel;
    const int last = last_step[r];
    // with a gradient: zero-gradient steps last+1 .. cur-1, then the real update at cur; without (catch-up of the rows
    // the next forward pass will read, or a flush of the whole table): zero-gradient steps last+1 .. cur
    const bool has_grad = grows != nullptr;
    const int replay_end = has_grad ? cur - 1 : cur;
    if (last >= cur) continue;                   // already up to date (a flush right after a flush)
    float* pr = p + (long long)r * cols;
    float* mr = m + (long long)r * cols;
    float* vr = v + (long long)r * cols;
    const float* gr = has_grad ? grows + (long long)sel * ld_g : nullptr;
    for (int c0 = 0; c0 < cols; c0 += 4 * LAZY_THREADS) {
      float pv[4], mv[4], vv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j * LAZY_THREADS + threadIdx.x;
        const bool ok = c < cols;
        pv[j] = ok ? pr[c] : 0.f; mv[j] = ok ? mr[c] : 0.f; vv[j] = ok ? vr[c] : 0.f;
      }
      // a row that never received a gradient has zero moments: its zero-gradient steps change nothing (without weight
      // decay), so the replay is skipped for the whole CTA when no thread holds a non-zero moment
      bool live = weight_decay != 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) live = live || mv[j] != 0.f || vv[j] != 0.f;
      const bool replay = __syncthreads_or(live ? 1 : 0) != 0;
      for (int s0 = last + 1; replay && s0 <= replay_end; s0 += LAZY_CHUNK) {
        const int n = min(LAZY_CHUNK, replay_end - s0 + 1);
        __syncthreads();
        if (threadIdx.x < n) {  // bias corrections of step s0 + t, same double arithmetic as adamw_coef
          const double st = (double)(s0 + threadIdx.x);
          s_step_size[threadIdx.x] = lr / (float)(1.0 - pow((double)beta1, st));
          s_bc2[threadIdx.x] = (float)sqrt(1.0 - pow((double)beta2, st));
        }
        __syncthreads();
        for (int t = 0; t < n; ++t) {
          kc.step_size = s_step_size[t];
          kc.bc2_sqrt = s_bc2[t];
#pragma unroll
          for (int j = 0; j < 4; ++j) adamw_update(kc, pv[j], 0.f, mv[j], vv[j]);
        }
      }
      if (has_grad) {
        const double st = (double)cur;
        kc.step_size = lr / (float)(1.0 - pow((double)beta1, st));
        kc.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, st));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + j * LAZY_THREADS + threadIdx.x;
          adamw_update(kc, pv[j], c < cols ? gr[c] : 0.f, mv[j], vv[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j * LAZY_THREADS + threadIdx.x;
        if (c < cols) { pr[c] = pv[j]; mr[c] = mv[j]; vr[c] = vv[j]; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) last_step[r] = cur;
  }
}

// out[r] = finish(sum_s rowpart[s, r]) in split order; mode 0: 1/sqrt (row inverse norm), mode 1: plain sum (base).
__global__ void adamw_row_finish_kernel(const float* __restrict__ rowpart, int splits, int rows, int mode, float* __restrict__ out) {
  pdl_entry();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < splits; ++q) s += rowpart[(long long)q * rows + r];
    out[r] = mode == 0 ? 1.0f / sqrtf(s) : s;
  }
}

}  // namespace opt
}  // namespace gd

using namespace gd;
using namespace gd::opt;

extern "C" int gdmcf_adamw_refresh_splits(int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 1;
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const int row_groups = (rows + TR - 1) / TR, tiles_c = (cols + TC - 1) / TC;
  int splits = std::max(1, std::min(tiles_c, (sms * 8 + row_groups - 1) / row_groups));
  const int per = (tiles_c + splits - 1) / splits;
  return (tiles_c + per - 1) / per;
}

extern "C" int gdmcf_adamw_refresh(float* p, const float* g, int64_t ld_g, float* m, float* v, int rows, int cols, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int step,
                                   const int64_t* step_dev, float grad_scale, const gdmcf_refresh* out,
                                   gdmcf_stream_t stream) {
  // g == NULL: refresh only (m, v unused): the derived tensors are recomputed from the current weights by the same code
  // that produces them during training, so both paths agree bit for bit (row sums have a fixed order per geometry)
  if (!p || !out || rows <= 0 || cols <= 0 || (g && (!m || !v || ld_g < cols || (step < 1 && !step_dev)))) {
    set_error("adamw_refresh: bad arguments");
    return GDMCF_EBADARG;
  }
  const gdmcf_refresh& o = *out;
  const int cols_used = o.cols_used > 0 ? o.cols_used : cols;
  const bool vec = (cols % 4 == 0) && (!g || ld_g % 4 == 0) && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  if (cols_used > cols || (o.hi && ((o.ld_hi & 7) || o.ld_hi < cols_used || ((uintptr_t)o.hi & 15) || ((uintptr_t)o.lo & 15))) ||
      (o.t_hi && ((o.ld_t & 7) || o.ld_t < rows || ((uintptr_t)o.t_hi & 15) || ((uintptr_t)o.t_lo & 15))) ||
      (o.lo && !o.hi) || (o.t_lo && !o.t_hi) ||
      (o.delta && ((cols_used & 1) || (o.ld_delta & 3) || o.ld_delta < rows || !o.base || ((uintptr_t)o.delta & 15))) ||
      (o.inv_norm && o.delta) || ((o.inv_norm || o.delta) && !o.rowpart) || (o.tcols && o.n_tcols != cols - cols_used)) {
    set_error("adamw_refresh: inconsistent derived-tensor description (rows=%d cols=%d cols_used=%d)", rows, cols, cols_used);
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Args a{};
  a.p = p; a.g = g; a.m = m; a.v = v; a.rows = rows; a.cols = cols; a.ld_g = ld_g;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  // same host-side double arithmetic as gdmcf_adamw_fused (ignored when step_dev is given)
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)std::max(step, 1)));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)std::max(step, 1)));
  a.grad_scale = grad_scale; a.step_dev = reinterpret_cast<const long long*>(step_dev);
  a.cols_used = cols_used;
  a.hi = (__nv_bfloat16*)o.hi; a.lo = (__nv_bfloat16*)o.lo; a.ld_hi = o.ld_hi;
  a.t_hi = (__nv_bfloat16*)o.t_hi; a.t_lo = (__nv_bfloat16*)o.t_lo; a.ld_t = o.ld_t;
  a.rowpart = (o.inv_norm || o.delta) ? o.rowpart : nullptr;
  a.delta = o.delta; a.ld_delta = o.ld_delta;
  a.tcols = o.tcols; a.n_tcols = o.n_tcols;
  a.row_coef = o.row_coef;
  const int row_groups = (rows + TR - 1) / TR, tiles_c = (cols + TC - 1) / TC;
  a.col_splits = gdmcf_adamw_refresh_splits(rows, cols);
  a.tiles_per_split = (tiles_c + a.col_splits - 1) / a.col_splits;
  const long long ctas = (long long)row_groups * a.col_splits;
  if (ctas > 0x7fffffffLL) { set_error("adamw_refresh: too many tiles"); return GDMCF_EBADARG; }
  const bool flat_ok = g && !vec && !a.t_hi && !a.delta && !a.rowpart && ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  if (!g) {
    if (vec) launch_kernel(adamw_refresh_kernel<4, false>, (int)ctas, THREADS, 0, st, a);
    else launch_kernel(adamw_refresh_kernel<1, false>, (int)ctas, THREADS, 0, st, a);
  } else if (vec) {
    launch_kernel(adamw_refresh_kernel<4, true>, (int)ctas, THREADS, 0, st, a);
  } else if (flat_ok) {
    const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
    const long long groups = ((long long)rows * cols + 3) / 4;
    launch_kernel(adamw_flat_hi_kernel, (int)std::min<long long>((groups + THREADS - 1) / THREADS, (long long)sms * 8), THREADS, 0, st, a);
  } else {
    launch_kernel(adamw_refresh_kernel<1, true>, (int)ctas, THREADS, 0, st, a);
  }
  if ((rc = cuda_check_launch("adamw_refresh_kernel"))) return rc;
  if (a.rowpart) {
    float* dst = o.inv_norm ? o.inv_norm : o.base;
    launch_kernel(adamw_row_finish_kernel, (rows + 255) / 256, 256, 0, st, a.rowpart, a.col_splits, rows, o.inv_norm ? 0 : 1, dst);
    rc = cuda_check_launch("adamw_row_finish_kernel");
  }
  return rc;
}

extern "C" int gdmcf_adamw_partitioned(float* p, const float* g, int64_t ld_g, float* m, float* v, int rows, int cols, float lr,
                                       float beta1, float beta2, float eps, float weight_decay, int step,
                                       const int64_t* step_dev, float grad_scale, const float* row_coef, int n_ctas,
                                       gdmcf_stream_t stream) {
  if (!p || !g || !m || !v || rows <= 0 || cols <= 0 || ld_g < cols || (step < 1 && !step_dev) || n_ctas < 2 ||
      ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) != 0)) {
    set_error("adamw_partitioned: bad arguments (16 B aligned p/m/v, n_ctas >= 2)");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(adamw_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PART_SMEM);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(adamw_partition)");
    attr_set = true;
  }
  Args a{};
  a.p = p; a.g = g; a.m = m; a.v = v; a.rows = rows; a.cols = cols; a.ld_g = ld_g;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)std::max(step, 1)));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)std::max(step, 1)));
  a.grad_scale = grad_scale; a.step_dev = reinterpret_cast<const long long*>(step_dev);
  a.row_coef = row_coef;
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const int ctas = std::max(2, std::min(n_ctas, sms) & ~1);
  launch_kernel(adamw_partition_kernel, ctas, PART_THREADS, PART_SMEM, reinterpret_cast<cudaStream_t>(stream), a);
  return cuda_check_launch("adamw_partition_kernel");
}

extern "C" int gdmcf_adamw_rows_lazy(float* p, float* m, float* v, int32_t* last_step, const int32_t* idx, const float* grad_rows,
                                     int64_t ld_g, int n_sel, int n_rows, int cols, float lr, float beta1, float beta2, float eps,
                                     float weight_decay, int step, const int64_t* step_dev, float grad_scale,
                                     gdmcf_stream_t stream) {
  if (!p || !m || !v || !last_step || n_rows <= 0 || cols <= 0 || (step < 1 && !step_dev) ||
      (idx && n_sel <= 0) || (grad_rows && (!idx || ld_g < cols))) {
    set_error("adamw_rows_lazy: bad arguments");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  const int n = idx ? n_sel : n_rows;
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  launch_kernel(adamw_rows_lazy_kernel, std::min(n, sms * 8), LAZY_THREADS, 0, reinterpret_cast<cudaStream_t>(stream), p, m, v,
                last_step, idx, grad_rows, (long long)ld_g, n, cols, lr, beta1, beta2, eps, weight_decay, grad_scale,
                reinterpret_cast<const long long*>(step_dev), step);
  return cuda_check_launch("adamw_rows_lazy_kernel");
}

// K13 — fused history-mask + top-K select per user row, and the ranking metrics on the device.
//
// mask_topk: one CTA per user row. The row is converted to order-preserving uint32 keys and staged in
// shared memory when it fits (Yelp/Amazon catalogues do: 34 395 / 94 949 items -> 134 / 371 KB... the
// latter does not, so the kernel falls back to re-reading the L2-resident row per pass), history items
// get the lowest key via a shared-memory bitmap (the score matrix is not modified, unlike the
// reference's in-place index_put of -inf at main.py:299). A 4-pass 8-bit radix select finds the exact
// K-th largest key, candidates are collected (ties by ascending item id, deterministic) and a bitonic
// network sorts the <= 1024 survivors. HBM traffic: one read of the row (+ one per pass if not staged).
//
// topn_metrics: evaluate_utils.py:6-52 per user (binary search in the sorted ground-truth CSR row), one
// warp per user, sequential accumulation in the reference's order; colsum_f64 reduces in fixed order.
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace topk {

constexpr int TPB = 512;
constexpr int MAX_K = 1024;

GD_DEV uint32_t f32_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
GD_DEV float key_f32(uint32_t k) {
  if (k == 0u) return __uint_as_float(0xFF800000u);  // masked history item: reported as -inf like main.py:299
  const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(b);
}

struct RowView {
  const float* g;        // global scores row
  const uint32_t* keys;  // staged keys (or nullptr)
  const uint32_t* bitmap;
  GD_DEV uint32_t key(int i) const {
    if (keys) return keys[i];
    if (bitmap[i >> 5] & (1u << (i & 31))) return 0u;
    return f32_key(g[i]);
  }
};

constexpr int CAND_CAP = 2048;  // candidate slots of the threshold path (>= MAX_K)

// Bitonic sort of n (power of two) (key, idx) pairs in shared memory, descending by key, ties by ascending idx.
GD_DEV void bitonic_desc(uint32_t* key, uint32_t* idx, int n, int tid) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n; i += TPB) {
        const int j = i ^ stride;
        if (j > i) {
          const uint32_t ki = key[i], kj = key[j], ii = idx[i], ij = idx[j];
          const bool i_before_j = (ki > kj) || (ki == kj && ii < ij);  // desired order: i first
          const bool desc = ((i & size) == 0);
          if (desc ? !i_before_j : i_before_j) {
            key[i] = kj; key[j] = ki; idx[i] = ij; idx[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
}

// Threshold path (k <= TPB): every thread keeps the maximum of its strided slice; the k-th largest of the TPB slice
// maxima is a lower bound L of the k-th largest key of the row (k distinct elements are >= L), so the exact top-k is
// among the elements >= L — a few dozen for real score rows. Two coalesced passes over the row, no histogram atomics.
// Rows where more than CAND_CAP elements reach L (massive ties) fall back to the exact 4-pass radix select.
__global__ void __launch_bounds__(TPB, 3)  // 3 CTAs per SM: 444 rows resident, a 400-user batch is a single wave
mask_topk_kernel(const float* __restrict__ scores, long long ld, int n_rows, int n_items, const int* __restrict__ users,
                 const int* __restrict__ h_rowptr, const int* __restrict__ h_col, const int* __restrict__ h_rowptr2,
                 const int* __restrict__ h_col2, int k, int kpad, int stage_keys, int* __restrict__ out_idx,
                 float* __restrict__ out_val) {
  pdl_entry();
  extern __shared__ uint32_t sm[];
  // layout: hist[256] | ctrl[8] | scan[TPB] | scan_idx[TPB] | cand_key[CAND_CAP] | cand_idx[CAND_CAP] | bitmap[words] | keys[n_items]?
  uint32_t* hist = sm;
  uint32_t* ctrl = hist + 256;
  uint32_t* scan = ctrl + 8;
  uint32_t* scan_idx = scan + TPB;
  uint32_t* cand_key = scan_idx + TPB;
  uint32_t* cand_idx = cand_key + CAND_CAP;
  const int words = (n_items + 31) >> 5;
  uint32_t* bitmap = cand_idx + CAND_CAP;
  uint32_t* keys = stage_keys ? bitmap + words : nullptr;
  const int tid = threadIdx.x;

  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const float* row = scores + (long long)r * ld;
    const int u = users ? users[r] : r;
    // ---- history bitmap
    for (int w = tid; w < words; w += TPB) bitmap[w] = 0u;
    __syncthreads();
    if (h_rowptr) {
      for (int j = h_rowptr[u] + tid; j < h_rowptr[u + 1]; j += TPB) {
        const int c = h_col[j];
        if (c < n_items) atomicOr(&bitmap[c >> 5], 1u << (c & 31));
      }
    }
    if (h_rowptr2) {
      for (int j = h_rowptr2[u] + tid; j < h_rowptr2[u + 1]; j += TPB) {
        const int c = h_col2[j];
        if (c < n_items) atomicOr(&bitmap[c >> 5], 1u << (c & 31));
      }
    }
    __syncthreads();
    RowView rv{row, keys, bitmap};
    int n_sort = 0;  // number of candidate slots to sort (power of two); 0 -> radix fallback

    if (k <= TPB) {
      // ---- pass 1: slice maxima (and key staging when the row fits in shared memory)
      uint32_t tmax = 0u;
#pragma unroll 4
      for (int i = tid; i < n_items; i += TPB) {
        const uint32_t kk = (bitmap[i >> 5] & (1u << (i & 31))) ? 0u : f32_key(__ldg(row + i));
        if (keys) keys[i] = kk;
        tmax = max(tmax, kk);
      }
      scan[tid] = tmax;
      scan_idx[tid] = (uint32_t)tid;
      if (tid == 0) ctrl[2] = 0u;
      __syncthreads();
      bitonic_desc(scan, scan_idx, TPB, tid);
      const uint32_t L = scan[k - 1];
      // ---- pass 2: collect everything >= L
      if (L > 0u) {
#pragma unroll 4
        for (int i = tid; i < n_items; i += TPB) {
          const uint32_t kk = rv.key(i);
          if (kk >= L) {
            const uint32_t slot = atomicAdd(&ctrl[2], 1u);
            if (slot < (uint32_t)CAND_CAP) { cand_key[slot] = kk; cand_idx[slot] = (uint32_t)i; }
          }
        }
      }
      __syncthreads();
      const uint32_t cnt = ctrl[2];
      if (L > 0u && cnt <= (uint32_t)CAND_CAP) {
        n_sort = 2;
        while (n_sort < (int)cnt) n_sort <<= 1;
        for (int i = (int)cnt + tid; i < n_sort; i += TPB) { cand_key[i] = 0u; cand_idx[i] = 0x7FFFFFFFu; }
      }
      __syncthreads();
    } else if (keys) {
      for (int i = tid; i < n_items; i += TPB)
        keys[i] = (bitmap[i >> 5] & (1u << (i & 31))) ? 0u : f32_key(row[i]);
      __syncthreads();
    }

    if (n_sort == 0) {
      // ---- exact radix select of the k-th largest key (4 passes of 8 bits)
      uint32_t prefix = 0u, known = 0u;
      int remaining = k;
      for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 256; b += TPB) hist[b] = 0u;
        __syncthreads();
        for (int i = tid; i < n_items; i += TPB) {
          const uint32_t kk = rv.key(i);
          if ((kk & known) == prefix) atomicAdd(&hist[(kk >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
          // warp 0: lanes own 8 bins each, scanned from the top bin down
          uint32_t local[8];
          uint32_t lsum = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) { local[j] = hist[255 - (tid * 8 + j)]; lsum += local[j]; }
          uint32_t incl = lsum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += n;
          }
          uint32_t before = incl - lsum;  // count in bins above this lane's range
          if (before < (uint32_t)remaining && incl >= (uint32_t)remaining) {
            uint32_t cum = before;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (cum < (uint32_t)remaining && cum + local[j] >= (uint32_t)remaining) {
                ctrl[0] = 255 - (tid * 8 + j);
                ctrl[1] = cum;
              }
              cum += local[j];
            }
          }
        }
        __syncthreads();
        prefix |= ctrl[0] << shift;
        known |= 255u << shift;
        remaining -= (int)ctrl[1];
        __syncthreads();
      }
      // prefix = key of the k-th largest element; `remaining` of the elements equal to it are selected.
      for (int i = tid; i < kpad; i += TPB) { cand_key[i] = 0u; cand_idx[i] = 0x7FFFFFFFu; }
      if (tid == 0) ctrl[2] = 0u;
      __syncthreads();
      for (int i = tid; i < n_items; i += TPB) {
        const uint32_t kk = rv.key(i);
        if (kk > prefix) {
          const uint32_t slot = atomicAdd(&ctrl[2], 1u);
          cand_key[slot] = kk;
          cand_idx[slot] = (uint32_t)i;
        }
      }
      // ties: ordered by item id. Each thread owns a contiguous index range.
      const int per = (n_items + TPB - 1) / TPB;
      const int lo = min(n_items, tid * per), hi = min(n_items, lo + per);
      uint32_t mine = 0;
      for (int i = lo; i < hi; ++i) mine += (rv.key(i) == prefix) ? 1u : 0u;
      scan[tid] = mine;
      __syncthreads();
      // Hillis-Steele inclusive scan over TPB entries
      for (int o = 1; o < TPB; o <<= 1) {
        const uint32_t v = (tid >= o) ? scan[tid - o] : 0u;
        __syncthreads();
        scan[tid] += v;
        __syncthreads();
      }
      const uint32_t n_gt = ctrl[2];  // == k - remaining
      uint32_t rank = scan[tid] - mine;
      for (int i = lo; i < hi && rank < (uint32_t)remaining; ++i) {
        if (rv.key(i) == prefix) {
          cand_key[n_gt + rank] = prefix;
          cand_idx[n_gt + rank] = (uint32_t)i;
          ++rank;
        }
      }
      __syncthreads();
      n_sort = kpad;
    }

    // ---- sort the candidates, descending by (key, -idx); the first k are the answer
    bitonic_desc(cand_key, cand_idx, n_sort, tid);
    for (int i = tid; i < k; i += TPB) {
      out_idx[(long long)r * k + i] = (int)cand_idx[i];
      if (out_val) out_val[(long long)r * k + i] = key_f32(cand_key[i]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// computeTopNAccuracy per user (evaluate_utils.py:6-52)
// ---------------------------------------------------------------------------------------------
constexpr int MW = 4;  // warps per CTA

__global__ void __launch_bounds__(MW * 32)
topn_metrics_kernel(const int* __restrict__ topk_idx, int ld_idx, int n_rows, const int* __restrict__ users,
                    const int* __restrict__ gt_rowptr, const int* __restrict__ gt_col, const int* __restrict__ topn,
                    int n_topn, int max_n, double* __restrict__ stats) {
  pdl_entry();
  extern __shared__ double smd[];
  double* inv_log = smd;                                                  // [max_n] 1/log2(j+2)
  uint32_t* hitbits = reinterpret_cast<uint32_t*>(inv_log + max_n);        // [MW][ceil(max_n/32)]
  const int hw = (max_n + 31) >> 5;
  for (int j = threadIdx.x; j < max_n; j += blockDim.x) inv_log[j] = 1.0 / log2((double)(j + 2));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* hb = hitbits + warp * hw;
  for (int r = blockIdx.x * MW + warp; r < n_rows; r += gridDim.x * MW) {
    const int u = users ? users[r] : r;
    const int gb = gt_rowptr[u], ge = gt_rowptr[u + 1];
    const int glen = ge - gb;
    for (int p0 = 0; p0 < max_n; p0 += 32) {
      const int p = p0 + lane;
      bool hit = false;
      if (p < max_n && glen > 0) {
        const int item = topk_idx[(long long)r * ld_idx + p];
        int lo = gb, hi = ge;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const int v = gt_col[mid];
          if (v < item) lo = mid + 1; else hi = mid;
        }
        hit = (lo < ge && gt_col[lo] == item);
      }
      const uint32_t bits = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) hb[p0 >> 5] = bits;
    }
    __syncwarp();
    for (int j = lane; j < n_topn; j += 32) {
      const int N = topn[j];
      double dcg = 0.0, idcg = 0.0, mrr = 0.0;
      int hits = 0, idcg_count = glen;
      bool mrr_flag = true;
      if (glen > 0) {
        for (int p = 0; p < min(N, max_n); ++p) {
          if (hb[p >> 5] & (1u << (p & 31))) {
            dcg += inv_log[p];
            if (mrr_flag) { mrr = 1.0 / ((double)p + 1.0); mrr_flag = false; }
            ++hits;
          }
          if (idcg_count > 0) { idcg += inv_log[p]; --idcg_count; }
        }
      }
      double* o = stats + ((long long)r * n_topn + j) * 4;
      o[0] = glen > 0 ? (double)hits / (double)N : 0.0;
      o[1] = glen > 0 ? (double)hits / (double)glen : 0.0;
      o[2] = (glen > 0 && idcg != 0.0) ? dcg / idcg : 0.0;
      o[3] = mrr;
    }
    __syncwarp();
  }
}

// out[c] = sum_r x[r, c], sequential per thread then fixed-shape tree: deterministic for fixed (rows, cols).
__global__ void colsum_f64_kernel(const double* __restrict__ x, int rows, int cols, double* __restrict__ out) {
  pdl_entry();
  __shared__ double red[256];
  for (int c = blockIdx.x; c < cols; c += gridDim.x) {
    double s = 0.0;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) s += x[(long long)r * cols + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = red[0];
    __syncthreads();
  }
}

}  // namespace topk
}  // namespace gd

using namespace gd;
using namespace gd::topk;

extern "C" int gdmcf_mask_topk(const float* scores, int64_t ld, int n_rows, int n_items, const int32_t* users,
                               const int32_t* hist_rowptr, const int32_t* hist_col, const int32_t* hist_rowptr2,
                               const int32_t* hist_col2, int k, int32_t* out_idx, float* out_val, gdmcf_stream_t stream) {
  if (!scores || !out_idx || n_rows <= 0 || n_items <= 0 || k <= 0 || k > MAX_K || k > n_items || ld < n_items ||
      ((hist_rowptr == nullptr) != (hist_col == nullptr)) || ((hist_rowptr2 == nullptr) != (hist_col2 == nullptr))) {
    set_error("mask_topk: bad arguments (1 <= k <= min(1024, n_items), ld >= n_items)");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int kpad = 2;
  while (kpad < k) kpad <<= 1;
  const int words = (n_items + 31) >> 5;
  const size_t base_bytes = (size_t)(256 + 8 + 2 * TPB + 2 * CAND_CAP + words) * 4;
  const size_t staged_bytes = base_bytes + (size_t)n_items * 4;
  const size_t limit = 200 * 1024;
  if (base_bytes > limit) { set_error("mask_topk: catalogue too wide for the history bitmap (%d items)", n_items); return GDMCF_EBADARG; }
  // Keys are staged in shared memory only on the radix-only path (k > 512); the threshold path re-reads the L2-resident
  // row once instead, which keeps shared memory at ~25 KB and four CTAs resident per SM.
  const int stage = (k > TPB && staged_bytes <= limit) ? 1 : 0;
  const size_t smem = stage ? staged_bytes : base_bytes;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t err = cudaFuncSetAttribute(mask_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(mask_topk)");
    attr_smem = limit;
  }
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const int grid = std::min(n_rows, sms * 4);
  launch_kernel(mask_topk_kernel, grid, TPB, smem, st, scores, ld, n_rows, n_items, users, hist_rowptr, hist_col, hist_rowptr2,
                                            hist_col2, k, kpad, stage, out_idx, out_val);
  return cuda_check_launch("mask_topk_kernel");
}

extern "C" int gdmcf_colsum_f64(const double* x, int rows, int cols, double* out, gdmcf_stream_t stream) {
  if (!x || !out || rows <= 0 || cols <= 0) { set_error("colsum_f64: bad arguments"); return GDMCF_EBADARG; }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  launch_kernel(colsum_f64_kernel, std::min(cols, 1024), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, rows, cols, out);
  return cuda_check_launch("colsum_f64_kernel");
}

extern "C" int gdmcf_topn_metrics(const int32_t* topk_idx, int ld_idx, int n_rows, const int32_t* users,
                                  const int32_t* gt_rowptr, const int32_t* gt_col, const int32_t* topn, int n_topn,
                                  double* stats, gdmcf_stream_t stream) {
  // `topn` is a DEVICE array of n_topn cutoffs; the largest cutoff must be <= ld_idx. The host passes
  // that maximum implicitly through ld_idx (cutoffs beyond ld_idx would read out of bounds).
  if (!topk_idx || !gt_rowptr || !gt_col || !topn || !stats || n_rows <= 0 || n_topn <= 0 || n_topn > 32 || ld_idx <= 0 ||
      ld_idx > MAX_K) {
    set_error("topn_metrics: bad arguments (n_topn <= 32, ld_idx <= 1024)");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  const int max_n = ld_idx;
  const size_t smem = (size_t)max_n * 8 + (size_t)MW * ((max_n + 31) / 32) * 4;
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const int grid = std::min((n_rows + MW - 1) / MW, sms * 8);
  launch_kernel(topn_metrics_kernel, grid, MW * 32, smem, reinterpret_cast<cudaStream_t>(stream), topk_idx, ld_idx, n_rows, users, gt_rowptr,
                                                                                     gt_col, topn, n_topn, max_n, stats);
  return cuda_check_launch("topn_metrics_kernel");
}

// Small glue kernels of the denoiser (O(B*d) work, never on the roofline): per-timestep bias tables,
// bias+activation rows, embedding row gathers, and a plain fp32 CUDA-core GEMM for the tiny contractions
// (time-embedding columns, nt_xent logits) where a 128x256 tensor-core tile would be >90 % padding.
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace small {

constexpr int TPB = 256;

static int grid_1d(long long work_items, int per_cta = TPB) {
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const long long ctas = (work_items + per_cta - 1) / per_cta;
  return (int)std::max<long long>(1, std::min<long long>(ctas, (long long)sms * 8));
}

// timestep_embedding (models/DNN.py:1806-1825) for integer t, dim e: [cos(t f_k) | sin(t f_k) | 0 if odd]
GD_DEV float temb_at(int t, int j, int e) {
  const int half = e / 2;
  if (j >= 2 * half) return 0.f;
  const int k = j < half ? j : j - half;
  const float freq = expf(-logf(10000.0f) * (float)k / (float)half);
  const float arg = (float)t * freq;
  return j < half ? cosf(arg) : sinf(arg);
}

// emb_table[t, i] = b_emb[i] + sum_j w_emb[i, j] * temb(t)[j]        (emb_layer, models/DNN.py:24/1122)
__global__ void time_emb_table_kernel(const float* __restrict__ w_emb, const float* __restrict__ b_emb, int T, int e,
                                      float* __restrict__ emb_table) {
  pdl_entry();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T * e; i += gridDim.x * blockDim.x) {
    const int t = i / e, r = i % e;
    float s = b_emb[r];
    for (int j = 0; j < e; ++j) s += w_emb[r * e + j] * temb_at(t, j, e);
    emb_table[i] = s;
  }
}
// bias_table[t, k] = b[k] + sum_j w_time[k, j] * emb_table[t, j]   (the `cat([x, emb])` columns of the first layer)
__global__ void time_bias_table_kernel(const float* __restrict__ emb_table, const float* __restrict__ w_time, long long ld_w,
                                       const float* __restrict__ b, int T, int e, int d, float* __restrict__ out, long long ld_out) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)T * d; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / d), k = (int)(i % d);
    float s = b ? b[k] : 0.f;
    for (int j = 0; j < e; ++j) s += w_time[(long long)k * ld_w + j] * emb_table[t * e + j];
    out[(long long)t * ld_out + k] = s;
  }
}

GD_DEV void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// out[r, c] = act(in[r, c] + bias[t(r), c])
__global__ void bias_act_rows_kernel(const float* __restrict__ in, long long ld_in, const float* __restrict__ bias,
                                     long long ld_bias, const int* __restrict__ row_t, int t_const, int act,
                                     float* __restrict__ out_f32, long long ld_of, __nv_bfloat16* __restrict__ out_hi,
                                     __nv_bfloat16* __restrict__ out_lo, long long ld_ob, int rows, int cols) {
  pdl_entry();
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const int t = row_t ? row_t[r] : t_const;
    float v = in[(long long)r * ld_in + c] + (bias ? bias[(long long)t * ld_bias + c] : 0.f);
    if (act == GDMCF_ACT_TANH) v = tanhf(v);
    else if (act == GDMCF_ACT_RELU) v = fmaxf(v, 0.f);
    if (out_f32) out_f32[(long long)r * ld_of + c] = v;
    if (out_hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      out_hi[(long long)r * ld_ob + c] = h;
      if (out_lo) out_lo[(long long)r * ld_ob + c] = l;
    }
  }
}

// out[r, :] = table[idx[r], :]    (nn.Embedding lookup, models/DNN.py:1265)
__global__ void gather_rows_kernel(const float* __restrict__ table, long long ld_t, const int* __restrict__ idx,
                                   float* __restrict__ out_f32, long long ld_of, __nv_bfloat16* __restrict__ out_hi,
                                   __nv_bfloat16* __restrict__ out_lo, long long ld_ob, int rows, int cols) {
  pdl_entry();
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = table[(long long)idx[r] * ld_t + c];
    if (out_f32) out_f32[(long long)r * ld_of + c] = v;
    if (out_hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      out_hi[(long long)r * ld_ob + c] = h;
      if (out_lo) out_lo[(long long)r * ld_ob + c] = l;
    }
  }
}

// C[m, n] = alpha * sum_k opA(A)[m, k] * opB(B)[k, n] + beta * C[m, n]   (fp32, 32x32 tiles through smem)
// ta == 0: A is [M, K] (lda); ta == 1: A is [K, M].  tb == 0: B is [K, N] (ldb); tb == 1: B is [N, K].
__global__ void sgemm_small_kernel(const float* __restrict__ A, long long lda, int ta, const float* __restrict__ B,
                                   long long ldb, int tb, float* __restrict__ C, long long ldc, int M, int N, int K,
                                   float alpha, float beta) {
  pdl_entry();
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int tiles_n = (N + 31) / 32;
  const int ntiles = ((M + 31) / 32) * tiles_n;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = (tile / tiles_n) * 32, n0 = (tile % tiles_n) * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
      for (int j = ty; j < 32; j += 8) {
        // sa[m][k], sb[k][n]
        float va = 0.f, vb = 0.f;
        if (!ta) { const int m = m0 + j, k = k0 + tx; if (m < M && k < K) va = A[(long long)m * lda + k]; sa[j][tx] = va; }
        else     { const int k = k0 + j, m = m0 + tx; if (m < M && k < K) va = A[(long long)k * lda + m]; sa[tx][j] = va; }
        if (!tb) { const int k = k0 + j, n = n0 + tx; if (k < K && n < N) vb = B[(long long)k * ldb + n]; sb[j][tx] = vb; }
        else     { const int n = n0 + j, k = k0 + tx; if (k < K && n < N) vb = B[(long long)n * ldb + k]; sb[tx][j] = vb; }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int mi = ty + 8 * q;
        float s = acc[q];
#pragma unroll 8
        for (int k = 0; k < 32; ++k) s = fmaf(sa[mi][k], sb[k][tx], s);
        acc[q] = s;
      }
      __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = m0 + ty + 8 * q, n = n0 + tx;
      if (m < M && n < N) {
        float* c = C + (long long)m * ldc + n;
        *c = alpha * acc[q] + (beta != 0.f ? beta * (*c) : 0.f);
      }
    }
  }
}

// out[c] = sum_r x[r, c]: 32 columns x 8 row-slices per CTA, fixed summation order (deterministic).
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, long long ld, int rows, int cols, float* __restrict__ out) {
  pdl_entry();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int c0 = blockIdx.x * 32; c0 < cols; c0 += gridDim.x * 32) {
    const int c = c0 + tx;
    float s = 0.f;
    if (c < cols) {
      // 4 independent partial sums per thread (rows ty, ty + 8, ty + 16, ty + 24 mod 32): the loads of a trip are in
      // flight together instead of one dependent add per L2 round trip; combined in a fixed order
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int r = ty;
      for (; r + 24 < rows; r += 32) {
        const float a0 = x[(long long)r * ld + c], a1 = x[(long long)(r + 8) * ld + c];
        const float a2 = x[(long long)(r + 16) * ld + c], a3 = x[(long long)(r + 24) * ld + c];
        s0 += a0; s1 += a1; s2 += a2; s3 += a3;
      }
      for (; r < rows; r += 8) s0 += x[(long long)r * ld + c];
      s = (s0 + s1) + (s2 + s3);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += red[j][tx];
      out[c] = t;
    }
    __syncthreads();
  }
}

// Skinny contractions (n <= 16): C[m, 0..n) = alpha * sum_k A[m,k] * B[k, 0..n) + beta * C.
// NN form (A is [M,K]): one warp per output row, lanes split K, warp-reduce.
template <int NMAX>
__global__ void __launch_bounds__(256)
skinny_nn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, int tb,
                 float* __restrict__ C, long long ldc, int M, int N, int K, float alpha, float beta) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int m = warp0; m < M; m += nwarps) {
    float acc[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) acc[j] = 0.f;
    // Branch-free inner loops: column j >= N re-reads column N - 1 (a valid address) into an accumulator that is never
    // stored, so the loads of a trip are independent of any predicate and issue back to back.
    const long long sk = tb ? 1 : ldb, sj = tb ? ldb : 1;
    for (int k0 = lane; k0 < K; k0 += 128) {  // 4 K-steps per trip, their loads issued together (same summation order)
      float a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = (k0 + 32 * u < K) ? A[(long long)m * lda + k0 + 32 * u] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = min(k0 + 32 * u, K - 1);   // past the end: a[u] = 0 multiplies a valid element
        float bv[NMAX];
        const float* b = B + (long long)k * sk;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) bv[j] = b[(long long)min(j, N - 1) * sj];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) acc[j] = fmaf(a[u], bv[j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < NMAX; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
      for (int j = 0; j < N; ++j) {
        float* c = C + (long long)m * ldc + j;
        *c = alpha * acc[j] + (beta != 0.f ? beta * (*c) : 0.f);
      }
    }
  }
}
// TN form (A stored [K,M]): 32 output rows x 8 K-slices per CTA (coalesced over m), smem reduce in fixed order.
// The skinny operand B [K, N] (16 KB for the time-embedding products) is staged in shared memory once per CTA when it
// fits, so the K loop issues only the independent A loads (4 in flight per thread) instead of N dependent global loads.
template <int NMAX>
__global__ void __launch_bounds__(256)
skinny_tn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, int tb,
                 float* __restrict__ C, long long ldc, int M, int N, int K, float alpha, float beta, int stage_b) {
  pdl_entry();
  __shared__ float red[8][NMAX][33];
  extern __shared__ float sB[];  // [K][N] when stage_b
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (stage_b) {
#pragma unroll 4
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
      const int k = i / N, j = i - k * N;
      sB[i] = tb ? B[(long long)j * ldb + k] : B[(long long)k * ldb + j];
    }
    __syncthreads();
  }
  for (int m0 = blockIdx.x * 32; m0 < M; m0 += gridDim.x * 32) {
    const int m = m0 + tx;
    float acc[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) acc[j] = 0.f;
    if (m < M) {
      if (stage_b) {
        int k = ty;
        for (; k + 24 < K; k += 32) {  // 4 K-slices of this thread per trip: independent A loads
          float a[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) a[u] = A[(long long)(k + 8 * u) * lda + m];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float* b = sB + (k + 8 * u) * N;  // same address for the whole warp: broadcast
#pragma unroll
            for (int j = 0; j < NMAX; ++j) acc[j] = fmaf(a[u], b[min(j, N - 1)], acc[j]);
          }
        }
        for (; k < K; k += 8) {
          const float a = A[(long long)k * lda + m];
          const float* b = sB + k * N;
#pragma unroll
          for (int j = 0; j < NMAX; ++j) acc[j] = fmaf(a, b[min(j, N - 1)], acc[j]);
        }
      } else {
        for (int k = ty; k < K; k += 8) {
          const float a = A[(long long)k * lda + m];
#pragma unroll
          for (int j = 0; j < NMAX; ++j)
            if (j < N) acc[j] = fmaf(a, tb ? B[(long long)j * ldb + k] : B[(long long)k * ldb + j], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NMAX; ++j) red[ty][j][tx] = acc[j];
    __syncthreads();
    if (ty == 0 && m < M) {
      for (int j = 0; j < N; ++j) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][j][tx];
        float* c = C + (long long)m * ldc + j;
        *c = alpha * t + (beta != 0.f ? beta * (*c) : 0.f);
      }
    }
    __syncthreads();
  }
}

}  // namespace small
}  // namespace gd

using namespace gd;
using namespace gd::small;

#define GD_PRE()                 \
  int rc = gdmcf_device_check(); \
  if (rc) return rc;             \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" int gdmcf_time_bias_table(const float* w_emb, const float* b_emb, const float* w_time, int64_t ld_w,
                                     const float* b, int T, int e, int d, float* emb_table, float* out, int64_t ld_out,
                                     gdmcf_stream_t stream) {
  if (!w_emb || !b_emb || !w_time || !emb_table || !out || T <= 0 || e <= 0 || d <= 0 || ld_w < e || ld_out < d) {
    set_error("time_bias_table: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(time_emb_table_kernel, grid_1d(T * e), TPB, 0, st, w_emb, b_emb, T, e, emb_table);
  if ((rc = cuda_check_launch("time_emb_table_kernel"))) return rc;
  launch_kernel(time_bias_table_kernel, grid_1d((long long)T * d), TPB, 0, st, emb_table, w_time, ld_w, b, T, e, d, out, ld_out);
  return cuda_check_launch("time_bias_table_kernel");
}

extern "C" int gdmcf_bias_act_rows(const float* in, int64_t ld_in, const float* bias, int64_t ld_bias, const int32_t* row_t,
                                   int t_const, int act, float* out_f32, int64_t ld_of, void* out_bf16, void* out_lo,
                                   int64_t ld_ob, int rows, int cols, gdmcf_stream_t stream) {
  if (!in || rows <= 0 || cols <= 0 || ld_in < cols || (!out_f32 && !out_bf16) || (out_f32 && ld_of < cols) ||
      (out_bf16 && ld_ob < cols) || (out_lo && !out_bf16)) {
    set_error("bias_act_rows: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(bias_act_rows_kernel, grid_1d((long long)rows * cols), TPB, 0, st, in, ld_in, bias, ld_bias, row_t, t_const, act, out_f32, ld_of,
                                                                        (__nv_bfloat16*)out_bf16, (__nv_bfloat16*)out_lo, ld_ob, rows, cols);
  return cuda_check_launch("bias_act_rows_kernel");
}

extern "C" int gdmcf_gather_rows(const float* table, int64_t ld_t, const int32_t* idx, float* out_f32, int64_t ld_of,
                                 void* out_bf16, void* out_lo, int64_t ld_ob, int rows, int cols, gdmcf_stream_t stream) {
  if (!table || !idx || rows <= 0 || cols <= 0 || ld_t < cols || (!out_f32 && !out_bf16) || (out_f32 && ld_of < cols) ||
      (out_bf16 && ld_ob < cols) || (out_lo && !out_bf16)) {
    set_error("gather_rows: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(gather_rows_kernel, grid_1d((long long)rows * cols), TPB, 0, st, table, ld_t, idx, out_f32, ld_of, (__nv_bfloat16*)out_bf16,
                                                                      (__nv_bfloat16*)out_lo, ld_ob, rows, cols);
  return cuda_check_launch("gather_rows_kernel");
}

extern "C" int gdmcf_sgemm_small(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                                 int64_t ldc, int m, int n, int k, float alpha, float beta, gdmcf_stream_t stream) {
  if (!A || !B || !C || m <= 0 || n <= 0 || k <= 0 || ldc < n) { set_error("sgemm_small: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  if (n <= 16) {  // time-embedding columns: a 32x32 tile would be mostly padding and the K loop would be serial
    if (trans_a) {
      const size_t sb_bytes = (size_t)k * n * sizeof(float);
      const int stage_b = sb_bytes <= 28 * 1024 ? 1 : 0;  // + 16.9 KB static: stays under the 48 KB default limit
      launch_kernel(skinny_tn_kernel<16>, grid_1d((m + 31) / 32, 1), 256, stage_b ? sb_bytes : 0, st, A, lda, B, ldb, trans_b, C, ldc, m, n, k,
                                                                                        alpha, beta, stage_b);
      return cuda_check_launch("skinny_tn_kernel");
    }
    launch_kernel(skinny_nn_kernel<16>, grid_1d((long long)m * 32), 256, 0, st, A, lda, B, ldb, trans_b, C, ldc, m, n, k, alpha, beta);
    return cuda_check_launch("skinny_nn_kernel");
  }
  const int ntiles = ((m + 31) / 32) * ((n + 31) / 32);
  launch_kernel(sgemm_small_kernel, grid_1d(ntiles, 1), TPB, 0, st, A, lda, trans_a, B, ldb, trans_b, C, ldc, m, n, k, alpha, beta);
  return cuda_check_launch("sgemm_small_kernel");
}

extern "C" int gdmcf_colsum_f32(const float* x, int64_t ld, int rows, int cols, float* out, gdmcf_stream_t stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || ld < cols) { set_error("colsum_f32: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  launch_kernel(colsum_f32_kernel, grid_1d((cols + 31) / 32, 1), 256, 0, st, x, ld, rows, cols, out);
  return cuda_check_launch("colsum_f32_kernel");
}

// Training-side HBM-bound kernels: loss gradient with fused operand production (bf16 G, its transpose, row and
// column reductions in one pass over the [B, n_item] score matrix), bf16 transposes for the wgrad operands,
// tanh/relu backward, sumW-mix backward, nt_xent softmax rows (loss + dS), embedding-row scatter.
// Reference: models/gaussian_diffusion.py:902-953 (loss), models/DNN.py:479-508 (nt_xent), :1288, :1304-1327.
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace train {

constexpr int TPB = 256;

static int sm_count() { return gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148; }
static int grid_1d(long long work_items, int per_cta = TPB) {
  const long long ctas = (work_items + per_cta - 1) / per_cta;
  return (int)std::max<long long>(1, std::min<long long>(ctas, (long long)sm_count() * 8));
}

GD_DEV void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------------------------
// dL/d(out) for loss[b] = gs-weighted mean_i (x0 - out)^2, one pass over [B, I]:
//   g[b,i]   = gs[b] * 2 * (out[b,i] - x0[b,i]) / I
//   G[b,i]   = bf16(g * rs[b] * cs[i])            (A operand of the dgrad GEMM)
//   GT[i,b]  = same, transposed                   (A operand of the wgrad GEMM)
//   colsum[i]      = sum_b (with_out ? g*out : g)
//   rowpart[cb, b] = sum over the 32 columns of block cb of (with_out ? g*out : g)
// Block = 32 columns x all rows; 32 x 8 threads; G^T goes through a 32x33 smem tile so both stores coalesce.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loss_grad_kernel(const float* __restrict__ out, long long ld_out, const float* __restrict__ x0, long long ld_x0,
                 const float* __restrict__ gs, const float* __restrict__ rs, const float* __restrict__ cs, int with_out,
                 __nv_bfloat16* __restrict__ G, __nv_bfloat16* __restrict__ G_lo, long long ld_g,
                 __nv_bfloat16* __restrict__ GT, __nv_bfloat16* __restrict__ GT_lo, long long ld_gt,
                 float* __restrict__ colsum, float* __restrict__ rowpart, int B, int I) {
  pdl_entry();
  __shared__ float tile[32][33];
  __shared__ float colred[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float inv_i2 = 2.0f / (float)I;
  const int n_cb = (I + 31) / 32;
  for (int cb = blockIdx.x; cb < n_cb; cb += gridDim.x) {
    const int c = cb * 32 + tx;
    const float csc = (cs && c < I) ? cs[c] : 1.0f;
    float colacc = 0.f;
    for (int r0 = 0; r0 < B; r0 += 32) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = r0 + ty + 8 * q;
        float gsv = 0.f, red = 0.f;
        if (r < B && c < I) {
          const float o = out[(long long)r * ld_out + c];
          const float g = gs[r] * ((o - x0[(long long)r * ld_x0 + c]) * inv_i2);
          red = with_out ? g * o : g;
          gsv = g * (rs ? rs[r] : 1.0f) * csc;
          __nv_bfloat16 hi, lo;
          split_bf16(gsv, hi, lo);
          G[(long long)r * ld_g + c] = hi;
          if (G_lo) G_lo[(long long)r * ld_g + c] = lo;
        }
        colacc += red;
        const float rsum = warp_sum(red);
        if (tx == 0 && r < B && rowpart) rowpart[(long long)cb * B + r] = rsum;
        tile[ty + 8 * q][tx] = gsv;
      }
      __syncthreads();
      if (GT) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = cb * 32 + ty + 8 * q, rr = r0 + tx;
          if (cc < I && rr < B) {
            __nv_bfloat16 hi, lo;
            split_bf16(tile[tx][ty + 8 * q], hi, lo);
            GT[(long long)cc * ld_gt + rr] = hi;
            if (GT_lo) GT_lo[(long long)cc * ld_gt + rr] = lo;
          }
        }
      }
      __syncthreads();
    }
    colred[ty][tx] = colacc;
    __syncthreads();
    if (ty == 0 && c < I && colsum) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += colred[j][tx];
      colsum[c] = s;
    }
    __syncthreads();
  }
}

// bf16 [rows, cols] (ld_in) -> bf16 [cols, rows] (ld_out); 32x32 tiles.
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out, long long ld_out,
                      int rows, int cols) {
  pdl_entry();
  __shared__ __nv_bfloat16 tile[32][34];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tr = (rows + 31) / 32, tc = (cols + 31) / 32;
  for (long long t = blockIdx.x; t < (long long)tr * tc; t += gridDim.x) {
    const int r0 = (int)(t % tr) * 32, c0 = (int)(t / tr) * 32;
    for (int j = ty; j < 32; j += 8) {
      const int r = r0 + j, c = c0 + tx;
      tile[j][tx] = (r < rows && c < cols) ? in[(long long)r * ld_in + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j, r = r0 + tx;
      if (c < cols && r < rows) out[(long long)c * ld_out + r] = tile[tx][j];
    }
    __syncthreads();
  }
}

// Same, 64x64 tiles with 16 B global accesses on both sides (needs 16 B-aligned bases and leading dimensions that are
// multiples of 8; reads the zero K-padding of the input up to ld_in, writes zero padding up to ld_out).
__global__ void __launch_bounds__(256)
transpose_bf16_vec_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                          long long ld_out, int rows, int cols) {
  pdl_entry();
  __shared__ uint32_t tile[64][33];  // 64 rows x 64 bf16 (32 words) + 1 pad word
  const int tid = threadIdx.x;
  const int tr = (rows + 63) / 64, tc = (cols + 63) / 64;
  for (long long t = blockIdx.x; t < (long long)tr * tc; t += gridDim.x) {
    const int r0 = (int)(t % tr) * 64, c0 = (int)(t / tr) * 64;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int rl = pass * 32 + (tid >> 3), c8 = (tid & 7) * 8;
      const int r = r0 + rl, c = c0 + c8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < rows && c + 8 <= ld_in) v = *reinterpret_cast<const uint4*>(in + (long long)r * ld_in + c);
      uint32_t* dst = &tile[rl][c8 >> 1];
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int cl = pass * 32 + (tid >> 3), r8 = (tid & 7) * 8;
      const int c = c0 + cl, r = r0 + r8;
      if (c < cols && r + 8 <= ld_out) {
        uint32_t h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t w = tile[r8 + j][cl >> 1];
          h[j] = (cl & 1) ? (w >> 16) : (w & 0xffffu);
        }
        *reinterpret_cast<uint4*>(out + (long long)c * ld_out + r) =
            make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
      }
    }
    __syncthreads();
  }
}

// out = op(a, b):  0: a * (b > 0)   1: a * (1 - b*b)   2: alpha*a + beta*b      (fp32 and/or bf16 hi(+lo) outputs)
__global__ void ew_binary_kernel(int op, const float* __restrict__ a, long long ld_a, const float* __restrict__ b, long long ld_b,
                                 float alpha, float beta, float* __restrict__ out_f32, long long ld_of,
                                 __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ld_ob, int rows,
                                 int cols) {
  pdl_entry();
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float x = a[(long long)r * ld_a + c], y = b[(long long)r * ld_b + c];
    float v;
    if (op == 0) v = y > 0.f ? x : 0.f;
    else if (op == 1) v = x * (1.0f - y * y);
    else v = alpha * x + beta * y;
    if (out_f32) out_f32[(long long)r * ld_of + c] = v;
    if (out_hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      out_hi[(long long)r * ld_ob + c] = h;
      if (out_lo) out_lo[(long long)r * ld_ob + c] = l;
    }
  }
}

// hc' = w*hc + (1-w)*g2  ->  d_hc = w * d_hcp ; d_g2 = (1-w) * d_hcp ; dw_rows[r] = sum_c d_hcp * (hc - g2)
__global__ void __launch_bounds__(TPB)
mix_backward_kernel(const float* __restrict__ d_hcp, long long ld_d, const float* __restrict__ hc, long long ld_hc,
                    const float* __restrict__ g2, long long ld_g, const float* __restrict__ sumw, float* __restrict__ d_hc,
                    long long ld_dh, float* __restrict__ d_g2, long long ld_dg, float* __restrict__ dw_rows, int rows, int cols) {
  pdl_entry();
  __shared__ float red[TPB / 32];
  const float w = sumw[0];
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    float acc = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      const float d = d_hcp[(long long)r * ld_d + c];
      acc += d * (hc[(long long)r * ld_hc + c] - g2[(long long)r * ld_g + c]);
      d_hc[(long long)r * ld_dh + c] = w * d;
      d_g2[(long long)r * ld_dg + c] = (1.0f - w) * d;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < TPB / 32; ++i) t += red[i];
      dw_rows[r] = t;
    }
    __syncthreads();
  }
}

// nt_xent rows (models/DNN.py:479-508). S = raw dot products h . h_U^T [n, n]; one CTA per row i:
//   p = softmax(S[i,:] / tau);  loss_rows[i] = -log((p_ii + eps) / sum_{j != i} p_ij)
//   dS[i,j] = dscale[0] * a_i * p_ii * (delta_ij - p_ij) / tau,  a_i = -(1/n) * (1/(p_ii + eps) + 1/sum_{j != i} p_ij)
// (gradient w.r.t. the RAW dot product, so the 1/tau is folded in; dscale = d(total loss)/d(closs)).
__global__ void __launch_bounds__(TPB)
ntxent_rows_kernel(const float* __restrict__ S, long long ld_s, int n, float tau, float eps, const float* __restrict__ dscale,
                   float* __restrict__ loss_rows, float* __restrict__ dS, long long ld_ds) {
  pdl_entry();
  __shared__ float red[TPB / 32];
  __shared__ float bc;
  const float inv_tau = 1.0f / tau;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const float* row = S + (long long)i * ld_s;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < n; j += blockDim.x) mx = fmaxf(mx, row[j] * inv_tau);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) { float m = red[0]; for (int w = 1; w < TPB / 32; ++w) m = fmaxf(m, red[w]); bc = m; }
    __syncthreads();
    mx = bc;
    float se = 0.f, sneg = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const float e = expf(row[j] * inv_tau - mx);
      se += e;
      if (j != i) sneg += e;
    }
    se = warp_sum(se);
    sneg = warp_sum(sneg);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < TPB / 32; ++w) t += red[w]; bc = t; }
    __syncthreads();
    const float denom = bc;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sneg;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < TPB / 32; ++w) t += red[w]; bc = t; }
    __syncthreads();
    const float neg = bc / denom;  // sum_{j != i} p_ij
    const float pii = expf(row[i] * inv_tau - mx) / denom;
    if (threadIdx.x == 0 && loss_rows) loss_rows[i] = -logf((pii + eps) / neg);
    if (dS) {
      const float a = -(1.0f / (float)n) * (1.0f / (pii + eps) + 1.0f / neg) * (dscale ? dscale[0] : 1.0f) * inv_tau;
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float pij = expf(row[j] * inv_tau - mx) / denom;
        dS[(long long)i * ld_ds + j] = a * pii * ((j == i ? 1.0f : 0.0f) - pij);
      }
    }
    __syncthreads();
  }
}

// grad[idx[r], :] += v[r, :]   (dense embedding gradient rows; atomics make duplicate ids safe)
__global__ void scatter_rows_add_kernel(const float* __restrict__ v, long long ld_v, const int* __restrict__ idx,
                                        float* __restrict__ grad, long long ld_g, int rows, int cols) {
  pdl_entry();
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    atomicAdd(grad + (long long)idx[r] * ld_g + c, v[(long long)r * ld_v + c]);
  }
}

// Lt_history / Lt_count update of training_losses (models/gaussian_diffusion.py:935-949) with the reference's
// sequential semantics: per timestep the new history is the last H entries of (old entries ++ this batch's losses for t,
// in batch order) — what the shift-left-and-append loop leaves behind. One warp per timestep: two ballot passes over the
// batch (count the matches, then place match number q at slot cnt_old + q - shift), no serial walk.
__global__ void __launch_bounds__(128)
lt_history_kernel(const long long* __restrict__ ts, const double* __restrict__ loss, double* __restrict__ hist,
                  long long* __restrict__ count, int B, int T, int H) {
  pdl_entry();
  __shared__ double rows[4][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.x * 4 + w;
  if (t >= T) return;  // whole warps leave together; no block-wide barrier below
  double* row = rows[w];
  const int cnt_old = (int)count[t];
  int n = 0;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    const bool match = b < B && ts[b] == (long long)t;
    n += __popc(__ballot_sync(0xffffffffu, match));
  }
  const int total = cnt_old + n;
  const int shift = max(0, total - H);
  if (lane < H) {
    const int src = lane + shift;  // surviving old entries move down by `shift`; untouched slots keep their value
    row[lane] = (src < cnt_old) ? hist[(long long)t * H + src] : hist[(long long)t * H + lane];
  }
  __syncwarp();
  int seen = 0;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    const bool match = b < B && ts[b] == (long long)t;
    const unsigned m = __ballot_sync(0xffffffffu, match);
    if (match) {
      const int q = seen + __popc(m & ((1u << lane) - 1u));
      const int pos = cnt_old + q - shift;
      if (pos >= 0) row[pos] = loss[b];  // pos < H by construction; entries pushed out by later ones are dropped
    }
    seen += __popc(m);
  }
  __syncwarp();
  if (lane < H) hist[(long long)t * H + lane] = row[lane];
  if (lane == 0) count[t] = (long long)min(total, H);
}

// Per-row loss terms of training_losses (models/gaussian_diffusion.py:906-957, START_X with reweighting) in one launch:
//   weight_b = t_b == 0 ? 1 : SNR(t_b - 1) - SNR(t_b),  SNR(t) = ac[t] / (1 - ac[t])       (:915-917, :163-165)
//   hist_b   = weight_b * mse_b                     -> Lt_history update                   (:935)
//   loss_b   = hist_b / pt_b + 0.1 * closs          -> the step's loss is mean_b(loss_b)   (:951-955)
//   g_b      = float(weight_b / pt_b * (1 / B))     = d mean(loss) / d mse_b, the seed of the backward pass
// fp64 like the reference's schedule tensors; every operation is a single correctly rounded IEEE operation in the order
// the tensor expressions evaluate them (no contraction), so the values equal the tensor-op form bit for bit.
__global__ void __launch_bounds__(256)
loss_terms_kernel(const long long* __restrict__ ts, const double* __restrict__ pt, const float* __restrict__ mse,
                  const double* __restrict__ ac, const float* __restrict__ closs, int B, int T, int reweight,
                  double* __restrict__ hist_loss, double* __restrict__ loss, float* __restrict__ g_mse) {
  pdl_entry();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long t = ts[b];
  double w = 1.0;
  if (reweight && t > 0 && t < T) {
    const double a1 = ac[t - 1], a0 = ac[t];
    w = __dsub_rn(__ddiv_rn(a1, __dsub_rn(1.0, a1)), __ddiv_rn(a0, __dsub_rn(1.0, a0)));
  }
  const double h = __dmul_rn(w, (double)mse[b]);
  hist_loss[b] = h;
  double l = __ddiv_rn(h, pt[b]);
  if (closs) l = __dadd_rn(l, (double)__fmul_rn(closs[0], 0.1f));
  loss[b] = l;
  g_mse[b] = __double2float_rn(__dmul_rn(__ddiv_rn(w, pt[b]), __ddiv_rn(1.0, (double)B)));
}

// sample_timesteps(method="importance") of models/gaussian_diffusion.py:959-986 without leaving the device: while any
// Lt_count[t] < H the draw is uniform with pt = 1 (:961-962); afterwards p = sqrt(mean(Lt_history^2)) normalised,
// mixed with uniform_prob, t ~ Categorical(p) by inverse CDF, pt = p[t] * T (:964-978). One CTA; thread 0 builds the
// CDF in index order (T <= 1024). ts_in != NULL: no draw, only pt for the given timesteps (parity tests inject ts).
__global__ void __launch_bounds__(256)
sample_timesteps_kernel(const double* __restrict__ hist, const long long* __restrict__ count, int T, int H, int B,
                        double uniform_prob, uint64_t seed, uint64_t offset0, const uint64_t* __restrict__ epoch,
                        const long long* __restrict__ ts_in, long long* __restrict__ ts, double* __restrict__ pt) {
  pdl_entry();
  __shared__ double s_p[1024];
  __shared__ double s_cdf[1024];
  int full = 1;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    full &= (count[t] == (long long)H) ? 1 : 0;
    double acc = 0.0;
    for (int j = 0; j < H; ++j) {
      const double x = hist[(long long)t * H + j];
      acc += x * x;
    }
    s_p[t] = sqrt(acc / (double)H);
  }
  const int ready = __syncthreads_and(full);
  if (threadIdx.x == 0 && ready) {
    double tot = 0.0;
    for (int t = 0; t < T; ++t) tot += s_p[t];
    double run = 0.0;
    for (int t = 0; t < T; ++t) {
      const double p = s_p[t] / tot * (1.0 - uniform_prob) + uniform_prob / (double)T;
      s_p[t] = p;
      run += p;
      s_cdf[t] = run;
    }
  }
  __syncthreads();
  const uint64_t offset = offset0;
  const uint64_t ep = epoch ? (epoch[0] << 8) : 0ull;  // high counter word = (epoch << 8) | sub-stream
  const Philox rng(seed);
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    long long t;
    if (ts_in) {
      t = ts_in[b];
    } else {
      const uint4 g = rng(offset + (uint64_t)b, ep | 7ull);
      const double u = ((double)(g.x >> 5) * 67108864.0 + (double)(g.y >> 6)) * (1.0 / 9007199254740992.0);  // [0,1), 53 bits
      if (ready) {
        const double target = u * s_cdf[T - 1];
        int lo = 0, hi = T - 1;
        while (lo < hi) {  // first t with cdf[t] > target
          const int mid = (lo + hi) >> 1;
          if (s_cdf[mid] > target) hi = mid; else lo = mid + 1;
        }
        t = lo;
      } else {
        t = min((long long)(u * (double)T), (long long)T - 1);
      }
      ts[b] = t;
    }
    pt[b] = ready ? s_p[t] * (double)T : 1.0;
  }
}

}  // namespace train
}  // namespace gd

using namespace gd;
using namespace gd::train;

extern "C" int gdmcf_lt_history_update(const int64_t* ts, const double* loss, double* lt_history, int64_t* lt_count, int batch,
                                       int steps, int history, gdmcf_stream_t stream) {
  if (!ts || !loss || !lt_history || !lt_count || batch <= 0 || steps <= 0 || history <= 0 || history > 32) {
    set_error("lt_history_update: bad arguments (history_num_per_term <= 32)");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  launch_kernel(lt_history_kernel, (steps + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const long long*>(ts), loss, lt_history, reinterpret_cast<long long*>(lt_count), batch, steps, history);
  return cuda_check_launch("lt_history_kernel");
}

extern "C" int gdmcf_loss_terms(const int64_t* ts, const double* pt, const float* mse, const double* alphas_cumprod,
                                const float* closs, int batch, int steps, int reweight, double* hist_loss, double* loss,
                                float* g_mse, gdmcf_stream_t stream) {
  if (!ts || !pt || !mse || !alphas_cumprod || !hist_loss || !loss || !g_mse || batch <= 0 || steps <= 0) {
    set_error("loss_terms: bad arguments");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  launch_kernel(loss_terms_kernel, (batch + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream),
                reinterpret_cast<const long long*>(ts), pt, mse, alphas_cumprod, closs, batch, steps, reweight, hist_loss, loss, g_mse);
  return cuda_check_launch("loss_terms_kernel");
}

extern "C" int gdmcf_sample_timesteps(const double* lt_history, const int64_t* lt_count, int steps, int history, int batch,
                                      double uniform_prob, uint64_t seed, uint64_t offset, const uint64_t* epoch_dev,
                                      const int64_t* ts_in, int64_t* ts_out, double* pt_out, gdmcf_stream_t stream) {
  if (!lt_history || !lt_count || !pt_out || (!ts_in && !ts_out) || steps <= 0 || steps > 1024 || history <= 0 || batch <= 0) {
    set_error("sample_timesteps: bad arguments (steps <= 1024)");
    return GDMCF_EBADARG;
  }
  int rc = gdmcf_device_check();
  if (rc) return rc;
  launch_kernel(sample_timesteps_kernel, 1, 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      lt_history, reinterpret_cast<const long long*>(lt_count), steps, history, batch, uniform_prob, seed, offset, epoch_dev,
      reinterpret_cast<const long long*>(ts_in), reinterpret_cast<long long*>(ts_out), pt_out);
  return cuda_check_launch("sample_timesteps_kernel");
}

#define GD_PRE()                 \
  int rc = gdmcf_device_check(); \
  if (rc) return rc;             \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" int gdmcf_loss_grad(const float* out, int64_t ld_out, const float* x0, int64_t ld_x0, const float* gs,
                               const float* row_scale, const float* col_scale, int with_out, void* g_bf16, void* g_lo,
                               int64_t ld_g, void* gt_bf16, void* gt_lo, int64_t ld_gt, float* colsum, float* rowpart,
                               int rows, int cols, gdmcf_stream_t stream) {
  if (!out || !x0 || !gs || !g_bf16 || rows <= 0 || cols <= 0 || ld_out < cols || ld_x0 < cols || ld_g < cols ||
      (gt_bf16 && ld_gt < rows) || (gt_lo && !gt_bf16)) {
    set_error("loss_grad: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  const int n_cb = (cols + 31) / 32;
  launch_kernel(loss_grad_kernel, std::min(n_cb, sm_count() * 8), 256, 0, st, out, ld_out, x0, ld_x0, gs, row_scale, col_scale, with_out,
                                                                   (__nv_bfloat16*)g_bf16, (__nv_bfloat16*)g_lo, ld_g,
                                                                   (__nv_bfloat16*)gt_bf16, (__nv_bfloat16*)gt_lo, ld_gt,
                                                                   colsum, rowpart, rows, cols);
  return cuda_check_launch("loss_grad_kernel");
}

extern "C" int gdmcf_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols,
                                    gdmcf_stream_t stream) {
  if (!in || !out || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < rows) { set_error("transpose_bf16: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  if ((ld_in & 7) == 0 && (ld_out & 7) == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
    const long long nt = (long long)((rows + 63) / 64) * ((cols + 63) / 64);
    launch_kernel(transpose_bf16_vec_kernel, grid_1d(nt, 1), 256, 0, st, (const __nv_bfloat16*)in, ld_in, (__nv_bfloat16*)out, ld_out, rows, cols);
    return cuda_check_launch("transpose_bf16_vec_kernel");
  }
  const long long nt = (long long)((rows + 31) / 32) * ((cols + 31) / 32);
  launch_kernel(transpose_bf16_kernel, grid_1d(nt, 1), 256, 0, st, (const __nv_bfloat16*)in, ld_in, (__nv_bfloat16*)out, ld_out, rows, cols);
  return cuda_check_launch("transpose_bf16_kernel");
}

extern "C" int gdmcf_ew_binary(int op, const float* a, int64_t ld_a, const float* b, int64_t ld_b, float alpha, float beta,
                               float* out_f32, int64_t ld_of, void* out_bf16, void* out_lo, int64_t ld_ob, int rows, int cols,
                               gdmcf_stream_t stream) {
  if (op < 0 || op > 2 || !a || !b || rows <= 0 || cols <= 0 || ld_a < cols || ld_b < cols || (!out_f32 && !out_bf16) ||
      (out_f32 && ld_of < cols) || (out_bf16 && ld_ob < cols) || (out_lo && !out_bf16)) {
    set_error("ew_binary: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(ew_binary_kernel, grid_1d((long long)rows * cols), TPB, 0, st, op, a, ld_a, b, ld_b, alpha, beta, out_f32, ld_of,
                                                                    (__nv_bfloat16*)out_bf16, (__nv_bfloat16*)out_lo, ld_ob, rows, cols);
  return cuda_check_launch("ew_binary_kernel");
}

extern "C" int gdmcf_mix_backward(const float* d_hcp, int64_t ld_d, const float* hc, int64_t ld_hc, const float* g2,
                                  int64_t ld_g, const float* sumw, float* d_hc, int64_t ld_dh, float* d_g2, int64_t ld_dg,
                                  float* dw_rows, int rows, int cols, gdmcf_stream_t stream) {
  if (!d_hcp || !hc || !g2 || !sumw || !d_hc || !d_g2 || !dw_rows || rows <= 0 || cols <= 0 || ld_d < cols || ld_hc < cols ||
      ld_g < cols || ld_dh < cols || ld_dg < cols) {
    set_error("mix_backward: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(mix_backward_kernel, grid_1d(rows, 1), TPB, 0, st, d_hcp, ld_d, hc, ld_hc, g2, ld_g, sumw, d_hc, ld_dh, d_g2, ld_dg, dw_rows, rows, cols);
  return cuda_check_launch("mix_backward_kernel");
}

extern "C" int gdmcf_ntxent_rows(const float* S, int64_t ld_s, int n, float tau, float eps, const float* dscale,
                                 float* loss_rows, float* dS, int64_t ld_ds, gdmcf_stream_t stream) {
  if (!S || n <= 1 || ld_s < n || tau <= 0.f || (!loss_rows && !dS) || (dS && ld_ds < n)) { set_error("ntxent_rows: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  launch_kernel(ntxent_rows_kernel, grid_1d(n, 1), TPB, 0, st, S, ld_s, n, tau, eps, dscale, loss_rows, dS, ld_ds);
  return cuda_check_launch("ntxent_rows_kernel");
}

extern "C" int gdmcf_scatter_rows_add(const float* v, int64_t ld_v, const int32_t* idx, float* grad, int64_t ld_g, int rows,
                                      int cols, gdmcf_stream_t stream) {
  if (!v || !idx || !grad || rows <= 0 || cols <= 0 || ld_v < cols || ld_g < cols) { set_error("scatter_rows_add: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  launch_kernel(scatter_rows_add_kernel, grid_1d((long long)rows * cols), TPB, 0, st, v, ld_v, idx, grad, ld_g, rows, cols);
  return cuda_check_launch("scatter_rows_add_kernel");
}

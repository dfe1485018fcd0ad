// HBM-bound kernels around the contractions: operand re-layout (fp32 -> bf16 hi/lo, transposes), CSR ->
// dense interaction rows, fused forward noising (q_sample + dropout), the 2-state discrete noise of the
// one-hot branch, the sparse one-hot encoder used at inference, user-tower finish (mix + row norms),
// loss rows, and fused AdamW. All are single-pass, vectorised, grid-stride over 148*k CTAs.
//
// Reference call sites are cited at each entry point in include/gdmcf_sm100.h.
#include "common.cuh"
#include "api_internal.h"

namespace gd {
namespace ew {

constexpr int TPB = 256;

static int grid_1d(long long work_items, int per_cta = TPB) {
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  const long long ctas = (work_items + per_cta - 1) / per_cta;
  return (int)std::max<long long>(1, std::min<long long>(ctas, (long long)sms * 8));
}

GD_DEV void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------------------------
// cast fp32 -> bf16 (hi[, lo]) with zero padding up to ld_out
// ---------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, long long ld_out, int rows, int cols) {
  pdl_entry();
  // one thread per 8 output columns: eight coalesced scalar loads (the fp32 rows of nn.Linear weights are not
  // 16 B aligned: ld_in = n_item + emb_size), one 16 B store per output
  const int groups = (int)(ld_out >> 3);  // ld_out % 8 == 0 (checked on the host)
  const long long total = (long long)rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups);
    const int c0 = (int)(i % groups) << 3;
    const float* src = in + (long long)r * ld_in + c0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < cols) ? src[j] : 0.f;
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * j], h0, l0);
      split_bf16(v[2 * j + 1], h1, l1);
      ph[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      pl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(hi + (long long)r * ld_out + c0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + (long long)r * ld_out + c0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// out[c, r] = in[r, c]; 32x32 tiles through shared memory, both sides coalesced.
__global__ void cast_bf16_transpose_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ hi,
                                           __nv_bfloat16* __restrict__ lo, long long ld_out, int rows, int cols) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int tiles_c = (cols + 31) / 32;
  const int tiles_r = (int)((ld_out + 31) / 32);  // output columns (= input rows) incl. padding
  const long long ntiles = (long long)tiles_c * tiles_r;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int r0 = (int)(t % tiles_r) * 32, c0 = (int)(t / tiles_r) * 32;
    for (int j = ty; j < 32; j += 8) {
      const int r = r0 + j, c = c0 + tx;
      tile[j][tx] = (r < rows && c < cols) ? in[(long long)r * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j, r = r0 + tx;  // output row c, output column r
      if (c < cols && r < ld_out) {
        __nv_bfloat16 h, l;
        split_bf16(tile[tx][j], h, l);
        hi[(long long)c * ld_out + r] = h;
        if (lo) lo[(long long)c * ld_out + r] = l;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// CSR rows -> dense {0,1} rows. One CTA per output row.
// ---------------------------------------------------------------------------------------------
__global__ void densify_rows_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                    const int* __restrict__ users, int n_rows, int n_items, float* __restrict__ out_f32,
                                    long long ld_f32, __nv_bfloat16* __restrict__ out_bf16, long long ld_bf16) {
  pdl_entry();
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    if (out_f32) {
      float4* p = reinterpret_cast<float4*>(out_f32 + (long long)r * ld_f32);
      for (long long i = threadIdx.x; i < ld_f32 / 4; i += blockDim.x) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (out_bf16) {
      uint4* p = reinterpret_cast<uint4*>(out_bf16 + (long long)r * ld_bf16);
      for (long long i = threadIdx.x; i < ld_bf16 / 8; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int u = users ? users[r] : r;
    const int b = rowptr[u], e = rowptr[u + 1];
    for (int j = b + threadIdx.x; j < e; j += blockDim.x) {
      const int c = col[j];
      if (c < n_items) {
        if (out_f32) out_f32[(long long)r * ld_f32 + c] = 1.0f;
        if (out_bf16) out_bf16[(long long)r * ld_bf16 + c] = __float2bfloat16_rn(1.0f);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// q_sample fused with dropout and the bf16 operand cast. One thread per 4 consecutive columns.
// ---------------------------------------------------------------------------------------------
__global__ void qsample_dropout_kernel(const float* __restrict__ x0, long long ld_x0, const int* __restrict__ row_t,
                                       int t_const, const float* __restrict__ sqrt_ab, const float* __restrict__ sqrt_1mab,
                                       const float* __restrict__ noise, const uint8_t* __restrict__ keep, float dropout_p,
                                       uint64_t seed, uint64_t offset0, const uint64_t* __restrict__ epoch,
                                       float* __restrict__ xt_f32, long long ld_xt,
                                       __nv_bfloat16* __restrict__ a_hi, __nv_bfloat16* __restrict__ a_lo, long long ld_a,
                                       int rows, int cols) {
  pdl_entry();
  // Philox counter: low word = offset0 (call site << 40) + element group, high word = (epoch << 8) | sub-stream.
  // `epoch` is a device-resident step counter, so a captured CUDA graph draws fresh numbers on every replay; it has
  // 56 bits of its own and cannot run into the call-site or element fields however long the run is.
  const uint64_t offset = offset0;
  const uint64_t ep = epoch ? (epoch[0] << 8) : 0ull;
  const int groups = (int)(ld_a / 4);  // ld_a % 8 == 0
  const long long total = (long long)rows * groups;
  const Philox rng(seed);
  const float keep_scale = dropout_p > 0.f ? 1.0f / (1.0f - dropout_p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups);
    const int c0 = (int)(i % groups) * 4;
    const int t = row_t ? row_t[r] : t_const;
    const bool noisy = sqrt_ab != nullptr;  // NULL tables: x_t = x0 (p_sample with sampling_steps == 0)
    const float ca = noisy ? sqrt_ab[t] : 1.f, cb = noisy ? sqrt_1mab[t] : 0.f;
    float eps[4] = {0.f, 0.f, 0.f, 0.f};
    uint4 ur = make_uint4(0, 0, 0, 0);
    if (noisy && !noise) {
      const uint4 g = rng(offset + (uint64_t)i, ep | 0ull);
      const float2 n0 = box_muller(g.x, g.y), n1 = box_muller(g.z, g.w);
      eps[0] = n0.x; eps[1] = n0.y; eps[2] = n1.x; eps[3] = n1.y;
    }
    if (dropout_p > 0.f && !keep) ur = rng(offset + (uint64_t)i, ep | 1ull);
    const uint32_t urr[4] = {ur.x, ur.y, ur.z, ur.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      float xt = 0.f, a = 0.f;
      if (c < cols) {
        const float x = x0[(long long)r * ld_x0 + c];
        if (noisy) {
          const float e = noise ? noise[(long long)r * cols + c] : eps[j];
          xt = ca * x + cb * e;  // same association as gaussian_diffusion.py:993-996
        } else {
          xt = x;
        }
        if (xt_f32) xt_f32[(long long)r * ld_xt + c] = xt;
        a = xt;
        if (dropout_p > 0.f) {
          const bool k = keep ? (keep[(long long)r * cols + c] != 0) : (u32_to_unit_open(urr[j]) > dropout_p);
          a = k ? xt * keep_scale : 0.f;
        }
      }
      split_bf16(a, h[j], l[j]);
    }
    __nv_bfloat16* ph = a_hi + (long long)r * ld_a + c0;
    *reinterpret_cast<uint2*>(ph) =
        make_uint2((uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
                   (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
    if (a_lo) {
      __nv_bfloat16* pl = a_lo + (long long)r * ld_a + c0;
      *reinterpret_cast<uint2*>(pl) =
          make_uint2((uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16),
                     (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Discrete (2-state) noise on the one-hot branch + dropout. One thread per 4 items -> 8 outputs.
// ---------------------------------------------------------------------------------------------
__global__ void onehot_noise_kernel(const float* __restrict__ x0, long long ld_x0, const int* __restrict__ ts,
                                    float discrete, float dropout_p, const float* __restrict__ u_keep,
                                    const float* __restrict__ u_drop, uint64_t seed, uint64_t offset0,
                                    const uint64_t* __restrict__ epoch, __nv_bfloat16* __restrict__ out, long long ld_out,
                                    int rows, int cols) {
  pdl_entry();
  const uint64_t offset = offset0;
  const uint64_t ep = epoch ? (epoch[0] << 8) : 0ull;  // high counter word = (epoch << 8) | sub-stream
  const int groups = (int)(ld_out / 8);  // 8 outputs = 4 items per thread
  const long long total = (long long)rows * groups;
  const Philox rng(seed);
  const float keep_val = dropout_p > 0.f ? 1.0f / (1.0f - dropout_p) : 1.0f;
  const __nv_bfloat16 one = __float2bfloat16_rn(keep_val), zero = __float2bfloat16_rn(0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups);
    const int i0 = (int)(i % groups) * 4;
    // a = ts / batch_size in fp32 (gaussian_diffusion.py:775); Q_bar = a*I + (1-a)*u_x (:601)
    const float a = ts ? (float)ts[r] / (float)rows : 1.0f;  // ts == NULL: no discrete noise (p_sample, steps == 0)
    const float q_one = a * 1.0f + (1.0f - a) * (1.0f - discrete);  // x0 == 1 keeps channel 1
    const float q_zero = a * 1.0f + (1.0f - a) * discrete;          // x0 == 0 keeps channel 0
    uint4 g0 = make_uint4(0, 0, 0, 0), g1 = make_uint4(0, 0, 0, 0);
    if (!u_keep && ts) g0 = rng(offset + (uint64_t)i, ep | 2ull);
    if (!u_drop && dropout_p > 0.f) g1 = rng(offset + (uint64_t)i, ep | 3ull);
    const uint32_t gk[4] = {g0.x, g0.y, g0.z, g0.w}, gd_[4] = {g1.x, g1.y, g1.z, g1.w};
    __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = i0 + j;
      o[2 * j] = zero;
      o[2 * j + 1] = zero;
      if (it < cols) {
        const int c = x0[(long long)r * ld_x0 + it] > 0.5f ? 1 : 0;
        bool kept = true;
        if (ts) {
          const float u = u_keep ? u_keep[(long long)r * cols + it] : (1.0f - u32_to_unit_open(gk[j]));  // [0,1)
          kept = u < (c ? q_one : q_zero);
        }
        if (kept && dropout_p > 0.f) {
          const float u = u_drop ? u_drop[(long long)r * 2 * cols + 2 * it + c] : (1.0f - u32_to_unit_open(gd_[j]));
          kept = u >= dropout_p;
        }
        if (kept) o[2 * j + c] = one;
      }
    }
    uint4 pk;
    pk.x = (uint32_t)__bfloat16_as_ushort(o[0]) | ((uint32_t)__bfloat16_as_ushort(o[1]) << 16);
    pk.y = (uint32_t)__bfloat16_as_ushort(o[2]) | ((uint32_t)__bfloat16_as_ushort(o[3]) << 16);
    pk.z = (uint32_t)__bfloat16_as_ushort(o[4]) | ((uint32_t)__bfloat16_as_ushort(o[5]) << 16);
    pk.w = (uint32_t)__bfloat16_as_ushort(o[6]) | ((uint32_t)__bfloat16_as_ushort(o[7]) << 16);
    *reinterpret_cast<uint4*>(out + (long long)r * ld_out + 2 * i0) = pk;
  }
}

// ---------------------------------------------------------------------------------------------
// out[r, c] = bf16(in[r, c] * col_scale[c]) (hi[, lo]); padding columns [cols, ld_out) are zeroed. Operand of the
// projected reverse loop: the first-layer weight with the items' inverse norms folded into its K dimension.
// One thread per 8 consecutive columns.
// ---------------------------------------------------------------------------------------------
__global__ void scale_cols_cast_kernel(const float* __restrict__ in, long long ld_in, const float* __restrict__ col_scale,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ld_out, int rows,
                                       int cols) {
  pdl_entry();
  const int groups = (int)(ld_out / 8);
  const long long total = (long long)rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups), c0 = (int)(i % groups) * 8;
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float v = c < cols ? in[(long long)r * ld_in + c] * col_scale[c] : 0.f;
      split_bf16(v, h[j], l[j]);
    }
    uint4 ph, pl;
    ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    ph.z = (uint32_t)__bfloat16_as_ushort(h[4]) | ((uint32_t)__bfloat16_as_ushort(h[5]) << 16);
    ph.w = (uint32_t)__bfloat16_as_ushort(h[6]) | ((uint32_t)__bfloat16_as_ushort(h[7]) << 16);
    *reinterpret_cast<uint4*>(hi + (long long)r * ld_out + c0) = ph;
    if (lo) {
      pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
      pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
      pl.z = (uint32_t)__bfloat16_as_ushort(l[4]) | ((uint32_t)__bfloat16_as_ushort(l[5]) << 16);
      pl.w = (uint32_t)__bfloat16_as_ushort(l[6]) | ((uint32_t)__bfloat16_as_ushort(l[7]) << 16);
      *reinterpret_cast<uint4*>(lo + (long long)r * ld_out + c0) = pl;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Faithful graph bookkeeping of p_sample (models/gaussian_diffusion.py:710-729), one reverse step:
//   x_i      = apply_noise(t, one_hot(state))           per entry: class 0 -> 1 w.p. (1 - a)(1 - p), a = t / batch
//   s_b      ~ Bernoulli(deg_b / max_b deg)             "degree-guided" draw per user (:711-716)
//   state   |= user_guided ? (x_i == 1 && s_b == 1) : (x_i == 1)      (:721-726, OR-accumulated over the steps)
// state: uint8 [rows, ld] (1 = edge user -> item). Entries already 1 stay 1 whatever is drawn. One thread per 4 entries.
// u_entry / u_user: optional injected uniforms in [0,1) (entry flips iff u >= P(0 -> 0); s_b = 1 iff u < deg fraction).
// ---------------------------------------------------------------------------------------------
__global__ void graph_noise_step_kernel(uint8_t* __restrict__ state, long long ld, const float* __restrict__ deg_frac, int t,
                                        int batch, float discrete, int user_guided, uint64_t seed, uint64_t offset0,
                                        const uint64_t* __restrict__ epoch, const float* __restrict__ u_entry,
                                        const float* __restrict__ u_user, int rows, int cols) {
  pdl_entry();
  const uint64_t ep = epoch ? (epoch[0] << 8) : 0ull;
  const Philox rng(seed);
  const float a = (float)t / (float)batch;                 // gaussian_diffusion.py:775
  const float q_zero = a * 1.0f + (1.0f - a) * discrete;   // P(class 0 stays 0), Q_bar = a*I + (1-a)*u_x (:601)
  const int groups = (cols + 3) / 4;
  const long long total = (long long)rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups), c0 = (int)(i % groups) * 4;
    bool guide = true;
    if (user_guided) {
      const float uu = u_user ? u_user[r] : (1.0f - u32_to_unit_open(rng(offset0 + (uint64_t)r, ep | 5ull).x));
      guide = uu < deg_frac[r];
    }
    const uint4 g = u_entry ? make_uint4(0, 0, 0, 0) : rng(offset0 + (uint64_t)i, ep | 6ull);
    const uint32_t gk[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      if (c < cols) {
        const float u = u_entry ? u_entry[(long long)r * cols + c] : (1.0f - u32_to_unit_open(gk[j]));
        if (guide && u >= q_zero) state[(long long)r * ld + c] = 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Sparse one-hot encoder (inference): S[r,:] = base + sum_{i in row} delta[i,:]
// ---------------------------------------------------------------------------------------------
// delta[i,k] = W2[k,2i+1] - W2[k,2i] via 32x32 transposing tiles; base accumulated separately.
__global__ void onehot_delta_kernel(const float* __restrict__ w2, long long ld_w, int d, int n_items,
                                    float* __restrict__ delta, long long ld_delta) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int tiles_i = (n_items + 31) / 32, tiles_k = (d + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < (long long)tiles_i * tiles_k; t += gridDim.x) {
    const int i0 = (int)(t % tiles_i) * 32, k0 = (int)(t / tiles_i) * 32;
    for (int j = ty; j < 32; j += 8) {
      const int k = k0 + j, i = i0 + tx;
      float v = 0.f;
      if (k < d && i < n_items) {
        const float2 w = *reinterpret_cast<const float2*>(w2 + (long long)k * ld_w + 2 * i);  // ld_w even, base 8B aligned
        v = w.y - w.x;
      }
      tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int i = i0 + j, k = k0 + tx;
      if (i < n_items && k < d) delta[(long long)i * ld_delta + k] = tile[tx][j];
    }
    __syncthreads();
  }
}
// base[k] = sum_i W2[k, 2i]  (fp64 accumulation, one CTA per k)
__global__ void onehot_base_kernel(const float* __restrict__ w2, long long ld_w, int d, int n_items, float* __restrict__ base) {
  pdl_entry();
  __shared__ double red[TPB / 32];
  for (int k = blockIdx.x; k < d; k += gridDim.x) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) s += (double)w2[(long long)k * ld_w + 2 * i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < TPB / 32; ++w) tot += red[w];
      base[k] = (float)tot;
    }
    __syncthreads();
  }
}
// Work unit = (user row, slice of its interactions). Most rows are one unit (their items fit one pass); rows of heavy
// users (the degree distribution has a long tail: up to n_item / 4 interactions) are cut into up to OH_MAX_SLICES slices
// that different CTAs gather concurrently — one CTA walking a 1 700-interaction row alone took 100 us at the Yelp shape.
// Slices write partial sums to `ws`; the last slice of a row to finish adds them in slice order (deterministic) and
// leaves the row's counter zeroed. Every CTA derives the unit list from rowptr itself (prefix over n_rows <= 1024 rows).
constexpr int OH_CHUNK = 64, OH_MAX_SLICES = 16, OH_SLAB = 512;

__global__ void encode_onehot_gather_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                            const int* __restrict__ users, int n_rows, const float* __restrict__ base,
                                            const float* __restrict__ delta, long long ld_delta, int d,
                                            float* __restrict__ out, long long ld_out, float* __restrict__ ws,
                                            int* __restrict__ counters) {
  pdl_entry();
  extern __shared__ int s_pref[];  // [n_rows + 1] first unit of every row
  __shared__ int s_col[OH_SLAB];
  __shared__ int s_last;
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
    const int u = users ? users[r] : r;
    const int nnz = rowptr[u + 1] - rowptr[u];
    s_pref[r + 1] = ws ? max(1, min(OH_MAX_SLICES, (nnz + OH_CHUNK - 1) / OH_CHUNK)) : 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s_pref[0] = 0;
    for (int r = 0; r < n_rows; ++r) s_pref[r + 1] += s_pref[r];
  }
  __syncthreads();
  const int n_units = s_pref[n_rows];
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    int lo = 0, hi = n_rows - 1;  // row of this unit: last r with s_pref[r] <= unit
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_pref[mid] <= unit) lo = mid; else hi = mid - 1;
    }
    const int r = lo, slice = unit - s_pref[r], n_slices = s_pref[r + 1] - s_pref[r];
    const int u = users ? users[r] : r;
    const int rb = rowptr[u], re = rowptr[u + 1];
    const int per = (re - rb + n_slices - 1) / n_slices;
    const int b = rb + slice * per, e = min(re, b + per);
    float s[4];  // up to 4 output columns per thread (d <= 4 * blockDim.x, checked on the host)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = threadIdx.x + q * blockDim.x;
      s[q] = (n_slices == 1 && k < d) ? base[k] : 0.f;
    }
    for (int j0 = b; j0 < e; j0 += OH_SLAB) {
      const int n = min(OH_SLAB, e - j0);
      __syncthreads();
      // the slice's item ids are staged in shared memory so that the delta-row gathers (coalesced over k) are
      // independent loads, 8 in flight per thread, instead of a chain id -> gather -> id -> gather
      for (int j = threadIdx.x; j < n; j += blockDim.x) s_col[j] = col[j0 + j];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = threadIdx.x + q * blockDim.x;
        if (k < d) {
          int j = 0;
          for (; j + 8 <= n; j += 8) {
            float x[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) x[t] = __ldg(delta + (long long)s_col[j + t] * ld_delta + k);
#pragma unroll
            for (int t = 0; t < 8; ++t) s[q] += x[t];
          }
          for (; j < n; ++j) s[q] += __ldg(delta + (long long)s_col[j] * ld_delta + k);
        }
      }
    }
    if (n_slices == 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = threadIdx.x + q * blockDim.x;
        if (k < d) out[(long long)r * ld_out + k] = s[q];
      }
      continue;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = threadIdx.x + q * blockDim.x;
      if (k < d) ws[(long long)unit * d + k] = s[q];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int done = atomicAdd(&counters[r], 1);
      s_last = (done == n_slices - 1) ? 1 : 0;
      if (s_last) counters[r] = 0;
      __threadfence();
    }
    __syncthreads();
    if (s_last) {
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        float acc = base[k];
        for (int t = 0; t < n_slices; ++t) acc += __ldcg(ws + (long long)(s_pref[r] + t) * d + k);
        out[(long long)r * ld_out + k] = acc;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// user tower finish: mix + row norm; plain row norms
// ---------------------------------------------------------------------------------------------
GD_DEV float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  __syncthreads();
  return tot;
}

__global__ void mix_rownorm_kernel(const float* __restrict__ hc, long long ld_hc, const float* __restrict__ g,
                                   long long ld_g, const float* __restrict__ sumw, float* __restrict__ out_f32,
                                   long long ld_of, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                   long long ld_ob, float* __restrict__ inv_norm, int rows, int cols) {
  pdl_entry();
  __shared__ float red[TPB / 32];
  const float w = (g && sumw) ? sumw[0] : 1.0f;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    float ss = 0.f;
    for (int c = threadIdx.x; c < (int)ld_ob; c += blockDim.x) {
      float v = 0.f;
      if (c < cols) {
        const float h = hc[(long long)r * ld_hc + c];
        // hc * sumW + all_embeddings[:B] * (1 - sumW)   (models/DNN.py:1288)
        v = g ? h * w + g[(long long)r * ld_g + c] * (1.0f - w) : h;
        ss += v * v;
        if (out_f32) out_f32[(long long)r * ld_of + c] = v;
      }
      if (out_hi) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        out_hi[(long long)r * ld_ob + c] = h;
        if (out_lo) out_lo[(long long)r * ld_ob + c] = l;
      }
    }
    const float tot = block_sum(ss, red);
    if (threadIdx.x == 0 && inv_norm) inv_norm[r] = 1.0f / sqrtf(tot);
  }
}

__global__ void row_inv_norm_kernel(const float* __restrict__ x, long long ld, float* __restrict__ inv_norm, int rows, int cols) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp0; r < rows; r += nwarps) {
    float ss = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float v = x[(long long)r * ld + c];
      ss += v * v;
    }
    ss = warp_sum(ss);
    if (lane == 0) inv_norm[r] = 1.0f / sqrtf(ss);
  }
}

// mean_flat((x0 - out)^2) per row (gaussian_diffusion.py:902,1194-1198): 512 threads per row, 16 B loads when both
// matrices have 16 B-aligned rows (the engine's [B, ld4] buffers do), fixed summation order per launch geometry.
__global__ void __launch_bounds__(512)
mse_rows_kernel(const float* __restrict__ out, long long ld_out, const float* __restrict__ x0, long long ld_x0, int rows,
                int cols, float* __restrict__ mse) {
  pdl_entry();
  __shared__ float red[16];
  const bool vec = (((ld_out | ld_x0) & 3) == 0) && ((((uintptr_t)out | (uintptr_t)x0) & 15) == 0);
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* po = out + (long long)r * ld_out;
    const float* px = x0 + (long long)r * ld_x0;
    float ss = 0.f;
    int c_begin = 0;
    if (vec) {
      const int n4 = cols >> 2;
      const float4* po4 = reinterpret_cast<const float4*>(po);
      const float4* px4 = reinterpret_cast<const float4*>(px);
#pragma unroll 4
      for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 a = __ldg(px4 + i), b = __ldg(po4 + i);
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        ss += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      }
      c_begin = n4 << 2;
    }
    for (int c = c_begin + threadIdx.x; c < cols; c += blockDim.x) {
      const float dlt = px[c] - po[c];
      ss += dlt * dlt;
    }
    const float tot = block_sum(ss, red);
    if (threadIdx.x == 0) mse[r] = tot / (float)cols;
  }
}

// ---------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW, amsgrad=False, maximize=False)
// ---------------------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float bc1, float bc2_sqrt, float grad_scale,
                             const long long* __restrict__ step_dev) {
  pdl_entry();
  const AdamwCoef kc = adamw_coef(lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale, step_dev);
  auto update = [&](float& param, float gr, float& mi, float& vi) { adamw_update(kc, param, gr, mi, vi); };
  // 128-bit streams (all four tensors are 16 B aligned: checked on the host), scalar tail
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    update(pp.x, gg.x, mm.x, vv.x);
    update(pp.y, gg.y, mm.y, vv.y);
    update(pp.z, gg.z, mm.z, vv.z);
    update(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pp = p[i], mm = m[i], vv = v[i];
    update(pp, g[i], mm, vv);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long inc) {
  pdl_entry();
  if (blockIdx.x == 0 && threadIdx.x == 0) ctr[0] += inc;
}

}  // namespace ew
}  // namespace gd

using namespace gd;
using namespace gd::ew;

#define GD_PRE()                               \
  int rc = gdmcf_device_check();               \
  if (rc) return rc;                           \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" int gdmcf_cast_bf16(const float* in, int64_t ld_in, void* out_hi, void* out_lo, int64_t ld_out, int rows,
                               int cols, gdmcf_stream_t stream) {
  if (!in || !out_hi || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < cols || (ld_out & 7) || ((uintptr_t)out_hi & 15) ||
      ((uintptr_t)out_lo & 15)) {
    set_error("cast_bf16: bad arguments (ld_out %% 8 == 0, 16 B aligned outputs)");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(cast_bf16_kernel, grid_1d((long long)rows * (ld_out / 8)), TPB, 0, st, in, ld_in, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_out, rows, cols);
  return cuda_check_launch("cast_bf16_kernel");
}

extern "C" int gdmcf_cast_bf16_transpose(const float* in, int64_t ld_in, void* out_hi, void* out_lo, int64_t ld_out,
                                         int rows, int cols, gdmcf_stream_t stream) {
  if (!in || !out_hi || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < rows) { set_error("cast_bf16_transpose: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  const long long ntiles = (long long)((cols + 31) / 32) * ((ld_out + 31) / 32);
  launch_kernel(cast_bf16_transpose_kernel, grid_1d(ntiles, 1), TPB, 0, st, in, ld_in, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_out, rows, cols);
  return cuda_check_launch("cast_bf16_transpose_kernel");
}

extern "C" int gdmcf_densify_rows(const int32_t* rowptr, const int32_t* col, const int32_t* users, int n_rows,
                                  int n_items, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                                  gdmcf_stream_t stream) {
  if (!rowptr || !col || n_rows <= 0 || n_items <= 0 || (!out_f32 && !out_bf16) ||
      (out_f32 && ((ld_f32 & 3) || ld_f32 < n_items || ((uintptr_t)out_f32 & 15))) ||
      (out_bf16 && ((ld_bf16 & 7) || ld_bf16 < n_items || ((uintptr_t)out_bf16 & 15)))) {
    set_error("densify_rows: need ld_f32 %% 4 == 0, ld_bf16 %% 8 == 0, ld >= n_items, 16B-aligned outputs");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(densify_rows_kernel, grid_1d(n_rows, 1), TPB, 0, st, rowptr, col, users, n_rows, n_items, out_f32, ld_f32, (__nv_bfloat16*)out_bf16, ld_bf16);
  return cuda_check_launch("densify_rows_kernel");
}

extern "C" int gdmcf_qsample_dropout(const float* x0, int64_t ld_x0, const int32_t* row_t, int t_const,
                                     const float* sqrt_ab, const float* sqrt_1mab, const float* noise,
                                     const uint8_t* keep, float dropout_p, uint64_t seed, uint64_t offset,
                                     const uint64_t* epoch_dev, float* xt_f32, int64_t ld_xt, void* a_bf16, void* a_lo,
                                     int64_t ld_a, int rows, int cols, gdmcf_stream_t stream) {
  if (!x0 || !a_bf16 || rows <= 0 || cols <= 0 || (ld_a & 7) || ld_a < cols || ld_x0 < cols || ((uintptr_t)a_bf16 & 15) ||
      ((uintptr_t)a_lo & 15) || (xt_f32 && ld_xt < cols) || dropout_p < 0.f || dropout_p >= 1.f || ((sqrt_ab == nullptr) != (sqrt_1mab == nullptr))) {
    set_error("qsample_dropout: bad arguments (ld_a %% 8 == 0, 0 <= dropout_p < 1, both or neither coefficient tables)");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(qsample_dropout_kernel, grid_1d((long long)rows * (ld_a / 4)), TPB, 0, st, 
      x0, ld_x0, row_t, t_const, sqrt_ab, sqrt_1mab, noise, keep, dropout_p, seed, offset, epoch_dev, xt_f32, ld_xt,
      (__nv_bfloat16*)a_bf16, (__nv_bfloat16*)a_lo, ld_a, rows, cols);
  return cuda_check_launch("qsample_dropout_kernel");
}

extern "C" int gdmcf_onehot_noise(const float* x0, int64_t ld_x0, const int32_t* ts, float discrete, float dropout_p,
                                  const float* u_keep, const float* u_drop, uint64_t seed, uint64_t offset,
                                  const uint64_t* epoch_dev, void* out_bf16, int64_t ld_out, int rows, int cols,
                                  gdmcf_stream_t stream) {
  if (!x0 || !out_bf16 || rows <= 0 || cols <= 0 || (ld_out & 7) || ld_out < 2LL * cols || ld_x0 < cols ||
      ((uintptr_t)out_bf16 & 15) || dropout_p < 0.f || dropout_p >= 1.f) {
    set_error("onehot_noise: bad arguments (ld_out %% 8 == 0 and >= 2*cols)");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(onehot_noise_kernel, grid_1d((long long)rows * (ld_out / 8)), TPB, 0, st, 
      x0, ld_x0, ts, discrete, dropout_p, u_keep, u_drop, seed, offset, epoch_dev, (__nv_bfloat16*)out_bf16, ld_out, rows, cols);
  return cuda_check_launch("onehot_noise_kernel");
}

extern "C" int gdmcf_onehot_tables(const float* w2, int64_t ld_w, int d, int n_items, float* base, float* delta,
                                   int64_t ld_delta, gdmcf_stream_t stream) {
  if (!w2 || !base || !delta || d <= 0 || n_items <= 0 || (ld_w & 1) || ld_w < 2LL * n_items || ld_delta < d || ((uintptr_t)w2 & 7)) {
    set_error("onehot_tables: bad arguments (ld_w even and >= 2*n_items, ld_delta >= d)");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  const long long ntiles = (long long)((n_items + 31) / 32) * ((d + 31) / 32);
  launch_kernel(onehot_delta_kernel, grid_1d(ntiles, 1), TPB, 0, st, w2, ld_w, d, n_items, delta, ld_delta);
  if ((rc = cuda_check_launch("onehot_delta_kernel"))) return rc;
  launch_kernel(onehot_base_kernel, grid_1d(d, 1), TPB, 0, st, w2, ld_w, d, n_items, base);
  return cuda_check_launch("onehot_base_kernel");
}

extern "C" size_t gdmcf_encode_onehot_gather_workspace_bytes(int n_rows, int d) {
  if (n_rows <= 0 || d <= 0) return 0;
  return (size_t)n_rows * OH_MAX_SLICES * d * sizeof(float);
}

extern "C" int gdmcf_encode_onehot_gather(const int32_t* rowptr, const int32_t* col, const int32_t* users, int n_rows,
                                          const float* base, const float* delta, int64_t ld_delta, int d, float* out,
                                          int64_t ld_out, float* workspace, size_t workspace_bytes, int32_t* counters,
                                          gdmcf_stream_t stream) {
  if (!rowptr || !col || !base || !delta || !out || n_rows <= 0 || d <= 0 || d > 4 * TPB || ld_delta < d || ld_out < d) {
    set_error("encode_onehot_gather: bad arguments (d <= 1024)");
    return GDMCF_EBADARG;
  }
  // without a workspace (or with more rows than the unit prefix can hold) every row is gathered by one CTA
  const bool split = workspace && counters && workspace_bytes >= gdmcf_encode_onehot_gather_workspace_bytes(n_rows, d);
  if (n_rows > 8192) { set_error("encode_onehot_gather: at most 8192 rows per call"); return GDMCF_EBADARG; }
  GD_PRE();
  const int sms = gdmcf_num_sms() > 0 ? gdmcf_num_sms() : 148;
  launch_kernel(encode_onehot_gather_kernel, std::min(n_rows * (split ? 2 : 1), sms * 8), TPB, (size_t)(n_rows + 1) * sizeof(int), st,
                rowptr, col, users, n_rows, base, delta, (long long)ld_delta, d, out, (long long)ld_out, split ? workspace : nullptr,
                split ? counters : nullptr);
  return cuda_check_launch("encode_onehot_gather_kernel");
}

extern "C" int gdmcf_mix_rownorm(const float* hc, int64_t ld_hc, const float* g, int64_t ld_g, const float* sumw,
                                 float* out_f32, int64_t ld_of, void* out_bf16, void* out_lo, int64_t ld_ob,
                                 float* inv_norm, int rows, int cols, gdmcf_stream_t stream) {
  if (!hc || rows <= 0 || cols <= 0 || ld_hc < cols || (g && (ld_g < cols || !sumw)) || (out_f32 && ld_of < cols) ||
      (out_bf16 && ld_ob < cols) || (!out_bf16 && out_lo)) {
    set_error("mix_rownorm: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  if (!out_bf16) ld_ob = cols;
  launch_kernel(mix_rownorm_kernel, grid_1d(rows, 1), TPB, 0, st, hc, ld_hc, g, ld_g, sumw, out_f32, ld_of, (__nv_bfloat16*)out_bf16,
                                                       (__nv_bfloat16*)out_lo, ld_ob, inv_norm, rows, cols);
  return cuda_check_launch("mix_rownorm_kernel");
}

extern "C" int gdmcf_row_inv_norm(const float* x, int64_t ld, float* inv_norm, int rows, int cols, gdmcf_stream_t stream) {
  if (!x || !inv_norm || rows <= 0 || cols <= 0 || ld < cols) { set_error("row_inv_norm: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  launch_kernel(row_inv_norm_kernel, grid_1d((long long)rows * 32), TPB, 0, st, x, ld, inv_norm, rows, cols);
  return cuda_check_launch("row_inv_norm_kernel");
}

extern "C" int gdmcf_mse_rows(const float* out, int64_t ld_out, const float* x0, int64_t ld_x0, int rows, int cols,
                              float* mse, gdmcf_stream_t stream) {
  if (!out || !x0 || !mse || rows <= 0 || cols <= 0 || ld_out < cols || ld_x0 < cols) { set_error("mse_rows: bad arguments"); return GDMCF_EBADARG; }
  GD_PRE();
  launch_kernel(mse_rows_kernel, grid_1d(rows, 1), 512, 0, st, out, ld_out, x0, ld_x0, rows, cols, mse);
  return cuda_check_launch("mse_rows_kernel");
}

extern "C" int gdmcf_adamw_fused(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, int step, const int64_t* step_dev,
                                 float grad_scale, gdmcf_stream_t stream) {
  if (!p || !g || !m || !v || n <= 0 || (step < 1 && !step_dev)) { set_error("adamw_fused: bad arguments (step counts from 1)"); return GDMCF_EBADARG; }
  if (n >= 4 && (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15)) {
    set_error("adamw_fused: p/g/m/v must be 16 B aligned");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  launch_kernel(adamw_kernel, grid_1d(n), TPB, 0, st, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale, reinterpret_cast<const long long*>(step_dev));
  return cuda_check_launch("adamw_kernel");
}

extern "C" int gdmcf_counter_add(uint64_t* counter_dev, uint64_t inc, gdmcf_stream_t stream) {
  if (!counter_dev) { set_error("counter_add: null pointer"); return GDMCF_EBADARG; }
  GD_PRE();
  launch_kernel(counter_add_kernel, 1, 32, 0, st, reinterpret_cast<unsigned long long*>(counter_dev), inc);
  return cuda_check_launch("counter_add_kernel");
}

extern "C" int gdmcf_graph_noise_step(uint8_t* state, int64_t ld, const float* deg_frac, int t, int batch, float discrete,
                                      int user_guided, uint64_t seed, uint64_t offset, const uint64_t* epoch_dev,
                                      const float* u_entry, const float* u_user, int rows, int cols, gdmcf_stream_t stream) {
  if (!state || rows <= 0 || cols <= 0 || ld < cols || batch <= 0 || (user_guided && !deg_frac)) {
    set_error("graph_noise_step: bad arguments");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(graph_noise_step_kernel, grid_1d((long long)rows * ((cols + 3) / 4)), TPB, 0, st, state, (long long)ld, deg_frac, t,
                batch, discrete, user_guided, seed, offset, epoch_dev, u_entry, u_user, rows, cols);
  return cuda_check_launch("graph_noise_step_kernel");
}

extern "C" int gdmcf_scale_cols_cast(const float* in, int64_t ld_in, const float* col_scale, void* out_hi, void* out_lo,
                                     int64_t ld_out, int rows, int cols, gdmcf_stream_t stream) {
  if (!in || !col_scale || !out_hi || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < cols || (ld_out & 7) ||
      ((uintptr_t)out_hi & 15) || ((uintptr_t)out_lo & 15)) {
    set_error("scale_cols_cast: bad arguments (ld_out %% 8 == 0, 16 B aligned outputs)");
    return GDMCF_EBADARG;
  }
  GD_PRE();
  launch_kernel(scale_cols_cast_kernel, grid_1d((long long)rows * (ld_out / 8)), TPB, 0, st, in, (long long)ld_in, col_scale,
                (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, (long long)ld_out, rows, cols);
  return cuda_check_launch("scale_cols_cast_kernel");
}

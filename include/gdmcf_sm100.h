/* gdmcf_sm100.h — C ABI of libgdmcf_sm100.so: the B200 (sm_100a) kernels under GDMCF's
 * train-and-rank hot path.
 *
 * The reference (GDMCF/GDMCF) has no FFI: the path sits behind plain Python calls that end in ATen
 * library kernels. Each entry point below replaces one of those ATen call sites (cited as
 * reference file:line); the Python mirror of the reference classes (gdmcf_b200/models/*.py) is the
 * only caller in this repo, INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - raw device pointers + explicit sizes / leading dimensions (in ELEMENTS); no torch types;
 *  - the caller allocates every buffer, including workspaces (see the *_workspace_bytes queries);
 *  - all device work is asynchronous on the given stream; no hidden syncs or allocations;
 *  - return 0 on success, GDMCF_EBADARG (-1) bad shape/alignment, GDMCF_EARCH (-2) device is not
 *    sm_100, GDMCF_ECUDA (-3) CUDA error; gdmcf_last_error() returns a thread-local message;
 *  - timesteps are int32 on the device (the reference uses int64 tensors, converted at the boundary);
 *  - bf16 matrices are row-major with the reduction (K) dimension contiguous ("K-major").
 */
#ifndef GDMCF_SM100_H_
#define GDMCF_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDMCF_OK 0
#define GDMCF_EBADARG (-1)
#define GDMCF_EARCH (-2)
#define GDMCF_ECUDA (-3)

#define GDMCF_ABI_VERSION 1

typedef void* gdmcf_stream_t; /* cudaStream_t */

const char* gdmcf_last_error(void);
int gdmcf_abi_version(void);
/* 0 if the current device is compute capability 10.x, GDMCF_EARCH otherwise. */
int gdmcf_device_check(void);
int gdmcf_num_sms(void);

/* ------------------------------------------------------------------------------------------------
 * K1 — normalized-adjacency propagation (CSR SpMM).
 * Replaces torch.sparse.mm(A_tilda, E) at lightGCN.py:185 (+ stack/mean at :188-189) and the
 * GCNConv.propagate scatter at models/DNN.py:1095,1100.
 * ---------------------------------------------------------------------------------------------- */

/* Host-side plan: splits rows into work items of at most `chunk` non-zeros so that no warp walks a
 * hub row alone. items_out: int32[4*cap] = {row, begin, end, slot}; slot = -1 for a whole row that
 * the warp stores directly, otherwise an index into the partial-sum scratch. long_out: int32[3*cap_long]
 * = {row, first_slot, n_slots}. Pass NULL outputs to query the counts. */
int gdmcf_spmm_plan(const int32_t* rowptr_host, int n_rows, int chunk, int32_t* items_out, int cap_items,
                    int32_t* long_out, int cap_long, int* n_items, int* n_long, int* n_slots);

/* Y[r,:] = alpha * sum_j val[j] * X[col[j],:] + beta * Z[r,:]   (Z may be NULL -> beta ignored).
 * X,Y,Z: fp32 [n_rows or n_cols, d] row-major, d a multiple of 64; scratch: fp32 [n_slots, d]. */
int gdmcf_spmm_csr_f32(const int32_t* col, const float* val, const int32_t* items, int n_items,
                       const int32_t* long_rows, int n_long, const float* X, const float* Z, float* Y,
                       float* scratch, int n_rows, int d, float alpha, float beta, gdmcf_stream_t stream);

/* LightGCN.propagate_through_layers (lightGCN.py:180-194): out = mean_{k=0..K} A^k E0, evaluated as
 * the Horner recurrence T <- A T + E0 (K launches, layer mean fused into the store).
 * tmp0,tmp1: fp32 [n, d] ping-pong buffers (unused when K == 1). */
int gdmcf_lightgcn_propagate_f32(const int32_t* col, const float* val, const int32_t* items, int n_items,
                                 const int32_t* long_rows, int n_long, const float* E0, float* tmp0,
                                 float* tmp1, float* out, float* scratch, int n, int d, int n_layers,
                                 gdmcf_stream_t stream);

/* get_A_tilda (lightGCN.py:145-178) on the device: given the user->item CSR of R (U x I) and its
 * transpose, writes the CSR of A~ = D^-1/2 [[0,R],[R^T,0]] D^-1/2 with d_inv = (rowsum + 1e-9)^-1/2.
 * rowptr_out int32[U+I+1], col_out int32[2 nnz], val_out fp32[2 nnz]. */
int gdmcf_build_norm_adj(const int32_t* r_rowptr, const int32_t* r_col, const int32_t* rt_rowptr,
                         const int32_t* rt_col, int n_users, int n_items, int32_t* rowptr_out,
                         int32_t* col_out, float* val_out, gdmcf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K3-K7 — denoiser contractions: C[M,N] = sum_s A_s[M,K_s] * B_s[N,K_s]^T, bf16 operands, fp32
 * accumulation in tensor memory (tcgen05.mma fed by TMA), fused epilogue.
 * Replaces nn.Linear/torch.mm at models/DNN.py:79-86, :1240-1252, :1082-1100 (user rows), :1320-1325.
 * Up to three K segments let one launch consume a concatenated input ([h, h_U, e_user]) or a
 * hi/lo bf16 split of fp32 operands ("fp32 mode": a_hi*b_hi + a_hi*b_lo + a_lo*b_hi).
 * ---------------------------------------------------------------------------------------------- */
#define GDMCF_MAX_SEG 3

typedef struct {
  const void* a[GDMCF_MAX_SEG]; /* bf16 [m, k[s]], leading dim lda[s] (multiple of 8), 16B aligned */
  const void* b[GDMCF_MAX_SEG]; /* bf16 [n, k[s]], leading dim ldb[s] (multiple of 8), 16B aligned */
  int64_t lda[GDMCF_MAX_SEG];
  int64_t ldb[GDMCF_MAX_SEG];
  int32_t k[GDMCF_MAX_SEG];
  int32_t n_seg;
  int32_t m, n;
} gdmcf_gemm_desc;

#define GDMCF_EPI_STORE 0    /* out = alpha*acc                                         */
#define GDMCF_EPI_BIAS_ACT 1 /* out = act(alpha*acc + bias[t(m)*ld_bias + n])            */
#define GDMCF_EPI_COSINE 2   /* s = alpha*acc*row_scale[m]*col_scale[n];
                                out = c1 ? c1[t(m)]*s + c2[t(m)]*xt[m,n] : s
                                (cosine scorer DNN.py:1304-1327 + posterior mean
                                 gaussian_diffusion.py:1041-1050)                        */
#define GDMCF_ACT_NONE 0
#define GDMCF_ACT_TANH 1
#define GDMCF_ACT_RELU 2

typedef struct {
  int32_t mode, act;
  float alpha;
  int32_t t_const;        /* timestep used when row_t == NULL                           */
  float* out_f32;         /* optional fp32 [m, n], leading dim ld_f32 (multiple of 4)   */
  void* out_bf16;         /* optional bf16 [m, n], leading dim ld_bf16 (multiple of 8)  */
  void* out_bf16_lo;      /* optional bf16 residual (v - bf16(v)), same ld as out_bf16  */
  int64_t ld_f32, ld_bf16;
  const float* bias;      /* BIAS_ACT: vector [n] (ld_bias = 0) or table [T, ld_bias]   */
  int64_t ld_bias;
  const int32_t* row_t;   /* optional per-row timestep [m]                              */
  const float* row_scale; /* COSINE: [m]                                                */
  const float* col_scale; /* COSINE: [n]                                                */
  const float* c1;        /* COSINE: optional posterior_mean_coef1 [T]                  */
  const float* c2;        /* COSINE: posterior_mean_coef2 [T]                           */
  const float* xt;        /* COSINE: fp32 x_t [m, n], leading dim ld_xt                 */
  int64_t ld_xt;
} gdmcf_epilogue;

/* Split count that fills the SMs for this shape (1 = no split-K). */
int gdmcf_gemm_auto_splits(int m, int n, int k_total);
/* Bytes of fp32 workspace needed for `splits` > 1 (0 for splits == 1). */
size_t gdmcf_gemm_workspace_bytes(int m, int n, int splits);
int gdmcf_gemm_bf16_tn(const gdmcf_gemm_desc* g, const gdmcf_epilogue* e, int splits, void* workspace,
                       size_t workspace_bytes, gdmcf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise / data-layout kernels around the contractions.
 * ---------------------------------------------------------------------------------------------- */

/* fp32 [rows, cols] (ld_in) -> bf16 hi (and optional lo residual) [rows, cols] (ld_out), zero padding
 * columns cols..ld_out. Weight re-layout after each optimizer step. */
int gdmcf_cast_bf16(const float* in, int64_t ld_in, void* out_hi, void* out_lo, int64_t ld_out, int rows,
                    int cols, gdmcf_stream_t stream);
/* Same, writing the transpose: out[cols, rows] (ld_out >= rows). */
int gdmcf_cast_bf16_transpose(const float* in, int64_t ld_in, void* out_hi, void* out_lo, int64_t ld_out,
                              int rows, int cols, gdmcf_stream_t stream);

/* Dense interaction rows from CSR (replaces the dense n_user x n_item host matrix of main.py:143-156):
 * for r < n_rows: row = CSR row users[r] scattered as 1.0 into out_f32[r,:] (optional) and
 * out_bf16[r,:] (optional); both are zero-filled first over their full leading dimension. */
int gdmcf_densify_rows(const int32_t* rowptr, const int32_t* col, const int32_t* users, int n_rows,
                       int n_items, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                       gdmcf_stream_t stream);

/* q_sample (gaussian_diffusion.py:988-996) fused with the input dropout of the denoiser
 * (models/DNN.py:78, :1232): x_t = sqrt_ab[t_r]*x0 + sqrt_1mab[t_r]*eps, eps ~ N(0,1) from Philox
 * (seed, offset) or read from `noise` when non-NULL (test injection); xt_f32 (optional) receives x_t;
 * a_bf16 (+ optional a_lo) receives dropout_p-dropped x_t scaled by 1/(1-p) (p = 0 -> plain cast);
 * keep-mask drawn from Philox or read from `keep` (uint8) when non-NULL. */
int gdmcf_qsample_dropout(const float* x0, int64_t ld_x0, const int32_t* row_t, int t_const,
                          const float* sqrt_ab, const float* sqrt_1mab, const float* noise,
                          const uint8_t* keep, float dropout_p, uint64_t seed, uint64_t offset,
                          float* xt_f32, int64_t ld_xt, void* a_bf16, void* a_lo, int64_t ld_a, int rows,
                          int cols, gdmcf_stream_t stream);

/* apply_noise + "& one_hot(x0)" + dropout on the one-hot branch (gaussian_diffusion.py:770-831,:851;
 * models/DNN.py:1224,1233): for entry (r,i) with class c = x0[r,i] in {0,1} the kept channel c survives
 * with probability a + (1-a)*(c ? 1-p : p), a = ts[r]/rows, p = `discrete`; each of the two interleaved
 * channels then passes dropout. Output bf16 [rows, 2*cols] (ld_out), values in {0, 1/(1-dropout_p)}.
 * u_keep / u_drop: optional injected uniforms [rows, cols] / [rows, 2*cols] (fp32 in [0,1)). */
int gdmcf_onehot_noise(const float* x0, int64_t ld_x0, const int32_t* ts, float discrete, float dropout_p,
                       const float* u_keep, const float* u_drop, uint64_t seed, uint64_t offset,
                       void* out_bf16, int64_t ld_out, int rows, int cols, gdmcf_stream_t stream);

/* Inference form of the one-hot encoder (models/DNN.py:1249-1251 with x_tU = one_hot(x0)):
 * S[r,:] = base[:] + sum_{i in row users[r]} delta[i,:], where base = sum_i W2[:,2i] and
 * delta[i,:] = W2[:,2i+1] - W2[:,2i] are fp32 tables prepared from in_layers2.0.weight. */
int gdmcf_encode_onehot_gather(const int32_t* rowptr, const int32_t* col, const int32_t* users, int n_rows,
                               const float* base, const float* delta, int64_t ld_delta, int d, float* out,
                               int64_t ld_out, gdmcf_stream_t stream);
/* Builds base/delta from W2 fp32 [d, ld_w] (columns 2i, 2i+1 interleaved; models/DNN.py:1224). */
int gdmcf_onehot_tables(const float* w2, int64_t ld_w, int d, int n_items, float* base, float* delta,
                        int64_t ld_delta, gdmcf_stream_t stream);

/* Row-wise finish of the user tower (models/DNN.py:1288, :1320-1321):
 * hc'[r,:] = sumW*hc[r,:] + (1-sumW)*g[r,:]; inv_norm[r] = 1/||hc'[r,:]||_2; hc' written as bf16
 * hi (+ optional lo) for the scorer GEMM and optionally as fp32. g may be NULL (sumW treated as 1). */
int gdmcf_mix_rownorm(const float* hc, int64_t ld_hc, const float* g, int64_t ld_g, const float* sumw,
                      float* out_f32, int64_t ld_of, void* out_bf16, void* out_lo, int64_t ld_ob,
                      float* inv_norm, int rows, int cols, gdmcf_stream_t stream);
/* inv_norm[r] = 1/||x[r,:]||_2 for an fp32 matrix (item table norms, models/DNN.py:1321). */
int gdmcf_row_inv_norm(const float* x, int64_t ld, float* inv_norm, int rows, int cols,
                       gdmcf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K13 — history-masked top-K (main.py:299-301) and the ranking metrics (evaluate_utils.py:6-52).
 * ---------------------------------------------------------------------------------------------- */

/* For r < n_rows: scores[r, hist row users[r]] are treated as -inf (scores is NOT modified), the k
 * largest remaining entries are written to out_idx[r, 0..k) (int32, descending score, ties by
 * ascending index) and out_val (optional). hist_rowptr2/hist_col2: optional second history CSR
 * (train + valid masking for the test split, main.py:177). k <= 1024. */
int gdmcf_mask_topk(const float* scores, int64_t ld, int n_rows, int n_items, const int32_t* users,
                    const int32_t* hist_rowptr, const int32_t* hist_col, const int32_t* hist_rowptr2,
                    const int32_t* hist_col2, int k, int32_t* out_idx, float* out_val,
                    gdmcf_stream_t stream);

/* Per-user partial sums of computeTopNAccuracy: for each user r and each cutoff topn[j]:
 * stats[r, j, 0..4) = {hits/topn, hits/len(gt), dcg/idcg, 1/rank_first_hit} as fp64, zeros when the
 * user's ground truth is empty (such users still count in the denominator, evaluate_utils.py:17,47).
 * gt CSR rows are indexed by users[r] and must be sorted ascending. */
int gdmcf_topn_metrics(const int32_t* topk_idx, int ld_idx, int n_rows, const int32_t* users,
                       const int32_t* gt_rowptr, const int32_t* gt_col, const int32_t* topn, int n_topn,
                       double* stats, gdmcf_stream_t stream);
/* Deterministic column sums of an fp64 [rows, cols] matrix (fixed-order tree) -> out[cols]. */
int gdmcf_colsum_f64(const double* x, int rows, int cols, double* out, gdmcf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training-side kernels.
 * ---------------------------------------------------------------------------------------------- */

/* K10 — loss epilogue (gaussian_diffusion.py:902-932): mse[r] = mean_i (x0 - out)^2;
 * grad[r,i] = gscale[r] * 2*(out - x0)/cols written as fp32 (optional) and bf16 (optional, + its
 * transpose grad_t_bf16 [cols, rows] optional) for the backward contractions. */
int gdmcf_mse_rows(const float* out, int64_t ld_out, const float* x0, int64_t ld_x0, int rows, int cols,
                   float* mse, gdmcf_stream_t stream);

/* K14 — fused AdamW over one flat fp32 parameter (torch.optim.AdamW semantics, main.py:258,351):
 * p *= 1 - lr*wd; m,v updated; p -= step_size * m_hat / (sqrt(v_hat) + eps). grad_scale multiplies
 * the gradient first (1/world_size after an allreduce-sum). Optionally refreshes the bf16 operand
 * copy (hi + lo) used by the contractions when the parameter is a [rows, cols] matrix. */
int gdmcf_adamw_fused(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int step, float grad_scale,
                      gdmcf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GDMCF_SM100_H_ */

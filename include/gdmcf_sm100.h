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
/* Number of kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
unsigned long long gdmcf_launch_count(void);

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
/* The same propagation for the separable normalisation A~ = D^-1/2 A D^-1/2 of a BINARY adjacency (lightGCN.py:145-178):
 * only the CSR pattern and dinv[n] = (deg + 1e-9)^-1/2 are needed. Iterates on U_k = D^-1/2 T_k, so neighbour sums are
 * unweighted (no value stream, no per-non-zero multiply). u0, tmp0, tmp1: caller-provided [n + 1, d] buffers whose last row
 * is zero (the kernels never write it; padding lanes of the gather read it). Same result as the value form up to fp32
 * rounding. */
int gdmcf_lightgcn_propagate_sym_f32(const int32_t* col, const float* dinv, const int32_t* items, int n_items,
                                     const int32_t* long_rows, int n_long, const float* E0, float* u0, float* tmp0,
                                     float* tmp1, float* out, float* scratch, int n, int d, int n_layers,
                                     gdmcf_stream_t stream);
/* dinv[r] = (deg_r + 1e-9)^-1/2 over the bipartite graph [[0, R], [R^T, 0]] (lightGCN.py:160-166). */
int gdmcf_norm_adj_dinv(const int32_t* r_rowptr, const int32_t* rt_rowptr, int n_users, int n_items, float* dinv,
                        gdmcf_stream_t stream);
int gdmcf_build_norm_adj(const int32_t* r_rowptr, const int32_t* r_col, const int32_t* rt_rowptr,
                         const int32_t* rt_col, int n_users, int n_items, int32_t* rowptr_out,
                         int32_t* col_out, float* val_out, gdmcf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K3-K7 — denoiser contractions: C[M,N] = sum_s A_s[M,K_s] * B_s[N,K_s]^T, bf16 operands, fp32
 * accumulation in tensor memory (tcgen05.mma fed by TMA), fused epilogue.
 * Replaces nn.Linear/torch.mm at models/DNN.py:79-86, :1240-1252, :1082-1100 (user rows), :1320-1325.
 * Up to GDMCF_MAX_SEG K segments let one launch consume a concatenated input ([h, h_U, e_user]), a
 * hi/lo bf16 split of fp32 operands ("fp32 mode": a_hi*b_hi + a_hi*b_lo + a_lo*b_hi), or the per-rank
 * factors of a data-parallel weight gradient (sum over ranks of Gs_r^T hc'_r as ONE contraction whose K
 * runs over the ranks: the factor exchange of engine.StepEngine, 8 ranks = 8 segments).
 * ---------------------------------------------------------------------------------------------- */
#define GDMCF_MAX_SEG 8

typedef struct {
  const void* a[GDMCF_MAX_SEG]; /* bf16 [m, k[s]], leading dim lda[s] (multiple of 8), 16B aligned */
  const void* b[GDMCF_MAX_SEG]; /* bf16 [n, k[s]], leading dim ldb[s] (multiple of 8), 16B aligned */
  int64_t lda[GDMCF_MAX_SEG];
  int64_t ldb[GDMCF_MAX_SEG];
  int32_t k[GDMCF_MAX_SEG];
  int32_t n_seg;
  int32_t m, n;
} gdmcf_gemm_desc;

/* Fused epilogue, one formula (absent pointers drop their factor/term):
 *   s   = act(alpha * acc * row_scale[m] * col_scale[n] + bias[t(m)*ld_bias + n])
 *   out = c1 ? c1[t(m)] * s + c2[t(m)] * xt[m,n] : s            t(m) = row_t ? row_t[m] : t_const
 * - bias + tanh/relu: nn.Linear + activation (models/DNN.py:79-86, :1240-1252, GCNConv bias :1082-1100);
 * - row_scale/col_scale: the cosine scorer's 1/(|u| |e_i|) (models/DNN.py:1304-1327);
 * - c1/c2/xt: posterior mean coef1*pred_xstart + coef2*x_t (models/gaussian_diffusion.py:1041-1050).
 * Padding rule: when an output's leading dimension is a multiple of 4 (fp32) / 8 (bf16) elements, the columns
 * [n, round_up(n, 4 | 8)) of each row may be overwritten (bulk tensor stores clip at 16 B granularity); columns
 * beyond that are never touched. Outputs with other leading dimensions are written exactly.
 * `mode` is informational (kept for ABI stability): */
#define GDMCF_EPI_STORE 0
#define GDMCF_EPI_BIAS_ACT 1
#define GDMCF_EPI_COSINE 2
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
/* Upper bound on the SMs the persistent contraction kernels occupy from now on (0 = all SMs): lets a data-parallel
 * caller leave SMs to NCCL while all-reduces overlap with contractions. Process-wide, not thread-safe. */
int gdmcf_gemm_set_sm_limit(int sms);
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
                          const uint64_t* epoch_dev, float* xt_f32, int64_t ld_xt, void* a_bf16, void* a_lo, int64_t ld_a, int rows,
                          int cols, gdmcf_stream_t stream);

/* apply_noise + "& one_hot(x0)" + dropout on the one-hot branch (gaussian_diffusion.py:770-831,:851;
 * models/DNN.py:1224,1233): for entry (r,i) with class c = x0[r,i] in {0,1} the kept channel c survives
 * with probability a + (1-a)*(c ? 1-p : p), a = ts[r]/rows, p = `discrete`; each of the two interleaved
 * channels then passes dropout. Output bf16 [rows, 2*cols] (ld_out), values in {0, 1/(1-dropout_p)}.
 * u_keep / u_drop: optional injected uniforms [rows, cols] / [rows, 2*cols] (fp32 in [0,1)). */
int gdmcf_onehot_noise(const float* x0, int64_t ld_x0, const int32_t* ts, float discrete, float dropout_p,
                       const float* u_keep, const float* u_drop, uint64_t seed, uint64_t offset,
                       const uint64_t* epoch_dev, void* out_bf16, int64_t ld_out, int rows, int cols, gdmcf_stream_t stream);

/* LightGCN propagation in bf16 mode (lightGCN.py:180-194): out = mean_{k<=K} (A~^k E0) for the binary pattern
 * `col_hot_first` with D^-1/2 = dinv, all K layers in one persistent launch. The iterated tables are bf16 rows (fp32
 * accumulation; E0 / out stay fp32), the gdmcf_lightgcn_hot_rows() most frequently gathered rows are staged in shared
 * memory every layer. Plan (host): items / long_rows from gdmcf_spmm_plan with items[.].row replaced by -(long index + 1)
 * for hub-row pieces; every row's neighbour list reordered hot-first with a hot neighbour encoded as 0x80000000 | slot;
 * item_mids[i] = end of item i's hot prefix; the first n_pieces items are the hub-row pieces (dealt round-robin, one
 * per warp); the whole-row items that warp slot w (= CTA * 32 + warp) walks, two at a time, are
 * items[n_pieces + warp_ptr[w] .. n_pieces + warp_ptr[w + 1]) — the plan balances estimated cost over the n_warp_slots
 * slots (a multiple of 32; the grid is n_warp_slots / 32 CTAs and must fit the device: cooperative launch).
 * hot_rows[slot] = row id. u0 / u1: bf16 [n + 1, 64] with row n zero; scratch: fp32 [n_slots, 64]; sync_block:
 * 33 + n_long + 32 + 512 zero-initialised uint32 (counters, left zeroed, + phase timestamps for diagnostics). d must
 * be 64. Deterministic. Normwise error vs fp32 ~1e-3. */
int gdmcf_lightgcn_hot_rows(void);
int gdmcf_lightgcn_propagate_bf16(const int32_t* col_hot_first, const int32_t* items, const int32_t* item_mids, int n_items, int n_pieces,
                                  const int32_t* warp_ptr, int n_warp_slots, const int32_t* long_rows,
                                  int n_long, const int32_t* hot_rows, int n_hot, const float* dinv, const float* E0,
                                  void* u0_bf16, void* u1_bf16, float* out, float* scratch, uint32_t* sync_block, int n, int d,
                                  int n_layers, gdmcf_stream_t stream);

/* out[r, c] = bf16(in[r, c] * col_scale[c]) as a K-major operand (hi[, lo] residual; columns [cols, ld_out) zeroed).
 * Builds W1 diag(1/||E_i||), the operand of the projected reverse loop: the next encoder pre-activation
 * W1 x_{t-1} = c1[t] * ru * hc' (E^T diag(ri) W1^T) + c2[t] * W1 x_t never needs the [B, n_item] scores of the
 * intermediate reverse steps (models/gaussian_diffusion.py:1047-1050 pushed through the linear first layer, DNN.py:1240). */
int gdmcf_scale_cols_cast(const float* in, int64_t ld_in, const float* col_scale, void* out_hi, void* out_lo,
                          int64_t ld_out, int rows, int cols, gdmcf_stream_t stream);

/* One reverse step of p_sample's random graph bookkeeping (models/gaussian_diffusion.py:710-729; "faithful graph" mode —
 * the edges only feed GCN item rows that the model never reads, so the default path skips it): state uint8 [rows, ld],
 * state |= guide_b & (class-0 entry flips to 1 w.p. (1 - t/batch)(1 - discrete)), guide_b ~ Bernoulli(deg_frac[b]) when
 * user_guided (deg_frac = row degree / max row degree of the batch), else 1. u_entry [rows, cols] / u_user [rows]: optional
 * injected uniforms replacing the Philox draws. */
int gdmcf_graph_noise_step(uint8_t* state, int64_t ld, const float* deg_frac, int t, int batch, float discrete,
                           int user_guided, uint64_t seed, uint64_t offset, const uint64_t* epoch_dev, const float* u_entry,
                           const float* u_user, int rows, int cols, gdmcf_stream_t stream);

/* Inference form of the one-hot encoder (models/DNN.py:1249-1251 with x_tU = one_hot(x0)):
 * S[r,:] = base[:] + sum_{i in row users[r]} delta[i,:], where base = sum_i W2[:,2i] and
 * delta[i,:] = W2[:,2i+1] - W2[:,2i] are fp32 tables prepared from in_layers2.0.weight. */
/* workspace (gdmcf_encode_onehot_gather_workspace_bytes) + counters (int32 [n_rows], zero-initialised, left zeroed):
 * optional; with them the rows of heavy users are gathered by several CTAs (deterministic slice-order reduction). */
size_t gdmcf_encode_onehot_gather_workspace_bytes(int n_rows, int d);
int gdmcf_encode_onehot_gather(const int32_t* rowptr, const int32_t* col, const int32_t* users, int n_rows,
                               const float* base, const float* delta, int64_t ld_delta, int d, float* out,
                               int64_t ld_out, float* workspace, size_t workspace_bytes, int32_t* counters,
                               gdmcf_stream_t stream);
/* Builds base/delta from W2 fp32 [d, ld_w] (columns 2i, 2i+1 interleaved; models/DNN.py:1224). */
int gdmcf_onehot_tables(const float* w2, int64_t ld_w, int d, int n_items, float* base, float* delta,
                        int64_t ld_delta, gdmcf_stream_t stream);

/* Per-timestep first-layer bias (the `cat([x, emb])` columns, models/DNN.py:72-80, :1227-1241): for every
 * t < T: emb_table[t,:] = emb_layer(timestep_embedding(t, e)) and out[t,k] = b[k] + sum_j w_time[k,j]*emb_table[t,j],
 * where w_time = &W[0, n_in] (the last e columns of the layer's weight, leading dim ld_w). b may be NULL. */
int gdmcf_time_bias_table(const float* w_emb, const float* b_emb, const float* w_time, int64_t ld_w,
                          const float* b, int T, int e, int d, float* emb_table, float* out, int64_t ld_out,
                          gdmcf_stream_t stream);
/* out[r,c] = act(in[r,c] + bias[t(r)*ld_bias + c]) as fp32 and/or bf16 hi(+lo). */
int gdmcf_bias_act_rows(const float* in, int64_t ld_in, const float* bias, int64_t ld_bias, const int32_t* row_t,
                        int t_const, int act, float* out_f32, int64_t ld_of, void* out_bf16, void* out_lo,
                        int64_t ld_ob, int rows, int cols, gdmcf_stream_t stream);
/* out[r,:] = table[idx[r],:] (nn.Embedding lookup, models/DNN.py:1265) as fp32 and/or bf16 hi(+lo). */
int gdmcf_gather_rows(const float* table, int64_t ld_t, const int32_t* idx, float* out_f32, int64_t ld_of,
                      void* out_bf16, void* out_lo, int64_t ld_ob, int rows, int cols, gdmcf_stream_t stream);
/* Small fp32 CUDA-core GEMM for the tiny contractions (time-embedding columns, nt_xent logits):
 * C = alpha * op(A) op(B) + beta * C; trans_a: A stored [k,m]; trans_b: B stored [n,k]. */
int gdmcf_sgemm_small(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                      int64_t ldc, int m, int n, int k, float alpha, float beta, gdmcf_stream_t stream);
/* out[c] = sum_r x[r,c] in row order (bias gradients). */
int gdmcf_colsum_f32(const float* x, int64_t ld, int rows, int cols, float* out, gdmcf_stream_t stream);

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
                      float beta2, float eps, float weight_decay, int step, const int64_t* step_dev,
                      float grad_scale, gdmcf_stream_t stream);
/* AdamW on a [rows, cols] fp32 weight fused with the refresh of the tensors the contractions derive from it (all
 * optional, NULL = skip): bf16 operand W[:, :cols_used] (hi[, lo], zero padded to ld_hi), its transpose (t_hi[, t_lo],
 * [cols_used, ld_t]), row inverse norms 1/||W[r,:]|| (inv_norm; models/DNN.py:1320), the one-hot encoder tables
 * delta[i,:] = W[:,2i+1] - W[:,2i] / base = sum_i W[:,2i] (models/DNN.py:1249-1251 at x_tU = one_hot(x0)), and a
 * contiguous copy of the trailing columns W[:, cols_used:] (tcols [rows, n_tcols]). rowpart: workspace of
 * gdmcf_adamw_refresh_splits(rows, cols) * rows floats (needed for inv_norm / delta). g may have a padded leading
 * dimension ld_g. row_coef (optional, [rows]): the effective gradient is g + row_coef[r] * W[r,:] — the norm term of the
 * cosine scorer's backward, d/dE of 1/||E_i||, is -E_i * c_i (models/DNN.py:1320-1325); deferring it to this pass saves
 * the wgrad contraction a full read of E. Same update arithmetic as gdmcf_adamw_fused.
 * g == NULL: refresh only — no update, m / v ignored; the derived tensors are recomputed from the current weights. */
typedef struct gdmcf_refresh {
  int32_t cols_used;      /* 0 = all columns */
  int32_t n_tcols;
  void* hi; void* lo; int64_t ld_hi;
  void* t_hi; void* t_lo; int64_t ld_t;
  float* inv_norm;
  float* delta; int64_t ld_delta; float* base;
  float* tcols;
  float* rowpart;
  const float* row_coef;
} gdmcf_refresh;
int gdmcf_adamw_refresh_splits(int rows, int cols);
int gdmcf_adamw_refresh(float* p, const float* g, int64_t ld_g, float* m, float* v, int rows, int cols, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step, const int64_t* step_dev,
                        float grad_scale, const gdmcf_refresh* out, gdmcf_stream_t stream);
/* User tower of DNNOneHotEmbeddingGCN in one launch (bf16 mode): LayerGCN on the user rows (models/DNN.py:1093-1100,
 * GCNConv with self loops only = two linears), the sumW mix (:1288) and the user norms of the cosine scorer (:1320):
 *   g1 = relu(hc W1^T + b1) [rows, hidden];  g2 = g1 W2^T + b2 [rows, n];  hc' = hc*sumW + g2*(1 - sumW);  inv_u = 1/||hc'||
 * hc_bf16 [rows, ld_hcb] / hc_f32 [rows, ld_hc]: the concatenated tower input (k1 = n columns); w1_bf16 [hidden, ld_w1],
 * w2_bf16 [n, ld_w2] K-major bf16 operands. Outputs: hcp_bf16 [rows, ld_hcp] (columns [n, ld_hcp) zeroed), inv_u [rows],
 * optional fp32 copies g1_f32 [rows, hidden], g2_f32, hcp_f32 (training keeps them for the backward pass). workspace:
 * gdmcf_user_tower_workspace_bytes(); sync_block: 64 zero-initialised bytes that the kernel leaves zeroed (grid barriers of
 * one resident wave of CTAs; max_ctas > 0 bounds the wave, 0 = one CTA per SM). Deterministic. */
size_t gdmcf_user_tower_workspace_bytes(int rows, int k1, int hidden, int n);
int gdmcf_user_tower(const void* hc_bf16, int64_t ld_hcb, const float* hc_f32, int64_t ld_hc, const void* w1_bf16,
                     int64_t ld_w1, const float* b1, const void* w2_bf16, int64_t ld_w2, const float* b2, const float* sumw,
                     int rows, int k1, int hidden, int n, void* hcp_bf16, int64_t ld_hcp, float* inv_u, float* g1_f32,
                     float* g2_f32, int64_t ld_g2, float* hcp_f32, int64_t ld_hcp32, void* workspace, size_t workspace_bytes,
                     uint32_t* sync_block, int max_ctas, gdmcf_stream_t stream);
/* AdamW only (no derived tensors) confined to `n_ctas` SMs: n_ctas CTAs (clusters of 2, one CTA per SM, each reserving
 * more than half of the SM's shared memory so that no tcgen05 contraction CTA can share it). Lets the HBM-bound optimizer
 * pass (main.py:351) run on a side stream next to the tensor-bound denoise + rank phase whose contractions were limited to
 * the other SMs with gdmcf_gemm_set_sm_limit. Same arithmetic and row_coef meaning as gdmcf_adamw_refresh. */
int gdmcf_adamw_partitioned(float* p, const float* g, int64_t ld_g, float* m, float* v, int rows, int cols, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int step, const int64_t* step_dev,
                            float grad_scale, const float* row_coef, int n_ctas, gdmcf_stream_t stream);
/* AdamW (torch.optim.AdamW semantics, main.py:258,351) on a [n_rows, cols] embedding table whose gradient is non-zero on
 * n_sel rows only: idx int32 [n_sel] (distinct rows), grad_rows [n_sel, ld_g]. last_step int32 [n_rows] (zero-initialised)
 * records the optimizer step each row is up to date with; a row that receives a gradient first replays the steps it
 * missed with a zero gradient (decaying moments keep moving the weights, exactly as the dense pass would), then takes the
 * real update — results are bit-identical to gdmcf_adamw_fused on the dense gradient. grad_rows == NULL: the rows idx (the
 * rows the next forward pass will read) or, with idx == NULL, all rows (flush: required before the table is read outside
 * the training step, saved or evaluated) are brought up to the current step with zero gradients. */
int gdmcf_adamw_rows_lazy(float* p, float* m, float* v, int32_t* last_step, const int32_t* idx, const float* grad_rows,
                          int64_t ld_g, int n_sel, int n_rows, int cols, float lr, float beta1, float beta2, float eps,
                          float weight_decay, int step, const int64_t* step_dev, float grad_scale, gdmcf_stream_t stream);
/* counter_dev[0] += inc on the stream: the device-resident step / RNG-epoch counters that keep captured CUDA graphs
 * advancing (Philox counter high word = (epoch << 8) | sub-stream; AdamW bias corrections from *step_dev when non-NULL). */
int gdmcf_counter_add(uint64_t* counter_dev, uint64_t inc, gdmcf_stream_t stream);

/* Backward of the weighted MSE (models/gaussian_diffusion.py:902,932,951) in one pass over out [rows, cols]:
 *   g[b,i] = gs[b] * 2*(out[b,i] - x0[b,i]) / cols          (gs[b] = dL/dloss[b] * weight[b] / pt[b])
 *   g_bf16[b,i]  = bf16(g * row_scale[b] * col_scale[i])     A operand of the dgrad contraction
 *   gt_bf16[i,b] = its transpose (optional)                  A operand of the wgrad contraction
 *   g_lo / gt_lo = optional bf16 residuals of the two (fp32 mode)
 *   colsum[i]      = sum_b (with_out ? g*out : g)            (optional)
 *   rowpart[cb, b] = the same quantity summed over the 32 columns of block cb (optional; [ceil(cols/32), rows]);
 * with_out = 1 yields the sums the cosine scorer's norm terms need (models/DNN.py:1320-1325). */
int gdmcf_loss_grad(const float* out, int64_t ld_out, const float* x0, int64_t ld_x0, const float* gs,
                    const float* row_scale, const float* col_scale, int with_out, void* g_bf16, void* g_lo,
                    int64_t ld_g, void* gt_bf16, void* gt_lo, int64_t ld_gt, float* colsum, float* rowpart,
                    int rows, int cols, gdmcf_stream_t stream);
/* bf16 [rows, cols] -> bf16 [cols, rows] (wgrad operands need the batch dimension contiguous). */
int gdmcf_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int rows, int cols,
                         gdmcf_stream_t stream);
/* out = op(a, b): 0: a*(b>0) (relu backward)  1: a*(1-b*b) (tanh backward)  2: alpha*a + beta*b. */
int gdmcf_ew_binary(int op, const float* a, int64_t ld_a, const float* b, int64_t ld_b, float alpha, float beta,
                    float* out_f32, int64_t ld_of, void* out_bf16, void* out_lo, int64_t ld_ob, int rows, int cols,
                    gdmcf_stream_t stream);
/* Backward of hc' = sumW*hc + (1-sumW)*g2 (models/DNN.py:1288): d_hc = sumW*d, d_g2 = (1-sumW)*d,
 * dw_rows[r] = sum_c d[r,c]*(hc[r,c]-g2[r,c]). */
int gdmcf_mix_backward(const float* d_hcp, int64_t ld_d, const float* hc, int64_t ld_hc, const float* g2,
                       int64_t ld_g, const float* sumw, float* d_hc, int64_t ld_dh, float* d_g2, int64_t ld_dg,
                       float* dw_rows, int rows, int cols, gdmcf_stream_t stream);
/* nt_xent_loss rows (models/DNN.py:479-508) from the raw logits S = h h_U^T [n,n]: loss_rows[i] =
 * -log((p_ii+eps)/sum_{j!=i} p_ij) with p = softmax(S/tau) (closs = mean(loss_rows)); dS = dscale[0] * d closs/dS. */
int gdmcf_ntxent_rows(const float* S, int64_t ld_s, int n, float tau, float eps, const float* dscale,
                      float* loss_rows, float* dS, int64_t ld_ds, gdmcf_stream_t stream);
/* Lt_history / Lt_count update (models/gaussian_diffusion.py:935-949), sequential semantics per timestep:
 * ts int64 [batch], loss fp64 [batch], lt_history fp64 [steps, history], lt_count int64 [steps]. */
int gdmcf_lt_history_update(const int64_t* ts, const double* loss, double* lt_history, int64_t* lt_count, int batch,
                            int steps, int history, gdmcf_stream_t stream);
/* Per-row loss terms of training_losses (models/gaussian_diffusion.py:906-957; START_X, optional reweighting) in one
 * launch: weight = t == 0 ? 1 : SNR(t-1) - SNR(t) with SNR = ac / (1 - ac); hist_loss = weight * mse (feeds
 * gdmcf_lt_history_update); loss = hist_loss / pt + 0.1 * closs[0] (closs may be NULL); g_mse = float(weight / pt / batch),
 * the backward seed d mean(loss) / d mse. ts int64, pt / alphas_cumprod / hist_loss / loss fp64, mse / closs / g_mse fp32.
 * Bit-identical to the tensor-op evaluation (single correctly rounded operations in the same order). */
int gdmcf_loss_terms(const int64_t* ts, const double* pt, const float* mse, const double* alphas_cumprod,
                     const float* closs, int batch, int steps, int reweight, double* hist_loss, double* loss,
                     float* g_mse, gdmcf_stream_t stream);
/* sample_timesteps(method="importance") (models/gaussian_diffusion.py:959-986) on the device, no host sync:
 * uniform draws with pt = 1 until every lt_count[t] == history, then t ~ Categorical(p), p = sqrt(mean(Lt_history^2))
 * normalised and mixed with uniform_prob, pt = p[t] * steps. ts_in != NULL: only pt for the given timesteps.
 * ts int64 [batch], pt fp64 [batch]; Philox counter = (offset + row, (epoch_dev[0] << 8) | 7). */
int gdmcf_sample_timesteps(const double* lt_history, const int64_t* lt_count, int steps, int history, int batch,
                           double uniform_prob, uint64_t seed, uint64_t offset, const uint64_t* epoch_dev,
                           const int64_t* ts_in, int64_t* ts_out, double* pt_out, gdmcf_stream_t stream);
/* grad[idx[r],:] += v[r,:] (dense nn.Embedding gradient rows, models/DNN.py:1265). */
int gdmcf_scatter_rows_add(const float* v, int64_t ld_v, const int32_t* idx, float* grad, int64_t ld_g, int rows,
                           int cols, gdmcf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GDMCF_SM100_H_ */

/* basd_b200.h -- C ABI of the B200-native BASD loss hot path (libbasd_b200.so).
 *
 * The reference (indrajeetadityaroy9/vit-inductive-bias-distillation) has no FFI: its
 * boundary is the Python module API of src/losses (combined.py:17-85,
 * layer_selector.py:40-152, relational.py:5-50), whose arithmetic bottoms out in
 * torch.linalg / torch.mm library calls.  This header is what a binding for that path
 * binds instead: one entry point per device stage, each citing the reference lines it
 * replaces.  The Python shim (vit-inductive-bias-distillation_b200/_native.py) loads the
 * library with ctypes and calls these from a torch.autograd.Function.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory unless marked "host";
 *     the caller (torch) owns all buffers, nothing is allocated or freed inside;
 *   - matrices are row-major; `ld` is the row pitch in elements, `stride` the distance
 *     between consecutive problems of a batch in elements;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work, none
 *     synchronises the host;
 *   - return value: 0 on success, a cudaError_t (>0) if a launch failed, <0 for an
 *     argument the kernels do not support.  No exceptions cross the boundary;
 *   - dtype codes: BASD_DTYPE_F32 = 0, BASD_DTYPE_BF16 = 1.
 */
#ifndef BASD_B200_H
#define BASD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BASD_DTYPE_F32 0
#define BASD_DTYPE_BF16 1

/* ---- dense contractions (fp32 SIMT; gemm_simt.cu) ---------------------------------- */

/* C[b] = alpha * (*alpha_dev) * op(A[b]) * op(B[b]) + beta * C[b].  A may be bf16; when
 * a_col_shift != NULL, a_col_shift[k] is subtracted from op(A)[m][k] on load (trans_a=0).
 * Replaces torch.mm / torch.bmm at layer_selector.py:72,88,99 and relational.py:47 and the
 * MmBackward/BmmBackward nodes of their autograd graph. */
int basd_sgemm_batched(int trans_a, int trans_b, int M, int N, int K, const void* A, int a_dtype,
                       int lda, long stride_a, const float* a_col_shift, const float* B, int ldb,
                       long stride_b, float* C, int ldc, long stride_c, int batch, float alpha,
                       const float* alpha_dev, float beta, void* stream);

/* ---- mean-shifted token statistics ----------------------------------------------------- */

/* The pooled covariance behind the MP rank and the layer subspaces (layer_selector.py:13,35,72,88,91) is
 * accumulated in a frame shifted by a rough mean mu0 (covariance is shift invariant), so that the Gram
 * accumulators never hold the M mu mu^T term the centring would have to cancel afterwards (DESIGN.md
 * section 3.1).  mu0[i] = mean of rows_sample evenly spaced rows of tensor i (count tensors, one launch).  The
 * Gram entry points below take mu0 (bf16-representable values) and return the statistics of x - mu0. */
int basd_rough_means(const void* const* tensors, int count, int dtype, long rows, int D, long rows_sample,
                     float* mu0, void* stream);
/* Data-parallel merge of statistics the ranks accumulated in their OWN frames mu0_r: gram (tensors, D, D) holds
 * sum_r G'_r on entry (after the all-reduce), d_slots / mu_slots (world, tensors, D) every rank's column sums
 * and shift; on exit gram, colsum and mu0 (tensors, D) are the statistics of the common frame mean_r mu0_r.
 * One all-reduce carries everything (the slots of the other ranks are zero in the local buffer). */
int basd_merge_shifted_stats(float* gram, float* colsum, float* mu0, const float* d_slots, const float* mu_slots,
                             int world, int tensors, int D, long rows_per_rank, void* stream);

/* Token-space second-moment statistics of one (rows x D) token matrix:
 * gram = X^T X (D x D, symmetric), colsum = X^T 1 (D).  Deterministic split-K.
 * Replaces `features.T @ features` (layer_selector.py:13) and the column mean (:35,:91);
 * the fixed projections proj_s/proj_t are applied to the small matrix afterwards. */
long basd_token_gram_simt_workspace_floats(long rows, int D);
int basd_token_gram_simt(const void* tokens, int dtype, long rows, int D, const float* mu0, float* gram,
                         float* colsum, float* workspace, void* stream);

/* Same statistic on the tensor cores (gram_tc.cu): TMA-fed tcgen05.mma (bf16 x bf16 -> fp32 in
 * TMEM), 128x128 upper-triangular tiles, split-K over the token rows. bf16 tokens, D % 128 == 0. */
long basd_token_gram_tc_workspace_bytes(long rows, int D);
int basd_token_gram_tc(const void* tokens, long rows, int D, const float* mu0, float* gram, float* colsum,
                       void* workspace, void* stream);

/* ---- small-matrix factorisations (jacobi.cu) ---------------------------------------- */

/* Rank-revealing (diagonally pivoted) Cholesky of `batch` PSD matrices.  K is destroyed.
 * LT row j holds the j-th column of L (K ~= LT^T LT); rows >= rank are zero.
 * dims (optional, device) = active leading dimension per problem. */
int basd_pivoted_cholesky(float* K, int n, int ld, long stride_k, float* LT, int ldl,
                          long stride_l, int batch, float rel_tol, int* rank_out, const int* dims,
                          void* stream);

/* One-sided Jacobi: orthogonalises the rows of each (n x m) matrix in place (rows end up
 * as sigma_j * v_j^T).  Shared-memory resident when the matrix fits, warp-shuffle dot
 * products, in-register rotations.  Serves torch.linalg.eigvalsh (layer_selector.py:16),
 * torch.linalg.svd (:36,:92), svdvals (:99) and matrix_norm(ord="nuc") (relational.py:48).
 * Requires ld % 4 == 0, stride % 4 == 0, 16-byte aligned base, m <= 1024. */
int basd_jacobi_rows(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                     float tol, int max_sweeps, int* sweeps_out, void* stream);

/* General form.  dims (square active size) and row_dims (leading non-zero rows) are mutually exclusive,
 * either may be null.  Two thresholds on the cosine between rows: a pair is rotated while |cos| > tol; the
 * sweeps stop after one in which every rotated pair had |cos| < stop_cos.  The entry points above and below
 * use stop_cos = sqrt(tol) (quadratic convergence of the sweep): enough for singular values and polar
 * factors.  Eigenvectors of a dense spectrum need stop_cos << relative eigenvalue gap (rows with nearly
 * equal norms keep mixing by cos / gap): the selector's symmetric eigenproblems pass 1e-5.
 * (torch.linalg.svd at layer_selector.py:36,92) */
int basd_jacobi_rows_ex(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                        const int* row_dims, float tol, float stop_cos, int max_sweeps, int* sweeps_out,
                        int* rot_out, void* stream);

/* Rank-deficient products G = F_q^T F_p: only the first row_dims[problem] rows are non-zero (the
 * pivoted-Cholesky rank); sweeps over those rows only.  row_dims may be null (= basd_jacobi_rows). */
int basd_jacobi_rows_ranked(float* G, int n, int m, int ld, long stride, int batch,
                            const int* row_dims, float tol, int max_sweeps, int* sweeps_out,
                            int* rot_out, void* stream);

/* Same, and rot_out[problem] += number of plane rotations applied (for the roofline accounting). */
int basd_jacobi_rows_counted(float* G, int n, int m, int ld, long stride, int batch,
                             const int* dims, float tol, int max_sweeps, int* sweeps_out,
                             int* rot_out, void* stream);

/* Row norms -> vals (norm or norm^2), unit rows -> V, optional descending sort. */
int basd_rows_normalize(const float* G, int n, int m, int ld, long stride, float* V, int ldv,
                        long stride_v, float* vals, int batch, int sort, int square,
                        float rel_floor, const int* dims, void* stream);

/* out[b*n + r] = <A[b][r,:], B[b][r,:]>  (Rayleigh-quotient refinement of eigenvalues). */
int basd_rowdot(const float* A, int lda, long stride_a, const float* B, int ldb, long stride_b,
                int n, int m, int batch, float* out, void* stream);

/* ---- selector glue (selector.cu) ---------------------------------------------------- */

/* K = sym(G) - inv_rows * colsum colsum^T (colsum may be NULL). layer_selector.py:35,91. */
int basd_center_gram(const float* G, const float* colsum, int D, float inv_rows, float* K,
                     int batch, void* stream);

/* Projection of the token-space statistics, computed in double precision (rotate_f64.cu):
 *   k_centred = fp32( sym(P G P^T) - inv_rows (P c)(P c)^T ),  chat = fp32(P c)
 * for a batch of (gram (d_in x d_in), colsum (d_in)) pairs sharing the fixed projection
 * proj (d_out x d_in).  Replaces `student_tokens @ proj_s.T` / `teacher_tokens @ proj_t.T`
 * (layer_selector.py:72,88) followed by the centring of :35,:91, in Gram form; fp64 because the
 * rotation of a spectrum spanning cond(X)^2 in fp32 is what limits the selector's gradient.
 * workspace: basd_rotate_stats_f64_workspace_bytes(...) bytes, 8-byte aligned. */
long basd_rotate_stats_f64_workspace_bytes(int d_out, int d_in, int batch);
int basd_rotate_stats_f64(const float* proj, int d_out, int d_in, const float* gram,
                          const float* colsum, int batch, double inv_rows, void* workspace,
                          float* k_centred, float* chat, void* stream);

/* Device-side marchenko_pastur_rank (layer_selector.py:8-20) + the cap of :74.
 * lam: (layers, D) spectrum of the uncentred second moment, any order.
 * edges (optional): (layers, 2) = median, lambda_plus. No host sync. */
int basd_mp_rank(const float* lam, int D, long rows, int cap, int* ranks, float* edges, int layers,
                 void* stream);

/* Same rank from the eigendecomposition of the CENTRED second moment: the uncentred one is a
 * rank-one update (+ rho c c^T, y = V^T c), whose eigenvalue counting function
 * #{mu > x} = #{lam > x} + [1 + rho sum y_j^2/(lam_j - x) < 0] gives the lower median by bisection
 * and the rank by one evaluation.  edges (optional): (layers, 3) = median, lambda_plus, tie flag. */
int basd_mp_rank_secular(const float* lam_c, const float* y, int D, long rows, int cap, int* ranks,
                         float* edges, int layers, void* stream);

int basd_expand_ranks(const int* ranks, int E, int L, int* dims, void* stream);
int basd_mask_block(const float* src, float* dst, int D, const int* dims, int batch, void* stream);

/* d[i,l] = sum_m sw_m acos(min(sig_m,1-eps))^2 / sum_m sw_m   (layer_selector.py:100-105) */
int basd_angle_distance(const float* sig, const float* lam_c, const int* ranks, int D, int E, int L,
                        float* dist, void* stream);

/* weights = softmax(-d / softplus(log_temp))                  (layer_selector.py:67,107-108) */
int basd_mix_weights(const float* dist, const float* log_temp, int E, int L, float* weights,
                     float* temps, void* stream);
int basd_mix_weights_bwd(const float* d_weights, const float* weights, const float* dist,
                         const float* log_temp, int E, int L, float scale, float* d_dist,
                         float* d_log_temp, void* stream);
int basd_scale_rows_dsigma(float* Uxt, const float* sig, const float* lam_c, const int* ranks,
                           const float* d_dist, int D, int E, int L, void* stream);
int basd_omega_accumulate(const float* block, const float* lam_s, const int* ranks, int D, int E,
                          int L, float* omega, void* stream);
int basd_symmetrize_add(const float* in, int D, float* out, int batch, void* stream);

/* K <- K + rel_shift * max(diag K) * I per problem: same eigenvectors, and the pivoted Cholesky behind the
 * symmetric eigenproblems (torch.linalg.svd / eigvalsh at layer_selector.py:16,36,92) never meets a
 * noise-sized pivot on a nearly singular Gram. */
int basd_shift_diag(float* K, int D, float rel_shift, int batch, void* stream);

/* Null-space completion for numerically rank-deficient student Grams (fewer token rows than dimensions, the
 * regime of layer_selector.py:14-15, or eigenvalues below the fp32 floor): P = I - sym(vtv) with
 * vtv = V_r^T V_r, dims_out[problem] = D if rows are missing else 0 (the `dims` of the Cholesky that follows,
 * so complete bases cost nothing); after basd_pivoted_cholesky(P) rows 0..rank_P-1 of its factor are an
 * orthonormal basis of the complement and are written into the zero rows D-rank_P..D-1 of Vt -- what the
 * (I - V V^T) term of the thin-SVD backward (layer_selector.py:92) needs. */
int basd_projector_complement(const float* vtv, int D, float* P, int* dims_out, int batch, void* stream);
int basd_place_complement(float* Vt, const float* LT, const int* rank_p, int D, int batch, void* stream);

/* ---- HBM-bound mixing (mix.cu) ------------------------------------------------------ */

/* attention map (B,H,q_rows,side) -> importance row (B, n_tok): CLS row mean over heads, or
 * mean over (heads, queries) without CLS.  q_rows = side for a full map, 1 when the caller
 * hands over only the CLS query row.                           (relational.py:22-27) */
int basd_attn_rows(const void* attn, int dtype, int B, int H, int side, int q_rows, int has_cls,
                   float* rows, void* stream);

/* out[i] = resample_{n_src->n_dst}( sum_l weights[i,l] * teacher_l ), all i in one pass.
 * teacher_layers: HOST array of L device pointers, each (B, n_src, D).
 *                                                (layer_selector.py:110-111, combined.py:9-14) */
int basd_mix_interp(const void* const* teacher_layers, int L, int E, const float* weights,
                    int in_dtype, int B, int n_src, int n_dst, int D, void* out, int out_dtype,
                    void* stream);

/* Flat layer mix out[i][x] = sum_l weights[i,l] * layer_l[x] over numel elements (any tensor shape), all E
 * outputs in one pass: the reference-shaped outputs of GrassmannianLayerSelector.forward -- mixed full
 * attention maps (layer_selector.py:112).  out: (E, numel) in the layers' dtype; layer pointers 16-byte
 * aligned. */
int basd_mix_flat(const void* const* layers, int L, int E, const float* weights, int dtype, long numel,
                  void* out, void* stream);

/* mixed + resampled + normalised token importance.   (layer_selector.py:112, relational.py:29-34) */
int basd_mix_rows(const float* rows, const float* weights, int E, int L, int B, int n_src,
                  int n_dst, float* w_out, float* totals, void* stream);

/* dL/dweights[i,l] = <Z_i, resample(teacher_l)> + gw_scale * <gw_i, resample(rows_l)>.
 * partial: scratch of basd_weight_grad_slices() * L * E floats.  n_src = tokens per teacher layer,
 * n_rows = length of the importance rows (B, n_rows): the attention map's own token count, which
 * differs from n_src when the tokens were resampled by the caller (relational.py:29-32). */
int basd_weight_grad_slices(void);
int basd_weight_grad(const void* const* teacher_layers, int L, int E, const float* Z,
                     const float* gw, const float* rows, int in_dtype, int B, int n_src, int n_rows,
                     int n_dst, int D, float gw_scale, const float* gw_scale_dev, float* partial,
                     float* d_weights, void* stream);

/* ---- tensor-core batched GEMM (gemm_tc3.cu) ----------------------------------------- */

/* C[b] (M x N, pitch ldc) = alpha * alpha_dev[0] * op(A[b]) op(B[b]) on tcgen05 tensor cores
 * with fp32-level accuracy: every operand is split into two TF32 terms and three
 * kind::tf32 MMAs per K step accumulate in TMEM ("3xTF32").  ta = 0: A stored M x K,
 * ta = 1: stored K x M;  tb = 0: B stored K x N, tb = 1: stored N x K.  Pointers 16-byte
 * aligned; M, N, K, pitches and strides multiples of 4 (basd_gemm_tc3_supported tells).
 * Operands are fetched by TMA (tensor maps over (inner, outer, batch), encoded per call on the
 * host): a stride of 0 shares an operand across the batch.  Returns 0, -3 (alignment / shape
 * rule), -4 (batch > 65535) or -5 (the driver refused the tensor map).
 * Replaces torch.bmm (relational.py:47) and the matmuls of linalg.svd's backward for the
 * per-sample Procrustes products. */
int basd_gemm_tc3_supported(int M, int N, int K, int lda, int ldb, int ldc, long sa, long sb,
                            long sc);
int basd_gemm_tc3_batched(int ta, int tb, int M, int N, int K, const float* A, int lda, long sa,
                          const float* B, int ldb, long sb, float* C, int ldc, long sc, int batch,
                          float alpha, const float* alpha_dev, void* stream);
/* Extended form: C = alpha * alpha_dev[0] * (op(A) op(B) - 1 col_sub^T).  a_dtype BF16 (ta = 0;
 * K, lda, sa multiples of 8): bf16 tokens feed the tensor cores exactly (bf16 is a subset of TF32,
 * two MMAs per K step); c_dtype BF16 stores the result in the token dtype; col_sub (N floats,
 * nullable) = mu^T B turns the product into (A - 1 mu^T) B -- the token centring of
 * layer_selector.py:90-91 applied in the epilogue. */
int basd_gemm_tc3_batched_ex(int ta, int tb, int M, int N, int K, const void* A, int a_dtype, int lda,
                             long sa, const float* B, int ldb, long sb, void* C, int c_dtype, int ldc,
                             long sc, int batch, float alpha, const float* alpha_dev,
                             const float* col_sub, void* stream);

/* ---- Procrustes glue (procrustes.cu) ------------------------------------------------ */

/* out = sqrt(w) * (X - sum_n w_n X_n) per sample.                 (relational.py:36-43) */
int basd_weighted_center(const void* X, int dtype, long stride_x, const float* W, long stride_w,
                         int N, int D, float* out, long stride_o, int batch, void* stream);
int basd_extract_diag(const float* K, int N, int ld, long stride, int batch, float* diag,
                      void* stream);
/* rows2 (rq x rq): row j = sigma_j q_j^T, Pt (rq x rp): unit rows p_j^T.  sig = row norms of
 * rows2, nuc = their sum, keep_j = sig_j > rel_floor * max; rows2_j <- keep q_j^T sig^(eq/2),
 * Pt_j <- keep p_j^T sig^(ep/2), pic_j = keep sig^(-(eq+ep)/2); eq, ep in {-1,0,1}.
 * The derived q_j rows additionally need sig_j > rel_floor_q * max (>= rel_floor).
 * (the singular values / vectors torch.linalg.matrix_norm(ord="nuc") differentiates through,
 * relational.py:48) */
int basd_procrustes_rows_finish(float* rows2, int rq, int ldr, long stride_r, float* Pt, int rp,
                                int ldp, long stride_p, int batch, float rel_floor, float rel_floor_q,
                                int eq, int ep, float* sig, float* nuc, float* pic, void* stream);
/* f = tr_s + tr_t - 2 nuc, gw = df/dw~ from the images IA, IB (rq x N); a non-null Y (N x N,
 * Gram side) becomes 2 diag(sqrt w)(I - Y) in place.                   (relational.py:34-50) */
int basd_procrustes_grad_prep(float* YA, float* YB, const float* IA, const float* IB, int N, int rq,
                              int ldi, long stride_i, int ldy, long stride_y, int batch,
                              const float* pic, const float* nuc, const float* ks, const float* kt,
                              const float* w, const float* totals, float* f_out, float* gw,
                              int with_grad, void* stream);
/* direct side (D <= N): T (N x D) <- 2 sqrt(w_n) (A - T). */
int basd_procrustes_direct_grad(const float* A, float* T, const float* w, int N, int D, int batch,
                                void* stream);
/* dst (dtype) = alpha * alpha_dev[0] * src (fp32). */
int basd_scale_out(const float* src, void* dst, int dtype, long n, float alpha,
                   const float* alpha_dev, void* stream);
/* geo_terms[i] = mean_b f[i,b], geo = mean_i geo_terms[i]  (relational.py:50, combined.py:76);
 * NaN when any of the n_weights mixing weights is not finite (layer_selector.py:105, k = 0). */
int basd_geo_reduce(const float* f, int E, int B, const float* weights, int n_weights,
                    float* geo_terms, float* geo, void* stream);
int basd_cast_out(const float* src, void* dst, int dtype, long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BASD_B200_H */

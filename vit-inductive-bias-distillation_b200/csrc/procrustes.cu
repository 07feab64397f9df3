// Attention-weighted Procrustes loss, per-sample glue kernels
// (reference: relational.py:34-50 and its autograd).  The dense work between these kernels
// (N x N Grams, factor products) runs in gemm_*.cu, the factorisations in jacobi.cu.
//
// Per sample:  A = sqrt(w)(S - mu_s), Bm = sqrt(w)(R' - mu_t)
//   K_s = A A^T = L_s L_s^T, K_t = Bm Bm^T = L_t L_t^T, X = L_s^T L_t = U S V^T
//   f = tr K_s + tr K_t - 2 sum(S)
//   df/dS  = 2 sqrt(w) (I - Y_A) A,   Y_A = (L_t V) S^+ (L_t V)^T
//   df/dR' = 2 sqrt(w) (I - Y_B) Bm,  Y_B = (L_s U) S^+ (L_s U)^T
//   df/dw_n = (K_s[n,n] + K_t[n,n] - 2 [L_s U (L_t V)^T]_{nn}) / w_n
#include "common.cuh"

namespace basd {

// out[n,:] = sqrt(w_n) (x[n,:] - sum_m w_m x[m,:])           (relational.py:36-43)
template <typename T>
__global__ void __launch_bounds__(256)
weighted_center_kernel(const T* __restrict__ X, long strideX, const float* __restrict__ W,
                       long strideW, int N, int D, float* __restrict__ out, long strideO) {
  extern __shared__ float sm[];
  float* w = sm;           // N
  float* rw = sm + N;      // N  sqrt(w)
  const int s = blockIdx.x;
  const T* x = X + (long)s * strideX;
  float* o = out + (long)s * strideO;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float v = W[(long)s * strideW + n];
    w[n] = v;
    rw[n] = sqrtf(v);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float mu = 0.f;
    for (int n = 0; n < N; ++n) mu = fmaf(w[n], to_f32<T>(x[(long)n * D + d]), mu);
    for (int n = 0; n < N; ++n)
      o[(long)n * D + d] = rw[n] * (to_f32<T>(x[(long)n * D + d]) - mu);
  }
}

// diag[s*N + n] = K[s][n][n]
__global__ void extract_diag_kernel(const float* __restrict__ K, int N, int ld, long stride,
                                    float* __restrict__ diag) {
  const int s = blockIdx.x;
  for (int n = threadIdx.x; n < N; n += blockDim.x)
    diag[(long)s * N + n] = K[(long)s * stride + (long)n * ld + n];
}

// rows2 (N x N) = U^T X, row j = sigma_j v_j^T.  Produces per sample:
//   sig[j] = |row j| (refined singular values), nuc = sum_j sig[j]
//   Vt'[j,:] = keep_j * row_j / sig_j^{3/2}      (i.e. v_j^T / sqrt(sig_j))
//   Ut'[j,:] = keep_j * Ut[j,:] / sqrt(sig_j)
// keep_j = sig_j > rel_floor * max(sig).
__global__ void __launch_bounds__(512)
procrustes_rows_finish_kernel(float* __restrict__ rows2, float* __restrict__ Ut, int N, int ld,
                              long stride, float rel_floor, float* __restrict__ sig,
                              float* __restrict__ nuc) {
  extern __shared__ float sm[];
  float* nrm = sm;        // N
  float* red = sm + N;    // 32
  const int s = blockIdx.x;
  float* R = rows2 + (long)s * stride;
  float* U = Ut + (long)s * stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < N; r += nw) {
    float a = 0.f;
    for (int c = lane; c < N; c += 32) { const float v = R[(long)r * ld + c]; a = fmaf(v, v, a); }
    a = warp_sum(a);
    if (lane == 0) nrm[r] = sqrtf(a);
  }
  __syncthreads();
  float mx = 0.f, tot = 0.f;
  for (int r = threadIdx.x; r < N; r += blockDim.x) { mx = fmaxf(mx, nrm[r]); tot += nrm[r]; }
  mx = block_max(mx, red);
  tot = block_sum(tot, red);
  const float floor_v = rel_floor * mx;
  for (int r = warp; r < N; r += nw) {
    const float sg = nrm[r];
    const bool keep = sg > floor_v && sg > 0.f;
    const float isq = keep ? rsqrtf(sg) : 0.f;
    const float iv = keep ? isq / sg : 0.f;
    for (int c = lane; c < N; c += 32) {
      R[(long)r * ld + c] *= iv;
      U[(long)r * ld + c] *= isq;
    }
    if (lane == 0) sig[(long)s * N + r] = sg;
  }
  if (threadIdx.x == 0) nuc[s] = tot;
}

// Per sample, after Y_A = FAt'^T FAt', Y_B = FBt'^T FBt' (N x N each):
//   f = tr_s + tr_t - 2 nuc
//   pi[n] = sum_j FBt'[j,n] FAt'[j,n] sig[j]
//   gw[n] = ((ks[n] + kt[n] - 2 pi[n]) / w[n] - f) / total        (d f / d w~)
//   M_A = 2 diag(sqrt w) (I - Y_A)   (in place over Y_A), same for M_B.
__global__ void __launch_bounds__(512)
procrustes_grad_prep_kernel(float* __restrict__ YA, float* __restrict__ YB,
                            const float* __restrict__ FAt, const float* __restrict__ FBt, int N,
                            int ld, long stride, const float* __restrict__ sig,
                            const float* __restrict__ nuc, const float* __restrict__ ks,
                            const float* __restrict__ kt, const float* __restrict__ w,
                            const float* __restrict__ totals, float* __restrict__ f_out,
                            float* __restrict__ gw, int with_grad) {
  __shared__ float red[32];
  __shared__ float f_sh;
  const int s = blockIdx.x;
  float a = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) a += ks[(long)s * N + n] + kt[(long)s * N + n];
  a = block_sum(a, red);
  if (threadIdx.x == 0) { f_sh = a - 2.f * nuc[s]; f_out[s] = f_sh; }
  __syncthreads();
  if (!with_grad) return;
  const float f = f_sh;
  const float* fa = FAt + (long)s * stride;
  const float* fb = FBt + (long)s * stride;
  const float tot = totals[s];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float pi = 0.f;
    for (int j = 0; j < N; ++j)
      pi = fmaf(fb[(long)j * ld + n] * fa[(long)j * ld + n], sig[(long)s * N + j], pi);
    const float wn = w[(long)s * N + n];
    gw[(long)s * N + n] = ((ks[(long)s * N + n] + kt[(long)s * N + n] - 2.f * pi) / wn - f) / tot;
  }
  float* ya = YA + (long)s * stride;
  float* yb = YB + (long)s * stride;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int r = e / N, c = e - r * N;
    const float sc = 2.f * sqrtf(w[(long)s * N + r]);
    const float id = (r == c) ? 1.f : 0.f;
    ya[(long)r * ld + c] = sc * (id - ya[(long)r * ld + c]);
    yb[(long)r * ld + c] = sc * (id - yb[(long)r * ld + c]);
  }
}

// geo_terms[i] = mean_b f[i,b]; geo = mean_i geo_terms[i]    (relational.py:50, combined.py:76)
__global__ void geo_reduce_kernel(const float* __restrict__ f, int E, int B,
                                  float* __restrict__ geo_terms, float* __restrict__ geo) {
  __shared__ float red[32];
  float total = 0.f;
  for (int i = 0; i < E; ++i) {
    float a = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) a += f[(long)i * B + b];
    a = block_sum(a, red) / (float)B;
    if (threadIdx.x == 0) geo_terms[i] = a;
    total += a;
    __syncthreads();
  }
  if (threadIdx.x == 0) *geo = total / (float)E;
}

// dst (token dtype) = src (fp32), elementwise.
template <typename T>
__global__ void cast_out_kernel(const float* __restrict__ src, T* __restrict__ dst, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = from_f32<T>(src[i]);
}

}  // namespace basd

using namespace basd;
#define ST ((cudaStream_t)stream)

extern "C" int basd_weighted_center(const void* X, int dtype, long stride_x, const float* W,
                                    long stride_w, int N, int D, float* out, long stride_o,
                                    int batch, void* stream) {
  if (batch <= 0) return 0;
  const size_t dyn = (size_t)2 * N * sizeof(float);
  if (dtype == BASD_DTYPE_BF16)
    weighted_center_kernel<__nv_bfloat16><<<batch, 256, dyn, ST>>>(
        (const __nv_bfloat16*)X, stride_x, W, stride_w, N, D, out, stride_o);
  else
    weighted_center_kernel<float><<<batch, 256, dyn, ST>>>((const float*)X, stride_x, W, stride_w,
                                                           N, D, out, stride_o);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_extract_diag(const float* K, int N, int ld, long stride, int batch, float* diag,
                                 void* stream) {
  if (batch <= 0) return 0;
  extract_diag_kernel<<<batch, 256, 0, ST>>>(K, N, ld, stride, diag);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_procrustes_rows_finish(float* rows2, float* Ut, int N, int ld, long stride,
                                           int batch, float rel_floor, float* sig, float* nuc,
                                           void* stream) {
  if (batch <= 0) return 0;
  procrustes_rows_finish_kernel<<<batch, 512, (N + 32) * sizeof(float), ST>>>(
      rows2, Ut, N, ld, stride, rel_floor, sig, nuc);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_procrustes_grad_prep(float* YA, float* YB, const float* FAt, const float* FBt,
                                         int N, int ld, long stride, int batch, const float* sig,
                                         const float* nuc, const float* ks, const float* kt,
                                         const float* w, const float* totals, float* f_out,
                                         float* gw, int with_grad, void* stream) {
  if (batch <= 0) return 0;
  procrustes_grad_prep_kernel<<<batch, 512, 0, ST>>>(YA, YB, FAt, FBt, N, ld, stride, sig, nuc, ks,
                                                     kt, w, totals, f_out, gw, with_grad);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_geo_reduce(const float* f, int E, int B, float* geo_terms, float* geo,
                               void* stream) {
  geo_reduce_kernel<<<1, 256, 0, ST>>>(f, E, B, geo_terms, geo);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_cast_out(const float* src, void* dst, int dtype, long n, void* stream) {
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == BASD_DTYPE_BF16)
    cast_out_kernel<__nv_bfloat16><<<grid, 256, 0, ST>>>(src, (__nv_bfloat16*)dst, n);
  else
    cast_out_kernel<float><<<grid, 256, 0, ST>>>(src, (float*)dst, n);
  BASD_LAUNCH_CHECK();
  return 0;
}

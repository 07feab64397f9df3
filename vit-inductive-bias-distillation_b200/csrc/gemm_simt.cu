// fp32 SIMT GEMMs: the exact-arithmetic workhorse for the small rotations, the per-sample
// N x N products and the fp32-token configurations (C1).  The bf16 token Grams at C2+
// run on tcgen05 (gram_tc.cu); both produce the same statistic.
//
//   basd_sgemm_batched : C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b]      (row-major)
//                        A may be fp32 or bf16 and may have a per-column shift subtracted
//                        on load (used for the centred-token gradient GEMM,
//                        reference: autograd of layer_selector.py:88-92).
//   basd_token_gram    : G = X^T X (token space, symmetric, split-K, deterministic) and
//                        c = X^T 1  (reference: layer_selector.py:13,35 after R4)
#include "common.cuh"

namespace basd {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB, typename AT>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const AT* __restrict__ A, int lda, long sA,
             const float* __restrict__ a_shift, const float* __restrict__ B, int ldb, long sB,
             float* __restrict__ C, int ldc, long sC, float alpha,
             const float* __restrict__ alpha_dev, float beta) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int batch = blockIdx.z;
  A += (long)batch * sA;
  B += (long)batch * sB;
  C += (long)batch * sC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    if (!TA) {  // A[m][k], k contiguous
      const int m = tid >> 2, kk = (tid & 3) * 4;
      const bool row_ok = (m0 + m) < M;
      const AT* src = A + (long)(m0 + m) * lda + k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        float v = 0.f;
        if (row_ok && k < K) {
          v = to_f32<AT>(src[i]);
          if (a_shift) v -= a_shift[k];
        }
        As[kk + i][m] = v;
      }
    } else {  // A stored [k][m], m contiguous
      const int kk = tid >> 4, mm = (tid & 15) * 4;
      const bool k_ok = (k0 + kk) < K;
      const AT* src = A + (long)(k0 + kk) * lda + m0 + mm;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (k_ok && (m0 + mm + i) < M) v = to_f32<AT>(src[i]);
        As[kk][mm + i] = v;
      }
    }
    if (!TB) {  // B[k][n], n contiguous
      const int kk = tid >> 4, nn = (tid & 15) * 4;
      const bool k_ok = (k0 + kk) < K;
      const float* src = B + (long)(k0 + kk) * ldb + n0 + nn;
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[kk][nn + i] = (k_ok && (n0 + nn + i) < N) ? src[i] : 0.f;
    } else {  // B stored [n][k], k contiguous
      const int n = tid >> 2, kk = (tid & 3) * 4;
      const bool col_ok = (n0 + n) < N;
      const float* src = B + (long)(n0 + n) * ldb + k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[kk + i][n] = (col_ok && (k0 + kk + i) < K) ? src[i] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float scale = alpha;
  if (alpha_dev) scale *= *alpha_dev;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* dst = C + (long)m * ldc + n;
      const float prev = (beta != 0.f) ? beta * (*dst) : 0.f;
      *dst = scale * acc[i][j] + prev;
    }
  }
}

// ---------------------------------------------------------------- token-space Gram
// Upper-triangular 64x64 tiles of X^T X over one slice of the rows; partial sums go to
// `partial[slice]` (full D x D layout) and are folded by gram_reduce_kernel in a fixed
// order, so the statistic is bit-reproducible from run to run.
template <typename T>
__global__ void __launch_bounds__(256)
token_gram_partial_kernel(const T* __restrict__ X, long rows, int D, long rows_per_slice,
                          float* __restrict__ partial) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tiles = (D + BM - 1) / BM;
  // decode upper-triangular tile pair
  int t = blockIdx.x, ti = 0;
  while (t >= tiles - ti) { t -= tiles - ti; ++ti; }
  const int tj = ti + t;
  const int a0 = ti * BM, b0 = tj * BN;
  const long r_begin = (long)blockIdx.y * rows_per_slice;
  const long r_end = min(rows, r_begin + rows_per_slice);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int kk = tid >> 4, dd = (tid & 15) * 4;
  for (long r0 = r_begin; r0 < r_end; r0 += BK) {
    const bool k_ok = (r0 + kk) < r_end;
    const T* row = X + (r0 + kk) * (long)D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[kk][dd + i] = (k_ok && (a0 + dd + i) < D) ? to_f32<T>(row[a0 + dd + i]) : 0.f;
      Bs[kk][dd + i] = (k_ok && (b0 + dd + i) < D) ? to_f32<T>(row[b0 + dd + i]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + (long)blockIdx.y * D * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + tx * 4 + j;
      if (b < D) out[(long)a * D + b] = acc[i][j];
    }
  }
}

// Folds the slices (fixed order) and mirrors the upper tiles into the lower triangle.
// `tile` is the granularity at which the producer skipped the lower triangle.
__global__ void gram_reduce_kernel(const float* __restrict__ partial, int slices, int D, int tile,
                                   float* __restrict__ gram, float beta) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  int a = idx / D, b = idx % D;
  const int ta = a / tile, tb = b / tile;
  if (ta > tb) { const int s = a; a = b; b = s; }
  float sum = 0.f;
  for (int s = 0; s < slices; ++s) sum += partial[(long)s * D * D + (long)a * D + b];
  gram[idx] = sum + (beta != 0.f ? beta * gram[idx] : 0.f);
}

// Column sums: one block per 32-column group, rows strided over blockIdx.y, then a
// deterministic second pass.
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ X, long rows, int D,
                                      long rows_per_slice, float* __restrict__ partial) {
  const int d = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;  // 8 row lanes
  const long r_begin = (long)blockIdx.y * rows_per_slice;
  const long r_end = min(rows, r_begin + rows_per_slice);
  float s = 0.f;
  if (d < D)
    for (long r = r_begin + sub; r < r_end; r += 8) s += to_f32<T>(X[r * D + d]);
  __shared__ float red[8][33];
  red[sub][threadIdx.x & 31] = s;
  __syncthreads();
  if (sub == 0 && d < D) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x & 31];
    partial[(long)blockIdx.y * D + d] = tot;
  }
}
__global__ void colsum_reduce_kernel(const float* __restrict__ partial, int slices, int D,
                                     float* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float s = 0.f;
  for (int i = 0; i < slices; ++i) s += partial[(long)i * D + d];
  out[d] = s;
}

template <bool TA, bool TB>
static int launch_sgemm(int a_dtype, int M, int N, int K, const void* A, int lda, long sA,
                        const float* a_shift, const float* B, int ldb, long sB, float* C, int ldc,
                        long sC, int batch, float alpha, const float* alpha_dev, float beta,
                        cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
  if (a_dtype == BASD_DTYPE_BF16)
    sgemm_kernel<TA, TB, __nv_bfloat16><<<grid, 256, 0, st>>>(
        M, N, K, (const __nv_bfloat16*)A, lda, sA, a_shift, B, ldb, sB, C, ldc, sC, alpha,
        alpha_dev, beta);
  else
    sgemm_kernel<TA, TB, float><<<grid, 256, 0, st>>>(M, N, K, (const float*)A, lda, sA, a_shift,
                                                      B, ldb, sB, C, ldc, sC, alpha, alpha_dev,
                                                      beta);
  BASD_LAUNCH_CHECK();
  return 0;
}

// Host-side helpers shared with gram_tc.cu (declared in common.cuh).
int launch_gram_reduce(const float* partial, int slices, int D, int tile, float* gram,
                       cudaStream_t st) {
  const long total = (long)D * D;
  gram_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, slices, D, tile, gram,
                                                                      0.f);
  BASD_LAUNCH_CHECK();
  return 0;
}

int launch_colsum_bf16(const void* tokens, long rows, int D, float* partial, float* out,
                       cudaStream_t st) {
  int cs = (int)((rows + 4095) / 4096);
  if (cs > 64) cs = 64;
  if (cs < 1) cs = 1;
  const long per = (rows + cs - 1) / cs;
  dim3 cgrid((D + 31) / 32, cs);
  colsum_partial_kernel<__nv_bfloat16><<<cgrid, 256, 0, st>>>((const __nv_bfloat16*)tokens, rows, D,
                                                              per, partial);
  colsum_reduce_kernel<<<(D + 127) / 128, 128, 0, st>>>(partial, cs, D, out);
  BASD_LAUNCH_CHECK();
  return 0;
}

}  // namespace basd

extern "C" int basd_sgemm_batched(int trans_a, int trans_b, int M, int N, int K, const void* A,
                                  int a_dtype, int lda, long stride_a, const float* a_col_shift,
                                  const float* B, int ldb, long stride_b, float* C, int ldc,
                                  long stride_c, int batch, float alpha, const float* alpha_dev,
                                  float beta, void* stream) {
  using namespace basd;
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  if (batch > 65535) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  if (!trans_a && !trans_b)
    return launch_sgemm<false, false>(a_dtype, M, N, K, A, lda, stride_a, a_col_shift, B, ldb,
                                      stride_b, C, ldc, stride_c, batch, alpha, alpha_dev, beta, st);
  if (!trans_a && trans_b)
    return launch_sgemm<false, true>(a_dtype, M, N, K, A, lda, stride_a, a_col_shift, B, ldb,
                                     stride_b, C, ldc, stride_c, batch, alpha, alpha_dev, beta, st);
  if (trans_a && !trans_b)
    return launch_sgemm<true, false>(a_dtype, M, N, K, A, lda, stride_a, a_col_shift, B, ldb,
                                     stride_b, C, ldc, stride_c, batch, alpha, alpha_dev, beta, st);
  return launch_sgemm<true, true>(a_dtype, M, N, K, A, lda, stride_a, a_col_shift, B, ldb,
                                  stride_b, C, ldc, stride_c, batch, alpha, alpha_dev, beta, st);
}

extern "C" long basd_token_gram_simt_workspace_floats(long rows, int D) {
  long slices = (rows + 2047) / 2048;
  if (slices > 64) slices = 64;
  if (slices < 1) slices = 1;
  return slices * ((long)D * D + D);
}

// gram[D*D] = X^T X, colsum[D] = X^T 1 for X = tokens viewed as (rows, D).
extern "C" int basd_token_gram_simt(const void* tokens, int dtype, long rows, int D, float* gram,
                                    float* colsum, float* workspace, void* stream) {
  using namespace basd;
  cudaStream_t st = (cudaStream_t)stream;
  long slices = (rows + 2047) / 2048;
  if (slices > 64) slices = 64;
  if (slices < 1) slices = 1;
  long per = (rows + slices - 1) / slices;
  per = (per + BK - 1) / BK * BK;
  const int tiles = (D + BM - 1) / BM;
  dim3 grid(tiles * (tiles + 1) / 2, (unsigned)slices);
  float* part_g = workspace;
  float* part_c = workspace + slices * (long)D * D;
  dim3 cgrid((D + 31) / 32, (unsigned)slices);
  if (dtype == BASD_DTYPE_BF16) {
    token_gram_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)tokens, rows, D, per, part_g);
    colsum_partial_kernel<__nv_bfloat16><<<cgrid, 256, 0, st>>>((const __nv_bfloat16*)tokens, rows,
                                                                D, per, part_c);
  } else {
    token_gram_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)tokens, rows, D, per,
                                                           part_g);
    colsum_partial_kernel<float><<<cgrid, 256, 0, st>>>((const float*)tokens, rows, D, per, part_c);
  }
  BASD_LAUNCH_CHECK();
  const long total = (long)D * D;
  gram_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part_g, (int)slices, D, BM,
                                                                      gram, 0.f);
  colsum_reduce_kernel<<<(D + 127) / 128, 128, 0, st>>>(part_c, (int)slices, D, colsum);
  BASD_LAUNCH_CHECK();
  return 0;
}

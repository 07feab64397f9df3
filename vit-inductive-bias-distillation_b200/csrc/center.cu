// Rough token means for the mean-shifted statistics (reference: layer_selector.py:13,35,72,88,91 -- the
// pooled Gram and column sums behind the Marchenko-Pastur rank and the layer subspaces).
//
// The selector needs the CENTRED covariance of M = B x N token rows.  Accumulating the uncentred Gram
// sum x x^T and subtracting M mu mu^T afterwards cancels catastrophically once |mu|^2 exceeds the
// eigenvalues of interest: the tensor-core accumulators carry ~3e-6 relative error over 50,176 rows
// (measured at C2, B = 256), i.e. 3e-6 |mu|^2 in absolute terms -- 3e-4 there, against an eigenvalue gap of
// 1.7e-4 at the rank boundary and a smallest eigenvalue of 1e-5.  The centred matrix then has negative
// eigenvalues, the pivoted Cholesky behind sym_eig divides by a noise-sized pivot, and the top-k student
// subspace comes out rotated by O(1) (selector-gradient cosine 0.87 against autograd through the reference,
// while every small-batch parity case passed).  ViT token means are large (massive activations), so this is
// not a corner of the synthetic data.
//
// Fix: the Gram kernels work in a frame shifted by a ROUGH mean mu0 (covariance is shift invariant) without
// touching the tokens: gram_tc.cu subtracts 64 mu0 mu0^T per 64-row stage with one extra negated UMMA,
// gemm_simt.cu subtracts mu0 while staging its fp32 operands; both return the Gram and the column sums of
// x - mu0.  This file only provides the shift:
//   basd_rough_means : mu0 = mean of the first rows_sample rows of every tensor of a group (one launch).
// The caller rounds mu0 to bf16 (the correction operand must be exact) and, with data-parallel ranks,
// all-reduces it first: the ranks' statistics only add if they share the shift.
#include "common.cuh"

namespace basd {

constexpr int CENTER_MAX_TENSORS = 64;
struct TensorPtrs { const void* in[CENTER_MAX_TENSORS]; };

template <typename T>
__global__ void __launch_bounds__(256)
rough_mean_kernel(TensorPtrs t, long rows_sample, int D, float* __restrict__ mu0) {
  const T* X = reinterpret_cast<const T*>(t.in[blockIdx.y]);
  const int d = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;
  float s = 0.f;
  if (d < D)
    for (long r = sub; r < rows_sample; r += 8) s += to_f32<T>(X[r * D + d]);
  __shared__ float red[8][33];
  red[sub][threadIdx.x & 31] = s;
  __syncthreads();
  if (sub == 0 && d < D) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x & 31];
    mu0[(long)blockIdx.y * D + d] = tot / (float)rows_sample;
  }
}

}  // namespace basd

using namespace basd;

static int fill(TensorPtrs& t, const void* const* in, int count) {
  if (count < 1 || count > CENTER_MAX_TENSORS) return -6;
  for (int i = 0; i < count; ++i) t.in[i] = in[i];
  return 0;
}

extern "C" int basd_rough_means(const void* const* tensors, int count, int dtype, long rows, int D,
                                long rows_sample, float* mu0, void* stream) {
  TensorPtrs t;
  if (int rc = fill(t, tensors, count)) return rc;
  if (rows_sample > rows) rows_sample = rows;
  if (rows_sample < 1) return -9;
  dim3 grid((D + 31) / 32, count);
  if (dtype == BASD_DTYPE_BF16)
    rough_mean_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(t, rows_sample, D, mu0);
  else
    rough_mean_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(t, rows_sample, D, mu0);
  BASD_LAUNCH_CHECK();
  return 0;
}


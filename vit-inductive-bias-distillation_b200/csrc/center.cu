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
//   basd_rough_means : mu0 = mean of rows_sample evenly spaced rows of every tensor of a group (one launch).
// The caller rounds mu0 to bf16 (the correction operand must be exact) and, with data-parallel ranks,
// all-reduces it first: the ranks' statistics only add if they share the shift.
#include "common.cuh"

namespace basd {

constexpr int CENTER_MAX_TENSORS = 64;
struct TensorPtrs { const void* in[CENTER_MAX_TENSORS]; };

template <typename T>
__global__ void __launch_bounds__(256)
rough_mean_kernel(TensorPtrs t, long rows_sample, long row_stride, int D, float* __restrict__ mu0) {
  const T* X = reinterpret_cast<const T*>(t.in[blockIdx.y]);
  const int d = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;
  float s = 0.f;
  if (d < D)
    for (long r = sub; r < rows_sample; r += 8) s += to_f32<T>(X[r * row_stride * D + d]);
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

// Data-parallel merge.  Rank r accumulated its statistics in its OWN frame mu0_r:
//   G'_r = sum_r (x - mu0_r)(x - mu0_r)^T,   d_r = sum_r (x - mu0_r).
// After ONE all-reduce that sums the G'_r and carries every rank's (d_r, mu0_r) in its own slot, the
// statistics of the common frame mu0 = mean_r mu0_r follow with delta_r = mu0_r - mu0:
//   G' = sum_r [ G'_r + d_r delta_r^T + delta_r d_r^T + M_r delta_r delta_r^T ],   d = sum_r (d_r + M_r delta_r).
// The delta_r are sampling-noise sized (rough means of the same distribution), so nothing cancels.
// grid (ceil(D D / 256), tensors); gram (tensors, D, D) holds sum_r G'_r on entry; d_slots / mu_slots are
// (world, tensors, D); colsum (tensors, D) and mu0 (tensors, D) are written.
__global__ void merge_shifted_stats_kernel(float* __restrict__ gram, float* __restrict__ colsum,
                                           float* __restrict__ mu0, const float* __restrict__ d_slots,
                                           const float* __restrict__ mu_slots, int world, int tensors, int D,
                                           float rows_per_rank) {
  const int t = blockIdx.y;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  const int a = idx / D, b = idx % D;
  const int lo = min(a, b), hi = max(a, b);              // evaluated on (min, max): bitwise symmetric
  float m_lo = 0.f, m_hi = 0.f;
  for (int r = 0; r < world; ++r) {
    const float* mu = mu_slots + ((long)r * tensors + t) * D;
    m_lo += mu[lo];
    m_hi += mu[hi];
  }
  m_lo /= (float)world;
  m_hi /= (float)world;
  float add = 0.f, dsum = 0.f;
  for (int r = 0; r < world; ++r) {
    const float* mu = mu_slots + ((long)r * tensors + t) * D;
    const float* d = d_slots + ((long)r * tensors + t) * D;
    const float dl = mu[lo] - m_lo, dh = mu[hi] - m_hi;
    add += fmaf(d[lo], dh, fmaf(dl, d[hi], rows_per_rank * dl * dh));
    if (a == 0) dsum += d[b] + rows_per_rank * dh;        // a == 0: hi == b
  }
  gram[(long)t * D * D + idx] += add;
  if (a == 0) {
    colsum[(long)t * D + b] = dsum;
    mu0[(long)t * D + b] = m_hi;
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
  const long row_stride = rows / rows_sample;            // every row_stride-th row: a sample of the whole batch
  dim3 grid((D + 31) / 32, count);
  if (dtype == BASD_DTYPE_BF16)
    rough_mean_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(t, rows_sample, row_stride, D, mu0);
  else
    rough_mean_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(t, rows_sample, row_stride, D, mu0);
  BASD_LAUNCH_CHECK();
  return 0;
}


extern "C" int basd_merge_shifted_stats(float* gram, float* colsum, float* mu0, const float* d_slots,
                                        const float* mu_slots, int world, int tensors, int D,
                                        long rows_per_rank, void* stream) {
  if (world < 1 || tensors < 1) return 0;
  dim3 grid((unsigned)(((long)D * D + 255) / 256), tensors);
  merge_shifted_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gram, colsum, mu0, d_slots, mu_slots, world,
                                                                  tensors, D, (float)rows_per_rank);
  BASD_LAUNCH_CHECK();
  return 0;
}

// Projection of the token-space statistics in double precision:
//   K = sym(P G P^T) - (P c)(P c)^T / M      (reference: layer_selector.py:72,88 then :35,91)
// with G = X^T X (fp32, D_in x D_in), c = X^T 1 and the fixed projection P (D_out x D_in).
//
// Why fp64: the spectrum of G spans cond(X)^2 and the rotation mixes every direction into every
// entry, so an fp32 product leaves ~sqrt(D) eps lambda_max of noise on eigenvalues that the
// selector's cross-block gaps 1/(lambda_a - lambda_j) divide by.  On tokens with condition number
// 2e3 the fp32 rotation alone takes the student-gradient cosine against the reference from 0.99999
// to 0.9989 (CPU model, oracle/kernel_model.py); accumulating G itself in fp64 changes nothing.
// The work is tiny (8 GFLOP at C2), so plain DFMA tiles are enough: 64 x 64 output tiles, 4 x 4
// doubles per thread (columns tx + 16 j: conflict-free shared-memory reads), operands converted to double
// on their way through registers into shared memory, the next slab prefetched during the multiply.
#include "common.cuh"

namespace basd {
namespace r64 {

constexpr int TM = 64, TN = 64, TK = 16;

// C (M x N, fp64) = A (M x K, TA) . op(B),  B fp32: K x N row-major, or N x K when TRANS_B.
template <typename TA, bool TRANS_B>
__global__ void __launch_bounds__(256)
dgemm_kernel(int M, int N, int K, const TA* __restrict__ A, int lda, long stride_a,
             const float* __restrict__ B, int ldb, long stride_b, double* __restrict__ C, int ldc,
             long stride_c) {
  __shared__ double As[TK][TM + 2];
  __shared__ double Bs[TK][TN + 2];
  const int prob = blockIdx.z;
  A += (long)prob * stride_a;
  B += (long)prob * stride_b;
  C += (long)prob * stride_c;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  // staging coordinates: a 64 x 16 slab read along its 16-wide (contiguous) side
  const int sr = tid >> 2, sk = (tid & 3) * 4;
  // and a 16 x 64 slab read along its 64-wide side
  const int br = tid >> 4, bc = (tid & 15) * 4;
  // the next slab travels in registers while the current one is multiplied
  double ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int m = m0 + sr, k = k0 + sk + e;
      ra[e] = (m < M && k < K) ? (double)A[(long)m * lda + k] : 0.0;
    }
    if (TRANS_B) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + sr, k = k0 + sk + e;
        rb[e] = (n < N && k < K) ? (double)B[(long)n * ldb + k] : 0.0;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + br, n = n0 + bc + e;
        rb[e] = (k < K && n < N) ? (double)B[(long)k * ldb + n] : 0.0;
      }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) As[sk + e][sr] = ra[e];
    if (TRANS_B) {
#pragma unroll
      for (int e = 0; e < 4; ++e) Bs[sk + e][sr] = rb[e];
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) Bs[br][bc + e] = rb[e];
    }
    __syncthreads();
    if (k0 + TK < K) fetch(k0 + TK);
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];       // one address per half-warp: broadcast
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];      // consecutive lanes, consecutive doubles
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < N) C[(long)m * ldc + n] = acc[i][j];
    }
  }
}

// chat64[prob][i] = sum_k P[i][k] c[prob][k]
__global__ void __launch_bounds__(128)
project_colsum_kernel(const float* __restrict__ P, int d_out, int d_in, const float* __restrict__ c,
                      double* __restrict__ chat64) {
  __shared__ double red[4];
  const int i = blockIdx.x, prob = blockIdx.y;
  const float* row = P + (long)i * d_in;
  const float* cv = c + (long)prob * d_in;
  double s = 0.0;
  for (int k = threadIdx.x; k < d_in; k += 128) s = fma((double)row[k], (double)cv[k], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) chat64[(long)prob * d_out + i] = red[0] + red[1] + red[2] + red[3];
}

// K32 = fp32( sym(K64) - inv_rows chat chat^T ),  chat32 = fp32(chat64): ONE rounding per entry.
__global__ void sym_center_round_kernel(const double* __restrict__ K64, const double* __restrict__ chat64,
                                        int D, double inv_rows, float* __restrict__ K32,
                                        float* __restrict__ chat32) {
  const int prob = blockIdx.y;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  const int i = idx / D, j = idx % D;
  const double* k = K64 + (long)prob * D * D;
  const double* c = chat64 + (long)prob * D;
  const double v = 0.5 * (k[(long)i * D + j] + k[(long)j * D + i]) - inv_rows * (c[i] * c[j]);   // (c_i c_j) first: bitwise symmetric
  K32[(long)prob * D * D + idx] = (float)v;
  if (i == 0) chat32[(long)prob * D + j] = (float)c[j];
}

}  // namespace r64
}  // namespace basd

using namespace basd;

extern "C" long basd_rotate_stats_f64_workspace_bytes(int d_out, int d_in, int batch) {
  return ((long)batch * d_out * d_in + (long)batch * d_out * d_out + (long)batch * d_out) *
         (long)sizeof(double);
}

// proj (d_out x d_in) fp32 shared by the batch; gram (batch, d_in, d_in), colsum (batch, d_in) fp32.
// k_centred (batch, d_out, d_out) and chat (batch, d_out) fp32, each entry rounded once from fp64.
extern "C" int basd_rotate_stats_f64(const float* proj, int d_out, int d_in, const float* gram,
                                     const float* colsum, int batch, double inv_rows, void* workspace,
                                     float* k_centred, float* chat, void* stream) {
  if (batch <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(workspace) & 7) return -3;
  cudaStream_t st = (cudaStream_t)stream;
  double* t1 = static_cast<double*>(workspace);                 // (batch, d_out, d_in)  P G
  double* k64 = t1 + (long)batch * d_out * d_in;                // (batch, d_out, d_out) P G P^T
  double* c64 = k64 + (long)batch * d_out * d_out;              // (batch, d_out)        P c
  dim3 g1((d_in + r64::TN - 1) / r64::TN, (d_out + r64::TM - 1) / r64::TM, batch);
  r64::dgemm_kernel<float, false><<<g1, 256, 0, st>>>(d_out, d_in, d_in, proj, d_in, 0, gram, d_in,
                                                      (long)d_in * d_in, t1, d_in, (long)d_out * d_in);
  BASD_LAUNCH_CHECK();
  dim3 g2((d_out + r64::TN - 1) / r64::TN, (d_out + r64::TM - 1) / r64::TM, batch);
  r64::dgemm_kernel<double, true><<<g2, 256, 0, st>>>(d_out, d_out, d_in, t1, d_in, (long)d_out * d_in,
                                                      proj, d_in, 0, k64, d_out, (long)d_out * d_out);
  BASD_LAUNCH_CHECK();
  dim3 g3(d_out, batch);
  r64::project_colsum_kernel<<<g3, 128, 0, st>>>(proj, d_out, d_in, colsum, c64);
  BASD_LAUNCH_CHECK();
  dim3 g4((unsigned)(((long)d_out * d_out + 255) / 256), batch);
  r64::sym_center_round_kernel<<<g4, 256, 0, st>>>(k64, c64, d_out, inv_rows, k_centred, chat);
  BASD_LAUNCH_CHECK();
  return 0;
}

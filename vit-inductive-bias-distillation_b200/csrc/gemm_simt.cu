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
#include <cstdlib>

namespace basd {

constexpr int BM = 64, BN = 64, BK = 16;

// Loads up to 4 consecutive elements (fewer when `remaining` < 4) as floats.
template <typename T>
__device__ __forceinline__ void load_quad(const T* __restrict__ src, int remaining, bool vec,
                                          float v[4]);
template <>
__device__ __forceinline__ void load_quad<float>(const float* __restrict__ src, int remaining,
                                                 bool vec, float v[4]) {
  if (vec && remaining >= 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (i < remaining) ? src[i] : 0.f;
  }
}
template <>
__device__ __forceinline__ void load_quad<__nv_bfloat16>(const __nv_bfloat16* __restrict__ src,
                                                         int remaining, bool vec, float v[4]) {
  if (vec && remaining >= 4) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(src));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (i < remaining) ? __bfloat162float(src[i]) : 0.f;
  }
}

// Tile shape is a template parameter: 64 or 68 rows/columns.  The per-sample matrices of the
// C2-C4 workloads have 196 rows (3 x 68 = 204 wastes 4 %; 4 x 64 = 256 wasted 31 %).
template <bool TA, bool TB, typename AT, int TM, int TN>
__global__ void __launch_bounds__((TM / 4) * (TN / 4))
sgemm_kernel(int M, int N, int K, const AT* __restrict__ A, int lda, long sA,
             const float* __restrict__ a_shift, const float* __restrict__ B, int ldb, long sB,
             float* __restrict__ C, int ldc, long sC, float alpha,
             const float* __restrict__ alpha_dev, float beta, bool vec) {
  constexpr int NT = (TM / 4) * (TN / 4), TXN = TN / 4;
  __shared__ __align__(16) float As[BK][TM + 4];
  __shared__ __align__(16) float Bs[BK][TN + 4];
  const int batch = blockIdx.z;
  A += (long)batch * sA;
  B += (long)batch * sB;
  C += (long)batch * sC;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // Tiles are filled in quads of 4 elements along the operand's contiguous dimension
    // (one 128-bit / 64-bit global load when the quad is aligned and in range).
    if (!TA) {  // A[m][k], k contiguous: TM rows x 4 quads
      for (int q = tid; q < TM * 4; q += NT) {
        const int m = q >> 2, kk = (q & 3) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if ((m0 + m) < M) {
          const AT* src = A + (long)(m0 + m) * lda + k0 + kk;
          load_quad<AT>(src, K - (k0 + kk), vec, v);
          if (a_shift) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (k0 + kk + i < K) v[i] -= a_shift[k0 + kk + i];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) As[kk + i][m] = v[i];
      }
    } else {  // A stored [k][m], m contiguous: BK rows x TM/4 quads
      for (int q = tid; q < BK * (TM / 4); q += NT) {
        const int kk = q / (TM / 4), mm = (q % (TM / 4)) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if ((k0 + kk) < K) load_quad<AT>(A + (long)(k0 + kk) * lda + m0 + mm, M - (m0 + mm), vec, v);
        *reinterpret_cast<float4*>(&As[kk][mm]) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    if (!TB) {  // B[k][n], n contiguous
      for (int q = tid; q < BK * (TN / 4); q += NT) {
        const int kk = q / (TN / 4), nn = (q % (TN / 4)) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if ((k0 + kk) < K) load_quad<float>(B + (long)(k0 + kk) * ldb + n0 + nn, N - (n0 + nn), vec, v);
        *reinterpret_cast<float4*>(&Bs[kk][nn]) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {  // B stored [n][k], k contiguous
      for (int q = tid; q < TN * 4; q += NT) {
        const int n = q >> 2, kk = (q & 3) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if ((n0 + n) < N) load_quad<float>(B + (long)(n0 + n) * ldb + k0 + kk, K - (k0 + kk), vec, v);
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[kk + i][n] = v[i];
      }
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

// Large-tile variant: 8x8 register micro-tiles (as four 4x4 quadrants so that every shared-memory
// read is a conflict-free 128-bit load), BK = 8, double-buffered shared memory with the next
// k-slab prefetched into registers while the current one is multiplied.  Tile widths 104 or 128:
// two 104-wide tiles cover the 196-row per-sample matrices with 6 % waste.
constexpr int BK8 = 8;

template <bool TA, bool TB, typename AT, int TM, int TN>
__global__ void __launch_bounds__((TM / 8) * (TN / 8))
sgemm8_kernel(int M, int N, int K, const AT* __restrict__ A, int lda, long sA,
              const float* __restrict__ a_shift, const float* __restrict__ B, int ldb, long sB,
              float* __restrict__ C, int ldc, long sC, float alpha,
              const float* __restrict__ alpha_dev, float beta, bool vec) {
  constexpr int NT = (TM / 8) * (TN / 8), TXN = TN / 8;
  constexpr int QA = TM * BK8 / 4, QB = TN * BK8 / 4;          // quads per slab
  constexpr int RA = (QA + NT - 1) / NT, RB = (QB + NT - 1) / NT;
  __shared__ __align__(16) float As[2][BK8][TM];
  __shared__ __align__(16) float Bs[2][BK8][TN];
  const int batch = blockIdx.z;
  A += (long)batch * sA;
  B += (long)batch * sB;
  C += (long)batch * sC;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[RA][4], rb[RB][4];

  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < RA; ++r) {
      const int q = tid + r * NT;
      ra[r][0] = ra[r][1] = ra[r][2] = ra[r][3] = 0.f;
      if (q < QA) {
        if (!TA) {   // A[m][k]: TM rows x 2 quads
          const int m = q >> 1, kk = (q & 1) * 4;
          if ((m0 + m) < M) {
            load_quad<AT>(A + (long)(m0 + m) * lda + k0 + kk, K - (k0 + kk), vec, ra[r]);
            if (a_shift) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (k0 + kk + i < K) ra[r][i] -= a_shift[k0 + kk + i];
            }
          }
        } else {     // A stored [k][m]
          const int kk = q / (TM / 4), mm = (q % (TM / 4)) * 4;
          if ((k0 + kk) < K) load_quad<AT>(A + (long)(k0 + kk) * lda + m0 + mm, M - (m0 + mm), vec, ra[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int q = tid + r * NT;
      rb[r][0] = rb[r][1] = rb[r][2] = rb[r][3] = 0.f;
      if (q < QB) {
        if (!TB) {   // B[k][n]
          const int kk = q / (TN / 4), nn = (q % (TN / 4)) * 4;
          if ((k0 + kk) < K) load_quad<float>(B + (long)(k0 + kk) * ldb + n0 + nn, N - (n0 + nn), vec, rb[r]);
        } else {     // B stored [n][k]
          const int n = q >> 1, kk = (q & 1) * 4;
          if ((n0 + n) < N) load_quad<float>(B + (long)(n0 + n) * ldb + k0 + kk, K - (k0 + kk), vec, rb[r]);
        }
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int r = 0; r < RA; ++r) {
      const int q = tid + r * NT;
      if (q < QA) {
        if (!TA) {
          const int m = q >> 1, kk = (q & 1) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) As[buf][kk + i][m] = ra[r][i];
        } else {
          const int kk = q / (TM / 4), mm = (q % (TM / 4)) * 4;
          *reinterpret_cast<float4*>(&As[buf][kk][mm]) = make_float4(ra[r][0], ra[r][1], ra[r][2], ra[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int q = tid + r * NT;
      if (q < QB) {
        if (!TB) {
          const int kk = q / (TN / 4), nn = (q % (TN / 4)) * 4;
          *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = make_float4(rb[r][0], rb[r][1], rb[r][2], rb[r][3]);
        } else {
          const int n = q >> 1, kk = (q & 1) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) Bs[buf][kk + i][n] = rb[r][i];
        }
      }
    }
  };

  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BK8) {
    const bool more = (k0 + BK8) < K;
    if (more) fetch(k0 + BK8);                    // global loads in flight during the FMAs below
#pragma unroll
    for (int kk = 0; kk < BK8; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][TM / 2 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][TN / 2 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  float scale = alpha;
  if (alpha_dev) scale *= *alpha_dev;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : TM / 2 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : TN / 2 + tx * 4 + (j - 4));
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
                          float* __restrict__ partial, const float* __restrict__ mu0) {
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
  float ma[4], mb[4];                                    // the shift of this thread's staging columns
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ma[i] = (mu0 && (a0 + dd + i) < D) ? mu0[a0 + dd + i] : 0.f;
    mb[i] = (mu0 && (b0 + dd + i) < D) ? mu0[b0 + dd + i] : 0.f;
  }
  for (long r0 = r_begin; r0 < r_end; r0 += BK) {
    const bool k_ok = (r0 + kk) < r_end;
    const T* row = X + (r0 + kk) * (long)D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[kk][dd + i] = (k_ok && (a0 + dd + i) < D) ? to_f32<T>(row[a0 + dd + i]) - ma[i] : 0.f;
      Bs[kk][dd + i] = (k_ok && (b0 + dd + i) < D) ? to_f32<T>(row[b0 + dd + i]) - mb[i] : 0.f;
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
// mu0 / dsum / coef (tensor-core path, nullable): the producer accumulated  Gs = sum x x^T - Mc mu0 mu0^T;
// the Gram of the shifted tokens is  G' = Gs + (Mc - M) mu0 mu0^T - mu0 d^T - d mu0^T  with d = sum (x - mu0)
// and coef = Mc - M (|coef| < one stage of rows): three small terms, no cancellation.
// One CTA per 32 x 32 block of the upper tile triangle: coalesced reads of the slices, the block and -- for
// off-diagonal tile pairs -- its mirror image written through a shared-memory transpose.  (The first version
// walked the output linearly and fetched the lower triangle from the mirrored position: 4-byte reads D floats
// apart, 14 us per 768 x 768 Gram.)  `tile` must be a multiple of 32.
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const float* __restrict__ partial, int slices, int D, int tile,
                   float* __restrict__ gram, float beta, const float* __restrict__ mu0,
                   const float* __restrict__ dsum, float coef) {
  __shared__ float tsm[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  const int ta = bi * 32 / tile, tb = bj * 32 / tile;
  if (ta > tb) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long plane = (long)D * D;
  for (int rr = ty; rr < 32; rr += 8) {
    const int a = bi * 32 + rr, b = bj * 32 + tx;
    float sum = 0.f;
    if (a < D && b < D) {
      const float* p = partial + (long)a * D + b;
      for (int s = 0; s < slices; ++s) sum += p[s * plane];
      if (mu0) {                                         // evaluated on (min, max) so that G'[a][b] == G'[b][a] bitwise
        const int lo = min(a, b), hi = max(a, b);
        sum += fmaf(coef * mu0[lo], mu0[hi], -fmaf(mu0[lo], dsum[hi], dsum[lo] * mu0[hi]));
      }
      const long idx = (long)a * D + b;
      gram[idx] = sum + (beta != 0.f ? beta * gram[idx] : 0.f);
    }
    tsm[rr][tx] = sum;
  }
  if (ta == tb) return;                                  // a diagonal tile holds both of its halves
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int a = bj * 32 + rr, b = bi * 32 + tx;        // the mirror image of (bi*32 + tx, bj*32 + rr)
    if (a < D && b < D) {
      const long idx = (long)a * D + b;
      gram[idx] = tsm[tx][rr] + (beta != 0.f ? beta * gram[idx] : 0.f);
    }
  }
}

// Column sums: one block per 32-column group, rows strided over blockIdx.y, then a
// deterministic second pass.
// mu0 (nullable): the sums are taken of x - mu0 (each term exact in fp32 for bf16 tokens and a bf16 shift), so
// that they stay small next to M mu0 and keep their relative accuracy (see the Gram entry points).
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ X, long rows, int D,
                                      long rows_per_slice, float* __restrict__ partial,
                                      const float* __restrict__ mu0) {
  const int d = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;  // 8 row lanes
  const long r_begin = (long)blockIdx.y * rows_per_slice;
  const long r_end = min(rows, r_begin + rows_per_slice);
  float s = 0.f;
  if (d < D) {
    const float m = mu0 ? mu0[d] : 0.f;
    for (long r = r_begin + sub; r < r_end; r += 8) s += to_f32<T>(X[r * D + d]) - m;
  }
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

static inline int pick_tile(int extent) {
  const int w64 = (extent + 63) / 64 * 64, w68 = (extent + 67) / 68 * 68;
  return w68 < w64 ? 68 : 64;
}

template <bool TA, bool TB, typename AT, int TM, int TN>
static int launch_sgemm_tile(int M, int N, int K, const void* A, int lda, long sA,
                             const float* a_shift, const float* B, int ldb, long sB, float* C,
                             int ldc, long sC, int batch, float alpha, const float* alpha_dev,
                             float beta, cudaStream_t st) {
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, batch);
  // 128-bit operand loads need 4-element pitches/strides and 16-byte aligned bases
  const bool vec = !(lda & 3) && !(ldb & 3) && !(sA & 3) && !(sB & 3) &&
                   !(reinterpret_cast<uintptr_t>(A) & 15) && !(reinterpret_cast<uintptr_t>(B) & 15);
  sgemm_kernel<TA, TB, AT, TM, TN><<<grid, (TM / 4) * (TN / 4), 0, st>>>(
      M, N, K, (const AT*)A, lda, sA, a_shift, B, ldb, sB, C, ldc, sC, alpha, alpha_dev, beta, vec);
  BASD_LAUNCH_CHECK();
  return 0;
}

static inline int pick_tile8(int extent) {
  const int w104 = (extent + 103) / 104 * 104, w128 = (extent + 127) / 128 * 128;
  return w104 < w128 ? 104 : 128;
}

template <bool TA, bool TB, typename AT, int TM, int TN>
static int launch_sgemm8_tile(int M, int N, int K, const void* A, int lda, long sA,
                              const float* a_shift, const float* B, int ldb, long sB, float* C,
                              int ldc, long sC, int batch, float alpha, const float* alpha_dev,
                              float beta, cudaStream_t st) {
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, batch);
  const bool vec = !(lda & 3) && !(ldb & 3) && !(sA & 3) && !(sB & 3) &&
                   !(reinterpret_cast<uintptr_t>(A) & 15) && !(reinterpret_cast<uintptr_t>(B) & 15);
  sgemm8_kernel<TA, TB, AT, TM, TN><<<grid, (TM / 8) * (TN / 8), 0, st>>>(
      M, N, K, (const AT*)A, lda, sA, a_shift, B, ldb, sB, C, ldc, sC, alpha, alpha_dev, beta, vec);
  BASD_LAUNCH_CHECK();
  return 0;
}

template <bool TA, bool TB, typename AT>
static int launch_sgemm_type(int M, int N, int K, const void* A, int lda, long sA,
                             const float* a_shift, const float* B, int ldb, long sB, float* C,
                             int ldc, long sC, int batch, float alpha, const float* alpha_dev,
                             float beta, cudaStream_t st) {
  if (M >= 96 && N >= 96) {
    const int tm8 = pick_tile8(M), tn8 = pick_tile8(N);
#define BASD_ARGS8 M, N, K, A, lda, sA, a_shift, B, ldb, sB, C, ldc, sC, batch, alpha, alpha_dev, beta, st
    if (tm8 == 104 && tn8 == 104) return launch_sgemm8_tile<TA, TB, AT, 104, 104>(BASD_ARGS8);
    if (tm8 == 104) return launch_sgemm8_tile<TA, TB, AT, 104, 128>(BASD_ARGS8);
    if (tn8 == 104) return launch_sgemm8_tile<TA, TB, AT, 128, 104>(BASD_ARGS8);
    return launch_sgemm8_tile<TA, TB, AT, 128, 128>(BASD_ARGS8);
#undef BASD_ARGS8
  }
  const int tm = pick_tile(M), tn = pick_tile(N);
#define BASD_ARGS M, N, K, A, lda, sA, a_shift, B, ldb, sB, C, ldc, sC, batch, alpha, alpha_dev, beta, st
  if (tm == 68 && tn == 68) return launch_sgemm_tile<TA, TB, AT, 68, 68>(BASD_ARGS);
  if (tm == 68) return launch_sgemm_tile<TA, TB, AT, 68, 64>(BASD_ARGS);
  if (tn == 68) return launch_sgemm_tile<TA, TB, AT, 64, 68>(BASD_ARGS);
  return launch_sgemm_tile<TA, TB, AT, 64, 64>(BASD_ARGS);
#undef BASD_ARGS
}

template <bool TA, bool TB>
static int launch_sgemm(int a_dtype, int M, int N, int K, const void* A, int lda, long sA,
                        const float* a_shift, const float* B, int ldb, long sB, float* C, int ldc,
                        long sC, int batch, float alpha, const float* alpha_dev, float beta,
                        cudaStream_t st) {
  if (a_dtype == BASD_DTYPE_BF16)
    return launch_sgemm_type<TA, TB, __nv_bfloat16>(M, N, K, A, lda, sA, a_shift, B, ldb, sB, C, ldc,
                                                    sC, batch, alpha, alpha_dev, beta, st);
  return launch_sgemm_type<TA, TB, float>(M, N, K, A, lda, sA, a_shift, B, ldb, sB, C, ldc, sC, batch,
                                          alpha, alpha_dev, beta, st);
}

// Host-side helpers shared with gram_tc.cu (declared in common.cuh).
int launch_gram_reduce(const float* partial, int slices, int D, int tile, float* gram,
                       cudaStream_t st, const float* mu0, const float* dsum, float coef) {
  if (tile % 32) return -3;
  const unsigned nb = (unsigned)((D + 31) / 32);
  gram_reduce_kernel<<<dim3(nb, nb), 256, 0, st>>>(partial, slices, D, tile, gram, 0.f, mu0, dsum, coef);
  BASD_LAUNCH_CHECK();
  return 0;
}

// out[d] = sum_slices partial[s][d] + coef * mu0[d]   (column sums that rode on the tensor-core Gram: its
// accumulator holds sum x - Mc mu0, the shifted column sum is sum x - M mu0, coef = Mc - M)
__global__ void colsum_fold_kernel(const float* __restrict__ partial, int slices, int D, float* __restrict__ out,
                                   const float* __restrict__ mu0, float coef) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float s = 0.f;
  for (int i = 0; i < slices; ++i) s += partial[(long)i * D + d];
  out[d] = mu0 ? fmaf(coef, mu0[d], s) : s;
}

int launch_colsum_fold(const float* partial, int slices, int D, float* out, const float* mu0, float coef,
                       cudaStream_t st) {
  colsum_fold_kernel<<<(D + 127) / 128, 128, 0, st>>>(partial, slices, D, out, mu0, coef);
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

// Split-K slices of the SIMT Gram: 2,048 rows each for large inputs, but never so few that the grid
// (upper tiles x slices) leaves most of the GPU idle -- C1 (1,024 rows, D = 192) ran 6 CTAs per launch.
static long simt_gram_slices(long rows, int D) {
  const long tiles = (D + basd::BM - 1) / basd::BM, upper = tiles * (tiles + 1) / 2;
  long slices = (rows + 2047) / 2048;
  const long want = (2 * 148 + upper - 1) / upper;            // about two waves of CTAs
  const long by_rows = (rows + 127) / 128;                    // at least 128 rows per slice
  if (slices < want) slices = want < by_rows ? want : by_rows;
  if (slices > 64) slices = 64;
  if (slices < 1) slices = 1;
  return slices;
}

extern "C" long basd_token_gram_simt_workspace_floats(long rows, int D) {
  return simt_gram_slices(rows, D) * ((long)D * D + D);
}

// gram[D*D] = X'^T X', colsum[D] = X'^T 1 for X' = tokens - 1 mu0^T viewed as (rows, D); mu0 (D floats,
// nullable = no shift) is subtracted while the operands are staged (exact in fp32 for bf16 tokens).
// colsum may be null (skipped).
extern "C" int basd_token_gram_simt(const void* tokens, int dtype, long rows, int D, const float* mu0,
                                    float* gram, float* colsum, float* workspace, void* stream) {
  using namespace basd;
  cudaStream_t st = (cudaStream_t)stream;
  const long slices = simt_gram_slices(rows, D);
  long per = (rows + slices - 1) / slices;
  per = (per + BK - 1) / BK * BK;
  const int tiles = (D + BM - 1) / BM;
  dim3 grid(tiles * (tiles + 1) / 2, (unsigned)slices);
  float* part_g = workspace;
  float* part_c = workspace + slices * (long)D * D;
  dim3 cgrid((D + 31) / 32, (unsigned)slices);
  if (dtype == BASD_DTYPE_BF16) {
    token_gram_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)tokens, rows, D, per, part_g, mu0);
    if (colsum)
      colsum_partial_kernel<__nv_bfloat16><<<cgrid, 256, 0, st>>>((const __nv_bfloat16*)tokens, rows,
                                                                  D, per, part_c, mu0);
  } else {
    token_gram_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)tokens, rows, D, per,
                                                           part_g, mu0);
    if (colsum)
      colsum_partial_kernel<float><<<cgrid, 256, 0, st>>>((const float*)tokens, rows, D, per, part_c, mu0);
  }
  BASD_LAUNCH_CHECK();
  static_assert(BM % 32 == 0, "gram_reduce_kernel works on 32 x 32 blocks of the tile triangle");
  const unsigned nb = (unsigned)((D + 31) / 32);
  gram_reduce_kernel<<<dim3(nb, nb), 256, 0, st>>>(part_g, (int)slices, D, BM, gram, 0.f, nullptr, nullptr, 0.f);
  if (colsum) colsum_reduce_kernel<<<(D + 127) / 128, 128, 0, st>>>(part_c, (int)slices, D, colsum);
  BASD_LAUNCH_CHECK();
  return 0;
}

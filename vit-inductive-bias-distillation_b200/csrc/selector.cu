// Small device-side pieces of the Grassmannian layer selector.  Everything here is O(D^2)
// per matrix at most; the heavy lifting is in gemm_*.cu and jacobi.cu.  No host syncs: the
// Marchenko-Pastur ranks live in device memory and downstream kernels read them there
// (the reference syncs the host twice per teacher layer, layer_selector.py:17,19).
#include "common.cuh"
#include <math.h>

namespace basd {

// K = sym(G) - inv_rows * c c^T      (reference: layer_selector.py:35,91 centring, in Gram form)
__global__ void center_gram_kernel(const float* __restrict__ G, const float* __restrict__ c,
                                   int D, float inv_rows, float* __restrict__ K) {
  const int prob = blockIdx.y;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  const int i = idx / D, j = idx % D;
  const float* g = G + (long)prob * D * D;
  float v = 0.5f * (g[(long)i * D + j] + g[(long)j * D + i]);
  if (c) v -= inv_rows * c[(long)prob * D + i] * c[(long)prob * D + j];
  K[(long)prob * D * D + idx] = v;
}

// MP rank of each teacher layer from the spectrum of the uncentred second moment
// (reference: layer_selector.py:8-20 and :74).  lam: (L, D) in any order.
// n_eig = min(M, D): with fewer rows than dimensions the reference switches to the M x M Gram
// (layer_selector.py:14-15), whose spectrum is the n_eig largest eigenvalues of the D x D one --
// the median is taken over that population (the D - n_eig structural zeros are left out).
__global__ void mp_rank_kernel(const float* __restrict__ lam, int D, int n_eig, float aspect /* D/M */,
                               int cap, int* __restrict__ ranks, float* __restrict__ edges) {
  extern __shared__ float v[];
  __shared__ float median;
  __shared__ int count;
  const int layer = blockIdx.x;
  const float* l = lam + (long)layer * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) v[i] = l[i];
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  const int want = (D - n_eig) + (n_eig - 1) / 2;   // torch.median: lower middle of the ascending order
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    int below = 0;
    const float mine = v[i];
    for (int j = 0; j < D; ++j) below += (v[j] < mine) || (v[j] == mine && j < i);
    if (below == want) median = mine;
  }
  __syncthreads();
  const float root = 1.f + sqrtf(aspect);
  const float edge = median * root * root;
  int local = 0;
  for (int i = threadIdx.x; i < D; i += blockDim.x) local += (v[i] > edge);
  atomicAdd(&count, local);
  __syncthreads();
  if (threadIdx.x == 0) {
    ranks[layer] = min(count, cap);
    if (edges) { edges[2 * layer] = median; edges[2 * layer + 1] = edge; }
  }
}

// MP rank of the UNCENTRED second moment from the eigendecomposition of the CENTRED one.
// K_u = K_c + rho * c c^T is a rank-one update, so with K_c = V diag(lam) V^T and y = V^T c the
// eigenvalues mu of K_u are the roots of the secular function
//     f(x) = 1 + rho * sum_j y_j^2 / (lam_j - x),
// they interlace the lam_j, and  #{mu > x} = #{lam > x} + [f(x) < 0].   That counting function
// is monotone in x, so the lower median (torch.median, layer_selector.py:17) is found by
// bisection on it and the rank (:19) is one more evaluation -- no second eigenproblem per
// teacher layer.  fp64 for the (tiny) secular sums.  edges: (layers, 3) = median, lambda_plus,
// tie flag (1 if the count changes within +-1e-4 relative of lambda_plus).
__device__ int secular_count(const double* lam, const double* y2, int D, double rho, double x,
                             double* red_d, int* red_i) {
  double f = 0.0;
  int c = 0;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    double diff = lam[j] - x;
    c += diff > 0.0;
    if (diff == 0.0) diff = 1e-300;
    f += y2[j] / diff;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    f += __shfl_xor_sync(0xffffffffu, f, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) { red_d[warp] = f; red_i[warp] = c; }
  __syncthreads();
  f = 0.0;
  c = 0;
  for (int w = 0; w < nw; ++w) { f += red_d[w]; c += red_i[w]; }
  return c + ((1.0 + rho * f) < 0.0 ? 1 : 0);
}

__global__ void mp_rank_secular_kernel(const float* __restrict__ lam_c, const float* __restrict__ y,
                                       int D, int n_eig, double rho, float aspect, int cap,
                                       int* __restrict__ ranks, float* __restrict__ edges) {
  extern __shared__ double sd[];
  double* lam = sd;          // D
  double* y2 = sd + D;       // D
  __shared__ double red_d[32];
  __shared__ int red_i[32];
  const int layer = blockIdx.x;
  double hi = 0.0, ysum = 0.0;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const double l = (double)lam_c[(long)layer * D + j];
    const double v = (double)y[(long)layer * D + j];
    lam[j] = l;
    y2[j] = v * v;
  }
  __syncthreads();
  for (int j = 0; j < D; ++j) { hi = fmax(hi, lam[j]); ysum += y2[j]; }   // tiny, every thread
  hi = hi + rho * ysum + 1e-30;                       // mu_max <= lam_max + rho |y|^2
  double lo = fmin(0.0, -hi);                         // K_c is PSD up to rounding
  const int want = n_eig - (n_eig - 1) / 2;           // lower median of the n_eig largest = want-th largest
  // largest x with #{mu > x} >= want  ==  the want-th largest root
  for (int it = 0; it < 80; ++it) {
    const double mid = 0.5 * (lo + hi);
    const int cnt = secular_count(lam, y2, D, rho, mid, red_d, red_i);
    if (cnt >= want) lo = mid; else hi = mid;
  }
  const double median = 0.5 * (lo + hi);
  const double root = 1.0 + sqrt((double)aspect);
  const double edge = median * root * root;
  const int rank = secular_count(lam, y2, D, rho, edge, red_d, red_i);
  const int r_lo = secular_count(lam, y2, D, rho, edge * (1.0 - 1e-4), red_d, red_i);
  const int r_hi = secular_count(lam, y2, D, rho, edge * (1.0 + 1e-4), red_d, red_i);
  if (threadIdx.x == 0) {
    ranks[layer] = min(rank, cap);
    if (edges) {
      edges[3 * layer] = (float)median;
      edges[3 * layer + 1] = (float)edge;
      edges[3 * layer + 2] = (r_lo != r_hi) ? 1.f : 0.f;
    }
  }
}

// dims[i*L + l] = ranks[l]
__global__ void expand_ranks_kernel(const int* __restrict__ ranks, int E, int L,
                                    int* __restrict__ dims) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < E * L) dims[idx] = ranks[idx % L];
}

// Zero everything outside the leading dims[prob] x dims[prob] block; src may equal dst.
__global__ void mask_block_kernel(const float* __restrict__ src, float* __restrict__ dst, int D,
                                  const int* __restrict__ dims) {
  const int prob = blockIdx.y;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  const int k = dims[prob];
  const int i = idx / D, j = idx % D;
  const long off = (long)prob * D * D + idx;
  dst[off] = (i < k && j < k) ? src[off] : 0.f;
}

// Spectrally weighted squared Grassmann distance (reference: layer_selector.py:100-105).
// sig: (E*L, D) principal-angle cosines (descending), lam_c: (L, D) centred teacher
// eigenvalues (descending) -> sw = sqrt(lam).
__global__ void angle_distance_kernel(const float* __restrict__ sig, const float* __restrict__ lam_c,
                                      const int* __restrict__ ranks, int D, int L,
                                      float* __restrict__ dist) {
  __shared__ float red[32];
  const int prob = blockIdx.x, layer = prob % L;
  const int k = ranks[layer];
  const float lim = 1.f - 1.1920929e-07f;
  float num = 0.f, den = 0.f;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const float sw = sqrtf(fmaxf(lam_c[(long)layer * D + j], 0.f));
    const float th = acosf(fminf(sig[(long)prob * D + j], lim));
    num = fmaf(sw, th * th, num);
    den += sw;
  }
  num = block_sum(num, red);
  den = block_sum(den, red);
  if (threadIdx.x == 0) dist[prob] = num / den;   // k == 0 -> 0/0 = NaN like the reference
}

// weights[i,:] = softmax(-dist[i,:] / softplus(log_temp[i]))   (reference: :67,:107-108)
__global__ void mix_weights_kernel(const float* __restrict__ dist, const float* __restrict__ log_temp,
                                   int L, float* __restrict__ weights, float* __restrict__ temps) {
  const int i = blockIdx.x;
  const float x = log_temp[i];
  const float tau = (x > 20.f) ? x : log1pf(expf(x));
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int l = 0; l < L; ++l) mx = fmaxf(mx, -dist[i * L + l] / tau);
    float z = 0.f;
    for (int l = 0; l < L; ++l) z += expf(-dist[i * L + l] / tau - mx);
    for (int l = 0; l < L; ++l) weights[i * L + l] = expf(-dist[i * L + l] / tau - mx) / z;
    temps[i] = tau;
  }
}

// Backward of softmax(-d/tau), tau = softplus(log_temp):
//   d_dist (E,L), d_log_temp (E)      (autograd of layer_selector.py:107-108)
__global__ void mix_weights_bwd_kernel(const float* __restrict__ d_weights,
                                       const float* __restrict__ weights,
                                       const float* __restrict__ dist,
                                       const float* __restrict__ log_temp, int L, float scale,
                                       float* __restrict__ d_dist, float* __restrict__ d_log_temp) {
  const int i = blockIdx.x;
  if (threadIdx.x != 0) return;
  const float x = log_temp[i];
  const float tau = (x > 20.f) ? x : log1pf(expf(x));
  float dot = 0.f;
  for (int l = 0; l < L; ++l) dot += weights[i * L + l] * d_weights[i * L + l];
  float d_tau = 0.f;
  for (int l = 0; l < L; ++l) {
    const float dy = weights[i * L + l] * (d_weights[i * L + l] - dot);
    d_dist[i * L + l] = -dy / tau;
    d_tau += dy * dist[i * L + l];
  }
  d_tau /= tau * tau;
  d_log_temp[i] = scale * d_tau / (1.f + expf(-x));
}

// Scales the rows of Uxt (E*L, D, D) by d sigma_m (autograd of acos/clamp/pow/weighted mean,
// layer_selector.py:100-105): row m *= d_dist * sw_m/sum(sw) * 2 theta_m * (-1/sqrt(1-s^2)).
__global__ void scale_rows_dsigma_kernel(float* __restrict__ Uxt, const float* __restrict__ sig,
                                         const float* __restrict__ lam_c,
                                         const int* __restrict__ ranks,
                                         const float* __restrict__ d_dist, int D, int L) {
  __shared__ float red[32];
  const int prob = blockIdx.x, layer = prob % L;
  const int k = ranks[layer];
  float den = 0.f;
  for (int j = threadIdx.x; j < k; j += blockDim.x)
    den += sqrtf(fmaxf(lam_c[(long)layer * D + j], 0.f));
  den = block_sum(den, red);
  const float lim = 1.f - 1.1920929e-07f;
  const float dd = d_dist[prob];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int m = warp; m < D; m += nw) {
    float f = 0.f;
    if (m < k) {
      const float s = sig[(long)prob * D + m];
      if (s < lim) {
        const float sw = sqrtf(fmaxf(lam_c[(long)layer * D + m], 0.f));
        const float th = acosf(s);
        f = dd * (sw / den) * 2.f * th * (-rsqrtf(fmaxf(1.f - s * s, 1e-30f)));
      }
    }
    float* row = Uxt + ((long)prob * D + m) * D;
    for (int a = lane; a < D; a += 32) row[a] *= f;
  }
}

// Omega_i[j,a] = sum_l [a < k_l <= j] block_l[j,a] / (lam_a - lam_j)   
// block: (E*L, D, D); lam_s: (E, D) descending student eigenvalues; out: (E, D, D).
__global__ void omega_accumulate_kernel(const float* __restrict__ block,
                                        const float* __restrict__ lam_s,
                                        const int* __restrict__ ranks, int D, int L,
                                        float* __restrict__ omega) {
  const int i = blockIdx.y;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  const int j = idx / D, a = idx % D;
  float acc = 0.f;
  if (j > a) {
    const float gap = lam_s[(long)i * D + a] - lam_s[(long)i * D + j];
    const float inv = (gap > 0.f) ? 1.f / gap : 0.f;
    for (int l = 0; l < L; ++l) {
      const int k = ranks[l];
      if (a < k && k <= j) acc += block[((long)(i * L + l) * D + j) * D + a];
    }
    acc *= inv;
  }
  omega[(long)i * D * D + idx] = acc;
}

// out = in + in^T (per matrix)
__global__ void symmetrize_add_kernel(const float* __restrict__ in, int D, float* __restrict__ out) {
  const int prob = blockIdx.y;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  const int i = idx / D, j = idx % D;
  const float* m = in + (long)prob * D * D;
  out[(long)prob * D * D + idx] = m[(long)i * D + j] + m[(long)j * D + i];
}

// K <- K + rel_shift * max(diag K) * I, one block per problem.  The eigenvectors of K + delta I are those
// of K; the shift only keeps the pivoted Cholesky behind sym_eig away from noise-sized pivots when fp32
// rounding has pushed the smallest eigenvalues of a nearly singular Gram to zero or below (a noise pivot
// divides noise off-diagonals into an O(1) garbage column; cutting the factorisation early instead drops
// up to (D - r) cut-sized pivots of mass, more than the eigenvalue gap at the rank boundary).
__global__ void shift_diag_kernel(float* __restrict__ K, int D, float rel_shift) {
  __shared__ float red[32];
  float* k = K + (long)blockIdx.x * D * D;
  float mx = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) mx = fmaxf(mx, k[(long)i * D + i]);
  mx = block_max(mx, red);
  const float delta = rel_shift * mx;
  for (int i = threadIdx.x; i < D; i += blockDim.x) k[(long)i * D + i] += delta;
}

// ---- null-space completion (fewer token rows than dimensions, layer_selector.py:14-15 regime) ----
// sym_eig leaves the eigenvector rows of a rank-deficient Gram beyond its rank r as zeros.  The thin-SVD
// backward needs them: its (I - V V^T) term acts on exactly that complement.  P = I - V_r^T V_r is the
// orthogonal projector onto it, and the pivoted Cholesky factor of a projector has orthonormal columns
// (P = L L^T and P^2 = P give L^T L = I), so rows 0..D-r-1 of LT are an orthonormal basis of the null space.
// dims_out[problem] = D when rows are missing (trace of V^T V = number of unit rows < D), else 0: the
// Cholesky launch that follows exits at once for complete bases (the common case costs four tiny launches).
__global__ void projector_complement_kernel(const float* __restrict__ vtv, int D, float* __restrict__ P,
                                            int* __restrict__ dims_out) {
  __shared__ float red[32];
  const int prob = blockIdx.y;
  const float* g = vtv + (long)prob * D * D;
  if (blockIdx.x == 0) {
    float tr = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) tr += g[(long)i * D + i];
    tr = block_sum(tr, red);
    if (threadIdx.x == 0) dims_out[prob] = ((float)D - tr > 0.5f) ? D : 0;
  }
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * D) return;
  const int i = idx / D, j = idx % D;
  P[(long)prob * D * D + idx] = (i == j ? 1.f : 0.f) - 0.5f * (g[(long)i * D + j] + g[(long)j * D + i]);
}

// Vt rows r..D-1 (zeros) <- LT rows 0..D-r-1, r = D - rank_P[problem]
__global__ void place_complement_kernel(float* __restrict__ Vt, const float* __restrict__ LT,
                                        const int* __restrict__ rank_p, int D) {
  const int prob = blockIdx.y;
  const int nfill = min(rank_p[prob], D);
  const int r = D - nfill;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nfill * D) return;
  const int i = idx / D, c = idx % D;
  Vt[((long)prob * D + r + i) * D + c] = LT[((long)prob * D + i) * D + c];
}

}  // namespace basd

using namespace basd;
#define ST ((cudaStream_t)stream)
static inline unsigned blocks_for(long n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" int basd_center_gram(const float* G, const float* colsum, int D, float inv_rows,
                                float* K, int batch, void* stream) {
  if (batch <= 0) return 0;
  dim3 grid(blocks_for((long)D * D, 256), batch);
  center_gram_kernel<<<grid, 256, 0, ST>>>(G, colsum, D, inv_rows, K);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_mp_rank(const float* lam, int D, long rows, int cap, int* ranks, float* edges,
                            int layers, void* stream) {
  if (layers <= 0) return 0;
  mp_rank_kernel<<<layers, 256, D * sizeof(float), ST>>>(lam, D, (int)(rows < D ? rows : D),
                                                         (float)((double)D / (double)rows),
                                                         cap, ranks, edges);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_mp_rank_secular(const float* lam_c, const float* y, int D, long rows, int cap,
                                    int* ranks, float* edges, int layers, void* stream) {
  if (layers <= 0) return 0;
  mp_rank_secular_kernel<<<layers, 128, 2 * D * sizeof(double), ST>>>(
      lam_c, y, D, (int)(rows < D ? rows : D), 1.0 / (double)rows, (float)((double)D / (double)rows), cap,
      ranks, edges);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_expand_ranks(const int* ranks, int E, int L, int* dims, void* stream) {
  expand_ranks_kernel<<<blocks_for(E * L, 128), 128, 0, ST>>>(ranks, E, L, dims);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_mask_block(const float* src, float* dst, int D, const int* dims, int batch,
                               void* stream) {
  if (batch <= 0) return 0;
  dim3 grid(blocks_for((long)D * D, 256), batch);
  mask_block_kernel<<<grid, 256, 0, ST>>>(src, dst, D, dims);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_angle_distance(const float* sig, const float* lam_c, const int* ranks, int D,
                                   int E, int L, float* dist, void* stream) {
  angle_distance_kernel<<<E * L, 128, 0, ST>>>(sig, lam_c, ranks, D, L, dist);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_mix_weights(const float* dist, const float* log_temp, int E, int L,
                                float* weights, float* temps, void* stream) {
  mix_weights_kernel<<<E, 32, 0, ST>>>(dist, log_temp, L, weights, temps);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_mix_weights_bwd(const float* d_weights, const float* weights, const float* dist,
                                    const float* log_temp, int E, int L, float scale, float* d_dist,
                                    float* d_log_temp, void* stream) {
  mix_weights_bwd_kernel<<<E, 32, 0, ST>>>(d_weights, weights, dist, log_temp, L, scale, d_dist,
                                          d_log_temp);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_scale_rows_dsigma(float* Uxt, const float* sig, const float* lam_c,
                                      const int* ranks, const float* d_dist, int D, int E, int L,
                                      void* stream) {
  scale_rows_dsigma_kernel<<<E * L, 256, 0, ST>>>(Uxt, sig, lam_c, ranks, d_dist, D, L);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_omega_accumulate(const float* block, const float* lam_s, const int* ranks, int D,
                                     int E, int L, float* omega, void* stream) {
  dim3 grid(blocks_for((long)D * D, 256), E);
  omega_accumulate_kernel<<<grid, 256, 0, ST>>>(block, lam_s, ranks, D, L, omega);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_shift_diag(float* K, int D, float rel_shift, int batch, void* stream) {
  if (batch <= 0) return 0;
  shift_diag_kernel<<<batch, 256, 0, ST>>>(K, D, rel_shift);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_projector_complement(const float* vtv, int D, float* P, int* dims_out, int batch,
                                         void* stream) {
  if (batch <= 0) return 0;
  dim3 grid(blocks_for((long)D * D, 256), batch);
  projector_complement_kernel<<<grid, 256, 0, ST>>>(vtv, D, P, dims_out);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_place_complement(float* Vt, const float* LT, const int* rank_p, int D, int batch,
                                     void* stream) {
  if (batch <= 0) return 0;
  dim3 grid(blocks_for((long)D * D, 256), batch);
  place_complement_kernel<<<grid, 256, 0, ST>>>(Vt, LT, rank_p, D);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_symmetrize_add(const float* in, int D, float* out, int batch, void* stream) {
  dim3 grid(blocks_for((long)D * D, 256), batch);
  symmetrize_add_kernel<<<grid, 256, 0, ST>>>(in, D, out);
  BASD_LAUNCH_CHECK();
  return 0;
}

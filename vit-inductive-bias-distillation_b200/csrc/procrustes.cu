// Attention-weighted Procrustes loss, per-sample glue kernels
// (reference: relational.py:34-50 and its autograd).  The dense work between these kernels
// (N x N Grams, factor products) runs in gemm_*.cu, the factorisations in jacobi.cu.
//
// Per sample:  A = sqrt(w)(S - mu_s), Bm = sqrt(w)(R' - mu_t)
//   each side has a factor F with F F^T = K (its N x N Gram): the pivoted-Cholesky factor of K
//   when D > N ("Gram side", r = N columns), the tokens themselves when D <= N ("direct side",
//   r = D).  X = F_s^T F_t = U S V^T,  f = tr K_s + tr K_t - 2 sum(S)
//   Gram side:   df/dS = 2 sqrt(w) (I - Y_A) A,   Y_A = (F_t V) S^+ (F_t V)^T
//   direct side: df/dS = 2 sqrt(w) (A - (F_t V) U^T)        (unit vectors only, no 1/sigma)
//   df/dw_n = (K_s[n,n] + K_t[n,n] - 2 [F_s U (F_t V)^T]_{nn}) / w_n
#include "common.cuh"

namespace basd {

// out[n,:] = sqrt(w_n) (x[n,:] - sum_m w_m x[m,:])           (relational.py:36-43)
// Columns are independent, so a sample is split over gridDim.y column slices of CW columns;
// inside a block every thread owns VEC adjacent columns (one 128-bit load) and the row groups
// stride over the tokens: pass 1 accumulates the weighted column means, pass 2 re-reads the
// (L2-resident) rows and writes fp32.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) { load8(p, v); }
};

template <typename T, int CW>
__global__ void __launch_bounds__(256)
weighted_center_vec_kernel(const T* __restrict__ X, long strideX, const float* __restrict__ W,
                           long strideW, int N, int D, float* __restrict__ out, long strideO) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int CT = CW / VEC;                 // threads across the columns of the slice
  constexpr int RG = 256 / CT;                 // row groups
  extern __shared__ float sm[];
  float* w = sm;                               // N
  float* rw = sm + N;                          // N
  float* part = rw + N;                        // RG x CW partial means
  const int s = blockIdx.x;
  const int ct = threadIdx.x % CT, rg = threadIdx.x / CT;
  const int c0 = blockIdx.y * CW + ct * VEC;
  const T* x = X + (long)s * strideX;
  float* o = out + (long)s * strideO;
  for (int n = threadIdx.x; n < N; n += 256) {
    const float v = W[(long)s * strideW + n];
    w[n] = v;
    rw[n] = sqrtf(v);
  }
  __syncthreads();
  const bool live = c0 < D && rg < RG;
  float mu[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) mu[i] = 0.f;
  if (live) {
    for (int n = rg; n < N; n += RG) {
      float v[VEC];
      Vec16<T>::load(x + (long)n * D + c0, v);
      const float wn = w[n];
#pragma unroll
      for (int i = 0; i < VEC; ++i) mu[i] = fmaf(wn, v[i], mu[i]);
    }
  }
  if (rg < RG) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) part[rg * CW + ct * VEC + i] = mu[i];
  }
  __syncthreads();
  if (live) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float t = 0.f;
      for (int g = 0; g < RG; ++g) t += part[g * CW + ct * VEC + i];   // same order in every row group
      mu[i] = t;
    }
    for (int n = rg; n < N; n += RG) {
      float v[VEC];
      Vec16<T>::load(x + (long)n * D + c0, v);
      const float r = rw[n];
      float* dst = o + (long)n * D + c0;
#pragma unroll
      for (int i = 0; i < VEC; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(r * (v[i] - mu[i]), r * (v[i + 1] - mu[i + 1]),
                                                          r * (v[i + 2] - mu[i + 2]), r * (v[i + 3] - mu[i + 3]));
    }
  }
}

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

// rows2 (rq x rq) = P^T X: row j = sigma_j q_j^T;  Pt (rq x rp): row j = p_j^T (unit).
// Produces per sample:
//   sig[j] = |rows2_j| (refined singular values), nuc = sum_j sig[j], keep_j = sig_j > rel_floor * max(sig)
//   rows2[j,:] <- keepq_j * q_j^T * sig_j^(eq/2)       Pt[j,:] <- keep_j * p_j^T * sig_j^(ep/2)
//   pic[j]      = keep_j * sig_j^(-(eq+ep)/2)          (eq, ep in {-1, 0, +1})
// so that images I = rows . F^T of the two sides recombine as  sum_j pic_j I_q[j,n] I_p[j,n].
// keepq_j = sig_j > rel_floor_q * max(sig) with rel_floor_q >= rel_floor: the q_j are DERIVED vectors
// (normalised rows of P^T G^T, direction error ~ eps sigma_max / sigma_j), while the p_j come out of the
// Jacobi sweep orthogonal to working precision.  The operator built from the q images,
// Y_p = sum_j (F_q q_j)(F_q q_j)^T / sigma_j, divides that error by sigma_j once more, so it takes the
// higher floor; the one built from the p images keeps every direction above rel_floor.
__device__ __forceinline__ float half_power(float sg, int e2) {
  return e2 == 0 ? 1.f : (e2 < 0 ? rsqrtf(sg) : sqrtf(sg));
}

__global__ void __launch_bounds__(512)
procrustes_rows_finish_kernel(float* __restrict__ rows2, int rq, int ldr, long stride_r,
                              float* __restrict__ Pt, int rp, int ldp, long stride_p,
                              float rel_floor, float rel_floor_q, int eq, int ep,
                              float* __restrict__ sig, float* __restrict__ nuc,
                              float* __restrict__ pic) {
  extern __shared__ float sm[];
  float* nrm = sm;        // rq
  float* red = sm + rq;   // 32
  const int s = blockIdx.x;
  float* R = rows2 + (long)s * stride_r;
  float* U = Pt + (long)s * stride_p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < rq; r += nw) {
    float a = 0.f;
    for (int c = lane; c < rq; c += 32) { const float v = R[(long)r * ldr + c]; a = fmaf(v, v, a); }
    a = warp_sum(a);
    if (lane == 0) nrm[r] = sqrtf(a);
  }
  __syncthreads();
  float mx = 0.f, tot = 0.f;
  for (int r = threadIdx.x; r < rq; r += blockDim.x) { mx = fmaxf(mx, nrm[r]); tot += nrm[r]; }
  mx = block_max(mx, red);
  tot = block_sum(tot, red);
  const float floor_v = rel_floor * mx;
  const float floor_q = fmaxf(rel_floor, rel_floor_q) * mx;
  for (int r = warp; r < rq; r += nw) {
    const float sg = nrm[r];
    const bool keep = sg > floor_v && sg > 0.f;
    const bool keepq = keep && sg > floor_q;
    const float fq = keepq ? half_power(sg, eq) / sg : 0.f;    // rows2_j / sig_j = q_j^T
    const float fp = keep ? half_power(sg, ep) : 0.f;
    for (int c = lane; c < rq; c += 32) R[(long)r * ldr + c] *= fq;
    for (int c = lane; c < rp; c += 32) U[(long)r * ldp + c] *= fp;
    if (lane == 0) {
      sig[(long)s * rq + r] = sg;
      if (pic) {
        const int e2 = -(eq + ep);                             // pic = sig^(e2/2), e2 in [-2, 2]
        const float h = half_power(sg, e2);
        pic[(long)s * rq + r] = keep ? ((e2 == 2 || e2 == -2) ? h * h : h) : 0.f;
      }
    }
  }
  if (threadIdx.x == 0) nuc[s] = tot;
}

// Per sample, with the images IA = rows2' . F_q^T and IB = Pt' . F_p^T (rq x N each):
//   f = tr_s + tr_t - 2 nuc
//   pi[n] = sum_j pic[j] IA[j,n] IB[j,n]
//   gw[n] = ((ks[n] + kt[n] - 2 pi[n]) / w[n] - f) / total        (d f / d w~)
//   for a Gram side, Y (N x N, = I_other^T I_other) becomes M = 2 diag(sqrt w) (I - Y) in place;
//   pass a null Y for a direct side.
__global__ void __launch_bounds__(512)
procrustes_grad_prep_kernel(float* __restrict__ YA, float* __restrict__ YB,
                            const float* __restrict__ IA, const float* __restrict__ IB, int N,
                            int rq, int ldi, long stride_i, int ldy, long stride_y,
                            const float* __restrict__ pic,
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
  const float* fa = IA + (long)s * stride_i;
  const float* fb = IB + (long)s * stride_i;
  const float tot = totals[s];
  // pi[n] = sum_j pic[j] IA[j,n] IB[j,n]: 4 groups of 128 threads split j, each thread one n
  // per pass (coalesced along n), partials folded through shared memory
  __shared__ float pis[4][128];
  const int tn = threadIdx.x & 127, tg = threadIdx.x >> 7;
  for (int n0 = 0; n0 < N; n0 += 128) {
    const int n = n0 + tn;
    float pi = 0.f;
    if (n < N)
      for (int j = tg; j < rq; j += 4)
        pi = fmaf(fb[(long)j * ldi + n] * fa[(long)j * ldi + n], pic[(long)s * rq + j], pi);
    pis[tg][tn] = pi;
    __syncthreads();
    if (tg == 0 && n < N) {
      pi = (pis[0][tn] + pis[1][tn]) + (pis[2][tn] + pis[3][tn]);
      const float wn = w[(long)s * N + n];
      gw[(long)s * N + n] = ((ks[(long)s * N + n] + kt[(long)s * N + n] - 2.f * pi) / wn - f) / tot;
    }
    __syncthreads();
  }
  float* ya = YA ? YA + (long)s * stride_y : nullptr;
  float* yb = YB ? YB + (long)s * stride_y : nullptr;
  if (!ya && !yb) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bool vec = ((N | ldy) & 3) == 0 && (stride_y & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(YA) | reinterpret_cast<uintptr_t>(YB)) & 15) == 0;
  for (int r = warp; r < N; r += nw) {
    const float sc = 2.f * sqrtf(w[(long)s * N + r]);
    if (vec) {                                        // 128-bit read-modify-write of the row
      for (int c = 4 * lane; c < N; c += 128) {
        const float4 id = make_float4(r == c ? 1.f : 0.f, r == c + 1 ? 1.f : 0.f, r == c + 2 ? 1.f : 0.f,
                                      r == c + 3 ? 1.f : 0.f);
        if (ya) {
          float4* q = reinterpret_cast<float4*>(ya + (long)r * ldy + c);
          const float4 v = *q;
          *q = make_float4(sc * (id.x - v.x), sc * (id.y - v.y), sc * (id.z - v.z), sc * (id.w - v.w));
        }
        if (yb) {
          float4* q = reinterpret_cast<float4*>(yb + (long)r * ldy + c);
          const float4 v = *q;
          *q = make_float4(sc * (id.x - v.x), sc * (id.y - v.y), sc * (id.z - v.z), sc * (id.w - v.w));
        }
      }
      continue;
    }
    for (int c = lane; c < N; c += 32) {
      const float id = (r == c) ? 1.f : 0.f;
      if (ya) ya[(long)r * ldy + c] = sc * (id - ya[(long)r * ldy + c]);
      if (yb) yb[(long)r * ldy + c] = sc * (id - yb[(long)r * ldy + c]);
    }
  }
}

// Direct side: T (N x D, in/out) <- 2 sqrt(w_n) (A - T)  with T = I_other^T . own_vectors on entry.
__global__ void __launch_bounds__(256)
procrustes_direct_grad_kernel(const float* __restrict__ A, float* __restrict__ T,
                              const float* __restrict__ w, int N, int D) {
  const int s = blockIdx.y;
  const long base = (long)s * N * D;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)N * D;
       e += (long)gridDim.x * blockDim.x) {
    const int n = (int)(e / D);
    T[base + e] = 2.f * sqrtf(w[(long)s * N + n]) * (A[base + e] - T[base + e]);
  }
}

// dst (token dtype) = alpha * alpha_dev[0] * src (fp32), elementwise.
template <typename T>
__global__ void scale_out_kernel(const float* __restrict__ src, T* __restrict__ dst, long n,
                                 float alpha, const float* __restrict__ alpha_dev) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const float sc = alpha_dev ? alpha * alpha_dev[0] : alpha;
  if (i < n) dst[i] = from_f32<T>(sc * src[i]);
}

// geo_terms[i] = mean_b f[i,b]; geo = mean_i geo_terms[i]    (relational.py:50, combined.py:76)
// A non-finite mixing weight (MP rank 0 -> 0/0 at layer_selector.py:105) poisons the mixed
// tokens of the reference; fmax/fmin-style guards in the kernels in between would swallow the
// NaN, so it is re-asserted here: the loss is NaN exactly when the reference's is.
__global__ void geo_reduce_kernel(const float* __restrict__ f, int E, int B,
                                  const float* __restrict__ weights, int n_weights,
                                  float* __restrict__ geo_terms, float* __restrict__ geo) {
  __shared__ float red[32];
  float bad = 0.f;
  for (int i = threadIdx.x; i < n_weights; i += blockDim.x)
    if (!isfinite(weights[i])) bad = 1.f;
  bad = block_max(bad, red);
  float total = 0.f;
  for (int i = 0; i < E; ++i) {
    float a = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) a += f[(long)i * B + b];
    a = block_sum(a, red) / (float)B;
    if (bad > 0.f) a = __int_as_float(0x7fc00000);
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
  constexpr int CW = 128;                                 // columns per block
  const bool aligned = (D % 8 == 0) && (stride_x % 8 == 0) && (stride_o % 4 == 0) &&
                       !((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(out)) & 15);
  if (aligned && batch <= 65535) {
    dim3 grid(batch, (D + CW - 1) / CW);
    if (dtype == BASD_DTYPE_BF16) {
      const size_t dyn = ((size_t)2 * N + (size_t)(256 / (CW / 8)) * CW) * sizeof(float);
      weighted_center_vec_kernel<__nv_bfloat16, CW><<<grid, 256, dyn, ST>>>(
          (const __nv_bfloat16*)X, stride_x, W, stride_w, N, D, out, stride_o);
    } else {
      const size_t dyn = ((size_t)2 * N + (size_t)(256 / (CW / 4)) * CW) * sizeof(float);
      weighted_center_vec_kernel<float, CW><<<grid, 256, dyn, ST>>>((const float*)X, stride_x, W,
                                                                    stride_w, N, D, out, stride_o);
    }
    BASD_LAUNCH_CHECK();
    return 0;
  }
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

extern "C" int basd_procrustes_rows_finish(float* rows2, int rq, int ldr, long stride_r, float* Pt,
                                           int rp, int ldp, long stride_p, int batch,
                                           float rel_floor, float rel_floor_q, int eq, int ep,
                                           float* sig, float* nuc, float* pic, void* stream) {
  if (batch <= 0) return 0;
  if (eq < -1 || eq > 1 || ep < -1 || ep > 1) return -2;
  procrustes_rows_finish_kernel<<<batch, 512, (rq + 32) * sizeof(float), ST>>>(
      rows2, rq, ldr, stride_r, Pt, rp, ldp, stride_p, rel_floor, rel_floor_q, eq, ep, sig, nuc, pic);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_procrustes_grad_prep(float* YA, float* YB, const float* IA, const float* IB,
                                         int N, int rq, int ldi, long stride_i, int ldy,
                                         long stride_y, int batch, const float* pic,
                                         const float* nuc, const float* ks, const float* kt,
                                         const float* w, const float* totals, float* f_out,
                                         float* gw, int with_grad, void* stream) {
  if (batch <= 0) return 0;
  procrustes_grad_prep_kernel<<<batch, 512, 0, ST>>>(YA, YB, IA, IB, N, rq, ldi, stride_i, ldy,
                                                     stride_y, pic, nuc, ks, kt, w, totals, f_out,
                                                     gw, with_grad);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_procrustes_direct_grad(const float* A, float* T, const float* w, int N, int D,
                                           int batch, void* stream) {
  if (batch <= 0) return 0;
  const long per = (long)N * D;
  dim3 grid((unsigned)((per + 1023) / 1024), batch);
  procrustes_direct_grad_kernel<<<grid, 256, 0, ST>>>(A, T, w, N, D);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_scale_out(const float* src, void* dst, int dtype, long n, float alpha,
                              const float* alpha_dev, void* stream) {
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == BASD_DTYPE_BF16)
    scale_out_kernel<__nv_bfloat16><<<grid, 256, 0, ST>>>(src, (__nv_bfloat16*)dst, n, alpha, alpha_dev);
  else
    scale_out_kernel<float><<<grid, 256, 0, ST>>>(src, (float*)dst, n, alpha, alpha_dev);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_geo_reduce(const float* f, int E, int B, const float* weights, int n_weights,
                               float* geo_terms, float* geo, void* stream) {
  geo_reduce_kernel<<<1, 256, 0, ST>>>(f, E, B, weights, weights ? n_weights : 0, geo_terms, geo);
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

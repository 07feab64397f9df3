// Batched small-matrix factorisation kernels (one CTA per problem, matrix resident in
// shared memory when it fits, otherwise L2-resident global memory):
//
//   basd_pivoted_cholesky : K (PSD, n x n) -> rows of L^T, diagonal pivoting, rank-revealing
//   basd_jacobi_rows      : one-sided (Hestenes) Jacobi that orthogonalises the ROWS of a
//                           row-major matrix with warp-shuffle dot products and in-register
//                           plane rotations, round-robin (circle) pair schedule.
//   basd_rows_normalize   : row norms, optional descending sort, unit rows.
//   basd_rowdot           : per-row dot product of two matrices (Rayleigh refinement).
//
// These serve: the symmetric eigenproblems of the projected Gram matrices
// (reference: torch.linalg.eigvalsh / svd at layer_selector.py:16,36,92 -> Cholesky factor +
// row-Jacobi = Veselic-Hari), the k x k principal-angle SVDs (layer_selector.py:99) and the
// per-sample Procrustes SVD (relational.py:48, reduced to N x N).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>
namespace cg = cooperative_groups;

namespace basd {

// ------------------------------------------------------------------ pivoted Cholesky
// K is overwritten (used as the Schur complement). Output LT (n x n, row j = j-th column of
// L, rows >= rank are zero) so that K ~= LT^T LT.  `dims` (optional) gives the active
// leading dimension per problem; everything outside it is written as zero.
__global__ void pivoted_cholesky_kernel(float* __restrict__ Kbase, int n, int ld, long strideK,
                                        float* __restrict__ LTbase, int ldl, long strideL,
                                        float rel_tol, int* __restrict__ rank_out,
                                        const int* __restrict__ dims, int use_smem) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned int best_val[2][32];     // per-warp best pivot (float bits), double-buffered
  __shared__ int best_idx[2][32];
  const int prob = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  float* Kg = Kbase + (long)prob * strideK;
  float* LT = LTbase + (long)prob * strideL;
  const int nn = dims ? min(dims[prob], n) : n;
  const int npad = (n + 3) & ~3;
  float* colv = smem;                // npad (zero beyond nn)
  float* diag = colv + npad;         // npad: remaining pivot candidates, -1 once eliminated
  float* red = diag + npad;          // 32
  float* As = red + 32;              // npad*npad when staged
  float* A = use_smem ? As : Kg;
  const int lda = use_smem ? ((nn + 3) & ~3) : ld;
  if (use_smem) {
    for (int r = warp; r < nn; r += nwarp)
      for (int c = lane; c < lda; c += 32) As[r * lda + c] = (c < nn) ? Kg[(long)r * ld + c] : 0.f;
  }
  for (int i = tid; i < npad; i += T) colv[i] = 0.f;
  if (tid < 64) { best_val[tid >> 5][tid & 31] = 0u; best_idx[tid >> 5][tid & 31] = 0; }
  __syncthreads();
  // warp-level argmax of non-negative floats: REDUX on the bit patterns, then the first lane
  // that holds the maximum supplies the index
  auto warp_argmax = [](float v, int i, unsigned& vbits, int& idx) {
    const unsigned bits = __float_as_uint(fmaxf(v, 0.f));
    vbits = __reduce_max_sync(0xffffffffu, bits);
    const unsigned who = __ballot_sync(0xffffffffu, bits == vbits);
    idx = __shfl_sync(0xffffffffu, i, __ffs(who) - 1);
  };
  float dmax = 0.f;
  {
    float bv = -1.f;
    int bi = 0;
    for (int i = tid; i < nn; i += T) {
      const float d = A[(long)i * lda + i];
      diag[i] = d;
      dmax = fmaxf(dmax, d);
      if (d > bv) { bv = d; bi = i; }
    }
    unsigned vb;
    int ib;
    warp_argmax(bv, bi, vb, ib);
    if (lane == 0) { best_val[0][warp] = vb; best_idx[0][warp] = ib; }
  }
  dmax = block_max(dmax, red);
  const float floor_v = rel_tol * dmax;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < nn; ++j) {
    unsigned vb;
    int p;
    {   // every warp reduces the per-warp candidates redundantly: no extra barrier, no atomics
      const unsigned cv = lane < nwarp ? best_val[j & 1][lane] : 0u;
      const int ci = lane < nwarp ? best_idx[j & 1][lane] : 0;
      vb = __reduce_max_sync(0xffffffffu, cv);
      const unsigned who = __ballot_sync(0xffffffffu, cv == vb);
      p = __shfl_sync(0xffffffffu, ci, __ffs(who) - 1);
    }
    const float best = __uint_as_float(vb);
    if (!(best > floor_v) || !(best > 0.f)) break;           // uniform across the block
    const float rs = rsqrtf(best);
    for (int i = tid; i < nn; i += T) {                      // column j of L = row p of the Schur complement
      float c = (diag[i] < 0.f) ? 0.f : A[(long)p * lda + i] * rs;
      if (i == p) c = best * rs;
      colv[i] = c;
      LT[(long)j * ldl + i] = c;
    }
    __syncthreads();
    {   // rank-1 update (one warp per row, 128-bit chunks) fused with the next pivot search
      const int quads = lda >> 2;
      float bv = -1.f;
      int bi = 0;
      for (int i = warp; i < nn; i += nwarp) {
        const float di = diag[i];
        if (di < 0.f) continue;                              // already eliminated
        const float ci = colv[i];
        if (i == p) { if (lane == 0) diag[i] = -1.f; continue; }
        if (ci != 0.f) {
          float4* row = reinterpret_cast<float4*>(A + (long)i * lda);
          for (int k4 = lane; k4 < quads; k4 += 32) {
            const float4 ck = *reinterpret_cast<const float4*>(colv + 4 * k4);
            float4 a = row[k4];
            a.x = fmaf(-ci, ck.x, a.x); a.y = fmaf(-ci, ck.y, a.y);
            a.z = fmaf(-ci, ck.z, a.z); a.w = fmaf(-ci, ck.w, a.w);
            row[k4] = a;
          }
        }
        const float nd = fmaxf(fmaf(-ci, ci, di), 0.f);
        if (lane == 0) diag[i] = nd;
        if (nd > bv) { bv = nd; bi = i; }                    // warp-uniform
      }
      if (lane == 0) {
        best_val[(j + 1) & 1][warp] = __float_as_uint(fmaxf(bv, 0.f));
        best_idx[(j + 1) & 1][warp] = bi;
      }
    }
    __syncthreads();
    rank = j + 1;
  }
  __syncthreads();
  // zero the rest: rows >= rank, and (for dims) columns >= nn of every row
  for (int e = tid; e < n * n; e += T) {
    const int r = e / n, c = e - r * n;
    if (r >= rank || c >= nn) LT[(long)r * ldl + c] = 0.f;
  }
  if (rank_out && tid == 0) rank_out[prob] = rank;
}

// Left-looking variant for matrices whose factor fits shared memory (n <= 224): step j forms only
// the new column  col = (K[:,p] - sum_{k<j} L[:,k] L[p,k]) / sqrt(pivot)  from the factor rows
// already in shared memory.  The Schur complement is never materialised (K stays untouched in
// global memory; its pivot row is fetched from L2 while the dot products run), so a step reads
// j*n floats and writes n, instead of reading and writing the whole n x n trailing matrix.
template <int PARTS>
__global__ void __launch_bounds__(1024, 1)
pivoted_cholesky_left_kernel(const float* __restrict__ Kbase, int n, int ld, long strideK,
                             float* __restrict__ LTbase, int ldl, long strideL, float rel_tol,
                             int* __restrict__ rank_out, const int* __restrict__ dims) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned int best_val[2][32];
  __shared__ int best_idx[2][32];
  __shared__ float red[32];
  const int prob = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const float* Kg = Kbase + (long)prob * strideK;
  float* LT = LTbase + (long)prob * strideL;
  const int nn = dims ? min(dims[prob], n) : n;
  const int npad = (n + 31) & ~31;
  float* Ls = smem;                       // npad x npad factor rows (row j = column j of L)
  float* diag = Ls + (size_t)npad * npad; // npad
  float* pacc = diag + npad;              // PARTS x npad partial dot products
  const int wpp = npad >> 5;              // warps per part
  const int part = warp / wpp;            // which slice of k this thread sums
  const int i = (warp % wpp) * 32 + lane; // which row of the column it produces
  const bool worker = part < PARTS;
  auto warp_argmax = [](float v, int idx_in, unsigned& vbits, int& idx) {
    const unsigned bits = __float_as_uint(fmaxf(v, 0.f));
    vbits = __reduce_max_sync(0xffffffffu, bits);
    const unsigned who = __ballot_sync(0xffffffffu, bits == vbits);
    idx = __shfl_sync(0xffffffffu, idx_in, __ffs(who) - 1);
  };
  if (tid < 64) { best_val[tid >> 5][tid & 31] = 0u; best_idx[tid >> 5][tid & 31] = 0; }
  __syncthreads();
  float dmax = 0.f;
  {
    float bv = -1.f;
    int bi = 0;
    for (int r = tid; r < npad; r += T) {
      const float d = r < nn ? Kg[(long)r * ld + r] : -1.f;
      diag[r] = d;
      dmax = fmaxf(dmax, d);
      if (d > bv) { bv = d; bi = r; }
    }
    unsigned vb;
    int ib;
    warp_argmax(bv, bi, vb, ib);
    if (lane == 0) { best_val[0][warp] = vb; best_idx[0][warp] = ib; }
  }
  dmax = block_max(dmax, red);
  const float floor_v = rel_tol * dmax;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < nn; ++j) {
    unsigned vb;
    int p;
    {
      const unsigned cv = lane < nwarp ? best_val[j & 1][lane] : 0u;
      const int ci = lane < nwarp ? best_idx[j & 1][lane] : 0;
      vb = __reduce_max_sync(0xffffffffu, cv);
      const unsigned who = __ballot_sync(0xffffffffu, cv == vb);
      p = __shfl_sync(0xffffffffu, ci, __ffs(who) - 1);
    }
    const float best = __uint_as_float(vb);
    if (!(best > floor_v) || !(best > 0.f)) break;           // uniform across the block
    // pivot row of K (== pivot column by symmetry): in flight while the dot products run
    float kp = 0.f;
    if (worker && part == 0 && i < nn) kp = __ldg(Kg + (long)p * ld + i);
    if (worker) {
      float acc = 0.f;
      for (int k = part; k < j; k += PARTS)
        acc = fmaf(Ls[(size_t)k * npad + i], Ls[(size_t)k * npad + p], acc);
      pacc[part * npad + i] = acc;
    }
    __syncthreads();
    float bv = -1.f;
    int bi = 0;
    if (worker && part == 0) {
      float c = 0.f;
      const float di = diag[i];
      if (i < nn && di >= 0.f) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < PARTS; ++q) acc += pacc[q * npad + i];
        c = (i == p) ? best * rsqrtf(best) : (kp - acc) * rsqrtf(best);
        const float nd = (i == p) ? -1.f : fmaxf(fmaf(-c, c, di), 0.f);
        diag[i] = nd;
        bv = nd;
        bi = i;
      }
      Ls[(size_t)j * npad + i] = c;
      if (i < nn) LT[(long)j * ldl + i] = c;
    }
    {
      unsigned vbn;
      int ibn;
      warp_argmax(bv, bi, vbn, ibn);
      if (lane == 0) { best_val[(j + 1) & 1][warp] = vbn; best_idx[(j + 1) & 1][warp] = ibn; }
    }
    __syncthreads();
    rank = j + 1;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += T) {
    const int r = e / n, c = e - r * n;
    if (r >= rank || c >= nn) LT[(long)r * ldl + c] = 0.f;
  }
  if (rank_out && tid == 0) rank_out[prob] = rank;
}

// Left-looking, four outputs per thread: thread (part, iq) accumulates rows 4iq..4iq+3 of the new
// column over the factor rows k = part, part + PARTS, ... with ONE 128-bit load of L[k][4iq..] and
// one broadcast load of L[k][p] per four FMAs (the one-output kernel above issues two loads per
// FMA and ncu shows it bound by shared-memory instructions).  PARTS = warps / (npad / 128).
__global__ void __launch_bounds__(1024, 1)
pivoted_cholesky_left4_kernel(const float* __restrict__ Kbase, int n, int ld, long strideK,
                              float* __restrict__ LTbase, int ldl, long strideL, float rel_tol,
                              int* __restrict__ rank_out, const int* __restrict__ dims) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned int best_val[2][32];
  __shared__ int best_idx[2][32];
  __shared__ float red[32];
  const int prob = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const float* Kg = Kbase + (long)prob * strideK;
  float* LT = LTbase + (long)prob * strideL;
  const int nn = dims ? min(dims[prob], n) : n;
  const int npad = (n + 127) & ~127;      // multiple of 128: whole warps of row quads
  float* Ls = smem;                       // n x npad factor rows (row j = column j of L)
  float* diag = Ls + (size_t)n * npad;    // npad
  float* pacc = diag + npad;              // parts x npad partial dot products
  const int wpp = npad >> 7;              // warps per part (32 lanes x 4 rows = 128 rows per warp)
  const int parts = nwarp / wpp;
  const int part = warp / wpp;
  const int i0 = ((warp % wpp) * 32 + lane) * 4;   // first of the four rows this thread produces
  const bool worker = part < parts;
  auto warp_argmax = [](float v, int idx_in, unsigned& vbits, int& idx) {
    const unsigned bits = __float_as_uint(fmaxf(v, 0.f));
    vbits = __reduce_max_sync(0xffffffffu, bits);
    const unsigned who = __ballot_sync(0xffffffffu, bits == vbits);
    idx = __shfl_sync(0xffffffffu, idx_in, __ffs(who) - 1);
  };
  if (tid < 64) { best_val[tid >> 5][tid & 31] = 0u; best_idx[tid >> 5][tid & 31] = 0; }
  __syncthreads();
  float dmax = 0.f;
  {
    float bv = -1.f;
    int bi = 0;
    for (int r = tid; r < npad; r += T) {
      const float d = r < nn ? Kg[(long)r * ld + r] : -1.f;
      diag[r] = d;
      dmax = fmaxf(dmax, d);
      if (d > bv) { bv = d; bi = r; }
    }
    unsigned vb;
    int ib;
    warp_argmax(bv, bi, vb, ib);
    if (lane == 0) { best_val[0][warp] = vb; best_idx[0][warp] = ib; }
  }
  dmax = block_max(dmax, red);
  const float floor_v = rel_tol * dmax;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < nn; ++j) {
    unsigned vb;
    int p;
    {
      const unsigned cv = lane < nwarp ? best_val[j & 1][lane] : 0u;
      const int ci = lane < nwarp ? best_idx[j & 1][lane] : 0;
      vb = __reduce_max_sync(0xffffffffu, cv);
      const unsigned who = __ballot_sync(0xffffffffu, cv == vb);
      p = __shfl_sync(0xffffffffu, ci, __ffs(who) - 1);
    }
    const float best = __uint_as_float(vb);
    if (!(best > floor_v) || !(best > 0.f)) break;           // uniform across the block
    // pivot row of K (== pivot column by symmetry): in flight while the dot products run
    const bool finisher = tid < npad;                        // thread i finishes row i
    float kp = 0.f;
    if (finisher && tid < nn) kp = __ldg(Kg + (long)p * ld + tid);
    if (worker) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = part; k < j; k += parts) {
        const float4 l = *reinterpret_cast<const float4*>(Ls + (size_t)k * npad + i0);
        const float lp = Ls[(size_t)k * npad + p];
        acc.x = fmaf(l.x, lp, acc.x); acc.y = fmaf(l.y, lp, acc.y);
        acc.z = fmaf(l.z, lp, acc.z); acc.w = fmaf(l.w, lp, acc.w);
      }
      *reinterpret_cast<float4*>(pacc + (size_t)part * npad + i0) = acc;
    }
    __syncthreads();
    float bv = -1.f;
    int bi = 0;
    if (finisher) {
      const int i = tid;
      float c = 0.f;
      const float di = diag[i];
      if (i < nn && di >= 0.f) {
        float acc = 0.f;
        for (int q = 0; q < parts; ++q) acc += pacc[(size_t)q * npad + i];
        c = (i == p) ? best * rsqrtf(best) : (kp - acc) * rsqrtf(best);
        const float nd = (i == p) ? -1.f : fmaxf(fmaf(-c, c, di), 0.f);
        diag[i] = nd;
        bv = nd;
        bi = i;
      }
      Ls[(size_t)j * npad + i] = c;
      if (i < nn) LT[(long)j * ldl + i] = c;
    }
    {
      unsigned vbn;
      int ibn;
      warp_argmax(bv, bi, vbn, ibn);
      if (lane == 0) { best_val[(j + 1) & 1][warp] = vbn; best_idx[(j + 1) & 1][warp] = ibn; }
    }
    __syncthreads();
    rank = j + 1;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += T) {
    const int r = e / n, c = e - r * n;
    if (r >= rank || c >= nn) LT[(long)r * ldl + c] = 0.f;
  }
  if (rank_out && tid == 0) rank_out[prob] = rank;
}

// Left-looking, four outputs per thread, factor rows dealt round-robin to the CTAs of a cluster:
// CTA c keeps rows k = c, c + C, ... of the factor in its shared memory (n / C rows), forms its
// share of  sum_k L[k][:] L[k][p]  and publishes that partial column in shared memory; after ONE
// cluster barrier every CTA sums the C partial columns through DSMEM and finishes the column
// redundantly (same data, same order -> identical diag / pivot choice everywhere), the owner of row j
// stores it.  Serves Grams that do not fit one SM's shared memory (the 384 x 384 selector Grams:
// 16 problems x 4 CTAs instead of 16 CTAs streaming the trailing matrix through L2).
template <int CSIZE>
__global__ void __launch_bounds__(1024, 1)
pivoted_cholesky_left4_cluster_kernel(const float* __restrict__ Kbase, int n, int ld, long strideK,
                                      float* __restrict__ LTbase, int ldl, long strideL,
                                      float rel_tol, int* __restrict__ rank_out,
                                      const int* __restrict__ dims) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = cluster.block_rank();
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned int best_val[2][32];
  __shared__ int best_idx[2][32];
  __shared__ float red[32];
  const int prob = blockIdx.x / CSIZE, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const float* Kg = Kbase + (long)prob * strideK;
  float* LT = LTbase + (long)prob * strideL;
  const int nn = dims ? min(dims[prob], n) : n;
  const int npad = (n + 127) & ~127;
  const int rows_local = (n + CSIZE - 1) / CSIZE;
  float* Ls = smem;                                  // rows_local x npad: factor rows k = crank + CSIZE * kk
  float* diag = Ls + (size_t)rows_local * npad;      // npad (replicated in every CTA)
  float* cpart = diag + npad;                        // 2 x npad: this CTA's partial column, double buffered
  float* pacc = cpart + 2 * npad;                    // parts x npad
  const int wpp = npad >> 7;
  const int parts = nwarp / wpp;
  const int part = warp / wpp;
  const int i0 = ((warp % wpp) * 32 + lane) * 4;
  const bool worker = part < parts;
  const float* remote[CSIZE];
#pragma unroll
  for (int c = 0; c < CSIZE; ++c) remote[c] = cluster.map_shared_rank(cpart, c);
  auto warp_argmax = [](float v, int idx_in, unsigned& vbits, int& idx) {
    const unsigned bits = __float_as_uint(fmaxf(v, 0.f));
    vbits = __reduce_max_sync(0xffffffffu, bits);
    const unsigned who = __ballot_sync(0xffffffffu, bits == vbits);
    idx = __shfl_sync(0xffffffffu, idx_in, __ffs(who) - 1);
  };
  if (tid < 64) { best_val[tid >> 5][tid & 31] = 0u; best_idx[tid >> 5][tid & 31] = 0; }
  __syncthreads();
  float dmax = 0.f;
  {
    float bv = -1.f;
    int bi = 0;
    for (int r = tid; r < npad; r += T) {
      const float d = r < nn ? Kg[(long)r * ld + r] : -1.f;
      diag[r] = d;
      dmax = fmaxf(dmax, d);
      if (d > bv) { bv = d; bi = r; }
    }
    unsigned vb;
    int ib;
    warp_argmax(bv, bi, vb, ib);
    if (lane == 0) { best_val[0][warp] = vb; best_idx[0][warp] = ib; }
  }
  dmax = block_max(dmax, red);
  const float floor_v = rel_tol * dmax;
  __syncthreads();
  int rank = 0;
  for (int j = 0; j < nn; ++j) {
    unsigned vb;
    int p;
    {
      const unsigned cv = lane < nwarp ? best_val[j & 1][lane] : 0u;
      const int ci = lane < nwarp ? best_idx[j & 1][lane] : 0;
      vb = __reduce_max_sync(0xffffffffu, cv);
      const unsigned who = __ballot_sync(0xffffffffu, cv == vb);
      p = __shfl_sync(0xffffffffu, ci, __ffs(who) - 1);
    }
    const float best = __uint_as_float(vb);
    if (!(best > floor_v) || !(best > 0.f)) break;           // identical in every CTA of the cluster
    const bool finisher = tid < npad;
    float kp = 0.f;
    if (finisher && tid < nn) kp = __ldg(Kg + (long)p * ld + tid);
    const int jl = (j - crank + CSIZE - 1) / CSIZE;          // local rows with global index < j
    if (worker) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kk = part; kk < jl; kk += parts) {
        const float* row = Ls + (size_t)kk * npad;
        const float4 l = *reinterpret_cast<const float4*>(row + i0);
        const float lp = row[p];
        acc.x = fmaf(l.x, lp, acc.x); acc.y = fmaf(l.y, lp, acc.y);
        acc.z = fmaf(l.z, lp, acc.z); acc.w = fmaf(l.w, lp, acc.w);
      }
      *reinterpret_cast<float4*>(pacc + (size_t)part * npad + i0) = acc;
    }
    __syncthreads();
    float* mine = cpart + (size_t)(j & 1) * npad;
    if (finisher) {
      float acc = 0.f;
      for (int q = 0; q < parts; ++q) acc += pacc[(size_t)q * npad + tid];
      mine[tid] = acc;
    }
    cluster.sync();                                          // partial columns visible cluster-wide
    float bv = -1.f;
    int bi = 0;
    if (finisher) {
      const int i = tid;
      float c = 0.f;
      const float di = diag[i];
      if (i < nn && di >= 0.f) {
        float acc = 0.f;
#pragma unroll
        for (int cc = 0; cc < CSIZE; ++cc) acc += remote[cc][(size_t)(j & 1) * npad + i];
        c = (i == p) ? best * rsqrtf(best) : (kp - acc) * rsqrtf(best);
        const float nd = (i == p) ? -1.f : fmaxf(fmaf(-c, c, di), 0.f);
        diag[i] = nd;
        bv = nd;
        bi = i;
      }
      if (j % CSIZE == crank) {                              // owner of factor row j
        Ls[(size_t)(j / CSIZE) * npad + i] = c;
        if (i < nn) LT[(long)j * ldl + i] = c;
      }
    }
    {
      unsigned vbn;
      int ibn;
      warp_argmax(bv, bi, vbn, ibn);
      if (lane == 0) { best_val[(j + 1) & 1][warp] = vbn; best_idx[(j + 1) & 1][warp] = ibn; }
    }
    __syncthreads();
    rank = j + 1;
  }
  cluster.sync();                                            // nobody leaves while a peer may still read its partials
  if (crank == 0) {
    for (int e = tid; e < n * n; e += T) {
      const int r = e / n, c = e - r * n;
      if (r >= rank || c >= nn) LT[(long)r * ldl + c] = 0.f;
    }
    if (rank_out && tid == 0) rank_out[prob] = rank;
  }
}

// ------------------------------------------------------------------ one-sided Jacobi on rows
template <int NV>
__global__ void __launch_bounds__((NV >= 3 ? 512 : 1024), 1)
jacobi_rows_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                   const int* __restrict__ dims, float tol, int max_sweeps, int use_smem,
                   int* __restrict__ sweeps_out) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int rot_count;
  const int prob = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  float* Gg = Gbase + (long)prob * stride;
  const int nn = dims ? min(dims[prob], n) : n;
  const int mm = dims ? min(dims[prob], m) : m;
  const int mv = (mm + 3) >> 2;            // float4 per row
  float* G = Gg;
  int ldw = ld;
  if (use_smem) {
    ldw = mv * 4;
    G = smem;
    for (int e = tid; e < nn * mv; e += T) {
      const int r = e / mv, c4 = (e - r * mv) * 4;
      float4 v = *reinterpret_cast<const float4*>(Gg + (long)r * ld + c4);
      if (c4 + 1 >= mm) v.y = 0.f;
      if (c4 + 2 >= mm) v.z = 0.f;
      if (c4 + 3 >= mm) v.w = 0.f;
      *reinterpret_cast<float4*>(G + (long)r * ldw + c4) = v;
    }
  }
  __syncthreads();
  const int ne = nn + (nn & 1);
  const int ring = ne - 1;
  int sweep = 0;
  for (; sweep < max_sweeps && nn >= 2; ++sweep) {
    if (tid == 0) rot_count = 0;
    __syncthreads();
    for (int step = 0; step < ring; ++step) {
      for (int pair = warp; pair < (ne >> 1); pair += nwarps) {
        int p, q;
        if (pair == 0) { p = ne - 1; q = step; }
        else { p = (step + pair) % ring; q = (step - pair + ring) % ring; }
        if (p >= nn || q >= nn) continue;   // bye (odd n)
        float4* rp = reinterpret_cast<float4*>(G + (long)p * ldw);
        float4* rq = reinterpret_cast<float4*>(G + (long)q * ldw);
        float4 x[NV], y[NV];
        float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int idx = lane + 32 * v;
          if (idx < mv) {
            x[v] = rp[idx];
            y[v] = rq[idx];
            if (!use_smem) {   // global rows may carry junk past mm only in the last quad
              const int c4 = idx * 4;
              if (c4 + 1 >= mm) { x[v].y = 0.f; y[v].y = 0.f; }
              if (c4 + 2 >= mm) { x[v].z = 0.f; y[v].z = 0.f; }
              if (c4 + 3 >= mm) { x[v].w = 0.f; y[v].w = 0.f; }
            }
          } else {
            x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            y[v] = x[v];
          }
          al = fmaf(x[v].x, x[v].x, al); al = fmaf(x[v].y, x[v].y, al);
          al = fmaf(x[v].z, x[v].z, al); al = fmaf(x[v].w, x[v].w, al);
          be = fmaf(y[v].x, y[v].x, be); be = fmaf(y[v].y, y[v].y, be);
          be = fmaf(y[v].z, y[v].z, be); be = fmaf(y[v].w, y[v].w, be);
          ga = fmaf(x[v].x, y[v].x, ga); ga = fmaf(x[v].y, y[v].y, ga);
          ga = fmaf(x[v].z, y[v].z, ga); ga = fmaf(x[v].w, y[v].w, ga);
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_sum(ga);
        const float lim = tol * sqrtf(al) * sqrtf(be);
        if (!(fabsf(ga) > lim) || !(lim > 0.f)) continue;   // warp-uniform
        const float zeta = (be - al) / (2.f * ga);
        float t;
        if (fabsf(zeta) > 1e8f) t = 0.5f / zeta;
        else t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
        const float c = rsqrtf(fmaf(t, t, 1.f)), s = c * t;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int idx = lane + 32 * v;
          if (idx < mv) {
            float4 a = x[v], b = y[v], na, nb;
            na.x = c * a.x - s * b.x; nb.x = s * a.x + c * b.x;
            na.y = c * a.y - s * b.y; nb.y = s * a.y + c * b.y;
            na.z = c * a.z - s * b.z; nb.z = s * a.z + c * b.z;
            na.w = c * a.w - s * b.w; nb.w = s * a.w + c * b.w;
            rp[idx] = na;
            rq[idx] = nb;
          }
        }
        if (lane == 0) atomicAdd(&rot_count, 1);
      }
      __syncthreads();
    }
    const int done = (rot_count == 0);
    __syncthreads();
    if (done) { ++sweep; break; }
  }
  if (use_smem) {
    for (int e = tid; e < nn * mv; e += T) {
      const int r = e / mv, c4 = (e - r * mv) * 4;
      *reinterpret_cast<float4*>(Gg + (long)r * ld + c4) =
          *reinterpret_cast<const float4*>(G + (long)r * ldw + c4);
    }
  }
  if (sweeps_out && tid == 0) sweeps_out[prob] = sweep;
}

// Shared-memory variant with LP lanes per row pair (LP = 8 or 16): 32/LP pairs per warp, so a
// whole round-robin step of an N<=224 problem runs in ONE round of the CTA instead of four.
// Squared row norms are cached in shared memory and updated analytically after each rotation
// (a_pp' = a_pp - t*g, a_qq' = a_qq + t*g), refreshed at the start of every sweep, so a pair
// costs one dot product instead of three; the rotation parameters use one fast divide and
// one rsqrt (+ a Newton step on c so that c^2 + s^2 = 1 to an ulp).
template <int LP, int NV, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
jacobi_rows_grouped_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                           const int* __restrict__ dims, float tol, int max_sweeps,
                           int* __restrict__ sweeps_out) {
  extern __shared__ __align__(16) float smem[];
  const int prob = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int gid = tid / LP, gl = tid % LP, groups = T / LP;
  float* Gg = Gbase + (long)prob * stride;
  const int nn = dims ? min(dims[prob], n) : n;
  const int mm = dims ? min(dims[prob], m) : m;
  const int mv = (mm + 3) >> 2;
  const int ldw = mv * 4;
  __shared__ float red_scratch[32];
  float* G = smem;
  float* nrm2 = smem + (size_t)n * (((size_t)m + 3) & ~(size_t)3);
  for (int e = tid; e < nn * mv; e += T) {
    const int r = e / mv, c4 = (e - r * mv) * 4;
    float4 v = *reinterpret_cast<const float4*>(Gg + (long)r * ld + c4);
    if (c4 + 1 >= mm) v.y = 0.f;
    if (c4 + 2 >= mm) v.z = 0.f;
    if (c4 + 3 >= mm) v.w = 0.f;
    *reinterpret_cast<float4*>(G + (long)r * ldw + c4) = v;
  }
  __syncthreads();
  const int ne = nn + (nn & 1), ring = ne - 1, half = ne >> 1;
  const float tol2 = tol * tol;
  int sweep = 0;
  for (; sweep < max_sweeps && nn >= 2; ++sweep) {
    for (int base = 0; base < nn; base += groups) {        // refresh the cached norms
      const int r = base + gid;
      float a = 0.f;
      if (r < nn) {
        const float4* row = reinterpret_cast<const float4*>(G + (long)r * ldw);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int idx = gl + LP * v;
          if (idx < mv) {
            const float4 x = row[idx];
            a = fmaf(x.x, x.x, a); a = fmaf(x.y, x.y, a); a = fmaf(x.z, x.z, a); a = fmaf(x.w, x.w, a);
          }
        }
      }
#pragma unroll
      for (int o = LP >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (r < nn && gl == 0) nrm2[r] = a;
    }
    __syncthreads();
    // rows more than 1e-7 below the largest row are numerically zero: never rotate them
    float mx = 0.f;
    for (int r = tid; r < nn; r += T) mx = fmaxf(mx, nrm2[r]);
    mx = block_max(mx, red_scratch);
    const float zero_thr = 1e-14f * mx;
    float worst = 0.f;        // largest cos^2 between two rows met in this sweep (before rotating)
    for (int step = 0; step < ring; ++step) {
      for (int base = 0; base < half; base += groups) {
        const int pair = base + gid;
        int p = 0, q = 0;
        bool valid = pair < half;
        if (valid) {
          if (pair == 0) { p = ne - 1; q = step; }
          else { p = step + pair; if (p >= ring) p -= ring; q = step - pair; if (q < 0) q += ring; }
          valid = (p < nn) && (q < nn);
        }
        float4* rp = reinterpret_cast<float4*>(G + (long)p * ldw);
        float4* rq = reinterpret_cast<float4*>(G + (long)q * ldw);
        float4 x[NV], y[NV];
        float ga = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int idx = gl + LP * v;
          if (valid && idx < mv) {
            x[v] = rp[idx];
            y[v] = rq[idx];
          } else {
            x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            y[v] = x[v];
          }
          ga = fmaf(x[v].x, y[v].x, ga); ga = fmaf(x[v].y, y[v].y, ga);
          ga = fmaf(x[v].z, y[v].z, ga); ga = fmaf(x[v].w, y[v].w, ga);
        }
#pragma unroll
        for (int o = LP >> 1; o > 0; o >>= 1) ga += __shfl_xor_sync(0xffffffffu, ga, o);
        if (!valid) continue;
        const float al = nrm2[p], be = nrm2[q];
        if (!(ga * ga > tol2 * al * be) || al <= zero_thr || be <= zero_thr) continue;   // group-uniform
        worst = fmaxf(worst, __fdividef(ga * ga, al * be));
        // t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (be - al) / (2 ga), rewritten as
        // t = sgn * 2|ga| / (|d| + sqrt(d^2 + 4 ga^2)) with d = be - al: one rsqrt, one divide.
        const float d = be - al;
        const float h = fmaf(d, d, 4.f * ga * ga);
        const float root = h * rsqrtf(h);
        float t = __fdividef(2.f * fabsf(ga), fabsf(d) + root);
        t = ((d < 0.f) != (ga < 0.f)) ? -t : t;
        const float w2 = fmaf(t, t, 1.f);
        float c = rsqrtf(w2);
        c = c * fmaf(-0.5f * w2, c * c, 1.5f);               // Newton: c^2 (1 + t^2) = 1 to an ulp
        const float sn = c * t;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int idx = gl + LP * v;
          if (idx < mv) {
            const float4 a = x[v], b = y[v];
            float4 na, nb;
            na.x = fmaf(c, a.x, -sn * b.x); nb.x = fmaf(sn, a.x, c * b.x);
            na.y = fmaf(c, a.y, -sn * b.y); nb.y = fmaf(sn, a.y, c * b.y);
            na.z = fmaf(c, a.z, -sn * b.z); nb.z = fmaf(sn, a.z, c * b.z);
            na.w = fmaf(c, a.w, -sn * b.w); nb.w = fmaf(sn, a.w, c * b.w);
            rp[idx] = na;
            rq[idx] = nb;
          }
        }
        if (gl == 0) {
          nrm2[p] = fmaxf(fmaf(-t, ga, al), 0.f);
          nrm2[q] = fmaxf(fmaf(t, ga, be), 0.f);
        }
      }
      __syncthreads();
    }
    // Jacobi converges quadratically: once every |cos| met in a sweep is below sqrt(tol), the
    // rows are orthogonal to ~tol after it -- no separate verification sweep is needed.
    worst = block_max(worst, red_scratch);
    if (worst < tol) { ++sweep; break; }
    __syncthreads();
  }
  for (int e = tid; e < nn * mv; e += T) {
    const int r = e / mv, c4 = (e - r * mv) * 4;
    *reinterpret_cast<float4*>(Gg + (long)r * ld + c4) =
        *reinterpret_cast<const float4*>(G + (long)r * ldw + c4);
  }
  if (sweeps_out && tid == 0) sweeps_out[prob] = sweep;
}

// Register-resident variant (odd-even transposition ordering).  Each group of LP lanes OWNS two
// rows in registers (positions 2g and 2g+1 of a line).  Even steps rotate the resident pair and
// touch no shared memory at all; odd steps rotate (2g+1, 2g+2): the left row of every group is
// parked in a shared-memory exchange buffer, the left neighbour pairs it with its own right row
// and hands the result back.  After every rotation the two rows trade places, so in n steps every
// pair has met exactly once (the odd-even transposition network).  Shared-memory traffic per pair
// is half that of the round-robin kernel, which ncu showed to be shared-memory-bandwidth bound.
// All reloads are unconditional so at most two rows are ever live in registers.
// Rows are held as  row = d * stored  with a per-row scale d ("fast" / scaled rotations): the
// plane rotation  x' = c x - s y, y' = s x + c y  becomes two FMAs per element pair on the stored
// values,  stored_y' = stored_y + (t dx/dy) stored_x,  stored_x' = stored_x - (t dy/dx) stored_y,
// with the cosine folded into the scales (d' = c d_partner) instead of two FMUL + two FFMA.
// Scales shrink by c >= 1/sqrt(2) per rotation and are folded back into the rows every 16 steps.
template <int NV>
__device__ __forceinline__ void fold_scale(float4 (&r)[NV], float& d) {
#pragma unroll
  for (int v = 0; v < NV; ++v) { r[v].x *= d; r[v].y *= d; r[v].z *= d; r[v].w *= d; }
  d = 1.f;
}

template <int LP, int NV>
__device__ __forceinline__ void rotate_and_swap(float4 (&x)[NV], float& nx, float& dx,
                                                float4 (&y)[NV], float& ny, float& dy, bool valid,
                                                float tol2, float zero_thr, float& worst, int& nrot) {
  float ga = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    ga = fmaf(x[v].x, y[v].x, ga); ga = fmaf(x[v].y, y[v].y, ga);
    ga = fmaf(x[v].z, y[v].z, ga); ga = fmaf(x[v].w, y[v].w, ga);
  }
#pragma unroll
  for (int o = LP >> 1; o > 0; o >>= 1) ga += __shfl_xor_sync(0xffffffffu, ga, o);
  if (!valid) return;
  ga *= dx * dy;
  const bool rot = (ga * ga > tol2 * nx * ny) && nx > zero_thr && ny > zero_thr;   // group-uniform
  float c = 1.f, t1 = 0.f, t2 = 0.f;
  if (rot) {
    ++nrot;
    worst = fmaxf(worst, __fdividef(ga * ga, nx * ny));
    const float d = ny - nx;
    const float h = fmaf(d, d, 4.f * ga * ga);
    const float root = h * rsqrtf(h);
    float t = __fdividef(2.f * fabsf(ga), fabsf(d) + root);
    t = ((d < 0.f) != (ga < 0.f)) ? -t : t;
    const float w2 = fmaf(t, t, 1.f);
    c = rsqrtf(w2);
    c = c * fmaf(-0.5f * w2, c * c, 1.5f);
    t1 = t * __fdividef(dx, dy);
    t2 = t * __fdividef(dy, dx);
    // New squared norms (the rows trade places below, so they swap as well).  The larger row
    // grows by |t g| (no cancellation); the analytic update of the smaller one, a - |t g|,
    // cancels when the rotation nearly annihilates it (graded matrices: a row 1e-4 of its partner
    // would keep a norm of pure rounding noise and every later angle it takes part in would be
    // wrong).  The 2x2 Gram determinant a_pp a_qq - g^2 is invariant under the rotation, so the
    // smaller norm is det / larger -- accurate unless the rows are parallel.
    const float tg = t * ga;                               // sign(tg) = sign(d)
    const float big = (d >= 0.f) ? ny + tg : nx - tg;
    const float r = __fdividef(1.f, big);
    const float small = fmaxf(fmaf(nx, ny * r, -(ga * r) * ga), 0.f);
    nx = (d >= 0.f) ? big : small;                         // x will hold y'
    ny = (d >= 0.f) ? small : big;                         // y will hold x'
  } else {
    const float tmp = nx;
    nx = ny;
    ny = tmp;
  }
  // y' = s x + c y = (c dy)(stored_y + t1 stored_x),  x' = c x - s y = (c dx)(stored_x - t2 stored_y);
  // then the rows trade places: x <- y', y <- x'
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 a = x[v], b = y[v];
    x[v].x = fmaf(t1, a.x, b.x); y[v].x = fmaf(-t2, b.x, a.x);
    x[v].y = fmaf(t1, a.y, b.y); y[v].y = fmaf(-t2, b.y, a.y);
    x[v].z = fmaf(t1, a.z, b.z); y[v].z = fmaf(-t2, b.z, a.z);
    x[v].w = fmaf(t1, a.w, b.w); y[v].w = fmaf(-t2, b.w, a.w);
  }
  const float ndx = c * dy;
  dy = c * dx;
  dx = ndx;
}

template <int LP, int NV, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
jacobi_rows_oddeven_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                           const int* __restrict__ dims, float tol, int max_sweeps,
                           int* __restrict__ sweeps_out, int dim_lo, int dim_hi,
                           int* __restrict__ rot_out) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_scratch[32];
  int nrot = 0;
  const int prob = blockIdx.x, tid = threadIdx.x;
  // size window: lets the host launch this kernel and the cluster kernel back to back on the
  // same batch, each taking the problems whose (device-resident) size suits it -- no host sync
  if (dims && (dims[prob] < dim_lo || dims[prob] > dim_hi)) return;
  const int gid = tid / LP, gl = tid % LP;
  float* Gg = Gbase + (long)prob * stride;
  const int nn = dims ? min(dims[prob], n) : n;
  const int mm = dims ? min(dims[prob], m) : m;
  const int mv = (mm + 3) >> 2;
  const int ldw = (((m + 3) >> 2) << 2);                 // exchange-buffer row pitch (>= 4*NV*LP? no: padded below)
  float* xbuf = smem;                                    // (cap/2 + 1) rows of LP*NV quads
  const int pitch = LP * NV * 4;                         // every lane may touch all its NV quads
  const int cap = dims ? min(n, dim_hi) : n;             // largest problem this launch accepts
  float2* xn = reinterpret_cast<float2*>(smem + (size_t)((cap + 1) / 2 + 1) * pitch);  // (squared norm, scale) of the parked rows
  (void)ldw;
  const int h = (nn + 1) >> 1;                           // groups; the last one has no right row if nn is odd
  const bool active = gid < h;
  const int row_a = 2 * gid, row_b = 2 * gid + 1;
  const bool has_b = active && row_b < nn;
  float4 a[NV], b[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int idx = gl + LP * v;
    a[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    b[v] = a[v];
    if (active && idx < mv) {
      const int c4 = idx * 4;
      a[v] = *reinterpret_cast<const float4*>(Gg + (long)row_a * ld + c4);
      if (has_b) b[v] = *reinterpret_cast<const float4*>(Gg + (long)row_b * ld + c4);
      if (c4 + 1 >= mm) { a[v].y = 0.f; b[v].y = 0.f; }
      if (c4 + 2 >= mm) { a[v].z = 0.f; b[v].z = 0.f; }
      if (c4 + 3 >= mm) { a[v].w = 0.f; b[v].w = 0.f; }
    }
  }
  const float tol2 = tol * tol;
  const int spare = (cap + 1) / 2;
  const int slot = min(gid, spare);                      // parking slot (inactive groups share the spare one)
  float4* my_slot = reinterpret_cast<float4*>(xbuf + (size_t)slot * pitch);
  float4* right_slot = reinterpret_cast<float4*>(xbuf + (size_t)min(gid + 1, spare) * pitch);
  int sweep = 0;
  float da = 1.f, db = 1.f;                               // row = scale * stored (fast rotations)
  for (; sweep < max_sweeps && nn >= 2; ++sweep) {
    fold_scale<NV>(a, da);
    fold_scale<NV>(b, db);
    float na = 0.f, nb = 0.f;                             // refresh the carried norms
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      na = fmaf(a[v].x, a[v].x, na); na = fmaf(a[v].y, a[v].y, na);
      na = fmaf(a[v].z, a[v].z, na); na = fmaf(a[v].w, a[v].w, na);
      nb = fmaf(b[v].x, b[v].x, nb); nb = fmaf(b[v].y, b[v].y, nb);
      nb = fmaf(b[v].z, b[v].z, nb); nb = fmaf(b[v].w, b[v].w, nb);
    }
#pragma unroll
    for (int o = LP >> 1; o > 0; o >>= 1) {
      na += __shfl_xor_sync(0xffffffffu, na, o);
      nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    const float mx = block_max(fmaxf(na, nb), red_scratch);
    const float zero_thr = 1e-14f * mx;
    float worst = 0.f;
    for (int step = 0; step < nn; ++step) {
      if ((step & 1) == 0) {
        if ((step & 15) == 0 && step) { fold_scale<NV>(a, da); fold_scale<NV>(b, db); }
        rotate_and_swap<LP, NV>(a, na, da, b, nb, db, has_b, tol2, zero_thr, worst, nrot);
      } else {
        // park the left row of every group; the left neighbour pairs it with its right row
#pragma unroll
        for (int v = 0; v < NV; ++v) my_slot[gl + LP * v] = a[v];
        if (gl == 0) xn[slot] = make_float2(na, da);
        __syncthreads();
        const bool pair_ok = has_b && (gid + 1 < h);
        float4 y[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) y[v] = right_slot[gl + LP * v];
        const float2 nd = xn[min(gid + 1, spare)];
        float ny = nd.x, dy = nd.y;
        rotate_and_swap<LP, NV>(b, nb, db, y, ny, dy, pair_ok, tol2, zero_thr, worst, nrot);
        if (pair_ok) {
#pragma unroll
          for (int v = 0; v < NV; ++v) right_slot[gl + LP * v] = y[v];
          if (gl == 0) xn[gid + 1] = make_float2(ny, dy);
        }
        __syncthreads();
#pragma unroll
        for (int v = 0; v < NV; ++v) a[v] = my_slot[gl + LP * v];
        const float2 mine = xn[slot];
        na = mine.x;
        da = mine.y;
      }
    }
    worst = block_max(worst, red_scratch);
    if (worst < tol) { ++sweep; break; }
  }
  fold_scale<NV>(a, da);
  fold_scale<NV>(b, db);
  if (active) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int idx = gl + LP * v;
      if (idx < mv) {
        *reinterpret_cast<float4*>(Gg + (long)row_a * ld + idx * 4) = a[v];
        if (has_b) *reinterpret_cast<float4*>(Gg + (long)row_b * ld + idx * 4) = b[v];
      }
    }
  }
  if (sweeps_out && tid == 0) sweeps_out[prob] = sweep;
  if (rot_out) {                                         // rotations applied (one count per group)
    const float tot = block_sum(gl == 0 ? (float)nrot : 0.f, red_scratch);
    if (tid == 0) atomicAdd(rot_out + prob, (int)tot);
  }
}

// Cluster-wide register-resident odd-even Jacobi: the rows of ONE problem are spread over the
// CTAs of a thread-block cluster (positions [crank*2*gpc, (crank+1)*2*gpc) live in CTA crank's
// registers).  Even steps are purely register-local.  In odd steps every group parks its left row
// in its CTA's shared memory; the only cross-CTA traffic is the single boundary row per CTA,
// read and written back through distributed shared memory.  Two cluster barriers per odd step
// replace the per-step cluster barrier + L2 round trips of jacobi_rows_cluster_kernel.
__device__ __forceinline__ float cluster_max_nonneg(float v, cg::cluster_group& cluster, int* flag0,
                                                    float* red) {
  v = block_max(v, red);
  if (threadIdx.x == 0) atomicMax(flag0, __float_as_int(v));
  cluster.sync();
  const float all = __int_as_float(*flag0);
  return all;
}

template <int LP, int NV, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
jacobi_rows_oe_cluster_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                              const int* __restrict__ dims, float tol, int max_sweeps,
                              int* __restrict__ sweeps_out, int dim_lo, int dim_hi,
                              int* __restrict__ rot_out) {
  int nrot = 0;
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = cluster.num_blocks(), crank = cluster.block_rank();
  const int prob = blockIdx.x / csize;
  if (dims && (dims[prob] < dim_lo || dims[prob] > dim_hi)) return;     // whole cluster exits
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_scratch[32];
  __shared__ int cflag[2];
  const int tid = threadIdx.x;
  const int gpc = blockDim.x / LP;                        // groups (row pairs) per CTA
  const int lgid = tid / LP, gl = tid % LP;
  const int gid = crank * gpc + lgid;                     // global group index
  float* Gg = Gbase + (long)prob * stride;
  const int nn = dims ? min(dims[prob], n) : n;
  const int mm = dims ? min(dims[prob], m) : m;
  const int mv = (mm + 3) >> 2;
  constexpr int pitch = LP * NV * 4;
  float* xbuf = smem;                                     // gpc + 1 parking slots
  float2* xn = reinterpret_cast<float2*>(smem + (size_t)(gpc + 1) * pitch);   // (squared norm, scale) of the parked rows
  const int h = (nn + 1) >> 1;
  const bool active = gid < h;
  const int row_a = 2 * gid, row_b = 2 * gid + 1;
  const bool has_b = active && row_b < nn;
  float4 a[NV], b[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int idx = gl + LP * v;
    a[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    b[v] = a[v];
    if (active && idx < mv) {
      const int c4 = idx * 4;
      a[v] = *reinterpret_cast<const float4*>(Gg + (long)row_a * ld + c4);
      if (has_b) b[v] = *reinterpret_cast<const float4*>(Gg + (long)row_b * ld + c4);
      if (c4 + 1 >= mm) { a[v].y = 0.f; b[v].y = 0.f; }
      if (c4 + 2 >= mm) { a[v].z = 0.f; b[v].z = 0.f; }
      if (c4 + 3 >= mm) { a[v].w = 0.f; b[v].w = 0.f; }
    }
  }
  const float tol2 = tol * tol;
  float4* my_slot = reinterpret_cast<float4*>(xbuf + (size_t)lgid * pitch);
  float2* my_norm = xn + lgid;
  // right neighbour's parking slot: next group of this CTA, or slot 0 of the next CTA (DSMEM)
  float4* right_slot;
  float2* right_norm;
  if (lgid + 1 < gpc) {
    right_slot = reinterpret_cast<float4*>(xbuf + (size_t)(lgid + 1) * pitch);
    right_norm = xn + lgid + 1;
  } else if (crank + 1 < csize) {
    right_slot = reinterpret_cast<float4*>(cluster.map_shared_rank(xbuf, crank + 1));
    right_norm = cluster.map_shared_rank(xn, crank + 1);
  } else {
    right_slot = reinterpret_cast<float4*>(xbuf + (size_t)gpc * pitch);   // spare (never paired)
    right_norm = xn + gpc;
  }
  int* flag0 = cluster.map_shared_rank(cflag, 0);
  int sweep = 0;
  float da = 1.f, db = 1.f;                               // row = scale * stored (fast rotations)
  for (; sweep < max_sweeps && nn >= 2; ++sweep) {
    if (crank == 0 && tid == 0) { cflag[0] = 0; cflag[1] = 0; }
    cluster.sync();
    fold_scale<NV>(a, da);
    fold_scale<NV>(b, db);
    float na = 0.f, nb = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      na = fmaf(a[v].x, a[v].x, na); na = fmaf(a[v].y, a[v].y, na);
      na = fmaf(a[v].z, a[v].z, na); na = fmaf(a[v].w, a[v].w, na);
      nb = fmaf(b[v].x, b[v].x, nb); nb = fmaf(b[v].y, b[v].y, nb);
      nb = fmaf(b[v].z, b[v].z, nb); nb = fmaf(b[v].w, b[v].w, nb);
    }
#pragma unroll
    for (int o = LP >> 1; o > 0; o >>= 1) {
      na += __shfl_xor_sync(0xffffffffu, na, o);
      nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    const float mx = cluster_max_nonneg(fmaxf(na, nb), cluster, flag0, red_scratch);
    const float zero_thr = 1e-14f * mx;
    float worst = 0.f;
    for (int step = 0; step < nn; ++step) {
      if ((step & 1) == 0) {
        if ((step & 15) == 0 && step) { fold_scale<NV>(a, da); fold_scale<NV>(b, db); }
        rotate_and_swap<LP, NV>(a, na, da, b, nb, db, has_b, tol2, zero_thr, worst, nrot);
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) my_slot[gl + LP * v] = a[v];
        if (gl == 0) *my_norm = make_float2(na, da);
        cluster.sync();
        const bool pair_ok = has_b && (gid + 1 < h);
        float4 y[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) y[v] = right_slot[gl + LP * v];
        const float2 nd = *right_norm;
        float ny = nd.x, dy = nd.y;
        rotate_and_swap<LP, NV>(b, nb, db, y, ny, dy, pair_ok, tol2, zero_thr, worst, nrot);
        if (pair_ok) {
#pragma unroll
          for (int v = 0; v < NV; ++v) right_slot[gl + LP * v] = y[v];
          if (gl == 0) *right_norm = make_float2(ny, dy);
        }
        cluster.sync();
#pragma unroll
        for (int v = 0; v < NV; ++v) a[v] = my_slot[gl + LP * v];
        const float2 mine = *my_norm;
        na = mine.x;
        da = mine.y;
      }
    }
    const float all_worst = cluster_max_nonneg(worst, cluster, flag0 + 1, red_scratch);
    cluster.sync();                                       // all have read before rank 0 resets
    if (all_worst < tol) { ++sweep; break; }
  }
  fold_scale<NV>(a, da);
  fold_scale<NV>(b, db);
  if (active) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int idx = gl + LP * v;
      if (idx < mv) {
        *reinterpret_cast<float4*>(Gg + (long)row_a * ld + idx * 4) = a[v];
        if (has_b) *reinterpret_cast<float4*>(Gg + (long)row_b * ld + idx * 4) = b[v];
      }
    }
  }
  if (sweeps_out && crank == 0 && tid == 0) sweeps_out[prob] = sweep;
  if (rot_out) {
    const float tot = block_sum(gl == 0 ? (float)nrot : 0.f, red_scratch);
    if (tid == 0) atomicAdd(rot_out + prob, (int)tot);
  }
}

// Cluster variant for the few-but-large problems (projected-Gram eigenproblems, k x k
// principal-angle SVDs: n up to 1024, a few dozen problems).  One thread-block CLUSTER per
// problem: the matrix stays L2-resident in global memory, the n/2 independent row pairs of a
// round-robin step are split over all warps of the cluster, and the step boundary is the
// hardware cluster barrier (release/acquire at cluster scope) instead of a grid-wide sync.
// Rows are read/written with L1-bypassing accesses so no SM ever sees a stale line.
template <int NV, int R, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
jacobi_rows_cluster_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                           const int* __restrict__ dims, float tol, int max_sweeps,
                           int* __restrict__ sweeps_out, int dim_lo, int dim_hi) {
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = cluster.num_blocks(), crank = cluster.block_rank();
  const int prob = blockIdx.x / csize;
  if (dims && (dims[prob] < dim_lo || dims[prob] > dim_hi)) return;   // whole cluster exits together
  __shared__ int flag;                                   // rank 0's copy is the cluster flag
  __shared__ float red_scratch[32];
  int* flag0 = cluster.map_shared_rank(&flag, 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int gwarp = crank * nwarps + warp, total_warps = csize * nwarps;
  float* G = Gbase + (long)prob * stride;
  const int nn = dims ? min(dims[prob], n) : n;
  const int mm = dims ? min(dims[prob], m) : m;
  const int mv = (mm + 3) >> 2;
  const int ne = nn + (nn & 1), ring = ne - 1, half = ne >> 1;
  const float tol2 = tol * tol;
  int sweep = 0;
  for (; sweep < max_sweeps && nn >= 2; ++sweep) {
    if (crank == 0 && tid == 0) flag = 0;
    cluster.sync();
    float worst = 0.f;
    for (int step = 0; step < ring; ++step) {
      for (int base = 0; base < half; base += total_warps * R) {
        float4 x[R][NV], y[R][NV];
        int pp[R], qq[R];
        bool ok[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {                    // issue every load of this warp first
          const int pair = base + gwarp + r * total_warps;
          int p = 0, q = 0;
          bool valid = pair < half;
          if (valid) {
            if (pair == 0) { p = ne - 1; q = step; }
            else { p = step + pair; if (p >= ring) p -= ring; q = step - pair; if (q < 0) q += ring; }
            valid = (p < nn) && (q < nn);
          }
          pp[r] = p; qq[r] = q; ok[r] = valid;
          const float4* rp = reinterpret_cast<const float4*>(G + (long)p * ld);
          const float4* rq = reinterpret_cast<const float4*>(G + (long)q * ld);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int idx = lane + 32 * v;
            if (valid && idx < mv) {
              x[r][v] = __ldcg(rp + idx);
              y[r][v] = __ldcg(rq + idx);
              const int c4 = idx * 4;                    // zero the tail past mm (dims < ld)
              if (c4 + 1 >= mm) { x[r][v].y = 0.f; y[r][v].y = 0.f; }
              if (c4 + 2 >= mm) { x[r][v].z = 0.f; y[r][v].z = 0.f; }
              if (c4 + 3 >= mm) { x[r][v].w = 0.f; y[r][v].w = 0.f; }
            } else {
              x[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
              y[r][v] = x[r][v];
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const float4 a = x[r][v], b = y[r][v];
            al = fmaf(a.x, a.x, al); al = fmaf(a.y, a.y, al); al = fmaf(a.z, a.z, al); al = fmaf(a.w, a.w, al);
            be = fmaf(b.x, b.x, be); be = fmaf(b.y, b.y, be); be = fmaf(b.z, b.z, be); be = fmaf(b.w, b.w, be);
            ga = fmaf(a.x, b.x, ga); ga = fmaf(a.y, b.y, ga); ga = fmaf(a.z, b.z, ga); ga = fmaf(a.w, b.w, ga);
          }
          al = warp_sum(al);
          be = warp_sum(be);
          ga = warp_sum(ga);
          if (!ok[r] || !(ga * ga > tol2 * al * be)) continue;      // warp-uniform
          worst = fmaxf(worst, __fdividef(ga * ga, al * be));
          const float d = be - al;
          const float h = fmaf(d, d, 4.f * ga * ga);
          const float root = h * rsqrtf(h);
          float t = __fdividef(2.f * fabsf(ga), fabsf(d) + root);
          t = ((d < 0.f) != (ga < 0.f)) ? -t : t;
          const float w2 = fmaf(t, t, 1.f);
          float c = rsqrtf(w2);
          c = c * fmaf(-0.5f * w2, c * c, 1.5f);
          const float sn = c * t;
          float4* rp = reinterpret_cast<float4*>(G + (long)pp[r] * ld);
          float4* rq = reinterpret_cast<float4*>(G + (long)qq[r] * ld);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int idx = lane + 32 * v;
            if (idx < mv) {
              const float4 a = x[r][v], b = y[r][v];
              float4 na, nb;
              na.x = fmaf(c, a.x, -sn * b.x); nb.x = fmaf(sn, a.x, c * b.x);
              na.y = fmaf(c, a.y, -sn * b.y); nb.y = fmaf(sn, a.y, c * b.y);
              na.z = fmaf(c, a.z, -sn * b.z); nb.z = fmaf(sn, a.z, c * b.z);
              na.w = fmaf(c, a.w, -sn * b.w); nb.w = fmaf(sn, a.w, c * b.w);
              __stcg(rp + idx, na);
              __stcg(rq + idx, nb);
            }
          }
        }
      }
      cluster.sync();                                    // step boundary for the whole cluster
    }
    // cluster-wide max of cos^2 (non-negative floats order like their bit patterns)
    worst = block_max(worst, red_scratch);
    if (tid == 0) atomicMax(flag0, __float_as_int(worst));
    cluster.sync();
    const float all_worst = __int_as_float(*flag0);
    cluster.sync();                                      // everyone has read before rank 0 resets
    if (all_worst < tol) { ++sweep; break; }             // quadratic convergence: see grouped kernel
  }
  if (sweeps_out && crank == 0 && tid == 0) sweeps_out[prob] = sweep;
}

// ------------------------------------------------------------------ normalise (+ sort) rows
// in:  G (n x m) whose rows are s_j * v_j^T ; out: Vt rows = v_j^T (zero if s_j <= floor),
// vals[j] = s_j (square=0) or s_j^2 (square=1); sorted descending when sort != 0.
// In-place (out == in) is allowed only when sort == 0.
__global__ void rows_normalize_kernel(const float* __restrict__ Gbase, int n, int m, int ld,
                                      long stride, float* __restrict__ Vbase, int ldv,
                                      long strideV, float* __restrict__ vals, int sort, int square,
                                      float rel_floor, const int* __restrict__ dims) {
  extern __shared__ float sm[];
  float* nrm = sm;                                // n
  int* pos = reinterpret_cast<int*>(sm + n);      // n
  float* red = sm + 2 * n;                        // 32
  const int prob = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float* G = Gbase + (long)prob * stride;
  float* V = Vbase + (long)prob * strideV;
  const int nn = dims ? min(dims[prob], n) : n;
  const int mm = dims ? min(dims[prob], m) : m;
  for (int r = warp; r < nn; r += nwarps) {
    float s = 0.f;
    for (int c = lane; c < mm; c += 32) { const float v = G[(long)r * ld + c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) nrm[r] = sqrtf(s);
  }
  __syncthreads();
  float mx = 0.f;
  for (int r = tid; r < nn; r += T) mx = fmaxf(mx, nrm[r]);
  mx = block_max(mx, red);
  const float floor_v = rel_floor * mx;
  for (int r = tid; r < nn; r += T) {
    int rank = r;
    if (sort) {
      rank = 0;
      const float mine = nrm[r];
      for (int j = 0; j < nn; ++j) {
        const float o = nrm[j];
        rank += (o > mine) || (o == mine && j < r);
      }
    }
    pos[r] = rank;
  }
  __syncthreads();
  for (int r = warp; r < nn; r += nwarps) {
    const float s = nrm[r];
    const bool keep = (s > floor_v) && (s > 0.f);
    const float inv = keep ? 1.f / s : 0.f;
    const int dst = pos[r];
    for (int c = lane; c < m; c += 32)
      V[(long)dst * ldv + c] = (c < mm) ? G[(long)r * ld + c] * inv : 0.f;
    if (lane == 0 && vals) vals[(long)prob * n + dst] = square ? s * s : s;
  }
  // rows beyond the active dimension: zero
  for (int r = nn + warp; r < n; r += nwarps) {
    for (int c = lane; c < m; c += 32) V[(long)r * ldv + c] = 0.f;
    if (lane == 0 && vals) vals[(long)prob * n + r] = 0.f;
  }
}

// out[prob*n + r] = <A[r,:], B[r,:]>
__global__ void rowdot_kernel(const float* __restrict__ A, int lda, long sA,
                              const float* __restrict__ B, int ldb, long sB, int n, int m,
                              float* __restrict__ out) {
  const int prob = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const float* a = A + (long)prob * sA + (long)r * lda;
  const float* b = B + (long)prob * sB + (long)r * ldb;
  float s = 0.f;
  for (int c = lane; c < m; c += 32) s = fmaf(a[c], b[c], s);
  s = warp_sum(s);
  if (lane == 0) out[(long)prob * n + r] = s;
}

static int smem_limit() {
  static int lim = -1;
  if (lim < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return lim;
}

template <int NV>
static int launch_jacobi(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                         float tol, int max_sweeps, int* sweeps_out, cudaStream_t st) {
  BASD_CUDA(cudaFuncSetAttribute(jacobi_rows_kernel<NV>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 0));
  int pairs = (n + 1) / 2;
  int threads = pairs * 32;
  const int cap = NV >= 3 ? 512 : 1024;
  if (threads > cap) threads = cap;
  if (threads < 64) threads = 64;
  jacobi_rows_kernel<NV><<<batch, threads, 0, st>>>(G, n, m, ld, stride, dims, tol, max_sweeps, 0,
                                                    sweeps_out);
  BASD_LAUNCH_CHECK();
  return 0;
}

static int sm_count() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int NV, int R, int MAXT>
static int launch_cluster(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                          float tol, int max_sweeps, int* sweeps_out, cudaStream_t st,
                          int dim_lo = 0, int dim_hi = 1 << 30) {
  int csize = 8;
  while (csize > 1 && (long)batch * csize > sm_count()) csize >>= 1;
  const int half = (n + 1) / 2;
  while (csize > 1 && (csize / 2) * (MAXT / 32) * R >= half) csize >>= 1;   // no idle CTAs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * csize);
  cfg.blockDim = dim3(MAXT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BASD_CUDA(cudaLaunchKernelEx(&cfg, jacobi_rows_cluster_kernel<NV, R, MAXT>, G, n, m, ld, stride,
                               dims, tol, max_sweeps, sweeps_out, dim_lo, dim_hi));
  return 0;
}

template <int LP, int NV, int MAXT>
static int launch_oe_cluster(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                             float tol, int max_sweeps, int* sweeps_out, cudaStream_t st, int dim_lo,
                             int dim_hi, int* rot_out) {
  const int gpc = MAXT / LP;                               // row pairs per CTA
  int csize = 1;
  while (csize < 8 && csize * gpc * 2 < n) csize <<= 1;
  if (csize * gpc * 2 < n) return -12;                     // does not fit a portable cluster
  const size_t dyn = ((size_t)(gpc + 1) * LP * NV * 4 + 2 * (gpc + 2)) * sizeof(float);
  BASD_CUDA(cudaFuncSetAttribute(jacobi_rows_oe_cluster_kernel<LP, NV, MAXT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * csize);
  cfg.blockDim = dim3(gpc * LP);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BASD_CUDA(cudaLaunchKernelEx(&cfg, jacobi_rows_oe_cluster_kernel<LP, NV, MAXT>, G, n, m, ld, stride,
                               dims, tol, max_sweeps, sweeps_out, dim_lo, dim_hi, rot_out));
  return 0;
}

template <int LP, int NV, int MAXT>
static int launch_oddeven(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                          float tol, int max_sweeps, int* sweeps_out, cudaStream_t st,
                          int dim_lo = 0, int dim_hi = 1 << 30, int* rot_out = nullptr) {
  const int cap = (dims && dim_hi < n) ? dim_hi : n;
  const size_t slots = (size_t)(cap + 1) / 2 + 1;
  const size_t dyn = (slots * LP * NV * 4 + 2 * slots + 4) * sizeof(float);
  BASD_CUDA(cudaFuncSetAttribute(jacobi_rows_oddeven_kernel<LP, NV, MAXT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int threads = ((cap + 1) / 2) * LP;
  threads = (threads + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  jacobi_rows_oddeven_kernel<LP, NV, MAXT><<<batch, threads, dyn, st>>>(
      G, n, m, ld, stride, dims, tol, max_sweeps, sweeps_out, dim_lo, dim_hi, rot_out);
  BASD_LAUNCH_CHECK();
  return 0;
}

template <int LP, int NV, int MAXT>
static int launch_grouped(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                          float tol, int max_sweeps, int* sweeps_out, size_t dyn, cudaStream_t st) {
  BASD_CUDA(cudaFuncSetAttribute(jacobi_rows_grouped_kernel<LP, NV, MAXT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int threads = ((n + 1) / 2) * LP;
  threads = (threads + 31) / 32 * 32;
  if (threads > MAXT) threads = MAXT / 32 * 32;
  if (threads < 64) threads = 64;
  jacobi_rows_grouped_kernel<LP, NV, MAXT><<<batch, threads, dyn, st>>>(
      G, n, m, ld, stride, dims, tol, max_sweeps, sweeps_out);
  BASD_LAUNCH_CHECK();
  return 0;
}

}  // namespace basd

namespace basd {
// cholesky_reg.cu, opt-in experiment (BASD_CHOL_REG=1): register-resident left-looking pivoted Cholesky
int launch_pivoted_cholesky_reg(const float* K, int n, int ld, long stride_k, float* LT, int ldl,
                                long stride_l, int batch, float rel_tol, int* rank_out, const int* dims,
                                cudaStream_t st, int lanes_per_row);
}  // namespace basd

extern "C" int basd_pivoted_cholesky(float* K, int n, int ld, long stride_k, float* LT, int ldl,
                                     long stride_l, int batch, float rel_tol, int* rank_out,
                                     const int* dims, void* stream) {
  using namespace basd;
  if (batch <= 0 || n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {   // opt-in experiment (cholesky_reg.cu): factor rows resident in registers, not yet measured
    static const int reg = getenv("BASD_CHOL_REG") ? atoi(getenv("BASD_CHOL_REG")) : 0;   // 1 or 4: four lanes per row, 2: two
    if (reg) {
      const int e = launch_pivoted_cholesky_reg(K, n, ld, stride_k, LT, ldl, stride_l, batch, rel_tol, rank_out,
                                                dims, st, reg == 2 ? 2 : 4);
      if (e != -100) return e;
    }
  }
  {   // left-looking, four outputs per thread: factor rows resident in shared memory
    const size_t npad = ((size_t)n + 127) & ~(size_t)127;
    const int wpp = (int)(npad >> 7);
    int warps = 28 / wpp * wpp;                              // <= 896 threads, whole parts
    if (warps < wpp) warps = wpp;
    const int parts = warps / wpp;
    const size_t dyn4 = ((size_t)n * npad + npad + (size_t)parts * npad) * sizeof(float);
    static const bool no_left4 = getenv("BASD_CHOL_LEFT1") != nullptr;
    // n <= 128 pads to 128 rows per warp group and sums 28 partials per output: the one-output
    // kernel below is faster there (N = 64, 4,096 problems: 15.3 vs 17.6 ms per step)
    if (!no_left4 && n > 128 && !getenv("BASD_CHOL_RIGHT") && dyn4 + 2048 <= (size_t)smem_limit() &&
        warps * 32 <= 1024 && warps * 32 >= (int)npad) {
      BASD_CUDA(cudaFuncSetAttribute(pivoted_cholesky_left4_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn4));
      pivoted_cholesky_left4_kernel<<<batch, warps * 32, dyn4, st>>>(K, n, ld, stride_k, LT, ldl,
                                                                    stride_l, rel_tol, rank_out, dims);
      BASD_LAUNCH_CHECK();
      return 0;
    }
  }
  {   // left-looking kernel: factor rows resident in shared memory
    constexpr int PARTS = 4;
    const size_t np32 = ((size_t)n + 31) & ~(size_t)31;
    const size_t dyn_left = (np32 * np32 + np32 * (1 + PARTS)) * sizeof(float);
    const int threads_left = (int)(np32 * PARTS);
    if (!getenv("BASD_CHOL_RIGHT") && dyn_left + 2048 <= (size_t)smem_limit() && threads_left <= 1024) {
      BASD_CUDA(cudaFuncSetAttribute(pivoted_cholesky_left_kernel<PARTS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_left));
      pivoted_cholesky_left_kernel<PARTS><<<batch, threads_left, dyn_left, st>>>(
          K, n, ld, stride_k, LT, ldl, stride_l, rel_tol, rank_out, dims);
      BASD_LAUNCH_CHECK();
      return 0;
    }
  }
  {   // factor rows dealt to a cluster of 4 CTAs -- or 16 (non-portable size, one cluster per GPC) when
      // a quarter of the factor does not fit one SM's shared memory (768 x 768: 48 rows per CTA)
    const size_t npc = ((size_t)n + 127) & ~(size_t)127;
    const int wpp = (int)(npc >> 7);
    const int warps = 28 / wpp * wpp;
    const int parts = warps > 0 ? warps / wpp : 0;
    static const bool no_chol_cluster = getenv("BASD_CHOL_NO_CLUSTER") != nullptr;
    // few large problems only: with thousands of them one CTA per problem keeps every SM busy
    // (N = 256, 1,024 problems: 91 vs 82 ms per step with the clusters)
    const bool eligible = !no_chol_cluster && !getenv("BASD_CHOL_RIGHT") && parts >= 1 &&
                          warps * 32 >= (int)npc;
    for (int cs = 4; eligible && cs <= 16; cs *= 4) {
      const size_t rows_local = ((size_t)n + cs - 1) / cs;
      const size_t dync = (rows_local * npc + 3 * npc + (size_t)parts * npc) * sizeof(float);
      if (dync + 2048 > (size_t)smem_limit() || batch * cs > 4 * sm_count()) continue;
      auto kern = cs == 4 ? pivoted_cholesky_left4_cluster_kernel<4> : pivoted_cholesky_left4_cluster_kernel<16>;
      BASD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dync));
      if (cs > 8) BASD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(batch * cs);
      cfg.blockDim = dim3(warps * 32);
      cfg.dynamicSmemBytes = dync;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cs;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaLaunchKernelEx(&cfg, kern, (const float*)K, n, ld, stride_k, LT, ldl, stride_l, rel_tol,
                             rank_out, dims) == cudaSuccess)
        return 0;
      (void)cudaGetLastError();                              // refused cluster shape: try the next route
    }
  }
  const size_t npad = ((size_t)n + 3) & ~(size_t)3;
  const size_t base = (size_t)(2 * npad + 32) * sizeof(float);
  const size_t staged = base + npad * npad * sizeof(float);
  if (ld & 3) return -3;
  const int use_smem = staged + 1024 <= (size_t)smem_limit();
  const size_t dyn = use_smem ? staged : base;
  BASD_CUDA(cudaFuncSetAttribute(pivoted_cholesky_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int threads = n > 256 ? 1024 : (n >= 128 ? 512 : 256);
  pivoted_cholesky_kernel<<<batch, threads, dyn, st>>>(K, n, ld, stride_k, LT, ldl, stride_l,
                                                       rel_tol, rank_out, dims, use_smem);
  BASD_LAUNCH_CHECK();
  return 0;
}

namespace basd {
// jacobi_oe8.cu, opt-in experiment (BASD_JACOBI_SPLIT=2|4): small full problems split over a cluster of
// 2 or 4 CTAs with several CTAs resident per SM.  0 = off (the default).
int launch_jacobi_oe8_split(float* G, int n, int m, int ld, long stride, int batch, const int* dims, float tol,
                            int max_sweeps, int* sweeps_out, cudaStream_t st, int* rot_out, int csize, int dim_lo,
                            int dim_hi, int rows_only = 0);
static int jacobi_split_csize() {
  static const int v = [] {
    const char* e = getenv("BASD_JACOBI_SPLIT");
    const int c = e ? atoi(e) : 0;
    return (c == 2 || c == 4) ? c : 0;
  }();
  return v;
}
}  // namespace basd

// Orthogonalises the rows of each (n x m) row-major matrix in place. ld % 4 == 0 and
// 16-byte aligned bases are required (128-bit row accesses).
extern "C" int basd_jacobi_rows_counted(float* G, int n, int m, int ld, long stride, int batch,
                                        const int* dims, float tol, int max_sweeps, int* sweeps_out,
                                        int* rot_out, void* stream);

extern "C" int basd_jacobi_rows(float* G, int n, int m, int ld, long stride, int batch,
                                const int* dims, float tol, int max_sweeps, int* sweeps_out,
                                void* stream) {
  return basd_jacobi_rows_counted(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out,
                                  nullptr, stream);
}

// Rank-deficient factor products (G = F_q^T F_p with a pivoted-Cholesky F_q of rank r): only the
// first row_dims[problem] rows are non-zero, the others are exact zeros that no rotation touches.
// The register-resident kernel then sweeps over the active rows only (C3: rank 48 of 196 -> a
// quarter of the steps); other shapes run the full problem, which gives the same result.
extern "C" int basd_jacobi_rows_ranked(float* G, int n, int m, int ld, long stride, int batch,
                                       const int* row_dims, float tol, int max_sweeps,
                                       int* sweeps_out, int* rot_out, void* stream) {
  using namespace basd;
  if (batch <= 0 || n <= 0) return 0;
  if ((ld & 3) || (stride & 3) || (reinterpret_cast<uintptr_t>(G) & 15)) return -3;
  static const bool no_oe8 = getenv("BASD_JACOBI_NO_OE8") != nullptr;
  if (row_dims && basd::jacobi_split_csize() && n <= 256 && m <= 208) {   // opt-in experiment (rank-aware split)
    const int e = launch_jacobi_oe8_split(G, n, m, ld, stride, batch, row_dims, tol, max_sweeps, sweeps_out,
                                          (cudaStream_t)stream, rot_out, basd::jacobi_split_csize(), 0, 1 << 30, 1);
    if (e != -100) return e;
  }
  if (row_dims && !no_oe8 && n <= 256 && m <= 256 && !basd::jacobi_split_csize()) {
    const int e = launch_jacobi_oe8(G, n, m, ld, stride, batch, row_dims, tol, max_sweeps, sweeps_out,
                                    (cudaStream_t)stream, 0, 1 << 30, rot_out, 1);
    if (e != -100) return e;
  }
  return basd_jacobi_rows_counted(G, n, m, ld, stride, batch, nullptr, tol, max_sweeps, sweeps_out,
                                  rot_out, stream);
}

// Same, additionally accumulating into rot_out[problem] the number of plane rotations applied
// (bench.py's roofline leg; only the register-resident kernels count, others leave it untouched).
extern "C" int basd_jacobi_rows_counted(float* G, int n, int m, int ld, long stride, int batch,
                                        const int* dims, float tol, int max_sweeps, int* sweeps_out,
                                        int* rot_out, void* stream) {
  using namespace basd;
  if (batch <= 0 || n <= 0) return 0;
  if ((ld & 3) || (stride & 3) || (reinterpret_cast<uintptr_t>(G) & 15)) return -3;
  cudaStream_t st = (cudaStream_t)stream;
  const int quads = (m + 3) / 4;
  // shared-memory resident path: LP lanes per pair
  const size_t need = ((size_t)n * quads * 4 + n) * sizeof(float);
  const bool legacy = getenv("BASD_JACOBI_LEGACY") != nullptr;   // A/B debugging aids
  const bool no_oddeven = getenv("BASD_JACOBI_ROUNDROBIN") != nullptr;
  // eight rows per 16-lane group (jacobi_oe8.cu): a quarter of the shared-memory exchange traffic
  static const bool no_oe8 = getenv("BASD_JACOBI_NO_OE8") != nullptr;
  // opt-in experiment: full (dims == null) small problems split over 2 or 4 CTAs, several CTAs per SM
  if (!dims && jacobi_split_csize() && n <= 256 && m <= 208) {
    const int e = launch_jacobi_oe8_split(G, n, m, ld, stride, batch, nullptr, tol, max_sweeps, sweeps_out, st,
                                          rot_out, jacobi_split_csize(), 0, 1 << 30);
    if (e != -100) return e;
  }
  if (!legacy && !no_oddeven && !no_oe8 && n <= 256 && m <= 256) {
    const int e = launch_jacobi_oe8(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st,
                                    0, 1 << 30, rot_out);
    if (e != -100) return e;
  }
  // register-resident odd-even kernel: every row pair needs its own group of 8 lanes
  if (!legacy && !no_oddeven && ((n + 1) / 2) * 8 <= 800 && quads <= 56) {
#define BASD_OE(NV, MAXT) \
  return launch_oddeven<8, NV, MAXT>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, 0, 1 << 30, rot_out)
    if (quads <= 8) BASD_OE(1, 800);
    if (quads <= 16) BASD_OE(2, 800);
    if (quads <= 24) BASD_OE(3, 800);
    if (quads <= 32) BASD_OE(4, 800);
    if (quads <= 40) BASD_OE(5, 800);
    if (quads <= 48) BASD_OE(6, 800);
    BASD_OE(7, 800);
#undef BASD_OE
  }
  if (!legacy && need + 1024 <= (size_t)smem_limit()) {
#define BASD_GROUPED(LP, NV, MAXT) \
  return launch_grouped<LP, NV, MAXT>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, need, st)
    if (quads <= 8) BASD_GROUPED(8, 1, 1024);
    if (quads <= 16) BASD_GROUPED(8, 2, 1024);
    if (quads <= 24) BASD_GROUPED(8, 3, 1024);
    if (quads <= 32) BASD_GROUPED(8, 4, 1024);
    if (quads <= 40) BASD_GROUPED(8, 5, 896);
    if (quads <= 48) BASD_GROUPED(8, 6, 800);
    if (quads <= 56) BASD_GROUPED(8, 7, 800);
    if (quads <= 64) BASD_GROUPED(16, 4, 1024);
    if (quads <= 96) BASD_GROUPED(16, 6, 832);
    if (quads <= 128) BASD_GROUPED(32, 4, 1024);
#undef BASD_GROUPED
  }
  const int nv = (quads + 31) / 32;
  int lo = 0;
  if (!legacy && !no_oddeven && dims && n == m) {
    // square problems with a device-side active size (k x k principal-angle SVDs): those with
    // k <= 200 run register/shared-memory resident, the rest on the cluster kernel below
    constexpr int SMALL = 200;
    int e = -100;
    if (jacobi_split_csize())                              // opt-in experiment, see launch_jacobi_oe8_split
      e = launch_jacobi_oe8_split(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, rot_out,
                                  jacobi_split_csize(), 0, SMALL);
    if (e == -100)
      e = no_oe8 ? -100 : launch_jacobi_oe8(G, n, m, ld, stride, batch, dims, tol, max_sweeps,
                                            sweeps_out, st, 0, SMALL, rot_out);
    if (e == -100)
      e = launch_oddeven<8, 7, 800>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out,
                                    st, 0, SMALL, rot_out);
    if (e) return e;
    lo = SMALL + 1;
  }
  // eight rows per group spread over a cluster (jacobi_oe8.cu): up to 768 rows x 384 columns
  static const bool no_oe8c = getenv("BASD_JACOBI_NO_OE8_CLUSTER") != nullptr;
  if (!legacy && !no_oddeven && !no_oe8 && !no_oe8c && (m <= 384 || !dims) && m <= 768 && n <= 768) {
    const int e = launch_jacobi_oe8_cluster(G, n, m, ld, stride, batch, dims, tol, max_sweeps,
                                            sweeps_out, st, lo, 1 << 30, rot_out);
    if (e == 0) return 0;                                  // (-100 or a refused non-portable cluster: fall through)
    (void)cudaGetLastError();
  }
  // wider allocations with a device-side active size (C4: k ~ 366 of 768): the problems whose
  // active size fits take the cluster kernel through its size window, the rest fall through
  if (!legacy && !no_oddeven && !no_oe8 && !no_oe8c && dims && n == m && m > 384 && lo <= 384) {
    const int e = launch_jacobi_oe8_cluster(G, n, m, ld, stride, batch, dims, tol, max_sweeps,
                                            sweeps_out, st, lo, 384, rot_out);
    if (e == 0) lo = 385;
    else if (e != -100) return e;
  }
  // register-resident rows spread over a cluster (16 lanes per pair, 48 pairs per CTA)
  static const bool no_oe_cluster = getenv("BASD_JACOBI_L2CLUSTER") != nullptr;
  if (!legacy && !no_oddeven && !no_oe_cluster && quads <= 16 * 7 && n <= 8 * 96) {
    const int q16 = (quads + 15) / 16;
#define BASD_OEC(NV) \
  return launch_oe_cluster<16, NV, 768>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, lo, 1 << 30, rot_out)
    if (q16 <= 4) BASD_OEC(4);
    if (q16 == 5) BASD_OEC(5);
    if (q16 == 6) BASD_OEC(6);
    BASD_OEC(7);
#undef BASD_OEC
  }
  if (!legacy) {
    switch (nv) {
      case 1: return launch_cluster<1, 2, 1024>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, lo);
      case 2: return launch_cluster<2, 2, 1024>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, lo);
      case 3: return launch_cluster<3, 2, 768>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, lo);
      case 4: return launch_cluster<4, 1, 1024>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, lo);
      case 5: case 6:
        return launch_cluster<6, 1, 512>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, lo);
      case 7: case 8:
        return launch_cluster<8, 1, 512>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st, lo);
      default: return -4;
    }
  }
  switch (nv) {
    case 1: return launch_jacobi<1>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st);
    case 2: return launch_jacobi<2>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st);
    case 3: return launch_jacobi<3>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st);
    case 4: return launch_jacobi<4>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st);
    case 5: case 6:
      return launch_jacobi<6>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st);
    case 7: case 8:
      return launch_jacobi<8>(G, n, m, ld, stride, batch, dims, tol, max_sweeps, sweeps_out, st);
    default: return -4;  // m > 1024 unsupported
  }
}

extern "C" int basd_rows_normalize(const float* G, int n, int m, int ld, long stride, float* V,
                                   int ldv, long stride_v, float* vals, int batch, int sort,
                                   int square, float rel_floor, const int* dims, void* stream) {
  using namespace basd;
  if (batch <= 0 || n <= 0) return 0;
  if (sort && G == V) return -5;
  const size_t dyn = (size_t)(2 * n + 32) * sizeof(float);
  rows_normalize_kernel<<<batch, 512, dyn, (cudaStream_t)stream>>>(
      G, n, m, ld, stride, V, ldv, stride_v, vals, sort, square, rel_floor, dims);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_rowdot(const float* A, int lda, long stride_a, const float* B, int ldb,
                           long stride_b, int n, int m, int batch, float* out, void* stream) {
  using namespace basd;
  if (batch <= 0 || n <= 0) return 0;
  dim3 grid((n + 7) / 8, batch);
  rowdot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, stride_a, B, ldb, stride_b, n, m,
                                                        out);
  BASD_LAUNCH_CHECK();
  return 0;
}

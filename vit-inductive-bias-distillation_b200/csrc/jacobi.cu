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

// Left-looking, four outputs per thread (thread (part, iq) accumulates rows 4iq..4iq+3 of the new column
// over the factor rows k = part, part + PARTS, ... with ONE 128-bit load of L[k][4iq..] and one broadcast
// load of L[k][p] per four FMAs), factor rows dealt round-robin to the CTAs of a cluster:
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
// Cluster variant for the few-but-large problems (projected-Gram eigenproblems, k x k
// principal-angle SVDs: n up to 1024, a few dozen problems).  One thread-block CLUSTER per
// problem: the matrix stays L2-resident in global memory, the n/2 independent row pairs of a
// round-robin step are split over all warps of the cluster, and the step boundary is the
// hardware cluster barrier (release/acquire at cluster scope) instead of a grid-wide sync.
// Rows are read/written with L1-bypassing accesses so no SM ever sees a stale line.
template <int NV, int R, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
jacobi_rows_cluster_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                           const int* __restrict__ dims, float tol, float stop2, int max_sweeps,
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
    if (all_worst < stop2) { ++sweep; break; }             // quadratic convergence: see grouped kernel
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

static int sm_count() {
  static int n = -1;
  if (n < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int NV, int R, int MAXT>
static int launch_cluster(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                          float tol, float stop2, int max_sweeps, int* sweeps_out, cudaStream_t st,
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
                               dims, tol, stop2, max_sweeps, sweeps_out, dim_lo, dim_hi));
  return 0;
}

}  // namespace basd

namespace basd {
// cholesky_reg.cu: register-resident left-looking pivoted Cholesky (128 < n <= 208)
int launch_pivoted_cholesky_reg(const float* K, int n, int ld, long stride_k, float* LT, int ldl,
                                long stride_l, int batch, float rel_tol, int* rank_out, const int* dims,
                                cudaStream_t st, int lanes_per_row);
}  // namespace basd

// Routes by size: n <= 128 one-output left-looking kernel (factor rows in shared memory); 128 < n <= 208
// register-resident rows, two lanes per row (1.73 -> 1.18 ms per 1,024 x 196^2 launch on B200); up to 224
// the shared-memory kernel again; few large problems (384, 768) a cluster of 4 / 16 CTAs; anything else the
// right-looking kernel on the L2-resident trailing matrix.
extern "C" int basd_pivoted_cholesky(float* K, int n, int ld, long stride_k, float* LT, int ldl,
                                     long stride_l, int batch, float rel_tol, int* rank_out,
                                     const int* dims, void* stream) {
  using namespace basd;
  if (batch <= 0 || n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int e = launch_pivoted_cholesky_reg(K, n, ld, stride_k, LT, ldl, stride_l, batch, rel_tol, rank_out,
                                              dims, st, 2);
    if (e != -100) return e;
  }
  {   // left-looking kernel: factor rows resident in shared memory
    constexpr int PARTS = 4;
    const size_t np32 = ((size_t)n + 31) & ~(size_t)31;
    const size_t dyn_left = (np32 * np32 + np32 * (1 + PARTS)) * sizeof(float);
    const int threads_left = (int)(np32 * PARTS);
    if (dyn_left + 2048 <= (size_t)smem_limit() && threads_left <= 1024) {
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
    // few large problems only: with thousands of them one CTA per problem keeps every SM busy
    // (N = 256, 1,024 problems: 91 vs 82 ms per step with the clusters)
    const bool eligible = parts >= 1 && warps * 32 >= (int)npc;
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
// jacobi_oe8.cu: the same sweep with one problem split over a cluster of 4 CTAs, three CTAs per SM
int launch_jacobi_oe8_split(float* G, int n, int m, int ld, long stride, int batch, const int* dims, float tol, float stop2,
                            int max_sweeps, int* sweeps_out, cudaStream_t st, int* rot_out, int csize, int dim_lo,
                            int dim_hi, int rows_only = 0);

// A launch with few problems (the k x k principal-angle SVDs: E x L_t = 48 at C2) leaves most SMs idle on the
// one-CTA-per-problem kernel; four CTAs per problem measured 2.21 -> 1.64 ms there.  With a full wave of
// problems the single-CTA kernel wins (1,024 x 196^2: 16.1 vs 19.4 ms), so the split is taken only while
// its CTAs still fit the GPU twice over.
static bool few_problems(int batch) { return (long)batch * 4 <= 2L * sm_count(); }
}  // namespace basd

// Orthogonalises the rows of each (n x m) row-major matrix in place. ld % 4 == 0 and
// 16-byte aligned bases are required (128-bit row accesses).
//
// Two thresholds on the cosine between two rows: a pair is rotated while |cos| > tol, and the sweeps stop
// after one in which every rotated pair had |cos| < stop_cos.  stop_cos = sqrt(tol) relies on the quadratic
// convergence of the sweep (what is left after such a sweep is O(stop_cos^2) = tol) and is what the
// legacy entry points use: right for singular values, polar factors and well-separated spectra.  It is
// NOT enough for eigenvectors of a dense spectrum: two rows whose norms differ by a relative gap g keep
// mixing by an angle ~ cos / g, and with 50,176 token rows the selector Grams have g ~ 1e-3 at the
// Marchenko-Pastur rank boundary (measured at C2, B = 256: selector gradient cosine 0.87 with
// stop_cos = 1e-3) -- sym_eig passes a tighter stop_cos through basd_jacobi_rows_ex.
static int jacobi_dispatch(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                           const int* row_dims, float tol, float stop2, int max_sweeps, int* sweeps_out,
                           int* rot_out, void* stream);

extern "C" int basd_jacobi_rows_ex(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                                   const int* row_dims, float tol, float stop_cos, int max_sweeps,
                                   int* sweeps_out, int* rot_out, void* stream) {
  if (dims && row_dims) return -5;
  return jacobi_dispatch(G, n, m, ld, stride, batch, dims, row_dims, tol, stop_cos * stop_cos, max_sweeps,
                         sweeps_out, rot_out, stream);
}

extern "C" int basd_jacobi_rows(float* G, int n, int m, int ld, long stride, int batch,
                                const int* dims, float tol, int max_sweeps, int* sweeps_out,
                                void* stream) {
  return jacobi_dispatch(G, n, m, ld, stride, batch, dims, nullptr, tol, tol, max_sweeps, sweeps_out, nullptr,
                         stream);
}

// Rank-deficient factor products (G = F_q^T F_p with a pivoted-Cholesky F_q of rank r): only the
// first row_dims[problem] rows are non-zero, the others are exact zeros that no rotation touches.
// The register-resident kernel then sweeps over the active rows only (C3: rank 48 of 196 -> a
// quarter of the steps); other shapes run the full problem, which gives the same result.
extern "C" int basd_jacobi_rows_ranked(float* G, int n, int m, int ld, long stride, int batch,
                                       const int* row_dims, float tol, int max_sweeps,
                                       int* sweeps_out, int* rot_out, void* stream) {
  return jacobi_dispatch(G, n, m, ld, stride, batch, nullptr, row_dims, tol, tol, max_sweeps, sweeps_out,
                         rot_out, stream);
}

// Same as basd_jacobi_rows, additionally accumulating into rot_out[problem] the number of plane rotations
// applied (bench.py's roofline leg; only the register-resident kernels count, others leave it untouched).
extern "C" int basd_jacobi_rows_counted(float* G, int n, int m, int ld, long stride, int batch,
                                        const int* dims, float tol, int max_sweeps, int* sweeps_out,
                                        int* rot_out, void* stream) {
  return jacobi_dispatch(G, n, m, ld, stride, batch, dims, nullptr, tol, tol, max_sweeps, sweeps_out, rot_out,
                         stream);
}

// Routes: rows and columns <= 256 -> one CTA per problem, eight rows per 16-lane group in registers
// (jacobi_oe8.cu), or four CTAs per problem when the launch has few problems; square problems with a
// device-side active size (dims) in a wider allocation -> the same kernels through their size window
// [0, 200], the rest on the cluster kernel; up to 768 x 768 -> the eight-row layout spread over a
// thread-block cluster; larger -> the L2-resident cluster kernel above.
static int jacobi_dispatch(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                           const int* row_dims, float tol, float stop2, int max_sweeps, int* sweeps_out,
                           int* rot_out, void* stream) {
  using namespace basd;
  if (batch <= 0 || n <= 0) return 0;
  if ((ld & 3) || (stride & 3) || (reinterpret_cast<uintptr_t>(G) & 15)) return -3;
  if (row_dims && n <= 256 && m <= 256) {
    const int e = launch_jacobi_oe8(G, n, m, ld, stride, batch, row_dims, tol, stop2, max_sweeps, sweeps_out,
                                    (cudaStream_t)stream, 0, 1 << 30, rot_out, 1);
    if (e != -100) return e;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 256 && m <= 256) {
    int e = -100;
    if (few_problems(batch))
      e = launch_jacobi_oe8_split(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, rot_out, 4, 0,
                                  1 << 30);
    if (e == -100)
      e = launch_jacobi_oe8(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, 0, 1 << 30, rot_out);
    if (e != -100) return e;
  }
  int lo = 0;
  if (dims && n == m) {
    constexpr int SMALL = 200;
    int e = -100;
    if (few_problems(batch))
      e = launch_jacobi_oe8_split(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, rot_out, 4, 0,
                                  SMALL);
    if (e == -100)
      e = launch_jacobi_oe8(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, 0, SMALL, rot_out);
    if (e == 0) lo = SMALL + 1;
    else if (e != -100) return e;
  }
  if ((m <= 384 || !dims) && m <= 768 && n <= 768) {
    const int e = launch_jacobi_oe8_cluster(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps,
                                            sweeps_out, st, lo, 1 << 30, rot_out);
    if (e == 0) return 0;                                  // (-100 or a refused non-portable cluster: fall through)
    (void)cudaGetLastError();
  }
  // wider allocations with a device-side active size (C4: k ~ 366 of 768): the problems whose
  // active size fits take the cluster kernel through its size window, the rest fall through
  if (dims && n == m && m > 384 && m <= 768 && lo <= 384) {
    const int e = launch_jacobi_oe8_cluster(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps,
                                            sweeps_out, st, lo, 384, rot_out);
    if (e == 0) lo = 385;
    else if (e != -100) return e;
  }
  const int quads = (m + 3) / 4;
  switch ((quads + 31) / 32) {
    case 1: return launch_cluster<1, 2, 1024>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, lo);
    case 2: return launch_cluster<2, 2, 1024>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, lo);
    case 3: return launch_cluster<3, 2, 768>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, lo);
    case 4: return launch_cluster<4, 1, 1024>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, lo);
    case 5: case 6:
      return launch_cluster<6, 1, 512>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, lo);
    case 7: case 8:
      return launch_cluster<8, 1, 512>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, lo);
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

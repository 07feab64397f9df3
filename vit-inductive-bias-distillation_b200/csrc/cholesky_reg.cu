// Register-resident, left-looking, diagonally pivoted Cholesky for the per-sample N x N Grams
// (128 < n <= 208; measured on B200, 1,024 x 196^2: 1.73 ms for the shared-memory kernel it replaced,
// 1.37 ms with four lanes per row, 1.18 ms with two -- the default; n = 150, 256 problems: 0.34 / 0.22 /
// 0.20 ms).
//
// The shared-memory kernel re-read all j finished columns at step j: ~100 KB of shared-memory traffic per
// step at n = 196, ~2,400 cycles per step measured.  Here every row i of L lives in the REGISTERS of the four lanes that own the row
// (lane q of the quad holds columns 16 g + 4 q .. + 3 of every 16-column group g), so a step only
// moves the pivot row: its four owners publish L[p][0 .. 16 (g + 1)) to shared memory, everyone
// reads it back as broadcast 128-bit loads and forms its row's dot product with (g + 1) x 4 FMAs.
// The group loop is unrolled, so the dot length, the slot a new entry is written to and the number
// of loads are compile-time constants -- no dynamically indexed registers; unset slots are zero and
// simply contribute nothing.
//
// Same contract as basd_pivoted_cholesky (reference: the Cholesky inside torch.linalg.svd /
// matrix_norm(ord="nuc") replacements, relational.py:48): LT row j = column j of L, rows >= rank
// zero, pivot = largest remaining diagonal (lowest index on ties), stop below rel_tol * max diag.
#include "common.cuh"
#include <cstdlib>

namespace basd {
namespace chreg {

constexpr int GW = 16;            // columns per group: LPR lanes x (16 / LPR) floats

// LPR lanes own a row: 4 (800 threads, 52 data registers) or 2 (416 threads, 104 data registers: half the
// warps, so the per-step overhead -- pivot reduction, quad reduce, argmax, barriers -- is issued half as often)
template <int NG, int LPR>        // column groups: n <= 16 NG
__global__ void __maxnreg__(LPR == 4 ? 72 : 128)   // warps are allocated in fours: 28 x 32 x 72 and 16 x 32 x 128 registers fit the SM
pivoted_cholesky_reg_kernel(const float* __restrict__ Kbase, int n, int ld, long strideK,
                            float* __restrict__ LTbase, int ldl, long strideL, float rel_tol,
                            int* __restrict__ rank_out, const int* __restrict__ dims) {
  extern __shared__ __align__(16) float Ks[];             // K itself, staged once (n x n): the pivot
                                                          // row of K costs a shared-memory load per step
                                                          // instead of an L2 round trip on the critical path
  __shared__ __align__(16) float prow[GW * NG];           // the pivot row of L, published per step
  __shared__ unsigned int best_val[2][32];
  __shared__ int best_idx[2][32];
  __shared__ float red[32];
  const int prob = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
  const int i = tid / LPR, q = tid % LPR;                 // row, lane within the row's quad
  const float* Kg = Kbase + (long)prob * strideK;
  float* LT = LTbase + (long)prob * strideL;
  const int nn = dims ? min(dims[prob], n) : n;
  const bool live_row = i < nn;

  auto warp_argmax = [](float v, int idx_in, unsigned& vbits, int& idx) {
    const unsigned bits = __float_as_uint(fmaxf(v, 0.f));
    vbits = __reduce_max_sync(0xffffffffu, bits);
    const unsigned who = __ballot_sync(0xffffffffu, bits == vbits);
    idx = __shfl_sync(0xffffffffu, idx_in, __ffs(who) - 1);
  };

  constexpr int W = GW / LPR, NV = W / 4;                 // floats / float4s per lane per group
  float4 l[NG][NV];                                       // this lane's slice of row i of L
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int v = 0; v < NV; ++v) l[g][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = tid; e < nn * nn; e += T) {
    const int r = e / nn, c = e - r * nn;
    Ks[e] = Kg[(long)r * ld + c];
  }
  if (tid < 64) { best_val[tid >> 5][tid & 31] = 0u; best_idx[tid >> 5][tid & 31] = 0; }
  __syncthreads();
  float di = live_row ? Ks[i * nn + i] : -1.f;            // remaining diagonal (replicated in the quad)
  {
    unsigned vb;
    int ib;
    warp_argmax(q == 0 ? di : -1.f, i, vb, ib);
    if (lane == 0) { best_val[0][warp] = vb; best_idx[0][warp] = ib; }
  }
  const float dmax = block_max(fmaxf(di, 0.f), red);      // (its barriers also publish best_val[0])
  const float floor_v = rel_tol * dmax;
  __syncthreads();

  int rank = 0;
  bool done = false;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
#pragma unroll 1
    for (int jq = 0; jq < LPR && !done; ++jq) {           // the lane that receives the new entries
#pragma unroll
      for (int js = 0; js < W; ++js) {                    // ... in slot js of its l[g]
        if (done) break;
        const int j = GW * g + W * jq + js;
        if (j >= nn) { done = true; break; }
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
        if (!(best > floor_v) || !(best > 0.f)) { done = true; break; }   // uniform across the block
        // pivot row of K (== pivot column by symmetry)
        const float kp = live_row ? Ks[p * nn + i] : 0.f;
        if (i == p) {
#pragma unroll
          for (int gg = 0; gg <= g; ++gg)
#pragma unroll
            for (int v = 0; v < NV; ++v) *reinterpret_cast<float4*>(prow + GW * gg + W * q + 4 * v) = l[gg][v];
        }
        __syncthreads();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int gg = 0; gg <= g; ++gg) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const float4 pv = *reinterpret_cast<const float4*>(prow + GW * gg + W * q + 4 * v);
            a0 = fmaf(l[gg][v].x, pv.x, a0);
            a1 = fmaf(l[gg][v].y, pv.y, a1);
            a2 = fmaf(l[gg][v].z, pv.z, a2);
            a3 = fmaf(l[gg][v].w, pv.w, a3);
          }
        }
        float acc = (a0 + a1) + (a2 + a3);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (LPR == 4) acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        float c = 0.f, nd = -1.f;
        if (live_row && di >= 0.f) {
          const float rs = rsqrtf(best);
          c = (i == p) ? best * rs : (kp - acc) * rs;
          nd = (i == p) ? -1.f : fmaxf(fmaf(-c, c, di), 0.f);
          di = nd;
        }
        if (q == jq) {                                    // static slot: component js % 4 of l[g][js / 4]
          if ((js & 3) == 0) l[g][js >> 2].x = c;
          else if ((js & 3) == 1) l[g][js >> 2].y = c;
          else if ((js & 3) == 2) l[g][js >> 2].z = c;
          else l[g][js >> 2].w = c;
        }
        if (q == 0 && live_row) LT[(long)j * ldl + i] = c;
        {
          unsigned vbn;
          int ibn;
          warp_argmax((q == 0 && live_row) ? nd : -1.f, i, vbn, ibn);
          if (lane == 0) { best_val[(j + 1) & 1][warp] = vbn; best_idx[(j + 1) & 1][warp] = ibn; }
        }
        __syncthreads();
        rank = j + 1;
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += T) {
    const int r = e / n, c = e - r * n;
    if (r >= rank || c >= nn) LT[(long)r * ldl + c] = 0.f;
  }
  if (rank_out && tid == 0) rank_out[prob] = rank;
}

}  // namespace chreg

// 128 < n <= 200 with four lanes per row (800 threads, 80 registers), n <= 208 with two (416 threads);
// returns -100 when the shape does not fit (the caller falls back to the shared-memory kernels).
int launch_pivoted_cholesky_reg(const float* K, int n, int ld, long stride_k, float* LT, int ldl,
                                long stride_l, int batch, float rel_tol, int* rank_out, const int* dims,
                                cudaStream_t st, int lanes_per_row) {
  if (n <= 128 || n > (lanes_per_row == 2 ? 208 : 200)) return -100;
  const size_t dyn = (size_t)n * n * sizeof(float);
  if (lanes_per_row == 2) {
    const int rows = (n + 15) & ~15;                      // whole warps of row pairs
    BASD_CUDA(cudaFuncSetAttribute(chreg::pivoted_cholesky_reg_kernel<13, 2>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    chreg::pivoted_cholesky_reg_kernel<13, 2><<<batch, rows * 2, dyn, st>>>(K, n, ld, stride_k, LT, ldl, stride_l,
                                                                         rel_tol, rank_out, dims);
  } else {
    const int rows = (n + 7) & ~7;                        // whole warps of row quads
    BASD_CUDA(cudaFuncSetAttribute(chreg::pivoted_cholesky_reg_kernel<13, 4>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    chreg::pivoted_cholesky_reg_kernel<13, 4><<<batch, rows * 4, dyn, st>>>(K, n, ld, stride_k, LT, ldl, stride_l,
                                                                         rel_tol, rank_out, dims);
  }
  BASD_LAUNCH_CHECK();
  return 0;
}

}  // namespace basd

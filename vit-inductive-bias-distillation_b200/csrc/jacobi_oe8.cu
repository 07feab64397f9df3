// One-sided Jacobi, register resident, EIGHT rows per 16-lane group.
//
// ncu on jacobi_rows_oddeven_kernel (two rows per 8-lane group) showed the odd steps of the
// odd-even transposition ordering bound by shared-memory wavefronts: every group parks a row,
// fetches its neighbour's, writes it back and reloads its own -- 4 row moves per row pair, 2,800
// wavefronts per odd step at N = 196, while the FMA pipe idles (fast rotations changed nothing).
// Here a group of 16 lanes owns 8 consecutive positions of the line: an even step rotates the
// four resident pairs (0,1)(2,3)(4,5)(6,7), an odd step rotates (1,2)(3,4)(5,6) locally and only
// (7, right neighbour's 0) crosses through shared memory, so the traffic per row drops 4x and the
// three local pairs are computed between the two barriers of the exchange.  Four independent
// pairs per thread also give the dot -> shuffle -> angle -> rotate chain instruction-level
// parallelism that two-row groups lack.
//
// Rows are kept as  row = scale * stored  (scaled rotations: two FMAs per element pair), squared
// norms are carried analytically with the determinant form for the shrinking row.
// Serves the same contract as basd_jacobi_rows (reference: torch.linalg.svd / svdvals /
// matrix_norm(ord="nuc"), layer_selector.py:92,99 and relational.py:48).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace basd {
namespace oe8 {

constexpr int R = 8;      // rows per group

// Floats per shared-memory row.  A warp of 16-lane groups touches two consecutive rows at once
// (lane gl of both groups reads column gl + 16 j): their banks differ only if the pitch is an odd
// multiple of 16.  ncu on the 384-column eigenproblems (pitch 384): 39 % of the shared-memory
// wavefronts were 2-way conflicts.
template <int G, int NF>
__host__ __device__ constexpr int row_pitch() { return (G == 16 && (NF * G) % 32 == 0) ? NF * G + 16 : NF * G; }

struct RowState { float n, d; };   // squared norm of the actual row, scale (actual = d * stored)

template <int NF>
__device__ __forceinline__ void fold(float (&r)[NF], RowState& s) {
#pragma unroll
  for (int e = 0; e < NF; ++e) r[e] *= s.d;
  s.d = 1.f;
}

template <int NF>
__device__ __forceinline__ float dot_local(const float (&x)[NF], const float (&y)[NF]) {
  float g0 = 0.f, g1 = 0.f;
#pragma unroll
  for (int e = 0; e + 1 < NF; e += 2) {
    g0 = fmaf(x[e], y[e], g0);
    g1 = fmaf(x[e + 1], y[e + 1], g1);
  }
  if (NF & 1) g0 = fmaf(x[NF - 1], y[NF - 1], g0);
  return g0 + g1;
}

// Rotation angle for the pair with (stored) dot product ga; updates norms / scales for the
// rows AFTER they trade places; returns the two stored-row multipliers.  Branch free: every lane
// evaluates the formulas, `valid` / `rot` only select the results.
__device__ __forceinline__ void angle(float ga, RowState& sx, RowState& sy, bool valid, float tol2,
                                      float zero_thr, float& worst, int& nrot, float& t1,
                                      float& t2) {
  ga *= sx.d * sy.d;
  const float nx = sx.n, ny = sy.n;
  const float gg = ga * ga, nxy = nx * ny;
  const bool rot = valid && (gg > tol2 * nxy) && nx > zero_thr && ny > zero_thr;
  nrot += rot ? 1 : 0;
  worst = fmaxf(worst, rot ? __fdividef(gg, nxy) : 0.f);
  const float d = ny - nx;
  const float h = fmaf(d, d, 4.f * gg);
  const float root = h * rsqrtf(fmaxf(h, 1e-37f));
  float t = __fdividef(2.f * fabsf(ga), fmaxf(fabsf(d) + root, 1e-37f));
  t = ((d < 0.f) != (ga < 0.f)) ? -t : t;
  t = rot ? t : 0.f;
  const float w2 = fmaf(t, t, 1.f);
  float c = rsqrtf(w2);
  c = c * fmaf(-0.5f * w2, c * c, 1.5f);
  const float dx = sx.d, dy = sy.d;
  t1 = t * __fdividef(dx, dy);
  t2 = t * __fdividef(dy, dx);
  t1 = rot ? t1 : 0.f;
  t2 = rot ? t2 : 0.f;
  // larger row grows by |t g|; the smaller one is det / larger (no cancellation, jacobi.cu)
  const float tg = t * ga;
  const bool ybig = d >= 0.f;
  const float big = ybig ? ny + tg : nx - tg;
  const float r = __fdividef(1.f, fmaxf(big, 1e-37f));
  const float small = fmaxf(fmaf(nx, ny * r, -(ga * r) * ga), 0.f);
  // rows trade places when the pair is valid (rotated or not)
  const float nxn = rot ? (ybig ? big : small) : ny;      // x will hold y'
  const float nyn = rot ? (ybig ? small : big) : nx;      // y will hold x'
  sx.n = valid ? nxn : nx;
  sy.n = valid ? nyn : ny;
  sx.d = valid ? c * dy : dx;
  sy.d = valid ? c * dx : dy;
}

// stored x <- stored y + t1 stored x ; stored y <- stored x - t2 stored y  (rows trade places)
// DESC walks the elements downwards.  Both outputs of an element need both inputs, so the first result
// lands in a spare register and the register file's view of the x row shifts by one register per pass;
// a second pass in the opposite direction shifts it back, so a loop body made of one ascending and one
// descending full step is a closed permutation (no register moves at the back edge: ptxas emitted ~50
// MOVs per half-step for the all-ascending body).
template <int NF, bool DESC = false>
__device__ __forceinline__ void apply(float (&x)[NF], float (&y)[NF], bool valid, float t1, float t2) {
  if (!valid) return;
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const int e = DESC ? NF - 1 - i : i;
    const float a = x[e], b = y[e];
    x[e] = fmaf(t1, a, b);
    y[e] = fmaf(-t2, b, a);
  }
}

// In-place pair in which the x row lives in shared memory (positions 0 of a group):
//   first pass: partial dot product, second pass (after the angle): rotate and trade places.
template <int G, int NF>
__device__ __forceinline__ float dot_smem(const float* __restrict__ xs_row, const float (&y)[NF]) {
  float g0 = 0.f, g1 = 0.f;
#pragma unroll
  for (int j = 0; j + 1 < NF; j += 2) {
    g0 = fmaf(xs_row[G * j], y[j], g0);
    g1 = fmaf(xs_row[G * (j + 1)], y[j + 1], g1);
  }
  if (NF & 1) g0 = fmaf(xs_row[G * (NF - 1)], y[NF - 1], g0);
  return g0 + g1;
}

// Sum four per-lane partials over the G lanes of a group with a transpose-reduce: the first
// two stages halve the number of live values (2 + 1 shuffles), the remaining stages carry one.
// Returns, on EVERY lane, the full sum of value (gl >> 1) & 3 -- exactly the pair a lane may own
// in the angle pass (owner lane p handles pair p >> 1) -- fetched with one last shuffle.
template <int G>
__device__ __forceinline__ float reduce4_owner(const float (&v)[4], int gl) {
  constexpr int H = G >> 1, Q = G >> 2;
  // stage 1 (xor H): lanes with bit H clear keep values {0,1}, the others {2,3}
  const bool up = (gl & H) != 0;
  const float s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];      // what this lane gives away
  const float k0 = up ? v[2] : v[0], k1 = up ? v[3] : v[1];      // what it keeps
  const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, H);
  const float a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, H);
  // stage 2 (xor Q): bit Q clear keeps the first of the two, set keeps the second
  const bool up2 = (gl & Q) != 0;
  float b = (up2 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up2 ? a0 : a1, Q);
#pragma unroll
  for (int o = Q >> 1; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
  // lane gl now holds the sum of value 2*[bit H] + [bit Q]; owner of pair k reads it from lane k*Q
  const int k = (gl >> 1) & 3;
  return __shfl_sync(0xffffffffu, b, ((k >> 1) * H) + ((k & 1) * Q), G);
}

__device__ __forceinline__ float sel4(const float (&v)[4], int k) {
  return k == 0 ? v[0] : (k == 1 ? v[1] : (k == 2 ? v[2] : v[3]));
}

// One angle pass of a step with parity ODD: lane p (p < 8, parity of p = ODD) owns the pair
// (p, p + 1); position 0 (even steps) and the right neighbour's position 0 (odd steps) keep
// their state in shared memory.
template <int G, int ODD, bool FULL = false>
__device__ __forceinline__ void angle_pass(float g_own, float& sn, float& sd, float2* xs,
                                           int slot, int right, int gl, int cnt, bool cross_ok,
                                           float tol2, float zero_thr, float& worst, int& nrot,
                                           float (&T1)[4], float (&T2)[4]) {
  const bool owner = gl < R && (gl & 1) == ODD;
  const bool from_smem_x = !ODD && gl == 0;            // x = position 0 (even steps)
  const bool from_smem_y = ODD && gl == R - 1;         // y = right neighbour's position 0 (odd steps)
  const float pn = __shfl_down_sync(0xffffffffu, sn, 1, G);
  const float pd = __shfl_down_sync(0xffffffffu, sd, 1, G);
  RowState sx{sn, sd}, sy{pn, pd};
  if (from_smem_x) { const float2 v = xs[slot]; sx.n = v.x; sx.d = v.y; }
  if (from_smem_y) { const float2 v = xs[right]; sy.n = v.x; sy.d = v.y; }
  const bool valid = FULL ? owner : (owner && (from_smem_y ? cross_ok : (gl + 1 < cnt)));
  float t1, t2;
  angle(g_own, sx, sy, valid, tol2, zero_thr, worst, nrot, t1, t2);
  // new states: x's position keeps sx, the partner position receives sy
  if (from_smem_x) xs[slot] = make_float2(sx.n, sx.d);
  else if (owner) { sn = sx.n; sd = sx.d; }
  if (from_smem_y && valid) xs[right] = make_float2(sy.n, sy.d);
  const float rn = __shfl_up_sync(0xffffffffu, sy.n, 1, G);
  const float rd = __shfl_up_sync(0xffffffffu, sy.d, 1, G);
  const bool receiver = gl >= 1 && gl < R && ((gl - 1) & 1) == ODD;
  if (receiver) { sn = rn; sd = rd; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    T1[k] = __shfl_sync(0xffffffffu, t1, 2 * k + ODD, G);
    T2[k] = __shfl_sync(0xffffffffu, t2, 2 * k + ODD, G);
  }
}

// The half-steps of one sweep.  FULL: every position of both groups of the warp exists and has a right
// neighbour (all but the last warp of a problem): the validity predicates are compile-time true, so
// the rotated rows trade places without the per-thread branches (and the register shuffling ptxas
// emits at their reconvergence points).  Both instantiations execute the same block-wide barriers.
template <int G, int NF, bool FULL, bool DESC>
__device__ __forceinline__ void full_step(float (&r)[R - 1][NF], float& sn, float& sd, float2* xs, int slot,
                                          int right, float* my_row, float* right_row, int gl, int cnt,
                                          bool cross_ok, bool has_odd, float tol2, float zero_thr, float& worst,
                                          int& nrot) {
  float ga[4], T1[4], T2[4];
  // ---------------- even step: (0,1) with position 0 in shared memory, (2,3) (4,5) (6,7)
  ga[0] = dot_smem<G, NF>(my_row, r[0]);
#pragma unroll
  for (int k = 1; k < 4; ++k) ga[k] = dot_local<NF>(r[2 * k - 1], r[2 * k]);
  angle_pass<G, 0, FULL>(reduce4_owner<G>(ga, gl), sn, sd, xs, slot, right, gl, cnt, cross_ok, tol2, zero_thr, worst,
                         nrot, T1, T2);
  if (FULL || 1 < cnt) {
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const float a = my_row[G * j], b = r[0][j];
      my_row[G * j] = fmaf(T1[0], a, b);
      r[0][j] = fmaf(-T2[0], b, a);
    }
  }
#pragma unroll
  for (int k = 1; k < 4; ++k) apply<NF, DESC>(r[2 * k - 1], r[2 * k], FULL || 2 * k + 1 < cnt, T1[k], T2[k]);
  __syncthreads();
  // ---------------- odd step: (1,2) (3,4) (5,6), (7, right neighbour's 0) through shared memory
  if (FULL || has_odd) {
    ga[3] = dot_smem<G, NF>(right_row, r[R - 2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) ga[k] = dot_local<NF>(r[2 * k], r[2 * k + 1]);
    angle_pass<G, 1, FULL>(reduce4_owner<G>(ga, gl), sn, sd, xs, slot, right, gl, cnt, cross_ok, tol2, zero_thr,
                           worst, nrot, T1, T2);
    if (FULL || cross_ok) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int j = DESC ? NF - 1 - i : i;
        const float a = r[R - 2][j], b = right_row[G * j];
        r[R - 2][j] = fmaf(T1[3], a, b);
        right_row[G * j] = fmaf(-T2[3], b, a);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) apply<NF, DESC>(r[2 * k], r[2 * k + 1], FULL || 2 * k + 2 < cnt, T1[k], T2[k]);
  }
  __syncthreads();
}

// The half-steps of one sweep.  FULL: every position of both groups of the warp exists and has a right
// neighbour and the row count is even (all but the last warp of a problem): the validity predicates are
// compile-time true.  Both instantiations execute the same block-wide barriers.
template <int G, int NF, bool FULL>
__device__ __forceinline__ void sweep_steps(float (&r)[R - 1][NF], float& sn, float& sd, float2* xs, int slot,
                                            int right, float* my_row, float* right_row, int gl, int cnt,
                                            bool cross_ok, int nn, float tol2, float zero_thr, float& worst,
                                            int& nrot) {
#pragma unroll 1
  for (int step = 0; step < nn; step += 4) {
    if ((step & 15) == 0 && step) {                  // fold the scales (they shrink by c per rotation)
      const float d0 = xs[slot].y;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NF; ++j) my_row[G * j] *= d0;
      if (gl == 0) xs[slot].y = 1.f;
#pragma unroll
      for (int i = 0; i < R - 1; ++i) {
        const float di = __shfl_sync(0xffffffffu, sd, i + 1, G);
#pragma unroll
        for (int e = 0; e < NF; ++e) r[i][e] *= di;
      }
      sd = 1.f;
      __syncwarp();
    }
    full_step<G, NF, FULL, false>(r, sn, sd, xs, slot, right, my_row, right_row, gl, cnt, cross_ok, step + 1 < nn,
                                  tol2, zero_thr, worst, nrot);
    if (step + 2 < nn)
      full_step<G, NF, FULL, true>(r, sn, sd, xs, slot, right, my_row, right_row, gl, cnt, cross_ok, step + 3 < nn,
                                   tol2, zero_thr, worst, nrot);
  }
}

// Row state (squared norm, scale) is DISTRIBUTED: lane p (1..7) of a group holds the state of
// position p, position 0's state lives in shared memory next to its row.  The four angles of a
// step are computed in ONE pass, each by the lane that owns the pair's left position (instead of
// four redundant evaluations on all 16 lanes); partner states travel by shuffles (width 16) and
// the two stored-row multipliers of every pair are broadcast back.
template <int G, int NF>    // G lanes per group; NF floats per lane per row: columns lane + G j, j < NF
__global__ void __launch_bounds__(G == 16 ? 512 : 896, 1)
jacobi_rows_oe8_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                       const int* __restrict__ dims, float tol, float stop2, int max_sweeps,
                       int* __restrict__ sweeps_out, int dim_lo, int dim_hi,
                       int* __restrict__ rot_out, int rows_only) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_scratch[32];
  const int prob = blockIdx.x, tid = threadIdx.x;
  // dims: active leading size of a square problem (rows and columns), or -- rows_only -- the
  // number of leading non-zero rows of a rank-deficient factor product (all columns active)
  if (dims && !rows_only && (dims[prob] < dim_lo || dims[prob] > dim_hi)) return;
  const int gid = tid / G, gl = tid % G;
  float* Gg = Gbase + (long)prob * stride;
  // rows_only: round the active row count up to even -- the extra (exactly zero) row is the "bye"
  // of the transposition ordering; an odd count measured one sweep more on full-rank problems
  // (not for square sub-problems: the rows trade places every step, so the zero row would end up
  // inside the leading k rows the caller reads back and a real one outside)
  const int nn = dims ? min(rows_only ? ((dims[prob] + 1) & ~1) : dims[prob], n) : n;
  const int mm = (dims && !rows_only) ? min(dims[prob], m) : m;
  const int groups = (nn + R - 1) / R;
  const int cnt = max(0, min(R, nn - gid * R));          // positions of this group that exist
  constexpr int PITCH = row_pitch<G, NF>();               // floats per shared-memory row
  const int nslots = blockDim.x / G + 1;
  float2* xs = reinterpret_cast<float2*>(smem + (size_t)nslots * PITCH);   // state of every position 0
  const int slot = min(gid, nslots - 1);
  const int right = min(gid + 1, nslots - 1);
  float* my_row = smem + (size_t)slot * PITCH + gl;      // position 0 of this group lives here
  float* right_row = smem + (size_t)right * PITCH + gl;  // position 0 of the right neighbour

  float r[R - 1][NF];                                    // r[i] = position i + 1
  float sn = 0.f, sd = 1.f;                              // state of position gl (lanes 1..7)
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    const int c = gl + G * j;
    my_row[G * j] = (cnt > 0 && c < mm) ? Gg[(long)(gid * R) * ld + c] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < R - 1; ++i) {
    const int row = gid * R + i + 1;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int c = gl + G * j;
      r[i][j] = (i + 1 < cnt && c < mm) ? Gg[(long)row * ld + c] : 0.f;
    }
  }
  const bool cross_ok = (cnt == R) && (gid + 1 < groups);   // right neighbour always owns a position 0
  const bool full_warp = __all_sync(0xffffffffu, cross_ok) && !(nn & 1);
  const float tol2 = tol * tol;
  int nrot = 0;
  int sweep = 0;
  for (; sweep < max_sweeps && nn >= 2; ++sweep) {
    // fold the scales into the rows, refresh the carried norms
    float nrm[R];
    {
      const float d0 = sweep ? xs[slot].y : 1.f;
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const float v = my_row[G * j] * d0;
        my_row[G * j] = v;
        a = fmaf(v, v, a);
      }
      nrm[0] = a;
    }
#pragma unroll
    for (int i = 0; i < R - 1; ++i) {
      const float di = __shfl_sync(0xffffffffu, sd, i + 1, G);
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < NF; ++e) {
        r[i][e] *= di;
        a = fmaf(r[i][e], r[i][e], a);
      }
      nrm[i + 1] = a;
    }
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < R; ++i) nrm[i] += __shfl_xor_sync(0xffffffffu, nrm[i], o);
    }
    sd = 1.f;
    sn = 0.f;
    float mxl = 0.f;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      mxl = fmaxf(mxl, nrm[i]);
      if (gl == i) sn = nrm[i];
    }
    if (gl == 0) xs[slot] = make_float2(nrm[0], 1.f);
    const float mx = block_max(mxl, red_scratch);        // (also orders the shared-memory writes above)
    const float zero_thr = 1e-14f * mx;
    float worst = 0.f;
    if (full_warp)
      sweep_steps<G, NF, true>(r, sn, sd, xs, slot, right, my_row, right_row, gl, cnt, cross_ok, nn, tol2, zero_thr,
                               worst, nrot);
    else
      sweep_steps<G, NF, false>(r, sn, sd, xs, slot, right, my_row, right_row, gl, cnt, cross_ok, nn, tol2, zero_thr,
                                worst, nrot);
    worst = block_max(worst, red_scratch);
    if (worst < stop2) { ++sweep; break; }
  }
  {
    const float d0 = (nn >= 2 && sweep) ? xs[slot].y : 1.f;
    if (cnt > 0) {
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int c = gl + G * j;
        if (c < mm) Gg[(long)(gid * R) * ld + c] = my_row[G * j] * d0;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R - 1; ++i) {
    const float di = __shfl_sync(0xffffffffu, sd, i + 1, G);
    const int row = gid * R + i + 1;
    if (i + 1 < cnt) {
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int c = gl + G * j;
        if (c < mm) Gg[(long)row * ld + c] = r[i][j] * di;
      }
    }
  }
  if (sweeps_out && tid == 0) sweeps_out[prob] = sweep;
  if (rot_out) {
    const float tot = block_sum((float)nrot, red_scratch);
    if (tid == 0) atomicAdd(rot_out + prob, (int)tot);
  }
}

// ---------------------------------------------------------------------------------------------
// Cluster variant (G = 16): the rows of ONE problem are spread over the CTAs of a thread-block
// cluster, gpc groups of eight rows per CTA (<= 12 groups = 192 threads, so a thread may hold
// seven 24-float row pieces in registers: 384 columns).  The only cross-CTA traffic is position 0
// of the first group of the next CTA: the last group of a CTA reads that row through distributed
// shared memory INTO REGISTERS at the start of the odd step -- the DSMEM latency then overlaps the
// three register-only pairs -- and writes it back after the rotation.  One hardware cluster
// barrier per step.  Serves the 384 x 384 selector eigenproblems (4 CTAs per problem).
template <int G, int ODD>
__device__ __forceinline__ void angle_pass_ptr(float g_own, float& sn, float& sd, float2* xs_my,
                                               float2* xs_right, int gl, int cnt, bool cross_ok,
                                               float tol2, float zero_thr, float& worst, int& nrot,
                                               float (&T1)[4], float (&T2)[4]) {
  const bool owner = gl < R && (gl & 1) == ODD;
  const bool from_smem_x = !ODD && gl == 0;
  const bool from_smem_y = ODD && gl == R - 1;
  const float pn = __shfl_down_sync(0xffffffffu, sn, 1, G);
  const float pd = __shfl_down_sync(0xffffffffu, sd, 1, G);
  RowState sx{sn, sd}, sy{pn, pd};
  if (from_smem_x) { const float2 v = *xs_my; sx.n = v.x; sx.d = v.y; }
  if (from_smem_y) { const float2 v = *xs_right; sy.n = v.x; sy.d = v.y; }
  const bool valid = owner && (from_smem_y ? cross_ok : (gl + 1 < cnt));
  float t1, t2;
  angle(g_own, sx, sy, valid, tol2, zero_thr, worst, nrot, t1, t2);
  if (from_smem_x) *xs_my = make_float2(sx.n, sx.d);
  else if (owner) { sn = sx.n; sd = sx.d; }
  if (from_smem_y && valid) *xs_right = make_float2(sy.n, sy.d);
  const float rn = __shfl_up_sync(0xffffffffu, sy.n, 1, G);
  const float rd = __shfl_up_sync(0xffffffffu, sy.d, 1, G);
  const bool receiver = gl >= 1 && gl < R && ((gl - 1) & 1) == ODD;
  if (receiver) { sn = rn; sd = rd; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    T1[k] = __shfl_sync(0xffffffffu, t1, 2 * k + ODD, G);
    T2[k] = __shfl_sync(0xffffffffu, t2, 2 * k + ODD, G);
  }
}

template <int G, int NF>
__device__ __forceinline__ void
jacobi_rows_oe8_cluster_body(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                             const int* __restrict__ dims, float tol, float stop2, int max_sweeps,
                             int* __restrict__ sweeps_out, int dim_lo, int dim_hi,
                             int* __restrict__ rot_out) {
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = cluster.num_blocks(), crank = cluster.block_rank();
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_scratch[32];
  __shared__ int cflag[2];
  const int prob = blockIdx.x / csize, tid = threadIdx.x;
  if (dims && (dims[prob] < dim_lo || dims[prob] > dim_hi)) return;            // whole cluster exits together
  const int gpc = blockDim.x / G;
  const int lgid = tid / G, gl = tid % G;
  const int gid = crank * gpc + lgid;
  float* Gg = Gbase + (long)prob * stride;
  const int nn = dims ? min(dims[prob], n) : n;
  const int mm = dims ? min(dims[prob], m) : m;
  const int groups = (nn + R - 1) / R;
  const int cnt = max(0, min(R, nn - gid * R));
  constexpr int PITCH = row_pitch<G, NF>();
  const int nslots = gpc + 1;                            // + a spare nobody pairs with
  float2* xs = reinterpret_cast<float2*>(smem + (size_t)nslots * PITCH);
  float* my_row = smem + (size_t)lgid * PITCH + gl;
  float2* xs_my = xs + lgid;
  float* right_row = smem + (size_t)(lgid + 1) * PITCH + gl;
  float2* xs_right = xs + lgid + 1;
  if (lgid + 1 == gpc && crank + 1 < csize) {            // first group of the next CTA (DSMEM)
    right_row = cluster.map_shared_rank(smem, crank + 1) + gl;
    xs_right = cluster.map_shared_rank(xs, crank + 1);
  }
  int* flag0 = cluster.map_shared_rank(cflag, 0);
  auto wide_max = [&](float v, int which) {
    v = block_max(v, red_scratch);
    if (tid == 0) atomicMax(flag0 + which, __float_as_int(v));     // non-negative floats order as ints
    cluster.sync();
    return __int_as_float(flag0[which]);
  };

  float r[R - 1][NF];
  float sn = 0.f, sd = 1.f;
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    const int c = gl + G * j;
    my_row[G * j] = (cnt > 0 && c < mm) ? Gg[(long)(gid * R) * ld + c] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < R - 1; ++i) {
    const int row = gid * R + i + 1;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int c = gl + G * j;
      r[i][j] = (i + 1 < cnt && c < mm) ? Gg[(long)row * ld + c] : 0.f;
    }
  }
  const bool cross_ok = (cnt == R) && (gid + 1 < groups);
  const float tol2 = tol * tol;
  int nrot = 0;
  int sweep = 0;
  for (; sweep < max_sweeps && nn >= 2; ++sweep) {
    if (crank == 0 && tid == 0) { cflag[0] = 0; cflag[1] = 0; }
    cluster.sync();
    float nrm[R];
    {
      const float d0 = sweep ? xs_my->y : 1.f;
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const float v = my_row[G * j] * d0;
        my_row[G * j] = v;
        a = fmaf(v, v, a);
      }
      nrm[0] = a;
    }
#pragma unroll
    for (int i = 0; i < R - 1; ++i) {
      const float di = __shfl_sync(0xffffffffu, sd, i + 1, G);
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < NF; ++e) {
        r[i][e] *= di;
        a = fmaf(r[i][e], r[i][e], a);
      }
      nrm[i + 1] = a;
    }
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < R; ++i) nrm[i] += __shfl_xor_sync(0xffffffffu, nrm[i], o);
    }
    sd = 1.f;
    sn = 0.f;
    float mxl = 0.f;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      mxl = fmaxf(mxl, nrm[i]);
      if (gl == i) sn = nrm[i];
    }
    if (gl == 0) *xs_my = make_float2(nrm[0], 1.f);
    const float mx = wide_max(mxl, 0);                   // (its cluster barrier orders the writes above)
    const float zero_thr = 1e-14f * mx;
    float worst = 0.f;
    for (int step = 0; step < nn; step += 2) {
      float ga[4], T1[4], T2[4], x0[NF];
      // ---------------- even step: position 0 comes into registers once, goes back once
      const bool fold_now = (step & 15) == 0 && step;
      const float d0 = fold_now ? xs_my->y : 1.f;
      if (fold_now) __syncwarp();
#pragma unroll
      for (int j = 0; j < NF; ++j) x0[j] = my_row[G * j] * d0;
      if (fold_now) {
        if (gl == 0) xs_my->y = 1.f;
#pragma unroll
        for (int i = 0; i < R - 1; ++i) {
          const float di = __shfl_sync(0xffffffffu, sd, i + 1, G);
#pragma unroll
          for (int e = 0; e < NF; ++e) r[i][e] *= di;
        }
        sd = 1.f;
        __syncwarp();
      }
      ga[0] = dot_local<NF>(x0, r[0]);
#pragma unroll
      for (int k = 1; k < 4; ++k) ga[k] = dot_local<NF>(r[2 * k - 1], r[2 * k]);
      angle_pass_ptr<G, 0>(reduce4_owner<G>(ga, gl), sn, sd, xs_my, xs_right, gl, cnt, cross_ok, tol2, zero_thr,
                        worst, nrot, T1, T2);
      apply<NF>(x0, r[0], 1 < cnt, T1[0], T2[0]);
#pragma unroll
      for (int j = 0; j < NF; ++j) my_row[G * j] = x0[j];
      // Split-phase cluster barrier: everything other groups read (position 0 and its state) is written,
      // so ARRIVE now and run the register-only work -- the three remaining rotations and the three local
      // dot products of the odd step -- under the barrier's ~380-cycle latency before WAITING.
      cluster.barrier_arrive();
#pragma unroll
      for (int k = 1; k < 4; ++k) apply<NF>(r[2 * k - 1], r[2 * k], 2 * k + 1 < cnt, T1[k], T2[k]);
      // ---------------- odd step: the neighbour's position 0 (possibly remote) into registers first
      if (step + 1 < nn) {
#pragma unroll
        for (int k = 0; k < 3; ++k) ga[k] = dot_local<NF>(r[2 * k], r[2 * k + 1]);
        cluster.barrier_wait();
#pragma unroll
        for (int j = 0; j < NF; ++j) x0[j] = right_row[G * j];
        ga[3] = dot_local<NF>(r[R - 2], x0);
        angle_pass_ptr<G, 1>(reduce4_owner<G>(ga, gl), sn, sd, xs_my, xs_right, gl, cnt, cross_ok, tol2, zero_thr,
                          worst, nrot, T1, T2);
        if (cross_ok) {
          apply<NF>(r[R - 2], x0, true, T1[3], T2[3]);
#pragma unroll
          for (int j = 0; j < NF; ++j) right_row[G * j] = x0[j];
        }
        cluster.barrier_arrive();
#pragma unroll
        for (int k = 0; k < 3; ++k) apply<NF>(r[2 * k], r[2 * k + 1], 2 * k + 2 < cnt, T1[k], T2[k]);
      }
      cluster.barrier_wait();
    }
    const float all_worst = wide_max(worst, 1);
    cluster.sync();                                      // everyone has read the flags before rank 0 resets them
    if (all_worst < stop2) { ++sweep; break; }
  }
  {
    const float d0 = (nn >= 2 && sweep) ? xs_my->y : 1.f;
    if (cnt > 0) {
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int c = gl + G * j;
        if (c < mm) Gg[(long)(gid * R) * ld + c] = my_row[G * j] * d0;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R - 1; ++i) {
    const float di = __shfl_sync(0xffffffffu, sd, i + 1, G);
    const int row = gid * R + i + 1;
    if (i + 1 < cnt) {
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int c = gl + G * j;
        if (c < mm) Gg[(long)row * ld + c] = r[i][j] * di;
      }
    }
  }
  if (sweeps_out && crank == 0 && tid == 0) sweeps_out[prob] = sweep;
  if (rot_out) {
    const float tot = block_sum((float)nrot, red_scratch);
    if (tid == 0) atomicAdd(rot_out + prob, (int)tot);
  }
}

template <int G, int NF>
__global__ void __launch_bounds__(192, 1)
jacobi_rows_oe8_cluster_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                               const int* __restrict__ dims, float tol, float stop2, int max_sweeps,
                               int* __restrict__ sweeps_out, int dim_lo, int dim_hi,
                               int* __restrict__ rot_out) {
  jacobi_rows_oe8_cluster_body<G, NF>(Gbase, n, m, ld, stride, dims, tol, stop2, max_sweeps, sweeps_out, dim_lo, dim_hi,
                                      rot_out);
}

// The SAME cluster sweep for small problems (<= 256 rows, <= 208 columns), each split over 4 CTAs of at
// most 128 threads so that a launch with FEW problems spreads over four times the SMs (the k x k
// principal-angle SVDs, 48 problems at C2: 2.21 -> 1.64 ms on B200).  Measured and rejected for full waves
// of problems (1,024 x 196^2: 19.4 ms with quarters, 21.3 ms with halves, 16.1 ms on the single-CTA
// kernel: the cluster barrier and the DSMEM hop of the boundary row cost more than the independent
// barriers of co-resident CTAs gain); results are bitwise equal to the single-CTA kernel's.
template <int G, int NF, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
jacobi_rows_oe8_split_kernel(float* __restrict__ Gbase, int n, int m, int ld, long stride,
                             const int* __restrict__ dims, float tol, float stop2, int max_sweeps,
                             int* __restrict__ sweeps_out, int dim_lo, int dim_hi,
                             int* __restrict__ rot_out) {
  jacobi_rows_oe8_cluster_body<G, NF>(Gbase, n, m, ld, stride, dims, tol, stop2, max_sweeps, sweeps_out, dim_lo, dim_hi,
                                      rot_out);
}

// MAXT / MINB: 128 threads x 3 CTAs per SM (<= 8 groups per CTA: 168 registers, no spills at 13 floats
// per row piece).
template <int NF, int MAXT, int MINB>
static int launch_split(float* Gm, int n, int m, int ld, long stride, int batch, const int* dims, float tol, float stop2,
                        int max_sweeps, int* sweeps_out, cudaStream_t st, int* rot_out, int csize, int dim_lo,
                        int dim_hi) {
  constexpr int G = 16;
  const int cap = (dims && dim_hi < n) ? dim_hi : n;          // device-side sizes: problems outside the window exit
  const int groups = (cap + R - 1) / R;
  int gpc = (groups + csize - 1) / csize;
  gpc = (gpc + 1) & ~1;                                   // whole warps
  const int threads = gpc * G;
  if (threads > MAXT || threads < 32) return -100;
  const size_t nslots = gpc + 1;
  const size_t dyn = (nslots * row_pitch<G, NF>() + 2 * nslots + 4) * sizeof(float);
  auto kernel = jacobi_rows_oe8_split_kernel<G, NF, MAXT, MINB>;
  BASD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * csize);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BASD_CUDA(cudaLaunchKernelEx(&cfg, kernel, Gm, n, m, ld, stride, dims, tol, stop2, max_sweeps, sweeps_out, dim_lo, dim_hi,
                               rot_out));
  return 0;
}

template <int G, int NF>
static int launch_cluster(float* Gm, int n, int m, int ld, long stride, int batch, const int* dims,
                          float tol, float stop2, int max_sweeps, int* sweeps_out, cudaStream_t st, int dim_lo,
                          int dim_hi, int* rot_out) {
  int GPC_MAX = 192 / G;                                  // 192 threads: up to 255 registers each
  const int cap = (dims && dim_hi < n) ? dim_hi : n;
  const int groups = (cap + R - 1) / R;
  // 16-lane groups: portable clusters (<= 8 CTAs).  32-lane groups (rows up to 768 floats): 6 groups
  // per CTA, so 768 rows need the non-portable cluster size 16 (one such cluster per GPC).
  const int cmax = (G == 32) ? 16 : 8;
  // while the whole launch still fits one wave, half as many groups per CTA on twice the CTAs shortens
  // the step (three warps per SM instead of six: 16 x 384^2 eigenproblems 5.51 -> 5.36 ms)
  if (G == 16) {
    int c2 = 1;
    while (c2 < cmax && c2 * (GPC_MAX / 2) < groups) c2 <<= 1;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (c2 * (GPC_MAX / 2) >= groups && (long)batch * c2 <= sms) GPC_MAX /= 2;
  }
  int csize = 1;
  while (csize < cmax && csize * GPC_MAX < groups) csize <<= 1;
  if (csize * GPC_MAX < groups) return -100;
  int gpc = (groups + csize - 1) / csize;
  if (G == 16) gpc = (gpc + 1) & ~1;                      // whole warps
  const int threads = gpc * G;
  const size_t nslots = gpc + 1;
  const size_t dyn = (nslots * row_pitch<G, NF>() + 2 * nslots + 4) * sizeof(float);
  BASD_CUDA(cudaFuncSetAttribute(jacobi_rows_oe8_cluster_kernel<G, NF>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  if (csize > 8)
    BASD_CUDA(cudaFuncSetAttribute(jacobi_rows_oe8_cluster_kernel<G, NF>,
                                   cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * csize);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BASD_CUDA(cudaLaunchKernelEx(&cfg, jacobi_rows_oe8_cluster_kernel<G, NF>, Gm, n, m, ld, stride, dims, tol, stop2,
                               max_sweeps, sweeps_out, dim_lo, dim_hi, rot_out));
  return 0;
}

template <int G, int NF>
static int launch(float* Gm, int n, int m, int ld, long stride, int batch, const int* dims, float tol, float stop2,
                  int max_sweeps, int* sweeps_out, cudaStream_t st, int dim_lo, int dim_hi,
                  int* rot_out, int rows_only) {
  const int cap = (dims && !rows_only && dim_hi < n) ? dim_hi : n;
  int threads = ((cap + R - 1) / R) * G;
  threads = (threads + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  const size_t nslots = threads / G + 1;
  const size_t dyn = (nslots * row_pitch<G, NF>() + 2 * nslots + 4) * sizeof(float);
  BASD_CUDA(cudaFuncSetAttribute(jacobi_rows_oe8_kernel<G, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)dyn));
  jacobi_rows_oe8_kernel<G, NF><<<batch, threads, dyn, st>>>(Gm, n, m, ld, stride, dims, tol, stop2, max_sweeps,
                                                          sweeps_out, dim_lo, dim_hi, rot_out, rows_only);
  BASD_LAUNCH_CHECK();
  return 0;
}

}  // namespace oe8

// Problems with at most 256 active rows and 256 active columns (the per-sample Procrustes SVDs
// and the k x k principal-angle SVDs).  Returns -100 when the shape does not fit.
int launch_jacobi_oe8(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                      float tol, float stop2, int max_sweeps, int* sweeps_out, cudaStream_t st, int dim_lo,
                      int dim_hi, int* rot_out, int rows_only) {
  const int cap_n = (dims && !rows_only && dim_hi < n) ? dim_hi : n;
  const int cap_m = (dims && !rows_only && dim_hi < m) ? dim_hi : m;
  if (cap_n > 256 || cap_m > 256) return -100;          // 16 warps x 128 registers per CTA
#define BASD_OE8(NF) \
  return oe8::launch<16, NF>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, dim_lo, dim_hi, rot_out, rows_only)
  if (cap_m <= 64) BASD_OE8(4);
  if (cap_m <= 96) BASD_OE8(6);
  if (cap_m <= 128) BASD_OE8(8);
  if (cap_m <= 160) BASD_OE8(10);
  if (cap_m <= 192) BASD_OE8(12);
  if (cap_m <= 208) BASD_OE8(13);
  if (cap_m <= 224) BASD_OE8(14);
  BASD_OE8(16);
#undef BASD_OE8
}

// Split of the small problems over 4 CTAs (see jacobi_rows_oe8_split_kernel), for launches with few
// problems: full problems (dims == null) or square problems with a device-side active size inside
// [dim_lo, dim_hi].  Returns -100 when the shape does not fit.
int launch_jacobi_oe8_split(float* G, int n, int m, int ld, long stride, int batch, const int* dims, float tol, float stop2,
                            int max_sweeps, int* sweeps_out, cudaStream_t st, int* rot_out, int csize, int dim_lo,
                            int dim_hi, int rows_only) {
  const int cap_n = (dims && dim_hi < n) ? dim_hi : n;
  const int cap_m = (dims && dim_hi < m) ? dim_hi : m;
  if (cap_n > 256 || cap_m > 208 || csize != 4 || rows_only) return -100;
#define BASD_OE8S(NF)                                                                                         \
  return oe8::launch_split<NF, 128, 3>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, rot_out, \
                                       csize, dim_lo, dim_hi)
  if (cap_m <= 128) BASD_OE8S(8);
  if (cap_m <= 192) BASD_OE8S(12);
  BASD_OE8S(13);
#undef BASD_OE8S
}

// Cluster variant: up to 768 active rows; up to 384 active columns with 16-lane groups (portable
// clusters), up to 768 with 32-lane groups (cluster size 16 for more than 384 rows).
// Returns -100 when the shape does not fit.
int launch_jacobi_oe8_cluster(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                              float tol, float stop2, int max_sweeps, int* sweeps_out, cudaStream_t st, int dim_lo,
                              int dim_hi, int* rot_out) {
  const int cap_n = (dims && dim_hi < n) ? dim_hi : n;
  const int cap_m = (dims && dim_hi < m) ? dim_hi : m;
  if (cap_n > 768 || cap_m > 768) return -100;
#define BASD_OE8C(GG, NF) \
  return oe8::launch_cluster<GG, NF>(G, n, m, ld, stride, batch, dims, tol, stop2, max_sweeps, sweeps_out, st, dim_lo, dim_hi, rot_out)
  // (321..384 columns on 32-lane groups with 12-float row pieces over 8-CTA clusters measured
  //  slower: 5.85 vs 4.55 ms for the 16 x 384^2 eigenproblems of C2)
  if (cap_m <= 256) BASD_OE8C(16, 16);
  if (cap_m <= 320) BASD_OE8C(16, 20);
  if (cap_m <= 384) BASD_OE8C(16, 24);
  if (cap_m <= 512) BASD_OE8C(32, 16);
  if (cap_m <= 640) BASD_OE8C(32, 20);
  BASD_OE8C(32, 24);
#undef BASD_OE8C
}

}  // namespace basd

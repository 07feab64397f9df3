// Batched FP32-accurate GEMM on the 5th-generation tensor cores:  C = alpha * op(A) op(B)
// with every fp32 operand split into two TF32 terms (x = hi + lo, both rounded to nearest) and
// three tcgen05.mma.kind::tf32 products per K step (lo*hi + hi*lo + hi*hi, fp32 accumulation in
// TMEM) -- the "3xTF32" scheme: the dropped lo*lo term and the rounding of lo are both
// ~2^-22 |a||b|, i.e. fp32-level products, so the per-sample Grams and factor products that feed
// rank decisions keep their noise floor while running on the tensor pipe instead of FFMA.
//
// The matrices of this path are small (N_tokens x N_tokens x {N_tokens, D}), batched by sample, so
// the kernel is bandwidth bound: one CTA owns a 128 x BN output tile of one problem and walks K.
//   warps 0..7 : producers.  Coalesced 128-bit global loads -> hi/lo split in registers ->
//                shared memory in the canonical K-major SWIZZLE_128B UMMA layout (rows of 32 fp32
//                = 128 B, 16-byte chunks XOR-swizzled by row % 8).  Operands whose contraction index
//                is the slow dimension in memory are transposed by the store pattern
//                (bank-conflict-free: 16 k x 2 column quads per warp), so one descriptor type
//                serves all four op() combinations.
//   warp 8     : one thread issues tcgen05.mma (M = 128, N = BN, K = 8), tcgen05.commit frees the
//                stage (one stage per CTA, two CTAs per SM).
//   warps 0..7 : epilogue, tcgen05.ld 32x32b -> registers -> alpha -> global (warps w and w + 4 share a
//                TMEM lane quarter and split its columns).
// Replaces the SIMT batched SGEMM (gemm_simt.cu) for the Procrustes products
// (reference: torch.bmm at relational.py:47 and the matmuls inside linalg.svd's backward).
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace basd {
namespace tc3 {

constexpr int TM = 128;                   // UMMA M
constexpr int KS = 32;                    // fp32 per K slab = one 128-byte swizzle row
constexpr int PRODUCER_WARPS = 8;
constexpr int PRODUCERS = PRODUCER_WARPS * 32;
constexpr int THREADS = PRODUCERS + 32;
constexpr int A_BYTES = TM * 128;         // one split term of the A slab
constexpr int TMEM_COLS = 256;
constexpr int SMEM_LIMIT = 227 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TC3_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TC3_DONE;\n"
      "bra TC3_WAIT;\n"
      "TC3_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: start address >> 4 in bits [0,14),
// LBO (unused for swizzled K-major) = 1 at [16,30), SBO = 1024 B (8 rows x 128 B) at [32,46),
// version 1 at [46,48), layout type 2 (SWIZZLE_128B) at [61,64).  A K step of 8 fp32 advances
// the start address by 32 bytes inside the swizzle row.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((1024 >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// x = hi + lo with hi = tf32_rn(x), lo = tf32_rn(x - hi)  (low 13 mantissa bits cleared, so the
// tensor core's own operand truncation is a no-op).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  const uint32_t u = __float_as_uint(x);
  hi = __uint_as_float((u + 0x1000u) & 0xffffe000u);
  const float r = x - hi;
  lo = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x);
  split_tf32(v.y, hi.y, lo.y);
  split_tf32(v.z, hi.z, lo.z);
  split_tf32(v.w, hi.w, lo.w);
}

// Stores through 32-bit shared-window addresses: the tile base is carved out of the dynamic shared memory by
// integer alignment, which makes every pointer derived from it GENERIC for the compiler -- the first version's
// producers issued ST.E with 64-bit address arithmetic (an IADD3 / IADD3.X pair per store) instead of STS with
// an immediate offset.
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B slab (k in [0,32))
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
  return static_cast<uint32_t>(((row >> 3) << 10) + ((row & 7) << 7) +
                               ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}

// The global loads of slab kb+1 are issued right after slab kb has been split and written to shared memory,
// so they are in flight while the tensor core works on slab kb; the second CTA of the SM covers the rest.
//
// Everything about a producer thread's items that does not depend on the slab -- shared-memory offsets,
// global row pointers, which items exist and which rows are inside the matrix -- is computed ONCE per tile
// (a "plan"): item u of a thread sits 4,096 bytes (8,192 for bf16) behind item u - 1 in shared memory and a
// fixed stride behind it in global memory, so the per-slab code is loads, the split and stores at immediate
// offsets.  (ncu on the first version, which recomputed the swizzled offset of every element in every slab:
// 1,080 instructions per warp per slab, 84 k warp instructions per 128 x 256 x 196 tile -- the producers'
// issue slots, not the tensor pipe (21 % active), bounded the tile.)
struct Plan {
  const float* g;        // global address of item 0 at k0 = 0
  int gstep;             // elements between consecutive items (k-contiguous operands) / the pitch (transposed)
  uint32_t soff[4];      // shared-memory byte offsets of item 0 (k-contiguous: [0] only; transposed: 4 rows)
  uint32_t exists;       // bit u: item u lies inside the slab (is stored)
  uint32_t inside;       // bit u: item u's rows lie inside the matrix (is loaded; zero otherwise)
  int k;                 // k offset of this thread inside the slab
};

// Operand whose contraction index is contiguous in memory (row-major R x K with pitch ld):
// one float4 = 4 consecutive k of one row; a quarter warp covers one 128-byte row -> conflict-free
// 128-bit stores.  Item f of a thread: f = ptid + u * PRODUCERS, row = f / 8, chunk = f % 8.
template <int CNT>
__device__ __forceinline__ Plan plan_kcontig(const float* __restrict__ g, int ld, int r0, int R, int row_limit,
                                             int ptid) {
  Plan pl;
  const int row = ptid >> 3, ch = ptid & 7;
  pl.k = ch * 4;
  pl.g = g + static_cast<long>(r0 + row) * ld + pl.k;
  pl.gstep = (PRODUCERS / 8) * ld;
  pl.soff[0] = sw_off(row, pl.k);
  pl.soff[1] = pl.soff[2] = pl.soff[3] = 0;
  pl.exists = pl.inside = 0;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const int ru = row + u * (PRODUCERS / 8);
    if (ru < R) pl.exists |= 1u << u;
    if (ru < R && r0 + ru < row_limit) pl.inside |= 1u << u;
  }
  return pl;
}
template <int CNT>
__device__ __forceinline__ void issue_kcontig(float4 (&v)[CNT], const Plan& pl, int k0, int K) {
  const bool kin = k0 + pl.k < K;
  const float* g = pl.g + k0;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kin && ((pl.inside >> u) & 1u)) v[u] = __ldg(reinterpret_cast<const float4*>(g + static_cast<long>(u) * pl.gstep));
  }
}
template <int CNT>
__device__ __forceinline__ void store_kcontig(const float4 (&v)[CNT], const Plan& pl, uint32_t hi, uint32_t lo) {
  hi += pl.soff[0];
  lo += pl.soff[0];
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    if ((pl.exists >> u) & 1u) {
      float4 h, l;
      split4(v[u], h, l);
      sts128(hi + u * 4096, h);                          // 32 rows further: four 1,024-byte swizzle atoms
      sts128(lo + u * 4096, l);
    }
  }
}

// Operand whose contraction index is the slow dimension in memory (row-major K x R with pitch
// ld): one float4 = 4 consecutive rows of the slab at one k; transposed by scalar stores.  A warp
// covers 16 k x 2 row-quads: the 32 scalar stores of each of the 4 components hit 32 banks.
// Block blk of a warp: blk = warp + u * PRODUCER_WARPS over (row octets) x (two halves of the slab).
template <int CNT>
__device__ __forceinline__ Plan plan_mncontig(const float* __restrict__ g, int ld, int r0, int R, int row_limit,
                                              int ptid) {
  Plan pl;
  const int lane = ptid & 31, warp = ptid >> 5;
  const int kk = lane & 15, ql = lane >> 4;
  pl.k = (warp & 1) * 16 + kk;
  const int row = (warp >> 1) * 8 + ql * 4;              // item u: row + 32 u (blk >> 1 = (warp >> 1) + 4 u)
  pl.g = g + static_cast<long>(pl.k) * ld + r0 + row;
  pl.gstep = ld;
#pragma unroll
  for (int i = 0; i < 4; ++i) pl.soff[i] = sw_off(row + i, pl.k);
  pl.exists = pl.inside = 0;
  const int blocks = (R >> 3) * 2;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const int blk = warp + u * PRODUCER_WARPS;
    if (blk < blocks) pl.exists |= 1u << u;
    if (blk < blocks && r0 + row + 32 * u < row_limit) pl.inside |= 1u << u;
  }
  return pl;
}
template <int CNT>
__device__ __forceinline__ void issue_mncontig(float4 (&v)[CNT], const Plan& pl, int k0, int K) {
  const bool kin = k0 + pl.k < K;
  const float* g = pl.g + static_cast<long>(k0) * pl.gstep;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kin && ((pl.inside >> u) & 1u)) v[u] = __ldg(reinterpret_cast<const float4*>(g + 32 * u));
  }
}
template <int CNT>
__device__ __forceinline__ void store_mncontig(const float4 (&v)[CNT], const Plan& pl, uint32_t hi, uint32_t lo) {
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    if ((pl.exists >> u) & 1u) {
      float4 h, l;
      split4(v[u], h, l);
      const float hv[4] = {h.x, h.y, h.z, h.w}, lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sts32(hi + pl.soff[i] + u * 4096, hv[i]);
        sts32(lo + pl.soff[i] + u * 4096, lv[i]);
      }
    }
  }
}

// bf16 A operand (row-major R x K, K contiguous): one 128-bit load = 8 consecutive k of one row;
// bf16 is a subset of TF32, so hi is exact, lo is zero and its MMA is skipped.
// Item f = ptid + u * PRODUCERS (u < 2): row = f / 4, octet = f % 4.
__device__ __forceinline__ Plan plan_kcontig_bf16(const __nv_bfloat16* __restrict__ g, int ld, int r0,
                                                  int row_limit, int ptid) {
  Plan pl;
  const int row = ptid >> 2, oc = ptid & 3;
  pl.k = oc * 8;
  pl.g = reinterpret_cast<const float*>(g + static_cast<long>(r0 + row) * ld + pl.k);
  pl.gstep = (PRODUCERS / 4) * ld;                       // in bf16 elements
  pl.soff[0] = sw_off(row, pl.k);
  pl.soff[1] = sw_off(row, pl.k + 4);
  pl.soff[2] = pl.soff[3] = 0;
  pl.exists = 3u;
  pl.inside = (r0 + row < row_limit ? 1u : 0u) | (r0 + row + PRODUCERS / 4 < row_limit ? 2u : 0u);
  return pl;
}
__device__ __forceinline__ void issue_kcontig_bf16(float4 (&v)[4], const Plan& pl, int k0, int K) {
  const bool kin = k0 + pl.k < K;
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(pl.g) + k0;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (kin && ((pl.inside >> u) & 1u)) raw = __ldg(reinterpret_cast<const uint4*>(g + static_cast<long>(u) * pl.gstep));
    v[u] = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z),
                       __uint_as_float(raw.w));
  }
}
__device__ __forceinline__ void store_kcontig_bf16(const float4 (&v)[4], const Plan& pl, uint32_t hi) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t w[4] = {__float_as_uint(v[u].x), __float_as_uint(v[u].y), __float_as_uint(v[u].z),
                           __float_as_uint(v[u].w)};
    const float4 lo4 = make_float4(__uint_as_float(w[0] << 16), __uint_as_float(w[0] & 0xffff0000u),
                                   __uint_as_float(w[1] << 16), __uint_as_float(w[1] & 0xffff0000u));
    const float4 hi4 = make_float4(__uint_as_float(w[2] << 16), __uint_as_float(w[2] & 0xffff0000u),
                                   __uint_as_float(w[3] << 16), __uint_as_float(w[3] & 0xffff0000u));
    sts128(hi + pl.soff[0] + u * 8192, lo4);                           // k .. k+3     (64 rows further)
    sts128(hi + pl.soff[1] + u * 8192, hi4);                           // k+4 .. k+7
  }
}

constexpr int A_ITEMS = TM * 8 / PRODUCERS;             // 4 float4 per producer thread
constexpr int B_ITEMS = 256 * 8 / PRODUCERS;            // 8 (BN <= 256)

struct Params {
  const float* A; const float* B; float* C;
  int M, N, K, lda, ldb, ldc;
  long sa, sb, sc;
  int ta, tb;             // 1: the operand is stored with its contraction index as the slow dimension
  int BN, tiles_n, stages;
  float alpha; const float* alpha_dev;
  int a_bf16;              // A is bf16 (ta = 0 only): exact in TF32, no lo term
  const float* col_sub;    // optional (N): C = alpha (acc - col_sub[col])  -- (A - 1 mu^T) B with col_sub = mu^T B
  int c_bf16;              // store C as bf16
};

// One shared-memory stage and one register set of raw tiles per CTA, compiled for TWO CTAs per SM (96
// registers, <= 97 KB of shared memory, 256 TMEM columns each).  A tile of these batched products is short
// (K = 196: seven slabs, then an epilogue nothing overlaps), so one pipelined CTA per SM (round 1: two or
// three stages, two register sets, 168 registers) left the SM idle through every prologue, load latency and
// epilogue: ncu showed tensor pipe 20 %, DRAM 14 %, issue slots 33 % -- nothing saturated.  Two resident
// CTAs fill each other's gaps: 5.97 -> 5.01 ms for the 30 launches of a C2 step, measured on B200.
__global__ void __launch_bounds__(THREADS, 2)
gemm_tc3_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_bytes = p.BN * 128;
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tmem_full_bar = empty_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x / p.tiles_n, tile_n = blockIdx.x - tile_m * p.tiles_n;
  const int m0 = tile_m * TM, n0 = tile_n * p.BN;
  const long prob = blockIdx.y;
  const float* A = p.A + prob * p.sa;
  const float* B = p.B + prob * p.sb;
  float* C = p.C + prob * p.sc;
  const int num_kb = (p.K + KS - 1) / KS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], PRODUCER_WARPS); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PRODUCER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PRODUCER_WARPS) {
    // ===== producers =====
    const int ptid = threadIdx.x;
    float4 va[1][A_ITEMS], vb[1][B_ITEMS];
    Plan pa, pb;
    if (p.a_bf16) pa = plan_kcontig_bf16(reinterpret_cast<const __nv_bfloat16*>(p.A) + prob * p.sa, p.lda, m0, p.M, ptid);
    else if (p.ta) pa = plan_mncontig<A_ITEMS>(A, p.lda, m0, TM, p.M, ptid);
    else      pa = plan_kcontig<A_ITEMS>(A, p.lda, m0, TM, p.M, ptid);
    if (p.tb) pb = plan_kcontig<B_ITEMS>(B, p.ldb, n0, p.BN, p.N, ptid);
    else      pb = plan_mncontig<B_ITEMS>(B, p.ldb, n0, p.BN, p.N, ptid);
    auto issue = [&](int set, int kb) {
      const int k0 = kb * KS;
      if (p.a_bf16) issue_kcontig_bf16(va[set], pa, k0, p.K);
      else if (p.ta) issue_mncontig<A_ITEMS>(va[set], pa, k0, p.K);
      else      issue_kcontig<A_ITEMS>(va[set], pa, k0, p.K);
      if (p.tb) issue_kcontig<B_ITEMS>(vb[set], pb, k0, p.K);
      else      issue_mncontig<B_ITEMS>(vb[set], pb, k0, p.K);
    };
    auto publish = [&](int set, int kb) {
      const int s = kb % p.stages;
      const uint32_t ph = (kb / p.stages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      const uint32_t a_hi = smem_u32(smem) + s * stage_bytes;
      const uint32_t a_lo = a_hi + A_BYTES;
      const uint32_t b_hi = a_lo + A_BYTES;
      const uint32_t b_lo = b_hi + b_bytes;
      if (p.a_bf16) store_kcontig_bf16(va[set], pa, a_hi);
      else if (p.ta) store_mncontig<A_ITEMS>(va[set], pa, a_hi, a_lo);
      else      store_kcontig<A_ITEMS>(va[set], pa, a_hi, a_lo);
      if (p.tb) store_kcontig<B_ITEMS>(vb[set], pb, b_hi, b_lo);
      else      store_mncontig<B_ITEMS>(vb[set], pb, b_hi, b_lo);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> UMMA reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    };
    if (num_kb > 0) issue(0, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      publish(0, kb);
      if (kb + 1 < num_kb) issue(0, kb + 1);             // in flight while the MMAs of slab kb run
    }
  } else if (lane == 0) {
    // ===== MMA issuer (one thread) =====
    // instruction descriptor, kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 and
    // 10-12), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                           (static_cast<uint32_t>(p.BN >> 3) << 17) |
                           (static_cast<uint32_t>(TM >> 4) << 24);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % p.stages;
      const uint32_t ph = (kb / p.stages) & 1;
      mbar_wait(&full_bar[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
      const uint32_t a_lo = a_hi + A_BYTES;
      const uint32_t b_hi = a_lo + A_BYTES;
      const uint32_t b_lo = b_hi + b_bytes;
      const int ksteps = min(KS / 8, (p.K - kb * KS + 7) / 8);
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t dah = make_desc(a_hi + k * 32), dal = make_desc(a_lo + k * 32);
        const uint64_t dbh = make_desc(b_hi + k * 32), dbl = make_desc(b_lo + k * 32);
        const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
        if (!p.a_bf16) umma_tf32(tmem_base, dal, dbh, idesc, acc0);
        umma_tf32(tmem_base, dah, dbl, idesc, p.a_bf16 ? acc0 : 1u);
        umma_tf32(tmem_base, dah, dbh, idesc, 1u);
      }
      umma_commit(&empty_bar[s]);                       // frees the stage when the MMAs retire
    }
    umma_commit(tmem_full_bar);
  }

  if (warp < PRODUCER_WARPS) {
    // ===== epilogue: TMEM -> registers -> global =====
    // All eight producer warps take part: warp w may read TMEM lanes 32 (w % 4) .. + 31, so warps w and
    // w + 4 share a row quarter and split its columns.  Two 16-column loads are in flight per wait.
    // (Four warps walking all BN columns one load at a time made the epilogue ~8 of the ~28 us a
    // 128 x 256 x 196 tile takes: the tile is short, nothing overlaps the epilogue.)
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float alpha = p.alpha_dev ? p.alpha * p.alpha_dev[0] : p.alpha;
    const int quarter = warp & 3, half = warp >> 2;
    const int row = m0 + quarter * 32 + lane;
    const int ncols = min(p.BN, p.N - n0);
    const int split = min(ncols, ((ncols + 31) / 32) * 16);          // multiple of 16
    const int c_begin = half ? split : 0, c_end = half ? ncols : split;
    auto store16 = [&](const uint32_t (&v)[16], int c0) {
      if (row >= p.M) return;
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(v[i]);
        if (p.col_sub && c0 + i < ncols) a -= p.col_sub[n0 + c0 + i];
        o[i] = alpha * a;
      }
      if (p.c_bf16) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + prob * p.sc +
                             static_cast<long>(row) * p.ldc + n0 + c0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + 4 * i < ncols) {
            const __nv_bfloat162 lo2 = __floats2bfloat162_rn(o[4 * i], o[4 * i + 1]);
            const __nv_bfloat162 hi2 = __floats2bfloat162_rn(o[4 * i + 2], o[4 * i + 3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo2);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi2);
            *reinterpret_cast<uint2*>(dst + 4 * i) = pk;
          }
        }
      } else {
        float* dst = C + static_cast<long>(row) * p.ldc + n0 + c0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + 4 * i < ncols)
            *reinterpret_cast<float4*>(dst + 4 * i) =
                make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
      }
    };
    auto load16 = [&](uint32_t (&v)[16], int c0) {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
            "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
    };
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      uint32_t v0[16], v1[16];
      const bool two = c0 + 16 < c_end;                  // warp-uniform
      load16(v0, c0);
      if (two) load16(v1, c0 + 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      store16(v0, c0);
      if (two) store16(v1, c0 + 16);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == PRODUCER_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================================
// TMA-fed, persistent, warp-specialised variant for fp32 operands (everything but the bf16-A launches).
//
// ncu on the register-staged kernel above (C2 step, `z = M_B B`: 6,144 tiles of 128 x 256 x 196): tensor
// pipe 21 % active, 53 % of the warp samples on the long scoreboard -- a slab's global loads are issued one
// slab ahead, their latency is longer than the MMAs of a slab, and split + store + MMA are serial on the
// single stage; cutting the producers' instruction count by 2.5x (staging plans, STS) changed nothing.
// Here the latency is taken off the warps altogether:
//   warp 12 : TMA.  cp.async.bulk.tensor.3d (batch = third coordinate) drops raw fp32 slabs into a ring of
//            RAW stages in the UMMA layout directly -- K-major boxes for operands whose contraction index
//            is contiguous in memory (SWIZZLE_128B), MN-major boxes (32 columns x 32 k rows, the 32-byte-atom
//            swizzle TF32 needs) for the others; the instruction descriptor's major bits tell the tensor core
//            which is which, so nothing is transposed by threads.  Out-of-range rows / k arrive as zeros.
//   warps 4..11 : split.  A slab is a linear array of 16-byte chunks whatever its layout: read x, write
//            hi = tf32_rn(x) back IN PLACE and lo = tf32_rn(x - hi) at the same offset of a LO stage
//            (the same descriptors serve both), fence.proxy.async, arrive.
//   warp 13 : one thread issues the three tcgen05.mma.kind::tf32 per K step (lo*hi + hi*lo + hi*hi) and
//            commits to the RAW and LO stage barriers.
//   warps 0..3 : epilogue of the PREVIOUS tile (tcgen05.ld -> alpha, column shift, cast -> global) from the
//            other of the two 256-column TMEM accumulators, overlapped with the current tile's main loop.
// One CTA per SM (ring: ~213 KB), tiles dealt round-robin.  Results are bitwise those of the kernel above.
constexpr int T_EPI_WARPS = 4, T_CONV_WARPS = 8;
constexpr int T_THREADS = (T_EPI_WARPS + T_CONV_WARPS + 2) * 32;
constexpr int T_TMEM_COLS = 512;

struct TParams {
  float* C;
  int M, N, K, ldc;
  long sc;
  int ta, tb;                 // 1: operand stored with the contraction index as the slow dimension
  int a_bf16;                 // A is bf16 (ta = 0): exact in TF32 -- expanded by the split warps, no lo term
  int a_batched, b_batched;   // 0: one operand shared by the whole batch
  int BN, bnp, tiles_m, tiles_n, batch;
  int raw_stages, lo_stages;
  float alpha; const float* alpha_dev;
  const float* col_sub;
  int c_bf16;
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// MN-major descriptor for 32-bit elements.  TF32 operands with the MN index contiguous have ONE legal
// shared-memory layout: 128-byte swizzle with 32-byte atoms (layout type 1; Swizzle<2,5,2>: the 32-byte chunk
// index of a 128-byte row is XORed with the row index mod 4), written by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  32-element MN chunks lie LBO = 4,096 bytes apart (one 32 x 32 box),
// groups of 4 k rows SBO = 512 bytes apart; a K step of 8 advances the start by 1,024 bytes.
// (The plain SWIZZLE_128B layout, which serves bf16 MN-major operands in gram_tc.cu, gives garbage here.)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((4096 >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((512 >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  d |= 1ull << 61;
  return d;
}

__global__ void __launch_bounds__(T_THREADS, 1)
gemm_tc3_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const TParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_bytes = p.bnp * 128;
  const int slab_bytes = A_BYTES + b_bytes;
  const uint32_t raw0 = smem_u32(smem);
  const uint32_t lo0 = raw0 + p.raw_stages * slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (p.raw_stages + p.lo_stages) * slab_bytes);
  uint64_t* raw_full = bars;            // [4]
  uint64_t* raw_empty = bars + 4;       // [4]
  uint64_t* lo_full = bars + 8;         // [4]
  uint64_t* lo_empty = bars + 12;       // [4]
  uint64_t* acc_full = bars + 16;       // [2]
  uint64_t* acc_empty = bars + 18;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + KS - 1) / KS;
  const int tpp = p.tiles_m * p.tiles_n;
  const long total = static_cast<long>(tpp) * p.batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], 1);
      mbar_init(&lo_full[s], T_CONV_WARPS);
      mbar_init(&lo_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], T_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == T_EPI_WARPS + T_CONV_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(T_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == T_EPI_WARPS + T_CONV_WARPS) {
    if (lane == 0) {
      // ===== TMA producer =====
      const uint32_t tx = (p.a_bf16 ? A_BYTES / 2 : A_BYTES) + (p.tb ? p.BN * 128 : b_bytes);
      long it = 0;
      for (long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int prob = static_cast<int>(tile / tpp), r = static_cast<int>(tile - static_cast<long>(prob) * tpp);
        const int m0 = (r / p.tiles_n) * TM, n0 = (r % p.tiles_n) * p.BN;
        const int pa = p.a_batched ? prob : 0, pb = p.b_batched ? prob : 0;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = static_cast<int>(it % p.raw_stages);
          const uint32_t ph = static_cast<uint32_t>(it / p.raw_stages) & 1u;
          mbar_wait(&raw_empty[s], ph ^ 1u);
          mbar_expect_tx(&raw_full[s], tx);
          const uint32_t a = raw0 + s * slab_bytes, b = a + A_BYTES;
          const int k0 = kb * KS;
          if (p.a_bf16) {                               // dense 128 x 32 bf16 rows in the upper half of the A slab
            tma_load_3d(a + A_BYTES / 2, &tmA, &raw_full[s], k0, m0, pa);
          } else if (p.ta) {
#pragma unroll
            for (int i = 0; i < TM / 32; ++i) tma_load_3d(a + i * 4096, &tmA, &raw_full[s], m0 + 32 * i, k0, pa);
          } else {
            tma_load_3d(a, &tmA, &raw_full[s], k0, m0, pa);
          }
          if (p.tb) {
            tma_load_3d(b, &tmB, &raw_full[s], k0, n0, pb);
          } else {
            for (int i = 0; i < p.bnp / 32; ++i) tma_load_3d(b + i * 4096, &tmB, &raw_full[s], n0 + 32 * i, k0, pb);
          }
        }
      }
    }
  } else if (warp == T_EPI_WARPS + T_CONV_WARPS + 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 and 10-12), bit 15 / 16: A / B is
      // MN-major, N >> 3 at bit 17, M >> 4 at bit 24.
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (p.ta ? (1u << 15) : 0u) |
                             (p.tb ? 0u : (1u << 16)) | (static_cast<uint32_t>(p.BN >> 3) << 17) |
                             (static_cast<uint32_t>(TM >> 4) << 24);
      const uint32_t a_step = p.ta ? 1024u : 32u, b_step = p.tb ? 32u : 1024u;
      long it = 0;
      int tcount = 0;
      for (long tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
        const int acc = tcount & 1;
        mbar_wait(&acc_empty[acc], (static_cast<uint32_t>(tcount >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = static_cast<int>(it % p.raw_stages), l = static_cast<int>(it % p.lo_stages);
          mbar_wait(&lo_full[l], static_cast<uint32_t>(it / p.lo_stages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = raw0 + s * slab_bytes, b_hi = a_hi + A_BYTES;
          const uint32_t a_lo = lo0 + l * slab_bytes, b_lo = a_lo + A_BYTES;
          const int ksteps = min(KS / 8, (p.K - kb * KS + 7) / 8);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t dah = p.ta ? make_desc_mn(a_hi + k * a_step) : make_desc(a_hi + k * a_step);
            const uint64_t dal = p.ta ? make_desc_mn(a_lo + k * a_step) : make_desc(a_lo + k * a_step);
            const uint64_t dbh = p.tb ? make_desc(b_hi + k * b_step) : make_desc_mn(b_hi + k * b_step);
            const uint64_t dbl = p.tb ? make_desc(b_lo + k * b_step) : make_desc_mn(b_lo + k * b_step);
            const uint32_t acc0 = (kb > 0 || k > 0) ? 1u : 0u;
            if (!p.a_bf16) umma_tf32(tmem_d, dal, dbh, idesc, acc0);
            umma_tf32(tmem_d, dah, dbl, idesc, p.a_bf16 ? acc0 : 1u);
            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
          }
          umma_commit(&raw_empty[s]);
          umma_commit(&lo_empty[l]);
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else if (warp >= T_EPI_WARPS) {
    // ===== split: hi in place, lo into the LO ring =====
    const int ctid = threadIdx.x - T_EPI_WARPS * 32;
    // fp32 A: the whole slab [A | B] is split; bf16 A: only B is, A is expanded (slabs are multiples of 4,096 bytes)
    const int first = p.a_bf16 ? (A_BYTES >> 4) / (T_CONV_WARPS * 32) : 0;
    const int per_thread = (slab_bytes >> 4) / (T_CONV_WARPS * 32);
    // bf16 A: chunk c of the dense 128 x 64-byte tile = 8 consecutive k of row c / 4
    const int brow = ctid >> 2, boc = ctid & 3;
    const uint32_t boff0 = sw_off(brow, boc * 8), boff1 = sw_off(brow, boc * 8 + 4);   // + 8,192 for row + 64
    long it = 0;
    for (long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = static_cast<int>(it % p.raw_stages), l = static_cast<int>(it % p.lo_stages);
        mbar_wait(&raw_full[s], static_cast<uint32_t>(it / p.raw_stages) & 1u);
        mbar_wait(&lo_empty[l], (static_cast<uint32_t>(it / p.lo_stages) & 1u) ^ 1u);
        if (p.a_bf16) {
          // bf16 -> fp32 is a 16-bit shift and exact in TF32.  The packed tile sits in the upper half of the
          // region its expansion fills, so every split thread reads its two chunks, all of them meet at a
          // named barrier, then they write.
          const uint32_t abase = raw0 + s * slab_bytes;
          uint32_t w[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 v = lds128(abase + A_BYTES / 2 + (ctid + u * T_CONV_WARPS * 32) * 16);
            w[u][0] = __float_as_uint(v.x); w[u][1] = __float_as_uint(v.y);
            w[u][2] = __float_as_uint(v.z); w[u][3] = __float_as_uint(v.w);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(T_CONV_WARPS * 32) : "memory");
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            sts128(abase + boff0 + u * 8192,
                   make_float4(__uint_as_float(w[u][0] << 16), __uint_as_float(w[u][0] & 0xffff0000u),
                               __uint_as_float(w[u][1] << 16), __uint_as_float(w[u][1] & 0xffff0000u)));
            sts128(abase + boff1 + u * 8192,
                   make_float4(__uint_as_float(w[u][2] << 16), __uint_as_float(w[u][2] & 0xffff0000u),
                               __uint_as_float(w[u][3] << 16), __uint_as_float(w[u][3] & 0xffff0000u)));
          }
        }
        const uint32_t src = raw0 + s * slab_bytes + ctid * 16, dst = lo0 + l * slab_bytes + ctid * 16;
        // four 16-byte chunks in flight per thread, two split warps per scheduler: the shared-memory load
        // latency was the whole cost of this pass with one warp per scheduler (ncu: every split warp always
        // busy, 40 % of its samples on the short scoreboard)
#pragma unroll 1
        for (int c0 = first; c0 < per_thread; c0 += 4) {
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c0 + j < per_thread) v[j] = lds128(src + static_cast<uint32_t>(c0 + j) * (T_CONV_WARPS * 32 * 16));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c0 + j < per_thread) {
              const uint32_t off = static_cast<uint32_t>(c0 + j) * (T_CONV_WARPS * 32 * 16);
              float4 h, lo4;
              split4(v[j], h, lo4);
              sts128(src + off, h);
              sts128(dst + off, lo4);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> UMMA reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&lo_full[l]);
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global, one tile behind the main loop =====
    const float alpha = p.alpha_dev ? p.alpha * p.alpha_dev[0] : p.alpha;
    int tcount = 0;
    for (long tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      const int prob = static_cast<int>(tile / tpp), r = static_cast<int>(tile - static_cast<long>(prob) * tpp);
      const int m0 = (r / p.tiles_n) * TM, n0 = (r % p.tiles_n) * p.BN;
      const int acc = tcount & 1;
      mbar_wait(&acc_full[acc], static_cast<uint32_t>(tcount >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + warp * 32 + lane;
      const int ncols = min(p.BN, p.N - n0);
      float* C = p.C + prob * p.sc;
      auto store16 = [&](const uint32_t (&v)[16], int c0) {
        if (row >= p.M) return;
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = __uint_as_float(v[i]);
          if (p.col_sub && c0 + i < ncols) a -= p.col_sub[n0 + c0 + i];
          o[i] = alpha * a;
        }
        if (p.c_bf16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + prob * p.sc + static_cast<long>(row) * p.ldc +
                               n0 + c0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (c0 + 4 * i < ncols) {
              const __nv_bfloat162 lo2 = __floats2bfloat162_rn(o[4 * i], o[4 * i + 1]);
              const __nv_bfloat162 hi2 = __floats2bfloat162_rn(o[4 * i + 2], o[4 * i + 3]);
              uint2 pk;
              pk.x = *reinterpret_cast<const uint32_t*>(&lo2);
              pk.y = *reinterpret_cast<const uint32_t*>(&hi2);
              *reinterpret_cast<uint2*>(dst + 4 * i) = pk;
            }
          }
        } else {
          float* dst = C + static_cast<long>(row) * p.ldc + n0 + c0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (c0 + 4 * i < ncols)
              *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
      };
      auto load16 = [&](uint32_t (&v)[16], int c0) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * 256 + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
      };
#pragma unroll 1
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v0[16], v1[16];
        const bool two = c0 + 16 < ncols;                  // warp-uniform
        load16(v0, c0);
        if (two) load16(v1, c0 + 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        store16(v0, c0);
        if (two) store16(v1, c0 + 16);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == T_EPI_WARPS + T_CONV_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T_TMEM_COLS)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// fp32 matrix (outer x inner, pitch ld) per problem -> 3-D map {inner, outer, batch}, box {32, box_outer, 1}
static CUresult encode_3d(EncodeTiledFn encode, CUtensorMap* map, const float* base, int inner, int outer, int ld,
                          long stride, int batch, int box_outer, bool mn_major) {
  const bool batched = stride != 0 && batch > 1;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer),
                              static_cast<cuuint64_t>(batched ? batch : 1)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 4,
                                 static_cast<cuuint64_t>(batched ? stride : static_cast<long>(ld) * outer) * 4};
  const cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_outer), 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estride,
                CU_TENSOR_MAP_INTERLEAVE_NONE,
                mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// bf16 matrix (outer x inner, pitch ld) per problem, dense (unswizzled) box {32, box_outer, 1}
static CUresult encode_3d_bf16(EncodeTiledFn encode, CUtensorMap* map, const void* base, int inner, int outer, int ld,
                               long stride, int batch, int box_outer) {
  const bool batched = stride != 0 && batch > 1;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer),
                              static_cast<cuuint64_t>(batched ? batch : 1)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2,
                                 static_cast<cuuint64_t>(batched ? stride : static_cast<long>(ld) * outer) * 2};
  const cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_outer), 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// Returns 0 on launch, -100 when this variant does not take the problem (the caller falls back).
static int launch_tma(int ta, int tb, int M, int N, int K, const void* A, int a_bf16, int lda, long sa, const float* B, int ldb,
                      long sb, void* C, int c_bf16, int ldc, long sc, int batch, float alpha, const float* alpha_dev,
                      const float* col_sub, cudaStream_t st) {
  EncodeTiledFn encode = encode_fn();
  if (!encode) return -100;
  TParams p;
  p.C = static_cast<float*>(C);
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.sc = sc;
  p.ta = ta ? 1 : 0; p.tb = tb ? 1 : 0;
  p.a_bf16 = a_bf16;
  p.a_batched = (sa != 0 && batch > 1) ? 1 : 0;
  p.b_batched = (sb != 0 && batch > 1) ? 1 : 0;
  // N tiles of at most 224 columns: five slabs (three raw, two lo) fit the SM, so the split of slab kb + 1
  // overlaps the MMAs of slab kb (256-column tiles leave room for one lo stage only: split and MMA serial,
  // 0.70 against 0.55 ms for 1,024 x (196 x 768 x 196))
  const int n_tiles = (N + 223) / 224;
  int bn = (N + n_tiles - 1) / n_tiles;
  bn = (bn + 15) & ~15;
  p.BN = bn;
  p.bnp = (bn + 31) & ~31;
  p.tiles_n = (N + bn - 1) / bn;
  p.tiles_m = (M + TM - 1) / TM;
  p.batch = batch;
  p.alpha = alpha; p.alpha_dev = alpha_dev; p.col_sub = col_sub; p.c_bf16 = c_bf16;
  const int slab = A_BYTES + p.bnp * 128;
  const int extra = 1024 + 256;
  if (5 * slab + extra <= SMEM_LIMIT) { p.raw_stages = 3; p.lo_stages = 2; }
  else if (4 * slab + extra <= SMEM_LIMIT) { p.raw_stages = 3; p.lo_stages = 1; }
  else return -100;
  const int dyn = (p.raw_stages + p.lo_stages) * slab + extra;
  CUtensorMap tmA, tmB;
  // A: ta = 0 stored M x K (K-major box {32 k, 128 rows}); ta = 1 stored K x M (MN-major boxes {32 rows, 32 k})
  const float* Af = static_cast<const float*>(A);
  if ((a_bf16 ? encode_3d_bf16(encode, &tmA, A, K, M, lda, sa, batch, TM)
              : ta ? encode_3d(encode, &tmA, Af, M, K, lda, sa, batch, 32, true)
                   : encode_3d(encode, &tmA, Af, K, M, lda, sa, batch, TM, false)) != CUDA_SUCCESS)
    return -100;
  // B: tb = 1 stored N x K (K-major box {32 k, BN rows}); tb = 0 stored K x N (MN-major boxes)
  if ((tb ? encode_3d(encode, &tmB, B, K, N, ldb, sb, batch, bn, false) : encode_3d(encode, &tmB, B, N, K, ldb, sb, batch, 32, true)) !=
      CUDA_SUCCESS)
    return -100;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long total = static_cast<long>(p.tiles_m) * p.tiles_n * batch;
  const int grid = static_cast<int>(total < sms ? total : sms);
  BASD_CUDA(cudaFuncSetAttribute(gemm_tc3_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  gemm_tc3_tma_kernel<<<grid, T_THREADS, dyn, st>>>(tmA, tmB, p);
  BASD_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc3
}  // namespace basd

// 1 when basd_gemm_tc3_batched accepts the problem (alignment / size rules below).
extern "C" int basd_gemm_tc3_supported(int M, int N, int K, int lda, int ldb, int ldc, long sa,
                                       long sb, long sc) {
  return M >= 1 && N >= 16 && K >= 8 && (N & 3) == 0 && (K & 3) == 0 && (M & 3) == 0 &&
         (lda & 3) == 0 && (ldb & 3) == 0 && (ldc & 3) == 0 && (sa & 3) == 0 && (sb & 3) == 0 &&
         (sc & 3) == 0;
}

// C[b] (M x N, pitch ldc) = alpha * alpha_dev[0] * (op(A[b]) op(B[b]) - 1 col_sub^T);  op as in
// basd_sgemm_batched: ta = 0: A stored M x K, ta = 1: stored K x M;  tb = 0: B stored K x N,
// tb = 1: stored N x K.  a_dtype = BASD_DTYPE_BF16 needs ta = 0 and K, lda, sa multiples of 8;
// c_dtype = BASD_DTYPE_BF16 stores bf16.  col_sub (N floats) may be null.
// All pointers 16-byte aligned, all pitches / strides multiples of 4 elements.
extern "C" int basd_gemm_tc3_batched_ex(int ta, int tb, int M, int N, int K, const void* A,
                                        int a_dtype, int lda, long sa, const float* B, int ldb,
                                        long sb, void* C, int c_dtype, int ldc, long sc, int batch,
                                        float alpha, const float* alpha_dev, const float* col_sub,
                                        void* stream) {
  using namespace basd;
  using namespace basd::tc3;
  if (batch <= 0 || M <= 0 || N <= 0) return 0;
  if (!basd_gemm_tc3_supported(M, N, K, lda, ldb, ldc, sa, sb, sc)) return -3;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) |
       reinterpret_cast<uintptr_t>(C)) & 15)
    return -3;
  const int a_bf16 = a_dtype == BASD_DTYPE_BF16;
  if (a_bf16 && (ta || (K & 7) || (lda & 7) || (sa & 7))) return -3;
  if (batch > 65535) return -4;
  if (!std::getenv("BASD_TC3_NO_TMA")) {
    const int e = tc3::launch_tma(ta, tb, M, N, K, A, a_bf16, lda, sa, B, ldb, sb, C,
                                  c_dtype == BASD_DTYPE_BF16, ldc, sc, batch, alpha, alpha_dev, col_sub,
                                  (cudaStream_t)stream);
    if (e != -100) return e;
  }
  Params p;
  p.A = static_cast<const float*>(A); p.B = B; p.C = static_cast<float*>(C);
  p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sa = sa; p.sb = sb; p.sc = sc;
  p.ta = ta ? 1 : 0;
  p.tb = tb ? 1 : 0;
  p.a_bf16 = a_bf16;
  p.col_sub = col_sub;
  p.c_bf16 = c_dtype == BASD_DTYPE_BF16;
  const int n_tiles = (N + 255) / 256;
  int bn = (N + n_tiles - 1) / n_tiles;
  bn = (bn + 15) & ~15;
  p.BN = bn;
  p.tiles_n = (N + bn - 1) / bn;
  const int stage_bytes = 2 * A_BYTES + 2 * bn * 128;
  p.stages = 1;
  p.alpha = alpha;
  p.alpha_dev = alpha_dev;
  const int dyn = stage_bytes + 1024 + 256;
  if (2 * dyn > SMEM_LIMIT) return -5;
  BASD_CUDA(cudaFuncSetAttribute(gemm_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  dim3 grid(((M + TM - 1) / TM) * p.tiles_n, batch);
  gemm_tc3_kernel<<<grid, THREADS, dyn, (cudaStream_t)stream>>>(p);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_gemm_tc3_batched(int ta, int tb, int M, int N, int K, const float* A, int lda,
                                     long sa, const float* B, int ldb, long sb, float* C, int ldc,
                                     long sc, int batch, float alpha, const float* alpha_dev,
                                     void* stream) {
  return basd_gemm_tc3_batched_ex(ta, tb, M, N, K, A, BASD_DTYPE_F32, lda, sa, B, ldb, sb, C,
                                  BASD_DTYPE_F32, ldc, sc, batch, alpha, alpha_dev, nullptr, stream);
}

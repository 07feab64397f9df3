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

// byte offset of element (row, k) inside a K-major SWIZZLE_128B slab (k in [0,32))
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
  return static_cast<uint32_t>(((row >> 3) << 10) + ((row & 7) << 7) +
                               ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}

// The global loads of slab kb+1 are issued right after slab kb has been split and written to shared memory,
// so they are in flight while the tensor core works on slab kb; the second CTA of the SM covers the rest.
//
// Operand whose contraction index is contiguous in memory (row-major R x K with pitch ld):
// one float4 = 4 consecutive k of one row; a quarter warp covers one 128-byte row -> conflict-free
// 128-bit stores.  Item f of a thread: f = ptid + u * PRODUCERS, row = f / 8, chunk = f % 8.
template <int CNT>
__device__ __forceinline__ void issue_kcontig(float4 (&v)[CNT], const float* __restrict__ g, int ld,
                                              int r0, int R, int row_limit, int k0, int K, int ptid) {
  const int items = R * 8;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const int f = ptid + u * PRODUCERS;
    const int row = f >> 3, ch = f & 7;
    const int gr = r0 + row, gk = k0 + ch * 4;
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < items && gr < row_limit && gk < K)
      v[u] = __ldg(reinterpret_cast<const float4*>(g + static_cast<long>(gr) * ld + gk));
  }
}
template <int CNT>
__device__ __forceinline__ void store_kcontig(const float4 (&v)[CNT], int R, uint8_t* hi, uint8_t* lo,
                                              int ptid) {
  const int items = R * 8;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const int f = ptid + u * PRODUCERS;
    if (f < items) {
      const int row = f >> 3, ch = f & 7;
      float4 h, l;
      split4(v[u], h, l);
      const uint32_t off = sw_off(row, ch * 4);
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  }
}

// Operand whose contraction index is the slow dimension in memory (row-major K x R with pitch
// ld): one float4 = 4 consecutive rows of the slab at one k; transposed by scalar stores.  A warp
// covers 16 k x 2 row-quads: the 32 scalar stores of each of the 4 components hit 32 banks.
// Block blk of a warp: blk = warp + u * PRODUCER_WARPS over (row octets) x (two halves of the slab).
template <int CNT>
__device__ __forceinline__ void issue_mncontig(float4 (&v)[CNT], const float* __restrict__ g, int ld,
                                               int r0, int R, int row_limit, int k0, int K, int ptid) {
  const int lane = ptid & 31, warp = ptid >> 5;
  const int kk = lane & 15, ql = lane >> 4;
  const int blocks = (R >> 3) * 2;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const int blk = warp + u * PRODUCER_WARPS;
    const int k = (blk & 1) * 16 + kk, row = (blk >> 1) * 8 + ql * 4;
    const int gk = k0 + k, gr = r0 + row;
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blk < blocks && gk < K && gr < row_limit)
      v[u] = __ldg(reinterpret_cast<const float4*>(g + static_cast<long>(gk) * ld + gr));
  }
}
template <int CNT>
__device__ __forceinline__ void store_mncontig(const float4 (&v)[CNT], int R, uint8_t* hi, uint8_t* lo,
                                               int ptid) {
  const int lane = ptid & 31, warp = ptid >> 5;
  const int kk = lane & 15, ql = lane >> 4;
  const int blocks = (R >> 3) * 2;
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const int blk = warp + u * PRODUCER_WARPS;
    if (blk < blocks) {
      const int k = (blk & 1) * 16 + kk, row = (blk >> 1) * 8 + ql * 4;
      float4 h, l;
      split4(v[u], h, l);
      const float hv[4] = {h.x, h.y, h.z, h.w}, lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = sw_off(row + i, k);
        *reinterpret_cast<float*>(hi + off) = hv[i];
        *reinterpret_cast<float*>(lo + off) = lv[i];
      }
    }
  }
}

// bf16 A operand (row-major R x K, K contiguous): one 128-bit load = 8 consecutive k of one row;
// bf16 is a subset of TF32, so hi is exact, lo is zero and its MMA is skipped.
// Item f = ptid + u * PRODUCERS (u < 2): row = f / 4, octet = f % 4.
__device__ __forceinline__ void issue_kcontig_bf16(float4 (&v)[4], const __nv_bfloat16* __restrict__ g,
                                                   int ld, int r0, int row_limit, int k0, int K,
                                                   int ptid) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int f = ptid + u * PRODUCERS;
    const int row = f >> 2, oc = f & 3;
    const int gr = r0 + row, gk = k0 + oc * 8;
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (gr < row_limit && gk < K)
      raw = __ldg(reinterpret_cast<const uint4*>(g + static_cast<long>(gr) * ld + gk));
    v[u] = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z),
                       __uint_as_float(raw.w));
  }
}
__device__ __forceinline__ void store_kcontig_bf16(const float4 (&v)[4], uint8_t* hi, int ptid) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int f = ptid + u * PRODUCERS;
    const int row = f >> 2, oc = f & 3;
    const uint32_t w[4] = {__float_as_uint(v[u].x), __float_as_uint(v[u].y), __float_as_uint(v[u].z),
                           __float_as_uint(v[u].w)};
    const float4 lo4 = make_float4(__uint_as_float(w[0] << 16), __uint_as_float(w[0] & 0xffff0000u),
                                   __uint_as_float(w[1] << 16), __uint_as_float(w[1] & 0xffff0000u));
    const float4 hi4 = make_float4(__uint_as_float(w[2] << 16), __uint_as_float(w[2] & 0xffff0000u),
                                   __uint_as_float(w[3] << 16), __uint_as_float(w[3] & 0xffff0000u));
    *reinterpret_cast<float4*>(hi + sw_off(row, oc * 8)) = lo4;        // k .. k+3
    *reinterpret_cast<float4*>(hi + sw_off(row, oc * 8 + 4)) = hi4;    // k+4 .. k+7
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
    auto issue = [&](int set, int kb) {
      const int k0 = kb * KS;
      if (p.a_bf16) issue_kcontig_bf16(va[set], reinterpret_cast<const __nv_bfloat16*>(p.A) + prob * p.sa,
                                       p.lda, m0, p.M, k0, p.K, ptid);
      else if (p.ta) issue_mncontig<A_ITEMS>(va[set], A, p.lda, m0, TM, p.M, k0, p.K, ptid);
      else      issue_kcontig<A_ITEMS>(va[set], A, p.lda, m0, TM, p.M, k0, p.K, ptid);
      if (p.tb) issue_kcontig<B_ITEMS>(vb[set], B, p.ldb, n0, p.BN, p.N, k0, p.K, ptid);
      else      issue_mncontig<B_ITEMS>(vb[set], B, p.ldb, n0, p.BN, p.N, k0, p.K, ptid);
    };
    auto publish = [&](int set, int kb) {
      const int s = kb % p.stages;
      const uint32_t ph = (kb / p.stages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* a_hi = smem + s * stage_bytes;
      uint8_t* a_lo = a_hi + A_BYTES;
      uint8_t* b_hi = a_lo + A_BYTES;
      uint8_t* b_lo = b_hi + b_bytes;
      if (p.a_bf16) store_kcontig_bf16(va[set], a_hi, ptid);
      else if (p.ta) store_mncontig<A_ITEMS>(va[set], TM, a_hi, a_lo, ptid);
      else      store_kcontig<A_ITEMS>(va[set], TM, a_hi, a_lo, ptid);
      if (p.tb) store_kcontig<B_ITEMS>(vb[set], p.BN, b_hi, b_lo, ptid);
      else      store_mncontig<B_ITEMS>(vb[set], p.BN, b_hi, b_lo, ptid);
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

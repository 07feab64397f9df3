// Batched FP32-accurate GEMM on the 5th-generation tensor cores:  C = alpha * op(A) op(B)
// with every fp32 operand split into two TF32 terms (x = hi + lo, both rounded to nearest) and
// three tcgen05.mma.kind::tf32 products per K step (lo*hi + hi*lo + hi*hi, fp32 accumulation in
// TMEM) -- the "3xTF32" scheme: the dropped lo*lo term and the rounding of lo are both
// ~2^-22 |a||b|, i.e. fp32-level products, so the per-sample Grams and factor products that feed
// rank decisions keep their noise floor while running on the tensor pipe instead of FFMA.
//
// The matrices of this path are small (N_tokens x N_tokens x {N_tokens, D}), batched by sample: short tiles
// (K = 196: seven slabs of 32) whose operands stream from L2 / HBM.  A persistent, warp-specialised CTA per
// SM works through the 128 x BN tiles (BN <= 224) of all problems:
//   warp 12 : TMA.  cp.async.bulk.tensor.3d (batch = third coordinate) drops raw fp32 slabs into a ring of
//            RAW stages in the UMMA layout directly -- K-major boxes for operands whose contraction index
//            is contiguous in memory (SWIZZLE_128B), MN-major boxes (32 columns x 32 k rows, the 32-byte-atom
//            swizzle TF32 needs) for the others; the instruction descriptor's major bits tell the tensor core
//            which is which, so nothing is transposed by threads.  Out-of-range rows / k arrive as zeros.
//   warps 4..11 : split.  A slab is a linear array of 16-byte chunks whatever its layout: read x, write
//            hi = tf32_rn(x) back IN PLACE and lo = tf32_rn(x - hi) at the same offset of a LO stage
//            (the same descriptors serve both), fence.proxy.async, arrive.  A bf16 A operand arrives as a
//            packed tile and is expanded here (exact in TF32: no lo term, two MMAs per K step).
//   warp 13 : one thread issues the tcgen05.mma.kind::tf32 (M = 128, N = BN, K = 8) and commits to the RAW
//            and LO stage barriers.
//   warps 0..3 : epilogue of the PREVIOUS tile (tcgen05.ld -> alpha, column shift, cast -> global) from the
//            other of the two 256-column TMEM accumulators, overlapped with the current tile's main loop.
// Ring: three raw + two lo slabs (<= 220 KB), tiles dealt round-robin.
//
// History: round 1 staged the operands through registers (coalesced loads -> split ->
// swizzled / transposing stores), first with two or three stages and one CTA per SM, then with one stage and
// two CTAs per SM (5.97 -> 5.01 ms for the 30 launches of a C2 step).  ncu on that kernel: tensor pipe 21 %
// active, 53 % of the warp samples on the long scoreboard -- a slab's global loads were issued one slab
// ahead, their latency is longer than the MMAs of a slab, and split + store + MMA were serial on the single
// stage; cutting the producers' instruction count by 2.5x changed nothing.  This kernel is bitwise equal to
// it and takes 3.9 ms for the same launches; what bounds it now is shared-memory bandwidth (TMA writes +
// split reads / writes + three operand reads per K step: ~310 KB per 45 KB slab).
// Replaces the SIMT batched SGEMM (gemm_simt.cu) for the Procrustes products
// (reference: torch.bmm at relational.py:47 and the matmuls inside linalg.svd's backward).
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace basd {
namespace tc3 {

constexpr int TM = 128;                   // UMMA M
constexpr int KS = 32;                    // fp32 per K slab = one 128-byte swizzle row
constexpr int A_BYTES = TM * 128;         // one split term of the A slab
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

// Shared-window (32-bit) addresses throughout: the ring base is carved out of the dynamic shared memory by
// integer alignment, which makes pointers derived from it GENERIC for the compiler (ST.E / LD.E with 64-bit
// address arithmetic instead of STS / LDS with an immediate offset).
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B slab (k in [0,32))
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
  return static_cast<uint32_t>(((row >> 3) << 10) + ((row & 7) << 7) +
                               ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}

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
  const int e = tc3::launch_tma(ta, tb, M, N, K, A, a_bf16, lda, sa, B, ldb, sb, C, c_dtype == BASD_DTYPE_BF16, ldc,
                                sc, batch, alpha, alpha_dev, col_sub, (cudaStream_t)stream);
  if (e != -100) return e;
  return -5;
}

extern "C" int basd_gemm_tc3_batched(int ta, int tb, int M, int N, int K, const float* A, int lda,
                                     long sa, const float* B, int ldb, long sb, float* C, int ldc,
                                     long sc, int batch, float alpha, const float* alpha_dev,
                                     void* stream) {
  return basd_gemm_tc3_batched_ex(ta, tb, M, N, K, A, BASD_DTYPE_F32, lda, sa, B, ldb, sb, C,
                                  BASD_DTYPE_F32, ldc, sc, batch, alpha, alpha_dev, nullptr, stream);
}

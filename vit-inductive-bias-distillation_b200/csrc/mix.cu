// HBM-bound kernels of the layer mixing (reference: layer_selector.py:110-112,
// combined.py:9-14, relational.py:22-34 and the autograd of those lines).
//
//  basd_attn_rows     : attention map -> per-token importance row (CLS row / query mean),
//                       reduced BEFORE mixing (mixing is linear) so the full
//                       (L,B,H,N+1,N+1) stack is never materialised.
//  basd_mix_interp    : all E mixed+aligned teacher tensors in ONE pass over the teacher
//                       stack, 128-bit loads, 1-D linear resampling fused in.
//  basd_mix_rows      : mixed, resampled, normalised importance weights.
//  basd_weight_grad   : dL/d(mixing weights): inner products of the upstream token gradient
//                       with every teacher layer, one pass over the stack.
#include "common.cuh"
#include <cstdlib>

namespace basd {

constexpr int MAX_E = 8;
constexpr int MAX_L = 64;

struct LayerPtrs { const void* p[MAX_L]; };

// 1-D linear taps, align_corners=False (ATen upsample_linear1d):
//   src = max((i + .5) * n_src/n_dst - .5, 0); lo = floor(src); hi = min(lo+1, n_src-1)
__device__ __forceinline__ void taps(int i, int n_src, int n_dst, int& lo, int& hi, float& f) {
  if (n_src == n_dst) { lo = hi = i; f = 0.f; return; }
  float src = (i + 0.5f) * ((float)n_src / (float)n_dst) - 0.5f;
  src = fmaxf(src, 0.f);
  lo = min((int)src, n_src - 1);
  hi = min(lo + 1, n_src - 1);
  f = src - (float)lo;
}

template <typename T>
__global__ void attn_rows_kernel(const T* __restrict__ attn, int H, int side, int q_rows,
                                 int has_cls, int n_tok, float* __restrict__ rows) {
  const int b = blockIdx.x;
  const T* base = attn + (long)b * H * q_rows * side;
  for (int n = threadIdx.x; n < n_tok; n += blockDim.x) {
    float s = 0.f;
    if (has_cls) {    // CLS query row (row 0), keys 1..N
      for (int h = 0; h < H; ++h) s += to_f32<T>(base[(long)h * q_rows * side + 1 + n]);
      s /= (float)H;
    } else {
      for (int h = 0; h < H; ++h)
        for (int q = 0; q < q_rows; ++q) s += to_f32<T>(base[((long)h * q_rows + q) * side + n]);
      s /= (float)(H * q_rows);
    }
    rows[(long)b * n_tok + n] = s;
  }
}

// One thread per (b, n_dst, 8-wide column group). VEC = 8 elements.
template <typename TIn, typename TOut, int NE, int UNR>
__global__ void __launch_bounds__(256, (NE <= 4 ? 2 : 1))
mix_interp_kernel(LayerPtrs layers, int L, int E, const float* __restrict__ weights, int B,
                  int n_src, int n_dst, int D, TOut* __restrict__ out) {
  __shared__ float w[MAX_E * MAX_L];
  for (int i = threadIdx.x; i < E * L; i += blockDim.x) w[i] = weights[i];
  __syncthreads();
  const int groups = D >> 3;
  const long total = (long)B * n_dst * groups;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % groups;
  const long tok = idx / groups;
  const int n = tok % n_dst;
  const int b = tok / n_dst;
  int lo, hi;
  float f;
  taps(n, n_src, n_dst, lo, hi, f);
  float acc[NE][8];
#pragma unroll
  for (int i = 0; i < NE; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  const long off_lo = ((long)b * n_src + lo) * D + g * 8;
  const long off_hi = ((long)b * n_src + hi) * D + g * 8;
  // layers are walked UNR (4 or 6) at a time with every 128-bit load of the group issued before
  // any of them is consumed.  Two 256-thread CTAs fit per SM at <= 128 registers, so the bytes in
  // flight per SM are 16 warps x 32 lanes x UNR x 16 B: 32 KB at UNR = 4, 48 KB at UNR = 6 --
  // Little's law at 6.5 TB/s and ~800 ns asks for ~36 KB per SM.
  for (int l0 = 0; l0 < L; l0 += UNR) {
    float v[UNR][8];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int l = min(l0 + u, L - 1);
      const TIn* src = reinterpret_cast<const TIn*>(layers.p[l]);
      if (sizeof(TIn) == 2) {
        load8(reinterpret_cast<const __nv_bfloat16*>(src) + off_lo, v[u]);
      } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off_lo);
        const float4 a = __ldg(p), bq = __ldg(p + 1);
        v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
        v[u][4] = bq.x; v[u][5] = bq.y; v[u][6] = bq.z; v[u][7] = bq.w;
      }
    }
    if (f != 0.f) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int l = min(l0 + u, L - 1);
        const TIn* src = reinterpret_cast<const TIn*>(layers.p[l]);
        float hv[8];
        if (sizeof(TIn) == 2) {
          load8(reinterpret_cast<const __nv_bfloat16*>(src) + off_hi, hv);
        } else {
          const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off_hi);
          const float4 c0 = __ldg(q), c1 = __ldg(q + 1);
          hv[0] = c0.x; hv[1] = c0.y; hv[2] = c0.z; hv[3] = c0.w;
          hv[4] = c1.x; hv[5] = c1.y; hv[6] = c1.z; hv[7] = c1.w;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) v[u][c] = fmaf(f, hv[c] - v[u][c], v[u][c]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (l0 + u < L) {
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          if (i < E) {
            const float wi = w[i * L + l0 + u];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(wi, v[u][c], acc[i][c]);
          }
        }
      }
    }
  }
  const long slab = (long)B * n_dst * D;
  const long o = ((long)b * n_dst + n) * D + g * 8;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    if (i < E) {
      TOut* dst = out + (long)i * slab + o;
      if (sizeof(TOut) == 2) {
        uint4 pk;
        uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __nv_bfloat162 h = __floats2bfloat162_rn(acc[i][2 * c], acc[i][2 * c + 1]);
          pw[c] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst) = pk;
      } else {
        float4* d4 = reinterpret_cast<float4*>(dst);
        d4[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        d4[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
  }
}

// rows: (L, B, n_src) fp32. out w: (E, B, n_dst) normalised, totals: (E, B).
__global__ void mix_rows_kernel(const float* __restrict__ rows, const float* __restrict__ weights,
                                int L, int B, int n_src, int n_dst, float* __restrict__ w_out,
                                float* __restrict__ totals) {
  __shared__ float red[32];
  const int b = blockIdx.x, i = blockIdx.y;
  float local = 0.f;
  for (int n = threadIdx.x; n < n_dst; n += blockDim.x) {
    int lo, hi;
    float f;
    taps(n, n_src, n_dst, lo, hi, f);
    float a = 0.f, c = 0.f;
    for (int l = 0; l < L; ++l) {
      const float* r = rows + ((long)l * B + b) * n_src;
      const float wl = weights[i * L + l];
      a = fmaf(wl, r[lo], a);
      c = fmaf(wl, r[hi], c);
    }
    const float v = a + f * (c - a);
    w_out[((long)i * B + b) * n_dst + n] = v;
    local += v;
  }
  const float tot = block_sum(local, red);
  __syncthreads();
  for (int n = threadIdx.x; n < n_dst; n += blockDim.x)
    w_out[((long)i * B + b) * n_dst + n] /= tot;
  if (threadIdx.x == 0) totals[(long)i * B + b] = tot;
}

// partial[(slice*L + l)*E + i] = sum over the slice of <Z_i, resample(T_l)>.
// Z: (E, B, n_dst, D) fp32. Grid: (slices, L).
template <typename TIn>
__global__ void __launch_bounds__(256)
weight_grad_kernel(LayerPtrs layers, int E, const float* __restrict__ Z, int B, int n_src,
                   int n_dst, int D, float* __restrict__ partial) {
  __shared__ float red[32];
  const int l = blockIdx.y;
  const TIn* src = reinterpret_cast<const TIn*>(layers.p[l]);
  const int groups = D >> 3;
  const long total = (long)B * n_dst * groups;
  const long slab = (long)B * n_dst * D;
  float acc[MAX_E];
#pragma unroll
  for (int i = 0; i < MAX_E; ++i) acc[i] = 0.f;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long)gridDim.x * blockDim.x) {
    const int g = idx % groups;
    const long tok = idx / groups;
    const int n = tok % n_dst;
    const int b = tok / n_dst;
    int lo, hi;
    float f;
    taps(n, n_src, n_dst, lo, hi, f);
    float v[8];
    const long off_lo = ((long)b * n_src + lo) * D + g * 8;
    const long off_hi = ((long)b * n_src + hi) * D + g * 8;
    if (sizeof(TIn) == 2) {
      load8(reinterpret_cast<const __nv_bfloat16*>(src) + off_lo, v);
      if (f != 0.f) {
        float u[8];
        load8(reinterpret_cast<const __nv_bfloat16*>(src) + off_hi, u);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaf(f, u[c] - v[c], v[c]);
      }
    } else {
      const float* p = reinterpret_cast<const float*>(src);
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = p[off_lo + c];
      if (f != 0.f) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaf(f, p[off_hi + c] - v[c], v[c]);
      }
    }
    const long o = ((long)b * n_dst + n) * D + g * 8;
#pragma unroll
    for (int i = 0; i < MAX_E; ++i) {
      if (i < E) {
        const float4* z = reinterpret_cast<const float4*>(Z + (long)i * slab + o);
        const float4 z0 = __ldg(z), z1 = __ldg(z + 1);
        float s = acc[i];
        s = fmaf(z0.x, v[0], s); s = fmaf(z0.y, v[1], s); s = fmaf(z0.z, v[2], s); s = fmaf(z0.w, v[3], s);
        s = fmaf(z1.x, v[4], s); s = fmaf(z1.y, v[5], s); s = fmaf(z1.z, v[6], s); s = fmaf(z1.w, v[7], s);
        acc[i] = s;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAX_E; ++i) {
    if (i < E) {
      const float s = block_sum(acc[i], red);
      if (threadIdx.x == 0) partial[((long)blockIdx.x * gridDim.y + l) * E + i] = s;
    }
  }
}

// The same partial sums with the upstream gradient read ONCE (default for E <= 4; measured on B200 at C2:
// 0.96 -> 0.59 ms).  weight_grad_kernel above runs one block column per teacher layer, so every block
// column re-reads its slice of Z (E*B*N*D fp32, 616 MB at C2) -- ncu shows 2.4x the algorithmic DRAM
// bytes.  Here a thread keeps its E x 8 values of Z in registers and walks LC teacher layers with them,
// accumulating E x LC partial sums (48 registers for E = 4, LC = 12); grid = (slices, ceil(L / LC)).
template <typename TIn, int EC, int LC, bool FAST>
__global__ void __launch_bounds__(256, 2)
weight_grad_onepass_kernel(LayerPtrs layers, int L, int E, const float* __restrict__ Z, int B, int n_src,
                           int n_dst, int D, float* __restrict__ partial) {
  __shared__ float red[8][EC * LC];
  const int l0 = blockIdx.y * LC;
  const int groups = D >> 3;
  const long total = (long)B * n_dst * groups;
  const long slab = (long)B * n_dst * D;
  float acc[EC][LC];
#pragma unroll
  for (int i = 0; i < EC; ++i)
#pragma unroll
    for (int k = 0; k < LC; ++k) acc[i][k] = 0.f;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long)gridDim.x * blockDim.x) {
    const int g = idx % groups;
    const long tok = idx / groups;
    const int n = tok % n_dst;
    const int b = tok / n_dst;
    int lo, hi;
    float f;
    taps(n, n_src, n_dst, lo, hi, f);
    const long off_lo = ((long)b * n_src + lo) * D + g * 8;
    const long off_hi = ((long)b * n_src + hi) * D + g * 8;
    const long o = ((long)b * n_dst + n) * D + g * 8;
    float z[EC][8];
#pragma unroll
    for (int i = 0; i < EC; ++i) {
      if (i < E) {
        const float4* zp = reinterpret_cast<const float4*>(Z + (long)i * slab + o);
        const float4 z0 = __ldg(zp), z1 = __ldg(zp + 1);
        z[i][0] = z0.x; z[i][1] = z0.y; z[i][2] = z0.z; z[i][3] = z0.w;
        z[i][4] = z1.x; z[i][5] = z1.y; z[i][6] = z1.z; z[i][7] = z1.w;
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) z[i][c] = 0.f;
      }
    }
    if (FAST && sizeof(TIn) == 2) {
      // bf16 stack, no resampling: the 128-bit loads of UNR layers are issued before any is consumed and
      // stay packed (4 registers each) until their turn
      constexpr int UNR = (LC % 6 == 0) ? 6 : 4;
#pragma unroll
      for (int k0 = 0; k0 < LC; k0 += UNR) {
        uint4 raw[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          raw[u] = __ldg(reinterpret_cast<const uint4*>(
              reinterpret_cast<const __nv_bfloat16*>(layers.p[min(l0 + k0 + u, L - 1)]) + off_lo));
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (l0 + k0 + u < L) {
            const uint32_t w4[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
            for (int i = 0; i < EC; ++i) {
              float sum = acc[i][k0 + u];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                sum = fmaf(z[i][2 * c], __uint_as_float(w4[c] << 16), sum);
                sum = fmaf(z[i][2 * c + 1], __uint_as_float(w4[c] & 0xffff0000u), sum);
              }
              acc[i][k0 + u] = sum;
            }
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int k = 0; k < LC; ++k) {
      if (l0 + k < L) {
        const TIn* src = reinterpret_cast<const TIn*>(layers.p[l0 + k]);
        float v[8];
        if (sizeof(TIn) == 2) {
          load8(reinterpret_cast<const __nv_bfloat16*>(src) + off_lo, v);
          if (f != 0.f) {
            float u[8];
            load8(reinterpret_cast<const __nv_bfloat16*>(src) + off_hi, u);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = fmaf(f, u[c] - v[c], v[c]);
          }
        } else {
          const float* p = reinterpret_cast<const float*>(src);
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = p[off_lo + c];
          if (f != 0.f) {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = fmaf(f, p[off_hi + c] - v[c], v[c]);
          }
        }
#pragma unroll
        for (int i = 0; i < EC; ++i) {
          float sum = acc[i][k];
#pragma unroll
          for (int c = 0; c < 8; ++c) sum = fmaf(z[i][c], v[c], sum);
          acc[i][k] = sum;
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < EC; ++i)
#pragma unroll
    for (int k = 0; k < LC; ++k) {
      float v = acc[i][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][i * LC + k] = v;
    }
  __syncthreads();
  if (threadIdx.x < EC * LC) {
    const int i = threadIdx.x / LC, k = threadIdx.x % LC;
    if (i < E && l0 + k < L) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      partial[((long)blockIdx.x * L + (l0 + k)) * E + i] = v;
    }
  }
}

// d_weights[i,l] = sum_slices partial + sum_{b,n} gw[i,b,n] * resample(rows[l,b,:])[n]
// One CTA of 1,024 threads per (l, i): thread (tx, ty) owns token n = tx (+ 256, ...) -- its two resampling
// taps are computed once -- and walks the samples b = ty, ty + 4, ...: coalesced along n, no per-element
// division (the first version's flat index cost a div / mod and a tap evaluation per element: 0.11 ms for
// 50,176 products per CTA).
__global__ void __launch_bounds__(1024)
weight_grad_finish_kernel(const float* __restrict__ partial, int slices,
                          const float* __restrict__ gw,
                          const float* __restrict__ rows, int E, int L, int B,
                          int n_rows, int n_dst, float gw_scale,
                          const float* __restrict__ gw_scale_dev,
                          float* __restrict__ d_weights) {
  __shared__ float red[32];
  const int l = blockIdx.x, i = blockIdx.y;
  const int tx = threadIdx.x & 255, ty = threadIdx.x >> 8;
  float s = 0.f, t = 0.f;
  for (int sl = threadIdx.x; sl < slices; sl += blockDim.x) t += partial[((long)sl * L + l) * E + i];
  for (int n = tx; n < n_dst; n += 256) {
    int lo, hi;
    float f;
    taps(n, n_rows, n_dst, lo, hi, f);
    for (int b = ty; b < B; b += 4) {
      const float* r = rows + ((long)l * B + b) * n_rows;
      const float v = r[lo] + f * (r[hi] - r[lo]);
      s = fmaf(gw[((long)i * B + b) * n_dst + n], v, s);
    }
  }
  if (gw_scale_dev) gw_scale *= *gw_scale_dev;
  s = block_sum(fmaf(gw_scale, s, t), red);
  if (threadIdx.x == 0) d_weights[i * L + l] = s;
}

// Flat layer mix for GrassmannianLayerSelector.forward's reference-shaped outputs (the mixed full
// attention maps, layer_selector.py:112): out[i][x] = sum_l w[i,l] layer_l[x] over `numel` elements of any
// shape, all E outputs in ONE pass over the L inputs, 128-bit loads on the aligned body, scalar tail.
template <typename T, int NE>
__global__ void __launch_bounds__(256)
mix_flat_kernel(LayerPtrs layers, int L, int E, const float* __restrict__ weights, long numel,
                T* __restrict__ out) {
  __shared__ float w[MAX_E * MAX_L];
  for (int i = threadIdx.x; i < E * L; i += blockDim.x) w[i] = weights[i];
  __syncthreads();
  constexpr int V = 16 / sizeof(T);                      // elements per 128-bit access
  const long groups = numel / V;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long)gridDim.x * blockDim.x) {
    float acc[NE][V];
#pragma unroll
    for (int i = 0; i < NE; ++i)
#pragma unroll
      for (int c = 0; c < V; ++c) acc[i][c] = 0.f;
    for (int l = 0; l < L; ++l) {
      float v[V];
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(layers.p[l]) + g * V));
      if (sizeof(T) == 2) {
        const uint32_t q[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          v[(2 * c) % V] = __uint_as_float(q[c] << 16);
          v[(2 * c + 1) % V] = __uint_as_float(q[c] & 0xffff0000u);
        }
      } else {
        v[0] = __uint_as_float(raw.x); v[1 % V] = __uint_as_float(raw.y);
        v[2 % V] = __uint_as_float(raw.z); v[3 % V] = __uint_as_float(raw.w);
      }
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        if (i < E) {
          const float wi = w[i * L + l];
#pragma unroll
          for (int c = 0; c < V; ++c) acc[i][c] = fmaf(wi, v[c], acc[i][c]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      if (i < E) {
        T* dst = out + (long)i * numel + g * V;
#pragma unroll
        for (int c = 0; c < V; ++c) dst[c] = from_f32<T>(acc[i][c]);
      }
    }
  }
  // tail (numel % V elements) and nothing else: one thread each in the first block
  const long tail0 = groups * V;
  if (blockIdx.x == 0 && tail0 + threadIdx.x < numel) {
    const long x = tail0 + threadIdx.x;
    for (int i = 0; i < E; ++i) {
      float a = 0.f;
      for (int l = 0; l < L; ++l) a = fmaf(w[i * L + l], to_f32<T>(reinterpret_cast<const T*>(layers.p[l])[x]), a);
      out[(long)i * numel + x] = from_f32<T>(a);
    }
  }
}

}  // namespace basd

using namespace basd;
#define ST ((cudaStream_t)stream)

extern "C" int basd_attn_rows(const void* attn, int dtype, int B, int H, int side, int q_rows,
                              int has_cls, float* rows, void* stream) {
  const int n_tok = has_cls ? side - 1 : side;
  if (dtype == BASD_DTYPE_BF16)
    attn_rows_kernel<__nv_bfloat16><<<B, 256, 0, ST>>>((const __nv_bfloat16*)attn, H, side, q_rows,
                                                       has_cls, n_tok, rows);
  else
    attn_rows_kernel<float><<<B, 256, 0, ST>>>((const float*)attn, H, side, q_rows, has_cls, n_tok,
                                               rows);
  BASD_LAUNCH_CHECK();
  return 0;
}

static int fill_layers(LayerPtrs& lp, const void* const* ptrs, int L) {
  if (L > MAX_L) return -6;
  for (int l = 0; l < L; ++l) lp.p[l] = ptrs[l];
  return 0;
}

// out: (E, B, n_dst, D) in out_dtype. teacher_layers: host array of L device pointers.
extern "C" int basd_mix_interp(const void* const* teacher_layers, int L, int E, const float* weights,
                               int in_dtype, int B, int n_src, int n_dst, int D, void* out,
                               int out_dtype, void* stream) {
  if (E > MAX_E || (D & 7)) return -7;
  LayerPtrs lp;
  if (int rc = fill_layers(lp, teacher_layers, L)) return rc;
  const long total = (long)B * n_dst * (D >> 3);
  const unsigned grid = (unsigned)((total + 255) / 256);
  // six layers in flight when that divides the stack (12, 18, 24 layers), four otherwise
  const bool six = L % 6 == 0;
#define BASD_MIX(TI, TO, NE)                                                                      \
  do {                                                                                            \
    if (six)                                                                                      \
      mix_interp_kernel<TI, TO, NE, 6><<<grid, 256, 0, ST>>>(lp, L, E, weights, B, n_src, n_dst, D, (TO*)out); \
    else                                                                                          \
      mix_interp_kernel<TI, TO, NE, 4><<<grid, 256, 0, ST>>>(lp, L, E, weights, B, n_src, n_dst, D, (TO*)out); \
  } while (0)
#define BASD_MIX_E(TI, TO)              \
  do {                                  \
    if (E <= 1) BASD_MIX(TI, TO, 1);    \
    else if (E <= 2) BASD_MIX(TI, TO, 2); \
    else if (E <= 4) BASD_MIX(TI, TO, 4); \
    else BASD_MIX(TI, TO, 8);           \
  } while (0)
  if (in_dtype == BASD_DTYPE_BF16 && out_dtype == BASD_DTYPE_BF16)
    BASD_MIX_E(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == BASD_DTYPE_BF16)
    BASD_MIX_E(__nv_bfloat16, float);
  else if (out_dtype == BASD_DTYPE_F32)
    BASD_MIX_E(float, float);
  else
    return -8;
#undef BASD_MIX_E
#undef BASD_MIX
  BASD_LAUNCH_CHECK();
  return 0;
}

// out: (E, numel) in the layers' dtype.  Layer pointers must be 16-byte aligned (torch allocations are;
// otherwise -9); outputs are written element-wise, so any numel works.
extern "C" int basd_mix_flat(const void* const* layers, int L, int E, const float* weights, int dtype,
                             long numel, void* out, void* stream) {
  if (E > MAX_E || E < 1) return -7;
  LayerPtrs lp;
  if (int rc = fill_layers(lp, layers, L)) return rc;
  const int elt = dtype == BASD_DTYPE_BF16 ? 2 : 4;
  uintptr_t bits = 0;
  for (int l = 0; l < L; ++l) bits |= reinterpret_cast<uintptr_t>(layers[l]);
  if (bits & 15) return -9;
  const long groups = numel / (16 / elt);
  const unsigned grid = (unsigned)min((long)148 * 8, max(1L, (groups + 255) / 256));
#define BASD_FLAT(T)                                                                                 \
  do {                                                                                               \
    if (E <= 1) mix_flat_kernel<T, 1><<<grid, 256, 0, ST>>>(lp, L, E, weights, numel, (T*)out);      \
    else if (E <= 2) mix_flat_kernel<T, 2><<<grid, 256, 0, ST>>>(lp, L, E, weights, numel, (T*)out); \
    else if (E <= 4) mix_flat_kernel<T, 4><<<grid, 256, 0, ST>>>(lp, L, E, weights, numel, (T*)out); \
    else mix_flat_kernel<T, 8><<<grid, 256, 0, ST>>>(lp, L, E, weights, numel, (T*)out);             \
  } while (0)
  if (dtype == BASD_DTYPE_BF16) BASD_FLAT(__nv_bfloat16);
  else BASD_FLAT(float);
#undef BASD_FLAT
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_mix_rows(const float* rows, const float* weights, int E, int L, int B, int n_src,
                             int n_dst, float* w_out, float* totals, void* stream) {
  dim3 grid(B, E);
  mix_rows_kernel<<<grid, 128, 0, ST>>>(rows, weights, L, B, n_src, n_dst, w_out, totals);
  BASD_LAUNCH_CHECK();
  return 0;
}

extern "C" int basd_weight_grad_slices(void) { return 148 * 2; }

// partial: (slices, L, E) scratch, d_weights: (E, L).  n_src: token count of the teacher layers,
// n_rows: length of the importance rows (the attention map's own token count; it differs from n_src
// when the caller hands over tokens that were already resampled, relational.py:29-32).
extern "C" int basd_weight_grad(const void* const* teacher_layers, int L, int E, const float* Z,
                                const float* gw, const float* rows, int in_dtype, int B, int n_src,
                                int n_rows, int n_dst, int D, float gw_scale, const float* gw_scale_dev,
                                float* partial, float* d_weights, void* stream) {
  if (E > MAX_E || (D & 7)) return -7;
  LayerPtrs lp;
  if (int rc = fill_layers(lp, teacher_layers, L)) return rc;
  const int slices = basd_weight_grad_slices();
  // E <= 4: one pass over the upstream gradient, 12 teacher layers per thread (0.96 -> 0.59 ms at C2 on
  // B200, DRAM traffic 2.4x -> ~1x the algorithmic bytes); more extraction points take the per-layer grid
  if (E <= 4) {
    constexpr int LC = 12;
    dim3 g1(slices, (L + LC - 1) / LC);
    if (in_dtype == BASD_DTYPE_BF16 && n_src == n_dst)
      weight_grad_onepass_kernel<__nv_bfloat16, 4, LC, true><<<g1, 256, 0, ST>>>(lp, L, E, Z, B, n_src, n_dst, D, partial);
    else if (in_dtype == BASD_DTYPE_BF16)
      weight_grad_onepass_kernel<__nv_bfloat16, 4, LC, false><<<g1, 256, 0, ST>>>(lp, L, E, Z, B, n_src, n_dst, D, partial);
    else
      weight_grad_onepass_kernel<float, 4, LC, false><<<g1, 256, 0, ST>>>(lp, L, E, Z, B, n_src, n_dst, D, partial);
    BASD_LAUNCH_CHECK();
    dim3 fg(L, E);
    weight_grad_finish_kernel<<<fg, 1024, 0, ST>>>(partial, slices, gw, rows, E, L, B, n_rows, n_dst, gw_scale,
                                                 gw_scale_dev, d_weights);
    BASD_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(slices, L);
  if (in_dtype == BASD_DTYPE_BF16)
    weight_grad_kernel<__nv_bfloat16><<<grid, 256, 0, ST>>>(lp, E, Z, B, n_src, n_dst, D, partial);
  else
    weight_grad_kernel<float><<<grid, 256, 0, ST>>>(lp, E, Z, B, n_src, n_dst, D, partial);
  BASD_LAUNCH_CHECK();
  dim3 fgrid(L, E);
  weight_grad_finish_kernel<<<fgrid, 1024, 0, ST>>>(partial, slices, gw, rows, E, L, B, n_rows, n_dst,
                                                   gw_scale, gw_scale_dev, d_weights);
  BASD_LAUNCH_CHECK();
  return 0;
}

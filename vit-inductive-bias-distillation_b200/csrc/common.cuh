// Shared device/host helpers for the BASD sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define BASD_DTYPE_F32 0
#define BASD_DTYPE_BF16 1

// Every launcher returns the CUDA error of its launch (0 == cudaSuccess); nothing syncs.
#define BASD_LAUNCH_CHECK()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return (int)e__;                  \
  } while (0)

#define BASD_CUDA(x)                                          \
  do {                                                        \
    cudaError_t e__ = (x);                                    \
    if (e__ != cudaSuccess) return (int)e__;                  \
  } while (0)

namespace basd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `scratch` needs >= 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? scratch[lane] : -3.4e38f;
  r = warp_max(r);
  return r;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 128-bit streaming loads of 8 bf16 / 4 fp32, returned as floats.
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float out[8]) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float(w[i] << 16);
    out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// gemm_simt.cu helpers reused by gram_tc.cu
int launch_gram_reduce(const float* partial, int slices, int D, int tile, float* gram,
                       cudaStream_t st, const float* mu0 = nullptr, const float* dsum = nullptr,
                       float coef = 0.f);
int launch_colsum_fold(const float* partial, int slices, int D, float* out, const float* mu0, float coef,
                       cudaStream_t st);

// jacobi_oe8.cu: register-resident Jacobi with eight rows per 16-lane group (<= 224 x 224 active);
// returns -100 when the shape does not fit.
int launch_jacobi_oe8(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                      float tol, float stop2, int max_sweeps, int* sweeps_out, cudaStream_t st, int dim_lo,
                      int dim_hi, int* rot_out, int rows_only = 0);

int launch_jacobi_oe8_cluster(float* G, int n, int m, int ld, long stride, int batch, const int* dims,
                              float tol, float stop2, int max_sweeps, int* sweeps_out, cudaStream_t st, int dim_lo,
                              int dim_hi, int* rot_out);

}  // namespace basd

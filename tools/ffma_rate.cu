// FP32 FMA issue rate on this GPU: scalar FFMA with three register operands against packed FFMA2
// (fma.rn.f32x2, sm_100+), in the two operand patterns of the Jacobi sweep (dot product: two fresh
// operands + accumulator; rotation: one broadcast multiplier + two fresh operands).
// Prints FMAs per clock per SM (128 = the nominal FP32 peak).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 tools/ffma_rate.cu -o build/ffma_rate
#include <cuda_runtime.h>
#include <cstdio>

constexpr int NV = 24;      // floats per thread per operand array
constexpr int ITERS = 2048;

template <bool DESC>
__device__ __forceinline__ void rot_scalar(float (&x)[NV], float (&y)[NV], float t1, float t2) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = DESC ? NV - 1 - k : k;
    const float a = x[i], b = y[i];
    x[i] = fmaf(t1, a, b);
    y[i] = fmaf(-t2, b, a);
  }
}
template <bool DESC>
__device__ __forceinline__ void rot_packed(float2 (&x)[NV / 2], float2 (&y)[NV / 2], float t1, float t2) {
  const float2 T1 = make_float2(t1, t1), T2 = make_float2(-t2, -t2);
#pragma unroll
  for (int k = 0; k < NV / 2; ++k) {
    const int i = DESC ? NV / 2 - 1 - k : k;
    const float2 a = x[i], b = y[i];
    x[i] = __ffma2_rn(T1, a, b);
    y[i] = __ffma2_rn(T2, b, a);
  }
}

// MODE 0/1: rotation scalar / packed (ascending + descending pass per iteration: closed register permutation)
// MODE 2/3: dot products scalar / packed (two accumulators each), one dependent FMA per iteration feeds back
template <int MODE>
__global__ void __launch_bounds__(512, 1) rate_kernel(float* out, float seed, long long* clocks) {
  float x[NV], y[NV];
  float2 px[NV / 2], py[NV / 2];
#pragma unroll
  for (int i = 0; i < NV; ++i) { x[i] = seed + i + threadIdx.x; y[i] = seed * 0.5f - i; }
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) { px[i] = make_float2(x[2 * i], x[2 * i + 1]); py[i] = make_float2(y[2 * i], y[2 * i + 1]); }
  float t1 = seed * 1e-3f, t2 = seed * 2e-3f;
  float acc = 0.f;
  const long long c0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
      rot_scalar<false>(x, y, t1, t2);
      rot_scalar<true>(x, y, t1, t2);
    } else if (MODE == 1) {
      rot_packed<false>(px, py, t1, t2);
      rot_packed<true>(px, py, t1, t2);
    } else if (MODE == 2) {
      float g0 = 0.f, g1 = 0.f, h0 = 0.f, h1 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        g0 = fmaf(x[i], y[i], g0);
        g1 = fmaf(x[i + 1], y[i + 1], g1);
        h0 = fmaf(x[i], x[i + 1], h0);
        h1 = fmaf(y[i], y[i + 1], h1);
      }
      acc = (g0 + g1) + (h0 + h1);
      x[0] = fmaf(acc, 1e-30f, x[0]);
    } else {
      float2 g = make_float2(0.f, 0.f), h = make_float2(0.f, 0.f), u = g, v = g;
#pragma unroll
      for (int i = 0; i < NV / 2; i += 2) {
        g = __ffma2_rn(px[i], py[i], g);
        h = __ffma2_rn(px[i + 1], py[i + 1], h);
        u = __ffma2_rn(px[i], px[i + 1], u);
        v = __ffma2_rn(py[i], py[i + 1], v);
      }
      acc = (g.x + g.y) + (h.x + h.y) + (u.x + u.y) + (v.x + v.y);
      px[0].x = fmaf(acc, 1e-30f, px[0].x);
    }
  }
  const long long c1 = clock64();
  float s = acc;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += x[i] + y[i];
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) s += px[i].x + px[i].y + py[i].x + py[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = c1 - c0;
}

template <int MODE>
static void run(const char* name, int threads, float* out, long long* clk, int sms) {
  rate_kernel<MODE><<<sms, threads>>>(out, 1.0f, clk);
  cudaDeviceSynchronize();
  rate_kernel<MODE><<<sms, threads>>>(out, 1.0f, clk);
  cudaDeviceSynchronize();
  long long h[1024];
  cudaMemcpy(h, clk, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += h[i];
  mean /= sms;
  const double fmas = (MODE < 2 ? 4.0 : 2.0) * NV * ITERS * threads;
  std::printf("%-28s %4d threads/SM: %7.1f FMA/clk/SM (%.0f clocks)\n", name, threads, fmas / mean, mean);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* clk;
  cudaMalloc(&out, sms * 512 * sizeof(float));
  cudaMalloc(&clk, sms * sizeof(long long));
  for (int threads : {128, 256, 416, 512}) {
    run<0>("rotation scalar FFMA", threads, out, clk, sms);
    run<1>("rotation packed FFMA2", threads, out, clk, sms);
    run<2>("dot scalar FFMA", threads, out, clk, sms);
    run<3>("dot packed FFMA2", threads, out, clk, sms);
  }
  std::printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Torch-free timing / determinism check of the batched row Jacobi (basd_jacobi_rows) through the C ABI.
// Starts in about a second on a fresh GPU box, where `import torch` alone takes a minute, so an A/B of
// two builds or two environment switches fits in a sub-minute GPU call:
//   build/jacobi_check <out.bin> [batch n m]   prints ms per launch, mean sweep count and the largest
//   |cos| between the resulting rows, writes rows + sweep counts to <out.bin>; `cmp` two outputs for a
//   bitwise A/B (the sweep is deterministic).  `build/jacobi_check chol [batch n inner]` times the pivoted
//   Cholesky the same way and checks LT^T LT against K.  Used at the end of round 1 for the neighbour-only
//   synchronisation experiment.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude tools/jacobi_check.cu \
//        -o build/jacobi_check -Lvit-inductive-bias-distillation_b200/lib -lbasd_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../vit-inductive-bias-distillation_b200/lib'
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>
#include "basd_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

// chol mode: build/jacobi_check chol [batch n inner]  -- K = A A^T with A (n x inner), graded columns;
// times basd_pivoted_cholesky and checks LT^T LT against K (first and last problem) and the ranks.
static int chol_main(int argc, char** argv) {
  const int batch = argc > 2 ? std::atoi(argv[2]) : 1024;
  const int n = argc > 3 ? std::atoi(argv[3]) : 196;
  const int inner = argc > 4 ? std::atoi(argv[4]) : 384;
  const size_t per = (size_t)n * n, total = per * batch;
  std::vector<float> hk(total);
  unsigned long long s = 1234567891234567ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((s >> 11) * (1.0 / 9007199254740992.0)) * 2.f - 1.f; };
  std::vector<double> a((size_t)n * inner);
  for (int b = 0; b < batch; ++b) {
    if (b < 4 || b == batch - 1) {                       // a few distinct problems, the rest are copies
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < inner; ++k) a[(size_t)i * inner + k] = rnd() * std::pow(10.0, -2.0 * k / inner);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
          double d = 0;
          for (int k = 0; k < inner; ++k) d += a[(size_t)i * inner + k] * a[(size_t)j * inner + k];
          hk[b * per + (size_t)i * n + j] = hk[b * per + (size_t)j * n + i] = (float)d;
        }
    } else {
      std::copy(hk.begin() + (b % 4) * per, hk.begin() + (b % 4 + 1) * per, hk.begin() + b * per);
    }
  }
  float *d_k = nullptr, *d_work = nullptr, *d_lt = nullptr;
  int* d_rank = nullptr;
  CK(cudaMalloc(&d_k, total * sizeof(float)));
  CK(cudaMalloc(&d_work, total * sizeof(float)));
  CK(cudaMalloc(&d_lt, total * sizeof(float)));
  CK(cudaMalloc(&d_rank, batch * sizeof(int)));
  CK(cudaMemcpy(d_k, hk.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaMemcpy(d_work, d_k, total * sizeof(float), cudaMemcpyDeviceToDevice));   // K is destroyed by contract
    CK(cudaMemset(d_lt, 0xff, total * sizeof(float)));
    CK(cudaEventRecord(e0, 0));
    const int rc = basd_pivoted_cholesky(d_work, n, n, (long)per, d_lt, n, (long)per, batch, 1e-5f, d_rank, nullptr, nullptr);
    CK(cudaEventRecord(e1, 0));
    if (rc) { std::fprintf(stderr, "basd_pivoted_cholesky rc %d\n", rc); return 3; }
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  std::vector<float> lt(total);
  std::vector<int> rank(batch);
  CK(cudaMemcpy(lt.data(), d_lt, total * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rank.data(), d_rank, batch * sizeof(int), cudaMemcpyDeviceToHost));
  double worst = 0, kmax = 0;
  for (int b : {0, batch - 1}) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double d = 0;
        for (int k = 0; k < n; ++k) d += (double)lt[b * per + (size_t)k * n + i] * lt[b * per + (size_t)k * n + j];
        worst = std::fmax(worst, std::fabs(d - hk[b * per + (size_t)i * n + j]));
        kmax = std::fmax(kmax, std::fabs((double)hk[b * per + (size_t)i * n + j]));
      }
  }
  long rsum = 0;
  for (int v : rank) rsum += v;
  std::printf("chol %s batch %d n %d inner %d: %.3f ms per launch, rank[0] %d rank[last] %d rank sum %ld, "
              "max |LT^T LT - K| / max |K| = %.2e\n", "lib", batch, n, inner,
              best, rank[0], rank[batch - 1], rsum, worst / kmax);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "chol") return chol_main(argc, argv);
  const char* out = argc > 1 ? argv[1] : "/tmp/jacobi_check.bin";
  const int batch = argc > 2 ? std::atoi(argv[2]) : 1024;
  const int n = argc > 3 ? std::atoi(argv[3]) : 196;
  const int m = argc > 4 ? std::atoi(argv[4]) : 196;
  // optional 5th argument k: square problems allocated n x n (n == m) with a device-side active size k,
  // the shape of the k x k principal-angle SVDs (e.g. `48 384 384 174`)
  const int kdim = argc > 5 ? std::atoi(argv[5]) : 0;
  // optional 6th argument r (with k = 0): only the first r rows are non-zero and the rank-aware entry point
  // basd_jacobi_rows_ranked is called with row_dims = r (C3: a 49-token teacher resampled to 196: `1024 196 196 0 48`)
  const int rdim = argc > 6 ? std::atoi(argv[6]) : 0;
  const size_t per = (size_t)n * m, total = per * batch;
  std::vector<float> host(total);
  // graded rows (like a product of two pivoted-Cholesky factors) plus a dense coupling
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((s >> 11) * (1.0 / 9007199254740992.0)) * 2.f - 1.f; };
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < n; ++i) {
      const float scale = std::pow(10.f, -3.f * i / n);
      for (int j = 0; j < m; ++j) {
        const float v = scale * (rnd() + (i == j ? 2.f : 0.f));
        host[b * per + (size_t)i * m + j] = ((kdim && (i >= kdim || j >= kdim)) || (rdim && i >= rdim)) ? 0.f : v;
      }
    }
  float *d_in = nullptr, *d_work = nullptr;
  int* d_sweeps = nullptr;
  CK(cudaMalloc(&d_in, total * sizeof(float)));
  CK(cudaMalloc(&d_work, total * sizeof(float)));
  CK(cudaMalloc(&d_sweeps, batch * sizeof(int)));
  int* d_dims = nullptr;
  if (kdim || rdim) {
    std::vector<int> hd(batch, kdim ? kdim : rdim);
    CK(cudaMalloc(&d_dims, batch * sizeof(int)));
    CK(cudaMemcpy(d_dims, hd.data(), batch * sizeof(int), cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(d_in, host.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaMemcpy(d_work, d_in, total * sizeof(float), cudaMemcpyDeviceToDevice));
    CK(cudaMemset(d_sweeps, 0, batch * sizeof(int)));
    CK(cudaEventRecord(e0, 0));
    const int rc = rdim ? basd_jacobi_rows_ranked(d_work, n, m, m, (long)per, batch, d_dims, 1e-6f, 18, d_sweeps,
                                                  nullptr, nullptr)
                        : basd_jacobi_rows(d_work, n, m, m, (long)per, batch, d_dims, 1e-6f, 18, d_sweeps, nullptr);
    CK(cudaEventRecord(e1, 0));
    if (rc) { std::fprintf(stderr, "basd_jacobi_rows rc %d\n", rc); return 3; }
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  std::vector<int> sweeps(batch);
  CK(cudaMemcpy(host.data(), d_work, total * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sweeps.data(), d_sweeps, batch * sizeof(int), cudaMemcpyDeviceToHost));
  double mean = 0, worst = 0;
  for (int v : sweeps) mean += v;
  // orthogonality of the first problem's rows (largest |cos| between rows above 1e-6 of the largest)
  {
    std::vector<double> nr(n);
    double mx = 0;
    for (int i = 0; i < n; ++i) { double a = 0; for (int j = 0; j < m; ++j) a += (double)host[(size_t)i * m + j] * host[(size_t)i * m + j]; nr[i] = std::sqrt(a); mx = std::fmax(mx, nr[i]); }
    for (int i = 0; i < n; ++i) for (int k = i + 1; k < n; ++k) {
      if (nr[i] < 1e-6 * mx || nr[k] < 1e-6 * mx) continue;
      double d = 0; for (int j = 0; j < m; ++j) d += (double)host[(size_t)i * m + j] * host[(size_t)k * m + j];
      worst = std::fmax(worst, std::fabs(d) / (nr[i] * nr[k]));
    }
  }
  std::printf("jacobi %s batch %d n %d m %d k %d r %d: %.3f ms per launch, sweeps mean %.2f, max |cos| %.2e\n",
              "lib", batch, n, m, kdim, rdim, best,
              mean / batch, worst);
  FILE* f = std::fopen(out, "wb");
  if (!f) return 4;
  std::fwrite(host.data(), sizeof(float), total, f);
  std::fwrite(sweeps.data(), sizeof(int), batch, f);
  std::fclose(f);
  return 0;
}

// Native (no Python) check of the tcgen05 GEMM through the C-ABI.  Test infrastructure only.
//   gemm_check M N K a_mn b_mn act bias res out_f32 [iters] [variant]
//   res: 0 none, 1 bf16, 2 fp32, 3 fp32 in place (out aliases the residual); variant: agb_gemm_set_variant
// Verifies sampled (or all) output entries against a double-precision host dot product of the
// bf16-rounded inputs and, when iters > 0, times the kernel with CUDA events.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../include/autognothi_b200.h"

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (uint32_t)(rng_state >> 32);
}
static inline float urand() { return (rnd() >> 8) * (1.0f / 16777216.0f) - 0.5f; }

static double gelu_ref(double x) { return 0.5 * x * (1.0 + erf(x / sqrt(2.0))); }

int main(int argc, char** argv) {
  if (argc < 10) {
    printf("usage: gemm_check M N K a_mn b_mn act bias res out_f32 [iters] [variant]\n");
    return 1;
  }
  const int M = atoi(argv[1]), N = atoi(argv[2]), K = atoi(argv[3]);
  const int a_mn = atoi(argv[4]), b_mn = atoi(argv[5]), act = atoi(argv[6]);
  const int use_bias = atoi(argv[7]), use_res = atoi(argv[8]), out_f32 = atoi(argv[9]);
  const int iters = argc > 10 ? atoi(argv[10]) : 0;
  const int variant = argc > 11 ? atoi(argv[11]) : 0;
  agb_gemm_set_variant(variant);
  if (use_res >= 2 && !out_f32) { printf("fp32 residual needs out_f32\n"); return 1; }

  // stored shapes: K-major [rows, K]; MN-major [K, rows]
  const size_t a_elems = (size_t)M * K, b_elems = (size_t)N * K;
  const size_t r_elems = use_res ? (size_t)M * N : 1;
  std::vector<__nv_bfloat16> hA(a_elems), hB(b_elems), hR(r_elems);
  std::vector<float> fA(a_elems), fB(b_elems), fR(r_elems), hBias(N);
  for (size_t i = 0; i < a_elems; ++i) { hA[i] = __float2bfloat16(urand()); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < b_elems; ++i) { hB[i] = __float2bfloat16(urand()); fB[i] = __bfloat162float(hB[i]); }
  for (size_t i = 0; i < r_elems; ++i) {
    const float v = urand();
    hR[i] = __float2bfloat16(v);
    fR[i] = use_res >= 2 ? v : __bfloat162float(hR[i]);
  }
  for (int i = 0; i < N; ++i) hBias[i] = urand();

  __nv_bfloat16 *dA, *dB, *dR;
  float *dBias, *dRf;
  void* dC;
  CK(cudaMalloc(&dA, a_elems * 2));
  CK(cudaMalloc(&dB, b_elems * 2));
  CK(cudaMalloc(&dR, r_elems * 2));
  CK(cudaMalloc(&dBias, N * 4));
  CK(cudaMalloc(&dRf, r_elems * 4));
  CK(cudaMemcpy(dRf, fR.data(), r_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dC, (size_t)M * N * (out_f32 ? 4 : 2)));
  CK(cudaMemcpy(dA, hA.data(), a_elems * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), b_elems * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, hR.data(), r_elems * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBias, hBias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * (out_f32 ? 4 : 2)));
  if (use_res == 3) CK(cudaMemcpy(dC, fR.data(), (size_t)M * N * 4, cudaMemcpyHostToDevice));

  const int lda = a_mn ? M : K, ldb = b_mn ? N : K;
  auto run = [&]() {
    return agb_gemm_bf16(dA, lda, a_mn, dB, ldb, b_mn, M, N, K, 1.0f, use_bias ? dBias : nullptr, act,
                         use_res == 1 ? dR : nullptr,
                         use_res == 2 ? dRf : (use_res == 3 ? static_cast<const float*>(dC) : nullptr), N, 0, 0, dC, N,
                         out_f32, nullptr);
  };
  int rc = run();
  if (rc != 0) { printf("agb_gemm_bf16 rc=%d: %s\n", rc, agb_last_error()); return 3; }
  CK(cudaDeviceSynchronize());

  std::vector<float> hC((size_t)M * N);
  if (out_f32) {
    CK(cudaMemcpy(hC.data(), dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> tmp((size_t)M * N);
    CK(cudaMemcpy(tmp.data(), dC, (size_t)M * N * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) hC[i] = __bfloat162float(tmp[i]);
  }

  const size_t total = (size_t)M * N;
  const size_t nsamples = total <= 400000 ? total : 200000;
  double max_err = 0.0, max_ref = 0.0;
  size_t bad = 0;
  for (size_t s = 0; s < nsamples; ++s) {
    size_t idx;
    if (nsamples == total) idx = s;
    else if (s < 4096) idx = (total - 1) - s;           // always cover the tail tile
    else idx = ((size_t)rnd() << 20 ^ rnd()) % total;
    const int m = (int)(idx / N), n = (int)(idx % N);
    double acc = 0.0;
    for (int k = 0; k < K; ++k) {
      const float a = a_mn ? fA[(size_t)k * M + m] : fA[(size_t)m * K + k];
      const float b = b_mn ? fB[(size_t)k * N + n] : fB[(size_t)n * K + k];
      acc += (double)a * b;
    }
    if (use_bias) acc += hBias[n];
    if (act == 1) acc = gelu_ref(acc);
    if (use_res) acc += fR[idx];
    const double err = fabs(acc - hC[idx]);
    const double tol = (out_f32 ? 2e-3 : 1e-2) * (1.0 + fabs(acc));
    if (!(err <= tol)) {
      if (bad < 8) printf("  mismatch at (%d,%d): got %f want %f\n", m, n, hC[idx], acc);
      ++bad;
    }
    if (err > max_err) max_err = err;
    if (fabs(acc) > max_ref) max_ref = fabs(acc);
  }
  printf("gemm v%d M=%d N=%d K=%d a_mn=%d b_mn=%d act=%d bias=%d res=%d f32=%d: checked=%zu bad=%zu max_err=%.3e max_ref=%.3e %s\n",
         variant, M, N, K, a_mn, b_mn, act, use_bias, use_res, out_f32, nsamples, bad, max_err, max_ref,
         bad ? "FAIL" : "PASS");

  if (iters > 0 && !bad) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) run();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) run();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double t = ms / iters * 1e-3;
    printf("  time %.3f us  %.1f TFLOP/s\n", t * 1e6, 2.0 * M * N * K / t * 1e-12);
  }
  return bad ? 4 : 0;
}

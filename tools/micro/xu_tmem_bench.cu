// Micro-benchmarks behind the attention kernel's design (test infrastructure; run under gpurun):
//   (1) MUFU.EX2 throughput per SM sub-partition with 1 / 2 / 4 resident warps per sub-partition, in the instruction mix of
//       the softmax inner loop (FFMA -> MUFU.EX2 -> FADD, F2FP every second element);
//   (2) tcgen05.ld 32x32b.x32 throughput (bytes per clock per SM) with 4 / 8 / 16 warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I autognothi_b200/csrc tools/micro/xu_tmem_bench.cu -o build/xu_tmem_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "agb_common.cuh"

using namespace agb;

template <int MODE>   // 0: FFMA+MUFU+FADD+F2FP interleaved by the compiler   1: MUFU only
__global__ void xu_kernel(float* out, long long* cyc, int iters) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x + j) * 1e-3f;
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a = v[j], b = v[j + 1];
      if (MODE == 0) {
        a = fmaf(a, 0.18f, -1.0f);
        b = fmaf(b, 0.18f, -1.0f);
      }
      a = ex2_approx(a);
      b = ex2_approx(b);
      if (MODE == 0) {
        sum0 += a;
        sum1 += b;
        acc ^= pack_bf16x2(a, b);
      }
      v[j] = a;
      v[j + 1] = b;
    }
  }
  const long long t1 = clock64();
  float s = sum0 + sum1;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// (2b) the softmax chunk as the attention kernel runs it: tcgen05.ld 32 columns -> wait -> row max (FMNMX3) -> vote -> 32 x
//      (FFMA, MUFU.EX2, FADD), 16 x F2FP -> tcgen05.st 16 columns, with 1 / 2 / 4 warps per sub-partition
__global__ void chunk_kernel(float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  // fill my 128 columns with small numbers so that exp2 stays in range
  {
    uint32_t z[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = __float_as_uint(-0.01f * (float)(j + (threadIdx.x & 31)));
    for (int c = 0; c < 128; c += 32) tmem_st32(base + c, z);
    tmem_wait_st();
  }
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  float m_s = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t s[32], pk[16];
    const uint32_t a = base + ((it & 1) ? 64 : 32);
    tmem_ld32(a, s);
    tmem_wait_ld();
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      m0 = fmaxf(m0, __uint_as_float(s[j]));
      m1 = fmaxf(m1, __uint_as_float(s[j + 1]));
      m2 = fmaxf(m2, __uint_as_float(s[j + 2]));
      m3 = fmaxf(m3, __uint_as_float(s[j + 3]));
    }
    const float lm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * 0.18f;
    if (__any_sync(0xffffffffu, lm > m_s + 24.f)) m_s = lm;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float e0 = ex2_approx(fmaf(__uint_as_float(s[2 * j]), 0.18f, -m_s));
      const float e1 = ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), 0.18f, -m_s));
      sum[(j & 1) * 2] += e0;
      sum[(j & 1) * 2 + 1] += e1;
      pk[j] = pack_bf16x2(e0, e1);
    }
    tmem_st16(base, pk);
  }
  tmem_wait_st();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (sum[0] + sum[1]) + (sum[2] + sum[3]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (warp == 0) tmem_dealloc(slot, 512);
}

// (3) the same MUFU loop in warps 8-15 (2 per sub-partition) while warps 0-7 spin on an mbarrier, as the service / epilogue
//     warps of the attention kernel do.  SPIN 0: no spinners (they exit), 1: try_wait loop, 2: try_wait with a suspend-time hint,
//     3: try_wait loop with __nanosleep(64) between polls
template <int SPIN>
__global__ void xu_spin_kernel(float* out, long long* cyc, int iters) {
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 8);
    fence_mbar_init();
  }
  __syncthreads();
  if (warp < 8) {
    if (SPIN == 0) return;
    const uint32_t b = smem_u32(&bar);
    if (SPIN == 1) {
      while (!mbar_try_wait(b, 0)) {}
    } else if (SPIN == 2) {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                     : "=r"(ok) : "r"(b), "r"(0), "r"(20000) : "memory");
      }
    } else {
      while (!mbar_try_wait(b, 0)) __nanosleep(64);
    }
    return;
  }
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (float)(threadIdx.x + j) * 1e-3f;
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a = fmaf(v[j], 0.18f, -1.0f), b = fmaf(v[j + 1], 0.18f, -1.0f);
      a = ex2_approx(a);
      b = ex2_approx(b);
      sum0 += a;
      sum1 += b;
      acc ^= pack_bf16x2(a, b);
      v[j] = a;
      v[j + 1] = b;
    }
  }
  const long long t1 = clock64();
  float s = sum0 + sum1;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
  if (threadIdx.x == 256) cyc[blockIdx.x] = t1 - t0;
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(smem_u32(&bar));
}

// VAR 0: one x32 load, wait   1: one x16 load, wait   2: two x32 loads in flight, wait   3: x32 as 16x256b.x8 (16 lanes x 64 columns)
template <int VAR>
__global__ void ldtm_kernel(float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t s[32], t[32];
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + ((it * 32 + warp * 64) & 448);
    if (VAR == 0) {
      tmem_ld32(a, s);
    } else if (VAR == 1) {
      tmem_ld16(a, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
    } else if (VAR == 2) {
      tmem_ld32(a, s);
      tmem_ld32(a ^ 32, t);
    } else {
      asm volatile(
          "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(s[9]),
            "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]), "=r"(s[16]), "=r"(s[17]), "=r"(s[18]),
            "=r"(s[19]), "=r"(s[20]), "=r"(s[21]), "=r"(s[22]), "=r"(s[23]), "=r"(s[24]), "=r"(s[25]), "=r"(s[26]), "=r"(s[27]),
            "=r"(s[28]), "=r"(s[29]), "=r"(s[30]), "=r"(s[31])
          : "r"(a)
          : "memory");
    }
    tmem_wait_ld();
    acc += __uint_as_float(s[it & 15]);
    if (VAR == 2) acc += __uint_as_float(t[it & 15]);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) xu_kernel<0><<<148, warps * 32>>>(out, cyc, iters);
        else xu_kernel<1><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double per = (double)h[0] / ((double)iters * 32 * (warps / 4));
      printf("xu mode %d (%s): %2d warps/CTA (%d per sub-partition): %8lld cycles, %.2f cycles per MUFU.EX2 warp-instruction per sub-partition "
             "=> %.1f exp2 / clk / SM\n", mode, mode == 0 ? "FFMA+MUFU+FADD+F2FP" : "MUFU only", warps, warps / 4, h[0], per, 32.0 / per * 4);
    }
  }
  for (int warps : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) {
      chunk_kernel<<<148, warps * 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("softmax chunk (tcgen05.ld x32, max, vote, 32 exp2, tcgen05.st x16): %2d warps/CTA (%d per sub-partition): %.1f cycles per chunk per warp, "
           "%.2f cycles per MUFU.EX2 per sub-partition\n", warps, warps / 4, (double)h[0] / iters, (double)h[0] / ((double)iters * 32 * (warps / 4)));
  }
  for (int spin = 0; spin < 4; ++spin) {
    for (int rep = 0; rep < 2; ++rep) {
      if (spin == 0) xu_spin_kernel<0><<<148, 512>>>(out, cyc, iters);
      if (spin == 1) xu_spin_kernel<1><<<148, 512>>>(out, cyc, iters);
      if (spin == 2) xu_spin_kernel<2><<<148, 512>>>(out, cyc, iters);
      if (spin == 3) xu_spin_kernel<3><<<148, 512>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* sn[4] = {"no spinning warps", "8 warps spin on mbarrier.try_wait", "8 warps in try_wait with suspend-time hint", "8 warps poll try_wait + nanosleep(64)"};
    printf("MUFU loop in 8 warps (2 per sub-partition), %-44s: %.2f cycles per MUFU.EX2 warp-instruction per sub-partition\n", sn[spin],
           (double)h[0] / ((double)iters * 32 * 2));
  }
  const char* names[4] = {"32x32b.x32, wait", "32x32b.x16, wait", "2 x 32x32b.x32, wait", "16x256b.x8 (16 lanes x 64 cols), wait"};
  const int bytes[4] = {4096, 2048, 8192, 4096};
  for (int var = 0; var < 4; ++var)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (var == 0) ldtm_kernel<0><<<148, warps * 32>>>(out, cyc, iters);
        if (var == 1) ldtm_kernel<1><<<148, warps * 32>>>(out, cyc, iters);
        if (var == 2) ldtm_kernel<2><<<148, warps * 32>>>(out, cyc, iters);
        if (var == 3) ldtm_kernel<3><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("tcgen05.ld %-38s %2d warps: %8lld cycles for %d iterations per warp => %6.1f B / clk / SM, %7.1f cycles per iteration\n",
             names[var], warps, h[0], iters, (double)iters * warps * bytes[var] / (double)h[0], (double)h[0] / iters);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Backward-pass kernels of explainer training (reference scripts/train_explainer.py:182-198 runs
// autograd over models/vanilla_vit.py:102-130 / models/vanilla_bert.py:123-162; here the adjoints are
// written out): GELU forward/backward, LayerNorm backward, bias (column-sum) gradients, masked-attention
// backward and the embedding adjoints.  The four GEMMs per Linear (dgrad + wgrad) reuse the tcgen05
// kernel through its MN-major operand modes, so these are the HBM-bound remainder plus the attention
// adjoint (CUDA-core v1: no score matrix in HBM, no atomics, deterministic).
#include <algorithm>

#include "agb_common.cuh"

namespace agb {

// ------------------------------------------------------------------------------------------------
// element access helpers (fp32 or bf16 storage, fp32 math)
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

template <typename T> __device__ __forceinline__ float4 ldv4(const T* p);
template <> __device__ __forceinline__ float4 ldv4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ldv4<bf16>(const bf16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}
template <typename T> __device__ __forceinline__ void stv4(T* p, float4 v);
template <> __device__ __forceinline__ void stv4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void stv4<bf16>(bf16* p, float4 v) {
  uint2 pk;
  pk.x = pack_bf16x2(v.x, v.y);
  pk.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = pk;
}

// ------------------------------------------------------------------------------------------------
// GELU (exact erf form) forward / backward, vectorised
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
template <typename T> __device__ __forceinline__ float gelu_fwd_of(float x) { return gelu_erf_exact(x); }
template <> __device__ __forceinline__ float gelu_fwd_of<bf16>(float x) { return gelu_erf_tanhform(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_of(float x) { return gelu_grad(x); }
template <> __device__ __forceinline__ float gelu_grad_of<bf16>(float x) { return gelu_grad_tanhform(x); }

template <typename T>
__global__ void gelu_fwd_kernel(const T* __restrict__ z, T* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = ldv4<T>(z + i * 4);
  v.x = gelu_fwd_of<T>(v.x); v.y = gelu_fwd_of<T>(v.y); v.z = gelu_fwd_of<T>(v.z); v.w = gelu_fwd_of<T>(v.w);
  stv4<T>(out + i * 4, v);
}
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, T* __restrict__ dz, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 g = ldv4<T>(dy + i * 4);
  const float4 v = ldv4<T>(z + i * 4);
  stv4<T>(dz + i * 4, make_float4(g.x * gelu_grad_of<T>(v.x), g.y * gelu_grad_of<T>(v.y), g.z * gelu_grad_of<T>(v.z),
                                  g.w * gelu_grad_of<T>(v.w)));
}

int gelu_fwd(const void* z, void* out, long long n, int is_bf16, cudaStream_t st) {
  AGB_REQUIRE(n >= 0 && (n % 4) == 0, "element count must be a multiple of 4");
  if (n == 0) return AGB_OK;
  AGB_REQUIRE(z && out, "null pointer");
  const long long n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256);
  if (is_bf16) gelu_fwd_kernel<bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(z), static_cast<bf16*>(out), n4);
  else gelu_fwd_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(z), static_cast<float*>(out), n4);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}
int gelu_bwd(const void* dy, const void* z, void* dz, long long n, int is_bf16, cudaStream_t st) {
  AGB_REQUIRE(n >= 0 && (n % 4) == 0, "element count must be a multiple of 4");
  if (n == 0) return AGB_OK;
  AGB_REQUIRE(dy && z && dz, "null pointer");
  const long long n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256);
  if (is_bf16) gelu_bwd_kernel<bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(z), static_cast<bf16*>(dz), n4);
  else gelu_bwd_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(dy), static_cast<const float*>(z), static_cast<float*>(dz), n4);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] (+)= sum_m Y[m, n].  Grid-stride over row slabs, fp32 atomics
// per slab (a few hundred adds per column in total).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ y, long long ld, int M, int N, float* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * rows_per;
  const int m1 = min(M, m0 + rows_per);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += ld1<T>(y + (long long)m * ld + col);
  if (m1 > m0) atomicAdd(out + col, s);
}

// Vectorised variant (16-byte loads: 8 bf16 or 4 fp32 columns per thread).  Block = 32 column-threads x 8 row-threads;
// the 8 row-threads interleave over the block's row slab, combine through shared memory, one atomic per column per block.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ y, long long ld, int M, int N, float* __restrict__ out) {
  __shared__ float part[8][32 * VEC + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * VEC;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  if (col < N) {
    for (int m = m0 + ty; m < m1; m += 8) {
      const uint4 raw = *reinterpret_cast<const uint4*>(y + (long long)m * ld + col);
      if (VEC == 8) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[2 * i] += bf16_lo(w[i]); acc[2 * i + 1] += bf16_hi(w[i]); }
      } else {
        acc[0] += __uint_as_float(raw.x); acc[1] += __uint_as_float(raw.y);
        acc[2] += __uint_as_float(raw.z); acc[3] += __uint_as_float(raw.w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) part[ty][tx * VEC + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += part[r][c];
    const int gc = blockIdx.x * 32 * VEC + c;
    if (gc < N && m1 > m0) atomicAdd(out + gc, s);
  }
}

int colsum(const void* y, int is_bf16, long long ld, int M, int N, float* out, cudaStream_t st) {
  AGB_REQUIRE(M >= 0 && N > 0, "shape");
  if (M == 0) return AGB_OK;
  AGB_REQUIRE(y && out, "null pointer");
  const int vec = is_bf16 ? 8 : 4;
  if ((N % vec) == 0 && (ld % vec) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && M >= 64) {
    const int gx = (N + 32 * vec - 1) / (32 * vec);
    int gy = (4 * sm_count() + gx - 1) / gx;               // ~4 blocks per SM in total
    gy = max(1, min(gy, M / 32));
    dim3 grid(gx, gy);
    if (is_bf16) colsum_vec_kernel<bf16, 8><<<grid, 256, 0, st>>>(static_cast<const bf16*>(y), ld, M, N, out);
    else colsum_vec_kernel<float, 4><<<grid, 256, 0, st>>>(static_cast<const float*>(y), ld, M, N, out);
    AGB_CHECK_CUDA(cudaGetLastError());
    return AGB_OK;
  }
  dim3 grid((N + 127) / 128, M >= 4096 ? 64 : (M >= 256 ? 16 : 1));
  if (is_bf16) colsum_kernel<bf16><<<grid, 128, 0, st>>>(static_cast<const bf16*>(y), ld, M, N, out);
  else colsum_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(y), ld, M, N, out);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  y = (x - mean) * rstd * gamma + beta
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma     (+ dres when given)
//   dgamma += sum_rows dy * xhat,  dbeta += sum_rows dy
// One warp per row for dx; per-CTA partial dgamma/dbeta in shared memory, then one atomic per column.
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TDY>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const TX* __restrict__ x, const TDY* __restrict__ dy, const float* __restrict__ gamma,
                     const float* __restrict__ dres, int rows, int H, float eps, float* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sm[];
  float* sg = sm;       // H
  float* sb = sm + H;   // H
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int row = blockIdx.x * nw + warp; row < rows; row += gridDim.x * nw) {
    const TX* xr = x + (long long)row * H;
    const TDY* gr = dy + (long long)row * H;
    float s = 0.f;
    for (int i = lane; i < H; i += 32) s += ld1<TX>(xr + i);
    const float mean = warp_sum(s) / (float)H;
    float ss = 0.f;
    for (int i = lane; i < H; i += 32) { const float d = ld1<TX>(xr + i) - mean; ss += d * d; }
    const float rstd = rsqrtf(warp_sum(ss) / (float)H + eps);
    float a = 0.f, b = 0.f;
    for (int i = lane; i < H; i += 32) {
      const float xh = (ld1<TX>(xr + i) - mean) * rstd;
      const float g = ld1<TDY>(gr + i) * gamma[i];
      a += g;
      b += g * xh;
    }
    a = warp_sum(a) / (float)H;
    b = warp_sum(b) / (float)H;
    for (int i = lane; i < H; i += 32) {
      const float xh = (ld1<TX>(xr + i) - mean) * rstd;
      const float d = ld1<TDY>(gr + i);
      const float g = d * gamma[i];
      float v = rstd * (g - a - xh * b);
      if (dres) v += dres[(long long)row * H + i];
      dx[(long long)row * H + i] = v;
      atomicAdd(sg + i, d * xh);
      atomicAdd(sb + i, d);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sg[i]);
    if (dbeta) atomicAdd(dbeta + i, sb[i]);
  }
}

// Register-resident variant for fp32 x / dy and H = 128 * NV4 <= 1024: every row is read from HBM once with 16-byte
// loads; dgamma / dbeta partials stay in registers across all rows a warp handles (each lane owns fixed columns), are
// combined across the CTA's warps in shared memory and leave with one atomic per column per CTA.
template <int NV4>
__global__ void __launch_bounds__(256)
layernorm_bwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                         const float* __restrict__ dres, int rows, int H, float eps, float* __restrict__ dx,
                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sm[];      // [2][H]
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float4 gam[NV4], ag[NV4], ab[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    gam[i] = __ldg(reinterpret_cast<const float4*>(gamma + lane * 4 + i * 128));
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invH = 1.0f / (float)H;
  for (int row = blockIdx.x * nw + warp; row < rows; row += gridDim.x * nw) {
    const float* xr = x + (long long)row * H;
    const float* gr = dy + (long long)row * H;
    float4 xv[NV4], dv[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      xv[i] = *reinterpret_cast<const float4*>(xr + lane * 4 + i * 128);
      dv[i] = *reinterpret_cast<const float4*>(gr + lane * 4 + i * 128);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    const float mean = warp_sum(s) * invH;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      ss += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
    const float rstd = rsqrtf(warp_sum(ss) * invH + eps);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;            // xhat
      const float gx = dv[i].x * gam[i].x, gy = dv[i].y * gam[i].y, gz = dv[i].z * gam[i].z, gw = dv[i].w * gam[i].w;
      a += (gx + gy) + (gz + gw);
      b += (gx * xv[i].x + gy * xv[i].y) + (gz * xv[i].z + gw * xv[i].w);
      ag[i].x += dv[i].x * xv[i].x; ag[i].y += dv[i].y * xv[i].y; ag[i].z += dv[i].z * xv[i].z; ag[i].w += dv[i].w * xv[i].w;
      ab[i].x += dv[i].x; ab[i].y += dv[i].y; ab[i].z += dv[i].z; ab[i].w += dv[i].w;
    }
    a = warp_sum(a) * invH;
    b = warp_sum(b) * invH;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      float4 o;
      o.x = rstd * (dv[i].x * gam[i].x - a - xv[i].x * b);
      o.y = rstd * (dv[i].y * gam[i].y - a - xv[i].y * b);
      o.z = rstd * (dv[i].z * gam[i].z - a - xv[i].z * b);
      o.w = rstd * (dv[i].w * gam[i].w - a - xv[i].w * b);
      if (dres) {
        const float4 r = *reinterpret_cast<const float4*>(dres + (long long)row * H + lane * 4 + i * 128);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(dx + (long long)row * H + lane * 4 + i * 128) = o;
    }
  }
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = lane * 4 + i * 128;
    atomicAdd(sm + c + 0, ag[i].x); atomicAdd(sm + c + 1, ag[i].y); atomicAdd(sm + c + 2, ag[i].z); atomicAdd(sm + c + 3, ag[i].w);
    atomicAdd(sm + H + c + 0, ab[i].x); atomicAdd(sm + H + c + 1, ab[i].y); atomicAdd(sm + H + c + 2, ab[i].z);
    atomicAdd(sm + H + c + 3, ab[i].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sm[i]);
    if (dbeta) atomicAdd(dbeta + i, sm[H + i]);
  }
}

int layernorm_bwd(const void* x, int x_bf16, const void* dy, int dy_bf16, const float* gamma, const float* dres,
                  int rows, int H, float eps, float* dx, float* dgamma, float* dbeta, cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && H > 0 && H <= 8192, "LayerNorm shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(x && dy && gamma && dx, "null pointer");
  const int blocks = min((rows + 7) / 8, 4 * sm_count());
  const size_t smem = 2 * (size_t)H * sizeof(float);
  const int nv4 = H / 128;
  const bool reg_ok = !x_bf16 && !dy_bf16 && (H % 128) == 0 && (nv4 <= 4 || nv4 == 6 || nv4 == 8) &&
                      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
                      (dres == nullptr || (reinterpret_cast<uintptr_t>(dres) & 15) == 0);
  if (reg_ok) {
    const int rb = min((rows + 7) / 8, 2 * sm_count());
#define LNR(NV4)                                                                                              \
  layernorm_bwd_reg_kernel<NV4><<<rb, 256, smem, st>>>(static_cast<const float*>(x), static_cast<const float*>(dy), \
                                                       gamma, dres, rows, H, eps, dx, dgamma, dbeta)
    switch (nv4) {
      case 1: LNR(1); break;
      case 2: LNR(2); break;
      case 3: LNR(3); break;
      case 4: LNR(4); break;
      case 6: LNR(6); break;
      default: LNR(8); break;
    }
#undef LNR
    AGB_CHECK_CUDA(cudaGetLastError());
    return AGB_OK;
  }
#define LNB(TX, TDY)                                                                                   \
  layernorm_bwd_kernel<TX, TDY><<<blocks, 256, smem, st>>>(static_cast<const TX*>(x), static_cast<const TDY*>(dy), \
                                                           gamma, dres, rows, H, eps, dx, dgamma, dbeta)
  if (x_bf16 && dy_bf16) LNB(bf16, bf16);
  else if (x_bf16) LNB(bf16, float);
  else if (dy_bf16) LNB(float, bf16);
  else LNB(float, float);
#undef LNB
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Masked attention backward (CUDA cores, fp32 math, fp32|bf16 I/O), one CTA per (row, head), T <= 256.
//   P = softmax(X), X_ij = m_j ? s_ij/sqrt(d) : 0 (mul0) | -inf (neginf);  O = P V
//   dV_j = sum_i P_ij dO_i;  dP_ij = dO_i . V_j;  D_i = sum_j P_ij dP_ij
//   dS_ij = m_j * P_ij (dP_ij - D_i) / sqrt(d);  dQ_i = sum_j dS_ij K_j;  dK_j = sum_i dS_ij Q_i
// (a masked key's logit is the constant 0 in mul0 mode, so it passes no gradient to Q/K but its V row
// still receives P_ij dO_i).  Pass A is query-owned (row statistics, dQ), pass B is key-owned (dK, dV)
// and recomputes the scores — nothing T x T is stored and no atomics are needed.
// ------------------------------------------------------------------------------------------------
constexpr int AB_D = 64;
constexpr int AB_LD = AB_D + 2;   // padded row pitch (bf16 elements) -> conflict-free row-per-lane reads

template <typename TIO>
__global__ void __launch_bounds__(256)
attention_bwd_kernel(const TIO* __restrict__ qkv, const TIO* __restrict__ dctx, const uint32_t* __restrict__ mask,
                     int words, int T, int H, int heads, int mode, TIO* __restrict__ dqkv) {
  extern __shared__ uint8_t smraw[];
  bf16* sQ = reinterpret_cast<bf16*>(smraw);
  bf16* sK = sQ + T * AB_LD;
  bf16* sV = sK + T * AB_LD;
  bf16* sdO = sV + T * AB_LD;
  float* lse = reinterpret_cast<float*>(sdO + T * AB_LD);  // T   (max + log-sum)
  float* Dv = lse + T;                                      // T
  float* strips = Dv + T;                                   // nw * 2 * T
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x / heads, head = blockIdx.x % heads;
  const long long base = (long long)row * T * 3 * H;
  const uint32_t* mrow = mask + (long long)row * words;
  const float scale = 0.125f;  // 1/sqrt(64)
  // stage Q, K, V, dO (bf16 in smem; fp32 inputs are rounded here only for the staging of the bf16 path —
  // the fp32 instantiation keeps a float copy instead, see sF below)
  for (int e = threadIdx.x; e < T * AB_D; e += blockDim.x) {
    const int t = e / AB_D, c = e % AB_D;
    const TIO* p = qkv + base + (long long)t * 3 * H + head * AB_D + c;
    sQ[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(p));
    sK[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(p + H));
    sV[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(p + 2 * H));
    sdO[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(dctx + ((long long)row * T + t) * H + head * AB_D + c));
  }
  __syncthreads();
  float* s0 = strips + warp * 2 * T;
  float* s1 = s0 + T;
  // ---------------- pass A: query-owned ----------------
  for (int i = warp; i < T; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll 8
      for (int c = 0; c < AB_D; ++c) {
        a = fmaf(__bfloat162float(sQ[i * AB_LD + c]), __bfloat162float(sK[j * AB_LD + c]), a);
        b = fmaf(__bfloat162float(sdO[i * AB_LD + c]), __bfloat162float(sV[j * AB_LD + c]), b);
      }
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      s0[j] = x;
      s1[j] = b;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float z = 0.f;
    for (int j = lane; j < T; j += 32) z += expf(s0[j] - mx);
    z = warp_sum(z);
    const float l = mx + logf(z);
    float dsum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float pj = expf(s0[j] - l);
      dsum += pj * s1[j];
      s0[j] = pj;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) { lse[i] = l; Dv[i] = dsum; }
    __syncwarp();
    // dS strip, then dQ_i[c] = sum_j dS_ij K_j[c]  (lanes over the head dim)
    for (int j = lane; j < T; j += 32) {
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      s0[j] = keep ? s0[j] * (s1[j] - dsum) * scale : 0.f;
    }
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float ds = s0[j];
      q0 = fmaf(ds, __bfloat162float(sK[j * AB_LD + lane]), q0);
      q1 = fmaf(ds, __bfloat162float(sK[j * AB_LD + lane + 32]), q1);
    }
    TIO* dq = dqkv + base + (long long)i * 3 * H + head * AB_D;
    st1<TIO>(dq + lane, q0);
    st1<TIO>(dq + lane + 32, q1);
    __syncwarp();
  }
  __syncthreads();
  // ---------------- pass B: key-owned ----------------
  for (int j = warp; j < T; j += nw) {
    const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
    for (int i = lane; i < T; i += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll 8
      for (int c = 0; c < AB_D; ++c) {
        a = fmaf(__bfloat162float(sQ[i * AB_LD + c]), __bfloat162float(sK[j * AB_LD + c]), a);
        b = fmaf(__bfloat162float(sdO[i * AB_LD + c]), __bfloat162float(sV[j * AB_LD + c]), b);
      }
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      const float pij = expf(x - lse[i]);
      s0[i] = pij;                                               // for dV
      s1[i] = keep ? pij * (b - Dv[i]) * scale : 0.f;            // dS for dK
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < T; ++i) {
      const float pij = s0[i], ds = s1[i];
      k0 = fmaf(ds, __bfloat162float(sQ[i * AB_LD + lane]), k0);
      k1 = fmaf(ds, __bfloat162float(sQ[i * AB_LD + lane + 32]), k1);
      v0 = fmaf(pij, __bfloat162float(sdO[i * AB_LD + lane]), v0);
      v1 = fmaf(pij, __bfloat162float(sdO[i * AB_LD + lane + 32]), v1);
    }
    TIO* dk = dqkv + base + (long long)j * 3 * H + H + head * AB_D;
    TIO* dv = dqkv + base + (long long)j * 3 * H + 2 * H + head * AB_D;
    st1<TIO>(dk + lane, k0); st1<TIO>(dk + lane + 32, k1);
    st1<TIO>(dv + lane, v0); st1<TIO>(dv + lane + 32, v1);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Long sequences (256 < T <= 512, e.g. the reference's checked-in BERT configuration with 512 positions): the four
// T x 64 operands no longer fit in shared memory together, so the two passes become two kernels.
//   A (query-owned): K, V of the unit staged + a 64-query block -> row statistics (to a global scratch), dQ
//   B (key-owned)  : Q, dO of the unit staged + a 64-key block  -> dK, dV   (re-computes the scores)
// Same arithmetic as attention_bwd_kernel.  grid = (ceil(T / 64), rows * heads).
// ------------------------------------------------------------------------------------------------
constexpr int ABL_BLK = 64;

template <typename TIO>
__global__ void __launch_bounds__(256)
attention_bwd_long_a_kernel(const TIO* __restrict__ qkv, const TIO* __restrict__ dctx, const uint32_t* __restrict__ mask,
                            int words, int T, int H, int heads, int mode, TIO* __restrict__ dqkv,
                            float* __restrict__ stat /* [rows*heads][2][T] */) {
  extern __shared__ uint8_t smraw[];
  bf16* sK = reinterpret_cast<bf16*>(smraw);
  bf16* sV = sK + T * AB_LD;
  bf16* sQ = sV + T * AB_LD;                 // ABL_BLK rows
  bf16* sdO = sQ + ABL_BLK * AB_LD;          // ABL_BLK rows
  float* strips = reinterpret_cast<float*>(sdO + ABL_BLK * AB_LD);   // nw * 2 * T
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int unit = blockIdx.y, row = unit / heads, head = unit % heads;
  const int i0 = blockIdx.x * ABL_BLK, i1 = min(T, i0 + ABL_BLK);
  const long long base = (long long)row * T * 3 * H;
  const uint32_t* mrow = mask + (long long)row * words;
  const float scale = 0.125f;
  for (int e = threadIdx.x; e < T * AB_D; e += blockDim.x) {
    const int t = e / AB_D, c = e % AB_D;
    const TIO* p = qkv + base + (long long)t * 3 * H + head * AB_D + c;
    sK[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(p + H));
    sV[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(p + 2 * H));
  }
  for (int e = threadIdx.x; e < (i1 - i0) * AB_D; e += blockDim.x) {
    const int t = e / AB_D, c = e % AB_D;
    sQ[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(qkv + base + (long long)(i0 + t) * 3 * H + head * AB_D + c));
    sdO[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(dctx + ((long long)row * T + i0 + t) * H + head * AB_D + c));
  }
  __syncthreads();
  float* s0 = strips + warp * 2 * T;
  float* s1 = s0 + T;
  for (int i = i0 + warp; i < i1; i += nw) {
    const int il = i - i0;
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll 8
      for (int c = 0; c < AB_D; ++c) {
        a = fmaf(__bfloat162float(sQ[il * AB_LD + c]), __bfloat162float(sK[j * AB_LD + c]), a);
        b = fmaf(__bfloat162float(sdO[il * AB_LD + c]), __bfloat162float(sV[j * AB_LD + c]), b);
      }
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      s0[j] = x;
      s1[j] = b;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float z = 0.f;
    for (int j = lane; j < T; j += 32) z += expf(s0[j] - mx);
    z = warp_sum(z);
    const float l = mx + logf(z);
    float dsum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float pj = expf(s0[j] - l);
      dsum += pj * s1[j];
      s0[j] = pj;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) {
      stat[((long long)unit * 2 + 0) * T + i] = l;
      stat[((long long)unit * 2 + 1) * T + i] = dsum;
    }
    __syncwarp();
    for (int j = lane; j < T; j += 32) {
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      s0[j] = keep ? s0[j] * (s1[j] - dsum) * scale : 0.f;
    }
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float ds = s0[j];
      q0 = fmaf(ds, __bfloat162float(sK[j * AB_LD + lane]), q0);
      q1 = fmaf(ds, __bfloat162float(sK[j * AB_LD + lane + 32]), q1);
    }
    TIO* dq = dqkv + base + (long long)i * 3 * H + head * AB_D;
    st1<TIO>(dq + lane, q0);
    st1<TIO>(dq + lane + 32, q1);
    __syncwarp();
  }
}

template <typename TIO>
__global__ void __launch_bounds__(256)
attention_bwd_long_b_kernel(const TIO* __restrict__ qkv, const TIO* __restrict__ dctx, const uint32_t* __restrict__ mask,
                            int words, int T, int H, int heads, int mode, TIO* __restrict__ dqkv,
                            const float* __restrict__ stat) {
  extern __shared__ uint8_t smraw[];
  bf16* sQ = reinterpret_cast<bf16*>(smraw);
  bf16* sdO = sQ + T * AB_LD;
  bf16* sK = sdO + T * AB_LD;                // ABL_BLK rows
  bf16* sV = sK + ABL_BLK * AB_LD;           // ABL_BLK rows
  float* lse = reinterpret_cast<float*>(sV + ABL_BLK * AB_LD);   // T
  float* Dv = lse + T;                                            // T
  float* strips = Dv + T;                                         // nw * 2 * T
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int unit = blockIdx.y, row = unit / heads, head = unit % heads;
  const int j0 = blockIdx.x * ABL_BLK, j1 = min(T, j0 + ABL_BLK);
  const long long base = (long long)row * T * 3 * H;
  const uint32_t* mrow = mask + (long long)row * words;
  const float scale = 0.125f;
  for (int e = threadIdx.x; e < T * AB_D; e += blockDim.x) {
    const int t = e / AB_D, c = e % AB_D;
    sQ[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(qkv + base + (long long)t * 3 * H + head * AB_D + c));
    sdO[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(dctx + ((long long)row * T + t) * H + head * AB_D + c));
  }
  for (int e = threadIdx.x; e < (j1 - j0) * AB_D; e += blockDim.x) {
    const int t = e / AB_D, c = e % AB_D;
    const TIO* p = qkv + base + (long long)(j0 + t) * 3 * H + head * AB_D + c;
    sK[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(p + H));
    sV[t * AB_LD + c] = __float2bfloat16(ld1<TIO>(p + 2 * H));
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    lse[t] = stat[((long long)unit * 2 + 0) * T + t];
    Dv[t] = stat[((long long)unit * 2 + 1) * T + t];
  }
  __syncthreads();
  float* s0 = strips + warp * 2 * T;
  float* s1 = s0 + T;
  for (int j = j0 + warp; j < j1; j += nw) {
    const int jl = j - j0;
    const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
    for (int i = lane; i < T; i += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll 8
      for (int c = 0; c < AB_D; ++c) {
        a = fmaf(__bfloat162float(sQ[i * AB_LD + c]), __bfloat162float(sK[jl * AB_LD + c]), a);
        b = fmaf(__bfloat162float(sdO[i * AB_LD + c]), __bfloat162float(sV[jl * AB_LD + c]), b);
      }
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      const float pij = expf(x - lse[i]);
      s0[i] = pij;
      s1[i] = keep ? pij * (b - Dv[i]) * scale : 0.f;
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < T; ++i) {
      const float pij = s0[i], ds = s1[i];
      k0 = fmaf(ds, __bfloat162float(sQ[i * AB_LD + lane]), k0);
      k1 = fmaf(ds, __bfloat162float(sQ[i * AB_LD + lane + 32]), k1);
      v0 = fmaf(pij, __bfloat162float(sdO[i * AB_LD + lane]), v0);
      v1 = fmaf(pij, __bfloat162float(sdO[i * AB_LD + lane + 32]), v1);
    }
    TIO* dk = dqkv + base + (long long)j * 3 * H + H + head * AB_D;
    TIO* dv = dqkv + base + (long long)j * 3 * H + 2 * H + head * AB_D;
    st1<TIO>(dk + lane, k0); st1<TIO>(dk + lane + 32, k1);
    st1<TIO>(dv + lane, v0); st1<TIO>(dv + lane + 32, v1);
    __syncwarp();
  }
}

// fp32-exact variant: same algorithm with float staging (used by the fp32 verification mode).
__global__ void __launch_bounds__(256)
attention_bwd_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx,
                         const uint32_t* __restrict__ mask, int words, int T, int H, int heads, int mode,
                         float* __restrict__ dqkv,
                         unsigned drop_thr, unsigned long long drop_seed, float drop_scale) {
  extern __shared__ uint8_t smraw[];
  constexpr int LD = AB_D + 1;
  float* sQ = reinterpret_cast<float*>(smraw);
  float* sK = sQ + T * LD;
  float* sV = sK + T * LD;
  float* sdO = sV + T * LD;
  float* lse = sdO + T * LD;
  float* Dv = lse + T;
  float* strips = Dv + T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x / heads, head = blockIdx.x % heads;
  const long long base = (long long)row * T * 3 * H;
  const uint32_t* mrow = mask + (long long)row * words;
  const float scale = 0.125f;
  for (int e = threadIdx.x; e < T * AB_D; e += blockDim.x) {
    const int t = e / AB_D, c = e % AB_D;
    const float* p = qkv + base + (long long)t * 3 * H + head * AB_D + c;
    sQ[t * LD + c] = p[0];
    sK[t * LD + c] = p[H];
    sV[t * LD + c] = p[2 * H];
    sdO[t * LD + c] = dctx[((long long)row * T + t) * H + head * AB_D + c];
  }
  __syncthreads();
  float* s0 = strips + warp * 2 * T;
  float* s1 = s0 + T;
  for (int i = warp; i < T; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll 8
      for (int c = 0; c < AB_D; ++c) {
        a = fmaf(sQ[i * LD + c], sK[j * LD + c], a);
        b = fmaf(sdO[i * LD + c], sV[j * LD + c], b);
      }
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      if (drop_thr) b = agb_attn_keep(drop_seed, blockIdx.x, i, j, drop_thr) ? b * drop_scale : 0.f;   // dP o M'
      s0[j] = x; s1[j] = b;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float z = 0.f;
    for (int j = lane; j < T; j += 32) z += expf(s0[j] - mx);
    z = warp_sum(z);
    const float l = mx + logf(z);
    float dsum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float pj = expf(s0[j] - l);
      dsum += pj * s1[j];
      s0[j] = pj;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) { lse[i] = l; Dv[i] = dsum; }
    __syncwarp();
    for (int j = lane; j < T; j += 32) {
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      s0[j] = keep ? s0[j] * (s1[j] - dsum) * scale : 0.f;
    }
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float ds = s0[j];
      q0 = fmaf(ds, sK[j * LD + lane], q0);
      q1 = fmaf(ds, sK[j * LD + lane + 32], q1);
    }
    float* dq = dqkv + base + (long long)i * 3 * H + head * AB_D;
    dq[lane] = q0; dq[lane + 32] = q1;
    __syncwarp();
  }
  __syncthreads();
  for (int j = warp; j < T; j += nw) {
    const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
    for (int i = lane; i < T; i += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll 8
      for (int c = 0; c < AB_D; ++c) {
        a = fmaf(sQ[i * LD + c], sK[j * LD + c], a);
        b = fmaf(sdO[i * LD + c], sV[j * LD + c], b);
      }
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      const float pij = expf(x - lse[i]);
      float mk = 1.f;
      if (drop_thr) mk = agb_attn_keep(drop_seed, blockIdx.x, i, j, drop_thr) ? drop_scale : 0.f;
      s0[i] = pij * mk;
      s1[i] = keep ? pij * (b * mk - Dv[i]) * scale : 0.f;
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int i = 0; i < T; ++i) {
      const float pij = s0[i], ds = s1[i];
      k0 = fmaf(ds, sQ[i * LD + lane], k0);
      k1 = fmaf(ds, sQ[i * LD + lane + 32], k1);
      v0 = fmaf(pij, sdO[i * LD + lane], v0);
      v1 = fmaf(pij, sdO[i * LD + lane + 32], v1);
    }
    float* dk = dqkv + base + (long long)j * 3 * H + H + head * AB_D;
    float* dv = dqkv + base + (long long)j * 3 * H + 2 * H + head * AB_D;
    dk[lane] = k0; dk[lane + 32] = k1;
    dv[lane] = v0; dv[lane + 32] = v1;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// nn.Dropout in training mode (reference models/vanilla_vit.py:253,501-503,512-516, vanilla_bert.py:325,559,603):
//   out = residual + keep * y / (1 - p),   keep = hash bits of (seed, tag, element index) >= thr16
// The adjoint is the same map applied to the incoming gradient (no residual), so no mask is stored.
// ------------------------------------------------------------------------------------------------
template <typename TY, typename TO>
__global__ void dropout_kernel(const TY* __restrict__ y, const float* __restrict__ res, TO* __restrict__ out, long long n,
                               uint32_t key, uint32_t thr, float scale) {
  const long long pairs = (n + 1) / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t x = agb_drop_bits(key, (uint32_t)i);
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const long long e = 2 * i + o;
      if (e >= n) break;
      const bool keep = (o ? (x >> 16) : (x & 0xFFFFu)) >= thr;
      float v = keep ? ld1<TY>(y + e) * scale : 0.f;
      if (res) v += res[e];
      st1<TO>(out + e, v);
    }
  }
}

int dropout(const void* y, int y_bf16, const float* residual, void* out, int out_bf16, long long n, unsigned thr16,
            unsigned long long seed, unsigned tag, cudaStream_t st) {
  AGB_REQUIRE(n >= 0 && thr16 < 65536u, "dropout arguments");
  if (n == 0) return AGB_OK;
  AGB_REQUIRE(y && out, "null pointer");
  AGB_REQUIRE(n < (1ll << 32), "dropout: at most 2^32 elements per call");
  const uint32_t key = agb_drop_key(seed, tag, 0x5bd1e995u);
  const float scale = 65536.0f / (65536.0f - (float)thr16);
  const int blocks = (int)std::min<long long>(((n + 1) / 2 + 255) / 256, 8ll * sm_count());
#define DROP_GO(TY, TO) \
  dropout_kernel<TY, TO><<<blocks, 256, 0, st>>>(static_cast<const TY*>(y), residual, static_cast<TO*>(out), n, key, thr16, scale)
  if (y_bf16 && out_bf16) DROP_GO(bf16, bf16);
  else if (y_bf16) DROP_GO(bf16, float);
  else if (out_bf16) DROP_GO(float, bf16);
  else DROP_GO(float, float);
#undef DROP_GO
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// dense keep mask of the attention-probability dropout, for tests: keep[(row, head, query, key)] in {0, 1}
__global__ void attention_dropout_mask_kernel(unsigned char* __restrict__ keep, long long n, int heads, int T, uint32_t thr,
                                              unsigned long long seed) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % T), i = (int)((e / T) % T);
    const uint32_t unit = (uint32_t)(e / ((long long)T * T));
    keep[e] = agb_attn_keep(seed, unit, (uint32_t)i, (uint32_t)j, thr) ? 1 : 0;
  }
}

int attention_dropout_mask(unsigned char* keep, int rows, int heads, int T, unsigned thr16, unsigned long long seed,
                           cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && heads > 0 && T > 0 && thr16 < 65536u, "arguments");
  const long long n = (long long)rows * heads * T * T;
  if (n == 0) return AGB_OK;
  AGB_REQUIRE(keep, "null pointer");
  attention_dropout_mask_kernel<<<(int)std::min<long long>((n + 255) / 256, 8ll * sm_count()), 256, 0, st>>>(keep, n, heads, T,
                                                                                                           thr16, seed);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Narrow heads (head dim 8 / 16 / 32: the side ladders of the LTT variants, hidden size s_attn_hidden_size split over the
// backbone's head count, reference models/ltt_vit.py:386-396).  Same two passes and the same arithmetic as
// attention_bwd_kernel with the operands staged as fp32 (so fp32 I/O is the exact mode here too); the head dim no
// longer fills a warp, so 32 / D lane groups split the reduction index and are folded with shuffles.
// ------------------------------------------------------------------------------------------------
template <typename TIO, int D>
__global__ void __launch_bounds__(256)
attention_bwd_small_kernel(const TIO* __restrict__ qkv, const TIO* __restrict__ dctx, const uint32_t* __restrict__ mask,
                           int words, int T, int H, int heads, int mode, float scale, TIO* __restrict__ dqkv,
                           unsigned drop_thr, unsigned long long drop_seed, float drop_scale) {
  constexpr int LD = D + 1;
  constexpr int G = 32 / D;
  extern __shared__ uint8_t smraw[];
  float* sQ = reinterpret_cast<float*>(smraw);
  float* sK = sQ + T * LD;
  float* sV = sK + T * LD;
  float* sdO = sV + T * LD;
  float* lse = sdO + T * LD;   // T   (max + log-sum)
  float* Dv = lse + T;         // T
  float* strips = Dv + T;      // nw * 2 * T
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x / heads, head = blockIdx.x % heads;
  const long long base = (long long)row * T * 3 * H;
  const uint32_t* mrow = mask + (long long)row * words;
  for (int e = threadIdx.x; e < T * D; e += blockDim.x) {
    const int t = e / D, c = e % D;
    const TIO* p = qkv + base + (long long)t * 3 * H + head * D + c;
    sQ[t * LD + c] = ld1<TIO>(p);
    sK[t * LD + c] = ld1<TIO>(p + H);
    sV[t * LD + c] = ld1<TIO>(p + 2 * H);
    sdO[t * LD + c] = ld1<TIO>(dctx + ((long long)row * T + t) * H + head * D + c);
  }
  __syncthreads();
  float* s0 = strips + warp * 2 * T;
  float* s1 = s0 + T;
  const int c = lane % D, g = lane / D;
  // ---------------- pass A: query-owned (row statistics, dQ) ----------------
  for (int i = warp; i < T; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        a = fmaf(sQ[i * LD + k], sK[j * LD + k], a);
        b = fmaf(sdO[i * LD + k], sV[j * LD + k], b);
      }
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      if (drop_thr)      // dP carries the forward's dropout mask: O = (P o M / (1-p)) V
        b = agb_attn_keep(drop_seed, blockIdx.x, i, j, drop_thr) ? b * drop_scale : 0.f;
      s0[j] = x;
      s1[j] = b;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float z = 0.f;
    for (int j = lane; j < T; j += 32) z += expf(s0[j] - mx);
    z = warp_sum(z);
    const float l = mx + logf(z);
    float dsum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float pj = expf(s0[j] - l);
      dsum += pj * s1[j];
      s0[j] = pj;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) { lse[i] = l; Dv[i] = dsum; }
    __syncwarp();
    for (int j = lane; j < T; j += 32) {
      const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
      s0[j] = keep ? s0[j] * (s1[j] - dsum) * scale : 0.f;
    }
    __syncwarp();
    float q = 0.f;
    for (int j = g; j < T; j += G) q = fmaf(s0[j], sK[j * LD + c], q);
#pragma unroll
    for (int o = D; o < 32; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (g == 0) st1<TIO>(dqkv + base + (long long)i * 3 * H + head * D + c, q);
    __syncwarp();
  }
  __syncthreads();
  // ---------------- pass B: key-owned (dK, dV; scores recomputed) ----------------
  for (int j = warp; j < T; j += nw) {
    const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
    for (int i = lane; i < T; i += 32) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        a = fmaf(sQ[i * LD + k], sK[j * LD + k], a);
        b = fmaf(sdO[i * LD + k], sV[j * LD + k], b);
      }
      float x = a * scale;
      if (!keep) x = (mode == AGB_MASK_MUL0) ? 0.f : -INFINITY;
      const float pij = expf(x - lse[i]);
      float mk = 1.f;
      if (drop_thr) mk = agb_attn_keep(drop_seed, blockIdx.x, i, j, drop_thr) ? drop_scale : 0.f;
      s0[i] = pij * mk;                                          // for dV
      s1[i] = keep ? pij * (b * mk - Dv[i]) * scale : 0.f;       // dS for dK
    }
    __syncwarp();
    float kk = 0.f, vv = 0.f;
    for (int i = g; i < T; i += G) {
      kk = fmaf(s1[i], sQ[i * LD + c], kk);
      vv = fmaf(s0[i], sdO[i * LD + c], vv);
    }
#pragma unroll
    for (int o = D; o < 32; o <<= 1) {
      kk += __shfl_xor_sync(0xffffffffu, kk, o);
      vv += __shfl_xor_sync(0xffffffffu, vv, o);
    }
    if (g == 0) {
      st1<TIO>(dqkv + base + (long long)j * 3 * H + H + head * D + c, kk);
      st1<TIO>(dqkv + base + (long long)j * 3 * H + 2 * H + head * D + c, vv);
    }
    __syncwarp();
  }
}

template <typename TIO, int D>
static int launch_attention_bwd_small(const void* qkv, const void* dctx, const uint32_t* mask, int words, int rows, int T,
                                      int H, int heads, int mode, void* dqkv, cudaStream_t st, unsigned drop_thr,
                                      unsigned long long drop_seed) {
  const int nw = 8;
  const size_t smem = ((size_t)4 * T * (D + 1) + 2 * (size_t)T + (size_t)nw * 2 * T) * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("agb_masked_attention_bwd: T = %d with head dim %d does not fit in shared memory", T, D);
    return AGB_ERR_UNSUPPORTED;
  }
  AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_small_kernel<TIO, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_bwd_small_kernel<TIO, D><<<rows * heads, nw * 32, smem, st>>>(
      static_cast<const TIO*>(qkv), static_cast<const TIO*>(dctx), mask, words, T, H, heads, mode, rsqrtf((float)D),
      static_cast<TIO*>(dqkv), drop_thr, drop_seed, 65536.0f / (65536.0f - (float)drop_thr));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int attention_bwd_tc(const bf16* qkv, const bf16* dctx, const uint32_t* mask, int words, int rows, int T, int H,
                     int heads, int mode, bf16* dqkv, cudaStream_t stream, unsigned drop_thr, unsigned long long drop_seed);
static int g_attention_bwd_variant = 0;   // 0 auto (tensor cores for bf16), 1 CUDA-core kernel only
void set_attention_bwd_variant(int v) { g_attention_bwd_variant = v; }
int get_attention_bwd_variant() { return g_attention_bwd_variant; }

int attention_bwd(const void* qkv, const void* dctx, int io_bf16, const uint32_t* mask, int words, int rows, int T,
                  int H, int heads, int mode, void* dqkv, cudaStream_t st, unsigned drop_thr, unsigned long long drop_seed) {
  AGB_REQUIRE(rows >= 0 && T > 0 && heads > 0 && H % heads == 0, "attention shape");
  AGB_REQUIRE(words * 32 >= T, "mask words");
  AGB_REQUIRE(mode == AGB_MASK_MUL0 || mode == AGB_MASK_NEGINF, "mask mode");
  const int hd = H / heads;
  if (hd != AB_D) {
    AGB_REQUIRE(hd == 8 || hd == 16 || hd == 32, "attention backward: head dim must be 64, 32, 16 or 8");
    if (rows == 0) return AGB_OK;
    AGB_REQUIRE(qkv && dctx && mask && dqkv, "null pointer");
#define ABS_LAUNCH(D)                                                                                                  \
  (io_bf16 ? launch_attention_bwd_small<bf16, D>(qkv, dctx, mask, words, rows, T, H, heads, mode, dqkv, st, drop_thr, drop_seed) \
           : launch_attention_bwd_small<float, D>(qkv, dctx, mask, words, rows, T, H, heads, mode, dqkv, st, drop_thr, drop_seed))
    return hd == 8 ? ABS_LAUNCH(8) : (hd == 16 ? ABS_LAUNCH(16) : ABS_LAUNCH(32));
#undef ABS_LAUNCH
  }
  if (T > 512) {
    set_last_error("agb_masked_attention_bwd supports T <= 512 (got %d)", T);
    return AGB_ERR_UNSUPPORTED;
  }
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(qkv && dctx && mask && dqkv, "null pointer");
  if (io_bf16 && (g_attention_bwd_variant == 0 || drop_thr > 0)) {
    const int rc2 = attention_bwd_tc(static_cast<const bf16*>(qkv), static_cast<const bf16*>(dctx), mask, words, rows, T, H,
                                     heads, mode, static_cast<bf16*>(dqkv), st, drop_thr, drop_seed);
    if (rc2 != AGB_ERR_UNSUPPORTED) return rc2;
  }
  if (drop_thr > 0 && (io_bf16 || T > 256)) {
    set_last_error("attention dropout adjoint: T <= 256 (tcgen05 kernel in bf16, CUDA-core kernel in fp32 up to T = 200), "
                   "or head dims 8/16/32 (T = %d)", T);
    return AGB_ERR_UNSUPPORTED;
  }
  if (T > 256) {
    // two-kernel form (operands staged in bf16, fp32 math): fp32 I/O is accepted but is not the exact mode here
    AGB_REQUIRE(rows * heads <= 65535, "grid limits (chunk the rows)");
    // row statistics (L, D) between the two kernels: stream-ordered scratch, released after the second kernel
    float* stat = nullptr;
    AGB_CHECK_CUDA(cudaMallocAsync(&stat, (size_t)rows * heads * 2 * T * sizeof(float), st));
    const int nwl = 8;
    const size_t smem_l = (size_t)2 * T * AB_LD * 2 + (size_t)2 * ABL_BLK * AB_LD * 2 + (size_t)2 * T * 4 +
                          (size_t)nwl * 2 * T * 4;
    dim3 grid((T + ABL_BLK - 1) / ABL_BLK, rows * heads);
#define ABL_LAUNCH(TIO)                                                                                               \
  do {                                                                                                                \
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_long_a_kernel<TIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l)); \
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_long_b_kernel<TIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l)); \
    attention_bwd_long_a_kernel<TIO><<<grid, nwl * 32, smem_l, st>>>(static_cast<const TIO*>(qkv), static_cast<const TIO*>(dctx), \
                                                                     mask, words, T, H, heads, mode, static_cast<TIO*>(dqkv), stat); \
    attention_bwd_long_b_kernel<TIO><<<grid, nwl * 32, smem_l, st>>>(static_cast<const TIO*>(qkv), static_cast<const TIO*>(dctx), \
                                                                     mask, words, T, H, heads, mode, static_cast<TIO*>(dqkv), stat); \
  } while (0)
    if (io_bf16) ABL_LAUNCH(bf16);
    else ABL_LAUNCH(float);
#undef ABL_LAUNCH
    const cudaError_t launch_err = cudaGetLastError();
    AGB_CHECK_CUDA(cudaFreeAsync(stat, st));
    AGB_CHECK_CUDA(launch_err);
    return AGB_OK;
  }
  const int nw = 8;
  if (io_bf16) {
    const size_t smem = (size_t)4 * T * AB_LD * 2 + (size_t)2 * T * 4 + (size_t)nw * 2 * T * 4;
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_bwd_kernel<bf16><<<rows * heads, nw * 32, smem, st>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(dctx),
                                                                    mask, words, T, H, heads, mode, static_cast<bf16*>(dqkv));
  } else {
    const size_t smem = (size_t)4 * T * (AB_D + 1) * 4 + (size_t)2 * T * 4 + (size_t)nw * 2 * T * 4;
    AGB_REQUIRE(smem <= 227 * 1024, "T too large for the fp32 attention backward (<= 200)");
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_bwd_f32_kernel<<<rows * heads, nw * 32, smem, st>>>(static_cast<const float*>(qkv), static_cast<const float*>(dctx),
                                                                  mask, words, T, H, heads, mode, static_cast<float*>(dqkv),
                                                                  drop_thr, drop_seed, 65536.0f / (65536.0f - (float)drop_thr));
  }
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// embedding adjoints
// ------------------------------------------------------------------------------------------------
// ViT (reference models/vanilla_vit.py:242-253): dpos[t] += sum_b dx[b,t]; dcls += sum_b dx[b,0];
// dpatch[b,p] = dx[b,1+p] (fp32 or bf16, feeds the patch-projection wgrad GEMM).
template <typename TOut>
__global__ void vit_embed_bwd_kernel(const float* __restrict__ dx, int B, int T, int H, float* __restrict__ dpos,
                                     float* __restrict__ dcls, TOut* __restrict__ dpatch) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)T * H) return;
  const int t = (int)(gid / H), c = (int)(gid % H);
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float v = dx[((long long)b * T + t) * H + c];
    s += v;
    if (t > 0) st1<TOut>(dpatch + ((long long)b * (T - 1) + (t - 1)) * H + c, v);
  }
  dpos[gid] += s;
  if (t == 0) dcls[c] += s;
}
int vit_embed_bwd(const float* dx, int B, int T, int H, float* dpos, float* dcls, void* dpatch, int dpatch_bf16,
                  cudaStream_t st) {
  AGB_REQUIRE(B > 0 && T > 1 && H > 0, "shape");
  AGB_REQUIRE(dx && dpos && dcls && dpatch, "null pointer");
  const long long n = (long long)T * H;
  const int blocks = (int)((n + 255) / 256);
  if (dpatch_bf16) vit_embed_bwd_kernel<bf16><<<blocks, 256, 0, st>>>(dx, B, T, H, dpos, dcls, static_cast<bf16*>(dpatch));
  else vit_embed_bwd_kernel<float><<<blocks, 256, 0, st>>>(dx, B, T, H, dpos, dcls, static_cast<float*>(dpatch));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// BERT (reference models/vanilla_bert.py:307-325): the embedding LayerNorm adjoint is layernorm_bwd on the
// recomputed pre-norm sum; this kernel rebuilds that sum (fwd) and scatters its gradient (bwd).
__global__ void bert_embed_sum_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                      const float* __restrict__ pos, const float* __restrict__ type0, int BT, int T,
                                      int H, int vocab, float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)BT * H) return;
  const int tok = (int)(gid / H), c = (int)(gid % H);
  long long id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  out[gid] = (word[id * H + c] + type0[c]) + pos[(long long)(tok % T) * H + c];
}
__global__ void bert_embed_scatter_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dsum, int BT,
                                          int T, int H, int vocab, int pad_id, float* __restrict__ dword,
                                          float* __restrict__ dpos, float* __restrict__ dtype0) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)BT * H) return;
  const int tok = (int)(gid / H), c = (int)(gid % H);
  long long id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float g = dsum[gid];
  if (id != pad_id) atomicAdd(dword + id * H + c, g);   // nn.Embedding(padding_idx): no gradient for [PAD]
  atomicAdd(dpos + (long long)(tok % T) * H + c, g);
  atomicAdd(dtype0 + c, g);
}
int bert_embed_sum(const int64_t* ids, const float* word, const float* pos, const float* type0, int BT, int T, int H,
                   int vocab, float* out, cudaStream_t st) {
  AGB_REQUIRE(BT > 0 && T > 0 && H > 0, "shape");
  AGB_REQUIRE(ids && word && pos && type0 && out, "null pointer");
  const long long n = (long long)BT * H;
  bert_embed_sum_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(ids, word, pos, type0, BT, T, H, vocab, out);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}
int bert_embed_scatter(const int64_t* ids, const float* dsum, int BT, int T, int H, int vocab, int pad_id,
                       float* dword, float* dpos, float* dtype0, cudaStream_t st) {
  AGB_REQUIRE(BT > 0 && T > 0 && H > 0, "shape");
  AGB_REQUIRE(ids && dsum && dword && dpos && dtype0, "null pointer");
  const long long n = (long long)BT * H;
  bert_embed_scatter_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(ids, dsum, BT, T, H, vocab, pad_id, dword, dpos, dtype0);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

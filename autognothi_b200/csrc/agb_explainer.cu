// Explainer-side kernels (all fp32 math — these produce the attributions and the loss, where bf16
// rounding would dominate the error budget, SURVEY.md §7 "Softmax-probability outputs"):
//   * explainer_head_fwd: last Linear E->C of `explainer_mlp` FUSED with the additive efficiency
//     normalisation over all T tokens (warp-shuffle reductions), CLS drop and (B,T,C)->(B,C,n)
//     transpose (reference models/vanilla_vit.py:123-129, models/shapley.py:82-93);
//   * explainer_head_bwd: adjoint of the above (dh, dW, db);
//   * shapley_loss_fwd/bwd: packed-mask gather/dot + squared error (reference models/shapley.py:9-53).
// HBM-bound: one CTA per input keeps that input's (T x C) / (C x n) slab in shared memory, every
// global access is a coalesced vector access, nothing is re-read.
#include "agb_common.cuh"

namespace agb {

constexpr int MAXC = 16;

template <typename TH>
__device__ __forceinline__ float4 ld4(const TH* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 ld4<bf16>(const bf16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}

// ------------------------------------------------------------------------------------------------
// forward: pred[t,c] = h[b,t,:] . W[c,:] + bias[c];  phi[b,c,t-1] = pred[t,c] + ((grand-null) - sum_t pred)/T
// ------------------------------------------------------------------------------------------------
template <typename TH>
__global__ void __launch_bounds__(256)
explainer_head_fwd_kernel(const TH* __restrict__ h, int T, int E, int C, const float* __restrict__ W,
                          const float* __restrict__ bias, const float* __restrict__ grand,
                          const float* __restrict__ null_v, int normalize, float* __restrict__ phi,
                          float* __restrict__ pred_out) {
  extern __shared__ float sm[];
  float* pred = sm;               // T*C
  float* diff = sm + T * C;       // C
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const TH* hb = h + (long long)b * T * E;
  for (int t = warp; t < T; t += nw) {
    float acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
    const TH* hr = hb + (long long)t * E;
    for (int e = lane * 4; e < E; e += 128) {
      const float4 x = ld4<TH>(hr + e);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)c * E + e));
          acc[c] = fmaf(x.x, w.x, acc[c]);
          acc[c] = fmaf(x.y, w.y, acc[c]);
          acc[c] = fmaf(x.z, w.z, acc[c]);
          acc[c] = fmaf(x.w, w.w, acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const float v = warp_sum(acc[c]);
        if (lane == 0) pred[t * C + c] = v + bias[c];
      }
    }
  }
  __syncthreads();
  // token-sum per class: one warp per class, shuffle reduction
  for (int c = warp; c < C; c += nw) {
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += pred[t * C + c];
    s = warp_sum(s);
    if (lane == 0)
      diff[c] = normalize ? ((grand[(long long)b * C + c] - null_v[c]) - s) / (float)T : 0.f;
  }
  __syncthreads();
  const int n = T - 1;
  for (int i = threadIdx.x; i < C * n; i += blockDim.x) {
    const int c = i / n, j = i % n;
    phi[((long long)b * C + c) * n + j] = pred[(j + 1) * C + c] + diff[c];
  }
  if (pred_out != nullptr)
    for (int i = threadIdx.x; i < T * C; i += blockDim.x) pred_out[(long long)b * T * C + i] = pred[i];
}

// ------------------------------------------------------------------------------------------------
// forward, clustered: one thread-block CLUSTER per input, CTA `rank` owns TPW * 8 consecutive tokens.  The head weights
// (C x E fp32) sit in shared memory; each warp keeps TPW tokens x CMAX classes of accumulators in registers, so one
// shared-memory read of W feeds TPW FMAs and h is streamed from HBM exactly once (8 B per lane, coalesced).  The token
// sum of the efficiency normalisation crosses the CTAs of an input through distributed shared memory
// (barrier.cluster + ld.shared::cluster): no second kernel, no atomics, deterministic.
// (The one-CTA-per-input kernel above re-read W from L2 for every token and ran on B SMs only: 680 us at the bench
// shape against 6 us of HBM time; profiles/r01_small_kernels_ncu.txt.)
// ------------------------------------------------------------------------------------------------
template <typename TH, int CMAX, int TPW>
__global__ void __launch_bounds__(256)
explainer_head_fwd_cluster_kernel(const TH* __restrict__ h, int T, int E, int C, const float* __restrict__ W,
                                  const float* __restrict__ bias, const float* __restrict__ grand,
                                  const float* __restrict__ null_v, int normalize, float* __restrict__ phi,
                                  float* __restrict__ pred_out) {
  constexpr int TOK = 8 * TPW;                 // tokens per CTA
  extern __shared__ __align__(16) float sm[];
  float* part = sm;                            // [MAXC] this CTA's token sums (same offset in every CTA of the cluster)
  float* diff = sm + MAXC;                     // [MAXC]
  float* pred = sm + 2 * MAXC;                 // [TOK][C]
  float* sW = pred + TOK * MAXC;               // [C][E]
  const int b = blockIdx.y;
  const uint32_t rank = cluster_ctarank(), csize = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 4; i < C * E; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(sW + i) = __ldg(reinterpret_cast<const float4*>(W + i));
  __syncthreads();
  const int t0 = (int)rank * TOK + warp * TPW;
  float acc[TPW][CMAX];
#pragma unroll
  for (int k = 0; k < TPW; ++k)
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[k][c] = 0.f;
  const TH* hb = h + (long long)b * T * E;
  for (int e = lane * 4; e < E; e += 128) {
    float4 w[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      w[c] = (c < C) ? *reinterpret_cast<const float4*>(sW + (long long)c * E + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    // all TPW rows are requested before the first FMA (tokens past the end re-read the last row; their results are
    // dropped below), so the HBM latency is paid once per step instead of once per token
    float4 x[TPW];
#pragma unroll
    for (int k = 0; k < TPW; ++k) x[k] = ld4<TH>(hb + (long long)min(t0 + k, T - 1) * E + e);
#pragma unroll
    for (int k = 0; k < TPW; ++k) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        acc[k][c] = fmaf(x[k].x, w[c].x, acc[k][c]);
        acc[k][c] = fmaf(x[k].y, w[c].y, acc[k][c]);
        acc[k][c] = fmaf(x[k].z, w[c].z, acc[k][c]);
        acc[k][c] = fmaf(x[k].w, w[c].w, acc[k][c]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < TPW; ++k)
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      const float v = warp_sum(acc[k][c]);
      if (lane == 0 && c < C) pred[(warp * TPW + k) * C + c] = (t0 + k < T) ? v + bias[c] : 0.f;
    }
  __syncthreads();
  for (int c = warp; c < C; c += 8) {          // this CTA's token sums: one warp per class
    float s = 0.f;
    for (int tt = lane; tt < TOK; tt += 32) s += pred[tt * C + c];
    s = warp_sum(s);
    if (lane == 0) part[c] = s;
  }
  cluster_sync_all();                          // every CTA's `part` is written and visible
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float tot = 0.f;
    for (uint32_t r = 0; r < csize; ++r) tot += ld_shared_cluster_f32(mapa_rank(smem_u32(part + c), r));
    diff[c] = normalize ? ((grand[(long long)b * C + c] - null_v[c]) - tot) / (float)T : 0.f;
  }
  cluster_sync_all();                          // nobody leaves (or reuses smem) while a peer may still read `part`
  const int n = T - 1;
  const int tb = (int)rank * TOK;
  for (int i = threadIdx.x; i < TOK * C; i += blockDim.x) {
    const int c = i / TOK, tt = i % TOK;       // consecutive threads = consecutive tokens: coalesced phi rows
    const int t = tb + tt;
    if (t >= 1 && t < T) phi[((long long)b * C + c) * n + (t - 1)] = pred[tt * C + c] + diff[c];
  }
  if (pred_out != nullptr)
    for (int i = threadIdx.x; i < TOK * C; i += blockDim.x)
      if (tb + i / C < T) pred_out[((long long)b * T + tb) * C + i] = pred[i];
}

template <typename TH, int CMAX, int TPW>
static int launch_head_fwd_cluster(const void* h, int B, int T, int E, int C, const float* W, const float* bias,
                                   const float* grand, const float* null_v, int normalize, float* phi, float* pred_out,
                                   cudaStream_t st) {
  constexpr int TOK = 8 * TPW;
  const int csize = (T + TOK - 1) / TOK;
  const size_t smem = ((size_t)2 * MAXC + (size_t)TOK * MAXC + (size_t)C * E) * sizeof(float);
  if (csize > 8 || smem > 227 * 1024 || B > 65535) return AGB_ERR_UNSUPPORTED;
  auto kern = explainer_head_fwd_cluster_kernel<TH, CMAX, TPW>;
  AGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize, B, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  AGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<const TH*>(h), T, E, C, W, bias, grand, null_v, normalize, phi,
                                    pred_out));
  return AGB_OK;
}

template <typename TH>
static int head_fwd_cluster(const void* h, int B, int T, int E, int C, const float* W, const float* bias, const float* grand,
                            const float* null_v, int normalize, float* phi, float* pred_out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(h) & 15)) return AGB_ERR_UNSUPPORTED;
#define HF(CM, TP) launch_head_fwd_cluster<TH, CM, TP>(h, B, T, E, C, W, bias, grand, null_v, normalize, phi, pred_out, st)
  if (C <= 2) return HF(2, 8);
  if (C <= 4) return HF(4, 8);
  if (C <= 10) return HF(10, 8);
  return HF(16, 4);
#undef HF
}

int explainer_head_fwd(const void* h, int h_bf16, int B, int T, int E, int C, const float* W,
                       const float* bias, const float* grand, const float* null_v, int normalize,
                       float* phi, float* pred_out, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && T > 1 && E % 4 == 0 && C > 0 && C <= MAXC, "explainer head shape (C <= 16)");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(h && W && bias && phi, "null pointer");
  AGB_REQUIRE(!normalize || (grand && null_v), "normalisation needs grand and null");
  {
    const int rc = h_bf16 ? head_fwd_cluster<bf16>(h, B, T, E, C, W, bias, grand, null_v, normalize, phi, pred_out, st)
                          : head_fwd_cluster<float>(h, B, T, E, C, W, bias, grand, null_v, normalize, phi, pred_out, st);
    if (rc != AGB_ERR_UNSUPPORTED) return rc;
  }
  const size_t smem = ((size_t)T * C + C) * sizeof(float);
  AGB_REQUIRE(smem <= 200 * 1024, "T*C too large");
  if (h_bf16) {
    if (smem > 48 * 1024)
      AGB_CHECK_CUDA(cudaFuncSetAttribute(explainer_head_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    explainer_head_fwd_kernel<bf16><<<B, 256, smem, st>>>(static_cast<const bf16*>(h), T, E, C, W, bias, grand,
                                                          null_v, normalize, phi, pred_out);
  } else {
    if (smem > 48 * 1024)
      AGB_CHECK_CUDA(cudaFuncSetAttribute(explainer_head_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    explainer_head_fwd_kernel<float><<<B, 256, smem, st>>>(static_cast<const float*>(h), T, E, C, W, bias, grand,
                                                           null_v, normalize, phi, pred_out);
  }
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// backward: dpred[t,c] = [t>=1] dphi[b,c,t-1] - (normalize ? sum_j dphi[b,c,j] / T : 0)
//   dh[b,t,:] = sum_c dpred[t,c] W[c,:]      dW[c,:] += sum_{b,t} dpred[t,c] h[b,t,:]     db[c] += sum dpred
// dW/db are accumulated with fp32 atomics across the B CTAs (each CTA first reduces over its tokens).
// ------------------------------------------------------------------------------------------------
template <typename TH>
__global__ void __launch_bounds__(256)
explainer_head_bwd_kernel(const float* __restrict__ dphi, const TH* __restrict__ h, int T, int E, int C,
                          const float* __restrict__ W, int normalize, TH* __restrict__ dh,
                          float* __restrict__ dW, float* __restrict__ db) {
  extern __shared__ float sm[];
  float* dpred = sm;            // T*C
  float* mean = sm + T * C;     // C
  const int b = blockIdx.x;
  const int n = T - 1;
  // gridDim.y token chunks per input (more CTAs than inputs: the training batch is 32 inputs on 148 SMs); every chunk
  // rebuilds the input's small dpred slab, handles its own tokens and adds its dW / db share with atomics
  const int chunk = (T + gridDim.y - 1) / gridDim.y;
  const int tc0 = blockIdx.y * chunk, tc1 = min(T, tc0 + chunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int c = warp; c < C; c += nw) {
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += dphi[((long long)b * C + c) * n + j];
    s = warp_sum(s);
    if (lane == 0) mean[c] = normalize ? s / (float)T : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    const int t = i / C, c = i % C;
    const float g = (t >= 1) ? dphi[((long long)b * C + c) * n + (t - 1)] : 0.f;
    dpred[i] = g - mean[c];
  }
  __syncthreads();
  // db
  if (db != nullptr) {
    for (int c = warp; c < C; c += nw) {
      float s = 0.f;
      for (int t = tc0 + lane; t < tc1; t += 32) s += dpred[t * C + c];
      s = warp_sum(s);
      if (lane == 0) atomicAdd(db + c, s);
    }
  }
  // dh (coalesced over e) and dW (each thread owns 4 consecutive e for all c; loops over tokens)
  const TH* hb = h + (long long)b * T * E;
  for (int e = threadIdx.x * 4; e < E; e += blockDim.x * 4) {
    float4 wreg[MAXC];
    float4 dwacc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      wreg[c] = (c < C) ? __ldg(reinterpret_cast<const float4*>(W + (long long)c * E + e)) : make_float4(0, 0, 0, 0);
      dwacc[c] = make_float4(0, 0, 0, 0);
    }
    for (int tg = tc0; tg < tc1; tg += 4) {
      // four rows requested before the first FMA: the HBM latency is paid once per group, not once per token
      float4 xs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xs[k] = ld4<TH>(hb + (long long)min(tg + k, tc1 - 1) * E + e);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = tg + k;
        if (t >= tc1) break;
        const float4 x = xs[k];
        float4 o = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          if (c < C) {
            const float g = dpred[t * C + c];
            o.x = fmaf(g, wreg[c].x, o.x); o.y = fmaf(g, wreg[c].y, o.y);
            o.z = fmaf(g, wreg[c].z, o.z); o.w = fmaf(g, wreg[c].w, o.w);
            dwacc[c].x = fmaf(g, x.x, dwacc[c].x); dwacc[c].y = fmaf(g, x.y, dwacc[c].y);
            dwacc[c].z = fmaf(g, x.z, dwacc[c].z); dwacc[c].w = fmaf(g, x.w, dwacc[c].w);
          }
        }
        if (dh != nullptr) {
          TH* dst = dh + ((long long)b * T + t) * E + e;
          if (sizeof(TH) == 4) {
            *reinterpret_cast<float4*>(dst) = o;
          } else {
            uint2 pk;
            pk.x = pack_bf16x2(o.x, o.y);
            pk.y = pack_bf16x2(o.z, o.w);
            *reinterpret_cast<uint2*>(dst) = pk;
          }
        }
      }
    }
    if (dW != nullptr) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          float* d = dW + (long long)c * E + e;
          atomicAdd(d, dwacc[c].x); atomicAdd(d + 1, dwacc[c].y);
          atomicAdd(d + 2, dwacc[c].z); atomicAdd(d + 3, dwacc[c].w);
        }
      }
    }
  }
}

int explainer_head_bwd(const float* dphi, const void* h, int h_bf16, int B, int T, int E, int C,
                       const float* W, int normalize, void* dh, float* dW, float* db, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && T > 1 && E % 4 == 0 && C > 0 && C <= MAXC, "explainer head shape (C <= 16)");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(dphi && h && W, "null pointer");
  const size_t smem = ((size_t)T * C + C) * sizeof(float);
  AGB_REQUIRE(smem <= 200 * 1024, "T*C too large");
  AGB_REQUIRE(B <= 65535, "batch too large (chunk it)");
  // token chunks so that about two waves of CTAs cover the SMs even for a small batch (at least ~16 tokens per chunk)
  int chunks = (2 * sm_count() + B - 1) / B;
  chunks = max(1, min(chunks, (T + 15) / 16));
  const dim3 grid(B, chunks);
  if (h_bf16) {
    if (smem > 48 * 1024)
      AGB_CHECK_CUDA(cudaFuncSetAttribute(explainer_head_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    explainer_head_bwd_kernel<bf16><<<grid, 256, smem, st>>>(dphi, static_cast<const bf16*>(h), T, E, C, W, normalize,
                                                             static_cast<bf16*>(dh), dW, db);
  } else {
    if (smem > 48 * 1024)
      AGB_CHECK_CUDA(cudaFuncSetAttribute(explainer_head_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    explainer_head_bwd_kernel<float><<<grid, 256, smem, st>>>(dphi, static_cast<const float*>(h), T, E, C, W, normalize,
                                                              static_cast<float*>(dh), dW, db);
  }
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// standalone efficiency normalisation (reference models/shapley.py:82-93) for callers that already
// hold pred (B,T,C): same arithmetic as the fused tail above.
// ------------------------------------------------------------------------------------------------
__global__ void normalize_kernel(const float* __restrict__ pred, const float* __restrict__ grand,
                                 const float* __restrict__ null_v, int T, int C, float* __restrict__ out) {
  __shared__ float diff[64];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* pb = pred + (long long)b * T * C;
  for (int c = warp; c < C; c += nw) {
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += pb[t * C + c];
    s = warp_sum(s);
    if (lane == 0) diff[c] = ((grand[(long long)b * C + c] - null_v[c]) - s) / (float)T;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) out[(long long)b * T * C + i] = pb[i] + diff[i % C];
}

int normalize_shapley(const float* pred, const float* grand, const float* null_v, int B, int T, int C,
                      float* out, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && T > 0 && C > 0 && C <= 64, "normalise shape (C <= 64)");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(pred && grand && null_v && out, "null pointer");
  normalize_kernel<<<B, 256, 0, st>>>(pred, grand, null_v, T, C, out);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Shapley loss (reference models/shapley.py:9-53)
//   approx[b,s,c] = v0[c] + sum_j mask[b,s,j] phi[b,c,j];  loss = n * mean((approx - v_s)^2)
// masks are the packed words (bit j+1 = player j).  CTA per input: phi[b] staged in smem once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shapley_loss_fwd_kernel(const uint32_t* __restrict__ packed, int words, const float* __restrict__ v0,
                        const float* __restrict__ v_s, const float* __restrict__ phi, int S, int n, int C,
                        float* __restrict__ resid, float* __restrict__ partial) {
  extern __shared__ float sm[];
  float* ph = sm;  // C*n
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < C * n; i += blockDim.x) ph[i] = phi[(long long)b * C * n + i];
  __syncthreads();
  float sq = 0.f;
  for (int s = warp; s < S; s += nw) {
    const uint32_t* mrow = packed + ((long long)b * S + s) * words;
    float acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
    for (int j = lane; j < n; j += 32) {
      const int tok = j + 1;
      const uint32_t keep = (mrow[tok >> 5] >> (tok & 31)) & 1u;
      if (keep) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) acc[c] += ph[c * n + j];
      }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const float a = warp_sum(acc[c]);
        if (lane == 0) {
          const long long idx = ((long long)b * S + s) * C + c;
          const float r = (v0[c] + a) - v_s[idx];
          resid[idx] = r;
          sq += r * r;
        }
      }
    }
  }
  if (lane == 0) red[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += red[w];
    partial[b] = t;
  }
}

// fixed-order final reduction -> deterministic loss
__global__ void shapley_loss_final_kernel(const float* __restrict__ partial, int B, float scale,
                                          float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    loss[0] = t * scale;
  }
}

int shapley_loss_fwd(const uint32_t* packed, int words, const float* v0, const float* v_s, const float* phi,
                     int B, int S, int n, int C, float* resid, float* partial, float* loss, cudaStream_t st) {
  AGB_REQUIRE(B > 0 && S > 0 && n > 0 && C > 0 && C <= MAXC, "loss shape (C <= 16)");
  AGB_REQUIRE(words * 32 >= n + 1, "mask words");
  AGB_REQUIRE(packed && v0 && v_s && phi && resid && partial && loss, "null pointer");
  const size_t smem = (size_t)C * n * sizeof(float);
  AGB_REQUIRE(smem <= 200 * 1024, "C*n too large");
  if (smem > 48 * 1024)
    AGB_CHECK_CUDA(cudaFuncSetAttribute(shapley_loss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  shapley_loss_fwd_kernel<<<B, 256, smem, st>>>(packed, words, v0, v_s, phi, S, n, C, resid, partial);
  AGB_CHECK_CUDA(cudaGetLastError());
  const float scale = (float)n / ((float)B * (float)S * (float)C);
  shapley_loss_final_kernel<<<1, 256, 0, st>>>(partial, B, scale, loss);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// dphi[b,c,j] = gscale * (2n / (B S C)) * sum_s mask[b,s,j] resid[b,s,c]
__global__ void __launch_bounds__(256)
shapley_loss_bwd_kernel(const uint32_t* __restrict__ packed, int words, const float* __restrict__ resid,
                        const float* __restrict__ gout, int S, int n, int C, float coef,
                        float* __restrict__ dphi) {
  extern __shared__ float sm[];
  float* rs = sm;                                            // S*C
  uint32_t* mk = reinterpret_cast<uint32_t*>(sm + S * C);    // S*words
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) rs[i] = resid[(long long)b * S * C + i];
  for (int i = threadIdx.x; i < S * words; i += blockDim.x) mk[i] = packed[(long long)b * S * words + i];
  __syncthreads();
  const float g = coef * (gout ? gout[0] : 1.0f);
  for (int i = threadIdx.x; i < C * n; i += blockDim.x) {
    const int c = i / n, j = i % n;
    const int tok = j + 1;
    float a = 0.f;
    for (int s = 0; s < S; ++s) {
      const uint32_t keep = (mk[s * words + (tok >> 5)] >> (tok & 31)) & 1u;
      a += keep ? rs[s * C + c] : 0.f;
    }
    dphi[(long long)b * C * n + i] = a * g;
  }
}

int shapley_loss_bwd(const uint32_t* packed, int words, const float* resid, const float* gout, int B, int S,
                     int n, int C, float* dphi, cudaStream_t st) {
  AGB_REQUIRE(B > 0 && S > 0 && n > 0 && C > 0 && C <= MAXC, "loss shape (C <= 16)");
  AGB_REQUIRE(packed && resid && dphi, "null pointer");
  const size_t smem = ((size_t)S * C + (size_t)S * words) * sizeof(float);
  AGB_REQUIRE(smem <= 200 * 1024, "S too large");
  if (smem > 48 * 1024)
    AGB_CHECK_CUDA(cudaFuncSetAttribute(shapley_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float coef = 2.0f * (float)n / ((float)B * (float)S * (float)C);
  shapley_loss_bwd_kernel<<<B, 256, smem, st>>>(packed, words, resid, gout, S, n, C, coef, dphi);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

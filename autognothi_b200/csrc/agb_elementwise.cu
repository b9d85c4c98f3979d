// HBM-bound glue kernels of the surrogate / explainer forward: LayerNorm, ViT patchify + embedding
// assembly, BERT embeddings, CLS classification heads.  All are vectorised row kernels (one warp per
// token row, 16-byte lane accesses) — no data reuse, so no shared-memory tiling; the aim is one
// read and one write of each activation at full HBM sector efficiency.
#include "agb_common.cuh"

namespace agb {

// ------------------------------------------------------------------------------------------------
// LayerNorm over the hidden dimension (reference nn.LayerNorm at models/vanilla_vit.py:213,369,373,
// 94; models/vanilla_bert.py:318,548,596).  Two-pass mean/variance in fp32 (matches torch's
// numerics to ~1 ulp), input fp32 or bf16, outputs bf16 and/or fp32.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__device__ __forceinline__ float4 load4(const TIn* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<bf16>(const bf16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}

template <typename TIn>
__global__ void layernorm_kernel(const TIn* __restrict__ x, long long in_stride, int rows, int H,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, bf16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                 long long out_stride) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TIn* xr = x + (long long)row * in_stride;
  float s = 0.f;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 v = load4<TIn>(xr + i);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)H;
  float ss = 0.f;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 v = load4<TIn>(xr + i);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)H + eps);
  for (int i = lane * 4; i < H; i += 128) {
    const float4 v = load4<TIn>(xr + i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + i));
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + b.x;
    o.y = (v.y - mean) * rstd * g.y + b.y;
    o.z = (v.z - mean) * rstd * g.z + b.z;
    o.w = (v.w - mean) * rstd * g.w + b.w;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)row * out_stride + i) = o;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(out_bf16 + (long long)row * out_stride + i) = pk;
    }
  }
}

// Register-resident variant for the common widths (H = 128*NV4 floats per warp pass, H <= 1024): the row is
// read from HBM exactly once (NV4 independent 16-byte loads per lane, all issued before the first use), the
// two-pass mean / variance runs on registers, and the normalised row is written once.
template <typename TIn, int NV4>
__global__ void __launch_bounds__(256)
layernorm_reg_kernel(const TIn* __restrict__ x, long long in_stride, int rows, int H,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     bf16* __restrict__ out_bf16, float* __restrict__ out_f32, long long out_stride) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TIn* xr = x + (long long)row * in_stride;
  float4 v[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) v[i] = load4<TIn>(xr + lane * 4 + i * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)H;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int col = lane * 4 + i * 128;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)row * out_stride + col) = o;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(out_bf16 + (long long)row * out_stride + col) = pk;
    }
  }
}

// Narrow rows (H = 32 * NV4 <= 128: the LTT side ladders, hidden/8 = 96 for ViT-Base and BERT-base): 8 lanes per row, 4 rows
// per warp, the row held in registers (NV4 independent 16-byte loads per lane), 3-step shuffles inside the 8-lane group.
template <typename TIn, int NV4>
__global__ void __launch_bounds__(256)
layernorm_narrow_kernel(const TIn* __restrict__ x, long long in_stride, int rows, int H,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        bf16* __restrict__ out_bf16, float* __restrict__ out_f32, long long out_stride) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const int row_raw = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + (lane >> 3);
  const bool live = row_raw < rows;
  const int row = live ? row_raw : rows - 1;         // dead groups redo the last row (shuffles stay full-warp), no store
  const TIn* xr = x + (long long)row * in_stride;
  float4 v[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) v[i] = load4<TIn>(xr + (i * 8 + sub) * 4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int sh = 4; sh >= 1; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  const float mean = s / (float)H;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int sh = 4; sh >= 1; sh >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, sh);
  const float rstd = rsqrtf(ss / (float)H + eps);
  if (!live) return;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int col = (i * 8 + sub) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (long long)row * out_stride + col) = o;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(out_bf16 + (long long)row * out_stride + col) = pk;
    }
  }
}

template <typename TIn>
static void launch_layernorm(const TIn* x, long long in_stride, int rows, int H, const float* gamma, const float* beta,
                             float eps, bf16* out_bf16, float* out_f32, long long out_stride, cudaStream_t st) {
  const int warps = 8;
  const int blocks = (rows + warps - 1) / warps;
#define AGB_LN(NV4)                                                                                                  \
  layernorm_reg_kernel<TIn, NV4><<<blocks, warps * 32, 0, st>>>(x, in_stride, rows, H, gamma, beta, eps, out_bf16, \
                                                                out_f32, out_stride)
  if (H % 32 == 0 && H < 128) {
    const int nblocks = (rows + warps * 4 - 1) / (warps * 4);
#define AGB_LNN(NV4)                                                                                                     \
  layernorm_narrow_kernel<TIn, NV4><<<nblocks, warps * 32, 0, st>>>(x, in_stride, rows, H, gamma, beta, eps, out_bf16, \
                                                                    out_f32, out_stride)
    if (H == 32) AGB_LNN(1);
    else if (H == 64) AGB_LNN(2);
    else AGB_LNN(3);
#undef AGB_LNN
    return;
  }
  switch ((H % 128 == 0 && H <= 1024) ? H / 128 : 0) {
    case 1: AGB_LN(1); break;
    case 2: AGB_LN(2); break;
    case 3: AGB_LN(3); break;
    case 4: AGB_LN(4); break;
    case 6: AGB_LN(6); break;
    case 8: AGB_LN(8); break;
    default:
      layernorm_kernel<TIn><<<blocks, warps * 32, 0, st>>>(x, in_stride, rows, H, gamma, beta, eps, out_bf16, out_f32,
                                                           out_stride);
  }
#undef AGB_LN
}

int layernorm(const void* x, int in_bf16, long long in_stride, int rows, int H, const float* gamma,
              const float* beta, float eps, bf16* out_bf16, float* out_f32, long long out_stride,
              cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && H > 0 && (H % 4) == 0, "LayerNorm width must be a multiple of 4");
  AGB_REQUIRE((in_stride % 4) == 0 && (out_stride % 4) == 0, "row strides must be multiples of 4");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(x && gamma && beta && (out_bf16 || out_f32), "null pointer");
  if (in_bf16)
    launch_layernorm<bf16>(static_cast<const bf16*>(x), in_stride, rows, H, gamma, beta, eps, out_bf16, out_f32, out_stride, st);
  else
    launch_layernorm<float>(static_cast<const float*>(x), in_stride, rows, H, gamma, beta, eps, out_bf16, out_f32, out_stride, st);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Entry point of the LayerNorm-folded GEMM chain (agb_gemm_bf16_fused): bf16 copy of the fp32 residual
// stream + per-row (sum, sum of squares).  One warp per row, the row is read from HBM once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rowstats_cast_kernel(const float* __restrict__ x, long long ldx, int rows, int H, bf16* __restrict__ out,
                     long long ldo, float2* __restrict__ stats) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)row * ldx;
  float s = 0.f, ss = 0.f;
  for (int c = lane * 4; c < H; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += (v.x + v.y) + (v.z + v.w);
    ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + (long long)row * ldo + c) = pk;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane == 0) stats[row] = make_float2(s, ss);
}

int rowstats_cast(const float* x, long long ldx, int rows, int H, bf16* out, long long ldo, float* stats,
                  cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && H > 0 && (H % 4) == 0 && (ldx % 4) == 0 && (ldo % 4) == 0, "row width / strides must be multiples of 4");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(x && out && stats, "null pointer");
  const int warps = 8;
  rowstats_cast_kernel<<<(rows + warps - 1) / warps, warps * 32, 0, st>>>(x, ldx, rows, H, out, ldo,
                                                                          reinterpret_cast<float2*>(stats));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (weights, once per load) and bf16 -> fp32
// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = pk;
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16(in[j]);
  }
}
int cast_f32_to_bf16(const float* in, bf16* out, long long n, cudaStream_t st) {
  if (n <= 0) return AGB_OK;
  AGB_REQUIRE(in && out, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "alignment");
  const long long threads = (n + 3) / 4;
  cast_f32_bf16_kernel<<<(int)((threads + 255) / 256), 256, 0, st>>>(in, out, n);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// ViT patchify (im2col of the stride-16 conv, reference models/vanilla_vit.py:279-284):
// images (B,C,px,px) fp32 NCHW -> patches (B*gh*gw, C*P*P) in Conv2d weight order (c, py, px).
// One thread moves 4 consecutive pixels of one patch row (16 B read, 8/16 B write).
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void vit_im2col_kernel(const float* __restrict__ img, int B, int C, int px, int P,
                                  TOut* __restrict__ out) {
  const int g = px / P;
  const int K = C * P * P;
  const long long total4 = (long long)B * g * g * K / 4;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total4) return;
  const long long e = gid * 4;
  const int k = (int)(e % K);
  const long long prow = e / K;
  const int b = (int)(prow / (g * g));
  const int pidx = (int)(prow % (g * g));
  const int gy = pidx / g, gx = pidx % g;
  const int c = k / (P * P), py = (k / P) % P, pxx = k % P;
  const float4 v = *reinterpret_cast<const float4*>(
      img + (((long long)b * C + c) * px + (gy * P + py)) * px + gx * P + pxx);
  if (sizeof(TOut) == 4) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + e) = v;
  } else {
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out) + e) = pk;
  }
}

int vit_im2col(const float* img, int B, int C, int px, int P, void* out, int out_bf16, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && C > 0 && P > 0 && px % P == 0 && P % 4 == 0, "patchify shape");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(img && out, "null pointer");
  const long long total4 = (long long)B * px * px * C / 4;
  const int blocks = (int)((total4 + 255) / 256);
  if (out_bf16) vit_im2col_kernel<bf16><<<blocks, 256, 0, st>>>(img, B, C, px, P, static_cast<bf16*>(out));
  else vit_im2col_kernel<float><<<blocks, 256, 0, st>>>(img, B, C, px, P, static_cast<float*>(out));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// ViT embedding assembly (reference models/vanilla_vit.py:242-253): x[b*S+s, 0] = cls + pos[0],
// x[b*S+s, 1+p] = patch_emb[b, p] + pos[1+p]; each image is broadcast to its S coalition rows here
// (this replaces the reference's Xs_EXT pixel replication, scripts/train_explainer.py:159-163).
// ------------------------------------------------------------------------------------------------
__global__ void vit_assemble_kernel(const float* __restrict__ patch_emb, const float* __restrict__ cls,
                                    const float* __restrict__ pos, int B, int S, int T, int H,
                                    float* __restrict__ x) {
  const long long total4 = (long long)B * T * H / 4;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total4) return;
  const long long e = gid * 4;
  const int hcol = (int)(e % H);
  const int t = (int)((e / H) % T);
  const int b = (int)(e / ((long long)H * T));
  float4 v = (t == 0) ? __ldg(reinterpret_cast<const float4*>(cls + hcol))
                      : *reinterpret_cast<const float4*>(patch_emb + ((long long)b * (T - 1) + (t - 1)) * H + hcol);
  const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + (long long)t * H + hcol));
  v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
  for (int s = 0; s < S; ++s)
    *reinterpret_cast<float4*>(x + (((long long)b * S + s) * T + t) * H + hcol) = v;
}

int vit_assemble(const float* patch_emb, const float* cls, const float* pos, int B, int S, int T, int H,
                 float* x, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && S > 0 && T > 1 && H % 4 == 0, "embedding shape");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(patch_emb && cls && pos && x, "null pointer");
  const long long total4 = (long long)B * T * H / 4;
  vit_assemble_kernel<<<(int)((total4 + 255) / 256), 256, 0, st>>>(patch_emb, cls, pos, B, S, T, H, x);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Row fan-out: S consecutive copies of every source row (16-byte vectors; one load, S coalesced stores per thread).
// The first-block sharing path embeds every input once and fans its residual stream out to the S coalition rows here.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
repeat_rows_kernel(const uint4* __restrict__ src, long long total16, long long row16, int S, uint4* __restrict__ dst) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total16) return;
  const long long b = gid / row16, c = gid - b * row16;
  const uint4 v = src[gid];
  uint4* d = dst + b * S * row16 + c;
#pragma unroll 4
  for (int s = 0; s < S; ++s) d[(long long)s * row16] = v;
}

int repeat_rows(const void* src, int B, long long row_bytes, int S, void* dst, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && S > 0 && row_bytes > 0 && (row_bytes % 16) == 0, "rows of whole 16-byte vectors");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(src && dst, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "alignment");
  const long long row16 = row_bytes / 16, total16 = row16 * B;
  AGB_REQUIRE((total16 + 255) / 256 <= 0x7fffffffLL, "grid limits");
  repeat_rows_kernel<<<(unsigned)((total16 + 255) / 256), 256, 0, st>>>(static_cast<const uint4*>(src), total16, row16, S,
                                                                       static_cast<uint4*>(dst));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Kept-first token order of every (input, coalition) row (ViT "logit := 0" masks, surrogate evaluation): a stable
// partition of the tokens into kept (bit set; CLS = bit 0 always is) and masked ones.  One warp per row:
//   order[r, q]  = token at position q          pos[r, t] = position of token t          nkeep[r] = kept tokens
//   prefix[r, :] = packed mask of the permuted row (bits [0, nkeep) set)
// Everything between the attentions is token-wise and attention is permutation-equivariant, so evaluating the permuted
// rows leaves the CLS output unchanged; the attention kernels then see the masked keys as one contiguous tail.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
kept_first_order_kernel(const uint32_t* __restrict__ packed, int rows, int words, int T, uint8_t* __restrict__ order,
                        uint8_t* __restrict__ pos, int* __restrict__ nkeep, uint32_t* __restrict__ prefix) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint32_t* m = packed + (long long)r * words;
  const int nw = (T + 31) >> 5;
  int total = 0;
  for (int w = 0; w < nw; ++w) {
    uint32_t bits = __ldg(m + w);
    if (w == 0) bits |= 1u;                                   // CLS is always kept
    if ((w + 1) * 32 > T) bits &= (T & 31) ? ((1u << (T & 31)) - 1u) : 0xFFFFFFFFu;
    total += __popc(bits);
  }
  int kept_before = 0;
  for (int w = 0; w < nw; ++w) {
    uint32_t bits = __ldg(m + w);
    if (w == 0) bits |= 1u;
    if ((w + 1) * 32 > T) bits &= (T & 31) ? ((1u << (T & 31)) - 1u) : 0xFFFFFFFFu;
    const int t = w * 32 + lane;
    if (t < T) {
      const int kb = kept_before + __popc(bits & ((1u << lane) - 1u));
      const bool kept = (bits >> lane) & 1u;
      const int q = kept ? kb : total + (t - kb);             // masked tokens follow the kept ones, in token order
      order[(long long)r * T + q] = (uint8_t)t;
      pos[(long long)r * T + t] = (uint8_t)q;
    }
    kept_before += __popc(bits);
  }
  if (lane == 0) nkeep[r] = total;
  for (int w = lane; w < words; w += 32) {
    const int lo = w * 32;
    prefix[(long long)r * words + w] = total >= lo + 32 ? 0xFFFFFFFFu : (total > lo ? ((1u << (total - lo)) - 1u) : 0u);
  }
}

int kept_first_order(const uint32_t* packed, int rows, int words, int T, uint8_t* order, uint8_t* pos, int* nkeep,
                     uint32_t* prefix, cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && T > 0 && T <= 256 && words * 32 >= T, "kept-first order: T <= 256 tokens");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(packed && order && pos && nkeep && prefix, "null pointer");
  kept_first_order_kernel<<<(rows + 7) / 8, 256, 0, st>>>(packed, rows, words, T, order, pos, nkeep, prefix);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// dst[(r, q), :] = src[(r / S, order[r, q]), :]: the residual stream of every coalition row in kept-first token order,
// gathered from the per-input embeddings (replaces repeat_rows on that path)
// one warp per destination token row (no integer divisions per element; 8 rows per CTA)
__global__ void __launch_bounds__(256)
gather_token_rows_kernel(const uint4* __restrict__ src, const uint8_t* __restrict__ order, long long n_tok, int row16, int T,
                         int S, uint4* __restrict__ dst) {
  const long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);           // destination token row = r * T + q
  if (tok >= n_tok) return;
  const int lane = threadIdx.x & 31;
  const long long r = tok / T;
  const int t = order[tok];
  const uint4* s = src + ((r / S) * T + t) * row16;
  uint4* d = dst + tok * row16;
#pragma unroll 2
  for (int c = lane; c < row16; c += 32) d[c] = __ldg(s + c);
}

// hi/lo planes of an fp32 value: hi = bf16(x), lo = bf16(x - hi); hi + lo carries 16 significant bits of x
__device__ __forceinline__ void split_hilo8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  hi.x = pack_bf16x2(a.x, a.y); hi.y = pack_bf16x2(a.z, a.w); hi.z = pack_bf16x2(b.x, b.y); hi.w = pack_bf16x2(b.z, b.w);
  lo.x = pack_bf16x2(a.x - bf16_lo(hi.x), a.y - bf16_hi(hi.x));
  lo.y = pack_bf16x2(a.z - bf16_lo(hi.y), a.w - bf16_hi(hi.y));
  lo.z = pack_bf16x2(b.x - bf16_lo(hi.z), b.y - bf16_hi(hi.z));
  lo.w = pack_bf16x2(b.z - bf16_lo(hi.w), b.w - bf16_hi(hi.w));
}

// gather_token_rows with the destination written as hi/lo planes (the residual stream of agb_gemm_bf16_hilo):
// one warp per destination token row, 8 fp32 values -> 16 B of each plane per lane and step
__global__ void __launch_bounds__(256)
gather_token_rows_hilo_kernel(const float4* __restrict__ src, const uint8_t* __restrict__ order, long long n_tok, int row8, int T,
                              int S, uint4* __restrict__ hi, uint4* __restrict__ lo) {
  const long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= n_tok) return;
  const int lane = threadIdx.x & 31;
  const long long r = tok / T;
  const int t = order[tok];
  const float4* s = src + ((r / S) * T + t) * (2LL * row8);
  uint4* dh = hi + tok * row8;
  uint4* dl = lo + tok * row8;
  for (int c = lane; c < row8; c += 32) {
    uint4 h, l;
    split_hilo8(__ldg(s + 2 * c), __ldg(s + 2 * c + 1), h, l);
    dh[c] = h;
    dl[c] = l;
  }
}

int gather_token_rows_hilo(const float* src, const uint8_t* order, int rows, int T, int S, int H, bf16* hi, bf16* lo,
                           cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && T > 0 && S > 0 && H > 0 && H % 8 == 0 && rows % S == 0, "shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(src && order && hi && lo, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(lo) & 15) == 0, "alignment");
  const long long n_tok = (long long)rows * T;
  AGB_REQUIRE((n_tok + 7) / 8 <= 0x7fffffffLL, "grid limits");
  gather_token_rows_hilo_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(reinterpret_cast<const float4*>(src), order, n_tok,
                                                                            H / 8, T, S, reinterpret_cast<uint4*>(hi),
                                                                            reinterpret_cast<uint4*>(lo));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

__global__ void __launch_bounds__(256)
split_hilo_kernel(const float4* __restrict__ x, long long n8, uint4* __restrict__ hi, uint4* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 h, l;
  split_hilo8(__ldg(x + 2 * i), __ldg(x + 2 * i + 1), h, l);
  hi[i] = h;
  lo[i] = l;
}

int split_hilo(const float* x, long long n, bf16* hi, bf16* lo, cudaStream_t st) {
  AGB_REQUIRE(n >= 0 && n % 8 == 0, "element count must be a multiple of 8");
  if (n == 0) return AGB_OK;
  AGB_REQUIRE(x && hi && lo, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(lo) & 15) == 0, "alignment");
  const long long n8 = n / 8;
  AGB_REQUIRE((n8 + 255) / 256 <= 0x7fffffffLL, "grid limits");
  split_hilo_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n8,
                                                                 reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int gather_token_rows(const void* src, const uint8_t* order, int rows, int T, int S, long long row_bytes, void* dst,
                      cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && T > 0 && S > 0 && rows % S == 0 && row_bytes > 0 && (row_bytes % 16) == 0, "gather_token_rows shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(src && order && dst, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "alignment");
  const long long row16 = row_bytes / 16, n_tok = (long long)rows * T;
  AGB_REQUIRE((n_tok + 7) / 8 <= 0x7fffffffLL && row16 <= 0x7fffffffLL, "grid limits");
  gather_token_rows_kernel<<<(unsigned)((n_tok + 7) / 8), 256, 0, st>>>(static_cast<const uint4*>(src), order, n_tok, (int)row16, T, S,
                                                                       static_cast<uint4*>(dst));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// BERT embeddings (reference models/vanilla_bert.py:307-325): LN(word[id] + type[tt] + pos[t]),
// token_type_ids are all zero on this path (reference recipes/vanilla_bert.py:289).  One warp per
// token; the normalised row is broadcast to the S coalition rows of its input.
// ------------------------------------------------------------------------------------------------
__global__ void bert_embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                  const float* __restrict__ pos, const float* __restrict__ type0,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                  int B, int S, int T, int H, int vocab, float* __restrict__ x) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= B * T) return;
  const int lane = threadIdx.x & 31;
  const int b = tok / T, t = tok % T;
  long long id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float* w = word + id * H;
  const float* pp = pos + (long long)t * H;
  float s = 0.f;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 a = *reinterpret_cast<const float4*>(w + i);
    const float4 c = __ldg(reinterpret_cast<const float4*>(type0 + i));
    const float4 d = __ldg(reinterpret_cast<const float4*>(pp + i));
    s += ((a.x + c.x) + d.x) + ((a.y + c.y) + d.y) + ((a.z + c.z) + d.z) + ((a.w + c.w) + d.w);
  }
  const float mean = warp_sum(s) / (float)H;
  float ss = 0.f;
  for (int i = lane * 4; i < H; i += 128) {
    const float4 a = *reinterpret_cast<const float4*>(w + i);
    const float4 c = __ldg(reinterpret_cast<const float4*>(type0 + i));
    const float4 d = __ldg(reinterpret_cast<const float4*>(pp + i));
    const float e0 = ((a.x + c.x) + d.x) - mean, e1 = ((a.y + c.y) + d.y) - mean;
    const float e2 = ((a.z + c.z) + d.z) - mean, e3 = ((a.w + c.w) + d.w) - mean;
    ss += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)H + eps);
  for (int i = lane * 4; i < H; i += 128) {
    const float4 a = *reinterpret_cast<const float4*>(w + i);
    const float4 c = __ldg(reinterpret_cast<const float4*>(type0 + i));
    const float4 d = __ldg(reinterpret_cast<const float4*>(pp + i));
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + i));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + i));
    float4 o;
    o.x = (((a.x + c.x) + d.x) - mean) * rstd * g.x + bb.x;
    o.y = (((a.y + c.y) + d.y) - mean) * rstd * g.y + bb.y;
    o.z = (((a.z + c.z) + d.z) - mean) * rstd * g.z + bb.z;
    o.w = (((a.w + c.w) + d.w) - mean) * rstd * g.w + bb.w;
    for (int sidx = 0; sidx < S; ++sidx)
      *reinterpret_cast<float4*>(x + (((long long)b * S + sidx) * T + t) * H + i) = o;
  }
}

int bert_embed(const int64_t* ids, const float* word, const float* pos, const float* type0,
               const float* gamma, const float* beta, float eps, int B, int S, int T, int H, int vocab,
               float* x, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && S > 0 && T > 0 && H % 4 == 0 && vocab > 0, "embedding shape");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(ids && word && pos && type0 && gamma && beta && x, "null pointer");
  const int warps = 8;
  bert_embed_kernel<<<(B * T + warps - 1) / warps, warps * 32, 0, st>>>(ids, word, pos, type0, gamma, beta, eps, B,
                                                                        S, T, H, vocab, x);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// CLS heads -> class PROBABILITIES (fp32 end to end: these are the Shapley value function v(S)).
//   ViT  (reference models/vanilla_vit.py:213, 52-56): softmax(W * LN_final(x[row, 0]) + b)
//   BERT (reference models/vanilla_bert.py:73-77, 615-619): softmax(W * tanh(Wp * x[row, 0] + bp) + b)
// One CTA per row (BERT with >= 512 rows: 8 rows per CTA, cls_head_pool_rows_kernel); H <= 4096, C <= 64.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

__global__ void cls_head_kernel(const float* __restrict__ x, long long row_stride, int H, int C, int mode,
                                const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
                                const float* __restrict__ wp, const float* __restrict__ bp,
                                const float* __restrict__ wc, const float* __restrict__ bc,
                                float* __restrict__ probs, float* __restrict__ logits_out) {
  extern __shared__ float sm[];
  float* h = sm;            // H
  float* h2 = sm + H;       // H (BERT pooled)
  float* lg = sm + 2 * H;   // C
  float* red = sm + 2 * H + 64;
  const float* xr = x + (long long)blockIdx.x * row_stride;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const float* feat = h;
  if (mode == 0) {
    float s = 0.f;
    for (int i = tid; i < H; i += nt) { h[i] = xr[i]; s += xr[i]; }
    const float mean = block_sum(s, red) / (float)H;
    float ss = 0.f;
    for (int i = tid; i < H; i += nt) { const float d = h[i] - mean; ss += d * d; }
    const float rstd = rsqrtf(block_sum(ss, red) / (float)H + eps);
    for (int i = tid; i < H; i += nt) h[i] = (h[i] - mean) * rstd * ln_g[i] + ln_b[i];
    __syncthreads();
  } else {
    for (int i = tid; i < H; i += nt) h[i] = xr[i];
    __syncthreads();
    for (int j = warp; j < H; j += nw) {
      const float* wrow = wp + (long long)j * H;
      float a = 0.f;
      for (int i = lane; i < H; i += 32) a = fmaf(h[i], wrow[i], a);
      a = warp_sum(a);
      if (lane == 0) h2[j] = tanhf(a + bp[j]);
    }
    __syncthreads();
    feat = h2;
  }
  for (int c = warp; c < C; c += nw) {
    const float* wrow = wc + (long long)c * H;
    float a = 0.f;
    for (int i = lane; i < H; i += 32) a = fmaf(feat[i], wrow[i], a);
    a = warp_sum(a);
    if (lane == 0) lg[c] = a + bc[c];
  }
  __syncthreads();
  if (tid == 0) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, lg[c]);
    float z = 0.f;
    for (int c = 0; c < C; ++c) z += expf(lg[c] - m);
    for (int c = 0; c < C; ++c) {
      probs[(long long)blockIdx.x * C + c] = expf(lg[c] - m) / z;
      if (logits_out) logits_out[(long long)blockIdx.x * C + c] = lg[c];
    }
  }
}

// BERT head for many rows: RB rows per 32-warp CTA, so the H x H pooler weight is streamed from L2 once per RB rows instead of
// once per row, with 8 loads in flight per lane (the one-row kernel's serial pooler loop — H / 8 output columns per warp, each
// a chain of L2-latency loads — cost 158 us per call at BERT-base with 1024 rows).  Every dot product
// keeps the one-row kernel's summation order (lane-strided FMAs, then the warp tree), so the probabilities are bit-identical.
template <int RB>
__global__ void __launch_bounds__(1024)
cls_head_pool_rows_kernel(const float* __restrict__ x, long long row_stride, int rows, int H, int C,
                          const float* __restrict__ wp, const float* __restrict__ bp, const float* __restrict__ wc,
                          const float* __restrict__ bc, float* __restrict__ probs, float* __restrict__ logits_out) {
  extern __shared__ float sm[];
  float* h = sm;                 // RB x H : x[row, 0]
  float* h2 = h + RB * H;        // RB x H : pooled
  float* lg = h2 + RB * H;       // RB x 64
  const int row0 = blockIdx.x * RB;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int idx = tid; idx < RB * H; idx += nt) {
    const int r = idx / H, i = idx - r * H;
    h[idx] = x[(long long)min(row0 + r, rows - 1) * row_stride + i];
  }
  __syncthreads();
  for (int j = warp; j < H; j += nw) {
    const float* wrow = wp + (long long)j * H;
    float a[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) a[r] = 0.f;
#pragma unroll 8
    for (int i = lane; i < H; i += 32) {       // 8 independent weight loads in flight per lane: the loop is L2-latency bound
      const float w = __ldg(wrow + i);
#pragma unroll
      for (int r = 0; r < RB; ++r) a[r] = fmaf(h[r * H + i], w, a[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) a[r] = warp_sum(a[r]);
    if (lane == 0) {
      const float b = bp[j];
#pragma unroll
      for (int r = 0; r < RB; ++r) h2[r * H + j] = tanhf(a[r] + b);
    }
  }
  __syncthreads();
  for (int c = warp; c < C; c += nw) {
    const float* wrow = wc + (long long)c * H;
    float a[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) a[r] = 0.f;
#pragma unroll 8
    for (int i = lane; i < H; i += 32) {       // 8 independent weight loads in flight per lane: the loop is L2-latency bound
      const float w = __ldg(wrow + i);
#pragma unroll
      for (int r = 0; r < RB; ++r) a[r] = fmaf(h2[r * H + i], w, a[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) a[r] = warp_sum(a[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < RB; ++r) lg[r * 64 + c] = a[r] + bc[c];
    }
  }
  __syncthreads();
  if (tid < RB && row0 + tid < rows) {
    const float* l = lg + tid * 64;
    const long long row = row0 + tid;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, l[c]);
    float z = 0.f;
    for (int c = 0; c < C; ++c) z += expf(l[c] - m);
    for (int c = 0; c < C; ++c) {
      probs[row * C + c] = expf(l[c] - m) / z;
      if (logits_out) logits_out[row * C + c] = l[c];
    }
  }
}

int cls_head(const float* x, long long row_stride, int rows, int H, int C, int mode, const float* ln_g,
             const float* ln_b, float eps, const float* wp, const float* bp, const float* wc,
             const float* bc, float* probs, float* logits_out, cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && H > 0 && H <= 4096 && C > 0 && C <= 64, "head shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(x && wc && bc && probs, "null pointer");
  AGB_REQUIRE(mode == 0 ? (ln_g && ln_b) : (wp && bp), "head parameters");
  constexpr int RB = 8;
  const size_t smem_rows = ((size_t)2 * RB * H + RB * 64) * sizeof(float);
  if (mode == 1 && rows >= 512 && smem_rows <= 100 * 1024) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(cls_head_pool_rows_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    cls_head_pool_rows_kernel<RB><<<(rows + RB - 1) / RB, 1024, smem_rows, st>>>(x, row_stride, rows, H, C, wp, bp, wc, bc, probs,
                                                                              logits_out);
    AGB_CHECK_CUDA(cudaGetLastError());
    return AGB_OK;
  }
  const size_t smem = (2 * H + 64 + 32) * sizeof(float);
  cls_head_kernel<<<rows, 256, smem, st>>>(x, row_stride, H, C, mode, ln_g, ln_b, eps, wp, bp, wc, bc, probs,
                                           logits_out);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

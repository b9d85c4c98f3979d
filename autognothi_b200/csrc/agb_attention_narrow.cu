// Narrow-head attention on the warp-level tensor-core path (mma.sync m16n8k8 / m16n8k16, bf16 operands, fp32 accumulators):
// the inference forward of the LTT side ladders (reference models/ltt_vit.py:386-396 / models/ltt_bert.py:437-451: hidden size
// s_attn_hidden_size split over the backbone's head count -> head dims 8 / 16 / 32) on every coalition row.
//
// A head this narrow cannot feed tcgen05 (one 128 x N x 16 UMMA would be 7/8 padding at d = 8 and the per-(row, head) problem
// is 197 x 197 x 8), so the tile is the warp-level 16 x 8: one CTA per (row, head), the kept keys compacted into shared memory
// once (K row-major, V transposed, both bf16), every warp owns 16-query tiles and streams the keys 16 at a time with an online
// softmax (exp2 domain) — S accumulators are re-packed in registers as the A operand of the P V product, P never leaves the
// register file.  Mask semantics as in the CUDA-core kernel it replaces (agb_simt.cu, attention_narrow_kernel):
//   AGB_MASK_NEGINF (BERT, reference models/vanilla_bert.py:520-523): masked keys are absent;
//   AGB_MASK_MUL0   (ViT,  reference models/vanilla_vit.py:449-450): masked keys carry the logit 0 — all of them together enter
//                   the softmax as ONE virtual key of weight n_masked whose value row is the fp32 sum of their V rows.
// Attention dropout (training mode) stays on the CUDA-core kernel.
#include <stdlib.h>

#include "agb_common.cuh"

namespace agb {

__device__ __forceinline__ void mma_bf16_16x8x16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                 uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_bf16_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

// shared-memory strides in 32-bit words, chosen so that the B-fragment reads (8 rows x 4 consecutive words) hit 32 banks
template <int D> struct NarrowLayout {
  static constexpr int KS = D / 2 + (D == 8 ? 0 : 4);                     // K row (one key): 4 / 12 / 20 words
  __host__ __device__ static int tp(int T) { return (T + 15) & ~15; }      // key capacity, whole 16-key blocks
  __host__ __device__ static int vs(int T) { return tp(T) / 2 + 4; }       // V^T row (one head dim): == 4 (mod 8) words
  __host__ __device__ static size_t bytes(int T, int words) {
    // the key-bit words are padded to whole 16-byte vectors: the compiler reads them with LDS.128 (sM is 16-byte aligned)
    return ((size_t)tp(T) * KS + (size_t)D * vs(T) + D + ((words + 3) & ~3)) * 4;
  }
};

template <int D>
__global__ void __launch_bounds__(256)
attention_narrow_mma_kernel(const bf16* __restrict__ qkv, const uint32_t* __restrict__ mask, int words, int T, int H, int heads,
                            int mode, float scale_log2, bf16* __restrict__ ctx, const int* __restrict__ cu) {
  // cu != nullptr: packed variable-length rows (masked-token dropping): row r owns the tokens [cu[r], cu[r+1]) of
  // qkv (total, 3H) / ctx (total, H), every packed token is a live key; T is then the capacity (max row length)
  typedef NarrowLayout<D> L;
  constexpr int KS = L::KS;
  constexpr int V8 = D / 8;
  extern __shared__ __align__(16) uint32_t sm_nm[];
  const int Tcap = T;
  const int row = blockIdx.x / heads, head = blockIdx.x % heads;
  int tok0 = 0;
  if (cu != nullptr) {
    tok0 = __ldg(cu + row);
    T = __ldg(cu + row + 1) - tok0;
  }
  const int Tp = L::tp(Tcap), VS = L::vs(Tcap);
  uint32_t* sK = sm_nm;                                   // Tp x KS : kept keys, compacted, bf16 pairs
  uint32_t* sVt = sK + Tp * KS;                           // D x VS  : their V rows transposed (word i = keys 2i, 2i+1)
  float* sVm = reinterpret_cast<float*>(sVt + D * VS);    // D       : fp32 sum of the V rows of the ViT-masked keys
  uint32_t* sM = reinterpret_cast<uint32_t*>(sVm + D);    // words (padded to a multiple of 4): key bits of this row, bits >= T cleared
  const long long first = cu != nullptr ? (long long)tok0 : (long long)row * T;
  const bf16* base = qkv + first * 3 * H + head * D;
  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;

  for (int e = tid; e < ((words + 3) & ~3); e += nthr) {  // pad words (whole 16-byte vectors) are zero
    uint32_t w = e >= words ? 0u : (cu != nullptr ? 0xFFFFFFFFu : mask[(long long)row * words + e]);
    const int live = T - 32 * e;                          // bits of this word that are tokens
    if (live < 32) w = live <= 0 ? 0u : (w & ((1u << live) - 1u));
    sM[e] = w;
  }
  for (int e = tid; e < D; e += nthr) sVm[e] = 0.f;
  __syncthreads();
  int nkept = 0;
  for (int w = 0; w < words; ++w) nkept += __popc(sM[w]);
  const int nkp = (nkept + 15) & ~15;
  const int nmasked = mode == AGB_MASK_MUL0 ? T - nkept : 0;
  {
    // one pass over the row's keys: a kept key goes to its compacted slot (= number of kept keys before it: popcount
    // prefix of the key bits), a ViT-masked key (logit 0) only adds its V row to the virtual key's value row
    float vm[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) vm[u] = 0.f;
    unsigned short* vt = reinterpret_cast<unsigned short*>(sVt);
    for (int e = tid; e < T * V8; e += nthr) {            // nthr % V8 == 0: a thread keeps its 8-dim slice v
      const int j = e / V8, v = e % V8;
      const uint32_t wj = sM[j >> 5];
      const bool kept = (wj >> (j & 31)) & 1u;
      if (!kept && mode != AGB_MASK_MUL0) continue;
      const bf16* src = base + (long long)j * 3 * H + v * 8;
      const uint4 v4 = *reinterpret_cast<const uint4*>(src + 2 * H);
      const uint32_t vw[4] = {v4.x, v4.y, v4.z, v4.w};
      if (kept) {
        const uint4 k4 = *reinterpret_cast<const uint4*>(src + H);
        int jj = __popc(wj & ((1u << (j & 31)) - 1u));
        for (int w = 0; w < (j >> 5); ++w) jj += __popc(sM[w]);
        *reinterpret_cast<uint4*>(sK + jj * KS + v * 4) = k4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          vt[(size_t)(v * 8 + 2 * u) * (2 * VS) + jj] = (unsigned short)(vw[u] & 0xFFFFu);
          vt[(size_t)(v * 8 + 2 * u + 1) * (2 * VS) + jj] = (unsigned short)(vw[u] >> 16);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          vm[2 * u] += bf16_lo(vw[u]);
          vm[2 * u + 1] += bf16_hi(vw[u]);
        }
      }
    }
    for (int e = tid; e < (nkp - nkept) * V8; e += nthr) {    // the pad keys of the last 16-key block: zero rows
      const int jj = nkept + e / V8, v = e % V8;
      *reinterpret_cast<uint4*>(sK + jj * KS + v * 4) = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 8; ++u) vt[(size_t)(v * 8 + u) * (2 * VS) + jj] = 0;
    }
    if (nmasked > 0) {                                        // block-uniform
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int sh = 16; sh >= V8; sh >>= 1) vm[u] += __shfl_xor_sync(0xffffffffu, vm[u], sh);
      }
      if (lane < V8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) atomicAdd(&sVm[lane * 8 + u], vm[u]);
      }
    }
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  for (int q0 = warp * 16; q0 < T; q0 += (nthr >> 5) * 16) {
    const int qa = q0 + g, qb = q0 + g + 8;                // this thread's two query rows (C-fragment rows g, g + 8)
    const bf16* pa = base + (long long)min(qa, T - 1) * 3 * H;
    const bf16* pb = base + (long long)min(qb, T - 1) * 3 * H;
    uint32_t aq[V8 == 1 ? 2 : 2 * V8];
    if (D == 8) {
      aq[0] = *reinterpret_cast<const uint32_t*>(pa + 2 * t);
      aq[1] = *reinterpret_cast<const uint32_t*>(pb + 2 * t);
    } else {
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        aq[4 * kk + 0] = *reinterpret_cast<const uint32_t*>(pa + 16 * kk + 2 * t);
        aq[4 * kk + 1] = *reinterpret_cast<const uint32_t*>(pb + 16 * kk + 2 * t);
        aq[4 * kk + 2] = *reinterpret_cast<const uint32_t*>(pa + 16 * kk + 8 + 2 * t);
        aq[4 * kk + 3] = *reinterpret_cast<const uint32_t*>(pb + 16 * kk + 8 + 2 * t);
      }
    }
    float o[V8][4];
#pragma unroll
    for (int dt = 0; dt < V8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
    // the virtual key (logit 0) seeds the running maximum, so it can be added after the loop without a rescale
    float ma = nmasked > 0 ? 0.f : -INFINITY, mb = ma, la = 0.f, lb = 0.f;
    for (int kb = 0; kb < nkp; kb += 16) {
      float s[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const uint32_t* kr = sK + (kb + nt * 8 + g) * KS;
        if (D == 8) {
          mma_bf16_16x8x8(s[nt], aq[0], aq[1], kr[t]);
        } else {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            mma_bf16_16x8x16(s[nt], aq[4 * kk], aq[4 * kk + 1], aq[4 * kk + 2], aq[4 * kk + 3], kr[8 * kk + t], kr[8 * kk + 4 + t]);
        }
        if (kb + 16 > nkept) {                               // last block: its pad keys leave the softmax
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (kb + nt * 8 + 2 * t + (e & 1) >= nkept) s[nt][e] = -INFINITY;
        }
      }
      // s holds the raw dot products; the logit scale (> 0) is applied to the row maximum and inside the exponent's FMA
      float xa = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
      float xb = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
      xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 1));
      xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 1));
      xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 2));
      xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 2));
      // every block holds at least one live key (kb < nkept), so the new maxima are finite
      const float na = fmaxf(ma, xa * scale_log2), nb = fmaxf(mb, xb * scale_log2);
      const float ca = ex2_approx(ma - na), cb = ex2_approx(mb - nb);
      ma = na;
      mb = nb;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        s[nt][0] = ex2_approx(fmaf(s[nt][0], scale_log2, -na));
        s[nt][1] = ex2_approx(fmaf(s[nt][1], scale_log2, -na));
        s[nt][2] = ex2_approx(fmaf(s[nt][2], scale_log2, -nb));
        s[nt][3] = ex2_approx(fmaf(s[nt][3], scale_log2, -nb));
      }
      la = fmaf(la, ca, (s[0][0] + s[0][1]) + (s[1][0] + s[1][1]));     // this thread's 4 columns; quad-reduced at the end
      lb = fmaf(lb, cb, (s[0][2] + s[0][3]) + (s[1][2] + s[1][3]));
      const uint32_t p0 = pack_bf16x2(s[0][0], s[0][1]), p1 = pack_bf16x2(s[0][2], s[0][3]);
      const uint32_t p2 = pack_bf16x2(s[1][0], s[1][1]), p3 = pack_bf16x2(s[1][2], s[1][3]);
#pragma unroll
      for (int dt = 0; dt < V8; ++dt) {
        o[dt][0] *= ca;
        o[dt][1] *= ca;
        o[dt][2] *= cb;
        o[dt][3] *= cb;
        const uint32_t* vr = sVt + (dt * 8 + g) * VS + (kb >> 1);
        mma_bf16_16x8x16(o[dt], p0, p1, p2, p3, vr[t], vr[t + 4]);
      }
    }
    la += __shfl_xor_sync(0xffffffffu, la, 1);
    lb += __shfl_xor_sync(0xffffffffu, lb, 1);
    la += __shfl_xor_sync(0xffffffffu, la, 2);
    lb += __shfl_xor_sync(0xffffffffu, lb, 2);
    if (nmasked > 0) {                                     // ma, mb >= 0 here
      const float wa = ex2_approx(-ma), wb = ex2_approx(-mb);
      la = fmaf(wa, (float)nmasked, la);
      lb = fmaf(wb, (float)nmasked, lb);
#pragma unroll
      for (int dt = 0; dt < V8; ++dt) {
        const float v0 = sVm[dt * 8 + 2 * t], v1 = sVm[dt * 8 + 2 * t + 1];
        o[dt][0] = fmaf(wa, v0, o[dt][0]);
        o[dt][1] = fmaf(wa, v1, o[dt][1]);
        o[dt][2] = fmaf(wb, v0, o[dt][2]);
        o[dt][3] = fmaf(wb, v1, o[dt][3]);
      }
    }
    const float ra = la > 0.f ? 1.f / la : 0.f, rb = lb > 0.f ? 1.f / lb : 0.f;
    bf16* ca_ = ctx + (first + qa) * H + head * D + 2 * t;
    bf16* cb_ = ctx + (first + qb) * H + head * D + 2 * t;
#pragma unroll
    for (int dt = 0; dt < V8; ++dt) {
      if (qa < T) *reinterpret_cast<uint32_t*>(ca_ + dt * 8) = pack_bf16x2(o[dt][0] * ra, o[dt][1] * ra);
      if (qb < T) *reinterpret_cast<uint32_t*>(cb_ + dt * 8) = pack_bf16x2(o[dt][2] * rb, o[dt][3] * rb);
    }
  }
}

// warps per (row, head) CTA: AGB_NARROW_WARPS (A/B runs); default 4 (measured at T = 197: 4 warps 188 us, 7 warps 203 us,
// 13 warps 280 us per 1024-row launch — the per-CTA staging favours many small CTAs per SM)
static int narrow_threads(int T) {
  static const int forced = [] {
    const char* e = getenv("AGB_NARROW_WARPS");
    return e != nullptr ? atoi(e) : 0;
  }();
  const int tiles = (T + 15) / 16;
  int warps = forced > 0 ? forced : 4;
  warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);      // __launch_bounds__(256)
  if (warps > tiles) warps = tiles;
  return warps * 32;
}

template <int D>
static int launch_narrow_mma(const bf16* qkv, const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode,
                             bf16* ctx, cudaStream_t st, const int* cu) {
  if (cu != nullptr) words = (T + 31) / 32;     // all-ones key bits, built in shared memory
  const size_t smem = NarrowLayout<D>::bytes(T, words);
  if (smem > 227 * 1024) return AGB_ERR_UNSUPPORTED;
  AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_narrow_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_narrow_mma_kernel<D><<<rows * heads, narrow_threads(T), smem, st>>>(qkv, mask, words, T, H, heads, mode,
                                                                   rsqrtf((float)D) * 1.4426950408889634f, ctx, cu);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// head dim 8 / 16 / 32, bf16 I/O, 16-byte aligned qkv / ctx, H % 8 == 0 (checked by the callers in agb_simt.cu);
// returns AGB_ERR_UNSUPPORTED when a row does not fit shared memory
int attention_narrow_mma(const bf16* qkv, const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode, bf16* ctx,
                         cudaStream_t st, const int* cu) {
  const int d = H / heads;
  return d == 8    ? launch_narrow_mma<8>(qkv, mask, words, rows, T, H, heads, mode, ctx, st, cu)
         : d == 16 ? launch_narrow_mma<16>(qkv, mask, words, rows, T, H, heads, mode, ctx, st, cu)
                   : launch_narrow_mma<32>(qkv, mask, words, rows, T, H, heads, mode, ctx, st, cu);
}

}  // namespace agb

// Coalition-mask kernels: bit packing (a2), the paired Shapley-kernel sampler (a1) and the uniform
// sampler (a15).  Integer/byte work, HBM-bound; one thread produces one packed 32-bit word so every
// store is a coalesced 4-byte lane write and the int64 {0,1} tensors of the reference are only
// touched when a caller asks for them.
//
// Layout contract (include/autognothi_b200.h): bit 0 of word 0 = CLS (always 1, reference
// recipes/vanilla_vit.py:219-224), bit j+1 = player j.
#include "agb_common.cuh"

namespace agb {

// ------------------------------------------------------------------------------------------------
// pack: (rows, n) int64 {0,1}  ->  (rows, words) uint32
// ------------------------------------------------------------------------------------------------
__global__ void pack_masks_kernel(const int64_t* __restrict__ mask, int rows, int n, int prepend_cls,
                                  uint32_t* __restrict__ packed, int words) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)rows * words) return;
  const int row = (int)(gid / words), w = (int)(gid % words);
  const int64_t* src = mask + (long long)row * n;
  uint32_t bits = 0;
#pragma unroll 4
  for (int b = 0; b < 32; ++b) {
    const int tok = w * 32 + b;          // token index (CLS = 0 when prepend_cls)
    const int j = tok - prepend_cls;     // player index
    uint32_t v = 0;
    if (prepend_cls && tok == 0) v = 1u;
    else if (j >= 0 && j < n) v = (src[j] != 0) ? 1u : 0u;
    bits |= v << b;
  }
  packed[gid] = bits;
}

// unpack: packed -> (rows, n) int64; `skip` leading tokens dropped (1 removes the CLS bit)
__global__ void unpack_masks_kernel(const uint32_t* __restrict__ packed, int rows, int n, int skip,
                                    int words, int64_t* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)rows * n) return;
  const int row = (int)(gid / n), j = (int)(gid % n);
  const int tok = j + skip;
  out[gid] = (packed[(long long)row * words + (tok >> 5)] >> (tok & 31)) & 1u;
}

// ------------------------------------------------------------------------------------------------
// Shapley-kernel sampler (reference models/shapley.py:56-79, 131-135)
// ------------------------------------------------------------------------------------------------
// inverse-CDF subset-size draw: position = max(count(u >= prefix[k]) - 1, 0); threshold =
// float32(1/n) * float32(position)   (python-float scalar times int64 tensor -> float32 product)
__device__ __forceinline__ float shapley_threshold(float u_size, const float* __restrict__ prefix,
                                                   int n, float inv_n) {
  int count = 0;
  for (int k = 0; k < n - 1; ++k) count += (u_size >= prefix[k]) ? 1 : 0;
  const int position = count > 0 ? count - 1 : 0;
  return __fmul_rn(inv_n, (float)position);
}

// Philox4x32-10 (counter-based; Salmon et al. 2011) — the on-device uniform source.
struct Philox {
  uint32_t k0, k1;
  __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi,
             c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// One thread = one packed word of one PAIR; writes the word for row 2i and its complement for row
// 2i+1 (complement restricted to real player bits; CLS bit stays 1 in both).
// MODE 0: uniforms given (bit-exact mirror of the reference given the same draws)
// MODE 1: uniforms drawn on device from Philox(seed, offset)
template <int MODE>
__global__ void shapley_masks_kernel(const float* __restrict__ u_players, const float* __restrict__ u_size,
                                     const float* __restrict__ prefix, uint64_t seed, uint64_t offset,
                                     int pairs, int n, float inv_n, uint32_t* __restrict__ packed,
                                     int words, int64_t* __restrict__ dense) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)pairs * words) return;
  const int pair = (int)(gid / words), w = (int)(gid % words);
  Philox rng(seed);
  float us;
  if (MODE == 0) us = u_size[pair];
  else us = u24(rng(offset + (uint64_t)pair, 0xFFFFFFFFull).x);
  const float thresh = shapley_threshold(us, prefix, n, inv_n);
  uint32_t bits = 0, valid = 0;
  for (int b4 = 0; b4 < 32; b4 += 4) {
    uint4 rnd4 = make_uint4(0, 0, 0, 0);
    if (MODE == 1) rnd4 = rng(offset + (uint64_t)pair, (uint64_t)(w * 8 + (b4 >> 2)));
    const uint32_t rr[4] = {rnd4.x, rnd4.y, rnd4.z, rnd4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b4 + i;
      const int tok = w * 32 + b;
      const int j = tok - 1;
      if (tok == 0) { bits |= 1u; continue; }
      if (j < n) {
        const float u = (MODE == 0) ? u_players[(long long)pair * n + j] : u24(rr[i]);
        const uint32_t keep = (u > thresh) ? 1u : 0u;
        bits |= keep << b;
        valid |= 1u << b;
        if (dense != nullptr) {
          dense[((long long)2 * pair) * n + j] = keep;
          dense[((long long)2 * pair + 1) * n + j] = 1 - (int64_t)keep;
        }
      }
    }
  }
  const uint32_t cls = (w == 0) ? 1u : 0u;
  packed[((long long)2 * pair) * words + w] = bits;
  packed[((long long)2 * pair + 1) * words + w] = ((~bits) & valid) | cls;
}

// mask_purely_uniform (reference models/shapley.py:109-115): keep player j iff u[j] > u_row
template <int MODE>
__global__ void uniform_masks_kernel(const float* __restrict__ u_players, const float* __restrict__ u_row,
                                     uint64_t seed, uint64_t offset, int rows, int n,
                                     uint32_t* __restrict__ packed, int words, int64_t* __restrict__ dense) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)rows * words) return;
  const int row = (int)(gid / words), w = (int)(gid % words);
  Philox rng(seed);
  const float thresh = (MODE == 0) ? u_row[row] : u24(rng(offset + (uint64_t)row, 0xFFFFFFFFull).x);
  uint32_t bits = 0;
  for (int b4 = 0; b4 < 32; b4 += 4) {
    uint4 rnd4 = make_uint4(0, 0, 0, 0);
    if (MODE == 1) rnd4 = rng(offset + (uint64_t)row, (uint64_t)(w * 8 + (b4 >> 2)));
    const uint32_t rr[4] = {rnd4.x, rnd4.y, rnd4.z, rnd4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b4 + i;
      const int tok = w * 32 + b;
      const int j = tok - 1;
      if (tok == 0) { bits |= 1u; continue; }
      if (j < n) {
        const float u = (MODE == 0) ? u_players[(long long)row * n + j] : u24(rr[i]);
        const uint32_t keep = (u > thresh) ? 1u : 0u;
        bits |= keep << b;
        if (dense != nullptr) dense[(long long)row * n + j] = keep;
      }
    }
  }
  packed[gid] = bits;
}

// Rank masks: the on-device form of two reference mask generators that are python loops there.
//   * faithfulness insertion / deletion curves (reference scripts/measure_faithfulness.py:225-251,
//     _get_perturbed_samples): players ranked by attribution, descending; mask_i = base with the top stops[i]
//     players flipped.  Rank rule = np.argsort(a)[::-1] with a stable sort: a_k ranks before a_j iff
//     a_k > a_j, or a_k == a_j and k > j.
//   * fixed-count random masks (reference models/shapley.py:118-128, mask_uniform_selective): scores are
//     random keys (caller-supplied or Philox), one stop = n_masked, base = 1 -> exactly n_masked players are 0.
// One block per score row; O(n^2) rank counting from shared memory (n <= 2048), output rows r*nstops + i.
template <int MODE>
__global__ void __launch_bounds__(256)
rank_masks_kernel(const float* __restrict__ scores, uint64_t seed, uint64_t offset, int n,
                  const int* __restrict__ stops, int nstops, int mask_base, uint32_t* __restrict__ packed, int words,
                  int64_t* __restrict__ dense) {
  extern __shared__ float rk_smem[];
  float* a = rk_smem;                                            // n scores
  unsigned short* rank = reinterpret_cast<unsigned short*>(a + n);   // n ranks
  const int row = blockIdx.x;
  if (MODE == 0) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) a[j] = scores[(long long)row * n + j];
  } else {
    Philox rng(seed);
    for (int j4 = threadIdx.x; j4 * 4 < n; j4 += blockDim.x) {
      const uint4 r4 = rng(offset + (uint64_t)row, (uint64_t)j4);
      const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
      for (int i = 0; i < 4; ++i)
        if (j4 * 4 + i < n) a[j4 * 4 + i] = u24(rr[i]);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float aj = a[j];
    int r = 0;
    for (int k = 0; k < n; ++k) {
      const float ak = a[k];
      r += (ak > aj || (ak == aj && k > j)) ? 1 : 0;
    }
    rank[j] = static_cast<unsigned short>(r);
  }
  __syncthreads();
  const uint32_t base = mask_base ? 1u : 0u;
  for (int e = threadIdx.x; e < nstops * words; e += blockDim.x) {
    const int i = e / words, w = e % words;
    const int stop = stops[i];
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
      const int tok = w * 32 + b, j = tok - 1;
      if (tok == 0) { bits |= 1u; continue; }
      if (j < n) {
        const uint32_t keep = base ^ ((int)rank[j] < stop ? 1u : 0u);
        bits |= keep << b;
        if (dense != nullptr) dense[((long long)row * nstops + i) * n + j] = keep;
      }
    }
    packed[((long long)row * nstops + i) * words + w] = bits;
  }
}

// Masked-token dropping for BERT-style (additive -inf) masks: a masked token has probability 0 as a key in every
// block and the head reads only token 0, so masked tokens cannot influence the surrogate output (SURVEY.md 8a-a8 /
// 8f-3).  The kept tokens of each row are packed back to back; these two kernels build the packing.
//   counts[r]            = number of kept tokens of row r (bits < T)
//   src[cu[r] + k]       = (r / S) * T + (position of the k-th kept token): row of the per-input embedding to gather
__global__ void mask_counts_kernel(const uint32_t* __restrict__ packed, int rows, int words, int T, int* __restrict__ counts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int c = 0;
  for (int w = 0; w < words; ++w) {
    uint32_t bits = packed[(long long)r * words + w];
    const int base = w * 32;
    if (base >= T) bits = 0;
    else if (T - base < 32) bits &= (1u << (T - base)) - 1u;
    c += __popc(bits);
  }
  counts[r] = c;
}

__global__ void packed_token_index_kernel(const uint32_t* __restrict__ packed, int rows, int words, int T, int S,
                                          const int* __restrict__ cu, long long* __restrict__ src) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per row
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  int out = cu[r];
  const long long base_row = (long long)(r / S) * T;
  for (int w = 0; w < words; ++w) {
    uint32_t bits = packed[(long long)r * words + w];
    const int base = w * 32;
    if (base >= T) bits = 0;
    else if (T - base < 32) bits &= (1u << (T - base)) - 1u;
    if ((bits >> lane) & 1u) src[out + __popc(bits & ((1u << lane) - 1u))] = base_row + base + lane;
    out += __popc(bits);
  }
}

static inline int blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }

int pack_masks_i64(const int64_t* mask, int rows, int n, int prepend_cls, uint32_t* packed, int words,
                   cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && n > 0 && words * 32 >= n + (prepend_cls ? 1 : 0), "mask shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(mask && packed, "null pointer");
  pack_masks_kernel<<<blocks_for((long long)rows * words, 256), 256, 0, st>>>(mask, rows, n, prepend_cls ? 1 : 0,
                                                                              packed, words);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int unpack_masks_i64(const uint32_t* packed, int rows, int n, int skip, int words, int64_t* out,
                     cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && n > 0 && words * 32 >= n + skip && skip >= 0, "mask shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(packed && out, "null pointer");
  unpack_masks_kernel<<<blocks_for((long long)rows * n, 256), 256, 0, st>>>(packed, rows, n, skip, words, out);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int shapley_masks(const float* u_players, const float* u_size, const float* prefix, int use_philox,
                  uint64_t seed, uint64_t offset, int pairs, int n, uint32_t* packed, int words,
                  int64_t* dense, cudaStream_t st) {
  AGB_REQUIRE(pairs >= 0 && n >= 2 && words * 32 >= n + 1, "sampler shape (n_players >= 2)");
  if (pairs == 0) return AGB_OK;
  AGB_REQUIRE(prefix && packed, "null pointer");
  const float inv_n = (float)(1.0 / (double)n);
  const int blocks = blocks_for((long long)pairs * words, 128);
  if (use_philox) {
    shapley_masks_kernel<1><<<blocks, 128, 0, st>>>(nullptr, nullptr, prefix, seed, offset, pairs, n, inv_n,
                                                    packed, words, dense);
  } else {
    AGB_REQUIRE(u_players && u_size, "uniforms required");
    shapley_masks_kernel<0><<<blocks, 128, 0, st>>>(u_players, u_size, prefix, 0, 0, pairs, n, inv_n, packed,
                                                    words, dense);
  }
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int uniform_masks(const float* u_players, const float* u_row, int use_philox, uint64_t seed,
                  uint64_t offset, int rows, int n, uint32_t* packed, int words, int64_t* dense,
                  cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && n >= 1 && words * 32 >= n + 1, "sampler shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(packed, "null pointer");
  const int blocks = blocks_for((long long)rows * words, 128);
  if (use_philox) {
    uniform_masks_kernel<1><<<blocks, 128, 0, st>>>(nullptr, nullptr, seed, offset, rows, n, packed, words, dense);
  } else {
    AGB_REQUIRE(u_players && u_row, "uniforms required");
    uniform_masks_kernel<0><<<blocks, 128, 0, st>>>(u_players, u_row, 0, 0, rows, n, packed, words, dense);
  }
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int rank_masks(const float* scores, int use_philox, uint64_t seed, uint64_t offset, int rows, int n, const int* stops,
               int nstops, int mask_base, uint32_t* packed, int words, int64_t* dense, cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && n >= 1 && n <= 2048 && nstops >= 1 && words * 32 >= n + 1, "rank-mask shape (n <= 2048)");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(packed && stops && (use_philox || scores), "null pointer");
  const size_t smem = (size_t)n * 4 + (size_t)n * 2;
  if (use_philox)
    rank_masks_kernel<1><<<rows, 256, smem, st>>>(nullptr, seed, offset, n, stops, nstops, mask_base, packed, words, dense);
  else
    rank_masks_kernel<0><<<rows, 256, smem, st>>>(scores, 0, 0, n, stops, nstops, mask_base, packed, words, dense);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int mask_counts(const uint32_t* packed, int rows, int words, int T, int* counts, cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && words > 0 && T > 0 && words * 32 >= T, "mask shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(packed && counts, "null pointer");
  mask_counts_kernel<<<blocks_for(rows, 128), 128, 0, st>>>(packed, rows, words, T, counts);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int packed_token_index(const uint32_t* packed, int rows, int words, int T, int S, const int* cu, long long* src,
                       cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && words > 0 && T > 0 && S >= 1 && words * 32 >= T, "mask shape");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(packed && cu && src, "null pointer");
  packed_token_index_kernel<<<blocks_for((long long)rows * 32, 256), 256, 0, st>>>(packed, rows, words, T, S, cu, src);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

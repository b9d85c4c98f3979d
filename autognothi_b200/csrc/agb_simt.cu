// fp32 CUDA-core kernels: the "exact" precision mode used to verify the path against the reference
// at fp32 rtol 1e-4 (BASELINE.json north_star, SURVEY.md §7 "fp32 rtol 1e-4 gate"): a register-tiled
// fp32 GEMM with the same fused epilogues as the tensor-core kernel, and a key-masked attention that
// never materialises the (N,h,T,T) score tensor.  These are correctness instruments, not the
// throughput path — the bf16 tcgen05 kernels are — but they run on the GPU with no library calls.
#include "agb_common.cuh"

namespace agb {

// ------------------------------------------------------------------------------------------------
// C[M,N] = act(alpha * A[M,K] * B[N,K]^T + bias) + residual      (all fp32, row-major)
// 64x64 tile, 16x16 threads, 4x4 outputs per thread, K tiles of 16 staged in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, int a_mn, const float* __restrict__ B, int ldb, int b_mn,
                int M, int N, int K, float alpha, const float* __restrict__ bias, int act,
                const float* __restrict__ res, int ldr, float* __restrict__ C, int ldc) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // 64 rows x 16 k per operand = 1024 elements, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * 256;
      const int r = e >> 4, kk = e & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? (a_mn ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk]) : 0.f;
      Bs[kk][r] = (gn < N && gk < K) ? (b_mn ? B[(long long)gk * ldb + gn] : B[(long long)gn * ldb + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] * alpha;
      if (bias) v += bias[gn];
      if (act == 1) v = gelu_erf_exact(v);
      if (res) v += res[(long long)gm * ldr + gn];
      C[(long long)gm * ldc + gn] = v;
    }
  }
}

int gemm_f32(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, int M, int N, int K, float alpha,
             const float* bias, int act, const float* res, int ldr, float* C, int ldc, cudaStream_t st) {
  AGB_REQUIRE(M >= 0 && N > 0 && K > 0, "GEMM shape");
  if (M == 0) return AGB_OK;
  AGB_REQUIRE(A && B && C, "null pointer");
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  AGB_REQUIRE(grid.y <= 65535, "M too large for the fp32 verification GEMM (chunk the rows)");
  gemm_f32_kernel<<<grid, 256, 0, st>>>(A, lda, a_mn, B, ldb, b_mn, M, N, K, alpha, bias, act, res, ldr, C, ldc);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Key-masked attention, CUDA-core version (fp32 math; fp32 or bf16 I/O).
//   qkv : (N, T, 3H) fused projections, q | k | v along the last axis
//   mask: packed key bitmask (N, words); bit t = token t kept
//   mode AGB_MASK_MUL0   : s = m_j ? q.k/sqrt(d) : 0          (reference models/vanilla_vit.py:444-454)
//   mode AGB_MASK_NEGINF : s = q.k/sqrt(d) + (1-m_j)*FLT_MIN  (reference models/vanilla_bert.py:520-523)
// One warp per (row, head, query).  Scores live in a per-warp shared-memory strip of T floats.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

template <typename TIO, int MAXD32>
__global__ void attention_simt_kernel(const TIO* __restrict__ qkv, const uint32_t* __restrict__ mask,
                                      int words, int T, int H, int heads, int mode, TIO* __restrict__ ctx,
                                      unsigned drop_thr, unsigned long long drop_seed, float drop_scale) {
  extern __shared__ float sc_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int d = H / heads;
  const int row = blockIdx.z, head = blockIdx.y;
  const int qi = blockIdx.x * nw + warp;
  if (qi >= T) return;
  float* sc = sc_all + warp * T;
  const long long base = (long long)row * T * 3 * H;
  const TIO* qp = qkv + base + (long long)qi * 3 * H + head * d;
  const uint32_t* mrow = mask + (long long)row * words;
  const float scale = 1.0f / sqrtf((float)d);
  float qreg[MAXD32];
#pragma unroll
  for (int i = 0; i < MAXD32; ++i) qreg[i] = (lane + 32 * i < d) ? ldf<TIO>(qp + lane + 32 * i) : 0.f;
  float mx = -INFINITY;
  for (int j = 0; j < T; ++j) {
    const TIO* kp = qkv + base + (long long)j * 3 * H + H + head * d;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < MAXD32; ++i)
      if (lane + 32 * i < d) a = fmaf(qreg[i], ldf<TIO>(kp + lane + 32 * i), a);
    a = warp_sum(a);
    float s = a / sqrtf((float)d);
    (void)scale;
    const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
    if (mode == AGB_MASK_MUL0) s = keep ? s : 0.f;
    else s = keep ? s : s + (-3.402823466e+38f);
    if (lane == 0) sc[j] = s;
    mx = fmaxf(mx, s);
  }
  __syncwarp();
  float z = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    z += e;
  }
  z = warp_sum(z);
  __syncwarp();
  float o[MAXD32];
#pragma unroll
  for (int i = 0; i < MAXD32; ++i) o[i] = 0.f;
  for (int j = 0; j < T; ++j) {
    float pj = sc[j] / z;
    if (drop_thr)      // training mode: dropout on the probabilities (the row sum above keeps all of them)
      pj = agb_attn_keep(drop_seed, (uint32_t)(row * heads + head), (uint32_t)qi, (uint32_t)j, drop_thr) ? pj * drop_scale : 0.f;
    const TIO* vp = qkv + base + (long long)j * 3 * H + 2 * H + head * d;
#pragma unroll
    for (int i = 0; i < MAXD32; ++i)
      if (lane + 32 * i < d) o[i] = fmaf(pj, ldf<TIO>(vp + lane + 32 * i), o[i]);
  }
  TIO* op = ctx + ((long long)row * T + qi) * H + head * d;
#pragma unroll
  for (int i = 0; i < MAXD32; ++i)
    if (lane + 32 * i < d) stf<TIO>(op + lane + 32 * i, o[i]);
}

// ------------------------------------------------------------------------------------------------
// CLS-query attention for the LAST encoder block of a surrogate / classifier: the head reads only token 0
// (reference models/vanilla_vit.py:51-56 `[:, 0, :]`, models/vanilla_bert.py:615-619 pooler), so in the last block
// only the CLS query row is needed — its keys / values still come from every token.  Exact work-skipping.
//   q  : (rows, ldq)       the CLS query of each row, heads along the columns
//   kv : (rows * T, ldkv)  per-token keys at column k_off + head*d and values at v_off + head*d
// Same arithmetic (order of operations) as attention_simt_kernel for query 0.  One warp per (row, head).
// ------------------------------------------------------------------------------------------------
template <typename TIO, int MAXD32>
__global__ void cls_attention_kernel(const TIO* __restrict__ q, long long ldq, const TIO* __restrict__ kv, long long ldkv,
                                     int k_off, int v_off, const uint32_t* __restrict__ mask, int words, int T, int H,
                                     int heads, int mode, TIO* __restrict__ ctx, long long ldc) {
  extern __shared__ float sc_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int d = H / heads;
  const int row = blockIdx.y, head = blockIdx.x * nw + warp;
  if (head >= heads) return;
  float* sc = sc_all + warp * T;
  const TIO* qp = q + (long long)row * ldq + head * d;
  const TIO* kvr = kv + (long long)row * T * ldkv;
  const uint32_t* mrow = mask + (long long)row * words;
  float qreg[MAXD32];
#pragma unroll
  for (int i = 0; i < MAXD32; ++i) qreg[i] = (lane + 32 * i < d) ? ldf<TIO>(qp + lane + 32 * i) : 0.f;
  float mx = -INFINITY;
  for (int j = 0; j < T; ++j) {
    const TIO* kp = kvr + (long long)j * ldkv + k_off + head * d;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < MAXD32; ++i)
      if (lane + 32 * i < d) a = fmaf(qreg[i], ldf<TIO>(kp + lane + 32 * i), a);
    a = warp_sum(a);
    float s = a / sqrtf((float)d);
    const uint32_t keep = (mrow[j >> 5] >> (j & 31)) & 1u;
    if (mode == AGB_MASK_MUL0) s = keep ? s : 0.f;
    else s = keep ? s : s + (-3.402823466e+38f);
    if (lane == 0) sc[j] = s;
    mx = fmaxf(mx, s);
  }
  __syncwarp();
  float z = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    z += e;
  }
  z = warp_sum(z);
  __syncwarp();
  float o[MAXD32];
#pragma unroll
  for (int i = 0; i < MAXD32; ++i) o[i] = 0.f;
  for (int j = 0; j < T; ++j) {
    const float pj = sc[j] / z;
    const TIO* vp = kvr + (long long)j * ldkv + v_off + head * d;
#pragma unroll
    for (int i = 0; i < MAXD32; ++i)
      if (lane + 32 * i < d) o[i] = fmaf(pj, ldf<TIO>(vp + lane + 32 * i), o[i]);
  }
  TIO* op = ctx + (long long)row * ldc + head * d;
#pragma unroll
  for (int i = 0; i < MAXD32; ++i)
    if (lane + 32 * i < d) stf<TIO>(op + lane + 32 * i, o[i]);
}

// bf16, head dim 64 fast path: lane = key for the logits (each lane reads whole 128-byte K rows with 16-byte loads, all
// independent -> memory-level parallelism), lane = 2 output dims for P V (coalesced 128-byte V rows).  fp32 math.
__global__ void __launch_bounds__(128)
cls_attention_d64_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ kv, long long ldkv, int k_off,
                         int v_off, const uint32_t* __restrict__ mask, int words, int T, int heads, int mode,
                         bf16* __restrict__ ctx, long long ldc, const int* __restrict__ cu) {
  extern __shared__ float sc_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.y, head = blockIdx.x * nw + warp;
  if (head >= heads) return;
  float* sq = sc_all + warp * (64 + ((T + 3) & ~3));   // [64] the query (16-byte aligned), then [T] probabilities
  float* sc = sq + 64;
  const bf16* qp = q + (long long)row * ldq + head * 64;
  // packed rows (cu != nullptr): this row's tokens are kv rows [cu[row], cu[row+1]) and every one is a live key
  const bf16* kvr = kv + (cu ? (long long)cu[row] : (long long)row * T) * ldkv;
  const uint32_t* mrow = cu ? nullptr : mask + (long long)row * words;
  if (cu) T = cu[row + 1] - cu[row];
  {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(qp + 2 * lane);
    sq[2 * lane] = bf16_lo(w);
    sq[2 * lane + 1] = bf16_hi(w);
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < T; j += 32) {
    const uint4* kp = reinterpret_cast<const uint4*>(kvr + (long long)j * ldkv + k_off + head * 64);
    uint4 kk[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) kk[c] = __ldg(kp + c);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 q0 = *reinterpret_cast<const float4*>(sq + 8 * c);
      const float4 q1 = *reinterpret_cast<const float4*>(sq + 8 * c + 4);
      a = fmaf(q0.x, bf16_lo(kk[c].x), a); a = fmaf(q0.y, bf16_hi(kk[c].x), a);
      a = fmaf(q0.z, bf16_lo(kk[c].y), a); a = fmaf(q0.w, bf16_hi(kk[c].y), a);
      a = fmaf(q1.x, bf16_lo(kk[c].z), a); a = fmaf(q1.y, bf16_hi(kk[c].z), a);
      a = fmaf(q1.z, bf16_lo(kk[c].w), a); a = fmaf(q1.w, bf16_hi(kk[c].w), a);
    }
    float sv = a * 0.125f;
    const uint32_t keep = mrow ? ((__ldg(mrow + (j >> 5)) >> (j & 31)) & 1u) : 1u;
    if (mode == AGB_MASK_MUL0) sv = keep ? sv : 0.f;
    else sv = keep ? sv : -INFINITY;
    sc[j] = sv;
    mx = fmaxf(mx, sv);
  }
  mx = warp_max(mx);
  float z = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    z += e;
  }
  z = warp_sum(z);
  __syncwarp();
  const float inv = 1.0f / z;
  float o0 = 0.f, o1 = 0.f;
  const bf16* vp = kvr + v_off + head * 64 + 2 * lane;
#pragma unroll 4
  for (int j = 0; j < T; ++j) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(vp + (long long)j * ldkv));
    const float pj = sc[j];
    o0 = fmaf(pj, bf16_lo(w), o0);
    o1 = fmaf(pj, bf16_hi(w), o1);
  }
  *reinterpret_cast<uint32_t*>(ctx + (long long)row * ldc + head * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
}

int cls_attention(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off, int io_bf16,
                  const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode, void* ctx, long long ldc,
                  cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && T > 0 && heads > 0 && H % heads == 0, "attention shape");
  AGB_REQUIRE(words * 32 >= T, "mask words");
  AGB_REQUIRE(mode == AGB_MASK_MUL0 || mode == AGB_MASK_NEGINF, "mask mode");
  AGB_REQUIRE(H / heads <= 128, "head dim <= 128");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(q && kv && mask && ctx, "null pointer");
  AGB_REQUIRE(rows <= 65535, "grid limits (chunk the rows)");
  const int nw = 4;
  dim3 grid((heads + nw - 1) / nw, rows);
  if (io_bf16 && H == heads * 64 && (ldq % 2) == 0 && (ldkv % 8) == 0 && (k_off % 8) == 0 && (v_off % 8) == 0 &&
      (ldc % 2) == 0 && (reinterpret_cast<uintptr_t>(q) & 3) == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(ctx) & 3) == 0) {
    const size_t smem64 = (size_t)nw * (64 + ((T + 3) & ~3)) * sizeof(float);
    cls_attention_d64_kernel<<<grid, nw * 32, smem64, st>>>(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(kv), ldkv,
                                                            k_off, v_off, mask, words, T, heads, mode,
                                                            static_cast<bf16*>(ctx), ldc, nullptr);
    AGB_CHECK_CUDA(cudaGetLastError());
    return AGB_OK;
  }
  const size_t smem = (size_t)nw * T * sizeof(float);
  if (io_bf16)
    cls_attention_kernel<bf16, 4><<<grid, nw * 32, smem, st>>>(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(kv),
                                                              ldkv, k_off, v_off, mask, words, T, H, heads, mode,
                                                              static_cast<bf16*>(ctx), ldc);
  else
    cls_attention_kernel<float, 4><<<grid, nw * 32, smem, st>>>(static_cast<const float*>(q), ldq,
                                                               static_cast<const float*>(kv), ldkv, k_off, v_off, mask, words,
                                                               T, H, heads, mode, static_cast<float*>(ctx), ldc);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// ------------------------------------------------------------------------------------------------
// Narrow heads (head dim 8 / 16 / 32, bf16 I/O): the side ladders of the LTT variants (reference models/ltt_vit.py:386-396,
// hidden size s_attn_hidden_size split over the backbone's head count) evaluated on every coalition row.  One CTA per
// (row, head): K and V staged once in shared memory as fp32, every thread owns TWO queries and streams the keys with an
// online softmax (exp2 domain), so each broadcast K/V read feeds 4 D FMAs.  fp32 math throughout; P never leaves registers.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
attention_narrow_kernel(const bf16* __restrict__ qkv, const uint32_t* __restrict__ mask, int words, int T, int H, int heads,
                        int mode, float scale_log2, bf16* __restrict__ ctx, unsigned drop_thr, unsigned long long drop_seed,
                        float drop_scale, const int* __restrict__ cu) {
  // cu != nullptr: packed variable-length rows (masked-token dropping): row r owns the tokens [cu[r], cu[r+1]) of
  // qkv (total, 3H) / ctx (total, H), every packed token is a live key; T is then the capacity (max row length)
  extern __shared__ __align__(16) float sm_n[];
  const int Tcap = T;
  int tok0 = 0;
  if (cu != nullptr) {
    tok0 = __ldg(cu + blockIdx.x / heads);
    T = __ldg(cu + blockIdx.x / heads + 1) - tok0;
  }
  float* sK = sm_n;            // Tcap x D
  float* sV = sK + Tcap * D;   // Tcap x D
  float* sVm = sV + Tcap * D;  // D: sum of the V rows of the ViT-masked keys
  int* sIdx = reinterpret_cast<int*>(sVm + D);   // Tcap: indices of the kept keys
  int* sCnt = sIdx + Tcap;
  uint32_t* sM = reinterpret_cast<uint32_t*>(sCnt + 1);
  const int row = blockIdx.x / heads, head = blockIdx.x % heads;
  const long long first = cu != nullptr ? (long long)tok0 : (long long)row * T;      // first token of this row
  const bf16* base = qkv + first * 3 * H + head * D;
  constexpr int V8 = D / 8;
  for (int e = threadIdx.x; e < T * V8; e += blockDim.x) {
    const int t = e / V8, v = e % V8;
    const uint4 k4 = *reinterpret_cast<const uint4*>(base + (long long)t * 3 * H + H + v * 8);
    const uint4 v4 = *reinterpret_cast<const uint4*>(base + (long long)t * 3 * H + 2 * H + v * 8);
    const uint32_t kw[4] = {k4.x, k4.y, k4.z, k4.w}, vw[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sK[t * D + v * 8 + 2 * u] = __uint_as_float(kw[u] << 16);
      sK[t * D + v * 8 + 2 * u + 1] = __uint_as_float(kw[u] & 0xFFFF0000u);
      sV[t * D + v * 8 + 2 * u] = __uint_as_float(vw[u] << 16);
      sV[t * D + v * 8 + 2 * u + 1] = __uint_as_float(vw[u] & 0xFFFF0000u);
    }
  }
  for (int e = threadIdx.x; e < words; e += blockDim.x) sM[e] = cu != nullptr ? 0xFFFFFFFFu : mask[(long long)row * words + e];
  for (int e = threadIdx.x; e < D; e += blockDim.x) sVm[e] = 0.f;
  __syncthreads();
  // kept keys compacted (one warp, ballot prefix); ViT-masked keys all carry the logit 0, so their V rows are summed
  // once per (row, head) and enter the softmax as ONE virtual key of weight n_masked — exact, and it halves the key
  // loop at the sampler's average coalition size
  if (threadIdx.x < 32) {
    int cnt = 0;
    for (int j0 = 0; j0 < T; j0 += 32) {
      const int j = j0 + threadIdx.x;
      // with attention dropout every key is visited on its own (no folding of the ViT-masked ones): their K rows
      // are zeroed below instead, which gives the same logit 0
      const bool keep = j < T && (((sM[j >> 5] >> (j & 31)) & 1u) || (drop_thr && mode == AGB_MASK_MUL0));
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      if (keep) sIdx[cnt + __popc(bal & ((1u << threadIdx.x) - 1u))] = j;
      cnt += __popc(bal);
    }
    if (threadIdx.x == 0) *sCnt = cnt;
  }
  if (mode == AGB_MASK_MUL0) {
    const int c = threadIdx.x % D, part = threadIdx.x / D, parts = blockDim.x / D;
    float acc = 0.f;
    for (int j = part; j < T; j += parts)
      if (!((sM[j >> 5] >> (j & 31)) & 1u)) {
        acc += sV[j * D + c];
        if (drop_thr) sK[j * D + c] = 0.f;
      }
    atomicAdd(&sVm[c], acc);
  }
  __syncthreads();
  const int nkept = *sCnt;
  const int nmasked = (mode == AGB_MASK_MUL0 && !drop_thr) ? T - nkept : 0;
  const int half = (T + 1) / 2;
  for (int i0 = threadIdx.x; i0 < half; i0 += blockDim.x) {
    const int i1 = i0 + half;                 // second query of this thread (may fall off the end)
    const bool has1 = i1 < T;
    float q0[D], q1[D], o0[D], o1[D];
#pragma unroll
    for (int v = 0; v < V8; ++v) {
      const uint4 a = *reinterpret_cast<const uint4*>(base + (long long)i0 * 3 * H + v * 8);
      const uint4 b = has1 ? *reinterpret_cast<const uint4*>(base + (long long)i1 * 3 * H + v * 8) : make_uint4(0, 0, 0, 0);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        q0[v * 8 + 2 * u] = __uint_as_float(aw[u] << 16) * scale_log2;
        q0[v * 8 + 2 * u + 1] = __uint_as_float(aw[u] & 0xFFFF0000u) * scale_log2;
        q1[v * 8 + 2 * u] = __uint_as_float(bw[u] << 16) * scale_log2;
        q1[v * 8 + 2 * u + 1] = __uint_as_float(bw[u] & 0xFFFF0000u) * scale_log2;
      }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) o0[k] = o1[k] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int jj = 0; jj < nkept; ++jj) {
      const int j = sIdx[jj];
      float s0 = 0.f, s1 = 0.f;
      const float4* kr = reinterpret_cast<const float4*>(sK + j * D);
#pragma unroll
      for (int v = 0; v < D / 4; ++v) {
        const float4 kk = kr[v];
        s0 = fmaf(q0[4 * v], kk.x, s0); s0 = fmaf(q0[4 * v + 1], kk.y, s0);
        s0 = fmaf(q0[4 * v + 2], kk.z, s0); s0 = fmaf(q0[4 * v + 3], kk.w, s0);
        s1 = fmaf(q1[4 * v], kk.x, s1); s1 = fmaf(q1[4 * v + 1], kk.y, s1);
        s1 = fmaf(q1[4 * v + 2], kk.z, s1); s1 = fmaf(q1[4 * v + 3], kk.w, s1);
      }
      if (s0 > m0) {
        const float c = exp2f(m0 - s0);
        l0 *= c;
#pragma unroll
        for (int k = 0; k < D; ++k) o0[k] *= c;
        m0 = s0;
      }
      if (s1 > m1) {
        const float c = exp2f(m1 - s1);
        l1 *= c;
#pragma unroll
        for (int k = 0; k < D; ++k) o1[k] *= c;
        m1 = s1;
      }
      float p0 = exp2f(s0 - m0), p1 = exp2f(s1 - m1);
      l0 += p0;
      l1 += p1;
      if (drop_thr) {      // the row sums keep every probability; dropped ones only leave the P V product
        if (!agb_attn_keep(drop_seed, blockIdx.x, i0, j, drop_thr)) p0 = 0.f;
        if (!agb_attn_keep(drop_seed, blockIdx.x, i1, j, drop_thr)) p1 = 0.f;
      }
      const float4* vr = reinterpret_cast<const float4*>(sV + j * D);
#pragma unroll
      for (int v = 0; v < D / 4; ++v) {
        const float4 vv = vr[v];
        o0[4 * v] = fmaf(p0, vv.x, o0[4 * v]); o0[4 * v + 1] = fmaf(p0, vv.y, o0[4 * v + 1]);
        o0[4 * v + 2] = fmaf(p0, vv.z, o0[4 * v + 2]); o0[4 * v + 3] = fmaf(p0, vv.w, o0[4 * v + 3]);
        o1[4 * v] = fmaf(p1, vv.x, o1[4 * v]); o1[4 * v + 1] = fmaf(p1, vv.y, o1[4 * v + 1]);
        o1[4 * v + 2] = fmaf(p1, vv.z, o1[4 * v + 2]); o1[4 * v + 3] = fmaf(p1, vv.w, o1[4 * v + 3]);
      }
    }
    if (nmasked > 0) {                         // the ViT-masked keys: logit 0, weight n_masked, summed V rows
      if (0.f > m0) {
        const float c = exp2f(m0);
        l0 *= c;
#pragma unroll
        for (int k = 0; k < D; ++k) o0[k] *= c;
        m0 = 0.f;
      }
      if (0.f > m1) {
        const float c = exp2f(m1);
        l1 *= c;
#pragma unroll
        for (int k = 0; k < D; ++k) o1[k] *= c;
        m1 = 0.f;
      }
      const float p0 = exp2f(-m0), p1 = exp2f(-m1);
      l0 = fmaf(p0, (float)nmasked, l0);
      l1 = fmaf(p1, (float)nmasked, l1);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        o0[k] = fmaf(p0, sVm[k], o0[k]);
        o1[k] = fmaf(p1, sVm[k], o1[k]);
      }
    }
    const float r0 = l0 > 0.f ? drop_scale / l0 : 0.f, r1 = l1 > 0.f ? drop_scale / l1 : 0.f;
    bf16* c0 = ctx + (first + i0) * H + head * D;
    bf16* c1 = ctx + (first + i1) * H + head * D;
#pragma unroll
    for (int v = 0; v < V8; ++v) {
      uint32_t w0[4], w1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(o0[v * 8 + 2 * u] * r0, o0[v * 8 + 2 * u + 1] * r0);
        const __nv_bfloat162 b = __floats2bfloat162_rn(o1[v * 8 + 2 * u] * r1, o1[v * 8 + 2 * u + 1] * r1);
        w0[u] = *reinterpret_cast<const uint32_t*>(&a);
        w1[u] = *reinterpret_cast<const uint32_t*>(&b);
      }
      *reinterpret_cast<uint4*>(c0 + v * 8) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
      if (has1) *reinterpret_cast<uint4*>(c1 + v * 8) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    }
  }
}

int attention_narrow_mma(const bf16* qkv, const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode, bf16* ctx,
                         cudaStream_t st, const int* cu);

static bool narrow_simt_forced() {
  static const bool forced = [] {
    const char* e = getenv("AGB_NARROW_SIMT");
    return e != nullptr && e[0] != '\0' && e[0] != '0';
  }();
  return forced;
}

template <int D>
static int launch_attention_narrow(const bf16* qkv, const uint32_t* mask, int words, int rows, int T, int H, int heads,
                                   int mode, bf16* ctx, cudaStream_t st, unsigned drop_thr, unsigned long long drop_seed,
                                   const int* cu = nullptr) {
  // no dropout (every inference call): the warp-level tensor-core kernel (agb_attention_narrow.cu);
  // AGB_NARROW_SIMT=1 keeps the CUDA-core kernel for A/B runs
  if (drop_thr == 0 && !narrow_simt_forced()) {
    const int rc = attention_narrow_mma(qkv, mask, words, rows, T, H, heads, mode, ctx, st, cu);
    if (rc != AGB_ERR_UNSUPPORTED) return rc;
  }
  if (cu != nullptr) words = (T + 31) / 32;     // all-ones key bits, built in shared memory
  const size_t smem = ((size_t)2 * T * D + D + T + 1) * sizeof(float) + (size_t)words * sizeof(uint32_t);
  if (smem > 227 * 1024) return AGB_ERR_UNSUPPORTED;
  AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_narrow_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = min(128, (((T + 1) / 2 + 31) / 32) * 32);
  attention_narrow_kernel<D><<<rows * heads, threads, smem, st>>>(qkv, mask, words, T, H, heads, mode,
                                                                   rsqrtf((float)D) * 1.4426950408889634f, ctx, drop_thr,
                                                                   drop_seed, 65536.0f / (65536.0f - (float)drop_thr), cu);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

// packed variable-length rows with narrow heads (the LTT ladder after masked-token dropping)
int attention_narrow_varlen(const bf16* qkv, const int* cu, int rows, int max_len, int H, int heads, bf16* ctx, cudaStream_t st) {
  const int d = H / heads;
  AGB_REQUIRE(rows >= 0 && max_len > 0 && heads > 0 && H == heads * d && (d == 8 || d == 16 || d == 32) && (H % 8) == 0,
              "narrow varlen attention shape (head dim 8 / 16 / 32)");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(qkv && cu && ctx, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0, "alignment");
  return d == 8 ? launch_attention_narrow<8>(qkv, nullptr, 0, rows, max_len, H, heads, AGB_MASK_NEGINF, ctx, st, 0, 0, cu)
       : d == 16 ? launch_attention_narrow<16>(qkv, nullptr, 0, rows, max_len, H, heads, AGB_MASK_NEGINF, ctx, st, 0, 0, cu)
                 : launch_attention_narrow<32>(qkv, nullptr, 0, rows, max_len, H, heads, AGB_MASK_NEGINF, ctx, st, 0, 0, cu);
}

int attention_simt(const void* qkv, int io_bf16, const uint32_t* mask, int words, int rows, int T, int H,
                   int heads, int mode, void* ctx, cudaStream_t st, unsigned drop_thr, unsigned long long drop_seed) {
  AGB_REQUIRE(rows >= 0 && T > 0 && heads > 0 && H % heads == 0, "attention shape");
  AGB_REQUIRE(words * 32 >= T, "mask words");
  AGB_REQUIRE(mode == AGB_MASK_MUL0 || mode == AGB_MASK_NEGINF, "mask mode");
  const int d = H / heads;
  AGB_REQUIRE(d <= 128, "head dim <= 128");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(qkv && mask && ctx, "null pointer");
  if (io_bf16 && (d == 8 || d == 16 || d == 32) && (H % 8) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(ctx) & 15) == 0) {
    const bf16* q16 = static_cast<const bf16*>(qkv);
    bf16* c16 = static_cast<bf16*>(ctx);
    const int rc = d == 8 ? launch_attention_narrow<8>(q16, mask, words, rows, T, H, heads, mode, c16, st, drop_thr, drop_seed)
                 : d == 16 ? launch_attention_narrow<16>(q16, mask, words, rows, T, H, heads, mode, c16, st, drop_thr, drop_seed)
                           : launch_attention_narrow<32>(q16, mask, words, rows, T, H, heads, mode, c16, st, drop_thr, drop_seed);
    if (rc != AGB_ERR_UNSUPPORTED) return rc;
  }
  AGB_REQUIRE(rows <= 65535 && heads <= 65535, "grid limits (chunk the rows)");
  const int nw = 4;
  dim3 grid((T + nw - 1) / nw, heads, rows);
  const size_t smem = (size_t)nw * T * sizeof(float);
  if (io_bf16)
    attention_simt_kernel<bf16, 4><<<grid, nw * 32, smem, st>>>(static_cast<const bf16*>(qkv), mask, words, T, H,
                                                                heads, mode, static_cast<bf16*>(ctx), drop_thr, drop_seed,
                                                                65536.0f / (65536.0f - (float)drop_thr));
  else
    attention_simt_kernel<float, 4><<<grid, nw * 32, smem, st>>>(static_cast<const float*>(qkv), mask, words, T, H,
                                                                 heads, mode, static_cast<float*>(ctx), drop_thr, drop_seed,
                                                                 65536.0f / (65536.0f - (float)drop_thr));
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int cls_attention_varlen(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off, int io_bf16,
                         const int* cu, int rows, int max_len, int H, int heads, void* ctx, long long ldc, cudaStream_t st) {
  AGB_REQUIRE(rows >= 0 && max_len > 0 && heads > 0, "attention shape");
  AGB_REQUIRE(io_bf16 && H == heads * 64, "packed CLS attention: bf16, head dim 64");
  AGB_REQUIRE((ldq % 2) == 0 && (ldkv % 8) == 0 && (k_off % 8) == 0 && (v_off % 8) == 0 && (ldc % 2) == 0 &&
                  (reinterpret_cast<uintptr_t>(kv) & 15) == 0, "alignment");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(q && kv && cu && ctx, "null pointer");
  AGB_REQUIRE(rows <= 65535, "grid limits (chunk the rows)");
  const int nw = 4;
  dim3 grid((heads + nw - 1) / nw, rows);
  const size_t smem64 = (size_t)nw * (64 + ((max_len + 3) & ~3)) * sizeof(float);
  cls_attention_d64_kernel<<<grid, nw * 32, smem64, st>>>(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(kv), ldkv,
                                                          k_off, v_off, nullptr, 0, max_len, heads, AGB_MASK_NEGINF,
                                                          static_cast<bf16*>(ctx), ldc, cu);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

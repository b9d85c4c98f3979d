// Fused key-masked attention, third generation: ONE group of softmax warps that never waits for the tensor core.
//
// Same contract as agb_attention_pipe.cu for the ViT mask semantics (reference models/vanilla_vit.py:444-463: the logit
// of a masked key is 0, "scores * mask"; bf16 in/out, head dim 64) at T <= 208 — the (N,h,T,T) score tensor never leaves
// the SM and masked copies of the input are never built.
//
// What the second generation left on the table (profiles/r01_attention_pipe_trace_vit.txt, r02_attention_split_ncu.txt):
// its two softmax groups each owned one TMEM region, so a group sat idle from "P written" until the P V MMA, its own
// epilogue AND the next S MMA into the same region had finished (~40 % of its time), it read S from TMEM twice (row max,
// then exp), and one thread walked all 208 key columns of a row.  XU (MUFU) pipe 31 % busy, tensor pipe 17 %.
//
// Per SM, one persistent CTA of 24 warps over "items" (unit = (row, head), m-tile); item k lives in TMEM region k & 1:
//   (24 warps)
//   warp 0       TMA producer: K/V of a unit into a 3-deep ring, Q tiles into a 2-deep ring
//   warp 1       tcgen05 issuer  S_k = Q_k K^T  (N trimmed to the live keys), as soon as region k & 1 is drained
//   warp 2       tcgen05 issuer  O_k = P_k V    the moment P_k is complete
//   warp 3       key prep: zero the K rows of masked keys (logit exactly 0) or, in kept-first order, build the V row of
//                the virtual key that stands for all masked keys
//   warps 4-7    epilogue: O_k / rowsum -> bf16 ctx (thread = query row), then release the region for S_{k+2}
//   warps 8-23   softmax: two groups of 8 warps, group g serves the items of parity g (its own TMEM region), so that one
//                group exponentiates while the other waits for  P V (k) -> epilogue (k) -> S (k+2)  on its region: the
//                XU (MUFU) pipe, 16 exp2 / clk / SM and the floor of this kernel, always has a group feeding it.  Inside a
//                group two warps per TMEM lane quarter split the key columns (half 0 overlays its P on the S columns it
//                has consumed, half 1 writes its P into the region's spare columns [208, 256)).  S is read ONCE: the
//                row's reference maximum is the maximum over the first chunk of both halves and is raised later only if a
//                chunk exceeds it by 2^24 (softmax is shift-invariant; the P already written is then rescaled in place by
//                an exact power of two).
// High-O layout (AGB_ATTN_DEBUG=16, OFF by default — measured 5 % slower): an item whose key columns fit in 192 TMEM columns (kept-first order, >= 16 masked keys: ~80 % of the rows)
// keeps its O accumulator at [192, 256) and the upper half's P on that half's own consumed S columns, so nothing of the
// item lives in [0, 192) once P V has completed: S of item k + 2 — if it also fits in 192 columns — is issued as soon as
// P V (k) is done and overlaps the epilogue of item k instead of waiting for it.
// Kept-first order (AP "prefix" mode, see agb_attention_pipe.cu): keys [0, nkeep) are live, the masked rest is ONE virtual
// key at column nkeep (logit log(n_masked) / scale, V row = mean of the masked V rows); S, the softmax and P V then only
// span round16(nkeep + 1) key columns.
#include <stdlib.h>

#include "agb_common.cuh"

namespace agb {

constexpr int A3_D = 64;
constexpr int A3_THREADS = 768;     // 4 service + 4 epilogue + 16 softmax warps
constexpr int A3_TMEM_COLS = 512;
constexpr int A3_REGION = 256;     // TMEM columns per region (item parity)
constexpr int A3_O_COL = 128;      // O accumulator at [128, 192) of the region (S columns consumed by then)
constexpr int A3_SPARE = 208;      // P of the upper key half at [208, 256)
constexpr int A3_O_HI = 192;       // "high-O" layout (kept-first items with <= 192 key columns): O at [192, 256), see below
constexpr int A3_KV_RING = 3;
constexpr int A3_MAX_NK = 208;
constexpr float A3_LAZY = 24.f;    // log2 units a chunk may exceed the row's reference maximum before it is raised

struct Att3Params {
  const uint32_t* mask;
  int words;
  int rows, T, H, heads;
  int NK;                 // T rounded up to 16
  int units, mtiles, share;
  int kvb;                // bytes of one K (or V) tile = NK * 128
  const int* nkeep;       // PREFIX: kept tokens of row r (CLS included) are its first nkeep[r] tokens
  const uint8_t* dst_pos; // optional (rows, T): query token t of row r is written to token position dst_pos[r, t]
  bf16* ctx;
  long long* trace;
  int debug;              // diagnostics (AGB_ATTN_DEBUG): 1 = softmax touches only its first chunk, 2 = no MMAs are issued,
                          // 4 = exact two-pass row maximum instead of the lazily raised reference maximum, 8 = spin-wait hand-offs,
                          // 16 = high-O layout (A/B: measured SLOWER, 403 vs 382 us kept-first on one box — off by default)
};

#define A3_TRACE(k, slot)                                                                                         \
  do {                                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && (k) < 64 && lane == 0) p.trace[(k) * 16 + (slot)] = clock64();    \
  } while (0)

__device__ __forceinline__ void a3_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int W>
__device__ __forceinline__ float a3_max(const uint32_t (&s)[W], float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < W; j += 4) {
    m0 = fmaxf(m0, __uint_as_float(s[j]));
    m1 = fmaxf(m1, __uint_as_float(s[j + 1]));
    m2 = fmaxf(m2, __uint_as_float(s[j + 2]));
    m3 = fmaxf(m3, __uint_as_float(s[j + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

template <int W>
__device__ __forceinline__ void a3_exp(const uint32_t (&s)[W], uint32_t (&pk)[W / 2], float scale_log2, float m_scaled,
                                       float (&sum)[4]) {
#pragma unroll
  for (int j = 0; j < W / 2; ++j) {
    const float e0 = ex2_approx(fmaf(__uint_as_float(s[2 * j]), scale_log2, -m_scaled));
    const float e1 = ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), scale_log2, -m_scaled));
    sum[(j & 1) * 2] += e0;
    sum[(j & 1) * 2 + 1] += e1;
    pk[j] = pack_bf16x2(e0, e1);
  }
}

template <bool PREFIX>
__global__ void __launch_bounds__(A3_THREADS, 1)
attention_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const Att3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int kvb = p.kvb;
  uint8_t* sQ = smem;                           // [2][128 x 128 B]
  uint8_t* sKV = smem + 2 * 16384;              // [ring][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + A3_KV_RING * 2 * kvb);
  uint64_t* q_full = bars;                      // [2]
  uint64_t* q_empty = bars + 2;                 // [2]
  uint64_t* kv_full = bars + 4;                 // [3]
  uint64_t* kv_prep = bars + 7;                 // [3]
  uint64_t* kv_empty = bars + 10;               // [3]
  uint64_t* s_full = bars + 13;                 // [2]
  uint64_t* p_full = bars + 15;                 // [2]
  uint64_t* o_full = bars + 17;                 // [2]
  uint64_t* o_free = bars + 19;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  float* xch = reinterpret_cast<float*>(bars + 24);   // [2 kinds][2 groups][4 quarters][2 halves][32 lanes] exchange between the halves
  float* rowsum = xch + 2 * 512;                       // [2 regions][2 halves][128 rows]

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&q_full[i]), 1);
      mbar_init(smem_u32(&q_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_full[i]), 8);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&o_free[i]), 4);
    }
    for (int i = 0; i < A3_KV_RING; ++i) {
      mbar_init(smem_u32(&kv_full[i]), 1);
      mbar_init(smem_u32(&kv_prep[i]), 32);
      mbar_init(smem_u32(&kv_empty[i]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), A3_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int H = p.H, T = p.T, NK = p.NK, mt = p.mtiles;
  const int grid = gridDim.x;
  const int nu = (p.units - (int)blockIdx.x + grid - 1) / grid;   // units of this CTA
  const int n_items = nu * mt;
  constexpr int R = A3_KV_RING;

  // key columns of a unit: all NK, or (PREFIX) the kept keys + one virtual key, rounded up to 16
  auto unit_keys = [&](int ui) -> int {
    if (!PREFIX) return NK;
    const int u = blockIdx.x + ui * grid;
    const int nk = __ldg(p.nkeep + u / p.heads);
    return (nk + (nk < T ? 1 : 0) + 15) & ~15;
  };

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    const uint32_t e = elect_one();
    for (int k = 0; k < n_items; ++k) {
      const int ui = k / mt, m = k - ui * mt;
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads, head = u - row * p.heads;
      const int brow = row / p.share;
      if (m == 0) {
        const int b = ui % R, n = ui / R;
        if (n > 0) mbar_wait(smem_u32(&kv_empty[b]), (n - 1) & 1);
        const uint32_t bar = smem_u32(&kv_full[b]);
        mbar_arrive_expect_tx_e(e, bar, 2 * kvb);
        const uint32_t dst = smem_u32(sKV + b * 2 * kvb);
        tma_load_3d_e(e, dst, &tmKV, bar, H + head * A3_D, 0, brow);
        tma_load_3d_e(e, dst + kvb, &tmKV, bar, 2 * H + head * A3_D, 0, brow);
        A3_TRACE(k, 0);
      }
      const int qb = k & 1, nq = k >> 1;
      if (nq > 0) mbar_wait(smem_u32(&q_empty[qb]), (nq - 1) & 1);
      const uint32_t qbar = smem_u32(&q_full[qb]);
      mbar_arrive_expect_tx_e(e, qbar, 16384);
      tma_load_3d_e(e, smem_u32(sQ + qb * 16384), &tmQ, qbar, head * A3_D, m * 128, brow);
    }
  } else if (warp == 1) {
    // ------------------------------ S issuer ------------------------------
    const uint32_t e = elect_one();
    const uint64_t dk0 = make_smem_desc_sw128(0, 16, 1024);       // K-major operands (Q, K)
    const uint32_t sq0 = smem_u32(sQ) >> 4, skv0 = smem_u32(sKV) >> 4;
    const uint32_t kvb16 = (uint32_t)kvb >> 4;
    uint32_t idesc = make_idesc_bf16(128, NK, 0, 0);
    for (int k = 0; k < n_items; ++k) {
      const int ui = k / mt, m = k - ui * mt;
      const int b = ui % R, g = k & 1, n = k >> 1, qb = k & 1, nq = k >> 1;
      if (m == 0) {
        if (PREFIX) idesc = make_idesc_bf16(128, unit_keys(ui), 0, 0);
        mbar_wait(smem_u32(&kv_prep[b]), (ui / R) & 1);
      }
      mbar_wait(smem_u32(&q_full[qb]), nq & 1);
      if (n > 0) {
        // region g still holds item k - 2.  If both that item and this one use the high-O layout, [0, 192) is free the
        // moment P V (k - 2) has completed (o_full); otherwise wait until its epilogue has drained O (o_free).
        const bool early = PREFIX && (p.debug & 16) && unit_keys(ui) <= A3_O_HI && unit_keys((k - 2) / mt) <= A3_O_HI;
        if (early) mbar_wait(smem_u32(&o_full[g]), (n - 1) & 1);
        else if (p.debug & 8) mbar_wait_spin(smem_u32(&o_free[g]), (n - 1) & 1);
        else mbar_wait(smem_u32(&o_free[g]), (n - 1) & 1);
      }
      tc_fence_after();
      const uint32_t aq = sq0 + qb * (16384 >> 4);
      const uint32_t ak = skv0 + b * 2 * kvb16;
      if (!(p.debug & 2)) {
#pragma unroll
        for (int kk = 0; kk < A3_D / 16; ++kk)
          umma_ss_e<1>(e, tmem_base + g * A3_REGION, dk0 + (aq + kk * 2), dk0 + (ak + kk * 2), idesc, kk != 0 ? 1u : 0u);
      }
      umma_commit_e<1>(e, smem_u32(&s_full[g]));
      umma_commit_e<1>(e, smem_u32(&q_empty[qb]));
      A3_TRACE(k, 1);
    }
  } else if (warp == 2) {
    // ------------------------------ PV issuer ------------------------------
    const uint32_t e = elect_one();
    const uint32_t idesc_o = make_idesc_bf16(128, A3_D, 0, 1);
    const uint64_t dv0 = make_smem_desc_sw128(0, 8192, 1024);     // MN-major V, one 64-column atom (LBO unused)
    const uint32_t skv0 = smem_u32(sKV) >> 4;
    const uint32_t kvb16 = (uint32_t)kvb >> 4;
    int n16 = NK >> 4;
    for (int k = 0; k < n_items; ++k) {
      const int ui = k / mt, m = k - ui * mt;
      const int b = ui % R, g = k & 1, n = k >> 1;
      if (PREFIX && m == 0) n16 = unit_keys(ui) >> 4;
      if (p.debug & 8) mbar_wait_spin(smem_u32(&p_full[g]), n & 1);
      else mbar_wait(smem_u32(&p_full[g]), n & 1);
      A3_TRACE(k, 2);
      tc_fence_after();
      const bool hi_o = PREFIX && (p.debug & 16) && n16 * 16 <= A3_O_HI;
      const uint32_t d_o = tmem_base + g * A3_REGION + (hi_o ? A3_O_HI : A3_O_COL);
      uint64_t dv = dv0 + (skv0 + b * 2 * kvb16 + kvb16);
      // the O columns of this item were last read by the epilogue of item k - 2 (the S issuer may not have waited for it)
      if (n > 0) mbar_wait(smem_u32(&o_free[g]), (n - 1) & 1);
      if (!(p.debug & 2)) {
        // P of the lower key half overlays the start of the region; the upper half's P lives in the spare columns, or
        // (high-O layout) on that half's own consumed S columns
        const int ca = (n16 + 1) >> 1;
        uint32_t a_p = tmem_base + g * A3_REGION;
        int ks = 0;
        for (; ks < ca; ++ks, a_p += 8, dv += (2048 >> 4)) umma_ts_e(e, d_o, a_p, dv, idesc_o, ks != 0 ? 1u : 0u);
        a_p = tmem_base + g * A3_REGION + (hi_o ? ca * 16 : A3_SPARE);
        for (; ks < n16; ++ks, a_p += 8, dv += (2048 >> 4)) umma_ts_e(e, d_o, a_p, dv, idesc_o, 1u);
      }
      umma_commit_e<1>(e, smem_u32(&o_full[g]));
      if (m == mt - 1) umma_commit_e<1>(e, smem_u32(&kv_empty[b]));
      A3_TRACE(k, 3);
    }
  } else if (warp == 3) {
    // ------------------------------ key prep ------------------------------
    const int tid = lane;
    for (int ui = 0; ui < nu; ++ui) {
      const int b = ui % R, n = ui / R;
      mbar_wait(smem_u32(&kv_full[b]), n & 1);
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads;
      if (PREFIX) {
        // V row of the virtual key = mean of the masked V rows [nk, T): lane = (16-byte chunk, row group of 4)
        const int nk = __ldg(p.nkeep + row);
        if (nk < T) {
          const uint8_t* sV = sKV + b * 2 * kvb + kvb;
          const uint32_t chunk = (uint32_t)tid & 7u;
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
          for (int j = nk + (tid >> 3); j < T; j += 4) {
            const uint4 w = *reinterpret_cast<const uint4*>(sV + j * 128 + ((chunk ^ ((uint32_t)j & 7u)) << 4));
            acc[0] += bf16_lo(w.x); acc[1] += bf16_hi(w.x);
            acc[2] += bf16_lo(w.y); acc[3] += bf16_hi(w.y);
            acc[4] += bf16_lo(w.z); acc[5] += bf16_hi(w.z);
            acc[6] += bf16_lo(w.w); acc[7] += bf16_hi(w.w);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
          }
          const float invn = 1.0f / (float)(T - nk);
          __syncwarp();              // every masked row (incl. row nk) has been read before row nk is overwritten
          if (tid < 8) {
            uint4 o;
            o.x = pack_bf16x2(acc[0] * invn, acc[1] * invn);
            o.y = pack_bf16x2(acc[2] * invn, acc[3] * invn);
            o.z = pack_bf16x2(acc[4] * invn, acc[5] * invn);
            o.w = pack_bf16x2(acc[6] * invn, acc[7] * invn);
            *reinterpret_cast<uint4*>(sKV + b * 2 * kvb + kvb + nk * 128 + ((chunk ^ ((uint32_t)nk & 7u)) << 4)) = o;
          }
        }
      } else {
        const uint32_t* mrow = p.mask + (long long)row * p.words;
        uint8_t* sK = sKV + b * 2 * kvb;
        for (int j = tid; j < T; j += 32) {
          if (!((__ldg(mrow + (j >> 5)) >> (j & 31)) & 1u)) {
            uint4* kr = reinterpret_cast<uint4*>(sK + j * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) kr[c] = make_uint4(0, 0, 0, 0);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&kv_prep[b]));
    }
  } else if (warp < 8) {
    // ------------------------------ epilogue: O / rowsum -> bf16 ctx ------------------------------
    const int qd = warp & 3;                        // TMEM lane quarter this warp may touch
    const int r = qd * 32 + lane;                   // query row within the tile = TMEM lane
    for (int k = 0; k < n_items; ++k) {
      const int g = k & 1, n = k >> 1;
      const int ui = k / mt, m = k - ui * mt;
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads, head = u - row * p.heads;
      const bool warp_live = (m * 128 + qd * 32) < T;
      const bool hi_o = PREFIX && (p.debug & 16) && unit_keys(ui) <= A3_O_HI;
      const uint32_t o_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + g * A3_REGION + (hi_o ? A3_O_HI : A3_O_COL);
      if (p.debug & 8) mbar_wait_spin(smem_u32(&o_full[g]), n & 1);
      else mbar_wait(smem_u32(&o_full[g]), n & 1);
      if (qd == 0) A3_TRACE(k, 6);
      tc_fence_after();
      // first 32 output dims -> packed bf16, then the other 32 (keeps the live registers under the 80 of this launch)
      uint32_t o[32], w0[16];
      float inv = 0.f;
      if (warp_live) {
        tmem_ld32(o_addr, o);
        inv = 1.0f / (rowsum[(g * 2 + 0) * 128 + r] + rowsum[(g * 2 + 1) * 128 + r]);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) w0[j] = pack_bf16x2(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
        tmem_ld32(o_addr + 32, o);
        tmem_wait_ld();
      }
      // O and the row sums are in registers: release the region (S of item k + 2 may overwrite it) BEFORE storing
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&o_free[g]));
      if (qd == 0) A3_TRACE(k, 7);
      if (warp_live) {
        const int tq = m * 128 + r;
        if (tq < T) {
          const int tdst = p.dst_pos != nullptr ? (int)__ldg(p.dst_pos + (long long)row * T + tq) : tq;
          bf16* dst = p.ctx + ((long long)row * T + tdst) * H + head * A3_D;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            *reinterpret_cast<uint4*>(dst + 8 * cc) = make_uint4(w0[4 * cc], w0[4 * cc + 1], w0[4 * cc + 2], w0[4 * cc + 3]);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[8 * cc + 0]) * inv, __uint_as_float(o[8 * cc + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[8 * cc + 2]) * inv, __uint_as_float(o[8 * cc + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[8 * cc + 4]) * inv, __uint_as_float(o[8 * cc + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[8 * cc + 6]) * inv, __uint_as_float(o[8 * cc + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + 32 + 8 * cc) = w;
          }
        }
      }
    }
  } else {
    // ------------------------------ softmax ------------------------------
    const int qd = warp & 3;                        // TMEM lane quarter this warp may touch
    const int g = ((warp - 8) >> 2) & 1;            // item parity (= TMEM region) this warp serves
    const int split = (warp - 8) >> 3;              // lower / upper half of the key columns
    const int r = qd * 32 + lane;
    const float scale_log2 = 0.125f * 1.4426950408889634f;
    float* xmax = xch + ((g * 4 + qd) * 2) * 32;    // [2 halves][32 lanes]
    float* xfin = xmax + 512;
    const int bar_id = 1 + g * 4 + qd;
    for (int k = g; k < n_items; k += 2) {
      const int n = k >> 1;
      const int ui = k / mt, m = k - ui * mt;
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads;
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + g * A3_REGION;
      const bool warp_live = (m * 128 + qd * 32) < T;     // warp-uniform, identical for both halves
      int Tk = T, vj = -1;
      float vval = 0.f;
      if (PREFIX) {
        const int nk = __ldg(p.nkeep + row);
        Tk = nk + (nk < T ? 1 : 0);
        if (nk < T) {
          vj = nk & 15;                                    // the virtual key sits in the last 16-column chunk
          vval = __log2f((float)(T - nk)) / scale_log2;   // exp2(vval * scale_log2 - m) = n_masked * exp2(0 - m)
        }
      }
      const int NKu = (Tk + 15) & ~15;
      const int n16 = NKu >> 4;
      // my key columns [c_beg, c_end): the lower half gets ceil(n16 / 2) of the 16-column steps
      const int ca = (n16 + 1) >> 1;
      const int c_beg = split ? ca * 16 : 0, c_end = split ? NKu : ca * 16;
      const int tail = NKu - 16;                           // the only chunk with dead columns / the virtual key
      const int fast_end = c_end < tail ? c_end : tail;
      const uint32_t tail_live = (Tk - tail) >= 16 ? 0xFFFFu : ((1u << (Tk - tail)) - 1u);
      // P (bf16 pairs): the lower half overlays the S columns it has consumed (writes trail reads), the upper half uses the
      // spare columns [208, 256): the O accumulator will overwrite [128, 192), and S columns the lower half still has to read
      // must not be touched
      const bool hi_o = PREFIX && (p.debug & 16) && NKu <= A3_O_HI;        // (high-O layout: see the file header)
      const uint32_t p_start = lane_addr + (split ? (uint32_t)(hi_o ? ca * 16 : A3_SPARE) : 0u);

      mbar_wait(smem_u32(&s_full[g]), n & 1);
      if (qd == 0 && split == 0) A3_TRACE(k, 4);
      if (qd == 0 && split == 1) A3_TRACE(k, 13);
      tc_fence_after();
      if (warp_live) {
        uint32_t s[32];
        // load the chunk at cc: width 32, 16 (upper half padded with -inf) or 0; the tail chunk gets its dead columns -> -inf
        // and the virtual key -> its logit
        auto load_chunk = [&](int cc) -> int {
          if (cc + 32 <= fast_end) {
            tmem_ld32(lane_addr + cc, s);
            tmem_wait_ld();
            return 32;
          }
#pragma unroll
          for (int j = 16; j < 32; ++j) s[j] = 0xff800000u;
          if (cc < c_end) {
            tmem_ld16(lane_addr + cc, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
            tmem_wait_ld();
            if (cc == tail) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (PREFIX && j == vj) s[j] = __float_as_uint(vval);
                if (!((tail_live >> j) & 1u)) s[j] = 0xff800000u;
              }
            }
            return 16;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = 0xff800000u;
          return 0;
        };
        uint32_t p_cur = p_start;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        // multiply the P columns [p_start, p_cur) and the partial sums by f (a power of two per lane; warp-collective)
        auto rescale = [&](float f) {
          tmem_wait_st();
          for (uint32_t a = p_start; a < p_cur; a += 8) {
            uint32_t q[8];
            tmem_ld8(a, q);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = pack_bf16x2(bf16_lo(q[j]) * f, bf16_hi(q[j]) * f);
            tmem_st8(a, q);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) sum[j] *= f;
        };
        int c = c_beg;
        float mysum;
        if (p.debug & 4) {
          // ---- A/B variant: exact row maximum first, one exchange between the halves, then the exponentials.  An unloaded
          // tcgen05.ld costs ~35 cycles per 32 columns (profiles/r02_xu_tmem_microbench.txt), but here it queues behind the
          // other group's MUFU stream in the MIO pipe and the extra pass costs ~2000 cycles per item: not the default ----
          float mx = -INFINITY;
          for (int cc = c_beg; cc < c_end;) {
            const int ww = load_chunk(cc);
            mx = a3_max<32>(s, mx);
            cc += ww;
          }
          if (qd == 0 && split == 0) A3_TRACE(k, 8);
          xmax[split * 32 + lane] = mx;
          a3_bar_sync(bar_id, 64);
          if (qd == 0 && split == 0) A3_TRACE(k, 9);
          const float m_s = fmaxf(xmax[lane], xmax[32 + lane]) * scale_log2;
          while (c < c_end) {
            const int w = load_chunk(c);
            if (w == 32) {
              uint32_t pk[16];
              a3_exp<32>(s, pk, scale_log2, m_s, sum);
              tmem_st16(p_cur, pk);
              p_cur += 16;
            } else {
              uint32_t pk[8];
              a3_exp<16>(*reinterpret_cast<uint32_t(*)[16]>(&s[0]), pk, scale_log2, m_s, sum);
              tmem_st8(p_cur, pk);
              p_cur += 8;
            }
            c += w;
            if (p.debug & 1) break;
          }
          if (qd == 0 && split == 0) A3_TRACE(k, 10);
          mysum = (sum[0] + sum[1]) + (sum[2] + sum[3]);
        } else {
        int w = load_chunk(c);
        if (qd == 0 && split == 0) A3_TRACE(k, 8);
        float lm = a3_max<32>(s, -INFINITY) * scale_log2;
        xmax[split * 32 + lane] = lm;
        a3_bar_sync(bar_id, 64);
        if (qd == 0 && split == 0) A3_TRACE(k, 9);
        float m_s = fmaxf(xmax[lane], xmax[32 + lane]);   // scaled reference maximum of the row
        while (w > 0) {
          if (w == 32) {
            uint32_t pk[16];
            a3_exp<32>(s, pk, scale_log2, m_s, sum);
            tmem_st16(p_cur, pk);
            p_cur += 16;
          } else {
            uint32_t pk[8];
            a3_exp<16>(*reinterpret_cast<uint32_t(*)[16]>(&s[0]), pk, scale_log2, m_s, sum);
            tmem_st8(p_cur, pk);
            p_cur += 8;
          }
          c += w;
          if (c >= c_end || (p.debug & 1)) break;
          w = load_chunk(c);
          lm = a3_max<32>(s, -INFINITY) * scale_log2;
          const bool raise = lm > m_s + A3_LAZY;
          if (__any_sync(0xffffffffu, raise)) {
            const float m_new = raise ? lm : m_s;
            rescale(ex2_approx(m_s - m_new));
            m_s = m_new;
          }
        }
        // ---- reconcile the two halves: common reference maximum, row sum ----
        if (qd == 0 && split == 0) A3_TRACE(k, 10);
        mysum = (sum[0] + sum[1]) + (sum[2] + sum[3]);
        xfin[split * 32 + lane] = m_s;
        a3_bar_sync(bar_id, 64);
        if (qd == 0 && split == 0) A3_TRACE(k, 11);
        const float m_fin = fmaxf(xfin[lane], xfin[32 + lane]);
        if (__any_sync(0xffffffffu, m_fin > m_s)) {                          // rare: another split raised its maximum
          const float f = ex2_approx(m_s - m_fin);
          rescale(f);
          mysum *= f;
        }
        }
        rowsum[(g * 2 + split) * 128 + r] = mysum;
        tmem_wait_st();
        if (qd == 0 && split == 0) A3_TRACE(k, 12);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[g]));
      if (qd == 0 && split == 0) A3_TRACE(k, 5);
      if (qd == 0 && split == 1) A3_TRACE(k, 15);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, A3_TMEM_COLS);
}

long long* get_attention_trace();

// -> AGB_ERR_UNSUPPORTED when the shape is outside this kernel's range (the caller falls back to the pipelined kernel)
int attention_split(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                    const int* nkeep, const uint8_t* dst_pos, bf16* ctx, cudaStream_t stream) {
  Att3Params p;
  p.NK = (T + 15) / 16 * 16;
  if (p.NK > A3_MAX_NK || H != heads * A3_D) return AGB_ERR_UNSUPPORTED;
  p.mask = mask; p.words = words; p.rows = rows; p.T = T; p.H = H; p.heads = heads;
  p.units = rows * heads;
  p.mtiles = (T + 127) / 128;
  p.share = share;
  p.kvb = p.NK * 128;
  p.nkeep = nkeep;
  p.dst_pos = dst_pos;
  p.ctx = ctx;
  p.trace = get_attention_trace();
  static const int debug = [] { const char* e = getenv("AGB_ATTN_DEBUG"); return e != nullptr ? atoi(e) : 0; }();
  p.debug = debug;
  CUtensorMap tmQ, tmKV;
  const uint64_t batch_dim = (uint64_t)(rows / share);
  int rc = encode_tmap_3d_bf16(&tmQ, qkv, 3 * (uint64_t)H, (uint64_t)T, batch_dim, (uint64_t)3 * H * 2, (uint64_t)T * 3 * H * 2,
                               A3_D, 128, 1);
  if (rc != AGB_OK) return rc;
  rc = encode_tmap_3d_bf16(&tmKV, qkv, 3 * (uint64_t)H, (uint64_t)T, batch_dim, (uint64_t)3 * H * 2, (uint64_t)T * 3 * H * 2,
                           A3_D, p.NK, 1);
  if (rc != AGB_OK) return rc;
  const int smem = 1024 + 2 * 16384 + A3_KV_RING * 2 * p.kvb + 24 * 8 + (2 * 512 + 1024) * 4 + 64;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  const int grid = p.units < sm_count() ? p.units : sm_count();
  if (nkeep != nullptr) attention_split_kernel<true><<<grid, A3_THREADS, smem, stream>>>(tmQ, tmKV, p);
  else attention_split_kernel<false><<<grid, A3_THREADS, smem, stream>>>(tmQ, tmKV, p);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

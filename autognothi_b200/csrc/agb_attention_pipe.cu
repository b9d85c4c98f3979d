// Fused key-masked attention, second generation: a warp-specialised software pipeline per SM.
//
// Same contract as agb_attention_tc.cu (reference models/vanilla_vit.py:444-463, models/vanilla_bert.py:
// 517-537; bf16 in/out, head dim 64, T <= 512) — the (N,h,T,T) score tensor never leaves the SM and
// masked copies of the input are never built.  Round-1 measurement: the first-generation kernel ran one
// (row, head, m-tile) at a time per CTA and spent 18 % of the step on 4 % of the FLOPs, ~3x off its MUFU
// bound, because load -> mask -> QK^T -> softmax -> PV -> store were serialised and only overlapped
// through 2 CTAs/SM.  Here ONE persistent CTA per SM pipelines "items" (unit = (row, head), m-tile):
//   warp 0      TMA producer: K/V of a unit into a 3-deep ring, Q tiles into a 2-deep ring
//   warp 1      tcgen05 issuer for S_k = Q_k K^T (into TMEM region k&1)
//   warp 2      tcgen05 issuer for O_k = P_k V, the moment P_k arrives (independent of warp 1: no head-of-line wait)
//   warp 3      mask prep (ViT): zero the K rows of masked keys in smem => logit exactly 0 ("scores * mask")
//   warps 4-11  two softmax groups of 4 warps (thread = query row = TMEM lane); group k&1 owns item k,
//               so the MUFU-bound exp of one item overlaps the MMAs / epilogue / loads of its neighbours.
// Softmax: pass 1 row max, pass 2 P = exp2((s - max) * scale*log2e) -> bf16 pairs written back into TMEM over
// the consumed S columns and used as the A operand FROM TENSOR MEMORY of the second MMA; the row sum is
// accumulated in fp32 registers (key padding and, for BERT, masked keys are forced to P = 0 = "finfo.min
// additive mask", so neither numerator nor denominator sees them).
#include <stdlib.h>

#include "agb_common.cuh"

namespace agb {

constexpr int AP_D = 64;
constexpr int AP_THREADS = 384;
constexpr int AP_TMEM_COLS = 512;
constexpr int AP_O_COL = 128;      // O accumulator at columns [128, 192) of the group's 256-column region
constexpr int AP_KV_RING = 3;
constexpr int AP_PREP_THREADS = 32;
constexpr int AP_MODE_VARLEN = 2;   // internal third mode: packed variable-length rows, every packed token is a live key
// internal fourth mode (ViT "logit := 0" masks with the tokens of every row permuted so that the kept ones come first):
// keys [0, nkeep[row]) are kept; the masked keys [nkeep, T) all carry the logit 0, so they are folded into ONE virtual
// key at column nkeep whose logit is log(n_masked) / scale and whose V row is the mean of the masked V rows — the same
// softmax, with about half the columns to exponentiate at the sampler's average coalition size
constexpr int AP_MODE_PREFIX = 3;

struct AttPipeParams {
  const uint32_t* mask;
  int words;
  int rows, T, H, heads, mode;
  int NK;            // keys padded to a multiple of 16
  int units;         // rows * heads
  int mtiles;        // ceil(T / 128)
  int share;         // consecutive mask rows that read the SAME qkv row (first block: one projection per input)
  int groups;        // softmax groups / TMEM regions in flight: 2 (NK <= 256) or 1 (256 < NK <= 512: S needs all 512 columns)
  int ring;          // K/V ring depth (units), limited by shared memory
  int kvb;           // bytes reserved per K (or V) tile: NK * 128, or 2 x 256-row TMA boxes when NK > 256
  const int* cu;     // AP_MODE_VARLEN: row r owns the packed tokens [cu[r], cu[r+1]) of qkv (total_tokens, 3H)
  const int* nkeep;  // AP_MODE_PREFIX: kept tokens of row r (CLS included) = its first nkeep[r] tokens
  bf16* ctx;
  long long* trace;  // optional [items][8] clock64 timestamps of CTA 0 (diagnostics), or nullptr
  // attention-probability dropout (training forward only; DROP instantiations): P kept iff 16 hash bits >= drop_thr
  unsigned drop_thr;
  unsigned long long drop_seed;
  float drop_scale;  // 1 / (1 - p)
  int softmax_pipe;  // 1: software-pipelined TMEM loads over the unmasked key region (AGB_ATTN_SOFTMAX_PIPE, A/B switch)
};

#define AP_TRACE(k, slot)                                                             \
  do {                                                                                \
    if (p.trace != nullptr && blockIdx.x == 0 && (k) < 64 && lane == 0) p.trace[(k) * 8 + (slot)] = clock64(); \
  } while (0)

// ---- softmax helpers (thread = one query row) ---------------------------------------------------------
template <int W>
__device__ __forceinline__ void ap_load(uint32_t addr, uint32_t (&s)[W]) {
  if (W == 64) {
    tmem_ld32(addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
    tmem_ld32(addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
  } else {
    tmem_ld16(addr, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
  }
  tmem_wait_ld();
}

template <int W, bool MASKED, bool VIRT = false>
__device__ __forceinline__ float ap_max_chunk(uint32_t addr, float m, uint32_t live_lo, uint32_t live_hi, int vj = -1,
                                              float vval = 0.f) {
  uint32_t s[W];
  ap_load<W>(addr, s);
#pragma unroll
  for (int j = 0; j < W; ++j) {
    float v = __uint_as_float(s[j]);
    if (VIRT && j == vj) v = vval;          // the virtual key that stands for all masked keys
    if (MASKED && !(((j < 32 ? live_lo : live_hi) >> (j & 31)) & 1u)) v = -INFINITY;
    m = fmaxf(m, v);
  }
  return m;
}

template <int W, bool MASKED, bool DROP = false, bool VIRT = false>
__device__ __forceinline__ float ap_exp_chunk(uint32_t s_addr, uint32_t p_addr, float scale_log2, float m_scaled,
                                              uint32_t live_lo, uint32_t live_hi, uint32_t dkey = 0u, uint32_t pair0 = 0u,
                                              uint32_t dthr = 0u, int vj = -1, float vval = 0.f) {
  uint32_t s[W];
  ap_load<W>(s_addr, s);
  uint32_t pk[W / 2];
  float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
  for (int j = 0; j < W / 2; ++j) {
    float a = __uint_as_float(s[2 * j]), b = __uint_as_float(s[2 * j + 1]);
    if (VIRT) {
      if (2 * j == vj) a = vval;
      if (2 * j + 1 == vj) b = vval;
    }
    if (MASKED) {
      const int j0 = 2 * j, j1 = 2 * j + 1;
      if (!(((j0 < 32 ? live_lo : live_hi) >> (j0 & 31)) & 1u)) a = -INFINITY;
      if (!(((j1 < 32 ? live_lo : live_hi) >> (j1 & 31)) & 1u)) b = -INFINITY;
    }
    const float e0 = ex2_approx(fmaf(a, scale_log2, -m_scaled));
    const float e1 = ex2_approx(fmaf(b, scale_log2, -m_scaled));
    if (j & 1) { sum2 += e0; sum3 += e1; } else { sum0 += e0; sum1 += e1; }
    if (DROP) {      // the row sum keeps every probability; dropped ones only leave the P V product
      const uint32_t x = agb_drop_bits(dkey, pair0 + j);
      pk[j] = pack_bf16x2((x & 0xFFFFu) >= dthr ? e0 : 0.f, (x >> 16) >= dthr ? e1 : 0.f);
    } else
    pk[j] = pack_bf16x2(e0, e1);
  }
  if (W == 64) tmem_st32(p_addr, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
  else         tmem_st8(p_addr, *reinterpret_cast<uint32_t(*)[8]>(&pk[0]));
  return (sum0 + sum1) + (sum2 + sum3);
}

// ---- software-pipelined unmasked region [0, n_fast) (n_fast % 64 == 0): 32-column halves in two register buffers, the
// tcgen05.ld of the next half in flight while the current one is reduced / exponentiated (the serial ld -> wait -> use of
// the 64-column chunks left the softmax warps in long_scoreboard: profiles/r01_layer_ncu_full_end.txt) --------------
__device__ __forceinline__ float ap_max32(const uint32_t (&s)[32], float m) {
  float m0 = m, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    m0 = fmaxf(m0, __uint_as_float(s[j]));
    m1 = fmaxf(m1, __uint_as_float(s[j + 1]));
  }
  return fmaxf(m0, m1);
}

__device__ __forceinline__ float ap_max_fast(uint32_t addr, int n_fast, float m) {
  uint32_t a[32], b[32];
  tmem_ld32(addr, a);
  tmem_wait_ld_dep32(a);
  for (int c = 0; c < n_fast; c += 64) {
    tmem_ld32(addr + c + 32, b);
    m = ap_max32(a, m);
    tmem_wait_ld_dep32(b);
    tmem_ld32(addr + min(c + 64, n_fast - 32), a);      // branch-free: the last iteration re-reads the final half
    m = ap_max32(b, m);
    tmem_wait_ld_dep32(a);
  }
  return m;
}

__device__ __forceinline__ void ap_exp32(const uint32_t (&s)[32], uint32_t (&pk)[16], float scale_log2, float m_scaled,
                                         float (&sum)[4]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float e0 = ex2_approx(fmaf(__uint_as_float(s[2 * j]), scale_log2, -m_scaled));
    const float e1 = ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), scale_log2, -m_scaled));
    sum[(j & 1) * 2] += e0;
    sum[(j & 1) * 2 + 1] += e1;
    pk[j] = pack_bf16x2(e0, e1);
  }
}

// P (bf16 pairs) overlays S columns that are already in registers: the pairs of columns [c, c + 32) go to [c / 2, c / 2 + 16),
// always below the half whose load is in flight
__device__ __forceinline__ float ap_exp_fast(uint32_t s_addr, uint32_t p_addr, int n_fast, float scale_log2, float m_scaled) {
  uint32_t a[32], b[32], pk[16];
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  tmem_ld32(s_addr, a);
  tmem_wait_ld_dep32(a);
  for (int c = 0; c < n_fast; c += 64) {
    tmem_ld32(s_addr + c + 32, b);
    ap_exp32(a, pk, scale_log2, m_scaled, sum);
    tmem_st16(p_addr + (c >> 1), pk);
    tmem_wait_ld_dep32(b);
    // branch-free prefetch; in the last iteration there is nothing left to fetch: read the (unused) columns of this
    // chunk's second half again — NOT columns P has already overwritten
    tmem_ld32(s_addr + min(c + 64, n_fast - 32), a);
    ap_exp32(b, pk, scale_log2, m_scaled, sum);
    tmem_wait_ld_dep32(a);
    tmem_st16(p_addr + (c >> 1) + 16, pk);
  }
  return (sum[0] + sum[1]) + (sum[2] + sum[3]);
}

// live-column bits of the W-wide chunk starting at key c0: key < T, and (BERT) coalition bit set
__device__ __forceinline__ void ap_live_bits(const uint32_t* mrow, int words, int mode, int T, int c0, int W,
                                             uint32_t& lo, uint32_t& hi) {
  const int nvalid = min(max(T - c0, 0), W);
  const uint64_t valid = nvalid >= 64 ? ~0ull : ((1ull << nvalid) - 1ull);
  uint64_t live = valid;
  if (mode == AGB_MASK_NEGINF) {
    const int w0 = c0 >> 5;              // c0 is a multiple of 16; chunks of 64 start on a word boundary
    uint64_t bits = 0;
    if ((c0 & 31) == 0) {
      bits = (w0 < words ? (uint64_t)__ldg(mrow + w0) : 0ull) | ((w0 + 1 < words ? (uint64_t)__ldg(mrow + w0 + 1) : 0ull) << 32);
    } else {
      bits = (w0 < words ? (uint64_t)(__ldg(mrow + w0) >> 16) : 0ull) |
             ((w0 + 1 < words ? (uint64_t)__ldg(mrow + w0 + 1) : 0ull) << 16);
    }
    live &= bits;
  }
  lo = (uint32_t)live;
  hi = (uint32_t)(live >> 32);
}

template <int MODE, int G, bool DROP = false>
__global__ void __launch_bounds__(AP_THREADS, 1)
attention_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const AttPipeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int kvb = p.kvb;                       // bytes of one K (or V) tile
  const int R = (G == 2) ? AP_KV_RING : p.ring;   // two groups <=> NK <= 256 <=> the full ring always fits
  constexpr int rstride = 512 / G;             // TMEM columns per group region
  constexpr int o_col = (G == 2) ? AP_O_COL : 256; // O accumulator overlays S columns the softmax has already consumed
  uint8_t* sQ = smem;                          // [2][128 x 128 B]
  uint8_t* sKV = smem + 2 * 16384;             // [ring][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + p.ring * 2 * p.kvb);
  uint64_t* q_full = bars;                     // [2]
  uint64_t* q_empty = bars + 2;                // [2]
  uint64_t* kv_full = bars + 4;                // [3]
  uint64_t* kv_prep = bars + 7;                // [3]
  uint64_t* kv_empty = bars + 10;              // [3]
  uint64_t* s_full = bars + 13;                // [2]
  uint64_t* p_full = bars + 15;                // [2]
  uint64_t* o_full = bars + 17;                // [2]
  uint64_t* o_free = bars + 19;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&q_full[i]), 1);
      mbar_init(smem_u32(&q_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_full[i]), 4);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&o_free[i]), 4);
    }
    for (int i = 0; i < AP_KV_RING; ++i) {
      mbar_init(smem_u32(&kv_full[i]), 1);
      mbar_init(smem_u32(&kv_prep[i]), AP_PREP_THREADS);
      mbar_init(smem_u32(&kv_empty[i]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), AP_TMEM_COLS);
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

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    const uint32_t e = elect_one();
    for (int k = 0; k < n_items; ++k) {
      const int ui = k / mt, m = k - ui * mt;
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads, head = u - row * p.heads;
      // token coordinate of the unit's first token and batch coordinate of the TMA maps: (0, row / share) for the
      // fixed-length layout (rows, T, 3H); (cu[row], 0) for packed rows (total_tokens, 3H)
      const int tok0 = (MODE == AP_MODE_VARLEN) ? __ldg(p.cu + row) : 0;
      const int brow = (MODE == AP_MODE_VARLEN) ? 0 : row / p.share;
      if (m == 0) {
        const int b = ui % R, n = ui / R;
        if (n > 0) mbar_wait(smem_u32(&kv_empty[b]), (n - 1) & 1);
        const uint32_t bar = smem_u32(&kv_full[b]);
        mbar_arrive_expect_tx_e(e, bar, 2 * kvb);
        const uint32_t dst = smem_u32(sKV + b * 2 * kvb);
        // one TMA box per tile (NK <= 256 rows) or two 256-row boxes (rows past T are zero-filled)
        tma_load_3d_e(e, dst, &tmKV, bar, H + head * AP_D, tok0, brow);
        tma_load_3d_e(e, dst + kvb, &tmKV, bar, 2 * H + head * AP_D, tok0, brow);
        if (G == 1) {       // second 256-row box
          tma_load_3d_e(e, dst + 256 * 128, &tmKV, bar, H + head * AP_D, tok0 + 256, brow);
          tma_load_3d_e(e, dst + kvb + 256 * 128, &tmKV, bar, 2 * H + head * AP_D, tok0 + 256, brow);
        }
        AP_TRACE(k, 0);
      }
      const int qb = k & 1, nq = k >> 1;
      if (nq > 0) mbar_wait(smem_u32(&q_empty[qb]), (nq - 1) & 1);
      const uint32_t qbar = smem_u32(&q_full[qb]);
      mbar_arrive_expect_tx_e(e, qbar, 16384);
      tma_load_3d_e(e, smem_u32(sQ + qb * 16384), &tmQ, qbar, head * AP_D, tok0 + m * 128, brow);
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t e = elect_one();
    const int n0 = NK > 256 ? 256 : NK, n1 = NK - n0;                // S = Q K^T is issued in key chunks of <= 256 (UMMA N limit)
    const uint32_t idesc_s0 = make_idesc_bf16(128, n0, 0, 0);
    const uint32_t idesc_s1 = make_idesc_bf16(128, n1 > 0 ? n1 : 16, 0, 0);
    const uint64_t dk0 = make_smem_desc_sw128(0, 16, 1024);       // K-major operands (Q, K)
    const uint32_t sq0 = smem_u32(sQ) >> 4, skv0 = smem_u32(sKV) >> 4;
    const uint32_t kvb16 = (uint32_t)kvb >> 4;

    // S issuer.  S_k reuses the TMEM region of item k-G: o_free (that item's epilogue has drained O) implies its
    // PV has completed, so no ordering with the PV issuer (warp 2) is needed beyond the barriers.
    for (int k = 0; k < n_items; ++k) {
      const int ui = k / mt, m = k - ui * mt;
      const int b = ui % R, g = k % G, n = k / G, qb = k & 1, nq = k >> 1;
      if (m == 0) mbar_wait(smem_u32(&kv_prep[b]), (ui / R) & 1);
      mbar_wait(smem_u32(&q_full[qb]), nq & 1);
      if (n > 0) mbar_wait(smem_u32(&o_free[g]), (n - 1) & 1);
      tc_fence_after();
      const uint32_t aq = sq0 + qb * (16384 >> 4);
      const uint32_t ak = skv0 + b * 2 * kvb16;
#pragma unroll
      for (int kk = 0; kk < AP_D / 16; ++kk)
        umma_ss_e<1>(e, tmem_base + g * rstride, dk0 + (aq + kk * 2), dk0 + (ak + kk * 2), idesc_s0, kk != 0 ? 1u : 0u);
      if (G == 1 && n1 > 0) {
#pragma unroll
        for (int kk = 0; kk < AP_D / 16; ++kk)
          umma_ss_e<1>(e, tmem_base + g * rstride + 256, dk0 + (aq + kk * 2), dk0 + (ak + (256 * 128 >> 4) + kk * 2), idesc_s1,
                       kk != 0 ? 1u : 0u);
      }
      umma_commit_e<1>(e, smem_u32(&s_full[g]));
      umma_commit_e<1>(e, smem_u32(&q_empty[qb]));
      AP_TRACE(k, 1);
    }
  } else if (warp == 2) {
    // ------------------------------ PV issuer: O_k = P_k V as soon as group k&1 has written P_k ------------------------------
    const uint32_t e = elect_one();
    const uint32_t idesc_o = make_idesc_bf16(128, AP_D, 0, 1);
    const uint64_t dv0 = make_smem_desc_sw128(0, 8192, 1024);     // MN-major V, one 64-column atom (LBO unused)
    const uint32_t skv0 = smem_u32(sKV) >> 4;
    const uint32_t kvb16 = (uint32_t)kvb >> 4;
    for (int k = 0; k < n_items; ++k) {
      const int ui = k / mt, m = k - ui * mt;
      const int b = ui % R, g = k % G, n = k / G;
      mbar_wait(smem_u32(&p_full[g]), n & 1);
      AP_TRACE(k, 2);
      tc_fence_after();
      const uint32_t av = skv0 + b * 2 * kvb16 + kvb16;
      const uint32_t d_o = tmem_base + g * rstride + o_col, a_p = tmem_base + g * rstride;
      const uint64_t dv = dv0 + av;
      int nks = NK / 16;
      if (MODE == AP_MODE_PREFIX) {           // only the kept keys + the virtual key carry probability mass
        const int u = blockIdx.x + ui * grid;
        const int nk = __ldg(p.nkeep + u / p.heads);
        nks = (nk + (nk < T ? 1 : 0) + 15) / 16;     // P beyond these columns is zero or never written: not needed
      }
#pragma unroll 4
      for (int ks = 0; ks < nks; ++ks)
        umma_ts_e(e, d_o, a_p + ks * 8, dv + ks * (2048 >> 4), idesc_o, ks != 0 ? 1u : 0u);
      umma_commit_e<1>(e, smem_u32(&o_full[g]));
      if (m == mt - 1) umma_commit_e<1>(e, smem_u32(&kv_empty[b]));
      AP_TRACE(k, 3);
    }
  } else if (warp == 3) {
    // ------------------------------ mask prep (ViT: zero masked K rows) ------------------------------
    const int tid = lane;
    for (int ui = 0; ui < nu; ++ui) {
      const int b = ui % R, n = ui / R;
      mbar_wait(smem_u32(&kv_full[b]), n & 1);
      if (MODE == AP_MODE_PREFIX) {
        // V row of the virtual key = mean of the masked V rows [nk, T).  Lane = (16-byte chunk of the row, row group): four
        // row groups stride the masked rows, each lane sums its 8 dims, the groups are folded with shuffles.
        const int u = blockIdx.x + ui * grid;
        const int nk = __ldg(p.nkeep + u / p.heads);
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
      }
      if (MODE == AGB_MASK_MUL0) {
        const int u = blockIdx.x + ui * grid;
        const int row = u / p.heads;
        const uint32_t* mrow = p.mask + (long long)row * p.words;
        uint8_t* sK = sKV + b * 2 * kvb;
        for (int j = tid; j < T; j += AP_PREP_THREADS) {
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
  } else {
    // ------------------------------ softmax + epilogue (2 groups x 128 threads) ------------------------------
    const int g = (warp - 4) >> 2;
    const int qd = warp & 3;                        // TMEM lane quarter this warp may touch
    const int r = qd * 32 + lane;                   // query row within the tile = TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + g * rstride;
    const float scale_log2 = 0.125f * 1.4426950408889634f;
    for (int k = g; k < n_items && g < G; k += G) {     // with one group (long sequences) warps 8-11 have nothing to do
      const int n = k / G;
      const int ui = k / mt, m = k - ui * mt;
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads, head = u - row * p.heads;
      const uint32_t* mrow = p.mask + (long long)row * p.words;
      int tok0 = 0, Tr = T;                                // first token / length of this row
      if (MODE == AP_MODE_VARLEN) {
        tok0 = __ldg(p.cu + row);
        Tr = __ldg(p.cu + row + 1) - tok0;
      }
      const bool warp_live = (m * 128 + qd * 32) < Tr;    // warp-uniform: any real query row in this warp?
      // key columns the softmax has to visit: all of them, or (AP_MODE_PREFIX) the kept keys + one virtual key
      int Tk = Tr, NKu = NK, vcol = -1;
      float vval = 0.f;
      if (MODE == AP_MODE_PREFIX) {
        const int nk = __ldg(p.nkeep + row);
        if (nk < T) {
          vcol = nk;
          vval = __log2f((float)(T - nk)) / scale_log2;   // exp2(vval * scale_log2 - m) = n_masked * exp2(0 - m)
        }
        Tk = nk + (nk < T ? 1 : 0);
        NKu = (Tk + 15) / 16 * 16;
      }
      mbar_wait(smem_u32(&s_full[g]), n & 1);
      if (qd == 0) AP_TRACE(k, 4);
      tc_fence_after();
      float inv = 0.f;
      if (warp_live) {
        // ViT: keys [0, n_fast) are all live (masked keys keep their exact-0 logit) -> no per-element selects;
        // the chunk holding the T boundary (and every BERT chunk) takes the masked variant.
        constexpr bool VIRT = (MODE == AP_MODE_PREFIX);
        // (AP_MODE_PREFIX: the chunk that holds the virtual column always takes the masked variant)
        const int n_fast = (MODE == AP_MODE_PREFIX) ? ((Tk - 1) / 64) * 64 : (MODE != AGB_MASK_NEGINF) ? (Tr / 64) * 64 : 0;
        // pass 1: row maximum
        float mx = -INFINITY;
        int c0 = 0;
        if (p.softmax_pipe && n_fast > 0) {
          mx = ap_max_fast(lane_addr, n_fast, mx);
          c0 = n_fast;
        }
        for (; c0 < n_fast; c0 += 64) mx = ap_max_chunk<64, false>(lane_addr + c0, mx, 0u, 0u);
        for (; c0 + 64 <= NKu; c0 += 64) {
          uint32_t lo, hi;
          ap_live_bits(mrow, p.words, MODE, Tk, c0, 64, lo, hi);
          mx = ap_max_chunk<64, true, VIRT>(lane_addr + c0, mx, lo, hi, vcol - c0, vval);
        }
        for (; c0 < NKu; c0 += 16) {
          uint32_t lo, hi;
          ap_live_bits(mrow, p.words, MODE, Tk, c0, 16, lo, hi);
          mx = ap_max_chunk<16, true, VIRT>(lane_addr + c0, mx, lo, hi, vcol - c0, vval);
        }
        const float m_scaled = mx * scale_log2;
        // pass 2: P = exp2(s*scale - max*scale) -> bf16 pairs -> TMEM (overlaying consumed S columns)
        float sum = 0.f;
        c0 = 0;
        const uint32_t dkey = DROP ? agb_drop_key(p.drop_seed, (uint32_t)u, (uint32_t)(m * 128 + r)) : 0u;
        if (!DROP && p.softmax_pipe && n_fast > 0) {
          sum += ap_exp_fast(lane_addr, lane_addr, n_fast, scale_log2, m_scaled);
          c0 = n_fast;
        }
        for (; c0 < n_fast; c0 += 64)
          sum += ap_exp_chunk<64, false, DROP>(lane_addr + c0, lane_addr + (c0 >> 1), scale_log2, m_scaled, 0u, 0u, dkey,
                                               c0 >> 1, p.drop_thr);
        for (; c0 + 64 <= NKu; c0 += 64) {
          uint32_t lo, hi;
          ap_live_bits(mrow, p.words, MODE, Tk, c0, 64, lo, hi);
          sum += ap_exp_chunk<64, true, DROP, VIRT>(lane_addr + c0, lane_addr + (c0 >> 1), scale_log2, m_scaled, lo, hi, dkey,
                                                    c0 >> 1, p.drop_thr, vcol - c0, vval);
        }
        for (; c0 < NKu; c0 += 16) {
          uint32_t lo, hi;
          ap_live_bits(mrow, p.words, MODE, Tk, c0, 16, lo, hi);
          sum += ap_exp_chunk<16, true, DROP, VIRT>(lane_addr + c0, lane_addr + (c0 >> 1), scale_log2, m_scaled, lo, hi, dkey,
                                                    c0 >> 1, p.drop_thr, vcol - c0, vval);
        }
        tmem_wait_st();
        inv = (DROP ? p.drop_scale : 1.0f) / sum;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[g]));
      if (qd == 0) AP_TRACE(k, 5);
      // epilogue: O / rowsum -> bf16 ctx
      mbar_wait(smem_u32(&o_full[g]), n & 1);
      if (qd == 0) AP_TRACE(k, 6);
      tc_fence_after();
      uint32_t o[64];
      if (warp_live) ap_load<64>(lane_addr + o_col, o);
      // O is in registers: release the TMEM region (the next S_k+2 may overwrite it) BEFORE scaling / storing
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&o_free[g]));
      if (qd == 0) AP_TRACE(k, 7);
      if (warp_live) {
        const int tq = m * 128 + r;
        if (tq < Tr) {
          bf16* dst = p.ctx + ((MODE == AP_MODE_VARLEN ? (long long)tok0 : (long long)row * T) + tq) * H + head * AP_D;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + 8 * c) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, AP_TMEM_COLS);
}

static long long* g_attention_trace = nullptr;
void set_attention_trace(long long* t) { g_attention_trace = t; }
long long* get_attention_trace() { return g_attention_trace; }
// 0 auto (split-softmax kernel where it applies, else the pipelined one), 1 first-generation kernel, 2 pipelined kernel
static int g_attention_variant = 0;
int attention_split(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                    const int* nkeep, const uint8_t* dst_pos, bf16* ctx, cudaStream_t stream);
void set_attention_variant(int v) { g_attention_variant = v; }
int get_attention_variant() { return g_attention_variant; }

// T = sequence length (fixed layout) or an upper bound of the row lengths (packed layout, cu != nullptr)
static int attention_pipe_launch(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H,
                                 int heads, int mode, bf16* ctx, const int* cu, int total_tokens, cudaStream_t stream,
                                 unsigned drop_thr = 0, unsigned long long drop_seed = 0, const int* nkeep = nullptr) {
  AttPipeParams p;
  p.nkeep = nkeep;
  p.drop_thr = drop_thr;
  p.drop_seed = drop_seed;
  p.drop_scale = 65536.0f / (65536.0f - (float)drop_thr);
  static const int softmax_pipe = [] {
    const char* e = getenv("AGB_ATTN_SOFTMAX_PIPE");
    return e != nullptr ? atoi(e) : 0;
  }();
  p.softmax_pipe = softmax_pipe;
  p.mask = mask; p.words = words; p.rows = rows; p.T = T; p.H = H; p.heads = heads; p.mode = mode;
  p.NK = (T + 15) / 16 * 16;
  p.units = rows * heads;
  p.mtiles = (T + 127) / 128;
  p.share = share;
  p.cu = cu;
  p.ctx = ctx;
  p.trace = g_attention_trace;
  // NK <= 256: two softmax groups ping-pong over 2 x 256 TMEM columns, one TMA box per K / V tile, K/V ring of 3 units.
  // 256 < NK <= 512: S needs all 512 columns (one group), K / V arrive as two 256-row boxes, ring as deep as smem allows.
  const int box_rows = p.NK > 256 ? 256 : p.NK;
  p.groups = p.NK > 256 ? 1 : 2;
  p.kvb = p.NK > 256 ? 2 * 256 * 128 : p.NK * 128;
  const int smem_budget = 227 * 1024 - (1024 + 2 * 16384 + 256);
  p.ring = smem_budget / (2 * p.kvb);
  if (p.ring > AP_KV_RING) p.ring = AP_KV_RING;
  if (p.ring < 1) return AGB_ERR_UNSUPPORTED;
  // fixed layout: (rows / share, T, 3H); packed layout: one "batch" of total_tokens rows
  const uint64_t tok_dim = cu ? (uint64_t)total_tokens : (uint64_t)T;
  const uint64_t batch_dim = cu ? 1 : (uint64_t)(rows / share);
  CUtensorMap tmQ, tmKV;
  int rc = encode_tmap_3d_bf16(&tmQ, qkv, 3 * (uint64_t)H, tok_dim, batch_dim, (uint64_t)3 * H * 2, tok_dim * 3 * H * 2,
                               AP_D, 128, 1);
  if (rc != AGB_OK) return rc;
  rc = encode_tmap_3d_bf16(&tmKV, qkv, 3 * (uint64_t)H, tok_dim, batch_dim, (uint64_t)3 * H * 2, tok_dim * 3 * H * 2,
                           AP_D, box_rows, 1);
  if (rc != AGB_OK) return rc;
  const int smem = 1024 + 2 * 16384 + p.ring * 2 * p.kvb + 256;
  static int configured_smem = 0;
  if (smem > configured_smem) {
#define AP_SET(M_, G_) \
  AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<M_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
    AP_SET(AGB_MASK_MUL0, 2); AP_SET(AGB_MASK_NEGINF, 2); AP_SET(AP_MODE_VARLEN, 2);
    AP_SET(AGB_MASK_MUL0, 1); AP_SET(AGB_MASK_NEGINF, 1); AP_SET(AP_MODE_VARLEN, 1);
#undef AP_SET
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<AGB_MASK_MUL0, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<AGB_MASK_NEGINF, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<AP_MODE_PREFIX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  const int grid = p.units < sm_count() ? p.units : sm_count();
#define AP_GO(M_, G_) attention_pipe_kernel<M_, G_><<<grid, AP_THREADS, smem, stream>>>(tmQ, tmKV, p)
  const int kmode = cu ? AP_MODE_VARLEN : mode;
  if (nkeep != nullptr) {
    // kept-first token order (ViT masks): keys [0, nkeep[row]) + one virtual key for the masked rest
    if (cu != nullptr || share != 1 || p.groups != 2 || mode != AGB_MASK_MUL0 || drop_thr > 0) {
      set_last_error("prefix-mask attention: ViT mask semantics, fixed layout, T <= 256 (T = %d)", T);
      return AGB_ERR_UNSUPPORTED;
    }
    attention_pipe_kernel<AP_MODE_PREFIX, 2><<<grid, AP_THREADS, smem, stream>>>(tmQ, tmKV, p);
    AGB_CHECK_CUDA(cudaGetLastError());
    return AGB_OK;
  }
  if (drop_thr > 0) {
    // training forward with attention-probability dropout: fixed layout, T <= 256 (the adjoint's range)
    if (cu != nullptr || share != 1 || p.groups != 2) {
      set_last_error("attention dropout: fixed-layout rows with T <= 256 only (T = %d)", T);
      return AGB_ERR_UNSUPPORTED;
    }
    if (kmode == AGB_MASK_MUL0) attention_pipe_kernel<AGB_MASK_MUL0, 2, true><<<grid, AP_THREADS, smem, stream>>>(tmQ, tmKV, p);
    else attention_pipe_kernel<AGB_MASK_NEGINF, 2, true><<<grid, AP_THREADS, smem, stream>>>(tmQ, tmKV, p);
    AGB_CHECK_CUDA(cudaGetLastError());
    return AGB_OK;
  }
  if (p.groups == 2) {
    if (kmode == AGB_MASK_MUL0) AP_GO(AGB_MASK_MUL0, 2);
    else if (kmode == AGB_MASK_NEGINF) AP_GO(AGB_MASK_NEGINF, 2);
    else AP_GO(AP_MODE_VARLEN, 2);
  } else {
    if (kmode == AGB_MASK_MUL0) AP_GO(AGB_MASK_MUL0, 1);
    else if (kmode == AGB_MASK_NEGINF) AP_GO(AGB_MASK_NEGINF, 1);
    else AP_GO(AP_MODE_VARLEN, 1);
  }
#undef AP_GO
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

int attention_pipe(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                   int mode, bf16* ctx, cudaStream_t stream) {
  if (g_attention_variant == 1) return AGB_ERR_UNSUPPORTED;
  if ((g_attention_variant == 0 || g_attention_variant == 3) && mode == AGB_MASK_MUL0) {     // third generation (agb_attention_split.cu): T <= 208
    const int rc = attention_split(qkv, mask, words, rows, share, T, H, heads, nullptr, nullptr, ctx, stream);
    if (rc != AGB_ERR_UNSUPPORTED) return rc;
  }
  return attention_pipe_launch(qkv, mask, words, rows, share, T, H, heads, mode, ctx, nullptr, 0, stream);
}

// ViT-masked attention whose output rows are written in another token order: query token t of row r goes to position
// dst_pos[r, t] (the first block of the kept-first evaluation order: projections shared per input, output per coalition)
int attention_scatter(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                      const uint8_t* dst_pos, bf16* ctx, cudaStream_t stream) {
  AGB_REQUIRE(share >= 1 && rows % share == 0, "share must divide the number of mask rows");
  AGB_REQUIRE(rows >= 0 && T > 0 && T <= 256 && heads > 0 && H == heads * AP_D, "scatter attention shape (head dim 64, T <= 256)");
  AGB_REQUIRE(words * 32 >= T, "mask words");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(qkv && mask && ctx && dst_pos, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0, "alignment");
  const int rc = attention_split(qkv, mask, words, rows, share, T, H, heads, nullptr, dst_pos, ctx, stream);
  if (rc == AGB_ERR_UNSUPPORTED) set_last_error("scatter attention: T <= 208 (T = %d)", T);
  return rc;
}

// ViT-masked attention on rows whose tokens were permuted so that the kept ones come first (nkeep[row] of them, CLS
// included): the masked keys are folded into one virtual key (see AP_MODE_PREFIX).
int attention_pipe_prefix(const bf16* qkv, const int* nkeep, int rows, int T, int H, int heads, bf16* ctx, cudaStream_t stream) {
  AGB_REQUIRE(rows >= 0 && T > 0 && heads > 0 && H == heads * AP_D, "prefix attention shape (head dim 64)");
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(qkv && nkeep && ctx, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0, "alignment");
  // third generation (agb_attention_split.cu) for T <= 208: 338 vs 395 us at the bench shape (profiles/r02_attention_split_notes.txt)
  if (g_attention_variant == 0 || g_attention_variant == 3) {
    const int rc = attention_split(qkv, nullptr, 0, rows, 1, T, H, heads, nkeep, nullptr, ctx, stream);
    if (rc != AGB_ERR_UNSUPPORTED) return rc;
  }
  return attention_pipe_launch(qkv, nullptr, 0, rows, 1, T, H, heads, AGB_MASK_MUL0, ctx, nullptr, 0, stream, 0, 0, nkeep);
}

int attention_pipe_dropout(const bf16* qkv, const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode,
                           bf16* ctx, unsigned thr16, unsigned long long seed, cudaStream_t stream) {
  return attention_pipe_launch(qkv, mask, words, rows, 1, T, H, heads, mode, ctx, nullptr, 0, stream, thr16, seed);
}

// Packed variable-length rows (masked-token dropping): plain attention inside each segment [cu[r], cu[r+1]).
int attention_varlen(const bf16* qkv, const int* cu, int rows, int max_len, int total_tokens, int H, int heads, bf16* ctx,
                     cudaStream_t stream) {
  AGB_REQUIRE(rows >= 0 && max_len > 0 && max_len <= 512 && heads > 0 && H == heads * AP_D, "varlen attention shape");
  if (rows == 0 || total_tokens == 0) return AGB_OK;
  AGB_REQUIRE(qkv && cu && ctx, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0, "alignment");
  return attention_pipe_launch(qkv, nullptr, 0, rows, 1, max_len, H, heads, AGB_MASK_NEGINF, ctx, cu, total_tokens, stream);
}

}  // namespace agb

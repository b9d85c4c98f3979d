// Adjoint of the fused key-masked attention on tcgen05 / TMEM / TMA (bf16, head dim 64, T <= 256).
//
// Used by explainer training (reference: autograd through models/vanilla_vit.py:444-463 and
// models/vanilla_bert.py:517-537).  The round-1 CUDA-core adjoint took 60 % of the explainer step
// (tools/train_profile.py); this kernel keeps every contraction on the tensor cores and, like the
// forward, never materialises a T x T tensor in HBM and needs no atomics:
//   phase 0 (per 128-query tile):  S = Q K~^T and dP = dO V^T for ALL keys (TMEM 2 x 256 columns) ->
//           per-row log-sum-exp L_i and D_i = sum_j P_ij dP_ij  (kept in registers, thread = query row)
//   main    (per 128-key block jb, per query tile m):
//           S_blk, dP_blk (tcgen05, TMEM) -> P = exp2(S*c - L), dS = P (dP - D) / sqrt(d) -> bf16 -> smem
//           dV_jb += P^T dO_m,  dK_jb += dS^T Q_m,  dQ_m += dS K~_jb   (P / dS are read from smem both as
//           K-major and, transposed for free, as MN-major operands; dO / Q / K~ tiles double as MN-major B)
//   TMEM:   [S_blk 128 | dP_blk 128 | dV 64 | dK 64 | dQ_0 64 | dQ_1 64] = 512 columns, all accumulation on chip.
// Mask semantics as in the forward: ViT keys with bit 0 have their K rows zeroed in smem (logit exactly 0,
// still a live softmax column, no gradient to that key's K); BERT masked keys have P = 0.
// One CTA per SM, one (row, head) unit at a time: warp 0 TMA, warp 1 MMA issue, warps 2-5 = 128 row threads.
#include "agb_common.cuh"

namespace agb {

constexpr int BT_D = 64;
constexpr int BT_THREADS = 192;
constexpr int BT_TMEM_COLS = 512;
constexpr int BT_COL_S = 0, BT_COL_DP = 128, BT_COL_DV = 256, BT_COL_DK = 320, BT_COL_DQ = 384;
constexpr int BT_TILE = 16384;   // one 128-row x 128-byte SW128 tile

struct AttBwdParams {
  const uint32_t* mask;
  int words;
  int rows, T, H, heads, mode;
  int NK;            // keys padded to a multiple of 16
  int units;         // rows * heads
  int mtiles;        // ceil(T / 128) query tiles
  int njb;           // ceil(NK / 128) key blocks
  bf16* dqkv;
  // attention-probability dropout of the forward (DROP instantiation): O = (P o M / (1-p)) V with M regenerated here
  unsigned drop_thr;
  unsigned long long drop_seed;
  float drop_scale;
};

__device__ __forceinline__ uint32_t bt_live_word(const uint32_t* mrow, int words, int mode, int T, int j0) {
  // live bits of keys j0 .. j0+31 (j0 a multiple of 32): key < T and, for BERT, coalition bit set
  const int nvalid = min(max(T - j0, 0), 32);
  uint32_t live = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
  if (mode == AGB_MASK_NEGINF) live &= ((j0 >> 5) < words) ? __ldg(mrow + (j0 >> 5)) : 0u;
  return live;
}

template <bool DROP>
__global__ void __launch_bounds__(BT_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmdO, const AttBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int kvb = p.NK * 128;
  uint8_t* sQ = smem;                     // [2][BT_TILE]
  uint8_t* sdO = sQ + 2 * BT_TILE;        // [2][BT_TILE]
  uint8_t* sP = sdO + 2 * BT_TILE;        // [2 key atoms][BT_TILE]
  uint8_t* sdS = sP + 2 * BT_TILE;        // [2 key atoms][BT_TILE]
  uint8_t* sK = sdS + 2 * BT_TILE;        // [NK x 128 B]
  uint8_t* sV = sK + kvb;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kvb);
  uint64_t* bar_load = bars + 0;
  uint64_t* bar_prep = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_g = bars + 4;
  uint64_t* bar_e = bars + 5;      // dV / dK accumulators of a key block drained (once per key block)
  uint64_t* bar_done = bars + 6;   // unit fully drained: smem + TMEM reusable (once per unit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmdO);
    mbar_init(smem_u32(bar_load), 1);
    mbar_init(smem_u32(bar_prep), 128);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_p), 128);
    mbar_init(smem_u32(bar_g), 1);
    mbar_init(smem_u32(bar_e), 128);
    mbar_init(smem_u32(bar_done), 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), BT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int H = p.H, T = p.T, NK = p.NK, mt = p.mtiles, njb = p.njb;
  const int grid = gridDim.x;
  const int nu = (p.units - (int)blockIdx.x + grid - 1) / grid;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    const uint32_t e = elect_one();
    for (int ui = 0; ui < nu; ++ui) {
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads, head = u - row * p.heads;
      if (ui > 0) mbar_wait(smem_u32(bar_done), (ui - 1) & 1);      // previous unit fully drained
      const uint32_t bar = smem_u32(bar_load);
      mbar_arrive_expect_tx_e(e, bar, 2 * mt * BT_TILE + 2 * kvb);
      for (int m = 0; m < mt; ++m) {
        tma_load_3d_e(e, smem_u32(sQ + m * BT_TILE), &tmQ, bar, head * BT_D, m * 128, row);
        tma_load_3d_e(e, smem_u32(sdO + m * BT_TILE), &tmdO, bar, head * BT_D, m * 128, row);
      }
      tma_load_3d_e(e, smem_u32(sK), &tmKV, bar, H + head * BT_D, 0, row);
      tma_load_3d_e(e, smem_u32(sV), &tmKV, bar, 2 * H + head * BT_D, 0, row);
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t e = elect_one();
    const uint64_t d_kmaj = make_smem_desc_sw128(0, 16, 1024);          // K-major [rows x 64] tile
    const uint64_t d_mn1 = make_smem_desc_sw128(0, BT_TILE, 1024);      // MN-major, 64-wide atoms BT_TILE apart
    const uint32_t aQ = smem_u32(sQ) >> 4, adO = smem_u32(sdO) >> 4, aP = smem_u32(sP) >> 4, adS = smem_u32(sdS) >> 4;
    const uint32_t aK = smem_u32(sK) >> 4, aV = smem_u32(sV) >> 4;
    const uint32_t T16 = BT_TILE >> 4;
    const uint32_t idesc_full = make_idesc_bf16(128, NK, 0, 0);
    const uint32_t idesc_tt = make_idesc_bf16(128, BT_D, 1, 1);         // A^T (MN-major) x B (MN-major)
    const uint32_t idesc_kt = make_idesc_bf16(128, BT_D, 0, 1);         // A (K-major)   x B (MN-major)
    uint32_t cp = 0, ce = 0;     // phases of bar_p / bar_e consumed so far (every phase is waited on, in order)
    for (int ui = 0; ui < nu; ++ui) {
      mbar_wait(smem_u32(bar_load), ui & 1);
      mbar_wait(smem_u32(bar_prep), ui & 1);
      tc_fence_after();
      // ---- phase 0: full-row S and dP per query tile ----
      for (int m = 0; m < mt; ++m) {
        if (m > 0) { mbar_wait(smem_u32(bar_p), cp & 1); ++cp; tc_fence_after(); }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss_e<1>(e, tmem_base + 0, d_kmaj + (aQ + m * T16 + kk * 2), d_kmaj + (aK + kk * 2), idesc_full, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss_e<1>(e, tmem_base + 256, d_kmaj + (adO + m * T16 + kk * 2), d_kmaj + (aV + kk * 2), idesc_full, kk != 0);
        umma_commit_e<1>(e, smem_u32(bar_s));
      }
      mbar_wait(smem_u32(bar_p), cp & 1); ++cp;
      tc_fence_after();
      // ---- main: key blocks x query tiles ----
      for (int jb = 0; jb < njb; ++jb) {
        const int nb = min(128, NK - jb * 128);
        const uint32_t idesc_blk = make_idesc_bf16(128, nb, 0, 0);
        for (int m = 0; m < mt; ++m) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_e<1>(e, tmem_base + BT_COL_S, d_kmaj + (aQ + m * T16 + kk * 2), d_kmaj + (aK + jb * T16 + kk * 2),
                         idesc_blk, kk != 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_e<1>(e, tmem_base + BT_COL_DP, d_kmaj + (adO + m * T16 + kk * 2), d_kmaj + (aV + jb * T16 + kk * 2),
                         idesc_blk, kk != 0);
          umma_commit_e<1>(e, smem_u32(bar_s));
          mbar_wait(smem_u32(bar_p), cp & 1); ++cp;      // P / dS are in smem, S / dP columns consumed
          if (m == 0 && (ui | jb) != 0) { mbar_wait(smem_u32(bar_e), ce & 1); ++ce; }   // dV / dK accumulators drained
          tc_fence_after();
          // dV_jb += P^T dO_m ; dK_jb += dS^T Q_m   (K dimension = the 128 query rows of tile m)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            umma_ss_e<1>(e, tmem_base + BT_COL_DV, d_mn1 + (aP + ks * (2048 >> 4)), d_mn1 + (adO + m * T16 + ks * (2048 >> 4)),
                         idesc_tt, (m | ks) != 0);
            umma_ss_e<1>(e, tmem_base + BT_COL_DK, d_mn1 + (adS + ks * (2048 >> 4)), d_mn1 + (aQ + m * T16 + ks * (2048 >> 4)),
                         idesc_tt, (m | ks) != 0);
          }
          // dQ_m += dS K~_jb   (K dimension = the nb keys of this block)
          for (int ks = 0; ks < nb / 16; ++ks)
            umma_ss_e<1>(e, tmem_base + BT_COL_DQ + m * 64, d_kmaj + (adS + (ks >> 2) * T16 + (ks & 3) * 2),
                         d_mn1 + (aK + jb * T16 + ks * (2048 >> 4)), idesc_kt, (jb | ks) != 0);
          umma_commit_e<1>(e, smem_u32(bar_g));
        }
      }
    }
  } else {
    // ------------------------------ row threads (prep, statistics, P / dS, epilogues) ------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;                                    // TMEM lane = row inside a 128-tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c2 = 0.125f * 1.4426950408889634f;                  // 1/sqrt(d) * log2(e)
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    uint32_t cs = 0, cg = 0;
    for (int ui = 0; ui < nu; ++ui) {
      const int u = blockIdx.x + ui * grid;
      const int row = u / p.heads, head = u - row * p.heads;
      const uint32_t* mrow = p.mask + (long long)row * p.words;
      mbar_wait(smem_u32(bar_load), ui & 1);
      if (p.mode == AGB_MASK_MUL0) {
        for (int j = r; j < T; j += 128) {
          if (!((__ldg(mrow + (j >> 5)) >> (j & 31)) & 1u)) {
            uint4* kr = reinterpret_cast<uint4*>(sK + j * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) kr[c] = make_uint4(0, 0, 0, 0);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(bar_prep));

      // ---- phase 0: L_i (log2 domain) and D_i for this thread's row of each query tile ----
      float L[2] = {0.f, 0.f}, Dv[2] = {0.f, 0.f};
      uint32_t dkey[2] = {0u, 0u};
      if (DROP) {
        dkey[0] = agb_drop_key(p.drop_seed, (uint32_t)u, (uint32_t)r);
        dkey[1] = agb_drop_key(p.drop_seed, (uint32_t)u, (uint32_t)(128 + r));
      }
      for (int m = 0; m < mt; ++m) {
        mbar_wait(smem_u32(bar_s), cs & 1); ++cs;
        tc_fence_after();
        const bool row_live = (m * 128 + r) < T;
        const bool warp_live = (m * 128 + q * 32) < T;
        if (warp_live) {
          float mx = -INFINITY;
          for (int j0 = 0; j0 < NK; j0 += 32) {     // NK is a multiple of 16: the last chunk may be half valid
            uint32_t s[32];
            if (j0 + 32 <= NK) tmem_ld32(lane_addr + j0, s);
            else { tmem_ld16(lane_addr + j0, *reinterpret_cast<uint32_t(*)[16]>(&s[0])); }
            tmem_wait_ld();
            const uint32_t live = bt_live_word(mrow, p.words, p.mode, T, j0);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((live >> j) & 1u) mx = fmaxf(mx, __uint_as_float(s[j]));
          }
          float sum = 0.f, dsum = 0.f;
          const float mxs = mx * c2;
          for (int j0 = 0; j0 < NK; j0 += 32) {
            uint32_t s[32], d[32];
            if (j0 + 32 <= NK) { tmem_ld32(lane_addr + j0, s); tmem_ld32(lane_addr + 256 + j0, d); }
            else {
              tmem_ld16(lane_addr + j0, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
              tmem_ld16(lane_addr + 256 + j0, *reinterpret_cast<uint32_t(*)[16]>(&d[0]));
            }
            tmem_wait_ld();
            const uint32_t live = bt_live_word(mrow, p.words, p.mode, T, j0);
            uint32_t xb = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (DROP && (j & 1) == 0) xb = agb_drop_bits(dkey[m], (uint32_t)(j0 + j) >> 1);
              if ((live >> j) & 1u) {
                const float pe = ex2_approx(fmaf(__uint_as_float(s[j]), c2, -mxs));
                sum += pe;
                float dj = __uint_as_float(d[j]);
                if (DROP) dj = (((j & 1) ? (xb >> 16) : (xb & 0xFFFFu)) >= p.drop_thr) ? dj * p.drop_scale : 0.f;
                dsum = fmaf(pe, dj, dsum);
              }
            }
          }
          if (row_live) {
            L[m] = mxs + log2f(sum);
            Dv[m] = dsum / sum;
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(bar_p));
      }

      // ---- main ----
      for (int jb = 0; jb < njb; ++jb) {
        const int nb = min(128, NK - jb * 128);
        for (int m = 0; m < mt; ++m) {
          mbar_wait(smem_u32(bar_s), cs & 1); ++cs;     // also implies the previous block's gradient MMAs retired
          tc_fence_after();
          const bool row_live = (m * 128 + r) < T;
          const float Li = L[m], Di = Dv[m];
          for (int jc = 0; jc < nb; jc += 32) {
            uint32_t s[32], d[32];
            if (jc + 32 <= nb) { tmem_ld32(lane_addr + BT_COL_S + jc, s); tmem_ld32(lane_addr + BT_COL_DP + jc, d); }
            else {
              tmem_ld16(lane_addr + BT_COL_S + jc, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
              tmem_ld16(lane_addr + BT_COL_DP + jc, *reinterpret_cast<uint32_t(*)[16]>(&d[0]));
            }
            tmem_wait_ld();
            uint32_t live = bt_live_word(mrow, p.words, p.mode, T, jb * 128 + jc);
            if (!row_live) live = 0u;
            const int ncol = min(32, nb - jc);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {           // 8 keys = one 16-byte chunk of P and of dS
              if (c8 * 8 < ncol) {
                uint32_t pp[4], dd[4];
#pragma unroll
                for (int h2 = 0; h2 < 4; ++h2) {
                  float pv[2], dv[2];
                  const uint32_t xb = DROP ? agb_drop_bits(dkey[m], (uint32_t)(jb * 128 + jc + c8 * 8 + h2 * 2) >> 1) : 0u;
#pragma unroll
                  for (int o = 0; o < 2; ++o) {
                    const int j = c8 * 8 + h2 * 2 + o;
                    const bool lv = (live >> j) & 1u;
                    const float pe = lv ? ex2_approx(fmaf(__uint_as_float(s[j]), c2, -Li)) : 0.f;
                    float dj = __uint_as_float(d[j]);
                    float pk = pe;
                    if (DROP) {      // dP and the P of dV carry the dropout mask; dS = P (dP o M' - D) keeps the full P
                      const float mk = ((o ? (xb >> 16) : (xb & 0xFFFFu)) >= p.drop_thr) ? p.drop_scale : 0.f;
                      dj *= mk;
                      pk *= mk;
                    }
                    pv[o] = pk;
                    dv[o] = pe * (dj - Di) * 0.125f;
                  }
                  pp[h2] = pack_bf16x2(pv[0], pv[1]);
                  dd[h2] = pack_bf16x2(dv[0], dv[1]);
                }
                const int jj = jc + c8 * 8;                         // key offset inside the block
                const uint32_t off = (jj >> 6) * BT_TILE + r * 128 + ((((jj & 63) >> 3) ^ swz) << 4);
                *reinterpret_cast<uint4*>(sP + off) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
                *reinterpret_cast<uint4*>(sdS + off) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
              }
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(smem_u32(bar_p));

          if (m == mt - 1) {
            // this key block's dV / dK are complete once the gradient MMAs of (jb, m) retire
            cg += mt;
            mbar_wait(smem_u32(bar_g), (cg - 1) & 1);
            tc_fence_after();
            const int j = jb * 128 + r;
            const bool warp_has = (jb * 128 + q * 32) < T;
            if (warp_has) {
              uint32_t v[64];
              tmem_ld32(lane_addr + BT_COL_DV, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
              tmem_ld32(lane_addr + BT_COL_DV + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
              tmem_wait_ld();
              if (j < T) {
                bf16* dst = p.dqkv + ((long long)row * T + j) * 3 * H + 2 * H + head * BT_D;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  *reinterpret_cast<uint4*>(dst + 8 * c) = make_uint4(
                      pack_bf16x2(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1])),
                      pack_bf16x2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])),
                      pack_bf16x2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])),
                      pack_bf16x2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])));
              }
              tmem_ld32(lane_addr + BT_COL_DK, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
              tmem_ld32(lane_addr + BT_COL_DK + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
              tmem_wait_ld();
              if (j < T) {
                // ViT: a masked key's logit is the constant 0 -> no gradient reaches its K row
                const bool keep = (p.mode != AGB_MASK_MUL0) || ((__ldg(mrow + (j >> 5)) >> (j & 31)) & 1u);
                const float kz = keep ? 1.f : 0.f;
                bf16* dst = p.dqkv + ((long long)row * T + j) * 3 * H + H + head * BT_D;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  *reinterpret_cast<uint4*>(dst + 8 * c) = make_uint4(
                      pack_bf16x2(__uint_as_float(v[8 * c + 0]) * kz, __uint_as_float(v[8 * c + 1]) * kz),
                      pack_bf16x2(__uint_as_float(v[8 * c + 2]) * kz, __uint_as_float(v[8 * c + 3]) * kz),
                      pack_bf16x2(__uint_as_float(v[8 * c + 4]) * kz, __uint_as_float(v[8 * c + 5]) * kz),
                      pack_bf16x2(__uint_as_float(v[8 * c + 6]) * kz, __uint_as_float(v[8 * c + 7]) * kz));
              }
            }
            if (jb == njb - 1) {
              // dQ of both query tiles is complete as well
              for (int mm = 0; mm < mt; ++mm) {
                const int i = mm * 128 + r;
                if ((mm * 128 + q * 32) < T) {
                  uint32_t v[64];
                  tmem_ld32(lane_addr + BT_COL_DQ + mm * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                  tmem_ld32(lane_addr + BT_COL_DQ + mm * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                  tmem_wait_ld();
                  if (i < T) {
                    bf16* dst = p.dqkv + ((long long)row * T + i) * 3 * H + head * BT_D;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                      *reinterpret_cast<uint4*>(dst + 8 * c) = make_uint4(
                          pack_bf16x2(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1])),
                          pack_bf16x2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])),
                          pack_bf16x2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])),
                          pack_bf16x2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])));
                  }
                }
              }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(bar_e));
            if (jb == njb - 1) mbar_arrive(smem_u32(bar_done));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BT_TMEM_COLS);
}

// Returns AGB_ERR_UNSUPPORTED for shapes this kernel does not cover (the caller falls back to the CUDA-core adjoint).
int attention_bwd_tc(const bf16* qkv, const bf16* dctx, const uint32_t* mask, int words, int rows, int T, int H,
                     int heads, int mode, bf16* dqkv, cudaStream_t stream, unsigned drop_thr, unsigned long long drop_seed) {
  if (T > 256 || H != heads * BT_D) return AGB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(dctx) & 15) ||
      (reinterpret_cast<uintptr_t>(dqkv) & 15))
    return AGB_ERR_UNSUPPORTED;
  AttBwdParams p;
  p.mask = mask; p.words = words; p.rows = rows; p.T = T; p.H = H; p.heads = heads; p.mode = mode;
  p.NK = (T + 15) / 16 * 16;
  p.units = rows * heads;
  p.mtiles = (T + 127) / 128;
  p.njb = (p.NK + 127) / 128;
  p.dqkv = dqkv;
  p.drop_thr = drop_thr;
  p.drop_seed = drop_seed;
  p.drop_scale = 65536.0f / (65536.0f - (float)drop_thr);
  CUtensorMap tmQ, tmKV, tmdO;
  int rc = encode_tmap_3d_bf16(&tmQ, qkv, 3 * (uint64_t)H, T, rows, (uint64_t)3 * H * 2, (uint64_t)T * 3 * H * 2,
                               BT_D, 128, 1);
  if (rc != AGB_OK) return rc;
  rc = encode_tmap_3d_bf16(&tmKV, qkv, 3 * (uint64_t)H, T, rows, (uint64_t)3 * H * 2, (uint64_t)T * 3 * H * 2, BT_D,
                           p.NK, 1);
  if (rc != AGB_OK) return rc;
  rc = encode_tmap_3d_bf16(&tmdO, dctx, (uint64_t)H, T, rows, (uint64_t)H * 2, (uint64_t)T * H * 2, BT_D, 128, 1);
  if (rc != AGB_OK) return rc;
  const int smem = 1024 + 8 * BT_TILE + 2 * p.NK * 128 + 128;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  const int grid = p.units < sm_count() ? p.units : sm_count();
  if (drop_thr > 0) attention_bwd_tc_kernel<true><<<grid, BT_THREADS, smem, stream>>>(tmQ, tmKV, tmdO, p);
  else attention_bwd_tc_kernel<false><<<grid, BT_THREADS, smem, stream>>>(tmQ, tmKV, tmdO, p);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

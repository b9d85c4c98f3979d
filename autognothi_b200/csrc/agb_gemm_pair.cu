// Dense bf16 GEMM, second generation:  C[M,N] = act(alpha * A[M,K] * B[N,K]^T + bias) (+ residual)
//
// Same contract as agb_gemm_tc.cu (the nn.Linear replacement; reference models/vanilla_vit.py:437-441,
// 473-479, 487-493, 506-513 and models/vanilla_bert.py:503-537, 556-604) but built around the two
// findings of the round-1 ncu captures (profiles/r01_gemm_ncu.txt): the K=768 shapes were paced by the
// epilogue (tensor pipe 30-45 % active) and the 1-CTA 128x256 tile pays full B-operand smem traffic.
//   * CG = 2: a CTA pair (cluster of 2, same TPC) computes one 256x256 tile with
//     tcgen05.mma.cta_group::2 — each CTA stages its own 128 A rows and HALF of the B tile, the
//     leader's single thread issues M=256 UMMAs that read both halves, and each CTA keeps its
//     128x256 fp32 accumulator in its own TMEM (double-buffered, 512 columns).
//     CG = 1 keeps one CTA per 128x256 tile (small problems, odd tile counts).
//   * TMA epilogue: thread = accumulator row; tcgen05.ld 32 columns -> alpha/bias/GELU in registers ->
//     st.shared into a SWIZZLE_128B staging box (conflict-free) -> ONE cp.async.bulk.tensor store per
//     32x128-byte box.  The fp32 residual is TMA-LOADED into the same box ahead of time
//     (NBUF-deep per-warp ring, prefetched across tile boundaries), so the epilogue issues ~3
//     instructions per output element instead of ~14 and touches global memory only through TMA.
//   * GELU costs one MUFU and 8 issue slots (gelu_erf_tanhform) instead of two MUFU and ~16.
//   * producer and MMA warps run convergently with elected-lane predication (uniform-register operands).
//   * round 2: further epilogue modes on the same skeleton — hi/lo residual planes updated in place (RES = 3), GELU with the
//     pre-activation as a second output (ACT = 2), GELU adjoint with a TMA-loaded z tile (RES = 4), dropout + fp32 residual.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (leader CTA only) + TMEM owner, 2..9 = epilogue.
#include <cstdlib>
#include <type_traits>

#include "agb_common.cuh"

namespace agb {

constexpr int PG_BM = 128;   // accumulator rows per CTA (= TMEM lanes)
constexpr int PG_BN = 256;   // tile columns
constexpr int PG_BK = 64;    // K per pipeline stage (one 128-byte swizzle atom of bf16)
constexpr int PG_EPI_WARPS = 8;
constexpr int PG_THREADS = 64 + 32 * PG_EPI_WARPS;
constexpr int PG_BOX_BYTES = 4096;  // 32 rows x 128 bytes

struct PairGemmParams {
  int M, N, K;
  const float* bias;  // [N] fp32 or nullptr
  float alpha;
  int a_mn, b_mn;     // operand majorness (0 = K-major, 1 = MN-major)
  // LNIN: LayerNorm folded into this GEMM.  A holds the UN-normalised rows (bf16 copy of x), B = gamma-scaled
  // weights, bias already includes W beta; per-row statistics arrive as ln_parts partial (sum, sum of squares)
  // pairs and the epilogue applies  rstd_i * (acc_ij - mean_i * colsum_j) + bias_j.
  const float* ln_stats;   // [M][ln_parts][2]
  int ln_parts;
  const float* ln_colsum;  // [N]  sum_k B[j][k] (of the bf16-rounded, gamma-scaled weights)
  float ln_eps;
  // STATS: besides the fp32 output, emit a bf16 copy (tmOut2) and this GEMM's own per-row partial statistics
  float* stats_out;        // [M][2 * tiles_n][2]
  // split-K (wgrad: few output tiles, very long K): tile t = (split, m, n); split sp covers k-blocks
  // [sp * kb_per_split, ...) and stores its fp32 partial at output rows sp * M + m (reduced by splitk_reduce_kernel)
  int splits, kb_per_split;
  // fp32 residual epilogue with nn.Dropout on the GEMM output (training): out = residual + keep * (alpha acc + bias) / (1 - p),
  // keep regenerated from the counter hash of agb_dropout (same key, element index = row * N + col; contiguous output)
  unsigned drop_thr, drop_key;
  float drop_scale;
};

template <int CG, int STAGES, int NBUF, int ACT, int RES, int OUT_F32, int LNIN, int STATS>
__global__ void __launch_bounds__(PG_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmOut2, const PairGemmParams p) {
  // RES == 3 ("hi/lo"): the residual stream lives in HBM as two bf16 planes, x = hi + lo (hi = bf16(x), lo = bf16(x - hi):
  // 16 significant bits).  The epilogue TMA-loads both planes of a 32 x 64 chunk, adds the accumulator, re-splits and
  // stores both planes IN PLACE: the hi plane is at the same time the bf16 copy the next (LayerNorm-folded) GEMM reads as
  // its A operand, so the separate copy of the fp32 variant (2 of its 12 bytes per element) is never written.
  constexpr bool HL = RES == 3;
  // Training epilogues (bf16 output).  ACT == 2: GELU with the pre-activation as a second output (tmOut2): z for the adjoint,
  // GELU(z) for the next GEMM, from one pass over the accumulator.  RES == 4: the dgrad GEMM feeding a GELU adjoint,
  // out = acc * gelu'(z) with the z tile TMA-loaded like a residual (tmRes) and overwritten in place in its staging box.
  constexpr bool DUAL = ACT == 2;
  constexpr bool GBWD = RES == 4;
  static_assert(!DUAL || (OUT_F32 == 0 && RES == 0 && STATS == 0), "GELU + pre-activation: bf16 outputs, no residual");
  static_assert(!GBWD || (OUT_F32 == 0 && ACT == 0 && LNIN == 0 && STATS == 0), "GELU adjoint epilogue: bf16 output");
  static_assert(RES == 0 || HL || GBWD || OUT_F32 == 1, "fp32 residual pairs with fp32 output");
  static_assert(!HL || (OUT_F32 == 0 && STATS == 1 && ACT == 0 && LNIN == 0), "hi/lo residual: bf16 planes + statistics");
  static_assert(!STATS || (RES == 2 && OUT_F32 == 1) || HL, "statistics ride the residual epilogues");
  static_assert(!LNIN || (RES == 0 && OUT_F32 == 0), "folded LayerNorm feeds the bf16-output epilogues");
  constexpr int SLOT_BYTES = (HL || DUAL) ? 2 * PG_BOX_BYTES : PG_BOX_BYTES;   // one ring slot (two boxes back to back: hi / lo planes, GELU(z) / z)
  constexpr int NBOX = (HL || DUAL) ? 2 * NBUF : NBUF + (STATS ? 1 : 0);       // staging boxes per epilogue warp
  constexpr int B_ROWS = PG_BN / CG;                 // B rows staged by this CTA
  constexpr int A_BYTES = PG_BM * PG_BK * 2;
  constexpr int B_BYTES = B_ROWS * PG_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ATOM_BYTES = PG_BK * 128;            // one MN-major atom: 64 k-rows x 128 B
  constexpr int TMEM_COLS = 2 * PG_BN;
  constexpr int CHUNK_COLS = OUT_F32 ? 32 : 64;      // output columns per 128-byte staging row
  constexpr int NCHUNK = (PG_BN / 2) / CHUNK_COLS;   // chunks per warp per tile (warp owns 128 columns)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;                            // [EPI_WARPS][NBUF][4096]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + PG_EPI_WARPS * NBOX * PG_BOX_BYTES);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + STAGES;
  uint64_t* bar_tfull = bars + 2 * STAGES;
  uint64_t* bar_tempty = bars + 2 * STAGES + 2;
  uint64_t* bar_res = bars + 2 * STAGES + 4;                                  // [EPI_WARPS][NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_res + PG_EPI_WARPS * NBUF);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  const int pair = blockIdx.x / CG;
  const int num_pairs = gridDim.x / CG;
  const int tiles_m = (p.M + PG_BM * CG - 1) / (PG_BM * CG);
  const int tiles_n = (p.N + PG_BN - 1) / PG_BN;
  const int tiles_mn = tiles_m * tiles_n;
  const int num_tiles = tiles_mn * p.splits;
  const int num_kb = (p.K + PG_BK - 1) / PG_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (RES) tma_prefetch_desc(&tmRes);
    if (STATS || DUAL) tma_prefetch_desc(&tmOut2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), PG_EPI_WARPS * CG);
    }
    for (int s = 0; s < PG_EPI_WARPS * NBUF; ++s) mbar_init(smem_u32(&bar_res[s]), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish_pair(); }
    else         { tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);      tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs of a pair) ------------------------------
    // Whole warp runs the loop convergently; only the elected lane's instructions take effect.
    const uint32_t e = elect_one();
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
    uint32_t stage = 0, phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int sp = t / tiles_mn, tt = t - sp * tiles_mn;
      const int m0 = (tt / tiles_n) * (PG_BM * CG) + (int)cta_rank * PG_BM;
      const int n0 = (tt % tiles_n) * PG_BN + (int)cta_rank * B_ROWS;
      const int kb_end = min(num_kb, (sp + 1) * p.kb_per_split);
      for (int kb = sp * p.kb_per_split; kb < kb_end; ++kb) {
        mbar_wait(empty0 + stage * 8, phase ^ 1);
        const uint32_t full = full0 + stage * 8;
        if (leader) mbar_arrive_expect_tx_e(e, full, STAGE_BYTES * CG);   // both CTAs' bytes land on the leader
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        const int k0 = kb * PG_BK;
        if (CG == 2) {
          if (!p.a_mn) {
            tma_load_2d_pair_e(e, sa, &tmA, full, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < PG_BM / 64; ++j) tma_load_2d_pair_e(e, sa + j * ATOM_BYTES, &tmA, full, m0 + 64 * j, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_pair_e(e, sb, &tmB, full, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < B_ROWS / 64; ++j) tma_load_2d_pair_e(e, sb + j * ATOM_BYTES, &tmB, full, n0 + 64 * j, k0);
          }
        } else {
          if (!p.a_mn) {
            tma_load_2d_e(e, sa, &tmA, full, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < PG_BM / 64; ++j) tma_load_2d_e(e, sa + j * ATOM_BYTES, &tmA, full, m0 + 64 * j, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_e(e, sb, &tmB, full, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < B_ROWS / 64; ++j) tma_load_2d_e(e, sb + j * ATOM_BYTES, &tmB, full, n0 + 64 * j, k0);
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA) ------------------------------
    if (leader) {
      const uint32_t e = elect_one();
      const uint32_t idesc = make_idesc_bf16(PG_BM * CG, PG_BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? ATOM_BYTES : 16, a_kstep = (p.a_mn ? 2048 : 32) >> 4;
      const uint32_t b_lbo = p.b_mn ? ATOM_BYTES : 16, b_kstep = (p.b_mn ? 2048 : 32) >> 4;
      // descriptor = constant fields + (smem address >> 4); addresses stay below 2^18 so the add never carries
      const uint64_t da0 = make_smem_desc_sw128(0, a_lbo, 1024), db0 = make_smem_desc_sw128(0, b_lbo, 1024);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      const uint32_t tfull0 = smem_u32(&bar_tfull[0]), tempty0 = smem_u32(&bar_tempty[0]);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        mbar_wait(tempty0 + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PG_BN;
        const int sp = t / tiles_mn;
        const int kb_begin = sp * p.kb_per_split, kb_end = min(num_kb, (sp + 1) * p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t sa = (smem_base + stage * STAGE_BYTES) >> 4;
          const uint32_t sb = sa + (A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < PG_BK / 16; ++k) {
            umma_ss_e<CG>(e, d_tmem, da0 + (sa + k * a_kstep), db0 + (sb + k * b_kstep), idesc,
                          ((kb - kb_begin) | k) != 0 ? 1u : 0u);
          }
          umma_commit_e<CG>(e, empty0 + stage * 8);
          if (kb == kb_end - 1) umma_commit_e<CG>(e, tfull0 + acc * 8);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs) ------------------------------
    const int e = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch (rows q*32 .. +31)
    const int h = e >> 2;                   // column half (128 columns)
    const uint32_t stg_u32 = smem_u32(stg_base + e * NBOX * PG_BOX_BYTES);
    const uint32_t buf16 = stg_u32 + NBUF * PG_BOX_BYTES;     // STATS: bf16 copy box (64 columns x 32 rows)
    const uint32_t row_off = lane * 128;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint64_t* my_res = bar_res + e * NBUF;
    const uint32_t tempty_remote0 = (CG == 2) ? mapa_rank(smem_u32(&bar_tempty[0]), 0) : smem_u32(&bar_tempty[0]);

    // coordinates of this warp's running chunk g (tile sequence is static): returns false if the box is
    // entirely outside the output (M / N tails, the idle half of a pair on the last M tile)
    auto chunk_coords = [&](int g, int& row0, int& col0) -> bool {
      const int t = pair + (g / NCHUNK) * num_pairs;
      if (t >= num_tiles) return false;
      const int tt = t % tiles_mn;     // (residual loads are never combined with split-K)
      row0 = (tt / tiles_n) * (PG_BM * CG) + (int)cta_rank * PG_BM + q * 32;
      col0 = (tt % tiles_n) * PG_BN + h * (PG_BN / 2) + (g % NCHUNK) * CHUNK_COLS;
      return row0 < p.M && col0 < p.N;
    };
    auto issue_res = [&](int g) {  // lane 0 only
      int row0, col0;
      if (!chunk_coords(g, row0, col0)) return;
      const int b = g % NBUF;
      const uint32_t bar = smem_u32(&my_res[b]);
      mbar_arrive_expect_tx(bar, HL ? 2 * PG_BOX_BYTES : PG_BOX_BYTES);
      tma_load_2d(stg_u32 + b * SLOT_BYTES, &tmRes, bar, col0, row0);
      if (HL) tma_load_2d(stg_u32 + b * SLOT_BYTES + PG_BOX_BYTES, &tmOut2, bar, col0, row0);   // lo plane
    };

    if (RES && lane == 0) {
#pragma unroll
      for (int g = 0; g < NBUF; ++g) issue_res(g);
    }
    uint32_t res_phase = 0;  // bit b = parity of the next wait on my_res[b]
    uint32_t acc = 0, acc_phase = 0;
    const uint32_t tfull0 = smem_u32(&bar_tfull[0]);
    int g = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int sp = t / tiles_mn, tt = t - sp * tiles_mn;
      const int row0 = (tt / tiles_n) * (PG_BM * CG) + (int)cta_rank * PG_BM + q * 32;
      const int tcol0 = (tt % tiles_n) * PG_BN + h * (PG_BN / 2);
      const bool row_ok = row0 < p.M;
      const int out_row0 = row0 + sp * p.M;     // split-K partials are stacked along the rows of the workspace
      float ln_rstd = 1.f, ln_nmr = 0.f;      // LNIN: rstd_i and -mean_i * rstd_i of this thread's row
      float2 ln_part[LNIN ? 8 : 1];           // partial (sum, sumsq): loads are issued here, consumed after the TMEM wait
      if (LNIN) {
        const int my_row = row0 + lane;
        const float2* sp = reinterpret_cast<const float2*>(p.ln_stats) + (long long)min(my_row, p.M - 1) * p.ln_parts;
#pragma unroll
        for (int i = 0; i < 8; ++i) ln_part[i] = (i < p.ln_parts) ? __ldg(sp + i) : make_float2(0.f, 0.f);
      }
      float st_sum = 0.f, st_sq = 0.f;        // STATS: this thread's row over the warp's 128 columns
      mbar_wait(tfull0 + acc * 8, acc_phase);
      tc_fence_after();
      if (LNIN) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s1 += ln_part[i].x; s2 += ln_part[i].y; }
        const float invk = 1.0f / (float)p.K;
        const float mean = s1 * invk;
        const float var = fmaxf(fmaf(-mean, mean, s2 * invk), 0.f);
        ln_rstd = rsqrtf(var + p.ln_eps);
        ln_nmr = -mean * ln_rstd;
      }
      const uint32_t tm_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * PG_BN + h * (PG_BN / 2);
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c, ++g) {
        const int col0 = tcol0 + c * CHUNK_COLS;
        const bool active = row_ok && col0 < p.N;
        const int b = g % NBUF;
        const uint32_t buf = stg_u32 + b * SLOT_BYTES;
        uint32_t r[CHUNK_COLS];
        if (active) {
          tmem_ld32(tm_addr + c * CHUNK_COLS, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
          if (!OUT_F32)
            tmem_ld32(tm_addr + c * CHUNK_COLS + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[CHUNK_COLS - 32]));
          if (RES) {
            mbar_wait(smem_u32(&my_res[b]), (res_phase >> b) & 1u);
            res_phase ^= 1u << b;
          } else {
            if (lane == 0) bulk_wait_read<NBUF - 1>();   // the store that last used this box has read it
            __syncwarp();
          }
          tmem_wait_ld();
        }
        if (c == NCHUNK - 1) {
          // every TMEM read of this accumulator stage has landed in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(tempty_remote0 + acc * 8);
            else         mbar_arrive(tempty_remote0 + acc * 8);
          }
        }
        if (active) {
          // bias / colsum columns past N are clamped to a valid address (those outputs are clipped by the TMA store);
          // full chunks — every chunk of the hot shapes — take the copy of the loop with immediate offsets
          const float* bias_c = p.bias + col0;
          const int n_last = p.N - 4 - col0;
          auto chunk_math = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t addr = buf + row_off + ((static_cast<uint32_t>(j) ^ sw) << 4);
            if constexpr (HL) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_c + (FULL ? 8 * j : min(8 * j, n_last))));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_c + (FULL ? 8 * j + 4 : min(8 * j + 4, n_last))));
              uint32_t hw[4], lw[4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(hw[0]), "=r"(hw[1]), "=r"(hw[2]), "=r"(hw[3]) : "r"(addr) : "memory");
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(lw[0]), "=r"(lw[1]), "=r"(lw[2]), "=r"(lw[3]) : "r"(addr + PG_BOX_BYTES) : "memory");
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t nh[4], nl[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                // hi + lo is exact in fp32 (16 significant bits); one rounding for the sum with the accumulator
                const float v0 = fmaf(__uint_as_float(r[8 * j + 2 * i + 0]), p.alpha, bb[2 * i + 0]) + (bf16_lo(hw[i]) + bf16_lo(lw[i]));
                const float v1 = fmaf(__uint_as_float(r[8 * j + 2 * i + 1]), p.alpha, bb[2 * i + 1]) + (bf16_hi(hw[i]) + bf16_hi(lw[i]));
                st_sum += v0 + v1;
                st_sq = fmaf(v0, v0, fmaf(v1, v1, st_sq));
                nh[i] = pack_bf16x2(v0, v1);
                nl[i] = pack_bf16x2(v0 - bf16_lo(nh[i]), v1 - bf16_hi(nh[i]));
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(nh[0]), "r"(nh[1]), "r"(nh[2]), "r"(nh[3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + PG_BOX_BYTES), "r"(nl[0]), "r"(nl[1]), "r"(nl[2]),
                           "r"(nl[3])
                           : "memory");
            } else if (OUT_F32) {
              const float4 bias4 = __ldg(reinterpret_cast<const float4*>(bias_c + (FULL ? 4 * j : min(4 * j, n_last))));
              float4 v;
              v.x = fmaf(__uint_as_float(r[4 * j + 0]), p.alpha, bias4.x);
              v.y = fmaf(__uint_as_float(r[4 * j + 1]), p.alpha, bias4.y);
              v.z = fmaf(__uint_as_float(r[4 * j + 2]), p.alpha, bias4.z);
              v.w = fmaf(__uint_as_float(r[4 * j + 3]), p.alpha, bias4.w);
              if (ACT == 1) {
                v.x = gelu_erf_tanhform(v.x); v.y = gelu_erf_tanhform(v.y);
                v.z = gelu_erf_tanhform(v.z); v.w = gelu_erf_tanhform(v.w);
              }
              if (RES == 2 && !STATS && p.drop_thr != 0u) {
                const uint32_t pair0 = (static_cast<uint32_t>(row0 + lane) * static_cast<uint32_t>(p.N) +
                                        static_cast<uint32_t>(col0 + 4 * j)) >> 1;
                const uint32_t x0 = agb_drop_bits(p.drop_key, pair0), x1 = agb_drop_bits(p.drop_key, pair0 + 1u);
                v.x = (x0 & 0xFFFFu) >= p.drop_thr ? v.x * p.drop_scale : 0.f;
                v.y = (x0 >> 16) >= p.drop_thr ? v.y * p.drop_scale : 0.f;
                v.z = (x1 & 0xFFFFu) >= p.drop_thr ? v.z * p.drop_scale : 0.f;
                v.w = (x1 >> 16) >= p.drop_thr ? v.w * p.drop_scale : 0.f;
              }
              if (RES) {
                float4 rr;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(rr.x), "=f"(rr.y), "=f"(rr.z), "=f"(rr.w)
                             : "r"(addr)
                             : "memory");
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
              }
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                           "f"(v.w)
                           : "memory");
              if (STATS) {
                st_sum += (v.x + v.y) + (v.z + v.w);
                st_sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, st_sq))));
                // bf16 copy: chunk c fills the left (even c) or right (odd c) 64 bytes of the 128-byte row;
                // 8-byte piece j of this chunk = columns 4j..4j+3
                const uint32_t pos = static_cast<uint32_t>((c & 1) * 64 + j * 8);              // byte offset in the row
                const uint32_t a16 = buf16 + row_off + ((((pos >> 4) ^ sw) << 4) | (pos & 8u));
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a16), "r"(pack_bf16x2(v.x, v.y)),
                             "r"(pack_bf16x2(v.z, v.w))
                             : "memory");
              }
            } else {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_c + (FULL ? 8 * j : min(8 * j, n_last))));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_c + (FULL ? 8 * j + 4 : min(8 * j + 4, n_last))));
              float v[8];
              if (LNIN) {
                const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + (FULL ? 8 * j : min(8 * j, n_last))));
                const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + (FULL ? 8 * j + 4 : min(8 * j + 4, n_last))));
                v[0] = fmaf(__uint_as_float(r[8 * j + 0]), ln_rstd, fmaf(ln_nmr, c0.x, b0.x));
                v[1] = fmaf(__uint_as_float(r[8 * j + 1]), ln_rstd, fmaf(ln_nmr, c0.y, b0.y));
                v[2] = fmaf(__uint_as_float(r[8 * j + 2]), ln_rstd, fmaf(ln_nmr, c0.z, b0.z));
                v[3] = fmaf(__uint_as_float(r[8 * j + 3]), ln_rstd, fmaf(ln_nmr, c0.w, b0.w));
                v[4] = fmaf(__uint_as_float(r[8 * j + 4]), ln_rstd, fmaf(ln_nmr, c1.x, b1.x));
                v[5] = fmaf(__uint_as_float(r[8 * j + 5]), ln_rstd, fmaf(ln_nmr, c1.y, b1.y));
                v[6] = fmaf(__uint_as_float(r[8 * j + 6]), ln_rstd, fmaf(ln_nmr, c1.z, b1.z));
                v[7] = fmaf(__uint_as_float(r[8 * j + 7]), ln_rstd, fmaf(ln_nmr, c1.w, b1.w));
              } else {
              v[0] = fmaf(__uint_as_float(r[8 * j + 0]), p.alpha, b0.x);
              v[1] = fmaf(__uint_as_float(r[8 * j + 1]), p.alpha, b0.y);
              v[2] = fmaf(__uint_as_float(r[8 * j + 2]), p.alpha, b0.z);
              v[3] = fmaf(__uint_as_float(r[8 * j + 3]), p.alpha, b0.w);
              v[4] = fmaf(__uint_as_float(r[8 * j + 4]), p.alpha, b1.x);
              v[5] = fmaf(__uint_as_float(r[8 * j + 5]), p.alpha, b1.y);
              v[6] = fmaf(__uint_as_float(r[8 * j + 6]), p.alpha, b1.z);
              v[7] = fmaf(__uint_as_float(r[8 * j + 7]), p.alpha, b1.w);
              }
              if constexpr (GBWD) {
                uint32_t zw[4];
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(zw[0]), "=r"(zw[1]), "=r"(zw[2]), "=r"(zw[3]) : "r"(addr) : "memory");
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v[2 * i + 0] *= gelu_grad_tanhform(bf16_lo(zw[i]));
                  v[2 * i + 1] *= gelu_grad_tanhform(bf16_hi(zw[i]));
                }
              }
              if constexpr (DUAL) {
                // the adjoint differentiates at the ROUNDED pre-activation, as the unfused path (bf16 z -> gelu kernel) does
                const uint32_t z0 = pack_bf16x2(v[0], v[1]), z1 = pack_bf16x2(v[2], v[3]), z2 = pack_bf16x2(v[4], v[5]),
                               z3 = pack_bf16x2(v[6], v[7]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + PG_BOX_BYTES), "r"(z0), "r"(z1), "r"(z2), "r"(z3)
                             : "memory");
                v[0] = bf16_lo(z0); v[1] = bf16_hi(z0); v[2] = bf16_lo(z1); v[3] = bf16_hi(z1);
                v[4] = bf16_lo(z2); v[5] = bf16_hi(z2); v[6] = bf16_lo(z3); v[7] = bf16_hi(z3);
              }
              if (ACT != 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = gelu_erf_tanhform(v[i]);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[0], v[1])),
                           "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
                           : "memory");
            }
          }
          };
          if (col0 + CHUNK_COLS <= p.N) chunk_math(std::true_type{});
          else chunk_math(std::false_type{});
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmOut, buf, col0, out_row0);
            if (HL || DUAL) tma_store_2d(&tmOut2, buf + PG_BOX_BYTES, col0, row0);         // lo plane / pre-activation
            else if (STATS && (c & 1)) tma_store_2d(&tmOut2, buf16, col0 - CHUNK_COLS, row0);   // 64 bf16 columns
            bulk_commit();
            if (RES) bulk_wait_read<0>();   // boxes handed to the store engine: free for the next residual / copy
          }
        }
        if (RES && lane == 0) issue_res(g + NBUF);   // residual NBUF chunks ahead (crosses tile boundaries)
        __syncwarp();
      }
      if (STATS) {
        const int my_row = row0 + lane;
        if (my_row < p.M)
          reinterpret_cast<float2*>(p.stats_out)[(long long)my_row * (2 * tiles_n) + (tt % tiles_n) * 2 + h] =
              make_float2(st_sum, st_sq);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else         tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int CG, int STAGES, int NBUF, int STATS, int HL = 0>
constexpr int pair_smem_bytes() {
  return STAGES * (PG_BM * PG_BK * 2 + (PG_BN / CG) * PG_BK * 2) + PG_EPI_WARPS * (HL ? 2 * NBUF : NBUF + STATS) * PG_BOX_BYTES +   // HL: any two-box slot
         (2 * STAGES + 4 + PG_EPI_WARPS * NBUF) * 8 + 16 + 1024;
}

template <int CG, int STAGES, int NBUF, int ACT, int RES, int OUT_F32, int LNIN, int STATS>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                       const CUtensorMap& tmRes, const CUtensorMap& tmOut2, const PairGemmParams& p,
                       cudaStream_t stream) {
  constexpr int SMEM = pair_smem_bytes<CG, STAGES, NBUF, STATS, (RES == 3 || ACT == 2)>();
  static_assert(SMEM <= 232448, "shared memory budget");
  auto kern = gemm_pair_kernel<CG, STAGES, NBUF, ACT, RES, OUT_F32, LNIN, STATS>;
  static int max_pairs = 0;   // co-resident CTAs (CG = 1) or clusters (CG = 2)
  if (max_pairs == 0) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (CG == 2) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(sm_count(), 1, 1);
      qc.blockDim = dim3(PG_THREADS, 1, 1);
      qc.dynamicSmemBytes = SMEM;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      qc.attrs = at; qc.numAttrs = 1;
      int n = 0;
      AGB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &qc));
      AGB_REQUIRE(n > 0, "no co-resident CTA pair fits");
      max_pairs = n < sm_count() / 2 ? n : sm_count() / 2;
    } else {
      max_pairs = sm_count();
    }
  }
  const int tiles = ((p.M + PG_BM * CG - 1) / (PG_BM * CG)) * ((p.N + PG_BN - 1) / PG_BN) * p.splits;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pairs * CG, 1, 1);
  cfg.blockDim = dim3(PG_THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  AGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmRes, tmOut2, p));
  return AGB_OK;
}

// fixed-order sum of the split-K partials: out[m][n] = sum_sp ws[sp][m][n]  (float4 per thread; deterministic)
__global__ void splitk_reduce_kernel(const float4* __restrict__ ws, int splits, long long mn4, long long ld4,
                                     long long n4, float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn4) return;
  float4 a = ws[i];
  for (int s = 1; s < splits; ++s) {
    const float4 b = ws[(long long)s * mn4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[(i / n4) * ld4 + (i % n4)] = a;
}

static float* g_splitk_ws = nullptr;      // grow-only workspace (single-stream use, like the rest of the library)
static size_t g_splitk_ws_bytes = 0;

static int g_gemm_variant = 0;  // 0 auto, 1 legacy kernel only, 2 force CG=1, 3 force CG=2
static int g_hl_cfg = [] { const char* e = getenv("AGB_GEMM_HILO_CFG"); return e ? atoi(e) : 0; }();   // A/B switch, see below
void set_gemm_variant(int v) { g_gemm_variant = v; }
int get_gemm_variant() { return g_gemm_variant; }

// Returns AGB_ERR_UNSUPPORTED when this kernel does not cover the request (the caller then uses the
// first-generation kernel): narrow N, bf16 residual, fp32 residual with bf16 output, GELU + residual.
// Field meanings: PairGemmCall in agb_common.cuh.
// Optional fusions (nullptr = off):
//   ln_stats / ln_colsum : LayerNorm of the A rows folded into the epilogue (bf16 output only, no residual)
//   out16 / stats_out    : with an fp32 residual epilogue, also emit a bf16 copy of the output and its per-row
//                          partial (sum, sum of squares) over each 128-column slab: stats_out [M][2*ceil(N/256)][2]
int gemm_bf16_pair_call(const PairGemmCall& call, cudaStream_t stream) {
  const bf16* A = call.A; const int lda = call.lda, a_mn = call.a_mn;
  const bf16* B = call.B; const int ldb = call.ldb, b_mn = call.b_mn;
  const int M = call.M, N = call.N, K = call.K;
  const float alpha = call.alpha;
  const float* bias = call.bias;
  const int act = call.act;
  const bf16* res_bf16 = call.res_bf16;
  const float* res_f32 = call.res_f32; const int ldr = call.ldr;
  void* out = call.out; int ldo = call.ldo, out_f32 = call.out_f32;
  const float* ln_stats = call.ln_stats; const int ln_parts = call.ln_parts;
  const float* ln_colsum = call.ln_colsum; const float ln_eps = call.ln_eps;
  bf16* out16 = call.out16; const int ldo16 = call.ldo16;
  float* stats_out = call.stats_out;
  bf16* hl_hi = call.hl_hi; bf16* hl_lo = call.hl_lo; const int ld_hl = call.ld_hl;
  const unsigned drop_thr = call.drop_thr, drop_key = call.drop_key;
  bf16* z_out = call.z_out; const int ldz_out = call.ldz_out;
  const bf16* z_in = call.z_in; const int ldz_in = call.ldz_in;
  const bool lnin = ln_stats != nullptr, stats = stats_out != nullptr;
  const bool hl = hl_hi != nullptr;
  if (g_gemm_variant == 1) return AGB_ERR_UNSUPPORTED;
  if (N < 192 || res_bf16 != nullptr) return AGB_ERR_UNSUPPORTED;
  if (hl) {   // hi/lo residual planes, updated in place: x = hi + lo  <-  x + alpha A B^T + bias
    if (hl_lo == nullptr || res_f32 != nullptr || act != 0 || lnin || !stats || (N % PG_BN) != 0 || (ld_hl % 8) != 0 ||
        (reinterpret_cast<uintptr_t>(hl_hi) & 15) != 0 || (reinterpret_cast<uintptr_t>(hl_lo) & 15) != 0)
      return AGB_ERR_UNSUPPORTED;
    out = hl_hi; ldo = ld_hl; out_f32 = 0;
  }
  if (res_f32 != nullptr && (!out_f32 || act != 0)) return AGB_ERR_UNSUPPORTED;
  const int oes = out_f32 ? 4 : 2;
  if (((long long)ldo * oes) % 16 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return AGB_ERR_UNSUPPORTED;
  if (res_f32 && ((((long long)ldr * 4) % 16) != 0 || (reinterpret_cast<uintptr_t>(res_f32) & 15) != 0))
    return AGB_ERR_UNSUPPORTED;
  if (lnin && (out_f32 || res_f32 || a_mn || ln_colsum == nullptr || ln_parts <= 0 || ln_parts > 8 || alpha != 1.0f ||
               (reinterpret_cast<uintptr_t>(ln_colsum) & 15) != 0))
    return AGB_ERR_UNSUPPORTED;
  // training epilogues: act == 2 (GELU + pre-activation z_out), z_in (out = acc * gelu'(z_in)); bf16 outputs, nothing else fused
  const bool dual = act == 2, gbwd = z_in != nullptr;
  if (dual && (z_out == nullptr || out_f32 || res_f32 || lnin || stats || hl || (ldz_out % 8) != 0 ||
               (reinterpret_cast<uintptr_t>(z_out) & 15) != 0))
    return AGB_ERR_UNSUPPORTED;
  if (gbwd && (act != 0 || out_f32 || res_f32 || lnin || stats || hl || bias != nullptr || (ldz_in % 8) != 0 ||
               (reinterpret_cast<uintptr_t>(z_in) & 15) != 0))
    return AGB_ERR_UNSUPPORTED;
  if (act != 0 && act != 1 && !dual) return AGB_ERR_UNSUPPORTED;
  if (drop_thr != 0u && (!res_f32 || !out_f32 || stats || lnin || hl || ldo != N || (N % 4) != 0 ||
                         (long long)M * N >= (1ll << 32)))
    return AGB_ERR_UNSUPPORTED;
  if (stats && !hl && (!res_f32 || out16 == nullptr || (N % PG_BN) != 0 || (ldo16 % 8) != 0 ||
                       (reinterpret_cast<uintptr_t>(out16) & 15) != 0))
    return AGB_ERR_UNSUPPORTED;

  // CTA pairs pay off once there is at least ~one full wave of 256-row tiles; small problems keep CG = 1
  const long long tiles_pair = (long long)((M + 255) / 256) * ((N + PG_BN - 1) / PG_BN);
  int cg = tiles_pair >= 2 * (sm_count() / 2) ? 2 : 1;
  if (g_gemm_variant == 2) cg = 1;
  if (g_gemm_variant == 3) cg = 2;

  // split-K: few output tiles and a very long K (the wgrad GEMMs dW = dY^T X with K = rows of the batch)
  int splits = 1;
  const int num_kb = (K + PG_BK - 1) / PG_BK;
  void* final_out = out;
  const int final_ldo = ldo;
  if (cg == 1 && out_f32 && !res_f32 && !hl && !lnin && !stats && act == 0 && bias == nullptr && alpha == 1.0f &&
      (M % PG_BM) == 0 && (N % 4) == 0 && (ldo % 4) == 0 && num_kb >= 32 && g_gemm_variant == 0) {
    const long long tiles1 = (long long)(M / PG_BM) * ((N + PG_BN - 1) / PG_BN);
    int want = (int)(sm_count() / tiles1);
    if (want > 16) want = 16;
    if (want > num_kb / 8) want = num_kb / 8;
    if (want >= 2) {
      const size_t need = (size_t)want * M * N * sizeof(float);
      if (need > g_splitk_ws_bytes && need <= ((size_t)1 << 30)) {
        if (g_splitk_ws != nullptr) {
          AGB_CHECK_CUDA(cudaStreamSynchronize(stream));
          AGB_CHECK_CUDA(cudaFree(g_splitk_ws));
        }
        AGB_CHECK_CUDA(cudaMalloc(&g_splitk_ws, need));
        g_splitk_ws_bytes = need;
      }
      if (need <= g_splitk_ws_bytes) {
        splits = want;
        out = g_splitk_ws;
        ldo = N;
      }
    }
  }
  int kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per_split - 1) / kb_per_split;      // no empty splits

  CUtensorMap tmA, tmB, tmOut, tmRes, tmOut2;
  int rc;
  if (!a_mn) rc = encode_tmap_2d_bf16(&tmA, A, K, M, (uint64_t)lda * 2, PG_BK, PG_BM);
  else       rc = encode_tmap_2d_bf16(&tmA, A, M, K, (uint64_t)lda * 2, 64, PG_BK);
  if (rc != AGB_OK) return rc;
  if (!b_mn) rc = encode_tmap_2d_bf16(&tmB, B, K, N, (uint64_t)ldb * 2, PG_BK, PG_BN / cg);
  else       rc = encode_tmap_2d_bf16(&tmB, B, N, K, (uint64_t)ldb * 2, 64, PG_BK);
  if (rc != AGB_OK) return rc;
  rc = encode_tmap_2d(&tmOut, out, oes, N, (uint64_t)M * splits, (uint64_t)ldo * oes, out_f32 ? 32 : 64, 32);
  if (rc != AGB_OK) return rc;
  if (res_f32)   rc = encode_tmap_2d(&tmRes, res_f32, 4, N, M, (uint64_t)ldr * 4, 32, 32);
  else if (gbwd) rc = encode_tmap_2d(&tmRes, z_in, 2, N, M, (uint64_t)ldz_in * 2, 64, 32);
  else           tmRes = tmOut;               // hi/lo: the hi plane is residual and output
  if (rc != AGB_OK) return rc;
  if (hl)         rc = encode_tmap_2d(&tmOut2, hl_lo, 2, N, M, (uint64_t)ld_hl * 2, 64, 32);
  else if (dual)  rc = encode_tmap_2d(&tmOut2, z_out, 2, N, M, (uint64_t)ldz_out * 2, 64, 32);
  else if (stats) rc = encode_tmap_2d(&tmOut2, out16, 2, N, M, (uint64_t)ldo16 * 2, 64, 32);
  else            tmOut2 = tmOut;
  if (rc != AGB_OK) return rc;

  if (bias == nullptr) {   // the epilogue reads bias unconditionally: substitute zeros
    static float* zero_bias = nullptr;
    constexpr int ZB = 16384;
    if (N > ZB) return AGB_ERR_UNSUPPORTED;
    if (zero_bias == nullptr) {
      AGB_CHECK_CUDA(cudaMalloc(&zero_bias, ZB * sizeof(float)));
      AGB_CHECK_CUDA(cudaMemset(zero_bias, 0, ZB * sizeof(float)));
    }
    bias = zero_bias;
  }
  PairGemmParams p;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.alpha = alpha; p.a_mn = a_mn; p.b_mn = b_mn;
  p.splits = splits; p.kb_per_split = kb_per_split;
  p.ln_stats = ln_stats; p.ln_parts = ln_parts; p.ln_colsum = ln_colsum; p.ln_eps = ln_eps; p.stats_out = stats_out;
  p.drop_thr = drop_thr; p.drop_key = drop_key; p.drop_scale = 65536.0f / (65536.0f - (float)drop_thr);
  const int res = hl ? 3 : (res_f32 ? 2 : 0);
  if (dual) {   // two-box slots: 5 stages (160 KB) + one slot per warp (64 KB)
    if (cg == 2) return launch_pair<2, 5, 1, 2, 0, 0, 0, 0>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
    return launch_pair<1, 3, 1, 2, 0, 0, 0, 0>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
  }
  if (gbwd) {
    if (cg == 2) {
      if (K < 2048) return launch_pair<2, 4, 2, 0, 4, 0, 0, 0>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
      return launch_pair<2, 5, 2, 0, 4, 0, 0, 0>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
    }
    return launch_pair<1, 3, 2, 0, 4, 0, 0, 0>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
  }
  if (hl) {
    // ring slots are 8 KB here (both planes of a 32 x 64 chunk).  Short K (the HBM-bound out-projection): 3 stages + two
    // slots per warp (342-345 us in-step at the bench shape), or (AGB_GEMM_HILO_CFG=1) 4 stages + one slot (350-368 us);
    // long K (FC2): 5 stages + one slot.  An L2 prefetch of the residual chunks (cp.async.bulk.prefetch.tensor) 2-16 chunks
    // ahead of their TMA load was measured and made every residual GEMM SLOWER (profiles/r02_hilo_residual_ab.txt).
    if (cg == 2) {
      if (K < 2048) {
        if (g_hl_cfg == 1) return launch_pair<2, 4, 1, 0, 3, 0, 0, 1>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
        return launch_pair<2, 3, 2, 0, 3, 0, 0, 1>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
      }
      return launch_pair<2, 5, 1, 0, 3, 0, 0, 1>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
    }
    return launch_pair<1, 3, 1, 0, 3, 0, 0, 1>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);
  }
  // (stages, residual ring) per configuration: CTA pairs stage 32 KB per k-block (5 stages: the K = 3072 GEMM was
  // latency-starved with 4), single CTAs 48 KB (3 stages); the statistics variant trades one ring slot for the copy box
  auto finish = [&](int lrc) -> int {
    if (lrc != AGB_OK || splits == 1) return lrc;
    const long long mn4 = (long long)M * N / 4;
    splitk_reduce_kernel<<<(unsigned)((mn4 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(g_splitk_ws), splits, mn4, final_ldo / 4, N / 4, reinterpret_cast<float4*>(final_out));
    AGB_CHECK_CUDA(cudaGetLastError());
    return AGB_OK;
  };
#define AGB_PAIR_CASE(A_, R_, O_, L_, S_)                                                                          \
  if (act == A_ && res == R_ && out_f32 == O_ && (int)lnin == L_ && (int)stats == S_) {                            \
    if (cg == 2)                                                                                                   \
    {                                                                                                              \
      /* long-K residual GEMMs are latency-starved with 4 stages (5 stages, short residual ring); short-K ones are   \
         HBM-bound and want the deeper residual prefetch instead (4 stages, ring of 2 + copy box / ring of 3) */   \
      if (R_ && K < 2048)                                                                                          \
        return launch_pair<2, (R_ ? 4 : 5), (R_ ? (S_ ? 2 : 3) : 2), A_, R_, O_, L_, S_>(tmA, tmB, tmOut, tmRes, tmOut2, p, \
                                                                                         stream);                  \
      return launch_pair<2, 5, (S_ ? 1 : 2), A_, R_, O_, L_, S_>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream);       \
    }                                                                                                              \
    return finish(launch_pair<1, 3, (S_ ? 1 : 2), A_, R_, O_, L_, S_>(tmA, tmB, tmOut, tmRes, tmOut2, p, stream)); \
  }
  AGB_PAIR_CASE(0, 0, 0, 0, 0)
  AGB_PAIR_CASE(0, 0, 1, 0, 0)
  AGB_PAIR_CASE(0, 2, 1, 0, 0)
  AGB_PAIR_CASE(1, 0, 0, 0, 0)
  AGB_PAIR_CASE(1, 0, 1, 0, 0)
  AGB_PAIR_CASE(0, 0, 0, 1, 0)
  AGB_PAIR_CASE(1, 0, 0, 1, 0)
  AGB_PAIR_CASE(0, 2, 1, 0, 1)
#undef AGB_PAIR_CASE
  return AGB_ERR_UNSUPPORTED;
}

int gemm_bf16_pair(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N, int K,
                   float alpha, const float* bias, int act, const bf16* res_bf16, const float* res_f32, int ldr,
                   void* out, int ldo, int out_f32, cudaStream_t stream) {
  PairGemmCall c;
  c.A = A; c.lda = lda; c.a_mn = a_mn; c.B = B; c.ldb = ldb; c.b_mn = b_mn; c.M = M; c.N = N; c.K = K;
  c.alpha = alpha; c.bias = bias; c.act = act; c.res_bf16 = res_bf16; c.res_f32 = res_f32; c.ldr = ldr;
  c.out = out; c.ldo = ldo; c.out_f32 = out_f32;
  return gemm_bf16_pair_call(c, stream);
}

}  // namespace agb

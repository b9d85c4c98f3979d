// Fused key-masked attention on tcgen05 / TMEM / TMA (bf16 in/out, head dim 64, T <= 256).
//
// Replaces the reference's 6-kernel score pipeline (matmul, /sqrt(d), *mask or +mask, softmax,
// dropout(eval), matmul; models/vanilla_vit.py:444-463 and models/vanilla_bert.py:517-537): the
// (N,h,T,T) score tensor never leaves the SM and the masked copies of the input are never built.
//
// How the coalition bitmask enters (no per-element masking in the softmax loop at all):
//   * S = Q K^T is a tcgen05.mma with K staged by TMA; the rows of K whose key bit is 0 are ZEROED in
//     shared memory first, so a masked key's logit is exactly 0 — the ViT semantics ("scores * mask",
//     masked keys keep weight exp(0)/Z and their V rows still count).
//   * the softmax denominator comes out of the second MMA: V is extended by a "ones" column
//     (N = 64 + 16), so O[:,64] = sum_j P_ij * ones_j with the SAME bf16-rounded P that multiplies V.
//     ones_j = 1 for every real key (ViT) or only for kept keys (BERT); for BERT the masked V rows are
//     zeroed as well, which is exactly "probability 0" (additive finfo.min).  Zero-padded key rows
//     (j >= T) have ones_j = 0 and V_j = 0 and therefore drop out by themselves.
//   * P (bf16) is written back into TMEM over S and fed to the second MMA as the A operand from
//     tensor memory (tcgen05.mma ... [tmem_a]) — no shared-memory round trip for P.
// One CTA = one (row, head) unit at a time, 1 TMA warp + 1 MMA warp + 4 softmax/epilogue warps
// (thread = query row = TMEM lane).  The CTA is sequential inside; two CTAs are co-resident per SM
// (<= 110 KB smem, 256 TMEM columns each) so one CTA's MUFU-bound softmax overlaps the other's
// TMA/MMA phases.
#include "agb_common.cuh"

namespace agb {

constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_TMEM_COLS = 256;
constexpr int ATT_O_COL = 128;   // O accumulator lives at columns [128, 208) (overlays dead S columns)
constexpr int ATT_NV = 80;       // 64 value columns + 16 (ones column padded to the N%16 rule)

struct AttParams {
  const uint32_t* mask;
  int words;
  int rows, T, H, heads, mode;
  int NK;            // keys padded to a multiple of 16
  int units;         // rows * heads
  int mtiles;        // ceil(T / 128)
  bf16* ctx;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int kv_bytes = p.NK * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 128 * 128;
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sV1 = sV + kv_bytes;          // second MN atom of V: column 64 = ones, 65..79 = 0
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV1 + kv_bytes);
  uint64_t* bar_load = bars + 0;
  uint64_t* bar_prep = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint64_t* bar_free = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(smem_u32(bar_load), 1);
    mbar_init(smem_u32(bar_prep), 128);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_p), 128);
    mbar_init(smem_u32(bar_o), 1);
    mbar_init(smem_u32(bar_free), 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), ATT_TMEM_COLS);
    tmem_relinquish();
  }
  // zero the ones/padding atom once (TMA never writes it)
  for (int i = threadIdx.x * 16; i < kv_bytes; i += ATT_THREADS * 16)
    *reinterpret_cast<uint4*>(sV1 + i) = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int H = p.H, T = p.T, NK = p.NK;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int row = u / p.heads, head = u % p.heads;
        for (int mt = 0; mt < p.mtiles; ++mt, ++it) {
          if (it > 0) {
            mbar_wait(smem_u32(bar_s), (it - 1) & 1);                 // Q buffer consumed
            if (mt == 0) mbar_wait(smem_u32(bar_o), (it - 1) & 1);    // K/V consumed
          }
          const uint32_t bl = smem_u32(bar_load);
          mbar_arrive_expect_tx(bl, 128 * 128 + (mt == 0 ? 2 * kv_bytes : 0));
          tma_load_3d(smem_u32(sQ), &tmQ, bl, head * ATT_D, mt * 128, row);
          if (mt == 0) {
            tma_load_3d(smem_u32(sK), &tmKV, bl, H + head * ATT_D, 0, row);
            tma_load_3d(smem_u32(sV), &tmKV, bl, 2 * H + head * ATT_D, 0, row);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc_s = make_idesc_bf16(128, NK, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, ATT_NV, 0, 1);
    const uint32_t v_lbo = static_cast<uint32_t>(sV1 - sV);
    int it = 0, un = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++un) {
      for (int mt = 0; mt < p.mtiles; ++mt, ++it) {
        mbar_wait(smem_u32(bar_load), it & 1);
        if (mt == 0) mbar_wait(smem_u32(bar_prep), un & 1);
        if (it > 0) mbar_wait(smem_u32(bar_free), (it - 1) & 1);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(smem_u32(sK) + k * 32, 16, 1024);
            umma_ss(tmem_base, da, db, idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(bar_s));
        }
        __syncwarp();
        mbar_wait(smem_u32(bar_p), it & 1);
        tc_fence_after();
        if (lane == 0) {
          for (int ks = 0; ks < NK / 16; ++ks) {
            const uint64_t db = make_smem_desc_sw128(smem_u32(sV) + ks * 2048, v_lbo, 1024);
            umma_ts(tmem_base + ATT_O_COL, tmem_base + ks * 8, db, idesc_o, ks != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(bar_o));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ softmax + epilogue (128 threads) ------------------------------
    const int q = warp & 3;                         // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;                    // query row within the tile = TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float scale_log2 = 0.125f * 1.4426950408889634f;
    int it = 0, un = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++un) {
      const int row = u / p.heads, head = u % p.heads;
      const uint32_t* mrow = p.mask + (long long)row * p.words;
      for (int mt = 0; mt < p.mtiles; ++mt, ++it) {
        if (mt == 0) {
          mbar_wait(smem_u32(bar_load), it & 1);
          // apply the coalition mask to the staged K / V rows of this unit
          for (int j = r; j < NK; j += 128) {
            const bool real = j < T;
            const bool keep = real && ((mrow[j >> 5] >> (j & 31)) & 1u);
            if (real && !keep) {
              uint4* kr = reinterpret_cast<uint4*>(sK + j * 128);
#pragma unroll
              for (int c = 0; c < 8; ++c) kr[c] = make_uint4(0, 0, 0, 0);
              if (p.mode == AGB_MASK_NEGINF) {
                uint4* vr = reinterpret_cast<uint4*>(sV + j * 128);
#pragma unroll
                for (int c = 0; c < 8; ++c) vr[c] = make_uint4(0, 0, 0, 0);
              }
            }
            const bool counts = (p.mode == AGB_MASK_NEGINF) ? keep : real;
            // logical 16-byte chunk 0 of row j (value columns 64..71) sits at physical chunk (j & 7)
            *reinterpret_cast<uint4*>(sV1 + j * 128 + ((j & 7) << 4)) =
                make_uint4(counts ? 0x00003F80u : 0u, 0, 0, 0);
          }
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(bar_prep));
        }
        mbar_wait(smem_u32(bar_s), it & 1);
        tc_fence_after();
        // pass 1: row maximum of the raw scores (masked / padded keys contribute their exact 0)
        float m = -INFINITY;
        for (int c0 = 0; c0 < NK; c0 += 16) {
          uint32_t s[16];
          tmem_ld16(lane_addr + c0, s);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(s[j]));
        }
        const float m_scaled = m * scale_log2;
        // pass 2: P = exp2(s*scale - m*scale) -> bf16 pairs -> TMEM (overlaying consumed S columns)
        for (int c0 = 0; c0 < NK; c0 += 16) {
          uint32_t s[16];
          tmem_ld16(lane_addr + c0, s);
          tmem_wait_ld();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(s[2 * j]), scale_log2, -m_scaled));
            const float e1 = ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), scale_log2, -m_scaled));
            pk[j] = pack_bf16x2(e0, e1);
          }
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(
                  lane_addr + (c0 >> 1)),
              "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
              : "memory");
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(smem_u32(bar_p));
        // epilogue: O / rowsum -> bf16 ctx
        mbar_wait(smem_u32(bar_o), it & 1);
        tc_fence_after();
        const int tq = mt * 128 + r;
        uint32_t o[16];
        tmem_ld16(lane_addr + ATT_O_COL + 64, o);
        tmem_wait_ld();
        const float inv = 1.0f / __uint_as_float(o[0]);
        bf16* dst = p.ctx + ((long long)row * T + tq) * H + head * ATT_D;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          tmem_ld16(lane_addr + ATT_O_COL + c0, o);
          tmem_wait_ld();
          if (tq < T) {
            uint4 w0, w1;
            w0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
            w0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
            w0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
            w0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
            w1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
            w1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
            w1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
            w1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
            *reinterpret_cast<uint4*>(dst + c0) = w0;
            *reinterpret_cast<uint4*>(dst + c0 + 8) = w1;
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(bar_free));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ATT_TMEM_COLS);
}

int attention_pipe(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                   int mode, bf16* ctx, cudaStream_t stream);

int attention_tc(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                 int mode, bf16* ctx, cudaStream_t stream) {
  AGB_REQUIRE(share >= 1 && rows % share == 0, "share must divide the number of mask rows");
  AGB_REQUIRE(rows >= 0 && T > 0 && heads > 0 && H == heads * ATT_D, "tensor-core attention needs head dim 64");
  AGB_REQUIRE(words * 32 >= T, "mask words");
  AGB_REQUIRE(mode == AGB_MASK_MUL0 || mode == AGB_MASK_NEGINF, "mask mode");
  if (T > 512) {
    set_last_error("agb_masked_attention_bf16 supports T <= 512 (got %d); use agb_masked_attention_simt", T);
    return AGB_ERR_UNSUPPORTED;
  }
  if (rows == 0) return AGB_OK;
  AGB_REQUIRE(qkv && mask && ctx, "null pointer");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0, "alignment");
  {  // second-generation pipelined kernel (agb_attention_pipe.cu); this file keeps the first generation
    const int rc2 = attention_pipe(qkv, mask, words, rows, share, T, H, heads, mode, ctx, stream);
    if (rc2 != AGB_ERR_UNSUPPORTED) return rc2;
  }
  if (share != 1 || T > 256) {
    set_last_error("shared-qkv attention and T > 256 need the pipelined kernel (agb_attention_set_variant(0))");
    return AGB_ERR_UNSUPPORTED;
  }
  AttParams p;
  p.mask = mask; p.words = words; p.rows = rows; p.T = T; p.H = H; p.heads = heads; p.mode = mode;
  p.NK = (T + 15) / 16 * 16;
  p.units = rows * heads;
  p.mtiles = (T + 127) / 128;
  p.ctx = ctx;
  CUtensorMap tmQ, tmKV;
  int rc = encode_tmap_3d_bf16(&tmQ, qkv, 3 * (uint64_t)H, T, rows, (uint64_t)3 * H * 2, (uint64_t)T * 3 * H * 2,
                               ATT_D, 128, 1);
  if (rc != AGB_OK) return rc;
  rc = encode_tmap_3d_bf16(&tmKV, qkv, 3 * (uint64_t)H, T, rows, (uint64_t)3 * H * 2, (uint64_t)T * 3 * H * 2,
                           ATT_D, p.NK, 1);
  if (rc != AGB_OK) return rc;
  const int smem = 1024 + 128 * 128 + 3 * p.NK * 128 + 128;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  const int max_ctas = 2 * sm_count();
  const int grid = p.units < max_ctas ? p.units : max_ctas;
  attention_tc_kernel<<<grid, ATT_THREADS, smem, stream>>>(tmQ, tmKV, p);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

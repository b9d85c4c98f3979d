// Dense bf16 GEMM on the 5th-gen tensor cores:  C[M,N] = epi(A[M,K] * B[N,K]^T + bias) (+ residual)
//
// This is the workhorse behind every Linear of the frozen surrogate backbone and the explainer
// (reference: nn.Linear calls at models/vanilla_vit.py:437-441, 473-479, 487-493, 506-513 and the
// BERT twins models/vanilla_bert.py:503-537, 556-604).  Design (B200-first, not a port):
//   * persistent grid, one CTA per SM, static tile striding with N fastest so the three/nine/twelve
//     N-tiles that share one A row-panel run concurrently and hit L2;
//   * warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B boxes, STAGES-deep mbarrier ring);
//   * warp 1 = single-thread tcgen05.mma issuer, 128 x BN x 16 UMMAs, fp32 accumulators in TMEM,
//     two accumulator stages (2*BN columns) so the epilogue of tile i overlaps the MMAs of i+1;
//   * warps 2..9 = epilogue: tcgen05.ld -> bias / erf-GELU / residual -> bf16 or fp32 global store.
//   * operands may be K-major (row-major [rows,K], the nn.Linear layout) or MN-major (row-major
//     [K,rows]); the latter serves dgrad (dX = dY * W) and wgrad (dW = dY^T * X) without any
//     transposed copies.
#include "agb_common.cuh"

namespace agb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;

struct GemmParams {
  int M, N, K;
  const float* bias;        // [N] fp32 or nullptr
  const bf16* res_bf16;     // [*, ldr] bf16 residual or nullptr
  const float* res_f32;     // [*, ldr] fp32 residual or nullptr
  int ldr;
  // residual row remap: r = (m / res_group) * res_rows + (m % res_rows) when res_group > 0
  // (broadcast of a per-image tensor over its coalitions); identity when res_group == 0.
  int res_group, res_rows;
  void* out;
  int ldo;
  int out_f32;
  int act;                  // 0 none, 1 erf-GELU
  int a_mn, b_mn;           // operand majorness (0 = K-major, 1 = MN-major)
  float alpha;              // scale applied to the accumulator before bias
};

template <int BN, int STAGES, int ACT, int RES, int OUT_F32>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ATOM_BYTES = GEMM_BK * 128;  // one MN-major atom: 64 k-rows x 128 B
  constexpr int TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + STAGES;
  uint64_t* bar_tfull = bars + 2 * STAGES;
  uint64_t* bar_tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * GEMM_BM;
        const int n0 = (t % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_arrive_expect_tx(full, STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const int k0 = kb * GEMM_BK;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, full, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j)
              tma_load_2d(sa + j * ATOM_BYTES, &tmA, full, m0 + 64 * j, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmB, full, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * ATOM_BYTES, &tmB, full, n0 + 64 * j, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, p.a_mn, p.b_mn);
    const uint32_t a_lbo = p.a_mn ? ATOM_BYTES : 16, a_kstep = p.a_mn ? 2048 : 32;
    const uint32_t b_lbo = p.b_mn ? ATOM_BYTES : 16, b_kstep = p.b_mn ? 2048 : 32;
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (kb == num_kb - 1) umma_commit(smem_u32(&bar_tfull[acc]));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    // Two phases per 32-column chunk so that every global access is coalesced:
    //  (1) thread = TMEM lane = output row: tcgen05.ld 32 fp32 columns, park them in a per-warp
    //      4 KB staging tile (16-byte pieces XOR-swizzled by row -> conflict-free both ways);
    //  (2) 8 lanes per row x 4 rows per pass: read 4 columns back, apply alpha/bias/GELU/residual
    //      and store 128 B (fp32) or 64 B (bf16) contiguous per row.
    // ACT / RES / OUT_F32 are compile-time so the inner passes carry no branches or pointer tests.
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int h = (warp - 2) >> 2;          // column half
    constexpr int COLS_PER_WARP = BN / 2;
    constexpr int OUT_ES = OUT_F32 ? 4 : 2;                 // output element size
    constexpr int RES_ES = (RES == 2) ? 4 : 2;              // residual element size
    const uint32_t stage_u32 = smem_u32(smem + STAGES * STAGE_BYTES + 256 + (warp - 2) * 4096);
    const int piece = lane & 7, rsub = lane >> 3;
    const uint32_t st_base = stage_u32 + lane * 128;
    const uint32_t st_xor = static_cast<uint32_t>(lane & 7);
    // read-back address of pass ps: row_l = 4*ps + rsub, (row_l & 7) = rsub + 4*(ps & 1)
    const uint32_t ld_even = stage_u32 + rsub * 128 + ((piece ^ rsub) << 4);
    const uint32_t ld_odd = stage_u32 + (rsub + 4) * 128 + ((piece ^ (rsub + 4)) << 4);
    const long long out_pitch4 = 4ll * p.ldo * OUT_ES;      // bytes between passes
    const long long res_pitch4 = 4ll * p.ldr * RES_ES;
    uint32_t acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * GEMM_BM;
      const int n0 = (t % tiles_n) * BN;
      mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
      tc_fence_after();
      const int row0 = m0 + q * 32 + rsub;                  // row of pass 0 for this lane
      const int rows_left = p.M - row0;                     // pass ps valid iff 4*ps < rows_left
      const int ncol0 = n0 + h * COLS_PER_WARP + piece * 4;
      uint8_t* out_row = reinterpret_cast<uint8_t*>(p.out) + ((long long)row0 * p.ldo + ncol0) * OUT_ES;
      const uint8_t* res_row = nullptr;
      if (RES == 1) res_row = reinterpret_cast<const uint8_t*>(p.res_bf16) + ((long long)row0 * p.ldr + ncol0) * 2;
      if (RES == 2) res_row = reinterpret_cast<const uint8_t*>(p.res_f32) + ((long long)row0 * p.ldr + ncol0) * 4;
      const uint32_t tm_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + h * COLS_PER_WARP;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tm_addr + c * 32, r);
        const int n = ncol0 + c * 32;
        const bool n_ok = n < p.N;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr && n_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        // residual prefetch for the 8 passes (coalesced, issued before the TMEM wait)
        uint2 rb[8];
        float4 rf[8];
        if (RES != 0) {
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const bool ok = n_ok && (4 * ps < rows_left);
            const uint8_t* rp = res_row + ps * res_pitch4 + (long long)c * 32 * RES_ES;
            if (RES == 1) rb[ps] = ok ? __ldg(reinterpret_cast<const uint2*>(rp)) : make_uint2(0u, 0u);
            if (RES == 2) rf[ps] = ok ? __ldg(reinterpret_cast<const float4*>(rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_base + ((j ^ st_xor) << 4)),
                       "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
          float4 v;
          const uint32_t addr = ((ps & 1) ? ld_odd : ld_even) + (ps >> 1) * 1024;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(addr)
                       : "memory");
          v.x = fmaf(v.x, p.alpha, bias4.x);
          v.y = fmaf(v.y, p.alpha, bias4.y);
          v.z = fmaf(v.z, p.alpha, bias4.z);
          v.w = fmaf(v.w, p.alpha, bias4.w);
          if (ACT == 1) {
            v.x = gelu_erf_fast(v.x); v.y = gelu_erf_fast(v.y);
            v.z = gelu_erf_fast(v.z); v.w = gelu_erf_fast(v.w);
          }
          if (RES == 1) {
            v.x += bf16_lo(rb[ps].x); v.y += bf16_hi(rb[ps].x);
            v.z += bf16_lo(rb[ps].y); v.w += bf16_hi(rb[ps].y);
          }
          if (RES == 2) { v.x += rf[ps].x; v.y += rf[ps].y; v.z += rf[ps].z; v.w += rf[ps].w; }
          if (n_ok && 4 * ps < rows_left) {
            uint8_t* op = out_row + ps * out_pitch4 + (long long)c * 32 * OUT_ES;
            if (OUT_F32) {
              *reinterpret_cast<float4*>(op) = v;
            } else {
              uint2 pk;
              pk.x = pack_bf16x2(v.x, v.y);
              pk.y = pack_bf16x2(v.z, v.w);
              *reinterpret_cast<uint2*>(op) = pk;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, int ACT, int RES, int OUT_F32>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                       cudaStream_t stream) {
  constexpr int SMEM = STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 1024 + 256 + GEMM_EPI_WARPS * 4096;
  static bool configured = false;
  if (!configured) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, ACT, RES, OUT_F32>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tc_kernel<BN, STAGES, ACT, RES, OUT_F32><<<grid, GEMM_THREADS, SMEM, stream>>>(tmA, tmB, p);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

template <int BN, int STAGES>
static int dispatch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  const int res = p.res_bf16 ? 1 : (p.res_f32 ? 2 : 0);
#define AGB_CASE(A, R, O) \
  if (p.act == A && res == R && p.out_f32 == O) return launch_gemm<BN, STAGES, A, R, O>(tmA, tmB, p, st);
  AGB_CASE(0, 0, 0) AGB_CASE(0, 0, 1) AGB_CASE(0, 1, 0) AGB_CASE(0, 1, 1) AGB_CASE(0, 2, 0) AGB_CASE(0, 2, 1)
  AGB_CASE(1, 0, 0) AGB_CASE(1, 0, 1) AGB_CASE(1, 1, 0) AGB_CASE(1, 1, 1) AGB_CASE(1, 2, 0) AGB_CASE(1, 2, 1)
#undef AGB_CASE
  set_last_error("unsupported GEMM epilogue (act=%d)", p.act);
  return AGB_ERR_INVALID;
}

int gemm_bf16_pair(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N, int K,
                   float alpha, const float* bias, int act, const bf16* res_bf16, const float* res_f32, int ldr,
                   void* out, int ldo, int out_f32, cudaStream_t stream);

// Host entry used by the C-ABI (agb_api.cu).  lda/ldb are row pitches in elements of the stored
// matrices: K-major operand = [rows, K] (pitch >= K); MN-major operand = [K, rows] (pitch >= rows).
int gemm_bf16_tc(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N,
                 int K, float alpha, const float* bias, int act, const bf16* res_bf16,
                 const float* res_f32, int ldr, int res_group, int res_rows, void* out, int ldo,
                 int out_f32, cudaStream_t stream) {
  AGB_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  AGB_REQUIRE(A && B && out, "null operand");
  AGB_REQUIRE((N % 4) == 0, "N must be a multiple of 4");
  AGB_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0, "operand pitch must be a multiple of 8 elements");
  AGB_REQUIRE((ldo % (out_f32 ? 4 : 8)) == 0, "output pitch alignment");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "operands must be 16-byte aligned");
  AGB_REQUIRE(!(res_bf16 || res_f32) || (ldr % 8) == 0, "residual pitch alignment");
  AGB_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "bias alignment");
  AGB_REQUIRE(res_group == 0 && res_rows == 0, "residual row remap is reserved (pass 0, 0)");
  AGB_REQUIRE(act == 0 || act == 1, "activation");
  AGB_REQUIRE(!(res_bf16 && res_f32), "one residual at most");

  // second-generation kernel (CTA pairs + TMA epilogue) covers the hot shapes; this file keeps the rest
  {
    const int rc2 = gemm_bf16_pair(A, lda, a_mn, B, ldb, b_mn, M, N, K, alpha, bias, act, res_bf16, res_f32, ldr,
                                   out, ldo, out_f32, stream);
    if (rc2 != AGB_ERR_UNSUPPORTED) return rc2;
  }
  const bool wide = N >= 192;
  const int BN = wide ? 256 : 128;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = encode_tmap_2d_bf16(&tmA, A, K, M, (uint64_t)lda * 2, GEMM_BK, GEMM_BM);
  else       rc = encode_tmap_2d_bf16(&tmA, A, M, K, (uint64_t)lda * 2, 64, GEMM_BK);
  if (rc != AGB_OK) return rc;
  if (!b_mn) rc = encode_tmap_2d_bf16(&tmB, B, K, N, (uint64_t)ldb * 2, GEMM_BK, BN);
  else       rc = encode_tmap_2d_bf16(&tmB, B, N, K, (uint64_t)ldb * 2, 64, GEMM_BK);
  if (rc != AGB_OK) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.res_bf16 = res_bf16; p.res_f32 = res_f32; p.ldr = ldr;
  p.res_group = res_group; p.res_rows = res_rows;
  p.out = out; p.ldo = ldo; p.out_f32 = out_f32; p.act = act;
  p.a_mn = a_mn; p.b_mn = b_mn; p.alpha = alpha;
  if (wide) return dispatch_gemm<256, 4>(tmA, tmB, p, stream);
  return dispatch_gemm<128, 6>(tmA, tmB, p, stream);
}

}  // namespace agb

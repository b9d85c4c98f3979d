// KernelSHAP weighted least squares, batched over explained samples (a14).
//
// In the reference the solve happens inside the third-party `shap.KernelExplainer` on the CPU in
// float64 numpy/LAPACK (call site reference models/kernel_shap_bert.py:170-185).  Here, per explained
// sample b: coalitions Z_b (S x d, packed bits), kernel weights w_b (S), background-averaged model outputs
// p_b (S x C), the sample's own output f_b (C) and the null output f0 (C) go in; attributions phi (C x d)
// come out, all on the device and in float64:
//     y = link(p) - link(f0),  delta = link(f) - link(f0)
//     E = Z[:, :-1] - Z[:, -1:],  y~ = y - Z[:, -1:] delta           (efficiency constraint eliminated)
//     A = E^T W E  ((d-1) x (d-1) Gram),  R = E^T W y~                (kernel 1: register-tiled fp64 accumulation)
//     A = L L^T (Cholesky), phi[:-1] = A^-1 R, phi[-1] = delta - sum  (kernel 2: blocked, one CTA per sample)
// shap's optional l1_reg feature pre-selection is NOT part of this path (documented in oracle/kernelshap.py).
// Bytes per sample (S=2048, d=127, C=2): packed Z 32 KB + p 16 KB in, A 127 KB out/in, phi 2 KB out.
#include <stdlib.h>

#include "agb_common.cuh"

namespace agb {

constexpr int KG_T = 64;         // Gram tile edge
constexpr int KG_CH = 32;        // coalitions staged per shared-memory chunk (= bits of one column word)
constexpr int KG_THREADS = 64;   // 8 x 8 threads, 8 x 8 accumulators each

__device__ __forceinline__ double ks_link(double p, int link) {
  return link ? log(p / (1.0 - p)) : p;
}

// high word of the double 1.0 when (bits & mask) != 0, else 0 (the double 0.0): bit test into a predicate + select
__device__ __forceinline__ int ks_one_if(uint32_t bits, uint32_t mask) {
  int hi;
  asm("{\n"
      ".reg .pred p;\n"
      ".reg .b32 tt;\n"
      "and.b32 tt, %1, %2;\n"
      "setp.ne.u32 p, tt, 0;\n"
      "selp.b32 %0, 0x3FF00000, 0, p;\n"
      "}\n"
      : "=r"(hi)
      : "r"(bits), "r"(mask));
  return hi;
}

// bits of columns [c0, c0 + 32) that belong to the (d-1) x (d-1) system
__device__ __forceinline__ uint32_t ks_valid_bits(int c0, int n) {
  if (c0 + 32 <= n) return 0xFFFFFFFFu;
  if (c0 >= n) return 0u;
  return (1u << (n - c0)) - 1u;
}

// grid: (lower-triangle 64x64 tiles of A + one 64 x C strip of R per tile row, B).  fp64-pipe bound.
// With z' = z XOR z_last (per coalition) the eliminated design matrix is E[s,j] = sigma_s z'[s,j], sigma_s = +-1, so
//   A[j,k] = sum_s w_s z'[s,j] z'[s,k]          R[j,c] = sum_s (w_s z'[s,j]) (sigma_s y~[s,c])
// i.e. A accumulates w_s wherever both bits are set: no multiplications and no cancellation.  Per chunk of 32
// coalitions the j side is staged as doubles (w_s or 0) and the k side as one 32-bit word per column (bit ss = z');
// each thread keeps an 8 x 8 block in registers and, per coalition and column, turns the column's bit into the double
// 1.0 / 0.0 with two integer instructions that feed eight DFMAs, so shared memory delivers only the j side: 64 B per
// thread per coalition for 64 DFMAs (an 8 x 4 block with both sides in shared memory was LSU-bound at 45 % of the
// fp64 pipe, profiles/r01_kernelshap_ncu.txt).
template <int CMAX>   // classes carried by the rhs strip (4 or 16)
__global__ void __launch_bounds__(KG_THREADS, 6)
kernelshap_gram_kernel(const uint32_t* __restrict__ Z, int words, const double* __restrict__ w,
                       const double* __restrict__ probs, const double* __restrict__ fx,
                       const double* __restrict__ f0, int S, int d, int C, int link, double* __restrict__ A,
                       double* __restrict__ R, int rhs_only) {
  const int n = d - 1;
  const int tiles = (n + KG_T - 1) / KG_T;
  const int n_lower = tiles * (tiles + 1) / 2;
  const int b = blockIdx.y;
  const int tile = blockIdx.x + (rhs_only ? n_lower : 0);     // rhs_only: the Gram tiles come from the tensor-core kernel
  const bool is_rhs = tile >= n_lower;
  int tj = 0, tk = 0;                      // tile row, tile column (tk <= tj: lower triangle, symmetric matrix)
  if (is_rhs) {
    tj = tk = tile - n_lower;
  } else {
    while ((tj + 1) * (tj + 2) / 2 <= tile) ++tj;
    tk = tile - tj * (tj + 1) / 2;
  }
  __shared__ __align__(16) double wa[KG_CH][KG_T];    // w_s * z'[s, j-tile]
  __shared__ uint32_t cbits[KG_T];                    // column k of the k-tile: bit ss = z'[s0 + ss, k]
  __shared__ double us[KG_CH][CMAX];                  // sigma_s * y~[s, c] for the rhs strip (C <= CMAX)
  __shared__ double l0s[CMAX], dls[CMAX];             // link(f0[c]) and link(fx[b, c]) - link(f0[c])
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int ty = t >> 3, tx = t & 7;
  const uint32_t* Zb = Z + (long long)b * S * words;
  const int last = d - 1;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = 0.0;
  if (is_rhs) {
    for (int e = t; e < KG_CH * CMAX; e += KG_THREADS) us[e / CMAX][e % CMAX] = 0.0;
    if (t < C) {
      l0s[t] = ks_link(f0[t], link);
      dls[t] = ks_link(fx[(long long)b * C + t], link) - l0s[t];
    }
  }
  const int cj = tj * KG_T + warp * 32, ck = tk * KG_T + warp * 32;   // this warp's 32 columns of either tile
  const uint32_t vj = ks_valid_bits(cj, n), vk = ks_valid_bits(ck, n);
  // lane = coalition of the chunk: its two column words, XORed with the eliminated feature's bit, and its weight;
  // fetched one chunk ahead so that the global-memory latency hides behind the DFMAs of the current chunk
  uint32_t nxt_j = 0, nxt_k = 0;
  double nxt_w = 0.0;
  auto fetch = [&](int s0) {
    const int s = s0 + lane;
    nxt_j = nxt_k = 0;
    nxt_w = 0.0;
    if (s < S) {
      const uint32_t* zr = Zb + (long long)s * words;
      const uint32_t flip = ((zr[last >> 5] >> (last & 31)) & 1u) ? 0xFFFFFFFFu : 0u;
      nxt_w = w[(long long)b * S + s];
      if (vj) nxt_j = (zr[cj >> 5] ^ flip) & vj;
      if (vk) nxt_k = (zr[ck >> 5] ^ flip) & vk;
    }
  };
  fetch(0);
  for (int s0 = 0; s0 < S; s0 += KG_CH) {
    __syncthreads();
    {
      const uint32_t wordj = nxt_j, wordk = nxt_k;
      const double wv = nxt_w;
      fetch(s0 + KG_CH);
#pragma unroll 8
      for (int ss = 0; ss < KG_CH; ++ss) {
        const uint32_t wj = __shfl_sync(0xffffffffu, wordj, ss);
        const double ws = __shfl_sync(0xffffffffu, wv, ss);
        wa[ss][warp * 32 + lane] = ((wj >> lane) & 1u) ? ws : 0.0;
      }
      if (!is_rhs) {
        uint32_t mine = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const uint32_t bal = __ballot_sync(0xffffffffu, (wordk >> c) & 1u);
          if (lane == c) mine = bal;
        }
        cbits[warp * 32 + lane] = mine;
      } else {
        for (int e = t; e < KG_CH * C; e += KG_THREADS) {
          const int ss = e / C, c = e % C;
          const int s2 = s0 + ss;
          double v = 0.0;
          if (s2 < S) {
            const uint32_t* zr = Zb + (long long)s2 * words;
            const double zl = (double)((zr[last >> 5] >> (last & 31)) & 1);
            const double yv = ks_link(probs[((long long)b * S + s2) * C + c], link) - l0s[c];
            v = (1.0 - 2.0 * zl) * (yv - zl * dls[c]);
          }
          us[ss][c] = v;
        }
      }
    }
    __syncthreads();
    if (!is_rhs) {
      uint32_t cb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) cb[q] = cbits[tx + 8 * q];
#pragma unroll
      for (int ss = 0; ss < KG_CH; ++ss) {
        const double2* pa = reinterpret_cast<const double2*>(&wa[ss][ty * 8]);
        const double2 a01 = pa[0], a23 = pa[1], a45 = pa[2], a67 = pa[3];
        const double av[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
        // each column's bit as the double 1.0 / 0.0 (two integer instructions, all eight before the DFMAs so that
        // the integer chains overlap the fp64 pipe), then 8 x 8 DFMAs
        double bv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) bv[q] = __hiloint2double(ks_one_if(cb[q], 1u << ss), 0);
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i][q] = fma(av[i], bv[q], acc[i][q]);
      }
    } else {
      // rhs strip: thread = row j of the tile, CMAX class accumulators in acc[0..1][0..7]
#pragma unroll 4
      for (int ss = 0; ss < KG_CH; ++ss) {
        const double a = wa[ss][t];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) acc[c >> 3][c & 7] = fma(a, us[ss][c], acc[c >> 3][c & 7]);
      }
    }
  }
  if (!is_rhs) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = tj * KG_T + ty * 8 + i;
      if (j >= n) continue;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = tk * KG_T + tx + 8 * q;
        if (k < n) A[((long long)b * n + j) * n + k] = acc[i][q];
      }
    }
  } else {
    const int j = tj * KG_T + t;
    if (j < n) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) R[((long long)b * n + j) * C + c] = acc[c >> 3][c & 7];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Gram matrix on the tensor cores, exactly (round 2).  A[j,k] = sum_s w_s z'[s,j] z'[s,k] with z' in {0,1}: the only
// non-integer factor is the weight.  Per sample the weights are put on a common fixed-point grid, q_s = round(|w_s| *
// 2^(55 - e_max)) < 2^56, and cut into seven 8-bit limbs; then
//     A = 2^(e_max - 55) * sum_l 2^(8 l) * ( (Z' o limb_l)^T Z' )
// and every bracket is a product of small INTEGERS: limb values 0..255 and bits 0/1 are exact in bf16, their products are
// exact, and a sum of S <= 65536 of them stays below 2^24, so tcgen05.mma kind::f16 with its fp32 accumulator computes
// each bracket without any rounding.  The seven brackets are recombined in float64 (one rounding per addition, fixed
// order: bit-reproducible; the result is the correctly scaled sum of the 56-bit fixed-point weights, i.e. at least as
// accurate as the float64 FMA chain it replaces, whose own rounding error is S * 2^-53).
// One CTA = one 128 (j) x 64 (k) tile of one sample's lower triangle; operands are GENERATED in shared memory, MN-major
// SWIZZLE_128B tiles of 32 coalitions: for a coalition row s and 8 consecutive features the 8 mask bits index a 256-entry
// table of 16-byte lane masks, which is ANDed with the broadcast value — 3 instructions per 8 operand elements.
// The j side carries the plain bits (one A tile, shared by all limbs), the k side the bits times the limb value: the seven
// limb tiles sit back to back in shared memory, so they are ONE B operand with N = 7 x 64 = 448 accumulator columns
// (issued as N = 256 + N = 192): four tcgen05.mma per stage instead of fourteen, and A is read twice instead of 7 times.
//   warp 0        tcgen05 issuer: per stage 2 K-steps x (N = 256, N = 192), D_l = TMEM columns [64 l, 64 l + 64)
//   warps 1-8     operand generators (3-stage ring), then the epilogue: fp32 integers -> float64 recombination -> A
// ------------------------------------------------------------------------------------------------------------------
// explicit shared-space accesses (generic ld / st on a shared pointer go through the slower generic path of the LSU:
// 11 % of the kernel's stall samples were `stall_lg` / MIO on ST.E.128 / LD.E.128, profiles/r02_kernelshap_gram_ncu.txt)
__device__ __forceinline__ void kt_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 kt_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// |w| * 2^(56 - e_max) rounded to nearest, as a 56-bit integer, with INTEGER instructions only (the float64 pipe is slow:
// the DMUL + F2I pair was 10 % of the stall samples).  w = mant * 2^(exp - 1075) with the implicit bit set.
__device__ __forceinline__ unsigned long long kt_to_fixed(double w, int e_max) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(w) & 0x7FFFFFFFFFFFFFFFull;
  const int ex = (int)(bits >> 52);
  if (ex == 0) return 0ull;                                  // zero / subnormal: below the grid
  const unsigned long long mant = (bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
  const int sh = ex - 1075 + 56 - e_max;                     // q = mant * 2^sh; sh <= 3 because |w| < 2^e_max
  unsigned long long q;
  if (sh >= 0) q = mant << sh;
  else if (sh < -54) q = 0ull;
  else q = (mant + (1ull << (-sh - 1))) >> (-sh);
  return q >> 56 ? (1ull << 56) - 1 : q;
}

constexpr int KT_BM = 128, KT_BN = 64, KT_BK = 32, KT_LIMBS = 7, KT_STAGES = 4;
constexpr int KT_A_BYTES = KT_BK * KT_BM * 2;                 // A tile = the j-features' bits as bf16 0 / 1: 32 rows x 256 B (two mn atoms)
constexpr int KT_B_BYTES = KT_BK * KT_BN * 2;                 // one limb's B tile: the k-features' bits x limb value, 32 rows x 128 B
constexpr int KT_STAGE_BYTES = KT_A_BYTES + KT_LIMBS * KT_B_BYTES;     // 36 864
constexpr int KT_THREADS = 288;                               // warp 0 issuer + 8 generator warps
constexpr int KT_SMEM = 1024 + KT_STAGES * KT_STAGE_BYTES + 4096 /*lut*/ + KT_BK * KT_STAGES * 8 * 4 /*limbs*/ + 512;

__global__ void __launch_bounds__(KT_THREADS, 1)
kernelshap_gram_tc_kernel(const uint32_t* __restrict__ Z, int words, const double* __restrict__ w, int S, int d,
                          double* __restrict__ A) {
  extern __shared__ uint8_t kt_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(kt_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stages = smem;
  uint4* lut = reinterpret_cast<uint4*>(smem + KT_STAGES * KT_STAGE_BYTES);           // [256] 8 x u16 lane masks
  uint32_t* limbs = reinterpret_cast<uint32_t*>(lut + 256);                            // [stage][32 coalitions][8] bf16 pairs
  uint64_t* bars = reinterpret_cast<uint64_t*>(limbs + KT_STAGES * KT_BK * 8);
  uint64_t* full = bars;              // [4]
  uint64_t* empty = bars + 4;         // [4]
  uint64_t* done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  double* red = reinterpret_cast<double*>(bars + 10);                                  // [9] per-warp maxima

  const int n = d - 1;
  const int b = blockIdx.y;
  // tile index -> (tj, tk): j-tile tj owns the k-tiles 0 .. 2 tj + 1 (lower triangle incl. the diagonal block)
  int tj = 0, rest = blockIdx.x;
  while (rest >= 2 * tj + 2) { rest -= 2 * tj + 2; ++tj; }
  const int tk = rest;
  const int j0 = tj * KT_BM, k0 = tk * KT_BN;
  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t* Zb = Z + (long long)b * S * words;
  const double* wb = w + (long long)b * S;
  const int last = d - 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < KT_STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 8);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(done), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  // lane-mask table and the sample's largest |weight| (common fixed-point exponent)
  for (int e = threadIdx.x; e < 256; e += KT_THREADS) {
    uint4 m;
    m.x = ((e & 1) ? 0x0000FFFFu : 0u) | ((e & 2) ? 0xFFFF0000u : 0u);
    m.y = ((e & 4) ? 0x0000FFFFu : 0u) | ((e & 8) ? 0xFFFF0000u : 0u);
    m.z = ((e & 16) ? 0x0000FFFFu : 0u) | ((e & 32) ? 0xFFFF0000u : 0u);
    m.w = ((e & 64) ? 0x0000FFFFu : 0u) | ((e & 128) ? 0xFFFF0000u : 0u);
    lut[e] = m;
  }
  double wmax = 0.0;
  for (int s = threadIdx.x; s < S; s += KT_THREADS) wmax = fmax(wmax, fabs(wb[s]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) red[warp] = wmax;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  wmax = 0.0;
#pragma unroll
  for (int i = 0; i < KT_THREADS / 32; ++i) wmax = fmax(wmax, red[i]);
  int e_max = 0;
  if (wmax > 0.0) frexp(wmax, &e_max);               // wmax = m * 2^e_max, m in [0.5, 1)
  const int n_stages = (S + KT_BK - 1) / KT_BK;

  if (warp == 0) {
    // ------------------------------ tcgen05 issuer ------------------------------
    const uint32_t idesc_lo = make_idesc_bf16(KT_BM, 4 * KT_BN, 1, 1);     // limbs 0-3: N = 256
    const uint32_t idesc_hi = make_idesc_bf16(KT_BM, 3 * KT_BN, 1, 1);     // limbs 4-6: N = 192
    for (int st = 0; st < n_stages; ++st) {
      const int sl = st % KT_STAGES;
      mbar_wait(smem_u32(&full[sl]), (st / KT_STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(stages + sl * KT_STAGE_BYTES);
        const uint32_t sb = sa + KT_A_BYTES;
#pragma unroll
        for (int kk = 0; kk < KT_BK / 16; ++kk) {
          const uint64_t da = make_smem_desc_sw128(sa + kk * 2048, KT_BK * 128, 1024);
          const uint64_t db0 = make_smem_desc_sw128(sb + kk * 2048, KT_BK * 128, 1024);
          const uint64_t db1 = make_smem_desc_sw128(sb + 4 * KT_B_BYTES + kk * 2048, KT_BK * 128, 1024);
          umma_ss(tmem_base, da, db0, idesc_lo, (st | kk) != 0 ? 1u : 0u);
          umma_ss(tmem_base + 4 * KT_BN, da, db1, idesc_hi, (st | kk) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty[sl]));
        if (st == n_stages - 1) umma_commit(smem_u32(done));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ operand generators ------------------------------
    const int gt = threadIdx.x - 32;                   // 0 .. 255
    const uint32_t stages_s = smem_u32(stages), lut_s = smem_u32(lut);
    // my tasks of a stage: two A tasks (coalition row r, 16-byte chunk c of the 128 j-features) and one B task
    const int ra0 = gt >> 4, ca0 = gt & 15;            // task gt
    const int ra1 = (gt + 256) >> 4;                   // task gt + 256 (same chunk column, row + 16)
    const int rb = gt >> 3, cb = gt & 7;
    const int ja = j0 + ca0 * 8, kf = k0 + cb * 8;
    const uint32_t va = ja < n ? ks_valid_bits(ja & ~31, n) : 0u, vb = kf < n ? ks_valid_bits(kf & ~31, n) : 0u;
    // the mask bytes (and, for the first 32 threads, the weight) of the NEXT stage are fetched while the current one is
    // generated: the two dependent-free global loads per task are the long-latency part of a stage
    // RAW words are kept in registers and only turned into bytes one iteration later, so the loads stay in flight
    uint32_t nf_a0 = 0, nd_a0 = 0, nf_a1 = 0, nd_a1 = 0, nf_b = 0, nd_b = 0;
    double nx_w = 0.0;
    const int wl = last >> 5, wa_ = ja >> 5, wb_ = kf >> 5;
    auto fetch = [&](int st) {
      const int s0 = st * KT_BK;
      const int sa0 = s0 + ra0, sa1 = s0 + ra1, sbr = s0 + rb;
      nf_a0 = nd_a0 = nf_a1 = nd_a1 = nf_b = nd_b = 0u;
      if (sa0 < S && va) { nf_a0 = __ldg(Zb + (long long)sa0 * words + wl); nd_a0 = __ldg(Zb + (long long)sa0 * words + wa_); }
      if (sa1 < S && va) { nf_a1 = __ldg(Zb + (long long)sa1 * words + wl); nd_a1 = __ldg(Zb + (long long)sa1 * words + wa_); }
      if (sbr < S && vb) { nf_b = __ldg(Zb + (long long)sbr * words + wl); nd_b = __ldg(Zb + (long long)sbr * words + wb_); }
      nx_w = (sbr < S) ? wb[sbr] : 0.0;                   // weight of my B task's coalition
    };
    auto to_byte = [&](uint32_t fw, uint32_t dw, uint32_t valid, int f) -> uint32_t {
      const uint32_t flip = ((fw >> (last & 31)) & 1u) ? 0xFFFFFFFFu : 0u;
      return (((dw ^ flip) & valid) >> (f & 31)) & 0xFFu;
    };
    fetch(0);
    for (int st = 0; st < n_stages; ++st) {
      const int sl = st % KT_STAGES;
      // (rows past S were fetched as zero words with a zero flip word: byte 0)
      const uint32_t b_a0 = to_byte(nf_a0, nd_a0, va, ja), b_a1 = to_byte(nf_a1, nd_a1, va, ja), b_b = to_byte(nf_b, nd_b, vb, kf);
      const double wv = nx_w;
      if (st + 1 < n_stages) fetch(st + 1);
      if (st >= KT_STAGES) mbar_wait(smem_u32(&empty[sl]), ((st / KT_STAGES) - 1) & 1);
      // (1) limb values of my B task's coalition: bf16 pair (v, v) of sign * ((q >> 8 l) & 255).  Every thread converts its
      // own weight (8 threads share a coalition: redundant, but no exchange and no barrier between the generator warps,
      // so their per-stage latency chains overlap instead of adding up)
      uint32_t lv[KT_LIMBS];
      {
        const unsigned long long q = kt_to_fixed(wv, e_max);
        const bool neg = __double2hiint(wv) < 0;
#pragma unroll
        for (int l = 0; l < KT_LIMBS; ++l) {
          const float v = (float)(uint32_t)((q >> (8 * l)) & 255ull);
          const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(neg ? -v : v));
          lv[l] = h | (h << 16);
        }
      }
      const uint32_t sa = stages_s + sl * KT_STAGE_BYTES;
      const uint32_t sb = sa + KT_A_BYTES;
      // (2) A tile (j side): bf16 1.0 where the bit is set
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const int r = rep ? ra1 : ra0;
        const uint4 m = kt_lds128(lut_s + (rep ? b_a1 : b_a0) * 16);
        const uint32_t one = 0x3F803F80u;
        const uint32_t off = (uint32_t)(ca0 >> 3) * (KT_BK * 128) + (uint32_t)r * 128 + ((((uint32_t)ca0 & 7u) ^ ((uint32_t)r & 7u)) << 4);
        kt_sts128(sa + off, m.x & one, m.y & one, m.z & one, m.w & one);
      }
      // (3) B tiles (k side), one per limb, back to back: bit x limb value
      {
        const uint4 m = kt_lds128(lut_s + b_b * 16);
        const uint32_t off = (uint32_t)rb * 128 + ((((uint32_t)cb) ^ ((uint32_t)rb & 7u)) << 4);
#pragma unroll
        for (int l = 0; l < KT_LIMBS; ++l)
          kt_sts128(sb + l * KT_B_BYTES + off, m.x & lv[l], m.y & lv[l], m.z & lv[l], m.w & lv[l]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&full[sl]));
    }
    // ------------------------------ epilogue: recombine the limbs in float64 ------------------------------
    mbar_wait(smem_u32(done), 0);
    tc_fence_after();
    const int gw = warp - 1;                     // 0 .. 7
    const int qd = warp & 3;                     // TMEM lane quarter this warp may touch
    const int chalf = (gw >> 2) & 1;             // which 32 of the 64 k-columns (two warps share a lane quarter)
    // warps 1..8: (warp & 3) = 1,2,3,0,1,2,3,0 -> every quarter is covered twice, once per column half
    const int j = j0 + qd * 32 + lane;
    double res[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) res[i] = 0.0;
    const double inv_fixed = ldexp(1.0, e_max - 56);
#pragma unroll
    for (int l = KT_LIMBS - 1; l >= 0; --l) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + l * KT_BN + chalf * 32, v);
      tmem_wait_ld();
      const double sc = (double)(1ull << (8 * l));
#pragma unroll
      for (int i = 0; i < 32; ++i) res[i] = fma((double)__uint_as_float(v[i]), sc, res[i]);
    }
    if (j < n) {
      double* arow = A + ((long long)b * n + j) * n;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = k0 + chalf * 32 + i;
        if (k < n) arow[k] = res[i] * inv_fixed;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// Right-hand sides R[j,c] = sum_s z'[s,j] * (w_s sigma_s y~[s,c]) when the Gram matrix comes from the tensor-core kernel:
// one CTA per sample, thread = feature j.  Per chunk of 128 coalitions the per-coalition values v[s][c] = w_s sigma_s y~[s,c]
// and the flipped mask words are staged once; every thread then walks the chunk, testing its own bit (the word is a warp
// broadcast) and adding v — a masked float64 sum, 2 S C additions per feature.
constexpr int KR_CH = 128;
// grid (B, G): slice g of G handles the coalitions [g * S / G, (g + 1) * S / G) and adds into R (zeroed by the caller when
// G > 1); inside the CTA, `parts` threads share a feature (interleaved coalitions, combined through shared memory)
template <int CMAX>
__global__ void __launch_bounds__(256)
kernelshap_rhs_kernel(const uint32_t* __restrict__ Z, int words, const double* __restrict__ w, const double* __restrict__ probs,
                      const double* __restrict__ fx, const double* __restrict__ f0, int S, int d, int C, int link,
                      double* __restrict__ R) {
  __shared__ double v[KR_CH][CMAX];
  __shared__ uint32_t zs[KR_CH][33];          // up to 32 mask words per coalition (d <= 1024), padded against bank conflicts
  __shared__ double l0s[CMAX], dls[CMAX];
  __shared__ double comb[256][CMAX > 4 ? 1 : CMAX];      // partial sums of the parts (only used when parts > 1, C <= 4)
  const int n = d - 1, b = blockIdx.x, t = threadIdx.x, last = d - 1;
  const int G = gridDim.y, g = blockIdx.y;
  const int s_lo = (int)((long long)S * g / G), s_hi = (int)((long long)S * (g + 1) / G);
  const uint32_t* Zb = Z + (long long)b * S * words;
  if (t < C) {
    l0s[t] = ks_link(f0[t], link);
    dls[t] = ks_link(fx[(long long)b * C + t], link) - l0s[t];
  }
  __syncthreads();
  const int npad = (n + 31) & ~31;
  const int parts = (CMAX <= 4 && npad <= 128) ? 256 / npad : 1;      // threads per feature
  const int span = parts > 1 ? npad : 256;
  for (int jb = 0; jb < n; jb += span) {
    const int j = jb + (parts > 1 ? t % npad : t);
    const int part = parts > 1 ? t / npad : 0;
    double acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.0;
    for (int s0 = s_lo; s0 < s_hi; s0 += KR_CH) {
      __syncthreads();
      for (int e = t; e < KR_CH * words; e += 256) {
        const int ss = e / words, wi = e - ss * words;
        uint32_t val = 0;
        if (s0 + ss < s_hi) {
          const uint32_t* zr = Zb + (long long)(s0 + ss) * words;
          const uint32_t flip = ((__ldg(zr + (last >> 5)) >> (last & 31)) & 1u) ? 0xFFFFFFFFu : 0u;
          val = __ldg(zr + wi) ^ flip;
        }
        zs[ss][wi] = val;
      }
      for (int e = t; e < KR_CH * C; e += 256) {
        const int ss = e / C, c = e - ss * C;
        double val = 0.0;
        if (s0 + ss < s_hi) {
          const uint32_t* zr = Zb + (long long)(s0 + ss) * words;
          const double zl = (double)((__ldg(zr + (last >> 5)) >> (last & 31)) & 1u);
          const double yv = ks_link(probs[((long long)b * S + s0 + ss) * C + c], link) - l0s[c];
          val = w[(long long)b * S + s0 + ss] * (1.0 - 2.0 * zl) * (yv - zl * dls[c]);
        }
        v[ss][c] = val;
      }
      __syncthreads();
      if (j < n && part < parts) {
        const int wi = j >> 5, sh = j & 31;
#pragma unroll 4
        for (int ss = part; ss < KR_CH; ss += parts) {
          const bool on = (zs[ss][wi] >> sh) & 1u;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) acc[c] += on ? v[ss][c] : 0.0;
        }
      }
    }
    if (parts > 1) {
      __syncthreads();
      if (CMAX <= 4) {
#pragma unroll
        for (int c = 0; c < (CMAX > 4 ? 1 : CMAX); ++c) comb[t][c] = acc[c];
      }
      __syncthreads();
      if (part == 0 && CMAX <= 4) {
        for (int q = 1; q < parts; ++q)
#pragma unroll
          for (int c = 0; c < (CMAX > 4 ? 1 : CMAX); ++c) acc[c] += comb[q * npad + (t % npad)][c];
      }
    }
    if (j < n && part == 0) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          double* dst = R + ((long long)b * n + j) * C + c;
          if (G > 1) atomicAdd(dst, acc[c]);
          else *dst = acc[c];
        }
    }
  }
}

// One CTA per sample: blocked right-looking Cholesky of the lower triangle (32-column panels), with the C right-hand
// sides carried as C extra rows below the matrix so that the forward substitution L z = R falls out of the panel
// solves and trailing updates; then a blocked backward substitution L^T x = z and the eliminated feature.
//   per panel: (1) 32x32 diagonal block factored in shared memory; (2) every row below it solved against the block,
//   one thread per row, the result kept in shared memory transposed (Pt[c][row]); (3) trailing matrix (and the extra
//   rows) updated in 64x64 tiles with 4x4 register blocks, read-modify-write in L2.
constexpr int KC_NB = 32;
constexpr int KC_LD = KC_NB + 1;
constexpr int KC_THREADS = 256;

__host__ __device__ inline int kc_ldp(int n, int C) { return ((n + C + 63) / 64) * 64; }
__host__ __device__ inline size_t kc_smem_doubles(int n, int C) {
  const size_t panel = (size_t)KC_NB * kc_ldp(n, C);
  const size_t back = (size_t)n * C + (size_t)(KC_THREADS / 32) * KC_NB * 16;
  return (size_t)KC_NB * KC_LD + KC_NB + (panel > back ? panel : back);
}

__global__ void __launch_bounds__(KC_THREADS, 2)
kernelshap_solve_kernel(double* __restrict__ A, double* __restrict__ R, const double* __restrict__ fx,
                        const double* __restrict__ f0, int d, int C, int link, double* __restrict__ phi,
                        int* __restrict__ info) {
  extern __shared__ __align__(16) double ks_smem[];
  const int n = d - 1;
  const int b = blockIdx.x;
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  double* a = A + (long long)b * n * n;
  double* r = R + (long long)b * n * C;
  double* Lkk = ks_smem;                       // [32][33] diagonal block (identity padded)
  double* invd = Lkk + KC_NB * KC_LD;          // [32] reciprocals of its diagonal
  double* Pt = invd + KC_NB;                   // [32][ldp] panel, transposed   (16-byte aligned: 1088 doubles before)
  const int ldp = kc_ldp(n, C);
  const int next = n + C;                      // rows incl. the right-hand sides
  __shared__ int bad;
  if (t == 0) bad = 0;
  // element (i, j) of the extended matrix: rows >= n are the right-hand sides, stored (n, C) row-major
  auto at = [&](int i, int j) -> double* { return i < n ? a + (long long)i * n + j : r + (long long)j * C + (i - n); };

  for (int kb = 0; kb < n; kb += KC_NB) {
    const int nb = min(KC_NB, n - kb);
    __syncthreads();
    for (int e = t; e < KC_NB * KC_NB; e += KC_THREADS) {
      const int i = e >> 5, j = e & 31;
      double v = (i == j) ? 1.0 : 0.0;
      if (i < nb && j < nb) v = (j <= i) ? a[(long long)(kb + i) * n + kb + j] : 0.0;
      Lkk[i * KC_LD + j] = v;
    }
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
      const double v = Lkk[k * KC_LD + k];
      const bool ok = v > 0.0;
      const double piv = ok ? sqrt(v) : 1.0;
      const double inv = 1.0 / piv;
      if (!ok && t == 0 && bad == 0) bad = kb + k + 1;
      double upd[KC_NB * KC_NB / KC_THREADS];
#pragma unroll
      for (int u = 0; u < KC_NB * KC_NB / KC_THREADS; ++u) {
        const int e = t + u * KC_THREADS;
        const int i = e >> 5, j = e & 31;
        upd[u] = 0.0;
        if (j > k && j <= i && i < nb) upd[u] = (Lkk[i * KC_LD + k] * inv) * (Lkk[j * KC_LD + k] * inv);
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < KC_NB * KC_NB / KC_THREADS; ++u) {
        const int e = t + u * KC_THREADS;
        const int i = e >> 5, j = e & 31;
        if (j > k && j <= i && i < nb) Lkk[i * KC_LD + j] -= upd[u];
      }
      if (t == k) Lkk[k * KC_LD + k] = piv;
      else if (t > k && t < nb) Lkk[t * KC_LD + k] *= inv;
      __syncthreads();
    }
    for (int e = t; e < KC_NB * KC_NB; e += KC_THREADS) {
      const int i = e >> 5, j = e & 31;
      if (i < nb && j <= i) a[(long long)(kb + i) * n + kb + j] = Lkk[i * KC_LD + j];
    }
    if (t < KC_NB) invd[t] = 1.0 / Lkk[t * KC_LD + t];
    __syncthreads();
    const int r0 = kb + nb;
    const int mext = next - r0;                // rows below the block (matrix rows + right-hand sides)
    const int m = n - r0;                      // trailing columns
    for (int row = t; row < mext; row += KC_THREADS) {
      const int i = r0 + row;
      double x[KC_NB];
#pragma unroll
      for (int c = 0; c < KC_NB; ++c) x[c] = (c < nb) ? *at(i, kb + c) : 0.0;
#pragma unroll
      for (int c = 0; c < KC_NB; ++c) {
        double s = x[c];
#pragma unroll
        for (int p = 0; p < c; ++p) s = fma(-x[p], Lkk[c * KC_LD + p], s);
        x[c] = s * invd[c];
      }
#pragma unroll
      for (int c = 0; c < KC_NB; ++c) {
        if (c < nb) *at(i, kb + c) = x[c];
        Pt[c * ldp + row] = x[c];
      }
    }
    __syncthreads();
    if (m <= 0) continue;
    const int ty = t >> 4, tx = t & 15;
    for (int ti = 0; ti < mext; ti += 64) {
      const int jmax = min(ti + 64, m);        // tiles entirely above the diagonal are skipped
      for (int tc = 0; tc < jmax; tc += 64) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] = 0.0;
#pragma unroll 8
        for (int c = 0; c < KC_NB; ++c) {
          const double2* pr = reinterpret_cast<const double2*>(Pt + c * ldp + ti + ty * 4);
          const double2 r01 = pr[0], r23 = pr[1];
          const double rv[4] = {r01.x, r01.y, r23.x, r23.y};
          double cv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) cv[q] = Pt[c * ldp + tc + tx + 16 * q];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = fma(rv[i], cv[q], acc[i][q]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gi = r0 + ti + ty * 4 + i;
          if (gi >= next) continue;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int gj = r0 + tc + tx + 16 * q;
            if (gj < n && (gi >= n || gj <= gi)) *at(gi, gj) -= acc[i][q];
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- backward substitution L^T x = z (z sits in r), 32-row blocks from the bottom up ----
  double* xs = Pt;                             // x[j][c], (n, C)
  double* part = Pt + (size_t)n * C;           // [8 warps][32 columns][16 classes] partial dot products
  const int nblk = (n + KC_NB - 1) / KC_NB;
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int kb = blk * KC_NB;
    const int nb = min(KC_NB, n - kb);
    // partial sums over the rows below: warp w takes rows r0 + w, r0 + w + 8, ...; lane = column of the block
    double pacc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) pacc[c] = 0.0;
    for (int i = kb + nb + warp; i < n; i += KC_THREADS / 32) {
      const double lv = (lane < nb) ? a[(long long)i * n + kb + lane] : 0.0;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < C) pacc[c] = fma(lv, xs[(size_t)i * C + c], pacc[c]);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) part[(warp * KC_NB + lane) * 16 + c] = pacc[c];
    for (int e = t; e < KC_NB * KC_NB; e += KC_THREADS) {
      const int i = e >> 5, j = e & 31;
      double v = (i == j) ? 1.0 : 0.0;
      if (i < nb && j < nb) v = (j <= i) ? a[(long long)(kb + i) * n + kb + j] : 0.0;
      Lkk[i * KC_LD + j] = v;
    }
    __syncthreads();
    if (t < KC_NB) invd[t] = 1.0 / Lkk[t * KC_LD + t];
    __syncthreads();
    for (int c = warp; c < C; c += KC_THREADS / 32) {      // one warp per class, lane = row of the block
      double zv = 0.0, xv = 0.0;
      if (lane < nb) {
        zv = r[(long long)(kb + lane) * C + c];
#pragma unroll
        for (int ww = 0; ww < KC_THREADS / 32; ++ww) zv -= part[(ww * KC_NB + lane) * 16 + c];
      }
      for (int q = nb - 1; q >= 0; --q) {
        const double xq = __shfl_sync(0xffffffffu, zv, q) * invd[q];
        if (lane == q) xv = xq;
        if (lane < q) zv = fma(-Lkk[q * KC_LD + lane], xq, zv);
      }
      if (lane < nb) xs[(size_t)(kb + lane) * C + c] = xv;
    }
    __syncthreads();
  }
  for (int c = warp; c < C; c += KC_THREADS / 32) {
    double* out = phi + ((long long)b * C + c) * d;
    double tot = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double v = xs[(size_t)i * C + c];
      out[i] = v;
      tot += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) {
      const double l0 = ks_link(f0[c], link);
      out[n] = (ks_link(fx[(long long)b * C + c], link) - l0) - tot;
    }
  }
  if (t == 0 && info) info[b] = bad;
}

int kernelshap_solve(const uint32_t* Z, int words, const double* w, const double* probs, const double* fx,
                     const double* f0, int B, int S, int d, int C, int link, double* A, double* R, double* phi,
                     int* info, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && S > 0 && d >= 2 && C > 0 && C <= 16, "KernelSHAP shape (C <= 16, d >= 2)");
  AGB_REQUIRE(words * 32 >= d, "mask words");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(Z && w && probs && fx && f0 && A && R && phi, "null pointer");
  AGB_REQUIRE(B <= 65535, "batch too large (chunk it)");
  const int n = d - 1;
  const size_t smem = kc_smem_doubles(n, C) * sizeof(double);
  AGB_REQUIRE(smem <= 227 * 1024, "KernelSHAP: d too large for the shared-memory panel (d <= 1024)");
  const int tiles = (n + KG_T - 1) / KG_T;
  // Gram matrix: tensor cores (exact integer limbs, see kernelshap_gram_tc_kernel) unless AGB_KS_GRAM=fp64 asks for the
  // float64 FMA kernel; the right-hand sides E^T W y~ (real-valued y) always take the float64 kernel's rhs tiles
  static const bool use_tc = [] { const char* e = getenv("AGB_KS_GRAM"); return !(e != nullptr && e[0] == 'f'); }();
  const bool tc = use_tc && S <= 65536;
  if (tc) {
    static bool configured = false;
    if (!configured) {
      AGB_CHECK_CUDA(cudaFuncSetAttribute(kernelshap_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM));
      configured = true;
    }
    const int tjn = (n + KT_BM - 1) / KT_BM;
    dim3 gtc(tjn * (tjn + 1), B);                 // sum over tj of (2 tj + 2) k-tiles
    kernelshap_gram_tc_kernel<<<gtc, KT_THREADS, KT_SMEM, st>>>(Z, words, w, S, d, A);
    AGB_CHECK_CUDA(cudaGetLastError());
  }
  if (tc && words <= 32) {
    int G = (4 * sm_count()) / B;                 // few samples: slice the coalitions over several CTAs per sample
    G = G < 1 ? 1 : (G > 8 ? 8 : G);
    if (G > 1) AGB_CHECK_CUDA(cudaMemsetAsync(R, 0, sizeof(double) * (size_t)B * n * C, st));
    dim3 gr(B, G);
    if (C <= 4) kernelshap_rhs_kernel<4><<<gr, 256, 0, st>>>(Z, words, w, probs, fx, f0, S, d, C, link, R);
    else        kernelshap_rhs_kernel<16><<<gr, 256, 0, st>>>(Z, words, w, probs, fx, f0, S, d, C, link, R);
  } else {
    dim3 grid(tc ? tiles : tiles * (tiles + 1) / 2 + tiles, B);
    if (C <= 4) kernelshap_gram_kernel<4><<<grid, KG_THREADS, 0, st>>>(Z, words, w, probs, fx, f0, S, d, C, link, A, R, tc ? 1 : 0);
    else        kernelshap_gram_kernel<16><<<grid, KG_THREADS, 0, st>>>(Z, words, w, probs, fx, f0, S, d, C, link, A, R, tc ? 1 : 0);
  }
  AGB_CHECK_CUDA(cudaGetLastError());
  AGB_CHECK_CUDA(cudaFuncSetAttribute(kernelshap_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernelshap_solve_kernel<<<B, KC_THREADS, smem, st>>>(A, R, fx, f0, d, C, link, phi, info);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

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
                       double* __restrict__ R) {
  const int n = d - 1;
  const int tiles = (n + KG_T - 1) / KG_T;
  const int n_lower = tiles * (tiles + 1) / 2;
  const int b = blockIdx.y;
  const int tile = blockIdx.x;
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

__global__ void __launch_bounds__(KC_THREADS)
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
  dim3 grid(tiles * (tiles + 1) / 2 + tiles, B);
  if (C <= 4) kernelshap_gram_kernel<4><<<grid, KG_THREADS, 0, st>>>(Z, words, w, probs, fx, f0, S, d, C, link, A, R);
  else        kernelshap_gram_kernel<16><<<grid, KG_THREADS, 0, st>>>(Z, words, w, probs, fx, f0, S, d, C, link, A, R);
  AGB_CHECK_CUDA(cudaGetLastError());
  AGB_CHECK_CUDA(cudaFuncSetAttribute(kernelshap_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernelshap_solve_kernel<<<B, KC_THREADS, smem, st>>>(A, R, fx, f0, d, C, link, phi, info);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

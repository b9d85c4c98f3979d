// KernelSHAP weighted least squares, batched over explained samples (a14).
//
// In the reference the solve happens inside the third-party `shap.KernelExplainer` on the CPU in
// float64 numpy/LAPACK (call site reference models/kernel_shap_bert.py:170-185).  Here, per explained
// sample b: coalitions Z_b (S x d, packed bits), kernel weights w_b (S), background-averaged model outputs
// p_b (S x C), the sample's own output f_b (C) and the null output f0 (C) go in; attributions phi (C x d)
// come out, all on the device and in float64:
//     y = link(p) - link(f0),  delta = link(f) - link(f0)
//     E = Z[:, :-1] - Z[:, -1:],  y~ = y - Z[:, -1:] delta           (efficiency constraint eliminated)
//     A = E^T W E  ((d-1) x (d-1) Gram),  R = E^T W y~                (kernel 1: tiled accumulation)
//     A = L L^T (Cholesky), phi[:-1] = A^-1 R, phi[-1] = delta - sum  (kernel 2: one CTA per sample)
// shap's optional l1_reg feature pre-selection is NOT part of this path (documented in oracle/kernelshap.py).
// Bytes per sample (S=2048, d=127, C=2): packed Z 32 KB + p 16 KB in, A 127 KB out/in, phi 2 KB out.
#include "agb_common.cuh"

namespace agb {

constexpr int KS_TILE = 32;

__device__ __forceinline__ double ks_link(double p, int link) {
  return link ? log(p / (1.0 - p)) : p;
}

// grid: (tiles_j * tiles_k [+ rhs tiles], B).  Each CTA owns a 32x32 tile of A (or a 32 x C strip of R) and
// streams the S coalitions through shared memory in chunks.
__global__ void __launch_bounds__(KS_TILE * KS_TILE)
kernelshap_gram_kernel(const uint32_t* __restrict__ Z, int words, const double* __restrict__ w,
                       const double* __restrict__ probs, const double* __restrict__ fx,
                       const double* __restrict__ f0, int S, int d, int C, int link, double* __restrict__ A,
                       double* __restrict__ R) {
  const int n = d - 1;
  const int tiles = (n + KS_TILE - 1) / KS_TILE;
  const int b = blockIdx.y;
  const int tile = blockIdx.x;
  const bool is_rhs = tile >= tiles * tiles;
  const int tj = is_rhs ? (tile - tiles * tiles) : tile / tiles;
  const int tk = is_rhs ? 0 : tile % tiles;
  if (!is_rhs && tk > tj) return;  // symmetric: lower triangle only, mirrored by the solve kernel's reads
  const int tx = threadIdx.x % KS_TILE, ty = threadIdx.x / KS_TILE;
  const int j = tj * KS_TILE + ty;      // row of A / row of R
  const int k = tk * KS_TILE + tx;      // col of A, or class index for the rhs strip
  constexpr int CH = 96;                // coalitions per shared-memory chunk
  __shared__ float ej[CH][KS_TILE + 1];   // E[s, j-tile] in {-1,0,1}
  __shared__ float ek[CH][KS_TILE + 1];   // E[s, k-tile]
  __shared__ double ws[CH];
  __shared__ double yt[CH][16];           // y~[s, c] for the rhs strip (C <= 16)
  const uint32_t* Zb = Z + (long long)b * S * words;
  const int last = d - 1;
  double acc = 0.0;
  for (int s0 = 0; s0 < S; s0 += CH) {
    __syncthreads();
    for (int e = threadIdx.x; e < CH * KS_TILE; e += blockDim.x) {
      const int ss = e / KS_TILE, c = e % KS_TILE;
      const int s = s0 + ss;
      float vj = 0.f, vk = 0.f;
      if (s < S) {
        const uint32_t* zr = Zb + (long long)s * words;
        const int zl = (zr[last >> 5] >> (last & 31)) & 1;
        const int fj = tj * KS_TILE + c, fk = tk * KS_TILE + c;
        if (fj < n) vj = (float)((int)((zr[fj >> 5] >> (fj & 31)) & 1) - zl);
        if (fk < n) vk = (float)((int)((zr[fk >> 5] >> (fk & 31)) & 1) - zl);
      }
      ej[ss][c] = vj;
      ek[ss][c] = vk;
    }
    for (int ss = threadIdx.x; ss < CH; ss += blockDim.x) {
      const int s = s0 + ss;
      ws[ss] = (s < S) ? w[(long long)b * S + s] : 0.0;
    }
    if (is_rhs) {
      for (int e = threadIdx.x; e < CH * C; e += blockDim.x) {
        const int ss = e / C, c = e % C;
        const int s = s0 + ss;
        double v = 0.0;
        if (s < S) {
          const uint32_t* zr = Zb + (long long)s * words;
          const double zl = (double)((zr[last >> 5] >> (last & 31)) & 1);
          const double l0 = ks_link(f0[c], link);
          const double yv = ks_link(probs[((long long)b * S + s) * C + c], link) - l0;
          const double dl = ks_link(fx[(long long)b * C + c], link) - l0;
          v = yv - zl * dl;
        }
        yt[ss][c] = v;
      }
    }
    __syncthreads();
    if (!is_rhs) {
#pragma unroll 8
      for (int ss = 0; ss < CH; ++ss) acc += ws[ss] * (double)(ej[ss][ty] * ek[ss][tx]);
    } else if (tx < C) {
#pragma unroll 8
      for (int ss = 0; ss < CH; ++ss) acc += ws[ss] * (double)ej[ss][ty] * yt[ss][tx];
    }
  }
  if (j < n) {
    if (!is_rhs) {
      if (k < n) A[((long long)b * n + j) * n + k] = acc;
    } else if (tx < C) {
      R[((long long)b * n + j) * C + tx] = acc;
    }
  }
}

// One CTA per sample: in-place Cholesky of the lower triangle (right-looking, column by column), forward and
// backward substitution for the C right-hand sides, then the eliminated feature.  The matrix stays in L1/L2.
__global__ void __launch_bounds__(256)
kernelshap_solve_kernel(double* __restrict__ A, double* __restrict__ R, const double* __restrict__ fx,
                        const double* __restrict__ f0, int d, int C, int link, double* __restrict__ phi,
                        int* __restrict__ info) {
  const int n = d - 1;
  const int b = blockIdx.x;
  double* a = A + (long long)b * n * n;
  double* r = R + (long long)b * n * C;
  __shared__ double piv;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) {
      const double v = a[(long long)k * n + k];
      if (!(v > 0.0)) { bad = k + 1; piv = 1.0; }
      else piv = sqrt(v);
      a[(long long)k * n + k] = piv;
    }
    __syncthreads();
    const double inv = 1.0 / piv;
    for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) a[(long long)i * n + k] *= inv;
    __syncthreads();
    // trailing update of the lower triangle: a[i][j] -= a[i][k] * a[j][k],  k < j <= i < n
    const int m = n - k - 1;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = k + 1 + e / m, jj = k + 1 + e % m;
      if (jj <= i) a[(long long)i * n + jj] -= a[(long long)i * n + k] * a[(long long)jj * n + k];
    }
    __syncthreads();
  }
  // L z = R (forward), L^T x = z (backward); one thread per class, columns are short (n <= 511)
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    for (int i = 0; i < n; ++i) {
      double s = r[(long long)i * C + c];
      for (int jj = 0; jj < i; ++jj) s -= a[(long long)i * n + jj] * r[(long long)jj * C + c];
      r[(long long)i * C + c] = s / a[(long long)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = r[(long long)i * C + c];
      for (int jj = i + 1; jj < n; ++jj) s -= a[(long long)jj * n + i] * r[(long long)jj * C + c];
      r[(long long)i * C + c] = s / a[(long long)i * n + i];
    }
    double tot = 0.0;
    double* out = phi + ((long long)b * C + c) * d;
    for (int i = 0; i < n; ++i) { out[i] = r[(long long)i * C + c]; tot += out[i]; }
    const double l0 = ks_link(f0[c], link);
    out[n] = (ks_link(fx[(long long)b * C + c], link) - l0) - tot;
  }
  __syncthreads();
  if (threadIdx.x == 0 && info) info[b] = bad;
}

int kernelshap_solve(const uint32_t* Z, int words, const double* w, const double* probs, const double* fx,
                     const double* f0, int B, int S, int d, int C, int link, double* A, double* R, double* phi,
                     int* info, cudaStream_t st) {
  AGB_REQUIRE(B >= 0 && S > 0 && d >= 2 && C > 0 && C <= 16, "KernelSHAP shape (C <= 16, d >= 2)");
  AGB_REQUIRE(words * 32 >= d, "mask words");
  if (B == 0) return AGB_OK;
  AGB_REQUIRE(Z && w && probs && fx && f0 && A && R && phi, "null pointer");
  const int n = d - 1;
  const int tiles = (n + KS_TILE - 1) / KS_TILE;
  dim3 grid(tiles * tiles + tiles, B);
  AGB_REQUIRE(B <= 65535, "batch too large (chunk it)");
  kernelshap_gram_kernel<<<grid, KS_TILE * KS_TILE, 0, st>>>(Z, words, w, probs, fx, f0, S, d, C, link, A, R);
  AGB_CHECK_CUDA(cudaGetLastError());
  kernelshap_solve_kernel<<<B, 256, 0, st>>>(A, R, fx, f0, d, C, link, phi, info);
  AGB_CHECK_CUDA(cudaGetLastError());
  return AGB_OK;
}

}  // namespace agb

// extern "C" surface of libautognothi_b200.so — thin argument checks + dispatch to the launchers.
// Signatures are declared (with reference citations) in include/autognothi_b200.h.
#include "../../include/autognothi_b200.h"

#include "agb_common.cuh"

namespace agb {
const char* last_error();
int gemm_bf16_tc(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N,
                 int K, float alpha, const float* bias, int act, const bf16* res_bf16,
                 const float* res_f32, int ldr, int res_group, int res_rows, void* out, int ldo,
                 int out_f32, cudaStream_t stream);
int rank_masks(const float* scores, int use_philox, uint64_t seed, uint64_t offset, int rows, int n, const int* stops,
               int nstops, int mask_base, uint32_t* packed, int words, int64_t* dense, cudaStream_t st);
int cls_attention(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off, int io_bf16,
                  const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode, void* ctx, long long ldc,
                  cudaStream_t st);
int mask_counts(const uint32_t* packed, int rows, int words, int T, int* counts, cudaStream_t st);
int packed_token_index(const uint32_t* packed, int rows, int words, int T, int S, const int* cu, long long* src,
                       cudaStream_t st);
int attention_pipe_prefix(const bf16* qkv, const int* nkeep, int rows, int T, int H, int heads, bf16* ctx, cudaStream_t stream);
int attention_narrow_varlen(const bf16* qkv, const int* cu, int rows, int max_len, int H, int heads, bf16* ctx, cudaStream_t st);
int attention_varlen(const bf16* qkv, const int* cu, int rows, int max_len, int total_tokens, int H, int heads, bf16* ctx,
                     cudaStream_t stream);
int cls_attention_varlen(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off, int io_bf16,
                         const int* cu, int rows, int max_len, int H, int heads, void* ctx, long long ldc, cudaStream_t st);
void set_attention_variant(int v);
void set_attention_bwd_variant(int v);
int get_attention_bwd_variant();
void set_attention_trace(long long* t);
int get_attention_variant();
int gather_token_rows_hilo(const float* src, const uint8_t* order, int rows, int T, int S, int H, bf16* hi, bf16* lo,
                           cudaStream_t st);
int split_hilo(const float* x, long long n, bf16* hi, bf16* lo, cudaStream_t st);
int rowstats_cast(const float* x, long long ldx, int rows, int H, bf16* out, long long ldo, float* stats,
                  cudaStream_t st);
void set_gemm_variant(int v);
int get_gemm_variant();
int gemm_f32(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, int M, int N, int K, float alpha,
             const float* bias, int act, const float* res, int ldr, float* C, int ldc, cudaStream_t st);
int gelu_fwd(const void* z, void* out, long long n, int is_bf16, cudaStream_t st);
int gelu_bwd(const void* dy, const void* z, void* dz, long long n, int is_bf16, cudaStream_t st);
int colsum(const void* y, int is_bf16, long long ld, int M, int N, float* out, cudaStream_t st);
int layernorm_bwd(const void* x, int x_bf16, const void* dy, int dy_bf16, const float* gamma, const float* dres,
                  int rows, int H, float eps, float* dx, float* dgamma, float* dbeta, cudaStream_t st);
int dropout(const void* y, int y_bf16, const float* residual, void* out, int out_bf16, long long n, unsigned thr16,
            unsigned long long seed, unsigned tag, cudaStream_t st);
int attention_dropout_mask(unsigned char* keep, int rows, int heads, int T, unsigned thr16, unsigned long long seed,
                           cudaStream_t st);
int attention_pipe_dropout(const bf16* qkv, const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode,
                           bf16* ctx, unsigned thr16, unsigned long long seed, cudaStream_t stream);
int attention_bwd(const void* qkv, const void* dctx, int io_bf16, const uint32_t* mask, int words, int rows, int T,
                  int H, int heads, int mode, void* dqkv, cudaStream_t st, unsigned drop_thr, unsigned long long drop_seed);
int vit_embed_bwd(const float* dx, int B, int T, int H, float* dpos, float* dcls, void* dpatch, int dpatch_bf16,
                  cudaStream_t st);
int bert_embed_sum(const int64_t* ids, const float* word, const float* pos, const float* type0, int BT, int T, int H,
                   int vocab, float* out, cudaStream_t st);
int bert_embed_scatter(const int64_t* ids, const float* dsum, int BT, int T, int H, int vocab, int pad_id,
                       float* dword, float* dpos, float* dtype0, cudaStream_t st);
int pack_masks_i64(const int64_t* mask, int rows, int n, int prepend_cls, uint32_t* packed, int words,
                   cudaStream_t st);
int unpack_masks_i64(const uint32_t* packed, int rows, int n, int skip, int words, int64_t* out,
                     cudaStream_t st);
int shapley_masks(const float* u_players, const float* u_size, const float* prefix, int use_philox,
                  uint64_t seed, uint64_t offset, int pairs, int n, uint32_t* packed, int words,
                  int64_t* dense, cudaStream_t st);
int uniform_masks(const float* u_players, const float* u_row, int use_philox, uint64_t seed,
                  uint64_t offset, int rows, int n, uint32_t* packed, int words, int64_t* dense,
                  cudaStream_t st);
int layernorm(const void* x, int in_bf16, long long in_stride, int rows, int H, const float* gamma,
              const float* beta, float eps, bf16* out_bf16, float* out_f32, long long out_stride,
              cudaStream_t st);
int cast_f32_to_bf16(const float* in, bf16* out, long long n, cudaStream_t st);
int vit_im2col(const float* img, int B, int C, int px, int P, void* out, int out_bf16, cudaStream_t st);
int repeat_rows(const void* src, int B, long long row_bytes, int S, void* dst, cudaStream_t st);
int kept_first_order(const uint32_t* packed, int rows, int words, int T, uint8_t* order, uint8_t* pos, int* nkeep,
                     uint32_t* prefix, cudaStream_t st);
int gather_token_rows(const void* src, const uint8_t* order, int rows, int T, int S, long long row_bytes, void* dst,
                      cudaStream_t st);
int attention_scatter(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                      const uint8_t* dst_pos, bf16* ctx, cudaStream_t stream);
int vit_assemble(const float* patch_emb, const float* cls, const float* pos, int B, int S, int T, int H,
                 float* x, cudaStream_t st);
int bert_embed(const int64_t* ids, const float* word, const float* pos, const float* type0,
               const float* gamma, const float* beta, float eps, int B, int S, int T, int H, int vocab,
               float* x, cudaStream_t st);
int cls_head(const float* x, long long row_stride, int rows, int H, int C, int mode, const float* ln_g,
             const float* ln_b, float eps, const float* wp, const float* bp, const float* wc,
             const float* bc, float* probs, float* logits_out, cudaStream_t st);
int attention_simt(const void* qkv, int io_bf16, const uint32_t* mask, int words, int rows, int T, int H,
                   int heads, int mode, void* ctx, cudaStream_t st, unsigned drop_thr, unsigned long long drop_seed);
int attention_tc(const bf16* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H, int heads,
                 int mode, bf16* ctx, cudaStream_t st);
int explainer_head_fwd(const void* h, int h_bf16, int B, int T, int E, int C, const float* W,
                       const float* bias, const float* grand, const float* null_v, int normalize,
                       float* phi, float* pred_out, cudaStream_t st);
int explainer_head_bwd(const float* dphi, const void* h, int h_bf16, int B, int T, int E, int C,
                       const float* W, int normalize, void* dh, float* dW, float* db, cudaStream_t st);
int normalize_shapley(const float* pred, const float* grand, const float* null_v, int B, int T, int C,
                      float* out, cudaStream_t st);
int shapley_loss_fwd(const uint32_t* packed, int words, const float* v0, const float* v_s, const float* phi,
                     int B, int S, int n, int C, float* resid, float* partial, float* loss, cudaStream_t st);
int shapley_loss_bwd(const uint32_t* packed, int words, const float* resid, const float* gout, int B, int S,
                     int n, int C, float* dphi, cudaStream_t st);
int kernelshap_solve(const uint32_t* Z, int words, const double* w, const double* probs, const double* fx,
                     const double* f0, int B, int S, int d, int C, int link, double* A, double* R, double* phi,
                     int* info, cudaStream_t st);
}  // namespace agb

using agb::bf16;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

const char* agb_last_error(void) { return agb::last_error(); }
int agb_version(void) { return 100; }

int agb_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major,
                  int M, int N, int K, float alpha, const float* bias, int act,
                  const void* residual_bf16, const float* residual_f32, int ldr, int res_group,
                  int res_rows, void* out, int ldo, int out_is_f32, void* stream) {
  return agb::gemm_bf16_tc(static_cast<const bf16*>(A), lda, a_mn_major, static_cast<const bf16*>(B),
                           ldb, b_mn_major, M, N, K, alpha, bias, act,
                           static_cast<const bf16*>(residual_bf16), residual_f32, ldr, res_group,
                           res_rows, out, ldo, out_is_f32, static_cast<cudaStream_t>(stream));
}


int agb_rank_masks(const float* scores, int use_philox, uint64_t seed, uint64_t offset, int rows, int n_players,
                   const int* stops, int nstops, int mask_base, uint32_t* packed, int words, int64_t* dense,
                   void* stream) {
  return agb::rank_masks(scores, use_philox, seed, offset, rows, n_players, stops, nstops, mask_base, packed, words, dense,
                         ST(stream));
}

int agb_cls_attention(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off, int io_is_bf16,
                      const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode, void* ctx,
                      long long ldc, void* stream) {
  return agb::cls_attention(q, ldq, kv, ldkv, k_off, v_off, io_is_bf16, mask, words, rows, T, H, heads, mode, ctx, ldc,
                            ST(stream));
}

int agb_mask_counts(const uint32_t* packed, int rows, int words, int T, int* counts, void* stream) {
  return agb::mask_counts(packed, rows, words, T, counts, ST(stream));
}
int agb_packed_token_index(const uint32_t* packed, int rows, int words, int T, int S, const int* cu, int64_t* src,
                           void* stream) {
  return agb::packed_token_index(packed, rows, words, T, S, cu, reinterpret_cast<long long*>(src), ST(stream));
}
int agb_attention_bf16_varlen(const void* qkv, const int* cu, int rows, int max_len, int total_tokens, int H, int heads,
                              void* ctx, void* stream) {
  if (heads > 0 && H != heads * 64)      // narrow heads (LTT side ladder): CUDA-core kernel, one CTA per (row, head)
    return agb::attention_narrow_varlen(static_cast<const bf16*>(qkv), cu, rows, max_len, H, heads, static_cast<bf16*>(ctx),
                                        ST(stream));
  return agb::attention_varlen(static_cast<const bf16*>(qkv), cu, rows, max_len, total_tokens, H, heads,
                               static_cast<bf16*>(ctx), ST(stream));
}
int agb_attention_bf16_prefix(const void* qkv, const int* nkeep, int rows, int T, int H, int heads, void* ctx, void* stream) {
  return agb::attention_pipe_prefix(static_cast<const bf16*>(qkv), nkeep, rows, T, H, heads, static_cast<bf16*>(ctx), ST(stream));
}
int agb_cls_attention_varlen(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off,
                             int io_is_bf16, const int* cu, int rows, int max_len, int H, int heads, void* ctx,
                             long long ldc, void* stream) {
  return agb::cls_attention_varlen(q, ldq, kv, ldkv, k_off, v_off, io_is_bf16, cu, rows, max_len, H, heads, ctx, ldc,
                                   ST(stream));
}

int agb_attention_set_variant(int variant) {
  const int prev = agb::get_attention_variant();
  agb::set_attention_variant(variant);
  return prev;
}

int agb_attention_set_trace(void* trace) {
  agb::set_attention_trace(static_cast<long long*>(trace));
  return AGB_OK;
}

int agb_attention_bwd_set_variant(int variant) {
  const int prev = agb::get_attention_bwd_variant();
  agb::set_attention_bwd_variant(variant);
  return prev;
}

int agb_gemm_bf16_fused(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                        int act, const float* residual_f32, int ldr, void* out, int ldo, int out_is_f32,
                        const float* ln_stats, int ln_parts, const float* ln_colsum, float ln_eps,
                        void* out_bf16_copy, int ldo_copy, float* stats_out, void* stream) {
  AGB_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && out, "operands");
  AGB_REQUIRE((N % 4) == 0 && (lda % 8) == 0 && (ldb % 8) == 0, "alignment");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "alignment");
  agb::PairGemmCall c;
  c.A = static_cast<const bf16*>(A); c.lda = lda; c.B = static_cast<const bf16*>(B); c.ldb = ldb; c.M = M; c.N = N; c.K = K;
  c.bias = bias; c.act = act; c.res_f32 = residual_f32; c.ldr = ldr; c.out = out; c.ldo = ldo; c.out_f32 = out_is_f32;
  c.ln_stats = ln_stats; c.ln_parts = ln_parts; c.ln_colsum = ln_colsum; c.ln_eps = ln_eps;
  c.out16 = static_cast<bf16*>(out_bf16_copy); c.ldo16 = ldo_copy; c.stats_out = stats_out;
  const int rc = agb::gemm_bf16_pair_call(c, ST(stream));
  if (rc == AGB_ERR_UNSUPPORTED)
    agb::set_last_error("agb_gemm_bf16_fused: shape / epilogue combination not covered (M=%d N=%d K=%d act=%d)", M, N, K, act);
  return rc;
}
int agb_gemm_bf16_hilo(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, void* x_hi,
                       void* x_lo, int ldx, float* stats_out, void* stream) {
  AGB_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && x_hi && x_lo && stats_out, "operands");
  AGB_REQUIRE((N % 256) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 && (ldx % 8) == 0, "alignment");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "alignment");
  agb::PairGemmCall c;
  c.A = static_cast<const bf16*>(A); c.lda = lda; c.B = static_cast<const bf16*>(B); c.ldb = ldb; c.M = M; c.N = N; c.K = K;
  c.bias = bias; c.stats_out = stats_out;
  c.hl_hi = static_cast<bf16*>(x_hi); c.hl_lo = static_cast<bf16*>(x_lo); c.ld_hl = ldx;
  const int rc = agb::gemm_bf16_pair_call(c, ST(stream));
  if (rc == AGB_ERR_UNSUPPORTED)
    agb::set_last_error("agb_gemm_bf16_hilo: shape not covered (M=%d N=%d K=%d)", M, N, K);
  return rc;
}
int agb_gemm_bf16_dropout_residual(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                                   const float* residual_f32, int ldr, float* out, int thr16, uint64_t seed, int tag, void* stream) {
  AGB_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && out && residual_f32, "operands");
  AGB_REQUIRE(thr16 > 0 && thr16 < 65536, "dropout threshold");
  AGB_REQUIRE((N % 4) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 && (ldr % 4) == 0, "alignment");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "alignment");
  const unsigned key = agb::agb_drop_key(seed, (uint32_t)tag, 0x5bd1e995u);     // the stream of agb_dropout(seed, tag)
  agb::PairGemmCall c;
  c.A = static_cast<const bf16*>(A); c.lda = lda; c.B = static_cast<const bf16*>(B); c.ldb = ldb; c.M = M; c.N = N; c.K = K;
  c.bias = bias; c.res_f32 = residual_f32; c.ldr = ldr; c.out = out; c.ldo = N; c.out_f32 = 1;
  c.drop_thr = (unsigned)thr16; c.drop_key = key;
  const int rc = agb::gemm_bf16_pair_call(c, ST(stream));
  if (rc == AGB_ERR_UNSUPPORTED)
    agb::set_last_error("agb_gemm_bf16_dropout_residual: shape not covered (M=%d N=%d K=%d)", M, N, K);
  return rc;
}
int agb_gemm_bf16_gelu_dual(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, void* z_out,
                            void* gelu_out, int ldo, void* stream) {
  AGB_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && z_out && gelu_out, "operands");
  AGB_REQUIRE((N % 4) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 && (ldo % 8) == 0, "alignment");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "alignment");
  agb::PairGemmCall c;
  c.A = static_cast<const bf16*>(A); c.lda = lda; c.B = static_cast<const bf16*>(B); c.ldb = ldb; c.M = M; c.N = N; c.K = K;
  c.bias = bias; c.act = 2; c.out = gelu_out; c.ldo = ldo; c.z_out = static_cast<bf16*>(z_out); c.ldz_out = ldo;
  const int rc = agb::gemm_bf16_pair_call(c, ST(stream));
  if (rc == AGB_ERR_UNSUPPORTED)
    agb::set_last_error("agb_gemm_bf16_gelu_dual: shape not covered (M=%d N=%d K=%d)", M, N, K);
  return rc;
}
int agb_gemm_bf16_gelu_bwd(const void* dY, int ldy, const void* W, int ldw, int M, int N, int K, const void* z, int ldz,
                           void* dz, int ldo, void* stream) {
  AGB_REQUIRE(M > 0 && N > 0 && K > 0 && dY && W && z && dz, "operands");
  AGB_REQUIRE((N % 4) == 0 && (ldy % 8) == 0 && (ldw % 8) == 0 && (ldo % 8) == 0 && (ldz % 8) == 0, "alignment");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(dY) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "alignment");
  // W is the forward weight [K, N] (out_features K x in_features N): the MN-major B operand of dY W
  agb::PairGemmCall c;
  c.A = static_cast<const bf16*>(dY); c.lda = ldy; c.B = static_cast<const bf16*>(W); c.ldb = ldw; c.b_mn = 1;
  c.M = M; c.N = N; c.K = K; c.out = dz; c.ldo = ldo; c.z_in = static_cast<const bf16*>(z); c.ldz_in = ldz;
  const int rc = agb::gemm_bf16_pair_call(c, ST(stream));
  if (rc == AGB_ERR_UNSUPPORTED)
    agb::set_last_error("agb_gemm_bf16_gelu_bwd: shape not covered (M=%d N=%d K=%d)", M, N, K);
  return rc;
}
int agb_split_hilo(const float* x, long long n, void* hi, void* lo, void* stream) {
  return agb::split_hilo(x, n, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ST(stream));
}
int agb_gemm_stats_parts(int N) { return 2 * ((N + 255) / 256); }
int agb_rowstats_cast(const float* x, long long ldx, int rows, int H, void* out_bf16, long long ldo, float* stats,
                      void* stream) {
  return agb::rowstats_cast(x, ldx, rows, H, static_cast<bf16*>(out_bf16), ldo, stats, ST(stream));
}

int agb_gemm_set_variant(int variant) {
  const int prev = agb::get_gemm_variant();
  agb::set_gemm_variant(variant);
  return prev;
}

int agb_gemm_f32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, int M, int N,
                 int K, float alpha, const float* bias, int act, const float* residual, int ldr, float* C,
                 int ldc, void* stream) {
  return agb::gemm_f32(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, alpha, bias, act, residual, ldr, C, ldc,
                       ST(stream));
}
int agb_gelu_fwd(const void* z, void* out, long long n, int is_bf16, void* stream) {
  return agb::gelu_fwd(z, out, n, is_bf16, ST(stream));
}
int agb_gelu_bwd(const void* dy, const void* z, void* dz, long long n, int is_bf16, void* stream) {
  return agb::gelu_bwd(dy, z, dz, n, is_bf16, ST(stream));
}
int agb_colsum(const void* y, int is_bf16, long long ld, int M, int N, float* out, void* stream) {
  return agb::colsum(y, is_bf16, ld, M, N, out, ST(stream));
}
int agb_layernorm_bwd(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, const float* gamma,
                      const float* dres, int rows, int H, float eps, float* dx, float* dgamma, float* dbeta,
                      void* stream) {
  return agb::layernorm_bwd(x, x_is_bf16, dy, dy_is_bf16, gamma, dres, rows, H, eps, dx, dgamma, dbeta, ST(stream));
}
int agb_masked_attention_bwd(const void* qkv, const void* dctx, int io_is_bf16, const uint32_t* mask, int words,
                             int rows, int T, int H, int heads, int mode, void* dqkv, void* stream) {
  return agb::attention_bwd(qkv, dctx, io_is_bf16, mask, words, rows, T, H, heads, mode, dqkv, ST(stream), 0, 0);
}
int agb_dropout(const void* y, int y_is_bf16, const float* residual, void* out, int out_is_bf16, long long n,
                int thr16, uint64_t seed, int tag, void* stream) {
  return agb::dropout(y, y_is_bf16, residual, out, out_is_bf16, n, (unsigned)thr16, seed, (unsigned)tag, ST(stream));
}
int agb_attention_dropout_mask(void* keep, int rows, int heads, int T, int thr16, uint64_t seed,
                               void* stream) {
  return agb::attention_dropout_mask(static_cast<unsigned char*>(keep), rows, heads, T, (unsigned)thr16, seed, ST(stream));
}
int agb_masked_attention_dropout_fwd(const void* qkv, int io_is_bf16, const uint32_t* mask, int words, int rows, int T, int H,
                                     int heads, int mode, void* ctx, int thr16, uint64_t seed, void* stream) {
  if (io_is_bf16 && H == heads * 64)
    return agb::attention_pipe_dropout(static_cast<const bf16*>(qkv), mask, words, rows, T, H, heads, mode,
                                       static_cast<bf16*>(ctx), (unsigned)thr16, seed, ST(stream));
  return agb::attention_simt(qkv, io_is_bf16, mask, words, rows, T, H, heads, mode, ctx, ST(stream), (unsigned)thr16, seed);
}
int agb_masked_attention_dropout_bwd(const void* qkv, const void* dctx, int io_is_bf16, const uint32_t* mask, int words,
                                     int rows, int T, int H, int heads, int mode, void* dqkv, int thr16, uint64_t seed,
                                     void* stream) {
  return agb::attention_bwd(qkv, dctx, io_is_bf16, mask, words, rows, T, H, heads, mode, dqkv, ST(stream), (unsigned)thr16,
                            seed);
}
int agb_vit_embed_bwd(const float* dx, int B, int T, int H, float* dpos, float* dcls, void* dpatch,
                      int dpatch_is_bf16, void* stream) {
  return agb::vit_embed_bwd(dx, B, T, H, dpos, dcls, dpatch, dpatch_is_bf16, ST(stream));
}
int agb_bert_embed_sum(const int64_t* ids, const float* word, const float* pos, const float* type0, int BT, int T,
                       int H, int vocab, float* out, void* stream) {
  return agb::bert_embed_sum(ids, word, pos, type0, BT, T, H, vocab, out, ST(stream));
}
int agb_bert_embed_scatter(const int64_t* ids, const float* dsum, int BT, int T, int H, int vocab, int pad_id,
                           float* dword, float* dpos, float* dtype0, void* stream) {
  return agb::bert_embed_scatter(ids, dsum, BT, T, H, vocab, pad_id, dword, dpos, dtype0, ST(stream));
}
int agb_pack_masks_i64(const int64_t* mask, int rows, int n_players, int prepend_cls,
                       uint32_t* packed, int words, void* stream) {
  return agb::pack_masks_i64(mask, rows, n_players, prepend_cls, packed, words, ST(stream));
}
int agb_unpack_masks_i64(const uint32_t* packed, int rows, int n, int skip, int words, int64_t* out,
                         void* stream) {
  return agb::unpack_masks_i64(packed, rows, n, skip, words, out, ST(stream));
}
int agb_shapley_masks(const float* u_players, const float* u_size, const float* prefix,
                      int use_philox, uint64_t seed, uint64_t offset, int pairs, int n_players,
                      uint32_t* packed, int words, int64_t* dense, void* stream) {
  return agb::shapley_masks(u_players, u_size, prefix, use_philox, seed, offset, pairs, n_players, packed,
                            words, dense, ST(stream));
}
int agb_uniform_masks(const float* u_players, const float* u_row, int use_philox, uint64_t seed,
                      uint64_t offset, int rows, int n_players, uint32_t* packed, int words,
                      int64_t* dense, void* stream) {
  return agb::uniform_masks(u_players, u_row, use_philox, seed, offset, rows, n_players, packed, words,
                            dense, ST(stream));
}
int agb_layernorm(const void* x, int x_is_bf16, long long in_stride, int rows, int H,
                  const float* gamma, const float* beta, float eps, void* out_bf16, float* out_f32,
                  long long out_stride, void* stream) {
  return agb::layernorm(x, x_is_bf16, in_stride, rows, H, gamma, beta, eps, static_cast<bf16*>(out_bf16),
                        out_f32, out_stride, ST(stream));
}
int agb_cast_f32_to_bf16(const float* in, void* out_bf16, long long n, void* stream) {
  return agb::cast_f32_to_bf16(in, static_cast<bf16*>(out_bf16), n, ST(stream));
}
int agb_vit_im2col(const float* images, int B, int C, int px, int P, void* out, int out_is_bf16,
                   void* stream) {
  return agb::vit_im2col(images, B, C, px, P, out, out_is_bf16, ST(stream));
}
int agb_vit_assemble(const float* patch_emb, const float* cls_token, const float* pos_emb, int B,
                     int S, int T, int H, float* x, void* stream) {
  return agb::vit_assemble(patch_emb, cls_token, pos_emb, B, S, T, H, x, ST(stream));
}
int agb_repeat_rows(const void* src, int B, long long row_bytes, int S, void* dst, void* stream) {
  return agb::repeat_rows(src, B, row_bytes, S, dst, ST(stream));
}
int agb_kept_first_order(const uint32_t* packed, int rows, int words, int T, uint8_t* order, uint8_t* pos, int* nkeep,
                         uint32_t* prefix, void* stream) {
  return agb::kept_first_order(packed, rows, words, T, order, pos, nkeep, prefix, ST(stream));
}
int agb_gather_token_rows(const void* src, const uint8_t* order, int rows, int T, int S, long long row_bytes, void* dst,
                          void* stream) {
  return agb::gather_token_rows(src, order, rows, T, S, row_bytes, dst, ST(stream));
}
int agb_gather_token_rows_hilo(const float* src, const uint8_t* order, int rows, int T, int S, int H, void* hi, void* lo,
                               void* stream) {
  return agb::gather_token_rows_hilo(src, order, rows, T, S, H, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ST(stream));
}
int agb_masked_attention_bf16_scatter(const void* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H,
                                      int heads, const uint8_t* dst_pos, void* ctx, void* stream) {
  return agb::attention_scatter(static_cast<const bf16*>(qkv), mask, words, rows, share, T, H, heads, dst_pos,
                                static_cast<bf16*>(ctx), ST(stream));
}
int agb_bert_embed(const int64_t* ids, const float* word, const float* pos, const float* type0,
                   const float* gamma, const float* beta, float eps, int B, int S, int T, int H,
                   int vocab, float* x, void* stream) {
  return agb::bert_embed(ids, word, pos, type0, gamma, beta, eps, B, S, T, H, vocab, x, ST(stream));
}
int agb_cls_head(const float* x, long long row_stride, int rows, int H, int C, int mode,
                 const float* ln_gamma, const float* ln_beta, float eps, const float* w_pool,
                 const float* b_pool, const float* w_cls, const float* b_cls, float* probs,
                 float* logits_or_null, void* stream) {
  return agb::cls_head(x, row_stride, rows, H, C, mode, ln_gamma, ln_beta, eps, w_pool, b_pool, w_cls,
                       b_cls, probs, logits_or_null, ST(stream));
}
int agb_masked_attention_simt(const void* qkv, int io_is_bf16, const uint32_t* mask, int words,
                              int rows, int T, int H, int heads, int mode, void* ctx, void* stream) {
  return agb::attention_simt(qkv, io_is_bf16, mask, words, rows, T, H, heads, mode, ctx, ST(stream), 0, 0);
}
int agb_masked_attention_bf16(const void* qkv, const uint32_t* mask, int words, int rows, int T,
                              int H, int heads, int mode, void* ctx, void* stream) {
  return agb::attention_tc(static_cast<const bf16*>(qkv), mask, words, rows, 1, T, H, heads, mode,
                           static_cast<bf16*>(ctx), ST(stream));
}
int agb_masked_attention_bf16_shared(const void* qkv, const uint32_t* mask, int words, int rows, int share, int T,
                                     int H, int heads, int mode, void* ctx, void* stream) {
  return agb::attention_tc(static_cast<const bf16*>(qkv), mask, words, rows, share, T, H, heads, mode,
                           static_cast<bf16*>(ctx), ST(stream));
}
int agb_explainer_head_fwd(const void* h, int h_is_bf16, int B, int T, int E, int C, const float* W,
                           const float* bias, const float* grand, const float* null_v,
                           int normalize, float* phi, float* pred_or_null, void* stream) {
  return agb::explainer_head_fwd(h, h_is_bf16, B, T, E, C, W, bias, grand, null_v, normalize, phi,
                                 pred_or_null, ST(stream));
}
int agb_explainer_head_bwd(const float* dphi, const void* h, int h_is_bf16, int B, int T, int E,
                           int C, const float* W, int normalize, void* dh, float* dW, float* db,
                           void* stream) {
  return agb::explainer_head_bwd(dphi, h, h_is_bf16, B, T, E, C, W, normalize, dh, dW, db, ST(stream));
}
int agb_normalize_shapley(const float* pred, const float* grand, const float* null_v, int B, int T,
                          int C, float* out, void* stream) {
  return agb::normalize_shapley(pred, grand, null_v, B, T, C, out, ST(stream));
}
int agb_shapley_loss_fwd(const uint32_t* packed, int words, const float* v0, const float* v_s,
                         const float* phi, int B, int S, int n_players, int C, float* resid,
                         float* partial, float* loss, void* stream) {
  return agb::shapley_loss_fwd(packed, words, v0, v_s, phi, B, S, n_players, C, resid, partial, loss,
                               ST(stream));
}
int agb_shapley_loss_bwd(const uint32_t* packed, int words, const float* resid,
                         const float* grad_out, int B, int S, int n_players, int C, float* dphi,
                         void* stream) {
  return agb::shapley_loss_bwd(packed, words, resid, grad_out, B, S, n_players, C, dphi, ST(stream));
}

int agb_kernelshap_solve(const uint32_t* Z, int words, const double* weights, const double* probs,
                         const double* f_x, const double* f_null, int B, int S, int d, int C, int link_logit,
                         double* gram_ws, double* rhs_ws, double* phi, int* info, void* stream) {
  return agb::kernelshap_solve(Z, words, weights, probs, f_x, f_null, B, S, d, C, link_logit, gram_ws, rhs_ws, phi,
                               info, ST(stream));
}

}  // extern "C"

/*
 * autognothi_b200 — C-ABI of the B200 (sm_100a) coalition-masked evaluation hot path.
 *
 * The reference (gszfwsb/AutoGnothi) is pure Python/PyTorch and has no FFI; its plugin boundary is
 * the `ModelRecipe` dataclass (reference recipes/types.py:96-162) plus the free functions of
 * reference models/shapley.py.  This header is the boundary a native binding of that path would
 * use: every entry point takes plain device pointers, sizes and a cudaStream_t (passed as void*),
 * returns an int status (0 = ok) and never takes ownership of memory.  `agb_last_error()` returns a
 * thread-local description of the last failure.  Each declaration cites the reference code whose
 * arithmetic it replaces.
 *
 * Layout conventions
 *   - packed coalition masks: uint32 words, row-major [rows, words]; bit 0 of word 0 is the CLS
 *     token (always 1, reference recipes/vanilla_vit.py:219-224), bit j+1 is player j.
 *   - row order of per-coalition tensors is b * S + s (reference models/shapley.py:23-27).
 *   - dense matrices are row-major; bf16 = __nv_bfloat16 bit pattern (uint16).
 */
#ifndef AUTOGNOTHI_B200_H_
#define AUTOGNOTHI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGB_OK 0
#define AGB_ERR_INVALID 1
#define AGB_ERR_CUDA 2
#define AGB_ERR_UNSUPPORTED 3

#define AGB_MASK_MUL0 0   /* ViT: masked logit := 0 (reference models/vanilla_vit.py:446-454)   */
#define AGB_MASK_NEGINF 1 /* BERT: additive finfo.min (reference models/vanilla_bert.py:521-523) */

#define AGB_ACT_NONE 0
#define AGB_ACT_GELU 1 /* exact erf GELU, nn.GELU() default (reference models/vanilla_vit.py:488) */

/* ---- diagnostics ---------------------------------------------------------------------------- */
const char* agb_last_error(void);
int agb_version(void);

/* ---- dense GEMM: C = act(alpha * A * B^T + bias) + residual --------------------------------- */
/* Replaces every nn.Linear on the path (reference models/vanilla_vit.py:437-441,473-479,487-493,
 * 506-513; models/vanilla_bert.py:503-537,556-604).  tcgen05/TMA kernel, bf16 operands, fp32
 * accumulation.  Operand majorness: 0 = K-major (stored [rows, K]), 1 = MN-major (stored [K, rows]).
 * residual row remap r = (m / res_group) * res_rows + m % res_rows when res_group > 0. */
int agb_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major,
                  int M, int N, int K, float alpha, const float* bias, int act,
                  const void* residual_bf16, const float* residual_f32, int ldr, int res_group,
                  int res_rows, void* out, int ldo, int out_is_f32, void* stream);

/* LayerNorm-folded GEMM chain for pre-LN blocks (reference models/vanilla_vit.py:364-377: x + Attn(LN1(x)), then
 * y + W2 GELU(W1 LN2(y))).  Instead of a LayerNorm kernel between the GEMMs,
 *   - a residual GEMM (residual_f32 != NULL, fp32 out) can ALSO emit a bf16 copy of its output and per-row partial
 *     statistics: stats_out [M][agb_gemm_stats_parts(N)][2] = (sum, sum of squares) over each 128-column slab;
 *   - the consuming GEMM (bf16 out, no residual) takes that copy as A together with ln_stats / ln_parts, weights
 *     pre-scaled by gamma (B = W * gamma), bias' = b + W beta and ln_colsum[j] = sum_k B[j][k], and applies
 *     rstd_i * (acc_ij - mean_i * ln_colsum[j]) + bias'_j in its epilogue  ==  Linear(LayerNorm(x)).
 * All operands K-major ([rows, K]); returns AGB_ERR_UNSUPPORTED for shapes the tcgen05 pair kernel does not cover. */
int agb_gemm_bf16_fused(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                        int act, const float* residual_f32, int ldr, void* out, int ldo, int out_is_f32,
                        const float* ln_stats, int ln_parts, const float* ln_colsum, float ln_eps,
                        void* out_bf16_copy, int ldo_copy, float* stats_out, void* stream);
int agb_gemm_stats_parts(int N);
/* Residual GEMM of the same chain on a hi/lo residual stream (round 2; the two residual adds of reference
 * models/vanilla_vit.py:369-376).  The stream x is kept as two bf16 planes with x = hi + lo (hi = bf16(x), lo = bf16(x - hi):
 * 16 significant bits); the call updates both planes IN PLACE,  x <- x + A B^T + bias,  and writes the per-row partial
 * statistics of the new x exactly as agb_gemm_bf16_fused does.  The hi plane doubles as the bf16 copy the consuming
 * LayerNorm-folded GEMM takes as A, so no separate copy is written: 10 instead of 12 bytes of HBM traffic per element for
 * the HBM-bound attention output projection.  N % 256 == 0, ldx % 8 == 0; AGB_ERR_UNSUPPORTED otherwise. */
int agb_gemm_bf16_hilo(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, void* x_hi,
                       void* x_lo, int ldx, float* stats_out, void* stream);
/* Training: out = residual + Dropout(A B^T + bias) in ONE kernel (reference models/vanilla_vit.py:501-503, 512-516: dense ->
 * nn.Dropout -> residual add).  The keep mask is the counter-hash stream of agb_dropout(seed, tag) over the element index of the
 * contiguous (M, N) output, so the adjoint (agb_dropout on the gradient) regenerates it.  0 < thr16 < 65536; M * N < 2^32;
 * AGB_ERR_UNSUPPORTED when the tcgen05 pair kernel does not cover the shape (callers then run agb_gemm_bf16 + agb_dropout). */
int agb_gemm_bf16_dropout_residual(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                                   const float* residual_f32, int ldr, float* out, int thr16, uint64_t seed, int tag,
                                   void* stream);
/* Training, MLP of a block (reference models/vanilla_vit.py:487-493: dense -> nn.GELU): z = A B^T + bias (bf16, kept for the
 * adjoint) and gelu_out = GELU(z) (bf16, the next GEMM's operand) from ONE pass over the accumulator; both (M, N) with pitch ldo.
 * GELU is evaluated at the rounded z, as a separate agb_gelu_fwd on the stored z would. */
int agb_gemm_bf16_gelu_dual(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias, void* z_out,
                            void* gelu_out, int ldo, void* stream);
/* Its adjoint in the dgrad GEMM's epilogue: dz = (dY W) * GELU'(z), dY (M, K) bf16, W the forward weight stored [K, N]
 * (out_features x in_features), z / dz (M, N) bf16.  Replaces agb_gemm_bf16 + agb_gelu_bwd (the derivative is that of the
 * tanh-form approximant the bf16 kernels use, |error| <= 9e-4). */
int agb_gemm_bf16_gelu_bwd(const void* dY, int ldy, const void* W, int ldw, int M, int N, int K, const void* z, int ldz,
                           void* dz, int ldo, void* stream);
/* fp32 -> hi/lo planes (n % 8 == 0). */
int agb_split_hilo(const float* x, long long n, void* hi, void* lo, void* stream);
/* bf16 copy + one (sum, sum of squares) pair per row of an fp32 matrix: the entry of the chain above. */
int agb_rowstats_cast(const float* x, long long ldx, int rows, int H, void* out_bf16, long long ldo, float* stats,
                      void* stream);

/* Kernel selection for agb_gemm_bf16 (diagnostics / benchmarking): 0 = automatic, 1 = first-generation
 * 1-CTA kernel only, 2 = TMA-epilogue kernel with one CTA per tile, 3 = TMA-epilogue kernel on CTA pairs
 * (tcgen05 cta_group::2).  Returns the previous setting. */
int agb_gemm_set_variant(int variant);


/* fp32 CUDA-core GEMM with the same epilogue — the "exact" verification mode (fp32 rtol 1e-4 gate
 * of BASELINE.json).  A [M,K], B [N,K], C [M,N] row-major fp32. */
int agb_gemm_f32(const float* A, int lda, int a_mn_major, const float* B, int ldb, int b_mn_major, int M, int N,
                 int K, float alpha, const float* bias, int act, const float* residual, int ldr, float* C,
                 int ldc, void* stream);

/* ---- coalition masks (reference models/shapley.py:56-79,109-115,131-135; recipes/vanilla_vit.py:
 *      219-224) ------------------------------------------------------------------------------- */
/* (rows, n_players) int64 {0,1} -> packed words; prepend_cls=1 adds the always-on CLS bit 0. */
int agb_pack_masks_i64(const int64_t* mask, int rows, int n_players, int prepend_cls,
                       uint32_t* packed, int words, void* stream);
/* packed -> (rows, n) int64; `skip` leading tokens are dropped (1 removes the CLS bit). */
int agb_unpack_masks_i64(const uint32_t* packed, int rows, int n, int skip, int words, int64_t* out,
                         void* stream);
/* Paired Shapley-kernel sampler, `pairs` = n_mask_samples / 2 (reference models/shapley.py:56-79).
 * use_philox=0: u_players (pairs, n) and u_size (pairs) are the torch.rand draws of l.69 and l.133
 * — output is then bit-identical to mask_shapley_new under the same generator state.
 * use_philox=1: uniforms come from Philox4x32-10(seed, offset) on the device.
 * prefix = cumsum(p) - p (n-1 floats, l.132).  Row 2i+1 is the complement of row 2i (l.75-78).
 * dense (2*pairs, n) int64 is optional (the reference's return layout). */
int agb_shapley_masks(const float* u_players, const float* u_size, const float* prefix,
                      int use_philox, uint64_t seed, uint64_t offset, int pairs, int n_players,
                      uint32_t* packed, int words, int64_t* dense, void* stream);
/* mask_purely_uniform (reference models/shapley.py:109-115). */
int agb_uniform_masks(const float* u_players, const float* u_row, int use_philox, uint64_t seed,
                      uint64_t offset, int rows, int n_players, uint32_t* packed, int words,
                      int64_t* dense, void* stream);

/* Rank masks: rows of scores (rows, n) fp32 -> packed masks (rows * nstops, words), row r*nstops+i =
 * `mask_base` with the stops[i] highest-ranked players flipped (rank = descending score, ties: larger index first,
 * i.e. np.argsort(a)[::-1] with a stable sort).  Replaces the python loops of scripts/measure_faithfulness.py:225-251
 * (_get_perturbed_samples: scores = attributions, stops = linspace) and models/shapley.py:118-128
 * (mask_uniform_selective: scores = random keys, one stop = n_masked, base 1; use_philox draws the keys on device).
 * stops: device int32 [nstops]; dense (optional): int64 (rows * nstops, n), the reference's return layout. */
int agb_rank_masks(const float* scores, int use_philox, uint64_t seed, uint64_t offset, int rows, int n_players,
                   const int* stops, int nstops, int mask_base, uint32_t* packed, int words, int64_t* dense,
                   void* stream);

/* Masked-token dropping for additive (-inf) masks (reference models/vanilla_bert.py:520-523): a masked
 * token is never attended to and the surrogate head reads only token 0, so masked tokens cannot influence the output and
 * each row's kept tokens may be packed back to back (exact).  counts[r] = kept tokens of row r among the first T bits;
 * src[cu[r] + k] = (r / S) * T + position of the k-th kept token (cu = exclusive prefix sum of counts, rows + 1 ints). */
int agb_mask_counts(const uint32_t* packed, int rows, int words, int T, int* counts, void* stream);
int agb_packed_token_index(const uint32_t* packed, int rows, int words, int T, int S, const int* cu, int64_t* src,
                           void* stream);
/* Attention over packed variable-length rows: qkv (total_tokens, 3H) bf16, row r = tokens [cu[r], cu[r+1]) (every packed
 * token is a live key), max_len >= every row length (<= 512); ctx (total_tokens, H).  tcgen05 kernel for head dim 64;
 * head dims 8 / 16 / 32 (the LTT side ladder, reference models/ltt_bert.py:437-451) run the narrow-head kernel. */
int agb_attention_bf16_varlen(const void* qkv, const int* cu, int rows, int max_len, int total_tokens, int H, int heads,
                              void* ctx, void* stream);
/* ViT-masked attention (reference models/vanilla_vit.py:444-459: masked logit := 0) on rows whose tokens were permuted so
 * that the kept ones come first: keys [0, nkeep[row]) keep their logits, the masked keys [nkeep[row], T) all have the
 * logit 0 and are folded into ONE virtual key of weight (T - nkeep) whose value row is the mean of theirs — the same
 * softmax, with fewer columns to exponentiate.  qkv (rows, T, 3H) bf16 -> ctx (rows, T, H); head dim 64, T <= 256. */
int agb_attention_bf16_prefix(const void* qkv, const int* nkeep, int rows, int T, int H, int heads, void* ctx, void* stream);

/* ---- row kernels ---------------------------------------------------------------------------- */
/* nn.LayerNorm over the last dim (reference models/vanilla_vit.py:94,213,369,373; vanilla_bert.py:
 * 318,548,596).  x fp32 or bf16 [rows, in_stride]; writes bf16 and/or fp32 [rows, out_stride]. */
int agb_layernorm(const void* x, int x_is_bf16, long long in_stride, int rows, int H,
                  const float* gamma, const float* beta, float eps, void* out_bf16, float* out_f32,
                  long long out_stride, void* stream);
int agb_cast_f32_to_bf16(const float* in, void* out_bf16, long long n, void* stream);
/* stride-P patchify of NCHW fp32 images into Conv2d-weight order (reference models/vanilla_vit.py:
 * 279-284): (B,C,px,px) -> (B*(px/P)^2, C*P*P) fp32 or bf16. */
int agb_vit_im2col(const float* images, int B, int C, int px, int P, void* out, int out_is_bf16,
                   void* stream);
/* cat(cls, patches) + position embeddings, broadcast to S coalition rows per image (reference
 * models/vanilla_vit.py:242-253; replaces Xs_EXT of scripts/train_explainer.py:159-163). */
int agb_vit_assemble(const float* patch_emb, const float* cls_token, const float* pos_emb, int B,
                     int S, int T, int H, float* x, void* stream);
/* S consecutive copies of every source row: dst[(b*S + s)] = src[b], rows of row_bytes (multiple of 16) bytes.
 * The embedded input of the first-block sharing path fanned out to its S coalition rows on the device (replaces the
 * Xs_EXT replication loop of reference scripts/train_explainer.py:159-163 at the residual-stream level). */
int agb_repeat_rows(const void* src, int B, long long row_bytes, int S, void* dst, void* stream);
/* Kept-first token order of ViT coalition rows (surrogate evaluation; the mask semantics of reference
 * models/vanilla_vit.py:449-450 make every masked key's logit 0, so the masked keys of a row are interchangeable): a stable
 * partition of each row's tokens into kept (CLS = bit 0 always) and masked.  packed (rows, words) -> order (rows, T) uint8
 * [position -> token], pos (rows, T) uint8 [token -> position], nkeep (rows), prefix (rows, words) packed mask of the
 * permuted row (bits [0, nkeep) set).  T <= 256. */
int agb_kept_first_order(const uint32_t* packed, int rows, int words, int T, uint8_t* order, uint8_t* pos, int* nkeep,
                         uint32_t* prefix, void* stream);
/* dst[(r, q), :] = src[(r / S, order[r, q]), :] for rows of row_bytes bytes (multiple of 16): per-coalition residual stream
 * in kept-first order from the per-input embeddings (replaces the reference's Xs_EXT replication,
 * scripts/train_explainer.py:159-163, on that path). */
int agb_gather_token_rows(const void* src, const uint8_t* order, int rows, int T, int S, long long row_bytes, void* dst,
                          void* stream);
/* agb_gather_token_rows from fp32 rows of H values into hi/lo planes (the residual stream of agb_gemm_bf16_hilo). */
int agb_gather_token_rows_hilo(const float* src, const uint8_t* order, int rows, int T, int S, int H, void* hi, void* lo,
                               void* stream);
/* agb_masked_attention_bf16_shared whose query token t of row r is written to token position dst_pos[r, t] of ctx
 * (rows, T, H): the first block of the kept-first evaluation order.  T <= 208, head dim 64, ViT mask semantics. */
int agb_masked_attention_bf16_scatter(const void* qkv, const uint32_t* mask, int words, int rows, int share, int T, int H,
                                      int heads, const uint8_t* dst_pos, void* ctx, void* stream);
/* word + token_type(0) + position embeddings -> LayerNorm, broadcast to S rows per input
 * (reference models/vanilla_bert.py:307-325). ids (B,T) int64. */
int agb_bert_embed(const int64_t* ids, const float* word, const float* pos, const float* type0,
                   const float* gamma, const float* beta, float eps, int B, int S, int T, int H,
                   int vocab, float* x, void* stream);
/* CLS head -> class probabilities.  mode 0 (ViT): softmax(Wc LN(x[row,0]) + bc) (reference
 * models/vanilla_vit.py:213,52-56); mode 1 (BERT): softmax(Wc tanh(Wp x[row,0] + bp) + bc)
 * (reference models/vanilla_bert.py:73-77,615-619).  x fp32, row_stride = T*H. */
int agb_cls_head(const float* x, long long row_stride, int rows, int H, int C, int mode,
                 const float* ln_gamma, const float* ln_beta, float eps, const float* w_pool,
                 const float* b_pool, const float* w_cls, const float* b_cls, float* probs,
                 float* logits_or_null, void* stream);

/* ---- key-masked attention ------------------------------------------------------------------- */
/* qkv (rows, T, 3H) fused q|k|v; packed key mask (rows, words); ctx (rows, T, H).
 * mode AGB_MASK_MUL0 / AGB_MASK_NEGINF.  Scores are never written to HBM. */
int agb_masked_attention_simt(const void* qkv, int io_is_bf16, const uint32_t* mask, int words,
                              int rows, int T, int H, int heads, int mode, void* ctx, void* stream);
/* tcgen05/TMA version: bf16 in/out, head dim 64, T <= 512 (two softmax groups up to 256 keys, one group beyond). */
int agb_masked_attention_bf16(const void* qkv, const uint32_t* mask, int words, int rows, int T,
                              int H, int heads, int mode, void* ctx, void* stream);

/* First-block form: `share` consecutive mask rows (the coalitions of one input) read the SAME projections, qkv
 * (rows / share, T, 3H) — before the first attention every coalition of an input holds identical activations, so its
 * LayerNorm + QKV projection is computed once per input (exact work-skipping).  ctx (rows, T, H) as above. */
int agb_masked_attention_bf16_shared(const void* qkv, const uint32_t* mask, int words, int rows, int share, int T,
                                     int H, int heads, int mode, void* ctx, void* stream);

/* CLS-query attention for the LAST encoder block of a surrogate / classifier (exact work-skipping: the heads read only
 * token 0 — reference models/vanilla_vit.py:51-56, models/vanilla_bert.py:615-619 — so only the CLS query row of the last
 * block is needed; keys / values still come from all T tokens).  q (rows, ldq); kv (rows*T, ldkv) with keys at column
 * k_off + head*d and values at v_off + head*d; ctx (rows, ldc).  fp32 math, fp32 or bf16 I/O, same mask semantics. */
int agb_cls_attention(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off, int io_is_bf16,
                      const uint32_t* mask, int words, int rows, int T, int H, int heads, int mode, void* ctx,
                      long long ldc, void* stream);
/* Same over packed variable-length rows (every packed token is a live key): kv rows [cu[r], cu[r+1]) belong to row r. */
int agb_cls_attention_varlen(const void* q, long long ldq, const void* kv, long long ldkv, int k_off, int v_off,
                             int io_is_bf16, const int* cu, int rows, int max_len, int H, int heads, void* ctx,
                             long long ldc, void* stream);

/* Kernel selection for agb_masked_attention_bf16 (diagnostics): 0 = automatic (split-softmax kernel for ViT masks at
 * T <= 208 in token order, else the pipelined one), 1 = first-generation kernel, 2 = pipelined kernel, 3 = split-softmax
 * kernel for the kept-first order too.  Returns the previous setting. */
int agb_attention_set_variant(int variant);
/* Same for agb_masked_attention_bwd: 0 = automatic (tcgen05 kernel for bf16), 1 = CUDA-core kernel. */
int agb_attention_bwd_set_variant(int variant);
/* Diagnostics: device buffer of 64 x 8 int64 receiving clock64() pipeline timestamps of CTA 0 (NULL = off). */
int agb_attention_set_trace(void* trace);

/* ---- explainer head + efficiency normalisation (reference models/vanilla_vit.py:123-129,
 *      models/shapley.py:82-93) --------------------------------------------------------------- */
/* h (B*T, E) fp32|bf16, W (C,E), bias (C), grand (B,C), null (C) -> phi (B,C,T-1); pred (B,T,C)
 * optional.  The divisor is T (CLS included), CLS row dropped afterwards. */
int agb_explainer_head_fwd(const void* h, int h_is_bf16, int B, int T, int E, int C, const float* W,
                           const float* bias, const float* grand, const float* null_v,
                           int normalize, float* phi, float* pred_or_null, void* stream);
/* adjoint: dh (B*T,E) same dtype as h (nullable), dW (C,E) and db (C) ACCUMULATED (nullable). */
int agb_explainer_head_bwd(const float* dphi, const void* h, int h_is_bf16, int B, int T, int E,
                           int C, const float* W, int normalize, void* dh, float* dW, float* db,
                           void* stream);
/* normalize_shapley_explanation on a materialised pred (B,T,C) (reference models/shapley.py:82-93) */
int agb_normalize_shapley(const float* pred, const float* grand, const float* null_v, int B, int T,
                          int C, float* out, void* stream);

/* ---- Shapley loss (reference models/shapley.py:9-53) ---------------------------------------- */
/* packed (B*S, words), v0 (C), v_s (B*S,C), phi (B,C,n) -> resid (B*S,C), partial (B) scratch,
 * loss (1) = n * mean((v0 + mask.phi - v_s)^2); deterministic reduction order. */
int agb_shapley_loss_fwd(const uint32_t* packed, int words, const float* v0, const float* v_s,
                         const float* phi, int B, int S, int n_players, int C, float* resid,
                         float* partial, float* loss, void* stream);
/* dphi (B,C,n) = grad_out * dloss/dphi; grad_out (1) device scalar or NULL for 1. */
int agb_shapley_loss_bwd(const uint32_t* packed, int words, const float* resid,
                         const float* grad_out, int B, int S, int n_players, int C, float* dphi,
                         void* stream);

/* ---- explainer training: adjoints (the reference relies on torch autograd over models/vanilla_vit.py:
 *      102-130 / models/vanilla_bert.py:123-162 inside scripts/train_explainer.py:182-198) ------------- */
/* exact-erf GELU (nn.GELU(), reference models/vanilla_vit.py:488) forward and backward, fp32 or bf16 */
int agb_gelu_fwd(const void* z, void* out, long long n, int is_bf16, void* stream);
int agb_gelu_bwd(const void* dy, const void* z, void* dz, long long n, int is_bf16, void* stream);
/* bias gradient of an nn.Linear: out[n] += sum_m Y[m,n] */
int agb_colsum(const void* y, int is_bf16, long long ld, int M, int N, float* out, void* stream);
/* nn.LayerNorm adjoint: dx = dres + dLN(dy); dgamma/dbeta ACCUMULATED (nullable) */
int agb_layernorm_bwd(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, const float* gamma,
                      const float* dres, int rows, int H, float eps, float* dx, float* dgamma, float* dbeta,
                      void* stream);
/* adjoint of the key-masked attention (reference models/vanilla_vit.py:436-465, vanilla_bert.py:503-537):
 * qkv (rows,T,3H), dctx (rows,T,H) -> dqkv (rows,T,3H); head dim 64, T <= 512 (tcgen05 kernel up to 256, two-pass CUDA-core kernels beyond); a ViT-masked key passes no
 * gradient to Q/K (its logit is the constant 0) but its V row still receives P^T dO.  Head dims 8 / 16 / 32 (the side
 * ladders of reference models/ltt_vit.py:386-396, ltt_bert.py:437-451) run a CUDA-core kernel with fp32-staged operands. */
int agb_masked_attention_bwd(const void* qkv, const void* dctx, int io_is_bf16, const uint32_t* mask, int words,
                             int rows, int T, int H, int heads, int mode, void* dqkv, void* stream);
/* nn.Dropout in training mode (reference models/vanilla_vit.py:253,501-503,512-516; models/vanilla_bert.py:325,559,603):
 * out[e] = residual[e] + keep_e * y[e] / (1 - p), n elements, y fp32|bf16, residual fp32 (nullable), out fp32|bf16.
 * keep_e comes from a counter hash of (seed, tag, e) compared with thr16 = round(p * 65536): the adjoint is the same
 * call on the incoming gradient with the same (seed, tag), so no mask is stored. */
int agb_dropout(const void* y, int y_is_bf16, const float* residual, void* out, int out_is_bf16, long long n,
                int thr16, uint64_t seed, int tag, void* stream);
/* Key-masked attention with dropout on the probabilities (training mode of reference models/vanilla_vit.py:454-459,
 * models/vanilla_bert.py:527-532): ctx = (softmax(scores) o M / (1 - p)) V, M from the same counter hash keyed by
 * (seed, row * heads + head, query, key).  bf16: head dim 64 with T <= 256 (tcgen05 kernels) or head dims 8/16/32;
 * fp32 (the exact mode): CUDA-core kernels, adjoint up to T = 200 at head dim 64.
 * _bwd is its adjoint (regenerates M).  agb_attention_dropout_mask writes M as (rows, heads, T, T) bytes (tests). */
int agb_masked_attention_dropout_fwd(const void* qkv, int io_is_bf16, const uint32_t* mask, int words, int rows, int T, int H,
                                     int heads, int mode, void* ctx, int thr16, uint64_t seed, void* stream);
int agb_masked_attention_dropout_bwd(const void* qkv, const void* dctx, int io_is_bf16, const uint32_t* mask, int words,
                                     int rows, int T, int H, int heads, int mode, void* dqkv, int thr16, uint64_t seed,
                                     void* stream);
int agb_attention_dropout_mask(void* keep, int rows, int heads, int T, int thr16, uint64_t seed,
                               void* stream);
/* ViT embedding adjoint (reference models/vanilla_vit.py:242-253): dpos/dcls ACCUMULATED, dpatch (B*(T-1),H) */
int agb_vit_embed_bwd(const float* dx, int B, int T, int H, float* dpos, float* dcls, void* dpatch,
                      int dpatch_is_bf16, void* stream);
/* BERT embedding pre-LayerNorm sum and its scatter adjoint (reference models/vanilla_bert.py:307-325;
 * the padding_idx row of word_embeddings receives no gradient) */
int agb_bert_embed_sum(const int64_t* ids, const float* word, const float* pos, const float* type0, int BT, int T,
                       int H, int vocab, float* out, void* stream);
int agb_bert_embed_scatter(const int64_t* ids, const float* dsum, int BT, int T, int H, int vocab, int pad_id,
                           float* dword, float* dpos, float* dtype0, void* stream);

/* ---- KernelSHAP weighted least squares (reference models/kernel_shap_bert.py:170-185 hands this to the
 *      third-party shap.KernelExplainer; requirements.txt:10 pins shap~=0.44.1) ------------------------- */
/* Batched over B explained samples, float64 throughout.  Z (B,S,words) packed coalitions over the d = T token
 * features (bit j = feature j, no CLS offset); weights (B,S) kernel weights; probs (B,S,C) background-averaged
 * model outputs; f_x (B,C) outputs on the unperturbed rows; f_null (C) background expectation; link_logit = 1
 * applies shap's link="logit".  gram_ws (B,(d-1)^2) and rhs_ws (B,(d-1),C) are caller-provided scratch.
 * phi (B,C,d) satisfies sum_j phi = link(f_x) - link(f_null) exactly (efficiency constraint).
 * info[b] = 0, or k+1 if the Gram matrix lost positive definiteness at column k (LAPACK potrf convention).
 * shap's optional l1_reg feature pre-selection is not applied. */
int agb_kernelshap_solve(const uint32_t* Z, int words, const double* weights, const double* probs,
                         const double* f_x, const double* f_null, int B, int S, int d, int C, int link_logit,
                         double* gram_ws, double* rhs_ws, double* phi, int* info, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUTOGNOTHI_B200_H_ */

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

#ifdef __cplusplus
}
#endif
#endif /* AUTOGNOTHI_B200_H_ */

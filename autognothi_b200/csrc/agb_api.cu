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
}  // namespace agb

using agb::bf16;

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

}  // extern "C"

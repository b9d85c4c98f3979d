#!/bin/bash
# Correctness + timing sweep of the tcgen05 GEMM variants on a B200 (test infrastructure; run under gpurun).
#   variant 1 = first-generation kernel, 2 = TMA epilogue one CTA per tile, 3 = CTA pairs (cta_group::2)
cd "$(dirname "$0")/../.." || exit 1
G=build/gemm_check
fail=0
run() { timeout 120 $G "$@" || { echo "FAILED: $*"; fail=1; }; }
for v in 2 3 1; do
  # small / ragged shapes (tails in M, N, K; MN-major operands for dgrad / wgrad)
  run 300 256 64 0 0 0 1 0 0 0 $v
  run 1000 768 768 0 0 0 1 0 1 0 $v
  run 1000 768 768 0 0 0 1 2 1 0 $v
  run 1000 768 768 0 0 0 1 3 1 0 $v
  run 999 2304 776 0 0 1 1 0 0 0 $v
  run 777 328 200 0 0 0 1 2 1 0 $v
  run 6304 768 3072 0 0 0 1 2 1 0 $v
  run 640 768 512 0 1 0 0 0 1 0 $v
  run 768 3072 640 1 1 0 0 0 1 0 $v
  run 520 520 264 1 0 0 1 0 0 0 $v
  run 1000 768 768 0 0 1 1 0 1 0 $v
done
# the four ViT-B shapes at 512 and 1024 evals (M = evals * 197), timed
for M in 100864; do
  for v in 1 2 3; do
    run $M 2304 768 0 0 0 1 0 0 10 $v
    run $M 768 768 0 0 0 1 3 1 10 $v
    run $M 3072 768 0 0 1 1 0 0 10 $v
    run $M 768 3072 0 0 0 1 3 1 10 $v
  done
done
exit $fail

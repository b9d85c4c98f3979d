#!/bin/bash
# ncu --set full capture of one launch per ViT-B GEMM shape (test infrastructure; run under gpurun).
#   usage: ncu_gemm_shapes.sh <variant> <out-prefix>
cd "$(dirname "$0")/../.." || exit 1
V=${1:-3}; OUT=${2:-gpurun_out/gemm_v$V}
M=100864
i=0
for cfg in "2304 768 0 0 0 1 0 0" "768 768 0 0 0 1 3 1" "3072 768 0 0 1 1 0 0" "768 3072 0 0 0 1 3 1"; do
  ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -f -o ${OUT}_$i \
      build/gemm_check $M $cfg 3 $V > ${OUT}_$i.log 2>&1
  i=$((i+1))
done

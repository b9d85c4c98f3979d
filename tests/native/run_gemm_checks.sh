#!/bin/bash
# Runs the native tcgen05 GEMM checks, one process per case so a trapped kernel cannot poison the
# next case, each under its own timeout.  Usage: tests/native/run_gemm_checks.sh [logfile]
LOG=${1:-gpurun_out/gemm_checks.log}
mkdir -p "$(dirname "$LOG")"
: > "$LOG"
BIN=build/gemm_check
run() { echo "## $*" >> "$LOG"; timeout 90 $BIN "$@" >> "$LOG" 2>&1; echo "rc=$?" >> "$LOG"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> "$LOG" 2>&1
#    M     N    K  a_mn b_mn act bias res f32 iters
run 128   256   64   0 0 0 0 0 1
run 128   256  256   0 0 0 0 0 1
run 256   512  512   0 0 0 1 0 1
run 197   192  192   0 0 1 1 1 0
run 1000  128  320   0 0 0 1 0 0
run 128   256  200   0 0 0 0 0 1
run 300   264  136   0 0 1 1 1 1
run 256   256  128   1 0 0 0 0 1
run 256   256  128   0 1 0 0 0 1
run 256   256  128   1 1 0 0 0 1
run 336   320  200   1 1 0 1 0 1
run 336   128  200   1 1 0 1 0 0
run 25216  768  768  0 0 0 1 1 0 20
run 25216 2304  768  0 0 0 1 0 0 20
run 25216 3072  768  0 0 1 1 0 0 20
run 25216  768 3072  0 0 0 1 1 0 20
run 50432 3072 768  0 0 1 1 0 0 5
run 50432 768 3072  0 0 0 1 1 0 5
run 8192  8192 8192  0 0 0 0 0 0 5
run 25216  768  768  0 1 0 0 0 0 20
run 768   3072 25216 1 1 0 0 0 1 20
grep -cE "PASS" "$LOG" | sed 's/^/passes: /' >> "$LOG"
tail -n 80 "$LOG"

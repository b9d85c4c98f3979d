# A/B of the hi/lo residual stream, its ring configuration and the L2 prefetch distance of the residual chunks (bench.py, one box)
mkdir -p gpurun_out
out=gpurun_out/r02_hilo_ab2.txt
: > $out
for cfg in "1 0 0" "1 0 2" "1 0 4" "1 0 8" "1 1 0" "1 1 2" "1 1 4" "1 1 8" "0 0 0" "0 0 4" "0 0 8" "0 0 16" "1 0 0"; do
  set -- $cfg
  AGB_HILO_RESIDUAL=$1 AGB_GEMM_HILO_CFG=$2 AGB_GEMM_RES_PREFETCH=$3 python bench.py --no-ltt --no-cpu-baseline --no-train --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('hilo $1 cfg $2 prefetch $3 evals/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], [(s['gemm'].split(' ',1)[1], s['avg_us']) for s in d['roofline']['by_shape']])
" >> $out
done
cat $out

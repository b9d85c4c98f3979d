# A/B of the explainer-step epilogue fusions on ONE box (bench.py training leg, 64 images per step, dropout on)
mkdir -p gpurun_out
out=gpurun_out/r02_train_fusion_ab.txt
: > $out
for cfg in "1 1" "0 0" "1 1" "0 0" "1 0" "0 1" "1 1" "0 0"; do
  set -- $cfg
  AGB_FUSE_GELU=$1 AGB_FUSE_DROPOUT=$2 python bench.py --no-ltt --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
t = d['train']
print('fuse_gelu $1 fuse_dropout $2 train samples/s', round(t['value'],1), 'ms', round(t.get('ms_per_step',0),2), 'evals/s', round(d['value']), 'clk', d['clocks']['sm_mhz'])
" >> $out
done
cat $out

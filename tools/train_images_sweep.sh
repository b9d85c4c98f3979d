mkdir -p gpurun_out
: > gpurun_out/r02_train_images_sweep3.txt
for n in 64 63 62 61 60 58 56; do
  python bench.py --no-ltt --no-cpu-baseline --train-images $n --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
t = d['train']
print('train-images', $n, 'samples/s', round(t['value'],1), 'ms', round(t.get('ms_per_step',0),2), 'evals/s', round(d['value']), 'clk', d['clocks']['sm_mhz'])
" >> gpurun_out/r02_train_images_sweep3.txt
done
cat gpurun_out/r02_train_images_sweep3.txt

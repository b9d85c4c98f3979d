# End-of-round records on one fresh B200 box (run under gpurun): full GPU test suite, smoke(), the driver's bench command for both
# arms, and the other BASELINE.json configurations.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_full_end2.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_full_end2.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_end2.log 2>&1; tail -4 gpurun_out/r02_smoke_end2.log
echo "tests+smoke: $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference > gpurun_out/r02_bench_end2_ref.json 2> gpurun_out/r02_bench_end2_ref.err
echo "reference arm: $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_end2.json 2> gpurun_out/r02_bench_end2.err
echo "our arm: $(( $(date +%s) - t0 )) s"; t0=$(date +%s)
timeout 900 python bench.py --model vit_large --steps 10 --warmup 3 > gpurun_out/r02_bench_end2_vit_large.json 2> gpurun_out/r02_bench_end2_vit_large.err
timeout 900 python bench.py --model vit_tiny --steps 20 --warmup 5 > gpurun_out/r02_bench_end2_vit_tiny.json 2> gpurun_out/r02_bench_end2_vit_tiny.err
echo "vit_large + vit_tiny: $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
for f in ("r02_bench_end2_ref", "r02_bench_end2", "r02_bench_end2_vit_large", "r02_bench_end2_vit_tiny"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        t = d.get("train") or {}
        print(f, round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "train", round(t.get("value", 0), 1), "frac", round((d.get("roofline") or {}).get("frac", 0), 3),
              "whole", round((d.get("whole_path") or {}).get("frac_of_burst_peak", 0), 3), "clk", (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", e)
PY

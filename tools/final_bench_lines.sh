# Bench lines of every BASELINE.json configuration on one fresh B200 box (run under gpurun), both arms for the metric's config.
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference > gpurun_out/r02_bench_end3_ref.json 2> gpurun_out/r02_bench_end3_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_end3.json 2> gpurun_out/r02_bench_end3.err
timeout 900 python bench.py --model vit_large --steps 10 --warmup 3 > gpurun_out/r02_bench_end3_vit_large.json 2> gpurun_out/r02_bench_end3_vit_large.err
timeout 900 python bench.py --model vit_tiny --steps 20 --warmup 5 > gpurun_out/r02_bench_end3_vit_tiny.json 2> gpurun_out/r02_bench_end3_vit_tiny.err
timeout 400 python bench.py --workload bert_base_tayp_vanilla --steps 20 --warmup 5 > gpurun_out/r02_bench_end3_bert.json 2> gpurun_out/r02_bench_end3_bert.err
timeout 400 python bench.py --workload bert_base_tayp_kernel_shap --steps 20 --warmup 5 > gpurun_out/r02_bench_end3_kshap.json 2> gpurun_out/r02_bench_end3_kshap.err
python - <<'PY'
import json
for f in ("r02_bench_end3_ref", "r02_bench_end3", "r02_bench_end3_vit_large", "r02_bench_end3_vit_tiny", "r02_bench_end3_bert", "r02_bench_end3_kshap"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        t = d.get("train") or {}
        r = d.get("roofline") or {}
        print(f, round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "train", round(t.get("value", 0), 1), "frac", round(r.get("frac", 0), 3),
              "same-box", r.get("same_box_cublas_bf16_tflops_sustained"), r.get("frac_of_same_box_cublas"),
              "whole", round((d.get("whole_path") or {}).get("frac_of_burst_peak", 0), 3), "clk", (d.get("clocks") or {}).get("sm_mhz"), "cpu", round((d.get("cpu_baseline") or {}).get("value", 0), 1))
    except Exception as e:
        print(f, "ERR", e)
PY
